"""Print CTA 0's pipeline timeline (ns since kernel entry) for one tcgen05 conv launch.
usage: python tools/trace_conv.py n h w cin cout [pool]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qnn_b200 as q
from qnn_b200 import _lib as L, kernels as K

n, h, w, cin, cout = (int(v) for v in sys.argv[1:6])
pool = len(sys.argv) > 6 and sys.argv[6] == "pool"
rng = np.random.default_rng(0)
first = cin == 3
if first:
    x = torch.from_numpy(rng.integers(0, 256, size=(n, h, w, cin), dtype=np.uint8)).cuda()
else:
    x = torch.from_numpy(rng.integers(-8, 8, size=(n, h, w, cin)).astype(np.int8)).cuda()
wp = K.pack_weights(torch.from_numpy(rng.uniform(-1, 1, size=(3, 3, cin, cout)).astype(np.float32)).cuda(), L.W_QUANT, 4, 1.0, L.WFMT_I8)
epi = K.make_epilogue(1.0 / 64, act=L.ACT_QUANT, abits=4, pool=2 if pool else 0)
xq = K.QTensor("u8", x, 1 / 255.0, cin) if first else K.QTensor("i8", x, 0.125, cin)
for _ in range(3):
    K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_AUTO)
torch.cuda.synchronize()
buf = torch.zeros(16 * 1024, dtype=torch.int64, device="cuda")
L.check(L.lib().qnnb_debug_set_trace(L.ptr(buf), buf.numel()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_AUTO)
e1.record()
torch.cuda.synchronize()
L.check(L.lib().qnnb_debug_set_trace(None, 0))
b = buf.cpu().numpy()
acct = b[15 * 1024:15 * 1024 + 8]
if acct[0] == 5:
    print("MMA thread cycles: wait tempty %d, wait hfull %d, wait afull %d, issue %d, total %d over %d tiles" % tuple(int(v) for v in acct[2:8]))
ev = []
for wp_ in range(15):
    reg = b[wp_ * 1024:(wp_ + 1) * 1024]
    for i in range(int(reg[0])):
        ev.append((int(reg[3 + 2 * i]), int(reg[2 + 2 * i]) >> 32, int(reg[2 + 2 * i]) & 0xffffffff, wp_))
ev.sort()
t0 = ev[0][0]
names = {20: "epi chunk0 loaded", 21: "epi chunk1 loaded", 22: "epi chunk2 loaded", 23: "epi chunk3 loaded", 11: "epi loop top", 12: "epi chunk0 loaded", 13: "epi chunk1 loaded", 14: "epi math done", 10: "producer halo staged", 1: "entry", 2: "setup done", 3: "producer tile ready", 4: "mma acc free", 5: "mma first operands", 6: "mma tile issued", 7: "epi acc ready", 8: "epi stored", 9: "teardown"}
print("kernel event time %.1f us, %d trace events" % (e0.elapsed_time(e1) * 1e3, len(ev)))
for t, tag, idx, wp_ in ev[:int(os.environ.get("TRACE_MAX", "400"))]:
    print("%8d ns  w%-2d %-20s tile %d" % (t - t0, wp_, names.get(tag, tag), idx))
