"""CTA 0's timeline (ns since kernel entry) of one first-layer (K5) launch; needs a TRACE build:
  make -C <pkg>/csrc VARIANT=_trace TRACE=1 && QNNB_LIB=<pkg>/libqnnb200_trace.so python tools/k5_trace.py n cout pool abits"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qnn_b200 as q
from qnn_b200 import _lib as L, kernels as K
n, cout, pool, abits = (int(v) for v in sys.argv[1:5])
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.integers(0, 256, size=(n, 32, 32, 3), dtype=np.uint8)).cuda()
wp = K.pack_weights(torch.from_numpy(rng.uniform(-1, 1, size=(3, 3, 3, cout)).astype(np.float32)).cuda(), L.W_QUANT, min(abits, 8), 1.0, L.WFMT_I8)
inv = torch.from_numpy((rng.uniform(0.3, 0.9, cout) * rng.choice([1, 1, 1, -1], cout)).astype(np.float32)).cuda()
sh = torch.from_numpy(rng.uniform(-0.2, 0.2, cout).astype(np.float32)).cuda()
bias = torch.from_numpy(rng.uniform(-0.1, 0.1, cout).astype(np.float32)).cuda()
epi = K.make_epilogue(K.acc_scale(1.0 / 255.0, 1.0 / (1 << (min(abits, 8) - 1))), bias=bias, bn_inv=inv, bn_shift=sh, act=L.ACT_QUANT, abits=abits, pool=2 if pool else 0)
xq = K.QTensor("u8", x, 1.0 / 255.0, 3)
for _ in range(3):
    K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_TCGEN05)
torch.cuda.synchronize()
buf = torch.zeros(21 * 1024, dtype=torch.int64, device="cuda")
L.check(L.lib().qnnb_debug_set_trace(L.ptr(buf), buf.numel()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_TCGEN05)
e1.record()
torch.cuda.synchronize()
L.check(L.lib().qnnb_debug_set_trace(None, 0))
b = buf.cpu().numpy()
names = {1: "entry", 2: "neg scan done", 3: "A + constants built", 4: "setup sync done", 5: "griddep_wait done", 6: "stage: halo staged", 7: "stage: halo buffer free", 8: "stage: raw rows arrived",
         10: "MMA: halo ready", 11: "MMA: acc free", 12: "MMA: committed", 20: "epi: loop top", 21: "epi: acc ready", 22: "epi: TMEM loaded",
         23: "epi: next halo staged", 24: "epi: math done", 25: "epi: fence + store wait", 26: "epi: row-quarter barrier", 30: "teardown", 31: "exit"}
ev = []
for w in range(20):
    reg = b[w * 1024:(w + 1) * 1024]
    for i in range(int(reg[0])):
        ev.append((int(reg[3 + 2 * i]), int(reg[2 + 2 * i]) >> 32, int(reg[2 + 2 * i]) & 0xffffffff, w))
ev.sort()
t0 = ev[0][0]
show = set(int(v) for v in os.environ.get("TRACE_WARPS", "0,1,16").split(","))
print("kernel event time %.1f us, %d trace events" % (e0.elapsed_time(e1) * 1e3, len(ev)))
for t, tag, idx, w in ev:
    if w in show and (idx < int(os.environ.get("TRACE_TILES", "6")) or tag >= 30):
        print("%8d ns  w%-2d %-26s %d" % ((t - t0) / 1.965, w, names.get(tag, tag), idx))
