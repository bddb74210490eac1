"""CTA 0's timeline (ns since kernel entry) of one whole-network launch (csrc/net_fused.cu); needs a TRACE build:
  make -C <pkg>/csrc VARIANT=_trace TRACE=1 && QNNB_LIB=<pkg>/libqnnb200_trace.so python tools/net_trace.py [n] [cfg1|cifar]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qnn_b200 as q
from qnn_b200 import _lib as L
from helpers import make_cf, CONFIGS, assign_weights_from_spec
from oracle import netspec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
which = sys.argv[2] if len(sys.argv) > 2 else "cfg1"
kw = CONFIGS["cfg1"] if which == "cfg1" else dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG', nla=1, nfa=64, nlb=1, nfb=64, nlc=1, nfc=64)
cf = make_cf(**kw)
model = q.build_model(cf)
nodes = netspec.build_spec(cf)
netspec.set_weights(nodes, netspec.random_weights(nodes, seed=7, bias_range=0.1, bn="spread"))
assign_weights_from_spec(model, nodes)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(n, cf.dim, cf.dim, cf.channels), dtype=np.uint8)).cuda()
plan = model.plan()
assert plan.fused_available(x)
for _ in range(3):
    plan.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    plan.forward(x)
e1.record()
torch.cuda.synchronize()
print("eager back-to-back: %.2f us per launch" % (e0.elapsed_time(e1) * 1e3 / 20))
buf = torch.zeros(18 * 1024, dtype=torch.int64, device="cuda")
L.check(L.lib().qnnb_debug_set_trace(L.ptr(buf), buf.numel()))
plan.forward(x)
torch.cuda.synchronize()
L.check(L.lib().qnnb_debug_set_trace(None, 0))
b = buf.cpu().numpy()
names = {1: "entry", 2: "prologue done", 3: "griddep_wait done", 10: "MMA: layer input ready", 11: "MMA: acc free", 12: "MMA: committed",
         20: "epi: acc ready", 21: "epi: block done", 22: "epi:   chunk loaded", 23: "epi:   chunk stored", 30: "wrk: raw image arrived", 31: "wrk: im2col built", 33: "wrk: dense(prev) done",
         34: "wrk: last dense done", 40: "teardown"}
ev = []
for w in range(17):
    reg = b[w * 1024:(w + 1) * 1024]
    for i in range(int(reg[0])):
        ev.append((int(reg[3 + 2 * i]), int(reg[2 + 2 * i]) >> 32, int(reg[2 + 2 * i]) & 0xffffffff, w))
ev.sort()
t0 = ev[0][0]
show = set(int(v) for v in os.environ.get("TRACE_WARPS", "0,2,3").split(","))
lim = int(os.environ.get("TRACE_EVENTS", "120"))
k = 0
for t, tag, idx, w in ev:
    if w in show:
        print("%8d ns  w%-2d %-26s l=%d blk=%d" % ((t - t0) / 1.965, w, names.get(tag, tag), idx >> 4, idx & 15))
        k += 1
        if k >= lim:
            break
