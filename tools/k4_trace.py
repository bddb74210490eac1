"""CTA 0 timeline of one K4 launch (needs the TRACE=1 build: QNNB_LIB=.../libqnnb200_trace.so)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qnn_b200 as q
from qnn_b200 import _lib as L, kernels as K
n, h, w, c = (int(v) for v in sys.argv[1:5])
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.normal(0, 1, size=(n, h, w, c)).astype(np.float32)).cuda()
res = torch.from_numpy(rng.normal(0, 1, size=(n, h, w, c)).astype(np.float32)).cuda()
wp = K.pack_weights(torch.from_numpy(rng.uniform(-1, 1, size=(3, 3, c, c)).astype(np.float32)).cuda(), L.W_QUANT, 4, 1.0, L.WFMT_I8)
inv = torch.ones(c, device="cuda"); sh = torch.zeros(c, device="cuda")
epi = K.make_epilogue(0.125, bn_inv=inv, bn_shift=sh, residual=K.QTensor("f32", res, 1.0, c), res_mul=0.5, act=L.ACT_LEAKY)
xq = K.QTensor("f32", x, 1.0, c)
out = torch.empty_like(x)
for _ in range(3):
    K.conv2d(xq, wp, 3, 3, c, 1, epi, out=out)
torch.cuda.synchronize()
buf = torch.zeros(16 * 1024, dtype=torch.int64, device="cuda")
L.check(L.lib().qnnb_debug_set_trace(L.ptr(buf), buf.numel()))
K.conv2d(xq, wp, 3, 3, c, 1, epi, out=out)
torch.cuda.synchronize()
L.check(L.lib().qnnb_debug_set_trace(None, 0))
b = buf.cpu().numpy()
names = {1: "TMA   halo load issued", 2: "TMA   residual slot free", 3: "CVT   halo landed", 4: "CVT   plane slot free", 5: "CVT   planes published",
         6: "MMA   planes ready", 7: "MMA   tile issued", 8: "EPI   residual landed", 9: "EPI   accumulator ready", 10: "EPI   tile stored", 11: "CVT   loads issued", 12: "CVT   planes written", 13: "EPI   math done"}
ev = []
for role in range(4):
    reg = b[role * 1024:(role + 1) * 1024]
    for i in range(int(reg[0])):
        ev.append((int(reg[3 + 2 * i]), int(reg[2 + 2 * i]) >> 32, int(reg[2 + 2 * i]) & 0xffffffff))
ev.sort()
t0 = ev[0][0]
for t, tag, idx in ev[:int(os.environ.get("TRACE_MAX", "150"))]:
    print("%8d ns  %-26s tile %d" % (t - t0, names.get(tag, tag), idx))
