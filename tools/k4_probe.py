"""Time one K4 (fp32 tensor-core conv) launch shape under the QNNB_K4_EXP diagnostic modes."""
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
use_res = os.environ.get("PROBE_RES", "1") == "1"
epi = K.make_epilogue(0.125, bn_inv=inv, bn_shift=sh, residual=K.QTensor("f32", res, 1.0, c) if use_res else None, res_mul=0.5, act=L.ACT_LEAKY)
xq = K.QTensor("f32", x, 1.0, c)
out = torch.empty_like(x)
for _ in range(3):
    K.conv2d(xq, wp, 3, 3, c, 1, epi, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    K.conv2d(xq, wp, 3, 3, c, 1, epi, out=out)
e1.record(); torch.cuda.synchronize()
print("res=%s fst=%s rst=%s " % (os.environ.get("PROBE_RES","1"), os.environ.get("QNNB_K4_FST","4"), os.environ.get("QNNB_K4_RST","4")), end=""); print("exp=%s np=%s  %dx%dx%dx%d: %.1f us per launch" % (os.environ.get("QNNB_K4_EXP", "0"), os.environ.get("QNNB_K4_NP", "3"), n, h, w, c, e0.elapsed_time(e1) * 1e3 / 20))
