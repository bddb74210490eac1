"""Time one first-layer (K5) launch shape, replayed from a CUDA graph of REPS launches (PDL chain as in a real step).
usage: k5_probe.py n cout pool abits     env QNNB_K5_EXP = diagnostic bit mask (csrc/conv_first_tc.cu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qnn_b200 as q
from qnn_b200 import _lib as L, kernels as K
n, cout, pool, abits = (int(v) for v in sys.argv[1:5])
REPS = 20
rng = np.random.default_rng(0)
xs = [torch.from_numpy(rng.integers(0, 256, size=(n, 32, 32, 3), dtype=np.uint8)).cuda() for _ in range(4)]
wp = K.pack_weights(torch.from_numpy(rng.uniform(-1, 1, size=(3, 3, 3, cout)).astype(np.float32)).cuda(), L.W_QUANT, min(abits, 8), 1.0, L.WFMT_I8)
inv = torch.from_numpy((rng.uniform(0.3, 0.9, cout) * rng.choice([1, 1, 1, -1], cout)).astype(np.float32)).cuda()
sh = torch.from_numpy(rng.uniform(-0.2, 0.2, cout).astype(np.float32)).cuda()
bias = torch.from_numpy(rng.uniform(-0.1, 0.1, cout).astype(np.float32)).cuda()
epi = K.make_epilogue(K.acc_scale(1.0 / 255.0, 1.0 / (1 << (min(abits, 8) - 1))), bias=bias, bn_inv=inv, bn_shift=sh, act=L.ACT_QUANT, abits=abits, pool=2 if pool else 0)
oh = 16 if pool else 32
outs = [torch.empty((n, oh, oh, cout), dtype=torch.int8, device="cuda") for _ in range(4)]
def run(i):
    K.conv2d(K.QTensor("u8", xs[i % 4], 1.0 / 255.0, 3), wp, 3, 3, cout, 1, epi, impl=L.IMPL_TCGEN05, out=outs[i % 4])
for i in range(3):
    run(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(st):
    with torch.cuda.graph(g, stream=st):
        for i in range(REPS):
            run(i)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * REPS)
byts = n * (3072 + oh * oh * cout)
print("exp=%s n=%d cout=%d pool=%d a%d: %.2f us per launch, %.0f GB/s algorithmic" % (os.environ.get("QNNB_K5_EXP", "0"), n, cout, pool, abits, us, byts / us / 1e3))
