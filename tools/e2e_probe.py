"""Where does the end-to-end (host -> H2D -> plan -> D2H) step time go?  cfg3, batch 1024."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qnn_b200 as q
from helpers import make_cf, CONFIGS, assign_weights_from_spec
from oracle import netspec
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cf = make_cf(**CONFIGS[name])
model = q.build_model(cf)
nodes = netspec.build_spec(cf)
netspec.set_weights(nodes, netspec.random_weights(nodes, seed=7, bias_range=0.1, bn="spread"))
assign_weights_from_spec(model, nodes)
host = [torch.from_numpy(np.random.default_rng(i).integers(0, 256, size=(batch, cf.dim, cf.dim, cf.channels), dtype=np.uint8)).pin_memory() for i in range(8)]
dev = torch.empty_like(host[0], device="cuda")
nbytes = host[0].numel()
# 1. raw H2D rate, one stream, back to back
st = torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        for i in range(200):
            dev.copy_(host[i % 8], non_blocking=True)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("H2D %d B x200: %.1f GB/s (%.1f us per copy; host issue %.1f us per copy)" % (nbytes, nbytes * 200 / dt / 1e9, dt / 200 * 1e6, t_issue / 200 * 1e6))
# 2. the public path
for i in range(6):
    model.predict(host[i % 8])
torch.cuda.synchronize()
for depth in (3,):
    N = 400
    t0 = time.perf_counter()
    pend = []
    t_async = t_res = 0.0
    for i in range(N):
        a = time.perf_counter()
        pend.append(model.predict_async(host[i % 8]))
        b = time.perf_counter()
        t_async += b - a
        if len(pend) >= depth:
            pend.pop(0).result()
            t_res += time.perf_counter() - b
    for h in pend:
        h.result()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("predict_async depth %d: %.1f us per step = %.2f M img/s; host time in predict_async %.1f us, in result() %.1f us"
          % (depth, dt / N * 1e6, batch * N / dt / 1e6, t_async / N * 1e6, t_res / N * 1e6))
# 3. enqueue only (no result reads until the end): is the host the limit?
plan = model.plan()
t0 = time.perf_counter()
hs = []
for i in range(3):
    hs.append(model.predict_async(host[i % 8]))
t_issue3 = (time.perf_counter() - t0) / 3
for h in hs:
    h.result()
print("predict_async issue cost alone: %.1f us" % (t_issue3 * 1e6))
