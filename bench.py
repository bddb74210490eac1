#!/usr/bin/env python
"""Headline benchmark: images/sec of the CIFAR-10 VGG full-qnn w4a4 forward (BASELINE.json
configs[2], batch 1024 per GPU) through the fused plan; one JSON line on stdout (rank 0).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle O2a)

A *step* is one forward pass of the hot path over one batch.  Batches are independent, so N
GPUs run N shards with no data-path collective except the final logit all-gather (NCCL), which
is inside the timed region; scaling is weak (1024 images per GPU per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "images/sec CIFAR-10 VGG full-qnn w4a4 inference (whole job)"


def metric_name(workload):
    # the headline metric is quoted on cfg3 (BASELINE.json configs[2]); other workloads say so in the metric string
    return METRIC if workload == "cfg3" else "images/sec %s inference (whole job)" % workload

WORKLOADS = {
    # name -> (config kwargs, batch per GPU)
    "cfg3": (dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG', nla=1, nfa=64, nlb=1, nfb=128, nlc=1, nfc=256), 1024),
    "cfg4": (dict(network_type='full-qnn', wbits=8, abits=8, architecture='VGG', nla=3, nfa=256, nlb=3, nfb=256, nlc=3, nfc=256), 4096),
    "cfg2": (dict(network_type='full-bnn', architecture='VGG', nla=1, nfa=64, nlb=1, nfb=128, nlc=1, nfc=256), 256),
    "cfg1": (dict(network_type='full-qnn', wbits=2, abits=2, architecture='VGG', dataset='MNIST', dim=28, channels=1,
                  nla=1, nfa=64, nlb=1, nfb=64, nlc=1, nfc=64), 100),
    # BASELINE.json configs[4]: ResNet-(6n+2), n = 10, w4 kernels with fp32 LeakyReLU activations, and its ternary twin
    "cfg5": (dict(network_type='qnn', wbits=4, abits=4, architecture='RESNET', nres=10), 1024),
    "cfg5t": (dict(network_type='tnn', wbits=4, abits=4, architecture='RESNET', nres=10), 1024),
}


def make_cf(**kw):
    base = dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG', dataset='CIFAR-10', dim=32,
                channels=3, classes=10, nla=1, nfa=64, nlb=1, nfb=128, nlc=1, nfc=256, nres=3, pfilt=1,
                kernel_initializer='glorot_uniform', kernel_regularizer=0.)
    base.update(kw)
    return types.SimpleNamespace(**base)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def int8_peak_tops(pk):
    """Dense int8 tensor-core ceiling.  A tcgen05 kind::i8 micro-benchmark result (tools/) overrides the
    provisional 2 x measured bf16."""
    p = os.path.join(ROOT, "profiles", "int8_peak.json")
    if os.path.exists(p):
        return float(json.load(open(p))["int8_tops"]), "measured tcgen05 kind::i8 micro-benchmark"
    return 2.0 * pk["bf16_tflops"], "2 x %s bf16 burst (provisional)" % pk["source"]


# ------------------------------------------------------------------------------ algorithmic work
KIND_BYTES = {"u8": 1.0, "i8": 1.0, "b1": 1.0 / 8.0, "f32": 4.0}


def plan_work(plan, env):
    """Same accounting as layer_work but read off the executed plan (any architecture): per fused step
    (name, ops, bytes) with ops = 2*MACs and bytes = input + residual + output tensors at their stored width + packed
    kernel."""
    out = []
    ci = di = 0
    for st in plan.steps:
        if st.kind == "conv":
            lay = st.layer
            x, y = env[st.src], env[st.out]
            n, h, w, cin = x.shape
            _, oh, ow, cout = y.shape
            ph, pw = (2 * oh, 2 * ow) if st.pool else (oh, ow)
            macs = n * ph * pw * lay.kernel_size[0] * lay.kernel_size[1] * cin * cout
            byts = n * h * w * cin * KIND_BYTES[x.kind] + n * oh * ow * cout * KIND_BYTES[y.kind] + lay.kernel.size
            if st.res is not None:
                r = env[st.res]
                byts += float(np.prod(r.shape)) * KIND_BYTES[r.kind]
            out.append(("conv%d" % ci, 2 * macs, byts, ("conv", x.kind, h, w, cin, cout, lay.kernel_size, lay.strides, bool(st.pool), y.kind)))
            ci += 1
        elif st.kind == "dense":
            x, _, _ = plan._resolve_dense_input(st.src, env)
            n = int(x.shape[0])
            fin = int(np.prod(x.shape[1:]))
            units = st.layer.units
            out.append(("dense%s" % ("" if di == 0 else di), 2 * n * fin * units, n * fin * KIND_BYTES[x.kind] + n * units * 4 + fin * units,
                        ("dense", x.kind, fin, units)))
            di += 1
        else:
            t = env[st.out]
            out.append((st.layer.name, 0, 2 * float(np.prod(t.shape)) * KIND_BYTES[t.kind], ("layer", st.layer.name)))
    return out


def ncu_traffic(workload, step_indices, n_steps):
    """DRAM bytes per launch (read + write), averaged over the given kernels of one forward, from the committed
    Nsight Compute summary of this workload (profiles/r2_ncu_<workload>.csv, else round 1's; one row per launch of one
    forward)."""
    import csv
    path = os.path.join(ROOT, "profiles", "r2_ncu_%s.csv" % workload)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r1_ncu_%s.csv" % workload)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, body = rows[0], rows[1:]
    if len(body) > n_steps:
        return None
    step_indices = [k for k in step_indices if k < len(body)]      # a partial capture covers the first kernels of a forward
    if not step_indices:
        return None
    def col(prefix):
        for i, h in enumerate(hdr):
            if h.startswith(prefix):
                unit = h[h.index("[") + 1:h.index("]")] if "[" in h else "byte"
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
                return sum(float(body[k][i] or 0) for k in step_indices) * mult / len(step_indices)
        return 0.0
    return col("dram_read") + col("dram_write")


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled about every millisecond through NVML from a thread that is
    started BEFORE the warm-up (so nothing is forked or spawned inside a timed region).  ``summary(t0, t1)`` reports
    the samples that fall inside the host-clock window of the timed region."""
    HW_SLOWDOWN, SW_POWER_CAP, SW_THERMAL, HW_THERMAL = 0x8, 0x4, 0x20, 0x40

    def __init__(self, index=0, uuid=None):
        self.samples, self.stop, self.index, self.h, self.nv = [], False, index, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None
            self.max_mhz = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def _reasons(self):
        nv = self.nv
        for fn in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            if hasattr(nv, fn):
                try:
                    return int(getattr(nv, fn)(self.h))
                except Exception:
                    pass
        return 0

    def run(self):
        if self.h is None:
            return self.run_smi()
        nv = self.nv
        while not self.stop:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), self._reasons()))
            except Exception:
                pass
            time.sleep(0.001)

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def run_smi(self):
        # NVML python binding unavailable: poll nvidia-smi (coarse: one sample per ~100 ms)
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                f = [v.strip() for v in o.split(",")]
                bits = 0
                for v, b in zip(f[2:6], (self.HW_SLOWDOWN, self.HW_THERMAL, self.SW_THERMAL, self.SW_POWER_CAP)):
                    if v.lower().startswith("active"):
                        bits |= b
                self.max_mhz = float(f[1])
                self.samples.append((time.perf_counter(), float(f[0]), bits))
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.t.start()
        return self

    def close(self):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self, t0, t1):
        pad = 0.002
        inside = [s for s in self.samples if t0 - pad <= s[0] <= t1 + pad]
        where = "inside the timed region (+-2 ms)"
        if not inside and self.samples:           # region shorter than the sampling period: the nearest sample either side
            mid = 0.5 * (t0 + t1)
            inside = sorted(self.samples, key=lambda s: abs(s[0] - mid))[:2]
            where = "nearest samples to the timed region"
        bits = 0
        for s in inside:
            bits |= s[2]
        names = [(self.HW_SLOWDOWN, "hw_slowdown"), (self.HW_THERMAL, "hw_thermal_slowdown"), (self.SW_THERMAL, "sw_thermal_slowdown"),
                 (self.SW_POWER_CAP, "sw_power_cap")]
        return {"sm_mhz": float(np.median([s[1] for s in inside])) if inside else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in names if bits & b), "samples": len(inside), "window": where,
                "source": "nvml" if self.h is not None else "nvidia-smi"}


def measure_int8_peak():
    """Dense int8 tensor-core ceiling measured in THIS process by the tcgen05 kind::i8 micro-benchmark
    (tools/int8_peak.cu built as tools/libint8peak.so): (burst TOP/s, sustained TOP/s) or None."""
    import ctypes as C
    path = os.path.join(ROOT, "tools", "libint8peak.so")
    if not os.path.exists(path):
        return None
    try:
        h = C.CDLL(path)
        h.qnnb_int8_peak.restype = C.c_int
        h.qnnb_int8_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        burst, sus = C.c_double(), C.c_double()
        if h.qnnb_int8_peak(2000, 6, C.byref(burst), C.byref(sus)) != 0:
            return None
        return float(burst.value), float(sus.value)
    except Exception:
        return None


# ------------------------------------------------------------------------------ CPU reference arm
def build_spec(cfkw, seed=42):
    from oracle import netspec
    cf = make_cf(**cfkw)
    nodes = netspec.build_spec(cf)
    netspec.set_weights(nodes, netspec.random_weights(nodes, seed=seed, bias_range=0.1, bn="spread"))
    return cf, nodes


def cpu_reference_rate(nodes, cf, sample, repeats=3, warm=1):
    """images/sec of the reference's forward (oracle O2a: fp32, per-forward quantise, scaling identity,
    un-fused BN / activation / pooling) on all host cores."""
    import torch
    from oracle import refstate
    torch.set_num_threads(os.cpu_count() or 1)
    x = np.random.default_rng(99).integers(0, 256, size=(sample, cf.dim, cf.dim, cf.channels), dtype=np.uint8)
    for _ in range(warm):
        refstate.forward(nodes, x, trick=True)
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        refstate.forward(nodes, x, trick=True)
        ts.append(time.perf_counter() - t0)
    return sample / float(np.median(ts)), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfkw, batch = WORKLOADS[args.workload]
    cf, nodes = build_spec(cfkw)
    sample = min(batch, args.ref_sample)
    import torch
    from oracle import refstate
    torch.set_num_threads(os.cpu_count() or 1)
    x = np.random.default_rng(99).integers(0, 256, size=(sample, cf.dim, cf.dim, cf.channels), dtype=np.uint8)
    for _ in range(args.warmup):
        refstate.forward(nodes, x, trick=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        refstate.forward(nodes, x, trick=True)
    dt = time.perf_counter() - t0
    rate = sample * args.steps / dt
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": rate, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s %s %s w%da%d, reference CPU forward" % (args.workload, cf.dataset, cf.architecture, cf.network_type, cf.wbits, cf.abits),
                       "sample_images_per_step": sample},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d images per step (oracle O2a: torch-CPU fp32 restatement of the reference graph)" % sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
INTEGER_TYPES = ("full-qnn", "full-bnn", "qbnn", "qtnn")


def workload_label(name, cf, batch):
    if cf.architecture == "VGG":
        return "%s: %s VGG %s w%da%d %d/%d/%d x %d/%d/%d, batch %d per GPU" % (
            name, cf.dataset, cf.network_type, cf.wbits, cf.abits, cf.nla, cf.nlb, cf.nlc, cf.nfa, cf.nfb, cf.nfc, batch)
    return "%s: %s ResNet-%d %s w%da%d, batch %d per GPU" % (name, cf.dataset, 6 * cf.nres + 2, cf.network_type, cf.wbits, cf.abits, batch)


def layer_rooflines(work, per, pk, i8_burst, i8_src):
    """Per fused step: which roof bounds it (arithmetic intensity against the ridge), what it achieved, what fraction."""
    ridge = (i8_burst * 1e12) / (pk["hbm_gbs"] * 1e9)
    rows = []
    for (name, ops, byts, _sig), ms in zip(work, per):
        ms = float(ms)
        if ops / max(byts, 1.0) > ridge:
            ach = ops / (ms * 1e-3) / 1e12
            rows.append({"kernel": name, "ms": ms, "bound": "tensor", "achieved": ach, "unit": "TOP/s", "peak": i8_burst, "frac": ach / i8_burst,
                         "floor_ms": ops / (i8_burst * 1e12) * 1e3})
        else:
            ach = byts / (ms * 1e-3) / 1e9
            rows.append({"kernel": name, "ms": ms, "bound": "hbm", "achieved": ach, "unit": "GB/s", "peak": pk["hbm_gbs"], "frac": ach / pk["hbm_gbs"],
                         "floor_ms": byts / (pk["hbm_gbs"] * 1e9) * 1e3})
    return rows


def measure(args, name, ctx, steps, warmup, streams=0, full=True):
    """Build workload `name`, check it against the exact oracle, time `steps` forward passes (device events, inputs
    resident in HBM and larger than L2), time every fused kernel on its own, and -- full=True -- the end-to-end path."""
    import torch
    import torch.distributed as dist
    import qnn_b200 as q
    from qnn_b200 import kernels as K
    from helpers import assign_weights_from_spec
    world, rank, dev, clk = ctx["world"], ctx["rank"], ctx["dev"], ctx["clk"]

    cfkw, batch = WORKLOADS[name]
    if args.batch and name == args.workload:
        batch = args.batch
    cf, nodes = build_spec(cfkw)
    q.reset_names()
    model = q.build_model(cf)
    assign_weights_from_spec(model, nodes)
    impl = {"auto": 0, "generic": 1, "tcgen05": 2}[args.kernels]
    # SM share of the timed serving loop: (SMs per persistent kernel, streams).  Several independent batches in flight,
    # each kernel on ITS share of the SMs, hide each other's pipeline fill / drain and partial last waves (measured
    # sweep: profiles/r2_sm_share_sweep.txt).  The large configuration keeps the whole device per kernel.
    AUTO_SHARE = {"cfg3": (37, 8), "cfg2": (37, 8), "cfg1": (37, 8), "cfg5": (74, 2), "cfg5t": (74, 2)}
    share, auto_streams = AUTO_SHARE.get(name, (0, 0))
    if steps < 64:                                           # a short queue: fewer batches in flight, shorter drain
        if name == "cfg3":
            share, auto_streams = 74, 4
        elif name in ("cfg2", "cfg1"):
            share, auto_streams = 0, 0
    if args.sm_share >= 0:
        share = args.sm_share
    if streams > 0 and args.sm_share < 0 and streams != auto_streams:
        share = 0                                           # an explicit --streams keeps the whole-device kernels
    plan = model.plan(impl, sm_share=share)
    plan_alone = model.plan(impl)                           # per-kernel timing: each kernel alone on the whole device

    # ---- synthetic inputs: NBUF distinct resident batches, total > L2 (126 MB), rotated every step
    img_bytes = cf.dim * cf.dim * cf.channels
    nbuf = max(4, int(np.ceil(160e6 / (batch * img_bytes))))
    rng = np.random.default_rng(1234 + rank)
    host = [torch.from_numpy(rng.integers(0, 256, size=(batch, cf.dim, cf.dim, cf.channels), dtype=np.uint8)).pin_memory()
            for _ in range(min(nbuf, 8))]
    bufs = [host[i % len(host)].to(dev) for i in range(nbuf)]

    # warm-up (also packs weights / uploads constants)
    plan.launches = 0
    out0 = plan.forward(bufs[0])
    launches_per_step = plan.launches
    torch.cuda.synchronize()

    # ---- parity spot-check of the benchmarked configuration against the exact oracle (small slice)
    from oracle import exact
    xs = host[0][:8].numpy()
    got_s, want_s = model.predict(xs, impl=impl), exact.forward(nodes, xs)
    if cf.network_type in INTEGER_TYPES:
        ok = bool(np.array_equal(got_s, want_s))                      # integer paths: bit-exact
    else:                                                             # fp32 activations: north_star tolerance
        ok = bool(np.abs(got_s - want_s).max() <= 1e-4 * max(float(np.abs(want_s).max()), 1e-30))

    # ---- logit gather for N > 1.  "peer" (default): rank 0 exports a ring of [world*batch, classes] buffers over CUDA
    # IPC and every rank's final dense kernel stores its block straight into it over NVLink -- no per-step collective.
    # "nccl": one all-gather per step captured inside the step's CUDA graph (one communicator per stream).
    NS = streams if streams > 0 else (auto_streams if (share and auto_streams) else (1 if name in ("cfg4", "cfg5", "cfg5t") else 4))
    sts = [torch.cuda.Stream() for _ in range(NS)]
    mode, peer, groups, gflat = "none", None, None, None
    PEER_SLOTS = 8
    if world > 1:
        mode = args.gather
        if mode == "peer":
            from qnn_b200.sharding import PeerGather
            flag = torch.ones(1, device=dev)
            try:
                peer = PeerGather(batch, cf.classes, slots=PEER_SLOTS)
            except Exception as exc:
                print("rank %d: peer-mapped logit buffer unavailable (%s)" % (rank, exc), file=sys.stderr)
                flag.zero_()
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if float(flag.item()) < 1.0:
                peer, mode = None, "nccl"
        if mode == "nccl":
            groups = [dist.new_group(backend="nccl") for _ in range(NS)]
            gflat = [torch.empty((world * batch, cf.classes), dtype=torch.float32, device=dev) for _ in range(NS)]
            for s_i in range(NS):                      # communicator set-up cannot happen under capture: warm each one up
                with torch.cuda.stream(sts[s_i]):
                    sts[s_i].wait_stream(torch.cuda.current_stream())
                    dist.all_gather_into_tensor(gflat[s_i], out0, group=groups[s_i])
            torch.cuda.synchronize()

    # ---- CUDA graphs: one per input buffer.  Consecutive steps are independent batches, so they are replayed
    # round-robin on NS streams (graph i always on stream i % NS, with that stream's private memory pool): the
    # tail of one batch overlaps the head of the next, as in a serving loop.
    def step_body(i, b, pl=None):
        pl = pl or plan
        if peer is not None:
            return pl.forward(b, out=peer.block(i % PEER_SLOTS))
        o = pl.forward(b)
        if mode == "nccl":
            dist.all_gather_into_tensor(gflat[i % NS], o, group=groups[i % NS])
        return o

    # SPG consecutive steps (distinct input batches) are captured per graph, so the timed region is a handful of graph
    # replays issued up front: a step of the small configurations (cfg1: one 12 us launch; cfg2: 13-21 us) is about as long
    # as the host's cost of replaying a graph, and even for the 40 us steps a single host hiccup between replays would
    # show up in a 20-step window.  SPG divides `steps`, so the timed region is still exactly `steps` steps.
    SPG = 1
    if args.graphs and len(plan.steps) <= 6 and batch * img_bytes <= 4000000:      # short steps only: coarse groups lengthen the drain
        SPG = max(d for d in range(1, 9) if steps % d == 0 and nbuf // d >= NS)
    ngroups = nbuf // SPG
    graphs = None
    if args.graphs:
        import gc
        gc.collect()
        graphs = []
        pools = [torch.cuda.graph_pool_handle() for _ in range(NS)]
        for gi in range(ngroups):
            st = sts[gi % NS]
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pools[gi % NS], stream=st):
                    for t in range(SPG):
                        o = step_body(gi * SPG + t, bufs[gi * SPG + t])
            graphs.append((g, o))
        torch.cuda.synchronize()
    nround = (ngroups // NS) * NS                      # keeps graph index -> stream mapping fixed
    # With an SM share, the LAST batches of the queue get whole-device kernels again: once fewer batches are in flight than
    # shares exist, a kernel confined to its share would leave the other SMs idle (what a serving loop does when its
    # queue runs dry).  Same inputs, same arithmetic, the whole-device plan's graphs.
    G_TOTAL = steps // SPG
    TAIL = 0
    if share > 0 and graphs is not None and G_TOTAL > NS:
        TAIL = 1 if SPG > 1 else max(NS // 2, 1)
    graphs_tail = {}
    for gi in range(G_TOTAL - TAIL, G_TOTAL):
        j = gi % nround
        st = sts[j % NS]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pools[j % NS], stream=st):
                for t in range(SPG):
                    o = step_body(j * SPG + t, bufs[j * SPG + t], plan_alone)
        graphs_tail[j] = (g, o)
    torch.cuda.synchronize()

    def run_group(gi, tail=False):
        """SPG consecutive steps (one graph replay)."""
        j = gi % nround
        with torch.cuda.stream(sts[j % NS]):
            if graphs is not None:
                (graphs_tail if tail else graphs)[j][0].replay()   # forward(s) (+ logit hand-over)
            else:
                step_body(j, bufs[j])

    def fence_all(ev=None):
        cur = torch.cuda.current_stream()
        for st in sts:
            cur.wait_stream(st)
        if ev is not None:
            ev.record()
        for st in sts:
            st.wait_stream(cur)

    # warm-up: at least `warmup` steps, and every graph the timed region will replay once (the first launch of a CUDA
    # graph uploads it to the device: tens of microseconds that are not part of a step)
    for gi in range(max(-(-max(warmup, 3) // SPG), min(G_TOTAL, nround))):
        run_group(gi, tail=gi >= G_TOTAL - TAIL)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    fence_all(e0)
    for gi in range(G_TOTAL):
        run_group(gi, tail=gi >= G_TOTAL - TAIL)
    fence_all(e1)
    torch.cuda.synchronize()
    t_host1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()                                  # every rank's stores have landed in the gathering rank's memory
    clocks = clk.summary(t_host0, t_host1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = batch * world * steps / (ms * 1e-3)

    # ---- N > 1: the gathered logits of the last step on rank 0 must equal an NCCL all-gather of the same shards
    gather_ok = None
    if peer is not None:
        j = ((steps // SPG - 1) % nround) * SPG + SPG - 1          # the last step's input batch
        local = plan.forward(bufs[j])
        ref = torch.empty((world * batch, cf.classes), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(ref, local)
        torch.cuda.synchronize()
        if rank == 0:
            gather_ok = bool(torch.equal(peer.gathered(j % PEER_SLOTS), ref))

    # ---- per-kernel timing (CUDA events around a graph of REPS launches, same rotating inputs) for the roofline
    tplan = plan_alone
    envs = [tplan.run(bufs[i]) for i in range(2)]
    work = plan_work(tplan, envs[0])
    fused = tplan.fused_available(bufs[0])
    timed_steps = list(tplan.steps)
    if fused:
        # the whole net is ONE launch (csrc/net_fused.cu): one row -- ops of every layer; bytes = what crosses HBM, i.e. the
        # images in, the logits out and the packed kernels once
        kern_bytes = sum(float(st.layer.kernel.size) for st in tplan.steps)
        work = [("net (whole-network kernel)", float(sum(w[1] for w in work)), float(batch * (img_bytes + cf.classes * 4) + kern_bytes),
                 ("net", cf.dim, cf.channels, tuple(st.layer.kernel.shape[-1] for st in tplan.steps)))]
        timed_steps = [None]
    per = np.zeros(len(timed_steps))
    REPS = 10
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for si, st in enumerate(timed_steps):
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for r in range(REPS):
                    if st is None:
                        tplan.forward(bufs[r % 2])
                        continue
                    env = dict(envs[r % 2])
                    tplan.run_step(st, env)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        per[si] = a.elapsed_time(b) / (3 * REPS)
        del g
    del envs
    pk, (i8_burst, i8_sus, i8_src) = ctx["peaks"], ctx["int8"]
    rows = layer_rooflines(work, per, pk, i8_burst, i8_src)
    # dominant kernel = the launch shape (same kernel, same geometry) with the largest TOTAL time in one forward;
    # achieved = its algorithmic work per launch / its average launch duration
    shapes = {}
    for si, wk in enumerate(work):
        shapes.setdefault(wk[3], []).append(si)
    members = shapes[max(shapes, key=lambda k: sum(per[i] for i in shapes[k]))]
    nm = len(members)
    kname = work[members[0]][0] if nm == 1 else "%s..%s (%d launches of one shape)" % (work[members[0]][0], work[members[-1]][0], nm)
    ops = sum(work[i][1] for i in members) / nm
    byts = sum(work[i][2] for i in members) / nm
    kernel_ms = float(sum(per[i] for i in members) / nm)
    r0 = rows[members[0]]
    achieved = (ops / (kernel_ms * 1e-3) / 1e12) if r0["bound"] == "tensor" else (byts / (kernel_ms * 1e-3) / 1e9)
    roof = {"bound": r0["bound"], "achieved": achieved, "peak": r0["peak"], "unit": r0["unit"], "frac": achieved / r0["peak"],
            "traffic": ncu_traffic(name, members, len(timed_steps)), "kernel": kname, "kernel_ms": kernel_ms,
            "peak_source": i8_src if r0["bound"] == "tensor" else pk["source"] + " copy bandwidth",
            "share_of_step": float(sum(per[i] for i in members) / max(per.sum(), 1e-30)),
            "algorithmic": {"ops_per_launch": float(ops), "bytes_per_launch": float(byts)},
            "per_kernel_ms": {w[0]: float(p) for w, p in zip(work, per)},
            "layers": [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items() if k != "peak"} for r in rows],
            # the whole step against the sum of its layers' own roofline floors
            "step_floor_ms": float(sum(r["floor_ms"] for r in rows)),
            "step_frac": float(sum(r["floor_ms"] for r in rows) / (ms / steps))}
    total_ops = float(sum(w[1] for w in work))
    res = {"workload": workload_label(name, cf, batch), "value": value, "ms_per_step": ms / steps, "steps": steps, "batch": batch,
           "parity_vs_exact_oracle": ok, "clocks": clocks, "launches_per_step": launches_per_step, "streams": NS,
           "cuda_graphs": graphs is not None, "steps_per_graph": SPG, "sm_share": share, "tail_steps": TAIL * SPG, "nbuf": nbuf, "img_bytes": img_bytes, "classes": cf.classes, "cf": cf, "nodes": nodes,
           "roofline": roof, "gather": mode, "gather_verified": gather_ok,
           "int8_tops_achieved": total_ops / (ms / steps * 1e-3) / 1e12, "int8_frac_of_burst_peak": total_ops / (ms / steps * 1e-3) / 1e12 / i8_burst}

    # ---- end to end through the public API: pinned host batch -> H2D -> fused plan (CUDA graph) -> D2H logits,
    # every step; steps are software-pipelined PIPELINE_DEPTH (four) deep (model.predict_async), as a serving loop would.
    if full:
        torch.cuda.synchronize()
        for i in range(3):
            model.predict(host[i % len(host)], impl=impl)
        warm = [model.predict_async(host[i % len(host)], impl=impl) for i in range(max(warmup, 3))]   # the pipelined path itself, untimed
        for h in warm:
            h.result()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        depth = model.plan(impl).PIPELINE_DEPTH
        t0 = time.perf_counter()
        pending = []
        checksum = 0.0
        for i in range(steps):
            pending.append(model.predict_async(host[i % len(host)], impl=impl))
            if len(pending) >= depth:
                checksum += float(pending.pop(0).result()[0, 0])
        for h in pending:
            checksum += float(h.result()[0, 0])
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": batch * world * steps / float(te.item()), "unit": "images/s", "h2d_bytes_per_step": batch * img_bytes,
                      "d2h_bytes_per_step": batch * cf.classes * 4}

    # ---- release everything this workload holds on the device (the next one may need the room)
    graphs = None
    if peer is not None:
        torch.cuda.synchronize()
        peer.close()
    for pl in list(getattr(model, "_plans", {}).values()):
        pl.close()
    del bufs, plan, model
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    clk = ClockSampler(local, uuid).start()              # sampling runs from before the warm-up to the end

    pk = peaks()
    measured = measure_int8_peak()
    if measured:
        int8 = (measured[0], measured[1], "tcgen05 kind::i8 micro-benchmark measured in this process (tools/int8_peak.cu), burst")
    else:
        tops, src = int8_peak_tops(pk)
        int8 = (tops, tops, src)
    ctx = {"world": world, "rank": rank, "dev": dev, "clk": clk, "peaks": pk, "int8": int8}

    m = measure(args, args.workload, ctx, args.steps, args.warmup, streams=args.streams, full=True)

    # ---- secondary workloads on the same box, same process (N = 1, default headline run only): the large w8a8 net the
    # 60 % tensor-core target is stated on (BASELINE.json configs[3]), the XNOR config and the ResNet config
    secondary = {}
    if world == 1 and args.workload == "cfg3" and not args.no_secondary:
        for nm, st in (("cfg4", 20), ("cfg2", 200), ("cfg5", 20)):
            try:
                # own step counts (a secondary run is 2-100 ms of GPU time); tiny --steps (smoke runs) shrink them too
                r = measure(args, nm, ctx, st if args.steps >= 10 else min(st, max(args.steps, 3)), 3, full=False)
                ro = r["roofline"]
                secondary[nm] = {"workload": r["workload"], "value": r["value"], "unit": "images/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                                 "parity_vs_exact_oracle": r["parity_vs_exact_oracle"], "clocks": r["clocks"],
                                 "int8_tops_achieved": r["int8_tops_achieved"], "frac_of_int8_burst_peak": r["int8_frac_of_burst_peak"],
                                 "roofline": {k: ro[k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "kernel_ms", "share_of_step",
                                                                 "layers", "step_floor_ms", "step_frac")}}
            except Exception as exc:                     # a secondary workload never costs the headline line
                secondary[nm] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        cf, nodes, batch = m["cf"], m["nodes"], m["batch"]
        cpu = None
        if not args.no_cpu_baseline:
            sample = min(batch, args.ref_sample)
            rate, cores = cpu_reference_rate(nodes, cf, sample)
            cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "%d images, median of 3 (oracle O2a: torch-CPU fp32 restatement of the reference graph)" % sample}
        par = {"none": "single GPU", "peer": "batch-sharded x%d; every rank's final dense kernel stores its logit block into rank 0's "
               "buffer over NVLink (CUDA IPC peer mapping), no per-step collective; NCCL for barriers / timing only" % world,
               "nccl": "batch-sharded x%d, NCCL logit all-gather captured in the step's CUDA graph (one communicator per stream)" % world}[m["gather"]]
        line = {"metric": metric_name(args.workload), "value": m["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": ("int8 (int32 accumulate, fp32 epilogue)" if cf.network_type in INTEGER_TYPES
                          else "fp32 activations as an exact 3 x bf16 split x integer kernel levels (fp32 accumulate)"),
                "data": "synthetic",
                "config": {"workload": m["workload"], "global_batch": batch * world, "parallelism": par,
                           "l2_policy": "inputs larger than L2: %d distinct resident batches (%.0f MB) rotated every step" % (
                               m["nbuf"], m["nbuf"] * batch * m["img_bytes"] / 1e6),
                           "cuda_graphs": m["cuda_graphs"], "steps_per_graph": m["steps_per_graph"], "streams": m["streams"],
                           "sm_share": ("%d SMs per kernel: %d independent batches in flight, their kernels side by side (the last %d steps of the queue "
                                        "get whole-device kernels); roofline.per_kernel_ms times each kernel alone on the whole device"
                                        % (m["sm_share"], m["streams"], m["tail_steps"])) if m["sm_share"] else "whole device per kernel", "kernels": args.kernels,
                           "parity_vs_exact_oracle": m["parity_vs_exact_oracle"]},
                "clocks": m["clocks"], "gpu_launches": m["launches_per_step"] * args.steps,
                "e2e": m["e2e"], "roofline": m["roofline"],
                "int8_peak": {"burst_tops": int8[0], "sustained_tops": int8[1], "source": int8[2]}}
        if m["gather_verified"] is not None:
            line["config"]["peer_gather_equals_nccl_all_gather"] = m["gather_verified"]
        if secondary:
            line["secondary"] = secondary
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    clk.close()
    if world > 1:
        # leave without the NCCL destructor (graphs that touched communicators hung it on 2 GPUs in round 1): every
        # result is printed; drain, meet, exit 0
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--kernels", default="auto", choices=["auto", "generic", "tcgen05"])
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--streams", type=int, default=0, help="0 = auto: 4 for short steps, 1 for the large config")
    ap.add_argument("--sm-share", type=int, default=-1, help="SMs per persistent kernel in the timed loop (-1 = per-workload default, 0 = all)")
    ap.add_argument("--ref-sample", type=int, default=256)
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="world > 1: logits stored into rank 0's peer-mapped buffer by the dense kernel (NVLink), or one NCCL all-gather per step")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg4 / cfg2 / cfg5 block of the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
