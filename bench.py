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
    Nsight Compute summary of this workload (profiles/r1_ncu_<workload>.csv, one row per launch of one forward)."""
    import csv
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
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([s.strip() for s in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for nm, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


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
def run_ours(args):
    import torch
    import torch.distributed as dist
    import qnn_b200 as q
    from helpers import assign_weights_from_spec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    cfkw, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    cf, nodes = build_spec(cfkw)
    q.reset_names()
    model = q.build_model(cf)
    assign_weights_from_spec(model, nodes)
    impl = {"auto": 0, "generic": 1, "tcgen05": 2}[args.kernels]
    plan = model.plan(impl)

    # ---- synthetic inputs: NBUF distinct resident batches, total > L2 (126 MB), rotated every step
    img_bytes = cf.dim * cf.dim * cf.channels
    nbuf = max(4, int(np.ceil(160e6 / (batch * img_bytes))))
    rng = np.random.default_rng(1234 + rank)
    host = [torch.from_numpy(rng.integers(0, 256, size=(batch, cf.dim, cf.dim, cf.channels), dtype=np.uint8)).pin_memory()
            for _ in range(min(nbuf, 8))]
    bufs = [host[i % len(host)].to(dev) for i in range(nbuf)]
    gather = [torch.empty((batch, cf.classes), dtype=torch.float32, device=dev) for _ in range(world)] if world > 1 else None

    def step_fn(xb):
        out = plan.forward(xb)
        if world > 1:
            dist.all_gather(gather, out)
        return out

    # warm-up (also packs weights / uploads constants)
    plan.launches = 0
    out0 = step_fn(bufs[0])
    launches_per_step = plan.launches
    torch.cuda.synchronize()

    # ---- parity spot-check of the benchmarked configuration against the exact oracle (small slice)
    from oracle import exact
    xs = host[0][:8].numpy()
    got_s, want_s = model.predict(xs, impl=impl), exact.forward(nodes, xs)
    if cf.network_type in ("full-qnn", "full-bnn", "qbnn", "qtnn"):
        ok = bool(np.array_equal(got_s, want_s))                      # integer paths: bit-exact
    else:                                                             # fp32 activations: north_star tolerance
        ok = bool(np.abs(got_s - want_s).max() <= 1e-4 * max(float(np.abs(want_s).max()), 1e-30))

    # ---- CUDA graphs: one per input buffer.  Consecutive steps are independent batches, so they are replayed
    # round-robin on NS streams (graph i always on stream i % NS, with that stream's private memory pool): the
    # tail of one batch overlaps the head of the next, as in a serving loop.
    NS = args.streams if args.streams > 0 else (1 if args.workload in ("cfg4", "cfg5", "cfg5t") else 4)
    # world > 1: the per-step NCCL logit gather is issued eagerly from ONE dedicated communication stream, in step
    # order on every rank (a communicator must see the same sequence everywhere); forward graphs keep overlapping.
    comm = torch.cuda.Stream() if world > 1 else None
    gather_flat = torch.empty((world * batch, cf.classes), dtype=torch.float32, device=dev) if world > 1 else None
    last_comm = [None] * NS
    fwd_done = [torch.cuda.Event() for _ in range(NS)]
    graphs = None
    streams = [torch.cuda.Stream() for _ in range(NS)]
    # --gather graph (default for world > 1): the logit all-gather is CAPTURED inside each forward graph, on one NCCL
    # communicator per stream (a communicator must see the same collective sequence on every rank; graph i always runs
    # on stream i % NS, so per-stream communicators keep that true while the streams overlap).  One graph launch per
    # step then covers forward + gather, instead of ~40 us of eager NCCL enqueue per step on the host.
    gather_in_graph = world > 1 and args.graphs and args.gather == "graph"
    groups, gflat = None, None
    if gather_in_graph:
        try:
            groups = [dist.new_group(backend="nccl") for _ in range(NS)]
            gflat = [torch.empty((world * batch, cf.classes), dtype=torch.float32, device=dev) for _ in range(NS)]
            for s_i in range(NS):                      # communicator set-up cannot happen under capture: warm each one up
                with torch.cuda.stream(streams[s_i]):
                    streams[s_i].wait_stream(torch.cuda.current_stream())
                    dist.all_gather_into_tensor(gflat[s_i], out0, group=groups[s_i])
            torch.cuda.synchronize()
        except Exception as exc:
            if rank == 0:
                print("per-stream communicators unavailable (%s): eager gather" % exc, file=sys.stderr)
            gather_in_graph = False

    def capture_all(with_gather):
        gs = []
        pools = [torch.cuda.graph_pool_handle() for _ in range(NS)]
        for i, b in enumerate(bufs):
            st = streams[i % NS]
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pools[i % NS], stream=st):
                    o = plan.forward(b)
                    if with_gather:
                        dist.all_gather_into_tensor(gflat[i % NS], o, group=groups[i % NS])
            gs.append((g, o))
        torch.cuda.synchronize()
        return gs

    if args.graphs:
        try:
            graphs = capture_all(gather_in_graph)
        except Exception as exc:                       # e.g. a collective that cannot be captured
            if rank == 0:
                print("graph capture failed (%s): retrying without the collective / falling back to eager launches" % exc, file=sys.stderr)
            torch.cuda.synchronize()
            graphs = None
            if gather_in_graph:
                gather_in_graph = False
                try:
                    graphs = capture_all(False)
                except Exception:
                    graphs = None
                    torch.cuda.synchronize()
    nround = (nbuf // NS) * NS                         # keeps graph index -> stream mapping fixed

    def run_step(i):
        j = i % nround
        with torch.cuda.stream(streams[j % NS]):
            if graphs is not None and gather_in_graph:
                graphs[j][0].replay()                # forward + NCCL gather, one launch
            elif graphs is not None:
                if world > 1 and last_comm[j % NS] is not None:
                    streams[j % NS].wait_event(last_comm[j % NS])      # previous logits of this stream's pool were gathered
                graphs[j][0].replay()
                if world > 1:
                    fwd_done[j % NS].record(streams[j % NS])
                    comm.wait_event(fwd_done[j % NS])
                    with torch.cuda.stream(comm):
                        dist.all_gather_into_tensor(gather_flat, graphs[j][1])
                        ev = torch.cuda.Event()
                        ev.record(comm)
                    last_comm[j % NS] = ev
            else:
                step_fn(bufs[j])

    def fence_all(ev=None):
        cur = torch.cuda.current_stream()
        everyone = streams + ([comm] if comm is not None else [])
        for st in everyone:
            cur.wait_stream(st)
        if ev is not None:
            ev.record()
        for st in everyone:
            st.wait_stream(cur)

    for i in range(max(args.warmup, 3)):
        run_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        fence_all(e0)
        for i in range(args.steps):
            run_step(i)
        fence_all(e1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        time.sleep(0.25)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = batch * world * args.steps / (ms * 1e-3)

    # ---- per-kernel timing (CUDA events between launches, same rotating inputs) for the roofline
    envs = [plan.run(bufs[i]) for i in range(2)]
    work = plan_work(plan, envs[0])
    # Each kernel is replayed REPS times from its own CUDA graph (no host launch gaps) on the activations
    # the previous layer produced, alternating between two independent forward environments.
    per = np.zeros(len(plan.steps))
    from qnn_b200 import kernels as K
    REPS = 10
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for si, st in enumerate(plan.steps):
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for r in range(REPS):
                    env = dict(envs[r % 2])
                    plan.run_step(st, env)
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
    # dominant kernel = the launch shape (same kernel, same geometry) with the largest TOTAL time in one forward;
    # achieved = its algorithmic work per launch / its average launch duration
    groups = {}
    for si, wk in enumerate(work):
        groups.setdefault(wk[3], []).append(si)
    dom_sig = max(groups, key=lambda k: sum(per[i] for i in groups[k]))
    members = groups[dom_sig]
    nm = len(members)
    name = work[members[0]][0] if nm == 1 else "%s..%s (%d launches of one shape)" % (work[members[0]][0], work[members[-1]][0], nm)
    ops = sum(work[i][1] for i in members) / nm
    byts = sum(work[i][2] for i in members) / nm
    kernel_ms = float(sum(per[i] for i in members) / nm)
    pk = peaks()
    i8_peak, i8_src = int8_peak_tops(pk)
    ai = ops / byts
    tensor_bound = ai > (i8_peak * 1e12) / (pk["hbm_gbs"] * 1e9)
    traffic = ncu_traffic(args.workload, members, len(plan.steps))
    if tensor_bound:
        achieved = ops / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": i8_peak, "unit": "TOP/s", "frac": achieved / i8_peak,
                "traffic": traffic, "kernel": name, "kernel_ms": kernel_ms, "peak_source": i8_src}
    else:
        achieved = byts / (kernel_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                "traffic": traffic, "kernel": name, "kernel_ms": kernel_ms,
                "peak_source": pk["source"] + " copy bandwidth"}
    roof["share_of_step"] = float(sum(per[i] for i in members) / max(per.sum(), 1e-30))
    roof["algorithmic"] = {"ops_per_launch": float(ops), "bytes_per_launch": float(byts)}
    roof["per_kernel_ms"] = {w[0]: float(p) for w, p in zip(work, per)}

    # ---- end to end through the public API: pinned host batch -> H2D -> fused plan (CUDA graph) -> D2H logits,
    # every step; steps are software-pipelined three deep (model.predict_async), as a serving loop would.
    torch.cuda.synchronize()
    for i in range(3):
        model.predict(host[i % len(host)], impl=impl)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pending = []
    checksum = 0.0
    for i in range(args.steps):
        pending.append(model.predict_async(host[i % len(host)], impl=impl))
        if len(pending) >= 3:
            checksum += float(pending.pop(0).result()[0, 0])
    for h in pending:
        checksum += float(h.result()[0, 0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = batch * world * args.steps / float(te.item())

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            sample = min(batch, args.ref_sample)
            rate, cores = cpu_reference_rate(nodes, cf, sample)
            cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "%d images, median of 3 (oracle O2a: torch-CPU fp32 restatement of the reference graph)" % sample}
        line = {"metric": metric_name(args.workload), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": ("int8 (int32 accumulate, fp32 epilogue)" if cf.network_type in ("full-qnn", "full-bnn", "qbnn", "qtnn")
                          else "fp32 activations as an exact 3 x bf16 split x integer kernel levels (fp32 accumulate)"),
                "data": "synthetic",
                "config": {"workload": ("%s: %s VGG %s w%da%d %d/%d/%d x %d/%d/%d, batch %d per GPU" % (
                               args.workload, cf.dataset, cf.network_type, cf.wbits, cf.abits, cf.nla, cf.nlb, cf.nlc, cf.nfa, cf.nfb, cf.nfc, batch))
                           if cf.architecture == "VGG" else ("%s: %s ResNet-%d %s w%da%d, batch %d per GPU" % (
                               args.workload, cf.dataset, 6 * cf.nres + 2, cf.network_type, cf.wbits, cf.abits, batch)),
                           "global_batch": batch * world, "parallelism": "batch-sharded x%d, NCCL logit all-gather%s" % (
                               world, " captured in the step's CUDA graph (one communicator per stream)" if gather_in_graph else ""),
                           "l2_policy": "inputs larger than L2: %d distinct resident batches (%.0f MB) rotated every step" % (nbuf, nbuf * batch * img_bytes / 1e6),
                           "cuda_graphs": graphs is not None, "streams": NS, "kernels": args.kernels,
                           "parity_vs_exact_oracle": ok},
                "clocks": clk.summary(), "gpu_launches": launches_per_step * args.steps,
                "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": batch * img_bytes,
                        "d2h_bytes_per_step": batch * cf.classes * 4},
                "roofline": roof}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL collectives keep their communicators busy: tearing the process groups down
        # with those graphs alive hung on 2 GPUs.  Drop the graphs, drain the device, meet at a barrier, and leave
        # without the NCCL destructor (all results are printed; the driver only needs exit code 0).
        graphs = None
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
    ap.add_argument("--ref-sample", type=int, default=256)
    ap.add_argument("--gather", default="graph", choices=["graph", "eager"], help="world > 1: NCCL logit gather inside the CUDA graph or eager")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
