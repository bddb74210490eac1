"""Fused inference plan: lowers a model graph to a short list of libqnnb200 kernel launches.

Each custom conv / dense layer absorbs the layers that follow it in the reference's builders --
``BatchNormalization``, the residual ``add`` (+ ``Lambda(x*0.5)``), the activation quantiser and
``MaxPooling2D`` (models/vgg.py:15-37, models/resnet.py:57-129) -- into its kernel's epilogue, so
every layer makes exactly one HBM round trip and activations stay int8 / bit-packed between
layers.  ``AveragePooling2D(8) -> Flatten -> Dense`` (models/resnet.py:134-140) becomes one dense
kernel over the un-pooled map with the packed kernel replicated per pixel.
"""
from __future__ import annotations

import gc
import os
import weakref

import numpy as np
import torch

from . import _lib as L
from . import engine as E
from . import kernels as K
from .layers._base import QConv2DBase, QDenseBase

F32 = np.float32
MAX_CHUNK = 8192


class _Step:
    def __init__(self, kind, out, **kw):
        self.kind, self.out = kind, out
        self.__dict__.update(kw)
        self.dev = {}


def _act_of(layer):
    if isinstance(layer, (E.Activation, E.LeakyReLU)):
        return layer.act_spec()
    return None


class _Handle:
    """Result of Plan.predict_async: ``result()`` waits for the slot's D2H copy and returns a private copy.

    A slot is owned by at most one unread handle.  When the ring wraps around onto a slot whose handle has not been
    read yet, the plan first completes that handle (``_settle``: wait + private copy), so a late ``result()`` still
    returns ITS batch; ``gen`` guards against a stale handle touching a slot that has since been reused."""

    def __init__(self, slot):
        self._slot = slot
        self._gen = slot["gen"]
        self._out = None

    def _settle(self):
        slot = self._slot
        if self._out is None and slot is not None:
            if slot["gen"] != self._gen:
                raise RuntimeError("predict_async: the slot of this handle was recycled before its result was kept")
            slot["event"].synchronize()
            self._out = slot["host"].clone()
            slot["owner"] = None
            self._slot = None

    def result(self):
        self._settle()
        return self._out


class Plan:
    def __init__(self, model, impl=L.IMPL_AUTO, sm_share=0):
        """``sm_share``: CTAs (= SMs) each persistent kernel of this plan may take; 0 = the whole device.  A serving loop
        that keeps several independent batches in flight on different streams gives every plan a share (e.g. 37 of
        148 SMs with eight streams): kernels of different batches then run side by side and hide each other's pipeline
        fill / drain and partial last waves -- per-batch latency rises, throughput rises too."""
        # weak: the model owns its plans (model._plans); a strong back-reference would make every dead model -- with
        # the CUDA graphs and pool memory of its plans -- wait for the cyclic collector, which may then run in the
        # middle of somebody else's stream capture
        self._model_ref = weakref.ref(model)
        self.impl = int(impl)
        self.sm_share = int(sm_share)
        ins, outs = model._graph()
        self.order = E.topo_order(outs)
        self.index = {id(t): i for i, t in enumerate(self.order)}
        self.input_idx = self.index[id(ins[0])]
        self.output_idx = self.index[id(outs[0])]
        self.steps = []
        self.launches = 0
        self.launches_per_forward = 0
        self._slots = {}
        self._compile()
        self.launches_per_forward = len(self.steps)
        # whole-network launch (csrc/net_fused.cu): decided on the first forward (needs the input shape); None = not yet
        self.fuse = self.impl == L.IMPL_AUTO and os.environ.get("QNNB_FUSED_NET", "1") != "0"
        self._fused_ok = {}
        self._layers = [t.layer for t in self.order]
        self._epoch = -1
        self._versions = None
        self._sync_weights()

    @property
    def model(self):
        return self._model_ref()

    # ------------------------------------------------------------------ lifetime / staleness
    def close(self):
        """Drop everything that lives on the device on behalf of this plan: captured graphs with their private
        pools, static input / output buffers, pinned result buffers, cached per-step constants."""
        for ring in self._slots.values():
            for slot in ring["slots"]:
                owner = slot.get("owner")
                if owner is not None:
                    try:
                        owner._settle()                 # an outstanding handle keeps its result
                    except Exception:
                        pass
                slot.clear()
        self._slots = {}
        for st in self.steps:
            st.dev = {}

    def _sync_weights(self):
        """Layers bump a version on every ``set_weights`` (engine.Layer._touch).  If any layer of this plan changed
        since the last launch, the cached BN constants and every captured graph (they hold raw pointers to the old
        packed kernels / biases) are dropped and rebuilt lazily -- ``model.layers[i].set_weights(...)`` after a first
        ``predict`` must never replay stale buffers."""
        if self._epoch == E.weights_epoch():
            return
        versions = [getattr(l, "_wversion", 0) for l in self._layers]
        if self._versions is not None and versions != self._versions:
            self.close()
        self._versions = versions
        self._epoch = E.weights_epoch()

    # ------------------------------------------------------------------ compile
    def _consumers(self):
        cons = {i: [] for i in range(len(self.order))}
        for j, t in enumerate(self.order):
            for s in t.inputs:
                cons[self.index[id(s)]].append(j)
        return cons

    def _compile(self):
        order, index = self.order, self.index
        cons = self._consumers()
        done = set()
        self.alias = {}

        def sole(i):
            return cons[i][0] if len(cons[i]) == 1 else None

        for i, t in enumerate(order):
            if i in done:
                continue
            lay = t.layer
            src = [index[id(s)] for s in t.inputs]
            if isinstance(lay, E.InputLayer):
                continue
            if isinstance(lay, QConv2DBase):
                g = dict(layer=lay, src=src[0], bn=None, res=None, res_mul=1.0, act=lay._act, pool=False)
                state = 3 if lay._act is not None else 0
                cur = i
                while True:
                    nx = sole(cur)
                    if nx is None or nx in done:
                        break
                    nl = order[nx].layer
                    if isinstance(nl, E.BatchNormalization) and state < 1:
                        g["bn"], state = nl, 1
                    elif isinstance(nl, E.Add) and state < 2:
                        other = [index[id(s)] for s in order[nx].inputs if index[id(s)] != cur]
                        # the shortcut must already be materialised when this kernel runs
                        if len(other) != 1 or other[0] >= i:
                            break
                        g["res"], state = other[0], 2
                    elif isinstance(nl, E.Lambda) and state == 2 and g["res_mul"] == 1.0:
                        g["res_mul"] = float(nl.multiplier)
                    elif _act_of(nl) is not None and state < 3 and _act_of(nl)[0] in ("quant", "binary", "leaky", "linear"):
                        g["act"], state = _act_of(nl), 3
                    elif isinstance(nl, E.MaxPooling2D) and state < 4 and g["res"] is None:
                        g["pool"], state = True, 4
                    else:
                        break
                    done.add(nx)
                    cur = nx
                self.steps.append(_Step("conv", cur, **g))
            elif isinstance(lay, QDenseBase):
                g = dict(layer=lay, src=src[0], bn=None, softmax=(lay._act is not None and lay._act[0] == "softmax"))
                if lay._act is not None and not g["softmax"]:
                    raise ValueError("plan: Dense activation %r cannot be fused" % (lay._act,))
                cur = i
                nx = sole(cur)
                if nx is not None and isinstance(order[nx].layer, E.BatchNormalization) and not g["softmax"]:
                    g["bn"] = order[nx].layer
                    done.add(nx)
                    cur = nx
                self.steps.append(_Step("dense", cur, **g))
            elif isinstance(lay, E.Flatten):
                self.alias[i] = ("flatten", src[0])
            elif isinstance(lay, E.AveragePooling2D):
                self.alias[i] = ("avgpool", src[0], lay.pool_size)
            elif isinstance(lay, E.ZeroPadding2D):
                self.steps.append(_Step("layer", i, layer=lay, src=src))
            elif isinstance(lay, (E.BatchNormalization, E.Activation, E.LeakyReLU, E.MaxPooling2D, E.Add, E.Lambda)):
                # not adjacent to a conv/dense: run the stand-alone fp32 kernel
                self.steps.append(_Step("layer", i, layer=lay, src=src))
            else:
                raise ValueError("plan: layer %s (%s) is not supported on the accelerated path" % (lay.name, type(lay).__name__))
        # sanity: every kernel input is the model input or the output of an earlier step
        ready = {self.input_idx}
        for st in self.steps:
            needs = [st.src] if not isinstance(st.src, list) else list(st.src)
            if getattr(st, "res", None) is not None:
                needs.append(st.res)
            for nd in needs:
                while nd in self.alias:
                    nd = self.alias[nd][1]
                if nd not in ready:
                    raise ValueError("plan: step for %s reads tensor %d before it is produced" % (st.layer.name, nd))
            ready.add(st.out)
        # every avgpool alias must feed flatten -> dense
        for i, a in self.alias.items():
            if a[0] == "avgpool":
                for c in cons[i]:
                    if not (c in self.alias and self.alias[c][0] == "flatten"):
                        raise ValueError("plan: AveragePooling2D is only supported as AveragePooling2D -> Flatten -> Dense")

    # ------------------------------------------------------------------ run
    def _bn_dev(self, step, bn, dev):
        key = ("bn", str(dev))
        if key not in step.dev:
            inv, shift = bn.constants()
            step.dev[key] = (torch.from_numpy(inv).to(dev), torch.from_numpy(shift).to(dev))
        return step.dev[key]

    def _run_conv(self, st, env):
        lay = st.layer
        x = env[st.src]
        dev = x.data.device
        if x.kind == "b1" and lay.WEIGHT_KIND != "binary":
            x = K.QTensor("f32", x.to_float(), 1.0, x.channels)
        wfmt = L.WFMT_B1 if x.kind == "b1" else L.WFMT_I8
        if lay.WEIGHT_KIND == "float":
            # plain Conv2D ('float' networks): fp32 kernel x fp32 values; pixel levels become level / 255 first
            x = x if x.kind == "f32" else K.QTensor("f32", x.to_float(), 1.0, x.channels)
            wfmt = L.WFMT_F32
        _, _, _, wscale = lay.weight_mode()
        scale = K.acc_scale(x.scale if x.kind in ("u8", "i8") else 1.0, wscale)
        inv = shift = None
        if st.bn is not None:
            inv, shift = self._bn_dev(st, st.bn, dev)
        res = None
        if st.res is not None:
            res = env[st.res]
            if res.kind not in ("i8", "f32"):
                res = K.QTensor("f32", res.to_float(), 1.0, res.channels)
        act, abits, alpha = L.ACT_NONE, 0, 0.3
        if st.act is not None:
            if st.act[0] == "quant":
                act, abits = L.ACT_QUANT, int(st.act[1])
            elif st.act[0] == "binary":
                act = L.ACT_SIGN
            elif st.act[0] == "leaky":
                act, alpha = L.ACT_LEAKY, float(st.act[1])
        kw = dict(bias=lay.bias_tensor(dev), bn_inv=inv, bn_shift=shift, residual=res, res_mul=st.res_mul, abits=abits,
                  leaky_alpha=alpha, pool=2 if st.pool else 0)
        epi = K.make_epilogue(scale, act=act, **kw)
        if act == L.ACT_SIGN:
            # A +-1 map has two storage forms: bit-packed words (XNOR/popc kernels) or int8 levels.  When this layer
            # itself runs on the int8 tensor cores its output stays int8, so the next binary layer does too (measured:
            # tcgen05 kind::i8 on +-1 levels beats the CUDA-core XNOR-popc kernel by >10x, profiles/).
            epi8 = K.make_epilogue(scale, act=L.ACT_SIGN_I8, **kw)
            if K.conv2d_on_tensor_cores(x, lay.kernel_size[0], lay.kernel_size[1], lay.filters, lay.strides[0], epi8, self.impl):
                epi = epi8
        env[st.out] = K.conv2d(x, lay.packed_kernel(dev, wfmt), lay.kernel_size[0], lay.kernel_size[1], lay.filters,
                               lay.strides[0], epi, impl=self.impl, max_ctas=self.sm_share)
        self.launches += 1

    def _resolve_dense_input(self, idx, env):
        """Walk back through Flatten / AveragePooling2D aliases -> (QTensor [N, F], pool area)."""
        pool = 1
        cur = idx
        flat = False
        while cur in self.alias:
            a = self.alias[cur]
            if a[0] == "flatten":
                flat = True
            else:
                pool *= int(a[2][0]) * int(a[2][1])
                spatial = self.order[a[1]].shape
                if spatial[1] != a[2][0] or spatial[2] != a[2][1]:
                    raise ValueError("plan: AveragePooling2D must reduce the whole map (global pooling) to be fused")
            cur = a[1]
        return env[cur], pool, flat

    def _run_dense(self, st, env):
        lay = st.layer
        x, pool, _ = self._resolve_dense_input(st.src, env)
        dev = x.data.device
        if x.kind == "u8":
            x = K.QTensor("f32", x.to_float(), 1.0, x.channels)
        if x.kind == "b1" and (lay.WEIGHT_KIND != "binary" or x.channels % 32 != 0):
            x = K.QTensor("f32", x.to_float(), 1.0, x.channels)
        if x.kind == "i8" and x.channels % 4 != 0:
            x = K.QTensor("f32", x.to_float(), 1.0, x.channels)
        wfmt = L.WFMT_B1 if x.kind == "b1" else L.WFMT_I8
        if lay.WEIGHT_KIND == "float":
            x = x if x.kind == "f32" else K.QTensor("f32", x.to_float(), 1.0, x.channels)
            wfmt = L.WFMT_F32
        n = int(x.data.shape[0])
        ch = x.channels
        flat_shape = x.shape[1:]
        fin = int(np.prod(flat_shape))
        # pack the (in, units) kernel; under global average pooling the kernel has `ch` rows and is
        # replicated over the pooled pixels
        rows = lay.kernel.shape[0]
        avg_positions = 0
        if pool > 1 and x.kind == "f32" and ch <= 256:
            # fp32 maps: the kernel sums the pooled positions itself (one pass over x, a [units][ch] kernel)
            if rows != ch or fin != pool * ch:
                raise ValueError("plan: dense kernel %s does not match pooled features %d" % (lay.kernel.shape, ch))
            wp = lay.packed_kernel(dev, wfmt)
            avg_positions = pool
        elif pool > 1:
            if rows != ch or fin != pool * ch:
                raise ValueError("plan: dense kernel %s does not match pooled features %d" % (lay.kernel.shape, ch))
            key = "pool%d" % pool
            ck = (str(dev), wfmt, key)
            if ck not in st.dev:
                base = lay.packed_kernel(dev, wfmt)                 # [units][1][1][K]
                st.dev[ck] = base.reshape(lay.units, 1, -1).repeat(1, pool, 1).contiguous()
            wp = st.dev[ck]
        else:
            if rows != fin:
                raise ValueError("plan: dense kernel rows %d != features %d" % (rows, fin))
            wp = lay.packed_kernel(dev, wfmt)
        _, _, _, wscale = lay.weight_mode()
        xs = (x.scale if x.kind == "i8" else 1.0) / float(pool)
        scale = K.acc_scale(xs, wscale)
        inv = shift = None
        if st.bn is not None:
            inv, shift = self._bn_dev(st, st.bn, dev)
        epi = K.make_epilogue(scale, bias=lay.bias_tensor(dev), bn_inv=inv, bn_shift=shift)
        xin = K.QTensor(x.kind, x.data.reshape(n, -1), x.scale, fin if x.kind != "b1" else fin)
        # the caller's output buffer (Plan.forward(out=...): e.g. a peer-mapped block of the gathering rank) takes the
        # network output directly when this layer produces it
        dst = env.get("out_buffer") if st.out == self.output_idx else None
        out, logits = K.dense(xin, wp, lay.units, epi, softmax=st.softmax, want_logits=st.softmax, avg_positions=avg_positions,
                              out=dst, max_ctas=self.sm_share)
        if dst is not None:
            env["out_buffer_used"] = True
        env[st.out] = K.QTensor("f32", out, 1.0, lay.units)
        if logits is not None:
            env["logits"] = logits
        self.launches += 1

    def run_step(self, st, env):
        """Launch one fused step on the tensors in ``env`` (tensor index -> QTensor)."""
        if st.kind == "conv":
            self._run_conv(st, env)
        elif st.kind == "dense":
            self._run_dense(st, env)
        else:
            ins = [env[s] for s in st.src]
            y = st.layer.call(ins if isinstance(st.layer, E.Add) else ins[0])
            env[st.out] = y if isinstance(y, K.QTensor) else K.as_qtensor(y)
            self.launches += 1

    # ------------------------------------------------------------------ whole-network launch
    def _fused_chain(self):
        """The plan as (conv steps, dense step) when it is the plain models/vgg.py chain -- every conv 3x3 stride 1 fed by
        the previous one, quantised / binary activation, no residual, then Flatten -> Fc [-> BN] -- else None."""
        if len(self.steps) < 2 or self.steps[-1].kind != "dense" or any(st.kind != "conv" for st in self.steps[:-1]):
            return None
        if len(self.steps) - 1 > L.NET_MAX_CONVS:
            return None
        prev = self.input_idx
        for st in self.steps[:-1]:
            lay = st.layer
            if st.src != prev or st.res is not None or st.act is None or st.act[0] not in ("quant", "binary"):
                return None
            if lay.WEIGHT_KIND == "float":
                return None
            if tuple(lay.kernel_size) != (3, 3) or tuple(lay.strides) != (1, 1):
                return None
            prev = st.out
        dn = self.steps[-1]
        if dn.softmax or dn.out != self.output_idx:
            return None
        cur, flat = dn.src, False
        while cur in self.alias:
            if self.alias[cur][0] != "flatten":
                return None
            flat, cur = True, self.alias[cur][1]
        if cur != prev or not flat:
            return None
        return self.steps[:-1], dn

    def _fused_desc(self, x):
        """qnnb_vgg_desc of this plan for the uint8 batch ``x`` (QTensor), or None when the net is out of its scope."""
        chain = self._fused_chain()
        if chain is None or x.kind != "u8":
            return None
        convs, dn = chain
        dev = x.data.device
        n, h, w, cin = (int(v) for v in x.shape)
        scale_in, rows = x.scale, []
        for st in convs:
            lay = st.layer
            _, _, _, wscale = lay.weight_mode()
            inv = shift = None
            if st.bn is not None:
                inv, shift = self._bn_dev(st, st.bn, dev)
            if st.act[0] == "quant":
                act, abits, scale_out = L.ACT_QUANT, int(st.act[1]), 1.0 / float(1 << (int(st.act[1]) - 1))
            else:
                act, abits, scale_out = L.ACT_SIGN_I8, 0, 1.0
            epi = K.make_epilogue(K.acc_scale(scale_in, wscale), bias=lay.bias_tensor(dev), bn_inv=inv, bn_shift=shift,
                                  act=act, abits=abits, pool=2 if st.pool else 0)
            rows.append((lay.filters, 2 if st.pool else 0, lay.packed_kernel(dev, L.WFMT_I8), epi))
            scale_in = scale_out
            if st.pool:
                h, w = h // 2, w // 2
        lay = dn.layer
        if h < 1 or w < 1 or lay.kernel.shape[0] != h * w * convs[-1].layer.filters:
            return None
        _, _, _, wscale = lay.weight_mode()
        inv = shift = None
        if dn.bn is not None:
            inv, shift = self._bn_dev(dn, dn.bn, dev)
        depi = K.make_epilogue(K.acc_scale(scale_in, wscale), bias=lay.bias_tensor(dev), bn_inv=inv, bn_shift=shift)
        d = K.vgg_desc(n, int(x.shape[1]), int(x.shape[2]), cin, rows, lay.units, lay.packed_kernel(dev, L.WFMT_I8), depi,
                       max_ctas=self.sm_share)
        return d if K.vgg_forward_supported(d) else None

    def fused_available(self, x) -> bool:
        """Would ``forward`` run this batch as ONE whole-network launch?"""
        x = K.as_qtensor(x)
        key = (x.kind,) + tuple(x.shape[1:])
        if key not in self._fused_ok:
            self._fused_ok[key] = self.fuse and self._fused_desc(x) is not None
        return self._fused_ok[key]

    def run(self, x, out=None, fuse=False) -> dict:
        """One forward over a device batch.  Returns the environment (tensor index -> QTensor).  ``fuse``: take the
        whole-network launch when the net qualifies (the environment then holds the input and the output only)."""
        self._sync_weights()
        env = {self.input_idx: K.as_qtensor(x)}
        if fuse and int(env[self.input_idx].data.shape[0]) > 0 and self.fused_available(env[self.input_idx]):
            xq = env[self.input_idx]
            d = self._fused_desc(xq)
            # the resident image depends on the weights and the map geometry, not on the batch: cached with the other
            # per-step device constants (dropped by close() when a layer's weights change)
            key = ("vggblob", str(xq.data.device)) + tuple(xq.shape[1:])
            dev0 = self.steps[0].dev
            if key not in dev0:
                dev0[key] = K.vgg_pack(d, xq.data.device)
            y = K.vgg_forward(d, dev0[key], xq.data, out=out)
            env[self.output_idx] = K.QTensor("f32", y, 1.0, int(d.units))
            self.launches += 1
            return env
        if out is not None:
            env["out_buffer"] = out
        for st in self.steps:
            self.run_step(st, env)
        if out is not None and not env.get("out_buffer_used"):
            raise ValueError("plan: the network output is not produced by a dense layer; out= is not supported here")
        return env

    def forward(self, x, return_logits=False, out=None):
        """``out``: optional [N, units] fp32 destination of the network output (a CUDA tensor or an
        ``_lib.DeviceBuffer``, e.g. ``sharding.PeerGather.block()``), written by the final dense kernel itself."""
        env = self.run(x, out=out, fuse=True)
        out = env[self.output_idx]
        out = out.data if out.kind == "f32" else out.to_float()
        if return_logits:
            return out, env.get("logits", out)
        return out

    # ------------------------------------------------------------------ pipelined host path
    # A host batch goes through one of PIPELINE_DEPTH slots: its own stream, a static device input, the whole
    # fused plan captured once as a CUDA graph, and a pinned result buffer.  H2D copy, graph replay and D2H
    # copy are enqueued back to back on the slot's stream, so consecutive batches overlap copy and compute.
    # Four slots: with three, the host has ~40 us per step to re-issue a slot before the copy engine runs dry (a 3 MB batch is
    # 62 us of PCIe time); four keep the link busy through ordinary host jitter (13.9 -> 14.8 M img/s end to end on cfg3).
    PIPELINE_DEPTH = 4

    def _capture(self, x_dev, st):
        """Capture one forward over the static input ``x_dev`` on stream ``st``.  Returns (graph | None, output).

        Object finalisers must not run while the stream is capturing: destroying an old ``torch.cuda.CUDAGraph``
        (or freeing its pool) issues CUDA calls that are illegal under a global-mode capture and invalidate it.
        So: collect garbage BEFORE the capture, keep the cyclic collector off during it, and capture in thread-local
        mode (other threads' CUDA calls do not touch this capture either).  If the capture still fails the slot runs
        the same launches eagerly -- slower on the host, identical results."""
        gc.collect()
        torch.cuda.synchronize()
        st.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.stream(st):
                with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
                    out_dev = self.forward(x_dev)
            return g, out_dev
        except Exception:
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            del g
            with torch.cuda.stream(st):
                out_dev = self.forward(x_dev)
            st.synchronize()
            return None, out_dev
        finally:
            if gc_was_on:
                gc.enable()

    def _slot(self, shape, dtype):
        key = (tuple(shape), dtype)
        ring = self._slots.setdefault(key, {"slots": [], "next": 0})
        if len(ring["slots"]) < self.PIPELINE_DEPTH:
            dev = torch.device("cuda", torch.cuda.current_device())
            st = torch.cuda.Stream()
            x_dev = torch.zeros(shape, dtype=dtype, device=dev)
            if not ring["slots"]:
                self.forward(x_dev)                     # packs weights / uploads constants outside the capture
                torch.cuda.synchronize()
            before = self.launches
            g, out_dev = self._capture(x_dev, st)
            per_forward = self.launches - before           # 1 when the whole net is one launch
            self.launches = before
            out_host = torch.empty(out_dev.shape, dtype=out_dev.dtype).pin_memory()
            ring["slots"].append({"stream": st, "x": x_dev, "graph": g, "out": out_dev, "host": out_host,
                                  "event": torch.cuda.Event(), "owner": None, "gen": 0, "launches": per_forward})
        slot = ring["slots"][ring["next"] % len(ring["slots"])] if len(ring["slots"]) == self.PIPELINE_DEPTH else ring["slots"][-1]
        ring["next"] += 1
        return slot

    def predict_async(self, x):
        """Enqueue one host batch (numpy array or CPU tensor, ideally pinned) and return a handle whose
        ``result()`` blocks until the logits are back in host memory.  PIPELINE_DEPTH batches are in flight at most:
        when the ring wraps around onto a slot whose handle has not been read, that handle is completed first (its
        result is copied out), so handles may be read late and in any order."""
        if not torch.cuda.is_available():
            raise RuntimeError("predict needs a CUDA device: this package has no CPU path")
        src = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x.contiguous()
        if src.is_cuda:
            raise ValueError("predict_async takes host batches; use predict() for device tensors")
        self._sync_weights()
        slot = self._slot(src.shape, src.dtype)
        if slot["owner"] is not None:
            slot["owner"]._settle()
        slot["gen"] += 1
        with torch.cuda.stream(slot["stream"]):
            slot["x"].copy_(src, non_blocking=True)
            if slot["graph"] is not None:
                slot["graph"].replay()
            else:
                slot["out"].copy_(self.forward(slot["x"]))
            slot["host"].copy_(slot["out"], non_blocking=True)
            slot["event"].record(slot["stream"])
        self.launches += slot["launches"]
        h = _Handle(slot)
        slot["owner"] = h
        return h

    def predict(self, x, batch_size=None, return_logits=False):
        as_numpy = isinstance(x, np.ndarray)
        if not return_logits and (as_numpy or (isinstance(x, torch.Tensor) and not x.is_cuda)) and int(x.shape[0]) > 0:
            n = int(x.shape[0])
            bs = int(batch_size) if batch_size else min(n, MAX_CHUNK)
            if n % bs == 0 or n < bs:
                handles, outs = [], []
                for s in range(0, n, bs):
                    handles.append(self.predict_async(x[s:s + bs]))
                    if len(handles) >= self.PIPELINE_DEPTH:
                        outs.append(handles.pop(0).result())
                outs.extend(h.result() for h in handles)
                out = torch.cat(outs) if len(outs) > 1 else outs[0]
                return out.numpy() if as_numpy else out
        host_tensor = (not as_numpy) and isinstance(x, torch.Tensor) and not x.is_cuda
        if as_numpy or host_tensor:
            if not torch.cuda.is_available():
                raise RuntimeError("predict needs a CUDA device: this package has no CPU path")
            src = torch.from_numpy(np.ascontiguousarray(x)) if as_numpy else x.contiguous()
            xt = src.cuda(non_blocking=True)          # asynchronous when the host buffer is pinned
        else:
            xt = x
        n = int(xt.shape[0])
        bs = int(batch_size) if batch_size else min(max(n, 1), MAX_CHUNK)
        outs, logs = [], []
        for s in range(0, n, bs):
            r = self.forward(xt[s:s + bs].contiguous(), return_logits=return_logits)
            if return_logits:
                outs.append(r[0]); logs.append(r[1])
            else:
                outs.append(r)
        if n == 0:
            units = self.order[self.output_idx].shape[-1]
            out = torch.empty((0, units), dtype=torch.float32, device=xt.device)
            logit = out
        else:
            out = torch.cat(outs) if len(outs) > 1 else outs[0]
            logit = (torch.cat(logs) if len(logs) > 1 else logs[0]) if return_logits else None
        if as_numpy or host_tensor:
            out = out.cpu()
            logit = logit.cpu() if logit is not None else None
            if as_numpy:
                out = out.numpy()
                logit = logit.numpy() if logit is not None else None
        return (out, logit) if return_logits else out
