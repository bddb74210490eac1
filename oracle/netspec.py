"""Neutral graph description of the reference networks (TEST INFRASTRUCTURE).

Restates, as a flat list of nodes, the graphs built by the reference's
``models/model_factory.py:18-72`` (network_type switch), ``models/vgg.py:5-44`` and
``models/resnet.py:15-147``.  Both oracles (``oracle.exact`` and ``oracle.refstate``)
evaluate this list; the product builds its own graph from its own layer objects, so a
structural mistake on either side shows up as a parity failure.

A node is a dict: ``{"op": str, "in": [node indices], ...attrs}``; node 0 is the input.
Weights are attached with :func:`set_weights` in Keras order
(conv/dense: ``[kernel, bias?]``; BatchNormalization: ``[gamma, beta, mean, var]`` --
reference ``models/model_factory.py:91``).
"""
from __future__ import annotations

import numpy as np

LEAKY_ALPHA = 0.3          # keras.layers.LeakyReLU() default (model_factory.py:27,34,43,54)

_KINDS = {
    # network_type -> (weight kind, activation kind)        model_factory.py:24-58
    "qnn": ("quantized", "leaky"),
    "full-qnn": ("quantized", "quant"),
    "bnn": ("binary", "leaky"),
    "qbnn": ("binary", "quant"),
    "full-bnn": ("binary", "binary"),
    "tnn": ("ternary", "leaky"),
    "qtnn": ("ternary", "quant"),
    "float": ("float", "leaky"),            # plain Conv2D / Dense, model_factory.py:24-27
}


def kinds_for(network_type: str):
    if network_type not in _KINDS:
        # same message family as model_factory.py:61 (full-tnn / float are outside the path)
        raise ValueError("wrong network type, the supported network types in this repo are "
                         "float, qnn, full-qnn, bnn and full-bnn")
    return _KINDS[network_type]


def glorot_multiplier(kh, kw, cin, cout):
    """kernel_lr_multiplier == 'Glorot' (layers/quantized_layers.py:126-136)."""
    base = kh * kw
    nb_input = int(cin * base)
    nb_output = int(cout * base)
    return np.float32(1.0 / np.sqrt(1.5 / (nb_input + nb_output)))


def _conv(nodes, src, wkind, cf, filters, ksize, stride, use_bias, cin):
    nodes.append({"op": "conv", "in": [src], "wkind": wkind, "nb": int(cf.wbits), "H": 1.0,
                  "filters": int(filters), "ksize": int(ksize), "stride": int(stride),
                  "use_bias": bool(use_bias), "cin": int(cin),
                  "klm": glorot_multiplier(ksize, ksize, cin, filters)})
    return len(nodes) - 1


def _bn(nodes, src, eps, ch):
    nodes.append({"op": "bn", "in": [src], "eps": float(eps), "ch": int(ch)})
    return len(nodes) - 1


def _act(nodes, src, akind, cf):
    nodes.append({"op": "act", "in": [src], "akind": akind, "abits": int(cf.abits)})
    return len(nodes) - 1


def _dense_nb(cf):
    # QuantizedDense is built with nb=cf.abits, not wbits (model_factory.py:31)
    return int(cf.abits)


def vgg_spec(cf):
    """models/vgg.py:5-44 with the factories of model_factory.py:24-58."""
    wkind, akind = kinds_for(cf.network_type)
    nodes = [{"op": "input", "in": [], "shape": (cf.dim, cf.dim, cf.channels)}]
    cur, ch = 0, cf.channels
    plan = [(cf.nla, cf.nfa), (cf.nlb, cf.nfb), (cf.nlc, cf.nfc)]
    for bi, (nl, nf) in enumerate(plan):
        # block A always has its first conv (vgg.py:15) plus nla-1 more (vgg.py:19-22)
        count = max(nl, 1) if bi == 0 else nl
        for _ in range(count):
            cur = _conv(nodes, cur, wkind, cf, nf, 3, 1, True, ch)
            ch = nf
            cur = _bn(nodes, cur, 1e-4, ch)          # vgg.py:16 epsilon=0.0001
            cur = _act(nodes, cur, akind, cf)
        nodes.append({"op": "maxpool", "in": [cur]})   # vgg.py:23,30,37
        cur = len(nodes) - 1
    nodes.append({"op": "flatten", "in": [cur]})
    cur = len(nodes) - 1
    side = cf.dim // 8
    feat = side * side * ch
    nodes.append({"op": "dense", "in": [cur], "wkind": wkind, "nb": _dense_nb(cf), "H": 1.0,
                  "units": int(cf.classes), "use_bias": True, "fin": int(feat), "softmax": False})
    cur = len(nodes) - 1
    cur = _bn(nodes, cur, 1e-4, cf.classes)            # vgg.py:42
    return nodes


def resnet_spec(cf, use_bias=False, half=True):
    """models/resnet.py:15-147.  ``use_bias=True, half=False`` is the older revision the
    shipped ``results/RESNET3/weights_*.hdf5`` were trained with (SURVEY.md finding 7)."""
    wkind, akind = kinds_for(cf.network_type)
    pf = int(getattr(cf, "pfilt", 1))
    nodes = [{"op": "input", "in": [], "shape": (cf.dim, cf.dim, cf.channels)}]
    cur, ch = 0, cf.channels
    if cf.dataset in ("MNIST", "FASHION"):
        nodes.append({"op": "zeropad", "in": [cur], "pad": 2})      # resnet.py:101-102
        cur = len(nodes) - 1
    eps = 1e-3                                                      # Keras BN default (resnet.py:61)
    nf = 16
    cur = _conv(nodes, cur, wkind, cf, nf * pf, 3, 1, use_bias, ch)  # stem (resnet.py:105)
    ch = nf * pf
    cur = _act(nodes, _bn(nodes, cur, eps, ch), akind, cf)
    for stack in range(3):
        for blk in range(int(cf.nres)):
            stride = 2 if (stack > 0 and blk == 0) else 1
            x = cur
            y = _conv(nodes, x, wkind, cf, nf * pf, 3, stride, use_bias, ch)
            y = _act(nodes, _bn(nodes, y, eps, nf * pf), akind, cf)
            y = _conv(nodes, y, wkind, cf, nf * pf, 3, 1, use_bias, nf * pf)
            y = _bn(nodes, y, eps, nf * pf)
            if stack > 0 and blk == 0:
                # 1x1 stride-2 projection, no BN, no activation (resnet.py:121-126)
                x = _conv(nodes, x, wkind, cf, nf * pf, 1, stride, use_bias, ch)
            ch = nf * pf
            nodes.append({"op": "add", "in": [x, y], "mul": 0.5 if half else 1.0})  # resnet.py:127-128
            cur = _act(nodes, len(nodes) - 1, akind, cf)
        nf *= 2
    nodes.append({"op": "avgpool", "in": [cur], "size": 8})           # resnet.py:134
    nodes.append({"op": "flatten", "in": [len(nodes) - 1]})
    nodes.append({"op": "dense", "in": [len(nodes) - 1], "wkind": wkind, "nb": _dense_nb(cf), "H": 1.0,
                  "units": int(cf.classes), "use_bias": bool(use_bias), "fin": int(ch),
                  "softmax": True})                                    # resnet.py:136-140
    return nodes


def build_spec(cf, **kw):
    if cf.architecture == "VGG":
        return vgg_spec(cf)
    if cf.architecture == "RESNET":
        return resnet_spec(cf, **kw)
    raise ValueError("Error: type " + str(cf.architecture) + " is not supported")


def weight_shapes(nodes):
    """Keras-order list of (node index, name, shape)."""
    out = []
    for i, nd in enumerate(nodes):
        if nd["op"] == "conv":
            k = nd["ksize"]
            out.append((i, "kernel", (k, k, nd["cin"], nd["filters"])))
            if nd["use_bias"]:
                out.append((i, "bias", (nd["filters"],)))
        elif nd["op"] == "dense":
            out.append((i, "kernel", (nd["fin"], nd["units"])))
            if nd["use_bias"]:
                out.append((i, "bias", (nd["units"],)))
        elif nd["op"] == "bn":
            for nm in ("gamma", "beta", "mean", "var"):
                out.append((i, nm, (nd["ch"],)))
    return out


def set_weights(nodes, weights):
    shapes = weight_shapes(nodes)
    if len(shapes) != len(weights):
        raise ValueError("expected %d weight arrays, got %d" % (len(shapes), len(weights)))
    for (i, nm, shp), w in zip(shapes, weights):
        w = np.asarray(w, dtype=np.float32)
        if tuple(w.shape) != tuple(shp):
            raise ValueError("node %d %s: expected shape %s, got %s" % (i, nm, shp, w.shape))
        nodes[i][nm] = w
    return nodes


def random_weights(nodes, seed=42, bias_range=0.0, bn="identity"):
    """Seeded synthetic weights: kernels U(-1,1) (H=1,
    layers/quantized_layers.py:140), bias zeros or U(-r,r), BN either the literal Keras
    init or a 'spread' setting with non-trivial gamma/beta/mean/var (scaled so that the
    pre-activation distribution is not saturated)."""
    rng = np.random.default_rng(seed)
    out = []
    fan = {}
    for i, nd in enumerate(nodes):
        if nd["op"] == "conv":
            fan[i] = nd["ksize"] * nd["ksize"] * nd["cin"]
        elif nd["op"] == "dense":
            fan[i] = nd["fin"]
    last_fan = 1
    for (i, nm, shp) in weight_shapes(nodes):
        nd = nodes[i]
        if nm == "kernel":
            out.append(rng.uniform(-1.0, 1.0, size=shp).astype(np.float32))
            last_fan = fan[i]
        elif nm == "bias":
            if bias_range > 0:
                out.append(rng.uniform(-bias_range, bias_range, size=shp).astype(np.float32))
            else:
                out.append(np.zeros(shp, np.float32))
        elif bn == "identity":
            out.append({"gamma": np.ones, "beta": np.zeros, "mean": np.zeros, "var": np.ones}[nm](shp, np.float32))
        else:
            # 'spread': variance of a sum of `fan` products of U(-1,1) weights (var 1/3) with
            # activations of mean-square ~1/3 is ~ fan*0.11; keeps most outputs un-saturated.
            sd2 = max(last_fan * 0.11, 1e-3)
            if nm == "gamma":
                v = rng.uniform(0.3, 0.9, size=shp) * rng.choice([1.0, 1.0, 1.0, -1.0], size=shp)
            elif nm == "beta":
                v = rng.uniform(-0.2, 0.2, size=shp)
            elif nm == "mean":
                v = rng.uniform(-0.2, 0.2, size=shp) * np.sqrt(sd2)
            else:
                v = rng.uniform(0.5, 1.5, size=shp) * sd2
            out.append(v.astype(np.float32))
    return out
