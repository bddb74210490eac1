#!/usr/bin/env python
"""Generate the golden fixtures by running the REFERENCE'S OWN python sources.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

TensorFlow/Keras cannot be installed here, so ``tests/golden/keras_shim`` (a torch-CPU stand-in for the
handful of Keras/TF symbols the reference imports) is put first on ``sys.path`` and the reference modules
``layers.quantized_ops / binary_ops / ternary_ops``, ``layers.*_layers`` and ``models.model_factory /
vgg / resnet`` are imported UNMODIFIED from /root/reference.  What they compute on seeded inputs is stored
in small ``.npz`` files next to this script; ``tests/test_oracle_golden.py`` pins both oracles to them.

One adaptation is made after the models are built: ``kernel_lr_multiplier`` (a numpy float32 produced by
``layers/quantized_layers.py:133-136``) is converted to a python float holding the same value.  Under the
numpy 1.x the reference ran on, ``1./np.float32`` promotes to float64, so the scaling-identity constants
(``quantized_layers.py:167-180``) are float64 scalars that TF casts to fp32; numpy 2 would keep them in
float32 and hand torch a numpy scalar.  The conversion reproduces the original promotion exactly.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("QNNB_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "keras_shim"))

import torch  # noqa: E402
import keras  # noqa: E402  (the shim)
from keras import layers as KL, initializers as KI  # noqa: E402

from layers import quantized_ops as rq, binary_ops as rb, ternary_ops as rt  # noqa: E402  (reference)
from layers.quantized_layers import QuantizedConv2D, QuantizedDense  # noqa: E402
from layers.binary_layers import BinaryConv2D, BinaryDense  # noqa: E402
from layers.ternary_layers import TernaryConv2D, TernaryDense  # noqa: E402
from models.model_factory import build_model  # noqa: E402

from oracle import netspec  # noqa: E402

F32 = np.float32
assert keras.__version__.endswith("shim")


def T(a):
    return torch.as_tensor(np.asarray(a, F32))


def special_values():
    t = F32(2.0 ** -24)
    return np.array([0.0, -0.0, t, np.nextafter(t, F32(1)), -t, 0.5, -0.5, 0.25, -0.25, 0.125, 0.0625, 0.1875, -0.1875, 1.0, -1.0,
                     0.99999, 0.9375, 0.96875, 3 / 256, 5 / 256, -3 / 256, 1.5 / 128, 2.5 / 128, 0.75, -0.75, 1.7, -2.3], F32)


def gen_ops():
    rng = np.random.default_rng(11)
    x = np.concatenate([special_values(), rng.uniform(-1.2, 1.2, size=4096 - 27).astype(F32)])
    out = {"x": x}
    for nb in (2, 4, 8):
        out["quantize_%d" % nb] = rq.quantize(T(x), nb=nb).numpy()
        out["quantized_tanh_%d" % nb] = rq.quantized_tanh(T(x), nb=nb).numpy()
    out["round_through"] = rq.round_through(T(x * 8)).numpy()
    out["binary_tanh"] = rb.binary_tanh(T(x)).numpy()
    out["binarize_1"] = rb.binarize(T(x), H=1).numpy()
    out["binarize_075"] = rb.binarize(T(x), H=0.75).numpy()
    w = rng.uniform(-1, 1, size=(3, 3, 8, 16)).astype(F32)
    out["w"] = w
    out["ternarize_1"] = rt.ternarize(T(w), H=1).numpy()
    out["_ternarize_1"] = rt._ternarize(T(w), H=1).numpy()
    out["ternarize_05"] = rt.ternarize(T(w * 0.5), H=0.5).numpy()
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)
    print("ops.npz", len(out))


def pyfloat_multipliers(layers):
    for l in layers:
        if hasattr(l, "kernel_lr_multiplier") and isinstance(l.kernel_lr_multiplier, np.floating):
            l.kernel_lr_multiplier = float(l.kernel_lr_multiplier)


def gen_layers():
    rng = np.random.default_rng(12)
    out = {}
    xq = (rng.integers(-8, 8, size=(2, 8, 8, 16)).astype(F32) / 8).astype(F32)
    xf = rng.normal(0, 1, size=(2, 8, 8, 16)).astype(F32)
    out["xq"], out["xf"] = xq, xf
    cases = [("qconv_s1", QuantizedConv2D, dict(nb=4), 1, 3), ("qconv_s2", QuantizedConv2D, dict(nb=4), 2, 3),
             ("qconv_1x1s2", QuantizedConv2D, dict(nb=2), 2, 1), ("qconv8", QuantizedConv2D, dict(nb=8), 1, 3),
             ("bconv", BinaryConv2D, dict(), 1, 3), ("tconv", TernaryConv2D, dict(), 1, 3)]
    for name, cls, kw, stride, k in cases:
        lay = cls(filters=24, kernel_size=(k, k), strides=(stride, stride), padding="same", H=1., **kw)
        lay(T(xq))                                   # build
        kernel = rng.uniform(-1, 1, size=tuple(lay.weights[0].shape)).astype(F32)
        bias = rng.uniform(-0.3, 0.3, size=24).astype(F32)
        lay.set_weights([kernel, bias])
        out[name + "_klm"] = np.float32(lay.kernel_lr_multiplier)
        pyfloat_multipliers([lay])
        out[name + "_kernel"], out[name + "_bias"] = kernel, bias
        out[name + "_yq"] = lay(T(xq)).numpy()
        out[name + "_yf"] = lay(T(xf)).numpy()
    xd = (rng.integers(-8, 8, size=(5, 64)).astype(F32) / 8).astype(F32)
    out["xd"] = xd
    for name, cls, kw in [("qdense", QuantizedDense, dict(nb=4)), ("bdense", BinaryDense, dict()), ("tdense", TernaryDense, dict())]:
        lay = cls(10, H=1., **kw)
        lay(T(xd))
        kernel = rng.uniform(-1, 1, size=(64, 10)).astype(F32)
        bias = rng.uniform(-0.3, 0.3, size=10).astype(F32)
        lay.set_weights([kernel, bias])
        out[name + "_kernel"], out[name + "_bias"] = kernel, bias
        out[name + "_y"] = lay(T(xd)).numpy()
    np.savez_compressed(os.path.join(HERE, "layers.npz"), **out)
    print("layers.npz", len(out))


MODEL_CASES = {
    # name -> (cf kwargs, batch, weight seed, bn setting)
    "vgg_cfg1_w2a2": (dict(network_type='full-qnn', wbits=2, abits=2, architecture='VGG', dataset='MNIST', dim=28, channels=1,
                           nla=1, nfa=64, nlb=1, nfb=64, nlc=1, nfc=64), 8, 5, "spread"),
    "vgg_cfg3_w4a4": (dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG'), 4, 6, "spread"),
    "vgg_cfg3_identity": (dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG'), 4, 6, "identity"),
    "vgg_fullbnn": (dict(network_type='full-bnn', architecture='VGG'), 4, 7, "spread"),
    "vgg_qnn_w4": (dict(network_type='qnn', wbits=4, abits=4, architecture='VGG'), 4, 8, "spread"),
    "resnet1_fullqnn_w4a4": (dict(network_type='full-qnn', wbits=4, abits=4, architecture='RESNET', nres=1), 4, 9, "spread"),
    "resnet1_tnn": (dict(network_type='tnn', wbits=4, abits=4, architecture='RESNET', nres=1), 4, 10, "spread"),
    "resnet1_qbnn_a4": (dict(network_type='qbnn', wbits=4, abits=4, architecture='RESNET', nres=1), 4, 11, "spread"),
    # network_type 'float': keras Conv2D / Dense / LeakyReLU (model_factory.py:24-27)
    "vgg_float": (dict(network_type='float', architecture='VGG'), 4, 12, "spread"),
    "resnet1_float": (dict(network_type='float', architecture='RESNET', nres=1), 4, 13, "spread"),
}


def make_cf(**kw):
    base = dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG', dataset='CIFAR-10', dim=32, channels=3,
                classes=10, nla=1, nfa=64, nlb=1, nfb=128, nlc=1, nfc=256, nres=3, pfilt=1,
                kernel_initializer='glorot_uniform', kernel_regularizer=0.)
    base.update(kw)
    return types.SimpleNamespace(**base)


def fix_vgg_fc_quirk(cf):
    """models/vgg.py:41 calls ``Fc(cf.classes)`` positionally, which model_factory.py:31's ``lambda **kwargs``
    cannot accept for qnn/full-qnn (a TypeError in the shipped code, SURVEY.md finding 6).  The fixture needs the
    graph the author intended, so the one factory is wrapped to forward the positional argument."""
    import models.model_factory as mf
    import models.vgg as vgg_mod
    orig = vgg_mod.Vgg

    def patched(Conv, Act, Fc, cf_):
        def fc(*a, **k):
            if a:
                k["units"] = a[0]
            return Fc(**k)
        try:
            return orig(Conv, Act, Fc, cf_)
        except TypeError:
            KL.reset_names()
            return orig(Conv, Act, fc, cf_)
    mf.Vgg = patched


def gen_models(only=None):
    import models.model_factory  # noqa: F401
    fix_vgg_fc_quirk(None)
    for name, (cfkw, batch, seed, bn) in MODEL_CASES.items():
        if only and name not in only:
            continue
        KL.reset_names()
        KI.seed(seed)
        cf = make_cf(**cfkw)
        model = build_model(cf)
        nodes = netspec.build_spec(cf)
        netspec.set_weights(nodes, netspec.random_weights(nodes, seed=seed, bias_range=0.1, bn=bn))
        # copy the seeded weights into the reference model, matching layers by creation order (= name suffix)
        lin = [l for l in model.layers if hasattr(l, "kernel_lr_multiplier") or type(l).__name__ in ("Conv2D", "Dense")]
        bns = [l for l in model.layers if type(l).__name__ == "BatchNormalization"]
        lin.sort(key=lambda l: int(l.name.rsplit("_", 1)[1]) + (10000 if "dense" in l.name else 0))
        bns.sort(key=lambda l: int(l.name.rsplit("_", 1)[1]))
        spec_lin = [nd for nd in nodes if nd["op"] in ("conv", "dense")]
        spec_bn = [nd for nd in nodes if nd["op"] == "bn"]
        assert len(lin) == len(spec_lin) and len(bns) == len(spec_bn), (name, len(lin), len(spec_lin), len(bns), len(spec_bn))
        for lay, nd in zip(lin, spec_lin):
            lay.set_weights([nd["kernel"]] + ([nd["bias"]] if nd["use_bias"] else []))
            if hasattr(lay, "kernel_lr_multiplier"):
                assert abs(float(lay.kernel_lr_multiplier) - float(nd.get("klm", lay.kernel_lr_multiplier))) < 1e-3 or nd["op"] == "dense"
        for lay, nd in zip(bns, spec_bn):
            lay.set_weights([nd["gamma"], nd["beta"], nd["mean"], nd["var"]])
        pyfloat_multipliers(model.layers)
        x8 = np.random.default_rng(1000 + seed).integers(0, 256, size=(batch, cf.dim, cf.dim, cf.channels), dtype=np.uint8)
        x = x8.astype('float32') / 255                     # utils/load_data.py:40
        taps = []
        y = model.predict(x, taps=taps)
        out = {"x8": x8, "y": y, "param_count": np.int64(sum(int(np.prod(w.shape)) for w in model.get_weights())),
               "layer_names": np.array([n for n, _ in taps]),
               "tap_sum": np.array([float(np.float64(v.astype(np.float64).sum())) for _, v in taps]),
               "tap_abs": np.array([float(np.float64(np.abs(v.astype(np.float64)).sum())) for _, v in taps])}
        # full tensors of a few interior taps (first activation, last pooled/added map)
        act_names = [n for n, _ in taps if n.startswith("activation") or n.startswith("leaky")]
        for n, v in taps:
            if act_names and n in (act_names[0], act_names[-1]):
                out["tap_" + n] = v
        np.savez_compressed(os.path.join(HERE, "model_%s.npz" % name), **out)
        print("model_%s.npz" % name, y.shape, "params", int(out["param_count"]))


if __name__ == "__main__":
    if len(sys.argv) > 1:                     # python make_golden.py model_name ... : only these model fixtures
        gen_models(set(sys.argv[1:]))
    else:
        gen_ops()
        gen_layers()
        gen_models()
