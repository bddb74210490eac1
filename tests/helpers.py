"""Shared test helpers: single-layer oracle built from oracle.exact pieces, config bags."""
import types

import numpy as np

from oracle import exact

F32 = np.float32


def make_cf(**kw):
    base = dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG', dataset='CIFAR-10', dim=32,
                channels=3, classes=10, nla=1, nfa=64, nlb=1, nfb=128, nlc=1, nfc=256, nres=3, pfilt=1,
                kernel_initializer='glorot_uniform', kernel_regularizer=0.)
    base.update(kw)
    return types.SimpleNamespace(**base)


CONFIGS = {
    # BASELINE.json configs (batch sizes reduced where noted by the tests)
    "cfg1": dict(network_type='full-qnn', wbits=2, abits=2, architecture='VGG', dataset='MNIST', dim=28, channels=1,
                 nla=1, nfa=64, nlb=1, nfb=64, nlc=1, nfc=64),
    "cfg2": dict(network_type='full-bnn', architecture='VGG'),
    "cfg3": dict(network_type='full-qnn', wbits=4, abits=4, architecture='VGG'),
    "cfg4": dict(network_type='full-qnn', wbits=8, abits=8, architecture='VGG', nla=3, nfa=256, nlb=3, nfb=256, nlc=3, nfc=256),
    "cfg5": dict(network_type='qnn', wbits=4, abits=4, architecture='RESNET', nres=10),
}


def oracle_layer(x, in_kind, x_scale, kernel, wkind, nb, H, stride, bias=None, bn=None, eps=1e-4,
                 residual=None, res_mul=1.0, act=None, abits=4, alpha=0.3, pool=False, dense=False):
    """One fused conv/dense step evaluated with the exact oracle's primitives.
    x: integer levels (u8/i8), +-1 levels (b1) or fp32 values; returns (output array, int accumulators|None).
    Output: int levels for act 'quant', +-1 for 'binary', fp32 otherwise."""
    lv, ws = exact.weight_levels(kernel, wkind, nb, H)
    qt = exact.QT("f32", np.asarray(x, F32)) if in_kind == "f32" else exact.QT(in_kind, np.asarray(x).astype(np.int64), x_scale)
    c, iacc = exact.linear(qt, lv, ws, stride, dense=dense)
    if bias is not None:
        c = (c + np.asarray(bias, F32)).astype(F32)
    if bn is not None:
        inv, shift = exact.bn_constants(*bn, eps)
        c = (c * inv).astype(F32)
        c = (c + shift).astype(F32)
    if residual is not None:
        c = ((np.asarray(residual, F32) + c).astype(F32) * F32(res_mul)).astype(F32)
    if act == "quant":
        out = exact.act_quant_levels(c, abits)
    elif act == "binary":
        out = exact.act_binary_levels(c)
    elif act == "leaky":
        out = exact.leaky(c, alpha)
    else:
        out = c
    if pool:
        out = exact.maxpool2(out)
    return out, iacc


def unpack_bits(words, channels):
    """uint32/int32 words [..., W] -> +-1 levels [..., channels]."""
    w = np.asarray(words).astype(np.uint32)
    bits = ((w[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(w.shape[:-1] + (-1,))
    return bits[..., :channels].astype(np.int32) * 2 - 1


def assign_weights_from_spec(model, nodes):
    """Copy oracle-spec weights into the product model by walking both graphs in parallel."""
    from qnn_b200 import engine as E
    from qnn_b200.layers._base import QConv2DBase, QDenseBase
    order = E.topo_order(model._graph()[1])
    # map product tensors to spec nodes by structural matching along creation order of weighted layers
    spec_convs = [nd for nd in nodes if nd["op"] in ("conv", "dense")]
    spec_bns = [nd for nd in nodes if nd["op"] == "bn"]
    prod_lin = [t.layer for t in order if isinstance(t.layer, (QConv2DBase, QDenseBase))]
    prod_bn = [t.layer for t in order if isinstance(t.layer, E.BatchNormalization)]
    # creation order == name order in both builders
    prod_lin.sort(key=lambda l: int(l.name.rsplit("_", 1)[1]) + (10000 if isinstance(l, QDenseBase) else 0))
    prod_bn.sort(key=lambda l: int(l.name.rsplit("_", 1)[1]))
    assert len(prod_lin) == len(spec_convs) and len(prod_bn) == len(spec_bns)
    for lay, nd in zip(prod_lin, spec_convs):
        assert tuple(lay.kernel.shape) == tuple(nd["kernel"].shape), (lay.name, lay.kernel.shape, nd["kernel"].shape)
        lay.set_weights([nd["kernel"]] + ([nd["bias"]] if nd["use_bias"] else []))
    for lay, nd in zip(prod_bn, spec_bns):
        lay.set_weights([nd["gamma"], nd["beta"], nd["mean"], nd["var"]])
    model._invalidate()


def spec_weights_from_model(model, nodes):
    """The inverse of assign_weights_from_spec: copy the product model's weights into an oracle spec."""
    from qnn_b200 import engine as E
    from qnn_b200.layers._base import QConv2DBase, QDenseBase
    order = E.topo_order(model._graph()[1])
    spec_convs = [nd for nd in nodes if nd["op"] in ("conv", "dense")]
    spec_bns = [nd for nd in nodes if nd["op"] == "bn"]
    prod_lin = [t.layer for t in order if isinstance(t.layer, (QConv2DBase, QDenseBase))]
    prod_bn = [t.layer for t in order if isinstance(t.layer, E.BatchNormalization)]
    prod_lin.sort(key=lambda l: int(l.name.rsplit("_", 1)[1]) + (10000 if isinstance(l, QDenseBase) else 0))
    prod_bn.sort(key=lambda l: int(l.name.rsplit("_", 1)[1]))
    assert len(prod_lin) == len(spec_convs) and len(prod_bn) == len(spec_bns)
    for lay, nd in zip(prod_lin, spec_convs):
        assert tuple(lay.kernel.shape) == tuple(weight_shape(nd))
        nd["kernel"] = lay.kernel
        if nd["use_bias"]:
            nd["bias"] = lay.bias
    for lay, nd in zip(prod_bn, spec_bns):
        nd["gamma"], nd["beta"], nd["mean"], nd["var"] = lay.get_weights()
    return nodes


def weight_shape(nd):
    if nd["op"] == "conv":
        return (nd["ksize"], nd["ksize"], nd["cin"], nd["filters"])
    return (nd["fin"], nd["units"])
