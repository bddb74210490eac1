"""CPU tests of the host-side logic: model_factory switch, graph lowering / fusion plan, weight plumbing,
constructor surface, error behaviour, batch sharding over gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

import qnn_b200 as q
from helpers import make_cf, CONFIGS

F32 = np.float32


def test_network_type_switch_and_errors():
    for nt in ("float", "qnn", "full-qnn", "bnn", "qbnn", "full-bnn", "tnn", "qtnn"):
        for arch in ("VGG", "RESNET"):
            q.reset_names()
            m = q.build_model(make_cf(network_type=nt, architecture=arch, nres=1))
            assert m.output_shape == (None, 10)
    # 'float' is the plain keras Conv2D / Dense (model_factory.py:24-27): same graph, keras layer names, no quantiser fields
    q.reset_names()
    m = q.build_model(make_cf(network_type="float", architecture="RESNET", nres=3, kernel_regularizer=1e-4))
    assert m.layers[0].name == "conv2d_1" and m.layers[-1].name == "dense_1"
    assert "H" not in m.layers[0].get_config()
    with pytest.raises(ValueError, match="wrong network type"):
        q.build_model(make_cf(network_type="nope"))
    with pytest.raises(ValueError, match="is not supported"):
        q.build_model(make_cf(architecture="ALEXNET"))
    with pytest.raises(NotImplementedError):
        q.build_model(make_cf(network_type="full-tnn"))


def test_vgg_structure_and_param_counts():
    q.reset_names()
    m = q.build_model(make_cf(**CONFIGS["cfg3"]))
    names = [l.name for l in m.layers]
    assert names[:4] == ["quantized_conv2d_1", "batch_normalization_1", "activation_1", "max_pooling2d_1"]
    assert names[-3:] == ["flatten_1", "quantized_dense_1", "batch_normalization_4"]
    assert m.count_params() == 413618                     # kernels 411,328 + biases 458 + BN 4*458
    assert [l.output_shape for l in m.layers if "max_pooling" in l.name] == [(None, 16, 16, 64), (None, 8, 8, 128), (None, 4, 4, 256)]
    q.reset_names()
    m1 = q.build_model(make_cf(**CONFIGS["cfg1"]))
    assert m1.layers[-2].kernel.shape == (576, 10)        # 28 -> 14 -> 7 -> 3 (valid pooling), 3*3*64
    q.reset_names()
    m4 = q.build_model(make_cf(**CONFIGS["cfg4"]))
    assert sum(int(np.prod(l.kernel.shape)) for l in m4.layers if hasattr(l, "kernel")) == 4766464
    # QuantizedDense is built with nb = abits, convs with nb = wbits (model_factory.py:30-31)
    q.reset_names()
    mm = q.build_model(make_cf(network_type="full-qnn", wbits=8, abits=2))
    assert mm.layers[0].nb == 8 and mm.layers[-2].nb == 2


def test_resnet_structure_matches_reference_logs():
    for nres, want in ((3, 274442), (5, 470218), (10, 959658)):        # results/RESNET*/*.out (biased revision)
        q.reset_names()
        m = q.build_model(make_cf(architecture="RESNET", nres=nres), legacy_resnet=True)
        assert m.count_params() == want
    q.reset_names()
    m = q.build_model(make_cf(architecture="RESNET", nres=10, network_type="qnn"))
    convs = [l for l in m.layers if "conv2d" in l.name]
    assert len(convs) == 63 and m.depth == 62
    assert sum(int(np.prod(l.kernel.shape)) for l in m.layers if hasattr(l, "kernel")) == 948272
    assert all(not l.use_bias for l in convs)


def test_glorot_multipliers_match_stored_checkpoint_configs():
    # results/RESNET3/weights_44.hdf5 model_config: 10.677078 (3x3, 3->16) and 13.856406 (3x3, 16->16)
    q.reset_names()
    m = q.build_model(make_cf(architecture="RESNET", nres=3))
    convs = [l for l in m.layers if "conv2d" in l.name]
    assert abs(float(convs[0].kernel_lr_multiplier) - 10.677078) < 1e-5
    c16 = [l for l in convs if l.kernel.shape == (3, 3, 16, 16)][0]
    assert abs(float(c16.kernel_lr_multiplier) - 13.856406) < 1e-5
    cfg = convs[0].get_config()
    assert cfg["H"] == 1.0 and cfg["filters"] == 16 and cfg["padding"] == "same" and cfg["use_bias"] is False


def test_fusion_plan_of_vgg_and_resnet():
    q.reset_names()
    m = q.build_model(make_cf(**CONFIGS["cfg3"]))
    p = m.plan()
    assert [s.kind for s in p.steps] == ["conv", "conv", "conv", "dense"]
    assert all(s.bn is not None and s.pool and s.act == ("quant", 4) for s in p.steps[:3])
    assert p.steps[3].bn is not None and not p.steps[3].softmax
    q.reset_names()
    r = q.build_model(make_cf(architecture="RESNET", nres=2, network_type="full-qnn"))
    pr = r.plan()
    convs = [s for s in pr.steps if s.kind == "conv"]
    assert len(convs) == 1 + 3 * 2 * 2 + 2                   # stem + 12 block convs + 2 projections
    with_res = [s for s in convs if s.res is not None]
    assert len(with_res) == 6 and all(s.res_mul == 0.5 and s.bn is not None and s.act == ("quant", 4) for s in with_res)
    proj = [s for s in convs if s.layer.kernel_size == (1, 1)]
    assert len(proj) == 2 and all(s.bn is None and s.act is None and s.res is None for s in proj)
    assert pr.steps[-1].kind == "dense" and pr.steps[-1].softmax
    q.reset_names()
    old = q.build_model(make_cf(architecture="RESNET", nres=1), legacy_resnet=True)
    assert all(s.res_mul == 1.0 for s in old.plan().steps if s.kind == "conv" and s.res is not None)


def test_activation_probe_recognises_user_closures():
    from qnn_b200.layers.quantized_ops import quantized_tanh
    from qnn_b200.layers.binary_ops import binary_tanh
    assert q.Activation(lambda x: quantized_tanh(x, nb=3)).act_spec() == ("quant", 3)
    assert q.Activation(binary_tanh).act_spec() == ("binary",)
    with pytest.raises(ValueError):
        q.Activation(lambda x: x * 2)


def test_weights_roundtrip_and_shape_errors():
    q.reset_names()
    m = q.build_model(make_cf(**CONFIGS["cfg1"]))
    ws = m.get_weights()
    assert len(ws) == 3 * (2 + 4) + 2 + 4
    ws2 = [w + 1 for w in ws]
    m.set_weights(ws2)
    assert all(np.array_equal(a, b) for a, b in zip(m.get_weights(), ws2))
    with pytest.raises(ValueError):
        m.set_weights(ws2[:-1])
    with pytest.raises(ValueError):
        m.layers[0].set_weights([np.zeros((3, 3, 2, 64), F32), np.zeros(64, F32)])
    from qnn_b200.layers.quantized_layers import QuantizedConv2D
    with pytest.raises(ValueError):
        QuantizedConv2D(filters=8, kernel_size=3, padding="same", data_format="channels_first")
    with pytest.raises(ValueError):
        QuantizedConv2D(filters=8, kernel_size=5, padding="same")
    with pytest.raises(ValueError, match="nb=16"):
        QuantizedConv2D(filters=8, kernel_size=3, padding="same").weight_mode()      # reference default nb=16 is not int8


def test_shard_range_properties():
    from qnn_b200.sharding import shard_range
    for n in (0, 1, 7, 100, 1024, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, sizes, q_out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import qnn_b200  # noqa: F401
    from qnn_b200.sharding import predict_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fake_forward(xs):          # stands in for the CUDA plan: any per-image function
        t = torch.from_numpy(xs.astype("float32")).reshape(xs.shape[0], 48)
        return torch.stack([t.sum(1), t.sum(1) * 0.5 + 1, (t * torch.arange(48)).sum(1)], dim=1)

    ok = True
    for n in sizes:
        x = np.random.default_rng(5 + n).integers(0, 256, size=(n, 4, 4, 3), dtype=np.uint8)
        out = predict_sharded(None, x, forward=fake_forward)
        ok = ok and bool(torch.equal(out, fake_forward(x))) and tuple(out.shape) == (n, 3)
    q_out.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_predict_over_gloo_world_size_2():
    """Even, ragged (11 = 6 + 5) and degenerate (1 = 1 + 0) batches: every rank ends up with the full result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    qo = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, [10, 11, 1], qo)) for r in range(2)]
    for p in procs:
        p.start()
    res = [qo.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok for _, ok in res)


def test_hdf5_reader_on_reference_checkpoint():
    """Pure-Python HDF5 reader against a real Keras checkpoint of the reference (build container only) and against
    the converted fixture that travels with the repo."""
    ref = "/root/reference/results/RESNET3/weights_44.hdf5"
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trained_resnet3_44.npz"))
    names = list(gold["names"])
    assert len(names) == 120 and names[0] == "quantized_conv2d_1/kernel"
    assert sum(gold["w%03d" % i].size for i in range(120)) == 274442
    if not os.path.exists(ref):
        pytest.skip("reference checkpoints are only present in the build container")
    from qnn_b200.hdf5_lite import read_keras_weights
    t = read_keras_weights(ref)
    assert len(t) == 41 and sum(a.size for v in t.values() for a in v.values()) == 274442
    assert t["quantized_conv2d_1"]["kernel"].shape == (3, 3, 3, 16)
    assert abs(float(t["quantized_conv2d_1"]["kernel"].max()) - 0.9385) < 1e-3
    assert t["quantized_dense_2"]["kernel"].shape == (64, 10)
    q.reset_names()
    m = q.build_model(make_cf(architecture="RESNET", nres=3), legacy_resnet=True)
    m.load_weights(ref)
    for i, w in enumerate(m.get_weights()):
        assert np.array_equal(w, gold["w%03d" % i])
    with pytest.raises(ValueError):
        q.reset_names()
        q.build_model(make_cf(architecture="RESNET", nres=5), legacy_resnet=True).load_weights(ref)


def test_checkpoint_binding_ignores_absolute_layer_counters():
    """Keras binds checkpoint layers to model layers in order, not by auto-generated name: a model built after other
    models in the same process (shifted per-class counters) must load the reference checkpoint identically."""
    ref = "/root/reference/results/RESNET3/weights_44.hdf5"
    if not os.path.exists(ref):
        pytest.skip("reference checkpoints are only present in the build container")
    q.reset_names()
    a = q.build_model(make_cf(architecture="RESNET", nres=3), legacy_resnet=True)
    a.load_weights(ref)
    # no reset_names(): every auto-generated name of the second model carries a shifted counter
    q.build_model(make_cf(architecture="RESNET", nres=1), legacy_resnet=True)
    b = q.build_model(make_cf(architecture="RESNET", nres=3), legacy_resnet=True)
    assert b.layers[0].name != a.layers[0].name
    b.load_weights(ref)
    for wa, wb in zip(a.get_weights(), b.get_weights()):
        assert np.array_equal(wa, wb)
    # by_name=True only binds identical names: nothing matches the shifted model, nothing is touched
    from qnn_b200.hdf5_lite import load_keras_weights
    c = q.build_model(make_cf(architecture="RESNET", nres=3), legacy_resnet=True)
    before = [w.copy() for w in c.get_weights()]
    load_keras_weights(c, ref, by_name=True)
    assert all(np.array_equal(x, y) for x, y in zip(before, c.get_weights()))


def test_plans_do_not_keep_models_alive_and_track_weight_versions():
    """(i) model -> plan is the only strong edge (no reference cycle: a dropped model frees its plans by reference
    counting, not whenever the cyclic collector runs); (ii) a per-layer set_weights makes the plan stale."""
    import gc
    import weakref
    from qnn_b200 import engine as E
    q.reset_names()
    m = q.build_model(make_cf(**CONFIGS["cfg3"]))
    plan = m.plan()
    assert plan.model is m
    ref_m, ref_p = weakref.ref(m), weakref.ref(plan)
    gc.disable()
    try:
        del plan, m
        assert ref_m() is None and ref_p() is None, "model/plan survive without the cyclic collector: reference cycle"
    finally:
        gc.enable()
    q.reset_names()
    m = q.build_model(make_cf(**CONFIGS["cfg3"]))
    plan = m.plan()
    plan.steps[0].dev["sentinel"] = object()           # stands for cached device constants / captured graphs
    plan._sync_weights()
    assert "sentinel" in plan.steps[0].dev             # nothing changed: caches are kept
    bn = [l for l in m.layers if isinstance(l, E.BatchNormalization)][0]
    e0 = E.weights_epoch()
    bn.set_weights([w + 1 for w in bn.get_weights()])
    assert E.weights_epoch() == e0 + 1
    plan._sync_weights()
    assert "sentinel" not in plan.steps[0].dev         # stale caches dropped
    conv = m.layers[0]
    plan.steps[0].dev["sentinel"] = object()
    conv.set_weights(conv.get_weights())
    plan._sync_weights()
    assert "sentinel" not in plan.steps[0].dev
    # model-level set_weights drops the plans altogether
    m.set_weights(m.get_weights())
    assert m.plan() is not plan
