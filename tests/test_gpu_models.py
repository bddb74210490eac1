"""End-to-end parity of the model_factory networks on the GPU against both oracles.

* bit-exact vs O1 (oracle.exact): logits, arg-max labels and every intermediate activation level;
* vs O2b (reference restatement, no scaling identity): max abs logit error <= 1e-4 relative and
  top-1 agreement >= 99.9 % for w,a <= 4 (the north_star tolerance);
* vs O2a (with the scaling identity as the reference writes it): teacher-forced layer-wise
  agreement -- the end-to-end gap is the reference's own fp32 noise floor (SURVEY.md finding 5).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import exact, netspec, refstate  # noqa: E402
from helpers import make_cf, CONFIGS, assign_weights_from_spec  # noqa: E402

F32 = np.float32


def build(cfkw, bn="spread", bias_range=0.1, seed=42, legacy=False):
    import qnn_b200 as q
    q.reset_names()
    cf = make_cf(**cfkw)
    model = q.build_model(cf, legacy_resnet=legacy)
    nodes = netspec.build_spec(cf, **({"use_bias": True, "half": False} if legacy else {}))
    weights = netspec.random_weights(nodes, seed=seed, bias_range=bias_range, bn=bn)
    # the product and the oracle enumerate layers in different (both valid) orders for ResNet:
    # match weights through the per-layer structure instead of position
    netspec.set_weights(nodes, weights)
    assign_weights_from_spec(model, nodes)
    return cf, model, nodes


def images(cf, n, seed=1234):
    return np.random.default_rng(seed).integers(0, 256, size=(n, cf.dim, cf.dim, cf.channels), dtype=np.uint8)


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name,n", [("cfg1", 100), ("cfg3", 48), ("cfg2", 32), ("cfg4", 8)])
@pytest.mark.parametrize("bn", ["identity", "spread"])
def test_vgg_bit_exact_vs_exact_oracle(name, n, bn):
    cf, model, nodes = build(CONFIGS[name], bn=bn)
    x = images(cf, n)
    want, vals, info = exact.forward(nodes, x, return_all=True)
    got = model.predict(x)
    assert got.shape == want.shape == (n, 10)
    assert np.array_equal(got, want), "logit max abs diff %g" % np.abs(got - want).max()
    assert np.array_equal(got.argmax(1), want.argmax(1))
    # torch CUDA tensor in -> torch CUDA tensor out, same numbers
    got_t = model.predict(torch.from_numpy(x).cuda())
    assert got_t.is_cuda and np.array_equal(got_t.cpu().numpy(), want)
    # chunked execution gives identical results (images are independent)
    assert np.array_equal(model.predict(x, batch_size=7), want)


def test_vgg_intermediate_levels_bit_exact():
    """Every fused step's output tensor equals the oracle's value at the same graph position."""
    from qnn_b200 import kernels as K
    from helpers import unpack_bits
    for name in ("cfg3", "cfg2"):
        cf, model, nodes = build(CONFIGS[name], bn="spread")
        x = images(cf, 16)
        _, vals, _ = exact.forward(nodes, x, return_all=True)
        plan = model.plan()
        env = plan.run(torch.from_numpy(x).cuda())
        # pooled activations after each block: spec nodes with op maxpool, in order
        pools = [v for nd, v in zip(nodes, vals) if nd["op"] == "maxpool"]
        steps = [s for s in plan.steps if s.kind == "conv"]
        assert len(pools) == len(steps) == 3
        for st, want in zip(steps, pools):
            qt = env[st.out]
            got = qt.data.cpu().numpy()
            if qt.kind == "b1":
                got = unpack_bits(got, qt.channels)
            assert np.array_equal(got.astype(np.int32), want.data.astype(np.int32))


def test_vgg_vs_reference_restatement_tolerance():
    """north_star tolerance vs the fp32 restatement of the reference (O2b, w,a <= 4): max abs logit
    error <= 1e-4 relative, top-1 agreement >= 99.9 %; plus the O2a <-> O2b gap for context."""
    cf, model, nodes = build(CONFIGS["cfg3"], bn="spread")
    x = images(cf, 128)
    got = model.predict(x)
    o2b = refstate.forward(nodes, x, trick=False)
    assert rel_err(got, o2b) <= 1e-4
    assert (got.argmax(1) == o2b.argmax(1)).mean() >= 0.999
    cf1, model1, nodes1 = build(CONFIGS["cfg1"], bn="spread")
    x1 = images(cf1, 100)
    got1 = model1.predict(x1)
    o2b1 = refstate.forward(nodes1, x1, trick=False)
    assert rel_err(got1, o2b1) <= 1e-4
    assert (got1.argmax(1) == o2b1.argmax(1)).mean() >= 0.999


def test_vgg_teacher_forced_vs_reference_with_scaling_identity():
    """Layer-wise agreement with the reference AS WRITTEN (O2a): each layer of the restatement is fed the
    GPU's own activations; the fraction of activation levels that differ must stay <= 2e-4 (each by
    one level) and the dense head must agree to 1e-5 relative."""
    from helpers import unpack_bits
    cf, model, nodes = build(CONFIGS["cfg3"], bn="spread")
    x = images(cf, 64)
    plan = model.plan()
    env = plan.run(torch.from_numpy(x).cuda())
    steps = [s for s in plan.steps if s.kind == "conv"]
    pool_nodes = [i for i, nd in enumerate(nodes) if nd["op"] == "maxpool"]
    teacher = {}
    for st, ni in zip(steps, pool_nodes):
        teacher[ni] = env[st.out].to_float().cpu().numpy()
    out, vals, info = refstate.forward(nodes, x, trick=True, return_all=True, teacher=teacher)
    for ni in pool_nodes:
        own = info["own"][ni]
        diff = np.abs(own - teacher[ni])
        assert (diff > 0).mean() <= 2e-4
        assert diff.max() <= 1.0 / 8 + 1e-7
    got = env[plan.output_idx].data.cpu().numpy()
    assert rel_err(got, out) <= 1e-5


def test_forward_into_caller_buffer_and_peer_block():
    """Plan.forward(out=...) -- the hook of the NVLink logit path: the final dense kernel writes the network output
    into the caller's buffer (a tensor, or a raw device address as sharding.PeerGather hands out)."""
    from qnn_b200 import _lib as L
    cf, model, nodes = build(CONFIGS["cfg3"])
    x = images(cf, 16)
    want = exact.forward(nodes, x)
    plan = model.plan()
    xd = torch.from_numpy(x).cuda()
    ring = torch.full((3 * 16, 10), -7.0, dtype=torch.float32, device="cuda")
    ret = plan.forward(xd, out=ring[16:32])
    torch.cuda.synchronize()
    assert ret.data_ptr() == ring[16:32].data_ptr()
    assert np.array_equal(ring[16:32].cpu().numpy(), want)
    assert float(ring[:16].max()) == -7.0 and float(ring[32:].max()) == -7.0
    raw = L.DeviceBuffer(ring.data_ptr(), ring.shape, torch.float32)
    plan.forward(xd, out=raw.rows(32, 48))
    torch.cuda.synchronize()
    assert np.array_equal(ring[32:].cpu().numpy(), want)
    with pytest.raises(ValueError):
        plan.forward(xd, out=ring[:8])


def test_peer_gather_root_side(tmp_path):
    """sharding.PeerGather on its owning rank (world of one): exported buffer, per-slot blocks, torch view."""
    import torch.distributed as dist
    from qnn_b200.sharding import PeerGather
    cf, model, nodes = build(CONFIGS["cfg3"])
    x = images(cf, 8)
    want = exact.forward(nodes, x)
    started = not dist.is_initialized()
    if started:
        dist.init_process_group("gloo", init_method="file://%s" % (tmp_path / "rdzv"), world_size=1, rank=0)
    try:
        pg = PeerGather(8, 10, slots=3)
        plan = model.plan()
        xd = torch.from_numpy(x).cuda()
        plan.forward(xd, out=pg.block(2))
        pg.fence()
        assert np.array_equal(pg.gathered(2).cpu().numpy(), want)
        assert float(pg.gathered(0).abs().max()) == 0.0 and float(pg.gathered(1).abs().max()) == 0.0
        pg.close()
    finally:
        if started:
            dist.destroy_process_group()


@pytest.mark.parametrize("nt,legacy", [("full-qnn", False), ("full-qnn", True), ("full-bnn", False), ("qbnn", False), ("qtnn", False)])
def test_resnet_integer_types_bit_exact(nt, legacy):
    cf, model, nodes = build(dict(network_type=nt, wbits=4, abits=4, architecture='RESNET', nres=2), bn="spread",
                             bias_range=0.1 if legacy else 0.0, legacy=legacy)
    x = images(cf, 24)
    want, vals, info = exact.forward(nodes, x, return_all=True)
    got, logits = model.predict(x, return_logits=True)
    assert np.array_equal(logits, info["logits"]), "logit max abs diff %g" % np.abs(logits - info["logits"]).max()
    assert np.abs(got - want).max() <= 2e-6          # softmax: expf vs float64 exp
    assert np.array_equal(got.argmax(1), want.argmax(1))


@pytest.mark.parametrize("arch,nt", [("VGG", "qnn"), ("VGG", "bnn"), ("VGG", "tnn"), ("RESNET", "qnn"), ("RESNET", "tnn"), ("RESNET", "bnn")])
def test_float_activation_types_tolerance(arch, nt):
    """LeakyReLU nets keep fp32 activations: accumulation is floating point, so the bar is the north_star
    tolerance (<= 1e-4 relative on logits, top-1 >= 99.9 %) against O1 (float64 accumulation) and O2b."""
    cf, model, nodes = build(dict(network_type=nt, wbits=4, abits=4, architecture=arch, nres=2), bn="spread")
    x = images(cf, 32)
    want, vals, info = exact.forward(nodes, x, return_all=True)
    got, logits = model.predict(x, return_logits=True)
    assert rel_err(logits, info["logits"]) <= 1e-4
    assert (got.argmax(1) == want.argmax(1)).mean() >= 0.999
    o2b = refstate.forward(nodes, x, trick=False)
    assert rel_err(got, o2b) <= 1e-4


@pytest.mark.parametrize("nt", ["qnn", "tnn"])
def test_cfg5_resnet62_small_batch_tolerance(nt):
    """BASELINE.json configs[4] at its real depth (nres = 10, 63 convolutions) on 8 images: fp32 LeakyReLU activations,
    so the bar is the north_star tolerance against O1 (float64 accumulation) and O2b; arg-max identical."""
    cf, model, nodes = build(dict(CONFIGS["cfg5"], network_type=nt), bn="spread")
    assert model.depth == 62
    x = images(cf, 8)
    want, _, info = exact.forward(nodes, x, return_all=True)
    got, logits = model.predict(x, return_logits=True)
    assert rel_err(logits, info["logits"]) <= 1e-4
    assert np.array_equal(got.argmax(1), want.argmax(1))
    o2b = refstate.forward(nodes, x, trick=False)
    assert rel_err(got, o2b) <= 1e-4


def test_cfg5_integer_twin_full_qnn_resnet62_bit_exact():
    """Same depth with quantised activations (full-qnn): integer path, bit-exact logits."""
    cf, model, nodes = build(dict(CONFIGS["cfg5"], network_type="full-qnn"), bn="spread")
    x = images(cf, 8)
    want, _, info = exact.forward(nodes, x, return_all=True)
    got, logits = model.predict(x, return_logits=True)
    assert np.array_equal(logits, info["logits"])
    assert np.array_equal(got.argmax(1), want.argmax(1))


def test_mnist_resnet_zero_padding_path():
    cf, model, nodes = build(dict(network_type='full-qnn', wbits=4, abits=4, architecture='RESNET', nres=1,
                                  dataset='MNIST', dim=28, channels=1), bn="spread")
    x = images(cf, 8)
    want, _, info = exact.forward(nodes, x, return_all=True)
    got, logits = model.predict(x, return_logits=True)
    assert np.array_equal(logits, info["logits"])


def test_float_image_input_matches_uint8_input_within_tolerance():
    """The reference is fed uint8/255 as fp32 (utils/load_data.py:40); the same array given to predict() runs
    the first layer on the fp32 path and must agree with the integer first layer within tolerance."""
    cf, model, nodes = build(CONFIGS["cfg3"], bn="spread")
    x = images(cf, 32)
    a = model.predict(x)
    b = model.predict(x.astype("float32") / 255)
    assert (a.argmax(1) == b.argmax(1)).mean() >= 0.96      # first-layer 1-LSB flips can avalanche (SURVEY finding 5)
    o2b = refstate.forward(nodes, x, trick=False)
    assert rel_err(a, o2b) <= 1e-4


def test_layer_objects_standalone_call():
    """Un-fused use of the layer objects (build / call / get_weights / get_config), as a Keras user would."""
    import qnn_b200 as q
    from qnn_b200.layers.quantized_layers import QuantizedConv2D, QuantizedDense
    from qnn_b200.layers.binary_layers import BinaryConv2D
    from qnn_b200.layers.ternary_layers import TernaryConv2D
    from helpers import oracle_layer
    rng = np.random.default_rng(0)
    x = rng.normal(0, 1, size=(2, 8, 8, 16)).astype(F32)
    for cls, wkind, nb, kw in ((QuantizedConv2D, "quantized", 4, dict(nb=4)), (BinaryConv2D, "binary", 1, {}), (TernaryConv2D, "ternary", 2, {})):
        lay = cls(filters=24, kernel_size=(3, 3), padding='same', H=1., **kw)
        y = lay(torch.from_numpy(x).cuda())
        assert lay.built and lay.kernel.shape == (3, 3, 16, 24) and len(lay.get_weights()) == 2
        b = rng.uniform(-1, 1, 24).astype(F32)
        lay.set_weights([lay.kernel, b])
        y = lay(torch.from_numpy(x).cuda()).cpu().numpy()
        want, _ = oracle_layer(x, "f32", 1.0, lay.kernel, wkind, nb, 1.0, 1, bias=b)
        assert np.abs(y - want).max() <= 1e-4 * np.abs(want).max()
        cfg = lay.get_config()
        assert cfg["H"] == 1.0 and abs(cfg["kernel_lr_multiplier"] - np.sqrt((16 * 9 + 24 * 9) / 1.5)) < 1e-3
    d = QuantizedDense(10, nb=4)
    z = d(torch.from_numpy(x.reshape(2, -1)).cuda()).cpu().numpy()
    want, _ = oracle_layer(x.reshape(2, -1), "f32", 1.0, d.kernel, "quantized", 4, 1.0, 1, bias=d.bias, dense=True)
    assert np.abs(z - want).max() <= 1e-4 * np.abs(want).max()
    with pytest.raises(ValueError):
        QuantizedConv2D(filters=4, kernel_size=3, padding='same')(q.Input(shape=(8, 8, None)))


# --------------------------------------------------------------------------- host path: graphs, staleness, pipelining
def test_layer_set_weights_after_predict_is_not_stale():
    """The reference idiom ``model.layers[i].set_weights(...)`` (models/model_factory.py:77-105) AFTER a first predict:
    the captured CUDA graphs and cached BN constants of the plan must be rebuilt, not replayed."""
    import qnn_b200 as q
    from qnn_b200 import engine as E
    cf, model, nodes = build(CONFIGS["cfg3"], bn="spread")
    x = images(cf, 16)
    assert np.array_equal(model.predict(x), exact.forward(nodes, x))
    rng = np.random.default_rng(5)
    # new kernel for the second conv, new statistics for the first BN -- through the LAYERS, no model-level call
    conv = [l for l in model.layers if "conv2d" in l.name][1]
    bn = [l for l in model.layers if isinstance(l, E.BatchNormalization)][0]
    k2 = rng.uniform(-1, 1, size=conv.kernel.shape).astype(F32)
    conv.set_weights([k2, conv.bias])
    g, b, m, v = bn.get_weights()
    bn.set_weights([g * F32(0.9), b + F32(0.05), m, v * F32(1.1)])
    from helpers import spec_weights_from_model
    spec_weights_from_model(model, nodes)
    want = exact.forward(nodes, x)
    got = model.predict(x)                       # host path: CUDA-graph slots
    assert np.array_equal(got, want)
    got_dev = model.predict(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got_dev, want)


def test_predict_async_handles_survive_slot_recycling():
    """More outstanding handles than pipeline slots, read late and out of order: every handle returns ITS batch."""
    cf, model, nodes = build(CONFIGS["cfg3"], bn="spread")
    batches = [images(cf, 8, seed=100 + i) for i in range(8)]
    handles = [model.predict_async(b) for b in batches]          # PIPELINE_DEPTH slots, 8 handles
    for i in (7, 0, 3, 5, 1, 2, 6, 4):
        assert np.array_equal(handles[i].result().numpy(), exact.forward(nodes, batches[i])), "handle %d" % i


def test_many_models_one_process_graph_capture():
    """A long-lived process that builds model after model (each with its own CUDA-graph slots) must keep capturing:
    dead plans release their graphs deterministically, and no finaliser runs inside a capture."""
    import gc
    want_cache = {}
    for rep in range(12):
        name = ("cfg3", "cfg2", "cfg1")[rep % 3]
        cf, model, nodes = build(CONFIGS[name], bn="spread", seed=rep)
        x = images(cf, 4 + rep, seed=rep)
        got = model.predict(x)
        assert np.array_equal(got, exact.forward(nodes, x))
        if rep % 4 == 3:
            gc.collect()


# --------------------------------------------------------------------------- trained weights (reference checkpoints)
def _trained(code, nt):
    import os
    import qnn_b200 as q
    from helpers import spec_weights_from_model
    q.reset_names()
    cf = make_cf(network_type=nt, wbits=4, abits=4, architecture='RESNET', nres=3, kernel_initializer='he_normal', kernel_regularizer=1e-4)
    model = q.build_model(cf, legacy_resnet=True)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trained_resnet3_%s.npz" % code))
    model.set_weights([g["w%03d" % i] for i in range(len(g.files) - 1)])
    nodes = netspec.build_spec(cf, use_bias=True, half=False)
    spec_weights_from_model(model, nodes)
    return cf, model, nodes


@pytest.mark.parametrize("code,nt", [("44", "full-qnn"), ("bb", "full-bnn")])
def test_trained_reference_checkpoints_bit_exact(code, nt):
    """The reference's own trained ResNet-20 checkpoints (results/RESNET3/weights_{44,bb}.hdf5, converted by
    tests/golden/make_trained_fixture.py): real weight / BatchNorm statistics, older biased graph revision.
    Logits bit-exact vs O1 on noise and on smooth images; tolerance vs the fp32 restatement."""
    cf, model, nodes = _trained(code, nt)
    rng = np.random.default_rng(77)
    noise = rng.integers(0, 256, size=(32, 32, 32, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:32, 0:32]
    smooth = np.stack([np.stack([(127 + 120 * np.sin(xx / (3 + i) + c) * np.cos(yy / (4 + i))) for c in range(3)], -1) for i in range(16)])
    for x in (noise, smooth.astype(np.uint8)):
        want, vals, info = exact.forward(nodes, x, return_all=True)
        got, logits = model.predict(x, return_logits=True)
        assert np.array_equal(logits, info["logits"]), "max abs logit diff %g" % np.abs(logits - info["logits"]).max()
        assert np.array_equal(got.argmax(1), want.argmax(1))
        o2b = refstate.forward(nodes, x, trick=False)
        assert rel_err(got, o2b) <= 1e-4 and (got.argmax(1) == o2b.argmax(1)).mean() >= 0.999


@pytest.mark.parametrize("arch", ["VGG", "RESNET"])
def test_float_network_type_tolerance(arch):
    """network_type 'float' (keras Conv2D / Dense / LeakyReLU, model_factory.py:24-27): fp32 kernels x fp32 activations on
    the FFMA kernels.  Floating-point accumulation, so the bar is the north_star tolerance against O1 (float64
    accumulation) and against the fp32 restatement of the reference graph."""
    cf, model, nodes = build(dict(network_type='float', architecture=arch, nres=2), bn="spread")
    x = images(cf, 32)
    want, vals, info = exact.forward(nodes, x, return_all=True)
    if arch == "RESNET":
        got, logits = model.predict(x, return_logits=True)
        assert rel_err(logits, info["logits"]) <= 1e-4
    else:
        got = model.predict(x)
    assert rel_err(got, want) <= 1e-4
    assert (got.argmax(1) == want.argmax(1)).mean() >= 0.999
    ref = refstate.forward(nodes, x, trick=True)             # plain layers: no scaling identity either way
    assert rel_err(got, ref) <= 1e-4
    # same numbers for a float32 image batch (what the reference is fed: uint8 / 255, utils/load_data.py:40)
    got_f = model.predict((x.astype(np.float32) / 255))
    assert rel_err(got_f, want) <= 1e-4
    assert not model.plan().fused_available(torch.from_numpy(x).cuda())


def test_trained_float_checkpoint_tolerance():
    """The reference's trained float ResNet-20 (results/RESNET3/weights_ff.hdf5, test accuracy 0.8114 in results.out:29)
    through the fused plan vs both oracles on noise and smooth images."""
    cf, model, nodes = _trained("ff", "float")
    rng = np.random.default_rng(78)
    noise = rng.integers(0, 256, size=(32, 32, 32, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:32, 0:32]
    smooth = np.stack([np.stack([(127 + 120 * np.sin(xx / (3 + i) + c) * np.cos(yy / (4 + i))) for c in range(3)], -1) for i in range(16)])
    for x in (noise, smooth.astype(np.uint8)):
        want, vals, info = exact.forward(nodes, x, return_all=True)
        got, logits = model.predict(x, return_logits=True)
        assert rel_err(logits, info["logits"]) <= 1e-4
        assert (got.argmax(1) == want.argmax(1)).mean() >= 0.999
        ref = refstate.forward(nodes, x, trick=True)
        assert rel_err(got, ref) <= 1e-4


def test_evaluate_matches_host_metrics():
    cf, model, nodes = _trained("44", "full-qnn")
    x = images(cf, 64)
    labels = np.random.default_rng(3).integers(0, 10, size=64)
    y = np.eye(10, dtype=np.float32)[labels]
    loss, acc = model.evaluate(x, y)
    p = exact.forward(nodes, x)
    assert abs(acc - float((p.argmax(1) == labels).mean())) < 1e-9
    want_loss = float(np.mean(-np.log(np.clip(p[np.arange(64), labels], 1e-7, 1))))
    assert abs(loss - want_loss) <= 1e-4 * max(want_loss, 1.0)


# --------------------------------------------------------------------------- BASELINE.json full batch sizes: size-independent properties
@pytest.mark.parametrize("name,n,chunk", [("cfg3", 1024, 256), ("cfg2", 256, 64), ("cfg4", 4096, 1024), ("cfg5", 1024, 256)])
def test_full_size_batches_properties(name, n, chunk):
    """At the batch sizes BASELINE.json quotes, where the oracle cannot run the whole batch in seconds:
    (a) images are independent -- one forward over the full batch equals the same batch run in chunks, bit for bit;
    (b) permuting the batch permutes the logits; (c) a sample of images spread over the batch equals the exact oracle
    (bit-exact for the integer nets, north_star tolerance for the fp32-activation net)."""
    cf, model, nodes = build(CONFIGS[name], bn="spread")
    x = images(cf, n, seed=4321)
    full = model.predict(x)
    assert full.shape == (n, 10) and np.isfinite(full).all()
    assert np.array_equal(model.predict(x, batch_size=chunk), full)
    perm = np.random.default_rng(1).permutation(n)
    assert np.array_equal(model.predict(x[perm]), full[perm])
    idx = np.linspace(0, n - 1, 8).astype(int)
    want = exact.forward(nodes, x[idx])
    if cf.network_type in ("full-qnn", "full-bnn"):
        assert np.array_equal(full[idx], want)
    else:
        assert rel_err(full[idx], want) <= 1e-4
        assert np.array_equal(full[idx].argmax(1), want.argmax(1))


@pytest.mark.parametrize("name,n", [("cfg3", 300), ("cfg2", 64), ("cfg1", 100), ("cfg5", 8)])
def test_sm_share_gives_identical_results(name, n):
    """Plan(sm_share=k) sizes every persistent grid for k SMs (qnnb_*_desc.max_ctas) so that kernels of independent
    batches can run side by side; the tile -> CTA assignment changes, the arithmetic must not: bit-identical outputs
    for every share, also when batches on different streams really do overlap."""
    cf, model, nodes = build(CONFIGS[name], bn="spread")
    x = torch.from_numpy(images(cf, n)).cuda()
    want = model.plan().forward(x).clone()
    for share in (1, 37, 74, 1000):
        got = model.plan(sm_share=share).forward(x)
        assert torch.equal(got, want), "sm_share=%d changes the result" % share
    # four batches in flight on four streams, 37 SMs per kernel
    plan = model.plan(sm_share=37)
    xs = [torch.from_numpy(images(cf, n, seed=100 + i)).cuda() for i in range(4)]
    refs = [model.plan().forward(xi).clone() for xi in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]
    outs = []
    for rep in range(3):
        for xi, st in zip(xs, streams):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append(plan.forward(xi))
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        assert torch.equal(o, refs[i % 4])
