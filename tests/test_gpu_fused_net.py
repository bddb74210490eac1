"""Whole-network launch (csrc/net_fused.cu, qnnb_vgg_forward): one kernel runs models/vgg.py:15-42 end to end for the
nets that fit one SM's shared memory.  Bit-exact against the exact oracle O1 AND against the per-layer kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import exact  # noqa: E402
from helpers import CONFIGS  # noqa: E402
from test_gpu_models import build, images  # noqa: E402

MNIST = dict(architecture='VGG', dataset='MNIST', dim=28, channels=1, nla=1, nfa=64, nlb=1, nfb=64, nlc=1, nfc=64)
CIFAR = dict(architecture='VGG', nla=1, nfa=64, nlb=1, nfb=64, nlc=1, nfc=64)

NETS = {
    "cfg1": CONFIGS["cfg1"],
    "mnist_bnn": dict(network_type='full-bnn', **MNIST),
    "mnist_w4a4": dict(network_type='full-qnn', wbits=4, abits=4, **MNIST),
    "mnist_nla2": dict(network_type='full-qnn', wbits=2, abits=2, **dict(MNIST, nla=2, nfa=32)),
    "mnist_32_64_32": dict(network_type='full-qnn', wbits=4, abits=4, **dict(MNIST, nfa=32, nfc=32)),
    "cifar_64": dict(network_type='full-qnn', wbits=4, abits=4, **CIFAR),
    "cifar_64_w8a8": dict(network_type='full-qnn', wbits=8, abits=8, **CIFAR),
    "cifar_64_bnn": dict(network_type='full-bnn', **CIFAR),
    "cifar_64_qbnn": dict(network_type='qbnn', abits=4, **CIFAR),
    "cifar_64_qtnn": dict(network_type='qtnn', abits=4, **CIFAR),
    "cifar_nlb2": dict(network_type='full-qnn', wbits=4, abits=4, **dict(CIFAR, nlb=2)),
}


@pytest.mark.parametrize("name", sorted(NETS))
@pytest.mark.parametrize("bn", ["identity", "spread"])
def test_whole_network_launch_bit_exact(name, bn):
    cf, model, nodes = build(NETS[name], bn=bn)
    plan = model.plan()
    for n in (1, 37, 333):                      # 333 > 148 SMs: persistent CTAs take several images each
        x = images(cf, n, seed=n)
        xd = torch.from_numpy(x).cuda()
        assert plan.fused_available(xd), "net %s should run as one launch" % name
        before = plan.launches
        got = plan.forward(xd).cpu().numpy()
        assert plan.launches - before == 1
        # the per-layer kernels on the same batch
        layered = plan.run(xd)[plan.output_idx].data.cpu().numpy()
        assert np.array_equal(got, layered), "fused vs per-layer kernels: max abs diff %g" % np.abs(got - layered).max()
        if n <= 37:
            want = exact.forward(nodes, x)
            assert np.array_equal(got, want), "fused vs exact oracle: max abs diff %g" % np.abs(got - want).max()


def test_whole_network_launch_scope_and_switch(monkeypatch):
    # the headline net does not fit (288 KB of kernels in its last convolution): per-layer plan
    cf, model, nodes = build(CONFIGS["cfg3"])
    x = torch.from_numpy(images(cf, 8)).cuda()
    plan = model.plan()
    assert not plan.fused_available(x)
    before = plan.launches
    plan.forward(x)
    assert plan.launches - before == len(plan.steps) == 4
    # fp32 input: the first layer is not the uint8 one
    cf1, model1, _ = build(CONFIGS["cfg1"])
    assert not model1.plan().fused_available(torch.rand(4, 28, 28, 1, device="cuda"))
    # QNNB_FUSED_NET=0 switches the whole-network launch off
    monkeypatch.setenv("QNNB_FUSED_NET", "0")
    import qnn_b200 as q
    cf2, model2, nodes2 = build(CONFIGS["cfg1"])
    x2 = images(cf2, 16)
    plan2 = model2.plan()
    assert not plan2.fused_available(torch.from_numpy(x2).cuda())
    assert np.array_equal(model2.predict(x2), exact.forward(nodes2, x2))


def test_whole_network_launch_through_predict_and_graphs():
    """model.predict on a host batch: pinned copy + CUDA-graph replay of the single launch + copy back."""
    cf, model, nodes = build(CONFIGS["cfg1"], bn="spread")
    x = images(cf, 100)
    want = exact.forward(nodes, x)
    for _ in range(3):
        assert np.array_equal(model.predict(x), want)
    hs = [model.predict_async(x) for _ in range(5)]
    for h in hs:
        assert np.array_equal(h.result().numpy(), want)
    assert np.array_equal(model.predict(x, batch_size=25), want)
