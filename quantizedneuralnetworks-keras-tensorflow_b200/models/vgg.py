"""VGG-style network of the reference (models/vgg.py:5-44) built on the B200 layer set.

Three blocks of ``[conv3x3-same, BatchNormalization(eps=1e-4), Act]`` repeated nla / nlb / nlc
times, each block closed by a 2x2 max-pool, then ``Flatten -> Fc(classes) -> BatchNormalization``.
Logits are returned without softmax (the reference trains them with a squared hinge loss).
"""
from ..engine import Sequential, MaxPooling2D, BatchNormalization, Flatten

BN_EPS = 1e-4       # models/vgg.py:16
BN_MOMENTUM = 0.1


def Vgg(Conv, Act, Fc, cf):
    """``Conv``, ``Act``, ``Fc`` are the factories chosen by ``model_factory.build_model``."""
    reg = getattr(cf, "kernel_regularizer", 0.0)
    init = getattr(cf, "kernel_initializer", "glorot_uniform")

    def conv3(filters, **extra):
        return Conv(kernel_size=(3, 3), filters=filters, strides=(1, 1), padding='same',
                    kernel_initializer=init, kernel_regularizer=("l2", reg), **extra)

    def unit(net, filters, **extra):
        net.add(conv3(filters, **extra))
        net.add(BatchNormalization(momentum=BN_MOMENTUM, epsilon=BN_EPS))
        net.add(Act())

    net = Sequential()
    # block A: the first conv carries the input shape and always exists (vgg.py:15), then nla-1 more
    unit(net, cf.nfa, input_shape=(cf.dim, cf.dim, cf.channels))
    blocks = [(cf.nla - 1, cf.nfa), (cf.nlb, cf.nfb), (cf.nlc, cf.nfc)]
    for depth, width in blocks:
        for _ in range(depth):
            unit(net, width)
        net.add(MaxPooling2D(pool_size=(2, 2)))
    net.add(Flatten())
    net.add(Fc(cf.classes))
    net.add(BatchNormalization(momentum=BN_MOMENTUM, epsilon=BN_EPS))
    return net
