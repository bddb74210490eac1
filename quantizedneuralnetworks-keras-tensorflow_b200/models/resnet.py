"""CIFAR ResNet-v1 (depth 6*nres + 2) of the reference (models/resnet.py:15-147) on the B200
layer set: a 16-filter stem, three stacks of ``nres`` basic blocks with 16/32/64 (x pfilt)
filters, stride-2 + 1x1 projection at the start of stacks 1 and 2, block output
``Act(0.5 * (shortcut + y))``, then ``AveragePooling2D(8) -> Flatten -> Dense(softmax)``.

``legacy=True`` reproduces the older revision the shipped ``results/RESNET3/weights_*.hdf5`` were
trained with: biased convs/dense and no ``* 0.5`` (SURVEY.md finding 7).
"""
from ..engine import (Model, Input, BatchNormalization, AveragePooling2D, Flatten, ZeroPadding2D, Lambda, add)


def ResNet18(Conv2D, Activation, Dense, cf, legacy=False):
    n_blocks = int(cf.nres)
    depth = 6 * n_blocks + 2
    widen = int(getattr(cf, "pfilt", 1))
    init = getattr(cf, "kernel_initializer", "he_normal")
    reg = getattr(cf, "kernel_regularizer", 0.0)
    biased = bool(legacy)

    def stage(t, filters, kernel_size=3, strides=1, norm=True, act=True):
        """conv -> [BatchNormalization] -> [Act]   (resnet_layer with conv_first=True, resnet.py:26-70)"""
        t = Conv2D(filters=filters * widen, kernel_size=kernel_size, strides=strides, padding='same',
                   kernel_initializer=init, kernel_regularizer=("l2", reg), use_bias=biased)(t)
        if norm:
            t = BatchNormalization()(t)
        if act:
            t = Activation()(t)
        return t

    image = Input(shape=(cf.dim, cf.dim, cf.channels))
    t = image
    if cf.dataset in ("MNIST", "FASHION"):
        t = ZeroPadding2D(padding=(2, 2))(t)            # 28 -> 32 (resnet.py:101-102)
    t = stage(t, 16)
    filters = 16
    for stack in range(3):
        for block in range(n_blocks):
            downsample = stack > 0 and block == 0
            s = 2 if downsample else 1
            y = stage(t, filters, strides=s)
            y = stage(y, filters, act=False)
            if downsample:
                # linear 1x1 projection so the shortcut matches the new shape (resnet.py:117-126)
                t = stage(t, filters, kernel_size=1, strides=s, norm=False, act=False)
            t = add([t, y])
            if not legacy:
                t = Lambda(lambda v: v * 0.5)(t)        # resnet.py:128
            t = Activation()(t)
        filters *= 2
    t = AveragePooling2D(pool_size=8)(t)
    t = Flatten()(t)
    probs = Dense(units=cf.classes, activation='softmax', kernel_initializer=init, use_bias=biased)(t)
    model = Model(inputs=image, outputs=probs)
    model.depth = depth
    return model
