import torch.nn.functional as F

from . import Layer


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kwargs):
        super(LeakyReLU, self).__init__(**kwargs)
        self.alpha = alpha

    def call(self, x):
        return F.leaky_relu(x, negative_slope=self.alpha)
