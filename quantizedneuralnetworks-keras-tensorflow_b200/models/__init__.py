"""Model builders, named as in the reference's ``models/`` package."""
from . import model_factory, vgg, resnet  # noqa: F401
