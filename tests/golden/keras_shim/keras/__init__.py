"""A tiny stand-in for the parts of Keras 2.1 the reference imports, executed eagerly on torch-CPU fp32.

Purpose: TensorFlow/Keras are not installable in this image, so the reference cannot run as-is.  With this
package (and the sibling ``tensorflow`` stub) first on ``sys.path`` the reference's OWN sources --
``layers/*_ops.py``, ``layers/*_layers.py``, ``models/*.py`` -- import and run unmodified; their outputs on
seeded inputs are the golden fixtures in ``tests/golden/*.npz`` (see ``make_golden.py``).  Only third-party
semantics are restated here (from the Keras/TF documentation): tf.round = half-to-even, SAME padding,
BatchNormalization inference ``x*inv + (beta - mean*inv)`` with ``inv = rsqrt(var+eps)*gamma``, valid pooling,
Flatten in H,W,C order, LeakyReLU(0.3), softmax.  TEST INFRASTRUCTURE ONLY.
"""
from . import backend, layers, models, regularizers, constraints, initializers, activations  # noqa: F401

__version__ = "2.1.3-shim"
