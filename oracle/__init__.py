"""oracle/ -- CPU restatement of the reference's quantized-layer forward path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or the timed CPU baseline), never as a
fallback for the CUDA path.  The product package never imports ``oracle``.

Two oracles live here:

* ``oracle.exact``    (O1) NumPy, exact integer accumulators (float64 BLAS on exact
  integers, |sum| < 2^53) + one fixed fp32 epilogue order.  The CUDA kernels must match
  it BIT-FOR-BIT on integer accumulators, int8 activation levels, packed bits, logits
  produced from integer accumulators and arg-max labels.
* ``oracle.refstate`` (O2) torch-CPU fp32, op-for-op restatement of the reference graph
  (per-forward quantise of the kernels, un-fused bias / BatchNorm / activation /
  pooling), in two variants: O2a with the gradient-scaling identity exactly as the
  reference writes it (layers/quantized_layers.py:167-180, layers/binary_layers.py:163-176)
  and O2b without it.

Parity pin status
-----------------
The reference ships NO tests and NO golden vectors for this path, and TensorFlow /
Keras are not installable in this image, so the reference cannot be executed as-is.
The pin used instead: the reference's OWN python sources (``layers/*_ops.py``,
``layers/*_layers.py``, ``models/*.py``) are imported unmodified in the build
container on top of a small NumPy/torch stand-in for ``keras`` / ``tensorflow``
(``tests/golden/keras_shim``), and their outputs on seeded inputs are committed under
``tests/golden/*.npz`` by ``tests/golden/make_golden.py``.  ``tests/test_oracle_golden.py``
checks both oracles against those fixtures.  What stays unpinned is the arithmetic
inside TensorFlow itself (conv/dot reduction order, cuDNN algorithm choice), which no
file in the reference fixes; DESIGN.md states this.
"""
