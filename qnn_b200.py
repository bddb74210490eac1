"""Import alias for the hyphen-named package ``quantizedneuralnetworks-keras-tensorflow_b200``.

``import qnn_b200`` returns that package; its sub-modules are also reachable as
``qnn_b200.layers.quantized_layers`` etc. (registered in ``sys.modules`` under both names).
"""
import importlib
import os
import sys

_REAL = "quantizedneuralnetworks-keras-tensorflow_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules["qnn_b200" + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
