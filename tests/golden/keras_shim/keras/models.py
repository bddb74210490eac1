import numpy as np
import torch

from .layers import Layer, Tracer, InputLayer


class _Base(object):
    def summary(self, *a, **k):
        pass

    def get_weights(self):
        out = []
        for l in self.layers:
            out.extend(l.get_weights())
        return out

    def set_weights(self, ws):
        i = 0
        for l in self.layers:
            n = len(l.weights)
            l.set_weights(ws[i:i + n])
            i += n
        assert i == len(ws), (i, len(ws))

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(name)


class Sequential(_Base):
    def __init__(self):
        self.layers = []
        self._shape = None

    def add(self, layer):
        if self._shape is None:
            shp = layer._kw.get("input_shape")
            assert shp is not None, "first layer needs input_shape"
            self._probe = torch.zeros((1,) + tuple(shp))
        self._probe = layer(self._probe)           # builds the layer eagerly on a dummy sample
        self._shape = tuple(self._probe.shape)
        self.layers.append(layer)

    def predict(self, x, batch_size=None, verbose=0, taps=None):
        t = torch.as_tensor(np.asarray(x, dtype=np.float32))
        with torch.no_grad():
            for l in self.layers:
                t = l(t)
                if taps is not None:
                    taps.append((l.name, t.numpy().copy()))
        return t.numpy()


class Model(_Base):
    def __init__(self, inputs, outputs):
        self.input, self.output = inputs, outputs
        order, seen = [], set()

        def visit(tr):
            if id(tr) in seen:
                return
            seen.add(id(tr))
            for p in tr.parents:
                visit(p)
            order.append(tr)

        visit(outputs)
        self._order = order
        self.layers = []
        for tr in order:
            if tr.layer is not None and not isinstance(tr.layer, InputLayer) and tr.layer not in self.layers:
                self.layers.append(tr.layer)

    def predict(self, x, batch_size=None, verbose=0, taps=None):
        vals = {}
        t = torch.as_tensor(np.asarray(x, dtype=np.float32))
        with torch.no_grad():
            for tr in self._order:
                if not tr.parents:
                    vals[id(tr)] = t
                else:
                    ins = [vals[id(p)] for p in tr.parents]
                    vals[id(tr)] = tr.layer.call(ins if tr.multi else ins[0])
                    if taps is not None:
                        taps.append((tr.layer.name, vals[id(tr)].numpy().copy()))
        return vals[id(self._order[-1])].numpy()
