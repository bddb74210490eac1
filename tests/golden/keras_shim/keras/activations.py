from . import backend as K


def get(a):
    if a is None or a == "linear":
        return None
    if a == "softmax":
        return K.softmax
    if callable(a):
        return a
    raise ValueError(a)
