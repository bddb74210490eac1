"""Minimal pure-Python HDF5 reader for Keras-2.1 ``model.save`` / ``save_weights`` checkpoints
(the format of the reference's ``results/RESNET3/weights_*.hdf5``; h5py is not available here).

Covers exactly what those files use: superblock version 0, version-1 object headers (with continuation
blocks), old-style groups (symbol table message -> v1 B-tree + local heap + SNOD nodes), simple dataspaces,
fixed-point / IEEE little-endian datatypes and contiguous or compact dataset layouts.  Anything else raises
``ValueError`` -- there is no partial/approximate read.

``load_keras_weights(model, path)`` copies the tensors under ``/model_weights`` (or the file root for
``save_weights`` files) into a model built by ``models.model_factory.build_model`` by layer name, in Keras
order (conv/dense: kernel, bias; BatchNormalization: gamma, beta, moving_mean, moving_variance).
"""
from __future__ import annotations

import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.d = f.read()
        if self.d[:8] != _SIG:
            raise ValueError("%s: not an HDF5 file" % path)
        ver = self.d[8]
        if ver != 0:
            raise ValueError("unsupported HDF5 superblock version %d" % ver)
        self.so, self.sl = self.d[13], self.d[14]
        if (self.so, self.sl) != (8, 8):
            raise ValueError("unsupported offset/length sizes %d/%d" % (self.so, self.sl))
        # 8 sig + 8 versions/sizes + 2+2 K values + 4 flags = 24; then base, freespace, eof, driver addresses
        self.base = self._u64(24)
        root_entry = 24 + 4 * 8
        self.root_header = self._u64(root_entry + 8)

    # ---- primitives
    def _u16(self, o):
        return struct.unpack_from("<H", self.d, o)[0]

    def _u32(self, o):
        return struct.unpack_from("<I", self.d, o)[0]

    def _u64(self, o):
        return struct.unpack_from("<Q", self.d, o)[0]

    # ---- object headers
    def messages(self, addr):
        """Yield (type, flags, payload offset, payload size) of a version-1 object header, following continuations."""
        addr += self.base
        if self.d[addr] != 1:
            raise ValueError("unsupported object header version %d" % self.d[addr])
        nmsg = self._u16(addr + 2)
        hsize = self._u32(addr + 8)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u16(off), self._u16(off + 2), self.d[off + 4]
                body = off + 8
                if mtype == 0x10:                         # continuation
                    blocks.append((self._u64(body) + self.base, self._u64(body + 8)))
                out.append((mtype, flags, body, msize))
                off = body + msize
        return out

    # ---- groups
    def _heap_string(self, heap_addr, offset):
        h = heap_addr + self.base
        if self.d[h:h + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        data = self._u64(h + 24) + self.base
        s = data + offset
        e = self.d.index(b"\x00", s)
        return self.d[s:e].decode("utf-8")

    def _btree_entries(self, addr, heap):
        a = addr + self.base
        sig = self.d[a:a + 4]
        if sig == b"TREE":
            level, used = self.d[a + 5], self._u16(a + 6)
            off = a + 8 + 16
            children = []
            for i in range(used):
                off += 8                                  # key
                children.append(self._u64(off))
                off += 8
            for c in children:
                if level > 0:
                    yield from self._btree_entries(c, heap)
                else:
                    yield from self._snod_entries(c, heap)
        else:
            raise ValueError("bad B-tree signature %r" % sig)

    def _snod_entries(self, addr, heap):
        a = addr + self.base
        if self.d[a:a + 4] != b"SNOD":
            raise ValueError("bad symbol node signature")
        n = self._u16(a + 6)
        off = a + 8
        for _ in range(n):
            name = self._heap_string(heap, self._u64(off))
            yield name, self._u64(off + 8)
            off += 40

    def children(self, header_addr):
        """{name: object header address} for a group, {} for a dataset."""
        for mtype, _, body, _ in self.messages(header_addr):
            if mtype == 0x11:
                btree, heap = self._u64(body), self._u64(body + 8)
                return dict(self._btree_entries(btree, heap))
        return {}

    # ---- datasets
    def dataset(self, header_addr):
        shape = dtype = None
        data = None
        for mtype, _, body, size in self.messages(header_addr):
            if mtype == 0x01:
                ver, rank = self.d[body], self.d[body + 1]
                off = body + (8 if ver == 1 else 4)
                shape = tuple(self._u64(off + 8 * i) for i in range(rank))
            elif mtype == 0x03:
                cls = self.d[body] & 0x0F
                bits0 = self.d[body + 1]
                nbytes = self._u32(body + 4)
                if bits0 & 1:
                    raise ValueError("big-endian datasets are not supported")
                if cls == 1:
                    dtype = {2: np.float16, 4: np.float32, 8: np.float64}[nbytes]
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = np.dtype("%s%d" % ("i" if signed else "u", nbytes))
                else:
                    raise ValueError("unsupported datatype class %d" % cls)
            elif mtype == 0x08:
                ver = self.d[body]
                if ver != 3:
                    raise ValueError("unsupported data layout version %d" % ver)
                lclass = self.d[body + 1]
                if lclass == 1:
                    addr, nbytes = self._u64(body + 2), self._u64(body + 10)
                    data = (None, 0) if addr == UNDEF else (addr + self.base, nbytes)
                elif lclass == 0:
                    nbytes = self._u16(body + 2)
                    data = (body + 4, nbytes)
                else:
                    raise ValueError("chunked datasets are not supported")
        if shape is None or dtype is None or data is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        if data[0] is None:
            return np.zeros(shape, dtype)
        arr = np.frombuffer(self.d, dtype=dtype, count=count, offset=data[0])
        return arr.reshape(shape).copy()

    def walk(self, header_addr=None, prefix=""):
        """Yield (path, ndarray) for every dataset below the given group."""
        header_addr = self.root_header if header_addr is None else header_addr
        kids = self.children(header_addr)
        if not kids:
            arr = self.dataset(header_addr)
            if arr is not None:
                yield prefix, arr
            return
        for name, addr in kids.items():
            yield from self.walk(addr, prefix + "/" + name)


def _pad8(n):
    return (n + 7) // 8 * 8


def _attributes(f, header_addr):
    """{name: value} of the version-1 attribute messages of an object: fixed-length string arrays (Keras'
    ``layer_names`` / ``weight_names``) come back as lists of str, numeric ones as arrays; others are skipped."""
    out = {}
    for mtype, _, body, size in f.messages(header_addr):
        if mtype != 0x0C or f.d[body] != 1:
            continue
        nsz, tsz, ssz = f._u16(body + 2), f._u16(body + 4), f._u16(body + 6)
        off = body + 8
        name = f.d[off:off + nsz].split(b"\x00")[0].decode("utf-8")
        off += _pad8(nsz)
        cls, esize = f.d[off] & 0x0F, f._u32(off + 4)
        bits0 = f.d[off + 1]
        off += _pad8(tsz)
        sver, rank = f.d[off], f.d[off + 1]
        dims = tuple(f._u64(off + (8 if sver == 1 else 4) + 8 * i) for i in range(rank))
        off += _pad8(ssz)
        count = int(np.prod(dims)) if dims else 1
        if cls == 3:                                      # fixed-length strings
            out[name] = [f.d[off + i * esize:off + (i + 1) * esize].split(b"\x00")[0].decode("utf-8") for i in range(count)]
        elif cls in (0, 1) and not (bits0 & 1):
            dt = ({2: np.float16, 4: np.float32, 8: np.float64}[esize] if cls == 1
                  else np.dtype("%s%d" % ("i" if bits0 & 0x08 else "u", esize)))
            out[name] = np.frombuffer(f.d, dtype=dt, count=count, offset=off).reshape(dims).copy()
    return out


def read_keras_weights(path, with_order=False):
    """-> {layer name: {weight name: array}} from a Keras checkpoint (weight name without the ':0' suffix).
    ``with_order``: also return {layer name: [weight names in the file's ``weight_names`` order]} (only for layers
    whose group carries that attribute) and the file's ``layer_names`` list (or None)."""
    f = H5File(path)
    out = {}
    for p, arr in f.walk():
        parts = [q for q in p.split("/") if q]
        if parts and parts[0] == "optimizer_weights":
            continue
        if parts and parts[0] == "model_weights":
            parts = parts[1:]
        if len(parts) < 2:
            continue
        layer, wname = parts[0], parts[-1].split(":")[0]
        out.setdefault(layer, {})[wname] = arr
    if not with_order:
        return out
    root = f.root_header
    kids = f.children(root)
    if "model_weights" in kids:
        root = kids["model_weights"]
        kids = f.children(root)
    layer_names = _attributes(f, root).get("layer_names")
    if not isinstance(layer_names, list):
        layer_names = None
    worder = {}
    for lname, addr in kids.items():
        wn = _attributes(f, addr).get("weight_names")
        if isinstance(wn, list) and wn:
            worder[lname] = [w.split("/")[-1].split(":")[0] for w in wn]
    return out, worder, layer_names


_ORDER = {"kernel": 0, "bias": 1, "gamma": 0, "beta": 1, "moving_mean": 2, "moving_variance": 3}


def _split_auto_name(name):
    """'quantized_conv2d_12' -> ('quantized_conv2d', 12); names without a numeric suffix -> (name, None)."""
    base, _, suf = name.rpartition("_")
    return (base, int(suf)) if base and suf.isdigit() else (name, None)


def load_keras_weights(model, path, by_name=False):
    """``model.load_weights(path)`` (test_resnet.py:66, train.py:113-115).

    Keras assigns a checkpoint's weighted layers to the model's weighted layers IN ORDER and ignores their names
    (``by_name=False``).  Auto-generated names (``<class>_<k>``, k from a per-class process-wide counter) therefore
    carry no absolute meaning: the checkpoint may come from a session that had built other models first, and so may
    this model.  What both sides do agree on is the creation order of the layers of one class -- it is the order of
    the builder's code (models/vgg.py, models/resnet.py) -- so the k-th checkpoint layer of a class is bound to the
    k-th model layer of that class, whatever their counters say (this also absorbs the reference's QuantizedDense
    naming quirk, whose double ``__init__`` makes the first one ``quantized_dense_2``).  Layers with explicit names
    match by name.  Counts per class and every tensor shape are validated; ``by_name=True`` matches names only."""
    tensors, worder, layer_names = read_keras_weights(path, with_order=True)
    file_layers = [n for n in (layer_names or sorted(tensors)) if n in tensors and tensors[n]]
    for n in tensors:                                     # groups the attribute does not list (never in Keras files)
        if n not in file_layers and tensors[n]:
            file_layers.append(n)
    mine = [l for l in model.layers if l.weight_names()]
    binding = {}
    if by_name:
        for l in mine:
            if l.name in tensors:
                binding[l.name] = l.name
    else:
        fgroups, mgroups = {}, {}
        for n in file_layers:
            fgroups.setdefault(_split_auto_name(n)[0] if _split_auto_name(n)[1] is not None else ("=" + n), []).append(n)
        for l in mine:
            mgroups.setdefault(_split_auto_name(l.name)[0] if _split_auto_name(l.name)[1] is not None else ("=" + l.name), []).append(l.name)
        for base, names in mgroups.items():
            have = fgroups.get(base, [])
            if len(have) != len(names):
                raise ValueError("checkpoint %s holds %d weighted %r layers, the model has %d"
                                 % (path, len(have), base.lstrip("="), len(names)))
            if base.startswith("="):
                binding[names[0]] = have[0]
            else:
                have = sorted(have, key=lambda n: _split_auto_name(n)[1])
                for mn, fn in zip(sorted(names, key=lambda n: _split_auto_name(n)[1]), have):
                    binding[mn] = fn
        extra = set(fgroups) - set(mgroups)
        if extra:
            raise ValueError("checkpoint %s holds weighted layers the model lacks: %s" % (path, sorted(b.lstrip("=") for b in extra)))
    for layer in mine:
        src = binding.get(layer.name)
        if src is None:
            if by_name:
                continue
            raise ValueError("checkpoint %s has no weights for layer %s" % (path, layer.name))
        grp, want = tensors[src], layer.weight_names()
        if sorted(grp) != sorted(want):
            raise ValueError("checkpoint layer %s holds %s, layer %s expects %s" % (src, sorted(grp), layer.name, sorted(want)))
        order = worder.get(src)
        if not order or sorted(order) != sorted(want):
            order = sorted(want, key=lambda n: _ORDER[n])
        if order != want:
            # the file lists the tensors in another order than this layer's set_weights takes them: go by name
            order = want
        layer.set_weights([np.asarray(grp[n], np.float32) for n in order])      # validates every shape
    return model
