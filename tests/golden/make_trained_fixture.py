#!/usr/bin/env python
"""Convert two of the reference's trained checkpoints (results/RESNET3/weights_44.hdf5 = full-qnn w4a4,
weights_bb.hdf5 = full-bnn, weights_ff.hdf5 = float; all from the older, biased ResNet revision -- SURVEY.md finding 7) into compact
.npz weight lists with the repo's pure-Python HDF5 reader, so that GPU parity can run on REAL weight / BN
statistics without /root/reference.  Run in the build container only:

    python tests/golden/make_trained_fixture.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("QNNB_REFERENCE", "/root/reference")

import qnn_b200 as q  # noqa: E402

CASES = {"44": ("full-qnn", 4, 4), "bb": ("full-bnn", 4, 4), "ff": ("float", 4, 4)}
ONLY = set(sys.argv[1:])

for code, (nt, wb, ab) in CASES.items():
    if ONLY and code not in ONLY:
        continue
    cf = types.SimpleNamespace(network_type=nt, wbits=wb, abits=ab, architecture='RESNET', dataset='CIFAR-10', dim=32, channels=3,
                               classes=10, nres=3, pfilt=1, kernel_initializer='he_normal', kernel_regularizer=1e-4)
    q.reset_names()
    model = q.build_model(cf, legacy_resnet=True)
    model.load_weights(os.path.join(REF, "results", "RESNET3", "weights_%s.hdf5" % code))
    ws = model.get_weights()
    names = []
    for l in model.layers:
        names.extend("%s/%s" % (l.name, n) for n in l.weight_names())
    assert len(names) == len(ws) and sum(w.size for w in ws) == 274442
    out = {"w%03d" % i: w for i, w in enumerate(ws)}
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "trained_resnet3_%s.npz" % code), **out)
    print(code, len(ws), "arrays", sum(w.size for w in ws), "params")
