#!/usr/bin/env python
"""Generate tests/golden/ref_outputs.npz: outputs of the REFERENCE's own
kernels (oracle/_ref, built from /root/reference by oracle/build_ref.py) for
the test configurations.  The library itself cannot be rebuilt where the
reference tree is absent; these vectors pin the oracle port there too.

    python oracle/build_ref.py && python tools/make_ref_outputs.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402

cfgs = [H.golden_config(n) for n in H.golden_names()]
cfgs += [H.example_config("test_sersic_bulge"), H.example_config("full_mock_nopsf"),
         H.example_config("full_mock_psf"), H.example_config("full_mock_psf", ipp=False),
         H.synthetic_config("c4", 64), H.synthetic_config("c5", 64)]
out, meta = {}, {}
for cfg in cfgs:
    m = cfg.oracle(variant="ref")
    value, _ = m.render(cfg.params)
    lnew, model, _ = m.loglike(cfg.params, want_maps=True)
    out[cfg.name + "_value"] = value.astype(np.float32)
    out[cfg.name + "_model"] = model.astype(np.float32)
    meta[cfg.name] = dict(lnew=lnew, objects=cfg.objects, params=[float(p) for p in cfg.params])
    print(f"{cfg.name:32s} lnew {lnew:.9g}")
out["meta"] = np.array(json.dumps(meta))
np.savez_compressed(os.path.join(H.GOLDEN, "ref_outputs.npz"), **out)
