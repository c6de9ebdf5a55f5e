#!/usr/bin/env python
"""Full-size C5 log-likelihood fixture (tests/golden/c5_4096_lnew.npz).

The GPU tests compare images with the oracle at sizes the oracle finishes in
seconds; at 4096^2 (822 M rays per evaluation) this script does the oracle's
work once, in the build container, and commits the answers:

  * the observed image is made WITHOUT any renderer in the loop, so that the
    GPU box can rebuild it bit for bit from a small fixture: the strict-float32
    oracle's model of the C5 scene at 512^2 (committed, 1 MB) is block-replicated
    8 x 8 and divided by 64 (both exact in float32: the 4096^2 scene is the
    512^2 scene magnified 8 times at fixed magnitudes), plus seeded Gaussian
    noise of variance (model + offset)/gain; weight = gain/(image + offset)
    (tests/helpers.py: c5_fixture_observation);
  * lnew at the truth and at two perturbed points from the reference's own
    kernels compiled on the host (oracle/_ref, strict build), from the oracle
    port in float32 and in float64 (the noise floor).

Run where /root/reference exists (oracle/_ref built):  python tools/make_c5_fixture.py
Takes ~10 minutes on 8 cores.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from lensed_b200 import workloads  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def main():
    ncpu = len(os.sched_getaffinity(0))
    small = workloads.c5(512)
    blank = np.zeros((512, 512), np.float32)
    qq, ww = O.quad_rule(small["rule"])
    _, model512, _ = O.Model(small["objects"], blank, blank + 1, qq, ww, psf=small["psf"]).loglike(small["truth"], want_maps=True)
    model512 = np.asarray(model512, np.float32)
    w = workloads.c5(4096)
    image, weight = H.c5_fixture_observation(model512)
    P = np.concatenate([w["truth"][None, :], workloads.param_batch(w, 2)])
    out = dict(model512=model512, params=P, image_sha256=np.array(hashlib.sha256(image.tobytes()).hexdigest()),
               weight_sha256=np.array(hashlib.sha256(weight.tobytes()).hexdigest()))
    for variant in ("ref", "strict", "f64"):
        if not O.available(variant):
            print("skipping", variant)
            continue
        lib = O.lib(variant)
        lib.orc_set_threads(ncpu)
        om = O.Model(w["objects"], image, weight, qq, ww, psf=w["psf"], _lib=lib)
        vals = []
        for p in P:
            t0 = time.time()
            vals.append(om.loglike(p))
            print(variant, vals[-1], f"{time.time() - t0:.1f} s", flush=True)
        out["lnew_" + variant] = np.array(vals, np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c5_4096_lnew.npz"), **out)
    print({k: v for k, v in out.items() if k.startswith("lnew")}, "chi2/pixel", -2*out["lnew_f64"]/image.size)


if __name__ == "__main__":
    main()
