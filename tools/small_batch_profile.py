#!/usr/bin/env python
"""Stage breakdown of batched evaluation on the reference's small example images
(the sampler's regime: C2 / C3 / C1, B = 16 ... 4096 points per call).
Writes gpurun_out/small_batch.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lensed_b200 as L
import helpers as H

out = {}
ctx = L.Context(device=0, objects_dir=OBJECTS_DIR)
for name in ("full_mock_nopsf", "full_mock_psf", "test_sersic_bulge"):
    cfg = H.example_config(name)
    m = cfg.product(ctx, flags=L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
    rays = cfg.image.size*m.nq
    res = {}
    for B in (16, 64, 256, 1024, 4096):
        rng = np.random.default_rng(B)
        P = (cfg.params[None, :]*(1 + 1e-3*rng.uniform(-1, 1, (B, cfg.params.size)))).astype(np.float32)
        for _ in range(3):
            m.loglike_batch(P)
        n = max(3, int(2e5/B))
        t0 = time.perf_counter()
        for _ in range(n):
            m.loglike_batch(P)
        dt = (time.perf_counter() - t0)/n
        m.profile(True)
        for _ in range(5):
            m.loglike_batch(P)
        pr = m.profile_get()
        m.profile(False)
        res[B] = dict(evals_per_s=B/dt, grays_per_s=B/dt*rays/1e9, ms_per_call=dt*1e3,
                      stage_ms={k: pr[k]/5 for k in ("set_params_ms", "render_ms", "convolve_ms", "reduce_ms", "upload_ms", "download_ms")})
        print(name, B, json.dumps(res[B]), flush=True)
    out[name] = res
    m.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "small_batch.json"), "w"), indent=1)
