#!/usr/bin/env python
"""Parity + speed of the relaxed-math flag levels (exploratory)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")      # the reference's objects/*.cl, verbatim
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lensed_b200 as L
from lensed_b200 import workloads
import helpers as H

ctx = L.Context(device=0, objects_dir=OBJECTS_DIR)
cfgs = [H.golden_config(n) for n in ("sie", "epl", "point_mass", "nsie", "sersic", "devauc", "gauss", "sis_plus_shear")]
cfgs += [H.example_config("test_sersic_bulge"), H.example_config("full_mock_nopsf"), H.example_config("full_mock_psf")]
cfgs += [H.synthetic_config("c4", 128), H.synthetic_config("c4", 256), H.synthetic_config("c5", 256), H.synthetic_config("c5", 512),
         H.synthetic_config("c4", 256, psf=False), H.synthetic_config("c5", 256, psf=False)]
refs = []
for cfg in cfgs:
    om = cfg.oracle()
    value, _ = om.render(cfg.params)
    lnew, model, _ = om.loglike(cfg.params, want_maps=True)
    refs.append((value, model, lnew))
out = {}
for flags in (0, 4, 36):
    worst_raw = worst_model = worst_ln = 0
    for cfg, (value, model, lnew) in zip(cfgs, refs):
        m = cfg.product(ctx, flags=flags)
        o = m.render(cfg.params, error=False, chi=False)
        got = m.loglike(cfg.params)
        rr = H.rel_err(o["raw"], value); rm = H.rel_err(o["model"], model)
        lr = abs(got - lnew)/abs(lnew) if cfg.name.startswith(("C4", "C5", "full", "test")) else 0
        print(f"flags {flags:2d} {cfg.name:24s} raw max {rr.max():.2e} p99.9 {np.quantile(rr,0.999):.2e} model max {rm.max():.2e} lnew rel {lr:.2e}", flush=True)
        worst_raw = max(worst_raw, rr.max()); worst_model = max(worst_model, rm.max()); worst_ln = max(worst_ln, lr)
        m.close()
    # speed C4 1024 B=8
    w = workloads.c4(1024)
    blank = np.zeros((1024, 1024), np.float32)
    m = L.Model(ctx, w["objects"], blank, blank + 1, rule=w["rule"], psf=w["psf"], flags=flags)
    P = workloads.param_batch(w, 8)
    m.loglike_batch(P)
    t0 = time.perf_counter(); m.loglike_batch(P); m.loglike_batch(P); dt = (time.perf_counter() - t0)/2
    m.close()
    w5 = workloads.c5(2048)
    blank = np.zeros((2048, 2048), np.float32)
    m = L.Model(ctx, w5["objects"], blank, blank + 1, rule=w5["rule"], psf=w5["psf"], flags=flags)
    P5 = workloads.param_batch(w5, 4)
    m.loglike_batch(P5)
    t0 = time.perf_counter(); m.loglike_batch(P5); dt5 = (time.perf_counter() - t0)
    m.close()
    out[flags] = dict(raw=worst_raw, model=worst_model, lnew=worst_ln, c4_evals_s=8/dt, c5_2048_evals_s=4/dt5)
    print(f"== flags {flags}: worst raw {worst_raw:.2e} model {worst_model:.2e} lnew {worst_ln:.2e} | C4 {8/dt:.1f} evals/s | C5@2048 {4/dt5:.1f} evals/s", flush=True)
json.dump({str(k): v for k, v in out.items()}, open(os.path.join(ROOT, "gpurun_out", "probe2.json"), "w"), indent=1)
