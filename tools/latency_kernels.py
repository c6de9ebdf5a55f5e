#!/usr/bin/env python
"""A few single-point evaluations of the reference's example models, for an ncu
launch list (per-kernel durations on the one-point latency path):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/latency_launches.csv \
      python tools/latency_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")      # the reference's objects/*.cl, verbatim
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["LCU_NO_GRAPH"] = "1"
import lensed_b200 as L
import helpers as H

ctx = L.Context(device=0, objects_dir=OBJECTS_DIR)
for name in ("full_mock_nopsf", "full_mock_psf", "test_sersic_bulge"):
    cfg = H.example_config(name)
    m = cfg.product(ctx, flags=L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
    for _ in range(6):
        v = m.loglike(cfg.params)
    print(name, v)
    m.close()
