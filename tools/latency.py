#!/usr/bin/env python
"""Single-point latency of lcu_loglike (the sampler's one-point callback) on the
reference's example-sized images, with and without the CUDA-graph path, next to
the reference CPU build.  Writes gpurun_out/latency.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")      # the reference's objects/*.cl, verbatim
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lensed_b200 as L
import helpers as H

out = {}
ctx = L.Context(device=0, objects_dir=OBJECTS_DIR)
for name in ("full_mock_nopsf", "full_mock_psf", "test_sersic_bulge"):
    cfg = H.example_config(name)
    res = {}
    for mode in ("graph", "graph_one_kernel", "graph_one_point_per_pass", "graph_folded_set_params", "plain"):
        for env in ("LCU_NO_GRAPH", "LCU_FOLD_SETTER", "LCU_NO_SPLIT_PAIR", "LCU_FUSED_POINT"):
            os.environ.pop(env, None)
        if mode == "graph_one_kernel":
            os.environ["LCU_FUSED_POINT"] = "1"
        if mode == "plain":
            os.environ["LCU_NO_GRAPH"] = "1"
        elif mode == "graph_folded_set_params":
            os.environ["LCU_FOLD_SETTER"] = "1"
        elif mode == "graph_one_point_per_pass":
            os.environ["LCU_NO_SPLIT_PAIR"] = "1"
        m = cfg.product(ctx, flags=L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
        for _ in range(20):
            v = m.loglike(cfg.params)
        n = 2000
        t0 = time.perf_counter()
        for _ in range(n):
            v = m.loglike(cfg.params)
        dt = (time.perf_counter() - t0)/n
        res[mode] = dict(us_per_eval=dt*1e6, evals_per_s=1/dt, lnew=v)
        if mode == "graph":
            # two evaluations in flight (lcu_loglike_async / lcu_loglike_wait)
            t0 = time.perf_counter()
            t = m.loglike_async(cfg.params)
            for _ in range(n - 1):
                tn = m.loglike_async(cfg.params)
                v2 = m.loglike_wait(t)
                t = tn
            v2 = m.loglike_wait(t)
            dt = (time.perf_counter() - t0)/n
            assert v2 == v
            res["graph_two_in_flight"] = dict(us_per_eval=dt*1e6, evals_per_s=1/dt, lnew=v2)
        m.close()
    assert res["graph"]["lnew"] == res["plain"]["lnew"] == res["graph_folded_set_params"]["lnew"]
    assert res["graph"]["lnew"] == res["graph_one_point_per_pass"]["lnew"]
    assert res["graph"]["lnew"] == res["graph_one_kernel"]["lnew"]
    for env in ("LCU_NO_GRAPH", "LCU_FOLD_SETTER", "LCU_NO_SPLIT_PAIR", "LCU_FUSED_POINT"):
        os.environ.pop(env, None)
    try:
        from oracle import pyoracle as O
        variant = next(v for v in ("ref_simd512", "ref_simd", "ref_fast") if O.available(v))
        om = cfg.oracle(variant=variant, lib=O.lib(variant))
        res["cpu_build"] = variant
        om.loglike(cfg.params)
        n = 200
        t0 = time.perf_counter()
        for _ in range(n):
            om.loglike(cfg.params)
        dt = (time.perf_counter() - t0)/n
        res["cpu_reference"] = dict(us_per_eval=dt*1e6, evals_per_s=1/dt, threads=om.L.orc_max_threads())
    except Exception as e:
        res["cpu_reference"] = repr(e)
    out[name] = res
    print(name, json.dumps(res), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "latency.json"), "w"), indent=1)
