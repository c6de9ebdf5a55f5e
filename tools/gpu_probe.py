#!/usr/bin/env python
"""Exploratory GPU run: parity statistics and first timings in one go.
Writes gpurun_out/probe.json.  Not a test and not the bench."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")      # the reference's objects/*.cl, verbatim
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import lensed_b200 as L  # noqa: E402
from lensed_b200 import workloads  # noqa: E402
import helpers as H  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

out = {}
ctx = L.Context(device=0, objects_dir=OBJECTS_DIR)
out["fp32_peak_tflops"] = ctx.fp32_peak_tflops()
print("fp32 peak", out["fp32_peak_tflops"], flush=True)
print("host threads", O.lib().orc_max_threads(), flush=True)
out["host_threads"] = O.lib().orc_max_threads()


def parity(cfg, flags=0, tag=""):
    om = cfg.oracle()
    of = cfg.oracle(variant="f64")
    m = cfg.product(ctx, flags=flags)
    value, error = om.render(cfg.params)
    v64, _ = of.render(cfg.params)
    lnew, model, chi = om.loglike(cfg.params, want_maps=True)
    o = m.render(cfg.params)
    got = m.loglike(cfg.params)
    r_raw = H.rel_err(o["raw"], value)
    r_mod = H.rel_err(o["model"], model)
    floor = H.rel_err(value, v64)
    gpu64 = H.rel_err(o["raw"], v64)
    res = dict(raw_max=float(r_raw.max()), raw_p999=float(np.quantile(r_raw, 0.999)), raw_med=float(np.median(r_raw)),
               model_max=float(r_mod.max()), floor_max=float(floor.max()), floor_p999=float(np.quantile(floor, 0.999)),
               gpu_vs_f64_max=float(gpu64.max()),
               lnew=got, lnew_ref=lnew, lnew_rel=float(abs(got - lnew)/max(abs(lnew), 1e-300)))
    print(f"{cfg.name:34s}{tag:6s} raw max {res['raw_max']:.2e} p99.9 {res['raw_p999']:.2e} | model max {res['model_max']:.2e} "
          f"| f32-vs-f64 floor {res['floor_max']:.2e} | gpu-vs-f64 {res['gpu_vs_f64_max']:.2e} | lnew rel {res['lnew_rel']:.2e}", flush=True)
    m.close()
    return res


out["parity"] = {}
cfgs = [H.golden_config(n) for n in H.golden_names()]
cfgs += [H.example_config("test_sersic_bulge"), H.example_config("full_mock_nopsf"), H.example_config("full_mock_psf")]
cfgs += [H.synthetic_config("c4", 128), H.synthetic_config("c5", 128), H.synthetic_config("c4", 256)]
for cfg in cfgs:
    for flags, tag in ((0, ""), (L.LCU_FAST_MATH, " fast")):
        try:
            out["parity"][cfg.name + tag] = parity(cfg, flags, tag)
        except Exception as e:  # keep going: this is a probe
            print(cfg.name, tag, "FAILED", repr(e)[:500], flush=True)
            out["parity"][cfg.name + tag] = dict(error=repr(e)[:500])


def timing(w, nb_list, flags=0, tag="", env=None):
    size = w["width"]
    blank = np.zeros((size, size), np.float32)
    if env:
        os.environ.update(env)
    m0 = L.Model(ctx, w["objects"], blank, blank + 1, rule=w["rule"], psf=w["psf"], flags=flags)
    truth = m0.render(w["truth"], raw=False, error=False, chi=False)["model"]
    m0.close()
    image, weight = workloads.observe(truth, w["noise_seed"])
    m = L.Model(ctx, w["objects"], image, weight, rule=w["rule"], psf=w["psf"], flags=flags)
    res = {}
    for nb in nb_list:
        P = workloads.param_batch(w, nb)
        m.loglike_batch(P)
        m.profile(True)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ln = m.loglike_batch(P)
        dt = (time.perf_counter() - t0)/reps
        pr = m.profile_get()
        m.profile(False)
        res[nb] = dict(ms=dt*1e3, evals_per_s=nb/dt, chi2_dof=float(-2*ln[0]/image.size),
                       render_ms=pr["render_ms"]/reps, convolve_ms=pr["convolve_ms"]/reps, reduce_ms=pr["reduce_ms"]/reps,
                       set_params_ms=pr["set_params_ms"]/reps)
        print(f"{w['name']}{tag} B={nb}: {dt*1e3:.2f} ms/batch, {nb/dt:.1f} evals/s | render {res[nb]['render_ms']:.2f} "
              f"conv {res[nb]['convolve_ms']:.3f} reduce {res[nb]['reduce_ms']:.3f} set {res[nb]['set_params_ms']:.3f} ms", flush=True)
    m.close()
    if env:
        for k in env:
            os.environ.pop(k, None)
    return res


out["timing"] = {}
w4 = workloads.c4(1024)
out["timing"]["c4"] = timing(w4, [1, 8, 32])
out["timing"]["c4_fast"] = timing(w4, [8], flags=L.LCU_FAST_MATH, tag=" fast")
out["timing"]["c4_shared"] = timing(w4, [8], flags=L.LCU_OBJ_SHARED, tag=" smem-objs")
out["timing"]["c4_fastmath"] = timing(w4, [8], tag=" use_fast_math", env={"LCU_NVRTC_FLAGS": "--use_fast_math"})
w5 = workloads.c5(4096)
out["timing"]["c5"] = timing(w5, [1, 4])
w = workloads.c4(100)
w["rule"] = "g3k7"
out["timing"]["small100"] = timing(w, [1, 64, 512])
for s in ("1", "8"):
    out["timing"]["small100_split" + s] = timing(w, [1], tag=" split" + s, env={"LCU_SPLIT": s})

# CPU baseline sample: oracle fast build, all threads, 128 rows of C4
try:
    fast = O.lib("fast")
    band = H.Config("band", w4["objects"], w4["truth"], np.zeros((128, 1024), np.float32), np.ones((128, 1024), np.float32),
                    rule=w4["rule"], pcs=(1.0, 449.0, 1.0, 1.0))
    om = band.oracle(lib=fast)
    om.render(w4["truth"])
    t0 = time.perf_counter()
    om.render(w4["truth"])
    dt = time.perf_counter() - t0
    rays = 128*1024*225
    out["cpu_port_fast"] = dict(threads=fast.orc_max_threads(), rays_per_s=rays/dt, evals_per_s_c4=rays/dt/(1024*1024*225))
    print("cpu port fast:", out["cpu_port_fast"], flush=True)
    oms = band.oracle()
    t0 = time.perf_counter()
    oms.render(w4["truth"])
    dt = time.perf_counter() - t0
    out["cpu_port_strict"] = dict(threads=O.lib().orc_max_threads(), rays_per_s=rays/dt)
    print("cpu port strict:", out["cpu_port_strict"], flush=True)
except Exception as e:
    print("cpu baseline failed", repr(e))

out["launches"] = L.launch_count()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
print("done")
