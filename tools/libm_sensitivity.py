#!/usr/bin/env python
"""How far does a rendered image move when the math library is a different one?

Builds copies of the CPU oracle (test infrastructure) whose atan2f / sinf, cosf /
powf results are perturbed by a pseudo-random +-k ulp (glibc is < 1 ulp on
these; CUDA's libdevice documents 2-4 ulp; an OpenCL CPU runtime under
-cl-fast-relaxed-math is looser still) and prints, per scene, the maximum and
99.9th-percentile per-pixel relative change next to the oracle's float32-vs-
float64 distance.  Scenes where one ulp moves a pixel by more than 1e-5 cannot
meet the 1e-5 bound with *any* second implementation; tests/test_gpu_parity.py
screens its random scenes with this tool (CPU only):

    python tools/libm_sensitivity.py 13 20 2 5 21
"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import pyoracle as O

PERT_H = r"""
#include <math.h>
#include <stdint.h>
#include <string.h>
static inline float pert_(float r, float a, float b, float k)
{
    uint32_t ia, ib; memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    uint32_t h = ia*2654435761u ^ (ib + 0x9e3779b9u)*2246822519u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    float u = (float)(h & 0xffff)/32768.0f - 1.0f;
    return r*(1.0f + k*u*5.9604645e-8f);
}
#ifdef P_ATAN2
static inline float my_atan2f(float y, float x){ return pert_(atan2f(y, x), y, x, P_ATAN2); }
#define atan2f my_atan2f
#endif
#ifdef P_SINCOS
static inline float my_sinf(float x){ return pert_(sinf(x), x, 1.f, P_SINCOS); }
static inline float my_cosf(float x){ return pert_(cosf(x), x, 2.f, P_SINCOS); }
#define sinf my_sinf
#define cosf my_cosf
#endif
#ifdef P_POW
static inline float my_powf(float x, float y){ return pert_(powf(x, y), x, y, P_POW); }
#define powf my_powf
#endif
"""
VARIANTS = ["ATAN2=1", "ATAN2=3", "SINCOS=1", "SINCOS=2", "POW=4"]


def main():
    seeds = [int(a) for a in sys.argv[1:]] or [13, 20, 2, 5, 21]
    tmp = tempfile.mkdtemp(prefix="libm_sens_")
    open(os.path.join(tmp, "pert.h"), "w").write(PERT_H)
    for v in VARIANTS:
        out = os.path.join(tmp, f"liboracle_{v.replace('=', '_')}.so")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-w",
                        "-include", os.path.join(tmp, "pert.h"), f"-DP_{v}", "-o", out,
                        os.path.join(ROOT, "oracle", "lensed_oracle.c"), "-lm"], check=True)
        O._VARIANTS[v] = out
    for seed in seeds:
        cfg = H.random_config(seed)
        v, _ = cfg.oracle().render(cfg.params)
        v64, _ = cfg.oracle(variant="f64").render(cfg.params)
        line = f"seed {seed:3d} {'+'.join(cfg.objects)}: f32-f64 max {H.rel_err(v, v64).max():.2e}"
        for name in VARIANTS:
            vp, _ = cfg.oracle(variant=name).render(cfg.params)
            r = H.rel_err(vp, v)
            line += f" | {name} ulp: max {r.max():.2e} p99.9 {np.quantile(r, 0.999):.2e}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
