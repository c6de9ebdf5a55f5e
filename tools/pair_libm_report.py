#!/usr/bin/env python
"""CPU-side report on -DLCU_PF_LIBM_PAIR=1 (shim.cuh: atan2 / sincos / pow / powr
of the two-rays kernel with packed arithmetic).  No GPU needed:

  * SASS (NVRTC -> sm_100a cubin -> cuobjdump) of the written-out functions:
    the packed instructions must be exactly those the source names (ptxas must
    not have contracted a packed multiply into a packed add);
  * executed PTX instructions per pair of C5 rays in the interpreter
    (tools/ptx_emu.py), switch off and on;
  * registers of lcu_render_pair for the C5 program, off and on.

Usage: python tools/pair_libm_report.py > profiles/<name>.txt
"""
import collections
import os
import random
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
from cuda.bindings import nvrtc  # noqa: E402

import ptx_emu as E  # noqa: E402
import helpers as H  # noqa: E402
import lensed_b200 as L  # noqa: E402
import test_pair_math as TM  # noqa: E402
import test_pair_rays as TR  # noqa: E402


def cubin_of(src, extra):
    shim = open(os.path.join(ROOT, "lensed_b200", "kernel", "shim.cuh"), "rb").read()
    _, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"t.cu", 1, [shim], [b"shim.cuh"])
    opts = [o.replace("compute_100a", "sm_100a").encode() for o in TM.OPTIONS + list(extra)]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    cub = b" "*n
    nvrtc.nvrtcGetCUBIN(prog, cub)
    return cub


def sass_counts(cubin):
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(cubin)
        f.flush()
        txt = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True, check=True).stdout
    out = {}
    for m in re.finditer(r"Function : (\w+)(.*?)(?=Function :|\Z)", txt, re.S):
        ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", m.group(2), re.M))
        out[m.group(1)] = ops
    return out


def main():
    print("# -DLCU_PF_LIBM_PAIR=1: CPU-side evidence (tools/pair_libm_report.py)\n")
    print("## packed instructions in the SASS of the written-out pair functions (sm_100a, product options)")
    print("## expected from the source -- atan2: FMUL2 3 (1 without .ftz), FADD2 4, FFMA2 5; sincos: FMUL2 2, FFMA2 11;")
    print("## powr: FMUL2 8 (2 without .ftz), FADD2 11, FFMA2 19\n")
    counts = sass_counts(cubin_of(TM.HARNESS, ["-DLCU_PF_LIBM_PAIR=1"]))
    for fn in ("p_atan2", "p_sincos", "p_powr"):
        ops = counts[fn]
        packed = {k: v for k, v in sorted(ops.items()) if re.match(r"(FFMA2|FMUL2|FADD2)", k)}
        print(f"{fn:10s} {packed}")

    print("\n## executed PTX instructions per pair of rays of the C5 scene (interpreter, 20 random pairs, fast build)")
    ctx = L.Context(device=-1, objects_dir=os.path.join(ROOT, "tests", "golden", "objects"))
    cfg = H.synthetic_config("c5", 64)
    flags = L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH
    src, block = TR._program(cfg, ctx, flags)
    cnt = collections.Counter()
    orig = E._Machine.step

    def step(self, op, rest, line):
        cnt[op] += 1
        return orig(self, op, rest, line)
    E._Machine.step = step
    for extra in ([], ["-DLCU_PF_LIBM_PAIR=1"]):
        M = E.Module(TR._ray_ptx(src, flags, extra))
        rng = random.Random(3)
        tot = collections.Counter()
        n = 20
        for _ in range(n):
            xs = [E.f2b(rng.uniform(1, 64)) for _ in range(4)]
            cnt.clear()
            M.run("p_compute", [xs, block], 2)
            tot.update(cnt)
        work = sum(v for k, v in tot.items() if not k.startswith(("mov", "ld.", "st.", "cvta")))
        arith = sum(v for k, v in tot.items() if k.split(".")[0] in ("add", "sub", "mul", "fma") and "f32" in k)
        packed = sum(v for k, v in tot.items() if "f32x2" in k)
        print(f"{'on ' if extra else 'off'}: instructions other than moves / loads {work/n:6.1f}   FP32 add/mul/fma {arith/n:6.1f}"
              f" (packed {packed/n:6.1f})   IEEE divisions {tot['div.rn.ftz.f32']/n:.0f} (uniform series coefficients: hoisted in the kernel)")
    E._Machine.step = orig

    print("\n## lcu_render_pair of the C5 program (4096^2, fast build): registers, stack bytes")
    from lensed_b200 import workloads as W
    w = W.c5(4096)
    img = np.zeros((w["height"], w["width"]), np.float32)
    for tag, env in (("off", "-DLCU_PF_LIBM_PAIR=0"), ("on ", "-DLCU_PF_LIBM_PAIR=1")):
        os.environ["LCU_NVRTC_FLAGS"] = env
        m = L.Model(L.Context(device=-1, objects_dir=os.path.join(ROOT, "tests", "golden", "objects")), w["objects"], img, img, rule=w["rule"], psf=w["psf"], flags=flags)
        print(tag, m.kernel_usage("lcu_render_pair"))
    os.environ.pop("LCU_NVRTC_FLAGS", None)

    print("""
## bit parity of the switch, all on the CPU (PTX interpreter, tools/ptx_emu.py)
function level  tests/test_pair_math.py            atan2 / sincos / sin / cos / pow / powr: 600-850 argument pairs each in the suite
                                                   (random, extreme, special), 12 000 more per function by hand: every lane the scalar bits
ray level       tests/test_pair_rays.py            lcu_compute2 = lcu_compute on rays through the four EPL test configurations and C5,
                                                   also with the reference's own objects/*.cl
kernel level    tests/test_kernels_interpreted.py  lcu_render_pair = lcu_render_s1 (image and chi^2 partial sums) on an EPL scene and on
                                                   random models with power-law lenses; 42 more random models by hand
ptxas           tests/test_pair_rays.py            packed instruction counts of lcu_render_pair equal in PTX and SASS (no contraction)
hardware        tests/test_gpu_parity.py           test_two_rays_per_thread_same_bits_packed_libm_switch: the switch in either
                                                   position against the one-ray kernel, bit for bit (on by default since round 2)""")


if __name__ == "__main__":
    main()
