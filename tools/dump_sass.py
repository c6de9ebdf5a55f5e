#!/usr/bin/env python
"""SASS evidence for the render kernel (no GPU needed): compiles the C4 and C5
benchmark models (fast build, the reference's object files) with NVRTC for
sm_100a, disassembles lcu_render_pair with cuobjdump and writes

    profiles/<prefix>_sass_render_pair_<workload>.txt

= resource usage, the opcode histogram of the inner quadrature loop (the
innermost backward BRA.U of the kernel; slow-path blocks of the IEEE division /
sqrt / Payne-Hanek reduction are inside its address range but off the hot path,
so they are listed separately), and the loop's instructions as they are.

    python tools/dump_sass.py r02
"""
import collections
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lensed_b200 as L  # noqa: E402
from lensed_b200 import workloads  # noqa: E402

OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")


def main(prefix):
    ctx = L.Context(device=-1, objects_dir=OBJECTS_DIR)
    for tag, w in (("c4", workloads.c4(1024)), ("c5", workloads.c5(4096))):
        img = np.zeros((w["height"], w["width"]), np.float32)
        m = L.Model(ctx, w["objects"], img, img, rule=w["rule"], psf=w["psf"], flags=L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
        regs, stack = m.kernel_usage("lcu_render_pair")
        cubin = f"/tmp/lcu_{tag}_{os.getpid()}.cubin"
        open(cubin, "wb").write(m.cubin)
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", "lcu_render_pair", cubin], capture_output=True, text=True, check=True).stdout
        os.remove(cubin)
        ins = []
        for line in sass.splitlines():
            mm = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if mm:
                ins.append((int(mm.group(1), 16), mm.group(2).strip()))
        loops = []
        for a, t in ins:
            mm = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
            if mm and int(mm.group(1), 16) < a and "BRA.U" in t:
                loops.append((a - int(mm.group(1), 16), int(mm.group(1), 16), a))
        span, lo, hi = min(l for l in loops if l[0] > 0x400)
        body = [(a, t) for a, t in ins if lo <= a <= hi]

        def op(t):
            return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
        hist = collections.Counter(op(t).split(".")[0] for _, t in body)
        packed = {k: hist[k] for k in ("FFMA2", "FMUL2", "FADD2")}
        out = os.path.join(ROOT, "profiles", f"{prefix}_sass_render_pair_{tag}.txt")
        with open(out, "w") as f:
            f.write(f"# lcu_render_pair, {w['name']} ({'+'.join(w['objects'])}, rule {w['rule']}), fast build, sm_100a, "
                    f"objects = the reference's files\n")
            f.write(f"# registers {regs}, stack {stack} bytes; kernel {len(ins)} instructions; inner quadrature loop "
                    f"0x{lo:x}..0x{hi:x} = {len(body)} instructions in its address range (hot path + slow-path blocks)\n")
            f.write(f"# packed FP32 in the loop range: {packed}; MUFU {hist['MUFU']}; scalar FFMA/FMUL/FADD "
                    f"{hist['FFMA']}/{hist['FMUL']}/{hist['FADD']}; no HMMA / tensor instructions (nothing on this path is a contraction)\n")
            f.write("# opcode histogram of the loop range: " + ", ".join(f"{k} {v}" for k, v in hist.most_common()) + "\n\n")
            for a, t in body:
                f.write(f"/*{a:04x}*/  {t}\n")
        print(out, regs, stack, len(body), packed)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
