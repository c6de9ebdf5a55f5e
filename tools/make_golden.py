#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ from the reference tree.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):

    python tools/make_golden.py

Outputs
  tests/golden/quad_rules.npz    float32 (qq, ww) of every rule exactly as the
                                 reference's quad_rule() packs them
                                 (src/quadrature.c:32-43 over src/quad/*.c tables)
  tests/golden/ref_goldens.npz   the 16 golden model images of the reference's
                                 own test-suite (tests/{lens,source,foreground}/*.fits,
                                 primary HDU) as float32 + their parameters
  tests/golden/examples.npz      input data of examples/*.ini (C1-C3): images,
                                 PSFs, gain/offset, object lists, prior ranges
  tests/golden/objects/          the reference's objects/*.cl, byte for byte (test
                                 input: the plugin files a Lensed user has)
  tests/golden/ref_outputs.npz   (only if oracle/_ref is built) outputs of the
                                 reference's own kernels compiled on the host for
                                 the named configurations -- the parity vectors
"""
import glob
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("LENSED_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from lensed_b200.fits import read_hdus, read_image  # noqa: E402


def parse_c_table(path, symbol):
    text = re.sub(r"//.*", "", open(path).read())
    m = re.search(r"%s\s*\[\](?:\[2\])?\s*=\s*\{(.*?)\};" % symbol, text, re.S)
    return np.array([float(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", m.group(1))])


def quad_rules():
    out = {}
    for name in ("point", "sub2", "sub4", "gm75", "g3k7", "g5k11", "g7k15"):
        path = os.path.join(REF, "src", "quad", name + ".c")
        up = name.upper()
        pts = parse_c_table(path, f"QUAD_{up}_PTS").reshape(-1, 2)
        wht = parse_c_table(path, f"QUAD_{up}_WHT")
        err = parse_c_table(path, f"QUAD_{up}_ERR")
        # quad_rule() with sx = sy = 1, and with a non-trivial scale
        for tag, sx, sy in (("", 1.0, 1.0), ("_scaled", 0.75, -1.25)):
            qq = np.stack([(sx*pts[:, 0]).astype(np.float32), (sy*pts[:, 1]).astype(np.float32)], 1)
            ww = np.stack([wht.astype(np.float32), err.astype(np.float32)], 1)
            out[f"{name}{tag}_qq"] = qq
            out[f"{name}{tag}_ww"] = ww
    np.savez_compressed(os.path.join(OUT, "quad_rules.npz"), **out)
    print("quad_rules.npz", len(out), "arrays")


def read_ini(path):
    grp, objs, pri, opts = None, [], {}, {}
    for line in open(path):
        line = line.split(";")[0].strip()
        if not line:
            continue
        if line.startswith("["):
            grp = line.strip("[]")
            continue
        k, v = [s.strip() for s in line.split("=", 1)]
        if grp is None or grp == "options":
            opts[k] = v
        elif grp == "objects":
            objs.append((k, v))
        elif grp == "priors":
            pri[k] = v
    return opts, objs, pri


def ref_goldens():
    out = {}
    meta = {}
    for ini in sorted(glob.glob(os.path.join(REF, "tests", "*", "*.ini"))):
        opts, objs, pri = read_ini(ini)
        name = os.path.splitext(os.path.basename(ini))[0]
        hdus = read_hdus(os.path.join(os.path.dirname(ini), opts["image"]))
        img = next(d for h, d in hdus if d is not None and d.ndim == 2)
        out[name] = np.asarray(img, dtype=np.float32)
        meta[name] = dict(objects=objs, priors=pri, weight=float(opts["weight"]), rule=opts.get("rule", "g3k7"))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "ref_goldens.npz"), **out)
    print("ref_goldens.npz", len(meta), "configs")


def examples():
    out = {}
    meta = {}
    ex = os.path.join(REF, "examples")
    for name in ("test_sersic_bulge", "full_mock_nopsf", "full_mock_psf"):
        opts, objs, pri = read_ini(os.path.join(ex, name + ".ini"))
        img, pcs = read_image(os.path.join(ex, opts["image"]))
        out[name + "_image"] = img
        m = dict(objects=objs, priors=pri, gain=float(opts["gain"]), offset=float(opts["offset"]), pcs=pcs,
                 rule=opts.get("rule", "g3k7"), psf=None)
        if "psf" in opts:
            psf, _ = read_image(os.path.join(ex, opts["psf"]))
            out[name + "_psf"] = psf            # raw, un-normalised (read_psf normalises, src/data.c:354-370)
            m["psf"] = opts["psf"]
        meta[name] = m
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "examples.npz"), **out)
    print("examples.npz", list(meta))


def reference_objects():
    """The reference's objects/*.cl, verbatim, as the plugin directory every GPU
    test, smoke() and bench.py run on (tests/golden/objects/README.md)."""
    import hashlib
    import shutil
    dst = os.path.join(OUT, "objects")
    os.makedirs(dst, exist_ok=True)
    sums = []
    for path in sorted(glob.glob(os.path.join(REF, "objects", "*.cl"))):
        shutil.copyfile(path, os.path.join(dst, os.path.basename(path)))
        sums.append(f"{hashlib.sha256(open(path, 'rb').read()).hexdigest()}  {os.path.basename(path)}\n")
    shutil.copyfile(os.path.join(REF, "LICENSE.txt"), os.path.join(dst, "LICENSE.txt"))
    open(os.path.join(dst, "SHA256SUMS"), "w").writelines(sums)
    print("objects/", len(sums), "files")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    reference_objects()
    quad_rules()
    ref_goldens()
    examples()
