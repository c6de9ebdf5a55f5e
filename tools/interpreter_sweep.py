#!/usr/bin/env python
"""Wider sweeps of the interpreted-kernel checks than the test suite runs
(tests/test_kernels_interpreted.py), for use by hand -- no GPU needed:

  python tools/interpreter_sweep.py conv   FIRST LAST          random image / PSF shapes through both convolution kernels
  python tools/interpreter_sweep.py render FIRST LAST [-D...]  random object combinations: pair kernel = one-ray kernel, oracle bound
  python tools/interpreter_sweep.py full   FIRST LAST          random models with random PSFs: render + convolve + reduce against the oracle

FIRST / LAST are seeds; extra arguments are NVRTC options (e.g. -DLCU_PF_LIBM_PAIR=1).
Memory is strict (ptx_emu.StrictMemory): a load from an unwritten address raises.
The results on record (round 1): conv 50 shapes, render 84 models (42 with the
packed libm switch), full 45 models: no difference beyond the GPU tests' bounds.
"""
import dataclasses
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

import helpers as H  # noqa: E402
import lensed_b200 as L  # noqa: E402
import ptx_emu as E  # noqa: E402
import test_kernels_interpreted as K  # noqa: E402
from lensed_b200 import api  # noqa: E402


def random_objects(rng, w, h, host=True):
    objects, params = [], []

    def add(name, role):
        objects.append(name)
        params.extend(H._random_params(rng, name, w, h, role))
    if host and rng.random() < 0.3:
        add(str(rng.choice(H.SOURCES)), "host")
    for _ in range(int(rng.integers(1, 3))):
        add(str(rng.choice(H.LENSES)), "lens")
    for _ in range(int(rng.integers(1, 3))):
        add(str(rng.choice(H.SOURCES)), "source")
    if rng.random() < 0.7:
        add("sky", "sky")
    return objects, np.array(params, np.float32)


def quad_consts(cfg, block, psf=None):
    qq, ww = api.quad_rule(cfg.rule, 1, 1)
    c = {"lcu_quad": np.c_[qq, ww].astype(np.float32).view(np.uint32).ravel(), "lcu_objs_c": block}
    if psf is not None:
        c["lcu_psf"] = psf.view(np.uint32).ravel()
    return c


def conv(seed, extra):
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(1, 70)), int(rng.integers(1, 40))
    pw, ph = int(rng.integers(1, 8)), int(rng.integers(1, 8))
    psf = H.workloads.normalise_psf(rng.random((ph, pw)).astype(np.float32) + 0.1)
    cfg = dataclasses.replace(H.golden_config("sky"), name="conv", image=rng.random((h, w)).astype(np.float32),
                              weight=(rng.random((h, w)) + 0.5).astype(np.float32), rule="point", psf=psf)
    M, _, _ = K._program(cfg, L, extra)
    raw = rng.random((h, w)).astype(np.float32)*10
    mem = E.StrictMemory()
    K._put(mem, K.IMG, cfg.image)
    K._put(mem, K.WGT, cfg.weight)
    K._put(mem, K.RAW, raw)
    gpr = (w + 31)//32
    ng = h*gpr
    consts = {"lcu_psf": psf.view(np.uint32).ravel()}
    M.launch("lcu_convolve", ((w + 63)//64, (h + 31)//32, 1), 256, [K._convolve_args(K.RAW, K.MODEL, K.PART, h, ng, gpr, 5)], mem, consts)
    M.launch("lcu_convolve_small", ((w + 31)//32, (h + 7)//8, 1), 256, [K._convolve_args(K.RAW, K.MODEL1, K.PART1, h, ng, gpr, 5)], mem, consts)
    ref = np.asarray(cfg.oracle().convolve(raw), np.float32).view(np.uint32)
    ok = np.array_equal(K._get(mem, K.MODEL, (h, w)).view(np.uint32), ref) and np.array_equal(K._get(mem, K.MODEL1, (h, w)).view(np.uint32), ref) \
        and np.array_equal(K._get(mem, K.PART, (2*ng,)).view(np.uint32), K._get(mem, K.PART1, (2*ng,)).view(np.uint32))
    return ok, "image %dx%d psf %dx%d" % (w, h, pw, ph)


def render(seed, extra):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(6, 14)), int(rng.integers(8, 20))
    objects, params = random_objects(rng, w, h)
    cfg = H.Config(name="render%d" % seed, objects=objects, params=params, image=np.zeros((h, w), np.float32),
                   weight=np.ones((h, w), np.float32), rule=str(rng.choice(["point", "sub2"])), psf=None)
    M, text, words = K._program(cfg, L, extra)
    block = K._object_block(M, text, cfg, words)
    npix, ng = h*w, (h*w + 31)//32
    mem = E.StrictMemory()
    K._put(mem, K.IMG, cfg.image)
    K._put(mem, K.WGT, cfg.weight)
    consts = quad_consts(cfg, block)
    M.launch("lcu_render_pair", ((npix + 511)//512, 1), 256, [K._render_args(cfg, cfg.pcs, npix, K.RAW, K.PART, ng, 5)], mem, consts)
    M.launch("lcu_render_s1", ((npix + 255)//256, 1), 256, [K._render_args(cfg, cfg.pcs, npix, K.RAW1, K.PART1, ng, 5)], mem, consts)
    ref = np.asarray(cfg.oracle().render(cfg.params)[0])
    floor = H.rel_err(ref, np.asarray(cfg.oracle("f64").render(cfg.params)[0], np.float64)).max()
    rel = H.rel_err(K._get(mem, K.RAW, (h, w)), ref).max()
    same = np.array_equal(K._get(mem, K.RAW, (npix,)).view(np.uint32), K._get(mem, K.RAW1, (npix,)).view(np.uint32)) \
        and np.array_equal(K._get(mem, K.PART, (2*ng,)).view(np.uint32), K._get(mem, K.PART1, (2*ng,)).view(np.uint32))
    return same and rel <= max(1e-5, 1.5*floor), "%s %s %dx%d rel %.2e floor %.2e pair==one-ray %s" % (objects, cfg.rule, w, h, rel, floor, same)


def full(seed, extra):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(5, 14)), int(rng.integers(6, 40))
    objects, params = random_objects(rng, w, h, host=False)
    pw, ph = int(rng.integers(1, 7)), int(rng.integers(1, 7))
    psf = H.workloads.normalise_psf(rng.random((ph, pw)).astype(np.float32) + 0.1)
    cfg = H.Config(name="full%d" % seed, objects=objects, params=params, image=np.zeros((h, w), np.float32),
                   weight=np.ones((h, w), np.float32), rule=str(rng.choice(["point", "sub2"])), psf=psf)
    _, model, _ = cfg.oracle().loglike(cfg.params, want_maps=True)
    cfg.image, cfg.weight = H.workloads.observe(model, 500 + seed, gain=200.0, offset=0.5)
    M, text, words = K._program(cfg, L, extra)
    block = K._object_block(M, text, cfg, words)
    npix, gpr = h*w, (w + 31)//32
    ng = h*gpr
    pcs = list(cfg.pcs)                           # half-pixel shift of even PSFs, as lcu_model_create does (src/lensed.c:885-891)
    pcs[0] += 0.5*(pw % 2 == 0)
    pcs[1] += 0.5*(ph % 2 == 0)
    mem = E.StrictMemory()
    K._put(mem, K.IMG, cfg.image)
    K._put(mem, K.WGT, cfg.weight)
    consts = quad_consts(cfg, block, psf)
    M.launch("lcu_render_pair", ((npix + 511)//512, 1), 256, [K._render_args(cfg, pcs, npix, K.RAW, 0, (npix + 31)//32, 1)], mem, consts)
    M.launch("lcu_convolve", ((w + 63)//64, (h + 31)//32, 1), 256, [K._convolve_args(K.RAW, K.MODEL, K.PART, h, ng, gpr, 5)], mem, consts)
    M.launch("lcu_reduce", (1,), 256, [ng, K.PART, E.d2b(-0.5), K.LNEW], mem)
    ref_l, ref_model, _ = cfg.oracle().loglike(cfg.params, want_maps=True)
    rel = H.rel_err(K._get(mem, K.MODEL, (h, w)), ref_model).max()
    dl = abs(float(K._get(mem, K.LNEW, (1,), np.float64)[0]) - ref_l)/abs(ref_l)
    return rel <= 1e-5 and dl <= 4e-6, "%s %s %dx%d psf %dx%d model rel %.2e lnew rel %.2e" % (objects, cfg.rule, w, h, pw, ph, rel, dl)


if __name__ == "__main__":
    mode, first, last, extra = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4:]
    bad = 0
    for seed in range(first, last):
        t = time.time()
        ok, what = {"conv": conv, "render": render, "full": full}[mode](seed, extra)
        bad += not ok
        print(seed, "OK  " if ok else "FAIL", what, "%.1f s" % (time.time() - t), flush=True)
    print("%d of %d differ" % (bad, last - first))
    sys.exit(1 if bad else 0)
