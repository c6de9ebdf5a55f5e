"""Model configurations shared by the CPU and GPU tests: the reference's 16
known-answer test configurations (tests/*/*.ini of the reference, fixtures in
tests/golden/ref_goldens.npz), its three examples (C1-C3, tests/golden/examples.npz)
and the scaled synthetic C4 / C5 scenes of lensed_b200.workloads."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from lensed_b200 import workloads
from oracle import pyoracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# the reference's objects/*.cl, verbatim: the plugin directory of every test (golden/objects/README.md)
OBJECTS_DIR = os.path.join(GOLDEN, "objects")


@dataclass
class Config:
    name: str
    objects: list
    params: np.ndarray
    image: np.ndarray
    weight: np.ndarray
    rule: str = "g3k7"
    psf: Optional[np.ndarray] = None
    pcs: tuple = (1.0, 1.0, 1.0, 1.0)
    ipp: Optional[list] = None
    extra: dict = field(default_factory=dict)

    def oracle(self, variant="strict", lib=None):
        qq, ww = O.quad_rule(self.rule, self.pcs[2], self.pcs[3])
        flat = None
        if self.ipp is not None:
            flat = [int(v) for o in self.ipp for v in o]
        return O.Model(self.objects, self.image, self.weight, qq, ww, psf=self.psf, pcs=self.pcs, ipp=flat,
                       variant=variant, _lib=lib)

    def product(self, ctx, **kw):
        import lensed_b200 as L
        return L.Model(ctx, self.objects, self.image, self.weight, rule=self.rule, psf=self.psf, pcs=self.pcs,
                       ipp=self.ipp, **kw)


def _params_from_priors(objs, priors):
    """Fixed-value priors of the reference's test inis -> parameter vector in
    object order; parameters without a prior take the object's default."""
    vals = []
    for oid, name in objs:
        for p in O.object_info(name)["params"]:
            key = f"{oid}.{p['name']}"
            vals.append(float(priors[key]) if key in priors else p["defval"])
    return np.array(vals, np.float32)


def golden_names():
    with np.load(os.path.join(GOLDEN, "ref_goldens.npz")) as z:
        return sorted(json.loads(str(z["meta"])).keys())


def golden_config(name) -> Config:
    with np.load(os.path.join(GOLDEN, "ref_goldens.npz")) as z:
        meta = json.loads(str(z["meta"]))[name]
        img = z[name]
    objs = [tuple(o) for o in meta["objects"]]
    weight = np.full(img.shape, meta["weight"], np.float32)
    return Config(name=name, objects=[n for _, n in objs], params=_params_from_priors(objs, meta["priors"]),
                  image=img, weight=weight, rule=meta["rule"])


def _prior_mid(spec: str) -> float:
    toks = [t for t in spec.split() if t not in ("wrap", "image")]
    if toks[0] == "unif":
        return 0.5*(float(toks[1]) + float(toks[2]))
    if toks[0] == "norm":
        return float(toks[1])
    return float(toks[0])


def example_config(name: str, ipp: bool = True) -> Config:
    """C1 = test_sersic_bulge, C2 = full_mock_nopsf, C3 = full_mock_psf at the
    mid-points of their priors, image-plane priors as in the ini."""
    with np.load(os.path.join(GOLDEN, "examples.npz")) as z:
        meta = json.loads(str(z["meta"]))[name]
        img = z[name + "_image"]
        psf = z[name + "_psf"] if meta["psf"] else None
    objs = [tuple(o) for o in meta["objects"]]
    vals, flags = [], []
    for oid, oname in objs:
        f = []
        for p in O.object_info(oname)["params"]:
            spec = meta["priors"][f"{oid}.{p['name']}"]
            vals.append(_prior_mid(spec))
            f.append(int(ipp and "image" in spec.split()))
        flags.append(f)
    weight = workloads.make_weight(img, meta["gain"], meta["offset"])
    return Config(name=name + ("" if ipp else "-noipp"), objects=[n for _, n in objs], params=np.array(vals, np.float32),
                  image=img, weight=weight, rule=meta["rule"],
                  psf=workloads.normalise_psf(psf) if psf is not None else None,
                  pcs=tuple(meta["pcs"]), ipp=flags if ipp else None)


def synthetic_config(which: str, size: int, psf: bool = True, psf_shape=None, mask: float = 0.0, rule=None) -> Config:
    """Scaled C4 / C5 scene.  The observed image is the strict-float32 oracle
    model at the truth plus noise (seeded), as SURVEY.md section 8d specifies."""
    w = workloads.c4(size) if which == "c4" else workloads.c5(size)
    p = w["psf"] if psf else None
    if psf and psf_shape is not None:
        p = workloads.gaussian_psf(psf_shape[0], psf_shape[1], 2.0)
    cfg = Config(name=f"{w['name']}{'' if psf else '-nopsf'}", objects=w["objects"], params=w["truth"],
                 image=np.zeros((size, size), np.float32), weight=np.ones((size, size), np.float32),
                 rule=rule or w["rule"], psf=p, extra=dict(workload=w))
    _, model, _ = cfg.oracle().loglike(cfg.params, want_maps=True)
    cfg.image, cfg.weight = workloads.observe(model, w["noise_seed"], mask_fraction=mask)
    return cfg


def c5_fixture_observation(model512):
    """(image, weight) of the full-size C5 fixture (tools/make_c5_fixture.py):
    the committed 512^2 oracle model block-replicated 8 x 8 and divided by 64
    (exact in float32) plus seeded noise; no renderer involved, so every
    machine rebuilds the same bits (the fixture carries a SHA-256 to prove it)."""
    m = np.kron(np.asarray(model512, np.float32), np.ones((8, 8), np.float32))*np.float32(1.0/64)
    return workloads.observe(m, 12346)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b)/np.maximum(np.abs(b), 1e-30)


LENSES = ["sis", "sis_plus_shear", "sie", "sie_plus_shear", "nsis", "nsie", "epl", "epl_plus_shear", "point_mass"]
SOURCES = ["sersic", "devauc", "exponential", "gauss", "sersic-old"]
RULES = ["point", "sub2", "sub4", "gm75", "g3k7", "g5k11"]


def _random_params(rng, name, width, height, role):
    """Plausible parameters for one object from its metadata (kernel/object.cl
    types): positions near the image centre, radii of a few pixels, and the
    bounds the object declares for plain parameters."""
    cx, cy = 0.5*(width + 1), 0.5*(height + 1)
    vals = []
    for p in O.object_info(name)["params"]:
        t, pname = p["type"], p["name"]
        if t == 1:
            vals.append(cx + rng.uniform(-4, 4) + (rng.uniform(-6, 6) if role == "source" else 0))
        elif t == 2:
            vals.append(cy + rng.uniform(-4, 4) + (rng.uniform(-6, 6) if role == "source" else 0))
        elif t == 3:
            vals.append(rng.uniform(0.5, 2.0) if pname == "rc" else
                        rng.uniform(0.15, 0.3)*min(width, height) if role == "lens" else rng.uniform(1.5, 6.0))
        elif t == 4:
            vals.append(rng.uniform(-5.0, -2.0))
        elif t == 5:
            vals.append(rng.uniform(0.4, 0.95))
        elif t == 6:
            vals.append(rng.uniform(0.0, 180.0))
        elif pname in ("g1", "g2"):
            vals.append(rng.uniform(-0.06, 0.06))
        elif pname == "t":
            vals.append(rng.uniform(0.7, 1.4))
        elif pname == "n":
            vals.append(rng.uniform(0.6, 5.0))
        elif pname == "bg":
            vals.append(rng.uniform(0.01, 0.1))
        elif pname in ("dx", "dy"):
            vals.append(rng.uniform(-3e-4, 3e-4))
        else:
            lo, hi = p["bounds"]
            vals.append(rng.uniform(lo if np.isfinite(lo) else 0.1, hi if np.isfinite(hi) else 2.0))
    return vals


def random_config(seed: int) -> Config:
    """A random model: optional unlensed host galaxy, one or two lenses in one
    plane, one to three lensed sources, optional sky; random image shape,
    quadrature rule and PSF (none / odd / even / ragged); observed image = the
    strict-float32 oracle model plus seeded noise."""
    rng = np.random.default_rng(1000 + seed)
    height, width = int(rng.integers(36, 80)), int(rng.integers(36, 80))
    objects, params = [], []

    def add(name, role):
        objects.append(name)
        params.extend(_random_params(rng, name, width, height, role))
    if rng.random() < 0.3:
        add(str(rng.choice(SOURCES)), "host")
    for _ in range(int(rng.integers(1, 3))):
        add(str(rng.choice(LENSES)), "lens")
    for _ in range(int(rng.integers(1, 4))):
        add(str(rng.choice(SOURCES)), "source")
    if rng.random() < 0.7:
        add("sky", "sky")
    kind = rng.integers(0, 4)
    psf = None
    if kind == 1:
        psf = workloads.gaussian_psf(2*int(rng.integers(1, 5)) + 1, 2*int(rng.integers(1, 5)) + 1, rng.uniform(0.8, 2.0))
    elif kind == 2:
        psf = workloads.gaussian_psf(2*int(rng.integers(1, 4)), 2*int(rng.integers(1, 4)), rng.uniform(0.8, 2.0))
    elif kind == 3:
        psf = workloads.normalise_psf(rng.random((int(rng.integers(1, 9)), int(rng.integers(1, 9)))) + 0.05)
    cfg = Config(name=f"random-{seed}", objects=objects, params=np.array(params, np.float32),
                 image=np.zeros((height, width), np.float32), weight=np.ones((height, width), np.float32),
                 rule=str(rng.choice(RULES)), psf=psf)
    _, model, _ = cfg.oracle().loglike(cfg.params, want_maps=True)
    cfg.image, cfg.weight = workloads.observe(model, 500 + seed, gain=200.0, offset=0.5)
    return cfg
