"""Synthetic workloads of the named benchmark configurations.

C4: 1024^2, sie_plus_shear + 2 sersic + sky, 25x25 Gaussian PSF, rule g7k15.
C5: 4096^2, epl_plus_shear + 3 sersic + sky, 25x25 PSF, rule g3k7.
Both scale with ``size`` (positions and radii proportional) so that tests can
run the same scene at sizes a CPU oracle finishes in seconds.

Host-side numpy only.  The observed image is *input data*: any model image
(``truth_model``) plus Gaussian noise of variance (model + offset)/gain; the
weight map is built as the reference's make_weight() does
(src/data.c:314-330).
"""
from __future__ import annotations

import numpy as np

# gain / offset of examples/full_mock_psf.ini:24-25
GAIN = 1800.0
OFFSET = 2.9633


def gaussian_psf(width: int = 25, height: int = 25, sigma: float = 3.0) -> np.ndarray:
    """Circular Gaussian, normalised in double then narrowed to float, as
    read_psf() does (src/data.c:354-370)."""
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    cx, cy = 0.5*(width - 1), 0.5*(height - 1)
    p = np.exp(-0.5*((x - cx)**2 + (y - cy)**2)/sigma**2).astype(np.float32)
    return normalise_psf(p)


def normalise_psf(psf) -> np.ndarray:
    psf = np.asarray(psf, dtype=np.float32)
    norm = 0.0
    for v in psf.ravel().tolist():      # serial double sum in file order, like the reference
        norm += v
    return (psf.astype(np.float64)/norm).astype(np.float32)


def make_weight(image, gain=GAIN, offset=OFFSET) -> np.ndarray:
    """weight = gain/(image + offset) evaluated in double (src/data.c:326),
    negative weights clipped to zero."""
    image = np.asarray(image, dtype=np.float32)
    w = (np.float32(gain).astype(np.float64)/(image.astype(np.float64) + offset)).astype(np.float32)
    return np.where(w >= 0, w, 0).astype(np.float32)


def c4(size: int = 1024) -> dict:
    s = size/1024.0
    objects = ["sie_plus_shear", "sersic", "sersic", "sky"]
    truth = [
        512.5*s, 512.5*s, 200*s, 0.75, 45.0, 0.03, -0.02,           # lens: x y r q pa g1 g2
        497*s, 525*s, 40*s, -9.0, 3.18, 0.89, 30.0,                  # src1: x y r mag n q pa
        540*s, 500*s, 15*s, -7.5, 1.5, 0.7, 110.0,                   # src2
        0.05, 0.0, 0.0,                                              # sky: bg dx dy
    ]
    angles = [4, 13, 20]
    return dict(name=f"C4-{size}", objects=objects, truth=np.array(truth, np.float32), angle_idx=angles,
                width=size, height=size, rule="g7k15", psf=gaussian_psf(), noise_seed=12345,
                flops_per_ray=80, transc_per_ray=10)


def c5(size: int = 4096) -> dict:
    s = size/4096.0
    objects = ["epl_plus_shear", "sersic", "sersic", "sersic", "sky"]
    truth = [
        2048.5*s, 2048.5*s, 800*s, 1.1, 0.75, 45.0, 0.03, -0.02,     # lens: x y r t q pa g1 g2
        1988*s, 2100*s, 160*s, -9.0, 3.18, 0.89, 30.0,               # src1 (C4 x4)
        2160*s, 2000*s, 60*s, -7.5, 1.5, 0.7, 110.0,                 # src2 (C4 x4)
        2100*s, 2150*s, 90*s, -10.0, 2.2, 0.8, 70.0,                 # src3
        0.05, 0.0, 0.0,
    ]
    angles = [5, 14, 21, 28]
    return dict(name=f"C5-{size}", objects=objects, truth=np.array(truth, np.float32), angle_idx=angles,
                width=size, height=size, rule="g3k7", psf=gaussian_psf(), noise_seed=12346,
                flops_per_ray=237, transc_per_ray=24)


def observe(truth_model, seed: int, gain=GAIN, offset=OFFSET, mask_fraction: float = 0.0, mask_seed: int = 7):
    """(image, weight): noisy observation of a model image and its weight map;
    optionally a random fraction of pixels masked (weight 0, src/lensed.c:479-482)."""
    m = np.asarray(truth_model, dtype=np.float64)
    rng = np.random.default_rng(seed)
    sigma = np.sqrt(np.maximum(m + offset, 0.0)/gain)
    image = (m + sigma*rng.standard_normal(m.shape)).astype(np.float32)
    weight = make_weight(image, gain, offset)
    if mask_fraction > 0:
        mrng = np.random.default_rng(mask_seed)
        weight = np.where(mrng.random(m.shape) < mask_fraction, np.float32(0), weight).astype(np.float32)
    return image, weight


def param_batch(w: dict, nbatch: int, seed: int = 2024) -> np.ndarray:
    """B points around the truth: non-angle parameters x (1 + 0.01 U(-1,1)),
    angles +- 1 degree."""
    rng = np.random.default_rng(seed)
    t = w["truth"].astype(np.float64)
    u = rng.uniform(-1, 1, size=(nbatch, t.size))
    p = t[None, :]*(1 + 0.01*u)
    for i in w["angle_idx"]:
        p[:, i] = t[i] + u[:, i]
    return p.astype(np.float32)


def work_per_eval(w: dict, nq: int) -> dict:
    """Algorithmic work of one evaluation (SURVEY.md section 8d / BASELINE.md section 5)."""
    npix = w["width"]*w["height"]
    rays = npix*nq
    psf = w.get("psf")
    conv = 2.0*npix*psf.size if psf is not None else 0.0
    flops = rays*w["flops_per_ray"] + conv + 3.0*npix
    return dict(rays=rays, flops=flops, render_flops=rays*w["flops_per_ray"], convolve_flops=conv,
                transc=rays*w["transc_per_ray"], hbm_bytes=(16 if psf is not None else 8)*npix)
