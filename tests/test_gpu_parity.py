"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances are the north-star's: model images per-pixel relative error
<= 1e-5, log-likelihoods <= 1e-6 relative (BASELINE.json).  Integer /
structural results (object blocks, summation order, batch invariance) are
checked bit-exactly.
"""
import numpy as np
import pytest

import helpers as H
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu

PIXEL_TOL = 1e-5
LOGLIKE_TOL = 1e-6

# build modes under test: the default (one IEEE operation per source operation,
# accurate libdevice functions) and the relaxed build bench.py runs,
# LCU_FAST_INTRINSICS | LCU_FAST_ATANH: exp / log of source and foreground
# objects on the hardware exp2 / log2 units (compensated arguments), atanh of
# lens objects on the hardware log2 -- "fast-math only where the reference's
# tolerance allows": both have to meet the same bounds
FAST = 4 | 32
MATH_MODES = [pytest.param(0, id="strict"), pytest.param(FAST, id="fast")]


# ---------------------------------------------------------------------------
# Parity criteria.  Two are evaluated and RECORDED for every case (the record
# goes to gpurun_out/parity_report.json; a copy per round is committed under
# profiles/):
#
#   flat   the north-star's letter: per-pixel |gpu - o32|/|o32| <= 1e-5,
#          |lnew_gpu - lnew_o32|/|lnew_o32| <= 1e-6, o32 = the strict-float32 oracle;
#   floor  |gpu - f64| <= k |o32 - f64|: the CUDA path is no further from the
#          exact (float64) answer than the reference's own float32 arithmetic is.
#          Two independent float32 evaluations of an ill-conditioned scene
#          cannot agree better than either agrees with the exact result.  For a
#          scalar (lnew) one float32 realisation may land on the float64 value
#          by luck, so the floor there is the largest distance from float64 among
#          the float32 realisations at hand: the strict oracle, its -ffast-math
#          build (what -cl-fast-relaxed-math licenses the reference's compiler to
#          do, src/lensed.c:744-748) and, where oracle/_ref holds the configuration,
#          the reference's own kernels executed the way an OpenCL CPU runtime does
#          (work-items in SIMD lanes, vector math functions: liblensed_ref_simd.so).
#
# The 99.9th percentile of the per-pixel error has to meet the flat bound
# always.  The maximum and lnew have to meet the flat bound wherever a test
# says flat=True (the reference's 16 configurations, C1-C3, the C4 scenes, every
# truth point from 256^2 up) and flat-or-floor elsewhere (random ill-conditioned
# scenes, 1 %-off points with chi^2/dof ~ 10^2, pre-PSF C5).
# ---------------------------------------------------------------------------
# k = 2: the CUDA path's distance from float64 and the floor are both draws from
# rounding noise of the same scale; over the 141 log-likelihoods and 238 images of
# this suite the ratio of the two has median 0.70 / 0.91, 90th percentile 1.26 /
# 1.13 and maximum 1.81 / 1.67 (profiles/r02_parity_report.json): the CUDA path
# is typically CLOSER to the exact result than the reference's own float32
# realisations are, and never twice as far.  (The bulk of the image error is
# pinned separately by the flat 1e-5 bound on the 99.9th percentile.)
FLOOR_K = 2.0
FLOOR_K_MAX = 2.0
import os as _os
FULL_RECORD = bool(_os.environ.get("LCU_PARITY_FULL_RECORD"))
_REPORT = []


@pytest.fixture(scope="module", autouse=True)
def _parity_report():
    yield
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "parity_report.json"), "w") as f:
            json.dump(_REPORT, f, indent=0)
    except OSError:
        pass


def _realisations(cfg):
    """Other float32 realisations of the reference's arithmetic for this
    configuration, for the floor: the oracle's -ffast-math build and, where
    oracle/_ref holds the configuration, the reference's own kernels built with
    the flags the reference itself builds them with (-cl-fast-relaxed-math ~
    -ffast-math: ref_fast) and with their work-items packed into SIMD lanes and
    libmvec's vector math functions (what an OpenCL CPU runtime executes:
    ref_simd, oracle/ref_shim_simd.h)."""
    out = {"fast": cfg.oracle(variant="fast")}
    for v in ("ref_fast", "ref_simd"):
        if O.available(v):
            try:
                out[v] = cfg.oracle(variant=v, lib=O.lib(v))
            except Exception:
                pass            # oracle/_ref was not built with this object list
    return out


def _check_lnew(cfg, params, got, flat=False, tag="", om=None):
    """One log-likelihood against the oracle: flat 1e-6, or the floor criterion."""
    om = om or cfg.oracle()
    l32 = om.loglike(params)
    l64 = cfg.oracle(variant="f64").loglike(params)
    flat_ok = bool(abs(got - l32) <= LOGLIKE_TOL*abs(l32))
    # the other float32 realisations are evaluated where the verdict needs them, or everywhere for the
    # committed record (LCU_PARITY_FULL_RECORD=1: profiles/r02_parity_report.json)
    others = {k: v.loglike(params) for k, v in _realisations(cfg).items()} if (FULL_RECORD or not flat_ok) else {}
    floor = max([abs(l32 - l64)] + [abs(v - l64) for v in others.values()])
    n = cfg.image.size
    rec = dict(case=cfg.name, what="lnew", tag=tag, gpu=got, o32=l32, f64=l64, rel_vs_o32=abs(got - l32)/abs(l32),
               rel_vs_f64=abs(got - l64)/abs(l64), floor_rel=floor/abs(l64), chi2_per_pixel=-2*l64/n,
               others_rel_vs_f64={k: abs(v - l64)/abs(l64) for k, v in others.items()},
               flat_ok=bool(abs(got - l32) <= LOGLIKE_TOL*abs(l32)), floor_ok=bool(abs(got - l64) <= FLOOR_K*floor),
               flat_required=flat)
    _REPORT.append(rec)
    if flat:
        assert rec["flat_ok"], f"{cfg.name} {tag}: lnew {got} vs {l32} (rel {rec['rel_vs_o32']:.3e})"
    else:
        assert rec["flat_ok"] or rec["floor_ok"], \
            f"{cfg.name} {tag}: lnew {got} vs o32 {l32} (rel {rec['rel_vs_o32']:.3e}), vs f64 {rec['rel_vs_f64']:.3e}, floor {rec['floor_rel']:.3e}"
    return l32


def _check_images(out, cfg, om, flat=False, tag=""):
    """Per-pixel relative error of the raw (pre-PSF) and model images against
    the strict-float32 oracle: p99.9 <= 1e-5 always; the maximum <= 1e-5
    (flat=True), or else no further from the float64 twin than 1.5 x the largest
    distance from it among the float32 realisations of the reference's arithmetic
    (2 x for the single worst pixel)."""
    value, error = om.render(cfg.params)
    lnew, model, chi = om.loglike(cfg.params, want_maps=True)
    o64 = cfg.oracle(variant="f64")
    v64, _ = o64.render(cfg.params)
    _, m64, _ = o64.loglike(cfg.params, want_maps=True)
    others = {}
    if FULL_RECORD or any(H.rel_err(out[key], ref).max() > PIXEL_TOL for key, ref in (("raw", value), ("model", model))):
        for k, o in _realisations(cfg).items():
            others[k] = {"raw": o.render(cfg.params)[0], "model": o.loglike(cfg.params, want_maps=True)[1]}
    stats = {}
    for key, ref, ref64 in (("raw", value, v64), ("model", model, m64)):
        r = H.rel_err(out[key], ref)
        fl = H.rel_err(ref, ref64)
        r64 = H.rel_err(out[key], ref64)
        floors = {"strict": float(fl.max())}
        floors.update({k: float(H.rel_err(v[key], ref64).max()) for k, v in others.items()})
        floor = max(floors.values())
        need = bool(flat is True or (flat and key in flat))
        rec = dict(case=cfg.name, what=key, tag=tag, max=float(r.max()), p999=float(np.quantile(r, 0.999)),
                   floor_max=floor, floors=floors, floor_p999=float(np.quantile(fl, 0.999)), gpu_vs_f64_max=float(r64.max()),
                   flat_ok=bool(r.max() <= PIXEL_TOL), floor_ok=bool(r64.max() <= FLOOR_K_MAX*floor), flat_required=need)
        _REPORT.append(rec)
        stats[key] = (r.max(), floor)
        assert rec["p999"] <= PIXEL_TOL, f"{cfg.name}: {key} image p99.9 rel err {rec['p999']:.3e}"
        if need:
            assert rec["flat_ok"], f"{cfg.name}: {key} image max rel err {r.max():.3e}"
        else:
            assert rec["flat_ok"] or rec["floor_ok"], \
                f"{cfg.name}: {key} image max rel err {r.max():.3e} vs o32, {r64.max():.3e} vs f64 (f32 noise floors {floors})"
    # the quadrature error estimate is a sum with alternating-sign weights:
    # compare it on the scale of the value it estimates the error of
    scale = np.maximum(np.abs(value), 1e-30)
    e = (np.abs(out["error"].astype(np.float64) - error)/scale).max()
    assert e <= 10*PIXEL_TOL, f"{cfg.name}: error image differs by {e:.3e} of the value"
    return lnew, stats


def _check_block(m, om, cfg, atol=1e-30):
    """Object data block written by set_params: same layout word for word;
    values agree to ~2 ulp (the setters evaluate their math built-ins in double
    on the device and round once; the host libm is not correctly rounded for
    every function, e.g. tgammaf).  `atol`: for words that are a difference of
    large terms (a Sersic log0 that happens to come out near zero)."""
    blk = m.set_params(cfg.params).view(np.float32)
    ref = om.set_params(cfg.params).astype(np.float32)
    assert blk.size == ref.size, f"{cfg.name}: object block size {blk.size} vs {ref.size}"
    assert np.array_equal(blk == 0, ref == 0), f"{cfg.name}: object block layout differs"
    bad = ~np.isclose(blk, ref, rtol=5e-7, atol=atol)
    assert not bad.any(), f"{cfg.name}: object block words {np.nonzero(bad)[0]}: {blk[bad]} vs {ref[bad]}"


@pytest.mark.parametrize("flags", MATH_MODES)
@pytest.mark.parametrize("name", H.golden_names())
def test_reference_known_answer_configs(gpu_ctx, name, flags):
    """The reference's own 16 test configurations (tests/Makefile:1-17)."""
    cfg = H.golden_config(name)
    om = cfg.oracle()
    m = cfg.product(gpu_ctx, flags=flags)
    out = m.render(cfg.params)
    lnew, _ = _check_images(out, cfg, om, flat=True, tag=f"flags={flags}")
    # loose anchor against the reference's golden image: chi^2/dof << 1
    got = m.loglike(cfg.params)
    n = cfg.image.size
    assert -2*got/n < 0.01, f"{name}: chi2/dof {-2*got/n:.3e} against the reference's golden image"
    # chi^2 ~ 0 here (model == image up to the noise floor), so a relative
    # comparison is ill-conditioned (SURVEY.md 'hard parts'): compare on the
    # scale of the number of degrees of freedom instead
    assert abs(got - lnew) <= LOGLIKE_TOL*n, f"{name}: lnew {got} vs {lnew}"
    _check_block(m, om, cfg)


@pytest.mark.parametrize("flags", MATH_MODES)
@pytest.mark.parametrize("name,ipp", [("test_sersic_bulge", True), ("full_mock_nopsf", True),
                                      ("full_mock_psf", True), ("full_mock_psf", False)])
def test_examples(gpu_ctx, name, ipp, flags):
    """C1-C3: the reference's examples, image-plane priors included."""
    cfg = H.example_config(name, ipp)
    om = cfg.oracle()
    m = cfg.product(gpu_ctx, flags=flags)
    out = m.render(cfg.params)
    _check_images(out, cfg, om, flat=True, tag=f"flags={flags}")
    _check_lnew(cfg, cfg.params, m.loglike(cfg.params), flat=True, tag=f"flags={flags}", om=om)
    _check_block(m, om, cfg)
    # per-pixel chi^2 map (PVL layer of the dumper)
    _, _, chi = om.loglike(cfg.params, want_maps=True)
    assert np.allclose(out["chi"], chi, rtol=1e-4, atol=1e-6*chi.max())


@pytest.mark.parametrize("flags", MATH_MODES)
@pytest.mark.parametrize("which,size,psf", [("c4", 128, True), ("c4", 128, False), ("c5", 128, True), ("c5", 512, True),
                                            ("c4", 256, True)])
def test_synthetic_scenes(gpu_ctx, which, size, psf, flags):
    """Scaled C4 / C5 scenes on noisy images (chi^2 ~ N_pix: well-conditioned
    lnew).  Images: flat 1e-5 (the pre-PSF image of the EPL scene, whose oracle
    is itself 1e-5 from its float64 twin, by the floor criterion).  lnew at the
    truth: flat 1e-6 from 256^2 pixels up (rounding noise eps moves lnew by
    ~2 (S/N) eps / sqrt(N_pix); the named configurations are 1024^2 and 4096^2),
    flat-or-floor at 128^2 and at the 1 %-off points of the batch (chi^2/dof ~
    10^2: per-pixel differences add up coherently with the residuals)."""
    cfg = H.synthetic_config(which, size, psf=psf)
    om = cfg.oracle()
    m = cfg.product(gpu_ctx, flags=flags)
    out = m.render(cfg.params)
    _check_images(out, cfg, om, flat=True if which == "c4" else {"model"}, tag=f"flags={flags}")
    _check_lnew(cfg, cfg.params, m.loglike(cfg.params), flat=size >= 256, tag=f"truth flags={flags}", om=om)
    # batch of perturbed points (1 % off the truth: chi^2/dof ~ 10^2)
    P = H.workloads.param_batch(cfg.extra["workload"], 5)
    got = m.loglike_batch(P)
    for i, p in enumerate(P):
        _check_lnew(cfg, p, got[i], flat=False, tag=f"batch[{i}] flags={flags}", om=om)


def test_masked_pixels(gpu_ctx):
    cfg = H.synthetic_config("c4", 128, mask=0.1)
    assert (cfg.weight == 0).mean() > 0.05
    om, m = cfg.oracle(), cfg.product(gpu_ctx)
    _check_lnew(cfg, cfg.params, m.loglike(cfg.params), flat=True, tag="masked", om=om)


@pytest.mark.parametrize("psf_shape", [(9, 9), (8, 8), (6, 11), (25, 25), (1, 1)])
@pytest.mark.parametrize("small", ["0", "1"])
def test_convolution_shapes(gpu_ctx, psf_shape, small, monkeypatch):
    """Odd, even (half-pixel shift of src/lensed.c:885-891) and ragged PSFs; the
    image is smaller than a tile in one direction so every edge clamp is hit.
    Both convolution kernels: 64 x 32 tiles (small = 0) and 32 x 8 tiles."""
    monkeypatch.setenv("LCU_CONV_SMALL", small)
    rng = np.random.default_rng(5)
    psf = H.workloads.normalise_psf(rng.random((psf_shape[1], psf_shape[0])) + 0.01)
    w = H.workloads.c4(64)
    img = np.zeros((40, 64), np.float32)
    cfg = H.Config("conv", w["objects"], w["truth"], img, np.ones_like(img), rule="sub2", psf=psf)
    om, m = cfg.oracle(), cfg.product(gpu_ctx)
    out = m.render(cfg.params)
    _check_images(out, cfg, om)
    # convolution alone is bit-exact given the same input (same summation order)
    conv = om.convolve(out["raw"])
    assert np.array_equal(out["model"].view(np.uint32), conv.view(np.uint32))


@pytest.mark.parametrize("shape,psf_shape", [((100, 100), (9, 9)), ((45, 70), (13, 7)), ((33, 31), (4, 6)), ((120, 120), (21, 21))])
def test_convolution_kernels_same_bits(gpu_ctx, monkeypatch, shape, psf_shape):
    """The small-launch convolution (32 x 8 tiles, one pixel per thread) and the
    register-tiled one give the same model image, chi^2 map and log-likelihood
    to the last bit, for single points (graph path), batches and row strips."""
    rng = np.random.default_rng(11)
    psf = H.workloads.normalise_psf(rng.random((psf_shape[1], psf_shape[0])) + 0.01)
    w = H.workloads.c4(max(shape))
    img = rng.random(shape).astype(np.float32)
    wht = (rng.random(shape) > 0.1).astype(np.float32)*rng.random(shape).astype(np.float32)
    cfg = H.Config("conv2", w["objects"], w["truth"], img, wht, rule="sub2", psf=psf)
    P = np.stack([cfg.params, cfg.params*np.float32(1.001), cfg.params*np.float32(0.999)])
    res = {}
    for small in ("0", "1"):
        monkeypatch.setenv("LCU_CONV_SMALL", small)
        m = cfg.product(gpu_ctx)
        out = m.render(cfg.params)
        one = [m.loglike(p) for p in P]
        many = m.loglike_batch(P)
        m.set_rows(5, shape[0] - 7)
        strip = m.loglike_batch(P)
        m.close()
        res[small] = (out["model"], out["chi"], np.array(one), many, strip)
    for a, b in zip(res["0"], res["1"]):
        assert np.array_equal(a.view(np.uint64 if a.dtype == np.float64 else np.uint32),
                              b.view(np.uint64 if b.dtype == np.float64 else np.uint32))
    assert np.array_equal(res["0"][2], res["0"][3])


def test_batch_and_split_invariance(gpu_ctx, monkeypatch):
    """A point's log-likelihood does not depend on the batch it is evaluated
    in, on the batch chunking, or on how many warps share a pixel group."""
    cfg = H.synthetic_config("c4", 96)
    P = H.workloads.param_batch(cfg.extra["workload"], 7)
    m = cfg.product(gpu_ctx)
    base = m.loglike_batch(P)
    single = np.array([m.loglike(p) for p in P])
    assert np.array_equal(base, single)
    m2 = cfg.product(gpu_ctx, max_batch=3)
    assert m2.max_batch == 3
    assert np.array_equal(m2.loglike_batch(P), base)
    img0 = m.render(P[0])
    for split in ("1", "2", "4", "8"):
        monkeypatch.setenv("LCU_SPLIT", split)
        assert np.array_equal(m.loglike_batch(P), base), f"split {split}"
        img = m.render(P[0])
        assert np.array_equal(img["raw"], img0["raw"]) and np.array_equal(img["error"], img0["error"])
        # the split kernels shoot two quadrature points per pass for pairable models (lcu_render_q_s*): one per pass instead
        monkeypatch.setenv("LCU_NO_SPLIT_PAIR", "1")
        assert np.array_equal(m.loglike_batch(P), base), f"split {split}, one point per pass"
        monkeypatch.delenv("LCU_NO_SPLIT_PAIR")
    monkeypatch.delenv("LCU_SPLIT")
    # the same on the reference's examples (rule g3k7: 49 points = chunks of 32 + 17, an odd tail) and a one-point rule
    for cfg2 in (H.example_config("test_sersic_bulge"), H.example_config("full_mock_nopsf"),
                 H.Config("point-rule", cfg.objects, cfg.params, cfg.image, cfg.weight, rule="point", psf=cfg.psf),
                 H.Config("gm75-rule", cfg.objects, cfg.params, cfg.image, cfg.weight, rule="gm75")):
        m3 = cfg2.product(gpu_ctx)
        P3 = np.stack([cfg2.params, cfg2.params*np.float32(1.001)])
        a = m3.loglike_batch(P3)
        one = np.array([m3.loglike(p) for p in P3])
        monkeypatch.setenv("LCU_NO_SPLIT_PAIR", "1")
        m4 = cfg2.product(gpu_ctx)
        assert np.array_equal(m4.loglike_batch(P3), a) and np.array_equal(np.array([m4.loglike(p) for p in P3]), one), cfg2.name
        monkeypatch.delenv("LCU_NO_SPLIT_PAIR")
        assert np.array_equal(a, one), cfg2.name
        m3.close(); m4.close()


def test_row_strips_add_up(gpu_ctx):
    """Multi-GPU row-strip mode: strip log-likelihoods sum to the full one."""
    for psf in (True, False):
        cfg = H.synthetic_config("c4", 96, psf=psf)
        m = cfg.product(gpu_ctx)
        full = m.loglike(cfg.params)
        parts = []
        for r0, r1 in ((0, 17), (17, 64), (64, 96)):
            m.set_rows(r0, r1)
            parts.append(m.loglike(cfg.params))
        assert abs(sum(parts) - full) <= 1e-12*abs(full)
        m.set_rows(0, 96)
        assert m.loglike(cfg.params) == full


def test_shared_memory_object_blocks(gpu_ctx, monkeypatch):
    import lensed_b200 as L
    cfg = H.synthetic_config("c4", 96)
    a = cfg.product(gpu_ctx)
    b = cfg.product(gpu_ctx, flags=L.LCU_OBJ_SHARED)
    P = H.workloads.param_batch(cfg.extra["workload"], 3)
    base = a.loglike_batch(P)
    assert np.array_equal(base, b.loglike_batch(P))
    # the large-image kernels (two rays and one ray per thread) with the block in shared memory
    monkeypatch.setenv("LCU_SPLIT", "1")
    assert np.array_equal(base, b.loglike_batch(P))
    c = cfg.product(gpu_ctx, flags=L.LCU_OBJ_SHARED | L.LCU_NO_PAIR)
    assert np.array_equal(base, c.loglike_batch(P))


def test_device_resident_batch(gpu_ctx):
    torch = pytest.importorskip("torch")
    cfg = H.synthetic_config("c4", 96)
    m = cfg.product(gpu_ctx)
    P = H.workloads.param_batch(cfg.extra["workload"], 9)
    ref = m.loglike_batch(P)
    dp = torch.from_numpy(P).cuda()
    dl = torch.zeros(P.shape[0], dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream()
    m.loglike_batch_device(P.shape[0], dp.data_ptr(), dl.data_ptr(), s.cuda_stream)
    s.synchronize()
    assert np.array_equal(dl.cpu().numpy(), ref)


def test_full_size_c4_properties(gpu_ctx):
    """1024^2 / g7k15 / 25x25 PSF: properties that do not need the oracle at
    full size -- noise-free image gives lnew == 0 exactly, strips add up,
    batch invariance -- plus a banded oracle comparison of 64 rows."""
    w = H.workloads.c4(1024)
    img = np.zeros((1024, 1024), np.float32)
    cfg = H.Config("C4", w["objects"], w["truth"], img, np.ones_like(img), rule=w["rule"], psf=w["psf"])
    m = cfg.product(gpu_ctx)
    out = m.render(cfg.params, error=False, chi=False)
    m2 = H.Config("C4", w["objects"], w["truth"], out["model"], np.ones_like(img), rule=w["rule"], psf=w["psf"]).product(gpu_ctx)
    assert m2.loglike(cfg.params) == 0.0
    image, weight = H.workloads.observe(out["model"], w["noise_seed"])
    m3 = H.Config("C4", w["objects"], w["truth"], image, weight, rule=w["rule"], psf=w["psf"]).product(gpu_ctx)
    full = m3.loglike(cfg.params)
    assert 0.8 < -2*full/image.size < 1.2          # chi^2/dof ~ 1 at the truth
    m3.set_rows(0, 500)
    a = m3.loglike(cfg.params)
    m3.set_rows(500, 1024)
    b = m3.loglike(cfg.params)
    assert abs(a + b - full) <= 1e-12*abs(full)
    # banded oracle check: rows 480..544 of the raw image through a 64-row
    # crop with shifted pixel origin (same float coordinates -> same rays)
    band = H.Config("C4-band", w["objects"], w["truth"], np.zeros((64, 1024), np.float32), np.ones((64, 1024), np.float32),
                    rule=w["rule"], pcs=(1.0, 481.0, 1.0, 1.0))
    value, _ = band.oracle().render(cfg.params)
    r = H.rel_err(out["raw"][480:544], value)
    assert r.max() <= PIXEL_TOL, f"band max rel err {r.max():.3e}"


def test_non_finite_deflection_guard(gpu_ctx):
    """A ray that hits a singular lens centre exactly has a NaN deflection
    (0/0); compute() then throws it to (1e10, 1e10) (src/kernel.c:86-91) and it
    contributes the source's (vanishing) brightness there plus the sky."""
    img = np.zeros((9, 9), np.float32)
    for lens, lp in (("sis", [5.0, 5.0, 2.0]), ("point_mass", [5.0, 5.0, 2.0]), ("sis_plus_shear", [5.0, 5.0, 2.0, 0.1, 0.0])):
        params = np.array(lp + [5.0, 5.0, 1.5, -2.0, 1.0, 0.9, 10.0] + [0.25, 0.0, 0.0], np.float32)
        cfg = H.Config("guard-" + lens, [lens, "sersic", "sky"], params, img, np.ones_like(img), rule="point")
        om, m = cfg.oracle(), cfg.product(gpu_ctx)
        value, _ = om.render(params)
        out = m.render(params)
        assert np.isfinite(out["raw"]).all()
        assert out["raw"][4, 4] == np.float32(0.25) == value[4, 4]        # centre pixel: sky only
        assert H.rel_err(out["raw"], value).max() <= PIXEL_TOL


def test_empty_and_ragged_batches(gpu_ctx):
    cfg = H.synthetic_config("c4", 40, psf_shape=(3, 7))
    m = cfg.product(gpu_ctx, max_batch=4)
    assert m.loglike_batch(np.zeros((0, m.npars), np.float32)).shape == (0,)
    P = H.workloads.param_batch(cfg.extra["workload"], 11)          # 4 + 4 + 3
    got = m.loglike_batch(P)
    for i, p in enumerate(P):
        _check_lnew(cfg, p, got[i], flat=False, tag=f"ragged[{i}]")
    with pytest.raises(ValueError):
        m.loglike_batch(np.zeros((3, m.npars + 1), np.float32))


@pytest.mark.parametrize("shape,psf", [((1, 1), None), ((1, 1), (3, 3)), ((1, 33), None), ((1, 33), (4, 1)), ((3, 1), (3, 3)),
                                       ((3, 1), (4, 1)), ((2, 32), None), ((2, 32), (3, 3)), ((5, 67), (4, 1))])
def test_tiny_and_ragged_images(gpu_ctx, shape, psf):
    """Images smaller than a warp, a tile or the PSF itself (every pixel an edge
    pixel, most lanes dead): images, chi^2 map and log-likelihood against the
    oracle; single point (graph path), a batch larger than the image, both
    render kernels."""
    import lensed_b200 as L
    h, w = shape
    rng = np.random.default_rng(h*100 + w)
    wl = H.workloads.c4(max(8, max(shape)))
    p = H.workloads.normalise_psf(rng.random((psf[1], psf[0])) + 0.1) if psf else None
    img = rng.random(shape).astype(np.float32)
    wht = (0.5 + rng.random(shape)).astype(np.float32)
    cfg = H.Config(f"tiny-{h}x{w}", wl["objects"], wl["truth"], img, wht, rule="sub2", psf=p)
    om = cfg.oracle()
    P = H.workloads.param_batch(wl, 40)
    ref = np.array([om.loglike(q) for q in P])
    for flags in (0, L.LCU_NO_PAIR):
        m = cfg.product(gpu_ctx, flags=flags)
        out = m.render(cfg.params)
        _, model, chi = om.loglike(cfg.params, want_maps=True)
        assert H.rel_err(out["model"], model).max() <= PIXEL_TOL, cfg.name
        assert np.allclose(out["chi"], chi, rtol=1e-4, atol=1e-6*max(chi.max(), 1e-30))
        got = m.loglike_batch(P)
        # a handful of pixels: no averaging over rounding noise, and chi^2 = w (m - d)^2 amplifies the 1e-5 of m by 2 m/|m - d|
        assert np.all(np.abs(got - ref) <= 1e-4*np.abs(ref)), (cfg.name, np.abs(got - ref).max()/np.abs(ref).max())
        assert m.loglike(P[3]) == got[3]
        m.close()


def test_full_size_c5_properties(gpu_ctx):
    """4096^2 / epl_plus_shear + 3 sersic + sky / g3k7 / 25x25 PSF."""
    w = H.workloads.c5(4096)
    img = np.zeros((4096, 4096), np.float32)
    m0 = H.Config("C5", w["objects"], w["truth"], img, np.ones_like(img), rule=w["rule"], psf=w["psf"]).product(gpu_ctx, flags=FAST)
    model = m0.render(w["truth"], raw=False, error=False, chi=False)["model"]
    m0.close()
    image, weight = H.workloads.observe(model, w["noise_seed"])
    m = H.Config("C5", w["objects"], w["truth"], image, weight, rule=w["rule"], psf=w["psf"]).product(gpu_ctx, flags=FAST)
    full = m.loglike(w["truth"])
    assert 0.9 < -2*full/image.size < 1.1
    parts = []
    for r0, r1 in ((0, 1000), (1000, 1001), (1001, 4096)):
        m.set_rows(r0, r1)
        parts.append(m.loglike(w["truth"]))
    assert abs(sum(parts) - full) <= 1e-12*abs(full)
    m.set_rows(0, 4096)
    P = H.workloads.param_batch(w, 3)
    assert np.array_equal(m.loglike_batch(P), np.array([m.loglike(p) for p in P]))
    # banded oracle check of 32 rows through the lens centre (pre-PSF image)
    raw = m.render(w["truth"], model=False, error=False, chi=False)["raw"]
    band = H.Config("C5-band", w["objects"], w["truth"], np.zeros((32, 4096), np.float32), np.ones((32, 4096), np.float32),
                    rule=w["rule"], pcs=(1.0, 2033.0, 1.0, 1.0))
    value, _ = band.oracle().render(w["truth"])
    v64, _ = band.oracle(variant="f64").render(w["truth"])
    r = H.rel_err(raw[2032:2064], value)
    floor = H.rel_err(value, v64).max()
    assert np.quantile(r, 0.999) <= PIXEL_TOL and r.max() <= max(PIXEL_TOL, 1.5*floor), f"{r.max():.3e} floor {floor:.3e}"


@pytest.mark.parametrize("flags", MATH_MODES)
def test_full_size_c5_loglike_against_reference_fixture(gpu_ctx, flags):
    """lnew of C5 at its full 4096^2 (822 M rays per evaluation) at the truth and
    at two perturbed points against the committed answers of the reference's own
    kernels compiled on the host (oracle/_ref) and of the oracle port in float32
    and float64 (tests/golden/c5_4096_lnew.npz, tools/make_c5_fixture.py).  The
    observed image is rebuilt here from the committed 512^2 model without any
    renderer in the loop; its SHA-256 proves both sides saw the same pixels."""
    import hashlib
    import os
    path = os.path.join(H.GOLDEN, "c5_4096_lnew.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c5_4096_lnew.npz not generated")
    z = np.load(path)
    image, weight = H.c5_fixture_observation(z["model512"])
    if hashlib.sha256(image.tobytes()).hexdigest() != str(z["image_sha256"]) or \
            hashlib.sha256(weight.tobytes()).hexdigest() != str(z["weight_sha256"]):
        pytest.skip("this machine's numpy draws a different noise realisation than the fixture's: nothing to compare")
    w = H.workloads.c5(4096)
    m = H.Config("C5-4096-fixture", w["objects"], w["truth"], image, weight, rule=w["rule"], psf=w["psf"]).product(gpu_ctx, flags=flags)
    P = z["params"]
    got = m.loglike_batch(P)
    assert got[0] == m.loglike(P[0])
    ref = z["lnew_ref"] if "lnew_ref" in z else z["lnew_strict"]
    l64 = z["lnew_f64"]
    for i in range(len(P)):
        floor = max(abs(z[k][i] - l64[i]) for k in ("lnew_ref", "lnew_strict") if k in z)
        rec = dict(case="C5-4096-fixture", what="lnew", tag=f"point {i} flags={flags}", gpu=float(got[i]), o32=float(ref[i]),
                   f64=float(l64[i]), rel_vs_o32=abs(got[i] - ref[i])/abs(ref[i]), rel_vs_f64=abs(got[i] - l64[i])/abs(l64[i]),
                   floor_rel=floor/abs(l64[i]), chi2_per_pixel=-2*l64[i]/image.size,
                   flat_ok=bool(abs(got[i] - ref[i]) <= LOGLIKE_TOL*abs(ref[i])), floor_ok=bool(abs(got[i] - l64[i]) <= FLOOR_K*floor),
                   flat_required=i == 0)
        _REPORT.append(rec)
        assert rec["flat_ok"] or (i > 0 and rec["floor_ok"]), rec


@pytest.mark.parametrize("name", ["test_sersic_bulge", "full_mock_psf", "full_mock_nopsf"])
def test_dumper_layers_against_oracle(gpu_ctx, name):
    """The six result layers of the dumper (src/nested.c:178-253) for C1 / C3 / C2
    from lcu_render + the host arithmetic, against the oracle's restatement:
    IMG and RAW to the image bound, RES = image - IMG exactly, WHT exactly, ERR
    (a ratio of a cancelling sum to the value) on the scale of the value, PVL
    absolutely (it is a probability)."""
    from lensed_b200 import host
    cfg = H.example_config(name)
    om, m = cfg.oracle(), cfg.product(gpu_ctx)
    ref = om.dumper_layers(cfg.params)
    got = host.dumper_layers(m, cfg.params, cfg.image, cfg.weight)
    assert set(got) == set(ref) == {"IMG", "RES", "RAW", "ERR", "WHT", "PVL"}
    for k in ("IMG", "RAW"):
        assert H.rel_err(got[k], ref[k]).max() <= PIXEL_TOL, k
    assert np.array_equal(got["RES"], cfg.image - got["IMG"]) and np.array_equal(got["WHT"], ref["WHT"])
    assert np.abs(got["RES"] - ref["RES"]).max() <= PIXEL_TOL*np.abs(ref["IMG"]).max()
    # ERR = error/value: the error estimate is a sum with alternating-sign weights, accurate to ~1e-5 of the value
    assert np.nanmax(np.abs(got["ERR"].astype(np.float64) - ref["ERR"])) <= 10*PIXEL_TOL
    # PVL = erfc(sqrt(chi^2/2)): d/dchi^2 is bounded by 1/sqrt(2 pi chi^2); compare absolutely
    assert np.abs(got["PVL"].astype(np.float64) - ref["PVL"]).max() <= 1e-4
    assert np.all((got["PVL"] >= 0) & (got["PVL"] <= 1))


def test_single_point_graph_and_profile(gpu_ctx, monkeypatch):
    """The one-point entry replays a CUDA graph; it must return exactly what
    the plain launch sequence returns, also after the row range changes, and
    the per-stage profile must account for every evaluation."""
    import lensed_b200 as L
    cfg = H.example_config("full_mock_psf")
    P = np.stack([cfg.params*(1 + 1e-3*i) for i in range(4)]).astype(np.float32)
    big = np.repeat(P, 8, axis=0)                            # 32 points: the batched path with its own set_params kernel
    # opt-in (LCU_FOLD_SETTER=1): small launches run set_params inside the render blocks -- two kernels
    # per point (render + set_params, convolve + reduction), one without a PSF -- with the bits of the batched path
    monkeypatch.setenv("LCU_FOLD_SETTER", "1")
    mf = cfg.product(gpu_ctx)
    ref = mf.loglike_batch(big)[::8]
    n0 = L.launch_count()
    assert np.array_equal(np.array([mf.loglike(p) for p in P]), ref)
    assert L.launch_count() - n0 == 4*2
    assert np.array_equal(mf.loglike_batch(P), ref)          # four points: folded as well
    mf.close()
    cfg0 = H.example_config("full_mock_nopsf")
    P0 = np.stack([cfg0.params*(1 + 1e-3*i) for i in range(4)]).astype(np.float32)
    mf0 = cfg0.product(gpu_ctx)
    ref0 = mf0.loglike_batch(np.repeat(P0, 8, axis=0))[::8]
    n0 = L.launch_count()
    assert np.array_equal(np.array([mf0.loglike(p) for p in P0]), ref0)
    assert L.launch_count() - n0 == 4*1
    mf0.close()
    monkeypatch.delenv("LCU_FOLD_SETTER")
    # the default: set_params as a kernel of its own, same bits
    m = cfg.product(gpu_ctx)
    batch = m.loglike_batch(P)
    assert np.array_equal(batch, ref)
    n0 = L.launch_count()
    single = np.array([m.loglike(p) for p in P])           # graph path
    assert np.array_equal(single, batch)
    assert L.launch_count() - n0 == 4*3                      # set_params, render, convolve + reduction per point
    monkeypatch.setenv("LCU_NO_GRAPH", "1")
    m2 = cfg.product(gpu_ctx)
    assert np.array_equal(np.array([m2.loglike(p) for p in P]), batch)
    monkeypatch.delenv("LCU_NO_GRAPH")
    # the reduction as a kernel of its own, and waiting on the stream instead of
    # watching the mapped result word: same bits
    for env in ("LCU_NO_FUSED_REDUCE", "LCU_NO_POLL"):
        monkeypatch.setenv(env, "1")
        m3 = cfg.product(gpu_ctx)
        n0 = L.launch_count()
        assert np.array_equal(np.array([m3.loglike(p) for p in P]), batch)
        assert L.launch_count() - n0 == 4*(4 if env == "LCU_NO_FUSED_REDUCE" else 3)
        assert np.array_equal(m3.loglike_batch(P), batch)
        m3.close()
        monkeypatch.delenv(env)
    # without a PSF the split render kernels add the partials up themselves
    cfg0 = H.example_config("full_mock_nopsf")
    P0 = np.stack([cfg0.params*(1 + 1e-3*i) for i in range(4)]).astype(np.float32)
    res = []
    for env in (None, "LCU_NO_FUSED_REDUCE"):
        if env:
            monkeypatch.setenv(env, "1")
        m4 = cfg0.product(gpu_ctx)
        n0 = L.launch_count()
        one = np.array([m4.loglike(p) for p in P0])
        assert L.launch_count() - n0 == 4*(3 if env else 2)
        res.append((one, m4.loglike_batch(P0), m4.loglike_batch(np.repeat(P0, 64, axis=0))[::64]))
        m4.close()
        if env:
            monkeypatch.delenv(env)
    for r in res:
        assert np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][0]) and np.array_equal(r[2], res[0][0])
    # many evaluations in a row: the block counters of the fused reduction return to zero every time
    again = np.array([m.loglike(P[i % 4]) for i in range(200)])
    assert np.array_equal(again, np.tile(batch, 50))
    m.set_rows(10, 60)
    a = m.loglike(P[0])
    m.set_rows(0, cfg.image.shape[0])
    assert m.loglike(P[0]) == batch[0] and a != batch[0]
    m.profile(True)
    m.loglike_batch(P)
    m.loglike(P[0])
    pr = m.profile_get()
    m.profile(False)
    assert pr["evaluations"] == 5
    assert pr["render_ms"] > 0 and pr["convolve_ms"] > 0 and pr["reduce_ms"] >= 0 and pr["set_params_ms"] > 0


@pytest.mark.parametrize("name", ["full_mock_nopsf", "full_mock_psf", "test_sersic_bulge"])
def test_one_kernel_point_path(gpu_ctx, monkeypatch, name):
    """Opt-in (LCU_FUSED_POINT=1; measured ~1 us slower than the default, so an
    experiment on record): one point of a small image with set_params, render,
    convolve + chi^2 and the final sum in ONE kernel (lcu_point_s*), hand-overs
    through global memory instead of kernel boundaries -- the bits of the three-kernel sequence and of
    the batched path, one launch per evaluation, also with two in flight, after a
    change of the row range, and over many evaluations (the hand-over words return
    to zero every time).  A kernel that gives up waiting (forced here: block 0
    never raises its flag) leaves the result word pending; the host then falls
    back to the three-kernel sequence for good and still returns the right value."""
    import lensed_b200 as L
    cfg = H.example_config(name)
    P = np.stack([cfg.params*(1 + 1e-3*i) for i in range(6)]).astype(np.float32)
    nk = 3 if cfg.psf is not None else 2
    m3 = cfg.product(gpu_ctx)
    ref = m3.loglike_batch(np.repeat(P, 8, axis=0))[::8]
    n0 = L.launch_count()
    assert np.array_equal(np.array([m3.loglike(p) for p in P]), ref)
    assert L.launch_count() - n0 == len(P)*nk
    monkeypatch.setenv("LCU_FUSED_POINT", "1")               # opt-in: measured slower than the three launches
    m = cfg.product(gpu_ctx)
    n0 = L.launch_count()
    assert np.array_equal(np.array([m.loglike(p) for p in P]), ref)
    assert L.launch_count() - n0 == len(P)                    # one kernel per evaluation
    assert np.array_equal(np.array([m.loglike(P[i % 6]) for i in range(300)]), np.tile(ref, 50))
    t0, t1 = m.loglike_async(P[0]), m.loglike_async(P[1])
    assert (m.loglike_wait(t0), m.loglike_wait(t1)) == (ref[0], ref[1])
    assert np.array_equal(m.loglike_batch(P), ref)            # batches are unaffected
    h = cfg.image.shape[0]
    m.set_rows(7, h - 9)
    m3.set_rows(7, h - 9)
    assert m.loglike(P[2]) == m3.loglike(P[2]) != ref[2]
    m.set_rows(0, h)
    assert m.loglike(P[2]) == ref[2]
    m.close(); m3.close()
    # the give-up path
    monkeypatch.setenv("LCU_POINT_TEST_TIMEOUT", "1")
    mt = cfg.product(gpu_ctx)
    assert mt.loglike(P[0]) == ref[0]                         # ~35 ms: every block's wait runs out, then the fallback
    monkeypatch.delenv("LCU_POINT_TEST_TIMEOUT")
    n0 = L.launch_count()
    assert mt.loglike(P[1]) == ref[1] and L.launch_count() - n0 == nk       # three-kernel sequence from now on
    mt.close()
    monkeypatch.setenv("LCU_POINT_TEST_TIMEOUT", "1")
    ma = cfg.product(gpu_ctx)
    t0, t1 = ma.loglike_async(P[2]), ma.loglike_async(P[3])
    assert ma.loglike_wait(t1) == ref[3] and ma.loglike_wait(t0) == ref[2]
    monkeypatch.delenv("LCU_POINT_TEST_TIMEOUT")
    assert ma.loglike(P[4]) == ref[4]
    ma.close()


@pytest.mark.parametrize("env", [None, "LCU_NO_GRAPH", "LCU_GRAPH_COPIES", "LCU_NO_POLL"])
def test_async_pair_same_bits_two_in_flight(gpu_ctx, monkeypatch, env):
    """lcu_loglike_async / lcu_loglike_wait: two evaluations in flight, results
    identical to lcu_loglike, a third start and any synchronous evaluation refused
    until a ticket is redeemed; small image (graph path, set_params folded in) and
    an image large enough for the two-rays kernel."""
    import lensed_b200 as L
    if env:
        monkeypatch.setenv(env, "1")
    for cfg in (H.example_config("full_mock_psf"), H.synthetic_config("c4", 160)):
        m = cfg.product(gpu_ctx)
        P = np.stack([cfg.params*(1 + 1e-3*i) for i in range(9)]).astype(np.float32)
        ref = m.loglike_batch(np.repeat(P, 8, axis=0))[::8]
        assert np.array_equal(np.array([m.loglike(p) for p in P]), ref)
        t0 = m.loglike_async(P[0])
        t1 = m.loglike_async(P[1])
        assert {t0, t1} == {0, 1}
        with pytest.raises(L.LensedCudaError):
            m.loglike_async(P[2])
        with pytest.raises(L.LensedCudaError):
            m.loglike(P[2])
        for call in (lambda: m.set_rows(0, 10), lambda: m.render(P[2]), lambda: m.set_params(P[2]),
                     lambda: m.set_data(image=cfg.image), lambda: m.loglike_batch(P)):
            with pytest.raises(L.LensedCudaError, match="in flight"):
                call()                                        # nothing may touch the model's state while tickets are out
        assert m.loglike_wait(t0) == ref[0]
        with pytest.raises(L.LensedCudaError):
            m.loglike_wait(t0)                                # redeemed already
        # keep two in flight: start i + 1, then collect i
        got = []
        t = m.loglike_async(P[2])                             # t1 (point 1) still in flight
        got.append(m.loglike_wait(t1))
        for i in range(3, len(P)):
            tn = m.loglike_async(P[i])
            got.append(m.loglike_wait(t))
            t = tn
        got.append(m.loglike_wait(t))
        assert np.array_equal(np.array(got), ref[1:])
        assert m.loglike(P[4]) == ref[4]                      # the synchronous call works again
        m.close()


@pytest.mark.parametrize("flags", MATH_MODES)
def test_every_object_in_one_model(gpu_ctx, flags):
    """Host galaxy + foreground + three lenses summed in one plane + four
    lensed sources (one with image-plane priors) + an object the reference
    cannot even load (sersic-old): exercises the generated compute() /
    set_params() beyond the reference's own configurations."""
    objects = ["sersic", "sky", "sie", "point_mass", "nsis", "gauss", "devauc", "exponential", "sersic-old"]
    params = np.array(
        [30.5, 30.5, 6.0, -4.0, 2.5, 0.8, 20.0,            # host sersic (unlensed)
         0.02, 1e-4, -2e-4,                                # sky with gradient
         30.5, 30.5, 12.0, 0.7, 60.0,                      # sie
         38.0, 27.0, 2.0,                                  # point_mass
         25.0, 35.0, 3.0, 1.5,                             # nsis
         44.0, 30.0, 1.5, -3.0, 0.9, 15.0,                 # gauss, image-plane position
         31.0, 32.0, 2.0, -3.5, 0.7, 100.0,                # devauc
         29.0, 30.0, 1.0, -2.5, 0.6, 45.0,                 # exponential
         32.0, 29.0, 1.5, -3.0, 1.2, 0.85, 70.0], np.float32)   # sersic-old
    img = np.zeros((60, 60), np.float32)
    cfg = H.Config("zoo", objects, params, img, np.ones_like(img), rule="g5k11", psf=H.workloads.gaussian_psf(7, 5, 1.2),
                   ipp=[[0]*7, [0]*3, [0]*5, [0]*3, [0]*4, [1, 1, 0, 0, 0, 0], [0]*6, [0]*6, [0]*7])
    om = cfg.oracle()
    _, model, _ = om.loglike(params, want_maps=True)
    cfg.image, cfg.weight = H.workloads.observe(model, 99, gain=50.0, offset=0.5)
    om = cfg.oracle()
    m = cfg.product(gpu_ctx, flags=flags)
    assert m.npars == params.size and m.words == 12 + 4 + 16 + 4 + 4 + 12*4
    out = m.render(params)
    _check_images(out, cfg, om, tag=f"flags={flags}")
    _check_lnew(cfg, params, m.loglike(params), flat=False, tag=f"flags={flags}", om=om)          # 60^2 pixels
    _check_block(m, om, cfg)


# all 24 seeds of helpers.random_config, the ill-conditioned ones included (seeds
# 0, 1, 6, 7, 18, 22: the strict oracle is itself 5e-6 ... 9e-6 from its float64
# twin at the 99.9th percentile; seed 20 has a pixel on a critical curve of its
# epl_plus_shear lens that moves by 1.6e-5 when sin / cos change by one ulp,
# tools/libm_sensitivity.py): held to flat-or-floor, both recorded
RANDOM_SEEDS = list(range(24))


@pytest.mark.parametrize("flags", MATH_MODES)
@pytest.mark.parametrize("seed", RANDOM_SEEDS)
def test_random_models(gpu_ctx, seed, flags):
    """Random combinations of the 15 objects (host galaxy, one or two lenses,
    one to three sources, sky), random image shapes, quadrature rules and
    PSFs: images, object block and log-likelihood against the oracle, and the
    log-likelihood of a batch against the single-point path."""
    cfg = H.random_config(seed)
    om = cfg.oracle()
    m = cfg.product(gpu_ctx, flags=flags)
    out = m.render(cfg.params)
    _check_images(out, cfg, om, tag=f"flags={flags}")
    got = m.loglike(cfg.params)
    _check_lnew(cfg, cfg.params, got, flat=False, tag=f"flags={flags}", om=om)
    _check_block(m, om, cfg, atol=1e-6)
    P = np.stack([cfg.params, cfg.params*np.float32(1.0005)])
    assert np.array_equal(m.loglike_batch(P)[0], got)


@pytest.mark.parametrize("seed", [2, 4, 8, 12, 14, 16])
def test_random_models_with_image_plane_priors(gpu_ctx, seed):
    """The same scenes with image-plane priors on every lensed source: the
    generated set_params shoots the position through all lenses of the plane
    (src/kernel.c:499-564; the oracle is pinned bit for bit on exactly these
    configurations, tests/test_oracle.py)."""
    cfg = H.random_config(seed)
    seen_lens, flags = False, []
    for o in cfg.objects:
        info = O.object_info(o)
        seen_lens = seen_lens or info["type"] == "L"
        flags.append([int(info["type"] == "S" and seen_lens and p["type"] in (1, 2)) for p in info["params"]])
    cfg.ipp = flags
    assert any(any(f) for f in flags)
    om = cfg.oracle()
    m = cfg.product(gpu_ctx, flags=FAST)
    _check_block(m, om, cfg, atol=1e-6)
    out = m.render(cfg.params)
    _check_images(out, cfg, om, tag="ipp")
    _check_lnew(cfg, cfg.params, m.loglike(cfg.params), flat=False, tag="ipp", om=om)


def test_pixel_coordinate_system(gpu_ctx):
    """Image sections: origin and pixel scale (src/data.c:236-276) enter the
    pixel positions (kernel/lensed.cl:24) and scale the quadrature abscissae
    (src/quadrature.c:38-39)."""
    w = H.workloads.c4(128)
    img = np.zeros((48, 56), np.float32)
    for pcs in ((33.0, 41.0, 1.0, 1.0), (20.0, 30.0, 2.0, 2.0), (20.5, 30.25, 1.5, -0.75)):
        cfg = H.Config(f"pcs{pcs}", w["objects"], w["truth"], img, np.ones_like(img), rule="g3k7", psf=H.workloads.gaussian_psf(5, 5, 1.0), pcs=pcs)
        om, m = cfg.oracle(), cfg.product(gpu_ctx)
        _check_images(m.render(cfg.params), cfg, om)


def _pair_cases():
    cases = [("golden", n) for n in H.golden_names()]
    cases += [("example", "test_sersic_bulge"), ("example", "full_mock_psf")]
    cases += [("synthetic", ("c4", 160, True)), ("synthetic", ("c5", 96, True)), ("synthetic", ("c4", 75, False))]
    return cases


@pytest.mark.parametrize("flags", MATH_MODES)
def test_two_rays_per_thread_same_bits(gpu_ctx, monkeypatch, flags):
    """The packed two-rays-per-thread render kernel (FADD2 / FMUL2 / FFMA2,
    shim.cuh) against the one-ray kernel on every shipped object: value and
    error images, convolved model, chi^2 map and log-likelihood are the same
    bits.  LCU_SPLIT=1 makes both models take their large-image kernel at
    these test sizes; odd pixel counts cover the dead-lane tail."""
    import lensed_b200 as L
    monkeypatch.setenv("LCU_SPLIT", "1")
    for kind, arg in _pair_cases():
        cfg = H.golden_config(arg) if kind == "golden" else H.example_config(arg) if kind == "example" \
            else H.synthetic_config(*arg[:2], psf=arg[2])
        m2 = cfg.product(gpu_ctx, flags=flags)
        m1 = cfg.product(gpu_ctx, flags=flags | L.LCU_NO_PAIR)
        assert m2.rays_per_thread == 2 and m1.rays_per_thread == 1, cfg.name
        a, b = m2.render(cfg.params), m1.render(cfg.params)
        for key in ("raw", "error", "model", "chi"):
            if a.get(key) is None:
                continue
            same = a[key].view(np.uint32) == b[key].view(np.uint32)
            assert same.all(), f"{cfg.name}: {key} differs in {np.count_nonzero(~same)} pixels, first {np.argwhere(~same)[0]}"
        assert m2.loglike(cfg.params) == m1.loglike(cfg.params), cfg.name
        if kind == "synthetic":
            P = H.workloads.param_batch(cfg.extra["workload"], 6)
            assert np.array_equal(m2.loglike_batch(P), m1.loglike_batch(P)), cfg.name


@pytest.mark.parametrize("libm_pair", ["0", "1"])
@pytest.mark.parametrize("flags", MATH_MODES)
def test_two_rays_per_thread_same_bits_packed_libm_switch(gpu_ctx, monkeypatch, flags, libm_pair):
    """atan2 / sincos / powr of pairs as packed arithmetic (the default,
    LCU_PF_LIBM_PAIR=1) and lane by lane through libdevice (=0): the power-law
    lens configurations and C5 against the one-ray kernel, bit for bit, with the
    switch in either position."""
    import lensed_b200 as L
    monkeypatch.setenv("LCU_SPLIT", "1")
    cases = [("golden", n) for n in H.golden_names() if n.startswith("epl")] + [("synthetic", ("c5", 96, True))]
    for kind, arg in cases:
        cfg = H.golden_config(arg) if kind == "golden" else H.synthetic_config(*arg[:2], psf=arg[2])
        monkeypatch.setenv("LCU_NVRTC_FLAGS", "-DLCU_PF_LIBM_PAIR=" + libm_pair)
        m2 = cfg.product(gpu_ctx, flags=flags)
        monkeypatch.delenv("LCU_NVRTC_FLAGS")
        m1 = cfg.product(gpu_ctx, flags=flags | L.LCU_NO_PAIR)
        assert m2.rays_per_thread == 2 and m1.rays_per_thread == 1, cfg.name
        a, b = m2.render(cfg.params), m1.render(cfg.params)
        for key in ("raw", "error", "model", "chi"):
            if a.get(key) is None:
                continue
            same = a[key].view(np.uint32) == b[key].view(np.uint32)
            assert same.all(), f"{cfg.name}: {key} differs in {np.count_nonzero(~same)} pixels, first {np.argwhere(~same)[0]}"
        assert m2.loglike(cfg.params) == m1.loglike(cfg.params), cfg.name


def test_two_rays_per_thread_guard_and_zoo(gpu_ctx, monkeypatch):
    """Non-finite deflections (per-lane guard) and a nine-object model."""
    import lensed_b200 as L
    monkeypatch.setenv("LCU_SPLIT", "1")
    img = np.zeros((9, 9), np.float32)
    for lens, lp in (("sis", [5.0, 5.0, 2.0]), ("point_mass", [5.0, 5.0, 2.0]), ("nsis", [5.0, 5.0, 2.0, 0.0])):
        params = np.array(lp + [5.0, 5.0, 1.5, -2.0, 1.0, 0.9, 10.0] + [0.25, 0.0, 0.0], np.float32)
        cfg = H.Config("guard-" + lens, [lens, "sersic", "sky"], params, img, np.ones_like(img), rule="point")
        a = cfg.product(gpu_ctx).render(params)
        b = cfg.product(gpu_ctx, flags=L.LCU_NO_PAIR).render(params)
        assert np.isfinite(a["raw"]).all()
        assert np.array_equal(a["raw"].view(np.uint32), b["raw"].view(np.uint32)), lens
    objects = ["sersic", "sky", "sie", "point_mass", "nsis", "epl", "sis_plus_shear", "nsie",
               "gauss", "devauc", "exponential", "sersic-old"]
    params = np.array(
        [30.5, 30.5, 6.0, -4.0, 2.5, 0.8, 20.0, 0.02, 1e-4, -2e-4, 30.5, 30.5, 12.0, 0.7, 60.0, 38.0, 27.0, 2.0,
         25.0, 35.0, 3.0, 1.5, 20.0, 40.0, 4.0, 1.2, 0.8, 10.0, 41.0, 42.0, 3.0, 0.05, -0.03, 15.0, 18.0, 2.5, 0.5, 0.6, 130.0,
         44.0, 30.0, 1.5, -3.0, 0.9, 15.0, 31.0, 32.0, 2.0, -3.5, 0.7, 100.0,
         29.0, 30.0, 1.0, -2.5, 0.6, 45.0, 32.0, 29.0, 1.5, -3.0, 1.2, 0.85, 70.0], np.float32)
    img = np.zeros((61, 53), np.float32)
    cfg = H.Config("zoo2", objects, params, img, np.ones_like(img), rule="gm75", psf=H.workloads.gaussian_psf(4, 5, 1.2))
    m2, m1 = cfg.product(gpu_ctx), cfg.product(gpu_ctx, flags=L.LCU_NO_PAIR)
    assert m2.npars == params.size and m2.rays_per_thread == 2
    a, b = m2.render(params), m1.render(params)
    for key in ("raw", "error", "model", "chi"):
        assert np.array_equal(a[key].view(np.uint32), b[key].view(np.uint32)), key


def test_device_data_preparation(gpu_ctx):
    """Weight map from gain / offset / mask built on the device
    (src/data.c:314-330, src/lensed.c:470-482) is the host computation bit for
    bit; swapping image and weights of a live model gives the log-likelihoods
    of a model created with them."""
    cfg = H.synthetic_config("c4", 64, psf_shape=(5, 5))
    rng = np.random.default_rng(3)
    gain_map = rng.uniform(500, 3000, cfg.image.shape).astype(np.float32)
    mask = (rng.random(cfg.image.shape) < 0.1).astype(np.int32)
    offset = 2.9633
    m = cfg.product(gpu_ctx)
    for gain in (np.float32(1800.0), gain_map):
        for mk in (None, mask):
            ref = (np.asarray(gain, np.float32).astype(np.float64)/(cfg.image.astype(np.float64) + offset)).astype(np.float32)
            if mk is not None:
                ref = np.where(mk != 0, np.float32(0), ref)
            got = m.make_weight(gain, offset, mk)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    P = H.workloads.param_batch(cfg.extra["workload"], 4)
    fresh = H.Config("fresh", cfg.objects, cfg.params, cfg.image, ref, rule=cfg.rule, psf=cfg.psf).product(gpu_ctx)
    assert np.array_equal(m.loglike_batch(P), fresh.loglike_batch(P))
    assert m.loglike(P[0]) == fresh.loglike(P[0])
    # another observation into the same model: no recompilation, same answers as a new model
    img2 = (cfg.image + rng.normal(0, 0.01, cfg.image.shape)).astype(np.float32)
    w2 = H.workloads.make_weight(img2)
    m.set_data(image=img2, weight=w2)
    fresh2 = H.Config("fresh2", cfg.objects, cfg.params, img2, w2, rule=cfg.rule, psf=cfg.psf).product(gpu_ctx)
    assert np.array_equal(m.loglike_batch(P), fresh2.loglike_batch(P))
    assert m.loglike(P[1]) == fresh2.loglike(P[1])        # the single-point graph sees the new data too
    assert np.array_equal(m.weight_map(), w2)
    with pytest.raises(ValueError):
        m.set_data(image=np.zeros((3, 3), np.float32))


@pytest.mark.parametrize("flags", MATH_MODES)
def test_two_rays_per_thread_underflow(gpu_ctx, monkeypatch, flags):
    """Far wings of compact sources without sky: brightness runs through the
    smallest normal numbers into underflow.  The scalar build flushes denormals
    at every operation, the packed multiply leaves that to its consumer: the
    images must still agree -- bit for bit wherever the value is a normal
    number, and to within the smallest normal number elsewhere."""
    import lensed_b200 as L
    monkeypatch.setenv("LCU_SPLIT", "1")
    img = np.zeros((96, 96), np.float32)
    for objects, params in (
            (["gauss"], [48.3, 47.6, 1.1, -12.0, 0.8, 20.0]),
            (["sie", "gauss", "exponential"], [48.5, 48.5, 15.0, 0.7, 30.0, 50.0, 47.0, 0.9, -9.0, 0.9, 10.0,
                                                46.0, 50.0, 0.35, -8.0, 0.8, 70.0]),
            (["sersic", "sersic"], [48.2, 48.9, 0.8, -10.0, 0.5, 0.9, 40.0, 60.0, 30.0, 0.3, -6.0, 0.8, 0.7, 100.0])):
        params = np.array(params, np.float32)
        cfg = H.Config("underflow", objects, params, img, np.ones_like(img), rule="sub2")
        a = cfg.product(gpu_ctx, flags=flags).render(params)["raw"]
        b = cfg.product(gpu_ctx, flags=flags | L.LCU_NO_PAIR).render(params)["raw"]
        tiny = np.float32(1.1754944e-38)
        assert (b == 0).any() or (np.abs(b) < 1e-30).any(), f"{objects}: the scene does not reach underflow"
        normal = np.abs(b) >= tiny
        assert np.array_equal(a[normal].view(np.uint32), b[normal].view(np.uint32)), objects
        assert np.all(np.abs(a[~normal].astype(np.float64) - b[~normal]) <= tiny), objects


@pytest.mark.parametrize("flags", MATH_MODES)
def test_exponential_overflow_is_infinite(gpu_ctx, flags):
    """exp of an argument above 88.7 is +inf, never NaN, in both math modes and
    in both render kernels (the compensated hardware exp multiplies 2^t = inf by
    a correction factor, which must not be negative there): an absurdly bright
    Sersic source overflows in the pixels around its centre and is an ordinary
    profile further out."""
    import lensed_b200 as L
    img = np.zeros((48, 48), np.float32)
    params = np.array([24.4, 24.7, 3.0, -103.0, 1.0, 0.8, 30.0], np.float32)
    cfg = H.Config("overflow", ["sersic"], params, img, np.ones_like(img), rule="sub4")
    ref, _ = cfg.oracle().render(params)
    r64, _ = cfg.oracle(variant="f64").render(params)
    fmax = float(np.finfo(np.float32).max)
    over, under = r64 > 10*fmax, r64 < 0.1*fmax      # pixels on the threshold may go either way
    assert over.sum() >= 8 and under.sum() > 1000 and np.isinf(ref[over]).all() and not np.isnan(ref).any()
    for extra in (0, L.LCU_NO_PAIR):
        got = cfg.product(gpu_ctx, flags=flags | extra).render(params)["raw"]
        assert not np.isnan(got).any()
        assert np.isinf(got[over]).all() and (got[over] > 0).all()
        assert H.rel_err(got[under], ref[under]).max() <= PIXEL_TOL


def test_fma_contraction_flag(gpu_ctx, monkeypatch):
    """LCU_FAST_MATH (FMA contraction, what -cl-fast-relaxed-math allows the
    reference's compiler) is opt-in: it stays close to the strict result but
    is not held to the 1e-5 bar.  Both render kernels
    honour it: the packed multiply is then issued with .ftz, which lets ptxas
    contract it into FFMA2."""
    import lensed_b200 as L
    monkeypatch.setenv("LCU_SPLIT", "1")
    cfg = H.synthetic_config("c4", 128)
    strict = cfg.product(gpu_ctx).render(cfg.params)["raw"]
    for extra in (0, L.LCU_NO_PAIR):
        m = cfg.product(gpu_ctx, flags=L.LCU_FAST_MATH | extra)
        r = H.rel_err(m.render(cfg.params)["raw"], strict)
        assert 0 < r.max() <= 1e-4 and np.quantile(r, 0.999) <= 2e-5, (extra, r.max())
    assert cfg.product(gpu_ctx, flags=L.LCU_FAST_MATH).rays_per_thread == 2
