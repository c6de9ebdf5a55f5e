"""Host-side mirror (lensed_b200/host.py): priors, ini reader, parameter map,
unit-cube transform -- and, on a GPU, a reference-style end-to-end run of an
ini file with every parameter fixed (how the reference's own test-suite works,
tests/Makefile:27-28: min chi^2/dof of a fully specified model ~ 0)."""
import json
import os

import numpy as np
import pytest

import helpers as H
import lensed_b200 as L
from lensed_b200 import fits, host

INI = """
image  = img.fits
gain   = 1800
offset = 2.9633
psf    = psf.fits
rule   = g3k7

[objects]
host   = sersic
lens   = sie
source = sersic
sky    = sky

[priors]
host.x     = unif 55 65
host.y     = unif 55 65
host.r     = unif 1 50
host.mag   = norm -5 2
host.n     = 4.
host.q     = unif 0.5 1
host.pa    = wrap unif 0 180
lens.x     = unif 55 65
lens.y     = unif 55 65
lens.r     = unif 20 35
lens.q     = unif 0.1 1
lens.pa    = wrap unif 0 180
source.x   = image unif 80 85
source.y   = image unif 35 40
source.r   = unif 0.1 20
source.mag = unif -5 5
source.n   = unif 0.5 8.0
source.q   = unif 0.1 1
source.pa  = wrap unif 0 180
sky.bg     = unif 0 1

[labels]
host.x = x_H
"""


def _write_case(tmp_path):
    rng = np.random.default_rng(3)
    fits.write_layers(str(tmp_path / "img.fits"), [rng.random((40, 48)).astype(np.float32) + 1], ["IMG"])
    fits.write_layers(str(tmp_path / "psf.fits"), [rng.random((5, 5)).astype(np.float32)], ["PSF"])
    (tmp_path / "run.ini").write_text(INI)
    return str(tmp_path / "run.ini")


def test_priors():
    assert host.read_prior("4.").apply(0.3) == 4.0 and host.read_prior("4.").pseudo
    u = host.read_prior("unif 55 65")
    assert u.apply(0.0) == 55 and u.apply(1.0) == 65 and u.apply(0.25) == 57.5 and not u.pseudo
    n = host.read_prior("norm -5 2")
    assert abs(n.apply(0.5) + 5) < 2e-3                       # A&S 26.2.23 is accurate to 4.5e-4
    assert abs(n.apply(0.8413447) - (-5 + 2)) < 2e-3          # +1 sigma
    assert n.apply(0.1) == -10 - n.apply(0.9)                 # antisymmetric about the mean
    assert (n.lower(), n.upper()) == (-19, 9)
    for spec, msg in (("cauchy 0 1", "unknown prior: cauchy"), ("delta 1", "unknown prior: delta"),
                      ("unif 1", "invalid prior definition"), ("norm 0 x", "invalid prior definition")):
        with pytest.raises(ValueError) as e:                  # src/prior.c:119-147
            host.read_prior(spec)
        assert msg in str(e.value)


def test_ini_parameter_map_and_transform(compile_ctx, tmp_path):
    cfg, model, like = host.build(_write_case(tmp_path), compile_ctx)
    assert [o.name for o in cfg.objects] == ["sersic", "sie", "sersic", "sky"]
    assert like.npars == 7 + 5 + 7 + 3
    # derived: host.n (delta 4.) and the sky gradients (default -0.0f): last in sampler order
    assert like.ndims == like.npars - 3
    assert [like.pars[i].id for i in like.pmap[-3:]] == ["host.n", "sky.dx", "sky.dy"]
    src = next(o for o in cfg.objects if o.id == "source")
    assert [p.ipp for p in src.params] == [True, True] + [False]*5
    assert next(p for p in cfg.parameters if p.id == "host.pa").wrap
    assert next(p for p in cfg.parameters if p.id == "host.x").label == "x_H"
    # image-plane priors reach the generated set_params
    assert "x = lcu_float2(params[12], params[13]);" in model.source
    # even-free 5x5 PSF, section-free image: model geometry
    assert (model.width, model.height, model.has_psf) == (48, 40, True)
    cube = np.full(like.npars, 0.5)
    phys = like.physical(cube)
    params = like.device_params(phys)
    assert params.dtype == np.float32
    ids = [p.id for p in like.pars]
    assert params[ids.index("host.n")] == 4.0 and params[ids.index("host.x")] == 60.0
    assert params[ids.index("sky.dx")] == 0.0 and np.signbit(params[ids.index("sky.dx")])
    assert params[ids.index("lens.r")] == 27.5
    # default bounds: radius [0, inf), axis ratio [0, 1] (src/lensed.c:148-167)
    r = next(p for p in like.pars if p.id == "lens.r")
    q = next(p for p in like.pars if p.id == "lens.q")
    assert (r.lower, r.upper) == (0.0, float("inf")) and (q.lower, q.upper) == (0.0, 1.0)


def test_ini_errors(compile_ctx, tmp_path):
    path = _write_case(tmp_path)
    text = open(path).read()
    for bad, msg in ((text.replace("lens.q     = unif 0.1 1", "lens.q     = unif 2 3"), "prior does not include parameter bounds"),
                     (text.replace("sky    = sky", "sky    = sky\nlens2 = sis\nsrc2 = gauss\nlens3 = sis"), "multiple lensing planes"),
                     (text.replace("host.x     = unif 55 65", "nobody.x = 1"), "unknown object"),
                     (text.replace("sky.bg     = unif 0 1", ""), "missing prior: sky.bg"),
                     (text.replace("gain   = 1800", ""), "missing required option: gain"),
                     (text.replace("image  = img.fits", ""), "missing required option: image"),
                     (text.split("[objects]")[0], "no objects were given")):
        open(path, "w").write(bad)
        with pytest.raises(ValueError) as e:
            host.build(path, compile_ctx)
        assert msg in str(e.value)


def test_ini_syntax_of_the_reference(compile_ctx, tmp_path):
    """src/input/ini.c:16-25: '=' or ':' assigns, ';' or '#' ends a line,
    groups are [name] with optional blanks, unknown / unclosed groups are errors."""
    path = _write_case(tmp_path)
    text = open(path).read()
    alt = (text.replace("gain   = 1800", "gain   : 1800   # electrons per count")
               .replace("host.x     = unif 55 65", "host.x: unif 55 65 ; centre")
               .replace("[priors]", "[ priors ]"))
    open(path, "w").write(alt)
    cfg, _, like = host.build(path, compile_ctx)
    assert cfg.options["gain"] == "1800" and like.npars == 7 + 5 + 7 + 3
    assert cfg.objects[0].params[0].prior.apply(0.5) == 60.0
    for bad, msg in ((text.replace("[priors]", "[priors"), "missing closing ']'"),
                     (text.replace("[labels]", "[lables]"), "unknown group: lables"),
                     (text.replace("rule   = g3k7", "rule g3k7"), "does not assign anything")):
        open(path, "w").write(bad)
        with pytest.raises(ValueError) as e:
            host.build(path, compile_ctx)
        assert msg in str(e.value)


REFERENCE = os.environ.get("LENSED_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests")), reason="reference tree not present")
def test_reads_every_ini_file_of_the_reference(compile_ctx):
    """the reference's three examples and 16 test configurations, unmodified:
    objects, priors (incl. `wrap` / `image` keywords) and options come out as the files say"""
    import glob
    files = sorted(glob.glob(os.path.join(REFERENCE, "examples", "*.ini")) + glob.glob(os.path.join(REFERENCE, "tests", "*", "*.ini")))
    assert len(files) == 19
    for path in files:
        cfg = host.read_ini(path, compile_ctx)
        assert cfg.objects and "image" in cfg.options, path
        for p in cfg.parameters:
            assert p.prior is not None, (path, p.id)
        name = os.path.basename(path)[:-4]
        if "tests" in path.split(os.sep):
            assert all(p.prior.pseudo for p in cfg.parameters), path        # fully specified models
            assert any(o.name == name.split("-")[0] for o in cfg.objects), path   # tests/lens/sie.ini tests `sie`
    cfg = host.read_ini(os.path.join(REFERENCE, "examples", "test_sersic_bulge.ini"), compile_ctx)
    by_id = {p.id: p for p in cfg.parameters}
    assert by_id["source.x"].ipp and by_id["source.y"].ipp and by_id["lens.pa"].wrap and not by_id["lens.x"].ipp


@pytest.mark.gpu
def test_reference_style_known_answer_run(gpu_ctx, tmp_path):
    """tests/lens/sie.ini of the reference, reproduced: all parameters fixed,
    weight = 1000, golden image -> min chi^2/dof."""
    with np.load(os.path.join(H.GOLDEN, "ref_goldens.npz")) as z:
        meta = json.loads(str(z["meta"]))["sie"]
        fits.write_layers(str(tmp_path / "sie.fits"), [z["sie"]], ["IMG"])
    lines = ["image = sie.fits", "weight = 1000", "[objects]"] + [f"{i} = {n}" for i, n in meta["objects"]]
    lines += ["[priors]"] + [f"{k} = {v}" for k, v in meta["priors"].items()]
    (tmp_path / "sie.ini").write_text("\n".join(lines) + "\n")
    cfg, model, like = host.build(str(tmp_path / "sie.ini"), gpu_ctx)
    assert like.ndims == 0                                   # everything is a delta pseudo-prior
    lnew = like(np.zeros(like.npars))
    chi2_dof = -2*lnew/(model.width*model.height)
    assert 0 <= chi2_dof < 5e-3
    ref = H.golden_config("sie")
    assert abs(lnew - ref.oracle().loglike(ref.params)) <= 1e-6*ref.image.size
    # dumper layers of that point
    layers = host.dumper_layers(model, like.device_params(like.physical(np.zeros(like.npars))), ref.image, ref.weight)
    assert set(layers) == {"IMG", "RES", "RAW", "ERR", "WHT", "PVL"}
    assert np.allclose(layers["RES"], ref.image - layers["IMG"]) and np.all((layers["PVL"] >= 0) & (layers["PVL"] <= 1))
    host.write_results(str(tmp_path / "out.fits"), layers)
    back = fits.read_hdus(str(tmp_path / "out.fits"))
    # 7 HDUs as the reference writes them: empty primary, then IMG = HDU 1 ... PVL = HDU 6 (src/data.c:129-165)
    assert len(back) == 7 and back[0][1] is None
    assert [h["EXTNAME"] for h, _ in back[1:]] == ["IMG", "RES", "RAW", "ERR", "WHT", "PVL"]
    assert np.array_equal(back[1][1], layers["IMG"])


def test_find_mode_and_mask(compile_ctx, tmp_path):
    rng = np.random.default_rng(0)
    img = (0.3 + 0.05*rng.standard_normal((64, 64))).astype(np.float32)
    img[:4, :4] = 50.0                                        # a bright source does not move the mode
    mode, fwhm = host.find_mode(img, np.ones_like(img))
    assert abs(mode - 0.3) < 0.3 and 0 < fwhm                 # 100 bins over [min, 50]: coarse, like the reference
    mode, fwhm = host.find_mode(img[8:], None)
    assert abs(mode - 0.3) < 0.02 and abs(fwhm - 2.355*0.05) < 0.05
    assert host.find_mode(np.full(10, 2.0)) == (2.0, 0.0)
    # mask option: masked pixels get zero weight
    path = _write_case(tmp_path)
    mask = np.zeros((40, 48), np.float32)
    mask[5:9, 7:11] = 1
    fits.write_layers(str(tmp_path / "mask.fits"), [mask], ["MASK"])
    open(path, "w").write(INI.replace("rule   = g3k7", "rule   = g3k7\nmask   = mask.fits"))
    cfg, model, like = host.build(path, compile_ctx)
    assert model.npars == 22
