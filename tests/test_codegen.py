"""Generated lcu_compute() / lcu_set_params_body(): object order is the
physics."""
import re

import numpy as np

import lensed_b200 as L

IMG = np.zeros((8, 8), np.float32)


def _body(src, fn):
    i = src.index(fn)
    return src[i:src.index("\n}\n", i)]


def test_compute_order_host_lens_source_sky(compile_ctx):
    m = L.Model(compile_ctx, ["sersic", "sie", "sersic", "sky"], IMG, IMG)
    body = _body(m.source, "float lcu_compute(")
    lines = [l.strip() for l in body.splitlines() if "+=" in l or "-=" in l]
    lines = [l.strip() for l in body.splitlines() if " = " in l and ("deflection_" in l or "brightness_" in l or "foreground_" in l) or "+=" in l or "-=" in l]
    assert lines[0].startswith("f = brightness_sersic((struct data_sersic*)(data + 0), y)")        # unlensed host
    assert lines[1].startswith("a = deflection_sie((struct data_sie*)(data + 12), y)")
    assert lines[2].startswith("y -= dot(a, a) < HUGE_VALF ? a : lcu_float2(1E10f, 1E10f)")        # non-finite guard
    assert lines[3].startswith("f += brightness_sersic((struct data_sersic*)(data + 28), y)")      # lensed source
    assert lines[4].startswith("f += foreground_sky((struct data_sky*)(data + 40), x)")            # image plane
    assert m.words == 44 and m.npars == 7 + 5 + 7 + 3


def test_each_object_compiled_once(compile_ctx):
    m = L.Model(compile_ctx, ["sie", "sersic", "sersic", "sersic"], IMG, IMG)
    assert len(re.findall(r"#define data struct data_sersic\n", m.source)) == 2      # hot copy + setter copy
    assert m.source.count("lcu_meta_sersic[3]") == 1


def test_image_plane_priors_shoot_through_the_lens(compile_ctx):
    m = L.Model(compile_ctx, ["sersic", "sie", "sersic"], IMG, IMG, ipp=[None, None, [1, 1, 0, 0, 0, 0, 0]])
    body = _body(m.source, "void lcu_set_params_body(")
    i_set_sie = body.index("lcu_setter::set_sie(")
    i_pos = body.index("x = lcu_float2(params[12], params[13]);")
    i_defl = body.index("a += lcu_setter::deflection_sie((struct lcu_setter::data_sie*)(data + 12), x);")
    i_set_src = body.index("lcu_setter::set_sersic((struct lcu_setter::data_sersic*)(data + 28), x.x, x.y, params[14]")
    assert i_set_sie < i_pos < i_defl < i_set_src
    # without the flag the position parameters are passed straight through
    m2 = L.Model(compile_ctx, ["sersic", "sie", "sersic"], IMG, IMG)
    assert "(data + 28), params[12], params[13], params[14]" in m2.source


def test_build_options_are_the_reference_macro_names(compile_ctx):
    m = L.Model(compile_ctx, ["sky"], np.zeros((5, 7), np.float32), np.zeros((5, 7), np.float32), rule="sub4",
                psf=np.ones((3, 4), np.float32)/12)
    for line in ("#define IMAGE_SIZE 35", "#define IMAGE_WIDTH 7", "#define IMAGE_HEIGHT 5", "#define PSF 1",
                 "#define PSF_WIDTH 4", "#define PSF_HEIGHT 3", "#define QUAD_POINTS 16"):
        assert line in m.source
