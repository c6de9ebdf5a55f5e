"""Generated lcu_compute() / lcu_set_params_body(): object order is the
physics."""
import os
import re

import numpy as np
import pytest

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


def test_pair_copy_of_every_shipped_object(compile_ctx):
    """Every shipped object's per-ray function also compiles with float = a
    packed pair of rays (shim.cuh); the model then carries lcu_compute2 and the
    two-rays-per-thread kernels, unless LCU_NO_PAIR asks for one ray."""
    import os
    names = sorted(f[:-3] for f in os.listdir(os.path.join(os.path.dirname(L.__file__), "objects")) if f.endswith(".cl"))
    assert len(names) == 15
    for n in names:
        ok, why = compile_ctx.object_pairable(n)
        assert ok, f"{n}: {why}"
    m = L.Model(compile_ctx, ["sie_plus_shear", "sersic", "sky"], IMG, IMG)
    assert m.rays_per_thread == 2
    assert "lcu_pf lcu_compute2(const uint* data, lcu_pf2 x)" in m.source
    assert "a = lcu_pair::deflection_sie_plus_shear((struct data_sie_plus_shear*)(data + 0), y);" in m.source
    assert "y -= lcu_pair_guard(a);" in m.source
    m1 = L.Model(compile_ctx, ["sie_plus_shear", "sersic", "sky"], IMG, IMG, flags=L.LCU_NO_PAIR)
    assert m1.rays_per_thread == 1 and "lcu_pf lcu_compute2(" not in m1.source


def test_pair_copy_keeps_only_per_ray_code(compile_ctx):
    """type / params / data / set() are blanked out of the pair copy (line
    numbers preserved); helper functions and the per-ray function stay."""
    m = L.Model(compile_ctx, ["sersic"], IMG, IMG)
    src = m.source
    pair = src[src.index("namespace lcu_pair {"):src.index("} // namespace lcu_pair")]
    assert "brightness(" in pair and "tgamma" not in pair and "POSITION_X" not in pair and "half_inv_n;" not in pair
    scalar = src[src.index('#line 1 "objects/sersic.cl"'):src.index("namespace lcu_setter")]
    body = pair[pair.index('#line 1 "objects/sersic.cl"'):]
    # same number of lines up to the per-ray function: diagnostics point at the right place
    assert body[:body.index("brightness(")].count("\n") == scalar[:scalar.index("brightness(")].count("\n")


def test_object_that_cannot_be_paired_falls_back(tmp_path):
    """Branches on ray values, double arithmetic or integer casts cannot be
    typed as pairs: such an object is not an error, the model renders one ray
    per thread."""
    import shutil
    objdir = tmp_path / "objects"
    shutil.copytree(os.path.join(os.path.dirname(L.__file__), "objects"), objdir)
    (objdir / "ring.cl").write_text(
        "type = SOURCE;\n"
        "params { {\"x\", POSITION_X}, {\"y\", POSITION_Y}, {\"r\", RADIUS} };\n"
        "data { float2 c; float r; };\n"
        "static float brightness(local data* this, float2 x)\n"
        "{\n    float d = length(x - this->c);\n    if(d < this->r)\n        return 1.0f;\n    return d > 2*this->r ? 0.0f : 0.5f;\n}\n"
        "static void set(local data* this, float x, float y, float r)\n{\n    this->c = (float2)(x, y);\n    this->r = r;\n}\n")
    (objdir / "dbl.cl").write_text(
        "type = FOREGROUND;\nparams { {\"a\"} };\ndata { float a; };\n"
        "static float foreground(local data* this, float2 x)\n{\n    return this->a*0.5*x.x;\n}\n"
        "static void set(local data* this, float a)\n{\n    this->a = a;\n}\n")
    ctx = L.Context(device=-1, objects_dir=str(objdir))
    try:
        ok, why = ctx.object_pairable("ring")
        assert not ok and "ring.cl" in why
        ok, why = ctx.object_pairable("dbl")
        assert not ok
        assert ctx.object_pairable("sie")[0]
        m = L.Model(ctx, ["sie", "ring", "sky"], IMG, IMG)
        assert m.rays_per_thread == 1 and "lcu_pf lcu_compute2(" not in m.source
        assert L.Model(ctx, ["sie", "sersic", "sky"], IMG, IMG).rays_per_thread == 2
    finally:
        ctx.close()


def test_render_kernels_keep_the_object_block_out_of_vector_registers(compile_ctx):
    """The ray loops run in uniform control flow, so ptxas holds the object
    block (48 words for C4) in uniform registers: the two-rays kernel of the
    1024^2 benchmark model needs ~44 vector registers and no stack.  A branch
    on a per-thread condition around the loop would push it back to 80 (3
    resident blocks instead of 5)."""
    from lensed_b200 import workloads
    w = workloads.c4(1024)
    img = np.zeros((1024, 1024), np.float32)
    for flags in (0, L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH):
        m = L.Model(compile_ctx, w["objects"], img, img, rule=w["rule"], psf=w["psf"], flags=flags)
        regs, stack = m.kernel_usage("lcu_render_pair")
        assert regs <= 56 and stack == 0, (flags, regs, stack)
        assert "LDCU" not in m.source          # (the source is CUDA C++; the property is ptxas's doing)
    w5 = workloads.c5(4096)
    m = L.Model(compile_ctx, w5["objects"], img, img, rule=w5["rule"], psf=w5["psf"], flags=L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
    regs, stack = m.kernel_usage("lcu_render_pair")
    assert regs <= 85 and stack <= 64, (regs, stack)
    with pytest.raises(L.LensedCudaError):
        m.kernel_usage("no_such_kernel")
