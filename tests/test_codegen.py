"""Generated lcu_compute() / lcu_set_params_body(): object order is the
physics."""
import os
import re

import numpy as np
import pytest

import helpers as H
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
    assert lines[0].startswith("f = lcu_ray::brightness_sersic((struct lcu_ray::data_sersic*)(data + 0), y)")        # unlensed host
    assert lines[1].startswith("a = lcu_ray::deflection_sie((struct lcu_ray::data_sie*)(data + 12), y)")
    assert lines[2].startswith("y -= dot(a, a) < HUGE_VALF ? a : lcu_float2(1E10f, 1E10f)")        # non-finite guard
    assert lines[3].startswith("f += lcu_ray::brightness_sersic((struct lcu_ray::data_sersic*)(data + 28), y)")      # lensed source
    assert lines[4].startswith("f += lcu_ray::foreground_sky((struct lcu_ray::data_sky*)(data + 40), x)")            # image plane
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
    names = sorted(f[:-3] for f in os.listdir(H.OBJECTS_DIR) if f.endswith(".cl"))
    assert len(names) == 15
    for n in names:
        ok, why = compile_ctx.object_pairable(n)
        assert ok, f"{n}: {why}"
    m = L.Model(compile_ctx, ["sie_plus_shear", "sersic", "sky"], IMG, IMG)
    assert m.rays_per_thread == 2
    assert "lcu_pf lcu_compute2(const uint* data, lcu_pf2 x)" in m.source
    assert "a = lcu_pair::deflection_sie_plus_shear((struct lcu_ray::data_sie_plus_shear*)(data + 0), y);" in m.source
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
    shutil.copytree(H.OBJECTS_DIR, objdir)
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


TRICKY = r'''// a plugin written to trip a text-based loader: } { ; set( in comments and strings of macros
/* type = LENS;   params { {"nope"} };   static void set(local data* this) { } */
#define TWICE(v) ((v) + (v))   /* macros of the plugin's own: } */
#define AMPLITUDE(self) ((self)->amp)

type = SOURCE;

params {
    { "x", POSITION_X },   // comment with a brace }
    { "y", POSITION_Y },
    { "width", RADIUS, { 0.5f, 100.f }, 2.0f },
    { "amp", PARAMETER, POS_BOUND }
};

data
{
    float2 centre;   /* }; */
    float inv_w;
    float amp;
};

// helper functions before and after the entry points, with the OpenCL qualifiers the docs use
static float bump(float u)
{
    return exp(-0.5f*u);
}

static float brightness(constant data* this, float2 x)
{
    float2 d = (x - this->centre)*this->inv_w;
    float2 e = (float2)(TWICE(d.x), (float2)(d.y, 0.0f).x);
    return AMPLITUDE(this)*bump(0.25f*dot(e, e) + d.y*d.y*0.0f);
}

static float unused_helper(global data* this)
{
    return this->amp;   // set(this, 1, 2, 3, 4);
}

static void set(local data* this, float x, float y, float width, float amp)
{
    this->centre = (float2)(x, y);
    if(width > 0)       // branches on parameters are fine in set()
        this->inv_w = 1/width;
    else
        this->inv_w = 0;
    this->amp = amp;
}
'''


SWIZZLE = r'''type = LENS;
params { {"x", POSITION_X}, {"y", POSITION_Y}, {"r", RADIUS}, {"pa", POS_ANGLE} };
data { float2 c; mat22 m; float r; };
// a helper with vector arguments: every copy of the plugin text has its own,
// and argument-dependent lookup must not mix them up
static float2 rot(mat22 m, float2 v) { return (float2)(dot(m.lo, v), dot(m.hi, v)); }
static float2 deflection(local data* this, float2 x)
{
    float2 d = rot(this->m, x - this->c);
    float n = length(d);
    float2 u = normalize(d);
    float4 w = (float4)(u, d.s0, d.s1);          // mixed vector literal
    float4 z = (float4)(0.0f, u, 1.0f) + (float4)(n, n, w.zw);
    return this->r*(float2)(w.x + 0*z.z, w.s1 + 0*z.w)*fmin(n, 1.0f);
}
static void set(local data* this, float x, float y, float r, float pa)
{
    float c;
    float s = sincos(pa*DEG2RAD, &c);
    this->c = (float2)(x, y);
    this->m = (mat22)(c, s, -s, c);
    this->r = r;
}
'''


def test_plugin_text_with_comments_macros_helpers_and_qualifiers(tmp_path):
    """The loader works on text (name mangling by macros as src/kernel.c:153-172,
    vector literals rewritten, type / params / data / set() blanked out of the
    pair copy): braces and keywords in comments, the plugin's own macros, helper
    functions, nested vector literals and all three address-space qualifiers of
    docs/create.md must survive it -- in the scalar, the setter and the pair copy."""
    import shutil
    objdir = tmp_path / "objects"
    shutil.copytree(H.OBJECTS_DIR, objdir)
    (objdir / "tricky.cl").write_text(TRICKY)
    ctx = L.Context(device=-1, objects_dir=str(objdir))
    try:
        info = ctx.object_info("tricky")
        assert info.type == "S" and info.words == 4
        assert [p.name for p in info.params] == ["x", "y", "width", "amp"]
        assert info.params[2].bounds == (0.5, 100.0) and info.params[2].defval == 2.0
        assert info.params[3].bounds[0] == 0.0 and info.params[3].bounds[1] > 1e30
        ok, why = ctx.object_pairable("tricky")
        assert ok, why
        m = L.Model(ctx, ["sie", "tricky", "tricky", "sky"], IMG, IMG)
        assert m.rays_per_thread == 2 and m.npars == 5 + 4 + 4 + 3
        assert len(m.cubin) > 0
        # vector-typed helper, .lo/.hi/.s0/.zw swizzles, mixed vector literals; also as the
        # lens an image-plane prior is shot through (the setter copy calls its own helper)
        (objdir / "swizzle.cl").write_text(SWIZZLE)
        assert ctx.object_info("swizzle").words == 12 and ctx.object_pairable("swizzle")[0]
        m = L.Model(ctx, ["swizzle", "sie", "tricky"], IMG, IMG, ipp=[[0]*4, [0]*5, [1, 1, 0, 0]])
        assert m.rays_per_thread == 2 and "lcu_setter::deflection_swizzle" in _body(m.source, "void lcu_set_params_body")
    finally:
        ctx.close()


BUILTINS = r'''type = SOURCE;
params { {"x", POSITION_X}, {"y", POSITION_Y}, {"r", RADIUS}, {"a"} };
data { float2 c; float r; float a; float4 k; };

__constant float COEFF[4] = { 1.0f, 0.5f, 0.25f, 0.125f };
constant float SCALE = 2.0f;

static float poly(float u)
{
    float s = 0;
    for(int i = 0; i < 4; ++i)
        s = s*u + COEFF[i];
    return s;
}

static float brightness(__local data* this, float2 x)
{
    float2 d = (x - this->c)/this->r;
    float q = clamp(hypot(d.x, d.y), 0.0f, 10.0f);
    float t = fmax(fmin(q, 4.0f), 0.1f) + fabs(d.x)*0 + sign(d.y)*0 + step(0.5f, q)*0 + smoothstep(0.0f, 1.0f, q)*0;
    t += exp2(-q) + log1p(q)*0 + expm1(q)*0 + native_exp(-q)*0 + half_exp(-q)*0 + native_sqrt(q)*0 + rsqrt(q + 1.0f)*0;
    t += mix(0.0f, 1.0f, 0.5f)*0 + mad(q, 0.0f, 0.0f) + pown(q, 2)*0 + powr(q + 1, 0.5f)*0 + pow(q + 1, 2.0f)*0;
    t += degrees(0.0f) + radians(0.0f) + M_PI_F*0 + FLT_MAX*0 + FLT_EPSILON*0;
    t += (isnan(q) ? 1.0f : 0.0f)*0 + (isfinite(q) ? 0.0f : 1.0f) + (float)(isinf(q))*0;
    t += dot(this->k, this->k)*0 + length(this->k.xy)*0 + distance(d, d) + fast_length(d)*0;
    t += atan2(d.y, d.x)*0 + sinh(q)*0 + cosh(q)*0 + tanh(q)*0 + asin(0.5f)*0 + acos(0.5f)*0 + cbrt(q)*0 + erf(q)*0 + erfc(q)*0;
    t += floor(q)*0 + ceil(q)*0 + round(q)*0 + trunc(q)*0 + fmod(q, 2.0f)*0 + copysign(q, -1.0f)*0 + fdim(q, 1.0f)*0;
    return this->a*SCALE*poly(t);
}

static void set(__local data* this, float x, float y, float r, float a)
{
    this->c = (float2)(x, y);
    this->r = r;
    this->a = a*tgamma(2.0f)/exp(lgamma(2.0f));
    this->k = (float4)(1.0f);
    this->k.lo = (float2)(x, y);
    this->k.s3 = as_float(as_int(r));
    uint n = (uint)convert_int(r);
    this->k.z = (float)n + convert_float(3);
}
'''


def test_opencl_builtins_a_plugin_may_call(tmp_path):
    """OpenCL C built-ins beyond what the shipped objects use (common, math,
    geometric, reinterpretation and conversion functions, __-prefixed address
    spaces, program-scope constant arrays) compile in the scalar and the setter
    copy; the pair copy takes everything except the classification functions
    on ray values (isnan / isinf / isfinite: a per-ray decision), with which
    the model falls back to one ray per thread."""
    import shutil
    objdir = tmp_path / "objects"
    shutil.copytree(H.OBJECTS_DIR, objdir)
    (objdir / "builtins.cl").write_text(BUILTINS)
    (objdir / "builtins2.cl").write_text("\n".join(l for l in BUILTINS.splitlines() if "isnan(" not in l))
    ctx = L.Context(device=-1, objects_dir=str(objdir))
    try:
        assert ctx.object_info("builtins").words == 8
        ok, why = ctx.object_pairable("builtins")
        errors = [l for l in why.splitlines() if "error:" in l]
        assert not ok and len(errors) == 3 and all(("isnan" in e) or ("isinf" in e) or ("isfinite" in e) for e in errors), why
        assert L.Model(ctx, ["sie", "builtins"], IMG, IMG).rays_per_thread == 1
        ok, why = ctx.object_pairable("builtins2")
        assert ok, why
        assert L.Model(ctx, ["sie", "builtins2"], IMG, IMG).rays_per_thread == 2
    finally:
        ctx.close()


def test_object_file_encodings_and_function_specifiers(tmp_path):
    """Windows line ends, a UTF-8 byte order mark, a missing final newline;
    `static inline`, plain `inline` and non-static helper functions; the
    stray ';' after function bodies that docs/create.md:113,149 has."""
    import shutil
    objdir = tmp_path / "objects"
    src = H.OBJECTS_DIR
    shutil.copytree(src, objdir)
    sersic = open(os.path.join(src, "sersic.cl")).read()
    sie = open(os.path.join(src, "sie.cl")).read()
    (objdir / "crlf.cl").write_bytes(sersic.replace("\n", "\r\n").encode())
    (objdir / "bom.cl").write_bytes(b"\xef\xbb\xbf" + sersic.encode())
    (objdir / "nonl.cl").write_bytes(sersic.rstrip("\n").encode())
    helpers = sie.replace("static float2 deflection", "float sq(float v) { return v*v; }\ninline float tw(float v) { return v + v; }\n"
                          "static inline float2 deflection").replace("v.y*v.y", "sq(v.y) + 0*tw(v.x)")
    assert helpers != sie
    (objdir / "helpers.cl").write_text(helpers.replace("\n}\n", "\n};\n"))
    ctx = L.Context(device=-1, objects_dir=str(objdir))
    try:
        ref = ctx.object_info("sersic")
        for name in ("crlf", "bom", "nonl"):
            info = ctx.object_info(name)
            assert (info.type, info.words, [p.name for p in info.params]) == (ref.type, ref.words, [p.name for p in ref.params])
            assert ctx.object_pairable(name)[0]
        assert ctx.object_info("helpers").words == ctx.object_info("sie").words and ctx.object_pairable("helpers")[0]
        assert L.Model(ctx, ["helpers", "crlf", "bom", "nonl"], IMG, IMG).rays_per_thread == 2
    finally:
        ctx.close()


VECTOR_BUILTINS = r'''type = LENS;
params { {"x", POSITION_X}, {"y", POSITION_Y}, {"r", RADIUS} };
data { float2 c; float r; };
static float2 deflection(local data* this, float2 x)
{
    float2 d = x - this->c;
    float2 a = fabs(d);
    float2 b = sqrt(a*a + 1.0f);
    float2 e = exp(-a)*0.0f + log(b)*0.0f + sin(d)*0.0f + cos(d)*0.0f;
    float2 f = fmax(a, 0.5f) + fmin(a, (float2)(2.0f, 3.0f))*0.0f + clamp(d, -1.0f, 1.0f)*0.0f + mix(a, b, 0.5f)*0.0f;
    float2 g = pow(b, 2.0f)*0.0f + atan2(d, b)*0.0f + sign(d)*0.0f + floor(d)*0.0f;
    return this->r*d/(f + e + g);
}
static void set(local data* this, float x, float y, float r)
{
    this->c = (float2)(x, y);
    this->r = r;
}
'''


def test_gentype_builtins_on_vectors(tmp_path):
    """OpenCL math and common functions take vectors (gentype): exp(float2),
    fmax(float2, float), clamp(float2, float, float), ... -- component by
    component, in every copy of the text and under every math mode (the modes
    substitute function names, so each substituted name needs the vector forms
    too), also in the lens an image-plane prior is shot through."""
    import shutil
    objdir = tmp_path / "objects"
    shutil.copytree(H.OBJECTS_DIR, objdir)
    (objdir / "vecfn.cl").write_text(VECTOR_BUILTINS)
    ctx = L.Context(device=-1, objects_dir=str(objdir))
    try:
        assert ctx.object_info("vecfn").words == 4
        ok, why = ctx.object_pairable("vecfn")
        assert ok, why
        for flags in (0, L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH, L.LCU_FAST_INTRINSICS | L.LCU_FAST_LENS_INTRINSICS | L.LCU_FAST_ATANH):
            m = L.Model(ctx, ["vecfn", "sersic"], IMG, IMG, flags=flags, ipp=[[0]*3, [1, 1, 0, 0, 0, 0, 0]])
            assert m.rays_per_thread == 2
            assert "warning" not in m.build_log
    finally:
        ctx.close()
