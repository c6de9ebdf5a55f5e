"""The C-ABI shared library: loads without a GPU, exports every symbol that
include/lensed_cuda.h declares, answers metadata / compile requests on a
compile-only context and refuses compute calls there (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import lensed_b200 as L
from lensed_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "lensed_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lensed_cuda.h but not exported"
    # and the Python binding covers the same set
    assert sorted(_lib.SYMBOLS) == names


def test_no_link_time_dependency_on_libcuda():
    """The library must load on machines without a driver (compile-only use)."""
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out
    assert "libnvrtc" in out


def test_version_and_rules():
    assert _lib.lib.lcu_version() == 100
    rules = dict(L.quad_rules())
    assert list(rules) == ["point", "sub2", "sub4", "gm75", "g3k7", "g5k11", "g7k15"]
    with pytest.raises(ValueError):
        L.quad_rule("nope")


def test_compile_only_context_refuses_compute(compile_ctx):
    img = np.zeros((16, 16), np.float32)
    m = L.Model(compile_ctx, ["sie", "sersic"], img, img)
    assert m.npars == 12 and m.words == 28
    assert len(m.cubin) > 1000 and m.cubin[:4] == b"\x7fELF"
    for call in (lambda: m.loglike(np.zeros(12)), lambda: m.loglike_batch(np.zeros((2, 12))),
                 lambda: m.render(np.zeros(12)), lambda: m.set_params(np.zeros(12)),
                 lambda: m.loglike_batch_device(1, 1, 1)):
        with pytest.raises(L.LensedCudaError) as e:
            call()
        assert e.value.code == 5 and "no CPU fallback" in e.value.message or "compile-only" in e.value.message
    with pytest.raises(L.LensedCudaError):
        compile_ctx.fp32_peak_tflops()


def test_source_only_model(compile_ctx):
    """LCU_SOURCE_ONLY: the program text without building it (what the reference
    writes as <root>kernel.cl before its build, src/lensed.c:714-735)."""
    img = np.zeros((16, 16), np.float32)
    full = L.Model(compile_ctx, ["sie", "sersic"], img, img)
    text = L.Model(compile_ctx, ["sie", "sersic"], img, img, flags=L.LCU_SOURCE_ONLY)
    assert text.source == full.source and "lcu_compute2" in text.source
    assert (text.npars, text.words, text.rays_per_thread) == (full.npars, full.words, full.rays_per_thread)
    assert len(text.cubin) == 0
    with pytest.raises(L.LensedCudaError):
        text.kernel_usage("lcu_render_pair")
    with pytest.raises(L.LensedCudaError) as e:
        text.loglike(np.zeros(12))
    assert e.value.code == 5


def test_error_reporting(compile_ctx, tmp_path):
    img = np.zeros((8, 8), np.float32)
    with pytest.raises(L.LensedCudaError) as e:
        compile_ctx.object_info("no_such_object")
    assert e.value.code == 2 and 'could not load object "no_such_object"' in e.value.message
    with pytest.raises(L.LensedCudaError) as e:
        L.Model(compile_ctx, ["sie", "sersic", "sis", "sersic"], img, img)
    assert "multiple lensing planes are not supported" in e.value.message
    with pytest.raises(L.LensedCudaError) as e:
        L.Model(compile_ctx, ["sie", "sersic"], img, img, ipp=[None, [1, 0, 0, 0, 0, 0, 0]])
    assert "image plane prior requires pair (X,Y)" in e.value.message
    # a broken plugin: the NVRTC log comes back in the error message
    (tmp_path / "broken.cl").write_text("type = LENS;\nparams { {\"x\"} };\ndata { float a; };\n"
                                        "static float2 deflection(local data* this, float2 x) { return undefined_symbol; }\n"
                                        "static void set(local data* this, float x) { this->a = x; }\n")
    ctx = L.Context(device=-1, objects_dir=str(tmp_path))
    with pytest.raises(L.LensedCudaError) as e:
        ctx.object_info("broken")
    assert e.value.code == 3 and "undefined_symbol" in e.value.message
    # wrong type value
    (tmp_path / "badtype.cl").write_text("type = 7;\nparams { {\"x\"} };\ndata { float a; };\n"
                                         "static float brightness(local data* this, float2 x) { return 0; }\n"
                                         "static void set(local data* this, float x) { this->a = x; }\n")
    with pytest.raises(L.LensedCudaError) as e:
        ctx.object_info("badtype")
    assert "invalid type" in e.value.message
    with pytest.raises(L.LensedCudaError):
        L.Context(device=-1, kernel_dir=str(tmp_path))


def test_user_plugin_without_recompiling_host(compile_ctx, tmp_path):
    """docs/create.md:8-10: a new object file is picked up at run time; vector
    literals, swizzles, address-space qualifiers and OpenCL built-ins work."""
    (tmp_path / "ring.cl").write_text('''
type = SOURCE;
params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "w", RADIUS, POS_BOUND, 0.5f }, { "amp", PARAMETER, NEG_BOUND } };
data { float2 c; float4 k; float r; float w; float amp; };
static float brightness(constant data* this, float2 x)
{
    float2 d = x - this->c;
    float rad = length(d) - this->r;
    float2 u = (float2)(this->k.lo.x, this->k.hi.s1)*normalize(d);
    return this->amp*exp(-0.5f*rad*rad/(this->w*this->w))*(1 + 0*dot(u, IMAGE_CENTER));
}
static void set(global data* this, float x, float y, float r, float w, float amp)
{
    this->c = (float2)(x, y);
    this->k = (float4)(1, 2, 3, 4);
    this->r = r; this->w = w; this->amp = amp;
}
''')
    ctx = L.Context(device=-1, objects_dir=str(tmp_path))
    info = ctx.object_info("ring")
    assert info.type == "S" and info.words == 12 and [p.name for p in info.params] == ["x", "y", "r", "w", "amp"]
    assert info.params[3].bounds[0] == 0 and info.params[3].bounds[1] > 3e38 and info.params[3].has_default
    assert info.params[4].bounds[0] < -3e38 and not info.params[4].has_default
    img = np.zeros((8, 8), np.float32)
    m = L.Model(ctx, ["ring"], img, img)
    assert "brightness_ring" in m.source and m.npars == 5


def test_constant_bank_budget_is_checked(compile_ctx):
    """Quadrature table + PSF + object block have to fit the 64 KB constant
    bank: a rule that cannot is a clean LCU_E_ARG with the sizes in the message,
    not an NVRTC / ptxas failure."""
    img = np.zeros((8, 8), np.float32)
    n = 4100
    qq = np.zeros((n, 2), np.float32)
    ww = np.full((n, 2), 1.0/n, np.float32)
    with pytest.raises(L.LensedCudaError, match="constant memory"):
        L.Model(compile_ctx, ["sersic"], img, img, qq=qq, ww=ww)
    # a large but admissible rule still builds
    m = L.Model(compile_ctx, ["sersic"], img, img, qq=qq[:1024], ww=ww[:1024])
    assert m.nq == 1024 and m.max_batch >= 1
