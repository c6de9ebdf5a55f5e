"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs, never by the product package.

Two interchangeable back-ends export the same C entry points:
  * ``liboracle*.so``  -- the restatement in oracle/lensed_oracle.c ("port");
  * ``_ref/liblensed_ref*.so`` -- the reference's own objects/*.cl,
    kernel/lensed.cl and generated compute()/set_params() text compiled on the
    host by oracle/build_ref.py ("reference"; only for the model
    configurations baked into it at build time).
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_VARIANTS = {
    "strict": "liboracle.so",
    "f64": "liboracle_f64.so",
    "fast": "liboracle_fast.so",
    "ref": os.path.join("_ref", "liblensed_ref.so"),
    "ref_fast": os.path.join("_ref", "liblensed_ref_fast.so"),
    # the same reference text with float = 8 consecutive work-items (AVX2): timing only
    "ref_simd": os.path.join("_ref", "liblensed_ref_simd.so"),
    "ref_simd512": os.path.join("_ref", "liblensed_ref_simd512.so"),
}

_libs: dict = {}


def _cpu_has(flag: str) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


def available(variant: str = "strict") -> bool:
    if variant == "ref_simd512" and not _cpu_has("avx512f"):
        return False
    if variant == "ref_simd" and not _cpu_has("avx2"):
        return False
    return os.path.exists(os.path.join(HERE, _VARIANTS[variant]))


def lib(variant: str = "strict"):
    if variant in _libs:
        return _libs[variant]
    path = os.path.join(HERE, _VARIANTS[variant])
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle`")
    L = C.CDLL(path)
    L.orc_real_size.restype = C.c_int
    L.orc_object_count.restype = C.c_int
    L.orc_object_name.restype = C.c_char_p
    L.orc_object_name.argtypes = [C.c_int]
    L.orc_object_info.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.orc_object_param.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.orc_quad_size.argtypes = [C.c_char_p]
    L.orc_quad_rule.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    L.orc_model_create.restype = C.c_void_p
    L.orc_model_create.argtypes = [C.c_size_t, C.POINTER(C.c_char_p), C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                   C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.orc_model_free.argtypes = [C.c_void_p]
    L.orc_model_npars.restype = C.c_size_t
    L.orc_model_npars.argtypes = [C.c_void_p]
    L.orc_model_words.restype = C.c_size_t
    L.orc_model_words.argtypes = [C.c_void_p]
    L.orc_set_params.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_convolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_loglike.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
    if hasattr(L, "orc_dumper_layers"):
        L.orc_dumper_layers.argtypes = [C.c_void_p]*8
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.orc_make_weight.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_size_t, C.c_void_p]
    L.orc_normalise_psf.argtypes = [C.c_void_p, C.c_size_t]
    _libs[variant] = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def object_info(name: str, variant: str = "strict"):
    L = lib(variant)
    t, w, n = C.c_int(), C.c_size_t(), C.c_size_t()
    if L.orc_object_info(name.encode(), C.byref(t), C.byref(w), C.byref(n)):
        raise KeyError(name)
    pars = []
    for j in range(n.value):
        buf = C.create_string_buffer(16)
        pt, b, d = C.c_int(), (C.c_float * 2)(), C.c_float()
        L.orc_object_param(name.encode(), j, buf, C.byref(pt), b, C.byref(d))
        pars.append(dict(name=buf.value.decode(), type=pt.value, bounds=(b[0], b[1]), defval=d.value,
                         defval_bits=np.float32(d.value).view(np.uint32).item()))
    return dict(type=chr(t.value), words=w.value, npar=n.value, params=pars)


def object_names(variant: str = "strict"):
    L = lib(variant)
    return [L.orc_object_name(i).decode() for i in range(L.orc_object_count())]


def quad_rule(rule: str, sx: float = 1.0, sy: float = 1.0, variant: str = "strict"):
    L = lib(variant)
    n = L.orc_quad_size(rule.encode())
    if n < 0:
        raise KeyError(rule)
    qq = np.zeros((n, 2), np.float32)
    ww = np.zeros((n, 2), np.float32)
    L.orc_quad_rule(rule.encode(), sx, sy, _ptr(qq), _ptr(ww))
    return qq, ww


def make_weight(image, gain, offset):
    image = _f32(image)
    gain = _f32(np.broadcast_to(np.asarray(gain, np.float32), image.shape))
    w = np.empty_like(image)
    lib().orc_make_weight(_ptr(image), _ptr(gain), float(offset), image.size, _ptr(w))
    return w


def normalise_psf(psf):
    psf = _f32(psf).copy()
    lib().orc_normalise_psf(_ptr(psf), psf.size)
    return psf


class Model:
    """Mirror of the device state src/lensed.c:644-1112 builds, on the CPU."""

    def __init__(self, objects, image, weight, qq, ww, psf=None, pcs=(1, 1, 1, 1), ipp=None, variant="strict", _lib=None):
        self.L = _lib if _lib is not None else lib(variant)
        self.real = np.float64 if self.L.orc_real_size() == 8 else np.float32
        image = _f32(image)
        self.height, self.width = image.shape
        names = (C.c_char_p * len(objects))(*[o.encode() for o in objects])
        self._keep = (names, image, _f32(weight), _f32(qq), _f32(ww), _f32(np.asarray(pcs)),
                      _f32(psf) if psf is not None else None,
                      np.ascontiguousarray(ipp, dtype=np.int32) if ipp is not None else None)
        _, image, weight, qq, ww, pcs, psf, ippa = self._keep
        ph, pw = psf.shape if psf is not None else (0, 0)
        self.h = self.L.orc_model_create(len(objects), names, _ptr(ippa), self.width, self.height, _ptr(pcs),
                                         qq.shape[0], _ptr(qq), _ptr(ww), _ptr(image), _ptr(weight), _ptr(psf), pw, ph)
        if not self.h:
            raise ValueError(f"oracle cannot build model for objects {objects}")
        self.npars = self.L.orc_model_npars(self.h)
        self.words = self.L.orc_model_words(self.h)
        self.has_psf = psf is not None

    def __del__(self):
        try:
            if self.h:
                self.L.orc_model_free(self.h)
                self.h = None
        except Exception:
            pass

    def set_params(self, params):
        p = _f32(params)
        assert p.size == self.npars
        block = np.zeros(self.words, self.real)
        self.L.orc_set_params(self.h, _ptr(p), _ptr(block))
        return block

    def render(self, params):
        p = _f32(params)
        assert p.size == self.npars
        v = np.zeros((self.height, self.width), self.real)
        e = np.zeros_like(v)
        self.L.orc_render(self.h, _ptr(p), _ptr(v), _ptr(e))
        return v, e

    def convolve(self, img):
        a = np.ascontiguousarray(img, dtype=self.real)
        out = np.zeros_like(a)
        if self.L.orc_convolve(self.h, _ptr(a), _ptr(out)):
            raise ValueError("model has no PSF")
        return out

    def dumper_layers(self, params):
        """The six result layers of the reference's dumper (src/nested.c:219-253), float32."""
        p = _f32(params)
        assert p.size == self.npars
        names = ("IMG", "RES", "RAW", "ERR", "WHT", "PVL")
        out = {n: np.zeros((self.height, self.width), np.float32) for n in names}
        self.L.orc_dumper_layers(self.h, _ptr(p), *[_ptr(out[n]) for n in names])
        return out

    def loglike(self, params, want_maps=False):
        p = _f32(params)
        assert p.size == self.npars
        ln = C.c_double()
        if want_maps:
            model = np.zeros((self.height, self.width), self.real)
            chi = np.zeros_like(model)
            self.L.orc_loglike(self.h, _ptr(p), C.byref(ln), _ptr(model), _ptr(chi))
            return ln.value, model, chi
        self.L.orc_loglike(self.h, _ptr(p), C.byref(ln), None, None)
        return ln.value
