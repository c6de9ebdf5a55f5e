"""Host-side mirror of the reference's device interface, on top of the C ABI.

``Context``  ~ get_lensed_cl()            (reference src/opencl.c:132-241)
``Context.object_info``  ~ add_object()'s device round trip (src/input/objects.c:72-239)
``quad_rule``  ~ quad_rule()              (src/quadrature.c:32-43)
``Model``    ~ the kernel set-up of src/lensed.c:644-1112
``Model.loglike``  ~ loglike()            (src/nested.c:63-115), plus the batched entry
``Model.render``   ~ the dumper's re-render (src/nested.c:178-253)

Arrays are numpy on the host; ``loglike_batch_device`` takes raw device
pointers (e.g. ``torch.Tensor.data_ptr()``) and a CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import LensedCudaError, check, lib

LENS, SOURCE, FOREGROUND = "L", "S", "F"
PARAM_TYPES = ("PARAMETER", "POSITION_X", "POSITION_Y", "RADIUS", "MAGNITUDE", "AXIS_RATIO", "POS_ANGLE")


@dataclass
class Param:
    name: str
    type: int
    bounds: tuple
    defval: float
    has_default: bool      # defval > 0 or sign bit set (src/input/objects.c:225)


@dataclass
class ObjectInfo:
    name: str
    type: str
    words: int
    params: list

    @property
    def npars(self) -> int:
        return len(self.params)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def quad_rules():
    """Names and descriptions of the built-in rules (``lensed --rules``)."""
    return [(lib.lcu_quad_rule_name(i).decode(), lib.lcu_quad_rule_info(i).decode())
            for i in range(lib.lcu_quad_rule_count())]


def quad_rule(rule: str, sx: float = 1.0, sy: float = 1.0):
    """(qq[n,2], ww[n,2]) float32: scaled abscissae and (weight, error weight)."""
    n = lib.lcu_quad_rule(rule.encode(), sx, sy, None, None)
    if n < 0:
        raise ValueError(f"invalid quadrature rule: {rule}")
    qq = np.zeros((n, 2), np.float32)
    ww = np.zeros((n, 2), np.float32)
    lib.lcu_quad_rule(rule.encode(), sx, sy, _ptr(qq), _ptr(ww))
    return qq, ww


def launch_count() -> int:
    return int(lib.lcu_launch_count())


class Context:
    """A CUDA device (``device >= 0``) or a compile-only context (``device = -1``)."""

    def __init__(self, device: int = 0, objects_dir: Optional[str] = None, kernel_dir: Optional[str] = None):
        h = C.c_void_p()
        check(lib.lcu_create(int(device), kernel_dir.encode() if kernel_dir else None,
                             objects_dir.encode() if objects_dir else None, C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            lib.lcu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def object_info(self, name: str) -> ObjectInfo:
        t, w, n = C.c_int(), C.c_size_t(), C.c_size_t()
        pars = (_lib.LcuParam * 64)()
        check(lib.lcu_object_info(self._h, name.encode(), C.byref(t), C.byref(w), C.byref(n), pars, 64))
        out = []
        for i in range(n.value):
            p = pars[i]
            d = float(p.defval)
            out.append(Param(p.name.decode(), int(p.type), (float(p.bounds[0]), float(p.bounds[1])), d,
                             d > 0 or bool(np.signbit(np.float32(d)))))
        return ObjectInfo(name, chr(t.value), int(w.value), out)

    def object_pairable(self, name: str):
        """(True, "") if the object's per-ray code compiles for two rays per
        thread, else (False, compiler log)."""
        why = C.c_char_p()
        rc = lib.lcu_object_pairable(self._h, name.encode(), C.byref(why))
        if rc < 0:
            check(-rc)
        return bool(rc), (why.value or b"").decode()

    def fp32_peak_tflops(self) -> float:
        v = C.c_double()
        check(lib.lcu_measure_fp32_peak(self._h, C.byref(v)))
        return v.value


class Model:
    """One lens model on one device.

    objects : object file names in ini order (``[objects]`` group), e.g.
              ``["sie", "sersic"]``
    ipp     : optional list (one entry per object) of per-parameter
              image-plane-prior flags (the ``image`` keyword of ``[priors]``)
    """

    def __init__(self, ctx: Context, objects: Sequence[str], image, weight, rule: str = "g3k7",
                 psf=None, pcs=(1.0, 1.0, 1.0, 1.0), ipp=None, qq=None, ww=None,
                 max_batch: int = 0, flags: int = 0):
        self.ctx = ctx
        self.objects = list(objects)
        image = _f32(image)
        if image.ndim != 2:
            raise ValueError("image must be 2-D")
        weight = _f32(weight)
        if weight.shape != image.shape:
            raise ValueError("wrong dimensions for weight map")
        self.height, self.width = image.shape
        if qq is None or ww is None:
            qq, ww = quad_rule(rule, pcs[2], pcs[3])
        qq, ww = _f32(qq), _f32(ww)
        psf_a = _f32(psf) if psf is not None else None

        specs = (_lib.LcuObjectSpec * max(len(self.objects), 1))()
        keep = []
        for i, name in enumerate(self.objects):
            specs[i].name = name.encode()
            if ipp is not None and ipp[i] is not None and any(ipp[i]):
                arr = (C.c_int * len(ipp[i]))(*[int(bool(v)) for v in ipp[i]])
                keep.append(arr)
                specs[i].ipp = arr
        desc = _lib.LcuModelDesc()
        desc.width, desc.height = self.width, self.height
        desc.pcs = (C.c_float * 4)(*[float(v) for v in pcs])
        desc.nq = qq.shape[0]
        desc.qq, desc.ww = _ptr(qq), _ptr(ww)
        desc.image, desc.weight = _ptr(image), _ptr(weight)
        if psf_a is not None:
            desc.psf = _ptr(psf_a)
            desc.psf_height, desc.psf_width = psf_a.shape
        desc.max_batch = int(max_batch)
        desc.flags = int(flags)
        h = C.c_void_p()
        check(lib.lcu_model_create(ctx._h, specs, len(self.objects), C.byref(desc), C.byref(h)))
        self._h = h
        self.npars = int(lib.lcu_model_npars(h))
        self.words = int(lib.lcu_model_words(h))
        self.max_batch = int(lib.lcu_model_max_batch(h))
        self.nq = int(qq.shape[0])
        self.has_psf = psf_a is not None

    def close(self):
        if getattr(self, "_h", None):
            lib.lcu_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rays_per_thread(self) -> int:
        return int(lib.lcu_model_rays_per_thread(self._h))

    @property
    def source(self) -> str:
        return lib.lcu_model_source(self._h).decode()

    @property
    def build_log(self) -> str:
        return lib.lcu_model_build_log(self._h).decode()

    @property
    def cubin(self) -> bytes:
        img = C.c_void_p()
        n = lib.lcu_model_cubin(self._h, C.byref(img))
        return C.string_at(img, n)

    def kernel_usage(self, kernel: str):
        """(registers per thread, stack bytes) of a kernel of the compiled module."""
        r, st = C.c_uint(), C.c_uint()
        check(lib.lcu_model_kernel_usage(self._h, kernel.encode(), C.byref(r), C.byref(st)))
        return int(r.value), int(st.value)

    def set_rows(self, row0: int, row1: int):
        check(lib.lcu_model_set_rows(self._h, int(row0), int(row1)))

    def set_data(self, image=None, weight=None):
        """Replace the observed image and / or the weight map (same size)."""
        for a in (image, weight):
            if a is not None and np.asarray(a).shape != (self.height, self.width):
                raise ValueError("set_data: array shape differs from the model's image")
        im = _f32(image) if image is not None else None
        wt = _f32(weight) if weight is not None else None
        check(lib.lcu_model_set_data(self._h, _ptr(im), _ptr(wt)))

    def make_weight(self, gain, offset: float = 0.0, mask=None) -> np.ndarray:
        """weight = gain / (image + offset) on the device (src/data.c:314-330);
        gain is a number or a per-pixel map, masked pixels (non-zero) get
        weight 0 (src/lensed.c:470-482).  Returns the new map."""
        gm = None
        g0 = 0.0
        if np.ndim(gain) == 0:
            g0 = float(gain)
        else:
            gm = _f32(gain)
            if gm.shape != (self.height, self.width):
                raise ValueError("make_weight: gain map shape differs from the model's image")
        mk = None
        if mask is not None:
            mk = np.ascontiguousarray(mask, dtype=np.int32)
            if mk.shape != (self.height, self.width):
                raise ValueError("make_weight: mask shape differs from the model's image")
        check(lib.lcu_model_make_weight(self._h, _ptr(gm), C.c_float(g0), C.c_double(offset), _ptr(mk)))
        return self.weight_map()

    def weight_map(self) -> np.ndarray:
        out = np.empty((self.height, self.width), np.float32)
        check(lib.lcu_model_get_weight(self._h, _ptr(out)))
        return out

    def _params(self, params, batch: bool):
        p = _f32(params)
        if batch:
            if p.ndim != 2 or p.shape[1] != self.npars:
                raise ValueError(f"params must be [B, {self.npars}]")
        elif p.size != self.npars:
            raise ValueError(f"params must have {self.npars} entries")
        return p

    def loglike(self, params) -> float:
        p = self._params(params, False)
        v = C.c_double()
        check(lib.lcu_loglike(self._h, _ptr(p), C.byref(v)))
        return v.value

    def loglike_async(self, params) -> int:
        """Start one evaluation and return its ticket (at most two in flight)."""
        p = self._params(params, False)
        t = C.c_int()
        check(lib.lcu_loglike_async(self._h, _ptr(p), C.byref(t)))
        return t.value

    def loglike_wait(self, ticket: int) -> float:
        v = C.c_double()
        check(lib.lcu_loglike_wait(self._h, int(ticket), C.byref(v)))
        return v.value

    def loglike_batch(self, params) -> np.ndarray:
        p = self._params(params, True)
        out = np.zeros(p.shape[0], np.float64)
        check(lib.lcu_loglike_batch(self._h, p.shape[0], _ptr(p), _ptr(out)))
        return out

    def loglike_batch_device(self, nbatch: int, d_params: int, d_lnew: int, stream: int = 0):
        """Enqueue on ``stream`` (raw cudaStream_t; 0 = the CUDA default stream) with
        device-resident ``params[nbatch, npars]`` float32 / ``lnew[nbatch]`` float64."""
        check(lib.lcu_loglike_batch_device(self._h, int(nbatch), C.c_void_p(d_params), C.c_void_p(d_lnew),
                                           C.c_void_p(stream) if stream else None))

    def render(self, params, model=True, raw=True, error=True, chi=True) -> dict:
        p = self._params(params, False)
        shape = (self.height, self.width)
        bufs = {k: (np.zeros(shape, np.float32) if want else None)
                for k, want in (("model", model), ("raw", raw), ("error", error), ("chi", chi))}
        check(lib.lcu_render(self._h, _ptr(p), _ptr(bufs["model"]), _ptr(bufs["raw"]), _ptr(bufs["error"]), _ptr(bufs["chi"])))
        return {k: v for k, v in bufs.items() if v is not None}

    def set_params(self, params) -> np.ndarray:
        p = self._params(params, False)
        block = np.zeros(self.words, np.uint32)
        check(lib.lcu_set_params(self._h, _ptr(p), _ptr(block)))
        return block

    def profile(self, on: bool = True):
        check(lib.lcu_profile_enable(self._h, int(on)))

    def profile_get(self) -> dict:
        pr = _lib.LcuProfile()
        check(lib.lcu_profile_get(self._h, C.byref(pr)))
        return {f: getattr(pr, f) for f, _ in pr._fields_}
