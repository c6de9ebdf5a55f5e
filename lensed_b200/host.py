"""Host side around the device path: what a sampler driver needs to run a
Lensed configuration file against the CUDA layer.

Mirrors, in Python, the small amount of host logic that sits on either side of
the hot path in the reference (SURVEY.md section 8f "next" rows):

* priors ``delta``/``unif``/``norm``           src/prior.c:70-160, src/prior/*.c
* ini reader for [options]/[objects]/[priors]  src/input/ini.c:109-284,
  incl. the ``wrap`` / ``image`` keywords       src/input/objects.c:328-359
* parameter list, default bounds, parameter map (free dimensions first,
  derived ones last)                            src/lensed.c:107-271
* unit cube -> physical -> float32 params       src/nested.c:43-74
* dumper layers IMG RES RAW ERR WHT PVL         src/nested.c:219-253

Nothing here is on the timed path (O(npar) doubles per evaluation).
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import fits
from .api import Context, Model

POSITION_X, POSITION_Y, RADIUS, MAGNITUDE, AXIS_RATIO, POS_ANGLE = 1, 2, 3, 4, 5, 6


# ---------------------------------------------------------------------------
# priors
# ---------------------------------------------------------------------------
class Prior:
    pseudo = False

    def apply(self, u: float) -> float:
        raise NotImplementedError

    def apply_many(self, u):
        """apply() for an array of unit-interval values (batched samplers)."""
        return np.array([self.apply(float(v)) for v in np.asarray(u).ravel()], np.float64).reshape(np.shape(u))

    def lower(self) -> float:
        raise NotImplementedError

    def upper(self) -> float:
        raise NotImplementedError


class Delta(Prior):
    """Fixed value = pseudo-prior: the parameter is derived, not sampled
    (src/prior/delta.c:50-63, src/prior.c PRIORS[0].pseudo)."""
    pseudo = True

    def __init__(self, value: float):
        self.value = float(value)

    def apply(self, u):
        return self.value

    def apply_many(self, u):
        return np.full(np.shape(u), self.value, np.float64)

    def lower(self):
        return self.value

    def upper(self):
        return self.value


class Uniform(Prior):
    def __init__(self, a: float, b: float):
        self.a, self.b = float(a), float(b)

    def apply(self, u):
        return self.a + u*(self.b - self.a)          # src/prior/unif.c:68-73

    def apply_many(self, u):
        return self.a + np.asarray(u, np.float64)*(self.b - self.a)

    def lower(self):
        return self.a

    def upper(self):
        return self.b


class Normal(Prior):
    def __init__(self, mean: float, sigma: float):
        self.m, self.s = float(mean), float(sigma)

    @staticmethod
    def _gauss(u: float) -> float:
        # rational approximation of the normal quantile (Abramowitz & Stegun
        # 26.2.23), as src/prior/norm.c:9-15
        t = math.sqrt(-2.0*math.log(u)) if u < 0.5 else math.sqrt(-2.0*math.log(1 - u))
        t = t - ((0.010328*t + 0.802853)*t + 2.515517)/(((0.001308*t + 0.189269)*t + 1.432788)*t + 1.0)
        return -t if u < 0.5 else t

    def apply(self, u):
        return self.m + self.s*self._gauss(u)         # src/prior/norm.c:77-82

    def apply_many(self, u):
        # the same rational approximation, element-wise
        u = np.asarray(u, np.float64)
        low = u < 0.5
        t = np.sqrt(-2.0*np.log(np.where(low, u, 1 - u)))
        t = t - ((0.010328*t + 0.802853)*t + 2.515517)/(((0.001308*t + 0.189269)*t + 1.432788)*t + 1.0)
        return self.m + self.s*np.where(low, -t, t)

    def lower(self):
        return self.m - 7*self.s

    def upper(self):
        return self.m + 7*self.s


def read_prior(spec: str) -> Prior:
    """``"1.5"`` -> delta, ``"unif a b"``, ``"norm m s"`` (src/prior.c:70-160)."""
    args = spec.split()
    if len(args) == 1:
        try:
            return Delta(float(args[0]))
        except ValueError:
            pass
    if not args:
        raise ValueError("empty prior")
    # named priors: the list is searched from its second entry, so that "delta"
    # is reachable only as a bare number (src/prior.c:119-133)
    kinds = {"unif": Uniform, "norm": Normal}
    if args[0] not in kinds:
        raise ValueError(f"unknown prior: {args[0]}")
    try:
        if len(args) != 3:
            raise ValueError
        return kinds[args[0]](float(args[1]), float(args[2]))
    except ValueError:
        raise ValueError(f"invalid prior definition: {spec}") from None


# ---------------------------------------------------------------------------
# configuration
# ---------------------------------------------------------------------------
@dataclass
class Parameter:
    id: str                 # "<object id>.<param name>"
    name: str
    type: int
    lower: float
    upper: float
    prior: Optional[Prior] = None
    wrap: bool = False
    ipp: bool = False
    label: Optional[str] = None

    @property
    def bounded(self) -> bool:
        return bool(self.lower or self.upper)

    @property
    def derived(self) -> bool:
        return self.prior is not None and self.prior.pseudo


@dataclass
class ObjectEntry:
    id: str
    name: str
    type: str
    params: list = field(default_factory=list)


@dataclass
class Config:
    options: dict
    objects: list
    basedir: str = "."

    @property
    def parameters(self) -> list:
        return [p for o in self.objects for p in o.params]


def read_ini(path: str, ctx: Context) -> Config:
    """Parse a Lensed ini file (groups [options] (default) / [objects] /
    [priors] / [labels]).  Object metadata comes from the compiled object
    files, as src/input/objects.c:15-266 does."""
    options, objects, by_id = {}, [], {}
    grp = "options"
    nplanes, typ = 0, None
    with open(path) as f:
        for lineno, raw in enumerate(f, 1):
            # src/input/ini.c:16-25,152-218: a line ends at ';' or '#', a group is
            # [name], name and value are split at the first '=' or ':'
            line = re.split(r"[;#]", raw, maxsplit=1)[0].strip()
            if not line:
                continue
            if line.startswith("["):
                if not line.endswith("]"):
                    raise ValueError(f"{path}:{lineno}: missing closing ']' character for group")
                grp = line[1:-1].strip()
                if grp not in ("options", "objects", "priors", "labels"):
                    raise ValueError(f"{path}:{lineno}: unknown group: {grp}")
                continue
            parts = re.split(r"[=:]", line, maxsplit=1)
            if len(parts) != 2:
                raise ValueError(f"{path}:{lineno}: line does not assign anything to \"{line}\"")
            name, value = (s.strip() for s in parts)
            if grp == "options":
                options[name] = value
            elif grp == "objects":
                if name in by_id:
                    raise ValueError(f"{path}:{lineno}: duplicate object name: {name}")
                info = ctx.object_info(value)
                obj = ObjectEntry(name, value, info.type)
                for p in info.params:
                    par = Parameter(f"{name}.{p.name}", p.name, p.type, p.bounds[0], p.bounds[1])
                    if p.has_default:
                        par.prior = Delta(p.defval)       # src/input/objects.c:225-226
                    obj.params.append(par)
                # src/input/ini.c:249-260
                if info.type != typ and info.type != "F":
                    if info.type == "L":
                        nplanes += 1
                        if nplanes > 1:
                            raise ValueError(f"{path}:{lineno}: multiple lensing planes are not supported")
                    typ = info.type
                objects.append(obj)
                by_id[name] = obj
            elif grp in ("priors", "labels"):
                if "." not in name:
                    raise ValueError(f"{path}:{lineno}: object {name}: no parameter given (should be {name}.<param>)")
                oid, pname = name.split(".", 1)
                if oid not in by_id:
                    raise ValueError(f"{path}:{lineno}: unknown object: {oid} (check [objects] group)")
                par = next((p for p in by_id[oid].params if p.name == pname), None)
                if par is None:
                    raise ValueError(f"{path}:{lineno}: object {oid}: unknown parameter {pname}")
                if grp == "labels":
                    par.label = value
                    continue
                # keywords, src/input/objects.c:336-353
                words = value.split()
                while words and words[0] in ("wrap", "image"):
                    if words[0] == "wrap":
                        par.wrap = True
                    else:
                        par.ipp = True
                    words.pop(0)
                par.prior = read_prior(" ".join(words)) if words else None
            else:
                raise ValueError(f"{path}:{lineno}: unknown group [{grp}]")
    return Config(options, objects, os.path.dirname(os.path.abspath(path)))


# ---------------------------------------------------------------------------
# the likelihood as the sampler sees it
# ---------------------------------------------------------------------------
class Likelihood:
    """Parameter bookkeeping of src/lensed.c:107-271 plus loglike() of
    src/nested.c:17-131 on top of a device ``Model``."""

    def __init__(self, cfg: Config, model: Model):
        self.cfg = cfg
        self.model = model
        self.pars = cfg.parameters
        for p in self.pars:
            if p.prior is None:
                raise ValueError(f"missing prior: {p.id}")
            # default bounds, src/lensed.c:148-167
            if not p.lower and not p.upper:
                if p.type == RADIUS:
                    p.lower, p.upper = 0.0, math.inf
                elif p.type == AXIS_RATIO:
                    p.lower, p.upper = 0.0, 1.0
            # src/lensed.c:172-190: only a prior reaching outside the bounds is
            # examined; no overlap at all is an error, partial overlap a warning
            if p.bounded and (p.prior.lower() < p.lower or p.prior.upper() > p.upper):
                if p.prior.lower() >= p.upper or p.prior.upper() <= p.lower:
                    raise ValueError(f"{p.id}: prior does not include parameter bounds [{p.lower:g}, {p.upper:g}]")
        free = [i for i, p in enumerate(self.pars) if not p.derived]
        derived = [i for i, p in enumerate(self.pars) if p.derived]
        self.pmap = free + derived           # MultiNest keeps derived parameters last
        self.ndims = len(free)
        self.npars = len(self.pars)

    def physical(self, cube) -> np.ndarray:
        """Unit cube (first ndims entries used) -> physical parameters in
        sampler order (src/nested.c:43-61)."""
        out = np.empty(self.npars, np.float64)
        for i in range(self.npars):
            par = self.pars[self.pmap[i]]
            u = float(cube[i]) if i < self.ndims else 0.5
            phys = par.prior.apply(u)
            if par.bounded and (phys < par.lower or phys > par.upper):
                # the reference redraws with the same u, i.e. spins forever;
                # report instead
                raise ValueError(f"{par.id}: value {phys:g} outside parameter bounds [{par.lower:g}, {par.upper:g}]")
            out[i] = phys
        return out

    def device_params(self, phys) -> np.ndarray:
        """Sampler order -> object order, narrowed to float32 (src/nested.c:70-72)."""
        params = np.empty(self.npars, np.float32)
        for i in range(self.npars):
            params[self.pmap[i]] = phys[i]
        return params

    def physical_batch(self, cubes) -> np.ndarray:
        """physical() for [n][ndims] unit-cube points at once -> [n][npars],
        sampler order.  Points outside a parameter's hard bounds raise, as
        physical() does."""
        cubes = np.atleast_2d(np.asarray(cubes, np.float64))
        out = np.empty((cubes.shape[0], self.npars), np.float64)
        for i in range(self.npars):
            par = self.pars[self.pmap[i]]
            u = cubes[:, i] if i < self.ndims else np.full(cubes.shape[0], 0.5)
            phys = par.prior.apply_many(u)
            if par.bounded and ((phys < par.lower) | (phys > par.upper)).any():
                bad = phys[(phys < par.lower) | (phys > par.upper)][0]
                raise ValueError(f"{par.id}: value {bad:g} outside parameter bounds [{par.lower:g}, {par.upper:g}]")
            out[:, i] = phys
        return out

    def device_params_batch(self, phys) -> np.ndarray:
        """[n][npars] sampler order -> object order, float32."""
        phys = np.atleast_2d(phys)
        params = np.empty(phys.shape, np.float32)
        params[:, self.pmap] = phys
        return params

    def __call__(self, cube) -> float:
        return self.model.loglike(self.device_params(self.physical(cube)))

    def batch(self, cubes) -> np.ndarray:
        """Many points per launch: the batched entry point."""
        return self.model.loglike_batch(self.device_params_batch(self.physical_batch(cubes)))


def build(path: str, ctx: Context, **model_kw):
    """ini file -> (Config, Model, Likelihood): reads the image (with section),
    weight or gain/offset, mask-free, PSF (normalised as src/data.c:354-370)."""
    from . import workloads
    cfg = read_ini(path, ctx)
    opt = cfg.options
    # src/input.c:230-238: `image` is required, `gain` unless a weight map is given, and some objects
    for name in ("image",) + (() if "weight" in opt else ("gain",)):
        if name not in opt:
            raise ValueError(f"missing required option: {name}")
    if not cfg.objects:
        raise ValueError("no objects were given (check [objects] section)")

    def rel(p):
        return p if os.path.isabs(p) else os.path.join(cfg.basedir, p)

    image, pcs = fits.read_image(rel(opt["image"]))
    if "bscale" in opt:
        # image[i] *= bscale with a double bscale (src/lensed.c:430): widen, multiply, narrow once
        image = (image.astype(np.float64)*float(opt["bscale"])).astype(np.float32)

    def value_or_file(v):
        try:
            return np.full(image.shape, float(v), np.float32)
        except ValueError:
            arr, _ = fits.read_image(rel(v))
            if arr.shape != image.shape:
                raise ValueError(f"{v}: wrong dimensions {arr.shape[1]} x {arr.shape[0]}")
            return arr

    if "weight" in opt:
        weight = value_or_file(opt["weight"])
    else:
        gain = value_or_file(opt["gain"])
        # make_weight(), src/data.c:314-330
        weight = (gain.astype(np.float64)/(image.astype(np.float64) + float(opt.get("offset", 0)))).astype(np.float32)
    if "xweight" in opt:
        weight = weight*value_or_file(opt["xweight"])
    if "mask" in opt:
        # FITS mask: non-zero pixels are excluded by zero weight (src/lensed.c:471-486)
        mask, _ = fits.read_image(rel(opt["mask"]))
        if mask.shape != image.shape:
            raise ValueError(f"wrong dimensions {mask.shape[1]} x {mask.shape[0]} for mask")
        weight = np.where(mask != 0, np.float32(0), weight).astype(np.float32)
    psf = None
    if "psf" in opt:
        psf, _ = fits.read_image(rel(opt["psf"]))
        psf = workloads.normalise_psf(psf)
    ipp = [[int(p.ipp) for p in o.params] for o in cfg.objects]
    model = Model(ctx, [o.name for o in cfg.objects], image, weight, rule=opt.get("rule", "g3k7"), psf=psf, pcs=pcs,
                  ipp=ipp, **model_kw)
    return cfg, model, Likelihood(cfg, model)


def find_mode(values, mask=None, bins: int = 100):
    """Mode and FWHM of the pixel-value histogram, the reference's background
    check (src/data.c:421-501; src/lensed.c:519-547 warns when |mode| exceeds
    half the FWHM and there is no sky object)."""
    v = np.asarray(values, np.float64).ravel()
    if mask is not None:
        v = v[np.asarray(mask).ravel() != 0]
    if v.size == 0:
        return float("nan"), 0.0
    lo, hi = v.min(), v.max()
    if lo == hi:
        return float(lo), 0.0
    dx = (hi - lo)/bins
    j = np.minimum(((v - lo)/dx).astype(np.int64), bins - 1)
    counts = np.bincount(j, minlength=bins)
    peak = int(np.argmax(counts))                 # first of equal maxima, as the reference
    mode = lo + (peak + 0.5)*dx
    i = peak + 1
    while i < bins and counts[i] >= 0.5*counts[peak]:
        i += 1
    fwhm = i*dx
    i = peak
    while i > 0 and counts[i - 1] >= 0.5*counts[peak]:
        i -= 1
    return float(mode), float(fwhm - (i - 1)*dx)


# ---------------------------------------------------------------------------
# dumper
# ---------------------------------------------------------------------------
def layers_from_maps(model_img, raw, error, chi, image, weight) -> dict:
    """The host-side arithmetic of the dumper (src/nested.c:219-253) on the four
    maps it reads back from the device: RES = image - model (float), ERR =
    error/value (float division; 0/0 and x/0 are NaN / inf as in the
    reference), PVL = erfc(sqrt(0.5 chi^2)) evaluated in double and narrowed."""
    img = np.asarray(model_img, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        relerr = (np.asarray(error, np.float32)/np.asarray(raw, np.float32)).astype(np.float32)
    half = 0.5*np.asarray(chi, np.float32).astype(np.float64)
    pvl = np.array([math.erfc(math.sqrt(c)) if c >= 0 else math.nan for c in half.ravel().tolist()], np.float32).reshape(img.shape)
    return {"IMG": img, "RES": (np.asarray(image, np.float32) - img).astype(np.float32), "RAW": np.asarray(raw, np.float32),
            "ERR": relerr, "WHT": np.asarray(weight, np.float32), "PVL": pvl}


def dumper_layers(model: Model, params, image, weight) -> dict:
    """The six result layers of src/nested.c:178-253 for one parameter point:
    re-render on the device (lcu_render), layer arithmetic on the host as in the
    reference.  (The reference's PVL layer shows the chi^2 map of the last point
    MultiNest evaluated -- its dumper does not re-run the loglike kernel; here it
    is the map of the point given.)"""
    out = model.render(params)
    return layers_from_maps(out["model"], out["raw"], out["error"], out["chi"], image, weight)


def write_results(path: str, layers: dict):
    """Results file: primary HDU + IMAGE extensions named like the reference's
    (src/data.c:129-165, 372-391)."""
    names = ["IMG", "RES", "RAW", "ERR", "WHT", "PVL"]
    fits.write_layers(path, [layers[n] for n in names], names)
