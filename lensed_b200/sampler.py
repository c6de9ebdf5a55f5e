"""Batch-native nested sampling on top of the batched likelihood entry point.

The reference hands the evaluation of one parameter point at a time to
MultiNest (``run()`` call in src/lensed.c:1236-1288, callback ``loglike`` in
src/nested.c:17-131); MultiNest (a third-party Fortran library, "version 3.8
or later", docs/dependencies.md:6, absent here) draws one point, waits for its
likelihood, draws the next.  A GPU wants the opposite: ``lcu_loglike_batch``
evaluates B points per launch (include/lensed_cuda.h).  This module is the
driver that produces such batches: the nested
sampling algorithm of Skilling (2004) with MultiNest's ellipsoidal rejection
scheme (Feroz & Hobson 2008: the live points are partitioned recursively by
2-means into ellipsoids while that shrinks the bounded volume), restated so that the B candidate
points of one step are drawn *before* any of them is evaluated:

    bound   = union of enlarged ellipsoids around clusters of the live points
              (intersected with the unit cube)
    cand    = B points uniform in bound                    -> one batched launch
    for c in cand, in the order drawn:
        if L(c) > L_min(live): the worst live point dies (weight L_min dX),
                               c takes its place, X shrinks by exp(-1/nlive)

When rejection from the bound stops paying (curved posteriors in >~ 10
dimensions: acceptance ~1e-3), the candidates of a round come from constrained
random walks instead -- copies of live points take Metropolis steps inside the
current contour, all walkers advancing by one step per launch -- and are played
against the live set in the same way.

Every candidate is an independent uniform draw from a region that contains the
whole iso-likelihood contour of every threshold met during the step (contours
only shrink), so a candidate accepted against the threshold current at its turn
is a uniform draw from inside that contour -- the property nested sampling
needs.  Nothing depends on the evaluation order inside the launch.

Conventions follow MultiNest where Lensed exposes them
(src/input/options.c:148-232, passed on in src/lensed.c:1248-1275): ``nlive``
(default 300), ``tol`` (tolerance in log-evidence, 0.1), ``eff`` (MultiNest's
efr, from ``shf`` = 0.8: the bound's volume is enlarged by 1/eff), ``seed``,
``maxiter`` (0 = no limit).
Outputs are in physical parameters, as the reference's callback overwrites the
cube with them (src/nested.c:43-61); ``write_multinest`` writes the
``<root>.txt`` / ``<root>post_equal_weights.dat`` / ``<root>stats.dat`` files in
MultiNest's column layout.

Host-side numpy only; all device work goes through the likelihood callable.
With one process per GPU every rank runs the same sampler with the same seed
and evaluates its slice of each batch (lensed_b200.distributed).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np


@dataclass
class NestedResult:
    logz: float                     # ln evidence
    logz_err: float                 # sqrt(H / nlive)
    information: float              # H, nats
    samples: np.ndarray             # [n][ndims] unit-cube positions, dead points then final live points
    physical: Optional[np.ndarray]  # [n][npars] physical parameters (if a transform was given)
    loglike: np.ndarray             # [n]
    logwt: np.ndarray               # [n] ln posterior weight (normalised: logsumexp = 0)
    niter: int                      # dead points
    nevals: int                     # likelihood evaluations
    nbatches: int                   # batched launches
    efficiency: float               # accepted / evaluated after the initial live points
    stats: dict = field(default_factory=dict)

    @property
    def weights(self) -> np.ndarray:
        return np.exp(self.logwt)

    def mean(self, physical: bool = True) -> np.ndarray:
        x = self.physical if physical and self.physical is not None else self.samples
        return self.weights @ x

    def std(self, physical: bool = True) -> np.ndarray:
        x = self.physical if physical and self.physical is not None else self.samples
        m = self.weights @ x
        return np.sqrt(np.maximum(self.weights @ (x - m)**2, 0.0))

    def max_like(self, physical: bool = True) -> np.ndarray:
        x = self.physical if physical and self.physical is not None else self.samples
        return x[int(np.argmax(self.loglike))]

    def equal_weights(self, rng=None) -> np.ndarray:
        """Indices of an equally weighted posterior sample (MultiNest's
        post_equal_weights): point i is kept with probability w_i / max w."""
        rng = np.random.default_rng(rng)
        w = self.weights
        return np.nonzero(rng.random(w.size) < w/w.max())[0]


def _logaddexp(a: float, b: float) -> float:
    if a == -math.inf:
        return b
    if b == -math.inf:
        return a
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


class _Ellipsoid:
    """{u : (u - c)^T A^-1 (u - c) <= 1} around a set of points, enlarged so that
    its volume is `enlarge` times that of the smallest similar ellipsoid
    containing them all."""

    def __init__(self, pts: np.ndarray, enlarge: float, min_logvol: float = -math.inf):
        n, d = pts.shape
        self.d = d
        self.c = pts.mean(axis=0)
        dx = pts - self.c
        cov = dx.T @ dx/max(n - 1, 1)
        # regularise degenerate directions (n <= d, or points on a hyperplane)
        w, v = np.linalg.eigh(cov)
        w = np.maximum(w, max(w.max(), 1e-300)*1e-12)
        cov = (v*w) @ v.T
        inv = (v/w) @ v.T
        k = float(np.einsum("ij,jk,ik->i", dx, inv, dx).max())       # largest Mahalanobis distance^2
        k = max(k, 1e-300)*enlarge**(2.0/d)
        self.L = np.linalg.cholesky(cov*k)                            # A = L L^T
        self.logvol = float(np.log(np.diag(self.L)).sum()) + _log_unit_ball(d)
        # never smaller than the prior volume its points stand for (Feroz et al.
        # 2009, section 5.1.1): an ellipsoid fitted to few points in many
        # dimensions would otherwise cut into the iso-likelihood contour
        if self.logvol < min_logvol:
            self.L = self.L*math.exp((min_logvol - self.logvol)/d)
            self.logvol = min_logvol

    def draw(self, n: int, rng) -> np.ndarray:
        z = rng.standard_normal((n, self.d))
        z /= np.linalg.norm(z, axis=1, keepdims=True)
        r = rng.random(n)**(1.0/self.d)
        return self.c + (z*r[:, None]) @ self.L.T

    def contains(self, u: np.ndarray) -> np.ndarray:
        y = np.linalg.solve(self.L, (u - self.c).T)
        return (y*y).sum(axis=0) <= 1.0


def _log_unit_ball(d: int) -> float:
    return 0.5*d*math.log(math.pi) - math.lgamma(0.5*d + 1.0)


def _kmeans2(pts: np.ndarray, rng, iters: int = 20):
    """Two-means split of the live points (MultiNest partitions the live set
    recursively this way); returns a boolean label array or None."""
    n = pts.shape[0]
    i = int(rng.integers(n))
    j = int(np.argmax(((pts - pts[i])**2).sum(axis=1)))
    c = np.stack([pts[i], pts[j]])
    lab = None
    for _ in range(iters):
        d2 = ((pts[:, None, :] - c[None, :, :])**2).sum(axis=2)
        new = d2[:, 1] < d2[:, 0]
        if lab is not None and np.array_equal(new, lab):
            break
        lab = new
        if lab.all() or not lab.any():
            return None
        c = np.stack([pts[~lab].mean(axis=0), pts[lab].mean(axis=0)])
    return lab


class _Bound:
    """Union of ellipsoids, intersected with the unit cube.  The live points are
    partitioned recursively by 2-means, as MultiNest does; a split is kept
    when the two enlarged ellipsoids together have clearly less volume than
    the one around all the points.  Uniform sampling from the union: pick an
    ellipsoid by volume, draw in it, and accept with probability
    1 / (number of ellipsoids containing the point)."""

    def __init__(self, pts: np.ndarray, enlarge: float, rng, split: bool, logx: float = 0.0, max_ellipsoids: int = 16):
        n, d = pts.shape
        self.d = d
        self.ells = []
        min_pts = 2*(d + 1)

        def ell(p):
            # expected prior volume of the region these points sample: X n_k / N
            return _Ellipsoid(p, enlarge, logx + math.log(p.shape[0]/n) + math.log(enlarge))

        def build(p, e):
            if split and len(self.ells) + 1 < max_ellipsoids and p.shape[0] >= 2*min_pts:
                lab = _kmeans2(p, rng)
                if lab is not None and min(int(lab.sum()), int((~lab).sum())) >= min_pts:
                    a, b = ell(p[~lab]), ell(p[lab])
                    if np.logaddexp(a.logvol, b.logvol) < e.logvol + math.log(0.5):
                        build(p[~lab], a)
                        build(p[lab], b)
                        return
            self.ells.append(e)

        build(pts, ell(pts))
        lv = np.array([e.logvol for e in self.ells])
        self.logvol = float(np.logaddexp.reduce(lv))
        self.p = np.exp(lv - self.logvol)

    def draw(self, n: int, rng, max_tries: int = 64) -> np.ndarray:
        """n points uniform in (union of ellipsoids) x [0,1)^d.  Raises
        RuntimeError if practically none of the union lies inside the cube."""
        out = np.empty((0, self.d))
        want = n
        tried = kept = 0
        for _ in range(max_tries):
            # size the next draw by the fraction that fell inside the cube so far
            frac = max(kept/tried, 1e-4) if tried else 0.5
            m = int(min(max(want/frac*1.2, 16), 1 << 18))
            if len(self.ells) == 1:
                u = self.ells[0].draw(m, rng)
            else:
                which = rng.choice(len(self.ells), size=m, p=self.p)
                u = np.empty((m, self.d))
                for k, e in enumerate(self.ells):
                    sel = which == k
                    if sel.any():
                        u[sel] = e.draw(int(sel.sum()), rng)
                cnt = np.zeros(m, int)
                for e in self.ells:
                    cnt += e.contains(u)
                u = u[rng.random(m)*np.maximum(cnt, 1) < 1.0]
            u = u[np.all((u >= 0.0) & (u < 1.0), axis=1)]
            tried += m
            kept += u.shape[0]
            out = np.concatenate([out, u[:want]])
            want = n - out.shape[0]
            if want <= 0:
                return out
        raise RuntimeError("nested sampling: the bounding ellipsoids lie almost entirely outside the unit cube")


def _box_draw(pts: np.ndarray, n: int, enlarge: float, rng) -> np.ndarray:
    """Fallback bound: the axis-aligned box around the live points, enlarged by
    the same volume factor and clipped to the unit cube."""
    d = pts.shape[1]
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    grow = 0.5*(hi - lo)*(enlarge**(1.0/d) - 1.0) + 1e-12
    lo, hi = np.maximum(lo - grow, 0.0), np.minimum(hi + grow, 1.0)
    return lo + rng.random((n, d))*(hi - lo)


def nested_sample(loglike_batch: Callable[[np.ndarray], np.ndarray], ndims: int, *, nlive: int = 300,
                  batch: int = 64, tol: float = 0.1, eff: float = 0.8, seed: int = 0, maxiter: int = 0,
                  transform: Optional[Callable[[np.ndarray], np.ndarray]] = None, split: bool = False,
                  method: str = "auto", walks: int = 0, wrap: Optional[Sequence[bool]] = None,
                  callback: Optional[Callable[[dict], None]] = None, update_interval: int = 0) -> NestedResult:
    """Nested sampling of a likelihood over the unit cube [0,1)^ndims.

    loglike_batch : callable([n][ndims] float64) -> [n] log-likelihoods; called
                    with ``batch`` points per step (``nlive`` points at start,
                    in chunks of ``batch``)
    transform     : optional unit cube -> physical parameters, applied to the
                    returned samples (one point per call)
    tol           : stop when the live points can raise ln Z by less than this
    eff           : target efficiency; the bound's volume is enlarged by 1/eff
    maxiter       : stop after this many dead points (0 = no limit)
    wrap          : per dimension, whether it is periodic (the `wrap` keyword of
                    a prior, handed to MultiNest as pWrap by src/lensed.c:1254-1259).
                    The circle of a periodic dimension is cut opposite the live
                    points before every step, so that a posterior straddling
                    0 / 1 (a position angle near 0 = 180 degrees) is bounded by one
                    small ellipsoid instead of one as wide as the cube; random-walk
                    steps wrap around.  The cut is a measure-preserving
                    re-parametrisation: draws stay uniform inside the contour.
    method        : how new points are drawn inside the likelihood contour.
                    "reject": uniform draws from the ellipsoidal bound (exact, but the
                    acceptance falls to 1e-3 and below for curved posteriors in >~ 10
                    dimensions); "rwalk": ``walks`` Metropolis steps from copies of live
                    points, constrained to the contour, proposals shaped by the live
                    points' covariance -- one launch per step for all walkers of a
                    round (Skilling 2006; cost ~ walks evaluations per dead point in
                    any dimension); "auto" (default): rejection while more than one
                    candidate in ``walks`` is accepted, random walks from then on
    walks         : Metropolis steps per new point for "rwalk"; 0 = max(25, 8 ndims).  Too few
                    steps leave new points correlated with their starting points and
                    bias ln Z upwards (12-D test with a known answer: +0.5 at 25 steps,
                    +0.14 +- 0.07 at 100, against a statistical error of 0.33); the
                    posterior itself is far less sensitive
    split         : partition the live points into several ellipsoids.  Pays for
                    separated modes and curved degeneracies in few dimensions;
                    with ~nlive/10 points per ellipsoid in >~ 10 dimensions the
                    fitted ellipsoids cut into the contour and bias ln Z upwards
                    (measured: +1.1 at 3 sigma on a 12-D test), hence off by default
    """
    if ndims < 1 or nlive < 2 or batch < 1:
        raise ValueError("nested_sample: need ndims >= 1, nlive >= 2, batch >= 1")
    if not 0.0 < eff <= 1.0:
        raise ValueError("nested_sample: eff must be in (0, 1]")
    if method not in ("auto", "reject", "rwalk") or walks < 0:
        raise ValueError("nested_sample: method must be 'auto', 'reject' or 'rwalk', walks >= 0")
    if walks == 0:
        walks = max(25, 8*ndims)
    rng = np.random.default_rng(seed)
    wrap_idx = np.nonzero(np.asarray(wrap, dtype=bool))[0] if wrap is not None else np.empty(0, dtype=int)
    if wrap_idx.size and wrap_idx.max() >= ndims:
        raise ValueError("nested_sample: wrap has more entries than there are dimensions")

    def cut() -> np.ndarray:
        """offset per dimension that moves the cut of every periodic dimension
        half a turn away from the circular mean of the live points"""
        off = np.zeros(ndims)
        ang = 2.0*math.pi*live_u[:, wrap_idx]
        mean = np.arctan2(np.sin(ang).mean(axis=0), np.cos(ang).mean(axis=0))/(2.0*math.pi)
        off[wrap_idx] = (mean - 0.5) % 1.0
        return off

    nan_points = 0

    def evaluate(u: np.ndarray) -> np.ndarray:
        """One batched launch; a NaN likelihood (a parameter point at which the
        model itself is undefined, e.g. axis ratio 0) counts as zero
        likelihood, as MultiNest's logzero does (src/lensed.c:1244)."""
        nonlocal nan_points
        ll = np.array(loglike_batch(u), dtype=np.float64)
        bad = np.isnan(ll)
        if bad.any():
            nan_points += int(bad.sum())
            ll[bad] = -math.inf
        return ll

    live_u = rng.random((nlive, ndims))
    live_l = np.empty(nlive)
    nevals = nbatches = 0
    for i in range(0, nlive, batch):
        live_l[i:i + batch] = evaluate(live_u[i:i + batch])
        nbatches += 1
    nevals += nlive
    if not np.isfinite(live_l).any():
        raise ValueError("nested_sample: the likelihood is NaN or zero at every initial live point")

    dead_u, dead_l, dead_lw = [], [], []
    logz = -math.inf
    h = 0.0
    logx = 0.0                                   # ln prior volume left
    # ln(X_{i-1} - X_i) for X_i = exp(-i/nlive)
    log_dx_factor = math.log1p(-math.exp(-1.0/nlive))
    accepted = proposed = 0
    niter = 0
    enlarge = 1.0/eff

    def add_weight(ll: float, logw: float):
        """Accumulate ln Z and the information H (Skilling 2006, eq. 14 ff.)."""
        nonlocal logz, h
        new = _logaddexp(logz, ll + logw)
        if new == -math.inf:
            return
        t_new = math.exp(ll + logw - new)*ll if ll > -math.inf else 0.0
        t_old = math.exp(logz - new)*(h + logz) if logz > -math.inf else 0.0
        h = t_new + t_old - new
        logz = new

    use_walk = method == "rwalk"
    recent = []                                  # acceptance of the last rejection steps
    step_scale = 1.0                             # random-walk step, in units of the live covariance
    walk_rounds = walk_accept = 0

    def walk_round():
        """One round of constrained random walks: `walkers` copies of live points
        take `walks` Metropolis steps inside {L > L_min at the start of the
        round}; every step is one launch over all walkers.  Returns the end
        points that moved (a walker that never moved is a copy of a live point
        and is dropped)."""
        nonlocal nevals, nbatches, step_scale, walk_accept
        nw = max(1, min(batch, nlive//2))
        lmin0 = live_l.min()
        start = rng.integers(nlive, size=nw)
        u, ll = live_u[start].copy(), live_l[start].copy()
        cov = np.cov((live_u - cut()) % 1.0 if wrap_idx.size else live_u, rowvar=False).reshape(ndims, ndims)
        w, v = np.linalg.eigh(cov)
        w = np.maximum(w, max(w.max(), 1e-300)*1e-12)
        chol = np.linalg.cholesky((v*w) @ v.T)
        moved = np.zeros(nw, bool)
        nacc = 0
        for _ in range(walks):
            prop = u + step_scale*(rng.standard_normal((nw, ndims)) @ chol.T)
            if wrap_idx.size:
                prop[:, wrap_idx] %= 1.0
            inside = np.all((prop >= 0.0) & (prop < 1.0), axis=1)
            if inside.any():
                lp = evaluate(prop[inside])
                nevals += int(inside.sum())
                nbatches += 1
                ok = np.zeros(nw, bool)
                ok[inside] = lp > lmin0
                full = np.empty(nw)
                full[inside] = lp
                u[ok], ll[ok] = prop[ok], full[ok]
                moved |= ok
                nacc += int(ok.sum())
        # step size towards ~40 % acceptance (the usual target for constrained walks)
        rate = nacc/(nw*walks)
        step_scale *= math.exp((rate - 0.4)/2.0)
        step_scale = min(max(step_scale, 1e-6), 4.0)
        walk_accept += nacc
        return u[moved], ll[moved], nw*walks

    done = False
    while not done:
        if use_walk:
            cand, cl, cost = walk_round()
            walk_rounds += 1
            proposed += cost
        else:
            # bound from the current live points; while it is no smaller than the
            # cube itself, draw from the cube
            off = cut() if wrap_idx.size else None
            pts = (live_u - off) % 1.0 if wrap_idx.size else live_u
            bound = _Bound(pts, enlarge, rng, split, logx)
            if bound.logvol >= 0.0:
                cand = rng.random((batch, ndims))
            else:
                try:
                    cand = bound.draw(batch, rng)
                except RuntimeError:
                    # posterior pressed into a corner of the prior in many dimensions
                    cand = _box_draw(pts, batch, enlarge, rng)
                if wrap_idx.size:
                    cand = (cand + off) % 1.0
                    cand[cand >= 1.0] = 0.0       # (x + off) % 1 can round up to 1
            cl = evaluate(cand)
            nevals += batch
            nbatches += 1
            proposed += batch
        before = accepted
        # the threshold only rises during the step: candidates at or below the
        # threshold it starts with can never be accepted
        for k in np.nonzero(cl > live_l.min())[0]:
            u, ll = cand[k], cl[k]
            worst = int(np.argmin(live_l))
            lmin = live_l[worst]
            if not ll > lmin:
                continue
            # the worst live point dies with weight L_min (X_{i-1} - X_i)
            logw = logx + log_dx_factor
            add_weight(lmin, logw)
            dead_u.append(live_u[worst].copy())
            dead_l.append(lmin)
            dead_lw.append(lmin + logw)
            live_u[worst] = u
            live_l[worst] = ll
            logx -= 1.0/nlive
            niter += 1
            accepted += 1
            # termination: what the live points could still add (MultiNest's tol)
            remain = live_l.max() + logx
            if logz > -math.inf and _logaddexp(logz, remain) - logz < tol:
                done = True
                break
            if maxiter and niter >= maxiter:
                done = True
                break
        if not use_walk and method == "auto":
            recent.append((accepted - before)/batch)
            recent = recent[-8:]
            if len(recent) == 8 and sum(recent)/8 < 1.0/walks:
                use_walk = True
        if callback is not None and (update_interval <= 0 or nbatches % update_interval == 0 or done):
            callback(dict(niter=niter, nevals=nevals, nbatches=nbatches, logz=logz, logx=logx,
                          lmax=float(live_l.max()), efficiency=accepted/max(proposed, 1)))

    # the live points share what is left of the prior volume
    logw_live = logx - math.log(nlive)
    order = np.argsort(live_l)
    for i in order:
        add_weight(live_l[i], logw_live)
        dead_u.append(live_u[i].copy())
        dead_l.append(live_l[i])
        dead_lw.append(live_l[i] + logw_live)

    samples = np.array(dead_u)
    ll = np.array(dead_l)
    logwt = np.array(dead_lw) - logz
    phys = np.stack([np.asarray(transform(s), dtype=np.float64) for s in samples]) if transform is not None else None
    return NestedResult(logz=logz, logz_err=math.sqrt(max(h, 0.0)/nlive), information=h, samples=samples, physical=phys,
                        loglike=ll, logwt=logwt, niter=niter, nevals=nevals, nbatches=nbatches,
                        efficiency=accepted/max(proposed, 1),
                        stats={"nan_points": nan_points, "walk_rounds": walk_rounds, "walk_step": step_scale,
                               "walk_acceptance": walk_accept/max(walk_rounds*walks*max(1, min(batch, nlive//2)), 1)})


def run(like, *, nlive: int = 300, batch: int = 64, tol: float = 0.1, eff: float = 0.8, seed: int = 0,
        maxiter: int = 0, evaluate: Optional[Callable[[np.ndarray], np.ndarray]] = None, **kw) -> NestedResult:
    """Sample a ``lensed_b200.host.Likelihood``: priors and the parameter map
    turn each unit-cube point into the float32 parameter vector of the model
    (src/nested.c:43-74), the whole batch goes through ``loglike_batch`` in one
    call.  ``evaluate`` replaces the model call, e.g. by
    ``ShardedLikelihood.for_model(model).loglike_batch`` with one process per
    GPU (every rank then runs this function with the same seed)."""
    ev = evaluate if evaluate is not None else like.model.loglike_batch

    def lb(cubes: np.ndarray) -> np.ndarray:
        return np.asarray(ev(like.device_params_batch(like.physical_batch(cubes))), dtype=np.float64)

    # periodic parameters, sampler order (src/lensed.c:1254-1259)
    kw.setdefault("wrap", [bool(getattr(like.pars[like.pmap[i]], "wrap", False)) for i in range(like.ndims)])
    return nested_sample(lb, like.ndims, nlive=nlive, batch=batch, tol=tol, eff=eff, seed=seed, maxiter=maxiter,
                         transform=like.physical, **kw)


def write_multinest(root: str, res: NestedResult, labels: Optional[Sequence[str]] = None, seed: int = 0):
    """MultiNest-layout result files (what the reference leaves behind for its
    users' tools): ``<root>.txt`` = weight, -2 ln L, parameters;
    ``<root>post_equal_weights.dat`` = parameters, ln L;
    ``<root>stats.dat`` = evidence, mean / sigma, maximum-likelihood point."""
    x = res.physical if res.physical is not None else res.samples
    w = res.weights
    with open(root + ".txt", "w") as f:
        for wi, li, xi in zip(w, res.loglike, x):
            f.write(" ".join(f"{v: .18E}" for v in (wi, -2.0*li, *xi)) + "\n")
    idx = res.equal_weights(seed)
    with open(root + "post_equal_weights.dat", "w") as f:
        for i in idx:
            f.write(" ".join(f"{v: .18E}" for v in (*x[i], res.loglike[i])) + "\n")
    mean, sigma, ml = res.mean(), res.std(), res.max_like()
    with open(root + "stats.dat", "w") as f:
        f.write(f"Nested Sampling Global Log-Evidence           :  {res.logz: .18E}  +/-  {res.logz_err: .18E}\n\n")
        f.write("Dim No.       Mean        Sigma\n")
        for i, (m, s) in enumerate(zip(mean, sigma), 1):
            f.write(f"{i:5d}  {m: .18E}  {s: .18E}\n")
        f.write("\nMaximum Likelihood Parameters\nDim No.        Parameter\n")
        for i, v in enumerate(ml, 1):
            f.write(f"{i:5d}  {v: .18E}\n")
        if labels:
            f.write("\n" + "\n".join(f"# {i}: {lab}" for i, lab in enumerate(labels, 1)) + "\n")
