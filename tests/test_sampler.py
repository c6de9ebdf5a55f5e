"""Batch-native nested sampler (lensed_b200/sampler.py) against analytic
evidences; the likelihood is a host function here, so no GPU is needed.  The
GPU test at the end runs it on a small lens model through lcu_loglike_batch."""
import math
import os

import numpy as np
import pytest

from lensed_b200 import sampler as S


def _gauss(mu, sig):
    mu, sig = np.asarray(mu, float), np.asarray(sig, float)

    def f(u):
        u = np.atleast_2d(u)
        return -0.5*(((u - mu)/sig)**2).sum(axis=1) - np.log(sig*math.sqrt(2*math.pi)).sum()
    return f


def test_gaussian_evidence_and_moments():
    """Normalised Gaussian well inside the unit cube: Z = 1."""
    calls = []
    f = _gauss([0.4, 0.6, 0.5], [0.02, 0.05, 0.03])

    def lb(u):
        calls.append(len(u))
        return f(u)
    r = S.nested_sample(lb, 3, nlive=400, batch=64, tol=0.01, seed=3)
    assert abs(r.logz) < 4*r.logz_err + 0.05, (r.logz, r.logz_err)
    assert np.allclose(r.mean(), [0.4, 0.6, 0.5], atol=0.005)
    assert np.allclose(r.std(), [0.02, 0.05, 0.03], rtol=0.15)
    # information of a Gaussian w.r.t. a unit prior: H = -ln((2 pi e)^(d/2) |Sigma|^(1/2))
    h = -(1.5*math.log(2*math.pi*math.e) + math.log(0.02*0.05*0.03))
    assert abs(r.information - h) < 0.3
    # every call after the initial live points is one full batch
    assert set(calls[7:]) == {64} and sum(calls) == r.nevals and len(calls) == r.nbatches
    assert abs(np.exp(r.logwt).sum() - 1) < 1e-9
    assert r.efficiency > 0.2


def test_same_seed_same_run_and_batch_size_independent_answer():
    f = _gauss([0.5, 0.5], [0.05, 0.1])
    a = S.nested_sample(f, 2, nlive=200, batch=32, seed=11)
    b = S.nested_sample(f, 2, nlive=200, batch=32, seed=11)
    assert a.logz == b.logz and np.array_equal(a.samples, b.samples)
    c = S.nested_sample(f, 2, nlive=200, batch=1, seed=12)       # one point per call, MultiNest style
    d = S.nested_sample(f, 2, nlive=200, batch=256, seed=13)
    for r in (a, c, d):
        assert abs(r.logz) < 4*r.logz_err + 0.05
    assert d.nbatches < c.nbatches/50


def test_periodic_dimension_wraps_around():
    """`wrap` (src/lensed.c:1254-1259, MultiNest's pWrap): a posterior that
    straddles the 0 / 1 seam of a periodic parameter -- a position angle near
    0 = 180 degrees -- has Z = 1 counting both sides of the seam.  With wrap the
    circle is cut opposite the live points, so that one small ellipsoid bounds
    them; without it the bound spans the whole dimension."""
    sig = np.array([0.02, 0.03])

    def lb(u):
        u = np.atleast_2d(u)
        d0 = np.minimum(u[:, 0], 1.0 - u[:, 0])                  # circular distance from the seam
        return -0.5*((d0/sig[0])**2 + ((u[:, 1] - 0.5)/sig[1])**2) - np.log(sig*math.sqrt(2*math.pi)).sum()
    for method in ("reject", "rwalk"):
        w = S.nested_sample(lb, 2, nlive=300, batch=64, tol=0.01, seed=21, wrap=[True, False], method=method)
        assert abs(w.logz) < 4*w.logz_err + 0.05, (method, w.logz, w.logz_err)
        x = w.samples[w.equal_weights(1), 0]
        assert 0.3 < np.mean(x < 0.5) < 0.7                        # both sides of the seam are populated
        assert np.all(np.minimum(x, 1 - x) < 0.15)
    plain = S.nested_sample(lb, 2, nlive=300, batch=64, tol=0.01, seed=21, method="reject")
    wrapd = S.nested_sample(lb, 2, nlive=300, batch=64, tol=0.01, seed=21, wrap=[True, False], method="reject")
    assert abs(plain.logz) < 4*plain.logz_err + 0.05              # still exact without, only slower
    assert wrapd.efficiency > 3*plain.efficiency, (wrapd.efficiency, plain.efficiency)
    with pytest.raises(ValueError):
        S.nested_sample(lb, 2, wrap=[False, False, True])


def test_two_modes_are_both_found():
    f1, f2 = _gauss([0.25, 0.3], [0.02, 0.02]), _gauss([0.75, 0.7], [0.02, 0.02])

    def lb(u):
        return np.logaddexp(f1(u), f2(u)) - math.log(2)
    r = S.nested_sample(lb, 2, nlive=400, batch=64, seed=5, split=True)
    assert abs(r.logz) < 4*r.logz_err + 0.05
    r1 = S.nested_sample(lb, 2, nlive=400, batch=64, seed=5)           # one ellipsoid: same answer, more evaluations
    assert abs(r1.logz) < 4*r1.logz_err + 0.05 and r1.nevals > r.nevals
    w = r.weights
    left = w[r.samples[:, 0] < 0.5].sum()
    assert 0.35 < left < 0.65


def test_maxiter_transform_and_files(tmp_path):
    f = _gauss([0.5], [0.1])
    r = S.nested_sample(f, 1, nlive=50, batch=8, seed=1, maxiter=100, transform=lambda u: np.array([10*u[0], 7.0]))
    assert r.niter == 100 and r.samples.shape == (150, 1) and r.physical.shape == (150, 2)
    assert np.all(r.physical[:, 1] == 7.0) and np.allclose(r.physical[:, 0], 10*r.samples[:, 0])
    S.write_multinest(str(tmp_path / "run-"), r, labels=["x", "const"])
    txt = np.loadtxt(tmp_path / "run-.txt")
    assert txt.shape == (150, 4) and abs(txt[:, 0].sum() - 1) < 1e-9
    assert np.allclose(txt[:, 1], -2*r.loglike)
    assert "Global Log-Evidence" in open(tmp_path / "run-stats.dat").read()
    pew = np.loadtxt(tmp_path / "run-post_equal_weights.dat", ndmin=2)
    assert pew.shape[1] == 3 and len(pew) > 5
    with pytest.raises(ValueError):
        S.nested_sample(f, 1, nlive=1)
    with pytest.raises(ValueError):
        S.nested_sample(lambda u: np.full(len(u), np.nan), 1, nlive=10)


def test_nan_likelihood_counts_as_zero():
    """Points where the model is undefined (NaN) are treated like MultiNest's
    logzero: never accepted, zero weight; the evidence is that of the rest."""
    g = _gauss([0.5, 0.5], [0.05, 0.05])

    def lb(u):
        ll = g(u)
        ll[u[:, 0] < 0.1] = np.nan
        return ll
    r = S.nested_sample(lb, 2, nlive=200, batch=32, seed=9)
    assert r.stats["nan_points"] > 0
    assert abs(r.logz) < 4*r.logz_err + 0.05
    assert np.all(r.weights[r.samples[:, 0] < 0.1] == 0)


def test_run_maps_cube_to_device_parameters():
    """sampler.run() feeds the model's batched entry point with float32
    parameter vectors in object order (free dimensions first in the cube,
    derived parameters last: src/lensed.c:236-271, src/nested.c:43-74)."""
    from lensed_b200 import host as Hh

    class FakeModel:
        def __init__(self):
            self.seen = []

        def loglike_batch(self, P):
            assert P.dtype == np.float32 and P.shape[1] == 3
            self.seen.append(P.copy())
            return -0.5*((P[:, 0] - 3.0)/0.2)**2 - 0.5*((P[:, 2] - 1.0)/0.1)**2

    pars = [Hh.Parameter("a.x", "x", 0, 0, 0, Hh.Uniform(0, 10)), Hh.Parameter("a.k", "k", 0, 0, 0, Hh.Delta(5.0)),
            Hh.Parameter("a.z", "z", 0, 0, 0, Hh.Uniform(0, 2))]
    cfg = Hh.Config({}, [Hh.ObjectEntry("a", "fake", "S", pars)])
    fm = FakeModel()
    like = Hh.Likelihood(cfg, fm)
    assert like.ndims == 2 and like.npars == 3
    r = S.run(like, nlive=100, batch=16, seed=2)
    assert all(np.all(P[:, 1] == 5.0) for P in fm.seen)
    m = r.mean()            # sampler order: x, z, then the derived k
    assert abs(m[0] - 3.0) < 0.05 and abs(m[1] - 1.0) < 0.03 and abs(m[2] - 5.0) < 1e-9
    # Z = integral of L over the prior = (0.2 sqrt(2 pi)/10) (0.1 sqrt(2 pi)/2)
    z = math.log(0.2*math.sqrt(2*math.pi)/10) + math.log(0.1*math.sqrt(2*math.pi)/2)
    assert abs(r.logz - z) < 4*r.logz_err + 0.05


@pytest.mark.gpu
def test_sampler_on_a_lens_model(gpu_ctx):
    """Three free parameters of a small SIE + Sersic scene: the posterior
    brackets the truth, every step is one batched launch."""
    import helpers as H
    import lensed_b200 as L
    cfg = H.synthetic_config("c4", 64, psf_shape=(5, 5))
    m = cfg.product(gpu_ctx)
    truth = cfg.params.astype(np.float64)
    free = [2, 3, 10]                         # lens radius, lens axis ratio, source magnitude
    span = [1.0, 0.1, 0.3]
    n0 = L.launch_count()

    def lb(u):
        P = np.tile(truth, (len(u), 1))
        for k, (i, s) in enumerate(zip(free, span)):
            P[:, i] = truth[i] + (u[:, k] - 0.5)*2*s
        return m.loglike_batch(P.astype(np.float32))
    r = S.nested_sample(lb, 3, nlive=100, batch=50, tol=0.5, seed=4)
    launches = L.launch_count() - n0
    assert launches <= 5*r.nbatches                       # set_params, copy-free render, convolve, reduce per batch
    mean, std = r.mean(), r.std()
    assert np.all(np.abs(mean - 0.5) < 5*std + 0.02), (mean, std)
    assert np.all(std < 0.2)
    assert r.nevals == 100 + 50*(r.nbatches - 2)


def test_posterior_in_a_corner_of_the_prior():
    """A likelihood peaked at a corner of the unit cube in 8 dimensions: almost
    all of a bounding ellipsoid lies outside the cube; draws fall back to the
    box around the live points."""
    sig = 0.01

    def lb(u):
        return -0.5*((u/sig)**2).sum(axis=1)
    r = S.nested_sample(lb, 8, nlive=100, batch=64, seed=3, tol=0.5)
    z = 8*math.log(sig*math.sqrt(math.pi/2))           # half a Gaussian per dimension
    assert abs(r.logz - z) < 5*r.logz_err + 0.1, (r.logz, z)


def test_random_walks_take_over_when_rejection_stops_paying():
    """Curved, badly scaled posterior in 8 dimensions (two Rosenbrock-like
    pairs times a narrow Gaussian): rejection from one ellipsoid accepts a few
    per mille; the default hands over to constrained random walks and gets the
    evidence with a fraction of the evaluations."""
    def banana(x, y):
        X, Y = (x - 0.5)*6, (y - 0.3)*6
        return -((1 - X)**2 + 10*(Y - X**2)**2)
    g = np.linspace(0, 1, 1501)
    G = (g[:-1] + g[1:])/2
    XX, YY = np.meshgrid(G, G, indexing="ij")
    z2 = math.log(np.exp(banana(XX, YY)).mean())
    sig = np.array([0.002, 0.01, 0.003, 0.03])

    def lb(u):
        return (banana(u[:, 0], u[:, 1]) + banana(u[:, 2], u[:, 3]) - 0.5*(((u[:, 4:] - 0.5)/sig)**2).sum(axis=1)
                - np.log(sig*math.sqrt(2*math.pi)).sum())
    r = S.nested_sample(lb, 8, nlive=300, batch=128, seed=2)
    assert r.stats["walk_rounds"] > 0 and 0.2 < r.stats["walk_acceptance"] < 0.6
    assert abs(r.logz - 2*z2) < 3*r.logz_err + 0.1, (r.logz, 2*z2, r.logz_err)
    assert np.allclose(r.mean()[4:], 0.5, atol=3*sig.max()/10 + 5e-3)
    assert r.efficiency > 0.005
    e = S.nested_sample(lb, 8, nlive=300, batch=128, seed=2, method="reject", maxiter=3000)
    assert e.stats["walk_rounds"] == 0
    w = S.nested_sample(lb, 8, nlive=300, batch=128, seed=2, method="rwalk", walks=40, maxiter=3000)
    assert w.stats["walk_rounds"] > 0 and w.niter == e.niter == 3000
