"""N > 1 host logic on CPU: world_size 2, gloo backend.  The per-rank evaluator
is the CPU oracle here (the GPU model on the B200 box); what is under test is
the partition of points / rows and the single all-reduce of the lnew vector."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers as H
from lensed_b200.distributed import shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from lensed_b200.distributed import ShardedLikelihood
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cfg = H.synthetic_config("c4", 48, psf=(mode == "rows"))
        P = H.workloads.param_batch(cfg.extra["workload"], 5)
        full = cfg.oracle()
        ref = np.array([full.loglike(p) for p in P])
        if mode == "points":
            calls = []

            def evaluate(p):
                calls.append(p.shape[0])
                return np.array([full.loglike(x) for x in p])
            sh = ShardedLikelihood(evaluate, mode="points")
            got = sh.loglike_batch(P)
            q.put((rank, np.array_equal(got, ref), calls))
        else:
            # row strips: each rank's oracle sees only its rows (cropped image,
            # shifted pixel origin); without a PSF halo exchange the strips must
            # still add up to the full chi^2 for a PSF-free model
            cfg2 = H.synthetic_config("c4", 48, psf=False)
            full2 = cfg2.oracle()
            ref2 = np.array([full2.loglike(p) for p in P])
            state = {}

            def set_rows(r0, r1):
                strip = H.Config("strip", cfg2.objects, cfg2.params, cfg2.image[r0:r1], cfg2.weight[r0:r1], rule=cfg2.rule,
                                 pcs=(1.0, 1.0 + r0, 1.0, 1.0))
                state["m"] = strip.oracle()
            sh = ShardedLikelihood(lambda p: np.array([state["m"].loglike(x) for x in p]), mode="rows",
                                   set_rows=set_rows, height=48)
            got = sh.loglike_batch(P)
            q.put((rank, bool(np.all(np.abs(got - ref2) <= 1e-12*np.abs(ref2))), list(sh.rows)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["points", "rows"])
def test_two_ranks_gloo(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    if mode == "points":
        assert [c for _, _, c in res] == [[2], [3]]          # 5 points -> slices of 2 and 3
    else:
        assert [c for _, _, c in res] == [[0, 24], [24, 48]]


def _sampler_worker(rank, world, port, q):
    import math
    import torch.distributed as dist
    from lensed_b200.distributed import ShardedLikelihood
    from lensed_b200 import sampler as S
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        seen = []

        def evaluate(p):
            # this rank's slice of the batch only (float32 parameters, as the model gets them)
            seen.append(p.shape[0])
            p = p.astype(np.float64)
            return -0.5*(((p - 0.5)/0.05)**2).sum(axis=1) - p.shape[1]*math.log(0.05*math.sqrt(2*math.pi))
        sh = ShardedLikelihood(evaluate, mode="points")
        r = S.nested_sample(sh.loglike_batch, 2, nlive=100, batch=16, seed=7)
        q.put((rank, r.logz, r.logz_err, r.niter, sum(seen), r.nevals))
    finally:
        dist.destroy_process_group()


def test_sampler_two_ranks_same_chain_half_the_work_each():
    """One sampler per rank, same seed: the all-reduced batch is identical on
    both ranks, so both follow the same chain while each evaluates half of
    every batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, z0, e0, n0, w0, t0), (_, z1, e1, n1, w1, t1) = res
    assert z0 == z1 and n0 == n1 and t0 == t1
    assert w0 + w1 == t0 and abs(w0 - w1) <= 8
    assert abs(z0) < 4*e0 + 0.05
