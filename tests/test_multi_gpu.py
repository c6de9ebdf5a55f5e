"""N > 1 on hardware: world_size = number of visible GPUs (2 ... 8), NCCL.
Skipped below two devices (the round-end `pytest -m gpu` box has one GPU; run
with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).

rows    every rank evaluates all points on its strip of image rows (PSF halo
        re-rendered, not exchanged: SURVEY.md section 8e way 2); one NCCL
        all-reduce adds the strips' -chi^2/2; the sum must equal the single-GPU
        log-likelihood (same pixels, same per-32-pixel partial sums; only the
        final order of addition differs: 1e-12).
points  rank g evaluates its slice of the batch, the all-reduce assembles the
        vector: bit-identical to the single-GPU batch.
"""
import os
import socket
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, which, size, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import lensed_b200 as L
    from lensed_b200.distributed import ShardedLikelihood
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        ctx = L.Context(device=rank, objects_dir=H.OBJECTS_DIR)
        w = H.workloads.c4(size) if which == "c4" else H.workloads.c5(size)
        blank = np.zeros((size, size), np.float32)
        m0 = L.Model(ctx, w["objects"], blank, blank + 1, rule=w["rule"], psf=w["psf"])
        truth = m0.render(w["truth"], raw=False, error=False, chi=False)["model"]
        m0.close()
        image, weight = H.workloads.observe(truth, w["noise_seed"])
        flags = L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH
        m = L.Model(ctx, w["objects"], image, weight, rule=w["rule"], psf=w["psf"], flags=flags)
        P = H.workloads.param_batch(w, 5)
        full = m.loglike_batch(P)                         # single GPU, whole image, whole batch
        sh = ShardedLikelihood.for_model(m, mode=mode, device=f"cuda:{rank}")
        got = sh.loglike_batch(P)
        rows = list(getattr(sh, "rows", (0, size)))
        if mode == "points":
            ok = bool(np.array_equal(got, full))
        else:
            ok = bool(np.all(np.abs(got - full) <= 1e-12*np.abs(full)))
        q.put((rank, ok, rows, float(np.abs(got - full).max()/np.abs(full).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("mode,which,size", [("rows", "c4", 256), ("rows", "c5", 512), ("points", "c4", 256)])
def test_nccl_ranks_agree_with_one_gpu(mode, which, size):
    import torch.multiprocessing as mp
    world = min(_ngpus(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, which, size, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    if mode == "rows":
        # the strips tile the image
        assert [r[2][0] for r in res] == [size*g//world for g in range(world)]
        assert res[-1][2][1] == size
