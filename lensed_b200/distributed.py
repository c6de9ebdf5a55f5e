"""Multi-GPU evaluation of a batch of parameter points.

One process per GPU (``torchrun``); the image, weight map, PSF and tables are
replicated when each rank creates its ``Model``.  Two partitions:

* ``points``: rank g evaluates the contiguous slice [g*B/G, (g+1)*B/G) of the
  batch.  Each rank writes its slots of a zeroed ``lnew[B]`` vector and one
  all-reduce (sum) delivers the full vector everywhere -- the only collective
  on the path (the reference has none: MultiNest's MPI mode runs whole
  independent processes, src/lensed.c:1243).
* ``rows``: for very large images every rank evaluates all points on its strip
  of image rows (plus the PSF halo it renders itself); the all-reduce then
  adds the strips' -chi^2/2.

The per-rank evaluation is a callable so that the partition / collective logic
can be exercised on CPU with the gloo backend (tests/test_distributed.py).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous slice [n*rank/world, n*(rank+1)/world) (ragged when world does not divide n)."""
    return (n*rank)//world, (n*(rank + 1))//world


class ShardedLikelihood:
    def __init__(self, evaluate: Callable[[np.ndarray], np.ndarray], mode: str = "points", group=None,
                 device: Optional[str] = None, set_rows: Optional[Callable[[int, int], None]] = None,
                 height: Optional[int] = None):
        import torch.distributed as dist
        if mode not in ("points", "rows"):
            raise ValueError("mode must be 'points' or 'rows'")
        self.dist = dist
        self.evaluate = evaluate
        self.mode = mode
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device
        if mode == "rows":
            if set_rows is None or height is None:
                raise ValueError("rows mode needs set_rows and height")
            r0, r1 = shard_range(height, self.rank, self.world)
            if r1 <= r0:
                raise ValueError("more ranks than image rows")
            set_rows(r0, r1)
            self.rows = (r0, r1)

    @classmethod
    def for_model(cls, model, mode: str = "points", group=None, device: Optional[str] = None):
        """Bind to a lensed_b200.Model on this rank's GPU.  With a CUDA `device`
        the batch stays on the device between the evaluation and the
        all-reduce: parameters go up once from pinned memory, the C ABI's
        device-pointer entry writes this rank's slots of the lnew vector, NCCL
        reduces that vector in place and it comes down once."""
        self = cls(model.loglike_batch, mode=mode, group=group, device=device, set_rows=model.set_rows,
                   height=model.height)
        self.model = model if (device and str(device).startswith("cuda")) else None
        self._bufs = None
        return self

    def _device_batch(self, params: np.ndarray) -> np.ndarray:
        import torch
        nbatch, npars = params.shape
        if self._bufs is None or self._bufs[0].shape[0] < nbatch:
            cap = max(nbatch, 64)
            self._bufs = (torch.empty((cap, npars), dtype=torch.float32).pin_memory(),
                          torch.empty((cap, npars), dtype=torch.float32, device=self.device),
                          torch.empty(cap, dtype=torch.float64, device=self.device),
                          torch.empty(cap, dtype=torch.float64).pin_memory())
        h_par, d_par, d_lnew, h_lnew = self._bufs
        stream = torch.cuda.current_stream(self.device)
        h_par[:nbatch].copy_(torch.from_numpy(params))
        d_par[:nbatch].copy_(h_par[:nbatch], non_blocking=True)
        lnew = d_lnew[:nbatch]
        if self.mode == "points":
            b0, b1 = shard_range(nbatch, self.rank, self.world)
            lnew.zero_()
            if b1 > b0:
                self.model.loglike_batch_device(b1 - b0, d_par[b0:b1].data_ptr(), lnew[b0:b1].data_ptr(), stream.cuda_stream)
        else:
            self.model.loglike_batch_device(nbatch, d_par.data_ptr(), lnew.data_ptr(), stream.cuda_stream)
        if self.world > 1:
            self.dist.all_reduce(lnew, op=self.dist.ReduceOp.SUM, group=self.group)
        h_lnew[:nbatch].copy_(lnew, non_blocking=True)
        stream.synchronize()
        return h_lnew[:nbatch].numpy().copy()

    def loglike_batch(self, params) -> np.ndarray:
        """params[B, npars] (identical on every rank) -> lnew[B] on every rank."""
        import torch
        params = np.ascontiguousarray(params, dtype=np.float32)
        nbatch = params.shape[0]
        if getattr(self, "model", None) is not None and nbatch > 0:
            return self._device_batch(params)
        out = np.zeros(nbatch, np.float64)
        if self.mode == "points":
            b0, b1 = shard_range(nbatch, self.rank, self.world)
            if b1 > b0:
                out[b0:b1] = self.evaluate(params[b0:b1])
        else:
            out[:] = self.evaluate(params)
        if self.world > 1:
            t = torch.from_numpy(out)
            if self.device:
                t = t.to(self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            out = t.cpu().numpy()
        return out
