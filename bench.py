#!/usr/bin/env python
"""bench.py -- likelihood evaluations/s of the per-likelihood model-image hot
path on the synthetic 1024^2 configuration (C4: SIE+shear lens, 2 Sersic
sources + sky, 25x25 PSF, rule g7k15), N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU implementation

A "step" is one batched pass of the hot path (set_params -> render -> convolve
-> chi^2 -> reduce) over B parameter points per GPU.  `value` is whole-job
evaluations/s with parameters already resident in HBM (device-pointer entry
point, CUDA events on the launching stream, max over ranks); `e2e` is the same
metric through the host-buffer C-ABI call lcu_loglike_batch (pinned staging,
H2D of the parameters and D2H of the log-likelihoods inside the timed region).
Multi-GPU: parameter points are sharded across ranks (weak scaling, B points
per GPU); the only communication is the all-reduce of the log-likelihood vector
over NCCL.

The reference arm times the reference's own kernels compiled on the host
(oracle/_ref, built from /root/reference by oracle/build_ref.py; "reference")
or, where that library is absent, the oracle port ("port"), on all host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "likelihood_evals_per_s"
UNIT = "evals/s"

# stdout carries exactly one JSON line: libraries that print to stdout (NCCL's
# version banner, for one) are sent to stderr for the whole run
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clock / throttle samples during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s in sm if s > 0.5*max(mx or [1])] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(name):
    from lensed_b200 import workloads
    return workloads.c4(1024) if name == "c4" else workloads.c5(4096)


def cpu_library():
    """(ctypes lib, kind): the reference's own kernels compiled on the host if
    present, else the oracle port; both built -O3 -ffast-math with OpenMP."""
    from oracle import pyoracle as O
    if O.available("ref_fast"):
        return O.lib("ref_fast"), "reference"
    return O.lib("fast"), "port"


def cpu_model(w, image, weight):
    from oracle import pyoracle as O
    lib, kind = cpu_library()
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    lib.orc_set_threads(ncpu)
    qq, ww = O.quad_rule(w["rule"])
    return O.Model(w["objects"], image, weight, qq, ww, psf=w["psf"], _lib=lib), lib, kind


def synthetic_observation(w, ctx=None):
    """Observed image + weight map of the workload.  With a GPU context the
    truth model is rendered by the product (input data only); on the CPU arm it
    is rendered by the CPU library."""
    from lensed_b200 import workloads
    size = w["width"]
    blank = np.zeros((size, size), np.float32)
    if ctx is not None:
        import lensed_b200 as L
        m0 = L.Model(ctx, w["objects"], blank, blank + 1, rule=w["rule"], psf=w["psf"])
        truth = m0.render(w["truth"], raw=False, error=False, chi=False)["model"]
        m0.close()
    else:
        om, _, _ = cpu_model(w, blank, blank + 1)
        _, truth, _ = om.loglike(w["truth"], want_maps=True)
    return workloads.observe(truth, w["noise_seed"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from lensed_b200 import workloads
    w = workload(args.workload)
    image, weight = synthetic_observation(w)
    om, lib, kind = cpu_model(w, image, weight)
    cores = lib.orc_max_threads()
    P = workloads.param_batch(w, args.steps + args.warmup)
    for i in range(args.warmup):
        om.loglike(P[i])
    t0 = time.perf_counter()
    for i in range(args.steps):
        om.loglike(P[args.warmup + i])
    dt = time.perf_counter() - t0
    value = args.steps/dt
    nq = om.L.orc_quad_size(w["rule"].encode())
    sample = f"{args.steps} full {w['name']} evaluations (one per step), {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3*dt/args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {'+'.join(w['objects'])}, {w['width']}x{w['height']}, PSF 25x25, rule {w['rule']}",
                   "points_per_step": 1, "rays_per_eval": w["width"]*w["height"]*nq},
        "grays_per_s": value*w["width"]*w["height"]*nq/1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import lensed_b200 as L
    from lensed_b200 import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this benchmark has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = workload(args.workload)
    ctx = L.Context(device=local)
    image, weight = synthetic_observation(w, ctx)
    flags = 0 if args.math == "strict" else (L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
    if args.math == "contract":
        # opt-in, outside the parity bar: FMA contraction on top of the fast build
        flags |= L.LCU_FAST_MATH
    model = L.Model(ctx, w["objects"], image, weight, rule=w["rule"], psf=w["psf"], flags=flags)
    B = args.batch
    nq = model.nq
    work = workloads.work_per_eval(w, nq)

    # this rank's parameter points: a distinct slice of one global batch
    P_all = workloads.param_batch(w, B*world)
    P = np.ascontiguousarray(P_all[rank*B:(rank + 1)*B])
    d_params = torch.from_numpy(P).cuda()
    d_lnew = torch.zeros(B*world, dtype=torch.float64, device="cuda")
    mine = d_lnew[rank*B:(rank + 1)*B]
    stream = torch.cuda.current_stream()

    def step():
        if world > 1:
            d_lnew.zero_()
        model.loglike_batch_device(B, d_params.data_ptr(), mine.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(d_lnew)          # every rank fills its own slots: sum == gather

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp32_peak = ctx.fp32_peak_tflops()
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: device-resident -----------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    model.profile(True)
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    prof = model.profile_get()
    model.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B*world*args.steps/(ms*1e-3)
    lnew_dev = d_lnew.cpu().numpy().copy()

    # ---- timed region: end to end through the host-buffer C-ABI call ---------
    # (N > 1: through lensed_b200.distributed, which adds the NCCL all-reduce
    # that assembles the full lnew vector on every rank)
    from lensed_b200.distributed import ShardedLikelihood
    sharded = ShardedLikelihood.for_model(model, mode="points", device=f"cuda:{local}") if world > 1 else None

    def host_step():
        if sharded is None:
            return model.loglike_batch(P)
        return sharded.loglike_batch(P_all)[rank*B:(rank + 1)*B]

    for _ in range(2):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lnew_host = host_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = B*world*args.steps/float(t.item())
    assert np.array_equal(lnew_host, lnew_dev[rank*B:(rank + 1)*B]), "host and device entry points disagree"

    if rank == 0:
        pk = peaks()
        nchunk = max(args.steps, 1)
        render_ms = prof["render_ms"]/nchunk
        conv_ms = prof["convolve_ms"]/nchunk
        render_flops = work["render_flops"]*B
        achieved = render_flops/(render_ms*1e-3)/1e12 if render_ms > 0 else None
        traffic = pipe_busy = None
        tpath = os.path.join(ROOT, "profiles", "render_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get(f"{args.workload}_B{B}")
                if flags != 0 and model.rays_per_thread == 2:
                    pipe_busy = tj.get(f"{args.workload}_fp32_pipe_busy")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {'+'.join(w['objects'])}, {w['width']}x{w['height']}, PSF 25x25, rule {w['rule']}",
                       "points_per_gpu_per_step": B, "rays_per_eval": work["rays"], "parallelism": f"points x{world}",
                       "rays_per_thread": model.rays_per_thread,
                       "math": "strict: IEEE ops in source order, accurate libdevice functions, no FMA contraction" if flags == 0 else
                               "contract: the fast build plus FMA contraction (LCU_FAST_MATH); NOT held to the parity bar" if args.math == "contract" else
                               "LCU_FAST_INTRINSICS|LCU_FAST_ATANH: exp/log of source and foreground objects and atanh of lens objects "
                               "on the hardware exp2/log2 units; division, sqrt, atan, no FMA contraction and the summation order "
                               "as in the strict build (parity-tested to the same bounds, tests/test_gpu_parity.py)",
                       "l2": f"working set per step {(B*w['width']*w['height']*4 + 8*w['width']*w['height'])/1e6:.0f} MB of staged "
                             "images > 126 MB L2 (no explicit flush)" if B*w['width']*w['height']*4 > 126e6 else
                             "compute-bound kernel, working set fits L2; inputs re-read from L2 by design"},
            "grays_per_s": value*work["rays"]/1e9,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(P.nbytes), "d2h_bytes_per_step": int(8*B)},
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": "lcu_render_pair" if model.rays_per_thread == 2 else "lcu_render_s1", "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved/fp32_peak if achieved and fp32_peak else None,
                # the reference's arithmetic is FMUL + FADD where an FMA would do (no contraction, parity):
                # source-level flops can reach at most half the FFMA peak
                "frac_of_unfused_ceiling": achieved/(0.5*fp32_peak) if achieved and fp32_peak else None,
                "traffic": traffic,
                "peak_source": "FFMA micro-benchmark run in this process (MEASURED_PEAKS.json records no FP32 peak)",
                "algorithmic": f"{w['flops_per_ray']} flop + {w['transc_per_ray']} transcendental calls per ray x "
                               f"{work['rays']} rays x {B} points per launch",
                "launch_ms": render_ms,
                # share of cycles the FP32 pipe is busy, from the committed ncu capture of this kernel
                # (profiles/): the instruction-level roofline; `frac` counts source-level flops only
                "fp32_pipe_busy_ncu": pipe_busy,
                "transc_calls_per_s": work["transc"]*B/(render_ms*1e-3) if render_ms > 0 else None,
                "grays_per_s_kernel": work["rays"]*B/(render_ms*1e-3)/1e9 if render_ms > 0 else None,
            },
            "stage_ms_per_step": {"set_params": prof["set_params_ms"]/nchunk, "render": render_ms, "convolve_chi2": conv_ms,
                                  "reduce": prof["reduce_ms"]/nchunk},
            # the 25x25 direct convolution is FP32-bound, not HBM-bound: 2 Pw Ph flop per pixel, issued as
            # separate FMUL + FADD per tap (the reference's rounding, bit-exact against the oracle), so its
            # ceiling is half the FFMA peak; its HBM traffic (16 B/pixel) is a few per cent of the bandwidth
            "roofline_convolve": {
                "kernel": "lcu_convolve", "bound": "fp32",
                "achieved": work["convolve_flops"]*B/(conv_ms*1e-3)/1e12 if conv_ms > 0 else None,
                "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": (work["convolve_flops"]*B/(conv_ms*1e-3)/1e12)/fp32_peak if conv_ms > 0 and fp32_peak else None,
                "frac_of_unfused_ceiling": (work["convolve_flops"]*B/(conv_ms*1e-3)/1e12)/(0.5*fp32_peak)
                if conv_ms > 0 and fp32_peak else None,
                "hbm_gbs": work["hbm_bytes"]*B/(conv_ms*1e-3)/1e9 if conv_ms > 0 else None,
                "hbm_peak_gbs": pk.get("hbm_gbs"),
            },
        }
        # CPU baseline on this box's host cores, N = 1 only, bounded sample
        if world == 1 and not args.no_cpu_baseline:
            try:
                om, lib, kind = cpu_model(w, image, weight)
                cores = lib.orc_max_threads()
                om.loglike(P[0])
                n = 0
                t0 = time.perf_counter()
                while n < 3 or (time.perf_counter() - t0 < args.cpu_seconds and n < B):
                    ref = om.loglike(P[n % B])
                    n += 1
                dtc = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": n/dtc, "unit": UNIT, "cores": cores, "kind": kind,
                                        "sample": f"{n} full {w['name']} evaluations on {cores} host threads ({dtc:.1f} s)"}
                # in-bench parity check of one point against the strict-float32 oracle
                from oracle import pyoracle as O
                qq, ww = O.quad_rule(w["rule"])
                strict = O.Model(w["objects"], image, weight, qq, ww, psf=w["psf"]).loglike(P[0])
                line["parity_lnew_rel"] = abs(lnew_dev[0] - strict)/abs(strict)
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "unavailable", "sample": repr(e)[:200]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"])
    ap.add_argument("--batch", type=int, default=32, help="parameter points per GPU per step")
    ap.add_argument("--math", default="fast", choices=["strict", "fast", "contract"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
