#!/usr/bin/env python
"""bench.py -- likelihood evaluations/s of the per-likelihood model-image hot
path (set_params -> render -> convolve -> chi^2 -> reduce) on N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W              # C4-1024, B = 32 points per GPU per step (weak scaling)
    python bench.py --workload c5 --gpus N                     # C5-4096, 64 points per step in total (strong scaling)
    python bench.py --workload c5 --parallelism rows --gpus N  # C5-4096, image row strips across GPUs (SURVEY 8e way 2)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU implementation on the host cores

Workloads (BASELINE.json:configs[3] and [4]; lensed_b200/workloads.py):
  c4  1024^2, sie_plus_shear + 2 sersic + sky, 25x25 PSF, rule g7k15   (the configuration `metric` is quoted on)
  c5  4096^2, epl_plus_shear + 3 sersic + sky, 25x25 PSF, rule g3k7

A "step" is one batched pass of the hot path over the step's parameter points.
`value` is whole-job evaluations/s with the parameters already resident in HBM
(device-pointer entry point, CUDA events on the launching stream, max over
ranks); `e2e` is the same metric through the host-buffer C-ABI call
lcu_loglike_batch (pinned staging, H2D of the parameters and D2H of the
log-likelihoods inside the timed region).  The object plugins are the
reference's own objects/*.cl, unmodified (tests/golden/objects).

Multi-GPU.  points: the step's parameter points are sharded across ranks, every
rank fills its slots of a zeroed lnew vector and one NCCL all-reduce (sum ==
gather) assembles it everywhere.  rows: every rank evaluates all points on its
strip of image rows (PSF halo re-rendered, not exchanged) and the all-reduce
adds the strips' -chi^2/2.  No other communication.

The reference arm times the reference's own kernels compiled on the host
(oracle/_ref, built from /root/reference by oracle/build_ref.py; "reference")
or, where that library is absent, the oracle port ("port"), on all host cores.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# plugin directory: the reference's own objects/*.cl, verbatim (tests/golden/objects/README.md)
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")
OBJECTS_NOTE = "tests/golden/objects (the reference's objects/*.cl, byte for byte)"

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


def workloads_module():
    """lensed_b200/workloads.py (numpy only) loaded by path: the reference arm
    must not import the package, whose __init__ maps the CUDA library."""
    spec = importlib.util.spec_from_file_location("lcu_workloads", os.path.join(ROOT, "lensed_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ClockSampler:
    """nvidia-smi clock / throttle samples during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index = index
        self.period_ms = period_ms
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", str(self.period_ms)], stdout=subprocess.PIPE,
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
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s in sm if s > 0.5*max(mx or [1])] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_min_mhz": min(load) if load else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w": statistics.median(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def describe(w):
    return f"{w['name']}: {'+'.join(w['objects'])}, {w['width']}x{w['height']}, PSF 25x25, rule {w['rule']}"


# ---------------------------------------------------------------------------
# CPU side: the reference arm and the cpu_baseline leg (the only places that
# execute anything under oracle/)
# ---------------------------------------------------------------------------
CPU_KINDS = (
    # variant of oracle/pyoracle, kind, description
    ("ref_simd512", "reference", "the reference's objects/*.cl and generated compute() compiled on the host with float = 16 "
                                 "consecutive work-items (AVX-512 lanes, libmvec), -O3 -ffast-math, OpenMP over work-groups: "
                                 "what a vectorising OpenCL CPU runtime does with them"),
    ("ref_simd", "reference", "the reference's objects/*.cl and generated compute() compiled on the host with float = 8 "
                              "consecutive work-items (AVX2 lanes, libmvec), -O3 -ffast-math, OpenMP over work-groups: "
                              "what a vectorising OpenCL CPU runtime does with them"),
    ("ref_fast", "reference", "the reference's kernels compiled on the host one scalar work-item at a time, -O3 -ffast-math, "
                              "OpenMP over pixels"),
    ("fast", "port", "the oracle port (oracle/lensed_oracle.c), -O3 -ffast-math, OpenMP"),
)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_models(w, image, weight):
    """[(variant, kind, description, model, lib)] for every CPU build present,
    all host cores whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    from oracle import pyoracle as O
    out = []
    for variant, kind, text in CPU_KINDS:
        if not O.available(variant):
            continue
        if kind == "port" and out:
            continue                    # the port only stands in where no reference build exists
        lib = O.lib(variant)
        lib.orc_set_threads(host_threads())
        qq, ww = O.quad_rule(w["rule"])
        try:
            out.append((variant, kind, text, O.Model(w["objects"], image, weight, qq, ww, psf=w["psf"], _lib=lib), lib))
        except Exception:
            continue                    # a build that does not hold this object list
    if not out:
        raise RuntimeError("no CPU build of the hot path under oracle/ (run __graft_entry__.build())")
    return out


def time_cpu(om, P, seconds, min_evals=2, max_evals=None):
    om.loglike(P[0])
    n = 0
    t0 = time.perf_counter()
    while n < min_evals or (time.perf_counter() - t0 < seconds and (max_evals is None or n < max_evals)):
        om.loglike(P[n % len(P)])
        n += 1
    return n, time.perf_counter() - t0


def synthetic_observation(w, wl, ctx=None):
    """Observed image + weight map of the workload.  With a GPU context the
    truth model is rendered by the product (input data only); on the CPU arm it
    is rendered by the CPU library."""
    size = w["width"]
    blank = np.zeros((size, size), np.float32)
    if ctx is not None:
        import lensed_b200 as L
        m0 = L.Model(ctx, w["objects"], blank, blank + 1, rule=w["rule"], psf=w["psf"])
        truth = m0.render(w["truth"], raw=False, error=False, chi=False)["model"]
        m0.close()
    else:
        om = cpu_models(w, blank, blank + 1)[0][3]
        _, truth, _ = om.loglike(w["truth"], want_maps=True)
    return wl.observe(truth, w["noise_seed"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = workloads_module()
    w = wl.c4(1024) if args.workload == "c4" else wl.c5(4096)
    image, weight = synthetic_observation(w, wl)
    builds = cpu_models(w, image, weight)
    P = wl.param_batch(w, args.steps + args.warmup)
    # one evaluation of every build present; the arm then runs the fastest one
    probe = []
    for variant, kind, text, om, lib in builds:
        om.loglike(P[0])
        t0 = time.perf_counter()
        om.loglike(P[0])
        probe.append(time.perf_counter() - t0)
    best = int(np.argmin(probe))
    variant, kind, text, om, lib = builds[best]
    cores = lib.orc_max_threads()
    for i in range(args.warmup):
        om.loglike(P[i])
    t0 = time.perf_counter()
    for i in range(args.steps):
        om.loglike(P[args.warmup + i])
    dt = time.perf_counter() - t0
    value = args.steps/dt
    nq = om.L.orc_quad_size(w["rule"].encode())
    sample = f"{args.steps} full {w['name']} evaluations (one per step), {cores} host threads; {variant}: {text}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3*dt/args.steps, "higher_is_better": True,
        "scaling": "weak" if args.workload == "c4" else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(w), "points_per_step": 1, "rays_per_eval": w["width"]*w["height"]*nq,
                   "objects_dir": OBJECTS_NOTE if kind == "port" else "the reference's objects/*.cl (compiled into oracle/_ref)"},
        "grays_per_s": value*w["width"]*w["height"]*nq/1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "builds_probed": {b[0]: 1.0/t for b, t in zip(builds, probe)}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------
MATH_TEXT = {
    "strict": "strict: IEEE ops in source order, accurate libdevice functions, no FMA contraction",
    "contract": "contract: the fast build plus FMA contraction (LCU_FAST_MATH); NOT held to the parity bar",
    "fast": "LCU_FAST_INTRINSICS|LCU_FAST_ATANH: exp/log of source and foreground objects and atanh of lens objects "
            "on the hardware exp2/log2 units; division, sqrt, atan, no FMA contraction and the summation order "
            "as in the strict build (parity-tested to the same bounds, tests/test_gpu_parity.py)",
}


def math_flags(L, math):
    flags = 0 if math == "strict" else (L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
    if math == "contract":
        flags |= L.LCU_FAST_MATH      # opt-in, outside the parity bar
    return flags


def committed_profile(workload, B, pair):
    """DRAM traffic and FP32-pipe share of the render kernel from the committed
    ncu capture (profiles/render_traffic.json): NOT measured in this run."""
    tpath = os.path.join(ROOT, "profiles", "render_traffic.json")
    try:
        tj = json.load(open(tpath))
    except Exception:
        return None, None
    return tj.get(f"{workload}_B{B}"), (tj.get(f"{workload}_fp32_pipe_busy") if pair else None)


def render_roofline(w, work, B, render_ms, fp32_peak, model, traffic, pipe_busy):
    achieved = work["render_flops"]*B/(render_ms*1e-3)/1e12 if render_ms > 0 else None
    return {
        "kernel": "lcu_render_pair" if model.rays_per_thread == 2 else "lcu_render_s1", "bound": "fp32",
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved/fp32_peak if achieved and fp32_peak else None,
        # the reference's arithmetic is FMUL + FADD where an FMA would do (no contraction, parity):
        # source-level flops can reach at most half the FFMA peak
        "frac_of_unfused_ceiling": achieved/(0.5*fp32_peak) if achieved and fp32_peak else None,
        "traffic": traffic,
        "peak_source": "FFMA micro-benchmark run in this process (MEASURED_PEAKS.json records no FP32 peak)",
        "algorithmic": f"{w['flops_per_ray']} flop + {w['transc_per_ray']} transcendental calls per ray x "
                       f"{work['rays']} rays x {B} points per launch",
        "launch_ms": render_ms,
        # share of cycles the FP32 pipe is busy: the instruction-level roofline; `frac` counts source-level flops only
        "fp32_pipe_busy_ncu": pipe_busy,
        "from_committed_profile": {"fields": ["traffic", "fp32_pipe_busy_ncu"], "source": "profiles/render_traffic.json",
                                   "note": "ncu capture of this kernel committed under profiles/, not measured in this run"},
        "transc_calls_per_s": work["transc"]*B/(render_ms*1e-3) if render_ms > 0 else None,
        "grays_per_s_kernel": work["rays"]*B/(render_ms*1e-3)/1e9 if render_ms > 0 else None,
    }


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import lensed_b200 as L
    from lensed_b200 import workloads as wl
    from lensed_b200.distributed import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this benchmark has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = wl.c4(1024) if args.workload == "c4" else wl.c5(4096)
    rows_mode = args.parallelism == "rows"
    strong = rows_mode or args.scaling == "strong"
    if rows_mode:
        B = args.batch                  # every rank evaluates all B points on its rows
        Btot = B
    elif strong:
        Btot = args.batch               # the step's points, divided among the ranks
        if Btot % world:
            raise SystemExit(f"bench.py: --scaling strong needs --batch ({Btot}) divisible by the number of GPUs ({world})")
        B = Btot//world
    else:
        B = args.batch                  # points per GPU per step
        Btot = B*world

    ctx = L.Context(device=local, objects_dir=OBJECTS_DIR)
    image, weight = synthetic_observation(w, wl, ctx)
    flags = math_flags(L, args.math)
    model = L.Model(ctx, w["objects"], image, weight, rule=w["rule"], psf=w["psf"], flags=flags)
    nq = model.nq
    work = wl.work_per_eval(w, nq)
    H = w["height"]
    strip = (0, H)
    if rows_mode:
        strip = shard_range(H, rank, world)
        model.set_rows(*strip)

    # the step's parameter points; this rank's are a distinct slice of them (points mode)
    P_all = wl.param_batch(w, Btot)
    P = np.ascontiguousarray(P_all if rows_mode else P_all[rank*B:(rank + 1)*B])
    d_params = torch.from_numpy(P).cuda()
    d_lnew = torch.zeros(Btot, dtype=torch.float64, device="cuda")
    mine = d_lnew if rows_mode else d_lnew[rank*B:(rank + 1)*B]
    stream = torch.cuda.current_stream()

    def step():
        if world > 1 and not rows_mode:
            d_lnew.zero_()
        model.loglike_batch_device(B, d_params.data_ptr(), mine.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(d_lnew)          # points: every rank fills its own slots, sum == gather; rows: sum of strips

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(nsteps):
            step()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    ms = timed(args.steps)
    launches = L.launch_count() - launches0
    prof = model.profile_get()
    model.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    value = Btot*args.steps/(ms*1e-3)
    lnew_dev = d_lnew.cpu().numpy().copy()

    # ---- timed region: end to end through the host-buffer C-ABI call ---------
    # (N > 1: through lensed_b200.distributed, which adds the NCCL all-reduce
    # that assembles the full lnew vector on every rank)
    from lensed_b200.distributed import ShardedLikelihood
    sharded = None
    if world > 1:
        sharded = ShardedLikelihood.for_model(model, mode="rows" if rows_mode else "points", device=f"cuda:{local}")

    def host_step():
        if sharded is None:
            return model.loglike_batch(P)
        return sharded.loglike_batch(P_all)

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
    e2e = Btot*args.steps/float(t.item())
    if world > 1 and rows_mode:
        assert np.allclose(lnew_host, lnew_dev, rtol=1e-13, atol=0), "host and device entry points disagree"
    else:
        assert np.array_equal(lnew_host, lnew_dev), "host and device entry points disagree"

    # ---- sustained: the same step for >= args.sustain seconds, clocks sampled throughout ----
    sustained = None
    if args.sustain > 0:
        per_step = ms/args.steps*1e-3
        nsus = max(args.steps, int(np.ceil(args.sustain/per_step)))
        s2 = ClockSampler(local, period_ms=250)
        if rank == 0:
            s2.start()
        ms_sus = timed(nsus)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"value": Btot*nsus/(ms_sus*1e-3), "unit": UNIT, "seconds": ms_sus*1e-3, "steps": nsus, "clocks": c2}

    if rank == 0:
        pk = peaks()
        nchunk = max(args.steps, 1)
        render_ms = prof["render_ms"]/nchunk
        conv_ms = prof["convolve_ms"]/nchunk
        traffic, pipe_busy = committed_profile(args.workload, B, flags != 0 and model.rays_per_thread == 2)
        halo = 2*(w["psf"].shape[0]//2)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms/args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(w),
                       "objects_dir": OBJECTS_NOTE,
                       "points_per_gpu_per_step": B, "points_per_step": Btot, "rays_per_eval": work["rays"],
                       "parallelism": (f"rows x{world}: every GPU evaluates all {B} points on {H//world} of {H} image rows + "
                                       f"{halo} re-rendered halo rows" if rows_mode else f"points x{world}"),
                       "rays_per_thread": model.rays_per_thread,
                       "math": MATH_TEXT[args.math],
                       "l2": f"working set per step {(B*w['width']*w['height']*4 + 8*w['width']*w['height'])/1e6:.0f} MB of staged "
                             "images > 126 MB L2 (no explicit flush)" if B*w['width']*w['height']*4 > 126e6 else
                             "compute-bound kernel, working set fits L2; inputs re-read from L2 by design"},
            "grays_per_s": value*work["rays"]/1e9,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(P.nbytes), "d2h_bytes_per_step": int(8*Btot if world > 1 else 8*B)},
            "gpu_launches": int(launches),
            "sustained": sustained,
            "stage_ms_per_step": {"set_params": prof["set_params_ms"]/nchunk, "render": render_ms, "convolve_chi2": conv_ms,
                                  "reduce": prof["reduce_ms"]/nchunk},
        }
        if rows_mode:
            # rows actually rendered by a middle rank against the rows it owns
            own = strip[1] - strip[0]
            line["rows"] = {"rows_per_gpu": own, "halo_rows": halo, "predicted_overhead": halo*world/H if world > 1 else 0.0,
                            "rank0_rows_rendered": min(H, strip[1] + halo//2) - max(0, strip[0] - halo//2)}
            render_frac = (min(H, strip[1] + halo//2) - max(0, strip[0] - halo//2))/H
        else:
            render_frac = 1.0
        rl = render_roofline(w, work, B, render_ms, fp32_peak, model, traffic, pipe_busy)
        if rows_mode and rl["achieved"]:
            for k in ("achieved", "frac", "frac_of_unfused_ceiling", "transc_calls_per_s", "grays_per_s_kernel"):
                rl[k] = rl[k]*render_frac if rl[k] is not None else None
            rl["algorithmic"] += f" x {render_frac:.4f} (rank 0's strip + halo)"
        line["roofline"] = rl
        # the 25x25 direct convolution is FP32-bound, not HBM-bound: 2 Pw Ph flop per pixel, issued as
        # separate FMUL + FADD per tap (the reference's rounding, bit-exact against the oracle), so its
        # ceiling is half the FFMA peak; its HBM traffic (16 B/pixel) is a few per cent of the bandwidth
        conv_share = (strip[1] - strip[0])/H
        cf = work["convolve_flops"]*B*conv_share/(conv_ms*1e-3)/1e12 if conv_ms > 0 else None
        line["roofline_convolve"] = {
            "kernel": "lcu_convolve", "bound": "fp32", "achieved": cf, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": cf/fp32_peak if cf and fp32_peak else None,
            "frac_of_unfused_ceiling": cf/(0.5*fp32_peak) if cf and fp32_peak else None,
            "hbm_gbs": work["hbm_bytes"]*B*conv_share/(conv_ms*1e-3)/1e9 if conv_ms > 0 else None,
            "hbm_peak_gbs": pk.get("hbm_gbs"),
        }

        # ---- C5 sub-record in the default line (N = 1): 8 points of the 4096^2 EPL scene ----
        if args.workload == "c4" and world == 1 and not args.no_c5:
            try:
                line["c5"] = c5_subrecord(L, wl, ctx, torch, stream, fp32_peak, args)
            except Exception as e:
                line["c5"] = {"error": repr(e)[:300]}

        # ---- CPU baseline on this box's host cores + parity record, N = 1 only, bounded sample ----
        if world == 1 and not args.no_cpu_baseline:
            try:
                builds = cpu_models(w, image, weight)
                res = []
                for variant, kind, text, om, lib in builds:
                    n, dtc = time_cpu(om, P, args.cpu_seconds/len(builds), max_evals=B)
                    res.append({"value": n/dtc, "unit": UNIT, "cores": lib.orc_max_threads(), "kind": kind, "build": variant,
                                "sample": f"{n} full {w['name']} evaluations on {lib.orc_max_threads()} host threads ({dtc:.1f} s); {text}"})
                res.sort(key=lambda r: -r["value"])
                line["cpu_baseline"] = dict(res[0], others=res[1:])        # the fastest CPU build is the baseline
                line["parity"] = parity_record(model, w, image, weight, P, lnew_dev)
                line["parity_lnew_rel"] = line["parity"]["lnew"][0]["rel_vs_oracle_f32"]
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "unavailable", "sample": repr(e)[:200]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def c5_subrecord(L, wl, ctx, torch, stream, fp32_peak, args, B=8, steps=5, warmup=3):
    """8 points of C5 (4096^2, epl_plus_shear + 3 sersic + sky, g3k7) on this GPU, device-resident."""
    w = wl.c5(4096)
    image, weight = synthetic_observation(w, wl, ctx)
    model = L.Model(ctx, w["objects"], image, weight, rule=w["rule"], psf=w["psf"], flags=math_flags(L, args.math))
    work = wl.work_per_eval(w, model.nq)
    P = wl.param_batch(w, B)
    d_params = torch.from_numpy(P).cuda()
    d_lnew = torch.zeros(B, dtype=torch.float64, device="cuda")
    for _ in range(warmup):
        model.loglike_batch_device(B, d_params.data_ptr(), d_lnew.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    model.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        model.loglike_batch_device(B, d_params.data_ptr(), d_lnew.data_ptr(), stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = model.profile_get()
    model.profile(False)
    render_ms = prof["render_ms"]/steps
    traffic, pipe_busy = committed_profile("c5", B, model.rays_per_thread == 2 and args.math != "strict")
    rec = {"workload": describe(w), "points_per_step": B, "steps": steps, "warmup": warmup, "value": B*steps/(ms*1e-3), "unit": UNIT,
           "ms_per_step": ms/steps, "grays_per_s": B*steps/(ms*1e-3)*work["rays"]/1e9,
           "stage_ms_per_step": {"set_params": prof["set_params_ms"]/steps, "render": render_ms,
                                 "convolve_chi2": prof["convolve_ms"]/steps, "reduce": prof["reduce_ms"]/steps},
           "roofline": render_roofline(w, work, B, render_ms, fp32_peak, model, traffic, pipe_busy)}
    model.close()
    return rec


def parity_record(model, w, image, weight, P, lnew_gpu):
    """The CUDA path against the checker (oracle/, test infrastructure) on the
    benchmark configuration itself: per-pixel relative error of the raw and
    convolved model images at the truth against the strict-float32 oracle, next
    to that oracle's own distance from its float64 twin (the float32 noise
    floor of the scene), and the log-likelihood of the step's first points."""
    from oracle import pyoracle as O
    for v in ("strict", "f64"):
        O.lib(v).orc_set_threads(host_threads())
    qq, ww = O.quad_rule(w["rule"])
    o32 = O.Model(w["objects"], image, weight, qq, ww, psf=w["psf"])
    o64 = O.Model(w["objects"], image, weight, qq, ww, psf=w["psf"], variant="f64")
    out = model.render(w["truth"], error=False, chi=False)
    rec = {"point": "truth", "criterion": "flat: |gpu - o32|/|o32| <= 1e-5 per pixel, <= 1e-6 for lnew (north_star); "
           "floor: |gpu - f64| <= 2 |o32 - f64| (no further from the exact answer than the reference's own float32 arithmetic; "
           "tests/test_gpu_parity.py has the criterion and its ensemble of float32 realisations)"}
    raw32, _ = o32.render(w["truth"])
    raw64, _ = o64.render(w["truth"])
    l32, m32, _ = o32.loglike(w["truth"], want_maps=True)
    l64, m64, _ = o64.loglike(w["truth"], want_maps=True)

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return np.abs(a - b)/np.maximum(np.abs(b), 1e-300)
    for key, g, r32, r64 in (("raw", out["raw"], raw32, raw64), ("model", out["model"], m32, m64)):
        e, fl, e64 = rel(g, r32), rel(r32, r64), rel(g, r64)
        rec[key] = {"max": float(e.max()), "p99.9": float(np.quantile(e, 0.999)), "median": float(np.median(e)),
                    "floor_max": float(fl.max()), "floor_p99.9": float(np.quantile(fl, 0.999)),
                    "gpu_vs_f64_max": float(e64.max()), "gpu_vs_f64_p99.9": float(np.quantile(e64, 0.999)),
                    "flat_ok": bool(e.max() <= 1e-5), "floor_ok": bool(e64.max() <= 2.0*fl.max())}
    pts = [("truth", w["truth"], model.loglike(w["truth"]), l32, l64)]
    for i in range(min(2, len(P))):
        pts.append((f"batch[{i}] (1 % off the truth)", P[i], float(lnew_gpu[i]), o32.loglike(P[i]), o64.loglike(P[i])))
    rec["lnew"] = [{"point": name, "gpu": g, "oracle_f32": a, "oracle_f64": b, "rel_vs_oracle_f32": abs(g - a)/abs(a),
                    "rel_vs_f64": abs(g - b)/abs(b), "floor_rel": abs(a - b)/abs(b), "chi2_per_pixel": -2*b/image.size}
                   for name, _, g, a, b in pts]
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"])
    ap.add_argument("--batch", type=int, default=None,
                    help="weak scaling: parameter points per GPU per step (default 32 for c4); strong scaling and rows: "
                         "points per step in total (default 64 for c5 points, 4 for rows)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: weak for c4, strong for c5 (SURVEY 8d)")
    ap.add_argument("--parallelism", default="points", choices=["points", "rows"])
    ap.add_argument("--math", default="fast", choices=["strict", "fast", "contract"])
    ap.add_argument("--sustain", type=float, default=None, help="seconds of the sustained run after the timed region (0 = none)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 sub-record of the default line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.scaling is None:
        args.scaling = "weak" if args.workload == "c4" else "strong"
    if args.batch is None:
        args.batch = 4 if args.parallelism == "rows" else 32 if args.workload == "c4" else 64 if args.scaling == "strong" else 8
    if args.sustain is None:
        args.sustain = 10.0 if (args.workload == "c4" and args.parallelism == "points") else 0.0
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
