#!/usr/bin/env python
"""End-to-end run of the batch-native nested sampler on one of the reference's
examples (inputs from tests/golden/examples.npz: image, PSF, gain / offset,
objects and priors of examples/<name>.ini): posterior of all free parameters,
evidence, wall time and likelihood evaluations per second through
lcu_loglike_batch.  Also reports the batched throughput of the model at a few
batch sizes (the number a sampler sees), next to the single-point latency.

    python tools/sample_example.py full_mock_psf --nlive 300 --batch 256
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")      # the reference's objects/*.cl, verbatim
sys.path.insert(0, ROOT)

import lensed_b200 as L                                     # noqa: E402
from lensed_b200 import host as Hh, sampler as S, workloads  # noqa: E402

# input parameters quoted in the comments of the example files (source position in the source plane)
TRUTH = {
    "full_mock_psf": {"lens.x": 50.5, "lens.y": 50.5, "lens.r": 17.44, "lens.q": 0.75, "lens.pa": 45.0,
                      "source.r": 3.90, "source.mag": -3.07, "source.n": 3.18, "source.q": 0.89, "source.pa": 30.0},
}
TRUTH["full_mock_nopsf"] = TRUTH["full_mock_psf"]


def build(ctx, name, flags):
    with np.load(os.path.join(ROOT, "tests", "golden", "examples.npz")) as z:
        meta = json.loads(str(z["meta"]))[name]
        img = z[name + "_image"]
        psf = z[name + "_psf"] if meta["psf"] else None
    objects = []
    for oid, oname in meta["objects"]:
        info = ctx.object_info(oname)
        obj = Hh.ObjectEntry(oid, oname, info.type)
        for p in info.params:
            par = Hh.Parameter(f"{oid}.{p.name}", p.name, p.type, p.bounds[0], p.bounds[1])
            words = meta["priors"][par.id].split()
            while words and words[0] in ("wrap", "image"):
                if words[0] == "wrap":
                    par.wrap = True
                else:
                    par.ipp = True
                words.pop(0)
            par.prior = Hh.read_prior(" ".join(words))
            obj.params.append(par)
        objects.append(obj)
    cfg = Hh.Config({}, objects)
    weight = workloads.make_weight(img, meta["gain"], meta["offset"])
    model = L.Model(ctx, [o.name for o in objects], img, weight, rule=meta["rule"],
                    psf=workloads.normalise_psf(psf) if psf is not None else None, pcs=tuple(meta["pcs"]),
                    ipp=[[int(p.ipp) for p in o.params] for o in objects], flags=flags)
    return cfg, model, Hh.Likelihood(cfg, model)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", nargs="?", default="full_mock_psf")
    ap.add_argument("--nlive", type=int, default=300)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--tol", type=float, default=0.1)
    ap.add_argument("--eff", type=float, default=0.8)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--maxiter", type=int, default=0)
    ap.add_argument("--math", default="fast", choices=["strict", "fast"])
    ap.add_argument("--out", default=None, help="root for MultiNest-layout result files")
    args = ap.parse_args()

    ctx = L.Context(device=0, objects_dir=OBJECTS_DIR)
    flags = 0 if args.math == "strict" else (L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
    cfg, model, like = build(ctx, args.name, flags)
    rng = np.random.default_rng(0)

    # what a sampler sees: evaluations per second as a function of the batch size
    thr = {}
    for nb in (1, 16, 256, 1024):
        cubes = rng.random((nb, like.ndims))
        P = np.stack([like.device_params(like.physical(c)) for c in cubes])
        model.loglike_batch(P)
        reps = max(3, min(200, 4096//nb))
        t0 = time.perf_counter()
        for _ in range(reps):
            model.loglike_batch(P) if nb > 1 else model.loglike(P[0])
        thr[nb] = nb*reps/(time.perf_counter() - t0)

    t0 = time.perf_counter()
    res = S.run(like, nlive=args.nlive, batch=args.batch, tol=args.tol, eff=args.eff, seed=args.seed, maxiter=args.maxiter)
    dt = time.perf_counter() - t0
    if args.out:
        S.write_multinest(args.out, res, labels=[like.pars[i].id for i in like.pmap])

    mean, std, ml = res.mean(), res.std(), res.max_like()
    truth = TRUTH.get(args.name, {})
    rows = []
    for k, i in enumerate(like.pmap):
        pid = like.pars[i].id
        rows.append({"param": pid, "mean": float(mean[k]), "sigma": float(std[k]), "max_like": float(ml[k]), "truth": truth.get(pid)})
    out = {
        "example": args.name, "ndims": like.ndims, "image": list(model_shape(model)), "rays_per_thread": model.rays_per_thread,
        "nlive": args.nlive, "batch": args.batch, "tol": args.tol, "seed": args.seed, "math": args.math,
        "logz": res.logz, "logz_err": res.logz_err, "information": res.information, "max_loglike": float(res.loglike.max()),
        "dead_points": res.niter, "evaluations": res.nevals, "launch_batches": res.nbatches, "efficiency": res.efficiency,
        "wall_s": dt, "evals_per_s_sampler": res.nevals/dt, "evals_per_s_by_batch": thr, "posterior": rows,
    }
    print(json.dumps(out))
    for r in rows:
        t = "" if r["truth"] is None else f"   truth {r['truth']:g}"
        print(f"# {r['param']:12s} {r['mean']:10.4f} +- {r['sigma']:.4f}   ML {r['max_like']:10.4f}{t}", file=sys.stderr)
    print(f"# lnZ = {res.logz:.2f} +- {res.logz_err:.2f}; {res.nevals} evaluations in {res.nbatches} launches, {dt:.1f} s "
          f"({res.nevals/dt:.0f} evals/s, efficiency {res.efficiency:.3f})", file=sys.stderr)


def model_shape(model):
    return model.height, model.width


if __name__ == "__main__":
    main()
