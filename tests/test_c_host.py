"""The C ABI from C: tests/c/host_check.c is compiled with gcc against
include/lensed_cuda.h, linked to liblensed_cuda.so and run -- the binding the
reference's C host would use."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import pyoracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the C host passes objects_dir = NULL: the library looks in $LENSED_PATH/objects, as the
# reference does (src/kernel.c:11-13); here that is the verbatim copy of the reference's objects/
ENV = dict(os.environ, LENSED_PATH=os.path.dirname(H.OBJECTS_DIR))


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c") / "host_check")
    lib = os.path.join(ROOT, "lensed_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "host_check.c"), "-L", lib, "-llensed_cuda", f"-Wl,-rpath,{lib}",
                    "-lm", "-o", exe], check=True)
    return exe


def test_metadata_from_c(host_check):
    out = subprocess.run([host_check, "meta"], check=True, capture_output=True, text=True, env=ENV).stdout.splitlines()
    objs = [l.split() for l in out if l.startswith("object")]
    assert len(objs) == 15
    for f in objs:
        o = O.object_info(f[1])
        assert (f[2], int(f[3]), int(f[4])) == (o["type"], o["words"], o["npar"])
        for tok, p in zip(f[5:], o["params"]):
            name, typ, lo, hi, dflt = tok.split(":")
            assert name == p["name"] and int(typ) == p["type"]
            assert float(lo) == pytest.approx(p["bounds"][0]) and float(hi) == pytest.approx(p["bounds"][1], rel=1e-5)
            assert int(dflt) == int(p["defval"] > 0 or bool(np.signbit(np.float32(p["defval"]))))
    rules = {l.split()[1]: int(l.split()[2]) for l in out if l.startswith("rule")}
    assert rules == dict(point=1, sub2=4, sub4=16, gm75=17, g3k7=49, g5k11=121, g7k15=225)
    assert any(l.startswith('error could not load object "nonesuch"') for l in out)


@pytest.mark.gpu
def test_loglike_from_c(host_check):
    size = 64
    out = subprocess.run([host_check, "loglike", "0", str(size)], check=True, capture_output=True, text=True, env=ENV).stdout.splitlines()
    assert out[0] == "model npars 12 words 28"
    lnew = [float(v) for v in out[1].split()[1:]]
    assert lnew[0] == lnew[1]                                  # single == first of the batch
    # the same model through the oracle
    c = 0.5*(size + 1)
    weight = (1.0 + (np.arange(size*size) % 7)).astype(np.float32).reshape(size, size)
    psf = np.array([[0.05, 0.1, 0.05], [0.1, 0.4, 0.1], [0.05, 0.1, 0.05]], np.float32)
    for b in range(3):
        params = np.array([c, c, 0.2*size, 0.75 + 0.05*b, 45, c + 0.25*size, c + 0.1*size, 0.04*size, -3, 2, 0.8, 30], np.float32)
        cfg = H.Config("c-host", ["sie", "sersic"], params, np.zeros((size, size), np.float32), weight, psf=psf,
                       ipp=[[0]*5, [1, 1, 0, 0, 0, 0, 0]])
        ref = cfg.oracle().loglike(params)
        assert abs(lnew[1 + b] - ref) <= 3e-6*abs(ref), (b, lnew[1 + b], ref)
    flux = float(out[2].split()[1])
    assert flux > 0 and int(out[2].split()[3]) > 0
    # device-side weight map (image is all zeros here: gain/offset everywhere, 0 at the masked pixel)
    prep = out[3].split()
    assert prep[0] == "prep" and int(prep[1]) == 2
    assert np.float32(float(prep[2])) == np.float32(np.float32(1800.0)/np.float64(2.9633)) and float(prep[3]) == 0.0
    assert float(prep[4]) != lnew[0]
    assert out[4].split()[0] == "back" and float(out[4].split()[1]) == lnew[0]          # original weights restored


@pytest.mark.gpu
def test_single_point_latency_from_c(host_check):
    """The sampler's callback pattern from a C host: one point per call on a
    100 x 100 image (the reference's example size).  The reference's CPU build
    takes ~2 ms per call on 16 threads; anything above 200 us here means the
    graph path is not in use."""
    out = subprocess.run([host_check, "latency", "0", "100", "2000"], check=True, capture_output=True, text=True, env=ENV).stdout
    print(out)
    us = float(out.split()[1])
    assert 0 < us < 200
    # two in flight: same checksum (same bits in the same order), and no slower per point
    second = out.splitlines()[1].split()
    assert second[0] == "pipelined" and second[-1] == "same"
    assert 0 < float(second[1]) < 200
