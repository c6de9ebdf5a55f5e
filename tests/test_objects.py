"""Object metadata read from the NVRTC-compiled module (no device) against
the oracle catalogue and, where the reference tree is present, against the
reference's own object files consumed unmodified."""
import os

import numpy as np
import pytest

import helpers as H
import lensed_b200 as L
from oracle import pyoracle as O

NAMES = O.object_names()
REF_OBJECTS = os.path.join(os.environ.get("LENSED_REFERENCE", "/root/reference"), "objects")
# sizeof(data) in 4-byte words, SURVEY.md section 8
WORDS = dict(sis=4, nsis=4, point_mass=4, sky=4, sis_plus_shear=12, sersic=12, devauc=12, exponential=12, gauss=12,
             sie=16, nsie=16, epl=16, sie_plus_shear=20, epl_plus_shear=20)
WORDS["sersic-old"] = 12


def _same(info, o):
    assert info.type == o["type"] and info.words == o["words"] and info.npars == o["npar"]
    for p, q in zip(info.params, o["params"]):
        assert p.name == q["name"] and p.type == q["type"]
        assert p.bounds == tuple(q["bounds"])
        assert np.float32(p.defval).view(np.uint32) == q["defval_bits"]      # -0.0f survives


@pytest.mark.parametrize("name", NAMES)
def test_shipped_objects(compile_ctx, name):
    info = compile_ctx.object_info(name)
    _same(info, O.object_info(name))
    assert info.words == WORDS[name]


def test_sky_gradient_defaults(compile_ctx):
    p = compile_ctx.object_info("sky").params
    assert [q.has_default for q in p] == [False, True, True]
    assert np.signbit(np.float32(p[1].defval)) and p[1].defval == 0


def test_fixture_objects_are_the_reference_files():
    """tests/golden/objects/ -- the plugin directory every test, smoke() and
    bench.py run on -- holds the reference's objects/*.cl byte for byte: checked
    against the committed SHA-256 list everywhere and against the reference tree
    where it exists."""
    import hashlib
    sums = dict(reversed(l.split()) for l in open(os.path.join(H.OBJECTS_DIR, "SHA256SUMS")))
    files = sorted(f for f in os.listdir(H.OBJECTS_DIR) if f.endswith(".cl"))
    assert files == sorted(sums) and len(files) == 15
    for f in files:
        data = open(os.path.join(H.OBJECTS_DIR, f), "rb").read()
        assert hashlib.sha256(data).hexdigest() == sums[f], f
        if os.path.isdir(REF_OBJECTS):
            assert data == open(os.path.join(REF_OBJECTS, f), "rb").read(), f
    if os.path.isdir(REF_OBJECTS):
        assert files == sorted(f for f in os.listdir(REF_OBJECTS) if f.endswith(".cl"))


@pytest.mark.skipif(not os.path.isdir(REF_OBJECTS), reason="reference tree not present")
def test_same_device_code_from_the_reference_tree_and_from_the_fixture_directory():
    """The program text and the sm_100a image of the C4 and C5 models are the same
    bytes whether the plugin directory is /root/reference/objects or the committed
    copy every GPU run uses: what the B200 ran is what the reference's files give."""
    from lensed_b200 import workloads
    img = np.zeros((64, 64), np.float32)
    a, b = L.Context(device=-1, objects_dir=REF_OBJECTS), L.Context(device=-1, objects_dir=H.OBJECTS_DIR)
    for w in (workloads.c4(64), workloads.c5(64)):
        ma = L.Model(a, w["objects"], img, img, rule=w["rule"], psf=w["psf"], flags=L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
        mb = L.Model(b, w["objects"], img, img, rule=w["rule"], psf=w["psf"], flags=L.LCU_FAST_INTRINSICS | L.LCU_FAST_ATANH)
        assert ma.source == mb.source and ma.cubin == mb.cubin and len(ma.cubin) > 0
    a.close(); b.close()


def test_library_ships_no_object_files(monkeypatch, tmp_path):
    """No objects_dir and no LENSED_PATH: the first object load fails in the
    reference's words (src/kernel.c:757-759); LENSED_PATH/objects is found as the
    reference finds it (src/kernel.c:11-13)."""
    assert not os.path.isdir(os.path.join(os.path.dirname(L.__file__), "objects"))
    monkeypatch.delenv("LENSED_PATH", raising=False)
    ctx = L.Context(device=-1)
    with pytest.raises(L.LensedCudaError, match='could not load object "sersic"'):
        ctx.object_info("sersic")
    ctx.close()
    monkeypatch.setenv("LENSED_PATH", os.path.dirname(H.OBJECTS_DIR))
    ctx = L.Context(device=-1)
    _same(ctx.object_info("sersic"), O.object_info("sersic"))
    ctx.close()


@pytest.mark.skipif(not O.available("ref"), reason="oracle/_ref not built")
def test_oracle_catalogue_matches_reference_meta_kernels():
    """The reference's own meta_<name>/params_<name> kernels, run on the host."""
    for name in O.object_names("ref"):
        assert O.object_info(name, "ref") == O.object_info(name)
