"""Object metadata read from the NVRTC-compiled module (no device) against
the oracle catalogue and, where the reference tree is present, against the
reference's own object files consumed unmodified."""
import os

import numpy as np
import pytest

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


@pytest.mark.skipif(not os.path.isdir(REF_OBJECTS), reason="reference tree not present")
@pytest.mark.parametrize("name", NAMES)
def test_reference_object_files_unmodified(name):
    """Drop-in: the reference's objects/ directory compiles as is."""
    ctx = L.Context(device=-1, objects_dir=REF_OBJECTS)
    _same(ctx.object_info(name), O.object_info(name))


@pytest.mark.skipif(not O.available("ref"), reason="oracle/_ref not built")
def test_oracle_catalogue_matches_reference_meta_kernels():
    """The reference's own meta_<name>/params_<name> kernels, run on the host."""
    for name in O.object_names("ref"):
        assert O.object_info(name, "ref") == O.object_info(name)
