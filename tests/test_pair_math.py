"""Pair math of shim.cuh against the scalar functions it restates, on the CPU.

The two-rays-per-thread render kernel evaluates the plugin text with `float` =
`lcu_pf` (two rays in one 64-bit register, packed FADD2 / FMUL2 / FFMA2).  Its
sqrt / atan / exp / log / atanh -- and, behind -DLCU_PF_LIBM_PAIR=1, atan2 /
sincos / pow / powr -- are written out operation for operation after the scalar
code the one-ray kernel runs (libdevice, or shim.cuh's own fast versions), so
that each lane gets the same bits.  On a GPU that is checked end to end
(test_gpu_parity.py::test_two_rays_per_thread_same_bits).  Here the PTX that
NVRTC emits for both versions is run through a PTX interpreter
(tools/ptx_emu.py: exact IEEE arithmetic with the instruction's rounding mode
and .ftz behaviour, a deterministic stand-in for the special-function unit) on
random, extreme and special arguments and the results are compared bit for
bit.  No device needed: NVRTC compiles for compute_100a anywhere.
"""
import math
import os
import random
import re
import sys
import zlib

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import ptx_emu as E  # noqa: E402

nvrtc = pytest.importorskip("cuda.bindings.nvrtc")

ROOT = os.path.join(os.path.dirname(__file__), "..")

HARNESS = r"""
#include "shim.cuh"
#define K1(name, fn) \
extern "C" __global__ void s_##name(float* o, const float* a) { o[0] = fn(a[0]); } \
extern "C" __global__ void p_##name(float* o, const float* a) { lcu_pf r = fn(lcu_pf(a[0], a[1])); o[0] = r.lo(); o[1] = r.hi(); }
#define K2(name, fn) \
extern "C" __global__ void s_##name(float* o, const float* a, const float* b) { o[0] = fn(a[0], b[0]); } \
extern "C" __global__ void p_##name(float* o, const float* a, const float* b) \
{ lcu_pf r = fn(lcu_pf(a[0], a[1]), lcu_pf(b[0], b[1])); o[0] = r.lo(); o[1] = r.hi(); }
K1(sqrt, sqrt) K1(atan, atan) K1(exp, exp) K1(log, log) K1(atanh, atanh)
K1(fast_exp, lcu_fast_exp) K1(fast_log, lcu_fast_log) K1(fast_atanh, lcu_fast_atanh)
K2(atan2, atan2) K2(powr, powr) K2(pow, pow) K1(sin, sin) K1(cos, cos)
extern "C" __global__ void s_sincos(float* o, const float* a) { float c; o[0] = sincos(a[0], &c); o[1] = c; }
extern "C" __global__ void p_sincos(float* o, const float* a)
{ lcu_pf c; lcu_pf r = sincos(lcu_pf(a[0], a[1]), &c); o[0] = r.lo(); o[1] = c.lo(); o[2] = r.hi(); o[3] = c.hi(); }
"""

# the product's options for object code (lcu_program.cpp: build_options), PTX instead of a cubin
OPTIONS = ["--gpu-architecture=compute_100a", "--std=c++17", "--device-as-default-execution-space",
           "-diag-suppress=177,550", "--ftz=true", "--prec-div=true", "--prec-sqrt=true", "--fmad=false", "-DLCU_FMAD=0"]


def _compile_ptx(extra):
    shim = open(os.path.join(ROOT, "lensed_b200", "kernel", "shim.cuh"), "rb").read()
    err, prog = nvrtc.nvrtcCreateProgram(HARNESS.encode(), b"pair_math.cu", 1, [shim], [b"shim.cuh"])
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
    opts = [o.encode() for o in OPTIONS + list(extra)]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        _, n = nvrtc.nvrtcGetProgramLogSize(prog)
        log = b" "*n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError(log.decode(errors="replace"))
    _, n = nvrtc.nvrtcGetPTXSize(prog)
    ptx = b" "*n
    nvrtc.nvrtcGetPTX(prog, ptx)
    nvrtc.nvrtcDestroyProgram(prog)
    return ptx.decode().rstrip("\0")


@pytest.fixture(scope="module")
def ptx_packed():
    """every pair function in its written-out form"""
    return _compile_ptx(["-DLCU_PF_LIBM_PAIR=1"])


@pytest.fixture(scope="module")
def ptx_default():
    return _compile_ptx([])


@pytest.fixture(scope="module")
def ptx_lanewise():
    """atan2 / sincos / powr of pairs lane by lane through libdevice"""
    return _compile_ptx(["-DLCU_PF_LIBM_PAIR=0"])


SPECIAL = [0x00000000, 0x80000000, 0x3F800000, 0xBF800000, 0x7F800000, 0xFF800000, 0x7FC00000, 0x00000001, 0x80000400,
           0x00800000, 0x7F7FFFFF, 0xFF7FFFFF, 0x40000000, 0x3F000000, 0x40400000, 0xC0400000, 0x3F7FFFFF, 0x3F800001,
           0x47CE4780, 0x47CE477F, 0x4B000000, 0x7E000000, 0x0D000000, 0x0CFFFFFF, 0x42B17218, 0xC2CFF1B5]


def _u(rng, lo, hi):
    return E.f2b(rng.uniform(lo, hi))


def _lu(rng, lo, hi, signed=True):
    v = math.exp(rng.uniform(math.log(lo), math.log(hi)))
    return E.f2b(v*(rng.choice((-1, 1)) if signed else 1))


def _compare(M, name, nin, nout, draws):
    """pair entry on (lane0, lane1) against two runs of the scalar entry"""
    for args in draws:
        lanes = [[a[k] for a in args] for k in (0, 1)]            # args: per input (lane0, lane1)
        want = []
        for k in (0, 1):
            want += M.run("s_" + name, [[v] for v in lanes[k]], nout)
        got = M.run("p_" + name, [list(a) for a in args], 2*nout)
        assert got == want, "%s(%s): pair %s, scalar %s" % (
            name, ", ".join("%08x/%08x" % tuple(a) for a in args), ["%08x" % v for v in got], ["%08x" % v for v in want])
    assert nin == len(draws[0])


def _one_arg_draws(rng, ranges, n):
    draws = []
    for lo, hi in ranges:
        draws += [((_u(rng, lo, hi), _u(rng, lo, hi)),) for _ in range(n)]
    draws += [((rng.choice(SPECIAL), rng.choice(SPECIAL)),) for _ in range(n)]
    draws += [((rng.choice(SPECIAL), _u(rng, *ranges[0])),) for _ in range(n//2)]
    draws += [((_u(rng, *ranges[0]), rng.choice(SPECIAL)),) for _ in range(n//2)]
    return draws


@pytest.mark.parametrize("name,ranges", [
    ("sqrt", [(0.0, 1e6), (0.0, 1e-30)]),
    ("atan", [(-50.0, 50.0), (-1.5, 1.5)]),
    ("exp", [(-100.0, 100.0), (-1.0, 1.0)]),
    ("log", [(1e-6, 1e6), (0.5, 2.0)]),
    ("atanh", [(-0.999, 0.999), (-1e-3, 1e-3)]),
    ("fast_exp", [(-100.0, 100.0), (-1.0, 1.0)]),
    ("fast_log", [(1e-6, 1e6), (0.5, 2.0)]),
    ("fast_atanh", [(-0.999, 0.999), (-1e-3, 1e-3)]),
])
def test_shipped_pair_functions_same_bits(ptx_default, name, ranges):
    """what the default build runs: sqrt / atan and both math modes' exp / log / atanh"""
    M = E.Module(ptx_default)
    _compare(M, name, 1, 1, _one_arg_draws(random.Random(zlib.crc32(name.encode())), ranges, 60))


def test_pair_atan2_same_bits(ptx_packed):
    M = E.Module(ptx_packed)
    rng = random.Random(11)
    draws = [((_u(rng, -10, 10), _u(rng, -10, 10)), (_u(rng, -10, 10), _u(rng, -10, 10))) for _ in range(150)]
    draws += [((_lu(rng, 1e-30, 1e30), _lu(rng, 1e-42, 1e38)), (_lu(rng, 1e-30, 1e30), _lu(rng, 1e-42, 1e38))) for _ in range(150)]
    draws += [((rng.choice(SPECIAL), rng.choice(SPECIAL)), (rng.choice(SPECIAL), rng.choice(SPECIAL))) for _ in range(300)]
    draws += [((rng.choice(SPECIAL), _u(rng, -3, 3)), (_u(rng, -3, 3), rng.choice(SPECIAL))) for _ in range(150)]
    _compare(M, "atan2", 2, 1, draws)


def test_pair_sincos_same_bits(ptx_packed):
    M = E.Module(ptx_packed)
    rng = random.Random(12)
    draws = [((_u(rng, -7, 7), _u(rng, -7, 7)),) for _ in range(200)]
    draws += [((_lu(rng, 1e-40, 1e5), _lu(rng, 1e-10, 1e6)),) for _ in range(200)]
    draws += [((rng.choice(SPECIAL), _u(rng, -4, 4)),) for _ in range(80)]
    draws += [((_u(rng, -4, 4), _lu(rng, 1e5, 1e38)),) for _ in range(80)]      # Payne-Hanek lane next to a plain one
    draws += [((rng.choice(SPECIAL), rng.choice(SPECIAL)),) for _ in range(80)]
    _compare(M, "sincos", 1, 2, draws)


@pytest.mark.parametrize("name", ["sin", "cos"])
def test_pair_sin_cos_same_bits(ptx_packed, name):
    M = E.Module(ptx_packed)
    rng = random.Random(16)
    draws = [((_u(rng, -7, 7), _u(rng, -7, 7)),) for _ in range(200)]
    draws += [((_lu(rng, 1e-40, 1e5), _lu(rng, 1e-10, 1e6)),) for _ in range(200)]
    draws += [((rng.choice(SPECIAL), _u(rng, -4, 4)),) for _ in range(80)]
    draws += [((_u(rng, -4, 4), _lu(rng, 1e5, 1e38)),) for _ in range(80)]
    draws += [((rng.choice(SPECIAL), rng.choice(SPECIAL)),) for _ in range(80)]
    _compare(M, name, 1, 1, draws)


@pytest.mark.parametrize("name", ["powr", "pow"])
def test_pair_pow_same_bits(ptx_packed, name):
    M = E.Module(ptx_packed)
    rng = random.Random(13)
    draws = []
    for _ in range(150):                                                         # the EPL's powr(r, 1 - t): one exponent
        b = _u(rng, -3, 3)
        draws.append(((_lu(rng, 1e-3, 1e4, False), _lu(rng, 1e-3, 1e4, False)), (b, b)))
    draws += [((_lu(rng, 1e-38, 1e38, False), _lu(rng, 1e-30, 1e30, False)), (_lu(rng, 1e-3, 1e3), _lu(rng, 1e-2, 1e2)))
              for _ in range(150)]                                               # overflow / underflow of the result
    draws += [((rng.choice(SPECIAL), rng.choice(SPECIAL)), (rng.choice(SPECIAL), rng.choice(SPECIAL))) for _ in range(300)]
    draws += [((rng.choice(SPECIAL), _lu(rng, 0.1, 3, False)), (_u(rng, -3, 3), rng.choice(SPECIAL))) for _ in range(150)]
    draws += [((_lu(rng, 0.5, 2), _u(rng, -5, 5)), (_u(rng, -2, 2), _u(rng, -5, 5))) for _ in range(100)]   # negative bases
    _compare(M, name, 2, 1, draws if name == "powr" else draws[::3])      # pow is the same function


def test_interpreter_agrees_with_libm(ptx_packed):
    """the interpreted scalar functions are the functions: within 2 ulp of libm
    (atan2f and the plain path of sincosf use no special-function unit, so the
    interpreter reproduces libdevice exactly there)"""
    M = E.Module(ptx_packed)
    rng = random.Random(14)
    for _ in range(100):
        y, x = E.b2f(_u(rng, -5, 5)), E.b2f(_u(rng, -5, 5))
        got = E.b2f(M.run("s_atan2", [[E.f2b(y)], [E.f2b(x)]], 1)[0])
        assert abs(got - math.atan2(y, x)) <= 2.5e-7*abs(math.atan2(y, x)) + 1e-37
        t = E.b2f(_u(rng, -7, 7))
        s, c = (E.b2f(v) for v in M.run("s_sincos", [[E.f2b(t)]], 2))
        assert abs(s - math.sin(t)) <= 1.5e-7 and abs(c - math.cos(t)) <= 1.5e-7
        a, b = E.b2f(_lu(rng, 1e-2, 50, False)), E.b2f(_u(rng, -3, 3))
        got = E.b2f(M.run("s_powr", [[E.f2b(a)], [E.f2b(b)]], 1)[0])
        assert abs(got - a**b) <= 2.5e-7*a**b
    for t in (1e6, 123456.789, 3e9):                                             # Payne-Hanek reduction, integer code
        tb = E.f2b(t)
        s, c = (E.b2f(v) for v in M.run("s_sincos", [[tb]], 2))
        assert abs(s - math.sin(E.b2f(tb))) <= 1.5e-7 and abs(c - math.cos(E.b2f(tb))) <= 1.5e-7


def test_comparison_is_sensitive(ptx_packed):
    """negative control: one polynomial coefficient of the pair code off by one
    ulp (the scalar code untouched) is caught"""
    for name, const, nin in (("atan2", "0f419D92C8", 2), ("powr", "0f3F317218", 2), ("sincos", "0f3D2AAABB", 1)):
        entry = re.search(r"\.visible\s+\.entry\s+p_%s\b.*?\n\}" % name, ptx_packed, re.S)
        assert entry and const in entry.group(0), (name, const)
        bumped = "0f%08X" % (int(const[2:], 16) + 1)
        M = E.Module(ptx_packed.replace(entry.group(0), entry.group(0).replace(const, bumped)))
        rng = random.Random(15)
        differs = 0
        for _ in range(40):
            args = [(_lu(rng, 0.3, 3, False), _lu(rng, 0.3, 3, False)) for _ in range(nin)]
            nout = 2 if name == "sincos" else 1
            want = []
            for k in (0, 1):
                want += M.run("s_" + name, [[a[k]] for a in args], nout)
            differs += M.run("p_" + name, [list(a) for a in args], 2*nout) != want
        assert differs > 0, name


def test_default_build_is_the_packed_libm(ptx_default, ptx_packed, ptx_lanewise):
    """the switch is on by default: atan2 / sincos / powr of pairs are packed
    instructions unless -DLCU_PF_LIBM_PAIR=0 asks for the lane-by-lane libdevice calls"""
    for name in ("atan2", "sincos", "powr", "sin", "cos"):
        for ptx, packed in ((ptx_default, True), (ptx_packed, True), (ptx_lanewise, False)):
            entry = re.search(r"\.visible\s+\.entry\s+p_%s\b.*?\n\}" % name, ptx, re.S).group(0)
            assert ("f32x2" in entry) == packed, (name, packed)


def test_interpreter_fast_path_is_the_exact_path():
    """add / sub / mul / fma of normal numbers go through binary64 when that is
    provably exact; the result must be what exact rational arithmetic gives
    (near-cancellations, operands far apart, results next to the subnormal and
    overflow thresholds included)"""
    rng = random.Random(17)

    def draw():
        r = rng.random()
        if r < 0.4:
            return _u(rng, -4, 4)
        if r < 0.7:
            return _lu(rng, 1e-38, 3e38)
        if r < 0.9:
            return rng.getrandbits(32)
        return rng.choice(SPECIAL)
    fast = 0
    for _ in range(10000):
        a, b, c = draw(), draw(), draw()
        if rng.random() < 0.2:
            b = a ^ 0x80000000 ^ rng.getrandbits(3)                  # near-cancellation in add, ties in fma
        for op in ("add", "sub", "mul", "fma"):
            cc = c if op == "fma" else None
            for flush in (False, True):
                want = E._arith_exact(op, a, b, cc, "rn", flush)
                got = E._arith(op, a, b, cc, "rn", flush)
                assert got == want, (op, hex(a), hex(b), hex(c), flush, hex(got), hex(want))
            fast += E._fast_rn(op, a, b, cc) is not None
    assert fast > 10000                                              # the fast path does take most ordinary cases


RACY = """
.visible .entry k_race(.param .u64 k_race_param_0)
{
    .shared .align 4 .b8 buf[8];
    ld.param.u64 %rd1, [k_race_param_0];
    mov.u32 %r1, %tid.x;
    mov.u32 %r2, buf;
    st.shared.u32 [buf], %r1;
    RACE_BAR
    ld.shared.u32 %r3, [buf];
    st.global.u32 [%rd1], %r3;
    ret;
}
.visible .entry k_oob(.param .u64 k_oob_param_0)
{
    ld.param.u64 %rd1, [k_oob_param_0];
    ld.global.u32 %r1, [%rd1+4096];
    st.global.u32 [%rd1], %r1;
    ret;
}
"""


def test_interpreter_flags_races_and_unwritten_loads():
    """negative controls for the two checks the interpreted kernel tests rely on:
    two threads writing one shared word without a barrier, and a load from an
    address nobody wrote, both raise"""
    M = E.Module(RACY.replace("RACE_BAR", ""))
    with pytest.raises(RuntimeError, match="shared-memory hazard"):
        M.launch("k_race", (1,), 2, [0x1000], E.StrictMemory())
    M.launch("k_race", (1,), 1, [0x1000], E.StrictMemory())          # one thread: nothing to race with
    M.launch("k_race", (1,), 2, [0x1000], E.StrictMemory(), racecheck=False)
    with pytest.raises(KeyError, match="unwritten address"):
        M.launch("k_oob", (1,), 1, [0x1000], E.StrictMemory())
    mem = {}
    M.launch("k_oob", (1,), 1, [0x1000], mem)                        # a plain dict reads zeros
    assert mem[0x1000] == 0
