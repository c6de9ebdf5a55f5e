"""A small PTX interpreter for the device code of this repository: float32 ray
functions, float64 setters, and whole kernels as grids of thread blocks
(Module.launch: bar.sync, shfl.sync, shared / constant / parameter memory).

Test infrastructure (CPU only): it lets `tests/test_pair_math.py` run the PTX
that NVRTC emits for shim.cuh's pair functions (`lcu_pf`: the same quantity for
two rays in one 64-bit register, packed `.f32x2` arithmetic) next to the PTX of
the scalar function each of them restates (libdevice's expf / logf / atanf /
atanhf / atan2f / sincosf / powf as inlined by NVRTC), on the same inputs, and
compare the results bit for bit -- without a GPU.

Arithmetic follows the PTX ISA: IEEE-754 binary32 with the rounding mode the
instruction names (exact rational arithmetic + one rounding), `.ftz` flushes
subnormal inputs and results to signed zero, NaN results are the canonical
0x7fffffff.  The special-function-unit approximations (`rcp/rsqrt/ex2/lg2
.approx`) are NOT the hardware's tables: they are a deterministic stand-in
(correctly rounded result with the last bit flipped by a hash of the argument
bits), which is all a comparison of two instruction streams needs -- both must
hand the unit the same bits to get the same bits back.  `sqrt.rn`, `div.rn` and
`rcp.rn` are exact IEEE operations, as in PTX.

Only the instructions that occur in those functions are implemented; anything
else raises NotImplementedError with the offending line.
"""
from __future__ import annotations

import math
import re
import struct
from fractions import Fraction

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
CANON_NAN = 0x7FFFFFFF


# ---------------------------------------------------------------------------
# binary32 helpers on bit patterns
# ---------------------------------------------------------------------------
def f2b(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def b2f(b: int) -> float:
    return struct.unpack("<f", struct.pack("<I", b & M32))[0]


def d2b(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def b2d(b: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", b & M64))[0]


def is_nan(b: int) -> bool:
    return (b & 0x7F800000) == 0x7F800000 and (b & 0x007FFFFF) != 0


def is_inf(b: int) -> bool:
    return (b & 0x7FFFFFFF) == 0x7F800000


def is_zero(b: int) -> bool:
    return (b & 0x7FFFFFFF) == 0


def ftz(b: int) -> int:
    """flush a subnormal bit pattern to zero of the same sign"""
    if (b & 0x7F800000) == 0:
        return b & 0x80000000
    return b


def to_frac(b: int) -> Fraction:
    s = -1 if b & 0x80000000 else 1
    e = (b >> 23) & 0xFF
    m = b & 0x007FFFFF
    if e == 0:
        return Fraction(s * m, 1 << 149)
    return s * Fraction((1 << 23) | m) * Fraction(2) ** (e - 150)


def round_frac64(x: Fraction, mode: str = "rn", neg_zero: bool = False) -> int:
    """exact rational -> binary64 bits (round to nearest even; other modes as binary32's)"""
    if x == 0:
        return (1 << 63) if neg_zero else 0
    sign = (1 << 63) if x < 0 else 0
    a = -x if x < 0 else x
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    elif Fraction(2) ** (e + 1) <= a:
        e += 1
    q = max(e, -1022) - 52
    scaled = a / (Fraction(2) ** q)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    up = False
    if rem != 0:
        if mode == "rn":
            up = rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1))
        elif mode == "rm":
            up = bool(sign)
        elif mode == "rp":
            up = not sign
    if up:
        n += 1
    if n >= (1 << 53):
        n >>= 1
        q += 1
    if n < (1 << 52):
        return sign | n
    ebits = q + 52 + 1023
    if ebits >= 2047:
        return sign | (0x7FF << 52)
    return sign | (ebits << 52) | (n & ((1 << 52) - 1))


def to_frac64(b: int) -> Fraction:
    s = -1 if b >> 63 else 1
    e = (b >> 52) & 0x7FF
    m = b & ((1 << 52) - 1)
    if e == 0:
        return Fraction(s * m, 1 << 1074)
    return s * Fraction((1 << 52) | m) * Fraction(2) ** (e - 1075)


def is_nan64(b: int) -> bool:
    return (b >> 52) & 0x7FF == 0x7FF and (b & ((1 << 52) - 1)) != 0


def is_inf64(b: int) -> bool:
    return (b & ~(1 << 63)) == (0x7FF << 52)


def fma64(a: int, b: int, c: int) -> int:
    """fma.rn.f64 on bit patterns: exact product and sum, one rounding"""
    if is_nan64(a) or is_nan64(b) or is_nan64(c) or is_inf64(a) or is_inf64(b) or is_inf64(c):
        x, y, z = b2d(a), b2d(b), b2d(c)
        try:
            return d2b(x*y + z)                    # IEEE special values: the unfused expression gives the same
        except OverflowError:
            return d2b(float("nan"))
    p = to_frac64(a) * to_frac64(b)
    r = p + to_frac64(c)
    nz = False
    if r == 0:
        ps = (a ^ b) >> 63
        nz = bool(ps and (c >> 63)) if p == 0 and (c & ~(1 << 63)) == 0 else False
    return round_frac64(r, "rn", nz)


def round_frac(x: Fraction, mode: str = "rn", flush: bool = False, neg_zero: bool = False) -> int:
    """exact rational -> binary32 bits under rounding mode rn / rz / rm / rp"""
    if x == 0:
        return 0x80000000 if neg_zero else 0
    sign = 0x80000000 if x < 0 else 0
    a = -x if x < 0 else x
    # exponent e with 2^e <= a < 2^(e+1)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    elif Fraction(2) ** (e + 1) <= a:
        e += 1
    q = max(e, -126) - 23                         # weight of the last mantissa bit
    scaled = a / (Fraction(2) ** q)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    up = False
    if rem != 0:
        if mode == "rn":
            up = rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1))
        elif mode == "rz":
            up = False
        elif mode == "rm":
            up = bool(sign)
        elif mode == "rp":
            up = not sign
        else:
            raise NotImplementedError(mode)
    if up:
        n += 1
    if n >= (1 << 24):
        n >>= 1
        q += 1
    if n < (1 << 23):                             # subnormal (q == -149)
        bits = sign | n
        return (bits & 0x80000000) if flush else bits
    ebits = q + 23 + 127
    if ebits >= 255:
        if mode == "rz" or (mode == "rm" and not sign) or (mode == "rp" and sign):
            return sign | 0x7F7FFFFF
        return sign | 0x7F800000
    return sign | (ebits << 23) | (n & 0x007FFFFF)


_MIN_NORMAL = 2.0 ** -126


def _normal(b: int) -> bool:
    return 0 < (b & 0x7F800000) < 0x7F800000


def _fast_rn(op: str, a: int, b: int, c):
    """add / sub / mul / fma in round-to-nearest for normal operands and a normal
    result, through binary64: the product of two binary32 numbers is exact in
    binary64, and a sum is used only when the error-free transformation says
    it is exact too -- then the single rounding to binary32 is the operation's.
    Returns None when the case is not covered (the exact rational path takes it)."""
    if not (_normal(a) and _normal(b) and (c is None or _normal(c))):
        return None
    x, y = b2f(a), b2f(b)
    if op == "mul":
        r = x*y
    else:
        if op == "fma":
            x, y = x*y, b2f(c)
        elif op == "sub":
            y = -y
        r = x + y
        t = r - x
        if (x - (r - t)) + (y - t) != 0.0:
            return None
    if not _MIN_NORMAL <= abs(r) < 3.4e38:
        return None                                # zero, subnormal or overflowing results: signs, flushing, infinities
    bits = f2b(r)
    return bits if _normal(bits) else None


def _arith(op: str, a: int, b: int, c: int | None, mode: str, flush: bool) -> int:
    """add / sub / mul / fma on bit patterns"""
    if mode == "rn":
        r = _fast_rn(op, a, b, c)
        if r is not None:
            return r
    return _arith_exact(op, a, b, c, mode, flush)


def _arith_exact(op: str, a: int, b: int, c: int | None, mode: str, flush: bool) -> int:
    """the same through exact rational arithmetic: every rounding mode, special values, signed zeros, flushing"""
    if flush:
        a, b = ftz(a), ftz(b)
        if c is not None:
            c = ftz(c)
    if op == "sub":
        b ^= 0x80000000
        op = "add"
    if is_nan(a) or is_nan(b) or (c is not None and is_nan(c)):
        return CANON_NAN
    if op == "add":
        if is_inf(a) or is_inf(b):
            if is_inf(a) and is_inf(b) and (a ^ b) & 0x80000000:
                return CANON_NAN
            return a if is_inf(a) else b
        r = to_frac(a) + to_frac(b)
        nz = False
        if r == 0:                                # signed zero of an exact cancellation
            if is_zero(a) and is_zero(b):
                nz = bool(a & b & 0x80000000) if mode != "rm" else bool((a | b) & 0x80000000)
            else:
                nz = mode == "rm"
        return round_frac(r, mode, flush, nz)
    if op == "mul":
        sign = (a ^ b) & 0x80000000
        if is_inf(a) or is_inf(b):
            if is_zero(a) or is_zero(b):
                return CANON_NAN
            return sign | 0x7F800000
        return round_frac(to_frac(a) * to_frac(b), mode, flush, bool(sign))
    if op == "fma":
        psign = (a ^ b) & 0x80000000
        if is_inf(a) or is_inf(b):
            if is_zero(a) or is_zero(b):
                return CANON_NAN
            if is_inf(c) and (c ^ psign) & 0x80000000:
                return CANON_NAN
            return psign | 0x7F800000
        if is_inf(c):
            return c
        p = to_frac(a) * to_frac(b)
        r = p + to_frac(c)
        nz = False
        if r == 0:
            if p == 0 and is_zero(c):
                nz = bool(psign and (c & 0x80000000)) if mode != "rm" else bool(psign or (c & 0x80000000))
            else:
                nz = mode == "rm"
        return round_frac(r, mode, flush, nz)
    raise NotImplementedError(op)


def _hash_bit(b: int) -> int:
    return ((b * 2654435761) >> 13) & 1


def _approx(kind: str, a: int, flush: bool) -> int:
    """deterministic stand-in for a special-function-unit approximation"""
    if flush:
        a = ftz(a)
    if is_nan(a):
        return CANON_NAN
    x = b2f(a)
    neg = bool(a & 0x80000000)
    if kind == "rcp":
        if is_zero(a):
            return (a & 0x80000000) | 0x7F800000
        if is_inf(a):
            return a & 0x80000000
        r = round_frac(1 / to_frac(a), "rn", flush)
    elif kind == "rsqrt":
        if is_zero(a):
            return (a & 0x80000000) | 0x7F800000
        if neg:
            return CANON_NAN
        if is_inf(a):
            return 0
        r = f2b(1.0 / math.sqrt(x))
    elif kind == "ex2":
        if is_inf(a):
            return 0 if neg else 0x7F800000
        if x > 128.0:
            return 0x7F800000
        if x < -150.0:
            return 0
        r = f2b(2.0 ** x)
    elif kind == "lg2":
        if is_zero(a):
            return 0xFF800000
        if neg:
            return CANON_NAN
        if is_inf(a):
            return a
        r = f2b(math.log2(x))
    else:
        raise NotImplementedError(kind)
    if flush:
        r = ftz(r)
    if (r & 0x7F800000) not in (0, 0x7F800000):   # keep zeros / infinities / subnormals as they are
        r ^= _hash_bit(a)
    return r


def _ieee(kind: str, a: int, b: int | None, flush: bool) -> int:
    """sqrt.rn / div.rn / rcp.rn"""
    if flush:
        a = ftz(a)
        if b is not None:
            b = ftz(b)
    if is_nan(a) or (b is not None and is_nan(b)):
        return CANON_NAN
    if kind == "sqrt":
        if is_zero(a):
            return a
        if a & 0x80000000:
            return CANON_NAN
        if is_inf(a):
            return a
        fr = to_frac(a)
        # integer square root at 2x48 extra bits decides the rounding exactly
        k = 160
        n = (fr * Fraction(2) ** (2 * k))
        root = math.isqrt(n.numerator // n.denominator)
        exact = root * root == n and n.denominator == 1
        r = Fraction(root, 1 << k)
        if not exact:
            r += Fraction(1, 1 << (k + 2))        # sticky: strictly between root and root + 1
        return round_frac(r, "rn", flush)
    if kind == "rcp":
        b = a
        a = 0x3F800000
    sign = (a ^ b) & 0x80000000
    if is_inf(a):
        return CANON_NAN if is_inf(b) else sign | 0x7F800000
    if is_inf(b):
        return sign
    if is_zero(b):
        return CANON_NAN if is_zero(a) else sign | 0x7F800000
    return round_frac(to_frac(a) / to_frac(b), "rn", flush, bool(sign))


def _cmp_f(op: str, a: int, b: int, flush: bool) -> bool:
    if flush:
        a, b = ftz(a), ftz(b)
    nan = is_nan(a) or is_nan(b)
    if op == "num":
        return not nan
    if op == "nan":
        return nan
    base = {"eq": "eq", "ne": "ne", "lt": "lt", "le": "le", "gt": "gt", "ge": "ge",
            "equ": "eq", "neu": "ne", "ltu": "lt", "leu": "le", "gtu": "gt", "geu": "ge"}[op]
    if nan:
        return op in ("equ", "neu", "ltu", "leu", "gtu", "geu")
    x, y = to_frac(a), to_frac(b)
    return {"eq": x == y, "ne": x != y, "lt": x < y, "le": x <= y, "gt": x > y, "ge": x >= y}[base]


def _s32(v: int) -> int:
    v &= M32
    return v - (1 << 32) if v & 0x80000000 else v


def _s64(v: int) -> int:
    v &= M64
    return v - (1 << 64) if v & (1 << 63) else v


def _round_int(fr: Fraction, mode: str) -> int:
    fl = fr.numerator // fr.denominator
    if mode == "rzi":
        return fl if fr >= 0 else -((-fr).numerator // (-fr).denominator)
    if mode == "rmi":
        return fl
    if mode == "rpi":
        return fl if fr == fl else fl + 1
    if mode == "rni":
        rem = fr - fl
        if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (fl & 1)):
            return fl + 1
        return fl
    raise NotImplementedError(mode)


# ---------------------------------------------------------------------------
# the interpreter
# ---------------------------------------------------------------------------
class StrictMemory(dict):
    """address -> word memory for Module.launch / Module.run in which a load from
    an address nobody has written raises (a plain dict reads such words as
    zero): out-of-bounds and uninitialised reads of the kernels -- global and
    shared memory alike -- surface the way compute-sanitizer's initcheck and
    memcheck show them on a device.  Read results back with `peek`."""

    def get(self, key, default=None):
        if key not in self:
            raise KeyError("load from unwritten address %#x" % key)
        return self[key]

    def peek(self, key, default=0):
        return dict.get(self, key, default)


class Function:
    def __init__(self, name, params, body):
        self.name, self.params, self.body = name, params, body
        self.labels = {}
        self.insns = []
        pending = ""
        for line in body:
            m = re.match(r"^(\$?[\w$]+):$", line)
            if m and not pending:
                self.labels[m.group(1)] = len(self.insns)
                continue
            # a statement runs to its ';' (call statements span several lines)
            pending = (pending + " " + line).strip()
            if pending.endswith(";"):
                self.insns.append(pending)
                pending = ""


def _body_lines(text: str):
    out = []
    for raw in text.split("\n"):
        line = raw.strip()
        if not line or line in ("{", "}") or line.startswith((".reg", ".pragma", ".loc", ".local", ".param", ".shared")):
            continue
        out.append(line)
    return out


class Module:
    """entries and initialised global arrays of one PTX text"""

    def __init__(self, ptx: str):
        self.functions = {}
        self.globals = {}                          # name -> (address, bytes)
        self._next_global = 0x10000000
        text = re.sub(r"//[^\n]*", "", ptx)
        for m in re.finditer(r"\.global\s+\.align\s+\d+\s+\.b8\s+(\w+)\[(\d+)\]\s*=\s*\{([^}]*)\};", text):
            data = bytes(int(v) for v in m.group(3).split(","))
            self.globals[m.group(1)] = (self._next_global, data)
            self._next_global += (len(data) + 255) & ~255
        for m in re.finditer(r"\.global\s+\.align\s+\d+\s+\.(?:u32|b32|s32|f32)\s+(\w+)\s*=\s*(-?\w+);", text):
            v = m.group(2)
            word = int(v[2:], 16) if v[:2] in ("0f", "0F") else int(v, 0)
            self.globals[m.group(1)] = (self._next_global, (word & M32).to_bytes(4, "little"))
            self._next_global += 256
        # uninitialised __constant__ arrays (filled by the caller of launch()) and __shared__ variables
        self.consts, self.shared = {}, {}
        self.struct_params = set()
        for m in re.finditer(r"\.const\s+\.align\s+\d+\s+\.b8\s+(\w+)\[(\d+)\];", text):
            self.consts[m.group(1)] = (self._next_global, int(m.group(2)))
            self._next_global += (int(m.group(2)) + 255) & ~255
        off = 0
        for m in re.finditer(r"\.shared\s+\.align\s+(\d+)\s+\.\w+\s+(\w+)(?:\[(\d+)\])?;", text):
            al = int(m.group(1))
            off = (off + al - 1)//al*al
            self.shared[m.group(2)] = 0x30000000 + off
            off += int(m.group(3) or 8)
        for m in re.finditer(r"\.visible\s+\.entry\s+(\w+)\s*\(([^)]*)\)\s*(?:\.\w+[^\n{]*\n\s*)*\{(.*?)\n\}", text, re.S):
            params = re.findall(r"(\w+_param_\d+)", m.group(2))
            self.functions[m.group(1)] = Function(m.group(1), params, _body_lines(m.group(3)))
        # device functions that NVRTC did not inline (libdevice slow paths): name(params) { body }
        for m in re.finditer(r"\.func\s*(?:\([^)]*\))?\s*(\w+)\s*\(([^)]*)\)\s*\{(.*?)\n\}", text, re.S):
            params = re.findall(r"(\w+_param_\d+)", m.group(2))
            self.functions[m.group(1)] = Function(m.group(1), params, _body_lines(m.group(3)))

    def run(self, name: str, inputs, n_out: int, max_steps: int = 20000):
        """run entry `name(float* out, const float* in0, const float* in1, ...)`;
        inputs: one list of uint32 bit patterns per input pointer; returns n_out words"""
        fn = self.functions[name]
        mem = {}
        for addr, data in self.globals.values():
            for i in range(0, len(data), 4):
                mem[addr + i] = int.from_bytes(data[i:i + 4], "little")
        out_addr = 0x1000
        ptrs = [out_addr]
        for k, words in enumerate(inputs):
            base = 0x2000 + 0x1000 * k
            ptrs.append(base)
            for i, w in enumerate(words):
                mem[base + 4 * i] = w & M32
        assert len(ptrs) == len(fn.params), (name, len(ptrs), fn.params)
        env = dict(zip(fn.params, ptrs))
        _Machine(self, fn, env, mem).execute(max_steps)
        return [mem.get(out_addr + 4 * i) for i in range(n_out)]


    def launch(self, name: str, grid, block, params, mem, consts=None, max_steps: int = 5_000_000, racecheck: bool = True):
        """Run kernel `name` over a grid (x, y, z) of blocks of `block` threads.
        params: one entry per kernel parameter -- an int for a scalar, bytes for a
        struct passed by value.  mem: the address -> 32-bit word dictionary shared
        by all threads (global memory; the caller places its buffers in it).
        consts: __constant__ arrays by name (lists of 32-bit words).  Threads of a
        block run as coroutines that meet at bar.sync and, per warp, at shfl.sync:
        enough for kernels whose warps are converged at those points.
        racecheck: shared-memory hazards between two threads of a block without a
        bar.sync in between (write after write, read after write, write after
        read) raise -- threads run one after the other here, so a kernel with such
        a hazard would give an order-dependent answer that proves nothing."""
        fn = self.functions[name]
        for addr, data in self.globals.values():
            for i in range(0, len(data), 4):
                mem.setdefault(addr + i, int.from_bytes(data[i:i + 4], "little"))
        for cname, words in (consts or {}).items():
            base, size = self.consts[cname]
            assert 4*len(words) <= size, cname
            for i, w in enumerate(words):
                mem[base + 4*i] = int(w) & M32
        env = {}
        for i, (pname, value) in enumerate(zip(fn.params, params)):
            if isinstance(value, (bytes, bytearray)):          # struct by value: lives in param space, addressable
                env[pname] = 0x40000000 + 0x1000*i
                self.struct_params.add(pname)
                for k in range(0, len(value), 4):
                    mem[env[pname] + k] = int.from_bytes(value[k:k + 4], "little")
            else:
                env[pname] = value
        gx, gy, gz = (tuple(grid) + (1, 1, 1))[:3]
        for bz in range(gz):
            for by in range(gy):
                for bx in range(gx):
                    for a in [k for k in mem if 0x30000000 <= k < 0x30100000]:
                        del mem[a]                 # shared memory does not outlive the block
                    self._run_block(fn, env, mem, (bx, by, bz), (gx, gy, gz), block, max_steps, racecheck)

    def _run_block(self, fn, env, mem, bid, gdim, nthreads, max_steps, racecheck=True):
        threads = []
        hazards = {} if racecheck else None        # shared address -> [last writer, readers] since the last barrier
        for t in range(nthreads):
            mach = _Machine(self, fn, env, mem)
            mach.tid, mach.hazards = t, hazards
            mach.special = {"%tid.x": t, "%tid.y": 0, "%tid.z": 0, "%ntid.x": nthreads, "%ntid.y": 1, "%ntid.z": 1,
                            "%ctaid.x": bid[0], "%ctaid.y": bid[1], "%ctaid.z": bid[2],
                            "%nctaid.x": gdim[0], "%nctaid.y": gdim[1], "%nctaid.z": gdim[2],
                            "%laneid": t % 32, "%warpid": t//32}
            mach.local_base = 0x20000000 + t*0x10000
            threads.append(mach.steps(max_steps))
        waiting = {}                               # thread -> event it stopped at
        alive = set(range(nthreads))
        reply = {t: None for t in alive}
        while alive:
            for t in sorted(alive - set(waiting)):
                try:
                    waiting[t] = threads[t].send(reply[t])
                    reply[t] = None
                except StopIteration:
                    alive.discard(t)
            if not alive:
                break
            progressed = False
            # shuffles: all live threads of a warp stand at one
            for w in range((nthreads + 31)//32):
                lanes = [t for t in range(32*w, min(32*w + 32, nthreads)) if t in alive]
                if lanes and all(t in waiting and waiting[t][0] == "shfl" for t in lanes):
                    vals = {t % 32: waiting[t][2] for t in lanes}
                    for t in lanes:
                        _, mode, v, b, c = waiting[t]
                        lane = t % 32
                        seg, cval = (c >> 8) & 31, c & 31
                        top = (lane & seg) | (cval & ~seg & 31)
                        if mode == "down":
                            src = lane + b
                            ok = src <= top
                        elif mode == "up":
                            src = lane - b
                            ok = src >= (lane & seg)
                        elif mode == "bfly":
                            src = lane ^ b
                            ok = src <= top
                        else:                      # idx
                            src = (lane & seg) | (b & ~seg & 31)
                            ok = src <= top
                        ok = ok and src in vals     # a lane that has exited supplies nothing
                        reply[t] = (vals[src] if ok else v, ok)
                        del waiting[t]
                    progressed = True
            if alive and all(t in waiting and waiting[t][0] == "bar" for t in alive):
                for t in alive:
                    reply[t] = None
                waiting.clear()
                if hazards is not None:
                    hazards.clear()
                progressed = True
            if not progressed and all(t in waiting for t in alive):
                raise RuntimeError("deadlock in %s: threads wait at %s" % (fn.name, {e[0] for e in waiting.values()}))


_DECODED = {}          # instruction text -> (guard predicate, opcode, operand text, text without the guard)
_OPERANDS = {}         # instruction text without guard -> (name, modifiers, operands, ftz, rounding mode, type)


class _Machine:
    special = {}
    tid, hazards = 0, None

    def shared_access(self, addr: int, write: bool):
        """racecheck bookkeeping for one 32-bit word of shared memory"""
        if self.hazards is None or not 0x30000000 <= addr < 0x30100000:
            return
        writer, readers = self.hazards.setdefault(addr, [None, set()])
        if write:
            if writer not in (None, self.tid):
                raise RuntimeError("shared-memory hazard in %s: threads %d and %d write %#x without a barrier" % (self.fn.name, writer, self.tid, addr))
            if readers - {self.tid}:
                raise RuntimeError("shared-memory hazard in %s: thread %d writes %#x that thread %d read, without a barrier"
                                   % (self.fn.name, self.tid, addr, min(readers - {self.tid})))
            self.hazards[addr][0] = self.tid
        else:
            if writer not in (None, self.tid):
                raise RuntimeError("shared-memory hazard in %s: thread %d reads %#x that thread %d wrote, without a barrier"
                                   % (self.fn.name, self.tid, addr, writer))
            readers.add(self.tid)

    def __init__(self, module, fn, params, mem, depth: int = 0):
        self.module, self.fn, self.params, self.mem = module, fn, params, mem
        self.reg = {}
        self.depth = depth
        self.local_base = 0x20000000 + depth*0x100000
        self.pspace = {}                           # .param variables of call sequences: name -> {offset: value}
        self.retvals = {}                          # func_retval0 of this function: offset -> value

    # operand access -------------------------------------------------------
    def val(self, tok: str, width: int = 32) -> int:
        tok = tok.strip()
        if tok.startswith("%"):
            if tok in self.reg:
                return self.reg[tok]
            return self.special[tok]
        if tok.startswith("0f") or tok.startswith("0F"):
            return int(tok[2:], 16)
        if tok.startswith("0d") or tok.startswith("0D"):
            return int(tok[2:], 16)
        if tok in self.module.globals:
            return self.module.globals[tok][0]
        if tok in self.module.consts:
            return self.module.consts[tok][0]
        if tok in self.params and isinstance(self.params[tok], int):
            return self.params[tok]
        if tok in self.module.shared:
            return self.module.shared[tok]
        if tok.startswith("__local_depot"):
            return self.local_base
        if re.match(r"^-?(0x[0-9a-fA-F]+|\d+)$", tok):
            return int(tok, 0) & (M64 if width == 64 else M32)
        raise NotImplementedError("operand " + tok)

    def pred(self, tok: str) -> bool:
        tok = tok.strip()
        if tok.startswith("!"):
            return not self.reg[tok[1:]]
        return bool(self.reg[tok])

    def addr(self, tok: str) -> int:
        m = re.match(r"^\[([^\]+]+)(\+(-?\d+))?\]$", tok.strip())
        base = m.group(1)
        off = int(m.group(3)) if m.group(3) else 0
        if base in self.params:
            return ("param", base)
        return (self.val(base, 64) + off) & M64

    # execution ------------------------------------------------------------
    def execute(self, max_steps: int):
        """single thread: no collectives"""
        for event in self.steps(max_steps):
            raise NotImplementedError("%s outside launch()" % event[0])

    def steps(self, max_steps: int):
        """run until the end; yields at bar.sync and shfl.sync (see Module.launch)"""
        pc, steps = 0, 0
        insns = self.fn.insns
        while pc < len(insns):
            steps += 1
            if steps > max_steps:
                raise RuntimeError("step limit in " + self.fn.name)
            dec = _DECODED.get(insns[pc])
            if dec is None:
                line = insns[pc].rstrip(";").strip()
                guard = re.match(r"^@(!?%p\d+)\s+(.*)$", line)
                if guard:
                    line = guard.group(2)
                op, rest = (re.split(r"\s+", line, maxsplit=1) + [""])[:2]
                dec = _DECODED[insns[pc]] = (guard.group(1) if guard else None, op, rest.strip(), line)
            pc += 1
            gpred, op, rest, line = dec
            if gpred is not None and not self.pred(gpred):
                continue
            if op in ("ret", "exit"):
                return
            if op in ("bra", "bra.uni"):
                pc = self.fn.labels[rest]
                continue
            if op.startswith("bar.") or op.startswith("barrier."):
                yield ("bar",)
                continue
            if op.startswith("shfl.sync"):
                args = [a.strip() for a in rest.split(",")]
                dst, _, pdst = args[0].partition("|")
                value, ok = yield ("shfl", op.split(".")[2], self.val(args[1]), self.val(args[2]), self.val(args[3]))
                self.reg[dst] = value
                if pdst:
                    self.reg[pdst] = ok
                continue
            self.step(op, rest, line)

    def step(self, op: str, rest: str, line: str):
        dec = _OPERANDS.get(line)
        if dec is None:
            parts = op.split(".")
            mods = parts[1:]
            # operands: split on commas outside braces
            dec = _OPERANDS[line] = (parts[0], mods, [a.strip() for a in re.split(r",\s*(?![^{]*\})", rest)], "ftz" in mods,
                                     next((m for m in mods if m in ("rn", "rz", "rm", "rp")), "rn"), mods[-1] if mods else "")
        name, mods, args, flush, mode, typ = dec
        R = self.reg

        if name == "call":
            m = re.match(r"^(?:\((\w+)\)\s*,)?\s*(\w+)\s*,\s*\(([^)]*)\)$", rest)
            if not m or m.group(2) not in self.module.functions:
                raise NotImplementedError(line)
            callee = self.module.functions[m.group(2)]
            actual = [a.strip() for a in m.group(3).split(",") if a.strip()]
            sub = _Machine(self.module, callee, {p: self.pspace[a] for p, a in zip(callee.params, actual)}, self.mem,
                           self.depth + 1)
            sub.execute(200000)
            if m.group(1):
                self.pspace[m.group(1)] = sub.retvals
            return
        if name in ("ld", "st") and mods[0] == "param":
            tok = args[1] if name == "ld" else args[0]
            m = re.match(r"^\[(%?\w+)(?:\+(\d+))?\]$", tok)
            var, off = m.group(1), int(m.group(2) or 0)
            wide = typ in ("u64", "b64", "s64", "f64")
            if name == "ld" and (var.startswith("%") or var in self.module.struct_params):
                base = (self.val(var, 64) if var.startswith("%") else self.params[var]) + off
                size = 8 if wide else 4
                for i, d in enumerate(t.strip() for t in args[0].strip("{}").split(",")):
                    a = base + i*size
                    R[d] = self.mem.get(a, 0) | ((self.mem.get(a + 4, 0) << 32) if wide else 0)
                return
            if name == "st":
                space = self.retvals if var.startswith("func_retval") else self.pspace.setdefault(var, {})
                space[off] = self.val(args[1], 64) & (M64 if wide else M32)
                return
            src = self.params[var] if var in self.params else self.pspace[var]
            R[args[0]] = src[off] if isinstance(src, dict) else src
            return
        if name == "st" and mods[0] == "shared":
            a0 = self.addr(args[0])
            n = len(args[1].split(",")) if args[1].startswith("{") else 1
            for k in range(n*(2 if typ in ("u64", "b64", "s64", "f64") else 1)):
                self.shared_access(a0 + 4*k, True)
        if name == "st" and args[1].startswith("{"):          # st.global.v2 / .v4 (32-bit elements)
            a = self.addr(args[0])
            for i, t in enumerate(x.strip() for x in args[1].strip("{}").split(",")):
                self.mem[a + 4*i] = self.val(t) & M32
            return
        if name == "ld":
            a = self.addr(args[1])
            if mods[0] == "shared":
                n = len(args[0].split(",")) if args[0].startswith("{") else 1
                for k in range(n*(2 if typ in ("u64", "b64", "s64", "f64") else 1)):
                    self.shared_access(a + 4*k, False)
            if typ in ("u8", "s8", "b8"):
                R[args[0]] = self.mem.get(a, 0) & 0xFF
                return
            wide = typ in ("u64", "b64", "s64", "f64")
            word = (lambda at: self.mem.get(at, 0) | (self.mem.get(at + 4, 0) << 32)) if wide else (lambda at: self.mem.get(at, 0))
            if args[0].startswith("{"):            # ld.global.v2 / .v4
                for i, d in enumerate(t.strip() for t in args[0].strip("{}").split(",")):
                    R[d] = word(a + i*(8 if wide else 4))
            else:
                R[args[0]] = word(a)
            return
        if name == "st":
            a = self.addr(args[0])
            v = self.val(args[1], 64)
            self.mem[a] = v & M32
            if typ in ("u64", "b64", "s64", "f64"):
                self.mem[a + 4] = (v >> 32) & M32
            return
        if name == "cvta":
            R[args[0]] = self.val(args[1], 64)
            return
        if name == "mov":
            d, s = args
            if d.startswith("{"):
                lo, hi = [t.strip() for t in d.strip("{}").split(",")]
                v = self.val(s, 64)
                R[lo], R[hi] = v & M32, (v >> 32) & M32
            elif s.startswith("{"):
                lo, hi = [t.strip() for t in s.strip("{}").split(",")]
                R[d] = (self.val(lo) & M32) | ((self.val(hi) & M32) << 32)
            else:
                R[d] = self.val(s, 64 if typ in ("b64", "u64", "s64", "f64") else 32)
            return
        if name in ("add", "sub", "mul", "fma") and typ == "f32":
            c = self.val(args[3]) if name == "fma" else None
            R[args[0]] = _arith(name, self.val(args[1]), self.val(args[2]), c, mode, flush)
            return
        if name in ("add", "sub", "mul", "fma") and typ == "f32x2":
            a, b = self.val(args[1], 64), self.val(args[2], 64)
            c = self.val(args[3], 64) if name == "fma" else None
            out = 0
            for sh in (0, 32):
                cc = ((c >> sh) & M32) if c is not None else None
                out |= _arith(name, (a >> sh) & M32, (b >> sh) & M32, cc, mode, flush) << sh
            R[args[0]] = out
            return
        if typ == "f64" and name in ("add", "sub", "mul", "div", "fma", "abs", "neg", "rcp", "sqrt", "min", "max"):
            R[args[0]] = self.f64(name, mods, [self.val(a, 64) for a in args[1:]])
            return
        if name == "setp" and typ == "f64":
            x, y = b2d(self.val(args[1], 64)), b2d(self.val(args[2], 64))
            nan = x != x or y != y
            cmp = mods[0]
            if cmp in ("num", "nan"):
                R[args[0]] = nan == (cmp == "nan")
            elif nan:
                R[args[0]] = cmp in ("equ", "neu", "ltu", "leu", "gtu", "geu")
            else:
                R[args[0]] = {"eq": x == y, "ne": x != y, "lt": x < y, "le": x <= y, "gt": x > y, "ge": x >= y}[cmp[:2]]
            return
        if name == "abs" and typ == "s32":
            R[args[0]] = abs(_s32(self.val(args[1]))) & M32
            return
        if name in ("abs", "neg") and typ == "f32":
            v = self.val(args[1])
            if flush:
                v = ftz(v)
            R[args[0]] = (v & 0x7FFFFFFF) if name == "abs" else (v ^ 0x80000000)
            return
        if name == "copysign":
            R[args[0]] = (self.val(args[1]) & 0x80000000) | (self.val(args[2]) & 0x7FFFFFFF)
            return
        if name in ("max", "min") and typ == "f32":
            a, b = self.val(args[1]), self.val(args[2])
            if flush:
                a, b = ftz(a), ftz(b)
            if is_nan(a) and is_nan(b):
                R[args[0]] = CANON_NAN
            elif is_nan(a):
                R[args[0]] = b
            elif is_nan(b):
                R[args[0]] = a
            else:
                x, y = to_frac(a), to_frac(b)
                if x == y:                        # +0 / -0: max prefers +0, min -0
                    R[args[0]] = (a & b) if name == "max" else (a | b)
                else:
                    R[args[0]] = (a if x > y else b) if name == "max" else (a if x < y else b)
            return
        if name in ("max", "min") and typ in ("u32", "s32", "u64", "s64"):
            w = int(typ[1:])
            cv = (_s64 if w == 64 else _s32) if typ[0] == "s" else (lambda v: v & ((1 << w) - 1))
            a, b = cv(self.val(args[1], w)), cv(self.val(args[2], w))
            R[args[0]] = (max(a, b) if name == "max" else min(a, b)) & ((1 << w) - 1)
            return
        if name in ("rcp", "rsqrt", "ex2", "lg2") and "approx" in mods:
            R[args[0]] = _approx(name, self.val(args[1]), flush)
            return
        if name == "sqrt" and "rn" in mods:
            R[args[0]] = _ieee("sqrt", self.val(args[1]), None, flush)
            return
        if name == "rcp" and "rn" in mods:
            R[args[0]] = _ieee("rcp", self.val(args[1]), None, flush)
            return
        if name == "div" and "rn" in mods and typ == "f32":
            R[args[0]] = _ieee("div", self.val(args[1]), self.val(args[2]), flush)
            return
        if name == "setp":
            cmp = mods[0]
            a_tok, b_tok = args[1], args[2]
            if typ == "f32":
                R[args[0]] = _cmp_f(cmp, self.val(a_tok), self.val(b_tok), flush)
            else:
                w = 64 if typ.endswith("64") else 32
                a, b = self.val(a_tok, w), self.val(b_tok, w)
                if typ.endswith("16"):
                    a, b = a & 0xFFFF, b & 0xFFFF
                    if typ.startswith("s"):
                        a, b = a - ((a >> 15) << 16), b - ((b >> 15) << 16)
                elif typ.startswith("s"):
                    a, b = (_s64(a), _s64(b)) if w == 64 else (_s32(a), _s32(b))
                R[args[0]] = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b,
                              "lo": a < b, "ls": a <= b, "hi": a > b, "hs": a >= b}[cmp]
            return
        if name == "selp":
            R[args[0]] = self.val(args[1], 64) if self.pred(args[3]) else self.val(args[2], 64)
            if typ in ("f32", "b32", "u32", "s32"):
                R[args[0]] &= M32
            return
        if name in ("and", "or", "xor") and typ == "pred":
            a, b = self.pred(args[1]), self.pred(args[2])
            R[args[0]] = {"and": a and b, "or": a or b, "xor": a != b}[name]
            return
        if name == "not" and typ == "pred":
            R[args[0]] = not self.pred(args[1])
            return
        if name in ("and", "or", "xor"):
            w = 64 if typ == "b64" else 32
            a, b = self.val(args[1], w), self.val(args[2], w)
            R[args[0]] = {"and": a & b, "or": a | b, "xor": a ^ b}[name] & (M64 if w == 64 else M32)
            return
        if name == "not":
            w = 64 if typ == "b64" else 32
            R[args[0]] = ~self.val(args[1], w) & (M64 if w == 64 else M32)
            return
        if name == "shl":
            w = 64 if typ == "b64" else 32
            n = self.val(args[2]) & M32
            R[args[0]] = (self.val(args[1], w) << n) & (M64 if w == 64 else M32) if n < w else 0
            return
        if name == "shr" and typ == "u16":
            R[args[0]] = (self.val(args[1]) & 0xFFFF) >> min(self.val(args[2]) & M32, 16)
            return
        if name == "shr":
            w = 64 if typ.endswith("64") else 32
            n = min(self.val(args[2]) & M32, w)
            v = self.val(args[1], w)
            if typ.startswith("s"):
                v = _s64(v) if w == 64 else _s32(v)
                n = min(n, w - 1)
            R[args[0]] = (v >> n) & (M64 if w == 64 else M32)
            return
        if name == "shf":                          # shf.l.wrap.b32 d, lo, hi, n
            lo, hi, n = self.val(args[1]), self.val(args[2]), self.val(args[3]) & 31
            v = ((hi << 32) | lo) << n if mods[0] == "l" else ((hi << 32) | lo) >> n
            R[args[0]] = ((v >> 32) & M32) if mods[0] == "l" else (v & M32)
            return
        if name in ("add", "sub") and typ in ("s16", "u16", "s32", "u32", "s64", "u64"):
            w = int(typ[1:])
            a, b = self.val(args[1], w), self.val(args[2], w)
            R[args[0]] = (a + b if name == "add" else a - b) & ((1 << w) - 1)
            return
        if name == "atom" and mods[-2] == "add":               # atom.global.add.u32 d, [a], b
            a = self.addr(args[1])
            R[args[0]] = self.mem.get(a, 0)
            self.mem[a] = (R[args[0]] + self.val(args[2])) & M32
            return
        if name in ("membar", "fence"):
            return
        if name == "neg" and typ in ("s32", "s64"):
            R[args[0]] = (-self.val(args[1], 64)) & (M64 if typ == "s64" else M32)
            return
        if name == "mul" and "wide" in mods:
            a, b = self.val(args[1]), self.val(args[2])
            if typ == "s32":
                a, b = _s32(a), _s32(b)
            R[args[0]] = (a * b) & M64
            return
        if name == "mad" and "wide" in mods:
            a, b = self.val(args[1]), self.val(args[2])
            if typ == "s32":
                a, b = _s32(a), _s32(b)
            R[args[0]] = (a * b + self.val(args[3], 64)) & M64
            return
        if name in ("mul", "mad") and ("lo" in mods or "hi" in mods):
            w = int(typ[1:])
            mask = (1 << w) - 1
            sg = (lambda v: (v & mask) - ((v & mask) >> (w - 1) << w)) if typ.startswith("s") else (lambda v: v & mask)
            prod = sg(self.val(args[1], 64)) * sg(self.val(args[2], 64))
            r = prod >> w if "hi" in mods else prod
            if name == "mad":
                r += sg(self.val(args[3], 64))
            R[args[0]] = r & mask
            return
        if name in ("div", "rem") and typ in ("s32", "u32", "s64", "u64"):
            w = 64 if typ.endswith("64") else 32
            sg = (_s64 if w == 64 else _s32) if typ.startswith("s") else (lambda v: v & (M64 if w == 64 else M32))
            a, b = sg(self.val(args[1], w)), sg(self.val(args[2], w))
            q = abs(a)//abs(b) if b else 0
            q = q if (a < 0) == (b < 0) else -q              # truncation, as C
            R[args[0]] = (q if name == "div" else a - q*b) & (M64 if w == 64 else M32)
            return
        if name == "cvt":
            R[args[0]] = self.cvt(mods, self.val(args[1], 64))
            return
        raise NotImplementedError(line)

    @staticmethod
    def f64(name, mods, a):
        """binary64 arithmetic (rn) on bit patterns"""
        if name == "abs":
            return a[0] & ~(1 << 63)
        if name == "neg":
            return a[0] ^ (1 << 63)
        if name == "fma":
            return fma64(a[0], a[1], a[2])
        x = b2d(a[0])
        y = b2d(a[1]) if len(a) > 1 else None
        if name == "rcp":
            if "approx" in mods:
                # fast gross approximation: the low 32 bits of the argument are ignored, those of the result zero
                x = b2d(a[0] & ~M32)
                if x == 0 or x != x or math.isinf(x):
                    return d2b(math.copysign(math.inf, x)) if x == 0 else (d2b(math.copysign(0.0, x)) if math.isinf(x) else a[0])
                return d2b(1.0/x) & ~M32
            x, y = 1.0, x
            name = "div"
        try:
            if name == "add":
                return d2b(x + y)
            if name == "sub":
                return d2b(x - y)
            if name == "mul":
                return d2b(x * y)
            if name == "sqrt":
                return d2b(math.sqrt(x)) if x >= 0 else d2b(math.nan)
            if name in ("min", "max"):
                if x != x or y != y:
                    return a[1] if x != x else a[0]
                return d2b(min(x, y) if name == "min" else max(x, y))
            if name == "div":
                if y == 0.0:
                    if x == 0.0 or x != x:
                        return d2b(math.nan)
                    return d2b(math.copysign(math.inf, x)*math.copysign(1.0, y))
                return d2b(x / y)
        except OverflowError:
            return d2b(math.inf if (x > 0) == (y is None or y > 0) else -math.inf)
        raise NotImplementedError(name)

    def cvt(self, mods, v: int) -> int:
        dst, src = mods[-2], mods[-1]
        if src == "f64" and dst in ("f64", "s32", "s64", "u32", "u64"):
            x = b2d(v)
            im = next((m for m in mods if m in ("rni", "rzi", "rmi", "rpi")), None)
            if dst == "f64":
                if x != x or math.isinf(x) or im is None:
                    return v
                n = _round_int(Fraction(x), im)
                return d2b(math.copysign(float(n), x))
            if x != x:
                return 0
            bits = 64 if dst.endswith("64") else 32
            lo, hi = ((-(1 << (bits - 1)), (1 << (bits - 1)) - 1) if dst[0] == "s" else (0, (1 << bits) - 1))
            n = (lo if x < 0 else hi) if math.isinf(x) else max(lo, min(hi, _round_int(Fraction(x), im or "rzi")))
            return n & ((1 << bits) - 1)
        if dst == "f64" and src in ("s32", "u32", "u64"):
            return d2b(float(_s32(v) if src == "s32" else (v & M32 if src == "u32" else v & M64)))
        if src == "u64" and dst == "u32":
            return v & M32
        flush = "ftz" in mods
        sat = "sat" in mods
        imode = next((m for m in mods if m in ("rni", "rzi", "rmi", "rpi")), None)
        if src == "f32" and dst == "f32":
            v &= M32
            if flush:
                v = ftz(v)
            if is_nan(v):
                return 0 if sat else CANON_NAN
            if imode and not is_inf(v):
                fr = to_frac(v)
                n = _round_int(fr, imode)
                v = round_frac(Fraction(n), "rn", False, bool(v & 0x80000000))
            if sat:
                fr = Fraction(0) if is_zero(v) else (Fraction(2) if is_inf(v) and not v & 0x80000000 else (Fraction(-1) if is_inf(v) else to_frac(v)))
                if fr <= 0:
                    return 0
                if fr >= 1:
                    return 0x3F800000
            return v
        if src == "f32" and dst in ("s32", "u32"):
            v &= M32
            if flush:
                v = ftz(v)
            if is_nan(v):
                return 0
            lo, hi = ((-(1 << 31), (1 << 31) - 1) if dst == "s32" else (0, M32))
            if is_inf(v):
                return (lo if v & 0x80000000 else hi) & M32
            n = _round_int(to_frac(v), imode)
            return max(lo, min(hi, n)) & M32
        if src in ("s32", "u32", "s64", "u64") and dst == "f32":
            n = _s32(v) if src == "s32" else v & M32 if src == "u32" else _s64(v) if src == "s64" else v & M64
            mode = next((m for m in mods if m in ("rn", "rz", "rm", "rp")), "rn")
            return round_frac(Fraction(n), mode)
        if src == "u32" and dst in ("u64", "s64"):
            return v & M32
        if src == "s32" and dst in ("s64", "u64"):
            return _s32(v) & M64
        if src in ("u64", "s64") and dst in ("u32", "s32"):
            return v & M32
        if src in ("u16", "u8") and dst in ("u32", "s32", "u64"):
            return v & (0xFFFF if src == "u16" else 0xFF)
        if dst in ("u16", "s16") and src in ("u32", "s32", "u64", "s64"):
            return v & 0xFFFF
        if src == "s32" and dst == "s64":
            return _s32(v) & M64
        if src == "s64" and dst == "f64":
            return d2b(float(_s64(v)))            # |v| < 2^63: Python rounds to nearest even
        if src == "f64" and dst == "f32":
            x = b2d(v)
            if x != x:
                return CANON_NAN
            if math.isinf(x):
                return 0x7F800000 | (0x80000000 if x < 0 else 0)
            return round_frac(Fraction(x), "rn", flush, math.copysign(1.0, x) < 0)
        if src == "f32" and dst == "f64":
            v &= M32
            if flush:
                v = ftz(v)
            return d2b(b2f(v))
        raise NotImplementedError("cvt." + ".".join(mods))
