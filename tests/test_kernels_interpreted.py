"""The CUDA kernels themselves, run on the CPU.

`tools/ptx_emu.py` executes NVRTC's PTX of a model's program as a grid of
thread blocks (threads as coroutines that meet at `bar.sync` and, per warp, at
`shfl.sync`; global, constant, shared and parameter memory; exact IEEE
arithmetic; loads from addresses nobody wrote raise, so out-of-bounds and
uninitialised reads show as they would under compute-sanitizer).  That is enough to run the whole evaluation of a small image --
`lcu_set_params`' body, `lcu_render_pair[_err]`, `lcu_render_s1`, `lcu_render_s8`
with the fused final reduction, `lcu_convolve` / `lcu_convolve_small` with the
fused chi^2, `lcu_reduce` -- and hold the result
against the oracle exactly as the GPU parity tests do: model image per pixel,
log-likelihood, pair kernel = one-ray kernel bit for bit, both convolution
kernels bit for bit.  What it does not exercise is the hardware (the special-
function unit is a correctly rounded stand-in) and ptxas; for those see
tests/test_gpu_parity.py and test_pair_rays.py.  Sizes are tiny: the
interpreter runs ~10^5 instructions per second.
"""
import dataclasses
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import ptx_emu as E  # noqa: E402
import helpers as H  # noqa: E402
import test_pair_rays as R  # noqa: E402

pytest.importorskip("cuda.bindings.nvrtc")

IMG, WGT, RAW, MODEL, PART, LNEW, RAW1, PART1, MODEL1, ERR, OBJS, COUNTER, LNEW8, RAW8, PART8 = (0x100000*k for k in range(1, 16))
RAWQ, PARTQ, LNEWQ = 0x100000*30, 0x100000*31, 0x100000*32
OUT_VALUE, OUT_ERROR, OUT_CHI2 = 1, 2, 4


def _put(mem, addr, arr):
    for i, v in enumerate(np.ascontiguousarray(arr).view(np.uint32).ravel()):
        mem[addr + 4*i] = int(v)


def _get(mem, addr, shape, dtype=np.float32):
    n = int(np.prod(shape))*np.dtype(dtype).itemsize//4
    return np.array([dict.get(mem, addr + 4*i, 0) for i in range(n)], np.uint32).view(dtype).reshape(shape)


def _render_args(cfg, pcs, nk, value, partial, ngroups, mode, error=0, objs=0, tail=(0, 0, 0.0), k0=0):
    # lcu_render_args of kernel/lensed.cu: pcs, k0, nk, objs, value, error, image, weight, chimap, partial, ngroups, mode, tail
    # lcu_render_args (kernel/lensed.cu): pcs, k0, nk, objs, params (fold kernels only), value, error, image, weight, chimap, partial, ...
    a = struct.pack("<4fqqQQQQQQQQiiQQd", *pcs, k0, nk, objs, 0, value, error, IMG, WGT, 0, partial, ngroups, mode, *tail)
    return a + b"\0"*(128 - len(a))


def _convolve_args(raw, model, partial, rows, ngroups, gpr, mode):
    # lcu_convolve_args: raw, model, image, weight, chimap, partial, row0, row1, ngroups, gpr, mode, tail
    a = struct.pack("<QQQQQQiiiii", raw, model, IMG, WGT, 0, partial, 0, rows, ngroups, gpr, mode)
    a += b"\0"*(72 - len(a)) + struct.pack("<QQd", 0, 0, 0.0)
    return a


def _scene(psf):
    """a 20 x 12 scene with the lens and the source inside the frame"""
    import lensed_b200 as L
    base = H.golden_config("sie")                                    # sie + sersic
    h, w = 12, 20
    params = base.params.copy()
    params[:5] = [10.3, 6.2, 3.0, 0.8, 30.0]                         # lens x y r q pa
    params[5:12] = [11.0, 6.6, 1.5, -3.0, 1.5, 0.7, 100.0]           # source x y r mag n q pa
    cfg = dataclasses.replace(base, name="tiny-sie" + ("-psf" if psf is not None else ""), params=params,
                              image=np.zeros((h, w), np.float32), weight=np.ones((h, w), np.float32), rule="sub2", psf=psf)
    _, model, _ = cfg.oracle().loglike(params, want_maps=True)
    cfg.image, cfg.weight = H.workloads.observe(model, 5)
    cfg.weight[3, 4] = 0                                             # a masked pixel
    return cfg, L


def _program(cfg, L, extra=()):
    ctx = L.Context(device=-1, objects_dir=H.OBJECTS_DIR)
    try:
        m = cfg.product(ctx, flags=L.LCU_SOURCE_ONLY)
        text, words = m.source, m.words
    finally:
        ctx.close()
    return E.Module(R._compile_ptx(text, 0, extra)), text, words


def _object_block(M, text, cfg, words):
    """lcu_set_params' body, interpreted (tests/test_pair_rays.py::test_set_params_against_oracle checks it)"""
    head = text.partition("// kernel/lensed.cu\n")[0]
    S = E.Module(R._compile_ptx(head + R.SETTER, 0))
    out = S.run("k_set", [[E.f2b(float(v)) for v in cfg.params]], words, max_steps=2000000)
    return [v or 0 for v in out]


def test_render_and_reduce_without_psf():
    from lensed_b200 import api
    cfg, L = _scene(None)
    M, text, words = _program(cfg, L)
    block = _object_block(M, text, cfg, words)
    h, w = cfg.image.shape
    npix, ngroups = h*w, (h*w + 31)//32
    qq, ww = api.quad_rule(cfg.rule, cfg.pcs[2], cfg.pcs[3])
    consts = {"lcu_quad": np.c_[qq, ww].astype(np.float32).view(np.uint32).ravel(), "lcu_objs_c": block}
    mem = E.StrictMemory()        # a load from an address nobody wrote raises
    _put(mem, IMG, cfg.image)
    _put(mem, WGT, cfg.weight)
    M.launch("lcu_render_pair", ((npix + 511)//512, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW, PART, ngroups, OUT_VALUE | OUT_CHI2)], mem, consts)
    M.launch("lcu_reduce", (1,), 256, [ngroups, PART, E.d2b(-0.5), LNEW], mem)
    M.launch("lcu_render_s1", ((npix + 255)//256, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW1, PART1, ngroups, OUT_VALUE | OUT_CHI2)], mem, consts)
    om = cfg.oracle()
    ref_l, ref_model, _ = om.loglike(cfg.params, want_maps=True)
    got = _get(mem, RAW, (h, w))
    rel = H.rel_err(got, ref_model)
    assert rel.max() <= 1e-5, rel.max()                              # BASELINE.json: per-pixel relative error
    lnew = float(_get(mem, LNEW, (1,), np.float64)[0])
    assert abs(lnew - ref_l) <= 1e-6*abs(ref_l), (lnew, ref_l)       # BASELINE.json: log-likelihood
    # two rays per thread: the one-ray kernel's bits, values and chi^2 partial sums
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAW1, (npix,)).view(np.uint32))
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PART1, (2*ngroups,)).view(np.uint32))
    # the small-image path (single-point latency): eight warps share a 32-pixel group's quadrature points, object
    # blocks from global memory, and the block that finishes last adds the partial sums up itself
    # (lcu_fused_reduce: a counter, a fence) -- same image, same sums, same log-likelihood bits
    _put(mem, OBJS, np.array(block, np.uint32))
    mem[COUNTER] = 0                                                 # the runtime allocates the counters zeroed
    M.launch("lcu_render_s8", ((npix + 31)//32, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW8, PART8, ngroups, OUT_VALUE | OUT_CHI2, objs=OBJS, tail=(LNEW8, COUNTER, -0.5))],
             mem, consts)
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAW8, (npix,)).view(np.uint32))
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PART8, (2*ngroups,)).view(np.uint32))
    assert np.array_equal(_get(mem, LNEW, (2,)).view(np.uint32), _get(mem, LNEW8, (2,)).view(np.uint32))
    assert mem[COUNTER] == 0                                         # left at zero for the next launch
    # four warps per group, no fused tail: partial sums for lcu_reduce
    M.launch("lcu_render_s4", ((npix + 63)//64, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW8, PART8, ngroups, OUT_VALUE | OUT_CHI2, objs=OBJS)], mem, consts)
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAW8, (npix,)).view(np.uint32))
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PART8, (2*ngroups,)).view(np.uint32))
    # the split kernels of pairable models: two quadrature points per thread and pass through lcu_compute2 -- the same
    # image, partial sums and fused log-likelihood, bit for bit (eight warps with the fused tail, then two warps)
    mem[COUNTER] = 0
    M.launch("lcu_render_q_s8", ((npix + 31)//32, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAWQ, PARTQ, ngroups, OUT_VALUE | OUT_CHI2, objs=OBJS, tail=(LNEWQ, COUNTER, -0.5))],
             mem, consts)
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAWQ, (npix,)).view(np.uint32))
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PARTQ, (2*ngroups,)).view(np.uint32))
    assert np.array_equal(_get(mem, LNEW, (2,)).view(np.uint32), _get(mem, LNEWQ, (2,)).view(np.uint32))
    assert mem[COUNTER] == 0
    M.launch("lcu_render_q_s2", ((npix + 127)//128, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAWQ, PARTQ, ngroups, OUT_VALUE | OUT_CHI2, objs=OBJS)], mem, consts)
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAWQ, (npix,)).view(np.uint32))
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PARTQ, (2*ngroups,)).view(np.uint32))
    # the quadrature error image (the dumper's ERR layer) from the pair kernel
    M.launch("lcu_render_pair_err", ((npix + 511)//512, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW1, 0, ngroups, OUT_VALUE | OUT_ERROR, error=ERR)], mem, consts)
    ref_err = np.asarray(om.render(cfg.params)[1])
    e = np.abs(_get(mem, ERR, (h, w)).astype(np.float64) - ref_err)/np.maximum(np.abs(ref_model), 1e-30)
    assert e.max() <= 1e-4, e.max()                                  # alternating-sign weights: on the scale of the value
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAW1, (npix,)).view(np.uint32))


def test_render_convolve_reduce_with_psf():
    from lensed_b200 import api
    psf = H.workloads.gaussian_psf(5, 3, 1.0)                        # 5 wide, 3 high
    cfg, L = _scene(psf)
    M, text, words = _program(cfg, L)
    block = _object_block(M, text, cfg, words)
    h, w = cfg.image.shape
    npix, gpr = h*w, (w + 31)//32
    ngroups = h*gpr
    qq, ww = api.quad_rule(cfg.rule, cfg.pcs[2], cfg.pcs[3])
    consts = {"lcu_quad": np.c_[qq, ww].astype(np.float32).view(np.uint32).ravel(), "lcu_objs_c": block,
              "lcu_psf": psf.astype(np.float32).view(np.uint32).ravel()}
    mem = E.StrictMemory()        # a load from an address nobody wrote raises
    _put(mem, IMG, cfg.image)
    _put(mem, WGT, cfg.weight)
    M.launch("lcu_render_pair", ((npix + 511)//512, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW, 0, (npix + 31)//32, OUT_VALUE)], mem, consts)
    M.launch("lcu_convolve", ((w + 63)//64, (h + 31)//32, 1), 256,
             [_convolve_args(RAW, MODEL, PART, h, ngroups, gpr, OUT_VALUE | OUT_CHI2)], mem, consts)
    M.launch("lcu_reduce", (1,), 256, [ngroups, PART, E.d2b(-0.5), LNEW], mem)
    M.launch("lcu_convolve_small", ((w + 31)//32, (h + 7)//8, 1), 256,
             [_convolve_args(RAW, MODEL1, PART1, h, ngroups, gpr, OUT_VALUE | OUT_CHI2)], mem, consts)
    om = cfg.oracle()
    ref_l, ref_model, _ = om.loglike(cfg.params, want_maps=True)
    raw = _get(mem, RAW, (h, w))
    # the convolution is the reference's arithmetic in the reference's order: bit-exact given the same input
    assert np.array_equal(_get(mem, MODEL, (h, w)).view(np.uint32), np.asarray(om.convolve(raw), np.float32).view(np.uint32))
    assert H.rel_err(_get(mem, MODEL, (h, w)), ref_model).max() <= 1e-5
    lnew = float(_get(mem, LNEW, (1,), np.float64)[0])
    assert abs(lnew - ref_l) <= 1e-6*abs(ref_l), (lnew, ref_l)
    # the small-launch convolution kernel: same bits, image and partial sums
    assert np.array_equal(_get(mem, MODEL, (npix,)).view(np.uint32), _get(mem, MODEL1, (npix,)).view(np.uint32))
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PART1, (2*ngroups,)).view(np.uint32))


def test_power_law_lens_with_written_out_pair_math():
    """epl_plus_shear + sersic through lcu_render_pair built with
    -DLCU_PF_LIBM_PAIR=1 (atan2 / sincos / powr of pairs as packed arithmetic,
    the default since round 2): the one-ray kernel's bits, the oracle's image."""
    from lensed_b200 import api
    base = H.golden_config("epl_plus_shear")
    h, w = 8, 12
    params = base.params.copy()
    params[:8] = [6.3, 4.2, 2.0, 1.2, 0.7, 40.0, 0.03, -0.02]         # lens x y r t q pa g1 g2
    params[8:15] = [6.8, 4.5, 1.0, -3.0, 1.5, 0.7, 100.0]            # source x y r mag n q pa
    cfg = dataclasses.replace(base, name="tiny-epl", params=params, image=np.zeros((h, w), np.float32),
                              weight=np.ones((h, w), np.float32), rule="sub2", psf=None)
    import lensed_b200 as L
    M, text, words = _program(cfg, L, ["-DLCU_PF_LIBM_PAIR=1"])
    assert sum("fma.rn.ftz.f32x2" in line for line in M.functions["lcu_render_pair"].body) > 40   # the packed libm is in
    block = _object_block(M, text, cfg, words)
    npix, ngroups = h*w, (h*w + 31)//32
    qq, ww = api.quad_rule(cfg.rule, cfg.pcs[2], cfg.pcs[3])
    consts = {"lcu_quad": np.c_[qq, ww].astype(np.float32).view(np.uint32).ravel(), "lcu_objs_c": block}
    mem = E.StrictMemory()        # a load from an address nobody wrote raises
    _put(mem, IMG, cfg.image)
    _put(mem, WGT, cfg.weight)
    M.launch("lcu_render_pair", ((npix + 511)//512, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW, PART, ngroups, OUT_VALUE | OUT_CHI2)], mem, consts)
    M.launch("lcu_render_s1", ((npix + 255)//256, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW1, PART1, ngroups, OUT_VALUE | OUT_CHI2)], mem, consts)
    ref = np.asarray(cfg.oracle().render(cfg.params)[0])
    assert H.rel_err(_get(mem, RAW, (h, w)), ref).max() <= 1e-5
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAW1, (npix,)).view(np.uint32))
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PART1, (2*ngroups,)).view(np.uint32))


def test_batch_index_and_row_strip():
    """blockIdx.y is the parameter point (object blocks and outputs strided by
    point); a launch restricted to image rows [2, 7) -- how lcu_model_set_rows
    shards very large images over GPUs -- writes the same bits into those rows."""
    from lensed_b200 import api
    cfg, L = _scene(None)
    cfg = dataclasses.replace(cfg, rule="point")                      # one ray per pixel: this test is about indexing
    M, text, words = _program(cfg, L)
    second = dataclasses.replace(cfg, params=cfg.params*np.float32(1.01))
    # the object blocks of both points from the set_params kernel itself (one thread per point)
    PARAMS = 0x100000*20
    mem0 = E.StrictMemory()
    _put(mem0, PARAMS, np.concatenate([cfg.params, second.params]).astype(np.float32))
    M.launch("lcu_set_params", (1,), 64, [2, PARAMS, OBJS], mem0)
    blocks = [int(v) for v in _get(mem0, OBJS, (2*words,), np.uint32)]
    assert blocks[:words] == _object_block(M, text, cfg, words) and blocks[words:] == _object_block(M, text, second, words)
    assert blocks[:words] != blocks[words:]
    h, w = cfg.image.shape
    npix, ngroups = h*w, (h*w + 31)//32
    qq, ww = api.quad_rule(cfg.rule, cfg.pcs[2], cfg.pcs[3])
    consts = {"lcu_quad": np.c_[qq, ww].astype(np.float32).view(np.uint32).ravel(), "lcu_objs_c": blocks}
    mem = E.StrictMemory()        # a load from an address nobody wrote raises
    _put(mem, IMG, cfg.image)
    _put(mem, WGT, cfg.weight)
    M.launch("lcu_render_pair", (1, 2), 256, [_render_args(cfg, cfg.pcs, npix, RAW, PART, ngroups, OUT_VALUE | OUT_CHI2)], mem, consts)
    M.launch("lcu_reduce", (2,), 256, [ngroups, PART, E.d2b(-0.5), LNEW], mem)
    full = _get(mem, RAW, (2, h, w))
    lnew = _get(mem, LNEW, (2,), np.float64)
    for b, c in enumerate((cfg, second)):
        ref_l, ref_model, _ = c.oracle().loglike(c.params, want_maps=True)
        assert H.rel_err(full[b], ref_model).max() <= 1e-5
        # the observation was made with another rule: a badly fitting point (chi^2/dof ~ 10), where per-pixel
        # rounding adds up coherently with the residuals -- the GPU tests' bound for such points
        assert abs(lnew[b] - ref_l) <= 4e-6*abs(ref_l)
    assert not np.array_equal(full[0], full[1])
    r0, r1 = 2, 7
    M.launch("lcu_render_pair", (1, 2), 256,
             [_render_args(cfg, cfg.pcs, (r1 - r0)*w, RAW1, PART1, ((r1 - r0)*w + 31)//32, OUT_VALUE | OUT_CHI2, k0=r0*w)], mem, consts)
    strip = _get(mem, RAW1, (2, h, w))
    assert np.array_equal(strip[:, r0:r1].view(np.uint32), full[:, r0:r1].view(np.uint32))
    assert not strip[:, :r0].any() and not strip[:, r1:].any()       # nothing outside the strip is touched


def test_weight_map_on_the_device():
    """lcu_make_weight: gain / (image + offset) in double, narrowed once, masked
    pixels zero -- the bits of make_weight() + the mask loop (src/data.c:314-330,
    src/lensed.c:470-482), with a gain value and with a gain map"""
    cfg, L = _scene(None)
    M, _, _ = _program(cfg, L)
    h, w = cfg.image.shape
    n = h*w
    rng = np.random.default_rng(2)
    gain_map = rng.uniform(500, 3000, (h, w)).astype(np.float32)
    mask = (rng.random((h, w)) < 0.2).astype(np.int32)
    offset = 2.9633
    GAIN, MASK, OUT = 0x100000*21, 0x100000*22, 0x100000*23
    for use_map in (False, True):
        mem = E.StrictMemory()        # a load from an address nobody wrote raises
        _put(mem, IMG, cfg.image)
        _put(mem, GAIN, gain_map)
        _put(mem, MASK, mask)
        M.launch("lcu_make_weight", (2,), 256, [n, IMG, GAIN if use_map else 0, E.f2b(1800.0), E.d2b(offset), MASK, OUT], mem)
        gain = gain_map if use_map else np.float32(1800.0)
        ref = (np.asarray(gain, np.float32).astype(np.float64)/(cfg.image.astype(np.float64) + offset)).astype(np.float32)
        ref = np.where(mask != 0, np.float32(0), ref)
        assert np.array_equal(_get(mem, OUT, (h, w)).view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("seed", [1, 3])
def test_convolution_kernels_on_random_shapes(seed):
    """both convolution kernels on a random image / PSF shape (odd and even
    PSF sizes, widths that are not a multiple of the tile): the oracle's
    convolution bit for bit, equal chi^2 partial sums"""
    import lensed_b200 as L
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(1, 70)), int(rng.integers(1, 40))
    pw, ph = int(rng.integers(1, 8)), int(rng.integers(1, 8))
    psf = H.workloads.normalise_psf(rng.random((ph, pw)).astype(np.float32) + 0.1)
    cfg = dataclasses.replace(H.golden_config("sky"), name="conv-%d" % seed, image=rng.random((h, w)).astype(np.float32),
                              weight=(rng.random((h, w)) + 0.5).astype(np.float32), rule="point", psf=psf)
    M, _, _ = _program(cfg, L)
    raw = rng.random((h, w)).astype(np.float32)*10
    mem = E.StrictMemory()        # a load from an address nobody wrote raises
    _put(mem, IMG, cfg.image)
    _put(mem, WGT, cfg.weight)
    _put(mem, RAW, raw)
    gpr = (w + 31)//32
    ngroups = h*gpr
    consts = {"lcu_psf": psf.view(np.uint32).ravel()}
    M.launch("lcu_convolve", ((w + 63)//64, (h + 31)//32, 1), 256,
             [_convolve_args(RAW, MODEL, PART, h, ngroups, gpr, OUT_VALUE | OUT_CHI2)], mem, consts)
    M.launch("lcu_convolve_small", ((w + 31)//32, (h + 7)//8, 1), 256,
             [_convolve_args(RAW, MODEL1, PART1, h, ngroups, gpr, OUT_VALUE | OUT_CHI2)], mem, consts)
    ref = np.asarray(cfg.oracle().convolve(raw), np.float32).view(np.uint32)
    assert np.array_equal(_get(mem, MODEL, (h, w)).view(np.uint32), ref), (w, h, pw, ph)
    assert np.array_equal(_get(mem, MODEL1, (h, w)).view(np.uint32), ref), (w, h, pw, ph)
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PART1, (2*ngroups,)).view(np.uint32))


@pytest.mark.parametrize("seed,libm", [(4, False), (102, True)])
def test_random_models_through_the_render_kernels(seed, libm):
    """random object combinations (optional host, one or two lenses, one or two
    sources, sky) in a small frame: the pair kernel writes the one-ray kernel's
    bits and both reproduce the oracle within the GPU tests' bound; seed 102
    draws a power-law lens and runs with -DLCU_PF_LIBM_PAIR=1.  (A wider sweep
    of this test -- 24 seeds, 50 convolution shapes -- was run once by hand.)"""
    import lensed_b200 as L
    from lensed_b200 import api
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(6, 14)), int(rng.integers(8, 20))
    objects, params = [], []

    def add(name, role):
        objects.append(name)
        params.extend(H._random_params(rng, name, w, h, role))
    if rng.random() < 0.3:
        add(str(rng.choice(H.SOURCES)), "host")
    for _ in range(int(rng.integers(1, 3))):
        add(str(rng.choice(H.LENSES)), "lens")
    for _ in range(int(rng.integers(1, 3))):
        add(str(rng.choice(H.SOURCES)), "source")
    if rng.random() < 0.7:
        add("sky", "sky")
    cfg = H.Config(name="random-tiny-%d" % seed, objects=objects, params=np.array(params, np.float32),
                   image=np.zeros((h, w), np.float32), weight=np.ones((h, w), np.float32),
                   rule=str(rng.choice(["point", "sub2"])), psf=None)
    assert not libm or any(o.startswith("epl") for o in objects)
    M, text, words = _program(cfg, L, ["-DLCU_PF_LIBM_PAIR=1"] if libm else [])
    block = _object_block(M, text, cfg, words)
    npix, ngroups = h*w, (h*w + 31)//32
    qq, ww = api.quad_rule(cfg.rule, 1, 1)
    consts = {"lcu_quad": np.c_[qq, ww].astype(np.float32).view(np.uint32).ravel(), "lcu_objs_c": block}
    mem = E.StrictMemory()        # a load from an address nobody wrote raises
    _put(mem, IMG, cfg.image)
    _put(mem, WGT, cfg.weight)
    M.launch("lcu_render_pair", ((npix + 511)//512, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW, PART, ngroups, OUT_VALUE | OUT_CHI2)], mem, consts)
    M.launch("lcu_render_s1", ((npix + 255)//256, 1), 256,
             [_render_args(cfg, cfg.pcs, npix, RAW1, PART1, ngroups, OUT_VALUE | OUT_CHI2)], mem, consts)
    ref = np.asarray(cfg.oracle().render(cfg.params)[0])
    floor = H.rel_err(ref, np.asarray(cfg.oracle("f64").render(cfg.params)[0], np.float64)).max()
    assert H.rel_err(_get(mem, RAW, (h, w)), ref).max() <= max(1e-5, 1.5*floor), objects
    assert np.array_equal(_get(mem, RAW, (npix,)).view(np.uint32), _get(mem, RAW1, (npix,)).view(np.uint32)), objects
    assert np.array_equal(_get(mem, PART, (2*ngroups,)).view(np.uint32), _get(mem, PART1, (2*ngroups,)).view(np.uint32))
