// lensed.cu -- hand-written sm_100a kernels of the per-likelihood hot path.
//
// Replaces the reference's kernel/lensed.cl (render :9-38, loglike :41-53,
// convolve :56-103), the generated set_params kernel (src/kernel.c:114-150)
// and the host-side chi^2 sum of src/nested.c:102-115.  This file is appended
// by the host (lcu_program.cpp) after the object plugins and the generated
// lcu_compute() / lcu_set_params_body(), and compiled with NVRTC.  Macros
// provided by the host: IMAGE_SIZE IMAGE_WIDTH IMAGE_HEIGHT PSF PSF_WIDTH
// PSF_HEIGHT QUAD_POINTS (src/kernel.c:890-896 names), LCU_QUAD_NI x LCU_QUAD_NJ
// (grid shape if the rule is Cartesian, else 0), LCU_WORDS (object block
// size in 4-byte words), LCU_NPARS, LCU_MAXB (parameter points per launch),
// LCU_OBJ_CONST (object blocks in the constant bank instead of shared memory).
//
// Stages, all batched over parameter points (blockIdx.y = point):
//   lcu_set_params   1 thread / point : params -> object block
//   lcu_render_s{S}  pixel x sub-pixel quadrature ray shooting; S warps share
//                    one 32-pixel group's quadrature points (S = 1 for large
//                    images, 8 for small ones); the quadrature sum is always
//                    accumulated in ascending point order, so results do not
//                    depend on S or on the batch size.  Without a PSF the
//                    chi^2 terms are fused in.
//   lcu_convolve     PSF convolution (true, flipped, edge-clamped) from a
//                    shared-memory tile with register reuse along x, PSF in the
//                    constant bank, fused with the masked weighted chi^2.
//   lcu_reduce       deterministic double-precision sum of the per-32-pixel
//                    partials -> log-likelihood.
//
// Arithmetic that the reference writes as separate operations is kept as
// separate IEEE operations (__fmul_rn / __fadd_rn never contract), so the
// images match the CPU oracle to the last bit wherever libm agrees.

#define LCU_OUT_VALUE  1
#define LCU_OUT_ERROR  2
#define LCU_OUT_CHI2   4
#define LCU_OUT_CHIMAP 8

#define LCU_BLOCK 256

// (qx, qy, weight, error weight) per quadrature point, src/quadrature.c:32-43
__constant__ float4 lcu_quad[QUAD_POINTS];

#if PSF
__constant__ float lcu_psf[PSF_WIDTH*PSF_HEIGHT];
#endif

#if LCU_OBJ_CONST
// object blocks of the points of the current launch; filled by a
// device-to-device copy of lcu_set_params' output between the two kernels
__constant__ uint4 lcu_objs_c[LCU_MAXB*LCU_WORDS/4];
#endif

// Final reduction fused into the kernel that writes the chi^2 partials (the
// small-launch kernels only: one kernel node less on the single-point latency
// path).  out == null: the host launches lcu_reduce instead.
struct lcu_tail
{
    double* out;            // [B] scale * sum of the point's partials
    unsigned* counter;      // [B] blocks of the point that have finished; zero between launches
    double scale;
};

struct lcu_render_args
{
    float4 pcs;             // (rx, ry, sx, sy), src/lensed.c:879-891
    long long k0;           // first pixel of the rendered range
    long long nk;           // number of pixels in the range
    const uint* objs;       // [B][LCU_WORDS]
    const float* params;    // [B][LCU_NPARS]: read by the lcu_render_fold_s* kernels, which run set_params
                            // themselves (one thread per block) instead of reading objs
    float* value;           // [B][IMAGE_SIZE] or null
    float* error;           // [B][IMAGE_SIZE] or null
    const float* image;     // [IMAGE_SIZE]
    const float* weight;    // [IMAGE_SIZE]
    float* chimap;          // [B][IMAGE_SIZE] or null
    double* partial;        // [B][ngroups]
    int ngroups;
    int mode;
    lcu_tail tail;
};

// ---------------------------------------------------------------------------
// set_params: one thread per parameter point
// ---------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(64)
lcu_set_params(int B, const float* __restrict__ params, uint* __restrict__ objs)
{
    const int b = blockIdx.x*blockDim.x + threadIdx.x;
    if(b >= B)
        return;

    __align__(16) uint blk[LCU_WORDS];
#pragma unroll
    for(int i = 0; i < LCU_WORDS; ++i)
        blk[i] = 0;

    lcu_set_params_body(blk, params + (size_t)b*LCU_NPARS);

    uint4* out = reinterpret_cast<uint4*>(objs + (size_t)b*LCU_WORDS);
#pragma unroll
    for(int i = 0; i < LCU_WORDS/4; ++i)
        out[i] = make_uint4(blk[4*i], blk[4*i+1], blk[4*i+2], blk[4*i+3]);
}

// set_params inside a render block (small launches: the single-point latency
// path).  Every block of a point computes the point's object block itself, in one
// thread, with exactly the code of lcu_set_params -- the same bits, redundantly
// -- which takes the set_params kernel and the dependency on it out of the
// launch sequence.  Not inlined: the double-precision setter code must not cost
// the ray loop registers.
__device__ __noinline__ void lcu_set_params_block(uint* sdata, const float* __restrict__ params)
{
    __align__(16) uint blk[LCU_WORDS];
#pragma unroll
    for(int i = 0; i < LCU_WORDS; ++i)
        blk[i] = 0;
    lcu_set_params_body(blk, params);
#pragma unroll
    for(int i = 0; i < LCU_WORDS; ++i)
        sdata[i] = blk[i];
}

// ---------------------------------------------------------------------------
// final reduction: one block per point, fixed summation shape
// out = scale * sum_g partial[g]   (scale = -0.5 gives the log-likelihood
// of src/nested.c:115; scale = 1 the chi^2 of one row strip)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void lcu_reduce_block(int ngroups, const double* p, double scale, double* out)
{
    __shared__ double sm[LCU_BLOCK/32];

    double s = 0;
    for(int g = threadIdx.x; g < ngroups; g += LCU_BLOCK)
        s += __ldcg(p + g);
#pragma unroll
    for(int off = 16; off > 0; off >>= 1)
        s += __shfl_down_sync(0xffffffffu, s, off);
    if((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if(threadIdx.x == 0)
    {
        double t = 0;
#pragma unroll
        for(int w = 0; w < LCU_BLOCK/32; ++w)
            t += sm[w];
        *out = scale*t;
    }
}

// Called by every thread of every block of point b once the block's partials
// are written: the block that finishes last adds them up, in lcu_reduce's shape.
__device__ __forceinline__ bool lcu_fused_reduce(const lcu_tail& t, int b, unsigned nblocks, int ngroups, const double* partial)
{
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if(threadIdx.x == 0)
        last = atomicAdd(t.counter + b, 1u) == nblocks - 1;
    __syncthreads();
    if(!last)
        return false;
    __threadfence();
    if(threadIdx.x == 0)
        t.counter[b] = 0;
    lcu_reduce_block(ngroups, partial + (size_t)b*ngroups, t.scale, t.out + b);
    return true;
}

// ---------------------------------------------------------------------------
// render
// ---------------------------------------------------------------------------

// quadrature points staged per chunk when several warps share a pixel group
#define LCU_CHUNK 32

// ERR: also accumulate the quadrature error estimate (second weight); only
// the dumper asks for it, the likelihood path does not pay for it
// FOLD: the block computes its point's object block itself (lcu_set_params_block)
// PAIRQ (split kernels of pairable models): a thread shoots two quadrature points
// of its pixel per pass through lcu_compute2() -- the packed two-rays code of
// shim.cuh with the lanes on consecutive points instead of on two pixels -- which
// halves the chain of dependent ray evaluations of a small launch (7 -> 4 for rule
// g3k7 shared by 8 warps).  Each lane computes the one-ray bits, the values land
// in the same shared-memory slots and are added up in the same order.
// EXT (lcu_point_*: one kernel for all stages of a single point): the object
// block and the block index come from the caller, the point is number 0 and the
// caller runs the tail.
template<int S, bool ERR, bool FOLD = false, bool PAIRQ = false, bool EXT = false>
__device__ __forceinline__ void lcu_render_impl(const lcu_render_args& a, const uint* ext_data = nullptr, unsigned ext_bx = 0)
{
    constexpr int P = LCU_BLOCK/S;          // pixels per block
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int pg = warp/S;                  // 32-pixel group within the block
    const int ns = warp%S;                  // which share of the points
    const int b = EXT ? 0 : blockIdx.y;
    const unsigned bx = EXT ? ext_bx : blockIdx.x;

    const long long kk = (long long)bx*P + pg*32 + lane;
    const bool live = kk < a.nk;
    const long long k = a.k0 + (live ? kk : 0);

    // Object block of this point: constant bank for the large-image kernel
    // (S = 1; the host copies the blocks there between set_params and render),
    // shared memory, filled from set_params' output in global memory, for the
    // split kernels -- those run on small images, where the extra copy node
    // would cost more than the block's 48 loads (single-point latency)
    const uint* data;
    if constexpr(EXT)
        data = ext_data;
    else
#if LCU_OBJ_CONST
    if constexpr(S == 1)
        data = reinterpret_cast<const uint*>(lcu_objs_c) + b*LCU_WORDS;
    else
#endif
    {
        __shared__ __align__(16) uint sdata[LCU_WORDS];
        if constexpr(FOLD)
        {
            // one coalesced read of the point's parameters per block (they may live in
            // mapped host memory: one bus transaction, not LCU_NPARS of them)
            __shared__ float sparams[LCU_NPARS > 0 ? LCU_NPARS : 1];
            if(threadIdx.x < LCU_NPARS)
                sparams[threadIdx.x] = a.params[(size_t)b*LCU_NPARS + threadIdx.x];
            __syncthreads();
            if(threadIdx.x == 0)
                lcu_set_params_block(sdata, sparams);
        }
        else
        {
            for(int i = threadIdx.x; i < LCU_WORDS; i += LCU_BLOCK)
                sdata[i] = a.objs[(size_t)b*LCU_WORDS + i];
        }
        __syncthreads();
        data = sdata;
    }

    // pixel position, kernel/lensed.cl:24
    const float px = (float)(k % IMAGE_WIDTH);
    const float py = (float)(k / IMAGE_WIDTH);
    const float2 x = float2(__fadd_rn(a.pcs.x, __fmul_rn(a.pcs.z, px)),
                            __fadd_rn(a.pcs.y, __fmul_rn(a.pcs.w, py)));

    // value and error of quadrature, kernel/lensed.cl:27-32
    float f0 = 0, f1 = 0;

    if(S == 1)
    {
        // no branch on `live` around the loop: threads past the end of the range
        // shoot the rays of pixel k0 and drop the result, the loop stays in
        // uniform control flow and the object block in uniform registers
        {
#if LCU_QUAD_NI > 0
            // Cartesian rule (first axis outer, second inner, as the tables
            // are laid out): same points, same order, but written as two
            // loops so that everything that depends on the x abscissa only
            // (x - centre, its products with the rotation matrices, ...) is
            // computed once per grid column instead of once per ray
#pragma unroll 1
            for(int i = 0; i < LCU_QUAD_NI; ++i)
            {
                const float rx = __fadd_rn(x.x, lcu_quad[i*LCU_QUAD_NJ].x);
#pragma unroll 1
                for(int j = 0; j < LCU_QUAD_NJ; ++j)
                {
                    const float4 q = lcu_quad[i*LCU_QUAD_NJ + j];
                    const float c = lcu_compute(data, float2(rx, __fadd_rn(x.y, q.y)));
                    f0 = __fadd_rn(f0, __fmul_rn(q.z, c));
                    if(ERR)
                        f1 = __fadd_rn(f1, __fmul_rn(q.w, c));
                }
            }
#else
#pragma unroll 1
            for(int n = 0; n < QUAD_POINTS; ++n)
            {
                const float4 q = lcu_quad[n];
                const float c = lcu_compute(data, float2(__fadd_rn(x.x, q.x), __fadd_rn(x.y, q.y)));
                f0 = __fadd_rn(f0, __fmul_rn(q.z, c));
                if(ERR)
                    f1 = __fadd_rn(f1, __fmul_rn(q.w, c));
            }
#endif
        }
    }
    else
    {
        // S warps evaluate interleaved quadrature points of the same 32
        // pixels into shared memory; warp ns == 0 then adds them up in
        // ascending order, exactly as the single-warp loop would
        __shared__ float sc[P][LCU_CHUNK + 1];
        const int pl = pg*32 + lane;
        for(int c0 = 0; c0 < QUAD_POINTS; c0 += LCU_CHUNK)
        {
            const int c1 = min(c0 + LCU_CHUNK, QUAD_POINTS);
#if LCU_PAIR
            if constexpr(PAIRQ)
            {
#pragma unroll 1
                for(int n = c0 + 2*ns; n < c1; n += 2*S)
                {
                    const int n1 = min(n + 1, c1 - 1);      // odd tail: the second lane repeats the last point, result dropped
                    const float4 q0 = lcu_quad[n], q1 = lcu_quad[n1];
                    const lcu_pf c = lcu_compute2(data, lcu_pf2(lcu_pf(x.x) + lcu_pf(q0.x, q1.x), lcu_pf(x.y) + lcu_pf(q0.y, q1.y)));
                    sc[pl][n - c0] = c.lo();
                    if(n + 1 < c1)
                        sc[pl][n + 1 - c0] = c.hi();
                }
            }
            else
#endif
            {
#pragma unroll 1
                for(int n = c0 + ns; n < c1; n += S)
                {
                    const float4 q = lcu_quad[n];
                    sc[pl][n - c0] = lcu_compute(data, float2(__fadd_rn(x.x, q.x), __fadd_rn(x.y, q.y)));
                }
            }
            __syncthreads();
            if(ns == 0)
            {
                for(int n = c0; n < c1; ++n)
                {
                    const float4 q = lcu_quad[n];
                    const float c = sc[pl][n - c0];
                    f0 = __fadd_rn(f0, __fmul_rn(q.z, c));
                    if(ERR)
                        f1 = __fadd_rn(f1, __fmul_rn(q.w, c));
                }
            }
            __syncthreads();
        }
    }
    // the warps that only helped with the quadrature points have no output
    const bool writer = S == 1 || ns == 0;

    // outputs, kernel/lensed.cl:35-37, and the fused loglike kernel
    // (kernel/lensed.cl:41-53) when there is no PSF
    const size_t o = (size_t)b*IMAGE_SIZE + k;
    float chi = 0;
    if(live && writer)
    {
        if(a.mode & LCU_OUT_VALUE)
            a.value[o] = f0;
        if(ERR && (a.mode & LCU_OUT_ERROR))
            a.error[o] = f1;
        if(a.mode & (LCU_OUT_CHI2 | LCU_OUT_CHIMAP))
        {
            const float d = __fadd_rn(f0, -a.image[k]);
            chi = __fmul_rn(__fmul_rn(a.weight[k], d), d);
            if(a.mode & LCU_OUT_CHIMAP)
                a.chimap[o] = chi;
        }
    }
    if((a.mode & LCU_OUT_CHI2) && writer)
    {
        // fixed-shape tree over the 32 pixels of the group, in double
        double s = chi;
#pragma unroll
        for(int off = 16; off > 0; off >>= 1)
            s += __shfl_down_sync(0xffffffffu, s, off);
        const long long g = ((long long)bx*P + pg*32) >> 5;
        if(lane == 0 && g < a.ngroups)
            a.partial[(size_t)b*a.ngroups + g] = s;
    }
    if constexpr(S > 1 && !EXT)
        if(a.tail.out)
            lcu_fused_reduce(a.tail, b, gridDim.x, a.ngroups, a.partial);
}

// Resident blocks per SM the compiler budgets registers for: 3 (<= 85
// registers, 24 warps/SM) lets it keep every object field of the ray loop in
// registers; the kernel is issue-bound with ~5 eligible warps per scheduler,
// so the lower occupancy costs nothing (profiles/).
#ifndef LCU_RENDER_MINBLOCKS
#define LCU_RENDER_MINBLOCKS 3
#endif

#define LCU_RENDER_KERNEL(S) \
    extern "C" __global__ void __launch_bounds__(LCU_BLOCK, LCU_RENDER_MINBLOCKS) \
    lcu_render_s##S(const __grid_constant__ lcu_render_args a) { lcu_render_impl<S, false>(a); } \
    extern "C" __global__ void __launch_bounds__(LCU_BLOCK, LCU_RENDER_MINBLOCKS) \
    lcu_render_err_s##S(const __grid_constant__ lcu_render_args a) { lcu_render_impl<S, true>(a); }
LCU_RENDER_KERNEL(1)
LCU_RENDER_KERNEL(2)
LCU_RENDER_KERNEL(4)
LCU_RENDER_KERNEL(8)

// the split kernels with set_params folded in (small launches, likelihood path only); an
// experiment that measured slower than the separate kernel: compiled
// only into models created with LCU_FOLD_SETTER=1 in the environment
#if LCU_FOLD
#define LCU_RENDER_FOLD_KERNEL(S) \
    extern "C" __global__ void __launch_bounds__(LCU_BLOCK, LCU_RENDER_MINBLOCKS) \
    lcu_render_fold_s##S(const __grid_constant__ lcu_render_args a) { lcu_render_impl<S, false, true>(a); }
LCU_RENDER_FOLD_KERNEL(2)
LCU_RENDER_FOLD_KERNEL(4)
LCU_RENDER_FOLD_KERNEL(8)
#endif

#if LCU_PAIR
// the split kernels with two quadrature points per thread and pass (small launches of pairable models)
#define LCU_RENDER_Q_KERNEL(S) \
    extern "C" __global__ void __launch_bounds__(LCU_BLOCK, LCU_RENDER_MINBLOCKS) \
    lcu_render_q_s##S(const __grid_constant__ lcu_render_args a) { lcu_render_impl<S, false, false, true>(a); }
LCU_RENDER_Q_KERNEL(2)
LCU_RENDER_Q_KERNEL(4)
LCU_RENDER_Q_KERNEL(8)
#endif

#if LCU_PAIR
// Two rays per thread (shim.cuh: packed pairs).  A warp covers 64 consecutive
// pixels: lane l shoots the rays of pixels l and 32 + l through
// lcu_compute2(), whose adds and multiplies are FADD2 / FMUL2 / FFMA2 -- half
// the FP32 instructions per ray of the one-ray kernel.  Per pixel the arithmetic
// and its order are those of lcu_render_s1, and each half of the warp is one
// 32-pixel group of the chi^2 partial sums with the same lane layout, so the
// outputs are the same bits.
template<bool ERR>
__device__ __forceinline__ void lcu_render_pair_impl(const lcu_render_args& a)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.y;

    const long long kk0 = (long long)blockIdx.x*(2*LCU_BLOCK) + warp*64 + lane;
    const long long kk1 = kk0 + 32;
    const bool live0 = kk0 < a.nk, live1 = kk1 < a.nk;
    const long long k0 = a.k0 + (live0 ? kk0 : 0);
    const long long k1 = a.k0 + (live1 ? kk1 : 0);

#if LCU_OBJ_CONST
    const uint* data = reinterpret_cast<const uint*>(lcu_objs_c) + b*LCU_WORDS;
#else
    __shared__ __align__(16) uint sdata[LCU_WORDS];
    for(int i = threadIdx.x; i < LCU_WORDS; i += LCU_BLOCK)
        sdata[i] = a.objs[(size_t)b*LCU_WORDS + i];
    __syncthreads();
    const uint* data = sdata;
#endif

    // pixel positions, kernel/lensed.cl:24
    const lcu_pf px((float)(k0 % IMAGE_WIDTH), (float)(k1 % IMAGE_WIDTH));
    const lcu_pf py((float)(k0 / IMAGE_WIDTH), (float)(k1 / IMAGE_WIDTH));
    const lcu_pf2 x(lcu_pf(a.pcs.x) + lcu_pf(a.pcs.z)*px, lcu_pf(a.pcs.y) + lcu_pf(a.pcs.w)*py);

    // value and error of quadrature, kernel/lensed.cl:27-32
    // (threads past the end of the range shoot the rays of pixel k0 and drop the
    // result: no divergent branch around the loop, so that the compiler can keep
    // the object block in uniform registers)
    lcu_pf f0 = 0.0f, f1 = 0.0f;
    {
#if LCU_QUAD_NI > 0
#pragma unroll 1
        for(int i = 0; i < LCU_QUAD_NI; ++i)
        {
            const lcu_pf rx = x.x + lcu_pf(lcu_quad[i*LCU_QUAD_NJ].x);
#pragma unroll 1
            for(int j = 0; j < LCU_QUAD_NJ; ++j)
            {
                const float4 q = lcu_quad[i*LCU_QUAD_NJ + j];
                const lcu_pf c = lcu_compute2(data, lcu_pf2(rx, x.y + lcu_pf(q.y)));
                f0 = f0 + lcu_pf(q.z)*c;
                if(ERR)
                    f1 = f1 + lcu_pf(q.w)*c;
            }
        }
#else
#pragma unroll 1
        for(int n = 0; n < QUAD_POINTS; ++n)
        {
            const float4 q = lcu_quad[n];
            const lcu_pf c = lcu_compute2(data, lcu_pf2(x.x + lcu_pf(q.x), x.y + lcu_pf(q.y)));
            f0 = f0 + lcu_pf(q.z)*c;
            if(ERR)
                f1 = f1 + lcu_pf(q.w)*c;
        }
#endif
    }

    // outputs and the fused loglike kernel, as in lcu_render_impl
    const long long kk[2] = { kk0, kk1 };
    const long long k[2] = { k0, k1 };
    const bool live[2] = { live0, live1 };
    const float v0[2] = { f0.lo(), f0.hi() };
    const float v1[2] = { f1.lo(), f1.hi() };
#pragma unroll
    for(int h = 0; h < 2; ++h)
    {
        const size_t o = (size_t)b*IMAGE_SIZE + k[h];
        float chi = 0;
        if(live[h])
        {
            if(a.mode & LCU_OUT_VALUE)
                a.value[o] = v0[h];
            if(ERR && (a.mode & LCU_OUT_ERROR))
                a.error[o] = v1[h];
            if(a.mode & (LCU_OUT_CHI2 | LCU_OUT_CHIMAP))
            {
                const float d = __fadd_rn(v0[h], -a.image[k[h]]);
                chi = __fmul_rn(__fmul_rn(a.weight[k[h]], d), d);
                if(a.mode & LCU_OUT_CHIMAP)
                    a.chimap[o] = chi;
            }
        }
        if(a.mode & LCU_OUT_CHI2)
        {
            double s = chi;
#pragma unroll
            for(int off = 16; off > 0; off >>= 1)
                s += __shfl_down_sync(0xffffffffu, s, off);
            const long long g = (kk[h] - lane) >> 5;
            if(lane == 0 && g < a.ngroups)
                a.partial[(size_t)b*a.ngroups + g] = s;
        }
    }
}

// 3 resident blocks (<= 80 registers): with two rays per thread the kernel is
// bound by the FP32 pipe, which a packed instruction occupies for two cycles;
// 24 warps per SM keep it ~77 % busy, 16 warps ~73 % (profiles/)
#ifndef LCU_PAIR_MINBLOCKS
#define LCU_PAIR_MINBLOCKS 3
#endif
extern "C" __global__ void __launch_bounds__(LCU_BLOCK, LCU_PAIR_MINBLOCKS)
lcu_render_pair(const __grid_constant__ lcu_render_args a) { lcu_render_pair_impl<false>(a); }
extern "C" __global__ void __launch_bounds__(LCU_BLOCK, LCU_PAIR_MINBLOCKS)
lcu_render_pair_err(const __grid_constant__ lcu_render_args a) { lcu_render_pair_impl<true>(a); }
#endif // LCU_PAIR

// ---------------------------------------------------------------------------
// convolve + chi^2
// ---------------------------------------------------------------------------
#if PSF

// Output tile 64 x LCU_CT_H per block of 8 warps.  A warp covers 32 columns x
// 8 rows: lane = 4*ly + lx computes the 8 consecutive pixels x0 = 8*lx ... +7
// of row ly.  For each PSF row the thread loads the 8 + PSF_WIDTH - 1 input
// values it needs once (128-bit shared loads) and reuses them from registers
// for all PSF_WIDTH x 8 products, so shared-memory traffic is ~1/6 load per
// multiply-add instead of 1.  The row pitch is 4 (mod 8) words, which makes
// the quarter-warp 128-bit accesses bank-conflict free.
#define LCU_CT_W 64
#define LCU_CV_N ((8 + PSF_WIDTH - 1 + 3)/4)             // float4 loads per thread per PSF row
#define LCU_CW_MIN (56 + 4*LCU_CV_N)
#define LCU_CW (LCU_CW_MIN + ((12 - LCU_CW_MIN%8)%8))    // pitch = 4 (mod 8)
#if (LCU_CW*(32 + PSF_HEIGHT - 1)*4 <= 48*1024)
#define LCU_CT_H 32
#elif (LCU_CW*(16 + PSF_HEIGHT - 1)*4 <= 48*1024)
#define LCU_CT_H 16
#else
#define LCU_CT_H 8
#endif
#define LCU_CH (LCU_CT_H + PSF_HEIGHT - 1)
#define LCU_CT_PASSES (32/LCU_CT_H)

struct lcu_convolve_args
{
    const float* raw;       // [B][IMAGE_SIZE] rendered images
    float* model;           // [B][IMAGE_SIZE] convolved images or null
    const float* image;
    const float* weight;
    float* chimap;          // [B][IMAGE_SIZE] or null
    double* partial;        // [B][ngroups]
    int row0, row1;         // output rows [row0, row1)
    int ngroups;
    int gpr;                // groups per row = ceil(IMAGE_WIDTH/32)
    int mode;
    lcu_tail tail;
};

extern "C" __global__ void __launch_bounds__(LCU_BLOCK)
lcu_convolve(const __grid_constant__ lcu_convolve_args a)
{
    __shared__ __align__(16) float tile[LCU_CH][LCU_CW];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int lx = lane & 3, ly = lane >> 2;
    const int b = blockIdx.z;
    const int gx0 = blockIdx.x*LCU_CT_W;
    const int gy0 = a.row0 + blockIdx.y*LCU_CT_H;

    const float* raw = a.raw + (size_t)b*IMAGE_SIZE;

    // cache origin and fill, kernel/lensed.cl:73-83: edge-clamped window
    const int cx = gx0 - PSF_WIDTH/2;
    const int cy = gy0 - PSF_HEIGHT/2;
    for(int i = threadIdx.x; i < LCU_CH*LCU_CW; i += LCU_BLOCK)
    {
        const int r = i/LCU_CW, c = i%LCU_CW;
        const int yy = min(max(cy + r, 0), IMAGE_HEIGHT - 1);
        const int xx = min(max(cx + c, 0), IMAGE_WIDTH - 1);
        tile[r][c] = raw[(size_t)yy*IMAGE_WIDTH + xx];
    }
    __syncthreads();

    // this thread's 8 output pixels; warps that fall outside a reduced-height
    // tile have nothing to do
    const int x0 = (warp & 1)*32 + lx*8;
    const int ty = (warp >> 1)*8 + ly;
    if(ty >= LCU_CT_H)
        return;

    float acc[8];
#pragma unroll
    for(int r = 0; r < 8; ++r)
        acc[r] = 0;

    // reference order: PSF rows outer, columns inner (kernel/lensed.cl:95-97).
    // "p*v + acc" is one FMUL + one FADD in the default (strict) build, which
    // makes the result bit-identical to the CPU oracle, and contracts to FFMA
    // when the model is built with LCU_FAST_MATH.
#pragma unroll 1
    for(int j = 0; j < PSF_HEIGHT; ++j)
    {
        float v[4*LCU_CV_N];
        const float4* row = reinterpret_cast<const float4*>(&tile[ty + PSF_HEIGHT - 1 - j][x0]);
#pragma unroll
        for(int k = 0; k < LCU_CV_N; ++k)
        {
            const float4 t = row[k];
            v[4*k] = t.x; v[4*k + 1] = t.y; v[4*k + 2] = t.z; v[4*k + 3] = t.w;
        }
#pragma unroll
        for(int i = 0; i < PSF_WIDTH; ++i)
        {
            const float p = lcu_psf[j*PSF_WIDTH + i];
#pragma unroll
            for(int r = 0; r < 8; ++r)
                acc[r] = acc[r] + p*v[r + PSF_WIDTH - 1 - i];
        }
    }

    const int gj = gy0 + ty;
    const bool row_live = gj < a.row1;
    double s = 0;
#pragma unroll
    for(int r = 0; r < 8; ++r)
    {
        const int gi = gx0 + x0 + r;
        if(row_live && gi < IMAGE_WIDTH)
        {
            const size_t k = (size_t)gj*IMAGE_WIDTH + gi;
            const size_t o = (size_t)b*IMAGE_SIZE + k;
            if(a.mode & LCU_OUT_VALUE)
                a.model[o] = acc[r];
            // loglike kernel, kernel/lensed.cl:41-53
            const float d = __fadd_rn(acc[r], -a.image[k]);
            const float chi = __fmul_rn(__fmul_rn(a.weight[k], d), d);
            if(a.mode & LCU_OUT_CHIMAP)
                a.chimap[o] = chi;
            s += (double)chi;
        }
    }
    if(a.mode & LCU_OUT_CHI2)
    {
        // the 4 lanes lx = 0..3 hold one 32-pixel group of a row
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const int g = blockIdx.x*2 + (warp & 1);
        if(lx == 0 && row_live && g < a.gpr)
            a.partial[(size_t)b*a.ngroups + (size_t)(gj - a.row0)*a.gpr + g] = s;
    }
}

// The same convolution for launches whose 64 x 32 tiles would leave most of the
// machine idle (one point of a 100 x 100 image is 8 tiles on 148 SMs): a 32 x 8
// tile per block, one pixel per thread, every tap read from shared memory.
// Per pixel the products and sums are those of lcu_convolve in the same order,
// and the chi^2 partial of a 32-pixel group is added up in the same shape
// (eight consecutive pixels in sequence, then (s0 + s1) + (s2 + s3)), so both
// kernels produce the same bits and the host picks by grid size alone.
#define LCU_CS_W 32
#define LCU_CS_H (LCU_BLOCK/32)
#define LCU_CS_CW (LCU_CS_W + PSF_WIDTH - 1)
#define LCU_CS_CH (LCU_CS_H + PSF_HEIGHT - 1)

// EXT (lcu_point_*): tile indices from the caller, point 0, the rendered image
// read past L1 (other blocks of the same kernel wrote it), the caller runs the tail.
template<bool EXT>
__device__ __forceinline__ void lcu_convolve_small_impl(const lcu_convolve_args& a, unsigned ext_bx = 0, unsigned ext_by = 0)
{
    __shared__ float tile[LCU_CS_CH][LCU_CS_CW];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = EXT ? 0 : blockIdx.z;
    const unsigned bx = EXT ? ext_bx : blockIdx.x, by = EXT ? ext_by : blockIdx.y;
    const int gx0 = bx*LCU_CS_W;
    const int gy0 = a.row0 + by*LCU_CS_H;

    const float* raw = a.raw + (size_t)b*IMAGE_SIZE;

    // edge-clamped window, kernel/lensed.cl:73-83
    const int cx = gx0 - PSF_WIDTH/2;
    const int cy = gy0 - PSF_HEIGHT/2;
    for(int i = threadIdx.x; i < LCU_CS_CH*LCU_CS_CW; i += LCU_BLOCK)
    {
        const int r = i/LCU_CS_CW, c = i%LCU_CS_CW;
        const int yy = min(max(cy + r, 0), IMAGE_HEIGHT - 1);
        const int xx = min(max(cx + c, 0), IMAGE_WIDTH - 1);
        tile[r][c] = EXT ? __ldcg(raw + (size_t)yy*IMAGE_WIDTH + xx) : raw[(size_t)yy*IMAGE_WIDTH + xx];
    }
    __syncthreads();

    // PSF rows outer, columns inner (kernel/lensed.cl:95-97)
    float acc = 0;
#pragma unroll 1
    for(int j = 0; j < PSF_HEIGHT; ++j)
    {
        const float* row = &tile[warp + PSF_HEIGHT - 1 - j][lane + PSF_WIDTH - 1];
#pragma unroll
        for(int i = 0; i < PSF_WIDTH; ++i)
            acc = acc + lcu_psf[j*PSF_WIDTH + i]*row[-i];
    }

    const int gi = gx0 + lane;
    const int gj = gy0 + warp;
    const bool row_live = gj < a.row1;
    double c = 0;
    if(row_live && gi < IMAGE_WIDTH)
    {
        const size_t k = (size_t)gj*IMAGE_WIDTH + gi;
        const size_t o = (size_t)b*IMAGE_SIZE + k;
        if(a.mode & LCU_OUT_VALUE)
            a.model[o] = acc;
        // loglike kernel, kernel/lensed.cl:41-53
        const float d = __fadd_rn(acc, -a.image[k]);
        const float chi = __fmul_rn(__fmul_rn(a.weight[k], d), d);
        if(a.mode & LCU_OUT_CHIMAP)
            a.chimap[o] = chi;
        c = (double)chi;
    }
    if(a.mode & LCU_OUT_CHI2)
    {
        // lcu_convolve's summation shape: a thread's 8 consecutive pixels one
        // after the other, then the four threads of a group as (s0 + s1) + (s2 + s3)
        double s = 0;
#pragma unroll
        for(int r = 0; r < 8; ++r)
            s += __shfl_sync(0xffffffffu, c, (lane & 24) + r);
        const double s01 = __shfl_sync(0xffffffffu, s, 0) + __shfl_sync(0xffffffffu, s, 8);
        const double s23 = __shfl_sync(0xffffffffu, s, 16) + __shfl_sync(0xffffffffu, s, 24);
        const int g = bx;
        if(lane == 0 && row_live && g < a.gpr)
            a.partial[(size_t)b*a.ngroups + (size_t)(gj - a.row0)*a.gpr + g] = s01 + s23;
    }
    if constexpr(!EXT)
        if(a.tail.out)
            lcu_fused_reduce(a.tail, b, gridDim.x*gridDim.y, a.ngroups, a.partial);
}

extern "C" __global__ void __launch_bounds__(LCU_BLOCK)
lcu_convolve_small(const __grid_constant__ lcu_convolve_args a) { lcu_convolve_small_impl<false>(a); }

#endif // PSF

#if LCU_POINT
// ---------------------------------------------------------------------------
// One kernel for all stages of ONE point on a small image (the sampler's
// one-point-per-call pattern): set_params -> render -> convolve + chi^2 -> sum,
// what lcu_set_params, lcu_render[_q]_s{S}, lcu_convolve_small and the fused
// final reduction do as three dependent launches, with the same device code in
// the same order (the same bits) and two hand-overs through global memory instead
// of two kernel boundaries:
//   * block 0 runs set_params (one thread, as lcu_set_params does), publishes the
//     object block and raises a flag; the other blocks wait for the flag;
//   * every block renders its 32*8/S pixels; with a PSF it then takes a ticket:
//     the LAST conv_blocks arrivers wait until all blocks have arrived and each
//     convolves one 32 x 8 tile (the others are done); the last tile to finish
//     adds the chi^2 partial sums up and stores the log-likelihood.
// Progress: a block waits either for block 0's flag (block 0 is dispatched first)
// or, as one of the last conv_blocks arrivers, for blocks that are running or
// still to be dispatched -- at most conv_blocks blocks ever hold a slot while
// they wait, and the host uses this kernel only when that is a small fraction
// of what the device holds at once.  Every wait is nevertheless bounded by a
// cycle count, after which the block gives up: the result word then keeps its
// "pending" pattern and the host falls back to the three-kernel path for good.
// Compiled only into models created with LCU_FUSED_POINT=1 in the environment:
// measured, the two hand-overs cost what the two kernel boundaries cost (20.9 us
// against 20.0 us per call on a 100 x 100 image).
// ---------------------------------------------------------------------------
struct lcu_point_args
{
    lcu_render_args r;          // r.objs unused (the object block is computed here), r.params = the point's parameters
#if PSF
    lcu_convolve_args c;
#endif
    uint* objs;                 // [LCU_WORDS] scratch: the point's object block, written by block 0
    unsigned* sync;             // [2], zero between launches: object block ready, blocks that have rendered
    int conv_gx, conv_blocks;   // convolution tiles per row / in total (0 without a PSF)
};

__device__ __forceinline__ unsigned lcu_ld_acquire(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// thread 0 of the block waits until *p >= target; false after ~2^26 cycles (tens of milliseconds)
__device__ __forceinline__ bool lcu_wait_for(const unsigned* p, unsigned target)
{
    __shared__ int ok;
    if(threadIdx.x == 0)
    {
        const long long t0 = clock64();
        int good = 1;
        while(lcu_ld_acquire(p) < target)
            if(clock64() - t0 > (1ll << 26))
            {
                good = 0;
                break;
            }
        ok = good;
    }
    __syncthreads();
    return ok != 0;
}

template<int S, bool PAIRQ>
__device__ __forceinline__ void lcu_point_impl(const lcu_point_args& a)
{
    __shared__ __align__(16) uint sdata[LCU_WORDS];
    const unsigned G = gridDim.x;

    // stage 1: the object block (src/nested.c:77)
    if(blockIdx.x == 0)
    {
        __shared__ float sparams[LCU_NPARS > 0 ? LCU_NPARS : 1];
        if(threadIdx.x < LCU_NPARS)
            sparams[threadIdx.x] = a.r.params[threadIdx.x];
        __syncthreads();
        if(threadIdx.x == 0)
            lcu_set_params_block(sdata, sparams);
        __syncthreads();
        for(int i = threadIdx.x; i < LCU_WORDS; i += LCU_BLOCK)
            a.objs[i] = sdata[i];
        __threadfence();
        __syncthreads();
        if(threadIdx.x == 0 && a.conv_gx >= 0)     // conv_gx < 0: the host's test of the give-up path (the flag is never raised)
            atomicExch(a.sync + 0, 1u);
    }
    else
    {
        if(!lcu_wait_for(a.sync + 0, 1u))
            return;
        for(int i = threadIdx.x; i < LCU_WORDS; i += LCU_BLOCK)
            sdata[i] = __ldcg(a.objs + i);
        __syncthreads();
    }

    // stage 2: render (src/nested.c:84); without a PSF the chi^2 terms are fused in
    lcu_render_impl<S, false, false, PAIRQ, true>(a.r, sdata, blockIdx.x);

#if PSF
    // stage 3: the last conv_blocks blocks to get here convolve (src/nested.c:89-97)
    __shared__ unsigned ticket;
    __threadfence();
    __syncthreads();
    if(threadIdx.x == 0)
        ticket = atomicAdd(a.sync + 1, 1u);
    __syncthreads();
    const unsigned first = G - (unsigned)a.conv_blocks;
    if(ticket < first)
        return;
    if(!lcu_wait_for(a.sync + 1, G))
        return;
    const unsigned tile = ticket - first;
    lcu_convolve_small_impl<true>(a.c, tile % (unsigned)a.conv_gx, tile / (unsigned)a.conv_gx);
    const bool last = lcu_fused_reduce(a.c.tail, 0, (unsigned)a.conv_blocks, a.c.ngroups, a.c.partial);
#else
    const bool last = lcu_fused_reduce(a.r.tail, 0, G, a.r.ngroups, a.r.partial);
#endif
    // the block that stored the result leaves the hand-over words at zero for the next launch
    if(last && threadIdx.x == 0)
    {
        a.sync[0] = 0;
        a.sync[1] = 0;
    }
}

#define LCU_POINT_KERNEL(S) \
    extern "C" __global__ void __launch_bounds__(LCU_BLOCK, LCU_RENDER_MINBLOCKS) \
    lcu_point_s##S(const __grid_constant__ lcu_point_args a) { lcu_point_impl<S, LCU_PAIR != 0>(a); }
LCU_POINT_KERNEL(4)
LCU_POINT_KERNEL(8)
#endif // LCU_POINT

// ---------------------------------------------------------------------------
// data preparation: weight map from gain and offset (src/data.c:314-330), masked
// pixels weight 0 (src/lensed.c:470-482).  Double division, narrowed once.
// ---------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(LCU_BLOCK)
lcu_make_weight(long long n, const float* __restrict__ image, const float* __restrict__ gain_map, float gain,
                double offset, const int* __restrict__ mask, float* __restrict__ weight)
{
    for(long long i = (long long)blockIdx.x*LCU_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x*LCU_BLOCK)
    {
        const double g = gain_map ? gain_map[i] : gain;
        float w = (float)(g/((double)image[i] + offset));
        if(mask && mask[i])
            w = 0;
        weight[i] = w;
    }
}

// the same reduction as a kernel of its own: one block per point
extern "C" __global__ void __launch_bounds__(LCU_BLOCK)
lcu_reduce(int ngroups, const double* __restrict__ partial, double scale, double* __restrict__ out)
{
    lcu_reduce_block(ngroups, partial + (size_t)blockIdx.x*ngroups, scale, out + blockIdx.x);
}
