// Throughput of the packed FP32 instructions (FFMA2/FADD2/FMUL2) against their
// scalar forms on sm_100a, alone and mixed with MUFU / ALU work: decides
// whether evaluating two rays per thread pays in the issue-bound render kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float mul1(float a, float b){ float r; asm volatile("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float add1(float a, float b){ float r; asm volatile("add.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ex2(float a){ float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float rcpa(float a){ float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float lg2a(float a){ float r; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 a){ float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x; }
__device__ __forceinline__ float hi(u64 a){ float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return y; }

#define N 8
#define ITERS 8192
// MODE 0: scalar FFMA x16 chains   1: FFMA2 x8 chains (same flops)
//      2: scalar FMUL+FADD         3: FMUL2+FADD2
//      4: scalar FFMA x16 + 4 MUFU 5: FFMA2 x8 + 4 MUFU (same work)
//      6: scalar 16 FFMA + 8 IADD/LOP  7: 8 FFMA2 + 8 IADD/LOP
//      8: FMUL2 x8   9: FADD2 x8   10: FFMA2 x8 + scalar FFMA x8   11: FFMA2 x8 + scalar FMUL x8
//      12: FMUL2 x8 + scalar FFMA x8  13: scalar FFMA x8   14: scalar FMUL x8
//      15: MUFU.EX2 x8   16: MUFU.RCP x8   17: MUFU.EX2 x8 + 16 ALU   18: MUFU.EX2 x8 + scalar FFMA x16
//      19: MUFU.EX2 x8 + FFMA2 x8   20: MUFU.EX2 x4 + MUFU.LG2 x4
//      21: FFMA2 x8, three distinct vector-register operands   22: FFMA2 x8, two distinct + one uniform
//      23: FMUL2 x8, two distinct vector-register operands     24: FADD2 x8, two distinct
//      25: scalar FFMA x16, three distinct vector-register operands
//      (modes 1, 8, 9 have one vector-register operand and uniform ones: the difference is what operand
//       fetch costs -- the render loop's packed instructions mostly have two or three vector operands)
// (the packed multiply is issued without .ftz: ptxas 12.9 contracts mul.ftz.f32x2 + add.ftz.f32x2 into FFMA2)
template<int MODE> __global__ void __launch_bounds__(256) kern(float* out, float s, float t)
{
    float a[2*N]; u64 p[N]; float m[4]; unsigned q[4]; float mm[8];
    for(int i = 0; i < 8; ++i) mm[i] = 0.3f + i*0.01f + threadIdx.x*1e-4f;
    for(int i = 0; i < 2*N; ++i) a[i] = threadIdx.x*1e-3f + i;
    for(int i = 0; i < N; ++i) p[i] = pk(a[2*i], a[2*i+1]);
    for(int i = 0; i < 4; ++i) { m[i] = i*0.1f + threadIdx.x*1e-4f; q[i] = threadIdx.x + i; }
    const u64 ss = pk(s, s), tt = pk(t, t);
#pragma unroll 1
    for(int it = 0; it < ITERS; ++it)
    {
        if(MODE == 0 || MODE == 4 || MODE == 6) {
#pragma unroll
            for(int i = 0; i < 2*N; ++i) a[i] = fma1(a[i], s, t);
        }
        if(MODE == 8 || MODE == 12) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = mul2(p[i], ss);
        }
        if(MODE == 9) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = add2(p[i], tt);
        }
        if(MODE == 10 || MODE == 12 || MODE == 13) {
#pragma unroll
            for(int i = 0; i < N; ++i) a[i] = fma1(a[i], s, t);
        }
        if(MODE == 11 || MODE == 14) {
#pragma unroll
            for(int i = 0; i < N; ++i) a[i] = mul1(a[i], s);
        }
        if(MODE == 1 || MODE == 5 || MODE == 7 || MODE == 10 || MODE == 11) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = fma2(p[i], ss, tt);
        }
        if(MODE == 2) {
#pragma unroll
            for(int i = 0; i < 2*N; ++i) a[i] = add1(mul1(a[i], s), t);
        }
        // operands from other chains: values stay bounded (|x| <= 1 keeps products and sums tame enough for timing)
        if(MODE == 21) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = fma2(p[i], p[(i + 3) % N], p[(i + 5) % N]);
        }
        if(MODE == 22) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = fma2(p[i], p[(i + 3) % N], tt);
        }
        if(MODE == 23) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = mul2(p[i], p[(i + 3) % N]);
        }
        if(MODE == 24) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = add2(p[i], p[(i + 3) % N]);
        }
        if(MODE == 25) {
#pragma unroll
            for(int i = 0; i < 2*N; ++i) a[i] = fma1(a[i], a[(i + 5) % (2*N)], a[(i + 11) % (2*N)]);
        }
        if(MODE == 3) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = add2(mul2(p[i], ss), tt);
        }
        if(MODE == 4 || MODE == 5) {
#pragma unroll
            for(int i = 0; i < 4; ++i) m[i] = ex2(m[i]);
        }
        if(MODE == 15 || MODE == 17 || MODE == 18 || MODE == 19) {
#pragma unroll
            for(int i = 0; i < 8; ++i) mm[i] = ex2(mm[i]);
        }
        if(MODE == 16) {
#pragma unroll
            for(int i = 0; i < 8; ++i) mm[i] = rcpa(mm[i]);
        }
        if(MODE == 20) {
#pragma unroll
            for(int i = 0; i < 4; ++i) { mm[i] = ex2(mm[i]); mm[i+4] = lg2a(mm[i+4]); }
        }
        if(MODE == 18) {
#pragma unroll
            for(int i = 0; i < 2*N; ++i) a[i] = fma1(a[i], s, t);
        }
        if(MODE == 19) {
#pragma unroll
            for(int i = 0; i < N; ++i) p[i] = fma2(p[i], ss, tt);
        }
        if(MODE == 17) {
#pragma unroll
            for(int i = 0; i < 4; ++i) { q[i] = (q[i] + 0x3f2aaaabu) & 0xff800fffu; q[i] ^= q[(i+1)&3] >> 3; q[i] = (q[i] + 0x3f2aaaabu) & 0xff800fffu; q[i] ^= q[(i+2)&3] >> 5; }
        }
        if(MODE == 6 || MODE == 7) {
#pragma unroll
            for(int i = 0; i < 4; ++i) { q[i] = (q[i] + 0x3f2aaaabu) & 0xff800fffu; q[i] ^= q[(i+1)&3] >> 3; }
        }
    }
    float r = 0;
    for(int i = 0; i < 2*N; ++i) r += a[i];
    for(int i = 0; i < N; ++i) r += lo(p[i]) + hi(p[i]);
    for(int i = 0; i < 4; ++i) r += m[i] + q[i];
    for(int i = 0; i < 8; ++i) r += mm[i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = r;
}

template<int MODE> void run(const char* name, float* out, double fp_per_iter, double other_per_iter)
{
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    const int blocks = pr.multiProcessorCount*8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for(int w = 0; w < 3; ++w) kern<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    const int reps = 5;
    for(int w = 0; w < reps; ++w) kern<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double threads = (double)blocks*256;
    const double lane_ops = threads*ITERS*fp_per_iter;      // FP32 lane operations (an FFMA2 counts 2)
    const double warp_instr_per_iter = 0;
    (void)warp_instr_per_iter;
    const double cyc = ms*1e-3*clk_khz*1e3;                 // at the nominal max clock
    printf("%-28s %8.3f ms  %7.2f T lane-FP-ops/s  %6.1f lane-FP-ops/clk/SM (nominal clk)  other/iter %.0f\n",
           name, ms, lane_ops/ms*1e-9, lane_ops/cyc/pr.multiProcessorCount, other_per_iter);
}

int main()
{
    float* out; cudaMalloc(&out, 148*8*256*4*4);
    run<0>("scalar FFMA x16", out, 16, 0);
    run<1>("FFMA2 x8", out, 16, 0);
    run<2>("scalar FMUL+FADD x16", out, 32, 0);
    run<3>("FMUL2+FADD2 x8", out, 32, 0);
    run<4>("scalar FFMA x16 + 4 MUFU", out, 16, 4);
    run<5>("FFMA2 x8 + 4 MUFU", out, 16, 4);
    run<6>("scalar FFMA x16 + 8 ALU", out, 16, 8);
    run<7>("FFMA2 x8 + 8 ALU", out, 16, 8);
    run<8>("FMUL2 x8", out, 16, 0);
    run<9>("FADD2 x8", out, 16, 0);
    run<13>("scalar FFMA x8", out, 8, 0);
    run<14>("scalar FMUL x8", out, 8, 0);
    run<10>("FFMA2 x8 + scalar FFMA x8", out, 24, 0);
    run<11>("FFMA2 x8 + scalar FMUL x8", out, 24, 0);
    run<12>("FMUL2 x8 + scalar FFMA x8", out, 24, 0);
    run<21>("FFMA2 x8, 3 vector operands", out, 16, 0);
    run<22>("FFMA2 x8, 2 vector + 1 uniform", out, 16, 0);
    run<23>("FMUL2 x8, 2 vector operands", out, 16, 0);
    run<24>("FADD2 x8, 2 vector operands", out, 16, 0);
    run<25>("scalar FFMA x16, 3 vector ops", out, 16, 0);
    run<0>("scalar FFMA x16 (again)", out, 16, 0);
    run<15>("MUFU.EX2 x8", out, 0, 8);
    run<16>("MUFU.RCP x8", out, 0, 8);
    run<20>("MUFU.EX2 x4 + LG2 x4", out, 0, 8);
    run<17>("MUFU.EX2 x8 + 16 ALU", out, 0, 24);
    run<18>("MUFU.EX2 x8 + scalar FFMA x16", out, 16, 8);
    run<19>("MUFU.EX2 x8 + FFMA2 x8", out, 16, 8);
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
