// lcu_bench.cu -- FFMA micro-benchmark: the measured FP32 (CUDA-core) peak that
// the render kernel's roofline fraction is quoted against.  MEASURED_PEAKS.json
// records HBM bandwidth and bf16 tensor throughput only; the hot path here is
// plain FP32 + SFU work, so its denominator has to be measured too.

#include <cuda_runtime.h>

namespace {

constexpr int CHAINS = 8;
constexpr int INNER = 4096;

__global__ void __launch_bounds__(256) ffma_kernel(float* out, float a, float b)
{
    float v[CHAINS];
#pragma unroll
    for(int i = 0; i < CHAINS; ++i)
        v[i] = threadIdx.x*0.001f + i;
    for(int it = 0; it < INNER; ++it)
    {
#pragma unroll
        for(int i = 0; i < CHAINS; ++i)
            v[i] = fmaf(v[i], a, b);
    }
    float s = 0;
#pragma unroll
    for(int i = 0; i < CHAINS; ++i)
        s += v[i];
    if(s == 123.456f)
        out[0] = s;
}

} // namespace

extern "C" int lcu_bench_ffma(int sm_count, double* tflops)
{
    float* out = nullptr;
    if(cudaMalloc(&out, sizeof(float)) != cudaSuccess)
        return 1;
    const int blocks = (sm_count > 0 ? sm_count : 148)*8*4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for(int rep = 0; rep < 5; ++rep)
    {
        cudaEventRecord(e0);
        ffma_kernel<<<blocks, 256>>>(out, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        if(cudaEventSynchronize(e1) != cudaSuccess)
        {
            cudaFree(out);
            return 1;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0*CHAINS*(double)INNER*256.0*blocks;
        const double t = flops/(ms*1e-3)/1e12;
        if(rep > 0 && t > best)
            best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return 0;
}
