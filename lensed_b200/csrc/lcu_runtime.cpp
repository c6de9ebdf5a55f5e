// lcu_runtime.cpp -- contexts, models, launches: the C ABI of include/lensed_cuda.h.
//
// Absorbs the raw OpenCL call sites of the reference: device/context set-up
// (src/opencl.c:132-241), program build, buffers, kernel arguments and work
// sizes (src/lensed.c:644-1112), the per-evaluation enqueue sequence and
// result read-back (src/nested.c:63-115), the dumper's re-render
// (src/nested.c:178-214) and the profiler (src/profile.c).
//
// Device code is compiled per model by NVRTC (lcu_program.cpp) and loaded
// through the driver API, whose entry points are resolved at run time with
// cudaGetDriverEntryPoint so that this library has no link-time dependency
// on libcuda and loads on machines without a GPU (compile-only contexts).

#include "lcu_internal.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

using namespace lcu;

extern "C" int lcu_bench_ffma(int sm_count, double* tflops);     // lcu_bench.cu

namespace {

std::atomic<unsigned long long> g_launches{0};

// ---- driver API, resolved lazily ---------------------------------------------
struct Driver
{
    bool ready = false;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
} drv;

template<typename F>
bool resolve(const char* name, F* fn)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p)
    {
        set_error("CUDA driver entry point %s not available", name);
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

bool init_driver()
{
    if(drv.ready)
        return true;
    if(!resolve("cuModuleLoadData", &drv.ModuleLoadData) || !resolve("cuModuleUnload", &drv.ModuleUnload)
       || !resolve("cuModuleGetFunction", &drv.ModuleGetFunction)
       || !resolve("cuModuleGetGlobal", &drv.ModuleGetGlobal)
       || !resolve("cuLaunchKernel", &drv.LaunchKernel) || !resolve("cuGetErrorString", &drv.GetErrorString)
       || !resolve("cuOccupancyMaxActiveBlocksPerMultiprocessor", &drv.OccupancyMaxActiveBlocks))
        return false;
    drv.ready = true;
    return true;
}

const char* cu_str(CUresult r)
{
    const char* s = nullptr;
    if(drv.GetErrorString && drv.GetErrorString(r, &s) == CUDA_SUCCESS && s)
        return s;
    return "unknown driver error";
}

#define RT_CHECK(expr) \
    do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { \
        set_error("%s: %s", #expr, cudaGetErrorString(e_)); return LCU_E_CUDA; } } while(0)

#define DRV_CHECK(expr) \
    do { CUresult r_ = (expr); if(r_ != CUDA_SUCCESS) { \
        set_error("%s: %s", #expr, cu_str(r_)); return LCU_E_CUDA; } } while(0)

std::string library_dir()
{
    Dl_info info;
    if(dladdr(reinterpret_cast<void*>(&library_dir), &info) && info.dli_fname)
    {
        std::string p = info.dli_fname;
        const size_t s = p.rfind('/');
        return s == std::string::npos ? std::string(".") : p.substr(0, s);
    }
    return ".";
}

size_t div_up(size_t a, size_t b) { return (a + b - 1)/b; }

// device-side argument blocks: must match kernel/lensed.cu
struct Tail
{
    double* out;
    unsigned* counter;
    double scale;
};

struct RenderArgs
{
    float pcs[4];
    long long k0, nk;
    const uint32_t* objs;
    const float* params;
    float* value;
    float* error;
    const float* image;
    const float* weight;
    float* chimap;
    double* partial;
    int ngroups;
    int mode;
    Tail tail;
};

struct ConvolveArgs
{
    const float* raw;
    float* model;
    const float* image;
    const float* weight;
    float* chimap;
    double* partial;
    int row0, row1;
    int ngroups;
    int gpr;
    int mode;
    Tail tail;
};

// lcu_point_args of kernel/lensed.cu: with and without its PSF member
struct PointArgsPsf
{
    RenderArgs r;
    ConvolveArgs c;
    uint32_t* objs;
    unsigned* sync;
    int conv_gx, conv_blocks;
};

struct PointArgsNoPsf
{
    RenderArgs r;
    uint32_t* objs;
    unsigned* sync;
    int conv_gx, conv_blocks;
};

enum { OUT_VALUE = 1, OUT_ERROR = 2, OUT_CHI2 = 4, OUT_CHIMAP = 8 };

} // namespace

struct lcu_model
{
    lcu_ctx* ctx = nullptr;
    std::vector<ModelObject> objs;
    size_t npars = 0, words = 0;
    size_t width = 0, height = 0, size = 0;
    float pcs[4] = { 1, 1, 1, 1 };
    size_t nq = 0;
    bool has_psf = false;
    size_t psfw = 0, psfh = 0;
    unsigned flags = 0;
    bool obj_const = true;
    size_t maxb = 1;
    size_t row0 = 0, row1 = 0;
    size_t conv_tile_h = 32;            // must match LCU_CT_H in kernel/lensed.cu
    std::string source, log;
    std::vector<char> cubin;

    // device state
    CUmodule mod = nullptr;
    CUfunction f_set = nullptr, f_render[4] = {}, f_render_err[4] = {}, f_conv = nullptr, f_conv_small = nullptr, f_reduce = nullptr;
    CUfunction f_render_pair = nullptr, f_render_pair_err = nullptr;   // two rays per thread, if pair
    CUfunction f_render_fold[4] = {};                                   // split kernels with set_params folded in (if fold)
    bool fold = false;                                                  // LCU_FOLD_SETTER=1 when the model was created
    CUfunction f_render_q[4] = {};                                      // split kernels, two quadrature points per pass (if pair)
    CUfunction f_point[4] = {};                                         // all stages of one point in one kernel (split 4 and 8)
    size_t point_capacity[4] = {};                                      // blocks of f_point[] the device holds at once
    bool point = false;                                                 // LCU_FUSED_POINT=1 when the model was created
    bool point_off = false;                                             // a wait inside the kernel timed out once: not used again
    bool point_ok = false;                                              // set while capturing a graph whose result word the host watches
    unsigned* d_sync = nullptr;                                         // [2] hand-over words of f_point, zero between launches
    CUfunction f_make_weight = nullptr;
    bool pair = false;
    CUdeviceptr c_objs = 0;
    cudaStream_t stream = nullptr;
    float *d_image = nullptr, *d_weight = nullptr;
    uint32_t* d_objs = nullptr;         // [maxb][words]
    float* d_raw = nullptr;             // [raw_cap][size] rendered images, PSF models only; grows on demand
    size_t raw_cap = 0;
    double* d_partial = nullptr;        // [maxb][max groups]
    unsigned* d_counter = nullptr;      // [maxb] finished blocks per point (fused reduction), zero between launches
    size_t partial_cap = 0;
    // staging for the host entry points
    float *d_params = nullptr, *h_params = nullptr;
    double *d_lnew = nullptr, *h_lnew = nullptr;
    size_t stage_cap = 0;
    // single-point evaluation (the sampler's one-point callback) as a CUDA graph:
    // upload, 3-4 kernels, constant-bank copy and read-back replayed by one launch
    cudaGraphExec_t graph1 = nullptr;
    cudaGraphExec_t graph1b = nullptr;  // the same graph on the second staging slot (lcu_loglike_async)
    int async_next = 0;                 // slot the next lcu_loglike_async takes
    bool async_busy[2] = { false, false };
    bool async_redone[2] = { false, false };    // the ticket's point was re-evaluated synchronously: answer in async_redo
    double async_redo[2] = { 0, 0 };
    cudaEvent_t async_done[2] = { nullptr, nullptr };   // completion of a slot's evaluation when it is not watched
    bool graph1_off = false;
    size_t graph1_rows[2] = { 0, 0 };
    int graph1_split = 0, graph1_conv = -1;
    unsigned graph1_nodes = 0;          // kernel nodes of the graph
    bool graph1_mapped = false;         // the graph writes its result straight into h_lnew
    // dumper buffers (one point), allocated on first lcu_render
    float *d_value1 = nullptr, *d_error1 = nullptr, *d_model1 = nullptr, *d_chi1 = nullptr;
    // profiling: one set of stage events per launched chunk, harvested lazily
    struct EventSet { cudaEvent_t e[5]; size_t nb; };
    bool profile = false;
    lcu_profile prof = {};
    std::vector<EventSet> evsets;       // pool
    size_t evused = 0;                  // sets recorded since the last harvest
    cudaEvent_t ev_io[4] = {};          // host path: upload begin/end, download begin/end
};

namespace {

// the single-point graphs refer to the staging / scratch buffers: drop them before any of those moves
void drop_point_graphs(lcu_model* m)
{
    if(m->graph1) cudaGraphExecDestroy(m->graph1);
    if(m->graph1b) cudaGraphExecDestroy(m->graph1b);
    m->graph1 = nullptr;
    m->graph1b = nullptr;
}

int launch(lcu_model* m, CUfunction f, dim3 grid, dim3 block, void** args, cudaStream_t stream)
{
    DRV_CHECK(drv.LaunchKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, 0,
                               reinterpret_cast<CUstream>(stream), args, nullptr));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    (void)m;
    return LCU_OK;
}

// rows that have to be rendered so that rows [row0, row1) can be convolved:
// the PSF window of kernel/lensed.cl:73-76,97 spans rows gj - Ph/2 ... gj - Ph/2 + Ph - 1
void render_rows(const lcu_model* m, size_t* r0, size_t* r1)
{
    if(!m->has_psf)
    {
        *r0 = m->row0;
        *r1 = m->row1;
        return;
    }
    const long long lo = (long long)m->row0 - (long long)(m->psfh/2);
    const long long hi = (long long)m->row1 - 1 - (long long)(m->psfh/2) + (long long)m->psfh;   // exclusive
    *r0 = (size_t)std::max<long long>(lo, 0);
    *r1 = (size_t)std::min<long long>(hi, (long long)m->height);
}

size_t group_count(const lcu_model* m)
{
    if(m->has_psf)
        return (m->row1 - m->row0)*div_up(m->width, 32);
    size_t r0, r1;
    render_rows(m, &r0, &r1);
    return div_up((r1 - r0)*m->width, 32);
}

// chi^2 partial sums of up to nb points per launch ([nb][groups] doubles): sized
// by the largest batch actually launched, not by maxb (a 4096^2 model without a
// PSF has 524288 groups per point: 4 MB per point)
int ensure_partial(lcu_model* m, size_t nb)
{
    nb = std::min(std::max<size_t>(nb, 1), m->maxb);
    const size_t need = nb*group_count(m);
    if(need <= m->partial_cap)
        return LCU_OK;
    drop_point_graphs(m);                   // the graphs refer to the old buffer
    if(m->d_partial)
        cudaFree(m->d_partial);
    m->d_partial = nullptr;
    m->partial_cap = 0;
    RT_CHECK(cudaMalloc(&m->d_partial, need*sizeof(double)));
    m->partial_cap = need;
    return LCU_OK;
}

// staging for the rendered (pre-PSF) images of up to nb points per launch
int ensure_raw(lcu_model* m, size_t nb)
{
    if(!m->has_psf || nb <= m->raw_cap)
        return LCU_OK;
    drop_point_graphs(m);                   // the graphs refer to the old buffer
    if(m->d_raw)
        cudaFree(m->d_raw);
    m->d_raw = nullptr;
    m->raw_cap = 0;
    RT_CHECK(cudaMalloc(&m->d_raw, nb*m->size*sizeof(float)));
    m->raw_cap = nb;
    return LCU_OK;
}

int ensure_stage(lcu_model* m, size_t nbatch)
{
    if(nbatch <= m->stage_cap)
        return LCU_OK;
    drop_point_graphs(m);                   // the single-point graphs refer to the staging buffers
    if(m->d_params) cudaFree(m->d_params);
    if(m->d_lnew) cudaFree(m->d_lnew);
    if(m->h_params) cudaFreeHost(m->h_params);
    if(m->h_lnew) cudaFreeHost(m->h_lnew);
    m->d_params = nullptr; m->d_lnew = nullptr; m->h_params = nullptr; m->h_lnew = nullptr;
    m->stage_cap = 0;
    const size_t cap = std::max<size_t>(nbatch, 64);
    RT_CHECK(cudaMalloc(&m->d_params, cap*std::max<size_t>(m->npars, 1)*sizeof(float)));
    RT_CHECK(cudaMalloc(&m->d_lnew, cap*sizeof(double)));
    RT_CHECK(cudaMallocHost(&m->h_params, cap*std::max<size_t>(m->npars, 1)*sizeof(float)));
    RT_CHECK(cudaMallocHost(&m->h_lnew, cap*sizeof(double)));
    m->stage_cap = cap;
    return LCU_OK;
}

// how many warps share one 32-pixel group's quadrature points: enough to put
// ~1024 threads on every SM even for small images / batches
int pick_split(const lcu_model* m, size_t npix, size_t nb)
{
    const char* force = getenv("LCU_SPLIT");
    if(force && *force)
    {
        const int s = atoi(force);
        if(s == 1 || s == 2 || s == 4 || s == 8)
            return s;
    }
    const size_t target = (size_t)std::max(m->ctx->sm_count, 1)*1024;
    int s = 1;
    while(s < 8 && npix*nb*(size_t)s < target)
        s *= 2;
    return s;
}

// enqueue set_params -> render -> (convolve) for `nb` points whose
// parameters start at d_params; per-point outputs are optional
// `reduce_out` (optional, with want_chi2): where scale * sum of each point's
// chi^2 partials goes if the kernel that writes the partials can add them up
// itself (the small-launch kernels); *reduced says whether it did
int enqueue_points(lcu_model* m, size_t nb, const float* d_params, cudaStream_t st,
                   float* value, float* error, float* model, float* chimap, bool want_chi2,
                   cudaEvent_t* ev, double* reduce_out = nullptr, double scale = 1, bool* reduced = nullptr)
{
    if(reduced) *reduced = false;
    const bool may_fuse = want_chi2 && reduce_out && reduced && !getenv("LCU_NO_FUSED_REDUCE");
    if(ev) cudaEventRecord(ev[0], st);
    size_t r0, r1;
    render_rows(m, &r0, &r1);
    const size_t nk = (r1 - r0)*m->width;
    const int ngroups = (int)group_count(m);
    const int split = pick_split(m, nk, nb);
    // Opt-in (LCU_FOLD_SETTER=1): for small launches (one to four points of a small
    // image) the split render kernels run set_params themselves, once per block, which
    // takes one kernel and one dependency out of the launch sequence
    // (lcu_set_params_block: the same code, the same bits).  Measured slower than the
    // separate kernel, hence not the default.
    const bool fold = m->fold && split > 1 && nb <= 4 && !error;
    // Opt-in (LCU_FUSED_POINT=1 when the model is created): one point of a small image
    // with nothing but the log-likelihood asked for runs all stages in one kernel
    // (lcu_point_s*, kernel/lensed.cu).  Correct and tested, but measured ~1 us SLOWER
    // than the three launches, hence not the default.
    if(nb == 1 && (split == 4 || split == 8) && want_chi2 && may_fuse && !value && !error && !model && !chimap && !fold && !ev
       && m->point && m->point_ok && !m->point_off)
    {
        const int idx = split == 4 ? 2 : 3;
        const size_t grid = div_up(nk, 256/(size_t)split);
        const size_t rows = m->row1 - m->row0;
        const size_t conv_gx = div_up(m->width, 32), conv_blocks = m->has_psf ? conv_gx*div_up(rows, 8) : 0;
        const char* force = getenv("LCU_CONV_SMALL");
        const bool conv_small_ok = !m->has_psf || !(force && *force == '0');
        // at most conv_blocks blocks wait while they hold a slot: a quarter of the device's capacity at most
        if(4*(conv_blocks + 1) <= m->point_capacity[idx] && conv_blocks <= grid && conv_small_ok)
        {
            const bool test_giveup = getenv("LCU_POINT_TEST_TIMEOUT") != nullptr;      // tests: block 0 never raises its flag
            RenderArgs r;
            memcpy(r.pcs, m->pcs, sizeof(r.pcs));
            r.k0 = (long long)(r0*m->width);
            r.nk = (long long)nk;
            r.objs = nullptr;
            r.params = d_params;
            r.value = m->has_psf ? m->d_raw : nullptr;
            r.error = nullptr;
            r.image = m->d_image;
            r.weight = m->d_weight;
            r.chimap = nullptr;
            r.partial = m->d_partial;
            r.ngroups = ngroups;
            r.mode = m->has_psf ? OUT_VALUE : OUT_CHI2;
            r.tail = m->has_psf ? Tail{ nullptr, nullptr, 0 } : Tail{ reduce_out, m->d_counter, scale };
            int rc;
            if(m->has_psf)
            {
                PointArgsPsf a;
                a.r = r;
                a.c.raw = m->d_raw;
                a.c.model = nullptr;
                a.c.image = m->d_image;
                a.c.weight = m->d_weight;
                a.c.chimap = nullptr;
                a.c.partial = m->d_partial;
                a.c.row0 = (int)m->row0;
                a.c.row1 = (int)m->row1;
                a.c.ngroups = ngroups;
                a.c.gpr = (int)div_up(m->width, 32);
                a.c.mode = OUT_CHI2;
                a.c.tail = Tail{ reduce_out, m->d_counter, scale };
                a.objs = m->d_objs;
                a.sync = m->d_sync;
                a.conv_gx = test_giveup ? -1 : (int)conv_gx;
                a.conv_blocks = (int)conv_blocks;
                void* args[] = { &a };
                rc = launch(m, m->f_point[idx], dim3((unsigned)grid), dim3(256), args, st);
            }
            else
            {
                PointArgsNoPsf a;
                a.r = r;
                a.objs = m->d_objs;
                a.sync = m->d_sync;
                a.conv_gx = test_giveup ? -1 : 0;
                a.conv_blocks = 0;
                void* args[] = { &a };
                rc = launch(m, m->f_point[idx], dim3((unsigned)grid), dim3(256), args, st);
            }
            if(rc) return rc;
            *reduced = true;
            return LCU_OK;
        }
    }
    // set_params, src/nested.c:77
    if(!fold)
    {
        int B = (int)nb;
        void* args[] = { &B, (void*)&d_params, &m->d_objs };
        int rc = launch(m, m->f_set, dim3((unsigned)div_up(nb, 64)), dim3(64), args, st);
        if(rc) return rc;
    }

    // the large-image kernels read the object blocks from the constant bank;
    // the split kernels (small images) take them from global memory themselves
    if(m->obj_const && split == 1)
        RT_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(m->c_objs), m->d_objs,
                                 nb*m->words*sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    if(ev) cudaEventRecord(ev[1], st);

    // render, src/nested.c:84
    {
        RenderArgs a;
        memcpy(a.pcs, m->pcs, sizeof(a.pcs));
        a.k0 = (long long)(r0*m->width);
        a.nk = (long long)nk;
        a.objs = m->d_objs;
        a.params = fold ? d_params : nullptr;
        a.value = m->has_psf ? (value ? value : m->d_raw) : value;
        a.error = error;
        a.image = m->d_image;
        a.weight = m->d_weight;
        a.chimap = m->has_psf ? nullptr : chimap;
        a.partial = m->d_partial;
        a.ngroups = ngroups;
        a.mode = 0;
        a.tail = Tail{ nullptr, nullptr, 0 };
        if(may_fuse && !m->has_psf && split > 1)
        {
            a.tail = Tail{ reduce_out, m->d_counter, scale };
            *reduced = true;
        }
        if(a.value) a.mode |= OUT_VALUE;
        if(a.error) a.mode |= OUT_ERROR;
        if(!m->has_psf)
        {
            if(want_chi2) a.mode |= OUT_CHI2;
            if(a.chimap) a.mode |= OUT_CHIMAP;
        }
        const int idx = split == 1 ? 0 : split == 2 ? 1 : split == 4 ? 2 : 3;
        void* args[] = { &a };
        int rc;
        if(split == 1 && m->pair)
            // two pixels per thread: 512 per block
            rc = launch(m, a.error ? m->f_render_pair_err : m->f_render_pair, dim3((unsigned)div_up(nk, 512), (unsigned)nb),
                        dim3(256), args, st);
        else
        {
            // split kernels: pairable models shoot two quadrature points per pass (LCU_NO_SPLIT_PAIR: one)
            const bool q2 = split > 1 && m->pair && !a.error && !fold && !getenv("LCU_NO_SPLIT_PAIR");
            rc = launch(m, a.error ? m->f_render_err[idx] : fold ? m->f_render_fold[idx] : q2 ? m->f_render_q[idx] : m->f_render[idx],
                        dim3((unsigned)div_up(nk, 256/split), (unsigned)nb), dim3(256), args, st);
        }
        if(rc) return rc;
    }
    if(ev) cudaEventRecord(ev[2], st);

    // convolve + loglike, src/nested.c:89-97
    if(m->has_psf)
    {
        ConvolveArgs c;
        c.raw = value ? value : m->d_raw;
        c.model = model;
        c.image = m->d_image;
        c.weight = m->d_weight;
        c.chimap = chimap;
        c.partial = m->d_partial;
        c.row0 = (int)m->row0;
        c.row1 = (int)m->row1;
        c.ngroups = ngroups;
        c.gpr = (int)div_up(m->width, 32);
        c.mode = 0;
        if(model) c.mode |= OUT_VALUE;
        if(chimap) c.mode |= OUT_CHIMAP;
        if(want_chi2) c.mode |= OUT_CHI2;
        void* args[] = { &c };
        // 64 x conv_tile_h tiles with register reuse; 32 x 8 tiles (same bits) when
        // those would not give every SM a block (LCU_CONV_SMALL=0/1 forces the choice)
        const size_t rows = m->row1 - m->row0;
        const size_t big = div_up(m->width, 64)*div_up(rows, m->conv_tile_h)*nb;
        const char* force = getenv("LCU_CONV_SMALL");
        const bool small = force && *force ? *force == '1' : big < (size_t)std::max(m->ctx->sm_count, 1);
        c.tail = Tail{ nullptr, nullptr, 0 };
        if(may_fuse && small)
        {
            c.tail = Tail{ reduce_out, m->d_counter, scale };
            *reduced = true;
        }
        int rc = small
            ? launch(m, m->f_conv_small, dim3((unsigned)div_up(m->width, 32), (unsigned)div_up(rows, 8), (unsigned)nb), dim3(256), args, st)
            : launch(m, m->f_conv, dim3((unsigned)div_up(m->width, 64), (unsigned)div_up(rows, m->conv_tile_h), (unsigned)nb),
                     dim3(256), args, st);
        if(rc) return rc;
    }
    if(ev) cudaEventRecord(ev[3], st);
    return LCU_OK;
}

// stage events of the next chunk (profiling only); the pool grows on demand
cudaEvent_t* next_event_set(lcu_model* m, size_t nb)
{
    if(!m->profile)
        return nullptr;
    if(m->evused == m->evsets.size())
    {
        lcu_model::EventSet es;
        es.nb = 0;
        for(cudaEvent_t& e : es.e)
            if(cudaEventCreate(&e) != cudaSuccess)
                return nullptr;
        m->evsets.push_back(es);
    }
    lcu_model::EventSet& es = m->evsets[m->evused++];
    es.nb = nb;
    return es.e;
}

// wait for the recorded chunks and add their stage times to the profile
void harvest_profile(lcu_model* m)
{
    for(size_t i = 0; i < m->evused; ++i)
    {
        lcu_model::EventSet& es = m->evsets[i];
        float t;
        if(cudaEventSynchronize(es.e[4]) != cudaSuccess)
            continue;
        m->prof.evaluations += es.nb;
        if(cudaEventElapsedTime(&t, es.e[0], es.e[1]) == cudaSuccess) m->prof.set_params_ms += t;
        if(cudaEventElapsedTime(&t, es.e[1], es.e[2]) == cudaSuccess) m->prof.render_ms += t;
        if(cudaEventElapsedTime(&t, es.e[2], es.e[3]) == cudaSuccess) m->prof.convolve_ms += t;
        if(cudaEventElapsedTime(&t, es.e[3], es.e[4]) == cudaSuccess) m->prof.reduce_ms += t;
    }
    m->evused = 0;
}

int enqueue_batch(lcu_model* m, size_t nbatch, const float* d_params, double* d_lnew, cudaStream_t st)
{
    int rc = ensure_partial(m, std::min(m->maxb, nbatch));
    if(rc) return rc;
    rc = ensure_raw(m, std::min(m->maxb, nbatch));
    if(rc) return rc;
    const int ngroups = (int)group_count(m);
    for(size_t b0 = 0; b0 < nbatch; b0 += m->maxb)
    {
        const size_t nb = std::min(m->maxb, nbatch - b0);
        cudaEvent_t* ev = next_event_set(m, nb);
        // host sum of src/nested.c:106-115, on the device: by the last block of the
        // kernel that writes the partials (small launches) or by a kernel of its own
        double scale = -0.5;
        double* out = d_lnew + b0;
        bool reduced = false;
        rc = enqueue_points(m, nb, d_params + b0*m->npars, st, nullptr, nullptr, nullptr, nullptr, true, ev,
                            out, scale, &reduced);
        if(rc) return rc;
        if(!reduced)
        {
            int ng = ngroups;
            void* args[] = { &ng, &m->d_partial, &scale, &out };
            rc = launch(m, m->f_reduce, dim3((unsigned)nb), dim3(256), args, st);
            if(rc) return rc;
        }
        if(ev) cudaEventRecord(ev[4], st);
    }
    return LCU_OK;
}

// a wait inside lcu_point_* timed out: leave its hand-over words clean and do not use it again
int point_kernel_failed(lcu_model* m)
{
    RT_CHECK(cudaStreamSynchronize(m->stream));
    m->point_off = true;
    drop_point_graphs(m);
    RT_CHECK(cudaMemset(m->d_sync, 0, 2*sizeof(unsigned)));
    RT_CHECK(cudaMemset(m->d_counter, 0, m->maxb*sizeof(unsigned)));
    return LCU_OK;
}

// (re)build the single-point graph if the launch configuration changed;
// returns false when graphs are unavailable (the plain path is used then)
bool single_point_graph(lcu_model* m)
{
    if(m->graph1_off)
        return false;
    const char* forced = getenv("LCU_SPLIT");
    const int split = forced && *forced ? atoi(forced) : 0;
    const char* cforced = getenv("LCU_CONV_SMALL");
    const int conv = cforced && *cforced ? (*cforced == '1') : -1;
    if(m->graph1 && m->graph1_rows[0] == m->row0 && m->graph1_rows[1] == m->row1 && m->graph1_split == split
       && m->graph1_conv == conv)
        return true;
    drop_point_graphs(m);
    if(getenv("LCU_NO_GRAPH") || ensure_partial(m, 1) != LCU_OK || ensure_raw(m, 1) != LCU_OK)
    {
        m->graph1_off = true;
        return false;
    }
    // one graph per staging slot: slot 0 serves lcu_loglike, both serve lcu_loglike_async
    bool mapped = false;
    for(int slot = 0; slot < 2; ++slot)
    {
        cudaGraph_t graph = nullptr;
        const unsigned long long launches0 = g_launches.load();
        bool ok = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if(ok)
        {
            // The pinned staging buffers are mapped into the device's address space
            // (unified addressing): set_params reads the point's few parameters straight
            // from host memory and the reduction writes the result there, which takes
            // the two copy nodes (and their dependencies) out of the graph.  LCU_GRAPH_COPIES
            // restores them.
            float* dp = nullptr;
            double* dl = nullptr;
            mapped = !getenv("LCU_GRAPH_COPIES")
                && cudaHostGetDevicePointer(reinterpret_cast<void**>(&dp), m->h_params, 0) == cudaSuccess
                && cudaHostGetDevicePointer(reinterpret_cast<void**>(&dl), m->h_lnew, 0) == cudaSuccess && dp && dl;
            if(!mapped)
                cudaGetLastError();
            const size_t po = (size_t)slot*m->npars;
            // the one-kernel path is used only where the host can see that it gave up (the result
            // word keeps its "pending" pattern): in graphs that write straight into mapped memory
            m->point_ok = mapped && !getenv("LCU_NO_POLL");
            if(mapped)
                ok = enqueue_batch(m, 1, dp + po, dl + slot, m->stream) == LCU_OK;
            else
            {
                ok = cudaMemcpyAsync(m->d_params + po, m->h_params + po, m->npars*sizeof(float), cudaMemcpyHostToDevice, m->stream) == cudaSuccess;
                ok = ok && enqueue_batch(m, 1, m->d_params + po, m->d_lnew + slot, m->stream) == LCU_OK;
                ok = ok && cudaMemcpyAsync(m->h_lnew + slot, m->d_lnew + slot, sizeof(double), cudaMemcpyDeviceToHost, m->stream) == cudaSuccess;
            }
            m->point_ok = false;
            ok = (cudaStreamEndCapture(m->stream, &graph) == cudaSuccess) && ok && graph;
        }
        // captured launches have not run: they count each time the graph is launched
        m->graph1_nodes = (unsigned)(g_launches.load() - launches0);
        g_launches.fetch_sub(m->graph1_nodes, std::memory_order_relaxed);
        if(ok)
            ok = cudaGraphInstantiate(slot == 0 ? &m->graph1 : &m->graph1b, graph, 0) == cudaSuccess;
        if(graph)
            cudaGraphDestroy(graph);
        if(!ok)
        {
            cudaGetLastError();
            drop_point_graphs(m);
            m->graph1_off = true;
            return false;
        }
    }
    m->graph1_rows[0] = m->row0;
    m->graph1_rows[1] = m->row1;
    m->graph1_split = split;
    m->graph1_conv = conv;
    m->graph1_mapped = mapped;
    return true;
}

void destroy_device_state(lcu_model* m)
{
    if(m->ctx && m->ctx->device >= 0)
        cudaSetDevice(m->ctx->device);
    drop_point_graphs(m);
    for(cudaEvent_t& e : m->async_done)
        if(e) { cudaEventDestroy(e); e = nullptr; }
    for(cudaEvent_t& e : m->ev_io)
        if(e) { cudaEventDestroy(e); e = nullptr; }
    for(lcu_model::EventSet& es : m->evsets)
        for(cudaEvent_t& e : es.e)
            cudaEventDestroy(e);
    m->evsets.clear();
    void* bufs[] = { m->d_image, m->d_weight, m->d_objs, m->d_raw, m->d_partial, m->d_counter, m->d_sync, m->d_params, m->d_lnew,
                     m->d_value1, m->d_error1, m->d_model1, m->d_chi1 };
    for(void* p : bufs)
        if(p) cudaFree(p);
    if(m->h_params) cudaFreeHost(m->h_params);
    if(m->h_lnew) cudaFreeHost(m->h_lnew);
    if(m->mod && drv.ModuleUnload) drv.ModuleUnload(m->mod);
    if(m->stream) cudaStreamDestroy(m->stream);
}

} // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int lcu_version(void) { return LCU_VERSION; }

const char* lcu_last_error(void) { return get_error(); }

unsigned long long lcu_launch_count(void) { return g_launches.load(); }

int lcu_create(int device, const char* kernel_dir, const char* objects_dir, lcu_ctx** out)
{
    if(!out)
    {
        set_error("lcu_create: null output");
        return LCU_E_ARG;
    }
    *out = nullptr;
    lcu_ctx* ctx = new lcu_ctx;
    const std::string base = library_dir();
    ctx->kernel_dir = kernel_dir ? kernel_dir : base + "/kernel";
    // the library ships no object files: the plugin directory is the caller's
    // (a Lensed installation's objects/), or $LENSED_PATH/objects as the
    // reference resolves it (src/kernel.c:11-13, src/path.c:56-78)
    if(objects_dir)
        ctx->objects_dir = objects_dir;
    else if(const char* lp = getenv("LENSED_PATH"))
        ctx->objects_dir = std::string(lp) + (*lp && lp[strlen(lp) - 1] == '/' ? "" : "/") + "objects";
    else
        ctx->objects_dir = base + "/objects";

    struct { const char* file; std::string* dst; } files[] = {
        { "shim.cuh", &ctx->shim }, { "object.cuh", &ctx->object_hdr }, { "lensed.cu", &ctx->kernels } };
    for(auto& f : files)
    {
        bool ok = false;
        *f.dst = read_text_file(ctx->kernel_dir + "/" + f.file, &ok);
        if(!ok)
        {
            // src/kernel.c:680-683
            set_error("could not load kernel \"%s\" (file not found: %s/%s)", f.file, ctx->kernel_dir.c_str(), f.file);
            delete ctx;
            return LCU_E_IO;
        }
    }

    ctx->device = device;
    if(device >= 0)
    {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if(e != cudaSuccess || device >= count)
        {
            set_error("CUDA device %d not available (%s, %d devices)", device,
                      e != cudaSuccess ? cudaGetErrorString(e) : "out of range", count);
            delete ctx;
            return LCU_E_NODEVICE;
        }
        if(cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess)
        {
            set_error("could not initialise CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
            delete ctx;
            return LCU_E_CUDA;
        }
        cudaDeviceProp prop;
        if(cudaGetDeviceProperties(&prop, device) == cudaSuccess)
        {
            ctx->sm_count = prop.multiProcessorCount;
            if(prop.major != 10)
            {
                set_error("device %d is compute capability %d.%d; this library is built for sm_100a only",
                          device, prop.major, prop.minor);
                delete ctx;
                return LCU_E_NODEVICE;
            }
        }
        if(!init_driver())
        {
            delete ctx;
            return LCU_E_CUDA;
        }
    }
    *out = ctx;
    return LCU_OK;
}

void lcu_destroy(lcu_ctx* ctx) { delete ctx; }

int lcu_object_pairable(lcu_ctx* ctx, const char* name, const char** why)
{
    if(!ctx || !name)
    {
        set_error("lcu_object_pairable: null argument");
        return -LCU_E_ARG;
    }
    const ObjectInfo* info = ctx->object(name);
    if(!info)
        return -LCU_E_COMPILE;
    if(why)
        *why = info->pair_log.c_str();
    return info->pairable ? 1 : 0;
}

int lcu_object_info(lcu_ctx* ctx, const char* name, int* type, size_t* words, size_t* npar,
                    lcu_param* params, size_t cap)
{
    if(!ctx || !name)
    {
        set_error("lcu_object_info: null argument");
        return LCU_E_ARG;
    }
    const ObjectInfo* info = ctx->object(name);
    if(!info)
    {
        const std::string msg = get_error();
        if(msg.find("could not load") != std::string::npos) return LCU_E_IO;
        if(msg.find("failed to build") != std::string::npos) return LCU_E_COMPILE;
        return LCU_E_OBJECT;
    }
    if(type) *type = info->type;
    if(words) *words = info->words;
    if(npar) *npar = info->params.size();
    if(params)
        for(size_t i = 0; i < info->params.size() && i < cap; ++i)
            params[i] = info->params[i];
    return LCU_OK;
}

int lcu_quad_rule_count(void) { return quad_rule_count(); }
const char* lcu_quad_rule_name(int i) { return quad_rule_name(i); }
const char* lcu_quad_rule_info(int i) { return quad_rule_info(i); }
int lcu_quad_rule(const char* rule, double sx, double sy, float* qq, float* ww) { return quad_rule(rule, sx, sy, qq, ww); }

int lcu_model_create(lcu_ctx* ctx, const lcu_object_spec* specs, size_t nobjs, const lcu_model_desc* desc, lcu_model** out)
{
    if(!ctx || !desc || !out || (nobjs && !specs))
    {
        set_error("lcu_model_create: null argument");
        return LCU_E_ARG;
    }
    *out = nullptr;
    if(desc->width == 0 || desc->height == 0 || desc->nq == 0 || !desc->qq || !desc->ww)
    {
        set_error("lcu_model_create: image size and quadrature rule are required");
        return LCU_E_ARG;
    }
    if(desc->width*desc->height > (size_t)1 << 30)
    {
        set_error("lcu_model_create: image too large");
        return LCU_E_ARG;
    }
    if(ctx->device >= 0 && (!desc->image || !desc->weight))
    {
        set_error("lcu_model_create: image and weight are required");
        return LCU_E_ARG;
    }
    if(desc->psf && (desc->psf_width == 0 || desc->psf_height == 0))
    {
        set_error("lcu_model_create: PSF given without size");
        return LCU_E_ARG;
    }

    lcu_model* m = new lcu_model;
    m->ctx = ctx;

    // object list, offsets: src/lensed.c:826-828 (data words), src/kernel.c:632-633 (params)
    size_t d = 0, p = 0;
    int nplanes = 0, typ = 0;
    for(size_t i = 0; i < nobjs; ++i)
    {
        if(!specs[i].name)
        {
            set_error("lcu_model_create: object %zu has no name", i);
            delete m;
            return LCU_E_ARG;
        }
        const ObjectInfo* info = ctx->object(specs[i].name);
        if(!info)
        {
            delete m;
            const std::string msg = get_error();
            if(msg.find("could not load") != std::string::npos) return LCU_E_IO;
            if(msg.find("failed to build") != std::string::npos) return LCU_E_COMPILE;
            return LCU_E_OBJECT;
        }
        // src/input/ini.c:249-260
        if(info->type != typ && info->type != LCU_FOREGROUND)
        {
            if(info->type == LCU_LENS && ++nplanes > 1)
            {
                set_error("multiple lensing planes are not supported");
                delete m;
                return LCU_E_ARG;
            }
            typ = info->type;
        }
        ModelObject mo;
        mo.info = info;
        mo.d = d;
        mo.p = p;
        mo.ipp.assign(info->params.size(), 0);
        for(size_t j = 0; j < info->params.size(); ++j)
        {
            mo.ipp[j] = specs[i].ipp ? (specs[i].ipp[j] != 0) : 0;
            if(!mo.ipp[j])
                continue;
            // src/lensed.c:196-234: only (X, Y) pairs can carry image plane priors
            const int pt = info->params[j].type;
            bool good = false;
            if(pt == LCU_POSITION_X)
                good = j + 1 < info->params.size() && info->params[j + 1].type == LCU_POSITION_Y
                       && specs[i].ipp[j + 1];
            else if(pt == LCU_POSITION_Y)
                good = j > 0 && info->params[j - 1].type == LCU_POSITION_X && specs[i].ipp[j - 1];
            if(!good)
            {
                set_error("object `%s`: image plane prior requires pair (X,Y) of parameters", specs[i].name);
                delete m;
                return LCU_E_ARG;
            }
        }
        m->objs.push_back(mo);
        // data blocks start on 16-byte boundaries (float4 members); every
        // object shipped with Lensed already has a multiple of 4 words
        d += (info->words + 3)/4*4;
        p += info->params.size();
    }
    m->words = std::max<size_t>(d, 4);
    m->npars = p;
    m->width = desc->width;
    m->height = desc->height;
    m->size = desc->width*desc->height;
    m->row0 = 0;
    m->row1 = m->height;
    memcpy(m->pcs, desc->pcs, sizeof(m->pcs));
    m->nq = desc->nq;
    m->has_psf = desc->psf != nullptr;
    m->psfw = m->has_psf ? desc->psf_width : 0;
    m->psfh = m->has_psf ? desc->psf_height : 0;
    m->flags = desc->flags;
    m->obj_const = !(desc->flags & LCU_OBJ_SHARED);
    // two rays per thread if every object's per-ray code can be typed as pairs
    {
        const char* fold_env = getenv("LCU_FOLD_SETTER");
        m->fold = fold_env && *fold_env == '1';
        const char* point_env = getenv("LCU_FUSED_POINT");
        m->point = point_env && *point_env == '1';
    }
    m->pair = !(desc->flags & LCU_NO_PAIR);
    for(const ModelObject& o : m->objs)
        m->pair = m->pair && o.info->pairable;
    {
        const char* env = getenv("LCU_PAIR");
        if(env && *env == '0')
            m->pair = false;
    }

    if(m->has_psf)
    {
        // convolution tile height: same arithmetic as LCU_CT_H in kernel/lensed.cu
        const size_t nv = (8 + m->psfw - 1 + 3)/4;
        const size_t cwmin = 56 + 4*nv;
        const size_t cw = cwmin + ((12 - cwmin%8)%8);
        m->conv_tile_h = cw*(32 + m->psfh - 1)*4 <= 48*1024 ? 32 : cw*(16 + m->psfh - 1)*4 <= 48*1024 ? 16 : 8;
        if(cw*(8 + m->psfh - 1)*4 > 48*1024)
        {
            set_error("lcu_model_create: PSF %zu x %zu too large for the convolution tile", m->psfw, m->psfh);
            delete m;
            return LCU_E_ARG;
        }
    }

    // fix coordinate system for the half-pixel offset of even PSFs, src/lensed.c:885-891
    if(m->has_psf)
    {
        if(m->psfw % 2 == 0) m->pcs[0] += 0.5f;
        if(m->psfh % 2 == 0) m->pcs[1] += 0.5f;
    }

    // points per launch: object blocks share the 64 KB constant bank with the
    // quadrature table and the PSF; rendered images of PSF models are staged
    // in HBM (budget LCU_RAW_BUDGET_MB, default 4096)
    {
        const size_t used = m->nq*16 + m->psfw*m->psfh*4 + 2048;
        if(used + m->words*4 > 65536)
        {
            // lcu_quad + lcu_psf + one object block have to fit the 64 KB constant bank
            set_error("lcu_model_create: quadrature rule (%zu points), PSF (%zu x %zu) and object block (%zu words) need %zu bytes "
                      "of constant memory, more than the 64 KB bank holds", m->nq, m->psfw, m->psfh, m->words, used + m->words*4);
            delete m;
            return LCU_E_ARG;
        }
        size_t maxb = (65536 - used)/(std::max<size_t>(m->words, 1)*4);
        maxb = std::min<size_t>(std::max<size_t>(maxb, 1), 1024);
        {
            // per point and launch: the staged pre-PSF image (PSF models) and one double per 32 pixels
            // of chi^2 partial sums; budget LCU_RAW_BUDGET_MB (default 4096) for the sum of the two
            const char* env = getenv("LCU_RAW_BUDGET_MB");
            const size_t budget = (env && *env ? (size_t)atoll(env) : 4096) << 20;
            const size_t per_point = (m->has_psf ? m->size*sizeof(float) : 0) + div_up(m->size, 32)*sizeof(double);
            maxb = std::min(maxb, std::max<size_t>(budget/std::max<size_t>(per_point, 1), 1));
        }
        if(desc->max_batch)
            maxb = std::min(maxb, desc->max_batch);
        m->maxb = maxb;
    }

    // Is the rule a Cartesian grid (first axis outer, second inner)?  All
    // Gauss-Kronrod and sub-sampling rules of src/quad/ are; gm75 is not.
    size_t quad_ni = 0, quad_nj = 0;
    for(size_t nj = 2; nj*nj <= m->nq*m->nq && nj <= m->nq/2; ++nj)
    {
        if(m->nq % nj)
            continue;
        bool grid = true;
        for(size_t n = 0; n < m->nq && grid; ++n)
            grid = desc->qq[2*n] == desc->qq[2*((n/nj)*nj)] && desc->qq[2*n + 1] == desc->qq[2*(n%nj) + 1];
        if(grid)
        {
            quad_nj = nj;
            quad_ni = m->nq/nj;
            break;
        }
    }

    // program text: main_program(), src/kernel.c:838-879 -- ABI headers,
    // each distinct object once, compute, set_params, kernels
    auto assemble = [&](int pair_minblocks)
    {
        std::ostringstream s;
        // kernel_options(), src/kernel.c:881-944
        s << "#define IMAGE_SIZE " << m->size << "\n"
          << "#define IMAGE_WIDTH " << m->width << "\n"
          << "#define IMAGE_HEIGHT " << m->height << "\n"
          << "#define PSF " << (m->has_psf ? 1 : 0) << "\n"
          << "#define PSF_WIDTH " << m->psfw << "\n"
          << "#define PSF_HEIGHT " << m->psfh << "\n"
          << "#define QUAD_POINTS " << m->nq << "\n"
          << "#define LCU_QUAD_NI " << quad_ni << "\n"
          << "#define LCU_QUAD_NJ " << quad_nj << "\n"
          << "#define LCU_WORDS " << m->words << "\n"
          << "#define LCU_NPARS " << std::max<size_t>(m->npars, 1) << "\n"
          << "#define LCU_MAXB " << m->maxb << "\n"
          << "#define LCU_OBJ_CONST " << (m->obj_const ? 1 : 0) << "\n"
          << "#define LCU_PAIR " << (m->pair ? 1 : 0) << "\n"
          << "#define LCU_FOLD " << (m->fold ? 1 : 0) << "\n"
          << "#define LCU_POINT " << (m->point ? 1 : 0) << "\n";
        if(pair_minblocks)
            s << "#define LCU_PAIR_MINBLOCKS " << pair_minblocks << "\n";
        s << "#include \"shim.cuh\"\n#include \"object.cuh\"\n\n";
        std::vector<const ObjectInfo*> uniq;
        for(const ModelObject& o : m->objs)
            if(std::find(uniq.begin(), uniq.end(), o.info) == uniq.end())
                uniq.push_back(o.info);
        for(const ObjectInfo* info : uniq)
            s << info->wrapped;
        s << "//----------------------------------------------------------------------------\n"
          << "// compute\n"
          << "//----------------------------------------------------------------------------\n"
          << generate_compute(m->objs)
          << (m->pair ? generate_compute(m->objs, true) : std::string())
          << "//----------------------------------------------------------------------------\n"
          << "// set_params\n"
          << "//----------------------------------------------------------------------------\n"
          << generate_set_params(m->objs)
          << "//----------------------------------------------------------------------------\n"
          << "// kernel/lensed.cu\n"
          << "//----------------------------------------------------------------------------\n"
          << "#line 1 \"kernel/lensed.cu\"\n"
          << ctx->kernels;
        return s.str();
    };

    m->source = assemble(0);
    if(desc->flags & LCU_SOURCE_ONLY)
    {
        if(ctx->device >= 0)
        {
            set_error("lcu_model_create: LCU_SOURCE_ONLY needs a context without a device");
            delete m;
            return LCU_E_ARG;
        }
        *out = m;
        return LCU_OK;
    }
    if(!ctx->compile(m->source, m->flags, &m->cubin, &m->log))
    {
        set_error("failed to build program\n%s", m->log.c_str());
        delete m;
        return LCU_E_COMPILE;
    }
    // The two-rays kernel is built for 3 resident blocks per SM (80 registers).
    // A model whose ray function needs many more (epl_plus_shear + 3 sources:
    // ~125) would spill inside the ray loop: build it for 2 blocks instead if
    // that removes the spills.
    if(m->pair)
    {
        const char* extra = getenv("LCU_NVRTC_FLAGS");
        unsigned regs = 0, stack = 0, regs2 = 0, stack2 = 0;
        if(!(extra && strstr(extra, "LCU_PAIR_MINBLOCKS")) && cubin_kernel_usage(m->cubin, "lcu_render_pair", &regs, &stack)
           && stack > 64)
        {
            const std::string src2 = assemble(2);
            std::vector<char> cubin2;
            std::string log2;
            if(ctx->compile(src2, m->flags, &cubin2, &log2) && cubin_kernel_usage(cubin2, "lcu_render_pair", &regs2, &stack2)
               && stack2 < stack)
            {
                m->source = src2;
                m->cubin.swap(cubin2);
                m->log = log2;
            }
        }
        // The other way round: a ray function that needs 65 ... 80 registers fits
        // 3 blocks per SM (24 warps); if ptxas can do it in 64 at the price of a few
        // spilled words outside the hot path, 4 blocks fit (32 warps).  Measured on
        // C5 (epl_plus_shear + 3 sersic: 76 -> 64 registers, stack 32 -> 48 bytes):
        // +2.9 % (profiles/r02_call2_summary.txt); building for 5 blocks gains nothing.
        else if(!(extra && strstr(extra, "LCU_PAIR_MINBLOCKS")) && regs > 64 && regs <= 80)
        {
            const std::string src4 = assemble(4);
            std::vector<char> cubin4;
            std::string log4;
            if(ctx->compile(src4, m->flags, &cubin4, &log4) && cubin_kernel_usage(cubin4, "lcu_render_pair", &regs2, &stack2)
               && regs2 <= 64 && stack2 <= stack + 32)
            {
                m->source = src4;
                m->cubin.swap(cubin4);
                m->log = log4;
            }
        }
    }

    if(ctx->device < 0)
    {
        *out = m;
        return LCU_OK;
    }

    // ---- device set-up ---------------------------------------------------
#define M_CHECK(expr) do { int rc_ = [&]() -> int { expr; return LCU_OK; }(); \
        if(rc_) { destroy_device_state(m); delete m; return rc_; } } while(0)

    M_CHECK(RT_CHECK(cudaSetDevice(ctx->device)));
    M_CHECK(RT_CHECK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking)));
    M_CHECK(DRV_CHECK(drv.ModuleLoadData(&m->mod, m->cubin.data())));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_set, m->mod, "lcu_set_params")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render[0], m->mod, "lcu_render_s1")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render[1], m->mod, "lcu_render_s2")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render[2], m->mod, "lcu_render_s4")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render[3], m->mod, "lcu_render_s8")));
    if(m->fold)
    {
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_fold[1], m->mod, "lcu_render_fold_s2")));
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_fold[2], m->mod, "lcu_render_fold_s4")));
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_fold[3], m->mod, "lcu_render_fold_s8")));
    }
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_err[0], m->mod, "lcu_render_err_s1")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_err[1], m->mod, "lcu_render_err_s2")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_err[2], m->mod, "lcu_render_err_s4")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_err[3], m->mod, "lcu_render_err_s8")));
    if(m->pair)
    {
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_pair, m->mod, "lcu_render_pair")));
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_q[1], m->mod, "lcu_render_q_s2")));
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_q[2], m->mod, "lcu_render_q_s4")));
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_q[3], m->mod, "lcu_render_q_s8")));
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_render_pair_err, m->mod, "lcu_render_pair_err")));
    }
    if(m->point)
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_point[2], m->mod, "lcu_point_s4")));
    if(m->point)
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_point[3], m->mod, "lcu_point_s8")));
    for(int i = 2; i < 4 && m->point; ++i)
    {
        int per_sm = 0;
        M_CHECK(DRV_CHECK(drv.OccupancyMaxActiveBlocks(&per_sm, m->f_point[i], 256, 0)));
        m->point_capacity[i] = (size_t)std::max(per_sm, 0)*(size_t)std::max(ctx->sm_count, 0);
    }
    M_CHECK(RT_CHECK(cudaMalloc(&m->d_sync, 2*sizeof(unsigned))));
    M_CHECK(RT_CHECK(cudaMemset(m->d_sync, 0, 2*sizeof(unsigned))));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_reduce, m->mod, "lcu_reduce")));
    M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_make_weight, m->mod, "lcu_make_weight")));
    if(m->has_psf)
    {
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_conv, m->mod, "lcu_convolve")));
        M_CHECK(DRV_CHECK(drv.ModuleGetFunction(&m->f_conv_small, m->mod, "lcu_convolve_small")));
    }

    // constant tables: quadrature rule (src/lensed.c:817-823) and PSF (:808)
    {
        std::vector<float> quad(4*m->nq);
        for(size_t n = 0; n < m->nq; ++n)
        {
            quad[4*n + 0] = desc->qq[2*n + 0];
            quad[4*n + 1] = desc->qq[2*n + 1];
            quad[4*n + 2] = desc->ww[2*n + 0];
            quad[4*n + 3] = desc->ww[2*n + 1];
        }
        CUdeviceptr ptr = 0;
        size_t bytes = 0;
        M_CHECK(DRV_CHECK(drv.ModuleGetGlobal(&ptr, &bytes, m->mod, "lcu_quad")));
        M_CHECK(RT_CHECK(cudaMemcpy(reinterpret_cast<void*>(ptr), quad.data(), quad.size()*sizeof(float), cudaMemcpyHostToDevice)));
        if(m->has_psf)
        {
            M_CHECK(DRV_CHECK(drv.ModuleGetGlobal(&ptr, &bytes, m->mod, "lcu_psf")));
            M_CHECK(RT_CHECK(cudaMemcpy(reinterpret_cast<void*>(ptr), desc->psf, m->psfw*m->psfh*sizeof(float), cudaMemcpyHostToDevice)));
        }
        if(m->obj_const)
        {
            M_CHECK(DRV_CHECK(drv.ModuleGetGlobal(&m->c_objs, &bytes, m->mod, "lcu_objs_c")));
        }
    }

    // data buffers, src/lensed.c:801-812
    M_CHECK(RT_CHECK(cudaMalloc(&m->d_image, m->size*sizeof(float))));
    M_CHECK(RT_CHECK(cudaMalloc(&m->d_weight, m->size*sizeof(float))));
    M_CHECK(RT_CHECK(cudaMemcpy(m->d_image, desc->image, m->size*sizeof(float), cudaMemcpyHostToDevice)));
    M_CHECK(RT_CHECK(cudaMemcpy(m->d_weight, desc->weight, m->size*sizeof(float), cudaMemcpyHostToDevice)));
    M_CHECK(RT_CHECK(cudaMalloc(&m->d_objs, m->maxb*m->words*sizeof(uint32_t))));
    M_CHECK(RT_CHECK(cudaMalloc(&m->d_counter, m->maxb*sizeof(unsigned))));
    M_CHECK(RT_CHECK(cudaMemset(m->d_counter, 0, m->maxb*sizeof(unsigned))));
    for(cudaEvent_t& e : m->ev_io)
        M_CHECK(RT_CHECK(cudaEventCreate(&e)));
    M_CHECK(return ensure_partial(m, 1));
    M_CHECK(return ensure_stage(m, 64));
#undef M_CHECK

    *out = m;
    return LCU_OK;
}

void lcu_model_destroy(lcu_model* m)
{
    if(!m)
        return;
    destroy_device_state(m);
    delete m;
}

size_t lcu_model_npars(const lcu_model* m) { return m ? m->npars : 0; }
size_t lcu_model_words(const lcu_model* m) { return m ? m->words : 0; }
size_t lcu_model_max_batch(const lcu_model* m) { return m ? m->maxb : 0; }
int lcu_model_rays_per_thread(const lcu_model* m) { return m ? (m->pair ? 2 : 1) : 0; }
const char* lcu_model_source(const lcu_model* m) { return m ? m->source.c_str() : nullptr; }
const char* lcu_model_build_log(const lcu_model* m) { return m ? m->log.c_str() : nullptr; }

size_t lcu_model_cubin(const lcu_model* m, const void** image)
{
    if(!m)
        return 0;
    if(image)
        *image = m->cubin.data();
    return m->cubin.size();
}

int lcu_model_kernel_usage(const lcu_model* m, const char* kernel, unsigned* registers, unsigned* stack_bytes)
{
    unsigned regs = 0, stack = 0;
    if(!m || !kernel || !cubin_kernel_usage(m->cubin, kernel, &regs, &stack))
    {
        set_error("lcu_model_kernel_usage: no kernel \"%s\" in the module", kernel ? kernel : "(null)");
        return LCU_E_ARG;
    }
    if(registers) *registers = regs;
    if(stack_bytes) *stack_bytes = stack;
    return LCU_OK;
}

int lcu_model_set_rows(lcu_model* m, size_t row0, size_t row1)
{
    if(!m || row0 >= row1 || row1 > m->height)
    {
        set_error("lcu_model_set_rows: invalid row range");
        return LCU_E_ARG;
    }
    if(m->async_busy[0] || m->async_busy[1])
    {
        set_error("lcu_model_set_rows: evaluations started with lcu_loglike_async are still in flight (lcu_loglike_wait them first)");
        return LCU_E_ARG;
    }
    m->row0 = row0;
    m->row1 = row1;
    return LCU_OK;
}

// `idle`: the call uses or changes state that evaluations started with lcu_loglike_async
// are still reading (staging slots, scratch buffers, data, row range): refuse it until they
// have been waited for
static int need_device(const lcu_model* m, const char* fn, bool idle = true)
{
    if(!m)
    {
        set_error("%s: null model", fn);
        return LCU_E_ARG;
    }
    if(m->ctx->device < 0)
    {
        set_error("%s: compile-only context has no device (there is no CPU fallback)", fn);
        return LCU_E_NODEVICE;
    }
    if(idle && (m->async_busy[0] || m->async_busy[1]))
    {
        set_error("%s: evaluations started with lcu_loglike_async are still in flight (lcu_loglike_wait them first)", fn);
        return LCU_E_ARG;
    }
    return LCU_OK;
}

int lcu_model_set_data(lcu_model* m, const float* image, const float* weight)
{
    int rc = need_device(m, "lcu_model_set_data");
    if(rc) return rc;
    RT_CHECK(cudaSetDevice(m->ctx->device));
    RT_CHECK(cudaStreamSynchronize(m->stream));
    if(image)
        RT_CHECK(cudaMemcpy(m->d_image, image, m->size*sizeof(float), cudaMemcpyHostToDevice));
    if(weight)
        RT_CHECK(cudaMemcpy(m->d_weight, weight, m->size*sizeof(float), cudaMemcpyHostToDevice));
    return LCU_OK;
}

int lcu_model_make_weight(lcu_model* m, const float* gain_map, float gain, double offset, const int* mask)
{
    int rc = need_device(m, "lcu_model_make_weight");
    if(rc) return rc;
    RT_CHECK(cudaSetDevice(m->ctx->device));
    float* d_gain = nullptr;
    int* d_mask = nullptr;
    auto cleanup = [&]() { if(d_gain) cudaFree(d_gain); if(d_mask) cudaFree(d_mask); };
    if(gain_map)
    {
        RT_CHECK(cudaMalloc(&d_gain, m->size*sizeof(float)));
        if(cudaMemcpyAsync(d_gain, gain_map, m->size*sizeof(float), cudaMemcpyHostToDevice, m->stream) != cudaSuccess)
        {
            cleanup();
            set_error("lcu_model_make_weight: copy of the gain map failed");
            return LCU_E_CUDA;
        }
    }
    if(mask)
    {
        if(cudaMalloc(&d_mask, m->size*sizeof(int)) != cudaSuccess
           || cudaMemcpyAsync(d_mask, mask, m->size*sizeof(int), cudaMemcpyHostToDevice, m->stream) != cudaSuccess)
        {
            cleanup();
            set_error("lcu_model_make_weight: copy of the mask failed");
            return LCU_E_CUDA;
        }
    }
    long long n = (long long)m->size;
    const unsigned blocks = (unsigned)std::min<size_t>(div_up(m->size, 256), (size_t)std::max(m->ctx->sm_count, 1)*8);
    void* args[] = { &n, &m->d_image, &d_gain, &gain, &offset, &d_mask, &m->d_weight };
    rc = launch(m, m->f_make_weight, dim3(blocks), dim3(256), args, m->stream);
    const cudaError_t e = cudaStreamSynchronize(m->stream);
    cleanup();
    if(rc) return rc;
    if(e != cudaSuccess)
    {
        set_error("lcu_model_make_weight: %s", cudaGetErrorString(e));
        return LCU_E_CUDA;
    }
    return LCU_OK;
}

int lcu_model_get_weight(lcu_model* m, float* weight)
{
    int rc = need_device(m, "lcu_model_get_weight");
    if(rc) return rc;
    if(!weight)
    {
        set_error("lcu_model_get_weight: null output");
        return LCU_E_ARG;
    }
    RT_CHECK(cudaSetDevice(m->ctx->device));
    RT_CHECK(cudaStreamSynchronize(m->stream));
    RT_CHECK(cudaMemcpy(weight, m->d_weight, m->size*sizeof(float), cudaMemcpyDeviceToHost));
    return LCU_OK;
}

int lcu_loglike_batch_device(lcu_model* m, size_t nbatch, const float* d_params, double* d_lnew, void* stream)
{
    int rc = need_device(m, "lcu_loglike_batch_device");
    if(rc) return rc;
    if(nbatch == 0)
        return LCU_OK;
    if(!d_params || !d_lnew)
    {
        set_error("lcu_loglike_batch_device: null argument");
        return LCU_E_ARG;
    }
    RT_CHECK(cudaSetDevice(m->ctx->device));
    return enqueue_batch(m, nbatch, d_params, d_lnew, static_cast<cudaStream_t>(stream));
}

int lcu_loglike_batch(lcu_model* m, size_t nbatch, const float* params, double* lnew)
{
    int rc = need_device(m, "lcu_loglike_batch");
    if(rc) return rc;
    if(nbatch == 0)
        return LCU_OK;
    if(!params || !lnew)
    {
        set_error("lcu_loglike_batch: null argument");
        return LCU_E_ARG;
    }
    RT_CHECK(cudaSetDevice(m->ctx->device));
    rc = ensure_stage(m, nbatch);
    if(rc) return rc;
    cudaEvent_t* ev = m->profile ? m->ev_io : nullptr;

    // the sampler's one-point call: one graph launch instead of 7 API calls
    if(nbatch == 1 && !m->profile && single_point_graph(m))
    {
        memcpy(m->h_params, params, m->npars*sizeof(float));
        // With mapped staging the last kernel stores the result in pinned host
        // memory: watching that word is a few microseconds quicker than waiting for
        // the runtime to report the stream idle.  The sentinel is a NaN pattern no
        // arithmetic produces; the stream is queried now and then so that a failed
        // launch ends the wait.
        static const unsigned long long pending = 0x7ff8dead5eed0001ull;
        volatile unsigned long long* word = reinterpret_cast<volatile unsigned long long*>(m->h_lnew);
        const bool watch = m->graph1_mapped && !getenv("LCU_NO_POLL");
        if(watch)
            *word = pending;
        RT_CHECK(cudaGraphLaunch(m->graph1, m->stream));
        g_launches.fetch_add(m->graph1_nodes, std::memory_order_relaxed);
        if(watch)
        {
            for(unsigned spins = 1; *word == pending; ++spins)
                if((spins & 0xfff) == 0 && cudaStreamQuery(m->stream) != cudaErrorNotReady)
                    break;
            if(*word == pending)
                RT_CHECK(cudaStreamSynchronize(m->stream));
            if(*word == pending)
            {
                // the one-kernel path gave up waiting (its grid was not resident at once): never again on this model
                rc = point_kernel_failed(m);
                if(rc) return rc;
                return lcu_loglike_batch(m, 1, params, lnew);
            }
        }
        else
            RT_CHECK(cudaStreamSynchronize(m->stream));
        *lnew = m->h_lnew[0];
        return LCU_OK;
    }

    // parameter upload, src/nested.c:67-74
    memcpy(m->h_params, params, nbatch*m->npars*sizeof(float));
    if(ev) cudaEventRecord(ev[0], m->stream);
    RT_CHECK(cudaMemcpyAsync(m->d_params, m->h_params, nbatch*m->npars*sizeof(float), cudaMemcpyHostToDevice, m->stream));
    if(ev) cudaEventRecord(ev[1], m->stream);
    rc = enqueue_batch(m, nbatch, m->d_params, m->d_lnew, m->stream);
    if(rc) return rc;
    // result read-back, src/nested.c:102-115 (8 bytes per point instead of the chi^2 map)
    if(ev) cudaEventRecord(ev[2], m->stream);
    RT_CHECK(cudaMemcpyAsync(m->h_lnew, m->d_lnew, nbatch*sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    if(ev) cudaEventRecord(ev[3], m->stream);
    RT_CHECK(cudaStreamSynchronize(m->stream));
    memcpy(lnew, m->h_lnew, nbatch*sizeof(double));

    if(ev)
    {
        float t;
        if(cudaEventElapsedTime(&t, ev[0], ev[1]) == cudaSuccess) m->prof.upload_ms += t;
        if(cudaEventElapsedTime(&t, ev[2], ev[3]) == cudaSuccess) m->prof.download_ms += t;
        harvest_profile(m);
    }
    return LCU_OK;
}

// the NaN pattern that marks a result word as pending (no arithmetic produces it)
static const unsigned long long LCU_PENDING = 0x7ff8dead5eed0001ull;

int lcu_loglike_async(lcu_model* m, const float* params, int* ticket)
{
    int rc = need_device(m, "lcu_loglike_async", false);
    if(rc) return rc;
    if(!params || !ticket)
    {
        set_error("lcu_loglike_async: null argument");
        return LCU_E_ARG;
    }
    const int slot = m->async_next;
    if(m->async_busy[slot])
    {
        set_error("lcu_loglike_async: two evaluations are in flight already (lcu_loglike_wait one of them first)");
        return LCU_E_ARG;
    }
    RT_CHECK(cudaSetDevice(m->ctx->device));
    rc = ensure_stage(m, 2);
    if(rc) return rc;
    if(!m->async_done[slot])
        RT_CHECK(cudaEventCreateWithFlags(&m->async_done[slot], cudaEventDisableTiming));
    const size_t po = (size_t)slot*m->npars;
    memcpy(m->h_params + po, params, m->npars*sizeof(float));
    volatile unsigned long long* word = reinterpret_cast<volatile unsigned long long*>(m->h_lnew + slot);
    *word = LCU_PENDING;
    if(!m->profile && single_point_graph(m))
    {
        RT_CHECK(cudaGraphLaunch(slot == 0 ? m->graph1 : m->graph1b, m->stream));
        g_launches.fetch_add(m->graph1_nodes, std::memory_order_relaxed);
    }
    else
    {
        RT_CHECK(cudaMemcpyAsync(m->d_params + po, m->h_params + po, m->npars*sizeof(float), cudaMemcpyHostToDevice, m->stream));
        rc = enqueue_batch(m, 1, m->d_params + po, m->d_lnew + slot, m->stream);
        if(rc) return rc;
        RT_CHECK(cudaMemcpyAsync(m->h_lnew + slot, m->d_lnew + slot, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    }
    RT_CHECK(cudaEventRecord(m->async_done[slot], m->stream));
    m->async_busy[slot] = true;
    m->async_next = 1 - slot;
    *ticket = slot;
    return LCU_OK;
}

int lcu_loglike_wait(lcu_model* m, int ticket, double* lnew)
{
    int rc = need_device(m, "lcu_loglike_wait", false);
    if(rc) return rc;
    if(ticket < 0 || ticket > 1 || !lnew || !m->async_busy[ticket])
    {
        set_error("lcu_loglike_wait: no evaluation in flight under ticket %d", ticket);
        return LCU_E_ARG;
    }
    if(m->async_redone[ticket])
    {
        // answered already while the other ticket was being redone (see below)
        m->async_redone[ticket] = false;
        m->async_busy[ticket] = false;
        *lnew = m->async_redo[ticket];
        return LCU_OK;
    }
    volatile unsigned long long* word = reinterpret_cast<volatile unsigned long long*>(m->h_lnew + ticket);
    if(!getenv("LCU_NO_POLL"))
        for(unsigned spins = 1; *word == LCU_PENDING; ++spins)
            if((spins & 0xfff) == 0 && cudaEventQuery(m->async_done[ticket]) != cudaErrorNotReady)
                break;
    m->async_busy[ticket] = false;
    if(*word == LCU_PENDING)
        RT_CHECK(cudaEventSynchronize(m->async_done[ticket]));
    if(*word == LCU_PENDING)
    {
        // The one-kernel path gave up waiting (its grid was not resident at once).  Redo every
        // outstanding point the three-kernel way now; the other ticket's answer is kept for its wait.
        RT_CHECK(cudaStreamSynchronize(m->stream));
        const int other = 1 - ticket;
        const bool other_busy = m->async_busy[other];
        std::vector<float> p[2];
        double done[2] = { m->h_lnew[0], m->h_lnew[1] };
        bool pending[2];
        for(int s = 0; s < 2; ++s)
        {
            p[s].assign(m->h_params + (size_t)s*m->npars, m->h_params + (size_t)(s + 1)*m->npars);
            pending[s] = reinterpret_cast<volatile unsigned long long*>(m->h_lnew)[s] == LCU_PENDING;
        }
        rc = point_kernel_failed(m);
        if(rc) return rc;
        m->async_busy[other] = false;
        rc = lcu_loglike_batch(m, 1, p[ticket].data(), lnew);
        if(rc) return rc;
        if(other_busy)
        {
            if(pending[other])
            {
                rc = lcu_loglike_batch(m, 1, p[other].data(), &done[other]);
                if(rc) return rc;
            }
            m->async_redo[other] = done[other];
            m->async_redone[other] = true;
            m->async_busy[other] = true;
        }
        return LCU_OK;
    }
    *lnew = m->h_lnew[ticket];
    if(m->profile)
        harvest_profile(m);
    return LCU_OK;
}

int lcu_loglike(lcu_model* m, const float* params, double* lnew)
{
    return lcu_loglike_batch(m, 1, params, lnew);
}

int lcu_render(lcu_model* m, const float* params, float* model_img, float* raw, float* err, float* chi)
{
    int rc = need_device(m, "lcu_render");
    if(rc) return rc;
    if(!params)
    {
        set_error("lcu_render: null parameters");
        return LCU_E_ARG;
    }
    RT_CHECK(cudaSetDevice(m->ctx->device));
    rc = ensure_stage(m, 1);
    if(rc) return rc;
    rc = ensure_partial(m, 1);
    if(rc) return rc;
    const size_t bytes = m->size*sizeof(float);
    float** bufs[] = { &m->d_value1, &m->d_error1, &m->d_model1, &m->d_chi1 };
    for(float** b : bufs)
        if(!*b)
        {
            RT_CHECK(cudaMalloc(b, bytes));
            RT_CHECK(cudaMemsetAsync(*b, 0, bytes, m->stream));
        }
    memcpy(m->h_params, params, m->npars*sizeof(float));
    RT_CHECK(cudaMemcpyAsync(m->d_params, m->h_params, m->npars*sizeof(float), cudaMemcpyHostToDevice, m->stream));
    rc = enqueue_points(m, 1, m->d_params, m->stream, m->d_value1, m->d_error1,
                        m->has_psf ? m->d_model1 : nullptr, m->d_chi1, false, nullptr);
    if(rc) return rc;
    RT_CHECK(cudaStreamSynchronize(m->stream));
    if(raw) RT_CHECK(cudaMemcpy(raw, m->d_value1, bytes, cudaMemcpyDeviceToHost));
    if(err) RT_CHECK(cudaMemcpy(err, m->d_error1, bytes, cudaMemcpyDeviceToHost));
    if(model_img) RT_CHECK(cudaMemcpy(model_img, m->has_psf ? m->d_model1 : m->d_value1, bytes, cudaMemcpyDeviceToHost));
    if(chi) RT_CHECK(cudaMemcpy(chi, m->d_chi1, bytes, cudaMemcpyDeviceToHost));
    return LCU_OK;
}

int lcu_set_params(lcu_model* m, const float* params, uint32_t* block)
{
    int rc = need_device(m, "lcu_set_params");
    if(rc) return rc;
    if(!params || !block)
    {
        set_error("lcu_set_params: null argument");
        return LCU_E_ARG;
    }
    RT_CHECK(cudaSetDevice(m->ctx->device));
    rc = ensure_stage(m, 1);
    if(rc) return rc;
    memcpy(m->h_params, params, m->npars*sizeof(float));
    RT_CHECK(cudaMemcpyAsync(m->d_params, m->h_params, m->npars*sizeof(float), cudaMemcpyHostToDevice, m->stream));
    int B = 1;
    void* args[] = { &B, &m->d_params, &m->d_objs };
    rc = launch(m, m->f_set, dim3(1), dim3(64), args, m->stream);
    if(rc) return rc;
    RT_CHECK(cudaStreamSynchronize(m->stream));
    RT_CHECK(cudaMemcpy(block, m->d_objs, m->words*sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return LCU_OK;
}

int lcu_profile_enable(lcu_model* m, int on)
{
    if(!m) return LCU_E_ARG;
    if(m->profile)
        harvest_profile(m);
    m->profile = on != 0;
    if(on)
        m->prof = lcu_profile{};
    return LCU_OK;
}

int lcu_profile_get(lcu_model* m, lcu_profile* out)
{
    if(!m || !out) return LCU_E_ARG;
    if(m->ctx->device >= 0)
    {
        cudaSetDevice(m->ctx->device);
        harvest_profile(m);         // waits for the chunks recorded so far
    }
    *out = m->prof;
    return LCU_OK;
}

int lcu_measure_fp32_peak(lcu_ctx* ctx, double* tflops)
{
    if(!ctx || !tflops)
    {
        set_error("lcu_measure_fp32_peak: null argument");
        return LCU_E_ARG;
    }
    if(ctx->device < 0)
    {
        set_error("lcu_measure_fp32_peak: compile-only context has no device");
        return LCU_E_NODEVICE;
    }
    RT_CHECK(cudaSetDevice(ctx->device));
    if(lcu_bench_ffma(ctx->sm_count, tflops) != 0)
    {
        set_error("FFMA micro-benchmark failed: %s", cudaGetErrorString(cudaGetLastError()));
        return LCU_E_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return LCU_OK;
}

} // extern "C"
