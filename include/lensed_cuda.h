/*
 * lensed_cuda.h -- C ABI of the B200-native replacement for Lensed's
 * per-likelihood model-image path (set_params -> render -> convolve ->
 * loglike -> reduce).
 *
 * The reference has no clean device interface: src/opencl.c exports three
 * functions (src/opencl.h:39-45), src/kernel.c four (src/kernel.h:3-16) and
 * everything else is raw OpenCL calls in the callers (src/lensed.c:644-1112,
 * src/nested.c:63-115 and :178-214, src/input/objects.c:72-265).  This header
 * is the boundary those call sites bind to instead; each entry point cites the
 * reference code it replaces.  Plain C: pointers and sizes only.
 *
 * Every function returning int returns 0 on success and a non-zero LCU_E_*
 * code on failure; lcu_last_error() then holds a message (including the NVRTC
 * build log for compile failures).  Nothing here ever calls exit() (the
 * reference's error() does, src/log.c:84-99).  There is no CPU fallback: a
 * context created on a machine without a CUDA device can assemble and
 * compile programs and answer metadata queries, but every compute entry
 * point fails with LCU_E_NODEVICE.
 *
 * Threading: calls on one lcu_model are not re-entrant (one stream per model,
 * as the reference's single in-order queue, src/lensed.c:702); distinct
 * models and contexts are independent.
 */
#ifndef LENSED_CUDA_H
#define LENSED_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCU_VERSION 100

enum
{
    LCU_OK = 0,
    LCU_E_ARG = 1,        /* invalid argument */
    LCU_E_IO = 2,         /* kernel or object file not found */
    LCU_E_COMPILE = 3,    /* NVRTC failure; log in lcu_last_error() */
    LCU_E_CUDA = 4,       /* CUDA runtime / driver error */
    LCU_E_NODEVICE = 5,   /* compute call on a compile-only context */
    LCU_E_OBJECT = 6      /* malformed object (type, params, functions) */
};

/* object types, kernel/object.cl:2-7 */
enum { LCU_LENS = 'L', LCU_SOURCE = 'S', LCU_FOREGROUND = 'F' };

/* parameter types, kernel/object.cl:10-19 */
enum
{
    LCU_PARAMETER = 0, LCU_POSITION_X, LCU_POSITION_Y, LCU_RADIUS,
    LCU_MAGNITUDE, LCU_AXIS_RATIO, LCU_POS_ANGLE
};

/* one entry of an object's parameter list: the device-side `struct param`
 * of kernel/object.cl:27-33, byte for byte */
typedef struct
{
    char  name[16];
    int   type;
    float bounds[2];      /* {0, 0} = unbounded */
    float defval;         /* > 0 or -0.0f = has a default (src/input/objects.c:225) */
} lcu_param;

typedef struct lcu_ctx lcu_ctx;
typedef struct lcu_model lcu_model;

int lcu_version(void);

/* message of the last failure on the calling thread */
const char* lcu_last_error(void);

/* number of kernels this library has launched in the calling process */
unsigned long long lcu_launch_count(void);

/*
 * Replaces get_lensed_cl() / free_lensed_cl() (src/opencl.c:132-241,
 * src/opencl.h:42-45).  device >= 0 selects a CUDA device; device < 0
 * creates a compile-only context (no GPU needed).  kernel_dir / objects_dir:
 * where kernel/{shim.cuh,object.cuh,lensed.cu} and objects/<name>.cl live
 * (the reference's LENSED_PATH/kernel and /objects, src/kernel.c:11-13).
 * kernel_dir NULL = the kernel/ directory shipped next to this library.
 * objects_dir is the objects/ directory of a Lensed installation -- its files
 * are consumed unmodified, and this library ships none of its own; NULL =
 * $LENSED_PATH/objects (as src/path.c:56-78 resolves it), else objects/ next
 * to this library.  A missing object file is reported by the call that first
 * needs it, in the reference's words (src/kernel.c:757-759).
 */
int lcu_create(int device, const char* kernel_dir, const char* objects_dir, lcu_ctx** ctx);
void lcu_destroy(lcu_ctx* ctx);

/*
 * Object metadata.  Replaces the device round trip of add_object()
 * (src/input/objects.c:72-239: build object_program(), run meta_<name> and
 * params_<name>, src/kernel.c:41-62).  The object file is compiled with NVRTC
 * and type, sizeof(data) and the parameter list are read from the compiled
 * module.  words = sizeof(data) in 4-byte words, rounded up
 * (src/input/objects.c:139).  params may be NULL; at most cap entries are
 * written.
 */
int lcu_object_info(lcu_ctx* ctx, const char* name, int* type, size_t* words,
                    size_t* npar, lcu_param* params, size_t cap);
/*
 * Can the object's per-ray function (deflection / brightness / foreground) be
 * compiled for two rays per thread (packed FP32 arithmetic)?  Returns 1 or 0,
 * or a negative error code if the object cannot be loaded; with 0, *why (if
 * not NULL) points at the compiler's reason, valid until the context is
 * destroyed.  No counterpart in the reference: a property of this back end.
 */
int lcu_object_pairable(lcu_ctx* ctx, const char* name, const char** why);

/*
 * Quadrature rules, replaces QUAD_RULES[] and quad_rule()
 * (src/quadrature.c:21-43, src/quad/ tables): point, sub2, sub4, gm75, g3k7,
 * g5k11, g7k15.  qq[n][2] = abscissae scaled by the pixel scale (sx, sy);
 * ww[n][2] = (weight, error weight).  lcu_quad_rule returns the number of
 * points or -1 for an unknown rule; call with qq = ww = NULL for the size.
 */
int lcu_quad_rule_count(void);
const char* lcu_quad_rule_name(int index);
const char* lcu_quad_rule_info(int index);
int lcu_quad_rule(const char* rule, double sx, double sy, float* qq, float* ww);

/* one object on the line of sight, in ini order (src/input.h:110-127) */
typedef struct
{
    const char* name;     /* object file name without .cl */
    const int*  ipp;      /* per-parameter image-plane-prior flags, or NULL
                             (src/input.h:92-93, "image" keyword) */
} lcu_object_spec;

/* model flags */
#define LCU_FAST_MATH   1u   /* allow FMA contraction in object code (reference
                                builds with -cl-fast-relaxed-math, src/lensed.c:744-748);
                                default is one IEEE operation per source operation */
#define LCU_OBJ_SHARED  2u   /* keep object blocks in shared memory instead of
                                the constant bank */
#define LCU_FAST_INTRINSICS 4u  /* exp/log/pow/sin/cos -> hardware exp2/log2/sin/cos
                                   approximations in SOURCE and FOREGROUND objects */
#define LCU_NO_PAIR     8u   /* render one ray per thread even if every object can be
                                compiled for two (packed FADD2/FMUL2/FFMA2 arithmetic);
                                both ways give the same bits, this one is slower */
#define LCU_FAST_LENS_INTRINSICS 16u /* the same in LENS objects (less accurate deflections) */
#define LCU_FAST_ATANH      32u /* atanh in LENS objects = (ln(1+x) - ln(1-x))/2 on the hardware
                                   log2: absolute error 2e-7 (isothermal ellipsoid deflections) */

#define LCU_SOURCE_ONLY     64u /* assemble the program text and stop: the model answers lcu_model_source /
                                   npars / words / rays_per_thread only (the reference writes the same text
                                   before it builds, src/lensed.c:714-735).  Contexts without a device only */

typedef struct
{
    size_t width, height;       /* image size */
    float  pcs[4];              /* pixel coordinate system (rx, ry, sx, sy), src/lensed.c:879-883 */
    size_t nq;                  /* quadrature points */
    const float* qq;            /* [nq][2] */
    const float* ww;            /* [nq][2] */
    const float* image;         /* [height][width] observed image */
    const float* weight;        /* [height][width] inverse variance, 0 = masked (src/lensed.c:479-482) */
    const float* psf;           /* [psf_height][psf_width] normalised PSF, or NULL */
    size_t psf_width, psf_height;
    size_t max_batch;           /* parameter points per launch; 0 = automatic */
    unsigned flags;
} lcu_model_desc;

/*
 * Replaces the device set-up of src/lensed.c:644-1112: main_program()
 * assembly (src/kernel.c:838-879) with generated compute() / set_params(),
 * the program build with the IMAGE_, PSF_ and QUAD_POINTS options
 * (src/kernel.c:881-944), all buffers and kernel arguments, and the
 * half-pixel shift for even PSF sizes (src/lensed.c:885-891).  All input
 * arrays are copied; the caller keeps ownership.
 */
int lcu_model_create(lcu_ctx* ctx, const lcu_object_spec* objs, size_t nobjs,
                     const lcu_model_desc* desc, lcu_model** model);
void lcu_model_destroy(lcu_model* model);

size_t lcu_model_npars(const lcu_model* model);    /* total parameters, object order */
size_t lcu_model_words(const lcu_model* model);    /* object block size in 4-byte words */
size_t lcu_model_max_batch(const lcu_model* model);
/* rays per thread of the render kernel for large images: 2 if every object is
   pairable and LCU_NO_PAIR is not set, else 1 */
int lcu_model_rays_per_thread(const lcu_model* model);
/* assembled program text (what `output = true` dumps as <root>kernel.cl, src/lensed.c:714-735) */
const char* lcu_model_source(const lcu_model* model);
const char* lcu_model_build_log(const lcu_model* model);
/* compiled sm_100a module image (for cuobjdump / caching); returns its size */
size_t lcu_model_cubin(const lcu_model* model, const void** image);
/* registers per thread and stack (spill) bytes of one kernel of the module, e.g.
   "lcu_render_pair", read from the module image; LCU_E_ARG if there is no such kernel */
int lcu_model_kernel_usage(const lcu_model* model, const char* kernel, unsigned* registers, unsigned* stack_bytes);

/*
 * Restrict the model to the image rows [row0, row1) (multi-GPU row strips):
 * log-likelihoods then are the strip's share -chi2_strip/2, which add up over
 * disjoint strips.  With a PSF the strip renders its own halo rows.
 */
int lcu_model_set_rows(lcu_model* model, size_t row0, size_t row1);

/*
 * Data preparation on the device (once per run, off the likelihood path).
 * lcu_model_set_data replaces the observed image and / or the weight map of an
 * existing model (either pointer may be NULL = keep): same kernels, no
 * recompilation, e.g. for a series of mock observations.  lcu_model_make_weight
 * builds the weight map from the image held by the model as make_weight() does
 * (src/data.c:314-330): weight = gain / (image + offset), evaluated in double and
 * narrowed to float, with gain a per-pixel map or, if gain_map is NULL, one value
 * (make_real(), src/data.c:286-311); pixels with a non-zero mask entry get
 * weight 0 (src/lensed.c:470-482; mask may be NULL).  lcu_model_get_weight
 * reads the current map back (the dumper's WHT layer, src/nested.c:236).
 */
int lcu_model_set_data(lcu_model* model, const float* image, const float* weight);
int lcu_model_make_weight(lcu_model* model, const float* gain_map, float gain, double offset, const int* mask);
int lcu_model_get_weight(lcu_model* model, float* weight);

/*
 * One likelihood evaluation: replaces the device part of loglike(),
 * src/nested.c:63-115.  params[npars] are the physical parameters in object
 * order, i.e. what the reference writes into its mapped buffer at
 * src/nested.c:70-72.  *lnew = -chi^2/2, chi^2 summed in double.
 */
int lcu_loglike(lcu_model* model, const float* params, double* lnew);

/* lcu_loglike in two halves, for a host that has work of its own between
 * proposing a point and needing its likelihood (the prior transform of the next
 * point, src/nested.c:43-61; book-keeping of the sampler): lcu_loglike_async
 * copies the parameters, starts the evaluation and returns a ticket at once;
 * lcu_loglike_wait returns the ticket's log-likelihood (src/nested.c:115).  Up
 * to two evaluations may be in flight (they run one after the other on the
 * model's stream; the second is queued while the first runs); a third
 * lcu_loglike_async, and any other evaluation call on the model, fails with
 * LCU_E_ARG until one of them has been waited for.  Same bits as lcu_loglike. */
int lcu_loglike_async(lcu_model* model, const float* params, int* ticket);
int lcu_loglike_wait(lcu_model* model, int ticket, double* lnew);

/* Batched entry point: B independent parameter points per call,
 * params[B][npars] -> lnew[B], host memory. */
int lcu_loglike_batch(lcu_model* model, size_t nbatch, const float* params, double* lnew);

/* Same with device-resident params / lnew, enqueued on the given CUDA stream
 * (a cudaStream_t passed as void*; NULL = the CUDA default stream) without
 * synchronising.  Must not overlap other calls on the same model. */
int lcu_loglike_batch_device(lcu_model* model, size_t nbatch, const float* d_params,
                             double* d_lnew, void* stream);

/*
 * Images of one parameter point for the dumper (src/nested.c:178-253).  Any
 * output may be NULL.  model = what is compared with the data (convolved if
 * there is a PSF), raw = quadrature value before convolution, err =
 * quadrature error estimate, chi = per-pixel weight*(model - image)^2.
 */
int lcu_render(lcu_model* model, const float* params, float* model_img, float* raw,
               float* err, float* chi);

/* object data block produced by set_params for one point (words 4-byte words) */
int lcu_set_params(lcu_model* model, const float* params, uint32_t* block);

/* Per-stage device times (CUDA events on the launching stream), replaces
 * --profile (src/profile.c:36-83).  Works for the host and the device entry
 * points; lcu_profile_get waits for the launches recorded so far. */
typedef struct
{
    unsigned long long evaluations;
    double upload_ms, set_params_ms, render_ms, convolve_ms, reduce_ms, download_ms;
} lcu_profile;
int lcu_profile_enable(lcu_model* model, int on);
int lcu_profile_get(lcu_model* model, lcu_profile* out);

/* FFMA micro-benchmark on the context's device: measured FP32 peak in
 * TFLOP/s, the denominator of the render roofline. */
int lcu_measure_fp32_peak(lcu_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif

#endif /* LENSED_CUDA_H */
