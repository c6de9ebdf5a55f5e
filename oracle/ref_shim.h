/*
 * ref_shim.h -- host (g++) environment for compiling the REFERENCE's own
 * OpenCL-C text: objects/<name>.cl, kernel/object.cl, kernel/constants.cl,
 * kernel/lensed.cl and the compute()/set_params() text its src/kernel.c
 * generates.  TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py).
 *
 * Work-items are emulated one at a time with a work-group size of 1: every
 * work-item copies the object block / fills the convolution cache itself and
 * barrier() is a no-op, which is a valid OpenCL execution of those kernels.
 * The image / PSF / quadrature macros the reference passes as -D options
 * (src/kernel.c:890-896) are bound to run-time variables so that one compiled
 * library serves every image size.
 *
 * Conventions for built-ins whose precision OpenCL leaves open (the oracle
 * port and the CUDA shim use the same): normalize(v) = v/|v| component-wise,
 * length = sqrt(dot), sincos(x, &c) = (sinf(x), c = cosf(x)), powr = powf.
 */
#ifndef REF_SHIM_H
#define REF_SHIM_H

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#ifdef REF_SIMD
#include <immintrin.h>        /* ref_shim_simd.h; system headers must precede the qualifier macros below */
#endif

/* OpenCL math built-ins are overloaded on float: make the unqualified calls in
 * the reference text pick the float overloads, not C's double functions */
using std::sqrt; using std::exp; using std::log; using std::sin; using std::cos; using std::tan;
using std::atan; using std::atanh; using std::atan2; using std::asin; using std::acos;
using std::sinh; using std::cosh; using std::tanh; using std::asinh; using std::acosh;
using std::pow; using std::tgamma; using std::lgamma; using std::fabs; using std::floor; using std::ceil;
using std::exp2; using std::log2; using std::log10; using std::log1p; using std::expm1; using std::hypot;
using std::fmin; using std::fmax; using std::fmod; using std::erf; using std::erfc; using std::cbrt;

typedef unsigned int uint;
typedef unsigned long ulong;

struct alignas(8) float2
{
    union
    {
        struct { float x, y; };
        struct { float s0, s1; };
    };
    float2() = default;
    float2(float v) : x(v), y(v) {}
    float2(float a, float b) : x(a), y(b) {}
};

/* half of a float4 (.lo .hi .xy .zw): a plain pair that converts to float2
 * (g++ does not allow members with constructors in anonymous structs) */
struct float2_half
{
    float x, y;
    operator float2() const { return float2(x, y); }
};

struct alignas(16) float4
{
    union
    {
        struct { float x, y, z, w; };
        struct { float s0, s1, s2, s3; };
        struct { float2_half lo, hi; };
        struct { float2_half xy, zw; };
    };
    float4() = default;
    float4(float v) : x(v), y(v), z(v), w(v) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};

struct char16 { char s[16]; };

#define REF_OP2(op) \
    static inline float2 operator op(float2 a, float2 b) { return float2(a.x op b.x, a.y op b.y); } \
    static inline float2 operator op(float2 a, float b) { return float2(a.x op b, a.y op b); } \
    static inline float2 operator op(float a, float2 b) { return float2(a op b.x, a op b.y); } \
    static inline float2& operator op##=(float2& a, float2 b) { a.x op##= b.x; a.y op##= b.y; return a; } \
    static inline float2& operator op##=(float2& a, float b) { a.x op##= b; a.y op##= b; return a; }
REF_OP2(+) REF_OP2(-) REF_OP2(*) REF_OP2(/)
#undef REF_OP2
static inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }

#define REF_OP4(op) \
    static inline float4 operator op(float4 a, float4 b) { return float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    static inline float4 operator op(float4 a, float b) { return float4(a.x op b, a.y op b, a.z op b, a.w op b); } \
    static inline float4 operator op(float a, float4 b) { return float4(a op b.x, a op b.y, a op b.z, a op b.w); }
REF_OP4(+) REF_OP4(-) REF_OP4(*) REF_OP4(/)
#undef REF_OP4

static inline float dot(float2 a, float2 b) { return a.x*b.x + a.y*b.y; }
static inline float length(float2 a) { return sqrtf(dot(a, a)); }
static inline float2 normalize(float2 a) { float l = length(a); return float2(a.x/l, a.y/l); }
static inline float powr(float x, float y) { return powf(x, y); }
static inline float sincos(float x, float* c) { *c = cosf(x); return sinf(x); }
static inline int mad24(int a, int b, int c) { return a*b + c; }
static inline float2 vload2(size_t i, const float* p) { return float2(p[2*i], p[2*i+1]); }
static inline char16 vload16(size_t i, const char* p) { char16 v; memcpy(v.s, p + 16*i, 16); return v; }
using std::min;
using std::max;

/* emulated work-item state (one work-item at a time per host thread) */
struct ref_workitem { size_t gid[3]; size_t gsz[3]; };
extern thread_local ref_workitem ref_wi;
static inline size_t get_global_id(uint d) { return ref_wi.gid[d]; }
static inline size_t get_global_size(uint d) { return ref_wi.gsz[d]; }
static inline size_t get_local_id(uint) { return 0; }
static inline size_t get_local_size(uint) { return 1; }
static inline size_t get_group_id(uint d) { return ref_wi.gid[d]; }
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2
static inline void barrier(int) {}

/* build options of src/kernel.c:890-896, as run-time values */
extern int ref_image_size, ref_image_width, ref_image_height;
extern int ref_psf, ref_psf_width, ref_psf_height, ref_quad_points;
#define IMAGE_SIZE ref_image_size
#define IMAGE_WIDTH ref_image_width
#define IMAGE_HEIGHT ref_image_height
#define PSF ref_psf
#define PSF_WIDTH ref_psf_width
#define PSF_HEIGHT ref_psf_height
#define QUAD_POINTS ref_quad_points

/* kernel entry points of one compiled program (one object list) */
struct ref_program
{
    void (*set_params)(ulong, int*, int*, const float*);
    void (*render)(ulong, const uint*, uint*, float4, const float2*, const float2*, float*, float*);
    void (*loglike)(const float*, const float*, const float*, float*);
    void (*convolve)(float*, const float*, float*, float*, float*);
};

/* metadata kernels of one object (src/kernel.c:41-62) */
struct ref_object
{
    const char* name;
    void (*meta)(int*, ulong*, ulong*);
    void (*params)(char16*, int*, float2*, float*);
};

/* OpenCL address-space and kernel qualifiers; must come after all system
 * headers.  `this` is an ordinary parameter name in object files. */
#define kernel
#define global
#define local
#define constant const
#define this this_

#endif
