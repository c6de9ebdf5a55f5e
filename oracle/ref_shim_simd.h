/*
 * ref_shim_simd.h -- the REFERENCE's per-ray plugin text and generated
 * compute() compiled with `float` = 8 (AVX2) or 16 (AVX-512) consecutive work-items, one
 * SIMD register, the way a vectorising OpenCL CPU runtime (the Intel runtime of the
 * reference's CI, .travis.yml:60-68) executes them: work-items of a work-group
 * are packed into SIMD lanes, uniform data stays scalar.  TEST / BASELINE
 * INFRASTRUCTURE ONLY (oracle/build_ref.py, liblensed_ref_simd.so): this build
 * is timed, never used as a parity truth (its math functions are libmvec's
 * 4-ulp vector variants, as -cl-fast-relaxed-math permits).
 *
 * Included after ref_shim.h (scalar float2 / float4, qualifier macros).  The
 * object data blocks keep their scalar layout; per-ray code sees
 *     float -> vfloat,  float2 -> vfloat2,  float4 / mat22 -> vfloat4
 * through macros defined by the generated translation unit around the plugin
 * text.  No `this`, `local`, `global`, `constant`, `kernel` in this header: they
 * are macros here.
 */
#ifndef REF_SHIM_SIMD_H
#define REF_SHIM_SIMD_H

/* REF_SIMD = 8: AVX2 (x86-64-v3), = 16: AVX-512 (x86-64-v4); glibc libmvec's
 * vector math functions under their x86_64 vector-ABI names */
#if REF_SIMD == 16
#define REF_LANES 16
typedef float ref_vf __attribute__((vector_size(64)));
typedef int ref_vi __attribute__((vector_size(64)));
#define REF_MVEC1(f) _ZGVeN16v_##f
#define REF_MVEC2(f) _ZGVeN16vv_##f
#define REF_BCAST(s) {s, s, s, s, s, s, s, s, s, s, s, s, s, s, s, s}
#else
#define REF_LANES 8
typedef float ref_vf __attribute__((vector_size(32)));
typedef int ref_vi __attribute__((vector_size(32)));
#define REF_MVEC1(f) _ZGVdN8v_##f
#define REF_MVEC2(f) _ZGVdN8vv_##f
#define REF_BCAST(s) {s, s, s, s, s, s, s, s}
#endif

extern "C" {
#define REF_DECL1(f) ref_vf REF_MVEC1(f)(ref_vf);
REF_DECL1(expf) REF_DECL1(logf) REF_DECL1(sinf) REF_DECL1(cosf) REF_DECL1(tanf) REF_DECL1(atanf) REF_DECL1(atanhf)
REF_DECL1(asinf) REF_DECL1(acosf) REF_DECL1(sinhf) REF_DECL1(coshf) REF_DECL1(tanhf) REF_DECL1(exp2f) REF_DECL1(log2f)
REF_DECL1(log10f) REF_DECL1(log1pf) REF_DECL1(expm1f)
#undef REF_DECL1
ref_vf REF_MVEC2(powf)(ref_vf, ref_vf);
ref_vf REF_MVEC2(atan2f)(ref_vf, ref_vf);
ref_vf REF_MVEC2(hypotf)(ref_vf, ref_vf);
}

struct vfloat
{
    ref_vf v;
    vfloat() = default;
    vfloat(float s) : v REF_BCAST(s) {}
    vfloat(int s) : vfloat((float)s) {}
    vfloat(double s) : vfloat((float)s) {}
    vfloat(ref_vf x) : v(x) {}
};
struct vmask { ref_vi m; };

#define REF_VOP(op) \
    static inline vfloat operator op(vfloat a, vfloat b) { return vfloat(a.v op b.v); } \
    static inline vfloat& operator op##=(vfloat& a, vfloat b) { a.v = a.v op b.v; return a; }
REF_VOP(+) REF_VOP(-) REF_VOP(*) REF_VOP(/)
#undef REF_VOP
/* exact matches for scalar operands: float also converts to float2 / vfloat2 / vfloat4 */
#define REF_VSOP(op, T) \
    static inline vfloat operator op(vfloat a, T b) { return a op vfloat(b); } \
    static inline vfloat operator op(T a, vfloat b) { return vfloat(a) op b; }
#define REF_VSOPS(T) REF_VSOP(+, T) REF_VSOP(-, T) REF_VSOP(*, T) REF_VSOP(/, T)
REF_VSOPS(float) REF_VSOPS(int) REF_VSOPS(double)
#undef REF_VSOPS
#undef REF_VSOP
static inline vfloat operator-(vfloat a) { return vfloat(-a.v); }
static inline vfloat operator+(vfloat a) { return a; }
#define REF_VCMP(op) static inline vmask operator op(vfloat a, vfloat b) { return vmask{a.v op b.v}; }
REF_VCMP(<) REF_VCMP(>) REF_VCMP(<=) REF_VCMP(>=) REF_VCMP(==) REF_VCMP(!=)
#undef REF_VCMP
static inline vfloat ref_select(vmask c, vfloat a, vfloat b) { return vfloat(c.m ? a.v : b.v); }

struct vfloat2
{
    vfloat x, y;
    vfloat2() = default;
    explicit vfloat2(vfloat s) : x(s), y(s) {}
    vfloat2(float s) : x(s), y(s) {}
    vfloat2(int s) : x(s), y(s) {}
    vfloat2(vfloat a, vfloat b) : x(a), y(b) {}
    vfloat2(float2 s) : x(s.x), y(s.y) {}           /* uniform data entering per-ray arithmetic */
};
struct vfloat4
{
    vfloat x, y, z, w;
    vfloat4() = default;
    explicit vfloat4(vfloat s) : x(s), y(s), z(s), w(s) {}
    vfloat4(float s) : x(s), y(s), z(s), w(s) {}
    vfloat4(vfloat a, vfloat b, vfloat c, vfloat d) : x(a), y(b), z(c), w(d) {}
    vfloat4(float4 s) : x(s.x), y(s.y), z(s.z), w(s.w) {}
};

#define REF_V2OP(op) \
    static inline vfloat2 operator op(vfloat2 a, vfloat2 b) { return vfloat2(a.x op b.x, a.y op b.y); } \
    static inline vfloat2 operator op(vfloat2 a, vfloat b) { return vfloat2(a.x op b, a.y op b); } \
    static inline vfloat2 operator op(vfloat a, vfloat2 b) { return vfloat2(a op b.x, a op b.y); } \
    static inline vfloat2 operator op(vfloat2 a, float2 b) { return vfloat2(a.x op vfloat(b.x), a.y op vfloat(b.y)); } \
    static inline vfloat2 operator op(float2 a, vfloat2 b) { return vfloat2(vfloat(a.x) op b.x, vfloat(a.y) op b.y); } \
    static inline vfloat2 operator op(vfloat2 a, float b) { return vfloat2(a.x op vfloat(b), a.y op vfloat(b)); } \
    static inline vfloat2 operator op(float a, vfloat2 b) { return vfloat2(vfloat(a) op b.x, vfloat(a) op b.y); } \
    static inline vfloat2 operator op(vfloat a, float2 b) { return vfloat2(a op vfloat(b.x), a op vfloat(b.y)); } \
    static inline vfloat2 operator op(float2 a, vfloat b) { return vfloat2(vfloat(a.x) op b, vfloat(a.y) op b); } \
    static inline vfloat2& operator op##=(vfloat2& a, vfloat2 b) { a.x op##= b.x; a.y op##= b.y; return a; } \
    static inline vfloat2& operator op##=(vfloat2& a, vfloat b) { a.x op##= b; a.y op##= b; return a; }
REF_V2OP(+) REF_V2OP(-) REF_V2OP(*) REF_V2OP(/)
#undef REF_V2OP
static inline vfloat2 operator-(vfloat2 a) { return vfloat2(-a.x, -a.y); }
static inline vfloat2 ref_select(vmask c, vfloat2 a, vfloat2 b) { return vfloat2(ref_select(c, a.x, b.x), ref_select(c, a.y, b.y)); }

#define REF_V4OP(op) \
    static inline vfloat4 operator op(vfloat4 a, vfloat4 b) { return vfloat4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    static inline vfloat4 operator op(vfloat4 a, vfloat b) { return vfloat4(a.x op b, a.y op b, a.z op b, a.w op b); } \
    static inline vfloat4 operator op(vfloat a, vfloat4 b) { return vfloat4(a op b.x, a op b.y, a op b.z, a op b.w); }
REF_V4OP(+) REF_V4OP(-) REF_V4OP(*) REF_V4OP(/)
#undef REF_V4OP

/* OpenCL built-ins on the vector types (same conventions as ref_shim.h) */
#define REF_VFN1(name, f) static inline vfloat name(vfloat a) { return vfloat(REF_MVEC1(f)(a.v)); }
REF_VFN1(exp, expf) REF_VFN1(log, logf) REF_VFN1(sin, sinf) REF_VFN1(cos, cosf) REF_VFN1(tan, tanf) REF_VFN1(atan, atanf)
REF_VFN1(atanh, atanhf) REF_VFN1(asin, asinf) REF_VFN1(acos, acosf) REF_VFN1(sinh, sinhf) REF_VFN1(cosh, coshf)
REF_VFN1(tanh, tanhf) REF_VFN1(exp2, exp2f) REF_VFN1(log2, log2f) REF_VFN1(log10, log10f) REF_VFN1(log1p, log1pf)
REF_VFN1(expm1, expm1f)
#undef REF_VFN1
#if REF_SIMD == 16
static inline vfloat sqrt(vfloat a) { return vfloat((ref_vf)_mm512_sqrt_ps((__m512)a.v)); }
#else
static inline vfloat sqrt(vfloat a) { return vfloat((ref_vf)_mm256_sqrt_ps((__m256)a.v)); }
#endif
static inline vfloat rsqrt(vfloat a) { return vfloat(1.0f)/sqrt(a); }
static inline vfloat fabs(vfloat a) { return vfloat(a.v < 0.0f ? -a.v : a.v); }
static inline vfloat pow(vfloat a, vfloat b) { return vfloat(REF_MVEC2(powf)(a.v, b.v)); }
static inline vfloat powr(vfloat a, vfloat b) { return vfloat(REF_MVEC2(powf)(a.v, b.v)); }
static inline vfloat atan2(vfloat a, vfloat b) { return vfloat(REF_MVEC2(atan2f)(a.v, b.v)); }
static inline vfloat hypot(vfloat a, vfloat b) { return vfloat(REF_MVEC2(hypotf)(a.v, b.v)); }
static inline vfloat sincos(vfloat a, vfloat* c) { *c = vfloat(REF_MVEC1(cosf)(a.v)); return vfloat(REF_MVEC1(sinf)(a.v)); }
static inline vfloat fmin(vfloat a, vfloat b) { return vfloat(a.v < b.v ? a.v : b.v); }
static inline vfloat fmax(vfloat a, vfloat b) { return vfloat(a.v > b.v ? a.v : b.v); }
static inline vfloat dot(vfloat2 a, vfloat2 b) { return a.x*b.x + a.y*b.y; }
static inline vfloat dot(float2 a, vfloat2 b) { return vfloat(a.x)*b.x + vfloat(a.y)*b.y; }
static inline vfloat dot(vfloat2 a, float2 b) { return a.x*vfloat(b.x) + a.y*vfloat(b.y); }
static inline vfloat length(vfloat2 a) { return sqrt(dot(a, a)); }
static inline vfloat2 normalize(vfloat2 a) { vfloat l = length(a); return vfloat2(a.x/l, a.y/l); }
/* kernel/constants.cl:32-35 for a uniform (scalar) or per-ray matrix */
static inline vfloat2 mv22(float4 m, vfloat2 v) { return vfloat2(vfloat(m.x)*v.x + vfloat(m.y)*v.y, vfloat(m.z)*v.x + vfloat(m.w)*v.y); }
static inline vfloat2 mv22(vfloat4 m, vfloat2 v) { return vfloat2(m.x*v.x + m.y*v.y, m.z*v.x + m.w*v.y); }

/* entry point of one configuration's SIMD unit: pixels [k0, k1) of the render
 * kernel (kernel/lensed.cl:9-38), 8 consecutive work-items at a time */
typedef void (*ref_simd_render_fn)(const uint* data, const float* pcs, const float2* qq, const float2* ww, int nq,
                                   long k0, long k1, long size, int width, float* value, float* error);

#endif
