/*
 * lensed_oracle.c -- CPU restatement of Lensed's per-likelihood model-image path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (lensed_b200/, include/)
 * may link, import or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the *checker* and as the timed CPU baseline.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit (strict
 * float32 build) against oracle/_ref/liblensed_ref.so, which is the
 * reference's own objects/<name>.cl + kernel/lensed.cl + generated
 * compute()/set_params() text compiled on the host (see oracle/build_ref.py),
 * and against the reference's 16 golden FITS images at the chi^2/dof << 1
 * level (tests/golden/ref_goldens.npz).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the upstream source tree).  Arithmetic is done in `real` (float by default,
 * double with -DORC_DOUBLE for the float32 noise-floor twin), one rounding per
 * operation in the order the reference source text writes them; build the
 * strict variant with -ffp-contract=off.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORC_DOUBLE
typedef double real;
#define RL(x) x
#define r_sqrt sqrt
#define r_exp exp
#define r_log log
#define r_sin sin
#define r_cos cos
#define r_atan atan
#define r_atanh atanh
#define r_atan2 atan2
#define r_pow pow
#define r_tgamma tgamma
#define R_HUGE HUGE_VAL
#else
typedef float real;
#define RL(x) x##f
#define r_sqrt sqrtf
#define r_exp expf
#define r_log logf
#define r_sin sinf
#define r_cos cosf
#define r_atan atanf
#define r_atanh atanhf
#define r_atan2 atan2f
#define r_pow powf
#define r_tgamma tgammaf
#define R_HUGE HUGE_VALF
#endif

/* kernel/constants.cl:6-14 */
#define K_PI RL(3.1415926535897932384626433832795028841971693993751)
#define K_DEG2RAD RL(0.0174532925199432957692369076848861271344287188854)
#define K_LOG_10 RL(2.3025850929940456840179914546843642076011014886288)
#define K_LOG_PI RL(1.1447298858494001741434273513530587116472948129153)
/* objects/devauc.cl:2-4 */
#define K_DEVAUC_B RL(7.6692494425008039044)
#define K_DEVAUC_C RL(296.826303766893)

/* kernel/object.cl:2-19 */
enum { T_LENS = 'L', T_SOURCE = 'S', T_FOREGROUND = 'F' };
enum { P_PARAMETER = 0, P_POSITION_X, P_POSITION_Y, P_RADIUS, P_MAGNITUDE, P_AXIS_RATIO, P_POS_ANGLE };

typedef struct { real x, y; } r2;
/* mat22 = float4 (kernel/constants.cl:29): lo = (a, b), hi = (c, d) */
typedef struct { real a, b, c, d; } r4;

static inline r2 mk2(real x, real y) { r2 v = { x, y }; return v; }
static inline r2 add2(r2 a, r2 b) { return mk2(a.x + b.x, a.y + b.y); }
static inline r2 sub2(r2 a, r2 b) { return mk2(a.x - b.x, a.y - b.y); }
static inline r2 scl2(real s, r2 v) { return mk2(s*v.x, s*v.y); }
static inline real dot2(r2 a, r2 b) { return a.x*b.x + a.y*b.y; }
static inline real len2(r2 v) { return r_sqrt(dot2(v, v)); }
/* kernel/constants.cl:32-35 */
static inline r2 mv22(r4 m, r2 v) { return mk2(m.a*v.x + m.b*v.y, m.c*v.x + m.d*v.y); }
/* OpenCL normalize(): v/|v| component-wise (the same convention is used by the
 * host shim of oracle/_ref and by the device shim of the product) */
static inline r2 norm2(r2 v) { real l = len2(v); return mk2(v.x/l, v.y/l); }

static inline r2 ld2(const real* D, int o) { return mk2(D[o], D[o+1]); }
static inline r4 ld4(const real* D, int o) { r4 m = { D[o], D[o+1], D[o+2], D[o+3] }; return m; }
static inline void st2(real* D, int o, real x, real y) { D[o] = x; D[o+1] = y; }
static inline void st4(real* D, int o, real a, real b, real c, real d) { D[o] = a; D[o+1] = b; D[o+2] = c; D[o+3] = d; }

/* ------------------------------------------------------------------------
 * object catalogue: parameter lists (objects/<name>.cl `params` block) and
 * OpenCL struct layout of the `data` block (float: 4/4, float2: 8/8,
 * float4: 16/16 size/alignment, struct padded to its largest alignment;
 * size reported in 4-byte words as src/input/objects.c:139).
 * ---------------------------------------------------------------------- */
#define MAXPAR 8
#define MAXFLD 8
typedef struct { const char* name; int type; real lo, hi; real defval; } pardef;

typedef void (*set_fn)(real* D, const int* o, const real* p);
typedef r2 (*defl_fn)(const real* D, const int* o, r2 x);
typedef real (*bright_fn)(const real* D, const int* o, r2 x);

typedef struct {
    const char* name;
    int type;
    int npar;
    pardef par[MAXPAR];
    int nfld;
    int fld[MAXFLD];      /* field widths in words: 1, 2 or 4 */
    set_fn set;
    defl_fn defl;         /* LENS */
    bright_fn bright;     /* SOURCE: brightness(); FOREGROUND: foreground() */
    int off[MAXFLD];      /* filled by layout() */
    int words;
} objdef;

#define UNB 0, 0
#define POSB 0, 3.402823466e+38F

/* ---- lenses ---------------------------------------------------------- */

/* objects/sis.cl:19-33 */
static r2 sis_defl(const real* D, const int* o, r2 x)
{ return scl2(D[o[1]], norm2(sub2(x, ld2(D, o[0])))); }
static void sis_set(real* D, const int* o, const real* p)
{ st2(D, o[0], p[0], p[1]); D[o[1]] = p[2]; }

/* objects/sis_plus_shear.cl:21-40 */
static r2 sisg_defl(const real* D, const int* o, r2 x)
{
    x = sub2(x, ld2(D, o[0]));
    return add2(scl2(D[o[2]], norm2(x)), mv22(ld4(D, o[1]), x));
}
static void sisg_set(real* D, const int* o, const real* p)
{
    st2(D, o[0], p[0], p[1]);
    D[o[2]] = p[2];
    st4(D, o[1], p[3], p[4], p[4], -p[3]);
}

/* objects/nsis.cl:21-40 */
static r2 nsis_defl(const real* D, const int* o, r2 x)
{
    x = sub2(x, ld2(D, o[0]));
    return scl2(D[o[1]]/(D[o[2]] + len2(x)), x);
}
static void nsis_set(real* D, const int* o, const real* p)
{ st2(D, o[0], p[0], p[1]); D[o[1]] = p[2]; D[o[2]] = p[3]; }

/* objects/point_mass.cl:18-32 */
static r2 pm_defl(const real* D, const int* o, r2 x)
{
    x = sub2(x, ld2(D, o[0]));
    return scl2(D[o[1]]/dot2(x, x), x);
}
static void pm_set(real* D, const int* o, const real* p)
{ st2(D, o[0], p[0], p[1]); D[o[1]] = p[2]*p[2]; }

/* objects/sie.cl:27-44 (fields: x m w q2 e d) */
static r2 sie_defl(const real* D, const int* o, r2 x)
{
    r2 y; real r;
    x = sub2(x, ld2(D, o[0]));
    y = mv22(ld4(D, o[1]), x);
    r = D[o[4]]/r_sqrt(D[o[3]]*y.x*y.x + y.y*y.y);
    y = scl2(D[o[5]], mk2(r_atan(y.x*r), r_atanh(y.y*r)));
    return mv22(ld4(D, o[2]), y);
}
/* objects/sie.cl:46-64 */
static void sie_set(real* D, const int* o, const real* p)
{
    real x = p[0], y = p[1], r = p[2], q = p[3], pa = p[4];
    real c = r_cos(pa*K_DEG2RAD);
    real s = r_sin(pa*K_DEG2RAD);
    st2(D, o[0], x, y);
    st4(D, o[1], c, s, -s, c);
    st4(D, o[2], c, -s, s, c);
    D[o[3]] = q*q;
    D[o[4]] = r_sqrt(1 - q*q);
    D[o[5]] = r*r_sqrt(q)/r_sqrt(1 - q*q);
}

/* objects/sie_plus_shear.cl:30-47 (fields: x m w g q2 e d) */
static r2 sieg_defl(const real* D, const int* o, r2 x)
{
    r2 y; real r;
    x = sub2(x, ld2(D, o[0]));
    y = mv22(ld4(D, o[1]), x);
    r = D[o[5]]/r_sqrt(D[o[4]]*y.x*y.x + y.y*y.y);
    y = scl2(D[o[6]], mk2(r_atan(y.x*r), r_atanh(y.y*r)));
    return add2(mv22(ld4(D, o[2]), y), mv22(ld4(D, o[3]), x));
}
/* objects/sie_plus_shear.cl:49-70 */
static void sieg_set(real* D, const int* o, const real* p)
{
    real x = p[0], y = p[1], r = p[2], q = p[3], pa = p[4], g1 = p[5], g2 = p[6];
    real c = r_cos(pa*K_DEG2RAD);
    real s = r_sin(pa*K_DEG2RAD);
    st2(D, o[0], x, y);
    st4(D, o[1], c, s, -s, c);
    st4(D, o[2], c, -s, s, c);
    st4(D, o[3], g1, g2, g2, -g1);
    D[o[4]] = q*q;
    D[o[5]] = r_sqrt(1 - q*q);
    D[o[6]] = r*r_sqrt(q)/r_sqrt(1 - q*q);
}

/* objects/nsie.cl:29-46 (fields: x m w rc q2 e d) */
static r2 nsie_defl(const real* D, const int* o, r2 x)
{
    r2 y; real r;
    real rc = D[o[3]], q2 = D[o[4]], e = D[o[5]], d = D[o[6]];
    x = sub2(x, ld2(D, o[0]));
    y = mv22(ld4(D, o[1]), x);
    r = r_sqrt(q2*y.x*y.x + y.y*y.y);
    y = scl2(d, mk2(r_atan(y.x*e/(rc + r)), r_atanh(y.y*e/(rc*q2 + r))));
    return mv22(ld4(D, o[2]), y);
}
/* objects/nsie.cl:48-70 */
static void nsie_set(real* D, const int* o, const real* p)
{
    real x = p[0], y = p[1], r = p[2], rc = p[3], q = p[4], pa = p[5];
    real c = r_cos(pa*K_DEG2RAD);
    real s = r_sin(pa*K_DEG2RAD);
    st2(D, o[0], x, y);
    st4(D, o[1], c, s, -s, c);
    st4(D, o[2], c, -s, s, c);
    D[o[3]] = rc;
    D[o[4]] = q*q;
    D[o[5]] = r_sqrt(1 - q*q);
    D[o[6]] = r*r_sqrt(q)/r_sqrt(1 - q*q);
}

/* Angular series of the elliptical power law, objects/epl.cl:37-64 and
 * objects/epl_plus_shear.cl:42-69: x is already in the scaled elliptical
 * frame.  Returns a = n r^(1-t) * sum_k A_k (before the back-rotation). */
static r2 epl_series(r2 x, real t, real f, real n)
{
    const real T = 2 - t;
    real r = len2(x);
    real phi = r_atan2(x.y, x.x);
    real s = r_sin(phi), c = r_cos(phi);
    real s2 = r_sin(2*phi), c2 = r_cos(2*phi);
    r4 R = { c2, -s2, s2, c2 };
    r2 a, A;
    int k;
    a = A = mk2(c, s);
    for(k = 1; k <= 10; ++k)
    {
        /* a += A = -f*(2*k - T)/(2*k + T)*mv22(R, A); */
        A = scl2(-f*(2*k - T)/(2*k + T), mv22(R, A));
        a = add2(a, A);
    }
    return scl2(n*r_pow(r, 1 - t), a);
}

/* objects/epl.cl:25-68 (fields: x m w t f n) */
static r2 epl_defl(const real* D, const int* o, r2 x)
{
    x = mv22(ld4(D, o[1]), sub2(x, ld2(D, o[0])));
    return mv22(ld4(D, o[2]), epl_series(x, D[o[3]], D[o[4]], D[o[5]]));
}
/* objects/epl.cl:70-93 */
static void epl_set(real* D, const int* o, const real* p)
{
    real x1 = p[0], x2 = p[1], r = p[2], t = p[3], q = p[4], pa = p[5];
    real c = r_cos(pa*K_DEG2RAD);
    real s = r_sin(pa*K_DEG2RAD);
    real k = 1/r/r_sqrt(q);
    st2(D, o[0], x1, x2);
    st4(D, o[1], k*(q*c), k*(q*s), k*(-s), k*c);
    st4(D, o[2], c, -s, s, c);
    D[o[3]] = t;
    D[o[4]] = (1 - q)/(1 + q);
    D[o[5]] = 2*r*r_sqrt(q)/(1 + q);
}

/* objects/epl_plus_shear.cl:28-76 (fields: x m w g t f n) */
static r2 eplg_defl(const real* D, const int* o, r2 x)
{
    r2 dx = sub2(x, ld2(D, o[0]));
    r2 y;
    x = mv22(ld4(D, o[1]), dx);
    y = mv22(ld4(D, o[2]), epl_series(x, D[o[4]], D[o[5]], D[o[6]]));
    return add2(y, mv22(ld4(D, o[3]), dx));
}
/* objects/epl_plus_shear.cl:78-104 */
static void eplg_set(real* D, const int* o, const real* p)
{
    real x1 = p[0], x2 = p[1], r = p[2], t = p[3], q = p[4], pa = p[5], g1 = p[6], g2 = p[7];
    real c = r_cos(pa*K_DEG2RAD);
    real s = r_sin(pa*K_DEG2RAD);
    real k = 1/r/r_sqrt(q);
    st2(D, o[0], x1, x2);
    st4(D, o[1], k*(q*c), k*(q*s), k*(-s), k*c);
    st4(D, o[2], c, -s, s, c);
    D[o[4]] = t;
    D[o[5]] = (1 - q)/(1 + q);
    D[o[6]] = 2*r*r_sqrt(q)/(1 + q);
    st4(D, o[3], g1, g2, g2, -g1);
}

/* ---- sources --------------------------------------------------------- */

/* objects/sersic.cl:23-27, objects/sersic-old.cl:23-27 (fields: x t log0 log1 m) */
static real sersic_bright(const real* D, const int* o, r2 x)
{
    r2 y = mv22(ld4(D, o[1]), sub2(x, ld2(D, o[0])));
    return r_exp(D[o[2]] - r_exp(D[o[3]] + D[o[4]]*r_log(dot2(y, y))));
}
static void sersic_fill(real* D, const int* o, const real* p, real b)
{
    real x = p[0], y = p[1], r = p[2], mag = p[3], n = p[4], q = p[5], a = p[6];
    real c = r_cos(a*K_DEG2RAD);
    real s = r_sin(a*K_DEG2RAD);
    real sq = r_sqrt(q);
    st2(D, o[0], x, y);
    st4(D, o[1], (q*c)/sq, (q*s)/sq, (-s)/sq, c/sq);
    D[o[2]] = RL(-0.4)*mag*K_LOG_10 + 2*n*r_log(b) - K_LOG_PI - 2*r_log(r) - r_log(r_tgamma(2*n+1));
    D[o[3]] = r_log(b) - r_log(r)/n;
    D[o[4]] = RL(0.5)/n;
}
/* objects/sersic.cl:29-47; the integer literals keep the reference's types
 * (int and 64-bit integer converted to floating point at the multiplication) */
static void sersic_set(real* D, const int* o, const real* p)
{
    real n = p[4];
    real b = n > RL(0.36)
        ? RL(2.0)*n - RL(1.0)/3 + RL(4.0)/(405*n) + RL(46.0)/(25515*(n*n)) + RL(131.0)/(1148175*(n*n*n)) - RL(2194697.0)/(30690717750*(n*n*n*n))
        : RL(0.01945) - RL(0.8902)*n + RL(10.95)*(n*n) - RL(19.67)*(n*n*n) + RL(13.43)*(n*n*n*n);
    sersic_fill(D, o, p, b);
}
/* objects/sersic-old.cl:29-45 */
static void sersic_old_set(real* D, const int* o, const real* p)
{
    real b = RL(1.9992)*p[4] - RL(0.3271);
    sersic_fill(D, o, p, b);
}

/* shared (x, t = plain rotate+squash, scale, norm) setter head of
 * objects/devauc.cl:32-48, exponential.cl:27-43, gauss.cl:28-42 */
static void disk_head(real* D, const int* o, const real* p)
{
    real q = p[4], pa = p[5];
    real c = r_cos(pa*K_DEG2RAD);
    real s = r_sin(pa*K_DEG2RAD);
    st2(D, o[0], p[0], p[1]);
    st4(D, o[1], q*c, q*s, -s, c);
}

/* objects/devauc.cl:26-30 (fields: x t rs norm) */
static real devauc_bright(const real* D, const int* o, r2 x)
{
    return D[o[3]]*r_exp(-K_DEVAUC_B*r_sqrt(r_sqrt(len2(mv22(ld4(D, o[1]), sub2(x, ld2(D, o[0]))))/D[o[2]])));
}
static void devauc_set(real* D, const int* o, const real* p)
{
    real r = p[2], mag = p[3], q = p[4];
    disk_head(D, o, p);
    D[o[2]] = r;
    D[o[3]] = r_exp(RL(-0.4)*mag*K_LOG_10)/K_PI/r/r/q*K_DEVAUC_C;
}

/* objects/exponential.cl:21-25 (fields: x t rs norm) */
static real expo_bright(const real* D, const int* o, r2 x)
{
    return D[o[3]]*r_exp(-len2(mv22(ld4(D, o[1]), sub2(x, ld2(D, o[0]))))/D[o[2]]);
}
static void expo_set(real* D, const int* o, const real* p)
{
    real rs = p[2], mag = p[3], q = p[4];
    disk_head(D, o, p);
    D[o[2]] = rs;
    D[o[3]] = r_exp(RL(-0.4)*mag*K_LOG_10)*RL(0.5)/K_PI/rs/rs/q;
}

/* objects/gauss.cl:21-26 (fields: x t s2 norm) */
static real gauss_bright(const real* D, const int* o, r2 x)
{
    r2 y = mv22(ld4(D, o[1]), sub2(x, ld2(D, o[0])));
    return D[o[3]]*r_exp(RL(-0.5)*dot2(y, y)/D[o[2]]);
}
static void gauss_set(real* D, const int* o, const real* p)
{
    real sigma = p[2], mag = p[3], q = p[4];
    disk_head(D, o, p);
    D[o[2]] = sigma*sigma;
    D[o[3]] = r_exp(RL(-0.4)*mag*K_LOG_10)*RL(0.5)/K_PI/D[o[2]]/q;
}

/* ---- foreground ------------------------------------------------------ */

/* objects/sky.cl:16-25 (fields: bg grad) */
static real sky_fore(const real* D, const int* o, r2 x)
{ return D[o[0]] + dot2(ld2(D, o[1]), sub2(x, mk2(1, 1))); }
static void sky_set(real* D, const int* o, const real* p)
{ D[o[0]] = p[0]; st2(D, o[1], p[1], p[2]); }

#define XY {"x", P_POSITION_X, UNB, 0}, {"y", P_POSITION_Y, UNB, 0}
#define PAR(n, t) {n, t, UNB, 0}

static objdef CATALOGUE[] = {
    { "sis", T_LENS, 3, { XY, PAR("r", P_RADIUS) }, 2, {2, 1}, sis_set, sis_defl, 0 },
    { "sis_plus_shear", T_LENS, 5, { XY, PAR("r", P_RADIUS), PAR("g1", 0), PAR("g2", 0) }, 3, {2, 4, 1}, sisg_set, sisg_defl, 0 },
    { "sie", T_LENS, 5, { XY, PAR("r", P_RADIUS), PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      6, {2, 4, 4, 1, 1, 1}, sie_set, sie_defl, 0 },
    { "sie_plus_shear", T_LENS, 7, { XY, PAR("r", P_RADIUS), PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE), PAR("g1", 0), PAR("g2", 0) },
      7, {2, 4, 4, 4, 1, 1, 1}, sieg_set, sieg_defl, 0 },
    { "nsis", T_LENS, 4, { XY, PAR("r", P_RADIUS), PAR("rc", P_RADIUS) }, 3, {2, 1, 1}, nsis_set, nsis_defl, 0 },
    { "nsie", T_LENS, 6, { XY, PAR("r", P_RADIUS), PAR("rc", P_RADIUS), PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      7, {2, 4, 4, 1, 1, 1, 1}, nsie_set, nsie_defl, 0 },
    { "point_mass", T_LENS, 3, { XY, PAR("r", P_RADIUS) }, 2, {2, 1}, pm_set, pm_defl, 0 },
    { "epl", T_LENS, 6, { XY, PAR("r", P_RADIUS), {"t", 0, 0, 2, 0}, PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      6, {2, 4, 4, 1, 1, 1}, epl_set, epl_defl, 0 },
    { "epl_plus_shear", T_LENS, 8, { XY, PAR("r", P_RADIUS), {"t", 0, 0, 2, 0}, PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE), PAR("g1", 0), PAR("g2", 0) },
      7, {2, 4, 4, 4, 1, 1, 1}, eplg_set, eplg_defl, 0 },
    { "sersic", T_SOURCE, 7, { XY, PAR("r", P_RADIUS), PAR("mag", P_MAGNITUDE), {"n", 0, POSB, 0}, PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      5, {2, 4, 1, 1, 1}, sersic_set, 0, sersic_bright },
    { "sersic-old", T_SOURCE, 7, { XY, PAR("r", P_RADIUS), PAR("mag", P_MAGNITUDE), {"n", 0, 0.5f, 8.0f, 0}, PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      5, {2, 4, 1, 1, 1}, sersic_old_set, 0, sersic_bright },
    { "devauc", T_SOURCE, 6, { XY, PAR("r", P_RADIUS), PAR("mag", P_MAGNITUDE), PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      4, {2, 4, 1, 1}, devauc_set, 0, devauc_bright },
    { "exponential", T_SOURCE, 6, { XY, PAR("rs", P_RADIUS), PAR("mag", P_MAGNITUDE), PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      4, {2, 4, 1, 1}, expo_set, 0, expo_bright },
    { "gauss", T_SOURCE, 6, { XY, PAR("sigma", P_RADIUS), PAR("mag", P_MAGNITUDE), PAR("q", P_AXIS_RATIO), PAR("pa", P_POS_ANGLE) },
      4, {2, 4, 1, 1}, gauss_set, 0, gauss_bright },
    /* objects/sky.cl:3-8: dx, dy default to "-0.0f" = zero default value */
    { "sky", T_FOREGROUND, 3, { PAR("bg", 0), {"dx", 0, UNB, -0.0f}, {"dy", 0, UNB, -0.0f} }, 2, {1, 2}, sky_set, 0, sky_fore },
};
#define NCAT ((int)(sizeof(CATALOGUE)/sizeof(CATALOGUE[0])))

static int catalogue_ready = 0;

/* OpenCL struct layout; size in 4-byte words as src/input/objects.c:139 */
static void layout(objdef* d)
{
    int pos = 0, maxal = 1, i;
    for(i = 0; i < d->nfld; ++i)
    {
        int al = d->fld[i];
        if(al > maxal) maxal = al;
        pos = (pos + al - 1)/al*al;
        d->off[i] = pos;
        pos += d->fld[i];
    }
    d->words = (pos + maxal - 1)/maxal*maxal;
}

static objdef* find_object(const char* name)
{
    int i;
    if(!catalogue_ready)
    {
        for(i = 0; i < NCAT; ++i)
            layout(&CATALOGUE[i]);
        catalogue_ready = 1;
    }
    for(i = 0; i < NCAT; ++i)
        if(strcmp(CATALOGUE[i].name, name) == 0)
            return &CATALOGUE[i];
    return NULL;
}

int orc_real_size(void) { return (int)sizeof(real); }
int orc_object_count(void) { return NCAT; }
const char* orc_object_name(int i) { return i >= 0 && i < NCAT ? CATALOGUE[i].name : NULL; }

/* meta_<name> kernel, src/kernel.c:41-49 + src/input/objects.c:118-144 */
int orc_object_info(const char* name, int* type, size_t* words, size_t* npar)
{
    objdef* d = find_object(name);
    if(!d) return 1;
    *type = d->type; *words = (size_t)d->words; *npar = (size_t)d->npar;
    return 0;
}

/* params_<name> kernel, src/kernel.c:52-62 */
int orc_object_param(const char* name, size_t j, char pname[16], int* type, float bounds[2], float* defval)
{
    objdef* d = find_object(name);
    if(!d || j >= (size_t)d->npar) return 1;
    memset(pname, 0, 16);
    strncpy(pname, d->par[j].name, 15);
    *type = d->par[j].type;
    bounds[0] = (float)d->par[j].lo; bounds[1] = (float)d->par[j].hi;
    *defval = (float)d->par[j].defval;
    return 0;
}

/* ------------------------------------------------------------------------
 * quadrature rules, src/quadrature.c:21-43 + src/quad/ tables.  The Cartesian
 * Gauss-Kronrod tables are rebuilt from their 1-D factors (x outer, y inner;
 * weight k_i*k_j; error weight (k_i-g_i)*(k_j-g_j)); the Genz-Malik rule from
 * its closed form (Genz & Malik 1980, n = 2, unit square).
 * ---------------------------------------------------------------------- */
#include "quad_tables.inc"

static int quad_build(const char* rule, double (*pts)[2], double* wht, double* err)
{
    int n = 0, i, j;
    if(strcmp(rule, "point") == 0)
    {
        n = 1;
        if(pts) { pts[0][0] = pts[0][1] = 0; wht[0] = 1; err[0] = 0; }
    }
    else if(strcmp(rule, "sub2") == 0 || strcmp(rule, "sub4") == 0)
    {
        int m = rule[3] - '0';
        n = m*m;
        if(pts)
            for(i = 0; i < m; ++i)
                for(j = 0; j < m; ++j)
                {
                    pts[i*m+j][0] = (i + 0.5)/m - 0.5;
                    pts[i*m+j][1] = (j + 0.5)/m - 0.5;
                    wht[i*m+j] = 1.0/n;
                    err[i*m+j] = 0;
                }
    }
    else if(strcmp(rule, "gm75") == 0)
    {
        n = 17;
        if(pts)
        {
            const double l2 = 0.5*sqrt(9.0/70), l3 = 0.5*sqrt(9.0/10), l5 = 0.5*sqrt(9.0/19);
            /* degree-7 weights and degree-5 embedded weights for n = 2,
             * normalised to unit volume */
            const double w1 = (12824.0 - 9120.0*2 + 400.0*4)/19683, w2 = 980.0/6561,
                         w3 = (1820.0 - 400.0*2)/19683, w4 = 200.0/19683, w5 = 6859.0/19683/4;
            const double e1 = (729.0 - 950.0*2 + 50.0*4)/729, e2 = 245.0/486,
                         e3 = (265.0 - 100.0*2)/1458, e4 = 25.0/729;
            const double ax[2] = { l2, l3 };
            const double aw[2] = { w2, w3 }, ae[2] = { e2, e3 };
            const double dg[2] = { l3, l5 };
            const double dw[2] = { w4, w5 }, de[2] = { e4, 0 };
            int k = 0, s;
            pts[k][0] = 0; pts[k][1] = 0; wht[k] = w1; err[k] = w1 - e1; ++k;
            for(s = 0; s < 2; ++s)
            {
                const double X[4][2] = { {-ax[s], 0}, {0, -ax[s]}, {0, ax[s]}, {ax[s], 0} };
                for(i = 0; i < 4; ++i, ++k)
                { pts[k][0] = X[i][0]; pts[k][1] = X[i][1]; wht[k] = aw[s]; err[k] = aw[s] - ae[s]; }
            }
            for(s = 0; s < 2; ++s)
            {
                const double X[4][2] = { {-dg[s], -dg[s]}, {-dg[s], dg[s]}, {dg[s], -dg[s]}, {dg[s], dg[s]} };
                for(i = 0; i < 4; ++i, ++k)
                { pts[k][0] = X[i][0]; pts[k][1] = X[i][1]; wht[k] = dw[s]; err[k] = dw[s] - de[s]; }
            }
        }
    }
    else
    {
        const double *X, *K, *G; int m;
        if(strcmp(rule, "g3k7") == 0) { m = 7; X = GK_G3K7_X; K = GK_G3K7_K; G = GK_G3K7_G; }
        else if(strcmp(rule, "g5k11") == 0) { m = 11; X = GK_G5K11_X; K = GK_G5K11_K; G = GK_G5K11_G; }
        else if(strcmp(rule, "g7k15") == 0) { m = 15; X = GK_G7K15_X; K = GK_G7K15_K; G = GK_G7K15_G; }
        else return -1;
        n = m*m;
        if(pts)
            for(i = 0; i < m; ++i)
                for(j = 0; j < m; ++j)
                {
                    pts[i*m+j][0] = X[i];
                    pts[i*m+j][1] = X[j];
                    wht[i*m+j] = K[i]*K[j];
                    err[i*m+j] = (K[i] - G[i])*(K[j] - G[j]);
                }
    }
    return n;
}

int orc_quad_size(const char* rule) { return quad_build(rule, NULL, NULL, NULL); }

/* quad_rule(), src/quadrature.c:32-43: abscissae scaled by the pixel scale in
 * double, then narrowed to float; ww = (weight, error weight) */
int orc_quad_rule(const char* rule, double sx, double sy, float* qq, float* ww)
{
    double pts[225][2], wht[225], err[225];
    int n = quad_build(rule, pts, wht, err), i;
    if(n < 0) return -1;
    for(i = 0; i < n; ++i)
    {
        qq[2*i+0] = (float)(sx*pts[i][0]);
        qq[2*i+1] = (float)(sy*pts[i][1]);
        ww[2*i+0] = (float)wht[i];
        ww[2*i+1] = (float)err[i];
    }
    return n;
}

/* ------------------------------------------------------------------------
 * model = what src/lensed.c:644-1112 sets up on the device
 * ---------------------------------------------------------------------- */
typedef struct {
    objdef* def;
    size_t d;       /* word offset of the data block (src/kernel.c:321-383 `d`) */
    size_t p;       /* offset of first parameter (src/kernel.c:590-633 `p`) */
    int ipp[MAXPAR];
} mobj;

typedef struct orc_model {
    size_t nobjs; mobj* objs;
    size_t npars, words;
    size_t width, height, size;
    real pcs[4];
    size_t nq; real* qq; real* ww;
    real* image; real* weight;
    real* psf; size_t psfw, psfh;
    real* block;
    real* value; real* error; real* conv; real* chi;
} orc_model;

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if(n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_model_free(orc_model* m)
{
    if(!m) return;
    free(m->objs); free(m->qq); free(m->ww); free(m->image); free(m->weight); free(m->psf);
    free(m->block); free(m->value); free(m->error); free(m->conv); free(m->chi);
    free(m);
}

static real* dup_real(const float* src, size_t n)
{
    real* r = malloc((n ? n : 1)*sizeof(real));
    size_t i;
    for(i = 0; i < n; ++i) r[i] = src[i];
    return r;
}

/* names[]: object names in ini order; ipp: one flag per parameter in object /
 * parameter order (NULL = none).  pcs = (rx, ry, sx, sy); the even-PSF
 * half-pixel shift of src/lensed.c:885-891 is applied here. */
orc_model* orc_model_create(size_t nobjs, const char* const* names, const int* ipp,
                            size_t width, size_t height, const float pcs[4],
                            size_t nq, const float* qq, const float* ww,
                            const float* image, const float* weight,
                            const float* psf, size_t psfw, size_t psfh)
{
    orc_model* m = calloc(1, sizeof(orc_model));
    size_t i, d = 0, p = 0;
    int j;
    m->nobjs = nobjs;
    m->objs = calloc(nobjs ? nobjs : 1, sizeof(mobj));
    for(i = 0; i < nobjs; ++i)
    {
        objdef* def = find_object(names[i]);
        if(!def) { orc_model_free(m); return NULL; }
        m->objs[i].def = def;
        m->objs[i].d = d;
        m->objs[i].p = p;
        for(j = 0; j < def->npar; ++j)
            m->objs[i].ipp[j] = ipp ? ipp[p + j] : 0;
        d += def->words;         /* src/lensed.c:826-828 */
        p += def->npar;
    }
    m->words = d; m->npars = p;
    m->width = width; m->height = height; m->size = width*height;
    for(j = 0; j < 4; ++j) m->pcs[j] = pcs[j];
    if(psf)
    {
        /* src/lensed.c:885-891 (the reference adds 0.5 in float) */
        if(psfw % 2 == 0) m->pcs[0] += RL(0.5);
        if(psfh % 2 == 0) m->pcs[1] += RL(0.5);
    }
    m->nq = nq;
    m->qq = dup_real(qq, 2*nq);
    m->ww = dup_real(ww, 2*nq);
    m->image = dup_real(image, m->size);
    m->weight = dup_real(weight, m->size);
    if(psf) { m->psf = dup_real(psf, psfw*psfh); m->psfw = psfw; m->psfh = psfh; }
    m->block = calloc(m->words ? m->words : 1, sizeof(real));
    m->value = calloc(m->size, sizeof(real));
    m->error = calloc(m->size, sizeof(real));
    m->conv = calloc(m->size, sizeof(real));
    m->chi = calloc(m->size, sizeof(real));
    return m;
}

size_t orc_model_npars(const orc_model* m) { return m->npars; }
size_t orc_model_words(const orc_model* m) { return m->words; }

/* "apply deflection to ray, if finite", src/kernel.c:86-91 and :147-150 */
static inline r2 deflect(r2 y, r2 a)
{
    if(dot2(a, a) < R_HUGE)
        return sub2(y, a);
    return sub2(y, mk2(RL(1E10), RL(1E10)));
}

/* generated compute(), src/kernel.c:65-111 (templates) and :321-383 (object
 * loop): objects are visited in ini order; a change of the non-foreground
 * type away from LENS closes the lens plane and deflects the ray; sources see
 * the current ray position y, foregrounds the image-plane position x. */
static real compute(const orc_model* m, const real* data, r2 x)
{
    r2 y = x, a = { 0, 0 };
    real f = 0;
    int trigger = 0, type = 0, open = 0;
    size_t i;
    for(i = 0; i < m->nobjs; ++i)
    {
        const mobj* ob = &m->objs[i];
        const objdef* def = ob->def;
        if(def->type != trigger && def->type != T_FOREGROUND)
        {
            if(trigger == T_LENS)
            {
                y = deflect(y, a);
                open = 0;
            }
            trigger = def->type;
        }
        if(def->type != type)
        {
            /* COMPLHED opens the plane's block with "float2 a = 0;".  (A lens
             * following a foreground inside one plane would open a second,
             * unbalanced block in the reference text; here the plane simply
             * continues, as in the product.) */
            if(def->type == T_LENS && !open)
            {
                a = mk2(0, 0);
                open = 1;
            }
            type = def->type;
        }
        if(type == T_LENS)
            a = add2(a, def->defl(data + ob->d, def->off, y));
        else if(type == T_SOURCE)
            f += def->bright(data + ob->d, def->off, y);
        else
            f += def->bright(data + ob->d, def->off, x);
    }
    if(trigger == T_LENS)
        y = deflect(y, a);
    (void)y;
    return f;
}

/* generated set_params(), src/kernel.c:114-150 (templates) and :455-633:
 * image-plane-prior positions are shot through all lenses that precede the
 * start of the current source plane before being handed to the setter. */
static void set_params(const orc_model* m, const real* params, real* data)
{
    int trigger = 0;
    size_t plane = 0, i, k;
    int j;
    r2 x = { 0, 0 }, a = { 0, 0 };
    for(i = 0; i < m->nobjs; ++i)
    {
        const mobj* ob = &m->objs[i];
        const objdef* def = ob->def;
        real args[MAXPAR];
        if(def->type != trigger && def->type != T_FOREGROUND)
        {
            if(trigger == T_LENS && def->type == T_SOURCE)
                plane = i;
            trigger = def->type;
        }
        for(j = 0; j < def->npar; ++j)
        {
            if(ob->ipp[j] && def->par[j].type == P_POSITION_X)
            {
                int trigger2 = 0;
                x = mk2(params[ob->p + j], params[ob->p + j + 1]);
                for(k = 0; k < plane; ++k)
                {
                    const mobj* lk = &m->objs[k];
                    if(lk->def->type != trigger2 && lk->def->type != T_FOREGROUND)
                    {
                        if(trigger2 == T_LENS) { x = deflect(x, a); a = mk2(0, 0); }
                        trigger2 = lk->def->type;
                    }
                    if(lk->def->type == T_LENS)
                        a = add2(a, lk->def->defl(data + lk->d, lk->def->off, x));
                }
                if(trigger2 == T_LENS) { x = deflect(x, a); a = mk2(0, 0); }
            }
        }
        for(j = 0; j < def->npar; ++j)
        {
            if(ob->ipp[j])
                args[j] = def->par[j].type == P_POSITION_X ? x.x : def->par[j].type == P_POSITION_Y ? x.y : 0;
            else
                args[j] = params[ob->p + j];
        }
        def->set(data + ob->d, def->off, args);
    }
}

/* params[]: physical parameters in object order, already narrowed to float as
 * src/nested.c:71-72 does.  block_out (nullable): the object data block. */
int orc_set_params(orc_model* m, const float* params, real* block_out)
{
    real* p = malloc((m->npars ? m->npars : 1)*sizeof(real));
    size_t i;
    for(i = 0; i < m->npars; ++i) p[i] = params[i];
    memset(m->block, 0, m->words*sizeof(real));
    set_params(m, p, m->block);
    if(block_out) memcpy(block_out, m->block, m->words*sizeof(real));
    free(p);
    return 0;
}

/* render kernel, kernel/lensed.cl:9-38 */
static void render(const orc_model* m, const real* data, real* value, real* error)
{
    const size_t W = m->width, size = m->size, nq = m->nq;
    long k;
#pragma omp parallel for schedule(dynamic, 256)
    for(k = 0; k < (long)size; ++k)
    {
        r2 x = mk2(m->pcs[0] + m->pcs[2]*(real)(k % W), m->pcs[1] + m->pcs[3]*(real)(k / W));
        real f0 = 0, f1 = 0;
        size_t n;
        for(n = 0; n < nq; ++n)
        {
            real c = compute(m, data, add2(x, mk2(m->qq[2*n], m->qq[2*n+1])));
            f0 += m->ww[2*n]*c;
            f1 += m->ww[2*n+1]*c;
        }
        value[k] = f0;
        error[k] = f1;
    }
}

/* convolve kernel, kernel/lensed.cl:56-103: the local-memory cache is an
 * edge-clamped window of the input whose origin is (-Pw/2, -Ph/2) from the
 * work-group origin; entry (lj+Ph-1-j, li+Pw-1-i) is therefore input pixel
 * (gj - Ph/2 + Ph-1-j, gi - Pw/2 + Pw-1-i), clamped. */
static void convolve(const orc_model* m, const real* in, real* out)
{
    const long W = (long)m->width, H = (long)m->height;
    const long PW = (long)m->psfw, PH = (long)m->psfh;
    long gj;
#pragma omp parallel for schedule(static)
    for(gj = 0; gj < H; ++gj)
    {
        long gi, i, j;
        for(gi = 0; gi < W; ++gi)
        {
            real x = 0;
            for(j = 0; j < PH; ++j)
            {
                long row = gj - PH/2 + PH - j - 1;
                row = row < 0 ? 0 : row > H-1 ? H-1 : row;
                for(i = 0; i < PW; ++i)
                {
                    long col = gi - PW/2 + PW - i - 1;
                    col = col < 0 ? 0 : col > W-1 ? W-1 : col;
                    x += m->psf[j*PW + i]*in[row*W + col];
                }
            }
            out[gj*W + gi] = x;
        }
    }
}

/* value/error: kernel/lensed.cl:36-37 outputs (before convolution) */
int orc_render(orc_model* m, const float* params, real* value, real* error)
{
    orc_set_params(m, params, NULL);
    render(m, m->block, m->value, m->error);
    if(value) memcpy(value, m->value, m->size*sizeof(real));
    if(error) memcpy(error, m->error, m->size*sizeof(real));
    return 0;
}

int orc_convolve(orc_model* m, const real* in, real* out)
{
    if(!m->psf) return 1;
    convolve(m, in, out);
    return 0;
}

/* one likelihood evaluation, src/nested.c:63-115: set_params, render,
 * convolve (if PSF), loglike kernel (kernel/lensed.cl:41-53), then the host
 * sums the chi^2 map in double in ascending pixel order and returns -chi2/2.
 * model_out (nullable): what the loglike kernel compares with the image
 * (src/lensed.c:1073); chi_out (nullable): the per-pixel chi^2 map. */
int orc_loglike(orc_model* m, const float* params, double* lnew, real* model_out, real* chi_out)
{
    const real* model;
    double chi2 = 0.0;
    size_t k;
    orc_set_params(m, params, NULL);
    render(m, m->block, m->value, m->error);
    if(m->psf) { convolve(m, m->value, m->conv); model = m->conv; }
    else model = m->value;
    for(k = 0; k < m->size; ++k)
    {
        real d = model[k] - m->image[k];
        m->chi[k] = m->weight[k]*d*d;
    }
    for(k = 0; k < m->size; ++k)
        chi2 += m->chi[k];
    *lnew = -0.5*chi2;
    if(model_out) memcpy(model_out, model, m->size*sizeof(real));
    if(chi_out) memcpy(chi_out, m->chi, m->size*sizeof(real));
    return 0;
}

/* the dumper's six result layers, src/nested.c:178-253: re-render the given
 * point (set_params, render, convolve: :187-205), then on the host
 *   IMG = the model image (convolved if there is a PSF, :208)
 *   RES = image - IMG                                   (:223-224)
 *   RAW = the rendered value image                      (:240)
 *   ERR = error / value, the relative quadrature error  (:230-231)
 *   WHT = the weight map                                (:242)
 *   PVL = erfc(sqrt(0.5*loglike))                       (:237-238; a float times the
 *         double constant 0.5: evaluated in double and narrowed on assignment)
 * all as float arrays.  The reference's dumper does not re-run its loglike
 * kernel, so its PVL layer shows the chi^2 map of the point MultiNest evaluated
 * last; here the map is that of the point passed in (what the reference shows
 * when the last sampled point is the maximum-likelihood one). */
int orc_dumper_layers(orc_model* m, const float* params, float* img, float* res, float* raw, float* err, float* wht, float* pvl)
{
    double lnew;
    size_t k;
    const real* model;
    orc_loglike(m, params, &lnew, NULL, NULL);
    model = m->psf ? m->conv : m->value;
    for(k = 0; k < m->size; ++k)
    {
        img[k] = (float)model[k];
        res[k] = (float)(m->image[k] - model[k]);
        raw[k] = (float)m->value[k];
        err[k] = (float)(m->error[k]/m->value[k]);
        wht[k] = (float)m->weight[k];
        pvl[k] = (float)erfc(sqrt(0.5*(double)m->chi[k]));
    }
    return 0;
}

/* data preparation helpers used to build fixtures the way the reference's
 * host does: weight map src/data.c:314-330, PSF normalisation :354-370 */
void orc_make_weight(const float* image, const float* gain, double offset, size_t n, float* weight)
{
    size_t i;
    for(i = 0; i < n; ++i)
        weight[i] = (float)(gain[i]/(image[i] + offset));
}

void orc_normalise_psf(float* psf, size_t n)
{
    double norm = 0;
    size_t i;
    for(i = 0; i < n; ++i) norm += psf[i];
    for(i = 0; i < n; ++i) psf[i] = (float)(psf[i]/norm);
}
