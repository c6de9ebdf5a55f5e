// object.cuh -- the plugin ABI every objects/<name>.cl file is written against.
//
// Same names, values and layout as the reference's kernel/object.cl:2-33
// (object / parameter type enums, bound macros, `struct param`) and
// kernel/constants.cl:6-35 (numeric constants, IMAGE_CENTER, mat22, mv22), so
// that existing object files compile unmodified.  IMAGE_WIDTH / IMAGE_HEIGHT
// come from the -D options the host passes (src/kernel.c:890-896).

// object types
enum { LENS = 'L', SOURCE = 'S', FOREGROUND = 'F' };

// parameter types
enum { PARAMETER = 0, POSITION_X, POSITION_Y, RADIUS, MAGNITUDE, AXIS_RATIO, POS_ANGLE };

// parameter bounds
#define UNBOUNDED {0, 0}
#define POS_BOUND {0, +FLT_MAX}
#define NEG_BOUND {-FLT_MAX, 0}

// one entry of an object's `params` list; 32 bytes, read back by the host
// from the compiled module (lcu_program.cpp: read_object_meta)
struct alignas(4) param
{
    char  name[16];
    int   type;
    float bounds[2];
    float defval;
};

// trigonometry
#define PI      3.1415926535897932384626433832795028841971693993751f
#define PI_HALF 1.5707963267948966192313216916397514420985846996876f
#define DEG2RAD 0.0174532925199432957692369076848861271344287188854f

// logarithms
#define LOG_10  2.3025850929940456840179914546843642076011014886288f
#define LOG_PI  1.1447298858494001741434273513530587116472948129153f
#define LOG_2PI 1.8378770664093454835606594728112352797227949472756f

// true centre of the image in pixel coordinates
#define IMAGE_CENTER (lcu_float2(0.5f*(IMAGE_WIDTH + 1), 0.5f*(IMAGE_HEIGHT + 1)))

// 2x2 matrices are row-major float4: (m.x m.y; m.z m.w)
typedef lcu_float4 mat22;

// matrix-vector product, rows dotted with v (kernel/constants.cl:32-35)
__device__ __forceinline__ lcu_float2 mv22(mat22 m, lcu_float2 v)
{
    return lcu_float2(dot(m.lo, v), dot(m.hi, v));
}

// the same for two rays at once (shim.cuh: packed pairs); the matrix is either
// a field of the object's data block (uniform) or a local of the plugin
__device__ __forceinline__ lcu_pf2 mv22(lcu_pf4 m, lcu_pf2 v)
{
    return lcu_pf2(dot(m.lo, v), dot(m.hi, v));
}
__device__ __forceinline__ lcu_pf2 mv22(lcu_float4 m, lcu_pf2 v)
{
    return lcu_pf2(lcu_pf(m.x)*v.x + lcu_pf(m.y)*v.y, lcu_pf(m.z)*v.x + lcu_pf(m.w)*v.y);
}
