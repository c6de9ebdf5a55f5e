// lcu_quad.cpp -- sub-pixel quadrature rules.
//
// Replaces src/quadrature.c:21-43 and the tables under src/quad/.  The
// reference stores every rule as a full list of 2-D points; here the
// Cartesian Gauss-Kronrod rules are generated from their 1-D factors
// (quad_tables.inc, derived from first principles by
// tools/gen_quad_tables.py) and the Genz-Malik rule from its closed form.
// Point order, weights and error weights equal the reference's after the
// narrowing to float that quad_rule() performs (tests/test_quadrature.py).

#include "lcu_internal.h"

#include <cmath>
#include <cstring>

namespace lcu {

#include "quad_tables.inc"

namespace {

struct Rule
{
    const char* name;
    const char* info;
    int size;
    void (*build)(const Rule&, double (*pts)[2], double* wht, double* err);
    int m;                          // points per axis for Cartesian rules
    const double *x, *k, *g;        // 1-D abscissae, Kronrod and Gauss weights
};

// m x m midpoint sub-sampling, equal weights, no error estimate
void build_grid(const Rule& r, double (*pts)[2], double* wht, double* err)
{
    const int m = r.m;
    for(int i = 0; i < m; ++i)
        for(int j = 0; j < m; ++j)
        {
            pts[i*m + j][0] = (i + 0.5)/m - 0.5;
            pts[i*m + j][1] = (j + 0.5)/m - 0.5;
            wht[i*m + j] = 1.0/(m*m);
            err[i*m + j] = 0.0;
        }
}

// tensor product of a 1-D Gauss-Kronrod pair: first axis outer, second inner;
// error weight = Kronrod-minus-Gauss along both axes
void build_kronrod(const Rule& r, double (*pts)[2], double* wht, double* err)
{
    const int m = r.m;
    for(int i = 0; i < m; ++i)
        for(int j = 0; j < m; ++j)
        {
            pts[i*m + j][0] = r.x[i];
            pts[i*m + j][1] = r.x[j];
            wht[i*m + j] = r.k[i]*r.k[j];
            err[i*m + j] = (r.k[i] - r.g[i])*(r.k[j] - r.g[j]);
        }
}

// Genz & Malik (1980) degree-7 rule with embedded degree-5 rule in two
// dimensions on the unit square: centre, two axial generators, two diagonal
// generators; error weight = degree-7 minus degree-5 weight
void build_genz_malik(const Rule&, double (*pts)[2], double* wht, double* err)
{
    const double n = 2.0, vol = 4.0;
    const double l2 = 0.5*std::sqrt(9.0/70.0), l3 = 0.5*std::sqrt(9.0/10.0), l5 = 0.5*std::sqrt(9.0/19.0);
    const double w7[5] = { (12824.0 - 9120.0*n + 400.0*n*n)/19683.0, 980.0/6561.0,
                           (1820.0 - 400.0*n)/19683.0, 200.0/19683.0, 6859.0/19683.0/vol };
    const double w5[5] = { (729.0 - 950.0*n + 50.0*n*n)/729.0, 245.0/486.0,
                           (265.0 - 100.0*n)/1458.0, 25.0/729.0, 0.0 };
    int k = 0;
    auto put = [&](double px, double py, int cls) {
        pts[k][0] = px; pts[k][1] = py; wht[k] = w7[cls]; err[k] = w7[cls] - w5[cls]; ++k;
    };
    put(0, 0, 0);
    const double axial[2] = { l2, l3 };
    for(int s = 0; s < 2; ++s)
    {
        put(-axial[s], 0, 1 + s);
        put(0, -axial[s], 1 + s);
        put(0, +axial[s], 1 + s);
        put(+axial[s], 0, 1 + s);
    }
    const double diag[2] = { l3, l5 };
    for(int s = 0; s < 2; ++s)
        for(int a = -1; a <= 1; a += 2)
            for(int b = -1; b <= 1; b += 2)
                put(a*diag[s], b*diag[s], 3 + s);
}

void build_point(const Rule&, double (*pts)[2], double* wht, double* err)
{
    pts[0][0] = pts[0][1] = 0.0;
    wht[0] = 1.0;
    err[0] = 0.0;
}

const Rule RULES[] = {
    { "point", "single point sampling without error estimate", 1, build_point, 1, nullptr, nullptr, nullptr },
    { "sub2", "2x2 subsampling without error estimate", 4, build_grid, 2, nullptr, nullptr, nullptr },
    { "sub4", "4x4 subsampling without error estimate", 16, build_grid, 4, nullptr, nullptr, nullptr },
    { "gm75", "Genz-Malik (7, 5) fully symmetric rule", 17, build_genz_malik, 0, nullptr, nullptr, nullptr },
    { "g3k7", "Gauss-Kronrod (7, 3) Cartesian rule", 49, build_kronrod, 7, GK_G3K7_X, GK_G3K7_K, GK_G3K7_G },
    { "g5k11", "Gauss-Kronrod (11, 5) Cartesian rule", 121, build_kronrod, 11, GK_G5K11_X, GK_G5K11_K, GK_G5K11_G },
    { "g7k15", "Gauss-Kronrod (15, 7) Cartesian rule", 225, build_kronrod, 15, GK_G7K15_X, GK_G7K15_K, GK_G7K15_G },
};
const int NRULES = (int)(sizeof(RULES)/sizeof(RULES[0]));

} // namespace

int quad_rule_count() { return NRULES; }
const char* quad_rule_name(int i) { return i >= 0 && i < NRULES ? RULES[i].name : nullptr; }
const char* quad_rule_info(int i) { return i >= 0 && i < NRULES ? RULES[i].info : nullptr; }

// quad_rule(), src/quadrature.c:32-43: scale abscissae by the pixel scale in
// double, then narrow everything to float
int quad_rule(const char* rule, double sx, double sy, float* qq, float* ww)
{
    if(!rule)
        return -1;
    for(int i = 0; i < NRULES; ++i)
    {
        if(std::strcmp(RULES[i].name, rule) != 0)
            continue;
        const Rule& r = RULES[i];
        if(qq && ww)
        {
            double pts[225][2], wht[225], err[225];
            r.build(r, pts, wht, err);
            for(int n = 0; n < r.size; ++n)
            {
                qq[2*n + 0] = (float)(sx*pts[n][0]);
                qq[2*n + 1] = (float)(sy*pts[n][1]);
                ww[2*n + 0] = (float)wht[n];
                ww[2*n + 1] = (float)err[n];
            }
        }
        return r.size;
    }
    return -1;
}

} // namespace lcu
