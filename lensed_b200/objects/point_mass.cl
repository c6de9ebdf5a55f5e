// Point-mass (Schwarzschild) lens.
//
//   alpha(x) = r^2 u/|u|^2,   u = x - c

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS } };

data { float2 centre; float einstein_sq; };

static float2 deflection(local data* this, float2 x)
{
    float2 u = x - this->centre;
    return this->einstein_sq/dot(u, u)*u;
}

static void set(local data* this, float x, float y, float r)
{
    this->centre      = (float2)(x, y);
    this->einstein_sq = r*r;
}
