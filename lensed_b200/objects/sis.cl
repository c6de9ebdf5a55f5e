// Singular isothermal sphere.
//
// Deflection of constant magnitude r (the Einstein radius) pointing away
// from the lens centre:  alpha(x) = r (x - c)/|x - c|.

type = LENS;

params
{
    { "x", POSITION_X },
    { "y", POSITION_Y },
    { "r", RADIUS     }
};

data
{
    float2 centre;
    float  einstein;
};

static float2 deflection(local data* this, float2 x)
{
    return this->einstein*normalize(x - this->centre);
}

static void set(local data* this, float x, float y, float r)
{
    this->centre   = (float2)(x, y);
    this->einstein = r;
}
