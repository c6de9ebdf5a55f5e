// Singular isothermal sphere.
//
// Deflection of constant magnitude r (the Einstein radius) pointing away
// from the lens centre:  alpha(x) = r (x - c)/|x - c|.

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS } };

data { float2 centre; float einstein; };

// The convergence kappa = r/(2|u|) gives a deflection whose length does not
// depend on the distance from the centre; at u = 0 the direction is undefined
// (0/0): the generated compute() sends such a ray to (1e10, 1e10).
static float2 deflection(local data* this, float2 x)
{
    return this->einstein*normalize(x - this->centre);
}

// nothing to precompute: centre and Einstein radius as given (pixels)
static void set(local data* this, float x, float y, float r)
{
    this->centre   = (float2)(x, y);
    this->einstein = r;
}
