// Isothermal sphere with a finite core (non-singular).
//
//   alpha(x) = r u/(rc + |u|),   u = x - c

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "rc", RADIUS } };

data { float2 centre; float einstein; float core; };

// Softened version of sis.cl: the core radius rc removes the singular centre
// (alpha -> r u/rc for |u| << rc) and leaves alpha -> r u/|u| far outside it.
static float2 deflection(local data* this, float2 x)
{
    float2 u = x - this->centre;
    return this->einstein/(this->core + length(u))*u;
}

// both radii in pixels; rc = 0 reproduces the singular sphere
static void set(local data* this, float x, float y, float r, float rc)
{
    this->centre   = (float2)(x, y);
    this->einstein = r;
    this->core     = rc;
}
