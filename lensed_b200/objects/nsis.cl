// Isothermal sphere with a finite core (non-singular).
//
//   alpha(x) = r u/(rc + |u|),   u = x - c

type = LENS;

params
{
    { "x",  POSITION_X },
    { "y",  POSITION_Y },
    { "r",  RADIUS     },
    { "rc", RADIUS     }
};

data
{
    float2 centre;
    float  einstein;
    float  core;
};

static float2 deflection(local data* this, float2 x)
{
    float2 u = x - this->centre;
    return this->einstein/(this->core + length(u))*u;
}

static void set(local data* this, float x, float y, float r, float rc)
{
    this->centre   = (float2)(x, y);
    this->einstein = r;
    this->core     = rc;
}
