// Isothermal ellipsoid with a finite core rc (non-singular).
//
// With psi = sqrt(q^2 u1^2 + u2^2) in the lens frame:
//
//   alpha' = r sqrt(q)/e (atan(e u1/(rc + psi)), atanh(e u2/(q^2 rc + psi)))

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "rc", RADIUS },
        { "q", AXIS_RATIO }, { "pa", POS_ANGLE } };

data { float2 centre; mat22 to_lens; mat22 to_image; float core; float q_sq; float ecc; float amp; };

static float2 deflection(local data* this, float2 x)
{
    float2 u = x - this->centre;
    float2 v = mv22(this->to_lens, u);
    float  psi = sqrt(this->q_sq*v.x*v.x + v.y*v.y);
    float2 a = this->amp*(float2)(atan(v.x*this->ecc/(this->core + psi)),
                                  atanh(v.y*this->ecc/(this->core*this->q_sq + psi)));
    return mv22(this->to_image, a);
}

static void set(local data* this, float x, float y, float r, float rc, float q, float pa)
{
    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);

    this->centre   = (float2)(x, y);
    this->to_lens  = (mat22)(cs, sn, -sn, cs);
    this->to_image = (mat22)(cs, -sn, sn, cs);
    this->core     = rc;
    this->q_sq     = q*q;
    this->ecc      = sqrt(1 - q*q);
    this->amp      = r*sqrt(q)/sqrt(1 - q*q);
}
