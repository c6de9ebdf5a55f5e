// Singular isothermal ellipsoid (Kormann, Schneider & Bartelmann 1994).
//
// In the frame rotated by the position angle, with axis ratio q,
// e = sqrt(1 - q^2) and psi = sqrt(q^2 u1^2 + u2^2):
//
//   alpha' = r sqrt(q)/e (atan(e u1/psi), atanh(e u2/psi))
//
// which is rotated back to the image frame.

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "q", AXIS_RATIO },
        { "pa", POS_ANGLE } };

// rotation into the lens frame; and back; e; r sqrt(q)/e
data { float2 centre; mat22 to_lens; mat22 to_image; float q_sq; float ecc; float amp; };

static float2 deflection(local data* this, float2 x)
{
    float2 u = x - this->centre;
    float2 v = mv22(this->to_lens, u);
    float  s = this->ecc/sqrt(this->q_sq*v.x*v.x + v.y*v.y);
    float2 a = this->amp*(float2)(atan(v.x*s), atanh(v.y*s));
    return mv22(this->to_image, a);
}

static void set(local data* this, float x, float y, float r, float q, float pa)
{
    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);

    this->centre   = (float2)(x, y);
    this->to_lens  = (mat22)(cs, sn, -sn, cs);
    this->to_image = (mat22)(cs, -sn, sn, cs);
    this->q_sq     = q*q;
    this->ecc      = sqrt(1 - q*q);
    this->amp      = r*sqrt(q)/sqrt(1 - q*q);
}
