// Singular isothermal ellipsoid plus external shear.
//
// SIE deflection as in sie.cl, plus  G u  with  G = (g1 g2; g2 -g1)  acting on
// the offset u from the lens centre.

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "q", AXIS_RATIO },
        { "pa", POS_ANGLE }, { "g1", PARAMETER }, { "g2", PARAMETER } };

data { float2 centre; mat22 to_lens; mat22 to_image; mat22 shear; float q_sq; float ecc; float amp; };

static float2 deflection(local data* this, float2 x)
{
    float2 u = x - this->centre;
    float2 v = mv22(this->to_lens, u);
    float  s = this->ecc/sqrt(this->q_sq*v.x*v.x + v.y*v.y);
    float2 a = this->amp*(float2)(atan(v.x*s), atanh(v.y*s));
    return mv22(this->to_image, a) + mv22(this->shear, u);
}

static void set(local data* this, float x, float y, float r, float q, float pa, float g1, float g2)
{
    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);

    this->centre   = (float2)(x, y);
    this->to_lens  = (mat22)(cs, sn, -sn, cs);
    this->to_image = (mat22)(cs, -sn, sn, cs);
    this->shear    = (mat22)(g1, g2, g2, -g1);
    this->q_sq     = q*q;
    this->ecc      = sqrt(1 - q*q);
    this->amp      = r*sqrt(q)/sqrt(1 - q*q);
}
