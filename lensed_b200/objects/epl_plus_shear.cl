// Elliptical power-law lens (Tessore & Metcalf 2015 series, see epl.cl) plus
// an external shear  G u,  G = (g1 g2; g2 -g1),  u the offset from the centre.

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS },
        { "t", PARAMETER, { 0.f, 2.f } }, { "q", AXIS_RATIO }, { "pa", POS_ANGLE },
        { "g1", PARAMETER }, { "g2", PARAMETER } };

data { float2 centre; mat22 to_lens; mat22 to_image; mat22 shear; float slope; float flat; float amp; };

static float2 deflection(local data* this, float2 x)
{
    const float T = 2 - this->slope;
    const float f = this->flat;

    float2 u = x - this->centre;
    float2 v = mv22(this->to_lens, u);

    float rad = length(v);
    float phi = atan2(v.y, v.x);

    float c1, c2;
    float s1 = sincos(phi, &c1);
    float s2 = sincos(2*phi, &c2);

    mat22 twist = (mat22)(c2, -s2, s2, c2);

    float2 term = (float2)(c1, s1);
    float2 sum = term;
    for(int k = 1; k <= 10; ++k)
    {
        term = -f*(2*k - T)/(2*k + T)*mv22(twist, term);
        sum += term;
    }

    sum *= this->amp*powr(rad, 1 - this->slope);

    float2 a = mv22(this->to_image, sum);
    return a + mv22(this->shear, u);
}

static void set(local data* this, float x, float y, float r, float t, float q, float pa, float g1, float g2)
{
    float cs;
    float sn = sincos(pa*DEG2RAD, &cs);

    this->centre   = (float2)(x, y);
    this->to_lens  = (1/r/sqrt(q))*(mat22)(q*cs, q*sn, -sn, cs);
    this->to_image = (mat22)(cs, -sn, sn, cs);
    this->slope    = t;
    this->flat     = (1 - q)/(1 + q);
    this->amp      = 2*r*sqrt(q)/(1 + q);
    this->shear    = (mat22)(g1, g2, g2, -g1);
}
