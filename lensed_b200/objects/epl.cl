// Elliptical power-law lens, deflection by the series of Tessore & Metcalf
// (2015), truncated after the tenth order.
//
// Surface density slope t (t = 1 is isothermal), axis ratio q, scale r.  In
// elliptical polar coordinates (R, phi) of the rotated, r sqrt(q)-scaled frame
//
//   alpha' = 2 r sqrt(q)/(1 + q) R^(1-t) sum_k A_k,
//   A_0 = e^{i phi},  A_k = -f (2k - (2-t))/(2k + (2-t)) e^{2 i phi} A_{k-1},
//
// with f = (1 - q)/(1 + q) the second flattening.

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS },
        { "t", PARAMETER, { 0.f, 2.f } }, { "q", AXIS_RATIO }, { "pa", POS_ANGLE } };

// rotate, squash and scale into the elliptical frame; plain rotation back
data { float2 centre; mat22 to_lens; mat22 to_image; float slope; float flat; float amp; };

static float2 deflection(local data* this, float2 x)
{
    const float T = 2 - this->slope;
    const float f = this->flat;

    float2 v = mv22(this->to_lens, x - this->centre);

    float rad = length(v);
    float phi = atan2(v.y, v.x);

    float c1, c2;
    float s1 = sincos(phi, &c1);
    float s2 = sincos(2*phi, &c2);

    // multiplying by e^{2 i phi} is a rotation by 2 phi
    mat22 twist = (mat22)(c2, -s2, s2, c2);

    float2 term = (float2)(c1, s1);
    float2 sum = term;
    for(int k = 1; k <= 10; ++k)
    {
        term = -f*(2*k - T)/(2*k + T)*mv22(twist, term);
        sum += term;
    }

    sum *= this->amp*powr(rad, 1 - this->slope);

    return mv22(this->to_image, sum);
}

static void set(local data* this, float x, float y, float r, float t, float q, float pa)
{
    float cs;
    float sn = sincos(pa*DEG2RAD, &cs);

    this->centre   = (float2)(x, y);
    this->to_lens  = (1/r/sqrt(q))*(mat22)(q*cs, q*sn, -sn, cs);
    this->to_image = (mat22)(cs, -sn, sn, cs);
    this->slope    = t;
    this->flat     = (1 - q)/(1 + q);
    this->amp      = 2*r*sqrt(q)/(1 + q);
}
