// Elliptical Gaussian:
//
//   I(x) = I0 exp(-R^2/(2 sigma^2)),   R = |(q u1, u2)|  in the rotated frame,
//   total flux 10^(-0.4 mag) = I0 2 pi sigma^2 q.

type = SOURCE;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "sigma", RADIUS }, { "mag", MAGNITUDE },
        { "q", AXIS_RATIO }, { "pa", POS_ANGLE } };

data { float2 centre; mat22 to_profile; float variance; float peak; };

static float brightness(local data* this, float2 x)
{
    float2 v = mv22(this->to_profile, x - this->centre);
    return this->peak*exp(-0.5f*dot(v, v)/this->variance);
}

// to_profile rotates by the position angle (counter-clockwise from the x axis,
// degrees) and squashes the first axis by q, so that R^2 = dot(v, v); the peak
// follows from the total flux of an elliptical Gaussian, 2 pi sigma^2 q I0.
static void set(local data* this, float x, float y, float sigma, float mag, float q, float pa)
{
    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);
    float flux = exp(-0.4f*mag*LOG_10);

    this->centre     = (float2)(x, y);
    this->to_profile = (mat22)(q*cs, q*sn, -sn, cs);
    this->variance   = sigma*sigma;
    this->peak       = flux*0.5f/PI/this->variance/q;
}
