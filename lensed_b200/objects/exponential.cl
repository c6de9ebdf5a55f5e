// Exponential disc:
//
//   I(x) = I0 exp(-R/rs),   R = |(q u1, u2)|  in the rotated frame,
//   total flux 10^(-0.4 mag) = I0 2 pi rs^2 q.

type = SOURCE;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "rs", RADIUS }, { "mag", MAGNITUDE },
        { "q", AXIS_RATIO }, { "pa", POS_ANGLE } };

data { float2 centre; mat22 to_profile; float scale; float peak; };

static float brightness(local data* this, float2 x)
{
    float2 v = mv22(this->to_profile, x - this->centre);
    return this->peak*exp(-length(v)/this->scale);
}

static void set(local data* this, float x, float y, float rs, float mag, float q, float pa)
{
    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);

    this->centre     = (float2)(x, y);
    this->to_profile = (mat22)(q*cs, q*sn, -sn, cs);
    this->scale      = rs;
    // total flux of an elliptical exponential disc: 2 pi rs^2 q I0
    float flux = exp(-0.4f*mag*LOG_10);
    this->peak       = flux*0.5f/PI/rs/rs/q;
}
