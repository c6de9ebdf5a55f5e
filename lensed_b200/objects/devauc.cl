// De Vaucouleurs r^(1/4) profile (Sersic n = 4 with fixed b):
//
//   I(x) = I0 exp(-b (R/r)^(1/4)),   R = |(q u1, u2)|  in the rotated frame,
//   total flux 10^(-0.4 mag) = I0 pi r^2 q 8!/b^8.

#define DEVAUC_B 7.6692494425008039044f
// b^8/8!
#define DEVAUC_C 296.826303766893f

type = SOURCE;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "mag", MAGNITUDE },
        { "q", AXIS_RATIO }, { "pa", POS_ANGLE } };

// rotate by pa, squash first axis by q
data { float2 centre; mat22 to_profile; float scale; float peak; };

static float brightness(local data* this, float2 x)
{
    float2 v = mv22(this->to_profile, x - this->centre);
    return this->peak*exp(-DEVAUC_B*sqrt(sqrt(length(v)/this->scale)));
}

static void set(local data* this, float x, float y, float r, float mag, float q, float pa)
{
    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);

    this->centre     = (float2)(x, y);
    this->to_profile = (mat22)(q*cs, q*sn, -sn, cs);
    this->scale      = r;
    // total flux of the r^(1/4) law: pi r^2 q I0 8!/b^8
    float flux = exp(-0.4f*mag*LOG_10);
    this->peak       = flux/PI/r/r/q*DEVAUC_C;
}
