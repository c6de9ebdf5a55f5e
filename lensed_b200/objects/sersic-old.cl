// Sersic profile with the linear b(n) = 1.9992 n - 0.3271 approximation
// (Capaccioli 1989), usable for 0.5 < n < 8.  Kept for old configurations;
// prefer sersic.cl.

type = SOURCE;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "mag", MAGNITUDE },
        { "n", PARAMETER, { 0.5f, 8.0f } }, { "q", AXIS_RATIO }, { "pa", POS_ANGLE } };

data { float2 centre; mat22 to_profile; float log0; float log1; float half_inv_n; };

static float brightness(local data* this, float2 x)
{
    float2 v = mv22(this->to_profile, x - this->centre);
    return exp(this->log0 - exp(this->log1 + this->half_inv_n*log(dot(v, v))));
}

static void set(local data* this, float x, float y, float r, float mag, float n, float q, float pa)
{
    float b = 1.9992f*n - 0.3271f;

    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);

    this->centre     = (float2)(x, y);
    this->to_profile = (mat22)(q*cs, q*sn, -sn, cs)/sqrt(q);

    this->log0 = -0.4f*mag*LOG_10 + 2*n*log(b) - LOG_PI - 2*log(r) - log(tgamma(2*n+1));
    this->log1 = log(b) - log(r)/n;
    this->half_inv_n = 0.5f/n;
}
