// Sersic profile, normalised to total magnitude `mag`:
//
//   I(x) = I0 exp(-b (R/r)^(1/n)),   R^2 = q u1^2 + u2^2/q  (rotated frame)
//
// evaluated in log space as exp(log0 - exp(log1 + log(R^2)/(2n))).  b(n) is
// the asymptotic expansion of Ciotti & Bertin (1999) with the higher-order
// terms and the small-n polynomial of MacArthur, Courteau & Holtzman (2003).

type = SOURCE;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "mag", MAGNITUDE },
        { "n", PARAMETER, POS_BOUND }, { "q", AXIS_RATIO }, { "pa", POS_ANGLE } };

// rotate by pa, scale axes by sqrt(q), 1/sqrt(q); log I0; log b - log(r)/n; 1/(2n)
data { float2 centre; mat22 to_profile; float log0; float log1; float half_inv_n; };

static float brightness(local data* this, float2 x)
{
    float2 v = mv22(this->to_profile, x - this->centre);
    return exp(this->log0 - exp(this->log1 + this->half_inv_n*log(dot(v, v))));
}

static void set(local data* this, float x, float y, float r, float mag, float n, float q, float pa)
{
    float b = n > 0.36f
            ? 2.0f*n - 1.0f/3 + 4.0f/(405*n) + 46.0f/(25515*(n*n)) + 131.0f/(1148175*(n*n*n)) - 2194697.0f/(30690717750*(n*n*n*n))
            : 0.01945f - 0.8902f*n + 10.95f*(n*n) - 19.67f*(n*n*n) + 13.43f*(n*n*n*n);

    float cs = cos(pa*DEG2RAD);
    float sn = sin(pa*DEG2RAD);

    this->centre     = (float2)(x, y);
    this->to_profile = (mat22)(q*cs, q*sn, -sn, cs)/sqrt(q);

    // total flux 10^(-0.4 mag) = I0 2 pi n r^2 Gamma(2n) b^(-2n)
    this->log0 = -0.4f*mag*LOG_10 + 2*n*log(b) - LOG_PI - 2*log(r) - log(tgamma(2*n+1));
    this->log1 = log(b) - log(r)/n;
    this->half_inv_n = 0.5f/n;
}
