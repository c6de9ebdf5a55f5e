// Singular isothermal sphere in an external shear field.
//
//   alpha(x) = r u/|u| + G u,   u = x - c,   G = (g1 g2; g2 -g1)

type = LENS;

params { { "x", POSITION_X }, { "y", POSITION_Y }, { "r", RADIUS }, { "g1", PARAMETER },
        { "g2", PARAMETER } };

data { float2 centre; mat22 shear; float einstein; };

// The shear matrix acts on the offset from the lens centre, not on the image
// position: the external field vanishes at the centre of the sphere.
static float2 deflection(local data* this, float2 x)
{
    float2 u = x - this->centre;
    return this->einstein*normalize(u) + mv22(this->shear, u);
}

static void set(local data* this, float x, float y, float r, float g1, float g2)
{
    this->centre   = (float2)(x, y);
    this->einstein = r;
    this->shear    = (mat22)(g1, g2, g2, -g1);
}
