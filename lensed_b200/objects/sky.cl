// Sky background: a plane  bg + dx (x - 1) + dy (y - 1)  in image pixel
// coordinates (first pixel centre = (1, 1)).  The gradients default to zero;
// the default is written as -0.0f because a positive value or a set sign bit
// marks "has a default" for the host (plain 0 means "no default").

type = FOREGROUND;

params { { "bg" }, { "dx", PARAMETER, UNBOUNDED, -0.0f }, { "dy", PARAMETER, UNBOUNDED, -0.0f } };

data { float level; float2 slope; };

static float foreground(local data* this, float2 x)
{
    return this->level + dot(this->slope, x - (float2)(1, 1));
}

static void set(local data* this, float bg, float dx, float dy)
{
    this->level = bg;
    this->slope = (float2)(dx, dy);
}
