"""lensed_b200 -- B200-native (sm_100a) replacement for the per-likelihood
model-image hot path of Lensed: set_params -> render -> convolve -> loglike ->
reduce, behind Lensed's own objects/*.cl plugin API.  See DESIGN.md."""
from .api import (Context, Model, ObjectInfo, Param, LensedCudaError, quad_rule, quad_rules,
                  launch_count, LENS, SOURCE, FOREGROUND)
from ._lib import LCU_FAST_MATH, LCU_OBJ_SHARED, LCU_FAST_INTRINSICS, LCU_NO_PAIR, LCU_FAST_LENS_INTRINSICS, LCU_FAST_ATANH, LCU_SOURCE_ONLY

__all__ = ["Context", "Model", "ObjectInfo", "Param", "LensedCudaError", "quad_rule", "quad_rules",
           "launch_count", "LENS", "SOURCE", "FOREGROUND", "LCU_FAST_MATH", "LCU_OBJ_SHARED", "LCU_FAST_INTRINSICS", "LCU_NO_PAIR", "LCU_FAST_LENS_INTRINSICS", "LCU_FAST_ATANH", "LCU_SOURCE_ONLY"]
