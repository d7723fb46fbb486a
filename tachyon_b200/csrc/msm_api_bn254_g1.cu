// Entry points tachyon_bn254_g1_* (include/tachyon_msm_b200.h) and the kernels they instantiate.
#include "msm_api_common.cuh"

TB200_INSTANTIATE_GROUP(bn254, g1, Bn254Curve)
