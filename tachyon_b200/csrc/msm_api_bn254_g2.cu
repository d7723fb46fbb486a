// Entry points tachyon_bn254_g2_* (include/tachyon_msm_b200.h) and the kernels they instantiate.
#include "msm_api_common.cuh"

TB200_INSTANTIATE_GROUP(bn254, g2, Bn254G2Curve)
