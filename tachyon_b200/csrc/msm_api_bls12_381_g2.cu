// Entry points tachyon_bls12_381_g2_* (include/tachyon_msm_b200.h) and the kernels they instantiate.
#include "msm_api_common.cuh"

TB200_INSTANTIATE_GROUP(bls12_381, g2, Bls381G2Curve)
