// The one definition of the non-template kernels of the two-level sort (msm_sort.cuh).
#define TB200_DEFINE_SORT_KERNELS
#include "msm_sort.cuh"
