// mgs_kernel_wide.cu - the environment-per-CTA variant of the rollout kernel (256 threads share one environment); see
// mgs_common.cuh (MGS_WIDE) and mgs_kernel_ops.h
#include <cuda_runtime.h>
#define MGS_WIDE 256
#define MGS_MAX_WARPS_PER_BLOCK 8
#define MGS_KERNEL_TAG wide
#include "mgs_kernel.cuh"
