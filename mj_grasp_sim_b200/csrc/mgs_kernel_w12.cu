// mgs_kernel_w12.cu - the 12-warp (168-register) variant of the rollout kernel; see mgs_kernel_ops.h
#include <cuda_runtime.h>
#define MGS_MAX_WARPS_PER_BLOCK 12
#define MGS_KERNEL_TAG w12
#include "mgs_kernel.cuh"
