// mgs_sampler.cu - the ray casting of the antipodal grasp sampler as a CUDA kernel.
//
// Reference: /root/reference/mgs/sampler/antipodal.py:115-151 - for every sampled surface point one trimesh ray query in the +dir
// and one in the -dir direction, hits closer than eps dropped, ONE of the remaining hits chosen at random.  Here: one warp per
// point, the triangles staged through shared memory in tiles shared by the 8 warps (= 8 points) of a CTA, Moeller-Trumbore in
// fp64 (the reference works in float64), two sweeps: count the valid hits, then pick hit number floor(u * count) in the order
// "+dir faces 0..F-1, then -dir faces 0..F-1" (the order of the batched host expression in mgs/sampler/antipodal.py).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

#include "../../include/mgs_b200.h"

#define SMP_TILE 128  // faces per shared-memory tile: 128 x 9 doubles = 9 KB
#define SMP_WARPS 8

__device__ __forceinline__ bool ray_tri(const double *T, double ox, double oy, double oz, double dx, double dy, double dz, double *t_out) {
  const double e1x = T[3] - T[0], e1y = T[4] - T[1], e1z = T[5] - T[2];
  const double e2x = T[6] - T[0], e2y = T[7] - T[1], e2z = T[8] - T[2];
  const double px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
  const double det = e1x * px + e1y * py + e1z * pz;
  if (!(fabs(det) > 1e-14)) return false;
  const double inv = 1.0 / det;
  const double tx = ox - T[0], ty = oy - T[1], tz = oz - T[2];
  const double u = (tx * px + ty * py + tz * pz) * inv;
  const double qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
  const double v = (dx * qx + dy * qy + dz * qz) * inv;
  const double t = (e2x * qx + e2y * qy + e2z * qz) * inv;
  *t_out = t;
  return u >= -1e-12 && v >= -1e-12 && u + v <= 1.0 + 1e-12 && t > 0;
}

// sweep 0 counts the valid hits; sweep 1 finds valid hit number `target` (order: +dir faces, then -dir faces)
__global__ void __launch_bounds__(SMP_WARPS * 32) mgs_antipodal_kernel(int n, const double *__restrict__ p1, const double *__restrict__ dirs, int nface,
                                                                          const double *__restrict__ tri, double eps, const double *__restrict__ pick_u,
                                                                          double *__restrict__ signed_t, int *__restrict__ nvalid_out) {
  __shared__ double tile[SMP_TILE * 9];
  const int lane = threadIdx.x & 31, pt = blockIdx.x * SMP_WARPS + (threadIdx.x >> 5);
  const bool live = pt < n;
  double o[3] = {0, 0, 0}, d[3] = {1, 0, 0};
  if (live)
    for (int k = 0; k < 3; k++) { o[k] = p1[3 * pt + k]; d[k] = dirs[3 * pt + k]; }
  int count = 0, target = 0, found = 0;
  double answer = nan("");
  for (int sweep = 0; sweep < 2; sweep++) {
    int seen = 0;
    for (int sgn = 0; sgn < 2; sgn++) {
      const double s = sgn ? -1.0 : 1.0;
      for (int f0 = 0; f0 < nface; f0 += SMP_TILE) {
        const int nt = min(SMP_TILE, nface - f0);
        __syncthreads();
        for (int k = threadIdx.x; k < nt * 9; k += blockDim.x) tile[k] = tri[(size_t)f0 * 9 + k];
        __syncthreads();
        for (int g = 0; g < nt; g += 32) {
          const int f = g + lane;
          double t = 0;
          const bool ok = live && f < nt && ray_tri(tile + 9 * f, o[0], o[1], o[2], s * d[0], s * d[1], s * d[2], &t) && t >= eps;
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (sweep == 1 && !found && seen + __popc(m) > target) {
            // the lane whose rank inside this group equals target - seen holds the answer
            const int want = target - seen;
            const bool mine = ok && __popc(m & ((1u << lane) - 1u)) == want;
            const unsigned who = __ballot_sync(0xffffffffu, mine);
            const double tt = __shfl_sync(0xffffffffu, t, __ffs(who) - 1);
            answer = s * tt;
            found = 1;
          }
          seen += __popc(m);
        }
      }
    }
    if (sweep == 0) {
      count = seen;
      const double u = live ? pick_u[pt] : 0.0;
      int pick = (int)(u * (double)(count > 1 ? count : 1));
      const int last = count > 0 ? count - 1 : 0;
      target = pick < last ? pick : last;
      if (count == 0) found = 1;  // nothing to look for (every warp still walks the tiles: they are loaded by the whole CTA)
    }
  }
  if (live && lane == 0) { signed_t[pt] = answer; nvalid_out[pt] = count; }
}

extern "C" const char *mgs_last_error(void);
int mgs_set_error(const char *msg);  // (mgs_b200.cu)

extern "C" int mgs_antipodal_hits(int device, int n, const double *p1, const double *dirs, int nface, const double *tri, double eps,
                                  const double *pick_u, double *signed_t_out, int *nvalid_out) {
  if (n <= 0) return 0;
  if (!p1 || !dirs || !tri || !pick_u || !signed_t_out || !nvalid_out || nface <= 0) return mgs_set_error("mgs_antipodal_hits: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return mgs_set_error("no CUDA device: libmgs_b200 has no CPU path");
  if (device < 0 || device >= ndev) return mgs_set_error("bad device index");
  cudaError_t rc = cudaSetDevice(device);
  double *d_p1 = nullptr, *d_dir = nullptr, *d_tri = nullptr, *d_u = nullptr, *d_t = nullptr;
  int *d_n = nullptr;
  const size_t b3 = (size_t)n * 3 * sizeof(double), bt = (size_t)nface * 9 * sizeof(double);
  if (rc == cudaSuccess) rc = cudaMalloc(&d_p1, b3);
  if (rc == cudaSuccess) rc = cudaMalloc(&d_dir, b3);
  if (rc == cudaSuccess) rc = cudaMalloc(&d_tri, bt);
  if (rc == cudaSuccess) rc = cudaMalloc(&d_u, (size_t)n * sizeof(double));
  if (rc == cudaSuccess) rc = cudaMalloc(&d_t, (size_t)n * sizeof(double));
  if (rc == cudaSuccess) rc = cudaMalloc(&d_n, (size_t)n * sizeof(int));
  if (rc == cudaSuccess) rc = cudaMemcpy(d_p1, p1, b3, cudaMemcpyHostToDevice);
  if (rc == cudaSuccess) rc = cudaMemcpy(d_dir, dirs, b3, cudaMemcpyHostToDevice);
  if (rc == cudaSuccess) rc = cudaMemcpy(d_tri, tri, bt, cudaMemcpyHostToDevice);
  if (rc == cudaSuccess) rc = cudaMemcpy(d_u, pick_u, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (rc == cudaSuccess) {
    mgs_antipodal_kernel<<<(n + SMP_WARPS - 1) / SMP_WARPS, SMP_WARPS * 32>>>(n, d_p1, d_dir, nface, d_tri, eps, d_u, d_t, d_n);
    rc = cudaGetLastError();
  }
  if (rc == cudaSuccess) rc = cudaMemcpy(signed_t_out, d_t, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
  if (rc == cudaSuccess) rc = cudaMemcpy(nvalid_out, d_n, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_p1); cudaFree(d_dir); cudaFree(d_tri); cudaFree(d_u); cudaFree(d_t); cudaFree(d_n);
  if (rc != cudaSuccess) return mgs_set_error((std::string("mgs_antipodal_hits: ") + cudaGetErrorString(rc)).c_str());
  return 0;
}
