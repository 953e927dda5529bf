// gradient.cu — GPMIntegrator::computeGradient (gvpm/gvpm.cpp:1205-1306) restricted to the volume
// terms (APA estimators: no division by m_totalEmittedVolume, :1258-1259) and the throughput plane
// (:479-500): the three RGB planes handed to poisson::Solver::importImagesMTS (gvpm.cpp:560-578).
//   Gx(x,y) = (S_R - W_R)(x,y) + (W_L - S_L)(x+1,y); border pixels keep the forward part only.
#include "gvpm_device.cuh"

namespace gvpm {

__global__ void k_gradient(const float *__restrict__ acc, int w, int h, int use_abs, float *__restrict__ thr,
                           float *__restrict__ gx, float *__restrict__ gy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t p = (size_t)y * w + x;
  const float *a = acc + p * GVPM_OUT_FLOATS;
  // layout: primal 0..2, shifted[k] at 3*(1+k), weighted[k] at 3*(5+k); k = L,R,T,B
  for (int c = 0; c < 3; ++c) {
    float vx = a[3 * (1 + 1) + c] - a[3 * (5 + 1) + c];
    if (x != w - 1) {
      const float *r = acc + (p + 1) * GVPM_OUT_FLOATS;
      vx = vx + (r[3 * (5 + 0) + c] - r[3 * (1 + 0) + c]);
    }
    float vy = a[3 * (1 + 2) + c] - a[3 * (5 + 2) + c];
    if (y != h - 1) {
      const float *t = acc + (p + w) * GVPM_OUT_FLOATS;
      vy = vy + (t[3 * (5 + 3) + c] - t[3 * (1 + 3) + c]);
    }
    thr[3 * p + c] = a[c];
    gx[3 * p + c] = use_abs ? fabsf(vx) : vx;
    gy[3 * p + c] = use_abs ? fabsf(vy) : vy;
  }
}

void launch_gradient(const float *acc, int w, int h, int use_abs, float *thr, float *gx, float *gy,
                     cudaStream_t st) {
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  k_gradient<<<g, b, 0, st>>>(acc, w, h, use_abs, thr, gx, gy);
}

}  // namespace gvpm
