// gradient.cu — GPMIntegrator::computeGradient (gvpm/gvpm.cpp:1205-1306) restricted to the volume
// terms (APA estimators: no division by m_totalEmittedVolume, :1258-1259) and the throughput plane
// (:479-500): the three RGB planes handed to poisson::Solver::importImagesMTS (gvpm.cpp:560-578).
//   Gx(x,y) = (S_R - W_R)(x,y) + (W_L - S_L)(x+1,y); border pixels keep the forward part only.
#include "gvpm_device.cuh"

namespace gvpm {

// reuse != 0: the throughput plane is the reusePrimal estimate of gvpm.cpp:503-532 instead of mediumFlux: the shifted
// flux the four neighbours send to this pixel plus the pixel's own four weighted fluxes, over 4 (and over
// m_totalEmittedVolume for the estimators that are not APA: inv_emitted = 1 / m_totalEmittedVolume, else 1).
__global__ void k_gradient(const float *__restrict__ acc, int w, int h, int use_abs, int reuse, float inv_emitted,
                           float *__restrict__ thr, float *__restrict__ gx, float *__restrict__ gy) {
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
    float tp = a[c];
    if (reuse) {
      float T = 0.f;
      if (x != w - 1) T += acc[(p + 1) * GVPM_OUT_FLOATS + 3 * (1 + 0) + c];   // right neighbour's ELeft shift
      if (x != 0) T += acc[(p - 1) * GVPM_OUT_FLOATS + 3 * (1 + 1) + c];       // left neighbour's ERight shift
      if (y != h - 1) T += acc[(p + w) * GVPM_OUT_FLOATS + 3 * (1 + 3) + c];   // (x, y + 1): EBottom
      if (y != 0) T += acc[(p - w) * GVPM_OUT_FLOATS + 3 * (1 + 2) + c];       // (x, y - 1): ETop
      T += ((a[3 * (5 + 3) + c] + a[3 * (5 + 2) + c]) + a[3 * (5 + 1) + c]) + a[3 * (5 + 0) + c];
      tp = (T * 0.25f) * inv_emitted;
    }
    thr[3 * p + c] = tp;
    gx[3 * p + c] = use_abs ? fabsf(vx) : vx;
    gy[3 * p + c] = use_abs ? fabsf(vy) : vy;
  }
}

void launch_gradient(const float *acc, int w, int h, int use_abs, float *thr, float *gx, float *gy,
                     cudaStream_t st, int reuse, float inv_emitted) {
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  k_gradient<<<g, b, 0, st>>>(acc, w, h, use_abs, reuse, inv_emitted, thr, gx, gy);
}

}  // namespace gvpm
