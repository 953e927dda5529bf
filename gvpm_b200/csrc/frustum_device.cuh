// frustum_device.cuh — device functions of the perspective (frustum) grid shared by the build (tree_build.cu) and the
// photon dispatch between GPUs (dispatch.cu): the projection, the coarse ray-occupancy mask and the key of a photon.
// A rank that sends photons evaluates frustum_key with the RECEIVER's grid and occupancy mask, i.e. the very function
// (same inputs, same arithmetic) the receiver's own build evaluates: what it does not send is exactly what the receiver
// would drop.
#pragma once
#include "gvpm_device.cuh"

namespace gvpm {

// the projection both sides use (photons with q = p - C, rays with q = d): plane coordinates at distance 1 along m
__device__ __forceinline__ void frustum_project(const float *m, const float *u, const float *v, float qx, float qy, float qz,
                                                float &x, float &y, float &z) {
  z = qx * m[0] + qy * m[1] + qz * m[2];
  const float iz = 1.f / z;
  x = (qx * u[0] + qy * u[1] + qz * u[2]) * iz;
  y = (qx * v[0] + qy * v[1] + qz * v[2]) * iz;
}

// Coarse occupancy of the rays' projected directions: kOccRes x kOccRes bits over [xmin, xmax] x [ymin, ymax], built in
// shared memory per CTA and OR-ed into global memory.  frustum_key drops a photon when no bit under its footprint
// box is set (no ray can reach it): this is what leaves most of the photon set out of a rank's sort when the image is
// sharded over GPUs.
constexpr int kOccRes = 256;
constexpr int kOccWords = kOccRes * kOccRes / 32;   // 8 KB
__device__ __forceinline__ int occ_cell(float x, float lo, float inv) {
  return min(max((int)floorf((x - lo) * inv), 0), kOccRes - 1);
}

// Key of the photon at (px, py, pz): footprint class + cell (see FrustumGrid), NEAR = grids * n_cells, DROP = NEAR + 1.
// parity(): pathID & 1 of the photon, only evaluated when the grid is split by parity and the photon is binned.
template <class ParityFn>
__device__ __forceinline__ uint32_t frustum_key(const FrustumGrid &G, const uint32_t *__restrict__ occ, float px, float py,
                                                float pz, ParityFn parity) {
  const float qx = px - G.C[0], qy = py - G.C[1], qz = pz - G.C[2];
  const float rho = sqrtf(qx * qx + qy * qy + qz * qz);
  const uint32_t grids = G.parity_split ? 2u : 1u;
  const uint32_t NEAR = grids * G.n_cells, DROP = NEAR + 1u;
  float x, y, z;
  frustum_project(G.m, G.u, G.v, qx, qy, qz, x, y, z);
  const float pr = G.pad_r * 1.001f + 1e-6f * rho;
  if (rho <= 2.f * pr) return NEAR;                       // alpha >= 30 degrees
  if (z < 0.1f * rho) return rho <= 6.f * pr ? NEAR : DROP;   // more than 84 degrees off axis: only reachable when very close
  // footprint w = tan(theta + alpha) - tan(theta) with tan(theta) = |(x, y)| and sin(alpha) = pr / rho, in algebraic form
  // (tan of a sum; no atan / asin / tan): tan(alpha) = s / sqrt(1 - s^2), slightly enlarged; theta + alpha beyond ~83
  // degrees (tan > 8.24, or the sum past 90 degrees: non-positive denominator) is the NEAR case
  const float tanT = sqrtf(x * x + y * y);
  const float sA = pr / rho;                                  // < 0.5 here (rho > 2 pr)
  const float tanA = sA * rsqrtf(1.f - sA * sA) * 1.001f + 1e-6f;
  const float den = 1.f - tanT * tanA;
  if (den <= 1e-3f) return NEAR;
  const float tanS = (tanT + tanA) / den;
  if (tanS >= 8.24f) return NEAR;
  const float wfoot = (tanS - tanT) * 1.01f + 1e-6f * (1.f + tanT);   // 1 % under the class's cell edge
  if (x < G.xmin - wfoot || x > G.xmax + wfoot || y < G.ymin - wfoot || y > G.ymax + wfoot) return DROP;
  int c = 0;
  while (c < G.classes && wfoot > G.csize[c]) ++c;
  if (c >= G.classes) return NEAR;
  const float ic = 1.f / G.csize[c];
  const int nx = (int)G.nx[c], ny = (int)G.ny[c];
  int cx = (int)floorf((x - G.gx0) * ic), cy = (int)floorf((y - G.gy0) * ic);
  cx = min(max(cx, 0), nx - 1);
  cy = min(max(cy, 0), ny - 1);
  // any ray under the photon's footprint box?  (coarse bitmask, conservative: the box is padded by one ulp-ish
  // margin through wfoot's own 1 % pad)
  const float oix = kOccRes / fmaxf(G.xmax - G.xmin, 1e-20f), oiy = kOccRes / fmaxf(G.ymax - G.ymin, 1e-20f);
  const int ox0 = occ_cell(x - wfoot, G.xmin, oix), ox1 = occ_cell(x + wfoot, G.xmin, oix);
  const int oy0 = occ_cell(y - wfoot, G.ymin, oiy), oy1 = occ_cell(y + wfoot, G.ymin, oiy);
  bool any = false;
  for (int yy = oy0; yy <= oy1 && !any; ++yy)
    for (int w0 = ox0 >> 5; w0 <= (ox1 >> 5); ++w0) {
      const int b0 = max(ox0 - 32 * w0, 0), b1 = min(ox1 - 32 * w0, 31);
      const uint32_t bits = (0xffffffffu >> (31 - b1)) & (0xffffffffu << b0);
      any = any || (__ldg(occ + yy * (kOccRes / 32) + w0) & bits) != 0u;
    }
  if (!any) return DROP;
  const uint32_t par = G.parity_split ? (parity() & 1u) : 0u;
  return par * G.n_cells + G.base[c] + (uint32_t)cy * (uint32_t)nx + (uint32_t)cx;
}

}  // namespace gvpm
