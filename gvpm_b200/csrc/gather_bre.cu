// gather_bre.cu — G-BRE gather in two kernels.
//
// k_bre_traverse: one warp per PACKET of 4 consecutive camera rays.  The warp walks the implicit
//   32-ary AABB hierarchy over the Morton-sorted photons without a node stack (one ballot mask per
//   level is the whole traversal state): lane c tests child c of the current node against the
//   packet's central ray, fattened by the packet's spread (two coalesced 128-bit loads per lane).  At
//   a leaf, lane c holds photon c (one 128-bit load) and tests it against each ray of the packet:
//   first a relaxed FMA pre-test, then — for the rare candidates — the reference's predicate in
//   strictly rounded arithmetic (gvpm_accel.h:293-301, shift_volume_photon.cpp:707-724).  Photons that
//   pass the depth/mode/pathSet filters are compacted through a per-warp shared-memory queue into a
//   global (ray, photon) pair list, 32 pairs per coalesced store.  Incoherent packets fall back to
//   one traversal per ray; results never depend on the packeting.
// k_bre_shade: one THREAD per (ray, photon) pair, so the divergent shift code (null shift, diffuse
//   reconnection, MIS; bre_device.cuh) always runs with full warps whatever the number of
//   neighbours per ray.  27 accumulators per thread in registers, a segmented warp scan by ray id
//   (shuffles) folds the lanes of one ray, and the last lane of each run adds 27 floats to the ray's
//   output row.
#include "bre_device.cuh"

namespace gvpm {

#ifndef GVPM_TRAV_WARPS
#define GVPM_TRAV_WARPS 4
#endif
#ifndef GVPM_TRAV_MIN_BLOCKS
#define GVPM_TRAV_MIN_BLOCKS 8
#endif
#ifndef GVPM_SHADE_THREADS
#define GVPM_SHADE_THREADS 128
#endif
#ifndef GVPM_SHADE_MIN_BLOCKS
#define GVPM_SHADE_MIN_BLOCKS 4
#endif
constexpr int kTravWarps = GVPM_TRAV_WARPS;
constexpr int PK = GVPM_PACKET;  // rays per packet
constexpr int kQueue = 64;

struct TravShared {
  float4 ray[PK][4];           // base record of each ray of the packet
  uint32_t queue[PK][kQueue];  // per ray: sorted photon slots waiting to be flushed
  uint32_t mask[GVPM_MAX_LEVELS];
  uint32_t base[GVPM_MAX_LEVELS];
};

// flush `take` entries from the top of one ray's queue to the global pair list (one ray per flush,
// so the pairs of a ray form contiguous runs for the shading kernel's segmented reduction)
__device__ __forceinline__ void flush_pairs(const GatherParams &P, const uint32_t *queue, uint32_t ray,
                                            uint32_t &qn, uint32_t take, int lane) {
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(P.pair_counter, (unsigned long long)take);
  base = __shfl_sync(0xffffffffu, base, 0);
  qn -= take;
  if ((uint32_t)lane < take) {
    const unsigned long long idx = base + lane;
    if (idx < P.pair_cap) P.pairs[idx] = make_uint2(ray, queue[qn + lane]);
  }
  __syncwarp();
}

template <bool DUMP>
__global__ void __launch_bounds__(kTravWarps * 32, GVPM_TRAV_MIN_BLOCKS)
k_bre_traverse(const __grid_constant__ GatherParams P) {
  __shared__ TravShared sh[kTravWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  TravShared &S = sh[w];
  const Tree &T = P.tree;
  const int top = T.levels - 1;
  const uint32_t nPackets = (P.ray_end - P.ray_begin + PK - 1) / PK;
  const float coordMag = T.n ? __ldg(P.bounds + 6) : 0.f;
  // packets whose rays drift apart by more than this are traversed ray by ray (performance only)
  float spreadMax = 16.f * P.radius;
  if (T.n) {
    const float ex = __ldg(P.bounds + 3) - __ldg(P.bounds), ey = __ldg(P.bounds + 4) - __ldg(P.bounds + 1),
                ez = __ldg(P.bounds + 5) - __ldg(P.bounds + 2);
    spreadMax = fmaxf(spreadMax, 0.01f * sqrtf(ex * ex + ey * ey + ez * ez));
  }

  for (;;) {
    uint32_t pk = 0;
    if (lane == 0) pk = atomicAdd(P.work_counter, 1u);
    pk = __shfl_sync(0xffffffffu, pk, 0);
    if (pk >= nPackets) break;
    const uint32_t r0 = P.ray_begin + pk * PK;
    const int nr = min((uint32_t)PK, P.ray_end - r0);
    __syncwarp();
    if (lane < 4 * nr) S.ray[lane >> 2][lane & 3] = ldg4(P.rays + (size_t)(r0 + (lane >> 2)) * GVPM_RAY_FLOAT4 + (lane & 3));
    __syncwarp();

    // per-ray state, warp-uniform registers (static indexing: loops over j are fully unrolled)
    float ox[PK], oy[PK], oz[PK], dx[PK], dy[PK], dz[PK], mint[PK], elen[PK];
    uint32_t nGeom[PK], nContrib[PK], qn[PK];
    uint32_t actMask = 0;
#pragma unroll
    for (int j = 0; j < PK; ++j) {
      qn[j] = 0;
      const int jj = j < nr ? j : 0;
      const float4 b0 = S.ray[jj][0], b1 = S.ray[jj][1], b2 = S.ray[jj][2];
      ox[j] = b0.x; oy[j] = b0.y; oz[j] = b0.z; mint[j] = b0.w;
      dx[j] = b1.x; dy[j] = b1.y; dz[j] = b1.z; elen[j] = b2.w;
      nGeom[j] = 0; nContrib[j] = 0;
      if (j < nr && T.n > 0 && elen[j] >= mint[j]) actMask |= 1u << j;
    }
    // packet spread: max distance between corresponding points of ray j and the central ray over
    // [0, tEnd] (linear in t, so attained at an end).  Coherent packets share one traversal.
    uint32_t groups[PK];
    int ng = 0;
    float spread = 0.f;
    if (actMask) {
      const int c0 = __ffs(actMask) - 1;
      float cox = 0, coy = 0, coz = 0, cdx = 0, cdy = 0, cdz = 0, tEnd = 0;
#pragma unroll
      for (int j = 0; j < PK; ++j) {
        if (j == c0) { cox = ox[j]; coy = oy[j]; coz = oz[j]; cdx = dx[j]; cdy = dy[j]; cdz = dz[j]; }
        if (actMask >> j & 1) tEnd = fmaxf(tEnd, elen[j]);
      }
      tEnd += P.radius;
#pragma unroll
      for (int j = 0; j < PK; ++j)
        if (actMask >> j & 1) {
          const float ax = ox[j] - cox, ay = oy[j] - coy, az = oz[j] - coz;
          const float bx = ax + tEnd * (dx[j] - cdx), by = ay + tEnd * (dy[j] - cdy), bz = az + tEnd * (dz[j] - cdz);
          spread = fmaxf(spread, fmaxf(sqrtf(ax * ax + ay * ay + az * az), sqrtf(bx * bx + by * by + bz * bz)));
        }
      spread *= 1.0001f;
      if (spread <= spreadMax) {
        groups[ng++] = actMask;
      } else {
        spread = 0.f;
#pragma unroll
        for (int j = 0; j < PK; ++j)
          if (actMask >> j & 1) groups[ng++] = 1u << j;
      }
    }

    for (int g = 0; g < ng; ++g) {
      const uint32_t gm = groups[g];
      const int c0 = __ffs(gm) - 1;
      float cox = 0, coy = 0, coz = 0, cdx = 1, cdy = 1, cdz = 1, tloG = 3.4e38f, thiG = -3.4e38f, omag = 0.f;
#pragma unroll
      for (int j = 0; j < PK; ++j) {
        if (j == c0) { cox = ox[j]; coy = oy[j]; coz = oz[j]; cdx = dx[j]; cdy = dy[j]; cdz = dz[j]; }
        if (gm >> j & 1) {
          tloG = fminf(tloG, mint[j]);
          thiG = fmaxf(thiG, elen[j]);
          omag = fmaxf(omag, fmaxf(fmaxf(fabsf(ox[j]), fabsf(oy[j])), fabsf(oz[j])));
        }
      }
      // conservative culling: relaxed arithmetic made safe by `fpad` (a few ulp of the coordinate
      // magnitudes: rounding of the slab test and of the strict predicate) plus the packet spread
      const float fpad = (omag + coordMag + fabsf(thiG) + P.radius) * 3.8147e-6f;  // 2^-18
      const float pad = fpad + ((gm & (gm - 1)) ? spread : 0.f);
      const float oxp = cox + pad, oxm = cox - pad, oyp = coy + pad, oym = coy - pad, ozp = coz + pad,
                  ozm = coz - pad;
      const float ix = 1.f / cdx, iy = 1.f / cdy, iz = 1.f / cdz;
      const float tlo = tloG - 4.f * fpad;
      const float thi = thiG + P.radius + 4.f * fpad;
      const float rpad2 = (P.radius + fpad) * (P.radius + fpad);

      uint32_t cur, base = 0;
      int l = top;
      cur = __ballot_sync(0xffffffffu, (uint32_t)lane < T.cnt[top] &&
                                           box_hit(T, T.off[top] + lane, oxp, oxm, oyp, oym, ozp, ozm, ix, iy,
                                                   iz, tlo, thi));
      for (;;) {
        if (cur == 0) {
          if (l == top) break;
          ++l;
          cur = S.mask[l];
          base = S.base[l];
          continue;
        }
        const int c = __ffs(cur) - 1;
        cur &= cur - 1;
        const uint32_t node = base + c;
        if (l > 0) {
          S.mask[l] = cur;
          S.base[l] = base;
          --l;
          base = node << 5;
          const uint32_t idx = base + lane;
          cur = __ballot_sync(0xffffffffu, idx < T.cnt[l] && box_hit(T, T.off[l] + idx, oxp, oxm, oyp, oym, ozp,
                                                                       ozm, ix, iy, iz, tlo, thi));
          continue;
        }
        // ---- leaf: lane holds photon (node*32 + lane) ----
        const uint32_t pi = (node << 5) + lane;
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t cand = 0;
        if (pi < T.n) {
          q0 = ldg4(P.planes + pi);
#pragma unroll
          for (int j = 0; j < PK; ++j) {
            // relaxed (FMA) pre-test, conservative by fpad
            const float cx = q0.x - ox[j], cy = q0.y - oy[j], cz = q0.z - oz[j];
            const float dd = cx * dx[j] + cy * dy[j] + cz * dz[j];
            const float qx = cx - dd * dx[j], qy = cy - dd * dy[j], qz = cz - dd * dz[j];
            if ((gm >> j & 1) && (qx * qx + qy * qy + qz * qz) < rpad2 && dd > tlo) cand |= 1u << j;
          }
        }
        const uint32_t anyc = __reduce_or_sync(0xffffffffu, cand);
        if (anyc == 0) continue;
#pragma unroll
        for (int j = 0; j < PK; ++j) {
          if (!(anyc >> j & 1)) continue;  // warp-uniform
          bool geom = false, contrib = false;
          if (cand >> j & 1) {
            const BaseRay R = load_base_ray(S.ray[j]);
            sf tB, pc;
            geom = base_distance(P, R, v3(q0.x, q0.y, q0.z), tB, pc);
            contrib = geom && filters_pass(P, R, __float_as_uint(q0.w));
          }
          const uint32_t gmask = __ballot_sync(0xffffffffu, geom), cmask = __ballot_sync(0xffffffffu, contrib);
          if (DUMP) {
            if (geom) {
              const uint32_t rank = __popc(gmask & ((1u << lane) - 1u));
              P.nbr_idx[P.nbr_offsets[r0 + j] + nGeom[j] + rank] = P.orig[pi] | (contrib ? 0x80000000u : 0u);
            }
          }
          nGeom[j] += __popc(gmask);
          nContrib[j] += __popc(cmask);
          if (DUMP || cmask == 0) continue;
          if (contrib) S.queue[j][qn[j] + __popc(cmask & ((1u << lane) - 1u))] = pi;
          qn[j] += __popc(cmask);
          __syncwarp();
          if (qn[j] >= 32) flush_pairs(P, S.queue[j], r0 + j, qn[j], 32, lane);
        }
      }
    }
    if (!DUMP) {
#pragma unroll
      for (int j = 0; j < PK; ++j)
        if (qn[j] > 0) flush_pairs(P, S.queue[j], r0 + j, qn[j], qn[j], lane);
    }
    if (P.counts && lane < nr) {
      uint32_t g = 0, c = 0;
#pragma unroll
      for (int j = 0; j < PK; ++j)
        if (lane == j) { g = nGeom[j]; c = nContrib[j]; }
      P.counts[2 * (size_t)(r0 + lane)] = g;
      P.counts[2 * (size_t)(r0 + lane) + 1] = c;
    }
  }
}

__global__ void __launch_bounds__(GVPM_SHADE_THREADS, GVPM_SHADE_MIN_BLOCKS)
k_bre_shade(const __grid_constant__ GatherParams P) {
  const int lane = threadIdx.x & 31;
  unsigned long long total = *P.pair_counter;
  if (total > P.pair_cap) total = P.pair_cap;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < total;
       i0 += stride) {
    const unsigned long long i = i0 + lane;
    const bool valid = i < total;
    uint2 pr = make_uint2(0xffffffffu, 0u);
    if (valid) pr = P.pairs[i];
    float a[GVPM_OUT_FLOATS];
#pragma unroll
    for (int j = 0; j < GVPM_OUT_FLOATS; ++j) a[j] = 0.f;
    if (valid) bre_photon(P, P.rays + (size_t)pr.x * GVPM_RAY_FLOAT4, pr.y, a);
    // segmented inclusive scan over RUNS of equal ray id (a ray's pairs arrive in contiguous runs,
    // one per flush; the same ray may own several runs, each adds its own partial sum)
    const uint32_t key = pr.x;
    const uint32_t kprev = __shfl_up_sync(0xffffffffu, key, 1);
    const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || kprev != key);
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // first lane of my run
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const bool same = lane - off >= start;
#pragma unroll
      for (int j = 0; j < GVPM_OUT_FLOATS; ++j) {
        const float vu = __shfl_up_sync(0xffffffffu, a[j], off);
        if (same) a[j] += vu;
      }
    }
    if (valid && (lane == 31 || (heads >> (lane + 1) & 1u))) {
      float *o = P.out + (size_t)key * GVPM_OUT_FLOATS;
#pragma unroll
      for (int j = 0; j < GVPM_OUT_FLOATS; ++j) atomicAdd(o + j, a[j]);
    }
  }
}

// ---- host-side launchers (called from gvpm_capi.cu) -------------------------------------------
static int g_trav_blocks[2] = {0, 0}, g_shade_blocks = 0;

cudaError_t launch_bre_traverse(const GatherParams &P, bool dump, int sm_count, cudaStream_t stream) {
  if (P.ray_end <= P.ray_begin) return cudaSuccess;
  int &bps = g_trav_blocks[dump ? 1 : 0];
  if (bps == 0) {
    if (dump) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_bre_traverse<true>, kTravWarps * 32, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_bre_traverse<false>, kTravWarps * 32, 0);
    if (bps < 1) bps = 1;
  }
  // persistent grid: a whole number of resident CTAs per SM; warps pull packets from a counter
  unsigned grid = (unsigned)(sm_count * bps);
  const unsigned packets = (P.ray_end - P.ray_begin + PK - 1) / PK;
  const unsigned need = (packets + kTravWarps - 1) / kTravWarps;
  if (grid > need) grid = need;
  if (dump) k_bre_traverse<true><<<grid, kTravWarps * 32, 0, stream>>>(P);
  else k_bre_traverse<false><<<grid, kTravWarps * 32, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_bre_shade(const GatherParams &P, unsigned long long total, int sm_count, cudaStream_t stream) {
  if (total == 0) return cudaSuccess;
  if (g_shade_blocks == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_shade_blocks, k_bre_shade, GVPM_SHADE_THREADS, 0);
    if (g_shade_blocks < 1) g_shade_blocks = 1;
  }
  unsigned long long need = (total + GVPM_SHADE_THREADS - 1) / GVPM_SHADE_THREADS;
  unsigned long long grid = (unsigned long long)sm_count * g_shade_blocks * 4;  // a few waves, grid-stride
  if (grid > need) grid = need;
  k_bre_shade<<<(unsigned)grid, GVPM_SHADE_THREADS, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace gvpm
