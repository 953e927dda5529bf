// generate.cu — the two host stages on either side of the gather, on the device, for the
// box scene class of include/gvpm_b200.h (gvpm_box_scene):
//   k_generate_rays   camera-ray medium segments + 4 offset segments of a pinhole sensor: what GatherPointMap::generate
//                     (gvpm/gvpm_gatherpoint.h:259-486), ShiftGatherPoint::generate (gvpm/shift/shift_cameraPath.h:146-413)
//                     and sensorMIS (gvpm_struct.h:608-631) produce per pixel, written straight into the ray staging
//                     arrays (then packed by gvpm_commit_rays)
//   k_trace_count / k_trace_emit   the light-path random walk of GradientPhotonProcess (gvpm/gvpm_proc.cpp:125-210)
//                     + GPhotonMap::tryAppend (gvpm_accel.h:119-199): one thread per light path, two passes (count,
//                     exclusive scan over the paths, emit), so the photons land in path order whatever the launch
//                     geometry and the set is bit-reproducible; photon flux / pdf bookkeeping as libbidir's
//                     (SURVEY.md §9.1).
// Random numbers: PCG32 keyed by (seed, pixel index) / (seed, path index).  Arithmetic: strictly rounded fp32 in a fixed
// operation order; log, exp, sin and cos are polynomial routines (pm_*) built from +, -, *, / only, so a CPU
// restatement compiled without FMA contraction reproduces every record bit for bit (tests/test_gpu_generate.py).
#include "gvpm_device.cuh"

namespace gvpm {

// ---- counter-based RNG ---------------------------------------------------------------------------------------------
struct Pcg32 {
  unsigned long long state, inc;
  __device__ Pcg32(unsigned long long seed, unsigned long long seq) {
    state = 0;
    inc = (seq << 1u) | 1u;
    next();
    state += seed;
    next();
  }
  __device__ uint32_t next() {
    const unsigned long long old = state;
    state = old * 6364136223846793005ULL + inc;
    const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
  }
  __device__ float uniform() { return __fmul_rn((float)(next() >> 8), 1.0f / 16777216.0f); }  // [0,1)
};

// ---- portable elementary functions (strictly rounded, fixed order) ----------------------------------------------------
__device__ __forceinline__ float pm_log(float x) {   // x > 0, normal
  const uint32_t bits = __float_as_uint(x);
  int e = (int)(bits >> 23) - 127;
  sf m(__uint_as_float((bits & 0x007fffffu) | 0x3f800000u));   // [1, 2)
  if (m.v > 1.41421356f) { m = m * sf(0.5f); e += 1; }
  const sf s = (m - sf(1.f)) / (m + sf(1.f));
  const sf s2 = s * s;
  sf p(0.11111111f);
  p = p * s2 + sf(0.14285715f);
  p = p * s2 + sf(0.2f);
  p = p * s2 + sf(0.33333334f);
  p = p * s2 + sf(1.0f);
  return ((sf(2.f) * s) * p + sf((float)e) * sf(0.69314718f)).v;
}
__device__ __forceinline__ float pm_exp(float x) {   // x <= ~0
  if (x < -87.f) return 0.f;
  const sf n(floorf((sf(x) * sf(1.44269504f) + sf(0.5f)).v));
  sf r = sf(x) - n * sf(0.693359375f);
  r = r - n * sf(-2.12194440e-4f);
  sf p(1.0f / 720.0f);
  p = p * r + sf(1.0f / 120.0f);
  p = p * r + sf(1.0f / 24.0f);
  p = p * r + sf(1.0f / 6.0f);
  p = p * r + sf(0.5f);
  p = p * r + sf(1.0f);
  p = p * r + sf(1.0f);
  const int ni = (int)n.v;
  return (sf(__uint_as_float((uint32_t)(ni + 127) << 23)) * p).v;
}
__device__ __forceinline__ void pm_sincos2pi(float u, float &sn, float &cs) {   // sin / cos of 2 pi u, u in [0, 1)
  const sf q(floorf((sf(u) * sf(4.f) + sf(0.5f)).v));
  const sf a = (sf(u) - q * sf(0.25f)) * sf(6.2831855f);   // [-pi/4, pi/4]
  const sf a2 = a * a;
  sf p(-1.9841270e-4f);
  p = p * a2 + sf(8.3333333e-3f);
  p = p * a2 + sf(-0.16666667f);
  p = p * a2 + sf(1.0f);
  const sf s = a * p;
  sf c(2.4801587e-5f);
  c = c * a2 + sf(-1.3888889e-3f);
  c = c * a2 + sf(4.1666668e-2f);
  c = c * a2 + sf(-0.5f);
  c = c * a2 + sf(1.0f);
  switch ((int)q.v & 3) {
    case 0: sn = s.v; cs = c.v; break;
    case 1: sn = c.v; cs = -s.v; break;
    case 2: sn = -s.v; cs = -c.v; break;
    default: sn = -c.v; cs = s.v; break;
  }
}

// ---- scene ---------------------------------------------------------------------------------------------------------
struct SceneHit { sf t; v3 n, albedo; bool escaped; };
__device__ __forceinline__ float comp(const v3 &a, int axis) { return axis == 0 ? a.x.v : (axis == 1 ? a.y.v : a.z.v); }
// nearest of the five walls, the open face and the rectangles, in this order (ties keep the earlier one)
__device__ __forceinline__ SceneHit intersect_scene(const gvpm_box_scene &S, v3 o, v3 d) {
  SceneHit h;
  h.t = sf(1e30f);
  h.escaped = false;
  h.n = v3(0.f, 0.f, 0.f);
  h.albedo = v3(0.7f, 0.7f, 0.7f);
  auto plane = [&](int axis, float pos, v3 n, const float *alb, bool esc) {
    const float dc = comp(d, axis), oc = comp(o, axis);
    if (dc == 0.f) return;
    const sf t = (sf(pos) - sf(oc)) / sf(dc);
    if (t.v > 1e-6f && t < h.t) {
      h.t = t; h.n = n; h.escaped = esc;
      h.albedo = esc ? v3(0.f, 0.f, 0.f) : v3(alb[0], alb[1], alb[2]);
    }
  };
  plane(0, S.lo[0], v3(1.f, 0.f, 0.f), S.face_albedo[0], false);
  plane(0, S.hi[0], v3(-1.f, 0.f, 0.f), S.face_albedo[1], false);
  plane(1, S.lo[1], v3(0.f, 1.f, 0.f), S.face_albedo[2], false);
  plane(1, S.hi[1], v3(0.f, -1.f, 0.f), S.face_albedo[3], false);
  plane(2, S.hi[2], v3(0.f, 0.f, -1.f), S.face_albedo[4], false);
  plane(2, S.lo[2], v3(0.f, 0.f, 1.f), S.face_albedo[4], true);
  if (d.y.v != 0.f)
    for (int r = 0; r < S.n_rects; ++r) {
      const sf t = (sf(S.rect[r].y) - o.y) / d.y;
      if (t.v > 1e-6f && t < h.t) {
        const sf x = o.x + d.x * t, z = o.z + d.z * t;
        if (x.v >= S.rect[r].x0 && x.v <= S.rect[r].x1 && z.v >= S.rect[r].z0 && z.v <= S.rect[r].z1) {
          h.t = t;
          h.n = d.y.v < 0.f ? v3(0.f, 1.f, 0.f) : v3(0.f, -1.f, 0.f);
          h.albedo = v3(S.rect[r].albedo[0], S.rect[r].albedo[1], S.rect[r].albedo[2]);
          h.escaped = false;
        }
      }
    }
  return h;
}
__device__ __forceinline__ v3 unit(v3 a) { return a * (sf(1.0f) / length(a)); }

// ---- f-2: camera rays --------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool make_ray(const RayGenParams &P, sf tx, sf ty, sf sx, sf sy, v3 &ro, v3 &rd, sf &rl) {
  const gvpm_box_scene &S = P.scene;
  const sf w((float)P.cam.film_w), h((float)P.cam.film_h);
  const v3 dir = unit(v3(((sx / w - sf(0.5f)) * sf(2.f)) * tx, ((sy / h - sf(0.5f)) * sf(2.f)) * ty, sf(1.0f)));
  const v3 cam(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]);
  rd = dir;
  if (P.cam.inside_medium) {
    ro = cam;
  } else {
    const sf tIn = (sf(S.lo[2]) - cam.z) / dir.z;
    ro = cam + dir * tIn;
    ro.z = sf(S.lo[2]);
    if (ro.x.v <= S.lo[0] || ro.x.v >= S.hi[0] || ro.y.v <= S.lo[1] || ro.y.v >= S.hi[1]) { rl = sf(0.f); return false; }
  }
  const SceneHit hit = intersect_scene(S, ro, rd);
  rl = hit.t;
  return hit.t.v > (sf(4.f) * sf(P.epsilon)).v && !hit.escaped;
}
__device__ __forceinline__ int compact1by1(uint32_t v) {
  v &= 0x55555555u;
  v = (v ^ (v >> 1)) & 0x33333333u;
  v = (v ^ (v >> 2)) & 0x0f0f0f0fu;
  v = (v ^ (v >> 4)) & 0x00ff00ffu;
  v = (v ^ (v >> 8)) & 0x0000ffffu;
  return (int)v;
}
__device__ __forceinline__ void put3(float *dst, size_t i, v3 v) { dst[3 * i] = v.x.v; dst[3 * i + 1] = v.y.v; dst[3 * i + 2] = v.z.v; }

// one CTA per block x block pixel block (block <= 32: 1024 threads); the pixels of a block that fall inside the image
// are compacted in block order (row-major or Z-order), blocks follow each other row by row
__global__ void __launch_bounds__(1024) k_generate_rays(const __grid_constant__ RayGenParams P) {
  __shared__ uint32_t warpTot[32];
  const int w = P.cam.film_w, bs = P.block;
  const int blocksX = (w + bs - 1) / bs;
  const int bx = (blockIdx.x % blocksX) * bs, by = P.y0 + (blockIdx.x / blocksX) * bs;
  const int i = threadIdx.x, lane = i & 31, wp = i >> 5;
  const int x = bx + (P.zorder ? compact1by1((uint32_t)i) : i % bs), y = by + (P.zorder ? compact1by1((uint32_t)i >> 1) : i / bs);
  const bool valid = i < bs * bs && x < w && y < P.y1;
  const uint32_t vm = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) warpTot[wp] = __popc(vm);
  __syncthreads();
  uint32_t before = __popc(vm & ((1u << lane) - 1u));
  for (int k = 0; k < wp; ++k) before += warpTot[k];
  if (!valid) return;
  const int rowsHere = min(bs, P.y1 - by);
  const size_t n = (size_t)w * (size_t)(by - P.y0) + (size_t)bx * (size_t)rowsHere + before;

  const sf tx(P.cam.tan_half_fov_x), ty = tx * sf((float)P.cam.film_h) / sf((float)w);
  Pcg32 rng(P.seed ^ 0x9E3779B97F4A7C15ULL, (unsigned long long)y * (unsigned long long)w + (unsigned long long)x);
  const sf sx = sf((float)x) + sf(rng.uniform()), sy = sf((float)y) + sf(rng.uniform());
  v3 ro, rd;
  sf rl;
  const bool ok = make_ray(P, tx, ty, sx, sy, ro, rd, rl);
  put3(P.o, n, ro);
  put3(P.d, n, rd);
  P.mint[n] = P.epsilon;
  P.maxt[n] = ok ? (rl - sf(P.epsilon)).v : 0.f;   // empty segment when the pixel misses the medium
  P.edge_len[n] = ok ? rl.v : 0.f;
  put3(P.eye_contrib, n, v3(1.f, 1.f, 1.f));
  P.xi[n] = rng.uniform();
  P.px[n] = x;
  P.py[n] = y;
  P.edge_id[n] = P.cam.inside_medium ? 1 : 2;
  const int off[4][2] = {{-1, 0}, {1, 0}, {0, 1}, {0, -1}};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v3 ko, kd;
    sf kl;
    const bool kv = make_ray(P, tx, ty, sx + sf((float)off[k][0]), sy + sf((float)off[k][1]), ko, kd, kl);
    P.off_valid[4 * n + k] = kv ? 1 : 0;
    put3(P.off_o, 4 * n + k, ko);
    put3(P.off_d, 4 * n + k, kd);
    P.off_len[4 * n + k] = kl.v;
    put3(P.off_eye, 4 * n + k, v3(1.f, 1.f, 1.f));
    P.off_sensor[4 * n + k] = 1.0f;   // sensorMIS of a pinhole (gvpm_struct.h:608-631: pdf ratio x Jacobian = 1)
  }
}

// ---- f-1: light-path random walk ----------------------------------------------------------------------------------------
__device__ __forceinline__ void frame_of(v3 n, v3 &s, v3 &t) {
  if (fabsf(n.x.v) > fabsf(n.y.v)) {
    const sf il = sf(1.0f) / ssqrt(n.x * n.x + n.z * n.z);
    t = v3(n.z * il, sf(0.f), -(n.x * il));
  } else {
    const sf il = sf(1.0f) / ssqrt(n.y * n.y + n.z * n.z);
    t = v3(sf(0.f), n.z * il, -(n.y * il));
  }
  s = cross(t, n);
}
__device__ __forceinline__ v3 to_world(v3 n, v3 l) {
  v3 s, t;
  frame_of(n, s, t);
  return (s * l.x + t * l.y) + n * l.z;
}
__device__ __forceinline__ v3 cosine_hemisphere(float u1, float u2) {
  float sn, cs;
  pm_sincos2pi(u2, sn, cs);
  const sf r = ssqrt(sf(u1));
  const sf x = r * sf(cs), y = r * sf(sn);
  return v3(x, y, safe_sqrt((sf(1.f) - x * x) - y * y));
}
__device__ __forceinline__ v3 uniform_sphere(float u1, float u2) {
  float sn, cs;
  pm_sincos2pi(u2, sn, cs);
  const sf z = sf(1.f) - sf(2.f) * sf(u1);
  const sf r = safe_sqrt(sf(1.f) - z * z);
  return v3(r * sf(cs), r * sf(sn), z);
}
__device__ __forceinline__ sf hg_eval(sf g, sf cosWiWo) {   // phase/hg.cpp:107-110 with dot(wi, wo)
  const sf temp = (sf(1.0f) + g * g) + (sf(2.0f) * g) * cosWiWo;
  return (sf(GVPM_INV_FOURPI) * (sf(1.f) - g * g)) / (temp * ssqrt(temp));
}

// The walk of light path `pathIdx`.  EMIT = false: returns the number of photons it would store.  EMIT = true: stores
// them at slots first, first + 1, ... (< n_total) with path id `pid`.
template <bool EMIT>
__device__ __forceinline__ uint32_t walk_path(const TraceParams &P, unsigned long long pathIdx, unsigned long long first,
                                              uint32_t pid) {
  const gvpm_box_scene &S = P.scene;
  Pcg32 rng(P.seed, pathIdx);
  const sf sigS(P.sigma_s), sigT = sf(P.sigma_s) + sf(P.sigma_a);
  // vertex 0 = emitter supernode (weight = power), vertex 1 = emitter sample
  v3 curPos, curN(0.f, -1.f, 0.f), curAlbedo(0.f, 0.f, 0.f), curWeight(1.f, 1.f, 1.f), prevPos(1.f, 1.f, 1.f);
  {
    const float u1 = rng.uniform(), u2 = rng.uniform();
    curPos = v3(sf(S.light_x0) + (sf(S.light_x1) - sf(S.light_x0)) * sf(u1), sf(S.light_y),
                sf(S.light_z0) + (sf(S.light_z1) - sf(S.light_z0)) * sf(u2));
  }
  int curType = GVPM_PARENT_EMITTER;
  sf curRr(1.f);
  v3 dir;
  {
    const float u1 = rng.uniform(), u2 = rng.uniform();
    dir = to_world(curN, cosine_hemisphere(u1, u2));
  }
  sf pdfOmega = smax(sf(0.f), dot(dir, curN)) * sf(GVPM_INV_PI);
  v3 thr(S.light_power, S.light_power, S.light_power);
  int ci = 1;
  uint32_t appended = 0;
  const int firstStored = max(2, P.min_depth + 1);
  for (;;) {
    const SceneHit h = intersect_scene(S, curPos, dir);
    const sf t = (-sf(pm_log((sf(1.0f) - sf(rng.uniform())).v))) / sigT;
    const bool inMedium = t < h.t;
    const sf L = inMedium ? t : h.t;
    if (!inMedium && h.escaped) break;
    const sf T(pm_exp(((-sigT) * L).v));
    const sf edgePdf = inMedium ? sigT * T : T;
    const sf ew = T / edgePdf;
    const v3 nvPos = curPos + dir * L;
    sf pdfArea;
    int nvType;
    v3 nvN(0.f, 0.f, 0.f), nvAlbedo(0.f, 0.f, 0.f), nvWeight;
    if (inMedium) {
      nvType = GVPM_PARENT_MEDIUM;
      nvWeight = v3(sigS, sigS, sigS);
      pdfArea = pdfOmega / (L * L);
    } else {
      nvType = GVPM_PARENT_SURFACE;
      nvN = h.n;
      nvAlbedo = h.albedo;
      nvWeight = h.albedo;
      pdfArea = (pdfOmega * sf(fabsf(dot(h.n, dir).v))) / (L * L);
    }
    const v3 prefix = thr;
    const v3 step((curWeight.x * curRr) * ew, (curWeight.y * curRr) * ew, (curWeight.z * curRr) * ew);
    const v3 flux = thr * step;
    const int ni = ci + 1;
    if (inMedium && ni >= firstStored) {
      if (EMIT) {
        const unsigned long long slot = first + appended;
        if (slot < P.n_total && P.aos != nullptr) {
          // the record the gather reads (gvpm_device.cuh: A0 pos, meta | A1 flux, parent pdf | A2 parent, edge pdf |
          // A3 predecessor, rr | A4 parent normal | A5 prefix flux | A6 parent albedo)
          float4 *r = P.aos + slot * 8;
          const v3 pred = ni >= 3 ? prevPos : v3(1.f, 1.f, 1.f);
          r[0] = make_float4(nvPos.x.v, nvPos.y.v, nvPos.z.v, __uint_as_float(pack_meta((uint32_t)curType, (uint32_t)(ni - 1), pid)));
          r[1] = make_float4(flux.x.v, flux.y.v, flux.z.v, pdfArea.v);
          r[2] = make_float4(curPos.x.v, curPos.y.v, curPos.z.v, edgePdf.v);
          r[3] = make_float4(pred.x.v, pred.y.v, pred.z.v, curRr.v);
          r[4] = make_float4(curN.x.v, curN.y.v, curN.z.v, 0.f);
          r[5] = make_float4(prefix.x.v, prefix.y.v, prefix.z.v, 0.f);
          r[6] = make_float4(curAlbedo.x.v, curAlbedo.y.v, curAlbedo.z.v, 0.f);
        } else if (slot < P.n_total) {
          put3(P.pos, slot, nvPos); put3(P.flux, slot, flux); put3(P.parent_pos, slot, curPos);
          put3(P.pred_pos, slot, ni >= 3 ? prevPos : v3(1.f, 1.f, 1.f));
          put3(P.parent_n, slot, curN); put3(P.prefix_flux, slot, prefix); put3(P.parent_albedo, slot, curAlbedo);
          P.parent_pdf[slot] = pdfArea.v; P.edge_pdf[slot] = edgePdf.v; P.rr_weight[slot] = curRr.v;
          P.parent_type[slot] = (uint8_t)curType; P.depth[slot] = (uint8_t)(ni - 1); P.path_id[slot] = pid;
        }
      }
      ++appended;
    }
    thr = flux;
    prevPos = curPos;
    curPos = nvPos; curN = nvN; curAlbedo = nvAlbedo; curWeight = nvWeight; curType = nvType; curRr = sf(1.f);
    ci = ni;
    if (ci >= P.max_depth) break;   // path length bound
    // russian roulette before sampling the next direction (vertex.cpp:291-302)
    if (ci - 1 >= P.rr_depth) {
      const sf m = smax(thr.x * curWeight.x, smax(thr.y * curWeight.y, thr.z * curWeight.z));
      const sf q(fminf(m.v, 0.95f));
      if (!(rng.uniform() < q.v)) break;
      curRr = sf(1.0f) / q;
    }
    const v3 inDir = dir;
    if (curType == GVPM_PARENT_MEDIUM) {
      if (P.phase_type == GVPM_PHASE_HG && fabsf(P.hg_g) > 1e-4f) {
        const sf g(P.hg_g);
        const float u = rng.uniform();
        const sf sq = (sf(1.f) - g * g) / ((sf(1.f) - g) + (sf(2.f) * g) * sf(u));
        const sf ct = ((sf(1.f) + g * g) - sq * sq) / (sf(2.f) * g);
        const sf st = safe_sqrt(sf(1.f) - ct * ct);
        float sn, cs;
        pm_sincos2pi(rng.uniform(), sn, cs);
        dir = to_world(inDir, v3(st * sf(cs), st * sf(sn), ct));
        pdfOmega = hg_eval(g, -dot(inDir, dir));   // wi = -inDir
      } else {
        const float u1 = rng.uniform(), u2 = rng.uniform();
        dir = uniform_sphere(u1, u2);
        pdfOmega = sf(GVPM_INV_FOURPI);
      }
    } else {
      const float u1 = rng.uniform(), u2 = rng.uniform();
      dir = to_world(curN, cosine_hemisphere(u1, u2));
      pdfOmega = smax(sf(0.f), dot(dir, curN)) * sf(GVPM_INV_PI);
      if (pdfOmega.v <= 0.f) break;
    }
    dir = unit(dir);
  }
  return appended;
}

// pass 1: photons per path; block totals of (photons, contributing paths) packed as (paths << 40 | photons)
__global__ void __launch_bounds__(256) k_trace_count(const __grid_constant__ TraceParams P, uint8_t *__restrict__ counts,
                                                      unsigned long long *__restrict__ block_tot) {
  __shared__ unsigned long long ws[8];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = 0;
  if (k < P.n_paths) {
    c = walk_path<false>(P, P.path_base + k, 0ull, 0u);
    counts[k] = (uint8_t)c;
  }
  unsigned long long v = (unsigned long long)c | ((unsigned long long)(c ? 1u : 0u) << 40);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += ws[i];
    block_tot[blockIdx.x] = t;
  }
}
// exclusive scan of the block totals by one block; total -> total_out[0]
__global__ void __launch_bounds__(1024) k_trace_scan(unsigned long long *__restrict__ block_tot, uint32_t nb,
                                                      unsigned long long *__restrict__ total_out) {
  __shared__ unsigned long long ws[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = i < nb ? block_tot[i] : 0ull;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
      const unsigned long long x = ws[lane];
      unsigned long long xi = x;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= o) xi += u;
      }
      ws[lane] = xi - x;
    }
    __syncthreads();
    const unsigned long long excl = carry + ws[w] + (inc - v);
    if (i < nb) block_tot[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) total_out[0] = carry;
}
// pass 2: the same walks, stored.  last_path[0] = 1 + index of the path that stored photon n_total - 1.
__global__ void __launch_bounds__(256) k_trace_emit(const __grid_constant__ TraceParams P, const uint8_t *__restrict__ counts,
                                                     const unsigned long long *__restrict__ block_off,
                                                     unsigned long long *__restrict__ last_path) {
  __shared__ unsigned long long ws[8];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t c = k < P.n_paths ? counts[k] : 0u;
  const unsigned long long v = (unsigned long long)c | ((unsigned long long)(c ? 1u : 0u) << 40);
  unsigned long long inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  unsigned long long excl = block_off[blockIdx.x] + (inc - v);
  for (int i = 0; i < w; ++i) excl += ws[i];
  if (c == 0) return;
  const unsigned long long first = P.slot_base + (excl & ((1ull << 40) - 1ull));
  const uint32_t pid = P.path_id_base + (uint32_t)(excl >> 40);
  if (first >= P.n_total) return;
  walk_path<true>(P, P.path_base + k, first, pid);
  if (first + c >= P.n_total) last_path[0] = P.path_base + k + 1ull;   // exactly one path crosses the end of the set
}

// ---- launchers ------------------------------------------------------------------------------------------------------
void launch_generate_rays(const RayGenParams &P, cudaStream_t st) {
  const int bs = P.block;
  const int blocksX = (P.cam.film_w + bs - 1) / bs, blocksY = (P.y1 - P.y0 + bs - 1) / bs;
  if (blocksX > 0 && blocksY > 0) k_generate_rays<<<blocksX * blocksY, 1024, 0, st>>>(P);
}
// counts: [n_paths] bytes; block_tot: [(n_paths + 255) / 256 + 1] words; totals: [0] batch total (paths << 40 | photons)
void launch_trace_count(const TraceParams &P, uint8_t *counts, unsigned long long *block_tot, unsigned long long *totals,
                        cudaStream_t st) {
  const uint32_t nb = (P.n_paths + 255) / 256;
  if (nb) k_trace_count<<<nb, 256, 0, st>>>(P, counts, block_tot);
  k_trace_scan<<<1, 1024, 0, st>>>(block_tot, nb, totals);
}
void launch_trace_emit(const TraceParams &P, const uint8_t *counts, const unsigned long long *block_off,
                       unsigned long long *last_path, cudaStream_t st) {
  const uint32_t nb = (P.n_paths + 255) / 256;
  if (nb) k_trace_emit<<<nb, 256, 0, st>>>(P, counts, block_off, last_path);
}

}  // namespace gvpm
