// gather_bre.cu — G-BRE gather: primal + 4 offset-path gradient contributions per camera-ray
// medium segment.  Replaces, for all gather points of an iteration,
//   GradientBeamRadianceEstimator::query            gvpm/gvpm_accel.h:268-312
//   VolumeGradientBREQuery::operator()              gvpm/shift/shift_volume_photon.cpp:658-856
//   shiftNull / shiftPhotonDiffuse / getShiftPos    shift_volume_photon.cpp:119-158,382-486,858-896
//   diffuseReconnection                             gvpm/shift/operation/shift_diffuse.cpp:11-134
//   HomogeneousMedium::eval, phase eval             medium/homogeneous.cpp:432-513, phase/hg.cpp:107
//
// Mapping: one warp per camera ray.  The warp walks the implicit 32-ary AABB
// hierarchy over the Morton-sorted photons without a node stack: one ballot mask per level is
// the whole traversal state.  Lane c tests child c of the current node (two coalesced 128-bit
// loads per lane); at a leaf lane c tests photon c with the reference's predicate in strictly
// rounded arithmetic.  Contributing photons are compacted into a per-warp shared-memory queue
// and evaluated 32 at a time (one photon per lane, 7 x 128-bit record loads), so the divergent
// shift code runs with full lanes; each lane keeps its own 27 accumulators, reduced across the
// warp with shuffles once per ray.
#include "gvpm_device.cuh"

namespace gvpm {

#ifndef GVPM_BRE_WARPS
#define GVPM_BRE_WARPS 4
#endif
#ifndef GVPM_BRE_MIN_BLOCKS
#define GVPM_BRE_MIN_BLOCKS 4
#endif
constexpr int kWarpsPerBlock = GVPM_BRE_WARPS;
constexpr int kQueue = 64;

struct WarpShared {
  float4 ray[GVPM_RAY_FLOAT4];  // 320 B
  uint32_t queue[kQueue];       // 256 B
  float acc[GVPM_OUT_FLOATS * 32];  // 3456 B: [27][32], column = lane
  uint32_t mask[GVPM_MAX_LEVELS];
  uint32_t base[GVPM_MAX_LEVELS];
};

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// conservative slab test of an (already inflated) box against the ray interval [tlo, thi]
__device__ __forceinline__ bool box_hit(const Tree &t, uint32_t i, float oxp, float oxm, float oyp, float oym,
                                        float ozp, float ozm, float ix, float iy, float iz, float tlo,
                                        float thi) {
  float4 lo = ldg4(t.lo + i), hi = ldg4(t.hi + i);
  // (lo - pad) - o == lo - (o + pad): the pad is folded into the origin
  float t1 = (lo.x - oxp) * ix, t2 = (hi.x - oxm) * ix;
  float tn = fminf(t1, t2), tf = fmaxf(t1, t2);
  t1 = (lo.y - oyp) * iy; t2 = (hi.y - oym) * iy;
  tn = fmaxf(tn, fminf(t1, t2)); tf = fminf(tf, fmaxf(t1, t2));
  t1 = (lo.z - ozp) * iz; t2 = (hi.z - ozm) * iz;
  tn = fmaxf(tn, fminf(t1, t2)); tf = fminf(tf, fmaxf(t1, t2));
  return tn <= tf && tf >= tlo && tn <= thi;
}

// 1/max(2*deltaT, 0.0001) as the reference evaluates it in double then rounds to Float
// (shift_volume_photon.cpp:723): identical to this fp32 form (test_chord_pdf).
__device__ __forceinline__ sf chord_pdf(sf deltaT) {
  sf x2 = deltaT * sf(2.f);
  return (x2.v <= 0.0001f) ? sf(10000.f) : sf(1.f) / x2;
}

struct MediumRec { sf T, pdfSuccess; };

// HomogeneousMedium::eval with equal sigma_t over channels (enforced at gvpm_set_medium as the
// reference does, homogeneous.cpp:188-201): transmittance is one scalar.
__device__ __forceinline__ MediumRec medium_eval(const GatherParams &P, sf mint, sf maxt) {
  MediumRec r;
  sf distance = maxt - mint;
  sf st(P.sigma_t[0]);
  sf tmp(expf(((-st) * distance).v));
  sf ps = st * tmp;
  ps = ((ps + ps) + ps) / sf(3.f);
  r.pdfSuccess = ps * sf(P.sampling_weight);
  r.T = tmp;
  if (r.T.v < 1e-20f) r.T = sf(0.f);
  return r;
}

__device__ __forceinline__ sf phase_eval(const GatherParams &P, v3 wi, v3 wo) {
  if (P.phase_type == GVPM_PHASE_ISOTROPIC) return sf(GVPM_INV_FOURPI);
  sf g(P.hg_g);
  sf temp = sf(1.f) + g * g + sf(2.f) * g * dot(wi, wo);
  return sf(GVPM_INV_FOURPI) * (sf(1.f) - g * g) / (temp * ssqrt(temp));
}

// Triangle::rayIntersect (include/mitsuba/core/triangle.h:109-145) any-hit over the occluder list,
// preceded by a conservative plane-distance cull (|d| = 1 so t >= distance to the plane).
__device__ __forceinline__ bool occluded(const GatherParams &P, v3 o, v3 d, sf mint, sf maxt) {
  if (maxt < mint) return false;
  for (uint32_t t = 0; t < P.n_tri; ++t) {
    float4 pl = ldg4(P.tri_plane + t);
    float dist = fabsf(pl.x * o.x.v + pl.y * o.y.v + pl.z * o.z.v + pl.w);
    if (dist > maxt.v * 1.001f + 1e-5f) continue;
    const float *tv = P.tri + 9 * t;
    v3 p0(__ldg(tv), __ldg(tv + 1), __ldg(tv + 2)), p1(__ldg(tv + 3), __ldg(tv + 4), __ldg(tv + 5)),
        p2(__ldg(tv + 6), __ldg(tv + 7), __ldg(tv + 8));
    v3 edge1 = p1 - p0, edge2 = p2 - p0;
    v3 pvec = cross(d, edge2);
    sf det = dot(edge1, pvec);
    if (det.v == 0.f) continue;
    sf inv_det = sf(1.f) / det;
    v3 tvec = o - p0;
    sf u = dot(tvec, pvec) * inv_det;
    if (u.v < 0.f || u.v > 1.f) continue;
    v3 qvec = cross(tvec, edge1);
    sf v = dot(d, qvec) * inv_det;
    if (v.v >= 0.f && (u + v).v <= 1.f) {
      sf tt = dot(edge2, qvec) * inv_det;
      if (tt >= mint && tt <= maxt) return true;
    }
  }
  return false;
}

// coordinateSystemCoherent, src/libcore/util.cpp:592-599
__device__ __forceinline__ void coherent_frame(v3 n, v3 &b1, v3 &b2) {
  const sf sign(copysignf(1.0f, n.z.v));
  const sf a = sf(-1.0f) / (sign + n.z);
  const sf b = n.x * n.y * a;
  b1 = v3(sf(1.0f) + sign * n.x * n.x * a, sign * b, -sign * n.x);
  b2 = v3(b, sign + n.y * n.y * a, -n.y);
}

struct BaseRay {
  v3 o, d, eye;
  sf mint, maxt, edgeLen, xi;
  int px, py, edgeId;
};

// kernel-chord sampling of the 3-D kernel (shift_volume_photon.cpp:707-724).  Returns false when
// the photon is outside the geometric neighbour set.
__device__ __forceinline__ bool base_distance(const GatherParams &P, const BaseRay &R, v3 p, sf &tBase,
                                              sf &pdfCam) {
  v3 oc = p - R.o;
  sf dd = dot(oc, R.d);
  sf distSqr = length_sq((R.o + dd * R.d) - p);
  if (!(dd > R.mint && distSqr < sf(P.radius_sq))) return false;  // gvpm_accel.h:297-301
  if (P.cfg.kernel_3d) {
    sf r(P.radius);
    sf deltaT = safe_sqrt(r * r - distSqr);
    sf tminKernel = dd - deltaT;
    sf tRand = tminKernel + (deltaT * sf(2.f)) * R.xi;
    if (tRand < R.mint || tRand > R.edgeLen) return false;
    tBase = tRand;
    pdfCam = chord_pdf(deltaT);
  } else {
    if (dd > R.edgeLen) return false;  // explicit bound, DESIGN.md §6 (bre.cpp:240-242)
    tBase = dd;
    pdfCam = sf(1.f);
  }
  return true;
}

__device__ __forceinline__ bool filters_pass(const GatherParams &P, const BaseRay &R, uint32_t meta) {
  int type = meta & 3, depth = (meta >> 2) & 255, parity = (meta >> 10) & 1;
  int pathLen = depth + R.edgeId;
  if (P.cfg.max_depth > 0 && pathLen > P.cfg.max_depth) return false;
  if (P.cfg.min_depth != 0 && pathLen < P.cfg.min_depth) return false;
  int m = P.cfg.lighting_mode;
  if (!((m & GVPM_SURF2MEDIA) && (m & GVPM_MEDIA2MEDIA))) {
    if (type == GVPM_PARENT_MEDIUM && !(m & GVPM_MEDIA2MEDIA)) return false;
    if (type != GVPM_PARENT_MEDIUM && !(m & GVPM_SURF2MEDIA)) return false;
  }
  if (P.cfg.path_set && parity != ((R.px + R.py) % 2)) return false;
  return true;
}

// per-warp accumulators live in shared memory, one conflict-free column per lane: [27][32]
__device__ __forceinline__ void acc_add(float *A, int j, v3 c) {
  A[(3 * j) * 32] += c.x.v; A[(3 * j + 1) * 32] += c.y.v; A[(3 * j + 2) * 32] += c.z.v;
}

// One contributing photon: VolumeGradientBREQuery::operator() after the filters.
__device__ __forceinline__ BaseRay load_base_ray(const float4 *sray) {
  BaseRay R;
  const float4 b0 = sray[0], b1 = sray[1], b2 = sray[2], b3 = sray[3];
  R.o = v3(b0.x, b0.y, b0.z); R.mint = sf(b0.w);
  R.d = v3(b1.x, b1.y, b1.z); R.maxt = sf(b1.w);
  R.eye = v3(b2.x, b2.y, b2.z); R.edgeLen = sf(b2.w);
  R.xi = sf(b3.x);
  R.px = (int)__float_as_uint(b3.y); R.py = (int)__float_as_uint(b3.z);
  R.edgeId = (int)__float_as_uint(b3.w);
  return R;
}

// Out of line on purpose: the shift code needs ~100 registers, the traversal loop ~40; keeping
// them in separate register frames leaves the hot loop spill-free.
__device__ __noinline__ void bre_photon(const GatherParams *Pp, const float4 *sray, uint32_t pi, float *A) {
  const GatherParams &P = *Pp;
  const BaseRay R = load_base_ray(sray);
  const uint32_t n = P.tree.n;
  const float4 q0 = ldg4(P.planes + pi);
  const float4 q1 = ldg4(P.planes + (size_t)n + pi);
  const float4 q2 = ldg4(P.planes + 2 * (size_t)n + pi);
  const float4 q3 = ldg4(P.planes + 3 * (size_t)n + pi);
  const float4 q4 = ldg4(P.planes + 4 * (size_t)n + pi);
  const float4 q5 = ldg4(P.planes + 5 * (size_t)n + pi);
  const float4 q6 = ldg4(P.planes + 6 * (size_t)n + pi);
  const v3 p(q0.x, q0.y, q0.z), flux(q1.x, q1.y, q1.z), parent(q2.x, q2.y, q2.z), pred(q3.x, q3.y, q3.z),
      pn(q4.x, q4.y, q4.z), prefix(q5.x, q5.y, q5.z), albedo(q6.x, q6.y, q6.z);
  const sf parentPdf(q1.w), edgePdf(q2.w), rrW(q3.w);
  const int ptype = __float_as_uint(q0.w) & 3;
  const sf r(P.radius), rr2 = r * r;
  const v3 sigS(P.sigma_s[0], P.sigma_s[1], P.sigma_s[2]);

  sf tBase, pdfCam;
  if (!base_distance(P, R, p, tBase, pdfCam)) return;  // cannot happen for a queued photon
  const sf rrG = P.cfg.path_set ? sf(2.f) : sf(1.f);
  const v3 wi = normalize(parent - p);
  const MediumRec mBase = medium_eval(P, R.mint, tBase);
  const v3 contrib = (sigS * flux) * phase_eval(P, wi, -R.d);
  const v3 baseContrib = (contrib * mBase.T) * R.eye;
  const sf norm = sf(P.kernel_vol) * pdfCam;
  const sf recip = sf(1.f) / norm;
  acc_add(A, 0, (baseContrib * recip) * rrG);

  const MediumRec mShift = medium_eval(P, sf(P.cfg.epsilon), tBase);
  const v3 zBase = R.o + tBase * R.d;

#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const float4 s0 = sray[4 * (k + 1)], s1 = sray[4 * (k + 1) + 1], s2 = sray[4 * (k + 1) + 2];
    sf weight(1.f);
    v3 S(0.f, 0.f, 0.f);
    if (__float_as_uint(s2.w) != 0u) {  // validVolumeEdge, shift_cameraPath.h:135-140
      const v3 ok(s0.x, s0.y, s0.z), dk(s1.x, s1.y, s1.z), eyeK(s2.x, s2.y, s2.z);
      const sf lenK(s0.w), sensor(s1.w);
      const v3 zShift = ok + tBase * dk;
      bool done = false;
      if (P.cfg.use_shift_null && P.cfg.kernel_3d) {  // :776-802
        sf ZPtoY = length_sq(zShift - p);
        if (ZPtoY < rr2 && tBase < lenK) {
          sf dd = dot(p - ok, dk);
          sf ds = length_sq((ok + dd * dk) - p);
          sf pdfShift = chord_pdf(safe_sqrt(rr2 - ds));
          // shiftNull, :119-158
          v3 c = (sigS * flux) * phase_eval(P, wi, -dk);
          S = (c * mShift.T) * eyeK;
          weight = sf(0.5f);
          if (P.cfg.use_mis) {
            if (pdfShift.v == 0.f || pdfCam.v == 0.f) weight = sf(1.f);
            else weight = sf(1.f) / (sf(1.f) + sensor * pdfShift / pdfCam);
          }
          done = true;
        }
      }
      if (!done && lenK >= tBase && ptype != GVPM_PARENT_OTHER) {  // :809-838
        // getShiftPos, :858-896
        v3 offsetPos = zShift + (p - zBase);
        if (!P.cfg.kernel_3d) {  // coherent frames for the 2-D kernel, :866-873
          v3 bs, bt, ns, nt;
          coherent_frame(R.d, bs, bt);
          coherent_frame(dk, ns, nt);
          const v3 v = p - zBase;
          const v3 local(dot(v, bs), dot(v, bt), dot(v, R.d));
          offsetPos = zShift + ((ns * local.x + nt * local.y) + dk * local.z);
        }
        if (P.cfg.use_shift_null) {
          sf offDistSqr = length_sq(zBase - offsetPos);
          if (offDistSqr < rr2) {
            v3 dShift = zShift - zBase;
            dShift = dShift / length(dShift);
            sf cosD = dot(dShift, -(offsetPos - zShift));
            offsetPos = offsetPos + (dShift * cosD) * sf(2.f);
          }
        }
        sf pdfShift(1.f);
        if (P.cfg.kernel_3d) {
          sf dd = dot(offsetPos - ok, dk);
          sf ds = length_sq((ok + dd * dk) - offsetPos);
          pdfShift = chord_pdf(safe_sqrt(rr2 - ds));
        }
        // shiftPhotonDiffuse, :382-486
        v3 dProj = offsetPos - parent;
        sf lProj = length(dProj);
        dProj = dProj / lProj;
        bool ok2 = !occluded(P, parent, dProj, sf(P.cfg.epsilon), lProj * sf(P.cfg.shadow_maxt_scale));
        if (ok2 && ptype != GVPM_PARENT_MEDIUM) {
          v3 edgeD = normalize(p - parent);
          sf signDot = dot(pn, dProj) / dot(pn, edgeD);
          if (signDot.v < 0.f) ok2 = false;
        }
        if (ok2) {
          // diffuseReconnection, shift_diffuse.cpp:11-134
          v3 thr(1.f, 1.f, 1.f);
          sf pdfValue(0.f);
          bool early = false;
          if (ptype == GVPM_PARENT_SURFACE) {
            v3 wiW = normalize(pred - parent);
            sf cosI = dot(pn, wiW), cosO = dot(pn, dProj);
            if (cosI.v <= 0.f || cosO.v <= 0.f) {
              thr = v3(0.f, 0.f, 0.f);
            } else {
              thr = thr * (albedo * (sf(GVPM_INV_PI) * cosO));
              pdfValue = sf(GVPM_INV_PI) * cosO;
            }
            if ((cosI * cosI).v <= 0.f || (cosO * cosO).v <= 0.f) early = true;
          } else if (ptype == GVPM_PARENT_MEDIUM) {
            v3 pWi = normalize(pred - parent);
            sf phv = phase_eval(P, pWi, dProj);
            thr = thr * (sigS * phv);
            pdfValue = phv;
          } else {  // emitter sample, emitters/area.cpp:132-150
            sf dp = dot(dProj, pn);
            if (dp.v < 0.f) dp = sf(0.f);
            sf e = sf(GVPM_INV_PI) * dp;
            thr = thr * v3(e, e, e);
            pdfValue = e;
          }
          sf sPdf(0.f);
          if (!early) {
            sf GOp = sf(1.f) / (lProj * lProj);
            sPdf = pdfValue * GOp;
            thr = thr * GOp;
            if (parentPdf.v == 0.f) {
              sPdf = sf(0.f);
            } else {
              thr = thr / parentPdf;
              thr = thr * rrW;
              MediumRec mr = medium_eval(P, sf(0.f), lProj);
              sPdf = sPdf * mr.pdfSuccess;
              sf te = mr.T * (sf(1.f) / edgePdf);  // Spectrum / Float = * (1/f), spectrum.h:415-425
              thr = thr * te;
            }
          }
          if (sPdf.v == 0.f) {
            weight = sf(1.f);
          } else {
            v3 photonWeight = prefix * thr;
            v3 c = (sigS * photonWeight) * phase_eval(P, -dProj, -dk);
            S = (c * mShift.T) * eyeK;
            weight = sf(0.5f);
            if (P.cfg.use_mis) {
              sf basePdf = pdfCam;
              basePdf = basePdf * parentPdf;
              basePdf = basePdf * edgePdf;
              sf offsetPdf = sPdf * pdfShift;
              if (offsetPdf.v == 0.f || basePdf.v == 0.f) {
                weight = sf(1.f);
              } else {
                sf q = sensor * (offsetPdf / basePdf);
                weight = P.cfg.power_heuristic ? sf(1.f) / (sf(1.f) + q * q) : sf(1.f) / (sf(1.f) + q);
              }
            }
          }
        }
      }
    }
    if ((k == 1 && R.px == P.cfg.film_w - 1) || (k == 2 && R.py == P.cfg.film_h - 1)) weight = sf(1.f);
    const sf rw = rrG * weight;
    acc_add(A, 5 + k, (baseContrib * rw) * recip);
    acc_add(A, 1 + k, (S * rw) * recip);
  }
}

template <bool DUMP>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, GVPM_BRE_MIN_BLOCKS) k_gather_bre(const __grid_constant__ GatherParams P) {
  __shared__ WarpShared sh[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  WarpShared &S = sh[w];
  const Tree &T = P.tree;
  const int top = T.levels - 1;

  for (;;) {
    uint32_t ray = 0;
    if (lane == 0) ray = atomicAdd(P.work_counter, 1u);
    ray = __shfl_sync(0xffffffffu, ray, 0);
    if (ray >= P.n_rays) break;
    __syncwarp();
    if (lane < GVPM_RAY_FLOAT4) S.ray[lane] = ldg4(P.rays + (size_t)ray * GVPM_RAY_FLOAT4 + lane);
    __syncwarp();
    const BaseRay R = load_base_ray(S.ray);
    float *A = S.acc + lane;
#pragma unroll
    for (int j = 0; j < GVPM_OUT_FLOATS; ++j) A[j * 32] = 0.f;
    uint32_t nGeom = 0, nContrib = 0, qn = 0;
    uint64_t dumpBase = 0;
    if (DUMP) dumpBase = P.nbr_offsets[ray];

    if (T.n > 0 && R.edgeLen.v >= R.mint.v) {
      // conservative culling: the slab test runs in relaxed arithmetic against boxes inflated by
      // the radius; `pad` absorbs the rounding of both the slab test and the strict predicate
      // (a few ulp of the coordinate magnitudes), folded into the origin so it costs nothing.
      const float mag = fmaxf(fmaxf(fabsf(R.o.x.v), fabsf(R.o.y.v)), fabsf(R.o.z.v)) + __ldg(P.bounds + 6) +
                        fabsf(R.edgeLen.v) + P.radius;
      const float pad = mag * 3.8147e-6f;  // 2^-18
      const float oxp = R.o.x.v + pad, oxm = R.o.x.v - pad, oyp = R.o.y.v + pad, oym = R.o.y.v - pad,
                  ozp = R.o.z.v + pad, ozm = R.o.z.v - pad;
      const float ix = 1.f / R.d.x.v, iy = 1.f / R.d.y.v, iz = 1.f / R.d.z.v;
      const float tlo = R.mint.v - 4.f * pad;
      const float thi = R.edgeLen.v + P.radius + 4.f * pad;
      const float rpad2 = (P.radius + pad) * (P.radius + pad);

      uint32_t cur, base = 0;
      int l = top;
      cur = __ballot_sync(0xffffffffu, (uint32_t)lane < T.cnt[top] &&
                                           box_hit(T, T.off[top] + lane, oxp, oxm, oyp, oym, ozp, ozm, ix, iy,
                                                   iz, tlo, thi));
      for (;;) {
        bool flush = false;
        if (cur == 0) {
          if (l == top) {
            flush = true;
          } else {
            ++l;
            cur = S.mask[l];
            base = S.base[l];
            continue;
          }
        }
        if (!flush) {
          const int c = __ffs(cur) - 1;
          cur &= cur - 1;
          const uint32_t node = base + c;
          if (l > 0) {
            S.mask[l] = cur;
            S.base[l] = base;
            --l;
            base = node << 5;
            const uint32_t idx = base + lane;
            cur = __ballot_sync(0xffffffffu, idx < T.cnt[l] && box_hit(T, T.off[l] + idx, oxp, oxm, oyp, oym,
                                                                         ozp, ozm, ix, iy, iz, tlo, thi));
            continue;
          }
          // ---- leaf: lane tests photon (node*32 + lane) with the reference predicate ----
          const uint32_t pi = (node << 5) + lane;
          bool geom = false, contrib = false;
          float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f);
          bool cand = false;
          if (pi < T.n) {
            q0 = ldg4(P.planes + pi);
            // relaxed (FMA) pre-test, conservative by `pad`: only candidates pay for the strictly
            // rounded predicate below
            const float cx = q0.x - R.o.x.v, cy = q0.y - R.o.y.v, cz = q0.z - R.o.z.v;
            const float dd = cx * R.d.x.v + cy * R.d.y.v + cz * R.d.z.v;
            const float qx = cx - dd * R.d.x.v, qy = cy - dd * R.d.y.v, qz = cz - dd * R.d.z.v;
            cand = (qx * qx + qy * qy + qz * qz) < rpad2 && dd > tlo;
          }
          if (!__any_sync(0xffffffffu, cand)) continue;
          if (cand) {
            sf tB, pc;
            geom = base_distance(P, R, v3(q0.x, q0.y, q0.z), tB, pc);
            contrib = geom && filters_pass(P, R, __float_as_uint(q0.w));
          }
          const uint32_t gm = __ballot_sync(0xffffffffu, geom), cm = __ballot_sync(0xffffffffu, contrib);
          if (DUMP) {
            if (geom) {
              const uint32_t rank = __popc(gm & ((1u << lane) - 1u));
              P.nbr_idx[dumpBase + nGeom + rank] = P.orig[pi] | (contrib ? 0x80000000u : 0u);
            }
          }
          nGeom += __popc(gm);
          nContrib += __popc(cm);
          if (DUMP || cm == 0) continue;
          if (contrib) S.queue[qn + __popc(cm & ((1u << lane) - 1u))] = pi;
          qn += __popc(cm);
          __syncwarp();
          if (qn < 32) continue;
        }
        // ---- evaluate up to 32 queued photons, one per lane ----
        if (!DUMP && qn > 0) {
          const uint32_t take = qn < 32u ? qn : 32u;
          qn -= take;
          uint32_t mine = 0;
          if ((uint32_t)lane < take) mine = S.queue[qn + lane];
          __syncwarp();
#ifndef GVPM_EXP_SKIP_SHADE
          if ((uint32_t)lane < take) bre_photon(&P, S.ray, mine, A);
#endif
          __syncwarp();
        }
        if (flush) break;
      }
    }

    if (!DUMP) {
      // warp-level reduction of the per-lane accumulators: lane j < 27 sums row j (32 columns,
      // rotated start so the 27 lanes hit distinct banks), one coalesced 108-byte store per ray
      __syncwarp();
      if (lane < GVPM_OUT_FLOATS) {
        const float *row = S.acc + lane * 32;
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) v += row[(c + lane) & 31];
        P.out[(size_t)ray * GVPM_OUT_FLOATS + lane] = v;
      }
      __syncwarp();
    }
    if (P.counts && lane == 0) {
      P.counts[2 * (size_t)ray] = nGeom;
      P.counts[2 * (size_t)ray + 1] = nContrib;
    }
  }
}

// host-side launcher (called from gvpm_capi.cu)
cudaError_t launch_gather_bre(const GatherParams &P, bool dump, int sm_count, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(P.work_counter, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  if (P.n_rays == 0) return cudaSuccess;
  int blocksPerSm = 0;
  if (dump)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, k_gather_bre<true>, kWarpsPerBlock * 32, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, k_gather_bre<false>, kWarpsPerBlock * 32, 0);
  if (blocksPerSm < 1) blocksPerSm = 1;
  // persistent grid: a whole number of resident CTAs per SM, warps pull rays from a counter
  unsigned grid = (unsigned)(sm_count * blocksPerSm);
  unsigned need = (P.n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (grid > need) grid = need;
  if (dump)
    k_gather_bre<true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(P);
  else
    k_gather_bre<false><<<grid, kWarpsPerBlock * 32, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace gvpm
