// gvpm_oracle_trace.cpp — TEST INFRASTRUCTURE (see gvpm_oracle.hpp).  CPU restatement of the on-device generators of
// rows f-1 / f-2 (include/gvpm_b200.h: gvpm_trace_photons, gvpm_generate_rays) for the box scene class: the light-path
// random walk with libbidir's bookkeeping (gvpm/gvpm_proc.cpp:125-210, gvpm_accel.h:119-199) and the
// pinhole gather rays with their four offsets (gvpm/gvpm_gatherpoint.h:259-486, shift_utilities.h:255-261).
// The specification both sides implement: PCG32(seed, path / pixel index) uniforms; plain IEEE fp32 with every
// operation rounded on its own (this file is compiled with -ffp-contract=off); log / exp / sin / cos through the
// fixed polynomial routines below (no libm), so the device records can be compared BIT FOR BIT
// (tests/test_gpu_generate.py).  Sequential: paths 0, 1, 2, ... until n photons are stored.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/gvpm_b200.h"

namespace {

struct Rng {  // PCG32 (O'Neill), stream = index
  uint64_t state, inc;
  Rng(uint64_t seed, uint64_t seq) : state(0), inc((seq << 1u) | 1u) {
    next();
    state += seed;
    next();
  }
  uint32_t next() {
    const uint64_t old = state;
    state = old * 6364136223846793005ULL + inc;
    const uint32_t x = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
    return (x >> rot) | (x << ((0u - rot) & 31u));
  }
  float uniform() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
};

inline float asFloat(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }
inline uint32_t asBits(float f) { uint32_t b; std::memcpy(&b, &f, 4); return b; }

// log x = 2 atanh((m - 1)/(m + 1)) + e ln 2 with m in [1/sqrt2, sqrt2)
float pmLog(float x) {
  const uint32_t bits = asBits(x);
  int e = (int)(bits >> 23) - 127;
  float m = asFloat((bits & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
  const float s = (m - 1.f) / (m + 1.f), s2 = s * s;
  float p = 0.11111111f;
  p = p * s2 + 0.14285715f;
  p = p * s2 + 0.2f;
  p = p * s2 + 0.33333334f;
  p = p * s2 + 1.0f;
  return (2.f * s) * p + (float)e * 0.69314718f;
}
// exp x = 2^n exp(r), n = floor(x / ln 2 + 1/2), r = x - n ln 2 in two steps
float pmExp(float x) {
  if (x < -87.f) return 0.f;
  const float n = std::floor(x * 1.44269504f + 0.5f);
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float p = 1.0f / 720.0f;
  p = p * r + 1.0f / 120.0f;
  p = p * r + 1.0f / 24.0f;
  p = p * r + 1.0f / 6.0f;
  p = p * r + 0.5f;
  p = p * r + 1.0f;
  p = p * r + 1.0f;
  return asFloat((uint32_t)((int)n + 127) << 23) * p;
}
// sin, cos of 2 pi u: quadrant q = floor(4u + 1/2), residual angle in [-pi/4, pi/4]
void pmSinCos2Pi(float u, float &sn, float &cs) {
  const float q = std::floor(u * 4.f + 0.5f);
  const float a = (u - q * 0.25f) * 6.2831855f, a2 = a * a;
  float p = -1.9841270e-4f;
  p = p * a2 + 8.3333333e-3f;
  p = p * a2 + -0.16666667f;
  p = p * a2 + 1.0f;
  const float s = a * p;
  float c = 2.4801587e-5f;
  c = c * a2 + -1.3888889e-3f;
  c = c * a2 + 4.1666668e-2f;
  c = c * a2 + -0.5f;
  c = c * a2 + 1.0f;
  switch ((int)q & 3) {
    case 0: sn = s; cs = c; break;
    case 1: sn = c; cs = -s; break;
    case 2: sn = -s; cs = -c; break;
    default: sn = -c; cs = s; break;
  }
}

struct V { float x, y, z; };
inline V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V operator-(V a) { return {-a.x, -a.y, -a.z}; }
inline V operator*(V a, float f) { return {a.x * f, a.y * f, a.z * f}; }
inline V mul(V a, V b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V unit(V a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline float axisOf(V a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

struct Hit { float t; V n, albedo; bool escaped; };

// walls x = lo, x = hi, y = lo, y = hi, z = hi, the open face z = lo, then the rectangles; a later surface replaces an
// earlier one only when it is strictly nearer
Hit intersect(const gvpm_box_scene &S, V o, V d) {
  Hit h{1e30f, {0, 0, 0}, {0.7f, 0.7f, 0.7f}, false};
  struct Face { int axis; float pos; V n; int alb; bool open; };
  const Face faces[6] = {{0, S.lo[0], {1, 0, 0}, 0, false}, {0, S.hi[0], {-1, 0, 0}, 1, false},
                         {1, S.lo[1], {0, 1, 0}, 2, false}, {1, S.hi[1], {0, -1, 0}, 3, false},
                         {2, S.hi[2], {0, 0, -1}, 4, false}, {2, S.lo[2], {0, 0, 1}, 4, true}};
  for (const Face &f : faces) {
    const float dc = axisOf(d, f.axis);
    if (dc == 0.f) continue;
    const float t = (f.pos - axisOf(o, f.axis)) / dc;
    if (t > 1e-6f && t < h.t) {
      h.t = t;
      h.n = f.n;
      h.escaped = f.open;
      h.albedo = f.open ? V{0, 0, 0} : V{S.face_albedo[f.alb][0], S.face_albedo[f.alb][1], S.face_albedo[f.alb][2]};
    }
  }
  if (d.y != 0.f)
    for (int r = 0; r < S.n_rects; ++r) {
      const auto &R = S.rect[r];
      const float t = (R.y - o.y) / d.y;
      if (!(t > 1e-6f && t < h.t)) continue;
      const float x = o.x + d.x * t, z = o.z + d.z * t;
      if (x < R.x0 || x > R.x1 || z < R.z0 || z > R.z1) continue;
      h.t = t;
      h.n = d.y < 0.f ? V{0, 1, 0} : V{0, -1, 0};
      h.albedo = {R.albedo[0], R.albedo[1], R.albedo[2]};
      h.escaped = false;
    }
  return h;
}

// orthonormal frame around n (tangent from the larger of |n.x|, |n.y|), local -> world
V toWorld(V n, V l) {
  V t;
  if (std::fabs(n.x) > std::fabs(n.y)) {
    const float il = 1.0f / std::sqrt(n.x * n.x + n.z * n.z);
    t = {n.z * il, 0.f, -(n.x * il)};
  } else {
    const float il = 1.0f / std::sqrt(n.y * n.y + n.z * n.z);
    t = {0.f, n.z * il, -(n.y * il)};
  }
  const V s = cross(t, n);
  return (s * l.x + t * l.y) + n * l.z;
}
V cosineHemisphere(float u1, float u2) {
  float sn, cs;
  pmSinCos2Pi(u2, sn, cs);
  const float r = std::sqrt(u1), x = r * cs, y = r * sn;
  return {x, y, std::sqrt(std::max(0.f, (1.f - x * x) - y * y))};
}
V uniformSphere(float u1, float u2) {
  float sn, cs;
  pmSinCos2Pi(u2, sn, cs);
  const float z = 1.f - 2.f * u1, r = std::sqrt(std::max(0.f, 1.f - z * z));
  return {r * cs, r * sn, z};
}
const float INV_PI = 0.31830988618379067154f, INV_FOURPI = 0.07957747154594766788f;
float hgEval(float g, float c) {
  const float temp = (1.0f + g * g) + (2.0f * g) * c;
  return (INV_FOURPI * (1.f - g * g)) / (temp * std::sqrt(temp));
}

struct Sink {  // where the photons of a walk go
  gvpm_photon_soa *out;
  size_t cap, filled;
  uint32_t pathId;
  bool storedAny;
};
void put(const float *dst, size_t i, V v) {
  float *d = const_cast<float *>(dst);
  d[3 * i] = v.x; d[3 * i + 1] = v.y; d[3 * i + 2] = v.z;
}

// one light path; photons beyond the capacity are counted but not stored
uint32_t walk(const gvpm_box_scene &S, const gvpm_medium &med, uint64_t seed, uint64_t pathIdx, int maxDepth, int rrDepth,
              int minDepth, Sink &sink) {
  Rng rng(seed, pathIdx);
  const float sigS = med.sigma_s[0], sigT = med.sigma_s[0] + med.sigma_a[0];
  float u1 = rng.uniform(), u2 = rng.uniform();
  V curPos{S.light_x0 + (S.light_x1 - S.light_x0) * u1, S.light_y, S.light_z0 + (S.light_z1 - S.light_z0) * u2};
  V curN{0, -1, 0}, curAlbedo{0, 0, 0}, curWeight{1, 1, 1}, prevPos{1, 1, 1};
  int curType = GVPM_PARENT_EMITTER, ci = 1;
  float curRr = 1.f;
  u1 = rng.uniform();
  u2 = rng.uniform();
  V dir = toWorld(curN, cosineHemisphere(u1, u2));
  float pdfOmega = std::max(0.f, dot(dir, curN)) * INV_PI;
  V thr{S.light_power, S.light_power, S.light_power};
  const int firstStored = std::max(2, minDepth + 1);
  uint32_t appended = 0;
  for (;;) {
    const Hit h = intersect(S, curPos, dir);
    const float t = -pmLog(1.0f - rng.uniform()) / sigT;
    const bool inMedium = t < h.t;
    const float L = inMedium ? t : h.t;
    if (!inMedium && h.escaped) break;
    const float T = pmExp(-sigT * L);
    const float edgePdf = inMedium ? sigT * T : T, ew = T / edgePdf;
    const V nvPos = curPos + dir * L;
    float pdfArea;
    int nvType;
    V nvN{0, 0, 0}, nvAlbedo{0, 0, 0}, nvWeight;
    if (inMedium) {
      nvType = GVPM_PARENT_MEDIUM;
      nvWeight = {sigS, sigS, sigS};
      pdfArea = pdfOmega / (L * L);
    } else {
      nvType = GVPM_PARENT_SURFACE;
      nvN = h.n;
      nvAlbedo = h.albedo;
      nvWeight = h.albedo;
      pdfArea = (pdfOmega * std::fabs(dot(h.n, dir))) / (L * L);
    }
    const V prefix = thr;
    const V step{(curWeight.x * curRr) * ew, (curWeight.y * curRr) * ew, (curWeight.z * curRr) * ew};
    const V flux = mul(thr, step);
    const int ni = ci + 1;
    if (inMedium && ni >= firstStored) {
      if (sink.filled < sink.cap) {
        const size_t i = sink.filled++;
        const gvpm_photon_soa &o = *sink.out;
        put(o.pos, i, nvPos); put(o.flux, i, flux); put(o.parent_pos, i, curPos);
        put(o.pred_pos, i, ni >= 3 ? prevPos : V{1, 1, 1});
        put(o.parent_n, i, curN); put(o.prefix_flux, i, prefix); put(o.parent_albedo, i, curAlbedo);
        const_cast<float *>(o.parent_pdf)[i] = pdfArea;
        const_cast<float *>(o.edge_pdf)[i] = edgePdf;
        const_cast<float *>(o.rr_weight)[i] = curRr;
        const_cast<uint8_t *>(o.parent_type)[i] = (uint8_t)curType;
        const_cast<uint8_t *>(o.depth)[i] = (uint8_t)(ni - 1);
        const_cast<uint32_t *>(o.path_id)[i] = sink.pathId;
        sink.storedAny = true;
      }
      ++appended;
    }
    thr = flux;
    prevPos = curPos;
    curPos = nvPos; curN = nvN; curAlbedo = nvAlbedo; curWeight = nvWeight; curType = nvType; curRr = 1.f;
    ci = ni;
    if (ci >= maxDepth) break;
    if (ci - 1 >= rrDepth) {
      const float m = std::max(thr.x * curWeight.x, std::max(thr.y * curWeight.y, thr.z * curWeight.z));
      const float q = std::min(m, 0.95f);
      if (!(rng.uniform() < q)) break;
      curRr = 1.0f / q;
    }
    const V inDir = dir;
    if (curType == GVPM_PARENT_MEDIUM) {
      if (med.phase_type == GVPM_PHASE_HG && std::fabs(med.hg_g) > 1e-4f) {
        const float g = med.hg_g, u = rng.uniform();
        const float sq = (1.f - g * g) / ((1.f - g) + (2.f * g) * u);
        const float ct = ((1.f + g * g) - sq * sq) / (2.f * g);
        const float st = std::sqrt(std::max(0.f, 1.f - ct * ct));
        float sn, cs;
        pmSinCos2Pi(rng.uniform(), sn, cs);
        dir = toWorld(inDir, {st * cs, st * sn, ct});
        pdfOmega = hgEval(g, -dot(inDir, dir));
      } else {
        u1 = rng.uniform();
        u2 = rng.uniform();
        dir = uniformSphere(u1, u2);
        pdfOmega = INV_FOURPI;
      }
    } else {
      u1 = rng.uniform();
      u2 = rng.uniform();
      dir = toWorld(curN, cosineHemisphere(u1, u2));
      pdfOmega = std::max(0.f, dot(dir, curN)) * INV_PI;
      if (pdfOmega <= 0.f) break;
    }
    dir = unit(dir);
  }
  return appended;
}

}  // namespace

extern "C" {

// fills `out` (caller-allocated arrays of n entries) with exactly n photons; returns the number of light paths traced
long long gvpm_oracle_trace_photons(const gvpm_box_scene *scene, const gvpm_medium *med, size_t n, uint64_t seed,
                                    int max_depth, int rr_depth, int min_depth, gvpm_photon_soa *out) {
  if (!scene || !med || !out) return -1;
  Sink sink{out, n, 0, 0, false};
  const int md = max_depth > 0 ? max_depth : 64;
  uint64_t path = 0;
  while (sink.filled < n) {
    sink.storedAny = false;
    walk(*scene, *med, seed, path, md, rr_depth, min_depth, sink);
    if (sink.storedAny) ++sink.pathId;
    ++path;
    if (path > (1ull << 40)) return -1;
  }
  return (long long)path;
}

// the elementary routines alone, for accuracy checks against libm (tests/test_oracle_trace.py)
void gvpm_oracle_pm_functions(size_t n, const float *x, float *logx, float *expmx, float *sin2pi, float *cos2pi) {
  for (size_t i = 0; i < n; ++i) {
    logx[i] = pmLog(x[i]);
    expmx[i] = pmExp(-x[i]);
    pmSinCos2Pi(x[i] - std::floor(x[i]), sin2pi[i], cos2pi[i]);
  }
}

}  // extern "C"
