// synth.cpp — seeded synthetic inputs for the gather: the data a
// dump hook at gvpm.cpp:1040-1042 / gvpm_accel.h:119-199 would write, produced without Mitsuba.
//
// Scene: unit Cornell box [0,1]^3 filled with a homogeneous medium, open towards -z (index-matched
// boundary at z=0), diffuse walls, a two-sided diffuse shelf, ceiling area light.  Light paths
// follow the bookkeeping of libbidir that the photon records inherit:
// vertex 0 = emitter supernode, vertex 1 = emitter sample, vertex >= 2 scattering events;
// photon flux = prod_{k<i} vertex_k.weight * rr_k * edge_k.weight (gvpm_accel.h:134-148);
// pdf[EImportance] is in area measure (vertex.cpp:315-322); edge pdf = pdfSuccess for a medium
// successor (edge.cpp:72-76); rrWeight = 1/q, q = min(max throughput, 0.95) from rrDepth on.
// Paths are generated from a counter-based RNG keyed by the path index, so the output does not
// depend on the number of worker threads.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/gvpm_b200.h"
#include "../host/gvpm_host.hpp"

namespace {

struct Rng {  // PCG32
  uint64_t state, inc;
  Rng(uint64_t seed, uint64_t seq) {
    state = 0;
    inc = (seq << 1u) | 1u;
    next();
    state += seed;
    next();
  }
  uint32_t next() {
    uint64_t old = state;
    state = old * 6364136223846793005ULL + inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
  }
  float uniform() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }  // [0,1)
};

struct Vec { float x, y, z; };
inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator*(Vec a, float f) { return {a.x * f, a.y * f, a.z * f}; }
inline float dot(Vec a, Vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float len(Vec a) { return std::sqrt(dot(a, a)); }
inline Vec norm(Vec a) { return a * (1.0f / len(a)); }

struct Hit { float t; Vec n; Vec albedo; bool escaped; };

// shelf: y = 0.5, x in [0.3,0.7], z in [0.4,0.8]
const float SH_Y = 0.5f, SH_X0 = 0.3f, SH_X1 = 0.7f, SH_Z0 = 0.4f, SH_Z1 = 0.8f;

Hit intersectScene(Vec o, Vec d) {
  Hit h;
  h.t = 1e30f;
  h.escaped = false;
  h.n = {0, 0, 0};
  h.albedo = {0.7f, 0.7f, 0.7f};
  auto plane = [&](int axis, float pos, Vec n, Vec alb, bool esc) {
    float dc = axis == 0 ? d.x : (axis == 1 ? d.y : d.z);
    float oc = axis == 0 ? o.x : (axis == 1 ? o.y : o.z);
    if (dc == 0) return;
    float t = (pos - oc) / dc;
    if (t > 1e-6f && t < h.t) { h.t = t; h.n = n; h.albedo = alb; h.escaped = esc; }
  };
  plane(0, 0.f, {1, 0, 0}, {0.63f, 0.065f, 0.05f}, false);
  plane(0, 1.f, {-1, 0, 0}, {0.14f, 0.45f, 0.091f}, false);
  plane(1, 0.f, {0, 1, 0}, {0.7f, 0.7f, 0.7f}, false);
  plane(1, 1.f, {0, -1, 0}, {0.7f, 0.7f, 0.7f}, false);
  plane(2, 1.f, {0, 0, -1}, {0.7f, 0.7f, 0.7f}, false);
  plane(2, 0.f, {0, 0, 1}, {0, 0, 0}, true);
  if (d.y != 0) {
    float t = (SH_Y - o.y) / d.y;
    if (t > 1e-6f && t < h.t) {
      float x = o.x + d.x * t, z = o.z + d.z * t;
      if (x >= SH_X0 && x <= SH_X1 && z >= SH_Z0 && z <= SH_Z1) {
        h.t = t;
        h.n = d.y < 0 ? Vec{0, 1, 0} : Vec{0, -1, 0};
        h.albedo = {0.6f, 0.6f, 0.6f};
        h.escaped = false;
      }
    }
  }
  return h;
}

void frame(Vec n, Vec &s, Vec &t) {
  if (std::fabs(n.x) > std::fabs(n.y)) {
    float il = 1.0f / std::sqrt(n.x * n.x + n.z * n.z);
    t = {n.z * il, 0.f, -n.x * il};
  } else {
    float il = 1.0f / std::sqrt(n.y * n.y + n.z * n.z);
    t = {0.f, n.z * il, -n.y * il};
  }
  s = {t.y * n.z - t.z * n.y, t.z * n.x - t.x * n.z, t.x * n.y - t.y * n.x};
}
Vec toWorld(Vec n, Vec l) {
  Vec s, t;
  frame(n, s, t);
  return s * l.x + t * l.y + n * l.z;
}
Vec cosineHemisphere(float u1, float u2) {
  float r = std::sqrt(u1), phi = 6.28318530718f * u2;
  float x = r * std::cos(phi), y = r * std::sin(phi);
  return {x, y, std::sqrt(std::max(0.f, 1.f - x * x - y * y))};
}
Vec uniformSphere(float u1, float u2) {
  float z = 1.f - 2.f * u1, r = std::sqrt(std::max(0.f, 1.f - z * z)), phi = 6.28318530718f * u2;
  return {r * std::cos(phi), r * std::sin(phi), z};
}
float hgEval(float g, float cosWiWo) {  // phase/hg.cpp:107-110 with dot(wi,wo)
  float temp = 1.0f + g * g + 2.0f * g * cosWiWo;
  return 0.07957747154594766788f * (1 - g * g) / (temp * std::sqrt(temp));
}

struct Params {
  float sigma_s, sigma_a, hg_g;
  int phase_type, max_depth, rr_depth, min_depth;
  float power;
  bool emit_beams = false;  // emit photon beams (light-path edges) instead of volume photons
};

struct PhotonRec {  // a volume photon, or (emit_beams) a photon beam: pos = end, parent = origin
  Vec pos, flux, parent, pred, pn, prefix, albedo;
  float ppdf, epdf, rr;
  uint8_t ptype, depth;
  uint32_t path;  // local path counter, fixed up on merge
  Vec endN = {0, 0, 0};
  uint8_t endSurf = 0;
};

struct Vtx {
  Vec pos, n, albedo;
  int type;          // gvpm_parent_type
  Vec weight;        // vertex weight[EImportance]
  float rr;          // rrWeight
  float pdfArea;     // pdf[EImportance] of sampling the successor, area measure
  float edgePdf;     // pdf of the edge leaving this vertex
  Vec edgeWeight;
};

// one light path; appends photons; returns the number appended
int walk(uint64_t seed, uint64_t pathIdx, const Params &P, std::vector<PhotonRec> &out) {
  Rng rng(seed, pathIdx);
  const float sigT = P.sigma_s + P.sigma_a;
  std::vector<Vtx> v;
  v.reserve(P.max_depth + 2);
  Vtx v0{};
  v0.type = -1;
  v0.weight = {P.power, P.power, P.power};
  v0.rr = 1;
  v0.edgeWeight = {1, 1, 1};
  v0.edgePdf = 1;
  v0.pdfArea = 1.0f / (0.3f * 0.3f);
  v.push_back(v0);
  Vtx v1{};
  v1.type = GVPM_PARENT_EMITTER;
  v1.pos = {0.35f + 0.3f * rng.uniform(), 0.999f, 0.35f + 0.3f * rng.uniform()};
  v1.n = {0, -1, 0};
  v1.albedo = {0, 0, 0};
  v1.weight = {1, 1, 1};
  v1.rr = 1;
  v.push_back(v1);
  Vec dir = toWorld(v1.n, cosineHemisphere(rng.uniform(), rng.uniform()));
  float pdfOmega = std::max(0.f, dot(dir, v1.n)) * 0.31830988618f;
  Vec thr = v0.weight;  // product up to (and including) edge k-1 for the vertex being created
  int appended = 0;
  for (;;) {
    Vtx &cur = v.back();
    size_t ci = v.size() - 1;
    Hit h = intersectScene(cur.pos, dir);
    float t = -std::log(1.0f - rng.uniform()) / sigT;
    bool inMedium = t < h.t;
    float L = inMedium ? t : h.t;
    if (!inMedium && h.escaped) break;
    float T = std::exp(-sigT * L);
    cur.edgePdf = inMedium ? sigT * T : T;
    float ew = T / cur.edgePdf;
    cur.edgeWeight = {ew, ew, ew};
    Vtx nv{};
    nv.pos = cur.pos + dir * L;
    nv.rr = 1;
    if (inMedium) {
      nv.type = GVPM_PARENT_MEDIUM;
      nv.weight = {P.sigma_s, P.sigma_s, P.sigma_s};
      cur.pdfArea = pdfOmega / (L * L);
    } else {
      nv.type = GVPM_PARENT_SURFACE;
      nv.n = h.n;
      nv.albedo = h.albedo;
      nv.weight = h.albedo;
      cur.pdfArea = pdfOmega * std::fabs(dot(h.n, dir)) / (L * L);
    }
    // throughput arriving at nv (includes the edge): photon flux, gvpm_accel.h:146-148
    Vec prefix = thr;  // product up to vertex ci-1 incl. its edge = prod_{k<ci}
    Vec step = {cur.weight.x * cur.rr * cur.edgeWeight.x, cur.weight.y * cur.rr * cur.edgeWeight.y,
                cur.weight.z * cur.rr * cur.edgeWeight.z};
    Vec flux = {thr.x * step.x, thr.y * step.y, thr.z * step.z};
    size_t ni = ci + 1;  // vertexId of nv
    if (P.emit_beams) {
      // photon beam = edge ci (vertex ci -> ci+1) inside the medium, LTBeamMap::tryAppendLT
      // (gvpm_beams.h:54-84): i >= max(minDepth, 1); flux excludes the edge's own transmittance (:29-35)
      if (ci >= (size_t)std::max(1, P.min_depth)) {
        PhotonRec r;
        r.pos = nv.pos;       // beam end
        r.parent = cur.pos;   // beam origin = parent vertex of the reconnection
        r.flux = {thr.x * cur.weight.x * cur.rr, thr.y * cur.weight.y * cur.rr, thr.z * cur.weight.z * cur.rr};
        r.pred = ci >= 2 ? v[ci - 1].pos : Vec{1, 1, 1};
        r.pn = cur.n;
        r.prefix = prefix;
        r.albedo = cur.albedo;
        r.ppdf = cur.pdfArea;
        r.epdf = cur.edgePdf;
        r.rr = cur.rr;
        r.ptype = (uint8_t)cur.type;
        r.depth = (uint8_t)ci;
        r.path = 0;
        r.endN = nv.n;
        r.endSurf = inMedium ? 0 : 1;
        out.push_back(r);
        ++appended;
      }
    } else if (inMedium && ni >= (size_t)std::max(2, P.min_depth + 1)) {
      PhotonRec r;
      r.pos = nv.pos;
      r.flux = flux;
      r.parent = cur.pos;
      r.pred = ni >= 3 ? v[ci - 1].pos : Vec{1, 1, 1};
      r.pn = cur.n;
      r.prefix = prefix;
      r.albedo = cur.albedo;
      r.ppdf = cur.pdfArea;
      r.epdf = cur.edgePdf;
      r.rr = cur.rr;
      r.ptype = (uint8_t)cur.type;
      r.depth = (uint8_t)(ni - 1);
      r.path = 0;
      out.push_back(r);
      ++appended;
    }
    thr = flux;
    v.push_back(nv);
    if ((int)v.size() - 1 >= P.max_depth) break;  // path length bound
    Vtx &nw = v.back();
    // russian roulette before sampling the next direction (vertex.cpp:291-302)
    int depth = (int)v.size() - 2;
    if (depth >= P.rr_depth) {
      float m = std::max(thr.x * nw.weight.x, std::max(thr.y * nw.weight.y, thr.z * nw.weight.z));
      float q = std::min(m, 0.95f);
      if (!(rng.uniform() < q)) break;
      nw.rr = 1.0f / q;
    }
    Vec inDir = dir;
    if (nw.type == GVPM_PARENT_MEDIUM) {
      if (P.phase_type == GVPM_PHASE_HG && std::fabs(P.hg_g) > 1e-4f) {
        float g = P.hg_g, u = rng.uniform();
        float sq = (1 - g * g) / (1 - g + 2 * g * u);
        float ct = (1 + g * g - sq * sq) / (2 * g);
        float st = std::sqrt(std::max(0.f, 1 - ct * ct)), phi = 6.28318530718f * rng.uniform();
        dir = toWorld(inDir, {st * std::cos(phi), st * std::sin(phi), ct});
        pdfOmega = hgEval(g, -dot(inDir, dir));  // wi = -inDir
      } else {
        dir = uniformSphere(rng.uniform(), rng.uniform());
        pdfOmega = 0.07957747154594766788f;
      }
    } else {
      dir = toWorld(nw.n, cosineHemisphere(rng.uniform(), rng.uniform()));
      pdfOmega = std::max(0.f, dot(dir, nw.n)) * 0.31830988618f;
      if (pdfOmega <= 0) break;
    }
    dir = norm(dir);
  }
  return appended;
}

void put3(float *dst, size_t i, Vec v) { dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z; }

}  // namespace

extern "C" {

}  // extern "C" (re-opened below)

namespace {
// Walks light paths 0, 1, 2, ... (chunks of paths on worker threads, merged in path order so the result
// does not depend on the thread count) until n records are stored; emit(slot, record, pathID).
template <typename Emit>
long long runWalks(uint64_t seed, size_t n, const Params &P, int threads, Emit &&emit) {
  if (threads < 1) threads = 1;
  const uint64_t chunk = 2048;  // paths per work item
  size_t filled = 0;
  uint64_t pathBase = 0;
  uint32_t pathIdCounter = 0;
  long long totalPaths = 0;
  while (filled < n) {
    size_t remaining = n - filled;
    size_t nChunks = std::max<size_t>((size_t)threads, remaining / (chunk * 2) + 1);
    nChunks = std::min<size_t>(nChunks, 4096);
    std::vector<std::vector<PhotonRec>> res(nChunks);
    std::vector<std::vector<uint32_t>> perPath(nChunks);
    std::atomic<size_t> next(0);
    auto worker = [&]() {
      for (;;) {
        size_t c = next.fetch_add(1);
        if (c >= nChunks) break;
        auto &r = res[c];
        auto &pp = perPath[c];
        pp.resize(chunk);
        for (uint64_t k = 0; k < chunk; ++k) {
          size_t before = r.size();
          walk(seed, pathBase + c * chunk + k, P, r);
          pp[k] = (uint32_t)(r.size() - before);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
    bool done = false;
    for (size_t c = 0; c < nChunks && !done; ++c) {
      size_t off = 0;
      for (uint64_t k = 0; k < chunk && !done; ++k) {
        uint32_t cnt = perPath[c][k];
        ++totalPaths;
        uint32_t stored = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
          if (filled >= n) break;  // map full: outCapacity, gvpm_accel.h:174-176 / beams.h tryAppend
          emit(filled, res[c][off + j], pathIdCounter);
          ++filled;
          ++stored;
        }
        if (stored) ++pathIdCounter;  // gvpm_accel.h:194-197, gvpm_beams.h:77-81
        off += cnt;
        if (filled >= n) done = true;
      }
    }
    pathBase += nChunks * chunk;
  }
  return totalPaths;
}

Params makeParams(const gvpm_medium *med, int max_depth, int rr_depth, int min_depth, float power) {
  Params P;
  P.sigma_s = med->sigma_s[0];
  P.sigma_a = med->sigma_a[0];
  P.hg_g = med->hg_g;
  P.phase_type = med->phase_type;
  P.max_depth = max_depth > 0 ? max_depth : 64;
  P.rr_depth = rr_depth;
  P.min_depth = min_depth;
  P.power = power;
  return P;
}
}  // namespace

extern "C" {

// Fills the SoA arrays (caller-allocated, n entries each) with exactly n photons.
// Returns the number of light paths traced (nbPathVolume incl. empty ones), or -1.
long long gvpm_synth_photons(uint64_t seed, size_t n, const gvpm_medium *med, int max_depth, int rr_depth,
                             int min_depth, float power, int threads, gvpm_photon_soa *out) {
  Params P = makeParams(med, max_depth, rr_depth, min_depth, power);
  float *pos = (float *)out->pos, *flux = (float *)out->flux, *ppos = (float *)out->parent_pos,
        *pred = (float *)out->pred_pos, *pn = (float *)out->parent_n, *prefix = (float *)out->prefix_flux,
        *alb = (float *)out->parent_albedo, *ppdf = (float *)out->parent_pdf, *epdf = (float *)out->edge_pdf,
        *rr = (float *)out->rr_weight;
  uint8_t *ptype = (uint8_t *)out->parent_type, *depth = (uint8_t *)out->depth;
  uint32_t *pid = (uint32_t *)out->path_id;
  return runWalks(seed, n, P, threads, [&](size_t i, const PhotonRec &r, uint32_t pathId) {
    put3(pos, i, r.pos); put3(flux, i, r.flux); put3(ppos, i, r.parent);
    put3(pred, i, r.pred); put3(pn, i, r.pn); put3(prefix, i, r.prefix);
    put3(alb, i, r.albedo);
    ppdf[i] = r.ppdf; epdf[i] = r.epdf; rr[i] = r.rr;
    ptype[i] = r.ptype; depth[i] = r.depth; pid[i] = pathId;
  });
}

// Same walks, emitting photon beams (every light-path edge inside the medium, LTBeamMap::tryAppendLT,
// gvpm_beams.h:54-84).  Returns the number of light paths traced (nbPathBeams).
long long gvpm_synth_beams(uint64_t seed, size_t n, const gvpm_medium *med, int max_depth, int rr_depth,
                           int min_depth, float power, int threads, gvpm_beam_soa *out) {
  Params P = makeParams(med, max_depth, rr_depth, min_depth, power);
  P.emit_beams = true;
  float *org = (float *)out->origin, *end = (float *)out->end, *flux = (float *)out->flux,
        *prefix = (float *)out->prefix_flux, *pn = (float *)out->parent_n, *alb = (float *)out->parent_albedo,
        *pred = (float *)out->pred_pos, *endn = (float *)out->end_n, *ppdf = (float *)out->parent_pdf,
        *rr = (float *)out->rr_weight;
  uint8_t *ptype = (uint8_t *)out->parent_type, *esurf = (uint8_t *)out->end_on_surface, *depth = (uint8_t *)out->depth;
  uint32_t *pid = (uint32_t *)out->path_id;
  return runWalks(seed, n, P, threads, [&](size_t i, const PhotonRec &r, uint32_t pathId) {
    put3(org, i, r.parent); put3(end, i, r.pos); put3(flux, i, r.flux); put3(prefix, i, r.prefix);
    put3(pn, i, r.pn); put3(alb, i, r.albedo); put3(pred, i, r.pred); put3(endn, i, r.endN);
    ppdf[i] = r.ppdf; rr[i] = r.rr;
    ptype[i] = r.ptype; esurf[i] = r.endSurf; depth[i] = r.depth; pid[i] = pathId;
  });
}

// Photon planes from photon beams, as computeVolumeGradientPlanes converts them (gvpm.cpp:790-797): one sampler
// (the block-0 sampler) consumed beam by beam through LTPhotonPlane::transformBeam (host mirror in
// ../host/gvpm_host.hpp).  edge_id = the beam's edge index (= depth).  Returns n.
size_t gvpm_synth_planes(uint64_t seed, const gvpm_beam_soa *beams, size_t n, const gvpm_medium *med,
                         gvpm_plane_soa *out) {
  struct Sampler {
    Rng rng;
    float next1D() { return rng.uniform(); }
    void next2D(float &x, float &y) { x = rng.uniform(); y = rng.uniform(); }
  } sampler{Rng(seed ^ 0xA0761D6478BD642FULL, 0)};
  float *org = (float *)out->origin, *w0 = (float *)out->w0, *l0 = (float *)out->length0, *w1 = (float *)out->w1,
        *l1 = (float *)out->length1, *flux = (float *)out->flux;
  int32_t *eid = (int32_t *)out->edge_id;
  for (size_t i = 0; i < n; ++i) {
    const float *o = beams->origin + 3 * i, *e = beams->end + 3 * i;
    // PhotonBeam::setEndPoint (beams_struct.h:73-81)
    float d[3] = {e[0] - o[0], e[1] - o[1], e[2] - o[2]};
    const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float rcp = 1.0f / len;
    d[0] *= rcp; d[1] *= rcp; d[2] *= rcp;
    float nw[3], nl;
    gvpm_host::transformBeam(d, *med, sampler, nw, nl);
    for (int a = 0; a < 3; ++a) {
      org[3 * i + a] = o[a];
      w0[3 * i + a] = d[a];
      w1[3 * i + a] = nw[a];
      flux[3 * i + a] = beams->flux[3 * i + a];
    }
    l0[i] = len;
    l1[i] = nl;
    eid[i] = beams->depth[i];
  }
  return n;
}

// Occluder triangles of the synthetic scene: 5 walls + shelf = 12 triangles.  out: [12*9].
size_t gvpm_synth_occluders(float *out) {
  const float q[6][12] = {
      {0, 0, 0, 0, 0, 1, 0, 1, 1, 0, 1, 0},                                     // x = 0
      {1, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1},                                     // x = 1
      {0, 0, 0, 1, 0, 0, 1, 0, 1, 0, 0, 1},                                     // y = 0
      {0, 1, 0, 0, 1, 1, 1, 1, 1, 1, 1, 0},                                     // y = 1
      {0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1},                                     // z = 1
      {SH_X0, SH_Y, SH_Z0, SH_X1, SH_Y, SH_Z0, SH_X1, SH_Y, SH_Z1, SH_X0, SH_Y, SH_Z1}};
  size_t n = 0;
  for (int i = 0; i < 6; ++i) {
    const float *p = q[i];
    const int idx[6] = {0, 1, 2, 0, 2, 3};
    for (int k = 0; k < 6; ++k) {
      if (out) memcpy(out + 9 * n + 3 * (k % 3), p + 3 * idx[k], 12);
      if (k % 3 == 2) ++n;
    }
  }
  return n;
}

// Camera-ray medium segments for a pinhole at (0.5, 0.5, -cam_dist) looking down +z; the medium
// edge is edge 2 (sensor outside the index-matched medium).  Rays are emitted
// block by block (block x block pixels, row-major inside a block) like m_gatherBlocks
// (gvpm.cpp:271-290).  One jittered sample position per pixel, offsets at +-1 pixel
// (shift_utilities.h:255-261).  y0,y1: row range [y0,y1) of the image to emit (tile sharding).
// Returns the number of rays written (= w*(y1-y0)).
size_t gvpm_synth_rays(uint64_t seed, int w, int h, int block, int y0, int y1, float cam_dist, float cover,
                       float epsilon, gvpm_ray_soa *out) {
  float *o = (float *)out->o, *d = (float *)out->d, *mint = (float *)out->mint, *maxt = (float *)out->maxt,
        *elen = (float *)out->edge_len, *eye = (float *)out->eye_contrib, *xi = (float *)out->xi;
  int32_t *px = (int32_t *)out->px, *py = (int32_t *)out->py, *eid = (int32_t *)out->edge_id;
  uint8_t *ov = (uint8_t *)out->off_valid;
  float *oo = (float *)out->off_o, *od = (float *)out->off_d, *ol = (float *)out->off_len,
        *oe = (float *)out->off_eye, *os = (float *)out->off_sensor;
  // cam_dist < 0: the sensor sits INSIDE the medium at z = -cam_dist (photon planes need it, gvpm.cpp:785-787);
  // the medium segment is then edge 1 starting at the sensor position and `cover` is tan(fov/2)
  const bool inside = cam_dist < 0.f;
  const Vec cam = {0.5f, 0.5f, -cam_dist};
  const float tx = inside ? cover : 0.5f * cover / cam_dist, ty = tx * (float)h / (float)w;
  auto makeRay = [&](float sx, float sy, Vec &ro, Vec &rd, float &rl) -> bool {
    Vec dir = norm(Vec{(sx / w - 0.5f) * 2 * tx, (sy / h - 0.5f) * 2 * ty, 1.0f});
    if (inside) {
      ro = cam;
      rd = dir;
      Hit hit = intersectScene(ro, rd);
      rl = hit.t;
      return hit.t > 4 * epsilon && !hit.escaped;
    }
    float tIn = cam_dist / dir.z;
    ro = cam + dir * tIn;
    ro.z = 0.f;
    rd = dir;
    if (ro.x <= 0 || ro.x >= 1 || ro.y <= 0 || ro.y >= 1) { rl = 0; return false; }
    Hit hit = intersectScene(ro, rd);
    rl = hit.t;
    return hit.t > 4 * epsilon && !hit.escaped;
  };
  size_t n = 0;
  const int off[4][2] = {{-1, 0}, {1, 0}, {0, 1}, {0, -1}};
  // block < 0: |block| x |block| blocks (power of two) walked in Z-order, so 4 consecutive rays are a
  // 2x2 pixel quad, 16 a 4x4 tile, ... (ray order is the caller's choice; results are per ray)
  const bool zorder = block < 0;
  if (zorder) block = -block;
  auto compact1by1 = [](uint32_t v) {
    v &= 0x55555555u;
    v = (v ^ (v >> 1)) & 0x33333333u;
    v = (v ^ (v >> 2)) & 0x0f0f0f0fu;
    v = (v ^ (v >> 4)) & 0x00ff00ffu;
    v = (v ^ (v >> 8)) & 0x0000ffffu;
    return (int)v;
  };
  for (int by = y0; by < y1; by += block)
    for (int bx = 0; bx < w; bx += block)
      for (int i = 0; i < block * block; ++i) {
        const int x = bx + (zorder ? compact1by1((uint32_t)i) : i % block);
        const int y = by + (zorder ? compact1by1((uint32_t)i >> 1) : i / block);
        if (x >= w || y >= y1) continue;
        {
          Rng rng(seed ^ 0x9E3779B97F4A7C15ULL, (uint64_t)y * (uint64_t)w + (uint64_t)x);
          float sx = x + rng.uniform(), sy = y + rng.uniform();
          Vec ro, rd;
          float rl;
          bool ok = makeRay(sx, sy, ro, rd, rl);
          put3(o, n, ro);
          put3(d, n, rd);
          mint[n] = epsilon;
          maxt[n] = ok ? rl - epsilon : 0.f;  // empty segment when the pixel misses the medium
          elen[n] = ok ? rl : 0.f;
          put3(eye, n, Vec{1, 1, 1});
          xi[n] = rng.uniform();
          px[n] = x;
          py[n] = y;
          eid[n] = inside ? 1 : 2;
          for (int k = 0; k < 4; ++k) {
            Vec ko, kd;
            float kl;
            bool kv = makeRay(sx + off[k][0], sy + off[k][1], ko, kd, kl);
            ov[4 * n + k] = kv ? 1 : 0;
            put3(oo, 4 * n + k, ko);
            put3(od, 4 * n + k, kd);
            ol[4 * n + k] = kl;
            put3(oe, 4 * n + k, Vec{1, 1, 1});
            os[4 * n + k] = 1.0f;
          }
          ++n;
        }
      }
  return n;
}

// G-VPM camera distance samples, as computeVolumeGradientPhoton draws them (gvpm.cpp:1141-1175) for
// gather points with ONE medium edge (selBeam = {1}, pdfSel = 1): per ray `nb` samples,
// randSample = next1D() (stratified: (i + rand)/nb), then HomogeneousMedium::sampleDistance(ray(o, d,
// Epsilon, beamDist), EDistanceAlwaysValid, randSample) with the balance strategy
// (homogeneous.cpp:293-430).  Rays whose medium segment is empty get no samples.
// Arrays are caller-allocated for n_rays*nb entries; returns the number of samples written.
size_t gvpm_synth_vpm_samples(uint64_t seed, const gvpm_ray_soa *rays, size_t n_rays, int nb, int stratified,
                              const gvpm_medium *med, float epsilon, const float *radius_per_ray,
                              gvpm_vpm_sample_soa *out) {
  uint32_t *ray = (uint32_t *)out->ray;
  float *t = (float *)out->t, *T = (float *)out->transmittance, *ps = (float *)out->pdf_success,
        *psel = (float *)out->pdf_sel, *rad = (float *)out->radius;
  const float sigT[3] = {med->sigma_s[0] + med->sigma_a[0], med->sigma_s[1] + med->sigma_a[1],
                         med->sigma_s[2] + med->sigma_a[2]};
  const float samplingDensity = sigT[1];  // channel = min(int(0.5*3), 2) = 1
  const float normalization = 1.f / nb;
  size_t n = 0;
  for (size_t r = 0; r < n_rays; ++r) {
    const float beamDist = rays->edge_len[r];
    if (!(beamDist > 4 * epsilon)) continue;
    Rng rng(seed ^ 0xD1B54A32D192ED03ULL, r);
    const float mint = epsilon, maxt = beamDist;
    for (int i = 0; i < nb; ++i) {
      float rand = rng.uniform();
      if (stratified) rand = i * normalization + rand * normalization;
      // sampleDistance, EDistanceAlwaysValid (currentMediumSampling = 1)
      const float maxDist = std::max((maxt - mint) - epsilon, 0.0f);
      const float norm = 1 - std::exp(-samplingDensity * maxDist);
      const float sampledDistance = -std::log(1 - rand * norm) / samplingDensity;
      const float distSurf = maxt - mint;
      if (!(sampledDistance < distSurf)) continue;  // "Failed to sample the distance": sample skipped
      const float tt = sampledDistance + mint;
      const float px = rays->o[3 * r] + tt * rays->d[3 * r], py = rays->o[3 * r + 1] + tt * rays->d[3 * r + 1],
                  pz = rays->o[3 * r + 2] + tt * rays->d[3 * r + 2];
      if (px == rays->o[3 * r] && py == rays->o[3 * r + 1] && pz == rays->o[3 * r + 2]) continue;
      float pdfSuccess = 0;
      const float maxDist2 = maxt - mint;
      for (int c = 0; c < 3; ++c) {
        const float nrm = 1 - std::exp(-sigT[c] * maxDist2);
        const float tmp = std::exp(-sigT[c] * sampledDistance);
        pdfSuccess += (sigT[c] / nrm) * tmp;
      }
      pdfSuccess /= 3;
      float tr[3];
      for (int c = 0; c < 3; ++c) tr[c] = std::exp(sigT[c] * (-sampledDistance));
      if (std::max(tr[0], std::max(tr[1], tr[2])) < 1e-20f) tr[0] = tr[1] = tr[2] = 0.f;
      ray[n] = (uint32_t)r;
      t[n] = tt;
      T[3 * n] = tr[0]; T[3 * n + 1] = tr[1]; T[3 * n + 2] = tr[2];
      ps[n] = pdfSuccess;
      psel[n] = 1.0f;
      rad[n] = radius_per_ray[r];
      ++n;
    }
  }
  return n;
}

}  // extern "C"
