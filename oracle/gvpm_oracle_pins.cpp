// gvpm_oracle_pins.cpp — TEST INFRASTRUCTURE (see gvpm_oracle.hpp).  Batch entry points that expose the
// oracle's restated building blocks one by one, with the same flat signatures as oracle/ref_harness.cpp, so
// that tests/test_oracle_ref_pin.py can compare restatement and reference (oracle/_ref/libgvpm_ref.so, the
// reference's own code) call for call: kd layout, traversal visit sequences, intersection routines.
#include "gvpm_oracle.hpp"

using namespace gvpm_oracle;

namespace {
gvpm_medium dummyMedium() {
  gvpm_medium m;
  std::memset(&m, 0, sizeof(m));
  for (int c = 0; c < 3; ++c) { m.sigma_s[c] = 1.f; m.sigma_a[c] = 1.f; }
  m.sampling_weight = 1.f;
  return m;
}
gvpm_config dummyConfig() {
  gvpm_config c;
  std::memset(&c, 0, sizeof(c));
  c.kernel_3d = 1;
  c.epsilon = 1e-4f;
  return c;
}
gvpm_photon_soa positionsOnly(const float *pos) {
  gvpm_photon_soa s;
  std::memset(&s, 0, sizeof(s));
  s.pos = pos;
  return s;
}
CamRay<float> plainRay(const float *o, const float *d, float mint, float maxt) {
  CamRay<float> r;
  r.o = V3<float>(o);
  r.d = V3<float>(d);
  r.mint = mint;
  r.maxt = maxt;
  return r;
}
template <typename A, typename B> long long flushCSR(const std::vector<std::vector<A>> &lists, uint64_t *offsets,
                                                     B *out, size_t cap) {
  long long total = 0;
  for (size_t i = 0; i < lists.size(); ++i) {
    offsets[i] = (uint64_t)total;
    for (const A &v : lists[i]) {
      if (out && (size_t)total < cap) out[total] = (B)v;
      ++total;
    }
  }
  offsets[lists.size()] = (uint64_t)total;
  return total;
}
}  // namespace

extern "C" {

// BreTree::build (sliding midpoint) layout: same outputs as ref_kd_layout(..., sliding = 1)
int gvpm_oracle_pin_kd_layout(const float *pos, size_t n, uint32_t *orig, uint32_t *right, uint8_t *leaf,
                              uint8_t *axis) {
  BreTree<float> t;
  t.build(positionsOnly(pos), n, 0.f);
  for (size_t i = 0; i < n; ++i) {
    orig[i] = t.nodes[i].orig;
    leaf[i] = t.nodes[i].leaf ? 1 : 0;
    right[i] = t.nodes[i].leaf ? 0 : t.nodes[i].right;
    axis[i] = t.nodes[i].leaf ? 0 : t.nodes[i].axis;
  }
  return (int)t.depth;
}

long long gvpm_oracle_pin_range_visits(const float *pos, size_t n, const float *q, const float *radius, size_t m,
                                       uint64_t *offsets, uint32_t *idx, size_t cap) {
  BreTree<float> t;
  t.build(positionsOnly(pos), n, 0.f);
  std::vector<std::vector<uint32_t>> lists(m);
  for (size_t j = 0; j < m; ++j)
    t.rangeQuery(V3<float>(q + 3 * j), radius[j], [&](uint32_t node) { lists[j].push_back(t.nodes[node].orig); });
  return flushCSR(lists, offsets, idx, cap);
}

long long gvpm_oracle_pin_bre_visits(const float *pos, size_t n, float radius, const float *ray_o, const float *ray_d,
                                     const float *ray_mint, const float *ray_maxt, size_t n_rays, uint64_t *offsets,
                                     uint32_t *idx, float *tdisk, size_t cap, int *depth) {
  BreTree<float> t;
  t.build(positionsOnly(pos), n, radius);
  if (depth) *depth = (int)t.depth;
  Scene<float> sc(dummyMedium(), dummyConfig(), radius);
  std::vector<std::vector<uint32_t>> li(n_rays);
  std::vector<std::vector<float>> lt(n_rays);
  for (size_t r = 0; r < n_rays; ++r) {
    CamRay<float> ray = plainRay(ray_o + 3 * r, ray_d + 3 * r, ray_mint[r], ray_maxt[r]);
    t.query(sc, ray, [&](uint32_t node, float dd) {
      li[r].push_back(t.nodes[node].orig);
      lt[r].push_back(dd);
    });
  }
  flushCSR(lt, offsets, tdisk, cap);
  return flushCSR(li, offsets, idx, cap);
}

void gvpm_oracle_pin_cylinder(const float *co, const float *cd, const float *cmaxt, const float *vo, const float *vd,
                              const float *vmaxt, const float *radius, size_t m, uint8_t *hit, double *tNear,
                              double *tFar) {
  for (size_t i = 0; i < m; ++i) {
    double a = 0, b = 0;
    hit[i] = Scene<float>::cylinderIntersection(V3<float>(co + 3 * i), V3<float>(cd + 3 * i), cmaxt[i],
                                                V3<float>(vo + 3 * i), V3<float>(vd + 3 * i), vmaxt[i], radius[i], a, b)
                 ? 1 : 0;
    tNear[i] = a;
    tFar[i] = b;
  }
}

void gvpm_oracle_pin_plane0d(const float *ori, const float *w0, const float *len0, const float *w1, const float *len1,
                             const float *ro, const float *rd, const float *rmint, const float *rmaxt, size_t m,
                             uint8_t *hit, float *out) {
  for (size_t i = 0; i < m; ++i) {
    Scene<float>::Plane p;
    p.ori = V3<float>(ori + 3 * i);
    p.w0 = V3<float>(w0 + 3 * i);
    p.w1 = V3<float>(w1 + 3 * i);
    p.length0 = len0[i];
    p.length1 = len1[i];
    Scene<float>::PlaneIts r{0, 0, 0, 0};
    hit[i] = Scene<float>::intersectPlane0D(p, V3<float>(ro + 3 * i), V3<float>(rd + 3 * i), rmint[i], rmaxt[i], r) ? 1 : 0;
    out[4 * i] = r.tCam; out[4 * i + 1] = r.t0; out[4 * i + 2] = r.t1; out[4 * i + 3] = r.invDet;
  }
}

void gvpm_oracle_pin_beam1d(const float *origin, const float *end, const float *radius, const float *ro, const float *rd,
                            const float *rmint, const float *rmaxt, const float *tmin, const float *tmax, size_t m,
                            uint8_t *hit, float *uvws) {
  for (size_t i = 0; i < m; ++i) {
    const V3<float> o(origin + 3 * i), e(end + 3 * i);
    V3<float> dir = e - o;                 // PhotonBeam::setEndPoint, beams_struct.h:73-81
    const float len = dir.length();
    dir = dir / len;
    float u = 0, v = 0, w = 0, s = 0;
    hit[i] = Scene<float>::beamIntersect1D(o, dir, len, radius[i], V3<float>(ro + 3 * i), V3<float>(rd + 3 * i), rmint[i],
                                           rmaxt[i], tmin[i], tmax[i], u, v, w, s) ? 1 : 0;
    uvws[4 * i] = u; uvws[4 * i + 1] = v; uvws[4 * i + 2] = w; uvws[4 * i + 3] = s;
  }
}

// Occluders::anyHit on ONE triangle per query with the interval [mint, maxt]
void gvpm_oracle_pin_triangle(const float *tri, const float *ro, const float *rd, const float *mint, const float *maxt,
                              size_t m, uint8_t *hit) {
  for (size_t i = 0; i < m; ++i) {
    Occluders<float> occ;
    occ.set(tri + 9 * i, 1);
    hit[i] = occ.anyHit(V3<float>(ro + 3 * i), V3<float>(rd + 3 * i), mint[i], maxt[i]) ? 1 : 0;
  }
}

void gvpm_oracle_pin_coordsys(const float *a, size_t m, int coherent, float *b, float *c) {
  for (size_t i = 0; i < m; ++i) {
    V3<float> s, t;
    if (coherent) coordinateSystemCoherent(V3<float>(a + 3 * i), s, t);
    else Scene<float>::coordinateSystem(V3<float>(a + 3 * i), s, t);
    b[3 * i] = s.x; b[3 * i + 1] = s.y; b[3 * i + 2] = s.z;
    c[3 * i] = t.x; c[3 * i + 1] = t.y; c[3 * i + 2] = t.z;
  }
}

void gvpm_oracle_pin_quadratic(const double *abc, size_t m, uint8_t *ok, double *x0, double *x1) {
  for (size_t i = 0; i < m; ++i) {
    double a = 0, b = 0;
    ok[i] = Scene<float>::solveQuadraticDouble(abc[3 * i], abc[3 * i + 1], abc[3 * i + 2], a, b) ? 1 : 0;
    x0[i] = a;
    x1[i] = b;
  }
}

}  // extern "C"

// ---- radiometric building blocks, same flat signatures as oracle/ref_physics.cpp (the reference's medium / phase /
// BSDF / emitter plugins and shift_diffuse.cpp compiled from /root/reference) --------------------------------------------
namespace {
gvpm_medium makeMedium(const float *sigS, const float *sigA, float samplingWeight, int phaseType, float g) {
  gvpm_medium m;
  std::memset(&m, 0, sizeof(m));
  for (int c = 0; c < 3; ++c) { m.sigma_s[c] = sigS[c]; m.sigma_a[c] = sigA[c]; }
  m.sampling_weight = samplingWeight;
  m.phase_type = phaseType;
  m.hg_g = g;
  return m;
}
}  // namespace

extern "C" {

// Medium<float>::eval: HomogeneousMedium::eval, src/medium/homogeneous.cpp:432-513
void gvpm_oracle_pin_medium_eval(const float *sigS, const float *sigA, float samplingWeight, size_t n, const float *mint,
                                 const float *maxt, float *T, float *pdfSuccess, float *pdfFailure) {
  const Medium<float> med(makeMedium(sigS, sigA, samplingWeight, GVPM_PHASE_ISOTROPIC, 0.f));
  for (size_t i = 0; i < n; ++i) {
    const Medium<float>::Rec r = med.eval(mint[i], maxt[i]);
    T[3 * i] = r.transmittance.x; T[3 * i + 1] = r.transmittance.y; T[3 * i + 2] = r.transmittance.z;
    pdfSuccess[i] = r.pdfSuccess;
    pdfFailure[i] = r.pdfFailure;
  }
}

// Medium<float>::phase: phase/isotropic.cpp:76, phase/hg.cpp:107-110 (eval == pdf for both)
void gvpm_oracle_pin_phase(int type, float g, size_t n, const float *wi, const float *wo, float *eval, float *pdf) {
  const float one[3] = {1.f, 1.f, 1.f};
  const Medium<float> med(makeMedium(one, one, 1.f, type, g));
  for (size_t i = 0; i < n; ++i) eval[i] = pdf[i] = med.phase(V3<float>(wi + 3 * i), V3<float>(wo + 3 * i));
}

// Scene<float>::diffuseReconnection (shift_diffuse.cpp:11-134) on flattened parent records.  ok mirrors the reference's
// return value: false when the surface side tests fail (:43-47), when the parent pdf is zero (:100-104), or for a
// parent that is not in scope.
void gvpm_oracle_pin_diffuse_reconnection(const float *sigS, const float *sigA, float samplingWeight, int phaseType, float g,
                                          size_t n, const uint8_t *parent_type, const float *parent_pos,
                                          const float *pred_pos, const float *parent_n, const float *albedo,
                                          const float *parent_pdf, const float *edge_pdf, const float *rr_weight,
                                          const float *newD, const float *newDLength, uint8_t *ok, float *throughput,
                                          float *pdf) {
  gvpm_config cfg = dummyConfig();
  const Scene<float> sc(makeMedium(sigS, sigA, samplingWeight, phaseType, g), cfg, 1.f);
  for (size_t i = 0; i < n; ++i) {
    Photon<float> ph;
    ph.parentType = parent_type[i];
    ph.parentPos = V3<float>(parent_pos + 3 * i);
    ph.predPos = V3<float>(pred_pos + 3 * i);
    ph.parentN = V3<float>(parent_n + 3 * i);
    ph.albedo = V3<float>(albedo + 3 * i);
    ph.parentPdf = parent_pdf[i];
    ph.edgePdf = edge_pdf[i];
    ph.rrWeight = rr_weight[i];
    V3<float> thr;
    float p = 0.f;
    const bool good = sc.diffuseReconnectionStatus(ph, V3<float>(newD + 3 * i), newDLength[i], thr, p);
    ok[i] = good ? 1 : 0;
    throughput[3 * i] = thr.x; throughput[3 * i + 1] = thr.y; throughput[3 * i + 2] = thr.z;
    pdf[i] = p;
  }
}

}  // extern "C"
