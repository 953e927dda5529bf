// gvpm_oracle.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's volumetric density-estimation gather (gradientpm/gvpm,
// src/integrators/photonmapper/...), used ONLY by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py as the checker and the timed CPU baseline.
// The product path (gvpm_b200/csrc) never includes, links or calls anything in oracle/.
//
// PARITY PINNED TO THE REFERENCE'S OWN COMPILED CODE.  The reference ships no test, golden vector or fixture for this
// path and Mitsuba as a whole cannot be built in this image (Boost, Eigen, Xerces, OpenEXR ... are
// absent), but the files of the path compile where they lie under /root/reference (oracle/Makefile,
// outputs in oracle/_ref/), and this restatement is compared with them BIT-EXACTLY, live and through committed golden
// vectors generated from them:
//   * everything that decides a neighbour index set - PointKDTree build + range query, the AABB slab test, GPhotonMap +
//     GradientBeamRadianceEstimator (hierarchy + traversal + neighbour predicate), SubBeamBVH, PhotonPlaneBVH,
//     cylinderIntersection, intersectPlane0D, rayIntersectInternal1D, Triangle::rayIntersect,
//     coordinateSystem(Coherent), solveQuadraticDouble (ref_harness.cpp -> _ref/libgvpm_ref.so;
//     tests/test_oracle_ref_pin.py, tests/golden/ref_pins.npz);
//   * the radiometric building blocks - HomogeneousMedium::eval, the phase functions, the diffuse BSDF, the area
//     emitter, diffuseReconnection (ref_physics.cpp -> _ref/libgvpm_physics_ref.so; tests/test_oracle_physics_pin.py,
//     tests/golden/physics_pins.npz);
//   * the shift functors as a whole - VolumeGradientBREQuery, VolumeGradientPositionQuery, BeamGradRadianceQuery (3-D
//     and 1-D kernels), PlaneGradRadianceQuery and sppm's BeamRadianceQuery, i.e. contributions, Jacobians, MIS
//     weights, filters, border rule and accumulation (ref_functor.cpp -> _ref/libgvpm_functor_ref.so;
//     tests/test_oracle_functor_pin.py, tests/golden/functor_pins.npz).
//   * sppm's primal functors - BeamRadianceQuery (four techniques) and the loop body of BeamRadianceEstimator::query
//     (same library and tests).
// Camera segments beyond the first medium edge agree to 5e-7 instead of bit for bit (sensorMIS's cancelling geometry
// terms), sppm's BRE under Henyey-Greenstein to 1.2e-6 (the flattened photon direction); DESIGN.md §5.
//
// Template parameter Real = float restates the SINGLE_PRECISION build
// (build/config-linux-gcc.py:7); Real = double is the error-budget variant.  Compile with
// -ffp-contract=off so no FMA is formed: the neighbour predicate must be bit-reproducible.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/gvpm_b200.h"

namespace gvpm_oracle {

template <typename Real> struct V3 {
  Real x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(Real a, Real b, Real c) : x(a), y(b), z(c) {}
  explicit V3(const float *p) : x((Real)p[0]), y((Real)p[1]), z((Real)p[2]) {}
  Real operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  Real &at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
  V3 operator+(const V3 &o) const { return V3(x + o.x, y + o.y, z + o.z); }
  V3 operator-(const V3 &o) const { return V3(x - o.x, y - o.y, z - o.z); }
  V3 operator-() const { return V3(-x, -y, -z); }
  V3 operator*(Real f) const { return V3(x * f, y * f, z * f); }
  // include/mitsuba/core/vector.h:535-542: division multiplies by the reciprocal
  V3 operator/(Real f) const { Real r = (Real)1 / f; return V3(x * r, y * r, z * r); }
  V3 &operator+=(const V3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
  // component-wise (Spectrum) product, spectrum.h:385-397
  V3 operator*(const V3 &o) const { return V3(x * o.x, y * o.y, z * o.z); }
  Real lengthSquared() const { return x * x + y * y + z * z; }
  Real length() const { return std::sqrt(lengthSquared()); }
  Real maxc() const { return std::max(x, std::max(y, z)); }
};
template <typename Real> inline V3<Real> operator*(Real f, const V3<Real> &v) { return v * f; }
template <typename Real> inline Real dot(const V3<Real> &a, const V3<Real> &b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;  // vector.h:609-611
}
template <typename Real> inline V3<Real> cross(const V3<Real> &a, const V3<Real> &b) {
  return V3<Real>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename Real> inline V3<Real> normalize(const V3<Real> &v) { return v / v.length(); }
template <typename Real> inline Real safe_sqrt(Real v) { return std::sqrt(std::max((Real)0, v)); }

// coordinateSystemCoherent, src/libcore/util.cpp:592-599 (Duff et al. 2017); sign, a, b are
// `float` in the reference whatever Float is.
template <typename Real> inline void coordinateSystemCoherent(const V3<Real> &n, V3<Real> &b1, V3<Real> &b2) {
  float sign = copysignf(1.0f, (float)n.z);
  const float a = (float)(-1.0f / (sign + n.z));
  const float b = (float)(n.x * n.y * a);
  b1 = V3<Real>((Real)(1.0f + sign * n.x * n.x * a), (Real)(sign * b), (Real)(-sign * n.x));
  b2 = V3<Real>((Real)b, (Real)(sign + n.y * n.y * a), (Real)(-n.y));
}

template <typename Real> struct Consts;
template <> struct Consts<float> {
  static constexpr float pi = 3.14159265358979323846f, inv_pi = 0.31830988618379067154f,
                         inv_fourpi = 0.07957747154594766788f;
};
template <> struct Consts<double> {
  static constexpr double pi = 3.14159265358979323846, inv_pi = 0.31830988618379067154,
                          inv_fourpi = 0.07957747154594766788;
};

// ---------------------------------------------------------------------------------------
// Homogeneous medium (balance strategy), src/medium/homogeneous.cpp:432-513
template <typename Real> struct Medium {
  V3<Real> sigmaS, sigmaA, sigmaT;
  int phaseType;
  Real g, samplingWeight;
  explicit Medium(const gvpm_medium &m)
      : sigmaS(m.sigma_s), sigmaA(m.sigma_a), phaseType(m.phase_type), g((Real)m.hg_g),
        samplingWeight((Real)m.sampling_weight) {
    sigmaT = sigmaS + sigmaA;
  }
  struct Rec { V3<Real> transmittance; Real pdfSuccess, pdfFailure; };
  // math::fastexp, include/mitsuba/core/math.h:175-187: on Linux / x86_64 the single-precision build evaluates exp in
  // DOUBLE and rounds the result to float (pinned bit-exactly against the reference's compiled HomogeneousMedium::eval,
  // tests/test_oracle_physics_pin.py)
  static Real fastexp(Real v) { return (Real)std::exp((double)v); }
  // eval(ray, mRec) with EDistanceNormal: distance = maxt - mint
  Rec eval(Real mint, Real maxt) const {
    Rec r;
    Real distance = maxt - mint;
    r.pdfSuccess = 0;
    r.pdfFailure = 0;
    for (int i = 0; i < 3; ++i) {  // :478-483
      Real tmp = fastexp(-sigmaT[i] * distance);
      r.pdfFailure += tmp;
      r.pdfSuccess += sigmaT[i] * tmp;
    }
    r.pdfSuccess /= 3;
    r.pdfFailure /= 3;
    r.transmittance = V3<Real>(fastexp(sigmaT.x * (-distance)), fastexp(sigmaT.y * (-distance)),
                               fastexp(sigmaT.z * (-distance)));  // :504
    r.pdfSuccess = r.pdfSuccess * samplingWeight;                    // :505
    r.pdfFailure = r.pdfFailure * samplingWeight + (1 - samplingWeight);
    if (r.transmittance.maxc() < (Real)1e-20) r.transmittance = V3<Real>();  // :511-512
    return r;
  }
  // PhaseFunction::eval(pRec(wi, wo)): phase/isotropic.cpp:76, phase/hg.cpp:107-110
  Real phase(const V3<Real> &wi, const V3<Real> &wo) const {
    if (phaseType == GVPM_PHASE_ISOTROPIC) return Consts<Real>::inv_fourpi;
    Real temp = (Real)1 + g * g + (Real)2 * g * dot(wi, wo);
    return Consts<Real>::inv_fourpi * (1 - g * g) / (temp * std::sqrt(temp));
  }
};

// ---------------------------------------------------------------------------------------
// Input views (float storage, converted on access)
template <typename Real> struct Photon {
  V3<Real> pos, flux, parentPos, predPos, parentN, prefix, albedo;
  Real parentPdf, edgePdf, rrWeight;
  int parentType, depth;
  uint32_t pathId;
};
template <typename Real> inline Photon<Real> loadPhoton(const gvpm_photon_soa &s, size_t i) {
  Photon<Real> p;
  p.pos = V3<Real>(s.pos + 3 * i);
  p.flux = V3<Real>(s.flux + 3 * i);
  p.parentPos = V3<Real>(s.parent_pos + 3 * i);
  p.predPos = V3<Real>(s.pred_pos + 3 * i);
  p.parentN = V3<Real>(s.parent_n + 3 * i);
  p.prefix = V3<Real>(s.prefix_flux + 3 * i);
  p.albedo = V3<Real>(s.parent_albedo + 3 * i);
  p.parentPdf = (Real)s.parent_pdf[i];
  p.edgePdf = (Real)s.edge_pdf[i];
  p.rrWeight = (Real)s.rr_weight[i];
  p.parentType = s.parent_type[i];
  p.depth = s.depth[i];
  p.pathId = s.path_id[i];
  return p;
}

template <typename Real> struct CamRay {
  V3<Real> o, d, eye;
  Real mint, maxt, edgeLen, xi;
  int px, py, edgeId;
  bool offValid[4];
  V3<Real> offO[4], offD[4], offEye[4];
  Real offLen[4], offSensor[4];
};
template <typename Real> inline CamRay<Real> loadRay(const gvpm_ray_soa &s, size_t i) {
  CamRay<Real> r;
  r.o = V3<Real>(s.o + 3 * i);
  r.d = V3<Real>(s.d + 3 * i);
  r.eye = V3<Real>(s.eye_contrib + 3 * i);
  r.mint = (Real)s.mint[i];
  r.maxt = (Real)s.maxt[i];
  r.edgeLen = (Real)s.edge_len[i];
  r.xi = (Real)s.xi[i];
  r.px = s.px[i];
  r.py = s.py[i];
  r.edgeId = s.edge_id[i];
  for (int k = 0; k < 4; ++k) {
    r.offValid[k] = s.off_valid[4 * i + k] != 0;
    r.offO[k] = V3<Real>(s.off_o + 3 * (4 * i + k));
    r.offD[k] = V3<Real>(s.off_d + 3 * (4 * i + k));
    r.offEye[k] = V3<Real>(s.off_eye + 3 * (4 * i + k));
    r.offLen[k] = (Real)s.off_len[4 * i + k];
    r.offSensor[k] = (Real)s.off_sensor[4 * i + k];
  }
  return r;
}

// ---------------------------------------------------------------------------------------
// Occluders: Triangle::rayIntersect, include/mitsuba/core/triangle.h:109-145, any-hit over a
// flat list with the (mint, maxt) interval of Scene::rayIntersect(ray).
template <typename Real> struct Occluders {
  std::vector<V3<Real>> v;  // 3 per triangle
  void set(const float *tri, size_t n) {
    v.resize(3 * n);
    for (size_t i = 0; i < 3 * n; ++i) v[i] = V3<Real>(tri + 3 * i);
  }
  bool anyHit(const V3<Real> &o, const V3<Real> &d, Real mint, Real maxt) const {
    for (size_t t = 0; t + 2 < v.size(); t += 3) {
      V3<Real> edge1 = v[t + 1] - v[t], edge2 = v[t + 2] - v[t];
      V3<Real> pvec = cross(d, edge2);
      Real det = dot(edge1, pvec);
      if (det == 0) continue;
      Real inv_det = (Real)1 / det;
      V3<Real> tvec = o - v[t];
      Real u = dot(tvec, pvec) * inv_det;
      if (u < 0 || u > 1) continue;
      V3<Real> qvec = cross(tvec, edge1);
      Real vv = dot(d, qvec) * inv_det;
      if (vv >= 0 && u + vv <= 1) {
        Real tt = dot(edge2, qvec) * inv_det;
        if (tt >= mint && tt <= maxt) return true;  // Ray interval, shape.h semantics
      }
    }
    return false;
  }
};

// ---------------------------------------------------------------------------------------
// Per-ray accumulators: AbstractVolumeGradientRecord, shift/shift_volume_photon.h:17-20
template <typename Real> struct Accum {
  V3<Real> mediumFlux, shifted[4], weighted[4];
  void store(float *out) const {
    auto put = [&](int j, const V3<Real> &v) {
      out[3 * j] = (float)v.x; out[3 * j + 1] = (float)v.y; out[3 * j + 2] = (float)v.z;
    };
    put(0, mediumFlux);
    for (int k = 0; k < 4; ++k) put(1 + k, shifted[k]);
    for (int k = 0; k < 4; ++k) put(5 + k, weighted[k]);
  }
};

template <typename Real> struct GradientSamplingResult {  // shift_utilities.h:17-23
  V3<Real> shiftedFlux;
  Real weight = 1, jacobian = 1;
};

// 1/max(2*deltaT, 0.0001) evaluated in double like the reference (shift_volume_photon.cpp:723)
template <typename Real> inline Real chordPdf(Real deltaT) {
  return (Real)(1.f / std::max((double)deltaT * 2.0, 0.0001));
}

template <typename Real> struct Scene {
  Medium<Real> medium;
  gvpm_config cfg;
  Occluders<Real> occ;
  Real radius;
  explicit Scene(const gvpm_medium &m, const gvpm_config &c, Real r) : medium(m), cfg(c), radius(r) {}

  // computeVolumeContribution, shift_utilities.h:233-253 (bsdfInteractionMode = EAll)
  bool lightingModeAccepts(int parentType) const {
    int m = cfg.lighting_mode;
    if ((m & GVPM_SURF2MEDIA) && (m & GVPM_MEDIA2MEDIA)) return true;
    if (parentType == GVPM_PARENT_MEDIUM && !(m & GVPM_MEDIA2MEDIA)) return false;
    if (parentType != GVPM_PARENT_MEDIUM && !(m & GVPM_SURF2MEDIA)) return false;
    return true;
  }

  // getVolumePhotonContrib, shift_volume_photon.h:79-86
  V3<Real> volumePhotonContrib(const V3<Real> &flux, const V3<Real> &wi, const V3<Real> &wo) const {
    return (medium.sigmaS * flux) * medium.phase(wi, wo);
  }

  // diffuseReconnection, shift/operation/shift_diffuse.cpp:11-134, for the in-scope vertex
  // types {area emitter (emitters/area.cpp:132-150), diffuse BSDF (bsdfs/diffuse.cpp:110-127),
  // medium}.  Returns throughput and pdf (pdf = 0 when the reconnection is impossible).
  void diffuseReconnection(const Photon<Real> &ph, const V3<Real> &newD, Real newDLength,
                           V3<Real> &throughput, Real &pdf) const {
    diffuseReconnectionStatus(ph, newD, newDLength, throughput, pdf);
  }
  // same, returning the reference function's own return value (false: the caller treats the shift as failed); throughput
  // and pdf hold what the reference's ShiftRecord holds at that point, which is what the pin against
  // shift_diffuse.cpp compares (tests/test_oracle_physics_pin.py)
  bool diffuseReconnectionStatus(const Photon<Real> &ph, const V3<Real> &newD, Real newDLength,
                                 V3<Real> &throughput, Real &pdf) const {
    throughput = V3<Real>(1, 1, 1);
    pdf = 0;
    Real pdfValue;
    const Real INV_PI = Consts<Real>::inv_pi;
    if (ph.parentType == GVPM_PARENT_SURFACE) {
      V3<Real> wiWorld = normalize(ph.predPos - ph.parentPos);
      Real cosI = dot(ph.parentN, wiWorld), cosO = dot(ph.parentN, newD);
      if (cosI <= 0 || cosO <= 0) {
        throughput = V3<Real>();
        pdfValue = 0;
      } else {
        throughput = throughput * (ph.albedo * (INV_PI * cosO));
        pdfValue = INV_PI * cosO;
      }
      // geometric == shading normal for the flattened record: :43-47
      if (cosI * cosI <= 0 || cosO * cosO <= 0) return false;
    } else if (ph.parentType == GVPM_PARENT_MEDIUM) {
      V3<Real> pWi = normalize(ph.predPos - ph.parentPos);
      Real phv = medium.phase(pWi, newD);
      throughput = throughput * (medium.sigmaS * phv);
      pdfValue = phv;
    } else if (ph.parentType == GVPM_PARENT_EMITTER) {
      Real dp = dot(newD, ph.parentN);
      if (dp < 0) dp = 0;
      throughput = throughput * V3<Real>(INV_PI * dp, INV_PI * dp, INV_PI * dp);
      pdfValue = INV_PI * dp;
    } else {
      return false;  // glossy parent: manifold shift, out of scope (treated as a failed shift)
    }
    Real GOp = 1 / (newDLength * newDLength);  // :89-92 (isVolumeBase)
    pdf = pdfValue * GOp;
    throughput = throughput * GOp;
    if (ph.parentPdf == 0) { pdf = 0; return false; }  // :100-104
    throughput = throughput / ph.parentPdf;       // :111
    throughput = throughput * ph.rrWeight;        // :112
    // edge->medium != nullptr always holds for a volume photon's last edge: :114-131
    typename Medium<Real>::Rec m = medium.eval(0, newDLength);
    pdf *= m.pdfSuccess;
    throughput = throughput * (m.transmittance / ph.edgePdf);
    return true;
  }

  // AbstractVolumeGradientRecord::shiftNull, shift_volume_photon.cpp:119-158
  void shiftNull(const Photon<Real> &ph, const V3<Real> &wi, const CamRay<Real> &ray, int k,
                 const typename Medium<Real>::Rec &mShift, GradientSamplingResult<Real> &res,
                 Real pdfBaseRay, Real pdfShiftRay) const {
    V3<Real> contrib = volumePhotonContrib(ph.flux, wi, -ray.offD[k]);
    res.shiftedFlux = ((mShift.transmittance * contrib) * ray.offEye[k]) * res.jacobian;
    res.weight = (Real)0.5;
    if (cfg.use_mis) {
      if (pdfShiftRay == 0 || pdfBaseRay == 0) { res.weight = 1; return; }
      res.weight = (Real)1 / ((Real)1 + ray.offSensor[k] * pdfShiftRay * res.jacobian / pdfBaseRay);
    }
  }

  // shiftPhoton -> shiftPhotonDiffuse, shift_volume_photon.cpp:49-117,382-486
  void shiftPhotonDiffuse(const Photon<Real> &ph, const V3<Real> &offsetPos, const CamRay<Real> &ray,
                          int k, const typename Medium<Real>::Rec &mShift,
                          GradientSamplingResult<Real> &res, Real pdfBaseRay, Real pdfShiftRay) const {
    if (ph.parentType == GVPM_PARENT_OTHER) return;  // EManifoldShift with useManifold=false
    V3<Real> dProj = offsetPos - ph.parentPos;
    Real lProj = dProj.length();
    dProj = dProj / lProj;
    // Ray projRay(parent, dProj, Epsilon, lProj * ShadowEpsilon): :396
    if (occ.anyHit(ph.parentPos, dProj, (Real)cfg.epsilon, lProj * (Real)cfg.shadow_maxt_scale)) return;
    if (ph.parentType != GVPM_PARENT_MEDIUM) {  // :404-412
      V3<Real> edgeD = normalize(ph.pos - ph.parentPos);
      Real signDot = dot(ph.parentN, dProj) / dot(ph.parentN, edgeD);
      if (signDot < 0) return;
    }
    V3<Real> thr;
    Real sPdf;
    diffuseReconnection(ph, dProj, lProj, thr, sPdf);
    if (sPdf == 0) { res.weight = 1; return; }
    V3<Real> photonWeight = ph.prefix * thr;
    V3<Real> contrib = volumePhotonContrib(photonWeight, -dProj, -ray.offD[k]);
    res.shiftedFlux = ((mShift.transmittance * contrib) * ray.offEye[k]) * res.jacobian;
    res.weight = (Real)0.5;
    if (cfg.use_mis) {
      Real basePdf = pdfBaseRay;
      basePdf *= ph.parentPdf;
      basePdf *= ph.edgePdf;
      Real offsetPdf = sPdf * pdfShiftRay;
      if (offsetPdf == 0 || basePdf == 0) { res.weight = 1; return; }
      Real q = ray.offSensor[k] * res.jacobian * (offsetPdf / basePdf);
      res.weight = cfg.power_heuristic ? (Real)1 / ((Real)1 + q * q) : (Real)1 / ((Real)1 + q);
    }
  }

  // AbstractVolumeGradientRecord::getShiftPos, shift_volume_photon.cpp:858-896 (3-D kernel)
  V3<Real> getShiftPos(const CamRay<Real> &ray, int k, Real tBase, const V3<Real> &basePhotonPos) const {
    return getShiftPos(ray, k, tBase, basePhotonPos, radius, !cfg.kernel_3d);
  }
  V3<Real> getShiftPos(const CamRay<Real> &ray, int k, Real tBase, const V3<Real> &basePhotonPos, Real radius,
                       bool coherent) const {
    V3<Real> zBase = ray.o + tBase * ray.d, zShift = ray.offO[k] + tBase * ray.offD[k];
    V3<Real> offsetPos = zShift + (basePhotonPos - zBase);
    if (coherent) {  // coherent frames for the 2-D kernel, :866-873
      V3<Real> bs, bt, ns, nt;
      coordinateSystemCoherent(ray.d, bs, bt);
      coordinateSystemCoherent(ray.offD[k], ns, nt);
      V3<Real> v = basePhotonPos - zBase;
      V3<Real> localD(dot(v, bs), dot(v, bt), dot(v, ray.d));          // Frame::toLocal
      offsetPos = zShift + ((ns * localD.x + nt * localD.y) + ray.offD[k] * localD.z);  // toWorld
    }
    if (cfg.use_shift_null) {
      Real offDistSqr = (zBase - offsetPos).lengthSquared();
      if (offDistSqr < radius * radius) {
        V3<Real> dShift = zShift - zBase;
        dShift = dShift / dShift.length();
        Real cosD = dot(dShift, -(offsetPos - zShift));
        offsetPos += (dShift * cosD) * (Real)2;
      }
    }
    return offsetPos;
  }

  // The neighbour predicate of GradientBeamRadianceEstimator::query, gvpm_accel.h:293-301.
  // Returns true and diskDistance when the photon reaches the functor.
  bool diskTest(const CamRay<Real> &ray, const V3<Real> &p, Real &diskDistance) const {
    V3<Real> originToCenter = p - ray.o;
    diskDistance = dot(originToCenter, ray.d);
    Real radSqr = radius * radius;
    Real distSqr = ((ray.o + diskDistance * ray.d) - p).lengthSquared();
    return diskDistance > ray.mint && distSqr < radSqr;
  }

  // ---- sppm primal BRE: BeamRadianceEstimator::query, photonmapper/bre.cpp:167-259, driven as sppm.cpp:960-981 -----
  // The per-photon sampler->next1D() of the 3-D kernel (:217) is replaced by a counter-based hash of the ray and the
  // caller's photon index (traversal-order independent); m_scaleFactor is left to the caller.
  static uint32_t sppmHash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
  }
  Real sppmUniform(const CamRay<Real> &ray, uint32_t photonIndex) const {
    uint32_t h = sppmHash32(cfg.rng_seed ^ 0x9E3779B9u);
    h = sppmHash32(h ^ (uint32_t)ray.px);
    h = sppmHash32(h ^ ((uint32_t)ray.py * 0x85EBCA6Bu));
    h = sppmHash32(h ^ ((uint32_t)ray.edgeId * 0xC2B2AE35u));
    h = sppmHash32(h ^ photonIndex);
    return (Real)((float)(h >> 8) * (1.0f / 16777216.0f));
  }
  // One photon of the query loop (:195-254) on the ray re-based at r(r.mint) (:169).  0 = rejected by the geometric
  // tests, 1 = passes them but is dropped by the depth filter (:195-198), 2 = contributes (result += ...).
  int sppmBreFunctor(const CamRay<Real> &ray, const Photon<Real> &ph, uint32_t photonIndex, V3<Real> &result) const {
    const V3<Real> ro = ray.o + ray.mint * ray.d;
    const Real rmaxt = ray.maxt - ray.mint;
    const Real r = radius;
    const V3<Real> originToCenter = ph.pos - ro;
    const Real diskDistance = dot(originToCenter, ray.d), radSqr = r * r;
    const Real distSqr = ((ro + diskDistance * ray.d) - ph.pos).lengthSquared();
    if (!(diskDistance > 0 && distSqr < radSqr)) return 0;
    Real tEval, scale, invPdf = 1;
    if (cfg.kernel_3d) {
      if (diskDistance - (r * 2) > rmaxt) return 0;
      const Real weight = (Real)(1 / ((4.0 / 3.0) * (double)Consts<Real>::pi * std::pow((double)r, 3)));
      const Real deltaT = std::sqrt(radSqr - distSqr);
      const Real tminKernel = diskDistance - deltaT;
      const Real diskDistanceRand = tminKernel + 2 * deltaT * sppmUniform(ray, photonIndex);
      if (diskDistanceRand < 0 || diskDistanceRand > rmaxt) return 0;
      const Real invPdfSampling = std::max((Real)(2.0f * deltaT), (Real)0.0001f);
      tEval = diskDistanceRand;
      scale = weight;              // (weight * m_scaleFactor) * invPdfSampling, in the reference's order (:233-235)
      invPdf = invPdfSampling;
    } else {
      if (diskDistance > rmaxt) return 0;
      tEval = diskDistance;
      scale = (Real)(1 / ((double)Consts<Real>::pi * std::pow((double)r, 2)));
    }
    if (cfg.max_depth != -1 && ph.depth > cfg.max_depth - ray.edgeId) return 1;
    const V3<Real> wi = normalize(ph.parentPos - ph.pos);   // = -photon.getDirection()
    typename Medium<Real>::Rec mRecBase = medium.eval(0, tEval);
    V3<Real> term = ((mRecBase.transmittance * ph.flux) * medium.phase(wi, -ray.d)) * scale;
    if (cfg.kernel_3d) term = term * invPdf;
    result += term * ray.eye;
    return 2;
  }

  // VolumeGradientBREQuery::operator(), shift_volume_photon.cpp:658-856.
  // Returns 0 = not in the geometric set, 1 = geometric only (filtered), 2 = contributes.
  int breFunctor(const CamRay<Real> &ray, const Photon<Real> &ph, Real diskDistance, Accum<Real> &acc) const {
    const Real r = radius;
    bool filtered = false;
    int pathLen = ph.depth + ray.edgeId;
    if (cfg.max_depth > 0 && pathLen > cfg.max_depth) filtered = true;       // :670
    if (cfg.min_depth != 0 && pathLen < cfg.min_depth) filtered = true;      // :672
    if (!lightingModeAccepts(ph.parentType)) filtered = true;                // :675-677
    Real rrGlobalWeight = 1;
    if (cfg.path_set) {                                                      // :689-697
      uint32_t currentGroup = (uint32_t)((ray.px + ray.py) % 2);
      if (ph.pathId % 2 != currentGroup) filtered = true;
      rrGlobalWeight = 2;
    }
    Real kernelVol, pdfCameraPos = 1, tBase = diskDistance;
    if (cfg.kernel_3d) {                                                     // :707-724
      kernelVol = (Real)((4.0 / 3.0) * (double)Consts<Real>::pi * std::pow((double)r, 3));
      Real distSqr = ((ray.o + diskDistance * ray.d) - ph.pos).lengthSquared();
      Real deltaT = safe_sqrt(r * r - distSqr);
      Real tminKernel = diskDistance - deltaT;
      Real diskDistanceRand = tminKernel + (deltaT * 2) * ray.xi;
      if (diskDistanceRand < ray.mint || diskDistanceRand > ray.edgeLen) return 0;
      tBase = diskDistanceRand;
      pdfCameraPos = chordPdf(deltaT);
    } else {
      kernelVol = (Real)((double)Consts<Real>::pi * std::pow((double)r, 2));
      // the reference leaves `baseProjDist > edgeLen` as an empty block (:726-731), which makes
      // its 2-D result depend on the tree shape past the ray end; restated with the explicit
      // bound sppm uses (bre.cpp:240-242).  Documented deviation, DESIGN.md §6.
      if (diskDistance > ray.edgeLen) return 0;
    }
    if (filtered) return 1;

    const V3<Real> wi = normalize(ph.parentPos - ph.pos);  // -edge(c-1).d
    typename Medium<Real>::Rec mBase = medium.eval(ray.mint, tBase);
    V3<Real> contrib = volumePhotonContrib(ph.flux, wi, -ray.d);
    V3<Real> baseContrib = (mBase.transmittance * contrib) * ray.eye;
    const Real norm = kernelVol * pdfCameraPos;
    acc.mediumFlux += (baseContrib / norm) * rrGlobalWeight;                 // :751

    for (int k = 0; k < 4; ++k) {
      GradientSamplingResult<Real> res;
      if (ray.offValid[k]) {
        const Real shiftDistTotal = ray.offLen[k];
        bool alreadyShift = false;
        if (cfg.use_shift_null && cfg.kernel_3d) {                           // :776-802
          V3<Real> zp = ray.offO[k] + tBase * ray.offD[k];
          Real ZPtoY = (zp - ph.pos).lengthSquared();
          if (ZPtoY < r * r && tBase < shiftDistTotal) {
            Real dd = dot(ph.pos - ray.offO[k], ray.offD[k]);
            Real ds = ((ray.offO[k] + dd * ray.offD[k]) - ph.pos).lengthSquared();
            Real pdfShiftPos = chordPdf(safe_sqrt(r * r - ds));
            typename Medium<Real>::Rec mShift = medium.eval((Real)cfg.epsilon, tBase);
            shiftNull(ph, wi, ray, k, mShift, res, pdfCameraPos, pdfShiftPos);
            alreadyShift = true;
          }
        }
        if (!alreadyShift && shiftDistTotal >= tBase) {                      // :809-838
          V3<Real> offsetPos = getShiftPos(ray, k, tBase, ph.pos);
          Real pdfShiftPos = 1;
          if (cfg.kernel_3d) {
            Real dd = dot(offsetPos - ray.offO[k], ray.offD[k]);
            Real ds = ((ray.offO[k] + dd * ray.offD[k]) - offsetPos).lengthSquared();
            pdfShiftPos = chordPdf(safe_sqrt(r * r - ds));
          }
          typename Medium<Real>::Rec mShift = medium.eval((Real)cfg.epsilon, tBase);
          shiftPhotonDiffuse(ph, offsetPos, ray, k, mShift, res, pdfCameraPos, pdfShiftPos);
        }
      }
      if ((k == 1 && ray.px == cfg.film_w - 1) || (k == 2 && ray.py == cfg.film_h - 1))
        res.weight = 1;                                                      // :843-846
      acc.weighted[k] += ((rrGlobalWeight * res.weight) * baseContrib) / norm;
      acc.shifted[k] += ((rrGlobalWeight * res.weight) * res.shiftedFlux) / norm;
    }
    return 2;
  }

  // ---- G-VPM -----------------------------------------------------------------------------
  // One camera distance sample (gvpm.cpp:1141-1175), with the four cached shiftMRec
  // (VolumeGradientPositionQuery, shift_volume_photon.cpp:543-567): they depend on the sample only.
  struct VpmSample {
    uint32_t ray;
    Real t, pdfSuccess, pdfSel, radius;
    V3<Real> T;
    bool validShiftDist[4];
    typename Medium<Real>::Rec shiftMRec[4];
  };
  VpmSample loadVpmSample(const gvpm_vpm_sample_soa &s, size_t i, const CamRay<Real> &ray) const {
    VpmSample v;
    v.ray = s.ray[i];
    v.t = (Real)s.t[i];
    v.pdfSuccess = (Real)s.pdf_success[i];
    v.pdfSel = (Real)s.pdf_sel[i];
    v.radius = (Real)s.radius[i];
    v.T = V3<Real>(s.transmittance + 3 * i);
    for (int k = 0; k < 4; ++k) {
      v.validShiftDist[k] = ray.offValid[k] && ray.offLen[k] >= v.t;  // :549-553
      v.shiftMRec[k].pdfSuccess = 0;
      v.shiftMRec[k].pdfFailure = 0;
      if (v.validShiftDist[k]) {
        // medium->eval(shiftRay(o_k, d_k, Epsilon, len_k), shiftMRec, EDistanceAlwaysValid) with mRec.t = t:
        // homogeneous.cpp:468-476,504-513 (currentMediumSampling forced to 1)
        const Real maxDist = ray.offLen[k] - (Real)cfg.epsilon, distance = v.t;
        Real ps = 0;
        for (int c = 0; c < 3; ++c) {
          const Real normalization = 1 - Medium<Real>::fastexp(-medium.sigmaT[c] * maxDist);
          const Real tmp = Medium<Real>::fastexp(-medium.sigmaT[c] * distance);
          ps += (medium.sigmaT[c] / normalization) * tmp;
        }
        ps /= 3;
        v.shiftMRec[k].pdfSuccess = ps;
        v.shiftMRec[k].transmittance = V3<Real>(Medium<Real>::fastexp(medium.sigmaT.x * (-distance)),
                                                Medium<Real>::fastexp(medium.sigmaT.y * (-distance)),
                                                Medium<Real>::fastexp(medium.sigmaT.z * (-distance)));
        if (v.shiftMRec[k].transmittance.maxc() < (Real)1e-20) v.shiftMRec[k].transmittance = V3<Real>();
      }
    }
    return v;
  }

  // PointKDTree::executeQuery predicate, include/mitsuba/core/kdtree.h:721-723
  bool sphereTest(const V3<Real> &queryPos, Real r, const V3<Real> &p) const {
    return (p - queryPos).lengthSquared() < r * r;
  }

  // VolumeGradientPositionQuery::operator(), shift_volume_photon.cpp:489-655.
  // 1 = found by the range query but filtered, 2 = contributes.
  int vpmFunctor(const CamRay<Real> &ray, const VpmSample &s, const Photon<Real> &ph, Accum<Real> &acc) const {
    const Real r = s.radius;
    const V3<Real> pos = ray.o + s.t * ray.d;
    const Real lengthSqr = (pos - ph.pos).lengthSquared();
    if ((r * r - lengthSqr) < 0) return 1;                                   // :497-500
    if (cfg.max_depth > 0 && ray.edgeId + ph.depth > cfg.max_depth) return 1;  // :503-505
    if (!lightingModeAccepts(ph.parentType)) return 1;                       // :508-510
    const V3<Real> wi = normalize(ph.parentPos - ph.pos);
    V3<Real> photonContrib = volumePhotonContrib(ph.flux, wi, -ray.d);
    V3<Real> baseContrib = (ray.eye * s.T) * photonContrib;                  // :529
    const Real kernelVol = (Real)((4.0 / 3.0) * (double)Consts<Real>::pi * std::pow((double)r, 3));
    const Real pdfBase = s.pdfSuccess * s.pdfSel;                            // pdfBaseRay(), .h:189-192
    const Real norm = kernelVol * pdfBase;
    acc.mediumFlux += baseContrib / norm;
    for (int k = 0; k < 4; ++k) {
      GradientSamplingResult<Real> res;
      if (s.validShiftDist[k]) {
        const Real pdfShift = s.shiftMRec[k].pdfSuccess * s.pdfSel;          // pdfShiftRay(), .h:193-197
        const V3<Real> zShift = ray.offO[k] + s.t * ray.offD[k];
        bool alreadyShifted = false;
        if (cfg.use_shift_null) {                                            // :584-602
          Real distSqr = (ph.pos - zShift).lengthSquared();
          if (distSqr < r * r) {
            alreadyShifted = true;
            shiftNull(ph, wi, ray, k, s.shiftMRec[k], res, pdfBase, pdfShift);
          }
        }
        if (!alreadyShifted) {                                               // :604-639
          V3<Real> offsetPos = getShiftPos(ray, k, s.t, ph.pos, r, false);
          shiftPhotonDiffuse(ph, offsetPos, ray, k, s.shiftMRec[k], res, pdfBase, pdfShift);
        }
      } else {
        res.weight = 1;
      }
      if ((k == 1 && ray.px == cfg.film_w - 1) || (k == 2 && ray.py == cfg.film_h - 1)) res.weight = 1;
      acc.shifted[k] += (res.weight * res.shiftedFlux) / norm;
      acc.weighted[k] += (res.weight * baseContrib) / norm;
    }
    return 2;
  }

  // ---- G-Beams 3D ("beam3d" = EBeamBeam3D_Optimized) ----------------------------------------
  struct Beam {  // LTPhotonBeam + the parent-vertex data the reconnection reads
    V3<Real> o, dir, end, flux, prefix, pn, albedo, pred, endN;
    Real length, parentPdf, rrWeight;
    int parentType, depth;
    bool endOnSurface;
    uint32_t pathId;
  };
  static Beam loadBeam(const gvpm_beam_soa &s, size_t i) {
    Beam b;
    b.o = V3<Real>(s.origin + 3 * i);
    b.end = V3<Real>(s.end + 3 * i);
    b.dir = b.end - b.o;             // PhotonBeam::setEndPoint, beams_struct.h:73-81
    b.length = b.dir.length();
    b.dir = b.dir / b.length;
    b.flux = V3<Real>(s.flux + 3 * i);
    b.prefix = V3<Real>(s.prefix_flux + 3 * i);
    b.pn = V3<Real>(s.parent_n + 3 * i);
    b.albedo = V3<Real>(s.parent_albedo + 3 * i);
    b.pred = V3<Real>(s.pred_pos + 3 * i);
    b.endN = V3<Real>(s.end_n + 3 * i);
    b.parentPdf = (Real)s.parent_pdf[i];
    b.rrWeight = (Real)s.rr_weight[i];
    b.parentType = s.parent_type[i];
    b.endOnSurface = s.end_on_surface[i] != 0;
    b.depth = s.depth[i];
    b.pathId = s.path_id[i];
    return b;
  }

  // coordinateSystem, src/libcore/util.cpp:600-609 (Frame(n) constructor)
  static void coordinateSystem(const V3<Real> &a, V3<Real> &b, V3<Real> &c) {
    if (std::abs(a.x) > std::abs(a.y)) {
      Real invLen = (Real)1 / std::sqrt(a.x * a.x + a.z * a.z);
      c = V3<Real>(a.z * invLen, 0, -a.x * invLen);
    } else {
      Real invLen = (Real)1 / std::sqrt(a.y * a.y + a.z * a.z);
      c = V3<Real>(0, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
  }

  // solveQuadraticDouble, src/libcore/util.cpp:487-525
  static bool solveQuadraticDouble(double a, double b, double c, double &x0, double &x1) {
    if (a == 0) {
      if (b != 0) { x0 = x1 = -c / b; return true; }
      return false;
    }
    double discrim = b * b - 4.0f * a * c;
    if (discrim < 0) return false;
    double temp, sqrtDiscrim = std::sqrt(discrim);
    if (b < 0) temp = -0.5f * (b - sqrtDiscrim); else temp = -0.5f * (b + sqrtDiscrim);
    x0 = temp / a;
    x1 = c / temp;
    if (x0 > x1) std::swap(x0, x1);
    return true;
  }

  // cylinderIntersection, photonmapper/beams_3d_intersections.h:77-140.  The cylinder is the segment
  // (co, cd, [0, cMaxt]) with radius `radius`; the "view" ray is (vo, vd, maxt = vMaxt).  The world ->
  // cylinder-frame transform follows the reference's 4x4 arithmetic (Transform::translate * Transform::fromFrame,
  // stored inverse applied to the view ray) so that tNear / tFar are bit-identical (tests/test_oracle_ref_pin.py).
  static bool cylinderIntersection(const V3<Real> &co, const V3<Real> &cd, Real cMaxt, const V3<Real> &vo,
                                   const V3<Real> &vd, Real vMaxt, Real radius, double &tNear, double &tFar) {
    const V3<Real> d1d2c = cross(vd, cd);
    const float sinThetaSqr = (float)dot(d1d2c, d1d2c);
    const float ad = (float)dot(co - vo, d1d2c);
    if (ad * ad >= (radius * radius) * sinThetaSqr) return false;
    V3<Real> s, t;
    coordinateSystem(cd, s, t);
    // worldToObject = (translate(co) * fromFrame(Frame(cd))).inverse(): the 4x4 product (matrix.h:744-756,
    // transform.cpp:28-45,216-227) keeps the rotation rows and puts -dot(axis, co) in the last column, and
    // Transform::operator()(Point) (transform.h:108-125) adds that column last
    const V3<Real> lo(dot(vo, s) - dot(co, s), dot(vo, t) - dot(co, t), dot(vo, cd) - dot(co, cd)),
        ld(dot(vd, s), dot(vd, t), dot(vd, cd));
    const Real lMax = cMaxt;
    const double ox = lo.x, oy = lo.y, dx = ld.x, dy = ld.y;
    const double A = dx * dx + dy * dy;
    const double B = 2 * (dx * ox + dy * oy);
    const double C = ox * ox + oy * oy - radius * radius;
    if (!solveQuadraticDouble(A, B, C, tNear, tFar)) return false;
    if (tNear > vMaxt || tFar < 0) return false;
    const double zPosNear = lo.z + ld.z * tNear;
    const double zPosFar = lo.z + ld.z * tFar;
    if (zPosNear < 0) {
      if (zPosFar < 0) return false;
      float th = (float)(tNear + (tFar - tNear) * (zPosNear) / (zPosNear - zPosFar));
      tNear = th;
      return true;
    } else if (zPosNear >= 0 && zPosNear < lMax) {
      return true;
    } else if (zPosNear > lMax) {
      if (zPosFar > lMax) return false;
      float th = (float)(tNear + (tFar - tNear) * (zPosNear - lMax) / (zPosNear - zPosFar));
      tNear = th;
      return true;
    }
    return false;
  }

  // The two uniforms per (camera ray, beam) that replace sampler->next1D()
  static uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
  }
  Real beamUniform(const CamRay<Real> &ray, uint32_t beamIndex, uint32_t dim) const {
    uint32_t h = hash32(cfg.rng_seed ^ 0x9E3779B9u);
    h = hash32(h ^ (uint32_t)ray.px);
    h = hash32(h ^ ((uint32_t)ray.py * 0x85EBCA6Bu));
    h = hash32(h ^ ((uint32_t)ray.edgeId * 0xC2B2AE35u));
    h = hash32(h ^ beamIndex);
    h = hash32(h ^ (dim * 0x27D4EB2Fu));
    return (Real)((float)(h >> 8) * (1.0f / 16777216.0f));
  }

  // BeamKernelRecord, gvpm/shift/shift_volume_beams.h:24-288
  struct BeamKernelRecord {
    Real v = 0, w = 0, pdfKernel = 0, pdfEdgeFailure = 0, weightKernel = 0, u = 0;
    V3<Real> beamTrans, contrib;
    bool isValid() const { return !(contrib.x == 0 && contrib.y == 0 && contrib.z == 0); }
    Real pdf() const { return pdfEdgeFailure * pdfKernel; }
  };

  // PhotonBeam::rayIntersectInternal1D, photonmapper/beams_struct.h:250-311 (code "taken from smallUPDT"): closest
  // approach of the camera line and the beam line.  The intermediates are `float` whatever Float is.
  static bool beamIntersect1D(const V3<Real> &p1, const V3<Real> &bdir, Real blen, Real radius, const V3<Real> &ro,
                              const V3<Real> &rd, Real rmint, Real rmaxt, Real tminBeam, Real tmaxBeam, Real &u, Real &v,
                              Real &w, Real &sinTheta) {
    const V3<Real> d1d2c = cross(rd, bdir);
    const float sinThetaSqr = (float)dot(d1d2c, d1d2c);
    const float ad = (float)dot(p1 - ro, d1d2c);
    if (ad * ad >= (radius * radius) * sinThetaSqr) return false;
    const float d1d2 = (float)dot(rd, bdir);
    const float d1d2Sqr = d1d2 * d1d2;
    const float d1d2SqrMinus1 = d1d2Sqr - 1.0f;
    if (d1d2SqrMinus1 < 1e-5f && d1d2SqrMinus1 > -1e-5f) return false;
    const float d1O1 = (float)dot(rd, ro);
    const float d1O2 = (float)dot(rd, p1);
    w = (d1O1 - d1O2 - d1d2 * (dot(bdir, ro) - dot(bdir, p1))) / d1d2SqrMinus1;
    if (w <= rmint || w >= rmaxt) return false;
    v = (w + d1O1 - d1O2) / d1d2;
    if (v <= 0.0 || v >= blen || std::isnan(v)) return false;
    if (tminBeam >= v || tmaxBeam < v) return false;
    const float sinThetaConst = std::sqrt(sinThetaSqr);
    u = std::abs(ad) / sinThetaConst;
    sinTheta = sinThetaConst;
    return true;
  }

  // BeamKernelRecord::eval, EBeamBeam1D branch (shift_volume_beams.h:169-194) + PhotonBeam::getContrib
  // (beams_struct.h:136-185), for the whole beam (tmin = 0, tmax = length): every sub-beam accepts v in (t1, t2],
  // so exactly one of them owns the hit.
  BeamKernelRecord beamKernelEval1D(const Beam &beam, const V3<Real> &camO, const V3<Real> &camD, Real camMint,
                                    Real camMaxt) const {
    BeamKernelRecord k;
    if (!beamIntersect1D(beam.o, beam.dir, beam.length, radius, camO, camD, camMint, camMaxt, (Real)0, beam.length,
                         k.u, k.v, k.w, k.pdfKernel))
      return k;
    typename Medium<Real>::Rec mRecCamera = medium.eval(0, k.w), mRec = medium.eval(0, k.v);
    k.weightKernel = (Real)0.5f / radius;
    k.beamTrans = mRec.transmittance;
    const Real phaseTerm = medium.phase(-beam.dir, -camD);
    V3<Real> beamContrib = (((mRec.transmittance * mRecCamera.transmittance) * medium.sigmaS) * beam.flux) * phaseTerm;
    if (!cfg.long_beams) {
      if (mRec.pdfFailure == 0 && !(mRec.transmittance.x == 0 && mRec.transmittance.y == 0 && mRec.transmittance.z == 0)) {
        k.pdfEdgeFailure = mRec.pdfFailure;
        return k;  // contrib stays 0: invalid record
      }
      beamContrib = beamContrib / mRec.pdfFailure;
      k.pdfEdgeFailure = mRec.pdfFailure;
    } else {
      k.pdfEdgeFailure = 1;
    }
    k.contrib = beamContrib;
    if (!(k.contrib.x == 0 && k.contrib.y == 0 && k.contrib.z == 0)) k.contrib = k.contrib / k.pdfKernel;
    return k;
  }

  // shift() + localMatrix(), shift_volume_beams.cpp:36-80: the point at distance u from the camera ray (at camera
  // distance w) in the plane through the beam origin `a`; Frame{r.d, s, t} is brace-initialised, so its (s, t, n)
  // are (r.d, s, t).
  static V3<Real> beamShift1D(const V3<Real> &ro, const V3<Real> &rd, const V3<Real> &a, Real u, Real w, bool flip) {
    const Real d0 = dot(a - ro, rd);
    const V3<Real> s = normalize(a - (ro + d0 * rd));
    const V3<Real> t = cross(rd, s);
    const Real d = dot(a - ro, rd);
    const V3<Real> tD = ro + d * rd;
    const V3<Real> rel = a - tD;
    const Real localAy = dot(rel, s);
    const Real q = u / std::abs(localAy);
    Real phi = (Real)(1.57079632679489661923 - (double)std::asin(std::min((Real)1, std::max((Real)-1, q))));
    if (flip) phi = -phi;
    const Real ly = u * std::cos(phi), lz = u * std::sin(phi);
    const V3<Real> worldU = (rd * (Real)0 + s * ly) + t * lz;
    return (ro + w * rd) + worldU;
  }
  // BeamGradRadianceQuery::getShiftPos1D, shift_volume_beams.cpp:81-96
  static V3<Real> beamShiftPos1D(const V3<Real> &bo, const V3<Real> &bd, const V3<Real> &so, const V3<Real> &sd,
                                 const V3<Real> &a, const V3<Real> &bBeamDir, Real w, Real u) {
    const V3<Real> baseShiftedBack = normalize(beamShift1D(bo, bd, a, u, w, false) - a);
    const bool flipAngle = (baseShiftedBack - bBeamDir).lengthSquared() > (Real)0.001;
    return beamShift1D(so, sd, a, u, w, flipAngle);
  }

  // BeamKernelRecord::eval, EBeamBeam3D_Optimized branch (shift_volume_beams.h:195-283), for the whole
  // beam (tmin = 0, tmax = length): the per-sub-beam ownership rule (:214-220) then reads
  // "tNear < 0 or 0 < tNear < length" (DESIGN.md §6).
  BeamKernelRecord beamKernelEval(const Beam &beam, const V3<Real> &camO, const V3<Real> &camD, Real camMint,
                                  Real camMaxt, Real xi1, Real xi2) const {
    BeamKernelRecord k;
    const Real r = radius;
    const V3<Real> camStart = camO + camMint * camD;
    double tNearBeam, tFarBeam;
    if (!cylinderIntersection(camStart, camD, camMaxt - camMint, beam.o, beam.dir, beam.length, r, tNearBeam, tFarBeam))
      return k;
    if (tNearBeam < 0) {
    } else if (tNearBeam > 0 && tNearBeam < beam.length) {
    } else {
      return k;
    }
    k.v = (Real)(tNearBeam + (tFarBeam - tNearBeam) * xi1);
    k.pdfKernel = (Real)(1.0 / std::max(tFarBeam - tNearBeam, 0.0001));
    if (k.v < 0 || k.v > beam.length) return k;
    const V3<Real> kernelCentroid = beam.o + beam.dir * k.v;
    const Real distToProj = dot(kernelCentroid - camO, camD);
    const Real distSqr = ((camO + distToProj * camD) - kernelCentroid).lengthSquared();
    const Real radSqr = r * r;
    if (distSqr >= radSqr) return k;
    const Real deltaT = safe_sqrt(radSqr - distSqr);
    k.w = distToProj - deltaT + 2 * deltaT * xi2;
    k.pdfKernel = (Real)(k.pdfKernel * (1.0 / std::max(2.0 * deltaT, 0.0001)));
    if (k.w < camMint || k.w > camMaxt) return k;
    typename Medium<Real>::Rec mRecBeam = medium.eval(0, k.v), mRecCamera = medium.eval(0, k.w);
    const Real phaseTerm = medium.phase(-beam.dir, -camD);
    const Real kernelVol = (Real)((4.0 / 3.0) * (double)Consts<Real>::pi * std::pow((double)r, 3));
    k.contrib = ((((beam.flux * mRecBeam.transmittance) * medium.sigmaS) * mRecCamera.transmittance) * phaseTerm) /
                k.pdfKernel;
    k.weightKernel = (Real)(1.0 / kernelVol);
    k.beamTrans = mRecBeam.transmittance;
    if (!cfg.long_beams) {
      k.contrib = k.contrib / mRecBeam.pdfFailure;
      k.pdfEdgeFailure = mRecBeam.pdfFailure;
    } else {
      k.pdfEdgeFailure = 1;
    }
    return k;
  }

  // BeamKernelRecord(ori, medium, beam, cameraRay): the null-shift re-evaluation, shift_volume_beams.h:39-143
  BeamKernelRecord beamKernelNull(const BeamKernelRecord &ori, const Beam &beam, const V3<Real> &camO,
                                  const V3<Real> &camD, Real camMint, Real camMaxt) const {
    BeamKernelRecord k;
    const Real r = radius;
    const V3<Real> camStart = camO + camMint * camD;
    double tNearBeam, tFarBeam;
    if (!cylinderIntersection(camStart, camD, camMaxt - camMint, beam.o, beam.dir, beam.length, r, tNearBeam, tFarBeam))
      return k;
    k.v = ori.v;
    k.pdfKernel = (Real)(1.0 / std::max(tFarBeam - tNearBeam, 0.0001));
    if (k.v < 0 || k.v > beam.length) return k;
    const V3<Real> kernelCentroid = beam.o + beam.dir * k.v;
    const Real distToProj = dot(kernelCentroid - camO, camD);
    const Real distSqr = ((camO + distToProj * camD) - kernelCentroid).lengthSquared();
    const Real radSqr = r * r;
    if (distSqr >= radSqr) return k;
    const Real deltaT = safe_sqrt(radSqr - distSqr);
    k.w = ori.w;
    k.pdfKernel = (Real)(k.pdfKernel * (1.0 / std::max(2.0 * deltaT, 0.0001)));
    if (k.w < camMint || k.w > camMaxt) return k;
    k.contrib = ori.contrib * (ori.pdfKernel / k.pdfKernel);
    k.weightKernel = ori.weightKernel;
    k.beamTrans = ori.beamTrans;
    k.pdfEdgeFailure = cfg.long_beams ? (Real)1 : ori.pdfEdgeFailure;
    return k;
  }

  // BeamKernelRecord::kernelPDF, shift_volume_beams.h:298-336 (3-D optimized)
  Real beamKernelPDF(const V3<Real> &camO, const V3<Real> &camD, Real camMaxt, const V3<Real> &orgBeam,
                     const V3<Real> &dBeam, Real newDLength) const {
    if (cfg.beam_kernel_1d) return std::sqrt(cross(camD, dBeam).lengthSquared());   // :299-300
    const Real r = radius;
    double tNearBeam, tFarBeam;
    if (cylinderIntersection(camO, camD, camMaxt, orgBeam, dBeam, (Real)INFINITY, r, tNearBeam, tFarBeam)) {
      Real pdfK = (Real)(1.0 / std::max(tFarBeam - tNearBeam, 0.0001));
      const V3<Real> kernelCentroid = orgBeam + dBeam * newDLength;
      const Real distToProj = dot(kernelCentroid - camO, camD);
      const Real distSqr = ((camO + distToProj * camD) - kernelCentroid).lengthSquared();
      const Real radSqr = r * r;
      if (distSqr < radSqr) {
        const Real deltaT = safe_sqrt(radSqr - distSqr);
        pdfK = (Real)(pdfK * (1.0 / std::max(2.0 * deltaT, 0.0001)));
        return pdfK;
      }
      return 0;
    }
    return 0;
  }

  // BeamGradRadianceQuery::getShiftPos, shift_volume_beams.cpp:98-137 (coherent = true)
  V3<Real> beamShiftPos(const CamRay<Real> &ray, int k, Real w, const V3<Real> &u, Real newW) const {
    const V3<Real> sAt = ray.offO[k] + newW * ray.offD[k];
    V3<Real> bs, bt, ns, nt;
    coordinateSystemCoherent(ray.d, bs, bt);
    coordinateSystemCoherent(ray.offD[k], ns, nt);
    const V3<Real> local(dot(u, bs), dot(u, bt), dot(u, ray.d));
    V3<Real> newPos = sAt + ((ns * local.x + nt * local.y) + ray.offD[k] * local.z);
    if (cfg.use_shift_null) {
      const V3<Real> bCamW = ray.o + w * ray.d;
      Real offDistSqr = (bCamW - newPos).lengthSquared();
      if (offDistSqr < radius * radius) {
        V3<Real> dShift = sAt - bCamW;
        dShift = dShift / dShift.length();
        const Real cosD = dot(dShift, -(newPos - sAt));
        newPos += (dShift * cosD) * (Real)2;
      }
    }
    return newPos;
  }

  // shiftBeamDiffuse + diffuseReconnectionPhotonBeam: shift_volume_beams.cpp:410-539,
  // shift/operation/shift_diffuse.cpp:136-268
  void shiftBeamDiffuse(const Beam &beam, const CamRay<Real> &ray, int k, Real shiftW, const BeamKernelRecord &kRec,
                        const V3<Real> &newPos, GradientSamplingResult<Real> &res) const {
    if (shiftW > ray.offLen[k]) { res.weight = 1; return; }                    // shiftBeam :364-367
    if (beam.parentType == GVPM_PARENT_OTHER) return;                          // manifold: out of scope
    V3<Real> newPBDir = newPos - beam.o;
    const Real newPBDist = newPBDir.length();
    newPBDir = newPBDir / newPBDist;
    if (occ.anyHit(beam.o, newPBDir, (Real)cfg.epsilon, newPBDist)) { res.weight = 1; return; }  // :420-426
    const V3<Real> basePos = beam.o + beam.dir * kRec.v;
    V3<Real> shiftPhotonWeight = beam.prefix;
    const Real pdfKernelAndDist = kRec.pdf();
    // diffuseReconnectionPhotonBeam
    V3<Real> thr(1, 1, 1);
    Real pdfValueSA = 0, sPdf = 0;
    const Real INV_PI = Consts<Real>::inv_pi;
    bool failed = false;
    if (beam.parentType == GVPM_PARENT_SURFACE) {
      V3<Real> wiWorld = normalize(beam.pred - beam.o);
      Real cosI = dot(beam.pn, wiWorld), cosO = dot(beam.pn, newPBDir);
      if (cosI <= 0 || cosO <= 0) { thr = V3<Real>(); pdfValueSA = 0; }
      else { thr = thr * (beam.albedo * (INV_PI * cosO)); pdfValueSA = INV_PI * cosO; }
      if (cosI * cosI <= 0 || cosO * cosO <= 0) failed = true;
    } else if (beam.parentType == GVPM_PARENT_MEDIUM) {
      V3<Real> pWi = normalize(beam.pred - beam.o);
      Real phv = medium.phase(pWi, newPBDir);
      thr = thr * (medium.sigmaS * phv);
      pdfValueSA = phv;
    } else {
      Real dp = dot(newPBDir, beam.pn);
      if (dp < 0) dp = 0;
      thr = thr * V3<Real>(INV_PI * dp, INV_PI * dp, INV_PI * dp);
      pdfValueSA = INV_PI * dp;
    }
    if (!failed) {
      const Real GOpNew = 1 / (newPBDist * newPBDist);
      sPdf = pdfValueSA * GOpNew;
      thr = thr * GOpNew;
      Real pdfBasePos = beam.parentPdf * (beam.o - beam.end).lengthSquared();
      if (beam.endOnSurface) pdfBasePos /= std::abs(dot(beam.endN, beam.dir));
      const Real GOpBase = (Real)1 / (beam.o - basePos).lengthSquared();
      pdfBasePos *= GOpBase;
      if (pdfBasePos == 0) {
        sPdf = 0;
      } else {
        thr = thr / pdfBasePos;
        thr = thr * beam.rrWeight;
        typename Medium<Real>::Rec m = medium.eval(0, newPBDist);
        if (!cfg.long_beams) sPdf *= m.pdfFailure;
        thr = thr * (m.transmittance / pdfKernelAndDist);
      }
    }
    if (sPdf == 0) { res.weight = 1; return; }
    const Real shiftKernelPDF = beamKernelPDF(ray.offO[k], ray.offD[k], ray.offLen[k], beam.o, newPBDir, newPBDist);
    if (shiftKernelPDF == 0) { res.weight = 1; return; }
    shiftPhotonWeight = shiftPhotonWeight * thr;
    typename Medium<Real>::Rec mRecShift = medium.eval(0, shiftW);
    const Real phaseTerm = medium.phase(-newPBDir, -ray.offD[k]);
    shiftPhotonWeight = shiftPhotonWeight * ((mRecShift.transmittance * medium.sigmaS) * phaseTerm);
    res.shiftedFlux = (shiftPhotonWeight * ray.offEye[k]) * res.jacobian;
    res.weight = (Real)0.5;
    if (cfg.use_mis) {
      Real basePdf = beam.parentPdf;
      basePdf *= (beam.o - beam.end).lengthSquared();
      if (beam.endOnSurface) basePdf /= std::abs(dot(beam.endN, beam.dir));
      basePdf /= (beam.o - basePos).lengthSquared();
      basePdf *= pdfKernelAndDist;
      Real offsetPdf = shiftKernelPDF;
      offsetPdf *= sPdf;
      if (offsetPdf == 0 || basePdf == 0) { res.weight = 1; return; }
      const Real q = ray.offSensor[k] * offsetPdf * res.jacobian / basePdf;
      res.weight = cfg.power_heuristic ? (Real)1 / ((Real)1 + q * q) : (Real)1 / ((Real)1 + q);
    }
  }

  // BeamGradRadianceQuery::operator(), shift_volume_beams.cpp:139-353 (beam3d, or beam1d with newShiftBeam).
  // 0 = no valid kernel record, 1 = valid but filtered, 2 = contributes.
  int beamFunctor(const CamRay<Real> &ray, const Beam &beam, uint32_t beamIndex, Accum<Real> &acc) const {
    bool filtered = false;
    if (cfg.max_depth > 0 && ray.edgeId + beam.depth > cfg.max_depth) filtered = true;   // :143-145
    if (!lightingModeAccepts(beam.parentType)) filtered = true;                         // :148-150
    Real rrGlobalWeight = 1;
    if (cfg.path_set) {                                                                 // :180-187
      if (beam.pathId % 2 != (uint32_t)((ray.px + ray.py) % 2)) filtered = true;
      rrGlobalWeight = 2;
    }
    const bool k1d = cfg.beam_kernel_1d != 0;
    const Real xi1 = beamUniform(ray, beamIndex, 0), xi2 = beamUniform(ray, beamIndex, 1);
    const BeamKernelRecord kRec = k1d ? beamKernelEval1D(beam, ray.o, ray.d, ray.mint, ray.maxt)
                                      : beamKernelEval(beam, ray.o, ray.d, ray.mint, ray.maxt, xi1, xi2);
    if (!kRec.isValid()) return 0;
    if (filtered) return 1;
    const Real r = radius;
    const V3<Real> baseContrib = (ray.eye * kRec.contrib) * kRec.weightKernel;          // :205
    acc.mediumFlux += baseContrib * rrGlobalWeight;
    for (int k = 0; k < 4; ++k) {
      GradientSamplingResult<Real> res;
      if (ray.offValid[k]) {
        const Real shiftDistMAX = ray.offLen[k], shiftW = kRec.w;
        bool alreadyShift = false;
        if (cfg.use_shift_null && !k1d) {                               // :254-289 ("Ignored in case of Beam 1D kernel")
          const V3<Real> kernelPos = beam.o + beam.dir * kRec.v;
          const Real ZPtoY = ((ray.offO[k] + shiftW * ray.offD[k]) - kernelPos).lengthSquared();
          if (ZPtoY < r * r && kRec.w <= shiftDistMAX) {
            BeamKernelRecord kS = beamKernelNull(kRec, beam, ray.offO[k], ray.offD[k], (Real)cfg.epsilon, shiftDistMAX);
            if (kS.isValid()) {
              // shiftNull3D, :748-786
              kS.contrib = kS.contrib * (kS.pdf() / kRec.pdf());
              res.jacobian = 1;
              res.shiftedFlux = (kS.contrib * ray.offEye[k]) * res.jacobian;
              res.weight = (Real)0.5;
              if (cfg.use_mis) {
                const Real basePdf = kRec.pdf(), offsetPdf = kS.pdf();
                if (offsetPdf == 0 || basePdf == 0) {
                  res.weight = 1;
                } else {
                  const Real q = ray.offSensor[k] * res.jacobian * (offsetPdf / basePdf);
                  res.weight = cfg.power_heuristic ? (Real)1 / ((Real)1 + q * q) : (Real)1 / ((Real)1 + q);
                }
              }
              alreadyShift = true;
            }
          }
        }
        if (!alreadyShift && kRec.w <= shiftDistMAX && k1d) {                           // newShiftBeam, :311-317
          const V3<Real> offsetPos = beamShiftPos1D(ray.o, ray.d, ray.offO[k], ray.offD[k], beam.o, beam.dir, kRec.w, kRec.u);
          shiftBeamDiffuse(beam, ray, k, shiftW, kRec, offsetPos, res);
        } else if (!alreadyShift && kRec.w <= shiftDistMAX) {                           // :293-310
          const Real dd = dot(beam.o - ray.offO[k], ray.offD[k]);
          const Real minDistSqr = (beam.o - (ray.offO[k] + dd * ray.offD[k])).lengthSquared();
          if (minDistSqr > 0) {  // kRec.u == 0 for the 3-D kernel
            const V3<Real> u = (beam.o + beam.dir * kRec.v) - (ray.o + kRec.w * ray.d);
            const V3<Real> offsetPos = beamShiftPos(ray, k, kRec.w, u, shiftW);
            shiftBeamDiffuse(beam, ray, k, shiftW, kRec, offsetPos, res);
          } else {
            res.weight = 1;
          }
        }
      } else {
        res.weight = 1;
      }
      res.shiftedFlux = res.shiftedFlux * kRec.weightKernel;                            // :337
      if ((k == 1 && ray.px == cfg.film_w - 1) || (k == 2 && ray.py == cfg.film_h - 1)) res.weight = 1;
      acc.shifted[k] += (res.weight * res.shiftedFlux) * rrGlobalWeight;
      acc.weighted[k] += (res.weight * baseContrib) * rrGlobalWeight;
    }
    return 2;
  }

  // ---- sppm primal photon beams: BeamRadianceQuery::operator(), photonmapper/beams.h:29-223, driven as
  // volumePhotonBeamPass (sppm.cpp:823-860); all four techniques of EVolumeTechnique (1D, 3D naive, 3D EGSR,
  // 3D optimized).  One call = one (camera beam, sub-beam [tmin, tmax]) visit of SubBeamBVH::query
  // (beams_accel.h:169-203).  `first` / `last` mark the beam's first / last sub-beam; `subIndex` its ordinal.
  // The sampler->next1D() draws are replaced by the counter-based hash (dims 0,1 per (ray, beam); the naive
  // technique samples per sub-beam: dims 2+2k, 3+2k).  Deviations: sub-beam ownership half-open
  // as in the gvpm restatement; the naive branch, which the reference lets through without a camera range test
  // (so that its result depends on which subtree boxes the ray happens to cross), gets the same
  // [mint, maxt] test as the other two 3-D branches (:160-162).
  // 0 = rejected, 1 = accepted but dropped by the depth filters (:34-39), 2 = contributes (Li += ...).
  int sppmBeamFunctor(const CamRay<Real> &ray, const Beam &beam, uint32_t beamIndex, Real tmin, Real tmax, bool first,
                      bool last, uint32_t subIndex, int technique, V3<Real> &Li) const {
    if (tmax > beam.length) tmax = beam.length;                                          // :30-32
    const int maxDepthQ = cfg.max_depth == -1 ? -1 : cfg.max_depth - ray.edgeId;         // sppm.cpp:853
    const int minDepthQ = std::max(0, cfg.min_depth - ray.edgeId);                       // sppm.cpp:854
    bool filtered = false;
    if (maxDepthQ != -1 && beam.depth > maxDepthQ) filtered = true;
    if (minDepthQ != 0 && beam.depth < minDepthQ) filtered = true;
    const Real r = radius, eps = (Real)cfg.epsilon;
    if (technique == GVPM_BEAM_1D) {                                                     // :41-68
      Real u, v, w, sinTheta;
      if (!beamIntersect1D(beam.o, beam.dir, beam.length, r, ray.o, ray.d, ray.mint, ray.maxt, first ? (Real)0 : tmin,
                           last ? beam.length : tmax, u, v, w, sinTheta))
        return 0;
      if (r <= u) return 0;
      if (filtered) return 1;
      typename Medium<Real>::Rec mRecCamera = medium.eval(eps, w), mRec = medium.eval(0, v);
      const Real weightKernel = (Real)0.5f / r;
      V3<Real> beamContrib =
          (((mRec.transmittance * mRecCamera.transmittance) * medium.sigmaS) * beam.flux) * medium.phase(-beam.dir, -ray.d);
      if (!cfg.long_beams) {                                                            // getContrib, beams_struct.h:157-172
        const bool tZero = mRec.transmittance.x == 0 && mRec.transmittance.y == 0 && mRec.transmittance.z == 0;
        if (mRec.pdfFailure == 0 && !tZero) return 2;   // contributes Spectrum(0)
        beamContrib = beamContrib / mRec.pdfFailure;
      }
      Li += ((beamContrib * weightKernel) / sinTheta) * ray.eye;
      return 2;
    }
    Real beamSegmentRand, cameraSegmentRand, invPDF;
    const Real radSqr = r * r;
    if (technique == GVPM_BEAM_3D_NAIVE) {                                               // :77-102
      const Real xi1 = beamUniform(ray, beamIndex, 2 + 2 * subIndex), xi2 = beamUniform(ray, beamIndex, 3 + 2 * subIndex);
      beamSegmentRand = tmin + (tmax - tmin) * xi1;
      invPDF = tmax - tmin;
      const V3<Real> kernelCentroid = beam.o + beam.dir * beamSegmentRand;
      const Real distToProj = dot(kernelCentroid - ray.o, ray.d);
      const Real distSqr = ((ray.o + distToProj * ray.d) - kernelCentroid).lengthSquared();
      if (distSqr >= radSqr) return 0;
      const Real deltaT = safe_sqrt(radSqr - distSqr);
      cameraSegmentRand = (distToProj - deltaT) + 2 * deltaT * xi2;
      invPDF = (Real)(invPDF * std::max(2.0 * deltaT, 0.0001));
      if (cameraSegmentRand < ray.mint || cameraSegmentRand > ray.maxt) return 0;       // deviation, see above
    } else {
      const Real xi1 = beamUniform(ray, beamIndex, 0), xi2 = beamUniform(ray, beamIndex, 1);
      const V3<Real> camStart = ray.o + ray.mint * ray.d;                                // _cam, :106-107
      double tNearBeam, tFarBeam;
      if (!cylinderIntersection(camStart, ray.d, ray.maxt - ray.mint, beam.o, beam.dir, beam.length, r, tNearBeam, tFarBeam))
        return 0;
      // ownership (:122-128) with half-open cuts: [tmin, tmax), the first sub-beam also takes tNear < 0
      if (!((first || tNearBeam >= (double)tmin) && (last || tNearBeam < (double)tmax))) return 0;
      if (!(tNearBeam < 0 || (tNearBeam > 0 && tNearBeam < beam.length))) return 0;
      beamSegmentRand = (Real)(tNearBeam + (tFarBeam - tNearBeam) * xi1);
      invPDF = (Real)std::max(tFarBeam - tNearBeam, 0.0001);
      if (beamSegmentRand < 0 || beamSegmentRand > beam.length) return 0;
      if (technique == GVPM_BEAM_3D_EGSR) {                                              // :138-150
        double tNearCam, tFarCam;
        if (!cylinderIntersection(beam.o, beam.dir, beam.length, camStart, ray.d, ray.maxt - ray.mint, r, tNearCam, tFarCam))
          return 0;
        cameraSegmentRand = (Real)(tNearCam + (tFarCam - tNearCam) * xi2);
        invPDF = (Real)(invPDF * std::max(tFarCam - tNearCam, 0.0001));
      } else {                                                                           // :151-170
        const V3<Real> kernelCentroid = beam.o + beam.dir * beamSegmentRand;
        const Real distToProj = dot(kernelCentroid - ray.o, ray.d);
        const Real distSqr = ((ray.o + distToProj * ray.d) - kernelCentroid).lengthSquared();
        if (distSqr >= radSqr) return 0;
        const Real deltaT = safe_sqrt(radSqr - distSqr);
        cameraSegmentRand = distToProj - deltaT + 2 * deltaT * xi2;
        invPDF = (Real)(invPDF * std::max(2.0 * deltaT, 0.0001));
      }
      if (cameraSegmentRand < ray.mint || cameraSegmentRand > ray.maxt) return 0;       // :173-175
      if (technique == GVPM_BEAM_3D_EGSR) {                                              // :178-187
        const V3<Real> kernelCentroid = beam.o + beam.dir * beamSegmentRand;
        const Real distSqr = ((ray.o + cameraSegmentRand * ray.d) - kernelCentroid).lengthSquared();
        if (distSqr >= radSqr) return 0;
      }
    }
    if (filtered) return 1;
    typename Medium<Real>::Rec mRecBeam = medium.eval(0, beamSegmentRand), mRecCamera = medium.eval(eps, cameraSegmentRand);
    const Real phaseTerm = medium.phase(-beam.dir, -ray.d);
    const Real kernelVol = (Real)((4.0 / 3.0) * (double)Consts<Real>::pi * std::pow((double)r, 3));
    V3<Real> beamContrib =
        ((((beam.flux * mRecBeam.transmittance) * medium.sigmaS) * mRecCamera.transmittance) * phaseTerm) * (invPDF / kernelVol);
    if (!cfg.long_beams) beamContrib = beamContrib / mRecBeam.pdfFailure;                // :213-217
    Li += beamContrib * ray.eye;                                                         // sppm.cpp:858
    return 2;
  }

  // ---- G-Planes 0D ("plane0d") -------------------------------------------------------------
  struct Plane {  // LTPhotonPlane, gvpm/gvpm_plane.h:18-46 + PhotonPlane, photonmapper/plane_struct.h:18-58
    V3<Real> ori, w0, w1, flux;
    Real length0, length1;
    int edgeID;
  };
  static Plane loadPlane(const gvpm_plane_soa &s, size_t i) {
    Plane p;
    p.ori = V3<Real>(s.origin + 3 * i);
    p.w0 = V3<Real>(s.w0 + 3 * i);
    p.w1 = V3<Real>(s.w1 + 3 * i);
    p.flux = V3<Real>(s.flux + 3 * i);
    p.length0 = (Real)s.length0[i];
    p.length1 = (Real)s.length1[i];
    p.edgeID = s.edge_id[i];
    return p;
  }
  struct PlaneIts { Real tCam, t0, t1, invDet; };  // PhotonPlane::IntersectionRecord, plane_struct.h:20-22
  static Real absDot(const V3<Real> &a, const V3<Real> &b) { return std::abs(dot(a, b)); }
  static V3<Real> cdiv(const V3<Real> &a, const V3<Real> &b) { return V3<Real>(a.x / b.x, a.y / b.y, a.z / b.z); }

  // PhotonPlane::intersectPlane0D, plane_struct.h:104-135 (`det` is a float whatever Float is)
  static bool intersectPlane0D(const Plane &pl, const V3<Real> &o, const V3<Real> &d, Real mint, Real maxt,
                               PlaneIts &r) {
    const V3<Real> e0 = pl.w0 * pl.length0, e1 = pl.w1 * pl.length1;
    const V3<Real> P = cross(d, e1);
    const float det = (float)dot(e0, P);
    if (std::abs(det) < 1e-5f) return false;
    r.invDet = (Real)(1.0f / det);
    const V3<Real> T = o - pl.ori;
    r.t0 = dot(T, P) * r.invDet;
    if (r.t0 < 0 || r.t0 > 1) return false;
    const V3<Real> Q = cross(T, e0);
    r.t1 = dot(d, Q) * r.invDet;
    if (r.t1 < 0 || r.t1 > 1) return false;
    r.tCam = dot(e1, Q) * r.invDet;
    if (r.tCam <= mint || r.tCam >= maxt) return false;
    r.t1 *= pl.length1;
    r.t0 *= pl.length0;
    return true;
  }
  // PhotonPlane::invJacobian, plane_struct.h:194-196 (1.0 / Float evaluated in double, rounded to Float)
  static Real planeInvJacobian(const Plane &pl, const V3<Real> &k) {
    return (Real)(1.0 / (double)absDot(pl.w0, cross(pl.w1, k)));
  }
  // PhotonPlane::getContrib0D, plane_struct.h:150-192
  V3<Real> planeContrib0D(const Plane &pl, const PlaneIts &its, const typename Medium<Real>::Rec &mRecCamera,
                          const V3<Real> &d) const {
    const Real phaseTerm = medium.phase(-pl.w1, -d);
    const typename Medium<Real>::Rec mRec0 = medium.eval(0, its.t0), mRec1 = medium.eval(0, its.t1);
    V3<Real> contrib = (((mRecCamera.transmittance * medium.sigmaS) * medium.sigmaS) * pl.flux) * phaseTerm;
    contrib = contrib * (mRec1.transmittance * mRec0.transmittance);
    contrib = contrib / mRec0.pdfFailure;
    contrib = contrib / mRec1.pdfFailure;
    contrib = contrib * planeInvJacobian(pl, d);
    return contrib;
  }
  // PlaneGradRadianceQuery::intersection, shift_volume_planes.h:426-453 (unit w0 / w1, no upper bound on t0, t1)
  static bool planeShiftIntersection(const V3<Real> &o, const V3<Real> &d, Real mint, Real maxt, const V3<Real> &ori,
                                     const V3<Real> &w0, const V3<Real> &w1, Real &tCam, Real &t0, Real &t1,
                                     Real &invDet) {
    const V3<Real> P = cross(d, w1);
    const Real det = dot(w0, P);
    if (std::abs(det) < 1e-8f) return false;
    invDet = 1.0f / det;
    const V3<Real> T = o - ori;
    t0 = dot(T, P) * invDet;
    if (t0 < 0.0f) return false;
    const V3<Real> Q = cross(T, w0);
    t1 = dot(d, Q) * invDet;
    if (t1 < 0.0f) return false;
    tCam = dot(w1, Q) * invDet;
    return !(tCam <= mint || tCam >= maxt);
  }
  // PlaneGradRadianceQuery::specularShift, shift_volume_planes.h:263-416 (BETTERSHIFT 0): origin and w0 of the
  // plane are kept, w1 is rotated so that the plane passes through the offset camera point at equal tCam.
  void planeSpecularShift(const Plane &pl, const CamRay<Real> &ray, int k, const PlaneIts &bRec,
                          const V3<Real> &baseContrib, GradientSamplingResult<Real> &res) const {
    const V3<Real> &so = ray.offO[k], &sd = ray.offD[k];
    const Real sMint = (Real)cfg.epsilon, sMaxt = ray.offLen[k];  // Ray(o, d, Epsilon, shiftDist, 0) :272-273
    const V3<Real> newIntersection = so + sd * bRec.tCam;
    V3<Real> orthNewW1 = newIntersection - (pl.ori + pl.w0 * dot(newIntersection - pl.ori, pl.w0));
    orthNewW1 = orthNewW1 / orthNewW1.length();
    const Real w0Dot = dot(pl.w0, pl.w1);
    const V3<Real> newW1 = std::sqrt(1 - (w0Dot * w0Dot)) * orthNewW1 + pl.w0 * w0Dot;
    Real t0New, t1New, tCamNew, invDetNew;
    if (!planeShiftIntersection(so, sd, sMint, sMaxt, pl.ori, pl.w0, newW1, tCamNew, t0New, t1New, invDetNew)) {
      res.weight = 1;
      return;
    }
    const typename Medium<Real>::Rec mRec1 = medium.eval(0, bRec.t1), mRec0 = medium.eval(0, bRec.t0),
                                     mRec1Shift = medium.eval(0, t1New), mRec0Shift = medium.eval(0, t0New);
    V3<Real> throughputShift = baseContrib;
    throughputShift = throughputShift * cdiv(mRec0Shift.transmittance, mRec0.transmittance);
    throughputShift = throughputShift * cdiv(mRec1Shift.transmittance, mRec1.transmittance);
    throughputShift = throughputShift / planeInvJacobian(pl, ray.d);
    throughputShift = throughputShift * (Real)(1.0 / (double)absDot(pl.w0, cross(newW1, sd)));
    res.jacobian = planeInvJacobian(pl, ray.d);
    res.jacobian *= absDot(pl.w0, cross(newW1, sd));
    res.jacobian /= t1New / bRec.t1;
    if (pl.edgeID != 1) res.jacobian /= t0New / bRec.t0;
    const Real phaseBase = medium.phase(-pl.w1, -ray.d), phaseNew = medium.phase(-newW1, -sd);
    throughputShift = throughputShift * phaseNew;
    throughputShift = throughputShift / phaseBase;
    res.weight = (Real)0.5;
    res.shiftedFlux = res.jacobian * throughputShift;
    if (cfg.use_mis) {
      Real basePdf = mRec0.pdfSuccess;
      basePdf *= mRec1.pdfSuccess;
      basePdf *= phaseBase;  // pdf == eval for the isotropic and HG phase functions
      Real offsetPdf = mRec0Shift.pdfSuccess;
      offsetPdf *= mRec1Shift.pdfSuccess;
      offsetPdf *= phaseNew;
      if (offsetPdf == 0 || basePdf == 0) {  // "Invalid path": weight 1, shiftedFlux is kept (:401-405)
        res.weight = 1;
        return;
      }
      res.weight = (Real)1 / ((Real)1 + ray.offSensor[k] * res.jacobian * offsetPdf / basePdf);
    }
  }
  // PlaneGradRadianceQuery::operator(), shift_volume_planes.h:57-101.  No eyeContrib, no depth / mode /
  // pathSet filter and no border rule in the reference functor.  Returns whether the plane was intersected.
  bool planeFunctor(const CamRay<Real> &ray, const Plane &pl, Accum<Real> &acc) const {
    PlaneIts bRec;
    if (!intersectPlane0D(pl, ray.o, ray.d, ray.mint, ray.maxt, bRec)) return false;
    const typename Medium<Real>::Rec mRecCam = medium.eval(0, bRec.tCam);
    const V3<Real> baseContrib = planeContrib0D(pl, bRec, mRecCam, ray.d);
    acc.mediumFlux += baseContrib;
    for (int k = 0; k < 4; ++k) {
      GradientSamplingResult<Real> res;
      if (ray.offValid[k]) planeSpecularShift(pl, ray, k, bRec, baseContrib, res);
      acc.shifted[k] += res.weight * res.shiftedFlux;
      acc.weighted[k] += res.weight * baseContrib;
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------
// PointKDTree<SimpleKDNode>::build with ESlidingMidpoint (include/mitsuba/core/kdtree.h:326-378,
// 921-1037) + GradientBeamRadianceEstimator::buildHierarchy (gvpm/gvpm_accel.cpp:35-55) +
// query (gvpm/gvpm_accel.h:268-312).
template <typename Real> struct BreTree {
  struct Node {
    V3<Real> pos, bmin, bmax;
    uint32_t right = 0, orig = 0;
    bool leaf = false;
    uint8_t axis = 0;
  };
  std::vector<Node> nodes;
  size_t depth = 0;
  Real radius = 0;

  void build(const gvpm_photon_soa &ph, size_t n, Real r) {
    radius = r;
    nodes.resize(n);
    if (n == 0) return;
    std::vector<V3<Real>> pts(n);
    V3<Real> amin(pts.empty() ? V3<Real>() : V3<Real>(ph.pos)), amax = amin;
    for (size_t i = 0; i < n; ++i) {
      pts[i] = V3<Real>(ph.pos + 3 * i);
      for (int a = 0; a < 3; ++a) {
        amin.at(a) = std::min(amin[a], pts[i][a]);
        amax.at(a) = std::max(amax[a], pts[i][a]);
      }
    }
    std::vector<uint32_t> ind(n);
    for (size_t i = 0; i < n; ++i) ind[i] = (uint32_t)i;
    std::vector<uint32_t> rightOf(n, 0);
    std::vector<uint8_t> leafOf(n, 0);
    depth = 0;
    buildRec(1, pts, ind, rightOf, leafOf, 0, n, amin, amax);
    for (size_t i = 0; i < n; ++i) {  // permute_inplace(indirection)
      Node &nd = nodes[i];
      nd.orig = ind[i];
      nd.pos = pts[ind[i]];
      nd.right = rightOf[ind[i]];
      nd.leaf = (leafOf[ind[i]] & 1) != 0;
      nd.axis = leafOf[ind[i]] >> 1;
    }
    hierarchy(0);
  }

  // PointKDTree::executeQuery, include/mitsuba/core/kdtree.h:675-731: visit(nodeIndex) for every node with
  // |node.pos - p|^2 < r^2; returns `found`.
  template <typename F> size_t rangeQuery(const V3<Real> &p, Real searchRadius, F &&visit) const {
    if (nodes.empty()) return 0;
    std::vector<uint32_t> stack(depth + 2);
    uint32_t index = 0, stackPos = 1, found = 0;
    const Real distSquared = searchRadius * searchRadius;
    stack[0] = 0;
    while (stackPos > 0) {
      const Node &node = nodes[index];
      uint32_t nextIndex;
      if (!node.leaf) {
        const Real distToPlane = p[node.axis] - node.pos[node.axis];
        const bool searchBoth = distToPlane * distToPlane <= distSquared;
        if (distToPlane > 0) {
          if (node.right != 0) {
            if (searchBoth) stack[stackPos++] = index + 1;
            nextIndex = node.right;
          } else if (searchBoth) {
            nextIndex = index + 1;
          } else {
            nextIndex = stack[--stackPos];
          }
        } else {
          if (searchBoth && node.right != 0) stack[stackPos++] = node.right;
          nextIndex = index + 1;
        }
      } else {
        nextIndex = stack[--stackPos];
      }
      const Real pointDistSquared = (node.pos - p).lengthSquared();
      if (pointDistSquared < distSquared) {
        ++found;
        visit(index);
      }
      index = nextIndex;
    }
    return found;
  }

  void buildRec(size_t d, const std::vector<V3<Real>> &pts, std::vector<uint32_t> &ind,
                std::vector<uint32_t> &rightOf, std::vector<uint8_t> &leafOf, size_t b, size_t e,
                V3<Real> &amin, V3<Real> &amax) {
    depth = std::max(d, depth);
    size_t count = e - b;
    if (count == 1) { leafOf[ind[b]] = 1; return; }
    V3<Real> ext = amax - amin;
    int axis = 0;
    for (int i = 1; i < 3; ++i) if (ext[i] > ext[axis]) axis = i;   // aabb.h:269-277
    Real midpoint = (Real)0.5 * (amax[axis] + amin[axis]);
    size_t nLT = 0;
    for (size_t i = b; i < e; ++i) nLT += pts[ind[i]][axis] <= midpoint;
    size_t split = b + nLT;
    if (split == b) ++split; else if (split == e) --split;
    std::nth_element(ind.begin() + b, ind.begin() + split, ind.begin() + e,
                     [&](uint32_t i1, uint32_t i2) { return pts[i1][axis] < pts[i2][axis]; });
    uint32_t sp = ind[split];
    leafOf[sp] = (uint8_t)(axis << 1);
    rightOf[sp] = (split + 1 != e) ? (uint32_t)(split + 1) : 0;
    std::swap(ind[b], ind[split]);
    Real splitPos = pts[sp][axis];
    Real temp = amax[axis];
    amax.at(axis) = splitPos;
    buildRec(d + 1, pts, ind, rightOf, leafOf, b + 1, split + 1, amin, amax);
    amax.at(axis) = temp;
    if (split + 1 != e) {
      temp = amin[axis];
      amin.at(axis) = splitPos;
      buildRec(d + 1, pts, ind, rightOf, leafOf, split + 1, e, amin, amax);
      amin.at(axis) = temp;
    }
  }

  void hierarchy(uint32_t root) {  // iterative post-order of gvpm_accel.cpp:35-55
    // children always have larger indices than their parent (left = self+1, right > self)
    for (size_t ii = nodes.size(); ii-- > 0;) {
      Node &nd = nodes[ii];
      V3<Real> rv(radius, radius, radius);
      nd.bmin = nd.pos - rv;
      nd.bmax = nd.pos + rv;
      if (!nd.leaf) {
        auto expand = [&](const Node &c) {
          for (int a = 0; a < 3; ++a) {
            nd.bmin.at(a) = std::min(nd.bmin[a], c.bmin[a]);
            nd.bmax.at(a) = std::max(nd.bmax[a], c.bmax[a]);
          }
        };
        expand(nodes[ii + 1]);
        if (nd.right) expand(nodes[nd.right]);
      }
    }
    (void)root;
  }

  // AABB::rayIntersect, include/mitsuba/core/aabb.h:310-340
  static bool slab(const V3<Real> &bmin, const V3<Real> &bmax, const V3<Real> &o, const V3<Real> &d,
                   const V3<Real> &dRcp, Real &nearT, Real &farT) {
    nearT = -INFINITY;
    farT = INFINITY;
    for (int i = 0; i < 3; ++i) {
      Real origin = o[i], minVal = bmin[i], maxVal = bmax[i];
      if (d[i] == 0) {
        if (origin < minVal || origin > maxVal) return false;
      } else {
        Real t1 = (minVal - origin) * dRcp[i], t2 = (maxVal - origin) * dRcp[i];
        if (t1 > t2) std::swap(t1, t2);
        nearT = std::max(t1, nearT);
        farT = std::min(t2, farT);
        if (!(nearT <= farT)) return false;
      }
    }
    return true;
  }

  // visit(nodeIndexInTree, diskDistance) for every photon passing the predicate
  template <typename F> void query(const Scene<Real> &sc, const CamRay<Real> &ray, F &&visit) const {
    if (nodes.empty()) return;
    std::vector<uint32_t> stack(depth + 2);
    uint32_t index = 0, stackPos = 1;
    V3<Real> dRcp((Real)1 / ray.d.x, (Real)1 / ray.d.y, (Real)1 / ray.d.z);
    while (stackPos > 0) {
      const Node &node = nodes[index];
      Real mint, maxt;
      if (!slab(node.bmin, node.bmax, ray.o, ray.d, dRcp, mint, maxt) || maxt < ray.mint || mint > ray.maxt) {
        index = stack[--stackPos];
        continue;
      }
      uint32_t self = index;
      if (!node.leaf) {
        if (node.right != 0) stack[stackPos++] = node.right;
        index = self + 1;
      } else {
        index = stack[--stackPos];
      }
      Real diskDistance;
      if (sc.diskTest(ray, node.pos, diskDistance)) visit(self, diskDistance);
    }
  }
};

}  // namespace gvpm_oracle
