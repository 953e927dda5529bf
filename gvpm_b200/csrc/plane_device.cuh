// plane_device.cuh — device functions of the G-Planes 0D gather ("plane0d").
// Reference lines restated here:
//   PhotonPlane::intersectPlane0D / getContrib0D / invJacobian   photonmapper/plane_struct.h:104-135,150-196
//   PlaneGradRadianceQuery::operator()                           gvpm/shift/shift_volume_planes.h:57-101
//   PlaneGradRadianceQuery::specularShift / intersection         shift_volume_planes.h:263-416,426-453
// Everything that decides whether a plane is intersected (base ray) or whether the shifted plane is intersected
// (offset ray) is strictly rounded fp32 in the reference's operation order.
#pragma once
#include "bre_device.cuh"

namespace gvpm {

// Plane records: 6 float4 planes of n entries each, in Morton order of the plane centres:
//   Q0 = ori.xyz, length0 | Q1 = e0.xyz (= w0*length0), length1 | Q2 = e1.xyz (= w1*length1), edgeID (bits)
//   Q3 = flux.xyz, -      | Q4 = w0.xyz, -                      | Q5 = w1.xyz, -
// The intersection test reads Q0-Q2 only (48 B/plane).

struct PlaneRec {
  v3 ori, e0, e1, flux, w0, w1;
  sf length0, length1;
  int edgeID;
};
struct PlaneIts { sf tCam, t0, t1, invDet; };

// PhotonPlane::intersectPlane0D, plane_struct.h:104-135
__device__ __forceinline__ bool plane_intersect(v3 ori, v3 e0, v3 e1, sf length0, sf length1, v3 o, v3 d, sf mint,
                                                sf maxt, PlaneIts &r) {
  const v3 Pv = cross(d, e1);
  const sf det = dot(e0, Pv);
  if (fabsf(det.v) < 1e-5f) return false;
  r.invDet = sf(1.0f) / det;
  const v3 T = o - ori;
  r.t0 = dot(T, Pv) * r.invDet;
  if (r.t0.v < 0.0f || r.t0.v > 1.0f) return false;
  const v3 Q = cross(T, e0);
  r.t1 = dot(d, Q) * r.invDet;
  if (r.t1.v < 0.0f || r.t1.v > 1.0f) return false;
  r.tCam = dot(e1, Q) * r.invDet;
  if (r.tCam <= mint || r.tCam >= maxt) return false;
  r.t1 = r.t1 * length1;
  r.t0 = r.t0 * length0;
  return true;
}

__device__ __forceinline__ sf abs_dot(v3 a, v3 b) { return sf(fabsf(dot(a, b).v)); }

// 1.0 / absDot(w0, cross(w1, k)): the reference divides in double and rounds to Float, which equals the fp32
// quotient (double rounding is innocuous for a 53-bit intermediate of 24-bit operands)
__device__ __forceinline__ sf plane_inv_jacobian(v3 w0, v3 w1, v3 k) { return sf(1.0f) / abs_dot(w0, cross(w1, k)); }

// PlaneGradRadianceQuery::intersection, shift_volume_planes.h:426-453
__device__ __forceinline__ bool plane_shift_intersection(v3 o, v3 d, sf mint, sf maxt, v3 ori, v3 w0, v3 w1, sf &t0,
                                                         sf &t1) {
  const v3 Pv = cross(d, w1);
  const sf det = dot(w0, Pv);
  if (fabsf(det.v) < 1e-8f) return false;
  const sf invDet = sf(1.0f) / det;
  const v3 T = o - ori;
  t0 = dot(T, Pv) * invDet;
  if (t0.v < 0.0f) return false;
  const v3 Q = cross(T, w0);
  t1 = dot(d, Q) * invDet;
  if (t1.v < 0.0f) return false;
  const sf tCam = dot(w1, Q) * invDet;
  return !(tCam <= mint || tCam >= maxt);
}

// PlaneGradRadianceQuery::operator() for one intersected (ray, plane) pair: base contribution + the four
// specular shifts.  `rec` = the ray's 20-float4 record.  a[27] += contributions.
__device__ __forceinline__ void plane_functor(const GatherParams &P, const float4 *__restrict__ rec, v3 rayD,
                                              const PlaneRec &pl, const PlaneIts &bRec, float *a) {
  const v3 sigS(P.sigma_s[0], P.sigma_s[1], P.sigma_s[2]);
  // getContrib0D.  From here on every quotient is radiometric only (no decision hangs on it; the shifted intersection
  // test below keeps its strictly rounded form): fast reciprocal / division (2 ulp), as in the BRE shading
  const MediumRec mCam = medium_eval(P, sf(0.f), bRec.tCam), mRec0 = medium_eval(P, sf(0.f), bRec.t0),
                  mRec1 = medium_eval(P, sf(0.f), bRec.t1);
  const sf phaseBase = phase_eval(P, -pl.w1, -rayD);
  const sf absBase = abs_dot(pl.w0, cross(pl.w1, rayD));
  const sf invJacBase = frcp(absBase);   // plane_inv_jacobian
  v3 baseContrib = (((sigS * mCam.T) * sigS) * pl.flux) * phaseBase;
  baseContrib = baseContrib * (mRec1.T * mRec0.T);
  baseContrib = baseContrib * frcp(mRec0.pdfFailure);
  baseContrib = baseContrib * frcp(mRec1.pdfFailure);
  baseContrib = baseContrib * invJacBase;
  acc_add(a, 0, baseContrib);
  const sf w0Dot = dot(pl.w0, pl.w1);
  const sf sinW = ssqrt(sf(1.f) - (w0Dot * w0Dot));
  const sf invT0 = frcp(mRec0.T), invT1 = frcp(mRec1.T), invPhaseBase = frcp(phaseBase);
  float Sx[4], Sy[4], Sz[4], Wk[4];  // staged so that a[] keeps static indices (registers)
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const float4 s0 = ldg4(rec + 4 * (k + 1)), s1 = ldg4(rec + 4 * (k + 1) + 1), s2 = ldg4(rec + 4 * (k + 1) + 2);
    sf weight(1.f);
    v3 S(0.f, 0.f, 0.f);
    if (__float_as_uint(s2.w) != 0u) {
      // specularShift
      const v3 so(s0.x, s0.y, s0.z), sd_(s1.x, s1.y, s1.z);
      const sf sMaxt(s0.w), sensor(s1.w);
      const v3 newIntersection = so + sd_ * bRec.tCam;
      v3 orthNewW1 = newIntersection - (pl.ori + pl.w0 * dot(newIntersection - pl.ori, pl.w0));
      orthNewW1 = orthNewW1 / length(orthNewW1);
      const v3 newW1 = sinW * orthNewW1 + pl.w0 * w0Dot;
      sf t0New, t1New;
      if (plane_shift_intersection(so, sd_, sf(P.cfg.epsilon), sMaxt, pl.ori, pl.w0, newW1, t0New, t1New)) {
        const MediumRec mRec1S = medium_eval(P, sf(0.f), t1New), mRec0S = medium_eval(P, sf(0.f), t0New);
        const sf absNew = abs_dot(pl.w0, cross(newW1, sd_));
        v3 thr = baseContrib;
        thr = thr * (mRec0S.T * invT0);
        thr = thr * (mRec1S.T * invT1);
        thr = thr * absBase;            // 1 / invJacBase
        thr = thr * frcp(absNew);
        sf jac = invJacBase;
        jac = jac * absNew;
        jac = jac * (bRec.t1 * frcp(t1New));                       // / (t1New / bRec.t1)
        if (pl.edgeID != 1) jac = jac * (bRec.t0 * frcp(t0New));   // / (t0New / bRec.t0)
        const sf phaseNew = phase_eval(P, -newW1, -sd_);
        thr = thr * phaseNew;
        thr = thr * invPhaseBase;
        weight = sf(0.5f);
        S = thr * jac;
        if (P.cfg.use_mis) {
          const sf basePdf = mRec0.pdfSuccess * mRec1.pdfSuccess * phaseBase;
          const sf offsetPdf = mRec0S.pdfSuccess * mRec1S.pdfSuccess * phaseNew;
          if (offsetPdf.v == 0.f || basePdf.v == 0.f) weight = sf(1.f);
          else weight = frcp(sf(1.f) + fdiv(sensor * jac * offsetPdf, basePdf));
        }
      }
    }
    Sx[k] = S.x.v; Sy[k] = S.y.v; Sz[k] = S.z.v; Wk[k] = weight.v;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const sf wk(Wk[k]);
    acc_add(a, 1 + k, v3(Sx[k], Sy[k], Sz[k]) * wk);
    acc_add(a, 5 + k, baseContrib * wk);
  }
}

}  // namespace gvpm
