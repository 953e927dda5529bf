// ref_physics.cpp — TEST INFRASTRUCTURE.  extern "C" marshalling around the REFERENCE'S OWN radiometric code, compiled
// from /root/reference where it lies (oracle/Makefile, target _ref/libgvpm_physics_ref.so; no reference source is
// copied).  It pins the radiometric half of oracle/gvpm_oracle.hpp (rows a9 and a17 of SURVEY.md §8): every function
// below only converts flat arrays into the reference's records, calls the reference, and flattens what comes back.
//
// What is pinned (reference file:line -> harness entry):
//   HomogeneousMedium::eval, src/medium/homogeneous.cpp:432-513                          -> ref_phys_medium_eval
//   IsotropicPhaseFunction::eval/pdf, src/phase/isotropic.cpp:76                           -> ref_phys_phase
//   HGPhaseFunction::eval/pdf, src/phase/hg.cpp:107-110                                    -> ref_phys_phase
//   SmoothDiffuse::eval/pdf, src/bsdfs/diffuse.cpp:110-127 (+ BSDF::pdfComponent)          -> ref_phys_diffuse_bsdf
//   AreaLight::evalDirection/pdfDirection, src/emitters/area.cpp:132-150                   -> ref_phys_area_emitter
//   diffuseReconnection, gvpm/shift/operation/shift_diffuse.cpp:11-134, driven with PathVertex / PathEdge records of
//     the three in-scope parent types (include/mitsuba/bidir/{vertex,edge}.h)              -> ref_phys_diffuse_reconnection
// The plugin classes have no headers: they are reached through the CreateInstance entry point every plugin exports
// (MTS_EXPORT_PLUGIN, include/mitsuba/core/cobject.h:99-107), renamed per file on the compiler command line.
#include <mitsuba/render/medium.h>
#include <mitsuba/render/phase.h>
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/emitter.h>
#include <mitsuba/render/shape.h>
#include <mitsuba/bidir/vertex.h>
#include <mitsuba/bidir/edge.h>
#include "gvpm/shift/operation/shift_diffuse.h"

#include <cstring>

using namespace mitsuba;

extern "C" {
void *CreateInstance_hom(const Properties &props);
void *CreateInstance_iso(const Properties &props);
void *CreateInstance_hg(const Properties &props);
void *CreateInstance_dif(const Properties &props);
void *CreateInstance_area(const Properties &props);
}

namespace {

inline Point P3(const float *p) { return Point(p[0], p[1], p[2]); }
inline Vector V3f(const float *p) { return Vector(p[0], p[1], p[2]); }
inline Spectrum S3(const float *p) {
  Spectrum s;
  s.fromLinearRGB(p[0], p[1], p[2]);   // RGB build: stores the three values as they are (spectrum.h)
  return s;
}
inline void putS(float *dst, const Spectrum &s) { dst[0] = s[0]; dst[1] = s[1]; dst[2] = s[2]; }

// Class::staticInitialization links every registered class to its superclass (src/libcore/class.cpp); the reference
// does this in its start-up code (mitsuba.cpp), the addChild() type checks depend on it
void initOnce() {
  static bool done = false;
  if (!done) { Class::staticInitialization(); done = true; }
}

ref<PhaseFunction> makePhase(int type, float g) {
  initOnce();
  ref<PhaseFunction> ph;
  if (type == 0) {
    ph = static_cast<PhaseFunction *>(CreateInstance_iso(Properties("isotropic")));
  } else {
    Properties p("hg");
    p.setFloat("g", g);
    ph = static_cast<PhaseFunction *>(CreateInstance_hg(p));
  }
  ph->configure();
  return ph;
}

ref<Medium> makeMedium(const float *sigS, const float *sigA, float samplingWeight, PhaseFunction *phase) {
  Properties p("homogeneous");
  p.setSpectrum("sigmaS", S3(sigS));
  p.setSpectrum("sigmaA", S3(sigA));
  p.setFloat("mediumSamplingWeight", samplingWeight);
  ref<Medium> m = static_cast<Medium *>(CreateInstance_hom(p));
  m->addChild("", phase);
  m->configure();
  return m;
}

// a shape whose only job is to own the BSDF an Intersection refers to (Intersection::getBSDF -> shape->getBSDF)
class HarnessShape : public Shape {
public:
  explicit HarnessShape(BSDF *bsdf) : Shape(Properties("harness")) { m_bsdf = bsdf; }
  AABB getAABB() const { return AABB(); }
  size_t getPrimitiveCount() const { return 1; }
  size_t getEffectivePrimitiveCount() const { return 1; }
};

ref<BSDF> makeDiffuse(const float *albedo) {
  Properties p("diffuse");
  p.setSpectrum("reflectance", S3(albedo));
  ref<BSDF> b = static_cast<BSDF *>(CreateInstance_dif(p));
  b->configure();
  return b;
}

}  // namespace

extern "C" {

int ref_phys_version() { return 1; }

// T: [3n], pdfSuccess / pdfFailure: [n]
void ref_phys_medium_eval(const float *sigS, const float *sigA, float samplingWeight, size_t n, const float *mint,
                          const float *maxt, float *T, float *pdfSuccess, float *pdfFailure) {
  ref<PhaseFunction> ph = makePhase(0, 0.f);
  ref<Medium> m = makeMedium(sigS, sigA, samplingWeight, ph.get());
  for (size_t i = 0; i < n; ++i) {
    MediumSamplingRecord mRec;
    Ray ray(Point(0.f), Vector(0.f, 0.f, 1.f), mint[i], maxt[i], 0.f);
    m->eval(ray, mRec);
    putS(T + 3 * i, mRec.transmittance);
    pdfSuccess[i] = mRec.pdfSuccess;
    pdfFailure[i] = mRec.pdfFailure;
  }
}

// type 0 isotropic, 1 Henyey-Greenstein; wi, wo: [3n] world directions as diffuseReconnection passes them
void ref_phys_phase(int type, float g, size_t n, const float *wi, const float *wo, float *eval, float *pdf) {
  ref<PhaseFunction> ph = makePhase(type, g);
  MediumSamplingRecord mRec;
  for (size_t i = 0; i < n; ++i) {
    PhaseFunctionSamplingRecord pRec(mRec, V3f(wi + 3 * i), V3f(wo + 3 * i), EImportance);
    eval[i] = ph->eval(pRec);
    pdf[i] = ph->pdf(pRec);
  }
}

// diffuse BSDF at a surface with geometric = shading normal `normal`; wi, wo in WORLD space (the harness forms the local
// directions through the reference's Frame, as Intersection::toLocal does)
void ref_phys_diffuse_bsdf(const float *albedo, size_t n, const float *normal, const float *wiWorld, const float *woWorld,
                           float *eval, float *pdf) {
  ref<BSDF> bsdf = makeDiffuse(albedo);
  ref<HarnessShape> shape = new HarnessShape(bsdf.get());
  for (size_t i = 0; i < n; ++i) {
    Intersection its;
    its.p = Point(0.f);
    its.geoFrame = Frame(Normal(V3f(normal + 3 * i)));
    its.shFrame = its.geoFrame;
    its.shape = shape.get();
    its.wi = its.toLocal(V3f(wiWorld + 3 * i));
    BSDFSamplingRecord bRec(its, its.wi, its.toLocal(V3f(woWorld + 3 * i)), EImportance);
    putS(eval + 3 * i, bsdf->eval(bRec, ESolidAngle));
    pdf[i] = bsdf->pdf(bRec, ESolidAngle) * bsdf->pdfComponent(bRec);
  }
}

void ref_phys_area_emitter(size_t n, const float *normal, const float *d, float *eval, float *pdf) {
  Properties p("area");
  float one[3] = {1.f, 1.f, 1.f};
  p.setSpectrum("radiance", S3(one));
  ref<Emitter> em = static_cast<Emitter *>(CreateInstance_area(p));
  for (size_t i = 0; i < n; ++i) {
    PositionSamplingRecord pRec;
    pRec.n = Normal(V3f(normal + 3 * i));
    pRec.measure = EArea;
    DirectionSamplingRecord dRec;
    dRec.d = V3f(d + 3 * i);
    dRec.measure = ESolidAngle;
    putS(eval + 3 * i, em->evalDirection(dRec, pRec));
    pdf[i] = em->pdfDirection(dRec, pRec);
  }
}

// diffuseReconnection(sRec, newIts, newD, newDLength, parentVertex, edge, isVolumeBase = true, predPos, adjointCorr = false)
// for n independent parent vertices.  parent_type: 0 emitter sample, 1 diffuse surface, 2 medium (gvpm_parent_type).
// Outputs: ok[n] (the function's return value), throughput[3n] and pdf[n] of the ShiftRecord it filled.
void ref_phys_diffuse_reconnection(const float *sigS, const float *sigA, float samplingWeight, int phaseType, float g,
                                   size_t n, const uint8_t *parent_type, const float *parent_pos, const float *pred_pos,
                                   const float *parent_n, const float *albedo, const float *parent_pdf,
                                   const float *edge_pdf, const float *rr_weight, const float *newD,
                                   const float *newDLength, uint8_t *ok, float *throughput, float *pdf) {
  ref<PhaseFunction> ph = makePhase(phaseType, g);
  ref<Medium> medium = makeMedium(sigS, sigA, samplingWeight, ph.get());
  Properties ep("area");
  float one[3] = {1.f, 1.f, 1.f};
  ep.setSpectrum("radiance", S3(one));
  ref<Emitter> em = static_cast<Emitter *>(CreateInstance_area(ep));
  for (size_t i = 0; i < n; ++i) {
    PathVertex v;
    std::memset(&v, 0, sizeof(v));
    PathEdge e;
    std::memset(&e, 0, sizeof(e));
    e.medium = medium.get();
    e.pdf[EImportance] = edge_pdf[i];
    v.pdf[EImportance] = parent_pdf[i];
    v.rrWeight = rr_weight[i];
    v.sampledComponentIndex = -1;
    const Point pos = P3(parent_pos + 3 * i), pred = P3(pred_pos + 3 * i);
    const Normal nrm(V3f(parent_n + 3 * i));
    ref<BSDF> bsdf;
    ref<HarnessShape> shape;
    if (parent_type[i] == 1) {
      v.type = PathVertex::ESurfaceInteraction;
      Intersection &its = v.getIntersection();
      new (&its) Intersection();
      its.p = pos;
      its.geoFrame = Frame(nrm);
      its.shFrame = its.geoFrame;
      bsdf = makeDiffuse(albedo + 3 * i);
      shape = new HarnessShape(bsdf.get());
      its.shape = shape.get();
      its.wi = its.toLocal(normalize(pred - pos));   // what the tracer stores: the direction back to the predecessor
      its.t = 1.f;
    } else if (parent_type[i] == 2) {
      v.type = PathVertex::EMediumInteraction;
      MediumSamplingRecord &mRec = v.getMediumSamplingRecord();
      new (&mRec) MediumSamplingRecord();
      mRec.p = pos;
      mRec.medium = medium.get();
      mRec.sigmaS = S3(sigS);
      mRec.sigmaA = S3(sigA);
    } else {
      v.type = PathVertex::EEmitterSample;
      PositionSamplingRecord &pRec = v.getPositionSamplingRecord();
      new (&pRec) PositionSamplingRecord();
      pRec.p = pos;
      pRec.n = nrm;
      pRec.measure = EArea;
      pRec.object = em.get();
    }
    ShiftRecord sRec;
    Intersection unusedIts;
    const bool good = diffuseReconnection(sRec, unusedIts, V3f(newD + 3 * i), newDLength[i], &v, &e, true, pred, false);
    ok[i] = good ? 1 : 0;
    putS(throughput + 3 * i, sRec.throughtput);
    pdf[i] = sRec.pdf;
  }
}

}  // extern "C"
