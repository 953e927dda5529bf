// ref_functor.cpp — TEST INFRASTRUCTURE.  Drives the REFERENCE'S OWN point-photon shift functors
//   VolumeGradientBREQuery::operator()                    gvpm/shift/shift_volume_photon.cpp:658-856   (G-BRE)
//   VolumeGradientPositionQuery::operator()               :489-655                                     (G-VPM)
//   shiftNull / shiftPhoton / shiftPhotonDiffuse           :119-158, :49-117, :382-486
//   getShiftPos                                           :858-896
//   + getTypeShift / VertexClassifier, diffuseReconnection, HomogeneousMedium::eval, the phase functions, the diffuse
//     BSDF, the area emitter
// compiled from /root/reference where it lies (oracle/Makefile, target functor_ref -> _ref/libgvpm_functor_ref.so; no
// reference source is copied) on the flattened inputs of the C ABI (gvpm_photon_soa / gvpm_ray_soa / gvpm_vpm_sample_soa,
// include/gvpm_b200.h), so that the oracle's restatement of the functors' CONTROL FLOW (which offsets take the null
// shift, which the diffuse reconnection, the border rule, the filters, the MIS weights, the accumulation) is pinned to
// reference output, not only its radiometric building blocks (ref_physics.cpp).
//
// What this file does: it rebuilds, for every (camera segment, photon) pair, the reference-side objects the functors
// read - a GatherPoint with its camera Path and cached vertex weights, four ShiftGatherPoints marked as generated, the
// photon's light Path (emitter sample, intermediate medium vertices, predecessor, emitter / diffuse surface / medium
// parent vertex, prefix weights), a GPhotonNodeKD - from the flattened arrays, evaluates the neighbour predicate with the
// statements of GradientBeamRadianceEstimator::query (gvpm_accel.h:293-301) / PointKDTree::executeQuery
// (kdtree.h:721-723), and calls the functor.  All radiometry and every branch is the reference's.  Restricted to first
// medium edges (edge_id = 1: sensorMIS has no geometry terms there, gvpm_struct.h:608-631), which is what primary camera
// rays are.
//
// Objects of the reference that cannot be constructed here (Scene, Sensor, Film, ShapeKDTree, GPMThreadData need the
// whole renderer) are raw zeroed storage with the two or three members the functor reads poked in; the shadow ray of
// the reconnection reaches ShapeKDTree::rayIntersect, which functor_stubs.cpp answers with the reference's own
// Triangle::rayIntersect over the harness' triangle list.
#include <array>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
// every standard header the reference's headers pull in, BEFORE the access override below (libstdc++ does not survive it)
#include <algorithm>
#include <cmath>
#include <deque>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <memory>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <stack>
#include <stdexcept>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <atomic>
#include <mutex>
#include <thread>
#include <random>
#include <chrono>
#include <complex>
#include <bitset>
#include <iterator>
#include <typeinfo>
#include <cassert>
#include <cstdio>
#include <cstdint>

// the functor's configuration, scene and film members are private / protected; the harness pokes them directly
#define private public
#define protected public
#include <mitsuba/render/scene.h>
#include <mitsuba/render/sensor.h>
#include <mitsuba/render/film.h>
#include <mitsuba/bidir/path.h>
#include <mitsuba/render/sampler.h>
#include "gvpm/gvpm_accel.h"
#include "beams_accel.h"
#include "plane_accel.h"
#include "gvpm/shift/shift_volume_photon.h"
#include "gvpm/shift/shift_volume_beams.h"
#include "gvpm/gvpm_plane.h"
#include "gvpm/shift/shift_volume_planes.h"
#include <mitsuba/render/photon.h>
#include "bre.h"
#undef private
#undef protected

#include "ref_physics.cpp"   // makePhase / makeMedium / makeDiffuse / HarnessShape and the ref_phys_* entries

#include "../include/gvpm_b200.h"
#include <mitsuba/render/trimesh.h>
#include "../gvpm_b200/host/gvpm_mitsuba_shim.hpp"   // the reference-side flattening shims, executed by ref_fn_shim_*

namespace mitsuba {
// the triangle list the stand-in of ShapeKDTree::rayIntersect walks (functor_stubs.cpp)
extern std::vector<std::array<Point, 3>> g_functor_occluders;
}

namespace {

template <class T> T *rawZeroed() {
  void *p = std::calloc(1, sizeof(T) + 64);
  return reinterpret_cast<T *>(p);
}

struct LightPath {
  std::vector<PathVertex> v;
  std::vector<PathEdge> e;
  std::vector<ref<BSDF>> bsdfs;
  std::vector<ref<HarnessShape>> shapes;
  Path path;
};

void zero(PathVertex &x) { std::memset(&x, 0, sizeof(x)); x.sampledComponentIndex = -1; x.rrWeight = 1.f; }
void zero(PathEdge &x) { std::memset(&x, 0, sizeof(x)); }


// A surface parent that VertexClassifier calls glossy (roughness 0 <= bounceRoughness): getTypeShift then answers
// EManifoldShift, which the functors refuse with useManifold = false (GVPM_PARENT_OTHER of the flattened form).  Nothing of
// it is ever evaluated.
class GlossyStub : public BSDF {
public:
  GlossyStub() : BSDF(Properties("harness_glossy")) {
    m_components.push_back(EGlossyReflection | EFrontSide);
    BSDF::configure();
  }
  Spectrum sample(BSDFSamplingRecord &, const Point2 &) const override { std::abort(); }
  Spectrum sample(BSDFSamplingRecord &, Float &, const Point2 &) const override { std::abort(); }
  Spectrum eval(const BSDFSamplingRecord &, EMeasure) const override { std::abort(); }
  Float pdf(const BSDFSamplingRecord &, EMeasure) const override { std::abort(); }
  Float getRoughness(const Intersection &, int) const override { return 0.f; }
};

// Everything the functors read that does not depend on the camera segment: medium, emitter, the stand-in scene, the
// configuration and one light Path + kd node per photon.
struct World {
  ref<PhaseFunction> phase;
  ref<Medium> medium;
  ref<Emitter> em;
  Scene *scene = nullptr;
  GPMThreadData *thdata = nullptr;
  GPMConfig config;
  int buildThreads = 1;
  std::vector<PathVertex> vArena;
  std::vector<PathEdge> eArena;
  std::vector<Path> paths;
  PathVertex unitEmitter, unitMedium;
  PathEdge unitEdge;
  std::map<std::array<uint32_t, 3>, std::pair<ref<BSDF>, ref<HarnessShape>>> bsdfCache;
  ref<BSDF> glossy;
  ref<HarnessShape> glossyShape;
  std::vector<GPhotonNodeKD> nodes;

  ~World() {
    for (auto &p : paths) { p.m_vertices.clear(); p.m_edges.clear(); }   // they point into the arenas
  }

  int build(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_medium *med, const gvpm_config *cfg, const float *tri,
            size_t n_tri, EVolumeTechnique technique) {
    common(med, cfg, tri, n_tri, technique);
    return photonPaths(ph, n_ph, med, buildThreads);
  }

  void common(const gvpm_medium *med, const gvpm_config *cfg, const float *tri, size_t n_tri, EVolumeTechnique technique) {
    initOnce();
    phase = makePhase(med->phase_type, med->hg_g);
    medium = makeMedium(med->sigma_s, med->sigma_a, med->sampling_weight, phase.get());
    Properties ep("area");
    float one[3] = {1.f, 1.f, 1.f};
    ep.setSpectrum("radiance", S3(one));
    em = static_cast<Emitter *>(CreateInstance_area(ep));

    g_functor_occluders.clear();
    for (size_t t = 0; t < n_tri; ++t)
      g_functor_occluders.push_back({P3(tri + 9 * t), P3(tri + 9 * t + 3), P3(tri + 9 * t + 6)});

    // Scene -> sensor -> film size, Scene -> kd-tree (any-hit): raw storage, members poked in
    scene = rawZeroed<Scene>();
    Sensor *sensor = rawZeroed<Sensor>();
    Film *film = rawZeroed<Film>();
    film->m_size = Vector2i(cfg->film_w, cfg->film_h);
    sensor->m_film.m_ptr = film;
    scene->m_sensor.m_ptr = sensor;
    scene->m_kdtree.m_ptr = rawZeroed<ShapeKDTree>();
    thdata = rawZeroed<GPMThreadData>();

    std::memset((void *)&config, 0, offsetof(GPMConfig, forceAPA));
    config.maxDepth = cfg->max_depth;
    config.minDepth = cfg->min_depth;
    config.lightingInteractionMode = cfg->lighting_mode;
    config.bsdfInteractionMode = BSDF::EAll;
    config.useManifold = false;
    config.useMIS = cfg->use_mis != 0;
    config.debugShift = EAllShift;
    config.noMediumShift = true;                    // plugin default, gvpm_struct.h:191
    config.volTechnique = technique;
    config.useShiftNull = cfg->use_shift_null != 0;
    config.pathSet = cfg->path_set != 0;
    config.powerHeuristic = cfg->power_heuristic != 0;
    VertexClassifier::roughnessThreshold = 0.05f;   // bounceRoughness default (gvpm_struct.h:236)
  }

  // One light Path + kd node per photon.  Per photon four vertices of its own (supernode carrying the prefix weight,
  // predecessor, parent, the photon's vertex) and the edge that carries it; the vertices in between, which the functor
  // only classifies and multiplies in as unit weights, are shared unit records.  Diffuse BSDFs are shared per albedo.
  int photonPaths(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_medium *med, int threads = 1) {
    for (size_t i = 0; i < n_ph; ++i) {
      const size_t c = (size_t)ph->depth[i] + 1;   // vertexId
      if (c < 2) return -2;
      const int ptype = ph->parent_type[i];
      if ((ptype == 0) != (c == 2)) return -3;     // the emitter sample is vertex 1: parent of the photons with vertexId 2 only
      if (ptype > 3) return -4;
      if (ptype == 3 && !glossy) {                  // GVPM_PARENT_OTHER: a glossy surface parent
        glossy = new GlossyStub();
        glossyShape = new HarnessShape(glossy.get());
      }
      if (ptype == 1) {
        std::array<uint32_t, 3> key;
        std::memcpy(key.data(), ph->parent_albedo + 3 * i, 12);
        if (!bsdfCache.count(key)) {
          ref<BSDF> b = makeDiffuse(ph->parent_albedo + 3 * i);
          bsdfCache[key] = std::make_pair(b, ref<HarnessShape>(new HarnessShape(b.get())));
        }
      }
    }
    vArena.resize(4 * n_ph);
    eArena.resize(n_ph);
    paths.resize(n_ph);
    nodes.resize(n_ph);
    zero(unitEmitter);
    unitEmitter.type = PathVertex::EEmitterSample;
    unitEmitter.weight[EImportance] = Spectrum(1.f);
    {
      PositionSamplingRecord &pr = unitEmitter.getPositionSamplingRecord();
      new (&pr) PositionSamplingRecord();
      pr.measure = EArea;
      pr.object = em.get();
    }
    zero(unitMedium);
    unitMedium.type = PathVertex::EMediumInteraction;
    unitMedium.weight[EImportance] = Spectrum(1.f);
    {
      MediumSamplingRecord &m = unitMedium.getMediumSamplingRecord();
      new (&m) MediumSamplingRecord();
      m.medium = medium.get();
    }
    zero(unitEdge);
    unitEdge.weight[EImportance] = Spectrum(1.f);
    unitEdge.medium = medium.get();

    auto one = [&](size_t i) {
      const size_t c = (size_t)ph->depth[i] + 1;
      const int ptype = ph->parent_type[i];
      PathVertex &v0 = vArena[4 * i], &vp = vArena[4 * i + 1], &v = vArena[4 * i + 2], &pv = vArena[4 * i + 3];
      PathEdge &pe = eArena[i];
      zero(v0); zero(vp); zero(v); zero(pv); zero(pe);
      // prefix: vertex(0).weight * rr * edge(0).weight * prod_{1 <= k < c-1} (...) = prefix_flux (every other factor is 1)
      v0.type = PathVertex::EEmitterSupernode;
      v0.weight[EImportance] = S3(ph->prefix_flux + 3 * i);
      vp.weight[EImportance] = v.weight[EImportance] = pv.weight[EImportance] = Spectrum(1.f);
      const Point pos = P3(ph->pos + 3 * i), parent = P3(ph->parent_pos + 3 * i), pred = P3(ph->pred_pos + 3 * i);
      const Normal nrm(V3f(ph->parent_n + 3 * i));
      // the predecessor (c-2) carries its position; it is the emitter sample when c = 3 (vertex 1 is always the emitter
      // sample: getTypeShift walks back to it and the classifier calls it diffuse, gvpm_struct.h:71)
      if (c == 3) {
        vp.type = PathVertex::EEmitterSample;
        PositionSamplingRecord &pr = vp.getPositionSamplingRecord();
        new (&pr) PositionSamplingRecord();
        pr.p = pred;
        pr.measure = EArea;
        pr.object = em.get();
      } else if (c > 3) {
        vp.type = PathVertex::EMediumInteraction;
        MediumSamplingRecord &m = vp.getMediumSamplingRecord();
        new (&m) MediumSamplingRecord();
        m.p = pred;
        m.medium = medium.get();
      }
      // parent vertex (c-1)
      v.pdf[EImportance] = ph->parent_pdf[i];
      v.rrWeight = ph->rr_weight[i];
      if (ptype == 1) {
        v.type = PathVertex::ESurfaceInteraction;
        Intersection &its = v.getIntersection();
        new (&its) Intersection();
        its.p = parent;
        its.geoFrame = Frame(nrm);
        its.shFrame = its.geoFrame;
        std::array<uint32_t, 3> key;
        std::memcpy(key.data(), ph->parent_albedo + 3 * i, 12);
        its.shape = bsdfCache.find(key)->second.second.get();
        its.wi = its.toLocal(normalize(pred - parent));
        its.t = 1.f;
      } else if (ptype == 3) {
        v.type = PathVertex::ESurfaceInteraction;
        Intersection &its = v.getIntersection();
        new (&its) Intersection();
        its.p = parent;
        its.geoFrame = Frame(nrm.lengthSquared() > 0 ? nrm : Normal(0.f, 0.f, 1.f));
        its.shFrame = its.geoFrame;
        its.shape = glossyShape.get();
        its.wi = its.toLocal(normalize(pred - parent));
        its.t = 1.f;
      } else if (ptype == 2) {
        v.type = PathVertex::EMediumInteraction;
        MediumSamplingRecord &m = v.getMediumSamplingRecord();
        new (&m) MediumSamplingRecord();
        m.p = parent;
        m.medium = medium.get();
        m.sigmaS = S3(med->sigma_s);
        m.sigmaA = S3(med->sigma_a);
      } else {
        v.type = PathVertex::EEmitterSample;
        PositionSamplingRecord &pr = v.getPositionSamplingRecord();
        new (&pr) PositionSamplingRecord();
        pr.p = parent;
        pr.n = nrm;
        pr.measure = EArea;
        pr.object = em.get();
      }
      // the photon's own vertex and the edge that carries it
      pv.type = PathVertex::EMediumInteraction;
      MediumSamplingRecord &pm = pv.getMediumSamplingRecord();
      new (&pm) MediumSamplingRecord();
      pm.p = pos;
      pm.medium = medium.get();
      pm.sigmaS = S3(med->sigma_s);
      pm.sigmaA = S3(med->sigma_a);
      Vector d = pos - parent;
      pe.weight[EImportance] = Spectrum(1.f);
      pe.medium = medium.get();
      pe.length = d.length();
      pe.d = d / pe.length;   // what the tracer stores; the flattened form recomputes wi = normalize(parent - pos) = -d
      pe.pdf[EImportance] = ph->edge_pdf[i];
      Path &path = paths[i];
      for (size_t k = 0; k <= c; ++k) {
        PathVertex *vk = k == 0 ? &v0 : k == c ? &pv : k == c - 1 ? &v : k == c - 2 ? &vp : k == 1 ? &unitEmitter : &unitMedium;
        path.append(vk);
        if (k < c) path.append(k == c - 1 ? &pe : &unitEdge);
      }
      nodes[i].setPosition(pos);
      nodes[i].setData(GPhotonNodeData(&path, (int)c, S3(ph->flux + 3 * i), ph->path_id[i]));
    };
    if (threads <= 1 || n_ph < 4096) {
      for (size_t i = 0; i < n_ph; ++i) one(i);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
          for (size_t i = n_ph * t / threads, e_ = n_ph * (t + 1) / threads; i < e_; ++i) one(i);
        });
      for (auto &t : pool) t.join();
    }
    return 0;
  }
};

// The two sampler->next1D() draws of BeamKernelRecord::eval (shift_volume_beams.h:209-236: the point on the photon beam,
// then the point on the camera segment) are inputs of the flattened form: the caller passes the two numbers per
// (ray, beam) pair that the C ABI derives from its counter-based hash.
class PresetSampler : public Sampler {
public:
  PresetSampler() : Sampler(Properties()) {}
  void preset(Float a, Float b) { v[0] = a; v[1] = b; k = 0; }
  Float next1D() override {
    if (k >= 2) { std::fprintf(stderr, "gvpm functor ref harness: third sampler draw\n"); std::abort(); }
    return v[k++];
  }
  Point2 next2D() override { Float a = next1D(); return Point2(a, next1D()); }
  ref<Sampler> clone() override { return NULL; }
private:
  Float v[2] = {0, 0};
  int k = 0;
};

// One light Path + LTPhotonBeam per flattened beam: the beam is edge i = depth from vertex(i) (origin = parent vertex of
// the reconnection, shift_volume_beams.cpp:430-436) to vertex(i + 1).
struct BeamWorld {
  std::vector<LightPath> lps;
  std::vector<LTPhotonBeam> beams;
  ~BeamWorld() {
    for (auto &L : lps) { L.path.m_vertices.clear(); L.path.m_edges.clear(); }
  }
  int build(World &W, const gvpm_beam_soa *bs, size_t n, const gvpm_medium *med, const gvpm_config *cfg, float radius) {
    lps.resize(n);
    beams.reserve(n);
    for (size_t j = 0; j < n; ++j) {
      LightPath &L = lps[j];
      const size_t i = (size_t)bs->depth[j];
      if (i < 1) return -2;
      const int ptype = bs->parent_type[j];
      if ((ptype == 0) != (i == 1)) return -3;   // vertex 1 is the emitter sample
      L.v.resize(i + 2);
      L.e.resize(i + 1);
      for (auto &x : L.v) zero(x);
      for (auto &x : L.e) zero(x);
      // prefix: vertex(0).weight * rr * edge(0).weight * prod_{1 <= k <= i-1} (...) = prefix_flux (every other factor 1)
      L.v[0].type = PathVertex::EEmitterSupernode;
      L.v[0].weight[EImportance] = S3(bs->prefix_flux + 3 * j);
      for (size_t k = 0; k <= i; ++k) { L.e[k].weight[EImportance] = Spectrum(1.f); L.e[k].medium = W.medium.get(); }
      for (size_t k = 1; k <= i + 1; ++k) L.v[k].weight[EImportance] = Spectrum(1.f);
      const Point origin = P3(bs->origin + 3 * j), end = P3(bs->end + 3 * j), pred = P3(bs->pred_pos + 3 * j);
      const Normal nrm(V3f(bs->parent_n + 3 * j));
      if (i >= 2) {   // vertex 1: the emitter sample; vertices 2 .. i-1: medium interactions at the predecessor's position
        L.v[1].type = PathVertex::EEmitterSample;
        PositionSamplingRecord &pr = L.v[1].getPositionSamplingRecord();
        new (&pr) PositionSamplingRecord();
        pr.p = pred;
        pr.measure = EArea;
        pr.object = W.em.get();
      }
      for (size_t k = 2; k + 1 <= i; ++k) {
        L.v[k].type = PathVertex::EMediumInteraction;
        MediumSamplingRecord &m = L.v[k].getMediumSamplingRecord();
        new (&m) MediumSamplingRecord();
        m.p = pred;
        m.medium = W.medium.get();
      }
      // parent vertex (i) = the beam origin
      PathVertex &v = L.v[i];
      v.pdf[EImportance] = bs->parent_pdf[j];
      v.rrWeight = bs->rr_weight[j];
      if (ptype == 1) {
        v.type = PathVertex::ESurfaceInteraction;
        Intersection &its = v.getIntersection();
        new (&its) Intersection();
        its.p = origin;
        its.geoFrame = Frame(nrm);
        its.shFrame = its.geoFrame;
        L.bsdfs.push_back(makeDiffuse(bs->parent_albedo + 3 * j));
        L.shapes.push_back(new HarnessShape(L.bsdfs.back().get()));
        its.shape = L.shapes.back().get();
        its.wi = its.toLocal(normalize(pred - origin));
        its.t = 1.f;
      } else if (ptype == 2) {
        v.type = PathVertex::EMediumInteraction;
        MediumSamplingRecord &m = v.getMediumSamplingRecord();
        new (&m) MediumSamplingRecord();
        m.p = origin;
        m.medium = W.medium.get();
        m.sigmaS = S3(med->sigma_s);
        m.sigmaA = S3(med->sigma_a);
      } else if (ptype == 0) {
        v.type = PathVertex::EEmitterSample;
        PositionSamplingRecord &pr = v.getPositionSamplingRecord();
        new (&pr) PositionSamplingRecord();
        pr.p = origin;
        pr.n = nrm;
        pr.measure = EArea;
        pr.object = W.em.get();
      } else if (ptype == 3) {   // GVPM_PARENT_OTHER: a glossy surface parent (the functor refuses the manifold shift)
        if (!W.glossy) {
          W.glossy = new GlossyStub();
          W.glossyShape = new HarnessShape(W.glossy.get());
        }
        v.type = PathVertex::ESurfaceInteraction;
        Intersection &its = v.getIntersection();
        new (&its) Intersection();
        its.p = origin;
        its.geoFrame = Frame(nrm.lengthSquared() > 0 ? nrm : Normal(0.f, 0.f, 1.f));
        its.shFrame = its.geoFrame;
        its.shape = W.glossyShape.get();
        its.wi = its.toLocal(normalize(pred - origin));
        its.t = 1.f;
      } else {
        return -4;
      }
      // end vertex (i + 1): on a surface (its geometric normal enters the base pdf, :506-507) or in the medium
      PathVertex &ev = L.v[i + 1];
      if (bs->end_on_surface[j]) {
        ev.type = PathVertex::ESurfaceInteraction;
        Intersection &its = ev.getIntersection();
        new (&its) Intersection();
        its.p = end;
        its.geoFrame = Frame(Normal(V3f(bs->end_n + 3 * j)));
        its.shFrame = its.geoFrame;
        its.t = 1.f;
      } else {
        ev.type = PathVertex::EMediumInteraction;
        MediumSamplingRecord &m = ev.getMediumSamplingRecord();
        new (&m) MediumSamplingRecord();
        m.p = end;
        m.medium = W.medium.get();
      }
      for (size_t k = 0; k <= i + 1; ++k) {
        L.path.append(&L.v[k]);
        if (k <= i) L.path.append(&L.e[k]);
      }
      beams.emplace_back(&L.path, i, radius, (int)bs->path_id[j]);
      LTPhotonBeam &b = beams.back();
      b.flux = S3(bs->flux + 3 * j);            // the constructor's product is an input of the flattened form
      b.longBeams = cfg->long_beams != 0;
      // the edge that carries the beam: direction and length as PhotonBeam::setEndPoint derives them
      L.e[i].d = b.getDir();
      L.e[i].length = b.getLength();
    }
    return 0;
  }
};

// One camera medium segment with its four offset segments as the functors see them: a GatherPoint with its camera Path
// and cached vertex weights, four ShiftGatherPoints marked as generated.  Holds pointers into itself: built in place.
struct CameraSide {
  std::vector<PathVertex> bv, sv[4];
  std::vector<PathEdge> be, se[4];
  GatherPoint gp;
  std::vector<ShiftGatherPoint> shiftGPs;
  CameraSide() : shiftGPs(4) {}
  CameraSide(const CameraSide &) = delete;
  ~CameraSide() {
    gp.path.m_vertices.clear(); gp.path.m_edges.clear();
    for (auto &s : shiftGPs) { s.path.m_vertices.clear(); s.path.m_edges.clear(); }
  }
  // vertices 0 .. e+1 and edges 0 .. e of one camera path whose medium segment is edge e: supernode, sensor sample
  // (pixel position), for e > 1 the vertices up to the segment start (medium-type records at the segment origin; only
  // the positions of vertex e and e+1 are read, by GOp in sensorMIS), the segment, its end vertex
  static void layout(std::vector<PathVertex> &v, std::vector<PathEdge> &ed, Path &path, size_t e, const Point &o,
                     const Vector &dir, Float len, const Medium *mediumPtr) {
    v.resize(e + 2);
    ed.resize(e + 1);
    for (auto &x : v) zero(x);
    for (auto &x : ed) zero(x);
    v[0].type = PathVertex::ESensorSupernode;
    v[1].type = PathVertex::ESensorSample;
    PositionSamplingRecord &pr = v[1].getPositionSamplingRecord();
    new (&pr) PositionSamplingRecord();
    pr.p = o;
    for (size_t k = 2; k <= e + 1; ++k) {
      v[k].type = PathVertex::EMediumInteraction;
      MediumSamplingRecord &m = v[k].getMediumSamplingRecord();
      new (&m) MediumSamplingRecord();
      m.p = k <= e ? o : o + dir * len;
      m.medium = mediumPtr;
    }
    for (size_t k = 0; k <= e + 1; ++k) {
      path.append(&v[k]);
      if (k <= e) path.append(&ed[k]);
    }
  }
  static void infos(std::vector<SVertexPDF> &info, size_t e, const Spectrum &weightBeam, Float sensorPdf) {
    info.resize(e + 1);
    for (auto &x : info) {
      x.weight = Spectrum(1.f);
      x.vertexWeight = Spectrum(1.f);
      x.pdf = x.jacobian = 1.f;
    }
    info[e - 1].weight = weightBeam;   // getWeightBeam(e - 1); getWeightVertex(e) = 1
    info[e].pdf = sensorPdf;           // sensorMIS(e, base, ., .) = (pdf / base pdf) * jacobian [* terms that cancel, e > 1]
  }
  // may be called again on the same object (bench: one CameraSide per worker thread, no allocation per ray)
  void build(const gvpm_ray_soa *ry, size_t r, const Medium *mediumPtr) {
    gp.path.m_vertices.clear(); gp.path.m_edges.clear();
    for (auto &s : shiftGPs) { s.path.m_vertices.clear(); s.path.m_edges.clear(); }
    const size_t e = (size_t)ry->edge_id[r];
    const Point o = P3(ry->o + 3 * r);
    const Vector d = V3f(ry->d + 3 * r);
    layout(bv, be, gp.path, e, o, d, ry->edge_len[r], mediumPtr);
    bv[1].getPositionSamplingRecord().uv = Point2((Float)ry->px[r] + 0.5f, (Float)ry->py[r] + 0.5f);
    be[e].d = d;
    be[e].length = ry->edge_len[r];
    be[e].medium = mediumPtr;
    infos(gp.info, e, S3(ry->eye_contrib + 3 * r), 1.f);
    for (int k = 0; k < 4; ++k) {
      const size_t q = 4 * r + k;
      ShiftGatherPoint &s = shiftGPs[k];
      s.generated = true;                                // nothing to trace: generate() returns at once
      const Vector od = V3f(ry->off_d + 3 * q);
      layout(sv[k], se[k], s.path, e, P3(ry->off_o + 3 * q), od, ry->off_len[q], mediumPtr);
      se[k][e].d = -od;                                  // the functor takes shiftDir = -edge(e).d
      se[k][e].length = ry->off_len[q];
      se[k][e].medium = ry->off_valid[q] ? mediumPtr : NULL;   // validVolumeEdge
      infos(s.info, e, S3(ry->off_eye + 3 * q), ry->off_sensor[q]);
    }
  }
};

}  // namespace

extern "C" {

int ref_fn_version() { return 12; }

// G-BRE.  out: [n_rays * 27] = mediumFlux, shiftedMediumFlux[4], weightedMediumFlux[4] summed over the photons of the
// neighbour set in photon order; counts: [n_rays] functor calls (geometric neighbours).  Returns < 0 on unsupported input.
int ref_fn_bre_gather(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_ray_soa *ry, size_t n_rays, const gvpm_medium *med,
                      const gvpm_config *cfg, const float *tri, size_t n_tri, float radius, float *out, uint32_t *counts) {
  World W;
  if (int rc = W.build(ph, n_ph, med, cfg, tri, n_tri, cfg->kernel_3d ? EVolBRE3D : EVolBRE2D)) return rc;
  for (size_t r = 0; r < n_rays; ++r) {
    for (int j = 0; j < 27; ++j) out[27 * r + j] = 0.f;
    if (counts) counts[r] = 0;
    const size_t e = (size_t)ry->edge_id[r];
    if (e < 1 || e > 8) return -5;
    CameraSide cam;
    cam.build(ry, r, W.medium.get());
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    VolumeGradientBREQuery gRec(W.scene, &cam.gp, W.config, *W.thdata, cam.shiftGPs, e, NULL);
    gRec.newRayBase(ray, W.medium.get());
    gRec.clear();
    for (size_t i = 0; i < n_ph; ++i) {
      // GradientBeamRadianceEstimator::query, gvpm_accel.h:293-305
      Vector originToCenter = W.nodes[i].getPosition() - ray.o;
      Float diskDistance = dot(originToCenter, ray.d), radSqr = radius * radius;
      Float distSqr = (ray(diskDistance) - W.nodes[i].getPosition()).lengthSquared();
      if (diskDistance > ray.mint && distSqr < radSqr) {
        Ray baseRay(ray);
        baseRay.maxt = diskDistance;
        gRec.newRayBase(baseRay, W.medium.get());
        gRec(W.nodes[i], radius, ry->xi[r]);
        if (counts) ++counts[r];
      }
    }
    float *o = out + 27 * r;
    putS(o, gRec.mediumFlux);
    for (int k = 0; k < 4; ++k) {
      putS(o + 3 * (1 + k), gRec.shiftedMediumFlux[k]);
      putS(o + 3 * (5 + k), gRec.weightedMediumFlux[k]);
    }
  }
  return 0;
}

// G-VPM: the per-sample part of computeVolumeGradientPhoton, gvpm.cpp:1141-1185, on the host-drawn distance samples of
// gvpm_vpm_sample_soa.  Per sample: changeEdge / newRayBase / clear, the functor VolumeGradientPositionQuery::operator()
// (shift_volume_photon.cpp:489-655) on every photon of PointKDTree's range predicate (kdtree.h:721-723) in photon order,
// then gp.* += gRec.* * normalization (:1177-1182).  out: [n_rays * 27]; mvol: [n_rays] photons found (MVol, :1175).
int ref_fn_vpm_gather(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_ray_soa *ry, size_t n_rays,
                      const gvpm_vpm_sample_soa *smp, size_t n_smp, const gvpm_medium *med, const gvpm_config *cfg,
                      const float *tri, size_t n_tri, int nb_camera_samples, float *out, float *mvol) {
  World W;
  if (int rc = W.build(ph, n_ph, med, cfg, tri, n_tri, EDistance)) return rc;
  struct Pixel { Spectrum mediumFlux, shifted[4], weighted[4]; };
  std::vector<Pixel> pix(n_rays);
  for (auto &p : pix) {
    p.mediumFlux = Spectrum(0.f);
    for (int k = 0; k < 4; ++k) p.shifted[k] = p.weighted[k] = Spectrum(0.f);
  }
  for (size_t r = 0; r < n_rays; ++r) mvol[r] = 0.f;
  const Float normalization = 1.f / nb_camera_samples;
  for (size_t s = 0; s < n_smp; ++s) {
    const size_t r = smp->ray[s];
    if (r >= n_rays) return -6;
    const size_t e = (size_t)ry->edge_id[r];
    if (e < 1 || e > 8) return -5;
    CameraSide cam;
    cam.build(ry, r, W.medium.get());
    VolumeGradientDistanceQuery gRec(W.scene, &cam.gp, W.config, *W.thdata, cam.shiftGPs, 0, NULL);
    gRec.changeEdge(e, smp->pdf_sel[s]);
    // Ray ray(oBeam, dBeam, Epsilon, beamDist, 0.f); sampleDistance fills mRec; ray.maxt = mRec.t
    Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->edge_len[r], 0.f);
    MediumSamplingRecord mRec;
    mRec.t = smp->t[s];
    mRec.p = ray(mRec.t);
    mRec.medium = W.medium.get();
    mRec.sigmaS = S3(med->sigma_s);
    mRec.sigmaA = S3(med->sigma_a);
    mRec.transmittance = S3(smp->transmittance + 3 * s);
    mRec.pdfSuccess = smp->pdf_success[s];
    mRec.pdfSuccessRev = mRec.pdfSuccess;
    mRec.pdfFailure = 0.f;
    mRec.time = 0.f;
    ray.maxt = mRec.t;
    const Float querySize = smp->radius[s];
    gRec.newRayBase(ray, mRec, querySize, mRec.pdfSuccess);
    gRec.clear();
    const Point q = ray.o + mRec.t * ray.d;
    size_t found = 0;
    for (size_t i = 0; i < n_ph; ++i) {
      // PointKDTree::executeQuery, include/mitsuba/core/kdtree.h:721-723
      const Float pointDistSquared = (W.nodes[i].getPosition() - q).lengthSquared();
      if (pointDistSquared < querySize * querySize) {
        gRec(W.nodes[i]);
        ++found;
      }
    }
    mvol[r] += (float)found;
    Pixel &g = pix[r];
    g.mediumFlux += (gRec.mediumFlux * normalization);
    for (int k = 0; k < 4; ++k) {
      g.shifted[k] += (gRec.shiftedMediumFlux[k] * normalization);
      g.weighted[k] += (gRec.weightedMediumFlux[k] * normalization);
    }
  }
  for (size_t r = 0; r < n_rays; ++r) {
    float *o = out + 27 * r;
    putS(o, pix[r].mediumFlux);
    for (int k = 0; k < 4; ++k) {
      putS(o + 3 * (1 + k), pix[r].shifted[k]);
      putS(o + 3 * (5 + k), pix[r].weighted[k]);
    }
  }
  return 0;
}

// G-Beams (beam3d = EBeamBeam3D_Optimized, or beam1d with newShiftBeam as gvpm.cpp:95-98 forces it): the functor
// BeamGradRadianceQuery::operator() (shift_volume_beams.cpp:139-353) with shiftBeam / shiftBeamDiffuse / shiftNull3D /
// getShiftPos / getShiftPos1D, BeamKernelRecord (shift_volume_beams.h:24-288) and diffuseReconnectionPhotonBeam, on every
// (camera segment, beam) pair in beam order with tmin = 0, tmax = infinity (what SubBeamBVH::query selects up to the
// sub-beam that owns tNear).  xi: [n_rays * n_beams * 2] the two sampler draws per pair.  counts: [n_rays * 2] =
// functor calls that returned true, and 0.
int ref_fn_beams_gather(const gvpm_beam_soa *bs, size_t n_beams, const gvpm_ray_soa *ry, size_t n_rays, const gvpm_medium *med,
                        const gvpm_config *cfg, const float *tri, size_t n_tri, float radius, const float *xi, float *out,
                        uint32_t *counts) {
  World W;
  W.common(med, cfg, tri, n_tri, cfg->beam_kernel_1d ? EBeamBeam1D : EBeamBeam3D_Optimized);
  W.config.newShiftBeam = cfg->beam_kernel_1d != 0;
  BeamWorld B;
  if (int rc = B.build(W, bs, n_beams, med, cfg, radius)) return rc;
  ref<PresetSampler> sampler = new PresetSampler();
  for (size_t r = 0; r < n_rays; ++r) {
    for (int j = 0; j < 27; ++j) out[27 * r + j] = 0.f;
    if (counts) counts[2 * r] = counts[2 * r + 1] = 0;
    const size_t e = (size_t)ry->edge_id[r];
    if (e < 1 || e > 8) return -5;
    CameraSide cam;
    cam.build(ry, r, W.medium.get());
    // gvpm.cpp:936: Ray ray(vertex(idEdge).position, d, Epsilon, distTotal - Epsilon, 0.f)
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    BeamGradRadianceQuery gRec(W.scene, &cam.gp, cam.shiftGPs, ray, W.medium.get(), W.config, *W.thdata, e, sampler.get());
    for (size_t j = 0; j < n_beams; ++j) {
      sampler->preset(xi[2 * (r * n_beams + j)], xi[2 * (r * n_beams + j) + 1]);
      if (gRec(&B.beams[j]) && counts) ++counts[2 * r];
    }
    float *o = out + 27 * r;
    putS(o, gRec.mediumFlux);
    for (int k = 0; k < 4; ++k) {
      putS(o + 3 * (1 + k), gRec.shiftedMediumFlux[k]);
      putS(o + 3 * (5 + k), gRec.weightedMediumFlux[k]);
    }
  }
  return 0;
}

// G-Planes 0D: the functor PlaneGradRadianceQuery::operator() (shift_volume_planes.h:56-101) with specularShift
// (:263-416) and its re-intersection (:427-453), PhotonPlane::intersectPlane0D / getContrib0D / invJacobian
// (photonmapper/plane_struct.h), on every (camera segment, plane) pair in plane order.  The functor reads no light-path
// data (its shift keeps origin and w0), so the LTPhotonPlane is filled from the flattened record directly.
// counts: [n_rays * 2] = planes the base ray intersects (mediumFlux += ...), and 0.
int ref_fn_planes_gather(const gvpm_plane_soa *ps, size_t n_planes, const gvpm_ray_soa *ry, size_t n_rays,
                         const gvpm_medium *med, const gvpm_config *cfg, float *out, uint32_t *counts) {
  World W;
  W.common(med, cfg, NULL, 0, EVolPlane0D);
  std::vector<LTPhotonPlane> planes(n_planes);
  for (size_t j = 0; j < n_planes; ++j) {
    LTPhotonPlane &p = planes[j];
    p._ori = P3(ps->origin + 3 * j);
    p._w0 = V3f(ps->w0 + 3 * j);
    p._length0 = ps->length0[j];
    p._w1 = V3f(ps->w1 + 3 * j);
    p._length1 = ps->length1[j];
    p.medium = W.medium.get();
    p._flux = S3(ps->flux + 3 * j);
    p.depth = p.edgeID = ps->edge_id[j];
    p.path = NULL;
    p.pathID = 0;
  }
  for (size_t r = 0; r < n_rays; ++r) {
    for (int j = 0; j < 27; ++j) out[27 * r + j] = 0.f;
    if (counts) counts[2 * r] = counts[2 * r + 1] = 0;
    const size_t e = (size_t)ry->edge_id[r];
    if (e < 1 || e > 8) return -5;
    CameraSide cam;
    cam.build(ry, r, W.medium.get());
    // gvpm.cpp:837: Ray ray(vertex(idEdge).position, d, Epsilon, distTotal - Epsilon, 0.f)
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    PlaneGradRadianceQuery gRec(W.scene, &cam.gp, cam.shiftGPs, ray, W.medium.get(), W.config, *W.thdata, (int)e);
    for (size_t j = 0; j < n_planes; ++j) {
      PhotonPlane::IntersectionRecord probe;
      if (counts && planes[j].intersectPlane0D(ray, probe)) ++counts[2 * r];
      gRec(&planes[j]);
    }
    float *o = out + 27 * r;
    putS(o, gRec.mediumFlux);
    for (int k = 0; k < 4; ++k) {
      putS(o + 3 * (1 + k), gRec.shiftedMediumFlux[k]);
      putS(o + 3 * (5 + k), gRec.weightedMediumFlux[k]);
    }
  }
  return 0;
}

// sppm primal photon beams: BeamRadianceQuery<PhotonBeam>::operator() (photonmapper/beams.h:29-223), all four techniques of
// EVolumeTechnique (1D, 3D naive, 3D EGSR, 3D optimized), driven as volumePhotonBeamPass does (sppm.cpp:846-857) with one
// call per (camera beam, sub-beam [t1, t2]) in the order of the caller's sub-beam table (the SubBeamBVH constructor's
// split, beams_accel.h:98-124).  technique: 0 = 1D, 1 = naive, 2 = EGSR, 3 = optimized.  xi: [n_rays * n_sub * 2] sampler
// draws per pair.  out: [n_rays * 3] = bRadQuery.Li (the caller multiplies by beam.weight, sppm.cpp:857);
// counts: [n_rays * 2] = calls that returned true, and 0.
int ref_fn_sppm_beams_gather(const gvpm_beam_soa *bs, size_t n_beams, const uint32_t *sub_beam, const float *sub_t12,
                             size_t n_sub, const gvpm_ray_soa *ry, size_t n_rays, const gvpm_medium *med,
                             const gvpm_config *cfg, float radius, int technique, const float *xi, float *out,
                             uint32_t *counts) {
  static const EVolumeTechnique kTech[4] = {EBeamBeam1D, EBeamBeam3D_Naive, EBeamBeam3D_EGSR, EBeamBeam3D_Optimized};
  if (technique < 0 || technique > 3) return -1;
  World W;
  W.common(med, cfg, NULL, 0, kTech[technique]);
  std::vector<PhotonBeam> beams;
  beams.reserve(n_beams);
  for (size_t j = 0; j < n_beams; ++j) {
    beams.emplace_back(P3(bs->origin + 3 * j), W.medium.get(), S3(bs->flux + 3 * j), (int)bs->depth[j], radius);
    beams.back().setEndPoint(P3(bs->end + 3 * j));
    beams.back().longBeams = cfg->long_beams != 0;
  }
  ref<PresetSampler> sampler = new PresetSampler();
  for (size_t r = 0; r < n_rays; ++r) {
    if (counts) counts[2 * r] = counts[2 * r + 1] = 0;
    const int depth = ry->edge_id[r];   // the camera beam's depth
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    BeamRadianceQuery<PhotonBeam> q(ray, W.medium.get(), cfg->max_depth == -1 ? -1 : cfg->max_depth - depth,
                                    std::max(0, cfg->min_depth - depth), sampler.get(), kTech[technique]);
    for (size_t s = 0; s < n_sub; ++s) {
      if (sub_beam[s] >= n_beams) return -6;
      sampler->preset(xi[2 * (r * n_sub + s)], xi[2 * (r * n_sub + s) + 1]);
      if (q(&beams[sub_beam[s]], sub_t12[2 * s], sub_t12[2 * s + 1]) && counts) ++counts[2 * r];
    }
    putS(out + 3 * r, q.Li);
  }
  return 0;
}

// sppm primal BRE: the loop body of BeamRadianceEstimator::query (photonmapper/bre.cpp:167-259: depth filter, neighbour
// predicate on the ray re-based at r(r.mint), 3-D kernel with its early-out and per-photon random chord position, 2-D
// kernel with its segment bound, transmittance, phase function, kernel weight), driven as volumePhotonPassBRE does
// (sppm.cpp:968-973).  The estimator is raw storage holding ONE leaf node with an all-enclosing box, re-filled per photon:
// query() then runs its loop body on exactly that photon, and the per-ray sum is taken over the photons in index order
// (the reference sums in traversal order).  The stock Photon stores its power in RGBE (photon.h:40-44): flux must be
// RGBE-representable (ref_fn_rgbe_roundtrip); dir = photon.getDirection() is stored unquantised in this fork
// (MTS_DISCRETIZED_PHOTON 0).  xi: [n_rays * n_ph] the sampler->next1D() of :217 per pair.  out: [n_rays * 3] WITHOUT
// beam.weight and m_scaleFactor; counts: [n_rays * 2] = 0.
int ref_fn_sppm_bre_gather(const float *pos, const float *dir, const float *flux, const uint8_t *depth, size_t n_ph,
                           const gvpm_ray_soa *ry, size_t n_rays, const gvpm_medium *med, const gvpm_config *cfg,
                           float radius, const float *xi, float *out) {
  World W;
  W.common(med, cfg, NULL, 0, cfg->kernel_3d ? EVolBRE3D : EVolBRE2D);
  BeamRadianceEstimator *bre = rawZeroed<BeamRadianceEstimator>();
  BeamRadianceEstimator::BRENode *node =
      reinterpret_cast<BeamRadianceEstimator::BRENode *>(std::calloc(1, sizeof(BeamRadianceEstimator::BRENode)));
  bre->m_nodes = node;
  bre->m_scaleFactor = 1.f;
  bre->m_photonCount = 1;
  bre->m_depth = 1;
  node->aabb = AABB(Point(-1e30f, -1e30f, -1e30f), Point(1e30f, 1e30f, 1e30f));
  node->radius = radius;
  ref<PresetSampler> sampler = new PresetSampler();
  for (size_t r = 0; r < n_rays; ++r) {
    const int beamDepth = ry->edge_id[r];
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    Spectrum sum(0.f);
    for (size_t i = 0; i < n_ph; ++i) {
      Photon &p = node->photon;
      p.setPosition(P3(pos + 3 * i));
      p.setLeaf(true);
      S3(flux + 3 * i).toRGBE(p.data.power);
      p.data.wi = V3f(dir + 3 * i);
      p.data.depth = depth[i];
      sampler->preset(xi[r * n_ph + i], 0.f);
      sum += bre->query(ray, W.medium.get(), cfg->max_depth == -1 ? -1 : cfg->max_depth - beamDepth, cfg->kernel_3d != 0,
                        sampler.get());
    }
    putS(out + 3 * r, sum);
  }
  std::free(node);
  return 0;
}

// Spectrum::toRGBE / fromRGBE (what Photon does to its power, photon.h:40-44,124-132)
void ref_fn_rgbe_roundtrip(const float *in, size_t n, float *out) {
  for (size_t i = 0; i < n; ++i) {
    uint8_t rgbe[4];
    S3(in + 3 * i).toRGBE(rgbe);
    Spectrum s;
    s.fromRGBE(rgbe);
    putS(out + 3 * i, s);
  }
}

// sppm primal photon planes: PhotonPlaneQuery::operator() (photonmapper/plane_struct.h:238-256) on every (camera beam,
// plane) pair in plane order.  out: [n_rays * 3] = Li (the caller multiplies by beam.weight); counts: [n_rays * 2] =
// planes hit, and 0.
int ref_fn_sppm_planes_gather(const gvpm_plane_soa *ps, size_t n_planes, const gvpm_ray_soa *ry, size_t n_rays,
                              const gvpm_medium *med, const gvpm_config *cfg, float *out, uint32_t *counts) {
  World W;
  W.common(med, cfg, NULL, 0, EVolPlane0D);
  std::vector<LTPhotonPlane> planes(n_planes);
  for (size_t j = 0; j < n_planes; ++j) {
    LTPhotonPlane &p = planes[j];
    p._ori = P3(ps->origin + 3 * j);
    p._w0 = V3f(ps->w0 + 3 * j);
    p._length0 = ps->length0[j];
    p._w1 = V3f(ps->w1 + 3 * j);
    p._length1 = ps->length1[j];
    p.medium = W.medium.get();
    p._flux = S3(ps->flux + 3 * j);
    p.depth = p.edgeID = ps->edge_id[j];
    p.path = NULL;
    p.pathID = 0;
  }
  for (size_t r = 0; r < n_rays; ++r) {
    if (counts) counts[2 * r] = counts[2 * r + 1] = 0;
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    PhotonPlaneQuery q(ray, W.medium.get(), -1, 0, NULL, EVolPlane0D);
    for (size_t j = 0; j < n_planes; ++j)
      if (q(&planes[j]) && counts) ++counts[2 * r];
    putS(out + 3 * r, q.Li);
  }
  return 0;
}

// G-BRE, the whole gather pass on the reference's own code: GPhotonMap::build (PointKDTree, sliding midpoint) +
// GradientBeamRadianceEstimator (hierarchy) + bre->query (traversal, neighbour predicate) + VolumeGradientBREQuery (functor)
// per camera segment, i.e. computeVolumeGradientPhotonBRE's inner loop (gvpm.cpp:994-1042) without its per-pass
// normalisation.  Photons enter the map through the protected kd-tree (tryAppend walks whole light paths).  Rays are
// split over `threads` std::threads (the reference: BlockScheduler over image blocks).  Sums are in traversal order.
// times_ms: [3] = kd build, hierarchy, gather.
namespace {
class FunctorMap : public GPhotonMap {
public:
  explicit FunctorMap(size_t n) : GPhotonMap(n, false, Point(0.f), 0.f) {}
  void add(const GPhotonNodeKD &node) { m_kdtree.push_back(node); }
};
double nowMs() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

// Persistent form (bench.py's reference arm: one build, many sampled gathers).
struct BrePass {
  World W;
  ref<FunctorMap> map;
  ref<GradientBeamRadianceEstimator> bre;
  float radius = 0;
};

// times_ms: [3] = light-path records (harness work, not the reference's), kd build, hierarchy
void *ref_fn_bre_open(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_medium *med, const gvpm_config *cfg, const float *tri,
                      size_t n_tri, float radius, int threads, double *times_ms) {
  BrePass *P = new BrePass();
  P->radius = radius;
  P->W.buildThreads = threads;
  double t0 = nowMs();
  if (P->W.build(ph, n_ph, med, cfg, tri, n_tri, cfg->kernel_3d ? EVolBRE3D : EVolBRE2D)) { delete P; return NULL; }
  double t1 = nowMs();
  P->map = new FunctorMap(n_ph);
  for (size_t i = 0; i < n_ph; ++i) P->map->add(P->W.nodes[i]);
  std::vector<GPhotonNodeKD>().swap(P->W.nodes);                                                   // the map holds them now
  P->map->build(true);                                                                             // gvpm.cpp:453
  double t2 = nowMs();
  P->bre = new GradientBeamRadianceEstimator(P->map.get(), radius);                                // gvpm.cpp:994
  double t3 = nowMs();
  if (times_ms) { times_ms[0] = t1 - t0; times_ms[1] = t2 - t1; times_ms[2] = t3 - t2; }
  return P;
}

void ref_fn_bre_close(void *h) { delete (BrePass *)h; }

int ref_fn_bre_run(void *h, const gvpm_ray_soa *ry, size_t ray_begin, size_t ray_end, int threads, float *out,
                   uint32_t *counts, double *gather_ms) {
  BrePass *P = (BrePass *)h;
  World &W = P->W;
  for (size_t r = ray_begin; r < ray_end; ++r)
    if (ry->edge_id[r] < 1 || ry->edge_id[r] > 8) return -5;
  const double t2 = nowMs();
  std::atomic<size_t> next(ray_begin);
  auto worker = [&]() {
    CameraSide cam;
    for (;;) {
      const size_t b = next.fetch_add(256);
      if (b >= ray_end) break;
      const size_t e_ = std::min(ray_end, b + 256);
      for (size_t r = b; r < e_; ++r) {
        struct Counting : VolumeGradientBREQuery {
          using VolumeGradientBREQuery::VolumeGradientBREQuery;
          uint32_t calls = 0;
          void operator()(const GPhotonNodeKD &n, Float rad, Float xi) { ++calls; VolumeGradientBREQuery::operator()(n, rad, xi); }
        };
        cam.build(ry, r, W.medium.get());
        const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
        Counting gRec(W.scene, &cam.gp, W.config, *W.thdata, cam.shiftGPs, (size_t)ry->edge_id[r], NULL);
        gRec.newRayBase(ray, W.medium.get());
        gRec.clear();
        P->bre->query(ray, W.medium.get(), gRec, ry->xi[r]);                                       // gvpm.cpp:1042
        if (out) {
          float *o = out + 27 * (r - ray_begin);
          putS(o, gRec.mediumFlux);
          for (int k = 0; k < 4; ++k) {
            putS(o + 3 * (1 + k), gRec.shiftedMediumFlux[k]);
            putS(o + 3 * (5 + k), gRec.weightedMediumFlux[k]);
          }
        }
        if (counts) counts[r - ray_begin] = gRec.calls;
      }
    }
  };
  if (threads <= 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
  }
  if (gather_ms) *gather_ms = nowMs() - t2;
  return 0;
}

int ref_fn_bre_pass(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_ray_soa *ry, size_t ray_begin, size_t ray_end,
                    const gvpm_medium *med, const gvpm_config *cfg, const float *tri, size_t n_tri, float radius, int threads,
                    float *out, uint32_t *counts, double *times_ms) {
  double tm[3] = {0, 0, 0}, g = 0;
  void *h = ref_fn_bre_open(ph, n_ph, med, cfg, tri, n_tri, radius, 1, tm);
  if (!h) return -2;
  const int rc = ref_fn_bre_run(h, ry, ray_begin, ray_end, threads, out, counts, &g);
  ref_fn_bre_close(h);
  if (times_ms) { times_ms[0] = tm[1]; times_ms[1] = tm[2]; times_ms[2] = g; }
  return rc;
}

// ---- whole passes of the other techniques on the reference's own acceleration structures --------------------------------------
// Each splits the rays over `threads` std::threads (the reference: BlockScheduler over image blocks) and sums in the
// structure's traversal order.  times_ms: [2] = structure build, gather.
}  // extern "C"

namespace {
// The C ABI's counter-based uniform numbers for the beam kernel record (oracle: Scene::beamUniform; CUDA: beam_device.cuh):
// an input convention of the flattened form, restated here so that a gather needs no [rays x beams] table.
uint32_t abiHash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
float abiBeamUniform(uint32_t seed, int px, int py, int edge, uint32_t beam, uint32_t dim) {
  uint32_t h = abiHash32(seed ^ 0x9E3779B9u);
  h = abiHash32(h ^ (uint32_t)px);
  h = abiHash32(h ^ ((uint32_t)py * 0x85EBCA6Bu));
  h = abiHash32(h ^ ((uint32_t)edge * 0xC2B2AE35u));
  h = abiHash32(h ^ beam);
  h = abiHash32(h ^ (dim * 0x27D4EB2Fu));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}
template <class F> void parallelRays(size_t n, int threads, F f) {
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    CameraSide cam;
    for (;;) {
      const size_t b = next.fetch_add(64);
      if (b >= n) break;
      for (size_t r = b, e_ = std::min(n, b + 64); r < e_; ++r) f(r, cam);
    }
  };
  if (threads <= 1) { worker(); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
  for (auto &t : pool) t.join();
}
void putAll(float *o, const Spectrum &m, const Spectrum *sh, const Spectrum *we) {
  putS(o, m);
  for (int k = 0; k < 4; ++k) { putS(o + 3 * (1 + k), sh[k]); putS(o + 3 * (5 + k), we[k]); }
}
}  // namespace

extern "C" {

// Persistent handles (bench.py's technique reference arms: one build, many sampled gathers); the *_pass entries are
// open + run + close.
struct TechPass {
  World W;
  gvpm_config cfg;
  gvpm_medium med;
  // beams
  BeamWorld B;
  std::vector<std::pair<int, LTPhotonBeam>> list;
  ref<SubBeamBVH<LTPhotonBeam>> beamBVH;
  // planes
  std::vector<LTPhotonPlane> planes;
  ref<PhotonPlaneBVH<LTPhotonPlane>> planeBVH;
  // point photons
  ref<FunctorMap> map;
};

void ref_fn_tech_close(void *h) { delete (TechPass *)h; }

// G-Beams: SubBeamBVH<LTPhotonBeam> (sub-beam split, kd-tree, hierarchy, beams_accel.h:90-243) + BeamGradRadianceQuery per
// camera segment (gvpm.cpp:936-941).  The functor's two sampler draws are preset per (ray, beam) before each call: from the
// caller's table xi [n_rays * n_beams * 2], or (xi = NULL) from the C ABI's counter-based hash evaluated here.
// times_ms: [2] = light-path + beam records (harness work), SubBeamBVH construction.
void *ref_fn_beams_open(const gvpm_beam_soa *bs, size_t n_beams, const gvpm_medium *med, const gvpm_config *cfg, const float *tri,
                        size_t n_tri, float radius, double *times_ms) {
  TechPass *T = new TechPass();
  T->cfg = *cfg;
  T->med = *med;
  const double t0 = nowMs();
  T->W.common(med, cfg, tri, n_tri, cfg->beam_kernel_1d ? EBeamBeam1D : EBeamBeam3D_Optimized);
  T->W.config.newShiftBeam = cfg->beam_kernel_1d != 0;
  if (T->B.build(T->W, bs, n_beams, med, cfg, radius)) { delete T; return NULL; }
  T->list.reserve(n_beams);
  for (size_t j = 0; j < n_beams; ++j) T->list.push_back(std::make_pair((int)j, T->B.beams[j]));
  const double t1 = nowMs();
  T->beamBVH = new SubBeamBVH<LTPhotonBeam>(T->list);
  if (times_ms) { times_ms[0] = t1 - t0; times_ms[1] = nowMs() - t1; }
  return T;
}

int ref_fn_beams_run(void *h, const gvpm_ray_soa *ry, size_t n_rays, const float *xi, int threads, float *out,
                     uint32_t *counts, double *gather_ms) {
  TechPass *T = (TechPass *)h;
  World &W = T->W;
  const size_t n_beams = T->list.size();
  for (size_t r = 0; r < n_rays; ++r)
    if (ry->edge_id[r] < 1 || ry->edge_id[r] > 8) return -5;
  const double t1 = nowMs();
  const LTPhotonBeam *first = &T->list[0].second;
  const size_t stride = sizeof(std::pair<int, LTPhotonBeam>);
  const uint32_t seed = T->cfg.rng_seed;
  parallelRays(n_rays, threads, [&](size_t r, CameraSide &cam) {
    cam.build(ry, r, W.medium.get());
    ref<PresetSampler> sampler = new PresetSampler();
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    BeamGradRadianceQuery gRec(W.scene, &cam.gp, cam.shiftGPs, ray, W.medium.get(), W.config, *W.thdata,
                               (size_t)ry->edge_id[r], sampler.get());
    struct Forward {
      const Ray &baseCameraRay;
      BeamGradRadianceQuery &q;
      PresetSampler *s;
      const float *xi;
      const char *first;
      size_t stride;
      uint32_t seed;
      int px, py, edge;
      uint32_t accepted;
      bool operator()(const LTPhotonBeam *b, Float t1_, Float t2_) {
        const size_t j = (size_t)(((const char *)b - first) / stride);
        if (xi) s->preset(xi[2 * j], xi[2 * j + 1]);
        else s->preset(abiBeamUniform(seed, px, py, edge, (uint32_t)j, 0), abiBeamUniform(seed, px, py, edge, (uint32_t)j, 1));
        const bool ok = q(b, t1_, t2_);
        accepted += ok ? 1u : 0u;
        return ok;
      }
    } fwd{ray, gRec, sampler.get(), xi ? xi + 2 * r * n_beams : NULL, (const char *)first, stride, seed,
          ry->px[r], ry->py[r], ry->edge_id[r], 0u};
    T->beamBVH->query(fwd);
    if (out) putAll(out + 27 * r, gRec.mediumFlux, gRec.shiftedMediumFlux, gRec.weightedMediumFlux);
    if (counts) { counts[2 * r] = fwd.accepted; counts[2 * r + 1] = 0; }
  });
  if (gather_ms) *gather_ms = nowMs() - t1;
  return 0;
}

int ref_fn_beams_pass(const gvpm_beam_soa *bs, size_t n_beams, const gvpm_ray_soa *ry, size_t n_rays, const gvpm_medium *med,
                      const gvpm_config *cfg, const float *tri, size_t n_tri, float radius, const float *xi, int threads,
                      float *out, uint32_t *counts, double *times_ms) {
  double tm[2] = {0, 0}, g = 0;
  void *h = ref_fn_beams_open(bs, n_beams, med, cfg, tri, n_tri, radius, tm);
  if (!h) return -2;
  const int rc = ref_fn_beams_run(h, ry, n_rays, xi, threads, out, counts, &g);
  ref_fn_tech_close(h);
  if (times_ms) { times_ms[0] = tm[1]; times_ms[1] = g; }
  return rc;
}

// G-Planes: PhotonPlaneBVH<LTPhotonPlane> (plane_accel.h:93-185) + PlaneGradRadianceQuery per camera segment (gvpm.cpp:837-841)
void *ref_fn_planes_open(const gvpm_plane_soa *ps, size_t n_planes, const gvpm_medium *med, const gvpm_config *cfg,
                         double *times_ms) {
  TechPass *T = new TechPass();
  T->cfg = *cfg;
  T->med = *med;
  const double t0 = nowMs();
  T->W.common(med, cfg, NULL, 0, EVolPlane0D);
  T->planes.resize(n_planes);
  for (size_t j = 0; j < n_planes; ++j) {
    LTPhotonPlane &p = T->planes[j];
    p._ori = P3(ps->origin + 3 * j);
    p._w0 = V3f(ps->w0 + 3 * j);
    p._length0 = ps->length0[j];
    p._w1 = V3f(ps->w1 + 3 * j);
    p._length1 = ps->length1[j];
    p.medium = T->W.medium.get();
    p._flux = S3(ps->flux + 3 * j);
    p.depth = p.edgeID = ps->edge_id[j];
    p.path = NULL;
    p.pathID = 0;
  }
  const double t1 = nowMs();
  T->planeBVH = new PhotonPlaneBVH<LTPhotonPlane>(T->planes);
  if (times_ms) { times_ms[0] = t1 - t0; times_ms[1] = nowMs() - t1; }
  return T;
}

int ref_fn_planes_run(void *h, const gvpm_ray_soa *ry, size_t n_rays, int threads, float *out, double *gather_ms) {
  TechPass *T = (TechPass *)h;
  World &W = T->W;
  for (size_t r = 0; r < n_rays; ++r)
    if (ry->edge_id[r] < 1 || ry->edge_id[r] > 8) return -5;
  const double t1 = nowMs();
  parallelRays(n_rays, threads, [&](size_t r, CameraSide &cam) {
    cam.build(ry, r, W.medium.get());
    const Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->maxt[r], 0.f);
    PlaneGradRadianceQuery gRec(W.scene, &cam.gp, cam.shiftGPs, ray, W.medium.get(), W.config, *W.thdata, ry->edge_id[r]);
    T->planeBVH->query(gRec);
    if (out) putAll(out + 27 * r, gRec.mediumFlux, gRec.shiftedMediumFlux, gRec.weightedMediumFlux);
  });
  if (gather_ms) *gather_ms = nowMs() - t1;
  return 0;
}

int ref_fn_planes_pass(const gvpm_plane_soa *ps, size_t n_planes, const gvpm_ray_soa *ry, size_t n_rays, const gvpm_medium *med,
                       const gvpm_config *cfg, int threads, float *out, double *times_ms) {
  double tm[2] = {0, 0}, g = 0;
  void *h = ref_fn_planes_open(ps, n_planes, med, cfg, tm);
  if (!h) return -2;
  const int rc = ref_fn_planes_run(h, ry, n_rays, threads, out, &g);
  ref_fn_tech_close(h);
  if (times_ms) { times_ms[0] = tm[1]; times_ms[1] = g; }
  return rc;
}

// G-VPM: GPhotonMap::build + GPhotonMap::evaluate (PointKDTree range query) + VolumeGradientDistanceQuery per distance sample,
// folded per pixel as gvpm.cpp:1175-1182.  Threads take whole pixels; a first pass buckets the samples per ray, in table
// order.  times_ms: [2] = light-path records (harness work), kd build.
void *ref_fn_vpm_open(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_medium *med, const gvpm_config *cfg, const float *tri,
                      size_t n_tri, int threads, double *times_ms) {
  TechPass *T = new TechPass();
  T->cfg = *cfg;
  T->med = *med;
  const double t0 = nowMs();
  T->W.buildThreads = threads;
  if (T->W.build(ph, n_ph, med, cfg, tri, n_tri, EDistance)) { delete T; return NULL; }
  const double t1 = nowMs();
  T->map = new FunctorMap(n_ph);
  for (size_t i = 0; i < n_ph; ++i) T->map->add(T->W.nodes[i]);
  std::vector<GPhotonNodeKD>().swap(T->W.nodes);
  T->map->build(true);
  if (times_ms) { times_ms[0] = t1 - t0; times_ms[1] = nowMs() - t1; }
  return T;
}

int ref_fn_vpm_run(void *h, const gvpm_ray_soa *ry, size_t n_rays, const gvpm_vpm_sample_soa *smp, size_t n_smp,
                   int nb_camera_samples, int threads, float *out, float *mvol, double *gather_ms) {
  TechPass *T = (TechPass *)h;
  World &W = T->W;
  const gvpm_medium *med = &T->med;
  std::vector<std::vector<uint32_t>> perRay(n_rays);
  for (size_t s = 0; s < n_smp; ++s) {
    if (smp->ray[s] >= n_rays) return -6;
    perRay[smp->ray[s]].push_back((uint32_t)s);
  }
  for (size_t r = 0; r < n_rays; ++r)
    if (ry->edge_id[r] < 1 || ry->edge_id[r] > 8) return -5;
  const double t1 = nowMs();
  const Float normalization = 1.f / nb_camera_samples;
  parallelRays(n_rays, threads, [&](size_t r, CameraSide &cam) {
    Spectrum flux(0.f), sh[4], we[4];
    for (int k = 0; k < 4; ++k) sh[k] = we[k] = Spectrum(0.f);
    float found = 0.f;
    for (uint32_t s : perRay[r]) {
      cam.build(ry, r, W.medium.get());
      const size_t e = (size_t)ry->edge_id[r];
      VolumeGradientDistanceQuery gRec(W.scene, &cam.gp, W.config, *W.thdata, cam.shiftGPs, 0, NULL);
      gRec.changeEdge(e, smp->pdf_sel[s]);
      Ray ray(P3(ry->o + 3 * r), V3f(ry->d + 3 * r), ry->mint[r], ry->edge_len[r], 0.f);
      MediumSamplingRecord mRec;
      mRec.t = smp->t[s];
      mRec.p = ray(mRec.t);
      mRec.medium = W.medium.get();
      mRec.sigmaS = S3(med->sigma_s);
      mRec.sigmaA = S3(med->sigma_a);
      mRec.transmittance = S3(smp->transmittance + 3 * s);
      mRec.pdfSuccess = smp->pdf_success[s];
      mRec.pdfSuccessRev = mRec.pdfSuccess;
      mRec.pdfFailure = 0.f;
      mRec.time = 0.f;
      ray.maxt = mRec.t;
      const Float querySize = smp->radius[s];
      gRec.newRayBase(ray, mRec, querySize, mRec.pdfSuccess);
      gRec.clear();
      found += (float)T->map->evaluate(gRec, ray.o + mRec.t * ray.d, querySize);                   // gvpm.cpp:1175
      flux += (gRec.mediumFlux * normalization);
      for (int k = 0; k < 4; ++k) {
        sh[k] += (gRec.shiftedMediumFlux[k] * normalization);
        we[k] += (gRec.weightedMediumFlux[k] * normalization);
      }
    }
    if (out) putAll(out + 27 * r, flux, sh, we);
    if (mvol) mvol[r] = found;
  });
  if (gather_ms) *gather_ms = nowMs() - t1;
  return 0;
}

int ref_fn_vpm_pass(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_ray_soa *ry, size_t n_rays, const gvpm_vpm_sample_soa *smp,
                    size_t n_smp, const gvpm_medium *med, const gvpm_config *cfg, const float *tri, size_t n_tri,
                    int nb_camera_samples, int threads, float *out, float *mvol, double *times_ms) {
  double tm[2] = {0, 0}, g = 0;
  void *h = ref_fn_vpm_open(ph, n_ph, med, cfg, tri, n_tri, threads, tm);
  if (!h) return -2;
  const int rc = ref_fn_vpm_run(h, ry, n_rays, smp, n_smp, nb_camera_samples, threads, out, mvol, &g);
  ref_fn_tech_close(h);
  if (times_ms) { times_ms[0] = tm[1]; times_ms[1] = g; }
  return rc;
}

// ---- round trips through the reference-side shims (gvpm_b200/host/gvpm_mitsuba_shim.hpp, rows a1 / a10) -----------------------
// flattened arrays -> the reference objects this harness rebuilds -> the shim a maintainer compiles into the plugin ->
// flattened arrays again.  The outputs are written through the (const) pointers of the caller's empty records.
}  // extern "C"

namespace {
template <class T> void copyOut(const T *dst, const std::vector<T> &src) {
  std::memcpy(const_cast<T *>(dst), src.data(), src.size() * sizeof(T));
}
}  // namespace

extern "C" {

// gvpm_shim::flattenPhotonMap over a GPhotonMap holding the rebuilt photons (insertion order, not built)
int ref_fn_shim_photons(const gvpm_photon_soa *ph, size_t n_ph, const gvpm_medium *med, const gvpm_config *cfg,
                        const gvpm_photon_soa *out) {
  World W;
  if (int rc = W.build(ph, n_ph, med, cfg, NULL, 0, EVolBRE3D)) return rc;
  ref<FunctorMap> map = new FunctorMap(n_ph);
  for (size_t i = 0; i < n_ph; ++i) map->add(W.nodes[i]);
  gvpm_shim::PhotonArrays a;
  gvpm_shim::flattenPhotonMap(*map, a, W.config.noMediumShift);
  if (a.size() != n_ph) return -7;
  copyOut(out->pos, a.pos); copyOut(out->flux, a.flux); copyOut(out->parent_pos, a.parent_pos);
  copyOut(out->pred_pos, a.pred_pos); copyOut(out->parent_n, a.parent_n); copyOut(out->prefix_flux, a.prefix_flux);
  copyOut(out->parent_albedo, a.parent_albedo); copyOut(out->parent_pdf, a.parent_pdf); copyOut(out->edge_pdf, a.edge_pdf);
  copyOut(out->rr_weight, a.rr_weight); copyOut(out->parent_type, a.parent_type); copyOut(out->depth, a.depth);
  copyOut(out->path_id, a.path_id);
  return 0;
}

// gvpm_shim::appendLightPathBeams on the rebuilt light path of every beam (minDepth = the beam's edge: that edge only)
int ref_fn_shim_beams(const gvpm_beam_soa *bs, size_t n_beams, const gvpm_medium *med, const gvpm_config *cfg, float radius,
                      const gvpm_beam_soa *out) {
  World W;
  W.common(med, cfg, NULL, 0, EBeamBeam3D_Optimized);
  BeamWorld B;
  if (int rc = B.build(W, bs, n_beams, med, cfg, radius)) return rc;
  gvpm_shim::BeamArrays a;
  for (size_t j = 0; j < n_beams; ++j) {
    unsigned int pathCounter = bs->path_id[j];
    size_t skipped = 0;
    if (gvpm_shim::appendLightPathBeams(a, &B.lps[j].path, (int)bs->depth[j], n_beams, Point(0.f), 0.f, pathCounter, skipped,
                                        W.config.noMediumShift) != 1) return -7;
  }
  copyOut(out->origin, a.origin); copyOut(out->end, a.end); copyOut(out->flux, a.flux); copyOut(out->prefix_flux, a.prefix_flux);
  copyOut(out->parent_n, a.parent_n); copyOut(out->parent_albedo, a.parent_albedo); copyOut(out->pred_pos, a.pred_pos);
  copyOut(out->end_n, a.end_n); copyOut(out->parent_pdf, a.parent_pdf); copyOut(out->rr_weight, a.rr_weight);
  copyOut(out->parent_type, a.parent_type); copyOut(out->end_on_surface, a.end_on_surface); copyOut(out->depth, a.depth);
  copyOut(out->path_id, a.path_id);
  return 0;
}

// gvpm_shim::appendGatherPoint on the rebuilt GatherPoint + ShiftGatherPoints of every camera segment
int ref_fn_shim_rays(const gvpm_ray_soa *ry, size_t n_rays, const gvpm_medium *med, const gvpm_config *cfg,
                     const gvpm_ray_soa *out) {
  World W;
  W.common(med, cfg, NULL, 0, EVolBRE3D);
  gvpm_shim::RayArrays a;
  ref<PresetSampler> sampler = new PresetSampler();
  for (size_t r = 0; r < n_rays; ++r) {
    if (ry->edge_id[r] < 1 || ry->edge_id[r] > 8) return -5;
    CameraSide cam;
    cam.build(ry, r, W.medium.get());
    cam.gp.pixel = Point2i(ry->px[r], ry->py[r]);
    sampler->preset(ry->xi[r], 0.f);
    gvpm_shim::appendGatherPoint(a, r, cam.gp, cam.shiftGPs.data(), W.medium.get(), 0, -1, sampler.get());
    if (a.size() != r + 1) return -7;
  }
  copyOut(out->o, a.o); copyOut(out->d, a.d); copyOut(out->mint, a.mint); copyOut(out->maxt, a.maxt);
  copyOut(out->edge_len, a.edge_len); copyOut(out->eye_contrib, a.eye_contrib); copyOut(out->xi, a.xi);
  copyOut(out->px, a.px); copyOut(out->py, a.py); copyOut(out->edge_id, a.edge_id); copyOut(out->off_valid, a.off_valid);
  copyOut(out->off_o, a.off_o); copyOut(out->off_d, a.off_d); copyOut(out->off_len, a.off_len);
  copyOut(out->off_eye, a.off_eye); copyOut(out->off_sensor, a.off_sensor);
  return 0;
}

// gvpm_shim::flattenMedium on the reference's HomogeneousMedium built from the record
int ref_fn_shim_medium(const gvpm_medium *med, gvpm_medium *out) {
  initOnce();
  ref<PhaseFunction> phase = makePhase(med->phase_type, med->hg_g);
  ref<Medium> medium = makeMedium(med->sigma_s, med->sigma_a, med->sampling_weight, phase.get());
  *out = gvpm_shim::flattenMedium(medium.get(), false, med->sampling_weight);
  return 0;
}

}  // extern "C"
