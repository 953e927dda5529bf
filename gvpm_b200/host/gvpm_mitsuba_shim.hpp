// gvpm_mitsuba_shim.hpp — the reference-side half of the boundary (rows a1 / a10 of SURVEY.md §8): flattening of the
// integrator's own objects (light Paths, GPhotonMap / LTBeamMap nodes, GatherPoints with their ShiftGatherPoints, the
// Medium, the Scene's triangle meshes) into the SoA records of include/gvpm_b200.h.  This header is compiled INSIDE the
// gvpm plugin, against the reference's own headers (include/mitsuba/..., src/integrators/photonmapper/gvpm/...): it is
// not part of libgvpm_b200.so and nothing in this repository's product path includes it.  tests/test_mitsuba_shim.py
// compile-checks it against the reference tree where that tree is present.
//
//   #include "gvpm_mitsuba_shim.hpp"          // in gvpm.cpp, after the plugin's own includes
//   gvpm_shim::PhotonArrays ph;  gvpm_shim::flattenPhotonMap(*gradientPhotonMap, ph);        // gvpm.cpp:453-467
//   gvpm_shim::RayArrays ry;     gvpm_shim::flattenGatherPoints(m_gatherBlocks, ...);        // gvpm.cpp:999-1042
//   gvpm_medium med = gvpm_shim::flattenMedium(scene->getMedia()[0].get(), /*volumeOnly*/true);
//   std::vector<float> tris = gvpm_shim::flattenOccluders(scene);
//
// Reference lines restated here (every other value is read from the reference's objects as they are):
//   GPhotonMap::tryAppend + cameraHit          gvpm/gvpm_accel.h:119-199,221-229 (appendLightPath: photons flattened
//                                              straight from a light Path, without building kd nodes first)
//   LTBeamMap::tryAppendLT + LTPhotonBeam      gvpm/gvpm_beams.h:18-43,54-98
//   isIntersectedPoint (camera sphere)         src/integrators/volume_utils.h:154-169
//   getTypeShift / VertexClassifier            gvpm/shift/shift_utilities.h:112-136, gvpm/gvpm_struct.h:46-104
//   the gather loop's ray construction         gvpm/gvpm.cpp:1018-1042
//   eyeContrib / sensorMIS / validVolumeEdge   gvpm/gvpm_struct.h:585-631, shift/shift_cameraPath.h:135-140
#pragma once
#include <vector>

#include "gvpm_b200.h"   // include/gvpm_b200.h

MTS_NAMESPACE_BEGIN
namespace gvpm_shim {

// ---- growable SoA containers ------------------------------------------------------------------------------------------------
struct PhotonArrays {
  std::vector<float> pos, flux, parent_pos, pred_pos, parent_n, prefix_flux, parent_albedo, parent_pdf, edge_pdf, rr_weight;
  std::vector<uint8_t> parent_type, depth;
  std::vector<uint32_t> path_id;
  size_t size() const { return parent_type.size(); }
  void clear() { *this = PhotonArrays(); }
  gvpm_photon_soa view() const {
    gvpm_photon_soa s;
    s.pos = pos.data(); s.flux = flux.data(); s.parent_pos = parent_pos.data(); s.pred_pos = pred_pos.data();
    s.parent_n = parent_n.data(); s.prefix_flux = prefix_flux.data(); s.parent_albedo = parent_albedo.data();
    s.parent_pdf = parent_pdf.data(); s.edge_pdf = edge_pdf.data(); s.rr_weight = rr_weight.data();
    s.parent_type = parent_type.data(); s.depth = depth.data(); s.path_id = path_id.data();
    return s;
  }
};
struct BeamArrays {
  std::vector<float> origin, end, flux, prefix_flux, parent_n, parent_albedo, pred_pos, end_n, parent_pdf, rr_weight;
  std::vector<uint8_t> parent_type, end_on_surface, depth;
  std::vector<uint32_t> path_id;
  size_t size() const { return parent_type.size(); }
  gvpm_beam_soa view() const {
    gvpm_beam_soa s;
    s.origin = origin.data(); s.end = end.data(); s.flux = flux.data(); s.prefix_flux = prefix_flux.data();
    s.parent_n = parent_n.data(); s.parent_albedo = parent_albedo.data(); s.pred_pos = pred_pos.data();
    s.end_n = end_n.data(); s.parent_pdf = parent_pdf.data(); s.rr_weight = rr_weight.data();
    s.parent_type = parent_type.data(); s.end_on_surface = end_on_surface.data(); s.depth = depth.data();
    s.path_id = path_id.data();
    return s;
  }
};
struct RayArrays {
  std::vector<float> o, d, mint, maxt, edge_len, eye_contrib, xi, off_o, off_d, off_len, off_eye, off_sensor;
  std::vector<int32_t> px, py, edge_id;
  std::vector<uint8_t> off_valid;
  std::vector<size_t> gather_point;   // index of the GatherPoint every record belongs to (results are summed per point)
  size_t size() const { return px.size(); }
  gvpm_ray_soa view() const {
    gvpm_ray_soa s;
    s.o = o.data(); s.d = d.data(); s.mint = mint.data(); s.maxt = maxt.data(); s.edge_len = edge_len.data();
    s.eye_contrib = eye_contrib.data(); s.xi = xi.data(); s.px = px.data(); s.py = py.data(); s.edge_id = edge_id.data();
    s.off_valid = off_valid.data(); s.off_o = off_o.data(); s.off_d = off_d.data(); s.off_len = off_len.data();
    s.off_eye = off_eye.data(); s.off_sensor = off_sensor.data();
    return s;
  }
};
inline void push3(std::vector<float> &v, const Point &p) { v.push_back((float)p.x); v.push_back((float)p.y); v.push_back((float)p.z); }
inline void push3(std::vector<float> &v, const Vector &p) { v.push_back((float)p.x); v.push_back((float)p.y); v.push_back((float)p.z); }
inline void push3(std::vector<float> &v, const Spectrum &s) {   // SPECTRUM_SAMPLES = 3 (RGB build)
  v.push_back((float)s[0]); v.push_back((float)s[1]); v.push_back((float)s[2]);
}

// ---- parent classification: which shift the reference would pick for this parent vertex ----------------------------------
// getTypeShift (shift_utilities.h:112-136) sends a parent to the diffuse reconnection when VertexClassifier calls it
// rough (roughness above GPMConfig's threshold; emitter samples always; medium vertices when the phase function's mean
// cosine is <= 0.5 or noMediumShift is set) and to the manifold shift otherwise, which is out of scope
// (GVPM_PARENT_OTHER: the gather fails the shift exactly as the reference does with useManifold = false).
inline uint8_t parentType(const PathVertex *v, bool noMediumShift) {
  switch (v->getType()) {
    case PathVertex::EEmitterSample: return GVPM_PARENT_EMITTER;
    case PathVertex::EMediumInteraction:
      if (noMediumShift) return GVPM_PARENT_MEDIUM;
      return VertexClassifier::type(*v, v->sampledComponentIndex) == VERTEX_TYPE_DIFFUSE ? GVPM_PARENT_MEDIUM : GVPM_PARENT_OTHER;
    case PathVertex::ESurfaceInteraction:
      return VertexClassifier::type(*v, v->sampledComponentIndex) == VERTEX_TYPE_DIFFUSE ? GVPM_PARENT_SURFACE : GVPM_PARENT_OTHER;
    default: return GVPM_PARENT_OTHER;
  }
}
inline Spectrum parentAlbedo(const PathVertex *v) {
  if (v->getType() != PathVertex::ESurfaceInteraction) return Spectrum(0.f);
  const Intersection &its = v->getIntersection();
  const BSDF *bsdf = its.getBSDF();
  return bsdf ? bsdf->getDiffuseReflectance(its) : Spectrum(0.f);
}
inline Vector parentNormal(const PathVertex *v) {
  return v->getType() == PathVertex::EMediumInteraction ? Vector(0.f) : Vector(v->getGeometricNormal());
}

// ---- a1: volume photons ---------------------------------------------------------------------------------------------------------
// one photon = vertex c of light path lt with running importance weight `weight` (GPhotonNodeData, gvpm_accel.h:17-30)
inline void appendPhotonRecord(PhotonArrays &out, const Path *lt, size_t c, const Spectrum &weight, unsigned int pathID,
                               bool noMediumShift) {
  const PathVertex *v = lt->vertex(c), *parent = lt->vertex(c - 1);
  push3(out.pos, v->getPosition());
  push3(out.flux, weight);
  push3(out.parent_pos, parent->getPosition());
  push3(out.pred_pos, c >= 3 ? lt->vertex(c - 2)->getPosition() : Point(1.f));   // shift_volume_photon.cpp:431
  push3(out.parent_n, parentNormal(parent));
  Spectrum prefix(1.f);                                                            // shift_volume_photon.cpp:415-422
  for (size_t i = 0; i + 1 < c; ++i)
    prefix *= lt->vertex(i)->weight[EImportance] * lt->vertex(i)->rrWeight * lt->edge(i)->weight[EImportance];
  push3(out.prefix_flux, prefix);
  push3(out.parent_albedo, parentAlbedo(parent));
  out.parent_pdf.push_back((float)parent->pdf[EImportance]);
  out.edge_pdf.push_back((float)lt->edge(c - 1)->pdf[EImportance]);
  out.rr_weight.push_back((float)parent->rrWeight);
  out.parent_type.push_back(parentType(parent, noMediumShift));
  out.depth.push_back((uint8_t)(c - 1));
  out.path_id.push_back(pathID);
}

// every node of a built GPhotonMap (the reference's tryAppend has already applied minDepth, the capacity and the
// camera-sphere skip).  Node order = the kd-tree's storage order after build(); the gather does not depend on it.
inline void flattenPhotonMap(const GPhotonMap &map, PhotonArrays &out, bool noMediumShift) {
  for (size_t i = 0; i < map.size(); ++i) {
    const GPhotonNodeData &d = map[i].getData();
    appendPhotonRecord(out, d.lightPath, d.vertexId, d.weight, d.pathID, noMediumShift);
  }
}

// GPhotonMap::tryAppend for the VOLUME map (m_storeSurface = false), restated on the flat arrays so that a plugin can
// skip the kd nodes altogether: same start index, same running weight, same capacity rule, same camera-sphere skip,
// same path numbering.  Returns the number of photons appended, -1 when the map was already full.
inline int appendLightPath(PhotonArrays &out, const Path *lightPath, int minDepth, size_t capacity, const Point &sensorPos,
                           Float cameraSphere, unsigned int &nbLightPathAdded, size_t &photonSkip, bool noMediumShift) {
  if (out.size() >= capacity) return -1;
  const size_t startIndex = (size_t)std::max(2, minDepth + 1);
  if (lightPath->vertexCount() <= startIndex) return 0;
  Spectrum importanceWeights(1.f);
  for (size_t i = 0; i < startIndex - 1; i++)
    importanceWeights *= lightPath->vertex(i)->weight[EImportance] * lightPath->vertex(i)->rrWeight *
                         lightPath->edge(i)->weight[EImportance];
  int nbAppend = 0;
  for (size_t i = startIndex; i < lightPath->vertexCount(); i++) {
    importanceWeights *= lightPath->vertex(i - 1)->weight[EImportance] * lightPath->vertex(i - 1)->rrWeight *
                         lightPath->edge(i - 1)->weight[EImportance];
    if (!lightPath->vertex(i)->isMediumInteraction()) continue;
    if (out.size() >= capacity) continue;
    // cameraHit (gvpm_accel.h:221-229): photons whose last segment passes through the sphere around the sensor are left out
    if (cameraSphere != 0.f &&
        isIntersectedPoint(sensorPos, lightPath->vertex(i - 1)->getPosition(), lightPath->vertex(i)->getPosition(), cameraSphere)) {
      photonSkip += 1;
      continue;
    }
    appendPhotonRecord(out, lightPath, i, importanceWeights, nbLightPathAdded, noMediumShift);
    nbAppend++;
  }
  if (nbAppend != 0) nbLightPathAdded += 1;
  return nbAppend;
}

// ---- a13: photon beams ---------------------------------------------------------------------------------------------------------
// LTBeamMap::tryAppendLT + the LTPhotonBeam constructor: beam = light-path edge i inside a medium, i >= max(minDepth, 1)
inline int appendLightPathBeams(BeamArrays &out, const Path *lt, int minDepth, size_t capacity, const Point &sensorPos,
                                Float cameraSphere, unsigned int &nbLightPathAdded, size_t &beamSkip, bool noMediumShift) {
  if (out.size() >= capacity) return -1;
  int nbAppendVol = 0;
  for (size_t i = (size_t)std::max(minDepth, 1); i < lt->edgeCount(); i++) {
    if (lt->edge(i)->medium == nullptr) continue;
    const PathVertex *vi = lt->vertex(i), *vn = lt->vertex(i + 1);
    if (cameraSphere != 0.f && isIntersectedPoint(sensorPos, vi->getPosition(), vn->getPosition(), cameraSphere)) {
      ++beamSkip;
      continue;
    }
    if (out.size() >= capacity) break;
    Spectrum flux(1.f), prefix(1.f);   // gvpm_beams.h:29-35 and shift_volume_beams.cpp:442-449
    for (size_t k = 0; k < i; k++)
      flux *= lt->vertex(k)->rrWeight * lt->vertex(k)->weight[EImportance] * lt->edge(k)->weight[EImportance];
    prefix = flux;
    flux *= vi->weight[EImportance];
    flux *= vi->rrWeight;
    push3(out.origin, vi->getPosition());
    push3(out.end, vn->getPosition());
    push3(out.flux, flux);
    push3(out.prefix_flux, prefix);
    push3(out.parent_n, parentNormal(vi));
    push3(out.parent_albedo, parentAlbedo(vi));
    push3(out.pred_pos, i >= 2 ? lt->vertex(i - 1)->getPosition() : Point(1.f));   // shift_volume_beams.cpp:457
    push3(out.end_n, vn->isOnSurface() ? Vector(vn->getGeometricNormal()) : Vector(0.f));
    out.parent_pdf.push_back((float)vi->pdf[EImportance]);
    out.rr_weight.push_back((float)vi->rrWeight);
    out.parent_type.push_back(parentType(vi, noMediumShift));
    out.end_on_surface.push_back(vn->isOnSurface() ? 1 : 0);
    out.depth.push_back((uint8_t)i);
    out.path_id.push_back(nbLightPathAdded);
    nbAppendVol++;
  }
  if (nbAppendVol == 0) return -1;
  nbLightPathAdded += 1;
  return nbAppendVol;
}

// ---- a10: gather points ---------------------------------------------------------------------------------------------------------
// One record per (gather point, medium edge), as the loop of computeVolumeGradientPhotonBRE builds them (gvpm.cpp:1018-1042).
// shiftGPs[k] must have been generate()d for this pixel (shift_cameraPath.h:29-133): the reference does that lazily inside
// the functor, the shim's caller does it up front.  xi: the sampler->next1D() of gvpm.cpp:1042.
inline void appendGatherPoint(RayArrays &out, size_t gpIndex, const GatherPoint &gp, const ShiftGatherPoint shiftGPs[4],
                              const Medium *medium, int minCameraDepth, int maxCameraDepth, Sampler *sampler) {
  for (int idEdge = 1; idEdge < int(gp.path.edgeCount()); idEdge++) {
    if (minCameraDepth > idEdge) continue;
    if (maxCameraDepth != -1 && idEdge > maxCameraDepth + 1) break;
    if (gp.path.edge(idEdge)->medium == nullptr) continue;
    const Point oBeam = gp.path.vertex(idEdge)->getPosition();
    Vector dBeam = gp.path.vertex(idEdge + 1)->getPosition() - oBeam;
    const Float beamDist = dBeam.length();
    dBeam /= beamDist;
    push3(out.o, oBeam);
    push3(out.d, dBeam);
    out.mint.push_back((float)Epsilon);
    out.maxt.push_back((float)(beamDist - Epsilon));
    out.edge_len.push_back((float)gp.path.edge(idEdge)->length);
    // eyeContrib = getWeightBeam(e - 1) * getWeightVertex(e), gvpm_struct.h:585-592 / shift_volume_photon.cpp:736-748
    push3(out.eye_contrib, gp.getWeightBeam(idEdge - 1) * gp.getWeightVertex(idEdge));
    out.xi.push_back((float)sampler->next1D());
    out.px.push_back(gp.pixel.x);
    out.py.push_back(gp.pixel.y);
    out.edge_id.push_back(idEdge);
    out.gather_point.push_back(gpIndex);
    for (int k = 0; k < 4; ++k) {
      const ShiftGatherPoint &s = shiftGPs[k];
      const bool valid = s.validVolumeEdge((size_t)idEdge, gp.path.edge(idEdge)->medium);   // shift_cameraPath.h:135-140
      out.off_valid.push_back(valid ? 1 : 0);
      if (valid) {
        const PathEdge *e = s.path.edge(idEdge);
        push3(out.off_o, s.path.vertex(idEdge)->getPosition());
        push3(out.off_d, -e->d);   // light-transport edges point towards the sensor: the camera segment runs the other way
        out.off_len.push_back((float)e->length);
        push3(out.off_eye, s.getWeightBeam(idEdge - 1) * s.getWeightVertex(idEdge));
        // sensorMIS (gvpm_struct.h:608-631): the distance factors cancel in ratio * jacobian, a per-offset-ray constant
        out.off_sensor.push_back((float)s.sensorMIS((size_t)idEdge, gp, (Float)1, (Float)1));
      } else {
        push3(out.off_o, Point(0.f)); push3(out.off_d, Vector(0.f, 0.f, 1.f)); out.off_len.push_back(0.f);
        push3(out.off_eye, Spectrum(0.f)); out.off_sensor.push_back(0.f);
      }
    }
  }
}

// ---- scene constants ----------------------------------------------------------------------------------------------------------------
// volumeOnly: GPMIntegrator calls computeOnlyVolumeInteraction() on every medium for volume-only renders (gvpm.cpp:135-141),
// which sets the sampling weight to 1; otherwise the plugin passes the medium's mediumSamplingWeight property.
inline gvpm_medium flattenMedium(const Medium *m, bool volumeOnly, float mediumSamplingWeight = 1.f) {
  gvpm_medium out;
  const Spectrum sigS = m->getSigmaS(), sigA = m->getSigmaA();
  for (int c = 0; c < 3; ++c) { out.sigma_s[c] = (float)sigS[c]; out.sigma_a[c] = (float)sigA[c]; }
  const PhaseFunction *ph = m->getPhaseFunction();
  const Float g = ph->getMeanCosine();
  out.phase_type = g == 0 ? GVPM_PHASE_ISOTROPIC : GVPM_PHASE_HG;   // isotropic.cpp / hg.cpp are the in-scope phase functions
  out.hg_g = (float)g;
  out.sampling_weight = volumeOnly ? 1.f : mediumSamplingWeight;
  return out;
}

// the triangle soup the reconnection shadow ray is tested against (scene->rayIntersect, shift_volume_photon.cpp:396-402)
inline std::vector<float> flattenOccluders(const Scene *scene) {
  std::vector<float> tri;
  const std::vector<TriMesh *> &meshes = scene->getMeshes();
  for (size_t m = 0; m < meshes.size(); ++m) {
    const TriMesh *mesh = meshes[m];
    const Point *pos = mesh->getVertexPositions();
    const Triangle *t = mesh->getTriangles();
    for (size_t i = 0; i < mesh->getTriangleCount(); ++i)
      for (int k = 0; k < 3; ++k) push3(tri, pos[t[i].idx[k]]);
  }
  return tri;
}

}  // namespace gvpm_shim
MTS_NAMESPACE_END
