// gvpm_host.hpp — host-side mirror of the reference's gather drivers, above the C ABI.
//
// GPMIntegrator keeps its XML parameters, its iteration loop, the per-iteration radius reduction,
// the per-pixel GatherPoint accumulators and the Poisson hand-off; only the bodies of the gather
// drivers change.  This header restates those driver bodies with the CPU gather replaced by calls
// into include/gvpm_b200.h, using the reference's names:
//     computeVolumeGradientPhotonBRE    gvpm/gvpm.cpp:988-1079
//     scaleVolumeAPA                    gvpm/gvpm.cpp:181-215
//     computeGradient                   gvpm/gvpm.cpp:1205-1306
// A patched gvpm.cpp would hold one VolumeGatherB200 next to m_gatherBlocks and call it from
// photonMapPass.  Errors: the reference raises through SLog(EError) (a
// std::runtime_error); so does this shim, carrying gvpm_last_error().
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gvpm_b200.h"

namespace gvpm_host {

// EVolumeTechnique values the gather supports (src/integrators/volume_utils.h:55-93)
enum EVolumeTechnique { EVolBRE2D = 0, EVolBRE3D = 1 };

// The GPMConfig fields the volume gather path reads (gvpm/gvpm_struct.h:181-333), same names.
struct GPMConfig {
  int maxDepth = -1, minDepth = 0;
  double alpha = 0.7;                 // Float alpha
  double initialScaleVolume = 1.0;
  int volTechnique = EVolBRE3D;
  int lightingInteractionMode = GVPM_ALL2MEDIA;
  bool useMIS = true, useShiftNull = false, pathSet = true, powerHeuristic = false;
  bool use3DKernelReduction = false;
  std::string forceAPA;               // "", "1D", "2D", "3D"
};

// GPMIntegrator::scaleVolumeAPA, gvpm.cpp:181-215 (m_independentScale = false): host-side, double.
inline void scaleVolumeAPA(double &globalScaleVolume, int it, const GPMConfig &config) {
  it -= 1;  // "Fix the bug as it == 1 at the first iteration."
  const double ratioVolAPA = (it + config.alpha) / (it + 1);
  const bool k3 = config.volTechnique == EVolBRE3D, k2 = config.volTechnique == EVolBRE2D;
  if (config.forceAPA.empty()) {
    if (k3 || config.use3DKernelReduction) globalScaleVolume *= std::cbrt(ratioVolAPA);
    else if (k2) globalScaleVolume *= std::sqrt(ratioVolAPA);
    else globalScaleVolume *= ratioVolAPA;
  } else if (config.forceAPA == "1D") {
    globalScaleVolume *= ratioVolAPA;
  } else if (config.forceAPA == "2D") {
    globalScaleVolume *= std::sqrt(ratioVolAPA);
  } else if (config.forceAPA == "3D") {
    globalScaleVolume *= std::cbrt(ratioVolAPA);
  } else {
    throw std::runtime_error("No Force APA: " + config.forceAPA);
  }
}

// GatherPoint's volume accumulators (gvpm/gvpm_struct.h:421-455), one per pixel, SoA of 27 floats:
// mediumFlux[3], shiftedMediumFlux[4][3], weightedMediumFlux[4][3].
class VolumeGatherB200 {
 public:
  VolumeGatherB200(int device, int width, int height, const GPMConfig &config, const gvpm_medium &medium,
                   float mediumBSphereRadius, const float *occluderTris, size_t nTris)
      : globalScaleVolume(config.initialScaleVolume), m_config(config), m_w(width), m_h(height),
        m_bsphereR(mediumBSphereRadius), m_acc((size_t)width * height * GVPM_OUT_FLOATS, 0.f),
        m_haveSmoke((size_t)width * height, 0) {
    check(gvpm_ctx_create(device, &m_ctx), "gvpm_ctx_create");
    check(gvpm_set_medium(m_ctx, &medium), "gvpm_set_medium");
    gvpm_config c{};
    c.max_depth = config.maxDepth;
    c.min_depth = config.minDepth;
    c.lighting_mode = config.lightingInteractionMode;
    c.use_mis = config.useMIS;
    c.use_shift_null = config.useShiftNull;
    c.path_set = config.pathSet;
    c.power_heuristic = config.powerHeuristic;
    c.kernel_3d = config.volTechnique == EVolBRE3D;
    c.film_w = width;
    c.film_h = height;
    c.shadow_maxt_scale = 1e-3f;  // ShadowEpsilon, shift_volume_photon.cpp:396
    c.epsilon = 1e-4f;            // Epsilon
    check(gvpm_set_config(m_ctx, &c), "gvpm_set_config");
    check(gvpm_set_occluders(m_ctx, occluderTris, nTris), "gvpm_set_occluders");
  }
  ~VolumeGatherB200() { if (m_ctx) gvpm_ctx_destroy(m_ctx); }
  VolumeGatherB200(const VolumeGatherB200 &) = delete;
  VolumeGatherB200 &operator=(const VolumeGatherB200 &) = delete;

  // breInitSize = m_smokeAABB.getBSphere().radius * globalScaleVolume * POURCENTAGE_BS (gvpm.cpp:989)
  float currentRadius() const { return (float)(m_bsphereR * (float)globalScaleVolume * 0.01f); }

  // gvpm.cpp:988-1079.  `rays` holds one record per (gather point, medium edge); several records may
  // share a pixel (px,py) and are summed like the idEdge loop does (:1018-1052).
  void computeVolumeGradientPhotonBRE(int it, const gvpm_photon_soa *photons, size_t nPhotons,
                                      const gvpm_ray_soa *rays, size_t nRays, size_t nbPathVolume) {
    check(gvpm_upload_photons(m_ctx, photons, nPhotons), "gvpm_upload_photons");
    check(gvpm_build_points(m_ctx, currentRadius()), "gvpm_build_points");
    check(gvpm_upload_rays(m_ctx, rays, nRays), "gvpm_upload_rays");
    m_iter.assign(nRays * GVPM_OUT_FLOATS, 0.f);
    check(gvpm_gather_bre(m_ctx, m_iter.data(), nullptr), "gvpm_gather_bre");
    // sum the medium edges of each pixel, normalise, fold into the APA running mean (:1054-1069)
    std::vector<float> pix(m_acc.size(), 0.f);
    for (size_t r = 0; r < nRays; ++r) {
      const int x = rays->px[r], y = rays->py[r];
      if (x < 0 || y < 0 || x >= m_w || y >= m_h) continue;
      const size_t p = (size_t)y * m_w + x;
      m_haveSmoke[p] = 1;
      for (int j = 0; j < GVPM_OUT_FLOATS; ++j) pix[p * GVPM_OUT_FLOATS + j] += m_iter[r * GVPM_OUT_FLOATS + j];
    }
    const float nb = (float)nbPathVolume;
    for (size_t i = 0; i < m_acc.size(); ++i) {
      const float fluxVolIter = pix[i] / nb;
      m_acc[i] = (m_acc[i] * (float)(it - 1) + fluxVolIter) / (float)it;  // APA estimator
    }
    scaleVolumeAPA(it);
  }

  void scaleVolumeAPA(int it) { gvpm_host::scaleVolumeAPA(globalScaleVolume, it, m_config); }

  // gvpm.cpp:1205-1306 + throughput plane (:479-500): interleaved RGB, row-major, poisson hand-off
  void computeGradient(float *throughput, float *gX, float *gY, bool useAbs) {
    check(gvpm_compute_gradient(m_ctx, m_acc.data(), m_w, m_h, useAbs ? 1 : 0, throughput, gX, gY),
          "gvpm_compute_gradient");
  }

  const std::vector<float> &accumulators() const { return m_acc; }
  const std::vector<uint8_t> &haveSmoke() const { return m_haveSmoke; }
  gvpm_ctx *context() { return m_ctx; }
  double globalScaleVolume;

 private:
  void check(int rc, const char *what) {
    if (rc != GVPM_OK)
      throw std::runtime_error(std::string(what) + " failed: " + gvpm_last_error(m_ctx));
  }
  GPMConfig m_config;
  gvpm_ctx *m_ctx = nullptr;
  int m_w, m_h;
  float m_bsphereR;
  std::vector<float> m_acc, m_iter;
  std::vector<uint8_t> m_haveSmoke;
};

}  // namespace gvpm_host
