// gvpm_host.hpp — host-side mirror of the reference's gather drivers, above the C ABI.
//
// GPMIntegrator keeps its XML parameters, its iteration loop, the per-iteration radius reduction,
// the per-pixel GatherPoint accumulators and the Poisson hand-off; only the bodies of the gather
// drivers change.  This header restates those driver bodies with the CPU gather replaced by calls
// into include/gvpm_b200.h, using the reference's names:
//     computeVolumeGradientPhotonBRE    gvpm/gvpm.cpp:988-1079
//     computeVolumeGradientPhoton       gvpm/gvpm.cpp:1081-1203   (G-VPM)
//     computeVolumeGradientBeams        gvpm/gvpm.cpp:880-986     (G-Beams)
//     computeVolumeGradientPlanes       gvpm/gvpm.cpp:782-878     (G-Planes 0D) + LTPhotonPlane::transformBeam
//     scaleVolumeAPA                    gvpm/gvpm.cpp:181-215
//     computeGradient                   gvpm/gvpm.cpp:1205-1306
// and, for the sppm plugin's primal volume passes (class SPPMVolumeGatherB200):
//     volumePhotonPassBRE               sppm.cpp:905-1001
//     volumePhotonBeamPass              sppm.cpp:765-880          (beam1d / beam3d_naive / beam3d_egsr / beam3d)
//     scaleVolumeAPA                    sppm.cpp:255-290
// A patched gvpm.cpp would hold one VolumeGatherB200 next to m_gatherBlocks and call it from
// photonMapPass.  Errors: the reference raises through SLog(EError) (a
// std::runtime_error); so does this shim, carrying gvpm_last_error().
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <map>
#include <string>
#include <vector>

#include "../../include/gvpm_b200.h"

namespace gvpm_host {

// EVolumeTechnique values the gather supports (src/integrators/volume_utils.h:55-93)
enum EVolumeTechnique { EVolBRE2D = 0, EVolBRE3D = 1, EVolVPM = 2, EVolBeam3D = 3, EVolPlane0D = 4, EVolBeam1D = 5 };

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

// Stand-in for mitsuba::Properties as GPMConfig::load uses it (include/mitsuba/core/properties.h): typed getters with
// a default, values kept as the strings of the scene XML.  A Mitsuba build passes its own Properties instead.
class Properties {
 public:
  void set(const std::string &key, const std::string &value) { m_values[key] = value; }
  bool has(const std::string &key) const { return m_values.count(key) != 0; }
  std::string getString(const std::string &key, const std::string &def) const {
    auto it = m_values.find(key);
    return it == m_values.end() ? def : it->second;
  }
  bool getBoolean(const std::string &key, bool def) const {
    if (!has(key)) return def;
    const std::string v = getString(key, "");
    if (v == "true") return true;
    if (v == "false") return false;
    throw std::runtime_error("Property \"" + key + "\" has the wrong type (expected <boolean>).");
  }
  long getInteger(const std::string &key, long def) const {
    if (!has(key)) return def;
    size_t used = 0;
    const std::string v = getString(key, "");
    long r = 0;
    try { r = std::stol(v, &used); } catch (...) { used = 0; }
    if (used != v.size() || v.empty()) throw std::runtime_error("Property \"" + key + "\" has the wrong type (expected <integer>).");
    return r;
  }
  double getFloat(const std::string &key, double def) const {
    if (!has(key)) return def;
    size_t used = 0;
    const std::string v = getString(key, "");
    double r = 0;
    try { r = std::stod(v, &used); } catch (...) { used = 0; }
    if (used != v.size() || v.empty()) throw std::runtime_error("Property \"" + key + "\" has the wrong type (expected <float>).");
    return r;
  }

 private:
  std::map<std::string, std::string> m_values;
};

// The further GPMConfig members that GPMConfig::load reads next to the gather's own (gvpm_struct.h:106-176), for the
// host code around the gather: same names, same defaults.
struct GPMConfigExtra {
  int rrDepth = 12, photonCount = 250000, volumePhotonCount = 250000, maxPasses = -1, dumpIteration = 5;
  bool reconstructL1 = false, reconstructL2 = true;
  double reconstructAlpha = 0.2;
  bool useManifold = true, noMediumShift = true, convertLong = false, newShiftBeam = false, deterministic = false;
  int nbCameraSamples = 40;
  long minCameraDepth = 0;
  int maxCameraDepth = -1;
  double cameraSphere = 1.0;
};

// parseVolumeTechnique (src/integrators/volume_utils.h:54-92) restricted to what the gvpm gather serves:
// "distance" is the point-photon (G-VPM) path of computeVolumeGradientPhoton, gvpm.cpp:1081-1203; BeamKernelRecord::eval
// has no naive / EGSR branch in gvpm (shift_volume_beams.h:169,195,285-287).
inline int parseVolumeTechnique(const std::string &volRenderingTech) {
  if (volRenderingTech == "distance") return EVolVPM;
  if (volRenderingTech == "bre" || volRenderingTech == "bre3d") return EVolBRE3D;
  if (volRenderingTech == "bre2d") return EVolBRE2D;
  if (volRenderingTech == "beam" || volRenderingTech == "beam1d") return EVolBeam1D;
  if (volRenderingTech == "beam3d" || volRenderingTech == "beam3d_optimized") return EVolBeam3D;
  if (volRenderingTech == "beam3d_naive" || volRenderingTech == "beam3d_egsr") throw std::runtime_error("Not supported kernel type");
  if (volRenderingTech == "plane0d") return EVolPlane0D;
  throw std::runtime_error("Unknow vol technique: " + volRenderingTech);
}
// parseMediaInteractionMode, volume_utils.h:109-128 (ELightingEffects bits)
inline int parseMediaInteractionMode(const std::string &m) {
  if (m == "all2media") return (1 << 2) | (1 << 4);
  if (m == "all2surf") return (1 << 1) | (1 << 3);
  if (m == "surf2surf") return 1 << 1;
  if (m == "media2surf") return 1 << 3;
  if (m == "surf2media") return 1 << 2;
  if (m == "media2media") return 1 << 4;
  if (m == "all2all") return (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4);
  throw std::runtime_error("Invalid media interaction mode: " + m);
}

// GPMConfig::load (gvpm_struct.h:181-333) + the GPMIntegrator constructor's fix-up (gvpm.cpp:93-98): the plugin's XML
// parameters, their defaults and their error conditions.  SLog(EError, ...) throws, as in the reference.
inline void loadGPMConfig(const Properties &props, GPMConfig &c, GPMConfigExtra &x) {
  auto lower = [](std::string s) { for (char &ch : s) ch = (char)std::tolower((unsigned char)ch); return s; };
  x.newShiftBeam = props.getBoolean("newShiftBeam", false);
  c.initialScaleVolume = props.getFloat("initialScaleVolume", 1.0);
  c.use3DKernelReduction = props.getBoolean("use3DKernelReduction", false);
  c.forceAPA = props.getString("forceAPA", "");
  x.noMediumShift = props.getBoolean("noMediumShift", true);
  c.powerHeuristic = props.getBoolean("powerHeuristic", false);
  const double relaxME = props.getFloat("relaxME", 1.0);
  if (relaxME != 0.0 && relaxME != 1.0) throw std::runtime_error("relaxME options need to be 0 or 1.");
  c.alpha = props.getFloat("alpha", .7);
  x.photonCount = (int)props.getInteger("photonCount", 250000);
  x.volumePhotonCount = (int)props.getInteger("volumePhotonCount", 250000);
  c.maxDepth = (int)props.getInteger("maxDepth", -1);
  c.minDepth = (int)props.getInteger("minDepth", 0);
  x.rrDepth = (int)props.getInteger("rrDepth", 12);
  x.maxPasses = (int)props.getInteger("maxPasses", -1);
  if (c.maxDepth <= 1 && c.maxDepth != -1) throw std::runtime_error("Maximum depth must be set to \"2\" or higher!");
  if (x.maxPasses <= 0 && x.maxPasses != -1)
    throw std::runtime_error("Maximum number of passes must either be set to \"-1\" or \"1\" or higher!");
  x.dumpIteration = (int)props.getInteger("dumpIteration", 5);
  x.reconstructL1 = props.getBoolean("reconstructL1", false);
  x.reconstructL2 = props.getBoolean("reconstructL2", true);
  x.reconstructAlpha = props.getFloat("reconstructAlpha", 0.2);
  const double bounceRoughness = props.getFloat("bounceRoughness", 0.001);
  if (bounceRoughness <= 0.0 || bounceRoughness > 1.0) throw std::runtime_error("Bad roughtness constant: " + std::to_string(bounceRoughness));
  x.useManifold = props.getBoolean("useManifold", true);
  const std::string mis = lower(props.getString("useMIS", "area"));
  if (mis == "area") c.useMIS = true;
  else if (mis == "none") c.useMIS = false;
  else throw std::runtime_error("useMIS: need to be 'none' or 'area'");
  std::string strLightingMode = props.getString("lightingInteractionMode", "");
  if (strLightingMode.empty()) strLightingMode = props.getString("interactionMode", "all2all");
  c.lightingInteractionMode = parseMediaInteractionMode(strLightingMode);
  x.convertLong = props.getBoolean("convertLong", false);
  c.volTechnique = parseVolumeTechnique(props.getString("volTechnique", "distance"));
  if (!(c.lightingInteractionMode & ((1 << 1) | (1 << 3)))) x.photonCount = 0;         // !needSurfaceRendering()
  if (!(c.lightingInteractionMode & ((1 << 2) | (1 << 4)))) x.volumePhotonCount = 0;   // !needVolumeRendering()
  c.useShiftNull = props.getBoolean("useShiftNull", false);
  x.nbCameraSamples = (int)props.getInteger("nbCameraSamples", 40);
  const bool use3DKernel = c.volTechnique == EVolVPM || c.volTechnique == EVolBRE3D || c.volTechnique == EVolBeam3D;
  if (c.useShiftNull && !use3DKernel && c.volTechnique != EVolBeam1D)
    throw std::runtime_error("Not possible to shift null without using 3D kernel");
  x.minCameraDepth = props.getInteger("minCameraDepth", 0);
  if (x.minCameraDepth < 0) throw std::runtime_error("minCamera depth need to be null or positive");
  x.maxCameraDepth = (int)props.getInteger("maxCameraDepth", -1);
  x.deterministic = props.getBoolean("deterministic", false);
  x.cameraSphere = props.getFloat("cameraSphere", 1.0);
  c.pathSet = props.getBoolean("pathSet", true);
  if (c.pathSet && x.deterministic)
    throw std::runtime_error("It is not possible to use pathSet and deterministic option at the same time");
  if (c.volTechnique == EVolBeam1D) x.newShiftBeam = true;   // "Use the correct fix here", gvpm.cpp:96-98
}

// GPMIntegrator::scaleVolumeAPA, gvpm.cpp:181-215 (m_independentScale = false): host-side.  Real = the reference build's
// Float: alpha and globalScaleVolume are Floats there, so `(it + alpha) / (it + 1)` is evaluated in Float before it is
// widened to the double ratioVolAPA, and the product with cbrt / sqrt (double) is rounded back to Float.  double = the
// reference's CMake default (DOUBLE_PRECISION) and what the drivers below keep; float reproduces a SINGLE_PRECISION build
// bit for bit (gvpm_host_scale_apa_f32; checked against the reference's compiled function).
template <typename Real>
inline void scaleVolumeAPA(Real &globalScaleVolume, int it, const GPMConfig &config) {
  it -= 1;  // "Fix the bug as it == 1 at the first iteration."
  const double ratioVolAPA = (double)(((Real)it + (Real)config.alpha) / (Real)(it + 1));
  // EVolumeTechniqueHelper::use3DKernel (volume_utils.h:35-41): EDistance, EVolBRE3D and every EBeamBeam3D_*;
  // beam1d and plane0d reduce linearly, bre2d by the square root
  const bool k3 = config.volTechnique == EVolVPM || config.volTechnique == EVolBRE3D || config.volTechnique == EVolBeam3D,
             k2 = config.volTechnique == EVolBRE2D;
  if (config.forceAPA.empty()) {
    if (k3 || config.use3DKernelReduction) globalScaleVolume = (Real)(globalScaleVolume * std::cbrt(ratioVolAPA));
    else if (k2) globalScaleVolume = (Real)(globalScaleVolume * std::sqrt(ratioVolAPA));
    else globalScaleVolume = (Real)(globalScaleVolume * ratioVolAPA);
  } else if (config.forceAPA == "1D") {
    globalScaleVolume = (Real)(globalScaleVolume * ratioVolAPA);
  } else if (config.forceAPA == "2D") {
    globalScaleVolume = (Real)(globalScaleVolume * std::sqrt(ratioVolAPA));
  } else if (config.forceAPA == "3D") {
    globalScaleVolume = (Real)(globalScaleVolume * std::cbrt(ratioVolAPA));
  } else {
    throw std::runtime_error("No Force APA: " + config.forceAPA);
  }
}

// LTPhotonPlane::transformBeam, gvpm/gvpm_plane.h:53-73: extends a photon beam (origin, unit dir) to a photon plane
// by sampling a free-flight distance (HomogeneousMedium::sampleDistance with EDistanceNormal on Ray(o, d, time):
// mint = Epsilon, maxt = inf; balance strategy => density sigma_t[1], src/medium/homogeneous.cpp:293-352) and a
// scattering direction from the phase function (isotropic: squareToUniformSphere, src/libcore/warp.cpp:25-31;
// HG: src/phase/hg.cpp:73-96 in FrameCoherent(dir)), rejecting directions parallel to the beam.  This consumes the
// integrator's sampler on the host, exactly like the reference (gvpm.cpp:793-797), so the planes that cross the
// C ABI are bit-identical inputs.  Sampler: float next1D(); void next2D(float &x, float &y).
template <typename Sampler>
inline void transformBeam(const float dir[3], const gvpm_medium &m, Sampler &sampler, float w1[3], float &length1) {
  const float Epsilon = 1e-4f, TWO_PI_F = 2.0f * 3.14159265358979323846f;
  float rand = sampler.next1D();
  if (!(rand < m.sampling_weight))
    // the reference reads an unset mRecNew.t in this case; volume-only renders force the weight to 1 (gvpm.cpp:135-141)
    throw std::runtime_error("transformBeam: mediumSamplingWeight must be 1 for photon planes");
  rand /= m.sampling_weight;
  const float samplingDensity = m.sigma_s[1] + m.sigma_a[1];
  const float sampledDistance = -std::log(1 - rand) / samplingDensity;
  length1 = sampledDistance + Epsilon;  // mRec.t = sampledDistance + ray.mint
  for (;;) {
    float sx, sy;
    sampler.next2D(sx, sy);
    if (m.phase_type == GVPM_PHASE_ISOTROPIC) {
      const float z = 1.0f - 2.0f * sy;
      const float r = std::sqrt(std::max(0.0f, 1.0f - z * z));
      const float phi = TWO_PI_F * sx;
      w1[0] = r * std::cos(phi); w1[1] = r * std::sin(phi); w1[2] = z;
    } else {
      const float g = m.hg_g;
      float cosTheta;
      if (std::fabs(g) < Epsilon) cosTheta = 1 - 2 * sx;
      else {
        const float sqrTerm = (1 - g * g) / (1 - g + 2 * g * sx);
        cosTheta = (1 + g * g - sqrTerm * sqrTerm) / (2 * g);
      }
      const float sinTheta = std::sqrt(std::max(0.0f, 1.0f - cosTheta * cosTheta));
      const float phi = TWO_PI_F * sy;
      const float lx = sinTheta * std::cos(phi), ly = sinTheta * std::sin(phi), lz = cosTheta;
      // FrameCoherent(-pRec.wi) with pRec.wi = -dir: coordinateSystemCoherent (src/libcore/util.cpp:592-599)
      const float sign = std::copysign(1.0f, dir[2]);
      const float a = -1.0f / (sign + dir[2]);
      const float b = dir[0] * dir[1] * a;
      const float s[3] = {1.0f + sign * dir[0] * dir[0] * a, sign * b, -sign * dir[0]};
      const float t[3] = {b, sign + dir[1] * dir[1] * a, -dir[1]};
      for (int i = 0; i < 3; ++i) w1[i] = s[i] * lx + t[i] * ly + dir[i] * lz;
    }
    if (std::fabs(dir[0] * w1[0] + dir[1] * w1[1] + dir[2] * w1[2]) != 1.0f) break;
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
    c.kernel_3d = config.volTechnique != EVolBRE2D;
    c.beam_kernel_1d = config.volTechnique == EVolBeam1D;  // "beam1d": newShiftBeam is forced on, gvpm.cpp:96-98
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
    foldIteration(it, rays, nRays, (float)nbPathVolume);  // :1054-1069
    scaleVolumeAPA(it);
  }

  // gvpm.cpp:880-986 (beam3d / beam1d).  beamRadius = bsphereR * globalScaleVolume * POURCENTAGE_BS (:881).
  void computeVolumeGradientBeams(int it, const gvpm_beam_soa *beams, size_t nBeams, const gvpm_ray_soa *rays,
                                  size_t nRays, size_t nbPathBeams) {
    check(gvpm_upload_beams(m_ctx, beams, nBeams), "gvpm_upload_beams");
    check(gvpm_build_beams(m_ctx, currentRadius()), "gvpm_build_beams");
    check(gvpm_upload_rays(m_ctx, rays, nRays), "gvpm_upload_rays");
    m_iter.assign(nRays * GVPM_OUT_FLOATS, 0.f);
    check(gvpm_gather_beams(m_ctx, m_iter.data(), nullptr), "gvpm_gather_beams");
    foldIteration(it, rays, nRays, (float)nbPathBeams);  // :958-976
    scaleVolumeAPA(it);
  }

  // gvpm.cpp:782-878.  `planes` were produced on the host by transformBeam from the iteration's beams (:793-797).
  void computeVolumeGradientPlanes(int it, const gvpm_plane_soa *planes, size_t nPlanes, const gvpm_ray_soa *rays,
                                   size_t nRays, size_t nbPathBeams) {
    check(gvpm_upload_planes(m_ctx, planes, nPlanes), "gvpm_upload_planes");
    check(gvpm_build_planes(m_ctx), "gvpm_build_planes");
    check(gvpm_upload_rays(m_ctx, rays, nRays), "gvpm_upload_rays");
    m_iter.assign(nRays * GVPM_OUT_FLOATS, 0.f);
    check(gvpm_gather_planes(m_ctx, m_iter.data(), nullptr), "gvpm_gather_planes");
    foldIteration(it, rays, nRays, (float)nbPathBeams);  // :850-868
    scaleVolumeAPA(it);
  }

  // gvpm.cpp:1081-1203 (G-VPM).  The distance samples are drawn on the host (:1141-1175); their radius is
  // BBPourcentageCONST * gp.scaleVol per pixel (:1131), `maxRadius` bounds them for the hierarchy.  mvol (may be
  // null) receives MVol per ray for the per-pixel radius update (:1191-1195), which stays with the caller.
  void computeVolumeGradientPhoton(int it, const gvpm_photon_soa *photons, size_t nPhotons, const gvpm_ray_soa *rays,
                                   size_t nRays, const gvpm_vpm_sample_soa *samples, size_t nSamples,
                                   int nbCameraSamples, float maxRadius, size_t nbPathVolume, uint32_t *mvol) {
    check(gvpm_upload_photons(m_ctx, photons, nPhotons), "gvpm_upload_photons");
    check(gvpm_build_points(m_ctx, maxRadius), "gvpm_build_points");
    check(gvpm_upload_rays(m_ctx, rays, nRays), "gvpm_upload_rays");
    check(gvpm_upload_vpm_samples(m_ctx, samples, nSamples), "gvpm_upload_vpm_samples");
    m_iter.assign(nRays * GVPM_OUT_FLOATS, 0.f);
    check(gvpm_gather_vpm(m_ctx, nbCameraSamples, m_iter.data(), mvol, nullptr), "gvpm_gather_vpm");
    // VPM is not an APA estimator: the gather adds onto gp.mediumFlux (:1176-1181) and the image assembly divides
    // by m_totalEmittedVolume (:487-491, isAPAVolumeEstimator() false)
    (void)it;
    for (size_t r = 0; r < nRays; ++r) {
      const int x = rays->px[r], y = rays->py[r];
      if (x < 0 || y < 0 || x >= m_w || y >= m_h) continue;
      const size_t p = (size_t)y * m_w + x;
      m_haveSmoke[p] = 1;
      for (int j = 0; j < GVPM_OUT_FLOATS; ++j) m_acc[p * GVPM_OUT_FLOATS + j] += m_iter[r * GVPM_OUT_FLOATS + j];
    }
    totalEmittedVolume += nbPathVolume;
  }
  // accumulators of the non-APA estimator divided by m_totalEmittedVolume (gvpm.cpp:487-491)
  std::vector<float> normalizedAccumulators() const {
    std::vector<float> a(m_acc);
    if (totalEmittedVolume) {
      const float inv = 1.0f / (float)totalEmittedVolume;   // Spectrum / Float: reciprocal multiply (spectrum.h:415-425)
      for (float &v : a) v *= inv;
    }
    return a;
  }

  void scaleVolumeAPA(int it) { gvpm_host::scaleVolumeAPA(globalScaleVolume, it, m_config); }

  // gvpm.cpp:1205-1306 + throughput plane (:479-500): interleaved RGB, row-major, poisson hand-off
  // reusePrimal (GPMConfig::reusePrimal): the throughput plane by gvpm.cpp:503-532 instead of mediumFlux
  void computeGradient(float *throughput, float *gX, float *gY, bool useAbs, bool reusePrimal = false) {
    if (reusePrimal) {
      const bool apa = m_config.volTechnique != EVolVPM;   // isAPAVolumeEstimator
      const float inv = apa || !totalEmittedVolume ? 1.f : 1.f / (float)totalEmittedVolume;
      check(gvpm_compute_gradient_reuse_primal(m_ctx, m_acc.data(), m_w, m_h, useAbs ? 1 : 0, inv, throughput, gX, gY),
            "gvpm_compute_gradient_reuse_primal");
      return;
    }
    check(gvpm_compute_gradient(m_ctx, m_acc.data(), m_w, m_h, useAbs ? 1 : 0, throughput, gX, gY),
          "gvpm_compute_gradient");
  }

  // gvpm.cpp:554-690: computeGradient + poisson::Solver (preset "L2D" for reconstructL2, "L1D" for reconstructL1,
  // alpha = reconstructAlpha) in one device round trip.  direct may be null; throughput / gX / gY may be null.
  void reconstruct(const char *preset, float reconstructAlpha, bool useAbs, const float *direct, float *throughput,
                   float *gX, float *gY, float *reconstruction) {
    gvpm_poisson_params p;
    check(gvpm_poisson_preset(preset, &p), "gvpm_poisson_preset");
    p.alpha = reconstructAlpha;
    check(gvpm_reconstruct(m_ctx, m_acc.data(), m_w, m_h, useAbs ? 1 : 0, direct, &p, throughput, gX, gY, reconstruction),
          "gvpm_reconstruct");
  }

  const std::vector<float> &accumulators() const { return m_acc; }
  const std::vector<uint8_t> &haveSmoke() const { return m_haveSmoke; }
  gvpm_ctx *context() { return m_ctx; }
  double globalScaleVolume;
  size_t totalEmittedVolume = 0;

 private:
  // sum the medium edges of each pixel, normalise, fold into the APA running mean (gvpm.cpp:1054-1069)
  void foldIteration(int it, const gvpm_ray_soa *rays, size_t nRays, float nb) {
    std::vector<float> pix(m_acc.size(), 0.f);
    for (size_t r = 0; r < nRays; ++r) {
      const int x = rays->px[r], y = rays->py[r];
      if (x < 0 || y < 0 || x >= m_w || y >= m_h) continue;
      const size_t p = (size_t)y * m_w + x;
      m_haveSmoke[p] = 1;
      for (int j = 0; j < GVPM_OUT_FLOATS; ++j) pix[p * GVPM_OUT_FLOATS + j] += m_iter[r * GVPM_OUT_FLOATS + j];
    }
    // Spectrum /= Float and Spectrum / Float multiply by the reciprocal (include/mitsuba/core/spectrum.h:415-456): the
    // same two roundings here, so that the running mean is the reference's bit for bit
    const float rnb = 1.0f / nb, rit = 1.0f / (float)it;
    for (size_t i = 0; i < m_acc.size(); ++i) {
      const float fluxVolIter = pix[i] * rnb;
      m_acc[i] = (m_acc[i] * (float)(it - 1) + fluxVolIter) * rit;  // APA estimator
    }
  }
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

// ---- sppm plugin: primal volume passes --------------------------------------------------------------------------
// volTechnique strings of sppm.cpp:208-209 / volume_utils.h:54-92 that the gather serves
enum ESPPMVolumeTechnique { ESppmBRE2D = 0, ESppmBRE3D = 1, ESppmBeam1D = 2, ESppmBeam3DNaive = 3, ESppmBeam3DEGSR = 4,
                            ESppmBeam3DOptimized = 5 };

struct SPPMConfig {   // the SPPMIntegrator members the volume passes read (sppm.cpp:163-241), same names
  int maxDepth = -1, minDepth = 0;
  double alpha = 0.7;
  double initialScaleVolume = 1.0;
  int volTechnique = ESppmBRE3D;
  std::string forceAPA;
  unsigned rngSeed = 0;   // replaces the per-thread Sampler of the 3-D kernels (counter-based hash)
};

// SPPMIntegrator::scaleVolumeAPA, sppm.cpp:255-290 (m_independentScale = false); Real as for the gvpm schedule above
template <typename Real>
inline void scaleVolumeAPA(Real &globalScaleVolume, int it, const SPPMConfig &config) {
  it -= 1;
  const double ratioVolAPA = (double)(((Real)it + (Real)config.alpha) / (Real)(it + 1));
  const int t = config.volTechnique;
  const bool use3D = t == ESppmBRE3D || t == ESppmBeam3DNaive || t == ESppmBeam3DEGSR || t == ESppmBeam3DOptimized ||
                     t == 6 /* ESppmDistance: EVolumeTechniqueHelper::use3DKernel counts EDistance */;
  if (config.forceAPA.empty()) {
    if (use3D) globalScaleVolume = (Real)(globalScaleVolume * std::cbrt(ratioVolAPA));
    else if (t == ESppmBRE2D) globalScaleVolume = (Real)(globalScaleVolume * std::sqrt(ratioVolAPA));
    else globalScaleVolume = (Real)(globalScaleVolume * ratioVolAPA);
  } else if (config.forceAPA == "1D") {
    globalScaleVolume = (Real)(globalScaleVolume * ratioVolAPA);
  } else if (config.forceAPA == "2D") {
    globalScaleVolume = (Real)(globalScaleVolume * std::sqrt(ratioVolAPA));
  } else if (config.forceAPA == "3D") {
    globalScaleVolume = (Real)(globalScaleVolume * std::cbrt(ratioVolAPA));
  } else {
    throw std::runtime_error("No Force APA: " + config.forceAPA);
  }
}

// SPPMIntegrator's constructor (sppm.cpp:163-241): the plugin's XML parameters, defaults and error conditions.
struct SPPMConfigExtra {
  int photonCount = 250000, volumePhotonCount = 250000, rrDepth = 3, maxPasses = -1, dumpIteration = 5, nbCameraSamples = 40;
  bool surfaceRendering = true, volumeRendering = true, convertLong = false, deterministic = false;
  long minCameraDepth = 0;
  int maxCameraDepth = -1;
  double cameraSphere = 1.0;
};
// techniques the sppm plugin parses but this mirror has no pass for (the calls exist in the C ABI: gvpm_gather_vpm /
// gvpm_gather_planes without offsets)
enum { ESppmDistance = 6, ESppmPlane0D = 7 };
inline void loadSPPMConfig(const Properties &props, SPPMConfig &c, SPPMConfigExtra &x) {
  c.initialScaleVolume = props.getFloat("initialScaleVolume", 1.0);
  c.alpha = props.getFloat("alpha", .7);
  x.photonCount = (int)props.getInteger("photonCount", 250000);
  x.volumePhotonCount = (int)props.getInteger("volumePhotonCount", 250000);
  c.maxDepth = (int)props.getInteger("maxDepth", -1);
  c.minDepth = (int)props.getInteger("minDepth", 0);
  x.rrDepth = (int)props.getInteger("rrDepth", 3);
  x.maxPasses = (int)props.getInteger("maxPasses", -1);
  if (c.maxDepth <= 1 && c.maxDepth != -1) throw std::runtime_error("Maximum depth must be set to \"2\" or higher!");
  if (x.maxPasses <= 0 && x.maxPasses != -1)
    throw std::runtime_error("Maximum number of Passes must either be set to \"-1\" or \"1\" or higher!");
  const long maxRenderingTime = props.getInteger("maxRenderingTime", 2147483647L);
  x.dumpIteration = (int)props.getInteger("dumpIteration", 5);
  if (maxRenderingTime != 2147483647L && x.maxPasses != 2147483647)   // as written (:196-198)
    throw std::runtime_error("Max pass and time is incompatible!");
  x.surfaceRendering = props.getBoolean("surfaceRendering", true);
  x.volumeRendering = props.getBoolean("volumeRendering", true);
  // the default "raymarching" is not a name parseVolumeTechnique knows: the scene has to choose one (:208-209)
  const std::string t = props.getString("volTechnique", "raymarching");
  if (t == "distance") c.volTechnique = ESppmDistance;
  else if (t == "bre" || t == "bre3d") c.volTechnique = ESppmBRE3D;
  else if (t == "bre2d") c.volTechnique = ESppmBRE2D;
  else if (t == "beam" || t == "beam1d") c.volTechnique = ESppmBeam1D;
  else if (t == "beam3d" || t == "beam3d_optimized") c.volTechnique = ESppmBeam3DOptimized;
  else if (t == "beam3d_naive") c.volTechnique = ESppmBeam3DNaive;
  else if (t == "beam3d_egsr") c.volTechnique = ESppmBeam3DEGSR;
  else if (t == "plane0d") c.volTechnique = ESppmPlane0D;
  else throw std::runtime_error("Unknow vol technique: " + t);
  x.convertLong = props.getBoolean("convertLong", false);
  x.deterministic = props.getBoolean("deterministic", false);
  x.nbCameraSamples = (int)props.getInteger("nbCameraSamples", 40);
  if (x.surfaceRendering && x.photonCount == 0) throw std::runtime_error("No surface photons and need to render surfaces LT");
  if (x.volumeRendering && x.volumePhotonCount == 0)
    throw std::runtime_error("No volume photons/beams and need to render volume LT");
  x.minCameraDepth = props.getInteger("minCameraDepth", 0);
  if (x.minCameraDepth < 0) throw std::runtime_error("minCamera depth need to be null or positive");
  x.maxCameraDepth = (int)props.getInteger("maxCameraDepth", -1);
  x.cameraSphere = props.getFloat("cameraSphere", 1.0);
  c.forceAPA = props.getString("forceAPA", "");
}

// One Spectrum per pixel: GatherPoint::fluxVol (photonmapper/gatherpoint.h), the APA running mean of sppm.cpp:871,992.
class SPPMVolumeGatherB200 {
 public:
  SPPMVolumeGatherB200(int device, int width, int height, const SPPMConfig &config, const gvpm_medium &medium,
                       float mediumBSphereRadius)
      : globalScaleVolume(config.initialScaleVolume), m_config(config), m_w(width), m_h(height),
        m_bsphereR(mediumBSphereRadius), m_fluxVol((size_t)width * height * 3, 0.f) {
    check(gvpm_ctx_create(device, &m_ctx), "gvpm_ctx_create");
    check(gvpm_set_medium(m_ctx, &medium), "gvpm_set_medium");
    gvpm_config c{};
    c.max_depth = config.maxDepth;
    c.min_depth = config.minDepth;
    c.lighting_mode = GVPM_ALL2MEDIA;
    c.kernel_3d = config.volTechnique != ESppmBRE2D;
    c.sppm_primal = 1;
    c.film_w = width;
    c.film_h = height;
    c.shadow_maxt_scale = 1e-3f;
    c.epsilon = 1e-4f;
    c.rng_seed = config.rngSeed;
    check(gvpm_set_config(m_ctx, &c), "gvpm_set_config");
    check(gvpm_set_occluders(m_ctx, nullptr, 0), "gvpm_set_occluders");   // the primal passes cast no shadow ray
  }
  ~SPPMVolumeGatherB200() { if (m_ctx) gvpm_ctx_destroy(m_ctx); }
  SPPMVolumeGatherB200(const SPPMVolumeGatherB200 &) = delete;
  SPPMVolumeGatherB200 &operator=(const SPPMVolumeGatherB200 &) = delete;

  // breInitSize / beamInitSize = m_smokeAABB.getBSphere().radius * globalScaleVolume * POURCENTAGE_BS (sppm.cpp:770,924)
  float currentRadius() const { return (float)(m_bsphereR * (float)globalScaleVolume * 0.01f); }

  // sppm.cpp:905-1001.  photons: pos, flux = getPower(), parent_pos = pos - getDirection(), depth; rays: one record per
  // (gather point, camera beam): o = beam.p1, d, mint = Epsilon, maxt = distTotal - Epsilon, eye_contrib = beam.weight,
  // edge_id = beam.depth.  shotParticles = proc->getShotParticles() (photonMap->setScaleFactor, :921).
  void volumePhotonPassBRE(int it, const gvpm_photon_soa *photons, size_t nPhotons, const gvpm_ray_soa *rays,
                           size_t nRays, size_t shotParticles) {
    check(gvpm_upload_photons(m_ctx, photons, nPhotons), "gvpm_upload_photons");
    check(gvpm_build_points(m_ctx, currentRadius()), "gvpm_build_points");
    check(gvpm_upload_rays(m_ctx, rays, nRays), "gvpm_upload_rays");
    m_iter.assign(nRays * 3, 0.f);
    check(gvpm_gather_sppm_bre(m_ctx, m_iter.data(), nullptr), "gvpm_gather_sppm_bre");
    foldIteration(it, rays, nRays, (float)shotParticles);   // :976-992
    scaleVolumeAPA(it);                                      // :997
  }

  // sppm.cpp:765-880.  beams: origin, end, flux, depth of the iteration's PhotonBeams; rays as above with
  // edge_len = distTotal.
  void volumePhotonBeamPass(int it, const gvpm_beam_soa *beams, size_t nBeams, const gvpm_ray_soa *rays, size_t nRays,
                            size_t shotParticles) {
    int technique;
    switch (m_config.volTechnique) {
      case ESppmBeam1D: technique = GVPM_BEAM_1D; break;
      case ESppmBeam3DNaive: technique = GVPM_BEAM_3D_NAIVE; break;
      case ESppmBeam3DEGSR: technique = GVPM_BEAM_3D_EGSR; break;
      case ESppmBeam3DOptimized: technique = GVPM_BEAM_3D_OPTIMIZED; break;
      default: throw std::runtime_error("Not supported kernel type");   // beams.h:219
    }
    check(gvpm_upload_beams(m_ctx, beams, nBeams), "gvpm_upload_beams");
    check(gvpm_build_beams(m_ctx, currentRadius()), "gvpm_build_beams");
    check(gvpm_upload_rays(m_ctx, rays, nRays), "gvpm_upload_rays");
    m_iter.assign(nRays * 3, 0.f);
    check(gvpm_gather_sppm_beams(m_ctx, technique, m_iter.data(), nullptr), "gvpm_gather_sppm_beams");
    foldIteration(it, rays, nRays, (float)shotParticles);   // :863-871
    scaleVolumeAPA(it);                                      // :876
  }

  void scaleVolumeAPA(int it) { gvpm_host::scaleVolumeAPA(globalScaleVolume, it, m_config); }
  const std::vector<float> &fluxVol() const { return m_fluxVol; }
  gvpm_ctx *context() { return m_ctx; }
  double globalScaleVolume;

 private:
  // sum the camera beams of each gather point, normalise by the shot particles, APA running mean - for EVERY pixel,
  // "even if there is no photon collected" (sppm.cpp:866-871)
  void foldIteration(int it, const gvpm_ray_soa *rays, size_t nRays, float shot) {
    std::vector<float> pix(m_fluxVol.size(), 0.f);
    for (size_t r = 0; r < nRays; ++r) {
      const int x = rays->px[r], y = rays->py[r];
      if (x < 0 || y < 0 || x >= m_w || y >= m_h) continue;
      const size_t p = ((size_t)y * m_w + x) * 3;
      for (int j = 0; j < 3; ++j) pix[p + j] += m_iter[r * 3 + j];
    }
    const float rshot = 1.0f / shot, rit = 1.0f / (float)it;   // reciprocal multiplies, as Spectrum's operators do
    for (size_t i = 0; i < m_fluxVol.size(); ++i)
      m_fluxVol[i] = (m_fluxVol[i] * (float)(it - 1) + pix[i] * rshot) * rit;
  }
  void check(int rc, const char *what) {
    if (rc != GVPM_OK) throw std::runtime_error(std::string(what) + " failed: " + gvpm_last_error(m_ctx));
  }
  SPPMConfig m_config;
  gvpm_ctx *m_ctx = nullptr;
  int m_w, m_h;
  float m_bsphereR;
  std::vector<float> m_fluxVol, m_iter;
};

}  // namespace gvpm_host
