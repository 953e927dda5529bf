// gvpm_host_capi.cpp — extern "C" handles on gvpm_host::VolumeGatherB200 so the host-side mirror can be
// driven from tests (ctypes).  A Mitsuba build would include gvpm_host.hpp directly.
#include <cstring>

#include "gvpm_fixture.hpp"
#include "gvpm_host.hpp"

using namespace gvpm_host;

static void set_err(char *err, size_t n, const std::string &m) {
  if (err && n) { strncpy(err, m.c_str(), n - 1); err[n - 1] = 0; }
}

extern "C" {

struct gvpm_host_params {
  int maxDepth, minDepth;
  double alpha, initialScaleVolume;
  int volTechnique, lightingInteractionMode;
  int useMIS, useShiftNull, pathSet, powerHeuristic, use3DKernelReduction;
  char forceAPA[8];
};

static GPMConfig to_cfg(const gvpm_host_params *p) {
  GPMConfig c;
  c.maxDepth = p->maxDepth; c.minDepth = p->minDepth; c.alpha = p->alpha;
  c.initialScaleVolume = p->initialScaleVolume; c.volTechnique = p->volTechnique;
  c.lightingInteractionMode = p->lightingInteractionMode; c.useMIS = p->useMIS;
  c.useShiftNull = p->useShiftNull; c.pathSet = p->pathSet; c.powerHeuristic = p->powerHeuristic;
  c.use3DKernelReduction = p->use3DKernelReduction; c.forceAPA = p->forceAPA;
  return c;
}

// GPMConfig::load on "key=value" lines (what the scene XML's <integrator type="gvpm"> block holds): no device needed
struct gvpm_host_extra {
  int rrDepth, photonCount, volumePhotonCount, maxPasses, dumpIteration, reconstructL1, reconstructL2;
  double reconstructAlpha;
  int useManifold, noMediumShift, convertLong, newShiftBeam, deterministic, nbCameraSamples, minCameraDepth, maxCameraDepth;
  double cameraSphere;
};
int gvpm_host_config_load(const char *text, gvpm_host_params *out, gvpm_host_extra *extra, char *err, size_t errlen) {
  try {
    Properties props;
    std::string all(text ? text : ""), line;
    size_t pos = 0;
    while (pos <= all.size()) {
      const size_t nl = all.find('\n', pos);
      line = all.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
      pos = nl == std::string::npos ? all.size() + 1 : nl + 1;
      const size_t eq = line.find('=');
      if (eq != std::string::npos) props.set(line.substr(0, eq), line.substr(eq + 1));
    }
    GPMConfig c;
    GPMConfigExtra x;
    loadGPMConfig(props, c, x);
    out->maxDepth = c.maxDepth; out->minDepth = c.minDepth; out->alpha = c.alpha;
    out->initialScaleVolume = c.initialScaleVolume; out->volTechnique = c.volTechnique;
    out->lightingInteractionMode = c.lightingInteractionMode; out->useMIS = c.useMIS; out->useShiftNull = c.useShiftNull;
    out->pathSet = c.pathSet; out->powerHeuristic = c.powerHeuristic; out->use3DKernelReduction = c.use3DKernelReduction;
    memset(out->forceAPA, 0, sizeof(out->forceAPA));
    strncpy(out->forceAPA, c.forceAPA.c_str(), sizeof(out->forceAPA) - 1);
    if (extra) {
      extra->rrDepth = x.rrDepth; extra->photonCount = x.photonCount; extra->volumePhotonCount = x.volumePhotonCount;
      extra->maxPasses = x.maxPasses; extra->dumpIteration = x.dumpIteration; extra->reconstructL1 = x.reconstructL1;
      extra->reconstructL2 = x.reconstructL2; extra->reconstructAlpha = x.reconstructAlpha; extra->useManifold = x.useManifold;
      extra->noMediumShift = x.noMediumShift; extra->convertLong = x.convertLong; extra->newShiftBeam = x.newShiftBeam;
      extra->deterministic = x.deterministic; extra->nbCameraSamples = x.nbCameraSamples;
      extra->minCameraDepth = (int)x.minCameraDepth; extra->maxCameraDepth = x.maxCameraDepth; extra->cameraSphere = x.cameraSphere;
    }
    return 0;
  } catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}

// the radius-reduction schedule alone (no device needed)
int gvpm_host_scale_apa(double *scale, int it, const gvpm_host_params *p, char *err, size_t errlen) {
  try { scaleVolumeAPA(*scale, it, to_cfg(p)); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}

// ... as a SINGLE_PRECISION build of the reference evaluates it (Float = float): bit-identical to GPMIntegrator::scaleVolumeAPA
int gvpm_host_scale_apa_f32(float *scale, int it, const gvpm_host_params *p, char *err, size_t errlen) {
  try { scaleVolumeAPA(*scale, it, to_cfg(p)); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}

void *gvpm_host_create(int device, int w, int h, const gvpm_host_params *p, const gvpm_medium *m,
                       float bsphereR, const float *tris, size_t nTris, char *err, size_t errlen) {
  try { return new VolumeGatherB200(device, w, h, to_cfg(p), *m, bsphereR, tris, nTris); }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return nullptr; }
}
void gvpm_host_destroy(void *h) { delete (VolumeGatherB200 *)h; }

int gvpm_host_bre_iteration(void *h, int it, const gvpm_photon_soa *ph, size_t n, const gvpm_ray_soa *rays,
                            size_t nRays, size_t nbPathVolume, char *err, size_t errlen) {
  try { ((VolumeGatherB200 *)h)->computeVolumeGradientPhotonBRE(it, ph, n, rays, nRays, nbPathVolume); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
// the other three gather drivers of the gvpm plugin (gvpm.cpp:880-986, 782-878, 1081-1203)
int gvpm_host_beams_iteration(void *h, int it, const gvpm_beam_soa *beams, size_t n, const gvpm_ray_soa *rays, size_t nRays,
                              size_t nbPathBeams, char *err, size_t errlen) {
  try { ((VolumeGatherB200 *)h)->computeVolumeGradientBeams(it, beams, n, rays, nRays, nbPathBeams); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
int gvpm_host_planes_iteration(void *h, int it, const gvpm_plane_soa *planes, size_t n, const gvpm_ray_soa *rays,
                               size_t nRays, size_t nbPathBeams, char *err, size_t errlen) {
  try { ((VolumeGatherB200 *)h)->computeVolumeGradientPlanes(it, planes, n, rays, nRays, nbPathBeams); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
int gvpm_host_vpm_iteration(void *h, int it, const gvpm_photon_soa *ph, size_t n, const gvpm_ray_soa *rays, size_t nRays,
                            const gvpm_vpm_sample_soa *samples, size_t nSamples, int nbCameraSamples, float maxRadius,
                            size_t nbPathVolume, uint32_t *mvol, char *err, size_t errlen) {
  try {
    ((VolumeGatherB200 *)h)->computeVolumeGradientPhoton(it, ph, n, rays, nRays, samples, nSamples, nbCameraSamples, maxRadius,
                                                        nbPathVolume, mvol);
    return 0;
  } catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
// VPM is not an APA estimator: accumulators / totalEmittedVolume (gvpm.cpp:487-491); out = [h][w][27]
int gvpm_host_normalized_accumulators(void *h, float *out, size_t n) {
  const std::vector<float> a = ((VolumeGatherB200 *)h)->normalizedAccumulators();
  if (n != a.size()) return -1;
  memcpy(out, a.data(), n * sizeof(float));
  return 0;
}
int gvpm_host_gradient(void *h, float *thr, float *gx, float *gy, int useAbs, char *err, size_t errlen) {
  try { ((VolumeGatherB200 *)h)->computeGradient(thr, gx, gy, useAbs != 0); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
int gvpm_host_reconstruct(void *h, const char *preset, float alpha, int useAbs, const float *direct, float *thr, float *gx,
                          float *gy, float *rec, char *err, size_t errlen) {
  try { ((VolumeGatherB200 *)h)->reconstruct(preset, alpha, useAbs != 0, direct, thr, gx, gy, rec); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
double gvpm_host_scale(void *h) { return ((VolumeGatherB200 *)h)->globalScaleVolume; }
float gvpm_host_radius(void *h) { return ((VolumeGatherB200 *)h)->currentRadius(); }
const float *gvpm_host_accumulators(void *h) { return ((VolumeGatherB200 *)h)->accumulators().data(); }

// ---- sppm mirror -----------------------------------------------------------------------------------------------------
struct gvpm_host_sppm_params {
  int maxDepth, minDepth;
  double alpha, initialScaleVolume;
  int volTechnique;
  unsigned rngSeed;
  char forceAPA[8];
};
static SPPMConfig to_sppm(const gvpm_host_sppm_params *p) {
  SPPMConfig c;
  c.maxDepth = p->maxDepth; c.minDepth = p->minDepth; c.alpha = p->alpha; c.initialScaleVolume = p->initialScaleVolume;
  c.volTechnique = p->volTechnique; c.rngSeed = p->rngSeed; c.forceAPA = p->forceAPA;
  return c;
}
struct gvpm_host_sppm_extra {
  int photonCount, volumePhotonCount, rrDepth, maxPasses, dumpIteration, nbCameraSamples, surfaceRendering, volumeRendering,
      convertLong, deterministic, minCameraDepth, maxCameraDepth;
  double cameraSphere;
};
static Properties parse_props(const char *text) {
  Properties props;
  std::string all(text ? text : "");
  size_t pos = 0;
  while (pos <= all.size()) {
    const size_t nl = all.find('\n', pos);
    const std::string line = all.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
    pos = nl == std::string::npos ? all.size() + 1 : nl + 1;
    const size_t eq = line.find('=');
    if (eq != std::string::npos) props.set(line.substr(0, eq), line.substr(eq + 1));
  }
  return props;
}
int gvpm_host_sppm_config_load(const char *text, gvpm_host_sppm_params *out, gvpm_host_sppm_extra *extra, char *err,
                               size_t errlen) {
  try {
    SPPMConfig c;
    SPPMConfigExtra x;
    loadSPPMConfig(parse_props(text), c, x);
    out->maxDepth = c.maxDepth; out->minDepth = c.minDepth; out->alpha = c.alpha;
    out->initialScaleVolume = c.initialScaleVolume; out->volTechnique = c.volTechnique; out->rngSeed = c.rngSeed;
    memset(out->forceAPA, 0, sizeof(out->forceAPA));
    strncpy(out->forceAPA, c.forceAPA.c_str(), sizeof(out->forceAPA) - 1);
    if (extra) {
      extra->photonCount = x.photonCount; extra->volumePhotonCount = x.volumePhotonCount; extra->rrDepth = x.rrDepth;
      extra->maxPasses = x.maxPasses; extra->dumpIteration = x.dumpIteration; extra->nbCameraSamples = x.nbCameraSamples;
      extra->surfaceRendering = x.surfaceRendering; extra->volumeRendering = x.volumeRendering;
      extra->convertLong = x.convertLong; extra->deterministic = x.deterministic;
      extra->minCameraDepth = (int)x.minCameraDepth; extra->maxCameraDepth = x.maxCameraDepth;
      extra->cameraSphere = x.cameraSphere;
    }
    return 0;
  } catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
int gvpm_host_sppm_scale_apa(double *scale, int it, const gvpm_host_sppm_params *p, char *err, size_t errlen) {
  try { scaleVolumeAPA(*scale, it, to_sppm(p)); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}

// ... as a SINGLE_PRECISION build of the reference evaluates it
int gvpm_host_sppm_scale_apa_f32(float *scale, int it, const gvpm_host_sppm_params *p, char *err, size_t errlen) {
  try { scaleVolumeAPA(*scale, it, to_sppm(p)); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
void *gvpm_host_sppm_create(int device, int w, int h, const gvpm_host_sppm_params *p, const gvpm_medium *m,
                            float bsphereR, char *err, size_t errlen) {
  try { return new SPPMVolumeGatherB200(device, w, h, to_sppm(p), *m, bsphereR); }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return nullptr; }
}
void gvpm_host_sppm_destroy(void *h) { delete (SPPMVolumeGatherB200 *)h; }
int gvpm_host_sppm_bre_pass(void *h, int it, const gvpm_photon_soa *ph, size_t n, const gvpm_ray_soa *rays, size_t nRays,
                            size_t shotParticles, char *err, size_t errlen) {
  try { ((SPPMVolumeGatherB200 *)h)->volumePhotonPassBRE(it, ph, n, rays, nRays, shotParticles); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
int gvpm_host_sppm_beam_pass(void *h, int it, const gvpm_beam_soa *beams, size_t n, const gvpm_ray_soa *rays,
                             size_t nRays, size_t shotParticles, char *err, size_t errlen) {
  try { ((SPPMVolumeGatherB200 *)h)->volumePhotonBeamPass(it, beams, n, rays, nRays, shotParticles); return 0; }
  catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}
double gvpm_host_sppm_scale(void *h) { return ((SPPMVolumeGatherB200 *)h)->globalScaleVolume; }
float gvpm_host_sppm_radius(void *h) { return ((SPPMVolumeGatherB200 *)h)->currentRadius(); }
const float *gvpm_host_sppm_flux_vol(void *h) { return ((SPPMVolumeGatherB200 *)h)->fluxVol().data(); }

// the fixture writer a Mitsuba-side dump hook uses (gvpm_fixture.hpp), reachable from the tests
int gvpm_host_write_bre_fixture(const char *path, const gvpm_medium *m, const gvpm_config *c, float radius,
                                const float *tris, size_t nTris, const gvpm_photon_soa *ph, size_t nPhotons,
                                const gvpm_ray_soa *rays, size_t nRays, const float *expectedOut,
                                const uint64_t *nbrOffsets, const uint32_t *nbrIdx, const char *producer, char *err,
                                size_t errlen) {
  try {
    gvpm_fixture::write_bre_fixture(path, *m, *c, radius, tris, nTris, *ph, nPhotons, *rays, nRays, expectedOut,
                                    nbrOffsets, nbrIdx, producer);
    return 0;
  } catch (const std::exception &e) { set_err(err, errlen, e.what()); return -1; }
}

}  // extern "C"
