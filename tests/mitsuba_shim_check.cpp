// compile check of gvpm_mitsuba_shim.hpp against the reference's headers (test infrastructure)
#include "gvpm/gvpm_accel.h"
#include "gvpm/gvpm_beams.h"
#include "gvpm/gvpm_struct.h"
#include "gvpm/shift/shift_cameraPath.h"
#include <mitsuba/render/scene.h>
#include <mitsuba/render/trimesh.h>
#include "gvpm_mitsuba_shim.hpp"
using namespace mitsuba;
// instantiate every entry point
void use(const GPhotonMap &map, const Path *lt, const GatherPoint &gp, const ShiftGatherPoint s[4], const Medium *m,
         const Scene *scene, Sampler *sampler) {
  gvpm_shim::PhotonArrays ph;
  gvpm_shim::flattenPhotonMap(map, ph, false);
  unsigned int added = 0; size_t skip = 0;
  gvpm_shim::appendLightPath(ph, lt, 0, 1000, Point(0.f), 1.f, added, skip, false);
  gvpm_shim::BeamArrays b;
  gvpm_shim::appendLightPathBeams(b, lt, 0, 1000, Point(0.f), 1.f, added, skip, false);
  gvpm_shim::RayArrays r;
  gvpm_shim::appendGatherPoint(r, 0, gp, s, m, 0, -1, sampler);
  gvpm_medium med = gvpm_shim::flattenMedium(m, true);
  std::vector<float> t = gvpm_shim::flattenOccluders(scene);
  (void)med; (void)t; (void)ph.view(); (void)b.view(); (void)r.view();
}
