// functor_stubs.cpp — TEST INFRASTRUCTURE.  Link-time stand-ins for what the reference's BRE shift functor
// (gvpm/shift/shift_volume_photon.cpp, compiled unmodified by oracle/Makefile target functor_ref) references but the pin
// never reaches, plus the one scene query it does reach:
//   * ShapeKDTree::rayIntersect(const Ray &) - the any-hit test of the reconnection's shadow ray
//     (shift_volume_photon.cpp:396-403).  Answered with the reference's own Triangle::rayIntersect
//     (include/mitsuba/core/triangle.h:109-145) over the harness' triangle list and the ray's [mint, maxt]; the kd-tree
//     around it (src/librender/skdtree.cpp) needs the whole renderer.
//   * the offset-path tracer of ShiftGatherPoint::generate (PathEdge::sampleNext, Path::initialize / release): the harness hands the functor gather points that are marked as generated.
//   * the manifold shift (SpecularManifold::det, generateShiftPathME, ShiftME, and the source-path cache of the beam
//     functor: Path::append(Path), PathEdge::clone): out of scope (useManifold = false,
//     SURVEY.md §2); parents that would need it are refused by the harness.
// Nothing here evaluates a radiometric quantity.
#include <array>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <mitsuba/mitsuba.h>
#include <mitsuba/core/stream.h>
#include <mitsuba/core/triangle.h>
#include <mitsuba/render/skdtree.h>
#include <mitsuba/bidir/path.h>
#include <mitsuba/bidir/manifold.h>
#include "gvpm/gvpm_struct.h"
#include "gvpm/shift/operation/shift_ME.h"

MTS_NAMESPACE_BEGIN

static void fn_unreachable(const char *what) {
  std::fprintf(stderr, "gvpm functor ref harness: unexpected call to %s\n", what);
  std::abort();
}

std::vector<std::array<Point, 3>> g_functor_occluders;

bool ShapeKDTree::rayIntersect(const Ray &ray) const {
  for (const auto &t : g_functor_occluders) {
    Float u, v, tt;
    if (Triangle::rayIntersect(t[0], t[1], t[2], ray, u, v, tt) && tt >= ray.mint && tt <= ray.maxt) return true;
  }
  return false;
}

Float VertexClassifier::roughnessThreshold = 0.05f;

// Sampler's serialization (src/librender/sampler.cpp is linked for the base class of the harness' preset sampler)
void Stream::writeULong(uint64_t) { fn_unreachable("Stream::writeULong"); }
uint64_t Stream::readULong() { fn_unreachable("Stream::readULong"); return 0; }
// ... and the stock Photon's (src/librender/photon.cpp, linked for sppm's BRE)
void Stream::writeUShort(unsigned short) { fn_unreachable("Stream::writeUShort"); }
unsigned short Stream::readUShort() { fn_unreachable("Stream::readUShort"); return 0; }

bool PathEdge::sampleNext(const Scene *, Sampler *, const PathVertex *, const Ray &, PathVertex *, ETransportMode, bool, bool) {
  fn_unreachable("PathEdge::sampleNext"); return false;
}
void Path::initialize(const Scene *, Float, ETransportMode, MemoryPool &) { fn_unreachable("Path::initialize"); }
void Path::release(MemoryPool &) { fn_unreachable("Path::release"); }
void Path::append(const Path &, size_t, size_t, bool) { fn_unreachable("Path::append(Path)"); }
PathEdge *PathEdge::clone(MemoryPool &) const { fn_unreachable("PathEdge::clone"); return NULL; }
Float SpecularManifold::det(const Path &, int, int) { fn_unreachable("SpecularManifold::det"); return 0; }
bool generateShiftPathME(const Path &, Path &, size_t, size_t, MemoryPool &, ManifoldPerturbation *, const PathVertex &, Float,
                         Point, Point) { fn_unreachable("generateShiftPathME"); return false; }
bool ShiftME(ShiftRecord &, const Path &, const Path &, size_t, size_t, bool) { fn_unreachable("ShiftME"); return false; }

MTS_NAMESPACE_END
