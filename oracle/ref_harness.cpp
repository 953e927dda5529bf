// ref_harness.cpp — TEST INFRASTRUCTURE.  extern "C" marshalling around the REFERENCE'S OWN code, compiled from
// /root/reference where it lies (oracle/Makefile, target _ref/libgvpm_ref.so; no reference source is copied).
// It exists to pin oracle/gvpm_oracle.hpp: every function here only converts flat arrays to the reference's
// types, calls the reference, and flattens what comes back.  The functors passed to the reference's query
// templates record their arguments and compute nothing.
//
// What is pinned (reference file:line -> harness entry):
//   PointKDTree::build ESlidingMidpoint / EBalanced, include/mitsuba/core/kdtree.h:326-378,921-1037 -> ref_kd_layout
//   PointKDTree::executeQuery, kdtree.h:675-731                                             -> ref_range_visits
//   GPhotonMap::build + GradientBeamRadianceEstimator ctor/buildHierarchy/query,
//     gvpm/gvpm_accel.h:201-203,268-312, gvpm/gvpm_accel.cpp:10-55 (incl. AABB::rayIntersect aabb.h:310-340) -> ref_bre_visits
//   SubBeamBVH ctor/buildHierarchy/query, photonmapper/beams_accel.h:90-243                  -> ref_subbeam_visits
//   PhotonPlaneBVH ctor/buildHierarchy/query, photonmapper/plane_accel.h:93-185              -> ref_plane_visits
//   cylinderIntersection, photonmapper/beams_3d_intersections.h:77-140                       -> ref_cylinder
//   PhotonBeam::rayIntersectInternal1D, photonmapper/beams_struct.h:250-311                  -> ref_beam1d
//   PhotonPlane::intersectPlane0D, photonmapper/plane_struct.h:104-135                       -> ref_plane0d
//   Triangle::rayIntersect, include/mitsuba/core/triangle.h:109-145                          -> ref_triangle
//   coordinateSystem / coordinateSystemCoherent, src/libcore/util.cpp:592-609                -> ref_coordsys
//   solveQuadraticDouble, src/libcore/util.cpp:487-525                                       -> ref_quadratic
// The shift functors themselves (gvpm/shift/*.cpp), which need Path / Medium / BSDF / Scene objects, are driven by a second
// harness: ref_functor.cpp -> _ref/libgvpm_functor_ref.so.
#include "gvpm/gvpm_accel.h"
#include "beams_accel.h"
#include "plane_accel.h"
#include "beams_3d_intersections.h"
#include <mitsuba/core/triangle.h>

using namespace mitsuba;

namespace {

inline Point P3(const float *p) { return Point(p[0], p[1], p[2]); }
inline Vector V3f(const float *p) { return Vector(p[0], p[1], p[2]); }

// exposes the protected kd-tree of GPhotonMap so that photons can be appended without a bidir Path
// (GPhotonMap::tryAppend reads the position from Path vertices, gvpm_accel.h:119-199)
class HarnessMap : public GPhotonMap {
public:
  explicit HarnessMap(size_t n) : GPhotonMap(n, false, Point(0.f), 0.f) {}
  void add(const Point &p, uint32_t i) {
    GPhotonNodeKD node;
    node.setPosition(p);
    GPhotonNodeData d;
    d.pathID = i;        // payload = caller's photon index
    d.vertexId = 2;
    node.setData(d);
    m_kdtree.push_back(node);
  }
  const PhotonTree &tree() const { return m_kdtree; }
};

struct CollectBRE {
  std::vector<uint32_t> *idx;
  std::vector<float> *tdisk;
  Float maxt = 0;
  void newRayBase(const Ray &r, const Medium *) { maxt = r.maxt; }
  void operator()(const GPhotonNodeKD &p, Float, Float) {
    idx->push_back(p.getData().pathID);
    tdisk->push_back((float)maxt);
  }
};

template <typename A, typename B> long long flushCSR(const std::vector<std::vector<A>> &lists, uint64_t *offsets,
                                                     B *out, size_t cap) {
  long long total = 0;
  for (size_t i = 0; i < lists.size(); ++i) {
    offsets[i] = (uint64_t)total;
    for (const A &v : lists[i]) {
      if (out && (size_t)total < cap) out[total] = (B)v;
      ++total;
    }
  }
  offsets[lists.size()] = (uint64_t)total;
  return total;
}

}  // namespace

extern "C" {

int ref_float_bytes(void) { return (int)sizeof(Float); }

// kd layout after PointKDTree::build: for tree slot i, the caller's point index, the stored right-child index,
// leaf flag and split axis.  heuristic: 0 = EBalanced, 3 = ESlidingMidpoint (kdtree.h enum order is read from the type).
int ref_kd_layout(const float *pos, size_t n, int sliding, uint32_t *orig, uint32_t *right, uint8_t *leaf,
                  uint8_t *axis) {
  typedef PointKDTree<SimpleKDNode<Point, uint32_t>> Tree;
  Tree t(0, sliding ? Tree::ESlidingMidpoint : Tree::EBalanced);
  t.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    SimpleKDNode<Point, uint32_t> node;
    node.setPosition(P3(pos + 3 * i));
    node.setData((uint32_t)i);
    t.push_back(node);
  }
  t.build(true);
  for (size_t i = 0; i < n; ++i) {
    orig[i] = t[i].getData();
    leaf[i] = t[i].isLeaf() ? 1 : 0;
    right[i] = t[i].isLeaf() ? 0 : (uint32_t)t[i].getRightIndex(i);
    axis[i] = t[i].isLeaf() ? 0 : (uint8_t)t[i].getAxis();
  }
  return (int)t.getDepth();
}

// PointKDTree::executeQuery(p, radius, functor) for m query points: visit order of the caller's indices (CSR).
long long ref_range_visits(const float *pos, size_t n, const float *q, const float *radius, size_t m,
                           uint64_t *offsets, uint32_t *idx, size_t cap) {
  ref<HarnessMap> map = new HarnessMap(n);
  for (size_t i = 0; i < n; ++i) map->add(P3(pos + 3 * i), (uint32_t)i);
  map->build(true);
  std::vector<std::vector<uint32_t>> lists(m);
  struct Q {
    std::vector<uint32_t> *l;
    void operator()(const GPhotonNodeKD &p) { l->push_back(p.getData().pathID); }
  };
  for (size_t j = 0; j < m; ++j) {
    Q fn{&lists[j]};
    map->evaluate(fn, P3(q + 3 * j), radius[j]);
  }
  return flushCSR(lists, offsets, idx, cap);
}

// GradientBeamRadianceEstimator over the given photons, then bre->query for every ray: the sequence of functor
// calls (photon index, baseRay.maxt = diskDistance) in the reference's traversal order.
long long ref_bre_visits(const float *pos, size_t n, float radius, const float *ray_o, const float *ray_d,
                         const float *ray_mint, const float *ray_maxt, size_t n_rays, uint64_t *offsets,
                         uint32_t *idx, float *tdisk, size_t cap, int *depth) {
  ref<HarnessMap> map = new HarnessMap(n);
  for (size_t i = 0; i < n; ++i) map->add(P3(pos + 3 * i), (uint32_t)i);
  map->build(true);                                                        // gvpm.cpp:453
  ref<GradientBeamRadianceEstimator> bre = new GradientBeamRadianceEstimator(map.get(), radius);  // gvpm.cpp:994
  if (depth) *depth = (int)map->getDepth();
  std::vector<std::vector<uint32_t>> li(n_rays);
  std::vector<std::vector<float>> lt(n_rays);
  for (size_t r = 0; r < n_rays; ++r) {
    Ray ray(P3(ray_o + 3 * r), V3f(ray_d + 3 * r), ray_mint[r], ray_maxt[r], 0.f);
    CollectBRE c;
    c.idx = &li[r];
    c.tdisk = &lt[r];
    bre->query(ray, nullptr, c, 0.5f);
  }
  flushCSR(lt, offsets, tdisk, cap);
  return flushCSR(li, offsets, idx, cap);
}

// SubBeamBVH<PhotonBeam> over beams (origin, end, radius): per ray the sequence of functor calls
// (beam index, t1, t2).  ray = (o, d, mint, maxt) as baseCameraRay.
long long ref_subbeam_visits(const float *origin, const float *end, size_t n, float radius, const float *ray_o,
                             const float *ray_d, const float *ray_mint, const float *ray_maxt, size_t n_rays,
                             uint64_t *offsets, uint32_t *idx, float *t1, float *t2, size_t cap) {
  std::vector<std::pair<int, PhotonBeam>> beams;
  beams.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    PhotonBeam b(P3(origin + 3 * i), nullptr, Spectrum(1.f), 1, radius);
    b.setEndPoint(P3(end + 3 * i));
    beams.push_back(std::make_pair((int)i, b));
  }
  ref<SubBeamBVH<PhotonBeam>> bvh = new SubBeamBVH<PhotonBeam>(beams);
  struct Q {
    Ray baseCameraRay;
    const PhotonBeam *first;
    size_t stride;
    std::vector<uint32_t> *li;
    std::vector<float> *l1, *l2;
    void operator()(const PhotonBeam *b, Float a, Float c) {
      li->push_back((uint32_t)(((const char *)b - (const char *)first) / stride));
      l1->push_back((float)a);
      l2->push_back((float)c);
    }
  };
  std::vector<std::vector<uint32_t>> li(n_rays);
  std::vector<std::vector<float>> l1(n_rays), l2(n_rays);
  for (size_t r = 0; r < n_rays; ++r) {
    Q q;
    q.baseCameraRay = Ray(P3(ray_o + 3 * r), V3f(ray_d + 3 * r), ray_mint[r], ray_maxt[r], 0.f);
    q.first = &beams[0].second;
    q.stride = sizeof(std::pair<int, PhotonBeam>);
    q.li = &li[r];
    q.l1 = &l1[r];
    q.l2 = &l2[r];
    bvh->query(q);
  }
  flushCSR(l1, offsets, t1, cap);
  flushCSR(l2, offsets, t2, cap);
  return flushCSR(li, offsets, idx, cap);
}

// PhotonPlaneBVH<PhotonPlane>: per ray the sequence of planes handed to the functor.
long long ref_plane_visits(const float *ori, const float *w0, const float *len0, const float *w1, const float *len1,
                           size_t n, const float *ray_o, const float *ray_d, const float *ray_mint,
                           const float *ray_maxt, size_t n_rays, uint64_t *offsets, uint32_t *idx, size_t cap) {
  std::vector<PhotonPlane> planes;
  planes.reserve(n);
  for (size_t i = 0; i < n; ++i)
    planes.push_back(PhotonPlane(P3(ori + 3 * i), V3f(w0 + 3 * i), len0[i], V3f(w1 + 3 * i), len1[i], nullptr,
                                 Spectrum(1.f), 1));
  ref<PhotonPlaneBVH<PhotonPlane>> bvh = new PhotonPlaneBVH<PhotonPlane>(planes);
  struct Q {
    Ray baseCameraRay;
    const PhotonPlane *first;
    std::vector<uint32_t> *li;
    void operator()(const PhotonPlane *p) { li->push_back((uint32_t)(p - first)); }
  };
  std::vector<std::vector<uint32_t>> li(n_rays);
  for (size_t r = 0; r < n_rays; ++r) {
    Q q;
    q.baseCameraRay = Ray(P3(ray_o + 3 * r), V3f(ray_d + 3 * r), ray_mint[r], ray_maxt[r], 0.f);
    q.first = planes.data();
    q.li = &li[r];
    bvh->query(q);
  }
  return flushCSR(li, offsets, idx, cap);
}

// cylinderIntersection(rCylinder = (co, cd, [0, cmaxt]), view = (vo, vd, [0, vmaxt]), radius) for m pairs.
void ref_cylinder(const float *co, const float *cd, const float *cmaxt, const float *vo, const float *vd,
                  const float *vmaxt, const float *radius, size_t m, uint8_t *hit, double *tNear, double *tFar) {
  for (size_t i = 0; i < m; ++i) {
    Ray c(P3(co + 3 * i), V3f(cd + 3 * i), 0.f, cmaxt[i], 0.f);
    Ray v(P3(vo + 3 * i), V3f(vd + 3 * i), 0.f, vmaxt[i], 0.f);
    double a = 0, b = 0;
    hit[i] = cylinderIntersection(c, v, radius[i], a, b) ? 1 : 0;
    tNear[i] = a;
    tFar[i] = b;
  }
}

// PhotonBeam(origin -> end, radius).rayIntersectInternal1D(radius, ray, tmin, tmax, u, v, w, sinTheta)
void ref_beam1d(const float *origin, const float *end, const float *radius, const float *ro, const float *rd,
                const float *rmint, const float *rmaxt, const float *tmin, const float *tmax, size_t m,
                uint8_t *hit, float *uvws /* [m*4] */) {
  for (size_t i = 0; i < m; ++i) {
    PhotonBeam b(P3(origin + 3 * i), nullptr, Spectrum(1.f), 1, radius[i]);
    b.setEndPoint(P3(end + 3 * i));
    Ray r(P3(ro + 3 * i), V3f(rd + 3 * i), rmint[i], rmaxt[i], 0.f);
    Float u = 0, v = 0, w = 0, s = 0;
    hit[i] = b.rayIntersectInternal1D(radius[i], r, tmin[i], tmax[i], u, v, w, s) ? 1 : 0;
    uvws[4 * i] = (float)u; uvws[4 * i + 1] = (float)v; uvws[4 * i + 2] = (float)w; uvws[4 * i + 3] = (float)s;
  }
}

// PhotonPlane::intersectPlane0D(ray, tCam, t0, t1, invDet)
void ref_plane0d(const float *ori, const float *w0, const float *len0, const float *w1, const float *len1,
                 const float *ro, const float *rd, const float *rmint, const float *rmaxt, size_t m, uint8_t *hit,
                 float *out /* [m*4] tCam t0 t1 invDet */) {
  for (size_t i = 0; i < m; ++i) {
    PhotonPlane p(P3(ori + 3 * i), V3f(w0 + 3 * i), len0[i], V3f(w1 + 3 * i), len1[i], nullptr, Spectrum(1.f), 1);
    Ray r(P3(ro + 3 * i), V3f(rd + 3 * i), rmint[i], rmaxt[i], 0.f);
    Float a = 0, b = 0, c = 0, d = 0;
    hit[i] = p.intersectPlane0D(r, a, b, c, d) ? 1 : 0;
    out[4 * i] = (float)a; out[4 * i + 1] = (float)b; out[4 * i + 2] = (float)c; out[4 * i + 3] = (float)d;
  }
}

// Triangle::rayIntersect(p0, p1, p2, ray, u, v, t); the interval test is the caller's (shape code), so it is
// reported raw.
void ref_triangle(const float *tri /* [m*9] */, const float *ro, const float *rd, size_t m, uint8_t *hit,
                  float *uvt /* [m*3] */) {
  for (size_t i = 0; i < m; ++i) {
    Ray r(P3(ro + 3 * i), V3f(rd + 3 * i), 0.f);
    Float u = 0, v = 0, t = 0;
    hit[i] = Triangle::rayIntersect(P3(tri + 9 * i), P3(tri + 9 * i + 3), P3(tri + 9 * i + 6), r, u, v, t) ? 1 : 0;
    uvt[3 * i] = (float)u; uvt[3 * i + 1] = (float)v; uvt[3 * i + 2] = (float)t;
  }
}

void ref_coordsys(const float *a, size_t m, int coherent, float *b, float *c) {
  for (size_t i = 0; i < m; ++i) {
    Vector s, t;
    if (coherent) coordinateSystemCoherent(V3f(a + 3 * i), s, t); else coordinateSystem(V3f(a + 3 * i), s, t);
    b[3 * i] = s.x; b[3 * i + 1] = s.y; b[3 * i + 2] = s.z;
    c[3 * i] = t.x; c[3 * i + 1] = t.y; c[3 * i + 2] = t.z;
  }
}

void ref_quadratic(const double *abc, size_t m, uint8_t *ok, double *x0, double *x1) {
  for (size_t i = 0; i < m; ++i) {
    double a = 0, b = 0;
    ok[i] = solveQuadraticDouble(abc[3 * i], abc[3 * i + 1], abc[3 * i + 2], a, b) ? 1 : 0;
    x0[i] = a;
    x1[i] = b;
  }
}

}  // extern "C"
