// gvpm_oracle_planes.cpp — TEST INFRASTRUCTURE (see gvpm_oracle.hpp).  G-Planes 0D gather: the functor
// PlaneGradRadianceQuery::operator() restated in gvpm_oracle.hpp, driven like computeVolumeGradientPlanes
// (gvpm/gvpm.cpp:782-878).  Two drivers:
//   mode 0  brute force: every plane offered to every ray (tree-independent definition of the result);
//   mode 1  the reference's structure: balanced kd-tree over the plane centres (PhotonPlaneMap uses
//           PointKDTree::EBalanced, photonmapper/plane_accel.h:38-41: median split along the axis of largest
//           extent, include/mitsuba/core/kdtree.h:921-1037), bottom-up AABBs of the plane corners
//           (buildHierarchy, plane_accel.h:160-175; PhotonPlane::getAABB, plane_struct.h:68-74) and the
//           explicit-stack DFS of PhotonPlaneBVH::query (plane_accel.h:120-153), which calls the functor on
//           every node whose subtree box is hit by the ray re-based at ray(mint).
// Both select the same planes: a plane's own corners are inside every ancestor box, and the functor repeats
// the exact intersection test.
#include "gvpm_oracle.hpp"

#include <atomic>
#include <chrono>
#include <numeric>
#include <thread>

using namespace gvpm_oracle;

namespace {

template <typename Real> struct PlaneTree {
  using Plane = typename Scene<Real>::Plane;
  struct Node {
    V3<Real> bmin, bmax;
    uint32_t plane;
    int32_t left, right;  // -1 = none
  };
  std::vector<Node> nodes;
  int depth = 0;

  static void corners(const Plane &p, V3<Real> c[4]) {
    c[0] = p.ori;
    c[1] = p.ori + p.w0 * p.length0;
    c[2] = p.ori + p.w1 * p.length1;
    c[3] = (p.ori + p.w1 * p.length1) + p.w0 * p.length0;
  }
  // nodes laid out like the reference's in-place build: node, left subtree, right subtree
  int32_t build(const std::vector<Plane> &pl, std::vector<V3<Real>> &ctr, std::vector<uint32_t> &idx, size_t b, size_t e,
                int d) {
    if (b >= e) return -1;
    depth = std::max(depth, d);
    V3<Real> lo = ctr[idx[b]], hi = lo;
    for (size_t i = b + 1; i < e; ++i)
      for (int a = 0; a < 3; ++a) {
        lo.at(a) = std::min(lo[a], ctr[idx[i]][a]);
        hi.at(a) = std::max(hi[a], ctr[idx[i]][a]);
      }
    int axis = 0;
    for (int a = 1; a < 3; ++a)
      if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
    const size_t mid = b + (e - b) / 2;
    std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e,
                     [&](uint32_t x, uint32_t y) { return ctr[x][axis] < ctr[y][axis]; });
    const int32_t me = (int32_t)nodes.size();
    nodes.push_back(Node());
    nodes[me].plane = idx[mid];
    const int32_t l = build(pl, ctr, idx, b, mid, d + 1);
    const int32_t r = build(pl, ctr, idx, mid + 1, e, d + 1);
    nodes[me].left = l;
    nodes[me].right = r;
    V3<Real> c[4];
    corners(pl[idx[mid]], c);
    V3<Real> bmin = c[0], bmax = c[0];
    auto grow = [&](const V3<Real> &p) {
      for (int a = 0; a < 3; ++a) {
        bmin.at(a) = std::min(bmin[a], p[a]);
        bmax.at(a) = std::max(bmax[a], p[a]);
      }
    };
    for (int i = 1; i < 4; ++i) grow(c[i]);
    if (l >= 0) { grow(nodes[l].bmin); grow(nodes[l].bmax); }
    if (r >= 0) { grow(nodes[r].bmin); grow(nodes[r].bmax); }
    nodes[me].bmin = bmin;
    nodes[me].bmax = bmax;
    return me;
  }
  void build(const std::vector<Plane> &pl) {
    std::vector<V3<Real>> ctr(pl.size());
    for (size_t i = 0; i < pl.size(); ++i)  // PhotonPlane::getCenter, plane_struct.h:64-66
      ctr[i] = (pl[i].ori + (pl[i].w0 * pl[i].length0) * (Real)0.5) + (pl[i].w1 * pl[i].length1) * (Real)0.5;
    std::vector<uint32_t> idx(pl.size());
    std::iota(idx.begin(), idx.end(), 0u);
    nodes.reserve(pl.size());
    build(pl, ctr, idx, 0, pl.size(), 0);
  }
  // AABB::rayIntersect, include/mitsuba/core/aabb.h:310-340
  static bool boxHit(const Node &n, const V3<Real> &o, const V3<Real> &d, const V3<Real> &dRcp, Real &nearT, Real &farT) {
    nearT = -std::numeric_limits<Real>::infinity();
    farT = std::numeric_limits<Real>::infinity();
    for (int a = 0; a < 3; ++a) {
      const Real origin = o[a], minVal = n.bmin[a], maxVal = n.bmax[a];
      if (d[a] == 0) {
        if (origin < minVal || origin > maxVal) return false;
      } else {
        Real t1 = (minVal - origin) * dRcp[a], t2 = (maxVal - origin) * dRcp[a];
        if (t1 > t2) std::swap(t1, t2);
        nearT = std::max(t1, nearT);
        farT = std::min(t2, farT);
        if (!(nearT <= farT)) return false;
      }
    }
    return true;
  }
  template <typename F> void query(const CamRay<Real> &ray, F &&fn) const {
    if (nodes.empty()) return;
    const V3<Real> o = ray.o + ray.d * ray.mint;  // Ray(r(r.mint), r.d, 0, r.maxt - r.mint), plane_accel.h:124
    const Real rmax = ray.maxt - ray.mint;
    const V3<Real> dRcp((Real)1 / ray.d.x, (Real)1 / ray.d.y, (Real)1 / ray.d.z);
    std::vector<int32_t> stack((size_t)depth + 2);
    size_t sp = 0;
    int32_t index = 0;
    for (;;) {
      const Node &n = nodes[index];
      Real mint, maxt;
      const bool hit = boxHit(n, o, ray.d, dRcp, mint, maxt) && !(maxt < 0 || mint > rmax);
      if (hit) {
        fn(n.plane);
        if (n.right >= 0) stack[sp++] = n.right;
        if (n.left >= 0) { index = n.left; continue; }
      }
      if (sp == 0) break;
      index = stack[--sp];
    }
  }
};

template <typename Real>
void planesRange(const gvpm_plane_soa &ps, size_t nPlanes, const gvpm_ray_soa &rays, size_t nRays, const Scene<Real> &sc,
                 int mode, int threads, float *out, uint32_t *counts, std::vector<std::vector<uint32_t>> *nbr,
                 double *build_ms) {
  std::vector<typename Scene<Real>::Plane> planes(nPlanes);
  for (size_t i = 0; i < nPlanes; ++i) planes[i] = Scene<Real>::loadPlane(ps, i);
  PlaneTree<Real> tree;
  auto tb = std::chrono::steady_clock::now();
  if (mode == 1) tree.build(planes);
  if (build_ms) *build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb).count();
  const size_t tile = 64;
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      size_t b = next.fetch_add(tile);
      if (b >= nRays) break;
      size_t e = std::min(nRays, b + tile);
      for (size_t i = b; i < e; ++i) {
        CamRay<Real> ray = loadRay<Real>(rays, i);
        Accum<Real> acc;
        uint32_t nHit = 0;
        auto offer = [&](uint32_t j) {
          if (sc.planeFunctor(ray, planes[j], acc)) {
            ++nHit;
            if (nbr) (*nbr)[i].push_back(j | 0x80000000u);
          }
        };
        if (ray.maxt > ray.mint) {
          if (mode == 1) tree.query(ray, offer);
          else
            for (size_t j = 0; j < nPlanes; ++j) offer((uint32_t)j);
        }
        if (nbr) std::sort((*nbr)[i].begin(), (*nbr)[i].end());
        acc.store(out + GVPM_OUT_FLOATS * i);
        if (counts) { counts[2 * i] = nHit; counts[2 * i + 1] = nHit; }
      }
    }
  };
  if (threads <= 1) worker();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
  }
}
}  // namespace

extern "C" long long gvpm_oracle_planes(const gvpm_plane_soa *planes, size_t nPlanes, const gvpm_ray_soa *rays,
                                        size_t nRays, const gvpm_medium *med, const gvpm_config *cfg, int mode,
                                        int use_double, int threads, float *out, uint32_t *counts,
                                        uint64_t *nbr_offsets, uint32_t *nbr_idx, size_t cap, double *gather_ms,
                                        double *build_ms) {
  std::vector<std::vector<uint32_t>> nbr;
  if (nbr_offsets) nbr.resize(nRays);
  double bms = 0;
  auto t0 = std::chrono::steady_clock::now();
  if (use_double) {
    Scene<double> sc(*med, *cfg, 0.0);
    planesRange<double>(*planes, nPlanes, *rays, nRays, sc, mode, threads, out, counts, nbr_offsets ? &nbr : nullptr, &bms);
  } else {
    Scene<float> sc(*med, *cfg, 0.f);
    planesRange<float>(*planes, nPlanes, *rays, nRays, sc, mode, threads, out, counts, nbr_offsets ? &nbr : nullptr, &bms);
  }
  const double total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (gather_ms) *gather_ms = total_ms - bms;
  if (build_ms) *build_ms = bms;
  long long total = 0;
  if (nbr_offsets) {
    for (size_t i = 0; i < nbr.size(); ++i) {
      nbr_offsets[i] = (uint64_t)total;
      for (uint32_t v : nbr[i]) {
        if ((size_t)total < cap && nbr_idx) nbr_idx[total] = v;
        ++total;
      }
    }
    nbr_offsets[nbr.size()] = (uint64_t)total;
  }
  return total;
}
