// gvpm_oracle.cpp — TEST INFRASTRUCTURE (see gvpm_oracle.hpp).  extern "C" entry points for
// ctypes: build the reference-shaped kd/AABB hierarchy, run the BRE gather over a ray range
// with worker threads that pull 1024-ray tiles dynamically (the reference's BlockScheduler:
// nCores threads pulling 32x32 blocks under a mutex, utilities/block_sched.h:87-113).
#include "gvpm_oracle.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace gvpm_oracle;

namespace {

struct TreeHandle {
  bool dbl;
  BreTree<float> tf;
  BreTree<double> td;
  size_t n;
  double build_ms;
};

template <typename Real>
void gatherRange(const BreTree<Real> *tree, const gvpm_photon_soa &ph, size_t n, const gvpm_ray_soa &rays,
                 const Scene<Real> &sc, size_t begin, size_t end, int threads, float *out,
                 uint32_t *counts, std::vector<std::vector<uint32_t>> *nbr) {
  const size_t tile = 1024;
  std::atomic<size_t> next(begin);
  auto worker = [&]() {
    for (;;) {
      size_t b = next.fetch_add(tile);
      if (b >= end) break;
      size_t e = std::min(end, b + tile);
      for (size_t i = b; i < e; ++i) {
        CamRay<Real> ray = loadRay<Real>(rays, i);
        Accum<Real> acc;
        uint32_t nGeom = 0, nContrib = 0;
        std::vector<uint32_t> *mine = nbr ? &(*nbr)[i - begin] : nullptr;
        auto visit = [&](uint32_t orig, Real diskDistance) {
          Photon<Real> p = loadPhoton<Real>(ph, orig);
          int r = sc.breFunctor(ray, p, diskDistance, acc);
          if (r >= 1) {
            ++nGeom;
            if (r == 2) ++nContrib;
            if (mine) mine->push_back(orig | (r == 2 ? 0x80000000u : 0u));
          }
        };
        if (tree) {
          tree->query(sc, ray, [&](uint32_t nodeIdx, Real dd) { visit(tree->nodes[nodeIdx].orig, dd); });
        } else {  // brute force over all photons: the tree-independent set S
          for (size_t j = 0; j < n; ++j) {
            Real dd;
            if (sc.diskTest(ray, V3<Real>(ph.pos + 3 * j), dd)) visit((uint32_t)j, dd);
          }
        }
        acc.store(out + GVPM_OUT_FLOATS * (i - begin));
        if (counts) {
          counts[2 * (i - begin)] = nGeom;
          counts[2 * (i - begin) + 1] = nContrib;
        }
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
}

}  // namespace

extern "C" {

int gvpm_oracle_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

// PointKDTree::build + GradientBeamRadianceEstimator ctor (kdtree.h:326-378, gvpm_accel.cpp:10-55)
void *gvpm_oracle_tree_build(const gvpm_photon_soa *ph, size_t n, float radius, int use_double) {
  auto t0 = std::chrono::steady_clock::now();
  TreeHandle *h = new TreeHandle();
  h->dbl = use_double != 0;
  h->n = n;
  if (h->dbl) h->td.build(*ph, n, (double)radius); else h->tf.build(*ph, n, radius);
  h->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return h;
}
double gvpm_oracle_tree_build_ms(void *h) { return ((TreeHandle *)h)->build_ms; }
size_t gvpm_oracle_tree_depth(void *h) {
  TreeHandle *t = (TreeHandle *)h;
  return t->dbl ? t->td.depth : t->tf.depth;
}
void gvpm_oracle_tree_free(void *h) { delete (TreeHandle *)h; }

// computeVolumeGradientPhotonBRE's gather loop (gvpm.cpp:999-1052) over rays [begin, end).
// tree == NULL: brute force.  out: [(end-begin)*27]; counts: [(end-begin)*2] or NULL.
// nbr_offsets/nbr_idx: optional CSR dump of the geometric neighbour set (bit 31 = contributes);
// returns the number of entries needed (>= 0) or a negative value on error.
long long gvpm_oracle_bre(void *tree, const gvpm_photon_soa *ph, size_t n, const gvpm_ray_soa *rays,
                          size_t begin, size_t end, const gvpm_medium *med, const gvpm_config *cfg,
                          const float *tri, size_t n_tri, float radius, int use_double, int threads,
                          float *out, uint32_t *counts, uint64_t *nbr_offsets, uint32_t *nbr_idx,
                          size_t cap, double *gather_ms) {
  if (end < begin) return -1;
  TreeHandle *h = (TreeHandle *)tree;
  if (h && (h->dbl != (use_double != 0) || h->n != n)) return -2;
  std::vector<std::vector<uint32_t>> nbr;
  if (nbr_offsets) nbr.resize(end - begin);
  auto t0 = std::chrono::steady_clock::now();
  if (use_double) {
    Scene<double> sc(*med, *cfg, (double)radius);
    sc.occ.set(tri, n_tri);
    gatherRange<double>(h ? &h->td : nullptr, *ph, n, *rays, sc, begin, end, threads, out, counts,
                        nbr_offsets ? &nbr : nullptr);
  } else {
    Scene<float> sc(*med, *cfg, radius);
    sc.occ.set(tri, n_tri);
    gatherRange<float>(h ? &h->tf : nullptr, *ph, n, *rays, sc, begin, end, threads, out, counts,
                       nbr_offsets ? &nbr : nullptr);
  }
  if (gather_ms)
    *gather_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  long long total = 0;
  if (nbr_offsets) {
    for (size_t i = 0; i < nbr.size(); ++i) {
      nbr_offsets[i] = (uint64_t)total;
      std::sort(nbr[i].begin(), nbr[i].end(),
                [](uint32_t a, uint32_t b) { return (a & 0x7fffffffu) < (b & 0x7fffffffu); });
      for (uint32_t v : nbr[i]) {
        if ((size_t)total < cap && nbr_idx) nbr_idx[total] = v;
        ++total;
      }
    }
    nbr_offsets[nbr.size()] = (uint64_t)total;
  }
  return total;
}

}  // extern "C"

// ---- sppm primal BRE (sppm.cpp:926-981 + bre.cpp:167-259) over rays [0, nRays).  tree == NULL: brute force.
// out: [nRays*3]; counts: [nRays*2] {geometric, contributing}; optional CSR dump like gvpm_oracle_bre.
extern "C" long long gvpm_oracle_sppm_bre(void *tree, const gvpm_photon_soa *ph, size_t n, const gvpm_ray_soa *rays,
                                          size_t nRays, const gvpm_medium *med, const gvpm_config *cfg, float radius,
                                          int threads, float *out, uint32_t *counts, uint64_t *nbr_offsets,
                                          uint32_t *nbr_idx, size_t cap, double *gather_ms) {
  TreeHandle *h = (TreeHandle *)tree;
  if (h && (h->dbl || h->n != n)) return -2;
  std::vector<std::vector<uint32_t>> nbr;
  if (nbr_offsets) nbr.resize(nRays);
  auto t0 = std::chrono::steady_clock::now();
  Scene<float> sc(*med, *cfg, radius);
  const size_t tile = 256;
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      size_t b = next.fetch_add(tile);
      if (b >= nRays) break;
      size_t e = std::min(nRays, b + tile);
      for (size_t i = b; i < e; ++i) {
        CamRay<float> ray = loadRay<float>(*rays, i);
        V3<float> result;
        uint32_t nGeom = 0, nContrib = 0;
        auto visit = [&](uint32_t orig) {
          Photon<float> p = loadPhoton<float>(*ph, orig);
          int r = sc.sppmBreFunctor(ray, p, orig, result);
          if (r >= 1) {
            ++nGeom;
            if (r == 2) ++nContrib;
            if (nbr_offsets) nbr[i].push_back(orig | (r == 2 ? 0x80000000u : 0u));
          }
        };
        if (h) {
          // const Ray ray(r(r.mint), r.d, 0, r.maxt - r.mint, r.time), bre.cpp:169: same stack DFS as the gvpm query
          CamRay<float> rb = ray;
          rb.o = ray.o + ray.mint * ray.d;
          rb.maxt = ray.maxt - ray.mint;
          rb.mint = 0;
          h->tf.query(sc, rb, [&](uint32_t nodeIdx, float) { visit(h->tf.nodes[nodeIdx].orig); });
        } else {
          for (size_t j = 0; j < n; ++j) visit((uint32_t)j);
        }
        out[3 * i] = result.x; out[3 * i + 1] = result.y; out[3 * i + 2] = result.z;
        if (counts) { counts[2 * i] = nGeom; counts[2 * i + 1] = nContrib; }
      }
    }
  };
  if (threads <= 1) worker();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
  }
  if (gather_ms)
    *gather_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  long long total = 0;
  if (nbr_offsets) {
    for (size_t i = 0; i < nbr.size(); ++i) {
      nbr_offsets[i] = (uint64_t)total;
      std::sort(nbr[i].begin(), nbr[i].end(),
                [](uint32_t a, uint32_t b) { return (a & 0x7fffffffu) < (b & 0x7fffffffu); });
      for (uint32_t v : nbr[i]) {
        if ((size_t)total < cap && nbr_idx) nbr_idx[total] = v;
        ++total;
      }
    }
    nbr_offsets[nbr.size()] = (uint64_t)total;
  }
  return total;
}

// ---- G-VPM: computeVolumeGradientPhoton's per-sample range queries (gvpm.cpp:1141-1185) -------------
namespace {
template <typename Real>
void vpmRange(const BreTree<Real> *tree, const gvpm_photon_soa &ph, size_t n, const gvpm_ray_soa &rays, size_t nRays,
              const gvpm_vpm_sample_soa &smp, size_t nSamples, const Scene<Real> &sc, int nbCameraSamples,
              int threads, float *out, float *mvol, uint32_t *sampleCounts,
              std::vector<std::vector<uint32_t>> *nbr) {
  // per-sample accumulators first (gRec), folded into the pixel with the reference's normalisation
  std::vector<Accum<Real>> perSample(nSamples);
  const size_t tile = 4096;
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      size_t b = next.fetch_add(tile);
      if (b >= nSamples) break;
      size_t e = std::min(nSamples, b + tile);
      for (size_t i = b; i < e; ++i) {
        CamRay<Real> ray = loadRay<Real>(rays, smp.ray[i]);
        typename Scene<Real>::VpmSample s = sc.loadVpmSample(smp, i, ray);
        Accum<Real> &acc = perSample[i];
        uint32_t found = 0, contrib = 0;
        std::vector<uint32_t> *mine = nbr ? &(*nbr)[i] : nullptr;
        const V3<Real> q = ray.o + s.t * ray.d;
        auto visit = [&](uint32_t orig) {
          Photon<Real> p = loadPhoton<Real>(ph, orig);
          int r = sc.vpmFunctor(ray, s, p, acc);
          ++found;
          if (r == 2) ++contrib;
          if (mine) mine->push_back(orig | (r == 2 ? 0x80000000u : 0u));
        };
        if (tree) {
          tree->rangeQuery(q, s.radius, [&](uint32_t nodeIdx) { visit(tree->nodes[nodeIdx].orig); });
        } else {
          for (size_t j = 0; j < n; ++j)
            if (sc.sphereTest(q, s.radius, V3<Real>(ph.pos + 3 * j))) visit((uint32_t)j);
        }
        if (sampleCounts) { sampleCounts[2 * i] = found; sampleCounts[2 * i + 1] = contrib; }
      }
    }
  };
  if (threads <= 1) worker();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
  }
  // gp.mediumFlux += gRec.mediumFlux * normalization (gvpm.cpp:1177-1182), MVol += found (:1175)
  const Real normalization = (Real)1 / (Real)nbCameraSamples;
  std::vector<Accum<Real>> pix(nRays);
  for (size_t r = 0; r < nRays; ++r) mvol[r] = 0.f;
  for (size_t i = 0; i < nSamples; ++i) {
    Accum<Real> &g = pix[smp.ray[i]];
    g.mediumFlux += perSample[i].mediumFlux * normalization;
    for (int k = 0; k < 4; ++k) {
      g.shifted[k] += perSample[i].shifted[k] * normalization;
      g.weighted[k] += perSample[i].weighted[k] * normalization;
    }
  }
  for (size_t r = 0; r < nRays; ++r) pix[r].store(out + GVPM_OUT_FLOATS * r);
  (void)nSamples;
}
}  // namespace

extern "C" long long gvpm_oracle_vpm(void *tree, const gvpm_photon_soa *ph, size_t n, const gvpm_ray_soa *rays,
                                     size_t nRays, const gvpm_vpm_sample_soa *smp, size_t nSamples,
                                     const gvpm_medium *med, const gvpm_config *cfg, const float *tri, size_t n_tri,
                                     int nbCameraSamples, int use_double, int threads, float *out, float *mvol,
                                     uint32_t *sampleCounts, uint64_t *nbr_offsets, uint32_t *nbr_idx, size_t cap,
                                     double *gather_ms) {
  TreeHandle *h = (TreeHandle *)tree;
  if (h && (h->dbl != (use_double != 0) || h->n != n)) return -2;
  std::vector<std::vector<uint32_t>> nbr;
  if (nbr_offsets) nbr.resize(nSamples);
  auto t0 = std::chrono::steady_clock::now();
  if (use_double) {
    Scene<double> sc(*med, *cfg, 0.0);
    sc.occ.set(tri, n_tri);
    vpmRange<double>(h ? &h->td : nullptr, *ph, n, *rays, nRays, *smp, nSamples, sc, nbCameraSamples, threads, out,
                     mvol, sampleCounts, nbr_offsets ? &nbr : nullptr);
  } else {
    Scene<float> sc(*med, *cfg, 0.f);
    sc.occ.set(tri, n_tri);
    vpmRange<float>(h ? &h->tf : nullptr, *ph, n, *rays, nRays, *smp, nSamples, sc, nbCameraSamples, threads, out,
                    mvol, sampleCounts, nbr_offsets ? &nbr : nullptr);
  }
  if (gather_ms)
    *gather_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  // MVol per ray = sum of `found` over its samples
  if (sampleCounts)
    for (size_t i = 0; i < nSamples; ++i) mvol[smp->ray[i]] += (float)sampleCounts[2 * i];
  long long total = 0;
  if (nbr_offsets) {
    for (size_t i = 0; i < nbr.size(); ++i) {
      nbr_offsets[i] = (uint64_t)total;
      std::sort(nbr[i].begin(), nbr[i].end(),
                [](uint32_t a, uint32_t b) { return (a & 0x7fffffffu) < (b & 0x7fffffffu); });
      for (uint32_t v : nbr[i]) {
        if ((size_t)total < cap && nbr_idx) nbr_idx[total] = v;
        ++total;
      }
    }
    nbr_offsets[nbr.size()] = (uint64_t)total;
  }
  return total;
}

// The per-(ray, photon) uniform number of the restated sppm BRE (3-D kernel: the sampler->next1D() of bre.cpp:217, here the
// counter-based hash Scene::sppmUniform).  Exported so that the reference harness (ref_functor.cpp) is fed the same draws.
// xi: [nRays * nPhotons].
extern "C" void gvpm_oracle_sppm_uniforms(const gvpm_ray_soa *rays, size_t nRays, size_t nPhotons, const gvpm_medium *med,
                                          const gvpm_config *cfg, float *xi) {
  Scene<float> sc(*med, *cfg, 1.f);
  for (size_t i = 0; i < nRays; ++i) {
    CamRay<float> ray = loadRay<float>(*rays, i);
    for (size_t j = 0; j < nPhotons; ++j) xi[i * nPhotons + j] = sc.sppmUniform(ray, (uint32_t)j);
  }
}
