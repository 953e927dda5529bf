// gvpm_oracle_beams.cpp — TEST INFRASTRUCTURE (see gvpm_oracle.hpp).  G-Beams 3D gather over every
// (camera ray, photon beam) pair: the functor BeamGradRadianceQuery::operator() restated in
// gvpm_oracle.hpp, driven like computeVolumeGradientBeams (gvpm/gvpm.cpp:880-986).  This entry point is
// the tree-independent brute force (every beam offered to every ray with tmin = 0, tmax = length), which
// is what the reference's SubBeamBVH traversal (photonmapper/beams_accel.h:169-203) selects up to
// measure-zero coincidences of tNear with a sub-beam boundary.
#include "gvpm_oracle.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace gvpm_oracle;

namespace {
template <typename Real>
void beamsRange(const gvpm_beam_soa &bs, size_t nBeams, const gvpm_ray_soa &rays, size_t nRays, const Scene<Real> &sc,
                int threads, float *out, uint32_t *counts, std::vector<std::vector<uint32_t>> *nbr) {
  std::vector<typename Scene<Real>::Beam> beams(nBeams);
  for (size_t i = 0; i < nBeams; ++i) beams[i] = Scene<Real>::loadBeam(bs, i);
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
        uint32_t nGeom = 0, nContrib = 0;
        if (ray.edgeLen >= ray.mint) {
          for (size_t j = 0; j < nBeams; ++j) {
            int r = sc.beamFunctor(ray, beams[j], (uint32_t)j, acc);
            if (r >= 1) {
              ++nGeom;
              if (r == 2) ++nContrib;
              if (nbr) (*nbr)[i].push_back((uint32_t)j | (r == 2 ? 0x80000000u : 0u));
            }
          }
        }
        acc.store(out + GVPM_OUT_FLOATS * i);
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
}
}  // namespace

extern "C" long long gvpm_oracle_beams(const gvpm_beam_soa *beams, size_t nBeams, const gvpm_ray_soa *rays,
                                       size_t nRays, const gvpm_medium *med, const gvpm_config *cfg,
                                       const float *tri, size_t n_tri, float radius, int use_double, int threads,
                                       float *out, uint32_t *counts, uint64_t *nbr_offsets, uint32_t *nbr_idx,
                                       size_t cap, double *gather_ms) {
  std::vector<std::vector<uint32_t>> nbr;
  if (nbr_offsets) nbr.resize(nRays);
  auto t0 = std::chrono::steady_clock::now();
  if (use_double) {
    Scene<double> sc(*med, *cfg, (double)radius);
    sc.occ.set(tri, n_tri);
    beamsRange<double>(*beams, nBeams, *rays, nRays, sc, threads, out, counts, nbr_offsets ? &nbr : nullptr);
  } else {
    Scene<float> sc(*med, *cfg, radius);
    sc.occ.set(tri, n_tri);
    beamsRange<float>(*beams, nBeams, *rays, nRays, sc, threads, out, counts, nbr_offsets ? &nbr : nullptr);
  }
  if (gather_ms)
    *gather_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
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

// ---- sppm primal photon beams (photonmapper/beams.h:29-223, sppm.cpp:823-860), brute force over every
// (camera beam, sub-beam) pair.  The sub-beam split restates the SubBeamBVH constructor (beams_accel.h:98-124) in
// the reference's single-precision arithmetic: size = average length / 10, ceil(length / size) equal pieces.
namespace {
struct SubBeam { uint32_t beam, ordinal; float t1, t2; bool first, last; };
std::vector<SubBeam> splitSubBeams(const gvpm_beam_soa &bs, size_t nBeams) {
  std::vector<float> len(nBeams);
  float avgSize = 0.f;
  for (size_t i = 0; i < nBeams; ++i) {
    const float *o = bs.origin + 3 * i, *e = bs.end + 3 * i;
    const float d[3] = {e[0] - o[0], e[1] - o[1], e[2] - o[2]};
    len[i] = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    avgSize += len[i];
  }
  if (nBeams) avgSize /= (float)nBeams;
  const float subbeamSize = avgSize / 10;
  std::vector<SubBeam> subs;
  for (size_t i = 0; i < nBeams; ++i) {
    int nSub = subbeamSize > 0.f ? (int)std::ceil(len[i] / subbeamSize) : 1;
    if (nSub < 1) nSub = 1;
    const float lengthSub = len[i] / nSub;
    for (int k = 0; k < nSub; ++k)
      subs.push_back({(uint32_t)i, (uint32_t)k, lengthSub * k, lengthSub * (k + 1), k == 0, k == nSub - 1});
  }
  return subs;
}

template <typename Real>
void sppmBeamsRange(const gvpm_beam_soa &bs, size_t nBeams, const gvpm_ray_soa &rays, size_t nRays,
                    const Scene<Real> &sc, int technique, int threads, float *out, uint32_t *counts,
                    std::vector<std::vector<uint32_t>> *nbr) {
  std::vector<typename Scene<Real>::Beam> beams(nBeams);
  for (size_t i = 0; i < nBeams; ++i) beams[i] = Scene<Real>::loadBeam(bs, i);
  const std::vector<SubBeam> subs = splitSubBeams(bs, nBeams);
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const size_t b = next.fetch_add(64);
      if (b >= nRays) break;
      const size_t e = std::min(nRays, b + 64);
      for (size_t i = b; i < e; ++i) {
        CamRay<Real> ray = loadRay<Real>(rays, i);
        V3<Real> Li(0, 0, 0);
        uint32_t nGeom = 0, nContrib = 0;
        if (ray.edgeLen >= ray.mint) {
          for (const SubBeam &s : subs) {
            const int r = sc.sppmBeamFunctor(ray, beams[s.beam], s.beam, (Real)s.t1, (Real)s.t2, s.first, s.last,
                                             s.ordinal, technique, Li);
            if (r >= 1) {
              ++nGeom;
              if (r == 2) ++nContrib;
              if (nbr) (*nbr)[i].push_back(s.beam | (r == 2 ? 0x80000000u : 0u));
            }
          }
        }
        out[3 * i] = (float)Li.x; out[3 * i + 1] = (float)Li.y; out[3 * i + 2] = (float)Li.z;
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
}
}  // namespace

extern "C" long long gvpm_oracle_sppm_beams(const gvpm_beam_soa *beams, size_t nBeams, const gvpm_ray_soa *rays,
                                            size_t nRays, const gvpm_medium *med, const gvpm_config *cfg, float radius,
                                            int technique, int use_double, int threads, float *out, uint32_t *counts,
                                            uint64_t *nbr_offsets, uint32_t *nbr_idx, size_t cap) {
  if (technique < GVPM_BEAM_1D || technique > GVPM_BEAM_3D_OPTIMIZED) return -1;
  std::vector<std::vector<uint32_t>> nbr;
  if (nbr_offsets) nbr.resize(nRays);
  if (use_double) {
    Scene<double> sc(*med, *cfg, (double)radius);
    sppmBeamsRange<double>(*beams, nBeams, *rays, nRays, sc, technique, threads, out, counts, nbr_offsets ? &nbr : nullptr);
  } else {
    Scene<float> sc(*med, *cfg, radius);
    sppmBeamsRange<float>(*beams, nBeams, *rays, nRays, sc, technique, threads, out, counts, nbr_offsets ? &nbr : nullptr);
  }
  long long total = 0;
  if (nbr_offsets) {
    for (size_t i = 0; i < nbr.size(); ++i) {
      nbr_offsets[i] = (uint64_t)total;
      std::stable_sort(nbr[i].begin(), nbr[i].end(),
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

// the sub-beam table itself (parity of the split with the CUDA build): returns the number of sub-beams; t12 / beam
// may be NULL or hold `cap` entries
extern "C" long long gvpm_oracle_subbeams(const gvpm_beam_soa *beams, size_t nBeams, float *t12, uint32_t *beam, size_t cap) {
  const std::vector<SubBeam> subs = splitSubBeams(*beams, nBeams);
  for (size_t i = 0; i < subs.size() && i < cap; ++i) {
    if (t12) { t12[2 * i] = subs[i].t1; t12[2 * i + 1] = subs[i].t2; }
    if (beam) beam[i] = subs[i].beam;
  }
  return (long long)subs.size();
}

// The two uniform numbers per (ray, beam) pair that the restated functor derives from its counter-based hash (dims 0, 1:
// the point on the photon beam, the point on the camera segment).  Exported so that the reference functor harness
// (ref_functor.cpp), whose BeamKernelRecord draws them from a Sampler, is fed the very same numbers.
// xi: [nRays * nBeams * 2].
extern "C" void gvpm_oracle_beam_uniforms(const gvpm_ray_soa *rays, size_t nRays, size_t nBeams, const gvpm_medium *med,
                                          const gvpm_config *cfg, float *xi) {
  Scene<float> sc(*med, *cfg, 1.f);
  for (size_t i = 0; i < nRays; ++i) {
    CamRay<float> ray = loadRay<float>(*rays, i);
    for (size_t j = 0; j < nBeams; ++j) {
      xi[2 * (i * nBeams + j)] = sc.beamUniform(ray, (uint32_t)j, 0);
      xi[2 * (i * nBeams + j) + 1] = sc.beamUniform(ray, (uint32_t)j, 1);
    }
  }
}

// Same, for arbitrary hash dimensions: entry k of the table is (beam index, dim0, dim1); xi: [nRays * n * 2].  The sppm
// techniques draw dims (0, 1) per (ray, beam), the naive one dims (2 + 2k, 3 + 2k) for the beam's k-th sub-beam.
extern "C" void gvpm_oracle_beam_uniform_dims(const gvpm_ray_soa *rays, size_t nRays, const uint32_t *beam,
                                              const uint32_t *dim0, const uint32_t *dim1, size_t n, const gvpm_medium *med,
                                              const gvpm_config *cfg, float *xi) {
  Scene<float> sc(*med, *cfg, 1.f);
  for (size_t i = 0; i < nRays; ++i) {
    CamRay<float> ray = loadRay<float>(*rays, i);
    for (size_t k = 0; k < n; ++k) {
      xi[2 * (i * n + k)] = sc.beamUniform(ray, beam[k], dim0[k]);
      xi[2 * (i * n + k) + 1] = sc.beamUniform(ray, beam[k], dim1[k]);
    }
  }
}
