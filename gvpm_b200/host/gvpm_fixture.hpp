// gvpm_fixture.hpp — on-disk fixture of one gather iteration.
//
// Header-only and free of Mitsuba and CUDA types, so that the same file serves (a) a ~40-line dump hook inside a
// real Mitsuba build of the reference (at gvpm.cpp:1040-1042 the integrator already holds the
// flattened inputs next to the GatherPoint results its own CPU gather produced) and (b) this repository's tests and
// tools (gvpm_b200/fixture.py reads and writes the same format).  A fixture written by the hook on a machine that can
// build the reference pins the shift functors (shift_volume_photon.cpp etc.) against the real renderer.
//
// Format, little-endian:
//   char magic[8] = "GVPMFIX1"; uint32 n_sections; uint32 reserved
//   per section: char name[32] (NUL padded); uint32 dtype (0 f32, 1 u8, 2 u32, 3 i32, 4 f64, 5 u64); uint32 reserved;
//                uint64 count; data (count elements), zero-padded to a multiple of 8 bytes
// Sections: "medium" f32[9] = sigma_s[3], sigma_a[3], phase_type, hg_g, sampling_weight; "config" f64[16] = the
// gvpm_config fields in declaration order; "radius" f32[1]; "occluders" f32[9*n_tri]; "photon.<array>" and
// "ray.<array>" = the arrays of gvpm_photon_soa / gvpm_ray_soa under their field names; optional results of the
// producer's own gather: "expected.out" f32[27*n_rays] (mediumFlux, shiftedMediumFlux[4], weightedMediumFlux[4] of the
// iteration, un-normalised), "expected.nbr_offsets" u64[n_rays+1] + "expected.nbr_idx" u32 (photon indices the
// functor was called with, per ray), "meta.producer" u8 (text).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gvpm_b200.h"

namespace gvpm_fixture {

enum DType : uint32_t { F32 = 0, U8 = 1, U32 = 2, I32 = 3, F64 = 4, U64 = 5 };
inline size_t dtype_size(uint32_t d) {
  static const size_t s[6] = {4, 1, 4, 4, 8, 8};
  if (d > 5) throw std::runtime_error("gvpm_fixture: unknown dtype");
  return s[d];
}

class Writer {
 public:
  explicit Writer(const std::string &path) : m_f(std::fopen(path.c_str(), "wb")) {
    if (!m_f) throw std::runtime_error("gvpm_fixture: cannot open " + path);
    const char magic[8] = {'G', 'V', 'P', 'M', 'F', 'I', 'X', '1'};
    const uint32_t zero[2] = {0, 0};
    put(magic, 8);
    put(zero, 8);  // n_sections patched in close()
  }
  ~Writer() { if (m_f) std::fclose(m_f); }
  void add(const char *name, DType dt, const void *data, uint64_t count) {
    char nm[32] = {0};
    std::strncpy(nm, name, 31);
    const uint32_t hdr[2] = {(uint32_t)dt, 0};
    put(nm, 32);
    put(hdr, 8);
    put(&count, 8);
    const size_t bytes = (size_t)count * dtype_size(dt);
    if (bytes) put(data, bytes);
    const char pad[8] = {0};
    if (bytes % 8) put(pad, 8 - bytes % 8);
    ++m_n;
  }
  void add(const char *name, const float *p, uint64_t n) { add(name, F32, p, n); }
  void add(const char *name, const uint8_t *p, uint64_t n) { add(name, U8, p, n); }
  void add(const char *name, const uint32_t *p, uint64_t n) { add(name, U32, p, n); }
  void add(const char *name, const int32_t *p, uint64_t n) { add(name, I32, p, n); }
  void add(const char *name, const double *p, uint64_t n) { add(name, F64, p, n); }
  void add(const char *name, const uint64_t *p, uint64_t n) { add(name, U64, p, n); }
  void close() {
    if (!m_f) return;
    std::fseek(m_f, 8, SEEK_SET);
    put(&m_n, 4);
    std::fclose(m_f);
    m_f = nullptr;
  }

 private:
  void put(const void *p, size_t n) {
    if (std::fwrite(p, 1, n, m_f) != n) throw std::runtime_error("gvpm_fixture: short write");
  }
  std::FILE *m_f;
  uint32_t m_n = 0;
};

// One G-BRE iteration: everything gvpm_gather_bre reads, plus (optionally) what the producer's own gather returned.
inline void write_bre_fixture(const std::string &path, const gvpm_medium &m, const gvpm_config &c, float radius,
                              const float *tris, size_t nTris, const gvpm_photon_soa &ph, size_t nPhotons,
                              const gvpm_ray_soa &r, size_t nRays, const float *expectedOut /* [27*nRays] or null */,
                              const uint64_t *nbrOffsets /* [nRays+1] or null */, const uint32_t *nbrIdx,
                              const char *producer) {
  Writer w(path);
  const float med[9] = {m.sigma_s[0], m.sigma_s[1], m.sigma_s[2], m.sigma_a[0], m.sigma_a[1], m.sigma_a[2],
                        (float)m.phase_type, m.hg_g, m.sampling_weight};
  w.add("medium", med, 9);
  const double cfg[16] = {(double)c.max_depth, (double)c.min_depth, (double)c.lighting_mode, (double)c.use_mis,
                          (double)c.use_shift_null, (double)c.path_set, (double)c.power_heuristic, (double)c.kernel_3d,
                          (double)c.film_w, (double)c.film_h, (double)c.shadow_maxt_scale, (double)c.epsilon,
                          (double)c.long_beams, (double)c.rng_seed, (double)c.beam_kernel_1d, (double)c.sppm_primal};
  w.add("config", cfg, 16);
  w.add("radius", &radius, 1);
  w.add("occluders", tris, 9 * nTris);
  const size_t n = nPhotons;
  w.add("photon.pos", ph.pos, 3 * n);
  w.add("photon.flux", ph.flux, 3 * n);
  w.add("photon.parent_pos", ph.parent_pos, 3 * n);
  w.add("photon.pred_pos", ph.pred_pos, 3 * n);
  w.add("photon.parent_n", ph.parent_n, 3 * n);
  w.add("photon.prefix_flux", ph.prefix_flux, 3 * n);
  w.add("photon.parent_albedo", ph.parent_albedo, 3 * n);
  w.add("photon.parent_pdf", ph.parent_pdf, n);
  w.add("photon.edge_pdf", ph.edge_pdf, n);
  w.add("photon.rr_weight", ph.rr_weight, n);
  w.add("photon.parent_type", ph.parent_type, n);
  w.add("photon.depth", ph.depth, n);
  w.add("photon.path_id", ph.path_id, n);
  const size_t q = nRays;
  w.add("ray.o", r.o, 3 * q);
  w.add("ray.d", r.d, 3 * q);
  w.add("ray.mint", r.mint, q);
  w.add("ray.maxt", r.maxt, q);
  w.add("ray.edge_len", r.edge_len, q);
  w.add("ray.eye_contrib", r.eye_contrib, 3 * q);
  w.add("ray.xi", r.xi, q);
  w.add("ray.px", r.px, q);
  w.add("ray.py", r.py, q);
  w.add("ray.edge_id", r.edge_id, q);
  w.add("ray.off_valid", r.off_valid, 4 * q);
  w.add("ray.off_o", r.off_o, 12 * q);
  w.add("ray.off_d", r.off_d, 12 * q);
  w.add("ray.off_len", r.off_len, 4 * q);
  w.add("ray.off_eye", r.off_eye, 12 * q);
  w.add("ray.off_sensor", r.off_sensor, 4 * q);
  if (expectedOut) w.add("expected.out", expectedOut, (uint64_t)GVPM_OUT_FLOATS * q);
  if (nbrOffsets && (nbrIdx || nbrOffsets[q] == 0)) {
    w.add("expected.nbr_offsets", nbrOffsets, q + 1);
    w.add("expected.nbr_idx", nbrIdx, nbrOffsets[q]);
  }
  if (producer) w.add("meta.producer", (const uint8_t *)producer, std::strlen(producer));
  w.close();
}

}  // namespace gvpm_fixture
