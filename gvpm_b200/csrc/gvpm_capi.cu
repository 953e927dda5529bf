// gvpm_capi.cu — the extern "C" boundary of include/gvpm_b200.h: context, grow-only device
// buffers, uploads, build and gather launches.  No torch types, no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "gvpm_device.cuh"

namespace gvpm {
size_t sort_temp_bytes(uint32_t n);
cudaError_t run_sort(void *temp, size_t temp_bytes, const uint32_t *kin, uint32_t *kout, const uint32_t *vin,
                     uint32_t *vout, uint32_t n, cudaStream_t st);
int bounds_blocks(uint32_t n);
void launch_bounds(const float *pos, uint32_t n, float *partial, float *bounds, cudaStream_t st, uint32_t stride = 3);
void launch_morton(const float *pos, uint32_t n, const float *bounds, uint32_t *keys, uint32_t *vals,
                   cudaStream_t st, uint32_t stride = 3);
int pinhole_blocks(uint32_t n);
void launch_pinhole_fit(const float4 *rays, uint32_t n, double *partial, float *fit, unsigned *stats, cudaStream_t st,
                        const float *view_dir);
size_t frustum_occ_bytes();
void launch_frustum_keys(const float4 *rays, uint32_t n_rays, const float *pos, uint32_t stride, const uint32_t *par_src,
                         uint32_t par_stride, uint32_t par_shift, uint32_t n,
                         const FrustumGrid &G, uint32_t *occ, uint32_t *keys, uint32_t *vals, unsigned *coord_mag,
                         uint32_t *keepmask, uint32_t *block_kept, int sm_count, cudaStream_t st, uint32_t region_cap = 0,
                         const uint32_t *region_count = nullptr, uint32_t *cell_count = nullptr);
size_t cell_scan_temp_bytes(uint32_t n_keys);
cudaError_t launch_cell_scan(void *temp, size_t temp_bytes, const uint32_t *cell_count, uint32_t *cell_start, uint32_t n_keys,
                             cudaStream_t st);
void launch_pack_scatter(const PhotonStaging &S, uint32_t n, const uint32_t *keepmask, const uint32_t *keys, const uint32_t *rank,
                         const uint32_t *cell_start, float4 *aos, float4 *planes, uint32_t *orig, cudaStream_t st,
                         bool records_ready, const uint32_t *kept_dev);
void launch_frustum_mark(const float4 *rays, uint32_t n_rays, const FrustumGrid &G, uint32_t *occ, int sm_count, cudaStream_t st);
void launch_dispatch(const DispatchParams &P, cudaStream_t st, int ctas);
void launch_dispatch_signal(const SignalParams &P, cudaStream_t st);
void launch_flag_wait(const uint32_t *flags, int n, uint32_t target, unsigned *timeout, cudaStream_t st);
void launch_flag_set(const FlagSetParams &P, cudaStream_t st);
void launch_or_flag(uint32_t *dst, const uint32_t *src, cudaStream_t st);
void launch_l2_read(const void *p, size_t bytes, int reps, unsigned *sink, int sm_count, cudaStream_t st);
void launch_translate_idx(uint32_t *idx, unsigned long long n, const float4 *aos, cudaStream_t st);
void launch_compact_kept(const uint32_t *keys, const uint32_t *keepmask, const uint32_t *block_off, uint32_t n,
                         uint32_t *keys_c, uint32_t *vals_c, cudaStream_t st, uint32_t limit, uint32_t *overflow);
void launch_fill_u32(uint32_t *p, uint32_t n, uint32_t value, cudaStream_t st);
void launch_pack_sorted_kept(const PhotonStaging &S, uint32_t n, const uint32_t *keepmask, const uint32_t *sorted, uint32_t m,
                             float4 *aos, float4 *planes, uint32_t *orig, cudaStream_t st, bool records_ready);
void launch_scan_u32(uint32_t *vals, uint32_t nb, uint32_t *total, cudaStream_t st);
size_t cell_starts_scratch_bytes(uint32_t n_keys);
void launch_cell_starts(const uint32_t *sorted_keys, uint32_t n, uint32_t n_keys, uint32_t *cell_start, void *scratch,
                        int sm_count, cudaStream_t st);
cudaError_t run_sort_bits(void *temp, size_t temp_bytes, const uint32_t *kin, uint32_t *kout, const uint32_t *vin,
                          uint32_t *vout, uint32_t n, int bits, cudaStream_t st);
size_t ray_grid_bytes();
size_t ray_mask_bytes();
void launch_ray_region(const float4 *rays, uint32_t nRays, float radius, float *box, void *grid, uint32_t *mask,
                       float *bounds, int sm_count, cudaStream_t st);
void launch_keep_pruned(const float *pos, uint32_t n, const void *grid, const uint32_t *mask, uint32_t *keepmask,
                        uint32_t *block_kept, cudaStream_t st);
void launch_keys_kept(const float *pos, const uint32_t *vals, uint32_t m, const void *grid, uint32_t *keys, cudaStream_t st);
void launch_pack_pruned(const PhotonStaging &S, uint32_t n, const uint32_t *keepmask, const uint32_t *sorted, uint32_t m,
                        float4 *aos, float4 *planes, uint32_t *orig, cudaStream_t st);
void launch_pack_sorted(const PhotonStaging &S, float4 *aos, const uint32_t *sorted, uint32_t n, float4 *planes,
                        uint32_t *orig, cudaStream_t st, bool records_ready = false);
void launch_leaf_boxes(const float4 *p0, uint32_t n, uint32_t nLeaves, float radius, float4 *lo, float4 *hi,
                       cudaStream_t st);
void launch_level_boxes(const float4 *clo, const float4 *chi, uint32_t nChild, uint32_t nParent, float4 *plo,
                        float4 *phi, cudaStream_t st);
void launch_pack_rays(const RayStaging &S, uint32_t n, float4 *rays, cudaStream_t st);
void launch_pack_rays_range(const RayStaging &S, uint32_t r0, uint32_t r1, float4 *rays, cudaStream_t st);
cudaError_t launch_bre_traverse(const GatherParams &P, bool dump, int sm_count, cudaStream_t stream);
cudaError_t launch_bre_shade(const GatherParams &P, unsigned long long total, int sm_count, cudaStream_t stream);
void launch_sub_gather(const float4 *raw, const uint32_t *sorted, uint32_t n, float4 *out, cudaStream_t st);
void launch_subbeam_leaf_boxes(const float4 *subs, const float4 *beams, uint32_t n, uint32_t nLeaves, float radius,
                               float4 *lo, float4 *hi, cudaStream_t st);
cudaError_t launch_beam_traverse(const GatherParams &P, int sm_count, cudaStream_t stream);
cudaError_t launch_beam_shade(const GatherParams &P, unsigned long long total, int sm_count, cudaStream_t stream);
cudaError_t launch_beam_shade_sppm(const GatherParams &P, unsigned long long total, int sm_count, cudaStream_t stream);
void launch_plane_pack_sorted(const float4 *raw, const uint32_t *sorted, uint32_t n, float4 *planes, uint32_t *orig,
                              cudaStream_t st);
void launch_plane_leaf_boxes(const float4 *planes, uint32_t n, uint32_t nLeaves, float4 *lo, float4 *hi, cudaStream_t st);
cudaError_t launch_plane_gather(const GatherParams &P, bool dump, int sm_count, cudaStream_t stream);
cudaError_t launch_vpm_traverse(const GatherParams &P, bool dump, int sm_count, cudaStream_t stream);
cudaError_t launch_vpm_shade(const GatherParams &P, unsigned long long total, int sm_count, cudaStream_t stream);
void launch_gradient(const float *acc, int w, int h, int use_abs, float *thr, float *gx, float *gy,
                     cudaStream_t st, int reuse = 0, float inv_emitted = 1.f);
void launch_generate_rays(const RayGenParams &P, cudaStream_t st);
void launch_trace_count(const TraceParams &P, uint8_t *counts, unsigned long long *block_tot, unsigned long long *totals,
                        cudaStream_t st);
void launch_trace_emit(const TraceParams &P, const uint8_t *counts, const unsigned long long *block_off,
                       unsigned long long *last_path, cudaStream_t st);
void launch_pack_beams(const BeamStaging &S, uint32_t n, float4 *rec, float *len, cudaStream_t st);
void launch_beam_subsize_seq(const float *len, uint32_t n, float *subsize, cudaStream_t st);
void launch_sub_count(const float *len, uint32_t n, const float *subsize, uint32_t *block_tot, uint32_t *total, cudaStream_t st);
void launch_sub_emit(const float4 *rec, const float *len, uint32_t n, const float *subsize, const uint32_t *block_off,
                     uint32_t cap, float *sub_pos, float4 *sub_raw, cudaStream_t st);
void launch_pack_planes(const PlaneStaging &S, uint32_t n, float4 *rec, float *centre, uint32_t *flag, cudaStream_t st);
void launch_pack_samples(const SampleStaging &S, uint32_t n, uint32_t n_rays, float4 *packed, uint32_t *stats, cudaStream_t st);
size_t poisson_workspace_floats(size_t n);
long long poisson_solve_device(const float *tp, const float *dx, const float *dy, const float *direct, int W, int H,
                               float alpha, int irlsIterMax, float irlsRegInit, float irlsRegIter, int cgIterMax,
                               int cgIterCheck, float cgTolerance, float *ws, float *host_rz, float *rec,
                               cudaStream_t st);
}  // namespace gvpm

using namespace gvpm;

namespace {

struct DevBuf {  // grow-only device allocation
  void *p = nullptr;
  size_t cap = 0;
  bool pinned = false;  // exported to peers (CUDA IPC): the allocation must not move any more
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (pinned) return cudaErrorInvalidValue;
    // the old block stays valid until the new one exists; if memory is too tight for both, give the old one up
    // first and leave the buffer empty (cap = 0) when the allocation still fails
    const size_t want = bytes + bytes / 8 + 256;
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      if (p) cudaFree(p);
      p = nullptr;
      cap = 0;
      e = cudaMalloc(&q, want);
      if (e != cudaSuccess) return e;
    } else if (p) {
      cudaFree(p);
    }
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return (T *)p; }
};

std::string g_create_error;

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

struct gvpm_ctx {
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;  // compute, H2D, D2H
  cudaEvent_t pipe_ev[2 * 32 + 1] = {};  // per chunk: inputs landed / results ready; [64]: staging free
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::string err;
  uint64_t launches = 0;

  gvpm_medium medium{};
  bool have_medium = false;
  gvpm_config cfg{};
  bool have_cfg = false;

  DevBuf tri, tri_plane, tri_aux;
  uint32_t n_tri = 0;

  DevBuf ph_staging, ph_staging_alt;  // the selected photon staging buffer and the other one (double buffering)
  int ph_staging_sel = 0;
  // peer exchange over NVLink copy engines (gvpm_peer_*): interprocess events + peer mappings of the staging buffers
  cudaEvent_t ev_free[2] = {nullptr, nullptr};   // staging buffer b has been consumed by this context's last build
  cudaEvent_t ev_pushed[2] = {nullptr, nullptr}; // this context's slice has landed in every peer's staging buffer b
  cudaStream_t push_streams[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t push_kernel_stream = nullptr;  // highest priority: the SM push kernel takes its CTA slots as soon as they free
  int push_ctas = 32;                         // 0: copy engines (cudaMemcpyAsync per field and peer)
  cudaEvent_t push_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [0..3] push stream done, [4] slice ready
  int n_peers = 0, peer_self = -1;
  std::vector<void *> peer_staging[2];
  std::vector<cudaEvent_t> peer_free[2], peer_pushed[2];
  void *staging_ptr(int which) const { return which == ph_staging_sel ? ph_staging.p : ph_staging_alt.p; }
  uint32_t n_photons = 0;
  bool photons_loaded = false;
  bool photons_direct = false;   // gvpm_trace_photons_direct: the 128-byte records in `aos` ARE the photon set (no staging, no packing)
  // dispatched photon set (gvpm_build_dispatched): the records live in an inbox the peers wrote; one region of
  // region_cap records per sender, region_count[s] (device) of them filled
  const uint32_t *build_ovf = nullptr;   // device flag of the last frustum build (bounded compaction overflowed)
  float4 *aos_override = nullptr;
  uint32_t region_cap = 0;
  const uint32_t *region_count = nullptr;
  float4 *records() { return aos_override ? aos_override : aos.as<float4>(); }
  struct RayFit { bool concurrent = false; float C[3], m[3], u[3], v[3], delta, cosmin, xmin, xmax, ymin, ymax, count; };
  // photon dispatch between ranks (dispatch.cu, gvpm_dispatch_*)
  struct Dispatch {
    int n_peers = 0, self = -1;
    uint32_t region_cap = 0;
    bool connected = false;
    DevBuf inbox[2], ctrl, occ_all, grids_dev[2], keepbits, block_cnt, owner_map;
    bool shared_frame = false;                  // every rank projects on the same plane: one classification per photon
    float ux0 = 0.f, uy0 = 0.f, ux1 = 0.f, uy1 = 0.f;   // bounds of the owner map (union of the ranks' ray bounds)
    RayFit fit[GVPM_MAX_PEERS];
    char *peer_inbox[2][GVPM_MAX_PEERS] = {};   // mapped (or, same process, raw) pointers; [self] = own
    uint32_t *peer_ctrl[GVPM_MAX_PEERS] = {};
    bool peer_mapped[GVPM_MAX_PEERS] = {};      // opened through CUDA IPC (to be closed)
    FrustumGrid *grids_host[2] = {nullptr, nullptr};   // pinned
    uint32_t gen_push[2] = {0, 0}, gen_build[2] = {0, 0}, gen_collect[2] = {0, 0};
    std::vector<void *> shared_owned, shared_opened;   // gvpm_shared_buffer_create / _open
    cudaEvent_t ev_src = nullptr;
  } disp;
  DevBuf aos, keys_in, keys_out, vals_in, vals_out, sort_temp, planes, orig, box_lo, box_hi, bounds_partial, bounds;
  Tree tree{};
  float radius = 0.f;
  bool built = false;
  float extent_hint = 1.f;  // diagonal of the photon AABB (host copy, refreshed lazily)
  // gvpm_build_points_for_rays: occupancy grid of the uploaded rays; the hierarchy then holds only the photons they can
  // reach and is valid for that ray set (rays_gen) alone
  DevBuf ray_region;        // [0,32): ray box, [32,64): kept-photon counter, [64,..): RayGrid, the 32 KB cell mask, one keep bit per photon
  uint64_t rays_gen = 0, pruned_for_gen = 0;
  bool pruned = false;
  uint32_t n_kept = 0;
  // frustum grid (gvpm_device.cuh FrustumGrid): chosen by gvpm_build_points_for_rays when the uploaded rays are concurrent
  enum { ACCEL_BVH = 0, ACCEL_FRUSTUM = 1 };
  int accel = ACCEL_BVH;
  bool force_bvh = false;            // GVPM_ACCEL=bvh: A/B switch for kernel experiments
  FrustumGrid grid{};
  DevBuf cell_start, grid_occ, pin_scratch, trace_scratch, keepmask, cell_count;
  bool radix_frustum = false;        // GVPM_FRUSTUM_SORT=radix: the key sort of the perspective grid by radix sort (A/B switch)
  double kept_fraction_hint = 1.0;   // share of the photons the last frustum build kept (sizes the next one's sort)
  bool kept_hint_valid = false;      // the fraction comes from a count read back from a build over the same kind of set
  bool force_exact_build = false;    // a bounded build overflowed: size the next one from its exact count (one host sync)
  cudaEvent_t ev_hint = nullptr;
  bool hint_pending = false;
  uint32_t hint_n = 0;
  uint32_t hint_rays = 0xffffffffu;   // ray count the hint belongs to
  double trace_photons_per_path = 0.0;   // running estimate (sizes the first batch of gvpm_trace_photons)    // pin_scratch: [0,64) fit floats, [64,96) stats words, [128,..) block partials (doubles)
  uint64_t pin_gen = ~0ull;          // rays_gen the ray analysis below belongs to
  int dispatch_ctas = 0;             // side-stream dispatch: CTAs of its persistent grids (GVPM_DISPATCH_CTAS; 0 = one per chunk)
  bool have_view_dir = false;        // gvpm_set_view_direction: axis of the perspective grid's projection plane
  float view_dir[3] = {0.f, 0.f, 1.f};
  RayFit pin;
  float *pin_host = nullptr;         // pinned: 16 fit floats + 8 stats words

  DevBuf ray_staging, rays;
  uint32_t n_rays = 0;
  bool rays_loaded = false;

  DevBuf out, counts, nbr_offsets, nbr_idx, work_counter, pairs;
  unsigned long long pair_cap = 0;         // capacity of `pairs` in entries
  unsigned long long *pair_count_host = nullptr;  // pinned read-back of the pair counter
  unsigned long long last_pairs = 0;
  // asynchronous BRE gathers in flight (pending_check): pair count slot = pair_count_host[8 + index]
  struct AsyncGather { cudaEvent_t done = nullptr; uint64_t gen = 0; float *out = nullptr; uint32_t *counts = nullptr; unsigned long long cap = 0; };
  static constexpr int kRing = 8;
  AsyncGather ring[kRing];
  int ring_tail = 0, ring_n = 0;
  uint64_t state_gen = 0;   // bumped by everything a gather's result depends on (photons, build, rays, medium, config)
  // G-Beams
  DevBuf beams, beam_bounds, sub_pos, sub_raw, subs, beam_box_lo, beam_box_hi;
  DevBuf beam_staging, beam_len, beam_aux;  // raw gvpm_beam_soa arrays; beam lengths; [0] subbeamSize, [1] sub-beam total, [16..] block offsets
  float beam_subsize = 0.f;                 // avgLength / 10 (host copy when the beams came through gvpm_upload_beams)
  uint32_t n_subs = 0;                      // sub-beams of the uploaded beam set
  uint32_t n_beams = 0;
  bool beams_loaded = false, beams_built = false;
  Tree beam_tree{};
  float beam_radius = 0.f;
  // G-Planes
  DevBuf plane_raw, plane_pos, plane_rec, plane_orig, plane_box_lo, plane_box_hi, plane_bounds;
  DevBuf plane_staging;                     // raw gvpm_plane_soa arrays
  uint32_t n_planes = 0;
  bool planes_loaded = false, planes_built = false;
  DevBuf samples, sample_counts, mvol, sample_staging;  // G-VPM distance samples (+ their raw SoA arrays)
  uint32_t *sample_stats_host = nullptr;  // pinned: [0] max sample radius (float bits), [1] bad ray index seen; [2] plane flag
  uint32_t n_samples = 0;
  bool samples_loaded = false;
  float sample_radius_max = 0.f;
  DevBuf grad_in, grad_out;
  DevBuf poisson_io, poisson_ws;   // device copies of the four input planes + the result; solver workspace
  float poisson_ms = 0.f;
  float build_ms = 0.f, gather_ms = 0.f;
  bool timed_build = false, timed_gather = false;
  bool split_timed = false;   // the last gather recorded ev[4] / ev[5] around its traversal kernel
};

namespace {

int fail(gvpm_ctx *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(ctx, GVPM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
  } while (0)

// offsets of the raw photon SoA arrays inside the staging buffer
struct PhotonLayout {
  size_t off[13];
  size_t bytes;
  explicit PhotonLayout(size_t n) {
    const size_t sz[13] = {12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 4 * n, 4 * n, 4 * n, n, n, 4 * n};
    size_t o = 0;
    for (int i = 0; i < 13; ++i) { off[i] = o; o += align256(sz[i]); }
    bytes = o;
  }
};
PhotonStaging photon_staging_ptrs(const void *base, size_t n) {
  PhotonLayout L(n);
  const char *b = (const char *)base;
  PhotonStaging S;
  S.pos = (const float *)(b + L.off[0]);
  S.flux = (const float *)(b + L.off[1]);
  S.parent_pos = (const float *)(b + L.off[2]);
  S.pred_pos = (const float *)(b + L.off[3]);
  S.parent_n = (const float *)(b + L.off[4]);
  S.prefix_flux = (const float *)(b + L.off[5]);
  S.parent_albedo = (const float *)(b + L.off[6]);
  S.parent_pdf = (const float *)(b + L.off[7]);
  S.edge_pdf = (const float *)(b + L.off[8]);
  S.rr_weight = (const float *)(b + L.off[9]);
  S.parent_type = (const uint8_t *)(b + L.off[10]);
  S.depth = (const uint8_t *)(b + L.off[11]);
  S.path_id = (const uint32_t *)(b + L.off[12]);
  return S;
}
struct RayLayout {
  size_t off[16];
  size_t bytes;
  explicit RayLayout(size_t n) {
    const size_t sz[16] = {12 * n, 12 * n, 4 * n, 4 * n, 4 * n, 12 * n, 4 * n, 4 * n, 4 * n, 4 * n,
                           4 * n,  48 * n, 48 * n, 16 * n, 48 * n, 16 * n};
    size_t o = 0;
    for (int i = 0; i < 16; ++i) { off[i] = o; o += align256(sz[i]); }
    bytes = o;
  }
};
RayStaging ray_staging_ptrs(const void *base, size_t n) {
  RayLayout L(n);
  const char *b = (const char *)base;
  RayStaging S;
  S.o = (const float *)(b + L.off[0]);
  S.d = (const float *)(b + L.off[1]);
  S.mint = (const float *)(b + L.off[2]);
  S.maxt = (const float *)(b + L.off[3]);
  S.edge_len = (const float *)(b + L.off[4]);
  S.eye_contrib = (const float *)(b + L.off[5]);
  S.xi = (const float *)(b + L.off[6]);
  S.px = (const int32_t *)(b + L.off[7]);
  S.py = (const int32_t *)(b + L.off[8]);
  S.edge_id = (const int32_t *)(b + L.off[9]);
  S.off_valid = (const uint8_t *)(b + L.off[10]);
  S.off_o = (const float *)(b + L.off[11]);
  S.off_d = (const float *)(b + L.off[12]);
  S.off_len = (const float *)(b + L.off[13]);
  S.off_eye = (const float *)(b + L.off[14]);
  S.off_sensor = (const float *)(b + L.off[15]);
  return S;
}

int fill_params(gvpm_ctx *ctx, GatherParams &P, float *out_dev, uint32_t *counts_dev) {
  if (!ctx->have_medium || !ctx->have_cfg) return fail(ctx, GVPM_ERR_INVALID, "medium/config not set");
  if (!ctx->built) return fail(ctx, GVPM_ERR_INVALID, "gvpm_build_points has not been called");
  if (!ctx->rays_loaded) return fail(ctx, GVPM_ERR_INVALID, "no rays uploaded");
  if (ctx->pruned && ctx->pruned_for_gen != ctx->rays_gen)
    return fail(ctx, GVPM_ERR_INVALID, "the hierarchy was built by gvpm_build_points_for_rays for another ray set: rebuild");
  memset(&P, 0, sizeof(P));
  P.tree = ctx->tree;
  if (ctx->accel == gvpm_ctx::ACCEL_FRUSTUM) {
    P.grid = ctx->grid;
    P.cell_start = ctx->cell_start.as<uint32_t>();
    P.build_ovf = ctx->build_ovf;
  }
  P.planes = ctx->planes.as<float4>();
  P.aos = ctx->records();
  P.orig = ctx->orig.as<uint32_t>();
  P.rays = ctx->rays.as<float4>();
  P.n_rays = ctx->n_rays;
  const float r = ctx->radius;
  P.radius = r;
  P.radius_sq = r * r;
  // Float kernelVol = (4.0/3.0)*M_PI*pow(r,3) evaluated in double (shift_volume_photon.cpp:707)
  if (ctx->cfg.kernel_3d)
    P.kernel_vol = (float)((4.0 / 3.0) * (double)GVPM_PI * std::pow((double)r, 3));
  else
    P.kernel_vol = (float)((double)GVPM_PI * std::pow((double)r, 2));
  P.bounds = ctx->bounds.as<float>();
  for (int i = 0; i < 3; ++i) {
    P.sigma_s[i] = ctx->medium.sigma_s[i];
    P.sigma_t[i] = ctx->medium.sigma_s[i] + ctx->medium.sigma_a[i];
  }
  P.phase_type = ctx->medium.phase_type;
  P.hg_g = ctx->medium.hg_g;
  P.sampling_weight = ctx->medium.sampling_weight;
  P.cfg = ctx->cfg;
  P.tri = ctx->tri.as<float>();
  P.tri_plane = ctx->tri_plane.as<float4>();
  P.tri_aux = ctx->tri_aux.as<float2>();
  P.n_tri = ctx->n_tri;
  P.out = out_dev;
  P.counts = counts_dev;
  P.work_counter = ctx->work_counter.as<uint32_t>();
  P.pair_counter = (unsigned long long *)(ctx->work_counter.as<char>() + 8);
  P.ray_begin = 0;
  P.ray_end = ctx->n_rays;
  P.pairs = ctx->pairs.as<uint2>();
  P.pair_cap = ctx->pair_cap;
  // packets whose rays drift apart by more than this are traversed ray by ray (performance only)
  P.packet_spread_max = 0.f;  // derived in the kernel from the photon bounds
  return GVPM_OK;
}

// ---- G-BRE gather: traverse + shade without a host round trip in between ---------------------------------------------
// The shading kernel reads the pair count on the device, so a ray range is two back-to-back launches; the count is
// copied to a pinned slot behind them and looked at later (gather_finish / pending_poll).  If the pair list turns out to
// have overflowed, the range is run again the slow way (gather_range_sync: grow the list or split the range, one host
// sync per attempt).  The list is grow-only, so after the first iteration of a render this does not happen again.
int enqueue_range(gvpm_ctx *ctx, GatherParams &P, uint32_t r0, uint32_t r1, unsigned long long *host_slot) {
  if (r1 <= r0) { *host_slot = 0; return GVPM_OK; }
  P.ray_begin = r0;
  P.ray_end = r1;
  P.pairs = ctx->pairs.as<uint2>();
  P.pair_cap = ctx->pair_cap;
  CK(cudaMemsetAsync(ctx->work_counter.p, 0, 16, ctx->stream));
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  CK(launch_bre_traverse(P, false, ctx->sm_count, ctx->stream));
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->split_timed = true;
  CK(launch_bre_shade(P, ~0ull, ctx->sm_count, ctx->stream));
  ctx->launches += 2;
  CK(cudaMemcpyAsync(host_slot, P.pair_counter, 8, cudaMemcpyDeviceToHost, ctx->stream));
  return GVPM_OK;
}

// pair count a gather reports when its perspective grid is incomplete (bounded compaction overflow, build_frustum)
static const unsigned long long kBuildOverflow = 1ull << 62;

// the slow path: one host sync per attempt.  Rows [r0, r1) of P.out are cleared first.
int gather_range_sync(gvpm_ctx *ctx, GatherParams &P, uint32_t r0, uint32_t r1, int depth) {
  if (r1 <= r0) return GVPM_OK;
  for (;;) {
    P.ray_begin = r0;
    P.ray_end = r1;
    P.pairs = ctx->pairs.as<uint2>();
    P.pair_cap = ctx->pair_cap;
    CK(cudaMemsetAsync(ctx->work_counter.p, 0, 16, ctx->stream));
    CK(cudaMemsetAsync(P.out + (size_t)r0 * GVPM_OUT_FLOATS, 0, (size_t)(r1 - r0) * GVPM_OUT_FLOATS * sizeof(float), ctx->stream));
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    CK(launch_bre_traverse(P, false, ctx->sm_count, ctx->stream));
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->split_timed = true;
  ctx->split_timed = true;
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->pair_count_host, P.pair_counter, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const unsigned long long total = *ctx->pair_count_host;
    if (total <= ctx->pair_cap) {
      ctx->last_pairs += total;
      CK(launch_bre_shade(P, total, ctx->sm_count, ctx->stream));
      if (total) ctx->launches += 1;
      return GVPM_OK;
    }
    // overflow: grow within the budget, else halve the range
    size_t freeB = 0, totalB = 0;
    cudaMemGetInfo(&freeB, &totalB);
    const unsigned long long want = total + total / 8 + 1024;
    const unsigned long long budget = (freeB / 2 + ctx->pairs.cap) / sizeof(uint2);
    if (want <= budget) {
      cudaError_t e = ctx->pairs.reserve(want * sizeof(uint2));
      ctx->pair_cap = ctx->pairs.cap / sizeof(uint2);   // 0 when the allocation failed: nothing is stored through a stale pointer
      if (e != cudaSuccess) return fail(ctx, GVPM_ERR_CUDA, std::string("pair list: ") + cudaGetErrorString(e));
      continue;
    }
    if (r1 - r0 <= 1 || depth > 40) return fail(ctx, GVPM_ERR_CUDA, "pair list does not fit in device memory");
    const uint32_t mid = r0 + (r1 - r0) / 2;
    int rc = gather_range_sync(ctx, P, r0, mid, depth + 1);
    if (rc) return rc;
    return gather_range_sync(ctx, P, mid, r1, depth + 1);
  }
}

int reserve_pairs(gvpm_ctx *ctx, size_t entries) {
  if (ctx->pair_cap >= entries) return GVPM_OK;
  cudaError_t e = ctx->pairs.reserve(entries * sizeof(uint2));
  ctx->pair_cap = ctx->pairs.cap / sizeof(uint2);
  if (e != cudaSuccess) return fail(ctx, GVPM_ERR_CUDA, std::string("pair list: ") + cudaGetErrorString(e));
  return GVPM_OK;
}

// Asynchronous gathers (gvpm_gather_bre_into / _device) in flight: their pair counts are looked at without blocking when
// the next call comes in, and for good at gvpm_sync, which re-runs the latest gather if its list overflowed and nothing
// it depends on has changed since.
int pending_check(gvpm_ctx *ctx, bool block) {
  while (ctx->ring_n > 0) {
    gvpm_ctx::AsyncGather &g = ctx->ring[ctx->ring_tail];
    if (!block) {
      cudaError_t q = cudaEventQuery(g.done);
      if (q == cudaErrorNotReady) break;
      if (q != cudaSuccess) return fail(ctx, GVPM_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(q));
    } else {
      CK(cudaEventSynchronize(g.done));
    }
    const unsigned long long total = ctx->pair_count_host[8 + ctx->ring_tail];
    ctx->ring_tail = (ctx->ring_tail + 1) % gvpm_ctx::kRing;
    --ctx->ring_n;
    ctx->last_pairs = total;
    if (total <= g.cap) continue;
    if (total >= kBuildOverflow) {   // the bounded frustum build was too small: nothing was gathered
      ctx->force_exact_build = true;
      return fail(ctx, GVPM_ERR_INVALID,
                  "the perspective grid is incomplete - it was sized from the previous iteration's photon count and this "
                  "iteration kept over 25 % more, or a peer's photon dispatch never arrived (gvpm_dispatch_status): the gather "
                  "did not run.  Build again (the next build reads its exact count) and gather again");
    }
    // overflow: make the list large enough for the next one
    cudaStreamSynchronize(ctx->stream);
    int rc = reserve_pairs(ctx, total + total / 8 + 1024);
    if (rc) return rc;
    const bool latest = ctx->ring_n == 0 && g.gen == ctx->state_gen;
    if (block && latest) {  // same photons, rays and output buffers: run it again, the slow and safe way
      GatherParams P;
      rc = fill_params(ctx, P, g.out, g.counts);
      if (rc) return rc;
      ctx->last_pairs = 0;
      rc = gather_range_sync(ctx, P, 0, ctx->n_rays, 0);
      if (rc) return rc;
      CK(cudaEventRecord(ctx->ev[3], ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    } else {
      return fail(ctx, GVPM_ERR_INVALID,
                  "an asynchronous gather (gvpm_gather_bre_into / _device) overflowed its pair list and its inputs have "
                  "been replaced since: its results are incomplete.  Call gvpm_sync after such a gather before the next "
                  "upload / build / gather (the list has been enlarged)");
    }
  }
  return GVPM_OK;
}

int gather_common(gvpm_ctx *ctx, float *out_dev, uint32_t *counts_dev, unsigned long long *host_slot) {
  GatherParams P;
  int rc = fill_params(ctx, P, out_dev, counts_dev);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  const size_t n = ctx->n_rays;
  *host_slot = 0;
  if (n) {
    if (ctx->pair_cap == 0) {
      rc = reserve_pairs(ctx, std::max<size_t>(1u << 20, 16 * n));
      if (rc) return rc;
      P.pairs = ctx->pairs.as<uint2>();
      P.pair_cap = ctx->pair_cap;
    }
    CK(cudaMemsetAsync(out_dev, 0, n * GVPM_OUT_FLOATS * sizeof(float), ctx->stream));
    rc = enqueue_range(ctx, P, 0, (uint32_t)n, host_slot);
    if (rc) return rc;
  }
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed_gather = true;
  return GVPM_OK;
}

// blocking tail of a gather whose results the caller reads right away: wait, check the pair count, re-run on overflow.
int gather_finish(gvpm_ctx *ctx, float *out_dev, uint32_t *counts_dev, const unsigned long long *host_slot, bool *redone) {
  CK(cudaStreamSynchronize(ctx->stream));
  if (redone) *redone = false;
  const unsigned long long total = *host_slot;
  ctx->last_pairs = total;
  if (total <= ctx->pair_cap) return GVPM_OK;
  if (total >= kBuildOverflow) {
    ctx->force_exact_build = true;
    return fail(ctx, GVPM_ERR_INVALID,
                "the perspective grid is incomplete - it was sized from the previous iteration's photon count and this "
                "iteration kept over 25 % more, or a peer's photon dispatch never arrived (gvpm_dispatch_status): the gather "
                "did not run.  Build again (the next build reads its exact count) and gather again");
  }
  int rc = reserve_pairs(ctx, total + total / 8 + 1024);
  if (rc) return rc;
  GatherParams P;
  rc = fill_params(ctx, P, out_dev, counts_dev);
  if (rc) return rc;
  ctx->last_pairs = 0;
  rc = gather_range_sync(ctx, P, 0, ctx->n_rays, 0);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (redone) *redone = true;
  return GVPM_OK;
}

// asynchronous variant: the count goes to a ring slot that pending_check looks at later
int gather_async(gvpm_ctx *ctx, float *out_dev, uint32_t *counts_dev) {
  int rc = pending_check(ctx, false);
  if (rc) return rc;
  if (ctx->ring_n == gvpm_ctx::kRing) {       // ring full: wait for the oldest
    gvpm_ctx::AsyncGather &g = ctx->ring[ctx->ring_tail];
    CK(cudaEventSynchronize(g.done));
    rc = pending_check(ctx, false);
    if (rc) return rc;
  }
  const int slot = (ctx->ring_tail + ctx->ring_n) % gvpm_ctx::kRing;
  gvpm_ctx::AsyncGather &g = ctx->ring[slot];
  rc = gather_common(ctx, out_dev, counts_dev, ctx->pair_count_host + 8 + slot);
  if (rc) return rc;
  g.gen = ctx->state_gen;
  g.out = out_dev;
  g.counts = counts_dev;
  g.cap = ctx->pair_cap;
  CK(cudaEventRecord(g.done, ctx->stream));
  ++ctx->ring_n;
  return GVPM_OK;
}

}  // namespace

extern "C" {

int gvpm_abi_version(void) { return GVPM_ABI_VERSION; }

int gvpm_ctx_create(int device, gvpm_ctx **out) {
  if (!out) return GVPM_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                     " (gvpm_b200 has no CPU fallback)";
    return GVPM_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return GVPM_ERR_INVALID; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return GVPM_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                     "; this library is built for sm_100a only";
    return GVPM_ERR_NO_DEVICE;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return GVPM_ERR_CUDA; }
  gvpm_ctx *ctx = new gvpm_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete ctx;
    return GVPM_ERR_CUDA;
  }
  for (auto &ev : ctx->ev) cudaEventCreate(&ev);
  cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking);
  for (auto &ev : ctx->pipe_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  for (int b = 0; b < 2; ++b) {
    cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming | cudaEventInterprocess);
    cudaEventCreateWithFlags(&ctx->ev_pushed[b], cudaEventDisableTiming | cudaEventInterprocess);
  }
  for (auto &ps : ctx->push_streams) cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&ctx->push_kernel_stream, cudaStreamNonBlocking, hi);
  }
  for (auto &ev : ctx->push_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  ctx->work_counter.reserve(256);
  // [0,8): pair / kept-photon counters, the solver's 3-float residual; [8,16): async gather ring; [16,48): ray chunks
  cudaHostAlloc((void **)&ctx->pair_count_host, 48 * sizeof(unsigned long long), cudaHostAllocDefault);
  memset(ctx->pair_count_host, 0, 48 * sizeof(unsigned long long));
  for (auto &g : ctx->ring) cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming);
  cudaHostAlloc((void **)&ctx->pin_host, 128, cudaHostAllocDefault);
  cudaEventCreateWithFlags(&ctx->ev_hint, cudaEventDisableTiming);
  { const char *e = getenv("GVPM_ACCEL"); ctx->force_bvh = e && !strcmp(e, "bvh"); }
  { const char *e = getenv("GVPM_FRUSTUM_SORT"); ctx->radix_frustum = e && !strcmp(e, "radix"); }
  // side-stream dispatch: one CTA per SM by default (N = 8, cfg5: 32 CTAs 1.96 ms / step, 96: 1.11, one per chunk: 1.13)
  { const char *e = getenv("GVPM_DISPATCH_CTAS"); ctx->dispatch_ctas = e ? std::max(0, atoi(e)) : -1; }
  cudaHostAlloc((void **)&ctx->sample_stats_host, 64, cudaHostAllocDefault);
  memset(ctx->sample_stats_host, 0, 64);
  ctx->bounds.reserve(256);
  ctx->bounds_partial.reserve(1024 * 6 * sizeof(float));
  *out = ctx;
  return GVPM_OK;
}

int gvpm_ctx_destroy(gvpm_ctx *ctx) {
  if (!ctx) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf *bufs[] = {&ctx->tri, &ctx->tri_plane, &ctx->tri_aux, &ctx->ph_staging, &ctx->ph_staging_alt, &ctx->keys_in, &ctx->keys_out, &ctx->vals_in,
                    &ctx->vals_out, &ctx->sort_temp, &ctx->planes, &ctx->orig, &ctx->box_lo, &ctx->box_hi,
                    &ctx->bounds_partial, &ctx->bounds, &ctx->ray_staging, &ctx->rays, &ctx->out, &ctx->counts,
                    &ctx->nbr_offsets, &ctx->nbr_idx, &ctx->work_counter, &ctx->grad_in, &ctx->grad_out, &ctx->pairs,
                    &ctx->samples, &ctx->sample_counts, &ctx->mvol, &ctx->beams, &ctx->beam_bounds, &ctx->sub_pos,
                    &ctx->sub_raw, &ctx->subs, &ctx->beam_box_lo, &ctx->beam_box_hi, &ctx->aos, &ctx->plane_raw,
                    &ctx->plane_pos, &ctx->plane_rec, &ctx->plane_orig, &ctx->plane_box_lo, &ctx->plane_box_hi,
                    &ctx->plane_bounds, &ctx->ray_region, &ctx->poisson_io, &ctx->poisson_ws, &ctx->beam_staging, &ctx->beam_len,
                    &ctx->beam_aux, &ctx->plane_staging, &ctx->sample_staging, &ctx->cell_start, &ctx->grid_occ, &ctx->pin_scratch, &ctx->trace_scratch, &ctx->keepmask, &ctx->cell_count,
                    &ctx->disp.inbox[0], &ctx->disp.inbox[1], &ctx->disp.ctrl, &ctx->disp.occ_all, &ctx->disp.grids_dev[0], &ctx->disp.grids_dev[1],
                    &ctx->disp.keepbits, &ctx->disp.block_cnt, &ctx->disp.owner_map};
  for (int p = 0; p < ctx->disp.n_peers; ++p)
    if (ctx->disp.peer_mapped[p]) {
      for (int b = 0; b < 2; ++b) if (ctx->disp.peer_inbox[b][p]) cudaIpcCloseMemHandle(ctx->disp.peer_inbox[b][p]);
      if (ctx->disp.peer_ctrl[p]) cudaIpcCloseMemHandle(ctx->disp.peer_ctrl[p]);
    }
  for (int b = 0; b < 2; ++b) if (ctx->disp.grids_host[b]) cudaFreeHost(ctx->disp.grids_host[b]);
  for (void *q : ctx->disp.shared_opened) cudaIpcCloseMemHandle(q);
  for (void *q : ctx->disp.shared_owned) cudaFree(q);
  if (ctx->disp.ev_src) cudaEventDestroy(ctx->disp.ev_src);
  if (ctx->pair_count_host) cudaFreeHost(ctx->pair_count_host);
  if (ctx->sample_stats_host) cudaFreeHost(ctx->sample_stats_host);
  if (ctx->pin_host) cudaFreeHost(ctx->pin_host);
  if (ctx->ev_hint) cudaEventDestroy(ctx->ev_hint);
  for (auto &g : ctx->ring) if (g.done) cudaEventDestroy(g.done);
  for (auto &ps : ctx->push_streams) if (ps) { cudaStreamSynchronize(ps); cudaStreamDestroy(ps); }
  if (ctx->push_kernel_stream) { cudaStreamSynchronize(ctx->push_kernel_stream); cudaStreamDestroy(ctx->push_kernel_stream); }
  for (auto &ev : ctx->push_ev) if (ev) cudaEventDestroy(ev);
  for (int b = 0; b < 2; ++b) {
    for (int p = 0; p < ctx->n_peers; ++p) {
      if (p == ctx->peer_self) continue;
      if (ctx->peer_staging[b][p]) cudaIpcCloseMemHandle(ctx->peer_staging[b][p]);
      if (ctx->peer_free[b][p]) cudaEventDestroy(ctx->peer_free[b][p]);
      if (ctx->peer_pushed[b][p]) cudaEventDestroy(ctx->peer_pushed[b][p]);
    }
    if (ctx->ev_free[b]) cudaEventDestroy(ctx->ev_free[b]);
    if (ctx->ev_pushed[b]) cudaEventDestroy(ctx->ev_pushed[b]);
  }
  for (DevBuf *b : bufs) b->release();
  for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (auto &ev : ctx->pipe_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->copy_in) { cudaStreamSynchronize(ctx->copy_in); cudaStreamDestroy(ctx->copy_in); }
  if (ctx->copy_out) { cudaStreamSynchronize(ctx->copy_out); cudaStreamDestroy(ctx->copy_out); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return GVPM_OK;
}

const char *gvpm_last_error(const gvpm_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gvpm_sync(gvpm_ctx *ctx) {
  if (!ctx) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  return pending_check(ctx, true);   // an asynchronous gather whose pair list overflowed is re-run here
}

void *gvpm_stream(gvpm_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int gvpm_set_medium(gvpm_ctx *ctx, const gvpm_medium *m) {
  if (!ctx || !m) return GVPM_ERR_INVALID;
  const float st0 = m->sigma_s[0] + m->sigma_a[0];
  for (int i = 1; i < 3; ++i)
    if (m->sigma_s[i] + m->sigma_a[i] != st0)
      // the reference aborts the same way: homogeneous.cpp:188-201
      return fail(ctx, GVPM_ERR_UNSUPPORTED, "sigma_t must be equal across channels (balance strategy)");
  if (m->phase_type != GVPM_PHASE_ISOTROPIC && m->phase_type != GVPM_PHASE_HG)
    return fail(ctx, GVPM_ERR_UNSUPPORTED, "phase function must be isotropic or hg");
  ctx->medium = *m;
  ctx->have_medium = true;
  ++ctx->state_gen;
  return GVPM_OK;
}

int gvpm_set_config(gvpm_ctx *ctx, const gvpm_config *c) {
  if (!ctx || !c) return GVPM_ERR_INVALID;
  if (c->use_shift_null && !c->kernel_3d && !c->sppm_primal)
    // gvpm_struct.h:305-308: "Not possible to shift null without using 3D kernel"
    return fail(ctx, GVPM_ERR_UNSUPPORTED, "useShiftNull requires the 3D kernel");
  ctx->cfg = *c;
  ctx->have_cfg = true;
  ++ctx->state_gen;
  return GVPM_OK;
}

int gvpm_set_occluders(gvpm_ctx *ctx, const float *tri_xyz, size_t n_tri) {
  if (!ctx || (n_tri && !tri_xyz)) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  ctx->n_tri = (uint32_t)n_tri;
  ++ctx->state_gen;
  if (n_tri == 0) return GVPM_OK;
  std::vector<float> planes(4 * n_tri), aux(2 * n_tri);
  for (size_t t = 0; t < n_tri; ++t) {
    const float *p = tri_xyz + 9 * t;
    double e1[3], e2[3], n[3];
    for (int a = 0; a < 3; ++a) { e1[a] = (double)p[3 + a] - p[a]; e2[a] = (double)p[6 + a] - p[a]; }
    n[0] = e1[1] * e2[2] - e1[2] * e2[1];
    n[1] = e1[2] * e2[0] - e1[0] * e2[2];
    n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    {
      // rounding slack of the strict ray-triangle parameter (bre_device.cuh occluded()): ~16 eps * |e1||e2|/|e1 x e2|
      // per unit of |o - p0| / |n.d|; aux = (that coefficient, largest |vertex coordinate|)
      const double l1 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]),
                   l2 = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
      double vmag = 0;
      for (int a = 0; a < 9; ++a) vmag = std::max(vmag, (double)std::fabs(p[a]));
      aux[2 * t] = (float)(len > 0 ? 2e-6 * (l1 * l2 / len) : 0.0);
      aux[2 * t + 1] = (float)vmag;
    }
    if (len > 0) {
      for (int a = 0; a < 3; ++a) n[a] /= len;
      planes[4 * t] = (float)n[0]; planes[4 * t + 1] = (float)n[1]; planes[4 * t + 2] = (float)n[2];
      planes[4 * t + 3] = (float)(-(n[0] * p[0] + n[1] * p[1] + n[2] * p[2]));
    } else {
      planes[4 * t] = planes[4 * t + 1] = planes[4 * t + 2] = planes[4 * t + 3] = 0.f;  // never culled
    }
  }
  CK(ctx->tri.reserve(9 * n_tri * sizeof(float)));
  CK(ctx->tri_plane.reserve(4 * n_tri * sizeof(float)));
  CK(ctx->tri_aux.reserve(2 * n_tri * sizeof(float)));
  CK(cudaMemcpyAsync(ctx->tri_aux.p, aux.data(), 2 * n_tri * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tri.p, tri_xyz, 9 * n_tri * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tri_plane.p, planes.data(), 4 * n_tri * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

int gvpm_photon_staging(gvpm_ctx *ctx, size_t n, void **dev, size_t *bytes) {
  if (!ctx || n > 0xfffffff0u) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  PhotonLayout L(n);
  if (ctx->ph_staging.pinned && (L.bytes ? L.bytes : 256) > ctx->ph_staging.cap)
    return fail(ctx, GVPM_ERR_INVALID, "the photon staging buffers were exported to peers (gvpm_peer_export) and cannot grow: "
                                        "size them for the largest iteration before exporting");
  CK(ctx->ph_staging.reserve(L.bytes ? L.bytes : 256));
  ctx->n_photons = (uint32_t)n;
  ctx->photons_loaded = true;
  ctx->photons_direct = false;
  ctx->aos_override = nullptr;
  ctx->region_cap = 0;
  ctx->region_count = nullptr;
  ctx->built = false;
  ++ctx->state_gen;
  if (dev) *dev = ctx->ph_staging.p;
  if (bytes) *bytes = L.bytes;
  return GVPM_OK;
}

int gvpm_photon_staging_select(gvpm_ctx *ctx, int which) {
  if (!ctx || (which != 0 && which != 1)) return GVPM_ERR_INVALID;
  if (which != ctx->ph_staging_sel) {
    std::swap(ctx->ph_staging, ctx->ph_staging_alt);
    ctx->ph_staging_sel = which;
    ctx->built = false;
    ++ctx->state_gen;
  }
  return GVPM_OK;
}

int gvpm_photon_staging_layout(size_t n, size_t field_offset[13], size_t field_elem_bytes[13]) {
  PhotonLayout L(n);
  const size_t elt[13] = {12, 12, 12, 12, 12, 12, 12, 4, 4, 4, 1, 1, 4};
  for (int i = 0; i < 13; ++i) {
    if (field_offset) field_offset[i] = L.off[i];
    if (field_elem_bytes) field_elem_bytes[i] = elt[i];
  }
  return GVPM_OK;
}

int gvpm_upload_photons_slice(gvpm_ctx *ctx, const gvpm_photon_soa *p, size_t n_total, size_t begin, size_t count,
                              void *stream) {
  if (!ctx || (count && !p) || begin + count > n_total) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->n_photons != n_total || !ctx->ph_staging.p || ctx->ph_staging.cap < PhotonLayout(n_total).bytes)
    return fail(ctx, GVPM_ERR_INVALID, "gvpm_photon_staging(n_total) must be called first for the selected buffer");
  if (count == 0) return GVPM_OK;
  PhotonLayout L(n_total);
  char *b = (char *)ctx->ph_staging.p;
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const void *src[13] = {p->pos, p->flux, p->parent_pos, p->pred_pos, p->parent_n, p->prefix_flux, p->parent_albedo,
                         p->parent_pdf, p->edge_pdf, p->rr_weight, p->parent_type, p->depth, p->path_id};
  const size_t elt[13] = {12, 12, 12, 12, 12, 12, 12, 4, 4, 4, 1, 1, 4};
  for (int i = 0; i < 13; ++i) {
    if (!src[i]) return fail(ctx, GVPM_ERR_INVALID, "null array in gvpm_photon_soa");
    CK(cudaMemcpyAsync(b + L.off[i] + begin * elt[i], src[i], count * elt[i], cudaMemcpyHostToDevice, st));
  }
  return GVPM_OK;
}

// ---- peer exchange of photon slices over NVLink copy engines ------------------------------------------------------
struct IpcBlob {  // what gvpm_peer_export writes (GVPM_PEER_BLOB_BYTES)
  cudaIpcMemHandle_t staging[2];
  cudaIpcEventHandle_t free_ev[2], pushed_ev[2];
};
static_assert(sizeof(IpcBlob) == GVPM_PEER_BLOB_BYTES, "blob size");

int gvpm_peer_export(gvpm_ctx *ctx, void *blob) {
  if (!ctx || !blob) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!ctx->ph_staging.p || !ctx->ph_staging_alt.p)
    return fail(ctx, GVPM_ERR_INVALID, "size both photon staging buffers (gvpm_photon_staging) before exporting them");
  IpcBlob *B = (IpcBlob *)blob;
  ctx->ph_staging.pinned = ctx->ph_staging_alt.pinned = true;
  for (int b = 0; b < 2; ++b) {
    CK(cudaIpcGetMemHandle(&B->staging[b], ctx->staging_ptr(b)));
    CK(cudaIpcGetEventHandle(&B->free_ev[b], ctx->ev_free[b]));
    CK(cudaIpcGetEventHandle(&B->pushed_ev[b], ctx->ev_pushed[b]));
    // make both events "recorded" so that a first wait on them is well defined
    CK(cudaEventRecord(ctx->ev_free[b], ctx->stream));
    CK(cudaEventRecord(ctx->ev_pushed[b], ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

int gvpm_peer_connect(gvpm_ctx *ctx, const void *blobs, int n_peers, int self_index) {
  if (!ctx || !blobs || n_peers < 1 || self_index < 0 || self_index >= n_peers) return GVPM_ERR_INVALID;
  if (ctx->n_peers) return fail(ctx, GVPM_ERR_INVALID, "peers already connected");
  cudaSetDevice(ctx->device);
  const IpcBlob *B = (const IpcBlob *)blobs;
  for (int b = 0; b < 2; ++b) {
    ctx->peer_staging[b].assign(n_peers, nullptr);
    ctx->peer_free[b].assign(n_peers, nullptr);
    ctx->peer_pushed[b].assign(n_peers, nullptr);
  }
  ctx->n_peers = n_peers;
  ctx->peer_self = self_index;
  for (int p = 0; p < n_peers; ++p) {
    if (p == self_index) continue;
    for (int b = 0; b < 2; ++b) {
      CK(cudaIpcOpenMemHandle(&ctx->peer_staging[b][p], B[p].staging[b], cudaIpcMemLazyEnablePeerAccess));
      CK(cudaIpcOpenEventHandle(&ctx->peer_free[b][p], B[p].free_ev[b]));
      CK(cudaIpcOpenEventHandle(&ctx->peer_pushed[b][p], B[p].pushed_ev[b]));
    }
  }
  return GVPM_OK;
}

// SM push: one kernel reads this rank's slice of the 13 field arrays once (128-bit loads) and stores every word into
// the same place of every peer's staging buffer over NVLink (peer-mapped pointers), so the 7 outgoing streams of an
// 8-GPU box are fed in parallel.  It runs on a highest-priority stream with a small grid: the copy engines top out
// near 220 GB/s per GPU on this many-small-copies pattern (13 fields x 7 peers), a few SMs' worth of stores do not.
struct PushParams {
  const char *src;
  char *dst[8];
  int n_dst;
  unsigned long long off[13];   // byte offset of each field's slice inside a staging buffer
  unsigned long long cum[14];   // prefix sums of the slice sizes in 16-byte words
};
__global__ void __launch_bounds__(512) k_peer_push(const __grid_constant__ PushParams P) {
  const unsigned long long total = P.cum[13], stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long w0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w0 < total; w0 += 4 * stride) {
    uint4 v[4];
    unsigned long long byte[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned long long w = w0 + u * stride;
      byte[u] = ~0ull;
      if (w < total) {
        int f = 0;
        while (w >= P.cum[f + 1]) ++f;
        byte[u] = P.off[f] + (w - P.cum[f]) * 16ull;
        v[u] = __ldcs((const uint4 *)(P.src + byte[u]));
      }
    }
    for (int d = 0; d < P.n_dst; ++d)
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (byte[u] != ~0ull) __stcs((uint4 *)(P.dst[d] + byte[u]), v[u]);
  }
}

int gvpm_peer_push_mode(gvpm_ctx *ctx, int sm_ctas) {
  if (!ctx || sm_ctas < 0 || sm_ctas > 1024) return GVPM_ERR_INVALID;
  ctx->push_ctas = sm_ctas;
  return GVPM_OK;
}

int gvpm_peer_push_photon_slice(gvpm_ctx *ctx, int which, size_t n_total, size_t begin, size_t count, void *after_stream) {
  if (!ctx || (which != 0 && which != 1) || begin + count > n_total) return GVPM_ERR_INVALID;
  if (!ctx->n_peers) return fail(ctx, GVPM_ERR_INVALID, "gvpm_peer_connect has not been called");
  cudaSetDevice(ctx->device);
  PhotonLayout L(n_total);
  const size_t elt[13] = {12, 12, 12, 12, 12, 12, 12, 4, 4, 4, 1, 1, 4};
  const char *mine = (const char *)ctx->staging_ptr(which);
  // the slice must be complete on this rank (work queued on after_stream, e.g. its H2D upload) ...
  cudaEvent_t src_ready = ctx->push_ev[4];
  CK(cudaEventRecord(src_ready, after_stream ? (cudaStream_t)after_stream : ctx->stream));
  bool aligned = ctx->n_peers <= 9;
  for (int f = 0; f < 13; ++f) aligned = aligned && (begin * elt[f]) % 16 == 0 && (count * elt[f]) % 16 == 0;
  if (ctx->push_ctas > 0 && aligned && count > 0) {
    PushParams PP{};
    PP.src = mine;
    cudaStream_t ks = ctx->push_kernel_stream;
    CK(cudaStreamWaitEvent(ks, src_ready, 0));
    for (int p = 0; p < ctx->n_peers; ++p) {
      if (p == ctx->peer_self) continue;
      CK(cudaStreamWaitEvent(ks, ctx->peer_free[which][p], 0));
      PP.dst[PP.n_dst++] = (char *)ctx->peer_staging[which][p];
    }
    for (int f = 0; f < 13; ++f) {
      PP.off[f] = L.off[f] + begin * elt[f];
      PP.cum[f + 1] = PP.cum[f] + count * elt[f] / 16;
    }
    k_peer_push<<<ctx->push_ctas, 512, 0, ks>>>(PP);
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaEventRecord(ctx->ev_pushed[which], ks));
    return GVPM_OK;
  }
  int si = 0;
  for (int p = 0; p < ctx->n_peers; ++p) {
    if (p == ctx->peer_self) continue;
    cudaStream_t ps = ctx->push_streams[si++ & 3];
    CK(cudaStreamWaitEvent(ps, src_ready, 0));
    // ... and the peer must have consumed what its buffer held (its last build from it)
    CK(cudaStreamWaitEvent(ps, ctx->peer_free[which][p], 0));
    char *dst = (char *)ctx->peer_staging[which][p];
    for (int f = 0; f < 13; ++f)
      CK(cudaMemcpyAsync(dst + L.off[f] + begin * elt[f], mine + L.off[f] + begin * elt[f], count * elt[f],
                         cudaMemcpyDeviceToDevice, ps));
  }
  // ev_pushed[which] = all push streams done: funnel them through stream 0
  for (int i = 1; i < 4; ++i) {
    CK(cudaEventRecord(ctx->push_ev[i], ctx->push_streams[i]));
    CK(cudaStreamWaitEvent(ctx->push_streams[0], ctx->push_ev[i], 0));
  }
  CK(cudaEventRecord(ctx->ev_pushed[which], ctx->push_streams[0]));
  return GVPM_OK;
}

int gvpm_peer_wait_photons(gvpm_ctx *ctx, int which) {
  if (!ctx || (which != 0 && which != 1)) return GVPM_ERR_INVALID;
  if (!ctx->n_peers) return fail(ctx, GVPM_ERR_INVALID, "gvpm_peer_connect has not been called");
  cudaSetDevice(ctx->device);
  for (int p = 0; p < ctx->n_peers; ++p)
    if (p != ctx->peer_self) CK(cudaStreamWaitEvent(ctx->stream, ctx->peer_pushed[which][p], 0));
  return GVPM_OK;
}

int gvpm_upload_photons(gvpm_ctx *ctx, const gvpm_photon_soa *p, size_t n) {
  if (!ctx || (n && !p)) return GVPM_ERR_INVALID;
  int rc = gvpm_photon_staging(ctx, n, nullptr, nullptr);
  if (rc) return rc;
  if (n == 0) return GVPM_OK;
  PhotonLayout L(n);
  char *b = (char *)ctx->ph_staging.p;
  const void *src[13] = {p->pos, p->flux, p->parent_pos, p->pred_pos, p->parent_n, p->prefix_flux, p->parent_albedo,
                         p->parent_pdf, p->edge_pdf, p->rr_weight, p->parent_type, p->depth, p->path_id};
  const size_t sz[13] = {12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 4 * n, 4 * n, 4 * n, n, n, 4 * n};
  for (int i = 0; i < 13; ++i) {
    if (!src[i]) return fail(ctx, GVPM_ERR_INVALID, "null array in gvpm_photon_soa");
    CK(cudaMemcpyAsync(b + L.off[i], src[i], sz[i], cudaMemcpyHostToDevice, ctx->stream));
  }
  return GVPM_OK;
}

int gvpm_build_points(gvpm_ctx *ctx, float radius) {
  if (!ctx) return GVPM_ERR_INVALID;
  if (!ctx->photons_loaded) return fail(ctx, GVPM_ERR_INVALID, "no photons uploaded");
  if (!(radius > 0.f)) return fail(ctx, GVPM_ERR_INVALID, "radius must be positive");
  cudaSetDevice(ctx->device);
  const uint32_t n = ctx->n_photons;
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[0], st));
  Tree T{};
  T.n = n;
  // level sizes
  uint32_t cnt = (n + 31) / 32, total = 0;
  int levels = 0;
  if (n > 0) {
    for (;;) {
      T.cnt[levels] = cnt;
      T.off[levels] = total;
      total += cnt;
      ++levels;
      if (cnt <= 32) break;
      if (levels >= GVPM_MAX_LEVELS) return fail(ctx, GVPM_ERR_INVALID, "too many photons");
      cnt = (cnt + 31) / 32;
    }
  }
  T.levels = levels;
  if (n > 0) {
    CK(ctx->keys_in.reserve(8 * (size_t)n));
    CK(ctx->keys_out.reserve(8 * (size_t)n));
    CK(ctx->vals_in.reserve(4 * (size_t)n));
    CK(ctx->vals_out.reserve(4 * (size_t)n));
    const size_t tb = sort_temp_bytes(n);
    CK(ctx->sort_temp.reserve(tb));
    CK(ctx->planes.reserve(16 * (size_t)n));
    CK(ctx->orig.reserve(4 * (size_t)n));
    CK(ctx->box_lo.reserve(16 * (size_t)total));
    CK(ctx->box_hi.reserve(16 * (size_t)total));
    const bool direct = ctx->photons_direct;
    PhotonStaging S{};
    if (!direct) S = photon_staging_ptrs(ctx->ph_staging.p, n);
    const float *pos = direct ? (const float *)ctx->records() : S.pos;
    const uint32_t stride = direct ? 32u : 3u;
    CK(ctx->bounds_partial.reserve((size_t)bounds_blocks(n) * 6 * sizeof(float)));
    launch_bounds(pos, n, ctx->bounds_partial.as<float>(), ctx->bounds.as<float>(), st, stride);
    launch_morton(pos, n, ctx->bounds.as<float>(), ctx->keys_in.as<uint32_t>(), ctx->vals_in.as<uint32_t>(), st, stride);
    CK(run_sort(ctx->sort_temp.p, ctx->sort_temp.cap, ctx->keys_in.as<uint32_t>(), ctx->keys_out.as<uint32_t>(),
                ctx->vals_in.as<uint32_t>(), ctx->vals_out.as<uint32_t>(), n, st));
    if (!direct) CK(ctx->aos.reserve(128 * (size_t)n));
    launch_pack_sorted(S, ctx->records(), ctx->vals_out.as<uint32_t>(), n, ctx->planes.as<float4>(),
                       ctx->orig.as<uint32_t>(), st, direct);
    CK(cudaEventRecord(ctx->ev_free[ctx->ph_staging_sel], st));   // last read of the staging buffer
    float4 *lo = ctx->box_lo.as<float4>(), *hi = ctx->box_hi.as<float4>();
    launch_leaf_boxes(ctx->planes.as<float4>(), n, T.cnt[0], radius, lo, hi, st);
    for (int l = 1; l < levels; ++l)
      launch_level_boxes(lo + T.off[l - 1], hi + T.off[l - 1], T.cnt[l - 1], T.cnt[l], lo + T.off[l], hi + T.off[l], st);
    ctx->launches += 6 + (levels - 1) + 4;  // + the radix sort's own passes (library, ~4 launches)
    CK(cudaGetLastError());
  } else {
    CK(cudaMemsetAsync(ctx->bounds.p, 0, 7 * sizeof(float), st));
  }
  T.lo = ctx->box_lo.as<float4>();
  T.hi = ctx->box_hi.as<float4>();
  ctx->tree = T;
  ctx->radius = radius;
  ctx->built = true;
  ctx->accel = gvpm_ctx::ACCEL_BVH;
  ++ctx->state_gen;
  ctx->pruned = false;
  ctx->n_kept = n;
  CK(cudaEventRecord(ctx->ev[1], st));
  ctx->timed_build = true;
  return GVPM_OK;
}

// Same hierarchy, over the photons the uploaded rays can reach only (tree_build.cu, "ray-region pruning").
static int build_pruned_bvh(gvpm_ctx *ctx, float radius, uint32_t *n_kept) {
  cudaSetDevice(ctx->device);
  if (ctx->photons_direct) {   // records traced in place: no staged SoA to prune from, build over all of them
    int rc = gvpm_build_points(ctx, radius);
    if (rc == GVPM_OK) { ctx->pruned = false; if (n_kept) *n_kept = ctx->n_photons; }
    return rc;
  }
  const uint32_t n = ctx->n_photons;
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[0], st));
  const size_t gridOff = 64, maskOff = align256(gridOff + ray_grid_bytes()), keepOff = align256(maskOff + ray_mask_bytes());
  CK(ctx->ray_region.reserve(keepOff + 4 * ((size_t)n / 32 + 1)));
  char *rr = ctx->ray_region.as<char>();
  float *box = (float *)rr;
  uint32_t *counter = (uint32_t *)(rr + 32);
  uint32_t *mask = (uint32_t *)(rr + maskOff), *keepmask = (uint32_t *)(rr + keepOff);
  const float boxInit[8] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, 0.f, 0.f};
  CK(cudaMemcpyAsync(box, boxInit, sizeof(boxInit), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(counter, 0, 32, st));
  CK(cudaMemsetAsync(mask, 0, ray_mask_bytes(), st));
  launch_ray_region(ctx->rays.as<float4>(), ctx->n_rays, radius, box, rr + gridOff, mask, ctx->bounds.as<float>(),
                    ctx->sm_count, st);
  ctx->launches += ctx->n_rays ? 3 : 1;
  uint32_t kept = 0;
  if (n > 0) {
    CK(ctx->keys_in.reserve(8 * (size_t)n));
    CK(ctx->keys_out.reserve(8 * (size_t)n));
    CK(ctx->vals_in.reserve(4 * (size_t)n));
    CK(ctx->vals_out.reserve(4 * (size_t)n));
    CK(ctx->sort_temp.reserve(sort_temp_bytes(n)));
    CK(ctx->planes.reserve(16 * (size_t)n));
    CK(ctx->orig.reserve(4 * (size_t)n));
    CK(ctx->aos.reserve(128 * (size_t)n));
    PhotonStaging S = photon_staging_ptrs(ctx->ph_staging.p, n);
    // kept indices in photon order (deterministic): per-CTA counts, exclusive scan, ordered compaction
    const uint32_t nb = (n + 255) / 256;
    CK(ctx->keepmask.reserve(4 * ((size_t)nb + 4)));
    uint32_t *block_kept = ctx->keepmask.as<uint32_t>();
    launch_keep_pruned(S.pos, n, rr + gridOff, mask, keepmask, block_kept, st);
    launch_scan_u32(block_kept, nb, block_kept + nb, st);
    launch_compact_kept(nullptr, keepmask, block_kept, n, nullptr, ctx->vals_in.as<uint32_t>(), st, n, counter);
    ctx->launches += 3;
    // the number of kept photons sizes the sort and the hierarchy levels: one 4-byte read-back
    CK(cudaMemcpyAsync(ctx->pair_count_host, block_kept + nb, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    kept = *(const uint32_t *)ctx->pair_count_host;
  }
  Tree T{};
  T.n = kept;
  uint32_t cnt = (kept + 31) / 32, total = 0;
  int levels = 0;
  if (kept > 0) {
    for (;;) {
      T.cnt[levels] = cnt;
      T.off[levels] = total;
      total += cnt;
      ++levels;
      if (cnt <= 32) break;
      if (levels >= GVPM_MAX_LEVELS) return fail(ctx, GVPM_ERR_INVALID, "too many photons");
      cnt = (cnt + 31) / 32;
    }
  }
  T.levels = levels;
  if (kept > 0) {
    CK(ctx->box_lo.reserve(16 * (size_t)total));
    CK(ctx->box_hi.reserve(16 * (size_t)total));
    PhotonStaging S = photon_staging_ptrs(ctx->ph_staging.p, n);
    launch_keys_kept(S.pos, ctx->vals_in.as<uint32_t>(), kept, rr + gridOff, ctx->keys_in.as<uint32_t>(), st);
    CK(run_sort(ctx->sort_temp.p, ctx->sort_temp.cap, ctx->keys_in.as<uint32_t>(), ctx->keys_out.as<uint32_t>(),
                ctx->vals_in.as<uint32_t>(), ctx->vals_out.as<uint32_t>(), kept, st));
    launch_pack_pruned(S, n, keepmask, ctx->vals_out.as<uint32_t>(), kept, ctx->aos.as<float4>(),
                       ctx->planes.as<float4>(), ctx->orig.as<uint32_t>(), st);
    float4 *lo = ctx->box_lo.as<float4>(), *hi = ctx->box_hi.as<float4>();
    launch_leaf_boxes(ctx->planes.as<float4>(), kept, T.cnt[0], radius, lo, hi, st);
    for (int l = 1; l < levels; ++l)
      launch_level_boxes(lo + T.off[l - 1], hi + T.off[l - 1], T.cnt[l - 1], T.cnt[l], lo + T.off[l], hi + T.off[l], st);
    ctx->launches += 4 + (levels - 1) + 4;
    CK(cudaGetLastError());
  }
  if (n > 0) CK(cudaEventRecord(ctx->ev_free[ctx->ph_staging_sel], st));   // last read of the staging buffer
  T.lo = ctx->box_lo.as<float4>();
  T.hi = ctx->box_hi.as<float4>();
  ctx->tree = T;
  ctx->radius = radius;
  ctx->built = true;
  ctx->accel = gvpm_ctx::ACCEL_BVH;
  ++ctx->state_gen;
  ctx->pruned = true;
  ctx->pruned_for_gen = ctx->rays_gen;
  ctx->n_kept = kept;
  if (n_kept) *n_kept = kept;
  CK(cudaEventRecord(ctx->ev[1], st));
  ctx->timed_build = true;
  return GVPM_OK;
}


// ---- frustum grid (gvpm_device.cuh FrustumGrid) -------------------------------------------------------------------------
static float decode_max_key(unsigned key) {
  const unsigned b = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
  float f;
  memcpy(&f, &b, 4);
  return f;
}
// Are the uploaded rays concurrent (do their lines meet in one point)?  Two reductions over the packed rays and one
// read-back, cached until other rays are uploaded.
static int analyse_rays(gvpm_ctx *ctx) {
  if (ctx->pin_gen == ctx->rays_gen) return GVPM_OK;
  gvpm_ctx::RayFit &F = ctx->pin;
  F.concurrent = false;
  ctx->pin_gen = ctx->rays_gen;
  const uint32_t n = ctx->n_rays;
  if (n == 0) return GVPM_OK;
  CK(ctx->pin_scratch.reserve(128 + (size_t)pinhole_blocks(n) * 16 * sizeof(double)));
  char *ps = ctx->pin_scratch.as<char>();
  launch_pinhole_fit(ctx->rays.as<float4>(), n, (double *)(ps + 128), (float *)ps, (unsigned *)(ps + 64), ctx->stream,
                     ctx->have_view_dir ? ctx->view_dir : nullptr);
  ctx->launches += 3;
  CK(cudaMemcpyAsync(ctx->pin_host, ps, 96, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const float *fit = ctx->pin_host;
  const unsigned *st = (const unsigned *)(ctx->pin_host + 16);
  F.count = fit[12];
  if (!(F.count >= 1.f) || !(fit[13] > 0.f)) return GVPM_OK;   // no active ray, or parallel rays (no common point)
  for (int k = 0; k < 3; ++k) { F.C[k] = fit[k]; F.m[k] = fit[3 + k]; F.u[k] = fit[6 + k]; F.v[k] = fit[9 + k]; }
  F.delta = decode_max_key(st[0]);
  F.cosmin = -decode_max_key(st[1]);
  F.xmin = -decode_max_key(st[2]); F.xmax = decode_max_key(st[3]);
  F.ymin = -decode_max_key(st[4]); F.ymax = decode_max_key(st[5]);
  F.concurrent = std::isfinite(F.delta) && F.cosmin > 0.35f && std::isfinite(F.xmin) && std::isfinite(F.xmax) &&
                 std::isfinite(F.ymin) && std::isfinite(F.ymax) && F.xmin <= F.xmax && F.ymin <= F.ymax;
  return GVPM_OK;
}

// The grid of a concurrent ray set (RayFit) for search radius `radius`: host arithmetic only, so that a rank that SENDS
// photons (dispatch.cu) derives the receiver's grid from the receiver's RayFit exactly as the receiver does.
static FrustumGrid make_frustum_grid(const gvpm_ctx::RayFit &F, float radius, bool parity_split) {
  FrustumGrid G{};
  for (int k = 0; k < 3; ++k) { G.C[k] = F.C[k]; G.m[k] = F.m[k]; G.u[k] = F.u[k]; G.v[k] = F.v[k]; }
  G.xmin = F.xmin; G.xmax = F.xmax; G.ymin = F.ymin; G.ymax = F.ymax;
  G.pad_r = radius + F.delta;
  // about one ray per class-0 cell, square cells, at most 4 M of them; every class doubles the cell edge
  const double ex = std::max((double)F.xmax - F.xmin, 1e-9), ey = std::max((double)F.ymax - F.ymin, 1e-9);
  const double target = std::min(std::max((double)F.count, 64.0), 4.0e6);
  double cell = std::sqrt(ex * ey / target);
  cell = std::max(cell, std::max(ex, ey) / 4096.0);
  G.gx0 = F.xmin;
  G.gy0 = F.ymin;
  uint32_t total = 0;
  int classes = 0;
  for (int c = 0; c < GVPM_GRID_CLASSES; ++c) {
    G.csize[c] = (float)(cell * std::pow(2.0, (double)c));
    G.nx[c] = (uint32_t)std::floor(ex / G.csize[c]) + 2u;
    G.ny[c] = (uint32_t)std::floor(ey / G.csize[c]) + 2u;
    G.base[c] = total;
    total += G.nx[c] * G.ny[c];
    classes = c + 1;
    if (G.nx[c] <= 2 && G.ny[c] <= 2) break;
  }
  for (int c = classes; c < GVPM_GRID_CLASSES; ++c) { G.csize[c] = G.csize[classes - 1]; G.nx[c] = G.ny[c] = 1; G.base[c] = total; }
  G.classes = classes;
  G.n_cells = total;
  G.parity_split = parity_split ? 1 : 0;
  return G;
}

static int build_frustum(gvpm_ctx *ctx, float radius, uint32_t *n_kept) {
  cudaSetDevice(ctx->device);
  const uint32_t n = ctx->n_photons;
  cudaStream_t st = ctx->stream;
  const gvpm_ctx::RayFit &F = ctx->pin;
  CK(cudaEventRecord(ctx->ev[0], st));
  const FrustumGrid G = make_frustum_grid(F, radius, ctx->have_cfg && ctx->cfg.path_set && !ctx->cfg.sppm_primal);
  const uint32_t total = G.n_cells;
  const uint32_t n_keys = (G.parity_split ? 2u : 1u) * total + 2;   // + NEAR + DROP
  int bits = 1;
  while ((1ull << bits) < (unsigned long long)n_keys) ++bits;
  CK(ctx->cell_start.reserve(((size_t)n_keys + 1) * 4));
  // how much of the set did the previous build keep?  (read without blocking: sizes this build's sort)
  if (ctx->hint_pending && cudaEventQuery(ctx->ev_hint) == cudaSuccess) {
    ctx->hint_pending = false;
    const uint32_t keptPrev = *(const uint32_t *)(ctx->pin_host + 30);
    if (ctx->hint_n) { ctx->kept_fraction_hint = (double)keptPrev / (double)ctx->hint_n; ctx->kept_hint_valid = true; }
  }
  uint32_t m = n;   // sorted entries
  if (n > 0) {
    const uint32_t nb = (n + 255) / 256;
    CK(ctx->keys_in.reserve(8 * (size_t)n));
    CK(ctx->keys_out.reserve(8 * (size_t)n));
    CK(ctx->vals_in.reserve(4 * (size_t)n));
    CK(ctx->vals_out.reserve(4 * (size_t)n));
    CK(ctx->sort_temp.reserve(sort_temp_bytes(n)));
    CK(ctx->planes.reserve(16 * (size_t)n));
    CK(ctx->orig.reserve(4 * (size_t)n));
    CK(ctx->keepmask.reserve(4 * ((size_t)n / 32 + 1) + 4 * ((size_t)nb + 4)));
    const bool direct = ctx->photons_direct;
    if (!direct) CK(ctx->aos.reserve(128 * (size_t)n));
    CK(ctx->grid_occ.reserve(frustum_occ_bytes() + cell_starts_scratch_bytes(n_keys)));
    uint32_t *keepmask = ctx->keepmask.as<uint32_t>(), *block_kept = keepmask + (n / 32 + 1);
    PhotonStaging S{};
    if (!direct) S = photon_staging_ptrs(ctx->ph_staging.p, n);
    // Counting sort (default): the key pass counts the photons of every cell and gives each its rank in the cell; the
    // exclusive scan of the counters IS cell_start; the packing pass (or a scatter pass over records that exist already)
    // puts position plane entry and index at cell_start[key] + rank.  One streaming pass over the keys less than the
    // radix sort takes for each of its three digits, no gather of the sorted plane, no cell-start search, and nothing
    // to size from a count: dropped photons and the empty part of a dispatched inbox are simply never written.  The
    // order inside a cell follows the atomics (the gather's float atomics are unordered anyway).
    const bool counting = !ctx->radix_frustum;
    uint32_t *cell_count = nullptr;
    if (counting) {
      CK(ctx->cell_count.reserve(((size_t)n_keys + 2) * 4));
      cell_count = ctx->cell_count.as<uint32_t>();
      CK(cudaMemsetAsync(cell_count, 0, ((size_t)n_keys + 2) * 4, st));
      CK(ctx->sort_temp.reserve(std::max(sort_temp_bytes(n), cell_scan_temp_bytes(n_keys))));
    }
    if (direct)   // position = first three floats of the record, path parity = bit 10 of its meta word
      launch_frustum_keys(ctx->rays.as<float4>(), ctx->n_rays, (const float *)ctx->records(), 32u, (const uint32_t *)ctx->records() + 3, 32u, 10u, n,
                          G, ctx->grid_occ.as<uint32_t>(), ctx->keys_in.as<uint32_t>(), ctx->vals_in.as<uint32_t>(),
                          ctx->bounds.as<unsigned>() + 6, keepmask, block_kept, ctx->sm_count, st, ctx->region_cap, ctx->region_count,
                          cell_count);
    else
      launch_frustum_keys(ctx->rays.as<float4>(), ctx->n_rays, S.pos, 3u, S.path_id, 1u, 0u, n, G, ctx->grid_occ.as<uint32_t>(),
                          ctx->keys_in.as<uint32_t>(), ctx->vals_in.as<uint32_t>(), ctx->bounds.as<unsigned>() + 6, keepmask,
                          block_kept, ctx->sm_count, st, 0, nullptr, cell_count);
    const uint32_t *sortedKeys = nullptr, *sortedVals = nullptr;
    uint32_t *ovf = block_kept + nb + 1;   // set when a bounded compaction meets more kept photons than it was sized for
    CK(cudaMemsetAsync(ovf, 0, 4, st));
    if (counting) {
      CK(launch_cell_scan(ctx->sort_temp.p, ctx->sort_temp.cap, cell_count, ctx->cell_start.as<uint32_t>(), n_keys, st));
      launch_pack_scatter(S, n, keepmask, ctx->keys_in.as<uint32_t>(), ctx->vals_in.as<uint32_t>(), ctx->cell_start.as<uint32_t>(),
                          ctx->records(), ctx->planes.as<float4>(), ctx->orig.as<uint32_t>(), st, direct,
                          ctx->cell_start.as<uint32_t>() + n_keys - 1);
      sortedKeys = sortedVals = nullptr;
      ctx->launches += 4;
    } else if (ctx->kept_fraction_hint < 0.5) {
      // sharded image: most photons are out of this rank's reach.  Compact the kept (key, index) pairs in index order
      // (deterministic) and sort those only.  Their number sizes the sort: taken from the previous build's count (read
      // back asynchronously) with a 25 % margin, the slots past the real count filled with DROP keys - no host round
      // trip in the step.  Without a previous count (first build, or after an overflow) the exact count is read back.
      launch_scan_u32(block_kept, nb, block_kept + nb, st);
      const bool bounded = ctx->kept_hint_valid && !ctx->force_exact_build;
      if (bounded) {
        const uint32_t DROP = n_keys - 1;
        m = (uint32_t)std::min<double>((double)n, ctx->kept_fraction_hint * 1.25 * (double)n + 65536.0);
        launch_fill_u32(ctx->keys_out.as<uint32_t>(), m, DROP, st);
        CK(cudaMemsetAsync(ctx->vals_out.p, 0, 4 * (size_t)m, st));
        launch_compact_kept(ctx->keys_in.as<uint32_t>(), keepmask, block_kept, n, ctx->keys_out.as<uint32_t>(),
                            ctx->vals_out.as<uint32_t>(), st, m, ovf);
        CK(cudaMemcpyAsync(ctx->pin_host + 30, block_kept + nb, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev_hint, st));
        ctx->hint_pending = true;
        ctx->hint_n = n;
        ctx->launches += 1;
      } else {
        launch_compact_kept(ctx->keys_in.as<uint32_t>(), keepmask, block_kept, n, ctx->keys_out.as<uint32_t>(),
                            ctx->vals_out.as<uint32_t>(), st, n, ovf);
        CK(cudaMemcpyAsync(ctx->pair_count_host, block_kept + nb, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        m = *(const uint32_t *)ctx->pair_count_host;
        ctx->kept_fraction_hint = (double)m / (double)n;
        ctx->kept_hint_valid = true;
        ctx->force_exact_build = false;
        ctx->hint_pending = false;
      }
      if (m) CK(run_sort_bits(ctx->sort_temp.p, ctx->sort_temp.cap, ctx->keys_out.as<uint32_t>(), ctx->keys_in.as<uint32_t>(),
                              ctx->vals_out.as<uint32_t>(), ctx->vals_in.as<uint32_t>(), m, bits, st));
      sortedKeys = ctx->keys_in.as<uint32_t>();
      sortedVals = ctx->vals_in.as<uint32_t>();
      ctx->launches += 2;
    } else {
      CK(run_sort_bits(ctx->sort_temp.p, ctx->sort_temp.cap, ctx->keys_in.as<uint32_t>(), ctx->keys_out.as<uint32_t>(),
                       ctx->vals_in.as<uint32_t>(), ctx->vals_out.as<uint32_t>(), n, bits, st));
      sortedKeys = ctx->keys_out.as<uint32_t>();
      sortedVals = ctx->vals_out.as<uint32_t>();
    }
    ctx->build_ovf = ovf;
    if (!counting) {
      launch_cell_starts(sortedKeys, m, n_keys, ctx->cell_start.as<uint32_t>(),
                         ctx->grid_occ.as<char>() + frustum_occ_bytes(), ctx->sm_count, st);
      launch_pack_sorted_kept(S, n, keepmask, sortedVals, m, ctx->records(), ctx->planes.as<float4>(),
                              ctx->orig.as<uint32_t>(), st, direct);
    }
    CK(cudaEventRecord(ctx->ev_free[ctx->ph_staging_sel], st));   // last read of the staging buffer
    if (!counting && m == n) {   // kept count of this build -> hint of the next one (no blocking)
      CK(cudaMemcpyAsync(ctx->pin_host + 30, ctx->cell_start.as<uint32_t>() + n_keys - 1, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(ctx->ev_hint, st));
      ctx->hint_pending = true;
      ctx->hint_n = n;
    }
    ctx->launches += 9 + 4;
    CK(cudaGetLastError());
  } else {
    CK(cudaMemsetAsync(ctx->bounds.p, 0, 7 * sizeof(float), st));
    CK(cudaMemsetAsync(ctx->cell_start.p, 0, ((size_t)n_keys + 1) * 4, st));
  }
  Tree T{};
  T.n = m;
  ctx->tree = T;
  ctx->grid = G;
  ctx->radius = radius;
  ctx->built = true;
  ctx->accel = gvpm_ctx::ACCEL_FRUSTUM;
  ++ctx->state_gen;
  ctx->pruned = true;
  ctx->pruned_for_gen = ctx->rays_gen;
  ctx->n_kept = n;
  CK(cudaEventRecord(ctx->ev[1], st));
  ctx->timed_build = true;
  if (n_kept) {   // photons some ray can reach = everything in front of the DROP bucket (one read-back, only on request)
    CK(cudaMemcpyAsync(ctx->pair_count_host, ctx->cell_start.as<uint32_t>() + n_keys - 1, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->n_kept = *(const uint32_t *)ctx->pair_count_host;
    *n_kept = ctx->n_kept;
  }
  return GVPM_OK;
}

int gvpm_build_points_for_rays(gvpm_ctx *ctx, float radius, uint32_t *n_kept) {
  if (!ctx) return GVPM_ERR_INVALID;
  if (!ctx->photons_loaded) return fail(ctx, GVPM_ERR_INVALID, "no photons uploaded");
  if (!ctx->rays_loaded) return fail(ctx, GVPM_ERR_INVALID, "gvpm_build_points_for_rays needs the rays first (gvpm_upload_rays)");
  if (!(radius > 0.f)) return fail(ctx, GVPM_ERR_INVALID, "radius must be positive");
  cudaSetDevice(ctx->device);
  if (!ctx->force_bvh) {
    int rc = analyse_rays(ctx);
    if (rc) return rc;
    // concurrent rays whose common point is sharp compared with the search radius: perspective grid, no hierarchy
    if (ctx->pin.concurrent && ctx->pin.delta <= 0.25f * radius) return build_frustum(ctx, radius, n_kept);
  }
  return build_pruned_bvh(ctx, radius, n_kept);
}


// ---- on-device generators (csrc/generate.cu) -----------------------------------------------------
int gvpm_box_scene_default(gvpm_box_scene *s) {
  if (!s) return GVPM_ERR_INVALID;
  memset(s, 0, sizeof(*s));
  for (int a = 0; a < 3; ++a) { s->lo[a] = 0.f; s->hi[a] = 1.f; }
  const float alb[5][3] = {{0.63f, 0.065f, 0.05f}, {0.14f, 0.45f, 0.091f}, {0.7f, 0.7f, 0.7f}, {0.7f, 0.7f, 0.7f}, {0.7f, 0.7f, 0.7f}};
  memcpy(s->face_albedo, alb, sizeof(alb));
  s->n_rects = 1;   // the shelf
  s->rect[0].y = 0.5f; s->rect[0].x0 = 0.3f; s->rect[0].x1 = 0.7f; s->rect[0].z0 = 0.4f; s->rect[0].z1 = 0.8f;
  s->rect[0].albedo[0] = s->rect[0].albedo[1] = s->rect[0].albedo[2] = 0.6f;
  s->light_y = 0.999f; s->light_x0 = 0.35f; s->light_x1 = 0.65f; s->light_z0 = 0.35f; s->light_z1 = 0.65f;
  s->light_power = 100.f;
  return GVPM_OK;
}

static int check_scene(gvpm_ctx *ctx, const gvpm_box_scene *s) {
  if (s->n_rects < 0 || s->n_rects > 4) return fail(ctx, GVPM_ERR_INVALID, "gvpm_box_scene: n_rects must be in [0, 4]");
  for (int a = 0; a < 3; ++a)
    if (!(s->lo[a] < s->hi[a])) return fail(ctx, GVPM_ERR_INVALID, "gvpm_box_scene: empty box");
  return GVPM_OK;
}

int gvpm_generate_rays(gvpm_ctx *ctx, const gvpm_box_scene *scene, const gvpm_pinhole_camera *cam, uint64_t seed,
                       int block, int y0, int y1, float epsilon) {
  if (!ctx || !scene || !cam) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  int rc = check_scene(ctx, scene);
  if (rc) return rc;
  const bool zorder = block < 0;
  const int bs = zorder ? -block : block;
  if (bs < 1 || bs > 32 || (zorder && (bs & (bs - 1)))) return fail(ctx, GVPM_ERR_INVALID, "block must be in [1, 32] (a power of two for Z-order)");
  if (cam->film_w <= 0 || cam->film_h <= 0 || y0 < 0 || y1 > cam->film_h || y0 > y1)
    return fail(ctx, GVPM_ERR_INVALID, "bad film size / row range");
  const size_t n = (size_t)cam->film_w * (size_t)(y1 - y0);
  void *dev = nullptr;
  rc = gvpm_ray_staging(ctx, n, &dev, nullptr);
  if (rc) return rc;
  if (n) {
    RayStaging S = ray_staging_ptrs(dev, n);
    RayGenParams P{};
    P.scene = *scene;
    P.cam = *cam;
    P.seed = seed;
    P.block = bs;
    P.zorder = zorder ? 1 : 0;
    P.y0 = y0;
    P.y1 = y1;
    P.epsilon = epsilon;
    P.o = (float *)S.o; P.d = (float *)S.d; P.mint = (float *)S.mint; P.maxt = (float *)S.maxt;
    P.edge_len = (float *)S.edge_len; P.eye_contrib = (float *)S.eye_contrib; P.xi = (float *)S.xi;
    P.px = (int32_t *)S.px; P.py = (int32_t *)S.py; P.edge_id = (int32_t *)S.edge_id;
    P.off_valid = (uint8_t *)S.off_valid; P.off_o = (float *)S.off_o; P.off_d = (float *)S.off_d;
    P.off_len = (float *)S.off_len; P.off_eye = (float *)S.off_eye; P.off_sensor = (float *)S.off_sensor;
    launch_generate_rays(P, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
  }
  rc = gvpm_commit_rays(ctx);
  if (rc) return rc;
  // The rays of a pinhole are concurrent by construction: what analyse_rays would measure on the device (and read back)
  // follows from the camera.  Every ray is pos + t * dir up to the rounding of the entry point (a few ulp of the
  // coordinates, bounded here by 4e-6 * (1 + |pos|)); projected directions span [-tx, tx] x rows [y0, y1).
  {
    gvpm_ctx::RayFit &F = ctx->pin;
    const float tx = cam->tan_half_fov_x, ty = tx * (float)cam->film_h / (float)cam->film_w;
    for (int k = 0; k < 3; ++k) { F.C[k] = cam->pos[k]; F.m[k] = F.u[k] = F.v[k] = 0.f; }
    F.m[2] = 1.f; F.u[0] = 1.f; F.v[1] = 1.f;
    F.xmin = -tx * 1.0001f; F.xmax = tx * 1.0001f;
    F.ymin = ((float)y0 / (float)cam->film_h - 0.5f) * 2.f * ty;
    F.ymax = ((float)y1 / (float)cam->film_h - 0.5f) * 2.f * ty;
    const float ypad = 1e-4f * ty + 1e-7f;
    F.ymin -= ypad; F.ymax += ypad;
    const float amax = std::max(std::fabs(cam->pos[0]), std::max(std::fabs(cam->pos[1]), std::fabs(cam->pos[2])));
    F.delta = 4e-6f * (1.f + amax);
    F.cosmin = 1.f / std::sqrt(1.f + tx * tx + ty * ty);
    F.count = (float)n;
    F.concurrent = n > 0 && F.cosmin > 0.35f;
    ctx->pin_gen = ctx->rays_gen;
  }
  return GVPM_OK;
}

static int trace_impl(gvpm_ctx *ctx, const gvpm_box_scene *scene, size_t n, uint64_t seed, int max_depth, int rr_depth,
                      int min_depth, uint64_t *n_paths, bool direct) {
  if (!ctx || !scene || n > 0xfffffff0u) return GVPM_ERR_INVALID;
  if (!ctx->have_medium) return fail(ctx, GVPM_ERR_INVALID, "gvpm_set_medium first");
  cudaSetDevice(ctx->device);
  int rc = check_scene(ctx, scene);
  if (rc) return rc;
  void *dev = nullptr;
  PhotonStaging S{};
  if (direct) {
    // no staging: the records are written where the gather reads them
    CK(ctx->aos.reserve(128 * std::max<size_t>(n, 1)));
    ctx->n_photons = (uint32_t)n;
    ctx->photons_loaded = true;
    ctx->photons_direct = true;
    ctx->aos_override = nullptr;
    ctx->region_cap = 0;
    ctx->region_count = nullptr;
    ctx->built = false;
    ++ctx->state_gen;
  } else {
    rc = gvpm_photon_staging(ctx, n, &dev, nullptr);
    if (rc) return rc;
    S = photon_staging_ptrs(dev, n);
  }
  if (n_paths) *n_paths = 0;
  if (n == 0) return GVPM_OK;
  cudaStream_t st = ctx->stream;
  TraceParams P{};
  P.aos = direct ? ctx->aos.as<float4>() : nullptr;
  P.scene = *scene;
  P.sigma_s = ctx->medium.sigma_s[0];
  P.sigma_a = ctx->medium.sigma_a[0];
  P.hg_g = ctx->medium.hg_g;
  P.phase_type = ctx->medium.phase_type;
  P.max_depth = max_depth > 0 ? max_depth : 64;
  P.rr_depth = rr_depth;
  P.min_depth = min_depth;
  P.seed = seed;
  P.n_total = n;
  P.pos = (float *)S.pos; P.flux = (float *)S.flux; P.parent_pos = (float *)S.parent_pos; P.pred_pos = (float *)S.pred_pos;
  P.parent_n = (float *)S.parent_n; P.prefix_flux = (float *)S.prefix_flux; P.parent_albedo = (float *)S.parent_albedo;
  P.parent_pdf = (float *)S.parent_pdf; P.edge_pdf = (float *)S.edge_pdf; P.rr_weight = (float *)S.rr_weight;
  P.parent_type = (uint8_t *)S.parent_type; P.depth = (uint8_t *)S.depth; P.path_id = (uint32_t *)S.path_id;
  // batches of paths until n photons exist; the first one is sized from the photons-per-path ratio seen so far
  unsigned long long filled = 0, pathBase = 0, *host = ctx->pair_count_host;   // host[0] batch total, host[1] last path
  uint32_t pidBase = 0;
  host[1] = 0;
  for (int batch = 0; filled < n; ++batch) {
    if (batch > 64) return fail(ctx, GVPM_ERR_INVALID, "gvpm_trace_photons: the light paths store (almost) no photon in this scene");
    const double ratio = ctx->trace_photons_per_path > 0.0 ? ctx->trace_photons_per_path : 1.0;
    double want = (double)(n - filled) / ratio * 1.03 + 4096.0;
    if (want > 2.0e9) want = 2.0e9;
    const uint32_t np = (uint32_t)want;
    const uint32_t nb = (np + 255) / 256;
    const size_t offCounts = 64, offBlocks = align256(offCounts + np);
    CK(ctx->trace_scratch.reserve(offBlocks + ((size_t)nb + 1) * 8));
    char *sc = ctx->trace_scratch.as<char>();
    unsigned long long *totals = (unsigned long long *)sc;   // [0] batch total, [1] last path
    if (batch == 0) CK(cudaMemsetAsync(sc, 0, 64, st));
    P.path_base = pathBase;
    P.n_paths = np;
    P.slot_base = filled;
    P.path_id_base = pidBase;
    launch_trace_count(P, (uint8_t *)(sc + offCounts), (unsigned long long *)(sc + offBlocks), totals, st);
    launch_trace_emit(P, (const uint8_t *)(sc + offCounts), (const unsigned long long *)(sc + offBlocks), totals + 1, st);
    ctx->launches += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, totals, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const unsigned long long photons = host[0] & ((1ull << 40) - 1ull), paths = host[0] >> 40;
    filled += photons;
    pidBase += (uint32_t)paths;
    pathBase += np;
    ctx->trace_photons_per_path = (double)filled / (double)pathBase;
  }
  if (n_paths) *n_paths = host[1];
  return GVPM_OK;
}

int gvpm_trace_photons(gvpm_ctx *ctx, const gvpm_box_scene *scene, size_t n, uint64_t seed, int max_depth, int rr_depth,
                       int min_depth, uint64_t *n_paths) {
  return trace_impl(ctx, scene, n, seed, max_depth, rr_depth, min_depth, n_paths, false);
}
int gvpm_trace_photons_direct(gvpm_ctx *ctx, const gvpm_box_scene *scene, size_t n, uint64_t seed, int max_depth,
                              int rr_depth, int min_depth, uint64_t *n_paths) {
  return trace_impl(ctx, scene, n, seed, max_depth, rr_depth, min_depth, n_paths, true);
}


// ---- photon dispatch between ranks (dispatch.cu; include/gvpm_b200.h gvpm_dispatch_*) -----------------------------------
// Control block of a rank (its own device memory; the peers write into it, only the owner reads it), in 32-bit words:
enum {
  DC_PUSHED = 0,                           // [2][MAX_PEERS] generation of sender s's last dispatch into my inbox b
  DC_COUNT = 2 * GVPM_MAX_PEERS,           // [2][MAX_PEERS] records sender s wrote into its region of my inbox b
  DC_FREED = 4 * GVPM_MAX_PEERS,           // [2][MAX_PEERS] generation up to which RECEIVER d has released its inbox b
  DC_COLLECTED = 6 * GVPM_MAX_PEERS,       // [2][MAX_PEERS] generation of rank s's last result copy into MY image buffer b
  DC_TIMEOUT = 8 * GVPM_MAX_PEERS,         // set by k_flag_wait when a peer never signalled
  DC_OVERFLOW = 8 * GVPM_MAX_PEERS + 1,
  DC_OCC = 8 * GVPM_MAX_PEERS + 16,        // my ray-occupancy mask (frustum_occ_bytes), copied by the peers at connect
};
struct DispatchBlob {   // what gvpm_dispatch_export writes (GVPM_DISPATCH_BLOB_BYTES)
  long long pid;
  int device, n_peers;
  unsigned long long region_cap;
  void *raw_inbox[2], *raw_ctrl;           // valid inside the exporting process
  cudaIpcMemHandle_t inbox[2], ctrl;
  gvpm_ctx::RayFit fit;
};
static_assert(sizeof(DispatchBlob) <= GVPM_DISPATCH_BLOB_BYTES, "blob size");

int gvpm_dispatch_export(gvpm_ctx *ctx, int n_peers, size_t region_cap, void *blob) {
  if (!ctx || !blob || n_peers < 1 || n_peers > GVPM_MAX_PEERS || region_cap == 0 ||
      (unsigned long long)n_peers * region_cap > 0xfffffff0ull)
    return GVPM_ERR_INVALID;
  if (!ctx->rays_loaded) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_export needs this rank's rays first (gvpm_upload_rays)");
  if (ctx->disp.n_peers) return fail(ctx, GVPM_ERR_INVALID, "dispatch buffers already exported");
  cudaSetDevice(ctx->device);
  int rc = analyse_rays(ctx);
  if (rc) return rc;
  if (!ctx->pin.concurrent)
    return fail(ctx, GVPM_ERR_UNSUPPORTED, "photon dispatch needs concurrent rays (a pinhole's primary rays); use the all-gather exchange");
  gvpm_ctx::Dispatch &D = ctx->disp;
  const size_t ctrl_bytes = (size_t)DC_OCC * 4 + frustum_occ_bytes();
  for (int b = 0; b < 2; ++b) {
    CK(D.inbox[b].reserve((size_t)n_peers * region_cap * 128));
    D.inbox[b].pinned = true;
    CK(D.grids_dev[b].reserve(sizeof(FrustumGrid) * GVPM_MAX_PEERS));
    if (!D.grids_host[b]) CK(cudaHostAlloc((void **)&D.grids_host[b], sizeof(FrustumGrid) * GVPM_MAX_PEERS, cudaHostAllocDefault));
  }
  CK(D.ctrl.reserve(ctrl_bytes));
  D.ctrl.pinned = true;
  CK(D.occ_all.reserve(frustum_occ_bytes() * (size_t)n_peers));
  if (!D.ev_src) CK(cudaEventCreateWithFlags(&D.ev_src, cudaEventDisableTiming));
  CK(cudaMemsetAsync(D.ctrl.p, 0, ctrl_bytes, ctx->stream));
  // the occupancy mask depends on the rays alone (projection basis and bounds of the fit), not on the radius
  const FrustumGrid G = make_frustum_grid(ctx->pin, 1.f, false);
  launch_frustum_mark(ctx->rays.as<float4>(), ctx->n_rays, G, D.ctrl.as<uint32_t>() + DC_OCC, ctx->sm_count, ctx->stream);
  ctx->launches += 1;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  DispatchBlob B{};
  B.pid = (long long)getpid();
  B.device = ctx->device;
  B.n_peers = n_peers;
  B.region_cap = region_cap;
  for (int b = 0; b < 2; ++b) {
    B.raw_inbox[b] = D.inbox[b].p;
    CK(cudaIpcGetMemHandle(&B.inbox[b], D.inbox[b].p));
  }
  B.raw_ctrl = D.ctrl.p;
  CK(cudaIpcGetMemHandle(&B.ctrl, D.ctrl.p));
  B.fit = ctx->pin;
  memset(blob, 0, GVPM_DISPATCH_BLOB_BYTES);
  memcpy(blob, &B, sizeof(B));
  D.n_peers = n_peers;
  D.region_cap = (uint32_t)region_cap;
  return GVPM_OK;
}

int gvpm_dispatch_connect(gvpm_ctx *ctx, const void *blobs, int n_peers, int self_index) {
  if (!ctx || !blobs || self_index < 0 || self_index >= n_peers) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (D.n_peers != n_peers) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_export first, with the same number of peers");
  if (D.connected) return fail(ctx, GVPM_ERR_INVALID, "dispatch peers already connected");
  cudaSetDevice(ctx->device);
  const long long me = (long long)getpid();
  for (int p = 0; p < n_peers; ++p) {
    DispatchBlob B;
    memcpy(&B, (const char *)blobs + (size_t)p * GVPM_DISPATCH_BLOB_BYTES, sizeof(B));
    if (B.n_peers != n_peers || B.region_cap != D.region_cap)
      return fail(ctx, GVPM_ERR_INVALID, "dispatch blobs disagree on the number of peers / the region size");
    D.fit[p] = B.fit;
    if (B.pid == me) {   // same process (several contexts, tests): the pointers are valid as they are
      if (B.device != ctx->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, ctx->device, B.device));
        if (!can) return fail(ctx, GVPM_ERR_UNSUPPORTED, "no peer access between two devices of this process");
        cudaError_t e = cudaDeviceEnablePeerAccess(B.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
      }
      for (int b = 0; b < 2; ++b) D.peer_inbox[b][p] = (char *)B.raw_inbox[b];
      D.peer_ctrl[p] = (uint32_t *)B.raw_ctrl;
    } else {
      for (int b = 0; b < 2; ++b) CK(cudaIpcOpenMemHandle((void **)&D.peer_inbox[b][p], B.inbox[b], cudaIpcMemLazyEnablePeerAccess));
      CK(cudaIpcOpenMemHandle((void **)&D.peer_ctrl[p], B.ctrl, cudaIpcMemLazyEnablePeerAccess));
      D.peer_mapped[p] = true;
    }
    // the receiver's occupancy mask, read once: the classification then touches local memory only
    CK(cudaMemcpyAsync(D.occ_all.as<char>() + frustum_occ_bytes() * (size_t)p, D.peer_ctrl[p] + DC_OCC, frustum_occ_bytes(),
                       cudaMemcpyDeviceToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  D.self = self_index;
  D.connected = true;
  // Do all ranks project on the same plane (same axis and basis - gvpm_set_view_direction - and, up to rounding, the same
  // centre)?  Then a photon is classified ONCE: its footprint box is looked up in an owner map - kOccRes^2 cells over the
  // union of the ranks' ray bounds, one bit per rank whose own occupancy mask has a ray under the cell (dilated by a cell)
  // - instead of one key evaluation per receiver.
  D.shared_frame = n_peers <= 8;
  float cmag = 0.f;
  for (int k = 0; k < 3; ++k) cmag = std::max(cmag, std::fabs(D.fit[0].C[k]));
  for (int p = 1; p < n_peers && D.shared_frame; ++p)
    for (int k = 0; k < 3; ++k)
      D.shared_frame = D.shared_frame && D.fit[p].m[k] == D.fit[0].m[k] && D.fit[p].u[k] == D.fit[0].u[k] &&
                       D.fit[p].v[k] == D.fit[0].v[k] && std::fabs(D.fit[p].C[k] - D.fit[0].C[k]) <= 1e-5f * (1.f + cmag);
  if (D.shared_frame) {
    const int R = (int)std::lround(std::sqrt((double)frustum_occ_bytes() * 8.0));   // kOccRes
    D.ux0 = D.uy0 = 3.4e38f; D.ux1 = D.uy1 = -3.4e38f;
    for (int p = 0; p < n_peers; ++p) {
      D.ux0 = std::min(D.ux0, D.fit[p].xmin); D.ux1 = std::max(D.ux1, D.fit[p].xmax);
      D.uy0 = std::min(D.uy0, D.fit[p].ymin); D.uy1 = std::max(D.uy1, D.fit[p].ymax);
    }
    const double uw = std::max((double)D.ux1 - D.ux0, 1e-20), uh = std::max((double)D.uy1 - D.uy0, 1e-20);
    std::vector<uint32_t> occ(frustum_occ_bytes() / 4 * (size_t)n_peers);
    CK(cudaMemcpy(occ.data(), D.occ_all.p, occ.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> map((size_t)R * R, 0);
    for (int p = 0; p < n_peers; ++p) {
      const gvpm_ctx::RayFit &F = D.fit[p];
      const double wx = std::max((double)F.xmax - F.xmin, 1e-20) / R, wy = std::max((double)F.ymax - F.ymin, 1e-20) / R;
      const uint32_t *o = occ.data() + (frustum_occ_bytes() / 4) * (size_t)p;
      for (int j = 0; j < R; ++j)
        for (int i = 0; i < R; ++i) {
          if (!(o[(j * R + i) >> 5] >> ((j * R + i) & 31) & 1u)) continue;
          const int i0 = (int)std::floor((F.xmin + i * wx - D.ux0) / uw * R) - 1, i1 = (int)std::floor((F.xmin + (i + 1) * wx - D.ux0) / uw * R) + 1;
          const int j0 = (int)std::floor((F.ymin + j * wy - D.uy0) / uh * R) - 1, j1 = (int)std::floor((F.ymin + (j + 1) * wy - D.uy0) / uh * R) + 1;
          for (int jj = std::max(j0, 0); jj <= std::min(j1, R - 1); ++jj)
            for (int ii = std::max(i0, 0); ii <= std::min(i1, R - 1); ++ii) map[(size_t)jj * R + ii] |= (uint8_t)(1u << p);
        }
    }
    CK(D.owner_map.reserve(map.size()));
    CK(cudaMemcpy(D.owner_map.p, map.data(), map.size(), cudaMemcpyHostToDevice));
  }
  return GVPM_OK;
}

int gvpm_dispatch_photons(gvpm_ctx *ctx, int which, size_t n_total, size_t begin, size_t count, float radius, void *after_stream) {
  if (!ctx || (which != 0 && which != 1) || begin + count > n_total || !(radius > 0.f)) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (!D.connected) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_connect has not been called");
  if (count > D.region_cap) return fail(ctx, GVPM_ERR_INVALID, "slice larger than the exported region size");
  if (!ctx->ph_staging.p || ctx->ph_staging.cap < PhotonLayout(n_total).bytes)
    return fail(ctx, GVPM_ERR_INVALID, "size the selected photon staging buffer for n_total first (gvpm_photon_staging)");
  cudaSetDevice(ctx->device);
  // after_stream == the context's own stream: the dispatch runs IN that stream (between a build and its gather, say: a
  // side stream only gets CTA slots as the persistent gather kernels retire theirs, and every rank's next build waits for
  // the slowest rank's dispatch); any other stream / NULL: on the internal highest-priority stream, after that work
  cudaStream_t ks = ctx->push_kernel_stream;
  if (after_stream == (void *)ctx->stream) {
    ks = ctx->stream;
  } else {
    CK(cudaEventRecord(D.ev_src, after_stream ? (cudaStream_t)after_stream : ctx->stream));
    CK(cudaStreamWaitEvent(ks, D.ev_src, 0));
  }
  const uint32_t gen = ++D.gen_push[which];
  uint32_t *ctrl = D.ctrl.as<uint32_t>();
  // every receiver must have released its inbox `which` gen - 1 times (flags the receivers write into MY control block)
  if (gen > 1) {
    launch_flag_wait(ctrl + DC_FREED + which * GVPM_MAX_PEERS, D.n_peers, gen - 1, ctrl + DC_TIMEOUT, ks);
    ctx->launches += 1;
  }
  const bool parity = ctx->have_cfg && ctx->cfg.path_set && !ctx->cfg.sppm_primal;
  const uint32_t nb = (uint32_t)((count + 255) / 256);
  CK(D.keepbits.reserve(std::max<size_t>(count, 256)));
  CK(D.block_cnt.reserve((size_t)GVPM_MAX_PEERS * (nb + 2) * 4));
  static_assert(sizeof(DispatchParams) <= 4000, "kernel parameter space");
  DispatchParams P{};
  P.S = photon_staging_ptrs(ctx->ph_staging.p, n_total);
  P.begin = (uint32_t)begin;
  P.count = (uint32_t)count;
  P.n_dst = D.n_peers;
  for (int d = 0; d < D.n_peers; ++d) P.grids[d] = make_frustum_grid(D.fit[d], radius, parity);
  P.region_cap = D.region_cap;
  P.keepbits = D.keepbits.as<uint8_t>();
  P.block_cnt = D.block_cnt.as<uint32_t>();
  P.nb = nb;
  P.overflow = ctrl + DC_OVERFLOW;
  SignalParams Sg{};
  Sg.n_dst = D.n_peers;
  Sg.self = D.self;
  Sg.nb = nb;
  Sg.block_cnt = P.block_cnt;
  Sg.gen = gen;
  for (int d = 0; d < D.n_peers; ++d) {
    P.occ[d] = D.occ_all.as<uint32_t>() + (frustum_occ_bytes() / 4) * (size_t)d;
    P.inbox[d] = (float4 *)(D.peer_inbox[which][d] + (size_t)D.self * D.region_cap * 128);
    Sg.count_dst[d] = D.peer_ctrl[d] + DC_COUNT + which * GVPM_MAX_PEERS + D.self;
    Sg.flag_dst[d] = D.peer_ctrl[d] + DC_PUSHED + which * GVPM_MAX_PEERS + D.self;
  }
  P.owner_map = D.shared_frame ? D.owner_map.as<uint8_t>() : nullptr;
  if (D.shared_frame) {
    P.ux0 = D.ux0; P.uy0 = D.uy0;
    P.uix = (float)(std::sqrt((double)frustum_occ_bytes() * 8.0) / std::max((double)D.ux1 - D.ux0, 1e-20));
    P.uiy = (float)(std::sqrt((double)frustum_occ_bytes() * 8.0) / std::max((double)D.uy1 - D.uy0, 1e-20));
    P.ux1 = D.ux1; P.uy1 = D.uy1;
    P.pad_r_max = 0.f;
    for (int d = 0; d < D.n_peers; ++d) P.pad_r_max = std::max(P.pad_r_max, P.grids[d].pad_r);
    P.pad_r_max += 2e-5f * (1.f + std::fabs(P.grids[0].C[0]) + std::fabs(P.grids[0].C[1]) + std::fabs(P.grids[0].C[2]));   // the centres agree to 1e-5
  }
  // on the side stream: a small persistent grid (GVPM_DISPATCH_CTAS, default one CTA per SM); in-stream: one CTA per chunk
  launch_dispatch(P, ks, ks == ctx->stream ? 0 : (ctx->dispatch_ctas < 0 ? ctx->sm_count : ctx->dispatch_ctas));
  launch_dispatch_signal(Sg, ks);
  ctx->launches += 4;
  CK(cudaGetLastError());
  return GVPM_OK;
}

int gvpm_build_dispatched(gvpm_ctx *ctx, int which, float radius, uint32_t *n_kept) {
  if (!ctx || (which != 0 && which != 1) || !(radius > 0.f)) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (!D.connected) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_connect has not been called");
  if (!ctx->rays_loaded) return fail(ctx, GVPM_ERR_INVALID, "no rays uploaded");
  cudaSetDevice(ctx->device);
  int rc = analyse_rays(ctx);
  if (rc) return rc;
  const gvpm_ctx::RayFit &Fa = ctx->pin, &Fb = D.fit[D.self];
  bool same = Fa.concurrent && Fb.concurrent && Fa.delta == Fb.delta && Fa.xmin == Fb.xmin && Fa.xmax == Fb.xmax &&
              Fa.ymin == Fb.ymin && Fa.ymax == Fb.ymax && Fa.count == Fb.count;
  for (int k = 0; k < 3; ++k) same = same && Fa.C[k] == Fb.C[k] && Fa.m[k] == Fb.m[k] && Fa.u[k] == Fb.u[k] && Fa.v[k] == Fb.v[k];
  if (!same)
    return fail(ctx, GVPM_ERR_INVALID, "the rays changed since gvpm_dispatch_export: the peers dispatch for the exported ray set");
  const uint32_t gen = ++D.gen_build[which];
  uint32_t *ctrl = D.ctrl.as<uint32_t>();
  launch_flag_wait(ctrl + DC_PUSHED + which * GVPM_MAX_PEERS, D.n_peers, gen, ctrl + DC_TIMEOUT, ctx->stream);
  ctx->launches += 1;
  ctx->n_photons = (uint32_t)((size_t)D.n_peers * D.region_cap);
  ctx->photons_loaded = true;
  ctx->photons_direct = true;
  ctx->aos_override = D.inbox[which].as<float4>();
  ctx->region_cap = D.region_cap;
  ctx->region_count = ctrl + DC_COUNT + which * GVPM_MAX_PEERS;
  ctx->built = false;
  ++ctx->state_gen;
  if (gen == 1 && which == 0) {   // first build over an inbox: mostly empty regions; its exact count sizes the following ones
    ctx->kept_fraction_hint = 0.25;
    ctx->kept_hint_valid = false;
  }
  rc = build_frustum(ctx, radius, n_kept);
  if (rc) return rc;
  // a peer that never signalled (k_flag_wait timed out) or a region overflow: the grid is incomplete, and the gather says
  // so (the same device flag a bounded build sets when it overflows); gvpm_dispatch_status names the cause
  if (ctx->build_ovf) {
    launch_or_flag(const_cast<uint32_t *>(ctx->build_ovf), ctrl + DC_TIMEOUT, ctx->stream);
    launch_or_flag(const_cast<uint32_t *>(ctx->build_ovf), ctrl + DC_OVERFLOW, ctx->stream);
    ctx->launches += 2;
    CK(cudaGetLastError());
  }
  return GVPM_OK;
}

int gvpm_dispatch_release(gvpm_ctx *ctx, int which) {
  if (!ctx || (which != 0 && which != 1)) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (!D.connected) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_connect has not been called");
  cudaSetDevice(ctx->device);
  FlagSetParams F{};
  F.n = D.n_peers;
  F.value = D.gen_build[which];
  for (int s = 0; s < D.n_peers; ++s) F.dst[s] = D.peer_ctrl[s] + DC_FREED + which * GVPM_MAX_PEERS + D.self;
  launch_flag_set(F, ctx->stream);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return GVPM_OK;
}


// ---- result collection over the copy engines (rank 0's image buffer, written by every rank) -------------------------------
int gvpm_shared_buffer_create(gvpm_ctx *ctx, size_t bytes, void **dev, void *handle) {
  if (!ctx || !dev || !handle || bytes == 0) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  void *p = nullptr;
  CK(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); CK(e); }
  ctx->disp.shared_owned.push_back(p);
  static_assert(sizeof(cudaIpcMemHandle_t) + 16 <= GVPM_SHARED_HANDLE_BYTES, "handle size");
  memset(handle, 0, GVPM_SHARED_HANDLE_BYTES);
  memcpy(handle, &h, sizeof(h));
  const long long pid = (long long)getpid();
  memcpy((char *)handle + sizeof(h), &pid, 8);
  memcpy((char *)handle + sizeof(h) + 8, &p, 8);
  *dev = p;
  return GVPM_OK;
}

int gvpm_shared_buffer_open(gvpm_ctx *ctx, const void *handle, void **dev) {
  if (!ctx || !dev || !handle) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaIpcMemHandle_t h;
  long long pid = 0;
  void *raw = nullptr;
  memcpy(&h, handle, sizeof(h));
  memcpy(&pid, (const char *)handle + sizeof(h), 8);
  memcpy(&raw, (const char *)handle + sizeof(h) + 8, 8);
  if (pid == (long long)getpid()) { *dev = raw; return GVPM_OK; }   // same process: the pointer is valid as it is
  void *p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->disp.shared_opened.push_back(p);
  *dev = p;
  return GVPM_OK;
}

int gvpm_collect_signal(gvpm_ctx *ctx, int which, int root, void *stream) {
  if (!ctx || (which != 0 && which != 1)) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (!D.connected || root < 0 || root >= D.n_peers) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_connect has not been called");
  cudaSetDevice(ctx->device);
  FlagSetParams F{};
  F.n = 1;
  F.value = ++D.gen_collect[which];
  F.dst[0] = D.peer_ctrl[root] + DC_COLLECTED + which * GVPM_MAX_PEERS + D.self;
  launch_flag_set(F, stream ? (cudaStream_t)stream : ctx->stream);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return GVPM_OK;
}

int gvpm_collect_wait(gvpm_ctx *ctx, int which) {
  if (!ctx || (which != 0 && which != 1)) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (!D.connected) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_connect has not been called");
  cudaSetDevice(ctx->device);
  uint32_t *ctrl = D.ctrl.as<uint32_t>();
  // every rank signals once per iteration and buffer: wait until all have caught up with this rank's own count
  launch_flag_wait(ctrl + DC_COLLECTED + which * GVPM_MAX_PEERS, D.n_peers, D.gen_collect[which], ctrl + DC_TIMEOUT, ctx->stream);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return GVPM_OK;
}

int gvpm_dispatch_join(gvpm_ctx *ctx) {
  if (!ctx) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (!D.connected) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_connect has not been called");
  cudaSetDevice(ctx->device);
  CK(cudaEventRecord(D.ev_src, ctx->push_kernel_stream));
  CK(cudaStreamWaitEvent(ctx->stream, D.ev_src, 0));
  return GVPM_OK;
}

int gvpm_dispatch_status(gvpm_ctx *ctx, uint32_t counts[GVPM_MAX_PEERS], int which) {
  if (!ctx || (which != 0 && which != 1)) return GVPM_ERR_INVALID;
  gvpm_ctx::Dispatch &D = ctx->disp;
  if (!D.connected) return fail(ctx, GVPM_ERR_INVALID, "gvpm_dispatch_connect has not been called");
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(ctx->push_kernel_stream));
  uint32_t h[DC_OCC];
  CK(cudaMemcpy(h, D.ctrl.p, sizeof(h), cudaMemcpyDeviceToHost));
  if (counts) for (int s = 0; s < GVPM_MAX_PEERS; ++s) counts[s] = s < D.n_peers ? h[DC_COUNT + which * GVPM_MAX_PEERS + s] : 0u;
  if (h[DC_TIMEOUT]) return fail(ctx, GVPM_ERR_CUDA, "photon dispatch: a peer never signalled (flag wait timed out)");
  if (h[DC_OVERFLOW]) return fail(ctx, GVPM_ERR_INVALID, "photon dispatch: inbox region overflow");
  return GVPM_OK;
}

int gvpm_measure_read_bandwidth(gvpm_ctx *ctx, size_t bytes, int reps, double *gb_per_s) {
  if (!ctx || !gb_per_s || bytes < 4096 || bytes > (1ull << 32) || reps < 1) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CK(ctx->poisson_io.reserve(bytes + 256));
  CK(cudaMemsetAsync(ctx->poisson_io.p, 1, bytes + 256, ctx->stream));
  unsigned *sink = (unsigned *)(ctx->poisson_io.as<char>() + bytes);
  launch_l2_read(ctx->poisson_io.p, bytes, 2, sink, ctx->sm_count, ctx->stream);   // warm: the buffer is in L2 (if it fits)
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  launch_l2_read(ctx->poisson_io.p, bytes, reps, sink, ctx->sm_count, ctx->stream);
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
  ctx->launches += 2;
  *gb_per_s = (double)(bytes / 16 * 16) * reps / ((double)ms * 1e-3) / 1e9;
  return GVPM_OK;
}

int gvpm_read_device(gvpm_ctx *ctx, const void *dev, void *host, size_t bytes) {
  if (!ctx || (bytes && (!dev || !host))) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (bytes) CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

int gvpm_staging_peek(gvpm_ctx *ctx, int which, void **dev, size_t *count) {
  if (!ctx || !dev || (which != 0 && which != 1)) return GVPM_ERR_INVALID;
  *dev = which == 0 ? ctx->ph_staging.p : ctx->ray_staging.p;
  if (count) *count = which == 0 ? ctx->n_photons : ctx->n_rays;
  return GVPM_OK;
}

int gvpm_accel_kind(const gvpm_ctx *ctx) { return ctx ? ctx->accel : 0; }

int gvpm_set_view_direction(gvpm_ctx *ctx, const float dir[3]) {
  if (!ctx) return GVPM_ERR_INVALID;
  if (dir) {
    const double l = std::sqrt((double)dir[0] * dir[0] + (double)dir[1] * dir[1] + (double)dir[2] * dir[2]);
    if (!(l > 0.0) || !std::isfinite(l)) return fail(ctx, GVPM_ERR_INVALID, "view direction must be a non-zero vector");
    for (int k = 0; k < 3; ++k) ctx->view_dir[k] = dir[k];
  }
  ctx->have_view_dir = dir != nullptr;
  ctx->pin_gen = ~0ull;   // the rays are analysed again with the new axis
  return GVPM_OK;
}

int gvpm_ray_staging(gvpm_ctx *ctx, size_t n, void **dev, size_t *bytes) {
  if (!ctx || n > 0xfffffff0u) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  RayLayout L(n);
  CK(ctx->ray_staging.reserve(L.bytes ? L.bytes : 256));
  ctx->n_rays = (uint32_t)n;
  ctx->rays_loaded = false;
  ++ctx->rays_gen;
  ++ctx->state_gen;
  if (dev) *dev = ctx->ray_staging.p;
  if (bytes) *bytes = L.bytes;
  return GVPM_OK;
}

int gvpm_commit_rays(gvpm_ctx *ctx) {
  if (!ctx) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const uint32_t n = ctx->n_rays;
  CK(ctx->rays.reserve((size_t)n * GVPM_RAY_FLOAT4 * 16 + 256));
  CK(ctx->out.reserve((size_t)n * GVPM_OUT_FLOATS * 4 + 256));
  CK(ctx->counts.reserve((size_t)n * 8 + 256));
  if (n > 0) {
    launch_pack_rays(ray_staging_ptrs(ctx->ray_staging.p, n), n, ctx->rays.as<float4>(), ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
  }
  ctx->rays_loaded = true;
  if (n != ctx->hint_rays) {   // another ray set (not a re-jittered one): what the last build kept says nothing about the next
    ctx->hint_rays = n;
    ctx->kept_hint_valid = false;
    ctx->hint_pending = false;
    ctx->kept_fraction_hint = 1.0;
  }
  return GVPM_OK;
}

int gvpm_upload_rays(gvpm_ctx *ctx, const gvpm_ray_soa *r, size_t n) {
  if (!ctx || (n && !r)) return GVPM_ERR_INVALID;
  int rc = gvpm_ray_staging(ctx, n, nullptr, nullptr);
  if (rc) return rc;
  if (n > 0) {
    RayLayout L(n);
    char *b = (char *)ctx->ray_staging.p;
    const void *src[16] = {r->o, r->d, r->mint, r->maxt, r->edge_len, r->eye_contrib, r->xi, r->px, r->py,
                           r->edge_id, r->off_valid, r->off_o, r->off_d, r->off_len, r->off_eye, r->off_sensor};
    const size_t sz[16] = {12 * n, 12 * n, 4 * n, 4 * n, 4 * n, 12 * n, 4 * n, 4 * n, 4 * n, 4 * n,
                           4 * n,  48 * n, 48 * n, 16 * n, 48 * n, 16 * n};
    for (int i = 0; i < 16; ++i) {
      if (!src[i]) return fail(ctx, GVPM_ERR_INVALID, "null array in gvpm_ray_soa");
      CK(cudaMemcpyAsync(b + L.off[i], src[i], sz[i], cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  return gvpm_commit_rays(ctx);
}

int gvpm_gather_bre_into(gvpm_ctx *ctx, float *out_dev, uint32_t *counts_dev) {
  if (!ctx || !out_dev) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return gather_async(ctx, out_dev, counts_dev);
}

int gvpm_gather_bre_device(gvpm_ctx *ctx, const float **out_dev, const uint32_t **counts_dev) {
  if (!ctx) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  int rc = gather_async(ctx, ctx->out.as<float>(), ctx->counts.as<uint32_t>());
  if (rc) return rc;
  if (out_dev) *out_dev = ctx->out.as<float>();
  if (counts_dev) *counts_dev = ctx->counts.as<uint32_t>();
  return GVPM_OK;
}

int gvpm_gather_bre(gvpm_ctx *ctx, float *out, uint32_t *counts) {
  if (!ctx || !out) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  int rc = pending_check(ctx, true);
  if (rc) return rc;
  uint32_t *cdev = counts ? ctx->counts.as<uint32_t>() : nullptr;   // no counts wanted: the traversal filters before queueing
  rc = gather_common(ctx, ctx->out.as<float>(), cdev, ctx->pair_count_host);
  if (rc) return rc;
  const size_t n = ctx->n_rays;
  for (int pass = 0; pass < 2; ++pass) {   // second pass only if the pair list overflowed and the gather was re-run
    if (n) {
      CK(cudaMemcpyAsync(out, ctx->out.p, n * GVPM_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      if (counts)
        CK(cudaMemcpyAsync(counts, ctx->counts.p, n * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    bool redone = false;
    if (pass == 0) rc = gather_finish(ctx, ctx->out.as<float>(), cdev, ctx->pair_count_host, &redone);
    else CK(cudaStreamSynchronize(ctx->stream));
    if (rc) return rc;
    if (!redone) break;
  }
  return GVPM_OK;
}

// sppm's primal BRE (bre.cpp:167-259 driven as sppm.cpp:926-981): same traverse + shade kernels with
// gvpm_config.sppm_primal set; only the primal (first) spectrum of every ray is produced.
int gvpm_gather_sppm_bre(gvpm_ctx *ctx, float *out, uint32_t *counts) {
  if (!ctx || !out) return GVPM_ERR_INVALID;
  if (!ctx->have_cfg || !ctx->cfg.sppm_primal)
    return fail(ctx, GVPM_ERR_INVALID, "gvpm_gather_sppm_bre needs gvpm_config.sppm_primal = 1");
  cudaSetDevice(ctx->device);
  int rc = pending_check(ctx, true);
  if (rc) return rc;
  rc = gather_common(ctx, ctx->out.as<float>(), ctx->counts.as<uint32_t>(), ctx->pair_count_host);
  if (rc) return rc;
  rc = gather_finish(ctx, ctx->out.as<float>(), ctx->counts.as<uint32_t>(), ctx->pair_count_host, nullptr);
  if (rc) return rc;
  const size_t n = ctx->n_rays;
  if (n) {
    CK(cudaMemcpy2DAsync(out, 3 * sizeof(float), ctx->out.p, GVPM_OUT_FLOATS * sizeof(float), 3 * sizeof(float), n,
                         cudaMemcpyDeviceToHost, ctx->stream));
    if (counts)
      CK(cudaMemcpyAsync(counts, ctx->counts.p, n * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

// Pipelined iteration tail: rays are uploaded in chunks on a copy stream while earlier chunks are packed, traversed
// and shaded on the compute stream and their results stream back on a third one, so PCIe (both directions) and the
// SMs work at the same time.  Equivalent to gvpm_upload_rays + gvpm_gather_bre.
int gvpm_gather_bre_host(gvpm_ctx *ctx, const gvpm_ray_soa *r, size_t n, float *out) {
  if (!ctx || (n && (!r || !out))) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  int rc = gvpm_ray_staging(ctx, n, nullptr, nullptr);
  if (rc) return rc;
  CK(ctx->rays.reserve((size_t)n * GVPM_RAY_FLOAT4 * 16 + 256));
  CK(ctx->out.reserve((size_t)n * GVPM_OUT_FLOATS * 4 + 256));
  CK(ctx->counts.reserve((size_t)n * 8 + 256));
  ctx->rays_loaded = true;
  GatherParams P;
  rc = fill_params(ctx, P, ctx->out.as<float>(), nullptr);
  if (rc) return rc;
  rc = pending_check(ctx, true);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  ctx->last_pairs = 0;
  int used = 0;
  size_t per = 0;
  if (n) {
    const void *src[16] = {r->o, r->d, r->mint, r->maxt, r->edge_len, r->eye_contrib, r->xi, r->px, r->py,
                           r->edge_id, r->off_valid, r->off_o, r->off_d, r->off_len, r->off_eye, r->off_sensor};
    const size_t elt[16] = {12, 12, 4, 4, 4, 12, 4, 4, 4, 4, 4, 48, 48, 16, 48, 16};
    for (int i = 0; i < 16; ++i)
      if (!src[i]) return fail(ctx, GVPM_ERR_INVALID, "null array in gvpm_ray_soa");
    if (ctx->pair_cap == 0) {
      rc = reserve_pairs(ctx, std::max<size_t>(1u << 20, 16 * n));
      if (rc) return rc;
    }
    // chunks of whole 32-ray tiles, ~256k rays each
    int nChunks = (int)std::min<size_t>(32, std::max<size_t>(1, n / 262144));
    per = ((n + nChunks - 1) / nChunks + 31) & ~(size_t)31;
    RayLayout L(n);
    char *stg = (char *)ctx->ray_staging.p;
    const RayStaging S = ray_staging_ptrs(ctx->ray_staging.p, n);
    // the staging buffer may still be read by work queued earlier on the compute stream
    CK(cudaEventRecord(ctx->pipe_ev[64], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_in, ctx->pipe_ev[64], 0));
    CK(cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_ev[64], 0));
    for (int c = 0; c < nChunks; ++c) {
      const size_t r0 = (size_t)c * per, r1 = std::min(n, r0 + per);
      if (r0 >= r1) break;
      for (int i = 0; i < 16; ++i)
        CK(cudaMemcpyAsync(stg + L.off[i] + r0 * elt[i], (const char *)src[i] + r0 * elt[i], (r1 - r0) * elt[i],
                           cudaMemcpyHostToDevice, ctx->copy_in));
      CK(cudaEventRecord(ctx->pipe_ev[2 * c], ctx->copy_in));
      used = c + 1;
    }
    for (int c = 0; c < used; ++c) {
      const size_t r0 = (size_t)c * per, r1 = std::min(n, r0 + per);
      CK(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[2 * c], 0));
      launch_pack_rays_range(S, (uint32_t)r0, (uint32_t)r1, ctx->rays.as<float4>(), ctx->stream);
      ctx->launches += 1;
      CK(cudaMemsetAsync(ctx->out.as<float>() + r0 * GVPM_OUT_FLOATS, 0, (r1 - r0) * GVPM_OUT_FLOATS * sizeof(float),
                         ctx->stream));
      rc = enqueue_range(ctx, P, (uint32_t)r0, (uint32_t)r1, ctx->pair_count_host + 16 + c);
      if (rc) return rc;
      CK(cudaEventRecord(ctx->pipe_ev[2 * c + 1], ctx->stream));
      CK(cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_ev[2 * c + 1], 0));
      CK(cudaMemcpyAsync(out + r0 * GVPM_OUT_FLOATS, ctx->out.as<float>() + r0 * GVPM_OUT_FLOATS,
                         (r1 - r0) * GVPM_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
    }
  }
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed_gather = true;
  CK(cudaStreamSynchronize(ctx->copy_out));
  CK(cudaStreamSynchronize(ctx->stream));
  // chunks whose pair list overflowed are run again the slow way (first iteration of a render at most)
  for (int c = 0; c < used; ++c) {
    const unsigned long long total = ctx->pair_count_host[16 + c];
    if (total <= ctx->pair_cap) { ctx->last_pairs += total; continue; }
    const size_t r0 = (size_t)c * per, r1 = std::min(n, r0 + per);
    rc = reserve_pairs(ctx, total + total / 8 + 1024);
    if (rc) return rc;
    rc = gather_range_sync(ctx, P, (uint32_t)r0, (uint32_t)r1, 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out + r0 * GVPM_OUT_FLOATS, ctx->out.as<float>() + r0 * GVPM_OUT_FLOATS,
                       (r1 - r0) * GVPM_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return GVPM_OK;
}

int gvpm_dump_neighbours_bre(gvpm_ctx *ctx, uint64_t *offsets, uint32_t *idx, size_t cap) {
  if (!ctx || !offsets) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t n = ctx->n_rays;
  // pass 1: counts
  {
    int rc = pending_check(ctx, true);
    if (rc) return rc;
    rc = gather_common(ctx, ctx->out.as<float>(), ctx->counts.as<uint32_t>(), ctx->pair_count_host);
    if (rc) return rc;
    rc = gather_finish(ctx, ctx->out.as<float>(), ctx->counts.as<uint32_t>(), ctx->pair_count_host, nullptr);
    if (rc) return rc;
  }
  std::vector<uint32_t> counts(2 * n + 2);
  if (n) CK(cudaMemcpyAsync(counts.data(), ctx->counts.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) { offsets[i] = total; total += counts[2 * i]; }
  offsets[n] = total;
  if (total > cap || (total && !idx)) return fail(ctx, GVPM_ERR_INVALID, "neighbour buffer too small");
  if (total == 0) return GVPM_OK;
  CK(ctx->nbr_offsets.reserve((n + 1) * 8));
  CK(ctx->nbr_idx.reserve(total * 4));
  CK(cudaMemcpyAsync(ctx->nbr_offsets.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  GatherParams P;
  int rc = fill_params(ctx, P, ctx->out.as<float>(), nullptr);
  if (rc) return rc;
  P.nbr_offsets = ctx->nbr_offsets.as<uint64_t>();
  P.nbr_idx = ctx->nbr_idx.as<uint32_t>();
  CK(cudaMemsetAsync(ctx->work_counter.p, 0, 16, ctx->stream));
  CK(launch_bre_traverse(P, true, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  if (ctx->aos_override) {   // dispatched set: report the photons by their index in the whole set
    launch_translate_idx(ctx->nbr_idx.as<uint32_t>(), total, ctx->aos_override, ctx->stream);
    ctx->launches += 1;
  }
  CK(cudaMemcpyAsync(idx, ctx->nbr_idx.p, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

// ---- G-Planes 0D -------------------------------------------------------------------------------
struct PlaneLayout {
  size_t off[7], bytes;
  explicit PlaneLayout(size_t n) {
    const size_t sz[7] = {12 * n, 12 * n, 4 * n, 12 * n, 4 * n, 12 * n, 4 * n};
    size_t o = 0;
    for (int i = 0; i < 7; ++i) { off[i] = o; o += align256(sz[i]); }
    bytes = o;
  }
};
int gvpm_upload_planes(gvpm_ctx *ctx, const gvpm_plane_soa *p, size_t n) {
  if (!ctx || (n && !p) || n > 0x0ffffff0u) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  ctx->n_planes = (uint32_t)n;
  ctx->planes_loaded = true;
  ctx->planes_built = false;
  ++ctx->state_gen;
  if (n == 0) return GVPM_OK;
  if (!p->origin || !p->w0 || !p->length0 || !p->w1 || !p->length1 || !p->flux || !p->edge_id)
    return fail(ctx, GVPM_ERR_INVALID, "null array in gvpm_plane_soa");
  PlaneLayout L(n);
  CK(ctx->plane_staging.reserve(L.bytes));
  CK(ctx->plane_raw.reserve(GVPM_PLANE_PLANES * n * sizeof(float4)));
  CK(ctx->plane_pos.reserve(12 * n));
  const void *src[7] = {p->origin, p->w0, p->length0, p->w1, p->length1, p->flux, p->edge_id};
  const size_t sz[7] = {12 * n, 12 * n, 4 * n, 12 * n, 4 * n, 12 * n, 4 * n};
  char *stg = (char *)ctx->plane_staging.p;
  for (int i = 0; i < 7; ++i) CK(cudaMemcpyAsync(stg + L.off[i], src[i], sz[i], cudaMemcpyHostToDevice, ctx->stream));
  PlaneStaging S;
  S.origin = (const float *)(stg + L.off[0]); S.w0 = (const float *)(stg + L.off[1]); S.length0 = (const float *)(stg + L.off[2]);
  S.w1 = (const float *)(stg + L.off[3]); S.length1 = (const float *)(stg + L.off[4]); S.flux = (const float *)(stg + L.off[5]);
  S.edge_id = (const int32_t *)(stg + L.off[6]);
  uint32_t *flag = ctx->work_counter.as<uint32_t>() + 12;   // word 12 of the counter block
  CK(cudaMemsetAsync(flag, 0, 4, ctx->stream));
  launch_pack_planes(S, (uint32_t)n, ctx->plane_raw.as<float4>(), ctx->plane_pos.as<float>(), flag, ctx->stream);
  ctx->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->sample_stats_host + 2, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // the caller's arrays are only read during the call
  if (ctx->sample_stats_host[2]) {
    // PhotonPlane ctor asserts (plane_struct.h:45-46)
    ctx->planes_loaded = false;
    return fail(ctx, GVPM_ERR_INVALID, "photon plane with zero or non-finite edge length");
  }
  return GVPM_OK;
}

int gvpm_build_planes(gvpm_ctx *ctx) {
  if (!ctx) return GVPM_ERR_INVALID;
  if (!ctx->planes_loaded) return fail(ctx, GVPM_ERR_INVALID, "no planes uploaded");
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const uint32_t n = ctx->n_planes;
  CK(cudaEventRecord(ctx->ev[0], st));
  const uint32_t nLeaves = (n + 31) / 32;
  if (n > 0) {
    CK(ctx->plane_bounds.reserve(256));
    CK(ctx->bounds_partial.reserve((size_t)bounds_blocks(n) * 6 * sizeof(float)));
    CK(ctx->keys_in.reserve(8 * (size_t)n));
    CK(ctx->keys_out.reserve(8 * (size_t)n));
    CK(ctx->vals_in.reserve(4 * (size_t)n));
    CK(ctx->vals_out.reserve(4 * (size_t)n));
    CK(ctx->sort_temp.reserve(sort_temp_bytes(n)));
    CK(ctx->plane_rec.reserve(GVPM_PLANE_PLANES * (size_t)n * sizeof(float4)));
    CK(ctx->plane_orig.reserve(4 * (size_t)n));
    CK(ctx->plane_box_lo.reserve(16 * (size_t)nLeaves));
    CK(ctx->plane_box_hi.reserve(16 * (size_t)nLeaves));
    launch_bounds(ctx->plane_pos.as<float>(), n, ctx->bounds_partial.as<float>(), ctx->plane_bounds.as<float>(), st);
    launch_morton(ctx->plane_pos.as<float>(), n, ctx->plane_bounds.as<float>(), ctx->keys_in.as<uint32_t>(),
                  ctx->vals_in.as<uint32_t>(), st);
    CK(run_sort(ctx->sort_temp.p, ctx->sort_temp.cap, ctx->keys_in.as<uint32_t>(), ctx->keys_out.as<uint32_t>(),
                ctx->vals_in.as<uint32_t>(), ctx->vals_out.as<uint32_t>(), n, st));
    launch_plane_pack_sorted(ctx->plane_raw.as<float4>(), ctx->vals_out.as<uint32_t>(), n, ctx->plane_rec.as<float4>(),
                             ctx->plane_orig.as<uint32_t>(), st);
    launch_plane_leaf_boxes(ctx->plane_rec.as<float4>(), n, nLeaves, ctx->plane_box_lo.as<float4>(),
                            ctx->plane_box_hi.as<float4>(), st);
    ctx->launches += 4 + 4;
    CK(cudaGetLastError());
  }
  ctx->planes_built = true;
  CK(cudaEventRecord(ctx->ev[1], st));
  ctx->timed_build = true;
  return GVPM_OK;
}

static int plane_params(gvpm_ctx *ctx, GatherParams &P) {
  if (!ctx->have_medium || !ctx->have_cfg) return fail(ctx, GVPM_ERR_INVALID, "medium/config not set");
  if (!ctx->planes_built) return fail(ctx, GVPM_ERR_INVALID, "gvpm_build_planes has not been called");
  if (!ctx->rays_loaded) return fail(ctx, GVPM_ERR_INVALID, "no rays uploaded");
  memset(&P, 0, sizeof(P));
  P.tree.lo = ctx->plane_box_lo.as<float4>();
  P.tree.hi = ctx->plane_box_hi.as<float4>();
  P.tree.n = ctx->n_planes;
  P.rays = ctx->rays.as<float4>();
  P.n_rays = ctx->n_rays;
  for (int i = 0; i < 3; ++i) {
    P.sigma_s[i] = ctx->medium.sigma_s[i];
    P.sigma_t[i] = ctx->medium.sigma_s[i] + ctx->medium.sigma_a[i];
  }
  P.phase_type = ctx->medium.phase_type;
  P.hg_g = ctx->medium.hg_g;
  P.sampling_weight = ctx->medium.sampling_weight;
  P.cfg = ctx->cfg;
  P.out = ctx->out.as<float>();
  P.counts = ctx->counts.as<uint32_t>();
  P.work_counter = ctx->work_counter.as<uint32_t>();
  P.plane_rec = ctx->plane_rec.as<float4>();
  P.plane_orig = ctx->plane_orig.as<uint32_t>();
  P.n_planes = ctx->n_planes;
  return GVPM_OK;
}

int gvpm_gather_planes_device(gvpm_ctx *ctx, const float **out_dev, const uint32_t **counts_dev) {
  if (!ctx) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  GatherParams P;
  int rc = plane_params(ctx, P);
  if (rc) return rc;
  if (!counts_dev) P.counts = nullptr;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  CK(cudaMemsetAsync(ctx->work_counter.p, 0, 32, ctx->stream));
  ctx->split_timed = false;
  CK(launch_plane_gather(P, false, ctx->sm_count, ctx->stream));
  ctx->launches += ctx->n_rays ? 1 : 0;
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed_gather = true;
  if (out_dev) *out_dev = ctx->out.as<float>();
  if (counts_dev) *counts_dev = ctx->counts.as<uint32_t>();
  return GVPM_OK;
}

int gvpm_gather_planes(gvpm_ctx *ctx, float *out, uint32_t *counts) {
  if (!ctx || !out) return GVPM_ERR_INVALID;
  const uint32_t *cd = nullptr;
  int rc = gvpm_gather_planes_device(ctx, nullptr, counts ? &cd : nullptr);
  if (rc) return rc;
  const size_t n = ctx->n_rays;
  if (n) {
    CK(cudaMemcpyAsync(out, ctx->out.p, n * GVPM_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) CK(cudaMemcpyAsync(counts, ctx->counts.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

int gvpm_dump_neighbours_planes(gvpm_ctx *ctx, uint64_t *offsets, uint32_t *idx, size_t cap) {
  if (!ctx || !offsets) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t n = ctx->n_rays;
  std::vector<float> tmp(n * GVPM_OUT_FLOATS + 1);
  std::vector<uint32_t> counts(2 * n + 2);
  int rc = gvpm_gather_planes(ctx, tmp.data(), counts.data());
  if (rc) return rc;
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) { offsets[i] = total; total += counts[2 * i]; }
  offsets[n] = total;
  if (total > cap || (total && !idx)) return fail(ctx, GVPM_ERR_INVALID, "neighbour buffer too small");
  if (total == 0) return GVPM_OK;
  CK(ctx->nbr_offsets.reserve((n + 1) * 8));
  CK(ctx->nbr_idx.reserve(total * 4));
  CK(cudaMemcpyAsync(ctx->nbr_offsets.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  GatherParams P;
  rc = plane_params(ctx, P);
  if (rc) return rc;
  P.counts = nullptr;
  P.nbr_offsets = ctx->nbr_offsets.as<uint64_t>();
  P.nbr_idx = ctx->nbr_idx.as<uint32_t>();
  CK(cudaMemsetAsync(ctx->work_counter.p, 0, 32, ctx->stream));
  CK(launch_plane_gather(P, true, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(idx, ctx->nbr_idx.p, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n; ++i) std::sort(idx + offsets[i], idx + offsets[i + 1]);
  return GVPM_OK;
}

// ---- G-Beams 3D --------------------------------------------------------------------------------
// raw gvpm_beam_soa arrays back to back in one device buffer, each 256-byte aligned
struct BeamLayout {
  size_t off[14], bytes;
  explicit BeamLayout(size_t n) {
    const size_t sz[14] = {12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 4 * n, 4 * n, n, n, n, 4 * n};
    size_t o = 0;
    for (int i = 0; i < 14; ++i) { off[i] = o; o += align256(sz[i]); }
    bytes = o;
  }
};
static BeamStaging beam_staging_ptrs(const void *base, size_t n) {
  BeamLayout L(n);
  const char *b = (const char *)base;
  BeamStaging S;
  S.origin = (const float *)(b + L.off[0]); S.end = (const float *)(b + L.off[1]); S.flux = (const float *)(b + L.off[2]);
  S.prefix_flux = (const float *)(b + L.off[3]); S.parent_n = (const float *)(b + L.off[4]);
  S.parent_albedo = (const float *)(b + L.off[5]); S.pred_pos = (const float *)(b + L.off[6]);
  S.end_n = (const float *)(b + L.off[7]); S.parent_pdf = (const float *)(b + L.off[8]);
  S.rr_weight = (const float *)(b + L.off[9]); S.parent_type = (const uint8_t *)(b + L.off[10]);
  S.end_on_surface = (const uint8_t *)(b + L.off[11]); S.depth = (const uint8_t *)(b + L.off[12]);
  S.path_id = (const uint32_t *)(b + L.off[13]);
  return S;
}

// The 14 arrays go up as they are and a kernel assembles the 128-byte beam records (pack_prims.cu).  While the DMA
// runs the host does the one thing that has to be sequential: the fp32 running sum of the beam lengths that fixes the
// sub-beam size (beams_accel.h:98-104), and with it the number of sub-beams (sizes the sort without a read-back).
int gvpm_upload_beams(gvpm_ctx *ctx, const gvpm_beam_soa *b, size_t n) {
  if (!ctx || (n && !b) || n > 0x0ffffff0u) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  ctx->n_beams = (uint32_t)n;
  ctx->beams_loaded = true;
  ctx->beams_built = false;
  ctx->n_subs = 0;
  ctx->beam_subsize = 0.f;
  ++ctx->state_gen;
  if (n == 0) return GVPM_OK;
  if (!b->origin || !b->end || !b->flux || !b->prefix_flux || !b->parent_n || !b->parent_albedo || !b->pred_pos ||
      !b->end_n || !b->parent_pdf || !b->rr_weight || !b->parent_type || !b->end_on_surface || !b->depth || !b->path_id)
    return fail(ctx, GVPM_ERR_INVALID, "null array in gvpm_beam_soa");
  BeamLayout L(n);
  CK(ctx->beam_staging.reserve(L.bytes));
  CK(ctx->beams.reserve(8 * n * sizeof(float4)));
  CK(ctx->beam_len.reserve(4 * n));
  CK(ctx->beam_aux.reserve(64 + 4 * ((n + 255) / 256 + 1)));
  const void *src[14] = {b->origin, b->end, b->flux, b->prefix_flux, b->parent_n, b->parent_albedo, b->pred_pos, b->end_n,
                         b->parent_pdf, b->rr_weight, b->parent_type, b->end_on_surface, b->depth, b->path_id};
  const size_t sz[14] = {12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 12 * n, 4 * n, 4 * n, n, n, n, 4 * n};
  char *stg = (char *)ctx->beam_staging.p;
  for (int i = 0; i < 14; ++i) CK(cudaMemcpyAsync(stg + L.off[i], src[i], sz[i], cudaMemcpyHostToDevice, ctx->stream));
  launch_pack_beams(beam_staging_ptrs(stg, n), (uint32_t)n, ctx->beams.as<float4>(), ctx->beam_len.as<float>(), ctx->stream);
  ctx->launches += 1;
  // host, under the DMA: PhotonBeam::setEndPoint lengths, sequential sum, sub-beam count (SubBeamBVH ctor)
  std::vector<float> len(n);
  float avgSize = 0.f;
  for (size_t i = 0; i < n; ++i) {
    const float *o = b->origin + 3 * i, *e = b->end + 3 * i;
    const float d[3] = {e[0] - o[0], e[1] - o[1], e[2] - o[2]};
    len[i] = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    avgSize += len[i];
  }
  avgSize /= (float)n;
  const float subbeamSize = avgSize / 10;
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    int nSub = subbeamSize > 0.f ? (int)std::ceil(len[i] / subbeamSize) : 1;
    total += (uint64_t)(nSub < 1 ? 1 : nSub);
  }
  if (total > 0xfffffff0ull) return fail(ctx, GVPM_ERR_INVALID, "too many sub-beams");
  ctx->beam_subsize = subbeamSize;
  ctx->n_subs = (uint32_t)total;
  CK(cudaMemcpyAsync(ctx->beam_aux.p, &ctx->beam_subsize, 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));   // the caller's arrays are only read during the call
  return GVPM_OK;
}

int gvpm_build_beams(gvpm_ctx *ctx, float radius) {
  if (!ctx) return GVPM_ERR_INVALID;
  if (!ctx->beams_loaded) return fail(ctx, GVPM_ERR_INVALID, "no beams uploaded");
  if (!(radius > 0.f)) return fail(ctx, GVPM_ERR_INVALID, "radius must be positive");
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const uint32_t nb = ctx->n_beams, n = ctx->n_subs;
  CK(cudaEventRecord(ctx->ev[0], st));
  Tree T{};
  T.n = n;
  uint32_t cnt = (n + 31) / 32, total = 0;
  int levels = 0;
  if (n > 0) {
    for (;;) {
      T.cnt[levels] = cnt;
      T.off[levels] = total;
      total += cnt;
      ++levels;
      if (cnt <= 32) break;
      if (levels >= GVPM_MAX_LEVELS) return fail(ctx, GVPM_ERR_INVALID, "too many sub-beams");
      cnt = (cnt + 31) / 32;
    }
  }
  T.levels = levels;
  CK(ctx->beam_bounds.reserve(256));
  if (n > 0) {
    CK(ctx->sub_pos.reserve(12 * (size_t)n));
    CK(ctx->sub_raw.reserve(16 * (size_t)n));
    CK(ctx->subs.reserve(16 * (size_t)n));
    CK(ctx->keys_in.reserve(8 * (size_t)n));
    CK(ctx->keys_out.reserve(8 * (size_t)n));
    CK(ctx->vals_in.reserve(4 * (size_t)n));
    CK(ctx->vals_out.reserve(4 * (size_t)n));
    CK(ctx->sort_temp.reserve(sort_temp_bytes(n)));
    CK(ctx->bounds_partial.reserve((size_t)bounds_blocks(n) * 6 * sizeof(float)));
    CK(ctx->beam_box_lo.reserve(16 * (size_t)total));
    CK(ctx->beam_box_hi.reserve(16 * (size_t)total));
    // sub-beam cuts on the device (pack_prims.cu): per-beam counts -> offsets -> (midpoint, t1, t2, beam, flags)
    const float *subsize = ctx->beam_aux.as<float>();
    uint32_t *aux = ctx->beam_aux.as<uint32_t>();
    launch_sub_count(ctx->beam_len.as<float>(), nb, subsize, aux + 16, aux + 1, st);
    launch_sub_emit(ctx->beams.as<float4>(), ctx->beam_len.as<float>(), nb, subsize, aux + 16, n, ctx->sub_pos.as<float>(),
                    ctx->sub_raw.as<float4>(), st);
    launch_bounds(ctx->sub_pos.as<float>(), n, ctx->bounds_partial.as<float>(), ctx->beam_bounds.as<float>(), st);
    launch_morton(ctx->sub_pos.as<float>(), n, ctx->beam_bounds.as<float>(), ctx->keys_in.as<uint32_t>(),
                  ctx->vals_in.as<uint32_t>(), st);
    CK(run_sort(ctx->sort_temp.p, ctx->sort_temp.cap, ctx->keys_in.as<uint32_t>(), ctx->keys_out.as<uint32_t>(),
                ctx->vals_in.as<uint32_t>(), ctx->vals_out.as<uint32_t>(), n, st));
    launch_sub_gather(ctx->sub_raw.as<float4>(), ctx->vals_out.as<uint32_t>(), n, ctx->subs.as<float4>(), st);
    float4 *lo = ctx->beam_box_lo.as<float4>(), *hi = ctx->beam_box_hi.as<float4>();
    launch_subbeam_leaf_boxes(ctx->subs.as<float4>(), ctx->beams.as<float4>(), n, T.cnt[0], radius, lo, hi, st);
    for (int l = 1; l < levels; ++l)
      launch_level_boxes(lo + T.off[l - 1], hi + T.off[l - 1], T.cnt[l - 1], T.cnt[l], lo + T.off[l], hi + T.off[l], st);
    ctx->launches += 8 + (levels - 1) + 4;
    CK(cudaGetLastError());
  } else {
    CK(cudaMemsetAsync(ctx->beam_bounds.p, 0, 7 * sizeof(float), st));
  }
  T.lo = ctx->beam_box_lo.as<float4>();
  T.hi = ctx->beam_box_hi.as<float4>();
  ctx->beam_tree = T;
  ctx->beam_radius = radius;
  ctx->beams_built = true;
  ++ctx->state_gen;
  CK(cudaEventRecord(ctx->ev[1], st));
  ctx->timed_build = true;
  return GVPM_OK;
}

uint64_t gvpm_beam_subbeam_count(const gvpm_ctx *ctx) { return ctx ? ctx->n_subs : 0; }

static int beam_params(gvpm_ctx *ctx, GatherParams &P) {
  if (!ctx->have_medium || !ctx->have_cfg) return fail(ctx, GVPM_ERR_INVALID, "medium/config not set");
  if (!ctx->beams_built) return fail(ctx, GVPM_ERR_INVALID, "gvpm_build_beams has not been called");
  if (!ctx->rays_loaded) return fail(ctx, GVPM_ERR_INVALID, "no rays uploaded");
  memset(&P, 0, sizeof(P));
  P.tree = ctx->beam_tree;
  P.rays = ctx->rays.as<float4>();
  P.n_rays = ctx->n_rays;
  const float r = ctx->beam_radius;
  P.radius = r;
  P.radius_sq = r * r;
  P.kernel_vol = (float)((4.0 / 3.0) * (double)GVPM_PI * std::pow((double)r, 3));
  P.weight_kernel = (float)(1.0 / P.kernel_vol);  // shift_volume_beams.h:272
  if (ctx->cfg.beam_kernel_1d) P.weight_kernel = 0.5f / r;  // :188
  P.bounds = ctx->beam_bounds.as<float>();
  for (int i = 0; i < 3; ++i) {
    P.sigma_s[i] = ctx->medium.sigma_s[i];
    P.sigma_t[i] = ctx->medium.sigma_s[i] + ctx->medium.sigma_a[i];
  }
  P.phase_type = ctx->medium.phase_type;
  P.hg_g = ctx->medium.hg_g;
  P.sampling_weight = ctx->medium.sampling_weight;
  P.cfg = ctx->cfg;
  P.tri = ctx->tri.as<float>();
  P.tri_plane = ctx->tri_plane.as<float4>();
  P.tri_aux = ctx->tri_aux.as<float2>();
  P.n_tri = ctx->n_tri;
  P.out = ctx->out.as<float>();
  P.counts = ctx->counts.as<uint32_t>();
  P.work_counter = ctx->work_counter.as<uint32_t>();
  P.pair_counter = (unsigned long long *)(ctx->work_counter.as<char>() + 8);
  P.dump_counter = (unsigned long long *)(ctx->work_counter.as<char>() + 16);
  P.ray_begin = 0;
  P.ray_end = ctx->n_rays;
  P.beams = ctx->beams.as<float4>();
  P.subs = ctx->subs.as<float4>();
  P.n_beams = ctx->n_beams;
  P.sppm_beam_technique = -1;
  return GVPM_OK;
}

// traverse + shade with pair-list growth; dump != nullptr: fill the flat geometric pair list instead of shading
static int beams_run(gvpm_ctx *ctx, GatherParams &P, bool want_counts) {
  const size_t nr = ctx->n_rays;
  if (ctx->pair_cap == 0) {
    int rc = reserve_pairs(ctx, std::max<size_t>(1u << 20, 64 * nr));
    if (rc) return rc;
  }
  const bool sppm = P.sppm_beam_technique >= 0;
  P.beam_prefilter = (want_counts || P.dump_pairs || sppm) ? 0 : 1;
  unsigned long long total = 0;
  for (int attempt = 0; attempt < 4; ++attempt) {
    P.pairs = ctx->pairs.as<uint2>();
    P.pair_cap = ctx->pair_cap;
    CK(cudaMemsetAsync(ctx->work_counter.p, 0, 32, ctx->stream));
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    CK(launch_beam_traverse(P, ctx->sm_count, ctx->stream));
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->split_timed = true;
  ctx->split_timed = true;
    ctx->launches += nr ? 1 : 0;
    CK(cudaMemcpyAsync(ctx->pair_count_host, P.pair_counter, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    total = *ctx->pair_count_host;
    if (total <= ctx->pair_cap) break;
    if (attempt == 3) return fail(ctx, GVPM_ERR_CUDA, "beam pair list overflow");
    int rc = reserve_pairs(ctx, total + total / 8 + 1024);
    if (rc) return rc;
  }
  ctx->last_pairs = total;
  if (nr) {
    CK(cudaMemsetAsync(ctx->out.p, 0, nr * GVPM_OUT_FLOATS * sizeof(float), ctx->stream));
    CK(cudaMemsetAsync(ctx->counts.p, 0, nr * 8, ctx->stream));
  }
  CK(sppm ? launch_beam_shade_sppm(P, total, ctx->sm_count, ctx->stream)
          : launch_beam_shade(P, total, ctx->sm_count, ctx->stream));
  ctx->launches += total ? 1 : 0;
  return GVPM_OK;
}

int gvpm_gather_beams_device(gvpm_ctx *ctx, const float **out_dev, const uint32_t **counts_dev) {
  if (!ctx) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  GatherParams P;
  int rc = beam_params(ctx, P);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  rc = beams_run(ctx, P, counts_dev != nullptr);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed_gather = true;
  if (out_dev) *out_dev = ctx->out.as<float>();
  if (counts_dev) *counts_dev = ctx->counts.as<uint32_t>();
  return GVPM_OK;
}

int gvpm_gather_beams(gvpm_ctx *ctx, float *out, uint32_t *counts) {
  if (!ctx || !out) return GVPM_ERR_INVALID;
  const uint32_t *cd = nullptr;
  int rc = gvpm_gather_beams_device(ctx, nullptr, counts ? &cd : nullptr);
  if (rc) return rc;
  const size_t n = ctx->n_rays;
  if (n) {
    CK(cudaMemcpyAsync(out, ctx->out.p, n * GVPM_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) CK(cudaMemcpyAsync(counts, ctx->counts.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

int gvpm_dump_neighbours_beams(gvpm_ctx *ctx, uint64_t *offsets, uint32_t *idx, size_t cap) {
  if (!ctx || !offsets) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t n = ctx->n_rays;
  std::vector<float> tmp(n * GVPM_OUT_FLOATS + 1);
  std::vector<uint32_t> counts(2 * n + 2);
  int rc = gvpm_gather_beams(ctx, tmp.data(), counts.data());
  if (rc) return rc;
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) { offsets[i] = total; total += counts[2 * i]; }
  offsets[n] = total;
  if (total > cap || (total && !idx)) return fail(ctx, GVPM_ERR_INVALID, "neighbour buffer too small");
  if (total == 0) return GVPM_OK;
  CK(ctx->nbr_idx.reserve(total * sizeof(uint2)));
  GatherParams P;
  rc = beam_params(ctx, P);
  if (rc) return rc;
  P.counts = nullptr;
  P.dump_pairs = ctx->nbr_idx.as<uint2>();
  P.dump_cap = total;
  rc = beams_run(ctx, P, true);
  if (rc) return rc;
  std::vector<uint2> flat(total);
  CK(cudaMemcpyAsync(flat.data(), ctx->nbr_idx.p, total * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // bucket by ray, ascending beam index inside a ray
  std::vector<uint64_t> cursor(offsets, offsets + n);
  for (const uint2 &p : flat) idx[cursor[p.x]++] = p.y;
  for (size_t i = 0; i < n; ++i)
    std::sort(idx + offsets[i], idx + offsets[i + 1],
              [](uint32_t a, uint32_t b) { return (a & 0x7fffffffu) < (b & 0x7fffffffu); });
  return GVPM_OK;
}

// ---- sppm primal photon beams (sppm.cpp:823-860 + beams.h:29-223) ------------------------------------------------
static int sppm_beam_params(gvpm_ctx *ctx, int technique, GatherParams &P) {
  if (technique < GVPM_BEAM_1D || technique > GVPM_BEAM_3D_OPTIMIZED)
    return fail(ctx, GVPM_ERR_INVALID, "unknown gvpm_beam_technique");
  int rc = beam_params(ctx, P);
  if (rc) return rc;
  P.sppm_beam_technique = technique;
  return GVPM_OK;
}

int gvpm_gather_sppm_beams(gvpm_ctx *ctx, int technique, float *out, uint32_t *counts) {
  if (!ctx || !out) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  GatherParams P;
  int rc = sppm_beam_params(ctx, technique, P);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  rc = beams_run(ctx, P, counts != nullptr);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed_gather = true;
  const size_t n = ctx->n_rays;
  if (n) {
    CK(cudaMemcpyAsync(out, ctx->out.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) CK(cudaMemcpyAsync(counts, ctx->counts.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

int gvpm_dump_neighbours_sppm_beams(gvpm_ctx *ctx, int technique, uint64_t *offsets, uint32_t *idx, size_t cap) {
  if (!ctx || !offsets) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t n = ctx->n_rays;
  std::vector<float> tmp(n * 3 + 1);
  std::vector<uint32_t> counts(2 * n + 2);
  int rc = gvpm_gather_sppm_beams(ctx, technique, tmp.data(), counts.data());
  if (rc) return rc;
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) { offsets[i] = total; total += counts[2 * i]; }
  offsets[n] = total;
  if (total > cap || (total && !idx)) return fail(ctx, GVPM_ERR_INVALID, "neighbour buffer too small");
  if (total == 0) return GVPM_OK;
  CK(ctx->nbr_idx.reserve(total * sizeof(uint2)));
  GatherParams P;
  rc = sppm_beam_params(ctx, technique, P);
  if (rc) return rc;
  P.counts = nullptr;
  P.dump_pairs = ctx->nbr_idx.as<uint2>();
  P.dump_cap = total;
  rc = beams_run(ctx, P, true);
  if (rc) return rc;
  std::vector<uint2> flat(total);
  CK(cudaMemcpyAsync(flat.data(), ctx->nbr_idx.p, total * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<uint64_t> cursor(offsets, offsets + n);
  for (const uint2 &p : flat) idx[cursor[p.x]++] = p.y;
  // ascending beam index inside a ray; the naive technique can list a beam once per accepted sub-beam, entries of
  // one beam ordered by the filter bit so that the list is canonical
  for (size_t i = 0; i < n; ++i)
    std::sort(idx + offsets[i], idx + offsets[i + 1], [](uint32_t a, uint32_t b) {
      const uint32_t ai = a & 0x7fffffffu, bi = b & 0x7fffffffu;
      return ai != bi ? ai < bi : a < b;
    });
  return GVPM_OK;
}

// ---- G-VPM -------------------------------------------------------------------------------------
struct SampleLayout {
  size_t off[6], bytes;
  explicit SampleLayout(size_t n) {
    const size_t sz[6] = {4 * n, 4 * n, 12 * n, 4 * n, 4 * n, 4 * n};
    size_t o = 0;
    for (int i = 0; i < 6; ++i) { off[i] = o; o += align256(sz[i]); }
    bytes = o;
  }
};
// The six arrays go up as they are and one kernel packs them (2 float4 per sample), finds the largest sample radius and
// checks the ray indices (pack_prims.cu); both are looked at after the copy of two words at the end of this call.
int gvpm_upload_vpm_samples(gvpm_ctx *ctx, const gvpm_vpm_sample_soa *s, size_t n) {
  if (!ctx || (n && !s) || n > 0xfffffff0u) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  ctx->n_samples = (uint32_t)n;
  ctx->samples_loaded = true;
  ctx->sample_radius_max = 0.f;
  ++ctx->state_gen;
  if (n == 0) return GVPM_OK;
  if (!s->ray || !s->t || !s->transmittance || !s->pdf_success || !s->pdf_sel || !s->radius)
    return fail(ctx, GVPM_ERR_INVALID, "null array in gvpm_vpm_sample_soa");
  SampleLayout L(n);
  CK(ctx->sample_staging.reserve(L.bytes));
  CK(ctx->samples.reserve(2 * n * sizeof(float4)));
  CK(ctx->sample_counts.reserve(2 * n * sizeof(uint32_t)));
  const void *src[6] = {s->ray, s->t, s->transmittance, s->pdf_success, s->pdf_sel, s->radius};
  const size_t sz[6] = {4 * n, 4 * n, 12 * n, 4 * n, 4 * n, 4 * n};
  char *stg = (char *)ctx->sample_staging.p;
  for (int i = 0; i < 6; ++i) CK(cudaMemcpyAsync(stg + L.off[i], src[i], sz[i], cudaMemcpyHostToDevice, ctx->stream));
  SampleStaging S;
  S.ray = (const uint32_t *)(stg + L.off[0]); S.t = (const float *)(stg + L.off[1]);
  S.transmittance = (const float *)(stg + L.off[2]); S.pdf_success = (const float *)(stg + L.off[3]);
  S.pdf_sel = (const float *)(stg + L.off[4]); S.radius = (const float *)(stg + L.off[5]);
  uint32_t *stats = ctx->work_counter.as<uint32_t>() + 8;   // words 8, 9 of the counter block
  CK(cudaMemsetAsync(stats, 0, 8, ctx->stream));
  launch_pack_samples(S, (uint32_t)n, ctx->n_rays, ctx->samples.as<float4>(), stats, ctx->stream);
  ctx->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->sample_stats_host, stats, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // the caller's arrays are only read during the call
  if (ctx->sample_stats_host[1]) {
    ctx->samples_loaded = false;
    return fail(ctx, GVPM_ERR_INVALID, "sample refers to a ray that is not uploaded");
  }
  memcpy(&ctx->sample_radius_max, &ctx->sample_stats_host[0], 4);
  return GVPM_OK;
}

static int vpm_params(gvpm_ctx *ctx, GatherParams &P, int nb_camera_samples) {
  if (ctx->built && ctx->accel == gvpm_ctx::ACCEL_FRUSTUM) {
    // the range queries walk the box hierarchy: build it (over the photons the rays can reach) in place of the grid
    int rcb = build_pruned_bvh(ctx, ctx->radius, nullptr);
    if (rcb) return rcb;
  }
  int rc = fill_params(ctx, P, ctx->out.as<float>(), nullptr);
  if (rc) return rc;
  if (!ctx->samples_loaded) return fail(ctx, GVPM_ERR_INVALID, "no VPM samples uploaded");
  if (nb_camera_samples <= 0) return fail(ctx, GVPM_ERR_INVALID, "nbCameraSamples must be positive");
  if (ctx->sample_radius_max > ctx->radius)
    return fail(ctx, GVPM_ERR_INVALID, "gvpm_build_points radius is smaller than a sample radius");
  P.samples = ctx->samples.as<float4>();
  P.n_samples = ctx->n_samples;
  P.sample_counts = ctx->sample_counts.as<uint32_t>();
  P.vpm_normalization = 1.0f / (float)nb_camera_samples;
  return GVPM_OK;
}

int gvpm_gather_vpm_device(gvpm_ctx *ctx, int nb_camera_samples, const float **out_dev, const uint32_t **mvol_dev) {
  if (!ctx) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  GatherParams P;
  int rc = vpm_params(ctx, P, nb_camera_samples);
  if (rc) return rc;
  const size_t nr = ctx->n_rays, ns = ctx->n_samples;
  CK(ctx->mvol.reserve(nr * 4 + 256));
  P.mvol = ctx->mvol.as<uint32_t>();
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (ctx->pair_cap == 0) {
    rc = reserve_pairs(ctx, std::max<size_t>(1u << 20, 8 * ns));
    if (rc) return rc;
  }
  unsigned long long total = 0;
  for (int attempt = 0; attempt < 3; ++attempt) {
    P.pairs = ctx->pairs.as<uint2>();
    P.pair_cap = ctx->pair_cap;
    CK(cudaMemsetAsync(ctx->work_counter.p, 0, 16, ctx->stream));
    if (nr) CK(cudaMemsetAsync(ctx->out.p, 0, nr * GVPM_OUT_FLOATS * sizeof(float), ctx->stream));
    if (nr) CK(cudaMemsetAsync(ctx->mvol.p, 0, nr * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    CK(launch_vpm_traverse(P, false, ctx->sm_count, ctx->stream));
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->split_timed = true;
  ctx->split_timed = true;
    ctx->launches += ns ? 1 : 0;
    CK(cudaMemcpyAsync(ctx->pair_count_host, P.pair_counter, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    total = *ctx->pair_count_host;
    if (total <= ctx->pair_cap) break;
    if (attempt == 2) return fail(ctx, GVPM_ERR_CUDA, "VPM pair list overflow");
    rc = reserve_pairs(ctx, total + total / 8 + 1024);
    if (rc) return rc;
  }
  ctx->last_pairs = total;
  CK(launch_vpm_shade(P, total, ctx->sm_count, ctx->stream));
  ctx->launches += total ? 1 : 0;
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed_gather = true;
  if (out_dev) *out_dev = ctx->out.as<float>();
  if (mvol_dev) *mvol_dev = ctx->mvol.as<uint32_t>();
  return GVPM_OK;
}

int gvpm_gather_vpm(gvpm_ctx *ctx, int nb_camera_samples, float *out, uint32_t *mvol, uint32_t *sample_counts) {
  if (!ctx || !out) return GVPM_ERR_INVALID;
  int rc = gvpm_gather_vpm_device(ctx, nb_camera_samples, nullptr, nullptr);
  if (rc) return rc;
  const size_t nr = ctx->n_rays, ns = ctx->n_samples;
  if (nr) {
    CK(cudaMemcpyAsync(out, ctx->out.p, nr * GVPM_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (mvol) CK(cudaMemcpyAsync(mvol, ctx->mvol.p, nr * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (ns && sample_counts)
    CK(cudaMemcpyAsync(sample_counts, ctx->sample_counts.p, ns * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

int gvpm_dump_neighbours_vpm(gvpm_ctx *ctx, int nb_camera_samples, uint64_t *offsets, uint32_t *idx, size_t cap) {
  if (!ctx || !offsets) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t ns = ctx->n_samples;
  std::vector<float> tmp((size_t)ctx->n_rays * GVPM_OUT_FLOATS + 1);
  std::vector<uint32_t> counts(2 * ns + 2);
  int rc = gvpm_gather_vpm(ctx, nb_camera_samples, tmp.data(), nullptr, counts.data());
  if (rc) return rc;
  uint64_t total = 0;
  for (size_t i = 0; i < ns; ++i) { offsets[i] = total; total += counts[2 * i]; }
  offsets[ns] = total;
  if (total > cap || (total && !idx)) return fail(ctx, GVPM_ERR_INVALID, "neighbour buffer too small");
  if (total == 0) return GVPM_OK;
  CK(ctx->nbr_offsets.reserve((ns + 1) * 8));
  CK(ctx->nbr_idx.reserve(total * 4));
  CK(cudaMemcpyAsync(ctx->nbr_offsets.p, offsets, (ns + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  GatherParams P;
  rc = vpm_params(ctx, P, nb_camera_samples);
  if (rc) return rc;
  P.sample_counts = nullptr;
  P.nbr_offsets = ctx->nbr_offsets.as<uint64_t>();
  P.nbr_idx = ctx->nbr_idx.as<uint32_t>();
  CK(cudaMemsetAsync(ctx->work_counter.p, 0, 16, ctx->stream));
  CK(launch_vpm_traverse(P, true, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(idx, ctx->nbr_idx.p, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

static int compute_gradient_impl(gvpm_ctx *ctx, const float *acc, int w, int h, int use_abs, int reuse, float inv_emitted,
                                 float *throughput, float *gx, float *gy);
int gvpm_compute_gradient(gvpm_ctx *ctx, const float *acc, int w, int h, int use_abs, float *throughput,
                          float *gx, float *gy) {
  return compute_gradient_impl(ctx, acc, w, h, use_abs, 0, 1.f, throughput, gx, gy);
}
int gvpm_compute_gradient_reuse_primal(gvpm_ctx *ctx, const float *acc, int w, int h, int use_abs, float inv_emitted,
                                       float *throughput, float *gx, float *gy) {
  if (!(inv_emitted > 0.f)) return GVPM_ERR_INVALID;
  return compute_gradient_impl(ctx, acc, w, h, use_abs, 1, inv_emitted, throughput, gx, gy);
}
static int compute_gradient_impl(gvpm_ctx *ctx, const float *acc, int w, int h, int use_abs, int reuse, float inv_emitted,
                                 float *throughput, float *gx, float *gy) {
  if (!ctx || !acc || w <= 0 || h <= 0 || !throughput || !gx || !gy) return GVPM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t np = (size_t)w * h;
  CK(ctx->grad_in.reserve(np * GVPM_OUT_FLOATS * 4));
  CK(ctx->grad_out.reserve(np * 9 * 4));
  CK(cudaMemcpyAsync(ctx->grad_in.p, acc, np * GVPM_OUT_FLOATS * 4, cudaMemcpyHostToDevice, ctx->stream));
  float *o = ctx->grad_out.as<float>();
  launch_gradient(ctx->grad_in.as<float>(), w, h, use_abs, o, o + 3 * np, o + 6 * np, ctx->stream, reuse, inv_emitted);
  ctx->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(throughput, o, np * 12, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(gx, o + 3 * np, np * 12, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(gy, o + 6 * np, np * 12, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GVPM_OK;
}

// ---- screened-Poisson reconstruction (row f-3): poisson::Solver of the reference, gvpm.cpp:610-690 -------------------
int gvpm_poisson_preset(const char *preset, gvpm_poisson_params *p) {
  if (!preset || !p) return GVPM_ERR_INVALID;
  // Solver::Params::Params + setConfigPreset, Solver.cpp:59-158
  p->alpha = 0.2f;
  p->irls_iter_max = 1; p->irls_reg_init = 0.f; p->irls_reg_iter = 0.f;
  p->cg_iter_max = 1; p->cg_iter_check = 100; p->cg_precond = 0; p->cg_tolerance = 0.f;
  if (!strcmp(preset, "L1D")) { p->irls_iter_max = 20; p->irls_reg_init = 0.05f; p->irls_reg_iter = 0.5f; p->cg_iter_max = 50; }
  else if (!strcmp(preset, "L1Q")) { p->irls_iter_max = 64; p->irls_reg_init = 1.0f; p->irls_reg_iter = 0.7f; p->cg_iter_max = 1000; }
  else if (!strcmp(preset, "L1L")) { p->irls_iter_max = 7; p->irls_reg_init = 1.0e-4f; p->irls_reg_iter = 1.0e-1f; p->cg_iter_max = 20000; p->cg_tolerance = 1.0e-20f; }
  else if (!strcmp(preset, "L2D")) { p->cg_iter_max = 50; }
  else if (!strcmp(preset, "L2Q")) { p->cg_iter_max = 500; }
  else return GVPM_ERR_INVALID;
  return GVPM_OK;
}

int gvpm_poisson_solve(gvpm_ctx *ctx, int w, int h, const float *throughput, const float *dx, const float *dy,
                       const float *direct, const gvpm_poisson_params *params, float *reconstruction) {
  if (!ctx || w <= 0 || h <= 0 || !dx || !dy || !params || !reconstruction) return GVPM_ERR_INVALID;
  if (params->cg_precond)
    return fail(ctx, GVPM_ERR_UNSUPPORTED, "the preconditioned CG branch (cgPrecond, enabled by no preset) is not built");
  cudaSetDevice(ctx->device);
  // Solver::Params::sanitize, Solver.cpp:162-172
  const float alpha = std::max(params->alpha, 0.0f);
  const int irlsIterMax = std::max(params->irls_iter_max, 1), cgIterMax = std::max(params->cg_iter_max, 1),
            cgIterCheck = std::max(params->cg_iter_check, 1);
  const float irlsRegInit = std::max(params->irls_reg_init, 0.0f), irlsRegIter = std::max(params->irls_reg_iter, 0.0f),
              cgTolerance = std::max(params->cg_tolerance, 0.0f);
  const size_t n = (size_t)w * h, plane = 3 * n * sizeof(float);
  CK(ctx->poisson_io.reserve(5 * plane));
  CK(ctx->poisson_ws.reserve(poisson_workspace_floats(n) * sizeof(float)));
  float *io = ctx->poisson_io.as<float>();
  float *d_tp = io, *d_dx = io + 3 * n, *d_dy = io + 6 * n, *d_direct = io + 9 * n, *d_rec = io + 12 * n;
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[2], st));
  if (throughput) CK(cudaMemcpyAsync(d_tp, throughput, plane, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_dx, dx, plane, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_dy, dy, plane, cudaMemcpyHostToDevice, st));
  if (direct) CK(cudaMemcpyAsync(d_direct, direct, plane, cudaMemcpyHostToDevice, st));
  const long long launched = poisson_solve_device(throughput ? d_tp : nullptr, d_dx, d_dy, direct ? d_direct : nullptr, w, h,
                                                  alpha, irlsIterMax, irlsRegInit, irlsRegIter, cgIterMax, cgIterCheck,
                                                  cgTolerance, ctx->poisson_ws.as<float>(), (float *)ctx->pair_count_host,
                                                  d_rec, st);
  if (launched < 0) return fail(ctx, GVPM_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
  ctx->launches += (uint64_t)launched;
  CK(cudaMemcpyAsync(reconstruction, d_rec, plane, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->poisson_ms, ctx->ev[2], ctx->ev[3]));
  ctx->timed_gather = false;
  return GVPM_OK;
}

// computeGradient + reconstruction in one call: the three planes never leave the device (gvpm.cpp:554-690)
int gvpm_reconstruct(gvpm_ctx *ctx, const float *acc, int w, int h, int use_abs, const float *direct,
                     const gvpm_poisson_params *params, float *throughput, float *gx, float *gy, float *reconstruction) {
  if (!ctx || !acc || w <= 0 || h <= 0 || !params || !reconstruction) return GVPM_ERR_INVALID;
  if (params->cg_precond)
    return fail(ctx, GVPM_ERR_UNSUPPORTED, "the preconditioned CG branch (cgPrecond, enabled by no preset) is not built");
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)w * h, plane = 3 * n * sizeof(float);
  CK(ctx->grad_in.reserve(n * GVPM_OUT_FLOATS * 4));
  CK(ctx->grad_out.reserve(n * 9 * 4));
  CK(ctx->poisson_io.reserve(5 * plane));
  CK(ctx->poisson_ws.reserve(poisson_workspace_floats(n) * sizeof(float)));
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[2], st));
  CK(cudaMemcpyAsync(ctx->grad_in.p, acc, n * GVPM_OUT_FLOATS * 4, cudaMemcpyHostToDevice, st));
  float *o = ctx->grad_out.as<float>(), *io = ctx->poisson_io.as<float>();
  float *d_direct = io + 9 * n, *d_rec = io + 12 * n;
  if (direct) CK(cudaMemcpyAsync(d_direct, direct, plane, cudaMemcpyHostToDevice, st));
  launch_gradient(ctx->grad_in.as<float>(), w, h, use_abs, o, o + 3 * n, o + 6 * n, st);
  ctx->launches += 1;
  const long long launched = poisson_solve_device(
      o, o + 3 * n, o + 6 * n, direct ? d_direct : nullptr, w, h, std::max(params->alpha, 0.0f),
      std::max(params->irls_iter_max, 1), std::max(params->irls_reg_init, 0.0f), std::max(params->irls_reg_iter, 0.0f),
      std::max(params->cg_iter_max, 1), std::max(params->cg_iter_check, 1), std::max(params->cg_tolerance, 0.0f),
      ctx->poisson_ws.as<float>(), (float *)ctx->pair_count_host, d_rec, st);
  if (launched < 0) return fail(ctx, GVPM_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
  ctx->launches += (uint64_t)launched;
  if (throughput) CK(cudaMemcpyAsync(throughput, o, plane, cudaMemcpyDeviceToHost, st));
  if (gx) CK(cudaMemcpyAsync(gx, o + 3 * n, plane, cudaMemcpyDeviceToHost, st));
  if (gy) CK(cudaMemcpyAsync(gy, o + 6 * n, plane, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(reconstruction, d_rec, plane, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->poisson_ms, ctx->ev[2], ctx->ev[3]));
  ctx->timed_gather = false;
  return GVPM_OK;
}

float gvpm_last_poisson_ms(const gvpm_ctx *ctx) { return ctx ? ctx->poisson_ms : 0.f; }

int gvpm_last_timings(gvpm_ctx *ctx, float *build_ms, float *gather_ms) {
  if (!ctx) return GVPM_ERR_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  { int rc = pending_check(ctx, true); if (rc) return rc; }
  if (ctx->timed_build) CK(cudaEventElapsedTime(&ctx->build_ms, ctx->ev[0], ctx->ev[1]));
  if (ctx->timed_gather) CK(cudaEventElapsedTime(&ctx->gather_ms, ctx->ev[2], ctx->ev[3]));
  if (build_ms) *build_ms = ctx->build_ms;
  if (gather_ms) *gather_ms = ctx->gather_ms;
  return GVPM_OK;
}

int gvpm_last_gather_detail(gvpm_ctx *ctx, float *traverse_ms, float *shade_ms, uint64_t *pairs) {
  if (!ctx) return GVPM_ERR_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  { int rc = pending_check(ctx, true); if (rc) return rc; }
  float t = 0.f, s = 0.f;
  if (ctx->timed_gather && ctx->n_rays && ctx->split_timed) {
    CK(cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]));
    CK(cudaEventElapsedTime(&s, ctx->ev[5], ctx->ev[3]));
  }
  if (traverse_ms) *traverse_ms = t;
  if (shade_ms) *shade_ms = s;
  if (pairs) *pairs = ctx->last_pairs;
  return GVPM_OK;
}

uint64_t gvpm_launch_count(const gvpm_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
