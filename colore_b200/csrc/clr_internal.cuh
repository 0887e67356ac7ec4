// colore_b200 internal declarations: context, device-side table lookups, RNG, HEALPix arithmetic.
// sm_100a only. Reference citations are file:line under damonge/CoLoRe src/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/colore_b200.h"

#define CLR_SM_COUNT_FALLBACK 148
#define CLR_MAX_PEERS 16

// ---------------------------------------------------------------------------------------------
// device-visible parameter block (passed by value as __grid_constant__ where needed)
struct ClrDev {
  int n, nc, nz_here, iz0_here;     // grid side, n/2+1, slab
  // INTERNAL row pitch: ncp = nc rounded up to a multiple of 8 complex numbers, pitch = 2*ncp floats, so every row
  // starts on a 64-byte boundary (float4 streaming accesses, 16-byte cp.async, bulk-copy / TMA alignment, whole
  // 32-byte sectors per 8-mode tile). The reference layout of the host side (2*nc floats per row, fourier.c:46-51)
  // is converted in clr_grid_put / clr_grid_get.
  int ncp, pitch;
  int log2n;                        // log2(n) when n is a power of two, else -1 (index arithmetic)
  int nyl, ky0;                     // k-space slab of this rank: ky in [ky0, ky0+nyl), layout [kz][ky_local][kx]
  int bias_model, nside_base;
  float l_box;
  double pos_obs[3];
  double glob_idr, r_tab_max;       // 1/dr of the NA tables; r_arr[NA-1]
  double fgrowth_0, hubble_0, OmegaM, r_max;
  const double *r_arr, *z_arr, *d1_arr, *d2_arr, *v1_arr, *pd_arr, *ih_arr, *a2r_a, *a2r_r;
  const float *slice_left, *slice_right;   // z-halo planes of the potential (fourier.c:401-414)
  const float *z_f, *d1_f;                 // fp32 copies of z(r), D(r) for the streaming field kernels
  const float2 *d1_t;                      // {D(r_i), D(r_i+1)-D(r_i)}: one load per lerp
  // cell-node coordinates relative to the observer, tabulated once on the host with the reference's
  // own expressions: cf[ax][i] = (flouble)((i+0.0)*dx_f - pos_obs[ax]) (density.c:1087-1094, dx flouble)
  // and cd[ax][i] = (i+0.0)*dx_d - pos_obs[ax] (srcs.c:159-167, dx double); i is the GLOBAL index.
  const float *cf[3];
  const double *cd[3];
};

struct ClrPop {           // one tracer population: tables on the NA r-grid
  const double *nz, *bz, *norm;   // n(z) [or T(z)], b(z), normalisation
  double norm_0, norm_f;
};

// ---------------------------------------------------------------------------------------------
struct StageTime { float ms = 0; int launches = 0; };

struct clr_ctx {
  int device = 0, sm_count = CLR_SM_COUNT_FALLBACK;
  cudaStream_t stream = nullptr;
  clr_params p;                     // host copy (pointers re-targeted to host vectors below)
  std::vector<double> h_logk, h_pk, h_r, h_z, h_d1, h_d2, h_v1, h_pd, h_ih, h_a2r_a, h_a2r_r;
  ClrDev dev;
  // device memory
  float *d_dens = nullptr, *d_npot = nullptr;       // grids (+2 halo planes on npot)
  double *d_tables = nullptr;                        // 9 x NA
  float *d_tables_f = nullptr;                       // fp32 copies: z(r), D(r)
  float *d_coord_f = nullptr;                        // 3 x n cell coordinates (fp32 expression)
  double *d_coord_d = nullptr;                       // 3 x n cell coordinates (fp64 expression)
  double *d_pk = nullptr;                            // logk[numk], pk[numk]
  float2 *d_twiddle = nullptr;                       // exp(+2*pi*i*k/n), k<n
  float2 *d_pkt = nullptr, *d_sincos = nullptr;      // fp32 tables of the fast mode fill (clr_fields.cu)
  double *d_scratch = nullptr;                       // reductions / histograms
  // single-GPU c2r (clr_fft.cu): z-pass output in the kx-tile layout, ticket + per-plane-group counters of the fused y+x pass
  void *d_fft_tmp = nullptr; size_t fft_tmp_bytes = 0;
  unsigned *d_fft_sync = nullptr;
  int srcs_compact = 1;                              // option "srcs_compact": 0 = dense per-cell counts + full-array expansion
  size_t los_hess_bytes = 0;
  void *d_los_hess = nullptr;                        // per-cell Hessian of the potential for the kappa rays (clr_maps.cu), 32 B / cell
  int los_precompute = 1;                            // option "los_precompute": 0 = kappa rays evaluate the Hessian stencil per sample
  int fft_fused = 1;                                 // option "fft_fused": 0 = three separate axis passes
  int fill_fused = 1;                                // option "fill_fused": 0 = stand-alone mode fill + z pass
  int fill_cluster = -1;                             // option "fill_cluster": fused fill + z pass on CTA pairs (-1: where needed, i.e. n = 2048)
  int fill_w = 8;                                    // option "fill_w": kx lines per tile of the fused fill + z pass at n = 1024 (8 or 4)
  size_t scratch_bytes = 0;
  double sigma2_gauss = 0, mean_gauss = 0;
  // normalisation histogram produced by the fused lognormal + histogram pass (clr_fields.cu), valid for ONE later
  // clr_compute_density_normalization on the same field with the same populations
  double *d_hist = nullptr; size_t hist_bytes = 0;
  bool hist_valid = false; int hist_npop = 0, hist_nz = 0; const double *hist_bz[CLR_NPOP_MAX * 2] = {nullptr};
  int hist_fused = 1;                                // option "hist_fused": 0 = separate lognormal and histogram passes
  // populations
  struct Pop {
    bool set = false;
    std::vector<double> h_a, h_b, h_norm;
    double *d_a = nullptr, *d_b = nullptr, *d_norm = nullptr;
    double norm_0 = 1, norm_f = 1;
    bool have_norm = false;
    // imap shells
    int nside = 0, nr = 0;
    std::vector<float> r0, rf;
    // sources catalogue (device)
    // per-cell counts: dense int32 [nz][n][n], or (default) compact entries per 8192-cell super chunk stored in the
    // super chunk's own slice of this buffer + the number of entries per super chunk (clr_srcs.cu: poisson_kernel)
    int32_t *d_counts = nullptr;
    bool counts_compact = false, d_sup_entries_valid = false;
    int32_t *d_sup_entries = nullptr; size_t sup_entries_cap = 0;
    float *d_bound = nullptr;       // fp32 screening table of the Poisson pass (4 floats per r-bin)
    long long nsrc = 0;
    float *d_pos = nullptr; int32_t *d_ipix = nullptr; float *d_srcs = nullptr;
    // async_results: second Src buffer, so that the read-back of run s may take the whole of run s+1
    float *d_srcs_alt = nullptr;
    int srcs_buf = 0;               // which of the two buffers d_srcs currently is
    size_t cap_src = 0;
    // skewers of the last clr_srcs_get_beam_properties (clr_beam.cu): nsrc x skw_nr floats each
    float *d_skw_dg = nullptr, *d_skw_v = nullptr;
    long long skw_n = -1; int skw_nr = 0;
  } srcs[CLR_NPOP_MAX], imap[CLR_NPOP_MAX], cstm[CLR_NPOP_MAX];   // cstm: h_a = K(z) table (cosmo.c:659-664)
  double z0_norm = 0, zf_norm = 0;
  // fast-lensing shells of the last clr_lensing_get_beam_properties (clr_beam.cu), kept for the source interpolation
  float *d_lens_data = nullptr; long long lens_total = 0; int lens_nbeams = 0; std::vector<long long> lens_npp;
  // multi-GPU
  int rank = 0, nranks = 1;
  void *nccl_comm = nullptr;
  float *d_stage = nullptr;         // all-to-all staging buffer of the distributed FFT (one slab)
  // peer-memory transpose: every rank's staging buffer mapped into this process (CUDA IPC over NVLink), so the
  // FFT pass that produces the data stores it straight into the destination GPU (no separate all-to-all)
  float *peer_stage[CLR_MAX_PEERS] = {nullptr};
  bool p2p = false;
  int p2p_enabled = 1;              // option "p2p_fused"
  // Second transform pipeline (option "fft_overlap"): the c2r of the potential runs on its own stream with its own
  // staging buffer and flag barriers, so its NVLink-bound z pass overlaps the HBM-bound y / x passes, the lognormal
  // transform and the Poisson pass of the density on the main stream. Consumers of the potential call clr_npot_ready.
  struct StageSet { float *stage = nullptr; float *peer[CLR_MAX_PEERS] = {nullptr}; unsigned *flag_peer[CLR_MAX_PEERS] = {nullptr}; unsigned epoch = 0; } sets[2];
  int cur_set = 0;
  size_t stage_floats = 0;          // floats of one staging buffer (the flag words of the barriers sit behind them)
  bool flag_barrier = false;        // stream-ordered barriers through peer-memory flags instead of a 1-int all-reduce
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_z_done = nullptr, ev_npot = nullptr;
  cudaEvent_t ev_after_z = nullptr; // when set, the distributed c2r records it once its z pass has been exchanged
  bool npot_pending = false;        // the potential is still being transformed on stream2 (halo not exchanged yet)
  // Default OFF. Measured on 2 GPUs at 1024^3 (profiles/r2_bench_2gpu_overlap_experiment.json): step 14.56 ms against
  // 14.72 ms, end to end 46.9 ms against 17.2 ms. The passes are persistent kernels whose CTAs hold 132 KB of shared
  // memory and all 64K registers of an SM, so a z pass on the second stream and a y / x pass on the main stream never
  // share an SM: whichever starts first owns the GPU and the two pipelines serialise; the system-scope flag barriers
  // also stall behind the catalogue read-back of the end-to-end loop. A real overlap needs an SM partition between
  // the two pipelines (about 40 SMs saturate NVLink at 8 GPUs).
  int fft_overlap = 0;
  int p2p_tiled = -1;               // option "p2p_tiled": tile-major staging layout of the fused c2r (-1: auto)
  int *d_barrier = nullptr;
  double a2a_bytes = 0;             // bytes this rank has sent through the FFT all-to-all
  // bookkeeping
  long long launches = 0;
  bool profiling = false;
  int exact_math = 0;   // 1: field kernels evaluate the reference's double-precision expressions verbatim
  int lpt_interp_type = 1;          // field_par.lpt_interp_type: 0 NGP, 1 CIC, 2 TSC (common.h:67-69)
  int keep_particles = 0;           // keep the LPT particles on the device for write_lpt (io.c:619-695)
  float *d_lpt_pos[3] = {nullptr, nullptr, nullptr};
  long long lpt_sent = 0, lpt_received = 0;   // particles shipped to / received from other slabs in the last LPT run
  std::map<std::string, StageTime> stage;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr;
  // option async_results: catalogue read-back on its own stream, overlapping the next run
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_srcs_ready = nullptr, ev_copy_done = nullptr;
  // small results (moments, histograms, counters) reach the host through MAPPED pinned memory written by a tiny
  // kernel: a cudaMemcpy D2H on the main stream would queue behind the catalogue read-back on the copy engine
  void *h_small = nullptr, *d_small = nullptr;
  cudaEvent_t ev_buf_free[2] = {nullptr, nullptr};   // last read-back of Src buffer 0 / 1 has finished
  bool buf_busy[2] = {false, false};
  int async_results = 0;
  bool copy_pending = false;
  struct Pending { std::string name; int slot; int nl; };
  std::vector<cudaEvent_t> ev_pool;
  std::vector<Pending> ev_pending;
  int ev_used = 0;
};

void clr_set_error(const char *fmt, ...);
#define CLR_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      clr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)
#define CLR_CHECK(cond, ...)                     \
  do {                                           \
    if (!(cond)) { clr_set_error(__VA_ARGS__); return 1; } \
  } while (0)

// stage profiling: CUDA events around a kernel family on the context stream. Events are only
// RECORDED here (no synchronisation, ~1 us each); clr_get_stage_ms resolves them after a stream
// synchronise, so profiling can stay on inside a timed region.
struct StageScope {
  clr_ctx *c; const char *name; int nl; int slot = -1;
  StageScope(clr_ctx *ctx, const char *nm, int nlaunch) : c(ctx), name(nm), nl(nlaunch) {
    c->launches += nl;
    if (c->profiling) {
      if (c->ev_used + 2 > (int)c->ev_pool.size()) {
        for (int i = 0; i < 64; i++) { cudaEvent_t e; cudaEventCreate(&e); c->ev_pool.push_back(e); }
      }
      slot = c->ev_used;
      c->ev_used += 2;
      cudaEventRecord(c->ev_pool[slot], c->stream);
    }
  }
  ~StageScope() {
    if (slot >= 0) {
      cudaEventRecord(c->ev_pool[slot + 1], c->stream);
      c->ev_pending.push_back({std::string(name), slot, nl});
    }
  }
};

// kernels implemented in the other translation units
int clr_fft_c2r_impl(clr_ctx *c, float *grid, double norm, double *d_moments);
int clr_fft_r2c_impl(clr_ctx *c, float *grid);
int clr_fft_fill_c2r(clr_ctx *c, uint32_t seed, double norm, double *d_moments, bool *ran);
// mixed-radix path for grids that are not powers of two (clr_fft_generic.cu)
bool clr_fft_generic_ok(int n);
int clr_fft_generic_c2r(clr_ctx *c, float2 *g, float norm, double *mom);
int clr_fft_generic_r2c(clr_ctx *c, float2 *g);
int clr_fields_fill(clr_ctx *c, uint32_t seed);
int clr_fields_scale_moments(clr_ctx *c, double *out2);
int clr_fields_lognormal(clr_ctx *c, int clip);
int clr_fields_norm_hist(clr_ctx *c, int npop, const double *const *d_bz, int nz, double idz,
                         unsigned long long *h_n, double *h_z, double *h_b);
int clr_srcs_run(clr_ctx *c, int ipop, uint32_t seed);
int clr_srcs_local(clr_ctx *c, int ipop);
int clr_srcs_beam(clr_ctx *c, int ipop);
int clr_srcs_dense_counts(clr_ctx *c, int ipop, int32_t **d_dense, bool *owned);
int clr_srcs_distribute_impl(clr_ctx *c, int ipop, int beam_first, long long *nsrc_out);
int clr_maps_imap(clr_ctx *c, int ipop, float *h_data, int32_t *h_nadd);
int clr_maps_los(clr_ctx *c, int which, long long num_pix, const double *h_pos, int nplanes, const float *rf,
                 float *h_data);
int clr_beam_cstm(clr_ctx *c, int ipop, long long num_pix, const double *h_pos, float *h_data);
int clr_beam_srcs(clr_ctx *c, int ipop, int has_lensing, int has_skw, int skw_gauss, int rsd_done);
int clr_beam_get_skewers(clr_ctx *c, int ipop, float *h_dg, float *h_v);
int clr_beam_lens_shells(clr_ctx *c, int nbeams, int nr_sh, float *r_sh, const long long *npp, const double *h_pos, float *h_data);
int clr_beam_srcs_from_shells(clr_ctx *c, int ipop, int nr_sh, const float *r_sh, const int *nside_sh, int node, int nnodes,
                              long long *n_bad);
int clr_halo_update(clr_ctx *c);
int clr_npot_ready(clr_ctx *c);     // main stream waits for the potential pipeline, then exchanges the z halo
void clr_use_set(clr_ctx *c, int set);   // select the staging buffer / barrier flags (and stream) of pipeline 0 / 1
int clr_lpt_run(clr_ctx *c, int order);
int clr_lpt_particles(clr_ctx *c, float *x, float *y, float *z);
int clr_comm_destroy(clr_ctx *c);
int clr_comm_alltoall(clr_ctx *c, const void *send, void *recv, size_t block_floats);
int clr_comm_alltoallv(clr_ctx *c, const float *send, const size_t *send_off, const size_t *send_n, float *recv,
                       const size_t *recv_off, const size_t *recv_n);
int clr_comm_barrier(clr_ctx *c);
int clr_comm_all_ok(clr_ctx *c, int ok, const char *what);
int clr_comm_allreduce_f64(clr_ctx *c, double *dbuf, size_t n);
int clr_comm_allreduce_u64(clr_ctx *c, unsigned long long *dbuf, size_t n);
int clr_comm_allreduce_f32(clr_ctx *c, float *dbuf, size_t n);
int clr_comm_allreduce_i32(clr_ctx *c, int *dbuf, size_t n);
int clr_comm_halo(clr_ctx *c);
int clr_comm_dens_halo(clr_ctx *c, float *dst_plane);   // dst <- first density plane of the right neighbour
int clr_ensure_scratch(clr_ctx *c, size_t bytes);
#define CLR_SMALL_BYTES 65536
// host_dst <- dev_src (bytes a multiple of 4; meant for a few KB), ordered after the work queued on c->stream; blocks
int clr_read_small(clr_ctx *c, void *host_dst, const void *dev_src, size_t bytes);

// ---------------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__

// flat unpadded cell index i = ix + n*(iy + n*iz_local) -> (ix, iy, iz_local). 64-bit division by a
// run-time divisor costs ~100 instructions on the GPU, so powers of two take shifts.
__device__ __forceinline__ void clr_cell(const ClrDev &d, long long i, int &ix, int &iy, int &iz)
{
  if (d.log2n >= 0) {
    ix = (int)(i & (d.n - 1));
    iy = (int)((i >> d.log2n) & (d.n - 1));
    iz = (int)(i >> (2 * d.log2n));
  } else {
    long long row = i / d.n;
    ix = (int)(i - row * d.n);
    iz = (int)((unsigned)row / (unsigned)d.n);
    iy = (int)((unsigned)row - (unsigned)iz * (unsigned)d.n);
  }
}

// cosmo.c:30-38 f_of_r_linear
__device__ __forceinline__ double clr_lerp(const ClrDev &d, double r, const double *__restrict__ f,
                                           double f0, double ff)
{
  if (r <= 0) return f0;
  else if (r >= d.r_tab_max) return ff;
  int ir = (int)(r * d.glob_idr);
  double fa = __ldg(f + ir), fb = __ldg(f + ir + 1);
  return fa + (fb - fa) * (r - __ldg(d.r_arr + ir)) * d.glob_idr;
}
// cosmo.c:40-57 end values per tag
__device__ __forceinline__ double clr_bg_z(const ClrDev &d, double r) { return clr_lerp(d, r, d.z_arr, 0.0, __ldg(d.z_arr + CLR_NA - 1)); }
__device__ __forceinline__ double clr_bg_d1(const ClrDev &d, double r) { return clr_lerp(d, r, d.d1_arr, 1.0, __ldg(d.d1_arr + CLR_NA - 1)); }
__device__ __forceinline__ double clr_bg_v1(const ClrDev &d, double r) { return clr_lerp(d, r, d.v1_arr, __ldg(d.v1_arr), __ldg(d.v1_arr + CLR_NA - 1)); }
__device__ __forceinline__ double clr_bg_pd(const ClrDev &d, double r) { return clr_lerp(d, r, d.pd_arr, __ldg(d.pd_arr), __ldg(d.pd_arr + CLR_NA - 1)); }
__device__ __forceinline__ double clr_bg_ih(const ClrDev &d, double r) { return clr_lerp(d, r, d.ih_arr, __ldg(d.ih_arr), __ldg(d.ih_arr + CLR_NA - 1)); }
// cosmo.c:58-73
__device__ __forceinline__ double clr_bg_nz(const ClrDev &d, double r, const double *t) { return clr_lerp(d, r, t, 0.0, 0.0); }
__device__ __forceinline__ double clr_bg_bz(const ClrDev &d, double r, const double *t) { return clr_lerp(d, r, t, __ldg(t), 1.0); }

// cosmo.c:101-112
__device__ __forceinline__ double clr_r_of_z(const ClrDev &d, double z)
{
  double a = 1. / (1 + z);
  if (a >= 1) return 0;
  else if (a <= 0) return __ldg(d.a2r_r);
  int ia = (int)(a * (CLR_NA - 1));
  double r0 = __ldg(d.a2r_r + ia);
  return r0 + (__ldg(d.a2r_r + ia + 1) - r0) * (a - __ldg(d.a2r_a + ia)) * (CLR_NA - 1.);
}

// common.h:414-431
__device__ __forceinline__ double clr_bias_model(int model, double dd, double b)
{
  if (dd <= -1) return 0;
  if (model == 2) {
    if (dd < 0) return exp(b * dd / (1 + dd));
    return 1 + b * dd;
  } else if (model == 3) {
    double v = 1 + b * dd;
    return v > 0 ? v : 0;
  }
  return pow(1 + dd, b);
}

// Philox4x32-10 (Salmon et al. 2011), counter {index_lo, index_hi, block, stream}, key {seed, 0}.
// Word j of substream (seed, stream, index) = word j%4 of block j/4. Same definition as
// third_party/shim/gsl_shim.c:shim_philox_seek.
__device__ __forceinline__ void clr_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct ClrStream {   // sequential reader of one counter-based substream
  uint32_t k0, k1, i0, i1, stream, pos, buf[4];
  uint32_t first;            // per-cell Poisson streams: draw 0 comes from a block shared by 4 cells
  bool first_pending;
  __device__ __forceinline__ ClrStream(uint32_t seed, uint32_t strm, unsigned long long index)
      : k0(seed), k1(0), i0((uint32_t)index), i1((uint32_t)(index >> 32)), stream(strm), pos(0), first(0),
        first_pending(false) {}
  // third_party/shim/gsl_shim.c:shim_philox_seek_cell: draw 0 = `w0`, draws j>=1 = words j-1 of the substream
  __device__ __forceinline__ void set_first(uint32_t w0) { first = w0; first_pending = true; }
  // position the reader at word `p` of the substream
  __device__ __forceinline__ void seek(uint32_t p)
  {
    pos = p;
    if (pos & 3) clr_philox(i0, i1, pos >> 2, stream, k0, k1, buf);
  }
  __device__ __forceinline__ uint32_t next_u32()
  {
    if (first_pending) { first_pending = false; return first; }
    if ((pos & 3) == 0) clr_philox(i0, i1, pos >> 2, stream, k0, k1, buf);
    uint32_t w = (pos & 3) == 0 ? buf[0] : (pos & 3) == 1 ? buf[1] : (pos & 3) == 2 ? buf[2] : buf[3];
    pos++;
    return w;
  }
  // gsl_rng_uniform of mt19937: u32 / 2^32 (exact in double)
  __device__ __forceinline__ double next() { return next_u32() * (1.0 / 4294967296.0); }
  __device__ __forceinline__ double next_pos() { double x; do { x = next(); } while (x == 0); return x; }
};

// ---- HEALPix (Gorski et al. 2005) in the (x,y,face) formulation of the HEALPix C library -------
__device__ __forceinline__ double clr_fmodulo(double v1, double v2)
{
  if (v1 >= 0) return (v1 < v2) ? v1 : fmod(v1, v2);
  double tmp = fmod(v1, v2) + v2;
  return (tmp == v2) ? 0. : tmp;
}
__device__ __forceinline__ int clr_imodulo(int v1, int v2) { int v = v1 % v2; return (v >= 0) ? v : v + v2; }

// ang2pix_ring_z_phi == he_ang2pix (healpix_extra.c:141-172)
__device__ __forceinline__ long long clr_ang2pix_ring_zphi(int nside, double z, double phi)
{
  const double twopi = 6.283185307179586476925286766559005768394;
  const double twothird = 2.0 / 3.0;
  const double inv_halfpi = 0.6366197723675813430755350534900574;
  double za = fabs(z);
  double tt = clr_fmodulo(phi, twopi) * inv_halfpi;
  if (za <= twothird) {
    double temp1 = nside * (0.5 + tt);
    double temp2 = nside * z * 0.75;
    int jp = (int)(temp1 - temp2);
    int jm = (int)(temp1 + temp2);
    int ir = nside + 1 + jp - jm;
    int kshift = 1 - (ir & 1);
    int ip = (jp + jm - nside + kshift + 1) / 2;
    ip = clr_imodulo(ip, 4 * nside);
    return (long long)nside * (nside - 1) * 2 + (long long)(ir - 1) * 4 * nside + ip;
  } else {
    double tp = tt - (int)(tt);
    double tmp = nside * sqrt(3 * (1 - za));
    int jp = (int)(tp * tmp);
    int jm = (int)((1.0 - tp) * tmp);
    int ir = jp + jm + 1;
    int ip = (int)(tt * ir);
    ip = clr_imodulo(ip, 4 * ir);
    if (z > 0) return 2LL * ir * (ir - 1) + ip;
    return 12LL * nside * nside - 2LL * ir * (ir + 1) + ip;
  }
}

__device__ __forceinline__ int clr_isqrt(int v) { return (int)(sqrt(v + 0.5)); }
__device__ __forceinline__ int clr_spread_bits(int v)
{
  unsigned int x = (unsigned int)v & 0xffff;
  x = (x | (x << 8)) & 0x00ff00ff;
  x = (x | (x << 4)) & 0x0f0f0f0f;
  x = (x | (x << 2)) & 0x33333333;
  x = (x | (x << 1)) & 0x55555555;
  return (int)x;
}
// ring2nest for nside <= 8192 (pixel ids fit in 31 bits up to nside 8192: 12*2^26 < 2^31)
__device__ __forceinline__ int clr_ring2nest(int nside, int pix)
{
  const int jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
  const int jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
  int iring, iphi, kshift, nr, face;
  int ncap = 2 * nside * (nside - 1), npix = 12 * nside * nside, nl2 = 2 * nside;
  if (pix < ncap) {
    iring = (1 + clr_isqrt(1 + 2 * pix)) >> 1;
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0; nr = iring;
    face = (iphi - 1) / nr;
  } else if (pix < (npix - ncap)) {
    int ip = pix - ncap;
    iring = ip / (4 * nside) + nside;
    iphi = ip % (4 * nside) + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    int ire = iring - nside + 1;
    int irm = nl2 + 2 - ire;
    int ifm = (iphi - ire / 2 + nside - 1) / nside;
    int ifp = (iphi - irm / 2 + nside - 1) / nside;
    if (ifp == ifm) face = (ifp == 4) ? 4 : ifp + 4;
    else if (ifp < ifm) face = ifp;
    else face = ifm + 8;
  } else {
    int ip = npix - pix;
    iring = (1 + clr_isqrt(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0; nr = iring;
    iring = 2 * nl2 - iring;
    face = 8 + (iphi - 1) / nr;
  }
  int irt = iring - jrll[face] * nside + 1;
  int ipt = 2 * iphi - jpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  int ix = (ipt - irt) >> 1;
  int iy = (-(ipt + irt)) >> 1;
  return face * nside * nside + clr_spread_bits(ix) + (clr_spread_bits(iy) << 1);
}

// srcs.c:68-85
__device__ __forceinline__ void clr_cart2sph(double x, double y, double z, double *r, double *cth, double *phi)
{
  *r = sqrt(x * x + y * y + z * z);
  if ((*r) == 0) { *cth = 1; *phi = 0; }
  else {
    double xn = x / (*r), yn = y / (*r);
    *cth = z / (*r);
    *phi = atan2(yn, xn);
    if ((*phi) < 0) (*phi) += 2 * 3.14159265358979323846;
  }
}

// Conversion-free floor for 0 <= t < 2^23: adding 2^23 with round-toward-zero leaves floor(t) in the mantissa.
// F2I / I2F / FRND all issue on the quarter-rate XU pipe shared with rsqrt / ex2 / rcp; these forms stay on
// the FMA and ALU pipes. m = clr_floor_magic(t): index = clr_magic_int(m), (float)index = m - 2^23.
__device__ __forceinline__ float clr_floor_magic(float t) { return __fadd_rz(t, 8388608.f); }
__device__ __forceinline__ int clr_magic_int(float m) { return __float_as_int(m) & 0x7fffff; }
// r = sqrt(r2) through one MUFU.RSQ (rsqrtf() adds a denormal-scaling sequence); exact 0 at r2 = 0
__device__ __forceinline__ float clr_sqrt_fast(float r2)
{
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaxf(r2, 1e-30f)));
  return r2 * y;
}

// single-MUFU forms (the libdevice __expf / __fdividef wrap the MUFU in denormal-range scaling sequences)
__device__ __forceinline__ float clr_ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float clr_rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ double clr_warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif // __CUDACC__
