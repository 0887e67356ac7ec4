// C ABI of colore_b200 (include/colore_b200.h): context lifetime, host<->device grid transfer,
// orchestration of the stage kernels, and the host-side tail of compute_density_normalization.
#include "clr_internal.cuh"
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include <algorithm>

static thread_local char g_err[1024] = "";

void clr_set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {
__global__ void copy_small_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int n_words)
{
  for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
}
}  // namespace

int clr_read_small(clr_ctx *c, void *host_dst, const void *dev_src, size_t bytes)
{
  CLR_CHECK(bytes % 4 == 0, "clr_read_small: %zu bytes", bytes);
  for (size_t off = 0; off < bytes; off += CLR_SMALL_BYTES) {
    size_t nb = bytes - off < CLR_SMALL_BYTES ? bytes - off : CLR_SMALL_BYTES;
    copy_small_kernel<<<1, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(static_cast<const char *>(dev_src) + off),
                                                 static_cast<uint32_t *>(c->d_small), (int)(nb / 4));
    CLR_CUDA(cudaGetLastError());
    c->launches++;
    CLR_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(static_cast<char *>(host_dst) + off, c->h_small, nb);
  }
  return 0;
}

// The potential may still be in flight on the second stream (clr_create_cartesian_fields, multi-GPU): make the main
// stream wait for it and exchange its z halo (fourier.c:401-414). Called by everything that reads or writes the grid.
int clr_npot_ready(clr_ctx *c)
{
  if (!c->npot_pending) return 0;
  c->npot_pending = false;
  CLR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_npot, 0));
  return clr_halo_update(c);
}

extern "C" {

int clr_version(void) { return 100; }
const char *clr_last_error(void) { return g_err; }

int clr_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static void copy_tab(std::vector<double> &dst, const double *src, size_t n) { dst.assign(src, src + n); }

static int create_impl(clr_ctx *c, const clr_params *p, int device);

int clr_create(const clr_params *p, int device, clr_ctx **out)
{
  CLR_CHECK(p && out, "clr_create: null argument");
  *out = nullptr;
  CLR_CHECK(clr_device_count() > device, "clr_create: CUDA device %d not available (no CPU fallback exists)", device);
  // the FFT: power-of-two Stockham plans (clr_fft.cu) or the mixed-radix path (clr_fft_generic.cu): fail here, not at
  // the first transform
  CLR_CHECK(p->n_grid >= 16 && p->n_grid <= 4096 && ((p->n_grid & (p->n_grid - 1)) == 0 || clr_fft_generic_ok(p->n_grid)),
            "n_grid=%d unsupported: the GPU FFT takes multiples of 4 in [16,4096] without prime factors above 31", p->n_grid);
  CLR_CHECK(p->nz_here > 0 && p->iz0_here >= 0 && p->iz0_here + p->nz_here <= p->n_grid, "bad slab bounds");
  CLR_CHECK(p->r_arr_r2z && p->z_arr_r2z && p->growth_d_arr && p->growth_v_arr && p->pkarr && p->logkarr,
            "clr_create: missing tables");
  CLR_CUDA(cudaSetDevice(device));
  clr_ctx *c = new clr_ctx();
  if (create_impl(c, p, device)) {
    clr_destroy(c);                 // frees whatever the failed construction had allocated
    return 1;
  }
  *out = c;
  return 0;
}

static int create_impl(clr_ctx *c, const clr_params *p, int device)
{
  c->device = device;
  c->p = *p;
  cudaDeviceProp prop;
  CLR_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CLR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CLR_CUDA(cudaEventCreate(&c->ev0)); CLR_CUDA(cudaEventCreate(&c->ev1));
  CLR_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CLR_CUDA(cudaEventCreateWithFlags(&c->ev_srcs_ready, cudaEventDisableTiming));
  CLR_CUDA(cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming));
  for (int b = 0; b < 2; b++) CLR_CUDA(cudaEventCreateWithFlags(&c->ev_buf_free[b], cudaEventDisableTiming));
  CLR_CUDA(cudaHostAlloc(&c->h_small, CLR_SMALL_BYTES, cudaHostAllocMapped));
  CLR_CUDA(cudaHostGetDevicePointer(&c->d_small, c->h_small, 0));
  CLR_CUDA(cudaEventCreate(&c->evp0)); CLR_CUDA(cudaEventCreate(&c->evp1));
  // host copies of the tables
  copy_tab(c->h_logk, p->logkarr, p->numk); copy_tab(c->h_pk, p->pkarr, p->numk);
  copy_tab(c->h_r, p->r_arr_r2z, CLR_NA); copy_tab(c->h_z, p->z_arr_r2z, CLR_NA);
  copy_tab(c->h_d1, p->growth_d_arr, CLR_NA);
  copy_tab(c->h_d2, p->growth_d2_arr ? p->growth_d2_arr : p->growth_d_arr, CLR_NA);
  copy_tab(c->h_v1, p->growth_v_arr, CLR_NA);
  copy_tab(c->h_pd, p->growth_pd_arr ? p->growth_pd_arr : p->growth_d_arr, CLR_NA);
  copy_tab(c->h_ih, p->ihub_arr ? p->ihub_arr : p->growth_d_arr, CLR_NA);
  copy_tab(c->h_a2r_a, p->a_arr_a2r ? p->a_arr_a2r : p->r_arr_r2z, CLR_NA);
  copy_tab(c->h_a2r_r, p->r_arr_a2r ? p->r_arr_a2r : p->r_arr_r2z, CLR_NA);
  // device tables
  CLR_CUDA(cudaMalloc(&c->d_tables, 9 * CLR_NA * sizeof(double)));
  const std::vector<double> *tabs[9] = {&c->h_r, &c->h_z, &c->h_d1, &c->h_d2, &c->h_v1, &c->h_pd, &c->h_ih, &c->h_a2r_a, &c->h_a2r_r};
  for (int i = 0; i < 9; i++)
    CLR_CUDA(cudaMemcpy(c->d_tables + (size_t)i * CLR_NA, tabs[i]->data(), CLR_NA * sizeof(double), cudaMemcpyHostToDevice));
  CLR_CUDA(cudaMalloc(&c->d_pk, 2 * (size_t)p->numk * sizeof(double)));
  CLR_CUDA(cudaMemcpy(c->d_pk, c->h_logk.data(), p->numk * sizeof(double), cudaMemcpyHostToDevice));
  CLR_CUDA(cudaMemcpy(c->d_pk + p->numk, c->h_pk.data(), p->numk * sizeof(double), cudaMemcpyHostToDevice));
  // FFT master twiddles exp(+2 pi i k / n), evaluated in double
  {
    int n = p->n_grid;
    std::vector<float2> w(n);
    for (int k = 0; k < n; k++) {
      double a = 2.0 * M_PI * k / n;
      w[k] = make_float2((float)cos(a), (float)sin(a));
    }
    CLR_CUDA(cudaMalloc(&c->d_twiddle, n * sizeof(float2)));
    CLR_CUDA(cudaMemcpy(c->d_twiddle, w.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
  }
  // grids: allocate_fftw (fourier.c:211-238): dens slab, npot slab + 2 halo planes
  ClrDev &d = c->dev;
  d.n = p->n_grid; d.nc = p->n_grid / 2 + 1; d.nz_here = p->nz_here; d.iz0_here = p->iz0_here;
  d.ncp = (d.nc + 7) & ~7;
  d.pitch = 2 * d.ncp;
  d.nyl = d.n; d.ky0 = 0;
  d.log2n = -1;
  for (int b = 0; b < 31; b++) if ((1 << b) == d.n) d.log2n = b;
  d.bias_model = p->bias_model; d.nside_base = p->nside_base;
  d.l_box = p->l_box;
  for (int i = 0; i < 3; i++) d.pos_obs[i] = p->pos_obs[i];
  d.glob_idr = p->glob_idr; d.r_tab_max = c->h_r[CLR_NA - 1];
  d.fgrowth_0 = p->fgrowth_0; d.hubble_0 = p->hubble_0; d.OmegaM = p->OmegaM; d.r_max = p->r_max;
  d.r_arr = c->d_tables; d.z_arr = c->d_tables + CLR_NA; d.d1_arr = c->d_tables + 2 * CLR_NA;
  d.d2_arr = c->d_tables + 3 * CLR_NA; d.v1_arr = c->d_tables + 4 * CLR_NA; d.pd_arr = c->d_tables + 5 * CLR_NA;
  d.ih_arr = c->d_tables + 6 * CLR_NA; d.a2r_a = c->d_tables + 7 * CLR_NA; d.a2r_r = c->d_tables + 8 * CLR_NA;
  {
    std::vector<float> tf(4 * CLR_NA);
    for (int i = 0; i < CLR_NA; i++) {
      tf[i] = (float)c->h_z[i]; tf[CLR_NA + i] = (float)c->h_d1[i];
      tf[2 * CLR_NA + 2 * i] = (float)c->h_d1[i];
      tf[2 * CLR_NA + 2 * i + 1] = (float)(c->h_d1[i + 1 < CLR_NA ? i + 1 : i] - c->h_d1[i]);
    }
    CLR_CUDA(cudaMalloc(&c->d_tables_f, 4 * CLR_NA * sizeof(float)));
    CLR_CUDA(cudaMemcpy(c->d_tables_f, tf.data(), 4 * CLR_NA * sizeof(float), cudaMemcpyHostToDevice));
    d.z_f = c->d_tables_f; d.d1_f = c->d_tables_f + CLR_NA;
    d.d1_t = reinterpret_cast<const float2 *>(c->d_tables_f + 2 * CLR_NA);
  }
  {
    // coordinate tables, evaluated with the reference's expressions (see ClrDev)
    const int n = d.n;
    std::vector<float> cf(3 * (size_t)n);
    std::vector<double> cd(3 * (size_t)n);
    const float dxf = p->l_box / n;                 // flouble dx (density.c:1079)
    const double dxd = p->l_box / n;                // double dx from the float division (srcs.c:147)
    for (int ax = 0; ax < 3; ax++)
      for (int i = 0; i < n; i++) {
        cf[(size_t)ax * n + i] = (float)((i + 0.0) * dxf - p->pos_obs[ax]);
        cd[(size_t)ax * n + i] = (i + 0.0) * dxd - p->pos_obs[ax];
      }
    CLR_CUDA(cudaMalloc(&c->d_coord_f, cf.size() * sizeof(float)));
    CLR_CUDA(cudaMalloc(&c->d_coord_d, cd.size() * sizeof(double)));
    CLR_CUDA(cudaMemcpy(c->d_coord_f, cf.data(), cf.size() * sizeof(float), cudaMemcpyHostToDevice));
    CLR_CUDA(cudaMemcpy(c->d_coord_d, cd.data(), cd.size() * sizeof(double), cudaMemcpyHostToDevice));
    for (int ax = 0; ax < 3; ax++) { d.cf[ax] = c->d_coord_f + (size_t)ax * n; d.cd[ax] = c->d_coord_d + (size_t)ax * n; }
  }
  size_t plane = (size_t)d.pitch * d.n;
  CLR_CUDA(cudaMalloc(&c->d_dens, plane * d.nz_here * sizeof(float)));
  // potential: slab + 4 halo planes stored behind it: [nz] = plane iz0-1, [nz+1] = plane iz0+nz (slice_left / slice_right,
  // fourier.c:401-414), [nz+2] = plane iz0-2, [nz+3] = plane iz0+nz+1 (second ring, needed by the CIC velocity stencil of
  // sources that sit in the first / last plane of a slab, srcs.c:486-504)
  CLR_CUDA(cudaMalloc(&c->d_npot, plane * (d.nz_here + 4) * sizeof(float)));
  // the padding columns are never written by the kernels: keep them finite
  CLR_CUDA(cudaMemsetAsync(c->d_dens, 0, plane * d.nz_here * sizeof(float), c->stream));
  CLR_CUDA(cudaMemsetAsync(c->d_npot, 0, plane * (d.nz_here + 4) * sizeof(float), c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  d.slice_left = c->d_npot + plane * d.nz_here;
  d.slice_right = c->d_npot + plane * (d.nz_here + 1);
  return 0;
}

static void free_pop(clr_ctx::Pop &P)
{
  cudaFree(P.d_a); cudaFree(P.d_b); cudaFree(P.d_norm); cudaFree(P.d_counts); cudaFree(P.d_bound);
  cudaFree(P.d_pos); cudaFree(P.d_ipix); cudaFree(P.d_srcs); cudaFree(P.d_srcs_alt); cudaFree(P.d_sup_entries);
  cudaFree(P.d_skw_dg); cudaFree(P.d_skw_v);
}

int clr_destroy(clr_ctx *c)
{
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  if (c->stream) cudaStreamSynchronize(c->stream);
  clr_comm_destroy(c);
  for (int i = 0; i < CLR_NPOP_MAX; i++) { free_pop(c->srcs[i]); free_pop(c->imap[i]); free_pop(c->cstm[i]); }
  cudaFree(c->d_dens); cudaFree(c->d_npot); cudaFree(c->d_tables); cudaFree(c->d_tables_f); cudaFree(c->d_pk);
  cudaFree(c->d_lens_data); cudaFree(c->d_los_hess);
  cudaFree(c->d_coord_f); cudaFree(c->d_coord_d); cudaFree(c->d_fft_tmp); cudaFree(c->d_fft_sync); cudaFree(c->d_hist);
  for (int i = 0; i < 3; i++) cudaFree(c->d_lpt_pos[i]);
  cudaFree(c->d_twiddle); cudaFree(c->d_scratch); cudaFree(c->d_pkt); cudaFree(c->d_sincos);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->evp0); cudaEventDestroy(c->evp1);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_srcs_ready); cudaEventDestroy(c->ev_copy_done);
  for (int b = 0; b < 2; b++) cudaEventDestroy(c->ev_buf_free[b]);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  if (c->h_small) cudaFreeHost(c->h_small);
  delete c;
  return 0;
}

int clr_synchronize(clr_ctx *c)
{
  if (clr_npot_ready(c)) return 1;
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->copy_stream));
  c->copy_pending = false;
  c->buf_busy[0] = c->buf_busy[1] = false;
  return 0;
}
long long clr_launch_count(clr_ctx *c) { return c->launches; }


static int set_pop(clr_ctx *c, clr_ctx::Pop &P, const double *a, const double *b)
{
  P.h_a.assign(a, a + CLR_NA);
  P.h_b.assign(b, b + CLR_NA);
  if (!P.d_a) CLR_CUDA(cudaMalloc(&P.d_a, CLR_NA * sizeof(double)));
  if (!P.d_b) CLR_CUDA(cudaMalloc(&P.d_b, CLR_NA * sizeof(double)));
  if (!P.d_norm) CLR_CUDA(cudaMalloc(&P.d_norm, CLR_NA * sizeof(double)));
  CLR_CUDA(cudaMemcpyAsync(P.d_a, a, CLR_NA * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(P.d_b, b, CLR_NA * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  P.set = true;
  P.have_norm = false;
  c->hist_valid = false;
  return 0;
}

int clr_set_srcs(clr_ctx *c, int ipop, const double *nz_arr, const double *bz_arr)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX, "population index %d out of range", ipop);
  return set_pop(c, c->srcs[ipop], nz_arr, bz_arr);
}

int clr_set_imap(clr_ctx *c, int ipop, const double *tz_arr, const double *bz_arr, int nside, int nr,
                 const float *r0, const float *rf)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX, "population index %d out of range", ipop);
  CLR_CHECK(nside > 0 && (nside & (nside - 1)) == 0 && nr > 0, "imap: bad nside/nr");
  clr_ctx::Pop &P = c->imap[ipop];
  if (set_pop(c, P, tz_arr, bz_arr)) return 1;
  P.nside = nside; P.nr = nr;
  // imap_preproc (imap.c:105-121): shells sorted by r0
  std::vector<int> order(nr);
  for (int i = 0; i < nr; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return r0[a] < r0[b]; });
  P.r0.resize(nr); P.rf.resize(nr);
  for (int i = 0; i < nr; i++) { P.r0[i] = r0[order[i]]; P.rf[i] = rf[order[i]]; }
  return 0;
}

int clr_set_cstm(clr_ctx *c, int ipop, const double *kz_arr, const double *bz_arr)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX, "population index %d out of range", ipop);
  return set_pop(c, c->cstm[ipop], kz_arr, bz_arr);
}

static float *grid_ptr(clr_ctx *c, int which) { return which == CLR_GRID_DENS ? c->d_dens : c->d_npot; }

// host side = the reference layout, rows of 2*(n/2+1) floats (fourier.c:46-51); device side = rows of dev.pitch floats
int clr_grid_put(clr_ctx *c, int which, const float *host)
{
  if (which == CLR_GRID_DENS) c->hist_valid = false;
  if (which == CLR_GRID_NPOT && clr_npot_ready(c)) return 1;
  const size_t hrow = (size_t)2 * c->dev.nc * sizeof(float), drow = (size_t)c->dev.pitch * sizeof(float);
  CLR_CUDA(cudaMemcpy2DAsync(grid_ptr(c, which), drow, host, hrow, hrow, (size_t)c->dev.n * c->dev.nz_here,
                             cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int clr_grid_get(clr_ctx *c, int which, float *host)
{
  if (which == CLR_GRID_NPOT && clr_npot_ready(c)) return 1;
  const size_t hrow = (size_t)2 * c->dev.nc * sizeof(float), drow = (size_t)c->dev.pitch * sizeof(float);
  CLR_CUDA(cudaMemcpy2DAsync(host, hrow, grid_ptr(c, which), drow, hrow, (size_t)c->dev.n * c->dev.nz_here,
                             cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int clr_grid_device_ptr(clr_ctx *c, int which, void **dptr)
{
  c->hist_valid = false;            // the caller may write the grid behind our back
  if (which == CLR_GRID_NPOT && clr_npot_ready(c)) return 1;
  *dptr = grid_ptr(c, which);
  return 0;
}
int clr_grid_pitch(clr_ctx *c, long long *pitch_floats)
{
  *pitch_floats = c->dev.pitch;
  return 0;
}

int clr_fill_modes(clr_ctx *c, uint32_t seed) { if (clr_npot_ready(c)) return 1; return clr_fields_fill(c, seed); }
int clr_fft_c2r(clr_ctx *c, int which)
{
  c->hist_valid = false;
  if (clr_npot_ready(c)) return 1;
  return clr_fft_c2r_impl(c, grid_ptr(c, which), 1.0, nullptr);
}
int clr_fft_r2c(clr_ctx *c, int which)
{
  c->hist_valid = false;
  if (clr_npot_ready(c)) return 1;
  return clr_fft_r2c_impl(c, grid_ptr(c, which));
}
int clr_update_halo(clr_ctx *c) { if (clr_npot_ready(c)) return 1; return clr_halo_update(c); }

static void finish_moments(clr_ctx *c, const double mom[2], double *out2)
{
  // compute_sigma_dens (fourier.c:61-75)
  double ng_tot = (double)c->dev.n * c->dev.n * c->dev.n;
  double mean = mom[0] / ng_tot, s2 = mom[1] / ng_tot;
  c->mean_gauss = mean;
  c->sigma2_gauss = s2 - mean * mean;
  if (out2) { out2[0] = mean; out2[1] = c->sigma2_gauss; }
}

int clr_normalize_fields(clr_ctx *c, double *out2)
{
  c->hist_valid = false;
  if (clr_npot_ready(c)) return 1;
  double mom[2];
  if (clr_fields_scale_moments(c, mom)) return 1;
  if (clr_halo_update(c)) return 1;
  finish_moments(c, mom, out2);
  return 0;
}

int clr_create_cartesian_fields(clr_ctx *c, uint32_t seed, int inject, double *out2)
{
  c->hist_valid = false;
  if (clr_npot_ready(c)) return 1;
  if (clr_ensure_scratch(c, 4096)) return 1;
  CLR_CUDA(cudaMemsetAsync(c->d_scratch, 0, 2 * sizeof(double), c->stream));
  double norm = pow(sqrt(2 * M_PI) / c->p.l_box, 3);      // fourier.c:389
  bool ran = false;
  // own stream on one GPU: mode fill fused into the z pass of both transforms (clr_fft.cu)
  if (!inject && clr_fft_fill_c2r(c, seed, norm, c->d_scratch, &ran)) return 1;
  if (!ran) {
    if (!inject && clr_fields_fill(c, seed)) return 1;
    if (c->nranks > 1 && c->stream2 && c->p2p && c->p2p_enabled && c->fft_overlap) {
      // Two pipelines: the density is transformed on the main stream; once its z pass has been exchanged (NVLink is
      // free again) the potential follows on the second stream with its own staging buffer, under the y / x passes of
      // the density and whatever the caller queues next (lognormal, normalisation, Poisson). The z halo of the
      // potential is exchanged by clr_npot_ready when the first consumer shows up.
      c->ev_after_z = c->ev_z_done;
      int bad = clr_fft_c2r_impl(c, c->d_dens, norm, c->d_scratch);
      c->ev_after_z = nullptr;
      if (bad) return 1;
      cudaStream_t main_stream = c->stream;
      CLR_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_z_done, 0));
      c->stream = c->stream2;
      clr_use_set(c, 1);
      bad = clr_fft_c2r_impl(c, c->d_npot, norm, nullptr);
      if (!bad && cudaEventRecord(c->ev_npot, c->stream) != cudaSuccess) bad = 1;
      c->stream = main_stream;
      clr_use_set(c, 0);
      if (bad) return 1;
      c->npot_pending = true;
    } else {
      if (clr_fft_c2r_impl(c, c->d_dens, norm, c->d_scratch)) return 1;   // scaling + moments fused in the x pass
      if (clr_fft_c2r_impl(c, c->d_npot, norm, nullptr)) return 1;
    }
  }
  if (!c->npot_pending && clr_halo_update(c)) return 1;
  if (clr_comm_allreduce_f64(c, c->d_scratch, 2)) return 1;            // fourier.c:69-70
  double mom[2];
  if (clr_read_small(c, mom, c->d_scratch, sizeof(mom))) return 1;
  finish_moments(c, mom, out2);
  return 0;
}

int clr_set_sigma2_gauss(clr_ctx *c, double s2) { c->sigma2_gauss = s2; return 0; }

int clr_set_option(clr_ctx *c, const char *name, int value)
{
  if (!strcmp(name, "exact_math")) { c->exact_math = value; c->hist_valid = false; return 0; }
  if (!strcmp(name, "hist_fused")) { c->hist_fused = value; c->hist_valid = false; return 0; }
  if (!strcmp(name, "lpt_interp_type")) { c->lpt_interp_type = value; return 0; }
  if (!strcmp(name, "keep_particles")) { c->keep_particles = value; return 0; }
  if (!strcmp(name, "async_results")) { c->async_results = value; return 0; }
  if (!strcmp(name, "srcs_compact")) { c->srcs_compact = value; return 0; }
  if (!strcmp(name, "fft_fused")) { c->fft_fused = value; return 0; }
  if (!strcmp(name, "los_precompute")) {
    c->los_precompute = value;
    if (!value) { cudaFree(c->d_los_hess); c->d_los_hess = nullptr; c->los_hess_bytes = 0; }
    return 0;
  }
  if (!strcmp(name, "fill_fused")) { c->fill_fused = value; return 0; }
  if (!strcmp(name, "fill_w")) { c->fill_w = value; return 0; }
  if (!strcmp(name, "fill_cluster")) { c->fill_cluster = value; return 0; }
  if (!strcmp(name, "p2p_fused")) { c->p2p_enabled = value; return 0; }
  if (!strcmp(name, "fft_overlap")) { if (clr_npot_ready(c)) return 1; c->fft_overlap = value; return 0; }
  if (!strcmp(name, "p2p_tiled")) { c->p2p_tiled = value; return 0; }
  clr_set_error("unknown option %s", name);
  return 1;
}

int clr_compute_physical_density_field(clr_ctx *c)
{
  c->hist_valid = false;
  if (c->p.dens_type == CLR_DENS_TYPE_LGNR) return clr_fields_lognormal(c, 0);
  if (c->p.dens_type == CLR_DENS_TYPE_CLIP) return clr_fields_lognormal(c, 1);
  if (clr_npot_ready(c)) return 1;          // LPT borrows the transform machinery and staging buffer
  if (c->p.dens_type == CLR_DENS_TYPE_1LPT) return clr_lpt_run(c, 1);
  if (c->p.dens_type == CLR_DENS_TYPE_2LPT) return clr_lpt_run(c, 2);
  clr_set_error("Density type %d not supported\n", c->p.dens_type);     // density.c:1119
  return 1;
}

// cosmo.c:30-38 on the host copy of the tables
static double host_lerp(const clr_ctx *c, double r, const std::vector<double> &f, double f0, double ff)
{
  if (r <= 0) return f0;
  else if (r >= c->h_r[CLR_NA - 1]) return ff;
  int ir = (int)(r * c->p.glob_idr);
  return f[ir] + (f[ir + 1] - f[ir]) * (r - c->h_r[ir]) * c->p.glob_idr;
}
static double host_bg_z(const clr_ctx *c, double r) { return host_lerp(c, r, c->h_z, 0, c->h_z[CLR_NA - 1]); }

// gsl linear spline evaluation (interval by bisection), used at density.c:1304-1352
static double lin_interp(const std::vector<double> &x, const std::vector<double> &y, double xv)
{
  size_t lo = 0, hi = x.size() - 1;
  while (hi - lo > 1) {
    size_t mid = (hi + lo) >> 1;
    if (x[mid] > xv) hi = mid; else lo = mid;
  }
  double h = x[hi] - x[lo];
  double A = (x[hi] - xv) / h, B = (xv - x[lo]) / h;
  return A * y[lo] + B * y[hi];
}

int clr_compute_density_normalization(clr_ctx *c)
{
  // density.c:1233-1245
  double zmax = host_bg_z(c, (double)(c->p.l_box * 0.5));
  int nz = (int)(zmax / 0.05) + 2;
  double idz = (nz - 2) / zmax;
  std::vector<clr_ctx::Pop *> pops;
  for (int i = 0; i < CLR_NPOP_MAX; i++) if (c->srcs[i].set) pops.push_back(&c->srcs[i]);
  for (int i = 0; i < CLR_NPOP_MAX; i++) if (c->imap[i].set) pops.push_back(&c->imap[i]);
  for (int i = 0; i < CLR_NPOP_MAX; i++) if (c->cstm[i].set) pops.push_back(&c->cstm[i]);     // density.c:1177-1178
  int npop = (int)pops.size();
  std::vector<const double *> d_bz(npop ? npop : 1);
  for (int i = 0; i < npop; i++) d_bz[i] = pops[i]->d_b;
  std::vector<unsigned long long> narr(nz);
  std::vector<double> zarr(nz), barr((size_t)(npop ? npop : 1) * nz);
  if (clr_fields_norm_hist(c, npop, d_bz.data(), nz, idz, narr.data(), zarr.data(), barr.data())) return 1;
  // (multi-GPU: the histograms are all-reduced here, density.c:1262-1269)
  // density.c:1272-1297
  for (int iz = 0; iz < nz; iz++) {
    if (narr[iz] > 0) {
      zarr[iz] /= narr[iz];
      for (int ip = 0; ip < npop; ip++) barr[(size_t)ip * nz + iz] = narr[iz] / barr[(size_t)ip * nz + iz];
    }
  }
  zarr[0] = 0;
  zarr[nz - 1] = host_bg_z(c, 0.5 * c->p.l_box);
  c->z0_norm = zarr[0];
  c->zf_norm = zarr[nz - 1];
  for (int ip = 0; ip < npop; ip++) {
    double *b = &barr[(size_t)ip * nz];
    b[0] = b[1];
    b[nz - 1] = b[nz - 2];
    clr_ctx::Pop &P = *pops[ip];
    P.norm_0 = b[0];
    P.norm_f = b[nz - 1];
    std::vector<double> yv(b, b + nz);
    P.h_norm.resize(CLR_NA);
    // density.c:1323-1355
    for (int ii = 0; ii < CLR_NA; ii++) {
      double z = host_bg_z(c, c->h_r[ii]);
      double nm;
      if (z < c->z0_norm) nm = P.norm_0;
      else if (z >= c->zf_norm) nm = P.norm_f;
      else nm = lin_interp(zarr, yv, z);
      P.h_norm[ii] = nm;
    }
    CLR_CUDA(cudaMemcpyAsync(P.d_norm, P.h_norm.data(), CLR_NA * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    P.have_norm = true;
  }
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static clr_ctx::Pop *pick_pop(clr_ctx *c, int kind, int ipop)
{
  if (ipop < 0 || ipop >= CLR_NPOP_MAX) return nullptr;
  return kind == 0 ? &c->srcs[ipop] : kind == 1 ? &c->imap[ipop] : &c->cstm[ipop];
}

int clr_get_norm(clr_ctx *c, int kind, int ipop, double *norm_arr, double *ends2, double *zends2)
{
  clr_ctx::Pop *P = pick_pop(c, kind, ipop);
  CLR_CHECK(P && P->have_norm, "no normalisation for population %d/%d", kind, ipop);
  if (norm_arr) memcpy(norm_arr, P->h_norm.data(), CLR_NA * sizeof(double));
  if (ends2) { ends2[0] = P->norm_0; ends2[1] = P->norm_f; }
  if (zends2) { zends2[0] = c->z0_norm; zends2[1] = c->zf_norm; }
  return 0;
}

int clr_set_norm(clr_ctx *c, int kind, int ipop, const double *norm_arr, const double *ends2)
{
  clr_ctx::Pop *P = pick_pop(c, kind, ipop);
  CLR_CHECK(P && P->set, "population %d/%d not set", kind, ipop);
  P->h_norm.assign(norm_arr, norm_arr + CLR_NA);
  P->norm_0 = ends2[0]; P->norm_f = ends2[1];
  CLR_CUDA(cudaMemcpyAsync(P->d_norm, norm_arr, CLR_NA * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  P->have_norm = true;
  return 0;
}

int clr_srcs_set_cartesian(clr_ctx *c, int ipop, uint32_t seed, long long *nsrc_out)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX, "population index %d out of range", ipop);
  if (clr_srcs_run(c, ipop, seed)) return 1;
  if (clr_srcs_local(c, ipop)) return 1;
  if (nsrc_out) *nsrc_out = c->srcs[ipop].nsrc;
  return 0;
}

int clr_srcs_get_counts(clr_ctx *c, int ipop, int32_t *nsources_padded)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  CLR_CHECK(P.d_counts, "no counts for population %d", ipop);
  // device layout is unpadded [nz][n][n]; the reference array has the padded pitch (srcs.c:125)
  const ClrDev &d = c->dev;
  const int hpitch = 2 * d.nc;
  int32_t *dense = nullptr;
  bool owned = false;
  if (clr_srcs_dense_counts(c, ipop, &dense, &owned)) return 1;
  cudaError_t e1 = cudaMemcpy2DAsync(nsources_padded, (size_t)hpitch * sizeof(int32_t), dense, (size_t)d.n * sizeof(int32_t),
                                     (size_t)d.n * sizeof(int32_t), (size_t)d.n * d.nz_here, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  if (owned) cudaFree(dense);
  CLR_CHECK(e1 == cudaSuccess && e2 == cudaSuccess, "clr_srcs_get_counts: copy failed");
  for (long long row = 0; row < (long long)d.n * d.nz_here; row++)
    for (int x = d.n; x < hpitch; x++) nsources_padded[row * hpitch + x] = 0;
  return 0;
}

int clr_srcs_get_cartesian(clr_ctx *c, int ipop, float *pos4, int32_t *ipix)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  if (P.nsrc == 0) return 0;
  if (pos4) CLR_CUDA(cudaMemcpyAsync(pos4, P.d_pos, (size_t)P.nsrc * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (ipix) CLR_CUDA(cudaMemcpyAsync(ipix, P.d_ipix, (size_t)P.nsrc * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int clr_srcs_get_local_properties(clr_ctx *c, int ipop, float *srcs9)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  if (P.nsrc == 0) return 0;
  if (c->async_results) {
    // the copy runs on its own stream behind the kernels queued so far; the next run only waits for it
    // (on the device) before it reuses the catalogue buffers. `srcs9` is valid after clr_synchronize.
    CLR_CUDA(cudaEventRecord(c->ev_srcs_ready, c->stream));
    CLR_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_srcs_ready, 0));
    CLR_CUDA(cudaMemcpyAsync(srcs9, P.d_srcs, (size_t)P.nsrc * 9 * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
    CLR_CUDA(cudaEventRecord(c->ev_copy_done, c->copy_stream));
    CLR_CUDA(cudaEventRecord(c->ev_buf_free[P.srcs_buf], c->copy_stream));
    c->buf_busy[P.srcs_buf] = true;
    c->copy_pending = true;
    return 0;
  }
  CLR_CUDA(cudaMemcpyAsync(srcs9, P.d_srcs, (size_t)P.nsrc * 9 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int clr_srcs_distribute(clr_ctx *c, int ipop, int beam_first, long long *nsrc_out)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX && c->srcs[ipop].set, "population index %d out of range", ipop);
  return clr_srcs_distribute_impl(c, ipop, beam_first, nsrc_out);
}
int clr_srcs_beam_rsd(clr_ctx *c, int ipop) { if (clr_npot_ready(c)) return 1; return clr_srcs_beam(c, ipop); }
int clr_srcs_get_beam_properties(clr_ctx *c, int ipop, int has_lensing, int has_skw, int skw_gauss, int rsd_done)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX && c->srcs[ipop].set, "population index %d out of range", ipop);
  if (clr_npot_ready(c)) return 1;
  return clr_beam_srcs(c, ipop, has_lensing, has_skw, skw_gauss, rsd_done);
}
int clr_srcs_get_skewers(clr_ctx *c, int ipop, float *dg_skw, float *v_skw)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX, "population index %d out of range", ipop);
  return clr_beam_get_skewers(c, ipop, dg_skw, v_skw);
}
int clr_cstm_get_beam_properties(clr_ctx *c, int ipop, long long num_pix, const double *pos3, float *data)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX, "population index %d out of range", ipop);
  return clr_beam_cstm(c, ipop, num_pix, pos3, data);
}
int clr_lensing_get_beam_properties(clr_ctx *c, int nbeams, int nr_sh, float *r_sh, const long long *npp, const double *pos3,
                                    float *data)
{
  if (clr_npot_ready(c)) return 1;
  return clr_beam_lens_shells(c, nbeams, nr_sh, r_sh, npp, pos3, data);
}
int clr_srcs_lensing_from_shells(clr_ctx *c, int ipop, int nr_sh, const float *r_sh, const int32_t *nside_sh, int node, int nnodes,
                                 long long *n_bad)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX && c->srcs[ipop].set, "population index %d out of range", ipop);
  return clr_beam_srcs_from_shells(c, ipop, nr_sh, r_sh, nside_sh, node, nnodes, n_bad);
}
int clr_lpt_get_particles(clr_ctx *c, float *x, float *y, float *z) { return clr_lpt_particles(c, x, y, z); }
int clr_lpt_exchange_counts(clr_ctx *c, long long *sent, long long *received)
{
  if (sent) *sent = c->lpt_sent;
  if (received) *received = c->lpt_received;
  return 0;
}

int clr_imap_set_cartesian(clr_ctx *c, int ipop, float *data, int32_t *nadd)
{ if (clr_npot_ready(c)) return 1; return clr_maps_imap(c, ipop, data, nadd); }
int clr_kappa_get_beam_properties(clr_ctx *c, long long num_pix, const double *pos3, int nplanes, const float *rf, float *data)
{ if (clr_npot_ready(c)) return 1; return clr_maps_los(c, 0, num_pix, pos3, nplanes, rf, data); }
int clr_isw_get_beam_properties(clr_ctx *c, long long num_pix, const double *pos3, int nplanes, const float *rf, float *data)
{ if (clr_npot_ready(c)) return 1; return clr_maps_los(c, 1, num_pix, pos3, nplanes, rf, data); }

int clr_timer_start(clr_ctx *c) { CLR_CUDA(cudaEventRecord(c->ev0, c->stream)); return 0; }
int clr_timer_stop_ms(clr_ctx *c, float *ms)
{
  CLR_CUDA(cudaEventRecord(c->ev1, c->stream));
  CLR_CUDA(cudaEventSynchronize(c->ev1));
  CLR_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return 0;
}
static int resolve_stage_events(clr_ctx *c)
{
  if (c->ev_pending.empty()) return 0;
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  if (c->stream2) CLR_CUDA(cudaStreamSynchronize(c->stream2));
  for (auto &pe : c->ev_pending) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev_pool[pe.slot], c->ev_pool[pe.slot + 1]);
    StageTime &s = c->stage[pe.name];
    s.ms += ms; s.launches += pe.nl;
  }
  c->ev_pending.clear();
  c->ev_used = 0;
  return 0;
}
int clr_set_profiling(clr_ctx *c, int on)
{
  if (resolve_stage_events(c)) return 1;
  c->profiling = on != 0;
  if (on) c->stage.clear();
  return 0;
}
int clr_get_stage_ms(clr_ctx *c, const char *stage, float *ms, int *launches)
{
  if (resolve_stage_events(c)) return 1;
  auto it = c->stage.find(stage);
  if (it == c->stage.end()) { *ms = 0; if (launches) *launches = 0; return 0; }
  *ms = it->second.ms;
  if (launches) *launches = it->second.launches;
  return 0;
}

}  // extern "C"
