// Field kernels: Gaussian mode fill, scaling + moments, z-halo, lognormal / clip transform,
// normalisation histogram. HBM-bound streaming kernels: 8-byte vector accesses along x (the
// reference row pitch 2*(n/2+1) floats only guarantees 8-byte alignment), grid-stride loops sized
// in multiples of the SM count.
// Compiled with -fmad=false so that double-precision table lookups round exactly like the
// reference's scalar C code (no fused multiply-add contraction).
#include "clr_internal.cuh"
#include "clr_fill.cuh"
#include <algorithm>
#include <string.h>

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// create_grids_fourier (fourier.c:285-359) + rng_delta_gauss (common.c:191-201) + pk_linear0
// (cosmo.c:291-308). One thread per mode (kz_local, ky, kx<=n/2). RNG: substream
// (seed, stream 0, global mode-pair index): the neighbouring modes 2p, 2p+1 of a row share one
// Philox block, words 0,1 -> phase, modulus of the even mode, words 2,3 -> the odd mode (phase first, as
// the reference draws them). Arithmetic in double like the reference; the two complex64 stores per
// mode are the only HBM traffic (8 B per real-space cell).
template <bool EXACT>
__global__ void __launch_bounds__(kThreads)
fill_modes_kernel(const ClrDev d, float2 *__restrict__ dens_f, float2 *__restrict__ npot_f, uint32_t seed,
                  const double *__restrict__ logkarr, const double *__restrict__ pkarr, int numk, double logkmin,
                  double logkmax, double idlogk, double n_scal, double prefac_lensing, double r2_smooth,
                  int do_smoothing, int smooth_potential, double lgdk)
{
  const long long n_modes = (long long)d.nz_here * d.n * d.nc;
  const double dk = 2 * 3.14159265358979323846 / d.l_box;
  const double idk3 = 1. / (dk * dk * dk);
  // a CTA takes groups of rows (kz,ky) so that all index divisions are 32-bit and by nc / n only
  // k space is held as [kz][ky_local][kx] with ky in [ky0, ky0+nyl) (single GPU: nyl = n, ky0 = 0, which
  // is the reference layout); rows = (kz, ky_local)
  const unsigned n_rows = (unsigned)d.n * (unsigned)d.nyl;
  // rows of nc modes: the CTA takes a group of `rpb` rows; inside a group, thread t walks (row, kx) pairs
  // with kx = t mod kxs so that no per-mode integer division is needed (kxs = blockDim / rows-in-flight)
  const unsigned rpb = max(1u, 1024u / (unsigned)d.nc);
  const unsigned n_groups = (n_rows + rpb - 1) / rpb;
  (void)n_modes;
  // rows handled concurrently by the CTA: as many as fit with at least nc threads' worth of lanes each
  const unsigned rows_par = min(rpb, max(1u, (unsigned)blockDim.x / (unsigned)d.nc));
  const unsigned kxs = (unsigned)blockDim.x / rows_par;         // lanes per row
  const unsigned my_r = threadIdx.x / kxs, my_k = threadIdx.x - my_r * kxs;
  for (unsigned grp = blockIdx.x; grp < n_groups; grp += gridDim.x)
  for (unsigned rl = my_r; rl < rpb; rl += rows_par)
  for (int kk = (int)my_k; kk < d.nc; kk += (int)kxs) {
    unsigned row = grp * rpb + rl;
    if (row >= n_rows || my_r >= rows_par) continue;
    int ii_true = (int)(row / (unsigned)d.nyl);
    int jj = d.ky0 + (int)(row - (unsigned)ii_true * (unsigned)d.nyl);
    long long idx = (long long)row * d.ncp + kk;
    double kz = (2 * ii_true <= d.n) ? ii_true * dk : -(d.n - ii_true) * dk;
    double ky = (2 * jj <= d.n) ? jj * dk : -(d.n - jj) * dk;
    double kx = (2 * kk <= d.n) ? kk * dk : -(d.n - kk) * dk;
    double k_mod2 = kx * kx + ky * ky + kz * kz;
    float2 dk_out = make_float2(0.f, 0.f), pk_out = make_float2(0.f, 0.f);
    if (k_mod2 > 0 && !EXACT) {
      // fp32 transcendentals, double only where it is cheap and protects the result:
      //  * log10(k): k^2/dk^2 = m is an integer < 2^24; log2(m) = e + log2f(f) with the sum in double,
      //    so the P(k) table position keeps ~1e-8 absolute accuracy (a float sum would lose it);
      //  * ln(1-u2): log1pf for small u2, logf of the exactly computed 1-u2 otherwise;
      //  * phase 2*pi*u1 through sincospif (exact argument reduction).
      int mi = (2 * ii_true <= d.n ? ii_true : d.n - ii_true), mj = (2 * jj <= d.n ? jj : d.n - jj);
      int m = kk * kk + mj * mj + mi * mi;
      int e2 = 31 - __clz(m);
      float fm = __int2float_rn(m) * __int_as_float((127 - e2) << 23);      // m * 2^-e2 in [1,2), exact
      // lg2.approx on [1,2): absolute error 2^-22, i.e. 7e-8 in log10 k
      double lgk = lgdk + 0.15051499783199060 * ((double)e2 + (double)__log2f(fm));   // 0.5*log10(2)
      double pk;
      int ik = (int)((lgk - logkmin) * idlogk);
      if (ik < 0) pk = __ldg(pkarr) * (double)exp10f((float)(n_scal * (lgk - logkmin)));
      else if (ik < numk) {
        double p0 = __ldg(pkarr + ik), p1 = (ik + 1 < numk) ? __ldg(pkarr + ik + 1) : p0;
        pk = p0 + (lgk - __ldg(logkarr + ik)) * (p1 - p0) * idlogk;
      } else pk = __ldg(pkarr + numk - 1) * (double)exp10f((float)(-3 * (lgk - logkmax)));
      float sigma2 = (float)(pk * idk3);
      uint32_t w[4];
      // the neighbouring modes 2p, 2p+1 share one Philox block: p + npair*(jj + n*ii), npair = ceil(nc/2)
      unsigned long long gidx = (unsigned long long)(kk >> 1) + (unsigned long long)((d.nc + 1) / 2) * ((unsigned long long)jj + (unsigned long long)d.n * ii_true);
      clr_philox((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, 0u, seed, 0u, w);
      if (kk & 1) { w[0] = w[2]; w[1] = w[3]; }
      // ln(1-u2) without a branch: T = 2^32 - w is 1-u2 in units of 2^-32 (exact integer). logf of the
      // float-rounded T plus the first-order term of the (exact) rounding residue keeps full relative
      // accuracy both for u2 -> 0 (T -> 2^32, result -> 0) and u2 -> 1.
      float l1 = 0.f;
      if (w[1]) {
        uint32_t T = 0u - w[1];
        float Tf = __uint2float_rn(T);
        float resid = (float)((long long)T - (long long)Tf);
        l1 = logf(Tf * 2.3283064365386963e-10f) + __fdividef(resid, Tf);
      }
      float delta_mod = sqrtf(-sigma2 * l1);
      float sn, cs;
      sincospif((float)(w[0] >> 7) * (1.f / 16777216.f), &sn, &cs);         // 2*u1 with 25 bits
      float dre = delta_mod * cs, dim = delta_mod * sn;
      float pk2 = -(float)prefac_lensing * __frcp_rn((float)k_mod2);
      float pre = pk2 * dre, pim = pk2 * dim;
      if (do_smoothing) {
        float sm = __expf((float)(-0.5 * r2_smooth * k_mod2));
        dre *= sm; dim *= sm;
        if (smooth_potential) { pre *= sm; pim *= sm; }
      }
      dk_out = make_float2(dre, dim);
      pk_out = make_float2(pre, pim);
    } else if (k_mod2 > 0) {
      double lgk = 0.5 * log10(k_mod2);
      // pk_linear0
      double pk;
      int ik = (int)((lgk - logkmin) * idlogk);
      if (ik < 0) pk = __ldg(pkarr) * pow(10., n_scal * (lgk - logkmin));
      else if (ik < numk) {
        double p0 = __ldg(pkarr + ik), p1 = (ik + 1 < numk) ? __ldg(pkarr + ik + 1) : p0;
        pk = p0 + (lgk - __ldg(logkarr + ik)) * (p1 - p0) * idlogk;
      } else pk = __ldg(pkarr + numk - 1) * pow(10., -3 * (lgk - logkmax));
      double sigma2 = pk * idk3;
      uint32_t w[4];
      // the neighbouring modes 2p, 2p+1 share one Philox block: p + npair*(jj + n*ii), npair = ceil(nc/2)
      unsigned long long gidx = (unsigned long long)(kk >> 1) + (unsigned long long)((d.nc + 1) / 2) * ((unsigned long long)jj + (unsigned long long)d.n * ii_true);
      clr_philox((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, 0u, seed, 0u, w);
      if (kk & 1) { w[0] = w[2]; w[1] = w[3]; }
      double u1 = w[0] * (1.0 / 4294967296.0), u2 = w[1] * (1.0 / 4294967296.0);
      double delta_mod = sqrt(-sigma2 * log(1 - u2));
      double sn, cs;
      sincospi(2.0 * u1, &sn, &cs);       // phase = 2*pi*u1, exact argument reduction
      float dre = (float)(delta_mod * cs), dim = (float)(delta_mod * sn);
      // potential from the UNSMOOTHED, already float-rounded delta_k (fourier.c:346)
      double fac = -prefac_lensing;
      float pre = (float)(fac * (double)dre / k_mod2), pim = (float)(fac * (double)dim / k_mod2);
      if (do_smoothing) {
        double sm = exp(-0.5 * r2_smooth * k_mod2);
        dre = (float)((double)dre * sm); dim = (float)((double)dim * sm);
        if (smooth_potential) { pre = (float)((double)pre * sm); pim = (float)((double)pim * sm); }
      }
      dk_out = make_float2(dre, dim);
      pk_out = make_float2(pre, pim);
    }
    dens_f[idx] = dk_out;
    npot_f[idx] = pk_out;
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 mode fill (default; exact_math=0): arithmetic in clr_fill.cuh. One thread per mode PAIR (2p, 2p+1) of a row:
// one Philox block and one 16-byte store per grid (rows are 64-byte aligned, the pair never straddles the pitch).
__global__ void __launch_bounds__(kThreads)
fill_modes_fast_kernel(const ClrDev d, float2 *__restrict__ dens_f, float2 *__restrict__ npot_f, uint32_t seed,
                       const float2 *__restrict__ pkt, const float2 *__restrict__ sct, const FillFastK k, int kxs_log2)
{
  const unsigned n_rows = (unsigned)d.n * (unsigned)d.nyl;          // rows (kz, ky_local) of nc modes
  const unsigned kxs = 1u << kxs_log2, rows_par = blockDim.x >> kxs_log2;
  const unsigned my_r = threadIdx.x >> kxs_log2, my_k = threadIdx.x & (kxs - 1);
  const int npair_row = (d.nc + 1) / 2;                              // pairs per row
  for (unsigned row = blockIdx.x * rows_par + my_r; row < n_rows; row += gridDim.x * rows_par) {
    const int ii = (int)(row / (unsigned)d.nyl);
    const int jj = d.ky0 + (int)(row - (unsigned)ii * (unsigned)d.nyl);
    const int mi = (2 * ii <= d.n ? ii : d.n - ii), mj = (2 * jj <= d.n ? jj : d.n - jj);
    const int m_row = mj * mj + mi * mi;
    const long long idx0 = (long long)row * d.ncp;
    const unsigned long long g0 = (unsigned long long)npair_row * ((unsigned long long)jj + (unsigned long long)d.n * ii);
    for (int kp = (int)my_k; kp < npair_row; kp += (int)kxs) {
      const unsigned long long gidx = g0 + (unsigned)kp;
      uint32_t w[4];
      clr_philox((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, 0u, seed, 0u, w);
      float2 dk2[2], pk2[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int kk = 2 * kp + h;
        const int m = kk * kk + m_row;
        clr_fill_mode(k, pkt, sct, m, kk < d.nc && m > 0, w[2 * h], w[2 * h + 1], dk2[h], pk2[h]);
      }
      *reinterpret_cast<float4 *>(dens_f + idx0 + 2 * kp) = make_float4(dk2[0].x, dk2[0].y, dk2[1].x, dk2[1].y);
      *reinterpret_cast<float4 *>(npot_f + idx0 + 2 * kp) = make_float4(pk2[0].x, pk2[0].y, pk2[1].x, pk2[1].y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fourier.c:394-397 scaling of both grids + compute_sigma_dens (fourier.c:24-79) as one streaming
// pass (used when the FFT epilogue fusion is not: clr_normalize_fields on injected real fields).
__global__ void __launch_bounds__(kThreads)
scale_moments_kernel(const ClrDev d, float *__restrict__ dens, float *__restrict__ npot, double norm,
                     double *__restrict__ mom)
{
  const int half = d.pitch / 2;                    // float2 per row, includes the padding pair
  const long long n2 = (long long)d.nz_here * d.n * half;
  float2 *dd = reinterpret_cast<float2 *>(dens);
  float2 *pp = reinterpret_cast<float2 *>(npot);
  double a1 = 0, a2 = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    int xq = (int)(i % half);
    float2 v = dd[i], w = pp[i];
    v.x = (float)((double)v.x * norm); v.y = (float)((double)v.y * norm);
    w.x = (float)((double)w.x * norm); w.y = (float)((double)w.y * norm);
    dd[i] = v; pp[i] = w;
    if (2 * xq < d.n) {                            // padding columns do not enter the moments
      a1 += (double)v.x + (double)v.y;
      a2 += (double)(v.x * v.x) + (double)(v.y * v.y);
    }
  }
  a1 = clr_warp_sum(a1); a2 = clr_warp_sum(a2);
  __shared__ double red[2][kThreads / 32];
  int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (ln == 0) { red[0][w] = a1; red[1][w] = a2; }
  __syncthreads();
  if (w == 0) {
    double a = ln < kThreads / 32 ? red[0][ln] : 0, b = ln < kThreads / 32 ? red[1][ln] : 0;
    a = clr_warp_sum(a); b = clr_warp_sum(b);
    if (ln == 0) { atomicAdd(mom, a); atomicAdd(mom + 1, b); }
  }
}

// ---------------------------------------------------------------------------------------------
// lognormalize (density.c:1070-1103) / densclip (density.c:1034-1067): in place, cell coordinates
// in fp32 exactly as the reference (flouble dx, x0, y0, z0), growth factor lerp and exp in double.
template <bool EXACT>
__global__ void __launch_bounds__(kThreads)
lognormal_kernel(const ClrDev d, float *__restrict__ dens, double sigma2, int clip)
{
  const int halfn = d.n / 2;                       // float2 per row holding real cells
  const long long n2 = (long long)d.nz_here * d.n * halfn;
  const float idr = (float)d.glob_idr, rtab = (float)d.r_tab_max, dlast = __ldg(d.d1_f + CLR_NA - 1);
  const float hs2 = (float)(0.5 * sigma2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    int ix0, iy, iz;
    clr_cell(d, 2 * i, ix0, iy, iz);
    int xq = ix0 >> 1;
    long long row = (long long)iz * d.n + iy;
    float z0 = __ldg(d.cf[2] + iz + d.iz0_here);
    float y0 = __ldg(d.cf[1] + iy);
    float2 *p = reinterpret_cast<float2 *>(dens + row * d.pitch) + xq;
    float2 v = *p;
    float out[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      int ix = 2 * xq + q;
      float x0 = __ldg(d.cf[0] + ix);
      // reference: sqrt(x0*x0+y0*y0+z0*z0) with float products, left-to-right float sums
      float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(y0, y0)), __fmul_rn(z0, z0));
      float delta = q ? v.y : v.x;
      if (EXACT) {
        double r = sqrt((double)r2);
        double dg = clr_bg_d1(d, r);
        double res = clip ? fmax(1 + dg * (double)delta, 0.) - 1 : exp(dg * ((double)delta - 0.5 * dg * sigma2)) - 1;
        out[q] = (float)res;
      } else {
        // fp32 evaluation (the grid is fp32): growth-factor lerp on the fp32 copy of the table, expm1f
        // keeps full relative accuracy for small arguments. |error| <~ 3e-7 (1+|result|).
        float r = sqrtf(r2);
        float dg;
        if (r <= 0.f) dg = 1.f;
        else if (r >= rtab) dg = dlast;
        else {
          float t = r * idr;
          int ir = (int)t;
          float fa = __ldg(d.d1_f + ir), fb = __ldg(d.d1_f + ir + 1);
          dg = fa + (fb - fa) * (t - (float)ir);
        }
        out[q] = clip ? fmaxf(1.f + dg * delta, 0.f) - 1.f : expm1f(dg * (delta - hs2 * dg));
      }
    }
    *p = make_float2(out[0], out[1]);
  }
}

// Default (fp32) variant: a run of 8 consecutive cells per thread so that the index arithmetic and the
// y/z coordinate loads are paid once per run; exp(x)-1 through ex2.approx (absolute error ~2e-7 exp(x),
// i.e. 2e-7 relative on the physical density 1+delta, inside the stated fp32 tolerance).
__global__ void __launch_bounds__(kThreads)
lognormal_fast_kernel(const ClrDev d, float *__restrict__ dens, float hs2, int clip)
{
  // a warp owns 256-cell segments of a row: lane <-> float2 number lane + 32*q, so every load / store
  // instruction of the warp covers 256 contiguous bytes (rows are only 8-byte aligned: no float4)
  const float idr = (float)d.glob_idr, rtab = (float)d.r_tab_max, dlast = __ldg(d.d1_f + CLR_NA - 1);
  const int lane = threadIdx.x & 31;
  const unsigned spr = ((unsigned)d.n + 255u) >> 8;                     // segments per row
  const unsigned n_seg = (unsigned)d.nz_here * (unsigned)d.n * spr;      // < 2^32 up to n = 4096
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seg < n_seg; seg += n_warps) {
    const unsigned row = seg / spr;
    const int s = (int)(seg - row * spr);
    const int iz = (int)(row / (unsigned)d.n), iy = (int)(row - (unsigned)iz * (unsigned)d.n);
    const float y0 = __ldg(d.cf[1] + iy), z0 = __ldg(d.cf[2] + iz + d.iz0_here);
    const float yz = y0 * y0 + z0 * z0;
    const int xq0 = s * 128 + lane;                                     // float2 index inside the row
    float2 *p = reinterpret_cast<float2 *>(dens + (long long)row * d.pitch) + xq0;
    const float2 *xc = reinterpret_cast<const float2 *>(d.cf[0]) + xq0;
    float2 v[4], x[4];
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      ok[q] = 2 * (xq0 + 32 * q) < d.n;
      if (ok[q]) { v[q] = p[32 * q]; x[q] = __ldg(xc + 32 * q); }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (!ok[q]) continue;
      float o[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        float x0 = h ? x[q].y : x[q].x, delta = h ? v[q].y : v[q].x;
        float r2 = fmaf(x0, x0, yz);
        float r = clr_sqrt_fast(r2);                                    // 0 at r2 = 0 -> D(r_0) = 1
        float t = fminf(r * idr, (float)(CLR_NA - 2) + 0.5f);
        float m = clr_floor_magic(t);
        float2 e = __ldg(d.d1_t + clr_magic_int(m));
        float dg = r >= rtab ? dlast : fmaf(e.y, t - (m - 8388608.f), e.x);
        o[h] = clip ? fmaxf(fmaf(dg, delta, 1.f), 0.f) - 1.f : clr_ex2_fast(1.4426950408889634f * dg * fmaf(-hs2, dg, delta)) - 1.f;
      }
      p[32 * q] = make_float2(o[0], o[1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// collect_density_normalization_from_grid (density.c:1128-1211): per redshift bin, the number of
// cells, the sum of z and the sum of bias_model(delta, b(r)) per population. Warp-aggregated
// shared-memory histogram (neighbouring cells mostly share a bin), one global atomic per bin and
// CTA at the end.
#define CLR_MAX_NORM_POP 20
#define CLR_MAX_NZ 512
struct NormPops { const double *bz[CLR_MAX_NORM_POP]; int npop; };

template <bool EXACT>
__global__ void __launch_bounds__(kThreads)
norm_hist_kernel(const ClrDev d, const float *__restrict__ dens, NormPops pops, int nz, double idz,
                 unsigned long long *__restrict__ g_n, double *__restrict__ g_z, double *__restrict__ g_b)
{
  const float idrf = (float)d.glob_idr, rtabf = (float)d.r_tab_max, zlastf = __ldg(d.z_f + CLR_NA - 1);
  const float idzf = (float)idz;
  extern __shared__ double sh[];          // [nz] z sums, [npop][nz] bias sums, then counts
  double *s_z = sh;
  double *s_b = sh + nz;
  unsigned long long *s_n = reinterpret_cast<unsigned long long *>(sh + (size_t)nz * (1 + pops.npop));
  for (int i = threadIdx.x; i < nz * (1 + pops.npop); i += blockDim.x) sh[i] = 0;
  for (int i = threadIdx.x; i < nz; i += blockDim.x) s_n[i] = 0;
  __syncthreads();
  const long long n_cells = (long long)d.nz_here * d.n * d.n;
  const long long n_iter = (n_cells + (long long)gridDim.x * blockDim.x - 1) / ((long long)gridDim.x * blockDim.x);
  for (long long it = 0; it < n_iter; it++) {
    long long i = (it * gridDim.x + blockIdx.x) * (long long)blockDim.x + threadIdx.x;
    int bin = -1;
    double redshift = 0, dcell = 0, r = 0;
    if (i < n_cells) {
      int ix, iy, iz;
      clr_cell(d, i, ix, iy, iz);
      long long row = (long long)iz * d.n + iy;
      float z0 = __ldg(d.cf[2] + iz + d.iz0_here);
      float y0 = __ldg(d.cf[1] + iy);
      float x0 = __ldg(d.cf[0] + ix);
      float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(y0, y0)), __fmul_rn(z0, z0));
      int ind_z;
      if (EXACT) {
        r = sqrt((double)r2);
        redshift = clr_bg_z(d, r);
        ind_z = (int)(redshift * idz) + 1;
      } else {
        // fp32 screening of the bin; cells closer than 1e-4 bins to an edge take the double path
        float rf = sqrtf(r2), zf;
        if (rf <= 0.f) zf = 0.f;
        else if (rf >= rtabf) zf = zlastf;
        else {
          float t = rf * idrf;
          int ir = (int)t;
          float fa = __ldg(d.z_f + ir), fb = __ldg(d.z_f + ir + 1);
          zf = fa + (fb - fa) * (t - (float)ir);
        }
        float tb = zf * idzf;
        if (fabsf(tb - rintf(tb)) < 1e-4f) {
          r = sqrt((double)r2);
          redshift = clr_bg_z(d, r);
          ind_z = (int)(redshift * idz) + 1;
        } else {
          r = rf;
          redshift = zf;
          ind_z = (int)tb + 1;
        }
      }
      if (ind_z >= 0 && ind_z < nz) { bin = ind_z; dcell = dens[row * d.pitch + ix]; }
    }
    // warp aggregation by bin
    unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
    const int lane = threadIdx.x & 31;
    while (todo) {
      int leader = __ffs(todo) - 1;
      int b = __shfl_sync(0xffffffffu, bin, leader);
      bool mine = (bin == b);
      unsigned grp = __ballot_sync(0xffffffffu, mine);
      double zsum = clr_warp_sum(mine ? redshift : 0.);
      if (lane == leader) { atomicAdd(&s_z[b], zsum); atomicAdd(&s_n[b], (unsigned long long)__popc(grp)); }
      for (int ip = 0; ip < pops.npop; ip++) {
        double bm = 0.;
        if (mine) {
          if (EXACT) bm = clr_bias_model(d.bias_model, dcell, clr_bg_bz(d, r, pops.bz[ip]));
          else {
            // fp32 bias lerp and bias_model (common.h:414-431); summed in double
            float rf = (float)r, dl = (float)dcell, bi;
            const double *tb = pops.bz[ip];
            if (rf <= 0.f) bi = (float)__ldg(tb);
            else if (rf >= rtabf) bi = 1.f;
            else {
              float t = rf * idrf;
              int ir = (int)t;
              float fa = (float)__ldg(tb + ir), fb = (float)__ldg(tb + ir + 1);
              bi = fa + (fb - fa) * (t - (float)ir);
            }
            float v;
            if (dl <= -1.f) v = 0.f;
            else if (d.bias_model == 2) v = dl < 0.f ? expf(bi * dl / (1.f + dl)) : 1.f + bi * dl;
            else if (d.bias_model == 3) v = fmaxf(1.f + bi * dl, 0.f);
            else v = powf(1.f + dl, bi);
            bm = v;
          }
        }
        bm = clr_warp_sum(bm);
        if (lane == leader) atomicAdd(&s_b[ip * nz + b], bm);
      }
      todo &= ~grp;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nz; i += blockDim.x) {
    if (s_n[i]) {
      atomicAdd(&g_n[i], s_n[i]);
      atomicAdd(&g_z[i], s_z[i]);
      for (int ip = 0; ip < pops.npop; ip++) atomicAdd(&g_b[ip * nz + i], s_b[ip * nz + i]);
    }
  }
}

// Fast variant of the same histogram (default path; kernel further down): fp32 evaluation of z(r), b(r) and
// bias_model with per-lane register sums; the bin index falls back to the exact double expression within
// 1e-4 bins of an edge, so the counts stay exact. Needs an even row length (float2 loads).
constexpr int kRun = 2;
constexpr int kFastPop = 4;
// lerp tables {f[i], f[i+1]-f[i]} in fp32: entry 0 = z(r), entries 1.. = b(r) per population
struct NormPopsF { const float2 *zt; const float4 *zb; const float2 *bt[kFastPop]; int npop; };
// lognormal / clip transform fused into the histogram walk (XFORM): growth table, 0.5*sigma^2, dens_type 3
struct XformArgs { const float2 *d1_t; float hs2, dlast; int clip; };

__device__ __forceinline__ float bias_model_f(int model, float dl, float bi)
{
  // branch free on the data (delta < 0 for ~60 % of the cells: both sides would run in every warp anyway)
  if (model == 2) {
    float e = clr_ex2_fast(1.4426950408889634f * bi * dl * clr_rcp_fast(fmaxf(1.f + dl, 1e-30f)));
    float v = dl < 0.f ? e : fmaf(bi, dl, 1.f);
    return dl <= -1.f ? 0.f : v;
  }
  if (dl <= -1.f) return 0.f;
  if (model == 3) return fmaxf(1.f + bi * dl, 0.f);
  return __powf(1.f + dl, bi);
}

// Column-strip walk: a warp owns a strip of 64 cells along x (lane <-> one float2) and walks DOWN the rows of a
// contiguous row range, so a lane's x never changes, y moves one cell per step and its redshift bin (a ~60-cell
// thick shell) changes every ~60 steps: the lane keeps {count, sum z, sum bias_model} of its current bin in
// registers and touches the shared-memory histogram only when the bin changes. The 8 warps of a CTA take 8
// neighbouring strips of the same rows (2 KB contiguous per row step).
// XFORM: the walk first applies lognormalize / densclip (density.c:1034-1103, the arithmetic of lognormal_fast_kernel) to
// the Gaussian cell, stores it, and bins the transformed value: one read of the field for both stages, radius and table
// position computed once.
// BM: bias model fixed at compile time (2 = the reference's default build, common.h:414-431), 0 = d.bias_model
template <int NPOP, int XFORM, int BM>      // XFORM: 0 = histogram only, 1 = lognormalize first, 2 = densclip first
__global__ void __launch_bounds__(kThreads)
norm_hist_fast_kernel(const ClrDev d, float *__restrict__ dens, NormPopsF pops, int nz, double idz,
                      unsigned long long *__restrict__ g_n, double *__restrict__ g_z, double *__restrict__ g_b,
                      const XformArgs xf)
{
  // CTA histogram in shared memory, updated with 32-bit integer atomics only (the float / double / 64-bit
  // shared atomics of sm_100 are CAS spin loops): counts as u32, sums as 64-bit fixed point (2^-20) held
  // as {lo, hi} words with an explicit carry. Integer sums also make the result independent of the
  // order in which the threads flush.
  extern __shared__ double sh[];
  unsigned *s_n = reinterpret_cast<unsigned *>(sh);                    // [nz]
  unsigned *s_q = s_n + nz;                                            // [(1+npop)][nz][2]: z, then b per population
  for (int i = threadIdx.x; i < nz * (3 + 2 * pops.npop); i += blockDim.x) s_n[i] = 0;
  __syncthreads();
  auto add_fixed = [&](int slot, int bin, float v) {
    long long f = __float2ll_rn(v * 1048576.f);
    unsigned lo = (unsigned)f, hi = (unsigned)((unsigned long long)f >> 32);
    unsigned *w = s_q + ((size_t)slot * nz + bin) * 2;
    unsigned old = atomicAdd(w, lo);
    hi += (old + lo < old) ? 1u : 0u;
    if (hi) atomicAdd(w + 1, hi);
  };
  const float idrf = (float)d.glob_idr, rtabf = (float)d.r_tab_max;
  const float zlastf = __ldg(&pops.zt[CLR_NA - 1].x);
  const float idzf = (float)idz;
  constexpr int npop = NPOP;          // compile-time population count: dead per-population code folds away
  const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int nstrips = (d.n + 63) >> 6;
  const int groups = (nstrips + wpc - 1) / wpc;                        // CTAs side by side along x
  const int n_ranges = max(1, (int)gridDim.x / groups);
  const int strip = (int)(blockIdx.x % groups) * wpc + wip;
  const int range = (int)(blockIdx.x / groups);
  const unsigned n_rows = (unsigned)d.nz_here * (unsigned)d.n;
  const unsigned per = (n_rows + n_ranges - 1) / n_ranges;
  const unsigned row0 = range * per, row1 = min(n_rows, row0 + per);
  const int xq = strip * 32 + lane;                                   // float2 index inside the row
  const bool ok = strip < nstrips && 2 * xq < d.n && range < n_ranges;
  int curbin = -1, cnt = 0;
  float zs = 0.f, bs[kFastPop];
#pragma unroll
  for (int ip = 0; ip < kFastPop; ip++) bs[ip] = 0.f;
  auto flush = [&]() {
    if (curbin >= 0 && cnt > 0) {
      atomicAdd(&s_n[curbin], (unsigned)cnt);
      add_fixed(0, curbin, zs);
#pragma unroll
      for (int ip = 0; ip < kFastPop; ip++) if (ip < npop) add_fixed(1 + ip, curbin, bs[ip]);
    }
    cnt = 0; zs = 0.f;
#pragma unroll
    for (int ip = 0; ip < kFastPop; ip++) bs[ip] = 0.f;
  };
  if (ok) {
    const float2 xc = __ldg(reinterpret_cast<const float2 *>(d.cf[0]) + xq);
    const float xx[2] = {__fmul_rn(xc.x, xc.x), __fmul_rn(xc.y, xc.y)};
    int iz = (int)(row0 / (unsigned)d.n), iy = (int)(row0 - (unsigned)iz * (unsigned)d.n);
    float zz = 0.f;
    { const float z0 = __ldg(d.cf[2] + iz + d.iz0_here); zz = __fmul_rn(z0, z0); }
    float2 *p = reinterpret_cast<float2 *>(dens + (long long)row0 * d.pitch) + xq;
    const int pitch2 = d.pitch >> 1;
    float2 dv = row0 < row1 ? *p : make_float2(0.f, 0.f);
    for (unsigned row = row0; row < row1; row++) {
      const float2 dcur = dv;
      float2 *pcur = p;
      float xo[2];
      p += pitch2;
      if (row + 1 < row1) dv = *p;                                     // next row's load flies under this row's math
      const float y0 = __ldg(d.cf[1] + iy);
      const float yy = __fmul_rn(y0, y0);
      int bin2[2];
      float zf2[2], bm2[2][kFastPop];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        float dl = h ? dcur.y : dcur.x;
        const float r2 = __fadd_rn(__fadd_rn(xx[h], yy), zz);          // same order as the reference
        const float rf = clr_sqrt_fast(r2);
        const float tr = fminf(rf * idrf, (float)(CLR_NA - 2) + 0.5f);
        const float mr = clr_floor_magic(tr);
        const int ir = clr_magic_int(mr);
        const float fr = tr - (mr - 8388608.f);
        // pops.zb: {z_i, z_i+1 - z_i, b_i, b_i+1 - b_i} of the first population: one 16-byte load per cell
        const float4 tz = __ldg(pops.zb + ir);
        const bool past = rf >= rtabf;
        if (XFORM) {
          const float2 e = __ldg(xf.d1_t + ir);
          const float dg = past ? xf.dlast : fmaf(e.y, fr, e.x);
          dl = XFORM == 2 ? fmaxf(fmaf(dg, dl, 1.f), 0.f) - 1.f : clr_ex2_fast(1.4426950408889634f * dg * fmaf(-xf.hs2, dg, dl)) - 1.f;
          xo[h] = dl;
        }
        const float zf = past ? zlastf : fmaf(tz.y, fr, tz.x);
        const float tb = zf * idzf;
        const float tbn = __fadd_rn(__fadd_rn(tb, 12582912.f), -12582912.f);   // nearest integer, conversion-free
        int ind_z = clr_magic_int(clr_floor_magic(tb)) + 1;            // tb >= 0
        if (fabsf(tb - tbn) < 1e-4f) {                                 // ~2e-4 of the cells: exact edge decision
          double redshift = clr_bg_z(d, sqrt((double)r2));            // density.c:1166-1168 verbatim
          ind_z = (int)(redshift * idz) + 1;
        }
        bin2[h] = ((unsigned)ind_z < (unsigned)nz) ? ind_z : -1;
        zf2[h] = zf;
#pragma unroll
        for (int ip = 0; ip < kFastPop; ip++) {
          bm2[h][ip] = 0.f;
          if (ip < npop) {
            float bi;
            if (ip == 0) bi = past ? 1.f : fmaf(tz.w, fr, tz.z);
            else { float2 t = __ldg(pops.bt[ip] + ir); bi = past ? 1.f : fmaf(t.y, fr, t.x); }
            bm2[h][ip] = bias_model_f(BM ? BM : d.bias_model, dl, bi);
          }
        }
      }
      if (XFORM) *pcur = make_float2(xo[0], xo[1]);
      if (bin2[0] == curbin && bin2[1] == curbin) {                    // the common case: plain register sums
        cnt += 2;
        zs += zf2[0] + zf2[1];
#pragma unroll
        for (int ip = 0; ip < kFastPop; ip++) if (ip < npop) bs[ip] += bm2[0][ip] + bm2[1][ip];
      } else {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (bin2[h] != curbin) { flush(); curbin = bin2[h]; }
          cnt++;
          zs += zf2[h];
#pragma unroll
          for (int ip = 0; ip < kFastPop; ip++) if (ip < npop) bs[ip] += bm2[h][ip];
        }
      }
      if (cnt >= 4096) flush();          // keep the fp32 partial sums short
      if (++iy == d.n) {
        iy = 0; iz++;
        if (row + 1 < row1) { const float z0 = __ldg(d.cf[2] + iz + d.iz0_here); zz = __fmul_rn(z0, z0); }
      }
    }
    flush();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nz; i += blockDim.x) {
    if (s_n[i]) {
      auto fixed = [&](int slot) {
        const unsigned *w = s_q + ((size_t)slot * nz + i) * 2;
        return (double)(long long)(((unsigned long long)w[1] << 32) | w[0]) * (1.0 / 1048576.0);
      };
      atomicAdd(&g_n[i], (unsigned long long)s_n[i]);
      atomicAdd(&g_z[i], fixed(0));
      for (int ip = 0; ip < npop; ip++) atomicAdd(&g_b[ip * nz + i], fixed(1 + ip));
    }
  }
}

// {z[i], z[i+1]-z[i], b[i], b[i+1]-b[i]}: redshift and first-population bias lerp entries side by side
__global__ void lerp_table2_kernel(const double *__restrict__ za, const double *__restrict__ ba, float4 *__restrict__ dst, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = i + 1 < n ? i + 1 : i;
  double z0 = za[i], z1 = za[j], b0 = ba ? ba[i] : 1., b1 = ba ? ba[j] : 1.;
  dst[i] = make_float4((float)z0, (float)(z1 - z0), (float)b0, (float)(b1 - b0));
}

// {f[i], f[i+1]-f[i]} in fp32 from a double table of n entries
__global__ void lerp_table_kernel(const double *__restrict__ src, float2 *__restrict__ dst, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = src[i], b = src[i + 1 < n ? i + 1 : i];
  dst[i] = make_float2((float)a, (float)(b - a));
}

// z-halo of the potential for a single slab: periodic wrap (fourier.c:412-413)
__global__ void halo_copy_kernel(float *__restrict__ dst, const float *__restrict__ src, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

int grid_for(clr_ctx *c, long long work_items, int per_sm)
{
  long long g = (work_items + kThreads - 1) / kThreads;
  long long cap = (long long)c->sm_count * per_sm;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int clr_ensure_scratch(clr_ctx *c, size_t bytes)
{
  if (bytes <= c->scratch_bytes) return 0;
  if (c->d_scratch) cudaFree(c->d_scratch);
  c->d_scratch = nullptr; c->scratch_bytes = 0;
  CLR_CUDA(cudaMalloc(&c->d_scratch, bytes));
  c->scratch_bytes = bytes;
  return 0;
}

int clr_fill_fast_setup(clr_ctx *c, FillFastK *kp)
{
  const double dk = 2 * M_PI / c->p.l_box, idk3 = 1. / (dk * dk * dk), lgdk = log10(dk);
  const int numk = c->p.numk;
  if (!c->d_pkt) {
    std::vector<float2> t(numk);
    for (int i = 0; i < numk; i++) {
      double p0 = c->h_pk[i], p1 = i + 1 < numk ? c->h_pk[i + 1] : p0;
      t[i] = make_float2((float)(p0 * idk3), (float)((p1 - p0) * idk3));
    }
    CLR_CUDA(cudaMalloc(&c->d_pkt, numk * sizeof(float2)));
    CLR_CUDA(cudaMemcpy(c->d_pkt, t.data(), numk * sizeof(float2), cudaMemcpyHostToDevice));
    std::vector<float2> sc(4096);
    for (int i = 0; i < 4096; i++) sc[i] = make_float2((float)cos(2 * M_PI * i / 4096.), (float)sin(2 * M_PI * i / 4096.));
    CLR_CUDA(cudaMalloc(&c->d_sincos, 4096 * sizeof(float2)));
    CLR_CUDA(cudaMemcpy(c->d_sincos, sc.data(), 4096 * sizeof(float2), cudaMemcpyHostToDevice));
  }
  FillFastK &k = *kp;
  for (int e = 0; e < 26; e++) {
    double v = (lgdk + 0.15051499783199060 * e - c->p.logkmin) * c->p.idlogk;
    double fl = floor(v);
    k.e_int[e] = (int)fl;
    k.e_frac[e] = (float)(v - fl);
  }
  k.c1 = (float)(0.15051499783199060 * c->p.idlogk);
  k.nscal_c = (float)(c->p.n_scal / c->p.idlogk);
  k.m3_c = (float)(-3. / c->p.idlogk);
  k.tmax = (float)((c->p.logkmax - c->p.logkmin) * c->p.idlogk);
  k.p_first = (float)(c->h_pk[0] * idk3);
  k.p_last = (float)(c->h_pk[numk - 1] * idk3);
  k.neg_prefac_idk2 = (float)(-c->p.prefac_lensing / (dk * dk));
  k.smooth_c = (float)(-0.5 * c->p.r2_smooth * dk * dk);
  k.smooth_c2 = (float)(-0.5 * c->p.r2_smooth * dk * dk * 1.4426950408889634);   // for ex2
  k.numk = numk; k.do_smoothing = c->p.do_smoothing; k.smooth_potential = c->p.smooth_potential;
  return 0;
}

// the fp32 fill needs k^2/dk^2 < 2^24 (exact integer in a float): n_grid <= 4096
bool clr_fill_fast_ok(const clr_ctx *c) { return !c->exact_math && 3LL * (c->dev.n / 2) * (c->dev.n / 2) < (1LL << 24); }

int clr_fields_fill_fast(clr_ctx *c, uint32_t seed)
{
  FillFastK k;
  if (clr_fill_fast_setup(c, &k)) return 1;
  int kxs_log2 = 5;      // lanes per row: one lane per mode PAIR, whole warps
  while ((1 << kxs_log2) < (c->dev.nc + 1) / 2 && (1 << kxs_log2) < kThreads) kxs_log2++;
  const int rows_par = kThreads >> kxs_log2;
  long long n_rows = (long long)c->dev.n * c->dev.nyl;
  fill_modes_fast_kernel<<<grid_for(c, (n_rows + rows_par - 1) / rows_par * kThreads, 8), kThreads, 0, c->stream>>>(
      c->dev, reinterpret_cast<float2 *>(c->d_dens), reinterpret_cast<float2 *>(c->d_npot), seed, c->d_pkt, c->d_sincos,
      k, kxs_log2);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

int clr_fields_fill(clr_ctx *c, uint32_t seed)
{
  StageScope sc(c, "fill_modes", 1);
  if (clr_fill_fast_ok(c)) return clr_fields_fill_fast(c, seed);
  long long n_modes = (long long)c->dev.nz_here * c->dev.n * c->dev.nc;
  double lgdk = log10(2 * M_PI / c->p.l_box);
  auto k = c->exact_math ? fill_modes_kernel<true> : fill_modes_kernel<false>;
  unsigned rpb = std::max(1u, 1024u / (unsigned)c->dev.nc);
  long long n_groups = ((long long)c->dev.n * c->dev.nyl + rpb - 1) / rpb;
  (void)n_modes;
  k<<<grid_for(c, n_groups * kThreads, 8), kThreads, 0, c->stream>>>(
      c->dev, reinterpret_cast<float2 *>(c->d_dens), reinterpret_cast<float2 *>(c->d_npot), seed, c->d_pk,
      c->d_pk + c->p.numk, c->p.numk, c->p.logkmin, c->p.logkmax, c->p.idlogk, c->p.n_scal, c->p.prefac_lensing,
      c->p.r2_smooth, c->p.do_smoothing, c->p.smooth_potential, lgdk);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

int clr_fields_scale_moments(clr_ctx *c, double *out2)
{
  if (clr_ensure_scratch(c, 4096)) return 1;
  CLR_CUDA(cudaMemsetAsync(c->d_scratch, 0, 2 * sizeof(double), c->stream));
  double norm = pow(sqrt(2 * M_PI) / c->p.l_box, 3);
  {
    StageScope sc(c, "scale_moments", 1);
    long long n2 = (long long)c->dev.nz_here * c->dev.n * (c->dev.pitch / 2);
    scale_moments_kernel<<<grid_for(c, n2, 8), kThreads, 0, c->stream>>>(c->dev, c->d_dens, c->d_npot, norm, c->d_scratch);
    CLR_CUDA(cudaGetLastError());
  }
  if (clr_comm_allreduce_f64(c, c->d_scratch, 2)) return 1;
  if (clr_read_small(c, out2, c->d_scratch, 2 * sizeof(double))) return 1;
  return 0;
}

// bins of compute_density_normalization (density.c:1233-1245): nz = (int)(z(L/2)/0.05)+2, idz = (nz-2)/z(L/2)
static void norm_bins(const clr_ctx *c, int *nz, double *idz)
{
  const double r = (double)(c->p.l_box * 0.5);
  double zmax;
  if (r <= 0) zmax = 0;
  else if (r >= c->h_r[CLR_NA - 1]) zmax = c->h_z[CLR_NA - 1];
  else {
    int ir = (int)(r * c->p.glob_idr);
    zmax = c->h_z[ir] + (c->h_z[ir + 1] - c->h_z[ir]) * (r - c->h_r[ir]) * c->p.glob_idr;
  }
  *nz = (int)(zmax / 0.05) + 2;
  *idz = (*nz - 2) / zmax;
}

// every population that enters the normalisation, in the order of density.c:1246-1260 (sources, then intensity maps)
static int norm_pops(clr_ctx *c, const double **d_bz)
{
  int npop = 0;
  for (int i = 0; i < CLR_NPOP_MAX; i++) if (c->srcs[i].set) { if (d_bz) d_bz[npop] = c->srcs[i].d_b; npop++; }
  for (int i = 0; i < CLR_NPOP_MAX; i++) if (c->imap[i].set) { if (d_bz) d_bz[npop] = c->imap[i].d_b; npop++; }
  return npop;
}

// Launch the fp32 histogram walk into c->d_hist ({counts[nz], sum z[nz], sum bias_model[npop][nz]}), optionally with
// the lognormal / clip transform fused in (xf != nullptr).
static int launch_hist_walk(clr_ctx *c, int npop, const double *const *d_bz, int nz, double idz, const XformArgs *xf)
{
  const size_t nd = (size_t)nz * (2 + npop);
  const size_t tab_bytes = (size_t)(kFastPop + 3) * CLR_NA * sizeof(float2) + 64;
  if (!c->d_hist || c->hist_bytes < nd * sizeof(double) + tab_bytes) {
    if (c->d_hist) cudaFree(c->d_hist);
    c->d_hist = nullptr;
    c->hist_bytes = nd * sizeof(double) + tab_bytes;
    CLR_CUDA(cudaMalloc(&c->d_hist, c->hist_bytes));
  }
  CLR_CUDA(cudaMemsetAsync(c->d_hist, 0, nd * sizeof(double), c->stream));
  unsigned long long *g_n = reinterpret_cast<unsigned long long *>(c->d_hist);
  double *g_z = c->d_hist + nz;
  double *g_b = c->d_hist + 2 * nz;
  // fp32 lerp tables of z(r) and the b(r) live behind the histograms
  NormPopsF pf;
  pf.npop = npop;
  float2 *tab = reinterpret_cast<float2 *>(c->d_hist + nd);
  lerp_table_kernel<<<(CLR_NA + 255) / 256, 256, 0, c->stream>>>(c->dev.z_arr, tab, CLR_NA);
  pf.zt = tab;
  for (int i = 0; i < npop; i++) {
    lerp_table_kernel<<<(CLR_NA + 255) / 256, 256, 0, c->stream>>>(d_bz[i], tab + (size_t)(i + 1) * CLR_NA, CLR_NA);
    pf.bt[i] = tab + (size_t)(i + 1) * CLR_NA;
  }
  for (int i = npop; i < kFastPop; i++) pf.bt[i] = tab;
  float4 *tab4 = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(tab + (size_t)(kFastPop + 1) * CLR_NA) + 15) & ~(uintptr_t)15);
  lerp_table2_kernel<<<(CLR_NA + 255) / 256, 256, 0, c->stream>>>(c->dev.z_arr, npop ? d_bz[0] : nullptr, tab4, CLR_NA);
  pf.zb = tab4;
  c->launches += 2 + npop;                     // the lerp-table kernels above
  // CTAs side by side along x (8 strips of 64 cells each) x row ranges; 8 resident CTAs per SM
  const int groups = ((c->dev.n + 63) / 64 + 7) / 8;
  long long n_rows = (long long)c->dev.nz_here * c->dev.n;
  long long ranges = std::min<long long>(n_rows, std::max<long long>(1, (long long)c->sm_count * 8 / groups));
  int grid = (int)(ranges * groups);
  const size_t smem = nd * sizeof(double);
  XformArgs x0{nullptr, 0.f, 0.f, 0};
#define CLR_HIST(NP)                                                                                                                 \
  if (xf && !xf->clip && c->dev.bias_model == 2)                                                                                     \
    norm_hist_fast_kernel<NP, 1, 2><<<grid, kThreads, smem, c->stream>>>(c->dev, c->d_dens, pf, nz, idz, g_n, g_z, g_b, *xf);        \
  else if (xf && !xf->clip) norm_hist_fast_kernel<NP, 1, 0><<<grid, kThreads, smem, c->stream>>>(c->dev, c->d_dens, pf, nz, idz, g_n, g_z, g_b, *xf); \
  else if (xf) norm_hist_fast_kernel<NP, 2, 0><<<grid, kThreads, smem, c->stream>>>(c->dev, c->d_dens, pf, nz, idz, g_n, g_z, g_b, *xf); \
  else norm_hist_fast_kernel<NP, 0, 0><<<grid, kThreads, smem, c->stream>>>(c->dev, c->d_dens, pf, nz, idz, g_n, g_z, g_b, x0)
  switch (npop) {
    case 0: CLR_HIST(0); break;
    case 1: CLR_HIST(1); break;
    case 2: CLR_HIST(2); break;
    case 3: CLR_HIST(3); break;
    default: CLR_HIST(4); break;
  }
#undef CLR_HIST
  CLR_CUDA(cudaGetLastError());
  return 0;
}

int clr_fields_lognormal(clr_ctx *c, int clip)
{
  c->hist_valid = false;
  long long n2 = (long long)c->dev.nz_here * c->dev.n * (c->dev.n / 2);
  // populations already known (the usual run flow: read_run_params before the fields): do the normalisation histogram
  // of compute_density_normalization in the same pass over the field and keep it for that call
  const double *d_bz[CLR_MAX_NORM_POP];
  const int npop = norm_pops(c, nullptr);
  if (!c->exact_math && c->hist_fused && npop >= 1 && npop <= kFastPop && c->dev.n % kRun == 0) {
    int nz; double idz;
    norm_bins(c, &nz, &idz);
    if (nz <= CLR_MAX_NZ) {
      norm_pops(c, d_bz);
      StageScope sc(c, "lognormal", 1);
      XformArgs xf{c->dev.d1_t, (float)(0.5 * c->sigma2_gauss), (float)c->h_d1[CLR_NA - 1], clip};
      if (launch_hist_walk(c, npop, d_bz, nz, idz, &xf)) return 1;
      c->hist_valid = true; c->hist_npop = npop; c->hist_nz = nz;
      for (int i = 0; i < npop; i++) c->hist_bz[i] = d_bz[i];
      return 0;
    }
  }
  StageScope sc(c, "lognormal", 1);
  if (c->exact_math)
    lognormal_kernel<true><<<grid_for(c, n2, 8), kThreads, 0, c->stream>>>(c->dev, c->d_dens, c->sigma2_gauss, clip);
  else if (c->dev.n % 8 == 0)
    lognormal_fast_kernel<<<grid_for(c, n2 / 4, 8), kThreads, 0, c->stream>>>(c->dev, c->d_dens, (float)(0.5 * c->sigma2_gauss), clip);
  else
    lognormal_kernel<false><<<grid_for(c, n2, 8), kThreads, 0, c->stream>>>(c->dev, c->d_dens, c->sigma2_gauss, clip);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

int clr_fields_norm_hist(clr_ctx *c, int npop, const double *const *d_bz, int nz, double idz,
                         unsigned long long *h_n, double *h_z, double *h_b)
{
  CLR_CHECK(npop <= CLR_MAX_NORM_POP && nz <= CLR_MAX_NZ, "normalisation: npop=%d nz=%d too large", npop, nz);
  size_t nd = (size_t)nz * (2 + npop);
  double *d_res = nullptr;
  bool cached = c->hist_valid && c->hist_npop == npop && c->hist_nz == nz && !c->exact_math;
  for (int i = 0; i < npop && cached; i++) cached = c->hist_bz[i] == d_bz[i];
  c->hist_valid = false;                         // one use: the all-reduce below runs in place
  if (cached) d_res = c->d_hist;                 // filled by the fused lognormal + histogram pass on this very field
  else if (!c->exact_math && npop <= kFastPop && c->dev.n % kRun == 0) {
    StageScope sc(c, "norm_hist", 1);
    if (launch_hist_walk(c, npop, d_bz, nz, idz, nullptr)) return 1;
    d_res = c->d_hist;
  } else {
    if (clr_ensure_scratch(c, nd * sizeof(double) + 64)) return 1;
    CLR_CUDA(cudaMemsetAsync(c->d_scratch, 0, nd * sizeof(double), c->stream));
    NormPops pops;
    pops.npop = npop;
    for (int i = 0; i < npop; i++) pops.bz[i] = d_bz[i];
    StageScope sc(c, "norm_hist", 1);
    long long n_cells = (long long)c->dev.nz_here * c->dev.n * c->dev.n;
    size_t smem = nd * sizeof(double);
    unsigned long long *g_n = reinterpret_cast<unsigned long long *>(c->d_scratch);
    if (c->exact_math)
      norm_hist_kernel<true><<<grid_for(c, n_cells, 8), kThreads, smem, c->stream>>>(c->dev, c->d_dens, pops, nz, idz, g_n, c->d_scratch + nz, c->d_scratch + 2 * nz);
    else
      norm_hist_kernel<false><<<grid_for(c, n_cells, 8), kThreads, smem, c->stream>>>(c->dev, c->d_dens, pops, nz, idz, g_n, c->d_scratch + nz, c->d_scratch + 2 * nz);
    CLR_CUDA(cudaGetLastError());
    d_res = c->d_scratch;
  }
  // density.c:1262-1269: histograms summed over the slabs
  if (clr_comm_allreduce_u64(c, reinterpret_cast<unsigned long long *>(d_res), nz)) return 1;
  if (clr_comm_allreduce_f64(c, d_res + nz, (size_t)nz * (1 + npop))) return 1;
  {
    // counts, z sums, bias sums are contiguous: one small read-back
    std::vector<double> tmp(nd);
    if (clr_read_small(c, tmp.data(), d_res, nd * sizeof(double))) return 1;
    memcpy(h_n, tmp.data(), nz * sizeof(unsigned long long));
    memcpy(h_z, tmp.data() + nz, nz * sizeof(double));
    if (npop) memcpy(h_b, tmp.data() + 2 * nz, (size_t)npop * nz * sizeof(double));
  }
  return 0;
}

int clr_halo_update(clr_ctx *c)
{
  // single slab: slice_left = last plane, slice_right = first plane (fourier.c:412-413); they are
  // materialised after the grid so that the stencil kernels index them like the reference.
  if (c->nranks > 1) {
    StageScope sc(c, "halo", 0);
    return clr_comm_halo(c);
  }
  StageScope sc(c, "halo", 2);
  long long plane = (long long)c->dev.pitch * c->dev.n;
  float *left = c->d_npot + (long long)c->dev.nz_here * plane;
  float *right = left + plane;
  halo_copy_kernel<<<grid_for(c, plane, 4), kThreads, 0, c->stream>>>(left, c->d_npot + (long long)(c->dev.nz_here - 1) * plane, plane);
  halo_copy_kernel<<<grid_for(c, plane, 4), kThreads, 0, c->stream>>>(right, c->d_npot, plane);
  CLR_CUDA(cudaGetLastError());
  return 0;
}
