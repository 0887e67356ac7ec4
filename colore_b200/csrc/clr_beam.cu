// Line-of-sight tracers that need the CIC branch of interpolate_from_grid (beaming.c:183-265) or per-source rays:
//   * per-source lensing: shear, convergence and deflection of every source (srcs.c:531-614, default build without
//     _USE_FAST_LENSING), NGP velocity + tidal stencils along the ray observer -> source
//   * skewers: density (or Gaussian-field, beaming.c:55-66) and radial-velocity samples along the line of sight of
//     every source (srcs.c:507-529), CIC; post-processing of srcs.c:725-733
//   * custom projected maps (cstm.c:68-145): radial kernel * (bias_model(delta_CIC) * norm - 1) per pixel
// Compiled with -fmad=false: the interpolation weights and sums follow the reference's float / double mix.
//
// Several GPUs: per-source quantities are linear in the field, so every GPU integrates the part of every ray that
// crosses ITS slab for ALL sources (positions all-gathered), the partial results travel back to the rank that owns
// the source and are summed there -- the reference reaches the same sums by rotating the slabs past the sources
// (beaming.c:325-352). The custom map is not linear in delta (bias_model): a sample belongs to the slab that holds
// the lower plane of its CIC pair, the upper plane comes from a one-plane halo of the density.
#include "clr_internal.cuh"
#include "clr_stencil.cuh"
#include <math.h>
#include <algorithm>
#include <vector>

namespace {

constexpr int kThreads = 256;
constexpr double kRtod = 57.2957795;    // common.h:124

int blocks_for(clr_ctx *c, long long items, int per_sm)
{
  long long g = (items + kThreads - 1) / kThreads, cap = (long long)c->sm_count * per_sm;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// trilinear set-up of interpolate_from_grid (beaming.c:186-203) and the eight weights (beaming.c:210-213, 240-243):
// h1x is a flouble, h0x a double, products evaluated left to right and rounded to flouble
struct Cic {
  int x0, x1, y0, y1;
  int z0, z1;               // GLOBAL planes, periodic wrap applied
  float w[2][4];            // [plane z0 / z1][corner (x0,y0), (x1,y0), (x0,y1), (x1,y1)]
};
__device__ __forceinline__ void dev_cic(const ClrDev &d, const double xn[3], Cic &q)
{
  long ix0[3], ix1[3];
  double h0[3];
  float h1[3];
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    ix0[ax] = (long)(xn[ax]);
    h0[ax] = xn[ax] - ix0[ax];
    h1[ax] = (float)(1 - h0[ax]);
    ix1[ax] = ix0[ax] + 1;
    if (ix0[ax] >= d.n) ix0[ax] -= d.n; else if (ix0[ax] < 0) ix0[ax] += d.n;
    if (ix1[ax] >= d.n) ix1[ax] -= d.n; else if (ix1[ax] < 0) ix1[ax] += d.n;
  }
  q.x0 = (int)ix0[0]; q.x1 = (int)ix1[0]; q.y0 = (int)ix0[1]; q.y1 = (int)ix1[1]; q.z0 = (int)ix0[2]; q.z1 = (int)ix1[2];
  q.w[0][0] = h1[2] * h1[1] * h1[0];          q.w[0][1] = (float)(h1[2] * h1[1] * h0[0]);
  q.w[0][2] = (float)(h1[2] * h0[1] * h1[0]); q.w[0][3] = (float)(h1[2] * h0[1] * h0[0]);
  q.w[1][0] = (float)(h0[2] * h1[1] * h1[0]); q.w[1][1] = (float)(h0[2] * h1[1] * h0[0]);
  q.w[1][2] = (float)(h0[2] * h0[1] * h1[0]); q.w[1][3] = (float)(h0[2] * h0[1] * h0[0]);
}

// r of a source as srcs.c:488 computes it: float products and sums, double square root
__device__ __forceinline__ double dev_src_r(const float4 &p)
{
  float r2 = __fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z));
  return sqrt((double)r2);
}

// beaming.c:55-66: the Gaussian field recovered from the lognormal density at grid point (ix, iy, local plane lz)
__device__ __forceinline__ float dev_gauss_element(const ClrDev &d, float dens, int ix, int iy, int lz, double sigma2)
{
  float x0 = __ldg(d.cf[0] + ix), y0 = __ldg(d.cf[1] + iy), z0 = __ldg(d.cf[2] + lz + d.iz0_here);
  float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(y0, y0)), __fmul_rn(z0, z0));
  double r = sqrt((double)r2);
  double dg = clr_bg_d1(d, r);
  float onep = __fadd_rn(1.f, dens);
  return (float)(log((double)onep) / dg + dg * sigma2 * 0.5);
}

// sample window of a ray with direction cosine uz inside the slab: samples whose plane coordinate r*uz lies in
// [za, zb), widened by one sample each side (the exact per-sample test still decides)
__device__ __forceinline__ void dev_window(double uz, double za, double zb, double dr, int &win_lo, int &win_hi)
{
  double ra, rb;
  if (fabs(uz) < 1e-12) { const bool in = za <= 0 && zb > 0; ra = in ? -1e300 : 1e300; rb = in ? 1e300 : -1e300; }
  else if (uz > 0) { ra = za / uz; rb = zb / uz; }
  else { ra = zb / uz; rb = za / uz; }
  double lo = floor(ra / dr - 0.5) - 1, hi = ceil(rb / dr - 0.5) + 1;
  win_lo = lo < 0 ? 0 : (lo > 2e9 ? 0x7fffffff : (int)lo);
  win_hi = hi < 0 ? -1 : (hi > 2e9 ? 0x7fffffff : (int)hi);
}

// ---- skewers (srcs.c:507-529): one thread per (source, radial sample) --------------------------------------------
template <bool GAUSS>
__global__ void __launch_bounds__(kThreads)
skw_kernel(const ClrDev d, const float *__restrict__ dens, const float *__restrict__ npot, const float4 *__restrict__ pos,
           long long nsrc, int nr, double dr, double sigma2, float *__restrict__ dg, float *__restrict__ vs)
{
  const double idx = (double)(d.n / d.l_box), idr = 1. / dr;
  const bool whole_box = d.nz_here == d.n;
  const long long ngx = d.pitch, plane = ngx * d.n, tot = nsrc * nr;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long is = e / nr;
    const int i_r = (int)(e - is * nr);
    const float4 p = pos[is];
    const double r = dev_src_r(p);
    int i_r_max = (int)(r * idr + 0.5);
    if (i_r_max > nr - 1) i_r_max = nr - 1;
    if (i_r > i_r_max) continue;
    const double ir = 1. / (r > 0.001 ? r : 0.001);
    const double u[3] = {p.x * ir, p.y * ir, p.z * ir};
    const double rm = (i_r + 0.5) * dr;
    double xn[3];
#pragma unroll
    for (int ax = 0; ax < 3; ax++) xn[ax] = (rm * u[ax] + d.pos_obs[ax]) * idx;
    Cic q;
    dev_cic(d, xn, q);
    float dsum = 0.f, v[3] = {0.f, 0.f, 0.f};
    bool added = false;
#pragma unroll
    for (int cz = 0; cz < 2; cz++) {
      const int lz = (cz ? q.z1 : q.z0) - d.iz0_here;
      if (lz < 0 || lz >= d.nz_here) continue;      // the other plane's share is added by the slab that holds it
      added = true;
      const float *w = q.w[cz];
      const int cx[4] = {q.x0, q.x1, q.x0, q.x1}, cy[4] = {q.y0, q.y0, q.y1, q.y1};
      float e4[4], v4[4][3];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float dv = dens[cx[k] + cy[k] * ngx + lz * plane];
        e4[k] = GAUSS ? dev_gauss_element(d, dv, cx[k], cy[k], lz, sigma2) : dv;
        dev_vel_element(d, npot, cx[k], cy[k], lz, whole_box, v4[k]);
      }
      dsum += (e4[0] * w[0] + e4[1] * w[1] + e4[2] * w[2] + e4[3] * w[3]);
#pragma unroll
      for (int ax = 0; ax < 3; ax++) v[ax] += (v4[0][ax] * w[0] + v4[1][ax] * w[1] + v4[2][ax] * w[2] + v4[3][ax] * w[3]);
    }
    if (added) {
      dg[e] = dsum;
      vs[e] = (float)(0.5 * idx * (v[0] * u[0] + v[1] * u[1] + v[2] * u[2]));
    }
  }
}

// Sum of the per-slab partial skewers (several GPUs) + srcs.c:725-733: v_skw *= V1((i+0.5) dr) * factor_vel. The
// reference's loop bound there is MAX((int)(r idr + 0.5), nr - 1): for a source beyond r_max - dr/2 the loop runs on
// into the first elements of the NEXT source's skewer, which so receive the factors of radii (nr + j + 0.5) dr before
// their own (sequential source order). Reproduced: fac has nr + kcap entries.
__global__ void __launch_bounds__(kThreads)
skw_finish_kernel(const ClrDev d, const float *__restrict__ srcs, long long nsrc, int nr, double dr, const double *__restrict__ fac,
                  int kcap, int nparts, long long part_stride, const float *__restrict__ parts_dg, const float *__restrict__ parts_v,
                  float *__restrict__ dg, float *__restrict__ vs)
{
  const double idr = 1. / dr;
  const long long tot = nsrc * nr;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long is = e / nr;
    const int i_r = (int)(e - is * nr);
    float a = vs[e], g = dg[e];
    if (nparts > 0) {
      a = 0.f; g = 0.f;
      for (int k = 0; k < nparts; k++) { a += parts_v[k * part_stride + e]; g += parts_dg[k * part_stride + e]; }
      dg[e] = g;
    }
    if (is > 0 && i_r < kcap) {
      double rp = clr_r_of_z(d, (double)srcs[9 * (is - 1) + 2]);
      int over = (int)(rp * idr + 0.5) - (nr - 1);
      if (over > kcap) over = kcap;
      if (i_r < over) a = (float)((double)a * __ldg(fac + nr + i_r));
    }
    vs[e] = (float)((double)a * __ldg(fac + i_r));
  }
}

// ---- per-source lensing (srcs.c:531-614): one thread per source ----------------------------------------------------
struct LensPlan {
  const double *fac0, *fac1, *fac2;   // D (1+z) dr, r D (1+z) dr, r^2 D (1+z) dr at the sample radii (srcs.c:466-481)
  int nr;
  double dr;
  int restrict_z;
  double za, zb;
};

__global__ void __launch_bounds__(kThreads)
src_lens_kernel(const ClrDev d, const float *__restrict__ npot, const float4 *__restrict__ pos, long long nsrc, LensPlan pl,
                float *__restrict__ out5)
{
  const double idx = (double)(d.n / d.l_box), idr = 1. / pl.dr;
  const bool whole_box = d.nz_here == d.n;
  for (long long ip = blockIdx.x * (long long)blockDim.x + threadIdx.x; ip < nsrc; ip += (long long)gridDim.x * blockDim.x) {
    const float4 p = pos[ip];
    const double r = dev_src_r(p);
    const double ir = 1. / (r > 0.001 ? r : 0.001);
    const double u[3] = {p.x * ir, p.y * ir, p.z * ir};
    double u_x[3], u_y[3], r_k[6], r_e1[6], r_e2[6];
    {
      double cth = u[2], sth, cph = 1, sph = 0;
      const double prefac = idx * idx * ir, prefac_m = 0.5 * idx * ir;
      if (cth >= 1) cth = 1;
      if (cth <= -1) cth = -1;
      sth = sqrt((1 - cth) * (1 + cth));
      if (sth != 0) { cph = u[0] / sth; sph = u[1] / sth; }
      u_x[0] = cth * cph * prefac_m; u_x[1] = cth * sph * prefac_m; u_x[2] = -sth * prefac_m;
      u_y[0] = -sph * prefac_m; u_y[1] = cph * prefac_m; u_y[2] = 0;
      r_k[0] = (cth * cth * cph * cph + sph * sph) * prefac;
      r_k[1] = (2 * cph * sph * (cth * cth - 1)) * prefac;
      r_k[2] = (-2 * cth * sth * cph) * prefac;
      r_k[3] = (cth * cth * sph * sph + cph * cph) * prefac;
      r_k[4] = (-2 * cth * sth * sph) * prefac;
      r_k[5] = (sth * sth) * prefac;
      r_e1[0] = (cth * cth * cph * cph - sph * sph) * prefac;
      r_e1[1] = (2 * cph * sph * (cth * cth + 1)) * prefac;
      r_e1[2] = (-2 * cth * sth * cph) * prefac;
      r_e1[3] = (cth * cth * sph * sph - cph * cph) * prefac;
      r_e1[4] = (-2 * cth * sth * sph) * prefac;
      r_e1[5] = (sth * sth) * prefac;
      r_e2[0] = (-2 * cth * cph * sph) * prefac;
      r_e2[1] = (2 * cth * (cph * cph - sph * sph)) * prefac;
      r_e2[2] = (2 * sth * sph) * prefac;
      r_e2[3] = (2 * cth * sph * cph) * prefac;
      r_e2[4] = (-2 * sth * cph) * prefac;
      r_e2[5] = 0;
    }
    int i_lo = 0, i_hi = (int)(r * idr + 0.5);
    if (i_hi > pl.nr - 1) i_hi = pl.nr - 1;
    if (pl.restrict_z) {
      int wl, wh;
      dev_window(u[2], pl.za, pl.zb, pl.dr, wl, wh);
      i_lo = max(i_lo, wl); i_hi = min(i_hi, wh);
    }
    double dtx = 0, dty = 0, kp = 0, e1 = 0, e2 = 0;
    for (int i_r = i_lo; i_r <= i_hi; i_r++) {
      const double rm = (i_r + 0.5) * pl.dr;
      double xn[3];
      int c[3];
#pragma unroll
      for (int ax = 0; ax < 3; ax++) xn[ax] = (rm * u[ax] + d.pos_obs[ax]) * idx;
      if (!dev_ngp(d, xn, c)) continue;
      float tp[6], vp[3];
      dev_tidal(d, npot, c[0], c[1], c[2], tp);
      dev_vel_element(d, npot, c[0], c[1], c[2], whole_box, vp);
      const double f0 = __ldg(pl.fac0 + i_r), f1 = __ldg(pl.fac1 + i_r), f2 = __ldg(pl.fac2 + i_r);
      const double fr = f1 * r - f2;
      const double frm = 2 * (f0 * r - f1);
      double dotvx = 0, dotvy = 0, dotk = 0, dote1 = 0, dote2 = 0;
#pragma unroll
      for (int ax = 0; ax < 6; ax++) {
        dote1 += r_e1[ax] * tp[ax];
        dote2 += r_e2[ax] * tp[ax];
        dotk += r_k[ax] * tp[ax];
      }
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        dotvx += u_x[ax] * vp[ax];
        dotvy += u_y[ax] * vp[ax];
      }
      e1 += dote1 * fr;
      e2 += dote2 * fr;
      kp += dotk * fr;
      dtx += dotvx * frm;
      dty += dotvy * frm;
    }
    float *o = out5 + 5 * ip;
    o[0] = (float)e1; o[1] = (float)e2; o[2] = (float)kp; o[3] = (float)dty; o[4] = (float)dtx;   // e1, e2, kappa, dra, ddec
  }
}

// srcs.c:609-613 (accumulate the slab contributions into the Src record) + 722-723 (deflections in degrees)
__global__ void __launch_bounds__(kThreads)
lens_finish_kernel(float *__restrict__ srcs, long long nsrc, int nparts, long long part_stride, const float *__restrict__ parts)
{
  for (long long ip = blockIdx.x * (long long)blockDim.x + threadIdx.x; ip < nsrc; ip += (long long)gridDim.x * blockDim.x) {
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nparts; k++)
#pragma unroll
      for (int j = 0; j < 5; j++) acc[j] = (float)((double)acc[j] + (double)parts[k * part_stride + 5 * ip + j]);
    float *o = srcs + 9 * ip;
    o[4] = acc[0]; o[5] = acc[1]; o[6] = acc[2];
    o[7] = (float)((double)acc[3] * kRtod);
    o[8] = (float)((double)acc[4] * kRtod);
  }
}

// ---- custom projected maps (cstm.c:68-145): one thread per pixel ---------------------------------------------------
struct CstmPlan {
  const double *kz, *bz, *normz;   // K(z)/H^-1, b, normalisation at the sample radii (cstm.c:85-95)
  int ir_min, ir_max;              // support of the kernel (cstm.c:97-113)
  double dr;
  int restrict_z;
  double za, zb;
};

__global__ void __launch_bounds__(kThreads)
cstm_kernel(const ClrDev d, const float *__restrict__ dens, const float *__restrict__ dens_halo, const double *__restrict__ pos,
            long long num_pix, CstmPlan pl, float *__restrict__ data)
{
  const double idx = (double)(d.n / d.l_box);
  const bool whole_box = d.nz_here == d.n;
  const long long ngx = d.pitch, plane = ngx * d.n;
  for (long long ip = blockIdx.x * (long long)blockDim.x + threadIdx.x; ip < num_pix; ip += (long long)gridDim.x * blockDim.x) {
    const double u[3] = {pos[3 * ip], pos[3 * ip + 1], pos[3 * ip + 2]};
    int i_lo = pl.ir_min, i_hi = pl.ir_max;
    if (pl.restrict_z) {
      int wl, wh;
      dev_window(u[2], pl.za, pl.zb, pl.dr, wl, wh);
      i_lo = max(i_lo, wl); i_hi = min(i_hi, wh);
    }
    double cval = 0;
    for (int irr = i_lo; irr <= i_hi; irr++) {
      const double rm = (irr + 0.5) * pl.dr;
      double xn[3];
#pragma unroll
      for (int ax = 0; ax < 3; ax++) xn[ax] = (rm * u[ax] + d.pos_obs[ax]) * idx;
      Cic q;
      dev_cic(d, xn, q);
      const float *p0, *p1;
      if (whole_box) {
        p0 = dens + q.z0 * plane; p1 = dens + q.z1 * plane;
      } else {
        const int lz = q.z0 - d.iz0_here;           // the slab of the lower plane owns the sample
        if (lz < 0 || lz >= d.nz_here) continue;
        p0 = dens + lz * plane;
        p1 = lz + 1 < d.nz_here ? p0 + plane : dens_halo;
      }
      const long long c00 = q.x0 + q.y0 * ngx, c01 = q.x1 + q.y0 * ngx, c10 = q.x0 + q.y1 * ngx, c11 = q.x1 + q.y1 * ngx;
      float dv = 0.f;
      dv += (p0[c00] * q.w[0][0] + p0[c01] * q.w[0][1] + p0[c10] * q.w[0][2] + p0[c11] * q.w[0][3]);
      dv += (p1[c00] * q.w[1][0] + p1[c01] * q.w[1][1] + p1[c10] * q.w[1][2] + p1[c11] * q.w[1][3]);
      cval += __ldg(pl.kz + irr) * (clr_bias_model(d.bias_model, (double)dv, __ldg(pl.bz + irr)) * __ldg(pl.normz + irr) - 1);
    }
    data[ip] = (float)((double)data[ip] + cval * pl.dr);
  }
}


// ---- fast-lensing shells (lensing.c:76-250, -D_USE_FAST_LENSING builds): one thread per pixel of the finest shell ----
struct ShellPlan {
  const double *fac0, *fac1, *fac2;     // lensing.c:118-127
  const int *irmin, *irmax;             // sample range of every shell (lensing.c:100-116)
  const double *inv_r_max, *inv_ratio;  // 1 / (i_r_here dr), 1 / npix_ratio
  const long long *ratio, *npp, *off;   // fine pixels per pixel of shell ir, pixels per beam, float offset of shell ir
  int nr_sh;
  long long npix_hi;
  double dr;
  int restrict_z;
  double za, zb;
};

__global__ void __launch_bounds__(kThreads)
lens_shell_kernel(const ClrDev d, const float *__restrict__ npot, const double *__restrict__ pos, long long n_fine, ShellPlan pl,
                  float *__restrict__ data)
{
  const double idx = (double)(d.n / d.l_box);
  const bool whole_box = d.nz_here == d.n;
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < n_fine; it += (long long)gridDim.x * blockDim.x) {
    const long long ib = it / pl.npix_hi, ip = it - ib * pl.npix_hi;
    const double u[3] = {pos[3 * it], pos[3 * it + 1], pos[3 * it + 2]};
    double u_x[3], u_y[3], r_k[6], r_e1[6], r_e2[6];
    {
      double cth = u[2], sth, cph = 1, sph = 0;
      const double prefac = idx * idx, prefac_m = 0.5 * idx;
      if (cth >= 1) cth = 1;
      if (cth <= -1) cth = -1;
      sth = sqrt((1 - cth) * (1 + cth));
      if (sth != 0) { cph = u[0] / sth; sph = u[1] / sth; }
      u_x[0] = cth * cph * prefac_m; u_x[1] = cth * sph * prefac_m; u_x[2] = -sth * prefac_m;
      u_y[0] = -sph * prefac_m; u_y[1] = cph * prefac_m; u_y[2] = 0;
      r_k[0] = (cth * cth * cph * cph + sph * sph) * prefac;
      r_k[1] = (2 * cph * sph * (cth * cth - 1)) * prefac;
      r_k[2] = (-2 * cth * sth * cph) * prefac;
      r_k[3] = (cth * cth * sph * sph + cph * cph) * prefac;
      r_k[4] = (-2 * cth * sth * sph) * prefac;
      r_k[5] = (sth * sth) * prefac;
      r_e1[0] = (cth * cth * cph * cph - sph * sph) * prefac;
      r_e1[1] = (2 * cph * sph * (cth * cth + 1)) * prefac;
      r_e1[2] = (-2 * cth * sth * cph) * prefac;
      r_e1[3] = (cth * cth * sph * sph - cph * cph) * prefac;
      r_e1[4] = (-2 * cth * sth * sph) * prefac;
      r_e1[5] = (sth * sth) * prefac;
      r_e2[0] = (-2 * cth * cph * sph) * prefac;
      r_e2[1] = (2 * cth * (cph * cph - sph * sph)) * prefac;
      r_e2[2] = (2 * sth * sph) * prefac;
      r_e2[3] = (2 * cth * sph * cph) * prefac;
      r_e2[4] = (-2 * sth * cph) * prefac;
      r_e2[5] = 0;
    }
    int win_lo = 0, win_hi = 0x7fffffff;
    if (pl.restrict_z) dev_window(u[2], pl.za, pl.zb, pl.dr, win_lo, win_hi);
    double dx_0 = 0, dx_1 = 0, dy_0 = 0, dy_1 = 0, kappa_1 = 0, kappa_2 = 0, s1_1 = 0, s1_2 = 0, s2_1 = 0, s2_2 = 0;
    for (int ish = 0; ish < pl.nr_sh; ish++) {
      const int irmin = max(__ldg(pl.irmin + ish), win_lo), irmax = min(__ldg(pl.irmax + ish), win_hi);
      for (int irr = irmin; irr <= irmax; irr++) {
        const double rm = (irr + 0.5) * pl.dr;
        double xn[3];
        int c[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) xn[ax] = (rm * u[ax] + d.pos_obs[ax]) * idx;
        if (!dev_ngp(d, xn, c)) continue;
        float t[6], v[3];
        dev_tidal(d, npot, c[0], c[1], c[2], t);
        dev_vel_element(d, npot, c[0], c[1], c[2], whole_box, v);
        double dotk = 0, dote1 = 0, dote2 = 0, dotvx = 0, dotvy = 0;
#pragma unroll
        for (int ax = 0; ax < 6; ax++) { dote1 += r_e1[ax] * t[ax]; dote2 += r_e2[ax] * t[ax]; dotk += r_k[ax] * t[ax]; }
#pragma unroll
        for (int ax = 0; ax < 3; ax++) { dotvx += u_x[ax] * v[ax]; dotvy += u_y[ax] * v[ax]; }
        const double f0 = __ldg(pl.fac0 + irr), f1 = __ldg(pl.fac1 + irr), f2 = __ldg(pl.fac2 + irr);
        dx_0 += dotvx * f0; dx_1 += dotvx * f1;
        dy_0 += dotvy * f0; dy_1 += dotvy * f1;
        kappa_1 += dotk * f1; kappa_2 += dotk * f2;
        s1_1 += dote1 * f1; s1_2 += dote1 * f2;
        s2_1 += dote2 * f1; s2_2 += dote2 * f2;
      }
      // several fine pixels share a pixel of a coarser shell (lensing.c:219-227): atomic float sums
      const double irm = __ldg(pl.inv_r_max + ish), w = __ldg(pl.inv_ratio + ish);
      const long long npp = __ldg(pl.npp + ish);
      float *o = data + __ldg(pl.off + ish) + ib * 5 * npp + 5 * (ip / __ldg(pl.ratio + ish));
      atomicAdd(o + 0, (float)((s1_1 - irm * s1_2) * w));
      atomicAdd(o + 1, (float)((s2_1 - irm * s2_2) * w));
      atomicAdd(o + 2, (float)((kappa_1 - irm * kappa_2) * w));
      atomicAdd(o + 3, (float)(2 * (dx_0 - irm * dx_1) * w));
      atomicAdd(o + 4, (float)(2 * (dy_0 - irm * dy_1) * w));
    }
  }
}

// srcs.c:666-723: shear / convergence / deflection of the sources interpolated between the two shells that bracket them
struct ShellLookup {
  const float *r_sh;          // snapped shell radii (lensing.c:233-236)
  const int *nside_sh;
  const long long *npp, *off;
  int nr_sh, nbeams, node, nnodes;
};
__device__ __forceinline__ long long dev_vec2pix_nest(int nside, double x, double y, double z)
{
  const double vlen = sqrt(x * x + y * y + z * z);
  return clr_ring2nest(nside, (int)clr_ang2pix_ring_zphi(nside, z / vlen, atan2(y, x)));
}
__global__ void __launch_bounds__(kThreads)
src_shell_lens_kernel(const ClrDev d, const float4 *__restrict__ pos, float *__restrict__ srcs, long long nsrc, ShellLookup L,
                      const float *__restrict__ data, unsigned long long *__restrict__ bad)
{
  for (long long ii = blockIdx.x * (long long)blockDim.x + threadIdx.x; ii < nsrc; ii += (long long)gridDim.x * blockDim.x) {
    float *o = srcs + 9 * ii;
    const double r = clr_r_of_z(d, (double)o[2]);
    // get_r_index_lensing (srcs.c:24-64): r_sh[i] <= r < r_sh[i+1], clamped to [0, nr-2]
    int ir = 0;
    while (ir < L.nr_sh - 2 && r >= (double)__ldg(L.r_sh + ir + 1)) ir++;
    const float ra = __ldg(L.r_sh + ir), rb = __ldg(L.r_sh + ir + 1);
    const double h = (r - (double)ra) / (double)(rb - ra);
    const float4 p = pos[ii];
    const long long ibase = dev_vec2pix_nest(d.nside_base, p.x, p.y, p.z);
    const long long npl = __ldg(L.npp + ir), npu = __ldg(L.npp + ir + 1);
    const long long ipl = dev_vec2pix_nest(__ldg(L.nside_sh + ir), p.x, p.y, p.z) - ibase * npl;
    const long long ipu = dev_vec2pix_nest(__ldg(L.nside_sh + ir + 1), p.x, p.y, p.z) - ibase * npu;
    if (ibase % L.nnodes != L.node || ipl < 0 || ipl >= npl || ipu < 0 || ipu >= npu) { atomicAdd(bad, 1ULL); continue; }
    const long long ibh = (ibase - L.node) / L.nnodes;
    const float *lo = data + __ldg(L.off + ir) + ibh * 5 * npl + 5 * ipl;
    const float *up = data + __ldg(L.off + ir + 1) + ibh * 5 * npu + 2 * ipu;    // stride 2, not 5: srcs.c:710-714 as is
    const double g1 = (double)lo[0] * (1 - h) + (double)up[0] * h, g2 = (double)lo[1] * (1 - h) + (double)up[1] * h;
    const double kp = (double)lo[2] * (1 - h) + (double)up[2] * h;
    const double dxv = (double)lo[3] * (1 - h) + (double)up[3] * h, dyv = (double)lo[4] * (1 - h) + (double)up[4] * h;
    o[4] = (float)g1; o[5] = (float)g2; o[6] = (float)kp;
    o[7] = (float)((double)(float)dyv * kRtod);
    o[8] = (float)((double)(float)dxv * kRtod);
  }
}

double host_lerp(const clr_ctx *c, double r, const std::vector<double> &f, double f0, double ff)
{
  if (r <= 0) return f0;
  else if (r >= c->h_r[CLR_NA - 1]) return ff;
  int ir = (int)(r * c->p.glob_idr);
  return f[ir] + (f[ir + 1] - f[ir]) * (r - c->h_r[ir]) * c->p.glob_idr;
}

// slab window [za, zb) of the plane coordinate r*u_z for NGP (half = 0.5) or floor (half = 0) plane assignment;
// only valid when no sample of any ray wraps around the box
bool slab_window(const clr_ctx *c, int nr, double dr, double half, double *za, double *zb)
{
  if (c->nranks <= 1) return false;
  const double idx = (double)(c->p.n_grid / c->p.l_box);
  double far = (nr * dr + fabs(c->p.pos_obs[2])) * idx + 1.0, near = (c->p.pos_obs[2] - nr * dr) * idx;
  if (!(far < c->p.n_grid && near >= 0)) return false;
  *za = (c->dev.iz0_here - half) / idx - c->p.pos_obs[2];
  *zb = (c->dev.iz0_here + c->dev.nz_here - half) / idx - c->p.pos_obs[2];
  return true;
}

struct DevBuf {     // device allocation freed on scope exit
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T *as() { return static_cast<T *>(p); }
};

}  // namespace

// cstm_beams_preproc + cstm_get_beam_properties (cstm.c:38-145) for the pixels `pos` (unit vectors)
int clr_beam_cstm(clr_ctx *c, int ipop, long long num_pix, const double *h_pos, float *h_data)
{
  clr_ctx::Pop &P = c->cstm[ipop];
  CLR_CHECK(P.set, "custom population %d not set", ipop);
  CLR_CHECK(P.have_norm, "custom population %d has no normalisation", ipop);
  CLR_CHECK(num_pix > 0, "custom map: no pixels");
  const int nr = c->p.n_grid / 2;                 // get_radial_params (common.c:333-337)
  const double dr = c->p.r_max / nr;
  std::vector<double> tab(3 * (size_t)nr);
  double *kz = tab.data(), *bz = kz + nr, *normz = bz + nr, k_max = -1E100;
  for (int ir = 0; ir < nr; ir++) {               // cstm.c:85-95
    double rm = (ir + 0.5) * dr;
    kz[ir] = host_lerp(c, rm, P.h_a, 0, 0) / host_lerp(c, rm, c->h_ih, c->h_ih[0], c->h_ih[CLR_NA - 1]);
    bz[ir] = host_lerp(c, rm, P.h_b, P.h_b[0], 1);
    normz[ir] = host_lerp(c, rm, P.h_norm, P.norm_0, P.norm_f);
    if (fabs(kz[ir]) > k_max) k_max = fabs(kz[ir]);
  }
  int ir_min = 0, ir_max = nr - 1;                // cstm.c:97-113
  for (int ir = 0; ir < nr; ir++) if (fabs(kz[ir]) > 1E-4 * k_max) { ir_min = ir; break; }
  for (int ir = nr - 1; ir >= 0; ir--) if (fabs(kz[ir]) > 1E-4 * k_max) { ir_max = ir; break; }
  CLR_CHECK(ir_max >= ir_min, "Custom kernel has no suppport");
  DevBuf b_pos, b_tab, b_data, b_halo;
  CLR_CUDA(cudaMalloc(&b_pos.p, (size_t)3 * num_pix * sizeof(double)));
  CLR_CUDA(cudaMalloc(&b_tab.p, tab.size() * sizeof(double)));
  CLR_CUDA(cudaMalloc(&b_data.p, (size_t)num_pix * sizeof(float)));
  CLR_CUDA(cudaMemcpyAsync(b_pos.p, h_pos, (size_t)3 * num_pix * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(b_tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemsetAsync(b_data.p, 0, (size_t)num_pix * sizeof(float), c->stream));
  if (c->nranks > 1) {
    CLR_CUDA(cudaMalloc(&b_halo.p, (size_t)c->dev.pitch * c->dev.n * sizeof(float)));
    if (clr_comm_dens_halo(c, b_halo.as<float>())) return 1;
  }
  CstmPlan pl{b_tab.as<double>(), b_tab.as<double>() + nr, b_tab.as<double>() + 2 * nr, ir_min, ir_max, dr, 0, 0., 0.};
  pl.restrict_z = slab_window(c, nr, dr, 0.0, &pl.za, &pl.zb) ? 1 : 0;
  {
    StageScope sc(c, "cstm_los", 1);
    cstm_kernel<<<blocks_for(c, num_pix, 8), kThreads, 0, c->stream>>>(c->dev, c->d_dens, b_halo.as<float>(), b_pos.as<double>(),
                                                                        num_pix, pl, b_data.as<float>());
    CLR_CUDA(cudaGetLastError());
  }
  if (clr_comm_allreduce_f32(c, b_data.as<float>(), (size_t)num_pix)) return 1;
  CLR_CUDA(cudaMemcpyAsync(h_data, b_data.p, (size_t)num_pix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}


// lensing_beams_preproc + lensing_get_beam_properties (lensing.c:39-250). r_sh: shell radii, sorted ascending on entry,
// snapped to the radial sampling on exit. npp[ir]: pixels per beam of shell ir; pos: unit vectors of the finest shell,
// [nbeams][npp[nr-1]][3]. data (host, may be NULL): shells concatenated, shell ir = [nbeams][5 * npp[ir]]. The result
// also stays on the device for clr_beam_srcs_from_shells.
int clr_beam_lens_shells(clr_ctx *c, int nbeams, int nr_sh, float *r_sh, const long long *npp, const double *h_pos, float *h_data)
{
  CLR_CHECK(nbeams > 0 && nr_sh >= 2 && nr_sh <= 4096, "lensing shells: bad beam / shell count");
  const int nr = c->p.n_grid / 2;
  const double dr = c->p.r_max / nr, idr = 1. / dr;
  const long long npix_hi = npp[nr_sh - 1], n_fine = (long long)nbeams * npix_hi;
  std::vector<int> ir(2 * (size_t)nr_sh);
  std::vector<double> dd(2 * (size_t)nr_sh + 3 * (size_t)nr);
  std::vector<long long> ll(3 * (size_t)nr_sh);
  int *irmin = ir.data(), *irmax = irmin + nr_sh;
  double *inv_r_max = dd.data(), *inv_ratio = inv_r_max + nr_sh, *fac = inv_ratio + nr_sh;
  long long *ratio = ll.data(), *nppv = ratio + nr_sh, *off = nppv + nr_sh, total = 0;
  for (int i = 0; i < nr_sh; i++) {                   // lensing.c:100-116
    CLR_CHECK(npp[i] > 0 && npix_hi % npp[i] == 0, "lensing shells: shell %d does not nest into the finest one", i);
    int i_r_here = (int)(r_sh[i] * idr + 0.5);
    inv_r_max[i] = 1. / (i_r_here * dr);
    irmax[i] = std::min(i_r_here, nr - 1);
    ratio[i] = npix_hi / npp[i];
    inv_ratio[i] = 1. / ((double)ratio[i]);
    nppv[i] = npp[i];
    off[i] = total;
    total += 5LL * nbeams * npp[i];
  }
  irmin[0] = 0;
  for (int i = 1; i < nr_sh; i++) irmin[i] = irmax[i - 1] + 1;
  for (int i = 0; i < nr; i++) {                      // lensing.c:118-127
    double rm = (i + 0.5) * dr;
    double pg = host_lerp(c, rm, c->h_d1, 1, c->h_d1[CLR_NA - 1]) * (1 + host_lerp(c, rm, c->h_z, 0, c->h_z[CLR_NA - 1]));
    fac[i] = pg * dr; fac[nr + i] = rm * pg * dr; fac[2 * nr + i] = rm * rm * pg * dr;
  }
  DevBuf b_i, b_d, b_l, b_pos;
  CLR_CUDA(cudaMalloc(&b_i.p, ir.size() * sizeof(int)));
  CLR_CUDA(cudaMalloc(&b_d.p, dd.size() * sizeof(double)));
  CLR_CUDA(cudaMalloc(&b_l.p, ll.size() * sizeof(long long)));
  CLR_CUDA(cudaMalloc(&b_pos.p, (size_t)3 * n_fine * sizeof(double)));
  cudaFree(c->d_lens_data); c->d_lens_data = nullptr; c->lens_total = 0;
  CLR_CUDA(cudaMalloc(&c->d_lens_data, (size_t)total * sizeof(float)));
  c->lens_total = total; c->lens_nbeams = nbeams;
  c->lens_npp.assign(npp, npp + nr_sh);
  CLR_CUDA(cudaMemcpyAsync(b_i.p, ir.data(), ir.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(b_d.p, dd.data(), dd.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(b_l.p, ll.data(), ll.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(b_pos.p, h_pos, (size_t)3 * n_fine * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemsetAsync(c->d_lens_data, 0, (size_t)total * sizeof(float), c->stream));
  ShellPlan pl;
  pl.fac0 = b_d.as<double>() + 2 * nr_sh; pl.fac1 = pl.fac0 + nr; pl.fac2 = pl.fac1 + nr;
  pl.irmin = b_i.as<int>(); pl.irmax = pl.irmin + nr_sh;
  pl.inv_r_max = b_d.as<double>(); pl.inv_ratio = pl.inv_r_max + nr_sh;
  pl.ratio = b_l.as<long long>(); pl.npp = pl.ratio + nr_sh; pl.off = pl.npp + nr_sh;
  pl.nr_sh = nr_sh; pl.npix_hi = npix_hi; pl.dr = dr;
  pl.restrict_z = slab_window(c, nr, dr, 0.5, &pl.za, &pl.zb) ? 1 : 0;
  {
    StageScope sc(c, "lensing_shells", 1);
    lens_shell_kernel<<<blocks_for(c, n_fine, 8), kThreads, 0, c->stream>>>(c->dev, c->d_npot, b_pos.as<double>(), n_fine, pl, c->d_lens_data);
    CLR_CUDA(cudaGetLastError());
  }
  if (clr_comm_allreduce_f32(c, c->d_lens_data, (size_t)total)) return 1;
  if (h_data) CLR_CUDA(cudaMemcpyAsync(h_data, c->d_lens_data, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nr_sh; i++) r_sh[i] = (float)(1. / inv_r_max[i]);     // lensing.c:233-236
  return 0;
}

// the lensing part of srcs_beams_postproc under _USE_FAST_LENSING (srcs.c:666-723) from the shells of the last
// clr_beam_lens_shells; beam ib holds base pixel ib * nnodes + node. *n_bad: sources outside the held base pixels.
int clr_beam_srcs_from_shells(clr_ctx *c, int ipop, int nr_sh, const float *r_sh, const int *nside_sh, int node, int nnodes,
                              long long *n_bad)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  CLR_CHECK(c->d_lens_data && (int)c->lens_npp.size() == nr_sh, "no lensing shells on the device (clr_lensing_get_beam_properties first)");
  CLR_CHECK(nnodes >= 1 && node >= 0 && node < nnodes, "bad node layout");
  if (n_bad) *n_bad = 0;
  if (P.nsrc == 0) return 0;
  std::vector<long long> ll(2 * (size_t)nr_sh);
  long long total = 0;
  for (int i = 0; i < nr_sh; i++) { ll[i] = c->lens_npp[i]; ll[nr_sh + i] = total; total += 5LL * c->lens_nbeams * c->lens_npp[i]; }
  DevBuf b_r, b_ns, b_l;
  CLR_CUDA(cudaMalloc(&b_r.p, nr_sh * sizeof(float)));
  CLR_CUDA(cudaMalloc(&b_ns.p, nr_sh * sizeof(int)));
  CLR_CUDA(cudaMalloc(&b_l.p, ll.size() * sizeof(long long)));
  CLR_CUDA(cudaMemcpyAsync(b_r.p, r_sh, nr_sh * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(b_ns.p, nside_sh, nr_sh * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(b_l.p, ll.data(), ll.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  if (clr_ensure_scratch(c, sizeof(unsigned long long))) return 1;
  unsigned long long *d_bad = reinterpret_cast<unsigned long long *>(c->d_scratch), h_bad = 0;
  CLR_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
  ShellLookup L{b_r.as<float>(), b_ns.as<int>(), b_l.as<long long>(), b_l.as<long long>() + nr_sh, nr_sh, c->lens_nbeams, node, nnodes};
  {
    StageScope sc(c, "srcs_shell_lensing", 1);
    src_shell_lens_kernel<<<blocks_for(c, P.nsrc, 8), kThreads, 0, c->stream>>>(c->dev, reinterpret_cast<const float4 *>(P.d_pos), P.d_srcs,
                                                                                 P.nsrc, L, c->d_lens_data, d_bad);
    CLR_CUDA(cudaGetLastError());
  }
  if (clr_read_small(c, &h_bad, d_bad, sizeof(h_bad))) return 1;
  if (n_bad) *n_bad = (long long)h_bad;
  return 0;
}

// srcs_beams_preproc + srcs_get_beam_properties + srcs_beams_postproc (srcs.c:425-744) for one population: RSD under
// beaming (clr_srcs_beam), then the skewers and the per-source lensing of the default build
int clr_beam_srcs(clr_ctx *c, int ipop, int has_lensing, int has_skw, int skw_gauss, int rsd_done)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  CLR_CHECK(!(has_skw && skw_gauss) || c->p.dens_type == CLR_DENS_TYPE_LGNR,
            "Cannot write Gaussian skewers with density type %d\n", c->p.dens_type);      // beaming.c:56-57
  // zeroes e1, e2 like srcs.c:425-443 and finishes dz_rsd; rsd_done: the sources were routed by pixel and carry the
  // estimator evaluated on their home slab (clr_srcs_distribute with beam_first)
  if (!rsd_done && clr_srcs_beam(c, ipop)) return 1;
  if (!has_lensing && !has_skw) return 0;
  const int R = c->nranks, nr = c->p.n_grid / 2;     // catalog_alloc (common.c:366-397): nr, dr of the skewers
  const double dr = c->p.r_max / nr;
  const long long n_own = P.nsrc;
  // ---- several GPUs: every rank needs the positions of all sources
  std::vector<unsigned long long> cnt(R, 0ULL);
  long long n_all = n_own, my_off = 0;
  DevBuf b_allpos;
  const float4 *d_pos_all = reinterpret_cast<const float4 *>(P.d_pos);
  if (R > 1) {
    if (clr_ensure_scratch(c, (size_t)R * sizeof(unsigned long long))) return 1;
    unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(c->d_scratch);
    cnt[c->rank] = (unsigned long long)n_own;
    CLR_CUDA(cudaMemcpyAsync(d_cnt, cnt.data(), R * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    if (clr_comm_allreduce_u64(c, d_cnt, R)) return 1;
    if (clr_read_small(c, cnt.data(), d_cnt, R * sizeof(unsigned long long))) return 1;
    n_all = 0;
    std::vector<size_t> off(R), num(R), s_off(R, 0), s_num(R);
    for (int k = 0; k < R; k++) { off[k] = (size_t)n_all * 4; num[k] = (size_t)cnt[k] * 4; s_num[k] = (size_t)n_own * 4; n_all += (long long)cnt[k]; }
    my_off = (long long)(off[c->rank] / 4);
    if (n_all > 0) {
      int ok = cudaMalloc(&b_allpos.p, (size_t)n_all * sizeof(float4)) == cudaSuccess;
      if (!ok) clr_set_error("srcs beams: out of device memory (positions of all sources)");
      if (clr_comm_all_ok(c, ok, "srcs beams: position buffer")) return 1;
      if (n_own > 0)
        CLR_CUDA(cudaMemcpyAsync(b_allpos.as<float4>() + my_off, P.d_pos, (size_t)n_own * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
      if (clr_comm_alltoallv(c, P.d_pos, s_off.data(), s_num.data(), b_allpos.as<float>(), off.data(), num.data())) return 1;
    }
    d_pos_all = b_allpos.as<float4>();
  }
  if (n_all == 0) return 0;
  // exchange of per-source partial results: block of rank k's sources goes to rank k; returns the R partials of the
  // own sources in `recv` ([R][n_own * width])
  auto exchange = [&](const float *part_all, int width, float *recv) -> int {
    std::vector<size_t> s_off(R), s_num(R), r_off(R), r_num(R);
    size_t o = 0;
    for (int k = 0; k < R; k++) {
      s_off[k] = o * width; s_num[k] = (size_t)cnt[k] * width; o += (size_t)cnt[k];
      r_off[k] = (size_t)k * n_own * width; r_num[k] = (size_t)n_own * width;
    }
    if (n_own > 0)
      CLR_CUDA(cudaMemcpyAsync(recv + r_off[c->rank], part_all + s_off[c->rank], r_num[c->rank] * sizeof(float),
                               cudaMemcpyDeviceToDevice, c->stream));
    return clr_comm_alltoallv(c, part_all, s_off.data(), s_num.data(), recv, r_off.data(), r_num.data());
  };
  if (has_lensing) {
    std::vector<double> fac(3 * (size_t)nr);
    for (int i = 0; i < nr; i++) {                   // srcs.c:466-481
      double rm = (i + 0.5) * dr;
      double pg = host_lerp(c, rm, c->h_d1, 1, c->h_d1[CLR_NA - 1]) * (1 + host_lerp(c, rm, c->h_z, 0, c->h_z[CLR_NA - 1]));
      fac[i] = pg * dr;
      fac[nr + i] = rm * pg * dr;
      fac[2 * nr + i] = rm * rm * pg * dr;
    }
    DevBuf b_fac, b_part, b_recv;
    CLR_CUDA(cudaMalloc(&b_fac.p, fac.size() * sizeof(double)));
    CLR_CUDA(cudaMemcpyAsync(b_fac.p, fac.data(), fac.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int ok = cudaMalloc(&b_part.p, (size_t)n_all * 5 * sizeof(float)) == cudaSuccess;
    if (ok && R > 1 && n_own > 0) ok = cudaMalloc(&b_recv.p, (size_t)R * n_own * 5 * sizeof(float)) == cudaSuccess;
    if (!ok) clr_set_error("per-source lensing: out of device memory");
    if (clr_comm_all_ok(c, ok, "per-source lensing: partial-result buffers")) return 1;
    LensPlan pl{b_fac.as<double>(), b_fac.as<double>() + nr, b_fac.as<double>() + 2 * nr, nr, dr, 0, 0., 0.};
    pl.restrict_z = slab_window(c, nr, dr, 0.5, &pl.za, &pl.zb) ? 1 : 0;
    {
      StageScope sc(c, "srcs_lensing", 2);
      src_lens_kernel<<<blocks_for(c, n_all, 8), kThreads, 0, c->stream>>>(c->dev, c->d_npot, d_pos_all, n_all, pl, b_part.as<float>());
      CLR_CUDA(cudaGetLastError());
      const float *parts = b_part.as<float>();
      int nparts = 1;
      if (R > 1) {
        if (exchange(b_part.as<float>(), 5, b_recv.as<float>())) return 1;
        parts = b_recv.as<float>(); nparts = R;
      }
      if (n_own > 0) {
        lens_finish_kernel<<<blocks_for(c, n_own, 8), kThreads, 0, c->stream>>>(P.d_srcs, n_own, nparts, n_own * 5, parts);
        CLR_CUDA(cudaGetLastError());
      }
    }
    CLR_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (has_skw) {
    // fac[i] = V1((i+0.5) dr) * factor_vel (srcs.c:641, 730-731), extended past nr for the overrun of srcs.c:726
    const double idr = 1. / dr, dx = c->p.l_box / c->p.n_grid;
    int kcap = (int)((0.5 * c->p.l_box + 20. + 2 * dx) * idr + 0.5) - (nr - 1) + 2;
    if (kcap < 1) kcap = 1;
    if (kcap > nr) kcap = nr;
    const double factor_vel = -c->p.fgrowth_0 / (1.5 * c->p.hubble_0 * c->p.OmegaM);
    std::vector<double> fac((size_t)nr + kcap);
    for (int i = 0; i < nr + kcap; i++)
      fac[i] = host_lerp(c, (i + 0.5) * dr, c->h_v1, c->h_v1[0], c->h_v1[CLR_NA - 1]) * factor_vel;
    DevBuf b_fac, b_pdg, b_pv, b_rdg, b_rv;
    CLR_CUDA(cudaMalloc(&b_fac.p, fac.size() * sizeof(double)));
    CLR_CUDA(cudaMemcpyAsync(b_fac.p, fac.data(), fac.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const size_t own_el = (size_t)n_own * nr, all_el = (size_t)n_all * nr;
    cudaFree(P.d_skw_dg); cudaFree(P.d_skw_v);
    P.d_skw_dg = P.d_skw_v = nullptr; P.skw_n = 0; P.skw_nr = nr;
    int ok = 1;
    if (n_own > 0) ok = cudaMalloc(&P.d_skw_dg, own_el * sizeof(float)) == cudaSuccess && cudaMalloc(&P.d_skw_v, own_el * sizeof(float)) == cudaSuccess;
    if (ok && R > 1) {
      ok = cudaMalloc(&b_pdg.p, all_el * sizeof(float)) == cudaSuccess && cudaMalloc(&b_pv.p, all_el * sizeof(float)) == cudaSuccess;
      if (ok && n_own > 0)
        ok = cudaMalloc(&b_rdg.p, (size_t)R * own_el * sizeof(float)) == cudaSuccess && cudaMalloc(&b_rv.p, (size_t)R * own_el * sizeof(float)) == cudaSuccess;
    }
    if (!ok) clr_set_error("skewers: out of device memory (2 x nsrc x n_grid/2 floats)");
    if (clr_comm_all_ok(c, ok, "skewers: buffers (nsrc * n_grid/2 floats each)")) return 1;
    float *w_dg = R > 1 ? b_pdg.as<float>() : P.d_skw_dg, *w_v = R > 1 ? b_pv.as<float>() : P.d_skw_v;
    CLR_CUDA(cudaMemsetAsync(w_dg, 0, all_el * sizeof(float), c->stream));
    CLR_CUDA(cudaMemsetAsync(w_v, 0, all_el * sizeof(float), c->stream));
    {
      StageScope sc(c, "srcs_skewers", 2);
      const int grid = blocks_for(c, (long long)all_el, 16);
      if (skw_gauss)
        skw_kernel<true><<<grid, kThreads, 0, c->stream>>>(c->dev, c->d_dens, c->d_npot, d_pos_all, n_all, nr, dr, c->sigma2_gauss, w_dg, w_v);
      else
        skw_kernel<false><<<grid, kThreads, 0, c->stream>>>(c->dev, c->d_dens, c->d_npot, d_pos_all, n_all, nr, dr, c->sigma2_gauss, w_dg, w_v);
      CLR_CUDA(cudaGetLastError());
      int nparts = 0;
      if (R > 1) {
        if (exchange(b_pdg.as<float>(), nr, b_rdg.as<float>())) return 1;
        if (exchange(b_pv.as<float>(), nr, b_rv.as<float>())) return 1;
        nparts = R;
      }
      if (n_own > 0) {
        skw_finish_kernel<<<blocks_for(c, (long long)own_el, 16), kThreads, 0, c->stream>>>(
            c->dev, P.d_srcs, n_own, nr, dr, b_fac.as<double>(), kcap, nparts, (long long)own_el, b_rdg.as<float>(), b_rv.as<float>(),
            P.d_skw_dg, P.d_skw_v);
        CLR_CUDA(cudaGetLastError());
      }
    }
    CLR_CUDA(cudaStreamSynchronize(c->stream));
    P.skw_n = n_own;
  }
  return 0;
}

int clr_beam_get_skewers(clr_ctx *c, int ipop, float *h_dg, float *h_v)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  CLR_CHECK(P.skw_n == P.nsrc, "no skewers for population %d (call clr_srcs_get_beam_properties with has_skw first)", ipop);
  if (P.skw_n == 0) return 0;
  const size_t bytes = (size_t)P.skw_n * P.skw_nr * sizeof(float);
  if (h_dg) CLR_CUDA(cudaMemcpyAsync(h_dg, P.d_skw_dg, bytes, cudaMemcpyDeviceToHost, c->stream));
  if (h_v) CLR_CUDA(cudaMemcpyAsync(h_v, P.d_skw_v, bytes, cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
