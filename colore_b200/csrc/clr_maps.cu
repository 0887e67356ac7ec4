// Map kernels: intensity-map painting (scatter with atomics) and the kappa / ISW line-of-sight
// integrals (gather with finite-difference stencils).
// Replaces imap_set_cartesian_single (imap.c:135-245), kappa_get_beam_properties (kappa.c:78-175),
// isw_get_beam_properties (isw.c:78-147) and the NGP branch of interpolate_from_grid /
// get_element (beaming.c:31-181). Compiled with -fmad=false (pixel indices must match the oracle).
//
// Multi-GPU note: every accumulator is linear in the field, so each GPU integrates the part of
// every ray that crosses ITS slab (interpolate_from_grid returns added=0 elsewhere, beaming.c:159)
// and the maps are summed afterwards; the reference's ring rotation of whole slabs
// (beaming.c:325-352) is not needed.
#include "clr_internal.cuh"
#include "clr_stencil.cuh"
#include <math.h>
#include <algorithm>

namespace {

constexpr int kThreads = 256;

struct ImapShells { const float *r0, *rf; const int *nsub; int nr, nsub_lo, nsub_hi; };

// imap.c:76-103 for contiguous sorted shells: index i with r0[i] <= r < rf[i]; -1 below the first
// shell, nr above the last. (The reference's cached linear search returns the same index; it never
// terminates if r falls in a gap between shells, which contiguous frequency tables never produce:
// such points are skipped here, return -2.)
__device__ __forceinline__ int dev_r_index(const ImapShells &sh, double r)
{
  if (r < (double)__ldg(sh.r0)) return -1;
  int lo = 0, hi = sh.nr - 1;          // invariant: r0[lo] <= r
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if ((double)__ldg(sh.r0 + mid) <= r) lo = mid; else hi = mid - 1;
  }
  if (r < (double)__ldg(sh.rf + lo)) return lo;
  return lo == sh.nr - 1 ? sh.nr : -2;
}

__device__ __forceinline__ double dev_get_rvel(const ClrDev &d, const float *__restrict__ npot, int ix, int iy, int iz,
                                               double x0, double y0, double z0, double rr)
{
  const double idx = (double)(d.n / d.l_box);
  const long long ngx = d.pitch, plane = ngx * d.n;
  int ix_hi = ix + 1 == d.n ? 0 : ix + 1, ix_lo = ix == 0 ? d.n - 1 : ix - 1;
  int iy_hi = iy + 1 == d.n ? 0 : iy + 1, iy_lo = iy == 0 ? d.n - 1 : iy - 1;
  long long pz_hi = (iz == d.nz_here - 1) ? (long long)(d.nz_here + 1) : iz + 1;
  long long pz_lo = (iz == 0) ? (long long)d.nz_here : iz - 1;
  double u0 = x0 / rr, u1 = y0 / rr, u2 = z0 / rr;
  float v0 = npot[ix_hi + iy * ngx + iz * plane] - npot[ix_lo + iy * ngx + iz * plane];
  float v1 = npot[ix + iy_hi * ngx + iz * plane] - npot[ix + iy_lo * ngx + iz * plane];
  float v2 = npot[ix + iy * ngx + pz_hi * plane] - npot[ix + iy * ngx + pz_lo * plane];
  return 0.5 * idx * (v0 * u0 + v1 * u1 + v2 * u2);
}

// one sub-cell, the reference's double arithmetic verbatim (imap.c:214-231)
__device__ __forceinline__ void imap_subcell_exact(const ImapShells &sh, int nside, long long num_pix, double x, double y,
                                                   double z, double dr_rsd, float temp, float *__restrict__ data,
                                                   int *__restrict__ nadd)
{
  double r, cth, phi;
  clr_cart2sph(x, y, z, &r, &cth, &phi);
  int ir = dev_r_index(sh, r + dr_rsd);
  if (ir >= 0 && ir < sh.nr) {
    long long pix = clr_ang2pix_ring_zphi(nside, cth, phi);
    atomicAdd(&data[ir * num_pix + pix], temp);
    atomicAdd(&nadd[ir * num_pix + pix], 1);
  }
}

// one thread per cell; sub-cells are painted with float / int atomics (imap.c:224-231)
__global__ void __launch_bounds__(kThreads)
imap_kernel(const ClrDev d, const float *__restrict__ dens, const float *__restrict__ npot, ClrPop pop, ImapShells sh,
            int nside, float *__restrict__ data, int *__restrict__ nadd, double rmin_here, double rmax_here)
{
  const long long n_cells = (long long)d.nz_here * d.n * d.n;
  const long long num_pix = 12LL * nside * nside;
  const double dx = (double)(d.l_box / d.n);
  const double factor_vel = -d.fgrowth_0 / (1.5 * d.hubble_0 * d.OmegaM);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    clr_cell(d, i, ix, iy, iz);
    long long row = (long long)iz * d.n + iy;
    double z0 = __ldg(d.cd[2] + iz + d.iz0_here);
    double y0 = __ldg(d.cd[1] + iy);
    double x0 = __ldg(d.cd[0] + ix);
    double r0 = sqrt(x0 * x0 + y0 * y0 + z0 * z0);
    if (!(r0 <= rmax_here && r0 >= rmin_here)) continue;
    double tmean = clr_lerp(d, r0, pop.nz, 0.0, 0.0);
    if (!(tmean > 0)) continue;
    double bias = clr_bg_bz(d, r0, pop.bz);
    double dnorm = clr_lerp(d, r0, pop.norm, pop.norm_0, pop.norm_f);
    double rvel = factor_vel * dev_get_rvel(d, npot, ix, iy, iz, x0, y0, z0, r0);
    double dr_rsd = rvel * clr_bg_v1(d, r0) * clr_bg_ih(d, r0);
    float temp = (float)(tmean * clr_bias_model(d.bias_model, (double)dens[row * d.pitch + ix], bias) * dnorm);
    int irad = dev_r_index(sh, r0);
    int nsub = irad < 0 ? sh.nsub_lo : (irad >= sh.nr ? sh.nsub_hi : __ldg(sh.nsub + irad));
    double dx_sub = dx / nsub;
    for (int izz = 0; izz < nsub; izz++) {
      double z = z0 + (izz + 0.5) * dx_sub;
      for (int iyy = 0; iyy < nsub; iyy++) {
        double y = y0 + (iyy + 0.5) * dx_sub;
        for (int ixx = 0; ixx < nsub; ixx++) {
          double x = x0 + (ixx + 0.5) * dx_sub;
          imap_subcell_exact(sh, nside, num_pix, x, y, z, dr_rsd, temp, data, nadd);
        }
      }
    }
  }
}

// ---- fp32-screened painter (default path) ---------------------------------------------------------
// The per-cell quantities (temperature, RSD shift, sub-sampling) keep the reference's double arithmetic; the
// per-SUB-CELL work -- 4*10^9 (r, cos theta, phi) -> (shell, pixel) evaluations at 1024^3 -- runs in fp32 with
// rigorous margins: whenever r + dr_rsd lies within kMarginR of a shell edge, or one of the floor() arguments
// of ang2pix_ring lies within `margin` of an integer (fp32 error budget: 2e-6 * nside), the sub-cell is queued
// in shared memory and the CTA re-does it with imap_subcell_exact on dense warps. Shell and pixel indices
// (hence the hit counts) are therefore bit-identical to the double path.
constexpr int kImapQCap = 6144;
constexpr float kMarginR = 5e-3f;      // Mpc/h: fp32 ulp of r ~ 1200 is 1.2e-4, rsqrt.approx adds 3e-4

__device__ __forceinline__ int fast_r_index(const float *s_r0, const float *s_rf, int nr, float r, bool &sure)
{
  if (r < s_r0[0]) { if (s_r0[0] - r < kMarginR) sure = false; return -1; }
  int lo = 0, hi = nr - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (s_r0[mid] <= r) lo = mid; else hi = mid - 1;
  }
  if (r - s_r0[lo] < kMarginR || fabsf(r - s_rf[lo]) < kMarginR) sure = false;
  if (lo + 1 < nr && s_r0[lo + 1] - r < kMarginR) sure = false;
  if (r < s_rf[lo]) return lo;
  return lo == nr - 1 ? nr : -2;
}

__device__ __forceinline__ bool near_int(float v, float fl, float margin) { return v - fl < margin || fl + 1.f - v < margin; }

// ang2pix_ring_z_phi in fp32 (same formulae as clr_ang2pix_ring_zphi); 1-|z| from (x^2+y^2)/(r(r+|z|)) so that
// the polar caps keep full relative accuracy
__device__ __forceinline__ long long fast_ang2pix(int nside, float nsf, float x, float y, float z, float r2, float rinv,
                                                  float margin, bool &sure)
{
  const float cth = z * rinv, za = fabsf(cth);
  float tt = atan2f(y, x) * 0.6366197723675814f;
  if (tt < 0.f) tt += 4.f;
  if (fabsf(za - 0.66666667f) < 4e-6f) sure = false;
  if (za <= 0.66666667f) {
    float t1 = nsf * (0.5f + tt), t2 = nsf * cth * 0.75f;
    float a = t1 - t2, b = t1 + t2, fa = floorf(a), fb = floorf(b);
    if (near_int(a, fa, margin) || near_int(b, fb, margin)) sure = false;
    int jp = (int)fa, jm = (int)fb;
    int ir = nside + 1 + jp - jm;
    int kshift = 1 - (ir & 1);
    int ip = (jp + jm - nside + kshift + 1) / 2;
    ip = clr_imodulo(ip, 4 * nside);
    return (long long)nside * (nside - 1) * 2 + (long long)(ir - 1) * 4 * nside + ip;
  }
  float ftt = floorf(tt), tp = tt - ftt;
  if (tp < 2e-5f || tp > 1.f - 2e-5f || tt >= 4.f) sure = false;
  float rr = r2 * rinv;
  float omz = (x * x + y * y) / (rr * (rr + fabsf(z)));
  float tmp = nsf * sqrtf(3.f * omz);
  float a = tp * tmp, b = (1.f - tp) * tmp, fa = floorf(a), fb = floorf(b);
  if (near_int(a, fa, margin) || near_int(b, fb, margin)) sure = false;
  int jp = (int)fa, jm = (int)fb;
  int ir = jp + jm + 1;
  float c = tt * (float)ir, fc = floorf(c);
  if (near_int(c, fc, margin)) sure = false;
  int ip = clr_imodulo((int)fc, 4 * ir);
  if (z > 0.f) return 2LL * ir * (ir - 1) + ip;
  return 12LL * nside * nside - 2LL * ir * (ir + 1) + ip;
}

struct ImapCell { double x0, y0, z0, dr_rsd; float temp; int nsub; };

__global__ void __launch_bounds__(kThreads)
imap_fast_kernel(const ClrDev d, const float *__restrict__ dens, const float *__restrict__ npot, ClrPop pop, ImapShells sh,
                 int nside, float *__restrict__ data, int *__restrict__ nadd, double rmin_here, double rmax_here)
{
  extern __shared__ unsigned char smem_raw[];
  ImapCell *s_cell = reinterpret_cast<ImapCell *>(smem_raw);                       // [kThreads]
  unsigned *s_q = reinterpret_cast<unsigned *>(s_cell + kThreads);                // [kImapQCap]
  float *s_r0 = reinterpret_cast<float *>(s_q + kImapQCap), *s_rf = s_r0 + sh.nr;  // [nr] each
  __shared__ int q_len;
  for (int i = threadIdx.x; i < sh.nr; i += blockDim.x) { s_r0[i] = sh.r0[i]; s_rf[i] = sh.rf[i]; }
  if (threadIdx.x == 0) q_len = 0;
  __syncthreads();
  const long long n_cells = (long long)d.nz_here * d.n * d.n;
  const long long num_pix = 12LL * nside * nside;
  const double dx = (double)(d.l_box / d.n);
  const double factor_vel = -d.fgrowth_0 / (1.5 * d.hubble_0 * d.OmegaM);
  const float nsf = (float)nside, margin = 8e-6f * nsf;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_iter = (n_cells + stride - 1) / stride;
  for (long long it = 0; it < n_iter; it++) {
    const long long i = it * stride + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int nsub = 0;
    double x0 = 0, y0 = 0, z0 = 0, dr_rsd = 0;
    float temp = 0.f;
    if (i < n_cells) {
      int ix, iy, iz;
      clr_cell(d, i, ix, iy, iz);
      z0 = __ldg(d.cd[2] + iz + d.iz0_here);
      y0 = __ldg(d.cd[1] + iy);
      x0 = __ldg(d.cd[0] + ix);
      double r0 = sqrt(x0 * x0 + y0 * y0 + z0 * z0);
      if (r0 <= rmax_here && r0 >= rmin_here) {
        double tmean = clr_lerp(d, r0, pop.nz, 0.0, 0.0);
        if (tmean > 0) {
          double bias = clr_bg_bz(d, r0, pop.bz);
          double dnorm = clr_lerp(d, r0, pop.norm, pop.norm_0, pop.norm_f);
          double rvel = factor_vel * dev_get_rvel(d, npot, ix, iy, iz, x0, y0, z0, r0);
          dr_rsd = rvel * clr_bg_v1(d, r0) * clr_bg_ih(d, r0);
          temp = (float)(tmean * clr_bias_model(d.bias_model, (double)dens[((long long)iz * d.n + iy) * d.pitch + ix], bias) * dnorm);
          int irad = dev_r_index(sh, r0);
          nsub = irad < 0 ? sh.nsub_lo : (irad >= sh.nr ? sh.nsub_hi : __ldg(sh.nsub + irad));
        }
      }
    }
    s_cell[threadIdx.x] = ImapCell{x0, y0, z0, dr_rsd, temp, nsub};
    if (nsub > 0) {
      const double dx_sub = dx / nsub;
      const float drf = (float)dr_rsd;
      for (int izz = 0; izz < nsub; izz++) {
        const float z = (float)(z0 + (izz + 0.5) * dx_sub);
        for (int iyy = 0; iyy < nsub; iyy++) {
          const float y = (float)(y0 + (iyy + 0.5) * dx_sub);
          const float yz2 = y * y + z * z;
          for (int ixx = 0; ixx < nsub; ixx++) {
            const float x = (float)(x0 + (ixx + 0.5) * dx_sub);
            const float r2 = fmaf(x, x, yz2);
            const float rinv = rsqrtf(r2);
            bool sure = r2 > 1e-6f && nsub < 256;
            int ir = fast_r_index(s_r0, s_rf, sh.nr, r2 * rinv + drf, sure);
            long long pix = 0;
            if (!sure || (ir >= 0 && ir < sh.nr)) pix = fast_ang2pix(nside, nsf, x, y, z, r2, rinv, margin, sure);
            if (sure) {
              if (ir >= 0 && ir < sh.nr) {
                atomicAdd(&data[ir * num_pix + pix], temp);
                atomicAdd(&nadd[ir * num_pix + pix], 1);
              }
            } else {
              int slot = nsub < 256 ? atomicAdd(&q_len, 1) : kImapQCap;
              if (slot < kImapQCap) s_q[slot] = ((unsigned)threadIdx.x << 24) | (unsigned)((izz * nsub + iyy) * nsub + ixx);
              else imap_subcell_exact(sh, nside, num_pix, x0 + (ixx + 0.5) * dx_sub, y0 + (iyy + 0.5) * dx_sub,
                                      z0 + (izz + 0.5) * dx_sub, dr_rsd, temp, data, nadd);
            }
          }
        }
      }
    }
    __syncthreads();
    // deferred sub-cells: the reference's double arithmetic on dense warps
    const int nq = min(q_len, kImapQCap);
    for (int k = threadIdx.x; k < nq; k += blockDim.x) {
      const unsigned e = s_q[k];
      const ImapCell c = s_cell[e >> 24];
      const int sub = (int)(e & 0xffffffu);
      const int ixx = sub % c.nsub, iyy = (sub / c.nsub) % c.nsub, izz = sub / (c.nsub * c.nsub);
      const double dx_sub = dx / c.nsub;
      imap_subcell_exact(sh, nside, num_pix, c.x0 + (ixx + 0.5) * dx_sub, c.y0 + (iyy + 0.5) * dx_sub,
                         c.z0 + (izz + 0.5) * dx_sub, c.dr_rsd, c.temp, data, nadd);
    }
    __syncthreads();
    if (threadIdx.x == 0) q_len = 0;
    __syncthreads();
  }
}

// ---- line-of-sight integrals ---------------------------------------------------------------------
struct LosPlan {
  const double *fac1, *fac2;    // kappa: r*D(1+z)*dr, r^2*D(1+z)*dr ; isw: fac1 = 2*pdot*dr
  const int *irmin, *irmax;     // sample range of each source plane (kappa.c:99-106)
  const double *inv_r_max;      // kappa only
  int nplanes;
  double dr;
  // several GPUs: a ray only meets this slab for r*u_z in [za, zb) (Mpc/h, observer frame); `restrict_z`
  // is set when no sample of any ray can wrap around the box, so the bound is safe
  int restrict_z;
  double za, zb;
  // kappa with the precomputed Hessian: H holds the local planes [zc0, zc1) only (a chunk of the slab); the partial
  // sums of the chunks are collected in double (accd) and rounded once at the end
  int zc0, zc1;
  double *accd;
};

// Hessian of the potential at every cell of the slab, computed ONCE (dev_tidal, the reference's float expressions) and
// stored as 8 floats per cell {xx, xy, xz, yy | yz, zz, -, -}: at nside 1024 on a 1024^3 grid ~11 samples of ~11
// different pixels land in every cell, so the kappa rays then fetch two 16-byte words per sample instead of the
// 19-point stencil. Cells no ray reaches (r > r_reach) are skipped.
constexpr int kTidalZ = 32;      // planes per CTA column
__global__ void __launch_bounds__(kThreads)
tidal_field_kernel(const ClrDev d, const float *__restrict__ npot, float4 *__restrict__ H, float r_reach, int zc0, int zc1)
{
  // 2.5-D blocking: a CTA owns a 32 (x) x 8 (y) column of cells and marches kTidalZ planes up z with a three-plane
  // register window {centre, x-, x+, y-, y+ (, the four xy diagonals of the middle plane)}: 9 loads per cell instead of
  // 19, every potential plane is fetched from DRAM once (+ 2 / kTidalZ), x / y neighbours come from L1. Same float
  // expressions as dev_tidal (beaming.c:85-116). The 32 B / cell of output are streaming stores.
  const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int z0 = zc0 + blockIdx.z * kTidalZ, z1 = min(z0 + kTidalZ, zc1);       // local planes of this chunk
  if (ix >= d.n || iy >= d.n) return;
  const long long ngx = d.pitch, plane = ngx * d.n;
  const int xh = ix + 1 == d.n ? 0 : ix + 1, xl = ix == 0 ? d.n - 1 : ix - 1;
  const long long y0 = (long long)iy * ngx, yh = (long long)(iy + 1 == d.n ? 0 : iy + 1) * ngx, yl = (long long)(iy == 0 ? d.n - 1 : iy - 1) * ngx;
  const float x = __ldg(d.cf[0] + ix), y = __ldg(d.cf[1] + iy);
  const float rr = r_reach * r_reach, xy2 = x * x + y * y;
  if (xy2 > rr) return;
  // storage index of local plane lz in [-1, nz]: the slab, then the halo planes behind it ([nz] = -1, [nz+1] = nz)
  auto pz = [&](int lz) -> const float * { return npot + (lz < 0 ? (long long)d.nz_here : lz >= d.nz_here ? (long long)d.nz_here + 1 : (long long)lz) * plane; };
  struct P5 { float c, xm, xp, ym, yp; };
  auto load5 = [&](const float *g) { P5 p; p.c = g[ix + y0]; p.xm = g[xl + y0]; p.xp = g[xh + y0]; p.ym = g[ix + yl]; p.yp = g[ix + yh]; return p; };
  struct D4 { float pp, mm, pm, mp; };
  auto load4 = [&](const float *g) { D4 q; q.pp = g[xh + yh]; q.mm = g[xl + yl]; q.pm = g[xh + yl]; q.mp = g[xl + yh]; return q; };
  P5 lo = load5(pz(z0 - 1)), mid = load5(pz(z0));
  D4 dmid = load4(pz(z0));
  for (int iz = z0; iz < z1; iz++) {
    const float *gh = pz(iz + 1);
    const P5 hi = load5(gh);
    const D4 dhi = iz + 1 < z1 ? load4(gh) : D4{0.f, 0.f, 0.f, 0.f};
    const float z = __ldg(d.cf[2] + iz + d.iz0_here);
    if (xy2 + z * z <= rr) {
      const float c = mid.c;
      const float t0 = (mid.xp + mid.xm - 2 * c);
      const float t3 = (mid.yp + mid.ym - 2 * c);
      const float t1 = (float)(0.25 * (double)(dmid.pp + dmid.mm - dmid.pm - dmid.mp));
      const float t5 = (hi.c + lo.c - 2 * c);
      const float t2 = (float)(0.25 * (double)(hi.xp + lo.xm - lo.xp - hi.xm));
      const float t4 = (float)(0.25 * (double)(hi.yp + lo.ym - lo.yp - hi.ym));
      const long long i = ix + (long long)d.n * (iy + (long long)d.n * (iz - zc0));
      __stcs(H + 2 * i, make_float4(t0, t1, t2, t3));
      __stcs(H + 2 * i + 1, make_float4(t4, t5, 0.f, 0.f));
    }
    lo = mid; mid = hi; dmid = dhi;
  }
}

template <bool KAPPA>
__global__ void __launch_bounds__(kThreads, 4)
los_kernel(const ClrDev d, const float *__restrict__ npot, const float4 *__restrict__ H, const double *__restrict__ pos,
           long long num_pix, LosPlan pl, float *__restrict__ data)
{
  const double idx = (double)(d.n / d.l_box);
  for (long long ip = blockIdx.x * (long long)blockDim.x + threadIdx.x; ip < num_pix; ip += (long long)gridDim.x * blockDim.x) {
    double u[3] = {pos[3 * ip], pos[3 * ip + 1], pos[3 * ip + 2]};
    double rot[6];
    if (KAPPA) {
      // kappa.c:128-145
      double prefac = idx * idx;
      double cth = u[2], sth, cph = 1, sph = 0;
      if (cth >= 1) cth = 1;
      if (cth <= -1) cth = -1;
      sth = sqrt((1 - cth) * (1 + cth));
      if (sth != 0) { cph = u[0] / sth; sph = u[1] / sth; }
      rot[0] = (cth * cth * cph * cph + sph * sph) * prefac;
      rot[1] = (2 * cph * sph * (cth * cth - 1)) * prefac;
      rot[2] = (-2 * cth * sth * cph) * prefac;
      rot[3] = (cth * cth * sph * sph + cph * cph) * prefac;
      rot[4] = (-2 * cth * sth * sph) * prefac;
      rot[5] = (sth * sth) * prefac;
    }
    double acc1 = 0, acc2 = 0;
    // samples whose NGP plane can lie in this slab: r_m * u_z in [za, zb), widened by one sample each side;
    // the exact test (dev_ngp) still decides inside the window
    int win_lo = 0, win_hi = 0x7fffffff;
    if (pl.restrict_z) {
      double ra, rb;
      if (fabs(u[2]) < 1e-12) { ra = (pl.za <= 0 && pl.zb > 0) ? -1e300 : 1e300; rb = (pl.za <= 0 && pl.zb > 0) ? 1e300 : -1e300; }
      else if (u[2] > 0) { ra = pl.za / u[2]; rb = pl.zb / u[2]; }
      else { ra = pl.zb / u[2]; rb = pl.za / u[2]; }
      double lo = floor(ra / pl.dr - 0.5) - 1, hi = ceil(rb / pl.dr - 0.5) + 1;
      win_lo = lo < 0 ? 0 : (lo > 2e9 ? 0x7fffffff : (int)lo);
      win_hi = hi < 0 ? -1 : (hi > 2e9 ? 0x7fffffff : (int)hi);
    }
    // NGP cell of sample ri = irr + 0.5 (beaming.c:148-157: (long)(x + 0.5) with one periodic wrap; x stays within
    // (-n, 2n), so the 32-bit truncating conversion gives the same integer); false when the plane is not in this slab
    auto ngp = [&](double ri, int c[3]) -> bool {
      const double rm = ri * pl.dr;
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        int v = __double2int_rz((rm * u[ax] + d.pos_obs[ax]) * idx + 0.5);
        if (v >= d.n) v -= d.n; else if (v < 0) v += d.n;
        c[ax] = v;
      }
      c[2] -= d.iz0_here;
      return c[2] >= pl.zc0 && c[2] < pl.zc1;
    };
    for (int ipl = 0; ipl < pl.nplanes; ipl++) {
      const int irmin = max(__ldg(pl.irmin + ipl), win_lo), irmax = min(__ldg(pl.irmax + ipl), win_hi);
      double ri = irmin + 0.5;                     // (irr + 0.5), exact in double: no int -> double conversion per sample
      if (KAPPA && H) {
        // software pipeline: the two 16-byte words of sample irr + 1 are requested before sample irr is consumed
        float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
        bool nin = false;
        if (irmin <= irmax) {
          int c[3];
          nin = ngp(ri, c);
          if (nin) {
            const long long cell = c[0] + (long long)d.n * (c[1] + (long long)d.n * (c[2] - pl.zc0));
            na = __ldg(H + 2 * cell); nb = __ldg(H + 2 * cell + 1);
          }
        }
        for (int irr = irmin; irr <= irmax; irr++) {
          const float4 a = na, b = nb;
          const bool in = nin;
          ri += 1.0;
          if (irr < irmax) {
            int c[3];
            nin = ngp(ri, c);
            if (nin) {
              const long long cell = c[0] + (long long)d.n * (c[1] + (long long)d.n * (c[2] - pl.zc0));
              na = __ldg(H + 2 * cell); nb = __ldg(H + 2 * cell + 1);
            }
          }
          if (in) {
            const float t[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
            double dotp = 0;
#pragma unroll
            for (int ax = 0; ax < 6; ax++) dotp += rot[ax] * t[ax];
            acc1 += dotp * __ldg(pl.fac1 + irr);
            acc2 += dotp * __ldg(pl.fac2 + irr);
          }
        }
      } else {
        for (int irr = irmin; irr <= irmax; irr++, ri += 1.0) {
          int c[3];
          if (!ngp(ri, c)) continue;
          if (KAPPA) {
            float t[6];
            dev_tidal(d, npot, c[0], c[1], c[2], t);
            double dotp = 0;
#pragma unroll
            for (int ax = 0; ax < 6; ax++) dotp += rot[ax] * t[ax];
            acc1 += dotp * __ldg(pl.fac1 + irr);
            acc2 += dotp * __ldg(pl.fac2 + irr);
          } else {
            float pd = npot[c[0] + (long long)c[1] * d.pitch + (long long)c[2] * d.pitch * d.n];
            acc1 += pd * __ldg(pl.fac1 + irr);
          }
        }
      }
      double add = KAPPA ? (acc1 - __ldg(pl.inv_r_max + ipl) * acc2) : acc1;
      if (KAPPA && pl.accd) pl.accd[(long long)ipl * num_pix + ip] += add;
      else {
        float *o = data + (long long)ipl * num_pix + ip;
        *o = (float)((double)(*o) + add);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
los_round_kernel(const double *__restrict__ accd, float *__restrict__ data, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    data[i] = (float)((double)data[i] + accd[i]);
}

int grid_for(clr_ctx *c, long long items, int per_sm)
{
  long long g = (items + kThreads - 1) / kThreads, cap = (long long)c->sm_count * per_sm;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

static double host_lerp(const clr_ctx *c, double r, const std::vector<double> &f, double f0, double ff)
{
  if (r <= 0) return f0;
  else if (r >= c->h_r[CLR_NA - 1]) return ff;
  int ir = (int)(r * c->p.glob_idr);
  return f[ir] + (f[ir + 1] - f[ir]) * (r - c->h_r[ir]) * c->p.glob_idr;
}

int clr_maps_imap(clr_ctx *c, int ipop, float *h_data, int32_t *h_nadd)
{
  clr_ctx::Pop &P = c->imap[ipop];
  CLR_CHECK(P.set && P.nr > 0, "imap population %d not set", ipop);
  CLR_CHECK(P.have_norm, "imap population %d has no normalisation", ipop);
  const int nr = P.nr;
  const long long num_pix = 12LL * P.nside * P.nside;
  // imap.c:151-174: sub-sampling of each shell (host, double)
  double dx = (double)(c->p.l_box / c->p.n_grid);
  double pixel_size = sqrt(4 * M_PI / num_pix);
  std::vector<int> nsub(nr);
  for (int i = 0; i < nr; i++) {
    double r0 = P.r0[i], rf = P.rf[i];
    double vol = (rf * rf * rf - r0 * r0 * r0) * pixel_size * pixel_size / 3;
    double sct = r0 * pixel_size, scr = rf - r0, scv = pow(vol, 0.333333);
    double sc0 = fmin(scv, fmin(sct, scr));
    nsub[i] = (int)(dx / sc0 + 0.5) + 1;
  }
  float *d_r0 = nullptr, *d_rf = nullptr, *d_data = nullptr;
  int *d_nsub = nullptr, *d_nadd = nullptr;
  CLR_CUDA(cudaMalloc(&d_r0, nr * sizeof(float)));
  CLR_CUDA(cudaMalloc(&d_rf, nr * sizeof(float)));
  CLR_CUDA(cudaMalloc(&d_nsub, nr * sizeof(int)));
  CLR_CUDA(cudaMalloc(&d_data, (size_t)nr * num_pix * sizeof(float)));
  CLR_CUDA(cudaMalloc(&d_nadd, (size_t)nr * num_pix * sizeof(int)));
  CLR_CUDA(cudaMemcpyAsync(d_r0, P.r0.data(), nr * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(d_rf, P.rf.data(), nr * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(d_nsub, nsub.data(), nr * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemsetAsync(d_data, 0, (size_t)nr * num_pix * sizeof(float), c->stream));
  CLR_CUDA(cudaMemsetAsync(d_nadd, 0, (size_t)nr * num_pix * sizeof(int), c->stream));
  ImapShells sh{d_r0, d_rf, d_nsub, nr, nsub[0], nsub[nr - 1]};
  ClrPop pop{P.d_a, P.d_b, P.d_norm, P.norm_0, P.norm_f};
  {
    StageScope sc(c, "imap_paint", 1);
    long long n_cells = (long long)c->dev.nz_here * c->dev.n * c->dev.n;
    if (c->exact_math)
      imap_kernel<<<grid_for(c, n_cells, 8), kThreads, 0, c->stream>>>(c->dev, c->d_dens, c->d_npot, pop, sh, P.nside, d_data, d_nadd,
                                                                        (double)P.r0[0] - 20., (double)P.rf[nr - 1] + 20.);
    else {
      size_t smem = kThreads * sizeof(ImapCell) + kImapQCap * sizeof(unsigned) + 2 * (size_t)nr * sizeof(float);
      CLR_CUDA(cudaFuncSetAttribute(imap_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      imap_fast_kernel<<<grid_for(c, n_cells, 4), kThreads, smem, c->stream>>>(c->dev, c->d_dens, c->d_npot, pop, sh, P.nside, d_data,
                                                                                d_nadd, (double)P.r0[0] - 20., (double)P.rf[nr - 1] + 20.);
    }
    CLR_CUDA(cudaGetLastError());
  }
  // every GPU painted its slab into a full-sky map (imap.c:123-132); sum them (io.c:727-735)
  if (clr_comm_allreduce_f32(c, d_data, (size_t)nr * num_pix)) return 1;
  if (clr_comm_allreduce_i32(c, d_nadd, (size_t)nr * num_pix)) return 1;
  CLR_CUDA(cudaMemcpyAsync(h_data, d_data, (size_t)nr * num_pix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaMemcpyAsync(h_nadd, d_nadd, (size_t)nr * num_pix * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_r0); cudaFree(d_rf); cudaFree(d_nsub); cudaFree(d_data); cudaFree(d_nadd);
  return 0;
}

int clr_maps_los(clr_ctx *c, int which, long long num_pix, const double *h_pos, int nplanes, const float *rf, float *h_data)
{
  CLR_CHECK(nplanes > 0 && nplanes <= CLR_NPLANES_MAX && num_pix > 0, "LOS maps: bad plane/pixel count");
  // kappa.c:89-116 / isw.c:88-113: radial sampling and kernels (host, double)
  int nr = c->p.n_grid / 2;                       // get_radial_params (common.c:333-337), NSAMP_RAD=1
  double dr = c->p.r_max / nr, idr = 1. / dr;
  std::vector<int> irmin(nplanes), irmax(nplanes);
  std::vector<double> inv_r_max(nplanes), fac1(nr), fac2(nr);
  for (int i = 0; i < nplanes; i++) {
    int i_r_here = (int)(rf[i] * idr + 0.5);
    inv_r_max[i] = 1. / (i_r_here * dr);
    irmax[i] = std::min(i_r_here, nr - 1);
  }
  irmin[0] = 0;
  for (int i = 1; i < nplanes; i++) irmin[i] = irmax[i - 1] + 1;
  for (int i = 0; i < nr; i++) {
    double rm = (i + 0.5) * dr;
    if (which == 0) {
      double pg = host_lerp(c, rm, c->h_d1, 1, c->h_d1[CLR_NA - 1]) * (1 + host_lerp(c, rm, c->h_z, 0, c->h_z[CLR_NA - 1]));
      fac1[i] = rm * pg * dr;
      fac2[i] = rm * rm * pg * dr;
    } else {
      fac1[i] = 2 * host_lerp(c, rm, c->h_pd, c->h_pd[0], c->h_pd[CLR_NA - 1]) * dr;
      fac2[i] = 0;
    }
  }
  double *d_pos = nullptr, *d_fac = nullptr, *d_inv = nullptr;
  int *d_ir = nullptr;
  float *d_data = nullptr;
  CLR_CUDA(cudaMalloc(&d_pos, (size_t)3 * num_pix * sizeof(double)));
  CLR_CUDA(cudaMalloc(&d_fac, (size_t)2 * nr * sizeof(double)));
  CLR_CUDA(cudaMalloc(&d_inv, nplanes * sizeof(double)));
  CLR_CUDA(cudaMalloc(&d_ir, 2 * nplanes * sizeof(int)));
  CLR_CUDA(cudaMalloc(&d_data, (size_t)nplanes * num_pix * sizeof(float)));
  CLR_CUDA(cudaMemcpyAsync(d_pos, h_pos, (size_t)3 * num_pix * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(d_fac, fac1.data(), nr * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(d_fac + nr, fac2.data(), nr * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(d_inv, inv_r_max.data(), nplanes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(d_ir, irmin.data(), nplanes * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemcpyAsync(d_ir + nplanes, irmax.data(), nplanes * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CLR_CUDA(cudaMemsetAsync(d_data, 0, (size_t)nplanes * num_pix * sizeof(float), c->stream));
  LosPlan pl{d_fac, d_fac + nr, d_ir, d_ir + nplanes, d_inv, nplanes, dr, 0, 0., 0., 0, c->dev.nz_here, nullptr};
  // NGP plane of a sample = (long)((r*u_z + pos_obs_z)*idx + 0.5) (beaming.c:148-157); it lies in the local planes
  // [p0, p1) iff r*u_z in [(iz0+p0-0.5)/idx - pos_obs_z, (iz0+p1-0.5)/idx - pos_obs_z) provided no sample wraps around the box
  const double idx = (double)(c->p.n_grid / c->p.l_box);
  const double far = (nr * dr + fabs(c->p.pos_obs[2])) * idx + 0.5, near = (c->p.pos_obs[2] - nr * dr) * idx + 0.5;
  const bool no_wrap = far < c->p.n_grid && near >= 0;
  auto window = [&](int p0, int p1) {
    pl.restrict_z = 1;
    pl.za = (c->dev.iz0_here + p0 - 0.5) / idx - c->p.pos_obs[2];
    pl.zb = (c->dev.iz0_here + p1 - 0.5) / idx - c->p.pos_obs[2];
  };
  if (c->nranks > 1 && no_wrap) window(0, c->dev.nz_here);
  // kappa: Hessian of every cell once (32 B / cell, tidal_field_kernel) when the rays oversample the grid; the rays then
  // fetch two 16-byte words per sample instead of the 19-point stencil. The Hessian lives in scratch memory the context
  // already owns (the z-pass buffer of the single-GPU transforms / the staging slab of the distributed ones: allocating
  // tens of GB would cost more than the kernels), so the slab is processed in chunks of planes that fit; the partial
  // sums of the chunks are collected in double and rounded once.
  const long long n_cells = (long long)c->dev.nz_here * c->dev.n * c->dev.n;
  const size_t plane_bytes = (size_t)c->dev.n * c->dev.n * 32;
  float4 *d_H = nullptr;
  size_t h_bytes = 0;
  double *d_accd = nullptr;
  if (which == 0 && c->los_precompute && (double)num_pix * nr * (c->dev.nz_here / (double)c->dev.n) > 2.0 * (double)n_cells) {
    if (c->d_fft_tmp && c->fft_tmp_bytes >= plane_bytes) { d_H = reinterpret_cast<float4 *>(c->d_fft_tmp); h_bytes = c->fft_tmp_bytes; }
    else if (c->d_stage && c->stage_floats * sizeof(float) >= plane_bytes) { d_H = reinterpret_cast<float4 *>(c->d_stage); h_bytes = c->stage_floats * sizeof(float); }
    else {
      // no scratch around (fields uploaded by the caller): own buffer of at most 1/8 of the slab, kept with the context
      const size_t want = plane_bytes * (size_t)std::max(1, std::min(c->dev.nz_here, std::max(8, c->dev.nz_here / 8)));
      if (c->d_los_hess && c->los_hess_bytes < want) { cudaFree(c->d_los_hess); c->d_los_hess = nullptr; }
      if (!c->d_los_hess) {
        if (cudaMalloc(&c->d_los_hess, want) == cudaSuccess) c->los_hess_bytes = want;
        else { cudaGetLastError(); c->d_los_hess = nullptr; c->los_hess_bytes = 0; }
      }
      d_H = reinterpret_cast<float4 *>(c->d_los_hess); h_bytes = c->los_hess_bytes;
    }
  }
  const int ppc = d_H ? (int)std::min<size_t>((size_t)c->dev.nz_here, h_bytes / plane_bytes) : 0;
  const int n_chunks = d_H ? (c->dev.nz_here + ppc - 1) / ppc : 0;
  if (n_chunks > 1 && !no_wrap) d_H = nullptr;                       // the chunk windows need rays that do not wrap
  if (d_H && n_chunks > 1) {
    if (cudaMalloc(&d_accd, (size_t)nplanes * num_pix * sizeof(double)) != cudaSuccess) { cudaGetLastError(); d_H = nullptr; }
    else CLR_CUDA(cudaMemsetAsync(d_accd, 0, (size_t)nplanes * num_pix * sizeof(double), c->stream));
  }
  if (d_H) {
    const float dxf = c->p.l_box / c->p.n_grid;
    pl.accd = d_accd;
    for (int ch = 0; ch < n_chunks; ch++) {
      pl.zc0 = ch * ppc; pl.zc1 = std::min(c->dev.nz_here, pl.zc0 + ppc);
      if (n_chunks > 1) window(pl.zc0, pl.zc1);
      {
        StageScope sc(c, "kappa_tidal", 1);
        dim3 tg((c->dev.n + 31) / 32, (c->dev.n + 7) / 8, (pl.zc1 - pl.zc0 + kTidalZ - 1) / kTidalZ);
        tidal_field_kernel<<<tg, kThreads, 0, c->stream>>>(c->dev, c->d_npot, d_H, (float)(nr * dr) + 2.f * dxf, pl.zc0, pl.zc1);
        CLR_CUDA(cudaGetLastError());
      }
      StageScope sc(c, "kappa_los", 1);
      los_kernel<true><<<grid_for(c, num_pix, 8), kThreads, 0, c->stream>>>(c->dev, c->d_npot, d_H, d_pos, num_pix, pl, d_data);
      CLR_CUDA(cudaGetLastError());
    }
    if (d_accd) {
      StageScope sc(c, "kappa_los", 1);
      los_round_kernel<<<grid_for(c, (long long)nplanes * num_pix, 8), kThreads, 0, c->stream>>>(d_accd, d_data, (long long)nplanes * num_pix);
      CLR_CUDA(cudaGetLastError());
    }
  } else {
    StageScope sc(c, which == 0 ? "kappa_los" : "isw_los", 1);
    if (which == 0)
      los_kernel<true><<<grid_for(c, num_pix, 8), kThreads, 0, c->stream>>>(c->dev, c->d_npot, nullptr, d_pos, num_pix, pl, d_data);
    else
      los_kernel<false><<<grid_for(c, num_pix, 8), kThreads, 0, c->stream>>>(c->dev, c->d_npot, nullptr, d_pos, num_pix, pl, d_data);
    CLR_CUDA(cudaGetLastError());
  }
  // slab-local ray segments: the accumulators are linear in the field, so the partial maps add up to the
  // full line-of-sight integral (replaces the ring rotation of beaming.c:325-352)
  if (clr_comm_allreduce_f32(c, d_data, (size_t)nplanes * num_pix)) return 1;
  CLR_CUDA(cudaMemcpyAsync(h_data, d_data, (size_t)nplanes * num_pix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_pos); cudaFree(d_fac); cudaFree(d_inv); cudaFree(d_ir); cudaFree(d_data); cudaFree(d_accd);
  return 0;
}
