// Source catalogue kernels: per-cell Poisson sampling, ordered scan, in-cell placement + RSD +
// base pixel, spherical properties, and the RSD estimator used under "beaming".
// Replaces srcs_set_cartesian_single (srcs.c:120-283), srcs_get_local_properties_single
// (srcs.c:386-416) and the RSD parts of the beam hooks (srcs.c:425-443, 486-504, 656-662).
// Compiled with -fmad=false: counts and pixel indices must be bit-exact against the CPU oracle
// for identical uniform draws, so double arithmetic must round like scalar C code.
#include "clr_internal.cuh"
#include "clr_stencil.cuh"
#include <math.h>
#include <utility>
#include <vector>

namespace {

constexpr int kThreads = 256;
constexpr int kCellsPerThread = 4;
constexpr int kChunk = kThreads * kCellsPerThread;   // cells per CTA chunk (flat, x fastest)
constexpr int kSub = 8; // chunks a Poisson CTA screens before it runs the exact path

// ---- gsl_ran_poisson (GSL randist/poisson.c) and its helpers, restated for the device ---------
__device__ __noinline__ double dev_gamma_large(ClrStream &s, double a)
{
  double sqa = sqrt(2 * a - 1), x, y, v;
  do {
    do {
      y = tan(3.14159265358979323846 * s.next());
      x = sqa * y + a - 1;
    } while (x <= 0);
    v = s.next();
  } while (v > (1 + y * y) * exp((a - 1) * log(x / (a - 1)) - sqa * y));
  return x;
}
__device__ double dev_gamma_int(ClrStream &s, unsigned int a)
{
  if (a < 12) {
    double prod = 1;
    for (unsigned int i = 0; i < a; i++) prod *= s.next_pos();
    return -log(prod);
  }
  return dev_gamma_large(s, (double)a);
}
__device__ double dev_stirling(double y1)
{
  double y2 = y1 * y1;
  return (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / y2) / y2) / y2) / y2) / y1 / 166320.0;
}
// Kachitvichyanukul & Schmeiser BINV/BTPE as laid out in GSL randist/binomial_tpe.c
__device__ __noinline__ unsigned int dev_binomial(ClrStream &s, double p, unsigned int n)
{
  int ix = 0, flipped = 0;
  if (n == 0) return 0;
  if (p > 0.5) { p = 1.0 - p; flipped = 1; }
  double q = 1 - p, sr = p / q, np = n * p;
  if (np < 14) {
    double f0 = pow(q, (double)n);
    bool done = false;
    while (!done) {
      double f = f0, u = s.next();
      for (ix = 0; ix <= 110; ++ix) {
        if (u < f) { done = true; break; }
        u -= f;
        f *= sr * (n - ix) / (ix + 1);
      }
    }
  } else {
    double ffm = np + p;
    int m = (int)ffm;
    double xm = m + 0.5, npq = np * q;
    double p1 = floor(2.195 * sqrt(npq) - 4.6 * q) + 0.5;
    double xl = xm - p1, xr = xm + p1;
    double c = 0.134 + 20.5 / (15.3 + m);
    double p2 = p1 * (1.0 + c + c);
    double al = (ffm - xl) / (ffm - xl * p);
    double lambda_l = al * (1.0 + 0.5 * al);
    double ar = (xr - ffm) / (xr * q);
    double lambda_r = ar * (1.0 + 0.5 * ar);
    double p3 = p2 + c / lambda_l;
    double p4 = p3 + c / lambda_r;
    for (;;) {
      double var, accept;
      double u = s.next() * p4;
      double v = s.next();
      if (u <= p1) { ix = (int)(xm - p1 * v + u); break; }
      else if (u <= p2) {
        double x = xl + (u - p1) / c;
        v = v * c + 1.0 - fabs(x - xm) / p1;
        if (v > 1.0 || v <= 0.0) continue;
        ix = (int)x;
      } else if (u <= p3) {
        ix = (int)(xl + log(v) / lambda_l);
        if (ix < 0) continue;
        v *= ((u - p2) * lambda_l);
      } else {
        ix = (int)(xr - log(v) / lambda_r);
        if (ix > (double)n) continue;
        v *= ((u - p3) * lambda_r);
      }
      int k = abs(ix - m);
      if (k <= 20) {
        double g = (n + 1) * sr, f = 1.0;
        var = v;
        if (m < ix) { for (int i = m + 1; i <= ix; i++) f *= (g / i - sr); }
        else if (m > ix) { for (int i = ix + 1; i <= m; i++) f /= (g / i - sr); }
        accept = f;
      } else {
        var = log(v);
        if (k < npq / 2 - 1) {
          double amaxp = k / npq * ((k * (k / 3.0 + 0.625) + (1.0 / 6.0)) / npq + 0.5);
          double ynorm = -(k * k / (2.0 * npq));
          if (var < ynorm - amaxp) break;
          if (var > ynorm + amaxp) continue;
        }
        double x1 = ix + 1.0, w1 = n - ix + 1.0, f1 = m + 1.0, z1 = n + 1.0 - m;
        accept = xm * log(f1 / x1) + (n - m + 0.5) * log(z1 / w1) + (ix - m) * log(w1 * p / (x1 * q))
                 + dev_stirling(f1) + dev_stirling(z1) - dev_stirling(x1) - dev_stirling(w1);
      }
      if (var <= accept) break;
    }
  }
  return flipped ? (n - ix) : (unsigned int)ix;
}
__device__ unsigned int dev_poisson(ClrStream &s, double mu)
{
  unsigned int k = 0;
  while (mu > 10) {
    unsigned int m = (unsigned int)(mu * (7.0 / 8.0));
    double X = dev_gamma_int(s, m);
    if (X >= mu) return k + dev_binomial(s, mu / X, m - 1);
    k += m;
    mu -= X;
  }
  double emu = exp(-mu), prod = 1.0;
  do {
    prod *= s.next();
    k++;
  } while (prod > emu);
  return k - 1;
}

// ---- pass 1: lambda and Poisson count per cell (srcs.c:156-184) --------------------------------
// fp32 screening of one cell: true when the cell is SURELY empty, i.e. its first uniform lies below a
// rigorous lower bound of exp(-lambda) (gsl_ran_poisson returns 0 iff u0 <= exp(-mu)). The bound comes
// from one float4 per r-bin (poisson_bound_kernel): {upper bound of n(r)*norm(r)*cell volume, lowest b(r),
// highest b(r)} over the bin and its neighbours, so the fp32 bin index may be off by one.
struct ScreenK { float rcutf, idrf, rtabf; int bias_model; };
__device__ __forceinline__ bool screen_cell(const ScreenK &k, const float4 *__restrict__ bound, float r2, float dl,
                                            uint32_t word)
{
  // branch free (the tests below are data dependent and would diverge in nearly every warp)
  const float rf = clr_sqrt_fast(r2);
  const bool outside = rf > k.rcutf + 0.05f;                 // outside the sampled sphere (srcs.c:169)
  if (outside) return true;                                  // box corners: spatially coherent early out
  const bool inside = rf < k.rcutf - 0.05f && rf > 0.05f && rf < k.rtabf - 1.f;   // else: exact path decides
  const float4 e = __ldg(bound + clr_magic_int(clr_floor_magic(fminf(rf * k.idrf, (float)(CLR_NA - 1)))));
  float bm;                                                  // upper bound of |bias_model(dl, b)|, b in [e.y, e.z]
  if (k.bias_model == 2) {
    const float ex = clr_ex2_fast(1.4426951f * e.y * dl * clr_rcp_fast(fmaxf(1.f + dl, 1e-30f)));
    const float li = fmaxf(fabsf(fmaf(e.y, dl, 1.f)), fabsf(fmaf(e.z, dl, 1.f)));
    bm = dl < 0.f ? ex : li;
  } else if (k.bias_model == 3) bm = fmaxf(fmaxf(1.f + e.y * dl, 1.f + e.z * dl), 0.f);
  else { float lg = __log2f(fmaxf(1.f + dl, 1e-30f)); bm = exp2f(fmaxf(e.y * lg, e.z * lg)); }
  bm = dl <= -1.f ? 0.f : bm;
  const float lam_hi = e.x * bm * 1.001f;
  const float e_lo = clr_ex2_fast(-1.4426951f * lam_hi) * (1.f - 1e-5f);
  // upper bound of the first uniform from its top 23 bits, conversion free: ((word >> 9) + 1) * 2^-23 >= u0
  const float u0_hi = (__uint_as_float(0x4B000000u | (word >> 9)) - 8388607.f) * 1.1920928955078125e-07f;
  // gsl_ran_poisson returns 0 iff u0 <= exp(-mu) only in its mu <= 10 branch (Knuth's product); above it the gamma /
  // binomial reduction consumes the draws differently, so such cells always take the exact path
  return outside || (inside && lam_hi < 9.99f && u0_hi <= e_lo);   // NaN tables: comparison false -> exact path
}

// Group screen: ONE bound for the four neighbouring cells a thread owns (they share a Philox block). With
// dmax = largest delta, [rmin, rmax] = their radii and umax = the largest of their first uniforms, all four are
// surely empty if umax <= exp(-lambda_hi(dmax)): bias_model is increasing in delta for b >= 0, and for delta < 0 the
// smallest b gives the largest value, for delta >= 0 the largest b. ~88 % of the groups inside the sphere pass; the
// rest is re-screened cell by cell on dense warps (phase 1b). Needs b >= 0 over the window, else no proof.
__device__ __forceinline__ bool screen_group(const ScreenK &k, const float4 *__restrict__ bound_grp, float r2min,
                                             float r2max, float dmax, uint32_t umax)
{
  const float rmin = clr_sqrt_fast(r2min), rmax = clr_sqrt_fast(r2max);
  const bool inside = rmax < k.rcutf - 0.05f && rmin > 0.05f && rmax < k.rtabf - 1.f;
  const float4 e = __ldg(bound_grp + clr_magic_int(clr_floor_magic(fminf(rmin * k.idrf, (float)(CLR_NA - 1)))));
  float bm;
  if (k.bias_model == 2) {
    const float ex = clr_ex2_fast(1.4426951f * e.y * dmax * clr_rcp_fast(fmaxf(1.f + dmax, 1e-30f)));
    bm = dmax < 0.f ? ex : fmaf(e.z, dmax, 1.f);
  } else if (k.bias_model == 3) bm = fmaxf(dmax < 0.f ? fmaf(e.y, dmax, 1.f) : fmaf(e.z, dmax, 1.f), 0.f);
  else return false;
  bm = dmax <= -1.f ? 0.f : bm;
  const float lam_hi = e.x * bm * 1.001f;
  const float e_lo = clr_ex2_fast(-1.4426951f * lam_hi) * (1.f - 1e-5f);
  const float u_hi = (__uint_as_float(0x4B000000u | (umax >> 9)) - 8388607.f) * 1.1920928955078125e-07f;
  return inside && e.y >= 0.f && lam_hi < 9.99f && u_hi <= e_lo;   // NaN anywhere: comparison false -> not proven
}

// one entry per r-bin of the NA grid; window [ir-1, ir+2] covers an off-by-one fp32 bin index
// `up`: how many bins above ir the window reaches (2 for one cell; wider for the 4-cell group table, whose
// entry is looked up at the SMALLEST radius of the group and must cover the largest one)
__global__ void poisson_bound_kernel(const ClrDev d, ClrPop pop, float vol, float4 *__restrict__ bound, int up)
{
  int ir = blockIdx.x * blockDim.x + threadIdx.x;
  if (ir >= CLR_NA) return;
  double a = 0, nm = fmax(fabs(pop.norm_0), fabs(pop.norm_f)), blo = 1e300, bhi = -1e300;
  bool bad = false;
  for (int k = ir - 1; k <= ir + up; k++) {
    int kk = k < 0 ? 0 : (k > CLR_NA - 1 ? CLR_NA - 1 : k);
    double n = pop.nz[kk], m = pop.norm[kk], b = pop.bz[kk];
    if (!(n == n) || !(m == m) || !(b == b)) bad = true;
    a = fmax(a, fabs(n)); nm = fmax(nm, fabs(m));
    blo = fmin(blo, b); bhi = fmax(bhi, b);
  }
  float4 e;
  e.x = bad ? __int_as_float(0x7fc00000) : __double2float_ru(a * nm * (double)vol * 1.001);
  e.y = __double2float_rd(blo); e.z = __double2float_ru(bhi); e.w = 0.f;
  bound[ir] = e;
}

// counts: int32 per cell, unpadded flat order ix + n*(iy + n*iz_local); chunk_tot[chunk] = sum.
// RNG: the first uniform of cell g is word g&3 of the Philox block shared by cells 4(g>>2)..+3
// (third_party/shim/gsl_shim.c:shim_philox_seek_cell), so one thread screens 4 neighbouring cells per block.
// Output (COMPACT = true, the default): no dense count array is written. A CTA keeps the counts of its 8192-cell super
// chunk in shared memory and emits only the occupied cells, in cell order, as 32-bit entries (local cell << 16 | count)
// into the super chunk's OWN slice of the count buffer (entries[sup * 8192 ...]: at most one entry per cell, so the slice
// always suffices), plus {entries, sources} per super chunk. At <N> = 0.03 that is ~1 KB written per 32 KB slice
// instead of the whole slice, and the expansion pass reads the entries instead of every count. Counts above 65535
// raise *overflow (the host then re-runs with COMPACT = false: dense int32 counts as in the reference, srcs.c:125).
template <bool COMPACT>
__global__ void __launch_bounds__(kThreads, 4)
poisson_kernel(const ClrDev d, const float *__restrict__ dens, ClrPop pop, const float4 *__restrict__ bound,
               const float4 *__restrict__ bound_grp, uint32_t seed, int ipop, int32_t *__restrict__ counts,
               int32_t *__restrict__ chunk_tot, int32_t *__restrict__ sup_entries, int *__restrict__ overflow,
               long long n_cells)
{
  const double dx = (double)(d.l_box / d.n);       // float division, as in the reference (srcs.c:147)
  const double cell_vol = dx * dx * dx;
  const double rcut = (double)(d.l_box / 2) + 20.;
  ScreenK sk;
  sk.rcutf = (float)rcut;
  sk.idrf = (float)d.glob_idr; sk.rtabf = (float)d.r_tab_max; sk.bias_model = d.bias_model;
  const uint32_t strm = 1 + 2 * ipop;
  const unsigned long long goff = (unsigned long long)d.n * d.n * (unsigned long long)d.iz0_here;
  const bool rows4 = (d.n & 3) == 0;               // 4-cell groups never straddle a row
  // A CTA works on kSub consecutive chunks at a time: phase 1 screens all of them and queues the cells that
  // need the exact path; phase 2 then runs over a queue long enough to fill every lane of the CTA.
  __shared__ int sub_tot[kSub];
  __shared__ unsigned short q_cell[kSub * kChunk];
  __shared__ unsigned short q_grp[kSub * kChunk / 4];
  // COMPACT: counts of the occupied cells of the super chunk (valid where the bitmap has the cell's bit set)
  __shared__ unsigned short cnt_s[COMPACT ? kSub * kChunk : 1];
  __shared__ unsigned bm_s[COMPACT ? kSub * kChunk / 32 : 1];
  __shared__ int wsum_e[kThreads / 32];
  static_assert(kSub * kChunk / 32 == kThreads, "one bitmap word per thread");
  __shared__ int q_len, g_len;
  const long long n_chunks = (n_cells + kChunk - 1) / kChunk;
  const long long n_super = (n_chunks + kSub - 1) / kSub;
  if (threadIdx.x == 0) { q_len = 0; g_len = 0; }
  if (threadIdx.x < kSub) sub_tot[threadIdx.x] = 0;
  __syncthreads();
  for (long long sup = blockIdx.x; sup < n_super; sup += gridDim.x) {
    const long long cell0 = sup * kSub * kChunk;
    if (COMPACT) bm_s[threadIdx.x] = 0u;            // (made visible by the barrier that ends phase 1)
    // ---- phase 1 (all cells, fp32). Everything that is not surely empty -- about 3% of the cells at
    // <N> = 0.03 -- is queued for the exact double-precision path.
#pragma unroll 1
    for (int sub = 0; sub < kSub; sub++) {
      const int lc0 = sub * kChunk + threadIdx.x * kCellsPerThread;
      const long long i0 = cell0 + lc0;
      if (i0 >= n_cells) break;
      if (rows4 && i0 + 3 < n_cells) {
        int ix, iy, iz;
        clr_cell(d, i0, ix, iy, iz);
        const float *drow = dens + ((long long)iz * d.n + iy) * d.pitch + ix;
        float2 da = __ldg(reinterpret_cast<const float2 *>(drow));
        float2 db = __ldg(reinterpret_cast<const float2 *>(drow) + 1);
        float dl[4] = {da.x, da.y, db.x, db.y};
        float yf = __ldg(d.cf[1] + iy), zf = __ldg(d.cf[2] + iz + d.iz0_here);
        float4 xf4 = __ldg(reinterpret_cast<const float4 *>(d.cf[0] + ix));
        float xf[4] = {xf4.x, xf4.y, xf4.z, xf4.w};
        float yz2 = yf * yf + zf * zf;
        // box corners (48 % of the cells lie outside the sampled sphere, srcs.c:169): all four cells beyond the
        // cut -> no random numbers, no screening; neighbouring threads agree, so the branch is coherent
        const float r2min = fminf(fminf(xf[0] * xf[0], xf[1] * xf[1]), fminf(xf[2] * xf[2], xf[3] * xf[3])) + yz2;
        {
          const float rc = sk.rcutf + 0.05f;
          if (r2min > rc * rc * 1.000001f) {
            if (!COMPACT) *reinterpret_cast<int4 *>(counts + i0) = make_int4(0, 0, 0, 0);
            continue;
          }
        }
        unsigned long long grp = ((unsigned long long)i0 + goff) >> 2;
        uint32_t w[4];
        clr_philox((uint32_t)grp, (uint32_t)(grp >> 32), 0u, strm | 0x80000000u, seed, 0u, w);
        if (!COMPACT) *reinterpret_cast<int4 *>(counts + i0) = make_int4(0, 0, 0, 0);
        const float r2max = fmaxf(fmaxf(xf[0] * xf[0], xf[1] * xf[1]), fmaxf(xf[2] * xf[2], xf[3] * xf[3])) + yz2;
        const float dmax = fmaxf(fmaxf(dl[0], dl[1]), fmaxf(dl[2], dl[3]));
        const uint32_t umax = max(max(w[0], w[1]), max(w[2], w[3]));
        if (!screen_group(sk, bound_grp, r2min, r2max, dmax, umax))
          q_grp[atomicAdd(&g_len, 1)] = (unsigned short)(lc0 >> 2);       // phase 1b looks at the cells one by one
      } else {
        for (int q = 0; q < kCellsPerThread; q++) {
          long long i = i0 + q;
          if (i >= n_cells) break;
          int ix, iy, iz;
          clr_cell(d, i, ix, iy, iz);
          float xf = __ldg(d.cf[0] + ix), yf = __ldg(d.cf[1] + iy), zf = __ldg(d.cf[2] + iz + d.iz0_here);
          float dl = dens[((long long)iz * d.n + iy) * d.pitch + ix];
          unsigned long long gcell = (unsigned long long)i + goff;
          uint32_t w[4];
          clr_philox((uint32_t)(gcell >> 2), (uint32_t)(gcell >> 34), 0u, strm | 0x80000000u, seed, 0u, w);
          if (screen_cell(sk, bound, xf * xf + yf * yf + zf * zf, dl, w[gcell & 3])) { if (!COMPACT) counts[i] = 0; }
          else q_cell[atomicAdd(&q_len, 1)] = (unsigned short)(lc0 + q);
        }
      }
    }
    __syncthreads();
    // ---- phase 1b (groups the group screen could not clear, ~12 %): the per-cell fp32 screen on dense warps
    const int ng = g_len;
    for (int k = threadIdx.x; k < ng; k += kThreads) {
      const int lc0 = (int)q_grp[k] << 2;
      const long long i0 = cell0 + lc0;
      int ix, iy, iz;
      clr_cell(d, i0, ix, iy, iz);
      const float *drow = dens + ((long long)iz * d.n + iy) * d.pitch + ix;
      float2 da = __ldg(reinterpret_cast<const float2 *>(drow));
      float2 db = __ldg(reinterpret_cast<const float2 *>(drow) + 1);
      float dl[4] = {da.x, da.y, db.x, db.y};
      float yf = __ldg(d.cf[1] + iy), zf = __ldg(d.cf[2] + iz + d.iz0_here);
      float4 xf4 = __ldg(reinterpret_cast<const float4 *>(d.cf[0] + ix));
      float xf[4] = {xf4.x, xf4.y, xf4.z, xf4.w};
      float yz2 = yf * yf + zf * zf;
      unsigned long long grp = ((unsigned long long)i0 + goff) >> 2;
      uint32_t w[4];
      clr_philox((uint32_t)grp, (uint32_t)(grp >> 32), 0u, strm | 0x80000000u, seed, 0u, w);
      unsigned pend = 0;
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (!screen_cell(sk, bound, xf[q] * xf[q] + yz2, dl[q], w[q])) pend |= 1u << q;
      if (pend) {
        int base = atomicAdd(&q_len, __popc(pend));
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (pend >> q & 1) q_cell[base++] = (unsigned short)(lc0 + q);
      }
    }
    __syncthreads();
    // ---- phase 2 (queued cells, double): the reference arithmetic, bit for bit
    const int nq = q_len;
    for (int k = threadIdx.x; k < nq; k += kThreads) {
      const int lc = q_cell[k];
      long long i = cell0 + lc;
      int ix, iy, iz;
      clr_cell(d, i, ix, iy, iz);
      long long row = (long long)iz * d.n + iy;
      double z0 = __ldg(d.cd[2] + iz + d.iz0_here);
      double y0 = __ldg(d.cd[1] + iy);
      double x0 = __ldg(d.cd[0] + ix);
      double r = sqrt(x0 * x0 + y0 * y0 + z0 * z0);
      int npp = 0;
      if (r < rcut) {
        double ndens = clr_bg_nz(d, r, pop.nz);
        if (ndens > 0) {
          double bias = clr_bg_bz(d, r, pop.bz);
          double dnorm = clr_lerp(d, r, pop.norm, pop.norm_0, pop.norm_f);
          double delta = dens[row * d.pitch + ix];
          double lambda = ndens * cell_vol * clr_bias_model(d.bias_model, delta, bias) * dnorm;
          unsigned long long gcell = (unsigned long long)i + goff;
          uint32_t w[4];
          clr_philox((uint32_t)(gcell >> 2), (uint32_t)(gcell >> 34), 0u, strm | 0x80000000u, seed, 0u, w);
          ClrStream s(seed, strm, gcell);
          s.set_first(w[gcell & 3]);
          npp = (int)dev_poisson(s, lambda);
        }
      }
      if (COMPACT) {
        if (npp) {
          cnt_s[lc] = (unsigned short)min(npp, 65535);
          atomicOr(&bm_s[lc >> 5], 1u << (lc & 31));
          atomicAdd(&sub_tot[0], npp);
          if (npp > 65535) *overflow = 1;
        }
      } else {
        counts[i] = npp;
        if (npp) atomicAdd(&sub_tot[lc / kChunk], npp);
      }
    }
    __syncthreads();
    if (COMPACT) {
      // ordered compaction: thread t owns bitmap word t = the cells [32 t, 32 t + 32) of the super chunk
      unsigned word = bm_s[threadIdx.x];
      const int occ = __popc(word);
      int incl = occ;
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (lane == 31) wsum_e[w] = incl;
      __syncthreads();
      int woff = 0, tot_e = 0;
#pragma unroll
      for (int k = 0; k < kThreads / 32; k++) { const int t = wsum_e[k]; if (k < w) woff += t; tot_e += t; }
      int32_t *o = counts + cell0 + woff + incl - occ;
      while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        const int lc = threadIdx.x * 32 + b;
        *o++ = (int32_t)(((unsigned)lc << 16) | cnt_s[lc]);
      }
      if (threadIdx.x == 0) { chunk_tot[sup] = sub_tot[0]; sup_entries[sup] = tot_e; sub_tot[0] = 0; }
    } else if (threadIdx.x < kSub) {
      long long chunk = sup * kSub + threadIdx.x;
      if (chunk < n_chunks) chunk_tot[chunk] = sub_tot[threadIdx.x];
      sub_tot[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) { q_len = 0; g_len = 0; }
    __syncthreads();
  }
}

__global__ void zero_words_kernel(uint32_t *p, int n) { if ((int)threadIdx.x < n) p[threadIdx.x] = 0u; }

// dense per-cell counts (the reference's nsources array, srcs.c:125) from the compact entries: for clr_srcs_get_counts
__global__ void __launch_bounds__(kThreads)
entries_to_counts_kernel(const int32_t *__restrict__ entries, const int32_t *__restrict__ sup_entries, int32_t *__restrict__ dense,
                         long long n_super, long long n_cells)
{
  for (long long sup = blockIdx.x; sup < n_super; sup += gridDim.x) {
    const long long cell0 = sup * kSub * kChunk;
    for (int k = threadIdx.x; k < kSub * kChunk && cell0 + k < n_cells; k += kThreads) dense[cell0 + k] = 0;
    __syncthreads();
    const int ne = sup_entries[sup];
    for (int k = threadIdx.x; k < ne; k += kThreads) {
      const unsigned e = (unsigned)entries[cell0 + k];
      dense[cell0 + (e >> 16)] = (int32_t)(e & 0xffffu);
    }
    __syncthreads();
  }
}

// expansion from the compact entries: one CTA per super chunk, entry k -> `count` source references at the ordered slots
__global__ void __launch_bounds__(kThreads)
expand_entries_kernel(const int32_t *__restrict__ entries, const int32_t *__restrict__ sup_entries,
                      const long long *__restrict__ sup_offs, unsigned long long *__restrict__ src_ref, long long n_super)
{
  __shared__ int wsum[kThreads / 32];
  for (long long sup = blockIdx.x; sup < n_super; sup += gridDim.x) {
    const int ne = sup_entries[sup];
    if (ne == 0) continue;                                 // uniform per CTA
    const long long cell0 = sup * kSub * kChunk;
    long long run = sup_offs[sup];
    for (int k0 = 0; k0 < ne; k0 += kThreads) {
      const int k = k0 + threadIdx.x;
      const unsigned e = k < ne ? (unsigned)entries[cell0 + k] : 0u;
      const int cnt = (int)(e & 0xffffu);
      int incl = cnt;
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      __syncthreads();
      if (lane == 31) wsum[w] = incl;
      __syncthreads();
      int woff = 0, tot = 0;
#pragma unroll
      for (int q = 0; q < kThreads / 32; q++) { const int t = wsum[q]; if (q < w) woff += t; tot += t; }
      const unsigned long long cell = (unsigned long long)(cell0 + (e >> 16));
      unsigned long long *o = src_ref + run + woff + incl - cnt;
      for (int ip = 0; ip < cnt; ip++) o[ip] = (cell << 24) | (unsigned)ip;
      run += tot;
    }
  }
}

// exclusive scan of the chunk totals, offs[n_chunks] = grand total. Two launches over kScanBlocks contiguous
// segments: (1) segment sums, (2) every CTA scans the kScanBlocks sums in shared memory to get its base, then
// scans its own segment (a single 1024-thread CTA took 0.42 ms for the 10^6 chunks of a 1024^3 grid).
constexpr int kScanBlocks = 256;
__global__ void __launch_bounds__(256)
scan_sums_kernel(const int32_t *__restrict__ tot, long long *__restrict__ seg_sum, long long n_chunks)
{
  const long long per = (n_chunks + kScanBlocks - 1) / kScanBlocks;
  const long long b = blockIdx.x * per, e = b + per < n_chunks ? b + per : n_chunks;
  long long sum = 0;
  for (long long i = b + threadIdx.x; i < e; i += blockDim.x) sum += tot[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __shared__ long long ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int k = 0; k < 8; k++) t += ws[k];
    seg_sum[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256)
scan_chunks_kernel(const int32_t *__restrict__ tot, const long long *__restrict__ seg_sum, long long *__restrict__ offs,
                   long long n_chunks)
{
  __shared__ long long base_s, ws[8];
  // base of this segment = sum of the earlier segment sums
  {
    long long v = threadIdx.x < blockIdx.x ? seg_sum[threadIdx.x] : 0;      // kScanBlocks == blockDim.x
    long long all = seg_sum[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
    __shared__ long long wa[8];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = v; wa[threadIdx.x >> 5] = all; }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0, ta = 0;
      for (int k = 0; k < 8; k++) { t += ws[k]; ta += wa[k]; }
      base_s = t;
      if (blockIdx.x == 0) offs[n_chunks] = ta;
    }
    __syncthreads();
  }
  const long long per = (n_chunks + kScanBlocks - 1) / kScanBlocks;
  const long long b = blockIdx.x * per, e = b + per < n_chunks ? b + per : n_chunks;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long run = base_s;
  for (long long i0 = b; i0 < e; i0 += blockDim.x) {
    const long long i = i0 + threadIdx.x;
    long long v = i < e ? tot[i] : 0, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    __syncthreads();
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    long long woff = 0, blk = 0;
    for (int k = 0; k < 8; k++) { long long t = ws[k]; if (k < w) woff += t; blk += t; }
    if (i < e) offs[i] = run + woff + incl - v;
    run += blk;
  }
}

// srcs.c:87-118 (get_rvel): central differences of the potential, periodic in x,y, z through the
// halo planes stored after the slab (plane nz_here = left neighbour, nz_here+1 = right neighbour)
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

// ---- pass 2: positions, RSD, base pixel (srcs.c:238-276) ---------------------------------------
// Two kernels so that the expensive per-source arithmetic runs on dense warps:
//  (a) expand_kernel, per cell chunk: in-CTA scan of the counts, then every occupied cell writes one
//      64-bit reference (local cell index << 24 | ip) per source at its ordered slot;
//  (b) place_src_kernel, one thread per source: decodes the reference and evaluates position, RSD and
//      base pixel. Ordering = cell order, as the reference's single-thread loop produces.
__global__ void __launch_bounds__(kThreads)
expand_kernel(const int32_t *__restrict__ counts, const long long *__restrict__ chunk_offs,
              unsigned long long *__restrict__ src_ref, long long n_cells)
{
  __shared__ int wsum[kThreads / 32];
  for (long long chunk = blockIdx.x; chunk * kChunk < n_cells; chunk += gridDim.x) {
    long long base = chunk_offs[chunk];
    if (chunk_offs[chunk + 1] == base) continue;           // uniform per CTA: empty chunk
    // thread t owns cells [4t, 4t+4) of the chunk so that the in-CTA scan follows the cell order
    long long i0 = chunk * kChunk + (long long)threadIdx.x * kCellsPerThread;
    int c[kCellsPerThread], tsum = 0;
    if (i0 + kCellsPerThread <= n_cells) {
      int4 v = *reinterpret_cast<const int4 *>(counts + i0);
      c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
#pragma unroll
      for (int q = 0; q < kCellsPerThread; q++) c[q] = (i0 + q < n_cells) ? counts[i0 + q] : 0;
    }
#pragma unroll
    for (int q = 0; q < kCellsPerThread; q++) tsum += c[q];
    int incl = tsum;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    // exclusive prefix of the 8 warp sums: every warp scans them redundantly with three shuffles
    int ws = lane < kThreads / 32 ? wsum[lane] : 0, wincl = ws;
#pragma unroll
    for (int o = 1; o < kThreads / 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, wincl, o); if (lane >= o) wincl += v; }
    const int woff = __shfl_sync(0xffffffffu, wincl - ws, w);
    long long off = base + woff + incl - tsum;
    __syncthreads();
    // one loop over the thread's sources (usually 0 or 1) instead of one data-dependent loop per cell
    for (int k = 0; k < tsum; k++) {
      int q = 0, ip = k;
      if (ip >= c[0]) { ip -= c[0]; q = 1; if (ip >= c[1]) { ip -= c[1]; q = 2; if (ip >= c[2]) { ip -= c[2]; q = 3; } } }
      src_ref[off + k] = ((unsigned long long)(i0 + q) << 24) | (unsigned)ip;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
place_src_kernel(const ClrDev d, const float *__restrict__ npot, const unsigned long long *__restrict__ src_ref,
                 uint32_t seed, int ipop, float4 *__restrict__ pos, int32_t *__restrict__ ipix, long long nsrc)
{
  const double dx = (double)(d.l_box / d.n);
  const double factor_vel = -d.fgrowth_0 / (1.5 * d.hubble_0 * d.OmegaM);
  for (long long is = blockIdx.x * (long long)blockDim.x + threadIdx.x; is < nsrc; is += (long long)gridDim.x * blockDim.x) {
    unsigned long long ref = src_ref[is];
    long long i = (long long)(ref >> 24);
    int ip = (int)(ref & 0xffffffu);
    int ix, iy, iz;
    clr_cell(d, i, ix, iy, iz);
    double z0 = __ldg(d.cd[2] + iz + d.iz0_here);
    double y0 = __ldg(d.cd[1] + iy);
    double x0 = __ldg(d.cd[0] + ix);
    double rr = sqrt(x0 * x0 + y0 * y0 + z0 * z0);
    double rvel = factor_vel * dev_get_rvel(d, npot, ix, iy, iz, x0, y0, z0, rr);
    float dz_rsd = (float)(rvel * clr_bg_v1(d, rr));
    unsigned long long gcell = (unsigned long long)ix + (unsigned long long)d.n * ((unsigned long long)iy + (unsigned long long)d.n * (iz + d.iz0_here));
    ClrStream s(seed, 2 + 2 * ipop, gcell);
    s.seek(3u * (uint32_t)ip);
    float px = (float)(x0 + dx * (s.next() - 0.5));
    float py = (float)(y0 + dx * (s.next() - 0.5));
    float pz = (float)(z0 + dx * (s.next() - 0.5));
    // vec2pix_ring (chealpix) on the float-rounded position, then ring2nest (srcs.c:265-270)
    double vx = px, vy = py, vz = pz;
    double vlen = sqrt(vx * vx + vy * vy + vz * vz);
    long long pr = clr_ang2pix_ring_zphi(d.nside_base, vz / vlen, atan2(vy, vx));
    pos[is] = make_float4(px, py, pz, dz_rsd);
    ipix[is] = clr_ring2nest(d.nside_base, (int)pr);
  }
}

// ---- srcs_get_local_properties_single (srcs.c:386-416) -----------------------------------------
__global__ void __launch_bounds__(kThreads)
local_props_kernel(const ClrDev d, const float4 *__restrict__ pos, float *__restrict__ srcs, long long nsrc)
{
  // the Src record is 9 floats (36 B): written directly, a warp store would touch 36 sectors for 128 useful
  // bytes, so the CTA assembles its 256 records in shared memory and streams them out as contiguous floats
  __shared__ float rec[kThreads * 9];
  const long long n_blocks = (nsrc + kThreads - 1) / kThreads;
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long i = blk * kThreads + threadIdx.x;
    if (i < nsrc) {
      float4 p = pos[i];
      double r, cth, phi;
      clr_cart2sph(p.x, p.y, p.z, &r, &cth, &phi);
      float *o = rec + 9 * threadIdx.x;
      o[0] = (float)(57.2957795 * phi);
      o[1] = (float)(90 - 57.2957795 * acos(cth));
      o[2] = (float)clr_bg_z(d, r);
      o[3] = p.w;
      o[4] = -1.f; o[5] = -1.f; o[6] = 0.f; o[7] = 0.f; o[8] = 0.f;
    }
    __syncthreads();
    const long long base = blk * kThreads * 9;
    const long long lim = (nsrc - blk * kThreads < kThreads ? nsrc - blk * kThreads : kThreads) * 9;
    for (int k = threadIdx.x; k < lim; k += kThreads) srcs[base + k] = rec[k];
    __syncthreads();
  }
}

// ---- RSD under beaming: CIC-interpolated potential gradient along the line of sight -----------
// beaming.c:31-117 (get_element, RETURN_VEL) + beaming.c:183-265 (trilinear branch)
__global__ void __launch_bounds__(kThreads)
beam_rsd_kernel(const ClrDev d, const float *__restrict__ npot, const float4 *__restrict__ pos, float *__restrict__ srcs,
                long long nsrc, int do_pre, int do_post)
{
  const double idx = (double)(d.n / d.l_box);
  const double factor_vel = -d.fgrowth_0 / (1.5 * d.hubble_0 * d.OmegaM);
  const bool whole_box = d.nz_here == d.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nsrc; i += (long long)gridDim.x * blockDim.x) {
    float4 p = pos[i];
    float *o = srcs + 9 * i;
    float acc = do_pre ? 0.f : o[3];
    float pp[3] = {p.x, p.y, p.z};
    float r2 = __fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z));
    double r = sqrt((double)r2);
    double ir = 1. / (r > 0.001 ? r : 0.001);
    double xn[3], u[3];
    long ix0[3], ix1[3];
    double h0x[3];
    float h1x[3];
    for (int ax = 0; ax < 3; ax++) {
      xn[ax] = (pp[ax] + d.pos_obs[ax]) * idx;
      u[ax] = pp[ax] * ir;
      ix0[ax] = (long)(xn[ax]);
      h0x[ax] = xn[ax] - ix0[ax];
      h1x[ax] = (float)(1 - h0x[ax]);
      ix1[ax] = ix0[ax] + 1;
      if (ix0[ax] >= d.n) ix0[ax] -= d.n; else if (ix0[ax] < 0) ix0[ax] += d.n;
      if (ix1[ax] >= d.n) ix1[ax] -= d.n; else if (ix1[ax] < 0) ix1[ax] += d.n;
    }
    // local plane of the two CIC corner planes. A source of this slab sits within half a cell of one of its planes, so
    // with several slabs the corners lie in [-1, nz_here] (periodic distance to the slab) and are served by the halo:
    // the reference sums exactly these contributions while the slabs rotate (beaming.c:325-352).
    long lz[2] = {ix0[2] - d.iz0_here, ix1[2] - d.iz0_here};
    if (!whole_box)
      for (int cz = 0; cz < 2; cz++) {
        if (lz[cz] > d.nz_here) lz[cz] -= d.n;           // e.g. rank 0, global plane n-1 = local -1
        else if (lz[cz] < -1) lz[cz] += d.n;             // last rank, global plane 0 = local nz_here
      }
    float v[3] = {0.f, 0.f, 0.f};
    bool added = false;
    for (int cz = 0; cz < 2; cz++) {
      long izc = lz[cz];
      if (whole_box ? (izc >= 0 && izc < d.nz_here) : (izc >= -1 && izc <= d.nz_here)) {
        float w4[4], v4[4][3];
        if (cz == 0) {
          w4[0] = h1x[2] * h1x[1] * h1x[0]; w4[1] = (float)(h1x[2] * h1x[1] * h0x[0]);
          w4[2] = (float)(h1x[2] * h0x[1] * h1x[0]); w4[3] = (float)(h1x[2] * h0x[1] * h0x[0]);
        } else {
          w4[0] = (float)(h0x[2] * h1x[1] * h1x[0]); w4[1] = (float)(h0x[2] * h1x[1] * h0x[0]);
          w4[2] = (float)(h0x[2] * h0x[1] * h1x[0]); w4[3] = (float)(h0x[2] * h0x[1] * h0x[0]);
        }
        added = true;
        dev_vel_element(d, npot, (int)ix0[0], (int)ix0[1], (int)izc, whole_box, v4[0]);
        dev_vel_element(d, npot, (int)ix1[0], (int)ix0[1], (int)izc, whole_box, v4[1]);
        dev_vel_element(d, npot, (int)ix0[0], (int)ix1[1], (int)izc, whole_box, v4[2]);
        dev_vel_element(d, npot, (int)ix1[0], (int)ix1[1], (int)izc, whole_box, v4[3]);
        for (int ax = 0; ax < 3; ax++)
          v[ax] += (v4[0][ax] * w4[0] + v4[1][ax] * w4[1] + v4[2][ax] * w4[2] + v4[3][ax] * w4[3]);
      }
    }
    if (added) {
      float vr = (float)(0.5 * idx * (v[0] * u[0] + v[1] * u[1] + v[2] * u[2]));
      acc += vr;
    }
    if (do_post) {
      double z = o[2];
      double rz = clr_r_of_z(d, z);
      double vg = clr_bg_v1(d, rz);
      acc = (float)((double)acc * (vg * factor_vel));
    }
    o[3] = acc;
    if (do_pre) { o[4] = 0.f; o[5] = 0.f; }
  }
}

// ---- srcs_distribute_single (srcs.c:296-373): route every source to rank ipix % NNodes, order preserved ----------
// Stable multi-split on the device: chunks of 256 sources, per-chunk counts per destination, offsets from a host
// scan of the (small) count table, then every source computes its slot from warp ballots.
constexpr int kDistThreads = 256;
__global__ void __launch_bounds__(kDistThreads)
dist_count_kernel(const int32_t *__restrict__ ipix, long long n, int nranks, int *__restrict__ cnt)
{
  __shared__ int h[CLR_MAX_PEERS];
  if (threadIdx.x < CLR_MAX_PEERS) h[threadIdx.x] = 0;
  __syncthreads();
  const long long i = (long long)blockIdx.x * kDistThreads + threadIdx.x;
  const int d = i < n ? ipix[i] % nranks : -1;
  for (int k = 0; k < nranks; k++) {
    const unsigned m = __ballot_sync(0xffffffffu, d == k);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&h[k], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < nranks) cnt[(long long)blockIdx.x * nranks + threadIdx.x] = h[threadIdx.x];
}
__global__ void __launch_bounds__(kDistThreads)
dist_scatter_kernel(const float4 *__restrict__ pos, const int32_t *__restrict__ ipix, long long n, int nranks,
                    const long long *__restrict__ off, float4 *__restrict__ spos, int32_t *__restrict__ sipix)
{
  __shared__ int wcnt[kDistThreads / 32][CLR_MAX_PEERS];
  const long long i = (long long)blockIdx.x * kDistThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int d = i < n ? ipix[i] % nranks : -1;
  int before = 0;
  for (int k = 0; k < nranks; k++) {
    const unsigned m = __ballot_sync(0xffffffffu, d == k);
    if (d == k) before = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[w][k] = __popc(m);
  }
  __syncthreads();
  if (d < 0) return;
  long long slot = off[(long long)blockIdx.x * nranks + d] + before;
  for (int ww = 0; ww < w; ww++) slot += wcnt[ww][d];
  spos[slot] = pos[i];
  sipix[slot] = ipix[i];
}
// dz_rsd of the Src records (after the beam estimator) back into the Cartesian catalogue, which is what travels
__global__ void rsd_to_pos_kernel(const float *__restrict__ srcs, float4 *__restrict__ pos, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pos[i].w = srcs[9 * i + 3];
}

int grid_for(clr_ctx *c, long long blocks, int per_sm)
{
  long long cap = (long long)c->sm_count * per_sm;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

int clr_srcs_run(clr_ctx *c, int ipop, uint32_t seed)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  CLR_CHECK(P.set, "srcs population %d has no tables (clr_set_srcs)", ipop);
  CLR_CHECK(P.have_norm, "srcs population %d has no normalisation (clr_compute_density_normalization)", ipop);
  const long long n_cells = (long long)c->dev.nz_here * c->dev.n * c->dev.n;
  const long long n_chunks = (n_cells + kChunk - 1) / kChunk;
  const long long n_super = (n_chunks + kSub - 1) / kSub;
  // the count buffer holds whole super-chunk slices (compact entries live in the slice of their super chunk)
  if (!P.d_counts) CLR_CUDA(cudaMalloc(&P.d_counts, (size_t)n_super * kSub * kChunk * sizeof(int32_t)));
  ClrPop pop{P.d_a, P.d_b, P.d_norm, P.norm_0, P.norm_f};
  if (!P.d_bound) CLR_CUDA(cudaMalloc(&P.d_bound, (size_t)2 * CLR_NA * 4 * sizeof(float)));   // per-cell + group table
  {
    StageScope sc(c, "srcs_bound", 1);
    const double dx = (double)(c->dev.l_box / c->dev.n);
    poisson_bound_kernel<<<(CLR_NA + 255) / 256, 256, 0, c->stream>>>(c->dev, pop, (float)(dx * dx * dx) * 1.0001f,
                                                                      reinterpret_cast<float4 *>(P.d_bound), 2);
    // group table: looked up at the smallest radius of 4 cells in a row, must reach 3 cells further out
    const int up = (int)ceil(3. * dx * c->p.glob_idr) + 3;
    poisson_bound_kernel<<<(CLR_NA + 255) / 256, 256, 0, c->stream>>>(c->dev, pop, (float)(dx * dx * dx) * 1.0001f,
                                                                      reinterpret_cast<float4 *>(P.d_bound) + CLR_NA, up);
    c->launches++;
    CLR_CUDA(cudaGetLastError());
  }
  long long total = 0;
  long long *d_offs = nullptr;
  bool compact = c->srcs_compact != 0;
  // entries per super chunk: kept with the catalogue (expansion, clr_srcs_get_counts). Written by the kernel itself: a
  // device-to-device cudaMemcpy on the main stream would queue behind the catalogue read-back on the copy engine.
  if (compact && (!P.d_sup_entries || P.sup_entries_cap < (size_t)n_super)) {
    if (P.d_sup_entries) cudaFree(P.d_sup_entries);
    P.d_sup_entries = nullptr;
    CLR_CUDA(cudaMalloc(&P.d_sup_entries, (size_t)n_super * sizeof(int32_t)));
    P.sup_entries_cap = (size_t)n_super;
  }
  for (int attempt = 0; attempt < 2; attempt++) {
    // scratch: offs[n_scan + 2] (exclusive scan, grand total, overflow flag) | tot[n_scan] | segment sums | entries per super chunk
    const long long n_scan = compact ? n_super : n_chunks;
    size_t need = (size_t)(n_scan + 2) * sizeof(long long) + (size_t)(n_scan + 2) / 2 * sizeof(long long) +
                  (size_t)kScanBlocks * sizeof(long long) + (size_t)(n_super + 2) / 2 * sizeof(long long) + 64;
    if (clr_ensure_scratch(c, need)) return 1;
    d_offs = reinterpret_cast<long long *>(c->d_scratch);
    int32_t *d_tot = reinterpret_cast<int32_t *>(d_offs + n_scan + 2);
    long long *d_seg = d_offs + n_scan + 2 + (n_scan + 2) / 2;
    int32_t *d_sup_entries = P.d_sup_entries;
    zero_words_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<uint32_t *>(d_offs + n_scan + 1), 2);   // overflow flag
    c->launches++;
    {
      StageScope sc(c, "srcs_poisson", 1);
      const int grid = grid_for(c, n_super, 8);
      if (compact)
        poisson_kernel<true><<<grid, kThreads, 0, c->stream>>>(
            c->dev, c->d_dens, pop, reinterpret_cast<const float4 *>(P.d_bound), reinterpret_cast<const float4 *>(P.d_bound) + CLR_NA,
            seed, ipop, P.d_counts, d_tot, d_sup_entries, reinterpret_cast<int *>(d_offs + n_scan + 1), n_cells);
      else
        poisson_kernel<false><<<grid, kThreads, 0, c->stream>>>(
            c->dev, c->d_dens, pop, reinterpret_cast<const float4 *>(P.d_bound), reinterpret_cast<const float4 *>(P.d_bound) + CLR_NA,
            seed, ipop, P.d_counts, d_tot, d_sup_entries, reinterpret_cast<int *>(d_offs + n_scan + 1), n_cells);
      CLR_CUDA(cudaGetLastError());
    }
    {
      StageScope sc(c, "srcs_scan", 2);
      scan_sums_kernel<<<kScanBlocks, 256, 0, c->stream>>>(d_tot, d_seg, n_scan);
      scan_chunks_kernel<<<kScanBlocks, 256, 0, c->stream>>>(d_tot, d_seg, d_offs, n_scan);
      CLR_CUDA(cudaGetLastError());
    }
    long long res[2] = {0, 0};
    if (clr_read_small(c, res, d_offs + n_scan, sizeof(res))) return 1;
    total = res[0];
    P.counts_compact = compact;
    P.d_sup_entries_valid = false;
    if (compact && res[1] != 0) { compact = false; continue; }       // a cell holds more than 65535 sources: dense counts
    if (compact) P.d_sup_entries_valid = true;
    break;
  }
  P.nsrc = total;
  if ((size_t)total > P.cap_src) {
    if (c->copy_pending) CLR_CUDA(cudaStreamSynchronize(c->copy_stream));   // a read-back may still use the old buffers
    c->copy_pending = false; c->buf_busy[0] = c->buf_busy[1] = false;
    if (P.d_pos) cudaFree(P.d_pos);
    if (P.d_ipix) cudaFree(P.d_ipix);
    if (P.d_srcs) cudaFree(P.d_srcs);
    if (P.d_srcs_alt) cudaFree(P.d_srcs_alt);
    P.d_pos = nullptr; P.d_ipix = nullptr; P.d_srcs = nullptr; P.d_srcs_alt = nullptr;
    size_t cap = (size_t)total + (size_t)total / 16 + 1024;
    CLR_CUDA(cudaMalloc(&P.d_pos, cap * 4 * sizeof(float)));
    CLR_CUDA(cudaMalloc(&P.d_ipix, cap * sizeof(int32_t)));
    CLR_CUDA(cudaMalloc(&P.d_srcs, cap * 9 * sizeof(float)));
    P.cap_src = cap;
  }
  if (c->async_results) {
    // two Src buffers: this run fills the one the LAST read-back did not use, so that copy may take the whole
    // of this run; only the read-back of two runs ago (same buffer) has to be finished, checked on the device
    if (!P.d_srcs_alt) CLR_CUDA(cudaMalloc(&P.d_srcs_alt, P.cap_src * 9 * sizeof(float)));
    std::swap(P.d_srcs, P.d_srcs_alt);
    P.srcs_buf ^= 1;
    if (c->buf_busy[P.srcs_buf]) CLR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_buf_free[P.srcs_buf], 0));
    c->buf_busy[P.srcs_buf] = false;
  } else if (c->copy_pending) {
    CLR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
    c->copy_pending = false;
  }
  if (clr_npot_ready(c)) return 1;               // placement reads the potential (RSD) and its z halo
  if (total > 0) {
    // the 9-float Src buffer is not written until clr_srcs_local: borrow it for the source references
    unsigned long long *d_ref = reinterpret_cast<unsigned long long *>(P.d_srcs);
    {
      StageScope sc(c, "srcs_expand", 1);
      if (P.counts_compact)
        expand_entries_kernel<<<grid_for(c, n_super, 8), kThreads, 0, c->stream>>>(P.d_counts, P.d_sup_entries, d_offs, d_ref, n_super);
      else
        expand_kernel<<<grid_for(c, n_chunks, 8), kThreads, 0, c->stream>>>(P.d_counts, d_offs, d_ref, n_cells);
      CLR_CUDA(cudaGetLastError());
    }
    StageScope sc(c, "srcs_place", 1);
    place_src_kernel<<<grid_for(c, (total + kThreads - 1) / kThreads, 8), kThreads, 0, c->stream>>>(
        c->dev, c->d_npot, d_ref, seed, ipop, reinterpret_cast<float4 *>(P.d_pos), P.d_ipix, total);
    CLR_CUDA(cudaGetLastError());
  }
  return 0;
}

int clr_srcs_local(clr_ctx *c, int ipop)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  if (P.nsrc == 0) return 0;
  StageScope sc(c, "srcs_local", 1);
  local_props_kernel<<<grid_for(c, (P.nsrc + kThreads - 1) / kThreads, 8), kThreads, 0, c->stream>>>(
      c->dev, reinterpret_cast<const float4 *>(P.d_pos), P.d_srcs, P.nsrc);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

int clr_srcs_beam(clr_ctx *c, int ipop)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  if (P.nsrc == 0) return 0;
  StageScope sc(c, "srcs_beam_rsd", 1);
  beam_rsd_kernel<<<grid_for(c, (P.nsrc + kThreads - 1) / kThreads, 8), kThreads, 0, c->stream>>>(
      c->dev, c->d_npot, reinterpret_cast<const float4 *>(P.d_pos), P.d_srcs, P.nsrc, 1, 1);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

// srcs_distribute_single (srcs.c:296-373). After the call this rank holds the sources whose base pixel satisfies
// ipix % nranks == rank, in the reference's order: blocks from rank-1, rank-2, ... (mod nranks), own sources last,
// the order inside a block = the sender's catalogue order. beam_first: evaluate the RSD-under-beaming estimator
// (srcs.c:486-504) on the slab that still holds the potential around each source and carry it in pos[3].
int clr_srcs_distribute_impl(clr_ctx *c, int ipop, int beam_first, long long *nsrc_out)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  const int R = c->nranks;
  if (beam_first) {
    if (clr_npot_ready(c)) return 1;
    if (clr_srcs_local(c, ipop)) return 1;
    if (clr_srcs_beam(c, ipop)) return 1;
    if (P.nsrc > 0) {
      rsd_to_pos_kernel<<<grid_for(c, (P.nsrc + kThreads - 1) / kThreads, 8), kThreads, 0, c->stream>>>(
          P.d_srcs, reinterpret_cast<float4 *>(P.d_pos), P.nsrc);
      CLR_CUDA(cudaGetLastError());
      c->launches++;
    }
  }
  if (R == 1) { if (nsrc_out) *nsrc_out = P.nsrc; return 0; }
  StageScope sc(c, "srcs_distribute", 2);
  const long long n = P.nsrc;
  const long long n_chunks = (n + kDistThreads - 1) / kDistThreads;
  std::vector<int> h_cnt((size_t)n_chunks * R);
  std::vector<long long> h_off((size_t)n_chunks * R);
  int *d_cnt = nullptr; long long *d_off = nullptr;
  float4 *d_spos = nullptr; int32_t *d_sipix = nullptr;
  int rc = 1;
  std::vector<unsigned long long> mat((size_t)R * R, 0ULL);
  do {
    if (n > 0) {
      if (cudaMalloc(&d_cnt, h_cnt.size() * sizeof(int)) != cudaSuccess) { clr_set_error("srcs_distribute: out of device memory"); break; }
      if (cudaMalloc(&d_off, h_off.size() * sizeof(long long)) != cudaSuccess) { clr_set_error("srcs_distribute: out of device memory"); break; }
      if (cudaMalloc(&d_spos, (size_t)n * sizeof(float4)) != cudaSuccess) { clr_set_error("srcs_distribute: out of device memory"); break; }
      if (cudaMalloc(&d_sipix, (size_t)n * sizeof(int32_t)) != cudaSuccess) { clr_set_error("srcs_distribute: out of device memory"); break; }
      dist_count_kernel<<<(unsigned)n_chunks, kDistThreads, 0, c->stream>>>(P.d_ipix, n, R, d_cnt);
      if (cudaMemcpyAsync(h_cnt.data(), d_cnt, h_cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) break;
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) break;
    }
    // destination-major offsets: start of destination d + sources of earlier chunks going to d
    std::vector<long long> tot(R, 0), start(R, 0);
    for (long long ch = 0; ch < n_chunks; ch++)
      for (int d = 0; d < R; d++) tot[d] += h_cnt[(size_t)ch * R + d];
    for (int d = 1; d < R; d++) start[d] = start[d - 1] + tot[d - 1];
    {
      std::vector<long long> run(start);
      for (long long ch = 0; ch < n_chunks; ch++)
        for (int d = 0; d < R; d++) { h_off[(size_t)ch * R + d] = run[d]; run[d] += h_cnt[(size_t)ch * R + d]; }
    }
    if (n > 0) {
      if (cudaMemcpyAsync(d_off, h_off.data(), h_off.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) break;
      dist_scatter_kernel<<<(unsigned)n_chunks, kDistThreads, 0, c->stream>>>(reinterpret_cast<const float4 *>(P.d_pos), P.d_ipix, n, R,
                                                                              d_off, d_spos, d_sipix);
      if (cudaGetLastError() != cudaSuccess) break;
    }
    // transfer matrix (MPI_Allgather of ns_to_nodes, srcs.c:311-316): row = sender
    if (clr_ensure_scratch(c, mat.size() * sizeof(unsigned long long))) break;
    for (int d = 0; d < R; d++) mat[(size_t)c->rank * R + d] = (unsigned long long)tot[d];
    if (cudaMemcpyAsync(c->d_scratch, mat.data(), mat.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) break;
    if (clr_comm_allreduce_u64(c, reinterpret_cast<unsigned long long *>(c->d_scratch), mat.size())) break;
    if (clr_read_small(c, mat.data(), c->d_scratch, mat.size() * sizeof(unsigned long long))) break;
    // receive layout: blocks from rank-1, rank-2, ..., rank (srcs.c:347-352)
    long long n_new = 0;
    std::vector<size_t> roff(R), rn(R), soff(R), sn(R);
    for (int ii = 0; ii < R; ii++) {
      const int from = ((c->rank - 1 - ii) % R + R) % R;
      roff[from] = (size_t)n_new;
      rn[from] = (size_t)mat[(size_t)from * R + c->rank];
      n_new += (long long)rn[from];
    }
    for (int d = 0; d < R; d++) { soff[d] = (size_t)start[d]; sn[d] = (size_t)tot[d]; }
    // new catalogue buffers
    float *n_pos = nullptr; int32_t *n_ipix = nullptr; float *n_srcs = nullptr;
    const size_t cap = (size_t)n_new + (size_t)n_new / 16 + 1024;
    if (cudaMalloc(&n_pos, cap * 4 * sizeof(float)) != cudaSuccess || cudaMalloc(&n_ipix, cap * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc(&n_srcs, cap * 9 * sizeof(float)) != cudaSuccess) { clr_set_error("srcs_distribute: out of device memory"); break; }
    // own block: device copy; the rest: exact-size exchange (4 floats per position, 1 word per pixel index)
    if (sn[c->rank]) {
      cudaMemcpyAsync(n_pos + roff[c->rank] * 4, d_spos + soff[c->rank], sn[c->rank] * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream);
      cudaMemcpyAsync(n_ipix + roff[c->rank], d_sipix + soff[c->rank], sn[c->rank] * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream);
    }
    std::vector<size_t> so4(R), sn4(R), ro4(R), rn4(R);
    for (int d = 0; d < R; d++) { so4[d] = soff[d] * 4; sn4[d] = sn[d] * 4; ro4[d] = roff[d] * 4; rn4[d] = rn[d] * 4; }
    if (clr_comm_alltoallv(c, reinterpret_cast<const float *>(d_spos), so4.data(), sn4.data(), n_pos, ro4.data(), rn4.data())) break;
    if (clr_comm_alltoallv(c, reinterpret_cast<const float *>(d_sipix), soff.data(), sn.data(), reinterpret_cast<float *>(n_ipix),
                           roff.data(), rn.data())) break;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) break;
    if (c->copy_pending) cudaStreamSynchronize(c->copy_stream);
    c->copy_pending = false; c->buf_busy[0] = c->buf_busy[1] = false;
    cudaFree(P.d_pos); cudaFree(P.d_ipix); cudaFree(P.d_srcs); cudaFree(P.d_srcs_alt);
    P.d_pos = n_pos; P.d_ipix = n_ipix; P.d_srcs = n_srcs; P.d_srcs_alt = nullptr;
    P.cap_src = cap; P.nsrc = n_new;
    rc = 0;
  } while (0);
  cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_spos); cudaFree(d_sipix);
  if (rc == 0) {
    if (clr_srcs_local(c, ipop)) return 1;           // Src records of the new catalogue (dz_rsd = pos[3])
    if (nsrc_out) *nsrc_out = P.nsrc;
  }
  return rc;
}

// dense per-cell counts on the device, unpadded [nz][n][n] (the reference's nsources array, srcs.c:125); *owned = the
// caller frees the buffer
int clr_srcs_dense_counts(clr_ctx *c, int ipop, int32_t **d_dense, bool *owned)
{
  clr_ctx::Pop &P = c->srcs[ipop];
  *owned = false;
  if (!P.counts_compact) { *d_dense = P.d_counts; return 0; }
  CLR_CHECK(P.d_sup_entries_valid, "no counts for population %d", ipop);
  const long long n_cells = (long long)c->dev.nz_here * c->dev.n * c->dev.n;
  const long long n_super = ((n_cells + kChunk - 1) / kChunk + kSub - 1) / kSub;
  int32_t *buf = nullptr;
  CLR_CUDA(cudaMalloc(&buf, (size_t)n_super * kSub * kChunk * sizeof(int32_t)));
  entries_to_counts_kernel<<<grid_for(c, n_super, 8), kThreads, 0, c->stream>>>(P.d_counts, P.d_sup_entries, buf, n_super, n_cells);
  if (cudaGetLastError() != cudaSuccess) { cudaFree(buf); clr_set_error("entries_to_counts_kernel failed"); return 1; }
  c->launches++;
  *d_dense = buf; *owned = true;
  return 0;
}
