// Lagrangian perturbation theory density: lpt_1 (density.c:376-644), lpt_2 (density.c:646-1031) and the
// mass deposits pos_2_ngp / pos_2_cic / pos_2_tsc (density.c:37-188), single GPU.
//   r2c(delta) -> psi1_k = i k delta_k / (k^2 N^3)  [2LPT: + the six d_i psi1_j] -> c2r
//   2LPT: Upsilon = sum_{i<j} (psi_ii psi_jj - psi_ij^2) -> r2c -> psi2_k = -i k Upsilon_k/(k^2 N^3) -> c2r
//   x = q + D(r) psi1 + D2(r) psi2 (periodic wrap) -> deposit -> delta = n - 1
// Work arrays follow the reference (3 extra complex fields for 1LPT, 8 for 2LPT, density.c:380-391,
// 650-667); 1LPT writes its particles to three separate unpadded buffers instead of un-padding in place.
// Several GPUs (share_particles, density.c:191-374): a particle is needed by every rank that owns one of
// the z planes its deposit stencil touches. Every rank deposits its own particles into its own slab
// straight from the particle arrays, and ships only the particles whose stencil reaches another slab:
// count per destination -> all-reduced P x P count matrix -> pack (block-aggregated cursors) -> one grouped
// ncclSend/ncclRecv with exact sizes -> deposit of the received particles. Unlike the reference there is no
// lpt_buffer_fraction to tune: the buffers are sized from the counts.
// Compiled with -fmad=false (k-space factors follow the reference's double expressions).
#include "clr_internal.cuh"

namespace {

constexpr int kThreads = 256;

struct LptFields { float2 *cdisp[3]; float2 *cdigrad[6]; };

// density.c:408-440 (order 1) / 684-723 (order 2). One thread per mode, reference mode order.
__global__ void __launch_bounds__(kThreads)
lpt_kspace1_kernel(const ClrDev d, const float2 *dens_f, LptFields f, int order)
{
  const double dk = 2 * 3.14159265358979323846 / d.l_box;
  const double fftnorm = (double)d.n * (double)d.n * (double)d.n;
  const unsigned n_rows = (unsigned)d.n * (unsigned)d.nyl;
  const unsigned rpb = max(1u, 1024u / (unsigned)d.nc);
  const unsigned n_groups = (n_rows + rpb - 1) / rpb;
  for (unsigned grp = blockIdx.x; grp < n_groups; grp += gridDim.x)
    for (unsigned tl = threadIdx.x; tl < rpb * (unsigned)d.nc; tl += blockDim.x) {
      unsigned rl = tl / (unsigned)d.nc;
      int kk = (int)(tl - rl * (unsigned)d.nc);
      unsigned row = grp * rpb + rl;
      if (row >= n_rows) continue;
      int ii_true = (int)(row / (unsigned)d.nyl);
      int jj = d.ky0 + (int)(row - (unsigned)ii_true * (unsigned)d.nyl);
      long long idx = (long long)row * d.ncp + kk;
      double kv[3];
      kv[2] = (2 * ii_true <= d.n) ? ii_true * dk : -(d.n - ii_true) * dk;
      kv[1] = (2 * jj <= d.n) ? jj * dk : -(d.n - jj) * dk;
      kv[0] = kk * dk;
      double k_mod2 = fftnorm * (kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2]);
      float2 dkv = dens_f[idx];
      float2 cd[3];
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        if (k_mod2 <= 0) cd[ax] = make_float2(0.f, 0.f);
        else {
          // I*kv*(a+ib)/k2 : C99 product (0 + i kv)(a + i b) = (0*a - kv*b) + i(0*b + kv*a)
          double re = 0.0 * (double)dkv.x - kv[ax] * (double)dkv.y;
          double im = 0.0 * (double)dkv.y + kv[ax] * (double)dkv.x;
          cd[ax] = make_float2((float)(re / k_mod2), (float)(im / k_mod2));
        }
      }
      if (order == 2) {
        // cdigrad = I*kv[a]*cdisp[b] from the float-rounded displacements (density.c:717-722)
        const int ia[6] = {0, 1, 2, 1, 2, 2}, ib[6] = {0, 0, 0, 1, 1, 2};
#pragma unroll
        for (int q = 0; q < 6; q++) {
          double a = cd[ib[q]].x, b = cd[ib[q]].y, k = kv[ia[q]];
          f.cdigrad[q][idx] = make_float2((float)(0.0 * a - k * b), (float)(0.0 * b + k * a));
        }
      }
#pragma unroll
      for (int ax = 0; ax < 3; ax++) f.cdisp[ax][idx] = cd[ax];
    }
}

// density.c:779-811: psi2_k = -I*kv*Upsilon_k/k2 into cdigrad[0..2]
__global__ void __launch_bounds__(kThreads)
lpt_kspace2_kernel(const ClrDev d, LptFields f)
{
  const double dk = 2 * 3.14159265358979323846 / d.l_box;
  const double fftnorm = (double)d.n * (double)d.n * (double)d.n;
  const unsigned n_rows = (unsigned)d.n * (unsigned)d.nyl;
  const unsigned rpb = max(1u, 1024u / (unsigned)d.nc);
  const unsigned n_groups = (n_rows + rpb - 1) / rpb;
  for (unsigned grp = blockIdx.x; grp < n_groups; grp += gridDim.x)
    for (unsigned tl = threadIdx.x; tl < rpb * (unsigned)d.nc; tl += blockDim.x) {
      unsigned rl = tl / (unsigned)d.nc;
      int kk = (int)(tl - rl * (unsigned)d.nc);
      unsigned row = grp * rpb + rl;
      if (row >= n_rows) continue;
      int ii_true = (int)(row / (unsigned)d.nyl);
      int jj = d.ky0 + (int)(row - (unsigned)ii_true * (unsigned)d.nyl);
      long long idx = (long long)row * d.ncp + kk;
      double kv[3];
      kv[2] = (2 * ii_true <= d.n) ? ii_true * dk : -(d.n - ii_true) * dk;
      kv[1] = (2 * jj <= d.n) ? jj * dk : -(d.n - jj) * dk;
      kv[0] = kk * dk;
      double k_mod2 = fftnorm * (kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2]);
      float2 u = f.cdigrad[5][idx];
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        float2 o = make_float2(0.f, 0.f);
        if (k_mod2 > 0) {
          // -I*kv = (-0) + i(-kv): product with (a + i b) = (-0*a + kv*b) + i(-0*b - kv*a)
          double re = -0.0 * (double)u.x - (-kv[ax]) * (double)u.y;
          double im = -0.0 * (double)u.y + (-kv[ax]) * (double)u.x;
          o = make_float2((float)(re / k_mod2), (float)(im / k_mod2));
        }
        f.cdigrad[ax][idx] = o;
      }
    }
}

// density.c:742-759: second-order source term, float arithmetic left to right
__global__ void __launch_bounds__(kThreads)
lpt_upsilon_kernel(const ClrDev d, const float *g0, const float *g1, const float *g2, const float *g3, const float *g4, float *g5)
{
  const long long n_cells = (long long)d.nz_here * d.n * d.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    clr_cell(d, i, ix, iy, iz);
    long long idx = ((long long)iz * d.n + iy) * d.pitch + ix;
    float xx = g0[idx], xy = g1[idx], xz = g2[idx], yy = g3[idx], yz = g4[idx], zz = g5[idx];
    g5[idx] = xx * yy + xx * zz + yy * zz - xy * xy - xz * xz - yz * yz;
  }
}

// density.c:473-501 / 850-882: particle positions (unpadded x, y, z) and zeroing of the density grid
__global__ void __launch_bounds__(kThreads)
lpt_positions_kernel(const ClrDev d, float *dens, const float *d0, const float *d1, const float *d2, const float *e0,
                     const float *e1, const float *e2, float *px, float *py, float *pz, int order)
{
  const long long n_cells = (long long)d.nz_here * d.n * d.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    clr_cell(d, i, ix, iy, iz);
    long long idx = ((long long)iz * d.n + iy) * d.pitch + ix;
    float xv[3] = {__ldg(d.cf[0] + ix), __ldg(d.cf[1] + iy), __ldg(d.cf[2] + iz + d.iz0_here)};
    float r2 = __fadd_rn(__fadd_rn(__fmul_rn(xv[0], xv[0]), __fmul_rn(xv[1], xv[1])), __fmul_rn(xv[2], xv[2]));
    double r = sqrt((double)r2);
    double dg = clr_bg_d1(d, r);
    double d2g = order == 2 ? clr_lerp(d, r, d.d2_arr, __ldg(d.d2_arr), __ldg(d.d2_arr + CLR_NA - 1)) : 0.;
    float psi1[3] = {d0[idx], d1[idx], d2[idx]};
    float psi2[3] = {0.f, 0.f, 0.f};
    if (order == 2) { psi2[0] = e0[idx]; psi2[1] = e1[idx]; psi2[2] = e2[idx]; }
    float p[3];
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      double v = order == 2 ? (double)xv[ax] + dg * psi1[ax] + d2g * psi2[ax] + d.pos_obs[ax]
                            : (double)xv[ax] + dg * psi1[ax] + d.pos_obs[ax];
      float q = (float)v;
      if (q < 0) q += d.l_box;
      if (q >= d.l_box) q -= d.l_box;
      p[ax] = q;
    }
    px[i] = p[0]; py[i] = p[1]; pz[i] = p[2];
    dens[idx] = 0.f;
  }
}

// z planes (global, wrapped) touched by the deposit stencil of a particle at height z: the same
// expressions as in lpt_deposit_kernel below, so routing and deposit always agree
__device__ __forceinline__ int lpt_planes(float z, int interp, int n, float i_agrid, int (&pl)[3])
{
  if (interp == 0) {
    int i0 = (int)((double)(z * i_agrid) + 0.5);
    if (i0 >= n) i0 -= n;
    if (i0 < 0) i0 += n;
    pl[0] = i0;
    return 1;
  } else if (interp == 1) {
    float s = z * i_agrid;
    int i0 = (int)s, i1 = i0 + 1;
    if (i0 < 0) i0 += n;
    if (i1 < 0) i1 += n;
    if (i0 >= n) i0 -= n;
    if (i1 >= n) i1 -= n;
    pl[0] = i0; pl[1] = i1;
    return 2;
  }
  float s = z * i_agrid;
  int c0 = (int)(floorf((float)((double)s + 0.5)));
  int cm = c0 - 1, cp = c0 + 1;
  if (cm < 0) cm += n;
  if (c0 < 0) c0 += n;
  if (cp < 0) cp += n;
  if (cm >= n) cm -= n;
  if (c0 >= n) c0 -= n;
  if (cp >= n) cp -= n;
  pl[0] = cm; pl[1] = c0; pl[2] = cp;
  return 3;
}

// Routing of the particles whose stencil reaches another rank's slab (share_particles, density.c:191-374).
// PACK = false: cnt[h] += number of my particles rank h needs. PACK = true: copy them as (x,y,z) triplets
// into segment h of `send` (segment offsets `off`, running cursors `cur`). Slots are reserved once per
// block and destination, so the global atomics stay in the thousands.
constexpr int kMaxRanks = 64;
template <bool PACK>
__global__ void __launch_bounds__(kThreads)
lpt_route_kernel(const ClrDev d, const float *px, const float *py, const float *pz, long long np, int interp, int me,
                 int nranks, unsigned long long *cnt, const unsigned long long *off, unsigned long long *cur, float *send)
{
  __shared__ unsigned int s_cnt[kMaxRanks];
  __shared__ unsigned long long s_base[kMaxRanks];
  const float i_agrid = d.n / d.l_box;
  const int nzl = d.n / nranks;
  const long long chunk = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < np; i0 += chunk) {
    const long long i = i0 + threadIdx.x;
    for (int h = threadIdx.x; h < nranks; h += blockDim.x) s_cnt[h] = 0;
    __syncthreads();
    int dest[3], slot[3], nd = 0;
    float z = 0.f;
    if (i < np) {
      z = pz[i];
      int pl[3];
      int npl = lpt_planes(z, interp, d.n, i_agrid, pl);
      for (int k = 0; k < npl; k++) {
        int h = pl[k] / nzl;
        bool dup = h == me;
        for (int q = 0; q < nd; q++) dup = dup || dest[q] == h;
        if (!dup) { dest[nd] = h; slot[nd] = (int)atomicAdd(&s_cnt[h], 1u); nd++; }
      }
    }
    __syncthreads();
    for (int h = threadIdx.x; h < nranks; h += blockDim.x)
      if (s_cnt[h]) s_base[h] = atomicAdd(PACK ? &cur[h] : &cnt[h], (unsigned long long)s_cnt[h]);
    if (PACK) {
      __syncthreads();
      if (nd) {
        float x = px[i], y = py[i];
        for (int q = 0; q < nd; q++) {
          float *o = send + 3 * (off[dest[q]] + s_base[dest[q]] + slot[q]);
          o[0] = x; o[1] = y; o[2] = z;
        }
      }
    }
    __syncthreads();
  }
}

// density.c:37-175: one thread per particle, float atomics (the reference's deposit is serial; the sum
// order differs here, which moves the result by fp32 rounding only)
__global__ void __launch_bounds__(kThreads)
lpt_deposit_kernel(const ClrDev d, const float *px, const float *py, const float *pz, int stride, float *delta, long long np,
                   int interp)
{
  const float i_agrid = d.n / d.l_box;
  const long long ngx = d.pitch;
  const int n = d.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < np; i += (long long)gridDim.x * blockDim.x) {
    float x[3] = {px[i * stride], py[i * stride], pz[i * stride]};
    if (interp == 0) {
      int i0[3];
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        i0[ax] = (int)((double)(x[ax] * i_agrid) + 0.5);
        if (i0[ax] >= n) i0[ax] -= n;
        if (i0[ax] < 0) i0[ax] += n;
      }
      i0[2] -= d.iz0_here;
      if (i0[2] >= 0 && i0[2] < d.nz_here) atomicAdd(&delta[i0[0] + ngx * (i0[1] + (long long)n * i0[2])], 1.f);
    } else if (interp == 1) {
      int i0[3], i1[3];
      float a0[3], a1[3];
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        float s = x[ax] * i_agrid;
        i0[ax] = (int)s;
        a1[ax] = s - i0[ax];
        a0[ax] = 1 - a1[ax];
        i1[ax] = i0[ax] + 1;
        if (i0[ax] < 0) i0[ax] += n;
        if (i1[ax] < 0) i1[ax] += n;
        if (i0[ax] >= n) i0[ax] -= n;
        if (i1[ax] >= n) i1[ax] -= n;
      }
      i0[2] -= d.iz0_here; i1[2] -= d.iz0_here;
      if (i0[2] >= 0 && i0[2] < d.nz_here) {
        long long b = (long long)n * i0[2];
        atomicAdd(&delta[i0[0] + ngx * (i0[1] + b)], a0[0] * a0[1] * a0[2]);
        atomicAdd(&delta[i1[0] + ngx * (i0[1] + b)], a1[0] * a0[1] * a0[2]);
        atomicAdd(&delta[i0[0] + ngx * (i1[1] + b)], a0[0] * a1[1] * a0[2]);
        atomicAdd(&delta[i1[0] + ngx * (i1[1] + b)], a1[0] * a1[1] * a0[2]);
      }
      if (i1[2] >= 0 && i1[2] < d.nz_here) {
        long long b = (long long)n * i1[2];
        atomicAdd(&delta[i0[0] + ngx * (i0[1] + b)], a0[0] * a0[1] * a1[2]);
        atomicAdd(&delta[i1[0] + ngx * (i0[1] + b)], a1[0] * a0[1] * a1[2]);
        atomicAdd(&delta[i0[0] + ngx * (i1[1] + b)], a0[0] * a1[1] * a1[2]);
        atomicAdd(&delta[i1[0] + ngx * (i1[1] + b)], a1[0] * a1[1] * a1[2]);
      }
    } else {
      int ic[3][3];        // [axis][m,0,p]
      float w[3][3];
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        float s = x[ax] * i_agrid;
        int c0 = (int)(floorf((float)((double)s + 0.5)));
        float a = s - c0;
        w[ax][0] = (float)(0.5 * (0.5 - (double)a) * (0.5 - (double)a));
        w[ax][2] = (float)(0.5 * (0.5 + (double)a) * (0.5 + (double)a));
        w[ax][1] = (float)(0.75 - (double)(a * a));
        int cm = c0 - 1, cp = c0 + 1;
        if (cm < 0) cm += n;
        if (c0 < 0) c0 += n;
        if (cp < 0) cp += n;
        if (cm >= n) cm -= n;
        if (c0 >= n) c0 -= n;
        if (cp >= n) cp -= n;
        ic[ax][0] = cm; ic[ax][1] = c0; ic[ax][2] = cp;
      }
#pragma unroll
      for (int cz = 0; cz < 3; cz++) {
        int iz = ic[2][cz] - d.iz0_here;
        if (!(iz >= 0 && iz < d.nz_here)) continue;
#pragma unroll
        for (int cy = 0; cy < 3; cy++)
#pragma unroll
          for (int cx = 0; cx < 3; cx++)
            atomicAdd(&delta[ic[0][cx] + ngx * (ic[1][cy] + (long long)n * iz)], w[0][cx] * w[1][cy] * w[2][cz]);
      }
    }
  }
}

// density.c:610-623: delta = n * inv_dens - 1 with inv_dens = 1
__global__ void __launch_bounds__(kThreads)
lpt_finalize_kernel(const ClrDev d, float *dens)
{
  const long long n_cells = (long long)d.nz_here * d.n * d.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    clr_cell(d, i, ix, iy, iz);
    long long idx = ((long long)iz * d.n + iy) * d.pitch + ix;
    dens[idx] = (float)((double)(dens[idx] * 1.f) - 1.);
  }
}

int grid_for(clr_ctx *c, long long items, int per_sm)
{
  long long g = (items + kThreads - 1) / kThreads, cap = (long long)c->sm_count * per_sm;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int clr_lpt_run(clr_ctx *c, int order)
{
  CLR_CHECK(c->nranks <= kMaxRanks, "LPT: at most %d GPUs", kMaxRanks);
  CLR_CHECK(order == 1 || order == 2, "LPT order %d", order);
  c->lpt_sent = c->lpt_received = 0;
  CLR_CHECK(c->lpt_interp_type >= 0 && c->lpt_interp_type <= 2, "Wrong interpolation type\n");
  const ClrDev &d = c->dev;
  const size_t slab = (size_t)d.pitch * d.n * d.nz_here * sizeof(float);
  const long long n_cells = (long long)d.nz_here * d.n * d.n;
  const long long n_rowgroups = ((long long)d.n * d.nyl + std::max(1, 1024 / d.nc) - 1) / std::max(1, 1024 / d.nc);
  float *buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float *pos[3] = {nullptr, nullptr, nullptr};
  int nbuf = order == 1 ? 3 : 8;
  float *d_send = nullptr, *d_recv = nullptr;
  auto cleanup = [&]() { for (int i = 0; i < 8; i++) cudaFree(buf[i]); cudaFree(d_send); cudaFree(d_recv); };
  for (int i = 0; i < nbuf; i++)
    if (cudaMalloc(&buf[i], slab) != cudaSuccess) { cleanup(); clr_set_error("LPT: out of device memory (%d work fields)", nbuf); return 1; }
  LptFields f;
  float *disp[3], *digrad[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (order == 1) { for (int i = 0; i < 3; i++) disp[i] = buf[i]; }
  else {
    disp[0] = buf[0]; disp[1] = buf[1]; disp[2] = c->d_dens;          // density.c:666-667
    for (int i = 0; i < 6; i++) digrad[i] = buf[2 + i];
  }
  for (int i = 0; i < 3; i++) f.cdisp[i] = reinterpret_cast<float2 *>(disp[i]);
  for (int i = 0; i < 6; i++) f.cdigrad[i] = reinterpret_cast<float2 *>(digrad[i]);
  int rc = 1;
  do {
    if (clr_fft_r2c_impl(c, c->d_dens)) break;
    { StageScope sc(c, "lpt_kspace", 1);
      lpt_kspace1_kernel<<<grid_for(c, n_rowgroups * kThreads, 8), kThreads, 0, c->stream>>>(d, reinterpret_cast<float2 *>(c->d_dens), f, order); }
    if (order == 2) {
      bool bad = false;
      for (int i = 0; i < 6 && !bad; i++) bad = clr_fft_c2r_impl(c, digrad[i], 1.0, nullptr) != 0;
      if (bad) break;
      { StageScope sc(c, "lpt_upsilon", 1);
        lpt_upsilon_kernel<<<grid_for(c, n_cells, 8), kThreads, 0, c->stream>>>(d, digrad[0], digrad[1], digrad[2], digrad[3], digrad[4], digrad[5]); }
      if (clr_fft_r2c_impl(c, digrad[5])) break;
      { StageScope sc(c, "lpt_kspace", 1);
        lpt_kspace2_kernel<<<grid_for(c, n_rowgroups * kThreads, 8), kThreads, 0, c->stream>>>(d, f); }
    }
    bool bad = false;
    for (int i = 0; i < 3 && !bad; i++) {
      bad = clr_fft_c2r_impl(c, disp[i], 1.0, nullptr) != 0;
      if (!bad && order == 2) bad = clr_fft_c2r_impl(c, digrad[i], 1.0, nullptr) != 0;
    }
    if (bad) break;
    // particle buffers: digrad[3..5] for 2LPT (as the reference); separate unpadded buffers for 1LPT
    if (order == 2) { for (int i = 0; i < 3; i++) pos[i] = digrad[3 + i]; }
    else {
      for (int i = 0; i < 3; i++) cudaFree(c->d_lpt_pos[i]);
      for (int i = 0; i < 3; i++) c->d_lpt_pos[i] = nullptr;
      bool oom = false;
      for (int i = 0; i < 3 && !oom; i++) oom = cudaMalloc(&c->d_lpt_pos[i], (size_t)n_cells * sizeof(float)) != cudaSuccess;
      if (oom) { clr_set_error("LPT: out of device memory (particles)"); break; }
      for (int i = 0; i < 3; i++) pos[i] = c->d_lpt_pos[i];
    }
    { StageScope sc(c, "lpt_positions", 1);
      lpt_positions_kernel<<<grid_for(c, n_cells, 8), kThreads, 0, c->stream>>>(d, c->d_dens, disp[0], disp[1], disp[2], digrad[0], digrad[1],
                                                                               digrad[2], pos[0], pos[1], pos[2], order); }
    { StageScope sc(c, "lpt_deposit", 1);
      lpt_deposit_kernel<<<grid_for(c, n_cells, 8), kThreads, 0, c->stream>>>(d, pos[0], pos[1], pos[2], 1, c->d_dens, n_cells, c->lpt_interp_type); }
    if (c->nranks > 1) {
      // share_particles (density.c:191-374): ship the particles whose stencil reaches another slab
      const int P = c->nranks;
      // rank-local failures (allocation, launch) are agreed on before every collective: nobody is left waiting
      int ok = clr_ensure_scratch(c, (size_t)(P * P + 2 * P) * sizeof(unsigned long long)) == 0;
      unsigned long long *d_mat = reinterpret_cast<unsigned long long *>(c->d_scratch);   // P x P counts [src][dst]
      unsigned long long *d_off = d_mat + (size_t)P * P, *d_cur = d_off + P;
      if (ok) ok = cudaMemsetAsync(d_mat, 0, (size_t)(P * P + 2 * P) * sizeof(unsigned long long), c->stream) == cudaSuccess;
      if (ok) ok = cudaGetLastError() == cudaSuccess;        // the deposit / position kernels above
      if (clr_comm_all_ok(c, ok, "LPT particle routing")) break;
      { StageScope sc(c, "lpt_route", 1);
        lpt_route_kernel<false><<<grid_for(c, n_cells, 8), kThreads, 0, c->stream>>>(d, pos[0], pos[1], pos[2], n_cells, c->lpt_interp_type, c->rank, P,
                                                                                    d_mat + (size_t)c->rank * P, nullptr, nullptr, nullptr); }
      if (clr_comm_allreduce_u64(c, d_mat, (size_t)P * P)) break;
      std::vector<unsigned long long> mat((size_t)P * P);
      if (cudaMemcpyAsync(mat.data(), d_mat, mat.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) break;
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) { clr_set_error("LPT: routing failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
      std::vector<unsigned long long> s_off(P + 1, 0), r_off(P + 1, 0);
      for (int h = 0; h < P; h++) {
        s_off[h + 1] = s_off[h] + mat[(size_t)c->rank * P + h];
        r_off[h + 1] = r_off[h] + mat[(size_t)h * P + c->rank];
      }
      const unsigned long long n_send = s_off[P], n_recv = r_off[P];
      ok = 1;
      if (cudaMalloc(&d_send, (size_t)(3 * n_send + 3) * sizeof(float)) != cudaSuccess ||
          cudaMalloc(&d_recv, (size_t)(3 * n_recv + 3) * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        clr_set_error("LPT: out of device memory (particle exchange: %llu out, %llu in)", n_send, n_recv);
        ok = 0;
      }
      if (ok) ok = cudaMemcpyAsync(d_off, s_off.data(), P * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
      if (clr_comm_all_ok(c, ok, "LPT particle exchange")) break;
      { StageScope sc(c, "lpt_route", 1);
        lpt_route_kernel<true><<<grid_for(c, n_cells, 8), kThreads, 0, c->stream>>>(d, pos[0], pos[1], pos[2], n_cells, c->lpt_interp_type, c->rank, P,
                                                                                   nullptr, d_off, d_cur, d_send); }
      std::vector<size_t> so(P), sn(P), ro(P), rn(P);
      for (int h = 0; h < P; h++) { so[h] = 3 * s_off[h]; sn[h] = 3 * (s_off[h + 1] - s_off[h]); ro[h] = 3 * r_off[h]; rn[h] = 3 * (r_off[h + 1] - r_off[h]); }
      { StageScope sc(c, "lpt_exchange", 0);
        if (clr_comm_alltoallv(c, d_send, so.data(), sn.data(), d_recv, ro.data(), rn.data())) break; }
      c->lpt_sent = (long long)n_send; c->lpt_received = (long long)n_recv;
      if (n_recv) {
        StageScope sc(c, "lpt_deposit", 1);
        lpt_deposit_kernel<<<grid_for(c, (long long)n_recv, 8), kThreads, 0, c->stream>>>(d, d_recv, d_recv + 1, d_recv + 2, 3, c->d_dens, (long long)n_recv,
                                                                                         c->lpt_interp_type);
      }
    }
    { StageScope sc(c, "lpt_finalize", 1);
      lpt_finalize_kernel<<<grid_for(c, n_cells, 8), kThreads, 0, c->stream>>>(d, c->d_dens); }
    if (cudaGetLastError() != cudaSuccess) { clr_set_error("LPT kernel launch failed"); break; }
    if (order == 2 && c->keep_particles) {
      // keep a copy of the particles for write_lpt (io.c:619-695)
      bool oom = false;
      for (int i = 0; i < 3; i++) { cudaFree(c->d_lpt_pos[i]); c->d_lpt_pos[i] = nullptr; }
      for (int i = 0; i < 3 && !oom; i++) oom = cudaMalloc(&c->d_lpt_pos[i], (size_t)n_cells * sizeof(float)) != cudaSuccess;
      if (oom) { clr_set_error("LPT: out of device memory (particle copy)"); break; }
      for (int i = 0; i < 3; i++) cudaMemcpyAsync(c->d_lpt_pos[i], pos[i], (size_t)n_cells * sizeof(float), cudaMemcpyDeviceToDevice, c->stream);
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { clr_set_error("LPT: stream error %s", cudaGetErrorString(cudaGetLastError())); break; }
    rc = 0;
  } while (0);
  cleanup();
  if (!c->keep_particles) { for (int i = 0; i < 3; i++) { cudaFree(c->d_lpt_pos[i]); c->d_lpt_pos[i] = nullptr; } }
  return rc;
}

int clr_lpt_particles(clr_ctx *c, float *x, float *y, float *z)
{
  CLR_CHECK(c->d_lpt_pos[0], "no LPT particles kept (set option keep_particles before the density call)");
  size_t bytes = (size_t)c->dev.nz_here * c->dev.n * c->dev.n * sizeof(float);
  float *h[3] = {x, y, z};
  for (int i = 0; i < 3; i++) CLR_CUDA(cudaMemcpyAsync(h[i], c->d_lpt_pos[i], bytes, cudaMemcpyDeviceToHost, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
