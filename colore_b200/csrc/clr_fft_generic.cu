// 3-D c2r / r2c FFT for grid sizes that are NOT powers of two (the reference takes any n_grid through FFTW,
// fourier.c:81-125): mixed-radix Stockham autosort passes in shared memory, radices 4, 2, 3, 5, 7 in registers and any
// other prime factor up to 31 as a direct DFT. Same transform definition, pass order and in-place padded layout as the
// power-of-two plans of clr_fft.cu (complex passes over z and y, half-complex pass over x last / first; the imaginary
// parts of the x-DC and x-Nyquist lines are dropped by the c2r like FFTW's rdft2). This is the general path, not the
// fast one: the power-of-two sizes of BASELINE.json never come here. One GPU only.
#include "clr_internal.cuh"
#include <vector>

namespace {

constexpr int kThreads = 256;
constexpr int kMaxFactors = 16;
constexpr int kMaxRadix = 31;

struct Factors { int nf; int r[kMaxFactors]; };

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// w^k of the length-`wn` master table exp(+2 pi i k / wn); S = -1 conjugates (forward transform)
__device__ __forceinline__ float2 tw(const float2 *__restrict__ W, int k, int S)
{
  float2 w = __ldg(W + k);
  if (S < 0) w.y = -w.y;
  return w;
}

// One Stockham stage of radix R on T interleaved lines: in / out are [point][T]. Ns = product of the radices done so far.
// idx = j * T + line, j in [0, M/R). wstep = wn / M (the master table may belong to a multiple of M).
template <int R>
__device__ __forceinline__ void stage_fixed(const float2 *__restrict__ in, float2 *__restrict__ out, int M, int T, int Ns, int S,
                                            const float2 *__restrict__ W, int wstep)
{
  const int nb = M / R;
  for (int idx = threadIdx.x; idx < nb * T; idx += blockDim.x) {
    const int j = idx / T, line = idx - j * T;
    const int k = j % Ns;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      v[r] = in[(j + r * nb) * T + line];
      if (r && k) v[r] = cmulf(v[r], tw(W, (int)(((long long)r * k * (M / (Ns * R))) % M) * wstep, S));
    }
    // DFT of length R, sign S
    float2 o[R];
    if constexpr (R == 2) {
      o[0] = make_float2(v[0].x + v[1].x, v[0].y + v[1].y);
      o[1] = make_float2(v[0].x - v[1].x, v[0].y - v[1].y);
    } else if constexpr (R == 4) {
      float2 a = make_float2(v[0].x + v[2].x, v[0].y + v[2].y), b = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
      float2 c = make_float2(v[1].x + v[3].x, v[1].y + v[3].y), d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
      float2 id = S > 0 ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
      o[0] = make_float2(a.x + c.x, a.y + c.y); o[2] = make_float2(a.x - c.x, a.y - c.y);
      o[1] = make_float2(b.x + id.x, b.y + id.y); o[3] = make_float2(b.x - id.x, b.y - id.y);
    } else {
#pragma unroll
      for (int q = 0; q < R; q++) {
        float2 acc = v[0];
#pragma unroll
        for (int r = 1; r < R; r++) {
          const float2 t = cmulf(v[r], tw(W, ((r * q) % R) * (M / R) * wstep, S));
          acc.x += t.x; acc.y += t.y;
        }
        o[q] = acc;
      }
    }
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) out[(j0 + r * Ns) * T + line] = o[r];
  }
}

// any prime radix up to kMaxRadix: direct DFT with run-time loops
__device__ __forceinline__ void stage_any(const float2 *__restrict__ in, float2 *__restrict__ out, int M, int T, int Ns, int R, int S,
                                          const float2 *__restrict__ W, int wstep)
{
  const int nb = M / R;
  for (int idx = threadIdx.x; idx < nb * T; idx += blockDim.x) {
    const int j = idx / T, line = idx - j * T;
    const int k = j % Ns;
    float2 v[kMaxRadix];
    for (int r = 0; r < R; r++) {
      v[r] = in[(j + r * nb) * T + line];
      if (r && k) v[r] = cmulf(v[r], tw(W, (int)(((long long)r * k * (M / (Ns * R))) % M) * wstep, S));
    }
    const int j0 = (j / Ns) * Ns * R + k;
    for (int q = 0; q < R; q++) {
      float2 acc = v[0];
      for (int r = 1; r < R; r++) {
        float2 t = cmulf(v[r], tw(W, ((r * q) % R) * nb * wstep, S));
        acc.x += t.x; acc.y += t.y;
      }
      out[(j0 + q * Ns) * T + line] = acc;
    }
  }
}

// all stages of a length-M transform on T lines held in `a` ([point][T]); returns the buffer that holds the result
__device__ __forceinline__ float2 *stockham(float2 *a, float2 *b, int M, int T, int S, const Factors &f, const float2 *__restrict__ W,
                                            int wstep)
{
  int Ns = 1;
  for (int s = 0; s < f.nf; s++) {
    const int R = f.r[s];
    __syncthreads();
    switch (R) {
      case 2: stage_fixed<2>(a, b, M, T, Ns, S, W, wstep); break;
      case 3: stage_fixed<3>(a, b, M, T, Ns, S, W, wstep); break;
      case 4: stage_fixed<4>(a, b, M, T, Ns, S, W, wstep); break;
      case 5: stage_fixed<5>(a, b, M, T, Ns, S, W, wstep); break;
      case 7: stage_fixed<7>(a, b, M, T, Ns, S, W, wstep); break;
      default: stage_any(a, b, M, T, Ns, R, S, W, wstep); break;
    }
    Ns *= R;
    float2 *t = a; a = b; b = t;
  }
  __syncthreads();
  return a;
}

// strided pass: `n_tiles` tiles of T consecutive complex numbers (line index = contiguous index), points `stride` apart;
// tile t of batch bb starts at g + bb * batch_stride + t * T
__global__ void __launch_bounds__(kThreads)
gen_strided_kernel(float2 *__restrict__ g, const float2 *__restrict__ W, int M, int T, int S, Factors f, long long stride,
                   long long batch_stride, int tiles_per_batch, long long n_tiles)
{
  extern __shared__ float2 sm[];
  float2 *a = sm, *b = sm + (size_t)M * T;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long bb = t / tiles_per_batch;
    float2 *base = g + bb * batch_stride + (t - bb * tiles_per_batch) * T;
    __syncthreads();
    for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
      const int p = idx / T, line = idx - p * T;
      a[idx] = base[p * stride + line];
    }
    float2 *res = stockham(a, b, M, T, S, f, W, 1);
    for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
      const int p = idx / T, line = idx - p * T;
      base[p * stride + line] = res[idx];
    }
  }
}

// x pass of the c2r: a row of n/2+1 complex numbers -> n reals in place, through a complex transform of length H = n/2
// (Z[k] = (X[k] + conj X[H-k]) + i w^k (X[k] - conj X[H-k]), x[2m] + i x[2m+1] = sum_k Z[k] w^{2mk}); optional scaling
// and {sum, sum of squares} of the output (compute_sigma_dens, fourier.c:24-79)
__global__ void __launch_bounds__(kThreads)
gen_c2r_x_kernel(float2 *__restrict__ g, const float2 *__restrict__ W, int n, Factors f, long long n_rows, int pitch_c, float norm,
                 double *__restrict__ mom)
{
  extern __shared__ float2 sm[];
  const int H = n / 2;
  float2 *a = sm, *b = sm + H, *x = sm + 2 * H;       // x: the H+1 input modes of the row
  double s1 = 0, s2 = 0;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    float2 *base = g + row * pitch_c;
    __syncthreads();
    for (int k = threadIdx.x; k <= H; k += blockDim.x) {
      float2 v = base[k];
      if (k == 0 || k == H) v.y = 0.f;
      x[k] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
      const float2 xa = x[k], xb = x[H - k];
      const float2 e = make_float2(xa.x + xb.x, xa.y - xb.y), o = make_float2(xa.x - xb.x, xa.y + xb.y);
      const float2 wo = cmulf(o, tw(W, k, +1));
      a[k] = make_float2(e.x - wo.y, e.y + wo.x);
    }
    float2 *res = stockham(a, b, H, 1, +1, f, W, 2);
    for (int m = threadIdx.x; m < H; m += blockDim.x) {
      const float2 v = make_float2(res[m].x * norm, res[m].y * norm);
      base[m] = v;
      s1 += (double)v.x + (double)v.y;
      s2 += (double)v.x * v.x + (double)v.y * v.y;
    }
  }
  if (mom) {
    s1 = clr_warp_sum(s1); s2 = clr_warp_sum(s2);
    __shared__ double red[2][kThreads / 32];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t1 = 0, t2 = 0;
      for (int w = 0; w < kThreads / 32; w++) { t1 += red[0][w]; t2 += red[1][w]; }
      atomicAdd(mom, t1); atomicAdd(mom + 1, t2);
    }
  }
}

// x pass of the r2c: n reals -> n/2+1 complex numbers in place (Z = FFT_H(x[2m] + i x[2m+1]),
// X[k] = (Z[k] + conj Z[H-k]) / 2 + w^-k (Z[k] - conj Z[H-k]) / (2i))
__global__ void __launch_bounds__(kThreads)
gen_r2c_x_kernel(float2 *__restrict__ g, const float2 *__restrict__ W, int n, Factors f, long long n_rows, int pitch_c)
{
  extern __shared__ float2 sm[];
  const int H = n / 2;
  float2 *a = sm, *b = sm + H;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    float2 *base = g + row * pitch_c;
    __syncthreads();
    for (int m = threadIdx.x; m < H; m += blockDim.x) a[m] = base[m];
    float2 *res = stockham(a, b, H, 1, -1, f, W, 2);
    for (int k = threadIdx.x; k <= H; k += blockDim.x) {
      const float2 za = res[k == H ? 0 : k], zb = res[k == 0 ? 0 : H - k];
      const float2 e = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
      const float2 d = make_float2(0.5f * (za.x - zb.x), 0.5f * (za.y + zb.y));     // (Z[k] - conj Z[H-k]) / 2
      const float2 o = make_float2(d.y, -d.x);                                       // / i
      const float2 wo = cmulf(o, tw(W, k, -1));
      base[k] = make_float2(e.x + wo.x, e.y + wo.y);
    }
  }
}

bool factorize(int m, Factors *f)
{
  f->nf = 0;
  while (m % 4 == 0) { if (f->nf == kMaxFactors) return false; f->r[f->nf++] = 4; m /= 4; }
  for (int p = 2; p <= kMaxRadix && m > 1; p++)
    while (m % p == 0) { if (f->nf == kMaxFactors) return false; f->r[f->nf++] = p; m /= p; }
  return m == 1;
}

int pick_tile(int M, size_t *smem)
{
  for (int T = 8; T >= 1; T >>= 1) {
    *smem = (size_t)2 * M * T * sizeof(float2);
    if (*smem <= 200 * 1024) return T;
  }
  return 0;
}

int strided_pass(clr_ctx *c, float2 *g, int S, const Factors &f, long long stride, long long batch_stride, int batches, int n_lines)
{
  const int M = c->dev.n;
  size_t smem;
  const int T = pick_tile(M, &smem);
  CLR_CHECK(T > 0, "n_grid=%d: a line does not fit shared memory", M);
  CLR_CUDA(cudaFuncSetAttribute(gen_strided_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles_per_batch = (n_lines + T - 1) / T;      // rows are padded to a multiple of 8 complex numbers
  const long long n_tiles = (long long)tiles_per_batch * batches;
  const int grid = (int)std::min<long long>(n_tiles, (long long)c->sm_count * 4);
  gen_strided_kernel<<<grid, kThreads, smem, c->stream>>>(g, c->d_twiddle, M, T, S, f, stride, batch_stride, tiles_per_batch, n_tiles);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// largest prime factor <= 31 and n a multiple of 4 (the pointwise kernels walk rows in groups of 4 cells)
bool clr_fft_generic_ok(int n)
{
  Factors f;
  return n >= 16 && n <= 4096 && n % 4 == 0 && factorize(n, &f) && factorize(n / 2, &f);
}

int clr_fft_generic_c2r(clr_ctx *c, float2 *g, float norm, double *mom)
{
  const int n = c->dev.n;
  const long long nc = c->dev.ncp;
  CLR_CHECK(c->nranks == 1, "n_grid=%d: only power-of-two grids run on several GPUs", n);
  Factors fn, fh;
  CLR_CHECK(factorize(n, &fn) && factorize(n / 2, &fh), "n_grid=%d has a prime factor above %d", n, kMaxRadix);
  { StageScope sc(c, "fft_z", 1); if (strided_pass(c, g, +1, fn, (long long)n * nc, 0, 1, (int)(n * nc))) return 1; }
  { StageScope sc(c, "fft_y", 1); if (strided_pass(c, g, +1, fn, nc, (long long)n * nc, n, n / 2 + 1)) return 1; }
  StageScope sc(c, "fft_x", 1);
  const size_t smem = (size_t)(3 * (n / 2) + 1) * sizeof(float2);
  CLR_CUDA(cudaFuncSetAttribute(gen_c2r_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long rows = (long long)n * n;
  gen_c2r_x_kernel<<<(int)std::min<long long>(rows, (long long)c->sm_count * 8), kThreads, smem, c->stream>>>(
      g, c->d_twiddle, n, fh, rows, (int)nc, norm, mom);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

int clr_fft_generic_r2c(clr_ctx *c, float2 *g)
{
  const int n = c->dev.n;
  const long long nc = c->dev.ncp;
  CLR_CHECK(c->nranks == 1, "n_grid=%d: only power-of-two grids run on several GPUs", n);
  Factors fn, fh;
  CLR_CHECK(factorize(n, &fn) && factorize(n / 2, &fh), "n_grid=%d has a prime factor above %d", n, kMaxRadix);
  {
    StageScope sc(c, "fft_x", 1);
    const size_t smem = (size_t)(2 * (n / 2)) * sizeof(float2);
    CLR_CUDA(cudaFuncSetAttribute(gen_r2c_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long rows = (long long)n * n;
    gen_r2c_x_kernel<<<(int)std::min<long long>(rows, (long long)c->sm_count * 8), kThreads, smem, c->stream>>>(
        g, c->d_twiddle, n, fh, rows, (int)nc);
    CLR_CUDA(cudaGetLastError());
  }
  { StageScope sc(c, "fft_y", 1); if (strided_pass(c, g, -1, fn, nc, (long long)n * nc, n, n / 2 + 1)) return 1; }
  StageScope sc(c, "fft_z", 1);
  return strided_pass(c, g, -1, fn, (long long)n * nc, 0, 1, (int)(n * nc));
}
