// 3-D c2r / r2c FFT of the slab-local grid: hand-written Stockham autosort kernels for sm_100a.
// Replaces fftw_wrap_c2r / fftw_wrap_r2c (fourier.c:81-125). No cuFFT anywhere.
//
// Transform definition (FFTW manual, what the reference links against): unnormalised, c2r =
// backward (exp(+i...)), r2c = forward; half-complex last axis with n/2+1 entries; in place with
// real rows padded to 2*(n/2+1). The c2r runs complex passes over z and y first and the
// half-complex -> real pass over x last, so the imaginary parts of the x-DC and x-Nyquist lines
// are dropped exactly as FFTW's rdft2 does (the reference fills those planes non-Hermitian,
// fourier.c:325-345).
//
// Kernel design
//  * One pass per axis, each pass = batched 1-D complex FFTs held in shared memory:
//    Stockham autosort, radix-8 butterflies in registers (first stage radix 2/4/8 so that any
//    power of two fits), 8 points per virtual thread, V virtual threads per thread.
//  * Strided axes (y, z): a CTA owns a tile of T consecutive lines (T contiguous complex numbers
//    per point of the line, 64 B with T=8), lane <-> line so that global accesses are coalesced
//    along the contiguous index and shared-memory accesses are conflict free without padding.
//  * Contiguous axis (x): lane <-> point; shared rows padded by 1 complex every 8 to spread the
//    stride-8 first-stage stores over the banks. The real transform of length n runs as a complex
//    transform of length n/2 plus a pre-twiddle (c2r) / post-twiddle (r2c) step.
//  * Twiddles come from one master table exp(2*pi*i*k/n) built in double on the host; each CTA
//    gathers the per-stage, per-thread factors once into shared memory (persistent CTAs).
//  * The last pass of the c2r optionally fuses the (sqrt(2 pi)/L)^3 scaling and the sum / sum of
//    squares needed by compute_sigma_dens (fourier.c:24-79, 394-397).
//
// Kernels of the product path (each described where it is defined):
//    one GPU      fill_z_kernel (mode fill + z pass of both fields, n <= 1024) / fill_z_cluster_kernel (the same on
//                 CTA pairs with distributed shared memory, n = 2048) -> yx_fused_kernel (y + x passes through L2,
//                 TMA bulk loads); fft_strided_kernel + fft_c2r_x_kernel / fft_r2c_x_kernel for grids handed in by the
//                 caller, the r2c, n < 128 and n = 4096
//    several GPUs fill_peer_kernel (mode fill + z pass + slab transpose as NVLink peer stores, one field per launch) or
//                 fft_strided_kernel<PEER> (the same pass on a grid that already holds modes) -> y pass out of the
//                 staging buffer -> x pass
#include "clr_internal.cuh"
#include "clr_fill.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <utility>

namespace {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{ return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <int S> __device__ __forceinline__ float2 mul_i(float2 a)
{ return S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <int S> __device__ __forceinline__ void dft2(float2 &a, float2 &b)
{
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}
// X_k = sum_n x_n exp(S*2*pi*i*n*k/4), natural order in and out
template <int S> __device__ __forceinline__ void dft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3)
{
  float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = mul_i<S>(csub(x1, x3));
  x0 = cadd(a, c); x2 = csub(a, c); x1 = cadd(b, d); x3 = csub(b, d);
}
template <int S> __device__ __forceinline__ void dft8(float2 (&v)[8])
{
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  b1 = make_float2(h * (b1.x - S * b1.y), h * (b1.y + S * b1.x));     // * exp(S*i*pi/4)
  b2 = mul_i<S>(b2);                                                    // * exp(S*i*pi/2)
  b3 = make_float2(h * (-b3.x - S * b3.y), h * (-b3.y + S * b3.x));   // * exp(S*3*i*pi/4)
  dft4<S>(a0, a1, a2, a3);
  dft4<S>(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// exp(S*2*pi*i*m/32) for compile-time m (folded after unrolling)
__device__ __forceinline__ float cos32(int m)
{
  m &= 31;
  if (m > 16) m = 32 - m;
  switch (m) {
    case 0: return 1.f;
    case 1: return 0.98078528040323044913f;
    case 2: return 0.92387953251128675613f;
    case 3: return 0.83146961230254523708f;
    case 4: return 0.70710678118654752440f;
    case 5: return 0.55557023301960222474f;
    case 6: return 0.38268343236508977173f;
    case 7: return 0.19509032201612826785f;
    case 8: return 0.f;
    case 9: return -0.19509032201612826785f;
    case 10: return -0.38268343236508977173f;
    case 11: return -0.55557023301960222474f;
    case 12: return -0.70710678118654752440f;
    case 13: return -0.83146961230254523708f;
    case 14: return -0.92387953251128675613f;
    case 15: return -0.98078528040323044913f;
    default: return -1.f;
  }
}
// a *= exp(S*2*pi*i*m/32)
template <int S> __device__ __forceinline__ float2 rot32(float2 a, int m)
{
  m &= 31;
  if (m == 0) return a;
  if (m == 8) return mul_i<S>(a);
  if (m == 16) return make_float2(-a.x, -a.y);
  if (m == 24) return mul_i<-S>(a);
  const float c = cos32(m), sn = S * cos32(m - 8);      // sin(x) = cos(x - pi/2)
  return make_float2(a.x * c - a.y * sn, a.x * sn + a.y * c);
}
// R = 4 * (R/4) Cooley-Tukey step in registers, natural order in and out
template <int S> __device__ __forceinline__ void dft16(float2 (&v)[16])
{
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) dft4<S>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
#pragma unroll
  for (int k1 = 1; k1 < 4; k1++)
#pragma unroll
    for (int n2 = 1; n2 < 4; n2++) v[4 * k1 + n2] = rot32<S>(v[4 * k1 + n2], 2 * n2 * k1);
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4<S>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // X[k1 + 4*k2] sits at v[4*k1 + k2]: transpose the 4x4 register tile
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = a + 1; b < 4; b++) { float2 t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
}
template <int S> __device__ __forceinline__ void dft32(float2 (&v)[32])
{
#pragma unroll
  for (int n2 = 0; n2 < 8; n2++) dft4<S>(v[n2], v[8 + n2], v[16 + n2], v[24 + n2]);
#pragma unroll
  for (int k1 = 1; k1 < 4; k1++)
#pragma unroll
    for (int n2 = 1; n2 < 8; n2++) v[8 * k1 + n2] = rot32<S>(v[8 * k1 + n2], n2 * k1);
  float2 o[32];
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    float2 t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = v[8 * k1 + i];
    dft8<S>(t);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) o[k1 + 4 * k2] = t[k2];
  }
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = o[i];
}
template <int R, int S> __device__ __forceinline__ void dftR(float2 (&t)[R])
{
  if constexpr (R == 8) dft8<S>(t);
  else if constexpr (R == 16) dft16<S>(t);
  else dft32<S>(t);
}

__host__ __device__ constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

// kernel-variant experiments (tools/fft_bench.py builds the library with -DCLR_FFT_VARIANT=k)
#ifndef CLR_FFT_VARIANT
#define CLR_FFT_VARIANT 0
#endif

// Stage plan of a length-M transform: radices (descending) R0*R1*R2 = M, every thread owns E = R0
// points of a line, so a 1024-point line is two radix-32 stages with ONE exchange through shared
// memory (the exchange traffic, not the flops, is what bounds a Stockham pass on this machine).
// kWide flags an alternative plan of the same length (FftPlan<2048 | kWide> = 32 x 8 x 8: 32 points per thread, half as
// many threads per line), used where the thread count of a tile must match another transform's (yx_fused_kernel)
constexpr int kWide = 1 << 20;
template <int MM> struct FftPlan {
  static constexpr int M = MM & (kWide - 1);
  static constexpr int NST = M <= 32 ? 1 : (M <= 1024 ? 2 : 3);
  static constexpr bool WIDE3 = ((MM & kWide) != 0 || CLR_FFT_VARIANT == 1 || CLR_FFT_VARIANT == 2) && M == 2048;   // 32 x 8 x 8
  static constexpr int R0 = M <= 32 ? M : (M == 64 ? 8 : (M == 128 || M == 256 ? 16 : (M <= 1024 || WIDE3 ? 32 : 16)));
  static constexpr int R1 = NST < 2 ? 1 : (M <= 128 ? 8 : (M <= 512 ? 16 : (M == 1024 ? 32 : (WIDE3 ? 8 : 16))));
  static constexpr int R2 = NST < 3 ? 1 : M / (R0 * R1);
  static constexpr int E = R0;                                              // points per thread
  static constexpr int TPL = M / E;                                         // threads per line
  static constexpr int PADSH = ilog2c(R0);
  static constexpr int LSTRIDE = M + (M >> PADSH);                          // padded line
  static constexpr int NB1 = (NST >= 2 ? (E / R1) * ilog2c(R1) : 0);        // base twiddles per thread, stage 1
  static constexpr int NB2 = (NST >= 3 ? (E / R2) * ilog2c(R2) : 0);
  static constexpr int NTW = (NB1 + NB2) * TPL;
  static_assert(R0 * R1 * R2 == M, "bad radix plan");
  __host__ __device__ static constexpr int radix(int s) { return s == 0 ? R0 : (s == 1 ? R1 : R2); }
  __host__ __device__ static constexpr int ns(int s) { return s == 0 ? 1 : (s == 1 ? R0 : R0 * R1); }
};

// shared-memory index of point p of line l. The pad of one slot every R0 points spreads the
// first-stage stores (stride R0 between neighbouring threads) over the banks in both layouts.
template <int M, bool STRIDED, int T> __device__ __forceinline__ int sidx(int p, int l)
{
  const int pp = p + (p >> FftPlan<M>::PADSH);
  return STRIDED ? pp * T + l : l * FftPlan<M>::LSTRIDE + pp;
}

// Base twiddles of this pass from the master table W[k] = exp(+2*pi*i*k/wn): for stage s >= 1 and
// butterfly jb the factors w^(2^q), w = exp(S*2*pi*i*(jb % Ns)/(Ns*R)). The other powers are products
// of at most log2(R) of these (tw_apply), so every factor stays within a few ulp of the exact value.
template <int M, int S> __device__ __forceinline__ void load_twiddles(float2 *tw, const float2 *__restrict__ W, int wn)
{
  using P = FftPlan<M>;
  for (int i = threadIdx.x; i < P::NTW; i += blockDim.x) {
    int slot = i / P::TPL, j = i % P::TPL;
    int s = slot < P::NB1 ? 1 : 2;
    if (s == 2) slot -= P::NB1;
    int R = s == 1 ? P::R1 : P::R2, Ns = s == 1 ? P::R0 : P::R0 * P::R1;
    int lg = s == 1 ? ilog2c(P::R1) : ilog2c(P::R2);
    int b = slot / lg, q = slot % lg;
    int jb = j + b * P::TPL;
    int t = ((jb % Ns) << q) * (wn / (Ns * R));
    float2 w = W[t];
    if (S < 0) w.y = -w.y;
    tw[i] = w;
  }
}

// t[r] *= w^r, r < R, from the base powers w^1, w^2, w^4, ... (stride `st` apart in shared memory)
template <int R> __device__ __forceinline__ void tw_apply(float2 (&t)[R], const float2 *base, int st)
{
  float2 lo[8];
  lo[1] = base[0]; lo[2] = base[st]; lo[4] = base[2 * st];
  lo[3] = cmul(lo[1], lo[2]); lo[5] = cmul(lo[1], lo[4]); lo[6] = cmul(lo[2], lo[4]); lo[7] = cmul(lo[3], lo[4]);
#pragma unroll
  for (int r = 1; r < 8; r++) t[r] = cmul(t[r], lo[r]);
  if constexpr (R >= 16) {
    float2 w8 = base[3 * st];
    t[8] = cmul(t[8], w8);
#pragma unroll
    for (int r = 1; r < 8; r++) t[8 + r] = cmul(t[8 + r], cmul(w8, lo[r]));
    if constexpr (R == 32) {
      float2 w16 = base[4 * st], w24 = cmul(w8, w16);
      t[16] = cmul(t[16], w16);
      t[24] = cmul(t[24], w24);
#pragma unroll
      for (int r = 1; r < 8; r++) {
        t[16 + r] = cmul(t[16 + r], cmul(w16, lo[r]));
        t[24 + r] = cmul(t[24 + r], cmul(w24, lo[r]));
      }
    }
  }
}

// One stage = B = E/R butterflies per thread on the register slots b + B*r; outputs go back to the
// same slots and (unless it is the last stage) to their Stockham positions in shared memory.
template <int M, bool STRIDED, int T>
__device__ __forceinline__ void stage_load(float2 (&v)[FftPlan<M>::E], const float2 *s, int j, int l)
{
  using P = FftPlan<M>;
#pragma unroll
  for (int i = 0; i < P::E; i++) v[i] = s[sidx<M, STRIDED, T>(j + i * P::TPL, l)];
}
template <int M, int S, int ST>
__device__ __forceinline__ void stage_math(float2 (&v)[FftPlan<M>::E], const float2 *tw, int j)
{
  using P = FftPlan<M>;
  constexpr int R = P::radix(ST), B = P::E / R, LG = ilog2c(R);
#pragma unroll
  for (int b = 0; b < B; b++) {
    float2 t[R];
#pragma unroll
    for (int r = 0; r < R; r++) t[r] = v[b + B * r];
    if constexpr (ST > 0) tw_apply<R>(t, tw + ((ST == 1 ? 0 : P::NB1) + b * LG) * P::TPL + j, P::TPL);
    dftR<R, S>(t);
#pragma unroll
    for (int r = 0; r < R; r++) v[b + B * r] = t[r];
  }
}
template <int M, bool STRIDED, int T, int ST>
__device__ __forceinline__ void stage_store(const float2 (&v)[FftPlan<M>::E], float2 *s, int j, int l)
{
  using P = FftPlan<M>;
  constexpr int R = P::radix(ST), B = P::E / R, Ns = P::ns(ST);
#pragma unroll
  for (int b = 0; b < B; b++) {
    const int jb = j + b * P::TPL;
    const int base = (jb / Ns) * Ns * R + (jb % Ns);
#pragma unroll
    for (int r = 0; r < R; r++) s[sidx<M, STRIDED, T>(base + r * Ns, l)] = v[b + B * r];
  }
}
template <int M, int S, bool STRIDED, int T, int ST>
__device__ __forceinline__ void fft_stage(float2 (&v)[FftPlan<M>::E], float2 *s, const float2 *tw, int j, int l)
{
  using P = FftPlan<M>;
  if constexpr (ST > 0) {
    __syncthreads();
    stage_load<M, STRIDED, T>(v, s, j, l);
  }
  stage_math<M, S, ST>(v, tw, j);
  if constexpr (ST < P::NST - 1) {
    if constexpr (ST > 0) __syncthreads();
    stage_store<M, STRIDED, T, ST>(v, s, j, l);
  }
}

// Full length-M transform of the lines held by this CTA. On entry v[i] = point j + i*TPL of line l;
// on exit the same slots hold the transform (natural order).
template <int M, int S, bool STRIDED, int T>
__device__ __forceinline__ void fft_lines(float2 (&v)[FftPlan<M>::E], float2 *s, const float2 *tw, int j, int l)
{
  using P = FftPlan<M>;
  fft_stage<M, S, STRIDED, T, 0>(v, s, tw, j, l);
  if constexpr (P::NST >= 2) fft_stage<M, S, STRIDED, T, 1>(v, s, tw, j, l);
  if constexpr (P::NST >= 3) fft_stage<M, S, STRIDED, T, 2>(v, s, tw, j, l);
}

// ------------------------------------------------------------------------------------------
// strided pass: lines of length M, element stride e_stride, T consecutive lines per tile
// Addressing of a strided line: element e lives at outer*outer_stride + (e >> lo_bits)*hi_stride +
// (e & mask)*lo_stride (+ the contiguous index). Single-level strides use lo_bits = 31. The two-level
// form lets the y pass of the distributed transform read straight out of the all-to-all staging
// buffer [source rank][z_local][ky_in_source][kx] (and the r2c write straight into it), so the slab
// transpose costs no extra pack / unpack pass over HBM.
struct LineAddr {
  long long outer_stride, hi_stride, lo_stride;
  int lo_bits;
  // tile-major staging layout of the distributed c2r (see c2r_3d_dist): [source][z-pass tile][z_local][T]
  int tiled = 0, tile_rows = 0;
  // kx-tile layout of the single-GPU c2r (see yx_fused_kernel): the z pass stores point z of line (ky, kx) at
  // [kx / 8][z / G][ky][z % G][kx % 8], so that the G*8 lines a y tile needs are ONE contiguous block of n*G*64 bytes.
  // kx / 8 is the SLOWEST index: the n points of a z line then sit n*64 bytes apart (16 / 32 planes per 2 MB page at
  // n = 2048 / 1024) instead of one full plane apart (one page per point: the translation misses of that layout
  // cost the z pass half its bandwidth at 2048^3), and the blocks of one kx tile are consecutive for consecutive z
  int tile8 = 0, t8_g_log2 = 0, t8_nkt = 0, t8_n = 0;
  __device__ __forceinline__ long long off(int e) const
  { return (long long)(e >> lo_bits) * hi_stride + (long long)(e & ((1 << lo_bits) - 1)) * lo_stride; }
};

// cp.async of one 8-byte point (zero fill when !ok)
__device__ __forceinline__ void cp_async8(float2 *dst_smem, const float2 *src, bool ok)
{
  unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  int sz = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Peer-memory output: element e of a line is stored on rank e >> aout.lo_bits, at peers.p[rank] (that rank's
// staging buffer, already offset to the block reserved for this source rank) + the usual in-rank offset.
struct PeerPtrs { float2 *p[CLR_MAX_PEERS]; };

template <int M, int S, int T>
struct StridedTile {
  using P = FftPlan<M>;
  const float2 *gin; float2 *gout; LineAddr ain, aout; int tiles_per_outer, n_inner, j, l;
  __device__ __forceinline__ void store_peer(long long tile, const float2 (&v)[P::E], const PeerPtrs &peers) const
  {
    long long io, oo;
    if (!locate(tile, io, oo)) return;
    const int mask = (1 << aout.lo_bits) - 1;
    if (aout.tiled) {
      // tile-major: the T lines of this tile at plane z_local sit next to those at z_local+1, so the two (or four)
      // consecutive points a warp stores per instruction form one 256-byte run on the destination GPU, and a
      // whole (tile, destination) block is contiguous -- NVLink writes of 64/128-byte pieces are what limits
      // the natural layout
      const long long tbase = (tile - (tile / tiles_per_outer) * tiles_per_outer) * (long long)(mask + 1) * T + l;
#pragma unroll
      for (int i = 0; i < P::E; i++) {
        const int e = j + i * P::TPL;
        peers.p[e >> aout.lo_bits][tbase + (long long)(e & mask) * T] = v[i];
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < P::E; i++) {
      const int e = j + i * P::TPL;
      peers.p[e >> aout.lo_bits][oo + (long long)(e & mask) * aout.lo_stride] = v[i];
    }
  }
  __device__ __forceinline__ bool locate(long long tile, long long &in_off, long long &out_off) const
  {
    long long outer = tile / tiles_per_outer;
    int inner0 = (int)(tile - outer * tiles_per_outer) * T;
    in_off = outer * ain.outer_stride + inner0 + l;
    out_off = outer * aout.outer_stride + inner0 + l;
    return inner0 + l < n_inner;
  }
  // every thread copies its OWN first-stage inputs into the slots it will read them from
  __device__ __forceinline__ void prefetch(long long tile, float2 *s) const
  {
    long long io, oo;
    bool ok = locate(tile, io, oo);
    const float2 *bin = ok ? gin + io : gin;
    if (ain.tiled) {
      // y pass of the distributed c2r reading the tile-major staging buffer: element e = ky lives in source block
      // e >> lo_bits, at z-pass position inner = ky_local * nc + kx, i.e. tile inner / T, row z_local, slot inner % T
      const long long outer = tile / tiles_per_outer;                   // z_local
      const int kx = (int)(tile - outer * tiles_per_outer) * T + l;
      const int mask = (1 << ain.lo_bits) - 1;
#pragma unroll
      for (int i = 0; i < P::E; i++) {
        const int e = j + i * P::TPL;
        const long long inner = (long long)(e & mask) * ain.lo_stride + kx;
        const long long a = (long long)(e >> ain.lo_bits) * ain.hi_stride + ((inner / T) * ain.tile_rows + outer) * T + (inner % T);
        cp_async8(s + sidx<M, true, T>(e, l), gin + (ok ? a : 0), ok);
      }
      return;
    }
    if (ain.lo_bits == 31) {             // single-level stride: step a pointer instead of 64-bit multiplies
      const float2 *p = bin + (ok ? (long long)j * ain.lo_stride : 0);
      const long long step = ok ? (long long)P::TPL * ain.lo_stride : 0;
#pragma unroll
      for (int i = 0; i < P::E; i++, p += step) cp_async8(s + sidx<M, true, T>(j + i * P::TPL, l), p, ok);
    } else {
#pragma unroll
      for (int i = 0; i < P::E; i++)
        cp_async8(s + sidx<M, true, T>(j + i * P::TPL, l), bin + (ok ? ain.off(j + i * P::TPL) : 0), ok);
    }
  }
  __device__ __forceinline__ void store(long long tile, const float2 (&v)[P::E]) const
  {
    long long io, oo;
    if (!locate(tile, io, oo)) return;
    if (aout.tile8) {
      // z pass of the single-GPU c2r: inner = ky * ncp + kx (ncp a multiple of 8, so an 8-block never straddles a row)
      const int inner = (int)(tile - (tile / tiles_per_outer) * tiles_per_outer) * T + l;
      const int b = inner >> 3, ky = b / aout.t8_nkt, kxt = b - ky * aout.t8_nkt;
      const int gl = aout.t8_g_log2;
      float2 *bo = gout + ((((long long)kxt * (M >> gl)) * aout.t8_n + ky) << (gl + 3)) + (inner & 7);
      const long long zstep = (long long)aout.t8_n << (gl + 3);                       // one group of G planes
#pragma unroll
      for (int i = 0; i < P::E; i++) {
        const int e = j + i * P::TPL;
        bo[(long long)(e >> gl) * zstep + ((e & ((1 << gl) - 1)) << 3)] = v[i];
      }
      return;
    }
    float2 *bout = gout + oo;
    if (aout.lo_bits == 31) {
      float2 *p = bout + (long long)j * aout.lo_stride;
      const long long step = (long long)P::TPL * aout.lo_stride;
#pragma unroll
      for (int i = 0; i < P::E; i++, p += step) *p = v[i];
    } else {
#pragma unroll
      for (int i = 0; i < P::E; i++) bout[aout.off(j + i * P::TPL)] = v[i];
    }
  }
};

// Persistent CTAs, one tile of T lines at a time. The loads of the NEXT tile are issued (cp.async, each
// thread into its own shared-memory slots) as soon as the last exchange of the current tile has been
// read back, so they fly under the last-stage butterflies and the global stores.
// (Measured, round 1: a strided pass moves ~35 G contiguous runs/s whatever the run length -- 32 B runs
// 1.16 TB/s, 64 B 2.3 TB/s, 128 B 4.1 TB/s on the z pass, where every point of a line sits in another 2 MB
// page -- and cp.async .L2::128B / .L2::256B prefetch hints change nothing, so the limit is per-request
// (translation), not DRAM row activation: tiles are as wide as shared memory allows.)
// (Tried and dropped: a separate full-tile landing buffer + half-tile exchange buffer, so that a whole
// tile of 8-byte cp.async is always in flight -- 1.6x SLOWER at n=1024; see DESIGN.md.)
template <int M, int S, int T, bool PEER>
__global__ void __launch_bounds__(T * FftPlan<M>::TPL, (T * FftPlan<M>::TPL * FftPlan<M>::E <= 8192) ? 2 : 1)
fft_strided_kernel(const float2 *gin, float2 *gout, LineAddr ain, LineAddr aout, const float2 *__restrict__ W, int wn,
                   long long n_tiles, int tiles_per_outer, int n_inner, const __grid_constant__ PeerPtrs peers)
{
  using P = FftPlan<M>;
  extern __shared__ float2 smem[];
  float2 *s = smem;
  float2 *tw = smem + P::LSTRIDE * T;
  const int tid = threadIdx.x;
  StridedTile<M, S, T> tl{gin, gout, ain, aout, tiles_per_outer, n_inner, tid / T, tid % T};
  const int j = tl.j, l = tl.l;
  load_twiddles<M, S>(tw, W, wn);
  if (blockIdx.x < n_tiles) tl.prefetch(blockIdx.x, s);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    float2 v[P::E];
    cp_async_wait_all();
    stage_load<M, true, T>(v, s, j, l);                    // own slots: no barrier needed
    if constexpr (P::NST >= 2) {
      stage_math<M, S, 0>(v, tw, j);
      __syncthreads();                                     // everybody has picked up its inputs (and the twiddles are in)
      stage_store<M, true, T, 0>(v, s, j, l);
      __syncthreads();
      stage_load<M, true, T>(v, s, j, l);
      if constexpr (P::NST >= 3) {
        stage_math<M, S, 1>(v, tw, j);
        __syncthreads();
        stage_store<M, true, T, 1>(v, s, j, l);
        __syncthreads();
        stage_load<M, true, T>(v, s, j, l);
      }
      __syncthreads();                                     // last exchange read back: the buffer is free
    }
    if (tile + gridDim.x < n_tiles) tl.prefetch(tile + gridDim.x, s);
    stage_math<M, S, P::NST - 1>(v, tw, j);
    if constexpr (PEER) tl.store_peer(tile, v, peers);
    else tl.store(tile, v);
  }
}

// ------------------------------------------------------------------------------------------
// x pass of the c2r: rows of nc = M+1 complex -> 2M reals, in place; M = n/2.
// Output scaled by `norm`; MOM: accumulate sum / sum of squares (double) into mom[0..1].
template <int M, int T, bool MOM>
__global__ void __launch_bounds__(T * FftPlan<M>::TPL, (T * FftPlan<M>::TPL <= 256) ? 2 : 1)
fft_c2r_x_kernel(float2 *__restrict__ g, const float2 *__restrict__ W, int wn, long long n_rows, int pitch_c,
                 float norm, double *__restrict__ mom)
{
  using P = FftPlan<M>;
  extern __shared__ float2 smem[];
  // s: landing area of the raw rows (M+1 complex each, row stride LSTRIDE) AND exchange buffer of the stages
  float2 *s = smem;
  float2 *tw = smem + T * P::LSTRIDE;
  float2 *wx = tw + P::NTW;   // exp(+2*pi*i*k/n), k < M
  const int tid = threadIdx.x, j = tid % P::TPL, l = tid / P::TPL;
  load_twiddles<M, +1>(tw, W, wn);
  for (int k = tid; k < M; k += blockDim.x) wx[k] = W[k * (wn / (2 * M))];
  double acc1 = 0, acc2 = 0;
  const long long n_tiles = (n_rows + T - 1) / T;
  // every element of a row is fetched ONCE, contiguously, with cp.async (the half-complex combination below
  // needs X[k] and X[M-k]); the next tile's rows are in flight while this tile's results are stored
  auto prefetch = [&](long long tile) {
    const long long row = tile * T + l;
    const bool ok = row < n_rows;
    const float2 *X = g + (ok ? row : 0) * pitch_c;
    float2 *dst = s + l * P::LSTRIDE;
#pragma unroll
    for (int i = 0; i < P::E; i++) cp_async8(dst + j + i * P::TPL, X + j + i * P::TPL, ok);
    if (j == 0) cp_async8(dst + M, X + M, ok);
  };
  if (blockIdx.x < n_tiles) prefetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row = tile * T + l;
    const bool ok = row < n_rows;
    float2 *X = g + row * pitch_c;
    float2 v[P::E];
    cp_async_wait_all();
    __syncthreads();                                 // the rows have landed (and, first time, the twiddles)
    const float2 *raw = s + l * P::LSTRIDE;
#pragma unroll
    for (int i = 0; i < P::E; i++) {
      int k = j + i * P::TPL;
      float2 a = raw[k], b = raw[M - k];
      if (k == 0) { a.y = 0.f; b.y = 0.f; }        // x-DC and x-Nyquist are taken as real
      float2 e = make_float2(a.x + b.x, a.y - b.y);
      float2 d = cmul(make_float2(a.x - b.x, a.y + b.y), wx[k]);
      v[i] = make_float2(e.x - d.y, e.y + d.x);
    }
    if constexpr (P::NST > 1) __syncthreads();      // everybody has picked up its inputs: s is the exchange buffer now
    fft_lines<M, +1, false, T>(v, s, tw, j, l);
    __syncthreads();                                 // last exchange read back: s is free for the next rows
    if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
    if (ok) {
      float s1 = 0, s2 = 0;
#pragma unroll
      for (int i = 0; i < P::E; i++) {
        float2 o = v[i];
        o.x *= norm; o.y *= norm;
        if (MOM) {
          s1 += o.x + o.y;
          s2 += o.x * o.x + o.y * o.y;
        }
        X[j + i * P::TPL] = o;
      }
      if (MOM) { acc1 += s1; acc2 += s2; }
    }
  }
  if (MOM) {
    acc1 = clr_warp_sum(acc1);
    acc2 = clr_warp_sum(acc2);
    __shared__ double red[2][32];
    int w = tid >> 5, ln = tid & 31;
    if (ln == 0) { red[0][w] = acc1; red[1][w] = acc2; }
    __syncthreads();
    if (w == 0) {
      int nw = (blockDim.x + 31) >> 5;
      double a = ln < nw ? red[0][ln] : 0, b = ln < nw ? red[1][ln] : 0;
      a = clr_warp_sum(a); b = clr_warp_sum(b);
      if (ln == 0) { atomicAdd(mom, a); atomicAdd(mom + 1, b); }
    }
  }
}

// x pass of the r2c: rows of 2M reals -> M+1 complex, in place.
template <int M, int T>
__global__ void __launch_bounds__(T * FftPlan<M>::TPL, (T * FftPlan<M>::TPL <= 256) ? 2 : 1)
fft_r2c_x_kernel(float2 *__restrict__ g, const float2 *__restrict__ W, int wn, long long n_rows, int pitch_c)
{
  using P = FftPlan<M>;
  extern __shared__ float2 smem[];
  float2 *s = smem;                              // always needed here (post-processing exchange)
  float2 *tw = smem + T * P::LSTRIDE;
  float2 *wx = tw + P::NTW;                      // exp(-2*pi*i*k/n), k < M
  const int tid = threadIdx.x, j = tid % P::TPL, l = tid / P::TPL;
  load_twiddles<M, -1>(tw, W, wn);
  for (int k = tid; k < M; k += blockDim.x) { float2 w = W[k * (wn / (2 * M))]; w.y = -w.y; wx[k] = w; }
  __syncthreads();
  const long long n_tiles = (n_rows + T - 1) / T;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    long long row = tile * T + l;
    bool ok = row < n_rows;
    float2 *X = g + row * pitch_c;
    float2 v[P::E];
#pragma unroll
    for (int i = 0; i < P::E; i++) v[i] = ok ? X[j + i * P::TPL] : make_float2(0.f, 0.f);
    fft_lines<M, -1, false, T>(v, s, tw, j, l);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < P::E; i++) s[sidx<M, false, T>(j + i * P::TPL, l)] = v[i];
    __syncthreads();
    if (ok) {
#pragma unroll
      for (int i = 0; i < P::E; i++) {
        int k = j + i * P::TPL;
        float2 zk = v[i];
        if (k == 0) {
          X[0] = make_float2(zk.x + zk.y, 0.f);
          X[M] = make_float2(zk.x - zk.y, 0.f);
        } else {
          float2 zc = s[sidx<M, false, T>(M - k, l)];
          zc.y = -zc.y;                                        // conj(Z[M-k])
          float2 e = cadd(zk, zc);
          float2 d = cmul(csub(zk, zc), wx[k]);                // (Z[k]-conj Z[M-k]) * exp(-2 pi i k/n)
          // X[k] = 0.5*(e - i*d)
          X[k] = make_float2(0.5f * (e.x + d.y), 0.5f * (e.y - d.x));
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// y pass + x pass of the c2r in ONE persistent kernel (single GPU / local slab).
//
// Why: a plane of the half-spectrum (4.3 MB at n = 1024, 17 MB at 2048) fits the 126 MB L2 many times over, so the
// y pass of plane z can hand its output to the x pass of the same plane THROUGH L2: DRAM sees one read of the z-pass
// output and one write of the real field (8 B/cell) instead of two reads and two writes (measured with copy kernels
// of the same shape, tools/membench.cu: 1.76 ms against 3.50 ms for a 1024^3 field).
//
// Input `src`: z-pass output in the kx-tile layout [kx / 8][z / G][ky][z % G][kx % 8] (LineAddr::tile8), so a y tile
// (G*8 lines x n points) is one contiguous block of n*G*64 bytes, fetched by the TMA engine with bulk copies
// (cp.async.bulk ... mbarrier::complete_tx) straight into the padded Stockham layout. Output `dst`: the real field in
// the usual layout [z][y][pitch].
//
// Work order: one global ticket counter. Per step s the tickets are: the NKT y tiles of plane group s, then the x tiles
// of group s - LAG. An x tile waits (acquire load) until all y tiles of its group have signalled (release). Every
// y ticket is handed out before the x tickets that depend on it and y tiles never wait, so with all CTAs co-resident
// the wait cannot deadlock; LAG covers the y tiles still in flight so that it is hardly ever entered.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy on the TMA engine (UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int N, int G> struct YxCfg {
  static constexpr int NY = N == 2048 ? (N | kWide) : N;         // 32 points per thread in both transforms
  using PY = FftPlan<NY>;
  using PX = FftPlan<N / 2>;
  static constexpr int TY = 8 * G;                               // lines per y tile
  static constexpr int THREADS = TY * PY::TPL;
  static constexpr int XR = THREADS / PX::TPL;                   // rows per x tile
  static_assert(THREADS % PX::TPL == 0 && XR >= 1 && N % XR == 0, "x tile shape");
  static_assert(PY::NST >= 2 && PX::NST >= 2, "fused y+x pass needs n >= 128");
  static constexpr int BUF = (PY::LSTRIDE * TY > PX::LSTRIDE * XR) ? PY::LSTRIDE * TY : PX::LSTRIDE * XR;   // float2
  static constexpr int G_LOG2 = ilog2c(G);
  static constexpr size_t SMEM = 128 + ((size_t)BUF + PY::NTW + PX::NTW + N / 2) * sizeof(float2);
  // registers: <= 128 per thread whenever the tile is small enough for several CTAs per SM
  static constexpr int MIN_CTAS = SMEM > 110 * 1024 ? 1 : (THREADS >= 512 ? 1 : (512 / THREADS > 8 ? 8 : 512 / THREADS));
};

struct YxArgs {
  const float2 *src; float2 *dst;
  const float2 *W; int wn;
  int nz, ncp, nkt, lag;            // planes of this slab, complex pitch, kx tiles per row, lag in plane groups
  float norm;
  double *mom;
  unsigned *ticket, *done;          // done[group] counts finished y tiles
};

template <int N, int G, bool MOM>
__global__ void __launch_bounds__(YxCfg<N, G>::THREADS, YxCfg<N, G>::MIN_CTAS)
yx_fused_kernel(const YxArgs a)
{
  using C = YxCfg<N, G>;
  using PY = typename C::PY;
  using PX = typename C::PX;
  constexpr int M = N / 2, NY = C::NY;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  // s_tk[0..1]: ticket of the current / next tile (double buffered), s_tk[2..3]: its loads have been issued
  volatile unsigned *s_tk = reinterpret_cast<volatile unsigned *>(smem_raw + 16);
  float2 *s = reinterpret_cast<float2 *>(smem_raw + 128);
  float2 *tw_y = s + C::BUF;
  float2 *tw_x = tw_y + PY::NTW;
  float2 *wx = tw_x + PX::NTW;                                    // exp(+2*pi*i*k/n), k < M
  const int tid = threadIdx.x;
  const int jy = tid / C::TY, ly = tid % C::TY;                   // y tile: lane <-> line
  const int jx = tid % PX::TPL, lx = tid / PX::TPL;               // x tile: lane <-> point
  const int np = a.nz >> C::G_LOG2;                               // plane groups
  const int nxt = N / C::XR;                                      // x tiles per plane
  const unsigned per = (unsigned)a.nkt + (unsigned)(G * nxt);
  const unsigned total = (unsigned)(np + a.lag) * per;
  // bytes of one row of the half-spectrum, rounded up to the 16 bytes a bulk copy moves (the pitch has room)
  constexpr uint32_t ROW_BYTES = ((uint32_t)(M + 1) * 8u + 15u) & ~15u;
  constexpr uint32_t YCHUNK_ROWS = 1u << PY::PADSH, YCHUNK_BYTES = YCHUNK_ROWS * C::TY * 8u, YCHUNKS = N / YCHUNK_ROWS;

  // warp 0 only: start the loads of the tile of ticket t. An x tile needs every y tile of its plane group to have
  // signalled; when `blocking` is false and they have not, nothing is issued and false is returned (the caller retries
  // with blocking = true once this CTA has no unsignalled y tile of its own left, else it could wait for itself).
  auto try_issue = [&](unsigned t, bool blocking) -> bool {
    if (t >= total) return true;
    const int st = (int)(t / per), r = (int)(t - (unsigned)st * per);
    if (r < a.nkt) {
      if (st >= np) return true;
      const float2 *blk = a.src + ((long long)r * np + st) * ((long long)N * C::TY);
      if (tid == 0) { fence_proxy_async(); mbar_expect_tx(bar, (uint32_t)N * C::TY * 8u); }
      __syncwarp();
      for (unsigned c = tid; c < YCHUNKS; c += 32)
        bulk_g2s(s + (size_t)(c * YCHUNK_ROWS + c) * C::TY, blk + (size_t)c * YCHUNK_ROWS * C::TY, YCHUNK_BYTES, bar);
      return true;
    }
    const int p = st - a.lag;
    if (p < 0) return true;
    unsigned ready = 0;
    if (tid == 0) {
      for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.done + p) : "memory");
        ready = v >= (unsigned)a.nkt;
        if (ready || !blocking) break;
        __nanosleep(100);
      }
      if (ready) {
        fence_proxy_async();             // the rows were written through the generic proxy, the copy reads them asynchronously
        mbar_expect_tx(bar, ROW_BYTES * C::XR);
      }
    }
    ready = __shfl_sync(0xffffffffu, ready, 0);
    if (!ready) return false;
    const int xb = r - a.nkt, z = (p << C::G_LOG2) + xb / nxt, row0 = (xb % nxt) * C::XR;
    const float2 *rows = a.dst + ((long long)z * N + row0) * a.ncp;
    for (int rr = tid; rr < C::XR; rr += 32)
      bulk_g2s(s + (size_t)rr * PX::LSTRIDE, rows + (long long)rr * a.ncp, ROW_BYTES, bar);
    return true;
  };
  // warp 0 only: take the next ticket, publish it in s_tk[slot], prefetch its tile if that is possible right now
  auto issue_next = [&](int slot) {
    unsigned t = 0;
    if (tid == 0) t = atomicAdd(a.ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    const bool issued = try_issue(t, false);
    if (tid == 0) { s_tk[slot] = t; s_tk[2 + slot] = issued ? 1u : 0u; }
  };

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  load_twiddles<NY, +1>(tw_y, a.W, a.wn);
  load_twiddles<M, +1>(tw_x, a.W, a.wn);
  for (int k = tid; k < M; k += blockDim.x) wx[k] = a.W[k * (a.wn / (2 * M))];
  __syncthreads();
  if (tid < 32) issue_next(0);
  __syncthreads();
  double acc1 = 0, acc2 = 0;
  uint32_t phase = 0;
  int slot = 0;
  for (unsigned t = s_tk[0]; t < total; t = s_tk[slot]) {
    const int st = (int)(t / per), r = (int)(t - (unsigned)st * per);
    const bool is_y = r < a.nkt;
    const int p = is_y ? st : st - a.lag;
    const bool valid = is_y ? st < np : p >= 0;
    const bool issued = s_tk[2 + slot] != 0;
    slot ^= 1;
    if (!valid) {                                            // head / tail of the ticket sequence: nothing to do
      if (tid < 32) issue_next(slot);
      __syncthreads();
      continue;
    }
    if (!issued && tid < 32) try_issue(t, true);             // x tile whose y tiles were still in flight at prefetch time
    mbar_wait(bar, phase);
    phase ^= 1;
    if (is_y) {
      float2 v[PY::E];
      stage_load<NY, true, C::TY>(v, s, jy, ly);
      stage_math<NY, +1, 0>(v, tw_y, jy);
      __syncthreads();
      stage_store<NY, true, C::TY, 0>(v, s, jy, ly);
      __syncthreads();
      stage_load<NY, true, C::TY>(v, s, jy, ly);
      if constexpr (PY::NST >= 3) {
        stage_math<NY, +1, 1>(v, tw_y, jy);
        __syncthreads();
        stage_store<NY, true, C::TY, 1>(v, s, jy, ly);
        __syncthreads();
        stage_load<NY, true, C::TY>(v, s, jy, ly);
      }
      __syncthreads();                                       // last exchange read back: the buffer is free
      if (tid < 32) issue_next(slot);
      stage_math<NY, +1, PY::NST - 1>(v, tw_y, jy);
      // line ly = (z % G, kx % 8) of plane group p, kx tile r; point e = ky
      const int z = (p << C::G_LOG2) + (ly >> 3);
      float2 *o = a.dst + (long long)z * N * a.ncp + (r << 3) + (ly & 7);
#pragma unroll
      for (int i = 0; i < PY::E; i++) o[(long long)(jy + i * PY::TPL) * a.ncp] = v[i];
      __syncthreads();
      // one RELEASE by one thread publishes the stores of the whole CTA (they happen before it through the barrier;
      // release is cumulative), instead of a device-wide fence in every thread
      if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(a.done + p), "r"(1u) : "memory");
    } else {
      const int xb = r - a.nkt, z = (p << C::G_LOG2) + xb / nxt, row0 = (xb % nxt) * C::XR;
      float2 *X = a.dst + ((long long)z * N + row0 + lx) * a.ncp;
      float2 v[PX::E];
      const float2 *raw = s + lx * PX::LSTRIDE;
#pragma unroll
      for (int i = 0; i < PX::E; i++) {
        const int k = jx + i * PX::TPL;
        float2 x0 = raw[k], x1 = raw[M - k];
        if (k == 0) { x0.y = 0.f; x1.y = 0.f; }             // x-DC and x-Nyquist are taken as real (FFTW c2r)
        const float2 e = make_float2(x0.x + x1.x, x0.y - x1.y);
        const float2 d = cmul(make_float2(x0.x - x1.x, x0.y + x1.y), wx[k]);
        v[i] = make_float2(e.x - d.y, e.y + d.x);
      }
      // a row belongs to the PX::TPL <= 32 threads of ONE warp (tid = lx * TPL + jx) and its exchange stays inside the
      // row's own stretch of s: warp barriers are enough until the buffer is handed back to the bulk loads
      static_assert(PX::TPL <= 32 && 32 % PX::TPL == 0, "x rows must not straddle warps");
      __syncwarp();                                          // the row's inputs are in registers: s is the exchange buffer now
      stage_math<M, +1, 0>(v, tw_x, jx);
      stage_store<M, false, C::XR, 0>(v, s, jx, lx);
      __syncwarp();
      stage_load<M, false, C::XR>(v, s, jx, lx);
      if constexpr (PX::NST >= 3) {
        stage_math<M, +1, 1>(v, tw_x, jx);
        __syncwarp();
        stage_store<M, false, C::XR, 1>(v, s, jx, lx);
        __syncwarp();
        stage_load<M, false, C::XR>(v, s, jx, lx);
      }
      __syncthreads();
      if (tid < 32) issue_next(slot);
      stage_math<M, +1, PX::NST - 1>(v, tw_x, jx);
      float s1 = 0, s2 = 0;
#pragma unroll
      for (int i = 0; i < PX::E; i++) {
        float2 o = v[i];
        o.x *= a.norm; o.y *= a.norm;
        if (MOM) { s1 += o.x + o.y; s2 += o.x * o.x + o.y * o.y; }
        __stcs(X + jx + i * PX::TPL, o);                     // final result: streams out, keep L2 for the planes in flight
      }
      if (MOM) { acc1 += s1; acc2 += s2; }
      __syncthreads();                                       // s_tk[slot] is visible
    }
  }
  if (MOM) {
    acc1 = clr_warp_sum(acc1);
    acc2 = clr_warp_sum(acc2);
    __shared__ double red[2][32];
    const int w = tid >> 5, ln = tid & 31;
    if (ln == 0) { red[0][w] = acc1; red[1][w] = acc2; }
    __syncthreads();
    if (w == 0) {
      const int nw = (blockDim.x + 31) >> 5;
      double x = ln < nw ? red[0][ln] : 0, y = ln < nw ? red[1][ln] : 0;
      x = clr_warp_sum(x); y = clr_warp_sum(y);
      if (ln == 0) { atomicAdd(a.mom, x); atomicAdd(a.mom + 1, y); }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Gaussian mode fill FUSED into the z pass of the c2r, for delta_k and phi_k at once (create_grids_fourier,
// fourier.c:285-359, + the first axis of fftw_wrap_c2r). The modes are never written to memory: a CTA generates the
// 8 kx x n kz modes of its tile for both fields into two shared-memory buffers (one Philox block per mode pair, the
// arithmetic of clr_fill.cuh), two thread groups transform one buffer each along z, and only the z-pass OUTPUT goes
// to HBM, in the kx-tile layout the fused y+x pass reads. Saves the 8 B/cell write of the stand-alone fill and the
// 8 B/cell read of two z passes; the kernel is bound by instruction issue (Philox + transcendentals + butterflies).
template <int N, int WW> struct FzCfg {
  static constexpr int NP = N == 2048 ? (N | kWide) : N;        // 32 points per thread at every size
  using P = FftPlan<NP>;
  static constexpr int W = WW;                                  // lines (kx) per tile and field: 8, or 4 (half a kx tile)
  static constexpr int H = 8 / W;                               // tiles per 64-byte run of the output layout
  static constexpr int GT = W * P::TPL;                         // threads of one field group
  static constexpr int THREADS = 2 * GT;
  static constexpr int BUF = P::LSTRIDE * W;                    // float2 per field
  static constexpr size_t SMEM = ((size_t)2 * BUF + P::NTW) * sizeof(float2);
  static constexpr int MIN_CTAS = (THREADS >= 512 || SMEM > 110 * 1024) ? 1 : (512 / THREADS > 4 ? 4 : 512 / THREADS);
  static_assert(P::NST >= 2, "fused fill + z pass needs n >= 64");
  static_assert(W == 4 || W == 8, "tile width");
};

struct FzArgs {
  float2 *out_d, *out_p;            // z-pass output of delta / phi, kx-tile layout [kx/8][z][ky_local][kx%8]
  const float2 *W; int wn;
  const float2 *pkt, *sct;
  uint32_t seed;
  int n, nc, nyl, ky0, nkt;         // grid side, n/2+1, ky slab of this rank, kx tiles per row
  FillFastK k;
};

template <int N, int WW>
__global__ void __launch_bounds__(FzCfg<N, WW>::THREADS, FzCfg<N, WW>::MIN_CTAS)
fill_z_kernel(const __grid_constant__ FzArgs a)
{
  using C = FzCfg<N, WW>;
  using P = typename C::P;
  constexpr int W = C::W, H = C::H, NP = C::NP;
  extern __shared__ float2 smem[];
  float2 *tw = smem + 2 * C::BUF;
  const int tid = threadIdx.x;
  const int grp = tid / C::GT, tg = tid - grp * C::GT;           // field of this thread's transform
  const int j = tg / W, l = tg % W;
  float2 *sg = smem + grp * C::BUF;
  float2 *outg = grp ? a.out_p : a.out_d;
  load_twiddles<NP, +1>(tw, a.W, a.wn);
  const int npair_row = (a.nc + 1) / 2;
  const long long n_tiles = (long long)a.nkt * a.nyl * H;
  const long long zstep = (long long)a.nyl * 8;                  // one z plane inside a kx tile of the kx-tile layout
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // consecutive tiles = the H pieces of one 64-byte output run, then consecutive ky of the same kx tile: pieces and
    // neighbouring runs are written at about the same time by neighbouring CTAs and meet in L2
    const long long run = tile / H;
    const int half = (int)(tile - run * H);
    const int kxt = (int)(run / a.nyl), kyl = (int)(run - (long long)kxt * a.nyl);
    const int jj = a.ky0 + kyl;
    const int mj = (2 * jj <= N ? jj : N - jj);
    // ---- fill: W/2 mode pairs per kz
#pragma unroll 2
    for (int q = tid; q < N * (W / 2); q += C::THREADS) {
      const int kz = q / (W / 2), pr = q - kz * (W / 2);
      const int mi = (2 * kz <= N ? kz : N - kz);
      const int m_row = mj * mj + mi * mi;
      const int kk0 = kxt * 8 + half * W + 2 * pr;
      const unsigned long long gidx = (unsigned long long)(kk0 >> 1) + (unsigned long long)npair_row * ((unsigned long long)jj + (unsigned long long)N * kz);
      float2 dk2[2], pk2[2];
      uint32_t w[4];
      clr_philox((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, 0u, a.seed, 0u, w);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int kk = kk0 + h, m = kk * kk + m_row;            // beyond the Nyquist column: padding lines, zero
        clr_fill_mode(a.k, a.pkt, a.sct, m, kk < a.nc && m > 0, w[2 * h], w[2 * h + 1], dk2[h], pk2[h]);
      }
      const int si = sidx<NP, true, W>(kz, 2 * pr);
      *reinterpret_cast<float4 *>(smem + si) = make_float4(dk2[0].x, dk2[0].y, dk2[1].x, dk2[1].y);
      *reinterpret_cast<float4 *>(smem + C::BUF + si) = make_float4(pk2[0].x, pk2[0].y, pk2[1].x, pk2[1].y);
    }
    __syncthreads();
    // ---- transform along z: group 0 = delta_k, group 1 = phi_k
    float2 v[P::E];
    stage_load<NP, true, W>(v, sg, j, l);
    stage_math<NP, +1, 0>(v, tw, j);
    __syncthreads();
    stage_store<NP, true, W, 0>(v, sg, j, l);
    __syncthreads();
    stage_load<NP, true, W>(v, sg, j, l);
    if constexpr (P::NST >= 3) {
      stage_math<NP, +1, 1>(v, tw, j);
      __syncthreads();
      stage_store<NP, true, W, 1>(v, sg, j, l);
      __syncthreads();
      stage_load<NP, true, W>(v, sg, j, l);
    }
    __syncthreads();                                             // buffers free: the next fill may overwrite them
    stage_math<NP, +1, P::NST - 1>(v, tw, j);
    float2 *o = outg + ((long long)kxt * N * a.nyl + kyl) * 8 + half * W + l;
#pragma unroll
    for (int i = 0; i < P::E; i++) o[(long long)(j + i * P::TPL) * zstep] = v[i];
  }
}

// ------------------------------------------------------------------------------------------
// The same fusion on a CLUSTER OF TWO CTAs, for grids whose tile (two fields x 8 lines x n points) does not fit the
// shared memory of one SM (n = 2048: 2 x 135 KB). CTA 0 of the pair transforms delta_k, CTA 1 phi_k, each with its own
// 8-line tile in its own shared memory. The fill is shared work: every CTA generates the modes of HALF of the kz range
// for BOTH fields (one Philox block and one P(k) / phase evaluation per mode pair, as in fill_z_kernel), keeps its own
// field's values and pushes the other field's values into the partner's shared memory through distributed shared
// memory (st.shared::cluster via cluster.map_shared_rank). Cluster barriers: "tile complete" (release / acquire over
// both CTAs' stores) before the transforms, and a split "buffer free" barrier -- arrive after the last read of the
// exchange buffer, wait at the top of the next fill -- so the partner never overwrites a tile that is still being read.
// Output runs are 64 bytes (8 lines) instead of the 32 bytes of the 4-line tiles of fill_z_kernel<2048, 4>.
template <int N> struct FzcCfg {
  static constexpr int NP = N == 2048 ? (N | kWide) : N;
  using P = FftPlan<NP>;
  static constexpr int W = 8;
  static constexpr int THREADS = W * P::TPL;                    // one field per CTA
  static constexpr int BUF = P::LSTRIDE * W;
  static constexpr size_t SMEM = ((size_t)BUF + P::NTW) * sizeof(float2);
  static_assert(P::NST >= 2 && N % 2 == 0, "cluster fill + z pass needs n >= 64");
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FzcCfg<N>::THREADS, 1)
fill_z_cluster_kernel(const __grid_constant__ FzArgs a)
{
  using C = FzcCfg<N>;
  using P = typename C::P;
  constexpr int W = C::W, NP = C::NP;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float2 smem[];
  float2 *tw = smem + C::BUF;
  const int tid = threadIdx.x;
  const unsigned field = cluster.block_rank();                   // 0: this CTA transforms delta_k, 1: phi_k
  float2 *mine = smem;
  float2 *theirs = cluster.map_shared_rank(smem, field ^ 1u);    // the partner's tile (distributed shared memory)
  float2 *buf_d = field == 0 ? mine : theirs, *buf_p = field == 0 ? theirs : mine;
  const int j = tid / W, l = tid % W;
  float2 *outg = field ? a.out_p : a.out_d;
  load_twiddles<NP, +1>(tw, a.W, a.wn);
  const int npair_row = (a.nc + 1) / 2;
  const long long n_tiles = (long long)a.nkt * a.nyl;
  const long long zstep = (long long)a.nyl * 8;
  const int n_cl = gridDim.x >> 1;
  const int kz_lo = (int)field * (N / 2);                        // this CTA fills kz in [kz_lo, kz_lo + N/2)
  cluster_arrive();                                              // opens the first "buffer free" phase
  for (long long tile = blockIdx.x >> 1; tile < n_tiles; tile += n_cl) {
    const int kxt = (int)(tile / a.nyl), kyl = (int)(tile - (long long)kxt * a.nyl);
    const int jj = a.ky0 + kyl;
    const int mj = (2 * jj <= N ? jj : N - jj);
    cluster_wait();                                              // both CTAs have read their previous tile back
#pragma unroll 2
    for (int q = tid; q < (N / 2) * (W / 2); q += C::THREADS) {
      const int kz = kz_lo + q / (W / 2), pr = q % (W / 2);
      const int mi = (2 * kz <= N ? kz : N - kz);
      const int m_row = mj * mj + mi * mi;
      const int kk0 = kxt * 8 + 2 * pr;
      const unsigned long long gidx = (unsigned long long)(kk0 >> 1) + (unsigned long long)npair_row * ((unsigned long long)jj + (unsigned long long)N * kz);
      float2 dk2[2], pk2[2];
      uint32_t w[4];
      clr_philox((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, 0u, a.seed, 0u, w);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int kk = kk0 + h, m = kk * kk + m_row;            // beyond the Nyquist column: padding lines, zero
        clr_fill_mode(a.k, a.pkt, a.sct, m, kk < a.nc && m > 0, w[2 * h], w[2 * h + 1], dk2[h], pk2[h]);
      }
      const int si = sidx<NP, true, W>(kz, 2 * pr);
      *reinterpret_cast<float4 *>(buf_d + si) = make_float4(dk2[0].x, dk2[0].y, dk2[1].x, dk2[1].y);
      *reinterpret_cast<float4 *>(buf_p + si) = make_float4(pk2[0].x, pk2[0].y, pk2[1].x, pk2[1].y);
    }
    cluster_arrive();                                            // tile complete: my stores (local and remote) are released ...
    cluster_wait();                                              // ... and the partner's half has arrived
    float2 v[P::E];
    stage_load<NP, true, W>(v, mine, j, l);
    stage_math<NP, +1, 0>(v, tw, j);
    __syncthreads();
    stage_store<NP, true, W, 0>(v, mine, j, l);
    __syncthreads();
    stage_load<NP, true, W>(v, mine, j, l);
    if constexpr (P::NST >= 3) {
      stage_math<NP, +1, 1>(v, tw, j);
      __syncthreads();
      stage_store<NP, true, W, 1>(v, mine, j, l);
      __syncthreads();
      stage_load<NP, true, W>(v, mine, j, l);
    }
    cluster_arrive();                                            // buffer free (waited for at the top of the next fill)
    stage_math<NP, +1, P::NST - 1>(v, tw, j);
    float2 *o = outg + ((long long)kxt * N * a.nyl + kyl) * 8 + l;
#pragma unroll
    for (int i = 0; i < P::E; i++) o[(long long)(j + i * P::TPL) * zstep] = v[i];
  }
  cluster_wait();                                                // pair up the last arrive: no CTA leaves while its partner may still push
}

// ------------------------------------------------------------------------------------------
// Several GPUs: mode fill fused into the z pass WITH the slab transpose (create_grids_fourier, fourier.c:285-359, + the
// first axis of fftw_wrap_c2r, fourier.c:81-102, + FFTW-MPI's transpose). One field per launch: a CTA generates the
// T lines x n kz modes of its tile (the same Philox blocks and arithmetic as the stand-alone fill, clr_fill.cuh),
// transforms them along z and stores plane z straight into the staging buffer of rank z / nz_local over NVLink
// (StridedTile::store_peer, natural or tile-major staging). The pass is bound by NVLink, so the fill arithmetic hides
// under the transfer: the stand-alone fill (8 B/cell written) and the 8 B/cell read of the z pass disappear.
// Tiles = T consecutive lines of the flattened (ky_local, kx) index with the row pitch ncp, like fft_strided_kernel.
struct FpArgs {
  LineAddr aout; int tiles_per_outer, n_inner;
  const float2 *W; int wn;
  const float2 *pkt, *sct;
  uint32_t seed;
  int nc, ncp, ky0;                 // n/2+1, complex row pitch, first ky of this rank's slab
  FillFastK k;
};

template <int N, int T, int FIELD>
__global__ void __launch_bounds__(T * FftPlan<N>::TPL, (T * FftPlan<N>::TPL * FftPlan<N>::E <= 8192) ? 2 : 1)
fill_peer_kernel(const __grid_constant__ FpArgs a, long long n_tiles, const __grid_constant__ PeerPtrs peers)
{
  using P = FftPlan<N>;
  constexpr int THREADS = T * P::TPL, HP = T / 2;                // mode pairs per kz of a tile
  static_assert(THREADS % HP == 0 && T % 2 == 0, "a thread keeps its pair of lines for the whole tile");
  extern __shared__ float2 smem[];
  float2 *s = smem;
  float2 *tw = smem + P::LSTRIDE * T;
  const int tid = threadIdx.x;
  StridedTile<N, +1, T> tl{nullptr, nullptr, a.aout, a.aout, a.tiles_per_outer, a.n_inner, tid / T, tid % T};
  const int j = tl.j, l = tl.l;
  load_twiddles<N, +1>(tw, a.W, a.wn);
  const int npair_row = (a.nc + 1) / 2;
  const int pr = tid % HP, kz0 = tid / HP;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- fill: this thread's pair of lines (flattened index inner, inner + 1: same ky, kx even)
    const long long inner = tile * T + 2 * pr;
    const int kyl = (int)(inner / a.ncp), kk0 = (int)(inner - (long long)kyl * a.ncp);
    const int jj = a.ky0 + kyl;
    const int mj = (2 * jj <= N ? jj : N - jj);
    const bool row_ok = inner < a.n_inner;
#pragma unroll 2
    for (int kz = kz0; kz < N; kz += THREADS / HP) {
      const int mi = (2 * kz <= N ? kz : N - kz);
      const int m_row = mj * mj + mi * mi;
      const unsigned long long gidx = (unsigned long long)(kk0 >> 1) + (unsigned long long)npair_row * ((unsigned long long)jj + (unsigned long long)N * kz);
      float2 dk2[2], pk2[2];
      uint32_t w[4];
      clr_philox((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, 0u, a.seed, 0u, w);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int kk = kk0 + h, m = kk * kk + m_row;            // beyond the Nyquist column: padding, zero
        clr_fill_mode(a.k, a.pkt, a.sct, m, row_ok && kk < a.nc && m > 0, w[2 * h], w[2 * h + 1], dk2[h], pk2[h]);
      }
      const float2 v0 = FIELD ? pk2[0] : dk2[0], v1 = FIELD ? pk2[1] : dk2[1];
      *reinterpret_cast<float4 *>(s + sidx<N, true, T>(kz, 2 * pr)) = make_float4(v0.x, v0.y, v1.x, v1.y);
    }
    __syncthreads();
    // ---- transform along z, peer stores
    float2 v[P::E];
    stage_load<N, true, T>(v, s, j, l);
    stage_math<N, +1, 0>(v, tw, j);
    __syncthreads();
    stage_store<N, true, T, 0>(v, s, j, l);
    __syncthreads();
    stage_load<N, true, T>(v, s, j, l);
    if constexpr (P::NST >= 3) {
      stage_math<N, +1, 1>(v, tw, j);
      __syncthreads();
      stage_store<N, true, T, 1>(v, s, j, l);
      __syncthreads();
      stage_load<N, true, T>(v, s, j, l);
    }
    __syncthreads();                                             // buffer free: the next fill may overwrite it
    stage_math<N, +1, P::NST - 1>(v, tw, j);
    tl.store_peer(tile, v, peers);
  }
}

// ------------------------------------------------------------------------------------------
// host-side dispatch
template <int M> struct Cfg {   // lines per CTA tile, per transform length
  static constexpr int TPL = FftPlan<M>::TPL;
  static constexpr int T_DEF = M >= 4096 ? 4 : (M == 1024 ? 16 : (256 / TPL > 64 ? 64 : (256 / TPL < 8 ? 8 : 256 / TPL)));
  static constexpr int T_STRIDED = (M == 2048 && (CLR_FFT_VARIANT == 2 || CLR_FFT_VARIANT == 3)) ? 4
                                   : ((M == 1024 && CLR_FFT_VARIANT >= 1) ? 8 : T_DEF);
  // z pass: every point of a line sits in another 2 MB page (plane stride), so wider tiles halve the
  // TLB misses per byte
  static constexpr int T_Z = ((M == 1024 && CLR_FFT_VARIANT == 0) || M == 512) ? 16 : T_STRIDED;
  static constexpr int T_X = M >= 2048 ? 4 : (256 / TPL > 64 ? 64 : 256 / TPL);
};

// grid of a persistent kernel = resident CTAs per SM x SMs (capped by the tile count). The attribute call and the
// occupancy query cost ~10 us each: done once per (kernel, block size, shared memory), not per launch.
template <typename K> int launch_cfg(clr_ctx *c, K kernel, int threads, size_t smem, long long n_tiles, int *grid)
{
  static std::map<std::pair<const void *, std::pair<int, size_t>>, int> cache;
  const auto key = std::make_pair(reinterpret_cast<const void *>(kernel), std::make_pair(threads, smem));
  auto it = cache.find(key);
  int per_sm = 0;
  if (it != cache.end()) per_sm = it->second;
  else {
    CLR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CLR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    CLR_CHECK(per_sm > 0, "FFT kernel does not fit on an SM (threads=%d smem=%zu)", threads, smem);
    cache[key] = per_sm;
  }
  long long g = (long long)per_sm * c->sm_count;
  if (g > n_tiles) g = n_tiles;
  *grid = (int)g;
  return 0;
}

template <int M, int S, int T = Cfg<M>::T_STRIDED>
int run_strided2(clr_ctx *c, const float2 *gin, float2 *gout, LineAddr ain, LineAddr aout, long long n_outer, int n_inner,
                 const PeerPtrs *peers = nullptr)
{
  using P = FftPlan<M>;
  constexpr int threads = T * P::TPL;
  size_t smem = ((size_t)P::LSTRIDE * T + P::NTW) * sizeof(float2);
  int tiles_per_outer = (n_inner + T - 1) / T;
  long long n_tiles = n_outer * tiles_per_outer;
  int grid;
  if (peers) {
    auto k = fft_strided_kernel<M, S, T, true>;
    if (launch_cfg(c, k, threads, smem, n_tiles, &grid)) return 1;
    k<<<grid, threads, smem, c->stream>>>(gin, gout, ain, aout, c->d_twiddle, c->dev.n, n_tiles, tiles_per_outer, n_inner, *peers);
  } else {
    auto k = fft_strided_kernel<M, S, T, false>;
    if (launch_cfg(c, k, threads, smem, n_tiles, &grid)) return 1;
    k<<<grid, threads, smem, c->stream>>>(gin, gout, ain, aout, c->d_twiddle, c->dev.n, n_tiles, tiles_per_outer, n_inner, PeerPtrs{});
  }
  CLR_CUDA(cudaGetLastError());
  return 0;
}

// staging-buffer pointers of every rank, offset to the block this rank writes (block index = source rank)
PeerPtrs peer_blocks(clr_ctx *c, size_t block_float2)
{
  PeerPtrs pp{};
  for (int h = 0; h < c->nranks; h++) pp.p[h] = reinterpret_cast<float2 *>(c->peer_stage[h]) + (size_t)c->rank * block_float2;
  return pp;
}

template <int M, int S, int T = Cfg<M>::T_STRIDED>
int run_strided(clr_ctx *c, float2 *g, long long n_outer, long long outer_stride, long long e_stride, int n_inner)
{
  LineAddr a{outer_stride, 0, e_stride, 31};
  return run_strided2<M, S, T>(c, g, g, a, a, n_outer, n_inner);
}

// ---- slab-decomposed transform (one process per GPU) ----------------------------------------------
// k space is held in y slabs, layout [kz][ky_local][kx] (the mode fill is layout free); real space in
// z slabs [z_local][y][x] like the reference (fourier.c:172-177). c2r: z pass local -> ONE all-to-all
// (block for rank h = the z planes of h: contiguous, no pack) -> y pass reading the staging buffer
// through the two-level LineAddr and writing the natural layout -> x pass. r2c runs the mirror image.
int ilog2_host(int v) { int b = 0; while ((1 << b) < v) b++; return b; }
template <int M, bool MOM> int run_c2r_x(clr_ctx *c, float2 *g, long long n_rows, int pitch_c, float norm, double *mom);
template <int M> int run_r2c_x(clr_ctx *c, float2 *g, long long n_rows, int pitch_c);

// z pass of the distributed c2r with the mode fill fused in (fill_peer_kernel): field 0 = delta_k, 1 = phi_k
struct FillSpec { uint32_t seed; int field; };

template <int N>
int run_fill_peer(clr_ctx *c, const FillSpec &fs, LineAddr aout, int n_inner, const PeerPtrs &pp)
{
  constexpr int T = Cfg<N>::T_STRIDED;
  using P = FftPlan<N>;
  constexpr int threads = T * P::TPL;
  const size_t smem = ((size_t)P::LSTRIDE * T + P::NTW) * sizeof(float2);
  FpArgs a;
  if (clr_fill_fast_setup(c, &a.k)) return 1;
  a.aout = aout; a.n_inner = n_inner; a.tiles_per_outer = (n_inner + T - 1) / T;
  a.W = c->d_twiddle; a.wn = c->dev.n; a.pkt = c->d_pkt; a.sct = c->d_sincos; a.seed = fs.seed;
  a.nc = c->dev.nc; a.ncp = c->dev.ncp; a.ky0 = c->dev.ky0;
  const long long n_tiles = a.tiles_per_outer;
  int grid;
  if (fs.field) {
    auto k = fill_peer_kernel<N, T, 1>;
    if (launch_cfg(c, k, threads, smem, n_tiles, &grid)) return 1;
    k<<<grid, threads, smem, c->stream>>>(a, n_tiles, pp);
  } else {
    auto k = fill_peer_kernel<N, T, 0>;
    if (launch_cfg(c, k, threads, smem, n_tiles, &grid)) return 1;
    k<<<grid, threads, smem, c->stream>>>(a, n_tiles, pp);
  }
  CLR_CUDA(cudaGetLastError());
  return 0;
}

template <int N>
int c2r_3d_dist(clr_ctx *c, float2 *g, float norm, double *mom, const FillSpec *fs = nullptr)
{
  const long long nc = c->dev.ncp;     // complex PITCH of a row (>= N/2+1, multiple of 8)
  const int P = c->nranks, nzl = N / P, nyl = N / P;
  float2 *stage = reinterpret_cast<float2 *>(c->d_stage);
  CLR_CHECK(!fs || (c->p2p && c->p2p_enabled), "fused fill + transpose needs the peer-memory transpose");
  // tile-major staging pays when the remote runs of the natural layout are short (T = 8 lines = 64 bytes at
  // n_grid = 2048: 346 -> 496 GB/s per direction on 8 GPUs, step 46.2 -> 42.5 ms) and most of the output leaves the
  // GPU; with 128-byte runs (n_grid = 1024) or 2 GPUs the gather it forces on the y pass costs more than it saves
  const bool tiled = c->p2p_tiled < 0 ? (c->nranks >= 4 && Cfg<N>::T_STRIDED <= 8) : c->p2p_tiled != 0;
  if (c->p2p && c->p2p_enabled && tiled) {
    // z pass with the slab transpose fused into its stores: plane z of the result belongs to rank z / nzl and is
    // written straight into that rank's staging buffer over NVLink (block = source rank), tile by tile while
    // the next tile is being transformed. Barriers: nobody still reads its staging buffer / everything arrived.
    constexpr int TT = Cfg<N>::T_STRIDED;
    const long long blk_t = (long long)nzl * TT * ((nyl * nc + TT - 1) / TT);     // tile-major block of one source
    { StageScope sc(c, "fft_z", 1);
      if (clr_comm_barrier(c)) return 1;
      PeerPtrs pp = peer_blocks(c, (size_t)blk_t);
      LineAddr ain{0, 0, (long long)nyl * nc, 31};
      LineAddr aout{0, 0, (long long)nyl * nc, ilog2_host(nzl)};
      aout.tiled = 1;
      if (fs ? run_fill_peer<N>(c, *fs, aout, (int)(nyl * nc), pp)
             : run_strided2<N, +1>(c, g, nullptr, ain, aout, 1, (int)(nyl * nc), &pp)) return 1;
      if (clr_comm_barrier(c)) return 1;
      if (c->ev_after_z) CLR_CUDA(cudaEventRecord(c->ev_after_z, c->stream));
      c->a2a_bytes += (double)nzl * nyl * nc * 8 * (c->nranks - 1); }
    { StageScope sc2(c, "fft_y", 1);
      LineAddr yin{0, blk_t, nc, ilog2_host(nyl)};
      yin.tiled = 1; yin.tile_rows = nzl;
      LineAddr yout{(long long)N * nc, 0, nc, 31};
      if (run_strided2<N, +1>(c, stage, g, yin, yout, nzl, N / 2 + 1)) return 1; }
    StageScope sc3(c, "fft_x", 1);
    if (mom) return run_c2r_x<N / 2, true>(c, g, (long long)nzl * N, (int)nc, norm, mom);
    return run_c2r_x<N / 2, false>(c, g, (long long)nzl * N, (int)nc, norm, nullptr);
  } else if (c->p2p && c->p2p_enabled) {
    // same fusion, natural staging layout [source][z_local][ky_in_source][kx]
    StageScope sc(c, "fft_z", 1);
    if (clr_comm_barrier(c)) return 1;
    PeerPtrs pp = peer_blocks(c, (size_t)nzl * nyl * nc);
    LineAddr ain{0, 0, (long long)nyl * nc, 31};
    LineAddr aout{0, 0, (long long)nyl * nc, ilog2_host(nzl)};
    if (fs ? run_fill_peer<N>(c, *fs, aout, (int)(nyl * nc), pp)
           : run_strided2<N, +1>(c, g, nullptr, ain, aout, 1, (int)(nyl * nc), &pp)) return 1;
    if (clr_comm_barrier(c)) return 1;
    if (c->ev_after_z) CLR_CUDA(cudaEventRecord(c->ev_after_z, c->stream));
    c->a2a_bytes += (double)nzl * nyl * nc * 8 * (c->nranks - 1);
  } else {
    { StageScope sc(c, "fft_z", 1);
      if (run_strided<N, +1>(c, g, 1, 0, (long long)nyl * nc, (int)(nyl * nc))) return 1; }
    { StageScope sc(c, "fft_a2a", 0);
      if (clr_comm_alltoall(c, g, stage, (size_t)nzl * nyl * nc * 2)) return 1; }
    if (c->ev_after_z) CLR_CUDA(cudaEventRecord(c->ev_after_z, c->stream));
  }
  { StageScope sc(c, "fft_y", 1);
    LineAddr ain{(long long)nyl * nc, (long long)nzl * nyl * nc, nc, ilog2_host(nyl)};
    LineAddr aout{(long long)N * nc, 0, nc, 31};
    if (run_strided2<N, +1>(c, stage, g, ain, aout, nzl, N / 2 + 1)) return 1; }
  StageScope sc(c, "fft_x", 1);
  if (mom) return run_c2r_x<N / 2, true>(c, g, (long long)nzl * N, (int)nc, norm, mom);
  return run_c2r_x<N / 2, false>(c, g, (long long)nzl * N, (int)nc, norm, nullptr);
}

template <int N>
int r2c_3d_dist(clr_ctx *c, float2 *g)
{
  const long long nc = c->dev.ncp;     // complex PITCH of a row (>= N/2+1, multiple of 8)
  const int P = c->nranks, nzl = N / P, nyl = N / P;
  float2 *stage = reinterpret_cast<float2 *>(c->d_stage);
  { StageScope sc(c, "fft_x", 1); if (run_r2c_x<N / 2>(c, g, (long long)nzl * N, (int)nc)) return 1; }
  if (c->p2p && c->p2p_enabled) {
    // y pass storing ky block s straight into rank s's staging buffer; the z pass then runs staging -> grid
    { StageScope sc(c, "fft_y", 1);
      if (clr_comm_barrier(c)) return 1;
      PeerPtrs pp = peer_blocks(c, (size_t)nzl * nyl * nc);
      LineAddr ain{(long long)N * nc, 0, nc, 31};
      LineAddr aout{(long long)nyl * nc, 0, nc, ilog2_host(nyl)};
      if (run_strided2<N, -1>(c, g, nullptr, ain, aout, nzl, N / 2 + 1, &pp)) return 1;
      if (clr_comm_barrier(c)) return 1;
      c->a2a_bytes += (double)nzl * nyl * nc * 8 * (c->nranks - 1); }
    StageScope sc(c, "fft_z", 1);
    LineAddr a{0, 0, (long long)nyl * nc, 31};
    return run_strided2<N, -1>(c, stage, g, a, a, 1, (int)(nyl * nc));
  }
  { StageScope sc(c, "fft_y", 1);
    LineAddr ain{(long long)N * nc, 0, nc, 31};
    LineAddr aout{(long long)nyl * nc, (long long)nzl * nyl * nc, nc, ilog2_host(nyl)};
    if (run_strided2<N, -1>(c, g, stage, ain, aout, nzl, N / 2 + 1)) return 1; }
  { StageScope sc(c, "fft_a2a", 0);
    if (clr_comm_alltoall(c, stage, g, (size_t)nzl * nyl * nc * 2)) return 1; }
  StageScope sc(c, "fft_z", 1);
  return run_strided<N, -1>(c, g, 1, 0, (long long)nyl * nc, (int)(nyl * nc));
}

template <int M, bool MOM>
int run_c2r_x(clr_ctx *c, float2 *g, long long n_rows, int pitch_c, float norm, double *mom)
{
  using P = FftPlan<M>;
  constexpr int T = Cfg<M>::T_X;
  constexpr int threads = T * P::TPL;
  size_t smem = ((size_t)T * P::LSTRIDE + P::NTW + M) * sizeof(float2);
  int grid;
  auto k = fft_c2r_x_kernel<M, T, MOM>;
  if (launch_cfg(c, k, threads, smem, (n_rows + T - 1) / T, &grid)) return 1;
  k<<<grid, threads, smem, c->stream>>>(g, c->d_twiddle, c->dev.n, n_rows, pitch_c, norm, mom);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

template <int M>
int run_r2c_x(clr_ctx *c, float2 *g, long long n_rows, int pitch_c)
{
  using P = FftPlan<M>;
  constexpr int T = Cfg<M>::T_X;
  constexpr int threads = T * P::TPL;
  size_t smem = ((size_t)T * P::LSTRIDE + P::NTW + M) * sizeof(float2);
  int grid;
  auto k = fft_r2c_x_kernel<M, T>;
  if (launch_cfg(c, k, threads, smem, (n_rows + T - 1) / T, &grid)) return 1;
  k<<<grid, threads, smem, c->stream>>>(g, c->d_twiddle, c->dev.n, n_rows, pitch_c);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

// scratch of the single-GPU c2r: the z-pass output in the kx-tile layout (one slab) + ticket / done counters
int ensure_fft_tmp(clr_ctx *c)
{
  const size_t bytes = (size_t)c->dev.nz_here * c->dev.n * c->dev.ncp * sizeof(float2);
  if (c->d_fft_tmp && c->fft_tmp_bytes >= bytes) return 0;
  if (c->d_fft_tmp) cudaFree(c->d_fft_tmp);
  c->d_fft_tmp = nullptr; c->fft_tmp_bytes = 0;
  CLR_CUDA(cudaMalloc(&c->d_fft_tmp, bytes));
  c->fft_tmp_bytes = bytes;
  if (!c->d_fft_sync) CLR_CUDA(cudaMalloc(&c->d_fft_sync, (size_t)(c->dev.n + 8) * sizeof(unsigned)));
  return 0;
}

template <int N, int G, bool MOM>
int run_yx_fused(clr_ctx *c, const float2 *src, float2 *dst, float norm, double *mom)
{
  using C = YxCfg<N, G>;
  auto k = yx_fused_kernel<N, G, MOM>;
  int grid;
  if (launch_cfg(c, k, C::THREADS, C::SMEM, 1LL << 30, &grid)) return 1;
  const int nz = c->dev.nz_here, np = nz / G;
  CLR_CHECK(nz % G == 0, "fused y+x pass: %d planes are not a multiple of the group size %d", nz, G);
  YxArgs a;
  a.src = src; a.dst = dst; a.W = c->d_twiddle; a.wn = c->dev.n;
  a.nz = nz; a.ncp = c->dev.ncp; a.nkt = c->dev.ncp / 8;
  // tickets of (grid / tickets-per-group) plane groups are in flight at any time: the x tiles trail by one more
  const int per = a.nkt + G * (N / C::XR);
  a.lag = std::min(np, grid / per + 2);
  a.norm = norm; a.mom = mom;
  a.ticket = c->d_fft_sync; a.done = c->d_fft_sync + 1;
  CLR_CUDA(cudaMemsetAsync(c->d_fft_sync, 0, (size_t)(np + 1) * sizeof(unsigned), c->stream));
  k<<<grid, C::THREADS, C::SMEM, c->stream>>>(a);
  CLR_CUDA(cudaGetLastError());
  return 0;
}

// does the fused path exist for this size? (n >= 128: both transforms need a shared-memory exchange; n = 4096: a y tile
// of 8 lines no longer fits shared memory)
template <int N> constexpr bool kHasYx = N >= 128 && N <= 2048;

template <int N>
int c2r_3d(clr_ctx *c, float2 *g, float norm, double *mom)
{
  const long long nc = c->dev.ncp;     // complex PITCH of a row (>= N/2+1, multiple of 8)
  if constexpr (kHasYx<N>) {
    if (c->fft_fused) {
      if (ensure_fft_tmp(c)) return 1;
      float2 *tmp = reinterpret_cast<float2 *>(c->d_fft_tmp);
      // z pass: lines along z, contiguous index = flattened (ky,kx) of a plane; output in the kx-tile layout
      { StageScope sc(c, "fft_z", 1);
        LineAddr ain{0, 0, (long long)N * nc, 31};
        LineAddr aout = ain;
        aout.tile8 = 1; aout.t8_g_log2 = 0; aout.t8_nkt = (int)(nc / 8); aout.t8_n = N;
        if (run_strided2<N, +1, Cfg<N>::T_Z>(c, g, tmp, ain, aout, 1, (int)(N * nc))) return 1; }
      // y + x passes, plane by plane through L2
      StageScope sc(c, "fft_yx", 1);
      if (mom) return run_yx_fused<N, 1, true>(c, tmp, g, norm, mom);
      return run_yx_fused<N, 1, false>(c, tmp, g, norm, nullptr);
    }
  }
  // z pass: lines along z, contiguous index = flattened (ky,kx) of a plane
  { StageScope sc(c, "fft_z", 1); if (run_strided<N, +1, Cfg<N>::T_Z>(c, g, 1, 0, (long long)N * nc, (int)(N * nc))) return 1; }
  // y pass: per z plane, lines along y, contiguous index = kx
  { StageScope sc(c, "fft_y", 1); if (run_strided<N, +1>(c, g, N, (long long)N * nc, nc, N / 2 + 1)) return 1; }
  // x pass: half-complex -> real
  StageScope sc(c, "fft_x", 1);
  if (mom) return run_c2r_x<N / 2, true>(c, g, (long long)N * N, (int)nc, norm, mom);
  return run_c2r_x<N / 2, false>(c, g, (long long)N * N, (int)nc, norm, nullptr);
}

template <int N>
int r2c_3d(clr_ctx *c, float2 *g)
{
  const long long nc = c->dev.ncp;     // complex PITCH of a row (>= N/2+1, multiple of 8)
  { StageScope sc(c, "fft_x", 1); if (run_r2c_x<N / 2>(c, g, (long long)N * N, (int)nc)) return 1; }
  { StageScope sc(c, "fft_y", 1); if (run_strided<N, -1>(c, g, N, (long long)N * nc, nc, N / 2 + 1)) return 1; }
  StageScope sc(c, "fft_z", 1);
  return run_strided<N, -1, Cfg<N>::T_Z>(c, g, 1, 0, (long long)N * nc, (int)(N * nc));
}

template <int N, int W>
int run_fill_c2r(clr_ctx *c, uint32_t seed, float norm, double *mom)
{
  using C = FzCfg<N, W>;
  if (ensure_fft_tmp(c)) return 1;
  FzArgs a;
  if (clr_fill_fast_setup(c, &a.k)) return 1;
  // z-pass output of delta -> scratch, of phi -> the (still unused) density grid; then phi: density grid -> potential
  // grid, delta: scratch -> density grid
  float2 *tmp = reinterpret_cast<float2 *>(c->d_fft_tmp), *dens = reinterpret_cast<float2 *>(c->d_dens),
         *npot = reinterpret_cast<float2 *>(c->d_npot);
  a.out_d = tmp; a.out_p = dens;
  a.W = c->d_twiddle; a.wn = c->dev.n; a.pkt = c->d_pkt; a.sct = c->d_sincos; a.seed = seed;
  a.n = N; a.nc = c->dev.nc; a.nyl = c->dev.nyl; a.ky0 = c->dev.ky0; a.nkt = c->dev.ncp / 8;
  { StageScope sc(c, "fill_fft_z", 1);
    auto k = fill_z_kernel<N, W>;
    int grid;
    if (launch_cfg(c, k, C::THREADS, C::SMEM, (long long)a.nkt * a.nyl * C::H, &grid)) return 1;
    k<<<grid, C::THREADS, C::SMEM, c->stream>>>(a);
    CLR_CUDA(cudaGetLastError()); }
  StageScope sc(c, "fft_yx", 2);
  if (run_yx_fused<N, 1, false>(c, dens, npot, norm, nullptr)) return 1;
  if (mom) return run_yx_fused<N, 1, true>(c, tmp, dens, norm, mom);
  return run_yx_fused<N, 1, false>(c, tmp, dens, norm, nullptr);
}

// the cluster variant: one field per CTA of a pair (fill_z_cluster_kernel)
template <int N>
int run_fill_c2r_cluster(clr_ctx *c, uint32_t seed, float norm, double *mom)
{
  using C = FzcCfg<N>;
  if (ensure_fft_tmp(c)) return 1;
  FzArgs a;
  if (clr_fill_fast_setup(c, &a.k)) return 1;
  float2 *tmp = reinterpret_cast<float2 *>(c->d_fft_tmp), *dens = reinterpret_cast<float2 *>(c->d_dens),
         *npot = reinterpret_cast<float2 *>(c->d_npot);
  a.out_d = tmp; a.out_p = dens;
  a.W = c->d_twiddle; a.wn = c->dev.n; a.pkt = c->d_pkt; a.sct = c->d_sincos; a.seed = seed;
  a.n = N; a.nc = c->dev.nc; a.nyl = c->dev.nyl; a.ky0 = c->dev.ky0; a.nkt = c->dev.ncp / 8;
  { StageScope sc(c, "fill_fft_z", 1);
    auto k = fill_z_cluster_kernel<N>;
    static int clusters = 0;                       // co-resident CTA pairs (a pair shares a GPC): asked once
    if (!clusters) {
      CLR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * c->sm_count); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      CLR_CUDA(cudaOccupancyMaxActiveClusters(&n, k, &cfg));
      CLR_CHECK(n > 0, "fill + z pass on CTA pairs does not fit this device (threads=%d smem=%zu)", C::THREADS, C::SMEM);
      clusters = n;
    }
    const long long n_tiles = (long long)a.nkt * a.nyl;
    const int grid = 2 * (int)std::min<long long>(clusters, n_tiles);
    k<<<grid, C::THREADS, C::SMEM, c->stream>>>(a);
    CLR_CUDA(cudaGetLastError()); }
  StageScope sc(c, "fft_yx", 2);
  if (run_yx_fused<N, 1, false>(c, dens, npot, norm, nullptr)) return 1;
  if (mom) return run_yx_fused<N, 1, true>(c, tmp, dens, norm, mom);
  return run_yx_fused<N, 1, false>(c, tmp, dens, norm, nullptr);
}

}  // namespace

// create_grids_fourier + both fftw_wrap_c2r of create_cartesian_fields (fourier.c:285-359, 81-102, 394-397) with the
// mode fill fused into the z pass. *ran = false when this path does not apply (exact_math, no peer-memory transpose on
// several GPUs, sizes outside [128,2048] on one GPU): the caller then runs the stand-alone fill and two transforms.
int clr_fft_fill_c2r(clr_ctx *c, uint32_t seed, double norm, double *d_moments, bool *ran)
{
  *ran = false;
  if (!c->fill_fused || !clr_fill_fast_ok(c)) return 0;
  if (c->nranks > 1) {
    // several GPUs: the fill rides on the NVLink-bound z pass of each transform (fill_peer_kernel); density first (its
    // moments), then the potential, like the separate passes
    if (!(c->p2p && c->p2p_enabled) || c->fft_overlap) return 0;
    float2 *dens = reinterpret_cast<float2 *>(c->d_dens), *npot = reinterpret_cast<float2 *>(c->d_npot);
    const FillSpec fd{seed, 0}, fp{seed, 1};
    *ran = true;
    switch (c->dev.n) {
#define CLR_FILL_DIST(NN) case NN: return c2r_3d_dist<NN>(c, dens, (float)norm, d_moments, &fd) || c2r_3d_dist<NN>(c, npot, (float)norm, nullptr, &fp);
      // (4096^3 keeps the stand-alone fill: that size has only been run through the separate passes)
      CLR_FILL_DIST(64) CLR_FILL_DIST(128) CLR_FILL_DIST(256) CLR_FILL_DIST(512) CLR_FILL_DIST(1024) CLR_FILL_DIST(2048)
#undef CLR_FILL_DIST
      default: *ran = false; return 0;
    }
  }
  if (!c->fft_fused) return 0;
  *ran = true;
  switch (c->dev.n) {
    case 128: return run_fill_c2r<128, 8>(c, seed, (float)norm, d_moments);
    case 256: return c->fill_cluster > 0 ? run_fill_c2r_cluster<256>(c, seed, (float)norm, d_moments)   // (oracle-sized tests)
                     : c->fill_w == 4 ? run_fill_c2r<256, 4>(c, seed, (float)norm, d_moments)
                                      : run_fill_c2r<256, 8>(c, seed, (float)norm, d_moments);
    case 512: return run_fill_c2r<512, 8>(c, seed, (float)norm, d_moments);
    // option "fill_w": 4 = half-width tiles (two CTAs per SM: the fill of one overlaps the butterflies / stores of the other)
    case 1024: return c->fill_cluster > 0 ? run_fill_c2r_cluster<1024>(c, seed, (float)norm, d_moments)
                      : c->fill_w == 4 ? run_fill_c2r<1024, 4>(c, seed, (float)norm, d_moments)
                                       : run_fill_c2r<1024, 8>(c, seed, (float)norm, d_moments);
    // 2048: two fields x 8 lines x 2048 points do not fit the shared memory of one SM: a CTA pair with one field each
    // (option "fill_cluster" = 0: 4-line tiles on single CTAs instead)
    case 2048: return c->fill_cluster != 0 ? run_fill_c2r_cluster<2048>(c, seed, (float)norm, d_moments)
                                           : run_fill_c2r<2048, 4>(c, seed, (float)norm, d_moments);
    default: *ran = false; return 0;
  }
}

// norm multiplies the output (1.0 = plain fftw_wrap_c2r); d_moments != NULL accumulates
// {sum, sum of squares} of the scaled output over the unpadded cells.
int clr_fft_c2r_impl(clr_ctx *c, float *grid, double norm, double *d_moments)
{
  float2 *g = reinterpret_cast<float2 *>(grid);
  if (c->nranks > 1) {
    CLR_CHECK(c->d_stage, "distributed FFT: no staging buffer (clr_comm_init first)");
    switch (c->dev.n) {
      case 64: return c2r_3d_dist<64>(c, g, (float)norm, d_moments);
      case 128: return c2r_3d_dist<128>(c, g, (float)norm, d_moments);
      case 256: return c2r_3d_dist<256>(c, g, (float)norm, d_moments);
      case 512: return c2r_3d_dist<512>(c, g, (float)norm, d_moments);
      case 1024: return c2r_3d_dist<1024>(c, g, (float)norm, d_moments);
      case 2048: return c2r_3d_dist<2048>(c, g, (float)norm, d_moments);
      case 4096: return c2r_3d_dist<4096>(c, g, (float)norm, d_moments);
      default: clr_set_error("n_grid=%d: the distributed FFT supports powers of two in [64,4096]", c->dev.n); return 1;
    }
  }
  switch (c->dev.n) {
    case 16: return c2r_3d<16>(c, g, (float)norm, d_moments);
    case 32: return c2r_3d<32>(c, g, (float)norm, d_moments);
    case 64: return c2r_3d<64>(c, g, (float)norm, d_moments);
    case 128: return c2r_3d<128>(c, g, (float)norm, d_moments);
    case 256: return c2r_3d<256>(c, g, (float)norm, d_moments);
    case 512: return c2r_3d<512>(c, g, (float)norm, d_moments);
    case 1024: return c2r_3d<1024>(c, g, (float)norm, d_moments);
    case 2048: return c2r_3d<2048>(c, g, (float)norm, d_moments);
    case 4096: return c2r_3d<4096>(c, g, (float)norm, d_moments);
    default:
      if (clr_fft_generic_ok(c->dev.n)) return clr_fft_generic_c2r(c, g, (float)norm, d_moments);
      clr_set_error("n_grid=%d: the FFT takes multiples of 4 in [16,4096] without prime factors above 31", c->dev.n); return 1;
  }
}

int clr_fft_r2c_impl(clr_ctx *c, float *grid)
{
  float2 *g = reinterpret_cast<float2 *>(grid);
  if (c->nranks > 1) {
    CLR_CHECK(c->d_stage, "distributed FFT: no staging buffer (clr_comm_init first)");
    switch (c->dev.n) {
      case 64: return r2c_3d_dist<64>(c, g);
      case 128: return r2c_3d_dist<128>(c, g);
      case 256: return r2c_3d_dist<256>(c, g);
      case 512: return r2c_3d_dist<512>(c, g);
      case 1024: return r2c_3d_dist<1024>(c, g);
      case 2048: return r2c_3d_dist<2048>(c, g);
      case 4096: return r2c_3d_dist<4096>(c, g);
      default: clr_set_error("n_grid=%d: the distributed FFT supports powers of two in [64,4096]", c->dev.n); return 1;
    }
  }
  switch (c->dev.n) {
    case 16: return r2c_3d<16>(c, g);
    case 32: return r2c_3d<32>(c, g);
    case 64: return r2c_3d<64>(c, g);
    case 128: return r2c_3d<128>(c, g);
    case 256: return r2c_3d<256>(c, g);
    case 512: return r2c_3d<512>(c, g);
    case 1024: return r2c_3d<1024>(c, g);
    case 2048: return r2c_3d<2048>(c, g);
    case 4096: return r2c_3d<4096>(c, g);
    default:
      if (clr_fft_generic_ok(c->dev.n)) return clr_fft_generic_r2c(c, g);
      clr_set_error("n_grid=%d: the FFT takes multiples of 4 in [16,4096] without prime factors above 31", c->dev.n); return 1;
  }
}
