// Access-pattern microbenchmarks behind the FFT design decisions of round 2.
// Not part of the product: a stand-alone binary run once on the GPU box.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/membench tools/membench.cu
//   tools/_build/membench > gpurun_out/membench.jsonl
//
// Patterns (complex64 grid [nz][n][ncp], ncp = n/2+2: the 16-byte aligned internal pitch):
//  1. z-pass tiles: T consecutive points of a plane x all nz planes, read + write in place, plain LDG/STG at high
//     occupancy = what DRAM gives for runs of 8*T bytes one plane stride apart (the FFT z pass cannot beat it).
//  2. the same tiles moved by TMA box loads / bulk tensor stores through shared memory (no LSU instructions).
//  3. y-pass tiles (T columns x n rows of one plane), x-pass row blocks, and the two FUSED in one persistent kernel in
//     ticket order Y(z+LAG) before X(z): does the 126 MB L2 forward the y-pass output to the x pass?
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <functional>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

// ---------------------------------------------------------------------------------------------------
// 1. strided tile copy with LDG/STG. tile = T complex (8T bytes) x rows, row stride `rs` (float4 units), tiles laid
// out next to each other along the contiguous index. MODE 0: read+write, 1: read only, 2: write only.
template <int T, int MODE>
__global__ void __launch_bounds__(256) tile_copy_kernel(float4 *g, long long rs, int rows, long long n_tiles, float *sink)
{
  constexpr int C = T / 2;                       // float4 per run
  constexpr int RPI = 256 / C;                   // rows per iteration of the CTA
  const int c = threadIdx.x % C, r0 = threadIdx.x / C;
  float acc = 0.f;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    float4 *base = g + tile * C + c;
    for (int r = r0; r < rows; r += RPI * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int rr = r + u * RPI;
        if (MODE != 2) v[u] = rr < rows ? __ldcg(base + (long long)rr * rs) : make_float4(0, 0, 0, 0);
        else v[u] = make_float4((float)rr, 1.f, 2.f, 3.f);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int rr = r + u * RPI;
        if (MODE == 1) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        else if (rr < rows) { v[u].x += 1.f; __stcg(base + (long long)rr * rs, v[u]); }
      }
    }
  }
  if (MODE == 1 && acc == 1.2345f) *sink = acc;
}

__global__ void __launch_bounds__(256) stream_kernel(float4 *g, long long n4)
{
  const long long stride = (long long)gridDim.x * 256 * 4;
  for (long long i = (long long)blockIdx.x * 1024 + threadIdx.x; i < n4; i += stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = i + u * 256 < n4 ? __ldcg(g + i + u * 256) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; u++) if (i + u * 256 < n4) { v[u].x += 1.f; __stcg(g + i + u * 256, v[u]); }
  }
}

// ---------------------------------------------------------------------------------------------------
// 2. TMA: box {2T floats, BR rows}; a ring of stages, loads signal an mbarrier, the same thread stores the stage back
// with a bulk tensor store. One producer thread per CTA does everything (pure DMA copy).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int T, int BR, int STAGES>
__global__ void __launch_bounds__(32) tile_copy_tma_kernel(const __grid_constant__ CUtensorMap map, int rows, long long n_tiles)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t STAGE_BYTES = 8u * T * BR;
  constexpr int BAR_BYTES = (STAGES * 8 + 127) / 128 * 128;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  unsigned char *buf = smem_raw + BAR_BYTES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  const int boxes_per_tile = rows / BR;
  const long long my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_box = my_tiles * boxes_per_tile;
  // issue-ahead ring: load box i+STAGES-1 while storing box i
  auto coords = [&](long long i, int &c0, int &c1) {
    long long t = blockIdx.x + (i / boxes_per_tile) * gridDim.x;
    c0 = (int)(t * 2 * T);
    c1 = (int)(i % boxes_per_tile) * BR;
  };
  for (long long j = 0; j < n_box && j < STAGES; j++) {
    int c0, c1; coords(j, c0, c1);
    mbar_expect_tx(bar + j, STAGE_BYTES);
    tma_load_2d(buf + j * STAGE_BYTES, &map, c0, c1, bar + j);
  }
  for (long long i = 0; i < n_box; i++) {
    const int s = (int)(i % STAGES);
    mbar_wait(bar + s, (uint32_t)((i / STAGES) & 1));
    int c0, c1; coords(i, c0, c1);
    tma_store_2d(&map, buf + s * STAGE_BYTES, c0, c1);
    tma_commit();
    if (i >= 1) {
      tma_wait_read<1>();                      // the store of box i-1 has finished reading its stage: refill it
      const long long j = i - 1 + STAGES;
      if (j < n_box) {
        const int s2 = (int)(j % STAGES);
        coords(j, c0, c1);
        mbar_expect_tx(bar + s2, STAGE_BYTES);
        tma_load_2d(buf + s2 * STAGE_BYTES, &map, c0, c1, bar + s2);
      }
    }
  }
  tma_wait_read<0>();
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// 3. the planned c2r data flow. src = z-pass output in the tile layout [z/G][kxt][ky][G][8 kx] (a y tile = one
// CONTIGUOUS block of n*G*64 bytes), dst = final layout [z][ky][ncp]. Y tile: contiguous read, 64-byte runs written to
// dst (meant to stay in L2); X block: XR rows of one plane, read (from L2, hopefully) + written in place.
template <int G>
__device__ __forceinline__ void y_tile_copy(const float4 *src, float4 *dst, int n, int ncp4, int nkt, int p, int kxt)
{
  // block of n*G*4 float4; element i = (ky*G + zz)*4 + q
  const float4 *blk = src + ((long long)p * nkt + kxt) * n * G * 4;
  const int total = n * G * 4;
  for (int i = threadIdx.x; i < total; i += blockDim.x * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { const int ii = i + u * blockDim.x; v[u] = ii < total ? __ldcs(blk + ii) : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int ii = i + u * blockDim.x;
      if (ii < total) {
        const int q = ii & 3, zz = (ii >> 2) % G, ky = (ii >> 2) / G;
        v[u].x += 1.f;
        __stcg(dst + ((long long)(p * G + zz) * n + ky) * ncp4 + kxt * 4 + q, v[u]);
      }
    }
  }
}
template <int XR>
__device__ __forceinline__ void x_block_copy(float4 *plane, int n, int ncp4, int blk)
{
  float4 *base = plane + (long long)blk * XR * ncp4;
  const int total = XR * ncp4;
  for (int i = threadIdx.x; i < total; i += blockDim.x * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { const int ii = i + u * blockDim.x; v[u] = ii < total ? __ldcg(base + ii) : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < 8; u++) { const int ii = i + u * blockDim.x; if (ii < total) { v[u].y += 1.f; __stcs(base + ii, v[u]); } }
  }
}

template <int G, int XR>
__global__ void __launch_bounds__(512) pass_kernel(const float4 *src, float4 *dst, int n, int ncp4, int nz, int which)
{
  const int nkt = ncp4 / 4, nx = n / XR;
  const long long plane = (long long)n * ncp4;
  const long long total = which == 0 ? (long long)(nz / G) * nkt : (long long)nz * nx;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    if (which == 0) y_tile_copy<G>(src, dst, n, ncp4, nkt, (int)(t / nkt), (int)(t % nkt));
    else x_block_copy<XR>(dst + (t / nx) * plane, n, ncp4, (int)(t % nx));
  }
}

template <int G, int XR>
__global__ void __launch_bounds__(512) fused_kernel(const float4 *src, float4 *dst, int n, int ncp4, int nz, int lag, unsigned *ticket, unsigned *done)
{
  const int nkt = ncp4 / 4, nx = n / XR, np = nz / G;
  const long long plane = (long long)n * ncp4;
  const unsigned per = nkt + G * nx, total = (unsigned)(np + lag) * per;
  __shared__ unsigned s_t;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned t = s_t;
    if (t >= total) break;
    const int s = (int)(t / per), r = (int)(t % per);
    if (r < nkt) {
      if (s >= np) continue;
      y_tile_copy<G>(src, dst, n, ncp4, nkt, s, r);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(done + s, 1u);
    } else {
      const int p = s - lag;
      if (p < 0) continue;
      if (threadIdx.x == 0) {
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done + p) : "memory"); if (v < (unsigned)nkt) __nanosleep(200); } while (v < (unsigned)nkt);
      }
      __syncthreads();
      const int xb = r - nkt;
      x_block_copy<XR>(dst + (long long)(p * G + xb / nx) * plane, n, ncp4, xb % nx);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float time_ms(int reps, const std::function<void()> &f)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();                                    // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; i++) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

template <int T> void run_tile_copy(float4 *g, int n, int ncp, int nz, int sms)
{
  const long long rs = (long long)n * ncp / 2;                  // plane stride in float4
  const long long n_tiles = (long long)n * ncp / T;
  const double bytes = 8.0 * n * ncp * nz;
  float *sink; CK(cudaMalloc(&sink, 4));
  float ms0 = time_ms(3, [&] { tile_copy_kernel<T, 0><<<sms * 8, 256>>>(g, rs, nz, n_tiles, sink); });
  float ms1 = time_ms(3, [&] { tile_copy_kernel<T, 1><<<sms * 8, 256>>>(g, rs, nz, n_tiles, sink); });
  float ms2 = time_ms(3, [&] { tile_copy_kernel<T, 2><<<sms * 8, 256>>>(g, rs, nz, n_tiles, sink); });
  printf("{\"test\":\"zpass_ldg\",\"n\":%d,\"nz\":%d,\"run_bytes\":%d,\"rw_ms\":%.3f,\"rw_GBps\":%.0f,\"r_ms\":%.3f,\"r_GBps\":%.0f,\"w_ms\":%.3f,\"w_GBps\":%.0f}\n",
         n, nz, 8 * T, ms0, 2 * bytes / ms0 * 1e-6, ms1, bytes / ms1 * 1e-6, ms2, bytes / ms2 * 1e-6);
  fflush(stdout);
  CK(cudaFree(sink));
}

template <int T, int BR, int STAGES> void run_tile_tma(EncodeFn enc, float4 *g, int n, int ncp, int nz, int sms, int ctas_per_sm,
                                                        CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE)
{
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)2 * n * ncp, (cuuint64_t)nz};
  cuuint64_t strides[1] = {(cuuint64_t)8 * n * ncp};
  cuuint32_t box[2] = {2 * T, BR};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("{\"test\":\"zpass_tma\",\"error\":\"encode %d\"}\n", (int)r); return; }
  const long long n_tiles = (long long)n * ncp / T;
  const double bytes = 8.0 * n * ncp * nz;
  size_t smem = (STAGES * 8 + 127) / 128 * 128 + (size_t)STAGES * 8 * T * BR;
  auto k = tile_copy_tma_kernel<T, BR, STAGES>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float ms = time_ms(3, [&] { k<<<sms * ctas_per_sm, 32, smem>>>(map, nz, n_tiles); });
  printf("{\"test\":\"zpass_tma\",\"n\":%d,\"nz\":%d,\"run_bytes\":%d,\"box_rows\":%d,\"stages\":%d,\"ctas_per_sm\":%d,\"l2promo\":%d,\"smem\":%zu,\"rw_ms\":%.3f,\"rw_GBps\":%.0f}\n",
         n, nz, 8 * T, BR, STAGES, ctas_per_sm, (int)promo, smem, ms, 2 * bytes / ms * 1e-6);
  fflush(stdout);
}

template <int G, int XR> void run_fused(const float4 *src, float4 *dst, int n, int nz, int sms, int threads)
{
  const int ncp = (n / 2 + 1 + 7) / 8 * 8, ncp4 = ncp / 2;
  const double bytes = 8.0 * n * ncp * nz;
  unsigned *ticket, *done;
  CK(cudaMalloc(&ticket, 4)); CK(cudaMalloc(&done, 4 * (nz + 64)));
  const int cps = threads == 512 ? 2 : 4;
  float msy = time_ms(3, [&] { pass_kernel<G, XR><<<sms * cps, threads>>>(src, dst, n, ncp4, nz, 0); });
  float msx = time_ms(3, [&] { pass_kernel<G, XR><<<sms * cps, threads>>>(src, dst, n, ncp4, nz, 1); });
  printf("{\"test\":\"yx_separate\",\"n\":%d,\"nz\":%d,\"G\":%d,\"threads\":%d,\"y_ms\":%.3f,\"y_GBps\":%.0f,\"x_ms\":%.3f,\"x_GBps\":%.0f}\n", n, nz, G, threads,
         msy, 2 * bytes / msy * 1e-6, msx, 2 * bytes / msx * 1e-6);
  fflush(stdout);
  for (int c = 1; c <= cps; c *= 2)
    for (int lag : {1, 2, 4, 8, 16}) {
      if ((double)lag * G * 8.0 * n * ncp > 90e6) continue;
      float ms = time_ms(3, [&] {
        CK(cudaMemsetAsync(ticket, 0, 4)); CK(cudaMemsetAsync(done, 0, 4 * (nz + 64)));
        fused_kernel<G, XR><<<sms * c, threads>>>(src, dst, n, ncp4, nz, lag, ticket, done);
      });
      printf("{\"test\":\"yx_fused\",\"n\":%d,\"nz\":%d,\"G\":%d,\"threads\":%d,\"ctas_per_sm\":%d,\"lag_pairs\":%d,\"ms\":%.3f,\"alg_GBps\":%.0f,\"vs_separate\":%.3f}\n", n,
             nz, G, threads, c, lag, ms, 4 * bytes / ms * 1e-6, ms / (msx + msy));
      fflush(stdout);
    }
  CK(cudaFree(ticket)); CK(cudaFree(done));
}

int main()
{
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  printf("{\"test\":\"device\",\"name\":\"%s\",\"sms\":%d,\"l2_bytes\":%d}\n", prop.name, sms, prop.l2CacheSize);
  EncodeFn enc = nullptr;
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    enc = (EncodeFn)fn;
  }
  // one 1024^3 half-spectrum grid with the aligned pitch (4.4 GB) + a second one for the out-of-place flow
  {
    const int n = 1024, ncp = (n / 2 + 1 + 7) / 8 * 8, nz = 1024;
    float4 *g, *g2;
    CK(cudaMalloc(&g, (size_t)8 * n * ncp * nz));
    CK(cudaMalloc(&g2, (size_t)8 * n * ncp * nz));
    CK(cudaMemset(g, 0, (size_t)8 * n * ncp * nz));
    CK(cudaMemset(g2, 0, (size_t)8 * n * ncp * nz));
    {
      const long long n4 = (long long)n * ncp * nz / 2;
      float ms = time_ms(3, [&] { stream_kernel<<<sms * 8, 256>>>(g, n4); });
      printf("{\"test\":\"stream_inplace\",\"rw_ms\":%.3f,\"rw_GBps\":%.0f}\n", ms, 32.0 * n4 / ms * 1e-6);
    }
    run_fused<1, 16>(g, g2, n, nz, sms, 256);
    run_fused<2, 32>(g, g2, n, nz, sms, 512);
    if (enc) {
      run_tile_tma<16, 32, 24>(enc, g, n, ncp, nz, sms, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
      run_tile_tma<16, 32, 48>(enc, g, n, ncp, nz, sms, 1);      // 192 KB ring
      run_tile_tma<16, 32, 48>(enc, g, n, ncp, nz, sms, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
      run_tile_tma<16, 256, 3>(enc, g, n, ncp, nz, sms, 2);
      run_tile_tma<32, 128, 6>(enc, g, n, ncp, nz, sms, 1);      // 256-byte runs
    }
    CK(cudaFree(g)); CK(cudaFree(g2));
  }
  // 2048: planes of 16.9 MB; 128 planes (2.2 GB per buffer)
  {
    const int n = 2048, ncp = (n / 2 + 1 + 7) / 8 * 8, nz = 128;
    float4 *g, *g2;
    CK(cudaMalloc(&g, (size_t)8 * n * ncp * nz));
    CK(cudaMalloc(&g2, (size_t)8 * n * ncp * nz));
    CK(cudaMemset(g, 0, (size_t)8 * n * ncp * nz));
    CK(cudaMemset(g2, 0, (size_t)8 * n * ncp * nz));
    run_fused<1, 8>(g, g2, n, nz, sms, 512);
    CK(cudaFree(g)); CK(cudaFree(g2));
  }
  return 0;
}
