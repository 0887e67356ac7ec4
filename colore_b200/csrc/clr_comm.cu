// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch. Replaces the reference's MPI
// layer (common.c:216-274 and the call sites listed in SURVEY.md section 2.3):
//   FFTW-MPI transposes        -> one grouped ncclSend/ncclRecv all-to-all per transform (clr_fft.cu)
//   MPI_Allreduce (2 doubles)  -> ncclAllReduce                      (fourier.c:69-70)
//   MPI_Sendrecv z-halo        -> ncclSend/ncclRecv pair per side    (fourier.c:406-410)
//   MPI_Allreduce histograms   -> ncclAllReduce                      (density.c:1262-1269)
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a torchrun process this is the
// copy PyTorch already loaded, and a single-GPU build never needs the library at all.
#include "clr_internal.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
  if (g_nccl.h) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  CLR_CHECK(h, "cannot load libnccl.so.2: %s", dlerror());
#define CLR_SYM(name)                                                        \
  *(void **)(&g_nccl.name) = dlsym(h, "nccl" #name);                         \
  CLR_CHECK(g_nccl.name, "libnccl lacks symbol nccl" #name)
  CLR_SYM(GetUniqueId); CLR_SYM(CommInitRank); CLR_SYM(CommDestroy); CLR_SYM(Send); CLR_SYM(Recv);
  CLR_SYM(AllReduce); CLR_SYM(AllGather); CLR_SYM(GroupStart); CLR_SYM(GroupEnd); CLR_SYM(GetErrorString);
#undef CLR_SYM
  g_nccl.h = h;
  return 0;
}

#define CLR_NCCL(call)                                                                          \
  do {                                                                                          \
    ncclResult_t r_ = (call);                                                                   \
    if (r_ != ncclSuccess) {                                                                    \
      clr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));   \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

}  // namespace

extern "C" int clr_comm_unique_id(void *id128)
{
  if (load_nccl()) return 1;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  CLR_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int clr_comm_init(clr_ctx *c, int rank, int nranks, const void *id128)
{
  CLR_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "clr_comm_init: bad rank %d of %d", rank, nranks);
  c->rank = rank; c->nranks = nranks;
  ClrDev &d = c->dev;
  if (nranks == 1) { d.nyl = d.n; d.ky0 = 0; return 0; }
  CLR_CHECK(d.n % nranks == 0, "n_grid=%d is not divisible by the number of GPUs %d", d.n, nranks);
  CLR_CHECK(d.log2n >= 6 && d.n <= 4096, "n_grid=%d: the distributed FFT takes powers of two in [64,4096]", d.n);
  CLR_CHECK(d.nz_here == d.n / nranks && d.iz0_here == rank * (d.n / nranks),
            "slab bounds (nz_here=%d, iz0_here=%d) do not match rank %d of %d", d.nz_here, d.iz0_here, rank, nranks);
  if (load_nccl()) return 1;
  CLR_CUDA(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  CLR_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  c->nccl_comm = comm;
  d.nyl = d.n / nranks;
  d.ky0 = rank * d.nyl;
  // staging buffers of the FFT slab transpose: one slab each (+ padding: the tile-major layout of the fused c2r transpose
  // rounds every source block up to whole tiles), followed by the flag words of the peer-memory barriers. The second
  // buffer belongs to the potential's pipeline (fft_overlap); it is skipped when memory is short (4096^3 on 8 GPUs).
  size_t floats = (size_t)d.pitch * d.n * d.nz_here + (size_t)nranks * d.nz_here * 64 * 2;
  floats = (floats + 63) & ~(size_t)63;
  c->stage_floats = floats;
  const size_t bytes = floats * sizeof(float) + 4096;
  CLR_CUDA(cudaMalloc(&c->d_stage, bytes));
  CLR_CUDA(cudaMemset(reinterpret_cast<char *>(c->d_stage) + floats * sizeof(float), 0, 4096));
  c->sets[0].stage = c->d_stage;
  {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t slab = (size_t)d.pitch * d.n * d.nz_here * sizeof(float);
    if (c->fft_overlap && free_b > bytes + 2 * slab + (8ULL << 30)) {       // leave room for counts + catalogue
      CLR_CUDA(cudaMalloc(&c->sets[1].stage, bytes));
      CLR_CUDA(cudaMemset(reinterpret_cast<char *>(c->sets[1].stage) + floats * sizeof(float), 0, 4096));
    }
  }
  CLR_CUDA(cudaMalloc(&c->d_barrier, sizeof(int)));
  CLR_CUDA(cudaMemset(c->d_barrier, 0, sizeof(int)));
  // Peer mapping of the staging buffers (CUDA IPC; the GPUs of one node see each other over NVLink / NVSwitch).
  // Handles travel through an ncclAllGather. If any rank cannot map a peer, every rank falls back to the NCCL
  // all-to-all (the decision is all-reduced so that the ranks never disagree).
  int ok = nranks <= CLR_MAX_PEERS ? 1 : 0;
  int have2 = c->sets[1].stage ? 1 : 0;
  for (int b = 0; b < 2; b++) {
    if (!c->sets[b].stage) continue;
    cudaIpcMemHandle_t mine;
    if (ok && cudaIpcGetMemHandle(&mine, c->sets[b].stage) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    unsigned char *d_h = nullptr;
    CLR_CUDA(cudaMalloc(&d_h, (size_t)nranks * sizeof(cudaIpcMemHandle_t)));
    CLR_CUDA(cudaMemset(d_h, 0, (size_t)nranks * sizeof(cudaIpcMemHandle_t)));
    if (ok) CLR_CUDA(cudaMemcpy(d_h + (size_t)rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    CLR_NCCL(g_nccl.AllGather(d_h + (size_t)rank * sizeof(mine), d_h, sizeof(mine), ncclChar, comm, c->stream));
    CLR_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<cudaIpcMemHandle_t> all(nranks);
    CLR_CUDA(cudaMemcpy(all.data(), d_h, (size_t)nranks * sizeof(mine), cudaMemcpyDeviceToHost));
    cudaFree(d_h);
    for (int h = 0; h < nranks && ok; h++) {
      if (h == rank) { c->sets[b].peer[h] = c->sets[b].stage; }
      else {
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[h], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        c->sets[b].peer[h] = static_cast<float *>(ptr);
      }
      c->sets[b].flag_peer[h] = reinterpret_cast<unsigned *>(c->sets[b].peer[h] + floats);
    }
  }
  for (int h = 0; h < nranks; h++) c->peer_stage[h] = c->sets[0].peer[h];
  // every rank must have the second pipeline, or none uses it
  CLR_CUDA(cudaMemcpy(c->d_barrier, &have2, sizeof(int), cudaMemcpyHostToDevice));
  CLR_NCCL(g_nccl.AllReduce(c->d_barrier, c->d_barrier, 1, ncclInt32, ncclMin, comm, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  CLR_CUDA(cudaMemcpy(&have2, c->d_barrier, sizeof(int), cudaMemcpyDeviceToHost));
  CLR_CUDA(cudaMemcpy(c->d_barrier, &ok, sizeof(int), cudaMemcpyHostToDevice));
  CLR_NCCL(g_nccl.AllReduce(c->d_barrier, c->d_barrier, 1, ncclInt32, ncclMin, comm, c->stream));
  CLR_CUDA(cudaStreamSynchronize(c->stream));
  CLR_CUDA(cudaMemcpy(&ok, c->d_barrier, sizeof(int), cudaMemcpyDeviceToHost));
  c->p2p = ok != 0;
  c->flag_barrier = c->p2p && have2;                 // only the two-pipeline mode needs barriers that are not NCCL calls
  if (!(c->p2p && have2)) {                          // no second pipeline: everything runs on the main stream
    if (c->sets[1].stage) {
      for (int h = 0; h < nranks; h++)
        if (c->sets[1].peer[h] && c->sets[1].peer[h] != c->sets[1].stage) cudaIpcCloseMemHandle(c->sets[1].peer[h]);
      cudaFree(c->sets[1].stage);
    }
    c->sets[1] = clr_ctx::StageSet();
  } else {
    CLR_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CLR_CUDA(cudaEventCreateWithFlags(&c->ev_z_done, cudaEventDisableTiming));
    CLR_CUDA(cudaEventCreateWithFlags(&c->ev_npot, cudaEventDisableTiming));
  }
  return 0;
}

// select the staging buffer + barrier flags (clr_fft.cu works on c->d_stage / c->peer_stage)
void clr_use_set(clr_ctx *c, int set)
{
  c->cur_set = set;
  c->d_stage = c->sets[set].stage;
  for (int h = 0; h < CLR_MAX_PEERS; h++) c->peer_stage[h] = c->sets[set].peer[h];
}

namespace {
// Barrier over all ranks, ordered in the stream: thread h tells rank h "I have reached epoch e" with a release store into
// h's flag array (peer memory), then waits until rank h has told me the same. Everything queued before the barrier on
// the peers' streams (their stores into my staging buffer) has completed when it returns.
struct FlagPtrs { unsigned *p[CLR_MAX_PEERS]; };
__global__ void flag_barrier_kernel(FlagPtrs f, int rank, int nranks, unsigned epoch)
{
  const int h = threadIdx.x;
  if (h >= nranks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[h] + rank), "r"(epoch) : "memory");
  unsigned v;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.p[rank] + h) : "memory");
    if ((int)(v - epoch) >= 0) break;
    __nanosleep(200);
  }
}
}  // namespace


extern "C" int clr_comm_p2p(clr_ctx *c) { return c->p2p && c->p2p_enabled ? 1 : 0; }

// all ranks have finished everything queued on their streams (of the current pipeline) so far
int clr_comm_barrier(clr_ctx *c)
{
  if (c->nranks == 1) return 0;
  if (c->flag_barrier) {
    clr_ctx::StageSet &S = c->sets[c->cur_set];
    FlagPtrs f;
    for (int h = 0; h < CLR_MAX_PEERS; h++) f.p[h] = S.flag_peer[h];
    flag_barrier_kernel<<<1, 32, 0, c->stream>>>(f, c->rank, c->nranks, ++S.epoch);
    CLR_CUDA(cudaGetLastError());
    return 0;
  }
  CLR_CHECK(c->cur_set == 0, "the potential pipeline needs peer-memory barriers");
  CLR_NCCL(g_nccl.AllReduce(c->d_barrier, c->d_barrier, 1, ncclInt32, ncclMax, (ncclComm_t)c->nccl_comm, c->stream));
  return 0;
}

// Agreement before a collective: returns 0 when EVERY rank passed ok != 0, else sets the error on all ranks and returns 1
// (a rank that failed locally, e.g. out of memory, must not leave the others waiting in the next collective).
int clr_comm_all_ok(clr_ctx *c, int ok, const char *what)
{
  if (c->nranks > 1) {
    int v = ok ? 1 : 0;
    CLR_CUDA(cudaMemcpyAsync(c->d_barrier, &v, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CLR_NCCL(g_nccl.AllReduce(c->d_barrier, c->d_barrier, 1, ncclInt32, ncclMin, (ncclComm_t)c->nccl_comm, c->stream));
    CLR_CUDA(cudaMemcpyAsync(&v, c->d_barrier, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CLR_CUDA(cudaStreamSynchronize(c->stream));
    if (ok && !v) clr_set_error("%s: another rank failed", what);
    ok = v;
  }
  return ok ? 0 : 1;
}

int clr_comm_destroy(clr_ctx *c)
{
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
  for (int b = 0; b < 2; b++) {
    for (int h = 0; h < CLR_MAX_PEERS; h++)
      if (c->sets[b].peer[h] && c->sets[b].peer[h] != c->sets[b].stage) cudaIpcCloseMemHandle(c->sets[b].peer[h]);
    if (c->sets[b].stage) cudaFree(c->sets[b].stage);
    c->sets[b] = clr_ctx::StageSet();
  }
  for (int h = 0; h < CLR_MAX_PEERS; h++) c->peer_stage[h] = nullptr;
  c->d_stage = nullptr;
  c->p2p = false;
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->ev_z_done) cudaEventDestroy(c->ev_z_done);
  if (c->ev_npot) cudaEventDestroy(c->ev_npot);
  c->stream2 = nullptr; c->ev_z_done = nullptr; c->ev_npot = nullptr;
  if (c->d_barrier) cudaFree(c->d_barrier);
  c->d_barrier = nullptr;
  return 0;
}

// block b of `send` (block_floats floats) goes to rank b; block s of `recv` comes from rank s
int clr_comm_alltoall(clr_ctx *c, const void *send, void *recv, size_t block_floats)
{
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  const float *s = static_cast<const float *>(send);
  float *r = static_cast<float *>(recv);
  CLR_CUDA(cudaMemcpyAsync(r + (size_t)c->rank * block_floats, s + (size_t)c->rank * block_floats,
                           block_floats * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  CLR_NCCL(g_nccl.GroupStart());
  for (int k = 1; k < c->nranks; k++) {
    int to = (c->rank + k) % c->nranks, from = (c->rank - k + c->nranks) % c->nranks;
    CLR_NCCL(g_nccl.Send(s + (size_t)to * block_floats, block_floats, ncclFloat, to, comm, c->stream));
    CLR_NCCL(g_nccl.Recv(r + (size_t)from * block_floats, block_floats, ncclFloat, from, comm, c->stream));
  }
  CLR_NCCL(g_nccl.GroupEnd());
  c->a2a_bytes += (double)block_floats * sizeof(float) * (c->nranks - 1);
  return 0;
}

// variable-size exchange (floats): segment h of `send` goes to rank h, segment s of `recv` comes from rank s;
// replaces the MPI_Sendrecv ring of share_particles (density.c:300-360)
int clr_comm_alltoallv(clr_ctx *c, const float *send, const size_t *send_off, const size_t *send_n, float *recv,
                       const size_t *recv_off, const size_t *recv_n)
{
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  CLR_NCCL(g_nccl.GroupStart());
  for (int k = 1; k < c->nranks; k++) {
    int to = (c->rank + k) % c->nranks, from = (c->rank - k + c->nranks) % c->nranks;
    if (send_n[to]) CLR_NCCL(g_nccl.Send(send + send_off[to], send_n[to], ncclFloat, to, comm, c->stream));
    if (recv_n[from]) CLR_NCCL(g_nccl.Recv(recv + recv_off[from], recv_n[from], ncclFloat, from, comm, c->stream));
  }
  CLR_NCCL(g_nccl.GroupEnd());
  return 0;
}

int clr_comm_allreduce_f64(clr_ctx *c, double *dbuf, size_t n)
{
  if (c->nranks == 1) return 0;
  CLR_NCCL(g_nccl.AllReduce(dbuf, dbuf, n, ncclDouble, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  return 0;
}

// map reductions (io.c:727-735, 836-842, 974-980: MPI_Reduce of the shells): slab-local partial maps summed
int clr_comm_allreduce_f32(clr_ctx *c, float *dbuf, size_t n)
{
  if (c->nranks == 1) return 0;
  CLR_NCCL(g_nccl.AllReduce(dbuf, dbuf, n, ncclFloat, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  return 0;
}
int clr_comm_allreduce_i32(clr_ctx *c, int *dbuf, size_t n)
{
  if (c->nranks == 1) return 0;
  CLR_NCCL(g_nccl.AllReduce(dbuf, dbuf, n, ncclInt32, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  return 0;
}

int clr_comm_allreduce_u64(clr_ctx *c, unsigned long long *dbuf, size_t n)
{
  if (c->nranks == 1) return 0;
  CLR_NCCL(g_nccl.AllReduce(dbuf, dbuf, n, ncclUint64, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  return 0;
}

// z-halo of the potential (fourier.c:401-414): my last plane -> right neighbour's slice_left, my first plane -> left
// neighbour's slice_right; plus the SECOND ring (my second-to-last / second plane), stored behind the first one, which
// the CIC velocity stencil of sources in the first / last plane of a slab reaches (srcs.c:486-504 gets them through the
// slab rotation of beaming.c:325-352)
int clr_comm_halo(clr_ctx *c)
{
  const ClrDev &d = c->dev;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  size_t plane = (size_t)d.pitch * d.n;
  int right = (c->rank + 1) % c->nranks, left = (c->rank - 1 + c->nranks) % c->nranks;
  float *slice_left = c->d_npot + plane * d.nz_here, *slice_right = slice_left + plane;
  float *left2 = slice_right + plane, *right2 = left2 + plane;
  const bool two = d.nz_here >= 2;
  CLR_NCCL(g_nccl.GroupStart());
  CLR_NCCL(g_nccl.Send(c->d_npot + plane * (d.nz_here - 1), plane, ncclFloat, right, comm, c->stream));
  CLR_NCCL(g_nccl.Recv(slice_left, plane, ncclFloat, left, comm, c->stream));
  CLR_NCCL(g_nccl.Send(c->d_npot, plane, ncclFloat, left, comm, c->stream));
  CLR_NCCL(g_nccl.Recv(slice_right, plane, ncclFloat, right, comm, c->stream));
  if (two) {
    CLR_NCCL(g_nccl.Send(c->d_npot + plane * (d.nz_here - 2), plane, ncclFloat, right, comm, c->stream));
    CLR_NCCL(g_nccl.Recv(left2, plane, ncclFloat, left, comm, c->stream));
    CLR_NCCL(g_nccl.Send(c->d_npot + plane, plane, ncclFloat, left, comm, c->stream));
    CLR_NCCL(g_nccl.Recv(right2, plane, ncclFloat, right, comm, c->stream));
  }
  CLR_NCCL(g_nccl.GroupEnd());
  return 0;
}

// One-plane halo of the DENSITY for the custom maps (clr_beam.cu): the CIC pair of a sample owned by this slab may
// reach plane iz0 + nz, the first plane of the right neighbour (periodic).
int clr_comm_dens_halo(clr_ctx *c, float *dst_plane)
{
  const ClrDev &d = c->dev;
  size_t plane = (size_t)d.pitch * d.n;
  if (c->nranks == 1) {
    CLR_CUDA(cudaMemcpyAsync(dst_plane, c->d_dens, plane * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
  }
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  int right = (c->rank + 1) % c->nranks, left = (c->rank - 1 + c->nranks) % c->nranks;
  CLR_NCCL(g_nccl.GroupStart());
  CLR_NCCL(g_nccl.Send(c->d_dens, plane, ncclFloat, left, comm, c->stream));
  CLR_NCCL(g_nccl.Recv(dst_plane, plane, ncclFloat, right, comm, c->stream));
  CLR_NCCL(g_nccl.GroupEnd());
  return 0;
}
