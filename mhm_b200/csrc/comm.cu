// comm.cu -- the exchanges between the GPUs of one box, below the C ABI.
//
// One process per GPU (the reference's MPI ranks, common/mo_common_read_config.F90:416-437).  The
// library owns an NCCL communicator of its own; the caller only has to carry 128 opaque bytes from
// rank 0 to the other ranks (MPI_Bcast in the Fortran driver, torch.distributed in bench.py).
// libnccl.so.2 is opened with dlopen when the communicator is created, so a single-GPU run does
// not need NCCL at all.
//
// Exchanges built on it:
//   * shared forcing (mhm_cuda_set_meteo_shared): the members of an ensemble / the parameter sets of
//     a calibration sweep run the SAME domain on every GPU, so a forcing chunk is identical on all
//     ranks.  Every rank copies 1/N of the chunk's meteo steps from its host memory and the ranks
//     all-gather the rest over NVLink on the upload stream -- host-DRAM / PCIe traffic per rank
//     drops by N, the NVLink time hides under the previous chunk's kernels (double buffered like
//     mhm_cuda_set_meteo_async).  float32 on the wire is widened on the device after the gather.
//   * cut-link outflow of a sub-catchment-sharded domain (mrm_cuda_exchange_outflow, routing.cu).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "context.h"

namespace mhm {

// the few NCCL entry points in use, declared here so that no NCCL header is needed to build
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclFloat32_ = 7, ncclFloat64_ = 8 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return 0;
  const char* names[] = {getenv("MHM_CUDA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  MHM_REQUIRE(h, "comm: libnccl.so.2 cannot be loaded (%s); set MHM_CUDA_NCCL_LIB", dlerror());
#define MHM_SYM(field, name)                                                     \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                     \
  MHM_REQUIRE(g_nccl.field, "comm: symbol %s missing from the NCCL library", name)
  MHM_SYM(GetUniqueId, "ncclGetUniqueId");
  MHM_SYM(CommInitRank, "ncclCommInitRank");
  MHM_SYM(CommDestroy, "ncclCommDestroy");
  MHM_SYM(AllGather, "ncclAllGather");
  MHM_SYM(Send, "ncclSend");
  MHM_SYM(Recv, "ncclRecv");
  MHM_SYM(GroupStart, "ncclGroupStart");
  MHM_SYM(GroupEnd, "ncclGroupEnd");
  MHM_SYM(GetErrorString, "ncclGetErrorString");
  MHM_SYM(GetVersion, "ncclGetVersion");
#undef MHM_SYM
  g_nccl.handle = h;
  return 0;
}

#define MHM_NCCL_OK(expr)                                                                  \
  do {                                                                                     \
    int r__ = (expr);                                                                      \
    if (r__ != ncclSuccess_) {                                                             \
      ::mhm::set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
      return 4;                                                                            \
    }                                                                                      \
  } while (0)

// rows (meteo steps) of a chunk of n_steps that rank r of N copies from its host: equal slices of
// ceil(n_steps / N) rows (the last ranks' slices may be short or empty); host helper, see
// mhm_cuda_meteo_shared_rows
static void shared_rows(int64_t n_steps, int nranks, int rank, int64_t* per_rank, int64_t* first, int64_t* count) {
  const int64_t rpr = (n_steps + nranks - 1) / nranks;
  int64_t lo = rpr * rank, hi = lo + rpr;
  if (lo > n_steps) lo = n_steps;
  if (hi > n_steps) hi = n_steps;
  *per_rank = rpr;
  *first = lo;
  *count = hi - lo;
}

int comm_send_recv(mhm_cuda_context* ctx, const double* send, const size_t* send_counts, double* recv,
                   const size_t* recv_counts, cudaStream_t st) {
  // one grouped exchange: this rank sends send_counts[r] doubles to rank r (consecutive in `send`,
  // by ascending r) and receives recv_counts[r] from rank r (consecutive in `recv`)
  MHM_REQUIRE(ctx->nccl_comm, "comm: no communicator (mhm_cuda_comm_init)");
  MHM_NCCL_OK(g_nccl.GroupStart());
  size_t so = 0, ro = 0;
  for (int r = 0; r < ctx->nranks; ++r) {
    if (send_counts && send_counts[r]) {
      MHM_NCCL_OK(g_nccl.Send(send + so, send_counts[r], ncclFloat64_, r, (ncclComm_t)ctx->nccl_comm, st));
      so += send_counts[r];
    }
    if (recv_counts && recv_counts[r]) {
      MHM_NCCL_OK(g_nccl.Recv(recv + ro, recv_counts[r], ncclFloat64_, r, (ncclComm_t)ctx->nccl_comm, st));
      ro += recv_counts[r];
    }
  }
  MHM_NCCL_OK(g_nccl.GroupEnd());
  return 0;
}

__global__ void widen_f32_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (double)in[i];
}

}  // namespace mhm

using namespace mhm;

extern "C" {

int mhm_cuda_comm_unique_id(char* id128) {
  MHM_REQUIRE(id128, "comm_unique_id: null buffer");
  if (int rc = nccl_load()) return rc;
  ncclUniqueId id;
  MHM_NCCL_OK(g_nccl.GetUniqueId(&id));
  std::memcpy(id128, id.internal, sizeof(id.internal));
  return 0;
}

int mhm_cuda_comm_init(mhm_cuda_context* ctx, int32_t nranks, int32_t rank, const char* id128) {
  MHM_REQUIRE(ctx && nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad rank %d of %d", rank, nranks);
  MHM_REQUIRE(!ctx->nccl_comm, "comm_init: the context already has a communicator");
  ctx->nranks = nranks;
  ctx->rank = rank;
  if (nranks == 1) return 0;  // nothing to exchange: no NCCL needed
  MHM_REQUIRE(id128, "comm_init: null unique id");
  if (int rc = nccl_load()) return rc;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128, sizeof(id.internal));
  ncclComm_t comm = nullptr;
  MHM_NCCL_OK(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  return 0;
}

int mhm_cuda_comm_finalize(mhm_cuda_context* ctx) {
  MHM_REQUIRE(ctx, "comm_finalize: null context");
  if (ctx->nccl_comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  ctx->nranks = 1;
  ctx->rank = 0;
  return 0;
}

int mhm_cuda_comm_info(mhm_cuda_context* ctx, int32_t* nranks, int32_t* rank, int32_t* nccl_version) {
  MHM_REQUIRE(ctx, "comm_info: null context");
  if (nranks) *nranks = ctx->nranks;
  if (rank) *rank = ctx->rank;
  if (nccl_version) {
    int v = 0;
    if (g_nccl.handle) g_nccl.GetVersion(&v);
    *nccl_version = v;
  }
  return 0;
}

int mhm_cuda_meteo_shared_rows(int64_t n_steps, int32_t nranks, int32_t rank, int64_t* first_row,
                               int64_t* n_rows) {
  MHM_REQUIRE(n_steps >= 1 && nranks >= 1 && rank >= 0 && rank < nranks && first_row && n_rows,
              "meteo_shared_rows: bad arguments");
  int64_t rpr;
  shared_rows(n_steps, nranks, rank, &rpr, first_row, n_rows);
  return 0;
}

int mhm_cuda_set_meteo_shared(mhm_cuda_context* ctx, int32_t iDomain, int32_t var, const void* base,
                              int32_t is_f32, int64_t ld, int64_t offset, int64_t first_step,
                              int64_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(var >= 0 && var < MHM_M_COUNT, "set_meteo_shared: bad variable %d", var);
  MHM_REQUIRE(base && ld >= d->cfg.nCells && offset >= 0 && first_step >= 1 && n_steps >= 1,
              "set_meteo_shared: bad base/ld/offset/steps");
  MHM_REQUIRE(ctx->nranks == 1 || ctx->nccl_comm, "set_meteo_shared: no communicator (mhm_cuda_comm_init)");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const int N = ctx->nranks;
  const size_t n = (size_t)d->cfg.nCells;
  int64_t rpr, lo, cnt;
  shared_rows(n_steps, N, ctx->rank, &rpr, &lo, &cnt);
  // the gather needs equal slices: the buffer holds N * rpr rows, rows >= n_steps are never read
  const size_t need = (size_t)N * (size_t)rpr * n;
  cudaStream_t cs = ctx->copy_stream;
  const int nb = d->met_owned[var] ? 1 - d->met_active[var] : 0;
  if (d->met_bufcap[var][nb] < need) {
    MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    MHM_CUDA_OK(cudaStreamSynchronize(cs));
    cudaFree(d->met_buf[var][nb]);
    d->met_buf[var][nb] = nullptr;
    d->met_bufcap[var][nb] = 0;
    MHM_CUDA_OK(cudaMalloc(&d->met_buf[var][nb], need * sizeof(double)));
    d->met_bufcap[var][nb] = need;
  }
  if (is_f32 && d->met_f32cap < need) {  // staging for float32 on the wire (ordered on the copy stream)
    MHM_CUDA_OK(cudaStreamSynchronize(cs));
    cudaFree(d->met_f32);
    d->met_f32 = nullptr;
    d->met_f32cap = 0;
    MHM_CUDA_OK(cudaMalloc(&d->met_f32, need * sizeof(float)));
    d->met_f32cap = need;
  }
  if (!d->met_ready[var]) MHM_CUDA_OK(cudaEventCreateWithFlags(&d->met_ready[var], cudaEventDisableTiming));
  if (d->met_free_set[var][nb])  // the last run that read this buffer must be done
    MHM_CUDA_OK(cudaStreamWaitEvent(cs, d->met_free[var][nb], 0));
  const size_t esz = is_f32 ? sizeof(float) : sizeof(double);
  char* dst = is_f32 ? (char*)d->met_f32 : (char*)d->met_buf[var][nb];
  if (cnt > 0)
    MHM_CUDA_OK(cudaMemcpy2DAsync(dst + (size_t)lo * n * esz, n * esz,
                                  (const char*)base + ((size_t)lo * (size_t)ld + (size_t)offset) * esz,
                                  (size_t)ld * esz, n * esz, (size_t)cnt, cudaMemcpyHostToDevice, cs));
  if (N > 1)  // in place: every rank's slice sits at its own position of the receive buffer
    MHM_NCCL_OK(g_nccl.AllGather(dst + (size_t)ctx->rank * (size_t)rpr * n * esz, dst, (size_t)rpr * n,
                                 is_f32 ? ncclFloat32_ : ncclFloat64_, (ncclComm_t)ctx->nccl_comm, cs));
  if (is_f32) {
    const size_t tot = (size_t)n_steps * n;
    widen_f32_kernel<<<(unsigned)std::min<size_t>((tot + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, cs>>>(
        d->met_f32, d->met_buf[var][nb], tot);
    MHM_CUDA_OK(cudaGetLastError());
  }
  MHM_CUDA_OK(cudaEventRecord(d->met_ready[var], cs));
  d->met_ready_set[var] = true;
  d->met_active[var] = nb;
  d->met_owned[var] = true;
  d->met[var] = d->met_buf[var][nb];
  d->met_first[var] = first_step;
  d->met_n[var] = n_steps;
  d->met_h2d_bytes += (size_t)cnt * n * esz;
  return 0;
}

int mhm_cuda_meteo_h2d_bytes(mhm_cuda_context* ctx, int32_t iDomain, int64_t* bytes) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(bytes, "meteo_h2d_bytes: null output");
  *bytes = (int64_t)d->met_h2d_bytes;
  return 0;
}

}  // extern "C"
