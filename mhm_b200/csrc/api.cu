// api.cu -- C ABI (include/mhm_cuda.h): lifecycle, parameter/state/flux/meteo transfer,
// calendar, and the drivers of the fused cell kernel (per-step seam and time blocks).
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>

#include "context.h"
#include "fastmath.cuh"

namespace mhm {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

Domain* find_domain(mhm_cuda_context* ctx, int32_t iDomain) {
  if (!ctx) {
    set_error("null context");
    return nullptr;
  }
  auto it = ctx->domains.find(iDomain);
  if (it == ctx->domains.end()) {
    set_error("domain %d is not registered", iDomain);
    return nullptr;
  }
  return it->second;
}

static size_t state_rows(const Domain* d, int id) {
  return id == MHM_S_SOILMOIST ? (size_t)d->cfg.nHorizons : 1;
}
static size_t flux_rows(const Domain* d, int id) {
  return (id == MHM_F_AETSOIL || id == MHM_F_INFILSOIL) ? (size_t)d->cfg.nHorizons : 1;
}

// copy a Fortran section (rows x nCells, row stride ld) host -> dense device rows
static int h2d_rows(mhm_cuda_context* ctx, double* dst, const double* base, int64_t ld,
                    int64_t offset, size_t n, size_t rows) {
  MHM_CUDA_OK(cudaMemcpy2DAsync(dst, n * sizeof(double), base + offset, (size_t)ld * sizeof(double),
                                n * sizeof(double), rows, cudaMemcpyHostToDevice, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
static int d2h_rows(mhm_cuda_context* ctx, double* base, int64_t ld, int64_t offset,
                    const double* src, size_t n, size_t rows) {
  MHM_CUDA_OK(cudaMemcpy2DAsync(base + offset, (size_t)ld * sizeof(double), src, n * sizeof(double),
                                n * sizeof(double), rows, cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static void free_domain(Domain* d) {
  for (auto& p : d->P) cudaFree(p);
  for (auto& p : d->S) cudaFree(p);
  for (auto& p : d->F) cudaFree(p);
  for (int v = 0; v < MHM_M_COUNT; ++v) {
    for (int b = 0; b < 2; ++b) {
      cudaFree(d->met_buf[v][b]);
      if (d->met_free[v][b]) cudaEventDestroy(d->met_free[v][b]);
    }
    if (d->met_ready[v]) cudaEventDestroy(d->met_ready[v]);
  }
  for (auto& p : d->weights) cudaFree(p);
  cudaFree(d->met_f32);
  cudaFree(d->runoff_hist);
  cudaFree(d->out_acc);
  cudaFree(d->out_win);
  for (int w = 0; w < 3; ++w) cudaFree(d->opt_data[w]);
  cudaFree(d->bfi_acc);
  if (d->rt) routing_free(d->rt);
  if (d->mpr) mpr_free(d->mpr);
  delete d;
}

}  // namespace mhm

using namespace mhm;

void mhm_cuda_context::stat_begin(int which) {
  if (!timing) return;
  TimedLaunch t;
  t.which = which;
  for (cudaEvent_t* e : {&t.a, &t.b}) {
    if (!ev_pool.empty()) {
      *e = ev_pool.back();
      ev_pool.pop_back();
    } else {
      cudaEventCreate(e);
    }
  }
  cudaEventRecord(t.a, stream);
  pending.push_back(t);
}
void mhm_cuda_context::stat_end(int which, int64_t launches) {
  stat_launches[which] += launches;
  if (!timing) return;
  cudaEventRecord(pending.back().b, stream);
}
int mhm_cuda_context::stat_flush() {
  if (pending.empty()) return 0;
  MHM_CUDA_OK(cudaStreamSynchronize(stream));
  for (auto& t : pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.a, t.b);
    stat_ms[t.which] += ms;
    ev_pool.push_back(t.a);
    ev_pool.push_back(t.b);
  }
  pending.clear();
  return 0;
}

extern "C" {

const char* mhm_cuda_last_error(void) { return g_err.c_str(); }

// host instantiations of the fast kernel's math (tests/test_fastmath.py measures their error)
double mhm_host_fast_log(double x) { return mhm::fm::log_pos(x); }
double mhm_host_fast_exp(double x) { return mhm::fm::exp_bounded(x); }
double mhm_host_fast_pow(double x, double y) { return mhm::fm::pow_pos(x, y); }
double mhm_host_fast_pow23(double x) { return mhm::fm::pow23_pos(x); }
double mhm_host_tab_log(double x) { return mhm::fm::log_tab(mhm::fm::h_tables, x); }
double mhm_host_tab_exp(double x) { return mhm::fm::exp_tab(mhm::fm::h_tables, x); }
double mhm_host_tab_pow(double x, double y) { return mhm::fm::pow_tab(mhm::fm::h_tables, x, y); }
const char* mhm_cuda_version(void) { return "mhm_cuda 0.1 (sm_100a)"; }

int mhm_cuda_init(int device, mhm_cuda_context** out) {
  MHM_REQUIRE(out, "mhm_cuda_init: null output handle");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("mhm_cuda_init: no CUDA device available (%s); there is no CPU fallback",
              cudaGetErrorString(e));
    return 3;
  }
  if (device < 0) MHM_CUDA_OK(cudaGetDevice(&device));
  MHM_REQUIRE(device < count, "mhm_cuda_init: device %d out of range (%d devices)", device, count);
  MHM_CUDA_OK(cudaSetDevice(device));
  auto* ctx = new mhm_cuda_context();
  ctx->device = device;
  MHM_CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  MHM_CUDA_OK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (auto& ev : ctx->ev) MHM_CUDA_OK(cudaEventCreate(&ev));
  // history buffers of a time block may take 45 % of the device's memory (long blocks amortise
  // the routing pipeline's fill/drain and the cell kernel's state/parameter loads)
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0)
    ctx->block_bytes = (size_t)((double)total_b * 0.45);
  if (const char* s = getenv("MHM_CUDA_BLOCK_BYTES")) ctx->block_bytes = (size_t)atoll(s);
  MHM_CUDA_OK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
  if (const char* s = getenv("MHM_CUDA_LAUNCH_LOG")) ctx->launch_log = fopen(s, "w");
  *out = ctx;
  return 0;
}

int mhm_cuda_finalize(mhm_cuda_context* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->domains) free_domain(kv.second);
  mhm_cuda_comm_finalize(ctx);
  if (ctx->launch_log) fclose(ctx->launch_log);
  for (auto& ev : ctx->ev) cudaEventDestroy(ev);
  for (auto& t : ctx->pending) {
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return 0;
}

int mhm_cuda_register_domain(mhm_cuda_context* ctx, int32_t iDomain, const mhm_domain_config* cfg) {
  MHM_REQUIRE(ctx && cfg, "register_domain: null argument");
  MHM_REQUIRE(ctx->domains.find(iDomain) == ctx->domains.end(), "domain %d already registered",
              iDomain);
  MHM_REQUIRE(cfg->nCells > 0, "register_domain: nCells = %d", cfg->nCells);
  MHM_REQUIRE(cfg->nHorizons >= 1 && cfg->nHorizons <= kMaxHorizons,
              "register_domain: nHorizons = %d outside 1..%d", cfg->nHorizons, kMaxHorizons);
  MHM_REQUIRE(cfg->nLAI >= 1 && cfg->nLCscenes >= 1 && cfg->nMembers >= 1,
              "register_domain: nLAI/nLCscenes/nMembers must be >= 1");
  MHM_REQUIRE(cfg->processMatrix && cfg->nProcesses >= 8,
              "register_domain: processMatrix with >= 8 rows required");
  MHM_REQUIRE(cfg->nLCscenes <= 32767 && cfg->nLAI <= 32767, "register_domain: nLCscenes / nLAI above 32767");
  // processMatrix(3,1), (5,1), (8,1)
  const int soil_case = cfg->processMatrix[2], pet_case = cfg->processMatrix[4];
  MHM_REQUIRE(soil_case >= 1 && soil_case <= 4, "soil moisture processCase %d unsupported", soil_case);
  MHM_REQUIRE(pet_case >= -1 && pet_case <= 3, "PET processCase %d unsupported", pet_case);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  // owned until registration succeeded: an early return frees the domain and its device memory
  std::unique_ptr<Domain, void (*)(Domain*)> own(new Domain(), free_domain);
  Domain* d = own.get();
  d->id = iDomain;
  d->cfg = *cfg;
  d->processMatrix.assign(cfg->processMatrix, cfg->processMatrix + (size_t)cfg->nProcesses * 3);
  d->cfg.processMatrix = d->processMatrix.data();
  d->soil_case = soil_case;
  d->pet_case = pet_case;
  d->rout_case = d->processMatrix[7];
  const size_t n = (size_t)cfg->nCells, M = (size_t)cfg->nMembers;
  for (int s = 0; s < MHM_S_COUNT; ++s) {
    const size_t sz = M * state_rows(d, s) * n * sizeof(double);
    MHM_CUDA_OK(cudaMalloc(&d->S[s], sz));
    MHM_CUDA_OK(cudaMemsetAsync(d->S[s], 0, sz, ctx->stream));
  }
  for (int f = 0; f < MHM_F_COUNT; ++f) {
    const size_t sz = M * flux_rows(d, f) * n * sizeof(double);
    MHM_CUDA_OK(cudaMalloc(&d->F[f], sz));
    MHM_CUDA_OK(cudaMemsetAsync(d->F[f], 0, sz, ctx->stream));
  }
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->domains[iDomain] = own.release();
  return 0;
}

int mhm_cuda_unregister_domain(mhm_cuda_context* ctx, int32_t iDomain) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  free_domain(d);
  ctx->domains.erase(iDomain);
  return 0;
}

int mhm_cuda_set_param(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t id,
                       const double* base, int64_t ld, int64_t offset, int32_t dim2, int32_t dim3) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(id >= 0 && id < MHM_P_COUNT, "set_param: bad param id %d", id);
  MHM_REQUIRE(member >= 0 && member < d->cfg.nMembers, "set_param: bad member %d", member);
  MHM_REQUIRE(base && ld >= d->cfg.nCells && offset >= 0, "set_param: bad base/ld/offset");
  // expected shape per the reference's allocation (mo_mpr_global_variables.f90:128-170)
  const int nH = d->cfg.nHorizons, nLAI = d->cfg.nLAI, nLC = d->cfg.nLCscenes;
  int e2 = 1, e3 = 1;
  switch (id) {
    case MHM_P_FSEALED: case MHM_P_ALPHA: case MHM_P_DEGDAYINC: case MHM_P_DEGDAYMAX:
    case MHM_P_DEGDAYNOPRE: case MHM_P_KFASTFLOW: case MHM_P_KSLOWFLOW: case MHM_P_KBASEFLOW:
    case MHM_P_KPERCO: case MHM_P_TEMPTHRESH: e3 = nLC; break;
    case MHM_P_FROOTS: case MHM_P_SOILMOISTFC: case MHM_P_SOILMOISTSAT: case MHM_P_SOILMOISTEXP:
    case MHM_P_WILTINGPOINT: e2 = nH; e3 = nLC; break;
    case MHM_P_MAXINTER: case MHM_P_PRIETAYALPHA: case MHM_P_SURFRESIST: e2 = nLAI; break;
    case MHM_P_PETLAICORFACTOR: case MHM_P_AERORESIST: e2 = nLAI; e3 = nLC; break;
    default: break;
  }
  MHM_REQUIRE(dim2 == e2 && dim3 == e3, "set_param(%d): shape (:,%d,%d) given, (:,%d,%d) expected",
              id, dim2, dim3, e2, e3);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells, rows = (size_t)dim2 * dim3;
  if (!d->P[id]) {
    MHM_CUDA_OK(cudaMalloc(&d->P[id], (size_t)d->cfg.nMembers * rows * n * sizeof(double)));
    d->P_dim2[id] = dim2;
    d->P_dim3[id] = dim3;
  }
  return h2d_rows(ctx, d->P[id] + (size_t)member * rows * n, base, ld, offset, n, rows);
}

int mhm_cuda_set_state(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t id,
                       const double* base, int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(id >= 0 && id < MHM_S_COUNT, "set_state: bad state id %d", id);
  MHM_REQUIRE(member >= 0 && member < d->cfg.nMembers, "set_state: bad member %d", member);
  MHM_REQUIRE(base && ld >= d->cfg.nCells && offset >= 0, "set_state: bad base/ld/offset");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells, rows = state_rows(d, id);
  return h2d_rows(ctx, d->S[id] + (size_t)member * rows * n, base, ld, offset, n, rows);
}

int mhm_cuda_get_state(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t id,
                       double* base, int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(id >= 0 && id < MHM_S_COUNT, "get_state: bad state id %d", id);
  MHM_REQUIRE(member >= 0 && member < d->cfg.nMembers, "get_state: bad member %d", member);
  MHM_REQUIRE(base && ld >= d->cfg.nCells && offset >= 0, "get_state: bad base/ld/offset");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells, rows = state_rows(d, id);
  return d2h_rows(ctx, base, ld, offset, d->S[id] + (size_t)member * rows * n, n, rows);
}

int mhm_cuda_get_flux(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t id,
                      double* base, int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(id >= 0 && id < MHM_F_COUNT, "get_flux: bad flux id %d", id);
  MHM_REQUIRE(member >= 0 && member < d->cfg.nMembers, "get_flux: bad member %d", member);
  MHM_REQUIRE(base && ld >= d->cfg.nCells && offset >= 0, "get_flux: bad base/ld/offset");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells, rows = flux_rows(d, id);
  return d2h_rows(ctx, base, ld, offset, d->F[id] + (size_t)member * rows * n, n, rows);
}

__global__ void fill_kernel(double* p, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// mHM/mo_init_states.f90:280-300; constants MPR/mo_mpr_constants.f90:26-30
int mhm_cuda_states_default_init(mhm_cuda_context* ctx, int32_t iDomain, const double* depth) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const int nH = d->cfg.nHorizons;
  MHM_REQUIRE(nH == 1 || depth, "states_default_init: HorizonDepth_mHM required");
  const size_t n = (size_t)d->cfg.nCells, M = (size_t)d->cfg.nMembers;
  auto fill = [&](double* p, size_t cnt, double v) {
    fill_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(p, cnt, v);
  };
  const double P1 = 0.0, P2 = 15.0, P3 = 10.0, P4 = 75.0, P5 = 1500.0, C1 = 0.25;
  fill(d->S[MHM_S_INTER], M * n, P1);
  fill(d->S[MHM_S_SNOWPACK], M * n, P2);
  fill(d->S[MHM_S_SEALSTW], M * n, P1);
  fill(d->S[MHM_S_UNSATSTW], M * n, P3);
  fill(d->S[MHM_S_SATSTW], M * n, P4);
  for (size_t m = 0; m < M; ++m) {
    for (int i = 0; i < nH; ++i) {
      double v;
      if (i == nH - 1) {
        v = (P5 - (nH > 1 ? depth[nH - 2] : 0.0)) * C1;
      } else if (i == 0) {
        v = depth[0] * C1;
      } else {
        v = (depth[i] - depth[i - 1]) * C1;
      }
      fill(d->S[MHM_S_SOILMOIST] + (m * nH + i) * n, n, v);
    }
  }
  for (int f = 0; f < MHM_F_COUNT; ++f)
    MHM_CUDA_OK(cudaMemsetAsync(d->F[f], 0, M * flux_rows(d, f) * n * sizeof(double), ctx->stream));
  MHM_CUDA_OK(cudaGetLastError());
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------------- meteo
int mhm_cuda_set_meteo_config(mhm_cuda_context* ctx, int32_t iDomain, const mhm_meteo_config* cfg) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(cfg, "set_meteo_config: null config");
  MHM_REQUIRE(cfg->pet_case == d->pet_case,
              "set_meteo_config: pet_case %d differs from processMatrix(5,1) = %d", cfg->pet_case,
              d->pet_case);
  MHM_REQUIRE(cfg->nTstepForcingDay >= 1 && cfg->nTstepForcingDay <= 24,
              "set_meteo_config: nTstepForcingDay = %d", cfg->nTstepForcingDay);
  // the reference rejects PET cases 1-3 with hourly forcing (mo_meteo_handler.f90:879-909)
  MHM_REQUIRE(!(cfg->is_hourly_forcing && cfg->pet_case > 0),
              "set_meteo_config: PET processCase %d needs daily forcing", cfg->pet_case);
  d->mcfg = *cfg;
  d->has_meteo_cfg = true;
  d->h_idx.clear();  // calendar depends on nTstepForcingDay
  return 0;
}

// upload into the buffer the in-flight block does not read; returns without waiting for the copy
static int meteo_upload(mhm_cuda_context* ctx, Domain* d, int var, const double* base, int64_t ld,
                        int64_t offset, int64_t first_step, int64_t n_steps) {
  const size_t n = (size_t)d->cfg.nCells, need = (size_t)n_steps * n;
  const int nb = d->met_owned[var] ? 1 - d->met_active[var] : 0;
  if (d->met_bufcap[var][nb] < need) {
    // (re)allocation: make sure nobody still reads the old allocation
    MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    MHM_CUDA_OK(cudaStreamSynchronize(ctx->copy_stream));
    cudaFree(d->met_buf[var][nb]);
    d->met_buf[var][nb] = nullptr;
    d->met_bufcap[var][nb] = 0;
    MHM_CUDA_OK(cudaMalloc(&d->met_buf[var][nb], need * sizeof(double)));
    d->met_bufcap[var][nb] = need;
  }
  if (!d->met_ready[var]) MHM_CUDA_OK(cudaEventCreateWithFlags(&d->met_ready[var], cudaEventDisableTiming));
  if (d->met_free_set[var][nb])  // the last run that read this buffer must be done
    MHM_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, d->met_free[var][nb], 0));
  MHM_CUDA_OK(cudaMemcpy2DAsync(d->met_buf[var][nb], n * sizeof(double), base + offset,
                                (size_t)ld * sizeof(double), n * sizeof(double), (size_t)n_steps,
                                cudaMemcpyHostToDevice, ctx->copy_stream));
  MHM_CUDA_OK(cudaEventRecord(d->met_ready[var], ctx->copy_stream));
  d->met_ready_set[var] = true;
  d->met_active[var] = nb;
  d->met_owned[var] = true;
  d->met[var] = d->met_buf[var][nb];
  d->met_first[var] = first_step;
  d->met_n[var] = n_steps;
  d->met_h2d_bytes += need * sizeof(double);
  return 0;
}

int mhm_cuda_set_meteo_async(mhm_cuda_context* ctx, int32_t iDomain, int32_t var, const double* base,
                             int64_t ld, int64_t offset, int64_t first_step, int64_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(var >= 0 && var < MHM_M_COUNT, "set_meteo: bad variable %d", var);
  MHM_REQUIRE(base && ld >= d->cfg.nCells && offset >= 0 && first_step >= 1 && n_steps >= 1,
              "set_meteo: bad base/ld/offset/steps");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  return meteo_upload(ctx, d, var, base, ld, offset, first_step, n_steps);
}

int mhm_cuda_set_meteo(mhm_cuda_context* ctx, int32_t iDomain, int32_t var, const double* base,
                       int64_t ld, int64_t offset, int64_t first_step, int64_t n_steps) {
  if (int rc = mhm_cuda_set_meteo_async(ctx, iDomain, var, base, ld, offset, first_step, n_steps)) return rc;
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->copy_stream));  // the caller may reuse `base` at once
  return 0;
}

}  // extern "C"

// ---- N3: level-2 chunk -> packed L1 forcing on the device ---------------------------------
// one thread per (L1 cell, step); the valid level-2 cells of an L1 cell are summed with the
// second index outermost, like the reference's loops (mo_meteo_spatial_tools.f90:172-196)
template <class T>
__global__ void meteo_l2_to_l1_kernel(const T* __restrict__ data2, int nr2, int nc2, const int32_t* __restrict__ mask2,
                                      const int32_t* __restrict__ ci1, const int32_t* __restrict__ cj1, int nCells,
                                      int mode, int f, double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nCells) return;
  const size_t t = blockIdx.y;
  const T* d2 = data2 + t * (size_t)nr2 * nc2;
  const int i = ci1[k], j = cj1[k];  // 1-based L1 coordinates
  double v;
  if (mode == 0) {
    v = (double)d2[(size_t)(j - 1) * nr2 + (i - 1)];
  } else if (mode == 1) {  // aggregation, f = cellsize1 / cellsize2
    double s = 0.0;
    int cnt = 0;
    const int i0 = (i - 1) * f + 1, i1 = min(i * f, nr2), j0 = (j - 1) * f + 1, j1 = min(j * f, nc2);
    for (int jj = j0; jj <= j1; ++jj)
      for (int ii = i0; ii <= i1; ++ii) {
        if (!mask2[(size_t)(jj - 1) * nr2 + (ii - 1)]) continue;
        s = s + (double)d2[(size_t)(jj - 1) * nr2 + (ii - 1)];
        ++cnt;
      }
    v = s / (double)cnt;
  } else {  // disaggregation, f = cellsize2 / cellsize1
    const int ic = (i + f - 1) / f, jc = (j + f - 1) / f;
    v = mask2[(size_t)(jc - 1) * nr2 + (ic - 1)] ? (double)d2[(size_t)(jc - 1) * nr2 + (ic - 1)] : -9999.0;
  }
  out[t * (size_t)nCells + k] = v;
}

extern "C" {

int mhm_cuda_set_meteo_l2(mhm_cuda_context* ctx, int32_t iDomain, int32_t var, const void* data2,
                          int32_t is_f32, int32_t nrows2, int32_t ncols2, const int32_t* mask2,
                          double cellsize2, int32_t nrows1, int32_t ncols1, const int32_t* mask1,
                          double cellsize1, int64_t first_step, int64_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(var >= 0 && var < MHM_M_COUNT && data2 && mask2 && mask1 && nrows2 >= 1 && ncols2 >= 1 &&
                  nrows1 >= 1 && ncols1 >= 1 && cellsize1 > 0 && cellsize2 > 0 && first_step >= 1 && n_steps >= 1,
              "set_meteo_l2: bad arguments");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  int mode = 0, f = 1;
  const double r = cellsize1 / cellsize2;
  if (r > 1.0) {
    mode = 1;
    f = (int)std::lround(r);
    MHM_REQUIRE(fabs(r - f) < 1e-9, "set_meteo_l2: resolutions %g / %g are not multiples", cellsize1, cellsize2);
  } else if (r < 1.0) {
    mode = 2;
    f = (int)std::lround(1.0 / r);
    MHM_REQUIRE(fabs(1.0 / r - f) < 1e-9, "set_meteo_l2: resolutions %g / %g are not multiples", cellsize1, cellsize2);
  } else {
    MHM_REQUIRE(nrows1 == nrows2 && ncols1 == ncols2, "set_meteo_l2: equal resolutions need equal grids");
  }
  // coordinates of the packed L1 cells (element order of the mask)
  std::vector<int32_t> ci, cj;
  for (int j = 1; j <= ncols1; ++j)
    for (int i = 1; i <= nrows1; ++i)
      if (mask1[(size_t)(j - 1) * nrows1 + (i - 1)]) {
        ci.push_back(i);
        cj.push_back(j);
      }
  MHM_REQUIRE((int)ci.size() == d->cfg.nCells, "set_meteo_l2: mask1 holds %d cells, the domain %d", (int)ci.size(),
              d->cfg.nCells);
  const size_t n = (size_t)d->cfg.nCells, need = (size_t)n_steps * n;
  const size_t n2 = (size_t)nrows2 * ncols2, raw_bytes = n2 * (size_t)n_steps * (is_f32 ? 4 : 8);
  void* raw = nullptr;
  int32_t *dm2 = nullptr, *dci = nullptr, *dcj = nullptr;
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
  MHM_CUDA_OK(cudaMallocAsync(&raw, raw_bytes, cs));
  MHM_CUDA_OK(cudaMallocAsync((void**)&dm2, n2 * sizeof(int32_t), cs));
  MHM_CUDA_OK(cudaMallocAsync((void**)&dci, n * sizeof(int32_t), cs));
  MHM_CUDA_OK(cudaMallocAsync((void**)&dcj, n * sizeof(int32_t), cs));
  if (!d->met_ready[var]) MHM_CUDA_OK(cudaEventCreateWithFlags(&d->met_ready[var], cudaEventDisableTiming));
  if (d->met_free_set[var][nb]) MHM_CUDA_OK(cudaStreamWaitEvent(cs, d->met_free[var][nb], 0));
  MHM_CUDA_OK(cudaMemcpyAsync(raw, data2, raw_bytes, cudaMemcpyHostToDevice, cs));
  MHM_CUDA_OK(cudaMemcpyAsync(dm2, mask2, n2 * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  MHM_CUDA_OK(cudaMemcpyAsync(dci, ci.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  MHM_CUDA_OK(cudaMemcpyAsync(dcj, cj.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  const dim3 grid((unsigned)((n + 127) / 128), (unsigned)n_steps);
  if (is_f32)
    meteo_l2_to_l1_kernel<float><<<grid, 128, 0, cs>>>((const float*)raw, nrows2, ncols2, dm2, dci, dcj, (int)n, mode, f,
                                                       d->met_buf[var][nb]);
  else
    meteo_l2_to_l1_kernel<double><<<grid, 128, 0, cs>>>((const double*)raw, nrows2, ncols2, dm2, dci, dcj, (int)n, mode,
                                                        f, d->met_buf[var][nb]);
  MHM_CUDA_OK(cudaGetLastError());
  MHM_CUDA_OK(cudaFreeAsync(raw, cs));
  MHM_CUDA_OK(cudaFreeAsync(dm2, cs));
  MHM_CUDA_OK(cudaFreeAsync(dci, cs));
  MHM_CUDA_OK(cudaFreeAsync(dcj, cs));
  MHM_CUDA_OK(cudaEventRecord(d->met_ready[var], cs));
  MHM_CUDA_OK(cudaStreamSynchronize(cs));  // ci / cj / the caller's buffers may go away
  d->met_ready_set[var] = true;
  d->met_active[var] = nb;
  d->met_owned[var] = true;
  d->met[var] = d->met_buf[var][nb];
  d->met_first[var] = first_step;
  d->met_n[var] = n_steps;
  return 0;
}

int mhm_cuda_get_meteo(mhm_cuda_context* ctx, int32_t iDomain, int32_t var, double* out, int64_t ld,
                       int64_t first_step, int64_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(var >= 0 && var < MHM_M_COUNT && out && ld >= d->cfg.nCells && d->met[var] &&
                  first_step >= d->met_first[var] && first_step + n_steps <= d->met_first[var] + d->met_n[var],
              "get_meteo: steps outside the resident chunk");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  if (d->met_owned[var] && d->met_ready_set[var]) MHM_CUDA_OK(cudaStreamWaitEvent(ctx->stream, d->met_ready[var], 0));
  const size_t n = (size_t)d->cfg.nCells;
  MHM_CUDA_OK(cudaMemcpy2DAsync(out, (size_t)ld * sizeof(double),
                                d->met[var] + (size_t)(first_step - d->met_first[var]) * n, n * sizeof(double),
                                n * sizeof(double), (size_t)n_steps, cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mhm_cuda_set_meteo_device(mhm_cuda_context* ctx, int32_t iDomain, int32_t var,
                              const double* dev, int64_t first_step, int64_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(var >= 0 && var < MHM_M_COUNT, "set_meteo_device: bad variable %d", var);
  MHM_REQUIRE(dev && first_step >= 1 && n_steps >= 1, "set_meteo_device: bad arguments");
  d->met_owned[var] = false;
  d->met_ready_set[var] = false;
  d->met[var] = const_cast<double*>(dev);
  d->met_first[var] = first_step;
  d->met_n[var] = n_steps;
  return 0;
}

int mhm_cuda_set_meteo_weights(mhm_cuda_context* ctx, int32_t iDomain, int32_t var,
                               const double* base, int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(var == MHM_M_PRE || var == MHM_M_TEMP || var == MHM_M_PET,
              "set_meteo_weights: variable %d has no weights", var);
  MHM_REQUIRE(base && ld >= d->cfg.nCells && offset >= 0, "set_meteo_weights: bad base/ld/offset");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells;
  if (!d->weights[var]) MHM_CUDA_OK(cudaMalloc(&d->weights[var], 24 * 12 * n * sizeof(double)));
  return h2d_rows(ctx, d->weights[var], base, ld, offset, n, 24 * 12);
}

// ---------------------------------------------------------------------------------- time
int mhm_cuda_set_time(mhm_cuda_context* ctx, int32_t iDomain, const mhm_time_config* cfg) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(cfg && cfg->nTimeSteps >= 1, "set_time: bad config");
  d->axis.jul_start = cfg->jul_start;
  d->axis.nTimeSteps = cfg->nTimeSteps;
  d->axis.warming_days = cfg->warming_days;
  d->axis.timeStep_LAI_input = cfg->timeStep_LAI_input;
  d->axis.lc_year_start = cfg->lc_year_start;
  d->axis.LCyearId.clear();
  if (cfg->LCyearId && cfg->lc_nyears > 0)
    d->axis.LCyearId.assign(cfg->LCyearId, cfg->LCyearId + cfg->lc_nyears);
  d->has_time = true;
  d->h_idx.clear();
  return 0;
}

int mhm_time_indices(const mhm_time_config* cfg, int32_t timestep_h, int32_t nTstepForcingDay,
                     int32_t tt_first, int32_t n_steps, mhm_step_index* out) {
  MHM_REQUIRE(cfg && out && tt_first >= 1 && n_steps >= 0 && timestep_h >= 1 &&
                  nTstepForcingDay >= 1,
              "mhm_time_indices: bad arguments");
  TimeAxis ax;
  ax.jul_start = cfg->jul_start;
  ax.nTimeSteps = cfg->nTimeSteps;
  ax.warming_days = cfg->warming_days;
  ax.timeStep_LAI_input = cfg->timeStep_LAI_input;
  ax.lc_year_start = cfg->lc_year_start;
  if (cfg->LCyearId && cfg->lc_nyears > 0)
    ax.LCyearId.assign(cfg->LCyearId, cfg->LCyearId + cfg->lc_nyears);
  std::vector<StepIdx> v;
  fill_step_indices(ax, timestep_h, nTstepForcingDay, tt_first + n_steps - 1, v);
  for (int i = 0; i < n_steps; ++i) {
    const StepIdx& s = v[(size_t)tt_first - 1 + i];
    mhm_step_index& o = out[i];
    memset(&o, 0, sizeof(o));
    o.iMeteoTS = s.iMeteoTS;
    o.year = s.year;
    o.yId = s.yId;
    o.iLAI = s.iLAI;
    o.doy = s.doy;
    o.month = s.month;
    o.hour = s.hour;
    o.isday = s.isday;
  }
  return 0;
}

// ------------------------------------------------------------------------- cell kernel
static int ensure_calendar(mhm_cuda_context* ctx, Domain* d) {
  if (!d->h_idx.empty()) return 0;
  MHM_REQUIRE(d->has_time, "domain %d: mhm_cuda_set_time has not been called", d->id);
  MHM_REQUIRE(d->has_meteo_cfg, "domain %d: mhm_cuda_set_meteo_config has not been called", d->id);
  fill_step_indices(d->axis, d->cfg.timestep_h, d->mcfg.nTstepForcingDay, d->axis.nTimeSteps,
                    d->h_idx);
  return 0;
}

static int check_inputs(Domain* d, const StepIdx* idx, int n_steps) {
  MHM_REQUIRE(d->has_meteo_cfg, "domain %d: meteo config missing", d->id);
  static const int always[] = {MHM_P_FSEALED,     MHM_P_ALPHA,        MHM_P_DEGDAYINC,
                               MHM_P_DEGDAYMAX,   MHM_P_DEGDAYNOPRE,  MHM_P_FROOTS,
                               MHM_P_MAXINTER,    MHM_P_KARSTLOSS,    MHM_P_KFASTFLOW,
                               MHM_P_KSLOWFLOW,   MHM_P_KBASEFLOW,    MHM_P_KPERCO,
                               MHM_P_SOILMOISTFC, MHM_P_SOILMOISTSAT, MHM_P_SOILMOISTEXP,
                               MHM_P_JARVIS_C1,   MHM_P_TEMPTHRESH,   MHM_P_UNSATTHRESH,
                               MHM_P_SEALEDTHRESH, MHM_P_WILTINGPOINT};
  for (int id : always) MHM_REQUIRE(d->P[id], "domain %d: parameter %d has not been set", d->id, id);
  std::vector<int> needP, needM = {MHM_M_PRE, MHM_M_TEMP};
  switch (d->pet_case) {
    case -1: needP = {MHM_P_PETLAICORFACTOR}; needM.push_back(MHM_M_PET); break;
    case 0: needP = {MHM_P_FASP}; needM.push_back(MHM_M_PET); break;
    case 1:
      needP = {MHM_P_FASP, MHM_P_HARSAMCOEFF, MHM_P_LATITUDE};
      needM.push_back(MHM_M_TMIN);
      needM.push_back(MHM_M_TMAX);
      break;
    case 2: needP = {MHM_P_PRIETAYALPHA}; needM.push_back(MHM_M_NETRAD); break;
    case 3:
      needP = {MHM_P_AERORESIST, MHM_P_SURFRESIST};
      needM.push_back(MHM_M_NETRAD);
      needM.push_back(MHM_M_ABSVAPPRESS);
      needM.push_back(MHM_M_WINDSPEED);
      break;
  }
  for (int id : needP) MHM_REQUIRE(d->P[id], "domain %d: parameter %d has not been set", d->id, id);
  int64_t lo = idx[0].iMeteoTS, hi = idx[0].iMeteoTS;
  int ymax = 1, lmax = 1;
  for (int i = 0; i < n_steps; ++i) {
    lo = idx[i].iMeteoTS < lo ? idx[i].iMeteoTS : lo;
    hi = idx[i].iMeteoTS > hi ? idx[i].iMeteoTS : hi;
    ymax = idx[i].yId > ymax ? idx[i].yId : ymax;
    lmax = idx[i].iLAI > lmax ? idx[i].iLAI : lmax;
    MHM_REQUIRE(idx[i].yId >= 1 && idx[i].iLAI >= 1 && idx[i].month >= 1 && idx[i].month <= 12,
                "step index %d: yId/iLAI/month out of range", i);
  }
  MHM_REQUIRE(ymax <= d->cfg.nLCscenes, "yId %d exceeds nLCscenes %d", ymax, d->cfg.nLCscenes);
  MHM_REQUIRE(lmax <= d->cfg.nLAI, "iLAI %d exceeds nLAI %d", lmax, d->cfg.nLAI);
  for (int v : needM) {
    MHM_REQUIRE(d->met[v], "domain %d: meteo variable %d has not been set", d->id, v);
    MHM_REQUIRE(lo >= d->met_first[v] && hi < d->met_first[v] + d->met_n[v],
                "domain %d: meteo variable %d holds steps %lld..%lld, steps %lld..%lld needed",
                d->id, v, (long long)d->met_first[v], (long long)(d->met_first[v] + d->met_n[v] - 1),
                (long long)lo, (long long)hi);
  }
  if (!d->mcfg.is_hourly_forcing && d->mcfg.read_meteo_weights)
    MHM_REQUIRE(d->weights[MHM_M_PRE] && d->weights[MHM_M_TEMP] && d->weights[MHM_M_PET],
                "domain %d: meteo weights have not been set", d->id);
  return 0;
}

static void fill_args(mhm_cuda_context* ctx, Domain* d, CellArgs& a) {
  (void)ctx;
  memset(&a, 0, sizeof(a));
  a.nCells = d->cfg.nCells;
  a.nMembers = d->cfg.nMembers;
  a.nLC = d->cfg.nLCscenes;
  a.nLAI = d->cfg.nLAI;
  a.soil_case = d->soil_case;
  a.pet_case = d->pet_case;
  a.is_hourly = d->mcfg.is_hourly_forcing;
  a.read_weights = d->mcfg.read_meteo_weights;
  a.read_states = d->cfg.read_states;
  a.nTstepDay_dp = (double)(24 / d->cfg.timestep_h);
  a.c2TSTu = d->cfg.c2TSTu;
  for (int v = 0; v < MHM_M_COUNT; ++v) {
    a.met[v] = d->met[v];
    a.met_first[v] = d->met_first[v];
  }
  a.w_pre = d->weights[MHM_M_PRE];
  a.w_temp = d->weights[MHM_M_TEMP];
  a.w_pet = d->weights[MHM_M_PET];
  for (int p = 0; p < MHM_P_COUNT; ++p) a.P[p] = d->P[p];
  for (int s = 0; s < MHM_S_COUNT; ++s) a.S[s] = d->S[s];
  for (int f = 0; f < MHM_F_COUNT; ++f) a.F[f] = d->F[f];
  for (int m = 0; m < 12; ++m) {
    a.tab.fday_prec[m] = d->mcfg.fday_prec[m];
    a.tab.fnight_prec[m] = d->mcfg.fnight_prec[m];
    a.tab.fday_pet[m] = d->mcfg.fday_pet[m];
    a.tab.fnight_pet[m] = d->mcfg.fnight_pet[m];
    a.tab.fday_temp[m] = d->mcfg.fday_temp[m];
    a.tab.fnight_temp[m] = d->mcfg.fnight_temp[m];
    a.tab.evap_coeff[m] = d->mcfg.evap_coeff[m];
    a.tab.inv_evap_coeff[m] = 1.0 / d->mcfg.evap_coeff[m];
  }
}

// the main stream must see the uploads of the forcing it is about to read
static int meteo_acquire(mhm_cuda_context* ctx, Domain* d) {
  for (int v = 0; v < MHM_M_COUNT; ++v)
    if (d->met_owned[v] && d->met_ready_set[v])
      MHM_CUDA_OK(cudaStreamWaitEvent(ctx->stream, d->met_ready[v], 0));
  return 0;
}
// ... and later uploads must not overwrite a buffer a queued kernel still reads
static int meteo_release(mhm_cuda_context* ctx, Domain* d) {
  for (int v = 0; v < MHM_M_COUNT; ++v) {
    if (!d->met_owned[v]) continue;
    const int b = d->met_active[v];
    if (!d->met_free[v][b]) MHM_CUDA_OK(cudaEventCreateWithFlags(&d->met_free[v][b], cudaEventDisableTiming));
    MHM_CUDA_OK(cudaEventRecord(d->met_free[v][b], ctx->stream));
    d->met_free_set[v][b] = true;
  }
  return 0;
}

// ---- gridded outputs: window logic of mo_common_datetime_type.f90:157-184 ------------------
static inline bool out_active(const Domain* d, int32_t tt) {  // tIndex_out > 0, :621
  return tt - d->axis.warming_days * (24 / d->cfg.timestep_h) > 0;
}
static inline bool out_writes(const Domain* d, int32_t tt) {
  if (!out_active(d, tt)) return false;
  const int32_t nT = d->axis.nTimeSteps, ts = d->out_ts;
  const int32_t tIndex_out = tt - d->axis.warming_days * (24 / d->cfg.timestep_h);
  const int8_t fl = d->h_idx[(size_t)tt - 1].flags;
  if (tt == nT) return ts >= -3;
  if (ts > 0) return tIndex_out % ts == 0;
  if (ts == -1) return (fl & 1) != 0;
  if (ts == -2) return (fl & 2) != 0;
  if (ts == -3) return (fl & 4) != 0;
  return false;
}

__global__ void out_finalize_kernel(double* __restrict__ acc, double* __restrict__ win, size_t per_slot,
                                    int nslots, uint64_t avg_mask, double counter) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_slot) return;
  for (int s = 0; s < nslots; ++s) {  // writeVariableTimestep, mo_nc_output.f90:158-175
    double v = acc[(size_t)s * per_slot + i];
    if ((avg_mask >> s) & 1) v = v / counter;
    win[(size_t)s * per_slot + i] = v;
    acc[(size_t)s * per_slot + i] = 0.0;
  }
}

// ---- calibration aggregates: optidata_sim of FORCES mo_optimization_types, restated ------------
__global__ void optisim_average_kernel(double* __restrict__ col, size_t per, double counter) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < per) col[i] = col[i] / counter;
}
static inline bool optisim_flag(int32_t timeStepInput, int8_t fl) {
  return (timeStepInput == -1 && (fl & 1)) || (timeStepInput == -2 && (fl & 2)) ||
         (timeStepInput == -3 && (fl & 4));
}
static bool optisim_closes(const Domain* d, int8_t fl) {
  for (int w = 0; w < 3; ++w)
    if (d->opt_on[w] && optisim_flag(d->opt_ts[w], fl)) return true;
  return false;
}
// average_per_timestep (soil moisture, TWS) / increment_counter (ET) of a step whose date flags are fl
static int optisim_close(mhm_cuda_context* ctx, Domain* d, int8_t fl) {
  const size_t per = (size_t)d->cfg.nMembers * (size_t)d->cfg.nCells;
  for (int w = 0; w < 3; ++w) {
    if (!d->opt_on[w] || !optisim_flag(d->opt_ts[w], fl)) continue;
    if (w != 1 && d->opt_avg_ts[w] <= d->opt_ntime[w]) {
      optisim_average_kernel<<<(unsigned)((per + 255) / 256), 256, 0, ctx->stream>>>(
          d->opt_data[w] + (size_t)(d->opt_avg_ts[w] - 1) * per, per, (double)d->opt_avg_cnt[w]);
      MHM_CUDA_OK(cudaGetLastError());
    }
    d->opt_avg_ts[w] += 1;
    d->opt_avg_cnt[w] = 0;
  }
  return 0;
}

// One block of model steps = a sequence of launches of at most kIdxInline steps: the calendar
// of a launch travels inside the kernel arguments (constant bank), states are re-read at each
// launch.  With gridded outputs on, a launch also ends where an output window closes.
static int launch_cells(mhm_cuda_context* ctx, Domain* d, CellArgs& a, const StepIdx* idx,
                        bool block_mode) {
  const int32_t total = a.nSteps, tt0 = a.tt_first, wf = a.write_fluxes;
  double* const hist0 = a.runoff_hist;
  const size_t per_slot = (size_t)d->cfg.nMembers * (size_t)d->cfg.nCells;
  const bool outputs = block_mode && d->out_mask != 0;  // the per-step seam leaves outputs to the host
  const bool optisim = block_mode && (d->opt_on[0] || d->opt_on[1] || d->opt_on[2]);
  const bool aggregates = optisim || (block_mode && d->bfi_on);
  ctx->stat_begin(kStatCell);
  int rc = 0;
  int64_t launches = 0;
  for (int32_t t0 = 0; t0 < total && rc == 0;) {
    int32_t nb = total - t0 < kIdxInline ? total - t0 : kIdxInline;
    bool closes = false;
    int32_t out_first_l = 0, out_counted = 0;  // first accumulating step / steps counted into the open window
    a.out_mask = 0;
    a.agg_mask = 0;
    if (optisim) {
      // update_optisim (mo_mhm_interface_run.f90:776-857) runs after the step: a slot is closed
      // (average / increment_counter) BEFORE the step's own value is added, and the run's last
      // step adds nothing.  A launch therefore ends in front of every such step.
      const int32_t tt = tt0 + t0;
      if (out_active(d, tt))
        if (int r2 = optisim_close(ctx, d, d->h_idx[(size_t)tt - 1].flags)) return r2;
      for (int32_t t = 1; t < nb; ++t) {
        const int32_t tn = tt + t;
        if (out_active(d, tn) && (optisim_closes(d, d->h_idx[(size_t)tn - 1].flags) || tn == d->axis.nTimeSteps)) {
          nb = t;
          break;
        }
      }
    }
    if (outputs || aggregates) {
      int32_t first = nb;
      for (int32_t t = 0; t < nb; ++t) {
        const int32_t tt = tt0 + t0 + t;
        if (out_active(d, tt) && first == nb) first = t;
        a.out_yid[t] = (int16_t)(tt < d->axis.nTimeSteps ? d->h_idx[(size_t)tt].yId
                                                        : d->h_idx[(size_t)tt - 1].yId);
        if (outputs && out_writes(d, tt)) {
          nb = t + 1;
          closes = true;
          break;
        }
      }
      if (first < nb) {
        a.out_first = first;
        out_first_l = first;
        if (outputs) {
          a.out_mask = d->out_mask;
          a.out_acc = d->out_acc;
          a.out_nslots = d->out_nslots;
          d->out_counter += nb - first;
          out_counted = nb - first;
        }
        if (d->bfi_on && block_mode) {
          a.agg_mask |= 8u;
          a.bfi_acc = d->bfi_acc;
        }
        if (optisim && tt0 + t0 + first != d->axis.nTimeSteps) {  // [first, nb) never holds the last step otherwise
          const size_t per = (size_t)d->cfg.nMembers * (size_t)d->cfg.nCells;
          for (int w = 0; w < 3; ++w) {
            if (!d->opt_on[w] || d->opt_avg_ts[w] > d->opt_ntime[w]) continue;
            a.agg_mask |= 1u << w;
            a.agg_col[w] = d->opt_data[w] + (size_t)(d->opt_avg_ts[w] - 1) * per;
            if (w != 1) d->opt_avg_cnt[w] += nb - first;  // average_add
          }
          a.agg_nhor_sm = d->opt_nhor_sm;
        }
      }
    }
    // uniform calendar: the specialised kernels drop their per-step calendar work when all steps
    // of a launch share yId / iLAI / month and read consecutive meteo rows; a launch ends where
    // the calendar turns (about once a month)
    a.uniform_calendar = 0;
    a.forcing_tma = 0;
    // (the uniform kernels index a launch's forcing rows with 32 bits)
    // (with gridded outputs: only the default selection whose window the fast kernel keeps in registers)
    const bool out_uniform_ok = !a.out_mask || (a.out_mask == kOutDefaultMask && ctx->math_mode == 1 &&
                                                d->cfg.nHorizons <= 2 && !getenv("MHM_CUDA_NO_OUTPUT_REGISTERS") &&
                                                !getenv("MHM_CUDA_NO_UNIFORM_OUTPUTS"));
    const int32_t nb_before = nb;
    if (block_mode && ctx->uniform_calendar && out_uniform_ok && !a.agg_mask && a.is_hourly && a.pet_case <= 0 &&
        (uint64_t)kIdxInline * (uint64_t)a.nCells < ((uint64_t)1 << 32)) {
      const StepIdx& f = idx[t0];
      int32_t t = 1;
      for (; t < nb; ++t) {
        const StepIdx& g = idx[t0 + t];
        if (g.yId != f.yId || g.iLAI != f.iLAI || g.month != f.month || g.iMeteoTS != f.iMeteoTS + t) break;
      }
      if (t == nb || t >= 8) {  // a change in the first steps: one short general launch instead
        if (t < nb) closes = false;  // cut short: the output window stays open
        nb = t;
        a.uniform_calendar = 1;
        // TMA bulk copies need 16-byte aligned row stretches: an even number of cells, aligned bases
        // (MHM_CUDA_FORCING_TMA=0 keeps the per-lane loads)
        a.forcing_tma = ctx->forcing_tma && a.nCells % 2 == 0 &&
                        ((uintptr_t)a.met[MHM_M_PRE] | (uintptr_t)a.met[MHM_M_TEMP] | (uintptr_t)a.met[MHM_M_PET]) % 16 == 0;
      } else {
        for (t = 1; t < nb; ++t) {  // ... that ends where the uniform stretch begins
          const StepIdx &g = idx[t0 + t], &h = idx[t0 + t - 1];
          if (g.yId != h.yId || g.iLAI != h.iLAI || g.month != h.month) {
            nb = t;
            closes = false;
            break;
          }
        }
      }
    }
    if (a.out_mask && nb != nb_before) {  // the launch was cut short: fewer steps enter the open window
      const int32_t now = nb > out_first_l ? nb - out_first_l : 0;
      d->out_counter -= out_counted - now;
      if (now == 0) a.out_mask = 0;
    }
    a.nSteps = nb;
    a.tt_first = tt0 + t0;
    a.write_fluxes = (wf && t0 + nb >= total) ? 1 : 0;
    a.runoff_hist = hist0 ? hist0 + (size_t)t0 * per_slot : nullptr;
    a.qout_step0 = t0;
    memcpy(a.idx_in, idx + t0, (size_t)nb * sizeof(StepIdx));
    rc = ctx->math_mode == 1 ? launch_cell_block_fast(a, d->cfg.nHorizons, ctx->stream)
                             : launch_cell_block_strict(a, d->cfg.nHorizons, ctx->stream);
    if (ctx->launch_log)  // MHM_CUDA_LAUNCH_LOG: one line per cell-kernel launch (profiles/cell_profile.py)
      fprintf(ctx->launch_log, "cell %d %d %d %d %d\n", a.tt_first, nb, a.nCells, a.nMembers, a.uniform_calendar);
    ++launches;
    if (rc == 0 && closes) {
      const size_t w = d->out_win_tt.size();
      const size_t need = (w + 1) * (size_t)d->out_nslots * per_slot;
      if (need > d->out_win_cap) {
        rc = 2;  // sized by run_steps for every window of the call
      } else {
        out_finalize_kernel<<<(unsigned)((per_slot + 255) / 256), 256, 0, ctx->stream>>>(
            d->out_acc, d->out_win + w * (size_t)d->out_nslots * per_slot, per_slot, d->out_nslots,
            d->out_avg_mask, (double)d->out_counter);
        rc = (int)cudaGetLastError();
        d->out_win_tt.push_back(tt0 + t0 + nb - 1);
        d->out_counter = 0;
      }
    }
    t0 += nb;
  }
  ctx->stat_end(kStatCell, launches);
  MHM_REQUIRE(rc == 0, "cell kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  return 0;
}

int mhm_cuda_set_outputs(mhm_cuda_context* ctx, int32_t iDomain, const int32_t* outputFlxState,
                         int32_t timeStep_model_outputs) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(outputFlxState, "set_outputs: null outputFlxState");
  MHM_REQUIRE(!outputFlxState[17], "set_outputs: variable 18 (neutrons) is not supported");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  // order of the `ii = ii + 1` blocks of mHM_updateDataset, mo_write_fluxes_states.f90:326-436
  static const int order[20] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 20, 21};
  d->out_mask = 0;
  d->out_avg_mask = 0;
  d->out_slot_var.clear();
  d->out_slot_hor.clear();
  for (int v : order) {
    if (!outputFlxState[v - 1]) continue;
    d->out_mask |= 1u << v;
    const bool per_h = v == 3 || v == 4 || v == 17 || v == 19;
    for (int h = 0; h < (per_h ? d->cfg.nHorizons : 1); ++h) {
      if (v <= 8) d->out_avg_mask |= (uint64_t)1 << d->out_slot_var.size();
      d->out_slot_var.push_back((int8_t)v);
      d->out_slot_hor.push_back((int8_t)(per_h ? h : -1));
    }
  }
  d->out_nslots = (int32_t)d->out_slot_var.size();
  MHM_REQUIRE(d->out_nslots <= 64, "set_outputs: too many output slots (%d)", d->out_nslots);
  d->out_ts = timeStep_model_outputs;
  d->out_counter = 0;
  d->out_win_tt.clear();
  cudaFree(d->out_acc);
  d->out_acc = nullptr;
  if (d->out_nslots > 0) {
    const size_t bytes = (size_t)d->out_nslots * d->cfg.nMembers * d->cfg.nCells * sizeof(double);
    MHM_CUDA_OK(cudaMalloc(&d->out_acc, bytes));
    MHM_CUDA_OK(cudaMemsetAsync(d->out_acc, 0, bytes, ctx->stream));
  }
  return 0;
}

int mhm_cuda_set_optisim(mhm_cuda_context* ctx, int32_t iDomain, const mhm_optisim_config* cfg) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(cfg, "set_optisim: null config");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const size_t per = (size_t)d->cfg.nMembers * (size_t)d->cfg.nCells;
  const int32_t on[3] = {cfg->sm_on, cfg->et_on, cfg->tws_on};
  const int32_t ts[3] = {cfg->sm_timeStepInput, cfg->et_timeStepInput, cfg->tws_timeStepInput};
  const int32_t nt[3] = {cfg->sm_nTime, cfg->et_nTime, cfg->tws_nTime};
  for (int w = 0; w < 3; ++w) {
    if (on[w]) {
      MHM_REQUIRE(ts[w] >= -3 && ts[w] <= -1, "set_optisim: timeStepInput must be -1, -2 or -3 (got %d)", ts[w]);
      MHM_REQUIRE(nt[w] >= 1, "set_optisim: nTime must be positive");
    }
  }
  MHM_REQUIRE(!cfg->sm_on || (cfg->nSoilHorizons_sm_input >= 1 && cfg->nSoilHorizons_sm_input <= d->cfg.nHorizons),
              "set_optisim: nSoilHorizons_sm_input %d outside 1..%d", cfg->nSoilHorizons_sm_input,
              d->cfg.nHorizons);  // mo_mhm_read_config.f90:189
  for (int w = 0; w < 3; ++w) {
    cudaFree(d->opt_data[w]);
    d->opt_data[w] = nullptr;
    d->opt_on[w] = on[w] != 0;
    d->opt_ts[w] = ts[w];
    d->opt_ntime[w] = on[w] ? nt[w] : 0;
    d->opt_avg_ts[w] = 1;  // optidata_sim%init
    d->opt_avg_cnt[w] = 0;
    if (on[w]) {
      MHM_CUDA_OK(cudaMalloc(&d->opt_data[w], (size_t)nt[w] * per * sizeof(double)));
      MHM_CUDA_OK(cudaMemsetAsync(d->opt_data[w], 0, (size_t)nt[w] * per * sizeof(double), ctx->stream));
    }
  }
  d->opt_nhor_sm = cfg->nSoilHorizons_sm_input;
  cudaFree(d->bfi_acc);
  d->bfi_acc = nullptr;
  d->bfi_on = cfg->bfi_on != 0;
  if (d->bfi_on) {
    MHM_CUDA_OK(cudaMalloc(&d->bfi_acc, 2 * per * sizeof(double)));
    MHM_CUDA_OK(cudaMemsetAsync(d->bfi_acc, 0, 2 * per * sizeof(double), ctx->stream));
  }
  return 0;
}

int mhm_cuda_get_optisim(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t which,
                         double* base, int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(base && member >= 0 && member < d->cfg.nMembers && which >= 0 && which < 3 && d->opt_on[which],
              "get_optisim: bad member, or aggregate %d is not enabled", which);
  const size_t n = (size_t)d->cfg.nCells, per = n * (size_t)d->cfg.nMembers;
  MHM_REQUIRE(ld >= (int64_t)n && offset >= 0 && offset + (int64_t)n <= ld, "get_optisim: bad ld/offset");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaMemcpy2DAsync(base + offset, (size_t)ld * sizeof(double), d->opt_data[which] + (size_t)member * n,
                                per * sizeof(double), n * sizeof(double), (size_t)d->opt_ntime[which],
                                cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mhm_cuda_get_bfi_sums(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, const double* cellArea,
                          double* qBF_sum, double* qT_sum) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(d->bfi_on && cellArea && qBF_sum && qT_sum && member >= 0 && member < d->cfg.nMembers,
              "get_bfi_sums: BFI sums are not enabled, or bad arguments");
  const size_t n = (size_t)d->cfg.nCells, per = n * (size_t)d->cfg.nMembers;
  std::vector<double> h(2 * n);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaMemcpyAsync(h.data(), d->bfi_acc + (size_t)member * n, n * sizeof(double),
                              cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaMemcpyAsync(h.data() + n, d->bfi_acc + per + (size_t)member * n, n * sizeof(double),
                              cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  // sum over the evaluation period of sum(q * CellArea) / nCells, summed per cell first
  double sb = 0.0, st = 0.0;
  for (size_t k = 0; k < n; ++k) sb += h[k] * cellArea[k];
  for (size_t k = 0; k < n; ++k) st += h[n + k] * cellArea[k];
  *qBF_sum = sb / (double)n;
  *qT_sum = st / (double)n;
  return 0;
}

int mhm_cuda_get_output_windows(mhm_cuda_context* ctx, int32_t iDomain, int32_t* n_windows,
                                int32_t* tt_end, int32_t capacity) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(n_windows, "get_output_windows: null n_windows");
  *n_windows = (int32_t)d->out_win_tt.size();
  for (int32_t w = 0; tt_end && w < *n_windows && w < capacity; ++w) tt_end[w] = d->out_win_tt[(size_t)w];
  return 0;
}

int mhm_cuda_get_output(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t window,
                        int32_t variable, int32_t horizon, double* out) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(out && member >= 0 && member < d->cfg.nMembers && window >= 0 &&
                  window < (int32_t)d->out_win_tt.size(),
              "get_output: bad member/window (%d windows closed by the last run_steps)",
              (int)d->out_win_tt.size());
  int slot = -1;
  for (int s = 0; s < d->out_nslots; ++s)
    if (d->out_slot_var[(size_t)s] == variable &&
        (d->out_slot_hor[(size_t)s] < 0 || d->out_slot_hor[(size_t)s] == horizon - 1))
      slot = s;
  MHM_REQUIRE(slot >= 0, "get_output: variable %d (horizon %d) is not enabled", variable, horizon);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells, per_slot = n * (size_t)d->cfg.nMembers;
  MHM_CUDA_OK(cudaMemcpyAsync(out, d->out_win + ((size_t)window * d->out_nslots + slot) * per_slot + (size_t)member * n,
                              n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mhm_cuda_set_math_mode(mhm_cuda_context* ctx, int32_t mode) {
  MHM_REQUIRE(ctx && (mode == 0 || mode == 1), "set_math_mode: mode must be 0 or 1");
  ctx->math_mode = mode;
  return 0;
}

int mhm_cuda_cell_step(mhm_cuda_context* ctx, int32_t iDomain, int32_t tt,
                       const mhm_step_index* idx) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(idx && tt >= 1, "cell_step: bad arguments");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  StepIdx s{};
  s.iMeteoTS = idx->iMeteoTS;
  s.yId = idx->yId;
  s.iLAI = idx->iLAI;
  s.doy = idx->doy;
  s.year = (int16_t)idx->year;
  s.month = idx->month;
  s.hour = idx->hour;
  s.isday = idx->isday;
  if (int rc = check_inputs(d, &s, 1)) return rc;
  CellArgs a;
  fill_args(ctx, d, a);
  a.nSteps = 1;
  a.tt_first = tt;
  a.write_fluxes = 1;
  a.runoff_hist = nullptr;
  d->last_yId = s.yId;
  d->hist_steps = 0;
  if (int rc = meteo_acquire(ctx, d)) return rc;
  if (int rc = launch_cells(ctx, d, a, &s, false)) return rc;
  return meteo_release(ctx, d);
}

int mhm_cuda_run_steps(mhm_cuda_context* ctx, int32_t iDomain, int32_t tt_first, int32_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  if (int rc = ensure_calendar(ctx, d)) return rc;
  ctx->uniform_calendar = getenv("MHM_CUDA_NO_UNIFORM_CALENDAR") == nullptr;
  if (const char* e = getenv("MHM_CUDA_FORCING_TMA")) ctx->forcing_tma = atoi(e) != 0;
  MHM_REQUIRE(tt_first >= 1 && n_steps >= 1 && tt_first + n_steps - 1 <= d->axis.nTimeSteps,
              "run_steps: steps %d..%d outside 1..%d", tt_first, tt_first + n_steps - 1,
              d->axis.nTimeSteps);
  if (int rc = check_inputs(d, d->h_idx.data() + (tt_first - 1), n_steps)) return rc;
  const size_t n = (size_t)d->cfg.nCells, M = (size_t)d->cfg.nMembers;
  // time blocks sized by the history-buffer budget (runoff + routing histories)
  CellArgs probe;
  fill_args(ctx, d, probe);
  const bool fused = routing_fuse_qout(ctx, d, 1, &probe);  // routing input written by the cells
  size_t per_step = (fused ? 2 : 3) * M * n * sizeof(double);
  int32_t tb = (int32_t)(ctx->block_bytes / per_step);
  if (tb < 1) tb = 1;
  if (tb > n_steps) tb = n_steps;
  MHM_REQUIRE(!(d->rt && routing_is_deferred(d)) || n_steps <= tb,
              "run_steps: with deferred routing a call must fit one time block (%d steps)", tb);
  const size_t need = (size_t)tb * M * n;
  if (!fused && d->runoff_cap < need) {
    MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d->runoff_hist);
    d->runoff_hist = nullptr;
    d->runoff_cap = 0;
    MHM_CUDA_OK(cudaMalloc(&d->runoff_hist, need * sizeof(double)));
    d->runoff_cap = need;
  }
  d->out_win_tt.clear();
  if (d->out_mask) {  // every window this call closes gets a slot
    size_t nwin = 0;
    for (int32_t tt = tt_first; tt < tt_first + n_steps; ++tt) nwin += out_writes(d, tt) ? 1 : 0;
    const size_t need_w = nwin * (size_t)d->out_nslots * M * n;
    if (need_w > d->out_win_cap) {
      MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      cudaFree(d->out_win);
      d->out_win = nullptr;
      d->out_win_cap = 0;
      MHM_CUDA_OK(cudaMalloc(&d->out_win, need_w * sizeof(double)));
      d->out_win_cap = need_w;
    }
  }
  if (int rc = meteo_acquire(ctx, d)) return rc;
  for (int32_t t0 = 0; t0 < n_steps; t0 += tb) {
    const int32_t nb = (n_steps - t0 < tb) ? n_steps - t0 : tb;
    CellArgs a;
    fill_args(ctx, d, a);
    a.nSteps = nb;
    a.tt_first = tt_first + t0;
    a.write_fluxes = (t0 + nb >= n_steps) ? 1 : 0;  // fluxes of the call's last step
    a.runoff_hist = fused ? nullptr : d->runoff_hist;
    if (fused) routing_fuse_qout(ctx, d, nb, &a);
    if (int rc = launch_cells(ctx, d, a, d->h_idx.data() + (tt_first + t0 - 1), true)) return rc;
    d->hist_steps = fused ? 0 : nb;
    d->hist_tt_first = tt_first + t0;
    d->last_yId = d->h_idx[(size_t)(tt_first + t0 + nb - 2)].yId;
    if (d->rt && !routing_defer_block(d, tt_first + t0, nb, fused))
      if (int rc = routing_run_block(ctx, d, tt_first + t0, nb, fused)) return rc;
  }
  return meteo_release(ctx, d);
}

int mhm_cuda_keep_runoff_history(mhm_cuda_context* ctx, int32_t iDomain, int32_t keep) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  d->keep_runoff_hist = keep != 0;
  return 0;
}

int mhm_cuda_get_runoff_history(mhm_cuda_context* ctx, int32_t iDomain, int32_t member,
                                double* out, int64_t ld) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(member >= 0 && member < d->cfg.nMembers, "get_runoff_history: bad member");
  MHM_REQUIRE(d->hist_steps > 0 && d->runoff_hist,
              "get_runoff_history: no block has been run, or the history was fused into the routing "
              "input (call mhm_cuda_keep_runoff_history first)");
  MHM_REQUIRE(out && ld >= d->cfg.nCells, "get_runoff_history: bad out/ld");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells, M = (size_t)d->cfg.nMembers;
  MHM_CUDA_OK(cudaMemcpy2DAsync(out, (size_t)ld * sizeof(double), d->runoff_hist + (size_t)member * n,
                                M * n * sizeof(double), n * sizeof(double), (size_t)d->hist_steps,
                                cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// --------------------------------------------------------------------------- host coherence
int mhm_cuda_bind_host_state(mhm_cuda_context* ctx, int32_t iDomain, int32_t id, double* base,
                             int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(id >= 0 && id < MHM_S_COUNT && ld >= d->cfg.nCells && offset >= 0,
              "bind_host_state: bad arguments");
  d->sbind[id] = HostBind{base, ld, offset};
  return 0;
}
int mhm_cuda_bind_host_flux(mhm_cuda_context* ctx, int32_t iDomain, int32_t id, double* base,
                            int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(id >= 0 && id < MHM_F_COUNT && ld >= d->cfg.nCells && offset >= 0,
              "bind_host_flux: bad arguments");
  d->fbind[id] = HostBind{base, ld, offset};
  return 0;
}
int mhm_cuda_sync_to_host(mhm_cuda_context* ctx, int32_t iDomain) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells;
  for (int s = 0; s < MHM_S_COUNT; ++s) {
    const HostBind& b = d->sbind[s];
    if (!b.base) continue;
    MHM_CUDA_OK(cudaMemcpy2DAsync(b.base + b.offset, (size_t)b.ld * sizeof(double), d->S[s],
                                  n * sizeof(double), n * sizeof(double), state_rows(d, s),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  }
  for (int f = 0; f < MHM_F_COUNT; ++f) {
    const HostBind& b = d->fbind[f];
    if (!b.base) continue;
    MHM_CUDA_OK(cudaMemcpy2DAsync(b.base + b.offset, (size_t)b.ld * sizeof(double), d->F[f],
                                  n * sizeof(double), n * sizeof(double), flux_rows(d, f),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  }
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------ measurement
int mhm_cuda_event_record(mhm_cuda_context* ctx, int32_t slot) {
  MHM_REQUIRE(ctx && slot >= 0 && slot < 16, "event_record: bad slot");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaEventRecord(ctx->ev[slot], ctx->stream));
  return 0;
}
int mhm_cuda_event_elapsed_ms(mhm_cuda_context* ctx, int32_t a, int32_t b, double* ms) {
  MHM_REQUIRE(ctx && ms && a >= 0 && a < 16 && b >= 0 && b < 16, "event_elapsed_ms: bad arguments");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaEventSynchronize(ctx->ev[b]));
  float f = 0.f;
  MHM_CUDA_OK(cudaEventElapsedTime(&f, ctx->ev[a], ctx->ev[b]));
  *ms = f;
  return 0;
}
int mhm_cuda_synchronize(mhm_cuda_context* ctx) {
  MHM_REQUIRE(ctx, "synchronize: null context");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->copy_stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  MHM_CUDA_OK(cudaGetLastError());
  return 0;
}
int mhm_cuda_kernel_stats(mhm_cuda_context* ctx, int32_t which, double* ms, int64_t* launches) {
  MHM_REQUIRE(ctx && which >= 0 && which < kStatCount, "kernel_stats: bad class");
  if (int rc = ctx->stat_flush()) return rc;
  if (ms) *ms = ctx->stat_ms[which];
  if (launches) *launches = ctx->stat_launches[which];
  return 0;
}
int mhm_cuda_kernel_stats_reset(mhm_cuda_context* ctx, int32_t enable_timing) {
  MHM_REQUIRE(ctx, "kernel_stats_reset: null context");
  if (int rc = ctx->stat_flush()) return rc;
  for (int i = 0; i < kStatCount; ++i) {
    ctx->stat_ms[i] = 0.0;
    ctx->stat_launches[i] = 0;
  }
  ctx->timing = enable_timing != 0;
  return 0;
}

// dependent-chain-free DFMA loop: 8 independent accumulators per thread
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int mhm_cuda_measure_dfma_peak(mhm_cuda_context* ctx, double* dfma_per_s) {
  MHM_REQUIRE(ctx && dfma_per_s, "measure_dfma_peak: bad arguments");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  MHM_CUDA_OK(cudaGetDeviceProperties(&prop, ctx->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* buf = nullptr;
  MHM_CUDA_OK(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, ctx->stream);
    MHM_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)blocks * threads * iters * 8.0 / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *dfma_per_s = best;
  return 0;
}

}  // extern "C"
