// context.h -- host-side objects behind the opaque handles of include/mhm_cuda.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "calendar.h"
#include "device_types.h"

namespace mhm {

void set_error(const char* fmt, ...);

#define MHM_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::mhm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                      \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define MHM_REQUIRE(cond, ...)     \
  do {                             \
    if (!(cond)) {                 \
      ::mhm::set_error(__VA_ARGS__); \
      return 1;                    \
    }                              \
  } while (0)

struct HostBind {
  double* base = nullptr;
  int64_t ld = 0, offset = 0;
};

struct Routing;   // routing.cu
struct MprState;  // mpr.cu

// kernel classes for mhm_cuda_kernel_stats
enum { kStatCell = 0, kStatRouting = 1, kStatUpscale = 2, kStatCount = 3 };

struct Domain {
  int32_t id = 0;
  mhm_domain_config cfg{};
  std::vector<int32_t> processMatrix;
  int soil_case = 1, pet_case = 0, rout_case = 0;

  // effective parameters, device [member][dim3][dim2][nCells]
  double* P[MHM_P_COUNT] = {};
  int P_dim2[MHM_P_COUNT] = {}, P_dim3[MHM_P_COUNT] = {};
  // states / fluxes, device [member][(nH)][nCells]
  double* S[MHM_S_COUNT] = {};
  double* F[MHM_F_COUNT] = {};
  HostBind sbind[MHM_S_COUNT], fbind[MHM_F_COUNT];

  // meteo
  bool has_meteo_cfg = false;
  mhm_meteo_config mcfg{};
  double* met[MHM_M_COUNT] = {};      // buffer the next run reads (owned or bound zero-copy)
  bool met_owned[MHM_M_COUNT] = {};
  int64_t met_first[MHM_M_COUNT] = {}, met_n[MHM_M_COUNT] = {};
  // owned forcing is double buffered: an upload goes (on the copy stream) into the buffer the
  // running block does not read, so H2D of chunk c+1 overlaps the kernels of chunk c
  double* met_buf[MHM_M_COUNT][2] = {};
  size_t met_bufcap[MHM_M_COUNT][2] = {};   // capacity in doubles
  int met_active[MHM_M_COUNT] = {};
  cudaEvent_t met_ready[MHM_M_COUNT] = {};  // upload finished (copy stream)
  cudaEvent_t met_free[MHM_M_COUNT][2] = {};  // last run reading the buffer finished (main stream)
  bool met_ready_set[MHM_M_COUNT] = {}, met_free_set[MHM_M_COUNT][2] = {};
  float* met_f32 = nullptr;       // staging of float32-on-the-wire chunks (mhm_cuda_set_meteo_shared)
  size_t met_f32cap = 0;
  size_t met_h2d_bytes = 0;       // host bytes this process copied for forcing (measurement)
  double* weights[3] = {};

  // time axis
  bool has_time = false;
  TimeAxis axis;
  std::vector<StepIdx> h_idx;  // steps 1..nTimeSteps

  // total-runoff history of the last block, device [steps][member][nCells]
  double* runoff_hist = nullptr;
  size_t runoff_cap = 0;
  int32_t hist_steps = 0, hist_tt_first = 0;
  bool keep_runoff_hist = false;  // mhm_cuda_keep_runoff_history: never fuse the history away

  // gridded outputs (mhm_cuda_set_outputs): slots in mHM_updateDataset order
  uint32_t out_mask = 0;
  int32_t out_ts = 0, out_nslots = 0, out_counter = 0;
  uint64_t out_avg_mask = 0;            // bit s: slot s is averaged over its window
  std::vector<int8_t> out_slot_var, out_slot_hor;
  double* out_acc = nullptr;            // [slot][member][nCells], the open window
  double* out_win = nullptr;            // [window of the last run_steps][slot][member][nCells]
  size_t out_win_cap = 0;
  std::vector<int32_t> out_win_tt;      // model step that closed each window

  // calibration aggregates (mhm_cuda_set_optisim): 0 soil moisture, 1 evapotranspiration, 2 TWS
  bool opt_on[3] = {};
  int32_t opt_ts[3] = {}, opt_ntime[3] = {}, opt_nhor_sm = 0;
  int32_t opt_avg_ts[3] = {1, 1, 1}, opt_avg_cnt[3] = {};  // optidata_sim averageTimestep / averageCounter
  double* opt_data[3] = {};             // dataSim, [time][member][nCells]
  bool bfi_on = false;
  double* bfi_acc = nullptr;            // [2][member][nCells]: sums of baseflow, total runoff

  Routing* rt = nullptr;
  MprState* mpr = nullptr;
  int32_t last_yId = 1;  // scene of the last executed step (routing parameters, per-step seam)
};

struct TimedLaunch {
  cudaEvent_t a, b;
  int which;
};

}  // namespace mhm

struct mhm_cuda_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // forcing uploads
  std::map<int32_t, mhm::Domain*> domains;
  cudaEvent_t ev[16] = {};
  int math_mode = 0;
  // kernel statistics
  bool timing = false;
  double stat_ms[mhm::kStatCount] = {};
  int64_t stat_launches[mhm::kStatCount] = {};
  std::vector<mhm::TimedLaunch> pending;
  std::vector<cudaEvent_t> ev_pool;
  size_t block_bytes = (size_t)24 << 30;  // memory budget of per-block history buffers
  int sm_count = 148;                      // streaming multiprocessors of the device
  bool uniform_calendar = true;            // MHM_CUDA_NO_UNIFORM_CALENDAR (diagnostics) switches it off
  bool forcing_tma = false;                // forcing rows through cp.async.bulk + mbarrier (MHM_CUDA_FORCING_TMA)
  // the GPUs of one box, one process each (comm.cu): NCCL communicator owned by the library
  FILE* launch_log = nullptr;              // MHM_CUDA_LAUNCH_LOG (diagnostics)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;

  // bracket a kernel launch with events when timing is enabled
  void stat_begin(int which);
  void stat_end(int which, int64_t launches = 1);
  int stat_flush();
};

namespace mhm {
Domain* find_domain(mhm_cuda_context* ctx, int32_t iDomain);
// routing hooks used by api.cu
void routing_free(Routing* rt);
void mpr_free(MprState* s);
int routing_run_block(mhm_cuda_context* ctx, Domain* d, int32_t tt_first, int32_t n_steps,
                      bool qout_ready);
// can the cell kernel write the routing's node runoff itself for a block of n_steps?  (one cell
// per node, one model step per routing event, no inflow gauges); fills the CellArgs fields
// deferred routing (sub-catchment sharding): remember the block instead of routing it
bool routing_is_deferred(const Domain* d);
bool routing_defer_block(Domain* d, int32_t tt_first, int32_t n_steps, bool fused);
bool routing_fuse_qout(mhm_cuda_context* ctx, Domain* d, int32_t n_steps, CellArgs* a);
// comm.cu: one grouped NCCL exchange of doubles; send_counts / recv_counts are per peer rank, the
// pieces consecutive in `send` / `recv` by ascending rank; enqueued on `st`
int comm_send_recv(mhm_cuda_context* ctx, const double* send, const size_t* send_counts, double* recv,
                   const size_t* recv_counts, cudaStream_t st);
}  // namespace mhm
