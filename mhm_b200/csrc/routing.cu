// routing.cu -- mRM Muskingum routing on the device: time-blocked, chain-decomposed,
// warp-pipelined.
//
// Reference behaviour restated here (never copied):
//   mRM/mo_mrm_routing.f90:104-303 (mRM_routing), :380-481 (L11_routing)
//   mRM/mo_mrm_pre_routing.f90:77-143 (L11_runoff_acc), :179-214 (add_inflow)
//   mRM/mo_mrm_mpr.f90:61-119 (reg_rout)
//   mRM/mo_mrm_net_startup.f90:728-859 (L11_routing_order)
//   mHM/mo_mhm_interface_run.f90:460-612 (routing schedule, gauge back-fill)
//
// The reference sweeps the links serially in netPerm order once per routing step.  Its data
// dependence is (node, step) <- (upstream nodes, same step) and (node, step - 1).  Here a
// whole block of routing steps is routed at once:
//   * the river network is cut into chains (every node continues the chain of its highest
//     upstream node; the other inflowing links are tributaries) and the chains into segments
//     of at most 32 nodes;
//   * one warp owns one or several whole segments, lane j owning the j-th node, and runs a
//     software pipeline skewed by one routing step per lane: while lane j works on step s,
//     lane j+1 works on step s-1 and receives lane j's routed outflow through a register
//     shuffle.  Muskingum state stays in registers for the whole block;
//   * segments are level-scheduled on the (shallow) segment DAG: O(log N + longest chain/32)
//     launches per block instead of one per network level; only segment ends and tributary
//     mouths ever write their outflow series to memory.
// Upstream inflows are added in netPerm order and the node's own runoff last, exactly like
// the serial sweep, so the result is bit-identical to it.
//
// This file is compiled with -fmad=false: C1*(a-b) + C2*(c-d) must round as in Fortran.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "context.h"

namespace mhm {

constexpr int kSegLen = 32;   // nodes per segment = lanes of a warp
constexpr int kMetaUps = 4;   // upstream links described in the lane record itself
constexpr int kUpShuffle = -1;  // LaneMeta::up value: "outflow of the previous lane"

enum : int32_t {
  kEntValid = 1,      // lane holds a node (otherwise padding)
  kEntLink = 2,       // node has an outgoing link (Muskingum state); otherwise an outlet
  kEntAddQout = 4,    // node's own runoff is added to its inflow
  kEntZeroOut = 8,    // routed outflow feeds a non-headwater inflow gauge: set to zero
  kEntInflow = 16,    // node is an inflow gauge: add_inflow applies to its runoff
  kEntWriteHist = 32, // somebody reads this node's outflow series from memory
  kEntGhost = 64,     // from-node of a cut link owned by another shard: outflow is prescribed
  kEntLeafWarp = 128, // every lane of the warp is a headwater single-node segment (position = write shift)
};

// everything a lane needs to know about its node, one 32-byte record
struct alignas(16) LaneMeta {
  int32_t node;   // 0-based L11 node
  int32_t link;   // 0-based link (C1/C2 index)
  int32_t flags;  // kEnt* bits | number of upstream links << 8 | position in segment << 16 (5 bits)
                  // | write shift of the outflow series << 21 (3 bits), see kHistPad
  int32_t gslot;  // gauge slot or -1
  int32_t up[kMetaUps];  // first four upstream links in netPerm order: lane position of the
                         // link's from-node, or kUpShuffle
};
static_assert(sizeof(LaneMeta) == 32, "LaneMeta must be 32 bytes");

// History buffers (node runoff qOUT per event, routed outflow qTR per routing step) are tiled
// by 8 steps: [step / 8][member][lane][step % 8]; a node's 8 consecutive steps are one
// 64-byte run and neighbouring lanes are neighbouring runs.
constexpr int kHistTile = 8;
// The node-runoff history (qout_hist) is stored SKEWED: event e of lane p sits in slot
// e + skew(p), skew = position of the node in its segment (0..31).  In sub-step d of macro step
// S the routing pipeline works on event kWin * S + d - skew, i.e. on slot kWin * S + d for every
// lane: the lanes of a warp read the same aligned slots and the lean kernel fetches the kWin
// values of a macro step with one 256-bit load per lane (with unskewed storage every 8-byte
// load touched 32 different 64-byte runs and the L1 tag stage bounded the kernel).
// The routed-outflow history (qtr_hist) is stored SHIFTED for its one reader from memory (the
// lane of the downstream node, when that is not the next lane of the same segment): step r of a
// lane sits in slot r + ws, ws = (position of the reader in its segment) & 7.  The reader works on
// step 8 * k + j - skew in position j of its k-th tile, i.e. on slot 8 * (k - skew / 8) + j of each
// tributary row: the same window position as its node runoff, so a half window of a tributary is
// one aligned 32-byte piece fetched with one 256-bit load (with unshifted rows the four 8-byte
// loads of a window touched 32 different runs each and the load/store unit bounded the levels
// below the headwaters).  A warp that holds nothing but
// headwater single-node segments streams whole 64-byte runs; there the position field itself
// carries ws, so that node runoff in and routed outflow out share the shift and both stay aligned.
constexpr int kHistPad = 48;  // slots past the last event: largest skew + window overshoot
__host__ __device__ __forceinline__ int lane_skew(int flags) { return (flags >> 16) & 31; }
__host__ __device__ __forceinline__ int lane_wshift(int flags) { return (flags >> 21) & 7; }
__host__ __device__ __forceinline__ size_t hidx(int step, int M, int E, int m, int p) {
  return ((((size_t)(step >> 3) * M + m) * E + p) << 3) + (size_t)(step & 7);
}
__host__ __device__ __forceinline__ size_t hist_size(int steps, int M, int E) {
  return (size_t)((steps + kHistTile - 1) / kHistTile) * M * E * kHistTile;
}

struct DevEvent {
  int32_t tt;         // model step at which mRM_routing is called
  int32_t t0;         // first model step (index into the block) whose runoff is accumulated
  int32_t nacc;       // number of block steps accumulated
  int32_t use_carry;  // start from the carried RunToRout of the previous block
  int32_t rout_loop;  // routing sub-steps (mo_mrm_routing.f90:224)
  int32_t rs_first;   // index of the first sub-step in the block's outflow history
  int32_t backfill;   // nint(tsRoutFactorIn) when the gauge series is back-filled, else 0
  int32_t pad;
  double tst;         // HourSecs * timestep_rout
};

struct Routing {
  int32_t nCells1 = 0, nNodes = 0, nLinks = 0, nOutlets = 0, map_flag = 1, rout_case = 1;
  int32_t nGauges = 0, nInflowGauges = 0, nGaugesTotal = 0, nInflowTotal = 0, M = 1;
  int32_t E = 0;                 // lanes = nodes + padding
  std::vector<int32_t> lvl_ptr;  // lane range per segment level (multiples of 32)
  std::vector<int32_t> lvl_ku, lvl_mem;  // per level: upstream slots in use, any memory tributary
  std::vector<int32_t> lvl_plain;        // per level: no ghost / zeroed-outflow lanes, <= kMetaUps inflowing links
  bool lean_ok = true;                   // MHM_CUDA_NO_LEAN_ROUTING (diagnostics) forces the general kernel
  size_t pf_waves = 1;                   // levels of up to this many waves of CTAs run the prefetching variant
  std::vector<int32_t> gaugeIndexList, gaugeNodeList, inflowIndexList, inflowHeadwater,
      inflowNodeList;
  // device topology (shared by members)
  LaneMeta* meta = nullptr;
  int32_t *up_ptr = nullptr, *up_pos = nullptr;  // all upstream lanes (nodes with > 4 links)
  int32_t* node_lane = nullptr;                  // [nNodes] lane of a node
  int32_t *cell_ptr = nullptr, *cell_idx = nullptr;  // map_flag: L1 cells of each node, ascending
  int32_t* L11_L1_Id = nullptr;                      // !map_flag
  int32_t *d_inflow_node = nullptr, *d_inflow_index = nullptr, *d_inflow_head = nullptr;
  double *L1_area = nullptr, *L11_area = nullptr;
  int32_t nGslots = 0;
  int32_t *d_gauge_col = nullptr, *d_gauge_slot = nullptr;  // per gauge: column-1, slot
  // one-cell-per-node mapping: coalesced per-cell qOUT kernel
  bool bijective = false;
  // sub-catchment sharding
  int32_t nGhost = 0, nExport = 0;
  int32_t *d_ghost_lane = nullptr, *d_export_lane = nullptr;
  bool deferred = false;
  int32_t pend_tt = 0, pend_n = 0;  // block whose cells ran but whose routing is pending
  bool pend_fused = false;
  int32_t last_n = 0;               // steps of the last routed block (export)
  // cut-link exchange below the C ABI (mrm_cuda_set_exchange): links sent to / received from each
  // rank; the export / ghost lists are grouped by peer rank.  Device tables per list entry: first
  // link of its peer's piece, links in the piece, position inside the piece.
  std::vector<int32_t> xsend, xrecv;
  int32_t *d_xs_tab = nullptr, *d_xr_tab = nullptr;  // [3][nExport] / [3][nGhost]
  double *xsend_buf = nullptr, *xrecv_buf = nullptr;
  size_t xsend_cap = 0, xrecv_cap = 0;
  int32_t* d_cell_entry = nullptr;  // [nCells1] lane of the cell's node
  int8_t* d_cell_skew = nullptr;    // [nCells1] position of that node in its segment
  double* d_cell_area = nullptr;    // [nCells1] area factor of mo_mrm_pre_routing.f90:125/:141
  // per member state, device [M][...]
  double *C1 = nullptr, *C2 = nullptr, *qOUT = nullptr, *qMod = nullptr;
  double *qTIN = nullptr, *qTR = nullptr;  // [M][2][nNodes]
  // reg_rout inputs
  std::vector<std::vector<double>> param5;  // per member, empty = C1/C2 given
  std::vector<int32_t> c1c2_yId;            // scene the member's C1/C2 were computed for
  double *d_length = nullptr, *d_slope = nullptr, *d_fFPimp = nullptr;  // fFPimp [M][nLC][nNodes]
  double ssMax = 0.0, ssMax_global = 0.0;  // ssMax_global: maxval(slope) of the unsharded domain
  int32_t nLC = 1;
  double TSrout = 0.0;  // case 2/3 [s]: the members' common value, settled when a block is routed
  std::vector<double> TSrout_m;  // per member, as given to mrm_cuda_set_c1c2 (0 = not set)
  // inflow series, host (nDays, nInflowTotal) Fortran layout
  std::vector<double> inflowQ;
  int64_t nDays = 0;
  // scheduler state carried across blocks (mo_mhm_interface_run.f90:460-514)
  int32_t carry_steps = 0;
  std::vector<double> inflow_acc;
  double* carry = nullptr;  // device [M][nCells1]
  // gauge series, device [M][nGaugesTotal][nTimeSteps]
  double* gauge_hist = nullptr;
  int32_t nTimeSteps = 0;
  // block buffers
  double *qout_hist = nullptr, *qtr_hist = nullptr, *qmod_g = nullptr, *d_inflow_val = nullptr;
  DevEvent* d_events = nullptr;
  size_t qout_cap = 0, qtr_cap = 0, qmodg_cap = 0, inflow_cap = 0, ev_cap = 0;
  // pinned staging ring for the per-block event list / inflow values, so that run_steps never
  // has to wait for the device (an upload slot is reused only after its copy has executed)
  DevEvent* h_events[2] = {nullptr, nullptr};
  double* h_inflow[2] = {nullptr, nullptr};
  size_t h_ev_cap[2] = {0, 0}, h_inflow_cap[2] = {0, 0};
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
  bool stage_set[2] = {false, false};
  int stage_slot = 0;
};

void routing_free(Routing* rt) {
  if (!rt) return;
  void* ptrs[] = {rt->meta, rt->up_ptr, rt->up_pos, rt->node_lane, rt->cell_ptr, rt->cell_idx,
                  rt->L11_L1_Id, rt->d_inflow_node, rt->d_inflow_index, rt->d_inflow_head,
                  rt->L1_area, rt->L11_area, rt->d_gauge_col, rt->d_gauge_slot, rt->d_cell_entry, rt->d_cell_skew, rt->d_ghost_lane, rt->d_export_lane,
                  rt->d_cell_area, rt->C1, rt->C2, rt->qOUT, rt->qMod, rt->qTIN, rt->qTR,
                  rt->d_length, rt->d_slope, rt->d_fFPimp, rt->carry, rt->gauge_hist, rt->d_xs_tab, rt->d_xr_tab, rt->xsend_buf, rt->xrecv_buf,
                  rt->qout_hist, rt->qtr_hist, rt->qmod_g, rt->d_inflow_val, rt->d_events};
  for (void* p : ptrs) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    cudaFreeHost(rt->h_events[i]);
    cudaFreeHost(rt->h_inflow[i]);
    if (rt->stage_done[i]) cudaEventDestroy(rt->stage_done[i]);
  }
  delete rt;
}

template <class T>
static int upload(T** dst, const std::vector<T>& v, cudaStream_t st) {
  cudaFree(*dst);
  *dst = nullptr;
  const size_t bytes = (v.empty() ? 1 : v.size()) * sizeof(T);
  MHM_CUDA_OK(cudaMalloc(dst, bytes));
  if (!v.empty())
    MHM_CUDA_OK(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  MHM_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

template <class T>
static int ensure(T** p, size_t* cap, size_t need, cudaStream_t st) {
  if (*cap >= need && *p) return 0;
  MHM_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  MHM_CUDA_OK(cudaMalloc(p, (need ? need : 1) * sizeof(T)));
  MHM_CUDA_OK(cudaMemsetAsync(*p, 0, (need ? need : 1) * sizeof(T), st));  // lanes nobody writes read as 0
  *cap = need;
  return 0;
}

// --------------------------------------------------------------------------------- kernels

// reg_rout, mo_mrm_mpr.f90:97-117, one thread per link
__global__ void reg_rout_kernel(int nLinks, double p0, double p1, double p2, double p3, double p4,
                                const double* __restrict__ length, const double* __restrict__ slope,
                                const double* __restrict__ fFPimp, double ssMax, double TS,
                                double* __restrict__ C1, double* __restrict__ C2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nLinks) return;
  double K = p0 + p1 * (length[i] * 0.001) + p2 * slope[i] + p3 * fFPimp[i];
  double xi = p4 * (1.0 + slope[i] / ssMax);
  if (xi > 0.5) xi = 0.5;
  if (xi < 0.005) xi = 0.005;
  if (K > 0.5 * TS / xi) K = 0.5 * TS / xi;
  if (K < 0.5 * TS / (1.0 - xi)) K = 0.5 * TS / (1.0 - xi);
  const double c1 = TS / (K * (1.0 - xi) + 0.5 * TS);
  C1[i] = c1;
  C2[i] = 1.0 - c1 * K / TS;
}

struct QoutArgs {
  int32_t nCells1, nNodes, E, M, nEvents, map_flag, nInflowGauges, nInflowTotal;
  const DevEvent* events;
  const double* runoff_hist;  // [steps][M][nCells1]
  const double* carry;        // [M][nCells1]
  const LaneMeta* meta;
  const int32_t *cell_ptr, *cell_idx, *L11_L1_Id;
  const double *L1_area, *L11_area;
  const int32_t *inflow_node, *inflow_index, *inflow_head;
  const double* inflow_val;  // [nEvents][nInflowTotal]
  double* qout_hist;         // tiled, see hidx()
};

__device__ __forceinline__ double apply_inflow(const QoutArgs& a, int node, int ev, double v) {
  for (int g = 0; g < a.nInflowGauges; ++g) {  // add_inflow, mo_mrm_pre_routing.f90:203-213
    if (a.inflow_node[g] - 1 == node) {
      const double qi = a.inflow_val[(size_t)ev * a.nInflowTotal + a.inflow_index[g] - 1];
      v = a.inflow_head[g] ? v + qi : qi;
    }
  }
  return v;
}

// L11_runoff_acc + add_inflow for every (event, member, lane); a thread produces the 8 events
// of one history tile for its node (one 64-byte run).  General mapping (L11 coarser or finer
// than L1, accumulated runoff).
__global__ void __launch_bounds__(128) qout_kernel(const QoutArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.E) return;
  const int m = blockIdx.y, tile = blockIdx.z;
  const LaneMeta lm = a.meta[p];
  if (!(lm.flags & kEntValid)) return;
  const int node = lm.node;
  const size_t n1 = (size_t)a.nCells1;
  double q[kHistTile];
#pragma unroll
  for (int d = 0; d < kHistTile; ++d) {
    const int ev = tile * kHistTile + d;
    q[d] = 0.0;
    if (ev >= a.nEvents) continue;
    const DevEvent e = a.events[ev];
    auto run_to_rout = [&](int k) {  // RunToRout(k): accumulated in step order
      double acc = e.use_carry ? a.carry[(size_t)m * n1 + k] : 0.0;
      for (int j = 0; j < e.nacc; ++j)
        acc = acc + __ldcs(a.runoff_hist + ((size_t)(e.t0 + j) * a.M + m) * n1 + k);
      return acc;
    };
    double v;
    if (a.map_flag) {  // mo_mrm_pre_routing.f90:112-130
      v = 0.0;
      for (int c = a.cell_ptr[node]; c < a.cell_ptr[node + 1]; ++c) {
        const int k = a.cell_idx[c];
        v = v + run_to_rout(k) * a.L1_area[k];
      }
      v = v * 1000.0 / e.tst;
    } else {  // :132-141
      v = run_to_rout(a.L11_L1_Id[node] - 1);
      v = v * a.L11_area[node] * 1000.0 / e.tst;
    }
    if (lm.flags & kEntInflow) v = apply_inflow(a, node, ev, v);
    q[d] = v;
  }
  const int skew = lane_skew(lm.flags);  // skewed slots: two neighbouring runs of this lane
#pragma unroll
  for (int d = 0; d < kHistTile; ++d)
    if (tile * kHistTile + d < a.nEvents) a.qout_hist[hidx(tile * kHistTile + d + skew, a.M, a.E, m, p)] = q[d];
}

// Same for the common one-cell-per-node case with one model step per event: a thread takes one
// L1 cell, reads its runoff of the 8 events of a tile (coalesced over cells) and writes the
// finished 64-byte run at the lane of the cell's node.
__global__ void __launch_bounds__(128)
qout_cell_kernel(const QoutArgs a, const int32_t* __restrict__ cell_entry,
                 const double* __restrict__ cell_area) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.nCells1) return;
  const int m = blockIdx.y, tile = blockIdx.z;
  const size_t n1 = (size_t)a.nCells1;
  const int p = cell_entry[k];
  const double area = cell_area[k];
  const LaneMeta lm = a.meta[p];
  double q[kHistTile];
#pragma unroll
  for (int d = 0; d < kHistTile; ++d) {
    const int ev = tile * kHistTile + d;
    q[d] = 0.0;
    if (ev >= a.nEvents) continue;
    const DevEvent e = a.events[ev];
    const double r = 0.0 + __ldcs(a.runoff_hist + ((size_t)e.t0 * a.M + m) * n1 + k);
    // map_flag: (0 + qAll*efecArea) * 1000 / TST (:125,:129); else qAll * L11_area * 1000 / TST
    double v = a.map_flag ? (0.0 + r * area) : r * area;
    v = v * 1000.0 / e.tst;
    if (lm.flags & kEntInflow) v = apply_inflow(a, lm.node, ev, v);
    q[d] = v;
  }
  const int skew = lane_skew(lm.flags);
#pragma unroll
  for (int d = 0; d < kHistTile; ++d)
    if (tile * kHistTile + d < a.nEvents) a.qout_hist[hidx(tile * kHistTile + d + skew, a.M, a.E, m, p)] = q[d];
}

// carry = (carry) + sum of the block's not yet routed runoff
__global__ void carry_kernel(int nCells1, int M, int t0, int nacc, int use_carry,
                             const double* __restrict__ runoff_hist, double* __restrict__ carry) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nCells1) return;
  const int m = blockIdx.y;
  double acc = use_carry ? carry[(size_t)m * nCells1 + k] : 0.0;
  for (int j = 0; j < nacc; ++j)
    acc = acc + runoff_hist[((size_t)(t0 + j) * M + m) * nCells1 + k];
  carry[(size_t)m * nCells1 + k] = acc;
}

struct ChainArgs {
  int32_t lane0, lane1;  // lane range of the segment level (multiples of 32)
  int32_t ev0, ev1;      // events of the block routed by this launch (one land-cover scene)
  int32_t rs0;           // first routing sub-step of ev0
  int32_t rl;            // routing sub-steps per event (mo_mrm_routing.f90:224)
  int32_t E, M, nNodes, nGslots, single_node;
  const LaneMeta* meta;
  const int32_t *up_ptr, *up_pos;
  const double *C1, *C2;     // [M][nNodes], link indexed
  const double* qout_hist;   // tiled by event
  double* qtr_hist;          // tiled by routing sub-step
  double *qTIN, *qTR;        // [M][2][nNodes]
  double *qMod, *qOUT;       // [M][nNodes]
  double* qmod_g;            // [nEvents][M][nGslots]
};

// L11_routing (mo_mrm_routing.f90:428-478) for the segments of one level over a block of
// routing steps.  Lane j of a segment works on routing step (kWin * S + d - j) in sub-step d of
// macro step S, so that lane j-1 finished the same routing step one sub-step earlier and its
// outflow arrives by __shfl_up.  Every lane reads its own kWin-step windows (runoff, tributary
// outflows) with static register indexing.
// RL1: one routing step per event (the usual case) -- no index divisions in the inner loops.
#ifndef MHM_CHAIN_WIN
#define MHM_CHAIN_WIN 4
#endif
constexpr int kWin = MHM_CHAIN_WIN;  // routing steps per macro step (register windows of every lane)
#ifndef MHM_CHAIN_MIN_BLOCKS
#define MHM_CHAIN_MIN_BLOCKS 4
#endif
// KU: upstream-link slots any lane of the launch uses (1..kMetaUps); MEM: some lane reads a
// tributary's outflow series from memory (false for the headwater level: no windows at all)
template <bool RL1, int KU, bool MEM>
__global__ void __launch_bounds__(128, MHM_CHAIN_MIN_BLOCKS) route_chain_kernel(const ChainArgs a) {
  const int p = a.lane0 + blockIdx.x * blockDim.x + threadIdx.x;  // lane0, blockDim: multiples of 32
  if (p >= a.lane1) return;                                        // whole warps drop out together
  const int m = blockIdx.y;
  const LaneMeta lm = a.meta[p];
  const bool valid = lm.flags & kEntValid, is_link = lm.flags & kEntLink;
  const bool add_qout = lm.flags & kEntAddQout, zero_out = lm.flags & kEntZeroOut;
  const bool write_hist = lm.flags & kEntWriteHist, ghost = lm.flags & kEntGhost;
  const int nup = (lm.flags >> 8) & 0xff, skew = lane_skew(lm.flags);
  const int ws = lane_wshift(lm.flags);  // shift of this lane's own outflow series
  const int ts = skew & 7;               // shift of the series this lane reads (it is their reader)
  const int rl = RL1 ? 1 : a.rl;
  const int nRS = (a.ev1 - a.ev0) * rl;  // routing sub-steps of this launch
  const int lmax = __reduce_max_sync(0xffffffffu, valid ? skew + 1 : 0);
  double c1 = 0.0, c2 = 0.0;
  if (is_link) {
    c1 = a.C1[(size_t)m * a.nNodes + lm.link];
    c2 = a.C2[(size_t)m * a.nNodes + lm.link];
  }
  double* tin = a.qTIN + (size_t)m * 2 * a.nNodes;
  double* tr = a.qTR + (size_t)m * 2 * a.nNodes;
  double qtin1 = 0.0, qtr1 = 0.0, qout = 0.0, acc = 0.0, last_q = 0.0, qmod = 0.0;
  if (valid) {
    qtin1 = tin[lm.node];
    qtr1 = tr[lm.node];
  }
  const double rl_dp = (double)rl;
  const int u0 = nup > kMetaUps ? a.up_ptr[p] : 0;
  const size_t tile_stride = (size_t)a.M * a.E * kHistTile;  // doubles between history tiles
  const size_t lane_off = ((size_t)m * a.E + p) * kHistTile;
  size_t up_off[KU];
#pragma unroll
  for (int u = 0; u < KU; ++u)
    up_off[u] = ((size_t)m * a.E + (lm.up[u] > 0 ? lm.up[u] : 0)) * kHistTile;
  const int nMacro = (nRS + lmax - 1 + kWin - 1) / kWin;
  for (int S = 0; S < nMacro; ++S) {
    const int base = kWin * S - skew;  // routing sub-step (relative to rs0) of sub-step 0
    // ---- this lane's windows: own runoff and the outflow series it reads from memory ----
    double qo[kWin], t[MEM ? KU : 1][kWin];
#pragma unroll
    for (int d = 0; d < kWin; ++d) {
      const int r = base + d;
      const bool in = valid && r >= 0 && r < nRS;
      const int rs = a.rs0 + r;                       // absolute sub-step within the block
      const int ev = RL1 ? a.ev0 + r : a.ev0 + r / rl;  // its event
      const int evs = ev + skew;  // skewed slot of the node-runoff history
      const size_t oq = (size_t)(evs >> 3) * tile_stride + (size_t)(evs & 7);
      const size_t ot = (size_t)((rs + ts) >> 3) * tile_stride + (size_t)((rs + ts) & 7);
      qo[d] = (in && !ghost) ? a.qout_hist[oq + lane_off] : 0.0;
      if (MEM) {
#pragma unroll
        for (int u = 0; u < KU; ++u)
          t[u][d] = (in && u < nup && lm.up[u] != kUpShuffle) ? a.qtr_hist[ot + up_off[u]] : 0.0;
      }
    }
#pragma unroll
    for (int d = 0; d < kWin; ++d) {
      const double from_prev = __shfl_up_sync(0xffffffffu, last_q, 1);
      const int r = base + d;
      if (valid && r >= 0 && r < nRS) {
        const int rs = a.rs0 + r;
        const int ev = RL1 ? a.ev0 + r : a.ev0 + r / rl;
        const int sub = RL1 ? 0 : r % rl;
        qout = qo[d];
        double q_in;
        if (a.single_node) {  // nNodes == 1: L11_Qmod = L11_qOUT (mo_mrm_routing.f90:284)
          q_in = qout;
        } else {
          q_in = 0.0;  // :428, then upstream links in netPerm order :457
#pragma unroll
          for (int u = 0; u < KU; ++u)
            if (u < nup) q_in = q_in + ((!MEM || lm.up[u] == kUpShuffle) ? from_prev : t[u][d]);
          for (int u = kMetaUps; u < nup; ++u)
            q_in = q_in + a.qtr_hist[hidx(rs + ts, a.M, a.E, m, a.up_pos[u0 + u])];
          if (add_qout) q_in = q_in + qout;  // :441 / :466-467
          if (is_link) {
            double q;
            if (ghost) {  // routed by the shard that owns the node, received for the whole block
              q = a.qtr_hist[(size_t)((rs + ws) >> 3) * tile_stride + (size_t)((rs + ws) & 7) + lane_off];
            } else {
              q = qtr1 + c1 * (qtin1 - qtr1) + c2 * (q_in - qtin1);  // :443-445
              if (zero_out) q = 0.0;                                  // :447-452
              if (write_hist)
                a.qtr_hist[(size_t)((rs + ws) >> 3) * tile_stride + (size_t)((rs + ws) & 7) + lane_off] = q;
            }
            qtr1 = q;
            last_q = q;
          }
          qtin1 = q_in;
        }
        // mean over the sub-steps of the event, :263,:281
        if (RL1) {
          qmod = q_in;  // (0 + q) / 1
          if (lm.gslot >= 0) a.qmod_g[((size_t)ev * a.M + m) * a.nGslots + lm.gslot] = qmod;
        } else {
          acc = (sub == 0 ? 0.0 : acc) + q_in;
          if (sub == rl - 1) {
            qmod = a.single_node ? qout : acc / rl_dp;
            if (lm.gslot >= 0) a.qmod_g[((size_t)ev * a.M + m) * a.nGslots + lm.gslot] = qmod;
          }
        }
      }
    }
  }
  if (valid && nRS > 0) {
    if (!a.single_node) {
      tin[lm.node] = qtin1;
      tin[a.nNodes + lm.node] = qtin1;
      if (is_link) {
        tr[lm.node] = qtr1;
        tr[a.nNodes + lm.node] = qtr1;
      }
    }
    a.qMod[(size_t)m * a.nNodes + lm.node] = qmod;
    a.qOUT[(size_t)m * a.nNodes + lm.node] = qout;
  }
}

// Lean form of route_chain_kernel for the levels that need none of its options: one routing
// step per event with history indices rs == ev, more than one node, no ghost sources, no zeroed
// outflows, at most kMetaUps inflowing links per lane (Routing::lvl_plain).  Same operations in
// the same order -> bit-identical.  One iteration works on one history tile in two 4-slot half
// windows: the node runoff and every tributary series read from memory arrive with one 256-bit
// load per row and half window (skewed / reader-shifted storage, see kHistPad), all addresses
// are a per-lane base plus the tile's byte offset, and half windows in which every lane of the
// warp is inside its steps run unguarded.
#ifndef MHM_LEAN_MIN_BLOCKS
#define MHM_LEAN_MIN_BLOCKS 5  // measured on B200, 128-step block: 4 -> 20.3 ms, 5 -> 19.4, 6 -> 19.7 (spills)
#endif
template <int KU, bool MEM, bool PF>
__global__ void __launch_bounds__(128, PF ? 2 : MHM_LEAN_MIN_BLOCKS) route_chain_lean_kernel(const ChainArgs a) {
  const int p = a.lane0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.lane1) return;
  const int m = blockIdx.y;
  const LaneMeta lm = a.meta[p];
  const bool valid = lm.flags & kEntValid, is_link = lm.flags & kEntLink;
  const bool add_qout = lm.flags & kEntAddQout, write_hist = lm.flags & kEntWriteHist;
  const int nup = (lm.flags >> 8) & 0xff, skew = lane_skew(lm.flags), ws = lane_wshift(lm.flags);
  const int nRS = a.ev1 - a.ev0;
  const int lmax = __reduce_max_sync(0xffffffffu, valid ? skew + 1 : 0);
  double c1 = 0.0, c2 = 0.0;
  if (is_link) {
    c1 = a.C1[(size_t)m * a.nNodes + lm.link];
    c2 = a.C2[(size_t)m * a.nNodes + lm.link];
  }
  double* tin = a.qTIN + (size_t)m * 2 * a.nNodes;
  double* tr = a.qTR + (size_t)m * 2 * a.nNodes;
  double qtin1 = 0.0, qtr1 = 0.0, qout = 0.0, qmod = 0.0;
  if (valid) {
    qtin1 = tin[lm.node];
    qtr1 = tr[lm.node];
  }
  const long long tile_bytes = (long long)a.M * a.E * kHistTile * (long long)sizeof(double);
  if (!MEM && (a.ev0 & (kHistTile - 1)) == 0 && __all_sync(0xffffffffu, !valid || (lm.flags & kEntLeafWarp))) {
    // A warp of single-node segments (every level ends with them; on the headwater level these
    // are the leaves of the network, more than a third of all nodes): no lane waits for another
    // one, so every lane streams its own runs -- one 64-byte run of node runoff in, eight
    // Muskingum steps, one 64-byte run of routed outflow out; the next run is requested before
    // the current one is worked on.  Slot = step + skew for both series here (skew == ws), so a
    // lane's runs hold steps r0 .. r0 + 7 with r0 = -skew, 8 - skew, ...
    const size_t tile_elems = (size_t)(tile_bytes / (long long)sizeof(double));
    const size_t row = (size_t)(a.ev0 >> 3) * tile_elems + ((size_t)m * a.E + p) * kHistTile;
    const double* qp = a.qout_hist + row;
    double* hp = a.qtr_hist + row;
    double* const qg1 = lm.gslot >= 0 ? a.qmod_g + ((size_t)a.ev0 * a.M + m) * a.nGslots + lm.gslot : nullptr;
    const bool store = valid && is_link && write_hist;
    double nx[kHistTile];
    auto load_run = [&](const double* src) {
      asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                   : "=d"(nx[0]), "=d"(nx[1]), "=d"(nx[2]), "=d"(nx[3]) : "l"(src));
      asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                   : "=d"(nx[4]), "=d"(nx[5]), "=d"(nx[6]), "=d"(nx[7]) : "l"(src + 4));
    };
    load_run(qp);
    for (int r0 = -skew; r0 < nRS; r0 += kHistTile) {
      double v[kHistTile], o[kHistTile];
#pragma unroll
      for (int d = 0; d < kHistTile; ++d) v[d] = nx[d];
      qp += tile_elems;
      if (r0 + kHistTile < nRS) load_run(qp);
#pragma unroll
      for (int d = 0; d < kHistTile; ++d) {
        o[d] = 0.0;
        if (valid && (unsigned)(r0 + d) < (unsigned)nRS) {
          qout = v[d];
          double q_in = 0.0;                 // a segment head on this level has no inflowing link
          if (add_qout) q_in = q_in + qout;  // :441 / :466-467
          if (is_link) {
            const double q = qtr1 + c1 * (qtin1 - qtr1) + c2 * (q_in - qtin1);  // :443-445
            o[d] = q;
            qtr1 = q;
          }
          qtin1 = q_in;
          qmod = q_in;
          if (qg1) qg1[(size_t)(r0 + d) * (size_t)a.M * a.nGslots] = qmod;
        }
      }
      if (store) {
        if (r0 >= 0 && r0 + kHistTile <= nRS) {
          asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(hp), "d"(o[0]), "d"(o[1]), "d"(o[2]), "d"(o[3]) : "memory");
          asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(hp + 4), "d"(o[4]), "d"(o[5]), "d"(o[6]), "d"(o[7]) : "memory");
        } else {
#pragma unroll
          for (int d = 0; d < kHistTile; ++d)
            if ((unsigned)(r0 + d) < (unsigned)nRS) hp[d] = o[d];
        }
      }
      hp += tile_elems;
    }
    if (valid && nRS > 0) {
      tin[lm.node] = qtin1;
      tin[a.nNodes + lm.node] = qtin1;
      if (is_link) {
        tr[lm.node] = qtr1;
        tr[a.nNodes + lm.node] = qtr1;
      }
      a.qMod[(size_t)m * a.nNodes + lm.node] = qmod;
      a.qOUT[(size_t)m * a.nNodes + lm.node] = qout;
    }
    return;
  }
  // ---- pipelined segments: one iteration = one history tile (8 slots), two half windows ----
  // qout_hist slot of (lane, step r) = ev0 + r + skew: in half h of iteration k every lane works
  // on slots ev0 + 8k + 4h + d, i.e. on step r = 8k + 4h + d - skew (ev0 is a multiple of 8).
  const int tile0 = a.ev0 >> 3;
  const long long row = ((long long)m * a.E + p) * kHistTile * (long long)sizeof(double);
  const char* const qo_b = reinterpret_cast<const char*>(a.qout_hist) + tile0 * tile_bytes + row;
  // own outflow of step r: slot ev0 + r + ws = ev0 + 8(k + dq) + (4h + d + c); it lies in the next
  // tile (w_hi) from window position 8 - c on
  const int dq = (ws - skew) >> 3, c = (ws - skew) & 7;
  char* const w_lo = reinterpret_cast<char*>(a.qtr_hist) + (tile0 + dq) * tile_bytes + row + c * (long long)sizeof(double);
  char* const w_hi = w_lo + tile_bytes - kHistTile * (long long)sizeof(double);
  const int wrap_at = kHistTile - c;
  // tributary u read from memory: written with shift skew & 7 for this reader, so its step r is
  // in slot ev0 + 8(k - skew / 8) + 4h + d: the same window position as the node runoff
  const char* t_b[MEM ? KU : 1];
  bool up_mem[MEM ? KU : 1];
  bool up_any[MEM ? KU : 1];  // some lane of the warp reads slot u from memory (warp-uniform)
  unsigned sh_mask[KU];       // all ones when slot u is the outflow of the previous lane
#pragma unroll
  for (int u = 0; u < KU; ++u) {
    sh_mask[u] = (valid && u < nup && lm.up[u] == kUpShuffle) ? 0xffffffffu : 0u;
    asm volatile("" : "+r"(sh_mask[u]));  // kept as a bit mask: one LOP3 per half word and slot
    if (MEM) {
      up_mem[u] = valid && u < nup && lm.up[u] != kUpShuffle;
      up_any[u] = __any_sync(0xffffffffu, up_mem[u]);
      t_b[u] = reinterpret_cast<const char*>(a.qtr_hist) + (tile0 - (skew >> 3)) * tile_bytes +
               ((long long)m * a.E + (up_mem[u] ? lm.up[u] : p)) * kHistTile * (long long)sizeof(double);
    }
  }
  unsigned addq_mask = add_qout ? 0xffffffffu : 0u;
  asm volatile("" : "+r"(addq_mask));
  const bool warp_gauge = __any_sync(0xffffffffu, lm.gslot >= 0);
  double* const qg = lm.gslot >= 0 ? a.qmod_g + (size_t)m * a.nGslots + lm.gslot : nullptr;
  const size_t qg_stride = (size_t)a.M * a.nGslots;
  const int nRSv = valid ? nRS : 0;
  const int nIter = (nRS + lmax - 1 + kHistTile - 1) / kHistTile;
  static_assert(kWin == 4 && kHistTile == 8, "two 4-slot half windows per history tile");
  struct Window {
    double qo[kWin], t[MEM ? KU : 1][kWin];
  };
  // Slots nobody loads stay +0.0 for the whole launch: the inflow sum below adds every slot
  // unconditionally (x + (+0.0) == x bit for bit for every x the sum can hold: it starts from
  // +0.0 and can never become -0.0).
  Window wa, wb;
#pragma unroll
  for (int d = 0; d < kWin; ++d) {
    wa.qo[d] = wb.qo[d] = 0.0;
#pragma unroll
    for (int u = 0; u < (MEM ? KU : 1); ++u) wa.t[u][d] = wb.t[u][d] = 0.0;
  }
  // loads of half h of the tile at byte offset koff (one 256-bit load per row)
  // (measured on B200 and dropped: prefetch.global.L2 of the runs 1, 2 or 4 tiles ahead, 19.4 ->
  // 23.6 / 19.8 / 20.6 ms per 128-step block; routing the headwater leaves inside their reader
  // instead of writing and re-reading their series, 22.4 ms: the levels are bound by latency
  // and issue, not by DRAM bytes)
  auto load_half = [&](Window& w, const long long koff, const int r0, const int h) {
    // windows wholly outside the lane's steps (pipeline fill and drain) are not fetched
    const bool some = r0 + kWin > 0 && r0 < nRSv;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
                 "@p ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+d"(w.qo[0]), "+d"(w.qo[1]), "+d"(w.qo[2]), "+d"(w.qo[3])
                 : "l"(qo_b + koff + h * 32), "r"((int)some));
    if (MEM) {
      // tributary rows were written by earlier launches; window positions outside the launch's
      // steps hold other steps' values and never reach the state (guard below)
#pragma unroll
      for (int u = 0; u < KU; ++u)
        if (up_any[u])
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
                       "@p ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];\n\t}"
                       : "+d"(w.t[u][0]), "+d"(w.t[u][1]), "+d"(w.t[u][2]), "+d"(w.t[u][3])
                       : "l"(t_b[u] + koff + h * 32), "r"((int)(some && up_mem[u])));
    }
  };
  // the four routing steps of a half window; GUARD: some lane is outside its steps (pipeline
  // fill and drain), otherwise every lane is inside and nothing is predicated but the stores
  auto route_half = [&](const Window& w, const long long koff, const int r0, const int h, auto guard_tag) {
    constexpr bool GUARD = decltype(guard_tag)::value;
#pragma unroll
    for (int d = 0; d < kWin; ++d) {
      const double from_prev = __shfl_up_sync(0xffffffffu, qtr1, 1);
      const unsigned fp_lo = (unsigned)__double2loint(from_prev), fp_hi = (unsigned)__double2hiint(from_prev);
      const bool G = !GUARD || (unsigned)(r0 + d) < (unsigned)nRSv;
      double q_in = 0.0;  // :428, then upstream links in netPerm order :457
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        unsigned lo = fp_lo & sh_mask[u], hi = fp_hi & sh_mask[u];
        if (MEM) {
          lo |= (unsigned)__double2loint(w.t[u][d]);
          hi |= (unsigned)__double2hiint(w.t[u][d]);
        }
        q_in = q_in + __hiloint2double((int)hi, (int)lo);
      }
      q_in = q_in + __hiloint2double((int)((unsigned)__double2hiint(w.qo[d]) & addq_mask),
                                     (int)((unsigned)__double2loint(w.qo[d]) & addq_mask));  // :441 / :466-467
      // :443-445; lanes without a link carry c1 = c2 = 0 and never store or hand on q
      const double q = qtr1 + c1 * (qtin1 - qtr1) + c2 * (q_in - qtin1);
      const int idx = h * kWin + d;
      if (G && write_hist) *reinterpret_cast<double*>((idx >= wrap_at ? w_hi : w_lo) + koff + idx * 8) = q;
      if (warp_gauge) {
        if (G && qg) qg[(size_t)(a.ev0 + r0 + d) * qg_stride] = q_in;  // (0 + q) / 1, :263,:281
      }
      qtr1 = G ? q : qtr1;
      qtin1 = G ? q_in : qtin1;
    }
  };
  auto route = [&](const Window& w, const long long koff, const int r0, const int h, const int base) {
    // base = r0 + skew, the same for all lanes: the lane with the largest skew of the warp is at
    // step base - (lmax - 1), the one with skew 0 at base
    if (base - (lmax - 1) >= 0 && base + kWin - 1 < nRS) route_half(w, koff, r0, h, std::false_type{});
    else route_half(w, koff, r0, h, std::true_type{});
  };
  long long koff = 0;
  int r0 = -skew;
  if (PF) {
    // narrow levels (less than one wave of CTAs): nothing hides the load latency of a half
    // window, so the next one is requested before the current one is routed
    load_half(wa, koff, r0, 0);
    for (int k = 0; k < nIter; ++k, koff += tile_bytes, r0 += kHistTile) {
      load_half(wb, koff, r0 + kWin, 1);
      route(wa, koff, r0, 0, k * kHistTile);
      if (k + 1 < nIter) load_half(wa, koff + tile_bytes, r0 + kHistTile, 0);
      route(wb, koff, r0 + kWin, 1, k * kHistTile + kWin);
    }
  } else {
    for (int k = 0; k < nIter; ++k, koff += tile_bytes, r0 += kHistTile) {
      load_half(wa, koff, r0, 0);
      route(wa, koff, r0, 0, k * kHistTile);
      load_half(wa, koff, r0 + kWin, 1);
      route(wa, koff, r0 + kWin, 1, k * kHistTile + kWin);
    }
  }
  if (valid && nRS > 0) {
    tin[lm.node] = qtin1;
    tin[a.nNodes + lm.node] = qtin1;
    if (is_link) {
      tr[lm.node] = qtr1;
      tr[a.nNodes + lm.node] = qtr1;
    }
    a.qMod[(size_t)m * a.nNodes + lm.node] = qtin1;
    a.qOUT[(size_t)m * a.nNodes + lm.node] = a.qout_hist[hidx(a.ev0 + nRS - 1 + skew, a.M, a.E, m, p)];
  }
}

// mRM_runoff(tt, gaugeIndexList(gg)) = L11_Qmod(gaugeNodeList(gg)), plus the back-fill of
// mo_mhm_interface_run.f90:600-603
__global__ void gauge_kernel(int nEvents, int M, int nGauges, int nGslots, int nTimeSteps,
                             int nGaugesTotal, const DevEvent* __restrict__ events,
                             const int32_t* __restrict__ gauge_col,
                             const int32_t* __restrict__ gauge_slot,
                             const double* __restrict__ qmod_g, double* __restrict__ gauge_hist) {
  const int ev = blockIdx.x * blockDim.x + threadIdx.x;
  if (ev >= nEvents) return;
  const int m = blockIdx.y;
  const DevEvent e = events[ev];
  double* H = gauge_hist + (size_t)m * nGaugesTotal * nTimeSteps;
  for (int g = 0; g < nGauges; ++g)
    H[(size_t)gauge_col[g] * nTimeSteps + (e.tt - 1)] =
        qmod_g[((size_t)ev * M + m) * nGslots + gauge_slot[g]];
  if (e.backfill > 0)
    for (int c = 0; c < nGaugesTotal; ++c)
      for (int jj = 1; jj <= e.backfill; ++jj)
        if (e.tt - jj >= 0) H[(size_t)c * nTimeSteps + (e.tt - jj)] = H[(size_t)c * nTimeSteps + (e.tt - 1)];
}

// ------------------------------------------------------------------------------ host side

// L11_routing_order (mo_mrm_net_startup.f90:765-842) in O(nLinks).
// Headwater links (no link drains into their from-node) take ranks 1..nH in link order.
// The reference then sweeps the unranked links in ascending index again and again; a link
// is ranked as soon as every link draining into its from-node is ranked, also by the same
// sweep (the test reads rOrder live).  Hence sweep(i) = max(1, max_j(sweep(j) + (j > i)))
// over upstream links j, and rOrder is the rank of (sweep, i) in lexicographic order.
static int routing_order_linear(int32_t nNodes, int32_t nLinks, const int32_t* fromN,
                                const int32_t* toN, int32_t* rOrder, int32_t* netPerm) {
  std::vector<int32_t> link_of_node((size_t)nNodes + 1, -1), indeg((size_t)nLinks, 0),
      sweep((size_t)nLinks, 0);
  for (int i = 0; i < nLinks; ++i) {
    if (fromN[i] < 1 || fromN[i] > nNodes || toN[i] < 1 || toN[i] > nNodes) {
      set_error("routing_order: link %d has nodes outside 1..%d", i + 1, nNodes);
      return 1;
    }
    link_of_node[(size_t)fromN[i]] = i;
  }
  // number of upstream links of each link
  for (int j = 0; j < nLinks; ++j) {
    const int d = link_of_node[(size_t)toN[j]];
    if (d >= 0 && d != j) indeg[(size_t)d]++;
  }
  std::vector<int32_t> queue;
  queue.reserve((size_t)nLinks);
  for (int i = 0; i < nLinks; ++i)
    if (indeg[(size_t)i] == 0) queue.push_back(i);
  size_t done = 0;
  for (size_t q = 0; q < queue.size(); ++q) {
    const int j = queue[q];
    ++done;
    const int d = link_of_node[(size_t)toN[j]];
    if (d < 0 || d == j) continue;
    const int cand = sweep[(size_t)j] + (j > d ? 1 : 0);
    const int c1 = cand < 1 ? 1 : cand;
    if (c1 > sweep[(size_t)d]) sweep[(size_t)d] = c1;
    if (--indeg[(size_t)d] == 0) queue.push_back(d);
  }
  if (done != (size_t)nLinks) {
    set_error("routing_order: the link graph has a cycle");
    return 1;
  }
  // rank of (sweep, i): counting sort by sweep, stable in i
  int32_t smax = 0;
  for (int i = 0; i < nLinks; ++i) smax = std::max(smax, sweep[(size_t)i]);
  std::vector<int32_t> start((size_t)smax + 2, 0);
  for (int i = 0; i < nLinks; ++i) start[(size_t)sweep[(size_t)i] + 1]++;
  for (int s = 0; s <= smax; ++s) start[(size_t)s + 1] += start[(size_t)s];
  for (int i = 0; i < nLinks; ++i) {
    const int r = start[(size_t)sweep[(size_t)i]]++;
    rOrder[i] = r + 1;
    netPerm[r] = i + 1;
  }
  return 0;
}

static int build_topology(mhm_cuda_context* ctx, Domain* d, Routing* rt, const mrm_network* net) {
  const int nNodes = rt->nNodes, nLinks = rt->nLinks;
  std::vector<int32_t> rank((size_t)nLinks), link_of_node((size_t)nNodes, -1),
      down((size_t)nNodes, -1);
  for (int k = 0; k < nLinks; ++k) {
    const int i = net->netPerm[k] - 1;
    MHM_REQUIRE(i >= 0 && i < nLinks, "set_network: netPerm(%d) = %d outside 1..%d", k + 1, i + 1,
                nLinks);
    rank[(size_t)i] = k;
  }
  for (int i = 0; i < nLinks; ++i) {
    MHM_REQUIRE(net->fromN[i] >= 1 && net->fromN[i] <= nNodes && net->toN[i] >= 1 &&
                    net->toN[i] <= nNodes,
                "set_network: link %d has nodes outside 1..%d", i + 1, nNodes);
    link_of_node[(size_t)net->fromN[i] - 1] = i;
    down[(size_t)net->fromN[i] - 1] = net->toN[i] - 1;
  }
  // upstream links of every node, in netPerm order
  std::vector<std::vector<int32_t>> up((size_t)nNodes);
  for (int k = 0; k < nLinks; ++k) {
    const int i = net->netPerm[k] - 1;
    up[(size_t)net->toN[i] - 1].push_back(i);
  }
  // height of every node = longest chain of links above it; netPerm is a topological order
  std::vector<int32_t> height((size_t)nNodes, 0);
  for (int k = 0; k < nLinks; ++k) {
    const int i = net->netPerm[k] - 1;
    const int f = net->fromN[i] - 1, t = net->toN[i] - 1;
    for (int j : up[(size_t)f])
      MHM_REQUIRE(rank[(size_t)j] < k, "set_network: netPerm is not a topological order");
    height[(size_t)t] = std::max(height[(size_t)t], height[(size_t)f] + 1);
  }
  // ---- chains: every node continues the chain of its highest upstream node ----
  std::vector<int32_t> pred((size_t)nNodes, -1);
  for (int nd = 0; nd < nNodes; ++nd) {
    int best = -1, cnt = 0;
    for (int j : up[(size_t)nd]) {
      if (cnt++ == kMetaUps) break;  // the chain predecessor must be described in LaneMeta::up
      const int f = net->fromN[j] - 1;
      if (best < 0 || height[(size_t)f] > height[(size_t)best]) best = f;
    }
    pred[(size_t)nd] = best;
  }
  // segments of at most kSegLen nodes; seg_of[node], pos_in_seg[node]
  std::vector<int32_t> seg_of((size_t)nNodes, -1), pos_in((size_t)nNodes, 0);
  std::vector<std::vector<int32_t>> seg_nodes;
  for (int nd = 0; nd < nNodes; ++nd) {
    if (pred[(size_t)nd] >= 0) continue;  // not a chain head
    int x = nd, len = 0;
    while (true) {
      if (len % kSegLen == 0) seg_nodes.emplace_back();
      seg_of[(size_t)x] = (int32_t)seg_nodes.size() - 1;
      pos_in[(size_t)x] = len % kSegLen;
      seg_nodes.back().push_back(x);
      ++len;
      const int dn = down[(size_t)x];
      if (dn < 0 || pred[(size_t)dn] != x) break;
      x = dn;
    }
  }
  const int nSeg = (int)seg_nodes.size();
  // segment levels: a segment runs after every segment it reads from memory.  Nodes sorted by
  // height visit all inputs of a segment before the segments that depend on it.
  std::vector<int32_t> by_height((size_t)nNodes);
  for (int nd = 0; nd < nNodes; ++nd) by_height[(size_t)nd] = nd;
  std::stable_sort(by_height.begin(), by_height.end(),
                   [&](int x, int y) { return height[(size_t)x] < height[(size_t)y]; });
  std::vector<int32_t> seg_level((size_t)nSeg, 0);
  for (int nd : by_height) {
    const int s = seg_of[(size_t)nd];
    MHM_REQUIRE(s >= 0, "set_network: node %d is on no chain (cycle?)", nd + 1);
    for (int j : up[(size_t)nd]) {
      const int su = seg_of[(size_t)net->fromN[j] - 1];
      if (su != s) seg_level[(size_t)s] = std::max(seg_level[(size_t)s], seg_level[(size_t)su] + 1);
    }
  }
  // lanes: per level, segments by decreasing length, packed into warps (never straddling one)
  int nLv = 0;
  for (int s = 0; s < nSeg; ++s) nLv = std::max(nLv, seg_level[(size_t)s] + 1);
  std::vector<std::vector<int32_t>> lv_segs((size_t)nLv);
  for (int s = 0; s < nSeg; ++s) lv_segs[(size_t)seg_level[(size_t)s]].push_back(s);
  std::vector<int32_t> lane_of((size_t)nNodes, -1);
  rt->lvl_ptr.assign(1, 0);
  int lane = 0;
  // highest tributary slot (of the kMetaUps described in LaneMeta) a segment reads from memory:
  // the lean kernel skips, per warp, the slots none of its lanes reads, so segments with the
  // same need share warps
  std::vector<int32_t> seg_slots((size_t)nSeg, 0);
  for (int nd = 0; nd < nNodes; ++nd) {
    const int s = seg_of[(size_t)nd];
    if (s < 0) continue;
    int u = 0;
    for (int j : up[(size_t)nd]) {
      if (u == kMetaUps) break;
      const int f = net->fromN[j] - 1;
      const bool shuffle = f == pred[(size_t)nd] && seg_of[(size_t)f] == s;
      if (!shuffle) seg_slots[(size_t)s] = std::max(seg_slots[(size_t)s], u + 1);
      ++u;
    }
  }
  for (int l = 0; l < nLv; ++l) {
    auto& v = lv_segs[(size_t)l];
    std::stable_sort(v.begin(), v.end(), [&](int x, int y) {
      if (seg_nodes[(size_t)x].size() != seg_nodes[(size_t)y].size())
        return seg_nodes[(size_t)x].size() > seg_nodes[(size_t)y].size();
      return seg_slots[(size_t)x] > seg_slots[(size_t)y];
    });
    for (int s : v) {
      const int len = (int)seg_nodes[(size_t)s].size();
      if (lane % kSegLen + len > kSegLen) lane = (lane / kSegLen + 1) * kSegLen;
      for (int x : seg_nodes[(size_t)s]) lane_of[(size_t)x] = lane++;
    }
    lane = (lane + kSegLen - 1) / kSegLen * kSegLen;
    rt->lvl_ptr.push_back(lane);
  }
  const int E = lane;
  rt->E = E;
  // write shift of every node's outflow series (see kHistPad) and the warps of headwater
  // single-node segments, whose position field takes the shift
  std::vector<int32_t> wshift((size_t)nNodes, 0);
  for (int nd = 0; nd < nNodes; ++nd) {
    const int dn = down[(size_t)nd];
    if (dn < 0) continue;
    const bool by_shuffle = pred[(size_t)dn] == nd && seg_of[(size_t)dn] == seg_of[(size_t)nd];
    if (!by_shuffle) wshift[(size_t)nd] = pos_in[(size_t)dn] & 7;
  }
  std::vector<char> leaf_warp((size_t)E / kSegLen, 1);
  for (int nd = 0; nd < nNodes; ++nd)
    if (seg_nodes[(size_t)seg_of[(size_t)nd]].size() != 1 || !up[(size_t)nd].empty())
      leaf_warp[(size_t)lane_of[(size_t)nd] / kSegLen] = 0;
  for (int nd = 0; nd < nNodes; ++nd)
    if (leaf_warp[(size_t)lane_of[(size_t)nd] / kSegLen]) pos_in[(size_t)nd] = wshift[(size_t)nd];

  std::vector<char> is_ghost((size_t)nNodes, 0), is_export((size_t)nNodes, 0);
  for (int g = 0; g < net->nGhostSources; ++g) {
    MHM_REQUIRE(net->ghostSourceNodeList && net->ghostSourceNodeList[g] >= 1 && net->ghostSourceNodeList[g] <= nNodes,
                "set_network: ghostSourceNodeList(%d) outside 1..%d", g + 1, nNodes);
    is_ghost[(size_t)net->ghostSourceNodeList[g] - 1] = 1;
  }
  for (int g = 0; g < net->nExports; ++g) {
    MHM_REQUIRE(net->exportNodeList && net->exportNodeList[g] >= 1 && net->exportNodeList[g] <= nNodes,
                "set_network: exportNodeList(%d) outside 1..%d", g + 1, nNodes);
    is_export[(size_t)net->exportNodeList[g] - 1] = 1;
  }
  int last_sink = nLinks > 0 ? net->toN[net->netPerm[nLinks - 1] - 1] - 1 : -1;
  if (net->lastSinkNode != 0) last_sink = net->lastSinkNode > 0 ? net->lastSinkNode - 1 : -1;
  std::vector<LaneMeta> meta((size_t)E);
  std::memset(meta.data(), 0, meta.size() * sizeof(LaneMeta));
  std::vector<int32_t> up_ptr((size_t)E + 1, 0), up_pos;
  std::vector<int32_t> node_at((size_t)E, -1);
  for (int nd = 0; nd < nNodes; ++nd) node_at[(size_t)lane_of[(size_t)nd]] = nd;
  for (int p = 0; p < E; ++p) {
    LaneMeta& lm = meta[(size_t)p];
    lm.gslot = -1;
    const int nd = node_at[(size_t)p];
    if (nd < 0) {
      up_ptr[(size_t)p + 1] = (int32_t)up_pos.size();
      continue;
    }
    const int l = link_of_node[(size_t)nd];
    int fl = kEntValid | (leaf_warp[(size_t)p / kSegLen] ? kEntLeafWarp : 0);
    if (l >= 0) {
      fl |= kEntLink | kEntAddQout;                 // mo_mrm_routing.f90:441
      for (int g = 0; g < net->nInflowGauges; ++g)  // :447-452
        if (net->toN[l] == net->InflowGaugeNodeList[g] && !net->InflowGaugeHeadwater[g])
          fl |= kEntZeroOut;
      // the outflow series goes to memory unless the only reader is the next lane
      const int dn = down[(size_t)nd];
      const bool by_shuffle = pred[(size_t)dn] == nd && seg_of[(size_t)dn] == seg_of[(size_t)nd];
      if (!by_shuffle) fl |= kEntWriteHist;
    } else if (nd == last_sink) {
      fl |= kEntAddQout;  // :466-467: only the last link's sink adds its own runoff
    }
    for (int g = 0; g < net->nInflowGauges; ++g)
      if (net->InflowGaugeNodeList[g] - 1 == nd) fl |= kEntInflow;
    if (is_ghost[(size_t)nd]) {
      MHM_REQUIRE(l >= 0 && up[(size_t)nd].empty(), "set_network: ghost source %d must be a headwater with a link", nd + 1);
      fl = (fl | kEntGhost) & ~(kEntAddQout | kEntZeroOut | kEntWriteHist);
    }
    if (is_export[(size_t)nd]) {
      MHM_REQUIRE(l >= 0, "set_network: export node %d has no link", nd + 1);
      fl |= kEntWriteHist;
    }
    const int nup = (int)up[(size_t)nd].size();
    MHM_REQUIRE(nup < 256, "set_network: node %d has %d inflowing links", nd + 1, nup);
    lm.node = nd;
    lm.link = l >= 0 ? l : 0;
    lm.flags = fl | (nup << 8) | (pos_in[(size_t)nd] << 16) | (wshift[(size_t)nd] << 21);
    for (int u = 0; u < nup; ++u) {
      const int f = net->fromN[up[(size_t)nd][(size_t)u]] - 1;
      const bool shuffle = f == pred[(size_t)nd] && seg_of[(size_t)f] == seg_of[(size_t)nd];
      const int src = shuffle ? kUpShuffle : lane_of[(size_t)f];
      if (u < kMetaUps) lm.up[u] = src;
      MHM_REQUIRE(u < kMetaUps || !shuffle,
                  "set_network: node %d: chain predecessor beyond the %d-th inflowing link", nd + 1,
                  kMetaUps);
      up_pos.push_back(src);
    }
    up_ptr[(size_t)p + 1] = (int32_t)up_pos.size();
  }
  // per level: how many of the kMetaUps slots are in use, and whether any is read from memory
  rt->lvl_ku.assign(rt->lvl_ptr.size() - 1, 1);
  rt->lvl_mem.assign(rt->lvl_ptr.size() - 1, 0);
  rt->lvl_plain.assign(rt->lvl_ptr.size() - 1, 1);
  for (size_t l = 0; l + 1 < rt->lvl_ptr.size(); ++l)
    for (int p = rt->lvl_ptr[l]; p < rt->lvl_ptr[l + 1]; ++p) {
      const LaneMeta& lm = meta[(size_t)p];
      if (!(lm.flags & kEntValid)) continue;
      if ((lm.flags & (kEntGhost | kEntZeroOut)) || ((lm.flags >> 8) & 0xff) > kMetaUps) rt->lvl_plain[l] = 0;
      const int nup = std::min((lm.flags >> 8) & 0xff, kMetaUps);
      rt->lvl_ku[l] = std::max(rt->lvl_ku[l], nup);
      for (int u = 0; u < nup; ++u)
        if (lm.up[u] != kUpShuffle) rt->lvl_mem[l] = 1;
      if (((lm.flags >> 8) & 0xff) > kMetaUps) rt->lvl_mem[l] = 1;
    }
  // gauge slots: distinct gauge nodes
  std::vector<int32_t> gcol((size_t)net->nGauges), gslot((size_t)net->nGauges);
  rt->nGslots = 0;
  for (int g = 0; g < net->nGauges; ++g) {
    const int nd = net->gaugeNodeList[g] - 1;
    MHM_REQUIRE(nd >= 0 && nd < nNodes, "set_network: gauge node %d outside 1..%d", nd + 1, nNodes);
    MHM_REQUIRE(net->gaugeIndexList[g] >= 1 && net->gaugeIndexList[g] <= net->nGaugesTotal,
                "set_network: gaugeIndexList(%d) outside 1..nGaugesTotal", g + 1);
    int32_t& s = meta[(size_t)lane_of[(size_t)nd]].gslot;
    if (s < 0) s = rt->nGslots++;
    gslot[(size_t)g] = s;
    gcol[(size_t)g] = net->gaugeIndexList[g] - 1;
  }
  cudaStream_t st = ctx->stream;
  if (int rc = upload(&rt->meta, meta, st)) return rc;
  if (int rc = upload(&rt->up_ptr, up_ptr, st)) return rc;
  if (int rc = upload(&rt->up_pos, up_pos, st)) return rc;
  if (int rc = upload(&rt->node_lane, lane_of, st)) return rc;
  if (int rc = upload(&rt->d_gauge_col, gcol, st)) return rc;
  if (int rc = upload(&rt->d_gauge_slot, gslot, st)) return rc;
  {
    std::vector<int32_t> gl((size_t)net->nGhostSources), el((size_t)net->nExports);
    for (int g = 0; g < net->nGhostSources; ++g) gl[(size_t)g] = lane_of[(size_t)net->ghostSourceNodeList[g] - 1];
    for (int g = 0; g < net->nExports; ++g) el[(size_t)g] = lane_of[(size_t)net->exportNodeList[g] - 1];
    rt->nGhost = net->nGhostSources;
    rt->nExport = net->nExports;
    if (int rc = upload(&rt->d_ghost_lane, gl, st)) return rc;
    if (int rc = upload(&rt->d_export_lane, el, st)) return rc;
  }

  // L1 <-> L11 mapping
  const int n1 = d->cfg.nCells;
  if (rt->map_flag) {
    std::vector<int32_t> cptr((size_t)nNodes + 1, 0), cidx((size_t)n1);
    for (int k = 0; k < n1; ++k) {
      MHM_REQUIRE(net->L1_L11_Id[k] >= 1 && net->L1_L11_Id[k] <= nNodes,
                  "set_network: L1_L11_Id(%d) outside 1..%d", k + 1, nNodes);
      cptr[(size_t)net->L1_L11_Id[k]]++;
    }
    for (int nd = 0; nd < nNodes; ++nd) cptr[(size_t)nd + 1] += cptr[(size_t)nd];
    std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1);
    for (int k = 0; k < n1; ++k) cidx[(size_t)fill[(size_t)net->L1_L11_Id[k] - 1]++] = k;
    if (int rc = upload(&rt->cell_ptr, cptr, st)) return rc;
    if (int rc = upload(&rt->cell_idx, cidx, st)) return rc;
  } else {
    std::vector<int32_t> v(net->L11_L1_Id, net->L11_L1_Id + nNodes);
    for (int nd = 0; nd < nNodes; ++nd)
      MHM_REQUIRE(v[(size_t)nd] >= 1 && v[(size_t)nd] <= n1, "set_network: L11_L1_Id outside 1..%d", n1);
    if (int rc = upload(&rt->L11_L1_Id, v, st)) return rc;
  }
  // one-to-one mapping between L1 cells and L11 nodes?
  rt->bijective = false;
  const bool sharded = net->nGhostSources > 0 || net->nExports > 0;
  MHM_REQUIRE(!sharded || rt->map_flag, "set_network: sharded domains need map_flag (L11 >= L1)");
  if (n1 == nNodes || (sharded && n1 < nNodes)) {
    std::vector<int32_t> node_of_cell((size_t)n1, -1);
    std::vector<char> seen((size_t)nNodes, 0);
    bool ok = true;
    if (rt->map_flag) {
      for (int k = 0; k < n1 && ok; ++k) {
        const int nd = net->L1_L11_Id[k] - 1;
        ok = !seen[(size_t)nd];
        seen[(size_t)nd] = 1;
        node_of_cell[(size_t)k] = nd;
      }
    } else {
      for (int nd = 0; nd < nNodes && ok; ++nd) {
        const int k = net->L11_L1_Id[nd] - 1;
        ok = node_of_cell[(size_t)k] < 0;
        node_of_cell[(size_t)k] = nd;
      }
    }
    if (ok) {
      std::vector<int32_t> ce((size_t)n1);
      std::vector<int8_t> cs((size_t)n1);
      std::vector<double> ca((size_t)n1);
      for (int k = 0; k < n1; ++k) {
        const int nd = node_of_cell[(size_t)k];
        MHM_REQUIRE(!is_ghost[(size_t)nd], "set_network: L1 cell %d maps to a ghost node", k + 1);
        ce[(size_t)k] = lane_of[(size_t)nd];
        cs[(size_t)k] = (int8_t)lane_skew(meta[(size_t)lane_of[(size_t)nd]].flags);
        ca[(size_t)k] = rt->map_flag ? net->L1_areaCell[k] : net->L11_areaCell[nd];
      }
      if (int rc = upload(&rt->d_cell_entry, ce, st)) return rc;
      if (int rc = upload(&rt->d_cell_skew, cs, st)) return rc;
      if (int rc = upload(&rt->d_cell_area, ca, st)) return rc;
      rt->bijective = true;
    }
  }
  std::vector<double> a1(net->L1_areaCell, net->L1_areaCell + n1),
      a11(net->L11_areaCell, net->L11_areaCell + nNodes);
  if (int rc = upload(&rt->L1_area, a1, st)) return rc;
  if (int rc = upload(&rt->L11_area, a11, st)) return rc;
  rt->inflowIndexList.assign(net->InflowGaugeIndexList, net->InflowGaugeIndexList + net->nInflowGauges);
  rt->inflowHeadwater.assign(net->InflowGaugeHeadwater, net->InflowGaugeHeadwater + net->nInflowGauges);
  rt->inflowNodeList.assign(net->InflowGaugeNodeList, net->InflowGaugeNodeList + net->nInflowGauges);
  if (int rc = upload(&rt->d_inflow_node, rt->inflowNodeList, st)) return rc;
  if (int rc = upload(&rt->d_inflow_index, rt->inflowIndexList, st)) return rc;
  if (int rc = upload(&rt->d_inflow_head, rt->inflowHeadwater, st)) return rc;
  return 0;
}

static int alloc_zero(double** p, size_t n, cudaStream_t st) {
  MHM_CUDA_OK(cudaMalloc(p, (n ? n : 1) * sizeof(double)));
  MHM_CUDA_OK(cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(double), st));
  return 0;
}

// make sure every member's C1/C2 belong to land-cover scene yId (case 1: reg_rout is
// re-evaluated by the reference on every call with the current scene's fFPimp)
static int ensure_c1c2(mhm_cuda_context* ctx, const Domain* d, Routing* rt, int yId, double timestep_rout) {
  if (rt->rout_case != 1 || rt->nNodes <= 1) return 0;
  // restart run: the reference skips reg_rout (mo_mrm_routing.f90:211 `.not. read_states`) and
  // routes with the restart file's L11_C1 / L11_C2 -- they must have come through mrm_cuda_set_c1c2
  // (or mrm_cuda_set_state), whatever mrm_cuda_set_reg_rout was given
  if (d->cfg.read_states) return 0;
  for (int m = 0; m < rt->M; ++m) {
    if (rt->param5[(size_t)m].empty()) continue;  // C1/C2 supplied
    if (rt->c1c2_yId[(size_t)m] == yId) continue;
    const std::vector<double>& g = rt->param5[(size_t)m];
    const int nl = rt->nLinks;
    reg_rout_kernel<<<(nl + 127) / 128, 128, 0, ctx->stream>>>(
        nl, g[0], g[1], g[2], g[3], g[4], rt->d_length, rt->d_slope,
        rt->d_fFPimp + ((size_t)m * rt->nLC + (yId - 1)) * rt->nNodes, rt->ssMax, timestep_rout,
        rt->C1 + (size_t)m * rt->nNodes, rt->C2 + (size_t)m * rt->nNodes);
    MHM_CUDA_OK(cudaGetLastError());
    rt->c1c2_yId[(size_t)m] = yId;
  }
  return 0;
}

static long fortran_nint(double x) { return (long)std::lround(x); }

static bool routing_accumulates(const Domain* d, const Routing* rt) {
  return rt->rout_case != 1 && rt->TSrout / (d->cfg.timestep_h * 3600.0) >= 1.0;
}
static int routing_rout_loop(const Domain* d, const Routing* rt) {
  if (rt->rout_case == 1) return 1;
  const double f = rt->TSrout / (d->cfg.timestep_h * 3600.0);
  const long rl = fortran_nint(1.0 / f);
  return (int)(rl < 1 ? 1 : rl);
}

struct Segment {
  int32_t ev0, ev1, yId;
};

// route the events of one block.  `per_cell`: one cell per node and one model step per event,
// so the qOUT tiles are built by the coalesced per-cell kernel.
static int run_events(mhm_cuda_context* ctx, Domain* d, Routing* rt, std::vector<DevEvent>& ev,
                      const std::vector<Segment>& segs, const std::vector<double>& inflow_val,
                      const double* runoff_hist, bool per_cell, bool qout_ready,
                      double timestep_rout) {
  if (ev.empty()) return 0;
  cudaStream_t st = ctx->stream;
  const int nEv = (int)ev.size(), M = rt->M, E = rt->E;
  const int rl = ev[0].rout_loop;
  int RS = 0;
  for (auto& e : ev) {
    MHM_REQUIRE(e.rout_loop == rl, "routing: sub-step count changes inside a block (%d vs %d)",
                e.rout_loop, rl);
    e.rs_first = RS;
    RS += e.rout_loop;
  }
  if (int rc = ensure(&rt->d_events, &rt->ev_cap, (size_t)nEv, st)) return rc;
  if (int rc = ensure(&rt->qout_hist, &rt->qout_cap, hist_size(nEv + kHistPad, M, E), st)) return rc;
  if (int rc = ensure(&rt->qtr_hist, &rt->qtr_cap, hist_size(RS + kHistPad, M, E), st)) return rc;
  if (int rc = ensure(&rt->qmod_g, &rt->qmodg_cap, (size_t)nEv * M * std::max(1, rt->nGslots), st))
    return rc;
  if (int rc = ensure(&rt->d_inflow_val, &rt->inflow_cap, std::max<size_t>(1, inflow_val.size()), st))
    return rc;
  {  // stage through pinned memory; no host-device synchronisation on the hot path
    const int sl = rt->stage_slot;
    rt->stage_slot ^= 1;
    if (rt->stage_set[sl]) MHM_CUDA_OK(cudaEventSynchronize(rt->stage_done[sl]));
    if (rt->h_ev_cap[sl] < (size_t)nEv) {
      cudaFreeHost(rt->h_events[sl]);
      rt->h_events[sl] = nullptr;
      MHM_CUDA_OK(cudaHostAlloc(&rt->h_events[sl], (size_t)nEv * sizeof(DevEvent), cudaHostAllocDefault));
      rt->h_ev_cap[sl] = (size_t)nEv;
    }
    if (rt->h_inflow_cap[sl] < inflow_val.size()) {
      cudaFreeHost(rt->h_inflow[sl]);
      rt->h_inflow[sl] = nullptr;
      MHM_CUDA_OK(cudaHostAlloc(&rt->h_inflow[sl], inflow_val.size() * sizeof(double), cudaHostAllocDefault));
      rt->h_inflow_cap[sl] = inflow_val.size();
    }
    std::memcpy(rt->h_events[sl], ev.data(), (size_t)nEv * sizeof(DevEvent));
    MHM_CUDA_OK(cudaMemcpyAsync(rt->d_events, rt->h_events[sl], (size_t)nEv * sizeof(DevEvent),
                                cudaMemcpyHostToDevice, st));
    if (!inflow_val.empty()) {
      std::memcpy(rt->h_inflow[sl], inflow_val.data(), inflow_val.size() * sizeof(double));
      MHM_CUDA_OK(cudaMemcpyAsync(rt->d_inflow_val, rt->h_inflow[sl], inflow_val.size() * sizeof(double),
                                  cudaMemcpyHostToDevice, st));
    }
    if (!rt->stage_done[sl]) MHM_CUDA_OK(cudaEventCreateWithFlags(&rt->stage_done[sl], cudaEventDisableTiming));
    MHM_CUDA_OK(cudaEventRecord(rt->stage_done[sl], st));
    rt->stage_set[sl] = true;
  }

  ctx->stat_begin(kStatRouting);
  int64_t launched = 0;
  if (!qout_ready) {
    QoutArgs qa{};
    qa.nCells1 = rt->nCells1;
    qa.nNodes = rt->nNodes;
    qa.E = E;
    qa.M = M;
    qa.map_flag = rt->map_flag;
    qa.nInflowGauges = rt->nInflowGauges;
    qa.nInflowTotal = rt->nInflowTotal;
    qa.runoff_hist = runoff_hist;
    qa.carry = rt->carry;
    qa.meta = rt->meta;
    qa.cell_ptr = rt->cell_ptr;
    qa.cell_idx = rt->cell_idx;
    qa.L11_L1_Id = rt->L11_L1_Id;
    qa.L1_area = rt->L1_area;
    qa.L11_area = rt->L11_area;
    qa.inflow_node = rt->d_inflow_node;
    qa.inflow_index = rt->d_inflow_index;
    qa.inflow_head = rt->d_inflow_head;
    const int tiles = (nEv + kHistTile - 1) / kHistTile;
    for (int t0 = 0; t0 < tiles; t0 += 32768) {  // gridDim.z limit
      // the kernels index events and history tiles from `tile`; shift all three views
      qa.events = rt->d_events + (size_t)t0 * kHistTile;
      qa.nEvents = nEv - t0 * kHistTile;
      qa.inflow_val = rt->d_inflow_val + (size_t)t0 * kHistTile * rt->nInflowTotal;
      qa.qout_hist = rt->qout_hist + hidx(t0 * kHistTile, M, E, 0, 0);
      const int nt = std::min(32768, tiles - t0);
      if (per_cell)
        qout_cell_kernel<<<dim3((rt->nCells1 + 127) / 128, M, nt), 128, 0, st>>>(
            qa, rt->d_cell_entry, rt->d_cell_area);
      else
        qout_kernel<<<dim3((E + 127) / 128, M, nt), 128, 0, st>>>(qa);
      ++launched;
    }
    MHM_CUDA_OK(cudaGetLastError());
  }

  rt->lean_ok = getenv("MHM_CUDA_NO_LEAN_ROUTING") == nullptr;
  if (const char* w = getenv("MHM_CUDA_LEAN_PF_WAVES")) rt->pf_waves = (size_t)atoll(w);
  for (const Segment& sg : segs) {
    if (int rc = ensure_c1c2(ctx, d, rt, sg.yId, timestep_rout)) return rc;
    ChainArgs ca{};
    ca.ev0 = sg.ev0;
    ca.ev1 = sg.ev1;
    ca.rs0 = ev[(size_t)sg.ev0].rs_first;
    ca.rl = rl;
    ca.E = E;
    ca.M = M;
    ca.nNodes = rt->nNodes;
    ca.nGslots = std::max(1, rt->nGslots);
    ca.single_node = rt->nNodes <= 1;
    ca.meta = rt->meta;
    ca.up_ptr = rt->up_ptr;
    ca.up_pos = rt->up_pos;
    ca.C1 = rt->C1;
    ca.C2 = rt->C2;
    ca.qout_hist = rt->qout_hist;
    ca.qtr_hist = rt->qtr_hist;
    ca.qTIN = rt->qTIN;
    ca.qTR = rt->qTR;
    ca.qMod = rt->qMod;
    ca.qOUT = rt->qOUT;
    ca.qmod_g = rt->qmod_g;
    for (size_t l = 0; l + 1 < rt->lvl_ptr.size(); ++l) {
      ca.lane0 = rt->lvl_ptr[l];
      ca.lane1 = rt->lvl_ptr[l + 1];
      const int cnt = ca.lane1 - ca.lane0;
      const int threads = cnt >= 128 ? 128 : cnt;  // multiples of 32
      const dim3 grid((cnt + threads - 1) / threads, M);
      const int ku = rt->lvl_ku[l];
      const bool mem = rt->lvl_mem[l] != 0;
#define MHM_CHAIN(R, K, Mm) route_chain_kernel<R, K, Mm><<<grid, threads, 0, st>>>(ca)
#define MHM_CHAIN_K(R)                                   \
  if (!mem) MHM_CHAIN(R, 1, false);                      \
  else if (ku <= 1) MHM_CHAIN(R, 1, true);               \
  else if (ku == 2) MHM_CHAIN(R, 2, true);               \
  else if (ku == 3) MHM_CHAIN(R, 3, true);               \
  else MHM_CHAIN(R, 4, true)
      const bool lean = rl == 1 && ca.rs0 == ca.ev0 && ca.ev0 % kHistTile == 0 && !ca.single_node && rt->lvl_plain[l] && rt->lean_ok;
      if (lean) {
        // less than one wave of CTAs: the variant that keeps the next window's loads in flight
        const bool pf = (size_t)grid.x * grid.y <= (size_t)ctx->sm_count * MHM_LEAN_MIN_BLOCKS * rt->pf_waves;
#define MHM_LEAN(K, Mm)                                                      \
  if (pf) route_chain_lean_kernel<K, Mm, true><<<grid, threads, 0, st>>>(ca); \
  else route_chain_lean_kernel<K, Mm, false><<<grid, threads, 0, st>>>(ca)
        if (!mem) { MHM_LEAN(1, false); }
        else if (ku <= 1) { MHM_LEAN(1, true); }
        else if (ku == 2) { MHM_LEAN(2, true); }
        else if (ku == 3) { MHM_LEAN(3, true); }
        else { MHM_LEAN(4, true); }
#undef MHM_LEAN
      } else if (rl == 1) {
        MHM_CHAIN_K(true);
      } else {
        MHM_CHAIN_K(false);
      }
#undef MHM_CHAIN_K
#undef MHM_CHAIN
      ++launched;
    }
    MHM_CUDA_OK(cudaGetLastError());
  }
  if (rt->nGauges > 0 || rt->nGaugesTotal > 0) {
    gauge_kernel<<<dim3((nEv + 63) / 64, M), 64, 0, st>>>(
        nEv, M, rt->nGauges, std::max(1, rt->nGslots), rt->nTimeSteps, rt->nGaugesTotal,
        rt->d_events, rt->d_gauge_col, rt->d_gauge_slot, rt->qmod_g, rt->gauge_hist);
    ++launched;
    MHM_CUDA_OK(cudaGetLastError());
  }
  ctx->stat_end(kStatRouting, launched);
  return 0;
}

// routing of model steps tt_first .. tt_first+n_steps-1 whose total runoff is in
// d->runoff_hist; restates the schedule of mo_mhm_interface_run.f90:460-612
bool routing_is_deferred(const Domain* d) { return d->rt && d->rt->deferred; }
bool routing_defer_block(Domain* d, int32_t tt_first, int32_t n_steps, bool fused) {
  Routing* rt = d->rt;
  if (!rt || !rt->deferred) return false;
  rt->pend_tt = tt_first;
  rt->pend_n = n_steps;
  rt->pend_fused = fused;
  return true;
}

bool routing_fuse_qout(mhm_cuda_context* ctx, Domain* d, int32_t n_steps, CellArgs* a) {
  Routing* rt = d->rt;
  if (!rt || !rt->bijective || rt->nInflowGauges > 0 || rt->nInflowTotal > 0 || d->keep_runoff_hist ||
      routing_accumulates(d, rt) || getenv("MHM_CUDA_NO_ALIGNED") || getenv("MHM_CUDA_NO_FUSED_QOUT"))
    return false;
  if (ensure(&rt->qout_hist, &rt->qout_cap, hist_size(n_steps + kHistPad, rt->M, rt->E), ctx->stream)) return false;
  a->qout_hist = rt->qout_hist;
  a->cell_lane = rt->d_cell_entry;
  a->cell_skew = rt->d_cell_skew;
  a->cell_area = rt->d_cell_area;
  a->qout_step0 = 0;
  a->qout_E = rt->E;
  a->qout_map_flag = rt->map_flag;
  a->qout_tst = 3600.0 * d->cfg.timestep_h;
  a->qout_scale = 1000.0 / a->qout_tst;
  return true;
}

int routing_run_block(mhm_cuda_context* ctx, Domain* d, int32_t tt_first, int32_t n_steps,
                      bool qout_ready) {
  Routing* rt = d->rt;
  MHM_REQUIRE(rt->nTimeSteps == d->axis.nTimeSteps && rt->gauge_hist,
              "routing: time axis changed after mrm_cuda_set_network");
  const int nTstepDay = 24 / d->cfg.timestep_h;
  const int nT = d->axis.nTimeSteps;
  if (rt->rout_case != 1) {
    // the members of a domain are routed side by side on one schedule, so they must share the
    // adaptive routing step mrm_update_param derived for them (mo_mrm_mpr.f90:294-321); checked
    // here and not in set_c1c2, because every evaluation of a calibration run sets all members anew
    double ts = 0.0;
    for (int m = 0; m < rt->M; ++m) {
      const double v = rt->TSrout_m[(size_t)m];
      MHM_REQUIRE(v > 0.0, "routing: mrm_cuda_set_c1c2 has not been called for member %d", m);
      MHM_REQUIRE(ts == 0.0 || ts == v,
                  "routing: members of one domain must share L11_TSrout (member %d: %g, others: %g)", m, v, ts);
      ts = v;
    }
    rt->TSrout = ts;
  }
  if (rt->inflow_acc.size() != (size_t)rt->nInflowTotal) rt->inflow_acc.assign((size_t)rt->nInflowTotal, 0.0);
  auto inflow_at = [&](int g, int day) -> double {  // InflowGauge%Q(day, g), 1-based day
    if (rt->inflowQ.empty()) return 0.0;
    return rt->inflowQ[(size_t)g * rt->nDays + (size_t)(day - 1)];
  };
  const bool accumulates = routing_accumulates(d, rt);
  const bool per_cell = rt->bijective && !accumulates && !getenv("MHM_CUDA_NO_ALIGNED");
  std::vector<DevEvent> ev;
  std::vector<Segment> segs;  // runs of events of one land-cover scene (C1/C2 of case 1)
  std::vector<double> inflow_val;
  int32_t acc_t0 = 0;                     // first block step not yet routed
  bool carry_live = rt->carry_steps > 0;  // rt->carry holds runoff of steps before acc_t0
  for (int32_t t = 0; t < n_steps; ++t) {
    const int tt = tt_first + t;
    const int day = (tt + nTstepDay - 1) / nTstepDay;  // iDischargeTS, :463
    DevEvent e{};
    bool fire = false;
    if (!accumulates) {  // case 1 (:465-474) and adaptive step shorter than the model step
      e.rout_loop = routing_rout_loop(d, rt);
      e.tst = 3600.0 * d->cfg.timestep_h;
      for (int g = 0; g < rt->nInflowTotal; ++g) rt->inflow_acc[(size_t)g] = inflow_at(g, day);
      e.t0 = t;
      e.nacc = 1;
      e.use_carry = 0;
      fire = true;
    } else {  // routing step longer than the model step: :493-512
      double fin = rt->TSrout / (d->cfg.timestep_h * 3600.0);
      for (int g = 0; g < rt->nInflowTotal; ++g)
        rt->inflow_acc[(size_t)g] = rt->inflow_acc[(size_t)g] + inflow_at(g, day);
      if (tt == nT && (tt % fortran_nint(fin)) != 0) fin = (double)(tt % fortran_nint(fin));
      if ((tt % fortran_nint(fin)) == 0 || tt == nT) {
        for (int g = 0; g < rt->nInflowTotal; ++g)
          rt->inflow_acc[(size_t)g] = rt->inflow_acc[(size_t)g] / fin;
        e.tst = 3600.0 * (double)(d->cfg.timestep_h * (int)fortran_nint(fin));
        long r = fortran_nint(1.0 / fin);
        e.rout_loop = (int32_t)(r < 1 ? 1 : r);
        e.backfill = (int32_t)fortran_nint(fin);
        e.t0 = acc_t0;
        e.nacc = t - acc_t0 + 1;
        e.use_carry = carry_live ? 1 : 0;
        fire = true;
      }
    }
    if (!fire) continue;
    e.tt = tt;
    const int yId = d->h_idx[(size_t)(tt - 1)].yId;  // scene of the step that calls mRM_routing
    if (segs.empty() || segs.back().yId != yId) segs.push_back(Segment{(int32_t)ev.size(), 0, yId});
    ev.push_back(e);
    segs.back().ev1 = (int32_t)ev.size();
    inflow_val.insert(inflow_val.end(), rt->inflow_acc.begin(), rt->inflow_acc.end());
    std::fill(rt->inflow_acc.begin(), rt->inflow_acc.end(), 0.0);  // :596, :606
    acc_t0 = t + 1;
    carry_live = false;
  }
  rt->last_n = n_steps;
  MHM_REQUIRE((rt->nGhost == 0 && rt->nExport == 0) || (!accumulates && routing_rout_loop(d, rt) == 1),
              "routing: sharded domains need one routing step per model step");
  MHM_REQUIRE(!qout_ready || (per_cell && (int32_t)ev.size() == n_steps),
              "routing: fused node runoff does not match the block's routing schedule");
  if (int rc = run_events(ctx, d, rt, ev, segs, inflow_val, d->runoff_hist, per_cell, qout_ready,
                          (double)d->cfg.timestep_h))
    return rc;
  // steps at the end of the block that wait for a later routing call
  if (accumulates && acc_t0 < n_steps) {
    carry_kernel<<<dim3((rt->nCells1 + 127) / 128, rt->M), 128, 0, ctx->stream>>>(
        rt->nCells1, rt->M, acc_t0, n_steps - acc_t0, carry_live ? 1 : 0, d->runoff_hist, rt->carry);
    MHM_CUDA_OK(cudaGetLastError());
    rt->carry_steps = (carry_live ? rt->carry_steps : 0) + (n_steps - acc_t0);
  } else {
    rt->carry_steps = 0;
  }
  return 0;
}

}  // namespace mhm

using namespace mhm;

extern "C" {


// ---- sub-catchment sharding ---------------------------------------------------------------
__global__ void export_outflow_kernel(int nList, int M, int E, int T, const LaneMeta* __restrict__ meta,
    const int32_t* __restrict__ lanes,
                                      const double* __restrict__ qtr_hist, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y, m = blockIdx.z;
  if (t >= T) return;
  out[((size_t)m * nList + e) * T + t] = qtr_hist[hidx(t + lane_wshift(meta[lanes[e]].flags), M, E, m, lanes[e])];
}
__global__ void import_outflow_kernel(int nList, int M, int E, int T, const LaneMeta* __restrict__ meta,
    const int32_t* __restrict__ lanes,
                                      const double* __restrict__ in, double* __restrict__ qtr_hist) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y, m = blockIdx.z;
  if (t >= T) return;
  qtr_hist[hidx(t + lane_wshift(meta[lanes[e]].flags), M, E, m, lanes[e])] = in[((size_t)m * nList + e) * T + t];
}

// the same for the exchange buffers of mrm_cuda_shard_run_steps: one contiguous piece
// [member][link of the piece][step] per peer rank, pieces in rank order
__global__ void xchg_pack_kernel(int nList, int M, int E, int T, const LaneMeta* __restrict__ meta,
    const int32_t* __restrict__ lanes,
                                 const int32_t* __restrict__ tab, const double* __restrict__ qtr_hist,
                                 double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y, m = blockIdx.z;
  if (t >= T) return;
  const size_t first = (size_t)tab[e], cnt = (size_t)tab[nList + e], loc = (size_t)tab[2 * nList + e];
  out[(first * M + (size_t)m * cnt + loc) * T + t] = qtr_hist[hidx(t + lane_wshift(meta[lanes[e]].flags), M, E, m, lanes[e])];
}
__global__ void xchg_unpack_kernel(int nList, int M, int E, int T, const LaneMeta* __restrict__ meta,
    const int32_t* __restrict__ lanes,
                                   const int32_t* __restrict__ tab, const double* __restrict__ in,
                                   double* __restrict__ qtr_hist) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y, m = blockIdx.z;
  if (t >= T) return;
  const size_t first = (size_t)tab[e], cnt = (size_t)tab[nList + e], loc = (size_t)tab[2 * nList + e];
  qtr_hist[hidx(t + lane_wshift(meta[lanes[e]].flags), M, E, m, lanes[e])] = in[(first * M + (size_t)m * cnt + loc) * T + t];
}

int mrm_partition_subcatchments(int32_t nNodes, int32_t nLinks, const int32_t* fromN,
                                const int32_t* toN, const int32_t* netPerm, int32_t nParts,
                                int32_t* part_of_node) {
  MHM_REQUIRE(nNodes >= 1 && nLinks >= 0 && nLinks <= nNodes && nParts >= 1 && part_of_node &&
                  (nLinks == 0 || (fromN && toN && netPerm)),
              "partition_subcatchments: bad arguments");
  // subtree weights; netPerm is a topological order (upstream links first)
  std::vector<int64_t> w((size_t)nNodes, 1);
  std::vector<int32_t> down((size_t)nNodes, -1);
  for (int k = 0; k < nLinks; ++k) {
    const int i = netPerm[k] - 1;
    MHM_REQUIRE(i >= 0 && i < nLinks, "partition_subcatchments: netPerm(%d) outside 1..%d", k + 1, nLinks);
    const int f = fromN[i] - 1, t = toN[i] - 1;
    MHM_REQUIRE(f >= 0 && f < nNodes && t >= 0 && t < nNodes, "partition_subcatchments: link %d", i + 1);
    w[(size_t)t] += w[(size_t)f];
    down[(size_t)f] = t;
  }
  // skeleton = nodes whose subtree holds at least total / (8 nParts) nodes: the trunk, shard 0's.
  // Every maximal subtree beside the skeleton is a sub-catchment: it hangs off the skeleton by
  // one link (the cut) or is a whole small basin with its own outlet (no cut at all).
  // Sub-catchments smaller than 1/64 of the threshold stay with the trunk (too many cuts else).
  const int64_t thr = std::max<int64_t>(1, (int64_t)nNodes / ((int64_t)nParts * 8));
  const int64_t minw = std::max<int64_t>(1, thr / 64);
  std::vector<char> skel((size_t)nNodes, 0);
  for (int nd = 0; nd < nNodes; ++nd) skel[(size_t)nd] = nParts > 1 && w[(size_t)nd] >= thr;
  std::vector<int32_t> roots;
  int64_t trunk = 0;
  for (int nd = 0; nd < nNodes; ++nd) {
    if (skel[(size_t)nd]) {
      ++trunk;
      continue;
    }
    const int dn = down[(size_t)nd];
    if (dn >= 0 && !skel[(size_t)dn]) continue;  // inside a sub-catchment
    if (nParts > 1 && (dn < 0 || w[(size_t)nd] >= minw)) roots.push_back(nd);
    else trunk += w[(size_t)nd];
  }
  std::stable_sort(roots.begin(), roots.end(), [&](int a, int b) { return w[(size_t)a] > w[(size_t)b]; });
  std::vector<int64_t> load((size_t)nParts, 0);
  load[0] = trunk;
  std::vector<int32_t> owner((size_t)nNodes, -2);  // -2 unknown
  for (int rnode : roots) {  // heaviest first onto the least loaded shard
    int best = 0;
    for (int p = 1; p < nParts; ++p)
      if (load[(size_t)p] < load[(size_t)best]) best = p;
    owner[(size_t)rnode] = best;
    load[(size_t)best] += w[(size_t)rnode];
  }
  for (int nd = 0; nd < nNodes; ++nd)
    if (owner[(size_t)nd] == -2 && (skel[(size_t)nd] || down[(size_t)nd] < 0 || skel[(size_t)down[(size_t)nd]]))
      owner[(size_t)nd] = 0;  // skeleton, and small sub-catchments kept with the trunk
  // everything else takes the shard of the node below it (reverse topological order)
  for (int k = nLinks - 1; k >= 0; --k) {
    const int i = netPerm[k] - 1;
    const int f = fromN[i] - 1, t = toN[i] - 1;
    if (owner[(size_t)f] == -2) owner[(size_t)f] = owner[(size_t)t];
  }
  for (int nd = 0; nd < nNodes; ++nd) part_of_node[nd] = owner[(size_t)nd] < 0 ? 0 : owner[(size_t)nd];
  return 0;
}

int mrm_cuda_set_deferred(mhm_cuda_context* ctx, int32_t iDomain, int32_t deferred) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(d->rt, "set_deferred: no network set");
  d->rt->deferred = deferred != 0;
  d->rt->pend_n = 0;
  return 0;
}

int mrm_cuda_route_pending(mhm_cuda_context* ctx, int32_t iDomain) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt && rt->deferred && rt->pend_n > 0, "route_pending: no block is waiting for its routing");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const int32_t tt = rt->pend_tt, n = rt->pend_n;
  rt->pend_n = 0;
  return routing_run_block(ctx, d, tt, n, rt->pend_fused);
}

int mrm_cuda_export_outflow(mhm_cuda_context* ctx, int32_t iDomain, double* dev_out, int32_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt && dev_out && n_steps >= 1 && n_steps == rt->last_n && rt->qtr_hist,
              "export_outflow: the last routed block has %d steps", rt ? rt->last_n : 0);
  if (rt->nExport == 0) return 0;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  export_outflow_kernel<<<dim3((n_steps + 63) / 64, rt->nExport, rt->M), 64, 0, ctx->stream>>>(
      rt->nExport, rt->M, rt->E, n_steps, rt->meta, rt->d_export_lane, rt->qtr_hist, dev_out);
  MHM_CUDA_OK(cudaGetLastError());
  return 0;
}

int mrm_cuda_import_outflow(mhm_cuda_context* ctx, int32_t iDomain, const double* dev_in, int32_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt && dev_in && rt->deferred && n_steps == rt->pend_n,
              "import_outflow: the pending block has %d steps", rt ? rt->pend_n : 0);
  if (rt->nGhost == 0) return 0;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  if (int rc = ensure(&rt->qtr_hist, &rt->qtr_cap, hist_size(n_steps + kHistPad, rt->M, rt->E), ctx->stream)) return rc;
  import_outflow_kernel<<<dim3((n_steps + 63) / 64, rt->nGhost, rt->M), 64, 0, ctx->stream>>>(
      rt->nGhost, rt->M, rt->E, n_steps, rt->meta, rt->d_ghost_lane, dev_in, rt->qtr_hist);
  MHM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- sub-catchment sharding with the exchange below the C ABI --------------------------------
static std::vector<int32_t> piece_table(const std::vector<int32_t>& counts, int32_t n_list) {
  std::vector<int32_t> tab((size_t)3 * n_list);
  int32_t e = 0, first = 0;
  for (int32_t c : counts) {
    for (int32_t k = 0; k < c; ++k, ++e) {
      tab[(size_t)e] = first;
      tab[(size_t)n_list + e] = c;
      tab[(size_t)2 * n_list + e] = k;
    }
    first += c;
  }
  return tab;
}

int mrm_cuda_set_exchange(mhm_cuda_context* ctx, int32_t iDomain, const int32_t* send_links,
                          const int32_t* recv_links) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt && send_links && recv_links, "set_exchange: no network set, or null plan");
  MHM_REQUIRE(ctx->nranks == 1 || ctx->nccl_comm, "set_exchange: no communicator (mhm_cuda_comm_init)");
  rt->xsend.assign(send_links, send_links + ctx->nranks);
  rt->xrecv.assign(recv_links, recv_links + ctx->nranks);
  int64_t ns = 0, nr = 0;
  for (int r = 0; r < ctx->nranks; ++r) {
    MHM_REQUIRE(send_links[r] >= 0 && recv_links[r] >= 0 && (r != ctx->rank || (send_links[r] == 0 && recv_links[r] == 0)),
                "set_exchange: bad plan entry for rank %d", r);
    ns += send_links[r];
    nr += recv_links[r];
  }
  MHM_REQUIRE(ns == rt->nExport && nr == rt->nGhost,
              "set_exchange: the plan sends %lld / receives %lld links, the network has %d exports / %d ghost sources",
              (long long)ns, (long long)nr, rt->nExport, rt->nGhost);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  if (int rc = upload(&rt->d_xs_tab, piece_table(rt->xsend, rt->nExport), ctx->stream)) return rc;
  if (int rc = upload(&rt->d_xr_tab, piece_table(rt->xrecv, rt->nGhost), ctx->stream)) return rc;
  rt->deferred = true;
  rt->pend_n = 0;
  return 0;
}

// receive the ghost series of the pending block, route it, send the export series on
static int shard_route_pending(mhm_cuda_context* ctx, Domain* d) {
  Routing* rt = d->rt;
  const int32_t n = rt->pend_n, tt = rt->pend_tt;
  if (n <= 0) return 0;
  rt->pend_n = 0;
  const int N = ctx->nranks, M = rt->M;
  std::vector<size_t> cnt((size_t)N);
  if (rt->nGhost > 0) {
    for (int r = 0; r < N; ++r) cnt[(size_t)r] = (size_t)rt->xrecv[(size_t)r] * M * n;
    if (int rc = ensure(&rt->xrecv_buf, &rt->xrecv_cap, (size_t)rt->nGhost * M * n, ctx->stream)) return rc;
    if (int rc = ensure(&rt->qtr_hist, &rt->qtr_cap, hist_size(n + kHistPad, M, rt->E), ctx->stream)) return rc;
    if (int rc = comm_send_recv(ctx, nullptr, nullptr, rt->xrecv_buf, cnt.data(), ctx->stream)) return rc;
    xchg_unpack_kernel<<<dim3((n + 63) / 64, rt->nGhost, M), 64, 0, ctx->stream>>>(
        rt->nGhost, M, rt->E, n, rt->meta, rt->d_ghost_lane, rt->d_xr_tab, rt->xrecv_buf, rt->qtr_hist);
    MHM_CUDA_OK(cudaGetLastError());
  }
  if (int rc = routing_run_block(ctx, d, tt, n, rt->pend_fused)) return rc;
  if (rt->nExport > 0) {
    for (int r = 0; r < N; ++r) cnt[(size_t)r] = (size_t)rt->xsend[(size_t)r] * M * n;
    if (int rc = ensure(&rt->xsend_buf, &rt->xsend_cap, (size_t)rt->nExport * M * n, ctx->stream)) return rc;
    xchg_pack_kernel<<<dim3((n + 63) / 64, rt->nExport, M), 64, 0, ctx->stream>>>(
        rt->nExport, M, rt->E, n, rt->meta, rt->d_export_lane, rt->d_xs_tab, rt->qtr_hist, rt->xsend_buf);
    MHM_CUDA_OK(cudaGetLastError());
    if (int rc = comm_send_recv(ctx, rt->xsend_buf, cnt.data(), nullptr, nullptr, ctx->stream)) return rc;
  }
  return 0;
}

int mrm_cuda_shard_run_steps(mhm_cuda_context* ctx, int32_t iDomain, int32_t tt_first, int32_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt && rt->deferred && (int)rt->xsend.size() == ctx->nranks,
              "shard_run_steps: mrm_cuda_set_exchange has not been called");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  // a shard that receives works one block behind its senders: it routes the block whose ghost
  // series were sent during the senders' previous call, then runs this block's cells -- nobody waits
  if (rt->nGhost > 0)
    if (int rc = shard_route_pending(ctx, d)) return rc;
  if (int rc = mhm_cuda_run_steps(ctx, iDomain, tt_first, n_steps)) return rc;  // cells; routing pending
  if (rt->nGhost == 0)
    if (int rc = shard_route_pending(ctx, d)) return rc;
  return 0;
}

int mrm_cuda_shard_flush(mhm_cuda_context* ctx, int32_t iDomain) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt && rt->deferred, "shard_flush: the domain is not sharded");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  return shard_route_pending(ctx, d);
}

int mrm_routing_order(int32_t nNodes, int32_t nLinks, const int32_t* fromN, const int32_t* toN,
                      int32_t* rOrder, int32_t* netPerm) {
  MHM_REQUIRE(nNodes >= 1 && nLinks >= 0 && nLinks <= nNodes && fromN && toN && rOrder && netPerm,
              "mrm_routing_order: bad arguments");
  return routing_order_linear(nNodes, nLinks, fromN, toN, rOrder, netPerm);
}

int mrm_cuda_set_network(mhm_cuda_context* ctx, int32_t iDomain, const mrm_network* net) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(net && net->nNodes >= 1 && net->nOutlets >= 0 && net->nOutlets <= net->nNodes,
              "set_network: bad sizes");
  MHM_REQUIRE(d->has_time, "set_network: call mhm_cuda_set_time first (gauge series length)");
  MHM_REQUIRE(net->processCase >= 1 && net->processCase <= 3, "set_network: routing case %d",
              net->processCase);
  MHM_REQUIRE(net->nNodes == 1 || (net->netPerm && net->fromN && net->toN),
              "set_network: netPerm/fromN/toN required");
  MHM_REQUIRE(net->L1_areaCell && net->L11_areaCell, "set_network: cell areas required");
  MHM_REQUIRE(net->map_flag ? net->L1_L11_Id != nullptr : net->L11_L1_Id != nullptr,
              "set_network: L1_L11_Id (map_flag) or L11_L1_Id required");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (d->rt) routing_free(d->rt);
  d->rt = nullptr;
  auto* rt = new Routing();
  rt->nCells1 = d->cfg.nCells;
  rt->nNodes = net->nNodes;
  rt->nOutlets = net->nOutlets;
  rt->nLinks = net->nNodes - net->nOutlets;
  rt->map_flag = net->map_flag;
  rt->rout_case = net->processCase;
  rt->nGauges = net->nGauges;
  rt->nInflowGauges = net->nInflowGauges;
  rt->nGaugesTotal = net->nGaugesTotal;
  rt->nInflowTotal = net->nInflowTotal;
  rt->M = d->cfg.nMembers;
  rt->nLC = d->cfg.nLCscenes;
  rt->nTimeSteps = d->axis.nTimeSteps;
  rt->ssMax_global = net->ssMax;
  rt->gaugeIndexList.assign(net->gaugeIndexList, net->gaugeIndexList + net->nGauges);
  rt->gaugeNodeList.assign(net->gaugeNodeList, net->gaugeNodeList + net->nGauges);
  if (int rc = build_topology(ctx, d, rt, net)) {
    routing_free(rt);
    return rc;
  }
  const size_t M = (size_t)rt->M, nn = (size_t)rt->nNodes;
  cudaStream_t st = ctx->stream;
  int rc = 0;
  rc |= alloc_zero(&rt->C1, M * nn, st);
  rc |= alloc_zero(&rt->C2, M * nn, st);
  rc |= alloc_zero(&rt->qOUT, M * nn, st);
  rc |= alloc_zero(&rt->qMod, M * nn, st);
  rc |= alloc_zero(&rt->qTIN, M * 2 * nn, st);
  rc |= alloc_zero(&rt->qTR, M * 2 * nn, st);
  rc |= alloc_zero(&rt->carry, M * (size_t)rt->nCells1, st);
  rc |= alloc_zero(&rt->gauge_hist, M * (size_t)std::max(1, rt->nGaugesTotal) * rt->nTimeSteps, st);
  if (rc) {
    routing_free(rt);
    return 2;
  }
  MHM_CUDA_OK(cudaStreamSynchronize(st));
  rt->param5.assign(M, {});
  rt->c1c2_yId.assign(M, -1);
  rt->TSrout_m.assign(M, 0.0);
  d->rt = rt;
  return 0;
}

int mrm_cuda_set_reg_rout(mhm_cuda_context* ctx, int32_t iDomain, int32_t member,
                          const double* param5, const double* length, const double* slope,
                          const double* fFPimp) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt, "set_reg_rout: no network set");
  MHM_REQUIRE(member >= 0 && member < rt->M && param5 && length && slope && fFPimp,
              "set_reg_rout: bad arguments");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t nn = (size_t)rt->nNodes;
  cudaStream_t st = ctx->stream;
  if (!rt->d_length) {
    MHM_CUDA_OK(cudaMalloc(&rt->d_length, nn * sizeof(double)));
    MHM_CUDA_OK(cudaMalloc(&rt->d_slope, nn * sizeof(double)));
    MHM_CUDA_OK(cudaMalloc(&rt->d_fFPimp, (size_t)rt->M * rt->nLC * nn * sizeof(double)));
  }
  // L11_length(s11:e11-1), L11_slope(s11:e11-1): nNodes-1 values (mo_mhm_interface_run.f90:577)
  const size_t ns = nn > 1 ? nn - 1 : 0;
  if (ns) {
    MHM_CUDA_OK(cudaMemcpyAsync(rt->d_length, length, ns * sizeof(double), cudaMemcpyHostToDevice, st));
    MHM_CUDA_OK(cudaMemcpyAsync(rt->d_slope, slope, ns * sizeof(double), cudaMemcpyHostToDevice, st));
    double mx = slope[0];
    for (size_t i = 1; i < ns; ++i) mx = slope[i] > mx ? slope[i] : mx;  // maxval(slope(:))
    rt->ssMax = rt->ssMax_global > 0.0 ? rt->ssMax_global : mx;
  }
  MHM_CUDA_OK(cudaMemcpyAsync(rt->d_fFPimp + (size_t)member * rt->nLC * nn, fFPimp,
                              (size_t)rt->nLC * nn * sizeof(double), cudaMemcpyHostToDevice, st));
  MHM_CUDA_OK(cudaStreamSynchronize(st));
  rt->param5[(size_t)member].assign(param5, param5 + 5);
  rt->c1c2_yId[(size_t)member] = -1;
  return 0;
}

int mrm_cuda_set_c1c2(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, const double* C1,
                      const double* C2, double TSrout) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt, "set_c1c2: no network set");
  MHM_REQUIRE(member >= 0 && member < rt->M && C1 && C2, "set_c1c2: bad arguments");
  MHM_REQUIRE(rt->rout_case == 1 || TSrout > 0.0, "set_c1c2: L11_TSrout must be > 0 for case 2/3");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t nn = (size_t)rt->nNodes;
  MHM_CUDA_OK(cudaMemcpyAsync(rt->C1 + (size_t)member * nn, C1, nn * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream));
  MHM_CUDA_OK(cudaMemcpyAsync(rt->C2 + (size_t)member * nn, C2, nn * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  rt->param5[(size_t)member].clear();
  if (rt->rout_case != 1) {
    rt->TSrout_m[(size_t)member] = TSrout;
    rt->TSrout = TSrout;  // settled (and checked across members) when the next block is routed
  }
  return 0;
}

static int mrm_state_ptr(Routing* rt, int id, int member, double** p, size_t* rows) {
  const size_t nn = (size_t)rt->nNodes;
  switch (id) {
    case MRM_S_QOUT: *p = rt->qOUT + (size_t)member * nn; *rows = 1; break;
    case MRM_S_QMOD: *p = rt->qMod + (size_t)member * nn; *rows = 1; break;
    case MRM_S_C1: *p = rt->C1 + (size_t)member * nn; *rows = 1; break;
    case MRM_S_C2: *p = rt->C2 + (size_t)member * nn; *rows = 1; break;
    case MRM_S_QTIN: *p = rt->qTIN + (size_t)member * 2 * nn; *rows = 2; break;
    case MRM_S_QTR: *p = rt->qTR + (size_t)member * 2 * nn; *rows = 2; break;
    default: set_error("mrm state id %d unknown", id); return 1;
  }
  return 0;
}

int mrm_cuda_set_state(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t id,
                       const double* base, int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt, "mrm_set_state: no network set");
  MHM_REQUIRE(member >= 0 && member < rt->M && base && ld >= rt->nNodes && offset >= 0,
              "mrm_set_state: bad arguments");
  double* p;
  size_t rows;
  if (int rc = mrm_state_ptr(rt, id, member, &p, &rows)) return rc;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaMemcpy2DAsync(p, (size_t)rt->nNodes * sizeof(double), base + offset,
                                (size_t)ld * sizeof(double), (size_t)rt->nNodes * sizeof(double), rows,
                                cudaMemcpyHostToDevice, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mrm_cuda_get_state(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t id,
                       double* base, int64_t ld, int64_t offset) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt, "mrm_get_state: no network set");
  MHM_REQUIRE(member >= 0 && member < rt->M && base && ld >= rt->nNodes && offset >= 0,
              "mrm_get_state: bad arguments");
  double* p;
  size_t rows;
  if (int rc = mrm_state_ptr(rt, id, member, &p, &rows)) return rc;
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaMemcpy2DAsync(base + offset, (size_t)ld * sizeof(double), p,
                                (size_t)rt->nNodes * sizeof(double), (size_t)rt->nNodes * sizeof(double),
                                rows, cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mrm_cuda_set_inflow(mhm_cuda_context* ctx, int32_t iDomain, const double* Q, int64_t nDays) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt, "set_inflow: no network set");
  MHM_REQUIRE(Q && nDays >= 1, "set_inflow: bad arguments");
  rt->inflowQ.assign(Q, Q + (size_t)nDays * rt->nInflowTotal);
  rt->nDays = nDays;
  return 0;
}

int mrm_cuda_route(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t tt, int32_t yId,
                   const double* RunToRout, int32_t timestep_rout, double tsRoutFactorIn,
                   const double* InflowDischarge) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt, "route: no network set");
  MHM_REQUIRE(rt->M == 1 && member == 0, "route: the per-step seam serves single-member domains");
  MHM_REQUIRE(tt >= 1 && tt <= rt->nTimeSteps && timestep_rout >= 1 && tsRoutFactorIn > 0.0,
              "route: bad arguments");
  MHM_REQUIRE(yId >= 1 && yId <= rt->nLC, "route: yId %d outside 1..%d", yId, rt->nLC);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n1 = (size_t)rt->nCells1;
  const double* src = d->F[MHM_F_TOTAL_RUNOFF];
  if (RunToRout) {
    MHM_CUDA_OK(cudaMemcpyAsync(rt->carry, RunToRout, n1 * sizeof(double), cudaMemcpyHostToDevice,
                                ctx->stream));
    MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    src = rt->carry;
  }
  DevEvent e{};
  e.tt = tt;
  e.t0 = 0;
  e.nacc = 1;  // the "history" is the single row `src`
  e.use_carry = 0;
  long rl = fortran_nint(1.0 / tsRoutFactorIn);
  e.rout_loop = (int32_t)(rl < 1 ? 1 : rl);
  e.tst = 3600.0 * timestep_rout;
  e.backfill = 0;
  std::vector<DevEvent> ev{e};
  std::vector<double> inflow_val((size_t)rt->nInflowTotal, 0.0);
  if (InflowDischarge)
    for (int g = 0; g < rt->nInflowTotal; ++g) inflow_val[(size_t)g] = InflowDischarge[g];
  std::vector<Segment> segs{Segment{0, 1, yId}};
  return run_events(ctx, d, rt, ev, segs, inflow_val, src, false, false, (double)timestep_rout);
}

int mrm_cuda_get_runoff(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, double* out,
                        int64_t ld, int32_t tt_first, int32_t n_steps) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  Routing* rt = d->rt;
  MHM_REQUIRE(rt, "get_runoff: no network set");
  MHM_REQUIRE(member >= 0 && member < rt->M && out && ld >= tt_first + n_steps - 1 &&
                  tt_first >= 1 && n_steps >= 1 && tt_first + n_steps - 1 <= rt->nTimeSteps,
              "get_runoff: bad arguments");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const double* H = rt->gauge_hist + (size_t)member * rt->nGaugesTotal * rt->nTimeSteps;
  MHM_CUDA_OK(cudaMemcpy2DAsync(out + (tt_first - 1), (size_t)ld * sizeof(double), H + (tt_first - 1),
                                (size_t)rt->nTimeSteps * sizeof(double), (size_t)n_steps * sizeof(double),
                                (size_t)rt->nGaugesTotal, cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
