// routing.cu -- mRM Muskingum routing on the device, time-blocked and level-scheduled.
//
// Reference behaviour restated here (never copied):
//   mRM/mo_mrm_routing.f90:104-303 (mRM_routing), :380-481 (L11_routing)
//   mRM/mo_mrm_pre_routing.f90:77-143 (L11_runoff_acc), :179-214 (add_inflow)
//   mRM/mo_mrm_mpr.f90:61-119 (reg_rout)
//   mRM/mo_mrm_net_startup.f90:728-859 (L11_routing_order)
//   mHM/mo_mhm_interface_run.f90:460-612 (routing schedule, gauge back-fill)
//
// The reference sweeps the links serially in netPerm order once per routing step.  Its
// data dependence is (node, step) <- (upstream nodes, same step) and (node, step-1), so a
// whole block of routing steps can be done level by level: every link of one network
// level advances through all steps of the block with its Muskingum state in registers,
// reading the already finished outflow history of its upstream links.  Upstream inflows
// are added in netPerm order and the node's own runoff last, exactly like the serial
// sweep, so the result is bit-identical to it.
//
// This file is compiled with -fmad=false: C1*(a-b) + C2*(c-d) must round as in Fortran.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cooperative_groups.h>

#include "context.h"

namespace cg = cooperative_groups;

namespace mhm {

constexpr int kTailCluster = 8;       // CTAs per cluster in the tail kernel (portable maximum)
constexpr int kTailThreads = 256;     // threads per CTA in the tail kernel
constexpr int kTailMaxLevel = 2048;   // levels wider than this get their own launch

// everything a routing thread needs to know about its entry, one 32-byte record
struct alignas(16) EntMeta {
  int32_t node;    // 0-based L11 node
  int32_t link;    // 0-based link (C1/C2 index)
  int32_t flags;   // kEnt* bits | number of upstream links << 8
  int32_t gslot;   // gauge slot or -1
  int32_t up[4];   // entry positions of the first four upstream links (netPerm order)
};
constexpr int kMetaUps = 4;
static_assert(sizeof(EntMeta) == 32, "EntMeta must be 32 bytes");

// History buffers (node runoff qOUT and routed outflow qTR of every routing step of a block)
// are tiled by 8 steps: [step / 8][member][entry][step % 8].  An entry's 8 consecutive steps
// are one 64-byte run, so a level thread streams its own and its upstream links' series with
// 128-bit loads, and neighbouring entries of a level are neighbouring runs.
constexpr int kHistTile = 8;
__host__ __device__ __forceinline__ size_t hidx(int step, int M, int E, int m, int p) {
  return ((((size_t)(step >> 3) * M + m) * E + p) << 3) + (size_t)(step & 7);
}
__host__ __device__ __forceinline__ size_t hist_size(int steps, int M, int E) {
  return (size_t)((steps + kHistTile - 1) / kHistTile) * M * E * kHistTile;
}

enum : int32_t {
  kEntLink = 1,     // entry is a link (has Muskingum state); otherwise an outlet node
  kEntAddQout = 2,  // node's own runoff is added to its inflow
  kEntZeroOut = 4,  // routed outflow feeds a non-headwater inflow gauge: set to zero
  kEntInflow = 8,   // node is an inflow gauge: add_inflow applies to its runoff
};

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
  int32_t E = 0;                  // entries = links + outlet nodes
  std::vector<int32_t> lvl_ptr;   // entry range per level
  std::vector<int32_t> gaugeIndexList, gaugeNodeList, inflowIndexList, inflowHeadwater,
      inflowNodeList;
  // device topology (shared by members)
  int32_t *ent_node = nullptr, *ent_link = nullptr, *ent_flags = nullptr, *ent_gslot = nullptr;
  int32_t *up_ptr = nullptr, *up_pos = nullptr;
  EntMeta* meta = nullptr;
  int32_t* d_lvl_ptr = nullptr;
  int32_t tail_level = 0;  // levels >= tail_level are swept by one clustered wavefront kernel
  // aligned mode: L1 cells and L11 nodes map one to one, so the cell kernel hands its runoff
  // straight to the routing history (no L11_runoff_acc pass)
  bool bijective = false;
  int32_t* d_cell_entry = nullptr;  // [nCells1] routing entry of the cell's node
  double* d_cell_area = nullptr;    // [nCells1] area factor of mo_mrm_pre_routing.f90:125/:141
  int32_t *cell_ptr = nullptr, *cell_idx = nullptr;  // map_flag: L1 cells of each node, ascending
  int32_t* L11_L1_Id = nullptr;                      // !map_flag
  int32_t *d_inflow_node = nullptr, *d_inflow_index = nullptr, *d_inflow_head = nullptr;
  double *L1_area = nullptr, *L11_area = nullptr;
  int32_t nGslots = 0;
  int32_t *d_gauge_col = nullptr, *d_gauge_slot = nullptr;  // per gauge: column-1, slot
  // per member state, device [M][...]
  double *C1 = nullptr, *C2 = nullptr, *qOUT = nullptr, *qMod = nullptr;
  double *qTIN = nullptr, *qTR = nullptr;  // [M][2][nNodes]
  // reg_rout inputs
  std::vector<std::vector<double>> param5;  // per member, empty = C1/C2 given
  std::vector<int32_t> c1c2_yId;            // scene the member's C1/C2 were computed for
  double *d_length = nullptr, *d_slope = nullptr, *d_fFPimp = nullptr;  // fFPimp [M][nLC][nNodes]
  double ssMax = 0.0;
  int32_t nLC = 1;
  double TSrout = 0.0;  // case 2/3 [s]
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
};

void routing_free(Routing* rt) {
  if (!rt) return;
  void* ptrs[] = {rt->ent_node, rt->ent_link,  rt->ent_flags, rt->ent_gslot, rt->up_ptr,
                  rt->up_pos,   rt->meta, rt->d_lvl_ptr, rt->d_cell_entry, rt->d_cell_area, rt->cell_ptr,  rt->cell_idx,  rt->L11_L1_Id, rt->d_inflow_node,
                  rt->d_inflow_index, rt->d_inflow_head, rt->L1_area, rt->L11_area,
                  rt->d_gauge_col, rt->d_gauge_slot, rt->C1, rt->C2, rt->qOUT, rt->qMod,
                  rt->qTIN, rt->qTR, rt->d_length, rt->d_slope, rt->d_fFPimp, rt->carry,
                  rt->gauge_hist, rt->qout_hist, rt->qtr_hist, rt->qmod_g, rt->d_inflow_val,
                  rt->d_events};
  for (void* p : ptrs) cudaFree(p);
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
  const int32_t *ent_node, *cell_ptr, *cell_idx, *L11_L1_Id;
  const double *L1_area, *L11_area;
  const int32_t *inflow_node, *inflow_index, *inflow_head;
  const double* inflow_val;  // [nEvents][nInflowTotal]
  double* qout_hist;         // tiled, see hidx()
};

// L11_runoff_acc + add_inflow for every (event, member, entry); a thread produces the 8
// events of one history tile for its entry (one 64-byte run)
__global__ void __launch_bounds__(128) qout_kernel(const QoutArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.E) return;
  const int m = blockIdx.y, tile = blockIdx.z;
  const int node = a.ent_node[p];  // 0-based
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
    for (int g = 0; g < a.nInflowGauges; ++g) {  // add_inflow :203-213
      if (a.inflow_node[g] - 1 == node) {
        const double qi = a.inflow_val[(size_t)ev * a.nInflowTotal + a.inflow_index[g] - 1];
        v = a.inflow_head[g] ? v + qi : qi;
      }
    }
    q[d] = v;
  }
  double2* dst = reinterpret_cast<double2*>(a.qout_hist + hidx(tile * kHistTile, a.M, a.E, m, p));
#pragma unroll
  for (int d = 0; d < kHistTile / 2; ++d) dst[d] = make_double2(q[2 * d], q[2 * d + 1]);
}

// Same for the common one-cell-per-node case: a thread takes one L1 cell, reads its runoff
// of the 8 events of a tile (coalesced over cells) and writes the finished 64-byte run at the
// cell's routing entry.  Only used when every event is a single model step.
struct QoutCellArgs {
  int32_t nCells1, E, M, nEvents, map_flag, nInflowGauges, nInflowTotal;
  const DevEvent* events;
  const double* runoff_hist;  // [steps][M][nCells1]
  const int32_t* cell_entry;  // [nCells1]
  const double* cell_area;    // [nCells1]
  const EntMeta* meta;
  const int32_t *inflow_node, *inflow_index, *inflow_head;
  const double* inflow_val;
  double* qout_hist;
};
__global__ void __launch_bounds__(128) qout_cell_kernel(const QoutCellArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.nCells1) return;
  const int m = blockIdx.y, tile = blockIdx.z;
  const size_t n1 = (size_t)a.nCells1;
  const int p = a.cell_entry[k];
  const double area = a.cell_area[k];
  const EntMeta em = a.meta[p];
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
    if (em.flags & kEntInflow) {
      for (int g = 0; g < a.nInflowGauges; ++g) {
        if (a.inflow_node[g] - 1 == em.node) {
          const double qi = a.inflow_val[(size_t)ev * a.nInflowTotal + a.inflow_index[g] - 1];
          v = a.inflow_head[g] ? v + qi : qi;
        }
      }
    }
    q[d] = v;
  }
  double2* dst = reinterpret_cast<double2*>(a.qout_hist + hidx(tile * kHistTile, a.M, a.E, m, p));
#pragma unroll
  for (int d = 0; d < kHistTile / 2; ++d) dst[d] = make_double2(q[2 * d], q[2 * d + 1]);
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

struct LevelArgs {
  int32_t p0, p1;  // entry range of the level
  int32_t E, M, nNodes, ev0, ev1, single_node;
  const DevEvent* events;
  const int32_t *ent_node, *ent_link, *ent_flags, *ent_gslot, *up_ptr, *up_pos;
  const double *C1, *C2;      // [M][nNodes], link indexed
  const double* qout_hist;    // [nEvents][M][E]
  double* qtr_hist;           // [RS][M][E]
  double *qTIN, *qTR;         // [M][2][nNodes]
  double *qMod, *qOUT;        // [M][nNodes]
  double* qmod_g;             // [nEvents][M][nGslots]
  int32_t nGslots;
};

// L11_routing (mo_mrm_routing.f90:428-478) for all links of one level over a block of events
__global__ void route_level_kernel(const LevelArgs a) {
  const int p = a.p0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.p1) return;
  const int m = blockIdx.y;
  const int node = a.ent_node[p], flags = a.ent_flags[p], gslot = a.ent_gslot[p];
  const int u0 = a.up_ptr[p], u1 = a.up_ptr[p + 1];
  const bool is_link = flags & kEntLink;
  double c1 = 0.0, c2 = 0.0;
  if (is_link) {
    const int link = a.ent_link[p];
    c1 = a.C1[(size_t)m * a.nNodes + link];
    c2 = a.C2[(size_t)m * a.nNodes + link];
  }
  double* tin = a.qTIN + (size_t)m * 2 * a.nNodes;
  double* tr = a.qTR + (size_t)m * 2 * a.nNodes;
  double qtin1 = tin[node], qtr1 = tr[node];  // IT1 slot
  double qmod = 0.0, qout = 0.0;
  for (int ev = a.ev0; ev < a.ev1; ++ev) {
    const DevEvent e = a.events[ev];
    qout = a.qout_hist[hidx(ev, a.M, a.E, m, p)];
    if (a.single_node) {  // nNodes == 1: L11_Qmod = L11_qOUT (mo_mrm_routing.f90:284)
      qmod = qout;
    } else {
      double acc = 0.0;
      for (int s = 0; s < e.rout_loop; ++s) {
        const int rs = e.rs_first + s;
        double qin = 0.0;
        for (int u = u0; u < u1; ++u) qin = qin + a.qtr_hist[hidx(rs, a.M, a.E, m, a.up_pos[u])];
        if (flags & kEntAddQout) qin = qin + qout;
        if (is_link) {
          double q = qtr1 + c1 * (qtin1 - qtr1) + c2 * (qin - qtin1);
          if (flags & kEntZeroOut) q = 0.0;
          a.qtr_hist[hidx(rs, a.M, a.E, m, p)] = q;
          qtr1 = q;
        }
        qtin1 = qin;
        acc = acc + qin;
      }
      qmod = acc / (double)e.rout_loop;
    }
    if (gslot >= 0) a.qmod_g[((size_t)ev * a.M + m) * a.nGslots + gslot] = qmod;
  }
  if (a.ev1 > a.ev0) {
    tin[node] = qtin1;
    tin[a.nNodes + node] = qtin1;
    if (is_link) {
      tr[node] = qtr1;
      tr[a.nNodes + node] = qtr1;
    }
    a.qMod[(size_t)m * a.nNodes + node] = qmod;
    a.qOUT[(size_t)m * a.nNodes + node] = qout;
  }
}

// ---- fast path: every event is one routing step (rout_loop == 1, the usual case) ----------
struct FastArgs {
  int32_t p0, p1;          // entry range of the level (level kernel)
  int32_t ev0, ev1;        // event range [ev0, ev1) of the block handled by this launch
  int32_t E, M, nNodes, nEvents, nGslots;
  int32_t tail_level, nLevels;
  const EntMeta* meta;
  const int32_t *up_ptr, *up_pos, *lvl_ptr;
  const DevEvent* events;
  const double *C1, *C2;
  const double* qout_hist;  // tiled, see hidx()
  double* qtr_hist;         // tiled
  double *qTIN, *qTR, *qMod, *qOUT, *qmod_g;
};

__device__ __forceinline__ void load_tile(const double* src, double (&v)[kHistTile]) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
  for (int d = 0; d < kHistTile / 2; ++d) {
    const double2 t = s2[d];
    v[2 * d] = t.x;
    v[2 * d + 1] = t.y;
  }
}
__device__ __forceinline__ void load_tile_cg(const double* src, double (&v)[kHistTile]) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
  for (int d = 0; d < kHistTile / 2; ++d) {
    const double2 t = __ldcg(s2 + d);
    v[2 * d] = t.x;
    v[2 * d + 1] = t.y;
  }
}

// sum of the upstream links' routed outflow for the 8 events of a tile, added in netPerm
// order starting from 0 (mo_mrm_routing.f90:428,457); whole 64-byte runs per link
template <bool CG>
__device__ __forceinline__ void gather_upstream(const FastArgs& a, const EntMeta& em, int nup,
                                                int m, int p, int t0, double (&qin)[kHistTile]) {
  double t[kMetaUps][kHistTile];
#pragma unroll
  for (int u = 0; u < kMetaUps; ++u) {
    if (u < nup) {
      const double* src = a.qtr_hist + hidx(t0, a.M, a.E, m, em.up[u]);
      if (CG) load_tile_cg(src, t[u]); else load_tile(src, t[u]);
    }
  }
#pragma unroll
  for (int d = 0; d < kHistTile; ++d) qin[d] = 0.0;
#pragma unroll
  for (int u = 0; u < kMetaUps; ++u) {
    if (u < nup) {
#pragma unroll
      for (int d = 0; d < kHistTile; ++d) qin[d] = qin[d] + t[u][d];
    }
  }
  if (nup > kMetaUps) {  // rare: more than four inflowing links
    const int u0 = a.up_ptr[p];
    for (int u = kMetaUps; u < nup; ++u) {
      double x[kHistTile];
      load_tile_cg(a.qtr_hist + hidx(t0, a.M, a.E, m, a.up_pos[u0 + u]), x);
#pragma unroll
      for (int d = 0; d < kHistTile; ++d) qin[d] = qin[d] + x[d];
    }
  }
}

// the events [e_lo, e_hi) of one history tile for one entry; shared by both fast kernels
struct TileState {
  double qtin1, qtr1, qout;
};
__device__ __forceinline__ void route_tile(const FastArgs& a, const EntMeta& em, int m, int tile0,
                                           int e_lo, int e_hi, double c1, double c2,
                                           const double (&qo)[kHistTile],
                                           const double (&qup)[kHistTile],
                                           double (&qr)[kHistTile], TileState& st) {
  const bool is_link = em.flags & kEntLink;
#pragma unroll
  for (int d = 0; d < kHistTile; ++d) {
    const int ev = tile0 + d;
    qr[d] = 0.0;
    if (ev >= e_lo && ev < e_hi) {
      double qin = qup[d];
      st.qout = qo[d];
      if (em.flags & kEntAddQout) qin = qin + st.qout;
      if (is_link) {
        double q = st.qtr1 + c1 * (st.qtin1 - st.qtr1) + c2 * (qin - st.qtin1);
        if (em.flags & kEntZeroOut) q = 0.0;
        qr[d] = q;
        st.qtr1 = q;
      }
      st.qtin1 = qin;
      if (em.gslot >= 0) a.qmod_g[((size_t)ev * a.M + m) * a.nGslots + em.gslot] = qin;
    }
  }
}

// One wide network level: a thread owns one (entry, member), keeps qTIN/qTR in registers and
// walks through the events [ev0, ev1) of the block, one 8-event history tile (64-byte runs of
// its own runoff and of its first two upstream links, 128-bit loads) at a time.
__global__ void __launch_bounds__(128, 4) route_level_fast_kernel(const FastArgs a) {
  const int p = a.p0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.p1) return;
  const int m = blockIdx.y;
  const EntMeta em = a.meta[p];
  const bool is_link = em.flags & kEntLink;
  const int nup = em.flags >> 8;
  double c1 = 0.0, c2 = 0.0;
  if (is_link) {
    c1 = a.C1[(size_t)m * a.nNodes + em.link];
    c2 = a.C2[(size_t)m * a.nNodes + em.link];
  }
  double* tin = a.qTIN + (size_t)m * 2 * a.nNodes;
  double* tr = a.qTR + (size_t)m * 2 * a.nNodes;
  TileState st{tin[em.node], tr[em.node], 0.0};
  for (int t0 = a.ev0 & ~(kHistTile - 1); t0 < a.ev1; t0 += kHistTile) {
    double qo[kHistTile], qup[kHistTile], qr[kHistTile];
    load_tile(a.qout_hist + hidx(t0, a.M, a.E, m, p), qo);
    gather_upstream<false>(a, em, nup, m, p, t0, qup);
    double* own = a.qtr_hist + hidx(t0, a.M, a.E, m, p);
    // a tile shared with the previous launch (land-cover scene change inside the tile)
    // keeps the outflows that launch wrote
    const bool partial = t0 < a.ev0;
    double keep[kHistTile];
    if (partial && is_link) load_tile(own, keep);
    route_tile(a, em, m, t0, a.ev0, a.ev1, c1, c2, qo, qup, qr, st);
    if (is_link) {
      if (partial) {
#pragma unroll
        for (int d = 0; d < kHistTile; ++d)
          if (t0 + d < a.ev0) qr[d] = keep[d];
      }
      double2* dst = reinterpret_cast<double2*>(own);
#pragma unroll
      for (int d = 0; d < kHistTile / 2; ++d) dst[d] = make_double2(qr[2 * d], qr[2 * d + 1]);
    }
  }
  tin[em.node] = st.qtin1;
  tin[a.nNodes + em.node] = st.qtin1;
  if (is_link) {
    tr[em.node] = st.qtr1;
    tr[a.nNodes + em.node] = st.qtr1;
  }
  a.qMod[(size_t)m * a.nNodes + em.node] = st.qtin1;  // rout_loop == 1: qMod = qTIN(:, IT)
  a.qOUT[(size_t)m * a.nNodes + em.node] = st.qout;
}

// The narrow, deep part of the network (levels >= tail_level, every level at most
// kTailCluster * kTailThreads entries wide): instead of one launch per level, one thread-block
// cluster per member walks down the levels with a hardware cluster barrier between them.
// Thread i of the cluster owns entry i of the current level for all events of the range
// (state in registers, exactly like route_level_fast_kernel); what the next level needs from
// this one goes through L2 (.cg stores / loads) and is ordered by the barrier's
// release/acquire.  The per-level cost drops from a kernel launch to a cluster barrier.
__global__ void __cluster_dims__(kTailCluster, 1, 1) __launch_bounds__(kTailThreads, 2)
    route_tail_kernel(const FastArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ int32_t s_lvl[];  // lvl_ptr of the tail levels (or empty: read from L2)
  const int m = blockIdx.x / kTailCluster;
  const int ctid = (blockIdx.x % kTailCluster) * kTailThreads + threadIdx.x;
  const int nLv = a.nLevels - a.tail_level;
  const bool use_smem = a.p0 != 0;  // p0 doubles as "lvl_ptr fits in shared memory"
  if (use_smem)
    for (int i = threadIdx.x; i <= nLv; i += kTailThreads) s_lvl[i] = a.lvl_ptr[a.tail_level + i];
  __syncthreads();
  double* tin = a.qTIN + (size_t)m * 2 * a.nNodes;
  double* tr = a.qTR + (size_t)m * 2 * a.nNodes;
  for (int l = 0; l < nLv; ++l) {
    const int pb = use_smem ? s_lvl[l] : a.lvl_ptr[a.tail_level + l];
    const int pe = use_smem ? s_lvl[l + 1] : a.lvl_ptr[a.tail_level + l + 1];
    const int p = pb + ctid;
    if (p < pe) {
      const EntMeta em = a.meta[p];
      const bool is_link = em.flags & kEntLink;
      const int nup = em.flags >> 8;
      double c1 = 0.0, c2 = 0.0;
      if (is_link) {
        c1 = a.C1[(size_t)m * a.nNodes + em.link];
        c2 = a.C2[(size_t)m * a.nNodes + em.link];
      }
      TileState st{tin[em.node], tr[em.node], 0.0};
      for (int t0 = a.ev0 & ~(kHistTile - 1); t0 < a.ev1; t0 += kHistTile) {
        double qo[kHistTile], qup[kHistTile], qr[kHistTile];
        load_tile(a.qout_hist + hidx(t0, a.M, a.E, m, p), qo);
        gather_upstream<true>(a, em, nup, m, p, t0, qup);
        double* own = a.qtr_hist + hidx(t0, a.M, a.E, m, p);
        const bool partial = t0 < a.ev0;
        double keep[kHistTile];
        if (partial && is_link) load_tile_cg(own, keep);
        route_tile(a, em, m, t0, a.ev0, a.ev1, c1, c2, qo, qup, qr, st);
        if (is_link) {
          if (partial) {
#pragma unroll
            for (int d = 0; d < kHistTile; ++d)
              if (t0 + d < a.ev0) qr[d] = keep[d];
          }
          double2* dst = reinterpret_cast<double2*>(own);
#pragma unroll
          for (int d = 0; d < kHistTile / 2; ++d)
            __stcg(dst + d, make_double2(qr[2 * d], qr[2 * d + 1]));
        }
      }
      tin[em.node] = st.qtin1;
      tin[a.nNodes + em.node] = st.qtin1;
      if (is_link) {
        tr[em.node] = st.qtr1;
        tr[a.nNodes + em.node] = st.qtr1;
      }
      a.qMod[(size_t)m * a.nNodes + em.node] = st.qtin1;
      a.qOUT[(size_t)m * a.nNodes + em.node] = st.qout;
    }
    cluster.sync();
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
  std::vector<int32_t> rank((size_t)nLinks), link_of_node((size_t)nNodes, -1);
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
  }
  // upstream links of every node, in netPerm order
  std::vector<std::vector<int32_t>> up((size_t)nNodes);
  for (int k = 0; k < nLinks; ++k) {
    const int i = net->netPerm[k] - 1;
    up[(size_t)net->toN[i] - 1].push_back(i);
  }
  // level of every node = longest chain of links above it; netPerm is a topological order
  std::vector<int32_t> level((size_t)nNodes, 0);
  for (int k = 0; k < nLinks; ++k) {
    const int i = net->netPerm[k] - 1;
    const int f = net->fromN[i] - 1, t = net->toN[i] - 1;
    for (int j : up[(size_t)f])
      MHM_REQUIRE(rank[(size_t)j] < k, "set_network: netPerm is not a topological order");
    level[(size_t)t] = std::max(level[(size_t)t], level[(size_t)f] + 1);
  }
  // entries: all links (key: level, rank) then outlet nodes (key: level, nLinks + node)
  const int E = nNodes;
  std::vector<int32_t> ent((size_t)E);
  for (int nd = 0; nd < nNodes; ++nd) ent[(size_t)nd] = nd;
  auto key = [&](int nd) {
    const int l = link_of_node[(size_t)nd];
    return l >= 0 ? rank[(size_t)l] : nLinks + nd;
  };
  std::sort(ent.begin(), ent.end(), [&](int x, int y) {
    if (level[(size_t)x] != level[(size_t)y]) return level[(size_t)x] < level[(size_t)y];
    return key(x) < key(y);
  });
  std::vector<int32_t> pos_of_node((size_t)nNodes);
  for (int p = 0; p < E; ++p) pos_of_node[(size_t)ent[(size_t)p]] = p;
  rt->E = E;
  rt->lvl_ptr.clear();
  for (int p = 0; p < E; ++p)
    if (p == 0 || level[(size_t)ent[(size_t)p]] != level[(size_t)ent[(size_t)p - 1]])
      rt->lvl_ptr.push_back(p);
  rt->lvl_ptr.push_back(E);

  const int last_sink = nLinks > 0 ? net->toN[net->netPerm[nLinks - 1] - 1] - 1 : -1;
  std::vector<int32_t> ent_link((size_t)E), ent_flags((size_t)E), ent_gslot((size_t)E, -1),
      up_ptr((size_t)E + 1, 0), up_pos;
  up_pos.reserve((size_t)nLinks);
  for (int p = 0; p < E; ++p) {
    const int nd = ent[(size_t)p], l = link_of_node[(size_t)nd];
    int fl = 0;
    if (l >= 0) {
      fl |= kEntLink | kEntAddQout;  // mo_mrm_routing.f90:441
      for (int g = 0; g < net->nInflowGauges; ++g)  // :447-452
        if (net->toN[l] == net->InflowGaugeNodeList[g] && !net->InflowGaugeHeadwater[g])
          fl |= kEntZeroOut;
    } else if (nd == last_sink) {
      fl |= kEntAddQout;  // :466-467: only the last link's sink adds its own runoff
    }
    for (int g = 0; g < net->nInflowGauges; ++g)
      if (net->InflowGaugeNodeList[g] - 1 == nd) fl |= kEntInflow;
    ent_link[(size_t)p] = l >= 0 ? l : 0;
    ent_flags[(size_t)p] = fl;
    for (int j : up[(size_t)nd]) up_pos.push_back(pos_of_node[(size_t)net->fromN[j] - 1]);
    up_ptr[(size_t)p + 1] = (int32_t)up_pos.size();
  }
  // gauge slots: distinct gauge nodes
  std::vector<int32_t> gcol((size_t)net->nGauges), gslot((size_t)net->nGauges);
  rt->nGslots = 0;
  for (int g = 0; g < net->nGauges; ++g) {
    const int nd = net->gaugeNodeList[g] - 1;
    MHM_REQUIRE(nd >= 0 && nd < nNodes, "set_network: gauge node %d outside 1..%d", nd + 1, nNodes);
    MHM_REQUIRE(net->gaugeIndexList[g] >= 1 && net->gaugeIndexList[g] <= net->nGaugesTotal,
                "set_network: gaugeIndexList(%d) outside 1..nGaugesTotal", g + 1);
    int32_t& s = ent_gslot[(size_t)pos_of_node[(size_t)nd]];
    if (s < 0) s = rt->nGslots++;
    gslot[(size_t)g] = s;
    gcol[(size_t)g] = net->gaugeIndexList[g] - 1;
  }
  std::vector<EntMeta> meta((size_t)E);
  for (int p = 0; p < E; ++p) {
    EntMeta& em = meta[(size_t)p];
    em.node = ent[(size_t)p];
    em.link = ent_link[(size_t)p];
    em.flags = ent_flags[(size_t)p];
    em.gslot = ent_gslot[(size_t)p];
    const int nup = up_ptr[(size_t)p + 1] - up_ptr[(size_t)p];
    em.flags |= nup << 8;
    for (int u = 0; u < kMetaUps; ++u)
      em.up[u] = u < nup ? up_pos[(size_t)up_ptr[(size_t)p] + u] : 0;
  }
  // levels from tail_level on are all narrower than kTailMaxLevel
  const int nLv = (int)rt->lvl_ptr.size() - 1;
  rt->tail_level = nLv;
  int tail_max = kTailMaxLevel;
  if (const char* e = getenv("MHM_CUDA_TAIL_MAX")) tail_max = atoi(e);
  tail_max = std::min(tail_max, kTailCluster * kTailThreads);
  for (int l = nLv - 1; l >= 0; --l) {
    if (rt->lvl_ptr[(size_t)l + 1] - rt->lvl_ptr[(size_t)l] > tail_max) break;
    rt->tail_level = l;
  }
  if (nLv - rt->tail_level < 4) rt->tail_level = nLv;  // not worth a wavefront
  if (const char* e = getenv("MHM_CUDA_NO_TAIL")) if (e[0] == '1') rt->tail_level = nLv;
  cudaStream_t st = ctx->stream;
  if (int rc = upload(&rt->meta, meta, st)) return rc;
  if (int rc = upload(&rt->d_lvl_ptr, rt->lvl_ptr, st)) return rc;
  if (int rc = upload(&rt->ent_node, ent, st)) return rc;
  if (int rc = upload(&rt->ent_link, ent_link, st)) return rc;
  if (int rc = upload(&rt->ent_flags, ent_flags, st)) return rc;
  if (int rc = upload(&rt->ent_gslot, ent_gslot, st)) return rc;
  if (int rc = upload(&rt->up_ptr, up_ptr, st)) return rc;
  if (int rc = upload(&rt->up_pos, up_pos, st)) return rc;
  if (int rc = upload(&rt->d_gauge_col, gcol, st)) return rc;
  if (int rc = upload(&rt->d_gauge_slot, gslot, st)) return rc;

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
  {
    rt->bijective = false;
    if (n1 == nNodes) {
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
        std::vector<double> ca((size_t)n1);
        for (int k = 0; k < n1; ++k) {
          const int nd = node_of_cell[(size_t)k];
          ce[(size_t)k] = pos_of_node[(size_t)nd];
          ca[(size_t)k] = rt->map_flag ? net->L1_areaCell[k] : net->L11_areaCell[nd];
        }
        if (int rc = upload(&rt->d_cell_entry, ce, st)) return rc;
        if (int rc = upload(&rt->d_cell_area, ca, st)) return rc;
        rt->bijective = true;
      }
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
static int ensure_c1c2(mhm_cuda_context* ctx, Routing* rt, int yId, double timestep_rout) {
  if (rt->rout_case != 1 || rt->nNodes <= 1) return 0;
  for (int m = 0; m < rt->M; ++m) {
    if (rt->param5[(size_t)m].empty()) continue;  // C1/C2 supplied (read_states)
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
// aligned mode: one event per model step, one routing step per event, one cell per node
static bool routing_aligned(const Domain* d, const Routing* rt) {
  return rt->bijective && rt->nNodes > 1 && !routing_accumulates(d, rt) &&
         routing_rout_loop(d, rt) == 1;
}

struct Segment {
  int32_t ev0, ev1, yId;
};

// route the events of one block.  `aligned`: one cell per node and one model step per event,
// so the qOUT tiles are built by the coalesced per-cell kernel.
static int run_events(mhm_cuda_context* ctx, Domain* d, Routing* rt, std::vector<DevEvent>& ev,
                      const std::vector<Segment>& segs, const std::vector<double>& inflow_val,
                      const double* runoff_hist, bool aligned, double timestep_rout) {
  if (ev.empty()) return 0;
  cudaStream_t st = ctx->stream;
  const int nEv = (int)ev.size(), M = rt->M, E = rt->E;
  int RS = 0;
  bool fast = rt->nNodes > 1;
  for (auto& e : ev) {
    e.rs_first = RS;
    RS += e.rout_loop;
    fast = fast && e.rout_loop == 1;
  }
  if (getenv("MHM_CUDA_GENERIC_ROUTING")) fast = false;
  if (int rc = ensure(&rt->d_events, &rt->ev_cap, (size_t)nEv, st)) return rc;
  if (int rc = ensure(&rt->qout_hist, &rt->qout_cap, hist_size(nEv, M, E), st)) return rc;
  if (int rc = ensure(&rt->qtr_hist, &rt->qtr_cap, hist_size(RS, M, E), st)) return rc;
  if (int rc = ensure(&rt->qmod_g, &rt->qmodg_cap, (size_t)nEv * M * std::max(1, rt->nGslots), st))
    return rc;
  if (int rc = ensure(&rt->d_inflow_val, &rt->inflow_cap, std::max<size_t>(1, inflow_val.size()), st))
    return rc;
  MHM_CUDA_OK(cudaMemcpyAsync(rt->d_events, ev.data(), (size_t)nEv * sizeof(DevEvent),
                              cudaMemcpyHostToDevice, st));
  if (!inflow_val.empty())
    MHM_CUDA_OK(cudaMemcpyAsync(rt->d_inflow_val, inflow_val.data(),
                                inflow_val.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  MHM_CUDA_OK(cudaStreamSynchronize(st));  // host vectors go out of scope in the caller

  ctx->stat_begin(kStatRouting);
  int64_t launched = 0;
  const int tiles = (nEv + kHistTile - 1) / kHistTile;
  if (aligned) {
    QoutCellArgs qc{};
    qc.nCells1 = rt->nCells1;
    qc.E = E;
    qc.M = M;
    qc.map_flag = rt->map_flag;
    qc.nInflowGauges = rt->nInflowGauges;
    qc.nInflowTotal = rt->nInflowTotal;
    qc.runoff_hist = runoff_hist;
    qc.cell_entry = rt->d_cell_entry;
    qc.cell_area = rt->d_cell_area;
    qc.meta = rt->meta;
    qc.inflow_node = rt->d_inflow_node;
    qc.inflow_index = rt->d_inflow_index;
    qc.inflow_head = rt->d_inflow_head;
    for (int t0 = 0; t0 < tiles; t0 += 32768) {
      qc.events = rt->d_events + (size_t)t0 * kHistTile;
      qc.nEvents = nEv - t0 * kHistTile;
      qc.inflow_val = rt->d_inflow_val + (size_t)t0 * kHistTile * rt->nInflowTotal;
      qc.qout_hist = rt->qout_hist + hidx(t0 * kHistTile, M, E, 0, 0);
      const int nt = std::min(32768, tiles - t0);
      qout_cell_kernel<<<dim3((rt->nCells1 + 127) / 128, M, nt), 128, 0, st>>>(qc);
      ++launched;
    }
    MHM_CUDA_OK(cudaGetLastError());
  } else {
    QoutArgs qa{};
    qa.nCells1 = rt->nCells1;
    qa.nNodes = rt->nNodes;
    qa.E = E;
    qa.M = M;
    qa.nEvents = nEv;
    qa.map_flag = rt->map_flag;
    qa.nInflowGauges = rt->nInflowGauges;
    qa.nInflowTotal = rt->nInflowTotal;
    qa.runoff_hist = runoff_hist;
    qa.carry = rt->carry;
    qa.ent_node = rt->ent_node;
    qa.cell_ptr = rt->cell_ptr;
    qa.cell_idx = rt->cell_idx;
    qa.L11_L1_Id = rt->L11_L1_Id;
    qa.L1_area = rt->L1_area;
    qa.L11_area = rt->L11_area;
    qa.inflow_node = rt->d_inflow_node;
    qa.inflow_index = rt->d_inflow_index;
    qa.inflow_head = rt->d_inflow_head;
    for (int t0 = 0; t0 < tiles; t0 += 32768) {  // gridDim.z limit
      // the kernel indexes events and history tiles from `tile`; shift all three views
      qa.events = rt->d_events + (size_t)t0 * kHistTile;
      qa.nEvents = nEv - t0 * kHistTile;
      qa.inflow_val = rt->d_inflow_val + (size_t)t0 * kHistTile * rt->nInflowTotal;
      qa.qout_hist = rt->qout_hist + hidx(t0 * kHistTile, M, E, 0, 0);
      const int nt = std::min(32768, tiles - t0);
      qout_kernel<<<dim3((E + 127) / 128, M, nt), 128, 0, st>>>(qa);
      ++launched;
    }
    MHM_CUDA_OK(cudaGetLastError());
  }

  for (const Segment& sg : segs) {
    if (int rc = ensure_c1c2(ctx, rt, sg.yId, timestep_rout)) return rc;
    if (fast) {
      FastArgs fa{};
      fa.ev0 = sg.ev0;
      fa.ev1 = sg.ev1;
      fa.E = E;
      fa.M = M;
      fa.nNodes = rt->nNodes;
      fa.nEvents = nEv;
      fa.nGslots = std::max(1, rt->nGslots);
      fa.tail_level = rt->tail_level;
      fa.nLevels = (int)rt->lvl_ptr.size() - 1;
      fa.meta = rt->meta;
      fa.up_ptr = rt->up_ptr;
      fa.up_pos = rt->up_pos;
      fa.lvl_ptr = rt->d_lvl_ptr;
      fa.events = rt->d_events;
      fa.C1 = rt->C1;
      fa.C2 = rt->C2;
      fa.qout_hist = rt->qout_hist;
      fa.qtr_hist = rt->qtr_hist;
      fa.qTIN = rt->qTIN;
      fa.qTR = rt->qTR;
      fa.qMod = rt->qMod;
      fa.qOUT = rt->qOUT;
      fa.qmod_g = rt->qmod_g;
      for (int l = 0; l < rt->tail_level; ++l) {
        fa.p0 = rt->lvl_ptr[(size_t)l];
        fa.p1 = rt->lvl_ptr[(size_t)l + 1];
        const int cnt = fa.p1 - fa.p0;
        route_level_fast_kernel<<<dim3((cnt + 127) / 128, M), 128, 0, st>>>(fa);
        ++launched;
      }
      if (rt->tail_level < fa.nLevels) {
        const int nLvTail = fa.nLevels - rt->tail_level;
        const size_t smem = (size_t)(nLvTail + 1) * sizeof(int32_t);
        fa.p0 = smem <= 40 * 1024 ? 1 : 0;  // lvl_ptr of the tail staged in shared memory
        route_tail_kernel<<<dim3(M * kTailCluster), kTailThreads, fa.p0 ? smem : 0, st>>>(fa);
        ++launched;
      }
    } else {
      LevelArgs la{};
      la.E = E;
      la.M = M;
      la.nNodes = rt->nNodes;
      la.ev0 = sg.ev0;
      la.ev1 = sg.ev1;
      la.single_node = rt->nNodes <= 1;
      la.events = rt->d_events;
      la.ent_node = rt->ent_node;
      la.ent_link = rt->ent_link;
      la.ent_flags = rt->ent_flags;
      la.ent_gslot = rt->ent_gslot;
      la.up_ptr = rt->up_ptr;
      la.up_pos = rt->up_pos;
      la.C1 = rt->C1;
      la.C2 = rt->C2;
      la.qout_hist = rt->qout_hist;
      la.qtr_hist = rt->qtr_hist;
      la.qTIN = rt->qTIN;
      la.qTR = rt->qTR;
      la.qMod = rt->qMod;
      la.qOUT = rt->qOUT;
      la.qmod_g = rt->qmod_g;
      la.nGslots = std::max(1, rt->nGslots);
      for (size_t l = 0; l + 1 < rt->lvl_ptr.size(); ++l) {
        la.p0 = rt->lvl_ptr[l];
        la.p1 = rt->lvl_ptr[l + 1];
        const int cnt = la.p1 - la.p0;
        const int threads = cnt >= 128 ? 128 : 32;
        route_level_kernel<<<dim3((cnt + threads - 1) / threads, M), threads, 0, st>>>(la);
        ++launched;
      }
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
  (void)d;
  return 0;
}

// routing of model steps tt_first .. tt_first+n_steps-1 whose total runoff is in
// d->runoff_hist (or, aligned mode, already in the qOUT tiles); restates the schedule of
// mo_mhm_interface_run.f90:460-612
int routing_run_block(mhm_cuda_context* ctx, Domain* d, int32_t tt_first, int32_t n_steps) {
  Routing* rt = d->rt;
  MHM_REQUIRE(rt->nTimeSteps == d->axis.nTimeSteps && rt->gauge_hist,
              "routing: time axis changed after mrm_cuda_set_network");
  const int nTstepDay = 24 / d->cfg.timestep_h;
  const int nT = d->axis.nTimeSteps;
  if (rt->inflow_acc.size() != (size_t)rt->nInflowTotal) rt->inflow_acc.assign((size_t)rt->nInflowTotal, 0.0);
  auto inflow_at = [&](int g, int day) -> double {  // InflowGauge%Q(day, g), 1-based day
    if (rt->inflowQ.empty()) return 0.0;
    return rt->inflowQ[(size_t)g * rt->nDays + (size_t)(day - 1)];
  };
  const bool accumulates = routing_accumulates(d, rt);
  const bool aligned = routing_aligned(d, rt) && !getenv("MHM_CUDA_NO_ALIGNED");
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
        long rl = fortran_nint(1.0 / fin);
        e.rout_loop = (int32_t)(rl < 1 ? 1 : rl);
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
  if (int rc = run_events(ctx, d, rt, ev, segs, inflow_val, d->runoff_hist, aligned,
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
    rt->ssMax = mx;
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
  MHM_REQUIRE(rt->rout_case == 1 || rt->TSrout == 0.0 || rt->TSrout == TSrout,
              "set_c1c2: members of one domain must share L11_TSrout (%g vs %g)", rt->TSrout, TSrout);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t nn = (size_t)rt->nNodes;
  MHM_CUDA_OK(cudaMemcpyAsync(rt->C1 + (size_t)member * nn, C1, nn * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream));
  MHM_CUDA_OK(cudaMemcpyAsync(rt->C2 + (size_t)member * nn, C2, nn * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  rt->param5[(size_t)member].clear();
  if (rt->rout_case != 1) rt->TSrout = TSrout;
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
  return run_events(ctx, d, rt, ev, segs, inflow_val, src, false, (double)timestep_rout);
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
