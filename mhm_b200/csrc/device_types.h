// device_types.h -- argument blocks shared by the host API and the kernels.
#pragma once
#include <cstdint>

#include "../../include/mhm_cuda.h"

namespace mhm {

constexpr int kCellThreads = 128;
constexpr int kMaxHorizons = 8;
#ifndef MHM_IDX_INLINE
#define MHM_IDX_INLINE 128
#endif
constexpr int kIdxInline = MHM_IDX_INLINE;  // model steps per cell-kernel launch (calendar rides in the arguments)

// per-step calendar indices, 16 bytes (one LDG.128, warp-uniform)
struct alignas(16) StepIdx {
  int32_t iMeteoTS;  // 1-based meteo time index (mo_meteo_handler.f90:607)
  int16_t yId;       // 1-based land-cover scene
  int16_t iLAI;      // 1-based LAI step
  int16_t doy;
  int16_t year;
  int8_t month;      // 1..12
  int8_t hour;       // 0..23
  int8_t isday;      // hour > 6 .and. hour <= 18
  int8_t flags;      // after the step's date increment: 1 new day, 2 new month, 4 new year
};
static_assert(sizeof(StepIdx) == 16, "StepIdx must be 16 bytes");

struct MeteoTables {
  double fday_prec[12], fnight_prec[12], fday_pet[12], fnight_pet[12], fday_temp[12],
      fnight_temp[12], evap_coeff[12], inv_evap_coeff[12];
};

struct CellArgs {
  int32_t nCells, nMembers, nSteps, tt_first;
  int32_t nLC, nLAI;
  int32_t soil_case, pet_case, is_hourly, read_weights, read_states, write_fluxes;
  // the launch's steps share yId / iLAI / month and read consecutive meteo rows (set by the host,
  // which cuts launches where the calendar turns): the specialised kernels then drop the per-step
  // calendar work
  int32_t uniform_calendar;
  // forcing rows may go through the TMA unit (16-byte aligned rows; decided by the host)
  int32_t forcing_tma;
  double nTstepDay_dp, c2TSTu;
  StepIdx idx_in[kIdxInline];         // calendar of the launch's steps: constant-bank loads
  const double* met[MHM_M_COUNT];     // device, [rows][nCells]
  long long met_first[MHM_M_COUNT];   // iMeteoTS of row 0
  const double *w_pre, *w_temp, *w_pet;  // device, [24][12][nCells]
  const double* P[MHM_P_COUNT];       // device, [member][dim3][dim2][nCells]
  double* S[MHM_S_COUNT];             // device, [member][(nH)][nCells]
  double* F[MHM_F_COUNT];             // device, [member][(nH)][nCells]
  double* runoff_hist;                // device, [nSteps][member][nCells] or null
  // routing input produced in place (one cell per node, one model step per routing event):
  // node runoff qOUT of L11_runoff_acc in the tiled layout [slot/8][member][lane][slot%8],
  // slot = step + cell_skew
  double* qout_hist;                  // null: not fused
  const int32_t* cell_lane;           // [nCells] routing lane of the cell's node
  const double* cell_area;            // [nCells] area factor (mo_mrm_pre_routing.f90:125/:141)
  const int8_t* cell_skew;            // [nCells] position of the cell's node in its routing segment (history slot shift)
  int32_t qout_step0, qout_E, qout_map_flag;
  double qout_tst, qout_scale;        // seconds per model step; 1000 / tst
  // gridded outputs (mo_write_fluxes_states.f90:283-438): bit v = outputFlxState(v), slots in
  // the order of mHM_updateDataset; out_acc [slot][member][nCells] is the open window
  uint32_t out_mask;
  int32_t out_first;                  // first step of the launch with tIndex_out > 0
  int32_t out_nslots;                 // slots of out_acc in use
  double* out_acc;
  int16_t out_yid[kIdxInline];        // land-cover scene the driver holds after each step (StepIdx::yId's type)
  // calibration aggregates (mo_mhm_interface_run.f90:745-861) and BFI sums (:630-636): every step
  // t >= out_first adds to the open dataSim column of bit 0 soil moisture, bit 1 evapotranspiration,
  // bit 2 total water storage; bit 3 adds baseflow / total runoff to bfi_acc [2][member][nCells]
  uint32_t agg_mask;
  int32_t agg_nhor_sm;                // nSoilHorizons_sm_input
  double* agg_col[3];                 // [member][nCells]
  double* bfi_acc;
  MeteoTables tab;
};

// kernel launchers (cell_kernel_strict.cu / cell_kernel_fast.cu)
int launch_cell_block_strict(const CellArgs& a, int nH, void* stream);
int launch_cell_block_fast(const CellArgs& a, int nH, void* stream);

// the reference's default output selection (mhm_outputs.nml): bits 1..16, 19, 20, 21 of out_mask
constexpr uint32_t kOutDefaultMask = 0xffffu << 1 | 1u << 19 | 1u << 20 | 1u << 21;

}  // namespace mhm
