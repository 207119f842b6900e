// cell_kernel.cuh -- fused fp64 meteo prologue + L1 cell cascade for sm_100a.
//
// One thread owns one (cell, member) pair for a whole block of model steps and keeps
// its 5 + nH states and its effective parameters in registers; forcing is streamed
// [meteo step][cell] (cell-contiguous, the reference's own L1_pre(nCells, nSteps)
// layout) with the next step's loads issued before the current step's arithmetic.
// Members of one cell tile are adjacent in blockIdx, so a forcing row is fetched from
// HBM once and served to the other members from L2.
//
// Restates (never copies) the arithmetic of
//   meteo/mo_meteo_handler.f90:1053-1119,1166-1201,1247-1281 (PET select, disaggregation)
//   meteo/mo_meteo_temporal_tools.f90:57,90-99,142-159
//   mHM/mo_pet.f90:109-114,170-172,267-270,338-353,397,439
//   mHM/mo_mhm.f90:448-499, mo_canopy_interc.f90:105-131, mo_snow_accum_melt.f90:117-156,
//   mo_soil_moisture.f90:179-286,353-361,431-444, mo_runoff.f90:120-152,204-210,271-272
// in the same operation order.  This file is compiled twice (csrc/Makefile):
//   MHM_FAST=0  -fmad=false : no FMA contraction, IEEE division, literal formulas
//   MHM_FAST=1  FMA allowed + algebraically equivalent shortcuts (<= 1e-9 relative)
#pragma once
#include <cfloat>
#include <cstdint>
#include <type_traits>

#include "device_types.h"
#include "fastmath.cuh"

namespace mhm {

#ifndef MHM_FAST
#define MHM_FAST 0
#endif
#ifndef MHM_CELL_SMEM_BLOCKS
#define MHM_CELL_SMEM_BLOCKS 5
#endif
#ifndef MHM_TABLES_GLOBAL
#define MHM_TABLES_GLOBAL 0
#endif
#ifndef MHM_PARAMS_SMEM
#define MHM_PARAMS_SMEM 1
#endif
#ifndef MHM_SELECT_FORM
#define MHM_SELECT_FORM 1
#endif
#ifndef MHM_CELL_PIPELINE
#define MHM_CELL_PIPELINE 1  // uniform-calendar launches: stage A of step t+1 beside stage B of step t
#endif
#ifndef MHM_CELL_PIPE3
#define MHM_CELL_PIPE3 0  // uniform-calendar launches: three steps in flight (0: two; measured 10 % faster on B200)
#endif
#ifndef MHM_RESV_POW_POLY
#define MHM_RESV_POW_POLY 0  // slow-interflow power without table look-ups (more fp64, no bank conflicts)
#endif
#ifndef MHM_POW_COMPACT
#define MHM_POW_COMPACT 1  // infiltration powers of a warp compacted through shared memory (0: per lane)
#endif
#ifndef MHM_CELL_MIN_BLOCKS
#define MHM_CELL_MIN_BLOCKS 4
#endif

#if MHM_FAST && defined(__CUDA_ARCH__)
#define kEps (fm::c_coef.eps)  // constant bank: one uniform load instead of two 32-bit moves
#else
constexpr double kEps = 2.220446049250313e-16;  // epsilon(1.0_dp), mo_common_constants.f90:25
#endif
constexpr double kTwoThird = 0.6666666666666666666666666666666666667;  // FORCES twothird_dp
constexpr double kPi = 3.141592653589793238462643383279502884197;
constexpr double kTwoPi = 6.283185307179586476925286766559005768394;
constexpr double kDeg2Rad = kPi / 180.0;
constexpr double kT0 = 273.15;
constexpr double kDaySecs = 86400.0;
constexpr double kYearDays = 365.0;
constexpr double kSolarConst = 1367.0;
constexpr double kSpecHeatET = 2.45e06;
constexpr double kPsychro = 0.0646;
constexpr double kCp0 = 1005.0;
constexpr double kRho0 = 1.225;
constexpr double kHarSamConst = 17.800;  // mo_mhm_constants.f90:36
constexpr double kDuffieDr = 0.0330, kDuffieDelta1 = 0.4090, kDuffieDelta2 = 1.3900;
constexpr double kTetensC1 = 0.6108, kTetensC2 = 17.270, kTetensC3 = 237.30;
constexpr double kSatPressureSlope1 = 4098.0;

// ---- PET, mHM/mo_pet.f90 ---------------------------------------------------------
__device__ __forceinline__ bool le_eps(double a, double b) {  // FORCES mo_utils::le
  if ((kEps * fabs(b) - fabs(a - b)) < 0.0) return a < b;
  return true;
}
__device__ __forceinline__ double sat_vap_pressure(double tavg) {  // :439
  return kTetensC1 * exp(kTetensC2 * tavg / (tavg + kTetensC3));
}
__device__ __forceinline__ double slope_satpressure(double tavg) {  // :397
  return kSatPressureSlope1 * sat_vap_pressure(tavg) / exp(2.0 * log(tavg + kTetensC3));
}
__device__ __forceinline__ double extraterr_rad_approx(int doy, double latitude) {  // :338-353
  double dr = 1.0 + kDuffieDr * cos(kTwoPi * doy / kYearDays);
  double delta = kDuffieDelta1 * sin(kTwoPi * doy / kYearDays - kDuffieDelta2);
  double arg = -tan(latitude) * tan(delta);
  if (arg < -1.0) arg = -1.0;
  if (arg > 1.0) arg = 1.0;
  double omega = acos(arg);
  return kDaySecs / kPi / kSpecHeatET * kSolarConst * dr *
         (omega * sin(latitude) * sin(delta) + cos(latitude) * cos(delta) * sin(omega));
}
__device__ __forceinline__ double pet_hargreaves(double coeff, double tavg, double tmax,
                                                 double tmin, double latitude, int doy) {  // :109-114
  double delta_temp = tmax - tmin;
  if (le_eps(delta_temp, 0.0) || le_eps(tavg, -kHarSamConst)) return 0.0;
  return coeff * extraterr_rad_approx(doy, kDeg2Rad * latitude) * (tavg + kHarSamConst) *
         sqrt(delta_temp);
}
__device__ __forceinline__ double pet_priestly(double alpha, double Rn, double tavg) {  // :170-172
  double delta = slope_satpressure(tavg);
  return alpha * delta / (kPsychro + delta) * (Rn * kDaySecs / kSpecHeatET);
}
__device__ __forceinline__ double pet_penman(double net_rad, double tavg, double avp,
                                             double ra, double rs) {  // :267-270, a_s = a_sh = 1
  const double a_s = 1.0, a_sh = 1.0;
  return kDaySecs / kSpecHeatET *
         (slope_satpressure(tavg) * net_rad +
          kRho0 * kCp0 * (sat_vap_pressure(tavg) - avp) * a_sh / ra) /
         (slope_satpressure(tavg) + kPsychro * a_sh / a_s * (1.0 + rs / ra));
}

// ---- per-thread parameter set ------------------------------------------------------
template <int NH>
struct CellParams {
  // land-cover-scene (yId) indexed
  double fSealed, alpha, ddinc, ddmax_c, ddnop_c, ddthr;  // ddthr = (ddmax_c-ddnop_c)/ddinc
  double k0r, k1r, k2r, kpr;                              // c2TSTu / k
  double tthr;
  double fRoots[NH], FC[NH], SAT[NH], EXPN[NH], WP[NH];
  // constant
  double karst, jarvis_c1, unsatThr, sealedThr;
  // LAI-step (iLAI) indexed
  double maxInter;
  double petFac;  // petLAIcorFactor (case -1) or fAsp (case 0, 1)
#if MHM_FAST
  double inv_maxInter, inv_sealedThr, inv_SAT[NH], inv_FCWP[NH];
#endif
};

// The same parameter set as a structure of arrays in shared memory (one column per thread of
// the CTA, conflict-free): frees ~60 registers per thread for a higher occupancy.
template <int NH>
struct CellParamsShared {
  double fSealed[kCellThreads], alpha[kCellThreads], ddinc[kCellThreads], ddmax_c[kCellThreads],
      ddnop_c[kCellThreads], ddthr[kCellThreads];
  double k0r[kCellThreads], k1r[kCellThreads], k2r[kCellThreads], kpr[kCellThreads];
  double tthr[kCellThreads];
  double fRoots[NH][kCellThreads], FC[NH][kCellThreads], SAT[NH][kCellThreads], EXPN[NH][kCellThreads],
      WP[NH][kCellThreads];
  double karst[kCellThreads], jarvis_c1[kCellThreads], unsatThr[kCellThreads], sealedThr[kCellThreads];
  double maxInter[kCellThreads];
  double petFac[kCellThreads];
  double inv_maxInter[kCellThreads], inv_sealedThr[kCellThreads], inv_SAT[NH][kCellThreads],
      inv_FCWP[NH][kCellThreads];
};
// one access syntax for both stores: a member is a double (registers) or a column (shared)
__device__ __forceinline__ double& pcol(double& x) { return x; }
__device__ __forceinline__ const double& pcol(const double& x) { return x; }
__device__ __forceinline__ double& pcol(double (&x)[kCellThreads]) { return x[threadIdx.x]; }
__device__ __forceinline__ const double& pcol(const double (&x)[kCellThreads]) { return x[threadIdx.x]; }
#define PX(m_) pcol(p.m_)
#define PH(m_, h_) pcol(p.m_[h_])
struct EmptyParams {};
template <bool FIRST, class A, class B>
__device__ __forceinline__ auto& pick_ref(A& a, B& b) {
  if constexpr (FIRST) return a;
  else return b;
}
// fast build, up to 2 soil horizons: parameters in shared memory (specialised variants: 15 pairs,
// the 4 read most often in registers -- 96 registers, 29 KB of shared memory, 5 CTAs/SM; measured
// on B200 in cell-steps/s: all pairs in shared memory at 80 registers / 6 CTAs 6.68e10, 4 pairs in
// registers at 96 / 5 CTAs 6.93e10, 12 pairs at 128 / 4 CTAs 6.73e10, all 15 at 128 / 4 6.35e10: the
// kernel sits where the shared-memory data pipe (98 % busy with every pair in shared memory) and
// the latency the resident warps can hide meet); otherwise in registers, 128 registers, 4 CTAs/SM
template <int NH>
struct ParamPlace {
  static constexpr bool shared = MHM_FAST && MHM_PARAMS_SMEM && NH <= 2;
  static constexpr int min_blocks = shared ? MHM_CELL_SMEM_BLOCKS : MHM_CELL_MIN_BLOCKS;
};

// ---- one model step for one cell ---------------------------------------------------
template <int NH>
struct CellStates {
  double inter, snowpack, sealed, unsat, sat;
  double sm[NH];
};

// Fluxes leave the step through an emitter the moment they are final, so that none of them
// has to stay in a register until the end of the time loop.  They are only stored for the
// last step of a block: the time loop is compiled twice, with EMIT = false for steps
// 1..n-1 (the emitter vanishes) and EMIT = true for the last one.
// fluxes of the step the gridded outputs are built from (kept only by the OUT kernels)
struct FluxCapture {
  double v[MHM_F_COUNT];
  double aet_soil[kMaxHorizons], infil[kMaxHorizons];
};

template <bool EMIT, bool OUT = false>
struct FluxEmitter {
  static constexpr bool kEmit = EMIT, kOut = OUT;
  double* const* F;
  size_t mc, n, mh;  // member*n + cell; nCells; (member*NH)*n + cell
  bool on;
  FluxCapture* cap;
  __device__ __forceinline__ void operator()(int id, double v) const {
    if (EMIT) {
      if (on) F[id][mc] = v;
    }
    if (OUT) cap->v[id] = v;
  }
  __device__ __forceinline__ void operator()(int id, int h, double v) const {
    if (EMIT) {
      if (on) F[id][mh + (size_t)h * n] = v;
    }
    if (OUT) (id == MHM_F_AETSOIL ? cap->aet_soil : cap->infil)[h] = v;
  }
};

// mHM_updateDataset (mo_write_fluxes_states.f90:326-436): add the step's value of every enabled
// output variable to the open window, slots in the order of the reference's `ii = ii + 1`
// blocks.  fS / fNS / sat_o belong to the land-cover scene the driver holds after the step's
// date increment (mo_mhm_interface_run.f90:623-628, 690-696).
// The open window's sums of a (cell, member) live in (dynamic) shared memory for the whole
// launch, one conflict-free column per thread ([slot][thread]; loaded before the launch's first
// step, written back after its last): the additions happen in the same order as with one
// read-modify-write of global memory per step and slot, without the trips to L2 (a per-thread
// local array does not help: the parameter store leaves the L1 too small to hold it).
template <int NH>
__device__ __forceinline__ void accumulate_outputs(const uint32_t mask, double* acc,
                                                   const FluxCapture& f, const CellStates<NH>& s,
                                                   const double fS, const double* sat_o) {
  const double fNS = 1.0 - fS;  // L1_fNotSealed, mo_mhm_interface_run.f90:238-239
  auto add = [&](double v) {
    *acc = *acc + v;  // OutputVariable%updateVariable, mo_nc_output.f90:140-149
    acc += kCellThreads;
  };
  auto on = [&](int v) { return (mask >> v) & 1u; };
  if (on(1)) add(s.inter);
  if (on(2)) add(s.snowpack);
  if (on(3)) {
#pragma unroll
    for (int h = 0; h < NH; ++h) add(s.sm[h]);
  }
  if (on(4)) {
#pragma unroll
    for (int h = 0; h < NH; ++h) add(s.sm[h] / sat_o[h]);
  }
  if (on(5)) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h) a = a + s.sm[h];
#pragma unroll
    for (int h = 0; h < NH; ++h) b = b + sat_o[h];
    add(a / b);
  }
  if (on(6)) add(s.sealed);
  if (on(7)) add(s.unsat);
  if (on(8)) add(s.sat);
  if (on(9)) add(f.v[MHM_F_PET_CALC]);
  if (on(10)) {
    double a = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h) a = a + f.aet_soil[h];
    const double t1 = a * fNS, t2 = f.v[MHM_F_AETSEALED] * fS;
    add((t1 + f.v[MHM_F_AETCANOPY]) + t2);
  }
  if (on(11)) add(f.v[MHM_F_TOTAL_RUNOFF]);
  if (on(12)) add(f.v[MHM_F_RUNOFFSEAL] * fS);
  if (on(13)) add(f.v[MHM_F_FASTRUNOFF] * fNS);
  if (on(14)) add(f.v[MHM_F_SLOWRUNOFF] * fNS);
  if (on(15)) add(f.v[MHM_F_BASEFLOW] * fNS);
  if (on(16)) add(f.v[MHM_F_PERCOL] * fNS);
  if (on(17)) {
#pragma unroll
    for (int h = 0; h < NH; ++h) add(f.infil[h] * fNS);
  }
  if (on(19)) {
#pragma unroll
    for (int h = 0; h < NH; ++h) add(f.aet_soil[h] * fNS);
  }
  if (on(20)) add(f.v[MHM_F_PREEFFECT]);
  if (on(21)) add(f.v[MHM_F_MELT]);
}

// The reference's default mhm_outputs.nml (variables 1-16 and 19-21: 16 + 3 nH fields) with the open
// window's sums in REGISTERS: the same additions in the same order, but no read-modify-write of
// shared memory per step and field (22 fields cost 88 shared-memory wavefronts per warp-step, as much
// as the whole cascade) and no run-time selection.  The kernel gives up occupancy for it (168
// registers, 3 CTAs/SM).
template <int NH>
constexpr int out_default_slots() { return 16 + 3 * NH; }
template <int NH>
__device__ __forceinline__ void accumulate_outputs_default(double (&acc)[16 + 3 * NH], const FluxCapture& f,
                                                           const CellStates<NH>& s, const double fS,
                                                           const double* sat_o) {
  const double fNS = 1.0 - fS;
  int k = 0;
  auto add = [&](double v) {
    acc[k] = acc[k] + v;  // OutputVariable%updateVariable, mo_nc_output.f90:140-149
    ++k;
  };
  add(s.inter);
  add(s.snowpack);
#pragma unroll
  for (int h = 0; h < NH; ++h) add(s.sm[h]);
#pragma unroll
  for (int h = 0; h < NH; ++h) add(s.sm[h] / sat_o[h]);
  {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h) a = a + s.sm[h];
#pragma unroll
    for (int h = 0; h < NH; ++h) b = b + sat_o[h];
    add(a / b);
  }
  add(s.sealed);
  add(s.unsat);
  add(s.sat);
  add(f.v[MHM_F_PET_CALC]);
  {
    double a = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h) a = a + f.aet_soil[h];
    const double t1 = a * fNS, t2 = f.v[MHM_F_AETSEALED] * fS;
    add((t1 + f.v[MHM_F_AETCANOPY]) + t2);
  }
  add(f.v[MHM_F_TOTAL_RUNOFF]);
  add(f.v[MHM_F_RUNOFFSEAL] * fS);
  add(f.v[MHM_F_FASTRUNOFF] * fNS);
  add(f.v[MHM_F_SLOWRUNOFF] * fNS);
  add(f.v[MHM_F_BASEFLOW] * fNS);
  add(f.v[MHM_F_PERCOL] * fNS);
#pragma unroll
  for (int h = 0; h < NH; ++h) add(f.aet_soil[h] * fNS);
  add(f.v[MHM_F_PREEFFECT]);
  add(f.v[MHM_F_MELT]);
}

// The same additions split for the software-pipelined launches (stage A of step t+1 runs beside
// stage B of step t): the slots fed by canopy / snow / sealed store are added right after stage A,
// the others after stage B of the same step -- every slot still receives its steps in time order.
template <int NH>
__device__ __forceinline__ void accumulate_default_a(double (&acc)[16 + 3 * NH], const FluxCapture& f,
                                                     const CellStates<NH>& s, const double pet_calc,
                                                     const double fS) {
  acc[0] = acc[0] + s.inter;
  acc[1] = acc[1] + s.snowpack;
  acc[3 + 2 * NH] = acc[3 + 2 * NH] + s.sealed;
  acc[6 + 2 * NH] = acc[6 + 2 * NH] + pet_calc;
  acc[9 + 2 * NH] = acc[9 + 2 * NH] + f.v[MHM_F_RUNOFFSEAL] * fS;
  acc[14 + 3 * NH] = acc[14 + 3 * NH] + f.v[MHM_F_PREEFFECT];
  acc[15 + 3 * NH] = acc[15 + 3 * NH] + f.v[MHM_F_MELT];
}
template <int NH>
__device__ __forceinline__ void accumulate_default_b(double (&acc)[16 + 3 * NH], const FluxCapture& f,
                                                     const CellStates<NH>& s, const double fS,
                                                     const double* sat_o, const double aet_canopy,
                                                     const double aet_sealed) {
  const double fNS = 1.0 - fS;
#pragma unroll
  for (int h = 0; h < NH; ++h) acc[2 + h] = acc[2 + h] + s.sm[h];
#pragma unroll
  for (int h = 0; h < NH; ++h) acc[2 + NH + h] = acc[2 + NH + h] + s.sm[h] / sat_o[h];
  {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h) a = a + s.sm[h];
#pragma unroll
    for (int h = 0; h < NH; ++h) b = b + sat_o[h];
    acc[2 + 2 * NH] = acc[2 + 2 * NH] + a / b;
  }
  acc[4 + 2 * NH] = acc[4 + 2 * NH] + s.unsat;
  acc[5 + 2 * NH] = acc[5 + 2 * NH] + s.sat;
  {
    double a = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h) a = a + f.aet_soil[h];
    const double t1 = a * fNS, t2 = aet_sealed * fS;
    acc[7 + 2 * NH] = acc[7 + 2 * NH] + ((t1 + aet_canopy) + t2);
  }
  acc[8 + 2 * NH] = acc[8 + 2 * NH] + f.v[MHM_F_TOTAL_RUNOFF];
  acc[10 + 2 * NH] = acc[10 + 2 * NH] + f.v[MHM_F_FASTRUNOFF] * fNS;
  acc[11 + 2 * NH] = acc[11 + 2 * NH] + f.v[MHM_F_SLOWRUNOFF] * fNS;
  acc[12 + 2 * NH] = acc[12 + 2 * NH] + f.v[MHM_F_BASEFLOW] * fNS;
  acc[13 + 2 * NH] = acc[13 + 2 * NH] + f.v[MHM_F_PERCOL] * fNS;
#pragma unroll
  for (int h = 0; h < NH; ++h) acc[14 + 2 * NH + h] = acc[14 + 2 * NH + h] + f.aet_soil[h] * fNS;
}

// mhm_interface_run_update_optisim (mo_mhm_interface_run.f90:776-857): the step's soil-moisture
// fraction of the first nhor_sm horizons, total evapotranspiration and total water storage are
// added to the open dataSim column (optidata_sim%add); BFI sums (:630-636) per cell, the host
// applies the cell areas.  Same scene (after the date increment) as the gridded outputs.
template <int NH>
__device__ __forceinline__ void accumulate_aggregates(const uint32_t mask, const int nhor_sm,
                                                      double* const* col, double* bfi,
                                                      const size_t mc, const size_t per_slot,
                                                      const FluxCapture& f, const CellStates<NH>& s,
                                                      const double fS, const double* sat_o) {
  if (mask & 1u) {  // :785-786
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h)
      if (h < nhor_sm) a = a + s.sm[h];
#pragma unroll
    for (int h = 0; h < NH; ++h)
      if (h < nhor_sm) b = b + sat_o[h];
    col[0][mc] = col[0][mc] + a / b;
  }
  if (mask & 2u) {  // :827-830
    const double fNS = 1.0 - fS;
    double a = 0.0;
#pragma unroll
    for (int h = 0; h < NH; ++h) a = a + f.aet_soil[h];
    const double t1 = a * fNS, t2 = f.v[MHM_F_AETSEALED] * fS;
    col[1][mc] = col[1][mc] + ((t1 + f.v[MHM_F_AETCANOPY]) + t2);
  }
  if (mask & 4u) {  // :850-854: average_add of the five stores, then add per horizon
    double acc = col[2][mc];
    acc = acc + ((((s.inter + s.snowpack) + s.sealed) + s.unsat) + s.sat);
#pragma unroll
    for (int h = 0; h < NH; ++h) acc = acc + s.sm[h];
    col[2][mc] = acc;
  }
  if (mask & 8u) {
    bfi[mc] = bfi[mc] + f.v[MHM_F_BASEFLOW];
    bfi[per_slot + mc] = bfi[per_slot + mc] + f.v[MHM_F_TOTAL_RUNOFF];
  }
}

// kernel variants: the generic one takes every process selection at run time; the two
// specialised ones cover the configurations every large run uses (hourly forcing with PET as
// input, i.e. processCase(5) <= 0) with the soil-moisture scheme fixed at compile time.
enum CellVariant { kGeneric = 0, kHourlyFeddes = 1, kHourlyJarvis = 2 };

// returns total_runoff (mo_runoff.f90:271-272)
template <int NH, int VARIANT, bool EMIT, bool OUT, class PARAMS>
__device__ __forceinline__ double cascade_step(const PARAMS& p, CellStates<NH>& s,
                                               const double pet, const double temperature,
                                               const double prec, const int soil_case,
                                               const double evap_coeff,
#if MHM_FAST
                                               const double inv_evap_coeff, double2* warp_tasks,
                                               const fm::Tables& tab,
#endif
                                               const FluxEmitter<EMIT, OUT>& emit) {
  // ---- canopy_interc, mo_canopy_interc.f90:105-131 ----
  double throughfall, aet_canopy;
  {
    double aux = s.inter + prec;
    double ic;
    if (aux >= PX(maxInter)) {
      throughfall = aux - PX(maxInter);
      ic = PX(maxInter);
    } else {
      throughfall = 0.0;
      ic = aux;
    }
    double ev;
    if (PX(maxInter) > kEps) {
#if MHM_FAST
      // x**(2/3) = x * x**(-1/3); x = 0 is the common dry-canopy case
      const double x = ic * PX(inv_maxInter);
      ev = (x == 0.0) ? 0.0 : pet * fm::pow23_pos(x);
#else
      ev = pet * pow(ic / PX(maxInter), kTwoThird);
#endif
    } else {
      ev = 0.0;
    }
    if (ev < 0.0) ev = 0.0;
    if (ic > ev) {
      ic = ic - ev;
    } else {
      ev = ic;
      ic = 0.0;
    }
    s.inter = ic;
    aet_canopy = ev;
    emit(MHM_F_THROUGHFALL, throughfall);
    emit(MHM_F_AETCANOPY, aet_canopy);
  }

  // ---- snow_accum_melt, mo_snow_accum_melt.f90:117-156 ----
  double prec_effect;
  {
    const bool warm = temperature > PX(tthr);
    double snow, rain, melt, dd;
    if (warm) {
      snow = 0.0;
      rain = throughfall;
    } else {
      snow = throughfall;
      rain = 0.0;
    }
    if (prec <= PX(ddthr)) {
      dd = PX(ddnop_c) + PX(ddinc) * prec;
    } else {
      dd = PX(ddmax_c);
    }
    if (warm) {
      if (s.snowpack > 0.0) {
        double aux = dd * (temperature - PX(tthr));
        if (aux > s.snowpack) {
          melt = s.snowpack;
          s.snowpack = 0.0;
        } else {
          melt = aux;
          s.snowpack = s.snowpack - aux;
        }
      } else {
        melt = 0.0;
        s.snowpack = 0.0;
      }
    } else {
      melt = 0.0;
      s.snowpack = s.snowpack + snow;
    }
    prec_effect = melt + rain;
    emit(MHM_F_SNOW, snow);
    emit(MHM_F_RAIN, rain);
    emit(MHM_F_MELT, melt);
    emit(MHM_F_DEGDAY, dd);
    emit(MHM_F_PREEFFECT, prec_effect);
  }

  // ---- soil_moisture, mo_soil_moisture.f90:179-286 ----
  double runoff_sealed = 0.0, infil_last = 0.0;
  {
    double aet_sealed = 0.0;
    if (PX(fSealed) > 0.0) {
      double tmp = s.sealed + prec_effect;
      double st;
      if (tmp > PX(sealedThr)) {
        runoff_sealed = tmp - PX(sealedThr);
        st = PX(sealedThr);
      } else {
        runoff_sealed = 0.0;
        st = tmp;
      }
      if (PX(sealedThr) > kEps) {
#if MHM_FAST
        aet_sealed = (pet * inv_evap_coeff - aet_canopy) * (st * PX(inv_sealedThr));
#else
        aet_sealed = (pet / evap_coeff - aet_canopy) * (st / PX(sealedThr));
#endif
        if (aet_sealed < 0.0) aet_sealed = 0.0;
      } else {
        aet_sealed = DBL_MAX;  // huge(1.0_dp)
      }
      if (st > aet_sealed) {
        st = st - aet_sealed;
      } else {
        aet_sealed = st;
        st = 0.0;
      }
      s.sealed = st;
    }
    emit(MHM_F_RUNOFFSEAL, runoff_sealed);
    emit(MHM_F_AETSEALED, aet_sealed);

    double prec_effec_soil = prec_effect;
    double aet_pos_sum = 0.0;  // sum(aet(1:hh-1), mask = aet > 0), accumulated in index order
#if MHM_FAST
    // frac_runoff_h = (sm_h / sat_h) ** exp_h depends only on the states the step starts from,
    // so all horizons' powers are known up front.  Only cells that receive water need them
    // (a dry step leaves inf = 0 and sm unchanged whatever frac is): the (cell, horizon)
    // pairs of the warp that need one are compacted through shared memory and evaluated by
    // as few warp-wide pow_pos rounds as possible, instead of NH rounds at a fraction of
    // the lanes each.  Same function, same arguments -> same values as the direct evaluation.
    double frac_pre[NH];
    {
      const unsigned lane = threadIdx.x & 31u;
      const unsigned lt = (1u << lane) - 1u;
      const bool wet = prec_effect != 0.0;
      unsigned base = 0;
      unsigned slot[NH];
      bool need[NH];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        need[hh] = wet && s.sm[hh] > kEps && !(s.sm[hh] > PH(SAT, hh));
        const unsigned m = __ballot_sync(0xffffffffu, need[hh]);
        slot[hh] = base + __popc(m & lt);
        base += __popc(m);
        frac_pre[hh] = 0.0;
      }
      if (base != 0) {  // warp-uniform
#pragma unroll
        for (int hh = 0; hh < NH; ++hh)
          if (need[hh]) warp_tasks[slot[hh]] = make_double2(s.sm[hh] * PH(inv_SAT, hh), PH(EXPN, hh));
        __syncwarp();
        for (unsigned k = lane; k < base; k += 32u) {
          const double2 tk = warp_tasks[k];
          warp_tasks[k].x = fm::pow_tab(tab, tk.x, tk.y);
        }
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < NH; ++hh)
          if (need[hh]) frac_pre[hh] = warp_tasks[slot[hh]].x;
        __syncwarp();
      }
    }
#endif
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      double sm = s.sm[hh];
      const double sat = PH(SAT, hh);
      double inf;
      if (hh != 0) prec_effec_soil = infil_last;
      if (sm > sat) {
        inf = prec_effec_soil;
      } else {
#if MHM_FAST
        // prec_effec_soil == 0 (the common dry step) gives tmp = 0, inf = 0, sm unchanged
        // exactly, whatever frac_runoff is: skip the exp/log pair.
        if (prec_effec_soil == 0.0) {
          inf = 0.0;
        } else {
          const double frac_runoff = frac_pre[hh];
          double tmp = prec_effec_soil * (1.0 - frac_runoff);
          if ((sm + tmp) > sat) {
            inf = prec_effec_soil + (sm - sat);
            sm = sat;
          } else {
            inf = prec_effec_soil - tmp;
            sm = sm + tmp;
          }
        }
#else
        double frac_runoff;
        if (sm > kEps) {
          frac_runoff = exp(PH(EXPN, hh) * log(sm / sat));
        } else {
          frac_runoff = 0.0;
        }
        double tmp = prec_effec_soil * (1.0 - frac_runoff);
        if ((sm + tmp) > sat) {
          inf = prec_effec_soil + (sm - sat);
          sm = sat;
        } else {
          inf = prec_effec_soil - tmp;
          sm = sm + tmp;
        }
#endif
      }
      infil_last = inf;
      emit(MHM_F_INFILSOIL, hh, inf);

      double a = pet - aet_canopy;
      if (hh != 0) a = a - aet_pos_sum;
      double stress;
      const bool feddes = VARIANT == kHourlyFeddes ? true
                          : VARIANT == kHourlyJarvis ? false
                                                     : (soil_case == 1 || soil_case == 4);
      if (feddes) {  // feddes_et_reduction :353-361
        if (sm >= PH(FC, hh)) {
          stress = PH(fRoots, hh);
        } else if (sm > PH(WP, hh)) {
#if MHM_FAST
          stress = PH(fRoots, hh) * (sm - PH(WP, hh)) * PH(inv_FCWP, hh);
#else
          stress = PH(fRoots, hh) * (sm - PH(WP, hh)) / (PH(FC, hh) - PH(WP, hh));
#endif
        } else {
          stress = 0.0;
        }
      } else {  // jarvis_et_reduction :431-444 (cases 2, 3)
        double th = (sm - PH(WP, hh)) / (sat - PH(WP, hh));
        if (th < 0.0) th = 0.0;
        if (th > 1.0) th = 1.0;
        stress = 0.0;
        if (th >= PX(jarvis_c1)) {
          stress = PH(fRoots, hh);
        } else if (th < PX(jarvis_c1)) {
          stress = PH(fRoots, hh) * (th / PX(jarvis_c1));
        }
      }
      a = a * stress;
      if (a < 0.0) a = 0.0;
      if (sm > a) {
        sm = sm - a;
      } else {
        a = sm - kEps;
        sm = kEps;
      }
      if (sm < kEps) sm = kEps;
      emit(MHM_F_AETSOIL, hh, a);
      s.sm[hh] = sm;
      // running masked sum in index order (bit-identical to Fortran's sum(..., mask))
      if (a > 0.0) aet_pos_sum = aet_pos_sum + a;
    }
  }

  // ---- runoff_unsat_zone, mo_runoff.f90:120-152 ----
  double fast = 0.0, slow = 0.0;
  {
    double us = s.unsat + infil_last;
    if (us > PX(unsatThr)) fast = fmin(PX(k0r) * (us - PX(unsatThr)), us - kEps);
    us = us - fast;
    if (us > kEps) {
#if MHM_FAST
      slow = fmin(PX(k1r) * fm::pow_tab(tab, us, 1.0 + PX(alpha)), us - kEps);
#else
      slow = fmin(PX(k1r) * pow(us, 1.0 + PX(alpha)), us - kEps);
#endif
    }
    us = us - slow;
    double perc = PX(kpr) * us;
    if (us > perc) {
      us = us - perc;
      s.sat = s.sat + perc * PX(karst);
    } else {
      s.sat = s.sat + us * PX(karst);
      us = 0.0;
    }
    s.unsat = us;
    emit(MHM_F_FASTRUNOFF, fast);
    emit(MHM_F_SLOWRUNOFF, slow);
    emit(MHM_F_PERCOL, perc);
  }
  // ---- runoff_sat_zone :204-210 ----
  double baseflow;
  if (s.sat > 0.0) {
    baseflow = PX(k2r) * s.sat;
    s.sat = s.sat - baseflow;
  } else {
    baseflow = 0.0;
    s.sat = 0.0;
  }
  emit(MHM_F_BASEFLOW, baseflow);
  // ---- L1_total_runoff :271-272 ----
  const double total_runoff =
      ((baseflow + slow + fast) * (1.0 - PX(fSealed))) + (runoff_sealed * PX(fSealed));
  emit(MHM_F_TOTAL_RUNOFF, total_runoff);
  return total_runoff;
}

#if MHM_FAST
// (a < b) ? x : y and friends as setp + selp: written in PTX because the compiler recognises
// `a < b ? a : b` as fmin / fmax, whose sm_100 expansion (DSETP.MIN + NaN fix-up) costs six
// instructions instead of three.  NaN compares false, like in the C expression.
__device__ __forceinline__ double sel_lt(double a, double b, double x, double y) {
  double r;
  asm("{\n\t.reg .pred q;\n\tsetp.lt.f64 q, %1, %2;\n\tselp.f64 %0, %3, %4, q;\n\t}" : "=d"(r) : "d"(a), "d"(b), "d"(x), "d"(y));
  return r;
}
__device__ __forceinline__ double sel_gt(double a, double b, double x, double y) {
  double r;
  asm("{\n\t.reg .pred q;\n\tsetp.gt.f64 q, %1, %2;\n\tselp.f64 %0, %3, %4, q;\n\t}" : "=d"(r) : "d"(a), "d"(b), "d"(x), "d"(y));
  return r;
}
__device__ __forceinline__ double min_sel(double a, double b) { return sel_lt(a, b, a, b); }  // NaN a -> b
__device__ __forceinline__ double pos_part(double a) { return sel_gt(a, 0.0, a, 0.0); }       // NaN -> 0
#endif

#if MHM_FAST
// ---- select form of the cascade for the specialised fast variants ------------------------------
// The same water balance as cascade_step, written as straight-line code for the instruction issue
// port: the fused kernel issues ~3 instructions per fp64 instruction, so every select pair, every
// address computation and every parameter load counts.  Three means:
//  * exact floating-point identities instead of select pairs.  x - x = +0, x +- 0 = x and
//    a - (c - b) = a + (b - c) hold bit for bit, so "the rest" of a min() is one subtraction
//    (throughfall = aux - min(aux, maxInter); rain = throughfall - snow; unsat = us - min(us, perc) ...)
//    and the masked sums of the reference become plain sums of terms that are exactly zero;
//  * conditions that can only be false for a vanishing value are dropped: sat_storage > 0
//    (k2 * 0 = 0, 0 - 0 = 0), snow_pack > 0 (min(pot, 0) = 0), the positive-part mask of the aET
//    sum (aET >= 0).  Precondition of the fast mode: states and fluxes of the model are not
//    negative (the model keeps them so itself; strict mode is literal and has no precondition);
//  * parameters live in shared memory as PAIRS that are used together: one 128-bit load each.
// Values: identical to the branch form except where a reciprocal is hoisted or FMA-contracted
// differently (<= a few ulp per step; tests hold fast mode to 1e-9 against the oracle).
//
// Feddes and Jarvis share the code: both reduce the root water uptake to
//   aET_h = max(0, (pet_h * fRoots_h) * min(1, (sm_h - WP_h) * inv_range_h))
// with inv_range = 1 / (FC - WP) (Feddes, mo_soil_moisture.f90:353-361: sm >= FC <=> factor >= 1)
// or 1 / ((SAT - WP) * jarvis_c1) (Jarvis, :431-444: theta >= c1 <=> factor >= 1).
// Pair ids, ordered by how much shared-memory traffic a register copy saves: the first
// PairStore::kRegPairs ids live in registers, the rest in shared memory (the kernel is bound by the
// shared-memory data pipe: ~110 wavefronts per warp-step when every pair is re-read from there).
// Horizon h's pairs: pid_sat(h) first (read twice per step), the others at the end.
template <int NH>
struct PairIds {
  static constexpr int kSat0 = 0;             // SAT, 1 / SAT                       x NH
  static constexpr int kSeal2 = NH;           // fSealed, 1 - fSealed               (stage A and runoff)
  static constexpr int kDdBase = NH + 1;      // ddmax * c2TSTu, c2TSTu / kBaseFlow  (snow; baseflow)
  static constexpr int kPetTthr = NH + 2;     // petFac [iLAI, yId], tempThresh [yId]
  static constexpr int kMaxInter = NH + 3;    // maxInter, 1 / maxInter (0 when maxInter <= eps: no canopy evaporation)
  static constexpr int kDd = NH + 4;          // ddnoprec * c2TSTu, ddinc
  static constexpr int kSeal = NH + 5;        // sealedThr, 1 / sealedThr (+inf when sealedThr <= eps: everything evaporates)
  static constexpr int kUnsat = NH + 6;       // unsatThr, c2TSTu / kFastFlow
  static constexpr int kSlow = NH + 7;        // c2TSTu / kSlowFlow, 1 + alpha
  static constexpr int kPerc = NH + 8;        // c2TSTu / kPerco, karstLoss
  static constexpr int kRoot0 = NH + 9;       // fRoots, inv_range (see above)      x NH
  // exponents and wilting points, TWO horizons per pair (component h % 2 of pair h / 2): the horizon
  // loop reads one pair for both horizons' wilting points, the power tasks only ever touch the exponents
  static constexpr int kNHalf = (NH + 1) / 2;
  static constexpr int kWp0 = 2 * NH + 9;
  static constexpr int kExp0 = 2 * NH + 9 + kNHalf;
  static constexpr int kCount = 2 * NH + 9 + 2 * kNHalf;
};
#ifndef MHM_PARAM_REG_PAIRS
#define MHM_PARAM_REG_PAIRS 4  // pairs kept in registers when the rest is in shared memory (see ParamPlace)
#endif
// SHARED_STORE: pairs >= kRegPairs in shared memory, one conflict-free 16-byte column per thread;
// otherwise everything in registers (more than two horizons)
template <int NH, bool SHARED_STORE>
struct PairStore {
  static constexpr bool kShared = SHARED_STORE;
  static constexpr int kRegPairs = SHARED_STORE ? (MHM_PARAM_REG_PAIRS < PairIds<NH>::kCount ? MHM_PARAM_REG_PAIRS
                                                                                              : PairIds<NH>::kCount)
                                                : PairIds<NH>::kCount;
  static constexpr int kSmemPairs = PairIds<NH>::kCount - kRegPairs;
  struct Smem {
    double2 q[kSmemPairs > 0 ? kSmemPairs : 1][kCellThreads];
  };
  double2 r[kRegPairs > 0 ? kRegPairs : 1];
  Smem* sm;
  __device__ __forceinline__ double2 get(int k) const { return k < kRegPairs ? r[k] : sm->q[k - kRegPairs][threadIdx.x]; }
  // a pair of another thread's column (shared-memory pairs only)
  __device__ __forceinline__ double2 get_of(int k, unsigned tid) const { return sm->q[k - kRegPairs][tid]; }
  // one component (0: x, 1: y) of another thread's pair: a single 8-byte load
  __device__ __forceinline__ double get_half_of(int k, unsigned tid, unsigned half) const {
    return reinterpret_cast<const double*>(&sm->q[k - kRegPairs][tid])[half];
  }
  __device__ __forceinline__ void set(int k, double x, double y) {
    if (k < kRegPairs) r[k] = make_double2(x, y);
    else sm->q[k - kRegPairs][threadIdx.x] = make_double2(x, y);
  }
  __device__ __forceinline__ void setx(int k, double x) {
    if (k < kRegPairs) r[k].x = x;
    else sm->q[k - kRegPairs][threadIdx.x].x = x;
  }
  __device__ __forceinline__ void sety(int k, double y) {
    if (k < kRegPairs) r[k].y = y;
    else sm->q[k - kRegPairs][threadIdx.x].y = y;
  }
};
#define PID PairIds<NH>
#define P2(k_) p.get(PID::k_)
#define PH2(k_, h_) p.get(PID::k_ + (h_))

// What stage A (canopy, snow, sealed store: everything that needs the step's forcing) hands to
// stage B (soil horizons, unsaturated and saturated zone).  Stage A touches only the states
// inter / snowpack / sealed, stage B only soil moisture / unsat / sat, so stage A of step t+1
// may run beside stage B of step t (see the uniform-calendar time loop).
struct StageA {
  double prec_effect, pet_left, runoff_sealed;  // pet_left = pet - aet_canopy
  double aet_canopy = 0.0, aet_sealed = 0.0;    // carried to the step's output sums (OUT == 2 pipeline)
};
template <int NH, bool EMIT, class PARAMS, class EM>
__device__ __forceinline__ StageA cascade_stage_a_sel(const PARAMS& p, CellStates<NH>& s, const double pet,
                                                      const double temperature, const double prec,
                                                      const double tthr, const double inv_evap_coeff,
                                                      const EM& emit) {
  // ---- canopy_interc, mo_canopy_interc.f90:105-131 ----
  const double2 mi = P2(kMaxInter);
  const double aux = s.inter + prec;
  const double ic0 = min_sel(aux, mi.x);
  const double throughfall = aux - ic0;          // aux - maxInter, or exactly 0
  const double evr = pet * fm::pow23_nz(ic0 * mi.y);  // NaN for an empty canopy
  const double ev = pos_part(evr);               // NaN, negative -> 0 (:118-121)
  const double aet_canopy = min_sel(ev, ic0);
  s.inter = ic0 - aet_canopy;                    // ic0 - ev, or exactly 0
  emit(MHM_F_THROUGHFALL, throughfall);
  emit(MHM_F_AETCANOPY, aet_canopy);

  // ---- snow_accum_melt, mo_snow_accum_melt.f90:117-156 ----
  const double2 d2 = P2(kDd);
  const bool warm = temperature > tthr;
  const double rain = warm ? throughfall : 0.0;
  const double snow = throughfall - rain;        // exactly throughfall or 0
  // min(ddnoprec + ddinc * prec, ddmax): the reference's test prec <= (ddmax - ddnoprec) / ddinc
  // (:127) up to the rounding of that quotient
  const double dd = min_sel(d2.x + d2.y * prec, P2(kDdBase).x);
  const double pot = dd * (temperature - tthr);
  const double melt_w = sel_gt(pot, s.snowpack, s.snowpack, pot);  // snow_pack = 0 gives melt = 0
  const double melt = warm ? melt_w : 0.0;
  s.snowpack = s.snowpack + (snow - melt);       // one of the two terms is exactly 0
  const double prec_effect = melt + rain;
  emit(MHM_F_SNOW, snow);
  emit(MHM_F_RAIN, rain);
  emit(MHM_F_MELT, melt);
  emit(MHM_F_DEGDAY, dd);
  emit(MHM_F_PREEFFECT, prec_effect);

  // ---- sealed store, mo_soil_moisture.f90:179-215 ----
  const double2 se = P2(kSeal), fs = P2(kSeal2);
  const bool sealed_on = fs.x > 0.0;
  const double tmp_s = s.sealed + prec_effect;
  const double st0 = sel_gt(tmp_s, se.x, se.x, tmp_s);
  const double rs = tmp_s - st0;                 // tmp - sealedThr, or exactly 0
  const double aer = (pet * inv_evap_coeff - aet_canopy) * (st0 * se.y);
  const double ae = pos_part(aer);
  const double aes = min_sel(ae, st0);
  if (sealed_on) s.sealed = st0 - aes;           // st0 - ae, or exactly 0
  if (EM::kEmit || EM::kOut) {
    emit(MHM_F_RUNOFFSEAL, sealed_on ? rs : 0.0);
    emit(MHM_F_AETSEALED, sealed_on ? aes : 0.0);
  }
  StageA out;
  out.prec_effect = prec_effect;
  out.pet_left = pet - aet_canopy;
  out.runoff_sealed = rs;  // multiplied by fSealed = 0 where the reference has no sealed runoff
  return out;
}

// The warp's list of pending powers.  With the parameters in shared memory an entry is the 8-byte
// base plus one byte naming its source (lane | horizon << 5): the evaluating lane reads the exponent
// from the source lane's parameter column.  576 bytes per warp instead of 1 KB -- together with the
// 15 parameter pairs that is what lets a sixth CTA fit on the SM.  Otherwise (base, exponent) pairs.
template <int NH, bool BY_SOURCE>
struct WarpTasks;
template <int NH>
struct WarpTasks<NH, true> {
  static constexpr bool kBySource = true;
  double x[32 * NH];
  unsigned char src[32 * NH];
  template <class PARAMS>
  __device__ __forceinline__ void put(const PARAMS&, unsigned k, double base, int hh) {
    x[k] = base;
    src[k] = (unsigned char)((threadIdx.x & 31u) | ((unsigned)hh << 5));
  }
  template <class PARAMS>
  __device__ __forceinline__ void eval(const PARAMS& p, unsigned k, const fm::Tables& tab) {
    const unsigned sc = src[k];
    const unsigned hh = (sc >> 5) < (unsigned)NH ? (sc >> 5) : 0u;  // (a stale entry must stay in range)
    const double y = p.get_half_of(PID::kExp0 + (int)(hh >> 1), (threadIdx.x & ~31u) + (sc & 31u), hh & 1u);
    x[k] = fm::pow_tab(tab, x[k], y);
  }
  __device__ __forceinline__ double result(unsigned k) const { return x[k]; }
};
template <int NH>
struct WarpTasks<NH, false> {
  static constexpr bool kBySource = false;
  double2 xy[32 * NH];
  template <class PARAMS>
  __device__ __forceinline__ void put(const PARAMS& p, unsigned k, double base, int hh) {
    const double2 e2 = p.get(PID::kExp0 + hh / 2);
    xy[k] = make_double2(base, (hh & 1) ? e2.y : e2.x);
  }
  template <class PARAMS>
  __device__ __forceinline__ void eval(const PARAMS&, unsigned k, const fm::Tables& tab) {
    const double2 tk = xy[k];
    xy[k].x = fm::pow_tab(tab, tk.x, tk.y);
  }
  __device__ __forceinline__ double result(unsigned k) const { return xy[k].x; }
};

// infiltration powers of the warp, compacted (see cascade_step): frac_h = (sm_h / SAT_h) ** EXPN_h
// for the (cell, horizon) pairs that receive water; 0 elsewhere.  (The reference's sm > eps test
// is dropped: at sm = eps the power is below 1e-16 and 1 - frac rounds to the same 1.)
template <int NH, class PARAMS, class TASKS>
__device__ __forceinline__ void cascade_stage_b1_sel(const PARAMS& p, const CellStates<NH>& s,
                                                     const double prec_effect, TASKS& wt,
                                                     const fm::Tables& tab, double (&frac_pre)[NH]) {
  const bool wet = prec_effect != 0.0;
#if MHM_POW_COMPACT
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = (1u << lane) - 1u;
  unsigned base = 0;
  unsigned slot[NH];
  bool need[NH];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    need[hh] = wet && !(s.sm[hh] > PH2(kSat0, hh).x);
    const unsigned m = __ballot_sync(0xffffffffu, need[hh]);
    slot[hh] = base + __popc(m & lt);
    base += __popc(m);
    frac_pre[hh] = 0.0;
  }
  if (base != 0) {  // warp-uniform
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
      if (need[hh]) wt.put(p, slot[hh], s.sm[hh] * PH2(kSat0, hh).y, hh);
    __syncwarp();
    for (unsigned k = lane; k < base; k += 32u) wt.eval(p, k, tab);
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
      if (need[hh]) frac_pre[hh] = wt.result(slot[hh]);
    __syncwarp();
  }
#else
  // every lane evaluates its own powers (NH independent chains in straight-line code, no shared
  // memory round trip); a warp without a wet lane skips them
  (void)wt;
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) frac_pre[hh] = 0.0;
  if (__any_sync(0xffffffffu, wet)) {
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      const double2 sa = PH2(kSat0, hh);
      const double2 e2 = p.get(PID::kExp0 + hh / 2);
      const double f = fm::pow_tab(tab, s.sm[hh] * sa.y, (hh & 1) ? e2.y : e2.x);
      frac_pre[hh] = (wet && !(s.sm[hh] > sa.x)) ? f : 0.0;
    }
  }
#endif
}

// The same compaction in three parts, so that independent work of the thread can be placed between
// the warp barriers (the exchange through shared memory is a chain of four memory latencies):
//   pow_post    : which (cell, horizon) pairs need a power; their entries into the warp's task
//                 list; barrier
//   pow_round0  : this lane evaluates task `lane` -- unconditionally (a lane without a task works
//                 on a stale entry; pow_tab is safe for any bit pattern), so that the code is
//                 straight-line and shares its basic block with whatever follows
//   pow_collect : the rare further rounds (more than 32 tasks), barrier, results back, barrier
template <int NH>
struct PowTasks {
  unsigned slot[NH], base;
  bool need[NH];
};
template <int NH, class PARAMS, class TASKS>
__device__ __forceinline__ void pow_post(const PARAMS& p, const CellStates<NH>& s, const double prec_effect,
                                         TASKS& wt, PowTasks<NH>& pt) {
  const unsigned lt = (1u << (threadIdx.x & 31u)) - 1u;
  const bool wet = prec_effect != 0.0;
  pt.base = 0;
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    const double2 sa = PH2(kSat0, hh);
    pt.need[hh] = wet && !(s.sm[hh] > sa.x);
    const unsigned m = __ballot_sync(0xffffffffu, pt.need[hh]);
    pt.slot[hh] = pt.base + __popc(m & lt);
    pt.base += __popc(m);
    if (pt.need[hh]) wt.put(p, pt.slot[hh], s.sm[hh] * sa.y, hh);
  }
  __syncwarp();
}
template <class PARAMS, class TASKS>
__device__ __forceinline__ void pow_round0(const PARAMS& p, TASKS& wt, const fm::Tables& tab) {
  wt.eval(p, threadIdx.x & 31u, tab);
}
template <int NH, class PARAMS, class TASKS>
__device__ __forceinline__ void pow_collect(const PARAMS& p, const PowTasks<NH>& pt, TASKS& wt,
                                            const fm::Tables& tab, double (&frac_pre)[NH]) {
  if (NH > 1 && pt.base > 32u) {  // warp-uniform
    for (unsigned k = (threadIdx.x & 31u) + 32u; k < pt.base; k += 32u) wt.eval(p, k, tab);
  }
  __syncwarp();
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    const double f = wt.result(pt.slot[hh]);
    frac_pre[hh] = pt.need[hh] ? f : 0.0;
  }
  __syncwarp();
}

// soil horizons, mo_soil_moisture.f90:217-286; returns the infiltration leaving the last horizon
template <int NH, class PARAMS, class EM>
__device__ __forceinline__ double cascade_horizons_sel(const PARAMS& p, CellStates<NH>& s, const StageA& in,
                                                       const double (&frac_pre)[NH], const EM& emit) {
  double pe = in.prec_effect, aet_sum = 0.0;
  double2 wp2 = make_double2(0.0, 0.0);
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    const double2 sa = PH2(kSat0, hh), rt = PH2(kRoot0, hh);
    if ((hh & 1) == 0) wp2 = p.get(PID::kWp0 + hh / 2);
    const double wp = (hh & 1) ? wp2.y : wp2.x;
    const double sm0 = s.sm[hh];
    const double tmp = pe * (1.0 - frac_pre[hh]);
    const double u = sm0 + tmp;
    const double cap = sel_gt(sm0, sa.x, sm0, sa.x);  // an over-saturated horizon keeps its water (:226)
    const bool fill = u > sa.x;                  // also true whenever sm0 > SAT (frac = 0, pe >= 0)
    const double sm1 = fill ? cap : u;
    const double d = fill ? cap - sm0 : tmp;     // water the horizon takes up
    const double inf = pe - d;                   // pe + (sm0 - SAT) bit for bit, pe - tmp, or pe
    emit(MHM_F_INFILSOIL, hh, inf);
    const double A = hh == 0 ? in.pet_left : in.pet_left - aet_sum;
    const double w = (sm1 - wp) * rt.y;
    const double a = pos_part((A * rt.x) * sel_lt(w, 1.0, w, 1.0));  // <= 0 below the wilting point, :266 / :361
    const double a2 = sm1 > a ? a : sm1 - kEps;
    const double sm2 = sm1 - a2;                 // sm1 - a, or eps (0 for sm1 >> eps: floored next)
    emit(MHM_F_AETSOIL, hh, a2);
    s.sm[hh] = sel_lt(sm2, kEps, kEps, sm2);
    aet_sum = hh == 0 ? a2 : aet_sum + a2;       // sum(aet(1:hh), aet > 0): aet >= 0 here
    pe = inf;
  }
  return pe;
}

// unsaturated and saturated zone, total runoff: mo_runoff.f90:120-152, 204-210, 271-272
template <int NH, class PARAMS, class EM>
__device__ __forceinline__ double cascade_reservoirs_sel(const PARAMS& p, CellStates<NH>& s, const double infil_last,
                                                         const double runoff_sealed, const fm::Tables& tab,
                                                         const EM& emit) {
  const double2 un = P2(kUnsat), sl = P2(kSlow), pk = P2(kPerc), fs = P2(kSeal2);
  double us = s.unsat + infil_last;
  const double f1 = un.y * (us - un.x), ue = us - kEps;
  const double fmin1 = min_sel(f1, ue);
  const double fast = us > un.x ? fmin1 : 0.0;
  us = us - fast;
#if MHM_RESV_POW_POLY
  const double s1 = sl.x * fm::pow_pos(us, sl.y), ue2 = us - kEps;       // garbage for us <= eps: discarded
#else
  const double s1 = sl.x * fm::pow_tab(tab, us, sl.y), ue2 = us - kEps;  // garbage for us <= eps: discarded
#endif
  const double smin1 = min_sel(s1, ue2);
  const double slow = us > kEps ? smin1 : 0.0;
  us = us - slow;
  const double perc = pk.x * us;
  const double q = sel_gt(us, perc, perc, us);   // what leaves the unsaturated zone downwards
  s.sat = s.sat + q * pk.y;
  s.unsat = us - q;                              // us - perc, or exactly 0
  emit(MHM_F_FASTRUNOFF, fast);
  emit(MHM_F_SLOWRUNOFF, slow);
  emit(MHM_F_PERCOL, perc);
  const double baseflow = P2(kDdBase).y * s.sat;  // sat_storage = 0 gives 0 and leaves 0
  s.sat = __dadd_rn(s.sat, -baseflow);
  emit(MHM_F_BASEFLOW, baseflow);
  const double total_runoff = ((baseflow + slow + fast) * fs.y) + (runoff_sealed * fs.x);
  emit(MHM_F_TOTAL_RUNOFF, total_runoff);
  return total_runoff;
}

template <int NH, bool EMIT, class PARAMS, class EM>
__device__ __forceinline__ double cascade_stage_b2_sel(const PARAMS& p, CellStates<NH>& s, const StageA& in,
                                                       const double (&frac_pre)[NH], const fm::Tables& tab,
                                                       const EM& emit) {
  const double infil_last = cascade_horizons_sel<NH>(p, s, in, frac_pre, emit);
  return cascade_reservoirs_sel<NH>(p, s, infil_last, in.runoff_sealed, tab, emit);
}

template <int NH, bool EMIT, class PARAMS, class TASKS, class EM>
__device__ __forceinline__ double cascade_step_sel(const PARAMS& p, CellStates<NH>& s, const double pet,
                                                   const double temperature, const double prec,
                                                   const double inv_evap_coeff, TASKS& warp_tasks,
                                                   const fm::Tables& tab, const EM& emit) {
  const StageA sa = cascade_stage_a_sel<NH, EMIT>(p, s, pet, temperature, prec, P2(kPetTthr).y, inv_evap_coeff, emit);
  double frac_pre[NH];
  cascade_stage_b1_sel<NH>(p, s, sa.prec_effect, warp_tasks, tab, frac_pre);
  return cascade_stage_b2_sel<NH, EMIT>(p, s, sa, frac_pre, tab, emit);
}
#endif

#if MHM_FAST
// ---- forcing through the TMA unit (cp.async.bulk + mbarrier) ------------------------------------
// A warp's 32 cells are one contiguous 256-byte stretch of every forcing row.  Lane 0 of the warp
// keeps kFStages rows in flight: three 1-D bulk copies per row (precipitation, temperature, PET) land
// in the warp's slice of a shared-memory ring and complete the row's mbarrier; the lanes wait for the
// barrier's phase, read their three values, and the slot is refilled for the row kFStages later.  A
// warp waits only for itself (the barriers are per warp), no registers hold prefetched forcing and no
// per-lane global address is computed.  Needs 16-byte aligned rows (an even number of cells).
constexpr int kFStages = 4;
struct alignas(128) ForcingRing {
  double v[kFStages][3][kCellThreads];
  unsigned long long full[kCellThreads / 32][kFStages];
};
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
#endif

__device__ __forceinline__ double ldg_stream(const double* p) { return __ldg(p); }
// everything of the time loop that lives across steps besides states and parameters
struct CellCursor {
  int cur_y, cur_l;
  long long cur_row;
  double raw_pre, raw_temp, raw_pet;
  const double *ppre, *ptemp, *ppet;  // this cell in forcing row cur_row
  double* hist;                       // this cell/member in the total-runoff row of step t
  double* qp;                         // this cell's node in the tiled node-runoff history, step qst
  int qst;
};

// which parameter store a kernel variant uses
template <int NH, int VARIANT>
struct ParamStoreOf {
#if MHM_FAST
  static constexpr bool paired = VARIANT != kGeneric;  // specialised fast variants: select form
#else
  static constexpr bool paired = false;
#endif
};

// FUSED (uniform launches only): the step's runoff goes to the routing's tiled node-runoff history
// and nowhere else (true) / to the total-runoff history row and nowhere else (false) -- a block of
// steps always has exactly one of the two sinks
#ifndef MHM_OUT2_MIN_BLOCKS
#define MHM_OUT2_MIN_BLOCKS 3  // CTAs/SM of the launches that keep the default output window in registers
#endif
// OUT: 0 no gridded outputs / aggregates, 1 run-time selection with the window in shared memory,
// 2 the reference's default output set with the window in registers (accumulate_outputs_default)
// TMA (uniform launches): the forcing rows arrive through cp.async.bulk into a shared-memory ring
template <int NH, int VARIANT, int OUT, bool UNIFORM = false, bool FUSED = false, bool TMA = false>
__global__ void __launch_bounds__(kCellThreads, OUT == 2 ? MHM_OUT2_MIN_BLOCKS : ParamPlace<NH>::min_blocks)
MHM_KERNEL_NAME(const __grid_constant__ CellArgs a) {
  const int member = blockIdx.x % a.nMembers;
  const int cell = (blockIdx.x / a.nMembers) * kCellThreads + threadIdx.x;
#if MHM_FAST
  // the warps' lists of pending infiltration powers
  // (by source: the exponents must be in shared memory)
  constexpr bool kTasksBySource = ParamStoreOf<NH, VARIANT>::paired && ParamPlace<NH>::shared &&
                                  PairStore<NH, true>::kRegPairs <= PairIds<NH>::kExp0;
  __shared__ WarpTasks<NH, kTasksBySource> sh_tasks[kCellThreads / 32];
#if MHM_TABLES_GLOBAL
  const fm::Tables& sh_tab = fm::d_tables;  // served from L1
#else
  __shared__ fm::Tables sh_tab;  // log / exp tables of fastmath.cuh, 4 KB
  {
    static_assert(sizeof(fm::Tables) == 256 * sizeof(double2), "table layout");
    const double2* src = reinterpret_cast<const double2*>(&fm::d_tables);
    double2* dst = reinterpret_cast<double2*>(&sh_tab);
    for (int i = threadIdx.x; i < 256; i += kCellThreads) dst[i] = src[i];
    __syncthreads();
  }
#endif
  auto& warp_tasks = sh_tasks[threadIdx.x >> 5];
  // out-of-range lanes of the last tile stay alive (warp collectives) on a valid cell
  const bool live = cell < a.nCells;
  const int c = live ? cell : a.nCells - 1;
#else
  if (cell >= a.nCells) return;
  const bool live = true;
  const int c = cell;
#endif
  const size_t n = (size_t)a.nCells;
  const size_t mc = (size_t)member * n + c;
  constexpr bool kHourlyPetIn = VARIANT != kGeneric;
  constexpr bool kPaired = ParamStoreOf<NH, VARIANT>::paired;

  CellStates<NH> s;
  s.inter = a.S[MHM_S_INTER][mc];
  s.snowpack = a.S[MHM_S_SNOWPACK][mc];
  s.sealed = a.S[MHM_S_SEALSTW][mc];
  s.unsat = a.S[MHM_S_UNSATSTW][mc];
  s.sat = a.S[MHM_S_SATSTW][mc];
#pragma unroll
  for (int h = 0; h < NH; ++h) s.sm[h] = a.S[MHM_S_SOILMOIST][((size_t)member * NH + h) * n + c];

  constexpr bool kSharedParams = ParamPlace<NH>::shared;
#if MHM_FAST
  // specialised variants: parameter pairs, partly in registers, the rest in shared memory
  using Pairs = PairStore<NH, kSharedParams>;
  __shared__ std::conditional_t<kPaired && kSharedParams, typename Pairs::Smem, EmptyParams> pair_smem;
  std::conditional_t<kPaired, Pairs, EmptyParams> pairs;
  if constexpr (kPaired && kSharedParams) pairs.sm = &pair_smem;
#else
  EmptyParams pairs;
#endif
  // generic variant: one member per parameter, in shared memory or in registers
  __shared__ std::conditional_t<!kPaired && kSharedParams, CellParamsShared<NH>, EmptyParams> p_shared;
  std::conditional_t<!kPaired && !kSharedParams, CellParams<NH>, EmptyParams> p_regs;
  auto& p = pick_ref<kPaired>(pairs, pick_ref<kSharedParams>(p_shared, p_regs));
  // parameters that never change during a run
  if constexpr (!kPaired) {
    PX(karst) = a.P[MHM_P_KARSTLOSS][mc];
    PX(jarvis_c1) = a.P[MHM_P_JARVIS_C1][mc];
    PX(unsatThr) = a.P[MHM_P_UNSATTHRESH][mc];
    PX(sealedThr) = a.P[MHM_P_SEALEDTHRESH][mc];
#if MHM_FAST
    PX(inv_sealedThr) = 1.0 / PX(sealedThr);
#endif
  }
#if MHM_FAST
  else {
    const double thr = a.P[MHM_P_SEALEDTHRESH][mc];
    p.set(PID::kSeal, thr, thr > kEps ? 1.0 / thr : __longlong_as_double(0x7ff0000000000000LL));
  }
#endif
  CellCursor cu;
  cu.cur_y = -1;
  cu.cur_l = -1;
  cu.cur_row = a.idx_in[0].iMeteoTS;
  cu.ppre = a.met[MHM_M_PRE] + (size_t)(cu.cur_row - a.met_first[MHM_M_PRE]) * n + c;
  cu.ptemp = a.met[MHM_M_TEMP] + (size_t)(cu.cur_row - a.met_first[MHM_M_TEMP]) * n + c;
  cu.ppet = a.pet_case <= 0 ? a.met[MHM_M_PET] + (size_t)(cu.cur_row - a.met_first[MHM_M_PET]) * n + c
                            : nullptr;
  cu.hist = a.runoff_hist ? a.runoff_hist + (size_t)member * n + c : nullptr;
  const size_t hist_stride = (size_t)a.nMembers * n;
  // fused L11_runoff_acc (mo_mrm_pre_routing.f90:110-141): this cell is the only one of its node
  double* qout = nullptr;
  double qarea = 0.0;
  if (a.qout_hist) {
    qout = a.qout_hist + (((size_t)member * a.qout_E + a.cell_lane[c]) << 3);
    qarea = a.cell_area[c];
  }
#if MHM_FAST
  const double qscale = qarea * a.qout_scale;  // node runoff = total runoff * area * 1000 / TST
#endif
  const size_t qtile_stride = ((size_t)a.nMembers * a.qout_E) << 3;
  // the history is stored skewed: step e of a node sits in slot e + (position of the node in its
  // routing segment), the slot the routing pipeline touches in the same sub-step for all lanes
  cu.qst = a.qout_step0 + (qout ? (int)a.cell_skew[c] : 0);
  cu.qp = qout ? qout + (size_t)(cu.qst >> 3) * qtile_stride + (size_t)(cu.qst & 7) : nullptr;

  // gridded outputs: the open window of this (cell, member), see accumulate_outputs
  extern __shared__ double out_sh[];  // [out_nslots][kCellThreads], OUT launches only
  double* const out_acc_l = out_sh + threadIdx.x;
  int out_y = -1;
  double out_fS = 0.0, out_sat[NH];
  if (OUT == 1) {
    if (a.out_mask && live)
      for (int k = 0; k < a.out_nslots; ++k) out_acc_l[k * kCellThreads] = a.out_acc[(size_t)k * hist_stride + mc];
  }
  double oacc[OUT == 2 ? 16 + 3 * NH : 1];
  if (OUT == 2) {
#pragma unroll
    for (int k = 0; k < 16 + 3 * NH; ++k) oacc[OUT == 2 ? k : 0] = a.out_acc[(size_t)k * hist_stride + mc];
  }
  // parameters of the step's land-cover scene / LAI step, reloaded when they change (with a
  // uniform calendar: at the launch's first step only)
  auto load_params = [&](const int t) {
    const StepIdx si = a.idx_in[UNIFORM ? 0 : t];
    const int y = si.yId - 1, il = si.iLAI - 1;
    if ((!UNIFORM || t == 0) && y != cu.cur_y) {  // land-cover scene changed (new year): mo_mhm_interface_run.f90:626-628
      cu.cur_y = y;
      const size_t o1 = ((size_t)member * a.nLC + y) * n + c;  // (n, 1, nLC) arrays
      const double ddinc = a.P[MHM_P_DEGDAYINC][o1];
      const double ddmax_c = a.P[MHM_P_DEGDAYMAX][o1] * a.c2TSTu;    // mo_mhm.f90:463
      const double ddnop_c = a.P[MHM_P_DEGDAYNOPRE][o1] * a.c2TSTu;  // mo_mhm.f90:464
      const double ddthr = (ddmax_c - ddnop_c) / ddinc;              // mo_snow_accum_melt.f90:127
      const double k0r = a.c2TSTu / a.P[MHM_P_KFASTFLOW][o1];        // mo_mhm.f90:484
      const double k1r = a.c2TSTu / a.P[MHM_P_KSLOWFLOW][o1];
      const double kpr = a.c2TSTu / a.P[MHM_P_KPERCO][o1];
      const double k2r = a.c2TSTu / a.P[MHM_P_KBASEFLOW][o1];        // mo_mhm.f90:488
      if constexpr (!kPaired) {
        PX(fSealed) = a.P[MHM_P_FSEALED][o1];
        PX(alpha) = a.P[MHM_P_ALPHA][o1];
        PX(ddinc) = ddinc;
        PX(ddmax_c) = ddmax_c;
        PX(ddnop_c) = ddnop_c;
        PX(ddthr) = ddthr;
        PX(k0r) = k0r;
        PX(k1r) = k1r;
        PX(kpr) = kpr;
        PX(k2r) = k2r;
        PX(tthr) = a.P[MHM_P_TEMPTHRESH][o1];
      }
#if MHM_FAST
      else {
        const double fS = a.P[MHM_P_FSEALED][o1];
        p.set(PID::kSeal2, fS, 1.0 - fS);
        p.set(PID::kDd, ddnop_c, ddinc);
        p.set(PID::kDdBase, ddmax_c, k2r);
        p.set(PID::kUnsat, a.P[MHM_P_UNSATTHRESH][mc], k0r);
        p.set(PID::kSlow, k1r, 1.0 + a.P[MHM_P_ALPHA][o1]);
        p.set(PID::kPerc, kpr, a.P[MHM_P_KARSTLOSS][mc]);
        p.sety(PID::kPetTthr, a.P[MHM_P_TEMPTHRESH][o1]);
      }
#endif
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const size_t oh = (((size_t)member * a.nLC + y) * NH + h) * n + c;  // (n, nH, nLC)
        const double FC = a.P[MHM_P_SOILMOISTFC][oh], SAT = a.P[MHM_P_SOILMOISTSAT][oh];
        const double WP = a.P[MHM_P_WILTINGPOINT][oh];
        if constexpr (!kPaired) {
          PH(fRoots, h) = a.P[MHM_P_FROOTS][oh];
          PH(FC, h) = FC;
          PH(SAT, h) = SAT;
          PH(EXPN, h) = a.P[MHM_P_SOILMOISTEXP][oh];
          PH(WP, h) = WP;
#if MHM_FAST
          PH(inv_SAT, h) = 1.0 / SAT;
          PH(inv_FCWP, h) = 1.0 / (FC - WP);
#endif
        }
#if MHM_FAST
        else {
          const bool feddes = a.soil_case == 1 || a.soil_case == 4;
          const double range = feddes ? FC - WP : (SAT - WP) * a.P[MHM_P_JARVIS_C1][mc];
          p.set(PID::kSat0 + h, SAT, 1.0 / SAT);
          if (h & 1) {
            p.sety(PID::kExp0 + h / 2, a.P[MHM_P_SOILMOISTEXP][oh]);
            p.sety(PID::kWp0 + h / 2, WP);
          } else {
            p.set(PID::kExp0 + h / 2, a.P[MHM_P_SOILMOISTEXP][oh], 0.0);
            p.set(PID::kWp0 + h / 2, WP, 0.0);
          }
          p.set(PID::kRoot0 + h, a.P[MHM_P_FROOTS][oh], 1.0 / range);
        }
#endif
        if (t == 0 && a.tt_first == 1 && !a.read_states) s.sm[h] = 0.5 * FC;  // mo_mhm.f90:448-450
      }
      cu.cur_l = -1;  // petLAIcorFactor / aeroResist also depend on yId
    }
    if ((!UNIFORM || t == 0) && il != cu.cur_l) {  // LAI step changed: mo_common_datetime_type.f90:135-155
      cu.cur_l = il;
      const double maxInter = a.P[MHM_P_MAXINTER][((size_t)member * a.nLAI + il) * n + c];
      double petFac = 1.0;
      if (a.pet_case == -1) {
        petFac = a.P[MHM_P_PETLAICORFACTOR][(((size_t)member * a.nLC + y) * a.nLAI + il) * n + c];
      } else if (a.pet_case == 0 || a.pet_case == 1) {
        petFac = a.P[MHM_P_FASP][mc];
      }
      if constexpr (!kPaired) {
        PX(maxInter) = maxInter;
#if MHM_FAST
        PX(inv_maxInter) = 1.0 / maxInter;
#endif
        PX(petFac) = petFac;
      }
#if MHM_FAST
      else {
        p.set(PID::kMaxInter, maxInter, maxInter > kEps ? 1.0 / maxInter : 0.0);
        p.setx(PID::kPetTthr, petFac);
      }
#endif
    }
  };
  auto pet_factor = [&]() -> double {
#if MHM_FAST
    if constexpr (kPaired) return P2(kPetTthr).x;
    else
#endif
      return PX(petFac);
  };
  // node runoff of L11_runoff_acc from the step's total runoff (fused routing input)
  auto node_runoff = [&](const double total_runoff) -> double {
#if MHM_FAST
    return total_runoff * qscale;
#else
    const double r = 0.0 + total_runoff;
    const double v = a.qout_map_flag ? (0.0 + r * qarea) : r * qarea;
    return v * 1000.0 / a.qout_tst;
#endif
  };
  // the step's total runoff goes to the history row and / or the tiled node-runoff history
  auto put = [&](const double total_runoff) {
    if (UNIFORM ? !FUSED : cu.hist != nullptr) {
      if (live) __stcs(cu.hist, total_runoff);
      cu.hist += hist_stride;
    }
    if (UNIFORM ? FUSED : qout != nullptr) {
      const double v = node_runoff(total_runoff);
      if (live) *cu.qp = v;
      ++cu.qst;
      cu.qp += (cu.qst & 7) ? (size_t)1 : qtile_stride - 7;
    }
  };

  // one model step; EMIT is a compile-time tag so that steps 1..n-1 carry no flux stores
  auto step = [&](auto emit_tag, const int t) {
    constexpr bool EMIT = decltype(emit_tag)::value;
    const StepIdx si = a.idx_in[UNIFORM ? 0 : t];  // kernel-parameter space: uniform constant loads
    const int y = si.yId - 1, il = si.iLAI - 1, month = si.month - 1;

    load_params(t);

    // ---- forcing of this step: mo_meteo_handler.f90:595-618 (iMeteoTS); rows are
    //      [meteo step][cell]; the row of step t was loaded during step t-1 ----
    const long long row = cu.cur_row;
    const double raw_pre = cu.raw_pre, raw_temp = cu.raw_temp, raw_pet = cu.raw_pet;

    // ---- get_corrected_pet :1053-1119 ----
    double pet;
    if (kHourlyPetIn || a.pet_case <= 0) {
      pet = pet_factor() * raw_pet;
    } else if (a.pet_case == 1) {
      const double tmx = a.met[MHM_M_TMAX][(size_t)(row - a.met_first[MHM_M_TMAX]) * n + c];
      const double tmn = a.met[MHM_M_TMIN][(size_t)(row - a.met_first[MHM_M_TMIN]) * n + c];
      pet = pet_factor() * pet_hargreaves(a.P[MHM_P_HARSAMCOEFF][mc], raw_temp, tmx, tmn,
                                          a.P[MHM_P_LATITUDE][mc], si.doy);
    } else if (a.pet_case == 2) {
      const double rn = a.met[MHM_M_NETRAD][(size_t)(row - a.met_first[MHM_M_NETRAD]) * n + c];
      pet = pet_priestly(a.P[MHM_P_PRIETAYALPHA][((size_t)member * a.nLAI + il) * n + c],
                         fmax(rn, 0.0), raw_temp);
    } else {
      const double rn = a.met[MHM_M_NETRAD][(size_t)(row - a.met_first[MHM_M_NETRAD]) * n + c];
      const double avp =
          a.met[MHM_M_ABSVAPPRESS][(size_t)(row - a.met_first[MHM_M_ABSVAPPRESS]) * n + c];
      const double ws =
          a.met[MHM_M_WINDSPEED][(size_t)(row - a.met_first[MHM_M_WINDSPEED]) * n + c];
      const double ar =
          a.P[MHM_P_AERORESIST][(((size_t)member * a.nLC + y) * a.nLAI + il) * n + c];
      const double sr = a.P[MHM_P_SURFRESIST][((size_t)member * a.nLAI + il) * n + c];
      pet = pet_penman(fmax(rn, 0.0), raw_temp, avp / 1000.0, ar / ws, sr);
    }
    // ---- temporal disaggregation: mo_meteo_temporal_tools.f90 ----
    double pet_calc, temp_calc, prec_calc;
    if (kHourlyPetIn || a.is_hourly) {
      pet_calc = pet;
      temp_calc = raw_temp;
      prec_calc = raw_pre;
    } else if (a.read_weights) {
      const size_t wo = ((size_t)si.hour * 12 + month) * n + c;
      pet_calc = (pet + 0.0) * a.w_pet[wo] - 0.0;
      temp_calc = (raw_temp + kT0) * a.w_temp[wo] - kT0;
      prec_calc = (raw_pre + 0.0) * a.w_pre[wo] - 0.0;
    } else if (a.nTstepDay_dp > 1.0) {
      const double fpet = si.isday ? a.tab.fday_pet[month] : a.tab.fnight_pet[month];
      const double fpre = si.isday ? a.tab.fday_prec[month] : a.tab.fnight_prec[month];
      const double ftmp = si.isday ? a.tab.fday_temp[month] : a.tab.fnight_temp[month];
      pet_calc = 2.0 * pet * fpet / a.nTstepDay_dp;
      prec_calc = 2.0 * raw_pre * fpre / a.nTstepDay_dp;
      temp_calc = raw_temp + ftmp;
    } else {
      pet_calc = pet;
      temp_calc = raw_temp;
      prec_calc = raw_pre;
    }
    FluxCapture cap;
    const FluxEmitter<EMIT, OUT != 0> emit{a.F, mc, n, (size_t)member * NH * n + c, a.write_fluxes && live,
                                      &cap};
    emit(MHM_F_PET_CALC, pet_calc);
    emit(MHM_F_TEMP_CALC, temp_calc);
    emit(MHM_F_PREC_CALC, prec_calc);

    // the next step's forcing row is requested before this step's arithmetic
    if (t + 1 < a.nSteps) {
      if constexpr (UNIFORM) {  // consecutive rows
        cu.ppre += n;
        cu.ptemp += n;
        cu.ppet += n;
        cu.raw_pre = ldg_stream(cu.ppre);
        cu.raw_temp = ldg_stream(cu.ptemp);
        cu.raw_pet = ldg_stream(cu.ppet);
      } else {
        const long long nrow = (long long)a.idx_in[t + 1].iMeteoTS;
        if (nrow != row) {
          const size_t adv = (size_t)(nrow - row) * n;
          cu.cur_row = nrow;
          cu.ppre += adv;
          cu.ptemp += adv;
          cu.raw_pre = ldg_stream(cu.ppre);
          cu.raw_temp = ldg_stream(cu.ptemp);
          if (kHourlyPetIn || a.pet_case <= 0) {
            cu.ppet += adv;
            cu.raw_pet = ldg_stream(cu.ppet);
          }
        }
      }
    }

    double total_runoff;
#if MHM_FAST
    if constexpr (kPaired) {
      total_runoff = cascade_step_sel<NH, EMIT>(p, s, pet_calc, temp_calc, prec_calc,
                                                a.tab.inv_evap_coeff[month], warp_tasks, sh_tab, emit);
    } else {
      if constexpr (!kTasksBySource)
        total_runoff = cascade_step<NH, VARIANT, EMIT, OUT != 0>(p, s, pet_calc, temp_calc, prec_calc, a.soil_case,
                                                            a.tab.evap_coeff[month], a.tab.inv_evap_coeff[month],
                                                            warp_tasks.xy, sh_tab, emit);
    }
#else
    total_runoff = cascade_step<NH, VARIANT, EMIT, OUT != 0>(p, s, pet_calc, temp_calc, prec_calc, a.soil_case,
                                                        a.tab.evap_coeff[month], emit);
#endif

    if (OUT) {
      if (live && t >= a.out_first) {
        const int yo = a.out_yid[t] - 1;
        if (yo != out_y) {  // scene the driver holds after the step's date increment
          out_y = yo;
          out_fS = a.P[MHM_P_FSEALED][((size_t)member * a.nLC + yo) * n + c];
#pragma unroll
          for (int h = 0; h < NH; ++h)
            out_sat[h] = a.P[MHM_P_SOILMOISTSAT][(((size_t)member * a.nLC + yo) * NH + h) * n + c];
        }
        if constexpr (OUT == 2) accumulate_outputs_default<NH>(oacc, cap, s, out_fS, out_sat);
        else if (a.out_mask) accumulate_outputs<NH>(a.out_mask, out_acc_l, cap, s, out_fS, out_sat);
        if (a.agg_mask)
          accumulate_aggregates<NH>(a.agg_mask, a.agg_nhor_sm, a.agg_col, a.bfi_acc, mc, hist_stride, cap, s,
                                    out_fS, out_sat);
      }
    }
    put(total_runoff);
  };

  const int n_last = a.nSteps - 1;
#if MHM_FAST
  if constexpr (UNIFORM && kPaired && OUT != 1) {
    // Software pipeline over the steps of a uniform-calendar launch: stage A of step t+1 (canopy,
    // snow, sealed store -- it needs only the forcing and three states stage B never touches)
    // is issued in the same basic block as stage B2 of step t (horizons, reservoirs), so the
    // two dependent fp64 chains overlap.  Same operations on the same values as step().
    // The three forcing arrays are walked with ONE 32-bit index (row * nCells + cell, relative to
    // the launch's first row; the host keeps nSteps * nCells below 2^32).
    load_params(0);
    const long long row0 = a.idx_in[0].iMeteoTS;
    const double* bpre = a.met[MHM_M_PRE] + (size_t)(row0 - a.met_first[MHM_M_PRE]) * n;
    const double* btemp = a.met[MHM_M_TEMP] + (size_t)(row0 - a.met_first[MHM_M_TEMP]) * n;
    const double* bpet = a.met[MHM_M_PET] + (size_t)(row0 - a.met_first[MHM_M_PET]) * n;
    // (opaque to the optimiser: base + 8 * index is then one IMAD.WIDE per load)
    asm volatile("" : "+l"(bpre), "+l"(btemp), "+l"(bpet));
    unsigned fi = (unsigned)c;
    double nx_pre = 0.0, nx_temp = 0.0, nx_pet = 0.0;
    __shared__ std::conditional_t<TMA, ForcingRing, EmptyParams> fring;
    int f_row = 0;  // next forcing row (relative to the launch's first) to hand out
    const int n_rows = a.nSteps;
    // TMA: lane 0 requests row r of the warp's 32 cells (rows r - kFStages .. have been consumed)
    auto tma_request = [&](const int r) {
      if constexpr (TMA) {
        const int warp = threadIdx.x >> 5, sl = r % kFStages;
        const int cell0 = (blockIdx.x / a.nMembers) * kCellThreads + warp * 32;
        const int cnt = min(32, max(0, a.nCells - cell0));
        const unsigned bar = smem_addr(&fring.full[warp][sl]);
        mbar_expect_tx(bar, (unsigned)(3 * cnt * 8));
        if (cnt > 0) {
          const size_t off = (size_t)r * n + (size_t)cell0;
          tma_load_1d(smem_addr(&fring.v[sl][0][warp * 32]), bpre + off, (unsigned)(cnt * 8), bar);
          tma_load_1d(smem_addr(&fring.v[sl][1][warp * 32]), btemp + off, (unsigned)(cnt * 8), bar);
          tma_load_1d(smem_addr(&fring.v[sl][2][warp * 32]), bpet + off, (unsigned)(cnt * 8), bar);
        }
      }
    };
    if constexpr (TMA) {
      if ((threadIdx.x & 31) == 0) {
        for (int sl = 0; sl < kFStages; ++sl) mbar_init(smem_addr(&fring.full[threadIdx.x >> 5][sl]), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int r = 0; r < kFStages && r < n_rows; ++r) tma_request(r);
      }
      __syncwarp();
    } else {
      nx_pre = ldg_stream(bpre + fi);
      nx_temp = ldg_stream(btemp + fi);
      nx_pet = ldg_stream(bpet + fi);
    }
    // the next forcing row of this cell: (pre, temp, raw pet); `last_tag`: the launch's last row
    // (nothing is requested after it)
    auto next_row = [&](auto last_tag, double& pre, double& temp, double& pet) {
      constexpr bool LAST = decltype(last_tag)::value;
      if constexpr (TMA) {
        const int warp = threadIdx.x >> 5, sl = f_row % kFStages;
        mbar_wait(smem_addr(&fring.full[warp][sl]), (unsigned)((f_row / kFStages) & 1));
        pre = fring.v[sl][0][threadIdx.x];
        temp = fring.v[sl][1][threadIdx.x];
        pet = fring.v[sl][2][threadIdx.x];
        if constexpr (!LAST) {
          __syncwarp();  // every lane has read the slot before lane 0 lets the TMA unit overwrite it
          if ((threadIdx.x & 31) == 0 && f_row + kFStages < n_rows) tma_request(f_row + kFStages);
          ++f_row;
        }
      } else {
        pre = nx_pre;
        temp = nx_temp;
        pet = nx_pet;
        if constexpr (!LAST) {  // the row after it is requested (the last row is handed out with LAST)
          fi += (unsigned)a.nCells;
          nx_pre = ldg_stream(bpre + fi);
          nx_temp = ldg_stream(btemp + fi);
          nx_pet = ldg_stream(bpet + fi);
        }
      }
    };
    if (n_last > 0) {
      const int month = a.idx_in[0].month - 1;
      const double inv_ec = a.tab.inv_evap_coeff[month];
      // OUT == 2: the stages' fluxes are captured for the output sums; every step before the
      // launch's last one shares the launch's land-cover scene also after its date increment
      // (uniform calendar), so fSealed and the saturation depths are the parameter store's
      FluxCapture cap_a, cap_b;
      const FluxEmitter<false, OUT == 2> noemit{a.F, mc, n, (size_t)member * NH * n + c, false, &cap_a};
      const FluxEmitter<false, OUT == 2> noemit_b{a.F, mc, n, (size_t)member * NH * n + c, false, &cap_b};
      int a_step = 0;  // step of the next stage A
      // stage A of the next step on its forcing row
      auto stage_a_next = [&]() -> StageA {
        double pre, temp, pet;
        next_row(std::false_type{}, pre, temp, pet);
        const double2 pt = P2(kPetTthr);  // petFac, tempThresh: one load
        const double pet_calc = pt.x * pet;
        StageA r = cascade_stage_a_sel<NH, false>(p, s, pet_calc, temp, pre, pt.y, inv_ec, noemit);
        if constexpr (OUT == 2) {
          if (a_step >= a.out_first) accumulate_default_a<NH>(oacc, cap_a, s, pet_calc, P2(kSeal2).x);
          r.aet_canopy = cap_a.v[MHM_F_AETCANOPY];
          r.aet_sealed = cap_a.v[MHM_F_AETSEALED];
          ++a_step;
        }
        return r;
      };
      // stage B2 of step t (stage A result `in`), then the step's remaining output sums
      auto stage_b2 = [&](const StageA& in, const double (&frac_pre)[NH], const int t) -> double {
        const double r = cascade_stage_b2_sel<NH, false>(p, s, in, frac_pre, sh_tab, noemit_b);
        if constexpr (OUT == 2) {
          if (t >= a.out_first) {
            double sat_o[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) sat_o[h] = PH2(kSat0, h).x;
            accumulate_default_b<NH>(oacc, cap_b, s, P2(kSeal2).x, sat_o, in.aet_canopy, in.aet_sealed);
          }
        }
        return r;
      };
#if MHM_CELL_PIPE3
      // Three steps in flight.  With A = canopy / snow / sealed store, R = infiltration powers of
      // the warp (compacted, a chain of shared-memory round trips), H = soil horizons, V = unsaturated
      // and saturated zone + runoff, iteration t runs
      //      R(t+1)  ||  V(t)  ||  A(t+2)      -- one basic block, three independent chains
      //      H(t+1)
      // R(t+1) needs A(t+1) and the soil moisture H(t) left; V(t) needs what H(t) let through; the
      // three touch disjoint states.  Same operations on the same values as step().
      const int L = n_last;                // steps 0 .. L-1 go through the pipeline
      StageA sa0 = stage_a_next(), sa1;    // A(0), A(1)
      if (L > 1) sa1 = stage_a_next();
      double infil, rsl, frac_pre[NH];
      PowTasks<NH> pt;
      pow_post<NH>(p, s, sa0.prec_effect, warp_tasks, pt);  // R(0), H(0)
      pow_round0(p, warp_tasks, sh_tab);
      pow_collect<NH>(p, pt, warp_tasks, sh_tab, frac_pre);
      infil = cascade_horizons_sel<NH>(p, s, sa0, frac_pre, noemit);
      rsl = sa0.runoff_sealed;
      // cur = A(t+1) on entry; nxt receives A(t+2)
      auto iter = [&](auto with_a, const StageA& cur, StageA& nxt) {
        pow_post<NH>(p, s, cur.prec_effect, warp_tasks, pt);
        pow_round0(p, warp_tasks, sh_tab);
        put(cascade_reservoirs_sel<NH>(p, s, infil, rsl, sh_tab, noemit));
        if constexpr (decltype(with_a)::value) nxt = stage_a_next();
        pow_collect<NH>(p, pt, warp_tasks, sh_tab, frac_pre);
        infil = cascade_horizons_sel<NH>(p, s, cur, frac_pre, noemit);
        rsl = cur.runoff_sealed;
      };
      int t = 0;
      for (; t + 3 < L; t += 2) {  // two iterations per trip: the hand-over of A costs no moves
        iter(std::true_type{}, sa1, sa0);
        iter(std::true_type{}, sa0, sa1);
      }
      if (t + 2 < L) {
        iter(std::true_type{}, sa1, sa0);
        sa1 = sa0;
        ++t;
      }
      if (t + 1 < L) iter(std::false_type{}, sa1, sa0);  // H(L-1): no forcing row left for the pipeline
      put(cascade_reservoirs_sel<NH>(p, s, infil, rsl, sh_tab, noemit));  // V(L-1)
#else
      // Two steps in flight: stage A of step t+1 is issued in the same basic block as the horizons
      // and reservoirs of step t, so the two dependent fp64 chains overlap.
      StageA sa = stage_a_next();
      int t = 0;
      for (; t + 2 < n_last; t += 2) {  // unrolled by two: the hand-over sa <- sb costs no moves
        double frac_pre[NH];
        cascade_stage_b1_sel<NH>(p, s, sa.prec_effect, warp_tasks, sh_tab, frac_pre);
        const StageA sb = stage_a_next();
        put(stage_b2(sa, frac_pre, t));
        cascade_stage_b1_sel<NH>(p, s, sb.prec_effect, warp_tasks, sh_tab, frac_pre);
        sa = stage_a_next();
        put(stage_b2(sb, frac_pre, t + 1));
      }
      for (; t + 1 < n_last; ++t) {
        double frac_pre[NH];
        cascade_stage_b1_sel<NH>(p, s, sa.prec_effect, warp_tasks, sh_tab, frac_pre);
        const StageA sb = stage_a_next();  // step t+1; row t+2 <= n_last requested
        put(stage_b2(sa, frac_pre, t));
        sa = sb;
      }
      double frac_pre[NH];
      cascade_stage_b1_sel<NH>(p, s, sa.prec_effect, warp_tasks, sh_tab, frac_pre);
      put(stage_b2(sa, frac_pre, n_last - 1));  // step n_last - 1
#endif
    }
    // the launch's last step through the general path (it stores the fluxes)
    next_row(std::true_type{}, cu.raw_pre, cu.raw_temp, cu.raw_pet);
  } else
#endif
  {
    cu.raw_pre = ldg_stream(cu.ppre);
    cu.raw_temp = ldg_stream(cu.ptemp);
    cu.raw_pet = cu.ppet ? ldg_stream(cu.ppet) : 0.0;
    for (int t = 0; t < n_last; ++t) step(std::false_type{}, t);
  }
  step(std::true_type{}, n_last);

  if (OUT == 1) {
    if (a.out_mask && live)
      for (int k = 0; k < a.out_nslots; ++k) a.out_acc[(size_t)k * hist_stride + mc] = out_acc_l[k * kCellThreads];
  }
  if (OUT == 2) {
    if (live) {
#pragma unroll
      for (int k = 0; k < 16 + 3 * NH; ++k) a.out_acc[(size_t)k * hist_stride + mc] = oacc[OUT == 2 ? k : 0];
    }
  }
  // ---- write back states ----
  if (live) {
    a.S[MHM_S_INTER][mc] = s.inter;
    a.S[MHM_S_SNOWPACK][mc] = s.snowpack;
    a.S[MHM_S_SEALSTW][mc] = s.sealed;
    a.S[MHM_S_UNSATSTW][mc] = s.unsat;
    a.S[MHM_S_SATSTW][mc] = s.sat;
#pragma unroll
    for (int h = 0; h < NH; ++h) a.S[MHM_S_SOILMOIST][((size_t)member * NH + h) * n + c] = s.sm[h];
  }
}

}  // namespace mhm
