/*
 * mhm_oracle.h -- CPU restatement ("oracle") of the mHM L1 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / the CPU arm that is timed beside
 * the GPU.  The product (mhm_b200/, libmhm_cuda.so) never links or calls it.
 *
 * The reference (mhm-ufz/mHM 5.13.2-dev0, Fortran 2008 + FORCES v0.6.0) cannot be
 * compiled in the build container (no Fortran compiler, no NetCDF, FORCES not
 * vendored).  Every function here restates one reference routine with the same
 * operation order, compiled with -ffp-contract=off so that no FMA contraction
 * changes a rounding.  Reference file:line is cited on every function
 * (paths relative to /root/reference/src).
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - pinned by the reference's own pFUnit known-answer tests for canopy, snow,
 *     soil moisture, Feddes/Jarvis, runoff, PET, temporal disaggregation, grid
 *     (tests/test_oracle_kat.py transcribes src/tests/ *.pf);
 *   - routing / MPR: no unit-level vectors exist upstream -> see DESIGN.md.
 *
 * Struct layout rule: one "type name;" or "type name[N];" per line so that
 * tests/orc.py can mirror the struct in ctypes by parsing this header.
 */
#ifndef MHM_ORACLE_H
#define MHM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: FORCES v0.6.0 mo_constants (not vendored; values per SURVEY 8c,
 *      checked against test_pet.pf) and mo_common_constants.f90:25-31 ---------- */
#define ORC_EPS_DP      2.220446049250313e-16  /* epsilon(1.0_dp) */
#define ORC_NODATA_DP   (-9999.0)
#define ORC_NODATA_I4   (-9999)
#define ORC_TWOTHIRD    0.6666666666666666666666666666666666667
#define ORC_PI          3.141592653589793238462643383279502884197
#define ORC_TWOPI       6.283185307179586476925286766559005768394
#define ORC_DEG2RAD     (ORC_PI / 180.0)
#define ORC_T0          273.15
#define ORC_DAYSECS     86400.0
#define ORC_HOURSECS    3600.0
#define ORC_YEARDAYS    365.0
#define ORC_SOLARCONST  1367.0
#define ORC_SPECHEATET  2.45e06
#define ORC_PSYCHRO     0.0646
#define ORC_CP0         1005.0
#define ORC_RHO0        1.225
/* mo_mhm_constants.f90:36-48 */
#define ORC_HARSAMCONST 17.800
#define ORC_DUFFIEDR    0.0330
#define ORC_DUFFIEDELTA1 0.4090
#define ORC_DUFFIEDELTA2 1.3900
#define ORC_TETENS_C1   0.6108
#define ORC_TETENS_C2   17.270
#define ORC_TETENS_C3   237.30
#define ORC_SATPRESSURESLOPE1 4098.0

/* ---- scalar process routines ------------------------------------------------ */
void orc_canopy_interc(double pet, double interc_max, double precip,
                       double *interc, double *throughfall, double *evap_canopy);
void orc_snow_accum_melt(double deg_day_incr, double deg_day_max, double deg_day_noprec,
                         double prec, double temperature, double temperature_thresh,
                         double thrfall, double *snow_pack, double *deg_day, double *melt,
                         double *prec_effect, double *rain, double *snow);
double orc_feddes_et_reduction(double soil_moist, double soil_moist_FC, double wilting_point,
                               double frac_roots);
double orc_jarvis_et_reduction(double soil_moist, double soil_moist_sat, double wilting_point,
                               double frac_roots, double jarvis_thresh_c1);
/* horizon arrays are strided (stride in doubles) so that the caller can pass
 * Fortran (cell, horizon) sections without copying */
void orc_soil_moisture(int32_t processCase, double frac_sealed, double water_thresh_sealed,
                       double pet, double evap_coeff, int32_t nH, int64_t stride,
                       const double *soil_moist_sat, const double *frac_roots,
                       const double *soil_moist_FC, const double *wilting_point,
                       const double *soil_moist_exponen, double jarvis_thresh_c1,
                       double aet_canopy, double prec_effec, double *runoff_sealed,
                       double *storage_sealed, double *infiltration, double *soil_moist,
                       double *aet, double *aet_sealed);
void orc_runoff_unsat_zone(double k1, double kp, double k0, double alpha, double karst_loss,
                           double pefec_soil, double unsat_thresh, double *sat_storage,
                           double *unsat_storage, double *slow_interflow,
                           double *fast_interflow, double *perc);
void orc_runoff_sat_zone(double k2, double *sat_storage, double *baseflow);
void orc_L1_total_runoff(double fSealed, double fast_interflow, double slow_interflow,
                         double baseflow, double direct_runoff, double *total_runoff);

double orc_pet_hargreaves(double HarSamCoeff, double HarSamConst, double tavg, double tmax,
                          double tmin, double latitude, int32_t doy);
double orc_pet_priestly(double PrieTayParam, double Rn, double tavg);
double orc_pet_penman(double net_rad, double tavg, double act_vap_pressure,
                      double aerodyn_resistance, double bulksurface_resistance, double a_s,
                      double a_sh);
double orc_extraterr_rad_approx(int32_t doy, double latitude);
double orc_slope_satpressure(double tavg);
double orc_sat_vap_pressure(double tavg);

double orc_temporal_disagg_meteo_weights(double meteo_val_day, double meteo_val_weights,
                                         double weights_correction);
double orc_temporal_disagg_flux_daynight(int32_t isday, double ntimesteps_day,
                                         double meteo_val_day, double fday, double fnight);
double orc_temporal_disagg_state_daynight(int32_t isday, double ntimesteps_day,
                                          double meteo_val_day, double fday, double fnight,
                                          int32_t add_correction);

/* ---- calendar (FORCES mo_julian restated: Gregorian julday/caldat) ----------- */
int32_t orc_julday(int32_t dd, int32_t mm, int32_t yy);
void orc_caldat(int32_t julian, int32_t *dd, int32_t *mm, int32_t *yy);
int32_t orc_doy(int32_t dd, int32_t mm, int32_t yy);

/* ---- routing ------------------------------------------------------------------ */
void orc_L11_runoff_acc(int32_t nCells1, int32_t nNodes, const double *qAll,
                        const double *efecArea, const int32_t *L1_L11_Id,
                        const double *L11_areaCell, const int32_t *L11_L1_Id, int32_t TS,
                        int32_t map_flag, double *qAcc);
void orc_add_inflow(int32_t nInflowGauges, const int32_t *InflowIndexList,
                    const int32_t *InflowHeadwater, const int32_t *InflowNodeList,
                    const double *QInflow, double *qOut);
void orc_L11_routing(int32_t nNodes, int32_t nLinks, const int32_t *netPerm,
                     const int32_t *fromN, const int32_t *toN, const double *C1,
                     const double *C2, const double *qOUT, int32_t nInflowGauges,
                     const int32_t *InflowHeadwater, const int32_t *InflowNodeList,
                     double *qTIN /*[2][nNodes]*/, double *qTR /*[2][nNodes]*/, double *Qmod);
void orc_reg_rout(const double *param5, int32_t nLinks, int32_t nSlope, const double *length,
                  const double *slope, const double *fFPimp, double TS, double *C1, double *C2);
/* mrm_update_param, processCase 2 (constant celerity); returns chosen TSrout [s] */
double orc_mrm_update_param_case2(int32_t nNodes, int32_t nOutlets, const double *length,
                                  double celerity, double *C1, double *C2);

/* ---- routing case 3 (river-slope celerity): literal restatements, test infrastructure ---- */
/* L0_streamNet as left by L11_stream_features (mRM/mo_mrm_net_startup.f90:1370-1412) and the
 * 40th-percentile floor of the link lengths for cases 2/3 (:1440-1446).  2-D index (i, j) of a
 * Fortran (nrows0, ncols0) array is at [(j-1)*nrows0 + i-1]; fDir0 / id0 / streamNet0 are 2-D. */
/* L11_routing_order (mRM/mo_mrm_net_startup.f90:765-842) for networks too large for the
 * reference's O(nLinks^2) sweeps: the same rOrder / netPerm in linear time.  Used by the CPU arms of
 * bench.py to build the 1M-node network without touching the product library; checked against a
 * loop-for-loop transcription of the reference on random forests (tests/test_oracle_run.py). */
int32_t orc_routing_order_linear(int32_t nNodes, int32_t nLinks, const int32_t *fromN, const int32_t *toN,
                                 int32_t *rOrder, int32_t *netPerm);
void orc_stream_net(int32_t nrows0, int32_t ncols0, const int32_t *fDir0_2d, int32_t nLinks,
                    const int32_t *netPerm, const int32_t *fRow, const int32_t *fCol,
                    const int32_t *tRow, const int32_t *tCol, int32_t *streamNet0_2d);
void orc_length_floor(int32_t nNodes, double *length);
/* FORCES mo_mad::mad(arr, z, mask, tout='u', mval) (un-vendored; semantics pinned by check/case_13) */
void orc_mad_upper(int32_t n, double *arr, double z, const int32_t *mask, double mval);
/* L11_calc_celerity (mRM/mo_mrm_net_startup.f90:2212-2423); slope0 / streamNet0 packed */
void orc_calc_celerity(int32_t nrows0, int32_t ncols0, const int32_t *mask0_2d, const int32_t *fDir0_2d,
                       const int32_t *streamNet0_packed, const double *slope0_packed, int32_t nNodes,
                       int32_t nLinks, const int32_t *netPerm, const int32_t *fRow, const int32_t *fCol,
                       const int32_t *tRow, const int32_t *tCol, double param, double *celerity11);
/* mrm_update_param, processCase 3 (mRM/mo_mrm_mpr.f90:298-321); returns TSrout */
double orc_mrm_update_param_case3(int32_t nNodes, int32_t nOutlets, const double *length,
                                  const double *celerity11, double *C1, double *C2);
/* L11_flow_accumulation (mRM/mo_mrm_net_startup.f90:2022-2163), recursive like the reference */
void orc_flow_accumulation(int32_t nrows11, int32_t ncols11, const int32_t *mask11_2d,
                           const int32_t *fDir11_packed, const double *cellarea11_packed, double *fAcc11_packed);

/* ---- one domain, everything the time loop touches ---------------------------- */
typedef struct orc_domain {
  /* sizes */
  int32_t nCells;
  int32_t nH;
  int32_t nLAI;
  int32_t nLC;
  /* process switches: processMatrix(3,1), (5,1), (8,1) */
  int32_t pc_soil;
  int32_t pc_pet;
  int32_t pc_rout;
  int32_t read_states;
  /* time */
  int32_t timestep_h;
  int32_t nTstepDay;
  int32_t jul_start;
  int32_t nTimeSteps;
  int32_t warming_days;
  int32_t timeStep_LAI_input;
  int32_t lc_year_start;
  int32_t lc_nyears;
  const int32_t *LCyearId;
  /* meteo: forcing arrays are [nMeteoSteps][nCells] (Fortran (cell, step)) */
  int32_t nTstepForcingDay;
  int32_t is_hourly_forcing;
  int32_t read_meteo_weights;
  int32_t nMeteoSteps;
  const double *pre;
  const double *temp;
  const double *pet;
  const double *tmin;
  const double *tmax;
  const double *netrad;
  const double *absvappress;
  const double *windspeed;
  const double *pre_weights;
  const double *temp_weights;
  const double *pet_weights;
  double fday_prec[12];
  double fnight_prec[12];
  double fday_pet[12];
  double fnight_pet[12];
  double fday_temp[12];
  double fnight_temp[12];
  double evap_coeff[12];
  double c2TSTu;
  /* effective parameters, Fortran (cell, dim2, dim3) = C [dim3][dim2][nCells] */
  const double *fSealed;
  const double *alpha;
  const double *degDayInc;
  const double *degDayMax;
  const double *degDayNoPre;
  const double *fRoots;
  const double *maxInter;
  const double *karstLoss;
  const double *kFastFlow;
  const double *kSlowFlow;
  const double *kBaseFlow;
  const double *kPerco;
  const double *soilMoistFC;
  const double *soilMoistSat;
  const double *soilMoistExp;
  const double *jarvis_thresh_c1;
  const double *tempThresh;
  const double *unsatThresh;
  const double *sealedThresh;
  const double *wiltingPoint;
  const double *petLAIcorFactor;
  const double *fAsp;
  const double *HarSamCoeff;
  const double *PrieTayAlpha;
  const double *aeroResist;
  const double *surfResist;
  const double *latitude;
  /* states (inout) */
  double *inter;
  double *snowPack;
  double *sealSTW;
  double *soilMoist;
  double *unsatSTW;
  double *satSTW;
  /* fluxes of the last step (out) */
  double *pet_calc;
  double *temp_calc;
  double *prec_calc;
  double *aETSoil;
  double *aETCanopy;
  double *aETSealed;
  double *baseflow;
  double *infilSoil;
  double *fastRunoff;
  double *melt;
  double *percol;
  double *preEffect;
  double *rain;
  double *runoffSeal;
  double *slowRunoff;
  double *snow;
  double *throughfall;
  double *total_runoff;
  double *degDay;
  /* routing */
  int32_t do_routing;
  int32_t nNodes;
  int32_t nOutlets;
  int32_t map_flag;
  int32_t nGauges;
  int32_t nInflowGauges;
  int32_t nGaugesTotal;
  int32_t nInflowTotal;
  const double *L1_areaCell;
  const double *L11_areaCell;
  const int32_t *L1_L11_Id;
  const int32_t *L11_L1_Id;
  const int32_t *netPerm;
  const int32_t *fromN;
  const int32_t *toN;
  const int32_t *gaugeIndexList;
  const int32_t *gaugeNodeList;
  const int32_t *InflowGaugeIndexList;
  const int32_t *InflowGaugeHeadwater;
  const int32_t *InflowGaugeNodeList;
  const double *InflowQ;
  const double *L11_length;
  const double *L11_slope;
  const double *L11_nLinkFracFPimp;
  double rout_param[5];
  double L11_TSrout;
  double *L11_C1;
  double *L11_C2;
  double *L11_qOUT;
  double *L11_qTIN;
  double *L11_qTR;
  double *L11_qMod;
  /* outputs of a run */
  /* persistent scratch of the routing scheduler (mo_mhm_interface_run.f90:460-514) */
  double *RunToRout;
  double *InflowDischarge;
  int32_t nDays;
  /* outputs of a run: mRM_runoff is Fortran (nTimeSteps, nGaugesTotal) */
  double *mRM_runoff;
  double *flux_history;
  int32_t flux_history_stride;
  int32_t num_threads;
  /* gridded outputs: mHM_updateDataset + OutputVariable (mo_write_fluxes_states.f90:283-438,
   * mo_nc_output.f90:140-175).  out_flags[v-1] = outputFlxState(v), v = 1..21; slots follow the
   * order of mHM_updateDataset (orc_output_slots).  out_acc [nSlots][nCells] holds the open
   * window, out_win [out_max_windows][nSlots][nCells] the written ones, out_win_tt their step. */
  int32_t out_flags[21];
  int32_t timeStep_model_outputs;
  int32_t out_counter;
  int32_t out_nwin;
  int32_t out_max_windows;
  double *out_acc;
  double *out_win;
  int32_t *out_win_tt;
  /* calibration aggregates: mhm_interface_run_update_optisim (mo_mhm_interface_run.f90:745-861),
   * index 0 = soil moisture, 1 = evapotranspiration, 2 = total water storage.  The container
   * type optidata_sim lives in FORCES v0.6.0 mo_optimization_types (un-vendored, absent from
   * /root/reference): its published behaviour is restated in the .c (init: dataSim = 0,
   * averageTimestep = 1, averageCounter = 0; add; average; average_add; increment_counter;
   * average_per_timestep) -- PARITY UNPINNED for that part, no reference output holds dataSim.
   * opt_timestep = timeStepInput (-1 daily, -2 monthly, -3 yearly); opt_sm/et/tws = dataSim,
   * Fortran (nCells, opt_ntime[i]). */
  int32_t opt_on[3];
  int32_t opt_timestep[3];
  int32_t opt_ntime[3];
  int32_t opt_nhor_sm;
  int32_t opt_avg_ts[3];
  int32_t opt_avg_cnt[3];
  double *opt_sm;
  double *opt_et;
  double *opt_tws;
  /* BFI sums of the evaluation period (mo_mhm_interface_run.f90:630-636) */
  int32_t bfi_on;
  double bfi_qBF_sum;
  double bfi_qT_sum;
} orc_domain;

/* number of per-cell records written per step into flux_history (see .c) */
int32_t orc_flux_record_size(int32_t nH);

/* time loop: mo_mhm_eval.f90:136-150 + mo_mhm_interface_run.f90:341-638 restated.
 * Runs steps tt_first..tt_last (1-based, inclusive) continuing from the states in d.
 * tt_first must be 1 on the first call (state of the date stepping is recomputed
 * from tt, so any split of the time axis gives identical results). */
int32_t orc_run(orc_domain *d, int32_t tt_first, int32_t tt_last);
/* meteo_forcings_wrapper (meteo/mo_meteo_helper.f90:98-130): L2 chunk, Fortran (nr2, nc2, nT), ->
 * packed L1 (nCells1, nT) by spatial_aggregation / spatial_disaggregation
 * (meteo/mo_meteo_spatial_tools.f90:94-200, 313-377) or a plain copy; masks are 0/1 */
int32_t orc_meteo_l2_to_l1(const double *data2, int32_t nr2, int32_t nc2, int32_t nT, const int32_t *mask2,
                           double cellsize2, int32_t nr1, int32_t nc1, const int32_t *mask1,
                           double cellsize1, double *out_packed, double *out_grid);
/* slots of the enabled output variables in mHM_updateDataset order: var[s] in 1..21, hor[s] the
 * 0-based horizon of per-horizon variables (else -1), avg[s] = averaged over the window */
int32_t orc_output_slots(const int32_t *out_flags, int32_t nH, int32_t *var, int32_t *hor,
                         int32_t *avg);

/* per-step index vectors (month 1-12, hour, yId 1-based, iLAI 1-based, iMeteoTS 1-based,
 * isday, doy) for tt = 1..n; arrays of length n */
void orc_time_indices(const orc_domain *d, int32_t n, int32_t *month, int32_t *hour,
                      int32_t *yId, int32_t *iLAI, int32_t *iMeteoTS, int32_t *isday,
                      int32_t *doy, int32_t *year);

#ifdef __cplusplus
}
#endif
#endif
