/*
 * mhm_cuda.h -- C ABI of libmhm_cuda.so, the B200 (sm_100a) implementation of mHM's
 * L1 hot path: meteo prologue + cell process cascade, mRM Muskingum routing and the
 * MPR upscaling operators.
 *
 * Who binds this: the reference's Fortran driver through ISO_C_BINDING
 * (mhm_b200/fortran/mo_mhm_cuda.F90 holds the `interface ... bind(C)` blocks and
 * INTEGRATION.md shows the four patched call sites).  Every entry point names the
 * reference routine / call site it replaces (paths relative to /root/reference/src).
 *
 * Conventions
 *  - plain C types only; all arrays are caller-owned host memory unless the name ends
 *    in `_device`; the library never frees or keeps a host pointer after the call
 *    returns (except mhm_cuda_bind_*: kept until mhm_cuda_unregister_domain / finalize).
 *  - every function returns 0 on success, non-zero on failure; the message is available
 *    from mhm_cuda_last_error().  The Fortran shim turns non-zero into
 *    `call error_message('mhm_cuda: ', ...)` like every other fatal path of the reference.
 *  - Fortran array sections are passed as (base pointer of the whole module-global array,
 *    leading dimension ld = size(array,1), offset = s1-1 of the domain's first cell).
 *    Element (k, j, y) (1-based) of a (nCellsTot, dim2, dim3) array is
 *    base[(offset + k-1) + ld*((j-1) + dim2*(y-1))].
 *  - all ids that index Fortran arrays (yId, iLAI, node ids, netPerm ...) stay 1-based.
 *  - `member`: index (0-based) of an ensemble member = one parameter set / one
 *    mhm_eval(parameterset) evaluation (mHM/mo_mhm_eval.f90:94).  A plain run has
 *    nMembers = 1 and passes member = 0.
 *  - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef MHM_CUDA_H
#define MHM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mhm_cuda_context mhm_cuda_context; /* opaque; Fortran: type(c_ptr) */

/* ---------------------------------------------------------------------------------
 * B0  lifecycle.  Call sites: after mhm_initialize/mrm_init
 * (mHM/mo_mhm_interface.F90:169-174) and before restart writing (:462-476).
 * --------------------------------------------------------------------------------- */
int mhm_cuda_init(int device, mhm_cuda_context **ctx);      /* device < 0: current device */
int mhm_cuda_finalize(mhm_cuda_context *ctx);
const char *mhm_cuda_last_error(void);
const char *mhm_cuda_version(void);

/* sizes and process switches of one domain (common/mo_common_types.F90:52-88 Grid;
 * processMatrix as filled by MPR/mo_mpr_read_config.f90:390-988, passed verbatim,
 * column-major (nProcesses, 3)) */
typedef struct mhm_domain_config {
  int32_t nCells;        /* level1(iDomain)%nCells */
  int32_t nHorizons;     /* nSoilHorizons_mHM (1..8) */
  int32_t nLAI;          /* nLAI: 2nd dim of L1_maxInter */
  int32_t nLCscenes;     /* nLCoverScene */
  int32_t nMembers;      /* parameter sets evaluated side by side (>= 1) */
  int32_t nProcesses;    /* rows of processMatrix (11) */
  int32_t timestep_h;    /* timeStep [h] */
  int32_t read_states;   /* read_restart: skip the 0.5*FC soil moisture start */
  double c2TSTu;         /* mHM/mo_startup.f90:168, timeStep/24 */
  const int32_t *processMatrix;
} mhm_domain_config;

int mhm_cuda_register_domain(mhm_cuda_context *ctx, int32_t iDomain,
                             const mhm_domain_config *cfg);
int mhm_cuda_unregister_domain(mhm_cuda_context *ctx, int32_t iDomain);

/* effective parameters (MPR/mo_mpr_global_variables.f90:128-170).  Layout in the
 * reference: (nCellsTot, dim2, dim3). */
enum mhm_param_id {
  MHM_P_FSEALED = 0,      /* L1_fSealed        (:,1,nLC)    */
  MHM_P_ALPHA,            /* L1_alpha          (:,1,nLC)    */
  MHM_P_DEGDAYINC,        /* L1_degDayInc      (:,1,nLC)    */
  MHM_P_DEGDAYMAX,        /* L1_degDayMax      (:,1,nLC)    */
  MHM_P_DEGDAYNOPRE,      /* L1_degDayNoPre    (:,1,nLC)    */
  MHM_P_FROOTS,           /* L1_fRoots         (:,nH,nLC)   */
  MHM_P_MAXINTER,         /* L1_maxInter       (:,nLAI,1)   */
  MHM_P_KARSTLOSS,        /* L1_karstLoss      (:,1,1)      */
  MHM_P_KFASTFLOW,        /* L1_kFastFlow      (:,1,nLC)    */
  MHM_P_KSLOWFLOW,        /* L1_kSlowFlow      (:,1,nLC)    */
  MHM_P_KBASEFLOW,        /* L1_kBaseFlow      (:,1,nLC)    */
  MHM_P_KPERCO,           /* L1_kPerco         (:,1,nLC)    */
  MHM_P_SOILMOISTFC,      /* L1_soilMoistFC    (:,nH,nLC)   */
  MHM_P_SOILMOISTSAT,     /* L1_soilMoistSat   (:,nH,nLC)   */
  MHM_P_SOILMOISTEXP,     /* L1_soilMoistExp   (:,nH,nLC)   */
  MHM_P_JARVIS_C1,        /* L1_jarvis_thresh_c1 (:,1,1)    */
  MHM_P_TEMPTHRESH,       /* L1_tempThresh     (:,1,nLC)    */
  MHM_P_UNSATTHRESH,      /* L1_unsatThresh    (:,1,1)      */
  MHM_P_SEALEDTHRESH,     /* L1_sealedThresh   (:,1,1)      */
  MHM_P_WILTINGPOINT,     /* L1_wiltingPoint   (:,nH,nLC)   */
  MHM_P_PETLAICORFACTOR,  /* L1_petLAIcorFactor(:,nLAI,nLC) */
  MHM_P_FASP,             /* L1_fAsp           (:,1,1)      */
  MHM_P_HARSAMCOEFF,      /* L1_HarSamCoeff    (:,1,1)      */
  MHM_P_PRIETAYALPHA,     /* L1_PrieTayAlpha   (:,nLAI,1)   */
  MHM_P_AERORESIST,       /* L1_aeroResist     (:,nLAI,nLC) */
  MHM_P_SURFRESIST,       /* L1_surfResist     (:,nLAI,1)   */
  MHM_P_LATITUDE,         /* pack(level1%y, mask) (:,1,1)   */
  MHM_P_COUNT
};

/* upload one parameter array of one member: replaces the array-section arguments of
 * `call mhm(...)` (mHM/mo_mhm_interface_run.f90:394-457) and of get_corrected_pet
 * (:366-374).  Called after mpr_eval (:234) or read_restart_states (:224). */
int mhm_cuda_set_param(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                       int32_t param_id, const double *base, int64_t ld, int64_t offset,
                       int32_t dim2, int32_t dim3);

/* states (mHM/mo_global_variables.f90:130-136) */
enum mhm_state_id {
  MHM_S_INTER = 0,  /* L1_inter     */
  MHM_S_SNOWPACK,   /* L1_snowPack  */
  MHM_S_SEALSTW,    /* L1_sealSTW   */
  MHM_S_UNSATSTW,   /* L1_unsatSTW  */
  MHM_S_SATSTW,     /* L1_satSTW    */
  MHM_S_SOILMOIST,  /* L1_soilMoist (:, nH) */
  MHM_S_COUNT
};
int mhm_cuda_set_state(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                       int32_t state_id, const double *base, int64_t ld, int64_t offset);
int mhm_cuda_get_state(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                       int32_t state_id, double *base, int64_t ld, int64_t offset);
/* mHM/mo_init_states.f90:280-300 (fluxes_states_default_init) for all members, on device */
int mhm_cuda_states_default_init(mhm_cuda_context *ctx, int32_t iDomain,
                                 const double *HorizonDepth_mHM);

/* fluxes of the last executed step (mHM/mo_global_variables.f90:141-160) */
enum mhm_flux_id {
  MHM_F_PET_CALC = 0, MHM_F_TEMP_CALC, MHM_F_PREC_CALC, MHM_F_AETCANOPY, MHM_F_AETSEALED,
  MHM_F_BASEFLOW, MHM_F_FASTRUNOFF, MHM_F_MELT, MHM_F_PERCOL, MHM_F_PREEFFECT, MHM_F_RAIN,
  MHM_F_RUNOFFSEAL, MHM_F_SLOWRUNOFF, MHM_F_SNOW, MHM_F_THROUGHFALL, MHM_F_TOTAL_RUNOFF,
  MHM_F_DEGDAY,       /* L1_degDay(:,1,1), intent(out) of snow_accum_melt */
  MHM_F_AETSOIL,      /* L1_aETSoil  (:, nH) */
  MHM_F_INFILSOIL,    /* L1_infilSoil(:, nH) */
  MHM_F_COUNT
};
int mhm_cuda_get_flux(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                      int32_t flux_id, double *base, int64_t ld, int64_t offset);

/* ---------------------------------------------------------------------------------
 * meteo (meteo/mo_meteo_handler.f90).  Forcing arrays are L1_pre/L1_temp/... of shape
 * (nCellsTot, nMeteoSteps): cell-contiguous per meteo step.
 * --------------------------------------------------------------------------------- */
typedef struct mhm_meteo_config {
  int32_t pet_case;            /* processMatrix(5,1): -1, 0, 1, 2, 3 */
  int32_t nTstepForcingDay;    /* 1 = daily forcing, 24 = hourly */
  int32_t is_hourly_forcing;   /* self%is_hourly_forcing */
  int32_t read_meteo_weights;  /* self%read_meteo_weights */
  double fday_prec[12];
  double fnight_prec[12];
  double fday_pet[12];
  double fnight_pet[12];
  double fday_temp[12];
  double fnight_temp[12];
  double evap_coeff[12];       /* mhm.nml panEvapo */
} mhm_meteo_config;
int mhm_cuda_set_meteo_config(mhm_cuda_context *ctx, int32_t iDomain,
                              const mhm_meteo_config *cfg);

enum mhm_meteo_var {
  MHM_M_PRE = 0, MHM_M_TEMP, MHM_M_PET, MHM_M_TMIN, MHM_M_TMAX, MHM_M_NETRAD,
  MHM_M_ABSVAPPRESS, MHM_M_WINDSPEED, MHM_M_COUNT
};
/* upload meteo steps [first_step, first_step + n_steps) (1-based meteo time index
 * iMeteoTS, mo_meteo_handler.f90:607) of one variable; replaces what prepare_data
 * (:645-912) leaves in L1_<var>(s_meteo:e_meteo, :).  A chunk stays resident until
 * the next call for the same variable; run calls must stay inside the resident chunk. */
int mhm_cuda_set_meteo(mhm_cuda_context *ctx, int32_t iDomain, int32_t var,
                       const double *base, int64_t ld, int64_t offset, int64_t first_step,
                       int64_t n_steps);
/* same without waiting for the copy: `base` (ideally pinned memory) must stay unchanged until
 * the next mhm_cuda_run_steps / mhm_cuda_cell_step of the domain has been issued and
 * mhm_cuda_synchronize returned.  Uploads are double buffered and run on their own stream, so
 * the H2D of chunk c+1 overlaps the kernels of chunk c. */
int mhm_cuda_set_meteo_async(mhm_cuda_context *ctx, int32_t iDomain, int32_t var,
                             const double *base, int64_t ld, int64_t offset, int64_t first_step,
                             int64_t n_steps);
/* same, but `dev` is a device pointer to a dense [n_steps][nCells] array that the
 * caller keeps alive (zero copy; used when forcing is already resident in HBM) */
/* N3 forcing ingest: a chunk of the METEO grid (level 2) exactly as read from NetCDF -- Fortran
 * (nrows2, ncols2, n_steps), float64 or float32 (is_f32) -- is remapped to the packed L1 cells on
 * the device: spatial_aggregation (mean over the valid level-2 cells of an L1 cell, summed in the
 * reference's element order, meteo/mo_meteo_spatial_tools.f90:94-200), spatial_disaggregation
 * (value of the parent cell, :313-377) or plain packing for equal resolutions
 * (meteo/mo_meteo_helper.f90:98-130).  Replaces meteo_forcings_wrapper's remap + pack and the
 * upload of the packed L1 array; masks are int32 0/1 Fortran (nrows, ncols). */
int mhm_cuda_set_meteo_l2(mhm_cuda_context *ctx, int32_t iDomain, int32_t var, const void *data2,
                          int32_t is_f32, int32_t nrows2, int32_t ncols2, const int32_t *mask2,
                          double cellsize2, int32_t nrows1, int32_t ncols1, const int32_t *mask1,
                          double cellsize1, int64_t first_step, int64_t n_steps);
/* the resident L1 forcing of steps [first_step, first_step + n_steps), Fortran (nCells, n_steps) */
int mhm_cuda_get_meteo(mhm_cuda_context *ctx, int32_t iDomain, int32_t var, double *out, int64_t ld,
                       int64_t first_step, int64_t n_steps);
int mhm_cuda_set_meteo_device(mhm_cuda_context *ctx, int32_t iDomain, int32_t var,
                              const double *dev, int64_t first_step, int64_t n_steps);
/* ---------------------------------------------------------------------------------
 * the GPUs of one box, one process (MPI rank) per GPU -- the reference distributes whole domains
 * over ranks (common/mo_common_read_config.F90:416-437, common/mo_common_MPI_tools.F90:39-69).
 * The library owns an NCCL communicator: rank 0 calls mhm_cuda_comm_unique_id, the caller carries
 * the 128 opaque bytes to every rank (MPI_Bcast in the Fortran driver), then every rank calls
 * mhm_cuda_comm_init.  nranks = 1 needs no NCCL library.  libnccl.so.2 is opened at run time.
 * --------------------------------------------------------------------------------- */
#define MHM_CUDA_UNIQUE_ID_BYTES 128
int mhm_cuda_comm_unique_id(char *id128);
int mhm_cuda_comm_init(mhm_cuda_context *ctx, int32_t nranks, int32_t rank, const char *id128);
int mhm_cuda_comm_finalize(mhm_cuda_context *ctx);
int mhm_cuda_comm_info(mhm_cuda_context *ctx, int32_t *nranks, int32_t *rank, int32_t *nccl_version);
/* Forcing SHARED by all ranks (ensemble members / calibration parameter sets of ONE domain spread
 * over the GPUs, BASELINE config 5): like mhm_cuda_set_meteo_async, but every rank copies only its
 * own 1/nranks of the chunk's meteo steps from host memory -- rows first_row .. first_row+n_rows-1
 * (0-based, within the chunk) as mhm_cuda_meteo_shared_rows tells; the other rows of `base` are
 * never read and need not be valid -- and the ranks all-gather the chunk over NVLink on the upload
 * stream.  Collective: every rank of the communicator must call it with the same var / first_step /
 * n_steps, in the same order.  is_f32: `base` holds float32 (NetCDF forcing is usually stored so;
 * mo_read_nc.f90 widens on read): half the host and NVLink bytes, widened to float64 on the device. */
int mhm_cuda_meteo_shared_rows(int64_t n_steps, int32_t nranks, int32_t rank, int64_t *first_row,
                               int64_t *n_rows);
int mhm_cuda_set_meteo_shared(mhm_cuda_context *ctx, int32_t iDomain, int32_t var, const void *base,
                              int32_t is_f32, int64_t ld, int64_t offset, int64_t first_step,
                              int64_t n_steps);
/* host bytes of forcing this process has copied to the device for the domain so far (measurement) */
int mhm_cuda_meteo_h2d_bytes(mhm_cuda_context *ctx, int32_t iDomain, int64_t *bytes);
/* L1_pre_weights / L1_temp_weights / L1_pet_weights (nCellsTot, 12, 24); var = PRE/TEMP/PET */
int mhm_cuda_set_meteo_weights(mhm_cuda_context *ctx, int32_t iDomain, int32_t var,
                               const double *base, int64_t ld, int64_t offset);

/* ---------------------------------------------------------------------------------
 * time axis: common/mo_common_datetime_type.f90:71-155 restated inside the library so
 * that a block of steps can run without returning to the host.
 * --------------------------------------------------------------------------------- */
typedef struct mhm_time_config {
  int32_t jul_start;            /* simPer(iDomain)%julStart (Julian day number) */
  int32_t nTimeSteps;           /* domainDateTime%nTimeSteps */
  int32_t warming_days;         /* warmingDays(iDomain) */
  int32_t timeStep_LAI_input;   /* 0/1: iLAI = month; -1/-2/-3: running counter */
  int32_t lc_year_start;        /* first year covered by LCyearId */
  int32_t lc_nyears;
  const int32_t *LCyearId;      /* LCyearId(year, iDomain), 1-based scene ids */
} mhm_time_config;
int mhm_cuda_set_time(mhm_cuda_context *ctx, int32_t iDomain, const mhm_time_config *cfg);

/* per-step indices as the reference derives them; host helper (no GPU needed) so the
 * Fortran side / tests can cross-check the library's calendar */
typedef struct mhm_step_index {
  int32_t iMeteoTS;   /* 1-based */
  int32_t year;
  int16_t yId;        /* 1-based land-cover scene */
  int16_t iLAI;       /* 1-based */
  int16_t doy;
  int8_t month;       /* 1..12 */
  int8_t hour;        /* 0..23 */
  int8_t isday;       /* hour > 6 .and. hour <= 18 */
  int8_t pad_[3];
} mhm_step_index;
int mhm_time_indices(const mhm_time_config *cfg, int32_t timestep_h, int32_t nTstepForcingDay,
                     int32_t tt_first, int32_t n_steps, mhm_step_index *out);

/* ---------------------------------------------------------------------------------
 * B1  one model step (per-step parity seam).  Replaces get_corrected_pet + get_temp +
 * get_prec + `call mhm(...)` (mHM/mo_mhm_interface_run.f90:366-457).  All fluxes of the
 * step are left on the device (mhm_cuda_get_flux / mhm_cuda_sync_to_host).
 * --------------------------------------------------------------------------------- */
int mhm_cuda_cell_step(mhm_cuda_context *ctx, int32_t iDomain, int32_t tt,
                       const mhm_step_index *idx);

/* B2  a block of model steps tt_first .. tt_first+n_steps-1 (performance seam; replaces
 * the TimeLoop body mHM/mo_mhm_eval.f90:136-150 for those steps, including the routing
 * schedule mo_mhm_interface_run.f90:460-612 when a network is set).  Fluxes of the
 * block's last step are left on the device, the gauge series in the runoff buffer. */
int mhm_cuda_run_steps(mhm_cuda_context *ctx, int32_t iDomain, int32_t tt_first,
                       int32_t n_steps);
/* variant selection: 0 = strict (no FMA contraction, IEEE division, literal formulas;
 * default), 1 = fast (same algorithm, FMA + hoisted reciprocals; <= 1e-9 relative) */
int mhm_cuda_set_math_mode(mhm_cuda_context *ctx, int32_t mode);

/* ---------------------------------------------------------------------------------
 * A10 / N1  gridded outputs accumulated on the device while blocks of steps run.
 * Replaces, for run_steps, the per-step `call mHM_updateDataset(...)`
 * (mHM/mo_write_fluxes_states.f90:283-438, called from mo_mhm_interface_run.f90:690-717)
 * and OutputVariable%updateVariable / writeVariableTimestep (common/mo_nc_output.f90:140-175):
 * every step with tIndex_out > 0 adds the enabled variables (same derived expressions, same
 * land-cover scene -- the one AFTER the date increment -- and the same sequential summation
 * in time) to the open window; a window closes where datetimeinfo%writeout
 * (common/mo_common_datetime_type.f90:157-184) fires, averaged for the state variables
 * 1-8, summed for the fluxes.  outputFlxState is mhm_outputs.nml's array (1..21; 18 =
 * neutrons is not supported), timeStep_model_outputs its window selector (-3 yearly,
 * -2 monthly, -1 daily, 0 end of run, > 0 every n steps).  After a run_steps call the
 * windows it closed can be fetched; the Fortran writer then calls nc%setData for each.
 * --------------------------------------------------------------------------------- */
int mhm_cuda_set_outputs(mhm_cuda_context *ctx, int32_t iDomain, const int32_t *outputFlxState,
                         int32_t timeStep_model_outputs);
/* number of windows closed by the last run_steps call; tt_end[w] = model step that closed it */
int mhm_cuda_get_output_windows(mhm_cuda_context *ctx, int32_t iDomain, int32_t *n_windows,
                                int32_t *tt_end, int32_t capacity);
/* one variable (1..21; horizon 1..nH for variables 3, 4, 17, 19, else 0) of one closed
 * window of one member, out[nCells] */
int mhm_cuda_get_output(mhm_cuda_context *ctx, int32_t iDomain, int32_t member, int32_t window,
                        int32_t variable, int32_t horizon, double *out);

/* ---------------------------------------------------------------------------------
 * A10  calibration aggregates and BFI sums, accumulated on the device while blocks of steps run.
 * Replaces, for run_steps, `call mhm_interface_run_update_optisim(etOptiSim, twsOptiSim, ...,
 * smOptiSim)` (mHM/mo_mhm_interface_run.f90:745-861, called from mo_mhm_eval.f90:144) and the
 * BFI sums of mo_mhm_interface_run.f90:630-636.  The container optidata_sim (FORCES v0.6.0
 * mo_optimization_types, not vendored in the reference tree) is reproduced: dataSim starts at
 * zero; for every step of the evaluation period, in this order, the open slot is closed when
 * the step's date increment raised is_new_day / _month / _year (timeStepInput -1 / -2 / -3:
 * soil moisture and TWS are divided by the number of values added, ET only moves on), then
 * the step's value is added unless it is the run's last step:
 *   soil moisture  sum(soilMoist(:, 1:nSoilHorizons_sm_input)) / sum(soilMoistSat(:, 1:n, yId))
 *   ET             sum(aETSoil) * fNotSealed + aETCanopy + aETSealed * fSealed
 *   TWS            inter + snowPack + sealSTW + unsatSTW + satSTW, then + soilMoist(:, h) per horizon
 * nTime = size(dataObs, 2).  Neutrons are out of scope.  BFI: per member, sum over the steps
 * with tIndex_out > 0 of sum(L1_baseflow * CellArea) / nCells and the same for L1_total_runoff
 * (summed per cell over time first; <= 1e-13 relative to the reference's order).
 * --------------------------------------------------------------------------------- */
typedef struct mhm_optisim_config {
  int32_t sm_on;
  int32_t sm_timeStepInput;
  int32_t sm_nTime;
  int32_t nSoilHorizons_sm_input;
  int32_t et_on;
  int32_t et_timeStepInput;
  int32_t et_nTime;
  int32_t tws_on;
  int32_t tws_timeStepInput;
  int32_t tws_nTime;
  int32_t bfi_on;
} mhm_optisim_config;
int mhm_cuda_set_optisim(mhm_cuda_context *ctx, int32_t iDomain, const mhm_optisim_config *cfg);
/* which: 0 soil moisture, 1 ET, 2 TWS; dataSim(:, 1:nTime) of one member into the Fortran array
 * base(ld, nTime) starting at row offset */
int mhm_cuda_get_optisim(mhm_cuda_context *ctx, int32_t iDomain, int32_t member, int32_t which,
                         double *base, int64_t ld, int64_t offset);
/* cellArea = level1(iDomain)%CellArea, host, [nCells] */
int mhm_cuda_get_bfi_sums(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                          const double *cellArea, double *qBF_sum, double *qT_sum);

/* keep host module globals coherent (pybind get%L1_variable, restart writing): bind once,
 * then mhm_cuda_sync_to_host copies every bound state/flux of every member 0 array back */
int mhm_cuda_bind_host_state(mhm_cuda_context *ctx, int32_t iDomain, int32_t state_id,
                             double *base, int64_t ld, int64_t offset);
int mhm_cuda_bind_host_flux(mhm_cuda_context *ctx, int32_t iDomain, int32_t flux_id,
                            double *base, int64_t ld, int64_t offset);
int mhm_cuda_sync_to_host(mhm_cuda_context *ctx, int32_t iDomain);

/* total runoff history of the last run_steps block, Fortran (nCells, n_steps) per member
 * (what `RunToRout = L1_total_runoff(s1:e1)` saw at each step).  When the routing grid equals
 * the L1 grid the cell kernel writes the routing's node runoff directly (L11_runoff_acc,
 * mRM/mo_mrm_pre_routing.f90:110-141, fused) and no history is kept unless
 * mhm_cuda_keep_runoff_history(ctx, iDomain, 1) was called before run_steps. */
int mhm_cuda_keep_runoff_history(mhm_cuda_context *ctx, int32_t iDomain, int32_t keep);
int mhm_cuda_get_runoff_history(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                                double *out, int64_t ld);

/* ---------------------------------------------------------------------------------
 * B3  routing (mRM/mo_mrm_routing.f90:104-303, mo_mrm_pre_routing.f90:77-214,
 * mo_mrm_mpr.f90:61-119,241-329)
 * --------------------------------------------------------------------------------- */
typedef struct mrm_network {
  int32_t nNodes;               /* level11(iDomain)%nCells */
  int32_t nOutlets;             /* L11_nOutlets(iDomain) */
  int32_t map_flag;             /* ge(resolutionRouting, resolutionHydrology) */
  int32_t nGauges;              /* domain_mrm%nGauges */
  int32_t nInflowGauges;        /* domain_mrm%nInflowGauges */
  int32_t nGaugesTotal;         /* size(mRM_runoff, 2) */
  int32_t nInflowTotal;         /* size(InflowGauge%Q, 2) */
  int32_t processCase;          /* processMatrix(8,1): 1, 2 or 3 */
  const int32_t *L1_L11_Id;     /* (nCells1) */
  const int32_t *L11_L1_Id;     /* (nNodes) */
  const int32_t *netPerm;       /* (nNodes) valid 1..nLinks */
  const int32_t *fromN;         /* (nNodes) valid 1..nLinks */
  const int32_t *toN;           /* (nNodes) valid 1..nLinks */
  const double *L1_areaCell;    /* level1%CellArea * 1e-6 [km2] (nCells1) */
  const double *L11_areaCell;   /* level11%CellArea * 1e-6 [km2] (nNodes) */
  const int32_t *gaugeIndexList;        /* (nGauges) */
  const int32_t *gaugeNodeList;         /* (nGauges) */
  const int32_t *InflowGaugeIndexList;  /* (nInflowGauges) */
  const int32_t *InflowGaugeHeadwater;  /* (nInflowGauges) 0/1 */
  const int32_t *InflowGaugeNodeList;   /* (nInflowGauges) */
  /* sub-catchment sharding of one domain over several GPUs, zero / null for an
   * unsharded domain.  Ghost sources are local nodes that stand for the from-node of a cut
   * link owned by another shard: their link is in the local link list (so that the inflows of
   * its to-node are summed in the reference's netPerm order) but its routed outflow series is
   * received (mrm_cuda_import_outflow) instead of computed.  Exports are owned from-nodes of
   * cut links: their outflow series is kept for mrm_cuda_export_outflow.  ssMax > 0 replaces
   * the local maxval(L11_slope) of reg_rout (mRM/mo_mrm_mpr.f90:97) by the whole domain's. */
  int32_t nGhostSources;
  int32_t nExports;
  const int32_t *ghostSourceNodeList;   /* (nGhostSources) local node ids, 1-based */
  const int32_t *exportNodeList;        /* (nExports) local node ids, 1-based */
  double ssMax;
  /* the reference adds the own runoff of ONE sink only, the to-node of the last link in netPerm
   * (mo_mrm_routing.f90:466-467): 0 = derive it from the local netPerm (unsharded), > 0 = local
   * id of the whole domain's last sink if this shard owns it, < 0 = another shard owns it */
  int32_t lastSinkNode;
} mrm_network;
int mrm_cuda_set_network(mhm_cuda_context *ctx, int32_t iDomain, const mrm_network *net);

/* ---- sub-catchment sharding -------------------------------------------------------------
 * Host helper: cut the river forest into sub-catchments (whole subtrees hanging off one link)
 * of at least total/(8 nParts) nodes, pack them onto nParts shards by weight, and give the
 * remaining trunk (every node with a cut link above it, all outlets) to shard 0.  A cut link's
 * from-node is owned by an upstream shard, its to-node by shard 0, so the exchange has one
 * level: shards 1..nParts-1 route, send the outflow series of their cut links for the whole
 * time block, shard 0 routes.  part_of_node[nNodes] receives 0-based shard ids. */
int mrm_partition_subcatchments(int32_t nNodes, int32_t nLinks, const int32_t *fromN,
                                const int32_t *toN, const int32_t *netPerm, int32_t nParts,
                                int32_t *part_of_node);
/* With deferred routing mhm_cuda_run_steps only runs the cells of one time block (n_steps must
 * fit one block) and mrm_cuda_route_pending routes it later -- after the ghost outflows of the
 * block have arrived. */
int mrm_cuda_set_deferred(mhm_cuda_context *ctx, int32_t iDomain, int32_t deferred);
int mrm_cuda_route_pending(mhm_cuda_context *ctx, int32_t iDomain);
/* The same protocol with the exchange BELOW the C ABI (needs mhm_cuda_comm_init): the cut-link
 * series travel by NCCL send / recv on the library's stream, no host synchronisation.
 * mrm_cuda_set_exchange gives the plan -- send_links[r] / recv_links[r] = cut links whose series this
 * shard sends to / receives from rank r; exportNodeList / ghostSourceNodeList of mrm_network must be
 * grouped by peer rank in ascending rank order (and agree link by link with the peer's list).
 * mrm_cuda_shard_run_steps advances one time block (n_steps must fit one block): a shard that only
 * sends runs its cells, routes, sends; a shard that receives first receives and routes the block
 * left pending by its previous call, sends on what it exports, and then runs this block's cells --
 * it works one block behind its senders, so nobody waits for anybody's routing.  mrm_cuda_shard_flush
 * routes the last pending block (call it once per level of receiving shards after the last block). */
int mrm_cuda_set_exchange(mhm_cuda_context *ctx, int32_t iDomain, const int32_t *send_links,
                          const int32_t *recv_links);
int mrm_cuda_shard_run_steps(mhm_cuda_context *ctx, int32_t iDomain, int32_t tt_first, int32_t n_steps);
int mrm_cuda_shard_flush(mhm_cuda_context *ctx, int32_t iDomain);
/* routed outflow of the export nodes over the last routed block / of the ghost sources over the
 * pending block, DEVICE buffers laid out [member][node in list order][n_steps]; both are
 * enqueued on the library's stream (mhm_cuda_synchronize before handing the buffer to NCCL) */
int mrm_cuda_export_outflow(mhm_cuda_context *ctx, int32_t iDomain, double *dev_out, int32_t n_steps);
int mrm_cuda_import_outflow(mhm_cuda_context *ctx, int32_t iDomain, const double *dev_in, int32_t n_steps);

/* ---------------------------------------------------------------------------------
 * N2  river-network initialisation in linear time (host helper, no device needed).
 * Replaces, for the arrays the routing consumes, L11_flow_direction, L11_set_network_topology,
 * L11_routing_order, L11_link_location, L11_set_drain_outlet_gauges and the length / slope part
 * of L11_stream_features (mRM/mo_mrm_net_startup.f90:227-1477), called from mrm_init
 * (mRM/mo_mrm_init.f90:221-231).  2-D arrays are Fortran (nrows, ncols) (first index west ->
 * east), flow directions the rotated in-memory codes (mo_mrm_read_data.f90:527-600), ids and
 * coordinates 1-based; outputs sized nNodes are valid 1..nLinks like the reference's.
 * --------------------------------------------------------------------------------- */
typedef struct mrm_net_inputs {
  int32_t nrows0;                /* level0%nrows */
  int32_t ncols0;                /* level0%ncols */
  int32_t nrows11;               /* level11%nrows */
  int32_t ncols11;               /* level11%ncols */
  int32_t nNodes;                /* level11%nCells */
  int32_t nGauges;               /* domain_mrm%nGauges */
  int32_t coord_sys;             /* iFlag_cordinate_sys */
  int32_t outlet_capacity;       /* size of L0_rowOutlet / L0_colOutlet */
  double cellsize0;              /* level0%cellsize */
  double xllcorner0;
  double yllcorner0;
  const int32_t *mask0;          /* level0%mask 0/1 (nrows0, ncols0) */
  const int32_t *mask11;         /* level11%mask 0/1 (nrows11, ncols11) */
  const int32_t *fDir0;          /* L0_fDir packed (nCells0) */
  const int32_t *fAcc0;          /* L0_fAcc packed */
  const double *elev0;           /* L0_elev packed (null: no length / slope) */
  const int32_t *gaugeLoc0;      /* L0_gaugeLoc packed (null: no gauges) */
  const int32_t *gaugeIdList;    /* domain_mrm%gaugeIdList (nGauges) */
  const int32_t *upper_bound;    /* l0_l11_remap%upper_bound (nNodes) ... */
  const int32_t *lower_bound;
  const int32_t *left_bound;
  const int32_t *right_bound;
  const int32_t *lowres_id_on_highres; /* l0_l11_remap%lowres_id_on_highres (nrows0, ncols0) */
  /* flood plains (L11_stream_features, L11_fraction_sealed_floodplain); all optional */
  const double *cellArea0;       /* level0%CellArea packed (null: cellsize0^2) */
  const int32_t *LCover0;        /* L0_LCover (nCells0, nLCoverScene) */
  int32_t nLCoverScene;
  int32_t LCClassImp;            /* 2 in mrm_init (mo_mrm_init.f90:258) */
  int32_t routingCase;           /* processMatrix(8, 1): 2 / 3 raise the link lengths to their 40th
                                    percentile (L11_stream_features :1440-1446); 0 / 1: as computed */
} mrm_net_inputs;
typedef struct mrm_net_outputs {
  int32_t nCells0;
  int32_t nLinks;
  int32_t nOutlets11;            /* L11_nOutlets */
  int32_t L0_nOutlets;           /* domain_mrm%L0_Noutlet */
  int32_t *fDir11;               /* L11_fDir (nNodes) */
  int32_t *rowOut;               /* L11_rowOut (nNodes) */
  int32_t *colOut;               /* L11_colOut */
  int32_t *fromN;                /* L11_fromN (nNodes) */
  int32_t *toN;
  int32_t *rOrder;
  int32_t *netPerm;
  int32_t *fRow;                 /* L11_fRow (nNodes) */
  int32_t *fCol;
  int32_t *tRow;
  int32_t *tCol;
  int32_t *gaugeNodeList;        /* (nGauges) */
  int32_t *draSC0;               /* L0_draSC packed (nCells0) or null */
  int32_t *draCell0;             /* L0_draCell packed (nCells0) or null */
  int32_t *L0_rowOutlet;         /* (outlet_capacity) or null */
  int32_t *L0_colOutlet;
  double *length;                /* L11_length (nNodes) */
  double *slope;                 /* L11_slope */
  double *aFloodPlain;           /* L11_aFloodPlain (nNodes) or null: no flood plains */
  double *nLinkFracFPimp;        /* L11_nLinkFracFPimp (nNodes, nLCoverScene) or null */
  int32_t *floodPlain0;          /* L0_floodPlain packed (nCells0) or null */
  int32_t *streamNet0;           /* L0_streamNet packed (nCells0) or null (needs the flood plains) */
} mrm_net_outputs;
int mrm_net_init(const mrm_net_inputs *in, mrm_net_outputs *out);
/* L11_flow_accumulation (mRM/mo_mrm_net_startup.f90:2022-2163, called from mrm_init
 * mo_mrm_init.f90:224): L11_fAcc [km2] from L11_fDir and level11%cellarea [m2], both packed
 * (nNodes); mask11 is Fortran (nrows11, ncols11) 0/1.  Linear time, no recursion. */
int mrm_net_flow_accumulation(int32_t nrows11, int32_t ncols11, const int32_t *mask11,
                              const int32_t *fDir11, const double *cellarea11, double *fAcc11);
/* L11_calc_celerity (mRM/mo_mrm_net_startup.f90:2212-2423, called from mrm_update_param
 * mo_mrm_mpr.f90:302 for processCase(8) = 3): L0_slope [%] and L0_streamNet packed (nCells0),
 * the link locations of mrm_net_init, param(1) = slope_factor; fills L11_celerity (nNodes, valid
 * for the links) and, if not null, L0_celerity (nCells0). */
int mrm_net_calc_celerity(int32_t nrows0, int32_t ncols0, const int32_t *mask0, const int32_t *fDir0,
                          const int32_t *streamNet0, const double *slope0, int32_t nNodes,
                          int32_t nLinks, const int32_t *netPerm, const int32_t *fRow,
                          const int32_t *fCol, const int32_t *tRow, const int32_t *tCol,
                          double slope_factor, double *celerity11, double *celerity0);
/* mrm_update_param (mRM/mo_mrm_mpr.f90:241-329): travel time K = L11_length / celerity
 * (processCase(8) = 2: celerity_stride 0, one constant; 3: stride 1, L11_celerity), routing step
 * L11_TSrout from given_TS (mo_mrm_constants.F90:42-46), Muskingum C1 / C2 (nNodes) for
 * mrm_cuda_set_c1c2. */
int mrm_net_update_param(int32_t nNodes, int32_t nOutlets, const double *L11_length,
                         const double *celerity, int32_t celerity_stride, double *C1, double *C2,
                         double *TSrout);
/* L11_L1_mapping (mRM/mo_mrm_net_startup.f90:61-166): masks are Fortran (nrows, ncols) 0/1 */
int mrm_net_l1_l11_mapping(int32_t nrows1, int32_t ncols1, const int32_t *mask1, double cellsize1,
                           int32_t nrows11, int32_t ncols11, const int32_t *mask11, double cellsize11,
                           int32_t *L1_L11_Id, int32_t *L11_L1_Id);

/* L11_routing_order (mRM/mo_mrm_net_startup.f90:728-859) in O(nLinks): host helper that
 * yields the identical rOrder/netPerm as the reference's O(nLinks^2) sweeps */
int mrm_routing_order(int32_t nNodes, int32_t nLinks, const int32_t *fromN,
                      const int32_t *toN, int32_t *rOrder, int32_t *netPerm);

/* Muskingum parameters of one member.
 * case 1: reg_rout (mo_mrm_mpr.f90:61-119) is evaluated on the device whenever the land
 *         cover scene changes: pass the 5 routing gammas, L11_length(s11:e11-1),
 *         L11_slope(s11:e11-1) and L11_nLinkFracFPimp(s11:e11, 1:nLC).
 * case 2/3: pass C1/C2 as left by mrm_update_param (:241-329) and L11_TSrout. */
int mrm_cuda_set_reg_rout(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                          const double *param5, const double *L11_length,
                          const double *L11_slope, const double *L11_nLinkFracFPimp);
int mrm_cuda_set_c1c2(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                      const double *C1, const double *C2, double L11_TSrout);

enum mrm_state_id { MRM_S_QOUT = 0, MRM_S_QTIN, MRM_S_QTR, MRM_S_QMOD, MRM_S_C1, MRM_S_C2,
                    MRM_S_COUNT };
/* L11_qOUT, L11_qMod, L11_C1, L11_C2: (nNodes); L11_qTIN, L11_qTR: (nNodes, 2) */
int mrm_cuda_set_state(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                       int32_t state_id, const double *base, int64_t ld, int64_t offset);
int mrm_cuda_get_state(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                       int32_t state_id, double *base, int64_t ld, int64_t offset);
/* InflowGauge%Q (nDays, nInflowTotal) */
int mrm_cuda_set_inflow(mhm_cuda_context *ctx, int32_t iDomain, const double *Q,
                        int64_t nDays);

/* one call of mRM_routing for step tt (per-step seam, mo_mhm_interface_run.f90:552-590).
 * RunToRout: host (nCells1) or NULL to take the device's L1_total_runoff of the last
 * cell step; InflowDischarge: host (nInflowTotal) or NULL; yId selects nLinkFracFPimp. */
int mrm_cuda_route(mhm_cuda_context *ctx, int32_t iDomain, int32_t member, int32_t tt,
                   int32_t yId, const double *RunToRout, int32_t timestep_rout,
                   double tsRoutFactorIn, const double *InflowDischarge);
/* gauge series mRM_runoff(tt_first : tt_first+n_steps-1, :) of one member into the
 * Fortran (nTimeSteps, nGaugesTotal) array `out` with leading dimension ld */
int mrm_cuda_get_runoff(mhm_cuda_context *ctx, int32_t iDomain, int32_t member,
                        double *out, int64_t ld, int32_t tt_first, int32_t n_steps);

/* ---------------------------------------------------------------------------------
 * B4  MPR upscaling operators (MPR/mo_upscaling_operators.f90:152-227, 266-329, 369-432).
 * L0 data is the packed vector (nL0_cells); the remap is given by the L1 cell bounds
 * in the 2-D L0 grid (common/mo_grid.f90:97-175): upper/lower = first-index (row) range,
 * left/right = second-index (column) range, 1-based inclusive; mask0/cellId0 describe
 * the unpacking (mask0 is the nrows0 x ncols0 logical mask, Fortran order, as int32 0/1).
 * --------------------------------------------------------------------------------- */
typedef struct mpr_l0_grid mpr_l0_grid; /* opaque: device copy of mask/ids/bounds */
int mpr_cuda_grid_create(mhm_cuda_context *ctx, int32_t nrows0, int32_t ncols0,
                         const int32_t *mask0, int32_t nL1_cells,
                         const int32_t *upper_bound, const int32_t *lower_bound,
                         const int32_t *left_bound, const int32_t *right_bound,
                         const int32_t *n_subcells, mpr_l0_grid **grid);
int mpr_cuda_grid_destroy(mhm_cuda_context *ctx, mpr_l0_grid *grid);
int mpr_cuda_upscale_arithmetic_mean(mhm_cuda_context *ctx, const mpr_l0_grid *grid,
                                     double nodata, const double *L0_data, double *L1_out);
int mpr_cuda_upscale_harmonic_mean(mhm_cuda_context *ctx, const mpr_l0_grid *grid,
                                   double nodata, const double *L0_data, double *L1_out);
int mpr_cuda_upscale_geometric_mean(mhm_cuda_context *ctx, const mpr_l0_grid *grid,
                                    double nodata, const double *L0_data, double *L1_out);
int mpr_cuda_l0_fractional_cover(mhm_cuda_context *ctx, const mpr_l0_grid *grid,
                                 const int32_t *dataIn0, int32_t class_id, double *L1_out);

/* ---------------------------------------------------------------------------------
 * B4  MPR: gamma (global parameters) -> L1 effective parameters of one member on the device.
 * Replaces `call mpr(...)` (MPR/mo_mpr_eval.f90:133-152 -> mo_multi_param_reg.f90:67-654)
 * for iFlag_soilDB = 0 and processCase(10) = 0 (no neutrons).  The soil-class table of
 * mpr_sm (mo_mpr_soilmoist.f90:222-324) is evaluated on the host (a few thousand entries),
 * everything per L0 cell and every upscaling on the device.  Results land in the domain's
 * device parameter arrays (no host round trip); mhm_cuda_get_param copies them into the
 * L1_* module globals for restart / pybind coherence.
 * --------------------------------------------------------------------------------- */
typedef struct mpr_l0_inputs {
  int32_t nrows0;                /* level0%nrows (first index) */
  int32_t ncols0;                /* level0%ncols */
  const int32_t *mask0;          /* level0%mask as int32 0/1, Fortran (nrows0, ncols0) */
  const int32_t *upper_bound;    /* l0_l1_remap%upper_bound (nCells1) */
  const int32_t *lower_bound;
  const int32_t *left_bound;
  const int32_t *right_bound;
  const int32_t *n_subcells;
  const int32_t *geoUnit0;       /* L0_geoUnit (nL0) */
  const int32_t *soilId0;        /* L0_soilId(:, 1) (nL0) */
  const int32_t *LCover0;        /* L0_LCover (nL0, nLCscenes) */
  const double *Asp0;            /* L0_asp */
  const double *slope_emp0;      /* L0_slope_emp */
  const double *y0;              /* level0%y packed: latitude */
  const double *gridded_LAI0;    /* L0_gridded_LAI (nL0, nLAI) */
  /* A domain sharded by L1 cells (every shard holds the L0 cells under its own L1 cells): L0_soilId of
   * the WHOLE domain's last L0 cell.  The reference's root fractions of the last horizon use the root
   * zone depth the preceding loop left behind -- that of the last L0 cell's soil type
   * (MPR/mo_mpr_smhorizons.f90:424-427, 453-539) -- so a shard needs this one number to reproduce the
   * unsharded result.  0: this grid's own last cell (unsharded). */
  int32_t lastSoilId0;
} mpr_l0_inputs;
int mpr_cuda_set_l0(mhm_cuda_context *ctx, int32_t iDomain, const mpr_l0_inputs *in);

typedef struct mpr_soil_db {
  int32_t nSoilTypes;            /* size(soilDB%is_present) */
  int32_t maxHorizons;           /* size(soilDB%sand, 2) */
  int32_t nGeoUnits;             /* size(GeoUnitList) */
  const int32_t *is_present;     /* (nSoilTypes) */
  const int32_t *nHorizons;
  const int32_t *nTillHorizons;
  const double *sand;            /* (nSoilTypes, maxHorizons) */
  const double *clay;
  const double *DbM;
  const double *Wd;              /* (nSoilTypes, nSoilHorizons_mHM, maxHorizons) */
  const double *RZdepth;         /* (nSoilTypes) */
  const double *HorizonDepth_mHM; /* (nSoilHorizons_mHM) */
  const int32_t *GeoUnitList;    /* (nGeoUnits) */
  const int32_t *GeoUnitKar;     /* (nGeoUnits) */
  double fracSealed_CityArea;
} mpr_soil_db;
int mpr_cuda_set_soildb(mhm_cuda_context *ctx, int32_t iDomain, const mpr_soil_db *db);

/* param = the flat global parameter vector (global_parameters(:, 3) or the optimiser's
 * candidate), sliced inside by processMatrix exactly like the reference */
int mpr_cuda_eval(mhm_cuda_context *ctx, int32_t iDomain, int32_t member, const double *param,
                  int32_t nParam);
int mhm_cuda_get_param(mhm_cuda_context *ctx, int32_t iDomain, int32_t member, int32_t param_id,
                       double *base, int64_t ld, int64_t offset, int32_t dim2, int32_t dim3);

/* init_lowres_level (common/mo_grid.f90:58-183): host helper, integer maps bit-exact.
 * Call once with upper_bound = NULL to get nCells1 (return value written to *nCells1), then
 * with arrays of that size.  mask arrays are Fortran (nrows, ncols) int32 0/1. */
int mhm_grid_init_lowres_level(int32_t nrows0, int32_t ncols0, const int32_t *mask0,
                               const double *cellArea0, double cellsize0, double target_resolution,
                               int32_t *nrows1, int32_t *ncols1, int32_t *nCells1, int32_t *mask1,
                               int32_t *cellCoor, double *cellArea1, int32_t *upper_bound,
                               int32_t *lower_bound, int32_t *left_bound, int32_t *right_bound,
                               int32_t *n_subcells, int32_t *lowres_id_on_highres);

/* ---------------------------------------------------------------------------------
 * measurement hooks (bench.py): CUDA events on the library's own stream
 * --------------------------------------------------------------------------------- */
int mhm_cuda_event_record(mhm_cuda_context *ctx, int32_t slot);            /* slot 0..15 */
int mhm_cuda_event_elapsed_ms(mhm_cuda_context *ctx, int32_t a, int32_t b, double *ms);
int mhm_cuda_synchronize(mhm_cuda_context *ctx);
/* accumulated device time [ms] and launch count of kernel class `which` since the last
 * reset (0 = fused cell kernel, 1 = routing kernels, 2 = upscaling kernels) */
int mhm_cuda_kernel_stats(mhm_cuda_context *ctx, int32_t which, double *ms, int64_t *launches);
int mhm_cuda_kernel_stats_reset(mhm_cuda_context *ctx, int32_t enable_timing);
/* fp64 FMA peak micro-benchmark on this device: returns DFMA/s (x2 = FLOP/s) */
int mhm_cuda_measure_dfma_peak(mhm_cuda_context *ctx, double *dfma_per_s);

#ifdef __cplusplus
}
#endif
#endif /* MHM_CUDA_H */
