!> \file mo_mhm_cuda.F90
!> \brief ISO_C_BINDING interface of libmhm_cuda.so (include/mhm_cuda.h) for the mHM driver.
!> \details This module is what the reference's Fortran driver compiles (with -DMHM_CUDA) to
!! reach the B200 implementation of the L1 hot path.  It holds only `interface ... bind(C)`
!! blocks and thin wrappers that turn a non-zero status into the reference's own fatal path
!! (`error_message`, FORCES mo_message).  It cannot be compiled in the build container of this
!! repository (no Fortran compiler); it is written against the C header symbol by symbol and
!! the same symbols are exercised through ctypes by tests/.  INTEGRATION.md shows the four
!! call sites of the reference that use it.
!!
!! Conventions (see include/mhm_cuda.h): array sections are passed as the base address of the
!! whole module-global array (c_loc of its first element), its leading dimension
!! ld = size(array, 1) and offset = s1 - 1; logicals cross as integer(c_int32_t) 0/1.
module mo_mhm_cuda
  use, intrinsic :: iso_c_binding
  implicit none
  private

  public :: mhm_cuda_ctx, mhm_cuda_check
  public :: mhm_domain_config, mhm_meteo_config, mhm_time_config, mhm_step_index, mrm_network

  !> opaque library context (one per process / GPU), set by mhm_cuda_init
  type(c_ptr), save :: mhm_cuda_ctx = c_null_ptr

  ! enum mhm_param_id
  integer(c_int32_t), parameter, public :: MHM_P_FSEALED = 0, MHM_P_ALPHA = 1, MHM_P_DEGDAYINC = 2, &
    MHM_P_DEGDAYMAX = 3, MHM_P_DEGDAYNOPRE = 4, MHM_P_FROOTS = 5, MHM_P_MAXINTER = 6, MHM_P_KARSTLOSS = 7, &
    MHM_P_KFASTFLOW = 8, MHM_P_KSLOWFLOW = 9, MHM_P_KBASEFLOW = 10, MHM_P_KPERCO = 11, &
    MHM_P_SOILMOISTFC = 12, MHM_P_SOILMOISTSAT = 13, MHM_P_SOILMOISTEXP = 14, MHM_P_JARVIS_C1 = 15, &
    MHM_P_TEMPTHRESH = 16, MHM_P_UNSATTHRESH = 17, MHM_P_SEALEDTHRESH = 18, MHM_P_WILTINGPOINT = 19, &
    MHM_P_PETLAICORFACTOR = 20, MHM_P_FASP = 21, MHM_P_HARSAMCOEFF = 22, MHM_P_PRIETAYALPHA = 23, &
    MHM_P_AERORESIST = 24, MHM_P_SURFRESIST = 25, MHM_P_LATITUDE = 26
  ! enum mhm_state_id
  integer(c_int32_t), parameter, public :: MHM_S_INTER = 0, MHM_S_SNOWPACK = 1, MHM_S_SEALSTW = 2, &
    MHM_S_UNSATSTW = 3, MHM_S_SATSTW = 4, MHM_S_SOILMOIST = 5
  ! enum mhm_flux_id
  integer(c_int32_t), parameter, public :: MHM_F_PET_CALC = 0, MHM_F_TEMP_CALC = 1, MHM_F_PREC_CALC = 2, &
    MHM_F_AETCANOPY = 3, MHM_F_AETSEALED = 4, MHM_F_BASEFLOW = 5, MHM_F_FASTRUNOFF = 6, MHM_F_MELT = 7, &
    MHM_F_PERCOL = 8, MHM_F_PREEFFECT = 9, MHM_F_RAIN = 10, MHM_F_RUNOFFSEAL = 11, MHM_F_SLOWRUNOFF = 12, &
    MHM_F_SNOW = 13, MHM_F_THROUGHFALL = 14, MHM_F_TOTAL_RUNOFF = 15, MHM_F_DEGDAY = 16, &
    MHM_F_AETSOIL = 17, MHM_F_INFILSOIL = 18
  ! enum mhm_meteo_var
  integer(c_int32_t), parameter, public :: MHM_M_PRE = 0, MHM_M_TEMP = 1, MHM_M_PET = 2, MHM_M_TMIN = 3, &
    MHM_M_TMAX = 4, MHM_M_NETRAD = 5, MHM_M_ABSVAPPRESS = 6, MHM_M_WINDSPEED = 7
  ! enum mrm_state_id
  integer(c_int32_t), parameter, public :: MRM_S_QOUT = 0, MRM_S_QTIN = 1, MRM_S_QTR = 2, MRM_S_QMOD = 3, &
    MRM_S_C1 = 4, MRM_S_C2 = 5

  type, bind(C) :: mhm_domain_config
    integer(c_int32_t) :: nCells, nHorizons, nLAI, nLCscenes, nMembers, nProcesses, timestep_h, read_states
    real(c_double) :: c2TSTu
    type(c_ptr) :: processMatrix   !< c_loc(processMatrix(1,1)), int32 (nProcesses, 3)
  end type mhm_domain_config

  type, bind(C) :: mhm_meteo_config
    integer(c_int32_t) :: pet_case, nTstepForcingDay, is_hourly_forcing, read_meteo_weights
    real(c_double), dimension(12) :: fday_prec, fnight_prec, fday_pet, fnight_pet, fday_temp, fnight_temp, &
                                     evap_coeff
  end type mhm_meteo_config

  type, bind(C) :: mhm_time_config
    integer(c_int32_t) :: jul_start, nTimeSteps, warming_days, timeStep_LAI_input, lc_year_start, lc_nyears
    type(c_ptr) :: LCyearId        !< c_loc(LCyearId(lc_year_start, iDomain))
  end type mhm_time_config

  type, bind(C) :: mhm_step_index
    integer(c_int32_t) :: iMeteoTS, year
    integer(c_int16_t) :: yId, iLAI, doy
    integer(c_int8_t) :: month, hour, isday
    integer(c_int8_t) :: pad_(3)
  end type mhm_step_index

  type, bind(C) :: mrm_network
    integer(c_int32_t) :: nNodes, nOutlets, map_flag, nGauges, nInflowGauges, nGaugesTotal, nInflowTotal, &
                          processCase
    type(c_ptr) :: L1_L11_Id, L11_L1_Id, netPerm, fromN, toN, L1_areaCell, L11_areaCell, gaugeIndexList, &
                   gaugeNodeList, InflowGaugeIndexList, InflowGaugeHeadwater, InflowGaugeNodeList
    ! sub-catchment sharding (zero / c_null_ptr for an unsharded domain)
    integer(c_int32_t) :: nGhostSources = 0, nExports = 0
    type(c_ptr) :: ghostSourceNodeList = c_null_ptr, exportNodeList = c_null_ptr
    real(c_double) :: ssMax = 0.0_c_double
    integer(c_int32_t) :: lastSinkNode = 0
  end type mrm_network

  !> L0 inputs of mpr (MPR/mo_multi_param_reg.f90:67-75), see include/mhm_cuda.h
  type, bind(C) :: mpr_l0_inputs
    integer(c_int32_t) :: nrows0, ncols0
    type(c_ptr) :: mask0, upper_bound, lower_bound, left_bound, right_bound, n_subcells, geoUnit0, soilId0, &
                   LCover0, Asp0, slope_emp0, y0, gridded_LAI0
    integer(c_int32_t) :: lastSoilId0   !< sharded domains: L0_soilId of the whole domain's last L0 cell; 0 otherwise
  end type mpr_l0_inputs

  !> soilDB + geological units (MPR/mo_mpr_global_variables.f90)
  type, bind(C) :: mpr_soil_db
    integer(c_int32_t) :: nSoilTypes, maxHorizons, nGeoUnits
    type(c_ptr) :: is_present, nHorizons, nTillHorizons, sand, clay, DbM, Wd, RZdepth, HorizonDepth_mHM, &
                   GeoUnitList, GeoUnitKar
    real(c_double) :: fracSealed_CityArea
  end type mpr_soil_db

  !> calibration aggregates + BFI sums (mhm_cuda_set_optisim)
  type, bind(C) :: mhm_optisim_config
    integer(c_int32_t) :: sm_on = 0, sm_timeStepInput = -1, sm_nTime = 0, nSoilHorizons_sm_input = 1
    integer(c_int32_t) :: et_on = 0, et_timeStepInput = -1, et_nTime = 0
    integer(c_int32_t) :: tws_on = 0, tws_timeStepInput = -1, tws_nTime = 0
    integer(c_int32_t) :: bfi_on = 0
  end type mhm_optisim_config

  interface
    integer(c_int) function mhm_cuda_init(device, ctx) bind(C, name = 'mhm_cuda_init')
      import
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: ctx
    end function
    integer(c_int) function mhm_cuda_finalize(ctx) bind(C, name = 'mhm_cuda_finalize')
      import
      type(c_ptr), value :: ctx
    end function
    type(c_ptr) function mhm_cuda_last_error() bind(C, name = 'mhm_cuda_last_error')
      import
    end function
    integer(c_int) function mhm_cuda_register_domain(ctx, iDomain, cfg) bind(C, name = 'mhm_cuda_register_domain')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      type(mhm_domain_config), intent(in) :: cfg
    end function
    integer(c_int) function mhm_cuda_set_param(ctx, iDomain, member, param_id, base, ld, offset, dim2, dim3) &
        bind(C, name = 'mhm_cuda_set_param')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, member, param_id, dim2, dim3
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_set_state(ctx, iDomain, member, state_id, base, ld, offset) &
        bind(C, name = 'mhm_cuda_set_state')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, member, state_id
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_get_state(ctx, iDomain, member, state_id, base, ld, offset) &
        bind(C, name = 'mhm_cuda_get_state')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, member, state_id
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_get_flux(ctx, iDomain, member, flux_id, base, ld, offset) &
        bind(C, name = 'mhm_cuda_get_flux')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, member, flux_id
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_bind_host_state(ctx, iDomain, state_id, base, ld, offset) &
        bind(C, name = 'mhm_cuda_bind_host_state')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, state_id
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_bind_host_flux(ctx, iDomain, flux_id, base, ld, offset) &
        bind(C, name = 'mhm_cuda_bind_host_flux')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, flux_id
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_sync_to_host(ctx, iDomain) bind(C, name = 'mhm_cuda_sync_to_host')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
    end function
    integer(c_int) function mhm_cuda_set_meteo_config(ctx, iDomain, cfg) bind(C, name = 'mhm_cuda_set_meteo_config')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      type(mhm_meteo_config), intent(in) :: cfg
    end function
    integer(c_int) function mhm_cuda_set_meteo(ctx, iDomain, var, base, ld, offset, first_step, n_steps) &
        bind(C, name = 'mhm_cuda_set_meteo')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, var
      integer(c_int64_t), value :: ld, offset, first_step, n_steps
    end function
    integer(c_int) function mhm_cuda_set_meteo_async(ctx, iDomain, var, base, ld, offset, first_step, n_steps) &
        bind(C, name = 'mhm_cuda_set_meteo_async')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, var
      integer(c_int64_t), value :: ld, offset, first_step, n_steps
    end function
    !> rank 0 creates the 128-byte id; broadcast it (MPI_Bcast of 128 characters), then every rank
    !! calls mhm_cuda_comm_init(ctx, nproc, rank, id)
    integer(c_int) function mhm_cuda_comm_unique_id(id128) bind(C, name = 'mhm_cuda_comm_unique_id')
      import
      character(kind = c_char), dimension(128), intent(out) :: id128
    end function
    integer(c_int) function mhm_cuda_comm_init(ctx, nranks, rank, id128) bind(C, name = 'mhm_cuda_comm_init')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: nranks, rank
      character(kind = c_char), dimension(128), intent(in) :: id128
    end function
    integer(c_int) function mhm_cuda_comm_finalize(ctx) bind(C, name = 'mhm_cuda_comm_finalize')
      import
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function mhm_cuda_comm_info(ctx, nranks, rank, nccl_version) bind(C, name = 'mhm_cuda_comm_info')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), intent(out) :: nranks, rank, nccl_version
    end function
    integer(c_int) function mhm_cuda_meteo_shared_rows(n_steps, nranks, rank, first_row, n_rows) &
        bind(C, name = 'mhm_cuda_meteo_shared_rows')
      import
      integer(c_int64_t), value :: n_steps
      integer(c_int32_t), value :: nranks, rank
      integer(c_int64_t), intent(out) :: first_row, n_rows
    end function
    !> forcing shared by all ranks: each rank uploads 1/nranks of the chunk, NCCL all-gather
    integer(c_int) function mhm_cuda_set_meteo_shared(ctx, iDomain, var, base, is_f32, ld, offset, first_step, &
        n_steps) bind(C, name = 'mhm_cuda_set_meteo_shared')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, var, is_f32
      integer(c_int64_t), value :: ld, offset, first_step, n_steps
    end function
    integer(c_int) function mhm_cuda_meteo_h2d_bytes(ctx, iDomain, bytes) bind(C, name = 'mhm_cuda_meteo_h2d_bytes')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      integer(c_int64_t), intent(out) :: bytes
    end function
    !> sub-catchment sharding with the NCCL exchange inside the library
    integer(c_int) function mrm_cuda_set_exchange(ctx, iDomain, send_links, recv_links) &
        bind(C, name = 'mrm_cuda_set_exchange')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      integer(c_int32_t), dimension(*), intent(in) :: send_links, recv_links
    end function
    integer(c_int) function mrm_cuda_shard_run_steps(ctx, iDomain, tt_first, n_steps) &
        bind(C, name = 'mrm_cuda_shard_run_steps')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain, tt_first, n_steps
    end function
    integer(c_int) function mrm_cuda_shard_flush(ctx, iDomain) bind(C, name = 'mrm_cuda_shard_flush')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
    end function
    integer(c_int) function mhm_cuda_set_meteo_weights(ctx, iDomain, var, base, ld, offset) &
        bind(C, name = 'mhm_cuda_set_meteo_weights')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, var
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_set_time(ctx, iDomain, cfg) bind(C, name = 'mhm_cuda_set_time')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      type(mhm_time_config), intent(in) :: cfg
    end function
    integer(c_int) function mhm_cuda_cell_step(ctx, iDomain, tt, idx) bind(C, name = 'mhm_cuda_cell_step')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain, tt
      type(mhm_step_index), intent(in) :: idx
    end function
    integer(c_int) function mhm_cuda_run_steps(ctx, iDomain, tt_first, n_steps) bind(C, name = 'mhm_cuda_run_steps')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain, tt_first, n_steps
    end function
    integer(c_int) function mrm_cuda_set_network(ctx, iDomain, net) bind(C, name = 'mrm_cuda_set_network')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      type(mrm_network), intent(in) :: net
    end function
    integer(c_int) function mrm_cuda_set_reg_rout(ctx, iDomain, member, param5, length, slope, fFPimp) &
        bind(C, name = 'mrm_cuda_set_reg_rout')
      import
      type(c_ptr), value :: ctx, param5, length, slope, fFPimp
      integer(c_int32_t), value :: iDomain, member
    end function
    integer(c_int) function mrm_cuda_set_c1c2(ctx, iDomain, member, C1, C2, TSrout) bind(C, name = 'mrm_cuda_set_c1c2')
      import
      type(c_ptr), value :: ctx, C1, C2
      integer(c_int32_t), value :: iDomain, member
      real(c_double), value :: TSrout
    end function
    integer(c_int) function mrm_cuda_set_state(ctx, iDomain, member, state_id, base, ld, offset) &
        bind(C, name = 'mrm_cuda_set_state')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, member, state_id
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mrm_cuda_get_state(ctx, iDomain, member, state_id, base, ld, offset) &
        bind(C, name = 'mrm_cuda_get_state')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, member, state_id
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mrm_cuda_set_inflow(ctx, iDomain, Q, nDays) bind(C, name = 'mrm_cuda_set_inflow')
      import
      type(c_ptr), value :: ctx, Q
      integer(c_int32_t), value :: iDomain
      integer(c_int64_t), value :: nDays
    end function
    integer(c_int) function mrm_cuda_route(ctx, iDomain, member, tt, yId, RunToRout, timestep_rout, &
                                           tsRoutFactorIn, InflowDischarge) bind(C, name = 'mrm_cuda_route')
      import
      type(c_ptr), value :: ctx, RunToRout, InflowDischarge
      integer(c_int32_t), value :: iDomain, member, tt, yId, timestep_rout
      real(c_double), value :: tsRoutFactorIn
    end function
    integer(c_int) function mrm_cuda_get_runoff(ctx, iDomain, member, out, ld, tt_first, n_steps) &
        bind(C, name = 'mrm_cuda_get_runoff')
      import
      type(c_ptr), value :: ctx, out
      integer(c_int32_t), value :: iDomain, member, tt_first, n_steps
      integer(c_int64_t), value :: ld
    end function
    integer(c_int) function mpr_cuda_grid_create(ctx, nrows0, ncols0, mask0, nL1, upper, lower, left, right, &
                                                 n_subcells, grid) bind(C, name = 'mpr_cuda_grid_create')
      import
      type(c_ptr), value :: ctx, mask0, upper, lower, left, right, n_subcells
      integer(c_int32_t), value :: nrows0, ncols0, nL1
      type(c_ptr), intent(out) :: grid
    end function
    integer(c_int) function mpr_cuda_upscale_arithmetic_mean(ctx, grid, nodata, L0_data, L1_out) &
        bind(C, name = 'mpr_cuda_upscale_arithmetic_mean')
      import
      type(c_ptr), value :: ctx, grid, L0_data, L1_out
      real(c_double), value :: nodata
    end function
    integer(c_int) function mpr_cuda_upscale_harmonic_mean(ctx, grid, nodata, L0_data, L1_out) &
        bind(C, name = 'mpr_cuda_upscale_harmonic_mean')
      import
      type(c_ptr), value :: ctx, grid, L0_data, L1_out
      real(c_double), value :: nodata
    end function
    integer(c_int) function mpr_cuda_l0_fractional_cover(ctx, grid, dataIn0, class_id, L1_out) &
        bind(C, name = 'mpr_cuda_l0_fractional_cover')
      import
      type(c_ptr), value :: ctx, grid, dataIn0, L1_out
      integer(c_int32_t), value :: class_id
    end function
    ! ---- B4: the whole of mpr_eval on the device -------------------------------------------
    integer(c_int) function mpr_cuda_set_l0(ctx, iDomain, l0) bind(C, name = 'mpr_cuda_set_l0')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      type(mpr_l0_inputs), intent(in) :: l0
    end function
    integer(c_int) function mpr_cuda_set_soildb(ctx, iDomain, db) bind(C, name = 'mpr_cuda_set_soildb')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      type(mpr_soil_db), intent(in) :: db
    end function
    integer(c_int) function mpr_cuda_eval(ctx, iDomain, member, param, nParam) bind(C, name = 'mpr_cuda_eval')
      import
      type(c_ptr), value :: ctx, param
      integer(c_int32_t), value :: iDomain, member, nParam
    end function
    integer(c_int) function mhm_cuda_get_param(ctx, iDomain, member, param_id, base, ld, offset, dim2, dim3) &
        bind(C, name = 'mhm_cuda_get_param')
      import
      type(c_ptr), value :: ctx, base
      integer(c_int32_t), value :: iDomain, member, param_id, dim2, dim3
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_states_default_init(ctx, iDomain, HorizonDepth_mHM) &
        bind(C, name = 'mhm_cuda_states_default_init')
      import
      type(c_ptr), value :: ctx, HorizonDepth_mHM
      integer(c_int32_t), value :: iDomain
    end function
    ! ---- N3: meteo chunk on the level-2 grid, remapped and packed on the device --------------
    integer(c_int) function mhm_cuda_set_meteo_l2(ctx, iDomain, var, data2, is_f32, nrows2, ncols2, mask2, cellsize2, &
        nrows1, ncols1, mask1, cellsize1, first_step, n_steps) bind(C, name = 'mhm_cuda_set_meteo_l2')
      import
      type(c_ptr), value :: ctx, data2, mask2, mask1
      integer(c_int32_t), value :: iDomain, var, is_f32, nrows2, ncols2, nrows1, ncols1
      real(c_double), value :: cellsize2, cellsize1
      integer(c_int64_t), value :: first_step, n_steps
    end function
    ! ---- N2: river-network initialisation (host, linear time); mrm_net_inputs / mrm_net_outputs
    !      are plain structs of sizes and c_ptr in the order of include/mhm_cuda.h -------------
    integer(c_int) function mrm_net_init(net_in, net_out) bind(C, name = 'mrm_net_init')
      import
      type(c_ptr), value :: net_in, net_out
    end function
    integer(c_int) function mrm_net_l1_l11_mapping(nrows1, ncols1, mask1, cellsize1, nrows11, ncols11, mask11, &
        cellsize11, L1_L11_Id, L11_L1_Id) bind(C, name = 'mrm_net_l1_l11_mapping')
      import
      integer(c_int32_t), value :: nrows1, ncols1, nrows11, ncols11
      real(c_double), value :: cellsize1, cellsize11
      type(c_ptr), value :: mask1, mask11, L1_L11_Id, L11_L1_Id
    end function
    ! L11_flow_accumulation (mo_mrm_net_startup.f90:2022), L11_calc_celerity (:2212) and
    ! mrm_update_param (mo_mrm_mpr.f90:241) for the celerity-based routing cases 2 and 3
    integer(c_int) function mrm_net_flow_accumulation(nrows11, ncols11, mask11, fDir11, cellarea11, fAcc11) &
        bind(C, name = 'mrm_net_flow_accumulation')
      import
      integer(c_int32_t), value :: nrows11, ncols11
      type(c_ptr), value :: mask11, fDir11, cellarea11, fAcc11
    end function
    integer(c_int) function mrm_net_calc_celerity(nrows0, ncols0, mask0, fDir0, streamNet0, slope0, nNodes, nLinks, &
        netPerm, fRow, fCol, tRow, tCol, slope_factor, celerity11, celerity0) bind(C, name = 'mrm_net_calc_celerity')
      import
      integer(c_int32_t), value :: nrows0, ncols0, nNodes, nLinks
      type(c_ptr), value :: mask0, fDir0, streamNet0, slope0, netPerm, fRow, fCol, tRow, tCol, celerity11, celerity0
      real(c_double), value :: slope_factor
    end function
    integer(c_int) function mrm_net_update_param(nNodes, nOutlets, L11_length, celerity, celerity_stride, C1, C2, &
        TSrout) bind(C, name = 'mrm_net_update_param')
      import
      integer(c_int32_t), value :: nNodes, nOutlets, celerity_stride
      type(c_ptr), value :: L11_length, celerity, C1, C2
      real(c_double), intent(out) :: TSrout
    end function
    ! ---- A10: gridded outputs accumulated on the device ------------------------------------
    integer(c_int) function mhm_cuda_set_outputs(ctx, iDomain, outputFlxState, timeStep_model_outputs) &
        bind(C, name = 'mhm_cuda_set_outputs')
      import
      type(c_ptr), value :: ctx, outputFlxState        !< int32 (21), 0/1
      integer(c_int32_t), value :: iDomain, timeStep_model_outputs
    end function
    integer(c_int) function mhm_cuda_get_output_windows(ctx, iDomain, n_windows, tt_end, capacity) &
        bind(C, name = 'mhm_cuda_get_output_windows')
      import
      type(c_ptr), value :: ctx, tt_end
      integer(c_int32_t), value :: iDomain, capacity
      integer(c_int32_t), intent(out) :: n_windows
    end function
    integer(c_int) function mhm_cuda_get_output(ctx, iDomain, member, window, variable, horizon, out) &
        bind(C, name = 'mhm_cuda_get_output')
      import
      type(c_ptr), value :: ctx, out
      integer(c_int32_t), value :: iDomain, member, window, variable, horizon
    end function
    ! ---- A10: calibration aggregates (update_optisim) and BFI sums -------------------------
    integer(c_int) function mhm_cuda_set_optisim(ctx, iDomain, cfg) bind(C, name = 'mhm_cuda_set_optisim')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
      type(mhm_optisim_config), intent(in) :: cfg
    end function
    integer(c_int) function mhm_cuda_get_optisim(ctx, iDomain, member, which, base, ld, offset) &
        bind(C, name = 'mhm_cuda_get_optisim')
      import
      type(c_ptr), value :: ctx, base                  !< c_loc(xxOptiSim(iDomain)%dataSim)
      integer(c_int32_t), value :: iDomain, member, which
      integer(c_int64_t), value :: ld, offset
    end function
    integer(c_int) function mhm_cuda_get_bfi_sums(ctx, iDomain, member, cellArea, qBF_sum, qT_sum) &
        bind(C, name = 'mhm_cuda_get_bfi_sums')
      import
      type(c_ptr), value :: ctx, cellArea              !< c_loc(level1(iDomain)%CellArea)
      integer(c_int32_t), value :: iDomain, member
      real(c_double), intent(out) :: qBF_sum, qT_sum
    end function
    ! ---- sub-catchment sharding (one domain over several GPUs / MPI ranks) -----------------
    integer(c_int) function mrm_partition_subcatchments(nNodes, nLinks, fromN, toN, netPerm, nParts, part_of_node) &
        bind(C, name = 'mrm_partition_subcatchments')
      import
      integer(c_int32_t), value :: nNodes, nLinks, nParts
      type(c_ptr), value :: fromN, toN, netPerm, part_of_node
    end function
    integer(c_int) function mrm_cuda_set_deferred(ctx, iDomain, deferred) bind(C, name = 'mrm_cuda_set_deferred')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain, deferred
    end function
    integer(c_int) function mrm_cuda_route_pending(ctx, iDomain) bind(C, name = 'mrm_cuda_route_pending')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
    end function
    integer(c_int) function mrm_cuda_export_outflow(ctx, iDomain, dev_out, n_steps) &
        bind(C, name = 'mrm_cuda_export_outflow')
      import
      type(c_ptr), value :: ctx, dev_out               !< DEVICE buffer (member, export, step)
      integer(c_int32_t), value :: iDomain, n_steps
    end function
    integer(c_int) function mrm_cuda_import_outflow(ctx, iDomain, dev_in, n_steps) &
        bind(C, name = 'mrm_cuda_import_outflow')
      import
      type(c_ptr), value :: ctx, dev_in
      integer(c_int32_t), value :: iDomain, n_steps
    end function
    integer(c_int) function mrm_routing_order(nNodes, nLinks, fromN, toN, rOrder, netPerm) &
        bind(C, name = 'mrm_routing_order')
      import
      integer(c_int32_t), value :: nNodes, nLinks
      type(c_ptr), value :: fromN, toN, rOrder, netPerm
    end function
    ! ---- entry points without a call site of their own in the patched driver: teardown, math mode,
    ! ---- zero-copy forcing, diagnostics, host-side integer maps and the calendar
    integer(c_int) function mhm_cuda_unregister_domain(ctx, iDomain) bind(C, name = 'mhm_cuda_unregister_domain')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain
    end function
    type(c_ptr) function mhm_cuda_version() bind(C, name = 'mhm_cuda_version')
      import
    end function
    !> mode 0 = strict (reference operation order, libdevice), 1 = fast (tables, FMA contraction)
    integer(c_int) function mhm_cuda_set_math_mode(ctx, mode) bind(C, name = 'mhm_cuda_set_math_mode')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: mode
    end function
    integer(c_int) function mhm_cuda_synchronize(ctx) bind(C, name = 'mhm_cuda_synchronize')
      import
      type(c_ptr), value :: ctx
    end function
    !> dev: device address of a [n_steps][nCells] chunk owned by the caller (no copy)
    integer(c_int) function mhm_cuda_set_meteo_device(ctx, iDomain, var, dev, first_step, n_steps) &
        bind(C, name = 'mhm_cuda_set_meteo_device')
      import
      type(c_ptr), value :: ctx, dev
      integer(c_int32_t), value :: iDomain, var
      integer(c_int64_t), value :: first_step, n_steps
    end function
    integer(c_int) function mhm_cuda_get_meteo(ctx, iDomain, var, out, ld, first_step, n_steps) &
        bind(C, name = 'mhm_cuda_get_meteo')
      import
      type(c_ptr), value :: ctx, out
      integer(c_int32_t), value :: iDomain, var
      integer(c_int64_t), value :: ld, first_step, n_steps
    end function
    integer(c_int) function mhm_cuda_keep_runoff_history(ctx, iDomain, keep) &
        bind(C, name = 'mhm_cuda_keep_runoff_history')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: iDomain, keep
    end function
    integer(c_int) function mhm_cuda_get_runoff_history(ctx, iDomain, member, out, ld) &
        bind(C, name = 'mhm_cuda_get_runoff_history')
      import
      type(c_ptr), value :: ctx, out
      integer(c_int32_t), value :: iDomain, member
      integer(c_int64_t), value :: ld
    end function
    integer(c_int) function mhm_cuda_event_record(ctx, slot) bind(C, name = 'mhm_cuda_event_record')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: slot
    end function
    integer(c_int) function mhm_cuda_event_elapsed_ms(ctx, a, b, ms) bind(C, name = 'mhm_cuda_event_elapsed_ms')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: a, b
      real(c_double), intent(out) :: ms
    end function
    integer(c_int) function mhm_cuda_kernel_stats(ctx, which, ms, launches) bind(C, name = 'mhm_cuda_kernel_stats')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: which
      real(c_double), intent(out) :: ms
      integer(c_int64_t), intent(out) :: launches
    end function
    integer(c_int) function mhm_cuda_kernel_stats_reset(ctx, enable_timing) &
        bind(C, name = 'mhm_cuda_kernel_stats_reset')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: enable_timing
    end function
    integer(c_int) function mhm_cuda_measure_dfma_peak(ctx, dfma_per_s) bind(C, name = 'mhm_cuda_measure_dfma_peak')
      import
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: dfma_per_s
    end function
    integer(c_int) function mpr_cuda_grid_destroy(ctx, grid) bind(C, name = 'mpr_cuda_grid_destroy')
      import
      type(c_ptr), value :: ctx, grid
    end function
    integer(c_int) function mpr_cuda_upscale_geometric_mean(ctx, grid, nodata, L0_data, L1_out) &
        bind(C, name = 'mpr_cuda_upscale_geometric_mean')
      import
      type(c_ptr), value :: ctx, grid, L0_data, L1_out
      real(c_double), value :: nodata
    end function
    !> init_lowres_level (mo_grid.f90) without a device: masks and bounds as c_loc of the caller's arrays
    integer(c_int) function mhm_grid_init_lowres_level(nrows0, ncols0, mask0, cellArea0, cellsize0, &
        target_resolution, nrows1, ncols1, nCells1, mask1, cellCoor, cellArea1, upper_bound, lower_bound, &
        left_bound, right_bound, n_subcells, lowres_id_on_highres) bind(C, name = 'mhm_grid_init_lowres_level')
      import
      integer(c_int32_t), value :: nrows0, ncols0
      type(c_ptr), value :: mask0, cellArea0
      real(c_double), value :: cellsize0, target_resolution
      integer(c_int32_t), intent(out) :: nrows1, ncols1, nCells1
      type(c_ptr), value :: mask1, cellCoor, cellArea1, upper_bound, lower_bound, left_bound, right_bound, &
                            n_subcells, lowres_id_on_highres
    end function
    !> the per-step calendar of mhm_interface_run_do_time_step for steps tt_first .. tt_first+n_steps-1
    integer(c_int) function mhm_time_indices(cfg, timestep_h, nTstepForcingDay, tt_first, n_steps, out) &
        bind(C, name = 'mhm_time_indices')
      import
      type(mhm_time_config), intent(in) :: cfg
      integer(c_int32_t), value :: timestep_h, nTstepForcingDay, tt_first, n_steps
      type(mhm_step_index), intent(out) :: out(*)
    end function
  end interface

  public :: mhm_cuda_init, mhm_cuda_finalize, mhm_cuda_register_domain, mhm_cuda_set_param, &
            mhm_cuda_set_state, mhm_cuda_get_state, mhm_cuda_get_flux, mhm_cuda_bind_host_state, &
            mhm_cuda_bind_host_flux, mhm_cuda_sync_to_host, mhm_cuda_set_meteo_config, mhm_cuda_set_meteo, &
            mhm_cuda_set_meteo_async, mhm_cuda_set_meteo_weights, mhm_cuda_set_time, mhm_cuda_cell_step, mhm_cuda_run_steps, &
            mrm_cuda_set_network, mrm_cuda_set_reg_rout, mrm_cuda_set_c1c2, mrm_cuda_set_state, &
            mrm_cuda_get_state, mrm_cuda_set_inflow, mrm_cuda_route, mrm_cuda_get_runoff, &
            mpr_cuda_grid_create, mpr_cuda_upscale_arithmetic_mean, mpr_cuda_upscale_harmonic_mean, &
            mpr_cuda_l0_fractional_cover, mpr_cuda_set_l0, mpr_cuda_set_soildb, mpr_cuda_eval, &
            mhm_cuda_get_param, mhm_cuda_states_default_init, mhm_cuda_set_outputs, &
            mhm_cuda_get_output_windows, mhm_cuda_get_output, mrm_partition_subcatchments, &
            mhm_cuda_set_optisim, mhm_cuda_get_optisim, mhm_cuda_get_bfi_sums, &
            mrm_cuda_set_deferred, mrm_cuda_route_pending, mrm_cuda_export_outflow, mrm_cuda_import_outflow, &
            mrm_routing_order, mhm_cuda_set_meteo_l2, mrm_net_init, mrm_net_l1_l11_mapping, &
            mrm_net_flow_accumulation, mrm_net_calc_celerity, mrm_net_update_param, &
            mhm_cuda_unregister_domain, mhm_cuda_version, mhm_cuda_set_math_mode, mhm_cuda_synchronize, &
            mhm_cuda_set_meteo_device, mhm_cuda_get_meteo, mhm_cuda_keep_runoff_history, &
            mhm_cuda_get_runoff_history, mhm_cuda_event_record, mhm_cuda_event_elapsed_ms, &
            mhm_cuda_kernel_stats, mhm_cuda_kernel_stats_reset, mhm_cuda_measure_dfma_peak, &
            mpr_cuda_grid_destroy, mpr_cuda_upscale_geometric_mean, mhm_grid_init_lowres_level, &
            mhm_time_indices, mhm_cuda_comm_unique_id, mhm_cuda_comm_init, mhm_cuda_comm_finalize, &
            mhm_cuda_comm_info, mhm_cuda_meteo_shared_rows, mhm_cuda_set_meteo_shared, mhm_cuda_meteo_h2d_bytes, &
            mrm_cuda_set_exchange, mrm_cuda_shard_run_steps, mrm_cuda_shard_flush
  public :: mpr_l0_inputs, mpr_soil_db, mhm_optisim_config

contains

  !> turn a library status into the reference's fatal path (FORCES mo_message::error_message)
  subroutine mhm_cuda_check(ierr)
    use mo_message, only : error_message
    integer(c_int), intent(in) :: ierr
    character(kind = c_char), pointer :: cmsg(:)
    character(len = 1024) :: msg
    integer :: i
    if (ierr == 0) return
    call c_f_pointer(mhm_cuda_last_error(), cmsg, [1024])
    msg = ''
    do i = 1, 1024
      if (cmsg(i) == c_null_char) exit
      msg(i : i) = cmsg(i)
    end do
    call error_message('mhm_cuda: ', trim(msg))
  end subroutine mhm_cuda_check

end module mo_mhm_cuda
