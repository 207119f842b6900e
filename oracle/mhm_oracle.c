/*
 * mhm_oracle.c -- CPU restatement of the mHM L1 cell cascade, meteo prologue and
 * mRM Muskingum routing.  TEST INFRASTRUCTURE ONLY (see mhm_oracle.h).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off: gfortran -O3 on generic x86-64 emits no FMA, so neither may we.
 * exp/log/pow are glibc's, the same libm a gfortran build of the reference calls.
 */
#include "mhm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================================================================
 *  Cell processes
 * ==========================================================================*/

/* mHM/mo_canopy_interc.f90:75-133 */
void orc_canopy_interc(double pet, double interc_max, double precip, double *interc,
                       double *throughfall, double *evap_canopy) {
  double aux_help = *interc + precip; /* :105 */
  if (aux_help >= interc_max) {
    *throughfall = aux_help - interc_max;
    *interc = interc_max;
  } else {
    *throughfall = 0.0;
    *interc = aux_help;
  }
  if (interc_max > ORC_EPS_DP) { /* :116-121 */
    *evap_canopy = pet * pow(*interc / interc_max, ORC_TWOTHIRD);
  } else {
    *evap_canopy = 0.0;
  }
  if (*evap_canopy < 0.0) *evap_canopy = 0.0; /* :124 */
  if (*interc > *evap_canopy) {               /* :126-131 */
    *interc = *interc - *evap_canopy;
  } else {
    *evap_canopy = *interc;
    *interc = 0.0;
  }
}

/* mHM/mo_snow_accum_melt.f90:70-158 */
void orc_snow_accum_melt(double deg_day_incr, double deg_day_max, double deg_day_noprec,
                         double prec, double temperature, double temperature_thresh,
                         double thrfall, double *snow_pack, double *deg_day, double *melt,
                         double *prec_effect, double *rain, double *snow) {
  double aux_help;
  if (temperature > temperature_thresh) { /* :118-124 */
    *snow = 0.0;
    *rain = thrfall;
  } else {
    *snow = thrfall;
    *rain = 0.0;
  }
  if (prec <= (deg_day_max - deg_day_noprec) / deg_day_incr) { /* :127-131 */
    *deg_day = deg_day_noprec + deg_day_incr * prec;
  } else {
    *deg_day = deg_day_max;
  }
  if (temperature > temperature_thresh) { /* :134-153 */
    if (*snow_pack > 0.0) {
      aux_help = *deg_day * (temperature - temperature_thresh);
      if (aux_help > *snow_pack) {
        *melt = *snow_pack;
        *snow_pack = 0.0;
      } else {
        *melt = aux_help;
        *snow_pack = *snow_pack - aux_help;
      }
    } else {
      *melt = 0.0;
      *snow_pack = 0.0;
    }
  } else {
    *melt = 0.0;
    *snow_pack = *snow_pack + *snow;
  }
  *prec_effect = *melt + *rain; /* :156 */
}

/* mHM/mo_soil_moisture.f90:333-363 */
double orc_feddes_et_reduction(double soil_moist, double soil_moist_FC, double wilting_point,
                               double frac_roots) {
  if (soil_moist >= soil_moist_FC) {
    return frac_roots;
  } else if (soil_moist > wilting_point) {
    return frac_roots * (soil_moist - wilting_point) / (soil_moist_FC - wilting_point);
  }
  return 0.0;
}

/* mHM/mo_soil_moisture.f90:405-446 */
double orc_jarvis_et_reduction(double soil_moist, double soil_moist_sat, double wilting_point,
                               double frac_roots, double jarvis_thresh_c1) {
  double theta_inorm = (soil_moist - wilting_point) / (soil_moist_sat - wilting_point);
  double r = 0.0; /* the reference leaves the result undefined for NaN input */
  if (theta_inorm < 0.0) theta_inorm = 0.0;
  if (theta_inorm > 1.0) theta_inorm = 1.0;
  if (theta_inorm >= jarvis_thresh_c1) {
    r = frac_roots;
  } else if (theta_inorm < jarvis_thresh_c1) {
    r = frac_roots * (theta_inorm / jarvis_thresh_c1);
  }
  return r;
}

/* mHM/mo_soil_moisture.f90:94-290 */
void orc_soil_moisture(int32_t processCase, double frac_sealed, double water_thresh_sealed,
                       double pet, double evap_coeff, int32_t nH, int64_t st,
                       const double *soil_moist_sat, const double *frac_roots,
                       const double *soil_moist_FC, const double *wilting_point,
                       const double *soil_moist_exponen, double jarvis_thresh_c1,
                       double aet_canopy, double prec_effec, double *runoff_sealed,
                       double *storage_sealed, double *infiltration, double *soil_moist,
                       double *aet, double *aet_sealed) {
  int32_t hh, j;
  double prec_effec_soil, frac_runoff, soil_stress_factor = 0.0, tmp;

  *runoff_sealed = 0.0; /* :179-180 */
  *aet_sealed = 0.0;
  if (frac_sealed > 0.0) { /* :183-213 */
    tmp = *storage_sealed + prec_effec;
    if (tmp > water_thresh_sealed) {
      *runoff_sealed = tmp - water_thresh_sealed;
      *storage_sealed = water_thresh_sealed;
    } else {
      *runoff_sealed = 0.0;
      *storage_sealed = tmp;
    }
    if (water_thresh_sealed > ORC_EPS_DP) {
      *aet_sealed = (pet / evap_coeff - aet_canopy) * (*storage_sealed / water_thresh_sealed);
      if (*aet_sealed < 0.0) *aet_sealed = 0.0;
    } else {
      *aet_sealed = DBL_MAX; /* huge(1.0_dp) :199 */
    }
    if (*storage_sealed > *aet_sealed) {
      *storage_sealed = *storage_sealed - *aet_sealed;
    } else {
      *aet_sealed = *storage_sealed;
      *storage_sealed = 0.0;
    }
  }

  for (hh = 0; hh < nH; hh++) { /* :216-217 */
    aet[hh * st] = 0.0;
    infiltration[hh * st] = 0.0;
  }
  prec_effec_soil = prec_effec; /* :220 */

  for (hh = 0; hh < nH; hh++) { /* :222-286 */
    double sm = soil_moist[hh * st], sat = soil_moist_sat[hh * st], inf, a;
    if (hh != 0) prec_effec_soil = infiltration[(hh - 1) * st];
    if (sm > sat) { /* :227 */
      inf = prec_effec_soil;
    } else {
      if (sm > ORC_EPS_DP) { /* :232-236 */
        frac_runoff = exp(soil_moist_exponen[hh * st] * log(sm / sat));
      } else {
        frac_runoff = 0.0;
      }
      tmp = prec_effec_soil * (1.0 - frac_runoff); /* :238 */
      if ((sm + tmp) > sat) {                      /* :240-246 */
        inf = prec_effec_soil + (sm - sat);
        sm = sat;
      } else {
        inf = prec_effec_soil - tmp;
        sm = sm + tmp;
      }
    }
    infiltration[hh * st] = inf;

    a = pet - aet_canopy; /* :252 */
    if (hh != 0) {        /* :253: sum(aet(1:hh-1), mask = aet > 0) in index order */
      double s = 0.0;
      for (j = 0; j < hh; j++)
        if (aet[j * st] > 0.0) s = s + aet[j * st];
      a = a - s;
    }
    switch (processCase) { /* :256-268 */
      case 1:
      case 4:
        soil_stress_factor = orc_feddes_et_reduction(sm, soil_moist_FC[hh * st],
                                                     wilting_point[hh * st], frac_roots[hh * st]);
        break;
      case 2:
      case 3:
        soil_stress_factor = orc_jarvis_et_reduction(sm, sat, wilting_point[hh * st],
                                                     frac_roots[hh * st], jarvis_thresh_c1);
        break;
      default:
        break;
    }
    a = a * soil_stress_factor; /* :270 */
    if (a < 0.0) a = 0.0;       /* :273 */
    if (sm > a) {               /* :276-281 */
      sm = sm - a;
    } else {
      a = sm - ORC_EPS_DP;
      sm = ORC_EPS_DP;
    }
    if (sm < ORC_EPS_DP) sm = ORC_EPS_DP; /* :284 */
    aet[hh * st] = a;
    soil_moist[hh * st] = sm;
  }
}

/* mHM/mo_runoff.f90:74-154 */
void orc_runoff_unsat_zone(double k1, double kp, double k0, double alpha, double karst_loss,
                           double pefec_soil, double unsat_thresh, double *sat_storage,
                           double *unsat_storage, double *slow_interflow,
                           double *fast_interflow, double *perc) {
  double a, b;
  *unsat_storage = *unsat_storage + pefec_soil; /* :120 */
  *fast_interflow = 0.0;
  if (*unsat_storage > unsat_thresh) { /* :124-127 */
    a = k0 * (*unsat_storage - unsat_thresh);
    b = *unsat_storage - ORC_EPS_DP;
    *fast_interflow = (a < b) ? a : b;
  }
  *unsat_storage = *unsat_storage - *fast_interflow;
  *slow_interflow = 0.0;
  if (*unsat_storage > ORC_EPS_DP) { /* :133-136 */
    a = k1 * pow(*unsat_storage, 1.0 + alpha);
    b = *unsat_storage - ORC_EPS_DP;
    *slow_interflow = (a < b) ? a : b;
  }
  *unsat_storage = *unsat_storage - *slow_interflow;
  *perc = kp * *unsat_storage; /* :143 */
  if (*unsat_storage > *perc) {
    *unsat_storage = *unsat_storage - *perc;
    *sat_storage = *sat_storage + *perc * karst_loss;
  } else {
    *sat_storage = *sat_storage + *unsat_storage * karst_loss;
    *unsat_storage = 0.0;
  }
}

/* mHM/mo_runoff.f90:191-212 */
void orc_runoff_sat_zone(double k2, double *sat_storage, double *baseflow) {
  if (*sat_storage > 0.0) {
    *baseflow = k2 * *sat_storage;
    *sat_storage = *sat_storage - *baseflow;
  } else {
    *baseflow = 0.0;
    *sat_storage = 0.0;
  }
}

/* mHM/mo_runoff.f90:248-274 */
void orc_L1_total_runoff(double fSealed, double fast_interflow, double slow_interflow,
                         double baseflow, double direct_runoff, double *total_runoff) {
  *total_runoff = ((baseflow + slow_interflow + fast_interflow) * (1.0 - fSealed)) +
                  (direct_runoff * fSealed);
}

/* ============================================================================
 *  PET (mHM/mo_pet.f90) and temporal disaggregation (meteo/mo_meteo_temporal_tools.f90)
 * ==========================================================================*/

/* FORCES mo_utils::le (un-vendored): a <= b with an epsilon*|b| equality band */
static int orc_le(double a, double b) {
  if ((ORC_EPS_DP * fabs(b) - fabs(a - b)) < 0.0) return a < b;
  return 1;
}

/* mo_pet.f90:426-441 */
double orc_sat_vap_pressure(double tavg) {
  return ORC_TETENS_C1 * exp(ORC_TETENS_C2 * tavg / (tavg + ORC_TETENS_C3));
}

/* mo_pet.f90:384-399 */
double orc_slope_satpressure(double tavg) {
  return ORC_SATPRESSURESLOPE1 * orc_sat_vap_pressure(tavg) /
         exp(2.0 * log(tavg + ORC_TETENS_C3));
}

/* mo_pet.f90:314-355 */
double orc_extraterr_rad_approx(int32_t doy, double latitude) {
  double dr, delta, omega, arg;
  dr = 1.0 + ORC_DUFFIEDR * cos(ORC_TWOPI * doy / ORC_YEARDAYS);
  delta = ORC_DUFFIEDELTA1 * sin(ORC_TWOPI * doy / ORC_YEARDAYS - ORC_DUFFIEDELTA2);
  arg = -tan(latitude) * tan(delta);
  if (arg < -1.0) arg = -1.0;
  if (arg > 1.0) arg = 1.0;
  omega = acos(arg);
  return ORC_DAYSECS / ORC_PI / ORC_SPECHEATET * ORC_SOLARCONST * dr *
         (omega * sin(latitude) * sin(delta) + cos(latitude) * cos(delta) * sin(omega));
}

/* mo_pet.f90:73-116 */
double orc_pet_hargreaves(double HarSamCoeff, double HarSamConst, double tavg, double tmax,
                          double tmin, double latitude, int32_t doy) {
  double delta_temp = tmax - tmin;
  if (orc_le(delta_temp, 0.0) || orc_le(tavg, -HarSamConst)) return 0.0;
  return HarSamCoeff * orc_extraterr_rad_approx(doy, ORC_DEG2RAD * latitude) *
         (tavg + HarSamConst) * sqrt(delta_temp);
}

/* mo_pet.f90:149-174 */
double orc_pet_priestly(double PrieTayParam, double Rn, double tavg) {
  double delta = orc_slope_satpressure(tavg);
  return PrieTayParam * delta / (ORC_PSYCHRO + delta) * (Rn * ORC_DAYSECS / ORC_SPECHEATET);
}

/* mo_pet.f90:235-272 */
double orc_pet_penman(double net_rad, double tavg, double act_vap_pressure,
                      double aerodyn_resistance, double bulksurface_resistance, double a_s,
                      double a_sh) {
  return ORC_DAYSECS / ORC_SPECHEATET *
         (orc_slope_satpressure(tavg) * net_rad +
          ORC_RHO0 * ORC_CP0 * (orc_sat_vap_pressure(tavg) - act_vap_pressure) * a_sh /
              aerodyn_resistance) /
         (orc_slope_satpressure(tavg) +
          ORC_PSYCHRO * a_sh / a_s * (1.0 + bulksurface_resistance / aerodyn_resistance));
}

/* mo_meteo_temporal_tools.f90:32-59 */
double orc_temporal_disagg_meteo_weights(double v, double w, double c) {
  return (v + c) * w - c;
}

/* mo_meteo_temporal_tools.f90:66-101 */
double orc_temporal_disagg_flux_daynight(int32_t isday, double ntimesteps_day, double v,
                                         double fday, double fnight) {
  if (ntimesteps_day > 1.0) {
    if (isday) return 2.0 * v * fday / ntimesteps_day;
    return 2.0 * v * fnight / ntimesteps_day;
  }
  return v;
}

/* mo_meteo_temporal_tools.f90:108-161 */
double orc_temporal_disagg_state_daynight(int32_t isday, double ntimesteps_day, double v,
                                          double fday, double fnight, int32_t add_correction) {
  if (ntimesteps_day > 1.0) {
    if (add_correction) return isday ? v + fday : v + fnight;
    return isday ? 2.0 * v * fday : 2.0 * v * fnight;
  }
  return v;
}

/* ============================================================================
 *  Calendar.  FORCES mo_julian (un-vendored) julday/caldat are the Numerical
 *  Recipes routines with the Gregorian switch of 15 Oct 1582; restated here.
 * ==========================================================================*/
int32_t orc_julday(int32_t dd, int32_t mm, int32_t yy) {
  const int32_t IGREG = 15 + 31 * (10 + 12 * 1582);
  int32_t jy = yy, jm, ja, jul;
  if (jy < 0) jy = jy + 1;
  if (mm > 2) {
    jm = mm + 1;
  } else {
    jy = jy - 1;
    jm = mm + 13;
  }
  jul = (int32_t)floor(365.25 * jy) + (int32_t)floor(30.6001 * jm) + dd + 1720995;
  if (dd + 31 * (mm + 12 * yy) >= IGREG) {
    ja = (int32_t)(0.01 * jy);
    jul = jul + 2 - ja + (int32_t)(0.25 * ja);
  }
  return jul;
}

void orc_caldat(int32_t julian, int32_t *dd, int32_t *mm, int32_t *yy) {
  const int32_t IGREG = 2299161;
  int32_t ja, jalpha, jb, jc, jd, je;
  if (julian >= IGREG) {
    jalpha = (int32_t)(((julian - 1867216) - 0.25) / 36524.25);
    ja = julian + 1 + jalpha - (int32_t)(0.25 * jalpha);
  } else {
    ja = julian;
  }
  jb = ja + 1524;
  jc = (int32_t)(6680.0 + ((jb - 2439870) - 122.1) / 365.25);
  jd = 365 * jc + (int32_t)(0.25 * jc);
  je = (int32_t)((jb - jd) / 30.6001);
  *dd = jb - jd - (int32_t)(30.6001 * je);
  *mm = je - 1;
  if (*mm > 12) *mm = *mm - 12;
  *yy = jc - 4715;
  if (*mm > 2) *yy = *yy - 1;
  if (*yy <= 0) *yy = *yy - 1;
}

int32_t orc_doy(int32_t dd, int32_t mm, int32_t yy) {
  return orc_julday(dd, mm, yy) - orc_julday(1, 1, yy) + 1;
}

/* ============================================================================
 *  Routing
 * ==========================================================================*/

/* mRM/mo_mrm_pre_routing.f90:77-143 */
void orc_L11_runoff_acc(int32_t nCells1, int32_t nNodes, const double *qAll,
                        const double *efecArea, const int32_t *L1_L11_Id,
                        const double *L11_areaCell, const int32_t *L11_L1_Id, int32_t TS,
                        int32_t map_flag, double *qAcc) {
  int32_t k;
  double TST = ORC_HOURSECS * TS; /* :110 */
  if (map_flag) {                 /* :112-130 */
    for (k = 0; k < nNodes; k++) qAcc[k] = 0.0;
    for (k = 0; k < nCells1; k++)
      qAcc[L1_L11_Id[k] - 1] = qAcc[L1_L11_Id[k] - 1] + qAll[k] * efecArea[k];
    for (k = 0; k < nNodes; k++) qAcc[k] = qAcc[k] * 1000.0 / TST;
  } else { /* :132-141 */
    for (k = 0; k < nNodes; k++) qAcc[k] = qAll[L11_L1_Id[k] - 1];
    for (k = 0; k < nNodes; k++) qAcc[k] = qAcc[k] * L11_areaCell[k] * 1000.0 / TST;
  }
}

/* mRM/mo_mrm_pre_routing.f90:179-214 */
void orc_add_inflow(int32_t nInflowGauges, const int32_t *InflowIndexList,
                    const int32_t *InflowHeadwater, const int32_t *InflowNodeList,
                    const double *QInflow, double *qOut) {
  int32_t ii;
  for (ii = 0; ii < nInflowGauges; ii++) {
    if (InflowHeadwater[ii]) {
      qOut[InflowNodeList[ii] - 1] = qOut[InflowNodeList[ii] - 1] + QInflow[InflowIndexList[ii] - 1];
    } else {
      qOut[InflowNodeList[ii] - 1] = QInflow[InflowIndexList[ii] - 1];
    }
  }
}

/* mRM/mo_mrm_routing.f90:380-481.  qTIN/qTR are Fortran (nNodes, 2): [0]=IT1 (past), [1]=IT */
void orc_L11_routing(int32_t nNodes, int32_t nLinks, const int32_t *netPerm,
                     const int32_t *fromN, const int32_t *toN, const double *C1,
                     const double *C2, const double *qOUT, int32_t nInflowGauges,
                     const int32_t *InflowHeadwater, const int32_t *InflowNodeList,
                     double *qTIN, double *qTR, double *Qmod) {
  double *qTIN1 = qTIN, *qTIN2 = qTIN + nNodes, *qTR1 = qTR, *qTR2 = qTR + nNodes;
  int32_t g, i, k, iNode, tNode;
  for (k = 0; k < nNodes; k++) qTIN2[k] = 0.0; /* :428 */
  for (k = 0; k < nLinks; k++) {               /* :435-459 */
    i = netPerm[k] - 1;
    iNode = fromN[i] - 1;
    tNode = toN[i] - 1;
    qTIN2[iNode] = qTIN2[iNode] + qOUT[iNode];
    qTR2[iNode] = qTR1[iNode] + C1[i] * (qTIN1[iNode] - qTR1[iNode]) +
                  C2[i] * (qTIN2[iNode] - qTIN1[iNode]);
    if (nInflowGauges > 0) {
      for (g = 0; g < nInflowGauges; g++)
        if ((tNode + 1 == InflowNodeList[g]) && (!InflowHeadwater[g])) qTR2[iNode] = 0.0;
    }
    qTIN2[tNode] = qTIN2[tNode] + qTR2[iNode];
  }
  tNode = toN[netPerm[nLinks - 1] - 1] - 1; /* :466-467: ONLY the last link's sink */
  qTIN2[tNode] = qTIN2[tNode] + qOUT[tNode];
  for (k = 0; k < nNodes; k++) { /* :474-478 */
    Qmod[k] = qTIN2[k];
    qTR1[k] = qTR2[k];
    qTIN1[k] = qTIN2[k];
  }
}

/* mRM/mo_mrm_mpr.f90:61-119.  length/slope carry nSlope (= nNodes-1) entries at the
 * call site (mo_mhm_interface_run.f90:577-578); ssMax is the maxval over all of them. */
void orc_reg_rout(const double *param, int32_t nLinks, int32_t nSlope, const double *length,
                  const double *slope, const double *fFPimp, double TS, double *C1, double *C2) {
  int32_t i;
  double ssMax = slope[0], K, xi;
  for (i = 1; i < nSlope; i++)
    if (slope[i] > ssMax) ssMax = slope[i];
  for (i = 0; i < nLinks; i++) {
    K = param[0] + param[1] * (length[i] * 0.001) + param[2] * slope[i] + param[3] * fFPimp[i];
    xi = param[4] * (1.0 + slope[i] / ssMax);
    if (xi > 0.5) xi = 0.5;
    if (xi < 0.005) xi = 0.005;
    if (K > 0.5 * TS / xi) K = 0.5 * TS / xi;
    if (K < 0.5 * TS / (1.0 - xi)) K = 0.5 * TS / (1.0 - xi);
    C1[i] = TS / (K * (1.0 - xi) + 0.5 * TS);
    C2[i] = 1.0 - C1[i] * K / TS;
  }
}

/* mRM/mo_mrm_constants.F90:42-46 */
static const double given_TS[19] = {60.,   120.,  180.,  240.,   300.,   360.,   600.,
                                    720.,  900.,  1200., 1800.,  3600.,  7200.,  10800.,
                                    14400., 21600., 28800., 43200., 86400.};

/* FORCES mo_utils::locate (un-vendored): j with xx(j) <= x < xx(j+1); 0 / n outside */
static int32_t orc_locate(const double *xx, int32_t n, double x) {
  int32_t jl = 0, ju = n + 1, jm;
  while (ju - jl > 1) {
    jm = (ju + jl) / 2;
    if (x >= xx[jm - 1]) jl = jm; else ju = jm;
  }
  if (x == xx[0]) return 1;
  if (x == xx[n - 1]) return n - 1;
  return jl;
}

/* mRM/mo_mrm_mpr.f90:241-329, processCase(8)=2: K = length / celerity */
double orc_mrm_update_param_case2(int32_t nNodes, int32_t nOutlets, const double *length,
                                  double celerity, double *C1, double *C2) {
  int32_t i, ind;
  double xi = fabs(0.0), kmin, TSrout; /* rout_space_weight = 0 */
  double *K = (double *)malloc(sizeof(double) * (size_t)nNodes);
  for (i = 0; i < nNodes; i++) K[i] = length[i] / celerity;
  kmin = K[0];
  for (i = 1; i < nNodes - nOutlets; i++)
    if (K[i] < kmin) kmin = K[i];
  ind = orc_locate(given_TS, 19, kmin);
  if (ind < 1) ind = 1;
  TSrout = given_TS[ind - 1];
  for (i = 0; i < nNodes; i++) {
    C1[i] = TSrout / (K[i] * (1.0 - xi) + 0.5 * TSrout);
    C2[i] = 1.0 - C1[i] * K[i] / TSrout;
  }
  free(K);
  return TSrout;
}

/* ---- routing case 3 ------------------------------------------------------------------- */
#define A2(i, j, nr) ((size_t)((j) - 1) * (size_t)(nr) + (size_t)((i) - 1))
static void orc_move_down(int32_t fdir, int32_t *i, int32_t *j) { /* moveDownOneCell :1768-1801 */
  switch (fdir) {
    case 1: *j += 1; break;
    case 2: *i += 1; *j += 1; break;
    case 4: *i += 1; break;
    case 8: *i += 1; *j -= 1; break;
    case 16: *j -= 1; break;
    case 32: *i -= 1; *j -= 1; break;
    case 64: *i -= 1; break;
    case 128: *i -= 1; *j += 1; break;
    default: break;
  }
}
/* mRM/mo_mrm_net_startup.f90:1368-1412: streamNet0(frow, fcol) = ii for the from-cell and for
 * every cell entered by moveDownOneCell until the to-cell is reached */
void orc_stream_net(int32_t nrows0, int32_t ncols0, const int32_t *fDir0, int32_t nLinks,
                    const int32_t *netPerm, const int32_t *fRow, const int32_t *fCol,
                    const int32_t *tRow, const int32_t *tCol, int32_t *streamNet0) {
  int32_t rr, ii, fr, fc;
  size_t a;
  for (a = 0; a < (size_t)nrows0 * (size_t)ncols0; a++) streamNet0[a] = -9999;
  for (rr = 1; rr <= nLinks; rr++) {
    ii = netPerm[rr - 1];
    fr = fRow[ii - 1];
    fc = fCol[ii - 1];
    streamNet0[A2(fr, fc, nrows0)] = ii;
    while (!(fr == tRow[ii - 1] && fc == tCol[ii - 1])) { /* fId == tId <=> same cell */
      orc_move_down(fDir0[A2(fr, fc, nrows0)], &fr, &fc);
      streamNet0[A2(fr, fc, nrows0)] = ii;
    }
  }
}
static int orc_cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}
/* :1440-1446 with FORCES percentile (inverse empirical CDF): the ceiling(n k / 100)-th smallest */
void orc_length_floor(int32_t nNodes, double *length) {
  double *v = (double *)malloc(sizeof(double) * (size_t)nNodes), p;
  int32_t n = 0, i, kk;
  for (i = 0; i < nNodes; i++)
    if (length[i] >= 0.0) v[n++] = length[i];
  if (n > 2) {
    qsort(v, (size_t)n, sizeof(double), orc_cmp_double);
    kk = (int32_t)ceil((double)n * 40.0 / 100.0);
    if (kk < 1) kk = 1;
    if (kk > n) kk = n;
    p = v[kk - 1];
    for (i = 0; i < nNodes; i++)
      if (!(length[i] > p)) length[i] = p; /* merge(len, p, len > p) */
  }
  free(v);
}
static double orc_median_masked(int32_t n, const double *arr, const int32_t *mask) {
  double *v = (double *)malloc(sizeof(double) * (size_t)n), m;
  int32_t k = 0, i;
  for (i = 0; i < n; i++)
    if (mask[i]) v[k++] = arr[i];
  qsort(v, (size_t)k, sizeof(double), orc_cmp_double);
  m = (k % 2) ? v[k / 2] : 0.5 * (v[k / 2 - 1] + v[k / 2]);
  free(v);
  return m;
}
/* mad_val with tout = 'u': mval marks missing values -- entries equal to it take no part in the
 * statistics; entries of the sample above median + z * MAD / 0.6745 are cut back to that bound.
 * (FORCES source not available; this reading reproduces check/case_13's discharge to 1e-15,
 * "replace by mval" and "mval entries included" do not.) */
void orc_mad_upper(int32_t n, double *arr, double z, const int32_t *mask, double mval) {
  double med, mabsdev, thresh;
  double *d = (double *)malloc(sizeof(double) * (size_t)n);
  int32_t *m = (int32_t *)malloc(sizeof(int32_t) * (size_t)n), i, cnt = 0;
  for (i = 0; i < n; i++) {
    m[i] = mask[i] && arr[i] != mval;
    cnt += m[i];
  }
  if (cnt > 0) {
    med = orc_median_masked(n, arr, m);
    for (i = 0; i < n; i++) d[i] = fabs(arr[i] - med);
    mabsdev = orc_median_masked(n, d, m);
    thresh = mabsdev * z / 0.6745;
    for (i = 0; i < n; i++)
      if (m[i] && arr[i] > med + thresh) arr[i] = med + thresh;
  }
  free(m);
  free(d);
}
/* mRM/mo_mrm_net_startup.f90:2343-2409 */
void orc_calc_celerity(int32_t nrows0, int32_t ncols0, const int32_t *mask0, const int32_t *fDir0,
                       const int32_t *streamNet0, const double *slope0_packed, int32_t nNodes,
                       int32_t nLinks, const int32_t *netPerm, const int32_t *fRow, const int32_t *fCol,
                       const int32_t *tRow, const int32_t *tCol, double param, double *celerity11) {
  size_t ng = (size_t)nrows0 * (size_t)ncols0, a;
  int32_t nCells0 = 0, c, rr, ii, fr, fc, ns, cnt, k;
  double *slope_tmp, *slope0, *stack, s;
  int32_t *smask;
  for (k = 0; k < nNodes; k++) celerity11[k] = -9999.0;
  if (nNodes <= 1) {
    for (k = 0; k < nNodes; k++) celerity11[k] = 1.0;
    return;
  }
  for (a = 0; a < ng; a++) nCells0 += mask0[a] ? 1 : 0;
  slope_tmp = (double *)malloc(sizeof(double) * (size_t)nCells0);
  smask = (int32_t *)malloc(sizeof(int32_t) * (size_t)nCells0);
  cnt = 0;
  for (c = 0; c < nCells0; c++) {
    slope_tmp[c] = slope0_packed[c] < 0.1 ? 0.1 : slope0_packed[c]; /* :2350-2351 */
    smask[c] = streamNet0[c] != -9999;                               /* :2354 */
    cnt += smask[c];
  }
  if (cnt > 1) orc_mad_upper(nCells0, slope_tmp, 2.25, smask, 0.1); /* :2357-2359 */
  slope0 = (double *)malloc(sizeof(double) * ng); /* UNPACK :2361 */
  c = 0;
  for (a = 0; a < ng; a++) slope0[a] = mask0[a] ? slope_tmp[c++] : -9999.0;
  stack = (double *)malloc(sizeof(double) * (ng + 1));
  for (rr = 1; rr <= nLinks; rr++) { /* :2370-2401 */
    ii = netPerm[rr - 1];
    fr = fRow[ii - 1];
    fc = fCol[ii - 1];
    ns = 0;
    for (;;) {
      stack[ns++] = param * sqrt(slope0[A2(fr, fc, nrows0)] / 100.0);
      if (fr == tRow[ii - 1] && fc == tCol[ii - 1]) break;
      orc_move_down(fDir0[A2(fr, fc, nrows0)], &fr, &fc);
    }
    s = 0.0;
    for (k = 0; k < ns; k++) s = s + 1.0 / stack[k];
    celerity11[ii - 1] = (double)ns / s;
  }
  free(stack);
  free(slope0);
  free(smask);
  free(slope_tmp);
}
/* mRM/mo_mrm_mpr.f90:298-321 */
double orc_mrm_update_param_case3(int32_t nNodes, int32_t nOutlets, const double *length,
                                  const double *celerity11, double *C1, double *C2) {
  int32_t i, ind;
  double xi = fabs(0.0), kmin, TSrout;
  double *K = (double *)malloc(sizeof(double) * (size_t)nNodes);
  for (i = 0; i < nNodes; i++) K[i] = length[i] / celerity11[i];
  kmin = K[0];
  for (i = 1; i < nNodes - nOutlets; i++)
    if (K[i] < kmin) kmin = K[i];
  ind = orc_locate(given_TS, 19, kmin);
  if (ind < 1) ind = 1;
  TSrout = given_TS[ind - 1];
  for (i = 0; i < nNodes; i++) {
    C1[i] = TSrout / (K[i] * (1.0 - xi) + 0.5 * TSrout);
    C2[i] = 1.0 - C1[i] * K[i] / TSrout;
  }
  free(K);
  return TSrout;
}
/* mRM/mo_mrm_net_startup.f90:2084-2161 */
static void orc_facc_rec(const int32_t *fDir, double *fAcc, int32_t ii, int32_t jj, int32_t nrow, int32_t ncol) {
  static const int32_t di[8] = {0, 1, 1, 1, 0, -1, -1, -1}, dj[8] = {1, 1, 0, -1, -1, -1, 0, 1};
  static const int32_t code[8] = {16, 32, 64, 128, 1, 2, 4, 8};
  int q;
  for (q = 0; q < 8; q++) {
    int32_t ni = ii + di[q], nj = jj + dj[q];
    if (ni < 1 || ni > nrow || nj < 1 || nj > ncol) continue;
    if (fDir[A2(ni, nj, nrow)] == code[q]) {
      orc_facc_rec(fDir, fAcc, ni, nj, nrow, ncol);
      fAcc[A2(ii, jj, nrow)] = fAcc[A2(ii, jj, nrow)] + fAcc[A2(ni, nj, nrow)];
    }
  }
}
void orc_flow_accumulation(int32_t nrows11, int32_t ncols11, const int32_t *mask11,
                           const int32_t *fDir11_packed, const double *cellarea11_packed, double *fAcc11_packed) {
  size_t ng = (size_t)nrows11 * (size_t)ncols11, a, k = 0;
  int32_t *fd = (int32_t *)malloc(sizeof(int32_t) * ng), ii, jj;
  double *acc = (double *)malloc(sizeof(double) * ng);
  const double f = (double)1.e-6f; /* default-real literal in `cellarea * 1.e-6` :2057 */
  for (a = 0; a < ng; a++) {
    fd[a] = mask11[a] ? fDir11_packed[k] : -9999;
    acc[a] = mask11[a] ? cellarea11_packed[k] * f : -9999.0;
    k += mask11[a] ? 1 : 0;
  }
  for (jj = 1; jj <= ncols11; jj++)
    for (ii = 1; ii <= nrows11; ii++)
      if (fd[A2(ii, jj, nrows11)] == 0) orc_facc_rec(fd, acc, ii, jj, nrows11, ncols11);
  k = 0;
  for (a = 0; a < ng; a++)
    if (mask11[a]) fAcc11_packed[k++] = acc[a];
  free(fd);
  free(acc);
}

/* mRM/mo_mrm_routing.f90:104-303 for one call; GaugeDischarge = mRM_runoff(tt, :) with
 * Fortran (nTimeSteps, nGaugesTotal) layout -> element (tt, g) at [ (g-1)*nTimeSteps + tt-1 ] */
static void orc_mRM_routing(orc_domain *d, int32_t tt, const double *RunToRout,
                            int32_t timestep_rout, double tsRoutFactor, int32_t yId,
                            const double *InflowDischarge, double *qAcc) {
  int32_t nNodes = d->nNodes, nLinks = d->nNodes - d->nOutlets, k, gg, t, rout_loop;
  if (d->pc_rout == 1 && !d->read_states) { /* :211-218 */
    if (nNodes > 1)
      orc_reg_rout(d->rout_param, nLinks, nNodes - 1, d->L11_length, d->L11_slope,
                   d->L11_nLinkFracFPimp + (size_t)(yId - 1) * nNodes, (double)timestep_rout,
                   d->L11_C1, d->L11_C2);
  }
  rout_loop = (int32_t)lround(1.0 / tsRoutFactor); /* :224 nint */
  if (rout_loop < 1) rout_loop = 1;
  orc_L11_runoff_acc(d->nCells, nNodes, RunToRout, d->L1_areaCell, d->L1_L11_Id,
                     d->L11_areaCell, d->L11_L1_Id, timestep_rout, d->map_flag, d->L11_qOUT);
  orc_add_inflow(d->nInflowGauges, d->InflowGaugeIndexList, d->InflowGaugeHeadwater,
                 d->InflowGaugeNodeList, InflowDischarge, d->L11_qOUT);
  if (nNodes > 1) { /* :243-282 */
    for (k = 0; k < nNodes; k++) qAcc[k] = 0.0;
    for (t = 0; t < rout_loop; t++) {
      orc_L11_routing(nNodes, nLinks, d->netPerm, d->fromN, d->toN, d->L11_C1, d->L11_C2,
                      d->L11_qOUT, d->nInflowGauges, d->InflowGaugeHeadwater,
                      d->InflowGaugeNodeList, d->L11_qTIN, d->L11_qTR, d->L11_qMod);
      for (k = 0; k < nNodes; k++) qAcc[k] = qAcc[k] + d->L11_qMod[k];
    }
    for (k = 0; k < nNodes; k++) d->L11_qMod[k] = qAcc[k] / (double)rout_loop;
  } else {
    for (k = 0; k < nNodes; k++) d->L11_qMod[k] = d->L11_qOUT[k];
  }
  for (gg = 0; gg < d->nGauges; gg++) /* :299-301 */
    d->mRM_runoff[(size_t)(d->gaugeIndexList[gg] - 1) * d->nTimeSteps + (tt - 1)] =
        d->L11_qMod[d->gaugeNodeList[gg] - 1];
}

/* ============================================================================
 *  Time loop
 * ==========================================================================*/

/* common/mo_common_datetime_type.f90:28-155: date state carried through the loop */
typedef struct {
  int32_t year, month, day, hour, iLAI, yId;
  int32_t is_new_day, is_new_month, is_new_year;
} orc_dt;

static int32_t lc_yid(const orc_domain *d, int32_t year) {
  int32_t i = year - d->lc_year_start;
  if (i < 0) i = 0;
  if (i >= d->lc_nyears) i = d->lc_nyears - 1;
  return d->LCyearId[i];
}

static void dt_init(const orc_domain *d, orc_dt *s) { /* :71-108 */
  orc_caldat(d->jul_start, &s->day, &s->month, &s->year);
  s->is_new_day = s->is_new_month = s->is_new_year = 1;
  s->yId = lc_yid(d, s->year);
  s->hour = 0;
  s->iLAI = 0;
}

static void dt_update_LAI(const orc_domain *d, orc_dt *s) { /* :135-155 */
  switch (d->timeStep_LAI_input) {
    case 0:
    case 1: s->iLAI = s->month; break;
    case -1: if (s->is_new_day) s->iLAI++; break;
    case -2: if (s->is_new_month) s->iLAI++; break;
    case -3: if (s->is_new_year) s->iLAI++; break;
    default: break;
  }
}

static void dt_increment(const orc_domain *d, orc_dt *s) { /* :110-133 */
  int32_t pd = s->day, pm = s->month, py = s->year, jul;
  s->is_new_day = s->is_new_month = s->is_new_year = 0;
  s->hour = s->hour + d->timestep_h;
  /* newTime = julday + hour/24; int(newTime) = julday + hour/24 (integer division) */
  jul = orc_julday(s->day, s->month, s->year) + s->hour / 24;
  s->hour = s->hour % 24;
  orc_caldat(jul, &s->day, &s->month, &s->year);
  if (pd != s->day) s->is_new_day = 1;
  if (pm != s->month) s->is_new_month = 1;
  if (py != s->year) s->is_new_year = 1;
}

void orc_time_indices(const orc_domain *d, int32_t n, int32_t *month, int32_t *hour,
                      int32_t *yId, int32_t *iLAI, int32_t *iMeteoTS, int32_t *isday,
                      int32_t *doy, int32_t *year) {
  orc_dt s;
  int32_t tt, per = (int32_t)lround(24.0 / (double)d->nTstepForcingDay);
  dt_init(d, &s);
  for (tt = 1; tt <= n; tt++) {
    dt_update_LAI(d, &s);
    month[tt - 1] = s.month;
    hour[tt - 1] = s.hour;
    yId[tt - 1] = s.yId;
    iLAI[tt - 1] = s.iLAI;
    iMeteoTS[tt - 1] = (int32_t)ceil((double)tt / (double)per); /* mo_meteo_handler.f90:607 */
    isday[tt - 1] = (s.hour > 6) && (s.hour <= 18);               /* :1048 */
    doy[tt - 1] = orc_doy(s.day, s.month, s.year);
    year[tt - 1] = s.year;
    dt_increment(d, &s);
    if (s.is_new_year && tt < d->nTimeSteps) s.yId = lc_yid(d, s.year);
  }
}

#define P3(base, n, dim2, j, y) ((base) + ((size_t)(y) * (dim2) + (j)) * (size_t)(n))

int32_t orc_flux_record_size(int32_t nH) { return 22 + 3 * nH; }

/* one model step for all cells: meteo/mo_meteo_handler.f90:915-1285 (get_corrected_pet,
 * get_temp, get_prec) followed by mHM/mo_mhm.f90:229-521 */
static void orc_step_cells(orc_domain *d, int32_t tt, const orc_dt *s, int32_t iMeteoTS) {
  const int32_t n = d->nCells, nH = d->nH;
  const int32_t month = s->month, hour = s->hour, y = s->yId - 1, il = s->iLAI - 1;
  const int32_t isday = (hour > 6) && (hour <= 18);
  const int32_t doy = orc_doy(s->day, s->month, s->year);
  const size_t mo = (size_t)(iMeteoTS - 1) * n;
  const double nts = (double)d->nTstepDay;
  const double c2TSTu = d->c2TSTu;
  const double *fSealed = P3(d->fSealed, n, 1, 0, y), *alpha = P3(d->alpha, n, 1, 0, y);
  const double *ddinc = P3(d->degDayInc, n, 1, 0, y), *ddmax = P3(d->degDayMax, n, 1, 0, y);
  const double *ddnop = P3(d->degDayNoPre, n, 1, 0, y);
  const double *fRoots = P3(d->fRoots, n, nH, 0, y);
  const double *maxInter = P3(d->maxInter, n, d->nLAI, il, 0);
  const double *k0 = P3(d->kFastFlow, n, 1, 0, y), *k1 = P3(d->kSlowFlow, n, 1, 0, y);
  const double *k2 = P3(d->kBaseFlow, n, 1, 0, y), *kp = P3(d->kPerco, n, 1, 0, y);
  const double *FC = P3(d->soilMoistFC, n, nH, 0, y), *SAT = P3(d->soilMoistSat, n, nH, 0, y);
  const double *EXPN = P3(d->soilMoistExp, n, nH, 0, y), *WP = P3(d->wiltingPoint, n, nH, 0, y);
  const double *tthr = P3(d->tempThresh, n, 1, 0, y);
  const double *petLAI = d->petLAIcorFactor ? P3(d->petLAIcorFactor, n, d->nLAI, il, y) : 0;
  const double *PTa = d->PrieTayAlpha ? P3(d->PrieTayAlpha, n, d->nLAI, il, 0) : 0;
  const double *aeroR = d->aeroResist ? P3(d->aeroResist, n, d->nLAI, il, y) : 0;
  const double *surfR = d->surfResist ? P3(d->surfResist, n, d->nLAI, il, 0) : 0;
  int32_t k;

  if (tt == 1 && !d->read_states) { /* mo_mhm.f90:448-450 */
    size_t i;
    for (i = 0; i < (size_t)n * nH; i++) d->soilMoist[i] = 0.5 * FC[i];
  }

#pragma omp parallel for schedule(static) num_threads(d->num_threads)
  for (k = 0; k < n; k++) {
    double pet = 0.0, v;
    /* ---- get_corrected_pet, mo_meteo_handler.f90:1053-1119 ---- */
    switch (d->pc_pet) {
      case -1: pet = petLAI[k] * d->pet[mo + k]; break;
      case 0: pet = d->fAsp[k] * d->pet[mo + k]; break;
      case 1:
        pet = d->fAsp[k] * orc_pet_hargreaves(d->HarSamCoeff[k], ORC_HARSAMCONST,
                                              d->temp[mo + k], d->tmax[mo + k],
                                              d->tmin[mo + k], d->latitude[k], doy);
        break;
      case 2:
        v = d->netrad[mo + k];
        pet = orc_pet_priestly(PTa[k], v > 0.0 ? v : 0.0, d->temp[mo + k]);
        break;
      case 3:
        v = d->netrad[mo + k];
        pet = orc_pet_penman(v > 0.0 ? v : 0.0, d->temp[mo + k],
                             d->absvappress[mo + k] / 1000.0,
                             aeroR[k] / d->windspeed[mo + k], surfR[k], 1.0, 1.0);
        break;
      default: break;
    }
    if (d->is_hourly_forcing) {
      d->pet_calc[k] = pet;
    } else if (d->read_meteo_weights) {
      d->pet_calc[k] = orc_temporal_disagg_meteo_weights(
          pet, d->pet_weights[((size_t)hour * 12 + (month - 1)) * n + k], 0.0);
    } else {
      d->pet_calc[k] = orc_temporal_disagg_flux_daynight(
          isday, nts, pet, d->fday_pet[month - 1], d->fnight_pet[month - 1]);
    }
    /* ---- get_temp :1166-1201 ---- */
    if (d->is_hourly_forcing) {
      d->temp_calc[k] = d->temp[mo + k];
    } else if (d->read_meteo_weights) {
      d->temp_calc[k] = orc_temporal_disagg_meteo_weights(
          d->temp[mo + k], d->temp_weights[((size_t)hour * 12 + (month - 1)) * n + k], ORC_T0);
    } else {
      d->temp_calc[k] = orc_temporal_disagg_state_daynight(
          isday, nts, d->temp[mo + k], d->fday_temp[month - 1], d->fnight_temp[month - 1], 1);
    }
    /* ---- get_prec :1247-1281 ---- */
    if (d->is_hourly_forcing) {
      d->prec_calc[k] = d->pre[mo + k];
    } else if (d->read_meteo_weights) {
      d->prec_calc[k] = orc_temporal_disagg_meteo_weights(
          d->pre[mo + k], d->pre_weights[((size_t)hour * 12 + (month - 1)) * n + k], 0.0);
    } else {
      d->prec_calc[k] = orc_temporal_disagg_flux_daynight(
          isday, nts, d->pre[mo + k], d->fday_prec[month - 1], d->fnight_prec[month - 1]);
    }

    /* ---- mo_mhm.f90:458-499 ---- */
    orc_canopy_interc(d->pet_calc[k], maxInter[k], d->prec_calc[k], &d->inter[k],
                      &d->throughfall[k], &d->aETCanopy[k]);
    orc_snow_accum_melt(ddinc[k], ddmax[k] * c2TSTu, ddnop[k] * c2TSTu, d->prec_calc[k],
                        d->temp_calc[k], tthr[k], d->throughfall[k], &d->snowPack[k],
                        &d->degDay[k], &d->melt[k], &d->preEffect[k], &d->rain[k], &d->snow[k]);
    orc_soil_moisture(d->pc_soil, fSealed[k], d->sealedThresh[k], d->pet_calc[k],
                      d->evap_coeff[month - 1], nH, n, SAT + k, fRoots + k, FC + k, WP + k,
                      EXPN + k, d->jarvis_thresh_c1[k], d->aETCanopy[k], d->preEffect[k],
                      &d->runoffSeal[k], &d->sealSTW[k], d->infilSoil + k, d->soilMoist + k,
                      d->aETSoil + k, &d->aETSealed[k]);
    orc_runoff_unsat_zone(c2TSTu / k1[k], c2TSTu / kp[k], c2TSTu / k0[k], alpha[k],
                          d->karstLoss[k], d->infilSoil[(size_t)(nH - 1) * n + k],
                          d->unsatThresh[k], &d->satSTW[k], &d->unsatSTW[k],
                          &d->slowRunoff[k], &d->fastRunoff[k], &d->percol[k]);
    orc_runoff_sat_zone(c2TSTu / k2[k], &d->satSTW[k], &d->baseflow[k]);
    orc_L1_total_runoff(fSealed[k], d->fastRunoff[k], d->slowRunoff[k], d->baseflow[k],
                        d->runoffSeal[k], &d->total_runoff[k]);
  }

  if (d->flux_history) {
    const size_t rs = (size_t)orc_flux_record_size(nH);
    double *H = d->flux_history + (size_t)(tt - 1) * rs * n;
    const double *src[22] = {d->pet_calc,  d->temp_calc,  d->prec_calc,  d->aETCanopy,
                             d->aETSealed, d->baseflow,   d->fastRunoff, d->melt,
                             d->percol,    d->preEffect,  d->rain,       d->runoffSeal,
                             d->slowRunoff, d->snow,      d->throughfall, d->total_runoff,
                             d->degDay,    d->inter,      d->snowPack,   d->sealSTW,
                             d->unsatSTW,  d->satSTW};
    int32_t r, h;
    for (r = 0; r < 22; r++) memcpy(H + (size_t)r * n, src[r], sizeof(double) * n);
    for (h = 0; h < nH; h++) {
      memcpy(H + (size_t)(22 + h) * n, d->aETSoil + (size_t)h * n, sizeof(double) * n);
      memcpy(H + (size_t)(22 + nH + h) * n, d->infilSoil + (size_t)h * n, sizeof(double) * n);
      memcpy(H + (size_t)(22 + 2 * nH + h) * n, d->soilMoist + (size_t)h * n, sizeof(double) * n);
    }
  }
}

/* mo_mhm_eval.f90:136-150 + mo_mhm_interface_run.f90:341-638 */
/* ---- gridded outputs ------------------------------------------------------------------ */
int32_t orc_output_slots(const int32_t *f, int32_t nH, int32_t *var, int32_t *hor, int32_t *avg) {
  /* order of the `ii = ii + 1` blocks of mHM_updateDataset (mo_write_fluxes_states.f90:326-436);
   * averaged = created with avg=.true. in mHM_OutputDataset (:98-160): variables 1-8 and 18 */
  static const int32_t order[21] = {1, 2, 3, 4, 5, 6, 7, 8, 18, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 20, 21};
  int32_t n = 0, i, h;
  for (i = 0; i < 21; i++) {
    const int32_t v = order[i];
    const int32_t per_h = (v == 3 || v == 4 || v == 17 || v == 19);
    if (!f[v - 1]) continue;
    for (h = 0; h < (per_h ? nH : 1); h++) {
      if (var) var[n] = v;
      if (hor) hor[n] = per_h ? h : -1;
      if (avg) avg[n] = (v <= 8 || v == 18);
      n++;
    }
  }
  return n;
}

/* mHM_updateDataset for cell k with the land-cover scene yId the driver holds when it calls
 * write_output, i.e. AFTER the date increment of the step (mo_mhm_interface_run.f90:623-628,690) */
static double out_value(const orc_domain *d, int32_t v, int32_t hor, int32_t k, int32_t yId) {
  const size_t n = (size_t)d->nCells;
  const int32_t nH = d->nH;
  const double fS = d->fSealed[(size_t)(yId - 1) * n + k];
  const double fNS = 1.0 - fS; /* L1_fNotSealed, mo_mhm_interface_run.f90:238-239 */
  int32_t h;
  double a, b;
  switch (v) {
    case 1: return d->inter[k];
    case 2: return d->snowPack[k];
    case 3: return d->soilMoist[(size_t)hor * n + k];
    case 4: return d->soilMoist[(size_t)hor * n + k] / d->soilMoistSat[((size_t)(yId - 1) * nH + hor) * n + k];
    case 5:
      a = 0.0;
      b = 0.0;
      for (h = 0; h < nH; h++) a = a + d->soilMoist[(size_t)h * n + k];
      for (h = 0; h < nH; h++) b = b + d->soilMoistSat[((size_t)(yId - 1) * nH + h) * n + k];
      return a / b;
    case 6: return d->sealSTW[k];
    case 7: return d->unsatSTW[k];
    case 8: return d->satSTW[k];
    case 9: return d->pet_calc[k];
    case 10:
      a = 0.0;
      for (h = 0; h < nH; h++) a = a + d->aETSoil[(size_t)h * n + k];
      return a * fNS + d->aETCanopy[k] + d->aETSealed[k] * fS;
    case 11: return d->total_runoff[k];
    case 12: return d->runoffSeal[k] * fS;
    case 13: return d->fastRunoff[k] * fNS;
    case 14: return d->slowRunoff[k] * fNS;
    case 15: return d->baseflow[k] * fNS;
    case 16: return d->percol[k] * fNS;
    case 17: return d->infilSoil[(size_t)hor * n + k] * fNS;
    case 19: return d->aETSoil[(size_t)hor * n + k] * fNS;
    case 20: return d->preEffect[k];
    case 21: return d->melt[k];
    default: return 0.0; /* 18: neutrons, out of scope */
  }
}

/* mhm_interface_run_write_output (:641-742) for the gridded mHM outputs; s is the date AFTER
 * the increment of step tt */
static void write_output(orc_domain *d, int32_t tt, const orc_dt *s) {
  int32_t var[64], hor[64], avg[64], ns, sl, k, wr = 0;
  const int32_t tIndex_out = tt - d->warming_days * d->nTstepDay; /* :621 */
  const int32_t ts = d->timeStep_model_outputs;
  const size_t n = (size_t)d->nCells;
  if (!d->out_acc || tIndex_out <= 0) return;
  ns = orc_output_slots(d->out_flags, d->nH, var, hor, avg);
  if (ns == 0) return;
  for (sl = 0; sl < ns; sl++) /* updateVariable, mo_nc_output.f90:140-149 */
    for (k = 0; k < d->nCells; k++)
      d->out_acc[(size_t)sl * n + k] = d->out_acc[(size_t)sl * n + k] + out_value(d, var[sl], hor[sl], k, s->yId);
  d->out_counter += 1;
  /* datetimeinfo_writeout, mo_common_datetime_type.f90:157-184 */
  if (ts > 0) {
    wr = (tIndex_out % ts == 0) || tt == d->nTimeSteps;
  } else if (ts == 0) {
    wr = tt == d->nTimeSteps;
  } else if (ts == -1) {
    wr = s->is_new_day || tt == d->nTimeSteps;
  } else if (ts == -2) {
    wr = s->is_new_month || tt == d->nTimeSteps;
  } else if (ts == -3) {
    wr = s->is_new_year || tt == d->nTimeSteps;
  }
  if (!wr || d->out_nwin >= d->out_max_windows) return;
  for (sl = 0; sl < ns; sl++) /* writeVariableTimestep, mo_nc_output.f90:158-175 */
    for (k = 0; k < d->nCells; k++) {
      double v = d->out_acc[(size_t)sl * n + k];
      if (avg[sl]) v = v / (double)d->out_counter;
      d->out_win[((size_t)d->out_nwin * ns + sl) * n + k] = v;
      d->out_acc[(size_t)sl * n + k] = 0.0;
    }
  d->out_win_tt[d->out_nwin] = tt;
  d->out_nwin += 1;
  d->out_counter = 0;
}

/* ---- optidata_sim of FORCES mo_optimization_types (restated, see header) ---------------- */
static double *opt_data(orc_domain *d, int32_t w) { return w == 0 ? d->opt_sm : (w == 1 ? d->opt_et : d->opt_tws); }
static int32_t opt_flag(int32_t timeStepInput, const orc_dt *s) {
  if (timeStepInput == -1) return s->is_new_day;
  if (timeStepInput == -2) return s->is_new_month;
  if (timeStepInput == -3) return s->is_new_year;
  return 0;
}
static void opt_add(orc_domain *d, int32_t w, const double *data1d) { /* dataSim(:, averageTimestep) += */
  const size_t n = (size_t)d->nCells, col = (size_t)d->opt_avg_ts[w] - 1;
  double *x = opt_data(d, w);
  int32_t k;
  if ((int32_t)col >= d->opt_ntime[w]) return;
  for (k = 0; k < d->nCells; k++) x[col * n + k] = x[col * n + k] + data1d[k];
}
static void opt_average(orc_domain *d, int32_t w) { /* divide by the counter, next slot, counter = 0 */
  const size_t n = (size_t)d->nCells, col = (size_t)d->opt_avg_ts[w] - 1;
  double *x = opt_data(d, w);
  int32_t k;
  if ((int32_t)col < d->opt_ntime[w])
    for (k = 0; k < d->nCells; k++) x[col * n + k] = x[col * n + k] / (double)d->opt_avg_cnt[w];
  d->opt_avg_ts[w] += 1;
  d->opt_avg_cnt[w] = 0;
}

/* mhm_interface_run_update_optisim, mo_mhm_interface_run.f90:745-861 (neutrons out of scope);
 * s is the date AFTER the increment of step tt, s->yId the land-cover scene held then */
static void update_optisim(orc_domain *d, int32_t tt, const orc_dt *s, double *tmp) {
  const size_t n = (size_t)d->nCells;
  const int32_t nH = d->nH, y = s->yId - 1;
  int32_t k, h;
  if (tt - d->warming_days * d->nTstepDay <= 0) return;
  if (d->opt_on[0]) { /* :776-791 */
    if (opt_flag(d->opt_timestep[0], s)) opt_average(d, 0); /* average_per_timestep */
    if (tt != d->nTimeSteps) {
      for (k = 0; k < d->nCells; k++) {
        double a = 0.0, b = 0.0;
        for (h = 0; h < d->opt_nhor_sm; h++) a = a + d->soilMoist[(size_t)h * n + k];
        for (h = 0; h < d->opt_nhor_sm; h++) b = b + d->soilMoistSat[((size_t)y * nH + h) * n + k];
        tmp[k] = a / b;
      }
      opt_add(d, 0, tmp); /* average_add = add + counter */
      d->opt_avg_cnt[0] += 1;
    }
  }
  if (d->opt_on[1]) { /* :817-833 */
    if (opt_flag(d->opt_timestep[1], s)) d->opt_avg_ts[1] += 1; /* increment_counter */
    if (tt != d->nTimeSteps) {
      for (k = 0; k < d->nCells; k++) {
        const double fS = d->fSealed[(size_t)y * n + k], fNS = 1.0 - fS;
        double a = 0.0;
        for (h = 0; h < nH; h++) a = a + d->aETSoil[(size_t)h * n + k];
        tmp[k] = a * fNS + d->aETCanopy[k] + d->aETSealed[k] * fS;
      }
      opt_add(d, 1, tmp);
    }
  }
  if (d->opt_on[2]) { /* :840-857 */
    if (opt_flag(d->opt_timestep[2], s)) opt_average(d, 2);
    if (tt != d->nTimeSteps) {
      for (k = 0; k < d->nCells; k++)
        tmp[k] = d->inter[k] + d->snowPack[k] + d->sealSTW[k] + d->unsatSTW[k] + d->satSTW[k];
      opt_add(d, 2, tmp);
      d->opt_avg_cnt[2] += 1;
      for (h = 0; h < nH; h++) opt_add(d, 2, d->soilMoist + (size_t)h * n);
    }
  }
}

int32_t orc_run(orc_domain *d, int32_t tt_first, int32_t tt_last) {
  orc_dt s;
  int32_t tt, k, g, jj;
  const int32_t per = (int32_t)lround(24.0 / (double)d->nTstepForcingDay);
  double tsRoutFactor = 1.0, tsRoutFactorIn = 1.0;
  int32_t timestep_rout = d->timestep_h, doRoute;
  double *qAcc = 0, *opt_tmp = 0;

  if (d->do_routing) qAcc = (double *)malloc(sizeof(double) * (size_t)d->nNodes);
  if (d->opt_on[0] || d->opt_on[1] || d->opt_on[2]) opt_tmp = (double *)malloc(sizeof(double) * (size_t)d->nCells);
  dt_init(d, &s);
  for (tt = 1; tt <= tt_last; tt++) {
    int32_t iMeteoTS;
    dt_update_LAI(d, &s); /* mo_mhm_interface_run.f90:355 */
    if (tt >= tt_first) {
      iMeteoTS = (int32_t)ceil((double)tt / (double)per); /* mo_meteo_handler.f90:607 */
      orc_step_cells(d, tt, &s, iMeteoTS);

      /* ---- routing scheduling, mo_mhm_interface_run.f90:460-608 ---- */
      doRoute = 0;
      if (d->do_routing) {
        int32_t iDischargeTS = (int32_t)ceil((double)tt / (double)d->nTstepDay); /* :463 */
        if (d->pc_rout == 1) {                                                     /* :465-474 */
          doRoute = 1;
          tsRoutFactorIn = 1.0;
          timestep_rout = d->timestep_h;
          for (k = 0; k < d->nCells; k++) d->RunToRout[k] = d->total_runoff[k];
          for (g = 0; g < d->nInflowTotal; g++)
            d->InflowDischarge[g] = d->InflowQ[(size_t)g * d->nDays + iDischargeTS - 1];
        } else if (d->pc_rout == 2 || d->pc_rout == 3) { /* :476-513 */
          tsRoutFactor = d->L11_TSrout / (d->timestep_h * ORC_HOURSECS);
          if (tsRoutFactor < 1.0) {
            tsRoutFactorIn = tsRoutFactor;
            for (k = 0; k < d->nCells; k++) d->RunToRout[k] = d->total_runoff[k];
            for (g = 0; g < d->nInflowTotal; g++)
              d->InflowDischarge[g] = d->InflowQ[(size_t)g * d->nDays + iDischargeTS - 1];
            timestep_rout = d->timestep_h;
            doRoute = 1;
          } else {
            tsRoutFactorIn = tsRoutFactor;
            for (k = 0; k < d->nCells; k++)
              d->RunToRout[k] = d->RunToRout[k] + d->total_runoff[k];
            for (g = 0; g < d->nInflowTotal; g++)
              d->InflowDischarge[g] =
                  d->InflowDischarge[g] + d->InflowQ[(size_t)g * d->nDays + iDischargeTS - 1];
            if (tt == d->nTimeSteps && (tt % (int32_t)lround(tsRoutFactorIn)) != 0)
              tsRoutFactorIn = (double)(tt % (int32_t)lround(tsRoutFactorIn));
            if ((tt % (int32_t)lround(tsRoutFactorIn)) == 0 || tt == d->nTimeSteps) {
              for (g = 0; g < d->nInflowTotal; g++)
                d->InflowDischarge[g] = d->InflowDischarge[g] / tsRoutFactorIn;
              timestep_rout = d->timestep_h * (int32_t)lround(tsRoutFactorIn);
              doRoute = 1;
            }
          }
        }
        if (doRoute)
          orc_mRM_routing(d, tt, d->RunToRout, timestep_rout, tsRoutFactorIn, s.yId,
                          d->InflowDischarge, qAcc);
        if (d->pc_rout == 1) { /* :595-598 */
          for (g = 0; g < d->nInflowTotal; g++) d->InflowDischarge[g] = 0.0;
          for (k = 0; k < d->nCells; k++) d->RunToRout[k] = 0.0;
        } else if (d->pc_rout == 2 || d->pc_rout == 3) { /* :599-612 */
          if (!(tsRoutFactorIn < 1.0) && doRoute) {
            for (jj = 1; jj <= (int32_t)lround(tsRoutFactorIn); jj++)
              for (g = 0; g < d->nGaugesTotal; g++)
                d->mRM_runoff[(size_t)g * d->nTimeSteps + (tt - jj)] =
                    d->mRM_runoff[(size_t)g * d->nTimeSteps + (tt - 1)];
            for (g = 0; g < d->nInflowTotal; g++) d->InflowDischarge[g] = 0.0;
            for (k = 0; k < d->nCells; k++) d->RunToRout[k] = 0.0;
          }
        }
      }
    }
    dt_increment(d, &s); /* :623 */
    if (s.is_new_year && tt < d->nTimeSteps) s.yId = lc_yid(d, s.year); /* :626-628 */
    if (tt >= tt_first && d->bfi_on && tt - d->warming_days * d->nTstepDay > 0) { /* :630-636 */
      double sb = 0.0, st = 0.0;
      for (k = 0; k < d->nCells; k++) sb = sb + d->baseflow[k] * d->L1_areaCell[k];
      for (k = 0; k < d->nCells; k++) st = st + d->total_runoff[k] * d->L1_areaCell[k];
      d->bfi_qBF_sum = d->bfi_qBF_sum + sb / (double)d->nCells;
      d->bfi_qT_sum = d->bfi_qT_sum + st / (double)d->nCells;
    }
    if (tt >= tt_first) write_output(d, tt, &s);                        /* mo_mhm_eval.f90:141 */
    if (tt >= tt_first && opt_tmp) update_optisim(d, tt, &s, opt_tmp);  /* mo_mhm_eval.f90:144 */
  }
  free(opt_tmp);
  free(qAcc);
  return 0;
}

/* ---- forcing ingest: L2 -> L1 ---------------------------------------------------------- */
int32_t orc_meteo_l2_to_l1(const double *data2, int32_t nr2, int32_t nc2, int32_t nT, const int32_t *mask2,
                           double cellsize2, int32_t nr1, int32_t nc1, const int32_t *mask1,
                           double cellsize1, double *out_packed, double *out_grid) {
  const double f = cellsize1 / cellsize2; /* cellFactorHbyM, mo_meteo_helper.f90:110 */
  const size_t n2 = (size_t)nr2 * nc2, n1 = (size_t)nr1 * nc1;
  double *g = (double *)malloc(sizeof(double) * n1);
  int32_t *cnt = (int32_t *)calloc(n1, sizeof(int32_t));
  int32_t t, i, j, ncell = 0;
  for (i = 0; i < (int32_t)n1; i++) ncell += mask1[i] != 0;
  if (f > 1.0) /* nTCells, mo_meteo_spatial_tools.f90:160-170 */
    for (j = 1; j <= nc2; j++)
      for (i = 1; i <= nr2; i++) {
        const int32_t jc = (int32_t)ceil((double)j / f), ic = (int32_t)ceil((double)i / f);
        if (mask2[(size_t)(j - 1) * nr2 + (i - 1)]) cnt[(size_t)(jc - 1) * nr1 + (ic - 1)] += 1;
      }
  for (t = 0; t < nT; t++) {
    const double *d2 = data2 + (size_t)t * n2;
    int32_t k = 0;
    if (f > 1.0) { /* spatial_aggregation_3d :172-196 */
      for (i = 0; i < (int32_t)n1; i++) g[i] = 0.0;
      for (j = 1; j <= nc2; j++)
        for (i = 1; i <= nr2; i++) {
          const int32_t jc = (int32_t)ceil((double)j / f), ic = (int32_t)ceil((double)i / f);
          if (!mask2[(size_t)(j - 1) * nr2 + (i - 1)]) continue;
          g[(size_t)(jc - 1) * nr1 + (ic - 1)] = g[(size_t)(jc - 1) * nr1 + (ic - 1)] + d2[(size_t)(j - 1) * nr2 + (i - 1)];
        }
      for (i = 0; i < (int32_t)n1; i++) g[i] = mask1[i] ? g[i] / (double)cnt[i] : -9999.0;
    } else if (f < 1.0) { /* spatial_disaggregation_3d :353-372, cellFactor = cellsize2 / cellsize1 */
      const double fd = cellsize2 / cellsize1;
      for (i = 0; i < (int32_t)n1; i++) g[i] = -9999.0;
      for (j = 1; j <= nc1; j++)
        for (i = 1; i <= nr1; i++) {
          const int32_t jc = (int32_t)ceil((double)j / fd), ic = (int32_t)ceil((double)i / fd);
          if (!mask2[(size_t)(jc - 1) * nr2 + (ic - 1)]) continue;
          g[(size_t)(j - 1) * nr1 + (i - 1)] = d2[(size_t)(jc - 1) * nr2 + (ic - 1)];
        }
    } else {
      for (i = 0; i < (int32_t)n1; i++) g[i] = d2[i];
    }
    if (out_grid)
      for (i = 0; i < (int32_t)n1; i++) out_grid[(size_t)t * n1 + i] = g[i];
    if (out_packed)
      for (i = 0; i < (int32_t)n1; i++)
        if (mask1[i]) out_packed[(size_t)t * ncell + k++] = g[i];
  }
  free(g);
  free(cnt);
  return ncell;
}

/* ---------------------------------------------------------------------------------------------
 * L11_routing_order, mRM/mo_mrm_net_startup.f90:765-842, in linear time.
 * The reference numbers the headwater links 1..nHead in ascending link index (:779-803) and then
 * sweeps the remaining links in ascending index over and over; a link is numbered in the first
 * sweep in which all links entering its from-node carry a number already -- numbers given earlier
 * in the SAME sweep count (:812-838).  So a link i whose inflowing links are J gets numbered in
 *   pass(i) = max(1, max_{j in J} (pass(j) + (j > i)))       (headwater links: pass 0)
 * and its number is the rank of (pass(i), i).  One topological walk (links leave every node at most
 * once, so the link graph is a forest) and a counting sort by pass.
 * --------------------------------------------------------------------------------------------- */
int32_t orc_routing_order_linear(int32_t nNodes, int32_t nLinks, const int32_t *fromN, const int32_t *toN,
                                 int32_t *rOrder, int32_t *netPerm) {
  int32_t *out_link = (int32_t *)malloc(((size_t)nNodes + 1) * sizeof(int32_t));
  int32_t *waiting = (int32_t *)calloc((size_t)nLinks + 1, sizeof(int32_t));
  int32_t *pass = (int32_t *)calloc((size_t)nLinks + 1, sizeof(int32_t));
  int32_t *stack = (int32_t *)malloc(((size_t)nLinks + 1) * sizeof(int32_t));
  int32_t rc = 0, top = 0, seen = 0, maxpass = 0;
  if (!out_link || !waiting || !pass || !stack) { rc = 2; goto done; }
  for (int32_t nd = 0; nd <= nNodes; ++nd) out_link[nd] = -1;
  for (int32_t i = 0; i < nLinks; ++i) out_link[fromN[i]] = i;
  for (int32_t j = 0; j < nLinks; ++j) {        /* links entering the from-node of link d */
    const int32_t d = out_link[toN[j]];
    if (d >= 0 && d != j) waiting[d] += 1;
  }
  for (int32_t i = 0; i < nLinks; ++i)
    if (waiting[i] == 0) stack[top++] = i;      /* headwater links */
  while (top > 0) {
    const int32_t j = stack[--top];
    ++seen;
    if (pass[j] > maxpass) maxpass = pass[j];
    const int32_t d = out_link[toN[j]];
    if (d < 0 || d == j) continue;
    int32_t need = pass[j] + (j > d ? 1 : 0);
    if (need < 1) need = 1;
    if (need > pass[d]) pass[d] = need;
    if (--waiting[d] == 0) stack[top++] = d;
  }
  if (seen != nLinks) { rc = 1; goto done; }    /* a cycle */
  {
    int32_t *first = (int32_t *)calloc((size_t)maxpass + 2, sizeof(int32_t));
    if (!first) { rc = 2; goto done; }
    for (int32_t i = 0; i < nLinks; ++i) first[pass[i] + 1] += 1;
    for (int32_t s = 0; s <= maxpass; ++s) first[s + 1] += first[s];
    for (int32_t i = 0; i < nLinks; ++i) {      /* ascending i inside a pass */
      const int32_t r = first[pass[i]]++;
      rOrder[i] = r + 1;
      netPerm[r] = i + 1;
    }
    free(first);
  }
done:
  free(out_link); free(waiting); free(pass); free(stack);
  return rc;
}
