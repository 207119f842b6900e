/*
 * mpr_oracle.h -- declarations of the MPR part of the CPU oracle (test infrastructure only,
 * see mhm_oracle.h for the rules).  Struct members are one per line for tests/orc.py.
 */
#ifndef MPR_ORACLE_H
#define MPR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* L0 grid + L0 -> L1 remap (common/mo_common_types.F90:52-88): mask0 is Fortran
 * (nrows0, ncols0) as int32 0/1; bounds are 1-based inclusive */
typedef struct orc_l0_grid {
  int32_t nrows0;
  int32_t ncols0;
  int32_t nL1;
  const int32_t *mask0;
  const int32_t *upper;
  const int32_t *lower;
  const int32_t *left;
  const int32_t *right;
  const int32_t *nsub;
} orc_l0_grid;

/* inputs of mpr (MPR/mo_multi_param_reg.f90:67-75) for iFlag_soilDB = 0 */
typedef struct orc_mpr_in {
  int32_t nrows0;
  int32_t ncols0;
  int32_t nL0;
  int32_t nL1;
  int32_t nLC;
  int32_t nLAI;
  int32_t nH;
  int32_t nSoil;
  int32_t maxHor;
  int32_t nGeo;
  int32_t nProc;
  int32_t nParam;
  int32_t max_LCover;
  const int32_t *mask0;
  const int32_t *upper;
  const int32_t *lower;
  const int32_t *left;
  const int32_t *right;
  const int32_t *nsub;
  const int32_t *geoUnit0;
  const int32_t *soilId0;
  const int32_t *LCover0;
  const double *Asp0;
  const double *slope_emp0;
  const double *y0;
  const double *LAI0;
  const int32_t *is_present;
  const int32_t *nHorizons;
  const int32_t *nTillHorizons;
  const double *sand;
  const double *clay;
  const double *DbM;
  const double *Wd;
  const double *RZdepth;
  const double *HorizonDepth;
  const int32_t *GeoUnitList;
  const int32_t *GeoUnitKar;
  double fracSealed_CityArea;
  const int32_t *processMatrix;
  const double *param;
  int32_t lastSoilId0; /* > 0: soil id whose root zone depth the last horizon's root fractions use (a shard of a
                          domain); 0: the last L0 cell's, as the reference's loops leave it */
} orc_mpr_in;

/* L1 effective parameters, Fortran (nL1, dim2, dim3), see include/mhm_cuda.h mhm_param_id */
typedef struct orc_mpr_out {
  double *fSealed;
  double *alpha;
  double *degDayInc;
  double *degDayMax;
  double *degDayNoPre;
  double *fAsp;
  double *HarSamCoeff;
  double *PrieTayAlpha;
  double *aeroResist;
  double *surfResist;
  double *fRoots;
  double *kFastFlow;
  double *kSlowFlow;
  double *kBaseFlow;
  double *kPerco;
  double *karstLoss;
  double *soilMoistFC;
  double *soilMoistSat;
  double *soilMoistExp;
  double *jarvis_thresh_c1;
  double *tempThresh;
  double *unsatThresh;
  double *sealedThresh;
  double *wiltingPoint;
  double *maxInter;
  double *petLAIcorFactor;
} orc_mpr_out;

/* soil-class tables of mpr_sm, Fortran (nSoil, maxHor[, 3]) */
typedef struct orc_soil_table {
  int32_t nSoil;
  int32_t maxHor;
  double *thetaS_till;
  double *thetaFC_till;
  double *thetaPW_till;
  double *Ks;
  double *Db;
  double *thetaS;
  double *thetaFC;
  double *thetaPW;
} orc_soil_table;

void orc_calculate_grid_properties(int32_t nrowsIn, int32_t ncolsIn, double xllIn, double yllIn,
                                   double cellsizeIn, double aiming, int32_t *nrowsOut,
                                   int32_t *ncolsOut, double *xllOut, double *yllOut,
                                   double *cellsizeOut);
int32_t orc_init_lowres_level(int32_t nrows0, int32_t ncols0, const int32_t *mask0,
                              const double *cellArea0, double cellsize0, double target_resolution,
                              int32_t nrows1, int32_t ncols1, int32_t *mask1, int32_t *cellCoor,
                              double *cellArea1, int32_t *upper, int32_t *lower, int32_t *left,
                              int32_t *right, int32_t *n_subcells, int32_t *id_on_highres);
void orc_upscale_arithmetic_mean(const orc_l0_grid *g, const double *x0, double *out);
void orc_upscale_harmonic_mean(const orc_l0_grid *g, const double *x0, double *out);
void orc_upscale_geometric_mean(const orc_l0_grid *g, double nodata, const double *x0, double *out);
void orc_L0_fractionalCover_in_Lx(const orc_l0_grid *g, const int32_t *dataIn0, int32_t classId,
                                  double *out);
double orc_hydro_cond(const double *param4, double sand, double clay);
void orc_Genuchten(double *thetaS, double *n, double *alpha, const double *param6, double sand,
                   double clay, double Db);
double orc_field_cap(double Ks, double thetaS, double n);
double orc_PWP(double n, double alpha, double thetaS);
void orc_mpr_sm_table(const orc_mpr_in *in, const double *param13, orc_soil_table *t);
int32_t orc_mpr(const orc_mpr_in *in, orc_mpr_out *out);

#ifdef __cplusplus
}
#endif
#endif
