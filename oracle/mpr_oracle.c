/*
 * mpr_oracle.c -- CPU restatement of the MPR (multiscale parameter regionalisation) part of
 * the hot path: L0 -> L1 grid maps, upscaling operators and the transfer functions that fill
 * the L1 effective parameters.  TEST INFRASTRUCTURE ONLY (see mhm_oracle.h).
 *
 * Restates (paths relative to /root/reference/src):
 *   common/mo_grid.f90:58-183 (init_lowres_level), :489-556 (calculate_grid_properties)
 *   MPR/mo_upscaling_operators.f90:152-227, 266-329, 369-432, 469-535
 *   MPR/mo_multi_param_reg.f90:67-654 (mpr), :689-727, :800-852, :883-894, :944-1036,
 *       :1076-1159, :1203-1301
 *   MPR/mo_mpr_soilmoist.f90:100-455 (mpr_sm, iFlag_soilDB = 0), :489-772 (PWP, field_cap,
 *       Genuchten, hydro_cond)
 *   MPR/mo_mpr_smhorizons.f90:120-741 (mpr_SMhorizons, iFlag_soilDB = 0)
 *   MPR/mo_mpr_runoff.f90:74-194, MPR/mo_mpr_pet.f90:80-475, MPR/mo_mpr_constants.f90:21-74
 * Neutron (COSMIC) parameters are out of scope (processCase(10) = 0 in every BASELINE config).
 * Parity: no upstream unit test pins these routines except test_grid.pf (init_lowres_level);
 * see DESIGN.md.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mhm_oracle.h"
#include "mpr_oracle.h"

/* MPR/mo_mpr_constants.f90:34-74 */
static const double BulkDens_OrgMatter = 0.224, field_cap_c1 = -0.60, field_cap_c2 = 2.0;
static const double vG_sandtresh = 66.5;
static const double vG[19] = {0,      1.392, 0.418,  -0.024, 1.212, -0.704, -0.648, 0.023, 0.044, 3.168,
                              -2.562, 7.0E-9, 4.004, 3.750,  -0.016, -4.197, 0.013, 0.076, 0.276};
static const double Ks_c = 10.0, PWP_c = 1.0, PWP_matPot_ThetaR = 15000.0;
static const double WindMeasHeight = 10.0, karman = 0.41;
static const double LAI_factor_surfResi = 0.3, LAI_offset_surfResi = 1.2, max_surfResist = 250.0;

/* ---------------------------------------------------------------- grids */

/* common/mo_grid.f90:489-556 */
void orc_calculate_grid_properties(int32_t nrowsIn, int32_t ncolsIn, double xllIn, double yllIn,
                                   double cellsizeIn, double aiming, int32_t *nrowsOut,
                                   int32_t *ncolsOut, double *xllOut, double *yllOut,
                                   double *cellsizeOut) {
  double cellFactor = aiming / cellsizeIn, rounded = round(cellFactor);
  int32_t rounded_int = (int32_t)lround(cellFactor);
  *cellsizeOut = aiming;
  *ncolsOut = (int32_t)lround((double)ncolsIn / cellFactor);
  *nrowsOut = (int32_t)lround((double)nrowsIn / cellFactor);
  if (*ncolsOut * rounded_int < ncolsIn) *ncolsOut += 1;
  if (*nrowsOut * rounded_int < nrowsIn) *nrowsOut += 1;
  *xllOut = xllIn + (double)ncolsIn * aiming / rounded - (double)(*ncolsOut) * (*cellsizeOut);
  *yllOut = yllIn + (double)nrowsIn * aiming / rounded - (double)(*nrowsOut) * (*cellsizeOut);
}

/* common/mo_grid.f90:97-175.  mask arrays are Fortran (nrows, ncols) int32 0/1.
 * Outputs sized by the caller: mask1 (nrows1*ncols1); per-L1-cell arrays (nrows1*ncols1 max);
 * id_on_highres (nrows0*ncols0).  Returns nCells1. */
int32_t orc_init_lowres_level(int32_t nrows0, int32_t ncols0, const int32_t *mask0,
                              const double *cellArea0 /* packed, may be NULL */,
                              double cellsize0, double target_resolution, int32_t nrows1,
                              int32_t ncols1, int32_t *mask1, int32_t *cellCoor /* (nCells1,2) */,
                              double *cellArea1, int32_t *upper, int32_t *lower, int32_t *left,
                              int32_t *right, int32_t *n_subcells, int32_t *id_on_highres) {
  double cellFactor = round(target_resolution / cellsize0); /* anint(lowres%cellsize / highres%cellsize) */
  int32_t i, j, ic, jc, k, nCells1 = 0, cf;
  double *area2d = 0;
  for (k = 0; k < nrows1 * ncols1; k++) mask1[k] = 0;
  for (j = 1; j <= ncols0; j++) {
    jc = (int32_t)ceil((double)j / cellFactor);
    for (i = 1; i <= nrows0; i++) {
      if (!mask0[(size_t)(j - 1) * nrows0 + (i - 1)]) continue;
      ic = (int32_t)ceil((double)i / cellFactor);
      mask1[(size_t)(jc - 1) * nrows1 + (ic - 1)] = 1;
    }
  }
  for (k = 0; k < nrows1 * ncols1; k++) nCells1 += mask1[k] != 0;
  if (!upper) return nCells1;
  cellFactor = target_resolution / cellsize0; /* :113, not rounded here */
  cf = (int32_t)lround(cellFactor);
  if (cellArea0) { /* unpack(CellArea, mask, nodata) */
    size_t e, kk = 0;
    area2d = (double *)malloc(sizeof(double) * (size_t)nrows0 * ncols0);
    for (e = 0; e < (size_t)nrows0 * ncols0; e++) area2d[e] = mask0[e] ? cellArea0[kk++] : ORC_NODATA_DP;
  }
  for (k = 0; k < nrows0 * ncols0; k++) id_on_highres[k] = ORC_NODATA_I4;
  k = 0;
  for (jc = 1; jc <= ncols1; jc++) {
    for (ic = 1; ic <= nrows1; ic++) {
      int32_t iup, idown, jl, jr, cnt = 0;
      double a = 0.0;
      if (!mask1[(size_t)(jc - 1) * nrows1 + (ic - 1)]) continue;
      k++;
      cellCoor[k - 1] = ic;
      cellCoor[nCells1 + k - 1] = jc;
      iup = (ic - 1) * cf + 1;
      idown = ic * cf;
      jl = (jc - 1) * cf + 1;
      jr = jc * cf;
      if (iup < 1) iup = 1;
      if (idown > nrows0) idown = nrows0;
      if (jl < 1) jl = 1;
      if (jr > ncols0) jr = ncols0;
      upper[k - 1] = iup;
      lower[k - 1] = idown;
      left[k - 1] = jl;
      right[k - 1] = jr;
      for (j = jl; j <= jr; j++)
        for (i = iup; i <= idown; i++) {
          size_t e = (size_t)(j - 1) * nrows0 + (i - 1);
          if (mask0[e]) {
            cnt++;
            if (area2d) a = a + area2d[e];
          }
          id_on_highres[e] = k;
        }
      if (cellArea1) cellArea1[k - 1] = a;
      n_subcells[k - 1] = cnt;
    }
  }
  free(area2d);
  return nCells1;
}

/* ---------------------------------------------------------------- upscaling operators */

static void unpack_ids(const orc_l0_grid *g, int32_t *cell_of) {
  size_t e, k = 0;
  for (e = 0; e < (size_t)g->nrows0 * g->ncols0; e++) cell_of[e] = g->mask0[e] ? (int32_t)k++ : -1;
}

/* MPR/mo_upscaling_operators.f90:266-329; sum over the rectangle in array element order */
void orc_upscale_arithmetic_mean(const orc_l0_grid *g, const double *x0, double *out) {
  int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)g->nrows0 * g->ncols0);
  int32_t kk, i, j;
  unpack_ids(g, cell_of);
  for (kk = 0; kk < g->nL1; kk++) {
    double s = 0.0;
    for (j = g->left[kk] - 1; j <= g->right[kk] - 1; j++)
      for (i = g->upper[kk] - 1; i <= g->lower[kk] - 1; i++) {
        int32_t c = cell_of[(size_t)j * g->nrows0 + i];
        if (c >= 0) s = s + x0[c];
      }
    out[kk] = s / (double)g->nsub[kk];
  }
  free(cell_of);
}

/* :369-432 */
void orc_upscale_harmonic_mean(const orc_l0_grid *g, const double *x0, double *out) {
  int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)g->nrows0 * g->ncols0);
  int32_t kk, i, j;
  unpack_ids(g, cell_of);
  for (kk = 0; kk < g->nL1; kk++) {
    double s = 0.0;
    for (j = g->left[kk] - 1; j <= g->right[kk] - 1; j++)
      for (i = g->upper[kk] - 1; i <= g->lower[kk] - 1; i++) {
        int32_t c = cell_of[(size_t)j * g->nrows0 + i];
        if (c >= 0) s = s + 1.0 / x0[c];
      }
    out[kk] = (double)g->nsub[kk] / s;
  }
  free(cell_of);
}

static int ne_eps(double a, double b) { /* FORCES mo_utils::ne */
  return (ORC_EPS_DP * fabs(b) - fabs(a - b)) < 0.0;
}

/* :469-535 */
void orc_upscale_geometric_mean(const orc_l0_grid *g, double nodata, const double *x0, double *out) {
  int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)g->nrows0 * g->ncols0);
  int32_t kk, i, j;
  unpack_ids(g, cell_of);
  for (kk = 0; kk < g->nL1; kk++) {
    double p = 1.0;
    int32_t n = 0;
    for (j = g->left[kk] - 1; j <= g->right[kk] - 1; j++)
      for (i = g->upper[kk] - 1; i <= g->lower[kk] - 1; i++) {
        int32_t c = cell_of[(size_t)j * g->nrows0 + i];
        if (c >= 0 && ne_eps(x0[c], nodata)) {
          p = p * x0[c];
          n++;
        }
      }
    if (n == 0) out[kk] = 1.0;
    else if (ne_eps(p, 0.0)) out[kk] = pow(p, 1.0 / (double)n);
    else out[kk] = 0.0;
  }
  free(cell_of);
}

/* :152-227 */
void orc_L0_fractionalCover_in_Lx(const orc_l0_grid *g, const int32_t *dataIn0, int32_t classId,
                                  double *out) {
  int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)g->nrows0 * g->ncols0);
  int32_t kk, i, j;
  unpack_ids(g, cell_of);
  for (kk = 0; kk < g->nL1; kk++) {
    int32_t cnt = 0;
    for (j = g->left[kk] - 1; j <= g->right[kk] - 1; j++)
      for (i = g->upper[kk] - 1; i <= g->lower[kk] - 1; i++) {
        int32_t c = cell_of[(size_t)j * g->nrows0 + i];
        if (c >= 0 && dataIn0[c] == classId) cnt++;
      }
    out[kk] = (double)cnt / (double)g->nsub[kk];
  }
  free(cell_of);
}

/* ---------------------------------------------------------------- pedotransfer functions */

/* MPR/mo_mpr_soilmoist.f90:736-772 */
double orc_hydro_cond(const double *param4, double sand, double clay) {
  double x = param4[0] + param4[1] * sand - param4[2] * clay;
  double Ks = param4[3] * exp(x * log(Ks_c));
  if (Ks < 1.10) Ks = 1.10;
  return Ks;
}

/* :626-698 */
void orc_Genuchten(double *thetaS, double *n, double *alpha, const double *param6, double sand,
                   double clay, double Db) {
  double x;
  if (sand < vG_sandtresh) {
    *thetaS = param6[0] + param6[1] * clay + param6[2] * Db;
    *n = vG[1] - vG[2] * pow(sand, vG[3]) + vG[4] * pow(clay, vG[5]);
    x = vG[6] + vG[7] * sand + vG[8] * clay - vG[9] * Db;
  } else {
    *thetaS = param6[3] + param6[4] * clay + param6[5] * Db;
    *n = vG[10] + vG[11] * pow(sand, vG[12]) + vG[13] * pow(clay, vG[14]);
    x = vG[15] + vG[16] * sand + vG[17] * clay - vG[18] * Db;
  }
  *alpha = exp(x);
  if (*thetaS < 0.01) *thetaS = 0.01;
  if (*thetaS > 1.0) *thetaS = 1.0;
  if (*n < 1.01000) *n = 1.01000;
  if (*alpha < 0.00001) *alpha = 0.00001;
}

/* :560-584 */
double orc_field_cap(double Ks, double thetaS, double n) {
  double x = field_cap_c1 * (field_cap_c2 + log10(Ks));
  return thetaS * exp(x * log(n));
}

/* :489-520 */
double orc_PWP(double n, double alpha, double thetaS) {
  double m = PWP_c - (PWP_c / n);
  double x = PWP_c + exp(n * log(alpha * PWP_matPot_ThetaR));
  x = exp(m * log(x));
  if (x < 1.0) x = 1.0;
  return thetaS / x;
}

/* ---------------------------------------------------------------- soil-class table (mpr_sm, part 1) */

#define T3(a, s, j, L) (a)[((size_t)(L) * t->maxHor + (j)) * t->nSoil + (s)]
#define T2(a, s, j) (a)[(size_t)(j) * t->nSoil + (s)]

/* mo_mpr_soilmoist.f90:222-324 (iFlag_soilDB = 0).  Tables are (nSoil, maxHor[, 3]) Fortran
 * order; the reference stores the non-till tables shifted by minval(nTillHorizons) -- here they
 * are indexed by the data-base horizon itself, which is the same element. */
void orc_mpr_sm_table(const orc_mpr_in *in, const double *param13, orc_soil_table *t) {
  int32_t i, j, L, soil_case = in->processMatrix[2];
  double pOM_forest, pOM_imp = param13[1], pOM_perv = param13[2];
  size_t n3 = (size_t)in->nSoil * in->maxHor * 3, n2 = (size_t)in->nSoil * in->maxHor;
  t->nSoil = in->nSoil;
  t->maxHor = in->maxHor;
  pOM_forest = (soil_case == 1 || soil_case == 2) ? param13[2] + param13[0] : param13[0];
  memset(t->thetaS_till, 0, n3 * sizeof(double));
  memset(t->thetaFC_till, 0, n3 * sizeof(double));
  memset(t->thetaPW_till, 0, n3 * sizeof(double));
  memset(t->Ks, 0, n3 * sizeof(double));
  memset(t->Db, 0, n3 * sizeof(double));
  memset(t->thetaS, 0, n2 * sizeof(double));
  memset(t->thetaFC, 0, n2 * sizeof(double));
  memset(t->thetaPW, 0, n2 * sizeof(double));
  for (i = 0; i < in->nSoil; i++) {
    if (in->is_present[i] < 1) continue;
    for (j = 0; j < in->nHorizons[i]; j++) {
      double sand = T2(in->sand, i, j), clay = T2(in->clay, i, j), DbM = T2(in->DbM, i, j);
      double Ks_tmp = orc_hydro_cond(param13 + 9, sand, clay), n, alpha;
      for (L = 0; L < 3; L++) T3(t->Ks, i, j, L) = Ks_tmp;
      if (j + 1 <= in->nTillHorizons[i]) {
        for (L = 0; L < in->max_LCover; L++) {
          double pOM = L == 0 ? pOM_forest : (L == 1 ? pOM_imp : pOM_perv), pM = 100.0 - pOM;
          T3(t->Db, i, j, L) = 100.0 / ((pOM / BulkDens_OrgMatter) + (pM / DbM));
          T3(t->Ks, i, j, L) = T3(t->Ks, i, j, L) * (DbM / T3(t->Db, i, j, L));
          orc_Genuchten(&T3(t->thetaS_till, i, j, L), &n, &alpha, param13 + 3, sand, clay,
                        T3(t->Db, i, j, L));
          T3(t->thetaFC_till, i, j, L) =
              orc_field_cap(T3(t->Ks, i, j, L), T3(t->thetaS_till, i, j, L), n);
          T3(t->thetaPW_till, i, j, L) = orc_PWP(n, alpha, T3(t->thetaS_till, i, j, L));
        }
      } else {
        orc_Genuchten(&T2(t->thetaS, i, j), &n, &alpha, param13 + 3, sand, clay, DbM);
        T2(t->thetaFC, i, j) = orc_field_cap(T3(t->Ks, i, j, 0), T2(t->thetaS, i, j), n);
        T2(t->thetaPW, i, j) = orc_PWP(n, alpha, T2(t->thetaS, i, j));
      }
    }
  }
}

/* ---------------------------------------------------------------- mpr */

#define OUT3(a, n, dim2, j, y) ((a) + ((size_t)(y) * (dim2) + (j)) * (size_t)(n))

static void mpr_scene(const orc_mpr_in *in, orc_mpr_out *o, const orc_l0_grid *g, int32_t iLC,
                      double *w /* 8 x nL0 scratch */, double *w1 /* 4 x nL1 scratch */) {
  const int32_t n0 = in->nL0, n1 = in->nL1, nH = in->nH, nLAI = in->nLAI;
  const int32_t *pm = in->processMatrix, np = in->nProc;
  const int32_t *LC = in->LCover0 + (size_t)iLC * n0;
  const double *param = in->param;
  double *fForest = w1, *fPerm = w1 + n1;
  double *fSealed = OUT3(o->fSealed, n1, 1, 0, iLC);
  double *KsVar_H0 = w, *KsVar_V0 = w + n0, *SMs_FC0 = w + 2 * (size_t)n0, *tmp = w + 3 * (size_t)n0;
  orc_soil_table tab, *t = &tab;
  int32_t k, h, i, iStart, iEnd, iStart2, soil_case = pm[2], min_nTH;
  size_t n3 = (size_t)in->nSoil * in->maxHor * 3, n2 = (size_t)in->nSoil * in->maxHor;
  orc_mpr_in loc = *in;
#define PM(p, c) pm[((c)-1) * np + ((p)-1)]

  /* land-cover fractions, mo_multi_param_reg.f90:267-295 */
  orc_L0_fractionalCover_in_Lx(g, LC, 1, fForest);
  orc_L0_fractionalCover_in_Lx(g, LC, 2, fSealed);
  orc_L0_fractionalCover_in_Lx(g, LC, 3, fPerm);
  for (k = 0; k < n1; k++) {
    fSealed[k] = in->fracSealed_CityArea * fSealed[k];
    fPerm[k] = 1.0 - fSealed[k] - fForest[k];
  }
  /* snow_acc_melt_param :833-850 */
  {
    const double *p = param + (PM(2, 3) - PM(2, 2));
    double f_for = p[1], f_imp = p[1] + p[3] + p[2], f_per = p[1] + p[3];
    double m_for = p[1] + p[5], m_imp = p[1] + p[3] + p[2] + p[6], m_per = p[1] + p[3] + p[7];
    double *tt = OUT3(o->tempThresh, n1, 1, 0, iLC), *di = OUT3(o->degDayInc, n1, 1, 0, iLC);
    double *dn = OUT3(o->degDayNoPre, n1, 1, 0, iLC), *dm = OUT3(o->degDayMax, n1, 1, 0, iLC);
    for (k = 0; k < n1; k++) {
      tt[k] = p[0];
      di[k] = p[4];
      dn[k] = (f_for * fForest[k] + f_imp * fSealed[k] + f_per * fPerm[k]);
      dm[k] = (m_for * fForest[k] + m_imp * fSealed[k] + m_per * fPerm[k]);
    }
  }
  /* parameter windows of the soil-moisture process :357-388 */
  iStart = PM(3, 3) - PM(3, 2) + 1;
  switch (soil_case) {
    case 1: iEnd = PM(3, 3) - 4; iStart2 = PM(3, 3) - 4 + 1; break;
    case 2: iEnd = PM(3, 3) - 5; iStart2 = PM(3, 3) - 5 + 1; break;
    case 3: iEnd = PM(3, 3) - 9; iStart2 = PM(3, 3) - 8; break;
    default: iEnd = PM(3, 3) - 8; iStart2 = PM(3, 3) - 7; break;
  }
  (void)iEnd;
  if (soil_case == 2 || soil_case == 3)
    for (k = 0; k < n1; k++) o->jarvis_thresh_c1[k] = param[PM(3, 3) - 1];

  /* mpr_sm */
  loc.max_LCover = 0;
  for (k = 0; k < n0; k++)
    if (LC[k] > loc.max_LCover) loc.max_LCover = LC[k];
  tab.thetaS_till = (double *)malloc(n3 * sizeof(double));
  tab.thetaFC_till = (double *)malloc(n3 * sizeof(double));
  tab.thetaPW_till = (double *)malloc(n3 * sizeof(double));
  tab.Ks = (double *)malloc(n3 * sizeof(double));
  tab.Db = (double *)malloc(n3 * sizeof(double));
  tab.thetaS = (double *)malloc(n2 * sizeof(double));
  tab.thetaFC = (double *)malloc(n2 * sizeof(double));
  tab.thetaPW = (double *)malloc(n2 * sizeof(double));
  orc_mpr_sm_table(&loc, param + iStart - 1, t);
  { /* mo_mpr_soilmoist.f90:328-357 */
    const double p13 = param[iStart - 1 + 12];
    for (i = 0; i < n0; i++) {
      int32_t s = in->soilId0[i] - 1, L = LC[i] - 1, j;
      double kh = 0.0, kv = 0.0, fc = 0.0, tot = 0.0;
      for (j = 0; j < in->nHorizons[s]; j++) {
        if (j + 1 <= in->nTillHorizons[s]) {
          kh = kh + T3(t->thetaS_till, s, j, L) * T3(t->Ks, s, j, L);
          kv = kv + T3(t->thetaS_till, s, j, L) / T3(t->Ks, s, j, L);
          fc = fc + T3(t->thetaFC_till, s, j, L);
          tot = tot + T3(t->thetaS_till, s, j, L);
        } else {
          kh = kh + T2(t->thetaS, s, j) * T3(t->Ks, s, j, 0);
          kv = kv + T2(t->thetaS, s, j) / T3(t->Ks, s, j, 0);
          fc = fc + T2(t->thetaFC, s, j);
          tot = tot + T2(t->thetaS, s, j);
        }
      }
      SMs_FC0[i] = (tot - fc) / tot;
      KsVar_H0[i] = kh / tot / p13;
      KsVar_V0[i] = tot / kv / p13;
    }
  }

  /* mpr_SMhorizons, mo_mpr_smhorizons.f90:327-565 */
  {
    const double *p = param + iStart2 - 1;
    double c_forest = p[0], c_imp = p[1], c_perv, c_sand = 0.0, c_clay = 0.0, FCmin = 0.0, FCmax = 0.0;
    double *Bd0 = tmp, *SMs0 = tmp + n0, *FC0 = tmp + 2 * (size_t)n0, *PW0 = tmp + 3 * (size_t)n0,
           *fRoots0 = tmp + 4 * (size_t)n0;
    min_nTH = in->nTillHorizons[0];
    for (i = 1; i < in->nSoil; i++)
      if (in->nTillHorizons[i] < min_nTH) min_nTH = in->nTillHorizons[i];
    (void)min_nTH;
    if (soil_case == 1 || soil_case == 2) {
      c_perv = p[0] - p[2];
    } else {
      c_perv = p[2];
      c_sand = p[5] - p[4];
      c_clay = p[5];
      FCmin = p[6];
      FCmax = p[6] + p[7];
    }
    for (h = 0; h < nH; h++) {
      double dpth_f = 0.0, dpth_t = in->HorizonDepth[h];
      if (h > 0 && h < nH - 1) {
        dpth_f = in->HorizonDepth[h - 1];
        dpth_t = in->HorizonDepth[h];
      }
      for (k = 0; k < n0; k++) {
        int32_t L = LC[k] - 1, s = in->soilId0[k] - 1, j, nT = in->nTillHorizons[s],
                nHs = in->nHorizons[s];
        double a, b;
#define WD(s, h, j) in->Wd[((size_t)(j) * nH + (h)) * in->nSoil + (s)]
        a = 0.0;
        for (j = 0; j < nT; j++)
          if (WD(s, h, j) > 0.0) a = a + T3(t->Db, s, j, L) * WD(s, h, j);
        b = 0.0;
        for (j = nT; j < nHs; j++)
          if (WD(s, h, j) >= 0.0) b = b + T2(in->DbM, s, j) * WD(s, h, j);
        Bd0[k] = a + b;
        a = 0.0;
        for (j = 0; j < nT; j++)
          if (WD(s, h, j) > 0.0) a = a + T3(t->thetaS_till, s, j, L) * WD(s, h, j);
        b = 0.0;
        for (j = nT; j < nHs; j++)
          if (WD(s, h, j) > 0.0) b = b + T2(t->thetaS, s, j) * WD(s, h, j);
        SMs0[k] = a + b;
        a = 0.0;
        for (j = 0; j < nT; j++)
          if (WD(s, h, j) > 0.0) a = a + T3(t->thetaFC_till, s, j, L) * WD(s, h, j);
        b = 0.0;
        for (j = nT; j < nHs; j++)
          if (WD(s, h, j) > 0.0) b = b + T2(t->thetaFC, s, j) * WD(s, h, j);
        FC0[k] = a + b;
        a = 0.0;
        for (j = 0; j < nT; j++)
          if (WD(s, h, j) > 0.0) a = a + T3(t->thetaPW_till, s, j, L) * WD(s, h, j);
        b = 0.0;
        for (j = nT; j < nHs; j++)
          if (WD(s, h, j) > 0.0) b = b + T2(t->thetaPW, s, j) * WD(s, h, j);
        PW0[k] = a + b;
        if (h == nH - 1) { /* :424-427 -- stays set after the loop (see below) */
          dpth_f = in->HorizonDepth[nH - 2];
          dpth_t = in->RZdepth[s];
        }
        SMs0[k] = SMs0[k] * (dpth_t - dpth_f);
        FC0[k] = FC0[k] * (dpth_t - dpth_f);
        PW0[k] = PW0[k] * (dpth_t - dpth_f);
      }
      if (h == nH - 1 && in->lastSoilId0 > 0) dpth_t = in->RZdepth[in->lastSoilId0 - 1]; /* shard of a domain */
      /* root fractions :453-539.  NOTE the reference uses dpth_t/dpth_f as the previous loop
       * left them: for the last horizon that is RZdepth of the LAST L0 cell's soil type. */
      for (k = 0; k < n0; k++) {
        int32_t L = LC[k];
        double c;
        if (L == 1) c = c_forest;
        else if (L == 2) c = c_imp;
        else if (soil_case == 1 || soil_case == 2) c = c_perv;
        else {
          double FCnorm = (((FC0[k] / (dpth_t - dpth_f)) - FCmin) / (FCmax - FCmin));
          if (FCnorm < 0.0) FCnorm = 0.0;
          else if (FCnorm > 1.0) FCnorm = 1.0;
          c = (FCnorm * c_clay) + ((1 - FCnorm) * c_sand);
        }
        fRoots0[k] = (1.0 - pow(c, dpth_t * 0.1)) - (1.0 - pow(c, dpth_f * 0.1));
      }
      for (k = 0; k < n0; k++) Bd0[k] = Bd0[k] * p[3]; /* beta0 */
      orc_upscale_harmonic_mean(g, SMs0, OUT3(o->soilMoistSat, n1, nH, h, iLC));
      orc_upscale_harmonic_mean(g, Bd0, OUT3(o->soilMoistExp, n1, nH, h, iLC));
      orc_upscale_harmonic_mean(g, PW0, OUT3(o->wiltingPoint, n1, nH, h, iLC));
      orc_upscale_harmonic_mean(g, FC0, OUT3(o->soilMoistFC, n1, nH, h, iLC));
      orc_upscale_harmonic_mean(g, fRoots0, OUT3(o->fRoots, n1, nH, h, iLC));
    }
    /* :720-736 */
    for (h = 0; h < nH; h++) {
      double *S = OUT3(o->soilMoistSat, n1, nH, h, iLC), *F = OUT3(o->soilMoistFC, n1, nH, h, iLC);
      double *W = OUT3(o->wiltingPoint, n1, nH, h, iLC);
      for (k = 0; k < n1; k++) {
        if (F[k] > S[k]) F[k] = S[k] - 0.01 * S[k];
        if (W[k] > F[k]) W[k] = F[k] - 0.01 * F[k];
        if (S[k] < 0.0) S[k] = 0.0001;
        if (F[k] < 0.0) F[k] = 0.0001;
        if (W[k] < 0.0) W[k] = 0.0001;
      }
    }
    for (k = 0; k < n1; k++) {
      double tot = 0.0;
      for (h = 0; h < nH; h++) {
        double v = OUT3(o->fRoots, n1, nH, h, iLC)[k];
        if (v > 0.0) tot = tot + v;
      }
      for (h = 0; h < nH; h++) {
        double *v = OUT3(o->fRoots, n1, nH, h, iLC) + k;
        *v = tot > 0.0 ? *v / tot : 0.0;
      }
    }
  }

  /* PET-specific fields that depend on the land-cover scene :484-499 */
  if (PM(5, 1) == 3) { /* aerodynamical_resistance :1271-1297 */
    const double *p = param + (PM(5, 3) - PM(5, 2));
    double *maxLAI = tmp, *ch = tmp + n0, *ar0 = tmp + 2 * (size_t)n0;
    int32_t tt;
    for (k = 0; k < n0; k++) {
      double m = in->LAI0[k];
      for (tt = 1; tt < nLAI; tt++)
        if (in->LAI0[(size_t)tt * n0 + k] > m) m = in->LAI0[(size_t)tt * n0 + k];
      maxLAI[k] = m;
      ch[k] = ORC_NODATA_DP;
      if (LC[k] == 1) ch[k] = p[0];
      if (LC[k] == 2) ch[k] = p[1];
    }
    for (tt = 0; tt < nLAI; tt++) {
      for (k = 0; k < n0; k++) {
        double zm = WindMeasHeight, disp, zm0, zh0;
        if (LC[k] == 3) ch[k] = (p[2] * in->LAI0[(size_t)tt * n0 + k] / maxLAI[k]);
        if ((fabs(zm - ORC_NODATA_DP) > ORC_EPS_DP) && (zm < ch[k])) zm = ch[k] + zm;
        disp = p[3] * ch[k];
        zm0 = p[4] * ch[k];
        zh0 = p[5] * zm0;
        ar0[k] = log((zm - disp) / zm0) * log((zm - disp) / zh0) / pow(karman, 2.0);
      }
      orc_upscale_arithmetic_mean(g, ar0, OUT3(o->aeroResist, n1, nLAI, tt, iLC));
    }
  } else if (PM(5, 1) == -1) { /* pet_correctbyLAI, mo_mpr_pet.f90:140-161 */
    const double *p = param + (PM(5, 3) - PM(5, 2));
    int32_t tt;
    for (tt = 0; tt < nLAI; tt++) {
      for (k = 0; k < n0; k++) {
        double a = LC[k] == 1 ? p[0] : (LC[k] == 2 ? p[1] : p[2]);
        tmp[k] = a + (p[3] * (1.0 - exp(p[4] * in->LAI0[(size_t)tt * n0 + k])));
      }
      orc_upscale_harmonic_mean(g, tmp, OUT3(o->petLAIcorFactor, n1, nLAI, tt, iLC));
    }
  }

  /* mpr_runoff, mo_mpr_runoff.f90:142-191 */
  {
    const double *p = param + (PM(6, 3) - PM(6, 2));
    double *K0 = OUT3(o->kFastFlow, n1, 1, 0, iLC), *K1 = OUT3(o->kSlowFlow, n1, 1, 0, iLC);
    for (k = 0; k < n0; k++) tmp[k] = p[0] * SMs_FC0[k];
    orc_upscale_arithmetic_mean(g, tmp, o->unsatThresh);
    for (k = 0; k < n0; k++) {
      tmp[k] = p[1] * (2.0 - in->slope_emp0[k]);
      if (LC[k] == 1) tmp[k] = tmp[k] * p[2];
    }
    orc_upscale_arithmetic_mean(g, tmp, K0);
    for (k = 0; k < n1; k++)
      if (K0[k] < 1.0) K0[k] = 1.0;
    for (k = 0; k < n0; k++) tmp[k] = p[1] * (2.0 - in->slope_emp0[k]) + p[3] * (1.0 + KsVar_H0[k]);
    orc_upscale_arithmetic_mean(g, tmp, K1);
    for (k = 0; k < n1; k++)
      if (K1[k] < 2.0) K1[k] = 2.0;
    for (k = 0; k < n0; k++) tmp[k] = p[4] * (1.0 / KsVar_H0[k]) * (1.0 / (1.0 + SMs_FC0[k]));
    orc_upscale_arithmetic_mean(g, tmp, OUT3(o->alpha, n1, 1, 0, iLC));
    for (k = 0; k < n1; k++)
      if (K0[k] > K1[k]) K0[k] = K1[k];
  }
  /* karstic_layer, mo_multi_param_reg.f90:1009-1032 */
  {
    const double *p = param + (PM(7, 3) - PM(7, 2));
    double *Kp = OUT3(o->kPerco, n1, 1, 0, iLC), *fKar = w1 + 2 * (size_t)n1;
    for (k = 0; k < n0; k++) tmp[k] = p[0] * (1.0 + SMs_FC0[k]) / (1.0 + KsVar_V0[k]);
    orc_upscale_arithmetic_mean(g, tmp, Kp);
    for (k = 0; k < n1; k++) {
      if (Kp[k] < 2.0) Kp[k] = 2.0;
      fKar[k] = 0.0;
    }
    for (i = 0; i < in->nGeo; i++) { /* overwrites, does not accumulate (:1025-1029) */
      if (in->GeoUnitKar[i] == 0) continue;
      orc_L0_fractionalCover_in_Lx(g, in->geoUnit0, in->GeoUnitList[i], fKar);
    }
    for (k = 0; k < n1; k++) o->karstLoss[k] = 1.0 - (fKar[k] * p[1]);
  }
  free(tab.thetaS_till);
  free(tab.thetaFC_till);
  free(tab.thetaPW_till);
  free(tab.Ks);
  free(tab.Db);
  free(tab.thetaS);
  free(tab.thetaFC);
  free(tab.thetaPW);
}

/* MPR/mo_multi_param_reg.f90:67-654 */
int32_t orc_mpr(const orc_mpr_in *in, orc_mpr_out *o) {
  const int32_t n0 = in->nL0, n1 = in->nL1, nLAI = in->nLAI, nLC = in->nLC, np = in->nProc;
  const int32_t *pm = in->processMatrix;
  const double *param = in->param;
  orc_l0_grid g = {in->nrows0, in->ncols0, in->nL1, in->mask0, in->upper, in->lower,
                   in->left,   in->right,  in->nsub};
  double *w = (double *)malloc(sizeof(double) * 8 * (size_t)n0);
  double *w1 = (double *)malloc(sizeof(double) * 4 * (size_t)n1);
  int32_t iLC, k, tt;
  if (in->nH < 2) return 1; /* the reference indexes HorizonDepth(nH-1) */
  for (iLC = 0; iLC < nLC; iLC++) mpr_scene(in, o, &g, iLC, w, w1);

  for (k = 0; k < n1; k++) o->sealedThresh[k] = param[PM(4, 3) - 1]; /* iper_thres_runoff :893 */

  switch (PM(5, 1)) { /* :555-591 */
    case 0:
    case 1: { /* pet_correctbyASP, mo_mpr_pet.f90:245-270 */
      const double *p = param + (PM(5, 3) - PM(5, 2));
      double mx = p[0] + p[1];
      for (k = 0; k < n0; k++) {
        double asp = in->Asp0[k];
        double fN = asp < p[2] ? p[0] + (mx - p[0]) / p[2] * asp
                               : p[0] + (mx - p[0]) / (360. - p[2]) * (360. - asp);
        double fS = asp < p[2] ? p[0] + (mx - p[0]) / (360. - p[2]) * (360. - asp)
                               : p[0] + (mx - p[0]) / p[2] * asp;
        w[k] = in->y0[k] > 0.0 ? fN : fS;
      }
      orc_upscale_arithmetic_mean(&g, w, o->fAsp);
      if (PM(5, 1) == 1)
        for (k = 0; k < n1; k++) o->HarSamCoeff[k] = param[PM(5, 3) - 1];
      break;
    }
    case 2: { /* priestley_taylor_alpha :361-365 */
      const double *p = param + (PM(5, 3) - PM(5, 2));
      for (tt = 0; tt < nLAI; tt++) {
        for (k = 0; k < n0; k++) w[k] = p[0] + p[1] * in->LAI0[(size_t)tt * n0 + k];
        orc_upscale_arithmetic_mean(&g, w, OUT3(o->PrieTayAlpha, n1, nLAI, tt, 0));
      }
      break;
    }
    case 3: { /* bulksurface_resistance :460-471 */
      const double pr = param[PM(5, 3) - 1];
      for (tt = 0; tt < nLAI; tt++) {
        for (k = 0; k < n0; k++) {
          double lai = in->LAI0[(size_t)tt * n0 + k];
          double v = pr / (lai / (LAI_factor_surfResi * lai + LAI_offset_surfResi));
          w[k] = v > max_surfResist ? max_surfResist : v;
        }
        orc_upscale_arithmetic_mean(&g, w, OUT3(o->surfResist, n1, nLAI, tt, 0));
      }
      break;
    }
    default: break;
  }

  { /* baseflow_param :719-723 + :596-617 */
    const double *p = param + (PM(9, 3) - PM(9, 2));
    for (k = 0; k < n0; k++) {
      int32_t gi, best = 0, bd = abs(in->GeoUnitList[0] - in->geoUnit0[k]);
      for (gi = 1; gi < in->nGeo; gi++) { /* minloc: first minimum */
        int32_t dd = abs(in->GeoUnitList[gi] - in->geoUnit0[k]);
        if (dd < bd) {
          bd = dd;
          best = gi;
        }
      }
      w[k] = p[best];
    }
    orc_upscale_arithmetic_mean(&g, w, w1);
    for (iLC = 0; iLC < nLC; iLC++) {
      double *k2 = OUT3(o->kBaseFlow, n1, 1, 0, iLC), *k1 = OUT3(o->kSlowFlow, n1, 1, 0, iLC);
      for (k = 0; k < n1; k++) {
        k2[k] = w1[k];
        if (PM(7, 1) > 0 && k2[k] < k1[k]) k2[k] = k1[k];
      }
    }
  }

  { /* canopy_intercept_param :1142-1151 */
    const double gamma1 = param[PM(1, 3) - PM(1, 2)];
    for (tt = 0; tt < nLAI; tt++) {
      for (k = 0; k < n0; k++) w[k] = in->LAI0[(size_t)tt * n0 + k] * gamma1;
      orc_upscale_arithmetic_mean(&g, w, OUT3(o->maxInter, n1, nLAI, tt, 0));
    }
  }
  free(w);
  free(w1);
  return 0;
}
