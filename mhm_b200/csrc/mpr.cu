// mpr.cu -- MPR on the device: gamma -> L0 transfer functions -> L1 effective parameters.
//
// Restates (never copies), for iFlag_soilDB = 0 and without the neutron module:
//   MPR/mo_multi_param_reg.f90:67-654 (mpr), :689-727 baseflow_param, :800-852
//       snow_acc_melt_param, :883-894 iper_thres_runoff, :944-1036 karstic_layer, :1076-1159
//       canopy_intercept_param, :1203-1301 aerodynamical_resistance
//   MPR/mo_mpr_soilmoist.f90:100-455 (mpr_sm), :489-772 (PWP, field_cap, Genuchten, hydro_cond)
//   MPR/mo_mpr_smhorizons.f90:120-741 (mpr_SMhorizons)
//   MPR/mo_mpr_runoff.f90:74-194, MPR/mo_mpr_pet.f90:80-475, MPR/mo_mpr_constants.f90:21-74
//   common/mo_grid.f90:58-183 (init_lowres_level, host helper)
//
// Split of work: the soil-class table (nSoilTypes x horizons x 3 land-cover classes, a few
// thousand entries with pow/exp/log/log10) is evaluated on the host; every per-L0-cell field
// and every upscaling runs on the device and writes straight into the domain's parameter
// arrays.  Compiled with -fmad=false so that the per-cell sums round like the reference's.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "context.h"
#include "upscale.h"

namespace mhm {

// MPR/mo_mpr_constants.f90:34-74
static const double kBulkDensOrgMatter = 0.224, kFieldCapC1 = -0.60, kFieldCapC2 = 2.0;
static const double kVGsand = 66.5;
static const double kVG[19] = {0,      1.392,  0.418, -0.024, 1.212,  -0.704, -0.648, 0.023, 0.044, 3.168,
                               -2.562, 7.0E-9, 4.004, 3.750,  -0.016, -4.197, 0.013,  0.076, 0.276};
static const double kKsC = 10.0, kPwpC = 1.0, kPwpMatPot = 15000.0;
constexpr double kWindMeasHeight = 10.0, kKarman = 0.41;
constexpr double kLaiFactorSurfResi = 0.3, kLaiOffsetSurfResi = 1.2, kMaxSurfResist = 250.0;
constexpr double kNodata = -9999.0, kEpsDp = 2.220446049250313e-16;

struct MprState {
  mpr_l0_grid* grid = nullptr;
  int64_t nL0 = 0;
  int32_t nL1 = 0, nLC = 0, nLAI = 0, nH = 0;
  // L0 fields, device
  int32_t *geoUnit0 = nullptr, *soilId0 = nullptr, *LCover0 = nullptr;  // LCover0 [nLC][nL0]
  double *Asp0 = nullptr, *slope0 = nullptr, *y0 = nullptr, *LAI0 = nullptr;  // LAI0 [nLAI][nL0]
  int32_t last_soil = 0;                 // soilId0 of the last L0 cell (root-depth quirk)
  std::vector<int32_t> lc_max;           // maxval(LCover0(:, scene))
  // soil data base, host + device
  bool has_db = false;
  int32_t nSoil = 0, maxHor = 0;
  std::vector<int32_t> is_present, nHorizons, nTill, geoList, geoKar;
  std::vector<double> sand, clay, DbM, RZdepth, horizonDepth;
  double fracSealedCity = 1.0;
  int32_t *d_nHorizons = nullptr, *d_nTill = nullptr, *d_geoList = nullptr;
  double *d_DbM = nullptr, *d_Wd = nullptr, *d_RZdepth = nullptr;
  // device tables [3][maxHor][nSoil] and [maxHor][nSoil]
  double *t_thetaS_till = nullptr, *t_thetaFC_till = nullptr, *t_thetaPW_till = nullptr, *t_Ks = nullptr,
         *t_Db = nullptr, *t_thetaS = nullptr, *t_thetaFC = nullptr, *t_thetaPW = nullptr;
  // device scratch
  double* w0 = nullptr;  // [9][nL0]
  double* w1 = nullptr;  // [4][nL1]
  double* d_geoparam = nullptr;
};

void mpr_free(MprState* s) {
  if (!s) return;
  void* ptrs[] = {s->geoUnit0, s->soilId0, s->LCover0, s->Asp0, s->slope0, s->y0, s->LAI0, s->d_nHorizons,
                  s->d_nTill, s->d_geoList, s->d_DbM, s->d_Wd, s->d_RZdepth, s->t_thetaS_till,
                  s->t_thetaFC_till, s->t_thetaPW_till, s->t_Ks, s->t_Db, s->t_thetaS, s->t_thetaFC,
                  s->t_thetaPW, s->w0, s->w1, s->d_geoparam};
  for (void* p : ptrs) cudaFree(p);
  if (s->grid) mpr_cuda_grid_destroy(nullptr, s->grid);
  delete s;
}

template <class T>
static int to_device(T** dst, const T* src, size_t n, cudaStream_t st) {
  cudaFree(*dst);
  *dst = nullptr;
  MHM_CUDA_OK(cudaMalloc(dst, (n ? n : 1) * sizeof(T)));
  if (n) MHM_CUDA_OK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

// ------------------------------------------------------------------ host: soil-class table

// mo_mpr_soilmoist.f90:736-772
static double hydro_cond(const double* p, double sand, double clay) {
  const double x = p[0] + p[1] * sand - p[2] * clay;
  double Ks = p[3] * std::exp(x * std::log(kKsC));
  if (Ks < 1.10) Ks = 1.10;
  return Ks;
}
// :626-698
static void genuchten(double& thetaS, double& n, double& alpha, const double* p, double sand,
                      double clay, double Db) {
  double x;
  if (sand < kVGsand) {
    thetaS = p[0] + p[1] * clay + p[2] * Db;
    n = kVG[1] - kVG[2] * std::pow(sand, kVG[3]) + kVG[4] * std::pow(clay, kVG[5]);
    x = kVG[6] + kVG[7] * sand + kVG[8] * clay - kVG[9] * Db;
  } else {
    thetaS = p[3] + p[4] * clay + p[5] * Db;
    n = kVG[10] + kVG[11] * std::pow(sand, kVG[12]) + kVG[13] * std::pow(clay, kVG[14]);
    x = kVG[15] + kVG[16] * sand + kVG[17] * clay - kVG[18] * Db;
  }
  alpha = std::exp(x);
  if (thetaS < 0.01) thetaS = 0.01;
  if (thetaS > 1.0) thetaS = 1.0;
  if (n < 1.01000) n = 1.01000;
  if (alpha < 0.00001) alpha = 0.00001;
}
// :560-584
static double field_cap(double Ks, double thetaS, double n) {
  const double x = kFieldCapC1 * (kFieldCapC2 + std::log10(Ks));
  return thetaS * std::exp(x * std::log(n));
}
// :489-520
static double pwp(double n, double alpha, double thetaS) {
  const double m = kPwpC - (kPwpC / n);
  double x = kPwpC + std::exp(n * std::log(alpha * kPwpMatPot));
  x = std::exp(m * std::log(x));
  if (x < 1.0) x = 1.0;
  return thetaS / x;
}

struct HostTables {
  std::vector<double> thetaS_till, thetaFC_till, thetaPW_till, Ks, Db, thetaS, thetaFC, thetaPW;
};

// mo_mpr_soilmoist.f90:222-324; tables [L][j][s] / [j][s]
static void soil_table(const MprState* s, const double* p13, int soil_case, int max_lc, HostTables& t) {
  const size_t ns = (size_t)s->nSoil, n2 = ns * s->maxHor, n3 = n2 * 3;
  for (auto* v : {&t.thetaS_till, &t.thetaFC_till, &t.thetaPW_till, &t.Ks, &t.Db}) v->assign(n3, 0.0);
  for (auto* v : {&t.thetaS, &t.thetaFC, &t.thetaPW}) v->assign(n2, 0.0);
  const double pOM_forest = (soil_case == 1 || soil_case == 2) ? p13[2] + p13[0] : p13[0];
  const double pOM_imp = p13[1], pOM_perv = p13[2];
  auto i3 = [&](int i, int j, int L) { return ((size_t)L * s->maxHor + j) * ns + i; };
  auto i2 = [&](int i, int j) { return (size_t)j * ns + i; };
  for (int i = 0; i < s->nSoil; ++i) {
    if (s->is_present[(size_t)i] < 1) continue;
    for (int j = 0; j < s->nHorizons[(size_t)i]; ++j) {
      const double sand = s->sand[i2(i, j)], clay = s->clay[i2(i, j)], DbM = s->DbM[i2(i, j)];
      const double Ks_tmp = hydro_cond(p13 + 9, sand, clay);
      double n, alpha;
      for (int L = 0; L < 3; ++L) t.Ks[i3(i, j, L)] = Ks_tmp;
      if (j + 1 <= s->nTill[(size_t)i]) {
        for (int L = 0; L < max_lc && L < 3; ++L) {
          const double pOM = L == 0 ? pOM_forest : (L == 1 ? pOM_imp : pOM_perv), pM = 100.0 - pOM;
          t.Db[i3(i, j, L)] = 100.0 / ((pOM / kBulkDensOrgMatter) + (pM / DbM));
          t.Ks[i3(i, j, L)] = t.Ks[i3(i, j, L)] * (DbM / t.Db[i3(i, j, L)]);
          genuchten(t.thetaS_till[i3(i, j, L)], n, alpha, p13 + 3, sand, clay, t.Db[i3(i, j, L)]);
          t.thetaFC_till[i3(i, j, L)] = field_cap(t.Ks[i3(i, j, L)], t.thetaS_till[i3(i, j, L)], n);
          t.thetaPW_till[i3(i, j, L)] = pwp(n, alpha, t.thetaS_till[i3(i, j, L)]);
        }
      } else {
        genuchten(t.thetaS[i2(i, j)], n, alpha, p13 + 3, sand, clay, DbM);
        t.thetaFC[i2(i, j)] = field_cap(t.Ks[i3(i, j, 0)], t.thetaS[i2(i, j)], n);
        t.thetaPW[i2(i, j)] = pwp(n, alpha, t.thetaS[i2(i, j)]);
      }
    }
  }
}

// ------------------------------------------------------------------ device kernels

struct TabArgs {
  int32_t nSoil, maxHor, nH;
  const int32_t *nHorizons, *nTill;
  const double *thetaS_till, *thetaFC_till, *thetaPW_till, *Ks, *Db, *thetaS, *thetaFC, *thetaPW;
  const double *DbM, *Wd, *RZdepth;
};
#define TAB3(a, s, j, L) (a)[((size_t)(L) * t.maxHor + (j)) * t.nSoil + (s)]
#define TAB2(a, s, j) (a)[(size_t)(j) * t.nSoil + (s)]

// mo_mpr_soilmoist.f90:328-357: column integrals per L0 cell
__global__ void mpr_ksvar_kernel(int64_t n0, const int32_t* __restrict__ soilId0,
                                 const int32_t* __restrict__ LC, const TabArgs t, double p13,
                                 double* __restrict__ KsVar_H0, double* __restrict__ KsVar_V0,
                                 double* __restrict__ SMs_FC0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0) return;
  const int s = soilId0[i] - 1, L = LC[i] - 1;
  double kh = 0.0, kv = 0.0, fc = 0.0, tot = 0.0;
  const int nHs = t.nHorizons[s], nT = t.nTill[s];
  for (int j = 0; j < nHs; ++j) {
    if (j + 1 <= nT) {
      kh = kh + TAB3(t.thetaS_till, s, j, L) * TAB3(t.Ks, s, j, L);
      kv = kv + TAB3(t.thetaS_till, s, j, L) / TAB3(t.Ks, s, j, L);
      fc = fc + TAB3(t.thetaFC_till, s, j, L);
      tot = tot + TAB3(t.thetaS_till, s, j, L);
    } else {
      kh = kh + TAB2(t.thetaS, s, j) * TAB3(t.Ks, s, j, 0);
      kv = kv + TAB2(t.thetaS, s, j) / TAB3(t.Ks, s, j, 0);
      fc = fc + TAB2(t.thetaFC, s, j);
      tot = tot + TAB2(t.thetaS, s, j);
    }
  }
  SMs_FC0[i] = (tot - fc) / tot;
  KsVar_H0[i] = kh / tot / p13;
  KsVar_V0[i] = tot / kv / p13;
}

// mo_mpr_smhorizons.f90:386-448: depth-weighted horizon properties per L0 cell
__global__ void mpr_horizon_kernel(int64_t n0, int h, const int32_t* __restrict__ soilId0,
                                   const int32_t* __restrict__ LC, const TabArgs t, double dpth_f0,
                                   double dpth_t0, int last, double depth_last_f, double beta_par,
                                   double* __restrict__ beta0, double* __restrict__ SMs0,
                                   double* __restrict__ FC0, double* __restrict__ PW0) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n0) return;
  const int L = LC[k] - 1, s = soilId0[k] - 1, nT = t.nTill[s], nHs = t.nHorizons[s];
#define WDV(j) t.Wd[((size_t)(j) * t.nH + h) * t.nSoil + s]
  double a = 0.0, b = 0.0;
  for (int j = 0; j < nT; ++j)
    if (WDV(j) > 0.0) a = a + TAB3(t.Db, s, j, L) * WDV(j);
  for (int j = nT; j < nHs; ++j)
    if (WDV(j) >= 0.0) b = b + TAB2(t.DbM, s, j) * WDV(j);
  const double bd = a + b;
  a = 0.0, b = 0.0;
  for (int j = 0; j < nT; ++j)
    if (WDV(j) > 0.0) a = a + TAB3(t.thetaS_till, s, j, L) * WDV(j);
  for (int j = nT; j < nHs; ++j)
    if (WDV(j) > 0.0) b = b + TAB2(t.thetaS, s, j) * WDV(j);
  double sms = a + b;
  a = 0.0, b = 0.0;
  for (int j = 0; j < nT; ++j)
    if (WDV(j) > 0.0) a = a + TAB3(t.thetaFC_till, s, j, L) * WDV(j);
  for (int j = nT; j < nHs; ++j)
    if (WDV(j) > 0.0) b = b + TAB2(t.thetaFC, s, j) * WDV(j);
  double fc = a + b;
  a = 0.0, b = 0.0;
  for (int j = 0; j < nT; ++j)
    if (WDV(j) > 0.0) a = a + TAB3(t.thetaPW_till, s, j, L) * WDV(j);
  for (int j = nT; j < nHs; ++j)
    if (WDV(j) > 0.0) b = b + TAB2(t.thetaPW, s, j) * WDV(j);
  double pw = a + b;
  double dpth_f = dpth_f0, dpth_t = dpth_t0;
  if (last) {  // :424-427
    dpth_f = depth_last_f;
    dpth_t = t.RZdepth[s];
  }
  SMs0[k] = sms * (dpth_t - dpth_f);
  FC0[k] = fc * (dpth_t - dpth_f);
  PW0[k] = pw * (dpth_t - dpth_f);
  beta0[k] = bd * beta_par;  // :543
}

// :453-539 root fractions.  r_class: the three class constants evaluated on the host
// (soil cases 1, 2 and the forest / impervious classes of cases 3, 4)
__global__ void mpr_roots_kernel(int64_t n0, const int32_t* __restrict__ LC, int fc_dependent,
                                 double r_forest, double r_imp, double r_perv, double c_sand,
                                 double c_clay, double FCmin, double FCmax, double dpth_f, double dpth_t,
                                 const double* __restrict__ FC0, double* __restrict__ fRoots0) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n0) return;
  const int L = LC[k];
  double r;
  if (L == 1) {
    r = r_forest;
  } else if (L == 2) {
    r = r_imp;
  } else if (!fc_dependent) {
    r = r_perv;
  } else {
    double FCnorm = (((FC0[k] / (dpth_t - dpth_f)) - FCmin) / (FCmax - FCmin));
    if (FCnorm < 0.0) FCnorm = 0.0;
    else if (FCnorm > 1.0) FCnorm = 1.0;
    const double c = (FCnorm * c_clay) + ((1 - FCnorm) * c_sand);
    r = (1.0 - pow(c, dpth_t * 0.1)) - (1.0 - pow(c, dpth_f * 0.1));
  }
  fRoots0[k] = r;
}

enum {
  kL0UnsatThr, kL0K0, kL0K1, kL0Alpha, kL0Kp, kL0PetLai, kL0Fasp, kL0PtAlpha, kL0SurfRes, kL0K2,
  kL0MaxInter
};
struct L0Args {
  int64_t n0;
  int32_t op, nGeo;
  double p[6];
  const int32_t *LC, *geoUnit0, *geoList;
  const double *a, *b, *c;  // operand fields (meaning depends on op)
  const double* geoparam;
  double* out;
};
// element-wise L0 transfer functions
__global__ void mpr_l0_kernel(const L0Args q) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= q.n0) return;
  double v = 0.0;
  switch (q.op) {
    case kL0UnsatThr: v = q.p[0] * q.a[k]; break;  // mo_mpr_runoff.f90:142 (a = SMs_FC0)
    case kL0K0:                                    // :148-152 (a = slope_emp0)
      v = q.p[1] * (2.0 - q.a[k]);
      if (q.LC[k] == 1) v = v * q.p[2];
      break;
    case kL0K1: v = q.p[1] * (2.0 - q.a[k]) + q.p[3] * (1.0 + q.b[k]); break;           // :164 (b = KsVar_H0)
    case kL0Alpha: v = q.p[4] * (1.0 / q.b[k]) * (1.0 / (1.0 + q.a[k])); break;         // :178 (a = SMs_FC0)
    case kL0Kp: v = q.p[0] * (1.0 + q.a[k]) / (1.0 + q.b[k]); break;  // karstic_layer :1009 (b = KsVar_V0)
    case kL0PetLai: {                                                  // mo_mpr_pet.f90:140-152 (a = LAI)
      const int L = q.LC[k];
      const double a0 = L == 1 ? q.p[0] : (L == 2 ? q.p[1] : q.p[2]);
      v = a0 + (q.p[3] * (1.0 - exp(q.p[4] * q.a[k])));
      break;
    }
    case kL0Fasp: {  // pet_correctbyASP :245-270 (a = Asp0, b = latitude)
      const double asp = q.a[k], mx = q.p[0] + q.p[1];
      const double fN = asp < q.p[2] ? q.p[0] + (mx - q.p[0]) / q.p[2] * asp
                                     : q.p[0] + (mx - q.p[0]) / (360. - q.p[2]) * (360. - asp);
      const double fS = asp < q.p[2] ? q.p[0] + (mx - q.p[0]) / (360. - q.p[2]) * (360. - asp)
                                     : q.p[0] + (mx - q.p[0]) / q.p[2] * asp;
      v = q.b[k] > 0.0 ? fN : fS;
      break;
    }
    case kL0PtAlpha: v = q.p[0] + q.p[1] * q.a[k]; break;  // :363
    case kL0SurfRes: {                                     // :462-467
      const double lai = q.a[k];
      v = q.p[0] / (lai / (kLaiFactorSurfResi * lai + kLaiOffsetSurfResi));
      if (v > kMaxSurfResist) v = kMaxSurfResist;
      break;
    }
    case kL0K2: {  // baseflow_param, mo_multi_param_reg.f90:719-723: minloc(abs(list - unit))
      int best = 0, bd = abs(q.geoList[0] - q.geoUnit0[k]);
      for (int g = 1; g < q.nGeo; ++g) {
        const int dd = abs(q.geoList[g] - q.geoUnit0[k]);
        if (dd < bd) {
          bd = dd;
          best = g;
        }
      }
      v = q.geoparam[best];
      break;
    }
    case kL0MaxInter: v = q.a[k] * q.p[0]; break;  // :1143
  }
  q.out[k] = v;
}

// aerodynamical_resistance :1271-1297 for one LAI slice tt; canopy height of the pervious
// class is updated per slice exactly like the reference's running `canopy_height0`
__global__ void mpr_aero_kernel(int64_t n0, int tt, int nLAI, const int32_t* __restrict__ LC,
                                const double* __restrict__ LAI0, double p0, double p1, double p2,
                                double p3, double p4, double p5, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n0) return;
  double m = LAI0[k];
  for (int u = 1; u < nLAI; ++u) m = fmax(m, LAI0[(size_t)u * n0 + k]);
  double ch = kNodata;
  if (LC[k] == 1) ch = p0;
  if (LC[k] == 2) ch = p1;
  if (LC[k] == 3) ch = (p2 * LAI0[(size_t)tt * n0 + k] / m);
  double zm = kWindMeasHeight;
  if ((fabs(zm - kNodata) > kEpsDp) && (zm < ch)) zm = ch + zm;
  const double disp = p3 * ch, zm0 = p4 * ch, zh0 = p5 * zm0;
  out[k] = log((zm - disp) / zm0) * log((zm - disp) / zh0) / pow(kKarman, 2.0);
}

enum { kL1Fill, kL1ClampMin, kL1MinWith, kL1Snow, kL1Lc, kL1Karst, kL1MaxWith, kL1Copy };
struct L1Args {
  int32_t n1, op;
  double p[8];
  const double *a, *b, *c;
  double *x, *y, *z, *w;
};
__global__ void mpr_l1_kernel(const L1Args q) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= q.n1) return;
  switch (q.op) {
    case kL1Fill: q.x[k] = q.p[0]; break;
    case kL1ClampMin:
      if (q.x[k] < q.p[0]) q.x[k] = q.p[0];
      break;
    case kL1MinWith:  // x = merge(a, x, x > a)
      if (q.x[k] > q.a[k]) q.x[k] = q.a[k];
      break;
    case kL1MaxWith:  // x = merge(a, x, x < a)
      if (q.x[k] < q.a[k]) q.x[k] = q.a[k];
      break;
    case kL1Copy: q.x[k] = q.a[k]; break;
    case kL1Lc: {  // mo_multi_param_reg.f90:289-295: x = fSealed, y = fPerm, a = fForest
      const double fs = q.p[0] * q.x[k];
      q.x[k] = fs;
      q.y[k] = 1.0 - fs - q.a[k];
      break;
    }
    case kL1Snow: {  // :833-850: a = fForest, b = fSealed, c = fPerm
      q.x[k] = q.p[6];                                                      // tempThresh
      q.y[k] = q.p[7];                                                      // degDayInc
      q.z[k] = (q.p[0] * q.a[k] + q.p[1] * q.b[k] + q.p[2] * q.c[k]);       // degDayNoPre
      q.w[k] = (q.p[3] * q.a[k] + q.p[4] * q.b[k] + q.p[5] * q.c[k]);       // degDayMax
      break;
    }
    case kL1Karst: q.x[k] = 1.0 - (q.a[k] * q.p[0]); break;  // :1032
  }
}

// mo_mpr_smhorizons.f90:720-736 on the L1 arrays of one land-cover scene
__global__ void mpr_l1_soil_fix_kernel(int n1, int nH, double* __restrict__ SMs, double* __restrict__ FC,
                                       double* __restrict__ PW, double* __restrict__ fRoots) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n1) return;
  double tot = 0.0;
  for (int h = 0; h < nH; ++h) {
    const size_t o = (size_t)h * n1 + k;
    if (FC[o] > SMs[o]) FC[o] = SMs[o] - 0.01 * SMs[o];
    if (PW[o] > FC[o]) PW[o] = FC[o] - 0.01 * FC[o];
    if (SMs[o] < 0.0) SMs[o] = 0.0001;
    if (FC[o] < 0.0) FC[o] = 0.0001;
    if (PW[o] < 0.0) PW[o] = 0.0001;
    if (fRoots[o] > 0.0) tot = tot + fRoots[o];
  }
  for (int h = 0; h < nH; ++h) {
    const size_t o = (size_t)h * n1 + k;
    fRoots[o] = tot > 0.0 ? fRoots[o] / tot : 0.0;
  }
}

// ------------------------------------------------------------------ host driver

static int ensure_param(Domain* d, int id, int dim2, int dim3) {
  const size_t n = (size_t)d->cfg.nCells, rows = (size_t)dim2 * dim3;
  if (!d->P[id]) {
    MHM_CUDA_OK(cudaMalloc(&d->P[id], (size_t)d->cfg.nMembers * rows * n * sizeof(double)));
    MHM_CUDA_OK(cudaMemset(d->P[id], 0, (size_t)d->cfg.nMembers * rows * n * sizeof(double)));
    d->P_dim2[id] = dim2;
    d->P_dim3[id] = dim3;
  }
  MHM_REQUIRE(d->P_dim2[id] == dim2 && d->P_dim3[id] == dim3, "parameter %d has shape (:,%d,%d), MPR writes (:,%d,%d)",
              id, d->P_dim2[id], d->P_dim3[id], dim2, dim3);
  return 0;
}

static double* pslice(Domain* d, int id, int member, int j, int y) {
  const size_t n = (size_t)d->cfg.nCells;
  return d->P[id] + ((size_t)member * d->P_dim2[id] * d->P_dim3[id] + (size_t)y * d->P_dim2[id] + j) * n;
}

}  // namespace mhm

using namespace mhm;

extern "C" {

int mhm_grid_init_lowres_level(int32_t nrows0, int32_t ncols0, const int32_t* mask0,
                               const double* cellArea0, double cellsize0, double target_resolution,
                               int32_t* nrows1, int32_t* ncols1, int32_t* nCells1, int32_t* mask1,
                               int32_t* cellCoor, double* cellArea1, int32_t* upper, int32_t* lower,
                               int32_t* left, int32_t* right, int32_t* n_subcells,
                               int32_t* id_on_highres) {
  MHM_REQUIRE(mask0 && nrows1 && ncols1 && nCells1 && nrows0 > 0 && ncols0 > 0 && cellsize0 > 0,
              "init_lowres_level: bad arguments");
  // calculate_grid_properties, mo_grid.f90:534-554
  const double factor = target_resolution / cellsize0;
  const long ri = std::lround(factor);
  MHM_REQUIRE(std::fabs(std::round(factor) - factor) <= 1.e-7, "Two resolutions size do not confirm: %g %g",
              target_resolution, cellsize0);
  int32_t nc = (int32_t)std::lround((double)ncols0 / factor), nr = (int32_t)std::lround((double)nrows0 / factor);
  if (nc * ri < ncols0) ++nc;
  if (nr * ri < nrows0) ++nr;
  *nrows1 = nr;
  *ncols1 = nc;
  // low-resolution mask, :97-111
  std::vector<int32_t> m1((size_t)nr * nc, 0);
  const double cf_round = std::round(factor);
  for (int j = 1; j <= ncols0; ++j) {
    const int jc = (int)std::ceil((double)j / cf_round);
    for (int i = 1; i <= nrows0; ++i) {
      if (!mask0[(size_t)(j - 1) * nrows0 + (i - 1)]) continue;
      const int ic = (int)std::ceil((double)i / cf_round);
      m1[(size_t)(jc - 1) * nr + (ic - 1)] = 1;
    }
  }
  int32_t n1 = 0;
  for (int32_t v : m1) n1 += v != 0;
  *nCells1 = n1;
  if (mask1) std::memcpy(mask1, m1.data(), m1.size() * sizeof(int32_t));
  if (!upper) return 0;
  MHM_REQUIRE(lower && left && right && n_subcells, "init_lowres_level: bound arrays required");
  // remap, :113-175
  const int cf = (int)std::lround(factor);
  std::vector<double> area2d;
  if (cellArea0 && cellArea1) {
    area2d.resize((size_t)nrows0 * ncols0);
    size_t kk = 0;
    for (size_t e = 0; e < area2d.size(); ++e) area2d[e] = mask0[e] ? cellArea0[kk++] : kNodata;
  }
  if (id_on_highres)
    for (size_t e = 0; e < (size_t)nrows0 * ncols0; ++e) id_on_highres[e] = -9999;
  int k = 0;
  for (int jc = 1; jc <= nc; ++jc) {
    for (int ic = 1; ic <= nr; ++ic) {
      if (!m1[(size_t)(jc - 1) * nr + (ic - 1)]) continue;
      ++k;
      if (cellCoor) {
        cellCoor[k - 1] = ic;
        cellCoor[n1 + k - 1] = jc;
      }
      int iup = (ic - 1) * cf + 1, idown = ic * cf, jl = (jc - 1) * cf + 1, jr = jc * cf;
      iup = std::max(iup, 1);
      idown = std::min(idown, nrows0);
      jl = std::max(jl, 1);
      jr = std::min(jr, ncols0);
      upper[k - 1] = iup;
      lower[k - 1] = idown;
      left[k - 1] = jl;
      right[k - 1] = jr;
      int cnt = 0;
      double a = 0.0;
      for (int j = jl; j <= jr; ++j)
        for (int i = iup; i <= idown; ++i) {
          const size_t e = (size_t)(j - 1) * nrows0 + (i - 1);
          if (mask0[e]) {
            ++cnt;
            if (!area2d.empty()) a = a + area2d[e];
          }
          if (id_on_highres) id_on_highres[e] = k;
        }
      if (cellArea1) cellArea1[k - 1] = a;
      n_subcells[k - 1] = cnt;
    }
  }
  return 0;
}

int mpr_cuda_set_l0(mhm_cuda_context* ctx, int32_t iDomain, const mpr_l0_inputs* in) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(in && in->mask0 && in->geoUnit0 && in->soilId0 && in->LCover0 && in->Asp0 &&
                  in->slope_emp0 && in->y0 && in->gridded_LAI0,
              "mpr_set_l0: null input");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (d->mpr) mpr_free(d->mpr);
  d->mpr = nullptr;
  auto* s = new MprState();
  s->nL1 = d->cfg.nCells;
  s->nLC = d->cfg.nLCscenes;
  s->nLAI = d->cfg.nLAI;
  s->nH = d->cfg.nHorizons;
  if (int rc = mpr_cuda_grid_create(ctx, in->nrows0, in->ncols0, in->mask0, s->nL1, in->upper_bound,
                                    in->lower_bound, in->left_bound, in->right_bound, in->n_subcells,
                                    &s->grid)) {
    mpr_free(s);
    return rc;
  }
  s->nL0 = s->grid->nL0;
  const size_t n0 = (size_t)s->nL0;
  cudaStream_t st = ctx->stream;
  int rc = 0;
  rc |= to_device(&s->geoUnit0, in->geoUnit0, n0, st);
  rc |= to_device(&s->soilId0, in->soilId0, n0, st);
  rc |= to_device(&s->LCover0, in->LCover0, n0 * s->nLC, st);
  rc |= to_device(&s->Asp0, in->Asp0, n0, st);
  rc |= to_device(&s->slope0, in->slope_emp0, n0, st);
  rc |= to_device(&s->y0, in->y0, n0, st);
  rc |= to_device(&s->LAI0, in->gridded_LAI0, n0 * s->nLAI, st);
  if (!rc && cudaMalloc(&s->w0, 9 * n0 * sizeof(double)) != cudaSuccess) rc = 2;
  if (!rc && cudaMalloc(&s->w1, 4 * (size_t)s->nL1 * sizeof(double)) != cudaSuccess) rc = 2;
  cudaStreamSynchronize(st);
  if (rc) {
    set_error("mpr_set_l0: device allocation failed");
    mpr_free(s);
    return 2;
  }
  s->last_soil = in->lastSoilId0 > 0 ? in->lastSoilId0 : in->soilId0[n0 - 1];
  s->lc_max.assign((size_t)s->nLC, 0);
  for (int y = 0; y < s->nLC; ++y)
    for (size_t k = 0; k < n0; ++k) {
      const int32_t v = in->LCover0[(size_t)y * n0 + k];
      MHM_REQUIRE(v >= 1 && v <= 3, "mpr_set_l0: land-cover class %d outside 1..3", v);
      s->lc_max[(size_t)y] = std::max(s->lc_max[(size_t)y], v);
    }
  d->mpr = s;
  return 0;
}

int mpr_cuda_set_soildb(mhm_cuda_context* ctx, int32_t iDomain, const mpr_soil_db* db) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MprState* s = d->mpr;
  MHM_REQUIRE(s, "mpr_set_soildb: call mpr_cuda_set_l0 first");
  MHM_REQUIRE(db && db->nSoilTypes > 0 && db->maxHorizons > 0 && db->nGeoUnits > 0 && db->is_present &&
                  db->nHorizons && db->nTillHorizons && db->sand && db->clay && db->DbM && db->Wd &&
                  db->RZdepth && db->HorizonDepth_mHM && db->GeoUnitList && db->GeoUnitKar,
              "mpr_set_soildb: null input");
  MHM_REQUIRE(s->nH >= 2, "mpr: nSoilHorizons_mHM must be >= 2");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  s->nSoil = db->nSoilTypes;
  s->maxHor = db->maxHorizons;
  const size_t ns = (size_t)s->nSoil, n2 = ns * s->maxHor;
  s->is_present.assign(db->is_present, db->is_present + ns);
  s->nHorizons.assign(db->nHorizons, db->nHorizons + ns);
  s->nTill.assign(db->nTillHorizons, db->nTillHorizons + ns);
  s->sand.assign(db->sand, db->sand + n2);
  s->clay.assign(db->clay, db->clay + n2);
  s->DbM.assign(db->DbM, db->DbM + n2);
  s->RZdepth.assign(db->RZdepth, db->RZdepth + ns);
  s->horizonDepth.assign(db->HorizonDepth_mHM, db->HorizonDepth_mHM + s->nH);
  s->geoList.assign(db->GeoUnitList, db->GeoUnitList + db->nGeoUnits);
  s->geoKar.assign(db->GeoUnitKar, db->GeoUnitKar + db->nGeoUnits);
  s->fracSealedCity = db->fracSealed_CityArea;
  for (size_t i = 0; i < ns; ++i)
    MHM_REQUIRE(s->nHorizons[i] >= 1 && s->nHorizons[i] <= s->maxHor && s->nTill[i] >= 0 &&
                    s->nTill[i] <= s->nHorizons[i],
                "mpr_set_soildb: soil type %zu has nHorizons %d / nTillHorizons %d", i + 1,
                s->nHorizons[i], s->nTill[i]);
  cudaStream_t st = ctx->stream;
  int rc = 0;
  rc |= to_device(&s->d_nHorizons, s->nHorizons.data(), ns, st);
  rc |= to_device(&s->d_nTill, s->nTill.data(), ns, st);
  rc |= to_device(&s->d_geoList, s->geoList.data(), s->geoList.size(), st);
  rc |= to_device(&s->d_DbM, s->DbM.data(), n2, st);
  rc |= to_device(&s->d_Wd, db->Wd, n2 * s->nH, st);
  rc |= to_device(&s->d_RZdepth, s->RZdepth.data(), ns, st);
  for (double** p : {&s->t_thetaS_till, &s->t_thetaFC_till, &s->t_thetaPW_till, &s->t_Ks, &s->t_Db}) {
    cudaFree(*p);
    *p = nullptr;
    if (cudaMalloc(p, n2 * 3 * sizeof(double)) != cudaSuccess) rc = 2;
  }
  for (double** p : {&s->t_thetaS, &s->t_thetaFC, &s->t_thetaPW}) {
    cudaFree(*p);
    *p = nullptr;
    if (cudaMalloc(p, n2 * sizeof(double)) != cudaSuccess) rc = 2;
  }
  cudaFree(s->d_geoparam);
  s->d_geoparam = nullptr;
  if (cudaMalloc(&s->d_geoparam, s->geoList.size() * sizeof(double)) != cudaSuccess) rc = 2;
  MHM_CUDA_OK(cudaStreamSynchronize(st));
  MHM_REQUIRE(rc == 0, "mpr_set_soildb: device allocation failed");
  s->has_db = true;
  return 0;
}

int mpr_cuda_eval(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, const double* param,
                  int32_t nParam) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MprState* s = d->mpr;
  MHM_REQUIRE(s && s->has_db, "mpr_eval: mpr_cuda_set_l0 / mpr_cuda_set_soildb have not been called");
  MHM_REQUIRE(param && member >= 0 && member < d->cfg.nMembers, "mpr_eval: bad arguments");
  MHM_REQUIRE(d->cfg.nProcesses >= 9, "mpr_eval: processMatrix needs >= 9 rows");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const int np = d->cfg.nProcesses;
  const int32_t* pm = d->processMatrix.data();
  auto PM = [&](int p, int c) { return pm[(c - 1) * np + (p - 1)]; };
  for (int p = 1; p <= 9; ++p)
    MHM_REQUIRE(PM(p, 3) <= nParam && PM(p, 3) - PM(p, 2) >= 0, "mpr_eval: process %d needs parameters %d..%d of %d",
                p, PM(p, 3) - PM(p, 2) + 1, PM(p, 3), nParam);
  MHM_REQUIRE(PM(1, 1) == 1 && PM(2, 1) == 1 && PM(4, 1) == 1 && PM(6, 1) == 1 && PM(7, 1) == 1 && PM(9, 1) == 1,
              "mpr_eval: unsupported process case (interception/snow/direct runoff/interflow/percolation/baseflow must be 1)");
  const int soil_case = PM(3, 1), pet_case = PM(5, 1);
  MHM_REQUIRE(soil_case >= 1 && soil_case <= 4 && pet_case >= -1 && pet_case <= 3, "mpr_eval: bad soil/PET case");
  const int n1 = s->nL1, nH = s->nH, nLAI = s->nLAI, nLC = s->nLC;
  const int64_t n0 = s->nL0;
  cudaStream_t st = ctx->stream;
  const unsigned g0 = (unsigned)((n0 + 255) / 256), g1 = (unsigned)((n1 + 127) / 128);

  // parameter arrays this evaluation writes
  int rc = 0;
  for (int id : {MHM_P_FSEALED, MHM_P_ALPHA, MHM_P_DEGDAYINC, MHM_P_DEGDAYMAX, MHM_P_DEGDAYNOPRE,
                 MHM_P_KFASTFLOW, MHM_P_KSLOWFLOW, MHM_P_KBASEFLOW, MHM_P_KPERCO, MHM_P_TEMPTHRESH})
    rc |= ensure_param(d, id, 1, nLC);
  for (int id : {MHM_P_FROOTS, MHM_P_SOILMOISTFC, MHM_P_SOILMOISTSAT, MHM_P_SOILMOISTEXP, MHM_P_WILTINGPOINT})
    rc |= ensure_param(d, id, nH, nLC);
  for (int id : {MHM_P_KARSTLOSS, MHM_P_JARVIS_C1, MHM_P_UNSATTHRESH, MHM_P_SEALEDTHRESH})
    rc |= ensure_param(d, id, 1, 1);
  rc |= ensure_param(d, MHM_P_MAXINTER, nLAI, 1);
  if (pet_case == -1) rc |= ensure_param(d, MHM_P_PETLAICORFACTOR, nLAI, nLC);
  if (pet_case == 0 || pet_case == 1) rc |= ensure_param(d, MHM_P_FASP, 1, 1);
  if (pet_case == 1) rc |= ensure_param(d, MHM_P_HARSAMCOEFF, 1, 1);
  if (pet_case == 2) rc |= ensure_param(d, MHM_P_PRIETAYALPHA, nLAI, 1);
  if (pet_case == 3) {
    rc |= ensure_param(d, MHM_P_AERORESIST, nLAI, nLC);
    rc |= ensure_param(d, MHM_P_SURFRESIST, nLAI, 1);
  }
  if (rc) return rc;

  double *KsVar_H0 = s->w0, *KsVar_V0 = s->w0 + n0, *SMs_FC0 = s->w0 + 2 * n0, *tmp = s->w0 + 3 * n0;
  double *beta0 = tmp, *SMs0 = tmp + n0, *FC0 = tmp + 2 * n0, *PW0 = tmp + 3 * n0, *fRoots0 = tmp + 4 * n0;
  double *fForest = s->w1, *fPerm = s->w1 + n1, *fKar = s->w1 + 2 * (size_t)n1, *k2_1 = s->w1 + 3 * (size_t)n1;

  auto l0 = [&](int op, const double* p, int np_, const int32_t* LC, const double* a, const double* b,
                double* out) {
    L0Args q{};
    q.n0 = n0;
    q.op = op;
    q.nGeo = (int32_t)s->geoList.size();
    for (int i = 0; i < np_ && i < 6; ++i) q.p[i] = p[i];
    q.LC = LC;
    q.geoUnit0 = s->geoUnit0;
    q.geoList = s->d_geoList;
    q.a = a;
    q.b = b;
    q.geoparam = s->d_geoparam;
    q.out = out;
    mpr_l0_kernel<<<g0, 256, 0, st>>>(q);
  };
  auto l1 = [&](L1Args q) {
    q.n1 = n1;
    mpr_l1_kernel<<<g1, 128, 0, st>>>(q);
  };
  auto up = [&](int op, const double* x, const int32_t* xi, int cls, double* out) {
    return upscale_device(ctx, s->grid, op, kNodata, x, xi, cls, out);
  };

  HostTables tab;
  TabArgs t{};
  t.nSoil = s->nSoil;
  t.maxHor = s->maxHor;
  t.nH = nH;
  t.nHorizons = s->d_nHorizons;
  t.nTill = s->d_nTill;
  t.thetaS_till = s->t_thetaS_till;
  t.thetaFC_till = s->t_thetaFC_till;
  t.thetaPW_till = s->t_thetaPW_till;
  t.Ks = s->t_Ks;
  t.Db = s->t_Db;
  t.thetaS = s->t_thetaS;
  t.thetaFC = s->t_thetaFC;
  t.thetaPW = s->t_thetaPW;
  t.DbM = s->d_DbM;
  t.Wd = s->d_Wd;
  t.RZdepth = s->d_RZdepth;

  // parameter windows of the soil-moisture process, mo_multi_param_reg.f90:357-388
  const int iStart = PM(3, 3) - PM(3, 2) + 1;
  int iStart2;
  switch (soil_case) {
    case 1: iStart2 = PM(3, 3) - 4 + 1; break;
    case 2: iStart2 = PM(3, 3) - 5 + 1; break;
    case 3: iStart2 = PM(3, 3) - 8; break;
    default: iStart2 = PM(3, 3) - 7; break;
  }
  const double* p13 = param + iStart - 1;
  const double* ph = param + iStart2 - 1;

  for (int y = 0; y < nLC; ++y) {
    const int32_t* LC = s->LCover0 + (size_t)y * n0;
    double* fSealed = pslice(d, MHM_P_FSEALED, member, 0, y);
    // land-cover fractions :267-295
    if (int r = up(kOpFrac, nullptr, LC, 1, fForest)) return r;
    if (int r = up(kOpFrac, nullptr, LC, 2, fSealed)) return r;
    {
      L1Args q{};
      q.op = kL1Lc;
      q.p[0] = s->fracSealedCity;
      q.x = fSealed;
      q.y = fPerm;
      q.a = fForest;
      l1(q);
    }
    {  // snow_acc_melt_param :833-850
      const double* p = param + (PM(2, 3) - PM(2, 2));
      L1Args q{};
      q.op = kL1Snow;
      q.p[0] = p[1];
      q.p[1] = p[1] + p[3] + p[2];
      q.p[2] = p[1] + p[3];
      q.p[3] = p[1] + p[5];
      q.p[4] = p[1] + p[3] + p[2] + p[6];
      q.p[5] = p[1] + p[3] + p[7];
      q.p[6] = p[0];
      q.p[7] = p[4];
      q.a = fForest;
      q.b = fSealed;
      q.c = fPerm;
      q.x = pslice(d, MHM_P_TEMPTHRESH, member, 0, y);
      q.y = pslice(d, MHM_P_DEGDAYINC, member, 0, y);
      q.z = pslice(d, MHM_P_DEGDAYNOPRE, member, 0, y);
      q.w = pslice(d, MHM_P_DEGDAYMAX, member, 0, y);
      l1(q);
    }
    if (soil_case == 2 || soil_case == 3) {
      L1Args q{};
      q.op = kL1Fill;
      q.p[0] = param[PM(3, 3) - 1];
      q.x = pslice(d, MHM_P_JARVIS_C1, member, 0, 0);
      l1(q);
    }
    // mpr_sm: host table -> device, then the per-cell column integrals
    soil_table(s, p13, soil_case, s->lc_max[(size_t)y], tab);
    const size_t n2 = (size_t)s->nSoil * s->maxHor;
    struct { double* dst; const std::vector<double>* src; } cp[] = {
        {s->t_thetaS_till, &tab.thetaS_till}, {s->t_thetaFC_till, &tab.thetaFC_till},
        {s->t_thetaPW_till, &tab.thetaPW_till}, {s->t_Ks, &tab.Ks}, {s->t_Db, &tab.Db},
        {s->t_thetaS, &tab.thetaS}, {s->t_thetaFC, &tab.thetaFC}, {s->t_thetaPW, &tab.thetaPW}};
    for (auto& c : cp)
      MHM_CUDA_OK(cudaMemcpyAsync(c.dst, c.src->data(), c.src->size() * sizeof(double),
                                  cudaMemcpyHostToDevice, st));
    MHM_CUDA_OK(cudaStreamSynchronize(st));  // `tab` is reused by the next scene
    (void)n2;
    mpr_ksvar_kernel<<<g0, 256, 0, st>>>(n0, s->soilId0, LC, t, p13[12], KsVar_H0, KsVar_V0, SMs_FC0);

    // mpr_SMhorizons, mo_mpr_smhorizons.f90:327-565
    {
      const double c_forest = ph[0], c_imp = ph[1];
      double c_perv, c_sand = 0.0, c_clay = 0.0, FCmin = 0.0, FCmax = 0.0;
      const bool fc_dep = soil_case == 3 || soil_case == 4;
      if (!fc_dep) {
        c_perv = ph[0] - ph[2];
      } else {
        c_perv = ph[2];
        c_sand = ph[5] - ph[4];
        c_clay = ph[5];
        FCmin = ph[6];
        FCmax = ph[6] + ph[7];
      }
      for (int h = 0; h < nH; ++h) {
        double dpth_f = 0.0, dpth_t = s->horizonDepth[(size_t)h];
        if (h > 0 && h < nH - 1) {
          dpth_f = s->horizonDepth[(size_t)h - 1];
          dpth_t = s->horizonDepth[(size_t)h];
        }
        const int last = h == nH - 1;
        mpr_horizon_kernel<<<g0, 256, 0, st>>>(n0, h, s->soilId0, LC, t, dpth_f, dpth_t, last,
                                               s->horizonDepth[(size_t)nH - 2], ph[3], beta0, SMs0, FC0, PW0);
        // the reference's root-fraction loop reuses dpth_t/dpth_f as the loop above left them:
        // for the last horizon that is RZdepth of the LAST L0 cell's soil type (:424-427)
        if (last) {
          dpth_f = s->horizonDepth[(size_t)nH - 2];
          dpth_t = s->RZdepth[(size_t)s->last_soil - 1];
        }
        auto rf = [&](double c) {
          return (1.0 - std::pow(c, dpth_t * 0.1)) - (1.0 - std::pow(c, dpth_f * 0.1));
        };
        mpr_roots_kernel<<<g0, 256, 0, st>>>(n0, LC, fc_dep ? 1 : 0, rf(c_forest), rf(c_imp), rf(c_perv),
                                             c_sand, c_clay, FCmin, FCmax, dpth_f, dpth_t, FC0, fRoots0);
        if (int r = up(kOpHarm, SMs0, nullptr, 0, pslice(d, MHM_P_SOILMOISTSAT, member, h, y))) return r;
        if (int r = up(kOpHarm, beta0, nullptr, 0, pslice(d, MHM_P_SOILMOISTEXP, member, h, y))) return r;
        if (int r = up(kOpHarm, PW0, nullptr, 0, pslice(d, MHM_P_WILTINGPOINT, member, h, y))) return r;
        if (int r = up(kOpHarm, FC0, nullptr, 0, pslice(d, MHM_P_SOILMOISTFC, member, h, y))) return r;
        if (int r = up(kOpHarm, fRoots0, nullptr, 0, pslice(d, MHM_P_FROOTS, member, h, y))) return r;
      }
      mpr_l1_soil_fix_kernel<<<g1, 128, 0, st>>>(n1, nH, pslice(d, MHM_P_SOILMOISTSAT, member, 0, y),
                                                 pslice(d, MHM_P_SOILMOISTFC, member, 0, y),
                                                 pslice(d, MHM_P_WILTINGPOINT, member, 0, y),
                                                 pslice(d, MHM_P_FROOTS, member, 0, y));
    }
    // PET fields that depend on the land-cover scene :484-499
    if (pet_case == 3) {
      const double* p = param + (PM(5, 3) - PM(5, 2));
      for (int tt = 0; tt < nLAI; ++tt) {
        mpr_aero_kernel<<<g0, 256, 0, st>>>(n0, tt, nLAI, LC, s->LAI0, p[0], p[1], p[2], p[3], p[4], p[5], tmp);
        if (int r = up(kOpArith, tmp, nullptr, 0, pslice(d, MHM_P_AERORESIST, member, tt, y))) return r;
      }
    } else if (pet_case == -1) {
      const double* p = param + (PM(5, 3) - PM(5, 2));
      for (int tt = 0; tt < nLAI; ++tt) {
        l0(kL0PetLai, p, 5, LC, s->LAI0 + (size_t)tt * n0, nullptr, tmp);
        if (int r = up(kOpHarm, tmp, nullptr, 0, pslice(d, MHM_P_PETLAICORFACTOR, member, tt, y))) return r;
      }
    }
    {  // mpr_runoff, mo_mpr_runoff.f90:142-191
      const double* p = param + (PM(6, 3) - PM(6, 2));
      double *K0 = pslice(d, MHM_P_KFASTFLOW, member, 0, y), *K1 = pslice(d, MHM_P_KSLOWFLOW, member, 0, y);
      l0(kL0UnsatThr, p, 5, LC, SMs_FC0, nullptr, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, pslice(d, MHM_P_UNSATTHRESH, member, 0, 0))) return r;
      l0(kL0K0, p, 5, LC, s->slope0, nullptr, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, K0)) return r;
      L1Args q{};
      q.op = kL1ClampMin;
      q.p[0] = 1.0;
      q.x = K0;
      l1(q);
      l0(kL0K1, p, 5, LC, s->slope0, KsVar_H0, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, K1)) return r;
      q.p[0] = 2.0;
      q.x = K1;
      l1(q);
      l0(kL0Alpha, p, 5, LC, SMs_FC0, KsVar_H0, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, pslice(d, MHM_P_ALPHA, member, 0, y))) return r;
      L1Args m{};
      m.op = kL1MinWith;
      m.x = K0;
      m.a = K1;
      l1(m);
    }
    {  // karstic_layer :1009-1032
      const double* p = param + (PM(7, 3) - PM(7, 2));
      double* Kp = pslice(d, MHM_P_KPERCO, member, 0, y);
      l0(kL0Kp, p, 3, LC, SMs_FC0, KsVar_V0, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, Kp)) return r;
      L1Args q{};
      q.op = kL1ClampMin;
      q.p[0] = 2.0;
      q.x = Kp;
      l1(q);
      L1Args z{};
      z.op = kL1Fill;
      z.p[0] = 0.0;
      z.x = fKar;
      l1(z);
      for (size_t i = 0; i < s->geoList.size(); ++i) {  // overwrites, does not accumulate (:1025-1029)
        if (s->geoKar[i] == 0) continue;
        if (int r = up(kOpFrac, nullptr, s->geoUnit0, s->geoList[i], fKar)) return r;
      }
      L1Args kq{};
      kq.op = kL1Karst;
      kq.p[0] = p[1];
      kq.a = fKar;
      kq.x = pslice(d, MHM_P_KARSTLOSS, member, 0, 0);
      l1(kq);
    }
  }
  {  // iper_thres_runoff :893
    L1Args q{};
    q.op = kL1Fill;
    q.p[0] = param[PM(4, 3) - 1];
    q.x = pslice(d, MHM_P_SEALEDTHRESH, member, 0, 0);
    l1(q);
  }
  // PET :555-591
  if (pet_case == 0 || pet_case == 1) {
    const double* p = param + (PM(5, 3) - PM(5, 2));
    l0(kL0Fasp, p, 3, s->LCover0, s->Asp0, s->y0, tmp);
    if (int r = up(kOpArith, tmp, nullptr, 0, pslice(d, MHM_P_FASP, member, 0, 0))) return r;
    if (pet_case == 1) {
      L1Args q{};
      q.op = kL1Fill;
      q.p[0] = param[PM(5, 3) - 1];
      q.x = pslice(d, MHM_P_HARSAMCOEFF, member, 0, 0);
      l1(q);
    }
  } else if (pet_case == 2) {
    const double* p = param + (PM(5, 3) - PM(5, 2));
    for (int tt = 0; tt < nLAI; ++tt) {
      l0(kL0PtAlpha, p, 2, s->LCover0, s->LAI0 + (size_t)tt * n0, nullptr, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, pslice(d, MHM_P_PRIETAYALPHA, member, tt, 0))) return r;
    }
  } else if (pet_case == 3) {
    const double pr = param[PM(5, 3) - 1];
    for (int tt = 0; tt < nLAI; ++tt) {
      l0(kL0SurfRes, &pr, 1, s->LCover0, s->LAI0 + (size_t)tt * n0, nullptr, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, pslice(d, MHM_P_SURFRESIST, member, tt, 0))) return r;
    }
  }
  {  // baseflow_param :719-723 and :596-617
    const double* p = param + (PM(9, 3) - PM(9, 2));
    MHM_REQUIRE(PM(9, 2) == (int)s->geoList.size(), "mpr_eval: %d geo parameters for %zu geological units",
                PM(9, 2), s->geoList.size());
    MHM_CUDA_OK(cudaMemcpyAsync(s->d_geoparam, p, s->geoList.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    MHM_CUDA_OK(cudaStreamSynchronize(st));
    l0(kL0K2, nullptr, 0, s->LCover0, nullptr, nullptr, tmp);
    if (int r = up(kOpArith, tmp, nullptr, 0, k2_1)) return r;
    for (int y = 0; y < nLC; ++y) {
      L1Args q{};
      q.op = kL1Copy;
      q.a = k2_1;
      q.x = pslice(d, MHM_P_KBASEFLOW, member, 0, y);
      l1(q);
      if (PM(7, 1) > 0) {
        L1Args m{};
        m.op = kL1MaxWith;
        m.x = pslice(d, MHM_P_KBASEFLOW, member, 0, y);
        m.a = pslice(d, MHM_P_KSLOWFLOW, member, 0, y);
        l1(m);
      }
    }
  }
  {  // canopy_intercept_param :1142-1151
    const double gamma1 = param[PM(1, 3) - PM(1, 2)];
    for (int tt = 0; tt < nLAI; ++tt) {
      l0(kL0MaxInter, &gamma1, 1, s->LCover0, s->LAI0 + (size_t)tt * n0, nullptr, tmp);
      if (int r = up(kOpArith, tmp, nullptr, 0, pslice(d, MHM_P_MAXINTER, member, tt, 0))) return r;
    }
  }
  MHM_CUDA_OK(cudaGetLastError());
  MHM_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int mhm_cuda_get_param(mhm_cuda_context* ctx, int32_t iDomain, int32_t member, int32_t id, double* base,
                       int64_t ld, int64_t offset, int32_t dim2, int32_t dim3) {
  Domain* d = find_domain(ctx, iDomain);
  if (!d) return 1;
  MHM_REQUIRE(id >= 0 && id < MHM_P_COUNT && member >= 0 && member < d->cfg.nMembers && base &&
                  ld >= d->cfg.nCells && offset >= 0,
              "get_param: bad arguments");
  MHM_REQUIRE(d->P[id], "get_param: parameter %d has not been set", id);
  MHM_REQUIRE(d->P_dim2[id] == dim2 && d->P_dim3[id] == dim3, "get_param(%d): shape (:,%d,%d) held, (:,%d,%d) asked",
              id, d->P_dim2[id], d->P_dim3[id], dim2, dim3);
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->cfg.nCells, rows = (size_t)dim2 * dim3;
  MHM_CUDA_OK(cudaMemcpy2DAsync(base + offset, (size_t)ld * sizeof(double), d->P[id] + (size_t)member * rows * n,
                                n * sizeof(double), n * sizeof(double), rows, cudaMemcpyDeviceToHost, ctx->stream));
  MHM_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
