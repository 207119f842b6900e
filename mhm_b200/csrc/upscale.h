// upscale.h -- the L0 grid object shared by upscale.cu and mpr.cu
#pragma once
#include <cstdint>

#include "context.h"

struct mpr_l0_grid {
  int32_t nrows0 = 0, ncols0 = 0, nL1 = 0;
  int64_t nL0 = 0;          // number of unmasked L0 cells
  int32_t* cell_of = nullptr;  // device [ncols0][nrows0]: packed index or -1
  int32_t *iu = nullptr, *id = nullptr, *jl = nullptr, *jr = nullptr, *nsub = nullptr;
  double *d_in = nullptr, *d_out = nullptr;  // staging
  int32_t* d_in_i = nullptr;
};

namespace mhm {
enum { kOpArith = 0, kOpHarm = 1, kOpGeom = 2, kOpFrac = 3 };
// reduce a device-resident packed L0 field (double d_x, or int d_xi for kOpFrac) to L1
int upscale_device(mhm_cuda_context* ctx, const mpr_l0_grid* g, int op, double nodata,
                   const double* d_x, const int32_t* d_xi, int32_t class_id, double* d_out);
}  // namespace mhm
