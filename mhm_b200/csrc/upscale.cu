// upscale.cu -- MPR L0 -> L1 upscaling operators as segmented reductions.
//
// Restates MPR/mo_upscaling_operators.f90:
//   upscale_arithmetic_mean :266-329   sum(x[iu:id, jl:jr], mask) / n_subcells
//   upscale_harmonic_mean   :369-432   n_subcells / sum(1/x[iu:id, jl:jr], mask)
//   upscale_geometric_mean  :469-535   product(x != nodata) ** (1 / count)
//   L0_fractionalCover_in_Lx:152-227   count(class == c) / n_subcells
// The reference unpacks the packed L0 vector into a 2-D field on every call and then loops
// serially over the L1 cells.  Here the unpacking is a device-resident index map built once
// per grid, and every L1 cell is reduced
//   math mode 1 (fast)  : by one warp, lanes striding over the rectangle in Fortran element
//                         order, then a shuffle tree (rounding differs from a serial sum by
//                         O(n eps));
//   math mode 0 (strict): by one thread in Fortran element order = bit-identical to the
//                         reference's serial sum().
// Compiled with -fmad=false.
#include <vector>

#include "context.h"

#include "upscale.h"

namespace mhm {


struct UpArgs {
  int32_t nrows0, nL1, op, class_id;
  double nodata;
  const int32_t *cell_of, *iu, *id, *jl, *jr, *nsub;
  const double* x;
  const int32_t* xi;
  double* out;
};

__device__ __forceinline__ bool ne_eps(double a, double b) {  // FORCES mo_utils::ne
  return (2.220446049250313e-16 * fabs(b) - fabs(a - b)) < 0.0;
}

// value of element e (Fortran element order inside the rectangle) folded into (acc, cnt)
__device__ __forceinline__ void fold(const UpArgs& a, int i, int j, double& acc, int& cnt) {
  const int k = a.cell_of[(size_t)j * a.nrows0 + i];
  if (a.op == kOpFrac) {
    if (k >= 0 && a.xi[k] == a.class_id) cnt++;
    return;
  }
  if (k < 0) return;
  const double v = a.x[k];
  if (a.op == kOpArith) {
    acc = acc + v;
  } else if (a.op == kOpHarm) {
    acc = acc + 1.0 / v;
  } else if (ne_eps(v, a.nodata)) {
    acc = acc * v;
    cnt++;
  }
}

__device__ __forceinline__ double finish(const UpArgs& a, int kk, double acc, int cnt) {
  const double n = (double)a.nsub[kk];
  switch (a.op) {
    case kOpArith: return acc / n;
    case kOpHarm: return n / acc;
    case kOpFrac: return (double)cnt / n;
    default:  // geometric mean :527-531
      if (cnt == 0) return 1.0;  // product of an empty set, ** (1/0) = 1 ** inf
      if (ne_eps(acc, 0.0)) return pow(acc, 1.0 / (double)cnt);
      return 0.0;
  }
}

__global__ void upscale_serial_kernel(const UpArgs a) {
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= a.nL1) return;
  const int iu = a.iu[kk] - 1, id = a.id[kk] - 1, jl = a.jl[kk] - 1, jr = a.jr[kk] - 1;
  double acc = a.op == kOpGeom ? 1.0 : 0.0;
  int cnt = 0;
  for (int j = jl; j <= jr; ++j)
    for (int i = iu; i <= id; ++i) fold(a, i, j, acc, cnt);
  a.out[kk] = finish(a, kk, acc, cnt);
}

__global__ void upscale_warp_kernel(const UpArgs a) {
  const int kk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (kk >= a.nL1) return;
  const int iu = a.iu[kk] - 1, id = a.id[kk] - 1, jl = a.jl[kk] - 1, jr = a.jr[kk] - 1;
  const int h = id - iu + 1, total = h * (jr - jl + 1);
  double acc = a.op == kOpGeom ? 1.0 : 0.0;
  int cnt = 0;
  for (int e = lane; e < total; e += 32) fold(a, iu + e % h, jl + e / h, acc, cnt);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, acc, off);
    acc = a.op == kOpGeom ? acc * o : acc + o;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if (lane == 0) a.out[kk] = finish(a, kk, acc, cnt);
}

// device-resident operands (used by the MPR pipeline, mpr.cu)
int upscale_device(mhm_cuda_context* ctx, const mpr_l0_grid* g, int op, double nodata,
                   const double* d_x, const int32_t* d_xi, int32_t class_id, double* d_out) {
  cudaStream_t st = ctx->stream;
  UpArgs a{};
  a.nrows0 = g->nrows0;
  a.nL1 = g->nL1;
  a.op = op;
  a.class_id = class_id;
  a.nodata = nodata;
  a.cell_of = g->cell_of;
  a.iu = g->iu;
  a.id = g->id;
  a.jl = g->jl;
  a.jr = g->jr;
  a.nsub = g->nsub;
  a.x = d_x;
  a.xi = d_xi;
  a.out = d_out;
  ctx->stat_begin(kStatUpscale);
  if (ctx->math_mode == 1) {
    const int warps_per_block = 8;
    upscale_warp_kernel<<<(g->nL1 + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(a);
  } else {
    upscale_serial_kernel<<<(g->nL1 + 127) / 128, 128, 0, st>>>(a);
  }
  ctx->stat_end(kStatUpscale);
  MHM_CUDA_OK(cudaGetLastError());
  return 0;
}

static int run_upscale(mhm_cuda_context* ctx, const mpr_l0_grid* g, int op, double nodata,
                       const double* x, const int32_t* xi, int32_t class_id, double* out) {
  MHM_REQUIRE(ctx && g && out && (x || xi), "upscale: null argument");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  auto* gm = const_cast<mpr_l0_grid*>(g);
  if (x)
    MHM_CUDA_OK(cudaMemcpyAsync(gm->d_in, x, (size_t)g->nL0 * sizeof(double), cudaMemcpyHostToDevice, st));
  else
    MHM_CUDA_OK(cudaMemcpyAsync(gm->d_in_i, xi, (size_t)g->nL0 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (int rc = upscale_device(ctx, g, op, nodata, g->d_in, g->d_in_i, class_id, g->d_out)) return rc;
  MHM_CUDA_OK(cudaMemcpyAsync(out, g->d_out, (size_t)g->nL1 * sizeof(double), cudaMemcpyDeviceToHost, st));
  MHM_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace mhm

using namespace mhm;

extern "C" {

int mpr_cuda_grid_create(mhm_cuda_context* ctx, int32_t nrows0, int32_t ncols0, const int32_t* mask0,
                         int32_t nL1, const int32_t* upper, const int32_t* lower, const int32_t* left,
                         const int32_t* right, const int32_t* nsub, mpr_l0_grid** out) {
  MHM_REQUIRE(ctx && mask0 && upper && lower && left && right && nsub && out && nrows0 > 0 &&
                  ncols0 > 0 && nL1 > 0,
              "mpr_cuda_grid_create: bad arguments");
  MHM_CUDA_OK(cudaSetDevice(ctx->device));
  // unpack order of Fortran's unpack(vector, mask, field): array element order, first index fastest
  std::vector<int32_t> cell_of((size_t)nrows0 * ncols0);
  int64_t k = 0;
  for (size_t e = 0; e < cell_of.size(); ++e) cell_of[e] = mask0[e] ? (int32_t)k++ : -1;
  for (int c = 0; c < nL1; ++c) {
    MHM_REQUIRE(upper[c] >= 1 && lower[c] <= nrows0 && upper[c] <= lower[c] && left[c] >= 1 &&
                    right[c] <= ncols0 && left[c] <= right[c] && nsub[c] >= 0,
                "mpr_cuda_grid_create: bounds of L1 cell %d outside the L0 grid", c + 1);
  }
  auto* g = new mpr_l0_grid();
  g->nrows0 = nrows0;
  g->ncols0 = ncols0;
  g->nL1 = nL1;
  g->nL0 = k;
  cudaStream_t st = ctx->stream;
  auto up = [&](int32_t** dst, const int32_t* src, size_t n) -> int {
    MHM_CUDA_OK(cudaMalloc(dst, n * sizeof(int32_t)));
    MHM_CUDA_OK(cudaMemcpyAsync(*dst, src, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    return 0;
  };
  int rc = up(&g->cell_of, cell_of.data(), cell_of.size());
  rc |= up(&g->iu, upper, (size_t)nL1);
  rc |= up(&g->id, lower, (size_t)nL1);
  rc |= up(&g->jl, left, (size_t)nL1);
  rc |= up(&g->jr, right, (size_t)nL1);
  rc |= up(&g->nsub, nsub, (size_t)nL1);
  if (!rc && cudaMalloc(&g->d_in, (size_t)(k ? k : 1) * sizeof(double)) != cudaSuccess) rc = 2;
  if (!rc && cudaMalloc(&g->d_in_i, (size_t)(k ? k : 1) * sizeof(int32_t)) != cudaSuccess) rc = 2;
  if (!rc && cudaMalloc(&g->d_out, (size_t)nL1 * sizeof(double)) != cudaSuccess) rc = 2;
  cudaStreamSynchronize(st);
  if (rc) {
    set_error("mpr_cuda_grid_create: device allocation failed");
    mpr_cuda_grid_destroy(ctx, g);
    return 2;
  }
  *out = g;
  return 0;
}

int mpr_cuda_grid_destroy(mhm_cuda_context* ctx, mpr_l0_grid* g) {
  if (!g) return 0;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  void* ptrs[] = {g->cell_of, g->iu, g->id, g->jl, g->jr, g->nsub, g->d_in, g->d_in_i, g->d_out};
  for (void* p : ptrs) cudaFree(p);
  delete g;
  return 0;
}

int mpr_cuda_upscale_arithmetic_mean(mhm_cuda_context* ctx, const mpr_l0_grid* g, double nodata,
                                     const double* x, double* out) {
  return run_upscale(ctx, g, kOpArith, nodata, x, nullptr, 0, out);
}
int mpr_cuda_upscale_harmonic_mean(mhm_cuda_context* ctx, const mpr_l0_grid* g, double nodata,
                                   const double* x, double* out) {
  return run_upscale(ctx, g, kOpHarm, nodata, x, nullptr, 0, out);
}
int mpr_cuda_upscale_geometric_mean(mhm_cuda_context* ctx, const mpr_l0_grid* g, double nodata,
                                    const double* x, double* out) {
  return run_upscale(ctx, g, kOpGeom, nodata, x, nullptr, 0, out);
}
int mpr_cuda_l0_fractional_cover(mhm_cuda_context* ctx, const mpr_l0_grid* g, const int32_t* dataIn0,
                                 int32_t class_id, double* out) {
  return run_upscale(ctx, g, kOpFrac, 0.0, nullptr, dataIn0, class_id, out);
}

}  // extern "C"
