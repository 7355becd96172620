// Operator of the multigrid levels as seen by a cell (mg.cu, mg_coop.cu): coefficients rebuilt from the 1-D arrays of
// MGKernelBox (init_matrix_2d / init_matrix_3d, reference src/solver.f90:158-372,687-842), never stored per cell.
#pragma once
#include "common.h"

namespace snac {

struct MGCoef {  // coefficients of one cell
  double diag, w[3];
};

__device__ inline double mg_offdiag(const MGKernelBox &B, const double *x, long c, int i, int j, int k,
                                    double wx, double wy, double wz) {
  double off = 0.0;
  if (B.coupled[0]) off += wy * wz * (__ldg(B.am[0] + i) * x[c - 1] + __ldg(B.ap[0] + i) * x[c + 1]);
  if (B.coupled[1]) off += wx * wz * (__ldg(B.am[1] + j) * x[c - B.sy] + __ldg(B.ap[1] + j) * x[c + B.sy]);
  if (B.coupled[2]) off += wx * wy * (__ldg(B.am[2] + k) * x[c - B.sz] + __ldg(B.ap[2] + k) * x[c + B.sz]);
  return off;
}

__device__ inline double mg_diag(const MGKernelBox &B, int i, int j, int k, double wx, double wy, double wz,
                                 double alpha) {
  double lam = alpha;
  if (B.fft_dir >= 0) lam += __ldg(B.lam + (B.fft_dir == 0 ? i : (B.fft_dir == 1 ? j : k)));
  return wy * wz * __ldg(B.ad[0] + i) + wx * wz * __ldg(B.ad[1] + j) + wx * wy * __ldg(B.ad[2] + k) + lam * wx * wy * wz;
}

}  // namespace snac
