// Kernels for the general-preconditioner PCG path: the power-series-expansion preconditioner /
// initialisation of ITERATIVE_SCHUR (power_series_expansion_preconditioner.cc:57-82,
// iterative_schur_complement_solver.cc:100-111, :178-186) and a PCG that accepts any preconditioner and a non-zero
// initial guess (conjugate_gradients_solver.h:109-306 statement by statement).  The series operator
//   (F'F + D_f^2)^-1 F'E (E'E + D_e^2)^-1 E'F x  =  x - M^-1 (S x),      M = F'F + D_f^2  (block diagonal, 9x9 per camera)
// is evaluated with the same fused S*x product the PCG uses (implicit_schur_complement.cc:146-174 spells it as five
// SpMVs), so every path of that product (v4, v1, multi-GPU) serves it unchanged.  Scalars live on the host here
// (three small synchronisations per iteration): each iteration costs 1 + max_num_spse_iterations products, which
// dwarfs them.
#pragma once
#include "vector_kernels.cuh"

namespace b200 {

// y = M^-1 x   (one thread per entry; minv: [C][9][9] row-major)
__global__ void __launch_bounds__(256) block_apply_kernel(int n, const double* __restrict__ minv, const double* __restrict__ x,
                                                          double* __restrict__ y) {
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const double* m = minv + 9 * static_cast<size_t>(j);
    const double* xc = x + 9 * (j / 9);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc += m[k] * xc[k];
    y[j] = acc;
  }
}

// term = prev - M^-1 t  (t = S prev);  y += term;  partial[block] = |term|^2 over the block's entries
__global__ void __launch_bounds__(256) spse_term_kernel(int n, const double* __restrict__ minv, const double* __restrict__ prev,
                                                        const double* __restrict__ t, double* __restrict__ term, double* y,
                                                        double* partial) {
  __shared__ double scratch[32];
  double sq = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const double* m = minv + 9 * static_cast<size_t>(j);
    const double* tc = t + 9 * (j / 9);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc += m[k] * tc[k];
    const double v = prev[j] - acc;
    term[j] = v;
    y[j] += v;
    sq += v * v;
  }
  sq = block_sum<256>(sq, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = sq;
}

// partial[block*2 + {0,1}] = {a.b, c.d} over the block's entries (c, d may be null: second slot 0)
__global__ void __launch_bounds__(256) dot2_kernel(int n, const double* __restrict__ a, const double* __restrict__ b,
                                                   const double* __restrict__ c, const double* __restrict__ d, double* partial) {
  __shared__ double scratch[32];
  double s0 = 0.0, s1 = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    s0 += a[j] * b[j];
    if (c != nullptr) s1 += c[j] * d[j];
  }
  s0 = block_sum<256>(s0, scratch);
  s1 = block_sum<256>(s1, scratch);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 2 + 0] = s0;
    partial[blockIdx.x * 2 + 1] = s1;
  }
}

// out = a*x + b*y   (out may alias x or y)
__global__ void __launch_bounds__(256) axpby_kernel(int n, double a, const double* x, double b, const double* y, double* out) {
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) out[j] = a * x[j] + b * y[j];
}

// The tail of a PCG iteration (conjugate_gradients_solver.h:230-262): x += alpha p; r -= alpha q (or r = rhs - q when q
// holds S x after a residual reset: `reset` != 0, x is not touched then); partial {x.(rhs + r), r.r}.
__global__ void __launch_bounds__(256) cgg_update_kernel(int n, int reset, double alpha, const double* __restrict__ p,
                                                         const double* __restrict__ q, const double* __restrict__ rhs, double* x,
                                                         double* r, double* partial) {
  __shared__ double scratch[32];
  double s0 = 0.0, s1 = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    double xj = x[j], rj;
    if (reset) {
      rj = rhs[j] - q[j];
    } else {
      xj += alpha * p[j];
      x[j] = xj;
      rj = r[j] - alpha * q[j];
    }
    r[j] = rj;
    s0 += xj * (rhs[j] + rj);
    s1 += rj * rj;
  }
  s0 = block_sum<256>(s0, scratch);
  s1 = block_sum<256>(s1, scratch);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 2 + 0] = s0;
    partial[blockIdx.x * 2 + 1] = s1;
  }
}

}  // namespace b200
