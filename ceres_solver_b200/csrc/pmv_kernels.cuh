// The four single products of PartitionedMatrixView<2,3,9> (partitioned_matrix_view_impl.h:113-375) on the device Jacobian:
//   y += E x_e   (:113-137)      y += F x_f   (:140-191)      x_e += E'y   (:194-264)      x_f += F'y   (:267-375)
// On the solver path they only exist fused (S*x, J'J x, the implicit-Schur init); these stand-alone forms serve the
// synthetic block-SpMV sweep of BASELINE.json configs[4] (2x3-only, 2x9-only and both shapes) and callers that want the
// reference's partitioned interface.  One thread per row (right products, E'), one warp per camera slice with register
// accumulators (F'y: the reference's transpose block structure, block_sparse_matrix.cc:784-808); every cell is read once.
#pragma once
#include "kernels_v2b.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) pmv_right_e_kernel(ProblemView p, const double* __restrict__ x, double* y) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < static_cast<size_t>(p.N); r += stride) {
    const double2* ep = reinterpret_cast<const double2*>(p.E() + 6 * r);
    const double2 a0 = __ldg(ep), a1 = __ldg(ep + 1), a2 = __ldg(ep + 2);
    const double* xp = x + 3 * static_cast<size_t>(p.pt_of_row[r]);
    const double x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
    double2* yr = reinterpret_cast<double2*>(y + 2 * r);
    double2 v = *yr;
    v.x += a0.x * x0 + a0.y * x1 + a1.x * x2;
    v.y += a1.y * x0 + a2.x * x1 + a2.y * x2;
    *yr = v;
  }
}

__global__ void __launch_bounds__(256) pmv_right_f_kernel(ProblemView p, const double* __restrict__ x, double* y) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < static_cast<size_t>(p.N); r += stride) {
    const double2* fp = reinterpret_cast<const double2*>(p.F() + 18 * r);
    const double* xc = x + 9 * static_cast<size_t>(p.cam_idx[r]);
    double f[18], xv[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const double2 w = __ldg(fp + k);
      f[2 * k] = w.x;
      f[2 * k + 1] = w.y;
      xv[k] = __ldg(xc + k);
    }
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      t0 += f[k] * xv[k];
      t1 += f[9 + k] * xv[k];
    }
    double2* yr = reinterpret_cast<double2*>(y + 2 * r);
    double2 v = *yr;
    v.x += t0;
    v.y += t1;
    *yr = v;
  }
}

// one thread per point: the point's rows are contiguous
__global__ void __launch_bounds__(256) pmv_left_e_kernel(ProblemView p, const double* __restrict__ y, double* x) {
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < p.P; k += stride) {
    double u0 = 0.0, u1 = 0.0, u2 = 0.0;
    for (int r = p.pt_ptr[k]; r < p.pt_ptr[k + 1]; ++r) {
      const double2* ep = reinterpret_cast<const double2*>(p.E() + 6 * static_cast<size_t>(r));
      const double2 a0 = __ldg(ep), a1 = __ldg(ep + 1), a2 = __ldg(ep + 2);
      const double2 yr = *reinterpret_cast<const double2*>(y + 2 * static_cast<size_t>(r));
      u0 += a0.x * yr.x + a1.y * yr.y;
      u1 += a0.y * yr.x + a2.x * yr.y;
      u2 += a1.x * yr.x + a2.y * yr.y;
    }
    double* xp = x + 3 * static_cast<size_t>(k);
    xp[0] += u0;
    xp[1] += u1;
    xp[2] += u2;
  }
}

// one warp per camera slice (camera-major row lists), nine accumulators per lane, nine REDs per slice
__global__ void __launch_bounds__(256) pmv_left_f_kernel(ProblemView p, int num_items, const CamItem* __restrict__ items,
                                                         const int* __restrict__ cam_rows, const double* __restrict__ y, double* x) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int item = blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < num_items; item += gridDim.x * warps_per_block) {
    const CamItem it = items[item];
    double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = it.begin + lane; j < it.end; j += 32) {
      const size_t r = static_cast<size_t>(__ldg(cam_rows + j));
      const double2* fp = reinterpret_cast<const double2*>(p.F() + 18 * r);
      const double2 yr = *reinterpret_cast<const double2*>(y + 2 * r);
      double f[18];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double2 w = __ldg(fp + k);
        f[2 * k] = w.x;
        f[2 * k + 1] = w.y;
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) g[k] += f[k] * yr.x + f[9 + k] * yr.y;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      double v = g[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v != 0.0) red_add(x + 9 * static_cast<size_t>(it.cam) + k, v);
    }
  }
}

// fallback for structures without camera-major lists: one thread per row, nine REDs
__global__ void __launch_bounds__(256) pmv_left_f_rows_kernel(ProblemView p, const double* __restrict__ y, double* x) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < static_cast<size_t>(p.N); r += stride) {
    const double2* fp = reinterpret_cast<const double2*>(p.F() + 18 * r);
    const double2 yr = *reinterpret_cast<const double2*>(y + 2 * r);
    double* xc = x + 9 * static_cast<size_t>(p.cam_idx[r]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double2 a = __ldg(fp + k);
      red_add(xc + 2 * k, a.x * yr.x);
      red_add(xc + 2 * k + 1, a.y * yr.x);
    }
    {
      const double2 a = __ldg(fp + 4);
      red_add(xc + 8, a.x * yr.x);
      red_add(xc + 0, a.y * yr.y);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double2 a = __ldg(fp + 5 + k);
      red_add(xc + 2 * k + 1, a.x * yr.y);
      red_add(xc + 2 * k + 2, a.y * yr.y);
    }
  }
}

}  // namespace b200
