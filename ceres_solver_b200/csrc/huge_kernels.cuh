// Points observed by more than kTile cameras ("huge" points; real photo collections have tracks with thousands of
// observations).  Every kernel that needs no coupling between the rows of a point sees them as CHUNK tiles (TileDesc::chunk:
// a <= kTile-row slice of one point, point-sized outputs accumulated with REDs); the three operations that couple all rows of
// a point through a 3-vector -- (E'E + D^2)^-1 and the rhs, the S*x product, the back substitution -- run here: one CTA per
// huge point, two passes over its rows (first the point-sized sum, then the per-row update), E/F cells read straight from
// global memory (the second pass hits L2).  These launches only exist for problems that have such points.
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr int kHugeThreads = 256;

__global__ void __launch_bounds__(256) huge_zero3_kernel(int H, const int* __restrict__ pts, double* vec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * H) vec[3 * static_cast<size_t>(pts[i / 3]) + i % 3] = 0.0;
}

// Sums `v[0..kN)` over the CTA; the totals are valid in every thread afterwards.
template <int kN>
__device__ __forceinline__ void huge_block_sum(double (&v)[kN], double* scratch /* [kHugeThreads/32][kN] + [kN] */) {
#pragma unroll
  for (int k = 0; k < kN; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < kN; ++k) scratch[warp * kN + k] = v[k];
  __syncthreads();
  if (threadIdx.x < kN) {
    double acc = 0.0;
    for (int w = 0; w < kHugeThreads / 32; ++w) acc += scratch[w * kN + threadIdx.x];
    scratch[(kHugeThreads / 32) * kN + threadIdx.x] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kN; ++k) v[k] = scratch[(kHugeThreads / 32) * kN + k];
}

struct HugeRow {
  double2 e0, e1, e2;  // E cell [2][3] row-major: (e0.x e0.y e1.x ; e1.y e2.x e2.y)
  double f[18];
};
__device__ __forceinline__ void huge_load_row(const ProblemView& p, size_t row, HugeRow& r) {
  const double2* e = reinterpret_cast<const double2*>(p.E() + 6 * row);
  r.e0 = e[0];
  r.e1 = e[1];
  r.e2 = e[2];
  const double2* f = reinterpret_cast<const double2*>(p.F() + 18 * row);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const double2 v = f[k];
    r.f[2 * k] = v.x;
    r.f[2 * k + 1] = v.y;
  }
}

// ImplicitSchurComplement::Init for the huge points (implicit_schur_complement.cc:49-97, :251-276):
// ete_inv[k] = (E'E + D_e^2)^-1, ye[k] = ete_inv E'b, rhs_c += F'(b - E ye).
__global__ void __launch_bounds__(kHugeThreads) huge_schur_init_kernel(ProblemView p, int H, const int* __restrict__ pts, SchurState st) {
  __shared__ double scratch[(kHugeThreads / 32 + 1) * 9];
  for (int hidx = blockIdx.x; hidx < H; hidx += gridDim.x) {
    const int k = pts[hidx];
    const int r0 = p.pt_ptr[k], r1 = p.pt_ptr[k + 1];
    double m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = 0.0;
    for (int r = r0 + threadIdx.x; r < r1; r += kHugeThreads) {
      const double2* e = reinterpret_cast<const double2*>(p.E() + 6 * static_cast<size_t>(r));
      const double2 e0 = e[0], e1 = e[1], e2 = e[2];
      const double2 b = *reinterpret_cast<const double2*>(st.b + 2 * static_cast<size_t>(r));
      m[0] += e0.x * e0.x + e1.y * e1.y;
      m[1] += e0.x * e0.y + e1.y * e2.x;
      m[2] += e0.x * e1.x + e1.y * e2.y;
      m[3] += e0.y * e0.y + e2.x * e2.x;
      m[4] += e0.y * e1.x + e2.x * e2.y;
      m[5] += e1.x * e1.x + e2.y * e2.y;
      m[6] += e0.x * b.x + e1.y * b.y;
      m[7] += e0.y * b.x + e2.x * b.y;
      m[8] += e1.x * b.x + e2.y * b.y;
    }
    huge_block_sum<9>(m, scratch);
    if (st.D != nullptr) {
      const double d0 = st.D[3 * static_cast<size_t>(k)], d1 = st.D[3 * static_cast<size_t>(k) + 1], d2 = st.D[3 * static_cast<size_t>(k) + 2];
      m[0] += d0 * d0;
      m[3] += d1 * d1;
      m[5] += d2 * d2;
    }
    double inv[6];
    invert_sym3_llt(m, inv);
    const double v0 = inv[0] * m[6] + inv[1] * m[7] + inv[2] * m[8];
    const double v1 = inv[1] * m[6] + inv[3] * m[7] + inv[4] * m[8];
    const double v2 = inv[2] * m[6] + inv[4] * m[7] + inv[5] * m[8];
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) st.ete_inv[6 * static_cast<size_t>(k) + i] = inv[i];
      if (st.ye != nullptr) {
        st.ye[3 * static_cast<size_t>(k)] = v0;
        st.ye[3 * static_cast<size_t>(k) + 1] = v1;
        st.ye[3 * static_cast<size_t>(k) + 2] = v2;
      }
    }
    for (int r = r0 + threadIdx.x; r < r1; r += kHugeThreads) {
      HugeRow row;
      huge_load_row(p, r, row);
      const double2 b = *reinterpret_cast<const double2*>(st.b + 2 * static_cast<size_t>(r));
      const double t0 = b.x - (row.e0.x * v0 + row.e0.y * v1 + row.e1.x * v2);
      const double t1 = b.y - (row.e1.y * v0 + row.e2.x * v1 + row.e2.y * v2);
      double* rc = st.rhs + 9 * static_cast<size_t>(p.cam_idx[r]);
#pragma unroll
      for (int i = 0; i < 9; ++i) red_add(rc + i, row.f[i] * t0 + row.f[9 + i] * t1);
    }
    __syncthreads();
  }
}

// y += F'(F x - E (E'E+D^2)^-1 E'F x) for the huge points (implicit_schur_complement.cc:106-144).
__global__ void __launch_bounds__(kHugeThreads)
    huge_schur_mul_kernel(ProblemView p, int H, const int* __restrict__ pts, const double* __restrict__ ete_inv,
                          const double* __restrict__ x, double* y, const int* __restrict__ done_flag) {
  if (done_flag != nullptr && *done_flag != 0) return;
  __shared__ double scratch[(kHugeThreads / 32 + 1) * 3];
  for (int hidx = blockIdx.x; hidx < H; hidx += gridDim.x) {
    const int k = pts[hidx];
    const int r0 = p.pt_ptr[k], r1 = p.pt_ptr[k + 1];
    double u[3] = {0.0, 0.0, 0.0};
    for (int r = r0 + threadIdx.x; r < r1; r += kHugeThreads) {
      HugeRow row;
      huge_load_row(p, r, row);
      const double* xc = x + 9 * static_cast<size_t>(p.cam_idx[r]);
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const double xi = __ldcg(xc + i);
        t0 += row.f[i] * xi;
        t1 += row.f[9 + i] * xi;
      }
      u[0] += row.e0.x * t0 + row.e1.y * t1;
      u[1] += row.e0.y * t0 + row.e2.x * t1;
      u[2] += row.e1.x * t0 + row.e2.y * t1;
    }
    huge_block_sum<3>(u, scratch);
    const double* pi = ete_inv + 6 * static_cast<size_t>(k);
    const double v0 = -(pi[0] * u[0] + pi[1] * u[1] + pi[2] * u[2]);
    const double v1 = -(pi[1] * u[0] + pi[3] * u[1] + pi[4] * u[2]);
    const double v2 = -(pi[2] * u[0] + pi[4] * u[1] + pi[5] * u[2]);
    for (int r = r0 + threadIdx.x; r < r1; r += kHugeThreads) {
      HugeRow row;
      huge_load_row(p, r, row);
      const int cam = p.cam_idx[r];
      const double* xc = x + 9 * static_cast<size_t>(cam);
      double t0 = row.e0.x * v0 + row.e0.y * v1 + row.e1.x * v2;
      double t1 = row.e1.y * v0 + row.e2.x * v1 + row.e2.y * v2;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const double xi = __ldcg(xc + i);
        t0 += row.f[i] * xi;
        t1 += row.f[9 + i] * xi;
      }
      double* yc = y + 9 * static_cast<size_t>(cam);
#pragma unroll
      for (int i = 0; i < 9; ++i) red_add(yc + i, row.f[i] * t0 + row.f[9 + i] * t1);
    }
    __syncthreads();
  }
}

// y_e[k] = (E'E+D^2)^-1 E'(b - F x_f) for the huge points (implicit_schur_complement.cc:208-243); the camera part of y is
// copied by the caller.
__global__ void __launch_bounds__(kHugeThreads)
    huge_backsub_kernel(ProblemView p, int H, const int* __restrict__ pts, const double* __restrict__ ete_inv,
                        const double* __restrict__ b, const double* __restrict__ xf, double* y) {
  __shared__ double scratch[(kHugeThreads / 32 + 1) * 3];
  for (int hidx = blockIdx.x; hidx < H; hidx += gridDim.x) {
    const int k = pts[hidx];
    const int r0 = p.pt_ptr[k], r1 = p.pt_ptr[k + 1];
    double u[3] = {0.0, 0.0, 0.0};
    for (int r = r0 + threadIdx.x; r < r1; r += kHugeThreads) {
      HugeRow row;
      huge_load_row(p, r, row);
      const double2 bb = *reinterpret_cast<const double2*>(b + 2 * static_cast<size_t>(r));
      double t0 = bb.x, t1 = bb.y;
      if (xf != nullptr) {
        const double* xc = xf + 9 * static_cast<size_t>(p.cam_idx[r]);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          t0 -= row.f[i] * xc[i];
          t1 -= row.f[9 + i] * xc[i];
        }
      }
      u[0] += row.e0.x * t0 + row.e1.y * t1;
      u[1] += row.e0.y * t0 + row.e2.x * t1;
      u[2] += row.e1.x * t0 + row.e2.y * t1;
    }
    huge_block_sum<3>(u, scratch);
    if (threadIdx.x == 0) {
      const double* pi = ete_inv + 6 * static_cast<size_t>(k);
      y[3 * static_cast<size_t>(k)] = pi[0] * u[0] + pi[1] * u[1] + pi[2] * u[2];
      y[3 * static_cast<size_t>(k) + 1] = pi[1] * u[0] + pi[3] * u[1] + pi[4] * u[2];
      y[3 * static_cast<size_t>(k) + 2] = pi[2] * u[0] + pi[4] * u[1] + pi[5] * u[2];
    }
    __syncthreads();
  }
}

}  // namespace b200
