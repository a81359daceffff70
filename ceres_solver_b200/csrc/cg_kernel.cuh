// One cooperative kernel per PCG iteration carries everything that is not the S*p product:
//   phase A  q = D_f^2 p + (fixed-order sum of the per-CTA partial vectors of schur_mul_v2) [+ big-point RED buffer],
//            partial p.q
//   ---- grid sync ----
//   phase B  alpha = rho / p.q ; x += alpha p ; r -= alpha q ; z = M^-1 r (9x9 block per camera) ;
//            partial x.(b+r), r.r, r.z
//   ---- grid sync ----
//   phase C  the reference's termination tests (conjugate_gradients_solver.h:245-299) evaluated identically by
//            every CTA from the same partials, beta = rho_new / rho, p = z + beta p ; CTA 0 publishes the state.
// CTAs own whole cameras (28 cameras = 252 entries per 256-thread CTA), so the block-diagonal preconditioner
// needs only the CTA's own slice of r.  All dot products are reduced in a fixed order: the PCG is deterministic
// given q.  Replaces the reference's ~12 Eigen expressions + 3 host-synchronising dots per iteration
// (conjugate_gradients_solver.h:162-299; cuda_vector.cc:97-181 in its CUDA variant).
#pragma once
#include <cooperative_groups.h>

#include "vector_kernels.cuh"

namespace b200 {
namespace cg = cooperative_groups;

constexpr int kCgCamsPerCta = 28;
constexpr int kCgThreads = 256;

enum CgMode { CG_NORMAL = 0, CG_RESET_FIRST = 1, CG_RESET_SECOND = 2, CG_BEGIN = 3 };

struct CgVecArgs {
  CgParams prm;
  int mode;
  int C;
  // phase A inputs
  double* seed_target;         // non-null: this launch also writes the seed of the NEXT product there,
                               // seed_target = Df^2 * (p_new, or x_new in CG_RESET_FIRST)  (0 if Df is null)
  const double* Df;
  // preconditioner
  int precond;                 // 0 identity, 1 block-diagonal inverse blocks
  const double* minv;
  // vectors
  const double* rhs;
  double *x, *r, *z, *p, *q;   // q doubles as the S*x_new buffer in CG_RESET_SECOND
  double* red;                 // [gridDim.x][4] partial sums
  CgState* st;
};

__device__ __forceinline__ double cg_block_sum(double v, double* scratch) {
  return block_sum<kCgThreads>(v, scratch);
}

// Fixed-order sum of slot `slot` over all CTAs' partials; result broadcast to the whole CTA.
__device__ __forceinline__ double cg_total(const double* red, int nb, int slot, double* s_bcast) {
  __syncthreads();
  if (threadIdx.x < 32) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < nb; b += 32) acc += red[b * 4 + slot];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) *s_bcast = acc;
  }
  __syncthreads();
  return *s_bcast;
}

__global__ void __launch_bounds__(kCgThreads) cg_vector_kernel(CgVecArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double scratch[32];
  __shared__ double s_bcast;
  __shared__ double s_r[kCgCamsPerCta * 9];
  CgState* st = a.st;
  const int mode = a.mode;
  if (mode != CG_BEGIN && st->done) return;
  // state of the previous iteration, read before anybody rewrites it
  const double rho_old = (mode == CG_BEGIN) ? 1.0 : st->rho;
  const double Q0 = st->Q0, tol_r = st->tol_r;
  const int it = (mode == CG_BEGIN) ? 0 : st->iteration + (mode == CG_RESET_SECOND ? 0 : 1);
  const int tid = threadIdx.x;
  const int n = a.prm.n;
  const int nblocks = (a.C + kCgCamsPerCta - 1) / kCgCamsPerCta;
  const bool lane_ok = tid < kCgCamsPerCta * 9;
  // ------------------------------------------------------------------ phase A
  if (mode != CG_BEGIN) {
    double acc = 0.0;
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
      const int j = blk * kCgCamsPerCta * 9 + tid;
      if (lane_ok && j < n) {
        const double* src = (mode == CG_RESET_SECOND) ? a.x : a.p;   // vector the product was taken of
        const double qj = a.q[j];
        acc += src[j] * qj;
      }
    }
    acc = cg_block_sum(acc, scratch);
    if (tid == 0) a.red[blockIdx.x * 4 + 0] = acc;
    grid.sync();
  }

  // ------------------------------------------------------------------ phase B
  double alpha = 0.0;
  if (mode == CG_NORMAL || mode == CG_RESET_FIRST) {
    const double pq = cg_total(a.red, gridDim.x, 0, &s_bcast);
    bool stop = false;
    int term = 0, reason = 0;
    if (!(pq > 0.0) || isinf(pq)) {
      stop = true;
      term = isnan(pq) ? 2 : 1;
      reason = 6;
    } else {
      alpha = rho_old / pq;
      if (isinf(alpha)) {
        stop = true;
        term = 2;
        reason = 7;
      }
    }
    if (stop) {
      if (blockIdx.x == 0 && tid == 0) {
        st->pq = pq;
        st->done = 1;
        st->termination = term;
        st->reason = reason;
        st->iteration = it;
      }
      return;  // every CTA takes this branch together
    }
  }
  {
    double accQ = 0.0, accR = 0.0, accRho = 0.0;
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
      const int j = blk * kCgCamsPerCta * 9 + tid;
      const bool ok = lane_ok && j < n;
      double rj = 0.0;
      if (ok) {
        double xj;
        if (mode == CG_BEGIN) {
          xj = 0.0;
          rj = a.rhs[j];
          a.x[j] = 0.0;
        } else if (mode == CG_RESET_SECOND) {
          xj = a.x[j];
          rj = a.rhs[j] - a.q[j];           // r = b - S x   (q holds S x here)
        } else {
          xj = a.x[j] + alpha * a.p[j];
          a.x[j] = xj;
          rj = a.r[j] - alpha * a.q[j];
          if (mode == CG_RESET_FIRST && a.seed_target != nullptr)
            a.seed_target[j] = a.Df != nullptr ? a.Df[j] * a.Df[j] * xj : 0.0;
        }
        if (mode != CG_RESET_FIRST) {
          a.r[j] = rj;
          accQ += xj * (a.rhs[j] + rj);
          accR += rj * rj;
        }
      }
      if (mode != CG_RESET_FIRST) {
        __syncthreads();
        if (lane_ok) s_r[tid] = rj;
        __syncthreads();
        if (ok) {
          double zj;
          if (a.precond == 0) {
            zj = rj;
          } else {
            const int cl = tid / 9, row = tid - 9 * cl;
            const double* m = a.minv + 81 * static_cast<size_t>(blk * kCgCamsPerCta + cl) + 9 * row;
            const double* rc = s_r + 9 * cl;
            zj = 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k) zj += m[k] * rc[k];
          }
          a.z[j] = zj;
          accRho += rj * zj;
        }
      }
    }
    if (mode == CG_RESET_FIRST) {
      if (blockIdx.x == 0 && tid == 0) {
        st->alpha = alpha;
        st->iteration = it;   // iteration `it` is half done; the second half reads it back
      }
      return;
    }
    accQ = cg_block_sum(accQ, scratch);
    accR = cg_block_sum(accR, scratch);
    accRho = cg_block_sum(accRho, scratch);
    if (tid == 0) {
      a.red[blockIdx.x * 4 + 1] = accQ;
      a.red[blockIdx.x * 4 + 2] = accR;
      a.red[blockIdx.x * 4 + 3] = accRho;
    }
  }
  grid.sync();

  // ------------------------------------------------------------------ phase C
  const double dotQ = cg_total(a.red, gridDim.x, 1, &s_bcast);
  const double sqR = cg_total(a.red, gridDim.x, 2, &s_bcast);
  const double rho_new = cg_total(a.red, gridDim.x, 3, &s_bcast);
  const bool writer = (blockIdx.x == 0 && tid == 0);
  const double norm_r = sqrt(sqR);
  double Q0_next = 0.0;
  if (mode == CG_BEGIN) {
    if (writer) {
      st->norm_rhs = norm_r;
      st->tol_r = a.prm.r_tolerance * norm_r;
      st->norm_r = norm_r;
      st->Q0 = 0.0;
      st->iteration = 0;
      st->done = 0;
      st->termination = 1;
      st->reason = 0;
      st->last_rho = 1.0;
    }
    const double tol0 = a.prm.r_tolerance * norm_r;
    if (norm_r == 0.0 || (a.prm.min_iterations == 0 && norm_r <= tol0)) {
      if (writer) {
        st->done = 1;
        st->termination = 0;
        st->reason = norm_r == 0.0 ? 8 : 2;
      }
      return;
    }
  } else {
    // termination tests of iteration `it`
    const double Q1 = -dotQ;
    const double zeta = it * (Q1 - Q0) / Q1;
    int done = 0, term = 1, reason = 0;
    if (zeta < a.prm.q_tolerance && it >= a.prm.min_iterations) {
      done = 1; term = 0; reason = 1;
    } else if (norm_r <= tol_r && it >= a.prm.min_iterations) {
      done = 1; term = 0; reason = 2;
    } else if (it >= a.prm.max_iterations) {
      done = 1; term = 1; reason = 3;
    }
    if (done) {
      if (writer) {
        st->norm_r = norm_r;
        st->iteration = it;
        st->done = 1;
        st->termination = term;
        st->reason = reason;
      }
      return;
    }
    Q0_next = Q1;
  }
  // rho / beta checks of iteration it + 1, then p
  double beta = 0.0;
  {
    int fail_reason = 0;
    if (zero_or_inf(rho_new) || isnan(rho_new)) {
      fail_reason = 4;
    } else if (it >= 1) {
      beta = rho_new / rho_old;
      if (zero_or_inf(beta)) fail_reason = 5;
    }
    if (fail_reason) {
      if (writer) {
        st->iteration = it + 1;
        st->done = 1;
        st->termination = 2;
        st->reason = fail_reason;
      }
      return;
    }
  }
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int j = blk * kCgCamsPerCta * 9 + tid;
    if (lane_ok && j < n) {
      const double pj = (it == 0) ? a.z[j] : a.z[j] + beta * a.p[j];
      a.p[j] = pj;
      if (a.seed_target != nullptr) a.seed_target[j] = a.Df != nullptr ? a.Df[j] * a.Df[j] * pj : 0.0;
    }
  }
  if (writer) {
    st->norm_r = norm_r;
    st->last_rho = rho_old;
    st->rho = rho_new;
    st->Q0 = Q0_next;
    st->iteration = it;
  }
}

}  // namespace b200
