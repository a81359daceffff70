// Camera-sized (9C) PCG vector kernels and parameter-sized (3P+9C) LM vector kernels.
//
// The PCG of the reference (conjugate_gradients_solver.h:109-306) runs ~12 tiny Eigen expressions and
// 3 dot products per iteration, each a host round trip in its CUDA variant (cuda_vector.cc:97-181).  Here one
// single-CTA kernel before and one after the S*p product carry all of it, with every scalar (rho, alpha,
// beta, Q, |r|, iteration count, termination code) living in a device-side CgState so that the host never
// synchronises inside the iteration: kernels exit immediately once `done` is set.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kVecThreads = 1024;

struct CgState {
  double rho, last_rho, Q0, norm_rhs, tol_r, norm_r, alpha, pq, beta;
  int iteration;    // summary.num_iterations
  int done;         // 1 once a termination criterion fired
  int termination;  // B200_LS_*
  int reason;       // 1 zeta, 2 |r|, 3 max iterations, 4 rho, 5 beta, 6 pq, 7 alpha, 8 |b| = 0
};

struct CgParams {
  int n;            // 9C
  int min_iterations, max_iterations;
  double q_tolerance, r_tolerance;
};

__device__ __forceinline__ bool zero_or_inf(double x) { return x == 0.0 || isinf(x); }

// x = 0 (the reference starts ITERATIVE_SCHUR from zero, iterative_schur_complement_solver.cc:98-99),
// r = rhs - S*0 = rhs, Q0 = 0, rho = 1   (conjugate_gradients_solver.h:131-160)
__global__ void __launch_bounds__(kVecThreads)
    cg_begin_kernel(CgParams prm, const double* __restrict__ rhs, double* x, double* r, CgState* st) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < prm.n; i += blockDim.x) {
    const double v = rhs[i];
    x[i] = 0.0;
    r[i] = v;
    acc += v * v;
  }
  const double sq = block_sum<kVecThreads>(acc, scratch);
  if (threadIdx.x == 0) {
    const double norm_rhs = sqrt(sq);
    st->norm_rhs = norm_rhs;
    st->tol_r = prm.r_tolerance * norm_rhs;
    st->norm_r = norm_rhs;
    st->rho = 1.0;
    st->last_rho = 1.0;
    st->Q0 = 0.0;
    st->iteration = 0;
    st->done = 0;
    st->termination = 1;  // NO_CONVERGENCE until proven otherwise
    st->reason = 0;
    if (norm_rhs == 0.0) {
      st->done = 1;
      st->termination = 0;
      st->reason = 8;
    } else if (prm.min_iterations == 0 && norm_rhs <= st->tol_r) {
      st->done = 1;
      st->termination = 0;
      st->reason = 2;
    }
  }
}

// z = M^-1 r ; rho = r.z ; p = z (+ beta p) ; q = Df^2 p (seed of the S*p accumulation)
// precond: 0 identity, otherwise block-diagonal 9x9 inverse blocks in minv [81C].
__global__ void __launch_bounds__(kVecThreads)
    cg_pre_kernel(CgParams prm, int precond, const double* __restrict__ minv, const double* __restrict__ Df,
                  int seed_with_diagonal, const double* __restrict__ r, double* z, double* p, double* q, CgState* st) {
  __shared__ double scratch[32];
  __shared__ double s_beta;
  __shared__ int s_stop;
  if (st->done) return;
  const int it = st->iteration + 1;
  double acc = 0.0;
  for (int i = threadIdx.x; i < prm.n; i += blockDim.x) {
    double zi;
    if (precond == 0) {
      zi = r[i];
    } else {
      const int c = i / 9, row = i - 9 * c;
      const double* m = minv + 81 * static_cast<size_t>(c) + 9 * row;
      const double* rc = r + 9 * static_cast<size_t>(c);
      zi = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) zi += m[k] * rc[k];
    }
    z[i] = zi;
    acc += r[i] * zi;
  }
  const double rho = block_sum<kVecThreads>(acc, scratch);
  if (threadIdx.x == 0) {
    s_stop = 0;
    s_beta = 0.0;
    const double last_rho = st->rho;
    if (zero_or_inf(rho) || isnan(rho)) {
      st->done = 1;
      st->termination = 2;
      st->reason = 4;
      st->iteration = it;
      s_stop = 1;
    } else {
      if (it > 1) {
        const double beta = rho / last_rho;
        if (zero_or_inf(beta)) {
          st->done = 1;
          st->termination = 2;
          st->reason = 5;
          st->iteration = it;
          s_stop = 1;
        }
        s_beta = beta;
      }
      st->last_rho = last_rho;
      st->rho = rho;
      st->iteration = it;
    }
  }
  __syncthreads();
  if (s_stop) return;
  const double beta = s_beta;
  for (int i = threadIdx.x; i < prm.n; i += blockDim.x) {
    const double pi = (it == 1) ? z[i] : z[i] + beta * p[i];
    p[i] = pi;
    q[i] = (seed_with_diagonal && Df != nullptr) ? Df[i] * Df[i] * pi : 0.0;
  }
}

// mode 0: alpha = rho / p.q ; x += alpha p ; r -= alpha q ; then the termination tests
// mode 1: alpha, x += alpha p only            (first half of a residual-reset iteration)
// mode 2: r = rhs - Sx (sx given) ; then the termination tests   (second half)
__global__ void __launch_bounds__(kVecThreads)
    cg_post_kernel(CgParams prm, int mode, const double* __restrict__ rhs, const double* __restrict__ sx, double* x,
                   double* r, const double* __restrict__ p, const double* __restrict__ q, CgState* st) {
  __shared__ double scratch[32];
  __shared__ double s_alpha;
  __shared__ int s_stop;
  if (st->done) return;
  const int it = st->iteration;
  if (mode != 2) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < prm.n; i += blockDim.x) acc += p[i] * q[i];
    const double pq = block_sum<kVecThreads>(acc, scratch);
    if (threadIdx.x == 0) {
      s_stop = 0;
      s_alpha = 0.0;
      st->pq = pq;
      if (!(pq > 0.0) || isinf(pq)) {  // (pq <= 0) || isinf(pq); NaN also lands here
        st->done = 1;
        st->termination = isnan(pq) ? 2 : 1;
        st->reason = 6;
        s_stop = 1;
      } else {
        const double alpha = st->rho / pq;
        if (isinf(alpha)) {
          st->done = 1;
          st->termination = 2;
          st->reason = 7;
          s_stop = 1;
        }
        st->alpha = alpha;
        s_alpha = alpha;
      }
    }
    __syncthreads();
    if (s_stop) return;
  }
  const double alpha = (mode != 2) ? s_alpha : 0.0;
  double accQ = 0.0, accR = 0.0;
  for (int i = threadIdx.x; i < prm.n; i += blockDim.x) {
    double xi = x[i];
    double ri = r[i];
    if (mode != 2) {
      xi += alpha * p[i];
      x[i] = xi;
    }
    if (mode == 0) {
      ri -= alpha * q[i];
      r[i] = ri;
    } else if (mode == 2) {
      ri = rhs[i] - sx[i];
      r[i] = ri;
    }
    accQ += xi * (rhs[i] + ri);
    accR += ri * ri;
  }
  if (mode == 1) return;
  const double dotQ = block_sum<kVecThreads>(accQ, scratch);
  const double sqR = block_sum<kVecThreads>(accR, scratch);
  if (threadIdx.x == 0) {
    const double Q1 = -dotQ;
    const double zeta = it * (Q1 - st->Q0) / Q1;
    const double norm_r = sqrt(sqR);
    st->norm_r = norm_r;
    if (zeta < prm.q_tolerance && it >= prm.min_iterations) {
      st->done = 1;
      st->termination = 0;
      st->reason = 1;
    } else {
      st->Q0 = Q1;
      if (norm_r <= st->tol_r && it >= prm.min_iterations) {
        st->done = 1;
        st->termination = 0;
        st->reason = 2;
      } else if (it >= prm.max_iterations) {
        st->done = 1;
        st->termination = 1;
        st->reason = 3;
      }
    }
  }
}

// y[i] = d[i]^2 * x[i]  (or 0 when d == nullptr)
__global__ void __launch_bounds__(256) diag_sq_mul_kernel(int n, const double* __restrict__ d,
                                                          const double* __restrict__ x, double* y,
                                                          const int* __restrict__ done_flag) {
  if (done_flag != nullptr && *done_flag != 0) return;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = d != nullptr ? d[i] * d[i] * x[i] : 0.0;
}

__global__ void flag_to_double_kernel(const int* flag, double* out) { out[0] = flag[0] != 0 ? 1.0 : 0.0; }
__global__ void double_to_flag_kernel(const double* in, int* flag) { flag[0] = in[0] != 0.0 ? 1 : 0; }

__global__ void __launch_bounds__(256) fill_kernel(size_t n, double* y, double v) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = v;
}

// ---------------------------------------------------------------- LM vector kernels (n = 3P+9C)
// scale = 1 / (1 + sqrt(colnorm^2))     trust_region_minimizer.cc:263-274
__global__ void __launch_bounds__(256) jacobi_scale_kernel(int n, const double* __restrict__ sqnorm, double* scale) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) scale[i] = 1.0 / (1.0 + sqrt(sqnorm[i]));
}
// sq <- sq * scale^2 : squared column norms after J <- J diag(scale)
__global__ void __launch_bounds__(256) rescale_sq_kernel(int n, const double* __restrict__ scale, double* sq) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) sq[i] *= scale[i] * scale[i];
}
// diagonal = clamp(colnorm^2, min, max) (if refresh) ; D = sqrt(diagonal / radius)   levenberg_marquardt_strategy.cc:79-95
__global__ void __launch_bounds__(256) lm_diagonal_kernel(int n, int refresh, const double* __restrict__ sqnorm,
                                                          double* diagonal, double* D, double min_d, double max_d,
                                                          double radius) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = diagonal[i];
    if (refresh) {
      v = fmin(fmax(sqnorm[i], min_d), max_d);
      diagonal[i] = v;
    }
    D[i] = sqrt(v / radius);
  }
}

// Generic two-stage deterministic reductions: partial[blockIdx.x*kSlots + s], then reduce_final_kernel.
constexpr int kRedBlocks = 296;
// step = -y ; delta = step*scale ; cand = x + delta ; partials: {|delta|^2, |x|^2, nonfinite count}
__global__ void __launch_bounds__(256)
    lm_step_kernel(int n, const double* __restrict__ y, const double* __restrict__ scale, const double* __restrict__ x,
                   double* step, double* cand, double* partial) {
  __shared__ double scratch[32];
  double a = 0.0, b = 0.0, c = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double yi = y[i];
    const double si = -yi;
    step[i] = si;
    const double di = si * scale[i];
    const double xi = x[i];
    const double ci = xi + di;
    cand[i] = ci;
    const double dd = xi - ci;
    a += dd * dd;
    b += xi * xi;
    if (!isfinite(yi)) c += 1.0;
  }
  a = block_sum<256>(a, scratch);
  b = block_sum<256>(b, scratch);
  c = block_sum<256>(c, scratch);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3 + 0] = a;
    partial[blockIdx.x * 3 + 1] = b;
    partial[blockIdx.x * 3 + 2] = c;
  }
}
// partials: {max |g|, |g|^2}   (gradient norms: trust_region_minimizer.cc:290-302 with Euclidean Plus)
__global__ void __launch_bounds__(256) grad_norm_kernel(int n, const double* __restrict__ g, double* partial) {
  __shared__ double scratch[32];
  double mx = 0.0, sq = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = g[i];
    mx = fmax(mx, fabs(v));
    sq += v * v;
  }
  sq = block_sum<256>(sq, scratch);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < 8; ++w) m = fmax(m, scratch[w]);
    partial[blockIdx.x * 2 + 0] = m;
    partial[blockIdx.x * 2 + 1] = sq;
  }
}
// out[s] = reduce over blocks of partial[b*slots + s]; op_mask bit s set => max, else sum.
__global__ void __launch_bounds__(32) reduce_final_kernel(int blocks, int slots, unsigned op_mask,
                                                          const double* __restrict__ partial, double* out) {
  const int s = threadIdx.x;
  if (s >= slots) return;
  double acc = 0.0;
  const bool is_max = (op_mask >> s) & 1u;
  for (int b = 0; b < blocks; ++b) {
    const double v = partial[b * slots + s];
    acc = is_max ? fmax(acc, v) : acc + v;
  }
  out[s] = acc;
}
// out[0] = sum_t partial[t]   (per-tile cost partials, model cost partials): single CTA, fixed order
__global__ void __launch_bounds__(kVecThreads) sum_kernel(int n, const double* __restrict__ partial, double* out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = block_sum<kVecThreads>(acc, scratch);
  if (threadIdx.x == 0) out[0] = acc;
}

}  // namespace b200
