// Split-phase PCG (single GPU, camera-local problems): the two halves of the reference's iteration
// (conjugate_gradients_solver.h:162-299) are attached to the kernels that already run, so that no kernel needs a
// grid-wide barrier:
//
//   product kernel (148 persistent CTAs)   prologue: totals of the vector kernel's partial sums -> the reference's
//                                          termination tests of the iteration that just ended, beta, and the CTA's own
//                                          staged copy of p = z + beta p for its camera range (all CTAs compute the
//                                          same bits from the same partials; CTA 0 publishes the state) -- this runs
//                                          while the first tiles are already in flight;
//                                          body: q0 = F'(F p - E P E'F p) into a buffer zeroed one phase earlier;
//                                          flush: p.q partial = p . (its partial of q0) + sum over the cameras it OWNS
//                                          of D_f^2 p^2
//   vector kernel (28 cameras per CTA)     alpha = rho / p.q, p (re-formed for its own entries), x += alpha p,
//                                          r -= alpha (q0 + D_f^2 p), z = M^-1 r, partial sums of x.(b+r), r.r, r.z
//
// The state ping-pongs between two CgState slots (a product reads the slot the previous product wrote and writes the
// other one, so no CTA can see a half-updated state); slot 2 receives the final summary for the host.  Arithmetic and
// termination logic are those of cg_vector_kernel.  Opt-in (B200_SPLIT_PCG=1): measured 48.8 vs 49.6 us per CG iteration
// on Ladybug-1723, 4.41 vs 4.36 ms per LM step on Venice-1778, and 0.80 vs 0.70 ms per LM step on C16 -- the iteration
// is bound by its two kernel boundaries, not by the grid sync this design removes, so the default stays the simpler
// cooperative vector kernel.
#pragma once
#include "kernels_v2.cuh"
#include "vector_kernels.cuh"

namespace b200 {

enum PcgProductMode { PM_FIRST = 0, PM_NORMAL = 1, PM_RESET_X = 2 };
enum CgSplitMode { CA_BEGIN = 0, CA_NORMAL = 1, CA_RESET_FIRST = 2, CA_RESET_SECOND = 3 };

constexpr int kSplitCamsPerCta = 28;
constexpr int kSplitThreads = 256;

struct PcgLink {
  int mode;              // PcgProductMode
  int slot;              // state slot to read; the product writes 1 - slot
  CgState* st;           // [3]
  CgParams prm;
  const double* red;     // [red_n][4] partial sums of the vector kernel: 1 x.(b+r)  2 r.r  3 r.z
  int red_n;
  const double* z;       // [9C]
  const double* p;       // [9C] p of the previous iteration
  const double* xvec;    // [9C] (PM_RESET_X: the vector the product is applied to)
  const double* Df;      // camera part of the LM diagonal or null
  const int2* cta_own;   // per CTA: the camera sub-range whose D_f^2 p^2 it contributes to p.q
  double* pq_parts;      // [gridDim.x]
};

__device__ __forceinline__ void split_publish_final(CgState* st, int iteration, int termination, int reason, double norm_r) {
  CgState* f = st + 2;
  f->iteration = iteration;
  f->termination = termination;
  f->reason = reason;
  f->norm_r = norm_r;
  __threadfence();
  f->done = 1;
}

template <bool kOwned>
__global__ void __launch_bounds__(kV4MaxThreads, 1)
    schur_mul_v4_pcg_kernel(V2View v, const double* __restrict__ ete_inv, PcgLink L, double* y) {
  __shared__ double s_tot[4];
  __shared__ double s_pq;
  const V4Ctx c = v4_ctx(v);
  const int tid = threadIdx.x;
  v4_init(v, c);
  v4_prime(v, ete_inv, c);  // the first tiles stream in while the prologue below runs
  for (int i = tid; i < c.sy_stride * v.replicas; i += blockDim.x) c.sy()[i] = 0.0;
  if (tid == 0) s_pq = 0.0;
  const int range = 9 * (c.cr.y - c.cr.x);
  const size_t range0 = 9 * static_cast<size_t>(c.cr.x);
  const CgState* sr = L.st + L.slot;
  CgState* sw = L.st + (1 - L.slot);
  const bool writer = (blockIdx.x == 0 && tid == 0);
  const int prev_done = __ldcg(&sr->done);
  if (L.mode == PM_RESET_X) {
    if (prev_done) {
      v4_drain(v, c, 0);
      return;
    }
    for (int i = tid; i < range; i += blockDim.x) c.sx()[i] = __ldcg(L.xvec + range0 + i);
  } else {
    // everything that is needed from memory is requested at once
    const double rho_old = __ldcg(&sr->rho), Q0 = __ldcg(&sr->Q0);
    const double tol_prev = __ldcg(&sr->tol_r), nrhs_prev = __ldcg(&sr->norm_rhs);
    const int it = __ldcg(&sr->iteration) + (L.mode == PM_FIRST ? 0 : 1);   // the iteration whose tests run here
    {
      const int warp = tid >> 5, lane = tid & 31;
      if (warp < 3) {
        double acc = 0.0;
        for (int b = lane; b < L.red_n; b += 32) acc += __ldcg(L.red + b * 4 + 1 + warp);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_tot[warp] = acc;
      }
    }
    const bool pre = range <= 2 * static_cast<int>(blockDim.x);
    double z0 = 0.0, z1 = 0.0, p0 = 0.0, p1 = 0.0;
    if (pre) {
      if (tid < range) {
        z0 = __ldcg(L.z + range0 + tid);
        p0 = __ldcg(L.p + range0 + tid);
      }
      if (tid + static_cast<int>(blockDim.x) < range) {
        z1 = __ldcg(L.z + range0 + tid + blockDim.x);
        p1 = __ldcg(L.p + range0 + tid + blockDim.x);
      }
    }
    __syncthreads();
    if (prev_done) {  // keep the flag moving through the slots, nothing else to do
      if (writer) sw->done = 1;
      v4_drain(v, c, 0);
      return;
    }
    const double dotQ = s_tot[0], sqR = s_tot[1], rho_new = s_tot[2];
    const double norm_r = sqrt(sqR);
    int term = -1, reason = 0, fin_iter = it;
    double tol_r = tol_prev, norm_rhs = nrhs_prev, Q0_next = Q0, beta = 0.0;
    if (L.mode == PM_FIRST) {
      norm_rhs = norm_r;
      tol_r = L.prm.r_tolerance * norm_r;
      Q0_next = 0.0;
      if (norm_r == 0.0 || (L.prm.min_iterations == 0 && norm_r <= tol_r)) {
        term = 0;
        reason = norm_r == 0.0 ? 8 : 2;
        fin_iter = 0;
      } else if (zero_or_inf(rho_new) || isnan(rho_new)) {
        term = 2;
        reason = 4;
        fin_iter = 1;
      }
    } else {
      const double Q1 = -dotQ;
      const double zeta = it * (Q1 - Q0) / Q1;
      if (zeta < L.prm.q_tolerance && it >= L.prm.min_iterations) {
        term = 0;
        reason = 1;
      } else if (norm_r <= tol_prev && it >= L.prm.min_iterations) {
        term = 0;
        reason = 2;
      } else if (it >= L.prm.max_iterations) {
        term = 1;
        reason = 3;
      } else if (zero_or_inf(rho_new) || isnan(rho_new)) {
        term = 2;
        reason = 4;
        fin_iter = it + 1;
      } else {
        beta = rho_new / rho_old;
        if (zero_or_inf(beta)) {
          term = 2;
          reason = 5;
          fin_iter = it + 1;
        }
      }
      Q0_next = Q1;
    }
    if (writer) {
      sw->rho = rho_new;
      sw->last_rho = rho_old;
      sw->Q0 = Q0_next;
      sw->norm_rhs = norm_rhs;
      sw->tol_r = tol_r;
      sw->norm_r = norm_r;
      sw->beta = beta;
      sw->iteration = it;
      sw->termination = term >= 0 ? term : 1;
      sw->reason = reason;
      sw->done = term >= 0 ? 1 : 0;
      if (term >= 0) split_publish_final(L.st, fin_iter, term, reason, norm_r);
    }
    if (term >= 0) {
      v4_drain(v, c, 0);
      return;
    }
    // staged p of this CTA's camera range: z + beta p (beta = 0 right after the start: p = z)
    if (pre) {
      if (tid < range) c.sx()[tid] = z0 + beta * p0;
      if (tid + static_cast<int>(blockDim.x) < range) c.sx()[tid + blockDim.x] = z1 + beta * p1;
    } else {
      for (int i = tid; i < range; i += blockDim.x) c.sx()[i] = __ldcg(L.z + range0 + i) + beta * __ldcg(L.p + range0 + i);
    }
  }
  __syncthreads();
  uint32_t flip = 0;
  v4_tiles<kOwned>(v, ete_inv, c, flip);
  v4_big_points(v, c, ete_inv);
  __syncthreads();
  // flush into the zeroed output, p.q partial (incl. the D_f^2 p^2 of the cameras this CTA owns)
  const double* sy = c.sy();
  const double* sx = c.sx();
  const int2 own = L.cta_own[blockIdx.x];
  double pq = 0.0;
  for (int i = tid; i < range; i += blockDim.x) {
    double acc = sy[i];
    for (int r = 1; r < v.replicas; ++r) acc += sy[r * c.sy_stride + i];
    if (acc != 0.0) red_add(y + range0 + i, acc);
    const double pi = sx[i];
    double t = pi * acc;
    const int cam = c.cr.x + i / 9;
    if (L.Df != nullptr && cam >= own.x && cam < own.y) {
      const double d = __ldg(L.Df + range0 + i);
      t += d * d * pi * pi;
    }
    pq += t;
  }
  if (L.mode != PM_RESET_X) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pq += __shfl_xor_sync(0xffffffffu, pq, o);
    if ((tid & 31) == 0) atomicAdd(&s_pq, pq);
    __syncthreads();
    if (tid == 0) L.pq_parts[blockIdx.x] = s_pq;
  }
}

struct CgSplitArgs {
  CgParams prm;
  int mode;      // CgSplitMode
  int C;
  int slot;      // state slot to read (written by the preceding product)
  CgState* st;   // [3]
  const double* Df;
  int precond;
  const double* minv;
  const double* rhs;
  double *x, *r, *z, *p;
  const double* q;      // product output to consume (CA_NORMAL: S0 p, CA_RESET_SECOND: S0 x)
  double* zero_a;       // buffers this launch zeroes for later products (null: none)
  double* zero_b;
  const double* pq_parts;
  int num_pq_parts;
  double* red;          // [gridDim.x][4]
};

__global__ void __launch_bounds__(kSplitThreads) cg_split_kernel(CgSplitArgs a) {
  __shared__ double scratch[kSplitThreads / 32][3];
  __shared__ double s_pq;
  __shared__ double s_r[kSplitCamsPerCta * 9];
  CgState* st = a.st + a.slot;
  const int mode = a.mode;
  const int tid = threadIdx.x;
  const int j = blockIdx.x * kSplitCamsPerCta * 9 + tid;
  const bool lane_ok = tid < kSplitCamsPerCta * 9;
  const bool ok = lane_ok && j < a.prm.n;
  const bool writer = (blockIdx.x == 0 && tid == 0);

  // every operand is requested up front (one L2 round trip)
  int done = 0, st_it = 0;
  double rho = 1.0, beta = 0.0;
  if (mode != CA_BEGIN) {
    done = __ldcg(&st->done);
    st_it = __ldcg(&st->iteration);
    rho = __ldcg(&st->rho);
    beta = __ldcg(&st->beta);
  }
  double pq_lane = 0.0;
  if ((mode == CA_NORMAL || mode == CA_RESET_FIRST) && tid < 32)
    for (int b = tid; b < a.num_pq_parts; b += 32) pq_lane += __ldcg(a.pq_parts + b);
  double bj = 0.0, dj = 0.0, xj = 0.0, rj = 0.0, pj = 0.0, zj = 0.0, qj = 0.0;
  double mrow[9];
  if (ok) {
    bj = a.rhs[j];
    if (a.Df != nullptr) dj = a.Df[j];
    if (mode != CA_BEGIN) {
      xj = a.x[j];
      pj = a.p[j];
      if (mode != CA_RESET_SECOND) {
        rj = a.r[j];
        zj = a.z[j];
      }
      if (mode != CA_RESET_FIRST) qj = __ldcg(a.q + j);
    }
    if (a.precond != 0 && mode != CA_RESET_FIRST) {
      const double* m = a.minv + 9 * static_cast<size_t>(j);
#pragma unroll
      for (int k = 0; k < 9; ++k) mrow[k] = m[k];
    }
  }
  if (mode != CA_BEGIN && done) return;
  const int it = st_it + 1;

  if (mode == CA_NORMAL || mode == CA_RESET_FIRST) {
    if (tid < 32) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pq_lane += __shfl_xor_sync(0xffffffffu, pq_lane, o);
      if (tid == 0) s_pq = pq_lane;
    }
    __syncthreads();
    const double pq = s_pq;
    int term = -1, reason = 0;
    double alpha = 0.0;
    if (!(pq > 0.0) || isinf(pq)) {
      term = isnan(pq) ? 2 : 1;
      reason = 6;
    } else {
      alpha = rho / pq;
      if (isinf(alpha)) {
        term = 2;
        reason = 7;
      }
    }
    if (term >= 0) {  // every CTA takes this branch together
      if (writer) {
        st->pq = pq;
        st->termination = term;
        st->reason = reason;
        st->iteration = it;
        split_publish_final(a.st, it, term, reason, __ldcg(&st->norm_r));
        __threadfence();
        st->done = 1;
      }
      return;
    }
    if (ok) {
      pj = zj + beta * pj;   // p of this iteration (beta = 0 on the first one)
      a.p[j] = pj;
      xj += alpha * pj;
      a.x[j] = xj;
      if (mode == CA_NORMAL) {
        rj -= alpha * (qj + dj * dj * pj);
        a.r[j] = rj;
      }
      if (a.zero_a != nullptr) a.zero_a[j] = 0.0;
    }
    if (mode == CA_RESET_FIRST) {
      if (writer) st->alpha = alpha;
      return;
    }
  } else if (mode == CA_RESET_SECOND) {
    if (ok) {
      rj = bj - (qj + dj * dj * xj);   // r = b - S x
      a.r[j] = rj;
      if (a.zero_a != nullptr) a.zero_a[j] = 0.0;
    }
  } else {  // CA_BEGIN: x = 0, r = b
    if (ok) {
      xj = 0.0;
      rj = bj;
      a.x[j] = 0.0;
      a.r[j] = rj;
      a.p[j] = 0.0;
      if (a.zero_a != nullptr) a.zero_a[j] = 0.0;
      if (a.zero_b != nullptr) a.zero_b[j] = 0.0;
    }
    if (writer) {
      CgState* s0 = a.st;  // the first product reads slot 0
      s0->rho = 1.0;
      s0->last_rho = 1.0;
      s0->Q0 = 0.0;
      s0->norm_rhs = 0.0;
      s0->tol_r = 0.0;
      s0->norm_r = 0.0;
      s0->alpha = 0.0;
      s0->pq = 0.0;
      s0->beta = 0.0;
      s0->iteration = 0;
      s0->done = 0;
      s0->termination = 1;
      s0->reason = 0;
      a.st[1].done = 0;
      a.st[2].done = 0;
      a.st[2].iteration = 0;
      a.st[2].termination = 1;
    }
  }

  // z = M^-1 r (9x9 block per camera) and the partial sums
  double znew = rj;
  if (a.precond != 0) {
    if (lane_ok) s_r[tid] = ok ? rj : 0.0;
    __syncthreads();
    if (ok) {
      const double* rc = s_r + 9 * (tid / 9);
      znew = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) znew += mrow[k] * rc[k];
    }
  }
  double accQ = 0.0, accR = 0.0, accRho = 0.0;
  if (ok) {
    a.z[j] = znew;
    accQ = xj * (bj + rj);
    accR = rj * rj;
    accRho = rj * znew;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    accQ += __shfl_xor_sync(0xffffffffu, accQ, o);
    accR += __shfl_xor_sync(0xffffffffu, accR, o);
    accRho += __shfl_xor_sync(0xffffffffu, accRho, o);
  }
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    scratch[warp][0] = accQ;
    scratch[warp][1] = accR;
    scratch[warp][2] = accRho;
  }
  __syncthreads();
  if (tid == 0) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int w = 0; w < kSplitThreads / 32; ++w) {
      s0 += scratch[w][0];
      s1 += scratch[w][1];
      s2 += scratch[w][2];
    }
    a.red[blockIdx.x * 4 + 1] = s0;
    a.red[blockIdx.x * 4 + 2] = s1;
    a.red[blockIdx.x * 4 + 3] = s2;
  }
}

}  // namespace b200
