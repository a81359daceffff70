// The whole preconditioned CG of ITERATIVE_SCHUR as ONE persistent kernel (single GPU, camera-local problems):
// one CTA per SM runs, per iteration,
//     S*p over its warp tiles (v4 loop)  ->  grid barrier  ->  alpha, x, r, z = M^-1 r  ->  grid barrier  ->
//     termination tests, beta, p
// with no kernel boundary, no host round trip and no cold start of the TMA rings: the first tiles of the next
// product are requested before the flush of the current one, so the HBM stream never drains during the vector phases.
//
// What makes two barriers per iteration enough:
//   * q = S p is assembled by REDs into a buffer that its owners zeroed one iteration earlier (two buffers
//     alternate); the D_f^2 p term is added by the owner when it reads q, so nothing has to be seeded first;
//   * p.q comes out of the product itself: every CTA dots its private partial of q with its staged copy of p, the
//     owners add sum D_f^2 p^2 when they form p;
//   * every CTA updates its OWN staged copy of p (its camera range) as z + beta p after the second barrier -- the same
//     expression on the same operands as the owner's, so all copies are bit-identical and no third barrier is needed.
// The arithmetic and the termination logic are those of cg_vector_kernel (conjugate_gradients_solver.h:131-299),
// including the residual reset r = b - S x every `reset_period` iterations (two extra barriers on those iterations).
#pragma once
#include "kernels_v2.cuh"
#include "vector_kernels.cuh"

namespace b200 {

struct PcgArgs {
  V2View v;
  const double* ete_inv;
  CgParams prm;
  int C;
  int cams_per_cta;   // cameras owned by a CTA in the vector phases (9 * cams_per_cta <= blockDim)
  int reset_period;
  int precond;        // 0 identity, 1 block-diagonal inverse blocks
  const double* Df;   // camera part of the LM diagonal (may be null)
  const double* minv;
  const double* rhs;
  double *x, *r, *z, *p, *qa, *qb, *tmp;
  double* red;        // [gridDim.x][8] partial sums: 0 p.(S0 p)  1 sum Df^2 p^2  2 x.(b+r)  3 r.r  4 r.z
  CgState* st;
  unsigned* barrier;  // zeroed by the host before the launch
  unsigned long long* trace;  // optional [gridDim.x][trace_iters][8] globaltimer stamps (ns) of the phases, or null
  int trace_iters;
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PCG_STAMP(k)                                                                                     \
  do {                                                                                                   \
    if (a.trace != nullptr && tid == 0 && it <= a.trace_iters)                                           \
      a.trace[(static_cast<size_t>(blockIdx.x) * a.trace_iters + (it - 1)) * 8 + (k)] = global_ns();     \
  } while (0)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (co-resident, cooperative) grid.  `target` is the running arrival count every thread tracks.
// ctr[0] counts arrivals, ctr[1] is an abort word: a CTA that waits longer than ~2 s (a peer died or diverged) publishes
// its barrier sequence number there and every CTA leaves its loop -- the kernel always terminates.
// Returns false when the grid is aborting.
__device__ __forceinline__ bool pcg_grid_barrier(unsigned* ctr, unsigned& target, int* s_abort) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    int bad = 0;
    while (ld_acquire_u32(ctr) < target) {
      if (ld_acquire_u32(ctr + 1) != 0u) {
        bad = 1;
        break;
      }
      if (clock64() - t0 > 4000000000LL) {
        atomicCAS(ctr + 1, 0u, target / gridDim.x);
        bad = 1;
        break;
      }
    }
    *s_abort = bad;
    __threadfence();
  }
  __syncthreads();
  return *s_abort == 0;
}

// Scratch of warp w (96 doubles; idle outside the tile loop): [0,64) slice of the r exchange, [64,67) reduction slots,
// warp 0 [72,80) grid totals (v4_big_points uses warp 0's [64,80) too, but only while it runs).
__device__ __forceinline__ double* pcg_scratch(const V2View& v, const V4Ctx& c, int w) {
  return reinterpret_cast<double*>(c.ring() + static_cast<size_t>(w) * v.per_warp_bytes + v.stages * kV4StageBytes);
}

__device__ __forceinline__ void pcg_block_sum3(const V2View& v, const V4Ctx& c, double& a, double& b, double& d) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) {
    double* s = c.sW() + 64;
    s[0] = a;
    s[1] = b;
    s[2] = d;
  }
  __syncthreads();
  a = b = d = 0.0;
  (void)warp;
  for (int w = 0; w < v.warps; ++w) {
    const double* s = pcg_scratch(v, c, w) + 64;
    a += s[0];
    b += s[1];
    d += s[2];
  }
}

// Fixed-order totals of slots [slot0, slot0 + count) over all CTAs; every thread of every CTA gets the same bits.
__device__ __forceinline__ void pcg_totals(const V2View& v, const V4Ctx& c, const double* red, int slot0, int count, double* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* s_tot = pcg_scratch(v, c, 0) + 72;
  __syncthreads();
  if (warp < count) {
    double acc = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) acc += __ldcg(red + b * 8 + slot0 + warp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_tot[warp] = acc;
  }
  __syncthreads();
  for (int k = 0; k < count; ++k) out[k] = s_tot[k];
}

// z_j = (M^-1 r)_j for the entries this CTA owns: r is exchanged through the warps' scratch.
__device__ __forceinline__ double pcg_precondition(const V2View& v, const V4Ctx& c, int precond, bool ok, bool owner_lane,
                                                   double rj, const double (&mrow)[9]) {
  if (precond == 0) return rj;
  const int tid = threadIdx.x;
  __syncthreads();
  if (owner_lane) pcg_scratch(v, c, tid >> 6)[tid & 63] = ok ? rj : 0.0;
  __syncthreads();
  double zj = 0.0;
  if (ok) {
    const int base = 9 * (tid / 9);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int i = base + k;
      zj += mrow[k] * pcg_scratch(v, c, i >> 6)[i & 63];
    }
  }
  return zj;
}

// The tile loop as a real call: it needs the whole register file for itself, and the call boundary makes the compiler
// park the (few) values that live across a product on the stack once per product instead of demoting the tile loop's
// own arrays to local memory.
template <bool kOwned>
__device__ __noinline__ uint32_t pcg_product(const V2View& v, const double* ete_inv, V4Ctx c, uint32_t flip) {
  v4_tiles<kOwned>(v, ete_inv, c, flip);
  return flip;
}

// Loop state shared by the threads of a CTA (every CTA holds identical values): kept out of the register file, which
// the tile loop needs entirely.
struct PcgShared {
  double rho, last_rho, Q0, tol_r, norm_rhs, norm_r, pq;
  int it, term, reason, fin_iter, abort;
};

template <bool kOwned>
__global__ void __launch_bounds__(kV4MaxThreads, 1) pcg_kernel(const __grid_constant__ PcgArgs a) {
  __shared__ PcgShared s;
  const V2View& v = a.v;
  const V4Ctx c = v4_ctx(v);
  const int tid = threadIdx.x;
  const int j = blockIdx.x * a.cams_per_cta * 9 + tid;
  const bool owner_lane = tid < 9 * a.cams_per_cta;
  const bool ok = owner_lane && j < a.prm.n;
  const int range = 9 * (c.cr.y - c.cr.x);
  const size_t range0 = 9 * static_cast<size_t>(c.cr.x);
  unsigned bar_target = 0;
  double* red_mine = a.red + blockIdx.x * 8;

  uint32_t flip = 0;  // phase parities of this warp's ring slots
  v4_init(v, c);
  v4_prime(v, a.ete_inv, c);
  for (int i = tid; i < c.sy_stride * v.replicas; i += blockDim.x) c.sy()[i] = 0.0;

  // ------------------------------------------------------------------ begin: x = 0, r = b, z = M^-1 r, p = z
  // (per-entry operands are never kept in registers across a product; the owner re-reads its few values from L1/L2
  // while it waits at the barrier)
  bool running = true;
  {
    double bj = 0.0, rj = 0.0;
    double mrow[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) mrow[k] = 0.0;
    if (ok) {
      bj = a.rhs[j];
      if (a.precond != 0) {
        const double* m = a.minv + 9 * static_cast<size_t>(j);  // row (j % 9) of block (j / 9): 81 * cam + 9 * row
#pragma unroll
        for (int k = 0; k < 9; ++k) mrow[k] = m[k];
      }
      rj = bj;
      a.x[j] = 0.0;
      a.r[j] = rj;
      a.qa[j] = 0.0;
      a.qb[j] = 0.0;
    }
    const double zj = pcg_precondition(v, c, a.precond, ok, owner_lane, rj, mrow);
    double accQ = 0.0, accR = ok ? rj * rj : 0.0, accRho = ok ? rj * zj : 0.0;
    if (ok) a.z[j] = zj;
    pcg_block_sum3(v, c, accQ, accR, accRho);
    if (tid == 0) {
      red_mine[2] = accQ;
      red_mine[3] = accR;
      red_mine[4] = accRho;
    }
    const bool alive = pcg_grid_barrier(a.barrier, bar_target, &s.abort);
    double tot[3];
    pcg_totals(v, c, a.red, 2, 3, tot);
    const double norm_r = sqrt(tot[1]), rho_new = tot[2];
    const double tol_r = a.prm.r_tolerance * norm_r;
    if (tid == 0) {
      s.norm_rhs = norm_r;
      s.tol_r = tol_r;
      s.norm_r = norm_r;
      s.pq = 0.0;
      s.Q0 = 0.0;
      s.last_rho = 1.0;
      s.rho = rho_new;
      s.it = 0;
      s.term = 1;
      s.reason = 0;
      s.fin_iter = 0;
    }
    if (!alive) {
      running = false;
    } else if (norm_r == 0.0 || (a.prm.min_iterations == 0 && norm_r <= tol_r)) {
      running = false;
      if (tid == 0) {
        s.term = 0;
        s.reason = norm_r == 0.0 ? 8 : 2;
      }
    } else if (zero_or_inf(rho_new) || isnan(rho_new)) {
      running = false;
      if (tid == 0) {
        s.term = 2;
        s.reason = 4;
        s.fin_iter = 1;
      }
    } else {
      double seed = 0.0, d1 = 0.0, d2 = 0.0;
      if (ok) {
        const double dj = a.Df != nullptr ? a.Df[j] : 0.0;
        a.p[j] = zj;
        seed = dj * dj * zj * zj;
      }
      pcg_block_sum3(v, c, seed, d1, d2);
      if (tid == 0) red_mine[1] = seed;
      for (int i = tid; i < range; i += blockDim.x) c.sx()[i] = __ldcg(a.z + range0 + i);
    }
  }

  // ------------------------------------------------------------------ iterations
  while (running) {
    __syncthreads();  // staged p, zeroed private vectors, loop state
    const int it = s.it + 1;
    double* qcur = (it & 1) ? a.qa : a.qb;
    PCG_STAMP(0);
    // ---- q = S0 p (without the D_f^2 p term), p.q partial
    flip = pcg_product<kOwned>(v, a.ete_inv, c, flip);
    PCG_STAMP(1);
    v4_big_points(v, c, a.ete_inv);
    v4_prime(v, a.ete_inv, c);  // the next product's first tiles stream in during the flush and the vector phases
    __syncthreads();
    {
      double pq = 0.0, d1 = 0.0, d2 = 0.0;
      for (int i = tid; i < range; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < v.replicas; ++r) {
          acc += c.sy()[r * c.sy_stride + i];
          c.sy()[r * c.sy_stride + i] = 0.0;
        }
        if (acc != 0.0) red_add(qcur + range0 + i, acc);
        pq += c.sx()[i] * acc;
      }
      pcg_block_sum3(v, c, pq, d1, d2);
      if (tid == 0) red_mine[0] = pq;
    }
    PCG_STAMP(2);
    // own entries: fetched before the barrier
    double bj = 0.0, xj = 0.0, rj = 0.0, pj = 0.0, dj = 0.0;
    double mrow[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) mrow[k] = 0.0;
    if (ok) {
      bj = a.rhs[j];
      xj = a.x[j];
      rj = a.r[j];
      pj = a.p[j];
      if (a.Df != nullptr) dj = a.Df[j];
      if (a.precond != 0) {
        const double* m = a.minv + 9 * static_cast<size_t>(j);
#pragma unroll
        for (int k = 0; k < 9; ++k) mrow[k] = m[k];
      }
    }
    if (!pcg_grid_barrier(a.barrier, bar_target, &s.abort)) break;

    PCG_STAMP(3);
    // ---- alpha (the owner's q entry is requested before the totals so that both L2 round trips overlap)
    const double q_raw = ok ? __ldcg(qcur + j) : 0.0;
    const double rho = s.rho;
    double alpha;
    {
      double t[2];
      pcg_totals(v, c, a.red, 0, 2, t);
      const double pq = t[0] + t[1];
      if (tid == 0) s.pq = pq;
      if (!(pq > 0.0) || isinf(pq)) {
        if (tid == 0) {
          s.term = isnan(pq) ? 2 : 1;
          s.reason = 6;
          s.fin_iter = it;
        }
        break;
      }
      alpha = rho / pq;
      if (isinf(alpha)) {
        if (tid == 0) {
          s.term = 2;
          s.reason = 7;
          s.fin_iter = it;
        }
        break;
      }
    }
    double* qnext = (it & 1) ? a.qb : a.qa;
    const bool reset = a.reset_period > 0 && (it % a.reset_period) == 0;
    if (!reset) {
      if (ok) {
        const double qj = q_raw + dj * dj * pj;
        xj += alpha * pj;
        a.x[j] = xj;
        rj -= alpha * qj;
        a.r[j] = rj;
        qnext[j] = 0.0;
      }
    } else {
      // r = b - S x with the updated x: one more product, on x
      if (ok) {
        xj += alpha * pj;
        a.x[j] = xj;
        a.tmp[j] = 0.0;
      }
      if (!pcg_grid_barrier(a.barrier, bar_target, &s.abort)) break;
      for (int i = tid; i < range; i += blockDim.x) c.sx()[i] = __ldcg(a.x + range0 + i);
      __syncthreads();
      flip = pcg_product<kOwned>(v, a.ete_inv, c, flip);
      v4_big_points(v, c, a.ete_inv);
      v4_prime(v, a.ete_inv, c);
      __syncthreads();
      for (int i = tid; i < range; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < v.replicas; ++r) {
          acc += c.sy()[r * c.sy_stride + i];
          c.sy()[r * c.sy_stride + i] = 0.0;
        }
        if (acc != 0.0) red_add(a.tmp + range0 + i, acc);
      }
      if (!pcg_grid_barrier(a.barrier, bar_target, &s.abort)) break;
      if (ok) {
        rj = bj - (__ldcg(a.tmp + j) + dj * dj * xj);
        a.r[j] = rj;
        qnext[j] = 0.0;
      }
    }
    PCG_STAMP(4);
    const double zj = pcg_precondition(v, c, a.precond, ok, owner_lane, rj, mrow);
    {
      double accQ = ok ? xj * (bj + rj) : 0.0, accR = ok ? rj * rj : 0.0, accRho = ok ? rj * zj : 0.0;
      if (ok) a.z[j] = zj;
      pcg_block_sum3(v, c, accQ, accR, accRho);
      if (tid == 0) {
        red_mine[2] = accQ;
        red_mine[3] = accR;
        red_mine[4] = accRho;
      }
    }
    PCG_STAMP(5);
    if (!pcg_grid_barrier(a.barrier, bar_target, &s.abort)) break;

    PCG_STAMP(6);
    // ---- termination tests of iteration `it` (identical in every CTA), beta, p
    // (z of the camera range is requested before the totals: the round trips overlap)
    const bool z_pre = range <= 2 * static_cast<int>(blockDim.x);
    double z_pre0 = 0.0, z_pre1 = 0.0;
    if (z_pre) {
      if (tid < range) z_pre0 = __ldcg(a.z + range0 + tid);
      if (tid + static_cast<int>(blockDim.x) < range) z_pre1 = __ldcg(a.z + range0 + tid + blockDim.x);
    }
    double tot[3];
    pcg_totals(v, c, a.red, 2, 3, tot);
    const double norm_r = sqrt(tot[1]), rho_new = tot[2];
    const double Q1 = -tot[0];
    const double zeta = it * (Q1 - s.Q0) / Q1;
    int term = -1, reason = 0, fin_iter = it;
    if (zeta < a.prm.q_tolerance && it >= a.prm.min_iterations) {
      term = 0;
      reason = 1;
    } else if (norm_r <= s.tol_r && it >= a.prm.min_iterations) {
      term = 0;
      reason = 2;
    } else if (it >= a.prm.max_iterations) {
      term = 1;
      reason = 3;
    } else if (zero_or_inf(rho_new) || isnan(rho_new)) {
      term = 2;
      reason = 4;
      fin_iter = it + 1;
    }
    const double beta = rho_new / rho;
    if (term < 0 && zero_or_inf(beta)) {
      term = 2;
      reason = 5;
      fin_iter = it + 1;
    }
    __syncthreads();  // everybody has read the old loop state
    if (tid == 0) {
      s.norm_r = norm_r;
      if (term >= 0) {
        s.term = term;
        s.reason = reason;
        s.fin_iter = fin_iter;
      } else {
        s.Q0 = Q1;
        s.last_rho = rho;
        s.rho = rho_new;
        s.it = it;
      }
    }
    if (term >= 0) break;
    {
      // staged copy of p for this CTA's camera range (the old p comes back from memory after a reset product)
      if (reset) {
        for (int i = tid; i < range; i += blockDim.x) c.sx()[i] = __ldcg(a.z + range0 + i) + beta * __ldcg(a.p + range0 + i);
        if (!pcg_grid_barrier(a.barrier, bar_target, &s.abort)) break;  // everybody has read the old p before its owners overwrite it
      } else if (z_pre) {
        if (tid < range) c.sx()[tid] = z_pre0 + beta * c.sx()[tid];
        if (tid + static_cast<int>(blockDim.x) < range) c.sx()[tid + blockDim.x] = z_pre1 + beta * c.sx()[tid + blockDim.x];
      } else {
        for (int i = tid; i < range; i += blockDim.x) c.sx()[i] = __ldcg(a.z + range0 + i) + beta * c.sx()[i];
      }
      double seed = 0.0, d1 = 0.0, d2 = 0.0;
      if (ok) {
        pj = zj + beta * pj;
        a.p[j] = pj;
        seed = dj * dj * pj * pj;
      }
      pcg_block_sum3(v, c, seed, d1, d2);
      if (tid == 0) red_mine[1] = seed;
    }
    PCG_STAMP(7);
  }

  v4_drain(v, c, flip);
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) {
    CgState* st = a.st;
    st->rho = s.rho;
    st->last_rho = s.last_rho;
    st->Q0 = s.Q0;
    st->norm_rhs = s.norm_rhs;
    st->tol_r = s.tol_r;
    st->norm_r = s.norm_r;
    st->pq = s.pq;
    st->iteration = s.fin_iter;
    st->done = 1;
    const unsigned aborted = ld_acquire_u32(a.barrier + 1);
    st->termination = aborted ? 3 : s.term;          // B200_LS_FATAL_ERROR
    st->reason = aborted ? 1000 + static_cast<int>(aborted) : s.reason;
  }
}

}  // namespace b200
