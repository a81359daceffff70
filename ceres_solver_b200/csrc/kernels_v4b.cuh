// Once-per-LM-iteration kernels on the v4 machinery (kernels_v2.cuh: every operand of a warp tile through the warp's
// TMA slot, the CTA's cameras addressed through its camera list, per-warp private camera vectors, segmented sums by
// shuffles): the implicit-Schur initialisation fused with the per-row 2x2 blocks the camera-major block-diagonal pass
// needs, and that camera-major pass itself.  Together they are the "Schur eliminate" of the metric
// (SchurEliminator<2,3,9>::Eliminate against a block-diagonal lhs, schur_eliminator_impl.h:184-568, plus
// ImplicitSchurComplement::Init / UpdateRhs, implicit_schur_complement.cc:49-97, :251-276).
#pragma once
#include "kernels_v2b.cuh"

namespace b200 {

struct InitV4Args {
  const double* b;   // [2N]
  const double* D;   // [3P+9C] or null
  double* ete_inv;   // [6P]
  double* rhs;       // [9C], zeroed by the caller
  double* ye;        // [3P] or null
  double* q3;        // [kQStride N] or null: Q_r = I - E_r (E'E + D^2)^-1 E_r'  (2x2 symmetric: q00, q01, q11, pad) per row
};

// slot: [0,4608) F | [4608,6144) E | [6144,6656) b (16 B per row) | [6656,7440) the tile's slice of D_e (24 B per point, the
// 16-byte-aligned superset) | [7680,7840) descriptor block
__device__ __forceinline__ void init_v4_issue(const V2View& v, const InitV4Args& a, unsigned char* stage, uint64_t* bar, int tile,
                                              int row_begin, int pt_begin, int row_count, int pt_count) {
  const uint32_t xoff = (pt_begin & 1) ? 8u : 0u;
  const uint32_t dbytes = a.D != nullptr ? ((24u * pt_count + xoff + 15u) & ~15u) : 0u;
  mbar_arrive_expect_tx(bar, row_count * 208u + dbytes + kV4MetaWords * 4u);
  bulk_g2s(stage, v.p.F() + 18 * static_cast<size_t>(row_begin), row_count * 144u, bar);
  bulk_g2s(stage + 4608, v.p.E() + 6 * static_cast<size_t>(row_begin), row_count * 48u, bar);
  bulk_g2s(stage + 6144, a.b + 2 * static_cast<size_t>(row_begin), row_count * 16u, bar);
  if (a.D != nullptr)
    bulk_g2s(stage + 6656, reinterpret_cast<const unsigned char*>(a.D + 3 * static_cast<size_t>(pt_begin)) - xoff, dbytes, bar);
  bulk_g2s(stage + 7680, v.tile_meta + static_cast<size_t>(kV4MetaWords) * tile, kV4MetaWords * 4u, bar);
}

template <int K>
__device__ __forceinline__ void seg_suffix_sum(double (&w)[K], int seg_end, int maxlen) {
  const int lane = threadIdx.x & 31;
  for (int d = 1; d < maxlen; d <<= 1) {
    double a[K];
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = __shfl_down_sync(0xffffffffu, w[k], d);
    if (lane + d < seg_end) {
#pragma unroll
      for (int k = 0; k < K; ++k) w[k] += a[k];
    }
  }
}

// The 33..kTile-row points of the CTA for the implicit-Schur initialisation (staging as in schur_mul_big_points_impl): the
// nine sums of E'E and E'b go through a CTA reduction, every row then finishes like a warp-tile row; the camera part is
// added to replica 0 of the private camera vector with shared-memory atomics.
__device__ __forceinline__ void init_big_points_impl(const V2View& v, const BigStage& st, uint32_t& parity, double* sy_rep0, int2 cr,
                                                     const InitV4Args& a) {
  const int2 br = v.cta_big[blockIdx.x];
  const int tid = threadIdx.x;
  double* sU = st.sU;
  for (int b = br.x; b < br.y; ++b) {
    const TileDesc d = v.big_tiles[b];
    if (tid == 0) {
      mbar_arrive_expect_tx(st.bar, d.obs_count * 192u);
      for (int r0 = 0, k = 0; r0 < d.obs_count; r0 += st.chunk_rows, ++k) {
        const int rows = min(st.chunk_rows, d.obs_count - r0);
        unsigned char* dst = st.base + static_cast<size_t>(k) * st.chunk_stride;
        bulk_g2s(dst, v.p.F() + 18 * static_cast<size_t>(d.obs_begin + r0), rows * 144u, st.bar);
        bulk_g2s(dst + st.chunk_rows * 144, v.p.E() + 6 * static_cast<size_t>(d.obs_begin + r0), rows * 48u, st.bar);
      }
    }
    const bool active = tid < d.obs_count;
    const int chunk = tid / st.chunk_rows, rr = tid - chunk * st.chunk_rows;
    const double* sF = reinterpret_cast<const double*>(st.base + static_cast<size_t>(chunk) * st.chunk_stride) + rr * 18;
    const double* sE = reinterpret_cast<const double*>(st.base + static_cast<size_t>(chunk) * st.chunk_stride + st.chunk_rows * 144) + rr * 6;
    const size_t row = static_cast<size_t>(d.obs_begin) + tid;
    const size_t pt = static_cast<size_t>(d.pt_begin);
    int cam_l = 0;
    double2 bb = make_double2(0, 0);
    if (active) {
      cam_l = meta_local(v, __ldg(v.row_meta + row), cr);
      bb = *reinterpret_cast<const double2*>(a.b + 2 * row);
    }
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    if (a.D != nullptr) {
      d0 = __ldg(a.D + 3 * pt);
      d1 = __ldg(a.D + 3 * pt + 1);
      d2 = __ldg(a.D + 3 * pt + 2);
    }
    mbar_wait(st.bar, parity);
    parity ^= 1;
    double f[18];
    double2 e0 = make_double2(0, 0), e1 = e0, e2 = e0;
    double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double2 w = lds2(sF + 2 * k);
        f[2 * k] = w.x;
        f[2 * k + 1] = w.y;
      }
      e0 = lds2(sE);
      e1 = lds2(sE + 2);
      e2 = lds2(sE + 4);
      m[0] = e0.x * e0.x + e1.y * e1.y;
      m[1] = e0.x * e0.y + e1.y * e2.x;
      m[2] = e0.x * e1.x + e1.y * e2.y;
      m[3] = e0.y * e0.y + e2.x * e2.x;
      m[4] = e0.y * e1.x + e2.x * e2.y;
      m[5] = e1.x * e1.x + e2.y * e2.y;
      m[6] = e0.x * bb.x + e1.y * bb.y;
      m[7] = e0.y * bb.x + e2.x * bb.y;
      m[8] = e1.x * bb.x + e2.y * bb.y;
    }
    if (tid < kTile) {  // the first four warps hold all rows
#pragma unroll
      for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
      }
      if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) sU[(tid >> 5) * 9 + k] = m[k];
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int k = 0; k < 9; ++k) m[k] = (sU[k] + sU[9 + k]) + (sU[18 + k] + sU[27 + k]);
      m[0] += d0 * d0;
      m[3] += d1 * d1;
      m[5] += d2 * d2;
      double inv[6];
      invert_sym3_llt(m, inv);
      const double v0 = inv[0] * m[6] + inv[1] * m[7] + inv[2] * m[8];
      const double v1 = inv[1] * m[6] + inv[3] * m[7] + inv[4] * m[8];
      const double v2 = inv[2] * m[6] + inv[4] * m[7] + inv[5] * m[8];
      if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) a.ete_inv[6 * pt + k] = inv[k];
        if (a.ye != nullptr) {
          a.ye[3 * pt] = v0;
          a.ye[3 * pt + 1] = v1;
          a.ye[3 * pt + 2] = v2;
        }
      }
      const double t0 = bb.x - (e0.x * v0 + e0.y * v1 + e1.x * v2);
      const double t1 = bb.y - (e1.y * v0 + e2.x * v1 + e2.y * v2);
      double* yc = sy_rep0 + 9 * cam_l;
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(yc + k, f[k] * t0 + f[9 + k] * t1);
      if (a.q3 != nullptr) {
        const double pa = inv[0] * e0.x + inv[1] * e0.y + inv[2] * e1.x, pb = inv[1] * e0.x + inv[3] * e0.y + inv[4] * e1.x,
                     pc = inv[2] * e0.x + inv[4] * e0.y + inv[5] * e1.x;
        const double pd = inv[0] * e1.y + inv[1] * e2.x + inv[2] * e2.y, pe = inv[1] * e1.y + inv[3] * e2.x + inv[4] * e2.y,
                     pf = inv[2] * e1.y + inv[4] * e2.x + inv[5] * e2.y;
        double2* q = reinterpret_cast<double2*>(a.q3 + kQStride * row);
        q[0] = make_double2(1.0 - (e0.x * pa + e0.y * pb + e1.x * pc), -(e1.y * pa + e2.x * pb + e2.y * pc));
        q[1] = make_double2(1.0 - (e1.y * pd + e2.x * pe + e2.y * pf), 0.0);
      }
    }
    __syncthreads();  // staging and sU are reused by the next point
  }
}

// ete_inv[k] = (sum_rows E'E + D_k^2)^-1 ; ye = ete_inv E'b ; rhs += F'(b - E ye) ; q3[r] = I - E_r ete_inv E_r'
// Warp tiles and the CTA's 33..kTile-row points; the slices of larger points go through huge_schur_init_kernel and
// row_q_tiles_kernel.
template <bool kOwned>
__global__ void __launch_bounds__(kV4MaxThreads, 1) schur_init_v4_kernel(V2View v, InitV4Args a) {
  const V4Ctx c = v4_ctx(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int2 part = c.part, cr = c.cr;
  v4_init(v, c);
  if (lane == 0) {
    int t = part.x + warp;
    for (int s = 0; s < v.stages && t < part.y; ++s, t += v.warps) {
      const WarpTile wt = v.wtiles[t];
      init_v4_issue(v, a, c.wbase() + s * kV4StageBytes, c.bars() + s, t, wt.row_begin, wt.pt_begin, wt.row_count, wt.pt_count);
    }
  }
  {
    const int n = c.sy_stride * v.replicas;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c.sy()[i] = 0.0;
  }
  __syncthreads();
  double* my_y = c.sy() + (kOwned ? warp : warp % v.replicas) * c.sy_stride;
  const int reissue = v.warps * v.stages;
  int it = 0;
  for (int tile = part.x + warp; tile < part.y; tile += v.warps, ++it) {
    const int s = it % v.stages;
    const uint32_t parity = (it / v.stages) & 1u;
    unsigned char* stage = c.wbase() + s * kV4StageBytes;
    const double* sF = reinterpret_cast<const double*>(stage);
    const double* sE = reinterpret_cast<const double*>(stage + 4608);
    const double* sB = reinterpret_cast<const double*>(stage + 6144);
    const uint32_t* sM = reinterpret_cast<const uint32_t*>(stage + 7680);
    mbar_wait(c.bars() + s, parity);
    const uint4 own = *reinterpret_cast<const uint4*>(sM + 32);
    const uint4 nxt = *reinterpret_cast<const uint4*>(sM + 36);
    const int row_begin = static_cast<int>(own.x), pt_begin = static_cast<int>(own.y);
    const int row_count = static_cast<int>(own.z & 0xffffu);
    const bool active = lane < row_count;
    const uint32_t meta = active ? sM[lane] : 0u;
    const int cam_l = meta_local(v, meta, cr);
    const Seg sg = v2_segment(active && meta_head(meta), row_count);
    double f[18];
    double2 e0 = make_double2(0, 0), e1 = e0, e2 = e0, bb = e0;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    if (active) {
      const double* fr = sF + lane * 18;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double2 w = lds2(fr + 2 * k);
        f[2 * k] = w.x;
        f[2 * k + 1] = w.y;
      }
      e0 = lds2(sE + lane * 6);
      e1 = lds2(sE + lane * 6 + 2);
      e2 = lds2(sE + lane * 6 + 4);
      bb = lds2(sB + 2 * lane);
      if (a.D != nullptr) {
        const double* dp = reinterpret_cast<const double*>(stage + 6656 + ((pt_begin & 1) ? 8 : 0)) + 3 * sg.lpt;
        d0 = dp[0];
        d1 = dp[1];
        d2 = dp[2];
      }
    }
    __syncwarp();  // every lane is done with the ring slot
    if (lane == 0 && (nxt.z & 0xffffu) != 0u)
      init_v4_issue(v, a, stage, c.bars() + s, tile + reissue, static_cast<int>(nxt.x), static_cast<int>(nxt.y),
                    static_cast<int>(nxt.z & 0xffffu), static_cast<int>(nxt.z >> 16));
    // E'E (6 unique) and E'b (3) of the row, summed over the rows of its point
    double m[9];
    m[0] = e0.x * e0.x + e1.y * e1.y;
    m[1] = e0.x * e0.y + e1.y * e2.x;
    m[2] = e0.x * e1.x + e1.y * e2.y;
    m[3] = e0.y * e0.y + e2.x * e2.x;
    m[4] = e0.y * e1.x + e2.x * e2.y;
    m[5] = e1.x * e1.x + e2.y * e2.y;
    m[6] = e0.x * bb.x + e1.y * bb.y;
    m[7] = e0.y * bb.x + e2.x * bb.y;
    m[8] = e1.x * bb.x + e2.y * bb.y;
    seg_suffix_sum<9>(m, sg.end, static_cast<int>(own.w));
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = __shfl_sync(0xffffffffu, m[k], sg.first);
    double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
      m[0] += d0 * d0;
      m[3] += d1 * d1;
      m[5] += d2 * d2;
      double inv[6];
      invert_sym3_llt(m, inv);   // every lane of the point computes the same inverse (the FP64 pipe is idle anyway)
      const double v0 = inv[0] * m[6] + inv[1] * m[7] + inv[2] * m[8];
      const double v1 = inv[1] * m[6] + inv[3] * m[7] + inv[4] * m[8];
      const double v2 = inv[2] * m[6] + inv[4] * m[7] + inv[5] * m[8];
      const size_t pt = static_cast<size_t>(pt_begin + sg.lpt);
      if (lane == sg.first) {
        double2* pi = reinterpret_cast<double2*>(a.ete_inv + 6 * pt);
        pi[0] = make_double2(inv[0], inv[1]);
        pi[1] = make_double2(inv[2], inv[3]);
        pi[2] = make_double2(inv[4], inv[5]);
        if (a.ye != nullptr) {
          a.ye[3 * pt] = v0;
          a.ye[3 * pt + 1] = v1;
          a.ye[3 * pt + 2] = v2;
        }
      }
      const double t0 = bb.x - (e0.x * v0 + e0.y * v1 + e1.x * v2);
      const double t1 = bb.y - (e1.y * v0 + e2.x * v1 + e2.y * v2);
#pragma unroll
      for (int k = 0; k < 9; ++k) g[k] = f[k] * t0 + f[9 + k] * t1;
      if (a.q3 != nullptr) {
        // P e_r' for the two rows of E
        const double pa = inv[0] * e0.x + inv[1] * e0.y + inv[2] * e1.x, pb = inv[1] * e0.x + inv[3] * e0.y + inv[4] * e1.x,
                     pc = inv[2] * e0.x + inv[4] * e0.y + inv[5] * e1.x;
        const double pd = inv[0] * e1.y + inv[1] * e2.x + inv[2] * e2.y, pe = inv[1] * e1.y + inv[3] * e2.x + inv[4] * e2.y,
                     pf = inv[2] * e1.y + inv[4] * e2.x + inv[5] * e2.y;
        double2* q = reinterpret_cast<double2*>(a.q3 + kQStride * (static_cast<size_t>(row_begin) + lane));
        q[0] = make_double2(1.0 - (e0.x * pa + e0.y * pb + e1.x * pc), -(e1.y * pa + e2.x * pb + e2.y * pc));
        q[1] = make_double2(1.0 - (e1.y * pd + e2.x * pe + e2.y * pf), 0.0);
      }
    }
    if (kOwned) cam_accumulate9_owned(my_y, cam_l, active, g);
    else cam_accumulate9(my_y, cam_l, active, g);
  }
  {  // the CTA's 33..kTile-row points (uniform per CTA), processed by the whole CTA
    const int2 br = v.cta_big[blockIdx.x];
    if (br.y > br.x) {
      double* sw0 = reinterpret_cast<double*>(c.ring() + v.stages * kV4StageBytes);
      BigStage st;
      st.base = c.ring();
      st.chunk_rows = kV4BigChunkRows;
      st.chunk_stride = v.per_warp_bytes;
      st.sU = sw0 + 48;   // 4 warps x 9 partial sums
      st.bar = reinterpret_cast<uint64_t*>(c.ring() + v4_extra_offset(v.stages));
      __syncthreads();  // every warp is done with its ring slot
      uint32_t parity = 0;
      init_big_points_impl(v, st, parity, c.sy(), cr, a);
    }
  }
  v2_epilogue(v, c.sy(), cr, a.rhs);
}

// Q_r for the rows of CTA tiles that hold ONE point each (the 33..kTile-row points and the slices of larger ones).
__global__ void __launch_bounds__(kTile) row_q_tiles_kernel(ProblemView p, const double* __restrict__ ete_inv, double* q3) {
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const TileDesc d = p.tiles[tile];
    if (static_cast<int>(threadIdx.x) >= d.obs_count) continue;
    const size_t r = static_cast<size_t>(d.obs_begin) + threadIdx.x;
    const double2* ep = reinterpret_cast<const double2*>(p.E() + 6 * r);
    const double2 a0 = __ldg(ep), a1 = __ldg(ep + 1), a2 = __ldg(ep + 2);
    const double e00 = a0.x, e01 = a0.y, e02 = a1.x, e10 = a1.y, e11 = a2.x, e12 = a2.y;
    const double* pi = ete_inv + 6 * static_cast<size_t>(d.pt_begin);
    const double p0 = __ldg(pi), p1 = __ldg(pi + 1), p2 = __ldg(pi + 2), p3 = __ldg(pi + 3), p4 = __ldg(pi + 4), p5 = __ldg(pi + 5);
    const double a = p0 * e00 + p1 * e01 + p2 * e02, b = p1 * e00 + p3 * e01 + p4 * e02, c = p2 * e00 + p4 * e01 + p5 * e02;
    const double dd = p0 * e10 + p1 * e11 + p2 * e12, e = p1 * e10 + p3 * e11 + p4 * e12, f = p2 * e10 + p4 * e11 + p5 * e12;
    double2* q = reinterpret_cast<double2*>(q3 + kQStride * r);
    q[0] = make_double2(1.0 - (e00 * a + e01 * b + e02 * c), -(e10 * a + e11 * b + e12 * c));
    q[1] = make_double2(1.0 - (e10 * dd + e11 * e + e12 * f), 0.0);
  }
}

// ------------------------------------------------------------------------------------------------
// Camera-major block diagonal, second version (cam_blocks_kernel of kernels_v2b.cuh: ncu showed 8 resident warps per SM at
// 162 registers, 43 % long-scoreboard stalls, and 450 shuffles per item for the final reduction).  One warp per item (a
// slice of one camera's row list), 4-warp CTAs (three per SM at ~165 registers: 12 resident warps keep more loads in
// flight than one 8-warp CTA did); the 45 packed entries are reduced across the lanes by recursive halving (each lane ends with <= 2 entries: 46 64-bit exchanges
// instead of 225) and added with <= 2 REDs per lane.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

constexpr int kCamBlkRowBytes = 144 + 32;                 // F row + Q block (q00, q01, q11, pad)
constexpr int kCamBlkDepth = 3;                           // buffers per warp: data of two steps in flight behind the one computed
constexpr int kCamBlkWarpBytes = kCamBlkDepth * 32 * kCamBlkRowBytes;
constexpr int kCamBlkThreads = 128;

// Sum of the 46 (zero padded) per-lane entries over the lanes by recursive halving: 46 -> 23 -> 12 -> 6 -> 3 -> 2 entries per
// lane (46 64-bit exchanges instead of 45 x 5), then <= 2 REDs per lane into the camera's packed block.
__device__ __forceinline__ void cam_block_flush(double (&m)[46], double* dst) {
  const int lane = threadIdx.x & 31;
  int base = 0;
  double r1[23], r2[12], r3[6], r4[3], r5[2];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 23; ++k) {
      const double keep = up ? m[23 + k] : m[k], give = up ? m[k] : m[23 + k];
      r1[k] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
    base += up ? 23 : 0;
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const double hi = (12 + k < 23) ? r1[(12 + k < 23) ? 12 + k : 0] : 0.0;
      const double keep = up ? hi : r1[k], give = up ? r1[k] : hi;
      r2[k] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
    base += up ? 12 : 0;
  }
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double keep = up ? r2[6 + k] : r2[k], give = up ? r2[k] : r2[6 + k];
      r3[k] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    base += up ? 6 : 0;
  }
  {
    const bool up = (lane & 2) != 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double keep = up ? r3[3 + k] : r3[k], give = up ? r3[k] : r3[3 + k];
      r4[k] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
    }
    base += up ? 3 : 0;
  }
  {
    const bool up = (lane & 1) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double hi = (k == 0) ? r4[2] : 0.0;
      const double keep = up ? hi : r4[k], give = up ? r4[k] : hi;
      r5[k] = keep + __shfl_xor_sync(0xffffffffu, give, 1);
    }
    base += up ? 2 : 0;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (r5[k] != 0.0) red_add(dst + base + k, r5[k]);   // padding entries are exactly zero and never reach an index >= 45
}

// Camera-major block diagonal.  One warp walks a strided sequence of items (slices of one camera's row list) as ONE
// continuous stream of 32-row steps: each lane copies ITS (gathered) row of step s+2 into the warp's shared-memory ring with
// cp.async while the warp computes step s, and the row indices of step s+3 are already being fetched -- no load latency is
// exposed, neither inside an item nor between items (ncu on the register version: 43 % long-scoreboard stalls at 8-12
// resident warps).  The 45 packed entries stay in registers per lane and are reduced across the lanes when an item ends.
template <bool kSchur>
__global__ void __launch_bounds__(kCamBlkThreads, 3)
    cam_blocks_v2_kernel(ProblemView p, int num_items, const CamItem* __restrict__ items, const int* __restrict__ cam_rows,
                         const double* __restrict__ q3, double* out45) {
  extern __shared__ __align__(128) unsigned char cb_smem[];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int stride = gridDim.x * warps_per_block;
  unsigned char* wbuf = cb_smem + (threadIdx.x >> 5) * kCamBlkWarpBytes;
  struct Cursor {
    int item, iter, begin, end;
  };
  auto load_item = [&](Cursor& c) {
    c.iter = 0;
    if (c.item < num_items) {
      const CamItem it = items[c.item];
      c.begin = it.begin;
      c.end = it.end;
    } else {
      c.begin = c.end = 0;
    }
  };
  auto advance = [&](Cursor& c) {
    ++c.iter;
    if (c.begin + 32 * c.iter >= c.end) {
      c.item += stride;
      load_item(c);
    }
  };
  auto row_of = [&](const Cursor& c) -> int {
    const int j = c.begin + 32 * c.iter + lane;
    return (c.item < num_items && j < c.end) ? __ldg(cam_rows + j) : -1;
  };
  auto issue = [&](int r, int buf) {
    if (r >= 0) {
      unsigned char* dst = wbuf + buf * 32 * kCamBlkRowBytes + lane * 144;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(p.F() + 18 * static_cast<size_t>(r));
#pragma unroll
      for (int k = 0; k < 9; ++k) cp_async16(dst + 16 * k, src + 16 * k);
      if (kSchur) {
        unsigned char* dq = wbuf + buf * 32 * kCamBlkRowBytes + 32 * 144 + lane * 32;
        const double* sq = q3 + kQStride * static_cast<size_t>(r);
        cp_async16(dq, sq);
        cp_async16(dq + 16, sq + 2);
      }
    }
    cp_async_commit();
  };
  Cursor cp, ix;
  cp.item = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  load_item(cp);
  ix = cp;
  int r_cur = row_of(ix);
  advance(ix);
  int r_nxt = row_of(ix);
  advance(ix);
  int r_pend = row_of(ix);
  advance(ix);
  issue(r_cur, 0);
  issue(r_nxt, 1);
  double m[46];
#pragma unroll
  for (int k = 0; k < 46; ++k) m[k] = 0.0;
  int stage = 0;
  while (cp.item < num_items) {
    const int r_far = row_of(ix);   // row indices three steps ahead: their latency hides behind this step
    advance(ix);
    issue(r_pend, stage == 0 ? 2 : stage - 1);   // data two steps ahead
    cp_async_wait<2>();
    __syncwarp();
    if (r_cur >= 0) {
      const double* fr = reinterpret_cast<const double*>(wbuf + stage * 32 * kCamBlkRowBytes + lane * 144);
      double f[18];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double2 w = lds2(fr + 2 * k);
        f[2 * k] = w.x;
        f[2 * k + 1] = w.y;
      }
      double q00 = 1.0, q01 = 0.0, q11 = 1.0;
      if (kSchur) {
        const double* q = reinterpret_cast<const double*>(wbuf + stage * 32 * kCamBlkRowBytes + 32 * 144 + lane * 32);
        q00 = q[0];
        q01 = q[1];
        q11 = q[2];
      }
      int idx = 0;
#pragma unroll
      for (int aa = 0; aa < 9; ++aa) {
        const double ga = q00 * f[aa] + q01 * f[9 + aa], gb = q01 * f[aa] + q11 * f[9 + aa];   // row aa of F'Q
#pragma unroll
        for (int bb = aa; bb < 9; ++bb) {
          m[idx] += ga * f[bb] + gb * f[9 + bb];
          ++idx;
        }
      }
    }
    if (cp.begin + 32 * (cp.iter + 1) >= cp.end) {   // last step of the item (uniform over the warp)
      cam_block_flush(m, out45 + 45 * static_cast<size_t>(items[cp.item].cam));
#pragma unroll
      for (int k = 0; k < 46; ++k) m[k] = 0.0;
    }
    __syncwarp();   // the buffer is refilled by the issue of the next step
    advance(cp);
    r_cur = r_nxt;
    r_nxt = r_pend;
    r_pend = r_far;
    stage = stage == 2 ? 0 : stage + 1;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Camera-major block diagonal, CTA-local (third version).  The gather version above reads every F row exactly once but in
// camera-major order, i.e. as 144-byte pieces scattered over the whole array: measured 2.9 TB/s at best, whatever the
// latency hiding (register, cp.async double buffer, continuous stream: 42-48 us on Ladybug-1723).  Here every persistent CTA
// STREAMS its own contiguous row range (the same partition as the warp-tile kernels) through shared memory in chunks of up to
// 512 rows -- two bulk copies per chunk, DRAM sees a linear read -- and the regrouping by camera happens on chip: the host
// sorted the rows of every chunk by camera once (b200_create), a quarter-warp takes one (camera, chunk) segment, its eight
// lanes walk the segment's rows in shared memory with the 45 packed entries in registers, reduce them by recursive halving
// over the eight lanes and add the result to the CTA's private accumulator of that camera (plain read-modify-write: a
// segment has exactly one owner and chunks are separated by a CTA barrier).  One flush of <= span x 45 REDs per CTA.
// ------------------------------------------------------------------------------------------------
struct CbChunk {
  int row_begin, row_count;   // rows of the chunk (<= kCbChunkRows)
  int seg_begin, seg_count;   // its (camera, chunk) segments in CbView::segs, longest first (<= kCbMaxSegs)
  int slot_begin;             // its sorted row slots in CbView::slots (multiple of 8)
  int pad0, pad1, pad2;
};
struct CbView {
  const CbChunk* chunks;
  const int2* cta_chunks;     // per CTA: [begin, end) into chunks
  const uint2* segs;          // x = camera position in the CTA's list | rows << 16 ; y = offset into the chunk's slot list
  const unsigned short* slots;
};
constexpr int kCbChunkRows = 512;
constexpr int kCbMaxSegs = 128;
constexpr int kCbThreads = 256;
constexpr int kCbBufBytes = kCbChunkRows * 144 + kCbChunkRows * 32 + kCbMaxSegs * 8 + kCbChunkRows * 2;
__host__ __device__ inline size_t cb3_acc_bytes(int max_cam_span) { return (static_cast<size_t>(45) * max_cam_span * 8 + 127) & ~static_cast<size_t>(127); }
__host__ __device__ inline size_t cb3_smem_bytes(int max_cam_span) { return cb3_acc_bytes(max_cam_span) + 2 * kCbBufBytes + 64; }

template <bool kSchur>
__global__ void __launch_bounds__(kCbThreads, 1)
    cam_blocks_v3_kernel(V2View v, CbView cb, const double* __restrict__ q3, double* out45) {
  extern __shared__ __align__(128) unsigned char cb_smem[];
  double* sacc = reinterpret_cast<double*>(cb_smem);
  unsigned char* bufs = cb_smem + cb3_acc_bytes(v.max_cam_span);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bufs + 2 * kCbBufBytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int2 cr = v.cta_cam[blockIdx.x];
  const int2 cc = cb.cta_chunks[blockIdx.x];
  const int nacc = 45 * v2_span(v, cr);
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) sacc[i] = 0.0;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int k) {   // thread 0: chunk k of this CTA -> buffer k & 1
    const CbChunk ch = cb.chunks[cc.x + k];
    unsigned char* buf = bufs + (k & 1) * kCbBufBytes;
    uint64_t* bar = bars + (k & 1);
    const uint32_t fb = ch.row_count * 144u, qb = kSchur ? ch.row_count * 32u : 0u;
    const uint32_t sb = ((ch.seg_count * 8u) + 15u) & ~15u, lb = ((ch.row_count * 2u) + 15u) & ~15u;
    mbar_arrive_expect_tx(bar, fb + qb + sb + lb);
    bulk_g2s(buf, v.p.F() + 18 * static_cast<size_t>(ch.row_begin), fb, bar);
    if (kSchur) bulk_g2s(buf + kCbChunkRows * 144, q3 + kQStride * static_cast<size_t>(ch.row_begin), qb, bar);
    bulk_g2s(buf + kCbChunkRows * 176, cb.segs + ch.seg_begin, sb, bar);
    bulk_g2s(buf + kCbChunkRows * 176 + kCbMaxSegs * 8, cb.slots + ch.slot_begin, lb, bar);
  };
  const int nchunks = cc.y - cc.x;
  if (threadIdx.x == 0) {
    if (nchunks > 0) issue(0);
    if (nchunks > 1) issue(1);
  }
  const int quarter = lane >> 3, sub = lane & 7;
  const int base = ((lane & 4) ? 24 : 0) + ((lane & 2) ? 12 : 0) + ((lane & 1) ? 6 : 0);
  for (int k = 0; k < nchunks; ++k) {
    const int seg_count = cb.chunks[cc.x + k].seg_count;   // (L2 hit; in flight while the chunk lands)
    const unsigned char* buf = bufs + (k & 1) * kCbBufBytes;
    const double* sF = reinterpret_cast<const double*>(buf);
    const double* sQ = reinterpret_cast<const double*>(buf + kCbChunkRows * 144);
    const uint2* sSeg = reinterpret_cast<const uint2*>(buf + kCbChunkRows * 176);
    const unsigned short* sSlot = reinterpret_cast<const unsigned short*>(buf + kCbChunkRows * 176 + kCbMaxSegs * 8);
    mbar_wait(bars + (k & 1), (k >> 1) & 1);
    for (int s0 = 0; s0 < seg_count; s0 += 4 * (kCbThreads / 32)) {
      const int sidx = s0 + warp * 4 + quarter;
      int cam_l = 0, count = 0, slot0 = 0;
      if (sidx < seg_count) {
        const uint2 sg = sSeg[sidx];
        cam_l = static_cast<int>(sg.x & 0xffffu);
        count = static_cast<int>(sg.x >> 16);
        slot0 = static_cast<int>(sg.y);
      }
      int cmax = count;
#pragma unroll
      for (int o = 16; o >= 8; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
      double m[48];
#pragma unroll
      for (int q = 0; q < 48; ++q) m[q] = 0.0;
      for (int i = sub; i < cmax; i += 8) {
        if (i < count) {
          const int slot = sSlot[slot0 + i];
          const double* fr = sF + slot * 18;
          double f[18];
#pragma unroll
          for (int q = 0; q < 9; ++q) {
            const double2 w = lds2(fr + 2 * q);
            f[2 * q] = w.x;
            f[2 * q + 1] = w.y;
          }
          double q00 = 1.0, q01 = 0.0, q11 = 1.0;
          if (kSchur) {
            const double2 qa = lds2(sQ + slot * 4), qb2 = lds2(sQ + slot * 4 + 2);
            q00 = qa.x;
            q01 = qa.y;
            q11 = qb2.x;
          }
          int idx = 0;
#pragma unroll
          for (int aa = 0; aa < 9; ++aa) {
            const double ga = q00 * f[aa] + q01 * f[9 + aa], gb = q01 * f[aa] + q11 * f[9 + aa];   // row aa of F'Q
#pragma unroll
            for (int bb = aa; bb < 9; ++bb) {
              m[idx] += ga * f[bb] + gb * f[9 + bb];
              ++idx;
            }
          }
        }
      }
      // recursive halving over the eight lanes of the quarter: 48 -> 24 -> 12 -> 6 entries per lane
      double r1[24], r2[12], r3[6];
      {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int q = 0; q < 24; ++q) {
          const double keep = up ? m[24 + q] : m[q], give = up ? m[q] : m[24 + q];
          r1[q] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
        }
      }
      {
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int q = 0; q < 12; ++q) {
          const double keep = up ? r1[12 + q] : r1[q], give = up ? r1[q] : r1[12 + q];
          r2[q] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
        }
      }
      {
        const bool up = (lane & 1) != 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double keep = up ? r2[6 + q] : r2[q], give = up ? r2[q] : r2[6 + q];
          r3[q] = keep + __shfl_xor_sync(0xffffffffu, give, 1);
        }
      }
      if (sidx < seg_count) {
        double* acc = sacc + 45 * cam_l + base;
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if (base + q < 45) acc[q] += r3[q];
      }
    }
    __syncthreads();   // every warp is done with the buffer (and with this chunk's accumulator updates)
    if (threadIdx.x == 0 && k + 2 < nchunks) issue(k + 2);
  }
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    const double acc = sacc[i];
    if (acc != 0.0) red_add(out45 + v2_global_entry(v, cr, i, 45), acc);
  }
}

}  // namespace b200
