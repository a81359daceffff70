// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// CPU restatement of the reference's Schur-complement linear-solver path:
//   PartitionedView            internal/ceres/partitioned_matrix_view_impl.h:50-658
//   ImplicitSchur              internal/ceres/implicit_schur_complement.cc:49-276
//   SchurEliminator            internal/ceres/schur_eliminator_impl.h:87-721
//   RandomAccessLhs            internal/ceres/block_random_access_{dense,diagonal}_matrix.cc
//   ConjugateGradients         internal/ceres/conjugate_gradients_solver.h:109-306
//   IterativeSchurSolve        internal/ceres/iterative_schur_complement_solver.cc:64-157
//   DenseSchurSolve            internal/ceres/schur_complement_solver.cc:101-159 (Dense variant)
//   PowerSeriesExpansion       internal/ceres/power_series_expansion_preconditioner.cc:57-82,
//                              implicit_schur_complement.cc:146-174
// Parity: pinned (tests/test_oracle_kat.py, tests/test_oracle_ba.py: the reference's fixtures, known-answer tests and
// published per-iteration transcripts).
// Templated on <row, e, f> block sizes exactly like the reference
// (schur_eliminator.cc:56-135 instantiations); kDyn == Eigen::Dynamic.
#pragma once
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "block_sparse.h"

namespace orc {

enum Termination { SUCCESS = 0, NO_CONVERGENCE = 1, FAILURE = 2, FATAL_ERROR = 3 };
enum PreconditionerType { IDENTITY = 0, JACOBI = 1, SCHUR_JACOBI = 2, SCHUR_POWER_SERIES_EXPANSION = 3 };  // include/ceres/types.h:93-119

// ----------------------------------------------------------------------------
// PartitionedMatrixView<kR,kE,kF>
// ----------------------------------------------------------------------------
template <int kR, int kE, int kF>
struct PartitionedView {
  const BlockSparseMatrix& A;
  int num_col_blocks_e, num_col_blocks_f, num_row_blocks_e = 0, num_cols_e = 0, num_cols_f = 0;
  int num_threads;
  std::vector<long> e_cum, f_cum;  // cumulative nnz per transpose row (partition weights)

  // partitioned_matrix_view_impl.h:50-104
  PartitionedView(const BlockSparseMatrix& a, int num_eliminate_blocks, int nthreads)
      : A(a), num_col_blocks_e(num_eliminate_blocks), num_threads(nthreads) {
    num_col_blocks_f = static_cast<int>(A.bs.cols.size()) - num_col_blocks_e;
    for (const auto& row : A.bs.rows)
      if (row.cells[0].block_id < num_col_blocks_e) ++num_row_blocks_e;
    for (size_t c = 0; c < A.bs.cols.size(); ++c) {
      if (static_cast<int>(c) < num_col_blocks_e) num_cols_e += A.bs.cols[c].size;
      else num_cols_f += A.bs.cols[c].size;
    }
    const long e_base = 0;
    for (int c = 0; c < num_col_blocks_e; ++c) e_cum.push_back(A.tbs.rows[c].cumulative_nnz - e_base);
    const long f_base = num_col_blocks_e > 0 ? A.tbs.rows[num_col_blocks_e - 1].cumulative_nnz : 0;
    for (int c = num_col_blocks_e; c < num_col_blocks_e + num_col_blocks_f; ++c)
      f_cum.push_back(A.tbs.rows[c].cumulative_nnz - f_base);
  }

  // y += E x     (:113-137)
  void RightMultiplyAndAccumulateE(const double* x, double* y) const {
    const auto& bs = A.bs;
    const double* v = A.values.data();
    ParallelFor(0, num_row_blocks_e, num_threads, [&](int, int r) {
      const Cell& cell = bs.rows[r].cells[0];
      MatrixVectorMultiply<kR, kE, 1>(v + cell.position, bs.rows[r].block.size, bs.cols[cell.block_id].size,
                                      x + bs.cols[cell.block_id].position, y + bs.rows[r].block.position);
    });
  }

  // y += F x     (:140-191)
  void RightMultiplyAndAccumulateF(const double* x, double* y) const {
    const auto& bs = A.bs;
    const double* v = A.values.data();
    ParallelFor(0, num_row_blocks_e, num_threads, [&](int, int r) {
      const auto& cells = bs.rows[r].cells;
      for (size_t c = 1; c < cells.size(); ++c) {
        const Block& col = bs.cols[cells[c].block_id];
        MatrixVectorMultiply<kR, kF, 1>(v + cells[c].position, bs.rows[r].block.size, col.size,
                                        x + col.position - num_cols_e, y + bs.rows[r].block.position);
      }
    });
    ParallelFor(num_row_blocks_e, static_cast<int>(bs.rows.size()), num_threads, [&](int, int r) {
      for (const Cell& cell : bs.rows[r].cells) {
        const Block& col = bs.cols[cell.block_id];
        MatrixVectorMultiply<kDyn, kDyn, 1>(v + cell.position, bs.rows[r].block.size, col.size,
                                            x + col.position - num_cols_e, y + bs.rows[r].block.position);
      }
    });
  }

  // y += E' x    (:194-264)
  void LeftMultiplyAndAccumulateE(const double* x, double* y) const {
    const auto& bs = A.bs;
    const double* v = A.values.data();
    if (num_threads <= 1) {
      for (int r = 0; r < num_row_blocks_e; ++r) {
        const Cell& cell = bs.rows[r].cells[0];
        MatrixTransposeVectorMultiply<kR, kE, 1>(v + cell.position, bs.rows[r].block.size,
                                                 bs.cols[cell.block_id].size, x + bs.rows[r].block.position,
                                                 y + bs.cols[cell.block_id].position);
      }
      return;
    }
    const auto& tbs = A.tbs;
    ParallelForWeighted(0, num_col_blocks_e, num_threads, e_cum.data(), [&](int, int cb) {
      const CompressedRow& trow = tbs.rows[cb];
      for (const Cell& cell : trow.cells) {
        const Block& rb = tbs.cols[cell.block_id];
        MatrixTransposeVectorMultiply<kR, kE, 1>(v + cell.position, rb.size, trow.block.size, x + rb.position,
                                                 y + trow.block.position);
      }
    });
  }

  // y += F' x    (:267-375)
  void LeftMultiplyAndAccumulateF(const double* x, double* y) const {
    const auto& bs = A.bs;
    const double* v = A.values.data();
    if (num_threads <= 1) {
      for (int r = 0; r < num_row_blocks_e; ++r) {
        const auto& cells = bs.rows[r].cells;
        for (size_t c = 1; c < cells.size(); ++c) {
          const Block& col = bs.cols[cells[c].block_id];
          MatrixTransposeVectorMultiply<kR, kF, 1>(v + cells[c].position, bs.rows[r].block.size, col.size,
                                                   x + bs.rows[r].block.position, y + col.position - num_cols_e);
        }
      }
      for (size_t r = num_row_blocks_e; r < bs.rows.size(); ++r) {
        for (const Cell& cell : bs.rows[r].cells) {
          const Block& col = bs.cols[cell.block_id];
          MatrixTransposeVectorMultiply<kDyn, kDyn, 1>(v + cell.position, bs.rows[r].block.size, col.size,
                                                       x + bs.rows[r].block.position,
                                                       y + col.position - num_cols_e);
        }
      }
      return;
    }
    const auto& tbs = A.tbs;
    ParallelForWeighted(num_col_blocks_e, num_col_blocks_e + num_col_blocks_f, num_threads, f_cum.data(),
                        [&](int, int cb) {
                          const CompressedRow& trow = tbs.rows[cb];
                          size_t i = 0;
                          for (; i < trow.cells.size(); ++i) {
                            const Cell& cell = trow.cells[i];
                            if (cell.block_id >= num_row_blocks_e) break;
                            const Block& rb = tbs.cols[cell.block_id];
                            MatrixTransposeVectorMultiply<kR, kF, 1>(v + cell.position, rb.size, trow.block.size,
                                                                     x + rb.position,
                                                                     y + trow.block.position - num_cols_e);
                          }
                          for (; i < trow.cells.size(); ++i) {
                            const Cell& cell = trow.cells[i];
                            const Block& rb = tbs.cols[cell.block_id];
                            MatrixTransposeVectorMultiply<kDyn, kDyn, 1>(v + cell.position, rb.size,
                                                                         trow.block.size, x + rb.position,
                                                                         y + trow.block.position - num_cols_e);
                          }
                        });
  }

  // Positions of the diagonal cells of a block-diagonal matrix over column blocks
  // [start, end)   (:383-418)
  std::vector<int> DiagonalLayout(int start, int end, int* total) const {
    std::vector<int> pos;
    int p = 0;
    for (int c = start; c < end; ++c) {
      pos.push_back(p);
      p += A.bs.cols[c].size * A.bs.cols[c].size;
    }
    *total = p;
    return pos;
  }

  // out = blockdiag(E'E)   (:447-523)
  void UpdateBlockDiagonalEtE(const std::vector<int>& layout, double* out) const {
    const auto& tbs = A.tbs;
    const double* v = A.values.data();
    ParallelForWeighted(0, num_col_blocks_e, num_threads, e_cum.data(), [&](int, int cb) {
      const CompressedRow& trow = tbs.rows[cb];
      const int n = trow.block.size;
      double* cellv = out + layout[cb];
      for (int i = 0; i < n * n; ++i) cellv[i] = 0.0;
      for (const Cell& c : trow.cells) {
        const int nr = tbs.cols[c.block_id].size;
        MatrixTransposeMatrixMultiply<kR, kE, kR, kE, 1>(v + c.position, nr, n, v + c.position, nr, n, cellv, 0,
                                                         0, n, n);
      }
    });
  }

  // out = blockdiag(F'F)   (:531-658)
  void UpdateBlockDiagonalFtF(const std::vector<int>& layout, double* out) const {
    const auto& tbs = A.tbs;
    const double* v = A.values.data();
    ParallelForWeighted(num_col_blocks_e, num_col_blocks_e + num_col_blocks_f, num_threads, f_cum.data(),
                        [&](int, int cb) {
                          const CompressedRow& trow = tbs.rows[cb];
                          const int n = trow.block.size;
                          double* cellv = out + layout[cb - num_col_blocks_e];
                          for (int i = 0; i < n * n; ++i) cellv[i] = 0.0;
                          size_t i = 0;
                          for (; i < trow.cells.size(); ++i) {
                            const Cell& c = trow.cells[i];
                            if (c.block_id >= num_row_blocks_e) break;
                            const int nr = tbs.cols[c.block_id].size;
                            MatrixTransposeMatrixMultiply<kR, kF, kR, kF, 1>(v + c.position, nr, n,
                                                                             v + c.position, nr, n, cellv, 0, 0,
                                                                             n, n);
                          }
                          for (; i < trow.cells.size(); ++i) {
                            const Cell& c = trow.cells[i];
                            const int nr = tbs.cols[c.block_id].size;
                            MatrixTransposeMatrixMultiply<kDyn, kDyn, kDyn, kDyn, 1>(
                                v + c.position, nr, n, v + c.position, nr, n, cellv, 0, 0, n, n);
                          }
                        });
  }
};

// Block-diagonal matrix of square cells: values + per-block (size, vector position, cell position).
struct BlockDiagonal {
  std::vector<int> sizes, vec_pos, cell_pos;
  std::vector<double> values;
  // y += M x   (block_sparse_matrix.cc:239-274 on a block-diagonal BSM /
  //             block_random_access_diagonal_matrix.cc:102-116)
  void RightMultiplyAndAccumulate(const double* x, double* y, int num_threads) const {
    ParallelFor(0, static_cast<int>(sizes.size()), num_threads, [&](int, int i) {
      MatrixVectorMultiply<kDyn, kDyn, 1>(values.data() + cell_pos[i], sizes[i], sizes[i], x + vec_pos[i],
                                          y + vec_pos[i]);
    });
  }
};

// ----------------------------------------------------------------------------
// ImplicitSchurComplement
// ----------------------------------------------------------------------------
template <int kR, int kE, int kF>
struct ImplicitSchur {
  PartitionedView<kR, kE, kF> pmv;
  int num_threads;
  bool compute_ftf_inverse;
  const double* D = nullptr;
  const double* b = nullptr;
  BlockDiagonal ete_inv, ftf_inv;
  std::vector<double> rhs, tmp_rows, tmp_e_cols, tmp_e_cols_2, tmp_f_cols;

  ImplicitSchur(const BlockSparseMatrix& A, int num_eliminate_blocks, bool want_ftf, int nthreads)
      : pmv(A, num_eliminate_blocks, nthreads), num_threads(nthreads), compute_ftf_inverse(want_ftf) {
    SetupDiagonal(0, pmv.num_col_blocks_e, &ete_inv);
    if (compute_ftf_inverse)
      SetupDiagonal(pmv.num_col_blocks_e, pmv.num_col_blocks_e + pmv.num_col_blocks_f, &ftf_inv);
    rhs.assign(pmv.num_cols_f, 0.0);
    tmp_rows.assign(A.num_rows, 0.0);
    tmp_e_cols.assign(pmv.num_cols_e, 0.0);
    tmp_e_cols_2.assign(pmv.num_cols_e, 0.0);
    tmp_f_cols.assign(pmv.num_cols_f, 0.0);
  }

  void SetupDiagonal(int start, int end, BlockDiagonal* d) {
    int total = 0;
    d->cell_pos = pmv.DiagonalLayout(start, end, &total);
    d->values.assign(total, 0.0);
    int p = 0;
    for (int c = start; c < end; ++c) {
      d->sizes.push_back(pmv.A.bs.cols[c].size);
      d->vec_pos.push_back(p);
      p += pmv.A.bs.cols[c].size;
    }
  }

  int num_rows() const { return pmv.num_cols_f; }

  // implicit_schur_complement.cc:49-97
  void Init(const double* D_in, const double* b_in) {
    D = D_in;
    b = b_in;
    pmv.UpdateBlockDiagonalEtE(ete_inv.cell_pos, ete_inv.values.data());
    if (compute_ftf_inverse) pmv.UpdateBlockDiagonalFtF(ftf_inv.cell_pos, ftf_inv.values.data());
    AddDiagonalAndInvert(D, &ete_inv);
    if (compute_ftf_inverse) AddDiagonalAndInvert(D == nullptr ? nullptr : D + pmv.num_cols_e, &ftf_inv);
    UpdateRhs();
  }

  // :179-204   m += diag(D^2); m = llt(upper(m)).solve(I)
  void AddDiagonalAndInvert(const double* Dp, BlockDiagonal* m) {
    ParallelFor(0, static_cast<int>(m->sizes.size()), num_threads, [&](int, int i) {
      const int n = m->sizes[i];
      double* v = m->values.data() + m->cell_pos[i];
      if (Dp != nullptr)
        for (int k = 0; k < n; ++k) v[k * n + k] += Dp[m->vec_pos[i] + k] * Dp[m->vec_pos[i] + k];
      InvertSymmetricUpperLLT(n, v);
    });
  }

  // y = S x   (:106-144)   S = F'F + D_f^2 - F'E (E'E+D_e^2)^-1 E'F
  void RightMultiplyAndAccumulate(const double* x, double* y) {
    std::fill(tmp_rows.begin(), tmp_rows.end(), 0.0);
    pmv.RightMultiplyAndAccumulateF(x, tmp_rows.data());                         // y1 = F x
    std::fill(tmp_e_cols.begin(), tmp_e_cols.end(), 0.0);
    pmv.LeftMultiplyAndAccumulateE(tmp_rows.data(), tmp_e_cols.data());         // y2 = E' y1
    std::fill(tmp_e_cols_2.begin(), tmp_e_cols_2.end(), 0.0);
    ete_inv.RightMultiplyAndAccumulate(tmp_e_cols.data(), tmp_e_cols_2.data(), num_threads);
    for (double& t : tmp_e_cols_2) t = -t;                                       // y3 = -(E'E)^-1 y2
    pmv.RightMultiplyAndAccumulateE(tmp_e_cols_2.data(), tmp_rows.data());      // y1 += E y3
    const int n = num_rows();
    if (D != nullptr) {
      const double* Df = D + pmv.num_cols_e;
      for (int i = 0; i < n; ++i) y[i] = Df[i] * Df[i] * x[i];                   // y5 = D^2 x
    } else {
      for (int i = 0; i < n; ++i) y[i] = 0.0;
    }
    pmv.LeftMultiplyAndAccumulateF(tmp_rows.data(), y);                          // y = y5 + F' y1
  }

  // :146-174   y += (F'F + D_f^2)^-1 F'E (E'E + D_e^2)^-1 E'F x     (needs compute_ftf_inverse)
  void InversePowerSeriesOperatorRightMultiplyAccumulate(const double* x, double* y) {
    std::fill(tmp_rows.begin(), tmp_rows.end(), 0.0);
    pmv.RightMultiplyAndAccumulateF(x, tmp_rows.data());                         // y1 = F x
    std::fill(tmp_e_cols.begin(), tmp_e_cols.end(), 0.0);
    pmv.LeftMultiplyAndAccumulateE(tmp_rows.data(), tmp_e_cols.data());         // y2 = E' y1
    std::fill(tmp_e_cols_2.begin(), tmp_e_cols_2.end(), 0.0);
    ete_inv.RightMultiplyAndAccumulate(tmp_e_cols.data(), tmp_e_cols_2.data(), num_threads);  // y3 = (E'E)^-1 y2
    std::fill(tmp_rows.begin(), tmp_rows.end(), 0.0);
    pmv.RightMultiplyAndAccumulateE(tmp_e_cols_2.data(), tmp_rows.data());      // y1 = E y3
    std::fill(tmp_f_cols.begin(), tmp_f_cols.end(), 0.0);
    pmv.LeftMultiplyAndAccumulateF(tmp_rows.data(), tmp_f_cols.data());         // y4 = F' y1
    ftf_inv.RightMultiplyAndAccumulate(tmp_f_cols.data(), y, num_threads);       // y += (F'F)^-1 y4
  }

  // :208-243   y = [ (E'E)^-1 E'(b - F x) ; x ]
  void BackSubstitute(const double* x, double* y) {
    const int nr = pmv.A.num_rows;
    std::fill(tmp_rows.begin(), tmp_rows.end(), 0.0);
    if (x != nullptr) pmv.RightMultiplyAndAccumulateF(x, tmp_rows.data());
    for (int i = 0; i < nr; ++i) tmp_rows[i] = b[i] - tmp_rows[i];
    std::fill(tmp_e_cols.begin(), tmp_e_cols.end(), 0.0);
    pmv.LeftMultiplyAndAccumulateE(tmp_rows.data(), tmp_e_cols.data());
    for (int i = 0; i < pmv.A.num_cols; ++i) y[i] = 0.0;
    ete_inv.RightMultiplyAndAccumulate(tmp_e_cols.data(), y, num_threads);
    if (x != nullptr)
      for (int i = 0; i < pmv.num_cols_f; ++i) y[pmv.num_cols_e + i] = x[i];
  }

  // :251-276   rhs = F'(b - E (E'E)^-1 E' b)
  void UpdateRhs() {
    const int nr = pmv.A.num_rows;
    std::fill(tmp_e_cols.begin(), tmp_e_cols.end(), 0.0);
    pmv.LeftMultiplyAndAccumulateE(b, tmp_e_cols.data());
    std::fill(tmp_e_cols_2.begin(), tmp_e_cols_2.end(), 0.0);
    ete_inv.RightMultiplyAndAccumulate(tmp_e_cols.data(), tmp_e_cols_2.data(), num_threads);
    std::fill(tmp_rows.begin(), tmp_rows.end(), 0.0);
    pmv.RightMultiplyAndAccumulateE(tmp_e_cols_2.data(), tmp_rows.data());
    for (int i = 0; i < nr; ++i) tmp_rows[i] = b[i] - tmp_rows[i];
    std::fill(rhs.begin(), rhs.end(), 0.0);
    pmv.LeftMultiplyAndAccumulateF(tmp_rows.data(), rhs.data());
  }
};

// PowerSeriesExpansionPreconditioner::RightMultiplyAndAccumulate (power_series_expansion_preconditioner.cc:57-82):
// y = sum_{i=0..k} ((F'F)^-1 F'E (E'E)^-1 E'F)^i (F'F)^-1 x, k < max_num_spse_iterations, stopping early once a
// term's norm drops below spse_tolerance * |first term|.  y is overwritten.
template <int kR, int kE, int kF>
void PowerSeriesExpansion(ImplicitSchur<kR, kE, kF>* isc, int max_num_spse_iterations, double spse_tolerance,
                          const double* x, double* y) {
  const int n = isc->num_rows();
  std::vector<double> series_term(n), previous_series_term(n);
  for (int i = 0; i < n; ++i) y[i] = 0.0;
  isc->ftf_inv.RightMultiplyAndAccumulate(x, y, isc->num_threads);
  for (int i = 0; i < n; ++i) previous_series_term[i] = y[i];
  double sq = 0.0;
  for (int i = 0; i < n; ++i) sq += y[i] * y[i];
  const double norm_threshold = spse_tolerance * std::sqrt(sq);
  for (int i = 1;; ++i) {
    std::fill(series_term.begin(), series_term.end(), 0.0);
    isc->InversePowerSeriesOperatorRightMultiplyAccumulate(previous_series_term.data(), series_term.data());
    double tsq = 0.0;
    for (int k = 0; k < n; ++k) {
      y[k] += series_term[k];
      tsq += series_term[k] * series_term[k];
    }
    if (i >= max_num_spse_iterations || std::sqrt(tsq) < norm_threshold) break;
    std::swap(previous_series_term, series_term);
  }
}

// ----------------------------------------------------------------------------
// BlockRandomAccess{Dense,Diagonal}Matrix: the lhs the eliminator writes into.
// GetCell returns nullptr for cells the target does not store
// (block_random_access_diagonal_matrix.cc:62-81; dense: block_random_access_dense_matrix.cc).
// ----------------------------------------------------------------------------
struct RandomAccessLhs {
  bool diagonal_only;
  std::vector<int> sizes, layout;  // block sizes and their scalar offsets
  int n = 0;
  std::vector<double> values;      // dense: n*n row-major; diagonal: concatenated cells
  std::vector<int> cell_pos;       // diagonal only
  std::unique_ptr<std::mutex[]> locks;  // one per f block row (cell mutexes of the reference,
                                        // block_random_access_matrix.h:87-93, coarsened)
  RandomAccessLhs(const std::vector<int>& block_sizes, bool diag) : diagonal_only(diag), sizes(block_sizes) {
    int p = 0, cp = 0;
    for (int s : sizes) {
      layout.push_back(p);
      cell_pos.push_back(cp);
      p += s;
      cp += s * s;
    }
    n = p;
    values.assign(diag ? cp : static_cast<size_t>(n) * n, 0.0);
    locks.reset(new std::mutex[sizes.size() ? sizes.size() : 1]);
  }
  void SetZero() { std::fill(values.begin(), values.end(), 0.0); }
  double* GetCell(int rb, int cb, int* r, int* c, int* row_stride, int* col_stride) {
    if (diagonal_only) {
      if (rb != cb) return nullptr;
      *r = 0;
      *c = 0;
      *row_stride = sizes[rb];
      *col_stride = sizes[rb];
      return values.data() + cell_pos[rb];
    }
    *r = layout[rb];
    *c = layout[cb];
    *row_stride = n;
    *col_stride = n;
    return values.data();
  }
  // block_random_access_diagonal_matrix.cc:90-100
  void InvertDiagonal(int num_threads) {
    ParallelFor(0, static_cast<int>(sizes.size()), num_threads,
                [&](int, int i) { InvertSymmetricUpperLLT(sizes[i], values.data() + cell_pos[i]); });
  }
};

// ----------------------------------------------------------------------------
// SchurEliminator<kR,kE,kF>
// ----------------------------------------------------------------------------
template <int kR, int kE, int kF>
struct SchurEliminator {
  struct Chunk {
    int size = 0, start = 0;
    std::map<int, int> buffer_layout;  // f block id -> offset in the chunk buffer
  };
  int num_threads;
  int num_eliminate_blocks = 0;
  bool assume_full_rank_ete = false;
  int uneliminated_row_begins = 0;
  int buffer_size = 1;
  std::vector<Chunk> chunks;
  std::vector<int> lhs_row_layout;

  explicit SchurEliminator(int nthreads) : num_threads(nthreads) {}

  // schur_eliminator_impl.h:87-181
  void Init(int num_elim, bool full_rank_ete, const BlockStructure& bs) {
    num_eliminate_blocks = num_elim;
    assume_full_rank_ete = full_rank_ete;
    const int num_col_blocks = static_cast<int>(bs.cols.size());
    const int num_row_blocks = static_cast<int>(bs.rows.size());
    buffer_size = 1;
    chunks.clear();
    lhs_row_layout.assign(num_col_blocks - num_elim, 0);
    int lhs_num_rows = 0;
    for (int i = num_elim; i < num_col_blocks; ++i) {
      lhs_row_layout[i - num_elim] = lhs_num_rows;
      lhs_num_rows += bs.cols[i].size;
    }
    int r = 0;
    while (r < num_row_blocks) {
      const int chunk_block_id = bs.rows[r].cells.front().block_id;
      if (chunk_block_id >= num_elim) break;
      chunks.emplace_back();
      Chunk& chunk = chunks.back();
      chunk.start = r;
      int bsize = 0;
      const int e_block_size = bs.cols[chunk_block_id].size;
      while (r + chunk.size < num_row_blocks) {
        const CompressedRow& row = bs.rows[r + chunk.size];
        if (row.cells.front().block_id != chunk_block_id) break;
        for (size_t c = 1; c < row.cells.size(); ++c) {
          const Cell& cell = row.cells[c];
          if (chunk.buffer_layout.find(cell.block_id) == chunk.buffer_layout.end()) {
            chunk.buffer_layout[cell.block_id] = bsize;
            bsize += e_block_size * bs.cols[cell.block_id].size;
          }
        }
        buffer_size = std::max(bsize, buffer_size);
        ++chunk.size;
      }
      r += chunk.size;
    }
    uneliminated_row_begins = chunks.empty() ? 0 : chunks.back().start + chunks.back().size;
  }

  static void InvertPSD(bool full_rank, int n, const double* ete, double* inv) {
    // invert_psd_matrix.h:50-79: fixed 1..4 & full rank -> closed-form inverse; else LLT.
    // (The rank-deficient SVD pseudo-inverse branch is not on the BAL path: every
    //  caller on it passes D != 0 or assume_full_rank_ete = true.)
    if (full_rank && kE == 3 && n == 3) {
      Invert3x3(ete, inv);
      return;
    }
    if (full_rank && kE != kDyn && n == 1) {
      inv[0] = 1.0 / ete[0];
      return;
    }
    for (int i = 0; i < n * n; ++i) inv[i] = ete[i];
    InvertSymmetricUpperLLT(n, inv);
  }

  // :184-311
  void Eliminate(const BlockSparseMatrix& A, const double* b, const double* D, RandomAccessLhs* lhs,
                 double* rhs) {
    const BlockStructure& bs = A.bs;
    const int num_col_blocks = static_cast<int>(bs.cols.size());
    if (lhs->n > 0) {
      lhs->SetZero();
      if (rhs != nullptr)
        for (int i = 0; i < lhs->n; ++i) rhs[i] = 0.0;
    }
    if (D != nullptr) {
      for (int i = num_eliminate_blocks; i < num_col_blocks; ++i) {
        const int block_id = i - num_eliminate_blocks;
        int r, c, rs, cs;
        double* cell = lhs->GetCell(block_id, block_id, &r, &c, &rs, &cs);
        if (cell != nullptr) {
          const int n = bs.cols[i].size;
          for (int k = 0; k < n; ++k) cell[(r + k) * cs + c + k] += D[bs.cols[i].position + k] * D[bs.cols[i].position + k];
        }
      }
    }
    std::vector<double> buffers(static_cast<size_t>(buffer_size) * num_threads);
    std::vector<double> outer_buffers(static_cast<size_t>(buffer_size) * num_threads);
    ParallelFor(0, static_cast<int>(chunks.size()), num_threads, [&](int thread_id, int i) {
      double* buffer = buffers.data() + static_cast<size_t>(thread_id) * buffer_size;
      const Chunk& chunk = chunks[i];
      const int e_block_id = bs.rows[chunk.start].cells.front().block_id;
      const int e = bs.cols[e_block_id].size;
      for (int k = 0; k < buffer_size; ++k) buffer[k] = 0.0;
      double ete[16 * 16] = {0.0};
      if (D != nullptr)
        for (int k = 0; k < e; ++k) ete[k * e + k] = D[bs.cols[e_block_id].position + k] * D[bs.cols[e_block_id].position + k];
      double g[16] = {0.0};
      ChunkDiagonalBlockAndGradient(chunk, A, b, chunk.start, e, ete, g, buffer, lhs);
      double inverse_ete[16 * 16];
      InvertPSD(assume_full_rank_ete, e, ete, inverse_ete);
      if (rhs != nullptr) {
        double inverse_ete_g[16];
        MatrixVectorMultiply<kE, kE, 0>(inverse_ete, e, e, g, inverse_ete_g);
        UpdateRhs(chunk, A, b, chunk.start, inverse_ete_g, rhs, lhs);
      }
      ChunkOuterProduct(outer_buffers.data() + static_cast<size_t>(thread_id) * buffer_size, bs, e, inverse_ete,
                        buffer, chunk.buffer_layout, lhs);
    });
    NoEBlockRowsUpdate(A, b, uneliminated_row_begins, lhs, rhs);
  }

  // :314-380
  void BackSubstitute(const BlockSparseMatrix& A, const double* b, const double* D, const double* z, double* y) {
    const BlockStructure& bs = A.bs;
    const double* values = A.values.data();
    ParallelFor(0, static_cast<int>(chunks.size()), num_threads, [&](int, int i) {
      const Chunk& chunk = chunks[i];
      const int e_block_id = bs.rows[chunk.start].cells.front().block_id;
      const int e = bs.cols[e_block_id].size;
      double* y_ptr = y + bs.cols[e_block_id].position;
      double ete[16 * 16] = {0.0};
      if (D != nullptr)
        for (int k = 0; k < e; ++k) ete[k * e + k] = D[bs.cols[e_block_id].position + k] * D[bs.cols[e_block_id].position + k];
      for (int k = 0; k < e; ++k) y_ptr[k] = 0.0;
      for (int j = 0; j < chunk.size; ++j) {
        const CompressedRow& row = bs.rows[chunk.start + j];
        const Cell& e_cell = row.cells.front();
        double sj[16];
        for (int k = 0; k < row.block.size; ++k) sj[k] = b[row.block.position + k];
        for (size_t c = 1; c < row.cells.size(); ++c) {
          const int f_block_id = row.cells[c].block_id;
          MatrixVectorMultiply<kR, kF, -1>(values + row.cells[c].position, row.block.size, bs.cols[f_block_id].size,
                                           z + lhs_row_layout[f_block_id - num_eliminate_blocks], sj);
        }
        MatrixTransposeVectorMultiply<kR, kE, 1>(values + e_cell.position, row.block.size, e, sj, y_ptr);
        MatrixTransposeMatrixMultiply<kR, kE, kR, kE, 1>(values + e_cell.position, row.block.size, e,
                                                         values + e_cell.position, row.block.size, e, ete, 0, 0, e, e);
      }
      double inv[16 * 16], tmp[16];
      InvertPSD(assume_full_rank_ete, e, ete, inv);
      MatrixVectorMultiply<kE, kE, 0>(inv, e, e, y_ptr, tmp);
      for (int k = 0; k < e; ++k) y_ptr[k] = tmp[k];
    });
  }

 private:
  // :386-427
  void UpdateRhs(const Chunk& chunk, const BlockSparseMatrix& A, const double* b, int row_block_counter,
                 const double* inverse_ete_g, double* rhs, RandomAccessLhs* lhs) {
    const BlockStructure& bs = A.bs;
    const double* values = A.values.data();
    const int e_block_id = bs.rows[chunk.start].cells.front().block_id;
    const int e = bs.cols[e_block_id].size;
    int b_pos = bs.rows[row_block_counter].block.position;
    for (int j = 0; j < chunk.size; ++j) {
      const CompressedRow& row = bs.rows[row_block_counter + j];
      const Cell& e_cell = row.cells.front();
      double sj[16];
      for (int k = 0; k < row.block.size; ++k) sj[k] = b[b_pos + k];
      MatrixVectorMultiply<kR, kE, -1>(values + e_cell.position, row.block.size, e, inverse_ete_g, sj);
      for (size_t c = 1; c < row.cells.size(); ++c) {
        const int block_id = row.cells[c].block_id;
        const int block = block_id - num_eliminate_blocks;
        std::unique_lock<std::mutex> lk(lhs->locks[block], std::defer_lock);
        if (num_threads > 1) lk.lock();
        MatrixTransposeVectorMultiply<kR, kF, 1>(values + row.cells[c].position, row.block.size,
                                                 bs.cols[block_id].size, sj, rhs + lhs_row_layout[block]);
      }
      b_pos += row.block.size;
    }
  }

  // :449-512
  void ChunkDiagonalBlockAndGradient(const Chunk& chunk, const BlockSparseMatrix& A, const double* b,
                                     int row_block_counter, int e, double* ete, double* g, double* buffer,
                                     RandomAccessLhs* lhs) {
    const BlockStructure& bs = A.bs;
    const double* values = A.values.data();
    int b_pos = bs.rows[row_block_counter].block.position;
    for (int j = 0; j < chunk.size; ++j) {
      const CompressedRow& row = bs.rows[row_block_counter + j];
      if (row.cells.size() > 1) EBlockRowOuterProduct(A, row_block_counter + j, lhs);
      const Cell& e_cell = row.cells.front();
      MatrixTransposeMatrixMultiply<kR, kE, kR, kE, 1>(values + e_cell.position, row.block.size, e,
                                                       values + e_cell.position, row.block.size, e, ete, 0, 0, e, e);
      if (b != nullptr)
        MatrixTransposeVectorMultiply<kR, kE, 1>(values + e_cell.position, row.block.size, e, b + b_pos, g);
      for (size_t c = 1; c < row.cells.size(); ++c) {
        const int f_block_id = row.cells[c].block_id;
        const int f = bs.cols[f_block_id].size;
        double* buffer_ptr = buffer + chunk.buffer_layout.at(f_block_id);
        MatrixTransposeMatrixMultiply<kR, kE, kR, kF, 1>(values + e_cell.position, row.block.size, e,
                                                         values + row.cells[c].position, row.block.size, f,
                                                         buffer_ptr, 0, 0, e, f);
      }
      b_pos += row.block.size;
    }
  }

  // :519-568   S(i,j) -= b_i' ete^-1 b_j  for i <= j in the chunk
  void ChunkOuterProduct(double* b1_transpose_inverse_ete, const BlockStructure& bs, int e,
                         const double* inverse_ete, const double* buffer,
                         const std::map<int, int>& buffer_layout, RandomAccessLhs* lhs) {
    for (auto it1 = buffer_layout.begin(); it1 != buffer_layout.end(); ++it1) {
      const int block1 = it1->first - num_eliminate_blocks;
      const int block1_size = bs.cols[it1->first].size;
      MatrixTransposeMatrixMultiply<kE, kF, kE, kE, 0>(buffer + it1->second, e, block1_size, inverse_ete, e, e,
                                                       b1_transpose_inverse_ete, 0, 0, block1_size, e);
      for (auto it2 = it1; it2 != buffer_layout.end(); ++it2) {
        const int block2 = it2->first - num_eliminate_blocks;
        int r, c, rs, cs;
        double* cell = lhs->GetCell(block1, block2, &r, &c, &rs, &cs);
        if (cell != nullptr) {
          const int block2_size = bs.cols[it2->first].size;
          std::unique_lock<std::mutex> lk(lhs->locks[block1], std::defer_lock);
          if (num_threads > 1) lk.lock();
          MatrixMatrixMultiply<kF, kE, kE, kF, -1>(b1_transpose_inverse_ete, block1_size, e,
                                                   buffer + it2->second, e, block2_size, cell, r, c, rs, cs);
        }
      }
    }
  }

  // :575-606
  void NoEBlockRowsUpdate(const BlockSparseMatrix& A, const double* b, int row_block_counter,
                          RandomAccessLhs* lhs, double* rhs) {
    const BlockStructure& bs = A.bs;
    const double* values = A.values.data();
    for (; row_block_counter < static_cast<int>(bs.rows.size()); ++row_block_counter) {
      NoEBlockRowOuterProduct(A, row_block_counter, lhs);
      if (rhs == nullptr) continue;
      const CompressedRow& row = bs.rows[row_block_counter];
      for (const Cell& cell : row.cells) {
        const int block = cell.block_id - num_eliminate_blocks;
        MatrixTransposeVectorMultiply<kDyn, kDyn, 1>(values + cell.position, row.block.size,
                                                     bs.cols[cell.block_id].size, b + row.block.position,
                                                     rhs + lhs_row_layout[block]);
      }
    }
  }

  // :622-667
  void NoEBlockRowOuterProduct(const BlockSparseMatrix& A, int row_block_index, RandomAccessLhs* lhs) {
    RowOuterProduct<kDyn, kDyn>(A, row_block_index, 0, lhs, /*lock=*/false);
  }
  // :672-721
  void EBlockRowOuterProduct(const BlockSparseMatrix& A, int row_block_index, RandomAccessLhs* lhs) {
    RowOuterProduct<kR, kF>(A, row_block_index, 1, lhs, /*lock=*/num_threads > 1);
  }
  template <int kRr, int kFf>
  void RowOuterProduct(const BlockSparseMatrix& A, int row_block_index, int first_cell, RandomAccessLhs* lhs,
                       bool lock) {
    const BlockStructure& bs = A.bs;
    const double* values = A.values.data();
    const CompressedRow& row = bs.rows[row_block_index];
    for (size_t i = first_cell; i < row.cells.size(); ++i) {
      const int block1 = row.cells[i].block_id - num_eliminate_blocks;
      const int block1_size = bs.cols[row.cells[i].block_id].size;
      int r, c, rs, cs;
      double* cell = lhs->GetCell(block1, block1, &r, &c, &rs, &cs);
      if (cell != nullptr) {
        std::unique_lock<std::mutex> lk(lhs->locks[block1], std::defer_lock);
        if (lock) lk.lock();
        MatrixTransposeMatrixMultiply<kRr, kFf, kRr, kFf, 1>(values + row.cells[i].position, row.block.size,
                                                             block1_size, values + row.cells[i].position,
                                                             row.block.size, block1_size, cell, r, c, rs, cs);
      }
      for (size_t j = i + 1; j < row.cells.size(); ++j) {
        const int block2 = row.cells[j].block_id - num_eliminate_blocks;
        double* cell2 = lhs->GetCell(block1, block2, &r, &c, &rs, &cs);
        if (cell2 != nullptr) {
          const int block2_size = bs.cols[row.cells[j].block_id].size;
          std::unique_lock<std::mutex> lk(lhs->locks[block1], std::defer_lock);
          if (lock) lk.lock();
          MatrixTransposeMatrixMultiply<kRr, kFf, kRr, kFf, 1>(values + row.cells[i].position, row.block.size,
                                                               block1_size, values + row.cells[j].position,
                                                               row.block.size, block2_size, cell2, r, c, rs, cs);
        }
      }
    }
  }
};

// ----------------------------------------------------------------------------
// ConjugateGradientsSolver  (conjugate_gradients_solver.h:109-306), vector ops are
// the single-threaded Eigen expressions of eigen_vector_ops.h:46-101.
// ----------------------------------------------------------------------------
struct CGOptions {
  int min_num_iterations = 0;
  int max_num_iterations = 500;
  int residual_reset_period = 10;
  double r_tolerance = 0.0;
  double q_tolerance = 0.0;
};
struct LinearSummary {
  double residual_norm = -1.0;
  int num_iterations = -1;
  int termination_type = FATAL_ERROR;
  std::string message;
};

inline double Dot(const std::vector<double>& a, const std::vector<double>& b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}
inline double Norm(const std::vector<double>& a) { return std::sqrt(Dot(a, a)); }

template <typename Lhs, typename Precond>
LinearSummary ConjugateGradients(const CGOptions& options, Lhs&& lhs, const std::vector<double>& rhs,
                                 Precond&& preconditioner, std::vector<double>& solution) {
  auto IsZeroOrInfinity = [](double x) { return (x == 0.0) || std::isinf(x); };
  const size_t n = rhs.size();
  std::vector<double> p(n, 0.0), r(n, 0.0), z(n, 0.0), tmp(n, 0.0);
  LinearSummary summary;
  summary.termination_type = NO_CONVERGENCE;
  summary.message = "Maximum number of iterations reached.";
  summary.num_iterations = 0;

  const double norm_rhs = Norm(rhs);
  if (norm_rhs == 0.0) {
    std::fill(solution.begin(), solution.end(), 0.0);
    summary.termination_type = SUCCESS;
    summary.message = "Convergence. |b| = 0.";
    return summary;
  }
  const double tol_r = options.r_tolerance * norm_rhs;

  std::fill(tmp.begin(), tmp.end(), 0.0);
  lhs(solution.data(), tmp.data());
  for (size_t i = 0; i < n; ++i) r[i] = rhs[i] - tmp[i];
  double norm_r = Norm(r);
  if (options.min_num_iterations == 0 && norm_r <= tol_r) {
    summary.termination_type = SUCCESS;
    summary.message = "Convergence. |r| <= tol.";
    return summary;
  }

  double rho = 1.0;
  for (size_t i = 0; i < n; ++i) tmp[i] = rhs[i] + r[i];
  double Q0 = -Dot(solution, tmp);

  for (summary.num_iterations = 1;; ++summary.num_iterations) {
    std::fill(z.begin(), z.end(), 0.0);
    preconditioner(r.data(), z.data());

    const double last_rho = rho;
    rho = Dot(r, z);
    if (IsZeroOrInfinity(rho)) {
      summary.termination_type = FAILURE;
      summary.message = "Numerical failure. rho = r'z.";
      break;
    }
    if (summary.num_iterations == 1) {
      p = z;
    } else {
      const double beta = rho / last_rho;
      if (IsZeroOrInfinity(beta)) {
        summary.termination_type = FAILURE;
        summary.message = "Numerical failure. beta = rho_n / rho_{n-1}.";
        break;
      }
      for (size_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }

    std::vector<double>& q = z;
    std::fill(q.begin(), q.end(), 0.0);
    lhs(p.data(), q.data());
    const double pq = Dot(p, q);
    if ((pq <= 0) || std::isinf(pq)) {
      summary.termination_type = NO_CONVERGENCE;
      summary.message = "Matrix is indefinite, no more progress can be made.";
      break;
    }
    const double alpha = rho / pq;
    if (std::isinf(alpha)) {
      summary.termination_type = FAILURE;
      summary.message = "Numerical failure. alpha = rho / pq.";
      break;
    }
    for (size_t i = 0; i < n; ++i) solution[i] = solution[i] + alpha * p[i];

    if (summary.num_iterations % options.residual_reset_period == 0) {
      std::fill(tmp.begin(), tmp.end(), 0.0);
      lhs(solution.data(), tmp.data());
      for (size_t i = 0; i < n; ++i) r[i] = rhs[i] - tmp[i];
    } else {
      for (size_t i = 0; i < n; ++i) r[i] = r[i] - alpha * q[i];
    }

    for (size_t i = 0; i < n; ++i) tmp[i] = rhs[i] + r[i];
    const double Q1 = -Dot(solution, tmp);
    const double zeta = summary.num_iterations * (Q1 - Q0) / Q1;
    if (zeta < options.q_tolerance && summary.num_iterations >= options.min_num_iterations) {
      summary.termination_type = SUCCESS;
      summary.message = "Convergence: zeta < q_tolerance.";
      break;
    }
    Q0 = Q1;

    norm_r = Norm(r);
    if (norm_r <= tol_r && summary.num_iterations >= options.min_num_iterations) {
      summary.termination_type = SUCCESS;
      summary.message = "Convergence. |r| <= tol_r.";
      break;
    }
    if (summary.num_iterations >= options.max_num_iterations) break;
  }
  return summary;
}

// ----------------------------------------------------------------------------
// Linear solvers. Both solve  min |A x - b|^2 + |D x|^2  (linear_solver.h:237-256),
// A partitioned [E F] with num_eliminate_blocks leading column blocks.
// The solver objects persist across solves with identical sparsity
// (linear_solver.h:137-142).
// ----------------------------------------------------------------------------
struct LinearSolverBase {
  virtual ~LinearSolverBase() = default;
  virtual LinearSummary Solve(BlockSparseMatrix* A, const double* b, const double* D, double q_tolerance,
                              double r_tolerance, double* x) = 0;
};

struct IterativeSchurOptions {
  int num_eliminate_blocks = 0;
  int preconditioner_type = SCHUR_JACOBI;
  int min_num_iterations = 0;
  int max_num_iterations = 500;
  int residual_reset_period = 10;
  int num_threads = 1;
  int max_num_spse_iterations = 5;       // linear_solver.h:172
  bool use_spse_initialization = false;  // :177
  double spse_tolerance = 0.1;           // :183
};

template <int kR, int kE, int kF>
struct IterativeSchurSolver : LinearSolverBase {
  IterativeSchurOptions options;
  std::unique_ptr<ImplicitSchur<kR, kE, kF>> schur_complement;
  std::unique_ptr<SchurEliminator<kR, kE, kF>> eliminator;  // for SCHUR_JACOBI
  std::unique_ptr<RandomAccessLhs> m;
  std::vector<double> reduced_solution;
  explicit IterativeSchurSolver(const IterativeSchurOptions& o) : options(o) {}

  // iterative_schur_complement_solver.cc:64-157
  LinearSummary Solve(BlockSparseMatrix* A, const double* b, const double* D, double q_tolerance,
                      double r_tolerance, double* x) override {
    const int num_eliminate_blocks = options.num_eliminate_blocks;
    if (schur_complement == nullptr) {
      // compute_ftf_inverse: implicit_schur_complement.cc:61-64
      schur_complement.reset(new ImplicitSchur<kR, kE, kF>(
          *A, num_eliminate_blocks,
          options.use_spse_initialization || options.preconditioner_type == JACOBI ||
              options.preconditioner_type == SCHUR_POWER_SERIES_EXPANSION,
          options.num_threads));
    }
    schur_complement->Init(D, b);
    const int num_schur_complement_blocks = static_cast<int>(A->bs.cols.size()) - num_eliminate_blocks;
    if (num_schur_complement_blocks == 0) {
      LinearSummary summary;
      summary.num_iterations = 0;
      summary.termination_type = SUCCESS;
      schur_complement->BackSubstitute(nullptr, x);
      return summary;
    }
    reduced_solution.assign(schur_complement->num_rows(), 0.0);
    if (options.use_spse_initialization)  // iterative_schur_complement_solver.cc:100-111
      PowerSeriesExpansion(schur_complement.get(), options.max_num_spse_iterations, options.spse_tolerance,
                           schur_complement->rhs.data(), reduced_solution.data());

    // CreatePreconditioner (:159-199) + Update (:113-122)
    if (options.preconditioner_type == SCHUR_JACOBI) {
      if (eliminator == nullptr) {
        std::vector<int> sizes;
        for (size_t i = num_eliminate_blocks; i < A->bs.cols.size(); ++i) sizes.push_back(A->bs.cols[i].size);
        m.reset(new RandomAccessLhs(sizes, /*diag=*/true));
        eliminator.reset(new SchurEliminator<kR, kE, kF>(options.num_threads));
        eliminator->Init(num_eliminate_blocks, /*assume_full_rank_ete=*/true, A->bs);  // schur_jacobi_preconditioner.cc:79-84
      }
      eliminator->Eliminate(*A, nullptr, D, m.get(), nullptr);  // schur_jacobi_preconditioner.cc:87-97
      m->InvertDiagonal(options.num_threads);
    }

    CGOptions cg;
    cg.min_num_iterations = options.min_num_iterations;
    cg.max_num_iterations = options.max_num_iterations;
    cg.residual_reset_period = options.residual_reset_period;
    cg.q_tolerance = q_tolerance;
    cg.r_tolerance = r_tolerance;

    auto lhs = [&](const double* xx, double* yy) {
      // LinearOperatorAdapter: y += S x on a zeroed y; ISC::RightMultiplyAndAccumulate
      // overwrites y (it assigns y = D^2 x first), which is equivalent here.
      schur_complement->RightMultiplyAndAccumulate(xx, yy);
    };
    const int n = schur_complement->num_rows();
    auto precond = [&](const double* rr, double* zz) {
      switch (options.preconditioner_type) {
        case IDENTITY:
          for (int i = 0; i < n; ++i) zz[i] += rr[i];
          break;
        case JACOBI:
          schur_complement->ftf_inv.RightMultiplyAndAccumulate(rr, zz, options.num_threads);
          break;
        case SCHUR_POWER_SERIES_EXPANSION:  // :178-186: tolerance 0 keeps the preconditioner fixed during CG
          PowerSeriesExpansion(schur_complement.get(), options.max_num_spse_iterations, 0.0, rr, zz);
          break;
        default: {
          // apply m^-1 blockwise (block_random_access_diagonal_matrix.cc:102-116)
          ParallelFor(0, static_cast<int>(m->sizes.size()), options.num_threads, [&](int, int i) {
            MatrixVectorMultiply<kDyn, kDyn, 1>(m->values.data() + m->cell_pos[i], m->sizes[i], m->sizes[i],
                                                rr + m->layout[i], zz + m->layout[i]);
          });
        }
      }
    };
    LinearSummary summary = ConjugateGradients(cg, lhs, schur_complement->rhs, precond, reduced_solution);
    if (summary.termination_type != FAILURE && summary.termination_type != FATAL_ERROR)
      schur_complement->BackSubstitute(reduced_solution.data(), x);
    return summary;
  }
};

// DENSE_SCHUR: explicit S via the eliminator, dense Cholesky, back-substitute
// (schur_complement_solver.cc:101-159, DenseSchurComplementSolver :161-214). Stands in for
// SPARSE_SCHUR's CHOLMOD factorisation too: both are exact solves of the same reduced system.
template <int kR, int kE, int kF>
struct DenseSchurSolver : LinearSolverBase {
  int num_eliminate_blocks, num_threads;
  std::unique_ptr<SchurEliminator<kR, kE, kF>> eliminator;
  std::unique_ptr<RandomAccessLhs> lhs;
  std::vector<double> rhs;
  DenseSchurSolver(int num_elim, int nthreads) : num_eliminate_blocks(num_elim), num_threads(nthreads) {}

  LinearSummary Solve(BlockSparseMatrix* A, const double* b, const double* D, double, double, double* x) override {
    if (eliminator == nullptr) {
      std::vector<int> sizes;
      for (size_t i = num_eliminate_blocks; i < A->bs.cols.size(); ++i) sizes.push_back(A->bs.cols[i].size);
      lhs.reset(new RandomAccessLhs(sizes, /*diag=*/false));
      eliminator.reset(new SchurEliminator<kR, kE, kF>(num_threads));
      // SchurComplementSolver passes assume_full_rank_ete = false only when it has to cope with
      // rank-deficient E'E; with D != null (always on the LM path) E'E + D^2 is SPD, and the LLT
      // branch of InvertPSDMatrix is taken here.
      eliminator->Init(num_eliminate_blocks, /*assume_full_rank_ete=*/true, A->bs);
      rhs.assign(lhs->n, 0.0);
    }
    for (int i = 0; i < A->num_cols; ++i) x[i] = 0.0;
    eliminator->Eliminate(*A, b, D, lhs.get(), rhs.data());
    LinearSummary summary;
    summary.num_iterations = 0;
    summary.termination_type = SUCCESS;
    const int n = lhs->n;
    if (n > 0) {
      summary.num_iterations = 1;  // schur_complement_solver.cc:199
      // In-place dense Cholesky on the upper triangle (what selfadjointView<Upper>().llt() reads).
      std::vector<double>& S = lhs->values;
      std::vector<double> L(static_cast<size_t>(n) * n, 0.0);
      for (int j = 0; j < n; ++j) {
        double d = S[static_cast<size_t>(j) * n + j];
        for (int k = 0; k < j; ++k) d -= L[static_cast<size_t>(j) * n + k] * L[static_cast<size_t>(j) * n + k];
        if (!(d > 0.0)) {
          summary.termination_type = FAILURE;
          summary.message = "Eigen failure. Unable to perform dense Cholesky factorization.";
          return summary;
        }
        const double ljj = std::sqrt(d);
        L[static_cast<size_t>(j) * n + j] = ljj;
        ParallelFor(j + 1, n, (n - j > 256) ? num_threads : 1, [&](int, int i) {
          double s = S[static_cast<size_t>(j) * n + i];
          const double* Li = &L[static_cast<size_t>(i) * n];
          const double* Lj = &L[static_cast<size_t>(j) * n];
          for (int k = 0; k < j; ++k) s -= Li[k] * Lj[k];
          L[static_cast<size_t>(i) * n + j] = s / ljj;
        });
      }
      std::vector<double> y(n), z(n);
      for (int i = 0; i < n; ++i) {
        double s = rhs[i];
        for (int k = 0; k < i; ++k) s -= L[static_cast<size_t>(i) * n + k] * y[k];
        y[i] = s / L[static_cast<size_t>(i) * n + i];
      }
      for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= L[static_cast<size_t>(k) * n + i] * z[k];
        z[i] = s / L[static_cast<size_t>(i) * n + i];
      }
      const int num_cols_e = A->num_cols - n;
      for (int i = 0; i < n; ++i) x[num_cols_e + i] = z[i];
      eliminator->BackSubstitute(*A, b, D, z.data(), x);
    } else {
      eliminator->BackSubstitute(*A, b, D, nullptr, x);
    }
    return summary;
  }
};

}  // namespace orc
