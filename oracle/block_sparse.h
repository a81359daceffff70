// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// Block-row sparse matrix restating
//   internal/ceres/block_structure.h:52-177     (Block, Cell, CompressedRow, CompressedRowBlockStructure)
//   internal/ceres/block_sparse_matrix.cc:178-216 (ctor), :239-274 (y += A x),
//   :278-349 (y += A' x), :351-401 (SquaredColumnNorm), :403-450 (ScaleColumns),
//   :784-808 (transpose block structure)
// Cells hold their values row-major at values[cell.position].
#pragma once
#include <cstring>
#include <memory>
#include <vector>

#include "block_ops.h"
#include "parallel.h"

namespace orc {

struct Block {
  int size = 0;
  int position = 0;
};
struct Cell {
  int block_id = 0;
  int position = 0;
};
struct CompressedRow {
  Block block;
  std::vector<Cell> cells;
  long cumulative_nnz = 0;
};
struct BlockStructure {
  std::vector<Block> cols;
  std::vector<CompressedRow> rows;
};

struct BlockSparseMatrix {
  BlockStructure bs;
  BlockStructure tbs;  // transpose structure: rows = column blocks, cells = (row block id, position)
  std::vector<double> values;
  int num_rows = 0, num_cols = 0;
  long num_nonzeros = 0;

  // Finishes construction once bs.cols / bs.rows (with cell positions) are set.
  void Finalize() {
    num_rows = 0;
    num_cols = 0;
    num_nonzeros = 0;
    for (auto& c : bs.cols) num_cols += c.size;
    long cum = 0;
    for (auto& r : bs.rows) {
      num_rows += r.block.size;
      for (auto& c : r.cells) {
        const long n = static_cast<long>(r.block.size) * bs.cols[c.block_id].size;
        num_nonzeros += n;
        cum += n;
      }
      r.cumulative_nnz = cum;
    }
    values.assign(num_nonzeros, 0.0);
    BuildTranspose();
  }

  void BuildTranspose() {
    tbs = BlockStructure();
    tbs.rows.resize(bs.cols.size());
    for (size_t i = 0; i < bs.cols.size(); ++i) tbs.rows[i].block = bs.cols[i];
    tbs.cols.resize(bs.rows.size());
    for (size_t r = 0; r < bs.rows.size(); ++r) {
      tbs.cols[r] = bs.rows[r].block;
      for (auto& c : bs.rows[r].cells) {
        Cell t;
        t.block_id = static_cast<int>(r);
        t.position = c.position;
        tbs.rows[c.block_id].cells.push_back(t);
      }
    }
    long cum = 0;
    for (auto& r : tbs.rows) {
      for (auto& c : r.cells) cum += static_cast<long>(r.block.size) * tbs.cols[c.block_id].size;
      r.cumulative_nnz = cum;
    }
  }

  void SetZero() { std::memset(values.data(), 0, values.size() * sizeof(double)); }

  // y += A x   (block_sparse_matrix.cc:239-274; row-block parallel)
  void RightMultiplyAndAccumulate(const double* x, double* y, int num_threads) const {
    const double* v = values.data();
    ParallelFor(0, static_cast<int>(bs.rows.size()), num_threads, [&](int, int r) {
      const CompressedRow& row = bs.rows[r];
      for (const Cell& cell : row.cells) {
        const Block& col = bs.cols[cell.block_id];
        MatrixVectorMultiply<kDyn, kDyn, 1>(v + cell.position, row.block.size, col.size,
                                            x + col.position, y + row.block.position);
      }
    });
  }

  // y += A' x  (block_sparse_matrix.cc:278-349; serial walks rows, threaded walks the
  // transpose structure one column block per work item)
  void LeftMultiplyAndAccumulate(const double* x, double* y, int num_threads) const {
    const double* v = values.data();
    if (num_threads <= 1) {
      for (const CompressedRow& row : bs.rows) {
        for (const Cell& cell : row.cells) {
          const Block& col = bs.cols[cell.block_id];
          MatrixTransposeVectorMultiply<kDyn, kDyn, 1>(v + cell.position, row.block.size, col.size,
                                                       x + row.block.position, y + col.position);
        }
      }
      return;
    }
    ParallelFor(0, static_cast<int>(tbs.rows.size()), num_threads, [&](int, int c) {
      const CompressedRow& trow = tbs.rows[c];
      for (const Cell& cell : trow.cells) {
        const Block& rb = tbs.cols[cell.block_id];
        MatrixTransposeVectorMultiply<kDyn, kDyn, 1>(v + cell.position, rb.size, trow.block.size,
                                                     x + rb.position, y + trow.block.position);
      }
    });
  }

  // x[j] = sum_i A_ij^2   (block_sparse_matrix.cc:351-401)
  void SquaredColumnNorm(double* x, int num_threads) const {
    const double* v = values.data();
    if (num_threads <= 1) {
      std::memset(x, 0, sizeof(double) * num_cols);
      for (const CompressedRow& row : bs.rows) {
        for (const Cell& cell : row.cells) {
          const Block& col = bs.cols[cell.block_id];
          for (int r = 0; r < row.block.size; ++r)
            for (int c = 0; c < col.size; ++c) {
              const double a = v[cell.position + r * col.size + c];
              x[col.position + c] += a * a;
            }
        }
      }
      return;
    }
    ParallelFor(0, static_cast<int>(tbs.rows.size()), num_threads, [&](int, int cb) {
      const CompressedRow& trow = tbs.rows[cb];
      for (int c = 0; c < trow.block.size; ++c) x[trow.block.position + c] = 0.0;
      for (const Cell& cell : trow.cells) {
        const int nr = tbs.cols[cell.block_id].size;
        for (int r = 0; r < nr; ++r)
          for (int c = 0; c < trow.block.size; ++c) {
            const double a = v[cell.position + r * trow.block.size + c];
            x[trow.block.position + c] += a * a;
          }
      }
    });
  }

  // A <- A diag(scale)    (block_sparse_matrix.cc:403-450)
  void ScaleColumns(const double* scale, int num_threads) {
    double* v = values.data();
    ParallelFor(0, static_cast<int>(bs.rows.size()), num_threads, [&](int, int r) {
      const CompressedRow& row = bs.rows[r];
      for (const Cell& cell : row.cells) {
        const Block& col = bs.cols[cell.block_id];
        for (int i = 0; i < row.block.size; ++i)
          for (int c = 0; c < col.size; ++c) v[cell.position + i * col.size + c] *= scale[col.position + c];
      }
    });
  }
};

}  // namespace orc
