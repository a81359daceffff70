// TEST INFRASTRUCTURE -- a minimal restatement of the Ceres-internal interfaces adapter/b200_adapter.{h,cc} is written
// against, so that the adapter can be COMPILED and its host-side logic exercised in an image without Eigen (every real
// Ceres header includes it).  Only declarations the adapter touches exist here; signatures restate the reference's
// (file:line below; BlockSparseMatrix as adapter/ceres_b200.patch leaves it) and tests/test_adapter_mock.py checks each
// of them against the reference tree when it is present.  Nothing under ceres_solver_b200/ uses this directory.
#ifndef B200_TESTS_MOCK_CERES_ALL_H_
#define B200_TESTS_MOCK_CERES_ALL_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace ceres {

// include/ceres/types.h:57-91, 93-141, 166-185
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum PreconditionerType { IDENTITY, JACOBI, SCHUR_JACOBI, SCHUR_POWER_SERIES_EXPANSION, CLUSTER_JACOBI, CLUSTER_TRIDIAGONAL, SUBSET };
enum SparseLinearAlgebraLibraryType { SUITE_SPARSE, EIGEN_SPARSE, ACCELERATE_SPARSE, CUDA_SPARSE, NO_SPARSE };

class Manifold;
class EvaluationCallback;

// include/ceres/cost_function.h, loss_function.h:114,174-183 (loss_function.cc:52-66 for the Huber arithmetic)
class CostFunction {
 public:
  virtual ~CostFunction() = default;
  const std::vector<int32_t>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t> sizes_;
  int num_residuals_ = 0;
};

class LossFunction {
 public:
  virtual ~LossFunction() = default;
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

class HuberLoss final : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double* rho) const override {
    if (s > b_) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a_ * r - b_;
      rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s;
      rho[1] = 1.0;
      rho[2] = 0.0;
    }
  }

 private:
  const double a_;
  const double b_;
};

class CauchyLoss final : public LossFunction {   // a loss the device path does NOT implement (for the rejection test)
 public:
  explicit CauchyLoss(double a) : b_(a * a), c_(1.0 / b_) {}
  void Evaluate(double s, double* rho) const override {
    const double sum = 1.0 + s * c_, inv = 1.0 / sum;
    rho[0] = b_ * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c_ * (inv * inv);
  }

 private:
  const double b_, c_;
};

// include/ceres/autodiff_cost_function.h:151-246: only the functor() accessor and the static sizes matter to the adapter
template <typename CostFunctor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction final : public CostFunction {
 public:
  explicit AutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {
    num_residuals_ = kNumResiduals;
    sizes_ = {Ns...};
  }
  const CostFunctor& functor() const { return *functor_; }

 private:
  std::unique_ptr<CostFunctor> functor_;
};

namespace examples {
// examples/snavely_reprojection_error.h:53-104
struct SnavelyReprojectionError {
  SnavelyReprojectionError(double observed_x, double observed_y) : observed_x(observed_x), observed_y(observed_y) {}
  double observed_x;
  double observed_y;
};
struct SomeOtherError {   // a functor the device path does not implement
  double w = 0.0;
};
}  // namespace examples

namespace internal {

class ContextImpl;

// internal/ceres/casts.h:89-104
template <typename To, typename From>
inline To down_cast(From* f) {
  return static_cast<To>(f);
}

// internal/ceres/execution_summary.h:46-90
struct CallStatistics {
  double time{0.};
  int calls{0};
};
class ExecutionSummary {
 public:
  void IncrementTimeBy(const std::string& name, const double value) {
    CallStatistics& call_stats = statistics_[name];
    call_stats.time += value;
    ++call_stats.calls;
  }
  const std::map<std::string, CallStatistics>& statistics() const { return statistics_; }

 private:
  std::map<std::string, CallStatistics> statistics_;
};
class ScopedExecutionTimer {
 public:
  ScopedExecutionTimer(std::string name, ExecutionSummary* summary) : name_(std::move(name)), summary_(summary) {}
  ~ScopedExecutionTimer() { summary_->IncrementTimeBy(name_, 0.0); }

 private:
  std::string name_;
  ExecutionSummary* summary_;
};

// internal/ceres/block_structure.h:50-97
struct Block {
  Block() = default;
  Block(int size_, int position_) noexcept : size(size_), position(position_) {}
  int32_t size{-1};
  int position{-1};
};
struct Cell {
  Cell() = default;
  Cell(int block_id_, int position_) noexcept : block_id(block_id_), position(position_) {}
  int block_id{-1};
  int position{-1};
};
struct CompressedList {
  Block block;
  std::vector<Cell> cells;
  int nnz{-1};
  int cumulative_nnz{-1};
};
using CompressedRow = CompressedList;
struct CompressedRowBlockStructure {
  std::vector<Block> cols;
  std::vector<CompressedRow> rows;
};

// internal/ceres/linear_operator.h:46-83 (the Eigen-vector overloads are left out)
class LinearOperator {
 public:
  virtual ~LinearOperator() = default;
  virtual void RightMultiplyAndAccumulate(const double* x, double* y) const = 0;
  virtual void RightMultiplyAndAccumulate(const double* x, double* y, ContextImpl* context, int num_threads) const {
    (void)context; (void)num_threads;
    RightMultiplyAndAccumulate(x, y);
  }
  virtual void LeftMultiplyAndAccumulate(const double* x, double* y) const = 0;
  virtual void LeftMultiplyAndAccumulate(const double* x, double* y, ContextImpl* context, int num_threads) const {
    (void)context; (void)num_threads;
    LeftMultiplyAndAccumulate(x, y);
  }
  virtual int num_rows() const = 0;
  virtual int num_cols() const = 0;
};

// internal/ceres/sparse_matrix.h:65-117 (ToDenseMatrix / ToTextFile left out: Eigen, FILE dumps)
class SparseMatrix : public LinearOperator {
 public:
  ~SparseMatrix() override = default;
  using LinearOperator::RightMultiplyAndAccumulate;
  void RightMultiplyAndAccumulate(const double* x, double* y) const override = 0;
  void LeftMultiplyAndAccumulate(const double* x, double* y) const override = 0;
  virtual void SquaredColumnNorm(double* x) const = 0;
  virtual void SquaredColumnNorm(double* x, ContextImpl* context, int num_threads) const {
    (void)context; (void)num_threads;
    SquaredColumnNorm(x);
  }
  virtual void ScaleColumns(const double* scale) = 0;
  virtual void ScaleColumns(const double* scale, ContextImpl* context, int num_threads) {
    (void)context; (void)num_threads;
    ScaleColumns(scale);
  }
  virtual void SetZero() = 0;
  virtual void SetZero(ContextImpl* /*context*/, int /*num_threads*/) { SetZero(); }
  virtual double* mutable_values() = 0;
  virtual const double* values() const = 0;
  int num_rows() const override = 0;
  int num_cols() const override = 0;
  virtual int num_nonzeros() const = 0;
};

// internal/ceres/block_sparse_matrix.h:60-140 AS PATCHED by adapter/ceres_b200.patch: the class and the ten virtuals the
// device Jacobian overrides are no longer `final`; the accessors stay final, as in the reference.  A plain CPU
// implementation of the products is included so that the mock is a working matrix.
class BlockSparseMatrix : public SparseMatrix {
 public:
  explicit BlockSparseMatrix(CompressedRowBlockStructure* block_structure, bool use_page_locked_memory = false)
      : block_structure_(block_structure) {
    (void)use_page_locked_memory;
    num_rows_ = num_cols_ = num_nonzeros_ = 0;
    for (const Block& c : block_structure_->cols) num_cols_ += c.size;
    for (const CompressedRow& r : block_structure_->rows) {
      num_rows_ += r.block.size;
      for (const Cell& c : r.cells) num_nonzeros_ += r.block.size * block_structure_->cols[c.block_id].size;
    }
    values_.assign(num_nonzeros_, 0.0);
  }
  BlockSparseMatrix(const BlockSparseMatrix&) = delete;
  void operator=(const BlockSparseMatrix&) = delete;

  void SetZero() override { values_.assign(values_.size(), 0.0); }
  void SetZero(ContextImpl* context, int num_threads) override { (void)context; (void)num_threads; SetZero(); }
  void RightMultiplyAndAccumulate(const double* x, double* y) const override { Multiply(x, y, false); }
  void RightMultiplyAndAccumulate(const double* x, double* y, ContextImpl* context, int num_threads) const override {
    (void)context; (void)num_threads;
    Multiply(x, y, false);
  }
  void LeftMultiplyAndAccumulate(const double* x, double* y) const override { Multiply(x, y, true); }
  void LeftMultiplyAndAccumulate(const double* x, double* y, ContextImpl* context, int num_threads) const override {
    (void)context; (void)num_threads;
    Multiply(x, y, true);
  }
  void SquaredColumnNorm(double* x) const override {
    for (int i = 0; i < num_cols_; ++i) x[i] = 0.0;
    ForEach([&](int, int col, double v) { x[col] += v * v; });
  }
  void SquaredColumnNorm(double* x, ContextImpl* context, int num_threads) const override {
    (void)context; (void)num_threads;
    SquaredColumnNorm(x);
  }
  void ScaleColumns(const double* scale) override {
    size_t k = 0;
    ForEach([&](int, int col, double) { values_[k++] *= scale[col]; });
  }
  void ScaleColumns(const double* scale, ContextImpl* context, int num_threads) override {
    (void)context; (void)num_threads;
    ScaleColumns(scale);
  }

  int num_rows() const final { return num_rows_; }
  int num_cols() const final { return num_cols_; }
  int num_nonzeros() const final { return num_nonzeros_; }
  const double* values() const final { return values_.data(); }
  double* mutable_values() final { return values_.data(); }
  const CompressedRowBlockStructure* block_structure() const { return block_structure_.get(); }

 private:
  // visits the entries in storage order when every cell's values follow the previous cell's (true for the layouts built
  // by the mock writer only through `position`, so address them explicitly)
  template <typename F>
  void ForEach(F f) const {
    for (const CompressedRow& r : block_structure_->rows)
      for (const Cell& c : r.cells) {
        const Block& col = block_structure_->cols[c.block_id];
        for (int i = 0; i < r.block.size; ++i)
          for (int j = 0; j < col.size; ++j) f(r.block.position + i, col.position + j, values_[c.position + i * col.size + j]);
      }
  }
  void Multiply(const double* x, double* y, bool transpose) const {
    ForEach([&](int row, int col, double v) {
      if (transpose) y[col] += v * x[row];
      else y[row] += v * x[col];
    });
  }
  int num_rows_, num_cols_, num_nonzeros_;
  mutable std::vector<double> values_;
  std::unique_ptr<CompressedRowBlockStructure> block_structure_;
};

// internal/ceres/parameter_block.h:89,115,137
class ParameterBlock {
 public:
  ParameterBlock(double* user_state, int size, int index) : state_(user_state), size_(size), index_(index) {}
  int Size() const { return size_; }
  const Manifold* manifold() const { return manifold_; }
  int index() const { return index_; }
  void set_index(int index) { index_ = index; }
  void set_manifold_for_test(const Manifold* m) { manifold_ = m; }
  const double* state() const { return state_; }

 private:
  double* state_;
  int size_;
  int index_;
  const Manifold* manifold_ = nullptr;
};

// internal/ceres/residual_block.h:109-121
class ResidualBlock {
 public:
  ResidualBlock(const CostFunction* cost_function, const LossFunction* loss_function, const std::vector<ParameterBlock*>& blocks)
      : cost_function_(cost_function), loss_function_(loss_function), parameter_blocks_(new ParameterBlock*[blocks.size()]) {
    for (size_t i = 0; i < blocks.size(); ++i) parameter_blocks_[i] = blocks[i];
  }
  const CostFunction* cost_function() const { return cost_function_; }
  const LossFunction* loss_function() const { return loss_function_; }
  ParameterBlock* const* parameter_blocks() const { return parameter_blocks_.get(); }
  int NumParameterBlocks() const { return cost_function_->parameter_block_sizes().size(); }
  int NumResiduals() const { return cost_function_->num_residuals(); }

 private:
  const CostFunction* cost_function_;
  const LossFunction* loss_function_;
  std::unique_ptr<ParameterBlock*[]> parameter_blocks_;
};

// internal/ceres/program.h:65,158-162
class Program {
 public:
  const std::vector<ParameterBlock*>& parameter_blocks() const { return parameter_blocks_; }
  const std::vector<ResidualBlock*>& residual_blocks() const { return residual_blocks_; }
  std::vector<ParameterBlock*>* mutable_parameter_blocks() { return &parameter_blocks_; }
  std::vector<ResidualBlock*>* mutable_residual_blocks() { return &residual_blocks_; }
  int NumParameterBlocks() const { return static_cast<int>(parameter_blocks_.size()); }
  int NumParameters() const {
    int n = 0;
    for (const ParameterBlock* p : parameter_blocks_) n += p->Size();
    return n;
  }
  int NumEffectiveParameters() const { return NumParameters(); }
  int NumResiduals() const {
    int n = 0;
    for (const ResidualBlock* r : residual_blocks_) n += r->NumResiduals();
    return n;
  }

 private:
  std::vector<ParameterBlock*> parameter_blocks_;
  std::vector<ResidualBlock*> residual_blocks_;
};

// internal/ceres/evaluator.h:60-168
class Evaluator {
 public:
  virtual ~Evaluator() = default;
  struct Options {
    int num_threads = 1;
    int num_eliminate_blocks = -1;
    LinearSolverType linear_solver_type = DENSE_QR;
    SparseLinearAlgebraLibraryType sparse_linear_algebra_library_type = NO_SPARSE;
    bool dynamic_sparsity = false;
    ContextImpl* context = nullptr;
    EvaluationCallback* evaluation_callback = nullptr;
  };
  virtual std::unique_ptr<SparseMatrix> CreateJacobian() const = 0;
  struct EvaluateOptions {
    bool apply_loss_function = true;
    bool new_evaluation_point = true;
  };
  virtual bool Evaluate(const EvaluateOptions& evaluate_options, const double* state, double* cost, double* residuals,
                        double* gradient, SparseMatrix* jacobian) = 0;
  bool Evaluate(const double* state, double* cost, double* residuals, double* gradient, SparseMatrix* jacobian) {
    return Evaluate(EvaluateOptions(), state, cost, residuals, gradient, jacobian);
  }
  virtual bool Plus(const double* state, const double* delta, double* state_plus_delta) const = 0;
  virtual int NumParameters() const = 0;
  virtual int NumEffectiveParameters() const = 0;
  virtual int NumResiduals() const = 0;
  virtual std::map<std::string, CallStatistics> Statistics() const { return {}; }
};

// internal/ceres/block_jacobian_writer.h:67,76 + the layout of block_jacobian_writer.cc:68-167 for two-block residuals
// with the e block first in the elimination order: all E cells, then all F cells
class BlockJacobianWriter {
 public:
  BlockJacobianWriter(const Evaluator::Options& options, Program* program) : options_(options), program_(program) {}
  std::unique_ptr<SparseMatrix> CreateJacobian() const {
    auto* bs = new CompressedRowBlockStructure;
    int pos = 0;
    for (const ParameterBlock* p : program_->parameter_blocks()) {
      bs->cols.emplace_back(p->Size(), pos);
      pos += p->Size();
    }
    int e_total = 0;
    for (const ResidualBlock* r : program_->residual_blocks())
      for (int k = 0; k < r->NumParameterBlocks(); ++k)
        if (r->parameter_blocks()[k]->index() < options_.num_eliminate_blocks) e_total += r->NumResiduals() * r->parameter_blocks()[k]->Size();
    int row_pos = 0, e_pos = 0, f_pos = e_total;
    for (const ResidualBlock* r : program_->residual_blocks()) {
      CompressedRow row;
      row.block = Block(r->NumResiduals(), row_pos);
      row_pos += r->NumResiduals();
      std::vector<const ParameterBlock*> blocks(r->parameter_blocks(), r->parameter_blocks() + r->NumParameterBlocks());
      std::sort(blocks.begin(), blocks.end(), [](const ParameterBlock* a, const ParameterBlock* b) { return a->index() < b->index(); });
      for (const ParameterBlock* p : blocks) {
        const int size = r->NumResiduals() * p->Size();
        int& at = p->index() < options_.num_eliminate_blocks ? e_pos : f_pos;
        row.cells.emplace_back(p->index(), at);
        at += size;
      }
      bs->rows.push_back(row);
    }
    return std::make_unique<BlockSparseMatrix>(bs);
  }

 private:
  Evaluator::Options options_;
  Program* program_;
};

class LinearOperator;

// internal/ceres/linear_solver.h:57-74
enum class LinearSolverTerminationType { SUCCESS, NO_CONVERGENCE, FAILURE, FATAL_ERROR };

// internal/ceres/linear_solver.h:100-315 (fields the adapter does not read are kept so that the struct reads the same)
class LinearSolver {
 public:
  struct Options {
    LinearSolverType type = SPARSE_NORMAL_CHOLESKY;
    PreconditionerType preconditioner_type = JACOBI;
    SparseLinearAlgebraLibraryType sparse_linear_algebra_library_type = SUITE_SPARSE;
    bool dynamic_sparsity = false;
    bool use_explicit_schur_complement = false;
    int min_num_iterations = 1;
    int max_num_iterations = 1;
    int max_num_spse_iterations = 5;
    bool use_spse_initialization = false;
    double spse_tolerance = 0.1;
    int num_threads = 1;
    std::vector<int> elimination_groups;
    int residual_reset_period = 10;
    ContextImpl* context = nullptr;
  };
  struct PerSolveOptions {
    double* D = nullptr;
    LinearOperator* preconditioner = nullptr;
    double r_tolerance = 0.0;
    double q_tolerance = 0.0;
  };
  struct Summary {
    double residual_norm = -1.0;
    int num_iterations = -1;
    LinearSolverTerminationType termination_type = LinearSolverTerminationType::FAILURE;
    std::string message;
  };
  virtual ~LinearSolver() = default;
  virtual Summary Solve(LinearOperator* A, const double* b, const PerSolveOptions& per_solve_options, double* x) = 0;
  virtual std::map<std::string, CallStatistics> Statistics() const { return {}; }
};

// internal/ceres/linear_solver.h:341-387
template <typename MatrixType>
class TypedLinearSolver : public LinearSolver {
 public:
  LinearSolver::Summary Solve(LinearOperator* A, const double* b, const LinearSolver::PerSolveOptions& per_solve_options,
                              double* x) override {
    ScopedExecutionTimer total_time("LinearSolver::Solve", &execution_summary_);
    return SolveImpl(down_cast<MatrixType*>(A), b, per_solve_options, x);
  }
  std::map<std::string, CallStatistics> Statistics() const override { return execution_summary_.statistics(); }

 private:
  virtual LinearSolver::Summary SolveImpl(MatrixType* A, const double* b, const LinearSolver::PerSolveOptions& per_solve_options,
                                          double* x) = 0;
  ExecutionSummary execution_summary_;
};
using BlockSparseMatrixSolver = TypedLinearSolver<BlockSparseMatrix>;

}  // namespace internal
}  // namespace ceres

// absl/log/log.h, absl/log/check.h: LOG(ERROR) << ... and CHECK(cond)
namespace b200_mock_log {
struct Line {
  std::ostringstream s;
  ~Line() { std::cerr << s.str() << std::endl; }
};
}  // namespace b200_mock_log
#define LOG(severity) ::b200_mock_log::Line().s
#define CHECK(cond)                                              \
  do {                                                           \
    if (!(cond)) {                                               \
      std::fprintf(stderr, "CHECK failed: %s\n", #cond);         \
      std::abort();                                              \
    }                                                            \
  } while (0)

#endif  // B200_TESTS_MOCK_CERES_ALL_H_
