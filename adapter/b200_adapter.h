// Ceres-side adapters over libb200ba.so (include/b200ba.h).  Compiled INSIDE a Ceres tree (they include Ceres'
// internal headers, which need Eigen — not available in the development image, so these files are source only here;
// see INTEGRATION.md).  Three classes, all in ceres::internal:
//
//   B200Jacobian             : BlockSparseMatrix   the handle Solver::Solve / TrustRegionMinimizer see as "the
//                                                   Jacobian"; the values live in HBM, the virtuals the minimizer calls
//                                                   (sparse_matrix.h:67-116) forward to the C ABI
//   B200Evaluator            : Evaluator           evaluator.h:60-168
//   B200IterativeSchurSolver : BlockSparseMatrixSolver (TypedLinearSolver<BlockSparseMatrix>, linear_solver.h:366-387)
//
// Selected from the unmodified bundle_adjuster CLI with
//   --linear_solver=iterative_schur --sparse_linear_algebra_library=cuda_sparse --preconditioner=schur_jacobi
#ifndef CERES_INTERNAL_B200_ADAPTER_H_
#define CERES_INTERNAL_B200_ADAPTER_H_

#include <memory>
#include <string>

#include "b200ba.h"
#include "ceres/block_sparse_matrix.h"
#include "ceres/evaluator.h"
#include "ceres/execution_summary.h"
#include "ceres/linear_solver.h"
#include "ceres/program.h"

namespace ceres::internal {

// Shared owner of the device problem; evaluator, Jacobian and linear solver all point at it.
struct B200Context {
  b200_handle* handle = nullptr;
  // Host vector the last Evaluate() filled with residuals: when the minimizer hands the same pointer to the linear
  // solver (trust_region_minimizer.cc:399-402) the copy that is still in HBM is used instead of uploading it again.
  const double* last_residuals = nullptr;
  ~B200Context() { b200_destroy(handle); }
};

class B200Jacobian final : public BlockSparseMatrix {  // needs `final` dropped from block_sparse_matrix.h:60
 public:
  B200Jacobian(CompressedRowBlockStructure* bs, std::shared_ptr<B200Context> ctx)
      : BlockSparseMatrix(bs), ctx_(std::move(ctx)) {}
  b200_handle* handle() const { return ctx_->handle; }
  const B200Context& context() const { return *ctx_; }
  // -(J step)'(r + J step / 2) in one pass, r = residuals of the last Evaluate (optional minimizer hunk, ceres_b200.patch)
  bool ModelCostChange(const double* step, double* model_cost_change) const {
    return b200_model_cost_change(ctx_->handle, step, model_cost_change) == B200_OK;
  }

  // The four calls TrustRegionMinimizer / LevenbergMarquardtStrategy make on the Jacobian
  // (trust_region_minimizer.cc:269,277,431; levenberg_marquardt_strategy.cc:84).
  void SquaredColumnNorm(double* x) const final;
  void SquaredColumnNorm(double* x, ContextImpl*, int) const final { SquaredColumnNorm(x); }
  void ScaleColumns(const double* scale) final;
  void ScaleColumns(const double* scale, ContextImpl*, int) final { ScaleColumns(scale); }
  void RightMultiplyAndAccumulate(const double* x, double* y) const final;
  void RightMultiplyAndAccumulate(const double* x, double* y, ContextImpl*, int) const final {
    RightMultiplyAndAccumulate(x, y);
  }
  void LeftMultiplyAndAccumulate(const double* x, double* y) const final;
  void SetZero() final {}  // the evaluator overwrites every cell on the device
  // CPU consumers (problem dumps, CLUSTER_* preconditioners) pull the values explicitly:
  void SyncValuesToHost() { b200_jacobian_get_values(handle(), mutable_values()); }

 private:
  std::shared_ptr<B200Context> ctx_;
};

class B200Evaluator final : public Evaluator {
 public:
  // Returns nullptr + *error (as Evaluator::Create does, evaluator.cc:95-97) unless every residual block is
  // AutoDiffCostFunction<SnavelyReprojectionError, 2, 9, 3> on (camera[9], point[3]) with a null or Huber loss.
  static std::unique_ptr<Evaluator> Create(const Evaluator::Options& options, Program* program, std::string* error);

  std::unique_ptr<SparseMatrix> CreateJacobian() const final;
  bool Evaluate(const Evaluator::EvaluateOptions& evaluate_options, const double* state, double* cost,
                double* residuals, double* gradient, SparseMatrix* jacobian) final;
  bool Plus(const double* state, const double* delta, double* state_plus_delta) const final;
  int NumParameters() const final { return program_->NumParameters(); }
  int NumEffectiveParameters() const final { return program_->NumEffectiveParameters(); }
  int NumResiduals() const final { return program_->NumResiduals(); }
  std::map<std::string, CallStatistics> Statistics() const final { return execution_summary_.statistics(); }

 private:
  B200Evaluator(Program* program, std::shared_ptr<B200Context> ctx, int num_eliminate_blocks)
      : program_(program), ctx_(std::move(ctx)), num_eliminate_blocks_(num_eliminate_blocks) {}
  Program* program_;
  std::shared_ptr<B200Context> ctx_;
  int num_eliminate_blocks_;
  ExecutionSummary execution_summary_;
};

class B200IterativeSchurSolver final : public BlockSparseMatrixSolver {
 public:
  explicit B200IterativeSchurSolver(LinearSolver::Options options) : options_(std::move(options)) {}

 private:
  LinearSolver::Summary SolveImpl(BlockSparseMatrix* A, const double* b,
                                  const LinearSolver::PerSolveOptions& per_solve_options, double* x) final;
  LinearSolver::Options options_;
};

// DENSE_SCHUR / SPARSE_SCHUR on the device: explicit reduced camera system + Cholesky (b200_dense_schur_solve), the exact
// solve DenseSchurComplementSolver / SparseSchurComplementSolver perform (schur_complement_solver.cc:101-214, :224-408).
class B200DenseSchurSolver final : public BlockSparseMatrixSolver {
 public:
  explicit B200DenseSchurSolver(LinearSolver::Options options) : options_(std::move(options)) {}

 private:
  LinearSolver::Summary SolveImpl(BlockSparseMatrix* A, const double* b,
                                  const LinearSolver::PerSolveOptions& per_solve_options, double* x) final;
  LinearSolver::Options options_;
};

}  // namespace ceres::internal
#endif  // CERES_INTERNAL_B200_ADAPTER_H_
