// See b200_adapter.h.  Every C-ABI call is checked; failures map onto the reference's own error channels
// (Evaluate -> false, Create -> nullptr + *error, Solve -> Summary::termination_type), never exceptions.
#include "ceres/b200_adapter.h"

#include <cmath>
#include <vector>

#include "absl/log/check.h"
#include "absl/log/log.h"
#include "ceres/autodiff_cost_function.h"
#include "ceres/block_jacobian_writer.h"
#include "ceres/loss_function.h"
#include "ceres/parameter_block.h"
#include "ceres/residual_block.h"
#include "snavely_reprojection_error.h"  // examples/: the cost functor the device kernel implements

namespace ceres::internal {
namespace {
using SnavelyCost = AutoDiffCostFunction<examples::SnavelyReprojectionError, 2, 9, 3>;

bool Check(int rc, const char* what) {
  if (rc == B200_OK) return true;
  LOG(ERROR) << what << " failed: " << b200_last_error();
  return false;
}
}  // namespace

// ---------------------------------------------------------------- B200Jacobian
void B200Jacobian::SquaredColumnNorm(double* x) const { CHECK(Check(b200_jacobian_squared_column_norm(handle(), x), "SquaredColumnNorm")); }
void B200Jacobian::ScaleColumns(const double* scale) { CHECK(Check(b200_jacobian_scale_columns(handle(), scale), "ScaleColumns")); }
void B200Jacobian::RightMultiplyAndAccumulate(const double* x, double* y) const {
  CHECK(Check(b200_jacobian_right_multiply(handle(), x, y), "RightMultiplyAndAccumulate"));
}
void B200Jacobian::LeftMultiplyAndAccumulate(const double* x, double* y) const {
  CHECK(Check(b200_jacobian_left_multiply(handle(), x, y), "LeftMultiplyAndAccumulate"));
}

// ---------------------------------------------------------------- B200Evaluator
std::unique_ptr<Evaluator> B200Evaluator::Create(const Evaluator::Options& options, Program* program, std::string* error) {
  if (options.evaluation_callback != nullptr) {
    *error = "B200Evaluator: evaluation callbacks (evaluator.h:72) are not supported on the device path";
    return nullptr;
  }
  const int P = options.num_eliminate_blocks;  // e blocks come first in the reduced program (reorder_program.cc:262-273)
  const int C = program->NumParameterBlocks() - P;
  const auto& residual_blocks = program->residual_blocks();
  const int64_t N = static_cast<int64_t>(residual_blocks.size());
  std::vector<int32_t> cam_idx(N), pt_idx(N);
  std::vector<double> obs(2 * N);
  int loss_type = B200_LOSS_TRIVIAL;
  double loss_a = 1.0;
  for (int64_t i = 0; i < N; ++i) {
    const ResidualBlock* rb = residual_blocks[i];
    const auto* cost = dynamic_cast<const SnavelyCost*>(rb->cost_function());
    if (cost == nullptr || rb->NumParameterBlocks() != 2 || rb->parameter_blocks()[0]->Size() != 9 ||
        rb->parameter_blocks()[1]->Size() != 3 || rb->parameter_blocks()[0]->manifold() != nullptr ||
        rb->parameter_blocks()[1]->manifold() != nullptr) {
      *error = "B200Evaluator: residual block " + std::to_string(i) + " is not SnavelyReprojectionError<2,9,3> on Euclidean (camera, point) blocks";
      return nullptr;  // no silent CPU fallback on the hot path
    }
    const int cam_block = rb->parameter_blocks()[0]->index();  // index in the reduced program
    const int pt_block = rb->parameter_blocks()[1]->index();
    if (pt_block >= P || cam_block < P) {
      *error = "B200Evaluator: points must form the first elimination group (bundle_adjuster: --ordering_type=user)";
      return nullptr;
    }
    pt_idx[i] = pt_block;
    cam_idx[i] = cam_block - P;
    obs[2 * i] = cost->functor().observed_x;
    obs[2 * i + 1] = cost->functor().observed_y;
    // every residual block must carry the SAME loss: null, or HuberLoss(a) with one a
    int this_type = B200_LOSS_TRIVIAL;
    double this_a = 1.0;
    if (const LossFunction* loss = rb->loss_function()) {
      if (dynamic_cast<const HuberLoss*>(loss) == nullptr) {
        *error = "B200Evaluator: only the trivial and Huber losses are implemented";
        return nullptr;
      }
      this_type = B200_LOSS_HUBER;
      this_a = B200HuberScale(*loss);
      if (!std::isfinite(this_a)) {
        *error = "B200Evaluator: HuberLoss scale of residual block " + std::to_string(i) + " could not be recovered";
        return nullptr;
      }
    }
    if (i == 0) {
      loss_type = this_type;
      loss_a = this_a;
    } else if (this_type != loss_type || (this_type == B200_LOSS_HUBER && std::fabs(this_a - loss_a) > 1e-12 * loss_a)) {
      *error = "B200Evaluator: residual blocks with different loss functions (block " + std::to_string(i) + ")";
      return nullptr;
    }
  }
  b200_ba_desc desc{};
  desc.num_cameras = C;
  desc.num_points = P;
  desc.num_observations = N;
  desc.cam_idx = cam_idx.data();
  desc.pt_idx = pt_idx.data();
  desc.obs = obs.data();
  desc.loss_type = loss_type;
  desc.loss_a = loss_a;
  desc.device = 0;
  desc.world_size = 1;
  auto ctx = std::make_shared<B200Context>();
  if (b200_create(&desc, &ctx->handle) != B200_OK) {
    *error = std::string("B200Evaluator: ") + b200_last_error();
    return nullptr;
  }
  return std::unique_ptr<Evaluator>(new B200Evaluator(program, std::move(ctx), P));
}

std::unique_ptr<SparseMatrix> B200Evaluator::CreateJacobian() const {
  // Same block structure Ceres would build (block_jacobian_writer.cc:198-263) so DetectStructure and friends work.
  Evaluator::Options opts;
  opts.num_eliminate_blocks = num_eliminate_blocks_;
  BlockJacobianWriter writer(opts, program_);
  std::unique_ptr<SparseMatrix> plain = writer.CreateJacobian();
  auto* bsm = down_cast<BlockSparseMatrix*>(plain.get());
  auto* bs = new CompressedRowBlockStructure(*bsm->block_structure());
  return std::make_unique<B200Jacobian>(bs, ctx_);
}

bool B200Evaluator::Evaluate(const Evaluator::EvaluateOptions& evaluate_options, const double* state, double* cost,
                             double* residuals, double* gradient, SparseMatrix* jacobian) {
  // apply_loss_function = false (evaluator.h:101: Problem::Evaluate and the final cost of Solver::Summary ask for it)
  // switches the robust correction off on the device; new_evaluation_point only matters to evaluation callbacks
  if (evaluate_options.apply_loss_function != ctx_->apply_loss_function) {
    if (!Check(b200_set_apply_loss_function(ctx_->handle, evaluate_options.apply_loss_function ? 1 : 0), "b200_set_apply_loss_function"))
      return false;
    ctx_->apply_loss_function = evaluate_options.apply_loss_function;
  }
  ScopedExecutionTimer total("Evaluator::Total", &execution_summary_);
  ScopedExecutionTimer kind(gradient == nullptr && jacobian == nullptr ? "Evaluator::Residual" : "Evaluator::Jacobian",
                            &execution_summary_);  // the keys Solver::Summary reads (solver.cc:615-628)
  const int rc = b200_evaluate(ctx_->handle, state, cost, residuals, gradient, jacobian != nullptr);
  if (residuals != nullptr) ctx_->last_residuals = rc == B200_OK ? residuals : nullptr;
  if (rc == B200_ERR_EVALUATION_FAILED) return false;
  return Check(rc, "b200_evaluate");
}

bool B200Evaluator::Plus(const double* state, const double* delta, double* state_plus_delta) const {
  return b200_plus(ctx_->handle, state, delta, state_plus_delta) == B200_OK;
}

// ---------------------------------------------------------------- B200IterativeSchurSolver
LinearSolver::Summary B200IterativeSchurSolver::SolveImpl(BlockSparseMatrix* A, const double* b,
                                                          const LinearSolver::PerSolveOptions& per_solve_options,
                                                          double* x) {
  LinearSolver::Summary summary;
  auto* jac = dynamic_cast<B200Jacobian*>(A);
  if (jac == nullptr) {
    summary.termination_type = LinearSolverTerminationType::FATAL_ERROR;
    summary.message = "B200IterativeSchurSolver needs the Jacobian created by B200Evaluator.";
    return summary;
  }
  if (options_.use_explicit_schur_complement) {
    // Evaluator::Options carries no such flag, so the evaluator factory cannot opt out for it: refuse loudly here rather
    // than hand a Jacobian whose values live in HBM to SparseSchurComplementSolver
    summary.termination_type = LinearSolverTerminationType::FATAL_ERROR;
    summary.message = "use_explicit_schur_complement is not available on the B200 path (the Schur complement stays implicit).";
    return summary;
  }
  b200_solver_options o;
  b200_solver_options_default(&o);
  switch (options_.preconditioner_type) {
    case IDENTITY: o.preconditioner_type = B200_PRECOND_IDENTITY; break;
    case JACOBI: o.preconditioner_type = B200_PRECOND_JACOBI; break;
    case SCHUR_JACOBI: o.preconditioner_type = B200_PRECOND_SCHUR_JACOBI; break;
    case SCHUR_POWER_SERIES_EXPANSION: o.preconditioner_type = B200_PRECOND_SCHUR_POWER_SERIES_EXPANSION; break;
    default:
      summary.termination_type = LinearSolverTerminationType::FATAL_ERROR;
      summary.message = "Preconditioner not implemented on the B200 path.";
      return summary;
  }
  o.min_num_iterations = options_.min_num_iterations;
  o.max_num_iterations = options_.max_num_iterations;
  o.residual_reset_period = options_.residual_reset_period;
  o.max_num_spse_iterations = options_.max_num_spse_iterations;
  o.use_spse_initialization = options_.use_spse_initialization ? 1 : 0;
  o.spse_tolerance = options_.spse_tolerance;
  o.q_tolerance = per_solve_options.q_tolerance;
  o.r_tolerance = per_solve_options.r_tolerance;
  b200_solver_summary s{};
  // b is the residual vector of the last Evaluate in the LM loop: its device copy is still valid (b200ba.h)
  const double* b_arg = (b == jac->context().last_residuals) ? nullptr : b;
  if (b200_schur_solve(jac->handle(), b_arg, per_solve_options.D, &o, x, &s) != B200_OK) {
    summary.termination_type = LinearSolverTerminationType::FATAL_ERROR;
    summary.message = b200_last_error();
    return summary;
  }
  summary.num_iterations = s.num_iterations;
  summary.residual_norm = s.residual_norm;
  static_assert(static_cast<int>(LinearSolverTerminationType::SUCCESS) == B200_LS_SUCCESS &&
                    static_cast<int>(LinearSolverTerminationType::NO_CONVERGENCE) == B200_LS_NO_CONVERGENCE &&
                    static_cast<int>(LinearSolverTerminationType::FAILURE) == B200_LS_FAILURE &&
                    static_cast<int>(LinearSolverTerminationType::FATAL_ERROR) == B200_LS_FATAL_ERROR,
                "b200ba.h numbers the termination types like linear_solver.h:57-74");
  summary.termination_type = static_cast<LinearSolverTerminationType>(s.termination_type);
  return summary;
}

}  // namespace ceres::internal
