// TEST INFRASTRUCTURE.  Drives adapter/b200_adapter.{h,cc} -- compiled against tests/mock_ceres, a restatement of the
// Ceres-internal interfaces it is written for -- the way Ceres' own Evaluator::Create / LinearSolver::Create callers
// and the trust-region minimizer would:
//   adapter_mock_driver <case> [problem.bin]
// cases that need no device (the adapter must refuse BEFORE touching the library, each with its own message):
//   bad_order | other_functor | manifold | cauchy | mixed_loss | mixed_huber | callback | predicate | huber <a>
// case `solve problem.bin` (format of tests/test_c_abi_smoke.py): Create -> CreateJacobian -> Evaluate ->
//   SquaredColumnNorm -> B200IterativeSchurSolver::Solve with the residual pointer the evaluator filled ->
//   RightMultiplyAndAccumulate -> ModelCostChange, one "key value" line each.  Exit 3 when b200_create finds no device.
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ceres/b200_adapter.h"
#include "ceres/autodiff_cost_function.h"
#include "ceres/residual_block.h"
#include "snavely_reprojection_error.h"

using namespace ceres;            // NOLINT
using namespace ceres::internal;  // NOLINT
using SnavelyCost = AutoDiffCostFunction<examples::SnavelyReprojectionError, 2, 9, 3>;

namespace {
struct Owned {  // the mock Program does not own its blocks
  std::vector<std::unique_ptr<ParameterBlock>> params;
  std::vector<std::unique_ptr<ResidualBlock>> residuals;
  std::vector<std::unique_ptr<CostFunction>> costs;
  std::vector<std::unique_ptr<LossFunction>> losses;
  std::vector<double> state;
  Program program;
};

// reduced program of a BAL problem: points are blocks 0..P-1 (the first elimination group), cameras P..P+C-1
// (reorder_program.cc:262-273); residual i observes (camera cam[i], point pt[i]) in the functor's parameter order
void Build(Owned* o, int C, int P, int64_t N, const int32_t* cam, const int32_t* pt, const double* obs, const double* state,
           bool cameras_first = false) {
  o->state.assign(state, state + 3 * P + 9 * C);
  for (int i = 0; i < P + C; ++i) {
    const bool is_point = cameras_first ? i >= C : i < P;
    const int k = cameras_first ? (is_point ? i - C : i) : (is_point ? i : i - P);
    double* at = o->state.data() + (is_point ? 3 * k : 3 * P + 9 * k);
    o->params.push_back(std::make_unique<ParameterBlock>(at, is_point ? 3 : 9, i));
    o->program.mutable_parameter_blocks()->push_back(o->params.back().get());
  }
  for (int64_t i = 0; i < N; ++i) {
    o->costs.push_back(std::make_unique<SnavelyCost>(new examples::SnavelyReprojectionError(obs[2 * i], obs[2 * i + 1])));
    ParameterBlock* camera = o->params[cameras_first ? cam[i] : P + cam[i]].get();
    ParameterBlock* point = o->params[cameras_first ? C + pt[i] : pt[i]].get();
    o->residuals.push_back(std::make_unique<ResidualBlock>(o->costs.back().get(), nullptr, std::vector<ParameterBlock*>{camera, point}));
    o->program.mutable_residual_blocks()->push_back(o->residuals.back().get());
  }
}

void BuildTiny(Owned* o, bool cameras_first = false) {
  const int32_t cam[6] = {0, 1, 0, 1, 0, 1}, pt[6] = {0, 0, 1, 1, 2, 2};
  const double obs[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12};
  std::vector<double> state(3 * 3 + 9 * 2, 0.5);
  Build(o, 2, 3, 6, cam, pt, obs, state.data(), cameras_first);
}

Evaluator::Options DeviceOptions(int P) {
  Evaluator::Options eo;
  eo.num_eliminate_blocks = P;
  eo.linear_solver_type = ITERATIVE_SCHUR;
  eo.sparse_linear_algebra_library_type = CUDA_SPARSE;
  return eo;
}

int Refused(Owned* o, const Evaluator::Options& eo) {
  std::string error;
  auto ev = B200Evaluator::Create(eo, &o->program, &error);
  if (ev != nullptr) {
    std::printf("created\n");
    return 0;
  }
  std::printf("refused: %s\n", error.c_str());
  return 4;
}

void SetLoss(Owned* o, size_t i, LossFunction* loss) {
  o->losses.emplace_back(loss);
  ResidualBlock* old = o->residuals[i].get();
  auto nb = std::make_unique<ResidualBlock>(old->cost_function(), loss, std::vector<ParameterBlock*>{old->parameter_blocks()[0], old->parameter_blocks()[1]});
  (*o->program.mutable_residual_blocks())[i] = nb.get();
  o->residuals[i] = std::move(nb);
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  const std::string which = argv[1];
  Owned o;
  if (which == "predicate") {  // the one predicate both factory hunks use
    std::printf("%d %d %d %d\n", B200Selected(ITERATIVE_SCHUR, CUDA_SPARSE), B200Selected(SPARSE_SCHUR, CUDA_SPARSE),
                B200Selected(ITERATIVE_SCHUR, SUITE_SPARSE), B200Selected(CGNR, CUDA_SPARSE));
    return 0;
  }
  if (which == "huber") {
    if (argc < 3) return 1;
    HuberLoss loss(std::atof(argv[2]));
    std::printf("%.17g\n", B200HuberScale(loss));
    return 0;
  }
  if (which == "bad_order") {
    BuildTiny(&o, /*cameras_first=*/true);
    return Refused(&o, DeviceOptions(3));
  }
  if (which == "other_functor") {
    BuildTiny(&o);
    o.costs.push_back(std::make_unique<AutoDiffCostFunction<examples::SomeOtherError, 2, 9, 3>>(new examples::SomeOtherError));
    ResidualBlock* old = o.residuals[4].get();
    auto nb = std::make_unique<ResidualBlock>(o.costs.back().get(), nullptr, std::vector<ParameterBlock*>{old->parameter_blocks()[0], old->parameter_blocks()[1]});
    (*o.program.mutable_residual_blocks())[4] = nb.get();
    o.residuals[4] = std::move(nb);
    return Refused(&o, DeviceOptions(3));
  }
  if (which == "manifold") {
    BuildTiny(&o);
    o.params[3]->set_manifold_for_test(reinterpret_cast<const Manifold*>(&o));  // any non-null manifold (quaternion cameras)
    return Refused(&o, DeviceOptions(3));
  }
  if (which == "cauchy") {
    BuildTiny(&o);
    for (size_t i = 0; i < o.residuals.size(); ++i) SetLoss(&o, i, new CauchyLoss(1.0));
    return Refused(&o, DeviceOptions(3));
  }
  if (which == "mixed_loss") {
    BuildTiny(&o);
    SetLoss(&o, 2, new HuberLoss(1.0));
    return Refused(&o, DeviceOptions(3));
  }
  if (which == "mixed_huber") {
    BuildTiny(&o);
    for (size_t i = 0; i < o.residuals.size(); ++i) SetLoss(&o, i, new HuberLoss(i == 5 ? 2.0 : 1.0));
    return Refused(&o, DeviceOptions(3));
  }
  if (which == "callback") {
    BuildTiny(&o);
    Evaluator::Options eo = DeviceOptions(3);
    eo.evaluation_callback = reinterpret_cast<EvaluationCallback*>(&o);
    return Refused(&o, eo);
  }
  if (which != "solve" || argc < 3) return 1;

  FILE* f = std::fopen(argv[2], "rb");
  if (f == nullptr) return 1;
  int32_t C, P;
  int64_t N;
  if (std::fread(&C, 4, 1, f) != 1 || std::fread(&P, 4, 1, f) != 1 || std::fread(&N, 8, 1, f) != 1) return 1;
  std::vector<int32_t> cam(N), pt(N);
  std::vector<double> obs(2 * N), state(3 * P + 9 * C);
  if (std::fread(cam.data(), 4, N, f) != static_cast<size_t>(N) || std::fread(pt.data(), 4, N, f) != static_cast<size_t>(N) ||
      std::fread(obs.data(), 8, 2 * N, f) != static_cast<size_t>(2 * N) || std::fread(state.data(), 8, state.size(), f) != state.size())
    return 1;
  std::fclose(f);
  const bool huber = argc > 3 && std::strcmp(argv[3], "huber") == 0;
  Build(&o, C, P, N, cam.data(), pt.data(), obs.data(), state.data());
  if (huber) for (size_t i = 0; i < o.residuals.size(); ++i) SetLoss(&o, i, new HuberLoss(1.0));

  const Evaluator::Options eo = DeviceOptions(P);
  if (!B200Selected(eo.linear_solver_type, eo.sparse_linear_algebra_library_type)) return 2;
  std::string error;
  std::unique_ptr<Evaluator> ev = B200Evaluator::Create(eo, &o.program, &error);
  if (ev == nullptr) {
    std::fprintf(stderr, "create failed: %s\n", error.c_str());
    return error.find("no CPU fallback") != std::string::npos ? 3 : 2;
  }
  const int np = ev->NumParameters(), nr = ev->NumResiduals();
  std::printf("sizes %d %d %d\n", np, ev->NumEffectiveParameters(), nr);
  std::unique_ptr<SparseMatrix> jacobian = ev->CreateJacobian();
  std::printf("jacobian %d %d %d\n", jacobian->num_rows(), jacobian->num_cols(), jacobian->num_nonzeros());

  double cost = 0.0;
  std::vector<double> residuals(nr), gradient(np);
  if (!ev->Evaluate(o.state.data(), &cost, residuals.data(), gradient.data(), jacobian.get())) return 2;
  double gmax = 0.0;
  for (double g : gradient) gmax = std::max(gmax, std::fabs(g));
  std::printf("cost %.17g\ngradient_max_norm %.17g\n", cost, gmax);
  if (huber) {  // EvaluateOptions::apply_loss_function = false (Problem::Evaluate, Solver::Summary's final cost)
    Evaluator::EvaluateOptions raw;
    raw.apply_loss_function = false;
    double raw_cost = 0.0;
    if (!ev->Evaluate(raw, o.state.data(), &raw_cost, nullptr, nullptr, nullptr)) return 2;
    std::printf("raw_cost %.17g\n", raw_cost);
    if (!ev->Evaluate(o.state.data(), &cost, residuals.data(), gradient.data(), jacobian.get())) return 2;
  }

  // LevenbergMarquardtStrategy::ComputeStep (levenberg_marquardt_strategy.cc:84-95) on the unscaled Jacobian, radius 1e4
  std::vector<double> D(np);
  jacobian->SquaredColumnNorm(D.data());
  double colsum = 0.0;
  for (double d : D) colsum += d;
  std::printf("column_norm_sum %.17g\n", colsum);
  for (double& d : D) d = std::sqrt(std::min(std::max(d, 1e-6), 1e32) / 1e4);

  LinearSolver::Options lo;
  lo.type = ITERATIVE_SCHUR;
  lo.preconditioner_type = SCHUR_JACOBI;
  lo.sparse_linear_algebra_library_type = CUDA_SPARSE;
  lo.min_num_iterations = 0;
  lo.max_num_iterations = 500;
  std::unique_ptr<LinearSolver> solver = std::make_unique<B200IterativeSchurSolver>(lo);
  LinearSolver::PerSolveOptions per;
  per.D = D.data();
  per.q_tolerance = 1e-2;
  per.r_tolerance = -1.0;
  std::vector<double> x(np, 0.0);
  const LinearSolver::Summary s = solver->Solve(jacobian.get(), residuals.data(), per, x.data());
  double xn = 0.0;
  for (double v : x) xn += v * v;
  std::printf("solve iterations %d termination %d step_norm %.17g\n", s.num_iterations, static_cast<int>(s.termination_type), std::sqrt(xn));
  if (solver->Statistics().count("LinearSolver::Solve") != 1) return 2;

  // J x through the overridden product, and the minimizer's model cost change for step = -x
  std::vector<double> y(nr, 0.0);
  jacobian->RightMultiplyAndAccumulate(x.data(), y.data());
  double yn = 0.0;
  for (double v : y) yn += v * v;
  std::printf("jx_norm %.17g\n", std::sqrt(yn));
  std::vector<double> step(np);
  for (int i = 0; i < np; ++i) step[i] = -x[i];
  double mcc = 0.0;
  if (!down_cast<B200Jacobian*>(jacobian.get())->ModelCostChange(step.data(), &mcc)) return 2;
  std::printf("model_cost_change %.17g\n", mcc);
  std::vector<double> plus(np);
  if (!ev->Plus(o.state.data(), step.data(), plus.data())) return 2;
  double pn = 0.0;
  for (int i = 0; i < np; ++i) pn += (plus[i] - o.state[i] - step[i]) * (plus[i] - o.state[i] - step[i]);
  std::printf("plus_error %.17g\n", std::sqrt(pn));
  std::printf("evaluator_calls %d\n", ev->Statistics().at("Evaluator::Total").calls);
  std::printf("done\n");
  return 0;
}
