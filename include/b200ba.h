/* b200ba.h — C ABI of libb200ba.so: the B200-native (sm_100a) implementation of Ceres Solver's
 * Levenberg–Marquardt inner-loop hot path for bundle-adjustment-shaped problems
 * (row block 2, eliminated "e" blocks of size 3 = points, "f" blocks of size 9 = cameras).
 *
 * Ceres has no plugin registry or C ABI for evaluators/linear solvers; this header is what two thin
 * adapter classes (adapter/b200_evaluator.h, adapter/b200_iterative_schur_solver.h) bind, one entry
 * point per virtual of the two internal interfaces they subclass.  Each declaration cites the
 * reference interface it replaces (paths relative to the ceres-solver tree).
 *
 * Conventions: plain C types only; every array argument is a caller-owned HOST buffer unless the name
 * ends in _dev; functions return B200_OK (0) or a negative error code and never throw; the message of
 * the last error is available from b200_last_error().  A handle is NOT thread-safe (like the reference
 * objects: internal/ceres/program_evaluator.h:78-79).  All arithmetic is FP64.
 *
 * Vector layout (the reduced program's own order, internal/ceres/reorder_program.cc:262-273 and
 * block_jacobian_writer.cc:211-218):   x = [ e blocks: 3 doubles per point | f blocks: 9 per camera ].
 * Jacobian value layout (block_jacobian_writer.cc:68-167): all E cells [N][2][3] row-major first, then
 * all F cells [N][2][9] — identical to BlockSparseMatrix::values() for this structure.
 */
#ifndef B200BA_H_
#define B200BA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_handle b200_handle;

enum {
  B200_OK = 0,
  B200_ERR_INVALID_ARGUMENT = -1,
  B200_ERR_CUDA = -2,
  B200_ERR_EVALUATION_FAILED = -3, /* non-finite residual/Jacobian: Evaluator::Evaluate returns false */
  B200_ERR_NO_DEVICE = -4,
  B200_ERR_NCCL = -5,
  B200_ERR_UNSUPPORTED = -6
};

/* LinearSolverTerminationType, internal/ceres/linear_solver.h:58-74 (same numeric order). */
enum { B200_LS_SUCCESS = 0, B200_LS_NO_CONVERGENCE = 1, B200_LS_FAILURE = 2, B200_LS_FATAL_ERROR = 3 };
/* PreconditionerType subset valid for ITERATIVE_SCHUR here (include/ceres/types.h:93-119, same numeric order). */
enum {
  B200_PRECOND_IDENTITY = 0,
  B200_PRECOND_JACOBI = 1,
  B200_PRECOND_SCHUR_JACOBI = 2,
  B200_PRECOND_SCHUR_POWER_SERIES_EXPANSION = 3 /* power_series_expansion_preconditioner.cc:57-82 */
};
enum { B200_LOSS_TRIVIAL = 0, B200_LOSS_HUBER = 1 };
/* LinearSolverType subset (include/ceres/types.h): the implicit iterative solver and the exact solve on the explicit
 * reduced camera system (DENSE_SCHUR; stands in for SPARSE_SCHUR too: both are exact solves of the same system). */
enum { B200_ITERATIVE_SCHUR = 0, B200_DENSE_SCHUR = 1 };

/* Problem structure = what the adapters read off the reduced ceres::internal::Program
 * (residual_block->parameter_blocks()[j]->index(), SnavelyReprojectionError::observed_x/y,
 * residual_block->loss_function()).  Rows must be grouped by e block (pt_idx non-decreasing): the same
 * precondition SchurEliminator has (internal/ceres/schur_eliminator.h:85-99), established by
 * LexicographicallyOrderResidualBlocks (reorder_program.cc:278-359). */
typedef struct b200_ba_desc {
  int32_t num_cameras;       /* C: number of f blocks                                   */
  int32_t num_points;        /* P: number of e blocks (of this shard)                   */
  int64_t num_observations;  /* N: number of row blocks (of this shard)                 */
  const int32_t* cam_idx;    /* [N] f block id of row i  (cells[1].block_id - P)        */
  const int32_t* pt_idx;     /* [N] e block id of row i  (cells[0].block_id), sorted    */
  const double* obs;         /* [2N] observed_x, observed_y of row i                    */
  int32_t loss_type;         /* B200_LOSS_*  (examples/bundle_adjuster.cc:331-332)       */
  double loss_a;             /* HuberLoss(a)                                             */
  int32_t device;            /* CUDA device ordinal                                      */
  void* stream;              /* cudaStream_t to launch on; NULL = a private stream       */
  /* Multi-GPU: points (e blocks) are sharded, cameras replicated.  world_size==1 or
   * nccl_unique_id==NULL means single GPU.  nccl_unique_id = the 128 bytes of ncclUniqueId from
   * b200_nccl_unique_id() on rank 0, distributed by the caller (torch.distributed / MPI / files). */
  int32_t rank, world_size;
  const void* nccl_unique_id;
} b200_ba_desc;

/* The point order the library would keep privately for this structure (b200_create re-orders points -- with all their rows --
 * so that neighbouring points see the same few cameras, and undoes the permutation at every entry point: the layout above is
 * what the caller sees).  Host-only, needs no GPU: perm_out[k] = caller's e block at internal position k ([P], may be NULL);
 * metrics_out = distinct cameras per 1/num_chunks of the rows, summed, for {caller's order, by camera arc, by mean camera,
 * by smallest camera}; *choice_out = which of the four was taken (0 = the caller's order is kept). */
int b200_plan_point_order(const b200_ba_desc* desc, int num_chunks, int32_t* perm_out, int64_t metrics_out[4], int* choice_out);
int b200_nccl_unique_id(void* out128);
int b200_create(const b200_ba_desc* desc, b200_handle** out);
void b200_destroy(b200_handle* h);
const char* b200_last_error(void); /* thread-local message of the last failing call */
int b200_num_parameters(const b200_handle* h);     /* Evaluator::NumParameters   evaluator.h:151 */
int64_t b200_num_residuals(const b200_handle* h);  /* Evaluator::NumResiduals    evaluator.h:158 */

/* ---- Evaluator (internal/ceres/evaluator.h:116-121, ProgramEvaluator::Evaluate program_evaluator.h:137-304)
 * residuals / gradient may be NULL; want_jacobian != 0 refreshes the device-resident Jacobian.
 * gradient = J'r of the unscaled Jacobian.  Returns B200_ERR_EVALUATION_FAILED where Evaluate returns false. */
int b200_evaluate(b200_handle* h, const double* state, double* cost, double* residuals, double* gradient,
                  int want_jacobian);
/* Evaluator::EvaluateOptions::apply_loss_function (evaluator.h:101-102): apply == 0 makes the following evaluations
 * skip the robust correction (rho, Corrector) of the loss given at b200_create; apply != 0 (the default) restores it. */
int b200_set_apply_loss_function(b200_handle* h, int apply);
/* Evaluator::Plus (evaluator.h:146; Euclidean manifolds only): x_plus_delta = x + delta. */
int b200_plus(b200_handle* h, const double* x, const double* delta, double* x_plus_delta);

/* ---- SparseMatrix virtuals the minimizer calls on the Jacobian (internal/ceres/sparse_matrix.h:67-116) */
int b200_jacobian_squared_column_norm(b200_handle* h, double* x);            /* block_sparse_matrix.cc:351-401 */
int b200_jacobian_scale_columns(b200_handle* h, const double* scale);        /* :403-450 */
int b200_jacobian_right_multiply(b200_handle* h, const double* x, double* y);/* y += J x,  :239-274 */
int b200_jacobian_left_multiply(b200_handle* h, const double* x, double* y); /* y += J' x, :278-349 */
/* The one use the minimizer has for J*step, fused: model_cost_change = -(J step)'(r + J step / 2) with r the residuals
 * of the last b200_evaluate (still in HBM) -- trust_region_minimizer.cc:430-438 (ParallelSetZero +
 * RightMultiplyAndAccumulate + Dot) in one pass over J, returning one scalar instead of the 2N-vector J*step. */
int b200_model_cost_change(b200_handle* h, const double* step, double* model_cost_change);
int b200_jacobian_get_values(b200_handle* h, double* values);                /* BlockSparseMatrix::values(), 24N */
int b200_jacobian_set_values(b200_handle* h, const double* values);          /* mutable_values() */
/* The four single products of PartitionedMatrixView<2,3,9> (internal/ceres/partitioned_matrix_view_impl.h):
 *   B200_PMV_RIGHT_E  y[2N] += E x[3P]   (RightMultiplyAndAccumulateE, :113-137)
 *   B200_PMV_RIGHT_F  y[2N] += F x[9C]   (RightMultiplyAndAccumulateF, :140-191)
 *   B200_PMV_LEFT_E   y[3P] += E' x[2N]  (LeftMultiplyAndAccumulateE,  :194-264)
 *   B200_PMV_LEFT_F   y[9C] += F' x[2N]  (LeftMultiplyAndAccumulateF,  :267-375)
 * On the solver path these only run fused (b200_schur_multiply, b200_jtj_multiply); stand-alone they are the 2x3 / 2x9
 * block-SpMV shapes of the benchmark sweep.  Single GPU. */
enum { B200_PMV_RIGHT_E = 0, B200_PMV_RIGHT_F = 1, B200_PMV_LEFT_E = 2, B200_PMV_LEFT_F = 3 };
int b200_partitioned_multiply(b200_handle* h, int op, const double* x, double* y);
/* y = (J'J + diag(D)^2) x in one pass over J (D may be NULL).  The normal-equations product CGNR uses
 * (cgnr_solver.cc:90-115); here it is the north-star bandwidth kernel. */
int b200_jtj_multiply(b200_handle* h, const double* x, const double* D, double* y);

/* ---- LinearSolver (internal/ceres/linear_solver.h:339-342; IterativeSchurComplementSolver::SolveImpl,
 * iterative_schur_complement_solver.cc:64-157).  Solves min |J x - b|^2 + |D x|^2. */
typedef struct b200_solver_options { /* LinearSolver::Options + PerSolveOptions, linear_solver.h:150-315 */
  int32_t preconditioner_type;       /* B200_PRECOND_* */
  int32_t min_num_iterations;
  int32_t max_num_iterations;
  int32_t residual_reset_period;     /* linear_solver.h:211 (10) */
  double q_tolerance;                /* PerSolveOptions::q_tolerance (eta) */
  double r_tolerance;                /* PerSolveOptions::r_tolerance (-1 from LM) */
  int32_t max_num_spse_iterations;   /* linear_solver.h:172 (5): terms of the power series */
  int32_t use_spse_initialization;   /* :177 (0): start the PCG from the power series applied to the rhs
                                        (iterative_schur_complement_solver.cc:100-111) */
  double spse_tolerance;             /* :183 (0.1): early stop of that initialisation */
} b200_solver_options;
typedef struct b200_solver_summary { /* LinearSolver::Summary, linear_solver.h:320-326 */
  double residual_norm;
  int32_t num_iterations;
  int32_t termination_type;          /* B200_LS_* */
} b200_solver_summary;
void b200_solver_options_default(b200_solver_options* o);
/* b == NULL: b is the residual vector the last b200_evaluate produced, which is still in HBM (the minimizer passes
 * exactly that vector, trust_region_minimizer.cc:399-402 via levenberg_marquardt_strategy.cc:116; the adapter
 * compares the pointer with the one it filled in Evaluate and skips the 16N-byte upload). */
int b200_schur_solve(b200_handle* h, const double* b, const double* D, const b200_solver_options* opts,
                     double* x, b200_solver_summary* summary);

/* DenseSchurComplementSolver::SolveImpl (schur_complement_solver.cc:101-159, :161-214): explicit reduced camera system
 * S (dense 9C x 9C, assembled on the device), Cholesky (cuSOLVER potrf/potrs, loaded lazily), back substitution.
 * Single GPU, 9C up to ~75k.  summary: num_iterations 1, SUCCESS or FAILURE (S not positive definite). b == NULL as above. */
int b200_dense_schur_solve(b200_handle* h, const double* b, const double* D, double* x, b200_solver_summary* summary);

/* Finer-grained pieces of the same solve, for parity tests (each mirrors one reference class):
 *   ImplicitSchurComplement::Init / rhs / RightMultiplyAndAccumulate / BackSubstitute
 *     (implicit_schur_complement.cc:49-97, :251-276, :106-144, :208-243)
 *   SchurJacobiPreconditioner::UpdateImpl (schur_jacobi_preconditioner.cc:87-97) */
int b200_schur_init(b200_handle* h, const double* b, const double* D);
int b200_schur_rhs(b200_handle* h, double* rhs);                                /* [9C] */
int b200_schur_ete_inverse(b200_handle* h, double* out);                        /* [9P]: (E'E + D_e^2)^-1 */
int b200_schur_multiply(b200_handle* h, const double* x, double* y);            /* y = S x, [9C] */
int b200_schur_back_substitute(b200_handle* h, const double* z, double* y);     /* y [3P+9C] */
int b200_schur_jacobi_update(b200_handle* h, double* blocks, double* inverse);  /* each [81C], may be NULL */
int b200_block_jacobi_update(b200_handle* h, double* inverse);                  /* JACOBI: (F'F + D_f^2)^-1 blocks, [81C] */

/* ---- Device-resident trust-region loop (TrustRegionMinimizer::Minimize with
 * LevenbergMarquardtStrategy, trust_region_minimizer.cc:68-137 / levenberg_marquardt_strategy.cc:69-171).
 * State, residuals, Jacobian, D and the step never leave HBM; only scalars cross the bus. */
typedef struct b200_lm_options { /* Solver::Options subset, include/ceres/solver.h:232-632 */
  int32_t max_num_iterations;              /* bundle_adjuster.cc:121 (5) */
  int32_t jacobi_scaling;                  /* 1 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t linear_solver_type;        /* B200_ITERATIVE_SCHUR (default) or B200_DENSE_SCHUR */
  double eta;                              /* 1e-2 */
  double initial_trust_region_radius;      /* 1e4 */
  double max_trust_region_radius;          /* 1e16 */
  double min_trust_region_radius;          /* 1e-32 */
  double min_relative_decrease;            /* 1e-3 */
  double min_lm_diagonal, max_lm_diagonal; /* 1e-6, 1e32 */
  double function_tolerance, gradient_tolerance, parameter_tolerance; /* 1e-16 each in bundle_adjuster */
  b200_solver_options linear_solver;
} b200_lm_options;
typedef struct b200_lm_iteration { /* IterationSummary, include/ceres/iteration_callback.h */
  int32_t iteration, linear_solver_iterations, step_is_valid, step_is_successful;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease,
      trust_region_radius, model_cost_change;
} b200_lm_iteration;
void b200_lm_options_default(b200_lm_options* o);
/* state_inout: host [3P+9C] (read at entry, best state written back at exit).  trace: up to max_records
 * IterationSummary rows; returns the number written through *num_records.  If host_boundary != 0 the loop
 * is driven through the HOST-buffer entry points above exactly as the Ceres adapters would
 * (state/D/step/residual copies every iteration); otherwise everything stays device-resident. */
int b200_lm_solve(b200_handle* h, const b200_lm_options* opts, double* state_inout, b200_lm_iteration* trace,
                  int max_records, int* num_records, int host_boundary);

/* ---- Instrumentation */
typedef struct b200_kernel_stat {
  char name[32];
  int64_t launches;       /* kernel launches */
  int64_t operations;     /* logical operations (one S*x, one evaluate ...): an operation may take several launches
                             (main kernel + the few >32-row points + ...), all billed to it */
  double device_ms;       /* sum of CUDA-event times of all launches; only filled while profiling is enabled */
  double bytes_per_operation; /* algorithmic bytes moved by ONE OPERATION, 0 if not HBM-bound work:
                                 achieved GB/s = bytes_per_operation * operations / device_ms */
} b200_kernel_stat;
int b200_profile_enable(b200_handle* h, int on);   /* per-kernel cudaEvent timing on/off (off by default) */
int b200_stats_reset(b200_handle* h);
int b200_stats_get(b200_handle* h, b200_kernel_stat* out, int max_entries, int* num_entries);
int64_t b200_total_launches(const b200_handle* h); /* kernels launched since create / last reset */
int b200_synchronize(b200_handle* h);
/* h2d / d2h bytes moved by the host-buffer entry points since the last reset */
int b200_transfer_bytes(const b200_handle* h, int64_t* h2d, int64_t* d2h);

#ifdef __cplusplus
}
#endif
#endif /* B200BA_H_ */
