/* A plain C caller of libb200ba.so (include/b200ba.h): what a non-Python host -- the Ceres adapters, in C++ -- does.
 *   c_abi_smoke <problem.bin>
 * problem.bin (written by tests/test_c_abi_smoke.py): int32 C, P; int64 N; int32 cam_idx[N], pt_idx[N]; double obs[2N],
 * state[3P+9C].  Evaluates the problem, scales nothing, solves the damped normal equations once with ITERATIVE_SCHUR +
 * SCHUR_JACOBI and runs two LM iterations through the host-buffer boundary; prints one line per result for the test
 * to compare with the oracle.  Build: gcc -std=c99 -I include tests/c_abi_smoke.c -L ceres_solver_b200 -lb200ba -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200ba.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != B200_OK) {                                                    \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, b200_last_error());      \
      return 2;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t C, P;
  int64_t N;
  if (fread(&C, 4, 1, f) != 1 || fread(&P, 4, 1, f) != 1 || fread(&N, 8, 1, f) != 1) return 1;
  int32_t* cam = malloc(sizeof(int32_t) * N);
  int32_t* pt = malloc(sizeof(int32_t) * N);
  double* obs = malloc(sizeof(double) * 2 * N);
  const int np = 3 * P + 9 * C;
  double* state = malloc(sizeof(double) * np);
  if (fread(cam, 4, N, f) != (size_t)N || fread(pt, 4, N, f) != (size_t)N || fread(obs, 8, 2 * N, f) != (size_t)(2 * N) ||
      fread(state, 8, np, f) != (size_t)np)
    return 1;
  fclose(f);

  b200_ba_desc desc = {0};
  desc.num_cameras = C;
  desc.num_points = P;
  desc.num_observations = N;
  desc.cam_idx = cam;
  desc.pt_idx = pt;
  desc.obs = obs;
  desc.loss_type = B200_LOSS_TRIVIAL;
  desc.loss_a = 1.0;
  desc.world_size = 1;
  b200_handle* h = NULL;
  int rc = b200_create(&desc, &h);
  if (rc != B200_OK) {
    fprintf(stderr, "b200_create -> %d: %s\n", rc, b200_last_error());
    b200_destroy(h);
    return rc == B200_ERR_NO_DEVICE ? 3 : 2;
  }
  if (b200_num_parameters(h) != np || b200_num_residuals(h) != 2 * N) return 2;

  double cost = 0.0;
  double* residuals = malloc(sizeof(double) * 2 * N);
  double* gradient = malloc(sizeof(double) * np);
  CHECK(b200_evaluate(h, state, &cost, residuals, gradient, 1));
  double gmax = 0.0;
  for (int i = 0; i < np; ++i) gmax = fmax(gmax, fabs(gradient[i]));
  printf("cost %.17g\ngradient_max_norm %.17g\n", cost, gmax);

  /* LM diagonal of the unscaled Jacobian at radius 1e4 (levenberg_marquardt_strategy.cc:84-95) */
  double* D = malloc(sizeof(double) * np);
  CHECK(b200_jacobian_squared_column_norm(h, D));
  for (int i = 0; i < np; ++i) D[i] = sqrt(fmin(fmax(D[i], 1e-6), 1e32) / 1e4);
  b200_solver_options so;
  b200_solver_options_default(&so);
  so.q_tolerance = 1e-2;
  so.r_tolerance = -1.0;
  b200_solver_summary sum;
  double* x = malloc(sizeof(double) * np);
  CHECK(b200_schur_solve(h, residuals, D, &so, x, &sum));
  double xn = 0.0;
  for (int i = 0; i < np; ++i) xn += x[i] * x[i];
  printf("solve iterations %d termination %d step_norm %.17g\n", sum.num_iterations, sum.termination_type, sqrt(xn));

  b200_lm_options lo;
  b200_lm_options_default(&lo);
  lo.max_num_iterations = 2;
  b200_lm_iteration trace[8];
  int nrec = 0;
  CHECK(b200_lm_solve(h, &lo, state, trace, 8, &nrec, /*host_boundary=*/1));
  for (int i = 0; i < nrec; ++i)
    printf("lm %d cost %.17g step_norm %.17g cg %d\n", trace[i].iteration, trace[i].cost, trace[i].step_norm,
           trace[i].linear_solver_iterations);
  b200_destroy(h);
  printf("done\n");
  return 0;
}
