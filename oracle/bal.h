// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// CPU restatement of the bundle-adjustment front end of the reference's hot path:
//   BalProblem      examples/bal_problem.cc:73-133 (reader), :216-247 (camera<->centre),
//                   :249-292 (Normalize)
//   Jet             include/ceres/jet.h:298-400, :614-630, :735-748 (12-wide dual numbers)
//   Snavely         examples/snavely_reprojection_error.h:57-92 +
//                   include/ceres/rotation.h:864-930 (AngleAxisRotatePoint)
//   Huber/Corrector internal/ceres/loss_function.cc:52-66, corrector.cc:41-155,
//                   residual_block.cc:70-198
//   BaProgram       parameter/residual ordering: reorder_program.cc:217-276, :278-359;
//                   Jacobian layout: block_jacobian_writer.cc:68-167, :198-263
//   Evaluate        program_evaluator.h:137-304
//   Minimize        trust_region_minimizer.cc:68-137,246-304,381-462,726-845,
//                   levenberg_marquardt_strategy.cc:69-171, trust_region_step_evaluator.cc:49-109
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include "schur.h"

namespace orc {

// ---------------------------------------------------------------------------- Jet
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  explicit Jet(double value) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Jet(double value, int k) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
    v[k] = 1.0;
  }
};
#define ORC_JET_LOOP for (int i = 0; i < N; ++i)
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> r; r.a = -f.a; ORC_JET_LOOP r.v[i] = -f.v[i]; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a + g.a; ORC_JET_LOOP r.v[i] = f.v[i] + g.v[i]; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> r = f; r.a = f.a + s; return r; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> r = f; r.a = f.a + s; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a - g.a; ORC_JET_LOOP r.v[i] = f.v[i] - g.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> r = f; r.a = f.a - s; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a * g.a; ORC_JET_LOOP r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  Jet<N> r;
  r.a = f_a_by_g_a;
  ORC_JET_LOOP r.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return r;
}
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N> cos(const Jet<N>& f) { Jet<N> r; r.a = std::cos(f.a); const double s = -std::sin(f.a); ORC_JET_LOOP r.v[i] = s * f.v[i]; return r; }
template <int N> inline Jet<N> sin(const Jet<N>& f) { Jet<N> r; r.a = std::sin(f.a); const double c = std::cos(f.a); ORC_JET_LOOP r.v[i] = c * f.v[i]; return r; }
template <int N> inline Jet<N> hypot(const Jet<N>& x, const Jet<N>& y, const Jet<N>& z) {
  const double tmp = std::hypot(x.a, y.a, z.a);
  Jet<N> r;
  r.a = tmp;
  ORC_JET_LOOP r.v[i] = x.a / tmp * x.v[i] + y.a / tmp * y.v[i] + z.a / tmp * z.v[i];
  return r;
}
inline double hypot(double x, double y, double z) { return std::hypot(x, y, z); }
inline double ScalarOf(double x) { return x; }
template <int N> inline double ScalarOf(const Jet<N>& x) { return x.a; }
template <typename T> inline T FromScalar(double x) { return T(x); }

// ---------------------------------------------------------------------------- rotation.h:864-930
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  using std::cos;
  using std::sin;
  const T theta = hypot(angle_axis[0], angle_axis[1], angle_axis[2]);
  if (std::fpclassify(ScalarOf(theta)) != FP_ZERO) {
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = FromScalar<T>(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (FromScalar<T>(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1],
                             angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// ---------------------------------------------------------------------------- snavely_reprojection_error.h:57-92
template <typename T>
inline void SnavelyResidual(const T* camera, const T* point, double observed_x, double observed_y, T* residuals) {
  T p[3];
  AngleAxisRotatePoint(camera, point, p);
  p[0] += camera[3];
  p[1] += camera[4];
  p[2] += camera[5];
  const T xp = -p[0] / p[2];
  const T yp = -p[1] / p[2];
  const T& l1 = camera[7];
  const T& l2 = camera[8];
  const T r2 = xp * xp + yp * yp;
  const T distortion = 1.0 + r2 * (l1 + l2 * r2);
  const T& focal = camera[6];
  const T predicted_x = focal * distortion * xp;
  const T predicted_y = focal * distortion * yp;
  residuals[0] = predicted_x - observed_x;
  residuals[1] = predicted_y - observed_y;
}

// ---------------------------------------------------------------------------- BAL file
struct BalProblem {
  int num_cameras = 0, num_points = 0, num_observations = 0;
  std::vector<int> camera_index, point_index;
  std::vector<double> observations;  // 2N
  std::vector<double> parameters;    // 9C cameras then 3P points (file order)

  double* cameras() { return parameters.data(); }
  double* points() { return parameters.data() + 9 * num_cameras; }

  // bal_problem.cc:73-133
  bool Read(const char* filename) {
    FILE* f = std::fopen(filename, "r");
    if (f == nullptr) return false;
    bool ok = std::fscanf(f, "%d %d %d", &num_cameras, &num_points, &num_observations) == 3;
    if (ok) {
      camera_index.resize(num_observations);
      point_index.resize(num_observations);
      observations.resize(2 * static_cast<size_t>(num_observations));
      parameters.resize(9 * static_cast<size_t>(num_cameras) + 3 * static_cast<size_t>(num_points));
      for (int i = 0; ok && i < num_observations; ++i)
        ok = std::fscanf(f, "%d %d %lf %lf", &camera_index[i], &point_index[i], &observations[2 * i],
                         &observations[2 * i + 1]) == 4;
      for (size_t i = 0; ok && i < parameters.size(); ++i) ok = std::fscanf(f, "%lf", &parameters[i]) == 1;
    }
    std::fclose(f);
    return ok;
  }

  static double Median(std::vector<double>* data) {  // bal_problem.cc:64-68 (upper median)
    auto mid = data->begin() + data->size() / 2;
    std::nth_element(data->begin(), mid, data->end());
    return *mid;
  }
  // bal_problem.cc:216-247
  static void CameraToAngleAxisAndCenter(const double* camera, double* angle_axis, double* center) {
    for (int i = 0; i < 3; ++i) angle_axis[i] = camera[i];
    double inverse_rotation[3] = {-angle_axis[0], -angle_axis[1], -angle_axis[2]};
    AngleAxisRotatePoint(inverse_rotation, camera + 3, center);
    for (int i = 0; i < 3; ++i) center[i] *= -1.0;
  }
  static void AngleAxisAndCenterToCamera(const double* angle_axis, const double* center, double* camera) {
    for (int i = 0; i < 3; ++i) camera[i] = angle_axis[i];
    AngleAxisRotatePoint(angle_axis, center, camera + 3);
    for (int i = 0; i < 3; ++i) camera[3 + i] *= -1.0;
  }
  // bal_problem.cc:249-292
  void Normalize() {
    std::vector<double> tmp(num_points);
    double median[3];
    double* pts = points();
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < num_points; ++j) tmp[j] = pts[3 * j + i];
      median[i] = Median(&tmp);
    }
    for (int i = 0; i < num_points; ++i)
      tmp[i] = std::fabs(pts[3 * i] - median[0]) + std::fabs(pts[3 * i + 1] - median[1]) +
               std::fabs(pts[3 * i + 2] - median[2]);
    const double median_absolute_deviation = Median(&tmp);
    const double scale = 100.0 / median_absolute_deviation;
    for (int i = 0; i < num_points; ++i)
      for (int k = 0; k < 3; ++k) pts[3 * i + k] = scale * (pts[3 * i + k] - median[k]);
    double* cams = cameras();
    for (int i = 0; i < num_cameras; ++i) {
      double* camera = cams + 9 * i;
      double angle_axis[3], center[3];
      CameraToAngleAxisAndCenter(camera, angle_axis, center);
      for (int k = 0; k < 3; ++k) center[k] = scale * (center[k] - median[k]);
      AngleAxisAndCenterToCamera(angle_axis, center, camera);
    }
  }
};

// HuberLoss::Evaluate (loss_function.cc:52-66): rho = {rho(s), rho'(s), rho''(s)}, s = squared residual norm.
inline void HuberLossEvaluate(double a, double s, double rho[3]) {
  const double b = a * a;  // loss_function.h: HuberLoss(a): a_(a), b_(a * a)
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s;
    rho[1] = 1.0;
    rho[2] = 0.0;
  }
}

// Corrector (corrector.cc:41-155): rescales residuals and Jacobian so that the Gauss-Newton step of the corrected
// problem is the robustified one.
struct Corrector {
  double sqrt_rho1, residual_scaling, alpha_sq_norm;
  Corrector(double sq_norm, const double rho[3]) {
    sqrt_rho1 = std::sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) {  // :89-113
      residual_scaling = sqrt_rho1;
      alpha_sq_norm = 0.0;
      return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  void CorrectResiduals(int num_rows, double* residuals) const {  // :115-123
    for (int r = 0; r < num_rows; ++r) residuals[r] *= residual_scaling;
  }
  // jacobian: num_rows x num_cols row-major; residuals: the UNCORRECTED ones (:125-155)
  void CorrectJacobian(int num_rows, int num_cols, const double* residuals, double* jacobian) const {
    if (alpha_sq_norm == 0.0) {
      for (int k = 0; k < num_rows * num_cols; ++k) jacobian[k] *= sqrt_rho1;
      return;
    }
    for (int c = 0; c < num_cols; ++c) {
      double r_transpose_j = 0.0;
      for (int r = 0; r < num_rows; ++r) r_transpose_j += jacobian[r * num_cols + c] * residuals[r];
      for (int r = 0; r < num_rows; ++r)
        jacobian[r * num_cols + c] = sqrt_rho1 * (jacobian[r * num_cols + c] - alpha_sq_norm * residuals[r] * r_transpose_j);
    }
  }
};

// ---------------------------------------------------------------------------- BA program
struct BaProgram {
  int C = 0, P = 0, N = 0;
  // Program order (reduced program after ApplyOrdering + LexicographicallyOrderResidualBlocks):
  std::vector<int> point_of_eblock;   // e block j  -> input point id
  std::vector<int> camera_of_fblock;  // f block k  -> input camera id
  std::vector<int> obs_of_row;        // row i      -> input observation id
  std::vector<int> row_pt, row_cam;   // row i -> e block id, f block id (0-based within F)
  std::vector<double> row_obs;        // 2 per row
  bool use_huber = false;
  double huber_a = 1.0;
  int num_threads = 1;
  BlockSparseMatrix jacobian;         // structure built once (block_jacobian_writer.cc:198-263)

  int num_parameters() const { return 3 * P + 9 * C; }
  int num_residuals() const { return 2 * N; }

  void Build(int c, int p, int n, const int* cam_idx, const int* pt_idx, const double* obs) {
    C = c;
    P = p;
    N = n;
    // Problem::AddResidualBlock(cost, loss, camera, point) registers parameter blocks in first-use
    // order (camera, then point, per observation); ApplyOrdering keeps that order inside each
    // elimination group (reorder_program.cc:247-273): group 0 = points, group 1 = cameras.
    std::vector<int> e_of_point(P, -1), f_of_camera(C, -1);
    point_of_eblock.clear();
    camera_of_fblock.clear();
    for (int i = 0; i < N; ++i) {
      if (f_of_camera[cam_idx[i]] < 0) {
        f_of_camera[cam_idx[i]] = static_cast<int>(camera_of_fblock.size());
        camera_of_fblock.push_back(cam_idx[i]);
      }
      if (e_of_point[pt_idx[i]] < 0) {
        e_of_point[pt_idx[i]] = static_cast<int>(point_of_eblock.size());
        point_of_eblock.push_back(pt_idx[i]);
      }
    }
    P = static_cast<int>(point_of_eblock.size());   // unused blocks are dropped by the reduced program
    C = static_cast<int>(camera_of_fblock.size());
    // LexicographicallyOrderResidualBlocks (reorder_program.cc:278-359): bucket by e block, each
    // bucket filled back to front.
    std::vector<int> offsets(P + 1, 0);
    for (int i = 0; i < N; ++i) offsets[e_of_point[pt_idx[i]]]++;
    for (int j = 1; j <= P; ++j) offsets[j] += offsets[j - 1];
    obs_of_row.assign(N, -1);
    for (int i = 0; i < N; ++i) obs_of_row[--offsets[e_of_point[pt_idx[i]]]] = i;
    row_pt.resize(N);
    row_cam.resize(N);
    row_obs.resize(2 * static_cast<size_t>(N));
    for (int r = 0; r < N; ++r) {
      const int i = obs_of_row[r];
      row_pt[r] = e_of_point[pt_idx[i]];
      row_cam[r] = f_of_camera[cam_idx[i]];
      row_obs[2 * r] = obs[2 * i];
      row_obs[2 * r + 1] = obs[2 * i + 1];
    }
    // Jacobian structure: all E cells first, then all F cells (block_jacobian_writer.cc:68-167).
    BlockStructure& bs = jacobian.bs;
    bs.cols.resize(P + C);
    for (int j = 0; j < P; ++j) { bs.cols[j].size = 3; bs.cols[j].position = 3 * j; }
    for (int k = 0; k < C; ++k) { bs.cols[P + k].size = 9; bs.cols[P + k].position = 3 * P + 9 * k; }
    bs.rows.resize(N);
    for (int r = 0; r < N; ++r) {
      bs.rows[r].block.size = 2;
      bs.rows[r].block.position = 2 * r;
      bs.rows[r].cells.resize(2);
      bs.rows[r].cells[0].block_id = row_pt[r];
      bs.rows[r].cells[0].position = 6 * r;
      bs.rows[r].cells[1].block_id = P + row_cam[r];
      bs.rows[r].cells[1].position = 6 * N + 18 * r;
    }
    jacobian.Finalize();
  }

  // state = [points in e-block order ; cameras in f-block order]
  void StateFromParameters(const double* cameras, const double* points, double* state) const {
    for (int j = 0; j < P; ++j)
      for (int k = 0; k < 3; ++k) state[3 * j + k] = points[3 * point_of_eblock[j] + k];
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < 9; ++k) state[3 * P + 9 * c + k] = cameras[9 * camera_of_fblock[c] + k];
  }
  void ParametersFromState(const double* state, double* cameras, double* points) const {
    for (int j = 0; j < P; ++j)
      for (int k = 0; k < 3; ++k) points[3 * point_of_eblock[j] + k] = state[3 * j + k];
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < 9; ++k) cameras[9 * camera_of_fblock[c] + k] = state[3 * P + 9 * c + k];
  }

  // ResidualBlock::Evaluate (residual_block.cc:70-198) for one row.
  // jac_cam (2x9) / jac_pt (2x3) may be null together. Returns false on non-finite output.
  bool EvaluateRow(int r, const double* state, double* cost, double* residuals, double* jac_cam,
                   double* jac_pt) const {
    const double* camera = state + 3 * P + 9 * row_cam[r];
    const double* point = state + 3 * row_pt[r];
    const double ox = row_obs[2 * r], oy = row_obs[2 * r + 1];
    double res[2];
    if (jac_cam == nullptr) {
      SnavelyResidual<double>(camera, point, ox, oy, res);
    } else {
      // AutoDifferentiate<2, (9,3)> (include/ceres/internal/autodiff.h:252-312)
      typedef Jet<12> J;
      J cam[9], pt[3], out[2];
      for (int k = 0; k < 9; ++k) cam[k] = J(camera[k], k);
      for (int k = 0; k < 3; ++k) pt[k] = J(point[k], 9 + k);
      SnavelyResidual<J>(cam, pt, ox, oy, out);
      for (int i = 0; i < 2; ++i) {
        res[i] = out[i].a;
        for (int k = 0; k < 9; ++k) jac_cam[i * 9 + k] = out[i].v[k];
        for (int k = 0; k < 3; ++k) jac_pt[i * 3 + k] = out[i].v[9 + k];
      }
      for (int k = 0; k < 18; ++k)
        if (!std::isfinite(jac_cam[k])) return false;
      for (int k = 0; k < 6; ++k)
        if (!std::isfinite(jac_pt[k])) return false;
    }
    if (!std::isfinite(res[0]) || !std::isfinite(res[1])) return false;
    const double squared_norm = res[0] * res[0] + res[1] * res[1];
    if (!use_huber) {
      *cost = 0.5 * squared_norm;
    } else {
      double rho[3];
      HuberLossEvaluate(huber_a, squared_norm, rho);
      *cost = 0.5 * rho[0];
      if (jac_cam != nullptr || residuals != nullptr) {
        const Corrector correct(squared_norm, rho);   // residual_block.cc:170-195
        if (jac_cam != nullptr) {
          correct.CorrectJacobian(2, 9, res, jac_cam);
          correct.CorrectJacobian(2, 3, res, jac_pt);
        }
        correct.CorrectResiduals(2, res);
      }
    }
    if (residuals != nullptr) {
      residuals[0] = res[0];
      residuals[1] = res[1];
    }
    return true;
  }

  // ProgramEvaluator::Evaluate (program_evaluator.h:137-304). gradient is J'r of the UNSCALED J.
  bool Evaluate(const double* state, double* cost, double* residuals, double* gradient, bool want_jacobian) {
    const int nt = std::max(1, num_threads);
    const int np = num_parameters();
    std::vector<double> thread_cost(nt, 0.0);
    std::vector<std::vector<double>> thread_gradient;
    if (gradient != nullptr) thread_gradient.assign(nt, std::vector<double>(np, 0.0));
    if (want_jacobian) jacobian.SetZero();
    std::atomic<bool> abort(false);
    double* values = jacobian.values.data();
    ParallelFor(0, N, nt, [&](int tid, int r) {
      if (abort) return;
      double block_cost, res[2], jc[18], jp[6];
      const bool need_j = want_jacobian || gradient != nullptr;
      double* jac_pt = need_j ? (want_jacobian ? values + 6 * static_cast<size_t>(r) : jp) : nullptr;
      double* jac_cam = need_j ? (want_jacobian ? values + 6 * static_cast<size_t>(N) + 18 * static_cast<size_t>(r) : jc) : nullptr;
      double* block_res = residuals != nullptr ? residuals + 2 * static_cast<size_t>(r) : (gradient != nullptr ? res : nullptr);
      if (!EvaluateRow(r, state, &block_cost, block_res, jac_cam, jac_pt)) {
        abort = true;
        return;
      }
      thread_cost[tid] += block_cost;
      if (gradient != nullptr) {
        // parameter block order inside the residual block is (camera, point)
        MatrixTransposeVectorMultiply<kDyn, kDyn, 1>(jac_cam, 2, 9, block_res, thread_gradient[tid].data() + 3 * P + 9 * row_cam[r]);
        MatrixTransposeVectorMultiply<kDyn, kDyn, 1>(jac_pt, 2, 3, block_res, thread_gradient[tid].data() + 3 * row_pt[r]);
      }
    });
    if (abort) return false;
    *cost = 0.0;
    if (gradient != nullptr)
      for (int i = 0; i < np; ++i) gradient[i] = 0.0;
    for (int t = 0; t < nt; ++t) {
      *cost += thread_cost[t];
      if (gradient != nullptr)
        for (int i = 0; i < np; ++i) gradient[i] += thread_gradient[t][i];
    }
    return std::isfinite(*cost);
  }
};

// ---------------------------------------------------------------------------- LM minimizer
enum LinearSolverKind { ITERATIVE_SCHUR = 0, DENSE_SCHUR = 1 };

struct SolveOptions {
  int linear_solver = ITERATIVE_SCHUR;
  int preconditioner = SCHUR_JACOBI;
  int max_num_iterations = 5;               // bundle_adjuster.cc:121
  int max_linear_solver_iterations = 500;   // bundle_adjuster.cc:122
  int min_linear_solver_iterations = 0;
  double eta = 1e-2;                        // bundle_adjuster.cc:116
  double initial_trust_region_radius = 1e4; // include/ceres/solver.h
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  double function_tolerance = 1e-16;        // bundle_adjuster.cc:370-372
  double gradient_tolerance = 1e-16;
  double parameter_tolerance = 1e-16;
  int jacobi_scaling = 1;
  int max_num_consecutive_invalid_steps = 5;
  int num_threads = 1;
  int use_spse_initialization = 0;          // solver.h:612 (max_num_spse_iterations 5, spse_tolerance 0.1: :607, :618)
};

struct IterationRecord {
  int iteration;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease,
      trust_region_radius, model_cost_change;
  int linear_solver_iterations, step_is_valid, step_is_successful;
};

struct SolveTimes {
  double residual_eval = 0, jacobian_eval = 0, linear_solver = 0, total = 0;
  int num_residual_evals = 0, num_jacobian_evals = 0, num_linear_solves = 0;
};

inline double NowSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy, monotonic steps, no bounds,
// no inner iterations (the bundle_adjuster defaults).
inline int Minimize(BaProgram* program, const SolveOptions& opt, double* state_inout,
                    std::vector<IterationRecord>* trace, SolveTimes* times) {
  const int nt = opt.num_threads;
  program->num_threads = nt;
  const int np = program->num_parameters();
  const int nr = program->num_residuals();
  BlockSparseMatrix* J = &program->jacobian;

  std::unique_ptr<LinearSolverBase> linear_solver;
  if (opt.linear_solver == ITERATIVE_SCHUR) {
    IterativeSchurOptions so;
    so.num_eliminate_blocks = program->P;
    so.preconditioner_type = opt.preconditioner;
    so.use_spse_initialization = opt.use_spse_initialization != 0;
    so.min_num_iterations = opt.min_linear_solver_iterations;
    so.max_num_iterations = opt.max_linear_solver_iterations;
    so.num_threads = nt;
    linear_solver.reset(new IterativeSchurSolver<2, 3, 9>(so));
  } else {
    linear_solver.reset(new DenseSchurSolver<2, 3, 9>(program->P, nt));
  }

  std::vector<double> x(state_inout, state_inout + np), candidate_x(np), residuals(nr), gradient(np),
      step(np), delta(np), model_residuals(nr), jacobian_scaling(np, 1.0), diagonal(np), lm_diagonal(np);
  double x_cost = std::numeric_limits<double>::max(), candidate_cost = 0, model_cost_change = 0;
  double minimum_cost = x_cost;
  std::vector<double> best(x);
  const double t_start = NowSeconds();

  // LevenbergMarquardtStrategy state (levenberg_marquardt_strategy.cc:47-65)
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  // TrustRegionStepEvaluator, max_consecutive_nonmonotonic_steps = 0
  double se_minimum_cost = 0, se_current_cost = 0, se_reference_cost = 0, se_candidate_cost = 0;
  double se_acc_reference = 0, se_acc_candidate = 0;

  IterationRecord it{};
  int iteration = 0;

  // EvaluateGradientAndJacobian (trust_region_minimizer.cc:246-304)
  auto evaluate_gradient_and_jacobian = [&]() -> bool {
    const double t0 = NowSeconds();
    const bool ok = program->Evaluate(x.data(), &x_cost, residuals.data(), gradient.data(), true);
    if (times) { times->jacobian_eval += NowSeconds() - t0; times->num_jacobian_evals++; }
    if (!ok) return false;
    it.cost = x_cost;
    if (opt.jacobi_scaling) {
      if (iteration == 0) {
        J->SquaredColumnNorm(jacobian_scaling.data(), nt);
        for (int i = 0; i < np; ++i) jacobian_scaling[i] = 1.0 / (1.0 + std::sqrt(jacobian_scaling[i]));
      }
      J->ScaleColumns(jacobian_scaling.data(), nt);
    }
    double mx = 0, sq = 0;  // |Plus(x,-g) - x| with Euclidean manifolds and no bounds = |g|
    for (int i = 0; i < np; ++i) {
      const double d = x[i] - (x[i] + (-gradient[i]));
      mx = std::max(mx, std::fabs(d));
      sq += d * d;
    }
    it.gradient_max_norm = mx;
    it.gradient_norm = std::sqrt(sq);
    return true;
  };

  // IterationZero (:187-233)
  it = IterationRecord{};
  it.iteration = 0;
  if (!evaluate_gradient_and_jacobian()) return FAILURE;
  it.step_is_valid = 1;
  it.step_is_successful = 1;
  se_minimum_cost = se_current_cost = se_reference_cost = se_candidate_cost = x_cost;

  int num_consecutive_invalid_steps = 0;
  bool atleast_one_successful_step = false;
  int result = NO_CONVERGENCE;

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (:316-361)
    if (it.step_is_successful) {
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        best = x;
      }
    }
    it.trust_region_radius = radius;
    trace->push_back(it);
    if (it.iteration >= opt.max_num_iterations) break;
    if (it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) { result = SUCCESS; break; }
    if (it.trust_region_radius <= opt.min_trust_region_radius) { result = SUCCESS; break; }

    const double previous_gradient_norm = it.gradient_norm;
    const double previous_gradient_max_norm = it.gradient_max_norm;
    const int prev_iteration = it.iteration;
    it = IterationRecord{};
    it.iteration = prev_iteration + 1;
    iteration = it.iteration;

    // ComputeTrustRegionStep (:381-462) -> LevenbergMarquardtStrategy::ComputeStep (:69-156)
    it.step_is_valid = 0;
    if (!reuse_diagonal) {
      J->SquaredColumnNorm(diagonal.data(), nt);
      for (int i = 0; i < np; ++i) diagonal[i] = std::min(std::max(diagonal[i], opt.min_lm_diagonal), opt.max_lm_diagonal);
    }
    for (int i = 0; i < np; ++i) lm_diagonal[i] = std::sqrt(diagonal[i] / radius);
    for (int i = 0; i < np; ++i) step[i] = std::numeric_limits<double>::quiet_NaN();  // InvalidateArray
    const double t0 = NowSeconds();
    LinearSummary ls = linear_solver->Solve(J, residuals.data(), lm_diagonal.data(), opt.eta, -1.0, step.data());
    if (times) { times->linear_solver += NowSeconds() - t0; times->num_linear_solves++; }
    if (ls.termination_type == FATAL_ERROR) return FAILURE;
    if (ls.termination_type != FAILURE) {
      bool valid = true;
      for (int i = 0; i < np; ++i) valid = valid && std::isfinite(step[i]);
      if (!valid) ls.termination_type = FAILURE;
      else for (int i = 0; i < np; ++i) step[i] = -step[i];
    }
    reuse_diagonal = true;
    it.linear_solver_iterations = ls.num_iterations;

    if (ls.termination_type != FAILURE) {
      std::fill(model_residuals.begin(), model_residuals.end(), 0.0);
      J->RightMultiplyAndAccumulate(step.data(), model_residuals.data(), nt);
      double dot = 0.0;
      for (int i = 0; i < nr; ++i) dot += model_residuals[i] * (residuals[i] + model_residuals[i] / 2.0);
      model_cost_change = -dot;
      it.model_cost_change = model_cost_change;
      it.step_is_valid = model_cost_change > 0.0;
      if (it.step_is_valid) {
        for (int i = 0; i < np; ++i) delta[i] = step[i] * jacobian_scaling[i];
        num_consecutive_invalid_steps = 0;
      }
    }

    if (!it.step_is_valid) {
      // HandleInvalidStep (:467-500)
      if (++num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) return FAILURE;
      // LevenbergMarquardtStrategy::StepIsInvalid == StepRejected(0.0) (levenberg_marquardt_strategy.h:63-69)
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      it.cost = x_cost;
      it.cost_change = 0.0;
      it.gradient_max_norm = trace->back().gradient_max_norm;
      it.gradient_norm = trace->back().gradient_norm;
      it.step_norm = 0.0;
      it.relative_decrease = 0.0;
      continue;
    }

    // ComputeCandidatePointAndEvaluateCost (:781-799)
    for (int i = 0; i < np; ++i) candidate_x[i] = x[i] + delta[i];
    {
      const double t1 = NowSeconds();
      if (!program->Evaluate(candidate_x.data(), &candidate_cost, nullptr, nullptr, false))
        candidate_cost = std::numeric_limits<double>::max();
      if (times) { times->residual_eval += NowSeconds() - t1; times->num_residual_evals++; }
    }

    // ParameterToleranceReached (:726-746)
    {
      double xn = 0, sn = 0;
      for (int i = 0; i < np; ++i) {
        xn += x[i] * x[i];
        sn += (x[i] - candidate_x[i]) * (x[i] - candidate_x[i]);
      }
      // IterationSummary::step_norm is only assigned inside ParameterToleranceReached(), which the minimizer does not call
      // before the first successful step (trust_region_minimizer.cc:113, :730): 0 until then (the "0.00" of iteration 1 in
      // the published transcripts)
      it.step_norm = atleast_one_successful_step ? std::sqrt(sn) : 0.0;
      if (atleast_one_successful_step &&
          it.step_norm <= opt.parameter_tolerance * (std::sqrt(xn) + opt.parameter_tolerance)) {
        result = SUCCESS;
        break;
      }
    }
    // FunctionToleranceReached (:749-769)
    it.cost_change = x_cost - candidate_cost;
    if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) { result = SUCCESS; break; }

    // IsStepSuccessful (:801-827), StepQuality (trust_region_step_evaluator.cc:49-65)
    if (candidate_cost >= std::numeric_limits<double>::max()) {
      it.relative_decrease = std::numeric_limits<double>::lowest();
    } else {
      const double relative_decrease = (se_current_cost - candidate_cost) / model_cost_change;
      const double historical = (se_reference_cost - candidate_cost) / (se_acc_reference + model_cost_change);
      it.relative_decrease = std::max(relative_decrease, historical);
    }
    if (it.relative_decrease > opt.min_relative_decrease) {
      atleast_one_successful_step = true;
      // HandleSuccessfulStep (:832-845)
      x = candidate_x;
      if (!evaluate_gradient_and_jacobian()) return FAILURE;
      it.step_is_successful = 1;
      // StepAccepted (levenberg_marquardt_strategy.cc:158-165)
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      // TrustRegionStepEvaluator::StepAccepted (:67-107) with max_consecutive_nonmonotonic_steps = 0
      se_current_cost = candidate_cost;
      se_acc_candidate += model_cost_change;
      se_acc_reference += model_cost_change;
      int nonmono = 0;
      if (se_current_cost < se_minimum_cost) {
        se_minimum_cost = se_current_cost;
        se_candidate_cost = se_current_cost;
        se_acc_candidate = 0.0;
      } else {
        ++nonmono;
        if (se_current_cost > se_candidate_cost) {
          se_candidate_cost = se_current_cost;
          se_acc_candidate = 0.0;
        }
      }
      if (nonmono == 0) {
        se_reference_cost = se_candidate_cost;
        se_acc_reference = se_acc_candidate;
      }
    } else {
      it.step_is_successful = 0;
      it.cost = candidate_cost;
      it.gradient_norm = previous_gradient_norm;
      it.gradient_max_norm = previous_gradient_max_norm;
      // StepRejected (:167-171)
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }
  if (times) times->total = NowSeconds() - t_start;
  for (int i = 0; i < np; ++i) state_inout[i] = best[i];
  return result;
}

}  // namespace orc
