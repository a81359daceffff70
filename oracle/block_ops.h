// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// Tiny dense block kernels with the exact op semantics of the reference's
// internal/ceres/small_blas.h:164-555 (row-major storage; kOperation in
// {+1,-1,0} meaning C += , C -= , C = ; results written into a sub-block of C
// at (start_row, start_col) with C's own strides).  Sizes may be compile-time
// (>0) or dynamic (kDyn), as in the reference, so the <2,3,9> specialisation
// gets fully unrolled by the compiler while the fixture problems with odd
// block sizes run through the same code.
#pragma once

namespace orc {

constexpr int kDyn = -1;

#define ORC_DIM(k, n) ((k) != kDyn ? (k) : (n))

// C[start_row.., start_col..] op= A * B       (small_blas.h:164-213)
template <int kRowA, int kColA, int kRowB, int kColB, int kOp>
inline void MatrixMatrixMultiply(const double* A, int num_row_a, int num_col_a,
                                 const double* B, int num_row_b, int num_col_b,
                                 double* C, int start_row_c, int start_col_c,
                                 int /*row_stride_c*/, int col_stride_c) {
  const int NRA = ORC_DIM(kRowA, num_row_a);
  const int NCA = ORC_DIM(kColA, num_col_a);
  const int NCB = ORC_DIM(kColB, num_col_b);
  (void)num_row_b;
  for (int row = 0; row < NRA; ++row) {
    for (int col = 0; col < NCB; ++col) {
      double tmp = 0.0;
      for (int k = 0; k < NCA; ++k) tmp += A[row * NCA + k] * B[k * NCB + col];
      const int index = (row + start_row_c) * col_stride_c + start_col_c + col;
      if (kOp > 0) C[index] += tmp;
      else if (kOp < 0) C[index] -= tmp;
      else C[index] = tmp;
    }
  }
}

// C[start_row.., start_col..] op= A' * B      (small_blas.h:215-263)
template <int kRowA, int kColA, int kRowB, int kColB, int kOp>
inline void MatrixTransposeMatrixMultiply(const double* A, int num_row_a, int num_col_a,
                                          const double* B, int num_row_b, int num_col_b,
                                          double* C, int start_row_c, int start_col_c,
                                          int /*row_stride_c*/, int col_stride_c) {
  const int NRA = ORC_DIM(kRowA, num_row_a);
  const int NCA = ORC_DIM(kColA, num_col_a);
  const int NCB = ORC_DIM(kColB, num_col_b);
  (void)num_row_b;
  for (int row = 0; row < NCA; ++row) {
    for (int col = 0; col < NCB; ++col) {
      double tmp = 0.0;
      for (int k = 0; k < NRA; ++k) tmp += A[k * NCA + row] * B[k * NCB + col];
      const int index = (row + start_row_c) * col_stride_c + start_col_c + col;
      if (kOp > 0) C[index] += tmp;
      else if (kOp < 0) C[index] -= tmp;
      else C[index] = tmp;
    }
  }
}

// c op= A * b                                 (small_blas.h:265-331)
template <int kRowA, int kColA, int kOp>
inline void MatrixVectorMultiply(const double* A, int num_row_a, int num_col_a,
                                 const double* b, double* c) {
  const int NRA = ORC_DIM(kRowA, num_row_a);
  const int NCA = ORC_DIM(kColA, num_col_a);
  for (int row = 0; row < NRA; ++row) {
    double tmp = 0.0;
    for (int col = 0; col < NCA; ++col) tmp += A[row * NCA + col] * b[col];
    if (kOp > 0) c[row] += tmp;
    else if (kOp < 0) c[row] -= tmp;
    else c[row] = tmp;
  }
}

// c op= A' * b                                (small_blas.h:333-400)
template <int kRowA, int kColA, int kOp>
inline void MatrixTransposeVectorMultiply(const double* A, int num_row_a, int num_col_a,
                                          const double* b, double* c) {
  const int NRA = ORC_DIM(kRowA, num_row_a);
  const int NCA = ORC_DIM(kColA, num_col_a);
  for (int row = 0; row < NCA; ++row) {
    double tmp = 0.0;
    for (int col = 0; col < NRA; ++col) tmp += A[col * NCA + row] * b[col];
    if (kOp > 0) c[row] += tmp;
    else if (kOp < 0) c[row] -= tmp;
    else c[row] = tmp;
  }
}

// Dense SPD helpers standing in for the Eigen calls on this path (Eigen is a
// third-party dependency that is not vendored under /root/reference):
//   m.inverse() for fixed 3x3   invert_psd_matrix.h:62-64
//   m.selfadjointView<Upper>().llt().solve(I)   implicit_schur_complement.cc:201,
//                                               block_random_access_diagonal_matrix.cc:97
// Both are textbook algorithms; results agree with Eigen to rounding.

// Cholesky (upper triangle read) solve of M X = I, n <= 16; returns false if not SPD.
inline bool InvertSymmetricUpperLLT(int n, double* m /* n*n row-major, in/out */) {
  double L[16 * 16];
  if (n > 16) return false;
  for (int j = 0; j < n; ++j) {
    double d = m[j * n + j];  // upper == lower by symmetry; read upper: m[k][j], k<=j
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    if (!(d > 0.0)) return false;
    const double ljj = __builtin_sqrt(d);
    L[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double s = m[j * n + i];  // upper triangle entry (j,i), j<i
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / ljj;
    }
  }
  // Solve L L' X = I column by column.
  for (int c = 0; c < n; ++c) {
    double y[16];
    for (int i = 0; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
      y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * m[k * n + c];
      m[i * n + c] = s / L[i * n + i];
    }
  }
  return true;
}

// Cofactor inverse of a 3x3 (what Eigen's fixed-size inverse() evaluates).
inline void Invert3x3(const double* m, double* inv) {
  const double c00 = m[4] * m[8] - m[5] * m[7];
  const double c01 = m[5] * m[6] - m[3] * m[8];
  const double c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const double id = 1.0 / det;
  inv[0] = c00 * id;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  inv[3] = c01 * id;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  inv[6] = c02 * id;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

}  // namespace orc
