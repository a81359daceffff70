"""Pins the CPU oracle against the reference's own known-answer fixtures.

Mirrors the reference tests that hold these vectors:
  schur_eliminator_test.cc:82-177        eliminator vs dense R - Q'PQ           (1e-14 rel)
  implicit_schur_complement_test.cc:119-216  implicit S columns / rhs / back-substitution vs eliminator
  partitioned_matrix_view_test.cc:103-272    E/F products and block diagonals vs dense
  iterative_schur_complement_solver_test.cc:76-149, schur_complement_solver_test.cc:151-320
                                          solver solution vs DENSE_QR (numpy lstsq here)
  block_sparse_matrix_test.cc:191-674    SpMV / SquaredColumnNorm / ScaleColumns
plus the printed S, r, S\\r, A\\b, x, x_D of linear_least_squares_problems.cc.
"""
import numpy as np
import pytest

from tests.golden.llsq_fixtures import FIXTURES, dense, num_cols_e

IDS = [0, 1, 2, 3, 4, 5, 6]


def make(oracle, fx):
    return oracle.BlockSparseMatrix(fx["col_sizes"], fx["row_sizes"], fx["row_cells"], fx["values"])


def elim_blocks(pid, fx):
    return 1 if pid == 0 else fx["num_eliminate_blocks"]


def dense_schur(A, b, D, ne):
    """H = A'A + D^2; S = R - Q' P^-1 Q; rhs = g_f - Q' P^-1 g_e  (schur_eliminator_test.cc:82-121)."""
    H = A.T @ A + (np.diag(D * D) if D is not None else 0.0)
    g = A.T @ b
    P, Q, R = H[:ne, :ne], H[:ne, ne:], H[ne:, ne:]
    Pinv = np.linalg.inv(P)
    S = R - Q.T @ Pinv @ Q
    rhs = g[ne:] - Q.T @ Pinv @ g[:ne]
    return S, rhs, H, g


@pytest.mark.parametrize("pid", IDS)
def test_transcription_matches_printed_normal_matrix(pid):
    fx = FIXTURES[pid]
    if "AtA" not in fx:
        pytest.skip("no printed A'A for this fixture")
    A = dense(fx)
    assert np.allclose(A.T @ A, np.array(fx["AtA"], dtype=float), atol=1e-12)


@pytest.mark.parametrize("pid", IDS)
@pytest.mark.parametrize("nt", [1, 4])
def test_block_sparse_products(oracle, pid, nt):
    fx = FIXTURES[pid]
    A = dense(fx)
    M = make(oracle, fx)
    rng = np.random.RandomState(pid)
    x = rng.randn(A.shape[1])
    y = rng.randn(A.shape[0])
    assert np.allclose(M.right_multiply(x, nt), A @ x, rtol=0, atol=1e-12)
    assert np.allclose(M.left_multiply(y, nt), A.T @ y, rtol=0, atol=1e-12)
    assert np.allclose(M.squared_column_norm(nt), (A * A).sum(axis=0), rtol=1e-14, atol=0)
    s = rng.rand(A.shape[1]) + 0.5
    M.scale_columns(s, nt)
    assert np.allclose(M.right_multiply(x, nt), (A * s) @ x, rtol=0, atol=1e-12)


@pytest.mark.parametrize("pid", IDS)
@pytest.mark.parametrize("nt", [1, 2, 4, 8])
def test_partitioned_view(oracle, pid, nt):
    fx = FIXTURES[pid]
    ne_blocks = elim_blocks(pid, fx)
    A = dense(fx)
    ne = int(sum(fx["col_sizes"][:ne_blocks]))
    E, F = A[:, :ne], A[:, ne:]
    M = make(oracle, fx)
    rng = np.random.RandomState(10 + pid)
    xe, xf, y = rng.randn(ne), rng.randn(A.shape[1] - ne), rng.randn(A.shape[0])
    assert np.allclose(M.pmv(ne_blocks, 0, xe, A.shape[0], nt), E @ xe, atol=1e-12)
    assert np.allclose(M.pmv(ne_blocks, 1, xf, A.shape[0], nt), F @ xf, atol=1e-12)
    assert np.allclose(M.pmv(ne_blocks, 2, y, ne, nt), E.T @ y, atol=1e-12)
    assert np.allclose(M.pmv(ne_blocks, 3, y, A.shape[1] - ne, nt), F.T @ y, atol=1e-12)
    # block diagonals
    ete = M.block_diagonal(ne_blocks, 0, nt)
    p = 0
    c0 = 0
    for s in fx["col_sizes"][:ne_blocks]:
        blk = (E.T @ E)[c0:c0 + s, c0:c0 + s]
        assert np.allclose(ete[p:p + s * s].reshape(s, s), blk, atol=1e-12)
        p += s * s
        c0 += s
    ftf = M.block_diagonal(ne_blocks, 1, nt)
    p = 0
    c0 = 0
    for s in fx["col_sizes"][ne_blocks:]:
        blk = (F.T @ F)[c0:c0 + s, c0:c0 + s]
        assert np.allclose(ftf[p:p + s * s].reshape(s, s), blk, atol=1e-12)
        p += s * s
        c0 += s


@pytest.mark.parametrize("pid", [0, 2, 4, 5, 6])
@pytest.mark.parametrize("use_D", [False, True])
@pytest.mark.parametrize("nt", [1, 4])
@pytest.mark.parametrize("force_dynamic", [0, 1])
def test_schur_eliminator_vs_dense(oracle, pid, use_D, nt, force_dynamic):
    fx = FIXTURES[pid]
    if pid in (4, 6) and not use_D:
        pytest.skip("rank deficient without the diagonal (reference note :545-547)")
    ne_blocks = elim_blocks(pid, fx)
    A = dense(fx)
    ne = int(sum(fx["col_sizes"][:ne_blocks]))
    b = np.array(fx["b"], dtype=float)
    D = np.array(fx["D"], dtype=float) if use_D else None
    S, rhs, H, g = dense_schur(A, b, D, ne)
    M = make(oracle, fx)
    lhs, r = M.schur_eliminate(ne_blocks, b, D, nt=nt, force_dynamic=force_dynamic)
    # the eliminator fills the upper block triangle only
    lhs_full = np.triu(lhs) + np.triu(lhs, 1).T
    fsizes = fx["col_sizes"][ne_blocks:]
    # inside diagonal blocks both triangles are written; check them unsymmetrised too
    c0 = 0
    for s in fsizes:
        assert np.allclose(lhs[c0:c0 + s, c0:c0 + s], S[c0:c0 + s, c0:c0 + s], rtol=1e-13, atol=1e-13 * abs(S).max())
        c0 += s
    assert np.linalg.norm(lhs_full - S) / np.linalg.norm(S) < 1e-14 * 10
    assert np.linalg.norm(r - rhs) / np.linalg.norm(rhs) < 1e-13
    # back substitution: full solution of the regularised normal equations
    z = np.linalg.solve(S, rhs)
    sol = M.schur_back_substitute(ne_blocks, b, D, z, nt=nt, force_dynamic=force_dynamic)
    expect = np.linalg.solve(H, g)
    assert np.allclose(sol[:ne], expect[:ne], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("pid", [2, 5])
def test_printed_schur_complement(oracle, pid):
    """KAT-1 / KAT-2: printed S, r, S\\r, A\\b (4 decimals) for D = 0."""
    fx = FIXTURES[pid]
    M = make(oracle, fx)
    b = np.array(fx["b"], dtype=float)
    lhs, r = M.schur_eliminate(fx["num_eliminate_blocks"], b, None)
    S = np.triu(lhs) + np.triu(lhs, 1).T
    assert np.allclose(S, np.array(fx["S"]), atol=6e-5)
    assert np.allclose(r, np.array(fx["r"]), atol=6e-5)
    z = np.linalg.solve(S, r)
    assert np.allclose(z, np.array(fx["S_solve_r"]), atol=6e-5)
    x, its, term = M.linear_solve(fx["num_eliminate_blocks"], b, None, solver=1)
    assert term == 0
    assert np.allclose(x, np.array(fx["x"]), atol=1.1e-4)


def test_problem0_known_solutions(oracle):
    """KAT-3: x = [2,3]; with D = [1,2]: x_D = [1.78448275, 2.82327586]."""
    fx = FIXTURES[0]
    M = make(oracle, fx)
    b = np.array(fx["b"], dtype=float)
    x, _, term = M.linear_solve(1, b, None, solver=1)
    assert term == 0 and np.allclose(x, fx["x"], atol=1e-12)
    xd, _, term = M.linear_solve(1, b, np.array(fx["D"], dtype=float), solver=1)
    assert term == 0 and np.allclose(xd, fx["x_D"], atol=5e-9)
    xi, its, term = M.linear_solve(1, b, np.array(fx["D"], dtype=float), solver=0, r_tolerance=1e-12, max_iter=50)
    assert term == 0 and np.allclose(xi, fx["x_D"], atol=5e-9)


@pytest.mark.parametrize("pid", [2, 4, 5, 6])
@pytest.mark.parametrize("use_D", [False, True])
@pytest.mark.parametrize("force_dynamic", [0, 1])
def test_implicit_schur_complement(oracle, pid, use_D, force_dynamic):
    """implicit_schur_complement_test.cc:119-216: columns of S, rhs and back-substitution (1e-14 abs there)."""
    fx = FIXTURES[pid]
    if pid in (4, 6) and not use_D:
        pytest.skip("rank deficient without the diagonal")
    ne_blocks = fx["num_eliminate_blocks"]
    A = dense(fx)
    ne = num_cols_e(fx)
    b = np.array(fx["b"], dtype=float)
    D = np.array(fx["D"], dtype=float) if use_D else None
    S, rhs, H, g = dense_schur(A, b, D, ne)
    M = make(oracle, fx)
    isc = oracle.ImplicitSchur(M, ne_blocks, want_ftf=True, force_dynamic=force_dynamic)
    isc.init(D, b)
    nf = A.shape[1] - ne
    scale = abs(S).max()
    for i in range(nf):
        e = np.zeros(nf)
        e[i] = 1.0
        assert np.allclose(isc.right_multiply(e), S[:, i], rtol=0, atol=1e-13 * scale)
    assert np.allclose(isc.rhs(), rhs, rtol=0, atol=1e-13 * max(1.0, abs(rhs).max()))
    z = np.linalg.solve(S, rhs)
    sol = isc.back_substitute(z)
    expect = np.linalg.solve(H, g)
    assert np.allclose(sol, expect, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("pid", [2, 3, 5])
@pytest.mark.parametrize("precond", [0, 1, 2])
@pytest.mark.parametrize("use_D", [False, True])
def test_iterative_schur_vs_dense_qr(oracle, pid, precond, use_D):
    """iterative_schur_complement_solver_test.cc:76-115 (r_tolerance 1e-12, compare with DENSE_QR at 1e-14..)."""
    fx = FIXTURES[pid]
    A = dense(fx)
    b = np.array(fx["b"], dtype=float)
    D = np.array(fx["D"], dtype=float) if use_D else None
    Aaug = np.vstack([A, np.diag(D)]) if use_D else A
    baug = np.concatenate([b, np.zeros(A.shape[1])]) if use_D else b
    expect = np.linalg.lstsq(Aaug, baug, rcond=None)[0]
    M = make(oracle, fx)
    x, its, term = M.linear_solve(fx["num_eliminate_blocks"], b, D, solver=0, preconditioner=precond,
                                  r_tolerance=1e-12, max_iter=100)
    assert term == 0
    assert np.allclose(x, expect, rtol=0, atol=1e-10)


def _spse_setup(oracle, pid=5):
    """power_series_expansion_preconditioner_test.cc:44-78: fixture 5 with its D, dense S^-1 as the reference."""
    fx = FIXTURES[pid]
    A = dense(fx)
    ne = num_cols_e(fx)
    b = np.array(fx["b"], dtype=float)
    D = np.array(fx["D"], dtype=float)
    S, rhs, H, g = dense_schur(A, b, D, ne)
    M = make(oracle, fx)
    isc = oracle.ImplicitSchur(M, fx["num_eliminate_blocks"], want_ftf=True)
    isc.init(D, b)
    return fx, A, ne, D, S, isc, M, b


@pytest.mark.parametrize("tolerance", [1e-14, 0.0])
def test_power_series_expansion_is_the_schur_inverse(oracle, tolerance):
    """power_series_expansion_preconditioner_test.cc:80-124: 50 terms (tolerance 1e-14 reached, or tolerance 0 and all
    50 terms) reproduce every column of S^-1 to 1e-14."""
    fx, A, ne, D, S, isc, M, b = _spse_setup(oracle)
    Sinv = np.linalg.inv(S)
    nf = S.shape[0]
    for i in range(nf):
        e = np.zeros(nf)
        e[i] = 1.0
        y = isc.power_series(e, max_num_spse_iterations=50, spse_tolerance=tolerance)
        assert np.linalg.norm(y - Sinv[:, i]) < 1e-14


def test_power_series_expansion_bad_tolerance_stops_early(oracle):
    """:126-147: with tolerance 1/eps the series stops after one term and is NOT the inverse."""
    fx, A, ne, D, S, isc, M, b = _spse_setup(oracle)
    Sinv = np.linalg.inv(S)
    nf = S.shape[0]
    for i in range(nf):
        e = np.zeros(nf)
        e[i] = 1.0
        y = isc.power_series(e, max_num_spse_iterations=50, spse_tolerance=1e14)
        assert np.linalg.norm(y - Sinv[:, i]) > 1e-14


@pytest.mark.parametrize("terms", [1, 2, 5])
def test_power_series_expansion_partial_sums(oracle, terms):
    """The k-term value against the dense definition: sum_{i<=k} (M^-1 F'E P E'F)^i M^-1 x, M = F'F + D_f^2."""
    fx, A, ne, D, S, isc, M_, b = _spse_setup(oracle)
    E, F = A[:, :ne], A[:, ne:]
    P = np.linalg.inv(E.T @ E + np.diag(D[:ne] ** 2))
    Mi = np.linalg.inv(F.T @ F + np.diag(D[ne:] ** 2))
    T = Mi @ F.T @ E @ P @ E.T @ F
    rng = np.random.RandomState(7)
    x = rng.randn(S.shape[0])
    expect = Mi @ x
    term = expect.copy()
    for _ in range(terms):
        term = T @ term
        expect = expect + term
    got = isc.power_series(x, max_num_spse_iterations=terms, spse_tolerance=0.0)
    assert np.allclose(got, expect, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("pid", [2, 3, 5])
@pytest.mark.parametrize("spse_init", [False, True])
def test_iterative_schur_power_series_vs_dense_qr(oracle, pid, spse_init):
    """SCHUR_POWER_SERIES_EXPANSION as the preconditioner (iterative_schur_complement_solver.cc:178-186) and as the
    initial guess (:100-111) both reach the DENSE_QR solution."""
    fx = FIXTURES[pid]
    A = dense(fx)
    b = np.array(fx["b"], dtype=float)
    D = np.array(fx["D"], dtype=float)
    expect = np.linalg.lstsq(np.vstack([A, np.diag(D)]), np.concatenate([b, np.zeros(A.shape[1])]), rcond=None)[0]
    M = make(oracle, fx)
    x, its, term = M.linear_solve(fx["num_eliminate_blocks"], b, D, solver=0, preconditioner=3 if not spse_init else 2,
                                  r_tolerance=1e-12, max_iter=100, use_spse_initialization=spse_init)
    assert term == 0
    assert np.allclose(x, expect, rtol=0, atol=1e-10)


@pytest.mark.parametrize("pid", [2, 4, 5, 6])
def test_dense_schur_vs_dense_qr(oracle, pid):
    """schur_complement_solver_test.cc:127-131: |x - x_qr| / n < 1e-10 with the regulariser on."""
    fx = FIXTURES[pid]
    A = dense(fx)
    b = np.array(fx["b"], dtype=float)
    D = np.array(fx["D"], dtype=float)
    expect = np.linalg.lstsq(np.vstack([A, np.diag(D)]), np.concatenate([b, np.zeros(A.shape[1])]), rcond=None)[0]
    M = make(oracle, fx)
    x, _, term = M.linear_solve(fx["num_eliminate_blocks"], b, D, solver=1)
    assert term == 0
    assert np.linalg.norm(x - expect) / A.shape[1] < 1e-10


def test_schur_jacobi_diagonal_matches_full(oracle):
    """SCHUR_JACOBI asks the eliminator for diagonal cells only (schur_jacobi_preconditioner.cc:87-97):
    they must equal the diagonal blocks of the full S."""
    fx = FIXTURES[6]
    M = make(oracle, fx)
    D = np.array(fx["D"], dtype=float)
    ne_blocks = fx["num_eliminate_blocks"]
    lhs, _ = M.schur_eliminate(ne_blocks, None, D)
    fsizes = fx["col_sizes"][ne_blocks:]
    diag, _ = M.schur_eliminate(ne_blocks, None, D, diagonal_only=True, diag_len=sum(s * s for s in fsizes))
    p = 0
    c0 = 0
    for s in fsizes:
        assert np.allclose(diag[p:p + s * s].reshape(s, s), lhs[c0:c0 + s, c0:c0 + s], rtol=1e-14)
        p += s * s
        c0 += s


# ---------------------------------------------------------------------------------------------------------------------
# conjugate_gradients_solver_test.cc, loss_function_test.cc (HuberLoss), corrector_test.cc
# ---------------------------------------------------------------------------------------------------------------------
def test_cg_solves_3x3_identity_system(oracle):
    """conjugate_gradients_solver_test.cc:47-87: A = I, b = (1, 2, 3), x0 = (1, 1, 1): one iteration, exact."""
    x, its, term = oracle.cg_dense(np.eye(3), [1.0, 2.0, 3.0], [1.0, 1.0, 1.0], min_iter=1, max_iter=10,
                                   reset_period=20, q_tolerance=0.0, r_tolerance=1e-9)
    assert (term, its) == (0, 1)
    assert np.array_equal(x, [1.0, 2.0, 3.0])


def test_cg_solves_3x3_symmetric_system(oracle):
    """:89-147: the tridiagonal [2 -1 0; -1 2 -1; 0 -1 2], b = (-1, 0, 3), x0 = (1, 1, 1) -> (0, 1, 2)."""
    A = np.array([[2.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 2.0]])
    x, its, term = oracle.cg_dense(A, [-1.0, 0.0, 3.0], [1.0, 1.0, 1.0], min_iter=1, max_iter=10, reset_period=20,
                                   q_tolerance=0.0, r_tolerance=1e-9)
    assert term == 0
    assert np.allclose(x, [0.0, 1.0, 2.0], rtol=0, atol=1e-9)


@pytest.mark.parametrize("a", [0.7, 1.3])
@pytest.mark.parametrize("s", [0.357, 1.792])
def test_huber_loss_derivatives(oracle, a, s):
    """loss_function_test.cc:45-102: rho' and rho'' against symmetric finite differences of rho (h = 1e-4, 1e-6)."""
    h = 1e-4
    rho, fwd, bwd = oracle.huber_loss(a, s), oracle.huber_loss(a, s + h), oracle.huber_loss(a, s - h)
    assert abs((fwd[0] - bwd[0]) / (2 * h) - rho[1]) < 1e-6
    assert abs((fwd[0] - 2 * rho[0] + bwd[0]) / (h * h) - rho[2]) < 1e-6
    assert np.allclose(oracle.huber_loss(0.7, 0.0), [0.0, 1.0, 0.0], atol=1e-6)


@pytest.mark.parametrize("residual,rho", [(np.sqrt(3.0), [3.0, 0.1, -0.01]), (0.0, [0.0, 0.1, -0.01]),
                                          (np.sqrt(3.0), [3.0, 0.1, -0.1])])
def test_corrector_scalar_cases(oracle, residual, rho):
    """corrector_test.cc:56-138: rho'' < 0 (or a zero residual) clamps alpha to 0: both scale by sqrt(rho')."""
    r, J = oracle.corrector(residual * residual, rho, [residual], [10.0])
    assert abs(r[0] - residual * np.sqrt(rho[1])) < 1e-6
    assert abs(J[0] - np.sqrt(rho[1]) * 10.0) < 1e-6


def test_corrector_multidimensional_gauss_newton(oracle):
    """corrector_test.cc:140-205: corrected residuals / Jacobian against the closed forms and the robustified gradient."""
    rng = np.random.RandomState(0)
    for _ in range(2000):
        jac = rng.uniform(0.0, 1.0, (3, 2))
        res = rng.uniform(0.0, 1.0, 3)
        sq = float(res @ res)
        rho = [sq, rng.uniform(0.0, 1.0), rng.uniform(-1.0, 1.0)]
        kD = 1 + 2 * rho[2] / rho[1] * sq
        alpha = 1 - np.sqrt(kD) if rho[2] > 0.0 else 0.0
        g_res = np.sqrt(rho[1]) / (1.0 - alpha) * res
        g_jac = np.sqrt(rho[1]) * (jac - alpha / sq * np.outer(res, res) @ jac)
        g_grad = rho[1] * jac.T @ res
        r, J = oracle.corrector(sq, rho, res, jac)
        J = J.reshape(3, 2)
        assert np.linalg.norm(g_res - r) < 1e-10
        assert np.linalg.norm(g_jac - J) < 1e-10
        assert np.linalg.norm(g_grad - J.T @ r) < 1e-10


def test_angle_axis_rotate_point_matches_rotation_matrix(oracle):
    """rotation_test.cc:1809-1857 and :1873-1921 (tolerance 10 eps there; 1e-14 here, the comparison goes through scipy's
    rotation matrix): AngleAxisRotatePoint against R(angle_axis) p for angles across (-pi, pi) and for the near-zero
    branch (|theta| <= 1e-16, first-order Taylor expansion)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.RandomState(3)
    thetas = np.concatenate([(2.0 * np.arange(0, 10000, 37) * 0.0011 - 1.0) * np.pi,
                             (2.0 * np.arange(0, 10000, 97) * 0.0001 - 1.0) * 1e-16])
    for theta in thetas:
        axis = rng.uniform(-1.0, 1.0, 3)
        p = rng.uniform(-1.0, 1.0, 3)
        aa = axis * (theta / np.linalg.norm(axis))
        expect = Rotation.from_rotvec(aa).as_matrix() @ p
        got = oracle.angle_axis_rotate_point(aa, p)
        assert np.abs(got - expect).max() < 1e-14
