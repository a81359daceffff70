"""Small permuted-order problem through evaluate / S*x (for compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ceres_solver_b200 as cs
from ceres_solver_b200 import bal as B
from oracle import pyoracle as po
kind = sys.argv[1] if len(sys.argv) > 1 else "circle"
bal = B.synthetic_bal(400, 12000, 52000, seed=11) if kind == "circle" else B.synthetic("tiny")
rp = B.ReducedProgram(bal)
orc = po.BaProgram(bal.C, bal.P, bal.cam_idx, bal.pt_idx, np.ascontiguousarray(bal.obs).ravel())
gpu = cs.Problem(rp.C, rp.P, rp.row_cam, rp.row_pt, rp.row_obs)
state = rp.state(bal)
ok, cost, res, grad = gpu.evaluate(state)
ok_o, cost_o, res_o, grad_o = orc.evaluate(state, nt=8)
print("cost", cost, cost_o)
print("res err", np.linalg.norm(res - res_o) / np.linalg.norm(res_o), "grad err", np.linalg.norm(grad - grad_o) / np.linalg.norm(grad_o))
J = orc.jacobian()
s = 1.0 / (1.0 + np.sqrt(J.squared_column_norm()))
gpu.scale_columns(s); J.scale_columns(s, nt=8)
D = np.sqrt(np.clip(J.squared_column_norm(), 1e-6, 1e32) / 1e4)
isc = po.ImplicitSchur(J, gpu.P, want_ftf=False, nt=8)
isc.init(D, res_o)
gpu.schur_init(res_o, D)
print("rhs err", np.linalg.norm(gpu.schur_rhs() - isc.rhs()) / np.linalg.norm(isc.rhs()))
u = np.random.RandomState(1).randn(9 * gpu.C)
print("Su err", np.linalg.norm(gpu.schur_multiply(u) - isc.right_multiply(u)) / np.linalg.norm(isc.right_multiply(u)))
x, its, term = gpu.schur_solve(res_o, D, gpu.solver_options(q_tolerance=1e-2, r_tolerance=-1.0))
xo, its_o, term_o = J.linear_solve(gpu.P, res_o, D, solver=0, q_tolerance=1e-2, r_tolerance=-1.0, nt=8)
print("solve", its, term, its_o, term_o, np.linalg.norm(x - xo) / np.linalg.norm(xo))
