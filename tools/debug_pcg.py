"""Runs schur_solve on C16 (optionally a synthetic workload) and prints the summary; used to A/B the persistent PCG kernel
against the default multi-kernel path (B200_PCG_PERSISTENT=1 selects the persistent kernel).   python tools/debug_pcg.py [workload] [max_it] [precond]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ceres_solver_b200 as cs
from ceres_solver_b200 import bal as B

workload = sys.argv[1] if len(sys.argv) > 1 else "c16"
max_it = int(sys.argv[2]) if len(sys.argv) > 2 else 500
precond = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bal = B.normalize(B.read_bal(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "problem-16-22106-pre.txt.bz2"))) if workload == "c16" else B.synthetic(workload)
rp = B.ReducedProgram(bal)
gpu = cs.Problem(rp.C, rp.P, rp.row_cam, rp.row_pt, rp.row_obs)
state = rp.state(bal)
ok, cost, res, grad = gpu.evaluate(state)
s = 1.0 / (1.0 + np.sqrt(gpu.squared_column_norm()))
gpu.scale_columns(s)
D = np.sqrt(np.clip(gpu.squared_column_norm(), 1e-6, 1e32) / 1e4)
reps = int(os.environ.get('REPS', '1'))
for q_tol, r_tol in ((1e-2, -1.0), (0.0, 1e-10)) * reps:
    o = gpu.solver_options(preconditioner_type=precond, q_tolerance=q_tol, r_tolerance=r_tol, max_num_iterations=max_it)
    t0 = time.time()
    x, its, term = gpu.schur_solve(res, D, o)
    print("q_tol %g r_tol %g: its %d term %d |x| %.12e  (%.1f ms)" % (q_tol, r_tol, its, term, np.linalg.norm(x), 1e3 * (time.time() - t0)), flush=True)
