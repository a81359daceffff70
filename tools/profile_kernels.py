"""Runs each hot kernel of libb200ba a few times on a synthetic problem (for ncu captures and quick timing).
   python tools/profile_kernels.py [workload] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ceres_solver_b200 as cs
from ceres_solver_b200 import bal as B

workload = sys.argv[1] if len(sys.argv) > 1 else "ladybug-1723"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bal = B.normalize(B.read_bal(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "problem-16-22106-pre.txt.bz2"))) if workload == "c16" else B.synthetic(workload)
rp = B.ReducedProgram(bal)
gpu = cs.Problem(rp.C, rp.P, rp.row_cam, rp.row_pt, rp.row_obs)
state = rp.state(bal)
ok, cost, res, grad = gpu.evaluate(state)
s = 1.0 / (1.0 + np.sqrt(gpu.squared_column_norm()))
gpu.scale_columns(s)
D = np.sqrt(np.clip(gpu.squared_column_norm(), 1e-6, 1e32) / 1e4)
gpu.schur_init(res, D)
rng = np.random.RandomState(0)
x = rng.randn(9 * gpu.C)
xf = rng.randn(gpu.num_parameters)
gpu.stats_reset(); gpu.profile(True)
for _ in range(reps):
    gpu.schur_multiply(x)
    gpu.jtj_multiply(xf, D)
    gpu.schur_jacobi_update()
    gpu.schur_init(res, D)
    gpu.evaluate(state)
    gpu.schur_back_substitute(x)
    gpu.model_cost_change(xf)
st = gpu.stats()
for k, v in st.items():
    if v["launches"]:
        ops = max(v["operations"], 1)
        gb = v["bytes_per_operation"] * ops / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 and v["bytes_per_operation"] > 0 else float("nan")
        print("%-28s launches %4d  operations %4d  mean/op %.4f ms  %8.1f GB/s" % (k, v["launches"], v["operations"], v["ms"] / ops, gb))
