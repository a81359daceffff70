"""tests/c_abi_smoke.c -- a C program that binds include/b200ba.h directly -- is built with gcc against libb200ba.so (CPU
part: it compiles and links, i.e. the header is valid C and every function it uses is exported) and, on a GPU box, run on
a BAL-shaped problem and compared with the oracle (create -> evaluate -> schur_solve -> LM through host buffers -> destroy)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import ceres_solver_b200 as cs
    cs.lib()
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.join(ROOT, "ceres_solver_b200")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c"),
           "-o", exe, "-L", libdir, "-l:libb200ba.so", "-lm", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _write_problem(path, rp, state):
    with open(path, "wb") as f:
        f.write(np.int32(rp.C).tobytes())
        f.write(np.int32(rp.P).tobytes())
        f.write(np.int64(rp.N).tobytes())
        f.write(np.ascontiguousarray(rp.row_cam, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(rp.row_pt, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(rp.row_obs, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(state, dtype=np.float64).tobytes())


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_caller_builds_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the gpu-marked test runs the program")
    from ceres_solver_b200 import bal as B
    bal = B.synthetic("tiny")
    rp = B.ReducedProgram(bal)
    _write_problem(tmp_path / "p.bin", rp, rp.state(bal))
    r = subprocess.run([exe, str(tmp_path / "p.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr   # B200_ERR_NO_DEVICE, never a silent CPU path


@pytest.mark.gpu
def test_c_caller_matches_oracle(tmp_path, oracle):
    exe = _build(tmp_path)
    from ceres_solver_b200 import bal as B
    bal = B.synthetic_bal(64, 4000, 18000, seed=3)
    rp = B.ReducedProgram(bal)
    state = rp.state(bal)
    _write_problem(tmp_path / "p.bin", rp, state)
    r = subprocess.run([exe, str(tmp_path / "p.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("done"), (r.stdout, r.stderr)
    out = {}
    lm = []
    for line in r.stdout.splitlines():
        w = line.split()
        if w[0] == "cost":
            out["cost"] = float(w[1])
        elif w[0] == "gradient_max_norm":
            out["gmax"] = float(w[1])
        elif w[0] == "solve":
            out["its"], out["term"], out["xnorm"] = int(w[2]), int(w[4]), float(w[6])
        elif w[0] == "lm":
            lm.append((int(w[1]), float(w[3]), float(w[5]), int(w[7])))
    orc = oracle.BaProgram(bal.C, bal.P, bal.cam_idx, bal.pt_idx, np.ascontiguousarray(bal.obs).ravel())
    ok, cost_o, res_o, grad_o = orc.evaluate(state, nt=8)
    assert abs(out["cost"] - cost_o) <= 1e-12 * cost_o
    assert abs(out["gmax"] - np.abs(grad_o).max()) <= 1e-10 * np.abs(grad_o).max()
    J = orc.jacobian()
    D = np.sqrt(np.clip(J.squared_column_norm(), 1e-6, 1e32) / 1e4)
    xo, its_o, term_o = J.linear_solve(rp.P, res_o, D, solver=0, q_tolerance=1e-2, r_tolerance=-1.0, nt=8)
    assert (out["its"], out["term"]) == (its_o, term_o)
    assert abs(out["xnorm"] - np.linalg.norm(xo)) <= 1e-7 * np.linalg.norm(xo)
    o = orc.default_options()
    o.num_threads = 8
    o.max_num_iterations = 2
    _, recs_o, _ = orc.solve(state, o)
    assert len(lm) == len(recs_o)
    for (it, cost, sn, cg), b in zip(lm, recs_o):
        assert cg == int(b["ls_iterations"])
        assert abs(cost - b["cost"]) <= 1e-6 * abs(b["cost"])
        assert abs(sn - b["step_norm"]) <= 1e-6 * max(abs(b["step_norm"]), 1e-30)
