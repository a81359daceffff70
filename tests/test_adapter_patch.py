"""The Ceres-side patch of the drop-in must apply to the reference tree, and after it the virtuals B200Jacobian
overrides must no longer be `final` (round-1 verdict: the adapter could not compile as patched).  Needs the reference tree,
which exists in the development container only (skipped elsewhere)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = ["internal/ceres/block_sparse_matrix.h", "internal/ceres/evaluator.cc", "internal/ceres/linear_solver.cc",
         "internal/ceres/trust_region_minimizer.cc"]

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "internal", "ceres")) or shutil.which("patch") is None,
                                reason="needs the reference tree and patch(1)")


def test_patch_is_current_and_applies(tmp_path):
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_adapter_patch.py"), REF], capture_output=True, text=True)
    assert gen.returncode == 0, gen.stderr
    with open(os.path.join(ROOT, "adapter", "ceres_b200.patch")) as f:
        assert f.read() == gen.stdout, "adapter/ceres_b200.patch is stale: regenerate with tools/make_adapter_patch.py"
    for rel in FILES:
        dst = tmp_path / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, rel), dst)
    r = subprocess.run(["patch", "-p1", "-i", os.path.join(ROOT, "adapter", "ceres_b200.patch")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    patched = (tmp_path / FILES[0]).read_text()
    assert "class CERES_NO_EXPORT BlockSparseMatrix : public SparseMatrix" in patched
    # every virtual the adapter overrides is overridable after the patch
    header = open(os.path.join(ROOT, "adapter", "b200_adapter.h")).read()
    jac = header[header.index("class B200Jacobian"):header.index("class B200Evaluator")]
    names = set(re.findall(r"void (\w+)\(", jac)) - {"SyncValuesToHost"}
    assert names == {"SquaredColumnNorm", "ScaleColumns", "RightMultiplyAndAccumulate", "LeftMultiplyAndAccumulate", "SetZero"}
    iface = patched[patched.index("// Implementation of SparseMatrix interface."):patched.index("// Convert to CompressedRowSparseMatrix")]
    for n in names:
        decls = re.findall(r"void %s\([^;]*;" % n, iface, flags=re.S)
        assert decls and all("final" not in d for d in decls), (n, decls)
    for rel, needle in ((FILES[1], "B200Evaluator::Create(options, program, error)"), (FILES[2], "B200IterativeSchurSolver"),
                        (FILES[3], "ModelCostChange")):
        assert needle in (tmp_path / rel).read_text()
