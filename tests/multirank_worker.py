"""Worker of tests/test_gpu_multirank.py (one process per GPU under torch.distributed.run): the sharded CUDA path
(points sharded by observation count, cameras replicated) against the unsharded single-GPU result and
against the CPU oracle.  Rank 0 prints "MULTIRANK-OK" when every comparison holds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def main():
    import torch
    import torch.distributed as dist
    import ceres_solver_b200 as cs
    from ceres_solver_b200 import bal as B

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    workload = sys.argv[1] if len(sys.argv) > 1 else "trafalgar-257"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(cs.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().numpy().tobytes())

    bal = B.synthetic(workload)
    rp = B.ReducedProgram(bal)
    full = rp.state(bal)
    plo, phi, rlo, rhi = rp.shard(rank, world)
    gpu = cs.Problem(rp.C, phi - plo, rp.row_cam[rlo:rhi], rp.row_pt[rlo:rhi] - plo, rp.row_obs[rlo:rhi], device=local,
                     rank=rank, world_size=world, nccl_id=nccl_id)
    state = np.concatenate([full[3 * plo:3 * phi], full[3 * rp.P:]])
    nP = 3 * (phi - plo)

    # every rank takes part in the sharded calls (they contain collectives), in the same order
    ok, cost, res, grad = gpu.evaluate(state)
    sq = gpu.squared_column_norm()
    s = 1.0 / (1.0 + np.sqrt(sq))
    gpu.scale_columns(s)
    D = np.sqrt(np.clip(gpu.squared_column_norm(), 1e-6, 1e32) / 1e4)
    gpu.schur_init(res, D)
    rhs = gpu.schur_rhs()
    u = np.random.RandomState(7).randn(9 * rp.C)
    Su = gpu.schur_multiply(u)
    blocks, _ = gpu.schur_jacobi_update()
    xfull = np.random.RandomState(8).randn(3 * rp.P + 9 * rp.C)   # same vector on every rank; each takes its slice
    xj = np.concatenate([xfull[3 * plo:3 * phi], xfull[3 * rp.P:]])
    jtj = gpu.jtj_multiply(xj, D)
    _, recs = gpu.lm_solve(state, gpu.lm_options(max_num_iterations=3))
    _, recs_hb = gpu.lm_solve(state, gpu.lm_options(max_num_iterations=3), host_boundary=True)   # through the host buffers
    gpu.close()

    failures = []
    if rank == 0:
        def check(name, got, want, tol):
            e = relerr(got, want)
            if not e < tol:
                failures.append("%s: %.3e >= %.1e" % (name, e, tol))

        one = cs.Problem(rp.C, rp.P, rp.row_cam, rp.row_pt, rp.row_obs, device=local)
        ok1, cost1, res1, grad1 = one.evaluate(full)
        if not (ok and ok1):
            failures.append("evaluate failed")
        check("cost", [cost], [cost1], 1e-12)
        check("residuals(shard)", res, res1[2 * rlo:2 * rhi], 1e-13)
        check("gradient cameras", grad[nP:], grad1[3 * rp.P:], 1e-11)
        check("gradient points(shard)", grad[:nP], grad1[3 * plo:3 * phi], 1e-12)
        sq1 = one.squared_column_norm()
        check("column norms cameras", sq[nP:], sq1[3 * rp.P:], 1e-12)
        s1 = 1.0 / (1.0 + np.sqrt(sq1))
        one.scale_columns(s1)
        D1 = np.sqrt(np.clip(one.squared_column_norm(), 1e-6, 1e32) / 1e4)
        check("D cameras", D[nP:], D1[3 * rp.P:], 1e-12)
        one.schur_init(res1, D1)
        check("schur rhs", rhs, one.schur_rhs(), 1e-10)
        check("S u", Su, one.schur_multiply(u), 1e-10)
        b1, _ = one.schur_jacobi_update()
        check("schur jacobi blocks", blocks, b1, 1e-10)
        full_jtj = one.jtj_multiply(xfull, D1)
        check("J'J x points(shard)", jtj[:nP], full_jtj[3 * plo:3 * phi], 1e-11)
        check("J'J x cameras", jtj[nP:], full_jtj[3 * rp.P:], 1e-11)
        _, recs1 = one.lm_solve(full, one.lm_options(max_num_iterations=3))
        one.close()
        from oracle import pyoracle as po
        orc = po.BaProgram(bal.C, bal.P, bal.cam_idx, bal.pt_idx, np.ascontiguousarray(bal.obs).ravel())
        o = orc.default_options()
        o.num_threads = po.max_threads()
        o.max_num_iterations = 3
        _, recs_o, _ = orc.solve(full, o)
        for a, b in zip(recs_hb, recs):
            for key in ("cost", "step_norm", "gradient_max_norm", "tr_radius"):
                if abs(a[key] - b[key]) > 1e-6 * max(abs(b[key]), 1e-300):
                    failures.append("host-boundary iteration %d %s: %r vs %r" % (a["iteration"], key, a[key], b[key]))
            if a["ls_iterations"] != b["ls_iterations"]:
                failures.append("host-boundary iteration %d CG iterations %d vs %d" % (a["iteration"], a["ls_iterations"], b["ls_iterations"]))
        if len(recs_hb) != len(recs):
            failures.append("host-boundary trace length %d vs %d" % (len(recs_hb), len(recs)))
        if not (len(recs) == len(recs1) == len(recs_o)):
            failures.append("trace lengths %d %d %d" % (len(recs), len(recs1), len(recs_o)))
        else:
            for a, b, c in zip(recs, recs1, recs_o):
                for key, tol in (("cost", 1e-6), ("step_norm", 1e-6), ("gradient_max_norm", 1e-5), ("tr_radius", 1e-5)):
                    for other, who in ((b[key], "unsharded"), (float(c[key]), "oracle")):
                        if abs(a[key] - other) > tol * max(abs(other), 1e-300):
                            failures.append("iteration %d %s vs %s: %r vs %r" % (a["iteration"], key, who, a[key], other))
                if a["ls_iterations"] != b["ls_iterations"] or a["ls_iterations"] != int(c["ls_iterations"]):
                    failures.append("iteration %d CG iterations %d / %d / %d" % (a["iteration"], a["ls_iterations"],
                                                                                b["ls_iterations"], int(c["ls_iterations"])))
        print("MULTIRANK-OK" if not failures else "MULTIRANK-FAIL\n" + "\n".join(failures), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
