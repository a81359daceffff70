"""world_size-2 and -4 (gloo, CPU) check of the multi-GPU decomposition: points are sharded by observation
count, cameras replicated, and every camera-sized quantity of the Schur path is the SUM over shards — so a single
all-reduce of the 9C vector per CG iteration reproduces the unsharded product.  The per-shard arithmetic here is the
CPU oracle's; the same partition function (ReducedProgram.shard) and the same id/all-reduce plumbing drive the GPU
ranks in bench.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ceres_solver_b200 import bal as B
    from oracle import pyoracle as po

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # a 128-byte id travels from rank 0 to everybody exactly as the NCCL unique id does in bench.py
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(idt, 0)
    assert bytes(idt.numpy().tobytes()) == bytes(range(128))

    bal = B.synthetic("tiny")
    rp = B.ReducedProgram(bal)
    state = rp.state(bal)
    # full problem (reference result, computed redundantly on every rank)
    full = po.BaProgram(bal.C, bal.P, bal.cam_idx, bal.pt_idx, bal.obs.ravel())
    ok, cost, res, grad = full.evaluate(state)
    rng = np.random.RandomState(7)
    D = rng.rand(full.num_parameters) + 0.5
    x = rng.randn(9 * rp.C)
    isc = po.ImplicitSchur(full.jacobian(), rp.P)
    isc.init(D, res)
    Sx_full, rhs_full = isc.right_multiply(x), isc.rhs()

    # this rank's shard: a contiguous point range with all its rows, every camera
    plo, phi, rlo, rhi = rp.shard(rank, world)
    # cameras keep their global ids: give the oracle the shard's rows in input order with untouched camera ids
    obs_rows = rp.obs_of_row[rlo:rhi]
    order = np.argsort(obs_rows)  # back to input order so the oracle rebuilds the same row order
    shard = po.BaProgram(bal.C, bal.P, bal.cam_idx[obs_rows][order], bal.pt_idx[obs_rows][order],
                         bal.obs[obs_rows][order].ravel())
    # the shard program drops unused cameras/points: map its blocks back to the global ids
    cam_map = rp.camera_of_fblock.tolist()
    g_of_f = np.array([cam_map.index(c) for c in shard.camera_of_fblock])
    assert shard.P == phi - plo
    st = np.concatenate([state[3 * plo:3 * phi], state[3 * rp.P:].reshape(-1, 9)[g_of_f].ravel()])
    ok_s, cost_s, res_s, grad_s = shard.evaluate(st)
    assert np.allclose(res_s, res[2 * rlo:2 * rhi], rtol=1e-13, atol=1e-13)
    Dl = np.concatenate([D[3 * plo:3 * phi], np.zeros(9 * shard.C)])   # D_f^2 x is added once, after the reduction
    isc_s = po.ImplicitSchur(shard.jacobian(), shard.P)
    isc_s.init(Dl, res_s)
    xs = x.reshape(-1, 9)[g_of_f].ravel()
    part = np.zeros(9 * rp.C)
    part.reshape(-1, 9)[g_of_f] = isc_s.right_multiply(xs).reshape(-1, 9)
    rhs_part = np.zeros(9 * rp.C)
    rhs_part.reshape(-1, 9)[g_of_f] = isc_s.rhs().reshape(-1, 9)
    t = torch.from_numpy(np.stack([part, rhs_part]))
    dist.all_reduce(t)  # the one collective of a CG iteration
    Sx = t[0].numpy() + D[3 * rp.P:] ** 2 * x
    c = torch.tensor([cost_s], dtype=torch.float64)
    dist.all_reduce(c)
    np.save(os.path.join(out_dir, "r%d.npy" % rank),
            np.array([np.linalg.norm(Sx - Sx_full) / np.linalg.norm(Sx_full),
                      np.linalg.norm(t[1].numpy() - rhs_full) / np.linalg.norm(rhs_full),
                      abs(float(c[0]) - cost) / cost]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_schur_product_is_a_sum_over_shards(world, tmp_path, oracle):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err = np.load(tmp_path / ("r%d.npy" % r))
        assert err[0] < 1e-12 and err[1] < 1e-12 and err[2] < 1e-13, err
