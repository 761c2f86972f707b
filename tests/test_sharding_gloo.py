"""Multi-GPU path on CPU: world_size-2 gloo processes exercise the scatter -> per-rank solve ->
gather plumbing of bench.py / altro_cpp_b200.sharding with the CPU oracle standing in for the
device solve.  The sharded result must equal the unsharded one bit for bit (per-instance
arithmetic does not depend on placement) — the analogue of the reference's nthreads-equivalence
tests (test/ilqr/ilqr_class_test.cpp:130-160)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from altro_cpp_b200 import problems as P
from altro_cpp_b200.sharding import gather_rows, scatter_rows, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as ob
        spec = P.unicycle_problem(P.K_THREE_OBSTACLES, N=20)
        full = torch.from_numpy(P.perturbed_initial_states(spec, total, P.UNICYCLE_X0_SCALE)) if rank == 0 else None
        x0 = scatter_rows(full, total, (spec.n,), torch.float64, torch.device("cpu"))
        lo, hi = shard_bounds(total, world)[rank]
        assert x0.shape[0] == hi - lo
        res = ob.solve_batch(spec, x0.numpy(), nthreads=1, want_gains=False)
        cost = gather_rows(torch.from_numpy(res["cost"]).reshape(-1, 1), total)
        iters = gather_rows(torch.from_numpy(res["iters"].astype(np.int64)), total)
        X = gather_rows(torch.from_numpy(res["X"]), total)
        if rank == 0:
            np.savez(out_path, cost=cost.numpy()[:, 0], iters=iters.numpy(), X=X.numpy())
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_ragged_batches():
    for total in (1, 2, 7, 16, 65536):
        for world in (1, 2, 4, 8):
            b = shard_bounds(total, world)
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_scatter_solve_gather_equals_single(tmp_path, oracle):
    total = 11  # ragged: 6 + 5
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    got = np.load(out)
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES, N=20)
    X0 = P.perturbed_initial_states(spec, total, P.UNICYCLE_X0_SCALE)
    ref = oracle.solve_batch(spec, X0, nthreads=1, want_gains=False)
    assert np.array_equal(got["cost"], ref["cost"])
    assert np.array_equal(got["iters"], ref["iters"])
    assert np.array_equal(got["X"], ref["X"])
