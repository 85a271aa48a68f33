"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, solve their shard (with the CPU oracle —
this test is about sharding, not about the kernels) and gather; the result must equal the single-process solve."""
import os
import socket

import numpy as np
import pytest

from trajopt_b200 import problems, sharding


def test_shard_bounds_cover_the_batch():
    for total in (0, 1, 5, 8, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch.distributed as dist
    import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = problems.config2(B=5, T=10)  # odd batch: ranks get 3 and 2 trajectories
    mine = sharding.shard(desc, rank, world)
    r = oracle_lib.solve_batch(mine, n_threads=1)
    local = {k: r[k] for k in ("x", "status", "total_cost", "n_qp_solves")}
    full = sharding.gather_results(local, desc.B, dist)
    conv, secs = sharding.reduce_report(int((r["status"] == 0).sum()), 1.0 + rank, dist)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), conv=conv, secs=secs, **full)
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(oracle, tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    desc = problems.config2(B=5, T=10)
    ref = oracle.solve_batch(desc, n_threads=1)
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert (got["status"] == ref["status"]).all() and (got["n_qp_solves"] == ref["n_qp_solves"]).all()
        np.testing.assert_array_equal(got["x"], ref["x"])  # same arithmetic per trajectory, whatever the shard
        np.testing.assert_array_equal(got["total_cost"], ref["total_cost"])
        assert got["conv"] == float((ref["status"] == 0).sum()) and got["secs"] == 2.0
