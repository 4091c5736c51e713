"""Multi-GPU host logic on CPU (world_size 2, gloo): batch sharding, shard-local workload generation, result gather
and the whole-job throughput reduction of bench.py (no collective in the solve itself).  The per-rank
"solve" is the CPU oracle here -- the GPU box runs the same host logic around the CUDA kernel (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

from ik_b200 import sharding
from ik_b200 import workloads as W


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            rs = [sharding.shard_range(B, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = sharding.shard_sizes(B, world)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_workload_is_shard_reproducible():
    """Counter-based RNG keyed by (seed, b): generating a shard on its own equals slicing the full batch."""
    m = W.cassie_model()
    full = W.sample_configurations(m, 1000, seed=5)
    for world in (2, 8):
        parts = [W.sample_configurations(m, hi - lo, seed=5, b0=lo)
                 for lo, hi in (sharding.shard_range(1000, r, world) for r in range(world))]
        assert np.array_equal(np.concatenate(parts), full)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    import torch
    import torch.distributed as dist

    from oracle import oracle as O
    from oracle.bridge import make_workload, oracle_model, oracle_problem_like

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pb = W.cassie_feet_pelvis_problem()
        om = oracle_model("cassie")
        opb = oracle_problem_like(pb, om)
        lo, hi = sharding.shard_range(B, rank, world)
        q0, tg, _ = make_workload(pb, om, hi - lo, seed=77, standing=W.CASSIE_STANDING, b0=lo)
        q, ok, it, res = O.dls_batch(opb, q0, tg)
        local = {"q": torch.tensor(q.T.copy()), "success": torch.tensor(ok), "iters": torch.tensor(it),
                 "resid": torch.tensor(res)}
        allr = sharding.gather_results(local, B, dist)
        thr, t = sharding.reduce_throughput(int(ok.sum()), 0.5 + rank, dist)  # rank 1 is the slow one: 1.5 s
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), q=allr["q"].numpy().T, success=allr["success"].numpy(),
                     iters=allr["iters"].numpy(), resid=allr["resid"].numpy(), thr=thr, t=t)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle as O
    from oracle.bridge import make_workload, oracle_model, oracle_problem_like

    B, world = 301, 2  # odd: shards of 150 and 151 exercise the padding in gather_results
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    q0, tg, _ = make_workload(pb, om, B, seed=77, standing=W.CASSIE_STANDING)
    q, ok, it, res = O.dls_batch(oracle_problem_like(pb, om), q0, tg)
    assert np.array_equal(got["q"], q) and np.array_equal(got["success"], ok)
    assert np.array_equal(got["iters"], it) and np.array_equal(got["resid"], res)
    # whole-job throughput = converged problems of all ranks / MAX over ranks of the time
    assert abs(float(got["t"]) - 1.5) < 1e-12 and abs(float(got["thr"]) - ok.sum() / 1.5) < 1e-9
