"""world_size-2 gloo test (CPU) of the N>1 host logic: env/key sharding is a
partition of the single-process job, timings reduce with MAX, sweep shards gather."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rsr_mjx_b200 import airbot_spec as A, prng, sharding


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 6
        keys = sharding.shard_keys(0, n, rank, world)
        m = A.load_model("sf")
        qpos, qvel, ctrl = A.sample_reset(m, "sf", keys)
        tmax = sharding.reduce_max([1.0 + rank, 5.0 - rank])
        gathered = sharding.gather_concat(torch.arange(3, dtype=torch.float32) + 10 * rank)
        # PPO / SAC gradient pmean (RSR/train.py:261-262) and the normaliser's cross-rank statistics (:333-336)
        from rsr_mjx_b200 import ppo
        torch.manual_seed(0)
        net = ppo.MLP([3, 4, 2])
        extra = torch.nn.Parameter(torch.zeros(()))
        params = list(net.parameters()) + [extra]
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        ppo._flat_allreduce_mean(params)
        norm = ppo.RunningStatistics(2, "cpu")
        norm.update(torch.tensor([[1.0, 2.0], [3.0, 6.0]]) + 10.0 * rank)
        q.put((rank, keys, qpos, tmax, gathered.numpy(), list(sharding.env_range(n, rank)),
               [float(p.grad.flatten()[0]) for p in params], norm.mean.numpy(), float(norm.count), norm.std.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_env_sharding_world2():
    world, port = 2, 29517 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    all_keys = prng.split(prng.PRNGKey(0), 12)
    np.testing.assert_array_equal(np.concatenate([r[1] for r in res]), all_keys)
    m = A.load_model("sf")
    q_all, _, _ = A.sample_reset(m, "sf", all_keys)
    np.testing.assert_array_equal(np.concatenate([r[2] for r in res]), q_all)  # same envs as one process
    assert res[0][5] + res[1][5] == list(range(12))
    for r in res:
        assert r[3] == [2.0, 5.0]  # MAX over ranks, identical everywhere
        np.testing.assert_array_equal(r[4], [0, 1, 2, 10, 11, 12])
        assert r[6] == [1.5 * (i + 1) for i in range(5)]  # mean over ranks of (rank + 1) * (i + 1), every parameter
        # statistics of the union of both ranks' batches
        allx = np.array([[1.0, 2.0], [3.0, 6.0], [11.0, 12.0], [13.0, 16.0]])
        np.testing.assert_allclose(r[7], allx.mean(0), rtol=1e-6)
        assert r[8] == 4.0
        np.testing.assert_allclose(r[9], allx.std(0), rtol=1e-5)


def test_single_process_degenerates():
    assert sharding.reduce_max([3.0, 1.0]) == [3.0, 1.0]
    t = torch.arange(4.0)
    assert sharding.gather_concat(t) is t
    np.testing.assert_array_equal(sharding.shard_keys(3, 5, 0, 1), prng.split(prng.PRNGKey(3), 5))
