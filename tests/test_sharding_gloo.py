"""N>1 path on CPU: world_size-2 gloo processes, each stepping its env shard (oracle-backed here, CUDA on
the GPU box); the gathered shards must equal the single-process run bit for bit (RNG keyed by global env
id), and the statistics all-reduce must equal the single-process totals."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from roboschool_b200.sharding import shard_range, reduce_stats, summarize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, STEPS, SEED = 12, 15, 4


def test_shard_range_partitions():
    for n, w in [(12, 2), (13, 4), (262144, 8), (5, 8)]:
        blocks = [shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 3, 2)


def _rollout(lo, hi):
    from roboschool_b200.envspec import load_env_model
    from oracle.oracle import OracleWorld
    m = load_env_model("RoboschoolHopper-v1")
    o = OracleWorld(m, hi - lo)
    o.reset(seed=SEED, env_id0=lo)
    acts = np.random.RandomState(0).uniform(-1, 1, (STEPS, N_TOTAL, m.n_act))
    stats = np.zeros(8)
    for t in range(STEPS):
        obs, rew, done = o.step(acts[t, lo:hi], auto_reset=True, seed=SEED, env_id0=lo)
        stats[0] += rew.sum(); stats[1] += done.sum(); stats[6] += hi - lo
    return obs, o.get_state(), stats


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(N_TOTAL, rank, world)
    obs, state, stats = _rollout(lo, hi)
    st = reduce_stats(torch.tensor(stats, dtype=torch.float64))
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, obs, state))
    if rank == 0:
        torch.save({"stats": st, "shards": gathered}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path, oracle_lib):
    out = str(tmp_path / "gather.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    obs1, state1, stats1 = _rollout(0, N_TOTAL)
    obs2 = np.concatenate([s[2] for s in sorted(res["shards"], key=lambda s: s[0])])
    state2 = np.concatenate([s[3] for s in sorted(res["shards"], key=lambda s: s[0])])
    assert np.array_equal(obs1, obs2) and np.array_equal(state1, state2)
    assert np.allclose(res["stats"].numpy(), stats1, rtol=1e-12, atol=1e-12)
    assert summarize(res["stats"])["env_steps"] == N_TOTAL * STEPS
