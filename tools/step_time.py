"""Steady-state step-kernel time for (env, n_envs) pairs with the library's default launch shapes.
Usage: python tools/step_time.py RoboschoolHopper-v1:4096 RoboschoolAnt-v1:65536 ...   (one JSON line each)"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights
from roboschool_b200 import _lib

st = torch.cuda.current_stream().cuda_stream
# STEP_TIME_STATS=1: step through rs_world_step_stats (the entry point bench.py times) instead of rs_world_step
stats = torch.zeros(8, dtype=torch.float64, device="cuda") if os.environ.get("STEP_TIME_STATS") else None
for arg in sys.argv[1:]:
    env_id, n = arg.split(":")
    n = int(n)
    m = load_env_model(env_id)
    pol = ZooPolicy(load_zoo_weights(env_id), 0)
    W = BatchedWorld(m, n, 0)
    obs_p, act_p, rew_p, done_p = W.rollout_buffers()
    W.rollout_reset(seed=0, stream=st)
    W.rollout(pol, 300, stream=st)
    torch.cuda.synchronize()
    W.set_episode(frame=np.random.RandomState(7).randint(0, m.max_episode_steps - 1, size=n).astype(np.int32))
    W.rollout(pol, int(os.environ.get("STEP_TIME_WARM", "1000")), stream=st)     # steps between staggering the episode clocks and timing
    torch.cuda.synchronize()
    T = 100
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
    for t in range(T):
        pol.forward(obs_p, act_p, n, st)
        ev[t][0].record()
        if stats is not None:
            _lib.check(W.L.rs_world_step_stats(W.h, act_p, obs_p, rew_p, done_p, 1, stats.data_ptr(), st))
        else:
            _lib.check(W.L.rs_world_step(W.h, act_p, obs_p, rew_p, done_p, 1, st))
        ev[t][1].record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in ev]
    ms = float(np.median(ts))
    r, c = W.counts()
    print(json.dumps({"env": env_id, "envs": n, "k_step_ms": ms, "env_steps_per_s": n / ms * 1e3, "k_step_ms_mean": float(np.mean(ts)), "k_step_ms_max": float(np.max(ts)), "stats": stats is not None, "first_steps_ms": [round(x, 3) for x in ts[:24]], "mean_rows": float(r.mean()),
                      "variant": W.L.rs_world_kernel_variant(W.h)}), flush=True)
    del W
