"""Step-kernel time against the number of resident CTAs (7-warp CTAs, one per SM), steady-state batch.

Per-CTA work is the same at every size (one CTA = 224 envs on one SM), so the kernel time should be flat in the number of CTAs up to
148 -- unless something chip-wide runs out.  The per-thread local frame (world transforms, contact caches, row scalars, spills) times
the number of threads crosses the 126 MB L2 between 74 and 148 CTAs; this tool measures where the time rises.
Usage: RS_WPC=7 python tools/l2_cliff.py [env_id]     (prints one JSON line per size)"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("RS_WPC", "7")
import numpy as np
import torch
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights
from roboschool_b200 import _lib

env_id = sys.argv[1] if len(sys.argv) > 1 else "RoboschoolHumanoid-v1"
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [37, 74, 92, 110, 128, 148]
m = load_env_model(env_id)
pol = ZooPolicy(load_zoo_weights(env_id), 0)
st = torch.cuda.current_stream().cuda_stream
for ctas in sizes:
    n = min(ctas * 224, 32768 if ctas >= 147 else ctas * 224)
    W = BatchedWorld(m, n, 0)
    obs_p, act_p, rew_p, done_p = W.rollout_buffers()
    W.rollout_reset(seed=0, stream=st)
    W.rollout(pol, 300, stream=st)
    torch.cuda.synchronize()
    W.set_episode(frame=np.random.RandomState(7).randint(0, m.max_episode_steps - 1, size=n).astype(np.int32))
    W.rollout(pol, 100, stream=st)
    torch.cuda.synchronize()
    T = 100
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
    for t in range(T):
        pol.forward(obs_p, act_p, n, st)
        ev[t][0].record()
        _lib.check(W.L.rs_world_step(W.h, act_p, obs_p, rew_p, done_p, 1, st))
        ev[t][1].record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    print(json.dumps({"env": env_id, "ctas": (n + 223) // 224, "envs": n, "k_step_ms": ms, "env_steps_per_s": n / ms * 1e3,
                      "local_frame_mb": n * 5056 / 1e6}), flush=True)
    del W
