"""Per-phase warp-cycle breakdown of the step kernel in steady state (uses rs_world_step_profile)."""
import ctypes, sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights
from roboschool_b200 import _lib

env_id = sys.argv[1] if len(sys.argv) > 1 else "RoboschoolHumanoid-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 300
m = load_env_model(env_id)
W = BatchedWorld(m, n, 0)
random_actions = "random" in sys.argv[4:]      # uniform random actions from the device RNG instead of the zoo policy
pol = None if random_actions else ZooPolicy(load_zoo_weights(env_id), 0)
obs_p, act_p, rew_p, done_p = W.rollout_buffers()
st = torch.cuda.current_stream().cuda_stream
W.rollout_reset(seed=0, stream=st)
W.rollout(pol, warm, stream=st)
torch.cuda.synchronize()
if len(sys.argv) > 4 and "stagger" in sys.argv[4:]:
    W.set_episode(frame=np.random.RandomState(7).randint(0, m.max_episode_steps - 1, size=n).astype(np.int32))
    W.rollout(pol, 100, stream=st)
    torch.cuda.synchronize()
names = ["kinematics", "collide", "aba", "solve", "integrate", "epilogue+store", "load+action", "total",
         "solve:enumerate", "solve:limit_tasks", "solve:contact_tasks", "solve:rhs", "solve:pgs"]
acc = np.zeros(16)
for _ in range(5):
    if pol is not None:
        pol.forward(obs_p, act_p, n, st)
    cyc = (ctypes.c_ulonglong * 16)()
    _lib.check(W.L.rs_world_step_profile(W.h, act_p if pol is not None else None, obs_p, rew_p, done_p, cyc, st))
    acc += np.array(list(cyc), dtype=np.float64)
r, c = W.counts()
out = {"env": env_id, "n": n, "actions": "random" if random_actions else "policy", "mean_rows": float(r.mean()), "max_rows": int(r.max()), "mean_contacts": float(c.mean()),
       "share": {k: float(v / acc[7]) for k, v in zip(names, acc)}, "warp_cycles_per_env_step": float(acc[7] / 5 / (n / 32) )}
print(json.dumps(out))
