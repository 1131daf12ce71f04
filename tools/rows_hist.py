"""Constraint rows per env in steady state: distribution, and what the slowest lane costs its warp (sizing of the PGS phase)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights

env_id = sys.argv[1] if len(sys.argv) > 1 else "RoboschoolHumanoidFlagrunHarder-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
m = load_env_model(env_id)
W = BatchedWorld(m, n, 0)
pol = ZooPolicy(load_zoo_weights(env_id), 0)
W.rollout_reset(seed=0)
W.rollout(pol, 300)
torch.cuda.synchronize()
if len(sys.argv) > 3 and sys.argv[3] == "stagger":      # episode clocks spread over the time limit, like bench.py
    W.set_episode(frame=np.random.RandomState(7).randint(0, m.max_episode_steps - 1, size=n).astype(np.int32))
    W.rollout(pol, 100)
    torch.cuda.synchronize()
acc = []
for _ in range(10):
    W.rollout(pol, 7)
    torch.cuda.synchronize()
    r, c = W.counts()
    acc.append(r.copy())
r = np.concatenate(acc)
w = r.reshape(-1, 32)
out = {"env": env_id, "n": n, "mean_rows": float(r.mean()), "p50": float(np.percentile(r, 50)), "p90": float(np.percentile(r, 90)),
       "p99": float(np.percentile(r, 99)), "max": int(r.max()), "hist": np.bincount(np.minimum(r, 99) // 6).tolist(),
       "warp_max_mean": float(w.max(1).mean()), "warp_mean_mean": float(w.mean(1).mean()),
       "warp_2nd_max_mean": float(np.sort(w, 1)[:, -2].mean()), "warp_sum_over12_mean": float((w * (w > 12)).sum(1).mean()),
       "warp_max_le12_mean": float(np.where(w > 12, 0, w).max(1).mean()), "frac_envs_over12": float((r > 12).mean())}
print(json.dumps(out))
