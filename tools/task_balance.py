"""Row-building tasks per warp and per CTA (7-warp CTAs) in the steady state of FlagrunHarder: what balancing the tasks over the
whole CTA instead of the warp would save.  A contact with a robot side is one task (3 right-hand sides); a contact between two
robot bodies adds a second-body task.  Rounds of the current code: ceil(tasks_w / 32) per warp, the CTA waits for its slowest
warp (solve-phase re-alignment); CTA-wide pool: ceil(sum / 224).
Usage: python tools/task_balance.py [env_id] [n_envs] [sampled CTAs]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights

env_id = sys.argv[1] if len(sys.argv) > 1 else "RoboschoolHumanoidFlagrunHarder-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
n_cta = int(sys.argv[3]) if len(sys.argv) > 3 else 12
m = load_env_model(env_id)
W = BatchedWorld(m, n, 0)
pol = ZooPolicy(load_zoo_weights(env_id), 0)
W.rollout_reset(seed=0)
W.rollout(pol, 300)
torch.cuda.synchronize()
W.set_episode(frame=np.random.RandomState(7).randint(0, m.max_episode_steps - 1, size=n).astype(np.int32))
W.rollout(pol, 1000)
torch.cuda.synchronize()
ng, npair = m.n_geoms, m.n_pairs
first, second, limit_rows = [], [], []
r, c = W.counts()
for cta in np.linspace(0, n // 224 - 1, n_cta).astype(int):
    f = np.zeros(224, dtype=int); s = np.zeros(224, dtype=int)
    for k in range(224):
        keys, _ = W.contacts(int(cta) * 224 + k)
        robot_side = keys != ng + npair                      # everything but cube-floor
        f[k] = int(robot_side.sum())
        s[k] = int(((keys >= ng) & (keys < ng + npair)).sum())
    first.append(f.reshape(7, 32).sum(1)); second.append(s.reshape(7, 32).sum(1))
    rr = r[int(cta) * 224:int(cta) * 224 + 224] - 3 * c[int(cta) * 224:int(cta) * 224 + 224]
    limit_rows.append(rr.reshape(7, 32).sum(1))
first, second, limit_rows = np.array(first), np.array(second), np.array(limit_rows)
ceil = lambda a, b: -(-a // b)
cur = (ceil(first, 32) + ceil(second, 32)).max(1)
pooled = ceil(first.sum(1), 224) + ceil(second.sum(1), 224)
cur_l, pooled_l = ceil(limit_rows, 32).max(1), ceil(limit_rows.sum(1), 224)
print(json.dumps({"env": env_id, "ctas": int(n_cta), "first_tasks_per_warp_mean": float(first.mean()), "first_tasks_per_warp_max": int(first.max()),
                  "second_tasks_per_warp_mean": float(second.mean()), "limit_rows_per_warp_mean": float(limit_rows.mean()),
                  "contact_rounds_now_mean": float(cur.mean()), "contact_rounds_cta_pool_mean": float(pooled.mean()),
                  "limit_rounds_now_mean": float(cur_l.mean()), "limit_rounds_cta_pool_mean": float(pooled_l.mean()),
                  "per_warp_first": first.tolist()[:4], "per_warp_second": second.tolist()[:4]}))
