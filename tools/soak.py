import sys, json
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, numpy as np
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights
for env_id, n in [("RoboschoolHumanoid-v1", 32768), ("RoboschoolAnt-v1", 65536), ("RoboschoolHumanoidFlagrunHarder-v1", 32768), ("RoboschoolHalfCheetah-v1", 16384)]:
    m = load_env_model(env_id); W = BatchedWorld(m, n, 0); pol = ZooPolicy(load_zoo_weights(env_id), 0)
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    W.rollout_reset(seed=3, stream=st)
    W.rollout(pol, 5000, seed=3, stats=stats, stream=st)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    state = W.get_state()
    print(env_id, "mean reward/step %.3f episodes %d mean len %.1f nonfinite %d rows %.2f contacts %.2f state finite %s max|q| %.1f" % (
        s[0]/s[6], s[1], s[2]/max(s[1],1), s[3], s[4]/s[6], s[5]/s[6], bool(np.isfinite(state).all()), float(np.abs(state).max())), flush=True)

# uniform random actions (policy=None: counter-based RNG on the device), every shipped env
from roboschool_b200.envspec import ENV_SPECS
for env_id in sorted(ENV_SPECS):
    n = 8192
    m = load_env_model(env_id); W = BatchedWorld(m, n, 0)
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    W.rollout_reset(seed=4, stream=st)
    W.rollout(None, 3000, seed=4, stats=stats, stream=st)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    state = W.get_state()
    print("random", env_id, "mean reward/step %.3f episodes %d mean len %.1f nonfinite %d rows %.2f contacts %.2f state finite %s max|state| %.1f" % (
        s[0]/s[6], s[1], s[2]/max(s[1],1), s[3], s[4]/s[6], s[5]/s[6], bool(np.isfinite(state).all()), float(np.abs(state).max())), flush=True)
