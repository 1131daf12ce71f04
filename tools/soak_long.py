"""Long soak of the benchmark configuration: 32768 Humanoid envs under the zoo policy for 100 000 env steps (3.3 G env-steps,
~100 episodes per env), statistics accumulated on the device; then 20 000 steps of FlagrunHarder."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, numpy as np
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights
for env_id, n, T in [("RoboschoolHumanoid-v1", 32768, 100000), ("RoboschoolHumanoidFlagrunHarder-v1", 32768, 20000)]:
    m = load_env_model(env_id); W = BatchedWorld(m, n, 0); pol = ZooPolicy(load_zoo_weights(env_id), 0)
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    W.rollout_reset(seed=21, stream=st)
    t0 = time.time()
    for chunk in range(T // 10000):
        W.rollout(pol, 10000, seed=21, stats=stats, stream=st)
        torch.cuda.synchronize()
        s = stats.cpu().numpy()
        print(env_id, "steps %d: reward/step %.3f episodes %d mean len %.1f nonfinite %d  (%.1f M env-steps/s wall)" % (
            (chunk + 1) * 10000, s[0] / s[6], s[1], s[2] / max(s[1], 1), s[3], s[6] / (time.time() - t0) / 1e6), flush=True)
    state = W.get_state()
    print(env_id, "final state finite", bool(np.isfinite(state).all()), "max|state| %.1f" % float(np.abs(state).max()), flush=True)
