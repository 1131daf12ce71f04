"""Development check of the tcgen05 policy kernel against numpy float64 (run on the GPU box, under `timeout`)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from roboschool_b200.policy import ZooPolicy, load_zoo_weights
from oracle.oracle import mlp_forward
for env_id, O, A in [("RoboschoolHumanoid-v1", 44, 17), ("RoboschoolHopper-v1", 15, 3), ("RoboschoolAnt-v1", 28, 8)]:
    wts = load_zoo_weights(env_id)
    pol = ZooPolicy(wts, 0)
    for n in (1, 100, 128, 1000, 32768):
        X = np.random.RandomState(n).uniform(-3, 3, (n, O)).astype(np.float32)
        x = torch.tensor(X, device="cuda"); act = torch.full((n, A), float("nan"), device="cuda")
        pol.forward(x, act); torch.cuda.synchronize()
        ref = mlp_forward(wts, X.astype(np.float64)); got = act.cpu().numpy()
        err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
        print(env_id, "n", n, "rel err", float(err), "finite", bool(np.isfinite(got).all()), flush=True)
    x = torch.rand((32768, O), device="cuda"); act = torch.empty((32768, A), device="cuda")
    for _ in range(10): pol.forward(x, act)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): pol.forward(x, act)
    e1.record(); torch.cuda.synchronize()
    print(env_id, "us per forward (32768 envs):", e0.elapsed_time(e1) * 10, flush=True)
