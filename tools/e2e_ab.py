"""End-to-end host-buffer step (rs_world_policy_step_host): graph replay vs direct launches, same settled world.

Checks that both forms give bit-identical observations / rewards / done flags over STEPS steps from the same state, then times
each (wall clock around STEPS dependent steps: the output buffer of one step is the input of the next).
Usage: python tools/e2e_ab.py [env_id] [n_envs] [steps]      (prints one JSON line)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from roboschool_b200.envspec import load_env_model
from roboschool_b200.world import BatchedWorld
from roboschool_b200.policy import ZooPolicy, load_zoo_weights

env_id = sys.argv[1] if len(sys.argv) > 1 else "RoboschoolHumanoid-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
m = load_env_model(env_id)
pol = ZooPolicy(load_zoo_weights(env_id), 0)
st = torch.cuda.current_stream().cuda_stream


def settled(graph):
    os.environ["RS_E2E_GRAPH"] = "1" if graph else "0"
    W = BatchedWorld(m, n, 0)
    W.rollout_reset(seed=0, stream=st)
    W.rollout(pol, 300, stream=st)
    torch.cuda.synchronize()
    W.set_episode(frame=np.random.RandomState(7).randint(0, m.max_episode_steps - 1, size=n).astype(np.int32))
    bufs = [W.pinned_host_buffers(), W.pinned_host_buffers()]
    o, r, d = W.policy_step_host(pol, None, out=bufs[0])
    return W, bufs, o


def run(W, bufs, o, k, graph, record=None):
    os.environ["RS_E2E_GRAPH"] = "1" if graph else "0"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(k):
        o, r, d = W.policy_step_host(pol, o, out=bufs[(i + 1) & 1])
        if record is not None:
            record.append((o.copy(), r.copy(), d.copy()))
    torch.cuda.synchronize()
    return o, time.perf_counter() - t0


Wg, bg, og = settled(True)
Wd, bd, od = settled(False)
assert (og == od).all()
rg, rd = [], []
og, _ = run(Wg, bg, og, 12, True, rg)
od, _ = run(Wd, bd, od, 12, False, rd)
same = all((a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all() for a, b in zip(rg, rd))
out = {"env": env_id, "envs": n, "steps": steps, "bit_identical_12_steps": bool(same)}
for rep in range(2):
    og, tg = run(Wg, bg, og, steps, True)
    od, td = run(Wd, bd, od, steps, False)
    out["graph_env_steps_per_s_%d" % rep] = n * steps / tg
    out["direct_env_steps_per_s_%d" % rep] = n * steps / td
    out["graph_ms_per_step_%d" % rep] = tg / steps * 1e3
    out["direct_ms_per_step_%d" % rep] = td / steps * 1e3
print(json.dumps(out))
