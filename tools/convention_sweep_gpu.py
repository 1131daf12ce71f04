"""GPU leg of the convention sweep (tools/convention_sweep.py is the CPU-oracle leg with the oracle-only knobs).

Evaluates a GRID of recalled importer / solver conventions on the CUDA path, 2048 envs x one full episode per
(env, policy generation), scored like the CPU leg: mean first-episode return / registered reward_threshold
(roboschool/__init__.py:11-77) for every pretrained policy of the reference at once.

  here      : python tools/convention_sweep_gpu.py prepare      # MJCF variants -> tools/sweep_models/ (needs /root/reference)
  GPU box   : python tools/convention_sweep_gpu.py run [n_envs] # -> gpurun_out/r02/sweep_gpu.txt
"""
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from roboschool_b200 import envspec                      # noqa: E402
from roboschool_b200.model import ModelDesc              # noqa: E402
from roboschool_b200.policy import load_zoo_weights      # noqa: E402

ASSETS = "/root/reference/roboschool/mujoco_assets"
OUT = os.path.join(ROOT, "tools", "sweep_models")
ENVS = ["RoboschoolHopper-v1", "RoboschoolWalker2d-v1", "RoboschoolHalfCheetah-v1", "RoboschoolAnt-v1", "RoboschoolHumanoid-v1"]
GENS = ["v0", "v1"]

MJCF_GRID = {           # compile_mjcf keyword -> alternatives (first = shipped convention)
    "honour_stiffness": [False, True],
    "honour_armature": [False, True],
    "com_rule": ["last_fromto", "centroid"],
    "aabb_margin": [0.04, 0.001],
}
DESC_GRID = {           # descriptor field -> alternatives
    "erp": [0.2, 0.5],
    "split_thresh": [-0.04, -1e30],
    "lin_ang_damping": [0.04, 0.0],
}


def variants(grid):
    keys = list(grid)
    for combo in itertools.product(*[grid[k] for k in keys]):
        yield dict(zip(keys, combo))


def vname(v):
    return "_".join("%s-%s" % (k[:8], str(x)) for k, x in v.items())


def prepare():
    os.makedirs(OUT, exist_ok=True)
    for v in variants(MJCF_GRID):
        d = os.path.join(OUT, vname(v))
        os.makedirs(d, exist_ok=True)
        for env_id in ENVS:
            envspec.compile_env(env_id, ASSETS, **v).save(os.path.join(d, env_id + ".json"))
    print("wrote", len(list(variants(MJCF_GRID))), "variants to", OUT)


def run(n=2048, T=1000):
    import torch
    from roboschool_b200.world import BatchedWorld
    from roboschool_b200.policy import ZooPolicy
    dev = torch.device("cuda:0")
    pols = {}
    for env_id in ENVS:
        for gen in GENS:
            w = [a.copy() for a in load_zoo_weights(env_id, gen)]
            if gen == "v0" and "Humanoid" in env_id:    # agent_zoo/RoboschoolHumanoid_v0_2017may.py:15 shifts obs[0]
                w[1] = w[1] + w[0][0] * (-1.4 + 0.8)
            pols[(env_id, gen)] = ZooPolicy(w, 0)
    header = " ".join("%-10s" % (e.replace("Roboschool", "")[:6] + "-" + g) for e in ENVS for g in GENS)
    print("%-72s %s" % ("setting (first value of every knob = shipped)", header), flush=True)
    rows = []
    bufs = {}
    for mv in variants(MJCF_GRID):
        d = os.path.join(OUT, vname(mv))
        for dv in variants(DESC_GRID):
            cells, vals = [], []
            t0 = time.time()
            for env_id in ENVS:
                m = ModelDesc.load(os.path.join(d, env_id + ".json"))
                m.erp = dv["erp"]; m.split_thresh = dv["split_thresh"]
                m.c_erp, m.c_split_thresh = m.erp, m.split_thresh
                m.lin_damping = m.ang_damping = dv["lin_ang_damping"]
                thr = m.meta["reward_threshold"]
                world = BatchedWorld(m, n, 0)
                key = (m.obs_dim, m.n_act)
                if key not in bufs:
                    bufs[key] = (torch.empty((T + 1, n, m.obs_dim), device=dev), torch.empty((T, n, m.n_act), device=dev),
                                 torch.empty((T, n), device=dev), torch.empty((T, n), dtype=torch.uint8, device=dev))
                obs_T, act_T, rew_T, done_T = bufs[key]
                for gen in GENS:
                    world.rollout_reset(seed=1, env_id0=0)
                    world.rollout_record(pols[(env_id, gen)], T, obs_T, act_T, rew_T, done_T)
                    torch.cuda.synchronize()
                    done = done_T.bool()
                    first = torch.where(done.any(0), done.float().argmax(0), torch.full((n,), T - 1, device=dev))
                    live = (torch.arange(T, device=dev)[:, None] <= first[None, :]).float()
                    ret = (rew_T * live).sum(0)
                    r = float(ret.mean()) / thr
                    ln = float((first + 1).float().mean())
                    cells.append("%5.2f/%4d" % (r, ln)); vals.append(r)
                world.close()
            name = " ".join("%s=%s" % (k[7:11] if k.startswith("honour_") else k[:5], x) for k, x in {**mv, **dv}.items())
            line = "%-72s %s | min %5.2f mean %5.2f (%.0fs)" % (name[:72], " ".join(cells), min(vals), float(np.mean(vals)), time.time() - t0)
            print(line, flush=True)
            rows.append((min(vals), float(np.mean(vals)), line))
    print("\nbest 10 by min ratio:")
    for r in sorted(rows, key=lambda r: -r[0])[:10]:
        print(r[2])
    print("\nbest 10 by mean ratio:")
    for r in sorted(rows, key=lambda r: -r[1])[:10]:
        print(r[2])


if __name__ == "__main__":
    if sys.argv[1] == "prepare":
        prepare()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 2048)
