"""Joint sweep of the recalled Bullet importer / solver conventions, scored by every zoo policy at once.

The physics of the path lives in an un-vendored Bullet fork: every convention that could not be read is
either descriptor data (mjcf.compile_mjcf keyword arguments, rs_model_desc solver fields) or an oracle knob
(oracle_set_knob).  This tool evaluates a setting on the CPU oracle with ALL pretrained policies of the reference
(agent_zoo May-2017 "v0" and July-2017 "v1" generations x Hopper, Walker2d, HalfCheetah, Ant, Humanoid) and reports,
per (env, generation), return / registered reward_threshold (roboschool/__init__.py:11-77).  The setting to keep is the
single one that maximises the minimum ratio -- not one chosen per env.

Needs /root/reference (MJCF files); runs in the build container only.  Results: profiles/r02_convention_sweep.md.

  python tools/convention_sweep.py scan            # one knob at a time around the current setting
  python tools/convention_sweep.py grid            # full grid over the knobs the scan found influential
  python tools/convention_sweep.py one k=v k=v ... # one setting
"""
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from roboschool_b200 import envspec            # noqa: E402
from roboschool_b200.policy import load_zoo_weights   # noqa: E402
from oracle import oracle as O                 # noqa: E402

ASSETS = "/root/reference/roboschool/mujoco_assets"
ENVS = ["RoboschoolHopper-v1", "RoboschoolWalker2d-v1", "RoboschoolHalfCheetah-v1", "RoboschoolAnt-v1", "RoboschoolHumanoid-v1"]
GENS = ["v0", "v1"]

# knob -> (kind, default); kind: "mjcf" compile_mjcf kwarg, "desc" descriptor field, "knob" oracle_set_knob
KNOBS = {
    "honour_density": ("mjcf", False), "honour_damping": ("mjcf", False), "damping_scale": ("mjcf", 1.0),
    "honour_axisangle": ("mjcf", True), "com_rule": ("mjcf", "last_fromto"), "aabb_margin": ("mjcf", 0.04),
    "honour_armature": ("mjcf", False), "honour_stiffness": ("mjcf", False), "honour_settotalmass": ("mjcf", False),
    "contact_erp": ("mjcf", "split"),
    "erp": ("desc", 0.2), "split_thresh": ("desc", -0.04), "lin_damping": ("desc", 0.04), "ang_damping": ("desc", 0.04),
    "max_coord_vel": ("desc", 100.0), "use_gyro": ("desc", 1.0), "n_iter": ("desc", 5), "floor_friction": ("desc", 0.8),
    "max_limit_impulse": ("desc", 100.0),
    "limit_always": ("knob", 0), "friction_dirs": ("knob", 2), "torque_first_only": ("knob", 0), "warmstart": ("knob", 0.0),
    "damping_mode": ("knob", 0),
}


def build_model(env_id, cfg):
    kw = {k: v for k, v in cfg.items() if KNOBS[k][0] == "mjcf"}
    m = envspec.compile_env(env_id, ASSETS, **kw)
    for k, v in cfg.items():
        if KNOBS[k][0] == "desc":
            setattr(m, k, v)
    if "erp" in cfg or "split_thresh" in cfg:      # contact rows follow the limit-row constants in "split" mode
        if cfg.get("contact_erp", "split") == "split":
            m.c_erp, m.c_split_thresh = m.erp, m.split_thresh
    return m


def rollout(env_id, gen, m, n=8, T=1000, seed=1, threads=8):
    wts = load_zoo_weights(env_id, gen)
    w = O.OracleWorld(m, n)
    w.reset(seed=seed)
    obs = w.obs()[0]
    tot = np.zeros(n); alive = np.ones(n, bool); length = np.zeros(n)
    shift = -1.4 + 0.8 if (gen == "v0" and "Humanoid" in env_id) else 0.0     # agent_zoo/RoboschoolHumanoid_v0_2017may.py:15
    for t in range(T):
        x = obs.copy()
        x[:, 0] += shift
        obs, rew, done = w.step(O.mlp_forward(wts, x), threads=threads)
        tot += rew * alive; length += alive
        alive &= ~done.astype(bool)
        if not alive.any():
            break
    return tot, length


def evaluate(cfg, envs=ENVS, gens=GENS, n=8, T=1000):
    L = O.lib()
    for k, (kind, dflt) in KNOBS.items():
        if kind == "knob":
            L.oracle_set_knob(k.encode(), float(cfg.get(k, dflt)))
    out = {}
    for env_id in envs:
        m = build_model(env_id, cfg)
        thr = m.meta["reward_threshold"]
        for gen in gens:
            tot, length = rollout(env_id, gen, m, n=n, T=T)
            out[(env_id, gen)] = (float(tot.mean()) / thr, float(length.mean()))
    for k, (kind, dflt) in KNOBS.items():
        if kind == "knob":
            L.oracle_set_knob(k.encode(), float(dflt))
    return out


def fmt(out):
    cells = []
    for env_id in ENVS:
        for gen in GENS:
            if (env_id, gen) in out:
                r, ln = out[(env_id, gen)]
                cells.append("%5.2f/%4d" % (r, ln))
    vals = [v[0] for v in out.values()]
    return " ".join(cells) + "  | min %5.2f mean %5.2f" % (min(vals), float(np.mean(vals)))


HEADER = " ".join("%-10s" % (e.replace("Roboschool", "")[:6] + "-" + g) for e in ENVS for g in GENS)

SCAN = {
    "contact_erp": ["erp2", 0.5], "erp": [0.9, 0.05], "split_thresh": [-1e30], "limit_always": [1], "friction_dirs": [1],
    "torque_first_only": [1], "warmstart": [0.85], "lin_damping": [0.0], "ang_damping": [0.0], "use_gyro": [0.0],
    "honour_damping": [True], "honour_armature": [True], "honour_stiffness": [True], "honour_density": [True],
    "honour_axisangle": [False], "com_rule": ["origin", "centroid"], "aabb_margin": [0.0, 0.08],
    "honour_settotalmass": [True], "floor_friction": [0.5, 1.2], "n_iter": [10, 50], "max_limit_impulse": [1000.0],
}


def parse_val(s):
    if s in ("True", "False"):
        return s == "True"
    try:
        return int(s)
    except ValueError:
        try:
            return float(s)
        except ValueError:
            return s


def main():
    oc_knob = O.lib().oracle_set_knob
    import ctypes
    oc_knob.argtypes = [ctypes.c_char_p, ctypes.c_double]
    mode = sys.argv[1] if len(sys.argv) > 1 else "scan"
    base = {}
    args = sys.argv[2:]
    n = 8
    for a in list(args):
        if a.startswith("n="):
            n = int(a[2:]); args.remove(a)
    for a in args:
        if a.startswith("base:"):
            k, v = a[5:].split("=")
            base[k] = parse_val(v)
    args = [a for a in args if not a.startswith("base:")]
    print("%-44s %s" % ("setting", HEADER))
    if mode == "one":
        cfg = dict(base)
        for a in args:
            k, v = a.split("=")
            cfg[k] = parse_val(v)
        t = time.time()
        print("%-44s %s  (%.0fs)" % (json.dumps(cfg)[:44], fmt(evaluate(cfg, n=n)), time.time() - t))
    elif mode == "scan":
        print("%-44s %s" % ("base " + json.dumps(base)[:38], fmt(evaluate(base, n=n))), flush=True)
        for k, vals in SCAN.items():
            for v in vals:
                cfg = dict(base); cfg[k] = v
                print("%-44s %s" % ("%s=%s" % (k, v), fmt(evaluate(cfg, n=n))), flush=True)
    elif mode == "grid":
        # knobs=... as k:v1,v2;k:v1,v2
        spec = {}
        for a in args:
            k, vs = a.split(":")
            spec[k] = [parse_val(v) for v in vs.split(",")]
        keys = list(spec)
        rows = []
        for combo in itertools.product(*[spec[k] for k in keys]):
            cfg = dict(base); cfg.update(dict(zip(keys, combo)))
            out = evaluate(cfg, n=n)
            rows.append((min(v[0] for v in out.values()), cfg, out))
            print("%-44s %s" % (" ".join("%s=%s" % (k[:10], v) for k, v in zip(keys, combo))[:44], fmt(out)), flush=True)
        rows.sort(key=lambda r: -r[0])
        print("\nbest by min ratio:")
        for r in rows[:10]:
            print("%-44s %s" % (json.dumps({k: r[1][k] for k in keys})[:44], fmt(r[2])))


if __name__ == "__main__":
    main()
