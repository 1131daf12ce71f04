"""Compile the reference's MJCF robots into flat descriptors (roboschool_b200/models/*.json).

Usage: python tools/compile_models.py [/root/reference/roboschool/mujoco_assets]
The GPU box has no /root/reference; these JSON descriptors are what travels.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from roboschool_b200 import envspec

assets = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/roboschool/mujoco_assets"
os.makedirs(envspec.MODEL_DIR, exist_ok=True)
for env_id in envspec.ENV_SPECS:
    m = envspec.compile_env(env_id, assets)
    m.save(os.path.join(envspec.MODEL_DIR, env_id + ".json"))
    tot = m.base_mass + sum(m.mass[i] for i in range(m.n_links))
    print("%-32s links=%2d dof=%2d geoms=%2d pairs=%3d fixed_base=%d mass=%.2f parts=%d" % (
        env_id, m.n_links, m.n_dof, m.n_geoms, m.n_pairs, m.fixed_base, tot, m.n_parts))
