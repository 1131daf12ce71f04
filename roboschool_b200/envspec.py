"""Per-environment task constants layered on a compiled model.

Restates the class attributes / constructor arguments of
roboschool/gym_mujoco_walkers.py:14-136, gym_forward_walker.py:11-22,83-87,
gym_pendulums.py:70-127 and the gym registration in roboschool/__init__.py:11-110.
"""
import math
import os

import numpy as np

from .model import (ModelDesc, RS_MAX_ACT, RS_MAX_FEET, RS_MAX_PARTS)
from . import mjcf

_HERE = os.path.dirname(os.path.abspath(__file__))
MODEL_DIR = os.path.join(_HERE, "models")

RS_ALIVE_Z_AND_PITCH, RS_ALIVE_Z, RS_ALIVE_ALWAYS, RS_ALIVE_CHEETAH = 0, 1, 2, 3
RS_TASK_WALKER, RS_TASK_PENDULUM, RS_TASK_DOUBLE_PENDULUM, RS_TASK_REACHER = 0, 1, 2, 3

ENV_SPECS = {
    # id: xml, robot_name, A, O, power, foot_list, alive rule/params, per-joint coef overrides
    "RoboschoolHopper-v1": dict(
        xml="hopper.xml", robot="torso", A=3, O=15, power=0.75, feet=["foot"],
        alive_rule=RS_ALIVE_Z_AND_PITCH, alive_z=0.8, alive_bonus=1.0, cost_scale=1.0,
        initial_z=None, coef={}, max_contacts=8, reward_threshold=2500.0),
    "RoboschoolWalker2d-v1": dict(
        xml="walker2d.xml", robot="torso", A=6, O=22, power=0.40, feet=["foot", "foot_left"],
        alive_rule=RS_ALIVE_Z_AND_PITCH, alive_z=0.8, alive_bonus=1.0, cost_scale=1.0,
        initial_z=None, coef={"foot_joint": 30.0, "foot_left_joint": 30.0}, reward_threshold=2500.0),
    "RoboschoolReacher-v1": dict(          # gym_reacher.py: zero gravity, arm + target (two top-level bodies), never done
        xml="reacher.xml", robot="body0", A=2, O=9, power=1.0, feet=[],
        alive_rule=RS_ALIVE_ALWAYS, alive_z=0.0, alive_bonus=0.0, cost_scale=1.0,
        initial_z=None, coef={"joint0": 0.05, "joint1": 0.05}, no_floor=True, timestep=0.0165, frame_skip=1, gravity=0.0,
        act_joints=["joint0", "joint1"], max_contacts=4, reward_threshold=18.0, max_episode_steps=150,
        task=dict(kind=RS_TASK_REACHER, dofs=["joint0", "joint1", "target_x", "target_y"], link="fingertip", link2="target",
                  reset=[("target_x", 0.0, 0.27), ("target_y", 0.0, 0.27), ("joint0", 0.0, 3.14), ("joint1", 0.0, 3.14)])),
    "RoboschoolHalfCheetah-v1": dict(      # gym_mujoco_walkers.py:40-59
        xml="half_cheetah.xml", robot="torso", A=6, O=26, power=0.90,
        feet=["ffoot", "fshin", "fthigh", "bfoot", "bshin", "bthigh"],
        alive_rule=RS_ALIVE_CHEETAH, alive_z=0.0, alive_bonus=1.0, cost_scale=1.0, initial_z=None,
        coef={"bthigh": 120.0, "bshin": 90.0, "bfoot": 60.0, "fthigh": 140.0, "fshin": 60.0, "ffoot": 30.0},
        reward_threshold=3000.0),
    "RoboschoolAnt-v1": dict(
        xml="ant.xml", robot="torso", A=8, O=28, power=2.5,
        feet=["front_left_foot", "front_right_foot", "left_back_foot", "right_back_foot"],
        alive_rule=RS_ALIVE_Z, alive_z=0.26, alive_bonus=1.0, cost_scale=1.0,
        initial_z=None, coef={}, reward_threshold=2500.0),
    "RoboschoolHumanoid-v1": dict(
        xml="humanoid_symmetric.xml", robot="torso", A=17, O=44, power=0.41,
        feet=["right_foot", "left_foot"],
        alive_rule=RS_ALIVE_Z, alive_z=0.78, alive_bonus=2.0, cost_scale=4.25,
        initial_z=0.8,
        coef={"abdomen_z": 100, "abdomen_y": 100, "abdomen_x": 100,
              "right_hip_x": 100, "right_hip_z": 100, "right_hip_y": 300, "right_knee": 200,
              "left_hip_x": 100, "left_hip_z": 100, "left_hip_y": 300, "left_knee": 200,
              "right_shoulder1": 75, "right_shoulder2": 75, "right_elbow": 75,
              "left_shoulder1": 75, "left_shoulder2": 75, "left_elbow": 75},
        reset_base=dict(z=1.4, yaw_spread=math.pi / 16), reward_threshold=3500.0),
    # Humanoid running to a flag that moves when reached or after 150 steps (gym_humanoid_flagrun.py:6-45):
    # same model and physics; electricity cost halved (:14); stadium-centred target box (scene_stadium.py:6-7)
    "RoboschoolHumanoidFlagrun-v1": dict(
        xml="humanoid_symmetric.xml", robot="torso", A=17, O=44, power=0.41,
        feet=["right_foot", "left_foot"],
        alive_rule=RS_ALIVE_Z, alive_z=0.78, alive_bonus=2.0, cost_scale=4.25, electricity_scale=0.5,
        initial_z=0.8,
        coef={"abdomen_z": 100, "abdomen_y": 100, "abdomen_x": 100,
              "right_hip_x": 100, "right_hip_z": 100, "right_hip_y": 300, "right_knee": 200,
              "left_hip_x": 100, "left_hip_z": 100, "left_hip_y": 300, "left_knee": 200,
              "right_shoulder1": 75, "right_shoulder2": 75, "right_elbow": 75,
              "left_shoulder1": 75, "left_shoulder2": 75, "left_elbow": 75},
        reset_base=dict(z=1.4, yaw_spread=math.pi / 16), reward_threshold=2000.0,
        flag=dict(halflen=105 * 0.25, halfwidth=50 * 0.25, timeout=600)),
    # Flagrun + stand-up / roll-over starts chosen by a per-env curriculum, potential "leak" alive bonus, crawl-proof potential
    # (gym_humanoid_flagrun.py:73-169), and the flying cube that pelts the robot every 30 frames (:81-118,
    # models_household/cube.urdf: 0.1 m box, 1.2 kg): a second free body per env with box-floor and box-vs-geom contacts.
    "RoboschoolHumanoidFlagrunHarder-v1": dict(
        xml="humanoid_symmetric.xml", robot="torso", A=17, O=44, power=0.41,
        feet=["right_foot", "left_foot"],
        alive_rule=RS_ALIVE_Z, alive_z=0.78, alive_bonus=2.0, cost_scale=4.25, electricity_scale=0.5,
        initial_z=0.8,
        coef={"abdomen_z": 100, "abdomen_y": 100, "abdomen_x": 100,
              "right_hip_x": 100, "right_hip_z": 100, "right_hip_y": 300, "right_knee": 200,
              "left_hip_x": 100, "left_hip_z": 100, "left_hip_y": 300, "left_knee": 200,
              "right_shoulder1": 75, "right_shoulder2": 75, "right_elbow": 75,
              "left_shoulder1": 75, "left_shoulder2": 75, "left_elbow": 75},
        reset_base=dict(z=1.4, yaw_spread=math.pi / 16), reward_threshold=None,
        flag=dict(halflen=105 * 0.25, halfwidth=50 * 0.25, timeout=600, harder=True, pelvis="pelvis"),
        # cube.urdf:8,16,22 -- inertia is NOT the file's (0.001): load_urdf passes no URDF_USE_INERTIA_FROM_FILE flag
        # (physics-bullet.cpp:70-93), so Bullet recomputes it from the collision shape's AABB box; lateral friction is the
        # URDF importer's default (no <contact> element)
        cube=dict(half=(0.05, 0.05, 0.05), mass=1.2, start=(-1.5, 0.0, 0.05), friction=0.5)),
    # BASELINE.json configs[0] and its siblings (gym_pendulums.py): empty scene, dt 0.0165, frame_skip 1, no contacts
    "RoboschoolInvertedPendulum-v1": dict(
        xml="inverted_pendulum.xml", robot="cart", A=1, O=5, power=1.0, feet=[],
        alive_rule=RS_ALIVE_ALWAYS, alive_z=0.0, alive_bonus=0.0, cost_scale=1.0,
        initial_z=None, coef={"slider": 100.0}, no_floor=True, timestep=0.0165, frame_skip=1,
        act_joints=["slider"], max_contacts=4, reward_threshold=950.0,
        task=dict(kind=RS_TASK_PENDULUM, dofs=["slider", "hinge"], reset=[("hinge", 0.0)])),
    "RoboschoolInvertedPendulumSwingup-v1": dict(
        xml="inverted_pendulum.xml", robot="cart", A=1, O=5, power=1.0, feet=[],
        alive_rule=RS_ALIVE_ALWAYS, alive_z=0.0, alive_bonus=0.0, cost_scale=1.0,
        initial_z=None, coef={"slider": 100.0}, no_floor=True, timestep=0.0165, frame_skip=1,
        act_joints=["slider"], max_contacts=4, reward_threshold=800.0,
        # As registered, the reference's Swingup class is NOT a swing-up task: `swingup = True` is a class attribute that
        # RoboschoolInvertedPendulum.__init__(self, swingup=False) overrides on the instance (gym_pendulums.py:77-79,126-127),
        # so gym.make('RoboschoolInvertedPendulumSwingup-v1') behaves like InvertedPendulum-v1.  The fixture made from the
        # reference's own class pins that; set task swingup=1 and reset centre 3.1415 on a cloned model for the real thing.
        task=dict(kind=RS_TASK_PENDULUM, swingup=0, dofs=["slider", "hinge"], reset=[("hinge", 0.0)])),
    "RoboschoolInvertedDoublePendulum-v1": dict(
        xml="inverted_double_pendulum.xml", robot="cart", A=1, O=9, power=1.0, feet=[],
        alive_rule=RS_ALIVE_ALWAYS, alive_z=0.0, alive_bonus=0.0, cost_scale=1.0,
        initial_z=None, coef={"slider": 200.0}, no_floor=True, timestep=0.0165, frame_skip=1,
        act_joints=["slider"], max_contacts=4, reward_threshold=9100.0,
        task=dict(kind=RS_TASK_DOUBLE_PENDULUM, dofs=["slider", "hinge", "hinge2"], link="pole2",
                  reset=[("hinge", 0.0), ("hinge2", 0.0)])),
}


def apply_task(m: ModelDesc, spec: dict) -> ModelDesc:
    """Fill the task section of the descriptor from an ENV_SPECS entry."""
    jn = m.meta["joint_names"]
    ln = m.meta["link_names"]
    # ordered_joints: non-"ignore" joints in link order (gym_mujoco_xml_env.py:63-70)
    if "act_joints" in spec:
        ordered = [jn.index(n) for n in spec["act_joints"]]
    else:
        ordered = [i for i, n in enumerate(jn) if n is not None and not n.startswith("ignore")]
    assert len(ordered) == spec["A"], (len(ordered), spec["A"])
    m.n_act = spec["A"]
    for k, li in enumerate(ordered):
        m.act_dof[k] = m.dof_idx[li]
        m.act_gain[k] = spec["power"] * float(spec["coef"].get(jn[li], 100.0))
    m.meta["ordered_joints"] = [jn[li] for li in ordered]
    m.meta["ordered_joint_links"] = ordered
    # feet
    m.n_feet = len(spec["feet"])
    for k, f in enumerate(spec["feet"]):
        m.foot_link[k] = ln.index(f)
    # parts dict de-duplicates by name (gym_mujoco_xml_env.py:57-62); root_part is NOT in parts
    seen, parts = set(), []
    for i, n in enumerate(ln):
        if n in seen:
            parts = [p for p in parts if ln[p] != n]
        seen.add(n)
        parts.append(i)
    m.n_parts = len(parts)
    for k, p in enumerate(parts):
        m.part_link[k] = p
    # robot body
    if m.meta["base_name"] == spec["robot"] and spec["robot"] not in ln:
        m.body_link = -1
    else:
        m.body_link = ln.index(spec["robot"])
    m.obs_dim = spec["O"]
    m.alive_rule = spec["alive_rule"]
    m.alive_z = spec["alive_z"]
    m.alive_bonus = spec["alive_bonus"]
    m.electricity_cost = -2.0 * spec["cost_scale"] * spec.get("electricity_scale", 1.0)   # gym_forward_walker.py:83, gym_mujoco_walkers.py:86, gym_humanoid_flagrun.py:14
    m.stall_torque_cost = -0.1 * spec["cost_scale"]      # :84 / :87
    m.foot_collision_cost = -1.0                         # :85
    m.joints_at_limit_cost = -0.2                        # :87
    m.initial_z = -1.0 if spec["initial_z"] is None else spec["initial_z"]
    m.walk_target_x, m.walk_target_y = 1e3, 0.0          # gym_forward_walker.py:13-14
    m.reset_joint_spread = 0.1                           # gym_forward_walker.py:26, gym_pendulums.py:24,89
    task = spec.get("task")
    if task:
        m.task_kind = task["kind"]
        m.task_swingup = task.get("swingup", 0)
        for k, n in enumerate(task["dofs"]):
            m.task_dof[k] = m.dof_idx[jn.index(n)]
        m.task_link = ln.index(task["link"]) if "link" in task else -1
        m.task_link2 = ln.index(task["link2"]) if "link2" in task else -1
        m.n_reset = len(task["reset"])
        for k, r in enumerate(task["reset"]):
            m.reset_dof[k] = m.dof_idx[jn.index(r[0])]
            m.reset_center[k] = r[1]
            m.reset_spread[k] = r[2] if len(r) > 2 else 0.1
    else:   # forward walkers: every ordered joint, U(-0.1, 0.1) about 0 (gym_forward_walker.py:25-26)
        m.task_kind = RS_TASK_WALKER
        m.n_reset = m.n_act
        for k in range(m.n_act):
            m.reset_dof[k] = m.act_dof[k]
            m.reset_center[k] = 0.0
            m.reset_spread[k] = 0.1
    rb = spec.get("reset_base")
    if rb:
        m.reset_set_base = 1
        m.arr("reset_base_pos")[:] = [0.0, 0.0, rb["z"]]
        m.arr("reset_base_quat")[:] = [0, 0, 0, 1]
        m.reset_yaw_spread = rb["yaw_spread"]
    fl = spec.get("flag")
    if fl:
        m.flag_task = 2 if fl.get("harder") else 1
        if fl.get("harder"):
            m.task_link = ln.index(fl["pelvis"])
        m.flag_timeout_steps = int(fl["timeout"] // m.frame_skip)   # 600/frame_skip (gym_humanoid_flagrun.py:35)
        m.flag_halflen, m.flag_halfwidth = fl["halflen"], fl["halfwidth"]
    cb = spec.get("cube")
    if cb:
        m.has_cube = 1
        hx, hy, hz = cb["half"]
        m.arr("cube_half")[:] = [hx, hy, hz]
        m.cube_mass = cb["mass"]
        lx, ly, lz = 2 * hx, 2 * hy, 2 * hz     # btCompoundShape::calculateLocalInertia: box of the AABB
        m.arr("cube_inertia")[:] = [cb["mass"] / 12.0 * (ly * ly + lz * lz), cb["mass"] / 12.0 * (lx * lx + lz * lz), cb["mass"] / 12.0 * (lx * lx + ly * ly)]
        m.cube_friction = cb["friction"]
        m.cube_break_thresh = 0.02 * math.sqrt(hx * hx + hy * hy + hz * hz)   # gContactBreakingThreshold x getAngularMotionDisc()
        m.arr("cube_start")[:] = list(cb["start"])
    m.max_episode_steps = spec.get("max_episode_steps", 1000)
    m.max_contacts = spec.get("max_contacts", 16)
    m.meta["env_id"] = spec.get("env_id", "")
    m.meta["reward_threshold"] = spec.get("reward_threshold")
    return m


def compile_env(env_id, assets_dir, native=False, **kw):
    """MJCF (from a roboschool ``mujoco_assets`` dir) -> full descriptor for ``env_id``.
    ``native``: compile through the library's C entry ``rs_model_compile`` instead of the Python compiler (same result)."""
    spec = dict(ENV_SPECS[env_id], env_id=env_id)
    comp = mjcf.compile_mjcf_native if native else mjcf.compile_mjcf
    m = comp(os.path.join(assets_dir, spec["xml"]),
                          has_floor=not spec.get("no_floor", False),
                          timestep=spec.get("timestep", 0.0165 / 4), gravity=spec.get("gravity", 9.8),
                          frame_skip=spec.get("frame_skip", 4), **kw)
    return apply_task(m, spec)


def load_env_model(env_id):
    """Load the pre-compiled descriptor shipped in ``roboschool_b200/models``.

    Generated by ``tools/compile_models.py`` from the reference's MJCF files; the
    GPU box has no /root/reference, so the compiled descriptors travel instead.
    """
    path = os.path.join(MODEL_DIR, env_id + ".json")
    if not os.path.exists(path):
        raise FileNotFoundError("no compiled model for %s (run tools/compile_models.py)" % env_id)
    return ModelDesc.load(path)
