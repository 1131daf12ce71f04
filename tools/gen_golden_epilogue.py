"""Golden vectors for the obs/reward/done epilogue, produced by the REFERENCE'S OWN Python code.

The reference's env layer (roboschool/gym_forward_walker.py calc_state/step, gym_mujoco_walkers.py
alive_bonus, gym_mujoco_xml_env.py reset) is imported UNMODIFIED from /root/reference and driven
through a stub `gym` package and a stub `cpp_household` module whose World replays a recorded
trajectory of joint/body/contact state (taken from an oracle rollout).  What the stub restates (it
is C++ in the reference and cannot run here) is limited to Pose::rpy/xyz (python-binding.cpp:57-73)
and Joint::joint_current_relative_position (physics-bullet.cpp:547-566); everything downstream --
observation layout, clipping, alive/progress/electricity/joints-at-limit/feet-collision terms, the
one-step-stale feet_contact -- is the reference's code.

Writes tests/golden/epilogue_<env>.npz with the recorded inputs and the reference's outputs.
This script needs /root/reference (absent on the GPU box); the fixtures travel instead.

Usage: python tools/gen_golden_epilogue.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = "/root/reference"


# ---------------------------------------------------------------- stub gym
def install_stub_gym():
    gym = types.ModuleType("gym")

    class Env(object):
        metadata = {}
        spec = None

    class Box(object):
        def __init__(self, low, high, dtype=np.float32):
            self.low, self.high, self.dtype = low, high, dtype
            self.shape = np.asarray(low).shape

    gym.Env = Env
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")

    def np_random(seed=None):
        return np.random.RandomState(seed), seed
    seeding.np_random = np_random
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda **kw: None
    envs.registration = registration
    gym.spaces, gym.utils, gym.envs = spaces, utils, envs
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils), ("gym.utils.seeding", seeding),
                      ("gym.envs", envs), ("gym.envs.registration", registration)]:
        sys.modules[name] = mod


# ---------------------------------------------------------------- stub cpp_household replaying a trajectory
class Pose(object):
    def __init__(self):
        self.x = self.y = self.z = 0.0
        self.qx = self.qy = self.qz = 0.0
        self.qw = 1.0

    def set_xyz(self, x, y, z): self.x, self.y, self.z = x, y, z
    def move_xyz(self, x, y, z): self.x += x; self.y += y; self.z += z
    def set_rpy(self, r, p, y): pass
    def set_quaternion(self, x, y, z, w): self.qx, self.qy, self.qz, self.qw = x, y, z, w
    def xyz(self): return (self.x, self.y, self.z)

    def rpy(self):  # python-binding.cpp:60-73 (btScalar = double)
        qx, qy, qz, qw = self.qx, self.qy, self.qz, self.qw
        sqw, sqx, sqy, sqz = qw * qw, qx * qx, qy * qy, qz * qz
        t2 = -2.0 * (qx * qz - qy * qw) / (sqx + sqy + sqz + sqw)
        yaw = np.arctan2(2.0 * (qx * qy + qz * qw), (sqx - sqy - sqz + sqw))
        roll = np.arctan2(2.0 * (qy * qz + qx * qw), (-sqx - sqy + sqz + sqw))
        t2 = min(1.0, max(-1.0, t2))
        return (roll, np.arcsin(t2), yaw)


class Replay(object):
    """Trajectory store: per frame t, link poses (1+L,10), joint q/qd, per-foot (ground, other) flags."""
    t = 0
    poses = None
    q = None
    qd = None
    foot_flags = None


class Thingy(object):
    def __init__(self, name, body_index, replay, foot_index=None):
        self.name, self.b, self.R, self.foot_index = name, body_index, replay, foot_index

    def pose(self):
        p = Pose()
        row = self.R.poses[self.R.t][self.b]
        p.set_xyz(*row[0:3])
        p.set_quaternion(*row[3:7])
        return p

    def speed(self):
        return tuple(self.R.poses[self.R.t][self.b][7:10])

    def contact_list(self):
        out = []
        if self.foot_index is not None:
            g, o = self.R.foot_flags[self.R.t][self.foot_index]
            if g: out.append(Thingy("floor", -1, self.R))
            if o: out.append(Thingy("some_other_part", -1, self.R))
        return out

    def __hash__(self): return hash((self.name, self.b))
    def __eq__(self, other): return (self.name, self.b) == (other.name, other.b)


class Joint(object):
    def __init__(self, name, dof, lo, hi, revolute, replay):
        self.name, self.dof, self.R, self.revolute = name, dof, replay, revolute
        self.lo, self.hi = np.float32(lo), np.float32(hi)      # float joint_limit1/2 (household.h:78-79)
        self.torque = 0.0

    def set_motor_torque(self, t): self.torque = np.float32(t)
    def reset_current_position(self, pos, vel): pass
    def limits(self): return (float(self.lo), float(self.hi), 0.0, 0.0)

    def current_position(self):
        return (float(np.float32(self.R.q[self.R.t][self.dof])), float(np.float32(self.R.qd[self.R.t][self.dof])))

    def current_relative_position(self):   # physics-bullet.cpp:547-566, float arithmetic
        rpos = np.float32(self.R.q[self.R.t][self.dof])
        rspeed = np.float32(self.R.qd[self.R.t][self.dof])
        if self.lo < self.hi:
            pos_mid = np.float32(0.5 * (float(self.lo) + float(self.hi)))
            rpos = np.float32(2) * (rpos - pos_mid) / (self.hi - self.lo)
        rspeed = rspeed * (np.float32(0.1) if self.revolute else np.float32(0.5))   # joint_max_velocity == 0 for MJCF
        return (float(rpos), float(rspeed))


class Robot(object):
    def __init__(self, root, parts, joints):
        self.root_part, self.parts, self.joints = root, parts, joints

    throws = None      # list that receives (frame index, x, y, z, vx, vy, vz) of every set_pose_and_speed on the cube
    replay = None

    def query_position(self): pass
    def set_pose(self, p): pass

    def set_pose_and_speed(self, p, vx, vy, vz):
        if self.throws is not None:
            self.throws.append((self.replay.t,) + tuple(p.xyz()) + (float(np.float32(vx)), float(np.float32(vy)), float(np.float32(vz))))


def make_cpp_household(model, replay, foot_names):
    mod = types.ModuleType("roboschool.cpp_household")
    meta = model.meta

    class World(object):
        def __init__(self, gravity, timestep):
            self.ts = 0

        def clean_everything(self): pass
        def set_glsl_path(self, p): pass
        def test_window_big_caption(self, *a): pass
        def test_window_print(self, *a): pass
        def test_window_history_reset(self): pass
        def test_window_history_advance(self): pass
        def test_window_observations(self, *a): pass
        def test_window_actions(self, *a): pass
        def test_window_rewards(self, *a): pass
        def test_window_score(self, *a): pass
        def test_window(self): return True
        def load_thingy(self, *a): return Thingy("stadium", -1, replay)
        def new_camera_free_float(self, *a): return None
        def debug_sphere(self, *a): return None
        def load_urdf(self, *a):
            r = Robot(Thingy("cube", -1, replay), [], [])
            r.throws, r.replay = replay.throws, replay
            return r

        def load_mjcf(self, fn):
            if fn.endswith("ground_plane.xml"):
                return [Robot(Thingy("floor", -1, replay), [], [])]
            parts = []
            for i, n in enumerate(meta["link_names"]):
                parts.append(Thingy(n, i + 1, replay, foot_names.index(n) if n in foot_names else None))
            joints = []
            for i, jn in enumerate(meta["joint_names"]):
                if jn is None:
                    continue
                joints.append(Joint(jn, model.dof_idx[i], model.lim_lo[i], model.lim_hi[i], model.jtype[i] == 1, replay))
            return [Robot(Thingy(meta["base_name"], 0, replay), parts, joints)]

        def step(self, frame_skip):
            replay.t += 1
            return False

    mod.World, mod.Pose = World, Pose
    return mod


class CounterRNG(object):
    """np_random stand-in for the reference classes: replays the counter-based stream the oracle / CUDA path draw from
    (oracle/rs_oracle.c rs_uniform), in the order the reference consumes np_random at reset and in flag_reposition:
    A joint draws (k = 0..A-1), the yaw draw (k = 1000), then the flag draws (k = 2000, 2001, ...)."""
    def __init__(self, seed, env_id, episode, n_act):
        from oracle.oracle import lib
        self.h = lib().oracle_hash
        self.seed, self.env_id, self.episode, self.A, self.calls = seed, env_id, episode, n_act, 0

    def _u(self):
        c = self.calls
        self.calls += 1
        k = c if c < self.A else (1000 if c == self.A else 2000 + (c - self.A - 1))
        return float(self.h(self.seed, self.env_id, self.episode, k) >> 8) * (1.0 / 16777216.0)

    def uniform(self, low=0.0, high=1.0, size=None):
        if size is not None:      # FlagrunHarder's cube noise: consumes one draw per element
            n = int(np.prod(size))
            return np.array([low + (high - low) * self._u() for _ in range(n)]).reshape(size)
        return low + (high - low) * self._u()

    def choice_u(self):
        """the draw RepeatUnderlearnedTasks.decide_best_task takes (np.random.choice in the reference): stream slot 1500"""
        return float(self.h(self.seed, self.env_id, self.episode, 1500) >> 8) * (1.0 / 16777216.0)

    def randint(self, n):
        assert n == 2
        return 0 if self._u() < 0.5 else 1


def generate(env_id, cls_name, n_steps=60, seed=3, counter_rng=False, multi_episode=False, zoo_policy=False, tag=""):
    from roboschool_b200 import envspec
    from oracle.oracle import OracleWorld
    m = envspec.compile_env(env_id, os.path.join(REF, "roboschool", "mujoco_assets"))
    spec = envspec.ENV_SPECS[env_id]
    o = OracleWorld(m, 1)
    o.reset(seed=seed)
    rng = np.random.RandomState(seed)
    acts = rng.uniform(-1.3, 1.3, (n_steps, m.n_act))      # some outside [-1,1]: clip vs electricity cost
    R = Replay()
    R.poses, R.q, R.qd, R.foot_flags = [], [], [], []
    R.throws = []
    oracle_cube = []      # cube pos3 + vel3 after every step (FlagrunHarder)
    if zoo_policy:        # actions of the reference's own pretrained policy (+ noise): the robot has to stay on its feet
        from roboschool_b200.policy import load_zoo_weights    # beyond frame 100 for the cube to be thrown at it
        from oracle.oracle import mlp_forward
        wts = load_zoo_weights(env_id)
    oracle_obs, oracle_rew, oracle_done, oracle_r5 = [], [], [], []

    def record():
        s = o.get_state(0)
        nd = m.n_dof
        if m.has_cube:
            oracle_cube.append(np.concatenate([s[-13:-10], s[-3:]]))
            s = s[:-13]
        R.poses.append(o.link_poses(0).copy())
        R.q.append(s[-2 * nd:-nd].copy())
        R.qd.append(s[-nd:].copy())
        g, ot = o.foot_flags(0)
        R.foot_flags.append(list(zip(g.tolist(), ot.tolist())))

    record()
    obs0 = o.obs()[0][0].copy()
    reset_after = []          # steps after which the env was reset (multi_episode): the next frame is the reset state
    oracle_reset_obs = {}
    cur_obs = o.obs()[0]
    for t in range(n_steps):
        if zoo_policy:
            acts[t] = mlp_forward(wts, cur_obs)[0] + 0.05 * rng.standard_normal(m.n_act)
        ob, rw, dn = o.step(acts[t:t + 1])
        cur_obs = ob
        record()
        oracle_obs.append(ob[0].copy()); oracle_rew.append(rw[0]); oracle_done.append(dn[0]); oracle_r5.append(o.rewards5(0))
        if multi_episode and dn[0]:
            o.L.oracle_reset(o.h, 0, seed, 0)
            record()
            reset_after.append(t)
            oracle_reset_obs[t] = o.obs()[0][0].copy()
            cur_obs = o.obs()[0]

    # ---- run the reference's Python over the recorded trajectory
    install_stub_gym()
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k.startswith("roboschool") and not k.startswith("roboschool_b200")]:
        del sys.modules[k]
    stub = make_cpp_household(m, R, spec["feet"])
    sys.modules["roboschool.cpp_household"] = stub
    import warnings
    warnings.simplefilter("ignore")
    import roboschool
    roboschool.cpp_household = stub
    env = getattr(roboschool, cls_name)()

    class Spec(object):
        max_episode_steps = 1000
    env.spec = Spec()
    R.t = 0
    episode = 1
    if counter_rng:
        env.np_random = CounterRNG(seed, 0, episode, m.n_act)
    if multi_episode:     # RepeatUnderlearnedTasks draws from the global numpy RNG: route it to the counter stream
        def choice(rng, p):
            cdf = np.cumsum(np.asarray(p, dtype=np.float64)); cdf = cdf / cdf[-1]
            return min(int(np.count_nonzero(cdf <= env.np_random.choice_u())), len(cdf) - 1)
        np_choice, np.random.choice = np.random.choice, choice
    ref_flags, ref_task, ref_reset_obs = [], [], {}
    ref_obs0 = env.reset()
    ref_obs, ref_rew, ref_done, ref_r5 = [], [], [], []
    for t in range(n_steps):
        ob, rw, dn, _ = env.step(acts[t])
        ref_obs.append(np.asarray(ob, dtype=np.float64)); ref_rew.append(rw); ref_done.append(dn); ref_r5.append((list(env.rewards) + [0.0] * 5)[:5])
        ref_flags.append([getattr(env, "walk_target_x", 0.0), getattr(env, "walk_target_y", 0.0), getattr(env, "flag_timeout", 0)])
        ref_task.append(getattr(env, "task", 0))
        if multi_episode and t in reset_after:
            episode += 1
            env.np_random = CounterRNG(seed, 0, episode, m.n_act)
            R.t += 1                       # the frame recorded right after the oracle's reset
            ref_reset_obs[t] = np.asarray(env.reset(), dtype=np.float64)
    if multi_episode:
        np.random.choice = np_choice
        for t in reset_after:
            assert np.abs(ref_reset_obs[t] - oracle_reset_obs[t]).max() < 1e-6, ("reset obs", t)
    throws = np.array(R.throws, dtype=np.float64).reshape(len(R.throws), 7)
    if m.has_cube:        # every throw of the reference class lands on the oracle's cube: same pose, same (float) velocity
        assert len(throws) >= 3 or not zoo_policy, "fixture without throws"
        oc = np.array(oracle_cube)
        for row in throws:
            fi = int(row[0])
            assert np.abs(oc[fi][:3] - row[1:4]).max() < 1e-9 and np.abs(oc[fi][3:] - row[4:7]).max() < 1e-12, ("throw", fi, oc[fi], row)
    out = os.path.join(ROOT, "tests", "golden", "epilogue_%s%s.npz" % (env_id, tag))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, seed=seed, actions=acts, poses=np.array(R.poses), q=np.array(R.q), qd=np.array(R.qd),
                        foot_flags=np.array(R.foot_flags, dtype=np.int32).reshape(len(R.foot_flags), -1, 2),
                        ref_obs0=np.asarray(ref_obs0, dtype=np.float64), ref_obs=np.array(ref_obs), ref_rew=np.array(ref_rew),
                        ref_done=np.array(ref_done), ref_rewards5=np.array(ref_r5), ref_flags=np.array(ref_flags, dtype=np.float64),
                        ref_task=np.array(ref_task, dtype=np.int32), reset_after=np.array(reset_after, dtype=np.int32),
                        ref_throws=throws,
                        ref_reset_obs=np.array([ref_reset_obs[t] for t in reset_after]).reshape(len(reset_after), m.obs_dim))
    d_obs = np.abs(np.array(ref_obs) - np.array(oracle_obs)).max()
    d_rew = np.abs(np.array(ref_rew) - np.array(oracle_rew)).max()
    print("%-26s steps=%d  |ref-oracle| obs %.2e rew %.2e obs0 %.2e done equal=%s" % (
        env_id, n_steps, d_obs, d_rew, np.abs(ref_obs0 - obs0).max(), np.array_equal(np.array(ref_done), np.array(oracle_done).astype(bool))))


if __name__ == "__main__":
    generate("RoboschoolInvertedPendulum-v1", "RoboschoolInvertedPendulum")
    generate("RoboschoolInvertedPendulumSwingup-v1", "RoboschoolInvertedPendulumSwingup")
    generate("RoboschoolInvertedDoublePendulum-v1", "RoboschoolInvertedDoublePendulum")
    generate("RoboschoolReacher-v1", "RoboschoolReacher")
    generate("RoboschoolHopper-v1", "RoboschoolHopper")
    generate("RoboschoolWalker2d-v1", "RoboschoolWalker2d")
    generate("RoboschoolHalfCheetah-v1", "RoboschoolHalfCheetah")
    generate("RoboschoolAnt-v1", "RoboschoolAnt")
    generate("RoboschoolHumanoid-v1", "RoboschoolHumanoid")
    # Flagrun: 170 steps so that the 150-step flag timeout (and the reposition after it) is part of the fixture
    generate("RoboschoolHumanoidFlagrun-v1", "RoboschoolHumanoidFlagrun", n_steps=170, counter_rng=True)
    # FlagrunHarder: several episodes (the task curriculum and the stand-up / roll-over starts are part of the fixture)
    generate("RoboschoolHumanoidFlagrunHarder-v1", "RoboschoolHumanoidFlagrunHarder", n_steps=420, counter_rng=True, multi_episode=True)
    # FlagrunHarder driven by the zoo policy, so that the robot walks past frame 100 and is pelted by the cube (throws, hits on
    # the robot, feet touching the cube) -- every set_pose_and_speed of the reference class is checked against the oracle's cube
    generate("RoboschoolHumanoidFlagrunHarder-v1", "RoboschoolHumanoidFlagrunHarder", n_steps=700, counter_rng=True, multi_episode=True,
             zoo_policy=True, tag="_cube")
