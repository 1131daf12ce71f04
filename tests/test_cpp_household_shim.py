"""The drop-in boundary: the reference's UNMODIFIED Python env classes (from /root/reference) run on top
of roboschool_b200.cpp_household.  On this GPU-less box the shim's backend is swapped for the oracle
(tests/oracle_backend.py); the shim logic under test -- load_mjcf, joints/parts ordering, torque
caching, step(frame_skip), pose/speed readback, contact_list, reset_current_position,
set_pose_and_speed -- is identical for the CUDA backend."""
import os
import sys
import warnings

import numpy as np
import pytest

from tests.conftest import HAVE_REF

pytestmark = pytest.mark.skipif(not HAVE_REF, reason="reference Python sources not present")


@pytest.fixture()
def ref_roboschool(monkeypatch, oracle_lib):
    from tools.gen_golden_epilogue import install_stub_gym
    import roboschool_b200.cpp_household as shim
    from tests.oracle_backend import OracleBackend
    monkeypatch.setattr(shim, "_make_backend", lambda model, n_envs=1, device=0: OracleBackend(model, n_envs))
    shim.World._model_cache.clear()
    install_stub_gym()
    for k in [k for k in sys.modules if k == "roboschool" or k.startswith("roboschool.")]:
        del sys.modules[k]
    monkeypatch.syspath_prepend("/root/reference")
    sys.modules["roboschool.cpp_household"] = shim
    warnings.simplefilter("ignore")
    import roboschool
    yield roboschool
    for k in [k for k in sys.modules if k == "roboschool" or k.startswith("roboschool.")]:
        del sys.modules[k]


class _Spec:
    max_episode_steps = 1000


@pytest.mark.parametrize("cls,env_id", [("RoboschoolHopper", "RoboschoolHopper-v1"), ("RoboschoolAnt", "RoboschoolAnt-v1"),
                                        ("RoboschoolHumanoid", "RoboschoolHumanoid-v1"), ("RoboschoolWalker2d", "RoboschoolWalker2d-v1"),
                                        ("RoboschoolHalfCheetah", "RoboschoolHalfCheetah-v1"),
                                        ("RoboschoolInvertedPendulum", "RoboschoolInvertedPendulum-v1"),
                                        ("RoboschoolInvertedDoublePendulum", "RoboschoolInvertedDoublePendulum-v1"),
                                        ("RoboschoolReacher", "RoboschoolReacher-v1")])
def test_reference_env_runs_on_shim_and_matches_fused_path(cls, env_id, ref_roboschool, oracle_lib):
    from roboschool_b200 import envspec
    env = getattr(ref_roboschool, cls)()
    env.spec = _Spec()
    env.seed(3)
    obs0 = env.reset()
    m = envspec.load_env_model(env_id)
    assert obs0.shape == (m.obs_dim,)
    if m.task_kind == 0:
        assert obs0.dtype == np.float32
        assert [j.name for j in env.ordered_joints] == m.meta["ordered_joints"]
    # take the shim's post-reset state as the start of a fused (batched-path) oracle rollout
    robot = env.cpp_robot
    s0 = robot._backend.get_state()[0]
    o = oracle_lib.OracleWorld(m, 1)
    o.set_state(0, s0)
    iz = getattr(env, "initial_z", None)
    o.set_episode(0, env.potential, iz if iz is not None else -1.0, np.zeros(m.n_feet), 0)
    rng = np.random.RandomState(0)
    for t in range(30):
        a = rng.uniform(-1, 1, m.n_act).astype(np.float32)
        ob, rew, done, _ = env.step(a)
        oo, orw, od = o.step(a[None].astype(np.float64))
        assert np.abs(ob - oo[0]).max() < 1e-5, (t, np.abs(ob - oo[0]).max())
        assert abs(rew - orw[0]) < 1e-4
        assert bool(done) == bool(od[0])


def test_pose_math_roundtrip():
    import roboschool_b200.cpp_household as shim
    p = shim.Pose()
    p.set_rpy(0.3, -0.2, 1.1)
    r, pi, y = p.rpy()
    assert abs(r - 0.3) < 1e-12 and abs(pi + 0.2) < 1e-12 and abs(y - 1.1) < 1e-12
    p.set_xyz(1, 2, 3); p.move_xyz(1, 0, -1)
    assert p.xyz() == (2, 2, 2)
    q = shim.Pose(); q.set_rpy(0, 0, 0.5)
    q.rotate_z(0.25)
    assert abs(q.rpy()[2] - 0.75) < 1e-12
    a = shim.Pose(); a.set_rpy(0, 0, np.pi / 2); a.set_xyz(1, 0, 0)
    b = shim.Pose(); b.set_xyz(1, 0, 0)
    c = a.dot(b)
    assert np.allclose(c.xyz(), (1, 1, 0), atol=1e-12)


def test_load_failure_is_silent_like_reference(capsys):
    import roboschool_b200.cpp_household as shim
    w = shim.World(9.8, 0.0165 / 4)
    assert w.load_mjcf("/nonexistent/robot.xml") == []
    assert "cannot load MJCF" in capsys.readouterr().err
    assert w.step(4) is False


def test_joint_motor_state_machine_on_shim(ref_roboschool, oracle_lib):
    """Joint.set_servo_target / set_target_speed / set_motor_torque (python-binding.cpp:654-665, physics-bullet.cpp:497-537)
    through the shim on a reference env's own joints: a position servo pulls a knee towards its target, a velocity motor
    spins a hip at the commanded speed, and the next set_motor_torque switches the motor off again (first_torque_call)
    so that the joint is torque-controlled as before."""
    env = ref_roboschool.RoboschoolHumanoid()
    env.spec = _Spec()
    env.seed(1)
    env.reset()
    world = env.scene.cpp_world
    knee, hip = env.jdict["right_knee"], env.jdict["left_hip_y"]
    for j in env.ordered_joints:
        j.set_motor_torque(0.0)
    world.step(1)                                        # the server's time step is now timestep * 1
    q0 = knee.current_position()[0]
    knee.set_servo_target(-1.0, 0.1, 0.4, 80.0)
    hip.set_target_speed(-1.5, 0.5, 200.0)
    assert knee._first_torque_call and not knee._torque_need_repeat
    for k in range(12):
        world.step(4)
    assert knee.current_position()[0] < q0 - 0.3         # moved well towards -1.0
    assert abs(hip.current_position()[1] + 1.5) < 0.4     # spinning at about the commanded speed
    knee.set_motor_torque(0.0); hip.set_motor_torque(0.0)
    assert not knee._first_torque_call and knee._torque_need_repeat
    b = env.cpp_robot._backend
    s_before = b.get_state().copy()
    world.step(1)
    # same step on a world that never had motors: identical (the motor rows are gone, not merely weak)
    from tests.oracle_backend import OracleBackend
    ref = OracleBackend(env.cpp_robot._model, 1)
    ref.set_state(s_before, clear_contacts=True)
    b2 = OracleBackend(env.cpp_robot._model, 1); b2.set_state(s_before, clear_contacts=True)
    b2.set_joint_motor(knee._dof, 0.1, 0.4, -1.0, 0.0, 0.0)
    ref.substeps(1, np.zeros(env.cpp_robot._model.n_dof)); b2.substeps(1, np.zeros(env.cpp_robot._model.n_dof))
    assert np.array_equal(ref.get_state(), b2.get_state())


@pytest.mark.parametrize("cls,env_id", [("RoboschoolAnt", "RoboschoolAnt-v1"), ("RoboschoolHopper", "RoboschoolHopper-v1")])
def test_multiplayer_stadium_scene_on_shim(cls, env_id, ref_roboschool, oracle_lib):
    """SURVEY 8f.4, the reference's own multiplayer loop (multiplayer.py:237-288 without the shared-memory pipes): three robots
    loaded into ONE cpp_household.World by the reference's MultiplayerStadiumScene (scene_stadium.py:23-29: player n runs in lane
    y = n - 1), actions applied to all, ONE scene.global_step(), then every env's step() reads its own result.  Each robot
    reproduces the single-player rollout of the same seed and actions shifted into its lane (the robots are three envs of the
    batch: no inter-robot contacts -- DESIGN.md section 7)."""
    from roboschool.scene_stadium import MultiplayerStadiumScene
    Cls = getattr(ref_roboschool, cls)
    scene = MultiplayerStadiumScene(gravity=9.8, timestep=0.0165 / 4, frame_skip=4)
    envs = []
    for n in range(3):
        e = Cls()
        e.spec = _Spec()
        e.scene = scene
        e.player_n = n
        envs.append(e)
    scene.episode_restart()
    obs = []
    for n, e in enumerate(envs):
        e.seed(5 + n)
        obs.append(e.reset())
    assert len(scene.multiplayer_robots) == 3
    # the same three episodes, single player
    solo = []
    for n in range(3):
        e = Cls()
        e.spec = _Spec()
        e.seed(5 + n)
        o0 = e.reset()
        solo.append(e)
        lane = n - 1
        assert abs(envs[n].cpp_robot.root_part.pose().xyz()[1] - (e.cpp_robot.root_part.pose().xyz()[1] + lane)) < 1e-6
        if lane == 0:
            assert np.abs(o0 - obs[n]).max() < 1e-6
    rng = np.random.RandomState(1)
    for t in range(20):
        acts = [rng.uniform(-1, 1, e.action_space.shape[0]).astype(np.float32) for e in envs]
        for e, a in zip(envs, acts):
            e.apply_action(a)
        scene.global_step()
        for n, (e, a) in enumerate(zip(envs, acts)):
            ob, rew, done, _ = e.step(a)
            ob1, rew1, done1, _ = solo[n].step(a)
            # lanes differ in y only: joint, height, velocity and feet entries agree; the two angle-to-target entries see the
            # 1 m lane offset against a target 1 km away (< 1e-3 rad)
            keep = np.ones(ob.shape[0], dtype=bool)
            keep[1:3] = False
            assert np.abs(ob[keep] - ob1[keep]).max() < 1e-5, (t, n)
            assert np.abs(ob[1:3] - ob1[1:3]).max() < 2e-3
            assert bool(done) == bool(done1)
            by = e.body_xyz[1] - solo[n].body_xyz[1]
            assert abs(by - (n - 1)) < 1e-4, (t, n, by)


def test_flagrun_harder_with_thrown_cube_on_shim(ref_roboschool, oracle_lib):
    """BASELINE config 5 at the drop-in boundary: the reference's own RoboschoolHumanoidFlagrunHarder class
    (gym_humanoid_flagrun.py:72-169) on the shim.  Its reset loads models_household/cube.urdf through World.load_urdf -- the free
    box becomes the second body of the humanoid's world -- and its alive_bonus throws it with Robot.set_pose_and_speed every 30
    frames once the robot has been up for 100.  Driven by the shipped zoo policy so that the robot stays up: throws happen, every
    throw starts 4 m from the predicted robot position and 1 m above it at 20-31 m/s, the cube then FLIES (gravity + drag only,
    until it hits something), and the physics the class sees is the oracle's world with the cube (same state words)."""
    from roboschool_b200.policy import load_zoo_weights
    from oracle.oracle import mlp_forward
    env = ref_roboschool.RoboschoolHumanoidFlagrunHarder()
    env.spec = _Spec()
    env.seed(0)
    for k in range(64):            # the class draws its task from the GLOBAL numpy RNG (RepeatUnderlearnedTasks.decide_best_task):
        np.random.seed(k)          # take the first seed that gives the walking task -- throws need the robot on its feet
        obs = env.reset()
        if env.task == env.TASK_WALK:
            break
    assert env.task == env.TASK_WALK
    assert obs.shape == (44,)
    cube = env.aggressive_cube
    host = env.cpp_robot
    assert host._model.has_cube == 1 and host._backend.S == 13 + 2 * 17 + 13
    cube.query_position()
    assert np.allclose(cube.root_part.pose().xyz(), (-1.5, 0, 0.05), atol=1e-9)      # where the class loads it
    wts = load_zoo_weights("RoboschoolHumanoidFlagrunHarder-v1")
    throws, flights = [], 0
    calls = []
    orig = cube.set_pose_and_speed
    def spy(p, vx, vy, vz):
        calls.append((np.array(p.xyz()), np.array([vx, vy, vz])))
        return orig(p, vx, vy, vz)
    cube.set_pose_and_speed = spy
    prev_c = None
    for t in range(260):
        a = mlp_forward(wts, obs[None].astype(np.float64))[0]
        n_before = len(calls)
        obs, rew, done, _ = env.step(a)
        assert np.isfinite(obs).all() and np.isfinite(rew)
        c = np.asarray(host._backend.get_state())[0, -13:]
        if len(calls) > n_before:                      # thrown inside this step's alive_bonus (after the physics)
            pos, vel = calls[-1]
            assert np.allclose(c[0:3], pos, atol=1e-9) and np.allclose(c[10:13], vel.astype(np.float32), atol=0) and np.all(c[7:10] == 0)
            body = np.array(env.body_xyz)
            d = pos - body
            assert 3.0 < np.hypot(d[0], d[1]) < 5.0 and abs(d[2] - 1.0) < 0.2      # 4 m away (plus the aim prediction), 1 m above
            assert 19.0 < np.linalg.norm(vel) < 32.0
            throws.append(t)
        elif prev_c is not None and throws and t == throws[-1] + 1:
            # the first step after a throw is free flight: x(t + dt) = x + v dt up to gravity and 0.04 (1 + |v|) drag over 0.0165 s
            dt = 0.0165
            assert np.abs(c[0:3] - (prev_c[0:3] + prev_c[10:13] * dt)).max() < 0.02 * np.linalg.norm(prev_c[10:13]) * dt + 5e-3
            flights += 1
        prev_c = c
        if done:
            break
    assert len(throws) >= 3 and flights >= 3, (throws, flights)
    assert throws[0] == 120 and all(b - a == 30 for a, b in zip(throws, throws[1:])), throws     # alive_bonus runs before the frame counter is advanced: frames 120, 150, ...
