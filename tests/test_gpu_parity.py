"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerances: BASELINE.json's north_star asks for joint positions/velocities within a
per-step relative tolerance of 1e-4 in fp32 and exact contact-pair counts/indices per step."""
import numpy as np
import pytest

from roboschool_b200.envspec import load_env_model

pytestmark = pytest.mark.gpu

ENVS = ["RoboschoolHopper-v1", "RoboschoolAnt-v1", "RoboschoolHumanoid-v1", "RoboschoolInvertedPendulum-v1",
        "RoboschoolHumanoidFlagrun-v1", "RoboschoolWalker2d-v1", "RoboschoolHalfCheetah-v1", "RoboschoolInvertedDoublePendulum-v1", "RoboschoolReacher-v1",
        "RoboschoolHumanoidFlagrunHarder-v1"]
STEP_ENVS = ENVS
# Per env-step (= 4 sub-steps) error of the fp32 CUDA path against the fp64 oracle, relative to
# max(1, |state|_inf), over seeded random-action rollouts.  The dynamics are non-smooth (contact
# make/break, friction-cone and limit clamps), so a clamp that flips in one sub-step gives an outlier:
# the tolerance is stated on percentiles -- median <= 1e-5, >= 95 % of env-steps within the
# north-star's 1e-4, and every env-step within 0.2 (bounded divergence).
STEP_RTOL = 1e-4


def _worlds(env_id, n, seed):
    from roboschool_b200.world import BatchedWorld
    from oracle.oracle import OracleWorld
    m = load_env_model(env_id)
    w = BatchedWorld(m, n, 0)
    o = OracleWorld(m, n)
    w.reset(seed=seed)
    o.reset(seed=seed)
    return m, w, o


def _contact_keys(o, env):
    c = o.contacts(env)
    return c[:, 0].astype(np.int32)


@pytest.mark.parametrize("env_id", ENVS)
def test_reset_matches_oracle(env_id, oracle_lib):
    m, w, o = _worlds(env_id, 256, 3)
    assert np.abs(w.get_state() - o.get_state()).max() < 1e-6


PARITY_REPORT = {}     # env_id -> numbers of the teacher-forced run (written to gpurun_out/parity_report.json at the end)


def _write_parity_report():
    import json, os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.json"), "w") as f:
            json.dump(PARITY_REPORT, f, indent=1, sort_keys=True)


@pytest.mark.parametrize("env_id", STEP_ENVS)
def test_step_teacher_forced(env_id, oracle_lib):
    """Every step starts from the oracle's state AND the oracle's persistent contact manifolds (teacher forcing of the
    whole env, so each (env, step) is an independent sample).  Checked per (env, step):
      * contact manifolds after the step -- ids, per-manifold slots, count -- identical (north star: "contact-pair counts
        and indices match exactly per step"); the exact-match fraction is reported and must be >= 99.5 %: fp32 vs fp64
        can straddle a make/break threshold, nothing else may differ;
      * state within STEP_RTOL (percentile statement, see above);
      * reward, its five terms (rewards5) and the done flag against the oracle's for every sample whose manifolds match."""
    n, T = 128, 40
    m, w, o = _worlds(env_id, n, 5)
    rng = np.random.RandomState(1)
    errs, rew_errs = [], []
    n_samples = n_match = n_done_ok = n_done = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (n, m.n_act))
        so = o.get_state()
        w.set_state(so.astype(np.float32), clear_contacts=False)
        w.set_contacts(*o.contact_cache(max_pts=16))
        ep = [o.get_episode(i) for i in range(n)]
        w.set_episode(potential=[e[0] for e in ep], initial_z=[e[1] for e in ep],
                      feet_contact=np.array([e[2] for e in ep]).reshape(n, m.n_feet), frame=[e[3] for e in ep])
        if m.flag_task:
            fl = [o.get_flag(i) for i in range(n)]
            w.set_flags(target_xy=[[f[0], f[1]] for f in fl], timeout=[f[2] for f in fl], draws=[f[3] for f in fl])
        if m.flag_task == 2:
            w.set_task_state(np.stack([o.get_task_state(i) for i in range(n)]))
        for i in range(n):   # the oracle starts from the float32-rounded state as well
            o.set_state(i, so[i].astype(np.float32).astype(np.float64), clear_contacts=False)
        obs_g, rew_g, done_g = w.step_host(a.astype(np.float32))
        obs_o, rew_o, done_o = o.step(a.astype(np.float32).astype(np.float64))
        r5_g = w.rewards5()
        sg, so2 = w.get_state(), o.get_state()
        cnt_o, keys_o, _ = o.contact_cache(max_pts=16)
        match = np.zeros(n, dtype=bool)
        for i in range(n):
            kg, _ = w.contacts(i)
            match[i] = len(kg) == cnt_o[i] and np.array_equal(kg, keys_o[i, :cnt_o[i]])
        n_samples += n; n_match += int(match.sum())
        for i in np.nonzero(match)[0]:
            scale = max(1.0, np.abs(so2[i]).max())
            r = np.abs(sg[i] - so2[i]).max() / scale
            errs.append(r)
            n_done += 1; n_done_ok += int(bool(done_g[i]) == bool(done_o[i]))
            if r < STEP_RTOL:
                assert np.abs(obs_g[i] - obs_o[i]).max() < 2e-3
                # reward: the progress term is a position difference / 0.0165 s, so it amplifies the state error ~60x
                rew_errs.append(abs(float(rew_g[i]) - float(rew_o[i])))
                r5_o = o.rewards5(i)
                assert abs(r5_g[i, 0] - r5_o[0]) < 1e-5 or bool(done_g[i]) != bool(done_o[i]), ("alive", t, i, r5_g[i], r5_o)
                assert abs(r5_g[i, 4] - r5_o[4]) < 1e-6, ("feet_collision", t, i, r5_g[i], r5_o)
                assert abs(r5_g[i, 2] - r5_o[2]) < 1e-3 * max(1.0, abs(r5_o[2])), ("electricity", t, i, r5_g[i], r5_o)
        if m.flag_task:   # same flag bookkeeping (position, timeout, RNG draws) on both sides
            tg, tog, dg = w.get_flags()
            fo = np.array([o.get_flag(i) for i in range(n)])
            assert np.abs(tg[match] - fo[match, :2]).max() < 1e-9 and (tog[match] == fo[match, 2]).all() and (dg[match] == fo[match, 3]).all()
        if m.flag_task == 2:
            tsg = w.get_task_state()
            tso = np.stack([o.get_task_state(i) for i in range(n)])
            same = (tsg[:, [0, 1, 2, 5, 6, 7]] == tso[:, [0, 1, 2, 5, 6, 7]]).all(axis=1)
            assert (same | ~match).mean() >= 0.995      # on_ground / task history flip only with a straddled z threshold
    errs = np.array(errs); rew_errs = np.array(rew_errs)
    frac = n_match / n_samples
    PARITY_REPORT[env_id] = dict(samples=n_samples, manifold_exact_match=frac, state_err_median=float(np.median(errs)),
                                 state_err_p95=float(np.percentile(errs, 95)), state_err_max=float(errs.max()),
                                 frac_within_1e_4=float((errs <= STEP_RTOL).mean()),
                                 reward_err_median=float(np.median(rew_errs)), reward_err_p99=float(np.percentile(rew_errs, 99)),
                                 done_agree=n_done_ok / max(n_done, 1))
    print("\n[parity] %s: manifolds identical in %.4f of %d (env, step) samples; state err median %.2e p95 %.2e max %.2e; "
          "reward err median %.2e p99 %.2e; done agrees in %.4f" % (env_id, frac, n_samples, np.median(errs), np.percentile(errs, 95),
                                                                      errs.max(), np.median(rew_errs), np.percentile(rew_errs, 99), n_done_ok / max(n_done, 1)))
    _write_parity_report()
    assert np.median(errs) <= 1e-5, np.median(errs)
    assert (errs <= STEP_RTOL).mean() >= 0.95, (errs <= STEP_RTOL).mean()
    assert errs.max() < 0.2, errs.max()
    assert frac >= 0.995, "contact manifolds differ in %d of %d (env, step) samples" % (n_samples - n_match, n_samples)
    assert np.median(rew_errs) < 1e-3 and np.percentile(rew_errs, 99) < 5e-2, (np.median(rew_errs), np.percentile(rew_errs, 99))
    assert n_done_ok / max(n_done, 1) >= 0.998, n_done_ok / max(n_done, 1)


def test_cube_throw_flight_and_hits(oracle_lib):
    """FlagrunHarder's cube on the GPU (gym_humanoid_flagrun.py:81-118, BASELINE config 5): episodes advanced to frame 110 under
    the zoo policy on the oracle, then 45 teacher-forced steps covering the throw at frame 120 (five RNG draws ->
    set_pose_and_speed), the flight (drag, gravity), the hit on the robot (box vs capsule / sphere, two-body contact rows)
    and the cube coming to rest on the floor (box-plane manifold).  State compared includes the cube's 13 floats."""
    from roboschool_b200.policy import load_zoo_weights
    env_id = "RoboschoolHumanoidFlagrunHarder-v1"
    n = 128
    m, w, o = _worlds(env_id, n, 2)
    assert m.has_cube == 1 and w.S == m.state_dim == 13 + 2 * m.n_dof + 13
    wts = load_zoo_weights(env_id)
    obs = o.obs()[0]
    for t in range(110):
        obs, _, _ = o.step(oracle_lib.mlp_forward(wts, obs), auto_reset=True, seed=2, threads=8)
    S, nman = m.state_dim, m.n_geoms + m.n_pairs
    errs, thrown, cube_floor, cube_hits, n_match, n_samples = [], 0, 0, 0, 0, 0
    for t in range(45):
        a = oracle_lib.mlp_forward(wts, obs).astype(np.float32)
        so = o.get_state()
        w.set_state(so.astype(np.float32), clear_contacts=False)
        w.set_contacts(*o.contact_cache(max_pts=16))
        ep = [o.get_episode(i) for i in range(n)]
        w.set_episode(potential=[e[0] for e in ep], initial_z=[e[1] for e in ep],
                      feet_contact=np.array([e[2] for e in ep]).reshape(n, m.n_feet), frame=[e[3] for e in ep])
        fl = [o.get_flag(i) for i in range(n)]
        w.set_flags(target_xy=[[f[0], f[1]] for f in fl], timeout=[f[2] for f in fl], draws=[f[3] for f in fl])
        w.set_task_state(np.stack([o.get_task_state(i) for i in range(n)]))
        for i in range(n):
            o.set_state(i, so[i].astype(np.float32).astype(np.float64), clear_contacts=False)
        obs_g, rew_g, done_g = w.step_host(a)
        obs, rew_o, done_o = o.step(a.astype(np.float64), threads=8)
        sg, so2 = w.get_state(), o.get_state()
        thrown += int((np.abs(so2[:, S - 13:S - 10] - so[:, S - 13:S - 10]).max(axis=1) > 2.0).sum())
        cnt_o, keys_o, _ = o.contact_cache(max_pts=16)
        for i in range(n):
            kg, _ = w.contacts(i)
            n_samples += 1
            if not (len(kg) == cnt_o[i] and np.array_equal(kg, keys_o[i, :cnt_o[i]])):
                continue
            n_match += 1
            cube_floor += int((kg == nman).sum()); cube_hits += int((kg > nman).sum())
            r = np.abs(sg[i] - so2[i]).max() / max(1.0, np.abs(so2[i]).max())
            errs.append(r)
            if r < STEP_RTOL:
                assert np.abs(obs_g[i] - obs[i]).max() < 2e-3 and bool(done_g[i]) == bool(done_o[i])
    errs = np.array(errs)
    print("\n[cube] throws %d, cube-floor points %d, cube-robot points %d, manifolds identical %.4f, state err median %.2e p95 %.2e max %.2e"
          % (thrown, cube_floor, cube_hits, n_match / n_samples, np.median(errs), np.percentile(errs, 95), errs.max()))
    PARITY_REPORT["cube"] = dict(throws=thrown, cube_floor_points=cube_floor, cube_robot_points=cube_hits, manifold_exact_match=n_match / n_samples,
                                 state_err_median=float(np.median(errs)), state_err_p95=float(np.percentile(errs, 95)), state_err_max=float(errs.max()))
    _write_parity_report()
    assert thrown >= n // 2 and cube_floor > 0 and cube_hits > 0, (thrown, cube_floor, cube_hits)
    assert n_match / n_samples >= 0.99
    assert np.median(errs) <= 1e-5 and (errs <= STEP_RTOL).mean() >= 0.95 and errs.max() < 0.5, (np.median(errs), (errs <= STEP_RTOL).mean(), errs.max())


def test_cube_world_physics_only_substeps_match_oracle(oracle_lib):
    """rs_world_substeps on a world with the cube (the drop-in shim's World.step when the reference's own FlagrunHarder class
    drives it, tests/test_cpp_household_shim.py::test_flagrun_harder_with_thrown_cube_on_shim): physics-only sub-steps on the
    static-topology code (k_substeps_static), raw joint torques, no epilogue.  Teacher-forced against oracle_substeps through a
    throw placed by hand (Robot.set_pose_and_speed = 13 state words), the flight, hits on the robot and the floor."""
    from roboschool_b200.policy import load_zoo_weights
    env_id = "RoboschoolHumanoidFlagrunHarder-v1"
    n = 96
    m, w, o = _worlds(env_id, n, 4)
    wts = load_zoo_weights(env_id)
    obs = o.obs()[0]
    for t in range(60):
        obs, _, _ = o.step(oracle_lib.mlp_forward(wts, obs), auto_reset=True, seed=4, threads=8)
    S, nman = m.state_dim, m.n_geoms + m.n_pairs
    rng = np.random.RandomState(3)
    errs, n_match, n_samples, cube_floor, cube_hits = [], 0, 0, 0, 0
    for t in range(36):
        so = o.get_state()
        if t == 2:      # throw: 4 m from the torso, 1 m above it, 25 m/s at it
            for i in range(n):
                ang = rng.uniform(-3.14, 3.14)
                cp = so[i, 0:3] + np.array([4 * np.cos(ang), 4 * np.sin(ang), 1.0])
                v = so[i, 0:3] - cp
                so[i, S - 13:S - 10] = cp
                so[i, S - 10:S - 6] = [0, 0, 0, 1]
                so[i, S - 6:S - 3] = 0
                so[i, S - 3:S] = 25.0 * v / np.linalg.norm(v)
        so32 = so.astype(np.float32)
        w.set_state(so32, clear_contacts=False)
        w.set_contacts(*o.contact_cache(max_pts=16))
        tau = rng.uniform(-15, 15, (n, m.n_dof)).astype(np.float32)
        for i in range(n):
            o.set_state(i, so32[i].astype(np.float64), clear_contacts=False)
            o.set_tau(i, tau[i].astype(np.float64))
            o.substeps(i, 4)
        w.substeps(4, tau)
        sg, so2 = w.get_state(), o.get_state()
        cnt_o, keys_o, _ = o.contact_cache(max_pts=16)
        for i in range(n):
            kg, _ = w.contacts(i)
            n_samples += 1
            if not (len(kg) == cnt_o[i] and np.array_equal(kg, keys_o[i, :cnt_o[i]])):
                continue
            n_match += 1
            cube_floor += int((kg == nman).sum()); cube_hits += int((kg > nman).sum())
            errs.append(np.abs(sg[i] - so2[i]).max() / max(1.0, np.abs(so2[i]).max()))
    errs = np.array(errs)
    print("\n[cube, physics-only sub-steps] cube-floor points %d, cube-robot points %d, manifolds identical %.4f, state err median %.2e p95 %.2e max %.2e"
          % (cube_floor, cube_hits, n_match / n_samples, np.median(errs), np.percentile(errs, 95), errs.max()))
    assert cube_floor > 0 and cube_hits > 0, (cube_floor, cube_hits)
    assert n_match / n_samples >= 0.99
    assert np.median(errs) <= 1e-5 and (errs <= STEP_RTOL).mean() >= 0.95 and errs.max() < 0.5, (np.median(errs), (errs <= STEP_RTOL).mean(), errs.max())


@pytest.mark.parametrize("env_id", STEP_ENVS)
def test_free_running_bounded_divergence(env_id, oracle_lib):
    """No teacher forcing: fp32 GPU vs fp64 oracle over a short horizon; divergence stays bounded and
    observations/rewards agree early in the rollout."""
    n = 64
    m, w, o = _worlds(env_id, n, 9)
    rng = np.random.RandomState(2)
    for t in range(10):
        a = rng.uniform(-1, 1, (n, m.n_act)).astype(np.float32)
        obs_g, rew_g, done_g = w.step_host(a)
        obs_o, rew_o, done_o = o.step(a.astype(np.float64))
        err = np.abs(obs_g - obs_o).max(axis=1)
        assert np.median(err) < 1e-3, (t, np.median(err))
        assert np.isfinite(obs_g).all()
        close = err < 1e-3                      # envs that have not been split by a clamp flip yet
        assert np.median(np.abs(rew_g - rew_o)[close]) < 5e-3, (t, np.median(np.abs(rew_g - rew_o)[close]))
        assert (done_g.astype(bool) == done_o.astype(bool))[close].mean() >= 0.98
    assert (np.abs(obs_g - obs_o).max(axis=1) < 0.05).mean() > 0.9


def test_pendulum_physics_only(oracle_lib):
    """configs[0] plumbing: InvertedPendulum sub-steps with explicit torques, no contacts."""
    m, w, o = _worlds("RoboschoolInvertedPendulum-v1", 32, 1)
    rng = np.random.RandomState(3)
    s0 = rng.uniform(-0.5, 0.5, (32, 4))
    w.set_state(s0.astype(np.float32))
    for i in range(32):
        o.set_state(i, s0[i].astype(np.float32).astype(np.float64))
    tau = rng.uniform(-10, 10, (32, 2)).astype(np.float32)
    tau[:, 1] = 0
    for i in range(32):
        o.set_tau(i, tau[i])
        o.substeps(i, 20)
    w.substeps(20, tau)
    assert np.abs(w.get_state() - o.get_state()).max() < 2e-4


@pytest.mark.parametrize("env_id", ["RoboschoolInvertedPendulum-v1", "RoboschoolHopper-v1", "RoboschoolHumanoid-v1"])
def test_joint_motor_rows_match_oracle(env_id, oracle_lib):
    """rs_world_set_joint_motor (btMultiBodyJointMotor rows behind Joint::set_servo_target / set_target_speed): PD servos with
    per-env targets on every joint and a velocity motor on the first one, physics-only sub-steps, teacher-forced against the
    oracle; and the full env step of a world with motors (runtime-topology kernel) against the oracle's."""
    m, w, o = _worlds(env_id, 16, 2)
    n, nd, dt = 16, m.n_dof, m.timestep
    rs = np.random.RandomState(7)
    for e in range(n):
        for d in range(nd):
            args = (0.1, 0.4, float(rs.uniform(-0.3, 0.3)), 0.0, float(rs.choice([0.05, 2.0, 50.0])) * dt)
            o.set_joint_motor(e, d, *args); w.set_joint_motor(d, *args, env=e)
    args = (0.0, 0.8, 0.0, 0.5, 20.0 * dt)
    o.set_joint_motor(-1, 0, *args); w.set_joint_motor(0, *args)           # env -1: every env
    errs = []
    for t in range(40):
        w.set_state(o.get_state().astype(np.float32), clear_contacts=False)
        for e in range(n):
            o.substeps(e, 1)
        w.substeps(1)
        so, sw = o.get_state(), w.get_state()
        errs.append(np.abs(so - sw).max(axis=1) / np.maximum(1.0, np.abs(so).max(axis=1)))
    errs = np.concatenate(errs)
    assert np.median(errs) <= 1e-5 and np.mean(errs <= 1e-4) >= 0.95 and errs.max() <= 0.2, (np.median(errs), errs.max())
    # full env step (actions ignored by joints whose motor is on? no: torques still apply, the motor rows come on top)
    import torch
    A, O = m.n_act, m.obs_dim
    act = rs.uniform(-1, 1, (n, A))
    w.set_state(o.get_state().astype(np.float32), clear_contacts=False)
    obs = torch.empty((n, O), device="cuda"); rew = torch.empty(n, device="cuda"); done = torch.empty(n, dtype=torch.uint8, device="cuda")
    w.step(torch.tensor(act, dtype=torch.float32, device="cuda"), obs, rew, done, auto_reset=False)
    torch.cuda.synchronize()
    oo, orw, od = o.step(act)
    assert np.abs(obs.cpu().numpy() - oo).max() < 5e-3


def test_link_poses_match(oracle_lib):
    m, w, o = _worlds("RoboschoolHumanoid-v1", 16, 4)
    lp = w.link_poses()
    for i in range(16):
        ref = o.link_poses(i)
        assert np.abs(lp[i][:, :3] - ref[:, :3]).max() < 1e-5
        # quaternions up to sign
        d = np.minimum(np.abs(lp[i][:, 3:7] - ref[:, 3:7]).max(axis=1), np.abs(lp[i][:, 3:7] + ref[:, 3:7]).max(axis=1))
        assert d.max() < 1e-5


def test_policy_known_answers():
    """MLP against the known answers computed from the shipped .weights."""
    import torch
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    from oracle.oracle import mlp_forward
    for env_id, O, A in [("RoboschoolHopper-v1", 15, 3), ("RoboschoolAnt-v1", 28, 8), ("RoboschoolHumanoid-v1", 44, 17)]:
        wts = load_zoo_weights(env_id)
        pol = ZooPolicy(wts, 0)
        X = np.random.RandomState(12345).uniform(-1, 1, (4, O))
        Xb = np.concatenate([np.zeros((1, O)), X, np.random.RandomState(7).uniform(-5, 5, (1000, O))]).astype(np.float32)
        x = torch.tensor(Xb, device="cuda")
        act = torch.empty((Xb.shape[0], A), device="cuda")
        pol.forward(x, act)
        torch.cuda.synchronize()
        ref = mlp_forward(wts, Xb.astype(np.float64))
        got = act.cpu().numpy()
        # fp16 operands (11-bit significand), fp32 accumulate: tolerance 5e-3 of the output scale
        assert np.abs(got - ref).max() <= 5e-3 * max(1.0, np.abs(ref).max()), env_id


def test_vec_env_full_size_properties(monkeypatch):
    """BASELINE-size run (32768 Humanoid envs): finite, deterministic, independent of batch position."""
    import torch
    from roboschool_b200.vec_env import VecEnv
    n = 32768
    env = VecEnv("RoboschoolHumanoid-v1", n, 0)
    obs = env.reset(seed=1).clone()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.rand((n, 17), device="cuda", generator=g) * 2 - 1 for _ in range(5)]
    outs = []
    for a in acts:
        o, r, d = env.step(a)
        outs.append((o.clone(), r.clone(), d.clone()))
    assert all(torch.isfinite(o).all() and torch.isfinite(r).all() for o, r, d in outs)
    # determinism: same seed, same actions -> bit-identical
    env2 = VecEnv("RoboschoolHumanoid-v1", n, 0)
    obs2 = env2.reset(seed=1)
    assert torch.equal(obs, obs2)
    for a, (o, r, d) in zip(acts, outs):
        o2, r2, d2 = env2.step(a)
        assert torch.equal(o, o2) and torch.equal(r, r2) and torch.equal(d, d2)
    # sharding invariance: envs [1000, 1256) stepped in a small world with env_id0=1000 give the same results (same compiled
    # kernel on both sides -- 7-warp CTAs -- since bit-identity cannot be asked of two different instantiations)
    monkeypatch.setenv("RS_WPC", "7")
    env3 = VecEnv("RoboschoolHumanoid-v1", 256, 0, env_id0=1000)
    o3 = env3.reset(seed=1)
    assert torch.equal(o3, obs[1000:1256])
    for a, (o, r, d) in zip(acts, outs):
        o3, r3, d3 = env3.step(a[1000:1256].contiguous())
        assert torch.equal(o3, o[1000:1256]) and torch.equal(r3, r[1000:1256])


def test_rollout_record_matches_stepping():
    """Rollout recorder (time-major tensors for trainers) == the same steps taken one by one through VecEnv.step
    with the zoo policy: bit-identical observations, actions, rewards, dones, across an auto-reset."""
    import torch
    from roboschool_b200.vec_env import VecEnv
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    env_id, n, T = "RoboschoolHopper-v1", 512, 48      # Hopper under its zoo policy falls within ~30 steps here: resets are covered
    pol = ZooPolicy(load_zoo_weights(env_id), 0)
    a_env = VecEnv(env_id, n, 0)
    a_env.reset(seed=4)
    obs_T, act_T, rew_T, done_T = a_env.rollout(pol, T)
    torch.cuda.synchronize()
    b_env = VecEnv(env_id, n, 0)
    obs = b_env.reset(seed=4)
    assert torch.equal(obs_T[0], obs)
    act = torch.empty((n, b_env.act_dim), device="cuda")
    for t in range(T):
        pol.forward(obs, act)
        assert torch.equal(act, act_T[t])
        obs, rew, done = b_env.step(act)
        assert torch.equal(obs, obs_T[t + 1]) and torch.equal(rew, rew_T[t]) and torch.equal(done, done_T[t])
    assert done_T.sum().item() > 0
    # replaying the recorded actions without a policy reproduces the rollout
    c_env = VecEnv(env_id, n, 0)
    c_env.reset(seed=4)
    o2, a2, r2, d2 = c_env.rollout(None, T, actions=act_T)
    assert torch.equal(o2, obs_T) and torch.equal(r2, rew_T) and torch.equal(d2, done_T)


@pytest.mark.parametrize("env_id", ["RoboschoolHumanoid-v1", "RoboschoolAnt-v1", "RoboschoolHopper-v1"])
def test_ragged_batch_sizes(env_id):
    """Ragged batches: N = 1, N not a multiple of the warp (32) or of the CTA (7 / 4 warps), tail lanes that own no
    env but still help build the warp's constraint rows.  Each small world must reproduce, bit for bit, the matching
    slice of one big world (RNG and physics are keyed by the global env id, never by the position in the batch).
    (7-warp CTAs against 1-warp CTAs: test_baseline_config_sizes_properties and the pipelined host-path test.)"""
    import torch
    from roboschool_b200.vec_env import VecEnv
    big_n = 1200
    big = VecEnv(env_id, big_n, 0)
    obs_b = big.reset(seed=7).clone()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [torch.rand((big_n, big.act_dim), device="cuda", generator=g) * 2 - 1 for _ in range(12)]
    outs = []
    for a in acts:
        o, r, d = big.step(a)
        outs.append((o.clone(), r.clone(), d.clone()))
    for n, off in [(1, 0), (1, 1199), (31, 5), (33, 64), (225, 700), (449, 751)]:
        small = VecEnv(env_id, n, 0, env_id0=off)
        o0 = small.reset(seed=7)
        assert torch.equal(o0, obs_b[off:off + n]), (n, off)
        for a, (o, r, d) in zip(acts, outs):
            o2, r2, d2 = small.step(a[off:off + n].contiguous())
            assert torch.equal(o2, o[off:off + n]) and torch.equal(r2, r[off:off + n]) and torch.equal(d2, d[off:off + n]), (n, off)
        small.close()


def test_behavioural_fixture_on_gpu_humanoid_walks():
    """The behavioural pin of the oracle (tests/test_oracle_cpu.py) on the CUDA path, at scale: the reference's pretrained
    Humanoid policy -- trained on the Bullet path -- walks on the fp32 batched stepper: over 1000 steps of 4096 envs the
    mean reward per step is above reward_threshold/1000 = 3.5 (roboschool/__init__.py:71-77), nearly every episode reaches
    the 1000-step time limit, and nothing goes non-finite."""
    import torch
    from roboschool_b200.world import BatchedWorld
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    env_id, n = "RoboschoolHumanoid-v1", 4096
    W = BatchedWorld(load_env_model(env_id), n, 0)
    pol = ZooPolicy(load_zoo_weights(env_id), 0)
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    W.rollout_reset(seed=11, stream=st)
    W.rollout(pol, 1000, seed=11, stats=stats, stream=st)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    assert s[3] == 0, "non-finite observations"
    assert s[0] / s[6] > 3.5, s[0] / s[6]
    assert s[1] >= 0.97 * n and s[2] / s[1] > 990, (s[1], s[2] / max(s[1], 1))


@pytest.mark.parametrize("env_id,n", [("RoboschoolHumanoid-v1", 25000), ("RoboschoolHopper-v1", 9000), ("RoboschoolAnt-v1", 33152)])
def test_policy_step_host_pipelined_matches_single_stream(env_id, n, monkeypatch):
    """rs_world_policy_step_host with page-locked buffers cuts the envs into CTA-aligned slices on separate streams (copies of
    one slice overlap the kernels of the others; from the second use of a buffer set on, the slices of a step are replayed as one
    CUDA graph).  Bit-identical to the single-stream form (RS_E2E_CHUNKS=1) and to pageable
    buffers (staged path), over steps that include auto-resets; n is ragged on purpose (partial tail CTA, partial tail slice)."""
    import torch
    from roboschool_b200.world import BatchedWorld
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    m = load_env_model(env_id)
    pol = ZooPolicy(load_zoo_weights(env_id), 0)
    st = torch.cuda.current_stream().cuda_stream
    def run(chunks, pinned, graph=True):
        monkeypatch.setenv("RS_E2E_CHUNKS", str(chunks))
        monkeypatch.setenv("RS_E2E_GRAPH", "1" if graph else "0")   # slices replayed as one CUDA graph per buffer set / launched directly
        W = BatchedWorld(m, n, 0)
        obs0 = torch.empty((n, m.obs_dim), dtype=torch.float32, device="cuda")
        W.reset(seed=9, obs=obs0, stream=st)
        torch.cuda.synchronize()
        if pinned:
            a, b = W.pinned_host_buffers(), W.pinned_host_buffers()
        else:
            a = (np.empty((n, m.obs_dim), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8))
            b = (np.empty((n, m.obs_dim), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8))
        a[0][:] = obs0.cpu().numpy()
        outs = []
        cur, nxt = a, b
        for t in range(40):
            W.policy_step_host(pol, cur[0], out=nxt)
            outs.append((nxt[0].copy(), nxt[1].copy(), nxt[2].copy()))
            cur, nxt = nxt, cur
        return outs
    ref = run(1, True)
    if env_id == "RoboschoolHopper-v1":        # its zoo policy falls within ~30 steps on the restated physics: auto-resets are covered
        assert sum(int(d.sum()) for _, _, d in ref) > 0
    for chunks, pinned, graph in [(6, True, True), (6, True, False), (4, True, True), (3, True, False), (4, False, True)]:
        got = run(chunks, pinned, graph)
        for (o, r, d), (o2, r2, d2) in zip(ref, got):
            assert np.array_equal(o, o2) and np.array_equal(r, r2) and np.array_equal(d, d2), (chunks, pinned, graph)


@pytest.mark.parametrize("env_id,min_return,min_len", [
    ("RoboschoolInvertedPendulum-v1", 950.0, 990), ("RoboschoolInvertedDoublePendulum-v1", 7000.0, 750),
    ("RoboschoolReacher-v1", 15.0, 150), ("RoboschoolHopper-v1", 2000.0, 950)])
def test_behavioural_fixture_on_gpu_v0_policies(env_id, min_return, min_len):
    """The second behavioural pin (tests/test_oracle_cpu.py::test_behavioural_fixture_v0_policies) on the CUDA path:
    the reference's independently trained May-2017 policies over one full episode of 2048 envs.  Mean return per
    finished episode against the registered reward_threshold (roboschool/__init__.py:11-47: 950 / 9100 / 18 / 2500);
    the double pendulum and the hopper are held to what the oracle reaches (a quarter of the double-pendulum starts
    fall early, the hopper runs the full episode at ~2250)."""
    import torch
    from roboschool_b200.world import BatchedWorld
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    m = load_env_model(env_id)
    n, T = 2048, m.max_episode_steps
    W = BatchedWorld(m, n, 0)
    pol = ZooPolicy(load_zoo_weights(env_id, "v0"), 0)
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    W.rollout_reset(seed=5, stream=st)
    W.rollout(pol, T, seed=5, stats=stats, stream=st)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    assert s[3] == 0, "non-finite observations"
    assert s[1] >= n, s[1]
    # return per episode: reward summed over all env-steps / episodes' share of them
    ret = s[0] / s[6] * (s[2] / s[1])
    assert s[2] / s[1] >= min_len, s[2] / s[1]
    assert ret >= min_return, ret


@pytest.mark.parametrize("env_id,n", [("RoboschoolHopper-v1", 4096), ("RoboschoolAnt-v1", 65536)])
def test_baseline_config_sizes_properties(env_id, n, monkeypatch):
    """BASELINE.json configs[1] (Hopper, 4096 envs) and configs[2] (Ant, 65536 envs) at full size, through the
    size-independent properties: finite outputs, bit-identical reruns, independence of the position in the batch
    (a slice stepped in its own small world with the matching env_id0), contact pools consistent with feet flags."""
    import torch
    from roboschool_b200.vec_env import VecEnv
    env = VecEnv(env_id, n, 0)
    obs0 = env.reset(seed=2).clone()
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = [torch.rand((n, env.act_dim), device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    outs = []
    for a in acts:
        o, r, d = env.step(a)
        outs.append((o.clone(), r.clone(), d.clone()))
    assert all(torch.isfinite(o).all() and torch.isfinite(r).all() for o, r, d in outs)
    env2 = VecEnv(env_id, n, 0)
    assert torch.equal(env2.reset(seed=2), obs0)
    for a, (o, r, d) in zip(acts, outs):
        o2, r2, d2 = env2.step(a)
        assert torch.equal(o, o2) and torch.equal(r, r2) and torch.equal(d, d2)
    off = n - 300
    # bit-identity is a property of ONE compiled kernel: a 300-env world would pick 1-warp CTAs, the full-size one 7-warp
    # CTAs from 24576 envs up, and ptxas contracts multiply-adds differently in the two instantiations (measured on the Ant:
    # observations differ by 1-7 ulp, 1e-7 .. 9e-7, between them; identical with the same CTA shape)
    # (the Ant at 65536 envs runs on 14-warp CTAs: one wave instead of two of 7-warp CTAs)
    if n >= 24576:
        wpc = env.world.L.rs_world_cta_warps(env.world.h)
        assert wpc == (14 if env_id == "RoboschoolAnt-v1" else 7), wpc
        monkeypatch.setenv("RS_WPC", str(wpc))
    env3 = VecEnv(env_id, 300, 0, env_id0=off)
    assert torch.equal(env3.reset(seed=2), obs0[off:])
    for a, (o, r, d) in zip(acts, outs):
        o3, r3, d3 = env3.step(a[off:].contiguous())
        assert torch.equal(o3, o[off:]) and torch.equal(r3, r[off:]) and torch.equal(d3, d[off:])
    # feet_contact observations are 0/1 flags
    F = env.model.n_feet
    feet = outs[-1][0][:, -F:]
    assert ((feet == 0) | (feet == 1)).all()


@pytest.mark.parametrize("variant", ["static-pooled", "generic"])
def test_free_flight_momentum_converges_on_gpu(variant, monkeypatch):
    """Physics invariant on the CUDA kernels themselves (no oracle involved): a floating Humanoid tumbling in zero
    gravity with internal joint motion, no damping / limits / floor.  Linear momentum is an invariant of the continuous
    dynamics, so its drift under the semi-implicit integrator must shrink ~linearly with dt -- a check of the ABA
    sweeps, Coriolis and bias terms of the step kernel (the compile-time-topology one and the runtime-topology twin)."""
    from roboschool_b200.world import BatchedWorld
    drift = []
    n = 64
    for div in (1, 4, 16):
        m = load_env_model("RoboschoolHumanoid-v1").clone()
        m.lin_damping = 0; m.ang_damping = 0; m.gravity = 0.0
        m.timestep = m.timestep / div
        for i in range(m.n_links):
            m.has_limit[i] = 0
        for k in range(m.n_act):
            m.act_gain[k] = 0.0          # actions produce no torque: free motion
        if variant == "generic":         # any model that is not one of the generated topologies runs on the runtime-topology
            m.mass[2] *= 1.01            # kernel (e.g. everything loaded through load_mjcf): 1 % more mass on one link is enough
        w = BatchedWorld(m, n, 0)
        assert (w.L.rs_world_kernel_variant(w.h) > 0) == (variant == "static-pooled")
        w.reset(seed=2)
        s = w.get_state()
        rng = np.random.RandomState(0)
        s[:, 13 + 17:] = rng.uniform(-3, 3, (n, 17))
        s[:, 7:13] = rng.uniform(-1, 1, (n, 6))
        s[:, 2] = 10.0                   # far above the floor plane (the geoms keep their floor flags: same topology)
        w.set_state(s)
        mass = np.array([m.base_mass] + [m.mass[i] for i in range(m.n_links)])
        p0 = (mass[None, :, None] * w.link_poses()[:, :, 7:10]).sum(1)
        a = np.zeros((n, m.n_act), dtype=np.float32)
        for _ in range(20 * div):
            w.step_host(a, auto_reset=False)
        p1 = (mass[None, :, None] * w.link_poses()[:, :, 7:10]).sum(1)
        drift.append(np.median(np.abs(p1 - p0).max(axis=1)))
        w.close()
    assert drift[1] < drift[0] / 2.5 and drift[2] < drift[1] / 2.0, drift


def test_throughput_floor_on_b200():
    """Guards the measured operating point against silent regressions (a wrong shared-memory carve-out, a fall back to the
    runtime-topology kernel, a lost phase barrier each cost 20-70 %): 32768 Humanoid envs under the zoo policy, device-resident
    rollout, CUDA events.  Measured 38-42 M env-steps/s on the pool's B200s; the floor leaves a wide margin for box-to-box spread."""
    import torch
    from roboschool_b200.world import BatchedWorld
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    if "B200" not in torch.cuda.get_device_name(0):
        pytest.skip("throughput floor is stated for B200")
    env_id, n = "RoboschoolHumanoid-v1", 32768
    W = BatchedWorld(load_env_model(env_id), n, 0)
    assert W.L.rs_world_kernel_variant(W.h) > 0, "static-topology kernel not selected"
    pol = ZooPolicy(load_zoo_weights(env_id), 0)
    st = torch.cuda.current_stream().cuda_stream
    W.rollout_reset(seed=1, stream=st)
    W.rollout(pol, 200, seed=1, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    W.rollout(pol, 100, seed=1, stream=st)
    e1.record()
    torch.cuda.synchronize()
    rate = n * 100 / (e0.elapsed_time(e1) * 1e-3)
    assert rate > 2.5e7, "%.1f M env-steps/s" % (rate / 1e6)


@pytest.mark.parametrize("env_id,n,wpc", [("RoboschoolHumanoid-v1", 32768, 7), ("RoboschoolHumanoid-v1", 37888, 8), ("RoboschoolAnt-v1", 65536, 14)])
def test_wide_cta_world_oracle_spot_check(env_id, n, wpc, oracle_lib):
    """Direct oracle check of the wide-CTA step kernels: a 32768-env Humanoid world (the bench
    operating point: `k_step_static<TopoHumanoid,16,0,7>`), a 37888-env one (8-warp CTAs) and a 65536-env Ant world (one wave of
    14-warp CTAs, row discard on) free-run 6 steps of random actions; 8 blocks of 32 envs spread over the batch (first / interior /
    last CTA, CTA-straddling offsets) are compared with oracle worlds that were reset with the matching global env ids."""
    from roboschool_b200.world import BatchedWorld
    from oracle.oracle import OracleWorld
    m = load_env_model(env_id)
    w = BatchedWorld(m, n, 0)
    assert w.L.rs_world_cta_warps(w.h) == wpc
    w.reset(seed=13)
    per = 32 * wpc
    offs = [0, 200, per * 3 - 16, 9999, 16384, 20000, n - per, n - 32]
    orc = []
    for off in offs:
        o = OracleWorld(m, 32)
        o.reset(seed=13, env_id0=off)
        orc.append(o)
    s0 = w.get_state()
    for off, o in zip(offs, orc):
        assert np.abs(s0[off:off + 32] - o.get_state()).max() < 1e-6
    rng = np.random.RandomState(4)
    for t in range(6):
        a = rng.uniform(-1, 1, (n, m.n_act)).astype(np.float32)
        obs_g, rew_g, done_g = w.step_host(a)
        errs = []
        for off, o in zip(offs, orc):
            obs_o, rew_o, done_o = o.step(a[off:off + 32].astype(np.float64))
            e = np.abs(obs_g[off:off + 32] - obs_o).max(axis=1)
            errs.append(e)
            close = e < 1e-3
            assert np.median(np.abs(rew_g[off:off + 32] - rew_o)[close]) < 5e-3
            assert (done_g[off:off + 32].astype(bool) == done_o.astype(bool))[close].all()
        errs = np.concatenate(errs)
        assert np.median(errs) < 1e-3, (t, np.median(errs))
    assert (errs < 0.05).mean() > 0.9


_SHIM_ROBOT_XML = """<mujoco model="shimtest">
  <compiler angle="degree" coordinate="local" inertiafromgeom="true"/>
  <default><joint limited="true"/><geom conaffinity="1" condim="3" contype="1" friction="0.8 .1 .1"/></default>
  <worldbody>
    <body name="torso" pos="0 0 0.9">
      <geom fromto="0 0 0.2 0 0 -0.2" name="torso_geom" size="0.06" type="capsule"/>
      <body name="thigh_r" pos="0 -0.1 -0.2">
        <joint axis="0 -1 0" name="hip_r" pos="0 0 0" range="-100 30" type="hinge"/>
        <geom fromto="0 0 0 0 0 -0.35" name="thigh_r_geom" size="0.05" type="capsule"/>
        <body name="shin_r" pos="0 0 -0.35">
          <joint axis="0 -1 0" name="knee_r" pos="0 0 0" range="-150 0" type="hinge"/>
          <geom fromto="0 0 0 0 0 -0.3" name="shin_r_geom" size="0.04" type="capsule"/>
          <geom name="foot_r_geom" pos="0.05 0 -0.3" size="0.06" type="sphere"/>
        </body>
      </body>
      <body name="thigh_l" pos="0 0.1 -0.2">
        <joint axis="1 0 0" name="hip_l_x" pos="0 0 0" range="-25 25" type="hinge"/>
        <joint axis="0 -1 0" name="hip_l" pos="0 0 0" range="-100 30" type="hinge"/>
        <geom fromto="0 0 0 0 0 -0.35" name="thigh_l_geom" size="0.05" type="capsule"/>
        <body name="shin_l" pos="0 0 -0.35">
          <joint axis="0 -1 0" name="knee_l" pos="0 0 0" range="-150 0" type="hinge"/>
          <geom fromto="0 0 0 0 0 -0.3" name="shin_l_geom" size="0.04" type="capsule"/>
        </body>
      </body>
    </body>
  </worldbody>
</mujoco>
"""
_SHIM_FLOOR_XML = """<mujoco model="ground_plane"><worldbody>
  <geom conaffinity="1" condim="3" name="floor" friction="0.8 0.1 0.1" pos="0 0 0" type="plane"/>
</worldbody></mujoco>
"""


def test_cpp_household_shim_on_cuda_backend(tmp_path, monkeypatch, oracle_lib):
    """The drop-in `cpp_household` module on the CUDA backend: the call sequence a reference env
    makes -- World(g, dt), load_mjcf(floor), load_mjcf(robot), Joint.reset_current_position, Robot.set_pose_and_speed,
    set_motor_torque, World.step(frame_skip), current_relative_position, Thingy.pose / speed / contact_list, and the motor
    modes -- runs once through rs_world_substeps / rs_world_link_poses / rs_world_get_contacts on the GPU and once with
    the shim's backend swapped for the oracle; every value returned to the "env" must agree.  (The reference's own env
    classes run over the same shim in tests/test_cpp_household_shim.py where /root/reference exists; the GPU box has no
    reference tree, so this test brings its own small MJCF: floating base, a two-joint body, self-collision, a floor.)"""
    import roboschool_b200.cpp_household as shim
    from tests.oracle_backend import OracleBackend
    robot_xml, floor_xml = tmp_path / "robot.xml", tmp_path / "ground_plane.xml"
    robot_xml.write_text(_SHIM_ROBOT_XML); floor_xml.write_text(_SHIM_FLOOR_XML)

    def drive(use_oracle):
        shim.World._model_cache.clear()
        if use_oracle:
            monkeypatch.setattr(shim, "_make_backend", lambda model, n_envs=1, device=0: OracleBackend(model, n_envs))
        else:
            monkeypatch.undo()
        world = shim.World(9.8, 0.0165 / 4)
        floor = world.load_mjcf(str(floor_xml))
        robots = world.load_mjcf(str(robot_xml))
        assert len(floor) == 1 and len(robots) == 1
        r = robots[0]
        assert [j.name for j in r.joints] == ["hip_r", "knee_r", "hip_l_x", "hip_l", "knee_l"]
        rng = np.random.RandomState(7)
        for j in r.joints:
            j.reset_current_position(float(rng.uniform(-0.1, 0.1)), 0.0)
        p = shim.Pose(); p.set_xyz(0.0, 0.0, 0.95); p.set_rpy(0.0, 0.05, 0.3)
        r.set_pose_and_speed(p, 0.2, 0.0, 0.0)
        r.query_position()
        log = []
        parts = {t.name: t for t in r.parts}
        for t in range(40):
            if t == 25:                       # motor modes half way: a servo on one knee, a velocity motor on a hip
                r.joints[1].set_servo_target(-0.8, 0.1, 0.4, 60.0)
                r.joints[3].set_target_speed(-1.0, 0.5, 100.0)
            for k, j in enumerate(r.joints):
                if t >= 25 and k in (1, 3):
                    continue
                j.set_motor_torque(float(np.float32(6.0 * np.sin(0.3 * t + k))))
            assert world.step(4) is False
            row = []
            for j in r.joints:
                row += list(j.current_relative_position()) + list(j.current_position())
            for t_ in [r.root_part] + r.parts:
                ps = t_.pose()
                row += list(ps.xyz()) + list(ps.rpy()) + list(t_.speed())
            contacts = tuple(sorted(c.name for c in parts["shin_r"].contact_list())) + ("|",) + \
                       tuple(sorted(c.name for c in parts["shin_l"].contact_list()))
            log.append((np.array(row, dtype=np.float64), contacts))
        world.clean_everything()
        return log

    got = drive(False)
    want = drive(True)
    n_contact_same = 0
    for t, ((rg, cg), (ro, co)) in enumerate(zip(got, want)):
        assert np.isfinite(rg).all()
        err = np.abs(rg - ro).max()
        assert err < (5e-3 if t < 25 else 0.1), (t, err)       # free-running fp32 vs fp64 over 40 env steps (contacts from ~t=26)
        n_contact_same += int(cg == co)
    assert any("floor" in c for _, c in got), "the robot never touched the floor: the contact_list path was not exercised"
    assert n_contact_same >= 30, n_contact_same
