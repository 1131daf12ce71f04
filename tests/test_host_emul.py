"""The fp32 per-env kernel code (roboschool_b200/csrc/rs_core.cuh) compiled for the host, against the
fp64 oracle.  This is what lets the GPU-less container check the kernel arithmetic; the same
comparison runs on the real GPU in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from roboschool_b200 import envspec


@pytest.fixture(scope="module")
def emul():
    from tests.host_emul import emul as e
    e.build()
    return e


@pytest.mark.parametrize("static", [False, True], ids=["generic", "static-pooled"])
@pytest.mark.parametrize("env_id", ["RoboschoolInvertedPendulum-v1", "RoboschoolInvertedDoublePendulum-v1", "RoboschoolReacher-v1", "RoboschoolHopper-v1", "RoboschoolWalker2d-v1", "RoboschoolHalfCheetah-v1", "RoboschoolAnt-v1",
                                    "RoboschoolHumanoid-v1", "RoboschoolHumanoidFlagrun-v1", "RoboschoolHumanoidFlagrunHarder-v1"])
def test_emul_teacher_forced(env_id, static, emul, oracle_lib):
    """static=True runs the compile-time-topology sweeps with the pooled row builder (rs_static.cuh,
    rs_pooled.cuh) -- the code the shipped models run on the GPU; False the runtime-topology twin."""
    m = envspec.load_env_model(env_id)
    n = 32
    o = oracle_lib.OracleWorld(m, n)
    e = emul.EmulWorld(m, n)
    if static:
        if m.task_kind != 0:
            pytest.skip("pendulums run on the runtime-topology kernel")
        assert e.use_static(True) > 0
    elif m.has_cube:
        pytest.skip("the cube is stepped by its static-topology kernel only")
    o.reset(seed=5); e.reset(seed=5)
    assert np.abs(np.stack([e.get_state(i) for i in range(n)]) - o.get_state()).max() < 1e-6
    rng = np.random.RandomState(1)
    errs, alive = [], np.ones(n, bool)
    for t in range(25):
        a = rng.uniform(-1, 1, (n, m.n_act)).astype(np.float32).astype(np.float64)
        so = o.get_state()
        for i in range(n):
            s32 = so[i].astype(np.float32).astype(np.float64)
            o.set_state(i, s32, clear_contacts=False); e.set_state(i, s32, clear_contacts=False)
            ep = o.get_episode(i); e.set_episode(i, ep[0], ep[1], ep[2], ep[3])
            if m.flag_task:
                e.set_flag(i, *o.get_flag(i))
            if m.flag_task == 2:
                e.set_task_state(i, o.get_task_state(i))
        oo, orw, od = o.step(a)
        eo, erw, ed = e.step(a)
        for i in range(n):
            if not alive[i]:
                continue
            co, ce = o.contacts(i), e.contacts(i)
            if len(co) != len(ce) or not np.array_equal(co[:, :2], ce[:, :2]):
                alive[i] = False
                continue
            s1, s2 = o.get_state(i), e.get_state(i)
            r = np.abs(s1 - s2).max() / max(1.0, np.abs(s1).max())
            errs.append(r)
            if r < 1e-4:
                assert np.abs(oo[i] - eo[i]).max() < 2e-3
                assert abs(orw[i] - erw[i]) < 5e-2
                if m.flag_task:
                    fo, fe = o.get_flag(i), e.get_flag(i)
                    assert fo[2:] == fe[2:] and abs(fo[0] - fe[0]) < 1e-9 and abs(fo[1] - fe[1]) < 1e-9
                if m.flag_task == 2:
                    to, te = o.get_task_state(i), e.get_task_state(i)
                    assert (to[[0, 1, 2, 5, 6, 7]] == te[[0, 1, 2, 5, 6, 7]]).all() and np.abs(to[3:5] - te[3:5]).max() < 1e-2
    errs = np.array(errs)
    assert alive.mean() >= 0.9
    assert np.median(errs) <= 1e-5 and (errs <= 1e-4).mean() >= 0.95 and errs.max() < 0.2


def test_emul_cube_throw_flight_and_hits(emul, oracle_lib):
    """FlagrunHarder's cube (gym_humanoid_flagrun.py:81-118): episodes advanced to frame 110 under the zoo policy (free running
    on the oracle), then 45 teacher-forced steps that cover the throw at frame 120, the flight (drag, gravity), the hit on the
    robot (box-vs-capsule/sphere contact rows between two bodies) and the cube coming to rest on the floor (box-plane manifold)."""
    from roboschool_b200.policy import load_zoo_weights
    m = envspec.load_env_model("RoboschoolHumanoidFlagrunHarder-v1")
    assert m.has_cube == 1
    n = 24
    o = oracle_lib.OracleWorld(m, n)
    e = emul.EmulWorld(m, n)
    assert e.use_static(True) == 6
    o.reset(seed=2); e.reset(seed=2)
    wts = load_zoo_weights("RoboschoolHumanoidFlagrunHarder-v1")
    obs = o.obs()[0]
    for t in range(110):
        obs, _, _ = o.step(oracle_lib.mlp_forward(wts, obs), auto_reset=True, seed=2)
    errs, alive, thrown, cube_rows, hit = [], np.ones(n, bool), 0, 0, 0
    S = m.state_dim
    for t in range(45):
        a = oracle_lib.mlp_forward(wts, obs).astype(np.float32).astype(np.float64)
        so = o.get_state()
        for i in range(n):
            s32 = so[i].astype(np.float32).astype(np.float64)
            o.set_state(i, s32, clear_contacts=False); e.set_state(i, s32, clear_contacts=False)
            ep = o.get_episode(i); e.set_episode(i, ep[0], ep[1], ep[2], ep[3])
            e.set_flag(i, *o.get_flag(i)); e.set_task_state(i, o.get_task_state(i))
        if t == 0:     # the emulator has not followed the free-running prefix: adopt the oracle's contact cache
            cnt, keys, data = o.contact_cache()
            for i in range(n):
                e.set_contacts(i, cnt[i], keys[i], data[i])
        cube_before = so[:, S - 13:S - 10].copy()
        obs, orw, od = o.step(a)
        eo, erw, ed = e.step(a)
        s_o = o.get_state()
        thrown += int((np.abs(s_o[:, S - 13:S - 10] - cube_before).max(axis=1) > 2.0).sum())      # teleported to the throw position
        for i in range(n):
            if not alive[i]:
                continue
            co, ce = o.contacts(i), e.contacts(i)
            if len(co) != len(ce) or not np.array_equal(co[:, :2], ce[:, :2]):
                alive[i] = False
                continue
            if len(co):
                nman = m.n_geoms + m.n_pairs
                cube_rows += int((co[:, 0] == nman).sum()); hit += int((co[:, 0] > nman).sum())
            s1, s2 = s_o[i], e.get_state(i)
            errs.append(np.abs(s1 - s2).max() / max(1.0, np.abs(s1).max()))
            if errs[-1] < 1e-4:
                assert abs(orw[i] - erw[i]) < 5e-2 and od[i] == ed[i]
    errs = np.array(errs)
    assert thrown >= n // 2, thrown                    # robots lying on the ground are not pelted
    assert cube_rows > 0 and hit > 0, (cube_rows, hit)
    assert alive.mean() >= 0.8, alive.mean()
    assert np.median(errs) <= 1e-5 and (errs <= 1e-4).mean() >= 0.9 and errs.max() < 0.5, (np.median(errs), (errs <= 1e-4).mean(), errs.max())


def test_emul_auto_reset_matches_oracle(emul, oracle_lib):
    """done -> on-device reset path: same counter-based RNG, same first observation of the new episode."""
    m = envspec.load_env_model("RoboschoolHopper-v1")
    n = 16
    o = oracle_lib.OracleWorld(m, n); e = emul.EmulWorld(m, n)
    o.reset(seed=9, env_id0=100); e.reset(seed=9, env_id0=100)
    rng = np.random.RandomState(4)
    n_done = 0
    for t in range(60):
        a = rng.uniform(-1, 1, (n, m.n_act)).astype(np.float32).astype(np.float64)
        oo, orw, od = o.step(a, auto_reset=True, seed=9, env_id0=100)
        eo, erw, ed = e.step(a, auto_reset=True, seed=9, env_id0=100)
        both = (od == 1) & (ed == 1)
        n_done += both.sum()
        # the first observation after a reset depends only on the RNG: must agree to float32 rounding
        if both.any():
            assert np.abs(oo[both] - eo[both]).max() < 1e-6
    assert n_done > 0


@pytest.mark.parametrize("env_id", ["RoboschoolInvertedPendulum-v1", "RoboschoolHopper-v1", "RoboschoolHumanoid-v1"])
def test_emul_joint_motor_rows_match_oracle(env_id, emul, oracle_lib):
    """Motor rows in the fp32 kernel code against the oracle: PD servos on every joint (impulse bounds between saturating
    and generous) plus a velocity motor on the first joint, 60 sub-steps from the reset state, teacher-forced."""
    m = envspec.load_env_model(env_id)
    o = oracle_lib.OracleWorld(m, 1); o.reset(seed=3)
    e = emul.EmulWorld(m, 1)
    e.set_state(0, o.get_state(0))
    rs = np.random.RandomState(5)
    dt = m.timestep
    for d in range(m.n_dof):
        args = (0.1, 0.4, float(rs.uniform(-0.3, 0.3)), 0.0, float(rs.choice([0.05, 2.0, 50.0])) * dt)
        o.set_joint_motor(0, d, *args); e.set_joint_motor(0, d, *args)
    args = (0.0, 0.8, 0.0, 0.5, 20.0 * dt)
    o.set_joint_motor(0, 0, *args); e.set_joint_motor(0, 0, *args)
    errs = []
    for t in range(60):
        e.set_state(0, o.get_state(0), clear_contacts=False)
        o.substeps(0, 1); e.substeps(0, 1)
        so, se = o.get_state(0), e.get_state(0)
        errs.append(np.abs(so - se).max() / max(1.0, np.abs(so).max()))
    errs = np.array(errs)
    assert np.median(errs) <= 1e-5 and errs.max() <= 2e-3, (np.median(errs), errs.max())
