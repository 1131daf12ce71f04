"""CPU tests of the oracle: closed-form physics, the reference-Python golden vectors for the
epilogue, the MLP known answers, and the behavioural fixture (zoo policy walks)."""
import os

import numpy as np
import pytest

from roboschool_b200 import envspec
from roboschool_b200.policy import load_zoo_weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_free_fall_closed_form(oracle_lib):
    m = envspec.load_env_model("RoboschoolAnt-v1").clone()
    m.lin_damping = 0; m.ang_damping = 0
    w = oracle_lib.OracleWorld(m, 1)
    w.reset(seed=1)
    z0 = w.get_state(0)[2]
    n = 20
    w.substeps(0, n)
    s = w.get_state(0)
    t, h = n * m.timestep, m.timestep
    # semi-implicit Euler: v_n = -g t ; z_n = z0 - g h^2 n(n+1)/2
    assert abs(s[12] + m.gravity * t) < 1e-12
    assert abs(s[2] - (z0 - m.gravity * h * h * n * (n + 1) / 2)) < 1e-12
    assert np.abs(s[7:10]).max() < 1e-12   # no spurious rotation


def test_pendulum_energy_first_order(oracle_lib):
    """Energy error of the undamped cart-pole shrinks ~linearly with dt (semi-implicit Euler)."""
    errs = []
    for div in (1, 4, 16):
        m = envspec.load_env_model("RoboschoolInvertedPendulum-v1").clone()
        m.lin_damping = 0; m.ang_damping = 0; m.timestep = m.timestep / div
        w = oracle_lib.OracleWorld(m, 1)
        w.set_state(0, np.array([0.0, 1.0, 0.0, 0.0]))
        e0 = w.energy(0)
        w.substeps(0, 60 * div)
        errs.append(abs(w.energy(0) - e0))
    assert errs[1] < errs[0] / 2.5 and errs[2] < errs[1] / 2.5


def test_momentum_and_energy_converge_in_free_flight(oracle_lib):
    """Floating 19-link humanoid tumbling with internal joint motion, g=0, no damping/limits/contacts:
    linear momentum and kinetic energy are invariants of the continuous dynamics, so their drift under
    the semi-implicit integrator must shrink ~linearly with dt (checks ABA, Coriolis and bias terms)."""
    dp, dE = [], []
    for div in (1, 4, 16):
        m = envspec.load_env_model("RoboschoolHumanoid-v1").clone()
        m.lin_damping = 0; m.ang_damping = 0; m.gravity = 0.0; m.has_floor = 0; m.n_pairs = 0
        m.timestep = m.timestep / div
        for g in range(m.n_geoms):
            m.g_floor[g] = 0
        for i in range(m.n_links):
            m.has_limit[i] = 0
        w = oracle_lib.OracleWorld(m, 1)
        w.reset(seed=2)
        s = w.get_state(0)
        rng = np.random.RandomState(0)
        s[13 + 17:] = rng.uniform(-3, 3, 17)
        s[7:13] = rng.uniform(-1, 1, 6)
        w.set_state(0, s)
        mass = np.array([m.base_mass] + [m.mass[i] for i in range(m.n_links)])
        p0 = (mass[:, None] * w.link_poses(0)[:, 7:10]).sum(0)
        e0 = w.energy(0)
        w.substeps(0, 80 * div)
        dp.append(np.abs((mass[:, None] * w.link_poses(0)[:, 7:10]).sum(0) - p0).max())
        dE.append(abs(w.energy(0) - e0))
    assert dp[1] < dp[0] / 3 and dp[2] < dp[1] / 3, dp
    assert dE[1] < dE[0] / 3 and dE[2] < dE[1] / 3, dE


def test_determinism_and_thread_independence(oracle_lib):
    m = envspec.load_env_model("RoboschoolHopper-v1")
    a = np.random.RandomState(0).uniform(-1, 1, (30, 8, 3))
    outs = []
    for threads in (1, 4):
        w = oracle_lib.OracleWorld(m, 8)
        w.reset(seed=5)
        for t in range(30):
            obs, rew, done = w.step(a[t], threads=threads)
        outs.append((obs.copy(), w.get_state().copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("env_id", ["RoboschoolInvertedPendulum-v1", "RoboschoolInvertedPendulumSwingup-v1",
                                    "RoboschoolInvertedDoublePendulum-v1", "RoboschoolReacher-v1", "RoboschoolHopper-v1", "RoboschoolWalker2d-v1",
                                    "RoboschoolHalfCheetah-v1", "RoboschoolAnt-v1", "RoboschoolHumanoid-v1", "RoboschoolHumanoidFlagrun-v1",
                                    "RoboschoolHumanoidFlagrunHarder-v1", "RoboschoolHumanoidFlagrunHarder-v1_cube"])
def test_epilogue_matches_reference_python(env_id, oracle_lib):
    """Golden vectors made by the reference's own gym_forward_walker.py (tools/gen_golden_epilogue.py).  `_cube`: FlagrunHarder
    under its zoo policy -- the reference class throws the cube (gym_humanoid_flagrun.py:101-118) and every
    set_pose_and_speed it issued is in the fixture."""
    g = np.load(os.path.join(GOLDEN, "epilogue_%s.npz" % env_id))
    env_id = env_id.replace("_cube", "")
    m = envspec.load_env_model(env_id)
    throws = {int(r[0]): r[1:] for r in g["ref_throws"]} if "ref_throws" in g else {}
    n_throws = 0
    w = oracle_lib.OracleWorld(m, 1)
    w.reset(seed=int(g["seed"]))
    assert np.abs(w.obs()[0][0] - g["ref_obs0"]).max() < 1e-6
    nd = m.n_dof
    row = 0     # index into the recorded frames (one per step, plus one per reset of a multi-episode fixture)
    for t in range(g["actions"].shape[0]):
        obs, rew, done = w.step(g["actions"][t:t + 1])
        s = w.get_state(0)
        row += 1
        if m.has_cube:
            cube, s = s[-13:], s[:-13]
            if row in throws:     # the reference class called aggressive_cube.set_pose_and_speed(...) in this step
                assert np.abs(cube[:3] - throws[row][:3]).max() < 1e-9 and np.abs(cube[10:] - throws[row][3:]).max() < 1e-12
                assert np.abs(cube[3:7] - [0, 0, 0, 1]).max() == 0 and np.abs(cube[7:10]).max() == 0
                n_throws += 1
        # the replayed trajectory must be the one the fixture was recorded from
        assert np.abs(s[-2 * nd:-nd] - g["q"][row]).max() < 1e-9, "fixture stale: rerun tools/gen_golden_epilogue.py"
        assert np.abs(obs[0] - g["ref_obs"][t]).max() < 1e-6
        assert abs(rew[0] - g["ref_rew"][t]) < 1e-6
        assert bool(done[0]) == bool(g["ref_done"][t])
        assert np.abs(w.rewards5(0) - g["ref_rewards5"][t]).max() < 1e-6
        if m.flag_task:   # flag position and timeout as kept by the reference's flag_reposition / calc_state
            tx, ty, to, _ = w.get_flag(0)
            assert abs(tx - g["ref_flags"][t][0]) < 1e-12 and abs(ty - g["ref_flags"][t][1]) < 1e-12
            assert to == int(g["ref_flags"][t][2])
        if m.flag_task == 2:   # FlagrunHarder: the task the reference's curriculum chose for this episode
            assert int(w.get_task_state(0)[0]) == int(g["ref_task"][t])
        if "reset_after" in g and t in g["reset_after"]:   # multi-episode fixture: the reference called env.reset() here
            k = list(g["reset_after"]).index(t)
            w.L.oracle_reset(w.h, 0, int(g["seed"]), 0)
            assert np.abs(w.obs()[0][0] - g["ref_reset_obs"][k]).max() < 1e-6
            row += 1
    if m.flag_task:
        assert len(np.unique(g["ref_flags"][:, 0])) >= 2, "fixture must contain a flag reposition"
    assert n_throws == len(throws)


def test_behavioural_fixture_flagrun_chases_flags(oracle_lib):
    """The reference's pretrained Flagrun policy on the restated physics: weaker pin than the Humanoid one -- the
    policy runs towards the moving flag (reward/step well above the alive bonus alone would give a faller, flags are
    reached before their 150-step timeout) but it does NOT reach the registered reward_threshold 2000
    (roboschool/__init__.py:78-84) in most episodes: turning makes it fall.  Recorded as a gap in DESIGN.md."""
    m = envspec.load_env_model("RoboschoolHumanoidFlagrun-v1")
    wts = load_zoo_weights("RoboschoolHumanoidFlagrun-v1")
    n = 16
    w = oracle_lib.OracleWorld(m, n)
    w.reset(seed=2)
    obs = w.obs()[0]
    tot = np.zeros(n); alive = np.ones(n, bool); reached = np.zeros(n); length = np.zeros(n)
    prev = [w.get_flag(i) for i in range(n)]
    for t in range(1000):
        obs, rew, done = w.step(oracle_lib.mlp_forward(wts, obs), threads=4)
        tot += rew * alive; length += alive
        alive &= ~done.astype(bool)
        for i in range(n):
            f = w.get_flag(i)
            if f[3] != prev[i][3] and prev[i][2] > 1 and alive[i]:
                reached[i] += 1          # moved although the timeout had not run out: the robot got within 1 m
            prev[i] = f
    assert np.median(length) >= 150, length
    assert (tot / length).mean() > 1.0, tot / length
    assert reached.sum() >= 3, reached


def test_mlp_known_answers(oracle_lib):
    """known answers computed from the shipped .weights files."""
    f = oracle_lib.mlp_forward
    w = load_zoo_weights("RoboschoolHopper-v1")
    assert np.allclose(f(w, np.zeros(15)), [-0.1601551671, 0.2929517778, -0.7377750787], atol=1e-9)
    X = np.random.RandomState(12345).uniform(-1, 1, (4, 15))
    assert np.allclose(f(w, X)[0], [0.8471205376, -3.6975331127, 0.4903137045], atol=1e-9)
    assert abs(f(w, X).sum() - 9.62946452367964) < 1e-9
    w = load_zoo_weights("RoboschoolAnt-v1")
    X = np.random.RandomState(12345).uniform(-1, 1, (4, 28))
    assert abs(f(w, X).sum() + 29.475036116473614) < 1e-9
    w = load_zoo_weights("RoboschoolHumanoid-v1")
    assert np.allclose(f(w, np.zeros(44))[:6], [-0.296643717, 1.0570575815, -0.0424037259, 0.3779573418, 0.4688812382, 0.9736019632], atol=1e-9)
    X = np.random.RandomState(12345).uniform(-1, 1, (4, 44))
    assert abs(f(w, X).sum() - 16.448666567345803) < 1e-9


def test_behavioural_fixture_humanoid_walks(oracle_lib):
    """The reference's pretrained Humanoid policy (trained on the Bullet path) must keep walking on the
    restated physics: reward_threshold 3500 per 1000 steps (roboschool/__init__.py:71-77)."""
    m = envspec.load_env_model("RoboschoolHumanoid-v1")
    wts = load_zoo_weights("RoboschoolHumanoid-v1")
    n = 8
    w = oracle_lib.OracleWorld(m, n)
    w.reset(seed=1)
    obs = w.obs()[0]
    tot = np.zeros(n); alive = np.ones(n, bool)
    for t in range(1000):
        obs, rew, done = w.step(oracle_lib.mlp_forward(wts, obs), threads=4)
        tot += rew * alive
        alive &= ~done.astype(bool)
    assert alive.all(), "humanoid fell"
    assert tot.min() > 3500.0, tot


def _score_v0(oracle_lib, env_id, n, seed, T=None):
    m = envspec.load_env_model(env_id)
    wts = load_zoo_weights(env_id, "v0")
    w = oracle_lib.OracleWorld(m, n)
    w.reset(seed=seed)
    obs = w.obs()[0]
    tot = np.zeros(n); alive = np.ones(n, bool); length = np.zeros(n)
    for t in range(T or m.max_episode_steps):
        obs, rew, done = w.step(oracle_lib.mlp_forward(wts, obs), threads=4)
        tot += rew * alive; length += alive
        alive &= ~done.astype(bool)
        if not alive.any():
            break
    return tot, length


def test_behavioural_fixture_v0_policies(oracle_lib):
    """A second, independent behavioural pin: the reference's May-2017 policies (agent_zoo/*_v0_2017may.py, trained on
    the Bullet path months before the v1 ones) on the restated physics, against the registered reward_threshold of each
    env (roboschool/__init__.py:11-47).  The cart-pole, the double pendulum and the reacher reach their thresholds, so the
    fixed-base slide/hinge chains, the contact-free tasks and their epilogues behave like the reference's.  Hopper runs
    the full 1000 steps at ~2250 (threshold 2500): the planar dummy-joint chain, foot contact and friction are close.
    Walker2d and HalfCheetah fall under both generations of policies -- see DESIGN.md "Known gaps"."""
    tot, length = _score_v0(oracle_lib, "RoboschoolInvertedPendulum-v1", 8, 1)
    assert tot.min() >= 950.0, tot
    tot, length = _score_v0(oracle_lib, "RoboschoolInvertedDoublePendulum-v1", 8, 1)
    assert np.median(tot) >= 9100.0, tot
    tot, length = _score_v0(oracle_lib, "RoboschoolReacher-v1", 16, 1)
    assert tot.mean() >= 18.0, tot
    tot, length = _score_v0(oracle_lib, "RoboschoolHopper-v1", 8, 1)
    assert (length == 1000).all(), length
    assert tot.min() > 2000.0, tot


def test_joint_motor_rows_closed_forms(oracle_lib):
    """btMultiBodyJointMotor rows (Joint::set_target_speed / set_servo_target, physics-bullet.cpp:510-537) on the cart-pole:
    a velocity motor with a generous impulse bound brings the slider to the target speed within one sub-step and holds
    it whatever the pole does; below saturation the impulse is exactly the bound (velocity change proportional to it);
    a PD servo on the hinge settles next to its target; bound 0 is torque control (no row)."""
    m = envspec.load_env_model("RoboschoolInvertedPendulum-v1")
    nd = m.n_dof
    w = oracle_lib.OracleWorld(m, 1)
    w.reset(seed=1)
    w.set_joint_motor(0, 0, 0.0, 1.0, 0.0, 0.7, 1000.0)
    for k in range(5):
        w.substeps(0, 1)
        assert abs(w.get_state(0)[nd] - 0.7) < 1e-12
    def fresh():                       # (a second reset of the same world would start episode 1: other joint draws)
        x = oracle_lib.OracleWorld(m, 1); x.reset(seed=1); return x
    free = fresh(); free.substeps(0, 1)
    dv = []
    for imp in (0.01, 0.02):
        w = fresh()
        w.set_joint_motor(0, 0, 0.0, 1.0, 0.0, 0.7, imp)
        w.substeps(0, 1)
        dv.append(w.get_state(0)[nd] - free.get_state(0)[nd])
    assert dv[0] > 0 and abs(dv[1] / dv[0] - 2.0) < 1e-9, dv
    w = fresh()
    w.set_joint_motor(0, 1, 0.1, 0.5, 0.3, 0.0, 1000.0)
    for k in range(300):
        w.substeps(0, 1)
    assert abs(w.get_state(0)[1] - 0.3) < 0.03, w.get_state(0)
    w = fresh()
    w.set_joint_motor(0, 1, 0.1, 0.5, 0.3, 0.0, 1000.0)
    w.set_joint_motor(0, 1, 0.1, 0.5, 0.3, 0.0, 0.0)           # off again
    w.substeps(0, 1)
    assert np.array_equal(w.get_state(0), free.get_state(0))


# ---------------------------------------------------------------------------------------------------------------------
# Mechanics identities: the oracle's articulated-body algorithm and its unit-impulse response are
# checked against the equations of motion built INDEPENDENTLY of any recursive algorithm -- the mass matrix and the
# gravity force are assembled in numpy from finite differences of the oracle's forward kinematics (link COM poses)
# alone:  M(x) = sum_i m_i Jv_i^T Jv_i + Jw_i^T (R_i I_i R_i^T) Jw_i ,  G = dV/dx ,  and at zero velocity
#   M a = tau - G        (ABA: computeAccelerationsArticulatedBodyAlgorithmMultiDof)
#   M u = f              (calcAccelerationDeltasMultiDof)
# with x = [base rotation vector (world), base position, q] and a / u / f in [w_base, v_base, qd] order.
# ---------------------------------------------------------------------------------------------------------------------
def _quat_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _perturbed_state(m, s, k, eps):
    """state with generalized coordinate k moved by eps: k<3 base rotation about world axis k, 3..5 base position, 6.. joints."""
    s = s.copy()
    if m.fixed_base:
        s[k - 6] += eps
        return s
    if k < 3:
        ax = np.zeros(3); ax[k] = 1.0
        dq = np.concatenate([ax * np.sin(0.5 * eps), [np.cos(0.5 * eps)]])
        x1, y1, z1, w1 = dq
        x2, y2, z2, w2 = s[3:7]
        s[3:7] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                  w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
    elif k < 6:
        s[k - 3] += eps
    else:
        s[13 + (k - 6)] += eps
    return s


def _mass_matrix_fd(oracle_lib, m, s, eps=1e-6):
    w = oracle_lib.OracleWorld(m, 1)
    nb = m.n_links + 1
    nd = 6 + m.n_dof
    k0 = 6 if m.fixed_base else 0

    def poses(state):
        w.set_state(0, state)
        P = w.link_poses(0)
        return P[:, 0:3].copy(), [_quat_mat(P[b, 3:7]) for b in range(nb)]
    p0, R0 = poses(s)
    Jv = np.zeros((nb, 3, nd)); Jw = np.zeros((nb, 3, nd))
    for k in range(k0, nd):
        pp, Rp = poses(_perturbed_state(m, s, k, eps))
        pm, Rm = poses(_perturbed_state(m, s, k, -eps))
        for b in range(nb):
            Jv[b, :, k] = (pp[b] - pm[b]) / (2 * eps)
            dR = (Rp[b] - Rm[b]) @ R0[b].T / (2 * eps)       # skew(omega) for a unit rate of coordinate k
            Jw[b, :, k] = [dR[2, 1], dR[0, 2], dR[1, 0]]
    masses = [m.base_mass] + [m.mass[i] for i in range(m.n_links)]
    inert = [np.array(m.arr("base_inertia"))] + [np.array(m.arr("inertia")[i]) for i in range(m.n_links)]
    M = np.zeros((nd, nd)); G = np.zeros(nd)
    for b in range(nb):
        if b == 0 and m.fixed_base:
            continue
        Iw = R0[b] @ np.diag(inert[b]) @ R0[b].T
        M += masses[b] * Jv[b].T @ Jv[b] + Jw[b].T @ Iw @ Jw[b]
        G += masses[b] * m.gravity * Jv[b][2]
    return M, G


@pytest.mark.parametrize("env_id", ["RoboschoolHopper-v1", "RoboschoolWalker2d-v1", "RoboschoolHalfCheetah-v1", "RoboschoolAnt-v1",
                                    "RoboschoolHumanoid-v1", "RoboschoolInvertedDoublePendulum-v1"])
def test_aba_and_response_satisfy_equations_of_motion(env_id, oracle_lib):
    m = envspec.load_env_model(env_id)
    rng = np.random.RandomState(5)
    nd = 6 + m.n_dof
    k0 = 6 if m.fixed_base else 0
    for trial in range(3):
        s = np.zeros(m.state_dim)
        q = rng.uniform(-0.6, 0.6, m.n_dof)
        for i in range(m.n_links):                # inside the joint ranges: no limit rows in this test
            if m.dof_idx[i] >= 0 and m.has_limit[i]:
                q[m.dof_idx[i]] = m.lim_lo[i] + rng.uniform(0.15, 0.85) * (m.lim_hi[i] - m.lim_lo[i])
        if m.fixed_base:
            s[:m.n_dof] = q
            names = m.meta["joint_names"]
            for i, nme in enumerate(names):       # lift the planar robots off the floor: no contact rows in this test
                if nme in ("ignorez", "ignore2"):
                    s[m.dof_idx[i]] = 1.0
        else:
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = rng.uniform(-1, 1)
            s[0:3] = [0.1, -0.2, 3.0]
            s[3:7] = np.concatenate([ax * np.sin(0.5 * ang), [np.cos(0.5 * ang)]])
            s[13:13 + m.n_dof] = q
        M, G = _mass_matrix_fd(oracle_lib, m, s)
        tau = rng.uniform(-20, 20, m.n_dof)
        w = oracle_lib.OracleWorld(m, 1)
        w.set_state(0, s)
        w.set_tau(0, tau)
        w.substeps(0, 1)
        s1 = w.get_state(0)
        dt = m.timestep
        a = np.zeros(nd)
        if m.fixed_base:
            a[6:] = s1[m.n_dof:] / dt
        else:
            a[0:3] = s1[7:10] / dt; a[3:6] = s1[10:13] / dt; a[6:] = s1[13 + m.n_dof:] / dt
        rhs = -G.copy(); rhs[6:] += tau
        res = (M @ a - rhs)[k0:]
        scale = max(1.0, np.abs(rhs[k0:]).max())
        assert np.abs(res).max() / scale < 1e-6, (env_id, trial, np.abs(res).max(), scale)
        # unit-impulse response: M u = f for random generalized forces (the solver's M^-1 J^T)
        w.set_state(0, s)
        for r in range(3):
            f = np.zeros(nd); f[k0:] = rng.normal(size=nd - k0)
            u = w.accel_deltas(0, f)
            assert np.abs((M @ u - f)[k0:]).max() < 1e-6 * max(1.0, np.abs(u).max() * np.abs(M).max()), (env_id, trial, r)


@pytest.mark.parametrize("world", ["oracle", "emulated-kernel"])
def test_cube_hit_exchanges_momentum_between_the_two_bodies(world, oracle_lib):
    """Two-body contact rows (robot geom vs FlagrunHarder's cube) are equal and opposite: with gravity and damping switched off
    and the floor out of reach, the cube thrown at the free-floating humanoid bounces off, each body's momentum changes by a clearly
    non-zero amount, and the TOTAL linear momentum of robot + cube is conserved up to the integrator's own first-order drift
    (the semi-implicit step keeps the joint velocities while it moves the joints: O(dt) per sub-step, see
    test_momentum_and_energy_converge_in_free_flight) -- so the mismatch must shrink with the time step and be small at a fine
    one; rows that were not action = -reaction would leave a mismatch that does not."""
    errs = []
    for div in (1, 4, 16):
        m = envspec.load_env_model("RoboschoolHumanoidFlagrunHarder-v1").clone()
        # no gravity, no damping; floor (5 m below), joint limits and self-collision pairs stay: they are internal to the robot
        # or out of reach, and the generated topology of the kernel (TopoHumanoidCube) is matched on them
        m.lin_damping = 0; m.ang_damping = 0; m.gravity = 0.0
        m.timestep = m.timestep / div
        S = m.state_dim
        mass = np.array([m.base_mass] + [m.mass[i] for i in range(m.n_links)])
        if world == "oracle":
            w = oracle_lib.OracleWorld(m, 1)
        else:
            from tests.host_emul import emul
            emul.build()
            w = emul.EmulWorld(m, 1)
            assert w.use_static(True) == 6              # the cube is stepped by the static-topology code only
        s0 = np.zeros(S)
        s0[0:3] = [0.0, 0.0, 5.0]; s0[6] = 1.0
        s0[S - 13:S - 10] = [-0.6, 0.02, 5.05]          # cube just in front of the torso
        s0[S - 7] = 1.0                                  # cube quaternion w
        s0[S - 3:S] = [20.0, 0.0, 0.0]                   # 20 m/s at the robot
        w.set_state(0, s0)
        n_contact = 0
        for _ in range(24 * div):                        # 0.1 s: approach, hit, separation
            w.substeps(0, 1)
            n_contact += int(len(w.contacts(0)) > 0)
        se = np.array(w.get_state(0), dtype=np.float64)
        p_robot = (mass[:, None] * np.asarray(w.link_poses(0))[:, 7:10]).sum(0)
        d_cube = m.cube_mass * (se[S - 3:S] - s0[S - 3:S])
        assert n_contact > 0 and abs(d_cube[0]) > 0.2 * m.cube_mass * 20.0, d_cube     # it hit: a good part of 24 kg m/s changed hands
        errs.append(np.abs(p_robot + d_cube).max() / np.abs(d_cube).max())
    assert errs[1] < errs[0] / 2 and errs[2] < errs[1] / 2 and errs[2] < 1e-2, errs      # measured 4.0 %, 1.3 %, 0.56 % at dt, dt / 4, dt / 16
