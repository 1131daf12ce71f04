"""Model compiler: structural fixtures derived from the reference MJCFs and the
shipped pre-compiled descriptors."""
import os

import numpy as np
import pytest

from roboschool_b200 import envspec
from roboschool_b200.model import ModelDesc
from tests.conftest import HAVE_REF, REF_ASSETS

EXPECT = {  # env: (links, dof, geoms, fixed_base, A, O, feet)
    "RoboschoolInvertedPendulum-v1": (2, 2, 2, 1, 1, 5, 0),
    "RoboschoolHopper-v1": (6, 6, 4, 1, 3, 15, 1),
    "RoboschoolWalker2d-v1": (9, 9, 7, 1, 6, 22, 2),
    "RoboschoolAnt-v1": (12, 8, 13, 0, 8, 28, 4),
    "RoboschoolHumanoid-v1": (19, 17, 17, 0, 17, 44, 2),
}


@pytest.mark.parametrize("env_id", sorted(EXPECT))
def test_shipped_descriptor_structure(env_id):
    m = envspec.load_env_model(env_id)
    L, D, G, fb, A, O, F = EXPECT[env_id]
    assert (m.n_links, m.n_dof, m.n_geoms, m.fixed_base, m.n_act, m.obs_dim, m.n_feet) == (L, D, G, fb, A, O, F)
    assert m.obs_dim == 8 + 2 * m.n_act + m.n_feet or env_id.startswith("RoboschoolInvertedPendulum")
    # tree is topologically ordered, quaternions unit, axes unit
    for i in range(m.n_links):
        assert -1 <= m.parent[i] < i
        assert abs(np.linalg.norm(m.arr("zero_rot")[i]) - 1) < 1e-12
        assert abs(np.linalg.norm(m.arr("axis")[i]) - 1) < 1e-12
    assert m.n_iter == 5 and m.frame_skip in (1, 4)
    assert m.gravity == float(np.float32(9.8))


def test_humanoid_details():
    m = envspec.load_env_model("RoboschoolHumanoid-v1")
    names = m.meta["ordered_joints"]
    assert names == ["abdomen_z", "abdomen_y", "abdomen_x", "right_hip_x", "right_hip_z", "right_hip_y", "right_knee",
                     "left_hip_x", "left_hip_z", "left_hip_y", "left_knee", "right_shoulder1", "right_shoulder2",
                     "right_elbow", "left_shoulder1", "left_shoulder2", "left_elbow"]
    gains = [m.act_gain[k] for k in range(17)]
    expect = [0.41 * c for c in [100, 100, 100, 100, 100, 300, 200, 100, 100, 300, 200, 75, 75, 75, 75, 75, 75]]
    assert np.allclose(gains, expect)
    assert m.body_link == -1 and m.alive_z == 0.78 and m.alive_bonus == 2.0
    assert abs(m.electricity_cost + 8.5) < 1e-12 and abs(m.stall_torque_cost + 0.425) < 1e-12
    # knee range -160..-2 degrees
    k = m.meta["joint_names"].index("right_knee")
    assert np.isclose(m.lim_lo[k], np.radians(-160)) and np.isclose(m.lim_hi[k], np.radians(-2))
    # self-collision pairs never involve ancestor/descendant links and never the base (ancestor of all)
    anc = {}
    for i in range(m.n_links):
        a, p = set(), m.parent[i]
        while p != -1:
            a.add(p); p = m.parent[p]
        a.add(-1)
        anc[i] = a
    for k in range(m.n_pairs):
        la, lb = m.g_link[m.pair_a[k]], m.g_link[m.pair_b[k]]
        assert la != lb and la != -1 and lb != -1 and la not in anc[lb] and lb not in anc[la]
    total = m.base_mass + sum(m.mass[i] for i in range(m.n_links))
    assert 35 < total < 50


def test_ant_floor_only():
    m = envspec.load_env_model("RoboschoolAnt-v1")
    assert m.n_pairs == 0 and all(m.g_floor[g] for g in range(m.n_geoms))
    assert sum(1 for i in range(m.n_links) if m.jtype[i] == 0) == 4   # four fixed 'leg' links


@pytest.mark.skipif(not HAVE_REF, reason="reference MJCFs not present")
@pytest.mark.parametrize("env_id", sorted(EXPECT))
def test_recompile_matches_shipped(env_id):
    a = envspec.compile_env(env_id, REF_ASSETS)
    b = envspec.load_env_model(env_id)
    assert bytes(a) == bytes(b), "shipped descriptor is stale: run tools/compile_models.py"


def test_descriptor_roundtrip(tmp_path):
    m = envspec.load_env_model("RoboschoolHopper-v1")
    p = tmp_path / "m.json"
    m.save(str(p))
    assert bytes(ModelDesc.load(str(p))) == bytes(m)
