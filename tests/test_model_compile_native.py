"""rs_model_compile (the C entry of the model compiler, roboschool_b200/csrc/rs_model_compile.cpp) against the Python
compiler roboschool_b200/mjcf.py: every physics field of the descriptor and the names, on synthetic MJCFs written here
(so the test needs no reference tree) and -- where /root/reference is present -- on every MJCF the reference ships for the
envs of this repo (World::load_mjcf, physics-bullet.cpp:95-131)."""
import ctypes
import os

import numpy as np
import pytest

from roboschool_b200 import _lib, mjcf
from roboschool_b200.model import ModelDesc, DEFINES
from tests.conftest import HAVE_REF, REF_ASSETS

NAME_LEN = 48


class ModelNames(ctypes.Structure):
    _fields_ = [("base_name", ctypes.c_char * NAME_LEN),
                ("link_names", (ctypes.c_char * NAME_LEN) * DEFINES["RS_MAX_LINKS"]),
                ("joint_names", (ctypes.c_char * NAME_LEN) * DEFINES["RS_MAX_LINKS"]),
                ("geom_names", (ctypes.c_char * NAME_LEN) * DEFINES["RS_MAX_GEOMS"])]


def native_compile(path, gravity=9.8, timestep=0.0165 / 4, frame_skip=4, has_floor=True):
    L = _lib.lib()
    m, names = ModelDesc(), ModelNames()
    _lib.check(L.rs_model_compile(path.encode(), gravity, timestep, frame_skip, int(has_floor), ctypes.byref(m), ctypes.byref(names)))
    return m, names


SKIP = {"max_contacts"}      # set by the task layer in Python (envspec.py), 16 by the C entry


def assert_same(a, b):
    for name, _t in ModelDesc._fields_:
        if name in SKIP:
            continue
        va, vb = getattr(a, name), getattr(b, name)
        if isinstance(va, (int, float)):
            assert va == pytest.approx(vb, rel=1e-13, abs=1e-15), name
        else:
            xa, xb = np.ctypeslib.as_array(va), np.ctypeslib.as_array(vb)
            if xa.dtype.kind == "i":
                assert (xa == xb).all(), name
            else:
                assert np.allclose(xa, xb, rtol=1e-13, atol=1e-15), (name, np.abs(xa - xb).max())


def check_file(path, **kw):
    py = mjcf.compile_mjcf(path, **kw)
    c, names = native_compile(path, **kw)
    assert_same(c, py)
    assert c.max_contacts == 16
    assert names.base_name.decode() == py.meta["base_name"]
    for i in range(py.n_links):
        assert names.link_names[i].value.decode() == py.meta["link_names"][i]
        assert names.joint_names[i].value.decode() == (py.meta["joint_names"][i] or "")
    for g in range(py.n_geoms):
        assert names.geom_names[g].value.decode() == py.meta["geom_names"][g]
    return c


SYNTH_FLOATING = """<?xml version="1.0"?>
<!-- a floating torso with a box foot, a rotated capsule, an axis-angle sphere body and a two-joint hip -->
<mujoco model="synthetic">
  <compiler angle="degree" inertiafromgeom="true"/>
  <default>
    <joint armature="1" damping="1" limited="true"/>
    <geom conaffinity="1" contype="1" friction="0.7 0.1 0.1" density="500"/>
  </default>
  <worldbody>
    <body name="torso" pos="0 0 1.25" quat="0.9238795 0 0.3826834 0">
      <geom name="torso_geom" type="capsule" fromto="0 -.07 0 0 .07 0" size="0.07"/>
      <geom name="head" type="sphere" pos="0 0 .19" size=".09"/>
      <body name="thigh" pos="0 -0.1 -0.04">
        <joint name="hip_x" type="hinge" axis="1 0 0" pos="0 0 0" range="-25 5"/>
        <joint name="hip_y" type="hinge" axis="0 1 0" pos="0 0 0.01" range="-110 20"/>
        <geom name="thigh_geom" type="capsule" fromto="0 0 0 0 0.01 -.34" size="0.06"/>
        <body name='shin' pos='0 0.01 -0.403'>
          <joint name="knee" type="hinge" axis="0 -1 0" range="-160 -2"/>
          <geom name="shin_geom" type="capsule" pos="0 0 -.15" size="0.049 .15" quat="0.9659258 0.258819 0 0"/>
          <body name="foot" pos="0 0 -.39">
            <geom name="foot_geom" type="box" pos="0.05 0 0" size="0.1 0.04 0.02" axisangle="0 1 0 10"/>
          </body>
        </body>
      </body>
      <body name="arm" pos="0 0.17 0.06">
        <joint name="shoulder" type="hinge" axis="2 1 1" range="-85 60" limited="false"/>
        <geom name="arm_geom" type="sphere" size="0.04" pos=".16 .16 .16" contype="0" conaffinity="0"/>
        <inertial mass="0.35" pos="0 0 0"/>
      </body>
    </body>
  </worldbody>
</mujoco>
"""

SYNTH_JOINTED_ROOTS = """<mujoco model="two_roots">
  <compiler angle="radian"/>
  <default><geom contype="0" friction="1 0.1 0.1"/></default>
  <worldbody>
    <geom name="ground" type="plane" size="1 1 1"/>
    <body name="cart" pos="0 0 0">
      <joint name="slider" type="slide" axis="1 0 0" pos="0 0 0" range="-1 1" limited="true"/>
      <geom name="cart_geom" type="capsule" fromto="-0.1 0 0 0.1 0 0" size="0.1"/>
      <body name="pole" pos="0 0 0">
        <joint name="hinge" type="hinge" axis="0 1 0" range="-1.57 1.57"/>
        <geom name="pole_geom" type="capsule" fromto="0 0 0 0.001 0 0.6" size="0.049"/>
      </body>
    </body>
    <body name="target" pos=".1 -.1 .01">
      <joint name="target_x" type="slide" axis="1 0 0" range="-.27 .27" limited="true"/>
      <joint name="target_y" type="slide" axis="0 1 0" range="-.27 .27" limited="true"/>
      <geom name="target_geom" type="sphere" size=".009" conaffinity="0"/>
    </body>
  </worldbody>
</mujoco>
"""


def test_native_matches_python_floating_base(tmp_path):
    p = tmp_path / "synth.xml"
    p.write_text(SYNTH_FLOATING)
    c = check_file(str(p))
    assert (c.n_links, c.n_dof, c.n_geoms, c.fixed_base) == (5, 4, 6, 0)
    assert c.jtype[3] == DEFINES["RS_JOINT_FIXED"] and c.dof_idx[3] == -1           # the jointless foot
    assert c.g_type[4] == DEFINES["RS_GEOM_BOX"]
    assert c.has_limit[4] == 0 and c.lim_lo[4] == 0.0 and c.lim_hi[4] == -1.0        # limited="false" -> Bullet's lo > hi


def test_native_matches_python_jointed_roots(tmp_path):
    p = tmp_path / "roots.xml"
    p.write_text(SYNTH_JOINTED_ROOTS)
    c = check_file(str(p), gravity=0.0, timestep=0.0165, frame_skip=1, has_floor=False)
    assert (c.n_links, c.n_dof, c.fixed_base, c.n_pairs) == (4, 4, 1, 0)
    assert c.gravity == 0.0 and c.timestep == float(np.float32(0.0165)) and c.frame_skip == 1 and c.has_floor == 0


def test_native_errors_are_reported(tmp_path):
    L = _lib.lib()
    m = ModelDesc()
    assert L.rs_model_compile(b"/nonexistent/file.xml", 9.8, 0.004125, 4, 1, ctypes.byref(m), None) != 0
    assert b"cannot open" in L.rs_last_error()
    bad = tmp_path / "bad.xml"
    bad.write_text("<mujoco><worldbody><body name='b'><geom type='mesh' size='1'/></body></worldbody></mujoco>")
    assert L.rs_model_compile(str(bad).encode(), 9.8, 0.004125, 4, 1, ctypes.byref(m), None) != 0
    assert b"not supported" in L.rs_last_error()
    broken = tmp_path / "broken.xml"
    broken.write_text("<mujoco><worldbody><body name='b'></worldbody></mujoco>")
    assert L.rs_model_compile(str(broken).encode(), 9.8, 0.004125, 4, 1, ctypes.byref(m), None) != 0
    assert b"XML" in L.rs_last_error()
    ball = tmp_path / "ball.xml"
    ball.write_text("<mujoco><worldbody><body name='b'><joint type='ball'/><geom type='sphere' size='1'/></body></worldbody></mujoco>")
    assert L.rs_model_compile(str(ball).encode(), 9.8, 0.004125, 4, 1, ctypes.byref(m), None) != 0
    assert b"joint type" in L.rs_last_error()


REF_XMLS = ["hopper.xml", "walker2d.xml", "half_cheetah.xml", "ant.xml", "humanoid_symmetric.xml", "inverted_pendulum.xml",
            "inverted_double_pendulum.xml", "reacher.xml"]


@pytest.mark.skipif(not HAVE_REF, reason="reference MJCFs not present")
@pytest.mark.parametrize("xml", REF_XMLS)
def test_native_matches_python_on_reference_models(xml):
    check_file(os.path.join(REF_ASSETS, xml))


@pytest.mark.skipif(not HAVE_REF, reason="reference MJCFs not present")
def test_native_descriptor_equals_shipped_model():
    """The physics part of the shipped (pre-compiled) Humanoid descriptor is what the C entry produces from the reference's file."""
    from roboschool_b200 import envspec
    shipped = envspec.load_env_model("RoboschoolHumanoid-v1")
    c, _ = native_compile(os.path.join(REF_ASSETS, "humanoid_symmetric.xml"))
    for name in ("n_links", "n_dof", "n_geoms", "n_pairs", "fixed_base"):
        assert getattr(c, name) == getattr(shipped, name), name
    for name in ("parent", "jtype", "dof_idx", "axis", "zero_rot", "e_vec", "d_vec", "mass", "inertia", "lim_lo", "lim_hi",
                 "g_link", "g_p0", "g_p1", "g_radius", "pair_a", "pair_b", "break_thresh", "base_inertia"):
        assert np.allclose(np.ctypeslib.as_array(getattr(c, name)), np.ctypeslib.as_array(getattr(shipped, name)), rtol=1e-13, atol=1e-15), name
    assert c.base_mass == pytest.approx(shipped.base_mass, rel=1e-13)


@pytest.mark.skipif(not HAVE_REF, reason="reference MJCFs not present")
@pytest.mark.parametrize("env_id", ["RoboschoolHopper-v1", "RoboschoolAnt-v1", "RoboschoolHumanoidFlagrunHarder-v1", "RoboschoolReacher-v1"])
def test_env_descriptor_through_native_compiler(env_id):
    """compile_env(native=True): C entry + the Python task layer == the all-Python path, byte for byte of the ABI struct's
    integer fields and to 1e-13 for the doubles."""
    from roboschool_b200 import envspec
    a = envspec.compile_env(env_id, REF_ASSETS, native=True)
    b = envspec.compile_env(env_id, REF_ASSETS)
    SKIP.discard("max_contacts")
    try:
        assert_same(a, b)
    finally:
        SKIP.add("max_contacts")
    assert a.meta["ordered_joints"] == b.meta["ordered_joints"]
