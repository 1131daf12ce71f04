"""Drop-in `cpp_household` module: Pose / Thingy / Joint / Robot / World with the reference's names,
argument order and error behaviour (roboschool/cpp-household/python-binding.cpp:611-713), implemented
as an N=1 view over the batched CUDA stepper's C ABI.

    import roboschool_b200.cpp_household as cpp_household
    sys.modules["roboschool.cpp_household"] = cpp_household

What is kept: World(gravity, timestep), clean_everything, load_mjcf, step(repeat) -> False, ts;
Robot.joints/parts/root_part/set_pose/set_pose_and_speed/query_position; Joint.name/type/
set_motor_torque/current_position/current_relative_position/reset_current_position/limits;
Thingy.pose/speed/name/contact_list/__hash__/__eq__; Pose math.  Rendering (Camera, test_window*,
load_thingy meshes) are inert stubs: the renderer is out of scope.
Bullet's PD/velocity motor modes (Joint.set_servo_target / set_target_speed, physics-bullet.cpp:510-537) are
btMultiBodyJointMotor rows of the solver (rs_world_set_joint_motor), with the reference's first_torque_call state
machine.  No exceptions on load failure: message on stderr + empty list, like the reference
(physics-bullet.cpp:105-108).
"""
import math
import os
import sys

import numpy as np

from . import mjcf
from .model import RS_JOINT_REVOLUTE, RS_JOINT_FIXED

SCALE = 1.0
COLLISION_MARGIN = 0.002 * SCALE
tip_z, tip_y = 0, 0
METACLASS_FLOOR, METACLASS_WALL, METACLASS_MYSELF, METACLASS_FURNITURE = 0x01, 0x02, 0x04, 0x08
METACLASS_HANDLE, METACLASS_ITEM = 0x10, 0x20


def _make_backend(model, n_envs=1, device=0):
    """The CUDA world (no CPU fallback).  tests/ monkeypatch this to drive the shim from the oracle."""
    from .world import BatchedWorld
    return BatchedWorld(model, n_envs, device)


class Pose(object):
    """python-binding.cpp:35-80"""

    def __init__(self):
        self.x = self.y = self.z = 0.0
        self.qx = self.qy = self.qz = 0.0
        self.qw = 1.0

    def set_xyz(self, x, y, z):
        self.x, self.y, self.z = x * SCALE, y * SCALE, z * SCALE

    def move_xyz(self, x, y, z):
        self.x += x * SCALE; self.y += y * SCALE; self.z += z * SCALE

    def set_quaternion(self, x, y, z, w):
        self.qx, self.qy, self.qz, self.qw = x, y, z, w

    def set_rpy(self, r, p, y):
        t0, t1 = math.cos(y * 0.5), math.sin(y * 0.5)
        t2, t3 = math.cos(r * 0.5), math.sin(r * 0.5)
        t4, t5 = math.cos(p * 0.5), math.sin(p * 0.5)
        self.qw = t0 * t2 * t4 + t1 * t3 * t5
        self.qx = t0 * t3 * t4 - t1 * t2 * t5
        self.qy = t0 * t2 * t5 + t1 * t3 * t4
        self.qz = t1 * t2 * t4 - t0 * t3 * t5

    def rotate_z(self, angle):
        s, c = math.sin(angle * 0.5), math.cos(angle * 0.5)
        # t = q(z, angle) * self
        x, y, z, w = self.qx, self.qy, self.qz, self.qw
        self.qx = c * x - s * y
        self.qy = c * y + s * x
        self.qz = c * z + s * w
        self.qw = c * w - s * z

    def xyz(self):
        return (self.x / SCALE, self.y / SCALE, self.z / SCALE)

    def quatertion(self):   # (sic) python-binding.cpp:58
        return (self.qx, self.qy, self.qz, self.qw)

    def rpy(self):
        qx, qy, qz, qw = self.qx, self.qy, self.qz, self.qw
        sqw, sqx, sqy, sqz = qw * qw, qx * qx, qy * qy, qz * qz
        t2 = -2.0 * (qx * qz - qy * qw) / (sqx + sqy + sqz + sqw)
        yaw = math.atan2(2.0 * (qx * qy + qz * qw), (sqx - sqy - sqz + sqw))
        roll = math.atan2(2.0 * (qy * qz + qx * qw), (-sqx - sqy + sqz + sqw))
        t2 = min(1.0, max(-1.0, t2))
        return (roll, math.asin(t2), yaw)

    def dot(self, other):
        # btTransform product: rotate other's origin by self, compose quaternions
        ax, ay, az, aw = self.qx, self.qy, self.qz, self.qw
        bx, by, bz, bw = other.qx, other.qy, other.qz, other.qw
        r = Pose()
        r.qx = aw * bx + ax * bw + ay * bz - az * by
        r.qy = aw * by - ax * bz + ay * bw + az * bx
        r.qz = aw * bz + ax * by - ay * bx + az * bw
        r.qw = aw * bw - ax * bx - ay * by - az * bz
        v = np.array([other.x, other.y, other.z])
        u = np.array([ax, ay, az])
        rv = v + 2.0 * np.cross(u, np.cross(u, v) + aw * v)
        r.x, r.y, r.z = self.x + rv[0], self.y + rv[1], self.z + rv[2]
        return r


class Thingy(object):
    """A robot part (root or link): python-binding.cpp:82-124."""

    def __init__(self, robot, body_index, name):
        self._robot, self._b, self.name = robot, body_index, name
        self.visibility_123 = 0

    def __hash__(self):
        return hash((id(self._robot), self._b))

    def __eq__(self, other):
        return isinstance(other, Thingy) and self._robot is other._robot and self._b == other._b

    def pose(self):
        p = Pose()
        row = self._robot._pose_row(self._b)
        off = self._robot._offset
        p.set_xyz((float(row[0]) + off[0]) / SCALE, (float(row[1]) + off[1]) / SCALE, (float(row[2]) + off[2]) / SCALE)
        p.set_quaternion(float(row[3]), float(row[4]), float(row[5]), float(row[6]))
        return p

    def speed(self):
        row = self._robot._pose_row(self._b)
        assert self._robot._queried, "speed() before any query_position() (python-binding.cpp:92)"
        return (float(row[7]) / SCALE, float(row[8]) / SCALE, float(row[9]) / SCALE)

    def contact_list(self):
        return self._robot._contact_list(self._b)

    def set_multiply_color(self, tex, color):
        pass

    def assign_metaclass(self, mclass):
        pass


class Joint(object):
    """python-binding.cpp:244-276; household.h:65-100."""

    def __init__(self, robot, link, name):
        self._robot, self._link, self.name = robot, link, name
        m = robot._model
        self._dof = m.dof_idx[link]
        self._revolute = m.jtype[link] == RS_JOINT_REVOLUTE
        self._l1, self._l2 = np.float32(m.lim_lo[link]), np.float32(m.lim_hi[link])
        self._max_vel = np.float32(m.max_vel[link])
        self._torque = np.float32(0.0)
        self._first_torque_call, self._torque_need_repeat = True, False

    def type(self):
        return "motor" if self._revolute else "linear_motor"

    def set_motor_torque(self, q):
        """physics-bullet.cpp:497-508: the first torque call after a motor command (or after loading) switches Bullet's
        joint motor off with set_servo_target(0, 0.1, 0.1, maxforce 0); the torque is then re-applied every step."""
        if self._first_torque_call:
            self.set_servo_target(0, 0.1, 0.1, 0)
            self._first_torque_call = False
        self._torque_need_repeat = True
        self._torque = np.float32(q)

    def _server_dt(self):
        # maxForce is turned into an impulse bound with the physics server's time step, which bullet_step sets to
        # timestep * skip_frames (physics-bullet.cpp:363-376); before the first step it still has Bullet's default 1/240
        w = self._robot._world
        return w._timestep * w._last_repeat if w._last_repeat else 1.0 / 240.0

    def _set_motor(self, kp, kd, target_pos, target_vel, maxforce):
        b = self._robot._backend
        if b is not None and self._dof >= 0:
            b.set_joint_motor(self._dof, kp, kd, target_pos, target_vel, float(np.float32(maxforce)) * self._server_dt())
        self._first_torque_call = True
        self._torque_need_repeat = False

    def set_servo_target(self, pos, kp, kd, maxforce):
        """physics-bullet.cpp:524-537 (CONTROL_MODE_POSITION_VELOCITY_PD: position target, velocity target 0)."""
        self._set_motor(float(np.float32(kp)), float(np.float32(kd)), float(np.float32(pos)), 0.0, maxforce)

    def set_target_speed(self, v, kd, maxforce):
        """physics-bullet.cpp:510-522 (CONTROL_MODE_VELOCITY: kp 0)."""
        self._set_motor(0.0, float(np.float32(kd)), 0.0, float(np.float32(v)), maxforce)

    def current_position(self):
        q, qd = self._robot._joint_state(self._dof)
        return (float(np.float32(q)), float(np.float32(qd)))

    def current_relative_position(self):
        """physics-bullet.cpp:547-566 (float arithmetic)."""
        q, qd = self._robot._joint_state(self._dof)
        rpos, rspeed = np.float32(q), np.float32(qd)
        if self._l1 < self._l2:
            pos_mid = np.float32(0.5 * (float(self._l1) + float(self._l2)))
            rpos = np.float32(2) * (rpos - pos_mid) / (self._l2 - self._l1)
        if self._max_vel > 0:
            rspeed = rspeed / self._max_vel
        else:
            rspeed = rspeed * (np.float32(0.1) if self._revolute else np.float32(0.5))
        return (float(rpos), float(rspeed))

    def reset_current_position(self, pos, vel):
        self._robot._set_joint(self._dof, float(np.float32(pos)), float(np.float32(vel)))

    def limits(self):
        return (float(self._l1), float(self._l2), 0.0, float(self._max_vel))


class Robot(object):
    """python-binding.cpp:278-295."""

    def __init__(self, world, model, static_name=None):
        self._world, self._model = world, model
        self._queried = False
        self._poses = None
        self._state = None
        # A fixed-base robot (planar walkers, pendulums) is anchored at the world origin in its own world; the floor is an infinite
        # plane, so moving its base (the multiplayer stadium puts player n in lane y = n - 1, scene_stadium.py:23-29) is a pure
        # translation of what the readback reports.  Floating-base robots carry their position in the state instead.
        self._offset = (0.0, 0.0, 0.0)
        if model is None:      # static body (floor): a root part only
            self.root_part = Thingy(self, 0, static_name)
            self.parts, self.joints = [], []
            self._backend = None
            return
        meta = model.meta
        self._backend = _make_backend(model, 1)
        self.root_part = Thingy(self, 0, meta["base_name"])
        self.parts = [Thingy(self, i + 1, n) for i, n in enumerate(meta["link_names"])]
        self.joints = [Joint(self, i, jn) for i, jn in enumerate(meta["joint_names"]) if jn is not None]
        # XML pose
        s = np.zeros((1, self._backend.S), dtype=np.float32)
        if not model.fixed_base:
            s[0, 0:3] = model.arr("reset_base_pos")
            s[0, 3:7] = model.arr("reset_base_quat")
        self._backend.set_state(s, clear_contacts=True)
        self._state = s

    # ---- internal ----
    def _attach_box(self, half, mass, pose):
        """Rebuild this robot's world with a free box as second body (load_urdf of a single-box URDF): same descriptor plus the
        cube block -- inertia of the collision AABB (load_urdf passes no URDF_USE_INERTIA_FROM_FILE), URDF importer default
        friction 0.5, contact-breaking threshold 0.02 x the half diagonal -- and the current state carried over."""
        m = self._model.clone()
        hx, hy, hz = half
        m.has_cube = 1
        m.arr("cube_half")[:] = [hx, hy, hz]
        m.cube_mass = mass
        lx, ly, lz = 2 * hx, 2 * hy, 2 * hz
        m.arr("cube_inertia")[:] = [mass / 12.0 * (ly * ly + lz * lz), mass / 12.0 * (lx * lx + lz * lz), mass / 12.0 * (lx * lx + ly * ly)]
        m.cube_friction = 0.5
        m.cube_break_thresh = 0.02 * math.sqrt(hx * hx + hy * hy + hz * hz)
        m.arr("cube_start")[:] = [pose.x, pose.y, pose.z]
        old = self._backend
        s_old = np.asarray(old.get_state(), dtype=np.float64)
        nb = _make_backend(m, 1)
        s = np.zeros((1, nb.S), dtype=np.float64)
        s[0, :s_old.shape[1]] = s_old[0]
        c0 = s_old.shape[1]
        s[0, c0:c0 + 3] = [pose.x, pose.y, pose.z]
        s[0, c0 + 3:c0 + 7] = [pose.qx, pose.qy, pose.qz, pose.qw]
        nb.set_state(s, clear_contacts=True)
        if hasattr(old, "close"):
            old.close()
        self._backend, self._model = nb, m
        self._state, self._poses = None, None

    def _refresh(self):
        if self._backend is None:
            self._poses = np.zeros((1, 10), dtype=np.float32)
            self._poses[0, 6] = 1.0
        else:
            self._state = self._backend.get_state()
            self._poses = self._backend.link_poses()[0]
        self._queried = True

    def _pose_row(self, b):
        if self._poses is None:
            self._refresh()
        return self._poses[b]

    def _joint_state(self, dof):
        if self._state is None:
            self._refresh()
        off = 0 if self._model.fixed_base else 13
        nd = self._model.n_dof
        return self._state[0, off + dof], self._state[0, off + nd + dof]

    def _set_joint(self, dof, q, qd):
        s = self._backend.get_state()
        off = 0 if self._model.fixed_base else 13
        nd = self._model.n_dof
        s[0, off + dof] = q
        s[0, off + nd + dof] = qd
        self._backend.set_state(s, clear_contacts=False)
        self._state = s
        self._poses = None

    def _contact_list(self, b):
        if self._backend is None:
            return []
        keys, _ = self._backend.contacts(0)
        m = self._model
        out = []
        for k in keys:
            k = int(k)
            if k < m.n_geoms:
                if m.g_link[k] + 1 == b and self._world._floor is not None:
                    out.append(self._world._floor.root_part)
            else:
                pk = k - m.n_geoms
                ba, bb = m.g_link[m.pair_a[pk]] + 1, m.g_link[m.pair_b[pk]] + 1
                if ba == b:
                    out.append(self.root_part if bb == 0 else self.parts[bb - 1])
                elif bb == b:
                    out.append(self.root_part if ba == 0 else self.parts[ba - 1])
        return out

    def _torques(self):
        tau = np.zeros((1, max(self._model.n_dof, 1)), dtype=np.float32)
        for j in self.joints:
            if j._torque_need_repeat:
                tau[0, j._dof] = j._torque
        return tau[:, :self._model.n_dof]

    # ---- API ----
    def pose(self):
        return self.root_part.pose()

    def query_position(self):
        self._refresh()

    def set_pose(self, p):
        self.set_pose_and_speed(p, 0, 0, 0)

    def set_pose_and_speed(self, p, vx, vy, vz):
        """World::robot_move (physics-bullet.cpp:583-594): base pose and base LINEAR velocity only."""
        if self._backend is None:
            return
        if self._model.fixed_base:
            raw = self._pose_row(0)
            self._offset = (p.x - float(raw[0]), p.y - float(raw[1]), p.z - float(raw[2]))
            return
        s = self._backend.get_state()
        s[0, 0:3] = [p.x, p.y, p.z]
        s[0, 3:7] = [p.qx, p.qy, p.qz, p.qw]
        s[0, 10:13] = [vx, vy, vz]
        self._backend.set_state(s, clear_contacts=False)
        self._state = s
        self._poses = None


def _parse_single_box_urdf(fn):
    """A URDF with one link, no joint and one <collision><geometry><box size=...>: (half extents, mass) -- the thrown cube of
    RoboschoolHumanoidFlagrunHarder (models_household/cube.urdf).  Anything else: None (the general URDF importer is SURVEY 8f.3)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(fn).getroot()
    links, joints = root.findall("link"), root.findall("joint")
    if root.tag != "robot" or len(links) != 1 or joints:
        return None
    col = links[0].findall("collision")
    if len(col) != 1:
        return None
    origin = col[0].find("origin")
    if origin is not None and any(abs(float(v)) > 0 for k in ("xyz", "rpy") for v in origin.attrib.get(k, "0 0 0").split()):
        return None
    box = col[0].find("geometry/box")
    mass = links[0].find("inertial/mass")
    if box is None or mass is None:
        return None
    size = [float(v) for v in box.attrib["size"].split()]
    return [0.5 * v * SCALE for v in size], float(mass.attrib["value"]), links[0].attrib.get("name", "cube")


class _FreeBoxRobot(Robot):
    """The robot handle World::load_urdf returns for a free box (python-binding.cpp:278-295 surface).  The box is the second
    free body of its host robot's world (FlagrunHarder's cube): the last 13 state words of the host's
    backend -- position, quaternion, angular and linear velocity."""

    def __init__(self, world, host, name):
        self._world, self._model, self._host = world, None, host
        self._queried, self._poses, self._state = False, None, None
        self._offset = (0.0, 0.0, 0.0)
        self._backend = None                 # stepped with its host
        self.root_part = Thingy(self, 0, name)
        self.parts, self.joints = [self.root_part], []

    def _refresh(self):
        c = np.asarray(self._host._backend.get_state())[0, -13:]
        row = np.zeros((1, 10), dtype=np.float64)
        row[0, 0:7] = c[0:7]
        row[0, 7:10] = c[10:13]
        self._poses = row
        self._queried = True

    def _contact_list(self, b):
        return []

    def set_pose_and_speed(self, p, vx, vy, vz):
        """World::robot_move (physics-bullet.cpp:579-591): pose in doubles, linear velocity through float, angular velocity
        zeroed by the pose command; the box's cached contact points go with the old pose."""
        h = self._host
        s = np.array(h._backend.get_state(), dtype=np.float64)
        c0 = s.shape[1] - 13
        s[0, c0:c0 + 3] = [p.x, p.y, p.z]
        s[0, c0 + 3:c0 + 7] = [p.qx, p.qy, p.qz, p.qw]
        s[0, c0 + 7:c0 + 10] = 0.0
        s[0, c0 + 10:c0 + 13] = [float(np.float32(vx * SCALE)), float(np.float32(vy * SCALE)), float(np.float32(vz * SCALE))]
        h._backend.set_state(s, clear_contacts=False)
        h._state = None
        self._poses = None


class Camera(object):
    """Inert stub (renderer out of scope)."""

    def __init__(self, name=""):
        self.name = name

    def __getattr__(self, item):
        def _noop(*a, **k):
            return None
        return _noop


class World(object):
    """python-binding.cpp:297-573; physics-bullet.cpp:12-131, 358-421."""

    _model_cache = {}

    def __init__(self, gravity, timestep):
        self._gravity = float(np.float32(gravity))
        self._timestep = float(np.float32(timestep))
        self._raw_timestep = timestep
        self._last_repeat = 0
        self.ts = 0.0
        self._robots = []
        self._floor = None

    # ---- scene management ----
    def clean_everything(self):
        for r in self._robots:
            if r._backend is not None and hasattr(r._backend, "close"):
                r._backend.close()
        self._robots = []
        self._floor = None
        self.ts = 0.0

    def load_mjcf(self, fn):
        try:
            import xml.etree.ElementTree as ET
            root = ET.parse(fn).getroot()
            wb = root.find("worldbody")
            bodies = [c for c in wb if c.tag == "body"]
            if not bodies:
                # world-level geoms only (ground_plane.xml): one static robot named after the geom
                planes = [c for c in wb if c.tag == "geom" and c.attrib.get("type") == "plane"]
                name = planes[0].attrib.get("name", "floor") if planes else "static"
                r = Robot(self, None, static_name=name)
                if planes:
                    self._floor = r
                self._robots.append(r)
                return [r]
            key = (os.path.abspath(fn), self._floor is not None, self._gravity, self._timestep)
            m = World._model_cache.get(key)
            if m is None:
                m = mjcf.compile_mjcf(fn, has_floor=self._floor is not None, gravity=self._gravity,
                                      timestep=self._raw_timestep, frame_skip=1)
                m.max_contacts = 16 if m.n_links > 6 else 8
                World._model_cache[key] = m
            r = Robot(self, m)
            self._robots.append(r)
            return [r]
        except Exception as e:   # reference behaviour: message on stderr, empty list
            sys.stderr.write("'%s': cannot load MJCF. (%s)\n" % (fn, e))
            return []

    def load_urdf(self, fn, pose, fixed_base, self_collision):
        """physics-bullet.cpp:70-93.  Supported: a free single-box URDF (FlagrunHarder's cube.urdf) next to a floating-base robot --
        it becomes the second free body of that robot's world; the host's backend is rebuilt on the cube topology with its state
        carried over.  Everything else (Atlas) fails the way the reference does: a message on stderr, an empty robot."""
        try:
            box = _parse_single_box_urdf(fn)
            host = next((r for r in reversed(self._robots) if r._backend is not None and not r._model.fixed_base
                         and not r._model.has_cube), None)
            if box is None or fixed_base or host is None:
                raise ValueError("only a free single-box URDF next to a floating-base robot is supported")
            half, mass, name = box
            host._attach_box(half, mass, pose)
            r = _FreeBoxRobot(self, host, name)
            self._robots.append(r)
            return r
        except Exception as e:
            sys.stderr.write("Cannot load URDF file '%s'. (%s)\n" % (fn, e))
            return Robot(self, None, static_name="")

    def load_sdf(self, fn):
        sys.stderr.write("'%s': cannot load SDF.\n" % fn)
        return []

    def load_thingy(self, fn, pose, scale, mass, color, decoration_only):
        return Thingy(Robot(self, None, static_name=os.path.basename(fn)), 0, os.path.basename(fn))

    # ---- stepping ----
    def step(self, repeat):
        """World::bullet_step(skip_frames) + query_positions()."""
        for r in self._robots:
            if r._backend is not None:
                r._backend.substeps(int(repeat), r._torques())
        self._last_repeat = int(repeat)
        self.ts += self._timestep * repeat
        for r in self._robots:
            r._refresh()
        return False

    # ---- inert rendering / HUD surface ----
    def new_camera_free_float(self, w, h, name):
        return Camera(name)

    def set_glsl_path(self, p):
        pass

    def set_key_callback(self, cb):
        pass

    def test_window(self):
        return True

    def __getattr__(self, item):
        if item.startswith("test_window") or item.startswith("debug_"):
            def _noop(*a, **k):
                return None
            return _noop
        raise AttributeError(item)
