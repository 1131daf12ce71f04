"""MJCF -> flat model descriptor ("model compiler").

Restates what the reference gets from its Bullet dependency when it calls
``b3LoadMJCFCommandInit(... URDF_USE_SELF_COLLISION|URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS)``
(roboschool/cpp-household/physics-bullet.cpp:95-131) followed by
``b3GetJointInfo`` (physics-bullet.cpp:206-251).

The importer that does this lives in the un-vendored dependency
``github.com/olegklimov/bullet3@roboschool_self_collision`` (Bullet 2.87,
install_bullet.sh:8-11).  The conventions below are RECALLED from upstream
``examples/Importers/ImportMJCFDemo/BulletMJCFImporter.cpp`` and
``ImportURDFDemo/URDF2Bullet.cpp``; each is a line of data in the descriptor, and
the knobs are keyword arguments so they can be re-pinned against a real Bullet
build:

 * every body with k>=1 joints becomes a chain of k one-dof links; the first k-1
   are massless "dummy" links located at the joint, the last carries mass+geoms;
 * a top-level body WITH joints hangs from a massless fixed base at the world
   origin; WITHOUT joints it is a floating base;
 * a jointless non-root body is a fixed (0-dof) link;
 * mass = 1000 kg/m^3 * sum(geom volumes) unless ``<inertial mass=..>``; the
   ``density`` attribute is not honoured (``honour_density=False``);
 * the inertial (COM) frame has identity rotation and sits at the centre of the
   LAST geom of the body if that geom is a ``fromto`` capsule, else at the body
   frame origin ("inertialShift");
 * inertia = box inertia of the compound collision shape's AABB
   (btCompoundShape::calculateLocalInertia); fromto capsules are
   btMultiSphereShape whose cached AABB carries one extra 0.04 margin;
 * joint ``range`` is converted with ``compiler angle``; damping/stiffness/
   armature/ref are ignored (``honour_damping=False``); box geoms (btBoxShape) collide with the floor plane only;
 * group<-contype, mask<-conaffinity with MuJoCo OR semantics; self-collision
   between links that are not ancestor/descendant.
"""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

from .model import (ModelDesc, RS_MAX_LINKS, RS_MAX_GEOMS, RS_MAX_PAIRS, RS_MAX_ACT,
                    RS_MAX_FEET, RS_MAX_PARTS, RS_JOINT_FIXED, RS_JOINT_REVOLUTE,
                    RS_JOINT_PRISMATIC, RS_GEOM_SPHERE, RS_GEOM_CAPSULE, RS_GEOM_BOX)

CONVEX_DISTANCE_MARGIN = 0.04      # btCollisionMargin.h
CONTACT_BREAKING_THRESHOLD = 0.02  # gContactBreakingThreshold


def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(v) != n:
        raise ValueError("expected %d numbers, got %r" % (n, s))
    return v


def quat_wxyz_to_mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def mat_to_quat_xyzw(R):
    """Rotation matrix -> quaternion (x,y,z,w), w>=0."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s
        y = (R[0, 2] - R[2, 0]) / s
        z = (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w = (R[2, 1] - R[1, 2]) / s
        x = 0.25 * s
        y = (R[0, 1] + R[1, 0]) / s
        z = (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w = (R[0, 2] - R[2, 0]) / s
        x = (R[0, 1] + R[1, 0]) / s
        y = 0.25 * s
        z = (R[1, 2] + R[2, 1]) / s
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w = (R[1, 0] - R[0, 1]) / s
        x = (R[0, 2] + R[2, 0]) / s
        y = (R[1, 2] + R[2, 1]) / s
        z = 0.25 * s
    q = np.array([x, y, z, w])
    if w < 0:
        q = -q
    return q / np.linalg.norm(q)


def axisangle_to_mat(aa, degrees):
    axis = aa[:3] / np.linalg.norm(aa[:3])
    ang = math.radians(aa[3]) if degrees else aa[3]
    x, y, z = axis
    c, s = math.cos(ang), math.sin(ang)
    C = 1 - c
    return np.array([
        [c + x * x * C, x * y * C - z * s, x * z * C + y * s],
        [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
        [z * x * C - y * s, z * y * C + x * s, c + z * z * C]])


class _Geom:
    __slots__ = ("name", "type", "p0", "p1", "radius", "fromto", "contype", "conaffinity",
                 "friction", "density", "volume", "rot", "halflen", "center", "half")


class _Joint:
    __slots__ = ("name", "type", "axis", "pos", "lo", "hi", "limited", "damping", "armature", "stiffness")


class _Body:
    __slots__ = ("name", "pos", "rot", "joints", "geoms", "children", "mass_override", "parent")


def _merged(defaults, tag, elem):
    d = dict(defaults.get(tag, {}))
    d.update(elem.attrib)
    return d


def _parse_geom(elem, defaults, degrees, honour_axisangle=True):
    a = _merged(defaults, "geom", elem)
    g = _Geom()
    g.name = a.get("name", "")
    t = a.get("type", "sphere")
    size = _vec(a.get("size"))
    pos = _vec(a.get("pos"), 3, [0, 0, 0])
    rot = np.eye(3)
    if "quat" in a:
        rot = quat_wxyz_to_mat(_vec(a["quat"], 4))
    elif "axisangle" in a and honour_axisangle:
        rot = axisangle_to_mat(_vec(a["axisangle"], 4), degrees)
    g.contype = int(a.get("contype", 1))
    g.conaffinity = int(a.get("conaffinity", 1))
    g.friction = float(a.get("friction", "1 0.005 0.0001").split()[0])
    g.density = float(a.get("density", 1000.0))
    g.rot = rot
    g.fromto = False
    g.halflen = 0.0
    if t == "sphere":
        g.type = RS_GEOM_SPHERE
        g.radius = float(size[0])
        g.p0 = pos.copy()
        g.p1 = pos.copy()
        g.volume = 4.0 / 3.0 * math.pi * g.radius ** 3
    elif t == "capsule":
        g.type = RS_GEOM_CAPSULE
        g.radius = float(size[0])
        if "fromto" in a:
            ft = _vec(a["fromto"], 6)
            g.p0, g.p1 = ft[:3].copy(), ft[3:].copy()
            g.fromto = True
            h = float(np.linalg.norm(g.p1 - g.p0))
        else:
            g.halflen = float(size[1])
            g.p0 = pos + rot @ np.array([0, 0, -g.halflen])
            g.p1 = pos + rot @ np.array([0, 0, +g.halflen])
            h = 2 * g.halflen
        g.volume = math.pi * g.radius ** 2 * h + 4.0 / 3.0 * math.pi * g.radius ** 3
    elif t == "box":
        # btBoxShape(half extents) under the geom's own transform; touches the floor plane only
        if "fromto" in a:
            raise NotImplementedError("box geom with fromto (model %s)" % g.name)
        g.type = RS_GEOM_BOX
        g.radius = 0.0
        g.half = np.array([float(size[0]), float(size[1]), float(size[2])])
        g.p0 = pos.copy()
        g.p1 = pos.copy()
        g.volume = 8.0 * float(np.prod(g.half))
    elif t == "plane":
        return None
    else:
        raise NotImplementedError("geom type %r (model %s)" % (t, g.name))
    g.center = 0.5 * (g.p0 + g.p1)
    return g


def _parse_joint(elem, defaults, degrees):
    a = _merged(defaults, "joint", elem)
    j = _Joint()
    j.name = a.get("name", "")
    t = a.get("type", "hinge")
    if t == "hinge":
        j.type = RS_JOINT_REVOLUTE
    elif t == "slide":
        j.type = RS_JOINT_PRISMATIC
    else:
        raise NotImplementedError("joint type %r" % t)
    ax = _vec(a.get("axis"), 3, [0, 0, 1])
    j.axis = ax / np.linalg.norm(ax)
    j.pos = _vec(a.get("pos"), 3, [0, 0, 0])
    j.limited = a.get("limited", "false") == "true" and "range" in a
    lo, hi = 0.0, -1.0   # Bullet "no limit" convention: lower > upper
    if "range" in a:
        r = _vec(a["range"], 2)
        lo, hi = float(r[0]), float(r[1])
        if degrees and j.type == RS_JOINT_REVOLUTE:
            lo, hi = math.radians(lo), math.radians(hi)
    if not j.limited:
        lo, hi = 0.0, -1.0
    j.lo, j.hi = lo, hi
    j.damping = float(a.get("damping", 0.0))
    j.armature = float(a.get("armature", 0.0))
    j.stiffness = float(a.get("stiffness", 0.0))
    return j


def _parse_body(elem, defaults, degrees, parent=None, honour_axisangle=True):
    b = _Body()
    b.name = elem.attrib.get("name", "")
    b.pos = _vec(elem.attrib.get("pos"), 3, [0, 0, 0])
    b.rot = quat_wxyz_to_mat(_vec(elem.attrib["quat"], 4)) if "quat" in elem.attrib else np.eye(3)
    b.parent = parent
    b.joints, b.geoms, b.children = [], [], []
    b.mass_override = None
    for ch in elem:
        if ch.tag == "joint":
            b.joints.append(_parse_joint(ch, defaults, degrees))
        elif ch.tag == "geom":
            g = _parse_geom(ch, defaults, degrees, honour_axisangle)
            if g is not None:
                b.geoms.append(g)
        elif ch.tag == "inertial":
            if "mass" in ch.attrib:
                b.mass_override = float(ch.attrib["mass"])
        elif ch.tag == "body":
            b.children.append(_parse_body(ch, defaults, degrees, b, honour_axisangle))
    return b


def _geom_aabb(g, com, margin=CONVEX_DISTANCE_MARGIN):
    """AABB of one child shape in the link COM frame, as Bullet's getAabb reports it."""
    if g.type == RS_GEOM_SPHERE:
        c = g.p0 - com
        return c - g.radius, c + g.radius
    if g.type == RS_GEOM_BOX:
        # btBoxShape::getAabb: |R| (half extents without margin + margin)
        ext = np.abs(g.rot) @ g.half
        c = g.center - com
        return c - ext, c + ext
    if g.fromto:
        # btMultiSphereShape: cached local AABB is exact, getAabb adds one more margin
        lo = np.minimum(g.p0, g.p1) - com - g.radius - margin
        hi = np.maximum(g.p0, g.p1) - com + g.radius + margin
        return lo, hi
    # btCapsuleShapeZ under a rotated child transform: OBB -> AABB via |R|
    he = np.array([g.radius, g.radius, g.radius + g.halflen])
    ext = np.abs(g.rot) @ he
    c = g.center - com
    return c - ext, c + ext


def _body_inertial(b, honour_density, com_rule="last_fromto", aabb_margin=CONVEX_DISTANCE_MARGIN):
    """(mass, com_in_body_frame, diag_inertia, break_thresh) for a real (non-dummy) link."""
    if b.geoms and com_rule == "last_fromto":
        last = b.geoms[-1]
        com = last.center.copy() if (last.type == RS_GEOM_CAPSULE and last.fromto) else np.zeros(3)
    elif b.geoms and com_rule == "centroid":
        com = sum(g.volume * g.center for g in b.geoms) / sum(g.volume for g in b.geoms)
    elif b.geoms and com_rule == "first":
        com = b.geoms[0].center.copy()
    else:
        com = np.zeros(3)
    if b.mass_override is not None:
        mass = b.mass_override
    else:
        mass = sum((g.density if honour_density else 1000.0) * g.volume for g in b.geoms)
    if not b.geoms:
        return mass, com, np.zeros(3), CONTACT_BREAKING_THRESHOLD
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for g in b.geoms:
        l, h = _geom_aabb(g, com, aabb_margin)
        lo, hi = np.minimum(lo, l), np.maximum(hi, h)
    ext = hi - lo
    lx, ly, lz = ext
    inertia = mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
    radius = 0.5 * np.linalg.norm(ext)
    center = 0.5 * (lo + hi)
    disc = radius + np.linalg.norm(center)
    return mass, com, inertia, CONTACT_BREAKING_THRESHOLD * disc


def compile_mjcf(xml_path, honour_density=False, honour_damping=False, has_floor=True,
                 gravity=9.8, timestep=0.0165 / 4, frame_skip=4, honour_axisangle=True, com_rule="last_fromto",
                 aabb_margin=CONVEX_DISTANCE_MARGIN, honour_armature=False, honour_stiffness=False,
                 honour_settotalmass=False, contact_erp="split", damping_scale=1.0):
    """Compile one robot MJCF into a :class:`ModelDesc` (physics part only).

    ``gravity``/``timestep`` are squeezed through float32 exactly as the reference's
    ``World(float gravity, float timestep)`` constructor does
    (roboschool/cpp-household/python-binding.cpp:297).
    """
    root = ET.parse(xml_path).getroot()
    comp = root.find("compiler")
    degrees = True   # MuJoCo default angle unit is degree
    if comp is not None and comp.attrib.get("angle", "degree") == "radian":
        degrees = False
    defaults = {}
    d = root.find("default")
    if d is not None:
        for ch in d:
            defaults[ch.tag] = dict(ch.attrib)
    wb = root.find("worldbody")
    bodies = [ch for ch in wb if ch.tag == "body"]
    if len(bodies) == 0:
        raise NotImplementedError("no top-level body")
    tops = [_parse_body(b, defaults, degrees, None, honour_axisangle) for b in bodies]
    if len(tops) > 1 and not all(t.joints for t in tops):
        raise NotImplementedError("several top-level bodies are supported only when each hangs from the world by joints")
    top = tops[0]

    m = ModelDesc()
    links = []   # dicts
    geoms = []   # (link_idx, _Geom, com)
    names_links, names_joints = [], []

    def add_geoms(link_idx, b, com):
        for g in b.geoms:
            geoms.append((link_idx, g, com))

    # ---- base ----
    if top.joints:
        m.fixed_base = 1
        m.base_mass = 0.0
        base_name = "world"
        base_thresh = CONTACT_BREAKING_THRESHOLD
        # anchor frame = world; each top body is attached by its own joints.  (Bullet makes one multibody per top-level
        # body -- reacher.xml: arm and target; hung from one fixed anchor they stay dynamically independent.)
        pending = [(t, -1, np.zeros(3), np.eye(3), True) for t in tops]
        base_pos = np.zeros(3)
        base_rot = np.eye(3)
    else:
        m.fixed_base = 0
        mass, com, inertia, thr = _body_inertial(top, honour_density, com_rule, aabb_margin)
        m.base_mass = mass
        m.arr("base_inertia")[:] = inertia
        base_name = top.name
        base_thresh = thr
        add_geoms(-1, top, com)
        pending = [(ch, -1, com, None, False) for ch in top.children]
        base_pos = top.pos + top.rot @ com
        base_rot = top.rot
    m.arr("break_thresh")[0] = base_thresh

    # depth-first, XML order (== Bullet link order == roboschool joint order)
    def emit(b, parent_link, parent_com, is_top):
        """Create the link chain of body ``b`` whose parent body's real link is
        ``parent_link`` with COM ``parent_com`` (parent body frame)."""
        mass, com, inertia, thr = _body_inertial(b, honour_density, com_rule, aabb_margin)
        pB = b.pos
        R_B = b.rot
        chain = b.joints if b.joints else [None]
        prev_link, prev_pivot_in_B = parent_link, None
        for k, j in enumerate(chain):
            real = (k == len(chain) - 1)
            L = {}
            L["parent"] = prev_link
            if j is None:
                L["jtype"] = RS_JOINT_FIXED
                L["axis"] = np.array([0.0, 0.0, 1.0])
                pj = np.zeros(3)
                L["lo"], L["hi"], L["limited"], L["damping"] = 0.0, -1.0, 0, 0.0
                L["armature"], L["stiffness"] = 0.0, 0.0
                jname = None
            else:
                L["jtype"] = j.type
                L["axis"] = j.axis
                pj = j.pos
                L["lo"], L["hi"], L["limited"] = j.lo, j.hi, int(j.limited)
                L["damping"] = j.damping * damping_scale if honour_damping else 0.0
                L["armature"] = j.armature if honour_armature else 0.0
                L["stiffness"] = j.stiffness if honour_stiffness else 0.0
                jname = j.name
            if k == 0:
                # parent frame = parent body frame; pivot = pB + R_B pj
                L["e"] = pB + R_B @ pj - parent_com
                L["zero_rot"] = mat_to_quat_xyzw(R_B.T)   # parent coords -> this coords
            else:
                L["e"] = pj - prev_pivot_in_B             # previous dummy COM sits at its pivot
                L["zero_rot"] = np.array([0.0, 0.0, 0.0, 1.0])
            if real:
                L["d"] = com - pj
                L["mass"], L["inertia"], L["thresh"] = mass, inertia, thr
                L["name"] = b.name
            else:
                L["d"] = np.zeros(3)
                L["mass"], L["inertia"], L["thresh"] = 0.0, np.zeros(3), CONTACT_BREAKING_THRESHOLD
                L["name"] = "dummy_%s" % jname
            L["jname"] = jname
            links.append(L)
            prev_link = len(links) - 1
            prev_pivot_in_B = pj
        real_idx = len(links) - 1
        add_geoms(real_idx, b, com)
        for ch in b.children:
            emit(ch, real_idx, com, False)

    for (b, pl, pcom, _, is_top) in pending:
        emit(b, pl, pcom, is_top)

    if len(links) > RS_MAX_LINKS:
        raise ValueError("too many links")
    m.n_links = len(links)
    ndof = 0
    for i, L in enumerate(links):
        m.parent[i] = L["parent"]
        m.jtype[i] = L["jtype"]
        if L["jtype"] == RS_JOINT_FIXED:
            m.dof_idx[i] = -1
        else:
            m.dof_idx[i] = ndof
            ndof += 1
        m.has_limit[i] = L["limited"]
        m.arr("axis")[i] = L["axis"]
        m.arr("zero_rot")[i] = L["zero_rot"]
        m.arr("e_vec")[i] = L["e"]
        m.arr("d_vec")[i] = L["d"]
        m.mass[i] = L["mass"]
        m.arr("inertia")[i] = L["inertia"]
        m.lim_lo[i], m.lim_hi[i] = L["lo"], L["hi"]
        m.jdamping[i] = L["damping"]
        m.armature[i] = L["armature"]
        m.stiffness[i] = L["stiffness"]
        m.max_vel[i] = 0.0
        m.arr("break_thresh")[i + 1] = L["thresh"]
    m.n_dof = ndof
    if honour_settotalmass and comp is not None and "settotalmass" in comp.attrib:
        tot = m.base_mass + sum(m.mass[i] for i in range(m.n_links))
        k = float(comp.attrib["settotalmass"]) / tot
        m.base_mass *= k
        m.arr("base_inertia")[:] *= k
        for i in range(m.n_links):
            m.mass[i] *= k
            m.arr("inertia")[i] *= k

    # ---- geoms ----
    if len(geoms) > RS_MAX_GEOMS:
        raise ValueError("too many geoms")
    m.n_geoms = len(geoms)
    floor_contype, floor_conaff = 1, 1   # ground_plane.xml:3 (contype default 1)
    for gi, (li, g, com) in enumerate(geoms):
        m.g_link[gi] = li
        m.g_type[gi] = g.type
        m.arr("g_p0")[gi] = g.p0 - com
        m.arr("g_p1")[gi] = g.half if g.type == RS_GEOM_BOX else g.p1 - com
        m.arr("g_quat")[gi] = mat_to_quat_xyzw(g.rot) if g.type == RS_GEOM_BOX else [0.0, 0.0, 0.0, 1.0]
        m.g_radius[gi] = g.radius
        m.g_friction[gi] = g.friction
        m.g_floor[gi] = int(has_floor and bool((g.contype & floor_conaff) or (floor_contype & g.conaffinity)))
    m.has_floor = int(has_floor)
    m.floor_friction = 0.8                # ground_plane.xml:3

    # ---- self-collision pairs (URDF_USE_SELF_COLLISION | EXCLUDE_ALL_PARENTS) ----
    def ancestors(li):
        out = set()
        while li != -1:
            li = links[li]["parent"]
            out.add(li)
        return out
    anc = {li: ancestors(li) for li in range(len(links))}
    anc[-1] = set()
    pairs = []
    for a in range(len(geoms)):
        for b2 in range(a + 1, len(geoms)):
            la, ga, _ = geoms[a]
            lb, gb, _ = geoms[b2]
            if la == lb or la in anc[lb] or lb in anc[la]:
                continue
            if not ((ga.contype & gb.conaffinity) or (gb.contype & ga.conaffinity)):
                continue
            if ga.type == RS_GEOM_BOX or gb.type == RS_GEOM_BOX:
                continue               # box-vs-convex needs GJK/EPA: only box-vs-plane is built
            pairs.append((a, b2))
    if len(pairs) > RS_MAX_PAIRS:
        raise ValueError("too many self-collision pairs (%d)" % len(pairs))
    m.n_pairs = len(pairs)
    for k, (a, b2) in enumerate(pairs):
        m.pair_a[k], m.pair_b[k] = a, b2

    # ---- world / solver constants ----
    m.gravity = float(np.float32(gravity))
    m.timestep = float(np.float32(timestep))
    m.frame_skip = frame_skip
    m.n_iter = 5                     # physics-bullet.cpp:370
    m.erp2 = 0.9                     # physics-bullet.cpp:371 (contactERP)
    m.erp = 0.2                      # btContactSolverInfo default m_erp
    m.split_thresh = -0.04           # m_splitImpulsePenetrationThreshold
    # contact rows: "split" = m_erp above the split-impulse threshold, positional term dropped below it;
    # "erp2" = contactERP for every penetration, always combined into the rhs
    if contact_erp == "split":
        m.c_erp, m.c_erp2, m.c_split_thresh = m.erp, m.erp2, m.split_thresh
    elif contact_erp == "erp2":
        m.c_erp, m.c_erp2, m.c_split_thresh = m.erp2, m.erp2, -1e30
    else:
        m.c_erp, m.c_erp2, m.c_split_thresh = float(contact_erp), float(contact_erp), -1e30
    m.lin_damping = 0.04             # btMultiBody ctor
    m.ang_damping = 0.04
    m.max_coord_vel = 100.0
    m.max_limit_impulse = 100.0      # btMultiBodyConstraint m_maxAppliedImpulse
    m.use_gyro = 1.0

    # reset defaults: base at its XML pose
    m.arr("reset_base_pos")[:] = base_pos
    m.arr("reset_base_quat")[:] = mat_to_quat_xyzw(base_rot)
    m.reset_joint_spread = 0.0
    m.reset_yaw_spread = 0.0
    m.reset_set_base = 0
    m.alive_rule = 2
    m.body_link = -1
    m.task_link = m.task_link2 = -1       # no task link until a task spec names one (a link-less body has no link 0)
    m.initial_z = -1.0
    m.max_episode_steps = 1000
    m.dt_env = timestep * frame_skip   # python double, scene_abstract.py:22

    m.meta = {
        "xml": os.path.basename(xml_path),
        "base_name": base_name,
        "link_names": [L["name"] for L in links],
        "joint_names": [L["jname"] for L in links],   # None for fixed
        "geom_names": [g.name for (_, g, _) in geoms],
    }
    return m


def compile_mjcf_native(xml_path, has_floor=True, gravity=9.8, timestep=0.0165 / 4, frame_skip=4):
    """Same descriptor through the library's C entry ``rs_model_compile`` (csrc/rs_model_compile.cpp): the compiler a
    native binding uses.  Default conventions only; the keyword knobs of :func:`compile_mjcf` exist for
    the convention sweeps and stay in Python."""
    import ctypes
    from . import _lib
    from .model import DEFINES

    n = 48   # RS_NAME_LEN

    class _Names(ctypes.Structure):
        _fields_ = [("base_name", ctypes.c_char * n), ("link_names", (ctypes.c_char * n) * DEFINES["RS_MAX_LINKS"]),
                    ("joint_names", (ctypes.c_char * n) * DEFINES["RS_MAX_LINKS"]), ("geom_names", (ctypes.c_char * n) * DEFINES["RS_MAX_GEOMS"])]

    m, names = ModelDesc(), _Names()
    _lib.check(_lib.lib().rs_model_compile(os.fsencode(xml_path), gravity, timestep, frame_skip, int(bool(has_floor)),
                                           ctypes.byref(m), ctypes.byref(names)))
    m.max_contacts = 0     # like compile_mjcf: the task layer sets it (envspec.apply_task)
    m.meta = {
        "xml": os.path.basename(xml_path),
        "base_name": names.base_name.decode(),
        "link_names": [names.link_names[i].value.decode() for i in range(m.n_links)],
        "joint_names": [(names.joint_names[i].value.decode() or None) for i in range(m.n_links)],
        "geom_names": [names.geom_names[g].value.decode() for g in range(m.n_geoms)],
    }
    return m
