"""Emit roboschool_b200/csrc/rs_topo_gen.h: compile-time topologies + constants of the shipped models.

The step kernel is specialised per model: tree loops are fully unrolled against these constexpr
tables, so parent/dof indices, joint axes and zero-rotation matrices become immediates and the tree
sweeps keep their accumulators in registers.  Models loaded at run time through load_mjcf() use the
generic (runtime-topology) kernel instead.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from roboschool_b200 import envspec

OUT = os.path.join(os.path.dirname(__file__), "..", "roboschool_b200", "csrc", "rs_topo_gen.h")


def f(x):
    v = float(np.float32(x))
    if v == 0.0:
        return "0.0f"
    s = "%.9g" % v
    if "." not in s and "e" not in s and "inf" not in s:
        s += ".0"
    return s + "f"


def arr(name, vals, ctype="float", fmt=f):
    return "    static constexpr %s %s[%d] = {%s};\n" % (ctype, name, max(len(vals), 1), ", ".join(fmt(v) for v in vals) if len(vals) else "0")


def quat_to_mat(q):
    x, y, z, w = q
    return [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]


def snap(v):
    """Exact zeros/ones survive float noise so the compiler can fold them."""
    out = []
    for x in v:
        if abs(x) < 1e-12: x = 0.0
        elif abs(x - 1) < 1e-12: x = 1.0
        elif abs(x + 1) < 1e-12: x = -1.0
        out.append(x)
    return out


def emit(env_id, cname):
    m = envspec.load_env_model(env_id)
    n = m.n_links
    s = "// %s (%s)\nstruct %s {\n" % (env_id, m.meta["xml"], cname)
    s += "    static constexpr int NL = %d, ND = %d, FIXED_BASE = %d;\n" % (n, m.n_dof, m.fixed_base)
    s += "    static constexpr int CUBE = 0;   // 1: a free cube (rs_model_desc.has_cube) steps next to the robot\n"
    I = lambda v: str(int(v))
    s += arr("parent", [m.parent[i] for i in range(n)], "int", I)
    s += arr("jtype", [m.jtype[i] for i in range(n)], "int", I)
    s += arr("dof", [m.dof_idx[i] for i in range(n)], "int", I)
    # first_fold[i]: 1 if link i is the highest-index child of its parent (first to fold in the leaves->root sweep)
    first = []
    for i in range(n):
        sib = [j for j in range(n) if m.parent[j] == m.parent[i]]
        first.append(1 if i == max(sib) else 0)
    s += arr("first_fold", first, "int", I)
    leaf = [0 if any(m.parent[j] == i for j in range(n)) else 1 for i in range(n)]
    s += arr("leaf", leaf, "int", I)
    base_leaf = 0 if any(m.parent[j] == -1 for j in range(n)) else 1
    s += "    static constexpr int base_leaf = %d;\n" % base_leaf
    R0, ax, e, d, sw, sv, inert = [], [], [], [], [], [], []
    for i in range(n):
        R0 += snap(quat_to_mat(list(m.zero_rot[i])))
        a = snap(list(m.axis[i])); dd = snap(list(m.d_vec[i]))
        ax += a; e += snap(list(m.e_vec[i])); d += dd
        if m.jtype[i] == 1:
            sw += a; sv += snap(list(np.cross(a, dd)))
        elif m.jtype[i] == 2:
            sw += [0, 0, 0]; sv += a
        else:
            sw += [0, 0, 0]; sv += [0, 0, 0]
        inert += list(m.inertia[i])
    s += arr("R0", R0) + arr("axis", ax) + arr("e", e) + arr("d", d) + arr("Sw", sw) + arr("Sv", sv)
    s += arr("mass", [m.mass[i] for i in range(n)]) + arr("inertia", inert)
    s += "    static constexpr float base_mass = %s;\n" % f(m.base_mass)
    s += arr("base_inertia", list(m.base_inertia))
    # collision geoms / self-collision pairs (static broadphase, rs_static.cuh static_collide)
    ng, npair = m.n_geoms, m.n_pairs
    s += "    static constexpr int NG = %d, NP = %d;\n" % (ng, npair)
    s += arr("g_link", [m.g_link[g] for g in range(ng)], "int", I)
    s += arr("g_type", [m.g_type[g] for g in range(ng)], "int", I)
    s += arr("g_floor", [m.g_floor[g] for g in range(ng)], "int", I)
    s += arr("g_p0", snap([m.g_p0[g][k] for g in range(ng) for k in range(3)]))
    s += arr("g_p1", snap([m.g_p1[g][k] for g in range(ng) for k in range(3)]))
    s += arr("g_radius", [m.g_radius[g] for g in range(ng)])
    s += arr("g_thresh", [m.break_thresh[m.g_link[g] + 1] for g in range(ng)])
    def bound(g):
        d = np.array([m.g_p1[g][k] - m.g_p0[g][k] for k in range(3)])
        return float(np.float32(0.5 * np.sqrt((d * d).sum()) + m.g_radius[g]) * np.float32(1.0001) + np.float32(1e-6))
    s += arr("pair_a", [m.pair_a[k] for k in range(npair)], "int", I)
    s += arr("pair_b", [m.pair_b[k] for k in range(npair)], "int", I)
    reach = []
    for k in range(npair):
        a, b = m.pair_a[k], m.pair_b[k]
        thr = min(m.break_thresh[m.g_link[a] + 1], m.break_thresh[m.g_link[b] + 1])
        reach.append(float(np.float32(bound(a)) + np.float32(bound(b)) + np.float32(thr)))
    s += arr("pair_reach", reach)
    s += "};\n\n"
    sig = "%d:%d:%d:" % (n, m.n_dof, m.fixed_base) + ",".join(str(m.parent[i]) for i in range(n))
    return s, sig


MODELS = [("RoboschoolHumanoid-v1", "TopoHumanoid"), ("RoboschoolAnt-v1", "TopoAnt"), ("RoboschoolHopper-v1", "TopoHopper"),
          ("RoboschoolWalker2d-v1", "TopoWalker2d"), ("RoboschoolHalfCheetah-v1", "TopoHalfCheetah")]

if __name__ == "__main__":
    out = "// GENERATED by tools/gen_topology.py from roboschool_b200/models/*.json -- do not edit.\n#pragma once\nnamespace rs {\n\n"
    for env_id, cname in MODELS:
        s, sig = emit(env_id, cname)
        out += s
    out += ("// RoboschoolHumanoidFlagrunHarder-v1: the Humanoid plus the thrown cube (gym_humanoid_flagrun.py:81-118)\n"
            "struct TopoHumanoidCube : TopoHumanoid {\n    static constexpr int CUBE = 1;\n};\n\n")
    out += "}  // namespace rs\n"
    open(OUT, "w").write(out)
    print("wrote", OUT, len(out), "bytes")
