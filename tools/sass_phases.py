"""Attribute SASS instructions (and, optionally, ncu per-instruction samples) of a step kernel to source phases.

  python tools/sass_phases.py <disasm from nvdisasm -g -c> <kernel-name substring> [ncu source-page csv]

Lines of the small-math helpers (rs_core.cuh top part) are ambiguous after inlining, so an instruction inherits
the phase of the closest preceding instruction whose line is unambiguous."""
import csv, re, sys, collections

dis, kname = sys.argv[1], sys.argv[2]
src_csv = sys.argv[3] if len(sys.argv) > 3 else None

import os
CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "roboschool_b200", "csrc")

def _anchors(fname, items):
    """[(first line of the text anchor, phase)] sorted; phase None = ambiguous helper region."""
    lines = open(os.path.join(CSRC, fname)).read().split("\n")
    out = []
    for text, nth, phase in items:
        hits = [i + 1 for i, l in enumerate(lines) if text in l]
        if len(hits) > nth:
            out.append((hits[nth], phase))
    return sorted(out)

RANGES = {
    "rs_kernels.cuh": [(1, "load/store/kernel-glue")],
    "rs_static.cuh": _anchors("rs_static.cuh", [
        ("namespace rs {", 0, None), ("RS_HD void static_body_geoms", 0, "kinematics"), ("RS_HD void static_collide", 0, "collide_broadphase"),
        ("RS_HD void static_aba", 0, "aba"), ("RS_HD void static_calc_state", 0, "calc_state")]),
    "rs_pooled.cuh": _anchors("rs_pooled.cuh", [
        ("namespace rs {", 0, None), ("RS_HD void pooled_response", 0, "resp_up"), ("float acc[NR][n + 1][6];", 0, "resp_root"),
        ("static_for<n>([&](auto ic) {", 0, "resp_down"), ("RS_HD void resolve_pooled", 0, "pgs_row"),
        ("RS_HD void solve_pooled", 0, "solve_owner_setup"), ("// ---- all lanes: limit-row tasks ----", 0, "solve_tasks_glue"),
        ("// ---- owner: right-hand sides", 0, "solve_owner_rhs+pgs"), ("RS_HD void substep_pooled", 0, "substep-glue")]),
    "rs_core.cuh": _anchors("rs_core.cuh", [
        ("namespace rs {", 0, None), ("RS_HD void kinematics(", 0, "generic_kin"), ("RS_HD void clamp_vel", 0, "clamp_vel"),
        ("RS_HD void aba(", 0, "generic_aba"), ("RS_HD void to_world", 0, None), ("RS_HD void seg_seg_closest", 0, "collide_narrow"),
        ("RS_HD void resolve_row", 0, "generic_solve"), ("RS_HD void integrate", 0, "integrate"), ("RS_HD void mat_to_quat", 0, "epilogue"),
        ("RS_HD void env_reset_with", 0, "reset"), ("RS_HD void apply_action", 0, "apply_action")]),
}

def phase_of(f, ln):
    f = f.split("/")[-1]
    ph = None
    for start, name in RANGES.get(f, []):
        if ln >= start:
            ph = name
    return ph

cur_fn, phase, in_fn = None, "prologue", False
count = collections.Counter()
addr_phase = {}
pending = None
for line in open(dis):
    if line.startswith(".text."):
        in_fn = kname in line
        phase = "prologue"
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        p = phase_of(m.group(1), int(m.group(2)))
        if p: phase = p
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(\S.*);', line)
    if m:
        count[phase] += 1
        addr_phase[int(m.group(1), 16)] = phase
tot = sum(count.values())
print("SASS instructions by phase (code size, 16 B each): total %d = %.0f KB" % (tot, tot * 16 / 1024))
for k, v in count.most_common():
    print("  %-28s %7d  %5.1f%%" % (k, v, 100.0 * v / tot))
if src_csv:
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    ia, isamp, iex, inoi, ilsb = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("stall_no_inst"), h.index("stall_long_sb")
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for r in rows[hi + 1:]:
        try: a = int(r[ia], 16)
        except ValueError: continue
        if base is None: base = a
        ph = addr_phase.get(a - base, "?")
        g = agg[ph]
        g[0] += int(r[isamp] or 0); g[1] += int(r[iex] or 0); g[2] += int(r[inoi] or 0); g[3] += int(r[ilsb] or 0)
    ts = sum(g[0] for g in agg.values()); te = sum(g[1] for g in agg.values())
    print("\nphase                         samples%%  executed%%  no_inst%%(of phase)  long_sb%%(of phase)")
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("  %-28s %6.1f    %6.1f      %6.1f            %6.1f" % (k, 100.0 * g[0] / ts, 100.0 * g[1] / te, 100.0 * g[2] / max(g[0], 1), 100.0 * g[3] / max(g[0], 1)))
