"""Attribute SASS instructions (and, optionally, ncu per-instruction samples) of a step kernel to source phases.

  python tools/sass_phases.py <disasm from nvdisasm -g -c> <kernel-name substring> [ncu source-page csv]

Lines of the small-math helpers (rs_core.cuh top part) are ambiguous after inlining, so an instruction inherits
the phase of the closest preceding instruction whose line is unambiguous."""
import csv, re, sys, collections

dis, kname = sys.argv[1], sys.argv[2]
src_csv = sys.argv[3] if len(sys.argv) > 3 else None

def phase_of(f, ln):
    f = f.split("/")[-1]
    if f == "rs_kernels.cuh":
        return "load/store/epilogue-glue"
    if f == "rs_static.cuh":
        if ln < 100: return None
        if ln < 131: return "kinematics"
        if ln < 290: return "aba"
        if ln < 385: return "old_response"
        return "calc_state"
    if f == "rs_pooled.cuh":
        if ln < 98: return None
        if ln < 166: return "resp_up"
        if ln < 205: return "resp_root"
        if ln < 235: return "resp_down"
        if ln < 250: return "resolve_row"
        if ln < 330: return "solve_owner_setup"
        if ln < 372: return "solve_tasks_glue"
        return "solve_owner_pgs"
    if f == "rs_core.cuh":
        if ln < 276: return None
        if ln < 346: return "generic_kin"
        if ln < 555: return "clamp/aba_generic"
        if ln < 762: return "collide"
        if ln < 905: return "solve_generic"
        if ln < 926: return "integrate"
        if ln < 1102: return "epilogue"
        return "reset/apply_action"
    return None

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
