"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

  python tools/summarize_ncu.py <launches.csv> <prof.ncu-rep> <tag>
writes profiles/<tag>_launches.csv (copy), profiles/<tag>_summary.md and profiles/k_step_traffic.json.
"""
import collections, csv, io, json, os, shutil, subprocess, sys

launches, rep, tag = sys.argv[1:4]
# optional: algorithmic bytes per env step and envs per launch of the capture (default: the Humanoid bench config)
alg_bytes = int(sys.argv[4]) if len(sys.argv) > 4 else 628
n_envs = int(sys.argv[5]) if len(sys.argv) > 5 else 32768
write_traffic = len(sys.argv) <= 4
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
shutil.copy(launches, os.path.join(out_dir, tag + "_launches.csv"))

rows = [r for r in csv.reader(open(launches)) if r]
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hdr_i]
agg = collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0]
    agg.setdefault(name, []).append(float(d["Metric Value"]))
tot = sum(sum(v) for v in agg.values())
md = ["# ncu summary `%s`" % tag, "", "## launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", "",
      "| kernel | launches | mean us | share of step |", "|---|---|---|---|"]
for k, v in agg.items():
    md.append("| `%s` | %d | %.1f | %.3f |" % (k[:70], len(v), sum(v) / len(v) / 1e3, sum(v) / tot))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units, vals = rr[0], rr[1], rr[2]
get = lambda k: vals[h.index(k)] if k in h else None
unit = lambda k: units[h.index(k)] if k in h else ""
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum"]
md += ["", "## `%s` (ncu --set full, one launch in steady state)" % get("Kernel Name")[:80], "", "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if get(k) is not None:
        md.append("| %s | %s | %s |" % (k, get(k), unit(k)))
stalls = []
for i, k in enumerate(h):
    if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
        try:
            stalls.append((float(vals[i]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
stalls.sort(reverse=True)
st = sum(v for v, _ in stalls) or 1.0
md += ["", "| stall reason (pc sampling) | share |", "|---|---|"] + ["| %s | %.1f %% |" % (k, 100 * v / st) for v, k in stalls[:8]]
def to_bytes(k):
    v, u = float(get(k)), unit(k).lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
if write_traffic:    # bench.py reads this file for roofline.traffic of its default (Humanoid) workload
    json.dump({"kernel": get("Kernel Name"), "dram_bytes_per_launch": traffic, "source": os.path.basename(rep), "tag": tag},
              open(os.path.join(out_dir, "k_step_traffic.json"), "w"))
md += ["", "DRAM traffic per launch: %.3f GB (algorithmic: %d B x %d envs = %.4f GB)" % (traffic / 1e9, alg_bytes, n_envs, alg_bytes * n_envs / 1e9)]
open(os.path.join(out_dir, tag + "_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md))
