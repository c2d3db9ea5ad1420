"""Turn the ncu outputs of a gpurun call (gpurun_out/) into the small tracked summaries under profiles/.
    python tools/summarize_profiles.py <launches.csv> <report.ncu-rep> <tag>"""
import collections, csv, io, os, re, subprocess, sys

launch_csv, rep, tag = sys.argv[1:4]
out = []
# ---- launch list: per-kernel share of the benchmark command
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10] if launch_csv != "-" else [["Kernel Name", "Metric Name", "Metric Value"]]
hdr = rows[0]
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", d["Kernel Name"])
    t[k][0] += 1
    t[k][1] += float(d["Metric Value"]) / 1e6
tot = sum(v[1] for v in t.values()) or 1.0
out.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), %s" % launch_csv)
out.append("")
out.append("| kernel | launches | total ms | share |")
out.append("|---|---|---|---|")
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005:
        continue
    out.append("| `%s` | %d | %.3f | %.1f %% |" % (k[:90], v[0], v[1], 100 * v[1] / tot))
out.append("")

# ---- full capture of the top kernel
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.max.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "launch__stack_size"]
for r in rr[2:]:
    d = dict(zip(h, r))
    out.append("## `ncu --set full` of `%s`, %s" % (d["Kernel Name"][:80], rep))
    out.append("")
    out.append("| metric | value | unit |")
    out.append("|---|---|---|")
    for k in keys:
        if k in d:
            out.append("| %s | %s | %s |" % (k, d[k], units[h.index(k)]))
    out.append("")

# ---- source page: opcode mix and stall reasons
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
sr = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(sr) if r and r[0] == "Address"][0]
sh = sr[hi]; ix = {k: i for i, k in enumerate(sh)}
byop, execs, stall = collections.Counter(), collections.Counter(), collections.Counter()
scols = [k for k in sh if k.startswith("stall_") and "Not Issued" not in k]
for r in sr[hi + 1:]:
    if len(r) < len(sh):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    cls = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else "IMAD.HI" if op.startswith("IMAD.HI") else "MOV" if op.startswith("IMAD.MOV") or op == "MOV" else op.split(".")[0]
    byop[cls] += int(r[ix["# Samples"]] or 0)
    execs[cls] += int(r[ix["Instructions Executed"]] or 0)
    for k in scols:
        stall[k] += int(r[ix[k]] or 0)
T, E = sum(byop.values()), sum(execs.values())
out.append("### Warp-sample and executed-instruction mix by opcode class (source page)")
out.append("")
out.append("| opcode class | warp samples | executed warp-instructions |")
out.append("|---|---|---|")
for k, v in byop.most_common(12):
    out.append("| %s | %.2f %% | %.2f %% (%d) |" % (k, 100 * v / T, 100 * execs[k] / E, execs[k]))
out.append("")
out.append("| stall reason | share of samples |")
out.append("|---|---|")
for k, v in stall.most_common(8):
    out.append("| %s | %.2f %% |" % (k, 100 * v / T))
out.append("")
open(os.path.join(os.environ.get("PROFILE_OUT_DIR", "profiles"), "%s.md" % tag), "w").write("# ncu summary %s\n\n" % tag + "\n".join(out) + "\n")
print("\n".join(out))
