#!/usr/bin/env python
"""Turn gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and a full-capture .ncu-rep into the
markdown summaries kept under profiles/.   usage: summarize_ncu.py <tag> [launches.csv] [prof.ncu-rep]"""
import collections, csv, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, "gpurun_out", "launches.csv")
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(REPO, "gpurun_out", "prof_mpc.ncu-rep")
out = [f"# ncu summary {tag}\n"]
if os.path.isfile(launches):
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki][:90], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out.append("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised: compare SHARES)\n")
    out.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        out.append(f"| `{k}` | {v[0]} | {v[1]/1e6:.3f} | {100*v[1]/tot:.1f}% |")
    out.append("")
if os.path.isfile(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
            "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append(f"## full capture (`ncu --set full --clock-control none --import-source on`): `{name[:80]}`\n")
        out.append("| metric | unit | value |\n|---|---|---:|")
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"| {k} | {units[i]} | {r[i]} |")
        out.append("")
path = os.path.join(REPO, "profiles", f"{tag}.md")
open(path, "w").write("\n".join(out) + "\n")
print(open(path).read())
