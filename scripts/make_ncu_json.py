#!/usr/bin/env python
"""profiles/<tag>.json from a full ncu capture of the pair kernel: DRAM bytes per trajectory and the headline counters, tied to
the kernel sources by the sha256 bench.py checks.   usage: make_ncu_json.py <tag> <prof.ncu-rep> <batch> [horizon]"""
import csv, hashlib, json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
horizon = int(sys.argv[4]) if len(sys.argv) > 4 else 10
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
def val(k, scale_unit=True):
    i = hdr.index(k)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    if scale_unit:
        v *= {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "ghz": 1.0, "mhz": 1e-3}.get(u, 1.0)
    return v
h = hashlib.sha256()
d = os.path.join(REPO, "forging_control_b200", "csrc")
for f in sorted(os.listdir(d)):
    h.update(open(os.path.join(d, f), "rb").read())
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {"file": "profiles/r02b_ncu_pair_kernel.md",
       "kernel": "fc::mpc_loss_pair_kernel", "batch": batch, "horizon": horizon,
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_trajectory": (rd + wr) / batch,
       "duration_ms": val("gpu__time_duration.sum"), "sm_ghz": val("sm__cycles_elapsed.avg.per_second"),
       "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
       "tensor_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
       "dram_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
       "long_scoreboard_per_issue": val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", False),
       "warp_instructions": val("smsp__inst_executed.sum", False),
       "source_hash": h.hexdigest()[:16],
       "capture": "ncu --set full --clock-control none --import-source on, python bench.py --steps 2 --warmup 3 --no-cpu-baseline "
                  f"--no-closed-loop --no-parity --batch-per-gpu {batch} (one pass: 148 CTAs x 2 tiles); scripts/ncu_captures.sh"}
path = os.path.join(REPO, "profiles", f"{tag}.json")
json.dump(out, open(path, "w"), indent=1)
print(open(path).read())
