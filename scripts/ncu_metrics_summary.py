"""Turns the long-format CSV of `ncu --metrics ... --csv --log-file X.csv` (one row per kernel launch and metric) into
profiles/r2_traffic.json: per kernel (first captured launch) duration, executed warp instructions, active threads per
instruction, issue-slot utilisation, occupancy, shared-memory bank conflicts and DRAM bytes.
Usage: python scripts/ncu_metrics_summary.py gpurun_out/ncu_light_TAG.csv profiles/r2_traffic.json"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
col = {name: h.index(name) for name in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launches = {}
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    launches.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]]})[r[col["Metric Name"]]] = (r[col["Metric Value"]], r[col["Metric Unit"]])
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}
out = {}
for lid, m in launches.items():
    name = re.sub(r"^void\s+", "", m["name"])
    name = re.sub(r"<unnamed>::", "", name)
    name = re.match(r"[A-Za-z0-9_]+", name).group(0)
    if name in out:
        continue

    def val(metric, scale_units=False):
        if metric not in m:
            return None
        v, u = m[metric]
        v = float(v.replace(",", ""))
        return v * SCALE.get(u, 1) if scale_units else v

    rd, wr = val("dram__bytes_read.sum", True), val("dram__bytes_write.sum", True)
    out[name] = {
        "ncu_duration_us": val("gpu__time_duration.sum", True),
        "warp_instructions_per_launch": int(val("smsp__inst_executed.sum") or 0),
        "threads_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct_elapsed": val("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "shared_bank_conflicts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "dram_bytes_per_launch": int(rd + wr) if rd is not None and wr is not None else None,
    }
out["_source"] = ("ncu --metrics (gpu__time_duration, smsp__inst_executed, thread_inst_executed_per_inst_executed, "
                  "issue_active, warps_active, shared bank conflicts, dram bytes) --clock-control none, C2 workload "
                  "(scripts/profile_step.py, steady-state step), first captured launch per kernel; csv: " + sys.argv[1])
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
