"""Turns an `ncu -i X.ncu-rep --page raw --csv` dump into profiles/r1_traffic.json (per kernel: DRAM bytes per launch,
duration, issue-slot and DRAM utilisation, occupancy, registers).  Usage: python scripts/ncu_summary.py raw.csv out.json"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
out = {}
for r in rows[2:]:
    g = lambda k: r[h.index(k)] if k in h else None
    name = re.sub(r"^void\s+", "", g("Kernel Name"))
    name = re.sub(r"<unnamed>::", "", name)
    name = re.match(r"[A-Za-z0-9_]+", name).group(0)
    if name in out:
        continue  # first captured launch per kernel
    f = lambda k: float(g(k)) if g(k) not in (None, "") else None
    unit = rows[1][h.index("dram__bytes_read.sum")]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    out[name] = {
        "dram_bytes_per_launch": int((f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * scale),
        "ncu_duration_us": f("gpu__time_duration.sum"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": int(f("launch__registers_per_thread")),
        "warp_instructions": int(f("smsp__inst_executed.sum")),
    }
out["_source"] = ("ncu --set full --clock-control none, C2 workload (scripts/profile_step.py, steady-state step), "
                  "first captured launch per kernel; raw page: " + sys.argv[1])
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
