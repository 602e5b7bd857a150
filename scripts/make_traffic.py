"""profiles/traffic.json from an `ncu --set full` capture of the dominant kernels (one launch each of the bench
workload): DRAM bytes per launch and the facts bench.py prints beside the roofline fraction, stamped with the sha1 of
the kernel sources the capture was taken from (bench.kernel_source_sha) -- bench.py refuses a stale file.

    python scripts/make_traffic.py gpurun_out/r2_final.ncu-rep "command line of the capture"
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
num = lambda r, k: float(r[hdr.index(k)].replace(",", ""))
tj = {"kernel_source_sha": bench.kernel_source_sha(), "source": f"{os.path.basename(rep)}: {how}"}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = "k_cn_hub_count" if "k_cn_hub_count" in name else ("k_cn_link" if "k_cn_link" in name else None)
    if short is None or short + "_dram_bytes_per_launch" in tj:
        continue
    def unit_bytes(k):
        v, u = num(r, k), rows[1][hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tj[short + "_dram_bytes_per_launch"] = int(unit_bytes("dram__bytes_read.sum") + unit_bytes("dram__bytes_write.sum"))
    tj[short + "_ncu"] = {
        "duration_us": num(r, "gpu__time_duration.sum"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": num(r, "smsp__inst_executed.sum"),
        "lanes_per_instruction": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
        "dram_pct_of_peak": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "long_scoreboard_stall_per_issue": num(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    }
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(tj, f, indent=1)
print(json.dumps(tj, indent=1))
