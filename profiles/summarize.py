"""Summarise an Nsight Compute report into a small JSON (run here, no GPU):
    python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/<name>.json
Keeps the metrics the roofline argument needs (B200_PROFILING.md)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    d[k] = {"value": r[i], "unit": units[i]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
        d["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rb = d["dram__bytes_read.sum"]["value"] * scale[d["dram__bytes_read.sum"]["unit"]]
            wb = d["dram__bytes_write.sum"]["value"] * scale[d["dram__bytes_write.sum"]["unit"]]
            d["traffic_bytes"] = rb + wb
        res.append(d)
    json.dump({"report": rep, "launches": res}, open(out, "w"), indent=1)
    print("wrote", out, "with", len(res), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
