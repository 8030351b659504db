#!/usr/bin/env python
"""Turn gpurun_out/ captures (profiles/collect.sh) into the tracked summaries under profiles/.

usage: python profiles/summarize.py TAG        (run in the build container; needs `ncu` for --page raw)
writes profiles/TAG_bench_c*.json (the bench lines as measured), profiles/TAG_launches_c2.csv,
profiles/TAG_ncu_<kernel>.json (selected metrics of the --set full capture) and
profiles/dominant_kernel_traffic.json (dram bytes per launch of the bench's dominant kernel; bench.py
reports it as roofline.traffic)."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_sample_buffer_full"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def main(tag):
    for f in sorted(os.listdir(OUT)):
        if f.startswith(f"bench_{tag}_") and f.endswith(".json") and os.path.getsize(os.path.join(OUT, f)) > 0:
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f"{tag}_" + f[len(f"bench_{tag}_"):].replace(".json", "_bench.json")))
    lc = os.path.join(OUT, f"launches_{tag}_c2.csv")
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(PROF, f"{tag}_launches_c2.csv"))
    dom = {}
    for f in sorted(os.listdir(OUT)):
        # raw-metrics CSV written on the GPU box by collect.sh (export_rep), or a report brought back whole
        if f.startswith(f"raw_{tag}_") and f.endswith(".csv"):
            kern = f[len(f"raw_{tag}_"):-len(".csv")]
            raw = open(os.path.join(OUT, f)).read()
        elif f.startswith(f"prof_{tag}_") and f.endswith(".ncu-rep"):
            kern = f[len(f"prof_{tag}_"):-len(".ncu-rep")]
            raw = subprocess.run(["ncu", "-i", os.path.join(OUT, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        else:
            continue
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {"Kernel Name": {"value": vals[hdr.index("Kernel Name")], "unit": ""}}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": vals[i], "unit": units[i]}
        json.dump(d, open(os.path.join(PROF, f"{tag}_ncu_{kern}.json"), "w"), indent=1)
        try:
            dom[kern] = to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"]) + \
                to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"])
        except KeyError:
            pass
    sor = [k for k in ("k_sor_ring", "k_sor_pair", "k_sor_reg", "k_sor") if k in dom]   # whichever sweep variant the capture ran
    if sor:
        json.dump({"kernel": sor[0], "dram_bytes_per_launch": dom[sor[0]], "source": f"profiles/{tag}_ncu_{sor[0]}.json",
                   "workload": "stack32 x 4096 worlds, step 306 (settled), one launch, ncu --set full --clock-control none"},
                  open(os.path.join(PROF, "dominant_kernel_traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
