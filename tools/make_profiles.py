#!/usr/bin/env python3
"""Turn the scratch ncu outputs of a gpurun call (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profiles.py --round r01 --launches gpurun_out/launches.csv --rep gpurun_out/prof_r01_chain.ncu-rep \
        --bench gpurun_out/bench_default.json [--tag chain]

Writes profiles/<round>_launches.csv (copy), profiles/<round>_launches_summary.txt (per-kernel share of the step),
profiles/<round>_<tag>_ncu.txt (the raw metrics DESIGN.md quotes, per profiled launch), profiles/<round>_<tag>_sass_mix.txt
(executed warp-instructions per opcode) and profiles/traffic.json (DRAM bytes per launch, read by bench.py).
"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg.per_second",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches_summary(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    d = collections.OrderedDict()
    for r in rows[1:]:
        try:
            d.setdefault((r[ki], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare shares)\n")
        f.write("# %-74s %-12s %-10s %6s %12s %12s %7s\n" % ("kernel", "grid", "block", "n", "total_us", "mean_us", "share"))
        for (k, g, b), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-76s %-12s %-10s %6d %12.1f %12.2f %7.4f\n" % (k[:76], g, b, len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, sum(v) / tot))


def ncu_raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    traffic = []
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; raw page, selected metrics per profiled launch\n")
        for r in rows[2:]:
            f.write("== launch %s  %s  block %s grid %s\n" % (r[0], r[hdr.index("Kernel Name")][:90], r[hdr.index("Block Size")], r[hdr.index("Grid Size")]))
            vals = {}
            for i, h in enumerate(hdr):
                if h in KEEP:
                    f.write("  %-90s %18s %s\n" % (h, r[i], units[i]))
                    vals[h] = (r[i], units[i])
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            try:
                rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * scale[vals["dram__bytes_read.sum"][1]]
                wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * scale[vals["dram__bytes_write.sum"][1]]
                traffic.append(rd + wr)
            except Exception:
                pass
    return traffic


def sass_mix(rep, out):
    txt = subprocess.run(["python3", os.path.join(ROOT, "tools", "ncu_source.py"), rep, "--top", "45"], capture_output=True, text=True).stdout
    with open(out, "w") as f:
        f.write("# executed warp-instructions per SASS opcode of the first profiled launch (ncu source page)\n" + txt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--launches", default="")
    ap.add_argument("--rep", default="")
    ap.add_argument("--tag", default="chain")
    ap.add_argument("--bench", default="")
    ap.add_argument("--traffic-json", action="store_true", help="update profiles/traffic.json from --rep")
    a = ap.parse_args()
    prof = os.path.join(ROOT, "profiles")
    os.makedirs(prof, exist_ok=True)
    if a.launches:
        shutil.copy(a.launches, os.path.join(prof, "%s_launches.csv" % a.round))
        launches_summary(a.launches, os.path.join(prof, "%s_launches_summary.txt" % a.round))
    if a.rep:
        traffic = ncu_raw(a.rep, os.path.join(prof, "%s_%s_ncu.txt" % (a.round, a.tag)))
        sass_mix(a.rep, os.path.join(prof, "%s_%s_sass_mix.txt" % (a.round, a.tag)))
        if a.traffic_json and traffic:
            with open(os.path.join(prof, "traffic.json"), "w") as f:
                json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic), "source": os.path.basename(a.rep),
                           "how": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, mean over the profiled launches"}, f, indent=1)
    if a.bench:
        shutil.copy(a.bench, os.path.join(prof, "%s_bench_n1.json" % a.round))


if __name__ == "__main__":
    main()
