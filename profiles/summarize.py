#!/usr/bin/env python
"""Summarise ncu output brought back from the GPU box into the tracked profiles/ directory.

    python profiles/summarize.py TAG          # reads gpurun_out/TAG_launches.csv and gpurun_out/TAG_prof.ncu-rep

Writes profiles/TAG_launches.csv (per-kernel launch list, our kernels + the torch ones of one step),
profiles/TAG_launch_shares.txt (share of the step per kernel) and profiles/TAG_ncu_full.txt (the
roofline-relevant metrics of every captured launch of the `--set full` pass).  Numbers printed under
ncu are never bench values; only shares and per-launch DRAM traffic are taken from here.
"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
]


def short(name):
    name = name.replace("void ", "")
    return name if len(name) < 110 else name[:107] + "..."


def launches(tag):
    src = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    out = os.path.join(ROOT, "profiles", tag + "_launches.csv")
    with open(out, "w") as f:
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in rows:
            f.write('%s,"%s","%s","%s",%s\n' % (r["ID"], short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], r["Metric Value"]))
    # share of one step: take the LAST occurrence of each of our kernels plus everything between the last
    # forward warp launch and the last backward warp launch
    ours = [i for i, r in enumerate(rows) if "pd::" in r["Kernel Name"]]
    with open(os.path.join(ROOT, "profiles", tag + "_launch_shares.txt"), "w") as f:
        if not ours:
            f.write("no library kernels in the launch list\n")
            return
        fwd = [i for i in ours if "warp_composite_fwd" in rows[i]["Kernel Name"] or "rows_fwd" in rows[i]["Kernel Name"]]
        bwd = [i for i in ours if "warp_composite_bwd" in rows[i]["Kernel Name"] or "rows_bwd" in rows[i]["Kernel Name"]]
        a, b = fwd[-1], bwd[-1]
        step = rows[a:b + 1]
        tot = sum(float(r["Metric Value"]) for r in step)
        f.write("one step of the launch list (ids %s..%s), ncu-serialised cold-cache durations: compare SHARES only\n" % (rows[a]["ID"], rows[b]["ID"]))
        f.write("%-9s %8s  %s\n" % ("share", "ns", "kernel"))
        for r in step:
            f.write("%8.1f%% %8.0f  %s\n" % (100 * float(r["Metric Value"]) / tot, float(r["Metric Value"]), short(r["Kernel Name"])))
        f.write("total %.0f ns; library kernels %.1f%% of the step\n" % (
            tot, 100 * sum(float(r["Metric Value"]) for r in step if "pd::" in r["Kernel Name"]) / tot))


def full(tag, suffix=""):
    """suffix "" = the main capture TAG_prof.ncu-rep; "_cfg3" etc. = secondary captures exported on the box as
    TAG_prof_cfg3_raw.csv (their .ncu-rep files are too large to bring back)."""
    rep = os.path.join(ROOT, "gpurun_out", tag + "_prof" + suffix + ".ncu-rep")
    pre = os.path.join(ROOT, "gpurun_out", tag + "_prof" + suffix + "_raw.csv")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    elif os.path.exists(pre):
        raw = open(pre).read()
    else:
        return
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(ROOT, "profiles", tag + "_ncu_full" + suffix + ".txt"), "w") as f:
        f.write("ncu --set full --clock-control none (one launch per block below); from gpurun_out/%s_prof%s.ncu-rep\n" % (tag, suffix))
        for r in rows[2:]:
            f.write("\n== %s\n" % short(r[hdr.index("Kernel Name")]))
            vals = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    vals[k] = r[i]
                    f.write("  %-75s %s %s\n" % (k, r[i], units[i]))
            try:
                sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(vals["dram__bytes_read.sum"]) * sc.get(units[hdr.index("dram__bytes_read.sum")], 1)
                wr = float(vals["dram__bytes_write.sum"]) * sc.get(units[hdr.index("dram__bytes_write.sum")], 1)
                f.write("  %-75s %.1f MB\n" % ("traffic = dram read + write per launch", (rd + wr) / 1e6))
            except Exception:
                pass


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    full(tag)
    for sfx in ("_cfg3", "_cfg4", "_resize"):
        full(tag, sfx)
    print("wrote profiles/%s_*" % tag)
