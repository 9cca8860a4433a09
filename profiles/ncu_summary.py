#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): headline metrics + per-source-line hot spots.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    out = []
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, unit = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append("## %s" % name)
        out.append("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append("| %s | %s | %s |" % (w, row[i], unit[i]))
    src = run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    rows = list(csv.reader(io.StringIO(src)))
    # sections: "Function Name",<kernel> ... header row ... (source line rows followed by their SASS rows)
    per = {}
    cur, H = None, None
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            cur = r[1].split("(")[0]
            H = None
            continue
        if "Instructions Executed" in r and "Source" in r:
            H = r
            ci, cs, csm = H.index("Instructions Executed"), H.index("Source"), H.index("# Samples")
            continue
        if cur is None or H is None or len(r) <= max(ci, cs, csm) or not r[0].isdigit():
            continue
        try:
            n, smp = int(r[ci]), int(r[csm])
        except ValueError:
            continue
        per.setdefault(cur, []).append((r[0], r[cs], n, smp))
    for fn, use in per.items():
        tot_i = sum(x[2] for x in use) or 1
        tot_s = sum(x[3] for x in use) or 1
        if tot_i < 1000:
            continue
        out.append("\n### %s: hottest source lines by executed warp instructions" % fn)
        out.append("| line | % inst | % stall samples | source |\n|---|---|---|---|")
        for a_, b_, n, smp in sorted(use, key=lambda x: -x[2])[:16]:
            out.append("| %s | %.1f | %.1f | `%s` |" % (a_, 100.0 * n / tot_i, 100.0 * smp / tot_s, b_.strip()[:110]))
        out.append("\n### %s: hottest source lines by stall samples" % fn)
        out.append("| line | % inst | % stall samples | source |\n|---|---|---|---|")
        for a_, b_, n, smp in sorted(use, key=lambda x: -x[3])[:8]:
            out.append("| %s | %.1f | %.1f | `%s` |" % (a_, 100.0 * n / tot_i, 100.0 * smp / tot_s, b_.strip()[:110]))
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
