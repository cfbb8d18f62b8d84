#!/usr/bin/env python
"""Short summaries of an .ncu-rep (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py metrics <rep>            # a fixed short list of raw metrics per captured kernel
    python tools/ncu_summary.py hot <rep> [N]            # top-N SASS instructions by stall samples (needs -lineinfo/--import-source)
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def metrics(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name", "?")[:90])
        for k in KEYS:
            if k in d:
                print("  %-82s %s %s" % (k, d[k], units[hdr.index(k)]))


def hot(rep, n=30):
    """sequential listing: the top-n SASS instructions by stall samples, in program order, with running context"""
    rows = ncu_csv(rep, "source")
    # row 0: kernel name, row 1: header
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) >= len(hdr) - 1 and r[0].startswith("0x")]
    def val(r):
        try:
            return float(r[idx["# Samples"]])
        except (ValueError, KeyError):
            return 0.0
    tot = sum(val(r) for r in body) or 1.0
    order = sorted(range(len(body)), key=lambda i: val(body[i]), reverse=True)[:n]
    print("total samples %d over %d instructions" % (tot, len(body)))
    for i in sorted(order):
        r = body[i]
        print("%4d %5.1f%%  %s" % (i, 100 * val(r) / tot, r[idx["Source"]].strip()[:90]))
    # by opcode class
    agg = {}
    for r in body:
        op = r[idx["Source"]].strip().split()[0].split(".")[0]
        if op.startswith("@"):
            op = r[idx["Source"]].strip().split()[1].split(".")[0]
        agg[op] = agg.get(op, 0) + val(r)
    print("by opcode:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]))


if __name__ == "__main__":
    cmd, rep = sys.argv[1], sys.argv[2]
    if cmd == "metrics":
        metrics(rep)
    else:
        hot(rep, int(sys.argv[3]) if len(sys.argv) > 3 else 30)
