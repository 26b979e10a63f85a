#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page raw --csv`: per launch the duration, DRAM traffic, achieved bandwidth, occupancy,
issue utilisation, warp-stall reasons per issued instruction and shared-memory bank conflicts."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "smem_dyn"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%")]
STALLS = ["long_scoreboard", "barrier", "wait", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "lg_throttle", "no_instruction",
          "not_selected", "branch_resolving", "dispatch_stall", "membar", "tex_throttle", "drain", "imc_miss", "sleeping"]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    idx = {n: i for i, n in enumerate(hdr)}
    for r in data:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1]
        out = ["%-28s" % name[:28]]
        vals = {}
        for m, short in COLS:
            if m in idx:
                vals[short] = num(r[idx[m]])
                u = units[idx[m]]
                out.append("%s=%s%s" % (short, r[idx[m]], (" " + u) if u and u not in ("%",) else ""))
        t, rd, wr = vals.get("time"), vals.get("dram_rd"), vals.get("dram_wr")
        print(" ".join(out))
        st = []
        for s in STALLS:
            m = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            m2 = "smsp__average_warp_latency_issue_stalled_%s.ratio" % s
            for k in (m, m2):
                if k in idx and num(r[idx[k]]) is not None and num(r[idx[k]]) >= 0.05:
                    st.append("%s %.2f" % (s, num(r[idx[k]])))
                    break
        print("    stalls per issue: " + (", ".join(st) if st else "(metrics absent)"))


if __name__ == "__main__":
    main(sys.argv[1])
