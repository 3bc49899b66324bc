"""Summaries of an .ncu-rep (read with the ncu CLI): key raw metrics, stall shares, hottest SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
       "launch__shared_mem_per_block_dynamic", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = ["long_sb", "short_sb", "wait", "branch_resolving", "not_selected", "selected", "mio", "math", "barrier",
          "no_inst", "lg", "dispatch", "sleep", "misc", "drain", "tex", "membar"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for i, h in enumerate(hdr):
        if h in RAW or h == "Kernel Name":
            print("%-70s %-12s %s" % (h, units[i], vals[i]))
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx]))))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("samples", tot, "warp instructions", sum(int(r[ix["Instructions Executed"]]) for r in data))
    agg = {k: sum(int(r[ix["stall_" + k]]) for r in data) for k in STALLS}
    print("stall shares:", {k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.005 * tot})
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top_n]
    for i in sorted(top):
        r = data[i]
        n = int(r[ix["# Samples"]])
        st = {k: int(r[ix["stall_" + k]]) for k in STALLS}
        st = {k: v for k, v in st.items() if v > 0.15 * n}
        print("%5d %-58s %6d %10s %s" % (i, r[ix["Source"]].strip()[:58], n, r[ix["Instructions Executed"]], st))


if __name__ == "__main__":
    main()
