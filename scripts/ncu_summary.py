"""Dev tool: condense one .ncu-rep (ncu --set full --import-source on) into the text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep "title / command line" > profiles/rNN_x.txt

Prints the launch metrics the roofline discussion in DESIGN.md uses and the share of executed warp-instructions
per CUDA source line range (function), from `--page source --print-source cuda,sass`.
"""
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_blocks",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    src_file = sys.argv[3] if len(sys.argv) > 3 else None
    print("# " + title)
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        print("## launch %d: %s" % (k, vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""))
        for name in KEEP:
            if name in hdr:
                i = hdr.index(name)
                print("%-72s %-16s %s" % (name, units[i], vals[i]))
    # stall reasons (warp state sampling)
    stall = [(h, rows[2][i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print("## warp stall reasons (avg warps stalled per issue-active cycle), launch 0, top 8")
    for h, v in sorted(stall, key=lambda t: -float(t[1] or 0))[:8]:
        print("%-90s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    out = ncu(rep, "--page", "source", "--print-source", "cuda,sass", "--csv")
    rows = list(csv.reader(io.StringIO(out)))
    hi = next((i for i, r in enumerate(rows) if r and r[0] == "Line No"), None)
    if hi is None:
        return
    hdr = rows[hi]
    ii = hdr.index("Instructions Executed")
    per = {}
    text = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0]:
            continue
        try:
            ln, inst = int(r[0]), int(r[ii])
        except ValueError:
            continue
        per[ln] = per.get(ln, 0) + inst
        text[ln] = r[1]
    tot = float(sum(per.values())) or 1.0
    print("## executed warp-instructions by CUDA source line (top 30 of %d lines; total %.4g)" % (len(per), tot))
    # group by enclosing function using the source text when available
    if src_file:
        src = open(src_file).read().split("\n")
        fn_of = {}
        cur = "?"
        pat = re.compile(r"__device__[^;(]*?\b(\w+)\s*\(|__global__[^;(]*?\b(\w+)\s*\(")
        for n, line in enumerate(src, 1):
            m = pat.search(line)
            if m and not line.strip().startswith("//"):
                cur = m.group(1) or m.group(2)
            fn_of[n] = cur
        agg = {}
        for ln, v in per.items():
            agg[fn_of.get(ln, "?")] = agg.get(fn_of.get(ln, "?"), 0) + v
        print("## share by function")
        for f, v in sorted(agg.items(), key=lambda t: -t[1]):
            if v / tot >= 0.001:
                print("%-28s %6.2f %%" % (f, 100 * v / tot))
        print("## top lines")
    for ln, v in sorted(per.items(), key=lambda t: -t[1])[:30]:
        print("%5d %6.2f %%  %s" % (ln, 100 * v / tot, text[ln].strip()[:120]))


if __name__ == "__main__":
    main()
