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
    # the page is a sequence of per-file tables: ["File Name", path], header row ["Line No", "Source", ...], lines
    per, samp, text, stalls = {}, {}, {}, {}
    # the combined cuda,sass view does not name the file of a line: resolve it by matching the line's text
    import os
    srcdir0 = os.path.dirname(src_file) if src_file else None
    file_lines = {}
    if srcdir0:
        for f in os.listdir(srcdir0):
            if f.endswith((".cuh", ".cu", ".hpp")):
                file_lines[f] = open(os.path.join(srcdir0, f)).read().split("\n")

    def resolve_file(ln, txt):
        t = txt.strip()
        main = os.path.basename(src_file) if src_file else None
        if main and ln <= len(file_lines.get(main, [])) and file_lines[main][ln - 1].strip() == t:
            return main
        for f, lines in file_lines.items():
            if ln <= len(lines) and lines[ln - 1].strip() == t:
                return f
        return "?"
    cur_file, hdr = "?", None
    STALL = ["stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_barrier", "stall_branch_resolving", "stall_mio", "stall_math", "stall_not_selected", "stall_selected", "stall_lg", "stall_dispatch"]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            ln = int(r[0]); inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        key = (resolve_file(ln, r[1]), ln)
        per[key] = per.get(key, 0) + inst
        try:
            samp[key] = samp.get(key, 0) + int(r[hdr.index("# Samples")])
        except ValueError:
            pass
        text[key] = r[1]
        st = stalls.setdefault(key, {})
        for name in STALL:
            if name in hdr:
                try:
                    st[name] = st.get(name, 0) + int(r[hdr.index(name)])
                except ValueError:
                    pass
    if not per:
        return
    tot = float(sum(per.values())) or 1.0
    stot = float(sum(samp.values())) or 1.0
    print("## executed warp-instructions and stall samples by CUDA source line (%d lines; total %.4g instructions, %.4g samples)" % (len(per), tot, stot))
    import os
    fn_of = {}
    srcdir = os.path.dirname(src_file) if src_file else None
    pat = re.compile(r"__device__[^;(]*?\b(\w+)\s*\(|__global__[^;(]*?\b(\w+)\s*\(")
    for f in set(k[0] for k in per):
        path = os.path.join(srcdir, f) if srcdir else None
        if path and os.path.exists(path):
            cur = "?"
            for n, line in enumerate(open(path).read().split("\n"), 1):
                m = pat.search(line)
                if m and not line.strip().startswith("//"):
                    cur = m.group(1) or m.group(2)
                fn_of[(f, n)] = cur
    agg, sagg, stagg = {}, {}, {}
    for key, v in per.items():
        fn = fn_of.get(key, key[0])
        agg[fn] = agg.get(fn, 0) + v
        sagg[fn] = sagg.get(fn, 0) + samp.get(key, 0)
        d = stagg.setdefault(fn, {})
        for name, c in stalls.get(key, {}).items():
            d[name] = d.get(name, 0) + c
    print("## share by function:  function, %% of executed instructions, %% of stall samples, top stall reasons (%% of the function's samples)")
    for f, v in sorted(sagg.items(), key=lambda t: -t[1]):
        if v / stot >= 0.002 or agg[f] / tot >= 0.002:
            d = stagg.get(f, {}); dt = float(sum(d.values())) or 1.0
            top = ", ".join("%s %.0f" % (n.replace("stall_", ""), 100 * c / dt) for n, c in sorted(d.items(), key=lambda t: -t[1])[:4])
            print("%-22s %6.2f %% inst %6.2f %% samples   %s" % (f, 100 * agg[f] / tot, 100 * v / stot, top))
    tots = {}
    for d in stalls.values():
        for n, c in d.items():
            tots[n] = tots.get(n, 0) + c
    tt = float(sum(tots.values())) or 1.0
    print("## stall samples by reason: " + ", ".join("%s %.1f%%" % (n.replace("stall_", ""), 100 * c / tt) for n, c in sorted(tots.items(), key=lambda t: -t[1])[:8]))
    print("## top lines by stall samples")
    for key, v in sorted(samp.items(), key=lambda t: -t[1])[:40]:
        d = stalls.get(key, {}); dt = float(sum(d.values())) or 1.0
        top = ", ".join("%s %.0f" % (n.replace("stall_", ""), 100 * c / dt) for n, c in sorted(d.items(), key=lambda t: -t[1])[:2])
        print("%-14s %5d %6.2f %% samples %6.2f %% inst  [%s]  %s" % (key[0][:14], key[1], 100 * v / stot, 100 * per[key] / tot, top, text[key].strip()[:90]))


if __name__ == "__main__":
    main()
