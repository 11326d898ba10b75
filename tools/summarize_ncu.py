"""Turn gpurun_out/*.ncu-rep / launch CSVs into the text summaries committed under profiles/ (run here, no GPU needed)."""
import collections
import csv
import re
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__block_size', 'smsp__inst_executed.sum']


def full(rep, out, header):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [header]
    for r in rows[2:]:
        lines.append("----")
        for w in WANT:
            if w in idx:
                lines.append(f"{w:78s} {r[idx[w]][:90]} {units[idx[w]]}")
    open(out, "w").write("\n".join(lines) + "\n")


def launches(csv_path, out, header):
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        name = re.sub(r"<.*", "", r[i_name]).replace("void ", "")
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("b2pc::"))
    lines = [header, f"# total {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches; b2pc kernels {100 * ours / tot:.1f} %"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append(f"{v[1] / 1e6:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  {k[:100]}")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    import os
    if len(sys.argv) >= 4 and sys.argv[1] == "full":          # python tools/summarize_ncu.py full <rep> <out> "<header>"
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "# ncu --set full")
    elif len(sys.argv) >= 4 and sys.argv[1] == "launches":    # python tools/summarize_ncu.py launches <csv> <out> "<header>"
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "# ncu launch list")
    else:
        print(__doc__)
