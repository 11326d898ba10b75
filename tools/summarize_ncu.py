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
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


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
    launches("gpurun_out/r01_launches.csv", "profiles/r01_ncu_launch_summary.txt",
             "# ncu launch list of one PT-v3m1-base training step (2 x 120k-voxel scenes): bench.py --steps 2 --warmup 3 under\n"
             "# ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 2700 (cold-cache, serialised: compare SHARES)")
    full("gpurun_out/r01_ncu_attn.ncu-rep", "profiles/r01_ncu_attention_full_metrics.txt",
         "# ncu --set full --clock-control none, tools/probe_attn.py time (H=2, T=241664, K=1024, bf16): attn_fwd_umma_kernel (TMA path)")
    full("gpurun_out/r01_ncu_conv.ncu-rep", "profiles/r01_ncu_conv_full_metrics.txt",
         "# ncu --set full --clock-control none, tools/probe_conv.py ONLY=0 (N=240000, C=32->32, 3^3, rows in shuffled order): "
         "gather_gemm_umma_kernel fwd, bwd-data, then bwd_weight_umma_kernel")
