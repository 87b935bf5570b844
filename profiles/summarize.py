"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/launches_r01.txt
    python profiles/summarize.py kernel   gpurun_out/prof_mask_r01.ncu-rep > profiles/mask_r01.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    acc = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        e = acc.setdefault(r[ki], [0, 0.0])
        e[0] += 1
        e[1] += v
    tot = sum(v[1] for v in acc.values())
    print(f"# {path}: per-kernel device time under ncu (cold cache, serialised: compare SHARES)")
    for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} launches={c:4d} total_us={t:11.1f} share={t / tot:.3f}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    print(f"# {path}")
    for D in rows[2:]:
        print(f"## {D[H.index('Kernel Name')]}  grid={D[H.index('Grid Size')]} block={D[H.index('Block Size')]}")
        for m in METRICS:
            if m in H:
                print(f"{m:90s} {D[H.index(m)]:>16s} {U[H.index(m)]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
