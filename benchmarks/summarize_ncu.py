"""Compact per-kernel summary of an `ncu --page raw --csv` export: the metrics DESIGN.md / bench.py quote.

    python benchmarks/summarize_ncu.py gpurun_out/r2_c4_kernels.raw.csv > profiles/r2_ncu_full_c4_kernels.csv
"""
import csv
import sys

KEYS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_static", "smem_static_KB"),
    ("launch__shared_mem_per_block_dynamic", "smem_dynamic_B"),
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor_hmma_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_cycles_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "stall_long_scoreboard_pct"),
    ("smsp__average_warp_latency_issue_stalled_barrier_per_warp_active.pct", "stall_barrier_pct"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard_per_warp_active.pct", "stall_short_scoreboard_pct"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    cols = []
    for key, name in KEYS:
        idx = [i for i, h in enumerate(hdr) if h == key]
        if idx:
            cols.append((idx[0], name + (f"[{units[idx[0]]}]" if units[idx[0]] and name not in ("kernel",) else "")))
    w = csv.writer(sys.stdout)
    w.writerow([n for _, n in cols])
    for r in rows[2:]:
        out = []
        for i, n in cols:
            v = r[i]
            if n == "kernel":
                v = v.split("(")[0].replace("void ", "")[:60]
            out.append(v)
        w.writerow(out)


if __name__ == "__main__":
    main()
