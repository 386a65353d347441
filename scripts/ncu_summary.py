"""Print the headline metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_lim_regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__cycles_active.avg", "cycles_active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"),
        ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall_short_sb"),
        ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall_lg_throttle"),
        ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall_mio"),
        ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall_barrier"),
        ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall_math"),
        ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall_wait"),
        ]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
ki = h.index("Kernel Name")
for r in data:
    print("==", r[ki][:90])
    parts = []
    for m, label in WANT:
        if m in h:
            i = h.index(m)
            parts.append(f"{label}={r[i]}{units[i] if units[i] not in ('', '%') else ''}")
    for i in range(0, len(parts), 6):
        print("   ", "  ".join(parts[i:i + 6]))
