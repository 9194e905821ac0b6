"""Summarise an .ncu-rep: python scripts/ncu_summary.py file.ncu-rep [pattern ...]  (reads `ncu --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
                        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "issue_stalled", "smsp__inst_executed.sum",
                        "smsp__inst_executed_pipe_fp64", "sm__inst_executed_pipe_", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum ",
                        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
print("kernels:", [r[ki][:40] for r in data])
for i, h in enumerate(hdr):
    if any(p in h for p in pats) and "Triage" not in h:
        vals = [r[i] for r in data]
        if all(v in ("0", "0.000000", "") for v in vals):
            continue
        print(f"{h} [{units[i]}]: {vals}")
