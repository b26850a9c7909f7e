"""Key metrics of one kernel from an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
        "sm__pipe_tensor_cycles_active.avg", "sm__throughput.avg", "smsp__issue_active.avg", "smsp__inst_executed.sum",
        "smsp__pcsamp_warps_issue_stalled", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "sm__warps_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "sm__cycles_active.avg", "sm__inst_executed_pipe_", "l1tex__t_requests_pipe_lsu_mem_global_op_st",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
        "l1tex__lsu_writeback", "smsp__average_warp", "sm__pipe_alu_cycles_active", "sm__pipe_fma_cycles_active",
        "sm__pipe_fmaheavy", "sm__inst_executed_pipe_xu", "l1tex__throughput", "lts__throughput")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print("#", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k) for k in KEYS) and not h.endswith("_not_issued"):
            print(f"{h:92s} {u:16s} {v}")
