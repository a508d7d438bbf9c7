"""Condense an ncu --set full report (raw page CSV on stdin) to the metrics the roofline discussion uses."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        # the tensor pipe per sub-partition: one elected thread of ONE warp issues every tcgen05.mma, so only one of the four
        # SMSPs of an SM ever shows tensor-pipe instructions; the *.avg over SMSPs is therefore a quarter of the busy SMSP's
        "smsp__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_tensor.max", "smsp__inst_executed_pipe_tensor.avg",
        "sm__inst_executed_pipe_tensor.sum", "smsp__pipe_tensor_cycles_active.max", "smsp__pipe_tensor_cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_active.avg"]
print("metric,unit," + ",".join(f"launch{i}" for i in range(len(rows) - 2)))
for w in want:
    for i, h in enumerate(hdr):
        if h == w or h.endswith("." + w):
            print(",".join([w, units[i]] + [r[i].replace(",", ";")[:60] for r in rows[2:]]))
            break
