import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','lts__t_bytes.sum ','smsp__inst_executed.sum ','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct','launch__grid_size','sm__pipe_tensor','smsp__average_warps_issue_stalled','lts__throughput.avg.pct','l1tex__throughput.avg.pct','Kernel Name','launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_tensor']
for i,h in enumerate(hdr):
    if any(w.strip() in h for w in want):
        print(f"{h[:88]:88s} {units[i][:10]:10s}", [r[i][:22] for r in rows[2:]])
