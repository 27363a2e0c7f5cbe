import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; data=rows[2:]
idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio']
sel2=[h for h in hdr if h.startswith('smsp__average_warp') and 'per_issue_active' in h]
for r in data:
    print('-----', r[idx['Kernel Name']][:50], r[idx['Grid Size']] if 'Grid Size' in idx else '', r[idx['Block Size']] if 'Block Size' in idx else '')
    for w in want:
        if w in idx: print(f"  {w:70s} {r[idx[w]][:20]} {units[idx[w]]}")
    vals=sorted(((float(r[idx[h]]),h) for h in sel2 if r[idx[h]] not in ('','n/a')),reverse=True)[:6]
    print('   stalls:', ', '.join(f"{h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}={v:.2f}" for v,h in vals))
