import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
keys=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__cycles_elapsed.max','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__waves_per_multiprocessor','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_bytes.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','sm__sass_thread_inst_executed_op_ffma_pred_on.sum','sm__sass_thread_inst_executed_op_fadd_pred_on.sum','sm__sass_thread_inst_executed_op_fmul_pred_on.sum','smsp__cycles_active.avg','sm__cycles_active.avg']
for r in rows[2:]:
    for k in keys:
        for i,h in enumerate(hdr):
            if h==k: print(h,'=',r[i],units[i])
    print('-- stalls (per issue active)')
    for i,h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and 'not_issued' not in h:
            try: v=float(r[i])
            except: continue
            if v>0.08: print('  ',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), '%.2f'%v)
    print('---')
