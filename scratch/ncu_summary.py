import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.per_cycle_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr: print('  %-70s %s'%(w,r[hdr.index(w)]))
    for name in hdr:
        if 'average_warps_issue_stalled' in name and name.endswith('per_issue_active.ratio'):
            v=r[hdr.index(name)]
            try:
                if float(v)>0.15: print('  stall %-60s %s'%(name.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v))
            except: pass
