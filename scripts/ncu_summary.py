import csv,sys,subprocess,io
out=subprocess.run(["ncu","-i",sys.argv[1],"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread ','launch__occupancy_limit_shared','launch__occupancy_limit_reg','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum ','smsp__issue_active.avg.pct','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ','launch__grid_size','smsp__thread_inst_executed_per_inst_executed','launch__shared_mem_per_block_dynamic','dram__bytes_read.sum ','dram__bytes_write.sum ','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ']
for h,u,v in zip(hdr,units,vals):
    if any(w in h+' ' for w in want): print(h,u,v)
st={h.replace('smsp__pcsamp_warps_issue_stalled_',''):int(v) for h,v in zip(hdr,vals) if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('not_issued')}
tot=sum(st.values())
print('stalls:', ', '.join(f"{k} {100*v/tot:.0f}%" for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:9]))
