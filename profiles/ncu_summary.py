import csv,collections,sys
raw,src=sys.argv[1],sys.argv[2]
rows=list(csv.reader(open(raw)))
hdr=rows[0]; units=rows[1]; idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','launch__block_size','launch__grid_size','launch__shared_mem_per_block_dynamic','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','sm__inst_executed_pipe_lsu.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('==',r[idx['Kernel Name']][:60])
    for w in want:
        if w in idx: print(' ',w.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''),'=',r[idx[w]],units[idx[w]])
rows=list(csv.reader(open(src)))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
tot=0; byop=collections.Counter(); recs=[]
for r in rows[2:]:
    try: n=int(r[idx['Instructions Executed']])
    except: continue
    s=r[idx['Source']]; toks=s.split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    byop[op.split('.')[0]]+=n; tot+=n; recs.append((n,s,int(r[idx['# Samples']] or 0)))
print('total',tot)
print(' '.join(f"{k}:{v/tot*100:.1f}%" for k,v in byop.most_common(18)))
cnts=collections.Counter(n for n,_,_ in recs)
for c,k in sorted(cnts.items(), key=lambda x:-x[0]*x[1])[:8]: print(" count",c,"x",k,"=>",round(c*k/tot*100,1),"%")
ts=sum(s for _,_,s in recs)
print("top stall-sample instrs:")
for n,s,sm in sorted(recs,key=lambda x:-x[2])[:14]: print(f"  {sm/ts*100:5.1f}%  n={n}  {s[:90]}")
