import csv, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
d=dict(zip(hdr,vals))
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__warps_eligible.avg.per_cycle_active','lts__t_bytes.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for k in keys:
    print(f"{k:70s} {d.get(k)} {units[hdr.index(k)] if k in hdr else ''}")
for h,v in zip(hdr,vals):
    if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h:
        try:
            if float(v)>0.04: print(f"  {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {float(v):.3f}")
        except: pass
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}; data=rows[2:]
def f(r,k):
    try: return float(r[idx[k]])
    except: return 0.0
tot=sum(f(r,'# Samples') for r in data)
print('samples',tot,'sass instrs',len(data))
top=sorted(data,key=lambda r:-f(r,'# Samples'))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in top:
    best=max(stalls,key=lambda k:f(r,k))
    print(f"  {f(r,'# Samples'):8.0f} {100*f(r,'# Samples')/tot:5.2f}%  {best:22s} {r[idx['Address']][-5:]} {r[idx['Source']][:80]}")
