"""Summarise an .ncu-rep (ncu --set full) into the few metrics quoted in DESIGN.md: python tools/ncu_summary.py file.ncu-rep"""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','smsp__warps_eligible.avg.per_cycle_active','derived__memory_l2_theoretical_sectors_global_excessive','lts__t_sectors.sum']
for r in rows[2:]:
    for w in want:
        if w in hdr: print(f'{w:75s} {r[hdr.index(w)]}')
    # stall reasons
    for i,h in enumerate(hdr):
        if h.startswith('smsp__average_warp') and h.endswith('per_issue_active.ratio') or ('warps_issue_stalled' in h and h.endswith('_per_warp_active.pct')):
            pass
    st=[(h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),float(r[i])) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and r[i]]
    for k,v in sorted(st,key=lambda x:-x[1])[:10]: print(f'   stall {k:30s} {v:.3f}')
