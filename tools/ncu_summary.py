"""Key counters of an `ncu --page raw --csv` export (one block per kernel).  python tools/ncu_summary.py raw.csv ..."""
import csv,sys
def summarize(path):
    rows=list(csv.reader(open(path)))
    h=rows[0]; u=rows[1]
    out={}
    for r in rows[2:]:
        d={h[i]:(r[i],u[i]) for i in range(len(h))}
        keys=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size',
         'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active',
         'smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
         'smsp__issue_active.avg.pct','sm__inst_executed.avg.per_cycle_elapsed','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum',
         'sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
         'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
         'sm__cycles_active.avg','smsp__cycles_active.avg','sm__cycles_elapsed.max']
        for k in keys:
            if k in d: print(f"  {k}: {d[k][0]} {d[k][1]}")
        st=[(k,float(v[0])) for k,v in d.items() if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and v[0]]
        print("  stalls per issue:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for k,v in sorted(st,key=lambda t:-t[1])[:7]))
for p in sys.argv[1:]:
    print(p); summarize(p)
