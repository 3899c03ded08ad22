import csv, sys, collections, subprocess
def launches(path):
    rows=[r for r in csv.reader(open(path)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    agg=collections.OrderedDict()
    for r in rows[1:]:
        try: v=float(r[vi].replace(',',''))
        except: continue
        k=r[ki].replace('void ','').replace('bm::<unnamed>::','')[:60]
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
    tot=sum(a[1] for a in agg.values())
    for k,a in sorted(agg.items(), key=lambda x:-x[1][1]):
        print(f"{a[1]/1e3:10.1f} us {a[0]:4d}x {a[1]/a[0]/1e3:9.1f} us avg {100*a[1]/tot:5.1f}%  {k}")
def raw(rep, pat=None):
    out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines())); hdr=rows[0]; units=rows[1]
    want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','smsp__issue_active.avg.pct','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','sm__cycles_active.avg','sm__cycles_elapsed.max',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio','smsp__average_warps_issue_stalled_selected_per_issue_active.ratio']
    for r in rows[2:]:
        name=r[hdr.index('Kernel Name')]
        if pat and pat not in name: continue
        print('----', name[:110])
        for w in want:
            if w in hdr:
                i=hdr.index(w); print(f"  {w:86s} {r[i]:>16s} {units[i]}")
if sys.argv[1]=='launches': launches(sys.argv[2])
else: raw(sys.argv[2], sys.argv[3] if len(sys.argv)>3 else None)
