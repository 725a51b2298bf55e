"""Key metrics + top stall-sample SASS lines of ncu reports:  python scripts/ncu_summary.py gpurun_out/ncu_*.ncu-rep"""
import csv, subprocess, sys
def summarize(rep, ntop=14):
    raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(raw.splitlines())); d=dict(zip(rows[0],rows[-1])); u=dict(zip(rows[0],rows[1]))
    out=[f"## {d['Kernel Name'][:90]}  grid {d.get('launch__grid_size')} block {d.get('launch__block_size')}"]
    for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg.per_second','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sectors.avg.pct_of_peak_sustained_elapsed','l1tex__t_sectors.avg.pct_of_peak_sustained_elapsed','smsp__pcsamp_sample_count','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_not_selected']:
        if k in d: out.append(f"  {k:72s} {d[k]:>14s} {u.get(k,'')}")
    src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
    rows=list(csv.reader(src.splitlines())); hdr=rows[1]
    isrc,isamp,iex=hdr.index("Source"),hdr.index("# Samples"),hdr.index("Instructions Executed")
    data=[(int(r[isamp] or 0), r[isrc].strip(), int(r[iex] or 0), i) for i,r in enumerate(rows[2:]) if len(r)>isamp]
    tot=sum(x[0] for x in data)
    out.append(f"  top stall-sample SASS instructions ({tot} samples, {len(data)} instructions):")
    for s_,srcl,ex,i in sorted(data,reverse=True)[:ntop]:
        out.append(f"    {s_:5d} {100*s_/max(tot,1):5.1f} %  #{i:4d} executed {ex:9d}  {srcl[:80]}")
    return "\n".join(out)
if __name__=="__main__":
    for r in sys.argv[1:]: print(summarize(r)); print()
