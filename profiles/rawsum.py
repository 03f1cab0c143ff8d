import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_static','launch__shared_mem_per_block_dynamic','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
idx=[(w,hdr.index(w)) for w in want if w in hdr]
ki=hdr.index('Kernel Name')
print("| kernel | "+" | ".join(w.split('.')[0].replace('__',':') for w,_ in idx)+" |")
print("|---|"+"---|"*len(idx))
for r in rows[2:]:
    name=r[ki].split('(')[0].replace('void ','').replace('<unnamed>::','')
    print("| %s | "%name+" | ".join((r[i]+(' '+units[i] if units[i] not in ('','%') else '')) for _,i in idx)+" |")
