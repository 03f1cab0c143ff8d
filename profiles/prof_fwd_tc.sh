O=gpurun_out/ev_tc; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_tc_kernel -s 4 -c 1 -o $O/prof_tc -f python bench.py --batch 16384 --seq-len 32 --no-sampler --no-cpu-baseline --no-vae --no-graph --steps 2 --warmup 2 > $O/ncu_tc.log 2>&1
python profiles/rawsum.py $O/prof_tc.ncu-rep > $O/ncu_full_lstm_fwd_tc_r2.md
ncu -i $O/prof_tc.ncu-rep --page source --csv > $O/src_tc.csv 2>/dev/null; python profiles/stalls.py $O/src_tc.csv 40 > $O/ncu_stalls_lstm_fwd_tc_kernel_B16384_r2.txt
ncu -i $O/prof_tc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[-1]
for k in ('smsp__cycles_active.avg','sm__inst_executed.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','lts__t_sectors_srcunit_tex_op_read.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active'):
    for i,n in enumerate(h):
        if n==k: print(k, v[i])
" > $O/raw_sel.txt
rm -f $O/src_tc.csv $O/prof_tc.ncu-rep
