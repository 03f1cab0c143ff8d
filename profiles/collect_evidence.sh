#!/bin/bash
# Round-1 evidence set: run on the GPU box through gpurun (`gpurun -- bash profiles/collect_evidence.sh`).
# .ncu-rep files are exported to tables on the box and deleted (gpurun_out/ is capped at 64 MiB).
set -x
O=gpurun_out/ev; mkdir -p $O
B="python bench.py --no-sampler --no-cpu-baseline --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 130 --csv --log-file $O/launches_warm.csv $B --steps 2 --warmup 2 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file $O/launches_cold.csv $B --steps 2 --warmup 2 > /dev/null 2>&1
# one whole B=200, L=16 step (25 launches), full counter set
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 75 -c 25 -o $O/prof_step -f $B --steps 2 --warmup 2 > $O/ncu_step.log 2>&1
python profiles/rawsum.py $O/prof_step.ncu-rep > $O/ncu_full_step_kernels.md
for k in lstm_fwd_kernel lstm_bwd_kernel; do ncu -i $O/prof_step.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > $O/src_$k.csv 2>/dev/null; python profiles/stalls.py $O/src_$k.csv 24 > $O/ncu_stalls_${k}_B200.txt; rm -f $O/src_$k.csv; done
rm -f $O/prof_step.ncu-rep
# large batch: the kernels that carry the step at B=16384, L=32
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_wgrad_tc|lstm_fwd_tc|inproj_tc|xhead|lstm_bwd|adamwn|keyenc" -c 14 -o $O/prof_big -f $B --batch 16384 --seq-len 32 --steps 1 --warmup 1 > $O/ncu_big.log 2>&1
python profiles/rawsum.py $O/prof_big.ncu-rep > $O/ncu_full_large_batch.md
for k in lstm_bwd_kernel lstm_fwd_tc_kernel lstm_wgrad_tc_kernel; do ncu -i $O/prof_big.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > $O/src_$k.csv 2>/dev/null; python profiles/stalls.py $O/src_$k.csv 24 > $O/ncu_stalls_${k}_B16384.txt; rm -f $O/src_$k.csv; done
rm -f $O/prof_big.ncu-rep
for bl in "200 16" "16384 32" "65536 32"; do set -- $bl; timeout 300 python profiles/kbench.py $1 $2 > $O/kbench_$1_$2.txt 2>&1; done
for bl in "64 32" "1024 32" "4096 32" "16384 32" "65536 32" "4096 128" "1024 512"; do set -- $bl; timeout 300 python bench.py --batch $1 --seq-len $2 --steps 5 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | grep "^{" > $O/sweep_$1_$2.json; done
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
du -sh $O
