#!/bin/bash
# Round-2 evidence set: run on the GPU box through gpurun (`gpurun -- bash profiles/collect_evidence_r2.sh`).
# .ncu-rep files are exported to tables on the box and deleted (gpurun_out/ is capped at 64 MiB).
set -x
O=gpurun_out/ev_r2; mkdir -p $O
B="python bench.py --no-sampler --no-cpu-baseline --no-vae --no-graph"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file $O/launches_r2_warm.csv $B --steps 2 --warmup 2 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r2_cold.csv $B --steps 2 --warmup 2 > /dev/null 2>&1
# one whole B=200, L=16 step (23 launches), full counter set
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 69 -c 23 -o $O/prof_step -f $B --steps 2 --warmup 2 > $O/ncu_step.log 2>&1
python profiles/rawsum.py $O/prof_step.ncu-rep > $O/ncu_full_step_kernels_r2.md
for k in lstm_pair_fwd_kernel lstm_bwd_kernel; do ncu -i $O/prof_step.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > $O/src_$k.csv 2>/dev/null; python profiles/stalls.py $O/src_$k.csv 24 > $O/ncu_stalls_${k}_B200_r2.txt; rm -f $O/src_$k.csv; done
rm -f $O/prof_step.ncu-rep
# CL-VAE fused step
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vae_fused -s 3 -c 1 -o $O/prof_vae -f python bench.py --no-sampler --no-cpu-baseline --no-graph --steps 2 --warmup 2 > $O/ncu_vae.log 2>&1
python profiles/rawsum.py $O/prof_vae.ncu-rep > $O/ncu_full_vae_fused_r2.md
ncu -i $O/prof_vae.ncu-rep --page source --csv > $O/src_vae.csv 2>/dev/null; python profiles/stalls.py $O/src_vae.csv 24 > $O/ncu_stalls_vae_fused_kernel_B100_r2.txt; rm -f $O/src_vae.csv $O/prof_vae.ncu-rep
timeout 120 python profiles/pbench.py 200 16 > $O/pbench_200_16.txt 2>&1
timeout 120 ./profiles/matvec_bench > $O/matvec_bench.txt 2>&1
for bl in "64 32" "1024 32" "4096 32" "16384 32" "4096 128" "1024 512"; do set -- $bl; timeout 200 python bench.py --batch $1 --seq-len $2 --steps 5 --warmup 3 --no-sampler --no-cpu-baseline --no-vae 2>/dev/null | grep "^{" > $O/sweep_$1_$2.json; done
timeout 400 python bench.py > $O/bench_default_r2.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_r2.json 2> $O/bench_reference.err
du -sh $O
