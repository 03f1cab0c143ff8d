#!/bin/bash
# Round-2 evidence, second set (tensor-core X head, refreshed sweep and default bench): gpurun -- bash profiles/collect_evidence_r2b.sh
set -x
O=gpurun_out/ev_r2b; mkdir -p $O
# tensor-core X head, 524 288 rows: full counter set + warp-stall sampling of one launch
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xhead_tc_kernel -s 27 -c 1 -o $O/prof_x -f python profiles/xbench.py > $O/ncu_x.log 2>&1
python profiles/rawsum.py $O/prof_x.ncu-rep > $O/ncu_full_xhead_tc_r2.md
ncu -i $O/prof_x.ncu-rep --page source --csv > $O/src_x.csv 2>/dev/null; python profiles/stalls.py $O/src_x.csv 24 > $O/ncu_stalls_xhead_tc_kernel_r2.txt; rm -f $O/src_x.csv $O/prof_x.ncu-rep
timeout 150 python profiles/xbench.py > $O/xbench_r2.txt 2>&1
# large-batch launch list (B = 16 384, L = 32)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 150 --csv --log-file $O/launches_r2_B16384_L32.csv python bench.py --batch 16384 --seq-len 32 --no-sampler --no-cpu-baseline --no-vae --no-graph --steps 2 --warmup 2 > /dev/null 2>&1
for bl in "64 32" "1024 32" "4096 32" "16384 32" "65536 32" "4096 128" "1024 512"; do set -- $bl; timeout 200 python bench.py --batch $1 --seq-len $2 --steps 5 --warmup 3 --no-sampler --no-cpu-baseline --no-vae 2>/dev/null | grep "^{" > $O/sweep_$1_$2.json; done
cat $O/sweep_*.json > $O/sweep_r2.jsonl
timeout 400 python bench.py > $O/bench_default_r2.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_r2.json 2> $O/bench_reference.err
du -sh $O
