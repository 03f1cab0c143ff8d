set -x
mkdir -p gpurun_out/ev
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 130 --csv --log-file gpurun_out/ev/launches_warm.csv python bench.py --steps 2 --warmup 2 --no-sampler --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/ev/launches_cold.csv python bench.py --steps 2 --warmup 2 --no-sampler --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 58 -c 29 -o gpurun_out/ev/prof_step_r1 -f python bench.py --steps 2 --warmup 2 --no-sampler --no-cpu-baseline --no-graph > gpurun_out/ev/ncu_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_wgrad_tc|lstm_fwd_tc|inproj_tc|xhead|lstm_bwd|adamwn|keyenc" -c 14 -o gpurun_out/ev/prof_big_r1 -f python bench.py --batch 16384 --seq-len 32 --steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-graph > gpurun_out/ev/ncu_big.log 2>&1
for bl in "200 16" "16384 32" "65536 32"; do set -- $bl; timeout 300 python profiles/kbench.py $1 $2 > gpurun_out/ev/kbench_$1_$2.txt 2>&1; done
for bl in "64 32" "1024 32" "4096 32" "16384 32" "65536 32" "4096 128" "1024 512"; do set -- $bl; timeout 300 python bench.py --batch $1 --seq-len $2 --steps 5 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/ev/sweep_$1_$2.json; done
timeout 600 python bench.py > gpurun_out/ev/bench_default.json 2> gpurun_out/ev/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev/bench_reference.json 2> gpurun_out/ev/bench_reference.err
tail -c 600 gpurun_out/ev/bench_default.json
