O=gpurun_out/ev2; mkdir -p $O
for bl in "200 16" "16384 32" "65536 32"; do set -- $bl; timeout 300 python profiles/kbench.py $1 $2 > $O/kbench_$1_$2.txt 2>&1; done
for bl in "64 32" "1024 32" "4096 32" "16384 32" "65536 32" "4096 128" "1024 512"; do set -- $bl; timeout 300 python bench.py --batch $1 --seq-len $2 --steps 5 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | grep "^{" > $O/sweep_$1_$2.json; done
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
