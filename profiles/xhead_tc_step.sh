timeout 200 python -m pytest tests/test_gpu_step.py -m gpu -x -q --timeout=120 -k "tensor_core_x_head" 2>&1 | tail -5
for cfg in "16384 32" "4096 32" "1024 32"; do set -- $cfg
for mn in 0 999999999; do
CLV_XHEAD_TC_MIN=$mn timeout 120 python bench.py --batch $1 --seq-len $2 --steps 10 --warmup 3 --no-sampler --no-vae --no-cpu 2>/dev/null | python -c "import sys,json
d=json.loads(sys.stdin.readlines()[-1]); print('B=$1 L=$2 xhead_tc_min=$mn', d['ms_per_step'], d['launches_per_step'], d['final_losses']['loss'])"
done; done
