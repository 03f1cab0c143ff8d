timeout 300 python -m pytest tests -m gpu -x -q --timeout=120 -k "lstm or tc or step" 2>&1 | tail -3
for cfg in "16384 32" "65536 32"; do set -- $cfg
timeout 120 python bench.py --batch $1 --seq-len $2 --steps 10 --warmup 3 --no-sampler --no-vae --no-cpu 2>/dev/null | python -c "import sys,json
d=json.loads(sys.stdin.readlines()[-1]); print('B=$1 L=$2', d['ms_per_step'], d['launches_per_step'], d['final_losses']['loss'])"
done
