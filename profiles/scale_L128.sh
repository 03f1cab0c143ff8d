# weak scaling at the top-of-sweep sequence length (BASELINE configs[3]: L=128): bash profiles/scale_L128.sh "1 2"
for N in ${1:-1 2}; do
if [ $N = 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"; fi
timeout 200 $CMD --batch 1024 --seq-len 128 --steps 30 --warmup 5 --no-sampler --no-vae --no-cpu 2>/dev/null | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({k: d.get(k) for k in ('n_gpus','ms_per_step','value','e2e','dp_parity_max_rel_err','exchange','p2p_exchange_rank0_us')}))"
done
