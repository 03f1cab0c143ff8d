# N-GPU comparison of the exchange variants (gpurun --gpus N -- 'bash profiles/run_nN.sh N')
N=${1:-2}
run() { tag=$1; shift; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-sampler --no-vae "$@" 2>gpurun_out/n${N}_$tag.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d.get('dp_parity_max_rel_err'), d.get('host_enqueue_ms_per_step'), d.get('p2p_exchange_rank0_us'))"; tail -3 gpurun_out/n${N}_$tag.err; }
run p2p --p2p 1
run nccl --p2p 0
run p2p400 --p2p 1 --batch 400
run nccl400 --p2p 0 --batch 400
