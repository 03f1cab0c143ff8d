# 2-GPU comparison of the exchange variants (gpurun --gpus 2 -- 'bash profiles/run_n2.sh')
run() { tag=$1; shift; timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-sampler --no-vae "$@" 2>gpurun_out/n2_$tag.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d.get('dp_parity_max_rel_err'), d.get('p2p_exchange_rank0_us'))"; tail -3 gpurun_out/n2_$tag.err; }
run p2p --p2p 1
run nccl --p2p 0
run p2p --p2p 1
