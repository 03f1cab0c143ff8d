run() { tag=$1; shift; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-sampler --no-vae "$@" 2>gpurun_out/d_$tag.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'], d.get('p2p_exchange_rank0_us'))"; tail -2 gpurun_out/d_$tag.err; }
CLV_P2P_DIAG_SELF=1 run symmself --p2p 1
