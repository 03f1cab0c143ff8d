N=${1:-2}
[ "$N" = 2 ] && timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_p2p_check.py 2>&1 | tail -4
run() { tag=$1; shift; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-sampler --no-vae "$@" 2>gpurun_out/n${N}_$tag.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d.get('dp_parity_max_rel_err'), d.get('exchange'), d.get('p2p_exchange_rank0_us'))"; tail -3 gpurun_out/n${N}_$tag.err; }
CLV_P2P_MC=1 run mc --p2p 1
CLV_P2P_MC=0 run oneshot --p2p 1
