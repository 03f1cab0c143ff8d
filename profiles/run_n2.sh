run() { tag=$1; shift; timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-sampler "$@" 2>/dev/null | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d.get('dp_parity_max_rel_err'))"; }
