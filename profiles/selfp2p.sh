timeout 600 python -m pytest tests -m gpu -x -q --timeout=60 2>&1 | tail -3
for m in -1 1 2 -1; do timeout 100 python bench.py --no-sampler --no-vae --no-cpu --p2p $m 2>gpurun_out/self_$m.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('p2p=$m', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'])"; tail -2 gpurun_out/self_$m.err; done
