set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout=60 2>&1 | tail -5
timeout 200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_v1.json; cat gpurun_out/bench_v1.json | cut -c1-600
timeout 120 python bench.py --batch 1024 --seq-len 32 --no-sampler --no-vae --steps 100 2>&1 | tail -1 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
