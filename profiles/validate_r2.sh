set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout=60 2>&1 | tail -3
timeout 200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_v1.json; cut -c1-300 gpurun_out/bench_v1.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_p2p_check.py 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 2>&1 | tail -1 > gpurun_out/bench_v1_n2.json; cut -c1-300 gpurun_out/bench_v1_n2.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
fi
