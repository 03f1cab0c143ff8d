# the driver's scaling command at N GPUs, both arms: bash profiles/scale_final.sh N
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>/dev/null | grep "^{" | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N 2>gpurun_out/scale_n$N.err | grep "^{" > gpurun_out/bench_n${N}_r2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n${N}_r2.json'))
print({k: d.get(k) for k in ('n_gpus','ms_per_step','value','dp_parity_max_rel_err','exchange','p2p_exchange_rank0_us','host_enqueue_ms_per_step')}); print(d['e2e']); print(d['sampler']['value'] if d.get('sampler') else None, d.get('clocks'))"
tail -3 gpurun_out/scale_n$N.err
