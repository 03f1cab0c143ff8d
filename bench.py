#!/usr/bin/env python
"""
bench.py -- headline benchmark of the CL-VRNN hot path on B200 (BASELINE.json metric:
"CL-VRNN train sequences/sec and sample timesteps/sec @1/2/4/8 B200; % roofline").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL).  One JSON line is printed by rank 0.
A "step" is one training step (forward, 4 losses, backward, gradient all-reduce, Adam-WN update) of
CL-VRNN on one batch of B=200 synthetic piano-roll windows of the JSB Chorales shape (L=16, 88 keys,
C=10 keys, z=2, --use_x_prev): BASELINE.json configs[1].  Weak scaling: every rank steps its own 200.
`value` is device-timed with the batch already in HBM; `e2e` goes through the public train_on_batch
call with pinned HOST buffers (H2D of windows+labels and D2H of the loss scalars inside the timed
region).  The same line carries the sampler metric (timesteps/s of generate_sample), the roofline of
the dominant kernel, and a CPU baseline (the oracle port timed on the host cores).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=200, L=16, D=88, H=88, Z=2, C=10, use_x_prev=True)      # BASELINE.json configs[1]
SAMPLER = dict(songs_per_gpu=12500, T_seed=16, nsteps=512, C=12)      # configs[4] / 8 GPUs


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU port timing
def cpu_train_port(steps, warmup, B=None):
    """The oracle restatement (PyTorch-CPU float32, per-timestep LSTM loop, unfused losses, eager
    Adam-WN) timed on the host cores.  This is the 'reference arm': the literal Keras-2.0.0 /
    TF-1.0.1 path cannot be installed here (Python 2)."""
    import numpy as np
    import torch
    from oracle import clv_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = dict(CFG)
    if B:
        c["B"] = B
    rng = np.random.default_rng(0)
    p = O.init_vrnn_params(rng, c["L"], c["D"], c["H"], c["Z"], c["C"], True, dtype=torch.float32)
    opt = O.AdamWN(p)
    pool = O.synth_rolls(rng, c["B"] * 4, c["L"] + 1, c["D"])
    labels = rng.integers(0, c["C"], c["B"] * 4)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        i0 = (it % 4) * c["B"]
        win = torch.tensor(pool[i0:i0 + c["B"]], dtype=torch.float32)
        wt = O.one_hot(labels[i0:i0 + c["B"]], c["C"], torch.float32)
        eps_w = torch.randn(c["B"], c["C"] - 1)
        eps_z = torch.randn(c["B"], c["L"], c["Z"])
        out, g = O.vrnn_loss_and_grads(p, win[:, 1:], win[:, :-1], wt, eps_w, eps_z, c["C"], True)
        p = opt.step(p, g)
        float(out["loss"])
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=c["B"] / (ms / 1e3), ms_per_step=ms, cores=cores, B=c["B"])


def cpu_sampler_port(songs, nsteps, T_seed=16, C=12):
    """Reference-faithful Python sampling loop (batch 1, two model calls per step) on the host."""
    import numpy as np
    import torch
    from oracle import clv_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    p = O.init_vrnn_params(rng, 16, 88, 88, 2, C, True, dtype=torch.float32)
    t0 = time.perf_counter()
    for s in range(songs):
        seed = torch.tensor(O.synth_rolls(rng, 1, T_seed)[0], dtype=torch.float32)
        T = T_seed + nsteps
        O.vrnn_generate_sample(p, seed, nsteps, O.one_hot([s % C], C, torch.float32),
                               torch.randn(T, 2), torch.rand(T, 88), True)
    dt = time.perf_counter() - t0
    return songs * (T_seed + nsteps) / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_train_port(args.steps, args.warmup)
    samp = cpu_sampler_port(2, 128)
    line = {
        "impl": "reference", "metric": "cl_vrnn_train_sequences_per_sec", "value": r["value"],
        "unit": "sequences/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
                         "sample": "%d full train steps of B=200 L=16 = one rank's share of the global batch "
                                   "(oracle port, PyTorch-CPU f32, all host threads)" % args.steps},
        "e2e": {"value": r["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sampler": {"metric": "cl_vrnn_sample_timesteps_per_sec", "value": samp, "unit": "timesteps/s",
                    "sample": "2 songs x (16 seed + 128) steps, batch 1 Python loop"},
        "note": "reference = restated CPU path (oracle port); Keras 2.0.0/TF 1.0.1/Python 2 not installable here",
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "cl_vrnn train step, JSB Chorales_all shape (BASELINE configs[1]): B=200/GPU, L=16, "
                        "D=H=88, C=10, z=2, use_x_prev, adam-wn",
            "global_batch": CFG["B"] * n, "seq_len": CFG["L"], "parallelism": "dp%d" % n,
            "l2": "inputs drawn from a 176 MB resident roll pool (> 126 MB L2), fresh windows every step",
            "noise": "in-kernel Philox", "graph": "one CUDA graph per step (fwd+bwd, Adam-WN scheduled inside at N=1; all-reduce then Adam-WN at N>1), programmatic dependent launches on the critical path"}


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sampler", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fused-opt", action="store_true", help="step then Adam-WN as two calls (N=1)")
    ap.add_argument("--p2p", type=int, default=-1, help="1/0: force the fused peer-memory all-reduce+Adam path on/off")
    ap.add_argument("--batch", type=int, default=CFG["B"], help="per-GPU batch (sweep points)")
    ap.add_argument("--seq-len", type=int, default=CFG["L"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    import clvae_b200  # noqa: F401
    from clvae_b200 import _lib
    from clvae_b200._lib import lib, check, ptr
    from clvae_b200.engine import Engine
    from clvae_b200.cl_vrnn.model import get_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    devn = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=devn)
    B, L, D, H, Z, Cc = args.batch, args.seq_len, CFG["D"], CFG["H"], CFG["Z"], CFG["C"]
    K, Wm = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=devn)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- model through the public API (mirrors cl_vrnn/train.py:45-46)
    extra = {} if args.p2p < 0 else {"p2p_allreduce": bool(args.p2p)}
    if args.no_fused_opt:
        extra["fused_optimizer"] = False
    model, _ = get_model(B, D, H, Z, L, Cc, True, "adam-wn", world_size=world, rank=rank,
                         use_graph=not args.no_graph, seed=1234, **extra)
    e = model.engine
    rng = np.random.default_rng(1711_07050 + rank)
    # resident synthetic roll pool > L2: Bernoulli(0.05) on keys 15..75
    n_frames = 2_000_000
    pool = torch.zeros(n_frames, D, dtype=torch.uint8, device=devn)
    pool[:, 15:76] = (torch.rand(n_frames, 61, device=devn) < 0.05).to(torch.uint8)
    e.set_resident_roll(pool)
    n_batches = 64
    offs = torch.randint(0, n_frames - L - 2, (n_batches, B), dtype=torch.int32, device=devn)
    labs = torch.randint(0, Cc, (n_batches, B), dtype=torch.int32, device=devn)

    def step_resident(i):
        e.stage_offsets(offs[i % n_batches], labs[i % n_batches])   # D2D of 1.6 KB: batch already in HBM
        e.run(train=True, gen_noise=True)

    for i in range(Wm):
        step_resident(i)
    l0 = lib().clv_launch_count()
    step_resident(0)
    launches_eager = None
    barrier()
    clk = ClockSampler(local) if rank == 0 else None   # one poller per job: nvidia-smi takes driver locks
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        step_resident(i)
    ev1.record()
    barrier()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / K)
    clocks = clk.stop() if clk is not None else None
    losses = e.read_losses()
    value = B * world / (ms_step / 1e3)

    # ---------------- end to end: pinned host windows -> train_on_batch -> host loss scalars
    # 8 host batches in pinned memory, each as ONE buffer [windows uint8 | labels int32] (the layout
    # the public call stages with a single H2D copy; separate tensors work too, with two copies)
    win_host, lab_host = [], []
    nw = B * (L + 1) * D
    for _ in range(8):
        buf = torch.empty(nw + 4 * B, dtype=torch.uint8).pin_memory()
        buf[:nw] = torch.from_numpy((rng.random(nw) < 0.05).astype(np.uint8))
        lab = buf[nw:].view(torch.int32) if nw % 4 == 0 else torch.empty(B, dtype=torch.int32).pin_memory()
        lab.copy_(torch.from_numpy(rng.integers(0, Cc, B).astype(np.int32)))
        win_host.append(buf[:nw].view(B, L + 1, D)); lab_host.append(lab)
    for i in range(Wm):
        model.train_on_batch_windows(win_host[i % 8], lab_host[i % 8])
    barrier()
    ev0.record()
    for i in range(K):
        model.train_on_batch_windows(win_host[i % 8], lab_host[i % 8])     # returns host floats
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1) / K)
    e2e = {"value": B * world / (ms_e2e / 1e3), "unit": "sequences/s",
           "h2d_bytes_per_step": B * (L + 1) * D + 4 * B, "d2h_bytes_per_step": 32,
           "ms_per_step": ms_e2e, "api": "model.train_on_batch_windows(pinned uint8 [B,L+1,88], int32 labels)"}

    # ---------------- per-kernel timing (CUDA events on the launching stream, L2 flushed between
    # launches) of the kernels that dominate the step, on this batch shape
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    G = 4 * H
    gates = torch.randn(B, L, G, device=devn) * 0.5
    U = e.view("encoder_h.recurrent_kernel")
    hbuf = torch.zeros(B, L, H, device=devn); cbuf = torch.zeros(B, L, H, device=devn)
    dh = torch.randn(B, L, H, device=devn); dAsum = torch.zeros(B, G, device=devn)
    dZb = torch.zeros(B, L, Z, device=devn); dWb = torch.zeros(B, Cc, device=devn)
    Wv = torch.rand(B, Cc, device=devn); Zs = torch.randn(B, L, Z, device=devn)
    Kd = e.view("decoder_h.kernel"); Ke = e.view("encoder_h.kernel")
    scratch = torch.zeros(lib().clv_inproj_tc_scratch_bytes() // 4, device=devn)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=devn)

    def time_kernel(fn, reps=20):
        tot = 0.0
        for _ in range(3):
            fn()
        for _ in range(reps):
            flush.zero_()                       # flush L2 between timed launches
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps

    t_fwd = time_kernel(lambda: check(lib().clv_lstm_fwd_fused(
        ptr(gates), 1, ptr(U), ptr(e.view("decoder_h.bias")), ptr(Wv), ptr(Kd[D + Z:]), Cc, ptr(Zs),
        ptr(Kd[D:D + Z]), Z, ptr(hbuf), ptr(cbuf), B, L, H, st)))
    t_bwd = time_kernel(lambda: check(lib().clv_lstm_bwd_fused(
        ptr(gates), ptr(U), ptr(cbuf), ptr(dh), ptr(dAsum), ptr(Kd[D + Z:]), Cc, ptr(dWb), 0,
        ptr(Kd[D:D + Z]), Z, ptr(dZb), B, L, H, st)))
    t_tc = time_kernel(lambda: check(lib().clv_inproj_tc(
        ptr(e.roll), ptr(e.win_off), L, 1, D, ptr(Ke), G, G, ptr(scratch), ptr(gates), G, B * L, None, 0, 0, st)))
    # algorithmic bytes per launch: streamed operands only, weights excluded
    bytes_fwd = 4 * L * (2 * G + 2 * H + Z) * B                 # read xproj+Zs, write gates+h+c
    bytes_bwd = 4 * L * (2 * G + 3 * H + Z) * B + 4 * (G + Cc) * B   # read gates,c,dh; write dA,dZ,dAsum,dW
    bytes_tc = (D + 4 * G) * B * L                              # read uint8 roll rows, write fp32 projection
    kern = {
        "clv_lstm_bwd_fused": {"ms": t_bwd, "bytes": bytes_bwd, "GBps": bytes_bwd / t_bwd / 1e6,
                               "launches_per_step": 2, "bound": "hbm (latency bound at B=200, FFMA-issue bound at large B)"},
        "clv_lstm_fwd_fused": {"ms": t_fwd, "bytes": bytes_fwd, "GBps": bytes_fwd / t_fwd / 1e6,
                               "launches_per_step": 2, "bound": "hbm (latency bound at B=200, FFMA-issue bound at large B)"},
        "clv_inproj_tc (tcgen05)": {"ms": t_tc, "bytes": bytes_tc, "GBps": bytes_tc / t_tc / 1e6,
                                    "launches_per_step": 2, "bound": "hbm"},
    }
    dom = "clv_lstm_bwd_fused" if t_bwd >= t_fwd else "clv_lstm_fwd_fused"
    # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu --set full capture
    # (profiles/ncu_full_step_kernels_r1.md, B=200 L=16; profiles/ncu_full_large_batch_B16384_L32_r1.md)
    ncu_traffic = {(200, 16): {"clv_lstm_bwd_fused": 7.01e6, "clv_lstm_fwd_fused": 4.74e6},
                   (16384, 32): {"clv_lstm_bwd_fused": 1.845e9}}
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["GBps"], "peak": hbm_peak,
                "unit": "GB/s", "frac": kern[dom]["GBps"] / hbm_peak,
                "traffic": ncu_traffic.get((B, L), {}).get(dom),
                "peak_source": peak_src,
                "share_of_step": 2 * kern[dom]["ms"] / ms_step,
                "note": "algorithmic bytes = streamed operands of one launch (DESIGN.md section 3); timed alone with "
                        "L2 flushed, so share_of_step is an upper bound (in the step the operands are L2-resident "
                        "and the prologue overlaps the predecessor via PDL). At B=200 the recurrence is 100 CTAs x 2 "
                        "rows x 16 serial steps: latency bound, DRAM traffic below the algorithmic bytes because "
                        "the whole working set sits in the 126 MB L2; the HBM-bound kernel of the path is the "
                        "tcgen05 projection (75-80% of measured peak at B>=16384, see `kernels` and profiles/)"}

    # ---------------- sampler: generate_sample for songs_per_gpu songs, Philox noise in-kernel
    sampler = None
    if not args.no_sampler:
        S, Ts, Ns, Cs = SAMPLER["songs_per_gpu"], SAMPLER["T_seed"], SAMPLER["nsteps"], SAMPLER["C"]
        smodel, _ = get_model(1, D, H, Z, L, Cs, True, "adam-wn", seed=7, use_graph=False)
        T = Ts + Ns
        seeds = (torch.rand(S, Ts, D, device=devn) < 0.05).to(torch.uint8)
        seeds[:, :, :15] = 0; seeds[:, :, 76:] = 0
        wkey = torch.zeros(S, Cs, device=devn)
        wkey[torch.arange(S), torch.randint(0, Cs, (S,), device=devn)] = 1.0
        out = torch.zeros(S, T, D, dtype=torch.uint8, device=devn)
        cfg = smodel.engine.cfg()

        def samp():
            check(lib().clv_vrnn_sample(C.byref(cfg), ptr(smodel.engine.params), None, None, None,
                                        ptr(seeds), Ts, Ns, ptr(wkey), None, None, 99, rank * S, S,
                                        ptr(out), None, st))
        for _ in range(2):
            samp()
        barrier()
        reps = 3
        ev0.record()
        for _ in range(reps):
            samp()
        ev1.record()
        barrier()
        ms_s = max_over_ranks(ev0.elapsed_time(ev1) / reps)
        # e2e: host seeds in, host rolls out
        seeds_h = seeds.cpu().pin_memory(); w_h = wkey.cpu().pin_memory()
        out_h = torch.zeros(S, T, D, dtype=torch.uint8).pin_memory()
        barrier()
        ev0.record()
        seeds.copy_(seeds_h, non_blocking=True); wkey.copy_(w_h, non_blocking=True)
        samp()
        out_h.copy_(out, non_blocking=True)
        ev1.record()
        barrier()
        ms_se = max_over_ranks(ev0.elapsed_time(ev1))
        flop = 2 * ((D + Cs) * G + H * G + H * 2 * Z + (D + Z + Cs) * G + H * G + H * D)
        sampler = {"metric": "cl_vrnn_sample_timesteps_per_sec", "value": S * T * world / (ms_s / 1e3),
                   "unit": "timesteps/s", "ms": ms_s, "songs": S * world, "steps_per_song": T,
                   "e2e": {"value": S * T * world / (ms_se / 1e3), "unit": "timesteps/s",
                           "h2d_bytes": S * Ts * D + 4 * S * Cs, "d2h_bytes": S * T * D},
                   "fp32_tflops": flop * S * T / (ms_s / 1e3) / 1e12,
                   "note": "given one-hot key, Philox noise keyed by (seed, global song, t); "
                           "density of generated notes depends on random-init weights"}

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_train_port(steps=20, warmup=2, B=CFG["B"])
        cpu = {"value": r["value"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
               "sample": "20 train steps of B=200 L=16 on the oracle port (PyTorch-CPU f32, %d threads)" % r["cores"],
               "ms_per_step": r["ms_per_step"]}

    if rank == 0:
        line = {
            "metric": "cl_vrnn_train_sequences_per_sec", "value": value, "unit": "sequences/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(e.launches_per_step) * K, "launches_per_step": int(e.launches_per_step),
            "roofline": roofline, "kernels": kern, "cpu_baseline": cpu, "sampler": sampler,
            "final_losses": losses,
        }
        if B != CFG["B"] or L != CFG["L"]:
            line["config"]["workload"] = "cl_vrnn train step, synthetic sweep point B=%d/GPU L=%d" % (B, L)
        print(json.dumps(line), flush=True)
    # teardown: NCCL kernels captured in CUDA graphs must be released before the communicator goes
    # away; a watchdog guarantees the process exits even if the communicator teardown stalls
    sys.stdout.flush()
    torch.cuda.synchronize()
    if world > 1:
        import threading
        wd = threading.Timer(60.0, lambda: os._exit(0))   # fallback only: fires if the NCCL teardown stalls
        wd.daemon = True
        wd.start()
        dist.barrier()
        e._graphs.clear()
        del model
        torch.cuda.synchronize()
        dist.destroy_process_group()
        wd.cancel()
    # normal return: interpreter exit hooks (atexit, the harness's loaded-library record) run


if __name__ == "__main__":
    main()
