"""Cost of the peer-memory exchange fused into the Adam-WN kernels, isolated from the rest of the step:
the three bucket updates of one step (local gradients vs peers' gradients over NVLink with in-kernel flags).
torchrun --nproc-per-node N profiles/p2p_probe.py"""
import ctypes as C
import os
import sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvae_b200  # noqa: F401
from clvae_b200._lib import lib, check, ptr
from clvae_b200.engine import Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
e = Engine("vrnn", 200, L=16, D=88, H=88, Z=2, n_classes=10, use_x_prev=True, world_size=world, rank=rank,
           use_graph=False, p2p_allreduce=True)
e.init_params(__import__("numpy").random.default_rng(0))
e.gradbuf.normal_(0, 1e-3)
L_ = lib()
cfg = e.cfg()
ranges = [(11, 16, 0, 0), (0, 11, 2, 1)]      # the peer-memory schedule: [decoder | X head], then the rest


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def local_step():
    for t0, t1, slot, adv in ranges:
        check(L_.clv_adamwn_step_range(C.byref(cfg), ptr(e.params), ptr(e.grads), ptr(e.opt_state), 1e-3, 0.9, 0.999, 1e-8,
                                       1.0, 1, t0, t1, adv, None, st()))


NOWAIT = os.environ.get("CLV_P2P_NOWAIT") == "1"


def p2p_step():
    if not NOWAIT:
        check(L_.clv_p2p_wait_done(C.byref(e.p2p), ptr(e.opt_state), C.byref(cfg), st()))
    for t0, t1, slot, adv in ranges:
        check(L_.clv_adamwn_step_range_p2p(C.byref(cfg), ptr(e.params), C.byref(e.p2p), ptr(e.opt_state), 1e-3, 0.9, 0.999,
                                           1e-8, 1, t0, t1, slot, adv, None, st()))


def timeit(name, fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%-40s N=%d  %.1f us per step" % (name, world, a.elapsed_time(b) / reps * 1e3), flush=True)
    dist.barrier()
    return g


# (peer form first: its step numbers must stay in sync with the flag blocks)
g2 = timeit("2 x Adam-WN range, peers over NVLink", p2p_step)
g1 = timeit("2 x Adam-WN range, local gradients", local_step)
del g1, g2
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
