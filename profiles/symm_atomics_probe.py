"""Is a peer-mapped (symmetric-memory) buffer slower as a TARGET of the backward pass' red.add traffic than a
plain cudaMalloc buffer?  torchrun --nproc-per-node 2 profiles/symm_atomics_probe.py"""
import ctypes as C
import os
import sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvae_b200  # noqa: F401
from clvae_b200._lib import lib, check, ptr

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
dev = torch.device("cuda", local)
B, L, H, G, D, Z = 200, 16, 88, 352, 88, 2
M = B * L
n = D * G + H * G + Z * G
plain = torch.zeros(n, device=dev)
sym = symm_mem.empty(n, dtype=torch.float32, device=dev); sym.zero_()
hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
dA = torch.randn(M, G, device=dev); hh = torch.randn(B, L, H, device=dev); Zs = torch.randn(M, Z, device=dev)
roll = (torch.rand(B * (L + 1) + 8, D, device=dev) < 0.05).to(torch.uint8)
off = (torch.arange(B, device=dev, dtype=torch.int32) * (L + 1)).contiguous()
L_ = lib()


def run(buf):
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    gKx, gU, gKz = buf[:D * G], buf[D * G:D * G + H * G], buf[D * G + H * G:]
    check(L_.clv_lstm_wgrad_tc(ptr(dA), ptr(roll), ptr(off), L, 1, D, ptr(hh), ptr(Zs), Z, ptr(gKx), ptr(gU), ptr(gKz), M, H, st))
    check(L_.clv_colsum(ptr(dA), G, M, G, ptr(gU), 1, st))


for name, buf in (("plain", plain), ("symmetric", sym), ("plain", plain), ("symmetric", sym)):
    for _ in range(5):
        run(buf)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        run(buf)
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%-10s target: wgrad_tc + colsum %.1f us per pair" % (name, a.elapsed_time(b) / 50 * 1e3), flush=True)
dist.barrier()
dist.destroy_process_group()
