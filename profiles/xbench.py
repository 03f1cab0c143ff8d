"""X head at large batch: tcgen05 form (clv_xhead_tc) against the SIMT kernels (clv_xhead_fwd_bwd); CUDA events,
L2 flushed between launches.  python profiles/xbench.py"""
import ctypes as C
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvae_b200  # noqa: F401
from clvae_b200._lib import lib, check, ptr

L = lib()
dev = "cuda"
D = H = 88
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for R, grp in ((32768, 32), (131072, 32), (524288, 32), (2097152, 32)):
    h = torch.rand(R, H, device=dev) * 2 - 1
    Kx = torch.randn(H, D, device=dev) * 0.3
    bx = torch.randn(D, device=dev) * 0.3
    nseq = R // grp
    roll = (torch.rand(nseq * (grp + 1) + 8, D, device=dev) < 0.05).to(torch.uint8)
    off = (torch.arange(nseq, device=dev, dtype=torch.int32) * (grp + 1)).contiguous()
    dl = torch.zeros(R, D, device=dev); dh = torch.zeros(R, H, device=dev); loss = torch.zeros(8, device=dev)
    scratch = torch.zeros(int(L.clv_xhead_tc_scratch_bytes()), dtype=torch.uint8, device=dev)
    gK = torch.zeros(H, D, device=dev); gb = torch.zeros(D, device=dev)

    def tc():
        check(L.clv_xhead_tc(ptr(h), ptr(Kx), ptr(bx), ptr(roll), ptr(off), grp, 1, ptr(loss), ptr(dl), ptr(dh),
                             ptr(gK), ptr(gb), ptr(scratch), R, H, D, 1.0 / R, st))

    def simt():
        check(L.clv_xhead_fwd_bwd(ptr(h), ptr(Kx), ptr(bx), ptr(roll), ptr(off), grp, 1, ptr(loss), ptr(dl), ptr(dh),
                                  R, H, D, 1.0 / R, 1, st))

    for name, fn in (("tcgen05", tc), ("simt", simt)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        byts = R * (3 * 352 + 88)
        print("R=%8d  %-8s %8.3f ms   %7.1f GB/s algorithmic (h in, dlogits + dh out, roll in)" % (R, name, ms, byts / ms / 1e6), flush=True)
