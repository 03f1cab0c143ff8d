"""Micro-benchmark of the wavefront pair kernels (lstm_pair.cu) against the kernels they replace
(CUDA events, warm L2 = the in-step condition).  Usage: python profiles/pbench.py [B] [L]
`pair ... (no wait)` pre-sets the hand-over counters to L so that the consumer never waits: the time
of the software-pipelined recurrence alone, without the producer->consumer latency."""
import ctypes as C
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvae_b200  # noqa: F401
from clvae_b200._lib import lib, check, ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H, G, D, Cc, Z = 88, 352, 88, 10, 2
dev = torch.device("cuda")
L_ = lib()


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)

g = torch.Generator(device=dev).manual_seed(1)


def rn(*s, scale=1.0):
    return torch.randn(*s, device=dev, generator=g) * scale


xproj_e, xproj_d = rn(B, L, G, scale=0.5), rn(B, L, G, scale=0.5)
gates_e, gates_d = torch.empty_like(xproj_e), torch.empty_like(xproj_d)
Ue, Ud = rn(H, G, scale=0.1), rn(H, G, scale=0.1)
be, bd = rn(G, scale=0.1), rn(G, scale=0.1)
Kew, Kdw, Kdz = rn(Cc, G, scale=0.1), rn(Cc, G, scale=0.1), rn(Z, G, scale=0.1)
Wv = torch.softmax(rn(B, Cc), -1).contiguous()
Kzm, Kzv, bzm, bzv = rn(H, Z, scale=0.1), rn(H, Z, scale=0.1), rn(Z, scale=0.1), rn(Z, scale=0.1)
h_e, c_e, h_d, c_d = (torch.zeros(B, L, H, device=dev) for _ in range(4))
eps = rn(B, L, Z)
Zargs, Zs = torch.zeros(B, L, 2 * Z, device=dev), torch.zeros(B, L, Z, device=dev)
loss = torch.zeros(8, device=dev)
ctr = torch.zeros(1, dtype=torch.int64, device=dev)
npairs = (B + 3) // 4
flags = torch.zeros(4 * npairs, dtype=torch.int32, device=dev)
flags_L = torch.full((4 * npairs,), L, dtype=torch.int32, device=dev)
flags_0 = torch.zeros(4 * npairs, dtype=torch.int32, device=dev)


def timeit(name, fn, reps=50):
    """fn captured into a CUDA graph (no host launch overhead in the timed region), replayed `reps` times"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record(); torch.cuda.synchronize()
    print("%-44s B=%d L=%d  %8.1f us" % (name, B, L, a.elapsed_time(b) / reps * 1e3), flush=True)


def old_fwd():
    gates_e.copy_(xproj_e); gates_d.copy_(xproj_d)
    check(L_.clv_lstm_fwd_fused(ptr(gates_e), 1, ptr(Ue), ptr(be), ptr(Wv), ptr(Kew), Cc, None, None, 0, ptr(h_e), ptr(c_e), B, L, H, st()))
    check(L_.clv_gauss_heads_fwd(ptr(h_e), ptr(Kzm), ptr(bzm), ptr(Kzv), ptr(bzv), ptr(eps), ptr(Zargs), ptr(Zs), ptr(loss),
                                 B * L, H, Z, 1.0, 0, 0, ptr(ctr), st()))
    check(L_.clv_lstm_fwd_fused(ptr(gates_d), 1, ptr(Ud), ptr(bd), ptr(Wv), ptr(Kdw), Cc, ptr(Zs), ptr(Kdz), Z, ptr(h_d), ptr(c_d), B, L, H, st()))


def pair_fwd(preset):
    gates_e.copy_(xproj_e); gates_d.copy_(xproj_d)
    if not preset:
        h_e.view(torch.int32).fill_(-1)       # 0xFFFFFFFF: the "not written yet" pattern the consumer polls for
    check(L_.clv_lstm_pair_fwd(ptr(gates_e), ptr(Ue), ptr(be), ptr(Kew), ptr(h_e), ptr(c_e), ptr(gates_d), 1, ptr(Ud), ptr(bd),
                               ptr(Kdw), ptr(Kdz), ptr(h_d), ptr(c_d), ptr(Wv), Cc, ptr(Kzm), ptr(bzm), ptr(Kzv), ptr(bzv),
                               ptr(eps), ptr(Zargs), ptr(Zs), ptr(loss), 1.0, 0, 0, ptr(ctr), B, L, H, Z, st()))


def copies():
    gates_e.copy_(xproj_e); gates_d.copy_(xproj_d); h_e.view(torch.int32).fill_(-1)


timeit("(copies of the two projections + flag fill)", copies)
timeit("old: lstm_fwd(enc) + heads + lstm_fwd(dec)", old_fwd)
ref = [t.clone() for t in (h_e, h_d, Zs, gates_d)]
timeit("pair fwd", lambda: pair_fwd(False))
for nm, a_, b_ in zip(("h_e", "h_d", "Zs", "gates_d"), ref, (h_e, h_d, Zs, gates_d)):
    print("   max |pair - old| %-8s %.3e" % (nm, (a_ - b_).abs().max().item()))
timeit("pair fwd (no wait)", lambda: pair_fwd(True))
if hasattr(L_, "clv_lstm_pair_bwd"):
    pass
