"""Sampler micro-benchmark: python profiles/sbench.py [S songs] [nsteps]"""
import ctypes as C, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvae_b200  # noqa
from clvae_b200._lib import lib, check, ptr
from clvae_b200.cl_vrnn.model import get_model
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
D, H, Z, Cs, Ts = 88, 88, 2, 12, 16
dev = torch.device("cuda")
m, _ = get_model(1, D, H, Z, 16, Cs, True, "adam-wn", seed=7, use_graph=False)
T = Ts + N
seeds = (torch.rand(S, Ts, D, device=dev) < 0.05).to(torch.uint8)
w = torch.zeros(S, Cs, device=dev); w[torch.arange(S), torch.randint(0, Cs, (S,), device=dev)] = 1
out = torch.zeros(S, T, D, dtype=torch.uint8, device=dev)
cfg = m.engine.cfg(); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    check(lib().clv_vrnn_sample(C.byref(cfg), ptr(m.engine.params), None, None, None, ptr(seeds), Ts, N, ptr(w),
                                None, None, 99, 0, S, ptr(out), None, st))
run(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print("sampler S=%d T=%d: %.2f ms  %.1f M timesteps/s  density %.3f" % (S, T, ms, S * T / ms / 1e3, out.float().mean().item()))
