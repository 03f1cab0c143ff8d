import sys, ctypes as C, torch
sys.path.insert(0, '.')
import os
from clvae_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from clvae_b200._lib import lib, check, ptr
B, L, H, G, Z, Cc = 200, 16, 88, 352, 2, 10
dev = 'cuda'
L_ = lib()
L_.clv_debug_prof.argtypes = [C.c_void_p, C.c_int]
gates = torch.randn(B, L, G, device=dev) * 0.5
U = torch.randn(H, G, device=dev) * 0.1
h = torch.zeros(B, L, H, device=dev); c = torch.zeros(B, L, H, device=dev)
bias = torch.zeros(G, device=dev); Wv = torch.rand(B, Cc, device=dev); Ww = torch.randn(Cc, G, device=dev) * .1
Zs = torch.randn(B, L, Z, device=dev); Kz = torch.randn(Z, G, device=dev) * .1
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = (C.c_longlong * 16)()
for name, fn in [("enc-like", lambda: check(L_.clv_lstm_fwd_fused(ptr(gates), 1, ptr(U), ptr(bias), ptr(Wv), ptr(Ww), Cc, None, None, 0, ptr(h), ptr(c), B, L, H, st))),
                 ("dec-like", lambda: check(L_.clv_lstm_fwd_fused(ptr(gates), 1, ptr(U), ptr(bias), ptr(Wv), ptr(Ww), Cc, ptr(Zs), ptr(Kz), Z, ptr(h), ptr(c), B, L, H, st)))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    L_.clv_debug_prof(None, 1)
    N = 20
    for _ in range(N): fn()
    torch.cuda.synchronize()
    L_.clv_debug_prof(out, 0)
    v = [x / N / L for x in out]
    print(name, "cycles/step: top %.0f matvec+reduce %.0f cell %.0f barrier %.0f  total %.0f" % (v[0], v[1], v[2], v[3], sum(v[:4])))
