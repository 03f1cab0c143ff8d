import sys, os, ctypes as C, torch
sys.path.insert(0, '.')
from clvae_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from clvae_b200._lib import lib, check, ptr
B, L, H, G, Z, D = 200, 16, 88, 352, 2, 88
dev = 'cuda'; L_ = lib()
L_.clv_debug_wprof.argtypes = [C.c_void_p, C.c_int]
M = B * L
roll = (torch.rand(B * (L + 1) + 8, D, device=dev) < 0.05).to(torch.uint8)
off = (torch.arange(B, device=dev, dtype=torch.int32) * (L + 1)).contiguous()
hh = torch.tanh(torch.randn(B, L, H, device=dev)); dAb = torch.randn(M, G, device=dev)
gKx = torch.zeros(D, G, device=dev); gU = torch.zeros(H, G, device=dev); gKz = torch.zeros(Z, G, device=dev)
Zsb = torch.randn(M, Z, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
fn = lambda: check(L_.clv_lstm_wgrad_tc(ptr(dAb), ptr(roll), ptr(off), L, 1, D, ptr(hh), ptr(Zsb), Z, ptr(gKx), ptr(gU), ptr(gKz), M, H, st))
out = (C.c_longlong * 16)()
for _ in range(5): fn()
torch.cuda.synchronize()
for rep in range(3):
    fn(); torch.cuda.synchronize()
    L_.clv_debug_wprof(out, 0)
    v = list(out)
    print("cycles from start: setup %d | stage0 %d stage1 %d stage2 %d stage3 %d | mma done %d | epilogue done %d | all warps %d" % tuple(x - v[0] for x in v[1:9]))
    print("   bulk copies of slab 0 / 1 issued at %d / %d; raw slab 0 / 1 / 2 landed at %d / %d / %d" % tuple(x - v[0] for x in v[9:14]))
