"""Micro-benchmark of individual libclv_b200 kernels (CUDA events, warm L2 = the in-step condition,
and cold = L2 flushed before every launch).  Usage: python profiles/kbench.py [B] [L]"""
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
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
gates = torch.randn(B, L, G, device=dev) * 0.5
U = torch.randn(H, G, device=dev) * 0.1
h = torch.zeros(B, L, H, device=dev); c = torch.zeros(B, L, H, device=dev)
dh = torch.randn(B, L, H, device=dev); dAsum = torch.zeros(B, G, device=dev)
bias = torch.zeros(G, device=dev); Wv = torch.rand(B, Cc, device=dev); Ww = torch.randn(Cc, G, device=dev) * 0.1
Zs = torch.randn(B, L, Z, device=dev); Kz = torch.randn(Z, G, device=dev) * 0.1
dZ = torch.zeros(B, L, Z, device=dev); dW = torch.zeros(B, Cc, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    warm = a.elapsed_time(b) / reps * 1e3
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    print("%-28s B=%d L=%d  warm %8.1f us   cold %8.1f us" % (name, B, L, warm, tot / 10 * 1e3), flush=True)


L_ = lib()
timeit("lstm_fwd", lambda: check(L_.clv_lstm_fwd(ptr(gates), ptr(U), ptr(h), ptr(c), None, None, B, L, H, st)))
timeit("lstm_fwd_fused(W,Z)", lambda: check(L_.clv_lstm_fwd_fused(ptr(gates), 1, ptr(U), ptr(bias), ptr(Wv), ptr(Ww), Cc,
                                                              ptr(Zs), ptr(Kz), Z, ptr(h), ptr(c), B, L, H, st)))
timeit("lstm_bwd", lambda: check(L_.clv_lstm_bwd(ptr(gates), ptr(U), ptr(h), ptr(c), ptr(dh), ptr(dAsum), B, L, H, st)))
timeit("lstm_bwd_fused(dW,dZ)", lambda: check(L_.clv_lstm_bwd_fused(ptr(gates), ptr(U), ptr(c), ptr(dh), ptr(dAsum), ptr(Ww), Cc,
                                                                ptr(dW), 0, ptr(Kz), Z, ptr(dZ), B, L, H, st)))

# ---- hoisted input projection: SIMT fp32 vs tcgen05 (bf16x3, fp32-exact)
from clvae_b200 import _lib
M = B * L
roll = (torch.rand(B * (L + 1) + 8, D, device=dev) < 0.05).to(torch.uint8)
off = (torch.arange(B, device=dev, dtype=torch.int32) * (L + 1)).contiguous()
Wk = torch.randn(D, G, device=dev) * 0.1
Cout = torch.zeros(M, G, device=dev)
scratch = torch.zeros(L_.clv_inproj_tc_scratch_bytes() // 4, device=dev)


def simt():
    a = _lib.clv_gemm_args(M=M, N=G, K=D, A=roll.data_ptr(), lda=D, a_u8=1, a_kmajor=1, a_off=off.data_ptr(),
                           a_grp=L, a_shift=1, Bm=Wk.data_ptr(), ldb=G, b_nmajor=1, C=Cout.data_ptr(), ldc=G,
                           split_k=1)
    check(L_.clv_gemm(C.byref(a), st))


timeit("inproj SIMT fp32", simt)
timeit("inproj tcgen05 bf16x3", lambda: check(L_.clv_inproj_tc(ptr(roll), ptr(off), L, 1, D, ptr(Wk), G, G, ptr(scratch),
                                                            ptr(Cout), G, M, None, 0, 0, st)))
print("   output bytes %.1f MB -> at HBM peak 6536 GB/s: %.1f us" % (M * G * 4 / 1e6, M * G * 4 / 6536e3))

# ---- LSTM weight gradients: tcgen05 (one launch) vs the three SIMT TN GEMMs
hh = torch.tanh(torch.randn(B, L, H, device=dev)); dAb = torch.randn(M, G, device=dev)
gKx = torch.zeros(D, G, device=dev); gU = torch.zeros(H, G, device=dev); gKz = torch.zeros(Z, G, device=dev)
Zsb = torch.randn(M, Z, device=dev)
timeit("lstm_wgrad tcgen05", lambda: check(L_.clv_lstm_wgrad_tc(ptr(dAb), ptr(roll), ptr(off), L, 1, D, ptr(hh), ptr(Zsb), Z,
                                                             ptr(gKx), ptr(gU), ptr(gKz), M, H, st)))
print("   operand bytes %.1f MB -> at HBM peak: %.1f us" % (M * (G * 4 + H * 4 + D) / 1e6, M * (G * 4 + H * 4 + D) / 6536e3))

# ---- tensor-core recurrence (forward)
gt = torch.randn(B, L, G, device=dev) * 0.5
uscr = torch.zeros(L_.clv_lstm_fwd_tc_scratch_bytes() // 4, device=dev)
timeit("lstm_fwd tcgen05", lambda: check(L_.clv_lstm_fwd_tc(ptr(gt), ptr(U), ptr(Zs), ptr(Kz), Z, ptr(h), ptr(c), ptr(uscr), B, L, H, st)))
print("   streamed bytes %.1f MB -> at HBM peak: %.1f us" % (B * L * (2 * G + 2 * H) * 4 / 1e6, B * L * (2 * G + 2 * H) * 4 / 6536e3))

# ---- input projection with the per-sequence addend (decoder form) and the fused X head
rb = torch.randn(B, G, device=dev)
timeit("inproj tcgen05 + rowadd", lambda: check(L_.clv_inproj_tc(ptr(roll), ptr(off), L, 0, D, ptr(Wk), G, G, ptr(scratch),
                                                              ptr(Cout), G, M, ptr(rb), G, L, st)))
Kxh = torch.randn(H, D, device=dev) * 0.1; bxh = torch.zeros(D, device=dev)
lacc = torch.zeros(8, device=dev); dlg = torch.zeros(M, D, device=dev); dhx = torch.zeros(M, H, device=dev)
timeit("xhead fwd+bwd", lambda: check(L_.clv_xhead_fwd_bwd(ptr(hh), ptr(Kxh), ptr(bxh), ptr(roll), ptr(off), L, 1, ptr(lacc),
                                                        ptr(dlg), ptr(dhx), M, H, D, 1.0 / M, 1, st)))
print("   2 x 88x88 FMA per row: %.2f GFMA -> at 36 TFMA/s: %.1f us" % (M * 2 * H * D / 1e9, M * 2 * H * D / 36e6))

# ---- Adam-WN per tensor range (the scheduled optimizer's three launches) and whole.  The partial
#      ranges do not advance `iterations`, so they re-derive the fp64 bias correction on every launch
#      here (in the step it is cached by the advancing launch): their numbers are upper bounds.
from clvae_b200.engine import Engine
eng = Engine("vrnn", 200, L=16, D=88, H=88, Z=2, n_classes=10, use_x_prev=True, use_graph=False)
cfgA = eng.cfg()
gr = torch.randn_like(eng.grads) * 1e-3
for name, t0, t1 in [("adam-wn [key encoder] 0..4", 0, 4), ("adam-wn [enc LSTM | Z heads] 4..11", 4, 11),
                     ("adam-wn [decoder | X head] 11..16", 11, 16), ("adam-wn all 0..16", 0, 16)]:
    timeit(name, lambda: check(L_.clv_adamwn_step_range(C.byref(cfgA), ptr(eng.params), ptr(gr), ptr(eng.opt_state),
                                                        1e-3, 0.9, 0.999, 1e-8, 1.0, 1, t0, t1, 1 if t1 == 16 else 0, None, st)))
