"""Per-kernel parity through the C-ABI against the oracle (float64 CPU) on seeded inputs.
Tolerance: 1e-4 relative (north_star: fp32 losses/gradients within 1e-4)."""
import ctypes as C
import numpy as np
import pytest
import torch

import util
from oracle import clv_oracle as O, manual_bwd as M

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _env():
    from clvae_b200 import _lib
    from clvae_b200._lib import lib, check, ptr
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return _lib, lib(), check, ptr, st


_KEEP = []   # device tensors must outlive the asynchronous launches that read them


@pytest.fixture(autouse=True)
def _clear_keep():
    _KEEP.clear()
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def dev(a, dtype=torch.float32):
    t = torch.tensor(np.ascontiguousarray(a), dtype=dtype).cuda()
    _KEEP.append(t)
    return t


def gemm(**kw):
    _lib, L, check, ptr, st = _env()
    a = _lib.clv_gemm_args(a_kmajor=1, b_nmajor=1, split_k=1)
    keep = []
    for k, v in kw.items():
        if torch.is_tensor(v):
            keep.append(v)
            v = v.data_ptr()
        setattr(a, k, v)
    check(L.clv_gemm(C.byref(a), st), "clv_gemm")
    torch.cuda.synchronize()


@pytest.mark.parametrize("M_,N,K", [(200, 88, 1408), (3200, 352, 88), (37, 18, 88), (1, 4, 5),
                                    (130, 65, 17)])
def test_gemm_nn_bias_relu_rowadd(M_, N, K):
    rng = np.random.default_rng(M_ + N + K)
    A = rng.normal(size=(M_, K)); B = rng.normal(size=(K, N)); bias = rng.normal(size=N)
    grp = 1 if M_ % 8 else 8
    ra = rng.normal(size=((M_ + grp - 1) // grp, N))
    ref = np.maximum(A @ B + bias + np.repeat(ra, grp, axis=0)[:M_], 0)
    Cd = torch.zeros(M_, N, device="cuda")
    gemm(M=M_, N=N, K=K, A=dev(A), lda=K, Bm=dev(B), ldb=N, C=Cd, ldc=N, bias=dev(bias),
         rowadd=dev(ra), ldra=N, ra_grp=grp, relu=1)
    assert util.rel_err(Cd.cpu().numpy(), ref) < TOL
    # accumulate on top
    Cd2 = Cd.clone()
    gemm(M=M_, N=N, K=K, A=dev(A), lda=K, Bm=dev(B), ldb=N, C=Cd2, ldc=N, accumulate=1)
    assert util.rel_err(Cd2.cpu().numpy(), ref + A @ B) < TOL


def test_gemm_u8_gather_nn_and_tn():
    rng = np.random.default_rng(7)
    B_, L, D, N = 9, 5, 88, 40
    roll = (rng.random((200, D)) < 0.1).astype(np.uint8)
    off = rng.integers(0, 200 - L - 1, B_).astype(np.int32)
    Wt = rng.normal(size=(D, N))
    X = np.stack([roll[o + 1:o + 1 + L] for o in off]).astype(np.float64)       # shift 1
    Cd = torch.zeros(B_ * L, N, device="cuda")
    gemm(M=B_ * L, N=N, K=D, A=dev(roll, torch.uint8), lda=D, a_u8=1, a_off=dev(off, torch.int32),
         a_grp=L, a_shift=1, Bm=dev(Wt), ldb=N, C=Cd, ldc=N)
    assert util.rel_err(Cd.cpu().numpy(), X.reshape(-1, D) @ Wt) < TOL
    # flat rows (hW): K = L*D contiguous
    Wf = rng.normal(size=(L * D, 16))
    Cd = torch.zeros(B_, 16, device="cuda")
    gemm(M=B_, N=16, K=L * D, A=dev(roll, torch.uint8), lda=D, a_u8=1, a_off=dev(off, torch.int32),
         a_grp=1, a_shift=1, Bm=dev(Wf), ldb=16, C=Cd, ldc=16)
    assert util.rel_err(Cd.cpu().numpy(), X.reshape(B_, -1) @ Wf) < TOL
    # TN (wgrad) with split-K atomics into a pre-filled buffer
    dC = rng.normal(size=(B_ * L, N))
    base = rng.normal(size=(D, N))
    Gd = dev(base)
    gemm(M=D, N=N, K=B_ * L, A=dev(roll, torch.uint8), lda=D, a_u8=1, a_kmajor=0,
         a_off=dev(off, torch.int32), a_grp=L, a_shift=1, Bm=dev(dC), ldb=N, C=Gd, ldc=N, split_k=3)
    assert util.rel_err(Gd.cpu().numpy(), base + X.reshape(-1, D).T @ dC) < TOL


def test_gemm_nt_mask_and_tn_shift():
    rng = np.random.default_rng(8)
    M_, N, K = 77, 88, 18
    dC = rng.normal(size=(M_, K)); Wt = rng.normal(size=(N, K)); mask = rng.normal(size=(M_, N))
    Cd = torch.zeros(M_, N, device="cuda")
    gemm(M=M_, N=N, K=K, A=dev(dC), lda=K, Bm=dev(Wt), ldb=K, b_nmajor=0, C=Cd, ldc=N,
         relu_mask=dev(mask), ldmask=N)
    assert util.rel_err(Cd.cpu().numpy(), (dC @ Wt.T) * (mask > 0)) < TOL
    # TN with the one-step shift used for dU = Hprev^T @ dA
    B_, L, H, G = 6, 7, 88, 352
    h = rng.normal(size=(B_, L, H)); dA = rng.normal(size=(B_, L, G))
    hprev = np.concatenate([np.zeros((B_, 1, H)), h[:, :-1]], axis=1)
    Gd = torch.zeros(H, G, device="cuda")
    gemm(M=H, N=G, K=B_ * L, A=dev(h), lda=H, a_kmajor=0, a_row_delta=-1, a_skip_grp=L, Bm=dev(dA),
         ldb=G, C=Gd, ldc=G, split_k=4)
    assert util.rel_err(Gd.cpu().numpy(), hprev.reshape(-1, H).T @ dA.reshape(-1, G)) < TOL


def test_colsum():
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(9)
    A = rng.normal(size=(1000, 352))
    out = dev(np.ones(352))
    check(L.clv_colsum(ptr(dev(A)), 352, 1000, 352, ptr(out), 1, st))
    torch.cuda.synchronize()
    assert util.rel_err(out.cpu().numpy(), 1 + A.sum(0)) < TOL


@pytest.mark.parametrize("B_,Cc", [(200, 10), (33, 2), (5, 16)])
def test_logitnormal_fwd_bwd(B_, Cc):
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(B_)
    C1 = Cc - 1
    Wargs = rng.normal(0, 0.7, size=(B_, 2 * C1)); eps = rng.standard_normal((B_, C1))
    labels = rng.integers(0, Cc, B_).astype(np.int32)
    wt = O.one_hot(labels, Cc).numpy()
    prior = 0.3
    W, wkl, wrec, corr = M.logitnormal_fwd(Wargs[:, :C1], Wargs[:, C1:], eps, wt, Cc, prior)
    Wd = torch.zeros(B_, Cc, device="cuda"); loss = torch.zeros(8, device="cuda")
    epsd = dev(eps)
    check(L.clv_logitnormal_fwd(ptr(dev(Wargs)), 2 * C1, ptr(epsd), ptr(dev(labels, torch.int32)),
                                ptr(Wd), ptr(loss), B_, Cc, prior, 1.0 / B_, 0, 0, None, st))
    torch.cuda.synchronize()
    assert util.rel_err(Wd.cpu().numpy(), W) < TOL
    lo = loss.cpu().numpy()
    assert abs(lo[1] - wkl.mean()) < TOL * abs(wkl.mean())
    assert abs(lo[2] - wrec.mean()) < TOL * abs(wrec.mean())
    assert abs(lo[4] - corr.mean()) < 1e-6
    dW_ext = rng.normal(size=(B_, Cc))
    dWm, dWlv = M.logitnormal_bwd(Wargs[:, :C1], Wargs[:, C1:], eps, wt, W, dW_ext, Cc, prior,
                                  0.7 / B_, 0.9 / B_)
    dWa = torch.zeros(B_, 2 * C1, device="cuda")
    check(L.clv_logitnormal_bwd(ptr(dev(Wargs)), 2 * C1, ptr(epsd), ptr(dev(labels, torch.int32)),
                                ptr(dev(W)), ptr(dev(dW_ext)), ptr(dWa), B_, Cc, prior, 0.7 / B_,
                                0.9 / B_, st))
    torch.cuda.synchronize()
    assert util.rel_err(dWa.cpu().numpy(), np.concatenate([dWm, dWlv], -1)) < TOL


@pytest.mark.parametrize("R,Z,relu_in", [(3200, 2, 0), (100, 4, 1), (17, 16, 0), (4099, 1, 0), (33, 2, 0)])
def test_gauss_heads_fwd_bwd(R, Z, relu_in):
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(R + Z)
    H = 88
    h = rng.normal(size=(R, H))
    if relu_in:
        h = np.maximum(h, 0)
    Km = rng.normal(0, 0.1, (H, Z)); Kv = rng.normal(0, 0.1, (H, Z))
    bm = rng.normal(0, 0.1, Z); bv = rng.normal(0, 0.1, Z); eps = rng.standard_normal((R, Z))
    mu = h @ Km + bm; lv = h @ Kv + bv
    Zs = mu + np.exp(lv / 2) * eps
    kl = -0.5 * (1 + lv - mu ** 2 - np.exp(lv)).sum(-1)
    Za = torch.zeros(R, 2 * Z, device="cuda"); Zd = torch.zeros(R, Z, device="cuda")
    loss = torch.zeros(8, device="cuda"); epsd = dev(eps); hd = dev(h); Kmd = dev(Km); Kvd = dev(Kv)
    check(L.clv_gauss_heads_fwd(ptr(hd), ptr(Kmd), ptr(dev(bm)), ptr(Kvd), ptr(dev(bv)), ptr(epsd),
                                ptr(Za), ptr(Zd), ptr(loss), R, H, Z, 1.0 / R, 0, 0, None, st))
    torch.cuda.synchronize()
    assert util.rel_err(Za.cpu().numpy(), np.concatenate([mu, lv], -1)) < TOL
    assert util.rel_err(Zd.cpu().numpy(), Zs) < TOL
    assert abs(loss.cpu().numpy()[3] - kl.mean()) < TOL * abs(kl.mean())
    dZ = rng.normal(size=(R, Z)); klw = 0.3 / R
    dmu = dZ + klw * mu
    dlv = dZ * eps * 0.5 * np.exp(lv / 2) + klw * 0.5 * (np.exp(lv) - 1)
    dh = dmu @ Km.T + dlv @ Kv.T
    if relu_in:
        dh = dh * (h > 0)
    dhd = torch.zeros(R, H, device="cuda")
    gKm = torch.zeros(H, Z, device="cuda"); gKv = torch.zeros(H, Z, device="cuda")
    gbm = torch.zeros(Z, device="cuda"); gbv = torch.zeros(Z, device="cuda")
    check(L.clv_gauss_heads_bwd(ptr(hd), ptr(Kmd), ptr(Kvd), ptr(epsd), ptr(Za), ptr(dev(dZ)), ptr(dhd),
                                ptr(gKm), ptr(gbm), ptr(gKv), ptr(gbv), R, H, Z, klw, relu_in, st))
    torch.cuda.synchronize()
    assert util.rel_err(dhd.cpu().numpy(), dh) < TOL
    assert util.rel_err(gKm.cpu().numpy(), h.T @ dmu) < TOL
    assert util.rel_err(gKv.cpu().numpy(), h.T @ dlv) < TOL
    assert util.rel_err(gbm.cpu().numpy(), dmu.sum(0)) < TOL
    assert util.rel_err(gbv.cpu().numpy(), dlv.sum(0)) < TOL
    if not relu_in:
        # weight gradients only (dh = NULL): the form the CL-VRNN step uses; H = 88, Z <= 2 takes the large-batch kernel
        for t_ in (gKm, gKv, gbm, gbv):
            t_.zero_()
        check(L.clv_gauss_heads_bwd(ptr(hd), ptr(Kmd), ptr(Kvd), ptr(epsd), ptr(Za), ptr(dev(dZ)), None,
                                    ptr(gKm), ptr(gbm), ptr(gKv), ptr(gbv), R, H, Z, klw, 0, st))
        torch.cuda.synchronize()
        assert util.rel_err(gKm.cpu().numpy(), h.T @ dmu) < TOL
        assert util.rel_err(gKv.cpu().numpy(), h.T @ dlv) < TOL
        assert util.rel_err(gbm.cpu().numpy(), dmu.sum(0)) < TOL
        assert util.rel_err(gbv.cpu().numpy(), dlv.sum(0)) < TOL


@pytest.mark.parametrize("B_,L", [(200, 16), (3, 5), (1, 1), (19, 33), (1200, 4)])
def test_lstm_fwd_bwd(B_, L):
    _lib, Lb, check, ptr, st = _env()
    rng = np.random.default_rng(B_ * 100 + L)
    H, G = 88, 352
    xproj = rng.normal(0, 1.2, size=(B_, L, G)); U = rng.normal(0, 0.15, size=(H, G))
    hs, cs, a_all = M.lstm_fwd(xproj, U)
    gates = dev(xproj); Ud = dev(U)
    hd = torch.zeros(B_, L, H, device="cuda"); cd = torch.zeros(B_, L, H, device="cuda")
    check(Lb.clv_lstm_fwd(ptr(gates), ptr(Ud), ptr(hd), ptr(cd), None, None, B_, L, H, st))
    torch.cuda.synchronize()
    assert util.rel_err(hd.cpu().numpy(), hs) < TOL
    assert util.rel_err(cd.cpu().numpy(), cs) < TOL
    g_ref = np.concatenate([M._hard_sigmoid(a_all[..., :2 * H]), np.tanh(a_all[..., 2 * H:3 * H]),
                            M._hard_sigmoid(a_all[..., 3 * H:])], -1)
    assert util.rel_err(gates.cpu().numpy(), g_ref) < TOL
    dh_out = rng.normal(size=(B_, L, H))
    dA = M.lstm_bwd(dh_out, hs, cs, a_all, U)
    dAsum = torch.zeros(B_, G, device="cuda")
    check(Lb.clv_lstm_bwd(ptr(gates), ptr(Ud), ptr(hd), ptr(cd), ptr(dev(dh_out)), ptr(dAsum), B_, L,
                          H, st))
    torch.cuda.synchronize()
    assert util.rel_err(gates.cpu().numpy(), dA) < TOL
    assert util.rel_err(dAsum.cpu().numpy(), dA.sum(1)) < TOL


def test_lstm_saturated_gates_have_zero_gradient():
    _lib, Lb, check, ptr, st = _env()
    H, G, B_, L = 88, 352, 2, 3
    xproj = np.zeros((B_, L, G)); xproj[..., :H] = 9.0; xproj[..., 3 * H:] = -9.0   # i clipped to 1, o to 0
    gates = dev(xproj); Ud = dev(np.zeros((H, G)))
    hd = torch.zeros(B_, L, H, device="cuda"); cd = torch.zeros(B_, L, H, device="cuda")
    check(Lb.clv_lstm_fwd(ptr(gates), ptr(Ud), ptr(hd), ptr(cd), None, None, B_, L, H, st))
    dAsum = torch.zeros(B_, G, device="cuda")
    check(Lb.clv_lstm_bwd(ptr(gates), ptr(Ud), ptr(hd), ptr(cd), ptr(dev(np.ones((B_, L, H)))),
                          ptr(dAsum), B_, L, H, st))
    torch.cuda.synchronize()
    g = gates.cpu().numpy()
    assert np.all(g[..., :H] == 0) and np.all(g[..., 3 * H:] == 0)
    assert np.all(hd.cpu().numpy() == 0)


@pytest.mark.parametrize("R", [3200, 7])
def test_bernoulli_ce(R):
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(R)
    D = 88
    logits = rng.normal(0, 3, size=(R, D)); logits[0, :4] = [30, -30, 17.5, -17.5]   # clip region
    roll = (rng.random((R + 5, D)) < 0.1).astype(np.uint8)
    off = np.arange(R, dtype=np.int32) + 2
    x = roll[off + 1].astype(np.float64)
    # float32 oracle: Keras computes this loss in fp32, where the clip bound 1-1e-7 rounds to
    # 1-2^-23; the float64 value of the clipped logit differs by 1% (16.12 vs 15.94)
    loss_ref, dl_ref = M.bernoulli_fwd_bwd(logits.astype(np.float32), x.astype(np.float32), np.float32(1.0 / R))
    loss_ref = loss_ref.astype(np.float64)
    lg = dev(logits); loss = torch.zeros(8, device="cuda")
    check(L.clv_bernoulli_ce_fwd_bwd(ptr(lg), ptr(dev(roll, torch.uint8)), ptr(dev(off, torch.int32)),
                                     1, 1, ptr(loss), R, D, 1.0 / R, 1, st))
    torch.cuda.synchronize()
    assert abs(loss.cpu().numpy()[0] - loss_ref.mean()) < TOL * loss_ref.mean()
    assert util.rel_err(lg.cpu().numpy(), dl_ref) < TOL
    assert lg.cpu().numpy()[0, 0] == 0.0 and lg.cpu().numpy()[0, 1] == 0.0


@pytest.mark.parametrize("weightnorm", [1, 0])
def test_adamwn_matches_oracle_trajectory(weightnorm):
    from clvae_b200.engine import Engine
    rng = np.random.default_rng(11)
    e = Engine("vrnn", 4, L=3, D=88, H=88, Z=2, n_classes=5, use_x_prev=True,
               optimizer="adam-wn" if weightnorm else "adam", use_graph=False)
    p0 = e.init_params(rng)
    params = {k: torch.tensor(v, dtype=torch.float64) for k, v in p0.items()}
    if weightnorm:
        opt = O.AdamWN(params)
    else:
        opt = O.AdamWN({k: v.reshape(-1) for k, v in params.items()})
        params = {k: v.reshape(-1) for k, v in params.items()}
    _lib, L, check, ptr, st = _env()
    for step in range(4):
        g = {k: rng.normal(0, 10.0 ** rng.integers(-3, 1), size=tuple(v.shape)) for k, v in params.items()}
        for k in e.names:
            e.grad_view(k).copy_(dev(g[k]).view_as(e.grad_view(k)))
        cfg = e.cfg()
        check(L.clv_adamwn_step(C.byref(cfg), ptr(e.params), ptr(e.grads), ptr(e.opt_state), 1e-3, 0.9,
                                0.999, 1e-8, 1.0, weightnorm, st))
        torch.cuda.synchronize()
        params = opt.step(params, {k: torch.tensor(v) for k, v in g.items()})
        got = e.get_params()
        for k in e.names:
            assert util.rel_err(got[k].reshape(-1), params[k].numpy().reshape(-1)) < TOL, (step, k)
    assert e.iterations == 4


# ---------------------------------------------------------------------------- fused kernels
@pytest.mark.parametrize("R,grp", [(3200, 16), (45, 5), (20031 // 33 * 33, 33)])   # last: the 64-row-tile kernel
def test_xhead_fused_matches_gemm_plus_bernoulli(R, grp):
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(R)
    D = H = 88
    h = rng.normal(0, 1.0, size=(R, H)); Kx = rng.normal(0, 0.3, size=(H, D)); bx = rng.normal(0, 0.3, D)
    roll = (rng.random((R + R // grp + 40, D)) < 0.1).astype(np.uint8)
    nseq = R // grp
    off = (np.arange(nseq) * (grp + 1)).astype(np.int32)
    rows = (off[:, None] + 1 + np.arange(grp)[None, :]).reshape(-1)
    x = roll[rows].astype(np.float32)
    logits = (h @ Kx + bx)
    loss_ref, dl_ref = M.bernoulli_fwd_bwd(logits.astype(np.float32), x, np.float32(1.0 / R))
    dh_ref = dl_ref.astype(np.float64) @ Kx.T
    dl = torch.zeros(R, D, device="cuda"); dh = torch.zeros(R, H, device="cuda")
    loss = torch.zeros(8, device="cuda")
    check(L.clv_xhead_fwd_bwd(ptr(dev(h)), ptr(dev(Kx)), ptr(dev(bx)), ptr(dev(roll, torch.uint8)),
                              ptr(dev(off, torch.int32)), grp, 1, ptr(loss), ptr(dl), ptr(dh), R, H, D,
                              1.0 / R, 1, st))
    torch.cuda.synchronize()
    assert abs(loss.cpu().numpy()[0] - loss_ref.astype(np.float64).mean()) < TOL * loss_ref.mean()
    assert util.rel_err(dl.cpu().numpy(), dl_ref) < TOL
    assert util.rel_err(dh.cpu().numpy(), dh_ref) < TOL


@pytest.mark.parametrize("R,grp", [(3200, 16), (45, 5), (128 * 150 * 3 + 33, 33), (77 * 7, 7)])
def test_xhead_tensor_core_matches_gemm_plus_bernoulli(R, grp):
    """clv_xhead_tc (two chained tcgen05 GEMMs, bf16 hi+mid operand splits) against float64 numpy: partial last
    tile, more tiles than SMs (both pipeline stages of every CTA reused), sequence groups crossing tiles."""
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(R + 1)
    D = H = 88
    h = rng.uniform(-1, 1, size=(R, H)); Kx = rng.normal(0, 0.3, size=(H, D)); bx = rng.normal(0, 0.3, D)
    nseq = (R + grp - 1) // grp
    roll = (rng.random((nseq * (grp + 1) + 40, D)) < 0.1).astype(np.uint8)
    off = (np.arange(nseq) * (grp + 1)).astype(np.int32)
    rows = (off[:, None] + 1 + np.arange(grp)[None, :]).reshape(-1)[:R]
    x = roll[rows].astype(np.float32)
    logits = (h @ Kx + bx)
    loss_ref, dl_ref = M.bernoulli_fwd_bwd(logits.astype(np.float32), x, np.float32(1.0 / R))
    dh_ref = dl_ref.astype(np.float64) @ Kx.T
    dl = torch.full((R, D), 7.0, device="cuda"); dh = torch.full((R, H), 7.0, device="cuda")
    loss = torch.zeros(8, device="cuda")
    scratch = torch.zeros(int(L.clv_xhead_tc_scratch_bytes()), dtype=torch.uint8, device="cuda")
    hd, Kd, bd, rd, od = dev(h), dev(Kx), dev(bx), dev(roll, torch.uint8), dev(off, torch.int32)
    gK = torch.zeros(H, D, device="cuda"); gb = torch.zeros(D, device="cuda")
    for _ in range(2):          # twice: the second call reuses the barriers' phases from scratch
        loss.zero_(); gK.zero_(); gb.zero_()
        check(L.clv_xhead_tc(ptr(hd), ptr(Kd), ptr(bd), ptr(rd), ptr(od), grp, 1, ptr(loss), ptr(dl), ptr(dh),
                             ptr(gK), ptr(gb), ptr(scratch), R, H, D, 1.0 / R, st))
        torch.cuda.synchronize()
    # the head's weight / bias gradients from the third GEMM
    assert util.rel_err(gK.cpu().numpy(), h.T @ dl_ref.astype(np.float64)) < TOL
    assert util.rel_err(gb.cpu().numpy(), dl_ref.astype(np.float64).sum(0)) < TOL
    assert abs(loss.cpu().numpy()[0] - loss_ref.astype(np.float64).mean()) < TOL * loss_ref.mean()
    assert util.rel_err(dl.cpu().numpy(), dl_ref) < TOL
    assert util.rel_err(dh.cpu().numpy(), dh_ref) < TOL
    # and against the SIMT kernel it replaces, much tighter than the parity tolerance
    dl2 = torch.zeros(R, D, device="cuda"); dh2 = torch.zeros(R, H, device="cuda"); loss2 = torch.zeros(8, device="cuda")
    check(L.clv_xhead_fwd_bwd(ptr(hd), ptr(Kd), ptr(bd), ptr(rd), ptr(od), grp, 1, ptr(loss2), ptr(dl2), ptr(dh2),
                              R, H, D, 1.0 / R, 1, st))
    torch.cuda.synchronize()
    assert util.rel_err(dl.cpu().numpy(), dl2.cpu().numpy()) < 2e-5
    assert util.rel_err(dh.cpu().numpy(), dh2.cpu().numpy()) < 2e-5
    assert abs(float(loss[0]) - float(loss2[0])) < 1e-5 * abs(float(loss2[0]))


@pytest.mark.parametrize("B_,Lq,Cc,dens", [(200, 16, 10, 0.05), (7, 3, 2, 0.5), (3, 64, 16, 0.0), (5, 128, 10, 0.004)])   # last: 34 KB dynamic + 17 KB static smem (opt-in path)
def test_keyenc_fused_fwd_bwd(B_, Lq, Cc, dens):
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(B_ + Lq)
    D, C1 = 88, Cc - 1
    roll = (rng.random((B_ * (Lq + 1) + 3, D)) < dens).astype(np.uint8)
    off = (np.arange(B_) * (Lq + 1)).astype(np.int32)
    X = np.stack([roll[o + 1:o + 1 + Lq] for o in off]).reshape(B_, -1).astype(np.float64)
    Khw = rng.normal(0, 0.1, (Lq * D, D)); bhw = rng.normal(0, 0.1, D)
    Kwa = rng.normal(0, 0.3, (D, 2 * C1)); bwa = rng.normal(0, 0.1, 2 * C1)
    eps = rng.standard_normal((B_, C1)); labels = rng.integers(0, Cc, B_).astype(np.int32)
    wt = O.one_hot(labels, Cc).numpy()
    hW = np.maximum(X @ Khw + bhw, 0); Wa = hW @ Kwa + bwa
    W, wkl, wrec, corr = M.logitnormal_fwd(Wa[:, :C1], Wa[:, C1:], eps, wt, Cc, 0.1)
    hWd = torch.zeros(B_, D, device="cuda"); Wad = torch.zeros(B_, 2 * C1, device="cuda")
    Wd = torch.zeros(B_, Cc, device="cuda"); loss = torch.zeros(8, device="cuda"); epsd = dev(eps)
    Kwad, labd = dev(Kwa), dev(labels, torch.int32)
    check(L.clv_keyenc_fwd(ptr(dev(roll, torch.uint8)), ptr(dev(off, torch.int32)), 1, Lq, D, ptr(dev(Khw)),
                           ptr(dev(bhw)), ptr(Kwad), ptr(dev(bwa)), ptr(epsd), ptr(labd), ptr(hWd), ptr(Wad),
                           ptr(Wd), ptr(loss), B_, Cc, 0.1, 1.0 / B_, 0, 0, None, st))
    torch.cuda.synchronize()
    assert util.rel_err(hWd.cpu().numpy(), hW) < TOL
    assert util.rel_err(Wad.cpu().numpy(), Wa) < TOL
    assert util.rel_err(Wd.cpu().numpy(), W) < TOL
    lo = loss.cpu().numpy()
    assert abs(lo[1] - wkl.mean()) < TOL * abs(wkl.mean()) and abs(lo[2] - wrec.mean()) < TOL * abs(wrec.mean())
    assert abs(lo[4] - corr.mean()) < 1e-6
    dW_ext = rng.normal(size=(B_, Cc))
    dWm, dWlv = M.logitnormal_bwd(Wa[:, :C1], Wa[:, C1:], eps, wt, W, dW_ext, Cc, 0.1, 0.7 / B_, 0.9 / B_)
    dWa_ref = np.concatenate([dWm, dWlv], -1)
    dhW_ref = (dWa_ref @ Kwa.T) * (hW > 0)
    dWa = torch.zeros(B_, 2 * C1, device="cuda"); dhW = torch.zeros(B_, D, device="cuda")
    check(L.clv_keyenc_bwd(ptr(Wad), ptr(epsd), ptr(labd), ptr(Wd), ptr(dev(dW_ext)), ptr(Kwad), ptr(hWd),
                           ptr(dWa), ptr(dhW), B_, Cc, D, 0.1, 0.7 / B_, 0.9 / B_, st))
    torch.cuda.synchronize()
    assert util.rel_err(dWa.cpu().numpy(), dWa_ref) < TOL
    assert util.rel_err(dhW.cpu().numpy(), dhW_ref) < TOL


@pytest.mark.parametrize("B_,L,Z,Cc,has_x", [(200, 16, 2, 10, 1), (5, 7, 4, 3, 0), (19, 3, 16, 16, 1)])
def test_lstm_fused_input_terms_and_extra_gradients(B_, L, Z, Cc, has_x):
    _lib, Lb, check, ptr, st = _env()
    rng = np.random.default_rng(B_ * 10 + L)
    H, G = 88, 352
    xp = rng.normal(0, 1.0, size=(B_, L, G)) * has_x
    U = rng.normal(0, 0.15, size=(H, G)); bias = rng.normal(0, 0.2, G)
    Wv = rng.dirichlet(np.ones(Cc), B_); Ww = rng.normal(0, 0.4, (Cc, G))
    Zs = rng.normal(size=(B_, L, Z)); Kz = rng.normal(0, 0.4, (Z, G))
    xproj = xp + bias + (Wv @ Ww)[:, None, :] + Zs @ Kz
    hs, cs, a_all = M.lstm_fwd(xproj, U)
    gates = dev(xp); Ud = dev(U); Wwd, Kzd = dev(Ww), dev(Kz)
    hd = torch.zeros(B_, L, H, device="cuda"); cd = torch.zeros(B_, L, H, device="cuda")
    check(Lb.clv_lstm_fwd_fused(ptr(gates), has_x, ptr(Ud), ptr(dev(bias)), ptr(dev(Wv)), ptr(Wwd), Cc,
                                ptr(dev(Zs)), ptr(Kzd), Z, ptr(hd), ptr(cd), B_, L, H, st))
    torch.cuda.synchronize()
    assert util.rel_err(hd.cpu().numpy(), hs) < TOL and util.rel_err(cd.cpu().numpy(), cs) < TOL
    dh_out = rng.normal(size=(B_, L, H))
    dA = M.lstm_bwd(dh_out, hs, cs, a_all, U)
    dAsum = torch.zeros(B_, G, device="cuda"); dZ = torch.zeros(B_, L, Z, device="cuda")
    base = rng.normal(size=(B_, Cc)); dWe = dev(base)
    check(Lb.clv_lstm_bwd_fused(ptr(gates), ptr(Ud), ptr(cd), ptr(dev(dh_out)), ptr(dAsum), ptr(Wwd), Cc,
                                ptr(dWe), 1, ptr(Kzd), Z, ptr(dZ), B_, L, H, st))
    torch.cuda.synchronize()
    assert util.rel_err(gates.cpu().numpy(), dA) < TOL
    assert util.rel_err(dZ.cpu().numpy(), dA @ Kz.T) < TOL
    assert util.rel_err(dWe.cpu().numpy(), base + dA.sum(1) @ Ww.T) < TOL


@pytest.mark.parametrize("B_,L,Z", [(200, 16, 2), (5, 7, 1), (37, 4, 2)])
def test_lstm_bwd_with_fused_z_head_exchange(B_, L, Z):
    """clv_lstm_bwd_heads: the decoder BPTT emits dLoss/d(Z_mean|Z_log_var) (reparametrisation backward +
    kl term), the encoder BPTT consumes it as dLoss/dh_e = dZargs @ [Kzm|Kzv]^T per cell -- against the
    numpy derivation (oracle/manual_bwd.py) of the same chain (cl_vrnn/model.py:200-216,236-239)."""
    _lib, Lb, check, ptr, st = _env()
    rng = np.random.default_rng(B_ * 7 + L + Z)
    H, G = 88, 352
    U_d = rng.normal(0, 0.15, (H, G)); U_e = rng.normal(0, 0.15, (H, G))
    Kz = rng.normal(0, 0.4, (Z, G)); Kzm = rng.normal(0, 0.3, (H, Z)); Kzv = rng.normal(0, 0.3, (H, Z))
    Zargs = rng.normal(0, 0.5, (B_, L, 2 * Z)); eps = rng.normal(size=(B_, L, Z))
    klw = 0.3 / (B_ * L)
    # two independent forward recurrences give consistent stashes for the two BPTTs
    xp_d = rng.normal(size=(B_, L, G)); xp_e = rng.normal(size=(B_, L, G))
    hs_d, cs_d, a_d = M.lstm_fwd(xp_d, U_d)
    hs_e, cs_e, a_e = M.lstm_fwd(xp_e, U_e)
    dh_d = rng.normal(size=(B_, L, H))
    dA_d = M.lstm_bwd(dh_d, hs_d, cs_d, a_d, U_d)
    dZ = dA_d @ Kz.T
    mu, lv = Zargs[..., :Z], Zargs[..., Z:]
    dmu = dZ + klw * mu
    dlv = dZ * eps * 0.5 * np.exp(0.5 * lv) + klw * 0.5 * (np.exp(lv) - 1.0)
    dh_e = dmu @ Kzm.T + dlv @ Kzv.T
    dA_e = M.lstm_bwd(dh_e, hs_e, cs_e, a_e, U_e)

    def run_fwd(xp, U):
        g = dev(xp); hd = torch.zeros(B_, L, H, device="cuda"); cd = torch.zeros(B_, L, H, device="cuda")
        check(Lb.clv_lstm_fwd(ptr(g), ptr(dev(U)), ptr(hd), ptr(cd), None, None, B_, L, H, st))
        return g, cd
    g_d, c_d = run_fwd(xp_d, U_d)
    g_e, c_e = run_fwd(xp_e, U_e)
    dAs_d = torch.zeros(B_, G, device="cuda"); dAs_e = torch.zeros(B_, G, device="cuda")
    dZd = torch.zeros(B_, L, Z, device="cuda"); dZa = torch.zeros(B_, L, 2 * Z, device="cuda")
    check(Lb.clv_lstm_bwd_heads(ptr(g_d), ptr(dev(U_d)), ptr(c_d), ptr(dev(dh_d)), ptr(dAs_d), None, 0, None, 0,
                                ptr(dev(Kz)), Z, ptr(dZd), ptr(dev(Zargs)), ptr(dev(eps)), klw, ptr(dZa),
                                None, None, None, 0, B_, L, H, st))
    check(Lb.clv_lstm_bwd_heads(ptr(g_e), ptr(dev(U_e)), ptr(c_e), None, ptr(dAs_e), None, 0, None, 0,
                                None, 0, None, None, None, 0.0, None,
                                ptr(dZa), ptr(dev(Kzm)), ptr(dev(Kzv)), Z, B_, L, H, st))
    torch.cuda.synchronize()
    assert util.rel_err(dZd.cpu().numpy(), dZ) < TOL
    assert util.rel_err(dZa.cpu().numpy(), np.concatenate([dmu, dlv], -1)) < TOL
    assert util.rel_err(g_d.cpu().numpy(), dA_d) < TOL
    assert util.rel_err(g_e.cpu().numpy(), dA_e) < TOL
    assert util.rel_err(dAs_e.cpu().numpy(), dA_e.sum(1)) < TOL


# ---------------------------------------------------------------------------- tcgen05 path
@pytest.mark.parametrize("B_,Lq,shift", [(200, 16, 1), (3, 5, 0), (1000, 33, 1)])
def test_inproj_tcgen05_is_fp32_exact(B_, Lq, shift):
    """tcgen05/TMEM input projection vs float64 numpy: the roll is {0,1} and the weights are split into
    bf16 hi+mid+lo, so the result must agree with fp32 to rounding (not to a bf16 tolerance)."""
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(B_ + Lq)
    D, N = 88, 352
    roll = (rng.random((B_ * (Lq + 1) + 8, D)) < 0.15).astype(np.uint8)
    off = (np.arange(B_) * (Lq + 1)).astype(np.int32)
    Wt = rng.normal(0, 0.3, (D, N)) * np.exp(rng.normal(0, 2, (D, N)))      # wide dynamic range
    ra = rng.normal(size=(B_, N))
    X = np.stack([roll[o + shift:o + shift + Lq] for o in off]).reshape(-1, D).astype(np.float64)
    ref = X @ Wt.astype(np.float32).astype(np.float64) + np.repeat(ra.astype(np.float32).astype(np.float64), Lq, axis=0)
    M_ = B_ * Lq
    Cd = torch.full((M_, N), float("nan"), device="cuda")
    scratch = torch.zeros(L.clv_inproj_tc_scratch_bytes() // 4, device="cuda")
    check(L.clv_inproj_tc(ptr(dev(roll, torch.uint8)), ptr(dev(off, torch.int32)), Lq, shift, D, ptr(dev(Wt)), N, N,
                          ptr(scratch), ptr(Cd), N, M_, ptr(dev(ra)), N, Lq, st))
    torch.cuda.synchronize()
    got = Cd.cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


@pytest.mark.parametrize("B_,Lq,Z,with_x", [(200, 16, 2, True), (7, 5, 8, True), (64, 33, 0, True), (300, 8, 2, False)])
def test_lstm_wgrad_tcgen05(B_, Lq, Z, with_x):
    """tcgen05 weight gradients (MN-major operands, hi+mid splits) vs float64: 1e-4 of the tensor's max."""
    _lib, L, check, ptr, st = _env()
    rng = np.random.default_rng(B_ * 3 + Lq)
    D, H, G = 88, 88, 352
    R = B_ * Lq
    roll = (rng.random((B_ * (Lq + 1) + 4, D)) < 0.1).astype(np.uint8)
    off = (np.arange(B_) * (Lq + 1)).astype(np.int32)
    X = np.stack([roll[o + 1:o + 1 + Lq] for o in off]).reshape(R, D).astype(np.float64)
    dA = rng.normal(0, 1, (R, G)) * np.exp(rng.normal(0, 1.5, (R, 1)))
    h = np.tanh(rng.normal(0, 1, (B_, Lq, H)))
    hprev = np.concatenate([np.zeros((B_, 1, H)), h[:, :-1]], axis=1).reshape(R, H)
    Zs = rng.normal(size=(R, max(Z, 1)))
    base = rng.normal(size=(D + H + max(Z, 1), G))
    gKx, gU, gKz = dev(base[:D]), dev(base[D:D + H]), dev(base[D + H:])
    check(L.clv_lstm_wgrad_tc(ptr(dev(dA)), ptr(dev(roll, torch.uint8)), ptr(dev(off, torch.int32)), Lq, 1, D,
                              ptr(dev(h)), ptr(dev(Zs)) if Z else None, Z, ptr(gKx) if with_x else None, ptr(gU),
                              ptr(gKz) if Z else None, R, H, st))
    torch.cuda.synchronize()
    dA32 = dA.astype(np.float32).astype(np.float64)
    ref_x = base[:D] + (X.T @ dA32 if with_x else 0)
    ref_u = base[D:D + H] + hprev.astype(np.float32).astype(np.float64).T @ dA32
    assert util.rel_err(gKx.cpu().numpy(), ref_x) < TOL
    assert util.rel_err(gU.cpu().numpy(), ref_u) < TOL
    if Z:
        ref_z = base[D + H:] + Zs.astype(np.float32).astype(np.float64).T @ dA32
        assert util.rel_err(gKz.cpu().numpy(), ref_z) < TOL


@pytest.mark.parametrize("B_,L,Z", [(128, 6, 0), (300, 16, 2), (1500, 5, 4)])
def test_lstm_fwd_tcgen05_matches_oracle(B_, L, Z):
    """tcgen05 recurrence (fp16 hi/lo splits, 3 products) vs the float64 cell: same 1e-4 bound as the
    FFMA kernel -- and in practice ~1e-6."""
    _lib, Lb, check, ptr, st = _env()
    rng = np.random.default_rng(B_ + L)
    H, G = 88, 352
    xproj = rng.normal(0, 1.2, size=(B_, L, G)); U = rng.normal(0, 0.15, size=(H, G))
    Zs = rng.normal(size=(B_, L, max(Z, 1))); Kz = rng.normal(0, 0.4, (max(Z, 1), G))
    full = xproj + (Zs @ Kz if Z else 0)
    hs, cs, a_all = M.lstm_fwd(full, U)
    gates = dev(xproj)
    hd = torch.zeros(B_, L, H, device="cuda"); cd = torch.zeros(B_, L, H, device="cuda")
    scratch = torch.zeros(Lb.clv_lstm_fwd_tc_scratch_bytes() // 4, device="cuda")
    check(Lb.clv_lstm_fwd_tc(ptr(gates), ptr(dev(U)), ptr(dev(Zs)) if Z else None, ptr(dev(Kz)) if Z else None, Z,
                             ptr(hd), ptr(cd), ptr(scratch), B_, L, H, st))
    torch.cuda.synchronize()
    err_h = util.rel_err(hd.cpu().numpy(), hs)
    assert err_h < 2e-5, err_h
    assert util.rel_err(cd.cpu().numpy(), cs) < 2e-5
    g_ref = np.concatenate([M._hard_sigmoid(a_all[..., :2 * H]), np.tanh(a_all[..., 2 * H:3 * H]),
                            M._hard_sigmoid(a_all[..., 3 * H:])], -1)
    assert util.rel_err(gates.cpu().numpy(), g_ref) < 2e-5
