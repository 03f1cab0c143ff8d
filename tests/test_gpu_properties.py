"""Size-independent properties at sweep-scale sizes (the oracle finishes only small cases in seconds):
permutation invariance, micro-batch linearity, eval idempotence, sampler threshold limits."""
import ctypes as C
import numpy as np
import pytest
import torch

import util
from oracle import clv_oracle as O

pytestmark = pytest.mark.gpu


def _engine(B, L, Cc=12, Z=2, seed=0, **kw):
    from clvae_b200.engine import Engine
    e = Engine("vrnn", B, L=L, D=88, H=88, Z=Z, n_classes=Cc, use_x_prev=True, use_graph=False, seed=seed, **kw)
    e.init_params(np.random.default_rng(seed))
    return e


def _stage(e, win, labels, eps_w, eps_z):
    e.stage_windows(torch.tensor(win).cuda(), torch.tensor(labels).cuda())
    e.eps_w.copy_(torch.tensor(eps_w, dtype=torch.float32).reshape(-1))
    e.eps_z.copy_(torch.tensor(eps_z, dtype=torch.float32).reshape(-1))


def test_batch_permutation_invariance_at_sweep_size():
    B, L, Cc, Z = 2048, 64, 12, 2
    rng = np.random.default_rng(0)
    win = O.synth_rolls(rng, B, L + 1); labels = rng.integers(0, Cc, B).astype(np.int32)
    eps_w = rng.standard_normal((B, Cc - 1)).astype(np.float32); eps_z = rng.standard_normal((B, L, Z)).astype(np.float32)
    e = _engine(B, L, Cc, Z)
    _stage(e, win, labels, eps_w, eps_z)
    e.run(train=True, gen_noise=False)
    l0 = e.read_losses(); g0 = e.grads.clone()
    perm = rng.permutation(B)
    e2 = _engine(B, L, Cc, Z)
    _stage(e2, win[perm], labels[perm], eps_w[perm], eps_z[perm])
    e2.run(train=True, gen_noise=False)
    l1 = e2.read_losses()
    for k in ("loss", "vae", "w_kl", "w_rec", "z_kl", "acc"):
        assert abs(l0[k] - l1[k]) <= 2e-5 * max(1.0, abs(l0[k])), k
    assert util.rel_err(e2.grads.cpu().numpy(), g0.cpu().numpy()) < 1e-4
    assert np.isfinite(g0.cpu().numpy()).all()


def test_microbatch_linearity_and_eval_idempotence_at_scale():
    from clvae_b200._lib import lib, check, ptr
    B, L, Cc, Z, parts = 1024, 32, 12, 2, 4
    rng = np.random.default_rng(1)
    win = O.synth_rolls(rng, B, L + 1); labels = rng.integers(0, Cc, B).astype(np.int32)
    eps_w = rng.standard_normal((B, Cc - 1)).astype(np.float32); eps_z = rng.standard_normal((B, L, Z)).astype(np.float32)
    full = _engine(B, L, Cc, Z)
    _stage(full, win, labels, eps_w, eps_z)
    full.run(train=False, gen_noise=False); a = full.read_losses()
    full.run(train=False, gen_noise=False); b = full.read_losses()
    # forward is idempotent up to the order of the fp32 atomics that accumulate the loss scalars
    assert all(abs(a[k] - b[k]) <= 2e-5 * max(1.0, abs(a[k])) for k in a)
    full.run(train=True, gen_noise=False)
    part = _engine(B // parts, L, Cc, Z)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    acc = torch.zeros_like(full.gradbuf)
    for i in range(parts):
        sl = slice(i * B // parts, (i + 1) * B // parts)
        _stage(part, win[sl], labels[sl], eps_w[sl], eps_z[sl])
        cfg = part.cfg(B_global=B, gen_noise=0, do_backward=1, accumulate=int(i > 0))
        check(lib().clv_train_step(C.byref(cfg), ptr(part.params), ptr(acc[:part.P]), ptr(acc[part.P:]),
                                   ptr(part.roll), ptr(part.win_off), ptr(part.labels), ptr(part.eps_w),
                                   ptr(part.eps_z), None, ptr(part.workspace), part.workspace.numel() * 4, st))
    torch.cuda.synchronize()
    assert util.rel_err(acc[:part.P].cpu().numpy(), full.grads.cpu().numpy()) < 1e-4
    assert util.rel_err(acc[part.P:part.P + 5].cpu().numpy(), full.loss_acc[:5].cpu().numpy()) < 1e-5


def test_sampler_threshold_limits_and_teacher_forcing_at_scale():
    from clvae_b200._lib import lib, check, ptr
    S, Ts, N, Cc, Z, D = 3000, 8, 40, 12, 2, 88
    e = _engine(1, 16, Cc, Z, seed=3)
    rng = np.random.default_rng(2)
    T = Ts + N
    seeds = torch.tensor(O.synth_rolls(rng, S, Ts)).cuda()
    w = torch.zeros(S, Cc, device="cuda"); w[torch.arange(S), torch.randint(0, Cc, (S,))] = 1.0
    eps = torch.randn(S, T, Z, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfg = e.cfg()
    outs = []
    for uval in (0.0, 2.0):   # u <= p always / never
        u = torch.full((S, T, D), uval, device="cuda")
        out = torch.zeros(S, T, D, dtype=torch.uint8, device="cuda")
        check(lib().clv_vrnn_sample(C.byref(cfg), ptr(e.params), None, None, None, ptr(seeds), Ts, N, ptr(w), ptr(eps),
                                    ptr(u), 0, 0, S, ptr(out), None, st))
        torch.cuda.synchronize()
        outs.append(out)
    assert int(outs[0].min()) == 1 and int(outs[1].max()) == 0
    # same tapes, different launch split (S is not a multiple of the 16/24-song tile): identical output
    u = torch.rand(S, T, D, device="cuda")
    full = torch.zeros(S, T, D, dtype=torch.uint8, device="cuda")
    check(lib().clv_vrnn_sample(C.byref(cfg), ptr(e.params), None, None, None, ptr(seeds), Ts, N, ptr(w), ptr(eps), ptr(u),
                                0, 0, S, ptr(full), None, st))
    cut = 1237
    tail = torch.zeros(S - cut, T, D, dtype=torch.uint8, device="cuda")
    check(lib().clv_vrnn_sample(C.byref(cfg), ptr(e.params), None, None, None, ptr(seeds[cut:].contiguous()), Ts, N,
                                ptr(w[cut:].contiguous()), ptr(eps[cut:].contiguous()), ptr(u[cut:].contiguous()), 0, 0,
                                S - cut, ptr(tail), None, st))
    torch.cuda.synchronize()
    assert torch.equal(full[cut:], tail)


def test_empty_batch_and_zero_steps_are_noops():
    from clvae_b200._lib import lib, check, ptr
    e = _engine(4, 3, 4, 2)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfg = e.cfg(B=0, B_global=4)
    before = e.gradbuf.clone()
    check(lib().clv_train_step(C.byref(cfg), ptr(e.params), ptr(e.grads), ptr(e.loss_acc), ptr(e.roll), ptr(e.win_off),
                               ptr(e.labels), ptr(e.eps_w), ptr(e.eps_z), None, ptr(e.workspace),
                               e.workspace.numel() * 4, st))
    out = torch.zeros(1, dtype=torch.uint8, device="cuda")
    check(lib().clv_vrnn_sample(C.byref(e.cfg()), ptr(e.params), None, None, None, ptr(out), 1, 5, ptr(e.eps_w), None,
                                None, 0, 0, 0, ptr(out), None, st))
    torch.cuda.synchronize()
    assert torch.equal(before, e.gradbuf)
