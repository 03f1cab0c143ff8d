"""Persistent samplers vs the oracle's restatement of generate_sample on recorded noise tapes:
bit-exact piano rolls wherever |p - u| > 1e-6 (north_star)."""
import ctypes as CT
import numpy as np
import pytest
import torch

import util
from oracle import clv_oracle as O

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("S,T_seed,nsteps,C,Z,xp", [(3, 4, 12, 10, 2, True), (17, 1, 6, 3, 4, False),
                                                    (1, 16, 20, 12, 2, True)])
def test_vrnn_sampler_bit_exact_away_from_threshold(S, T_seed, nsteps, C, Z, xp):
    from clvae_b200 import _lib
    from clvae_b200._lib import lib, check, ptr
    from clvae_b200.engine import Engine
    rng = np.random.default_rng(S * 31 + nsteps)
    L, D, H = 4, 88, 88
    p = O.init_vrnn_params(rng, L, D, H, Z, C, xp)
    p["X_decoded_mean.bias"] = p["X_decoded_mean.bias"] - 1.5
    e = Engine("vrnn", 1, L=L, D=D, H=H, Z=Z, n_classes=C, use_x_prev=xp, use_graph=False)
    e.set_params({k: v.numpy() for k, v in p.items()})
    T = T_seed + nsteps
    seeds = O.synth_rolls(rng, S, T_seed, D, 0.1)
    w = rng.dirichlet(np.ones(C), S)
    eps_z = rng.standard_normal((S, T, Z)).astype(np.float32)
    u = rng.random((S, T, D)).astype(np.float32)
    out = torch.zeros(S, T, D, dtype=torch.uint8, device="cuda")
    probs = torch.zeros(S, T, D, device="cuda")
    cfg = e.cfg()
    st = CT.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib().clv_vrnn_sample(CT.byref(cfg), ptr(e.params), None, None, None,
                                ptr(dev(seeds, torch.uint8)), T_seed, nsteps, ptr(dev(w)),
                                ptr(dev(eps_z)), ptr(dev(u)), 0, 0, S, ptr(out), ptr(probs), st))
    torch.cuda.synchronize()
    out, probs = out.cpu().numpy(), probs.cpu().numpy()
    n_checked = 0
    for s in range(S):
        # replay the oracle on the GPU's own history so one borderline flip cannot cascade
        Xs, Ps = O.vrnn_generate_sample(p, torch.tensor(seeds[s], dtype=torch.float64), nsteps,
                                        torch.tensor(w[s:s + 1]), torch.tensor(eps_z[s], dtype=torch.float64),
                                        torch.tensor(u[s], dtype=torch.float64), xp)
        Ps, Xs = Ps.numpy(), Xs.numpy()
        full = np.concatenate([np.zeros((T_seed, D)), Xs])
        for t in range(T):
            far = np.abs(Ps[t] - u[s, t]) > 1e-6
            ref_x = (u[s, t] <= Ps[t]).astype(np.uint8)
            assert np.abs(probs[s, t] - Ps[t]).max() < 2e-5, (s, t)
            assert np.array_equal(out[s, t][far], ref_x[far]), (s, t)
            n_checked += int(far.sum())
            if t >= T_seed and not np.array_equal(out[s, t], full[t].astype(np.uint8)):
                break  # a within-1e-6 flip changed the history; later steps are not comparable
    assert n_checked > 0.9 * S * T_seed * D


@pytest.mark.parametrize("S,nsteps,C,Z,xp,prior", [(5, 10, 2, 4, True, False), (16, 7, 10, 2, False, True)])
def test_vae_sampler_bit_exact_away_from_threshold(S, nsteps, C, Z, xp, prior):
    from clvae_b200._lib import lib, check, ptr
    from clvae_b200.engine import Engine
    rng = np.random.default_rng(S + nsteps)
    D, H, Hc = 88, 88, 88
    p = O.init_vae_params(rng, D, H, Z, Hc, C, xp)
    p["x_decoded_mean.bias"] = p["x_decoded_mean.bias"] - 1.0
    e = Engine("vae", 1, D=D, H=H, Z=Z, n_classes=C, use_x_prev=xp, Hc=Hc, use_graph=False)
    e.set_params({k: v.numpy() for k, v in p.items()})
    seeds = O.synth_rolls(rng, S, 1, D, 0.1)[:, 0]
    w = rng.dirichlet(np.ones(C), S)
    eps_z = rng.standard_normal((S, nsteps, Z)).astype(np.float32)
    u = rng.random((S, nsteps, D)).astype(np.float32)
    out = torch.zeros(S, nsteps, D, dtype=torch.uint8, device="cuda")
    probs = torch.zeros(S, nsteps, D, device="cuda")
    cfg = e.cfg()
    st = CT.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib().clv_vae_sample(CT.byref(cfg), ptr(e.params), ptr(dev(seeds, torch.uint8)), nsteps,
                               ptr(dev(w)), ptr(dev(eps_z)), ptr(dev(u)), 0, 0, S, int(prior),
                               ptr(out), ptr(probs), st))
    torch.cuda.synchronize()
    out, probs = out.cpu().numpy(), probs.cpu().numpy()
    for s in range(S):
        Xs, Ps = O.vae_generate_sample(p, torch.tensor(seeds[s], dtype=torch.float64), nsteps,
                                       torch.tensor(w[s:s + 1]), torch.tensor(eps_z[s], dtype=torch.float64),
                                       torch.tensor(u[s], dtype=torch.float64), xp, use_z_prior=prior)
        Xs, Ps = Xs.numpy(), Ps.numpy()
        for t in range(nsteps):
            far = np.abs(Ps[t] - u[s, t]) > 1e-6
            assert np.abs(probs[s, t] - Ps[t]).max() < 2e-5, (s, t)
            assert np.array_equal(out[s, t][far], (u[s, t] <= Ps[t]).astype(np.uint8)[far]), (s, t)
            if not np.array_equal(out[s, t], Xs[t].astype(np.uint8)):
                break


def test_sampler_philox_mode_is_deterministic_and_song_indexed():
    """Throughput mode: noise keyed by (seed, global song index, t) => output independent of how
    songs are split across launches / ranks."""
    from clvae_b200._lib import lib, check, ptr
    from clvae_b200.engine import Engine
    rng = np.random.default_rng(0)
    L, D, H, Z, C = 4, 88, 88, 2, 5
    e = Engine("vrnn", 1, L=L, D=D, H=H, Z=Z, n_classes=C, use_x_prev=True, use_graph=False)
    e.init_params(rng)
    S, T_seed, nsteps = 40, 2, 9
    T = T_seed + nsteps
    seeds = dev(O.synth_rolls(rng, S, T_seed, D, 0.1), torch.uint8)
    w = dev(rng.dirichlet(np.ones(C), S))
    st = CT.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfg = e.cfg()
    full = torch.zeros(S, T, D, dtype=torch.uint8, device="cuda")
    check(lib().clv_vrnn_sample(CT.byref(cfg), ptr(e.params), None, None, None, ptr(seeds), T_seed, nsteps,
                                ptr(w), None, None, 1234, 0, S, ptr(full), None, st))
    part = torch.zeros(S - 24, T, D, dtype=torch.uint8, device="cuda")
    check(lib().clv_vrnn_sample(CT.byref(cfg), ptr(e.params), None, None, None, ptr(seeds[24:].contiguous()),
                                T_seed, nsteps, ptr(w[24:].contiguous()), None, None, 1234, 24, S - 24,
                                ptr(part), None, st))
    torch.cuda.synchronize()
    assert torch.equal(full[24:], part)
    assert 0.0 < full.float().mean().item() < 1.0


def test_bit_packed_sampler_output_equals_the_uint8_rolls():
    """clv_vrnn_sample_bits: 11 bytes per 88-key frame (numpy.unpackbits little-endian) == clv_vrnn_sample."""
    from clvae_b200._lib import lib, check, ptr
    from clvae_b200.engine import Engine
    rng = np.random.default_rng(77)
    S, T_seed, nsteps, C, Z, L, D, H = 37, 3, 14, 10, 2, 4, 88, 88
    p = O.init_vrnn_params(rng, L, D, H, Z, C, True)
    e = Engine("vrnn", 1, L=L, D=D, H=H, Z=Z, n_classes=C, use_x_prev=True, use_graph=False)
    e.set_params({k: v.numpy() for k, v in p.items()})
    T = T_seed + nsteps
    seeds = dev(O.synth_rolls(rng, S, T_seed, D, 0.1), torch.uint8)
    w = dev(rng.dirichlet(np.ones(C), S))
    out = torch.zeros(S, T, D, dtype=torch.uint8, device="cuda")
    bits = torch.zeros(S, T, 11, dtype=torch.uint8, device="cuda")
    cfg = e.cfg()
    st = CT.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib().clv_vrnn_sample(CT.byref(cfg), ptr(e.params), None, None, None, ptr(seeds), T_seed, nsteps, ptr(w),
                                None, None, 5, 100, S, ptr(out), None, st))
    check(lib().clv_vrnn_sample_bits(CT.byref(cfg), ptr(e.params), None, None, None, ptr(seeds), T_seed, nsteps, ptr(w),
                                     None, None, 5, 100, S, ptr(bits), None, st))
    torch.cuda.synchronize()
    un = np.unpackbits(bits.cpu().numpy(), axis=-1, bitorder="little")[..., :D]
    assert np.array_equal(un, out.cpu().numpy())
    assert out.sum() > 0
