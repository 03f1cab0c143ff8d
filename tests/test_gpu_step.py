"""Whole training step (fwd + 4 losses + accuracy + bwd + Adam-WN) through clv_train_step against the
oracle on identical weights, inputs and injected noise.  1e-4 relative (fp32)."""
import numpy as np
import pytest
import torch

import util
from oracle import clv_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
KW = dict(class_weight=0.7, kl_weight=0.3, w_kl_weight=0.9, w_log_var_prior=0.2)


def check_step(e, out, g):
    e.run(train=True, gen_noise=False)
    lo = e.read_losses()
    for k in ("vae", "w_kl", "w_rec", "z_kl", "loss"):
        assert abs(lo[k] - float(out[k])) <= TOL * abs(float(out[k])), (k, lo[k], float(out[k]))
    assert abs(lo["acc"] - float(out["acc"])) < 1e-6
    for k in e.names:
        got = e.grad_view(k).cpu().numpy()
        assert util.rel_err(got, g[k].numpy().reshape(got.shape)) < TOL, k


@pytest.mark.parametrize("B,L,C,Z,xp", [(200, 16, 10, 2, True), (200, 16, 10, 2, False),
                                        (7, 3, 2, 4, True), (1, 1, 3, 1, True), (33, 40, 12, 2, True)])
def test_vrnn_step_matches_oracle(B, L, C, Z, xp):
    case = util.make_vrnn_case(B * 7 + L, B, L, C=C, Z=Z, use_x_prev=xp)
    out, g = util.oracle_vrnn(case, **KW)
    e = util.engine_for(case, "vrnn", use_graph=False, **KW)
    check_step(e, out, g)
    # intermediates the sampler sub-models expose
    assert util.rel_err(e.ws_view("W", (B, C)).cpu().numpy(), out["W"].numpy()) < TOL
    assert util.rel_err(e.ws_view("h_e", (B, L, 88)).cpu().numpy(), out["h_e"].numpy()) < TOL
    assert util.rel_err(e.ws_view("h_d", (B, L, 88)).cpu().numpy(), out["h_d"].numpy()) < TOL


@pytest.mark.parametrize("B,L,C,Z,xp", [(200, 16, 10, 2, True), (37, 5, 4, 1, False), (3, 9, 12, 2, True)])
def test_vrnn_step_with_the_backward_wavefront_matches_oracle(B, L, C, Z, xp):
    """clv_cfg.pair_bwd: decoder BPTT + Z-head exchange + encoder BPTT as one wavefront launch
    (clv_lstm_pair_bwd), partial last row group and Z = 1 included."""
    case = util.make_vrnn_case(B * 5 + L, B, L, C=C, Z=Z, use_x_prev=xp)
    out, g = util.oracle_vrnn(case, **KW)
    e = util.engine_for(case, "vrnn", use_graph=False, pair_bwd=True, **KW)
    check_step(e, out, g)
    # and twice in a row through a CUDA graph (the sentinel fills are part of the captured step)
    e2 = util.engine_for(case, "vrnn", use_graph=True, pair_bwd=True, **KW)
    e2.run(train=True, gen_noise=False)
    torch.cuda.synchronize()
    # the weight gradients are float atomics (order varies run to run) and the first Adam step is lr * g / |g|
    assert util.rel_err(e2.params.cpu().numpy(), e.params.cpu().numpy()) < 1e-5


def test_vrnn_step_with_the_tensor_core_x_head_matches_oracle():
    """B * L = 38 400 rows >= 2 x 128 x 148: the step takes the tcgen05 X head (clv_xhead_tc)."""
    B, L = 2400, 16
    case = util.make_vrnn_case(B + L + 5, B, L, C=10, Z=2, use_x_prev=True)
    out, g = util.oracle_vrnn(case, **KW)
    e = util.engine_for(case, "vrnn", use_graph=False, **KW)
    check_step(e, out, g)
    assert util.rel_err(e.ws_view("h_d", (B, L, 88)).cpu().numpy(), out["h_d"].numpy()) < TOL


@pytest.mark.parametrize("B,C,Z,xp", [(100, 2, 4, True), (100, 10, 2, False), (5, 3, 16, True)])
def test_vae_step_matches_oracle(B, C, Z, xp):
    case = util.make_vae_case(B + C, B, C=C, Z=Z, use_x_prev=xp)
    out, g = util.oracle_vae(case, **KW)
    e = util.engine_for(case, "vae", use_graph=False, **KW)
    check_step(e, out, g)


def test_vrnn_training_trajectory_and_graph_replay():
    """5 optimizer steps (Adam-WN) on a fixed batch: eager launches, CUDA-graph replay and the oracle
    stay together; validation pass leaves parameters untouched."""
    case = util.make_vrnn_case(99, 24, 6, C=4, Z=2, use_x_prev=True)
    engines = [util.engine_for(case, "vrnn", use_graph=ug) for ug in (False, True)]
    p = {k: v.clone() for k, v in case["p"].items()}
    opt = O.AdamWN(p)
    for step in range(5):
        c2 = dict(case, p=p)
        out, g = util.oracle_vrnn(c2)
        for e in engines:
            e.run(train=True, gen_noise=False)
            lo = e.read_losses()
            assert abs(lo["loss"] - float(out["loss"])) <= 2e-4 * abs(float(out["loss"])), step
        p = opt.step(p, g)
    for e in engines:
        got = e.get_params()
        for k in e.names:
            assert util.rel_err(got[k], p[k].numpy()) < 5e-4, k
        before = e.params.clone()
        e.run(train=False, gen_noise=False)
        assert torch.equal(before, e.params)
    a, b = engines[0].get_params(), engines[1].get_params()
    for k in a:
        assert util.rel_err(a[k], b[k]) < 1e-4   # atomics order differs between eager and graph replays


@pytest.mark.parametrize("model", ["vrnn", "vae"])
def test_step_with_scheduled_optimizer_equals_step_then_adam(model):
    """clv_train_step_opt (Adam-WN launched per tensor range inside the step's schedule) gives the same
    parameters, optimizer state and `iterations` as clv_train_step followed by clv_adamwn_step."""
    if model == "vrnn":
        case = util.make_vrnn_case(123, 48, 9, C=5, Z=2, use_x_prev=True)
    else:
        case = util.make_vae_case(77, 64, C=5, Z=3, use_x_prev=True)
    a = util.engine_for(case, model, use_graph=False, fused_optimizer=True)
    b = util.engine_for(case, model, use_graph=False, fused_optimizer=False)
    for _ in range(4):
        a.run(train=True, gen_noise=False)
        b.run(train=True, gen_noise=False)
    pa, pb = a.get_params(), b.get_params()
    for k in pa:
        assert util.rel_err(pa[k], pb[k]) < 2e-5, k
    n = a.opt_state.numel()
    assert util.rel_err(a.opt_state[:n - 2].cpu().numpy(), b.opt_state[:n - 2].cpu().numpy()) < 2e-5
    ia = a.opt_state[n - 2:].view(torch.int32).cpu().tolist()
    ib = b.opt_state[n - 2:].view(torch.int32).cpu().tolist()
    assert ia == ib == [4, 0]          # iterations advanced once per step, done-counter back at zero


def test_in_kernel_noise_is_standard_normal_and_fresh():
    case = util.make_vrnn_case(5, 256, 8, C=10, Z=2)
    e = util.engine_for(case, "vrnn", use_graph=True)
    e.run(train=False, gen_noise=True)
    z1 = e.eps_z.cpu().numpy().copy(); w1 = e.eps_w.cpu().numpy().copy()
    e.run(train=False, gen_noise=True)
    z2 = e.eps_z.cpu().numpy()
    assert abs(z1.mean()) < 0.1 and abs(z1.std() - 1) < 0.1 and abs(w1.std() - 1) < 0.1
    assert not np.array_equal(z1, z2)
    lo = e.read_losses()
    assert np.isfinite(lo["loss"])


def test_microbatch_accumulation_equals_full_batch():
    """accumulate=1 adds a second micro-batch into grads/losses (how the host fits the top of the
    sweep into HBM); B_global makes the means global."""
    import ctypes as C
    from clvae_b200._lib import lib, check, ptr
    case = util.make_vrnn_case(3, 16, 5, C=4, Z=2)
    out, g = util.oracle_vrnn(case)
    full = util.engine_for(case, "vrnn", use_graph=False)
    full.run(train=True, gen_noise=False)
    half = {}
    for i, sl in enumerate((slice(0, 8), slice(8, 16))):
        c = dict(case, B=8, win=case["win"][sl], labels=case["labels"][sl], eps_w=case["eps_w"][sl],
                 eps_z=case["eps_z"][sl])
        half[i] = util.engine_for(c, "vrnn", use_graph=False)
    e0, e1 = half[0], half[1]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i, e in enumerate((e0, e1)):
        cfg = e.cfg(B_global=16, gen_noise=0, do_backward=1, accumulate=i)
        check(lib().clv_train_step(C.byref(cfg), ptr(e.params), ptr(e0.grads), ptr(e0.loss_acc),
                                   ptr(e.roll), ptr(e.win_off), ptr(e.labels), ptr(e.eps_w),
                                   ptr(e.eps_z), None, ptr(e.workspace), e.workspace.numel() * 4, st))
    torch.cuda.synchronize()
    assert util.rel_err(e0.grads.cpu().numpy(), full.grads.cpu().numpy()) < 1e-4
    assert util.rel_err(e0.loss_acc.cpu().numpy()[:5], full.loss_acc.cpu().numpy()[:5]) < 2e-5


def test_vrnn_step_with_tcgen05_input_projections():
    """gemm_algo=1: the two hoisted input projections run on tcgen05 (bf16x3 split, fp32-exact); the
    step must hold the same 1e-4 bound as the SIMT path."""
    case = util.make_vrnn_case(77, 200, 16, C=10, Z=2, use_x_prev=True)
    out, g = util.oracle_vrnn(case, **KW)
    e = util.engine_for(case, "vrnn", use_graph=False, gemm_algo=1, **KW)
    check_step(e, out, g)


def test_vrnn_step_large_batch_tensor_core_recurrence():
    """B >= 1024 switches the forward recurrence to tcgen05 (plus tcgen05 projections and weight
    gradients): the whole step still meets the 1e-4 bound against the float64 oracle."""
    case = util.make_vrnn_case(1234, 1024, 6, C=10, Z=2, use_x_prev=True)
    out, g = util.oracle_vrnn(case, **KW)
    e = util.engine_for(case, "vrnn", use_graph=False, tc_lstm_min=512, **KW)
    check_step(e, out, g)


def test_fused_p2p_allreduce_adam_matches_nccl_on_two_gpus():
    """Needs >= 2 GPUs (skipped on the 1-GPU box): launches tests/dist_p2p_check.py under torchrun."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_p2p_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script],
                       capture_output=True, text=True, timeout=300)
    assert "P2P_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
