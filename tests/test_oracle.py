"""CPU tests of the oracle itself: hand-derived backward vs autograd, golden vectors produced by the
reference's own weightnorm.py, and the invariants the domain offers."""
import os
import numpy as np
import pytest
import torch

from oracle import clv_oracle as O, manual_bwd as M
import util

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("use_x_prev", [True, False])
def test_vrnn_manual_backward_matches_autograd(use_x_prev):
    case = util.make_vrnn_case(1, B=5, L=4, C=6, Z=3, use_x_prev=use_x_prev)
    kw = dict(class_weight=0.7, kl_weight=0.3, w_kl_weight=0.9, w_log_var_prior=0.2)
    out, g = util.oracle_vrnn(case, **kw)
    win = case["win"].astype(np.float64)
    X, Xp = (win[:, 1:], win[:, :-1]) if use_x_prev else (win, None)
    l2, g2 = M.vrnn_manual({k: v.numpy() for k, v in case["p"].items()}, X, Xp,
                           O.one_hot(case["labels"], case["C"]).numpy(), case["eps_w"],
                           case["eps_z"], case["C"], use_x_prev, **kw)
    assert abs(float(out["loss"]) - l2["loss"]) < 1e-10
    for k in g:
        assert util.rel_err(g2[k], g[k].numpy()) < 1e-9, k


@pytest.mark.parametrize("use_x_prev", [True, False])
def test_vae_manual_backward_matches_autograd(use_x_prev):
    case = util.make_vae_case(2, B=7, C=3, Z=4, use_x_prev=use_x_prev)
    kw = dict(class_weight=1.3, kl_weight=0.5, w_kl_weight=0.8, w_log_var_prior=-0.1)
    out, g = util.oracle_vae(case, **kw)
    win = case["win"].astype(np.float64)
    x, xp = (win[:, 1], win[:, 0]) if use_x_prev else (win[:, 0], None)
    l2, g2 = M.vae_manual({k: v.numpy() for k, v in case["p"].items()}, x, xp,
                          O.one_hot(case["labels"], case["C"]).numpy(), case["eps_w"], case["eps_z"],
                          case["C"], use_x_prev, **kw)
    assert abs(float(out["loss"]) - l2["loss"]) < 1e-10
    for k in g:
        assert util.rel_err(g2[k], g[k].numpy()) < 1e-9, k


def test_adamwn_oracle_matches_reference_weightnorm_golden():
    """tests/golden/adamwn.npz was produced by running the reference's utils/weightnorm.py
    (AdamWithWeightnorm.get_updates) on a numpy shim of keras/tf -- see make_golden.py."""
    G = np.load(os.path.join(HERE, "golden", "adamwn.npz"))
    n = 5
    params = {str(i): torch.tensor(G["p0_%d" % i]) for i in range(n)}
    opt = O.AdamWN(params)
    for s in range(4):
        grads = {str(i): torch.tensor(G["g%d_%d" % (s, i)]) for i in range(n)}
        params = opt.step(params, grads)
        for i in range(n):
            assert np.abs(params[str(i)].numpy() - G["p%d_%d" % (s + 1, i)]).max() < 1e-12
    # helper golden (get_weightnorm_params_and_grads)
    p, g = G["h_p"], G["h_g"]
    V_norm = np.sqrt((p * p).sum(0))
    assert np.allclose(V_norm, G["h_V_norm"], atol=1e-14)
    grad_g = (g * p).sum(0) / V_norm
    assert np.allclose(grad_g, G["h_grad_g"], atol=1e-13)
    assert np.allclose(g - grad_g / V_norm * p, G["h_grad_V"], atol=1e-13)


def test_adamwn_invariants():
    rng = np.random.default_rng(3)
    params = {"k": torch.tensor(rng.normal(size=(9, 5))), "b": torch.tensor(rng.normal(size=(5,)))}
    opt = O.AdamWN(params)
    for _ in range(3):
        grads = {k: torch.tensor(rng.normal(size=tuple(v.shape))) for k, v in params.items()}
        params = opt.step(params, grads)
        st = opt.state["k"]
        V = params["k"] / st["V_scaler"]
        # W == V_scaler * V and ||W||_col == g  (g = V_scaler * ||V||)
        assert torch.allclose(params["k"], st["V_scaler"] * V)
        assert torch.allclose(torch.sqrt((params["k"] ** 2).sum(0)), st["V_scaler"].abs() * torch.sqrt((V * V).sum(0)))


def test_loss_properties():
    rng = np.random.default_rng(4)
    mu = torch.tensor(rng.normal(size=(6, 9))); lv = torch.tensor(rng.normal(size=(6, 9)))
    W = O.logistic_normal(mu, lv, torch.tensor(rng.normal(size=(6, 9))))
    assert torch.allclose(W.sum(-1), torch.ones(6, dtype=W.dtype))
    # last logit pinned to 0: W[:, -1] == 1 / sum(exp(W0))
    assert (O.z_kl_loss(mu, lv) >= 0).all() and (O.w_kl_loss(mu, lv, 0.0) >= 0).all()
    assert torch.allclose(O.z_kl_loss(mu, lv), O.w_kl_loss(mu, lv, 0.0))
    # BCE: clip region has zero gradient
    logit = torch.tensor([40.0, -40.0, 0.3], requires_grad=True)
    O.keras_bce(torch.tensor([0.0, 1.0, 1.0]), torch.sigmoid(logit)).sum().backward()
    assert logit.grad[0] == 0 and logit.grad[1] == 0 and logit.grad[2] != 0


def test_sampler_oracle_teacher_forcing_and_shapes():
    rng = np.random.default_rng(5)
    L, D, H, Z, C = 4, 88, 88, 2, 3
    p = O.init_vrnn_params(rng, L, D, H, Z, C, True)
    seed = torch.tensor(O.synth_rolls(rng, 1, 6, D, 0.1)[0], dtype=torch.float64)
    w = O.one_hot([1], C)
    T = 6 + 5
    Xs, Ps = O.vrnn_generate_sample(p, seed, 5, w, torch.tensor(rng.standard_normal((T, Z))),
                                    torch.tensor(rng.random((T, D))), True)
    assert Xs.shape == (5, D) and Ps.shape == (T, D)
    assert set(np.unique(Xs.numpy())) <= {0.0, 1.0}
    # quirk Q2: chunk starts come from range(0, D, seq_length); only full chunks count
    w_inf = O.vrnn_infer_w(p, seed, L, C)
    assert w_inf.shape == (1, C) and abs(float(w_inf.sum()) - 1) < 1e-12
