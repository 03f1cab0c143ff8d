"""
Pins the oracle (oracle/clv_oracle.py) to the REFERENCE'S OWN SOURCE: tests/golden/{vrnn,vae}_train.npz
and {vrnn,vae}_sampler.npz were produced by executing /root/reference/code/cl_vrnn/model.py and
cl_vae/model.py (get_model, the loss closures, make_w_encoder / make_z_encoder / make_decoder,
generate_sample, sample_*) against tests/golden/keras_shim.py -- see tests/golden/make_golden.py.
CPU only; the GPU path is compared with the same fixtures in tests/test_gpu_golden.py.
"""
import numpy as np
import pytest
import torch

import util
from oracle import clv_oracle as O

KERAS_KEYS = {"vae": "X_decoded_mean_loss", "w_kl": "W_loss", "w_rec": "W2_loss", "z_kl": "Z_args_loss", "acc": "W_acc"}


def _close(a, b, rtol, atol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    assert not bad.any(), "max abs diff %.3e at %s" % (np.abs(a - b).max(), np.argwhere(bad)[:3].tolist())


def vrnn_case_from_golden(g):
    c = g["cfg"]
    shapes = O.vrnn_param_shapes(c["L"], c["D"], c["H"], c["Z"], c["C"], c["use_x_prev"])
    p = util.golden_params(g, shapes)
    return dict(p={k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}, win=g["win"], labels=g["labels"],
                eps_w=g["eps_w"], eps_z=g["eps_z"], B=c["B"], L=c["L"], C=c["C"], Z=c["Z"], D=c["D"], H=c["H"],
                use_x_prev=c["use_x_prev"], predict_next=c["predict_next"], kw=c["kw"])


def oracle_on_vrnn_golden(case, dtype=torch.float64):
    win = torch.tensor(case["win"], dtype=dtype)
    if case["use_x_prev"]:
        X, Xp, Y = win[:, 1:], win[:, :-1], None
    elif case["predict_next"]:
        X, Xp, Y = win[:, :-1], None, win[:, 1:]
    else:
        X, Xp, Y = win, None, None
    p = {k: v.to(dtype) for k, v in case["p"].items()}
    return O.vrnn_loss_and_grads(p, X, Xp, O.one_hot(case["labels"], case["C"], dtype),
                                 torch.tensor(case["eps_w"], dtype=dtype), torch.tensor(case["eps_z"], dtype=dtype),
                                 case["C"], case["use_x_prev"], Y=Y, **case["kw"])


@pytest.mark.parametrize("name", ["xprev", "noxprev", "predict_next"])
def test_oracle_vrnn_graph_matches_executed_reference_get_model(name):
    g = util.load_golden("vrnn_train.npz")[name]
    # the reference's own layer / output order
    assert g["cfg"]["output_names"] == ["X_decoded_mean", "W", "W2", "Z_args"]
    assert g["cfg"]["weighted_layers"] == ["hW", "Wargs", "encoder_h", "Z_mean", "Z_log_var", "decoder_h", "X_decoded_mean"]
    case = vrnn_case_from_golden(g)
    out, grads = oracle_on_vrnn_golden(case)
    res = g["cfg"]["res"]
    assert abs(float(out["loss"]) - res["loss"]) <= 1e-12 * abs(res["loss"])
    for k, kk in KERAS_KEYS.items():
        assert abs(float(out[k]) - res[kk]) <= 1e-12 * max(abs(res[kk]), 1.0), (k, float(out[k]), res[kk])
    _close(out["W"], g["out/W"], 1e-6, 1e-9)
    _close(torch.cat([out["Z_mean"], out["Z_log_var"]], -1), g["out/Z_args"], 1e-6, 1e-8)
    for k, v in grads.items():          # golden gradients are stored as float32
        _close(v.numpy(), g["g/" + k], 2e-6, 1e-9)


def vae_case_from_golden(g):
    c = g["cfg"]
    shapes = O.vae_param_shapes(c["D"], c["H"], c["Z"], c["Hc"], c["C"], c["use_x_prev"])
    p = util.golden_params(g, shapes)
    return dict(p={k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}, win=g["win"] if c["use_x_prev"] else g["win"][:, 1:],
                labels=g["labels"], eps_w=g["eps_w"], eps_z=g["eps_z"], B=c["B"], C=c["C"], Z=c["Z"], D=c["D"], H=c["H"],
                Hc=c["Hc"], use_x_prev=c["use_x_prev"], kw=c["kw"])


@pytest.mark.parametrize("name", ["xprev", "noxprev"])
def test_oracle_vae_graph_matches_executed_reference_get_model(name):
    g = util.load_golden("vae_train.npz")[name]
    assert g["cfg"]["output_names"] == ["x_decoded_mean", "w", "w2", "z_args"]
    assert g["cfg"]["weighted_layers"] == ["h_w", "w_mean", "w_log_var", "h", "z_mean", "z_log_var", "decoder_h", "x_decoded_mean"]
    case = vae_case_from_golden(g)
    out, grads = util.oracle_vae(case, **case["kw"])
    res = g["cfg"]["res"]
    assert abs(float(out["loss"]) - res["loss"]) <= 1e-12 * abs(res["loss"])
    for k, kk in KERAS_KEYS.items():
        kk = kk.replace("X_decoded_mean", "x_decoded_mean").replace("W2", "w2").replace("W_", "w_").replace("Z_args", "z_args")
        assert abs(float(out[k]) - res[kk]) <= 1e-12 * max(abs(res[kk]), 1.0), (k, float(out[k]), res[kk])
    _close(out["W"], g["out/w"], 1e-6, 1e-9)
    for k, v in grads.items():
        _close(v.numpy(), g["g/" + k], 2e-6, 1e-9)


# ------------------------------------------------------------------------------------------ samplers
def reference_noise_tapes(cfg, n_chunks, T, Z, D, C):
    """np.random draws of generate_sample in the reference's call order (cl_vrnn/model.py:34-59 /
    cl_vae/model.py:24-41; verified at fixture-generation time against the stream position after the
    reference run): per inferred-key chunk randn(1, C-1) (sample_w draws even with add_noise=False);
    one np.random.choice if w_discrete; then per step randn(Z) (sample_z) and rand(D) (sample_x)."""
    np.random.seed(cfg["np_seed"])
    for _ in range(n_chunks):
        np.random.randn(1, C - 1)
    return np.random


def test_oracle_vrnn_sampler_matches_executed_reference_generate_sample():
    for name, g in util.load_golden("vrnn_sampler.npz").items():
        c = g["cfg"]
        L, C, Z, D, H, xp = c["L"], c["C"], c["Z"], c["D"], c["H"], c["use_x_prev"]
        p = util.golden_params(g, O.vrnn_param_shapes(L, D, H, Z, C, xp))
        zl = util.golden_params(g, [("encoder_h.kernel", (D + C, 4 * H)), ("encoder_h.recurrent_kernel", (H, 4 * H)),
                                    ("encoder_h.bias", (4 * H,))], prefix="zenc/")
        p32 = {k: torch.tensor(v) for k, v in p.items()}
        enc = tuple(torch.tensor(zl[k]) for k in ("encoder_h.kernel", "encoder_h.recurrent_kernel", "encoder_h.bias"))
        x_seed = torch.tensor(g["x_seed"], dtype=torch.float32)
        one_d = x_seed.dim() == 1
        seed2d = x_seed[None, :] if one_d else x_seed
        nsteps = c["nsteps"] - 1 if one_d else c["nsteps"]      # a 1-D seed is its own first step (model.py:27-31)
        T = seed2d.shape[0] + nsteps
        np.random.seed(c["np_seed"])
        if c["infer"]:
            n_chunks = len([i for i in range(0, D, L) if i + L <= seed2d.shape[0]])     # quirk Q2
            assert n_chunks == c["n_w_calls"]
            for _ in range(n_chunks):
                np.random.randn(1, C - 1)
            w = O.vrnn_infer_w(p32, seed2d, L, C).double().numpy()
            if c["discrete"]:
                wn = np.zeros(C); wn[np.random.choice(C, p=w[0] / w[0].sum())] = 1.0
                w = wn[None, :]
        else:
            w = np.zeros((1, C)); w[0, c["label"]] = 1.0
        eps_z = np.zeros((T, Z), np.float32); u = np.zeros((T, D))
        for t in range(T):
            eps_z[t] = np.random.randn(Z); u[t] = np.random.rand(D)
        assert abs(np.random.rand() - c["stream_after"]) < 1e-15
        Xs, Ps = O.vrnn_generate_sample(p32, seed2d, nsteps, torch.tensor(w, dtype=torch.float32), torch.tensor(eps_z),
                                        torch.tensor(u, dtype=torch.float64), xp, enc_lstm=enc)
        _close(Ps.numpy(), g["probs"], 2e-5, 2e-6)
        out = Xs.numpy() if not one_d else np.concatenate([(u[:1] <= Ps.numpy()[:1]).astype(np.float32), Xs.numpy()])
        assert np.array_equal(out.astype(np.uint8), g["out"]), name


def test_oracle_vae_sampler_matches_executed_reference_generate_sample():
    for name, g in util.load_golden("vae_sampler.npz").items():
        c = g["cfg"]
        C, Z, D, H, Hc, xp = c["C"], c["Z"], c["D"], c["H"], c["Hc"], c["use_x_prev"]
        p = util.golden_params(g, O.vae_param_shapes(D, H, Z, Hc, C, xp))
        p32 = {k: torch.tensor(v) for k, v in p.items()}
        x_seed = torch.tensor(g["x_seed"], dtype=torch.float32)
        np.random.seed(c["np_seed"])
        if c["infer"]:
            np.random.randn(1, C - 1)
            w = O.vae_infer_w(p32, x_seed)
        else:
            w = torch.zeros(1, C); w[0, c["label"]] = 1.0
        T = c["nsteps"]
        eps_z = np.zeros((T, Z), np.float32); u = np.zeros((T, D))
        for t in range(T):
            eps_z[t] = np.random.randn(Z); u[t] = np.random.rand(D)
        assert abs(np.random.rand() - c["stream_after"]) < 1e-15
        Xs, Ps = O.vae_generate_sample(p32, x_seed, T, w, torch.tensor(eps_z), torch.tensor(u, dtype=torch.float64), xp,
                                       use_z_prior=c["use_z_prior"])
        _close(Ps.numpy(), g["probs"], 2e-5, 2e-6)
        assert np.array_equal(Xs.numpy().astype(np.uint8), g["out"]), name
