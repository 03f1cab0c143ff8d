"""
The CUDA path against fixtures produced by EXECUTING THE REFERENCE'S OWN SOURCE
(tests/golden/make_golden.py: cl_vrnn/model.py and cl_vae/model.py get_model + loss closures,
make_* sub-models, generate_sample, run over tests/golden/keras_shim.py):

  * one train step: the five loss scalars and every gradient tensor, 1e-4 (north_star tolerance);
  * generate_sample under the same np.random.seed: the piano roll, bit for bit (the fixtures keep a
    margin |p - u| > 1e-5 at every draw), and the per-step probabilities;
  * the reference's Python sampling loop re-run over the product's sub-model .predict() calls.
"""
import numpy as np
import pytest
import torch

import util
from oracle import clv_oracle as O
from test_golden_models import vrnn_case_from_golden, vae_case_from_golden

pytestmark = pytest.mark.gpu

KERAS_KEYS = {"vae": "X_decoded_mean_loss", "w_kl": "W_loss", "w_rec": "W2_loss", "z_kl": "Z_args_loss", "acc": "W_acc"}
TOL = 1e-4


def _check_grads(e, g):
    for k in e.names:
        got = e.grad_view(k).cpu().numpy()
        ref = g["g/" + k].reshape(got.shape)
        assert util.rel_err(got, ref) < TOL, (k, util.rel_err(got, ref))
        # element-wise form: small entries are checked against the tensor's own scale
        assert np.allclose(got, ref, rtol=1e-3, atol=TOL * np.abs(ref).max()), k


@pytest.mark.parametrize("name", ["xprev", "noxprev", "predict_next"])
@pytest.mark.parametrize("algo", [0, 1])
def test_vrnn_step_matches_executed_reference_graph(name, algo):
    g = util.load_golden("vrnn_train.npz")[name]
    case = vrnn_case_from_golden(g)
    case["p"] = {k: v.float() for k, v in case["p"].items()}
    e = util.engine_for(case, "vrnn", use_graph=False, gemm_algo=algo, predict_next=case["predict_next"], **case["kw"])
    e.run(train=True, gen_noise=False)
    lo = e.read_losses()
    res = g["cfg"]["res"]
    assert abs(lo["loss"] - res["loss"]) <= TOL * abs(res["loss"])
    for k, kk in KERAS_KEYS.items():
        assert abs(lo[k] - res[kk]) <= TOL * max(abs(res[kk]), 1e-3), (k, lo[k], res[kk])
    W = e.ws_view("W", (case["B"], case["C"])).cpu().numpy()
    assert np.allclose(W, g["out/W"], rtol=1e-4, atol=1e-6)
    Za = e.ws_view("Zargs", (case["B"], case["L"], 2 * case["Z"])).cpu().numpy()
    assert np.allclose(Za, g["out/Z_args"], rtol=1e-4, atol=1e-5)
    _check_grads(e, g)


@pytest.mark.parametrize("name", ["xprev", "noxprev"])
def test_vae_step_matches_executed_reference_graph(name):
    g = util.load_golden("vae_train.npz")[name]
    case = vae_case_from_golden(g)
    case["p"] = {k: v.float() for k, v in case["p"].items()}
    e = util.engine_for(case, "vae", use_graph=False, **case["kw"])
    e.run(train=True, gen_noise=False)
    lo = e.read_losses()
    res = g["cfg"]["res"]
    assert abs(lo["loss"] - res["loss"]) <= TOL * abs(res["loss"])
    for k, kk in KERAS_KEYS.items():
        kk = kk.replace("X_decoded_mean", "x_decoded_mean").replace("W2", "w2").replace("W_", "w_").replace("Z_args", "z_args")
        assert abs(lo[k] - res[kk]) <= TOL * max(abs(res[kk]), 1e-3), (k, lo[k], res[kk])
    _check_grads(e, g)


# ------------------------------------------------------------------------------------------ samplers
def _vrnn_model_from_golden(g):
    from clvae_b200.cl_vrnn import model as M
    c = g["cfg"]
    L, C, Z, D, H, xp = c["L"], c["C"], c["Z"], c["D"], c["H"], c["use_x_prev"]
    model, _ = M.get_model(8, D, H, Z, L, C, xp, "adam", seed=1, use_graph=False)
    model.engine.set_params(util.golden_params(g, O.vrnn_param_shapes(L, D, H, Z, C, xp)))
    zl = util.golden_params(g, [("encoder_h.kernel", (D + C, 4 * H)), ("encoder_h.recurrent_kernel", (H, 4 * H)),
                                ("encoder_h.bias", (4 * H,))], prefix="zenc/")
    w_enc = M.make_w_encoder(model, D, C, L)
    z_enc = M.make_z_encoder(model, D, C, (H, Z))
    # quirk Q1: the reference's z-encoder has its OWN encoder_h LSTM; the fixture recorded its weights
    z_enc.get_layer("encoder_h").set_weights([zl["encoder_h.kernel"], zl["encoder_h.recurrent_kernel"], zl["encoder_h.bias"]])
    dec = M.make_decoder(model, D, H, Z, C, xp)
    return M, model, w_enc, z_enc, dec


@pytest.mark.parametrize("name", ["given", "infer", "infer_discrete", "given_noxprev", "seed1d"])
def test_vrnn_generate_sample_equals_reference_run_under_the_same_numpy_seed(name):
    g = util.load_golden("vrnn_sampler.npz")[name]
    c = g["cfg"]
    M, model, w_enc, z_enc, dec = _vrnn_model_from_golden(g)
    w_val = None
    if not c["infer"]:
        w_val = np.zeros((1, c["C"])); w_val[0, c["label"]] = 1.0
    np.random.seed(c["np_seed"])
    out = M.generate_sample(dec, w_enc, z_enc, g["x_seed"].astype(np.float64), c["nsteps"], c["use_x_prev"], w_val=w_val,
                            w_discrete=c["discrete"], seq_length=c["L"])
    # the np.random stream was consumed draw for draw like the reference did
    assert abs(np.random.rand() - c["stream_after"]) < 1e-15
    assert out.shape == g["out"].shape and out.dtype == np.float64
    assert np.array_equal(out.astype(np.uint8), g["out"]), "min |p-u| margin of this fixture: %g" % c["min_margin"]


@pytest.mark.parametrize("name", ["given", "infer", "given_noxprev"])
def test_reference_python_loop_over_product_submodel_predicts(name):
    """generate_sample's loop as the reference wrote it (cl_vrnn/model.py:21-60), driving the product's
    make_* objects through .predict() / .reset_states(): per-step probabilities and Z heads equal the
    reference run's."""
    g = util.load_golden("vrnn_sampler.npz")[name]
    c = g["cfg"]
    M, model, w_enc, z_enc, dec = _vrnn_model_from_golden(g)
    x_seed = g["x_seed"].astype(np.float64)
    np.random.seed(c["np_seed"])
    for m in (dec, w_enc, z_enc):
        m.reset_states()
    nseed = x_seed.shape[0]
    if c["infer"]:
        w_ts = []
        for i in np.arange(0, x_seed.shape[1], c["L"]):
            xcs = x_seed[i:i + c["L"]]
            if xcs.shape[0] == c["L"]:
                w_ts.append(M.sample_w(w_enc.predict(xcs[None, :]), add_noise=False))
        w_t = np.vstack(w_ts).mean(axis=0)[None, :]
    else:
        w_t = np.zeros((1, c["C"])); w_t[0, c["label"]] = 1.0
    probs, zargs, Xs = [], [], []
    for t in range(c["nsteps"] + nseed):
        if t < nseed:
            x_prev = x_seed[t][None, None, :]
        za = z_enc.predict([x_prev, w_t])
        zargs.append(np.concatenate([a.reshape(-1) for a in za]))
        z_t = M.sample_z(za)
        p = dec.predict([z_t, x_prev, w_t] if c["use_x_prev"] else [z_t, w_t])
        probs.append(p.reshape(-1))
        x_t = M.sample_x(p)
        x_prev = x_t
        Xs.append(x_t.reshape(-1))
    assert np.allclose(np.stack(zargs), g["zargs"], rtol=2e-5, atol=2e-6)
    assert np.allclose(np.stack(probs), g["probs"], rtol=2e-5, atol=2e-6)
    assert np.array_equal(np.stack(Xs)[nseed:].astype(np.uint8), g["out"])


@pytest.mark.parametrize("name", ["given", "infer", "zprior_noxprev"])
def test_vae_generate_sample_equals_reference_run_under_the_same_numpy_seed(name):
    from clvae_b200.cl_vae import model as M
    g = util.load_golden("vae_sampler.npz")[name]
    c = g["cfg"]
    C, Z, D, H, Hc, xp = c["C"], c["Z"], c["D"], c["H"], c["Hc"], c["use_x_prev"]
    model, _ = M.get_model(1, D, (H, Z), (Hc, C), "adam", use_x_prev=xp, seed=1, use_graph=False)
    model.engine.set_params(util.golden_params(g, O.vae_param_shapes(D, H, Z, Hc, C, xp)))
    w_enc = M.make_w_encoder(model, D)
    z_enc = M.make_z_encoder(model, D, C, (H, Z))
    dec = M.make_decoder(model, (H, Z), C, use_x_prev=xp)
    w_val = None
    if not c["infer"]:
        w_val = np.zeros((1, C)); w_val[0, c["label"]] = 1.0
    np.random.seed(c["np_seed"])
    out = M.generate_sample(dec, w_enc, z_enc, g["x_seed"].astype(np.float64), c["nsteps"], w_val=w_val,
                            use_z_prior=c["use_z_prior"], use_x_prev=xp)
    assert abs(np.random.rand() - c["stream_after"]) < 1e-15
    assert np.array_equal(out.astype(np.uint8), g["out"])
    # the reference loop over the product's .predict() objects (cl_vae/model.py:20-42)
    np.random.seed(c["np_seed"])
    x_prev = np.expand_dims(g["x_seed"].astype(np.float64), axis=0)
    x_prev_t = x_prev
    w_t = M.sample_w(w_enc.predict(x_prev), add_noise=False) if c["infer"] else w_val
    probs = []
    for t in range(c["nsteps"]):
        z_mean, z_log_var = z_enc.predict([x_prev, w_t])
        z_t = M.sample_z((0 * z_mean, 0 * z_log_var)) if c["use_z_prior"] else M.sample_z((z_mean, z_log_var))
        p = dec.predict([w_t, z_t, x_prev_t] if xp else [w_t, z_t])
        probs.append(p.reshape(-1))
        x_t = M.sample_x(p)
        x_prev_t = x_prev
        x_prev = x_t
    assert np.allclose(np.stack(probs), g["probs"], rtol=2e-5, atol=2e-6)
