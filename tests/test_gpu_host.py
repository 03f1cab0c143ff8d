"""GPU tests of the reference-facing Python surface: get_model / fit / callbacks / save_weights /
load_model / generate_sample and the CLI entry points, on the bundled JSB pickles."""
import os
import types
import numpy as np
import pytest
import torch

import util
from oracle import clv_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data", "input")


def test_vrnn_fit_checkpoint_load_and_sample(tmp_path):
    from clvae_b200.cl_vrnn import model as M
    from clvae_b200.cl_vrnn.train import to_categorical
    from clvae_b200.utils.pianoroll import PianoData
    from clvae_b200.utils import model_utils
    import json
    P = PianoData(os.path.join(DATA, "JSB Chorales_all.pickle"), batch_size=200, seq_length=16, step_length=1,
                  return_y_next=True, return_y_hist=True, squeeze_x=False, squeeze_y=False)
    C = len(np.unique(P.train_song_keys))
    n, nv = 2000, 600
    w, wv = to_categorical(P.train_song_keys[:n], C), to_categorical(P.valid_song_keys[:nv], C)
    np.random.seed(0)
    model, _ = M.get_model(200, 88, 88, 2, 16, C, True, "adam-wn", seed=3)
    args = types.SimpleNamespace(model_dir=str(tmp_path), run_name="run", batch_size=200, seq_length=16,
                                 optimizer="adam-wn", original_dim=88, intermediate_dim=88, latent_dim=2,
                                 n_classes=C, use_x_prev=True, class_weight=1.0)
    model_utils.save_model_in_pieces(model, args)
    cbs = model_utils.get_callbacks(args, patience=5, min_epoch=1)
    hist = model.fit([P.y_train[:n], P.x_train[:n]], [P.y_train[:n], w, w, P.y_train[:n]], shuffle=True,
                     epochs=3, batch_size=200, callbacks=cbs,
                     validation_data=([P.y_valid[:nv], P.x_valid[:nv]], [P.y_valid[:nv], wv, wv, P.y_valid[:nv]]),
                     verbose=0)
    H = hist.history
    for k in ("loss", "X_decoded_mean_loss", "W_loss", "W2_loss", "Z_args_loss", "W_acc", "val_loss", "val_W_acc"):
        assert k in H and len(H[k]) == 3 and np.all(np.isfinite(H[k])), k
    assert H["loss"][-1] < H["loss"][0]                       # it learns
    assert os.path.exists(os.path.join(str(tmp_path), "run.h5"))
    assert json.load(open(os.path.join(str(tmp_path), "run.json")))["n_classes"] == C
    # reload (by order of weighted layers) and compare
    m2, _, margs = M.load_model(os.path.join(str(tmp_path), "run.h5"))
    assert margs["seq_length"] == 16 and margs["use_x_prev"] is True
    # the checkpoint holds the best-val epoch >= 1; the live model is at epoch 3 -- shapes must agree
    for a, b in zip(model.get_weights(), m2.get_weights()):
        assert a.shape == b.shape
    w_enc = M.make_w_encoder(m2, 88, C, 16)
    z_enc = M.make_z_encoder(m2, 88, C, (88, 2))
    dec = M.make_decoder(m2, 88, 88, 2, C, True)
    Ps = PianoData(os.path.join(DATA, "JSB Chorales_all.pickle"), batch_size=1, seq_length=32, squeeze_x=False)
    x_seed = Ps.x_test[5]
    s1 = M.generate_sample(dec, w_enc, z_enc, x_seed, 24, True, w_val=to_categorical(Ps.test_song_keys[5], C))
    s2 = M.generate_sample(dec, w_enc, z_enc, x_seed, 24, True, w_val=None, seq_length=16)
    for s in (s1, s2):
        assert s.shape == (24, 88) and s.dtype == np.float64 and set(np.unique(s)) <= {0.0, 1.0}


def test_generate_sample_matches_oracle_under_the_same_numpy_seed():
    from clvae_b200.cl_vrnn import model as M
    rng = np.random.default_rng(4)
    C, Z, L = 5, 2, 4
    p = O.init_vrnn_params(rng, L, 88, 88, Z, C, True)
    p["X_decoded_mean.bias"] = p["X_decoded_mean.bias"] - 1.5
    model, _ = M.get_model(1, 88, 88, Z, L, C, True, "adam-wn", seed=1, use_graph=False)
    model.engine.set_params({k: v.numpy() for k, v in p.items()})
    x_seed = O.synth_rolls(rng, 1, 8, 88, 0.1)[0]
    w_enc, z_enc, dec = M.make_w_encoder(model, 88, C, L), M.make_z_encoder(model, 88, C, (88, Z)), M.make_decoder(model, 88, 88, Z, C, True)
    np.random.seed(123)
    got = M.generate_sample(dec, w_enc, z_enc, x_seed, 10, True, w_val=None, seq_length=L)
    # oracle with the tape np.random would have produced in the reference's draw order
    np.random.seed(123)
    T = 8 + 10
    for _ in range(2):                 # two full 4-frame chunks of the 8-frame seed: sample_w draws per chunk
        np.random.randn(1, C - 1)      # (cl_vrnn/model.py:71-80; pinned by tests/golden/vrnn_sampler.npz)
    eps_z, u = np.zeros((T, Z), np.float32), np.zeros((T, 88), np.float32)
    for t in range(T):
        eps_z[t] = np.random.randn(Z); u[t] = np.random.rand(88)
    w = O.vrnn_infer_w(p, torch.tensor(x_seed, dtype=torch.float64), L, C)
    Xs, Ps = O.vrnn_generate_sample(p, torch.tensor(x_seed, dtype=torch.float64), 10, w,
                                    torch.tensor(eps_z, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), True)
    far = np.abs(Ps.numpy()[8:] - u[8:]) > 1e-6
    agree = (got == Xs.numpy())
    first_bad = np.argmax(~agree.all(axis=1)) if not agree.all() else len(agree)
    assert agree[:first_bad].all()
    assert first_bad == len(agree) or not far[first_bad].all()     # any divergence starts at a borderline draw


def test_keras_style_train_on_batch_with_independent_history():
    """current/history that do NOT overlap (arbitrary arrays) go through the [history | current]
    window layout (x_shift = L) and still match the oracle."""
    from clvae_b200.cl_vrnn import model as M
    case = util.make_vrnn_case(8, B=12, L=5, C=4, Z=2)
    rng = np.random.default_rng(1)
    cur = O.synth_rolls(rng, 12, 5, 88, 0.1); hist = O.synth_rolls(rng, 12, 5, 88, 0.1)
    model, _ = M.get_model(12, 88, 88, 2, 5, 4, True, "adam-wn", seed=2, use_graph=False)
    e = model.engine
    e.set_params({k: v.numpy() for k, v in case["p"].items()})
    wt = O.one_hot(case["labels"], 4).numpy()
    win = model._windows_from_inputs([cur, hist])
    assert win.shape == (12, 10, 88) and e.x_shift == 5
    e.stage_windows(torch.tensor(win).cuda(), torch.tensor(case["labels"]).cuda())
    e.eps_w.copy_(torch.tensor(case["eps_w"], dtype=torch.float32).reshape(-1))
    e.eps_z.copy_(torch.tensor(case["eps_z"], dtype=torch.float32).reshape(-1))
    e.run(train=True, gen_noise=False)
    lo = e.read_losses()
    p64 = {k: v for k, v in case["p"].items()}
    out, g = O.vrnn_loss_and_grads(p64, torch.tensor(cur, dtype=torch.float64), torch.tensor(hist, dtype=torch.float64),
                                   torch.tensor(wt, dtype=torch.float64), torch.tensor(case["eps_w"]),
                                   torch.tensor(case["eps_z"]), 4, True)
    assert abs(lo["loss"] - float(out["loss"])) < 1e-4 * float(out["loss"])
    for k in e.names:
        got = e.grad_view(k).cpu().numpy()
        assert util.rel_err(got, g[k].numpy().reshape(got.shape)) < 1e-4, k
    res = model.train_on_batch([cur, hist], [cur, wt, wt, cur])
    assert len(res) == len(model.metrics_names) == 6 and np.all(np.isfinite(res))


@pytest.mark.parametrize("use_graph", [False, True])
def test_packed_pinned_batch_and_mirrored_losses(use_graph):
    """train_on_batch_windows: a batch given as ONE pinned buffer [windows | labels] (single H2D copy) and
    as two separate tensors stages the same data, and the losses the scheduled-optimizer step mirrors
    into pinned host memory equal the ones copied back explicitly (eval pass, fused_optimizer off)."""
    from clvae_b200.cl_vrnn import model as M
    B, L, C = 16, 6, 4
    rng = np.random.default_rng(5)
    nw = B * (L + 1) * 88
    buf = torch.empty(nw + 4 * B, dtype=torch.uint8).pin_memory()
    buf[:nw] = torch.from_numpy((rng.random(nw) < 0.08).astype(np.uint8))
    lab = buf[nw:].view(torch.int32)
    lab.copy_(torch.from_numpy(rng.integers(0, C, B).astype(np.int32)))
    win = buf[:nw].view(B, L + 1, 88)
    models = []
    for fused in (True, False):
        m, _ = M.get_model(B, 88, 88, 2, L, C, True, "adam-wn", seed=3, use_graph=use_graph, fused_optimizer=fused)
        models.append(m)
    a, b = models
    b.engine.set_params(a.engine.get_params())
    for step in range(3):
        for m in (a, b):
            m.engine.rng_ctr.fill_(step)            # same Philox stream in both engines
        la = a.train_on_batch_windows(win, lab)                        # packed: one copy, mirrored losses
        lb = b.train_on_batch_windows(win.clone().pin_memory(), lab.clone().pin_memory())   # two copies, D2H
        assert torch.equal(a.engine.win_buf, b.engine.win_buf) and torch.equal(a.engine.labels, b.engine.labels)
        for k in la:
            assert abs(la[k] - lb[k]) <= 2e-5 * max(1.0, abs(lb[k])), (step, k, la[k], lb[k])


def test_vae_cli_train_then_sample_readme_example(tmp_path):
    """README.md:33-34: train a CL-VAE (latent_dim 4, --use_x_prev) on JSB Chorales_Cs, then sample."""
    from clvae_b200.cl_vae import train as T, sample as S
    mdir, sdir = str(tmp_path / "models"), str(tmp_path / "samples")
    a = T.build_parser().parse_args(["run1", "--use_x_prev", "--latent_dim", "4", "--num_epochs", "2",
                                     "--model_dir", mdir, "--train_file", os.path.join(DATA, "JSB Chorales_Cs.pickle")])
    model, best = T.train(a)
    assert a.n_classes == 2 and np.isfinite(best["val_loss"])
    assert os.path.exists(os.path.join(mdir, "run1.h5")) and os.path.exists(os.path.join(mdir, "run1.json"))
    s = S.build_parser().parse_args(["outfile", "--model_file", os.path.join(mdir, "run1.h5"), "--sample_dir", sdir,
                                     "-t", "16", "-n", "2", "--infer_w",
                                     "--train_file", os.path.join(DATA, "JSB Chorales_Cs.pickle")])
    S.sample(s)
    mids = sorted(os.listdir(sdir))
    assert mids == ["outfile_0.mid", "outfile_1.mid"]
    assert open(os.path.join(sdir, mids[0]), "rb").read(4) == b"MThd"


def test_fit_rolls_equals_fit_on_materialised_windows():
    """DeviceRolls (one roll per split + window offsets, no 17x window blow-up) feeds the same batches as the
    reference's materialised PianoData windows: same History under the same seeds."""
    from clvae_b200.cl_vrnn import model as M
    from clvae_b200.cl_vrnn.train import to_categorical
    from clvae_b200.utils.pianoroll import PianoData, DeviceRolls
    f = os.path.join(DATA, "JSB Chorales_all.pickle")
    P = PianoData(f, batch_size=200, seq_length=16, step_length=1, return_y_next=True, return_y_hist=True,
                  squeeze_x=False, squeeze_y=False)
    C = len(np.unique(P.train_song_keys))
    tr, _ = DeviceRolls.from_pickle(f, "train", 17, 200)
    va, _ = DeviceRolls.from_pickle(f, "valid", 17, 200)
    assert np.array_equal(tr.labels, P.train_song_keys) and np.array_equal(va.labels, P.valid_song_keys)
    w, wv = to_categorical(P.train_song_keys, C), to_categorical(P.valid_song_keys, C)
    hists = []
    for mode in ("windows", "rolls"):
        np.random.seed(5)
        model, _ = M.get_model(200, 88, 88, 2, 16, C, True, "adam-wn", seed=9)
        if mode == "windows":
            h = model.fit([P.y_train, P.x_train], [P.y_train, w, w, P.y_train], shuffle=True, epochs=1, batch_size=200,
                          validation_data=([P.y_valid, P.x_valid], [P.y_valid, wv, wv, P.y_valid]), verbose=0)
        else:
            h = model.fit_rolls(tr, va, shuffle=True, epochs=1, verbose=0)
        hists.append(h.history)
    for k in hists[0]:
        assert np.allclose(hists[0][k], hists[1][k], rtol=1e-6, atol=1e-7), (k, hists[0][k], hists[1][k])


def test_host_micro_batching_equals_the_full_batch_step():
    """Engine(micro_batch=Bm): B / Bm accumulated calls + one update == the one-call step (how B = 65 536 x
    L = 512, 177 GB of activations, is stepped in 180 GB of HBM)."""
    from clvae_b200.engine import Engine
    case = util.make_vrnn_case(11, B=24, L=6, C=5, Z=2)
    full = util.engine_for(case, "vrnn", use_graph=False)
    micro = util.engine_for(case, "vrnn", use_graph=False, micro_batch=8)
    assert micro.n_micro == 3 and micro.workspace.numel() < full.workspace.numel()
    for e in (full, micro):
        e.run(train=True, gen_noise=False)
    lf, lm = full.read_losses(), micro.read_losses()
    for k in lf:
        assert abs(lf[k] - lm[k]) <= 2e-5 * max(1.0, abs(lf[k])), (k, lf[k], lm[k])
    assert util.rel_err(micro.grads.cpu().numpy(), full.grads.cpu().numpy()) < 1e-4
    assert util.rel_err(micro.params.cpu().numpy(), full.params.cpu().numpy()) < 1e-5
    # automatic choice: a budget smaller than the full workspace forces a split
    auto = Engine("vrnn", 24, L=6, D=88, H=88, Z=2, n_classes=5, use_x_prev=True, use_graph=False,
                  workspace_budget_bytes=full.workspace.numel() * 4 // 2)
    assert auto.n_micro > 1 and 24 % auto.Bm == 0
