"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Hand-derived backward pass of the CL-VRNN / CL-VAE training graph in plain numpy.  The reference
never writes a backward (it is TF autodiff of cl_vrnn/model.py:164-264 / cl_vae/model.py:130-218);
this file is the derivation the CUDA kernels implement, and tests/test_oracle.py checks it against
torch.autograd of oracle/clv_oracle.py to 1e-9 in float64.  Same [K2-recall] assumptions as there.
"""
import math
import numpy as np

EPS = 1e-7


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _hard_sigmoid(a):
    return np.clip(0.2 * a + 0.5, 0.0, 1.0)


def _hs_pass(a):
    y = 0.2 * a + 0.5
    return ((y >= 0.0) & (y <= 1.0)).astype(a.dtype)  # closed interval [K2-recall (8)]


def lstm_fwd(xproj, U):
    """xproj [B,L,4H] (input projection incl. bias) -> h [B,L,H], stash (pre-activations a, c)."""
    B, L, H4 = xproj.shape
    H = H4 // 4
    h = np.zeros((B, H), xproj.dtype); c = np.zeros((B, H), xproj.dtype)
    hs = np.zeros((B, L, H), xproj.dtype); cs = np.zeros((B, L, H), xproj.dtype)
    a_all = np.zeros_like(xproj)
    for t in range(L):
        a = xproj[:, t] + h @ U
        i = _hard_sigmoid(a[:, :H]); f = _hard_sigmoid(a[:, H:2 * H])
        g = np.tanh(a[:, 2 * H:3 * H]); o = _hard_sigmoid(a[:, 3 * H:])
        c = f * c + i * g
        h = o * np.tanh(c)
        hs[:, t] = h; cs[:, t] = c; a_all[:, t] = a
    return hs, cs, a_all


def lstm_bwd(dh_out, hs, cs, a_all, U):
    """BPTT.  dh_out [B,L,H] = dLoss/dh_t from the layers above.  Returns dA [B,L,4H] (gradient wrt
    the pre-activations == wrt xproj) ; dU is Hprev^T @ dA."""
    B, L, H = hs.shape
    dA = np.zeros_like(a_all)
    dh_rec = np.zeros((B, H), hs.dtype); dc = np.zeros((B, H), hs.dtype)
    for t in range(L - 1, -1, -1):
        a = a_all[:, t]
        i = _hard_sigmoid(a[:, :H]); f = _hard_sigmoid(a[:, H:2 * H])
        g = np.tanh(a[:, 2 * H:3 * H]); o = _hard_sigmoid(a[:, 3 * H:])
        c_prev = cs[:, t - 1] if t > 0 else np.zeros((B, H), hs.dtype)
        tc = np.tanh(cs[:, t])
        dh = dh_out[:, t] + dh_rec
        do = dh * tc
        dc = dc + dh * o * (1.0 - tc * tc)
        di = dc * g; dg = dc * i; df = dc * c_prev
        dA[:, t, :H] = di * 0.2 * _hs_pass(a[:, :H])
        dA[:, t, H:2 * H] = df * 0.2 * _hs_pass(a[:, H:2 * H])
        dA[:, t, 2 * H:3 * H] = dg * (1.0 - g * g)
        dA[:, t, 3 * H:] = do * 0.2 * _hs_pass(a[:, 3 * H:])
        dh_rec = dA[:, t] @ U.T
        dc = dc * f
    return dA


def logitnormal_fwd(Wm, Wlv, eps_w, w_true, C, w_log_var_prior):
    """K2 forward: W, per-row w_kl (cl_vrnn/model.py:247-252), per-row w_rec on W2=W+1e-10
    (:244-245,255), correct flag (categorical accuracy)."""
    s = Wm + np.exp(Wlv / 2) * eps_w
    w0 = np.concatenate([s, np.zeros((s.shape[0], 1), s.dtype)], axis=-1)
    num = np.exp(w0)
    W = num / num.sum(-1, keepdims=True)
    ep = math.exp(w_log_var_prior)
    wkl = -0.5 * (1 - w_log_var_prior + Wlv - np.exp(Wlv) / ep - Wm ** 2 / ep).sum(-1)
    W2 = W + 1e-10
    S = W2.sum(-1, keepdims=True)
    q = W2 / S
    qc = np.clip(q, EPS, 1 - EPS)
    wrec = (C - 1) * -(w_true * np.log(qc)).sum(-1)
    correct = (W.argmax(-1) == w_true.argmax(-1)).astype(W.dtype)
    return W, wkl, wrec, correct


def logitnormal_bwd(Wm, Wlv, eps_w, w_true, W, dW_ext, C, w_log_var_prior, cw_over_B, wkl_over_B):
    """K2 backward.  dW_ext [B,C] = gradient reaching W from the two LSTMs' W columns.
    cw_over_B = class_weight/B, wkl_over_B = w_kl_weight/B.  Returns dWm, dWlv."""
    W2 = W + 1e-10
    S = W2.sum(-1, keepdims=True)
    q = W2 / S
    qpass = ((q >= EPS) & (q <= 1 - EPS)).astype(W.dtype)
    qc = np.clip(q, EPS, 1 - EPS)
    dq = -(C - 1) * w_true / qc * qpass * cw_over_B
    dW2 = dq / S - (dq * W2).sum(-1, keepdims=True) / (S * S)
    dW = dW_ext + dW2
    ds = (W * (dW - (dW * W).sum(-1, keepdims=True)))[:, :C - 1]
    ep = math.exp(w_log_var_prior)
    dWm = ds + wkl_over_B * Wm / ep
    dWlv = ds * eps_w * 0.5 * np.exp(Wlv / 2) + wkl_over_B * (-0.5) * (1 - np.exp(Wlv) / ep)
    return dWm, dWlv


def bernoulli_fwd_bwd(logits, x, scale):
    """K4: per-row vae loss (sum over keys of Keras BCE) and dlogits = scale*(p-x)*pass."""
    p = _sigmoid(logits)
    pc = np.clip(p, EPS, 1 - EPS)
    l = np.log(pc / (1 - pc))
    bce = np.maximum(l, 0) - l * x + np.log1p(np.exp(-np.abs(l)))
    ppass = ((p >= EPS) & (p <= 1 - EPS)).astype(p.dtype)
    return bce.sum(-1), scale * (pc - x) * ppass


def vrnn_manual(p, X, Xp, w_true, eps_w, eps_z, C, use_x_prev, class_weight=1.0, kl_weight=1.0,
                w_kl_weight=1.0, w_log_var_prior=0.0):
    """Full CL-VRNN forward + hand-derived backward in numpy.  p: dict name -> ndarray."""
    B, L, D = X.shape
    H = p["encoder_h.recurrent_kernel"].shape[0]
    Z = p["Z_mean.kernel"].shape[1]
    BL = B * L
    Xf = X.reshape(B, L * D)
    # ---- forward
    hW = np.maximum(Xf @ p["hW.kernel"] + p["hW.bias"], 0)
    Wargs = hW @ p["Wargs.kernel"] + p["Wargs.bias"]
    Wm, Wlv = Wargs[:, :C - 1], Wargs[:, C - 1:]
    W, wkl, wrec, correct = logitnormal_fwd(Wm, Wlv, eps_w, w_true, C, w_log_var_prior)
    Ke, Ue, be = p["encoder_h.kernel"], p["encoder_h.recurrent_kernel"], p["encoder_h.bias"]
    xproj_e = X @ Ke[:D] + (W @ Ke[D:] + be)[:, None, :]
    h_e, c_e, a_e = lstm_fwd(xproj_e, Ue)
    mu = h_e @ p["Z_mean.kernel"] + p["Z_mean.bias"]
    lv = h_e @ p["Z_log_var.kernel"] + p["Z_log_var.bias"]
    Zs = mu + np.exp(lv / 2) * eps_z
    zkl = -0.5 * (1 + lv - mu ** 2 - np.exp(lv)).sum(-1)
    Kd, Ud, bd = p["decoder_h.kernel"], p["decoder_h.recurrent_kernel"], p["decoder_h.bias"]
    xo = D if use_x_prev else 0
    xproj_d = Zs @ Kd[xo:xo + Z] + (W @ Kd[xo + Z:] + bd)[:, None, :]
    if use_x_prev:
        xproj_d = xproj_d + Xp @ Kd[:D]
    h_d, c_d, a_d = lstm_fwd(xproj_d, Ud)
    logits = h_d @ p["X_decoded_mean.kernel"] + p["X_decoded_mean.bias"]
    vae, dlogits = bernoulli_fwd_bwd(logits, X, 1.0 / BL)
    losses = dict(vae=vae.mean(), w_kl=wkl.mean(), w_rec=wrec.mean(), z_kl=zkl.mean(),
                  acc=correct.mean())
    losses["loss"] = (losses["vae"] + w_kl_weight * losses["w_kl"] + class_weight * losses["w_rec"]
                      + kl_weight * losses["z_kl"])
    # ---- backward
    g = {}
    g["X_decoded_mean.kernel"] = h_d.reshape(BL, H).T @ dlogits.reshape(BL, D)
    g["X_decoded_mean.bias"] = dlogits.sum((0, 1))
    dh_d = dlogits @ p["X_decoded_mean.kernel"].T
    dA_d = lstm_bwd(dh_d, h_d, c_d, a_d, Ud)
    hprev_d = np.concatenate([np.zeros((B, 1, H)), h_d[:, :-1]], axis=1)
    g["decoder_h.recurrent_kernel"] = hprev_d.reshape(BL, H).T @ dA_d.reshape(BL, 4 * H)
    g["decoder_h.bias"] = dA_d.sum((0, 1))
    dAsum_d = dA_d.sum(1)
    gKd = np.zeros_like(Kd)
    if use_x_prev:
        gKd[:D] = Xp.reshape(BL, D).T @ dA_d.reshape(BL, 4 * H)
    gKd[xo:xo + Z] = Zs.reshape(BL, Z).T @ dA_d.reshape(BL, 4 * H)
    gKd[xo + Z:] = W.T @ dAsum_d
    g["decoder_h.kernel"] = gKd
    dZ = dA_d @ Kd[xo:xo + Z].T
    dW_ext = dAsum_d @ Kd[xo + Z:].T
    dmu = dZ + (kl_weight / BL) * mu
    dlv = dZ * eps_z * 0.5 * np.exp(lv / 2) + (kl_weight / BL) * 0.5 * (np.exp(lv) - 1)
    g["Z_mean.kernel"] = h_e.reshape(BL, H).T @ dmu.reshape(BL, Z)
    g["Z_mean.bias"] = dmu.sum((0, 1))
    g["Z_log_var.kernel"] = h_e.reshape(BL, H).T @ dlv.reshape(BL, Z)
    g["Z_log_var.bias"] = dlv.sum((0, 1))
    dh_e = dmu @ p["Z_mean.kernel"].T + dlv @ p["Z_log_var.kernel"].T
    dA_e = lstm_bwd(dh_e, h_e, c_e, a_e, Ue)
    hprev_e = np.concatenate([np.zeros((B, 1, H)), h_e[:, :-1]], axis=1)
    g["encoder_h.recurrent_kernel"] = hprev_e.reshape(BL, H).T @ dA_e.reshape(BL, 4 * H)
    g["encoder_h.bias"] = dA_e.sum((0, 1))
    dAsum_e = dA_e.sum(1)
    gKe = np.zeros_like(Ke)
    gKe[:D] = X.reshape(BL, D).T @ dA_e.reshape(BL, 4 * H)
    gKe[D:] = W.T @ dAsum_e
    g["encoder_h.kernel"] = gKe
    dW_ext = dW_ext + dAsum_e @ Ke[D:].T
    dWm, dWlv = logitnormal_bwd(Wm, Wlv, eps_w, w_true, W, dW_ext, C, w_log_var_prior,
                                class_weight / B, w_kl_weight / B)
    dWargs = np.concatenate([dWm, dWlv], axis=-1)
    g["Wargs.kernel"] = hW.T @ dWargs
    g["Wargs.bias"] = dWargs.sum(0)
    dhW = (dWargs @ p["Wargs.kernel"].T) * (hW > 0)
    g["hW.kernel"] = Xf.T @ dhW
    g["hW.bias"] = dhW.sum(0)
    return losses, g


def vae_manual(p, x, xp, w_true, eps_w, eps_z, C, use_x_prev, class_weight=1.0, kl_weight=1.0,
               w_kl_weight=1.0, w_log_var_prior=0.0):
    """Full CL-VAE forward + hand-derived backward in numpy (cl_vae/model.py:130-218)."""
    B, D = x.shape
    Z = p["z_mean.kernel"].shape[1]
    h_w = np.maximum(x @ p["h_w.kernel"] + p["h_w.bias"], 0)
    Wm = h_w @ p["w_mean.kernel"] + p["w_mean.bias"]
    Wlv = h_w @ p["w_log_var.kernel"] + p["w_log_var.bias"]
    W, wkl, wrec, correct = logitnormal_fwd(Wm, Wlv, eps_w, w_true, C, w_log_var_prior)
    Kh = p["h.kernel"]
    h = np.maximum(x @ Kh[:D] + W @ Kh[D:] + p["h.bias"], 0)
    mu = h @ p["z_mean.kernel"] + p["z_mean.bias"]
    lv = h @ p["z_log_var.kernel"] + p["z_log_var.bias"]
    z = mu + np.exp(lv / 2) * eps_z
    zkl = -0.5 * (1 + lv - mu ** 2 - np.exp(lv)).sum(-1)
    Kdh = p["decoder_h.kernel"]
    xo = D if use_x_prev else 0
    pre = W @ Kdh[:C] + z @ Kdh[C + xo:] + p["decoder_h.bias"]
    if use_x_prev:
        pre = pre + xp @ Kdh[C:C + D]
    hd = np.maximum(pre, 0)
    logits = hd @ p["x_decoded_mean.kernel"] + p["x_decoded_mean.bias"]
    vae, dlogits = bernoulli_fwd_bwd(logits, x, 1.0 / B)
    losses = dict(vae=vae.mean(), w_kl=wkl.mean(), w_rec=wrec.mean(), z_kl=zkl.mean(),
                  acc=correct.mean())
    losses["loss"] = (losses["vae"] + w_kl_weight * losses["w_kl"] + class_weight * losses["w_rec"]
                      + kl_weight * losses["z_kl"])
    g = {}
    g["x_decoded_mean.kernel"] = hd.T @ dlogits
    g["x_decoded_mean.bias"] = dlogits.sum(0)
    dpre = (dlogits @ p["x_decoded_mean.kernel"].T) * (hd > 0)
    gK = np.zeros_like(Kdh)
    gK[:C] = W.T @ dpre
    if use_x_prev:
        gK[C:C + D] = xp.T @ dpre
    gK[C + xo:] = z.T @ dpre
    g["decoder_h.kernel"] = gK
    g["decoder_h.bias"] = dpre.sum(0)
    dW_ext = dpre @ Kdh[:C].T
    dz = dpre @ Kdh[C + xo:].T
    dmu = dz + (kl_weight / B) * mu
    dlv = dz * eps_z * 0.5 * np.exp(lv / 2) + (kl_weight / B) * 0.5 * (np.exp(lv) - 1)
    g["z_mean.kernel"] = h.T @ dmu; g["z_mean.bias"] = dmu.sum(0)
    g["z_log_var.kernel"] = h.T @ dlv; g["z_log_var.bias"] = dlv.sum(0)
    dh = (dmu @ p["z_mean.kernel"].T + dlv @ p["z_log_var.kernel"].T) * (h > 0)
    gKh = np.zeros_like(Kh)
    gKh[:D] = x.T @ dh
    gKh[D:] = W.T @ dh
    g["h.kernel"] = gKh; g["h.bias"] = dh.sum(0)
    dW_ext = dW_ext + dh @ Kh[D:].T
    dWm, dWlv = logitnormal_bwd(Wm, Wlv, eps_w, w_true, W, dW_ext, C, w_log_var_prior,
                                class_weight / B, w_kl_weight / B)
    g["w_mean.kernel"] = h_w.T @ dWm; g["w_mean.bias"] = dWm.sum(0)
    g["w_log_var.kernel"] = h_w.T @ dWlv; g["w_log_var.bias"] = dWlv.sum(0)
    dh_w = (dWm @ p["w_mean.kernel"].T + dWlv @ p["w_log_var.kernel"].T) * (h_w > 0)
    g["h_w.kernel"] = x.T @ dh_w; g["h_w.bias"] = dh_w.sum(0)
    return losses, g
