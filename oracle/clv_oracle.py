"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (PyTorch-CPU, float64 or float32) of the CL-VRNN / CL-VAE hot path of
mobeets/classifying-vae-lstm.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

PARITY STATUS.  The reference is Python 2 + Keras 2.0.0 + TensorFlow 1.0.1 (requirements.txt:1-2), none of
which exist in this environment, and it ships no tests, golden vectors or saved weights.  What pins this
oracle (tests/golden/, produced by tests/golden/make_golden.py IN the build container, where /root/reference
exists; checked by tests/test_golden_models.py, tests/test_oracle.py, tests/test_pianoroll.py):

  * PINNED to the reference's own source, executed: the whole of code/cl_vrnn/model.py and code/cl_vae/model.py
    -- get_model (graph wiring, concat orders, Lambdas, the four loss closures, loss_weights, metrics),
    make_w_encoder / make_z_encoder / make_decoder, generate_sample, sample_x/_w/_z/_w_discrete -- run (after a
    mechanical tuple-parameter rewrite, nothing else) against tests/golden/keras_shim.py, a recorder shim of
    the Keras functional API on PyTorch-CPU float64.  Losses agree with this oracle to 1e-12, every gradient
    tensor to float32 storage precision, sampled rolls bit for bit under the same np.random.seed (including the
    draw order: randn(1, C-1) per inferred-key chunk even without noise, one np.random.choice for w_discrete,
    then randn(z), rand(88) per step), for use_x_prev on/off, --predict_next, inferred / discrete / given keys,
    1-D seeds, CL-VAE use_z_prior.  Also executed unmodified: utils/pianoroll.py (PianoData on both bundled
    pickles) and utils/weightnorm.py (AdamWithWeightnorm.get_updates on a numpy shim).
  * STILL "[K2-recall]" (restated from the published Keras 2.0.0 / TF 1.0.1 sources inside the shim and here,
    because those packages cannot be installed): the PRIMITIVES the reference calls --

  (1) LSTM: gate order i,f,c,o; recurrent_activation hard_sigmoid = clip(0.2x+0.5,0,1);
      activation tanh; weights [kernel, recurrent_kernel, bias]; h0=c0=0.
  (2) _EPSILON = 1e-7.
  (3) K.binary_crossentropy = clip(p,eps,1-eps) -> log(p/(1-p)) -> sigmoid_cross_entropy_with_logits
      = max(l,0) - l*x + log1p(exp(-|l|)); losses.binary_crossentropy = mean over last axis.
  (4) K.categorical_crossentropy renormalises by the row sum, clips to [eps,1-eps], -sum t*log q.
  (5) every output's loss is averaged over all non-feature axes; total = loss-weighted sum.
  (6) metric 'accuracy' on W with a custom loss = categorical accuracy.
  (7) Keras Adam: p -= lr_t*m/(sqrt(v)+eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t).
  (8) clip_by_value passes its gradient on the closed interval (torch.clamp does the same).

All noise is an explicit input (eps_w, eps_z, u) so the CUDA path can be compared on identical draws.
Parameter containers are plain dicts name -> tensor in Keras [in,out] layout.
"""
import math
import numpy as np
import torch

EPS = 1e-7  # keras.backend.common._EPSILON [K2-recall]


# --------------------------------------------------------------------------------------
# parameter tables
# --------------------------------------------------------------------------------------
def vrnn_param_shapes(L, D, H, Z, C, use_x_prev):
    """Keras weighted-layer order of cl_vrnn/model.py:164-234 (hW, Wargs, encoder_h, Z_mean,
    Z_log_var, decoder_h, X_decoded_mean).  Note hW has `original_dim` units (model.py:174)."""
    in_d = (D if use_x_prev else 0) + Z + C
    return [
        ("hW.kernel", (L * D, D)), ("hW.bias", (D,)),
        ("Wargs.kernel", (D, 2 * (C - 1))), ("Wargs.bias", (2 * (C - 1),)),
        ("encoder_h.kernel", (D + C, 4 * H)), ("encoder_h.recurrent_kernel", (H, 4 * H)),
        ("encoder_h.bias", (4 * H,)),
        ("Z_mean.kernel", (H, Z)), ("Z_mean.bias", (Z,)),
        ("Z_log_var.kernel", (H, Z)), ("Z_log_var.bias", (Z,)),
        ("decoder_h.kernel", (in_d, 4 * H)), ("decoder_h.recurrent_kernel", (H, 4 * H)),
        ("decoder_h.bias", (4 * H,)),
        ("X_decoded_mean.kernel", (H, D)), ("X_decoded_mean.bias", (D,)),
    ]


def vae_param_shapes(D, H, Z, Hc, C, use_x_prev):
    """Keras weighted-layer order of cl_vae/model.py:130-188 (h_w, w_mean, w_log_var, h, z_mean,
    z_log_var, decoder_h, x_decoded_mean); H=latent_dim_0 (intermediate_dim), Hc=class_dim_0."""
    in_dec = C + (D if use_x_prev else 0) + Z
    return [
        ("h_w.kernel", (D, Hc)), ("h_w.bias", (Hc,)),
        ("w_mean.kernel", (Hc, C - 1)), ("w_mean.bias", (C - 1,)),
        ("w_log_var.kernel", (Hc, C - 1)), ("w_log_var.bias", (C - 1,)),
        ("h.kernel", (D + C, H)), ("h.bias", (H,)),
        ("z_mean.kernel", (H, Z)), ("z_mean.bias", (Z,)),
        ("z_log_var.kernel", (H, Z)), ("z_log_var.bias", (Z,)),
        ("decoder_h.kernel", (in_dec, H)), ("decoder_h.bias", (H,)),
        ("x_decoded_mean.kernel", (H, D)), ("x_decoded_mean.bias", (D,)),
    ]


def _glorot_uniform(rng, shape):
    lim = math.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape)


def _orthogonal(rng, shape):
    a = rng.standard_normal(shape)
    u, _, vt = np.linalg.svd(a, full_matrices=False)
    return u if u.shape == shape else vt


def init_vrnn_params(rng, L, D, H, Z, C, use_x_prev, dtype=torch.float64):
    """Keras default initialisers [K2-recall]: Dense glorot_uniform/zeros; LSTM kernel glorot_uniform,
    recurrent orthogonal, bias zeros with unit forget bias; Z heads and X head RandomNormal(0,0.1)
    (cl_vrnn/model.py:200-207,229-233)."""
    p = {}
    for name, shp in vrnn_param_shapes(L, D, H, Z, C, use_x_prev):
        if name.endswith(".bias"):
            v = np.zeros(shp)
            if name in ("encoder_h.bias", "decoder_h.bias"):
                v[H:2 * H] = 1.0
        elif name.endswith("recurrent_kernel"):
            v = _orthogonal(rng, shp)
        elif name.split(".")[0] in ("Z_mean", "Z_log_var", "X_decoded_mean"):
            v = rng.normal(0.0, 0.1, size=shp)
        else:
            v = _glorot_uniform(rng, shp)
        p[name] = torch.tensor(v, dtype=dtype)
    return p


def init_vae_params(rng, D, H, Z, Hc, C, use_x_prev, dtype=torch.float64):
    p = {}
    for name, shp in vae_param_shapes(D, H, Z, Hc, C, use_x_prev):
        v = np.zeros(shp) if name.endswith(".bias") else _glorot_uniform(rng, shp)
        p[name] = torch.tensor(v, dtype=dtype)
    return p


# --------------------------------------------------------------------------------------
# primitive ops at the reference's op granularity
# --------------------------------------------------------------------------------------
def hard_sigmoid(x):
    """keras.backend.tensorflow_backend.hard_sigmoid [K2-recall]: clip(0.2x+0.5, 0, 1)."""
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def lstm_forward(xin, kernel, rkernel, bias, h0=None, c0=None):
    """Keras 2.0.0 LSTM, implementation=0 (cl_vrnn/model.py:196-199,225-228) [K2-recall].
    xin [B,L,In] -> all h_t [B,L,H]; also returns final (h,c) for the stateful sampler."""
    B, L, _ = xin.shape
    H = rkernel.shape[0]
    xproj = xin @ kernel + bias  # hoisted input projection (preprocess_input)
    h = xin.new_zeros(B, H) if h0 is None else h0
    c = xin.new_zeros(B, H) if c0 is None else c0
    hs = []
    for t in range(L):
        a = xproj[:, t] + h @ rkernel
        i = hard_sigmoid(a[:, 0 * H:1 * H])
        f = hard_sigmoid(a[:, 1 * H:2 * H])
        g = torch.tanh(a[:, 2 * H:3 * H])
        o = hard_sigmoid(a[:, 3 * H:4 * H])
        c = f * c + i * g
        h = o * torch.tanh(c)
        hs.append(h)
    return torch.stack(hs, dim=1), h, c


def logistic_normal(w_mean, w_log_var, eps_w):
    """sampling_w, cl_vrnn/model.py:183-191 == w_sampling, cl_vae/model.py:146-157.
    Un-stabilised softmax over [s, 0]."""
    s = w_mean + torch.exp(w_log_var / 2) * eps_w
    w0 = torch.cat([s, s.new_zeros(s.shape[0], 1)], dim=-1)
    num = torch.exp(w0)
    return num / num.sum(dim=-1, keepdim=True)


def keras_bce(x, p):
    """K.binary_crossentropy(output=p, target=x) [K2-recall (3)]; elementwise."""
    pc = torch.clamp(p, EPS, 1.0 - EPS)
    l = torch.log(pc / (1.0 - pc))
    return torch.clamp(l, min=0.0) - l * x + torch.log1p(torch.exp(-torch.abs(l)))


def vae_loss(x, p, original_dim):
    """cl_vrnn/model.py:241-242 == cl_vae/model.py:190-191."""
    return original_dim * keras_bce(x, p).mean(dim=-1)


def keras_cce(t, q):
    """K.categorical_crossentropy(output=q, target=t) [K2-recall (4)]."""
    q = q / q.sum(dim=-1, keepdim=True)
    q = torch.clamp(q, EPS, 1.0 - EPS)
    return -(t * torch.log(q)).sum(dim=-1)


def w_rec_loss(w_true, w2, n_classes):
    """cl_vrnn/model.py:244-245."""
    return (n_classes - 1) * keras_cce(w_true, w2)


def w_kl_loss(w_mean, w_log_var, w_log_var_prior):
    """cl_vrnn/model.py:247-252 (closes over W_mean, W_log_var; ignores its arguments)."""
    ep = math.exp(w_log_var_prior)
    vs = 1 - w_log_var_prior + w_log_var - torch.exp(w_log_var) / ep - w_mean ** 2 / ep
    return -0.5 * vs.sum(dim=-1)


def z_kl_loss(z_mean, z_log_var):
    """kl_loss, cl_vrnn/model.py:236-239."""
    return -0.5 * (1 + z_log_var - z_mean ** 2 - torch.exp(z_log_var)).sum(dim=-1)


def categorical_accuracy(w_true, w):
    return (w_true.argmax(dim=-1) == w.argmax(dim=-1)).to(w.dtype).mean()


# --------------------------------------------------------------------------------------
# CL-VRNN training graph (cl_vrnn/model.py:164-267)
# --------------------------------------------------------------------------------------
def vrnn_forward(p, X, Xp, w_true, eps_w, eps_z, C, use_x_prev, class_weight=1.0, kl_weight=1.0,
                 w_kl_weight=1.0, w_log_var_prior=0.0, Y=None):
    """X=current [B,L,D], Xp=history [B,L,D] or None, w_true one-hot [B,C], eps_w [B,C-1],
    eps_z [B,L,Z]; Y = reconstruction target (cl_vrnn/train.py:51-63: y == current except with
    --predict_next, where the input is frames 0..L-1 and the target frames 1..L).  Returns dict with
    total loss, the 4 per-output mean losses, accuracy and intermediates."""
    B, L, D = X.shape
    hW = torch.relu(X.reshape(B, L * D) @ p["hW.kernel"] + p["hW.bias"])          # model.py:174
    Wargs = hW @ p["Wargs.kernel"] + p["Wargs.bias"]                              # model.py:175
    W_mean, W_log_var = Wargs[:, :C - 1], Wargs[:, C - 1:]                        # model.py:176-181
    W = logistic_normal(W_mean, W_log_var, eps_w)                                 # model.py:183-191
    Wrep = W[:, None, :].expand(B, L, C)
    XW = torch.cat([X, Wrep], dim=-1)                                             # model.py:193
    h_e, _, _ = lstm_forward(XW, p["encoder_h.kernel"], p["encoder_h.recurrent_kernel"],
                             p["encoder_h.bias"])                                 # model.py:196-199
    Z_mean = h_e @ p["Z_mean.kernel"] + p["Z_mean.bias"]                          # model.py:208
    Z_log_var = h_e @ p["Z_log_var.kernel"] + p["Z_log_var.bias"]                 # model.py:209
    Zs = Z_mean + torch.exp(Z_log_var / 2) * eps_z                                # model.py:212-216
    parts = ([Xp] if use_x_prev else []) + [Zs, Wrep]                             # model.py:218-222
    h_d, _, _ = lstm_forward(torch.cat(parts, dim=-1), p["decoder_h.kernel"],
                             p["decoder_h.recurrent_kernel"], p["decoder_h.bias"])  # :225-228
    logits = h_d @ p["X_decoded_mean.kernel"] + p["X_decoded_mean.bias"]
    P = torch.sigmoid(logits)                                                     # model.py:229-234
    W2 = W + 1e-10                                                                # model.py:255
    l_vae = vae_loss(X if Y is None else Y, P, D).mean()
    l_wkl = w_kl_loss(W_mean, W_log_var, w_log_var_prior).mean()
    l_wrec = w_rec_loss(w_true, W2, C).mean()
    l_zkl = z_kl_loss(Z_mean, Z_log_var).mean()
    total = 1.0 * l_vae + w_kl_weight * l_wkl + class_weight * l_wrec + kl_weight * l_zkl  # :261-264
    return dict(loss=total, vae=l_vae, w_kl=l_wkl, w_rec=l_wrec, z_kl=l_zkl,
                acc=categorical_accuracy(w_true, W), W=W, W_mean=W_mean, W_log_var=W_log_var,
                Z_mean=Z_mean, Z_log_var=Z_log_var, Z=Zs, h_e=h_e, h_d=h_d, P=P, hW=hW)


def vrnn_loss_and_grads(p, *args, **kw):
    """Gradients by torch.autograd (what TF autodiff computes inside model.fit)."""
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = vrnn_forward(q, *args, **kw)
    out["loss"].backward()
    grads = {k: v.grad.detach() for k, v in q.items()}
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads


# --------------------------------------------------------------------------------------
# CL-VAE training graph (cl_vae/model.py:130-224)
# --------------------------------------------------------------------------------------
def vae_forward(p, x, xp, w_true, eps_w, eps_z, C, use_x_prev, class_weight=1.0, kl_weight=1.0,
                w_kl_weight=1.0, w_log_var_prior=0.0):
    """x [B,D], xp [B,D] or None, eps_w [B,C-1], eps_z [B,Z]."""
    D = x.shape[1]
    h_w = torch.relu(x @ p["h_w.kernel"] + p["h_w.bias"])                         # model.py:141
    w_mean = h_w @ p["w_mean.kernel"] + p["w_mean.bias"]                          # model.py:142
    w_log_var = h_w @ p["w_log_var.kernel"] + p["w_log_var.bias"]                 # model.py:143
    w = logistic_normal(w_mean, w_log_var, eps_w)                                 # model.py:146-157
    xw = torch.cat([x, w], dim=-1)                                                # model.py:160
    h = torch.relu(xw @ p["h.kernel"] + p["h.bias"])                              # model.py:162
    z_mean = h @ p["z_mean.kernel"] + p["z_mean.bias"]
    z_log_var = h @ p["z_log_var.kernel"] + p["z_log_var.bias"]
    z = z_mean + torch.exp(z_log_var / 2) * eps_z                                 # model.py:170-174
    wz = torch.cat([w] + ([xp] if use_x_prev else []) + [z], dim=-1)              # model.py:177-181
    h_dec = torch.relu(wz @ p["decoder_h.kernel"] + p["decoder_h.bias"])          # model.py:184-186
    P = torch.sigmoid(h_dec @ p["x_decoded_mean.kernel"] + p["x_decoded_mean.bias"])
    w2 = w + 1e-10                                                                # model.py:208
    l_vae = vae_loss(x, P, D).mean()
    l_wkl = w_kl_loss(w_mean, w_log_var, w_log_var_prior).mean()
    l_wrec = w_rec_loss(w_true, w2, C).mean()
    l_zkl = z_kl_loss(z_mean, z_log_var).mean()
    total = 1.0 * l_vae + w_kl_weight * l_wkl + class_weight * l_wrec + kl_weight * l_zkl
    return dict(loss=total, vae=l_vae, w_kl=l_wkl, w_rec=l_wrec, z_kl=l_zkl,
                acc=categorical_accuracy(w_true, w), W=w, W_mean=w_mean, W_log_var=w_log_var,
                Z_mean=z_mean, Z_log_var=z_log_var, Z=z, P=P)


def vae_loss_and_grads(p, *args, **kw):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = vae_forward(q, *args, **kw)
    out["loss"].backward()
    grads = {k: v.grad.detach() for k, v in q.items()}
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads


# --------------------------------------------------------------------------------------
# AdamWithWeightnorm (utils/weightnorm.py:75-178)
# --------------------------------------------------------------------------------------
class AdamWN:
    """Eager restatement of AdamWithWeightnorm.get_updates (utils/weightnorm.py:75-143) with
    get_weightnorm_params_and_grads (:146-166) and add_weightnorm_param_updates (:169-178).
    State per >=2-D param: V_scaler (ones), m, v (param-shaped), m_g, v_g (per column);
    per 1-D param: m, v.  lr/betas/eps as init_adam_wn (utils/model_utils.py:52-57)."""

    def __init__(self, params, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.iterations = 0
        self.state = {}
        for k, p in params.items():
            st = dict(m=torch.zeros_like(p), v=torch.zeros_like(p))
            if p.dim() > 1:
                st.update(V_scaler=p.new_ones(p.shape[-1]), m_g=p.new_zeros(p.shape[-1]),
                          v_g=p.new_zeros(p.shape[-1]))
            self.state[k] = st

    def step(self, params, grads):
        t = self.iterations + 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        for k, p in params.items():
            g, st = grads[k], self.state[k]
            if p.dim() > 1:
                V = p / st["V_scaler"]
                V_norm = torch.sqrt((V * V).sum(dim=0))
                g_param = st["V_scaler"] * V_norm
                grad_g = (g * V).sum(dim=0) / V_norm
                grad_V = st["V_scaler"] * (g - (grad_g / V_norm) * V)
                st["m_g"] = self.b1 * st["m_g"] + (1 - self.b1) * grad_g
                st["v_g"] = self.b2 * st["v_g"] + (1 - self.b2) * grad_g ** 2
                new_g = g_param - lr_t * st["m_g"] / (torch.sqrt(st["v_g"]) + self.eps)
                st["m"] = self.b1 * st["m"] + (1 - self.b1) * grad_V
                st["v"] = self.b2 * st["v"] + (1 - self.b2) * grad_V ** 2
                new_V = V - lr_t * st["m"] / (torch.sqrt(st["v"]) + self.eps)
                new_V_norm = torch.sqrt((new_V * new_V).sum(dim=0))
                st["V_scaler"] = new_g / new_V_norm
                params[k] = st["V_scaler"] * new_V
            else:
                st["m"] = self.b1 * st["m"] + (1 - self.b1) * g
                st["v"] = self.b2 * st["v"] + (1 - self.b2) * g ** 2
                params[k] = p - lr_t * st["m"] / (torch.sqrt(st["v"]) + self.eps)
        self.iterations = t
        return params


# --------------------------------------------------------------------------------------
# samplers (cl_vrnn/model.py:9-96, cl_vae/model.py:9-74) driven by explicit noise tapes
# --------------------------------------------------------------------------------------
def vrnn_infer_w(p, x_seed, seq_length, C):
    """Key inference of generate_sample (cl_vrnn/model.py:34-41) with w_sample=False, INCLUDING
    quirk Q2: `ntms = x_seed.shape[1]` is the feature dim, so chunk starts are
    range(0, D, seq_length) and only full chunks count.  x_seed [T_seed, D] -> w [1, C]."""
    ntms = x_seed.shape[1]
    ws = []
    for i in range(0, ntms, seq_length):
        xcs = x_seed[i:i + seq_length]
        if xcs.shape[0] == seq_length:
            hW = torch.relu(xcs.reshape(1, -1) @ p["hW.kernel"] + p["hW.bias"])
            Wargs = hW @ p["Wargs.kernel"] + p["Wargs.bias"]
            w_mean = Wargs[:, :C - 1]
            ws.append(logistic_normal(w_mean, torch.zeros_like(w_mean), torch.zeros_like(w_mean)))
    return torch.cat(ws, dim=0).mean(dim=0, keepdim=True)


def vrnn_generate_sample(p, x_seed, nsteps, w, eps_z, u, use_x_prev, enc_lstm=None):
    """generate_sample, cl_vrnn/model.py:47-60, batch 1, stateful LSTMs, noise from tapes:
    eps_z [T_seed+nsteps, Z] (sample_z :90-96), u [T_seed+nsteps, D] (sample_x :62-63, x = u<=p).
    `enc_lstm` = (kernel, recurrent_kernel, bias) for the z-encoder LSTM; default is the trained
    encoder_h (the reference rebuilds it with fresh random weights -- quirk Q1).
    Returns (Xs[T_seed:], all per-step probabilities P [T_seed+nsteps, D])."""
    if enc_lstm is None:
        enc_lstm = (p["encoder_h.kernel"], p["encoder_h.recurrent_kernel"], p["encoder_h.bias"])
    T_seed, D = x_seed.shape
    H = p["decoder_h.recurrent_kernel"].shape[0]
    dt = w.dtype
    he = torch.zeros(1, H, dtype=dt); ce = torch.zeros(1, H, dtype=dt)
    hd = torch.zeros(1, H, dtype=dt); cd = torch.zeros(1, H, dtype=dt)
    Xs = torch.zeros(T_seed + nsteps, D, dtype=dt)
    Ps = torch.zeros(T_seed + nsteps, D, dtype=dt)
    x_prev = None
    for t in range(T_seed + nsteps):
        if t < T_seed:
            x_prev = x_seed[t][None, :]
        xw = torch.cat([x_prev, w], dim=-1)[:, None, :]
        h_seq, he, ce = lstm_forward(xw, *enc_lstm, h0=he, c0=ce)
        zm = h_seq[:, 0] @ p["Z_mean.kernel"] + p["Z_mean.bias"]
        zv = h_seq[:, 0] @ p["Z_log_var.kernel"] + p["Z_log_var.bias"]
        z_t = zm + torch.exp(zv / 2) * eps_z[t][None, :]
        parts = ([x_prev] if use_x_prev else []) + [z_t, w]
        h_seq, hd, cd = lstm_forward(torch.cat(parts, dim=-1)[:, None, :], p["decoder_h.kernel"],
                                     p["decoder_h.recurrent_kernel"], p["decoder_h.bias"],
                                     h0=hd, c0=cd)
        pr = torch.sigmoid(h_seq[:, 0] @ p["X_decoded_mean.kernel"] + p["X_decoded_mean.bias"])
        x_t = (u[t][None, :] <= pr).to(dt)
        Xs[t] = x_t[0]; Ps[t] = pr[0]
        x_prev = x_t
    return Xs[T_seed:], Ps


def vae_infer_w(p, x_seed):
    """cl_vae/model.py:24-25 with w_sample=False: softmax([w_mean, 0]).  x_seed [D]."""
    h_w = torch.relu(x_seed[None, :] @ p["h_w.kernel"] + p["h_w.bias"])
    w_mean = h_w @ p["w_mean.kernel"] + p["w_mean.bias"]
    return logistic_normal(w_mean, torch.zeros_like(w_mean), torch.zeros_like(w_mean))


def vae_generate_sample(p, x_seed, nsteps, w, eps_z, u, use_x_prev, use_z_prior=False):
    """generate_sample, cl_vae/model.py:9-42.  x_seed [D]; eps_z [nsteps,Z]; u [nsteps,D].
    Note the decoder's x input lags one step further behind (x_prev_t, :40-41)."""
    D = x_seed.shape[0]
    dt = w.dtype
    Xs = torch.zeros(nsteps, D, dtype=dt)
    Ps = torch.zeros(nsteps, D, dtype=dt)
    x_prev = x_seed[None, :]
    x_prev_t = x_prev
    for t in range(nsteps):
        h = torch.relu(torch.cat([x_prev, w], dim=-1) @ p["h.kernel"] + p["h.bias"])
        zm = h @ p["z_mean.kernel"] + p["z_mean.bias"]
        zv = h @ p["z_log_var.kernel"] + p["z_log_var.bias"]
        if use_z_prior:
            zm, zv = 0 * zm, 0 * zv
        z_t = zm + torch.exp(zv / 2) * eps_z[t][None, :]
        zc = torch.cat([w] + ([x_prev_t] if use_x_prev else []) + [z_t], dim=-1)
        hd = torch.relu(zc @ p["decoder_h.kernel"] + p["decoder_h.bias"])
        pr = torch.sigmoid(hd @ p["x_decoded_mean.kernel"] + p["x_decoded_mean.bias"])
        x_t = (u[t][None, :] <= pr).to(dt)
        Xs[t] = x_t[0]; Ps[t] = pr[0]
        x_prev_t = x_prev
        x_prev = x_t
    return Xs, Ps


# --------------------------------------------------------------------------------------
# helpers for tests / bench
# --------------------------------------------------------------------------------------
def synth_rolls(rng, B, L1, D=88, density=0.05, lo=15, hi=75):
    """Synthetic piano-roll windows: iid Bernoulli(density) on keys lo..hi."""
    X = np.zeros((B, L1, D), dtype=np.uint8)
    X[:, :, lo:hi + 1] = rng.random((B, L1, hi + 1 - lo)) < density
    return X


def one_hot(labels, C, dtype=torch.float64):
    w = torch.zeros(len(labels), C, dtype=dtype)
    w[torch.arange(len(labels)), torch.as_tensor(labels, dtype=torch.long)] = 1.0
    return w
