"""
Classifying VAE (CL-VAE, frame-wise, all-Dense): host-side mirror of code/cl_vae/model.py.
Same public names -- get_model, load_model, make_w_encoder, make_z_encoder, make_decoder,
generate_sample, sample_x / sample_w / sample_z -- over the sm_100a kernels of libclv_b200.
"""
import ctypes as C
import json
import numpy as np
import torch

from .. import _lib
from .._lib import lib, check, ptr
from ..engine import Engine, _stream
from ..keras_like import BaseModel, _binary_u8
from ..devops import dense, as_dev_f32
from ..cl_vrnn.model import sample_w, sample_z  # identical numpy helpers (cl_vae/model.py:47-74)


def sample_x(x_mean):
    return 1.0 * (np.random.rand(len(x_mean.squeeze())) <= x_mean)


class CLVAE(BaseModel):
    """get_model's return (cl_vae/model.py:130-224): outputs x_decoded_mean, w, w2, z_args."""
    output_names = ["x_decoded_mean", "w", "w2", "z_args"]
    acc_name = "w_acc"
    layer_tensors = {
        "h_w": ["h_w.kernel", "h_w.bias"], "w_mean": ["w_mean.kernel", "w_mean.bias"],
        "w_log_var": ["w_log_var.kernel", "w_log_var.bias"], "h": ["h.kernel", "h.bias"],
        "z_mean": ["z_mean.kernel", "z_mean.bias"], "z_log_var": ["z_log_var.kernel", "z_log_var.bias"],
        "decoder_h": ["decoder_h.kernel", "decoder_h.bias"],
        "x_decoded_mean": ["x_decoded_mean.kernel", "x_decoded_mean.bias"],
    }

    def __init__(self, engine, kl_weight, w_kl_weight, margs):
        super().__init__(engine, kl_weight, w_kl_weight)
        self.margs = margs
        xp = margs["use_x_prev"]
        self.all_layer_names = (["x", "h_w", "w_mean", "w_log_var", "w", "concatenate_1", "h", "z_mean",
                                 "z_log_var"] + (["history"] if xp else []) + ["z"]
                                + (["concatenate_2"] if xp else []) + ["concatenate_3", "decoder_h",
                                 "x_decoded_mean", "w2", "z_args"])

    def _overlaps(self, x, y=None):
        return True            # one storage form only: [history | x] / [x | y] / [x]

    def _windows_from_inputs(self, x, y=None, overlap=None):
        """[x, history] ([n, D] each) -> uint8 windows [n, 2, D] = [history | x]; x alone -> [n,1,D];
        --predict_next (cl_vae/train.py:18,63-69): [x | y] with the target y[0] = the next frame."""
        e = self.engine
        if e.predict_next:
            cur = _binary_u8(x[0] if isinstance(x, (list, tuple)) else x, "x")
            tgt = cur if y is None else _binary_u8(y[0], "target")
            return np.ascontiguousarray(np.stack([cur, tgt], axis=1))
        if e.use_x_prev:
            cur, hist = _binary_u8(x[0], "x"), _binary_u8(x[1], "history")
            return np.ascontiguousarray(np.stack([hist, cur], axis=1))
        cur = _binary_u8(x[0] if isinstance(x, (list, tuple)) else x, "x")
        return np.ascontiguousarray(cur[:, None, :])


def get_model(batch_size, original_dim, latent_dims, class_dims, optimizer, class_weight=1.0,
              kl_weight=1.0, use_x_prev=False, w_kl_weight=1.0, w_log_var_prior=0.0, **engine_kw):
    """cl_vae/model.py:130-134: latent_dims = (latent_dim_0, latent_dim) i.e. (intermediate_dim,
    latent_dim); class_dims = (class_dim_0, class_dim) i.e. (intermediate_class_dim, n_classes)."""
    latent_dim_0, latent_dim = latent_dims
    class_dim_0, class_dim = class_dims
    if latent_dim_0 <= 0:
        raise NotImplementedError("intermediate_dim=0 (no hidden layers) is not built")
    opt_name = optimizer if isinstance(optimizer, str) else getattr(optimizer, "name", "adam-wn")
    opt_kw = {}
    if not isinstance(optimizer, str):
        opt_kw = dict(lr=optimizer.lr, beta_1=optimizer.beta_1, beta_2=optimizer.beta_2,
                      epsilon=optimizer.epsilon)
    seed = engine_kw.pop("seed", None)
    seed = np.random.randint(0, 2 ** 31 - 1) if seed is None else seed
    engine_kw["predict_next"] = bool(engine_kw.pop("predict_next", False))
    eng = Engine("vae", batch_size, L=1, D=int(original_dim), H=latent_dim_0, Z=latent_dim,
                 n_classes=class_dim, use_x_prev=use_x_prev, Hc=class_dim_0,
                 class_weight=float(class_weight), kl_weight=float(kl_weight),
                 w_kl_weight=float(w_kl_weight), w_log_var_prior=float(w_log_var_prior),
                 optimizer=opt_name, seed=seed, **opt_kw, **engine_kw)
    eng.init_params(np.random.default_rng(seed))
    margs = dict(batch_size=batch_size, original_dim=int(original_dim), intermediate_dim=latent_dim_0,
                 latent_dim=latent_dim, intermediate_class_dim=class_dim_0, n_classes=class_dim,
                 use_x_prev=bool(use_x_prev), class_weight=class_weight)
    model = CLVAE(eng, kl_weight, w_kl_weight, margs)
    return model, EncModel(model)


def load_model(model_file, optimizer='adam', batch_size=1, no_x_prev=False):
    """cl_vae/model.py:226-239."""
    margs = json.load(open(model_file.replace('.h5', '.json')))
    batch_size = margs['batch_size'] if batch_size is None else batch_size
    if no_x_prev or 'use_x_prev' not in margs:
        margs['use_x_prev'] = False
    model, enc_model = get_model(batch_size, margs['original_dim'],
                                 (margs['intermediate_dim'], margs['latent_dim']),
                                 (margs['intermediate_class_dim'], margs['n_classes']), optimizer,
                                 margs['class_weight'], use_x_prev=margs['use_x_prev'])
    model.load_weights(model_file)
    return model, enc_model, margs


class EncModel:
    """`enc_model = Model([x, xp], [z_mean, w_mean])` (cl_vae/model.py:213-222): forward only, same weights
    (w inside the graph is the sampled simplex, so z_mean depends on this call's noise draw)."""
    def __init__(self, model):
        self.model = model

    def predict(self, x, batch_size=None):
        e = self.model.engine
        win = self.model._windows_from_inputs(x)
        n = win.shape[0]
        if n % e.B:
            raise ValueError("the graph has a static batch size (%d): got %d samples" % (e.B, n))
        lab = torch.zeros(e.B, dtype=torch.int32)
        zm, wm = [], []
        for i in range(0, n, e.B):
            e.stage_windows(torch.from_numpy(np.ascontiguousarray(win[i:i + e.B])), lab)
            e.run(train=False, gen_noise=True)
            zm.append(e.ws_view("Zargs", (e.B, 2 * e.Z)).cpu().numpy()[:, :e.Z].copy())
            wm.append(e.ws_view("Wargs", (e.B, 2 * (e.C - 1))).cpu().numpy()[:, :e.C - 1].copy())
        return [np.concatenate(zm), np.concatenate(wm)]


class _Sub:
    def __init__(self, model):
        self.model = model

    def reset_states(self):
        pass


class WEncoder(_Sub):
    """make_w_encoder (cl_vae/model.py:76-85): x [S, D] -> [w_mean, w_log_var]."""
    def predict(self, x):
        e = self.model.engine
        h_w = dense(as_dev_f32(x, e.dev), e.view("h_w.kernel"), e.view("h_w.bias"), act=1)
        return [dense(h_w, e.view("w_mean.kernel"), e.view("w_mean.bias")).cpu().numpy(),
                dense(h_w, e.view("w_log_var.kernel"), e.view("w_log_var.bias")).cpu().numpy()]


class ZEncoder(_Sub):
    """make_z_encoder (cl_vae/model.py:87-102): [x [S,D], w [S,C]] -> [z_mean, z_log_var]."""
    def predict(self, x):
        e = self.model.engine
        xw = torch.cat([as_dev_f32(x[0], e.dev), as_dev_f32(x[1], e.dev)], dim=-1).contiguous()
        h = dense(xw, e.view("h.kernel"), e.view("h.bias"), act=1)
        return [dense(h, e.view("z_mean.kernel"), e.view("z_mean.bias")).cpu().numpy(),
                dense(h, e.view("z_log_var.kernel"), e.view("z_log_var.bias")).cpu().numpy()]


class Decoder(_Sub):
    """make_decoder (cl_vae/model.py:104-128): inputs [w, z, xp] (or [w, z]); the Dense sees [w | xp | z]."""
    def __init__(self, model, use_x_prev):
        super().__init__(model)
        self.use_x_prev = bool(use_x_prev)

    def predict(self, x):
        e = self.model.engine
        w, z = as_dev_f32(x[0], e.dev), as_dev_f32(x[1], e.dev)
        parts = [w] + ([as_dev_f32(x[2], e.dev)] if self.use_x_prev else []) + [z]
        hd = dense(torch.cat(parts, dim=-1).contiguous(), e.view("decoder_h.kernel"), e.view("decoder_h.bias"), act=1)
        return dense(hd, e.view("x_decoded_mean.kernel"), e.view("x_decoded_mean.bias"), act=2).cpu().numpy()


def make_w_encoder(model, original_dim, batch_size=1):
    return WEncoder(model)


def make_z_encoder(model, original_dim, class_dim, latent_dims, batch_size=1):
    return ZEncoder(model)


def make_decoder(model, latent_dims, class_dim, original_dim=88, use_x_prev=False, batch_size=1):
    return Decoder(model, use_x_prev)


def infer_w_device(model, x_seeds_u8, w_sample=False):
    """w_t = sample_w(w_enc_model.predict(x_prev), add_noise=w_sample) (cl_vae/model.py:24-25) for S
    frames at once: h_w relu Dense, the two heads, softmax([w_mean (+noise), 0])."""
    e = model.engine
    S, D = x_seeds_u8.shape
    C1 = e.C - 1
    off = torch.arange(S, dtype=torch.int32, device=e.dev)
    h_w = torch.empty(S, e.Hc, device=e.dev)
    Wargs = torch.empty(S, 2 * C1, device=e.dev)
    a = _lib.clv_gemm_args(M=S, N=e.Hc, K=D, A=x_seeds_u8.data_ptr(), lda=D, a_u8=1, a_kmajor=1,
                           a_off=off.data_ptr(), a_grp=1, Bm=e.view("h_w.kernel").data_ptr(), ldb=e.Hc,
                           b_nmajor=1, C=h_w.data_ptr(), ldc=e.Hc, bias=e.view("h_w.bias").data_ptr(),
                           relu=1, split_k=1)
    check(lib().clv_gemm(C.byref(a), _stream()), "clv_gemm")
    for j, nm in enumerate(("w_mean", "w_log_var")):
        a = _lib.clv_gemm_args(M=S, N=C1, K=e.Hc, A=h_w.data_ptr(), lda=e.Hc, a_kmajor=1,
                               Bm=e.view(nm + ".kernel").data_ptr(), ldb=C1, b_nmajor=1,
                               C=Wargs.data_ptr() + 4 * j * C1, ldc=2 * C1,
                               bias=e.view(nm + ".bias").data_ptr(), split_k=1)
        check(lib().clv_gemm(C.byref(a), _stream()), "clv_gemm")
    # sample_w draws np.random.randn(1, C-1) per frame also with add_noise=False (cl_vae/model.py:47-58)
    draws = np.stack([np.random.randn(1, C1)[0] for _ in range(S)]).astype(np.float32)
    eps = torch.from_numpy(draws if w_sample else 0 * draws).to(e.dev)
    W = torch.empty(S, e.C, device=e.dev)
    scratch = torch.zeros(8, device=e.dev)
    labels = torch.zeros(S, dtype=torch.int32, device=e.dev)
    check(lib().clv_logitnormal_fwd(ptr(Wargs), 2 * C1, ptr(eps), ptr(labels), ptr(W), ptr(scratch), S,
                                    e.C, 0.0, 0.0, 0, 0, None, _stream()), "clv_logitnormal_fwd")
    return W


def generate_samples(model, x_seeds, nsteps, w_vals=None, use_z_prior=False, w_sample=False,
                     use_x_prev=None, noise=None, seed=0, song0=0, return_probs=False):
    """Batched persistent-kernel form of generate_sample: x_seeds [S, D] -> uint8 [S, nsteps, D]."""
    e = model.engine
    seeds = x_seeds if torch.is_tensor(x_seeds) else torch.from_numpy(np.ascontiguousarray(_binary_u8(x_seeds, "x_seed")))
    seeds = seeds.to(e.dev).contiguous()
    S, D = seeds.shape
    if w_vals is None:
        w = infer_w_device(model, seeds, w_sample)
    else:
        w = torch.as_tensor(np.asarray(w_vals), dtype=torch.float32).reshape(S, e.C).to(e.dev).contiguous()
    eps_z = u = None
    if noise is not None:
        eps_z = torch.as_tensor(noise[0], dtype=torch.float32).to(e.dev).contiguous()
        u = torch.as_tensor(noise[1], dtype=torch.float32).to(e.dev).contiguous()
    out = torch.empty(S, nsteps, D, dtype=torch.uint8, device=e.dev)
    probs = torch.empty(S, nsteps, D, device=e.dev) if return_probs else None
    cfg = e.cfg(use_x_prev=e.use_x_prev if use_x_prev is None else use_x_prev)
    check(lib().clv_vae_sample(C.byref(cfg), ptr(e.params), ptr(seeds), nsteps, ptr(w), ptr(eps_z), ptr(u),
                               seed, song0, S, int(bool(use_z_prior)), ptr(out), ptr(probs), _stream()),
          "clv_vae_sample")
    res = out.cpu().numpy()
    return (res, probs.cpu().numpy()) if return_probs else res


def generate_sample(dec_model, w_enc_model, z_enc_model, x_seed, nsteps, w_val=None, use_z_prior=False,
                    do_reset=True, w_sample=False, use_x_prev=False):
    """cl_vae/model.py:9-42, one song; the loop runs in the persistent kernel, noise drawn from
    np.random in the reference's per-step order (randn(z) then rand(D)).  Returns float64 [nsteps, D]."""
    model = dec_model.model
    e = model.engine
    x_seed = _binary_u8(np.asarray(x_seed).reshape(1, -1), "x_seed")
    if w_val is None:
        w = infer_w_device(model, torch.from_numpy(x_seed).to(e.dev), w_sample).cpu().numpy()
    else:
        w = np.asarray(w_val, dtype=np.float64).reshape(1, -1)
    eps_z = np.zeros((1, nsteps, e.Z), np.float32)
    u = np.zeros((1, nsteps, e.D), np.float32)
    for t in range(nsteps):
        eps_z[0, t] = np.random.randn(e.Z)
        u[0, t] = np.random.rand(e.D)
    xs = generate_samples(model, x_seed, nsteps, w_vals=w, use_z_prior=use_z_prior,
                          use_x_prev=use_x_prev, noise=(eps_z, u))
    return xs[0].astype(np.float64)
