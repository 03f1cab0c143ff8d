"""
Classifying VAE+LSTM (CL-VRNN): host-side mirror of code/cl_vrnn/model.py of the reference.
Same public names and argument meaning -- get_model, load_model, make_w_encoder, make_z_encoder,
make_decoder, generate_sample, sample_x / sample_w / sample_w_discrete / sample_z -- but the graph
runs as hand-written sm_100a kernels behind libclv_b200 (no Keras/TensorFlow, no CPU fallback).
"""
import ctypes as C
import json
import numpy as np
import torch

from .. import _lib
from .._lib import lib, check, ptr
from ..engine import Engine, _stream
from ..keras_like import BaseModel, _binary_u8
from ..devops import dense, as_dev_f32, StatefulLSTM


class CLVRNN(BaseModel):
    """What get_model returns (cl_vrnn/model.py:164-267): outputs X_decoded_mean, W, W2, Z_args with
    losses vae_loss, w_kl_loss, w_rec_loss, kl_loss; metric accuracy on W."""
    output_names = ["X_decoded_mean", "W", "W2", "Z_args"]
    acc_name = "W_acc"
    layer_tensors = {
        "hW": ["hW.kernel", "hW.bias"], "Wargs": ["Wargs.kernel", "Wargs.bias"],
        "encoder_h": ["encoder_h.kernel", "encoder_h.recurrent_kernel", "encoder_h.bias"],
        "Z_mean": ["Z_mean.kernel", "Z_mean.bias"], "Z_log_var": ["Z_log_var.kernel", "Z_log_var.bias"],
        "decoder_h": ["decoder_h.kernel", "decoder_h.recurrent_kernel", "decoder_h.bias"],
        "X_decoded_mean": ["X_decoded_mean.kernel", "X_decoded_mean.bias"],
    }

    def __init__(self, engine, kl_weight, w_kl_weight, margs):
        super().__init__(engine, kl_weight, w_kl_weight)
        self.margs = margs
        # model.layers order of the reference graph [K2-recall: topological, weight-less included]
        xp = margs["use_x_prev"]
        self.all_layer_names = (["current", "flatten_1", "hW", "Wargs", "lambda_1", "lambda_2", "W",
                                 "repeat_vector_1", "concatenate_1", "encoder_h", "Z_mean", "Z_log_var"]
                                + (["history"] if xp else []) + ["lambda_3"]
                                + (["concatenate_2"] if xp else []) + ["repeat_vector_2", "concatenate_3",
                                 "decoder_h", "X_decoded_mean", "W2", "Z_args"])

    def _overlaps(self, x, y=None):
        """True when the two arrays a window is built from are shifted copies of each other (PianoData
        windows): they are then stored once, as [n, L+1, D]; otherwise as [first | second], [n, 2L, D]."""
        e = self.engine
        if e.predict_next:
            if y is None:
                return True
            cur = _binary_u8(x[0] if isinstance(x, (list, tuple)) else x, "current")
            return bool(np.array_equal(cur[:, 1:], _binary_u8(y[0], "target")[:, :-1]))
        if e.use_x_prev:
            return bool(np.array_equal(_binary_u8(x[1], "history")[:, 1:], _binary_u8(x[0], "current")[:, :-1]))
        return True

    def _windows_from_inputs(self, x, y=None, overlap=None):
        """[current, history] ([n,L,D] each) -> uint8 windows.  PianoData windows overlap
        (history[:,1:] == current[:,:-1]) -> [n, L+1, D]; otherwise stored as [history | current].
        --predict_next (cl_vrnn/train.py:55-57): input x = frames 0..L-1, target y[0] = frames 1..L of
        the same L+1 window (or [x | y] when they do not overlap).  `overlap` forces the storage form
        (fit decides it once for the training and the validation split)."""
        e = self.engine
        if overlap is None:
            overlap = self._overlaps(x, y)
        if e.predict_next:
            cur = _binary_u8(x[0] if isinstance(x, (list, tuple)) else x, "current")
            tgt = np.concatenate([cur[:, 1:], cur[:, -1:]], axis=1) if y is None else _binary_u8(y[0], "target")
            if overlap:
                if (e.W, e.x_shift, e.y_shift) != (e.L + 1, 0, 1):
                    e.set_window(e.L + 1, 0, 1)
                return np.ascontiguousarray(np.concatenate([cur[:, :1], tgt], axis=1))
            if (e.W, e.x_shift, e.y_shift) != (2 * e.L, 0, e.L):
                e.set_window(2 * e.L, 0, e.L)
            return np.ascontiguousarray(np.concatenate([cur, tgt], axis=1))
        if e.use_x_prev:
            cur, hist = _binary_u8(x[0], "current"), _binary_u8(x[1], "history")
            if overlap:
                if e.x_shift != 0:
                    e.set_window(e.L + 1, 0)
                return np.ascontiguousarray(np.concatenate([hist[:, :1], cur], axis=1))
            if e.x_shift != e.L:
                e.set_window(2 * e.L, e.L)
            return np.ascontiguousarray(np.concatenate([hist, cur], axis=1))
        cur = _binary_u8(x[0] if isinstance(x, (list, tuple)) else x, "current")
        return np.ascontiguousarray(cur)


def get_model(batch_size, original_dim, intermediate_dim, latent_dim, seq_length, n_classes,
              use_x_prev, optimizer, class_weight=1.0, kl_weight=1.0, dropout=0.0, w_kl_weight=1.0,
              w_log_var_prior=0.0, **engine_kw):
    """cl_vrnn/model.py:164.  kl_weight / w_kl_weight may be floats or keras_like.Variable objects
    (annealed by AnnealLossWeight).  Returns (model, encoder) like the reference; `encoder` is the
    X -> [Z_mean, Z_log_var, W] view of the same weights."""
    if dropout:
        raise NotImplementedError("dropout is never set by the reference CLI (cl_vrnn/train.py:46)")
    predict_next = bool(engine_kw.pop("predict_next", False))
    opt_name = optimizer if isinstance(optimizer, str) else getattr(optimizer, "name", "adam-wn")
    opt_kw = {}
    if not isinstance(optimizer, str):
        opt_kw = dict(lr=optimizer.lr, beta_1=optimizer.beta_1, beta_2=optimizer.beta_2,
                      epsilon=optimizer.epsilon)
    seed = engine_kw.pop("seed", None)
    eng = Engine("vrnn", batch_size, L=seq_length, D=original_dim, H=intermediate_dim, Z=latent_dim,
                 n_classes=n_classes, use_x_prev=use_x_prev, class_weight=float(class_weight),
                 kl_weight=float(kl_weight), w_kl_weight=float(w_kl_weight),
                 w_log_var_prior=float(w_log_var_prior), optimizer=opt_name,
                 seed=np.random.randint(0, 2 ** 31 - 1) if seed is None else seed, predict_next=predict_next,
                 **opt_kw, **engine_kw)
    eng.init_params(np.random.default_rng(np.random.randint(0, 2 ** 31 - 1) if seed is None else seed))
    margs = dict(batch_size=batch_size, original_dim=original_dim, intermediate_dim=intermediate_dim,
                 latent_dim=latent_dim, seq_length=seq_length, n_classes=n_classes,
                 use_x_prev=bool(use_x_prev), class_weight=class_weight)
    model = CLVRNN(eng, kl_weight, w_kl_weight, margs)
    return model, EncoderView(model)


def load_model(model_file, batch_size=None, seq_length=None, optimizer='adam'):
    """cl_vrnn/model.py:269-282: read RUN.json next to RUN.h5, rebuild, load weights."""
    margs = json.load(open(model_file.replace('.h5', '.json')))
    optimizer = margs['optimizer'] if optimizer is None else optimizer
    batch_size = margs['batch_size'] if batch_size is None else batch_size
    seq_length = margs['seq_length'] if seq_length is None else seq_length
    model, enc_model = get_model(batch_size, margs['original_dim'], margs['intermediate_dim'],
                                 margs['latent_dim'], seq_length, margs['n_classes'],
                                 margs['use_x_prev'], optimizer, margs['class_weight'])
    model.load_weights(model_file)
    return model, enc_model, margs


# ---------------------------------------------------------------------- sampler sub-models
class EncoderView:
    """`encoder = Model(X, [Z_mean, Z_log_var, W])` (cl_vrnn/model.py:266): same weights, forward only.
    W is the SAMPLED simplex (the Lambda draws noise on every call, as in the reference)."""
    def __init__(self, model):
        self.model = model

    def predict(self, x, batch_size=None):
        e = self.model.engine
        cur = _binary_u8(x[0] if isinstance(x, (list, tuple)) else x, "current")
        n = cur.shape[0]
        if n % e.B:
            raise ValueError("the graph has a static batch size (%d): got %d samples" % (e.B, n))
        win = cur if not e.use_x_prev else np.concatenate([np.zeros_like(cur[:, :1]), cur], axis=1)
        if e.x_shift != 0:
            e.set_window(e.L + 1 if e.use_x_prev else e.L, 0)
        lab = torch.zeros(e.B, dtype=torch.int32)
        zm, zv, w = [], [], []
        for i in range(0, n, e.B):
            e.stage_windows(torch.from_numpy(np.ascontiguousarray(win[i:i + e.B])), lab)
            e.run(train=False, gen_noise=True)
            za = e.ws_view("Zargs", (e.B, e.L, 2 * e.Z)).cpu().numpy()
            zm.append(za[..., :e.Z].copy()); zv.append(za[..., e.Z:].copy())
            w.append(e.ws_view("W", (e.B, e.C)).cpu().numpy().copy())
        return [np.concatenate(zm), np.concatenate(zv), np.concatenate(w)]


class WEncoder:
    """make_w_encoder (cl_vrnn/model.py:98-114): x [S, seq_length, D] -> [w_mean, w_log_var]."""
    def __init__(self, model, seq_length):
        self.model, self.seq_length = model, seq_length

    def reset_states(self):
        pass

    def wargs_device(self, chunks_u8):
        """chunks_u8: device uint8 [M, L, D] -> device Wargs [M, 2(C-1)] (hW relu Dense, Wargs Dense)."""
        e = self.model.engine
        M, L, D = chunks_u8.shape
        assert L == e.L, "hW's kernel is [seq_length*D, D]: chunks must have the model's seq_length"
        C1 = e.C - 1
        off = (torch.arange(M, dtype=torch.int32, device=e.dev) * L).contiguous()
        hW = torch.empty(M, D, device=e.dev)
        a = _lib.clv_gemm_args(M=M, N=D, K=L * D, A=chunks_u8.data_ptr(), lda=D, a_u8=1, a_kmajor=1,
                               a_off=off.data_ptr(), a_grp=1, Bm=e.view("hW.kernel").data_ptr(), ldb=D,
                               b_nmajor=1, C=hW.data_ptr(), ldc=D, bias=e.view("hW.bias").data_ptr(),
                               relu=1, split_k=1)
        check(lib().clv_gemm(C.byref(a), _stream()), "clv_gemm")
        return dense(hW, e.view("Wargs.kernel"), e.view("Wargs.bias"))

    def predict(self, x):
        e = self.model.engine
        xs = torch.from_numpy(np.ascontiguousarray(_binary_u8(x, "x"))).to(e.dev)
        Wargs = self.wargs_device(xs).cpu().numpy()
        return [Wargs[:, :e.C - 1], Wargs[:, e.C - 1:]]


class _LstmWeights:
    """get_layer('encoder_h') of the z-encoder: get_weights / set_weights on its own LSTM tensors."""
    def __init__(self, owner):
        self.owner, self.name = owner, "encoder_h"

    def get_weights(self):
        return [t.detach().cpu().numpy().copy() for t in self.owner.lstm_tensors()]

    def set_weights(self, ws):
        e = self.owner.model.engine
        shapes = [(e.D + e.C, 4 * e.H), (e.H, 4 * e.H), (4 * e.H,)]
        assert len(ws) == 3 and all(tuple(np.shape(w)) == s for w, s in zip(ws, shapes))
        self.owner.lstm = [torch.tensor(np.asarray(w), dtype=torch.float32, device=e.dev).contiguous() for w in ws]
        self.owner._cell = None


class ZEncoder:
    """make_z_encoder (cl_vrnn/model.py:116-136): [x [S,1,D], w [S,C]] -> [z_mean, z_log_var], stateful.
    QUIRK Q1: the reference builds a FRESH encoder_h LSTM here and copies only the Z heads, so its
    sampling-time encoder LSTM is randomly initialised.  Default here is the documented fix (use the
    trained encoder_h); copy_encoder_weights=False reproduces the reference by drawing fresh
    Keras-default LSTM weights (get_layer('encoder_h').set_weights(...) installs given ones)."""
    def __init__(self, model, copy_encoder_weights=True, rng=None):
        self.model = model
        self.lstm = None
        self._cell = None
        if not copy_encoder_weights:
            e = model.engine
            rng = rng or np.random.default_rng(np.random.randint(0, 2 ** 31 - 1))
            r, c = e.D + e.C, 4 * e.H
            lim = np.sqrt(6.0 / (r + c))
            u, _, vt = np.linalg.svd(rng.standard_normal((e.H, c)), full_matrices=False)
            bias = np.zeros(c, np.float32); bias[e.H:2 * e.H] = 1.0
            self.lstm = [torch.tensor(rng.uniform(-lim, lim, (r, c)), dtype=torch.float32, device=e.dev),
                         torch.tensor(vt if vt.shape == (e.H, c) else u, dtype=torch.float32, device=e.dev).contiguous(),
                         torch.tensor(bias, device=e.dev)]

    def lstm_tensors(self):
        e = self.model.engine
        return self.lstm or [e.view("encoder_h.kernel"), e.view("encoder_h.recurrent_kernel"), e.view("encoder_h.bias")]

    def get_layer(self, name):
        if name == "encoder_h":
            return _LstmWeights(self)
        return self.model.get_layer(name)

    def reset_states(self):
        if self._cell is not None:
            self._cell.reset_states()

    def predict(self, x):
        e = self.model.engine
        xs, w = as_dev_f32(x[0], e.dev), as_dev_f32(x[1], e.dev)
        S, L, _ = xs.shape
        if self._cell is None:
            self._cell = StatefulLSTM(*self.lstm_tensors())
        xw = torch.cat([xs, w.reshape(S, 1, e.C).expand(S, L, e.C)], dim=-1).contiguous()   # concat = layout only
        h = self._cell(xw).reshape(S * L, e.H)
        zm = dense(h, e.view("Z_mean.kernel"), e.view("Z_mean.bias")).reshape(S, L, e.Z)
        zv = dense(h, e.view("Z_log_var.kernel"), e.view("Z_log_var.bias")).reshape(S, L, e.Z)
        return [zm.cpu().numpy(), zv.cpu().numpy()]


class Decoder:
    """make_decoder (cl_vrnn/model.py:138-162): [Z, Xp, W] (or [Z, W]) -> X_decoded_mean, stateful;
    shares decoder_h and X_decoded_mean weights with the trained model."""
    def __init__(self, model, use_x_prev=None):
        self.model = model
        self.use_x_prev = model.engine.use_x_prev if use_x_prev is None else bool(use_x_prev)
        self._cell = None

    def reset_states(self):
        if self._cell is not None:
            self._cell.reset_states()

    def predict(self, x):
        e = self.model.engine
        if self.use_x_prev:
            z, xp, w = (as_dev_f32(t, e.dev) for t in x)
        else:
            z, w = (as_dev_f32(t, e.dev) for t in x)
        S, L, _ = z.shape
        if self._cell is None:
            self._cell = StatefulLSTM(e.view("decoder_h.kernel"), e.view("decoder_h.recurrent_kernel"), e.view("decoder_h.bias"))
        wr = w.reshape(S, 1, e.C).expand(S, L, e.C)
        xin = torch.cat(([xp] if self.use_x_prev else []) + [z, wr], dim=-1).contiguous()      # [Xp | Z | W]
        h = self._cell(xin).reshape(S * L, e.H)
        p = dense(h, e.view("X_decoded_mean.kernel"), e.view("X_decoded_mean.bias"), act=2)
        return p.reshape(S, L, e.D).cpu().numpy()


def make_w_encoder(model, original_dim, n_classes, seq_length=1, batch_size=1):
    return WEncoder(model, seq_length)


def make_z_encoder(model, original_dim, n_classes, latent_dims, seq_length=1, batch_size=1,
                   stateful=True, copy_encoder_weights=True):
    return ZEncoder(model, copy_encoder_weights)


def make_decoder(model, original_dim, intermediate_dim, latent_dim, n_classes, use_x_prev,
                 seq_length=1, batch_size=1, stateful=True):
    return Decoder(model, use_x_prev)


# ---------------------------------------------------------------------- numpy samplers (host)
def sample_x(x_mean):
    return 1.0 * (np.random.rand(*x_mean.squeeze().shape) <= x_mean)


def sample_w_discrete(w):
    wn = np.zeros(w.shape)
    wn[np.random.choice(len(w), p=w / w.sum())] = 1.
    return wn


def sample_w(args, nsamps=1, nrm_samp=False, add_noise=True):
    w_mean, w_log_var = args
    if nsamps == 1:
        eps = np.random.randn(*((1, w_mean.flatten().shape[0])))
    else:
        eps = np.random.randn(*((nsamps,) + w_mean.shape))
    w_norm = w_mean + np.exp(w_log_var / 2) * eps if add_noise else w_mean + 0 * eps
    if nrm_samp:
        return w_norm
    if nsamps == 1:
        w_norm = np.hstack([w_norm, np.zeros((w_norm.shape[0], 1))])
        return np.exp(w_norm) / np.sum(np.exp(w_norm), axis=-1)[:, None]
    w_norm = np.dstack([w_norm, np.zeros(w_norm.shape[:-1] + (1,))])
    return np.exp(w_norm) / np.sum(np.exp(w_norm), axis=-1)[:, :, None]


def sample_z(args, nsamps=1):
    Z_mean, Z_log_var = args
    if nsamps == 1:
        eps = np.random.randn(*Z_mean.squeeze().shape)
    else:
        eps = np.random.randn(*((nsamps,) + Z_mean.squeeze().shape))
    return Z_mean + np.exp(Z_log_var / 2) * eps


# ---------------------------------------------------------------------- generation
def infer_w_device(w_enc_model, seeds_u8, seq_length, w_sample=False):
    """Key inference of generate_sample (cl_vrnn/model.py:34-41) for S songs at once, on the device.
    QUIRK Q2 reproduced: chunk starts are range(0, D, seq_length) (the reference reads the FEATURE
    dimension), and only chunks with seq_length full frames count.  seeds_u8: device [S, T_seed, D]."""
    e = w_enc_model.model.engine
    S, T_seed, D = seeds_u8.shape
    starts = [i for i in range(0, D, seq_length) if i + seq_length <= T_seed]
    if not starts:
        raise ValueError("seed shorter than seq_length: the reference fails here too (np.vstack([]))")
    chunks = torch.stack([seeds_u8[:, i:i + seq_length] for i in starts], dim=1).reshape(-1, seq_length, D).contiguous()
    M = chunks.shape[0]
    Wargs = w_enc_model.wargs_device(chunks)
    C1 = e.C - 1
    # sample_w draws np.random.randn(1, C-1) for EVERY chunk, also with add_noise=False (it multiplies
    # the draw by 0, cl_vrnn/model.py:71-80): consume the stream exactly like the reference does
    draws = np.stack([np.random.randn(1, C1)[0] for _ in range(M)]).astype(np.float32)
    eps = torch.from_numpy(draws if w_sample else 0 * draws).to(e.dev)
    Wc = torch.empty(M, e.C, device=e.dev)
    scratch = torch.zeros(8, device=e.dev)
    labels = torch.zeros(M, dtype=torch.int32, device=e.dev)
    check(lib().clv_logitnormal_fwd(ptr(Wargs), 2 * C1, ptr(eps), ptr(labels), ptr(Wc), ptr(scratch), M,
                                    e.C, 0.0, 0.0, 0, 0, None, _stream()), "clv_logitnormal_fwd")
    w = torch.empty(S, e.C, device=e.dev)
    check(lib().clv_chunk_mean(ptr(Wc), ptr(w), S, len(starts), e.C, _stream()), "clv_chunk_mean")
    return w


def generate_samples(dec_model, w_enc_model, z_enc_model, x_seeds, nsteps, use_x_prev, w_vals=None,
                     seq_length=None, w_sample=False, w_discrete=False, noise=None, seed=0, song0=0,
                     return_probs=False, keep_seed_steps=False):
    """Batched B200 form of generate_sample: S songs in one persistent kernel launch.
    x_seeds [S, T_seed, D] (numpy or device uint8); w_vals [S, C] or None (infer from the seed);
    noise = (eps_z [S,T,Z], u [S,T,D]) tapes or None (in-kernel Philox keyed by seed/song/t).
    Returns uint8 [S, nsteps, D] on the host (and probabilities [S, T, D] if asked)."""
    model = dec_model.model
    e = model.engine
    seeds = x_seeds if torch.is_tensor(x_seeds) else torch.from_numpy(np.ascontiguousarray(_binary_u8(x_seeds, "x_seed")))
    seeds = seeds.to(e.dev).contiguous()
    S, T_seed, D = seeds.shape
    T = T_seed + nsteps
    if w_vals is None:
        w = infer_w_device(w_enc_model, seeds, seq_length or e.L, w_sample)
        if w_discrete:
            wh = w.cpu().numpy().astype(np.float64)
            w = torch.tensor(np.stack([sample_w_discrete(r) for r in wh]), dtype=torch.float32, device=e.dev)
    else:
        w = torch.as_tensor(np.asarray(w_vals), dtype=torch.float32).reshape(S, e.C).to(e.dev).contiguous()
    eps_z = u = None
    if noise is not None:
        eps_z = torch.as_tensor(noise[0], dtype=torch.float32).to(e.dev).contiguous()
        u = torch.as_tensor(noise[1], dtype=torch.float32).to(e.dev).contiguous()
        assert eps_z.shape == (S, T, e.Z) and u.shape == (S, T, D)
    # the rolls leave the GPU bit-packed (11 bytes per 88-key frame, 8x less D2H) and are unpacked on the host
    nb = (D + 7) // 8
    out = torch.empty(S, T, nb, dtype=torch.uint8, device=e.dev)
    probs = torch.empty(S, T, D, device=e.dev) if return_probs else None
    lstm = z_enc_model.lstm or [None, None, None]
    cfg = e.cfg(use_x_prev=use_x_prev)
    check(lib().clv_vrnn_sample_bits(C.byref(cfg), ptr(e.params), ptr(lstm[0]), ptr(lstm[1]), ptr(lstm[2]),
                                     ptr(seeds), T_seed, nsteps, ptr(w), ptr(eps_z), ptr(u), seed, song0, S,
                                     ptr(out), ptr(probs), _stream()), "clv_vrnn_sample_bits")
    packed = (out if keep_seed_steps else out[:, T_seed:]).cpu().numpy()
    res = np.unpackbits(packed, axis=-1, bitorder="little")[..., :D]
    return (res, probs.cpu().numpy()) if return_probs else res


def generate_sample(dec_model, w_enc_model, z_enc_model, x_seed, nsteps, use_x_prev, w_val=None,
                    do_reset=True, seq_length=None, w_sample=False, w_discrete=False):
    """cl_vrnn/model.py:9-60, one song.  The whole `for t` loop runs inside the persistent kernel; the
    np.random stream is consumed draw for draw in the reference's order -- per inferred-key chunk
    randn(1, C-1) (drawn even when w_sample is off), one np.random.choice with w_discrete, then per step
    randn(z) and rand(88) -- so a fixed np.random.seed reproduces the reference's sample (pinned by
    tests/golden/vrnn_sampler.npz, produced by running the reference's own generate_sample).
    A 1-D x_seed is the reference's `nseedsteps == 0` form (:27-31): the seed frame is x_prev of the first
    generated step and all nsteps outputs are returned.  Returns float64 [nsteps, D]."""
    e = dec_model.model.engine
    x_seed = np.asarray(x_seed)
    one_d = x_seed.ndim == 1
    if one_d:
        if w_val is None:
            raise IndexError("a 1-D x_seed needs w_val: the reference reads x_seed.shape[1] to infer the key")
        x_seed = x_seed[None, :]
        nsteps -= 1                      # the seed step's own sampled frame is output 0
    T = x_seed.shape[0] + nsteps
    if w_val is None:
        w = infer_w_device(w_enc_model, torch.from_numpy(_binary_u8(x_seed, "x_seed")[None]).to(e.dev),
                           seq_length or e.L, w_sample).cpu().numpy().astype(np.float64)
        if w_discrete:
            w = sample_w_discrete(w[0])[None, :]
    else:
        w = np.asarray(w_val, dtype=np.float64).reshape(1, -1)
    eps_z = np.zeros((1, T, e.Z), np.float32)
    u = np.zeros((1, T, e.D), np.float32)
    for t in range(T):
        eps_z[0, t] = np.random.randn(e.Z)
        u[0, t] = np.random.rand(e.D)
    xs = generate_samples(dec_model, w_enc_model, z_enc_model, x_seed[None], nsteps, use_x_prev,
                          w_vals=w, noise=(eps_z, u), keep_seed_steps=one_d)
    return xs[0].astype(np.float64)
