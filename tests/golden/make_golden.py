"""
Generates tests/golden/*.json|npz by executing UNMODIFIED reference code from /root/reference in this
container (it cannot travel to the GPU box, hence the committed fixtures):

  * utils/pianoroll.py (PianoData) -- importable under Python 3 once `cPickle` and `xrange` are
    shimmed; run on the two bundled JSB pickles for the three configurations the CLIs use.
  * utils/weightnorm.py -- AdamWithWeightnorm.get_updates / get_weightnorm_params_and_grads /
    add_weightnorm_param_updates executed against a small numpy shim of the keras / tensorflow API
    (eager variables, K.update collected then applied simultaneously like one session.run).

Everything else on the hot path lives inside Keras 2.0.0 / TF 1.0.1, which are not installable here,
so it stays "parity unpinned" (see oracle/clv_oracle.py header).

Run:  python tests/golden/make_golden.py        (needs /root/reference; writes next to this file)
"""
import builtins, hashlib, zlib, importlib, json, os, pickle, sys, types
import numpy as np

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ shims
def install_py2_shims():
    cp = types.ModuleType("cPickle")

    def load(f):
        name = f.name
        f.close()
        with open(name, "rb") as fb:
            return pickle.load(fb, encoding="latin1")
    cp.load = load
    sys.modules["cPickle"] = cp
    builtins.xrange = range


class Var(np.ndarray):
    """numpy-backed stand-in for a Keras variable / TF tensor (hashable, like TF tensors)."""
    def __hash__(self):
        return id(self)


def as_var(a):
    return np.array(a, dtype=np.float64).view(Var)


class KerasShim:
    def __init__(self):
        self.created = []
        self.cursor = 0
        self.updates = []

    def _new(self, arr):
        if self.cursor < len(self.created):
            v = self.created[self.cursor]
        else:
            v = as_var(arr)
            self.created.append(v)
        self.cursor += 1
        return v

    def install(self):
        K = types.ModuleType("keras.backend")
        K.get_variable_shape = lambda p: tuple(p.shape)
        K.zeros = lambda shape: self._new(np.zeros(shape))
        K.ones = lambda shape: self._new(np.ones(shape))
        K.sqrt = np.sqrt
        K.pow = np.power
        K.square = np.square
        K.update = lambda var, val: (var, np.array(val))
        K.update_add = lambda var, inc: (var, np.array(var + inc))
        keras = types.ModuleType("keras")
        keras.backend = K
        opt = types.ModuleType("keras.optimizers")

        class _Opt(object):
            def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-8, decay=0.0, **kw):
                self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
                self.decay = self.initial_decay = decay
                self.iterations = as_var(0.0)
                self.momentum, self.nesterov = kw.get("momentum", 0.0), kw.get("nesterov", False)
                self._grads = None

            def get_gradients(self, loss, params):
                return self._grads
        opt.Adam = _Opt
        opt.SGD = _Opt
        keras.optimizers = opt
        tf = types.ModuleType("tensorflow")
        tf.reshape = lambda x, shape: np.reshape(x, shape)
        tf.sqrt = np.sqrt
        tf.square = np.square
        tf.reduce_sum = lambda x, axes: np.sum(x, axis=tuple(axes))
        sys.modules.update({"keras": keras, "keras.backend": K, "keras.optimizers": opt,
                            "tensorflow": tf})


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------ PianoData golden
def golden_pianodata():
    install_py2_shims()
    sys.path.insert(0, REF)
    pr = importlib.import_module("utils.pianoroll")
    out = {}
    cfgs = {
        "vrnn_train": dict(batch_size=200, seq_length=16, step_length=1, return_y_next=True,
                           return_y_hist=True, squeeze_x=False, squeeze_y=False),
        "vae_train": dict(batch_size=100, seq_length=1, step_length=1, return_y_next=True,
                          squeeze_x=True, squeeze_y=True),
        "vrnn_sample": dict(batch_size=1, seq_length=32, squeeze_x=False),
        "vae_sample": dict(batch_size=1, seq_length=32, squeeze_x=True),
    }
    for fn in ("JSB Chorales_all.pickle", "JSB Chorales_Cs.pickle"):
        for cname, kw in cfgs.items():
            P = pr.PianoData(os.path.join("/root/reference/data/input", fn), **kw)
            rec = {"key_map": {str(k): int(v) for k, v in P.key_map.items()}}
            for split in ("train", "valid", "test"):
                x = getattr(P, "x_" + split); y = getattr(P, "y_" + split)
                keys = getattr(P, split + "_song_keys"); inds = getattr(P, split + "_song_inds")
                rec[split] = dict(
                    x_shape=list(x.shape), y_shape=list(y.shape),
                    x_sum=float(x.sum()), y_sum=float(y.sum()),
                    x_sha=sha(x.astype(np.uint8)), y_sha=sha(y.astype(np.uint8)),
                    keys_sha=sha(keys.astype(np.int64)), inds_sha=sha(inds.astype(np.int64)),
                    keys_head=[int(k) for k in keys[:8]], keys_tail=[int(k) for k in keys[-8:]],
                    n_unique_keys=int(len(np.unique(keys))),
                    modes_sha=sha(np.asarray(getattr(P, split + "_song_modes")).astype(np.uint8)))
            out[fn + "::" + cname] = rec
    with open(os.path.join(HERE, "pianodata.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("pianodata.json:", len(out), "records")


# ------------------------------------------------------------------ Adam-WN golden
def golden_adamwn():
    shim = KerasShim()
    shim.install()
    sys.path.insert(0, REF)
    wn = importlib.import_module("utils.weightnorm")
    rng = np.random.default_rng(20171107)
    shapes = [(24, 40), (11, 6), (40,), (9, 2), (2,)]
    params = [as_var(rng.normal(0, 0.3, s)) for s in shapes]
    opt = wn.AdamWithWeightnorm(lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-08, decay=0.0)
    n_steps = 4
    grads_all = [[rng.normal(0, 1.0, s) * (10.0 ** rng.integers(-3, 1)) for s in shapes]
                 for _ in range(n_steps)]
    rec = {"p0_%d" % i: np.array(p) for i, p in enumerate(params)}
    for step in range(n_steps):
        opt._grads = [as_var(g) for g in grads_all[step]]
        shim.cursor = 0
        updates = opt.get_updates(params, {}, None)
        for var, val in updates:          # one session.run: all reads precede all assigns
            var[...] = val
        for i, p in enumerate(params):
            rec["g%d_%d" % (step, i)] = grads_all[step][i]
            rec["p%d_%d" % (step + 1, i)] = np.array(p)
    rec["iterations"] = np.array(opt.iterations)
    # direct helper golden: get_weightnorm_params_and_grads on a fresh matrix
    shim.cursor = len(shim.created)
    p = as_var(rng.normal(0, 1, (7, 5))); g = as_var(rng.normal(0, 1, (7, 5)))
    V, V_norm, V_scaler, g_param, grad_g, grad_V = wn.get_weightnorm_params_and_grads(p, g)
    rec.update(h_p=np.array(p), h_g=np.array(g), h_V=np.array(V), h_V_norm=np.array(V_norm),
               h_g_param=np.array(g_param), h_grad_g=np.array(grad_g), h_grad_V=np.array(grad_V))
    np.savez_compressed(os.path.join(HERE, "adamwn.npz"), **rec)
    print("adamwn.npz:", len(rec), "arrays, iterations =", float(opt.iterations))




# ------------------------------------------------------------------ model graphs + samplers golden
def py2_to_py3_source(src):
    """Mechanical Python-2 -> 3 rewrite of the ONLY py2-isms in cl_*/model.py: tuple parameters
    (`def f(a, (b, c), d)`, what lib2to3's fix_tuple_params does) -- `xrange` is shimmed as a builtin.
    Statements, names and order are untouched."""
    import re
    out, pos = [], 0
    for m in re.finditer(r"^def\s+\w+\(", src, flags=re.M):
        i = m.end(); depth = 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0); i += 1
        head_end = src.index(":", i) + 1
        params = src[m.end():i - 1]
        tup = re.findall(r"\(([^()]*)\)", params)
        if not tup:
            continue
        unpack = []
        for k, t in enumerate(tup):
            params = params.replace("(" + t + ")", "_tup%d" % k, 1)
            unpack.append("    (%s) = _tup%d" % (t, k))
        out.append(src[pos:m.end()] + params + src[i - 1:head_end] + "\n" + "\n".join(unpack))
        pos = head_end
    out.append(src[pos:])
    return "".join(out)


def load_reference_model_module(sub):
    """exec /root/reference/code/<sub>/model.py (its own statements) against tests/golden/keras_shim.py"""
    sys.path.insert(0, HERE)
    import keras_shim
    keras_shim.install()
    builtins.xrange = range
    src = py2_to_py3_source(open(os.path.join(REF, sub, "model.py")).read())
    mod = types.ModuleType("ref_%s_model" % sub)
    mod.__file__ = os.path.join(REF, sub, "model.py")
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod, keras_shim


def _randomise(model, seed, scale=1.0, prefix="", only=None):
    """weights from tests/util.golden_weight(seed, 'layer.weight'): the tests regenerate them"""
    for d in (os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))):
        if d not in sys.path:
            sys.path.insert(0, d)
    import util
    shas = {}
    for lay in model.layers:
        if lay.weight_names and lay.weights is not None and (only is None or lay.name in only):
            ws = []
            for n, w in zip(lay.weight_names, lay.get_weights()):
                v = util.golden_weight(seed, "%s%s.%s" % (prefix, lay.name, n), w.shape, scale)
                shas["%s%s.%s" % (prefix, lay.name, n)] = util.sha256(v)
                ws.append(v.astype(np.float64))
            lay.set_weights(ws)
    return shas


def _rolls(rng, *shape):
    x = np.zeros(shape, np.uint8)
    x[..., 15:76] = rng.random(shape[:-1] + (61,)) < 0.12
    return x


def _wl(model):
    return [l.name for l in model.layers if l.weight_names and l.weights is not None]


def golden_vrnn_train():
    """cl_vrnn/model.py get_model executed; fed like cl_vrnn/train.py:51-63: inputs [current, history]
    (or x alone), targets [y, w, w, y]; `predict_next`: input = frames 0..L-1, target = frames 1..L."""
    M, ks = load_reference_model_module("cl_vrnn")
    rec = {}
    cases = {
        "xprev": dict(B=6, L=4, C=4, Z=2, xp=True, pn=False,
                      kw=dict(class_weight=0.7, kl_weight=ks.K.variable(0.4), w_kl_weight=0.3, w_log_var_prior=0.2)),
        "noxprev": dict(B=5, L=3, C=3, Z=2, xp=False, pn=False, kw=dict()),
        "predict_next": dict(B=4, L=3, C=10, Z=3, xp=False, pn=True, kw=dict(w_kl_weight=ks.K.variable(0.0))),
    }
    for ci, (name, c) in enumerate(cases.items()):
        rng = np.random.default_rng(100 + ci)
        ks.set_rng(1)
        B, L, C, Z, D, H = c["B"], c["L"], c["C"], c["Z"], 88, 88
        n0 = len(ks.K._noise_nodes)
        model, enc = M.get_model(B, D, H, Z, L, C, c["xp"], "adam", **c["kw"])
        assert len(ks.K._noise_nodes) - n0 == 2
        wseed, wscale = 1000 + ci, 1.0
        shas = _randomise(model, wseed, wscale)
        win = _rolls(rng, B, L + 1, D) if (c["xp"] or c["pn"]) else _rolls(rng, B, L, D)
        labels = rng.integers(0, C, B)
        w_true = ks.to_categorical(labels, C)
        eps_w = rng.standard_normal((B, C - 1)).astype(np.float32)
        eps_z = rng.standard_normal((B, L, Z)).astype(np.float32)
        if c["xp"]:       # x,y = [P.y_train, P.x_train], P.y_train
            x, y = [win[:, 1:], win[:, :-1]], win[:, 1:]
        elif c["pn"]:     # x,y = P.x_train, P.y_train with return_y_next
            x, y = win[:, :-1], win[:, 1:]
        else:             # y == x
            x, y = win, win
        res, grads, outs = model.loss_and_grads(x, [y, w_true, w_true, y], [eps_w, eps_z])
        kwv = {k: (v.value if isinstance(v, ks.Variable) else v) for k, v in c["kw"].items()}
        rec[name + "/cfg"] = np.array(json.dumps(dict(
            B=B, L=L, C=C, Z=Z, D=D, H=H, use_x_prev=c["xp"], predict_next=c["pn"], kw=kwv, wseed=wseed, wscale=wscale,
            wsha=shas, output_names=model.output_names, weighted_layers=_wl(model), res=res)))
        rec[name + "/win"] = win; rec[name + "/labels"] = labels.astype(np.int32)
        rec[name + "/eps_w"] = eps_w; rec[name + "/eps_z"] = eps_z
        for k, v in grads.items():
            rec[name + "/g/" + k] = v.astype(np.float32)
        for k in ("W", "Z_args"):
            rec[name + "/out/" + k] = outs[k].astype(np.float32)
        print("vrnn_train", name, res)
    np.savez_compressed(os.path.join(HERE, "vrnn_train.npz"), **rec)


def golden_vae_train():
    """cl_vae/model.py get_model executed; fed like cl_vae/train.py ([x, history] / [x, w, w, x])."""
    M, ks = load_reference_model_module("cl_vae")
    rec = {}
    cases = {
        "xprev": dict(B=7, C=2, Z=4, H=88, Hc=88, xp=True, kw=dict(class_weight=0.6, kl_weight=0.5, w_kl_weight=0.8, w_log_var_prior=0.3)),
        "noxprev": dict(B=5, C=5, Z=3, H=64, Hc=40, xp=False, kw=dict()),
    }
    for ci, (name, c) in enumerate(cases.items()):
        rng = np.random.default_rng(200 + ci)
        ks.set_rng(2)
        B, C, Z, H, Hc, D = c["B"], c["C"], c["Z"], c["H"], c["Hc"], 88
        model, enc = M.get_model(B, D, (H, Z), (Hc, C), "adam", use_x_prev=c["xp"], **c["kw"])
        wseed, wscale = 2000 + ci, 1.5
        shas = _randomise(model, wseed, wscale)
        win = _rolls(rng, B, 2, D)
        x, xp = win[:, 1], win[:, 0]
        labels = rng.integers(0, C, B)
        w_true = ks.to_categorical(labels, C)
        eps_w = rng.standard_normal((B, C - 1)).astype(np.float32)
        eps_z = rng.standard_normal((B, Z)).astype(np.float32)
        res, grads, outs = model.loss_and_grads([x, xp] if c["xp"] else x, [x, w_true, w_true, x], [eps_w, eps_z])
        rec[name + "/cfg"] = np.array(json.dumps(dict(
            B=B, C=C, Z=Z, D=D, H=H, Hc=Hc, use_x_prev=c["xp"], kw=c["kw"], wseed=wseed, wscale=wscale, wsha=shas,
            output_names=model.output_names, weighted_layers=_wl(model), res=res)))
        rec[name + "/win"] = win; rec[name + "/labels"] = labels.astype(np.int32)
        rec[name + "/eps_w"] = eps_w; rec[name + "/eps_z"] = eps_z
        for k, v in grads.items():
            rec[name + "/g/" + k] = v.astype(np.float32)
        for k in ("w", "z_args"):
            rec[name + "/out/" + k] = outs[k].astype(np.float32)
        print("vae_train", name, res)
    np.savez_compressed(os.path.join(HERE, "vae_train.npz"), **rec)


class _Tape:
    """wraps a sub-model's predict to record what the reference's loop got back from it"""
    def __init__(self, mdl):
        self.mdl, self.calls = mdl, []

    def predict(self, x, **kw):
        y = self.mdl.predict(x, **kw)
        self.calls.append(y)
        return y

    def reset_states(self):
        self.mdl.reset_states()


def _margin(probs, np_seed, draws):
    """min |p - u| over the run, u re-drawn from the same np.random stream in the reference's order
    (`draws`: list of ('randn', n) / ('rand', n) / ('choice', C) in call order)"""
    np.random.seed(np_seed)
    us = []
    for kind, n in draws:
        if kind == "randn":
            np.random.randn(n)
        elif kind == "choice":
            np.random.choice(n, p=np.ones(n) / n)
        else:
            us.append(np.random.rand(n))
    return float(np.abs(np.stack(us) - probs).min())


def golden_vrnn_sampler():
    """cl_vrnn/model.py generate_sample + make_* executed under a fixed np.random.seed, wired like
    cl_vrnn/sample.py:30-40 (batch-1 stateful sub-models, FRESH encoder LSTM = quirk Q1)."""
    M, ks = load_reference_model_module("cl_vrnn")
    rec = {}
    cases = {
        "given": dict(L=4, C=4, Z=2, xp=True, T_seed=6, nsteps=9, infer=False, discrete=False, np_seed=11),
        "infer": dict(L=4, C=4, Z=2, xp=True, T_seed=21, nsteps=7, infer=True, discrete=False, np_seed=12),
        "infer_discrete": dict(L=5, C=3, Z=2, xp=True, T_seed=12, nsteps=6, infer=True, discrete=True, np_seed=13),
        "given_noxprev": dict(L=3, C=3, Z=3, xp=False, T_seed=4, nsteps=8, infer=False, discrete=False, np_seed=14),
        "seed1d": dict(L=4, C=4, Z=2, xp=True, T_seed=0, nsteps=6, infer=False, discrete=False, np_seed=15),
    }
    for ci, (name, c) in enumerate(cases.items()):
        rng = np.random.default_rng(300 + ci)
        ks.set_rng(3)
        L, C, Z, D, H = c["L"], c["C"], c["Z"], 88, 88
        model, _ = M.get_model(8, D, H, Z, L, C, c["xp"], "adam")
        wseed, wscale = 3000 + ci, 1.6
        shas = _randomise(model, wseed, wscale)
        w_enc = M.make_w_encoder(model, D, C, L)
        z_enc = M.make_z_encoder(model, D, C, (H, Z))
        # the FRESH encoder LSTM make_z_encoder builds (quirk Q1): give it known weights too
        shas.update(_randomise(z_enc, wseed, wscale, prefix="zenc/", only=("encoder_h",)))
        dec = M.make_decoder(model, D, H, Z, C, c["xp"])
        x_seed = _rolls(rng, max(c["T_seed"], 1), D).astype(np.float64)
        if c["T_seed"] == 0:
            x_seed = x_seed[0]
        label = int(rng.integers(0, C))
        w_val = None if c["infer"] else ks.to_categorical(label, C)
        tw, tz, td = _Tape(w_enc), _Tape(z_enc), _Tape(dec)
        np.random.seed(c["np_seed"])
        out = M.generate_sample(td, tw, tz, x_seed, c["nsteps"], c["xp"], w_val=w_val, w_discrete=c["discrete"], seq_length=L)
        after = np.random.rand()          # position of the np.random stream after the call
        probs = np.stack([np.asarray(p).reshape(-1) for p in td.calls]).astype(np.float32)
        draws = [("randn", C - 1)] * len(tw.calls) + ([("choice", C)] if c["discrete"] else [])
        for _ in range(len(td.calls)):
            draws += [("randn", Z), ("rand", D)]
        margin = _margin(probs, c["np_seed"], draws)
        assert abs(np.random.rand() - after) < 1e-15, "draw order of the re-derivation differs from the reference run"
        assert margin > 1e-5, (name, margin)
        rec[name + "/cfg"] = np.array(json.dumps(dict(
            L=L, C=C, Z=Z, D=D, H=H, use_x_prev=c["xp"], T_seed=c["T_seed"], nsteps=c["nsteps"], infer=c["infer"],
            discrete=c["discrete"], np_seed=c["np_seed"], label=label, n_w_calls=len(tw.calls), stream_after=after,
            min_margin=margin, wseed=wseed, wscale=wscale, wsha=shas)))
        rec[name + "/x_seed"] = x_seed.astype(np.uint8)
        rec[name + "/out"] = out.astype(np.uint8)
        rec[name + "/probs"] = probs
        rec[name + "/zargs"] = np.stack([np.concatenate([np.asarray(a).reshape(-1) for a in p]) for p in tz.calls]).astype(np.float32)
        print("vrnn_sampler", name, "notes", int(out.sum()), "of", out.size, "w calls", len(tw.calls), "margin %.2e" % margin)
    np.savez_compressed(os.path.join(HERE, "vrnn_sampler.npz"), **rec)


def golden_vae_sampler():
    """cl_vae/model.py generate_sample + make_* executed, wired like cl_vae/sample.py:17-22."""
    M, ks = load_reference_model_module("cl_vae")
    rec = {}
    cases = {
        "given": dict(C=2, Z=4, xp=True, nsteps=10, infer=False, zprior=False, np_seed=21),
        "infer": dict(C=3, Z=4, xp=True, nsteps=9, infer=True, zprior=False, np_seed=22),
        "zprior_noxprev": dict(C=2, Z=3, xp=False, nsteps=8, infer=True, zprior=True, np_seed=23),
    }
    for ci, (name, c) in enumerate(cases.items()):
        rng = np.random.default_rng(400 + ci)
        ks.set_rng(4)
        C, Z, D, H, Hc = c["C"], c["Z"], 88, 88, 88
        model, _ = M.get_model(1, D, (H, Z), (Hc, C), "adam", use_x_prev=c["xp"])
        wseed, wscale = 4000 + ci, 1.6
        shas = _randomise(model, wseed, wscale)
        w_enc = M.make_w_encoder(model, D)
        z_enc = M.make_z_encoder(model, D, C, (H, Z))
        dec = M.make_decoder(model, (H, Z), C, use_x_prev=c["xp"])
        x_seed = _rolls(rng, D).astype(np.float64)
        label = int(rng.integers(0, C))
        w_val = None if c["infer"] else ks.to_categorical(label, C)
        tw, tz, td = _Tape(w_enc), _Tape(z_enc), _Tape(dec)
        np.random.seed(c["np_seed"])
        out = M.generate_sample(td, tw, tz, x_seed, c["nsteps"], w_val=w_val, use_z_prior=c["zprior"], use_x_prev=c["xp"])
        after = np.random.rand()
        probs = np.stack([np.asarray(p).reshape(-1) for p in td.calls]).astype(np.float32)
        draws = [("randn", C - 1)] * len(tw.calls)
        for _ in range(len(td.calls)):
            draws += [("randn", Z), ("rand", D)]
        margin = _margin(probs, c["np_seed"], draws)
        assert abs(np.random.rand() - after) < 1e-15, "draw order of the re-derivation differs from the reference run"
        assert margin > 1e-5, (name, margin)
        rec[name + "/cfg"] = np.array(json.dumps(dict(
            C=C, Z=Z, D=D, H=H, Hc=Hc, use_x_prev=c["xp"], nsteps=c["nsteps"], infer=c["infer"], use_z_prior=c["zprior"],
            np_seed=c["np_seed"], label=label, stream_after=after, min_margin=margin, wseed=wseed, wscale=wscale, wsha=shas)))
        rec[name + "/x_seed"] = x_seed.astype(np.uint8)
        rec[name + "/out"] = out.astype(np.uint8)
        rec[name + "/probs"] = probs
        print("vae_sampler", name, "notes", int(out.sum()), "of", out.size, "margin %.2e" % margin)
    np.savez_compressed(os.path.join(HERE, "vae_sampler.npz"), **rec)


if __name__ == "__main__":
    which = sys.argv[1:] or ["pianodata", "adamwn", "vrnn_train", "vae_train", "vrnn_sampler", "vae_sampler"]
    for w in which:
        globals()["golden_" + w]()
