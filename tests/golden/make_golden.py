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
import builtins, hashlib, importlib, json, os, pickle, sys, types
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


if __name__ == "__main__":
    golden_pianodata()
    golden_adamwn()
