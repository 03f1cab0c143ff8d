"""
A recorder shim of the slice of the Keras 2.0.0 functional API that /root/reference/code/cl_*/model.py
uses, on PyTorch-CPU tensors -- TEST INFRASTRUCTURE (used only by tests/golden/make_golden.py, which
runs in the build container where /root/reference exists).

Purpose: EXECUTE the reference's own `get_model`, `make_w_encoder / make_z_encoder / make_decoder`,
`generate_sample` and `sample_*` source (after a mechanical Python-2 -> 3 rewrite of tuple
parameters) so that everything the REFERENCE FILES decide is pinned by running them:

  * graph wiring: which tensor feeds which layer, concat orders ([X|W], [Xp|Z|W], [w|xp|z]),
    the Wargs split, the logistic-normal Lambda (un-stabilised softmax over [s, 0]), W2 = W + 1e-10,
    Z_args = [Z_mean | Z_log_var], which inputs/targets each loss closure sees;
  * the four loss closures (executed symbolically, exactly like Keras runs them at compile time);
  * loss_weights / metrics passed to compile;
  * the sampler's control flow, sub-model wiring (incl. the fresh encoder LSTM of make_z_encoder,
    quirk Q1), and its np.random draw order.

What this shim itself restates from the published Keras 2.0.0 / TF 1.0.1 sources (still
"[K2-recall]", listed in oracle/clv_oracle.py): the primitives -- Dense, LSTM cell (gate order i,f,c,o,
hard_sigmoid, tanh), TimeDistributed, K.binary_crossentropy / K.categorical_crossentropy, the
weighted-mean loss reduction of Model.compile, categorical accuracy.

Everything is symbolic (like TF graph mode): layer calls and K.* calls build nodes, Lambda functions
and loss closures run ONCE at build time on symbolic tensors, Model.predict / Model.loss_and_grads
evaluate the nodes with torch (autograd gives the gradients TF's tf.gradients would).
K.random_normal nodes draw from an explicit noise tape, in creation order.
"""
import sys
import types
import numpy as np
import torch

EPS = 1e-7          # keras.backend.common._EPSILON
_STATE = {"rng": np.random.default_rng(0), "uid": 0}


def set_rng(seed):
    _STATE["rng"] = np.random.default_rng(seed)


class Ctx:
    def __init__(self, feeds, noise=None, dtype=torch.float64, dry=False):
        self.feeds, self.noise, self.dtype, self.memo, self.dry = feeds, (noise or {}), dtype, {}, dry


class T:
    """symbolic tensor"""
    def __init__(self, fn, inputs=(), layer=None, kind="op"):
        self.fn, self.inputs, self.layer, self.kind = fn, tuple(inputs), layer, kind
        _STATE["uid"] += 1
        self.uid = _STATE["uid"]

    def eval(self, ctx):
        if self.uid not in ctx.memo:
            if self.kind == "input":
                v = torch.zeros(self.kshape, dtype=ctx.dtype) if ctx.dry else ctx.feeds[self.uid]
            else:
                v = self.fn(ctx, *[ev(i, ctx) for i in self.inputs])
            ctx.memo[self.uid] = v
        return ctx.memo[self.uid]

    # operators used by the reference's Lambdas / losses
    def __add__(self, o): return T(lambda c, a, b: a + b, (self, o))
    def __radd__(self, o): return T(lambda c, a, b: b + a, (self, o))
    def __sub__(self, o): return T(lambda c, a, b: a - b, (self, o))
    def __rsub__(self, o): return T(lambda c, a, b: b - a, (self, o))
    def __mul__(self, o): return T(lambda c, a, b: a * b, (self, o))
    def __rmul__(self, o): return T(lambda c, a, b: b * a, (self, o))
    def __truediv__(self, o): return T(lambda c, a, b: a / b, (self, o))
    def __rtruediv__(self, o): return T(lambda c, a, b: b / a, (self, o))
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __neg__(self): return T(lambda c, a: -a, (self,))
    def __getitem__(self, idx): return T(lambda c, a: a[idx], (self,))


def ev(x, ctx):
    if isinstance(x, T):
        return x.eval(ctx)
    if isinstance(x, Variable):
        return x.value
    return x


class Variable:
    """K.variable (annealed loss weights)"""
    def __init__(self, value):
        self.value = float(value)


# ------------------------------------------------------------------ backend (keras.backend as K)
def _axis(a):
    return tuple(a) if isinstance(a, (list, tuple)) else a


K = types.ModuleType("keras.backend")
K.floatx = lambda: "float32"
K.epsilon = lambda: EPS
def _tt(a, c):
    return a if torch.is_tensor(a) else torch.as_tensor(a, dtype=c.dtype)


K.exp = lambda x: T(lambda c, a: torch.exp(_tt(a, c)), (x,))
K.log = lambda x: T(lambda c, a: torch.log(_tt(a, c)), (x,))
K.square = lambda x: T(lambda c, a: a * a, (x,))
K.abs = lambda x: T(lambda c, a: torch.abs(a), (x,))
K.sum = lambda x, axis=None, keepdims=False: T(
    lambda c, a: a.sum() if axis is None else a.sum(dim=_axis(axis), keepdim=keepdims), (x,))
K.mean = lambda x, axis=None, keepdims=False: T(
    lambda c, a: a.mean() if axis is None else a.mean(dim=_axis(axis), keepdim=keepdims), (x,))
K.zeros = lambda shape, dtype=None, name=None: T(lambda c: torch.zeros(tuple(shape), dtype=c.dtype))
K.variable = lambda value, dtype=None, name=None: Variable(value)
K.set_value = lambda v, x: setattr(v, "value", float(x))
K.get_value = lambda v: v.value
K.clip = lambda x, lo, hi: T(lambda c, a: torch.clamp(a, lo, hi), (x,))   # closed-interval gradient


def _random_normal(shape, mean=0., stddev=1.0, dtype=None, seed=None):
    node = T(None, kind="noise")
    K._noise_nodes.append(node)
    node.shape = tuple(shape)
    idx = len(K._noise_nodes) - 1

    def fn(c):
        if c.dry:
            return torch.zeros(node.shape, dtype=c.dtype)
        v = torch.as_tensor(np.asarray(c.noise[idx]), dtype=c.dtype)
        assert tuple(v.shape) == node.shape, (v.shape, node.shape)
        return mean + stddev * v
    node.fn = fn
    node.kind = "op"
    return node


K._noise_nodes = []
K.random_normal = _random_normal
_tf = types.ModuleType("tensorflow")
# tf.zeros(shape, dtype): the reference calls K.tf.zeros(batch_size, 1) (cl_vae/model.py:155, quirk Q8):
# shape = batch_size (a scalar -> 1-D), dtype enum 1 = DT_FLOAT
_tf.zeros = lambda shape, dtype=None, name=None: T(
    lambda c: torch.zeros((shape,) if np.isscalar(shape) else tuple(shape), dtype=c.dtype))
K.tf = _tf


def _k_binary_crossentropy(output, target, from_logits=False):
    """keras/backend/tensorflow_backend.py binary_crossentropy (2.0.0): clip -> logit ->
    tf.nn.sigmoid_cross_entropy_with_logits = max(l,0) - l*z + log(1+exp(-|l|))."""
    def fn(c, p, z):
        if not from_logits:
            p = torch.clamp(p, EPS, 1 - EPS)
            p = torch.log(p / (1 - p))
        return torch.clamp(p, min=0) - p * z + torch.log1p(torch.exp(-torch.abs(p)))
    return T(fn, (output, target))


def _k_categorical_crossentropy(output, target, from_logits=False):
    """... categorical_crossentropy (2.0.0): renormalise, clip, -sum(target*log(output))."""
    def fn(c, q, t):
        q = q / q.sum(dim=-1, keepdim=True)
        q = torch.clamp(q, EPS, 1 - EPS)
        return -(t * torch.log(q)).sum(dim=-1)
    return T(fn, (output, target))


K.binary_crossentropy = _k_binary_crossentropy
K.categorical_crossentropy = _k_categorical_crossentropy

losses = types.ModuleType("keras.losses")
losses.binary_crossentropy = lambda y_true, y_pred: K.mean(K.binary_crossentropy(y_pred, y_true), axis=-1)
losses.categorical_crossentropy = lambda y_true, y_pred: K.categorical_crossentropy(y_pred, y_true)


# ------------------------------------------------------------------ initializers
class RandomNormal:
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape):
        return _STATE["rng"].normal(self.mean, self.stddev, shape)


def _glorot_uniform(shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return _STATE["rng"].uniform(-lim, lim, shape)


def _orthogonal(shape):
    a = _STATE["rng"].standard_normal(shape)
    u, _, vt = np.linalg.svd(a, full_matrices=False)
    return u if u.shape == tuple(shape) else vt


def _init(spec, shape):
    if spec == "zeros":
        return np.zeros(shape)
    if spec in (None, "glorot_uniform"):
        return _glorot_uniform(shape)
    if spec == "orthogonal":
        return _orthogonal(shape)
    return spec(shape)


initializers = types.ModuleType("keras.initializers")
initializers.RandomNormal = RandomNormal


# ------------------------------------------------------------------ layers
class Layer:
    weight_names = ()

    def __init__(self, name=None):
        _STATE["uid"] += 1
        self.name = name or "%s_%d" % (type(self).__name__.lower(), _STATE["uid"])
        self.weights = None          # list of torch float64 leaves once built

    def build(self, in_dim):
        pass

    def _set(self, arrays):
        self.weights = [torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True) for a in arrays]

    def get_weights(self):
        return [w.detach().numpy().copy() for w in self.weights]

    def set_weights(self, ws):
        assert len(ws) == len(self.weights), "set_weights: %d given, layer %s has %d" % (len(ws), self.name, len(self.weights))
        for w, a in zip(self.weights, ws):
            assert tuple(w.shape) == tuple(np.shape(a)), (self.name, w.shape, np.shape(a))
        self._set(ws)

    def w(self, ctx, i):
        return self.weights[i].to(ctx.dtype)

    def reset_states(self):
        pass


def _shape_of(x):
    """static shape; nodes made by raw backend ops (inside Lambdas) get theirs from a dry evaluation on
    zeros (every Input of the reference has a full batch_shape)"""
    ks = getattr(x, "kshape", None)
    if ks is None:
        with torch.no_grad():
            ks = tuple(x.eval(Ctx({}, dry=True)).shape)
        x.kshape = ks
    return ks


def _mk(fn, inputs, layer, kshape):
    t = T(fn, inputs, layer=layer)
    t.kshape = kshape
    return t


def Input(batch_shape=None, shape=None, name=None):
    t = T(None, kind="input")
    t.kshape = tuple(batch_shape) if batch_shape is not None else (None,) + tuple(shape)
    t.name = name
    return t


_ACT = {None: lambda a: a, "linear": lambda a: a, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}


class Dense(Layer):
    weight_names = ("kernel", "bias")

    def __init__(self, units, activation=None, name=None, kernel_initializer=None, bias_initializer="zeros", **kw):
        super().__init__(name)
        self.units, self.activation = units, activation
        self.kinit, self.binit = kernel_initializer, bias_initializer

    def build(self, in_dim):
        if self.weights is None:
            self._set([_init(self.kinit, (in_dim, self.units)), _init(self.binit, (self.units,))])

    def __call__(self, x):
        ks = _shape_of(x)
        self.build(ks[-1])
        act = _ACT[self.activation]
        return _mk(lambda c, a: act(a @ self.w(c, 0) + self.w(c, 1)), (x,), self, ks[:-1] + (self.units,))


class TimeDistributed(Layer):
    """keras.layers.wrappers.TimeDistributed: the wrapped Dense applied to every timestep (with a
    static batch size Keras 2.0.0 runs it through K.rnn; same maths as a batched matmul)."""
    weight_names = ("kernel", "bias")

    def __init__(self, layer, name=None):
        super().__init__(name)
        self.layer = layer

    @property
    def weights(self):
        return self.layer.weights

    @weights.setter
    def weights(self, v):
        pass

    def get_weights(self):
        return self.layer.get_weights()

    def set_weights(self, ws):
        self.layer.set_weights(ws)

    def w(self, ctx, i):
        return self.layer.w(ctx, i)

    def __call__(self, x):
        y = self.layer(x)            # Dense broadcasts over the time axis
        y.layer = self
        return y


def hard_sigmoid(a):
    return torch.clamp(0.2 * a + 0.5, 0.0, 1.0)


class LSTM(Layer):
    """keras.layers.recurrent.LSTM 2.0.0, implementation=0: x@kernel+bias hoisted, per step
    z = x_t + h@recurrent_kernel; gate column blocks i,f,c,o; recurrent_activation hard_sigmoid,
    activation tanh; unit_forget_bias; glorot_uniform / orthogonal / zeros initialisers."""
    weight_names = ("kernel", "recurrent_kernel", "bias")

    def __init__(self, units, stateful=False, return_sequences=False, name=None, dropout=0.0, **kw):
        super().__init__(name)
        assert dropout == 0.0
        self.units, self.stateful, self.return_sequences = units, stateful, return_sequences
        self.states = None

    def build(self, in_dim):
        if self.weights is None:
            u = self.units
            b = np.zeros(4 * u); b[u:2 * u] = 1.0
            self._set([_glorot_uniform((in_dim, 4 * u)), _orthogonal((u, 4 * u)), b])

    def reset_states(self):
        self.states = None

    def __call__(self, x):
        ks = _shape_of(x)
        self.build(ks[-1])
        u = self.units

        def fn(c, a):
            B, L, _ = a.shape
            xp = a @ self.w(c, 0) + self.w(c, 2)
            U = self.w(c, 1)
            if self.stateful and self.states is not None:
                h, cc = self.states
            else:
                h, cc = a.new_zeros(B, u), a.new_zeros(B, u)
            hs = []
            for t in range(L):
                z = xp[:, t] + h @ U
                i = hard_sigmoid(z[:, :u]); f = hard_sigmoid(z[:, u:2 * u])
                g = torch.tanh(z[:, 2 * u:3 * u]); o = hard_sigmoid(z[:, 3 * u:])
                cc = f * cc + i * g
                h = o * torch.tanh(cc)
                hs.append(h)
            if self.stateful and not c.dry:
                self.states = (h.detach(), cc.detach())
            return torch.stack(hs, dim=1) if self.return_sequences else h
        return _mk(fn, (x,), self, (ks[0], ks[1], u) if self.return_sequences else (ks[0], u))


class Lambda(Layer):
    def __init__(self, function, output_shape=None, name=None):
        super().__init__(name)
        self.function = function

    def __call__(self, x):
        y = self.function(x)         # runs the reference's closure on symbolic tensors, once
        out = T(lambda c, a: a, (y,), layer=self)
        out.kshape = _shape_of(y)
        return out


class _Concat(Layer):
    pass


def concatenate(inputs, axis=-1, name=None):
    """keras.layers.concatenate -- also called on raw backend tensors inside the reference's Lambdas."""
    lay = _Concat(name)
    shapes = [_shape_of(i) for i in inputs]
    ks = shapes[0][:-1] + (sum(s[-1] for s in shapes),)
    return _mk(lambda c, *xs: torch.cat(list(xs), dim=axis), tuple(inputs), lay, ks)


class RepeatVector(Layer):
    def __init__(self, n, name=None):
        super().__init__(name)
        self.n = n

    def __call__(self, x):
        ks = _shape_of(x)
        return _mk(lambda c, a: a[:, None, :].expand(a.shape[0], self.n, a.shape[1]), (x,), self, (ks[0], self.n, ks[1]))


class Flatten(Layer):
    def __call__(self, x):
        ks = _shape_of(x)
        return _mk(lambda c, a: a.reshape(a.shape[0], -1), (x,), self, (ks[0], int(np.prod(ks[1:]))))


class Reshape(Layer):
    def __init__(self, target_shape, name=None):
        super().__init__(name)
        self.target_shape = tuple(target_shape)

    def __call__(self, x):
        ks = _shape_of(x)
        return _mk(lambda c, a: a.reshape((a.shape[0],) + self.target_shape), (x,), self, (ks[0],) + self.target_shape)


# ------------------------------------------------------------------ Model
def _categorical_accuracy(y_true, y_pred):
    return (y_true.argmax(dim=-1) == y_pred.argmax(dim=-1)).to(y_pred.dtype).mean()


class Model:
    def __init__(self, inputs, outputs):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        self.single_output = not isinstance(outputs, (list, tuple))
        self.layers = []
        seen = set()

        def walk(t):
            if not isinstance(t, T) or t.uid in seen:
                return
            seen.add(t.uid)
            for i in t.inputs:
                walk(i)
            if t.layer is not None and t.layer not in self.layers:
                self.layers.append(t.layer)
        for o in self.outputs:
            walk(o)
        self.output_names = [o.layer.name for o in self.outputs]
        self.stop_training = False

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError("No such layer: " + name)

    def reset_states(self):
        for l in self.layers:
            l.reset_states()

    def weighted(self):
        """(layer name, weight name, leaf) of every weighted layer, in graph order"""
        out = []
        for l in self.layers:
            if l.weight_names and l.weights is not None:
                for n, w in zip(l.weight_names, l.weights):
                    out.append((l.name, n, w))
        return out

    def _feeds(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        assert len(xs) == len(self.inputs), "model expects %d inputs" % len(self.inputs)
        return xs

    def predict(self, x, batch_size=None):
        """float32, like Keras (floatx) -- returns numpy float32 arrays"""
        xs = self._feeds(x)
        feeds = {}
        for t, v in zip(self.inputs, xs):
            v = torch.as_tensor(np.asarray(v), dtype=torch.float32)
            assert tuple(v.shape) == tuple(t.kshape), "input %s: shape %s, expected %s" % (t.name, tuple(v.shape), t.kshape)
            feeds[t.uid] = v
        ctx = Ctx(feeds, dtype=torch.float32)
        with torch.no_grad():
            outs = [o.eval(ctx).numpy().copy() for o in self.outputs]
        return outs[0] if self.single_output else outs

    def compile(self, optimizer=None, loss=None, loss_weights=None, metrics=None):
        self.optimizer, self.loss, self.loss_weights, self.metrics = optimizer, loss, loss_weights or {}, metrics or {}
        # like keras.engine.training.Model.compile: one target placeholder per output, loss function
        # called symbolically, weighted-mean reduction, total = sum_i loss_weight_i * loss_i
        self.targets, self.out_losses = [], []
        for name, o in zip(self.output_names, self.outputs):
            y_true = T(None, kind="input")
            self.targets.append(y_true)
            score = loss[name](y_true, o)

            def reduce_(c, s):           # _weighted_masked_objective with unit sample weights [B]
                if s.dim() > 1:
                    s = s.mean(dim=tuple(range(1, s.dim())))
                return s.mean()
            self.out_losses.append(T(reduce_, (score,)))

    def loss_and_grads(self, x, y, noise, dtype=torch.float64):
        """What one train_function call computes before the optimizer update: total loss, the per-output
        losses, metrics and d(total)/d(weights).  `noise`: list of arrays for the K.random_normal nodes of
        this graph, in creation order."""
        xs, ys = self._feeds(x), list(y)
        feeds = {}
        for t, v in zip(self.inputs, xs):
            feeds[t.uid] = torch.as_tensor(np.asarray(v), dtype=dtype)
        for t, v in zip(self.targets, ys):
            feeds[t.uid] = torch.as_tensor(np.asarray(v), dtype=dtype)
        mine = self._noise_index()
        assert len(mine) == len(noise), "graph has %d random_normal nodes" % len(mine)
        ctx = Ctx(feeds, dict(zip(mine, noise)), dtype)
        ws = self.weighted()
        for _, _, w in ws:
            w.grad = None
        per = [l.eval(ctx) for l in self.out_losses]
        total = 0.0
        for name, l in zip(self.output_names, per):
            total = total + float(ev(self.loss_weights.get(name, 1.0), ctx)) * l
        total.backward()
        res = {"loss": float(total)}
        for name, l in zip(self.output_names, per):
            res[name + "_loss"] = float(l)
        for name, m in self.metrics.items():
            assert m in ("accuracy", "acc")
            i = self.output_names.index(name)
            res[name + "_acc"] = float(_categorical_accuracy(feeds[self.targets[i].uid], self.outputs[i].eval(ctx)))
        grads = {"%s.%s" % (ln, wn): (w.grad.numpy().copy() if w.grad is not None else np.zeros(tuple(w.shape)))
                 for ln, wn, w in ws}
        outs = {n: o.eval(ctx).detach().numpy().copy() for n, o in zip(self.output_names, self.outputs)}
        return res, grads, outs

    def _noise_index(self):
        ids, seen = [], set()

        def walk(t):
            if not isinstance(t, T) or t.uid in seen:
                return
            seen.add(t.uid)
            for i in t.inputs:
                walk(i)
            if t in K._noise_nodes:
                ids.append(K._noise_nodes.index(t))
        for o in self.outputs:
            walk(o)
        return sorted(ids)


def to_categorical(y, num_classes=None):
    """keras.utils.np_utils.to_categorical 2.0.0: a scalar gives [1, C]."""
    y = np.array(y, dtype="int").ravel()
    if not num_classes:
        num_classes = np.max(y) + 1
    n = y.shape[0]
    out = np.zeros((n, num_classes))
    out[np.arange(n), y] = 1
    return out


def install():
    """register the shim as `keras` in sys.modules (only ever inside make_golden.py)"""
    keras = types.ModuleType("keras")
    layers = types.ModuleType("keras.layers")
    for k, v in dict(Input=Input, Dense=Dense, LSTM=LSTM, TimeDistributed=TimeDistributed, Lambda=Lambda,
                     concatenate=concatenate, RepeatVector=RepeatVector, Flatten=Flatten, Reshape=Reshape).items():
        setattr(layers, k, v)
    models = types.ModuleType("keras.models")
    models.Model = Model
    utils = types.ModuleType("keras.utils")
    utils.to_categorical = to_categorical
    keras.layers, keras.models, keras.backend, keras.losses, keras.initializers, keras.utils = \
        layers, models, K, losses, initializers, utils
    sys.modules.update({"keras": keras, "keras.layers": layers, "keras.models": models, "keras.backend": K,
                        "keras.losses": losses, "keras.initializers": initializers, "keras.utils": utils})
    return keras
