"""
Keras-shaped host objects over the device Engine: the subset of keras.models.Model the reference's
scripts actually use: fit / train_on_batch / evaluate / predict / save_weights /
load_weights / get_layer(...).get_weights()/set_weights() / to_yaml / stop_training, plus History and
the Callback protocol.  Loss/metric names follow Keras 2.0.0 [K2-recall]: loss, <output>_loss,
W_acc, val_*.
"""
import json
import time
import numpy as np
import torch

from .engine import Engine, LOSS_NAMES


class History:
    def __init__(self):
        self.history = {}
        self.epoch = []


class Callback:
    """keras.callbacks.Callback protocol (set_model, on_train_begin, on_epoch_begin/end)."""
    def __init__(self):
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


class Variable:
    """Stand-in for K.variable(value): a mutable scalar (annealed loss weights,
    cl_vrnn/train.py:32-43; K.set_value / K.eval in utils/model_utils.py:29,49-50)."""
    def __init__(self, value):
        self.value = float(value)

    def __float__(self):
        return self.value


def K_eval(v):
    return float(v)


def K_set_value(v, x):
    v.value = float(x)


class Layer:
    def __init__(self, model, name, tensors):
        self._model, self.name, self._tensors = model, name, tensors

    def get_weights(self):
        return [self._model.engine.view(t).detach().cpu().numpy().copy() for t in self._tensors]

    def set_weights(self, ws):
        assert len(ws) == len(self._tensors)
        for t, w in zip(self._tensors, ws):
            v = self._model.engine.view(t)
            v.copy_(torch.as_tensor(np.asarray(w), dtype=torch.float32).reshape(v.shape).to(v.device))


def _binary_u8(a, what):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        b = a.astype(np.uint8)
        if not np.array_equal(b, a):
            raise ValueError("%s must be a binary piano-roll (values in {0,1}); the B200 path keeps "
                             "rolls as uint8" % what)
        a = b
    elif a.size and a.max() > 1:
        raise ValueError("%s must be a binary piano-roll" % what)
    return a


class BaseModel:
    """Shared fit/evaluate machinery.  Subclasses set: output_names (4 outputs in Keras order),
    layer_tensors (layer name -> tensor names), all_layer_names (model.layers order)."""
    output_names = None
    acc_name = None

    def __init__(self, engine, kl_weight, w_kl_weight):
        self.engine = engine
        self._kl_w, self._wkl_w = kl_weight, w_kl_weight
        self.stop_training = False
        self.metrics_names = (["loss"] + [n + "_loss" for n in self.output_names] + [self.acc_name])

    # ------------------------------------------------------------------ weights
    def get_layer(self, name):
        return Layer(self, name, self.layer_tensors[name])

    @property
    def weighted_layers(self):
        return list(self.layer_tensors.keys())

    def get_weights(self):
        return [w for n in self.weighted_layers for w in self.get_layer(n).get_weights()]

    def set_weights(self, ws):
        i = 0
        for n in self.weighted_layers:
            k = len(self.layer_tensors[n])
            self.get_layer(n).set_weights(ws[i:i + k])
            i += k

    def save_weights(self, filepath, overwrite=True):
        """Keras-2.0.0 save_weights HDF5 layout (utils/model_utils.py:138).  Data parallel: the
        replicas are bit-identical, rank 0 writes."""
        if self.engine.rank != 0:
            return
        from .utils import hdf5
        layers = []
        for lname in self.all_layer_names:
            tens = self.layer_tensors.get(lname, [])
            names = ["%s/%s:0" % (lname, t.split(".", 1)[1]) for t in tens]
            layers.append((lname, names, [self.engine.view(t).detach().cpu().numpy() for t in tens]))
        hdf5.save_keras_weights(filepath, layers)

    def load_weights(self, filepath):
        """Loads by ORDER of weighted layers, like Keras' load_weights (cl_vrnn/model.py:281)."""
        from .utils import hdf5
        layers = [(n, ws) for n, ws in hdf5.load_keras_weights(filepath) if len(ws) > 0]
        if len(layers) != len(self.weighted_layers):
            raise ValueError("file has %d weighted layers, model has %d" % (len(layers), len(self.weighted_layers)))
        for (fname, ws), lname in zip(layers, self.weighted_layers):
            self.get_layer(lname).set_weights(ws)

    def to_yaml(self):
        """Structural YAML only (the reference's embeds Py2-marshalled Lambda bytecode and is never
        read back, cl_vrnn/model.py:276)."""
        e = self.engine
        lines = ["backend: b200-cuda", "class_name: Model", "config:", "  name: %s" % type(self).__name__,
                 "  layers:"]
        for lname in self.all_layer_names:
            lines.append("  - name: %s" % lname)
            for t in self.layer_tensors.get(lname, []):
                lines.append("    %s: %s" % (t.split(".", 1)[1], list(e.view(t).shape)))
        return "\n".join(lines) + "\n"

    # ------------------------------------------------------------------ batches
    def _sync_weights_of_losses(self):
        self.engine.set_loss_weights(kl_weight=float(self._kl_w), w_kl_weight=float(self._wkl_w))

    def _logs(self, d, prefix=""):
        o = self.output_names
        return {prefix + "loss": d["loss"], prefix + o[0] + "_loss": d["vae"], prefix + o[1] + "_loss": d["w_kl"],
                prefix + o[2] + "_loss": d["w_rec"], prefix + o[3] + "_loss": d["z_kl"],
                prefix + self.acc_name: d["acc"]}

    def _as_list(self, d):
        lg = self._logs(d)
        return [lg[k] for k in self.metrics_names]

    def train_on_batch_windows(self, win_u8, labels_i32):
        """Fast public path: one batch as uint8 windows [B, W, D] (W = L+1 with use_x_prev) + int32
        key labels, host (pinned) or device tensors.  H2D, one graph replay, D2H of the scalars."""
        self._sync_weights_of_losses()
        self.engine.stage_windows(win_u8, labels_i32)
        self.engine.run(train=True, gen_noise=True)
        return self.engine.read_losses()

    def _windows_from_inputs(self, x, y=None):
        raise NotImplementedError

    def _labels_from_targets(self, y):
        w = np.asarray(y[1])
        lab = w.argmax(-1).astype(np.int32)
        if not np.array_equal(np.eye(w.shape[-1], dtype=w.dtype)[lab], w):
            raise ValueError("key targets must be one-hot (to_categorical), as in the reference")
        return lab

    def train_on_batch(self, x, y):
        win = torch.from_numpy(self._windows_from_inputs(x, y))
        lab = torch.from_numpy(self._labels_from_targets(y))
        return self._as_list(self.train_on_batch_windows(win, lab))

    def test_on_batch(self, x, y):
        self._sync_weights_of_losses()
        self.engine.stage_windows(torch.from_numpy(self._windows_from_inputs(x, y)),
                                  torch.from_numpy(self._labels_from_targets(y)))
        self.engine.run(train=False, gen_noise=True)
        return self._as_list(self.engine.read_losses())

    # ------------------------------------------------------------------ epochs
    def _run_epoch(self, roll_dev, off_dev, labs_dev, order, train):
        """One pass over a device-resident split given as (roll, first frame of every window, key label);
        losses accumulate on the device, one D2H at the end."""
        e = self.engine
        B, Bg = e.B, e.B * e.world_size        # data parallel: every rank steps its slice of each global batch
        n = len(order)
        assert n % Bg == 0, "sample count must be a multiple of the global batch_size (PianoData guarantees it)"
        e.roll = roll_dev
        off_all = off_dev[order].contiguous()
        lab_all = labs_dev[order].contiguous()
        acc = torch.zeros(8, device=e.dev)
        for i in range(n // Bg):
            lo = i * Bg + e.rank * B            # == parallel.local_batch(order, i, B, world_size, rank)
            e.stage_offsets(off_all[lo:lo + B], lab_all[lo:lo + B])
            e.run(train=train, gen_noise=True)
            acc += e._loss_src                  # already the GLOBAL means (summed over ranks inside the step)
        v = (acc / (n // Bg)).tolist()
        d = dict(zip(LOSS_NAMES, v[:5]))
        h = e.hyper
        d["loss"] = d["vae"] + h["w_kl_weight"] * d["w_kl"] + h["class_weight"] * d["w_rec"] + h["kl_weight"] * d["z_kl"]
        return d

    def _split_from_windows(self, x, y, overlap):
        """materialised windows -> (roll, window offsets, labels) on the device"""
        e = self.engine
        wins = torch.from_numpy(self._windows_from_inputs(x, y, overlap=overlap)).to(e.dev).reshape(-1)
        labs = torch.from_numpy(self._labels_from_targets(y)).to(e.dev)
        off = (torch.arange(labs.numel(), dtype=torch.int32, device=e.dev) * e.W).contiguous()
        return wins, off, labs

    def _split_from_rolls(self, dr):
        """utils.pianoroll.DeviceRolls -> the same triple WITHOUT materialising the sliding windows: the roll of
        the whole split is uploaded once, a window is only its first-frame offset"""
        e = self.engine
        if dr.window != e.W or e.x_shift != 0 or e.y_shift not in (0, 1):
            raise ValueError("DeviceRolls window %d does not match the model's %d-frame windows" % (dr.window, e.W))
        return (torch.from_numpy(dr.roll).to(e.dev).reshape(-1), torch.from_numpy(dr.win_off).to(e.dev),
                torch.from_numpy(np.asarray(dr.labels, dtype=np.int32)).to(e.dev))

    def _fit_core(self, train, val, shuffle, epochs, callbacks, verbose):
        e = self.engine
        n = train[2].numel()
        callbacks = list(callbacks or [])
        hist = History()
        for cb in callbacks:
            cb.set_model(self)
            cb.on_train_begin({})
        self.stop_training = False
        for epoch in range(epochs):
            for cb in callbacks:
                cb.on_epoch_begin(epoch, {})
            self._sync_weights_of_losses()
            t0 = time.time()
            from .parallel import shared_permutation
            order = shared_permutation(n, shuffle, device=e.dev, group=e.pg).to(e.dev)   # one order for all ranks
            logs = self._logs(self._run_epoch(train[0], train[1], train[2], order, True))
            if val is not None:
                vorder = torch.arange(val[2].numel(), device=e.dev)
                logs.update(self._logs(self._run_epoch(val[0], val[1], val[2], vorder, False), "val_"))
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            if verbose and e.rank == 0:
                print("Epoch %d/%d - %.1fs - %s" % (epoch + 1, epochs, time.time() - t0,
                      " - ".join("%s: %.4f" % (k, logs[k]) for k in sorted(logs))))
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end({})
        self.history = hist
        return hist

    def fit(self, x, y, shuffle=True, epochs=1, batch_size=None, callbacks=None, validation_data=None,
            verbose=1):
        """model.fit as the reference calls it (cl_vrnn/train.py:66-71): per-epoch shuffle,
        validation pass every epoch with noise still sampled, callbacks, History."""
        e = self.engine
        if batch_size is not None and batch_size != e.B * e.world_size:
            raise ValueError("batch_size is baked into the model (%d x %d rank(s)), got %d" % (e.B, e.world_size, batch_size))
        # window geometry (frames per window, where `current` / the target sit) is decided ONCE for both
        # splits: a validation split stored differently from the training split would change e.W under
        # the training offsets
        ov = self._overlaps(x, y)
        if validation_data is not None:
            ov = ov and self._overlaps(validation_data[0], validation_data[1])
        train = self._split_from_windows(x, y, ov)
        val = None
        if validation_data is not None:
            val = self._split_from_windows(validation_data[0], validation_data[1], ov)
        return self._fit_core(train, val, shuffle, epochs, callbacks, verbose)

    def fit_rolls(self, train_rolls, valid_rolls=None, shuffle=True, epochs=1, callbacks=None, verbose=1):
        """fit on utils.pianoroll.DeviceRolls splits: identical batches and results to fit() on the PianoData
        windows of the same pickle (same enumeration, trim and labels), but the 17x-materialised
        [n, L+1, 88] windows never exist -- each split is one uint8 roll in HBM plus int32 window offsets
        that the kernels gather from."""
        train = self._split_from_rolls(train_rolls)
        val = self._split_from_rolls(valid_rolls) if valid_rolls is not None else None
        return self._fit_core(train, val, shuffle, epochs, callbacks, verbose)

    def evaluate(self, x, y, batch_size=None, verbose=0):
        e = self.engine
        wins, off, labs = self._split_from_windows(x, y, None)
        self._sync_weights_of_losses()
        return self._as_list(self._run_epoch(wins, off, labs, torch.arange(labs.numel(), device=e.dev), False))
