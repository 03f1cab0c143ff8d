"""
Epoch-level training-loop pieces: re-host of utils/model_utils.py (callbacks, checkpoint naming,
loss-weight annealing, model-args JSON).  Same classes, arguments and quirks.
"""
import json
import os.path
import numpy as np

from ..keras_like import Callback, K_eval, K_set_value
from .weightnorm import AdamWithWeightnorm, data_based_init  # noqa: F401


class AnnealLossWeight(Callback):
    """utils/model_utils.py:19-50: linear (or sigmoid, slope>0) ramp of a loss weight from its
    start value to final_value over n_epochs, set at on_epoch_begin."""
    def __init__(self, beta, name="beta", n_epochs=10, final_value=1.0, slope=0):
        super(AnnealLossWeight, self).__init__()
        self.beta = beta
        self.name = name
        self.slope = slope
        self.n_epochs = n_epochs
        self.start_value = K_eval(beta)
        self.final_value = final_value
        self.all_done = False

    def next_weight(self, x):
        if self.slope > 0:
            return 1 / (1 + np.exp(-self.slope * (x - 0.5)))
        return 1.0 * x

    def on_epoch_begin(self, epoch, logs={}):
        if self.all_done:
            return
        if epoch >= self.n_epochs:
            next_val = self.final_value
            self.all_done = True
        else:
            next_val = self.start_value + self.next_weight(1.0 * epoch / self.n_epochs) * (self.final_value - self.start_value)
        K_set_value(self.beta, next_val)
        print("+++++ {}: {}".format(self.name, K_eval(self.beta)))


def init_adam_wn(optimizer):
    if optimizer == 'adam-wn':
        return AdamWithWeightnorm(lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-08, decay=0.0), True
    return optimizer, False


class EarlyStoppingAfterEpoch(Callback):
    """utils/model_utils.py:59-104."""
    def __init__(self, monitor='val_loss', min_epoch=0, min_delta=0, patience=0, verbose=0, mode='auto'):
        super(EarlyStoppingAfterEpoch, self).__init__()
        self.monitor = monitor
        self.patience = patience
        self.verbose = verbose
        self.min_epoch = min_epoch
        self.min_delta = min_delta
        self.wait = 0
        self.stopped_epoch = 0
        assert mode in ['auto', 'min', 'max']
        if mode == 'min':
            self.monitor_op = np.less
        elif mode == 'max':
            self.monitor_op = np.greater
        elif 'acc' in self.monitor or self.monitor.startswith('fmeasure'):
            self.monitor_op = np.greater
        else:
            self.monitor_op = np.less
        self.min_delta *= 1 if self.monitor_op == np.greater else -1

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.stopped_epoch = 0
        self.best = np.inf if self.monitor_op == np.less else -np.inf

    def on_epoch_end(self, epoch, logs=None):
        if epoch < self.min_epoch:
            return
        current = logs.get(self.monitor)
        if self.monitor_op(current - self.min_delta, self.best):
            self.best = current
            self.wait = 0
        else:
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
            self.wait += 1


class ModelCheckpointAfterEpoch(Callback):
    """utils/model_utils.py:106-140: weights only, best val_loss only, only for epoch >= min_epoch."""
    def __init__(self, filepath, monitor, min_epoch=0, save_weights_only=True, save_best_only=True,
                 mode='auto', verbose=False):
        super(ModelCheckpointAfterEpoch, self).__init__()
        assert save_best_only and not verbose
        assert mode in ['auto', 'min', 'max']
        self.filepath = filepath
        self.monitor = monitor
        self.min_epoch = min_epoch
        self.save_weights_only = save_weights_only
        if mode == 'max' or (mode == 'auto' and ('acc' in self.monitor or self.monitor.startswith('fmeasure'))):
            self.monitor_op, self.best = np.greater, -np.inf
        else:
            self.monitor_op, self.best = np.less, np.inf

    def on_epoch_end(self, epoch, logs=None):
        if epoch < self.min_epoch:
            return
        logs = logs or {}
        filepath = self.filepath.format(epoch=epoch, **logs)
        current = logs.get(self.monitor)
        if self.monitor_op(current, self.best):
            self.best = current
            self.model.save_weights(filepath, overwrite=True)


def get_callbacks(args, patience=5, min_epoch=0, do_log=False):
    """utils/model_utils.py:142-158.  QUIRK Q5 reproduced: the SAME early-stopping object is appended
    twice, so `wait` advances twice per epoch.  --do_log (TensorBoard) has no equivalent here: the
    History object and the per-epoch prints carry the same scalars."""
    chkpt_filename = os.path.join(args.model_dir, args.run_name + '.h5')
    checkpt = ModelCheckpointAfterEpoch(chkpt_filename, min_epoch=min_epoch, monitor='val_loss',
                                        save_weights_only=True, save_best_only=True)
    callbacks = [checkpt]
    if do_log:
        print("--do_log: TensorBoard logging is not available; epoch logs are printed instead")
    if patience > 0:
        early_stop = EarlyStoppingAfterEpoch(monitor='val_loss', min_epoch=min_epoch, patience=patience, verbose=0)
        callbacks.append(early_stop)
        callbacks.append(early_stop)
    return callbacks


def save_model_in_pieces(model, args):
    """utils/model_utils.py:160-167: RUN.yaml (structure) + RUN.json (vars(args))."""
    outfile = os.path.join(args.model_dir, args.run_name + '.yaml')
    with open(outfile, 'w') as f:
        f.write(model.to_yaml())
    outfile = os.path.join(args.model_dir, args.run_name + '.json')
    d = {k: (int(v) if isinstance(v, (np.integer,)) else v) for k, v in vars(args).items()}
    json.dump(d, open(outfile, 'w'))
