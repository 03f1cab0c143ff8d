"""
Classifying VAE+LSTM (STORN) -- training CLI, same flags as code/cl_vrnn/train.py:76-119.
Run from the repo root:  python -m clvae_b200.cl_vrnn.train RUN --use_x_prev --train_file ...
(or  python classifying-vae-lstm_b200/cl_vrnn/train.py RUN ...).
"""
import argparse
import os
import sys
import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import clvae_b200  # noqa: F401
    __package__ = "clvae_b200.cl_vrnn"

from ..utils.pianoroll import PianoData
from ..utils.model_utils import get_callbacks, save_model_in_pieces, AnnealLossWeight, init_adam_wn
from ..utils.weightnorm import data_based_init
from ..keras_like import Variable
from .model import get_model
from ..parallel import init_from_env


def to_categorical(y, num_classes):
    y = np.asarray(y, dtype=int).ravel()
    out = np.zeros((len(y), num_classes), dtype=np.float32)
    out[np.arange(len(y)), y] = 1.0
    return out


def train(args):
    # one process per GPU under torchrun (RANK / LOCAL_RANK / WORLD_SIZE): --batch_size stays the GLOBAL batch,
    # every rank steps batch_size / WORLD_SIZE sequences of it and the gradients are all-reduced inside the step
    world, rank, _ = init_from_env()
    if args.batch_size % world:
        raise SystemExit("--batch_size %d must be a multiple of the %d ranks" % (args.batch_size, world))
    P = PianoData(args.train_file, batch_size=args.batch_size, seq_length=args.seq_length, step_length=1,
                  return_y_next=args.predict_next or args.use_x_prev, return_y_hist=True,
                  squeeze_x=False, squeeze_y=False)
    args.n_classes = len(np.unique(P.train_song_keys))
    w = to_categorical(P.train_song_keys, args.n_classes)
    wv = to_categorical(P.valid_song_keys, args.n_classes)
    print("Training with {} classes.".format(args.n_classes))
    assert not (args.predict_next and args.use_x_prev), "Can't use --predict_next if using --use_x_prev"
    callbacks = get_callbacks(args, patience=args.patience, min_epoch=max(args.kl_anneal, args.w_kl_anneal) + 1,
                              do_log=args.do_log)
    if args.kl_anneal > 0:
        assert args.kl_anneal <= args.num_epochs, "invalid kl_anneal"
        kl_weight = Variable(0.1)
        callbacks += [AnnealLossWeight(kl_weight, name="kl_weight", final_value=1.0, n_epochs=args.kl_anneal)]
    else:
        kl_weight = 1.0
    if args.w_kl_anneal > 0:
        assert args.w_kl_anneal <= args.num_epochs, "invalid w_kl_anneal"
        w_kl_weight = Variable(0.0)
        callbacks += [AnnealLossWeight(w_kl_weight, name="w_kl_weight", final_value=1.0, n_epochs=args.w_kl_anneal)]
    else:
        w_kl_weight = 1.0

    args.optimizer, was_adam_wn = init_adam_wn(args.optimizer)
    model, _ = get_model(args.batch_size // world, args.original_dim, args.intermediate_dim, args.latent_dim,
                         args.seq_length, args.n_classes, args.use_x_prev, args.optimizer, args.class_weight,
                         kl_weight, w_kl_weight=w_kl_weight, w_log_var_prior=args.w_log_var_prior,
                         predict_next=args.predict_next, world_size=world, rank=rank)
    args.optimizer = 'adam-wn' if was_adam_wn else args.optimizer
    os.makedirs(args.model_dir, exist_ok=True)
    if rank == 0:
        save_model_in_pieces(model, args)      # RUN.json keeps the GLOBAL batch_size

    print(P.x_train.shape, P.y_train.shape)
    if args.use_x_prev:
        x, y = [P.y_train, P.x_train], P.y_train
        xv, yv = [P.y_valid, P.x_valid], P.y_valid
    else:
        x, y = P.x_train, P.y_train
        xv, yv = P.x_valid, P.y_valid
    ytr = [y, w, w, y]
    yva = [yv, wv, wv, yv]
    data_based_init(model, x[:100])
    if getattr(args, 'materialize_windows', False):
        history = model.fit(x, ytr, shuffle=True, epochs=args.num_epochs, batch_size=args.batch_size,
                            callbacks=callbacks, validation_data=(xv, yva))
    else:
        # same batches, same labels, same results -- but each split lives in HBM as ONE uint8 roll plus window
        # offsets instead of the 17x-materialised [n, L(+1), 88] windows (utils/pianoroll.py:52-62)
        from ..utils.pianoroll import DeviceRolls
        window = args.seq_length + int(args.predict_next or args.use_x_prev)
        tr, _ = DeviceRolls.from_pickle(args.train_file, 'train', window, args.batch_size)
        va, _ = DeviceRolls.from_pickle(args.train_file, 'valid', window, args.batch_size)
        assert len(tr.labels) == len(P.train_song_keys) and np.array_equal(tr.labels, P.train_song_keys)
        history = model.fit_rolls(tr, va, shuffle=True, epochs=args.num_epochs, callbacks=callbacks)
    best_ind = np.argmin([v if i >= min(args.kl_anneal, args.w_kl_anneal) else np.inf
                          for i, v in enumerate(history.history['val_loss'])])
    best_loss = {k: history.history[k][best_ind] for k in history.history}
    return model, best_loss


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('run_name', type=str, help='tag for current run')
    parser.add_argument('--batch_size', type=int, default=200, help='batch size')
    parser.add_argument('--optimizer', type=str, default='adam-wn', help='optimizer name')
    parser.add_argument('--num_epochs', type=int, default=200, help='number of epochs')
    parser.add_argument('--original_dim', type=int, default=88, help='input dim')
    parser.add_argument('--latent_dim', type=int, default=2, help='latent dim')
    parser.add_argument('--intermediate_dim', type=int, default=88, help='intermediate dim')
    parser.add_argument('--seq_length', type=int, default=16, help='sequence length (to use as history)')
    parser.add_argument('--class_weight', type=float, default=1.0, help='relative weight on classifying key')
    parser.add_argument("--predict_next", action="store_true", help="use x_t to 'autoencode' x_{t+1}")
    parser.add_argument("--materialize_windows", action="store_true",
                        help="(B200 build) feed model.fit the materialised PianoData windows like the reference; "
                             "default: one device-resident roll per split + window offsets (same batches)")
    parser.add_argument("--do_log", action="store_true", help="save log files")
    parser.add_argument("--w_log_var_prior", type=float, default=0.0, help="log variance prior on w")
    parser.add_argument("--kl_anneal", type=int, default=0, help="number of epochs before kl loss term is 1.0")
    parser.add_argument("--w_kl_anneal", type=int, default=0, help="number of epochs before w's kl loss term is 1.0")
    parser.add_argument('--patience', type=int, default=5, help='# of epochs, for early stopping')
    parser.add_argument("--use_x_prev", action="store_true", help="use x_{t-1} to help z_t decode x_t")
    parser.add_argument('--log_dir', type=str, default='../data/logs', help='basedir for saving log files')
    parser.add_argument('--model_dir', type=str, default='../data/models', help='basedir for saving model weights')
    parser.add_argument('--train_file', type=str, default='../data/input/JSB Chorales_Cs.pickle',
                        help='file of training data (.pickle)')
    return parser


if __name__ == '__main__':
    _parser = build_parser()
    _args = _parser.parse_args()
    if _args.intermediate_dim != 88 or _args.original_dim != 88:
        _parser.error("the B200 recurrence kernels keep the 88x352 recurrent kernel of one LSTM in the registers of "
                      "one CTA (352 threads x 88): --intermediate_dim and --original_dim are fixed at 88, the value "
                      "every configuration of the reference uses")
    train(_args)
