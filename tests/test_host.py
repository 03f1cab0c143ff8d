"""CPU tests of the host-side drop-in pieces: HDF5 weights layout, MIDI events, callback semantics
(incl. the reference's quirks), CLI surface."""
import os
import struct
import types
import numpy as np
import pytest

import clvae_b200  # noqa: F401
from clvae_b200.utils import hdf5, midi_utils, model_utils
from clvae_b200.keras_like import Variable


def test_hdf5_keras_layout_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    layers = [("current", [], []),
              ("hW", ["hW/kernel:0", "hW/bias:0"], [rng.normal(size=(1408, 88)), rng.normal(size=(88,))]),
              ("lambda_1", [], []),
              ("encoder_h", ["encoder_h/kernel:0", "encoder_h/recurrent_kernel:0", "encoder_h/bias:0"],
               [rng.normal(size=(98, 352)), rng.normal(size=(88, 352)), rng.normal(size=(352,))]),
              ("Z_mean", ["Z_mean/kernel:0", "Z_mean/bias:0"], [rng.normal(size=(88, 2)), rng.normal(size=(2,))])]
    p = str(tmp_path / "w.h5")
    hdf5.save_keras_weights(p, layers)
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0                 # superblock v0
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)                  # end-of-file address
    out = hdf5.load_keras_weights(p)
    assert [n for n, _ in out] == [l[0] for l in layers]                    # layer_names order, weight-less kept
    for (_, _, A), (_, Bm) in zip(layers, out):
        assert len(A) == len(Bm)
        for a, b in zip(A, Bm):
            assert b.dtype == np.float32 and np.array_equal(a.astype(np.float32), b)
    r = hdf5._Reader(raw)
    at = r.attrs(r.root_hdr)
    assert at["backend"] == b"tensorflow" and at["keras_version"] == b"2.0.0"
    g = r.children(r.root_hdr)["encoder_h"]
    assert [n.decode() for n in r.attrs(g)["weight_names"]] == layers[3][1]
    assert sorted(r.children(r.find(g, "encoder_h"))) == ["bias:0", "kernel:0", "recurrent_kernel:0"]


def test_midi_events_follow_reference_logic(tmp_path):
    roll = np.zeros((4, 88)); roll[0, [39, 43]] = 1; roll[1, [39]] = 1; roll[3, [50]] = 1
    ev = midi_utils.MidiWriter().events(roll)
    # pitch = index + 21; only the first event of a frame carries the delta; offs before ons
    assert ev == [(120, b"\x90\x3c\x64"), (0, b"\x90\x40\x64"), (120, b"\x80\x40\x00"),
                  (120, b"\x80\x3c\x00"), (120, b"\x90\x47\x64"), (120, b"\x80\x47\x00")]
    f = midi_utils.write_sample(roll, str(tmp_path), "s", isHalfAsSlow=True)      # frames doubled
    raw = open(f, "rb").read()
    assert raw[:4] == b"MThd" and struct.unpack(">IHHH", raw[4:14]) == (6, 1, 2, 480)
    assert raw.count(b"MTrk") == 2 and raw.endswith(b"\x00\xff\x2f\x00")
    assert b"\xff\x58\x04\x04\x02\x18\x08" in raw                                  # 4/4, metronome 24, 8


class _FakeModel:
    def __init__(self):
        self.stop_training = False
        self.saved = []

    def save_weights(self, path, overwrite=True):
        self.saved.append(path)


def test_early_stopping_is_registered_twice_like_the_reference(tmp_path):
    args = types.SimpleNamespace(model_dir=str(tmp_path), run_name="r")
    cbs = model_utils.get_callbacks(args, patience=5, min_epoch=1)
    assert len(cbs) == 3 and cbs[1] is cbs[2]                                      # quirk Q5
    m = _FakeModel()
    for cb in cbs:
        cb.set_model(m); cb.on_train_begin({})
    losses = [5.0, 4.0, 4.1, 4.2, 4.3, 4.4, 4.5]
    stopped = None
    for ep, l in enumerate(losses):
        for cb in cbs:
            cb.on_epoch_end(ep, {"val_loss": l})
        if m.stop_training:
            stopped = ep
            break
    # wait advances 2/epoch: with patience 5 training stops on the 3rd non-improving epoch
    assert stopped == 4
    # checkpoint: best only, never before min_epoch (epoch 0 is not saved even though it is the first)
    assert m.saved == [os.path.join(str(tmp_path), "r.h5")]


def test_anneal_loss_weight_linear_schedule(capsys):
    v = Variable(0.1)
    cb = model_utils.AnnealLossWeight(v, name="kl_weight", final_value=1.0, n_epochs=4)
    vals = []
    for ep in range(6):
        cb.on_epoch_begin(ep)
        vals.append(float(v))
    assert np.allclose(vals, [0.1, 0.325, 0.55, 0.775, 1.0, 1.0])
    assert "+++++ kl_weight" in capsys.readouterr().out


def test_cli_surface_matches_reference_flags():
    from clvae_b200.cl_vrnn import train as vt, sample as vs
    from clvae_b200.cl_vae import train as at, sample as asmp
    a = vt.build_parser().parse_args(["run"])
    assert (a.batch_size, a.optimizer, a.num_epochs, a.original_dim, a.latent_dim, a.intermediate_dim,
            a.seq_length, a.class_weight, a.patience, a.kl_anneal, a.w_kl_anneal, a.w_log_var_prior) == \
           (200, "adam-wn", 200, 88, 2, 88, 16, 1.0, 5, 0, 0, 0.0)
    assert a.train_file == "../data/input/JSB Chorales_Cs.pickle" and a.model_dir == "../data/models"
    assert not a.use_x_prev and not a.predict_next and not a.do_log
    b = at.build_parser().parse_args(["run1", "--use_x_prev", "--latent_dim", "4"])         # README example
    assert (b.batch_size, b.seq_length, b.intermediate_class_dim, b.latent_dim, b.use_x_prev) == (100, 1, 88, 4, True)
    s = vs.build_parser().parse_args(["out", "--model_file", "m.h5", "--infer_w", "-t", "16", "-n", "3", "-c", "C"])
    assert (s.model_file, s.infer_w, s.discrete_w, s.t, s.n, s.c) == ("m.h5", True, False, 16, 3, "C")
    s2 = asmp.build_parser().parse_args(["out", "-i", "m.h5", "--use_z_prior", "--no_x_prev"])
    assert s2.model_file == "m.h5" and s2.use_z_prior and s2.no_x_prev and s2.t == 32 and s2.n == 1


def test_shard_range_partitions_exactly():
    from clvae_b200.parallel import shard_range
    for n in (0, 1, 7, 200, 100000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
