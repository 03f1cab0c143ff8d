"""PianoData re-host vs golden values produced by the reference's own utils/pianoroll.py
(tests/golden/make_golden.py), on the two bundled JSB pickles."""
import hashlib
import json
import os
import numpy as np
import pytest

import clvae_b200  # noqa: F401
from clvae_b200.utils import pianoroll as pr

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(os.path.dirname(HERE), "data", "input")
CFGS = {
    "vrnn_train": dict(batch_size=200, seq_length=16, step_length=1, return_y_next=True,
                       return_y_hist=True, squeeze_x=False, squeeze_y=False),
    "vae_train": dict(batch_size=100, seq_length=1, step_length=1, return_y_next=True,
                      squeeze_x=True, squeeze_y=True),
    "vrnn_sample": dict(batch_size=1, seq_length=32, squeeze_x=False),
    "vae_sample": dict(batch_size=1, seq_length=32, squeeze_x=True),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


GOLD = json.load(open(os.path.join(HERE, "golden", "pianodata.json")))


@pytest.mark.parametrize("key", sorted(GOLD))
def test_pianodata_matches_reference_golden(key):
    fn, cname = key.split("::")
    P = pr.PianoData(os.path.join(DATA, fn), **CFGS[cname])
    rec = GOLD[key]
    assert {k: int(v) for k, v in P.key_map.items()} == rec["key_map"]
    for split in ("train", "valid", "test"):
        x, y = getattr(P, "x_" + split), getattr(P, "y_" + split)
        r = rec[split]
        assert list(x.shape) == r["x_shape"] and list(y.shape) == r["y_shape"]
        assert sha(x.astype(np.uint8)) == r["x_sha"] and sha(y.astype(np.uint8)) == r["y_sha"]
        assert sha(getattr(P, split + "_song_keys").astype(np.int64)) == r["keys_sha"]
        assert sha(getattr(P, split + "_song_inds").astype(np.int64)) == r["inds_sha"]
        assert sha(np.asarray(getattr(P, split + "_song_modes")).astype(np.uint8)) == r["modes_sha"]


def test_device_rolls_equal_materialised_windows():
    fn = os.path.join(DATA, "JSB Chorales_all.pickle")
    P = pr.PianoData(fn, **CFGS["vrnn_train"])
    DR, key_map = pr.DeviceRolls.from_pickle(fn, "train", 17, 200)
    assert key_map == P.key_map
    assert len(DR.win_off) == len(P.x_train)
    idx = np.r_[0:50, len(DR.win_off) - 50:len(DR.win_off)]
    w = np.stack([DR.roll[o:o + 17] for o in DR.win_off[idx]])
    assert np.array_equal(w[:, 1:], P.y_train[idx]) and np.array_equal(w[:, :-1], P.x_train[idx])
    assert np.array_equal(DR.labels, P.train_song_keys)
    # the sampler window (33) drops short songs: label mis-alignment Q6 is reproduced
    P2 = pr.PianoData(fn, **CFGS["vrnn_sample"])
    DR2, _ = pr.DeviceRolls.from_pickle(fn, "test", 33, 1)
    assert np.array_equal(DR2.labels, P2.test_song_keys)
