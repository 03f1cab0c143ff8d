"""
Piano-roll pickle loader (host side) -- Python 3 re-host of utils/pianoroll.py of the reference.

Same constructor, attributes and quirks as `PianoData` (utils/pianoroll.py:73-158):
  * pickle schema: dict with train/valid/test = list of songs = list of frames = list of MIDI
    pitches, *_key (str), *_mode (bool); loaded with encoding='latin1' (the files are Python-2
    pickles, protocol 0/2, one of them embeds numpy scalars).
  * offset 21 with the +-12 adjustment (utils/pianoroll.py:37-42), `n - L` windows per song
    (sliding_inds, :49-50, quirk Q7), songs shorter than the window dropped BEFORE song indices are
    enumerated (:69-70, quirk Q6 -- labels are mis-aligned after the first dropped song, reproduced),
    tail trimmed to a multiple of batch_size (:154-158), relative-major key folding (:7-25,135-152).

Differences (documented): arrays are uint8 by default instead of float64 (values identical; pass
dtype=np.float64 for the reference dtype), and the windows are built with stride tricks instead of
dstack/swapaxes.  `DeviceRolls` is the B200-side form: one concatenated uint8 roll per split plus
a frame offset per window, so sliding windows are never materialised (the CUDA kernels gather rows
by offset).
"""
import pickle
import numpy as np

rel_keys = {'a': 'C', 'b-': 'D-', 'b': 'D', 'c': 'E-', 'c#': 'E', 'd-': 'F-', 'd': 'F', 'd#': 'F#',
            'e-': 'G-', 'e': 'G', 'f': 'A-', 'f#': 'A', 'g': 'B-', 'g#': 'B', 'a-': 'C-'}


def relative_major(k):
    return k if k.isupper() else rel_keys[k]


def pianoroll_to_song(roll, offset=21):
    return [(np.where(s)[0] + offset).tolist() for s in roll]


def song_to_pianoroll(song, offset=21, dtype=np.uint8):
    """[(60, 72, 79), (72, 79), ...] -> [n_frames, 88] binary roll (utils/pianoroll.py:31-47)."""
    all_notes = [y for x in song for y in x]
    if min(all_notes) - offset < 0:
        offset -= 12
    if max(all_notes) - offset > 87:
        offset += 12
    roll = np.zeros((len(song), 88), dtype=dtype)
    for t, notes in enumerate(song):
        roll[t, [int(n) - offset for n in notes]] = 1
    return roll


def sliding_inds(n, seq_length, step_length):
    return np.arange(n - seq_length, step=step_length)


def sliding_window(roll, seq_length, step_length=1):
    """[n_windows, seq_length, 88]; n - seq_length windows (not n - seq_length + 1)."""
    inds = sliding_inds(roll.shape[0], seq_length, step_length)
    if len(inds) == 0:
        return np.array([])
    v = np.lib.stride_tricks.sliding_window_view(roll, seq_length, axis=0)  # [n-L+1, 88, L]
    return np.ascontiguousarray(v[inds].transpose(0, 2, 1))


def songs_to_pianoroll(songs, seq_length, step_length, inner_fcn=song_to_pianoroll):
    rolls = [sliding_window(inner_fcn(s), seq_length, step_length) for s in songs]
    rolls = [r for r in rolls if len(r) > 0]
    inds = [i * np.ones((len(r),)) for i, r in enumerate(rolls)]
    return np.vstack(rolls), np.hstack(inds)


def load_pickle(train_file):
    with open(train_file, 'rb') as f:
        return pickle.load(f, encoding='latin1')


class PianoData:
    def __init__(self, train_file, batch_size=None, seq_length=1, step_length=1, return_y_next=True,
                 return_y_hist=False, squeeze_x=True, squeeze_y=True, use_rel_major=True,
                 dtype=np.uint8):
        D = load_pickle(train_file)
        self.train_file = train_file
        self.batch_size = batch_size
        self.seq_length = seq_length
        self.step_length = step_length
        self.return_y_next = return_y_next
        self.return_y_hist = return_y_hist
        self.squeeze_x = squeeze_x
        self.squeeze_y = squeeze_y
        self.use_rel_major = use_rel_major
        self.dtype = dtype

        self.x_train, self.y_train, self.train_song_inds = self.make_xy(D['train'])
        self.x_test, self.y_test, self.test_song_inds = self.make_xy(D['test'])
        self.x_valid, self.y_valid, self.valid_song_inds = self.make_xy(D['valid'])

        if 'train_mode' in D:
            self.train_song_modes = self.song_modes(D['train_mode'], self.train_song_inds)
            self.test_song_modes = self.song_modes(D['test_mode'], self.test_song_inds)
            self.valid_song_modes = self.song_modes(D['valid_mode'], self.valid_song_inds)
        if 'train_key' in D:
            D = self.update_keys(D)
            self.key_map = self.make_keymap(D)
            self.train_song_keys = self.song_keys(D['train_key'], self.train_song_inds)
            self.test_song_keys = self.song_keys(D['test_key'], self.test_song_inds)
            self.valid_song_keys = self.song_keys(D['valid_key'], self.valid_song_inds)

    def make_xy(self, songs):
        inner = lambda s: song_to_pianoroll(s, dtype=self.dtype)
        x_rolls, song_inds = songs_to_pianoroll(songs, self.seq_length + int(self.return_y_next),
                                                self.step_length, inner_fcn=inner)
        x_rolls = self.adjust_for_batch_size(x_rolls)
        song_inds = self.adjust_for_batch_size(song_inds)
        if self.return_y_next:
            y_rolls = x_rolls[:, 1:, :] if self.return_y_hist else x_rolls[:, -1, :]
            x_rolls = x_rolls[:, :-1, :]
        else:
            y_rolls = x_rolls
        if self.squeeze_x:
            x_rolls = x_rolls.squeeze()
        if self.squeeze_y:
            y_rolls = y_rolls.squeeze()
        return x_rolls, y_rolls, song_inds

    def song_modes(self, modes, song_inds):
        return np.array(modes)[song_inds.astype(int)]

    def update_keys(self, D):
        if not self.use_rel_major:
            return
        for s in ('train_key', 'test_key', 'valid_key'):
            D[s] = [relative_major(k) for k in D[s]]
        return D

    def make_keymap(self, D):
        all_keys = np.unique(np.hstack([D['train_key'], D['test_key'], D['valid_key']]))
        return dict(zip([str(k) for k in all_keys], range(len(all_keys))))

    def song_keys(self, keys, song_inds):
        key_inds = [self.key_map[k] for k in keys]
        return np.array(key_inds)[song_inds.astype(int)]

    def adjust_for_batch_size(self, items):
        if self.batch_size is None:
            return items
        mod = items.shape[0] % self.batch_size
        return items[:-mod] if mod > 0 else items


class DeviceRolls:
    """Device-side form of one split: `roll` uint8 [n_frames_total, 88] (songs concatenated) and
    `win_off` int32 [n_windows] = first frame of each sliding window, enumerated exactly as
    PianoData.make_xy does (same order, same tail trim, same dropped songs), plus the per-window key
    label.  Window i of length W is roll[win_off[i] : win_off[i]+W]; the CUDA GEMMs gather those
    rows themselves, so the 17x blow-up of the materialised windows never exists in HBM."""

    def __init__(self, songs, keys, key_map, window, batch_size=None, use_rel_major=True):
        rolls, offs, labels = [], [], []
        base = 0
        kept = 0
        for s in songs:
            r = song_to_pianoroll(s)
            n = r.shape[0] - window
            if n <= 0:
                continue  # dropped before enumeration (Q6): label index advances only for kept songs
            k = keys[kept]  # reproduces the reference's label mis-alignment after a dropped song
            k = relative_major(k) if use_rel_major else k
            rolls.append(r)
            offs.append(base + np.arange(n, dtype=np.int64))
            labels.append(np.full(n, key_map[k], dtype=np.int32))
            base += r.shape[0]
            kept += 1
        self.roll = np.ascontiguousarray(np.vstack(rolls))
        off = np.concatenate(offs)
        lab = np.concatenate(labels)
        if batch_size is not None and len(off) % batch_size:
            off, lab = off[:-(len(off) % batch_size)], lab[:-(len(lab) % batch_size)]
        self.win_off = off.astype(np.int32)
        self.labels = lab
        self.window = window

    @classmethod
    def from_pickle(cls, train_file, split, window, batch_size=None):
        D = load_pickle(train_file)
        all_keys = np.unique(np.hstack([[relative_major(k) for k in D[s + '_key']]
                                        for s in ('train', 'test', 'valid')]))
        key_map = dict(zip([str(k) for k in all_keys], range(len(all_keys))))
        return cls(D[split], D[split + '_key'], key_map, window, batch_size), key_map
