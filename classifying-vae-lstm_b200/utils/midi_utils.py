"""
Piano-roll -> Standard MIDI File (format 1) writer: re-host of utils/midi_utils.py, which relies on
the third-party `midi` package (python-midi, unpinned, absent here).  Same event logic as
MidiWriter.dump_sequence_to_midi (utils/midi_utils.py:26-98): resolution 480, 120 ticks per frame,
velocity 100, pitch = key index + 21, a 4/4 time-signature meta track (metronome 24, 8
thirty-seconds), note-offs before note-ons within a frame, only the first event of a frame carries
the frame's delta, trailing flush of sounding notes.  The SMF bytes are produced here directly
(python-midi's writer: running status is not used, every track ends with End-of-Track).
Byte-level parity with python-midi is unpinned (third-party, not installable).
"""
import os
import struct
import numpy as np

RANGE = 128


def _varlen(v):
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


class MidiWriter(object):
    def __init__(self, verbose=False, default_vel=100):
        self.verbose = verbose
        self.note_range = RANGE
        self.default_velocity = default_vel

    def note_off(self, val, tick):
        self.track.append((tick, bytes([0x80, val, 0])))
        return 0

    def note_on(self, val, tick):
        self.track.append((tick, bytes([0x90, val, self.default_velocity])))
        return 0

    def events(self, seq, time_step=120, offset=21):
        """The (delta_tick, message) list of the note track -- the part tests pin down."""
        sequence = np.asarray(seq)
        self.track = []
        tick = time_step
        notes_on = {n: False for n in range(self.note_range)}
        for seq_idx in range(sequence.shape[0]):
            notes = [int(n) + offset for n in np.nonzero(sequence[seq_idx, :])[0].tolist()]
            for n in notes_on:
                if notes_on[n] and n not in notes:
                    tick = self.note_off(n, tick)
                    notes_on[n] = False
            for note in notes:
                if not notes_on[note]:
                    tick = self.note_on(note, tick)
                    notes_on[note] = True
            tick += time_step
        for n in notes_on:
            if notes_on[n]:
                self.note_off(n, tick)
                tick = 0
                notes_on[n] = False
        return self.track

    def dump_sequence_to_midi(self, seq, output_filename, time_step=120, resolution=480,
                              metronome=24, offset=21, format='final'):
        if format == 'flat':
            seq = np.reshape(seq, [-1, self.note_range])
        elif format == 'icml':
            seq = np.array([[1 if i in tm else 0 for i in range(self.note_range)] for tm in seq])
        ev = self.events(seq, time_step, offset)

        def chunk(events):
            body = b"".join(_varlen(dt) + msg for dt, msg in events) + b"\x00\xff\x2f\x00"
            return b"MTrk" + struct.pack(">I", len(body)) + body
        meta = [(0, bytes([0xFF, 0x58, 4, 4, 2, metronome, 8]))]   # 4/4: denominator stored as log2
        data = b"MThd" + struct.pack(">IHHH", 6, 1, 2, resolution) + chunk(meta) + chunk(ev)
        with open(output_filename, "wb") as f:
            f.write(data)


def write_sample(sample, outdir, fnm, isHalfAsSlow=False):
    """utils/midi_utils.py:100-104 (JSB samples are frame-doubled)."""
    if isHalfAsSlow:
        sample = np.repeat(sample, 2, axis=0)
    fnm = os.path.join(outdir, fnm + '.mid')
    MidiWriter().dump_sequence_to_midi(sample, fnm)
    return fnm
