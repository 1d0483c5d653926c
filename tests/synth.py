"""Synthetic workload audio, shared by the tests, bench.py and the tools.  It is workload DATA, not a checker: it lives
outside oracle/ so that the measured product path never imports the oracle."""
import numpy as np


def synth_clip_labeled(seed, seconds=30.0, sr=22050):
    """synth_clip plus its ground truth: -> (audio float32, notes float64 [n, 3] = (start s, end s, MIDI pitch))."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    y = np.zeros(n, dtype=np.float64)
    detune = rng.uniform(-20.0, 20.0) / 100.0
    voices = int(rng.integers(1, 7))
    t_all = np.arange(n) / sr
    notes = []
    for _ in range(voices):
        pos = 0
        while pos < n:
            dur = int(rng.uniform(0.25, 0.5) * sr)
            end = min(n, pos + dur)
            midi = int(rng.integers(36, 97))
            notes.append((pos / sr, end / sr, midi))
            f0 = 440.0 * 2.0 ** ((midi - 69 + detune) / 12.0)
            tt = t_all[pos:end] - t_all[pos]
            env = np.minimum(1.0, np.minimum(tt / 0.01, (tt[-1] - tt) / 0.02 + 1e-3))
            ph = rng.uniform(0, 2 * np.pi, size=8)
            for k in range(1, 9):
                if f0 * k < 0.45 * sr:
                    y[pos:end] += env * np.sin(2 * np.pi * f0 * k * tt + ph[k - 1]) / k
            pos = end
    y += rng.standard_normal(n) * (10 ** (-40 / 20)) * np.max(np.abs(y) + 1e-9)
    y *= 0.5 / np.max(np.abs(y))
    return y.astype(np.float32), np.asarray(notes, dtype=np.float64)


def synth_clip(seed, seconds=30.0, sr=22050):
    """Seeded polyphonic test clip: 1-6 voices of 8-partial harmonic tones, notes of 0.25-0.5 s, MIDI 36-96,
    a random +-20 cent clip detune, white noise at -40 dB, peak-normalised to 0.5, float32."""
    return synth_clip_labeled(seed, seconds, sr)[0]


def piano_roll(notes, n_frames, fps=22050 / 512, min_pitch=24, n_out=72):
    """[n_frames, n_out] float32 frame labels of `notes` (a frame is active when its centre time lies inside the note)."""
    roll = np.zeros((n_frames, n_out), dtype=np.float32)
    for s, e, p in notes:
        a, b = int(np.ceil(s * fps)), int(np.ceil(e * fps))
        if 0 <= int(p) - min_pitch < n_out:
            roll[max(a, 0):min(b, n_frames), int(p) - min_pitch] = 1.0
    return roll
