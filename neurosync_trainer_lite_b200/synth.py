"""Deterministic synthetic waveforms and facial rows for the benchmark configurations.

SURVEY.md section 8(d): a voiced-like FM/AM tone + noise ("voiced"), white noise ("noise") and a
gated burst signal ("gated") that exercises the top-dB floor and near-silent frames.  Every signal
is peak-normalised the way ``utils/audio/load_audio.py:12-14`` of the reference does, so it can be
fed straight to ``extract_and_combine_features``.  Pure NumPy host code; no GPU involved.
"""
import io
import wave

import numpy as np

KINDS = ("voiced", "noise", "gated")


def peak_normalize(y):
    """y / max|y| in float32 when the peak is > 0 (reference load_audio.py:12-14)."""
    y = np.asarray(y, dtype=np.float32)
    peak = np.max(np.abs(y)) if y.size else np.float32(0)
    return y / peak if peak > 0 else y


def synth_clip(seconds, sr, seed=0, kind="voiced", normalize=True):
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    if kind == "voiced":
        f = 120.0 + 30.0 * np.sin(2 * np.pi * 0.7 * t)
        y = 0.5 * np.sin(2 * np.pi * f * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t))
        y = y + 0.05 * rng.standard_normal(n)
    elif kind == "noise":
        y = rng.standard_normal(n)
    elif kind == "gated":
        gate = (np.sin(2 * np.pi * 1.3 * t) > 0).astype(np.float64)
        y = rng.standard_normal(n) * gate * np.abs(np.sin(2 * np.pi * 5.0 * t)) ** 4
        y = y + 1e-4 * rng.standard_normal(n)
    else:
        raise ValueError(f"unknown kind {kind!r}")
    y = y.astype(np.float32)
    return peak_normalize(y) if normalize else y


def synth_facial(rows, seed=0, cols=61):
    """Blendshape-like rows in [0, 1] with slow temporal structure (float64 like the CSVs)."""
    rng = np.random.default_rng(10_000 + seed)
    t = np.arange(rows, dtype=np.float64)[:, None] / 60.0
    freq = rng.uniform(0.2, 3.0, size=(1, cols))
    phase = rng.uniform(0, 2 * np.pi, size=(1, cols))
    return 0.5 + 0.4 * np.sin(2 * np.pi * freq * t + phase) + 0.02 * rng.standard_normal((rows, cols))


def to_int16_pcm(y):
    """Quantise to int16 the way an ffmpeg/WAV writer would (round, clip)."""
    return np.clip(np.rint(np.asarray(y, dtype=np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def wav_bytes(pcm16, sr):
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(np.asarray(pcm16, dtype="<i2").tobytes())
    return buf.getvalue()
