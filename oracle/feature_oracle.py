"""CPU oracle: NumPy restatement of the reference audio feature front-end.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product package never does;
it fails loudly when its CUDA library is missing.

PARITY STATUS: **unpinned by the reference's own tests** (it ships none, and its single golden
artefact ``dataset/data/20241204_MySlate_166/audio_features.csv`` is not mounted).  The heavy
arithmetic lives in ``librosa`` (unpinned, not installable here), restated in
``oracle/librosa_standin``.  What *is* pinned: ``oracle/make_golden.py`` runs the reference's own
``.py`` files verbatim from ``/root/reference`` on top of the stand-in and asserts that every
function below reproduces them bit-for-bit; the resulting vectors are committed under
``tests/golden/``.  The stand-in itself is cross-checked against torchaudio / transformers / scipy
(``tests/test_oracle_standin.py``).

Every function cites the reference lines it follows (paths relative to ``/root/reference``).
The per-frame ``np.correlate`` loop is kept on purpose: it is the reference's cost signature and
this module doubles as the CPU baseline ("port") in ``bench.py``.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STANDIN = os.path.join(_HERE, "librosa_standin")


def _librosa():
    """Import the stand-in (or a real librosa, should one ever be installed)."""
    try:
        import librosa  # noqa: F401
    except ModuleNotFoundError:
        sys.path.insert(0, _STANDIN)
        import librosa  # noqa: F401
    return librosa


# ----------------------------------------------------------------------------------------------
# integer frame maths (must be bit-exact) -- utils/audio/extraction/extract_features.py:12-16
# ----------------------------------------------------------------------------------------------
MIN_FRAMES = 9


def frame_params(sr):
    """extract_features.py:12-13 -- note int(0.01667 * sr), not sr / 60."""
    frame_length = int(0.01667 * sr)
    return frame_length, frame_length // 2


def guard_frames(n_samples, frame_length, hop_length):
    """extract_features.py:16 -- un-padded frame count used only for the too-short guard."""
    return (n_samples - frame_length) // hop_length + 1


def hop_frames(n_samples, frame_length, hop_length):
    """T shared by both branches: pad F//2 each side, then 1 + (len - F) // H."""
    return 1 + (n_samples + 2 * (frame_length // 2) - frame_length) // hop_length


def feature_rows(n_samples, frame_length, hop_length):
    t = hop_frames(n_samples, frame_length, hop_length)
    return (t + 1) // 2


# ----------------------------------------------------------------------------------------------
# extract_features_utils.py
# ----------------------------------------------------------------------------------------------
def cmvn(coeffs):
    """extract_features_utils.py:5-8 -- per-coefficient mean / population std over the whole clip."""
    mu = np.mean(coeffs, axis=1, keepdims=True)
    sigma = np.std(coeffs, axis=1, keepdims=True)
    return (coeffs - mu) / (sigma + 1e-10)


def pair_reduce(channel_major):
    """extract_features_utils.py:33-44 -- mean of frame pairs; an odd last frame passes through."""
    n = channel_major.shape[1]
    even = n // 2 * 2
    out = channel_major[:, :even].reshape(channel_major.shape[0], -1, 2).mean(axis=2)
    if n % 2 == 1:
        out = np.hstack((out, channel_major[:, -1].reshape(-1, 1)))
    return out


def mfcc_block(y, sr, frame_length, hop_length, num_mfcc=23, include_deltas=True,
               include_cepstral=True):
    """extract_features_utils.py:17-30 -- MFCC -> CMVN -> delta, delta-delta -> (69, T)."""
    lr = _librosa()
    c = lr.feature.mfcc(y=y, sr=sr, n_mfcc=num_mfcc, n_fft=frame_length, hop_length=hop_length)
    if include_cepstral:
        c = cmvn(c)
    if not include_deltas:
        return c
    d1 = lr.feature.delta(c)
    d2 = lr.feature.delta(c, order=2)
    return np.vstack([c, d1, d2])


def mfcc_rows(y, sr, frame_length, hop_length, num_mfcc=23):
    """extract_features_utils.py:11-15 -- returns ((R, 69) float32, T)."""
    full = mfcc_block(y, sr, frame_length, hop_length, num_mfcc)
    return pair_reduce(full).T, full.shape[1]


def fix_edge_frames(ac, zero_threshold=1e-7):
    """extract_features_utils.py:105-113 -- near-silent first/last frame copies its neighbour."""
    if np.all(np.abs(ac[:, 0]) < zero_threshold):
        ac[:, 0] = ac[:, 1]
    if np.all(np.abs(ac[:, -1]) < zero_threshold):
        ac[:, -1] = ac[:, -2]
    return ac


def autocorr_block(y, sr, frame_length, hop_length, num_lags=187, pad_signal=True,
                   padding_mode="reflect", trim_padded=False):
    """extract_features_utils.py:54-102 -> (187, T) float64.  Defaults = the reference's only call
    (:119-121: reflect pad, no trimming); the three knobs follow :56-61 and :67-74."""
    lr = _librosa()
    pad = frame_length // 2
    padded = np.pad(y, pad_width=pad, mode=padding_mode) if pad_signal else y  # :56-61
    cols = lr.util.frame(padded, frame_length=frame_length, hop_length=hop_length)  # :64
    if pad_signal and trim_padded:                                            # :67-74
        start = np.arange(cols.shape[1]) * hop_length
        cols = cols[:, np.where((start >= pad) & (start + frame_length <= len(y) + pad))[0]]
    cols = cols - np.mean(cols, axis=0, keepdims=True)                        # :76
    cols = cols * np.hanning(frame_length)[:, np.newaxis]                      # :79-80 (-> float64)
    zero_lag = frame_length - 1
    out = []
    for col in cols.T:                                                        # :83-92
        full = np.correlate(col, col, mode="full")
        keep = full[zero_lag: zero_lag + num_lags + 1]
        if keep[0] != 0:
            keep = keep / keep[0]
        out.append(keep)
    ac = np.array(out).T[1:, :]                                                # :95-98
    return fix_edge_frames(ac)                                                 # :100


def autocorr_rows(y, sr, frame_length, hop_length, include_deltas=False):
    """extract_features_utils.py:116-128 -> (R, 187) float64 (561 columns with deltas)."""
    ac = autocorr_block(y, sr, frame_length, hop_length)
    if include_deltas:                                                        # :131-135
        lr = _librosa()
        ac = np.vstack([ac, lr.feature.delta(ac), lr.feature.delta(ac, order=2)])
    return pair_reduce(ac).T


def smooth_rows(rows):
    """extract_features_utils.py:47-51 -- row i <- mean(original row i-1, original row i)."""
    out = np.copy(rows)
    if len(rows) > 1:
        out[1:] = (rows[:-1] + rows[1:]) / 2
    return out


# ----------------------------------------------------------------------------------------------
# extract_features.py
# ----------------------------------------------------------------------------------------------
def extract_and_combine_features(y, sr, frame_length, hop_length, apply_smoothing=False,
                                 include_autocorr=True):
    """extract_features.py:26-46 -> (R, 69 | 256); float64 once the autocorr block is stacked."""
    blocks = [mfcc_rows(y, sr, frame_length, hop_length)[0]]
    if include_autocorr:
        blocks.append(autocorr_rows(y, sr, frame_length, hop_length))
    rows = np.hstack(blocks)
    return smooth_rows(rows) if apply_smoothing else rows


def peak_normalize(y):
    """utils/audio/load_audio.py:12-14."""
    peak = np.max(np.abs(y))
    return y / peak if peak > 0 else y


def extract_audio_features_from_array(y, sr):
    """extract_features.py:12-24 with the loader replaced by an in-memory, already decoded clip."""
    y = peak_normalize(np.asarray(y, dtype=np.float32))
    frame_length, hop_length = frame_params(sr)
    n = guard_frames(len(y), frame_length, hop_length)
    if n < MIN_FRAMES:
        print(f"Audio file is too short: {n} frames, required: {MIN_FRAMES} frames")
        return None, None
    return extract_and_combine_features(y, sr, frame_length, hop_length), y


# ----------------------------------------------------------------------------------------------
# dataset/data_processing.py -- collect_features family
# ----------------------------------------------------------------------------------------------
def interpolate_slower(data):
    """data_processing.py:84-106 -- (N, C) -> (2N-1, C): originals on even rows, midpoints on odd."""
    n, c = data.shape
    out = np.zeros((2 * n - 1, c))
    out[0::2] = data
    out[1::2] = (data[:-1] + data[1:]) / 2.0
    return out


def smooth_facial_data(facial):
    """data_processing.py:201-204."""
    out = np.copy(facial)
    out[1:] = (facial[:-1] + facial[1:]) / 2
    return out


def stack_with_blend(sequences, blend_frames):
    """data_processing.py:179-197 -- inclusive linspace cross-fade over min(blend, len, len) rows."""
    if not sequences:
        return None
    acc = sequences[0]
    for seq in sequences[1:]:
        n = min(blend_frames, acc.shape[0], seq.shape[0])
        if n <= 0:
            acc = np.vstack([acc, seq])
            continue
        w_out = np.linspace(1, 0, n).reshape(n, 1)
        w_in = np.linspace(0, 1, n).reshape(n, 1)
        acc = np.vstack([acc[:-n], w_out * acc[-n:] + w_in * seq[:n], seq[n:]])
    return acc


def match_lengths(audio_rows, facial_rows):
    """data_processing.py:126-145 -- centre-trim the longer stream, then clip both to the minimum."""
    la, lf = len(audio_rows), len(facial_rows)
    if la > lf:
        left = (la - lf) // 2
        audio_rows = audio_rows[left: la - ((la - lf) - left)]
    elif lf > la:
        left = (lf - la) // 2
        facial_rows = facial_rows[left: lf - ((lf - la) - left)]
    m = min(len(audio_rows), len(facial_rows))
    return audio_rows[:m], facial_rows[:m]


def collect_from_arrays(audio_rows, facial_rows, include_fast=True, include_slow=False,
                        blend_boundaries=True, blend_frames=30):
    """data_processing.py:126-177 with the CSV I/O (:112-123) stripped: arrays in, arrays out."""
    audio_rows, facial_rows = match_lengths(audio_rows, facial_rows)
    a_versions, f_versions = [audio_rows], [facial_rows]
    if include_fast:                                                           # :152-158
        a_versions.append(audio_rows[::2].copy())
        f_versions.append(facial_rows.copy()[::2].copy())
    if include_slow:                                                           # :161-167
        a_versions.append(interpolate_slower(audio_rows))
        f_versions.append(smooth_facial_data(interpolate_slower(facial_rows)))
    if blend_boundaries:                                                       # :170-175
        return stack_with_blend(a_versions, blend_frames), stack_with_blend(f_versions, blend_frames)
    return np.vstack(a_versions), np.vstack(f_versions)


def collected_rows(n_rows, include_fast=True, include_slow=False, blend_boundaries=True,
                   blend_frames=30):
    """Row count of collect_from_arrays for a matched length ``n_rows`` (integer arithmetic)."""
    total = n_rows
    for extra in ([(n_rows + 1) // 2] if include_fast else []) + \
                 ([2 * n_rows - 1] if include_slow else []):
        n = min(blend_frames, total, extra) if blend_boundaries else 0
        total = total + extra - max(n, 0)
    return total


# ----------------------------------------------------------------------------------------------
# dataset/dataset.py -- stride-1 windowing (the "next" row, section 8(f)-1)
# ----------------------------------------------------------------------------------------------
def window_examples(audio_rows, facial_rows, window=128):
    """dataset/dataset.py:58-98 -> list of (float32[window, Ca], float32[window, Cf]) arrays."""
    na, nf = len(audio_rows), len(facial_rows)
    top = max(na, nf)
    out = []
    for start in range(0, top - window + 1):                                   # :66-75
        a = np.zeros((window, audio_rows.shape[1]))
        f = np.zeros((window, facial_rows.shape[1]))
        a[:min(window, na - start)] = audio_rows[start:start + window]
        f[:min(window, nf - start)] = facial_rows[start:start + window]
        out.append((a.astype(np.float32), f.astype(np.float32)))
    if top % window != 0:                                                      # :77-96
        start = top - window
        seg_a, seg_f = audio_rows[start:top], facial_rows[start:top]
        a = np.zeros((window, audio_rows.shape[1]))
        f = np.zeros((window, facial_rows.shape[1]))
        a[:len(seg_a)] = seg_a
        a[len(seg_a):] = np.flip(seg_a, axis=0)[:window - len(seg_a)]
        f[:len(seg_f)] = seg_f
        f[len(seg_f):] = np.flip(seg_f, axis=0)[:window - len(seg_f)]
        out.append((a.astype(np.float32), f.astype(np.float32)))
    return out
