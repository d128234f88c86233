"""GPU-backed mirror of the reference's ``dataset/data_processing.py`` (dataset builders).

Host side (unchanged semantics): folder scan, CSV cache of ``audio_features.csv``, the facial CSV
minus ``Timecode`` / ``BlendshapeCount``, ``facial[:, :61] *= 100``.
Device side: feature extraction (``extract_audio_features``) and the whole ``collect_features``
augmentation - centre-trim, fast ``[::2]``, slow ``interpolate_slower`` (+ ``smooth_facial_data``),
``stack_with_blend`` - as one ``nsf_collect_host`` call in float64, bit-identical to the NumPy
arithmetic of the reference (same operation order, IEEE add/mul, no FMA contraction).

``load_data_batched`` is the B200-first builder: all clips of a dataset in one extraction batch and
one collect batch, optionally sharded by clip over ranks (``shard.py``).
"""
import os

import numpy as np
import pandas as pd

from .. import _native as nv
from .. import engine as _engine
from ..utils.audio.extraction.extract_features import extract_audio_features
from ..utils.video.mov_extraction import find_files, get_audio

COLUMNS_TO_DROP = ['Timecode', 'BlendshapeCount']


def _engine_any():
    f, h = _engine.frame_params(88200)
    return _engine.get_engine(88200, f, h)


def _rows64(a):
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def load_data(root_dir, sr, processed_folders):
    """reference :10-26 -- ``os.listdir`` order, skips (and records) processed folders."""
    examples = []
    for folder in os.listdir(root_dir):
        folder_path = os.path.join(root_dir, folder)
        if os.path.isdir(folder_path) and folder not in processed_folders:
            audio_features, facial_data = process_folder(folder_path, sr)
            if audio_features is not None and facial_data is not None:
                examples.append((audio_features, facial_data))
                processed_folders.add(folder)
    return examples


def scale_facial_data(facial_data, scale_factor=1.1):
    """reference :28-41 (unused helper; host NumPy): scale then clip to [-1, 1]."""
    return np.clip(np.asarray(facial_data) * scale_factor, -1, 1)


def process_folder(folder_path, sr, apply_smoothing=False, apply_over_scale=False):
    """reference :44-78."""
    mov_path, mp4_path, wav_path, facial_csv_path, audio_features_csv_path, _ = find_files(folder_path)
    video_path = mov_path or mp4_path
    have_cache = os.path.exists(audio_features_csv_path)
    if not (facial_csv_path and (video_path or wav_path or have_cache)):
        return None, None
    audio_path = get_audio(video_path, wav_path, folder_path) if (video_path or wav_path) else None
    if not (audio_path or have_cache):
        return None, None
    audio_features, facial_data = collect_features(audio_path if audio_path else _, audio_features_csv_path,
                                                   facial_csv_path, sr)
    if apply_over_scale:
        facial_data = scale_facial_data(facial_data)
    facial_data[:, :61] *= 100                                            # :68
    if apply_smoothing:
        facial_data = smooth_facial_data(facial_data)
    return audio_features, facial_data


def interpolate_slower(data):
    """reference :84-106 -- ``(N, F) -> (2N-1, F)``: originals on even rows, midpoints on odd rows."""
    data = _rows64(data)
    if data.shape[0] == 0:
        raise ValueError("interpolate_slower needs at least one row")
    out = _engine_any().rows_op(nv.ROWS_INTERP_SLOWER, data)
    return out.astype(np.float64, copy=False)  # the reference allocates np.zeros (float64)


def collect_arrays(audio_features, facial_data, include_fast=True, include_slow=False,
                   blend_boundaries=True, blend_frames=30):
    """The arithmetic of collect_features (reference :126-177) for one clip, on the device."""
    a = _rows64(audio_features)
    f = np.ascontiguousarray(facial_data, dtype=a.dtype)
    if not (include_fast or include_slow) or min(len(a), len(f)) == 0:
        # no augmentation: only the (host) length matching remains
        n = min(len(a), len(f))
        la = (len(a) - n) // 2 if len(a) > len(f) else 0
        lf = (len(f) - n) // 2 if len(f) > len(a) else 0
        return a[la:la + n], f[lf:lf + n]
    out_a, out_f, _ = _engine_any().collect_host(a, [0, len(a)], f, [0, len(f)], include_fast,
                                                 include_slow, blend_boundaries, blend_frames)
    return out_a, out_f


def collect_features(audio_path, audio_features_csv_path, facial_csv_path, sr,
                     include_fast=True, include_slow=False, blend_boundaries=True, blend_frames=30):
    """reference :108-177 -- same CSV cache side effects, same return shapes and dtypes."""
    if os.path.exists(audio_features_csv_path):                           # :112-114
        print(f"Loading audio features from {audio_features_csv_path}")
        audio_features = pd.read_csv(audio_features_csv_path).values
    else:                                                                 # :115-120
        print(f"Extracting audio features from {audio_path}")
        audio_features, _ = extract_audio_features(audio_path, sr)
        if audio_features is not None:
            pd.DataFrame(audio_features).to_csv(audio_features_csv_path, index=False)
            print(f"Audio features saved to {audio_features_csv_path}")
    facial_data = pd.read_csv(facial_csv_path).drop(columns=COLUMNS_TO_DROP).values   # :123
    if audio_features is None:
        # the reference fails here with TypeError: object of type 'NoneType' has no len() (:143)
        raise TypeError("object of type 'NoneType' has no len()")
    return collect_arrays(audio_features, facial_data, include_fast, include_slow, blend_boundaries,
                          blend_frames)


def stack_with_blend(sequences, blend_frames):
    """reference :179-197 -- inclusive-linspace cross-fade over ``min(blend, len, len)`` rows."""
    if not sequences:
        return None
    result = _rows64(sequences[0])
    eng = _engine_any()
    for seq in sequences[1:]:
        seq = np.ascontiguousarray(seq, dtype=result.dtype)
        if len(seq) == 0:
            continue
        if len(result) == 0:
            result = seq
            continue
        result = eng.rows_op(nv.ROWS_BLEND_STACK, result, seq, blend_frames)
    return result


def smooth_facial_data(facial_data):
    """reference :201-204 -- rows 1.. <- mean of (row i-1, row i), from the original rows."""
    x = _rows64(facial_data)
    if len(x) == 0:
        return x.copy()
    return _engine_any().rows_op(nv.ROWS_SMOOTH, x)


def remove_specified_dimensions(facial_data):
    """reference :208-212 (unused helper; host NumPy)."""
    cols = list(range(14)) + list(range(51, 61))
    return np.delete(facial_data, cols, axis=1)


def zero_specified_columns(facial_data):
    """reference :214-220 (unused helper; host NumPy, in place)."""
    cols = list(range(14)) + list(range(51, 61))
    facial_data[:, cols] = 0
    return facial_data


# ------------------------------------------------------------------------------------------------
# B200-first batched builder
# ------------------------------------------------------------------------------------------------
def collect_batch(audio_rows, facial_rows, include_fast=True, include_slow=False,
                  blend_boundaries=True, blend_frames=30, device=None, dtype=np.float32):
    """collect_features arithmetic for MANY clips in one device call.

    ``audio_rows`` / ``facial_rows``: lists of ``[N_i, 256]`` / ``[M_i, 61]`` arrays.  Returns
    ``(audio [sum, 256], facial [sum, 61], offsets)`` in ``dtype`` (float32 = training format,
    float64 = the reference's arithmetic bit for bit)."""
    f0, h0 = _engine.frame_params(88200)
    eng = _engine.get_engine(88200, f0, h0, device=device)
    a_off = np.zeros(len(audio_rows) + 1, dtype=np.int64)
    f_off = np.zeros(len(facial_rows) + 1, dtype=np.int64)
    np.cumsum([len(a) for a in audio_rows], out=a_off[1:])
    np.cumsum([len(f) for f in facial_rows], out=f_off[1:])
    a = np.concatenate([np.asarray(x, dtype=dtype) for x in audio_rows], axis=0)
    f = np.concatenate([np.asarray(x, dtype=dtype) for x in facial_rows], axis=0)
    return eng.collect_host(a, a_off, f, f_off, include_fast, include_slow, blend_boundaries,
                            blend_frames)
