"""GPU-backed mirror of the reference's ``utils/audio/extraction/extract_features.py``.

``extract_audio_features`` / ``extract_and_combine_features`` keep the reference signatures and
return conventions (``(R, 256)`` float64 rows, the peak-normalised float32 signal, ``(None, None)``
plus the reference's message for clips under 9 frames).  ``extract_features_batch`` is the batched
form the dataset builders and ``bench.py`` use: many clips, one C-ABI call, float32 rows.
"""
import numpy as np

from .... import _native as nv
from .... import engine as _engine
from ..load_audio import decode, decode_for_path

MIN_FRAMES = 9  # reference extract_features.py:14


def _too_short(num_frames):
    print(f"Audio file is too short: {num_frames} frames, required: {MIN_FRAMES} frames")
    return None, None


def extract_audio_features(audio_input, sr=88200, from_bytes=False):
    """reference :6-24.  Decode on the host, then ONE device pass that peak-normalises
    (load_audio.py:12-14) and extracts; returns ``(features float64 [R, 256], y float32 [L])``."""
    if from_bytes:
        import io
        pcm, sr = decode(io.BytesIO(audio_input), sr)
    else:
        pcm, sr = decode_for_path(audio_input, sr)
    frame_length, hop_length = _engine.frame_params(sr)                 # :12-13
    eng = _engine.get_engine(sr, frame_length, hop_length)
    num_frames = eng.plan.guard_frames(len(pcm))                        # :16
    if num_frames < MIN_FRAMES:
        return _too_short(num_frames)
    rows, y = eng.extract_host(pcm, [0, len(pcm)], nv.PEAK_NORMALIZE, want_y=True)
    return rows.astype(np.float64), y


def extract_and_combine_features(y, sr, frame_length, hop_length, apply_smoothing=False,
                                 include_autocorr=True):
    """reference :26-46 -> ``[R, 69 | 256]``; float64 when the autocorrelation block is present
    (the reference's hstack of float32 and float64), float32 otherwise."""
    y = np.asarray(y)
    if y.dtype != np.int16:
        y = np.ascontiguousarray(y, dtype=np.float32)
    flags = (0 if include_autocorr else nv.NO_AUTOCORR) | (nv.SMOOTH if apply_smoothing else 0)
    rows = _engine.get_engine(sr, frame_length, hop_length).extract_host(y, [0, len(y)], flags)
    return rows.astype(np.float64) if include_autocorr else rows


def extract_features_batch(clips, sr, frame_length=None, hop_length=None, peak_normalize=False,
                           device=None, out=None, flags=0):
    """Batched extraction: ``clips`` is a list of 1-D float32/int16 arrays (or a packed array plus
    offsets as ``(packed, offsets)``).  Returns ``(rows float32 [sum R_i, cols], row_offsets)``.
    One ``nsf_extract_host`` call; H2D, kernels and D2H are pipelined inside the library."""
    if frame_length is None:
        frame_length, hop_length = _engine.frame_params(sr)
    eng = _engine.get_engine(sr, frame_length, hop_length, device=device)
    if isinstance(clips, tuple):
        packed, offsets = clips
    else:
        packed, offsets = _engine.pack_clips(clips)
    flags |= nv.PEAK_NORMALIZE if peak_normalize else 0
    rows = eng.extract_host(packed, offsets, flags, out=out)
    return rows, eng.row_offsets(offsets, flags)
