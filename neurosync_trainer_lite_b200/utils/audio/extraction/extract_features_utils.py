"""GPU-backed mirror of the reference's ``utils/audio/extraction/extract_features_utils.py``.

Same function names, argument meaning and array layouts (channel-major ``[C, T]`` intermediates,
``.T`` at the end) as the reference; every function is one call into the C ABI (``include/nsf.h``)
and therefore needs an sm_100 GPU.  Nothing is computed with NumPy here beyond transposes and
dtype casts at the boundary.

dtypes: the device computes in float32.  Functions whose reference counterpart returns float64
(everything downstream of ``np.hanning``) up-cast the result so ``.dtype`` matches the reference.
"""
import numpy as np

from .... import _native as nv
from .... import engine as _engine

_DEFAULT_LAGS = 187


def _engine_for(sr, frame_length, hop_length, num_mfcc=23, n_lags=_DEFAULT_LAGS):
    return _engine.get_engine(sr, frame_length, hop_length, n_mfcc=num_mfcc, n_lags=n_lags)


def _generic_engine():
    """Engine for the array helpers that do not depend on the audio geometry."""
    f, h = _engine.frame_params(88200)
    return _engine.get_engine(88200, f, h)


def _signal(y):
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("expected a mono 1-D signal")
    if y.dtype != np.int16:
        y = np.ascontiguousarray(y, dtype=np.float32)
    return y


def _extract(y, sr, frame_length, hop_length, flags, num_mfcc=23, n_lags=_DEFAULT_LAGS):
    y = _signal(y)
    eng = _engine_for(sr, frame_length, hop_length, num_mfcc, n_lags)
    return eng.extract_host(y, [0, len(y)], flags)


def cepstral_mean_variance_normalization(mfcc):
    """reference :5-8 -- ``(x - mean) / (std + 1e-10)`` per coefficient over all frames."""
    x = np.asarray(mfcc)
    out = _generic_engine().post(np.ascontiguousarray(x.T, dtype=np.float32), nv.POST_CMVN)
    return out.T.astype(x.dtype if x.dtype in (np.float32, np.float64) else np.float32, copy=False)


def extract_mfcc_features(y, sr, frame_length, hop_length, num_mfcc=23):
    """reference :11-15 -> ``([R, 3*num_mfcc] float32, T)``."""
    rows = _extract(y, sr, frame_length, hop_length, nv.NO_AUTOCORR, num_mfcc)
    return rows, int(_engine.get_plan(sr, frame_length, hop_length, num_mfcc).hop_frames(len(y)))


def extract_overlapping_mfcc(chunk, sr, num_mfcc, frame_length, hop_length, include_deltas=True,
                             include_cepstral=True):
    """reference :17-30 -> ``[3*num_mfcc | num_mfcc, T]`` float32, one column per hop-frame."""
    flags = nv.NO_AUTOCORR | nv.NO_REDUCE
    if not include_deltas:
        flags |= nv.NO_DELTAS
    if not include_cepstral:
        flags |= nv.NO_CMVN
    return np.ascontiguousarray(_extract(chunk, sr, frame_length, hop_length, flags, num_mfcc).T)


def reduce_features(features):
    """reference :33-44 -- mean of frame pairs; an odd last frame passes through."""
    x = np.asarray(features)
    out = _generic_engine().post(np.ascontiguousarray(x.T, dtype=np.float32), nv.POST_REDUCE)
    return out.T.astype(x.dtype if x.dtype in (np.float32, np.float64) else np.float32, copy=False)


def smooth_features(features):
    """reference :47-51 -- row i <- (row i-1 + row i) / 2, computed from the original rows."""
    x = np.asarray(features)
    if x.shape[0] == 0:
        return np.copy(x)
    return _generic_engine().rows_op(nv.ROWS_SMOOTH, x)


def extract_overlapping_autocorr(y, sr, frame_length, hop_length, num_autocorr_coeff=187,
                                 pad_signal=True, padding_mode="reflect", trim_padded=False):
    """reference :54-102 -> ``[num_autocorr_coeff, T]`` float64 (lags 1..n, edge frames fixed).

    The defaults run as one fused device pass (reflect indexing inside the kernel).  The other knob
    settings keep the arithmetic on the device and do the *index* work on the host exactly where the
    reference does it: ``np.pad(y, F // 2, mode=padding_mode)`` before the call (:57-59), the kernel
    then frames the given signal as is (``NSF_AC_NO_PAD``), and ``trim_padded`` selects the frame
    columns of :67-74 before the edge fix (:100)."""
    y = _signal(y)
    base = nv.NO_MFCC | nv.NO_REDUCE
    if pad_signal and padding_mode == "reflect" and not trim_padded:
        rows = _extract(y, sr, frame_length, hop_length, base, n_lags=num_autocorr_coeff)
        return np.ascontiguousarray(rows.T, dtype=np.float64)
    pad = frame_length // 2
    if y.dtype == np.int16:
        y = y.astype(np.float32) / np.float32(32768)
    y_padded = np.pad(y, pad_width=pad, mode=padding_mode) if pad_signal else y          # :56-61
    eng = _engine_for(sr, frame_length, hop_length, n_lags=num_autocorr_coeff)
    if pad_signal and trim_padded:
        # :67-74 drops frames that touch the padding BEFORE fix_edge_frames_autocorr sees them, so the fix
        # must run on the kept columns only: extract with the in-kernel fix disabled (threshold 0 can never
        # be undercut), select, then apply the fix as its own call
        eng.set_edge_zero_threshold(0.0)
        try:
            rows = eng.extract_host(np.ascontiguousarray(y_padded, dtype=np.float32), [0, len(y_padded)],
                                    base | nv.AC_NO_PAD)
        finally:
            eng.set_edge_zero_threshold(1e-7)
        start = np.arange(rows.shape[0]) * hop_length
        valid = np.where((start >= pad) & (start + frame_length <= len(y) + pad))[0]
        feats = np.ascontiguousarray(rows[valid].T, dtype=np.float64)
        return fix_edge_frames_autocorr(feats) if feats.shape[1] >= 2 else feats
    rows = eng.extract_host(np.ascontiguousarray(y_padded, dtype=np.float32), [0, len(y_padded)],
                            base | nv.AC_NO_PAD)
    return np.ascontiguousarray(rows.T, dtype=np.float64)


def fix_edge_frames_autocorr(autocorr_features, zero_threshold=1e-7):
    """reference :105-113 -- a near-silent first/last frame copies its neighbour."""
    x = np.asarray(autocorr_features)
    eng = _generic_engine()
    if zero_threshold != 1e-7:
        eng.set_edge_zero_threshold(zero_threshold)
    try:
        out = eng.post(np.ascontiguousarray(x.T, dtype=np.float32), nv.POST_EDGEFIX)
    finally:
        if zero_threshold != 1e-7:
            eng.set_edge_zero_threshold(1e-7)
    return out.T.astype(x.dtype if x.dtype in (np.float32, np.float64) else np.float32, copy=False)


def extract_autocorrelation_features(y, sr, frame_length, hop_length, include_deltas=False):
    """reference :116-128 -> ``[R, 187]`` (``[R, 561]`` with deltas) float64."""
    flags = nv.NO_MFCC | (nv.AC_DELTAS if include_deltas else 0)
    return _extract(y, sr, frame_length, hop_length, flags).astype(np.float64)


def compute_autocorr_with_deltas(autocorr_base):
    """reference :131-135 -> ``vstack([x, delta(x), delta(x, order=2)])``."""
    x = np.asarray(autocorr_base)
    out = _generic_engine().post(np.ascontiguousarray(x.T, dtype=np.float32), nv.POST_DELTAS)
    return out.T.astype(x.dtype if x.dtype in (np.float32, np.float64) else np.float32, copy=False)
