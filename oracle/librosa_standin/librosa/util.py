"""librosa.util subset: ``frame`` and ``valid_audio`` (Appendix A.3)."""
import numpy as np


class ParameterError(Exception):
    pass


def valid_audio(y):
    if not isinstance(y, np.ndarray):
        raise ParameterError("Audio data must be of type numpy.ndarray")
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.ndim == 0:
        raise ParameterError("Audio data must be at least one-dimensional")
    if not np.isfinite(y).all():
        raise ParameterError("Audio buffer is not finite everywhere")
    return True


def frame(x, *, frame_length, hop_length, axis=-1):
    """Read-only strided view of shape (frame_length, T), T = 1 + (len(x) - F) // H."""
    x = np.asarray(x)
    if x.ndim != 1 or axis not in (-1, 0):
        raise ParameterError("stand-in frames 1-D signals only")
    if x.shape[0] < frame_length:
        raise ParameterError(
            f"Input is too short (n={x.shape[0]}) for frame_length={frame_length}")
    if hop_length < 1:
        raise ParameterError(f"Invalid hop_length: {hop_length}")
    view = np.lib.stride_tricks.sliding_window_view(x, frame_length)[::hop_length]
    return view.T  # (F, T); column t == x[t*H : t*H + F]
