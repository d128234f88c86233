"""librosa.feature subset: ``melspectrogram``, ``mfcc``, ``delta`` (Appendix A.1, A.2)."""
import numpy as np
import scipy.fftpack
import scipy.signal

from . import core, filters


def melspectrogram(*, y, sr, n_fft=2048, hop_length=512, win_length=None, window="hann",
                   center=True, pad_mode="constant", power=2.0, **mel_kwargs):
    S = np.abs(core.stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                         window=window, center=center, pad_mode=pad_mode)) ** power
    mel_basis = filters.mel(sr=sr, n_fft=n_fft, **mel_kwargs)
    return np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)


def mfcc(*, y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0, **kwargs):
    if S is None:
        S = core.power_to_db(melspectrogram(y=y, sr=sr, **kwargs))
    M = scipy.fftpack.dct(S, axis=-2, type=dct_type, norm=norm)[..., :n_mfcc, :]
    if lifter != 0:
        raise ValueError("stand-in: lifter unsupported (the reference leaves it at 0)")
    return M


def delta(data, *, width=9, order=1, axis=-1, mode="interp", **kwargs):
    data = np.atleast_1d(data)
    if mode == "interp" and width > data.shape[axis]:
        raise ValueError(
            f"when mode='interp', width={width} cannot exceed data.shape[axis]={data.shape[axis]}")
    if width < 3 or width % 2 != 1:
        raise ValueError("width must be an odd integer >= 3")
    kwargs.setdefault("polyorder", order)
    return scipy.signal.savgol_filter(data, width, deriv=order, axis=axis, mode=mode, **kwargs)
