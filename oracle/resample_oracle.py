"""CPU oracle of the sample-rate conversion step (TEST INFRASTRUCTURE ONLY - never imported by the
product package).

The reference resamples through librosa (``librosa.resample`` at ``utils/audio/load_audio.py:9``,
``librosa.load(..., sr=sr)`` at :19, :25, :36), i.e. soxr_hq - a third-party filter that is neither
vendored in /root/reference nor installable here, so THIS STEP IS PARITY-UNPINNED against the
reference.  What the product implements, and what this file restates in
plain float64 NumPy, is the rational polyphase resampler of ``scipy.signal.resample_poly``:

    up/down = target_sr/orig_sr (reduced),  half_len = 10 max(up, down)
    h = up * firwin(2 half_len + 1, 1 / max(up, down), window=('kaiser', 5.0))
    out[j] = sum_n hpad[(j + n_pre_remove) down - n up] x[n],   j < ceil(len(x) up / down)
    hpad = n_pre_pad zeros ++ h,  n_pre_pad = down - half_len % down,
    n_pre_remove = (half_len + n_pre_pad) // down

``tests/test_resample.py`` pins it to scipy itself (taps against ``scipy.signal.firwin``, output
against ``scipy.signal.resample_poly`` on float64 input).
"""
from math import gcd

import numpy as np


def design(orig_sr, target_sr):
    """-> (up, down, half_len, n_pre_pad, n_pre_remove, h float64)."""
    g = gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    max_rate = max(up, down)
    half_len = 10 * max_rate
    n = 2 * half_len + 1
    fc = 1.0 / max_rate
    m = np.arange(n, dtype=np.float64) - 0.5 * (n - 1)
    h = fc * np.sinc(fc * m) * np.kaiser(n, 5.0)
    h = h / h.sum() * up
    n_pre_pad = down - half_len % down
    return up, down, half_len, n_pre_pad, (half_len + n_pre_pad) // down, h


HQ_WIDTH, HQ_ROLLOFF, HQ_BETA = 64.0, 0.9475937167399596, 14.769656459379492     # the "kaiser_best" parameters


def design_hq(orig_sr, target_sr):
    """The high-quality design (what the loaders use): band-limited windowed sinc

        out[m] = sum_n x[n] g((n / down - m / up) f),  f = min(up, down) * rolloff,
        g(t) = (f / down) sinc(t) I0(beta sqrt(1 - (t / 64)^2)) / I0(beta)  for |t| <= 64, else 0

    i.e. torchaudio.functional.resample(..., lowpass_filter_width=64, rolloff=HQ_ROLLOFF,
    resampling_method="sinc_interp_kaiser", beta=HQ_BETA), expressed as a symmetric FIR on the up*down grid
    so that the polyphase indexing of ``resample`` applies unchanged.  ``tests/test_resample.py`` pins it to
    torchaudio itself."""
    g = gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    f = min(up, down) * HQ_ROLLOFF
    per_tap = f / (float(up) * down)
    half_len = int(np.floor(HQ_WIDTH / per_tap))
    t = (np.arange(2 * half_len + 1, dtype=np.float64) - half_len) * per_tap
    win = np.i0(HQ_BETA * np.sqrt(np.maximum(0.0, 1.0 - (t / HQ_WIDTH) ** 2))) / np.i0(HQ_BETA)
    h = (f / down) * np.sinc(t) * win
    n_pre_pad = down - half_len % down
    return up, down, half_len, n_pre_pad, (half_len + n_pre_pad) // down, h


def resample(x, orig_sr, target_sr, quality="poly"):
    """float64 polyphase resampling of a 1-D signal (direct evaluation of the sum above)."""
    x = np.asarray(x, dtype=np.float64)
    up, down, half_len, n_pre_pad, n_pre_remove, h = (design_hq if quality == "hq" else design)(orig_sr, target_sr)
    n_in = len(x)
    n_out = (n_in * up) // down + ((n_in * up) % down != 0)
    out = np.zeros(n_out, dtype=np.float64)
    j = np.arange(n_out, dtype=np.int64)
    t = (j + n_pre_remove) * down - n_pre_pad
    n_hi = t // up
    phase = t - n_hi * up
    kmax = (len(h) + up - 1) // up
    for k in range(kmax):
        q = phase + k * up              # tap index
        n = n_hi - k                    # input index
        ok = (q < len(h)) & (n >= 0) & (n < n_in)
        out[ok] += h[q[ok]] * x[n[ok]]
    return out
