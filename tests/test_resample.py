"""Sample-rate conversion (reference call sites utils/audio/load_audio.py:9,19,25,36).

The reference's resampler is soxr_hq (through librosa), which cannot be reproduced here: this step is
parity-UNPINNED against the reference and pinned instead to scipy.signal.resample_poly, whose arithmetic
the product implements (include/nsf.h, nsf_resample_*)."""
import io
import wave

import numpy as np
import pytest
from scipy import signal

from oracle import resample_oracle as ro

RATES = [(44100, 88200), (48000, 88200), (88200, 16000), (16000, 88200), (22050, 88200), (88200, 44100)]


@pytest.fixture(scope="module")
def nv():
    from neurosync_trainer_lite_b200 import _native
    return _native


def _sig(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.6 * np.sin(2 * np.pi * 220 * t) + 0.2 * rng.standard_normal(n)).astype(np.float32)


@pytest.mark.parametrize("orig,target", RATES)
def test_design_matches_scipy_firwin(nv, orig, target):
    import ctypes as C
    up, down, half_len, n_pre_pad, n_pre_remove, h = ro.design(orig, target)
    want = signal.firwin(2 * half_len + 1, 1.0 / max(up, down), window=("kaiser", 5.0)) * up
    np.testing.assert_allclose(h, want, rtol=0, atol=1e-13)
    vals = [C.c_int32() for _ in range(4)]
    n = nv.lib.nsf_resample_design(orig, target, None, 0, *[C.byref(v) for v in vals])
    assert n == 2 * half_len + 1
    assert [v.value for v in vals] == [up, down, n_pre_pad, n_pre_remove]
    taps = np.empty(n, dtype=np.float64)
    nv.lib.nsf_resample_design(orig, target, taps.ctypes.data_as(C.POINTER(C.c_double)), n, None, None, None, None)
    np.testing.assert_allclose(taps, want, rtol=0, atol=1e-13)


@pytest.mark.parametrize("orig,target", RATES)
@pytest.mark.parametrize("n", [1, 7, 1000, 4411])
def test_oracle_matches_scipy_resample_poly(nv, orig, target, n):
    x = _sig(n, seed=n).astype(np.float64)
    up, down = ro.design(orig, target)[:2]
    want = signal.resample_poly(x, up, down)
    got = ro.resample(x, orig, target)
    assert got.shape == want.shape
    assert nv.lib.nsf_resample_len(n, orig, target) == len(want)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_resample_len_edge_cases(nv):
    assert nv.lib.nsf_resample_len(0, 44100, 88200) == 0
    assert nv.lib.nsf_resample_len(10, 0, 88200) == 0
    assert nv.lib.nsf_resample_len(10, 88200, 88200) == 10
    assert nv.lib.nsf_resample_len(3, 88200, 16000) == 1


# ---- GPU: the kernel against the oracle, through the C ABI ------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("orig,target", RATES)
def test_gpu_resample_matches_oracle(orig, target):
    from neurosync_trainer_lite_b200 import engine
    f, h = engine.frame_params(88200)
    eng = engine.get_engine(88200, f, h)
    for n in (1, 333, 50001):
        x = _sig(n, seed=n)
        got = eng.resample_host(x, orig, target)
        want = ro.resample(x, orig, target)
        assert got.dtype == np.float32 and got.shape == want.shape
        # float64 accumulation on the device, one rounding to float32: half an ulp of the largest value
        np.testing.assert_allclose(got, want, rtol=0, atol=1.5e-7)


@pytest.mark.gpu
def test_gpu_resample_int16_and_loader(tmp_path):
    from neurosync_trainer_lite_b200 import engine
    from neurosync_trainer_lite_b200.utils.audio import load_audio as la
    f, h = engine.frame_params(88200)
    eng = engine.get_engine(88200, f, h)
    pcm = np.clip(np.rint(_sig(44100, seed=5) * 20000), -32768, 32767).astype(np.int16)
    got = eng.resample_host(pcm, 44100, 88200)
    want = ro.resample(pcm.astype(np.float32) / np.float32(32768), 44100, 88200)
    np.testing.assert_allclose(got, want, rtol=0, atol=1.5e-7)
    # file entry point: a 44.1 kHz WAV comes back at 88.2 kHz, peak-normalised (load_audio.py:6-16)
    path = tmp_path / "a.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    y, sr = la.load_and_preprocess_audio(str(path), sr=44100)
    assert sr == 88200 and len(y) == 2 * len(pcm) and y.dtype == np.float32
    ref = want / np.abs(want.astype(np.float32)).max()
    np.testing.assert_allclose(y, ref, rtol=0, atol=3e-7)
    assert np.abs(y).max() == pytest.approx(1.0, abs=1e-6)
    # bytes entry point converts to the requested rate only (load_audio.py:23-32)
    y2, sr2 = la.load_audio_from_bytes(path.read_bytes(), sr=16000)
    assert sr2 == 16000 and len(y2) == ro.resample(pcm, 44100, 16000).shape[0]
