"""Pin the librosa stand-in against independent implementations of the same published
definitions (SURVEY.md section 8(c) cross-checks 1-4).  librosa itself cannot be installed."""
import os
import sys

import numpy as np
import pytest
import scipy.fft
import scipy.signal

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "oracle", "librosa_standin"))
import librosa  # noqa: E402  (the stand-in)

from neurosync_trainer_lite_b200 import synth  # noqa: E402

# even n_fft only: for odd n_fft torchaudio/transformers put the last bin at sr/2 (linspace) while
# librosa uses rfftfreq; they are not comparable there
CASES = [(88200, 1470), (16000, 266)]


@pytest.mark.parametrize("sr,F", CASES)
def test_mel_basis_vs_torchaudio_and_transformers(sr, F):
    import types
    import torchaudio
    # transformers sees the stand-in as "librosa installed" and then wants soxr, which is absent
    sys.modules.setdefault("soxr", types.ModuleType("soxr"))
    from transformers.audio_utils import mel_filter_bank
    ours = librosa.filters.mel(sr=sr, n_fft=F, n_mels=128)
    assert ours.dtype == np.float32 and ours.shape == (128, F // 2 + 1)
    ta = torchaudio.functional.melscale_fbanks(F // 2 + 1, 0.0, sr / 2, 128, sr, norm="slaney",
                                               mel_scale="slaney").T.numpy()
    assert np.abs(ours - ta).max() < 5e-7
    hf = mel_filter_bank(F // 2 + 1, 128, 0.0, sr / 2, sr, norm="slaney", mel_scale="slaney").T
    assert np.abs(ours - hf).max() < 1e-8


def test_mel_sparsity_facts():
    m = librosa.filters.mel(sr=88200, n_fft=1470, n_mels=128)
    assert np.count_nonzero(m) == 1442 and np.all(m.sum(axis=1) > 0)
    m16 = librosa.filters.mel(sr=16000, n_fft=266, n_mels=128)
    assert int(np.sum(m16.sum(axis=1) == 0)) == 11


def test_delta_closed_form_and_edges():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 40)).astype(np.float32)
    d1 = librosa.feature.delta(x)
    d2 = librosa.feature.delta(x, order=2)
    k = np.arange(-4, 5)
    c1 = k / 60.0
    c2 = np.array([28, 7, -8, -17, -20, -17, -8, 7, 28]) / 462.0
    for t in range(4, 36):
        np.testing.assert_allclose(d1[:, t], x[:, t - 4:t + 5].astype(np.float64) @ c1, atol=2e-6)
        np.testing.assert_allclose(d2[:, t], x[:, t - 4:t + 5].astype(np.float64) @ c2, atol=2e-6)
    for t in range(4):
        np.testing.assert_allclose(d1[:, t], d1[:, 4], atol=2e-6)
        np.testing.assert_allclose(d2[:, t], d2[:, 4], atol=2e-6)
        np.testing.assert_allclose(d1[:, -1 - t], d1[:, -5], atol=2e-6)
        np.testing.assert_allclose(d2[:, -1 - t], d2[:, -5], atol=2e-6)
    with pytest.raises(ValueError):
        librosa.feature.delta(x[:, :8])


def test_dct_matrix_form():
    rng = np.random.default_rng(1)
    s = rng.standard_normal((128, 7))
    m = np.arange(128)
    D = np.sqrt(2.0 / 128) * np.cos(np.pi * np.outer(np.arange(23), 2 * m + 1) / 256.0)
    D[0] *= 1 / np.sqrt(2.0)
    ref = scipy.fft.dct(s, type=2, norm="ortho", axis=0)[:23]
    assert np.abs(D @ s - ref).max() < 1e-12


def test_frame_and_stft_alignment():
    x = np.arange(1000, dtype=np.float32)
    fr = librosa.util.frame(x, frame_length=100, hop_length=30)
    assert fr.shape == (100, 1 + (1000 - 100) // 30)
    np.testing.assert_array_equal(fr[:, 7], x[210:310])
    y = synth.synth_clip(0.3, 16000, seed=2)
    D = librosa.stft(y, n_fft=266, hop_length=133)
    assert D.dtype == np.complex64 and D.shape == (134, 1 + len(y) // 133)
    w = scipy.signal.get_window("hann", 266, fftbins=True)
    np.testing.assert_allclose(w, 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(266) / 266), atol=1e-15)
    yp = np.pad(y, 133)
    np.testing.assert_allclose(D[:, 5], np.fft.rfft(w * yp[5 * 133:5 * 133 + 266]), atol=1e-5)


@pytest.mark.parametrize("sr,F", [(88200, 1470), (16000, 266)])
def test_mfcc_vs_torchaudio_end_to_end(sr, F):
    import torch
    import torchaudio
    y = synth.synth_clip(2.0, sr, seed=4)
    ours = librosa.feature.mfcc(y=y, sr=sr, n_mfcc=23, n_fft=F, hop_length=F // 2)
    tm = torchaudio.transforms.MFCC(
        sample_rate=sr, n_mfcc=23, dct_type=2, norm="ortho", log_mels=False,
        melkwargs=dict(n_fft=F, hop_length=F // 2, n_mels=128, f_min=0.0, f_max=sr / 2,
                       center=True, pad_mode="constant", power=2.0, norm="slaney",
                       mel_scale="slaney"))
    theirs = tm(torch.from_numpy(y)).numpy()
    assert ours.shape == theirs.shape == (23, 1 + len(y) // (F // 2))
    scale = np.abs(ours).max()
    assert np.abs(ours - theirs).max() < 2e-5 * scale


def test_wav_decode_int16():
    pcm = np.array([0, 1, -1, 32767, -32768, 1000], dtype=np.int16)
    y, sr = librosa.load(synth.wav_bytes(pcm, 16000), sr=16000)
    assert sr == 16000 and y.dtype == np.float32
    np.testing.assert_array_equal(y, pcm.astype(np.float32) / np.float32(32768))
