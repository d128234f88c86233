"""CPU-only checks of the C ABI: the library loads, exports every symbol ``include/nsf.h`` declares,
its integer frame arithmetic is bit-exact against the oracle, its constant tables (windows, mel,
DCT, folded-DFT index tables) match the oracle's, and compute calls fail loudly without a GPU.
No kernel is launched here.
"""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle", "librosa_standin"))


@pytest.fixture(scope="module")
def nv():
    import __graft_entry__ as g
    g.build_library()
    from neurosync_trainer_lite_b200 import _native
    return _native


@pytest.fixture(scope="module")
def engine(nv):
    from neurosync_trainer_lite_b200 import engine
    return engine


def test_every_declared_symbol_is_exported(nv):
    header = open(os.path.join(ROOT, "include", "nsf.h")).read()
    declared = set(re.findall(r"NSF_API[^;(]*?\b(nsf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(nv.lib, name), f"{name} declared in nsf.h but not exported by libnsf.so"
    assert declared == set(nv.EXPORTED), declared ^ set(nv.EXPORTED)
    assert nv.lib.nsf_abi_version() == 2


@pytest.mark.parametrize("sr", [88200, 16000, 44100, 22050, 48000, 8000, 96000, 11025, 32000])
def test_frame_params_bit_exact(nv, oracle, sr):
    f = nv.lib.nsf_frame_length(sr)
    assert (f, nv.lib.nsf_hop_length(f)) == oracle.frame_params(sr)


def test_frame_counts_against_reference_kats(nv, oracle, golden):
    counts = golden("kat")["row_counts"]
    for n, t, r in counts:
        if t < 0:
            assert nv.lib.nsf_guard_frames(int(n), 1470, 735) < 9
        else:
            assert nv.lib.nsf_hop_frames(int(n), 1470, 735) == t
            assert nv.lib.nsf_feature_rows(int(n), 1470, 735) == r
    rng = np.random.default_rng(0)
    for F, H in [(1470, 735), (266, 133), (735, 367), (367, 183)]:
        for n in rng.integers(0, 400000, size=200):
            n = int(n)
            assert nv.lib.nsf_guard_frames(n, F, H) == oracle.guard_frames(n, F, H)
            if n + 2 * (F // 2) >= F:
                assert nv.lib.nsf_hop_frames(n, F, H) == oracle.hop_frames(n, F, H)
                assert nv.lib.nsf_feature_rows(n, F, H) == oracle.feature_rows(n, F, H)


def test_row_offsets_is_the_prefix_sum_of_the_per_clip_counts(nv, oracle):
    """nsf_row_offsets (one call per batch) against the per-clip functions and the oracle, for every flag
    that changes the row count; 10 000 clips must not cost a per-clip binding call each."""
    import time
    rng = np.random.default_rng(5)
    for F, H in [(1470, 735), (266, 133), (735, 367)]:
        lens = rng.integers(0, 200000, size=500).astype(np.int64)
        lens[:4] = [0, F - 1, F, F + H]
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        off += 12345                                       # offsets need not start at 0
        for flags in (0, nv.NO_REDUCE, nv.AC_NO_PAD, nv.AC_NO_PAD | nv.NO_REDUCE):
            got = np.full(len(off), -1, dtype=np.int64)
            assert nv.lib.nsf_row_offsets(F, H, off.ctypes.data_as(nv._i64p), len(lens), flags,
                                          got.ctypes.data_as(nv._i64p)) == nv.OK
            want = [0]
            for n in lens:
                n = int(n)
                if flags & nv.AC_NO_PAD:
                    t = max(0, oracle.guard_frames(n, F, H)) if n >= F else 0
                else:
                    t = nv.lib.nsf_hop_frames(n, F, H)
                    assert t == oracle.hop_frames(n, F, H)
                want.append(want[-1] + (t if flags & nv.NO_REDUCE else (t + 1) // 2))
            assert got.tolist() == want
    bad = np.array([0, 10, 5], dtype=np.int64)
    out = np.zeros(3, dtype=np.int64)
    assert nv.lib.nsf_row_offsets(266, 133, bad.ctypes.data_as(nv._i64p), 2, 0, out.ctypes.data_as(nv._i64p)) == nv.ERR_BAD_ARG
    off = np.arange(10001, dtype=np.int64) * 32000
    out = np.zeros(10001, dtype=np.int64)
    t0 = time.perf_counter()
    assert nv.lib.nsf_row_offsets(266, 133, off.ctypes.data_as(nv._i64p), 10000, 0, out.ctypes.data_as(nv._i64p)) == nv.OK
    assert time.perf_counter() - t0 < 2e-3 and out[-1] == 10000 * nv.lib.nsf_feature_rows(32000, 266, 133)


def test_collect_rows_match_oracle(nv, oracle):
    for n in [0, 1, 2, 5, 29, 30, 31, 59, 60, 61, 1800, 1801]:
        for fast in (False, True):
            for slow in (False, True):
                for blend in (False, True):
                    for bf in (0, 3, 30, 5000):
                        flags = (1 if fast else 0) | (2 if slow else 0) | (4 if blend else 0)
                        want = oracle.collected_rows(n, fast, slow, blend, bf) if n > 0 else 0
                        if n == 0:
                            continue
                        assert nv.lib.nsf_collect_rows(n, n + 7, flags, bf) == want
                        assert nv.lib.nsf_collect_rows(n + 3, n, flags, bf) == want


@pytest.mark.parametrize("sr,F,H", [(88200, 1470, 735), (16000, 266, 133), (44100, 735, 367),
                                    (22050, 367, 183)])
def test_tables_match_oracle(engine, nv, sr, F, H):
    import librosa
    import scipy.fftpack
    import scipy.signal
    plan = engine.Plan(sr, F, H)
    assert plan.bins == F // 2 + 1
    mel = librosa.filters.mel(sr=sr, n_fft=F, n_mels=128)
    np.testing.assert_array_equal(plan.mel_basis(), mel)             # float32, bit for bit
    dct = scipy.fftpack.dct(np.eye(128), axis=0, type=2, norm="ortho")[:23]
    np.testing.assert_allclose(plan.dct_matrix(), dct, rtol=0, atol=1e-7)
    np.testing.assert_allclose(plan.table(nv.TABLE_HANN_SYM), np.hanning(F), rtol=0, atol=6e-8)
    np.testing.assert_allclose(plan.table(nv.TABLE_HANN_PER),
                               scipy.signal.get_window("hann", F, fftbins=True), rtol=0, atol=6e-8)
    assert plan.feature_cols(0) == 256
    assert plan.feature_cols(nv.NO_AUTOCORR) == 69
    assert plan.feature_cols(nv.NO_MFCC | nv.AC_DELTAS) == 561


@pytest.mark.parametrize("F", [1470, 266, 735, 367, 64, 98])
def test_fold_tables_reproduce_the_rfft(engine, F):
    """The 4x-folded DFT is an exact re-indexing: checked in float64 on the host."""
    import scipy.signal
    plan = engine.Plan(88200, F, F // 2, n_lags=min(187, F - 1))
    rng = np.random.default_rng(F)
    for _ in range(3):
        frame = rng.standard_normal(F).astype(np.float32)
        want = np.fft.rfft(scipy.signal.get_window("hann", F, fftbins=True) * frame.astype(np.float64))
        got = plan.fold_check(frame)
        assert got.shape == want.shape
        # window taps are rounded to float32 inside the tables: ~1e-7 relative to the frame norm
        np.testing.assert_allclose(got, want, rtol=0, atol=3e-6 * np.sqrt(F))


def test_bad_geometry_is_rejected(nv):
    h = C.c_void_p()
    assert nv.lib.nsf_plan_create(88200, 2, 1, 23, 128, 187, C.byref(h)) == nv.ERR_BAD_ARG
    assert nv.lib.nsf_plan_create(88200, 8192, 4096, 23, 128, 187, C.byref(h)) == nv.ERR_UNSUPPORTED
    assert b"kernel limits" in nv.lib.nsf_last_error()


def test_no_gpu_means_loud_failure_not_fallback(nv, engine):
    """On a box without an sm_100 device every compute entry point must refuse to run."""
    if nv.lib.nsf_device_count() > 0:
        pytest.skip("a B200 is present; this test is for CPU-only boxes")
    plan = engine.Plan(88200, 1470, 735)
    with pytest.raises(nv.NsfError) as e:
        engine.Engine(plan, 0)
    assert e.value.status == nv.ERR_NO_DEVICE
    from neurosync_trainer_lite_b200.utils.audio.extraction.extract_features import (
        extract_and_combine_features)
    with pytest.raises(nv.NsfError):
        extract_and_combine_features(np.zeros(88200, np.float32), 88200, 1470, 735)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neurosync_trainer_lite_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "feature_oracle" not in text and "librosa_standin" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_build_stamp_is_path_independent_and_build_is_locked(tmp_path, monkeypatch):
    """The library built in this checkout must be recognised as current in a COPY of the checkout (the GPU
    box runs from a different path; a rebuild there raced between torchrun ranks), and concurrent callers
    must serialise on the build lock instead of loading a half-written file."""
    import shutil

    from neurosync_trainer_lite_b200 import build as b
    here = b._digest()
    root2 = tmp_path / "elsewhere"
    shutil.copytree(b.CSRC, root2 / "neurosync_trainer_lite_b200" / "csrc")
    (root2 / "include").mkdir()
    shutil.copy(os.path.join(b.ROOT, "include", "nsf.h"), root2 / "include" / "nsf.h")
    monkeypatch.setattr(b, "ROOT", str(root2))
    monkeypatch.setattr(b, "CSRC", str(root2 / "neurosync_trainer_lite_b200" / "csrc"))
    assert b._digest() == here
    src = open(b.__file__).read()
    assert "fcntl.flock" in src and "os.replace(lib_tmp, LIB_PATH)" in src


def test_chunk_count_matches_the_reference_loop(nv):
    """nsf_chunk_count == the number of iterations of process_audio_features' while loop
    (utils/audio/processing/audio_processing.py:62-84); invalid geometries give 0."""
    for n in (1, 7, 50, 112, 113, 128, 129, 230, 240, 300, 1801):
        for frame, overlap in ((128, 16), (64, 8), (128, 32), (64, 0)):
            count, start = 0, 0
            while start < n:
                count += 1
                start += frame - overlap
            assert nv.lib.nsf_chunk_count(n, frame, overlap) == count
    assert nv.lib.nsf_chunk_count(0, 128, 16) == 0 and nv.lib.nsf_chunk_count(10, 16, 16) == 0
    assert nv.lib.nsf_chunk_count(10, 0, 0) == 0 and nv.lib.nsf_chunk_count(10, 16, -1) == 0
