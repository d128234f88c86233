"""Golden vectors for tests/test_audio_processing.py, produced by the REFERENCE's own
utils/audio/processing/audio_processing.py (imported from /root/reference, which only exists in the build
container).  Run from the repo root:  python oracle/make_golden_chunker.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_audio_processing import CASES, ToyModel, _features  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_audio_processing",
                                              "/root/reference/utils/audio/processing/audio_processing.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
out = {}
for n, frame, overlap in CASES:
    out[f"out_{n}_{frame}_{overlap}"] = ref.process_audio_features(
        _features(n, seed=n), ToyModel(), "cpu", {"frame_size": frame, "overlap": overlap})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "audio_processing.npz"), **out)
print({k: v.shape for k, v in out.items()})
