"""Minimal stand-in for the five ``librosa`` entry points the reference hot path calls.

TEST INFRASTRUCTURE ONLY (part of ``oracle/``): never imported by the product package.

``librosa`` is an un-vendored, unpinned third-party dependency of the reference
(``/root/reference/requirements.txt:12``) and is not installable in this image.  This package
restates the *published* behaviour of librosa 0.10.2 / 0.11.0 for exactly the calls made on the
hot path:

* ``librosa.feature.mfcc``   - utils/audio/extraction/extract_features_utils.py:19
* ``librosa.feature.delta``  - extract_features_utils.py:25-26,132-133
* ``librosa.util.frame``     - extract_features_utils.py:64
* ``librosa.load``           - utils/audio/load_audio.py:19,25,36
* ``librosa.resample``       - utils/audio/load_audio.py:9

PARITY UNPINNED against real librosa (it cannot be run here).  The stand-in is cross-checked in
``tests/test_oracle_standin.py`` against torchaudio, transformers.audio_utils and scipy, which
implement the same published definitions independently.
"""
from . import feature, filters, util  # noqa: F401
from .core import load, resample, stft, power_to_db  # noqa: F401

__version__ = "0.10.2.standin"
