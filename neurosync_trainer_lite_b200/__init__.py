"""B200-native audio feature front-end of NeuroSync Trainer Lite (hot path only).

The package mirrors the reference's module paths for the path it replaces:

* ``utils.audio.extraction.extract_features``        (reference: same path)
* ``utils.audio.extraction.extract_features_utils``
* ``utils.audio.load_audio``
* ``dataset.data_processing`` / ``dataset.dataset``

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in ``include/nsf.h``
(``csrc/``, built to ``_lib/libnsf.so`` by ``__graft_entry__.build()``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

REFERENCE_MODULES = (
    "utils.audio.load_audio",
    "utils.audio.extraction.extract_features",
    "utils.audio.extraction.extract_features_utils",
    "utils.video.mov_extraction",
    "dataset.data_processing",
    "dataset.dataset",
)


def install_reference_aliases():
    """Register this package's modules under the reference's import paths (``utils.audio...``,
    ``dataset...``) so the reference's ``train.py`` / dataset loader pick them up unchanged.
    Call before the reference imports its own copies."""
    import importlib
    import sys
    for name in REFERENCE_MODULES:
        sys.modules[name] = importlib.import_module(__name__ + "." + name)
