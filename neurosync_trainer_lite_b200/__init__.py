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
