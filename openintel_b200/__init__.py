"""openintel_b200 — B200-native (sm_100a) hybrid retrieval engine: BM25 + cosine + RRF.

The product is libopenintel_gpu.so (CUDA kernels behind the C ABI in include/openintel_gpu.h).
This package is the thin Python binding used by the tests and bench.py; the C++ host layer in
openintel_b200/host/ mirrors the reference's port/adapter conventions.  There is no CPU
fallback: importing works anywhere, but every compute call needs the CUDA library and a GPU.
"""
from .capi import (GpuIndex, OiError, NO_DOC, DTYPE_F32, DTYPE_BF16, lib_path, load_library,  # noqa: F401
                   lexicon_analyze, version, GpuLexicon, pack_texts)
from . import fusion  # noqa: F401  (host-side crowding / alignment / confidence on the GPU analyzer's summary)

__all__ = ["GpuIndex", "OiError", "NO_DOC", "DTYPE_F32", "DTYPE_BF16", "lib_path", "load_library",
           "lexicon_analyze", "version", "GpuLexicon", "pack_texts", "fusion"]
