"""deflate.hpp_b200 -- B200-native DEFLATE / INFLATE hot path behind the API of HyperBitGore/deflate.hpp.

The product is the C-ABI shared library ``libb200deflate.so`` (hand-written sm_100a CUDA kernels,
``csrc/``) plus the header-only C++17 drop-ins in ``include/``.  This Python package is the thin
ctypes mirror of that ABI used by the tests and by ``bench.py``; it contains no compute and no CPU
fallback: importing :mod:`api` raises if the library is missing, and every call raises
:class:`B200Error` when there is no usable GPU.

The directory name contains a dot (it is the name the project was given), so it cannot be imported
with a plain ``import`` statement; ``deflate_hpp_b200.py`` at the repository root registers it under
the importable alias ``deflate_hpp_b200``.
"""
from .api import (  # noqa: F401
    B200Error,
    Context,
    CHUNK,
    LEVEL_BETTER,
    LEVEL_FAST,
    LEVEL_HUFFMAN,
    LEVEL_STORED,
    F_NOT_LAST,
    F_NO_INDEX,
    F_STRICT,
    F_ZLIB,
    F_GZIP,
    decompress_gzip,
    compress_file,
    decompress_file,
    compress,
    decompress,
    decompress_zlib,
    deflate_bound,
    launch_count,
    lib,
    lib_path,
)
