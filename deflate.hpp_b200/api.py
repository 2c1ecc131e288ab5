"""ctypes mirror of include/b200_deflate.h.

Function names and argument meaning follow the reference's public API
(HyperBitGore/deflate.hpp, include/deflate.hpp:753-815 and include/inflate.hpp:324-408):

    compress(data, level)            <-> deflate::compress(char*, size_t, int) / (vector&, int)
    decompress(data[, out_size])     <-> inflate::decompress(void*, size_t[, void*, size_t])
    decompress_zlib(data[, out_size])<-> inflate::decompressZlib(void*, size_t[, void*, size_t])

Errors: the reference inflater throws std::runtime_error("Reading bits beyond the alloted buffer
size!") on truncated input; here every non-zero return code of the C ABI raises B200Error (code and
the same message).  There is no CPU path: without libb200deflate.so the import fails, without a GPU
every call raises B200Error(B200_E_CUDA).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libb200deflate.so"

CHUNK = 65536
LEVEL_STORED, LEVEL_HUFFMAN, LEVEL_FAST, LEVEL_BETTER = 0, 1, 2, 3
F_NOT_LAST = 1
F_NO_INDEX = 2
F_ZLIB = 4
F_GZIP = 8
F_STRICT = 1
E_OVERRUN, E_DATA, E_OUTPUT, E_CUDA, E_ARG, E_NOMEM, E_IO = 1, 2, 3, 4, 5, 6, 7


class B200Error(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        msg = lib().b200_strerror(code).decode() if _lib is not None else str(code)
        super().__init__(f"{what + ': ' if what else ''}{msg} (code {code})")


_lib = None


def lib_path():
    return os.path.join(HERE, _LIB_NAME)


def lib():
    """Load libb200deflate.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python __graft_entry__.py` (or "
            f"`python deflate.hpp_b200/build.py`); there is no CPU fallback.")
    L = ctypes.CDLL(path)
    c_void_p, c_size_t, c_int, c_uint, c_u64 = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                                 ctypes.c_uint, ctypes.c_uint64)
    P = ctypes.POINTER
    L.b200_abi_version.restype = c_int
    L.b200_strerror.restype = ctypes.c_char_p
    L.b200_strerror.argtypes = [c_int]
    L.b200_launch_count.restype = c_u64
    L.b200_ctx_create.argtypes = [c_int, P(c_void_p)]
    L.b200_ctx_destroy.argtypes = [c_void_p]
    L.b200_ctx_profile.argtypes = [c_void_p, c_int]
    L.b200_ctx_profile.restype = c_int
    L.b200_ctx_profile_read.argtypes = [c_void_p, c_int, P(ctypes.c_double), P(c_u64)]
    L.b200_ctx_profile_read.restype = c_int
    L.b200_kernel_name.argtypes = [c_int]
    L.b200_kernel_name.restype = ctypes.c_char_p
    L.b200_deflate_bound.restype = c_size_t
    L.b200_deflate_bound.argtypes = [c_size_t]
    L.b200_deflate_compress.argtypes = [c_void_p, c_size_t, c_int, P(c_void_p), P(c_size_t)]
    L.b200_deflate_compress_into.argtypes = [c_void_p, c_size_t, c_int, c_void_p, c_size_t, P(c_size_t)]
    L.b200_inflate.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, P(c_size_t), P(c_size_t), c_uint]
    L.b200_inflate_alloc.argtypes = [c_void_p, c_size_t, P(c_void_p), P(c_size_t), c_uint]
    L.b200_inflate_zlib.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, P(c_size_t), P(c_size_t), c_uint]
    L.b200_inflate_zlib_alloc.argtypes = [c_void_p, c_size_t, P(c_void_p), P(c_size_t), c_uint]
    L.b200_deflate_compress_ex.argtypes = [c_void_p, c_size_t, c_int, c_uint, P(c_void_p), P(c_size_t)]
    L.b200_deflate_compress_into_ex.argtypes = [c_void_p, c_size_t, c_int, c_uint, c_void_p, c_size_t, P(c_size_t)]
    L.b200_inflate_gzip.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, P(c_size_t), P(c_size_t), c_uint]
    L.b200_inflate_gzip_alloc.argtypes = [c_void_p, c_size_t, P(c_void_p), P(c_size_t), c_uint]
    L.b200_crc32_dev.argtypes = [c_void_p, c_void_p, c_size_t, P(ctypes.c_uint32), c_void_p, c_void_p]
    for f in ("b200_deflate_compress_ex", "b200_deflate_compress_into_ex", "b200_inflate_gzip", "b200_inflate_gzip_alloc", "b200_crc32_dev"):
        getattr(L, f).restype = c_int
    L.b200_deflate_compress_file.argtypes = [ctypes.c_char_p, ctypes.c_char_p, c_int, c_uint, P(c_size_t), P(c_size_t)]
    L.b200_deflate_compress_file.restype = c_int
    L.b200_inflate_file.argtypes = [ctypes.c_char_p, ctypes.c_char_p, c_uint, P(c_size_t), P(c_size_t)]
    L.b200_inflate_file.restype = c_int
    L.b200_free.argtypes = [c_void_p]
    L.b200_deflate_compress_dev.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_uint, c_void_p, c_size_t,
                                            c_void_p, P(c_size_t), c_void_p, c_void_p]
    L.b200_deflate_compress_stage1_dev.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_uint, c_void_p, c_void_p]
    L.b200_deflate_compress_stage1_dev.restype = c_int
    L.b200_deflate_compress_stage2_dev.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]
    L.b200_deflate_compress_stage2_dev.restype = c_int
    L.b200_inflate_dev.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, P(c_size_t),
                                   P(c_size_t), c_void_p, c_uint, c_void_p]
    L.b200_inflate_shard_dev.argtypes = [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_int, c_void_p, c_size_t,
                                         P(c_size_t), P(c_size_t), P(c_size_t), c_uint, c_void_p]
    L.b200_inflate_shard_dev.restype = c_int
    L.b200_inflate_batch_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_size_t, c_uint, c_void_p]
    L.b200_corpus_generate_dev.argtypes = [c_void_p, c_u64, c_u64, c_u64, c_void_p]
    L.b200_deflate_compress_batch_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_uint, c_void_p,
                                                  c_size_t, c_void_p, P(c_size_t), c_void_p]
    L.b200_deflate_compress_batch_dev.restype = c_int
    L.b200_adler32_dev.argtypes = [c_void_p, c_void_p, c_size_t, P(ctypes.c_uint32), c_void_p, c_void_p]
    L.b200_adler32_dev.restype = c_int
    L.b200_publish_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.b200_publish_dev.restype = c_int
    for f in ("b200_ctx_create", "b200_deflate_compress", "b200_deflate_compress_into", "b200_inflate",
              "b200_inflate_alloc", "b200_inflate_zlib", "b200_inflate_zlib_alloc", "b200_deflate_compress_dev",
              "b200_inflate_dev", "b200_inflate_batch_dev", "b200_corpus_generate_dev"):
        getattr(L, f).restype = c_int
    _lib = L
    return L


def _as_buffer(data):
    """bytes-like -> (address, length, keepalive)."""
    if isinstance(data, bytes):
        return ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p).value or 0, len(data), data
    mv = memoryview(data).cast("B")
    if mv.readonly:
        b = mv.tobytes()
        return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value or 0, len(b), b
    if len(mv) == 0:
        return 0, 0, mv
    arr = (ctypes.c_char * len(mv)).from_buffer(mv)
    return ctypes.addressof(arr), len(mv), (arr, mv)


def _level(level):
    # README form: a bool picks fast (False) / better (True); the int form is 0..3 (deflate.hpp:675-680)
    if isinstance(level, bool):
        return LEVEL_BETTER if level else LEVEL_FAST
    return int(level)


def deflate_bound(n):
    return lib().b200_deflate_bound(n)


def launch_count():
    return lib().b200_launch_count()


def compress(data, level=LEVEL_FAST, flags=0):
    """deflate::compress(char* data, size_t data_size, int compression_level) -> vector<uint8_t>.
    flags: F_NO_INDEX, F_ZLIB (zlib framing, what zlib.decompress reads), F_GZIP (one gzip member)."""
    L = lib()
    addr, n, keep = _as_buffer(data)
    out = ctypes.c_void_p()
    out_n = ctypes.c_size_t()
    rc = L.b200_deflate_compress_ex(addr, n, _level(level), flags, ctypes.byref(out), ctypes.byref(out_n))
    del keep
    if rc:
        raise B200Error(rc, "deflate::compress")
    try:
        return ctypes.string_at(out.value, out_n.value)
    finally:
        L.b200_free(out)


def _inflate(fn_alloc, fn_into, data, out_size, flags, what):
    L = lib()
    addr, n, keep = _as_buffer(data)
    if out_size is None:
        out = ctypes.c_void_p()
        out_n = ctypes.c_size_t()
        rc = fn_alloc(addr, n, ctypes.byref(out), ctypes.byref(out_n), flags)
        del keep
        try:
            if rc:
                raise B200Error(rc, what)
            return ctypes.string_at(out.value, out_n.value)
        finally:
            if out.value:
                L.b200_free(out)
    buf = ctypes.create_string_buffer(max(out_size, 1))
    out_n = ctypes.c_size_t()
    full = ctypes.c_size_t()
    rc = fn_into(addr, n, buf, out_size, ctypes.byref(out_n), ctypes.byref(full), flags)
    del keep
    if rc:
        raise B200Error(rc, what)
    return buf.raw[:out_n.value]


def decompress(data, out_size=None, flags=0):
    """inflate::decompress(void* in, size_t in_size) -> vector, or with out_size the caller-buffer
    overload (inflate.hpp:338) which silently truncates at out_size like the reference."""
    L = lib()
    return _inflate(L.b200_inflate_alloc, L.b200_inflate, data, out_size, flags, "inflate::decompress")


def decompress_zlib(data, out_size=None, flags=0):
    """inflate::decompressZlib (inflate.hpp:326,352)."""
    L = lib()
    return _inflate(L.b200_inflate_zlib_alloc, L.b200_inflate_zlib, data, out_size, flags, "inflate::decompressZlib")


def compress_file(src, dst, level=LEVEL_FAST, flags=0):
    """deflate::compress(std::string file_path, std::string new_file, int level): streamed in slices.  -> (bytes in, bytes out)"""
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib().b200_deflate_compress_file(os.fsencode(src), os.fsencode(dst), _level(level), flags, ctypes.byref(a), ctypes.byref(b))
    if rc:
        raise B200Error(rc, "deflate::compress(file)")
    return a.value, b.value


def decompress_file(src, dst, flags=0):
    """inflate::decompress(std::string file_path, std::string new_file).  -> (bytes in, bytes out)"""
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib().b200_inflate_file(os.fsencode(src), os.fsencode(dst), flags, ctypes.byref(a), ctypes.byref(b))
    if rc:
        raise B200Error(rc, "inflate::decompress(file)")
    return a.value, b.value


def decompress_gzip(data, out_size=None, flags=0):
    """first member of a gzip buffer (RFC 1952); F_STRICT verifies CRC-32 and ISIZE on the GPU."""
    L = lib()
    return _inflate(L.b200_inflate_gzip_alloc, L.b200_inflate_gzip, data, out_size, flags, "inflate (gzip)")


class Context:
    """Device-resident API: raw device pointers (ints) and a cudaStream_t handle (int, 0 = default)."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        rc = lib().b200_ctx_create(device, ctypes.byref(self._h))
        if rc:
            raise B200Error(rc, "b200_ctx_create")
        self.device = device

    def close(self):
        if self._h:
            lib().b200_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def profile(self, enable):
        """Bracket every kernel launch of this context with CUDA events (bench.py roofline)."""
        rc = lib().b200_ctx_profile(self._h, 1 if enable else 0)
        if rc:
            raise B200Error(rc, "b200_ctx_profile")

    def profile_read(self):
        """-> {kernel name: (total device ms, launches)} for the records since profile(True)."""
        out = {}
        k = 0
        while True:
            name = lib().b200_kernel_name(k)
            if not name:
                return out
            ms, n = ctypes.c_double(), ctypes.c_uint64()
            rc = lib().b200_ctx_profile_read(self._h, k, ctypes.byref(ms), ctypes.byref(n))
            if rc:
                raise B200Error(rc, "b200_ctx_profile_read")
            if n.value:
                out[name.decode()] = (ms.value, n.value)
            k += 1

    def compress_dev(self, d_in, n, level, d_out, cap, flags=0, stream=0, d_out_n=0, d_chunk_off=0, sync=True):
        out_n = ctypes.c_size_t()
        rc = lib().b200_deflate_compress_dev(self._h, d_in, n, _level(level), flags, d_out, cap, d_out_n or None,
                                             ctypes.byref(out_n) if sync else None, d_chunk_off or None,
                                             stream or None)
        if rc:
            raise B200Error(rc, "b200_deflate_compress_dev")
        return out_n.value if sync else None

    def compress_stage1_dev(self, d_in, n, level, d_local_n, flags=0, stream=0):
        rc = lib().b200_deflate_compress_stage1_dev(self._h, d_in, n, _level(level), flags, d_local_n, stream or None)
        if rc:
            raise B200Error(rc, "b200_deflate_compress_stage1_dev")

    def compress_stage2_dev(self, d_in, n, d_out, d_base=0, stream=0):
        rc = lib().b200_deflate_compress_stage2_dev(self._h, d_in, n, d_out, d_base or None, stream or None)
        if rc:
            raise B200Error(rc, "b200_deflate_compress_stage2_dev")

    def inflate_dev(self, d_in, n, d_out, cap, flags=0, stream=0):
        """Returns (written, full_size)."""
        out_n = ctypes.c_size_t()
        full = ctypes.c_size_t()
        rc = lib().b200_inflate_dev(self._h, d_in, n, d_out, cap, None, ctypes.byref(out_n), ctypes.byref(full),
                                    None, flags, stream or None)
        if rc:
            raise B200Error(rc, "b200_inflate_dev")
        return out_n.value, full.value

    def inflate_shard_dev(self, d_in, n, lo, hi, first_is_start, ends_stream, d_out, cap, flags=0, stream=0):
        """A window of a longer stream (multi-GPU inflate).  Returns (decoded bytes, chunks, next_start)."""
        out_n, nch, nxt = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        rc = lib().b200_inflate_shard_dev(self._h, d_in, n, lo, hi, 1 if first_is_start else 0, 1 if ends_stream else 0,
                                          d_out, cap, ctypes.byref(out_n), ctypes.byref(nch), ctypes.byref(nxt),
                                          flags, stream or None)
        if rc:
            raise B200Error(rc, "b200_inflate_shard_dev")
        return out_n.value, nch.value, nxt.value

    def inflate_batch_dev(self, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap, d_out_len, d_status,
                          n_streams, flags=0, stream=0):
        rc = lib().b200_inflate_batch_dev(self._h, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap,
                                          d_out_len, d_status, n_streams, flags, stream or None)
        if rc:
            raise B200Error(rc, "b200_inflate_batch_dev")

    def compress_batch_dev(self, d_in, d_in_off, d_in_len, n_files, level, d_out, cap, d_out_off, flags=0, stream=0):
        """n_files independent inputs -> n_files independent streams; returns the total compressed size."""
        total = ctypes.c_size_t()
        rc = lib().b200_deflate_compress_batch_dev(self._h, d_in, d_in_off, d_in_len, n_files, _level(level), flags, d_out, cap,
                                                   d_out_off, ctypes.byref(total), stream or None)
        if rc:
            raise B200Error(rc, "b200_deflate_compress_batch_dev")
        return total.value

    def adler32_dev(self, d_data, n, stream=0):
        """Adler-32 of n device bytes, computed on the GPU."""
        out = ctypes.c_uint32()
        rc = lib().b200_adler32_dev(self._h, d_data or None, n, ctypes.byref(out), None, stream or None)
        if rc:
            raise B200Error(rc, "b200_adler32_dev")
        return out.value

    def crc32_dev(self, d_data, n, stream=0):
        """CRC-32 of n device bytes, computed on the GPU."""
        out = ctypes.c_uint32()
        rc = lib().b200_crc32_dev(self._h, d_data or None, n, ctypes.byref(out), None, stream or None)
        if rc:
            raise B200Error(rc, "b200_crc32_dev")
        return out.value

    def publish_dev(self, h_pinned_dst, d_src, n_words, stream=0):
        """n_words u64 from device memory into pinned host memory by a kernel on `stream` (no copy engine involved)."""
        rc = lib().b200_publish_dev(self._h, h_pinned_dst, d_src, n_words, stream or None)
        if rc:
            raise B200Error(rc, "b200_publish_dev")

    @staticmethod
    def corpus_generate_dev(d_out, seed, first_chunk, n_chunks, stream=0):
        rc = lib().b200_corpus_generate_dev(d_out, seed, first_chunk, n_chunks, stream or None)
        if rc:
            raise B200Error(rc, "b200_corpus_generate_dev")
