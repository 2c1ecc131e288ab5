"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/b200_deflate.h
declares; the drop-in headers compile; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200_deflate.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(b200):
    L = ctypes.CDLL(b200.lib_path())
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200_deflate.h but not exported"
    assert L.b200_abi_version() == 1


def test_bound_and_strerror(b200):
    assert b200.deflate_bound(0) >= 2
    for n in (1, 65535, 65536, 65537, 1 << 30):
        nch = (n + 65535) // 65536
        assert b200.deflate_bound(n) >= n + 10 * nch + 5
    L = b200.lib()
    assert b"beyond the alloted buffer size" in L.b200_strerror(1)   # the reference's message


def test_no_cpu_fallback(b200):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200.B200Error) as e:
        b200.compress(b"hello world")
    assert e.value.code == 4
    with pytest.raises(b200.B200Error):
        b200.decompress(b"\x03\x00")


def test_product_never_touches_oracle():
    """The product path must not include, import, link or dlopen anything under oracle/."""
    roots = [os.path.join(ROOT, "deflate.hpp_b200"), os.path.join(ROOT, "include")]
    for root in roots:
        for dirpath, _, files in os.walk(root):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    continue
                for line in open(os.path.join(dirpath, f), errors="ignore").read().splitlines():
                    s = line.strip()
                    if s.startswith(("//", "*", "/*", "#!", '"""')) or s.startswith("# "):
                        continue   # prose may cite the oracle; code may not reach it
                    if s.startswith(("#include", "import ", "from ")) or "CDLL(" in s or "dlopen(" in s:
                        assert "oracle" not in s, (f, line)


def test_dropin_headers_compile(tmp_path):
    exe = tmp_path / "test_dropin"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", str(exe),
                        os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"), "-lz", "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the reference's own callers use these exact expressions (test/libdeflate.cpp:123,134,180,195,215,231,290)
    probe = tmp_path / "sig.cpp"
    probe.write_text('''
#include "%s/include/deflate.hpp"
#include "%s/include/inflate.hpp"
int main() {
    std::vector<uint8_t> (*c1)(char*, size_t, int) = &deflate::compress;
    std::vector<uint8_t> (*c2)(std::vector<uint8_t>&, int) = &deflate::compress;
    size_t (*c3)(std::string, std::string, int) = &deflate::compress;
    size_t (*d1)(void*, size_t, void*, size_t) = &inflate::decompress;
    size_t (*d2)(void*, size_t, void*, size_t) = &inflate::decompressZlib;
    std::vector<uint8_t> (*d3)(void*, size_t) = &inflate::decompress;
    std::vector<uint8_t> (*d4)(void*, size_t) = &inflate::decompressZlib;
    std::vector<uint8_t> (*d5)(std::vector<uint8_t>) = &inflate::decompress;
    size_t (*d6)(std::string, std::string) = &inflate::decompress;
    // every integral level type the reference accepts must still resolve (ADVICE r1: the bool overloads made
    // unsigned / long / size_t ambiguous); only an exact bool picks the README form
    if (false) {
        std::vector<uint8_t> v;
        unsigned lu = 2; long ll = 2; size_t ls = 2; int64_t l64 = 2; short sh = 2; char ch = 2;
        deflate::compress(v, lu); deflate::compress(v, ll); deflate::compress(v, ls); deflate::compress(v, l64);
        deflate::compress(v, sh); deflate::compress(v, ch); deflate::compress(v, true);
        deflate::compress((char*)nullptr, (size_t)0, lu); deflate::compress((char*)nullptr, (size_t)0, false);
        deflate::compress("a", "b", ls); deflate::compress("a", "b", false);
    }
    return !(c1 && c2 && c3 && d1 && d2 && d3 && d4 && d5 && d6);
}
''' % (ROOT, ROOT))
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", str(probe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_host_prefault_touches_without_changing_zeroed_memory(b200):
    """b200_host_prefault (what the drop-in headers call on the std::vector they are about to fill) needs no GPU: it writes
    one zero byte per page from several threads -- fresh (zero) memory stays zero, sizes below its threshold and odd
    sizes / alignments are fine, NULL is ignored."""
    import mmap
    L = ctypes.CDLL(b200.lib_path())
    L.b200_host_prefault.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.b200_host_prefault.restype = None
    L.b200_host_prefault(None, 1 << 20)
    for n in (1, 4095, (4 << 20) + 123, (37 << 20) + 1):
        m = mmap.mmap(-1, n + 4096)
        buf = (ctypes.c_char * (n + 4096)).from_buffer(m)
        addr = ctypes.addressof(buf) + 3                   # unaligned start
        L.b200_host_prefault(addr, n)
        assert m[:n + 4096].count(b"\0") == n + 4096
        del buf
        m.close()
