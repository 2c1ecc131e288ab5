// common.hpp -- drop-in replacement for HyperBitGore/deflate.hpp's include/common.hpp.
//
// The reference's common.hpp holds the CPU toolbox (FlatHuffmanTree, RangeLookup, fixed tables).
// None of that exists here: the work happens in hand-written sm_100a CUDA kernels inside
// libb200deflate.so, reached through the C ABI declared in b200_deflate.h.  This header only loads
// that library (dlopen, once) so that -- like the reference -- "throw the include directory in your
// project" is all a user has to do; no link-time dependency is added.
//
// Library lookup: $B200_DEFLATE_LIB if set, else "libb200deflate.so" through the usual dlopen search
// (LD_LIBRARY_PATH, rpath).  There is no CPU fallback: if the library or a B200 is missing every call
// throws std::runtime_error.
#pragma once
#include <dlfcn.h>

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "b200_deflate.h"

#define KB32 32768   // kept for source compatibility (reference common.hpp:14); chunks here are 64 KiB

namespace b200_detail {

struct Api {
    void* handle = nullptr;
    decltype(&::b200_deflate_compress) compress = nullptr;
    decltype(&::b200_deflate_compress_into) compress_into = nullptr;
    decltype(&::b200_inflate) inflate = nullptr;
    decltype(&::b200_inflate_alloc) inflate_alloc = nullptr;
    decltype(&::b200_inflate_zlib) inflate_zlib = nullptr;
    decltype(&::b200_inflate_zlib_alloc) inflate_zlib_alloc = nullptr;
    decltype(&::b200_deflate_compress_view) compress_view = nullptr;
    decltype(&::b200_inflate_view) inflate_view = nullptr;
    decltype(&::b200_view_release) view_release = nullptr;
    decltype(&::b200_host_prefault) host_prefault = nullptr;
    decltype(&::b200_deflate_compress_file) compress_file = nullptr;
    decltype(&::b200_inflate_file) inflate_file = nullptr;
    decltype(&::b200_free) free_ = nullptr;
    decltype(&::b200_strerror) strerror_ = nullptr;
    decltype(&::b200_abi_version) abi_version = nullptr;
};

inline const Api& api() {
    static const Api a = [] {
        Api x;
        const char* path = std::getenv("B200_DEFLATE_LIB");
        x.handle = dlopen(path ? path : "libb200deflate.so", RTLD_NOW | RTLD_LOCAL);
        if (!x.handle)
            throw std::runtime_error(std::string("deflate.hpp (B200): cannot load libb200deflate.so: ") + dlerror() +
                                     " -- set B200_DEFLATE_LIB; there is no CPU fallback");
        auto sym = [&](const char* n) {
            void* p = dlsym(x.handle, n);
            if (!p) throw std::runtime_error(std::string("deflate.hpp (B200): missing symbol ") + n);
            return p;
        };
        x.compress = reinterpret_cast<decltype(x.compress)>(sym("b200_deflate_compress"));
        x.compress_into = reinterpret_cast<decltype(x.compress_into)>(sym("b200_deflate_compress_into"));
        x.inflate = reinterpret_cast<decltype(x.inflate)>(sym("b200_inflate"));
        x.inflate_alloc = reinterpret_cast<decltype(x.inflate_alloc)>(sym("b200_inflate_alloc"));
        x.inflate_zlib = reinterpret_cast<decltype(x.inflate_zlib)>(sym("b200_inflate_zlib"));
        x.inflate_zlib_alloc = reinterpret_cast<decltype(x.inflate_zlib_alloc)>(sym("b200_inflate_zlib_alloc"));
        x.compress_view = reinterpret_cast<decltype(x.compress_view)>(sym("b200_deflate_compress_view"));
        x.inflate_view = reinterpret_cast<decltype(x.inflate_view)>(sym("b200_inflate_view"));
        x.view_release = reinterpret_cast<decltype(x.view_release)>(sym("b200_view_release"));
        x.host_prefault = reinterpret_cast<decltype(x.host_prefault)>(sym("b200_host_prefault"));
        x.compress_file = reinterpret_cast<decltype(x.compress_file)>(sym("b200_deflate_compress_file"));
        x.inflate_file = reinterpret_cast<decltype(x.inflate_file)>(sym("b200_inflate_file"));
        x.free_ = reinterpret_cast<decltype(x.free_)>(sym("b200_free"));
        x.strerror_ = reinterpret_cast<decltype(x.strerror_)>(sym("b200_strerror"));
        x.abi_version = reinterpret_cast<decltype(x.abi_version)>(sym("b200_abi_version"));
        if (x.abi_version() != B200_DEFLATE_ABI_VERSION)
            throw std::runtime_error("deflate.hpp (B200): libb200deflate.so ABI version mismatch");
        return x;
    }();
    return a;
}

[[noreturn]] inline void fail(int code) { throw std::runtime_error(api().strerror_(code)); }

inline std::vector<uint8_t> take(void* p, size_t n) {
    std::vector<uint8_t> v;
    try {
        v.assign(static_cast<uint8_t*>(p), static_cast<uint8_t*>(p) + n);
    } catch (...) {
        api().free_(p);
        throw;
    }
    api().free_(p);
    return v;
}

// the library's pinned arena -> the vector the reference API returns: the vector's fresh pages are faulted in from several
// threads first (b200_host_prefault), then one pass (assign), then the arena is released
inline std::vector<uint8_t> take_view(const void* p, size_t n) {
    struct Release { ~Release() { api().view_release(); } } release;
    std::vector<uint8_t> v;
    v.reserve(n);
    api().host_prefault(v.data(), n);
    v.assign(static_cast<const uint8_t*>(p), static_cast<const uint8_t*>(p) + n);
    return v;
}

inline std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Failed to read file " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

inline void write_file(const std::string& path, const uint8_t* p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n));
    f.flush();
    if (!f) throw std::runtime_error("Failed to write file " + path);
}

}  // namespace b200_detail
