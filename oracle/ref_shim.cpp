// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Thin extern "C" shim around the UNMODIFIED reference headers, compiled from the sources where
// they lie (-I/root/reference/include, see oracle/Makefile) into oracle/_ref/libref_deflate.so.
// It is used (a) by tests/ to pin oracle/inflate_oracle.c and to decode GPU-compressed streams with
// the reference's own inflater, (b) by bench.py's cpu_baseline / --impl reference arm.
//
// Entry points wrap exactly the reference's public API:
//   deflate::compress(char*, size_t, int)                 reference include/deflate.hpp:779
//   inflate::decompress(void*, size_t, void*, size_t)     reference include/inflate.hpp:338
//   inflate::decompress(void*, size_t) -> vector          reference include/inflate.hpp:363
//   inflate::decompressZlib(void*, size_t) -> vector      reference include/inflate.hpp:352
// The reference prints "Code tree is over or under subscribed!" on std::cerr from
// common.hpp:399-400; ref_quiet() silences that stream.
#include "deflate.hpp"
#include "inflate.hpp"

#include <atomic>
#include <chrono>
#include <streambuf>
#include <thread>

namespace {
// Discards everything; stateless, so concurrent writers (the _mt helpers below) are harmless.
struct NullBuf : std::streambuf {
    int overflow(int c) override { return traits_type::not_eof(c); }
    std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
NullBuf g_sink;
std::streambuf* g_old = nullptr;
}  // namespace

extern "C" {

// Silence (on != 0) or restore (on == 0) the reference's std::cerr chatter.
void ref_quiet(int on) {
    if (on && !g_old) {
        g_old = std::cerr.rdbuf(&g_sink);
    } else if (!on && g_old) {
        std::cerr.rdbuf(g_old);
        g_old = nullptr;
    }
}

// deflate::compress(char*, size_t, int).  Returns the compressed size; copies min(size, cap) bytes.
// Returns -1 if the reference throws.
long long ref_compress(const void* in, size_t n, int level, void* out, size_t cap) {
    try {
        std::vector<uint8_t> v = deflate::compress((char*)in, n, level);
        size_t c = v.size() < cap ? v.size() : cap;
        if (out && c) std::memcpy(out, v.data(), c);
        return (long long)v.size();
    } catch (...) {
        return -1;
    }
}

// inflate::decompress(void*, size_t) -> vector.  Returns decoded size (copies min(size, cap));
// -1 if the reference throws std::runtime_error (truncated / garbage input), -2 on anything else.
long long ref_inflate(const void* in, size_t n, void* out, size_t cap) {
    try {
        std::vector<uint8_t> v = inflate::decompress((void*)in, n);
        size_t c = v.size() < cap ? v.size() : cap;
        if (out && c) std::memcpy(out, v.data(), c);
        return (long long)v.size();
    } catch (const std::runtime_error&) {
        return -1;
    } catch (...) {
        return -2;
    }
}

// inflate::decompress(void*, size_t, void*, size_t): caller buffer, silently truncating.
long long ref_inflate_into(const void* in, size_t n, void* out, size_t cap) {
    try {
        return (long long)inflate::decompress((void*)in, n, out, cap);
    } catch (const std::runtime_error&) {
        return -1;
    } catch (...) {
        return -2;
    }
}

// inflate::decompressZlib(void*, size_t) -> vector.
long long ref_inflate_zlib(const void* in, size_t n, void* out, size_t cap) {
    try {
        std::vector<uint8_t> v = inflate::decompressZlib((void*)in, n);
        size_t c = v.size() < cap ? v.size() : cap;
        if (out && c) std::memcpy(out, v.data(), c);
        return (long long)v.size();
    } catch (const std::runtime_error&) {
        return -1;
    } catch (...) {
        return -2;
    }
}

// CPU-baseline helpers: run the reference on `count` independent slices with `threads` host threads
// (the reference is single-threaded and re-entrant, SURVEY.md 8(b) "Threading").  Slice i is
// in + offs[i], lens[i] bytes.  Returns wall seconds; out_sizes[i] receives each compressed size.
double ref_compress_slices_mt(const void* in, const size_t* offs, const size_t* lens, size_t count,
                              int level, int threads, long long* out_sizes) {
    std::atomic<size_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= count) break;
                long long r = -1;
                try {
                    std::vector<uint8_t> v = deflate::compress((char*)in + offs[i], lens[i], level);
                    r = (long long)v.size();
                } catch (...) {
                }
                if (out_sizes) out_sizes[i] = r;
            }
        });
    }
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Same for inflate::decompress(void*, size_t, void*, size_t) over independent streams.
double ref_inflate_slices_mt(const void* in, const size_t* offs, const size_t* lens, size_t count,
                             void* out, const size_t* out_offs, const size_t* out_caps, int threads,
                             long long* out_sizes) {
    std::atomic<size_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= count) break;
                long long r = -1;
                try {
                    r = (long long)inflate::decompress((char*)in + offs[i], lens[i],
                                                       (char*)out + out_offs[i], out_caps[i]);
                } catch (...) {
                }
                if (out_sizes) out_sizes[i] = r;
            }
        });
    }
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
