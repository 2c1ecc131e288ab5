// inflate.hpp -- drop-in replacement for HyperBitGore/deflate.hpp's include/inflate.hpp.
//
// Same global class name, same static method signatures (reference include/inflate.hpp:324-408):
//   size_t               inflate::decompressZlib(void* in, size_t in_size, void* out, size_t out_size)  :326
//   size_t               inflate::decompress    (void* in, size_t in_size, void* out, size_t out_size)  :338
//   std::vector<uint8_t> inflate::decompressZlib(void* in, size_t in_size)                              :352
//   std::vector<uint8_t> inflate::decompress    (void* in, size_t in_size)                              :363
//   std::vector<uint8_t> inflate::decompress    (std::vector<uint8_t> in)                               :376
//   size_t               inflate::decompress    (std::string file_path, std::string new_file)           :390
// Error behaviour follows the reference: truncated or garbage input throws std::runtime_error (the
// reference's message "Reading bits beyond the alloted buffer size!" for an overrun, inflate.hpp:82);
// the caller-buffer overloads silently truncate at out_size (:345) and return the bytes written.
// The file overload decodes multi-block files correctly (the reference re-reads 32 KB per block and
// fails on them).
#pragma once
#include "common.hpp"

class inflate {
public:
    static size_t decompress(void* in, size_t in_size, void* out, size_t out_size) {
        size_t n = 0;
        const int rc = b200_detail::api().inflate(in, in_size, out, out_size, &n, nullptr, 0);
        if (rc) b200_detail::fail(rc);
        return n;
    }
    static size_t decompressZlib(void* in, size_t in_size, void* out, size_t out_size) {
        size_t n = 0;
        const int rc = b200_detail::api().inflate_zlib(in, in_size, out, out_size, &n, nullptr, 0);
        if (rc) b200_detail::fail(rc);
        return n;
    }
    static std::vector<uint8_t> decompress(void* in, size_t in_size) {
        void* out = nullptr;
        size_t n = 0;
        const int rc = b200_detail::api().inflate_alloc(in, in_size, &out, &n, 0);
        if (rc) { if (out) b200_detail::api().free_(out); b200_detail::fail(rc); }
        return b200_detail::take(out, n);
    }
    static std::vector<uint8_t> decompressZlib(void* in, size_t in_size) {
        void* out = nullptr;
        size_t n = 0;
        const int rc = b200_detail::api().inflate_zlib_alloc(in, in_size, &out, &n, 0);
        if (rc) { if (out) b200_detail::api().free_(out); b200_detail::fail(rc); }
        return b200_detail::take(out, n);
    }
    static std::vector<uint8_t> decompress(std::vector<uint8_t> in) { return decompress(in.data(), in.size()); }
    static size_t decompress(std::string file_path, std::string new_file) {
        std::vector<uint8_t> in = b200_detail::read_file(file_path);
        std::vector<uint8_t> out = decompress(in.data(), in.size());
        b200_detail::write_file(new_file, out.data(), out.size());
        return out.size();
    }
};
