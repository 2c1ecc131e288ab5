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
        const void* view = nullptr;
        size_t n = 0;
        const int rc = b200_detail::api().inflate_view(in, in_size, 0, 0, &view, &n);
        if (rc) b200_detail::fail(rc);
        return b200_detail::take_view(view, n);
    }
    static std::vector<uint8_t> decompressZlib(void* in, size_t in_size) {
        if (!in || in_size < 2) b200_detail::fail(B200_E_OVERRUN);
        const void* view = nullptr;
        size_t n = 0;
        const int rc = b200_detail::api().inflate_view(in, in_size, 0, 1, &view, &n);
        if (rc) b200_detail::fail(rc);
        return b200_detail::take_view(view, n);
    }
    static std::vector<uint8_t> decompress(std::vector<uint8_t> in) { return decompress(in.data(), in.size()); }
    static size_t decompress(std::string file_path, std::string new_file) {
        size_t out_n = 0;
        const int rc = b200_detail::api().inflate_file(file_path.c_str(), new_file.c_str(), 0, nullptr, &out_n);
        if (rc) b200_detail::fail(rc);
        return out_n;
    }
};
