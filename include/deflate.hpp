// deflate.hpp -- drop-in replacement for HyperBitGore/deflate.hpp's include/deflate.hpp.
//
// Same global class name, same static method signatures (reference include/deflate.hpp:753-815):
//   size_t               deflate::compress(std::string file_path, std::string new_file, int level)   :755
//   std::vector<uint8_t> deflate::compress(char* data, size_t data_size, int level)                  :779
//   std::vector<uint8_t> deflate::compress(std::vector<uint8_t>& data, int level)                    :798
// Levels are the reference's ints (deflate.hpp:675-680): 0 stored, 1 Huffman only, 2 "fast" matcher,
// 3 "better" matcher.  The README describes the level as a bool ("true will enable the slower
// compression"); exact-match bool overloads provide that form: false -> fast (2), true -> better (3).
// (In the reference a bool would silently convert to level 0 / 1.)
//
// The work is done by sm_100a CUDA kernels behind the C ABI in b200_deflate.h; the output is one
// valid raw RFC 1951 stream, but not the same bytes the reference emits (independent 64 KiB chunks,
// see b200_deflate.h).  Like the reference, compress() never reports a data-dependent error; it throws
// std::runtime_error only when the GPU library is unusable.
#pragma once
#include "common.hpp"

class deflate {
public:
    static std::vector<uint8_t> compress(char* data, size_t data_size, int compression_level) {
        const void* view = nullptr;
        size_t n = 0;
        const int rc = b200_detail::api().compress_view(data, data_size, compression_level, 0, &view, &n);
        if (rc) b200_detail::fail(rc);
        return b200_detail::take_view(view, n);
    }
    static std::vector<uint8_t> compress(std::vector<uint8_t>& data, int compression_level) {
        return compress(reinterpret_cast<char*>(data.data()), data.size(), compression_level);
    }
    // Streams file_path -> new_file.  The reference always returns 0 here (deflate.hpp:681,751,776);
    // this returns the number of compressed bytes written.
    // The file is streamed through the GPU in 64 MiB slices (pinned, double-buffered; bounded memory for any file size).
    static size_t compress(std::string file_path, std::string new_file, int compression_level) {
        size_t out_n = 0;
        const int rc = b200_detail::api().compress_file(file_path.c_str(), new_file.c_str(), compression_level, 0, nullptr, &out_n);
        if (rc) b200_detail::fail(rc);
        return out_n;
    }

    // README form: bool selects fast (false) or better (true).  Templates constrained to exactly `bool`, so every
    // other integral level type (unsigned, long, size_t, int64_t ...) still resolves to the int overloads above,
    // as it does in the reference.
    template <class B, typename std::enable_if<std::is_same<B, bool>::value, int>::type = 0>
    static std::vector<uint8_t> compress(char* data, size_t data_size, B better) {
        return compress(data, data_size, better ? B200_LEVEL_BETTER : B200_LEVEL_FAST);
    }
    template <class B, typename std::enable_if<std::is_same<B, bool>::value, int>::type = 0>
    static std::vector<uint8_t> compress(std::vector<uint8_t>& data, B better) {
        return compress(data, better ? B200_LEVEL_BETTER : B200_LEVEL_FAST);
    }
    template <class B, typename std::enable_if<std::is_same<B, bool>::value, int>::type = 0>
    static size_t compress(std::string file_path, std::string new_file, B better) {
        return compress(std::move(file_path), std::move(new_file), better ? B200_LEVEL_BETTER : B200_LEVEL_FAST);
    }
};
