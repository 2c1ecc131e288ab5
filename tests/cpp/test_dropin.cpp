// Mirrors the reference's test/libdeflate.cpp (main, :256-303) against the drop-in headers, with
// zlib 1.3 standing in for libdeflate (which the reference fetches from the network at configure
// time).  Unlike the reference's test, every step compares BYTES, not just "size != 0".
// Usage: test_dropin <fixture dir>      (exit code = number of failures)
#include <zlib.h>

#include <cstdio>
#include <iostream>

#include "../../include/deflate.hpp"
#include "../../include/inflate.hpp"

static int failures = 0;
#define CHECK(cond, what)                                              \
    do {                                                               \
        if (cond) std::cerr << "[PASS] " << what << "\n";              \
        else { std::cerr << "[FAIL] " << what << "\n"; failures++; }   \
    } while (0)

static std::vector<uint8_t> zlib_raw_deflate(const std::vector<uint8_t>& in, int level) {
    z_stream s{};
    deflateInit2(&s, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    std::vector<uint8_t> out(deflateBound(&s, in.size()) + 64);
    s.next_in = const_cast<Bytef*>(in.data()); s.avail_in = in.size();
    s.next_out = out.data(); s.avail_out = out.size();
    deflate(&s, Z_FINISH);
    out.resize(s.total_out);
    deflateEnd(&s);
    return out;
}
static bool zlib_raw_inflate(const std::vector<uint8_t>& in, std::vector<uint8_t>& out, size_t cap) {
    z_stream s{};
    inflateInit2(&s, -15);
    out.assign(cap + 16, 0);
    s.next_in = const_cast<Bytef*>(in.data()); s.avail_in = in.size();
    s.next_out = out.data(); s.avail_out = out.size();
    int rc = inflate(&s, Z_FINISH);
    out.resize(s.total_out);
    inflateEnd(&s);
    return rc == Z_STREAM_END;
}

// libdeflate.cpp:105-173 testDecompressionFile
static void round_trip(const std::string& path, int level) {
    std::vector<uint8_t> original = b200_detail::read_file(path);
    // step 1: third-party compress -> inflate.hpp decompress (caller buffer overload)
    std::vector<uint8_t> zc = zlib_raw_deflate(original, 1);
    std::vector<uint8_t> back(original.size());
    size_t n = inflate::decompress(zc.data(), zc.size(), back.data(), back.size());
    CHECK(n == original.size() && back == original, path + " L" + std::to_string(level) + " step1 zlib->inflate.hpp");
    // step 2: deflate.hpp compress -> third-party decompress
    std::vector<uint8_t> hc = deflate::compress(reinterpret_cast<char*>(original.data()), original.size(), level);
    std::vector<uint8_t> zout;
    bool ok = zlib_raw_inflate(hc, zout, original.size() * 2);
    CHECK(hc.size() != 0 && ok && zout == original, path + " L" + std::to_string(level) + " step2 deflate.hpp->zlib");
    // step 3: deflate.hpp compress -> inflate.hpp decompress
    std::vector<uint8_t> b2(original.size());
    n = inflate::decompress(hc.data(), hc.size(), b2.data(), b2.size());
    CHECK(n == original.size() && b2 == original, path + " L" + std::to_string(level) + " step3 deflate.hpp->inflate.hpp");
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? std::string(argv[1]) + "/" : "";
    // libdeflate.cpp:207-223 compareInflateLibVector
    {
        std::vector<uint8_t> original = b200_detail::read_file(dir + "test.bmp");
        std::vector<uint8_t> zc = zlib_raw_deflate(original, 1);
        std::vector<uint8_t> out = inflate::decompress(zc.data(), zc.size());
        CHECK(out == original, "inflate.hpp vector decompress matches original: test.bmp");
        CHECK(inflate::decompress(zc) == original, "inflate.hpp vector<uint8_t> overload: test.bmp");
    }
    // libdeflate.cpp:228-254 testInflateZlibFile
    for (const char* f : {"weird.dat", "zlib.dat"}) {
        std::vector<uint8_t> comp = b200_detail::read_file(dir + f);
        std::vector<uint8_t> out = inflate::decompressZlib(comp.data(), comp.size());
        std::vector<uint8_t> raw(comp.begin() + 2, comp.end()), zout;
        zlib_raw_inflate(raw, zout, out.size() + 1024);   // trailing Adler-32 -> not Z_STREAM_END-clean; compare bytes
        CHECK(!out.empty() && out == zout, std::string("inflate::decompressZlib matches zlib: ") + f);
        std::vector<uint8_t> buf(out.size());
        size_t n = inflate::decompressZlib(comp.data(), comp.size(), buf.data(), buf.size());
        CHECK(n == out.size() && buf == out, std::string("decompressZlib caller-buffer overload: ") + f);
        size_t half = out.size() / 2;
        n = inflate::decompressZlib(comp.data(), comp.size(), buf.data(), half);
        CHECK(n == half && std::memcmp(buf.data(), out.data(), half) == 0, std::string("silent truncation at out_size: ") + f);
    }
    // libdeflate.cpp:268-286 round trips at levels 0..3
    for (int level = 0; level <= 3; level++)
        for (const char* f : {"test.bmp", "tiny.bmp"}) round_trip(dir + f, level);
    // README form (bool)
    {
        std::vector<uint8_t> original = b200_detail::read_file(dir + "test.bmp");
        std::vector<uint8_t> fast = deflate::compress(original, false), better = deflate::compress(original, true);
        CHECK(inflate::decompress(fast) == original && inflate::decompress(better) == original, "bool overloads round-trip");
        CHECK(better.size() <= fast.size(), "better <= fast in size");
    }
    // libdeflate.cpp:288-296 file-path API round trip
    {
        const std::string tmp = "/tmp/b200_dropin_";
        deflate::compress(dir + "test.bmp", tmp + "deflated", 3);
        size_t n = inflate::decompress(tmp + "deflated", tmp + "inflated.bmp");
        CHECK(n == 21898 && b200_detail::read_file(tmp + "inflated.bmp") == b200_detail::read_file(dir + "test.bmp"),
              "test.bmp -> deflate(file) -> inflate(file) matches original");
        std::remove((tmp + "deflated").c_str());
        std::remove((tmp + "inflated.bmp").c_str());
    }
    // a multi-block FILE through the file-path overloads (the reference's inflate side fails on these):
    // 1.2 MB = 19 chunks, made from the fixture so that the test needs no other input
    {
        const std::string tmp = "/tmp/b200_dropin_big_";
        std::vector<uint8_t> one = b200_detail::read_file(dir + "test.bmp"), big;
        for (int k = 0; k < 55; k++) { big.insert(big.end(), one.begin(), one.end()); big.push_back((uint8_t)k); }
        b200_detail::write_file(tmp + "in", big.data(), big.size());
        for (int level : {0, 2, 3}) {
            size_t cn = deflate::compress(tmp + "in", tmp + "deflated", level);
            std::vector<uint8_t> comp = b200_detail::read_file(tmp + "deflated"), zout;
            bool ok = zlib_raw_inflate(comp, zout, big.size() + 64);
            CHECK(cn == comp.size() && ok && zout == big, "multi-block file, level " + std::to_string(level) + ": deflate(file) -> zlib");
            size_t n = inflate::decompress(tmp + "deflated", tmp + "inflated");
            CHECK(n == big.size() && b200_detail::read_file(tmp + "inflated") == big, "multi-block file: inflate(file) matches");
        }
        bool threw = false;
        try { deflate::compress(tmp + "no_such_file", tmp + "x", 2); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw, "missing input file throws std::runtime_error");
        for (const char* f : {"in", "deflated", "inflated"}) std::remove((tmp + f).c_str());
    }
    // error behaviour: truncated stream throws std::runtime_error (inflate.hpp:82)
    {
        std::vector<uint8_t> original = b200_detail::read_file(dir + "test.bmp");
        std::vector<uint8_t> zc = zlib_raw_deflate(original, 6);
        bool threw = false;
        try { inflate::decompress(zc.data(), zc.size() / 2); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw, "truncated input throws std::runtime_error");
    }
    std::cerr << (failures ? "=== FAILURES: " : "=== all passed: ") << failures << "\n";
    return failures;
}
