// e2e_dropin harness for bench.py: times what a USER of the drop-in headers gets --
//   deflate::compress(char*, size_t, int)               (reference include/deflate.hpp:779)
//   inflate::decompress(void*, size_t, void*, size_t)   (reference include/inflate.hpp:338)
// on ordinary pageable memory (std::vector / malloc), wall clock around each call, host <-> device copies,
// staging and the returned std::vector included.  Not a test: bench.py compiles it with g++ and reads the JSON
// line it prints.  Usage: bench_dropin <input file> <level> <reps>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "../../include/deflate.hpp"
#include "../../include/inflate.hpp"

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s <input file> <level> <reps>\n", argv[0]); return 2; }
    const int level = std::atoi(argv[2]);
    const int reps = std::atoi(argv[3]);
    std::vector<uint8_t> in = b200_detail::read_file(argv[1]);
    const size_t n = in.size();
    try {
        std::vector<uint8_t> comp;
        double c_best = 1e30, c_sum = 0;
        for (int i = 0; i <= reps; i++) {                 // first call = warm-up (context creation, workspace growth)
            const double t0 = now();
            comp = deflate::compress(reinterpret_cast<char*>(in.data()), n, level);
            const double t = now() - t0;
            if (i) { c_sum += t; if (t < c_best) c_best = t; }
        }
        std::vector<uint8_t> back(n);
        std::memset(back.data(), 0, n);                   // touch the pages: a caller's buffer normally is
        double d_best = 1e30, d_sum = 0;
        size_t got = 0;
        for (int i = 0; i <= reps; i++) {
            const double t0 = now();
            got = inflate::decompress(comp.data(), comp.size(), back.data(), back.size());
            const double t = now() - t0;
            if (i) { d_sum += t; if (t < d_best) d_best = t; }
        }
        const bool ok = got == n && std::memcmp(back.data(), in.data(), n) == 0;
        std::printf("{\"n\": %zu, \"comp\": %zu, \"level\": %d, \"reps\": %d, \"compress_s\": %.6f, \"compress_best_s\": %.6f, "
                    "\"inflate_s\": %.6f, \"inflate_best_s\": %.6f, \"round_trip\": %s}\n",
                    n, comp.size(), level, reps, c_sum / reps, c_best, d_sum / reps, d_best, ok ? "true" : "false");
    } catch (const std::exception& e) {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    return 0;
}
