// b200_deflate.cu -- C ABI (include/b200_deflate.h) and host orchestration of the sm_100a kernels.
//
// Pipeline (compress):  K1 lz77_kernel -> K2 huffman_kernel -> K3 scan_sizes_kernel -> K4 encode_kernel,
// per batch of chunks, all on one stream; chunk sizes are exact after K2 so K4 writes every chunk at
// its final byte offset (no compaction pass).
// Pipeline (inflate, single stream): find_sync (count, scan, write) -> per group of <= 32768 candidate chunks:
// classify -> segments (pass A: one thread per 4 KiB segment of every indexed chunk) -> fallback (one warp per
// chunk pass A could not take) -> copy (pass B: one warp per chunk applies the op lists) -> validate (optimistic
// layout chunk i -> out + i * 64 KiB); a stream that does not have this library's chunk structure falls back to
// one warp decoding it sequentially (inflate_batch_kernel with one stream).  Host buffers: slices of the input
// are decoded as they arrive while earlier output is on its way back (inflate_host_pipelined).
// There is no CPU fallback anywhere in this file.
#include "../../include/b200_deflate.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <fcntl.h>
#include <future>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "checksum.cuh"
#include "corpus.cuh"
#include "encode.cuh"
#include "huffman.cuh"
#include "inflate.cuh"
#include "inflate_tp.cuh"
#include "inflate_foreign.cuh"
#include "lz77.cuh"

using namespace b200;

namespace {

std::atomic<uint64_t> g_launches{0};

#define CK(expr)                                                                                 \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            if (getenv("B200_DEBUG"))                                                            \
                fprintf(stderr, "[b200] %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                     \
            return B200_E_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define LAUNCHED()                       \
    do {                                 \
        g_launches.fetch_add(1);         \
        CK(cudaGetLastError());          \
    } while (0)
// bracket a launch with profiling events when the context has profiling switched on
#define PROF_BEGIN(c, id, st) (c)->prof.begin((id), (st))
#define PROF_END(c, st) (c)->prof.end((st))

enum KernelId { K_LZ77 = 0, K_HUFFMAN, K_SCAN, K_ENCODE, K_FIND_SYNC, K_INFLATE_CHUNKS, K_VALIDATE, K_INFLATE_BATCH, K_CORPUS,
                K_INFLATE_SYMBOLS, K_INFLATE_FALLBACK, K_INFLATE_COPY, K_INFLATE_CLASSIFY,
                K_F_FIND, K_F_COUNT, K_F_EMIT, K_F_COPY, K_F_WINDOW, K_F_RESOLVE, K_COUNT };
const char* const kKernelNames[K_COUNT] = {"lz77_kernel", "huffman_kernel", "scan_sizes_kernel", "encode_kernel",
                                           "find_sync_kernel", "inflate_chunks_kernel", "validate_chunks_kernel",
                                           "inflate_batch_kernel", "corpus_kernels", "inflate_segments_kernel",
                                           "inflate_fallback_kernel", "inflate_copy_kernel", "inflate_classify_kernel",
                                           "foreign_find_blocks_kernel", "foreign_decode_kernel<count>", "foreign_decode_kernel<emit>",
                                           "foreign_copy_kernel", "foreign_window_kernels", "foreign_resolve_kernel"};

// Optional per-kernel timing: CUDA events recorded on the launching stream around every launch.
struct Prof {
    bool on = false;
    struct Rec { int id; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int id, cudaStream_t st) {
        if (!on) return;
        Rec r{id, get(), get()};
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end(cudaStream_t st) {
        if (!on) return;
        cudaEventRecord(recs.back().b, st);
    }
    void clear() {
        for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); }
        recs.clear();
    }
    void destroy() {
        clear();
        for (auto e : pool) cudaEventDestroy(e);
        pool.clear();
    }
};

// Public entry points run on the context's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;             // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(c)                        \
    DeviceGuard dev_guard__((c)->device);   \
    if (!dev_guard__.ok) { cudaGetLastError(); return B200_E_CUDA; }

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t need) {
        if (need <= cap) return B200_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = need + need / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            if (cudaMalloc(&p, need) != cudaSuccess) { cudaGetLastError(); return B200_E_NOMEM; }
            want = need;
        }
        cap = want;
        return B200_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct b200_ctx {
    int device = 0;
    bool attrs_set = false;
    uint32_t batch_chunks = 16384;   // chunks per compress batch (B200_BATCH_CHUNKS): 1 GiB of input, 4.4 GiB of scratch -- measured 4096 /
                                     // 8192 / 16384: 80.0 / 82.5 / 83.9 GB/s (fewer tails of the persistent matcher, fewer launches)
    // "better" level: chain depth / good-enough length.  0 = by input size: small inputs (< 1 MiB, where the time does not
    // matter) search like the reference does -- its level 3 looks at EVERY earlier position (deflate.hpp:280-296), and on
    // test.bmp that is worth 5 % -- large ones use 6 / 32 (the depth sweep in profiles/: 6.6 -> 21 GB/s for +1 % size)
    uint32_t better_depth = 0, better_nice = 0;
    // compress scratch
    Buf tok, ntok, hist, codes, hdr, desc, sizes, offsets, total;
    // inflate scratch
    Buf counts, woffs, cand, res, result, one_off, counter, cand16, ops, tpres, segnops, chunk_list, group_cnt, sync_cache, adler_parts, batch_nch, batch_first, batch_srcs;
    Buf f_cand, f_starts, f_stops, f_res, f_base, f_opsbase, f_sym, f_ops, f_gmap, f_gwin, f_slabs, f_misc, f_order;   // foreign streams
    bool size_probe = false;                  // set by inflate_host while it does not know the decoded size yet
    size_t foreign_min = (size_t)256 << 10;   // streams shorter than this stay with the one-warp decoder (B200_FOREIGN_MIN)
    uint32_t foreign_group = 0;               // units per window-propagation group (0 = auto); B200_FOREIGN_GROUP
    int foreign_tab = 0;                      // decode-kernel variant (inflate_foreign.cuh): 0 = 9/8-bit tables x 128 threads per SM,
                                              // 1 = 8/7 x 256, 2 = 8/6 x 320, 3 = 7/6 x 448; B200_FOREIGN_TAB
    std::vector<cudaEvent_t> group_events;
    cudaStream_t s_side = nullptr;   // inflate: copy pass of group g while group g + 1 is in pass A
    uint64_t inflate_group_chunks = 0;   // 0 = auto (32768 chunks); B200_INFLATE_GROUP
    bool inflate_overlap = false;        // B200_INFLATE_OVERLAP=1: copy pass of group g on a side stream while group g + 1 is in
                                         // pass A (measured: +1 % on a 4 GiB stream, so off by default)
    bool with_index = true;          // B200_NO_INDEX=1 / B200_F_NO_INDEX: no segment index in front of full chunks
    bool with_split = true;          // B200_NO_SPLIT=1: never cut a chunk into two blocks (huffman.cuh: block splitting)
    int sg_occ = 12, copy_occ = 8;   // resident CTAs per SM the two inflate passes are compiled for (tuning knobs; 8 x 4 warps at
                                     // 64 registers: no spills in the lane-local copies, measured best of 6 / 8 / 12)
    uint32_t sg_pad = 0;             // B200_SG_PAD: extra dynamic shared memory per pass-A CTA (occupancy experiments)
    unsigned copy_tune = 1;          // bit 0: prefetch the next step's sources into L2
    bool batch_two_pass = false;     // B200_BATCH_TP=1: batch inflate through the two-pass path, one THREAD per stream (measured
                                     // slower than one warp per stream on 1-64 KiB zlib streams: 15 vs 75 GB/s; kept for
                                     // workloads with very many tiny streams)
    bool inflate_warp_path = false;  // B200_INFLATE_WARP=1: the one-warp-per-unit decoder only (A/B comparisons)
    uint32_t num_sms = 148;
    uint32_t lzf_grid = 148 * 2;     // persistent two-phase matcher: SMs x resident CTAs
    uint32_t lzf_adaptive = 1;       // skip the match search where sample tiles find nothing (lz77.cuh); B200_LZF_ADAPTIVE=0: off
    uint32_t inf_grid = 148 * 7;     // persistent inflate grid: SMs x resident CTAs
    // host-API staging
    Buf d_in, d_out;
    cudaStream_t stream = nullptr;   // used by the host-buffer API (kernels)
    cudaStream_t s_in = nullptr, s_out = nullptr;   // host-buffer API: H2D / D2H copy streams
    std::vector<cudaEvent_t> events;
    uint64_t* mailbox = nullptr;     // pinned: per-slice end offsets
    size_t mailbox_cap = 0;
    unsigned long long* ctl = nullptr;   // pinned, 16 words: results the pipelined host inflate reads (publish_kernel)
    size_t host_inflate_slice = (size_t)128 << 20;  // host-buffer inflate: bytes of input per pipeline slice (0 = off; measured 32 / 64 /
                                                    // 128 / 256 MiB: 25.5 / 42.1 / 44.2 / 42.8 GB/s -- smaller groups of chunks run as
                                                    // partial waves, larger ones start the way back later); B200_HOST_INFLATE_SLICE
    uint32_t host_slice_chunks = 1024;   // 64 MiB: measured best (smaller slices starve the persistent matcher)
    // pageable caller memory: a ring of pinned staging buffers, filled / drained by a few host threads (host-buffer API)
    static constexpr int STG_N = 4;
    static constexpr size_t STG_BYTES = (size_t)32 << 20;      // per slot; a piece is copied by a few short-lived host threads, so few large
                                                               // pieces (8 MiB pieces: the thread start-up alone cost ~40 ms per GiB)
    void* stg_in[STG_N] = {};
    void* stg_out[STG_N] = {};
    cudaEvent_t stg_in_ev[STG_N] = {}, stg_out_ev[STG_N] = {};
    bool stg_in_busy[STG_N] = {};
    int stg_threads = 0;                            // 0 = auto (B200_STAGE_THREADS); 1 disables the thread fan-out
    bool stage_pageable = true;                     // B200_STAGE=0: hand pageable memory to cudaMemcpyAsync as round 1 did
    void* arena = nullptr;                          // pinned output arena of the *_view calls
    size_t arena_cap = 0;
    bool view_locked = false;
    // file API: pinned host staging (two slices in, two out), device ring
    void* pin_in[2] = {nullptr, nullptr};
    void* pin_out[2] = {nullptr, nullptr};
    size_t pin_in_cap[2] = {0, 0}, pin_out_cap[2] = {0, 0};
    Buf file_in[2], file_out[2];
    size_t file_slice = (size_t)64 << 20;           // bytes of input per slice (B200_FILE_SLICE; whole chunks)
    cudaEvent_t file_ev[2] = {nullptr, nullptr};
    std::mutex mu;
    Prof prof;
};

namespace {

int set_attrs(b200_ctx* c) {
    if (c->attrs_set) return B200_OK;
    CK(cudaFuncSetAttribute(lz77_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LZF_SMEM_BYTES));
    CK(cudaFuncSetAttribute(lz77_better_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LZB_SMEM_BYTES));
    CK(cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM_BYTES));
    CK(cudaFuncSetAttribute(inflate_segments_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_BYTES + 65536));
    CK(cudaFuncSetAttribute(inflate_segments_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_BYTES));
    CK(cudaFuncSetAttribute(inflate_symbols_kernel<BatchUnits>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM_BYTES));
#define FD_ATTR(LB, DB, NTH)                                                                                                    \
    CK(cudaFuncSetAttribute(foreign_decode_kernel<false, LB, DB, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fd_smem_bytes(LB, DB, NTH))); \
    CK(cudaFuncSetAttribute(foreign_decode_kernel<true, LB, DB, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fd_smem_bytes(LB, DB, NTH)));
    FD_ATTR(9, 8, 128) FD_ATTR(8, 7, 256) FD_ATTR(8, 6, 320) FD_ATTR(7, 6, 448)
#undef FD_ATTR
    CK(cudaFuncSetAttribute(inflate_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_OUT_BYTES + 16));
    CK(cudaFuncSetAttribute(foreign_window_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FW_SMEM_BYTES));
    CK(cudaFuncSetAttribute(foreign_window_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FW_SMEM_BYTES));
    CK(cudaFuncSetAttribute(foreign_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FW_SMEM_BYTES));
    c->attrs_set = true;
    return B200_OK;
}

// per-call override of the context's segment-index switch (B200_F_NO_INDEX)
struct IndexGuard {
    b200_ctx* c; bool saved;
    IndexGuard(b200_ctx* c_, unsigned flags) : c(c_), saved(c_->with_index) { if (flags & B200_F_NO_INDEX) c->with_index = false; }
    ~IndexGuard() { c->with_index = saved; }
};

// ---- batch compression helpers ------------------------------------------------------------------------------
// chunks per input (an empty input still gets one chunk: its stream is the two bytes 03 00) and the space the
// batch needs in the worst case
__global__ void batch_count_kernel(const uint64_t* __restrict__ in_len, uint64_t n, uint32_t* __restrict__ nch,
                                   unsigned long long* __restrict__ need) {
    const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const uint64_t len = in_len[f];
    const uint64_t k = len ? (len + CHUNK - 1) / CHUNK : 1;
    nch[f] = (uint32_t)k;
    atomicAdd(need, (unsigned long long)(len + 20 * ((len + CHUNK - 1) / CHUNK) + 34));      // = b200_deflate_bound(len)
}
__global__ void batch_srcs_kernel(const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                                  const uint64_t* __restrict__ first, uint64_t n, ChunkSrc* __restrict__ srcs) {
    const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const uint64_t len = in_len[f], off = in_off[f], c0 = first[f], k = first[f + 1] - c0;
    for (uint64_t j = 0; j < k; j++) {
        ChunkSrc s;
        s.off = off + j * CHUNK;
        s.clen = (uint32_t)(len - j * CHUNK < CHUNK ? len - j * CHUNK : CHUNK);
        s.last = (j + 1 == k ? 1u : 0u) | (len < SMALL_INPUT_BYTES ? 2u : 0u);
        srcs[c0 + j] = s;
    }
}
__global__ void batch_offsets_kernel(const uint64_t* __restrict__ first, const uint64_t* __restrict__ chunk_off, uint64_t n,
                                     uint64_t* __restrict__ out_off) {
    const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f <= n) out_off[f] = chunk_off[first[f]];
}

b200_ctx* g_default = nullptr;
std::mutex g_default_mu;

int default_ctx(b200_ctx** out) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default) {
        int dev = 0;
        if (const char* e = getenv("B200_DEVICE")) dev = atoi(e);
        else if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return B200_E_CUDA; }
        int rc = b200_ctx_create(dev, &g_default);
        if (rc) return rc;
    }
    *out = g_default;
    return B200_OK;
}


// Control words between host and device WITHOUT the copy engines.  A cudaMemcpyAsync of 8 bytes is queued on the same
// DMA engine as the bulk copies of the host-buffer pipeline and waits for whatever slice is on the wire (a 256 MiB slice:
// 5 ms) -- and the compute stream waits with it.  So the pipelined host inflate passes its few words through kernels:
// the device writes results straight into pinned (UVA-mapped) host memory, the host passes values as kernel arguments.
__global__ void publish_kernel(volatile unsigned long long* __restrict__ host, const unsigned long long* __restrict__ a, uint32_t na,
                               const unsigned long long* __restrict__ b, uint32_t nb) {
    const uint32_t t = threadIdx.x;
    if (t < na) host[t] = a[t];
    else if (t < na + nb) host[t] = b[t - na];
    __threadfence_system();
}
__global__ void poke_kernel(unsigned long long* p0, unsigned long long v0, unsigned long long* p1, unsigned long long v1,
                            unsigned long long* p2, unsigned long long v2, unsigned long long* p3, unsigned long long v3) {
    if (p0) *p0 = v0;
    if (p1) *p1 = v1;
    if (p2) *p2 = v2;
    if (p3) *p3 = v3;
}
__global__ void zero_words_kernel(unsigned long long* __restrict__ p, uint64_t nwords) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) p[i] = 0;
}

// Two-pass fast path (inflate_tp.cuh) for a set of units: pass A (one thread per unit: symbols, literals,
// op lists), the one-warp decoder for the units pass A gave up on, pass B (one warp per unit: copies).
// `res` receives one TpResult per unit.  cnt: 32 zeroed u64 counter words ([1] fallback queue, [2] copy queue,
// [3] fallback flag, [4] number of indexed chunks, [8..23] = 32 u32 bucket counts); chunk_list: 32 x nunits u32.  Pass B is enqueued on `st_copy` after pass A's kernels on
// `st` (event ev_a); with st_copy != st it overlaps whatever follows on `st`.
template <class Units>
static int inflate_two_pass(b200_ctx* c, const Units& U, uint64_t nunits, unsigned flags, TpResult* res,
                            unsigned long long* cnt, uint32_t* chunk_list, uint16_t* segnops, cudaStream_t st,
                            cudaStream_t st_copy, cudaEvent_t ev_a) {
    if constexpr (std::is_same<Units, ChunkUnits>::value) {
        PROF_BEGIN(c, K_INFLATE_CLASSIFY, st);
        inflate_classify_kernel<<<(uint32_t)((nunits + 255) / 256), 256, 0, st>>>(U, res, chunk_list, (uint32_t*)(cnt + 8), cnt + 4, flags, cnt + 3);
        LAUNCHED();
        PROF_END(c, st);
    }
    PROF_BEGIN(c, K_INFLATE_SYMBOLS, st);
    if constexpr (std::is_same<Units, ChunkUnits>::value) {
        if (c->sg_occ >= 16)
            inflate_segments_kernel<16><<<(uint32_t)((nunits + SG_CHUNKS - 1) / SG_CHUNKS), SG_THREADS, SG_SMEM_BYTES, st>>>(
                U, chunk_list, (const uint32_t*)(cnt + 8), cnt + 4, res, segnops, flags, cnt + 3);
        else
            inflate_segments_kernel<12><<<(uint32_t)((nunits + SG_CHUNKS - 1) / SG_CHUNKS), SG_THREADS, SG_SMEM_BYTES + c->sg_pad, st>>>(
                U, chunk_list, (const uint32_t*)(cnt + 8), cnt + 4, res, segnops, flags, cnt + 3);
    } else
        inflate_symbols_kernel<Units><<<(uint32_t)((nunits + TP_THREADS - 1) / TP_THREADS), TP_THREADS, TP_SMEM_BYTES, st>>>(U, res, flags, cnt + 3);
    LAUNCHED();
    PROF_END(c, st);
    const uint64_t want = (nunits + INF_WARPS - 1) / INF_WARPS;
    const uint32_t grid = (uint32_t)(want < c->inf_grid ? want : c->inf_grid);
    PROF_BEGIN(c, K_INFLATE_FALLBACK, st);
    inflate_fallback_kernel<Units><<<grid, INF_THREADS, 0, st>>>(U, res, flags, cnt + 3, cnt + 1);
    LAUNCHED();
    PROF_END(c, st);
    if (st_copy != st) {
        CK(cudaEventRecord(ev_a, st));
        CK(cudaStreamWaitEvent(st_copy, ev_a, 0));
    }
    PROF_BEGIN(c, K_INFLATE_COPY, st_copy);
    if (c->copy_occ >= 12) {
        const uint64_t g12 = (uint64_t)(c->inf_grid / INF_MAX_CTAS_PER_SM) * 12;
        inflate_copy_kernel<Units, 12><<<(uint32_t)(want < g12 ? want : g12), INF_THREADS, 0, st_copy>>>(U, res, cnt + 2, c->copy_tune);
    } else if (c->copy_occ >= 8) {
        const uint64_t g8 = (uint64_t)(c->inf_grid / INF_MAX_CTAS_PER_SM) * 8;
        inflate_copy_kernel<Units, 8><<<(uint32_t)(want < g8 ? want : g8), INF_THREADS, 0, st_copy>>>(U, res, cnt + 2, c->copy_tune);
    } else {
        inflate_copy_kernel<Units, 6><<<grid, INF_THREADS, 0, st_copy>>>(U, res, cnt + 2, c->copy_tune);
    }
    LAUNCHED();
    PROF_END(c, st_copy);
    return B200_OK;
}

// Chunk mode: the candidate chunks are cut into groups of 32768 (2 GiB of output) so that op-list scratch is
// bounded (2 bytes per output byte of a GROUP, not of the stream; a ring of two slots).  Optionally group g's
// copy pass runs on a side stream while group g + 1 is in pass A.  cand[ncand] must hold n (sentinel).
static int inflate_chunks_two_pass(b200_ctx* c, const uint8_t* in, uint64_t n, const uint64_t* cand, uint64_t ncand,
                                   uint8_t* out, uint64_t cap, unsigned flags, cudaStream_t st) {
    int rc;
    uint64_t G = c->inflate_group_chunks;
    if (G == 0) G = 32768;     // 2 GiB of output, 4 GiB of op-list scratch per slot; smaller groups run as partial waves (measured)
    if (G > ncand) G = ncand;
    const uint64_t ngroups = (ncand + G - 1) / G;
    const uint64_t nslots = ngroups < 2 ? ngroups : 2;
    if ((rc = c->tpres.ensure(ncand * sizeof(TpResult)))) return rc;
    if ((rc = c->chunk_list.ensure(G * SG_BUCKETS * 4))) return rc;
    if ((rc = c->segnops.ensure(ncand * NSEG * 2))) return rc;
    if ((rc = c->ops.ensure(nslots * G * OPS_PER_CHUNK * 8))) return rc;
    if ((rc = c->group_cnt.ensure(ngroups * 256))) return rc;
    zero_words_kernel<<<(uint32_t)((ngroups * 32 + 255) / 256), 256, 0, st>>>((unsigned long long*)c->group_cnt.p, ngroups * 32);   // (not a memset:
    LAUNCHED();                                                                    // see publish_kernel)
    while (c->group_events.size() < 2 * ngroups) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->group_events.push_back(e);
    }
    const bool overlap = ngroups > 1 && c->inflate_overlap;
    cudaStream_t side = overlap ? c->s_side : st;
    for (uint64_t g = 0; g < ngroups; g++) {
        const uint64_t g0 = g * G;
        const uint64_t ng = ncand - g0 < G ? ncand - g0 : G;
        if (overlap && g >= nslots) CK(cudaStreamWaitEvent(st, c->group_events[2 * (g - nslots) + 1], 0));   // slot's ops consumed
        const uint64_t o0 = g0 * CHUNK;
        ChunkUnits U{in, n, cand + g0, ng, out + o0, cap > o0 ? cap - o0 : 0, (uint64_t*)c->ops.p + (g % nslots) * G * OPS_PER_CHUNK,
                     (const uint16_t*)c->segnops.p + g0 * NSEG};
        if ((rc = inflate_two_pass(c, U, ng, flags, (TpResult*)c->tpres.p + g0, (unsigned long long*)c->group_cnt.p + g * 32,
                                   (uint32_t*)c->chunk_list.p, (uint16_t*)c->segnops.p + g0 * NSEG, st, side,
                                   c->group_events[2 * g])))
            return rc;
        if (overlap) CK(cudaEventRecord(c->group_events[2 * g + 1], side));
    }
    if (overlap)
        for (uint64_t g = ngroups >= nslots ? ngroups - nslots : 0; g < ngroups; g++)
            CK(cudaStreamWaitEvent(st, c->group_events[2 * g + 1], 0));
    return B200_OK;
}

}  // namespace

extern "C" {

int b200_abi_version(void) { return B200_DEFLATE_ABI_VERSION; }

const char* b200_strerror(int code) {
    switch (code) {
        case B200_OK: return "ok";
        case B200_E_OVERRUN: return "Reading bits beyond the alloted buffer size!";   // the reference's message
        case B200_E_DATA: return "invalid deflate stream";
        case B200_E_OUTPUT: return "output buffer too small";
        case B200_E_CUDA: return "CUDA error or no sm_100 device (this library has no CPU path)";
        case B200_E_ARG: return "bad argument";
        case B200_E_NOMEM: return "out of memory";
        case B200_E_IO: return "file I/O error";
        default: return "unknown error";
    }
}

uint64_t b200_launch_count(void) { return g_launches.load(); }

int b200_ctx_create(int device, b200_ctx** ctx) {
    if (!ctx) return B200_E_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return B200_E_CUDA; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return B200_E_CUDA;   // kernels are built for sm_100a only
    b200_ctx* c = new (std::nothrow) b200_ctx();
    if (!c) return B200_E_NOMEM;
    c->device = device;
    DeviceGuard dev_guard__(device);
    if (!dev_guard__.ok) { cudaGetLastError(); delete c; return B200_E_CUDA; }
    c->num_sms = (uint32_t)prop.multiProcessorCount;
    c->inf_grid = (uint32_t)prop.multiProcessorCount * INF_MAX_CTAS_PER_SM;
    c->lzf_grid = (uint32_t)prop.multiProcessorCount * 2;
    // B200_RESERVE_SMS=k: leave k SMs out of the persistent matcher's grid so that concurrently running
    // communication kernels (NCCL send/recv of the multi-GPU gather) can be scheduled while it runs
    if (const char* e = getenv("B200_RESERVE_SMS")) {
        int v = atoi(e);
        if (v > 0 && v < prop.multiProcessorCount) c->lzf_grid = (uint32_t)(prop.multiProcessorCount - v) * 2;
    }
    if (const char* e = getenv("B200_BETTER_DEPTH")) { int v = atoi(e); if (v > 0) c->better_depth = (uint32_t)v; }
    if (const char* e = getenv("B200_BETTER_NICE")) { int v = atoi(e); if (v >= 3) c->better_nice = (uint32_t)v; }
    if (const char* e = getenv("B200_LZF_ADAPTIVE")) c->lzf_adaptive = atoi(e) != 0;
    if (const char* e = getenv("B200_NO_SPLIT")) c->with_split = atoi(e) == 0;
    if (const char* e = getenv("B200_NO_INDEX")) c->with_index = atoi(e) == 0;
    if (const char* e = getenv("B200_INFLATE_GROUP")) { long v = atol(e); if (v > 0) c->inflate_group_chunks = (uint64_t)v; }
    if (const char* e = getenv("B200_INFLATE_OVERLAP")) c->inflate_overlap = atoi(e) != 0;
    if (const char* e = getenv("B200_SG_PAD")) c->sg_pad = (uint32_t)atoi(e);
    if (const char* e = getenv("B200_SG_OCC")) c->sg_occ = atoi(e);
    if (const char* e = getenv("B200_COPY_OCC")) c->copy_occ = atoi(e);
    if (const char* e = getenv("B200_COPY_TUNE")) c->copy_tune = (unsigned)atoi(e);
    if (const char* e = getenv("B200_BATCH_TP")) c->batch_two_pass = atoi(e) != 0;
    if (const char* e = getenv("B200_INFLATE_WARP")) c->inflate_warp_path = atoi(e) != 0;
    if (const char* e = getenv("B200_FOREIGN_MIN")) { long long v = atoll(e); if (v >= 0) c->foreign_min = (size_t)v; }
    if (const char* e = getenv("B200_FOREIGN_GROUP")) { int v = atoi(e); if (v > 0) c->foreign_group = (uint32_t)v; }
    if (const char* e = getenv("B200_FOREIGN_TAB")) { int v = atoi(e); if (v >= 0 && v <= 3) c->foreign_tab = v; }
    if (const char* e = getenv("B200_STAGE")) c->stage_pageable = atoi(e) != 0;
    if (const char* e = getenv("B200_STAGE_THREADS")) c->stg_threads = atoi(e);
    if (const char* e = getenv("B200_FILE_SLICE")) { long long v = atoll(e); if (v >= (long long)CHUNK) c->file_slice = (size_t)v / CHUNK * CHUNK; }
    if (const char* e = getenv("B200_BATCH_CHUNKS")) { int v = atoi(e); if (v > 0) c->batch_chunks = (uint32_t)v; }
    if (const char* e = getenv("B200_HOST_INFLATE_SLICE")) { long long v = atoll(e); if (v >= 0) c->host_inflate_slice = (size_t)v; }
    if (const char* e = getenv("B200_HOST_SLICE_CHUNKS")) { int v = atoi(e); if (v > 0) c->host_slice_chunks = (uint32_t)v; }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_side, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); b200_ctx_destroy(c); return B200_E_CUDA; }
    int rc = set_attrs(c);
    if (rc) { b200_ctx_destroy(c); return rc; }
    *ctx = c;
    return B200_OK;
}

void b200_ctx_destroy(b200_ctx* c) {
    if (!c) return;
    DeviceGuard dev_guard__(c->device);
    Buf* all[] = {&c->tok, &c->ntok, &c->hist, &c->codes, &c->hdr, &c->desc, &c->sizes, &c->offsets, &c->total,
                  &c->counts, &c->woffs, &c->cand, &c->res, &c->result, &c->one_off, &c->counter, &c->cand16, &c->ops, &c->tpres, &c->segnops, &c->chunk_list, &c->group_cnt, &c->sync_cache, &c->adler_parts, &c->batch_nch, &c->batch_first, &c->batch_srcs,
                  &c->d_in, &c->d_out, &c->f_cand, &c->f_starts, &c->f_stops, &c->f_res, &c->f_base, &c->f_opsbase, &c->f_sym, &c->f_ops,
                  &c->f_gmap, &c->f_gwin, &c->f_slabs, &c->f_misc, &c->f_order};
    for (Buf* b : all) b->release();
    c->prof.destroy();
    for (auto e : c->events) cudaEventDestroy(e);
    for (auto e : c->group_events) cudaEventDestroy(e);
    if (c->s_side) cudaStreamDestroy(c->s_side);
    if (c->mailbox) cudaFreeHost(c->mailbox);
    if (c->ctl) cudaFreeHost(c->ctl);
    if (c->arena) cudaFreeHost(c->arena);
    for (int k = 0; k < b200_ctx::STG_N; k++) {
        if (c->stg_in[k]) cudaFreeHost(c->stg_in[k]);
        if (c->stg_out[k]) cudaFreeHost(c->stg_out[k]);
        if (c->stg_in_ev[k]) cudaEventDestroy(c->stg_in_ev[k]);
        if (c->stg_out_ev[k]) cudaEventDestroy(c->stg_out_ev[k]);
    }
    for (int k = 0; k < 2; k++) {
        if (c->pin_in[k]) cudaFreeHost(c->pin_in[k]);
        if (c->pin_out[k]) cudaFreeHost(c->pin_out[k]);
        if (c->file_ev[k]) cudaEventDestroy(c->file_ev[k]);
        c->file_in[k].release(); c->file_out[k].release();
    }
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int b200_ctx_profile(b200_ctx* c, int enable) {
    if (!c) return B200_E_ARG;
    if (enable) c->prof.clear();   // records survive a stop so they can be read
    c->prof.on = enable != 0;
    return B200_OK;
}

int b200_ctx_profile_read(b200_ctx* c, int kernel_id, double* total_ms, uint64_t* launches) {
    if (!c || kernel_id < 0 || kernel_id >= K_COUNT) return B200_E_ARG;
    ON_DEVICE(c);
    double ms = 0;
    uint64_t n = 0;
    for (auto& r : c->prof.recs) {
        if (r.id != kernel_id) continue;
        CK(cudaEventSynchronize(r.b));
        float t = 0;
        CK(cudaEventElapsedTime(&t, r.a, r.b));
        ms += t; n++;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    return B200_OK;
}

const char* b200_kernel_name(int kernel_id) {
    return (kernel_id >= 0 && kernel_id < K_COUNT) ? kKernelNames[kernel_id] : nullptr;
}

size_t b200_deflate_bound(size_t n) {
    const size_t nchunks = (n + CHUNK - 1) / CHUNK;
    return n + 20 * nchunks + 34;   // per chunk: two stored blocks (2 x 5) + separator (10); + 16 slack + 18 for gzip framing
}

void b200_free(void* p) { free(p); }

// One batch of chunks through K1..K4 on stream `st`.  bin/bn: this batch's first input byte and the
// bytes from there to the end of the buffer; offs: the chunk-offset array (global chunk indexing).
// stages: bit 0 = K1..K3 (tokenise, code, size, scan), bit 1 = K4 (encode + write)
static int compress_batch(b200_ctx* c, const uint8_t* bin, uint64_t bn, uint32_t nb, uint64_t b0, bool final_batch,
                          int level, uint64_t* offs, uint64_t* d_total, void* d_out, cudaStream_t st,
                          int stages = 3, const uint64_t* d_extra_base = nullptr, const ChunkSrc* srcs = nullptr,
                          const uint64_t* d_first_base = nullptr) {
    // d_first_base: where the FIRST batch's first chunk goes (zlib / gzip framing puts a header in front); later batches
    // continue at the previous batch's end
    // srcs: batch compression -- chunk k of this batch is described by srcs[k] (absolute offsets into bin)
    if (stages & 1) {
    if (level == 2) {
        int rc;
        if ((rc = c->cand16.ensure((size_t)c->lzf_grid * CHUNK * 2))) return rc;
        if ((rc = c->counter.ensure(64))) return rc;
        CK(cudaMemsetAsync(c->counter.p, 0, 8, st));
    }
    if (level >= 1) {
        PROF_BEGIN(c, K_LZ77, st);
        if (level == 3) {
            // bytes of the whole input (a batch says it per chunk, in ChunkSrc)
            const bool small = !srcs && b0 * CHUNK + bn < SMALL_INPUT_BYTES;
            lz77_better_kernel<<<nb, LZB_THREADS, LZB_SMEM_BYTES, st>>>(bin, bn, (uint32_t*)c->tok.p, (uint32_t*)c->ntok.p, (uint16_t*)c->hist.p,
                                                                     c->better_depth ? c->better_depth : 6u, c->better_nice ? c->better_nice : 32u,
                                                                     c->better_depth ? c->better_depth : 128u, c->better_nice ? c->better_nice : 258u,
                                                                     small ? 1u : 0u, srcs);
        }
        else if (level == 2) {
            const uint32_t grid = nb < c->lzf_grid ? nb : c->lzf_grid;
            lz77_fast_kernel<<<grid, LZF_THREADS, LZF_SMEM_BYTES, st>>>(bin, bn, nb, (uint32_t*)c->tok.p, (uint32_t*)c->ntok.p,
                                                                     (uint16_t*)c->hist.p, (uint16_t*)c->cand16.p,
                                                                     (unsigned int*)c->counter.p, srcs, c->lzf_adaptive);
        }
        else
            lz77_literal_kernel<<<nb, LZL_THREADS, 0, st>>>(bin, bn, (uint32_t*)c->tok.p, (uint32_t*)c->ntok.p, (uint16_t*)c->hist.p, srcs);
        LAUNCHED();
        PROF_END(c, st);
    }
    PROF_BEGIN(c, K_HUFFMAN, st);
    huffman_kernel<<<(nb + HUF_WARPS - 1) / HUF_WARPS, HUF_THREADS, 0, st>>>(
        (const uint16_t*)c->hist.p, bn, nb, level, final_batch ? 1 : 0, c->with_index ? 1 : 0, c->with_split ? 1 : 0, srcs, (uint32_t*)c->codes.p,
        (uint32_t*)c->hdr.p, (BlockDesc*)c->desc.p, (uint32_t*)c->sizes.p);
    LAUNCHED();
    PROF_END(c, st);
    PROF_BEGIN(c, K_SCAN, st);
    scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)c->sizes.p, nb, b0 ? offs + b0 : d_first_base,
                                                 offs + b0, d_total);
    LAUNCHED();
    PROF_END(c, st);
    }   // stages & 1
    if (stages & 2) {
        PROF_BEGIN(c, K_ENCODE, st);
        encode_kernel<<<nb, ENC_THREADS, ENC_SMEM_BYTES, st>>>(bin, (const uint32_t*)c->tok.p, (const uint32_t*)c->ntok.p,
                                                              (const uint32_t*)c->codes.p, (const uint32_t*)c->hdr.p,
                                                              (const BlockDesc*)c->desc.p, offs + b0, d_extra_base, (uint8_t*)d_out, srcs);
        LAUNCHED();
        PROF_END(c, st);
    }
    return B200_OK;
}

// ---- zlib (RFC 1950) / gzip (RFC 1952) framing around the raw stream ------------------------------------------
static const uint8_t kZlibHeader[2] = {0x78, 0x9C};                                        // CM 8, 32 KiB window, default level, no dictionary
static const uint8_t kGzipHeader[10] = {0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF};            // no name, no mtime, OS unknown
static const uint64_t kFrameLen[3] = {0, 2, 10};
static int frame_mode(unsigned flags) { return (flags & B200_F_GZIP) ? 2 : (flags & B200_F_ZLIB) ? 1 : 0; }

// checksum of d_data[0..n) into *d_sum (device): Adler-32 (mode 1) or CRC-32 (mode 2)
// d_seed (may be NULL, may be d_sum): checksum of the bytes before this buffer, for streaming over slices
static int checksum_dev(b200_ctx* c, const uint8_t* d_data, uint64_t n, int mode, uint32_t* d_sum, cudaStream_t st,
                        const uint32_t* d_seed = nullptr) {
    int rc;
    const uint64_t nblocks = (n + CHUNK - 1) / CHUNK;
    if (mode == 1) {
        if ((rc = c->adler_parts.ensure((nblocks + 1) * sizeof(AdlerPart) + 16))) return rc;
        AdlerPart* parts = (AdlerPart*)c->adler_parts.p;
        if (nblocks) {
            adler_partial_kernel<<<(uint32_t)nblocks, ADLER_THREADS, 0, st>>>(d_data, n, parts);
            LAUNCHED();
        }
        adler_fold_kernel<<<1, 32, 0, st>>>(parts, nblocks, n, d_sum, d_seed);
        LAUNCHED();
    } else {
        if ((rc = c->adler_parts.ensure((nblocks + 1) * sizeof(CrcPart) + 16))) return rc;
        CrcPart* parts = (CrcPart*)c->adler_parts.p;
        if (nblocks) {
            crc32_partial_kernel<<<(uint32_t)nblocks, CRC_THREADS, 0, st>>>(d_data, n, parts);
            LAUNCHED();
        }
        crc32_fold_kernel<<<1, CRC_FOLD_THREADS, 0, st>>>(parts, nblocks, d_seed, d_sum);
        LAUNCHED();
    }
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------
int b200_deflate_compress_dev(b200_ctx* c, const void* d_in, size_t n, int level, unsigned flags, void* d_out,
                              size_t cap, uint64_t* d_out_n, size_t* h_out_n, uint64_t* d_chunk_off,
                              void* stream_) {
    if (!c || (!d_in && n) || !d_out || level < 0 || level > 3) return B200_E_ARG;
    if (cap < b200_deflate_bound(n)) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    const bool final_here = !(flags & B200_F_NOT_LAST);
    const int frame = frame_mode(flags);
    if (frame && !final_here) return B200_E_ARG;          // a frame wraps a whole stream, not a shard of one
    const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
    IndexGuard guard(c, flags);
    int rc;
    if ((rc = c->total.ensure(64))) return rc;
    uint64_t* d_total = (uint64_t*)c->total.p;            // [0] running total, [2] frame header length, [4] checksum
    const uint64_t* d_first_base = nullptr;
    if (frame) {
        CK(cudaMemcpyAsync(d_out, frame == 1 ? kZlibHeader : kGzipHeader, kFrameLen[frame], cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_total + 2, &kFrameLen[frame], 8, cudaMemcpyHostToDevice, st));
        d_first_base = d_total + 2;
    }
    auto finish_frame = [&]() -> int {
        if (!frame) return B200_OK;
        int rc2 = checksum_dev(c, (const uint8_t*)d_in, n, frame, (uint32_t*)(d_total + 4), st);
        if (rc2) return rc2;
        write_trailer_kernel<<<1, 1, 0, st>>>((uint8_t*)d_out, d_total, (const uint32_t*)(d_total + 4), (uint64_t)n, frame);
        LAUNCHED();
        return B200_OK;
    };

    if (nchunks == 0) {
        // empty input: a final empty fixed block (BFINAL=1, BTYPE=01, EOB) = 03 00
        static const uint8_t empty_final[2] = {0x03, 0x00};
        static uint64_t tots[3][2] = {{0, 2}, {2, 4}, {10, 12}};
        const uint64_t tot = tots[frame][final_here ? 1 : 0];
        if (final_here) CK(cudaMemcpyAsync((uint8_t*)d_out + kFrameLen[frame], empty_final, 2, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_total, &tots[frame][final_here ? 1 : 0], 8, cudaMemcpyHostToDevice, st));
        if (d_chunk_off) CK(cudaMemcpyAsync(d_chunk_off, d_total, 8, cudaMemcpyDeviceToDevice, st));
        if ((rc = finish_frame())) return rc;
        if (d_out_n) CK(cudaMemcpyAsync(d_out_n, d_total, 8, cudaMemcpyDeviceToDevice, st));
        if (h_out_n) { CK(cudaStreamSynchronize(st)); *h_out_n = (size_t)(tot + (frame == 1 ? 4 : frame == 2 ? 8 : 0)); }
        return B200_OK;
    }

    const uint64_t B = nchunks < c->batch_chunks ? nchunks : c->batch_chunks;
    if ((rc = c->tok.ensure(B * CHUNK * 4))) return rc;
    if ((rc = c->ntok.ensure(B * NSEG * 4))) return rc;
    if ((rc = c->hist.ensure(B * NSEG * NSYM * 2))) return rc;
    if ((rc = c->codes.ensure(B * 2 * NSYM * 4))) return rc;
    if ((rc = c->hdr.ensure(B * 2 * HDR_WORDS * 4))) return rc;
    if ((rc = c->desc.ensure(B * sizeof(BlockDesc)))) return rc;
    if ((rc = c->sizes.ensure(B * 4))) return rc;
    uint64_t* offs = d_chunk_off;
    if (!offs) {
        if ((rc = c->offsets.ensure((nchunks + 1) * 8))) return rc;
        offs = (uint64_t*)c->offsets.p;
    }

    const uint8_t* in = (const uint8_t*)d_in;
    for (uint64_t b0 = 0; b0 < nchunks; b0 += B) {
        const uint32_t nb = (uint32_t)((nchunks - b0 < B) ? nchunks - b0 : B);
        const uint8_t* bin = in + b0 * CHUNK;
        const uint64_t bn = n - b0 * CHUNK;
        const bool last_batch = b0 + nb == nchunks;
        if ((rc = compress_batch(c, bin, bn, nb, b0, last_batch && final_here, level, offs, d_total, d_out, st, 3, nullptr, nullptr,
                                 d_first_base)))
            return rc;
    }
    if ((rc = finish_frame())) return rc;
    if (d_out_n) CK(cudaMemcpyAsync(d_out_n, d_total, 8, cudaMemcpyDeviceToDevice, st));
    if (h_out_n) {
        uint64_t tot = 0;
        CK(cudaMemcpyAsync(&tot, d_total, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        *h_out_n = (size_t)tot;
    }
    return B200_OK;
}

// Batch compression: n_files independent inputs -> n_files independent raw DEFLATE streams, one launch sequence.
int b200_deflate_compress_batch_dev(b200_ctx* c, const void* d_in, const uint64_t* d_in_off, const uint64_t* d_in_len,
                                    size_t n_files, int level, unsigned flags, void* d_out, size_t cap,
                                    uint64_t* d_out_off, size_t* h_total, void* stream_) {
    if (!c || !d_in_off || !d_in_len || !d_out || !d_out_off || level < 0 || level > 3) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    if (n_files == 0) {
        CK(cudaMemsetAsync(d_out_off, 0, 8, st));
        if (h_total) { CK(cudaStreamSynchronize(st)); *h_total = 0; }
        return B200_OK;
    }
    IndexGuard guard(c, flags);
    int rc;
    if ((rc = c->batch_nch.ensure(n_files * 4))) return rc;
    if ((rc = c->batch_first.ensure((n_files + 1) * 8))) return rc;
    if ((rc = c->total.ensure(32))) return rc;
    unsigned long long* d_tot = (unsigned long long*)c->total.p;        // [0] compressed total, [1] chunks, [2] bytes needed
    CK(cudaMemsetAsync(d_tot, 0, 32, st));
    const uint32_t g = (uint32_t)((n_files + 255) / 256);
    batch_count_kernel<<<g, 256, 0, st>>>(d_in_len, n_files, (uint32_t*)c->batch_nch.p, d_tot + 2);
    LAUNCHED();
    // the scan kernel takes 32-bit counts: fine below 2^32 inputs
    scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)c->batch_nch.p, (uint32_t)n_files, nullptr,
                                                 (uint64_t*)c->batch_first.p, (uint64_t*)d_tot + 1);
    LAUNCHED();
    unsigned long long h[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(h, d_tot, 24, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t nchunks = h[1];
    if (h[2] > cap) return B200_E_ARG;                                  // cap must cover sum of b200_deflate_bound(len)
    if ((rc = c->batch_srcs.ensure(nchunks * sizeof(ChunkSrc)))) return rc;
    ChunkSrc* srcs = (ChunkSrc*)c->batch_srcs.p;
    batch_srcs_kernel<<<g, 256, 0, st>>>(d_in_off, d_in_len, (const uint64_t*)c->batch_first.p, n_files, srcs);
    LAUNCHED();
    const uint64_t B = nchunks < c->batch_chunks ? nchunks : c->batch_chunks;
    if ((rc = c->tok.ensure(B * CHUNK * 4))) return rc;
    if ((rc = c->ntok.ensure(B * NSEG * 4))) return rc;
    if ((rc = c->hist.ensure(B * NSEG * NSYM * 2))) return rc;
    if ((rc = c->codes.ensure(B * 2 * NSYM * 4))) return rc;
    if ((rc = c->hdr.ensure(B * 2 * HDR_WORDS * 4))) return rc;
    if ((rc = c->desc.ensure(B * sizeof(BlockDesc)))) return rc;
    if ((rc = c->sizes.ensure(B * 4))) return rc;
    if ((rc = c->offsets.ensure((nchunks + 1) * 8))) return rc;
    uint64_t* offs = (uint64_t*)c->offsets.p;
    for (uint64_t b0 = 0; b0 < nchunks; b0 += B) {
        const uint32_t nb = (uint32_t)((nchunks - b0 < B) ? nchunks - b0 : B);
        if ((rc = compress_batch(c, (const uint8_t*)d_in, 0, nb, b0, false, level, offs, (uint64_t*)d_tot, d_out, st, 3, nullptr,
                                 srcs + b0)))
            return rc;
    }
    batch_offsets_kernel<<<(uint32_t)((n_files + 256) / 256), 256, 0, st>>>((const uint64_t*)c->batch_first.p, offs, n_files, d_out_off);
    LAUNCHED();
    if (h_total) {
        uint64_t tot = 0;
        CK(cudaMemcpyAsync(&tot, d_tot, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        *h_total = (size_t)tot;
    }
    return B200_OK;
}

// Staged variant for the multi-GPU gather fused into the encoder.  Stage 1 = K1..K3 for one batch
// (n <= batch size), leaves this shard's compressed byte count in *d_local_n.  The caller exchanges the
// counts (an 8-byte all_gather), computes where the shard starts in the joined stream, and stage 2 = K4
// writes every chunk at d_out + *d_base + its local offset; d_out may be another GPU's memory.
int b200_deflate_compress_stage1_dev(b200_ctx* c, const void* d_in, size_t n, int level, unsigned flags,
                                     uint64_t* d_local_n, void* stream_) {
    if (!c || !d_in || !n || level < 0 || level > 3 || !d_local_n) return B200_E_ARG;
    const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
    if (nchunks > c->batch_chunks) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    IndexGuard guard(c, flags);
    int rc;
    const uint64_t B = nchunks;
    if ((rc = c->total.ensure(16))) return rc;
    if ((rc = c->tok.ensure(B * CHUNK * 4))) return rc;
    if ((rc = c->ntok.ensure(B * NSEG * 4))) return rc;
    if ((rc = c->hist.ensure(B * NSEG * NSYM * 2))) return rc;
    if ((rc = c->codes.ensure(B * 2 * NSYM * 4))) return rc;
    if ((rc = c->hdr.ensure(B * 2 * HDR_WORDS * 4))) return rc;
    if ((rc = c->desc.ensure(B * sizeof(BlockDesc)))) return rc;
    if ((rc = c->sizes.ensure(B * 4))) return rc;
    if ((rc = c->offsets.ensure((nchunks + 1) * 8))) return rc;
    if ((rc = compress_batch(c, (const uint8_t*)d_in, n, (uint32_t)nchunks, 0, !(flags & B200_F_NOT_LAST), level,
                             (uint64_t*)c->offsets.p, (uint64_t*)c->total.p, nullptr, st, 1)))
        return rc;
    CK(cudaMemcpyAsync(d_local_n, c->total.p, 8, cudaMemcpyDeviceToDevice, st));
    return B200_OK;
}

int b200_deflate_compress_stage2_dev(b200_ctx* c, const void* d_in, size_t n, void* d_out, const uint64_t* d_base,
                                     void* stream_) {
    if (!c || !d_in || !n || !d_out) return B200_E_ARG;
    const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
    if (nchunks > c->batch_chunks || !c->offsets.p) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    return compress_batch(c, (const uint8_t*)d_in, n, (uint32_t)nchunks, 0, false, 0, (uint64_t*)c->offsets.p,
                          (uint64_t*)c->total.p, d_out, st, 2, d_base);
}

// ------------------------------------------------------------------------------------------------
int b200_inflate_batch_dev(b200_ctx* c, const void* d_in, const uint64_t* d_in_off, const uint64_t* d_in_len,
                           void* d_out, const uint64_t* d_out_off, const uint64_t* d_out_cap, uint64_t* d_out_len,
                           int32_t* d_status, size_t n_streams, unsigned flags, void* stream_) {
    if (!c || !d_in_off || !d_in_len || !d_out_off || !d_out_cap || !d_out_len || !d_status) return B200_E_ARG;
    if (n_streams == 0) return B200_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    int rc;
    if ((rc = c->counter.ensure(64))) return rc;
    if (c->batch_two_pass && !c->inflate_warp_path) {
        // op-list scratch is addressed by output offset (2 bytes of scratch per output byte): its size is
        // the span of the output regions, known only on the device -> one 8-byte readback
        if ((rc = c->result.ensure(64))) return rc;
        unsigned long long* d_span = (unsigned long long*)c->result.p + 4;
        CK(cudaMemsetAsync(d_span, 0, 8, st));
        batch_span_kernel<<<(uint32_t)((n_streams + 255) / 256), 256, 0, st>>>(d_out_off, d_out_cap, n_streams, d_span);
        LAUNCHED();
        unsigned long long span = 0;
        CK(cudaMemcpyAsync(&span, d_span, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if ((rc = c->ops.ensure((span / 4 + 2) * 8))) return rc;
        BatchUnits U{(const uint8_t*)d_in, d_in_off, d_in_len, (uint8_t*)d_out, d_out_off, d_out_cap, (uint64_t)n_streams,
                     (uint64_t*)c->ops.p};
        if ((rc = c->tpres.ensure(n_streams * sizeof(TpResult)))) return rc;
        if ((rc = c->counter.ensure(256))) return rc;
        CK(cudaMemsetAsync(c->counter.p, 0, 256, st));
        if ((rc = inflate_two_pass(c, U, n_streams, flags, (TpResult*)c->tpres.p, (unsigned long long*)c->counter.p, nullptr,
                                   nullptr, st, st, nullptr)))
            return rc;
        batch_results_kernel<<<(uint32_t)((n_streams + 255) / 256), 256, 0, st>>>((const TpResult*)c->tpres.p, n_streams, d_out_len, d_status);
        LAUNCHED();
        return B200_OK;
    }
    CK(cudaMemsetAsync(c->counter.p, 0, 8, st));
    const uint64_t want = (n_streams + INF_WARPS - 1) / INF_WARPS;
    const uint64_t grid = want < c->inf_grid ? want : c->inf_grid;
    PROF_BEGIN(c, K_INFLATE_BATCH, st);
    inflate_batch_kernel<<<(uint32_t)grid, INF_THREADS, 0, st>>>((const uint8_t*)d_in, d_in_off, d_in_len, (uint8_t*)d_out,
                                                               d_out_off, d_out_cap, d_out_len, d_status, n_streams, flags,
                                                               (unsigned long long*)c->counter.p);
    LAUNCHED();
    PROF_END(c, st);
    return B200_OK;
}

// number of candidates below `hi` (candidates are sorted): one thread, binary search
__global__ void count_below_kernel(const uint64_t* __restrict__ cand, uint64_t ncand, uint64_t hi, unsigned long long* __restrict__ out) {
    uint64_t lo = 0, up = ncand;
    while (lo < up) {
        const uint64_t mid = (lo + up) >> 1;
        if (cand[mid] < hi) lo = mid + 1; else up = mid;
    }
    *out = lo;
}

// Chunk-parallel decode of in[0..n) (the stream, or a window of it).  Chunk starts = byte 0 (if first_is_start) and the
// byte after every separator tail; chunks that START in [lo, hi) are decoded to out + k * 64 KiB.  If the window ends
// the stream (expect_final) the last chunk must carry BFINAL, otherwise every decoded chunk must be a full one that ends
// at the next chunk start (or exactly at n).  *valid = the optimistic layout held; *total = decoded bytes; *nunits =
// chunks decoded; *next_start = offset of the first chunk that was NOT decoded (n if none).  Synchronizes `st`.
static int inflate_chunked(b200_ctx* c, const uint8_t* in, uint64_t n, uint64_t lo, uint64_t hi, bool first_is_start, bool expect_final,
                           uint8_t* out, uint64_t cap, unsigned flags, cudaStream_t st, bool* valid, uint64_t* total,
                           uint64_t* nunits, uint64_t* next_start, Buf* grow_out = nullptr) {
    // grow_out: decode into this buffer instead of out / cap, grown to 64 KiB per chunk once the chunk count is known
    int rc;
    *valid = false; *total = 0;
    if (nunits) *nunits = 0;
    if (next_start) *next_start = n;
    if ((rc = c->result.ensure(64))) return rc;
    unsigned long long* d_result = (unsigned long long*)c->result.p;   // [0] valid, [1] total, [2] marks, [3] units below hi
    const uint64_t nwarps = (n + SYNC_REGION - 1) / SYNC_REGION;
    const uint64_t cand_cap = n / 64 + 1024;
    if ((rc = c->counts.ensure(nwarps * 4))) return rc;
    if ((rc = c->woffs.ensure((nwarps + 1) * 8))) return rc;
    if ((rc = c->cand.ensure((cand_cap + 2) * 8))) return rc;
    if ((rc = c->sync_cache.ensure(nwarps * SYNC_CACHE * 4))) return rc;
    uint64_t* cand0 = (uint64_t*)c->cand.p;
    const uint32_t g = (uint32_t)((nwarps * 32 + 255) / 256);
    PROF_BEGIN(c, K_FIND_SYNC, st);
    const uint64_t min_cand = lo ? lo - 1 : 0;         // a chunk start s is reported iff s > min_cand
    find_sync_kernel<false><<<g, 256, 0, st>>>(in, n, (uint32_t*)c->counts.p, (uint32_t*)c->sync_cache.p, nullptr, nullptr, 0, min_cand);
    LAUNCHED();
    PROF_END(c, st);
    PROF_BEGIN(c, K_SCAN, st);
    scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)c->counts.p, (uint32_t)nwarps, nullptr,
                                                 (uint64_t*)c->woffs.p, (uint64_t*)d_result + 2);
    LAUNCHED();
    PROF_END(c, st);
    uint64_t nmark = 0;
    CK(cudaMemcpyAsync(&nmark, d_result + 2, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nmark + 1 > cand_cap) return B200_OK;          // not ours: far too many separators
    // not ours either: no separator at all in more bytes than one chunk can take (a foreign stream; without this its first
    // 64 KiB would be decoded by the one-warp decoder just to find out)
    if (nmark == 0 && first_is_start && n > (uint64_t)CHUNK + INDEX_BYTES + 64) return B200_OK;
    CK(cudaMemsetAsync(cand0, 0, 8, st));               // cand0[0] = 0
    if (nmark) {
        PROF_BEGIN(c, K_FIND_SYNC, st);
        find_sync_kernel<true><<<g, 256, 0, st>>>(in, n, (uint32_t*)c->counts.p, (uint32_t*)c->sync_cache.p,
                                                  (const uint64_t*)c->woffs.p, cand0 + 1, nmark, min_cand);
        LAUNCHED();
        PROF_END(c, st);
    }
    const uint64_t n64 = n;
    CK(cudaMemcpyAsync(cand0 + nmark + 1, &n64, 8, cudaMemcpyHostToDevice, st));      // sentinel: the last chunk ends at n
    uint64_t* cand = first_is_start ? cand0 : cand0 + 1;
    const uint64_t ncand = first_is_start ? nmark + 1 : nmark;
    uint64_t units = ncand;
    if (hi < n && ncand) {
        count_below_kernel<<<1, 1, 0, st>>>(cand, ncand, hi, d_result + 3);
        LAUNCHED();
        CK(cudaMemcpyAsync(&units, d_result + 3, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    if (units == 0) { *valid = true; return B200_OK; }
    if (grow_out) {
        if ((rc = grow_out->ensure(units * CHUNK + 64))) return rc;
        out = (uint8_t*)grow_out->p;
        cap = units * CHUNK;
    }
    const bool ends_stream = expect_final && units == ncand;
    const unsigned long long init[2] = {1ull, 0ull};
    CK(cudaMemcpyAsync(d_result, init, 16, cudaMemcpyHostToDevice, st));
    if (c->inflate_warp_path) {
        if (!ends_stream || !first_is_start) return B200_OK;         // A/B path: whole streams only
        if ((rc = c->res.ensure(ncand * sizeof(ChunkResult)))) return rc;
        if ((rc = c->counter.ensure(64))) return rc;
        CK(cudaMemsetAsync(c->counter.p, 0, 8, st));
        PROF_BEGIN(c, K_INFLATE_CHUNKS, st);
        {
            const uint64_t want = (ncand + INF_WARPS - 1) / INF_WARPS;
            inflate_chunks_kernel<<<(uint32_t)(want < c->inf_grid ? want : c->inf_grid), INF_THREADS, 0, st>>>(
                in, n, cand, ncand, out, cap, (ChunkResult*)c->res.p, flags, (unsigned long long*)c->counter.p);
        }
        LAUNCHED();
        PROF_END(c, st);
        PROF_BEGIN(c, K_VALIDATE, st);
        validate_chunks_kernel<<<(uint32_t)((ncand + 255) / 256), 256, 0, st>>>(cand, ncand, (const ChunkResult*)c->res.p, d_result);
        LAUNCHED();
        PROF_END(c, st);
    } else {
        if ((rc = inflate_chunks_two_pass(c, in, n, cand, units, out, cap, flags, st))) return rc;
        PROF_BEGIN(c, K_VALIDATE, st);
        validate_units_kernel<<<(uint32_t)((units + 255) / 256), 256, 0, st>>>(cand, units, (uint64_t)n, (const TpResult*)c->tpres.p, d_result,
                                                                              ends_stream ? 1 : 0);
        LAUNCHED();
        PROF_END(c, st);
    }
    unsigned long long verdict[2] = {0, 0};
    uint64_t nxt = n;
    CK(cudaMemcpyAsync(verdict, d_result, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&nxt, cand + units, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (verdict[0] == 1) {
        *valid = true; *total = verdict[1];
        if (nunits) *nunits = units;
        if (next_start) *next_start = nxt;
    }
    return B200_OK;
}

static int inflate_foreign(b200_ctx* c, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, unsigned flags, cudaStream_t st,
                           bool* done, uint64_t* full);

// foreign_decode_kernel in the table-width / CTA-size variant the context asks for (c->foreign_tab, B200_FOREIGN_TAB)
extern "C++" {
template <bool EMIT>
static void launch_foreign_decode(b200_ctx* c, cudaStream_t st, const uint8_t* in, uint64_t n, const uint64_t* starts, const uint64_t* stops,
                                  uint64_t nunits, FUnitRes* res, const uint64_t* out_base, const uint64_t* ops_base, uint16_t* S, uint64_t* ops,
                                  unsigned flags, const uint32_t* order, unsigned long long* queue) {
#define FD_LAUNCH(LB, DB, NTH)                                                                                                       \
    {                                                                                                                                \
        const uint64_t want = (nunits + NTH - 1) / NTH;                                                                              \
        foreign_decode_kernel<EMIT, LB, DB, NTH><<<(uint32_t)(want < c->num_sms ? want : c->num_sms), NTH, fd_smem_bytes(LB, DB, NTH), st>>>( \
            in, n, starts, stops, nunits, res, out_base, ops_base, S, ops, flags, order, queue);                                     \
    }
    switch (c->foreign_tab) {
        case 0: FD_LAUNCH(9, 8, 128) break;
        case 2: FD_LAUNCH(8, 6, 320) break;
        case 3: FD_LAUNCH(7, 6, 448) break;
        default: FD_LAUNCH(8, 7, 256) break;
    }
#undef FD_LAUNCH
}
}  // extern "C++"

// Single stream.  Synchronizes `stream` internally (the candidate count and the validation verdict
// steer the launches).
int b200_inflate_dev(b200_ctx* c, const void* d_in, size_t n, void* d_out, size_t cap, uint64_t* d_out_n,
                     size_t* h_out_n, size_t* h_full_n, int32_t* d_status, unsigned flags, void* stream_) {
    if (!c || (!d_in && n) || (!d_out && cap)) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    int rc;
    const uint8_t* in = (const uint8_t*)d_in;
    int status = B200_OK;
    uint64_t full = 0;
    bool done = false, small_done = false;

    // ---- a short stream into a small buffer: one warp, output window in shared memory (inflate_small_kernel) ----
    // (size_probe: the host does not know the decoded size yet -- if it fits the window the call is done, else the
    // regular paths run; with a caller buffer that small there is nothing to lose, the decoder counts beyond it anyway)
    if (n && n <= SMALL_IN_BYTES && (cap <= SMALL_OUT_BYTES || c->size_probe) && !c->inflate_warp_path &&
        !getenv("B200_NO_SMALL_INFLATE") && !getenv("B200_INFLATE_SEQUENTIAL")) {
        if ((rc = c->one_off.ensure(64))) return rc;
        if (!c->ctl) CK(cudaHostAlloc((void**)&c->ctl, 16 * 8, cudaHostAllocDefault));
        const uint32_t scap = (uint32_t)(cap < SMALL_OUT_BYTES ? cap : SMALL_OUT_BYTES);
        PROF_BEGIN(c, K_INFLATE_BATCH, st);
        inflate_small_kernel<<<1, SMALL_THREADS, scap + 16, st>>>(in, n, (uint8_t*)d_out, scap, flags, (unsigned long long*)c->one_off.p, c->ctl + 8);
        LAUNCHED();
        PROF_END(c, st);
        CK(cudaStreamSynchronize(st));
        const uint64_t got = c->ctl[8];
        if (c->ctl[9] == SMALL_ST_INDEXED) {
            // this library's stream with a segment index: the chunked path below
        } else if (got <= scap || cap <= scap) {       // everything is in the buffer, or the caller's buffer truncates anyway
            full = got;
            status = (int)(int32_t)(c->ctl[9] & 0xFFFFFFFFu);
            done = true;
            small_done = true;
        }
    }
    if (!small_done && n >= 8 && !getenv("B200_INFLATE_SEQUENTIAL")) {
        // ---- this library's own chunked streams ----
        if ((rc = inflate_chunked(c, in, n, 0, n, true, true, (uint8_t*)d_out, cap, flags, st, &done, &full, nullptr, nullptr))) return rc;
        // ---- anything else that is big enough to be worth it: block-parallel (inflate_foreign.cuh) ----
        if (!done && !c->inflate_warp_path && !getenv("B200_NO_FOREIGN_PARALLEL"))
            if ((rc = inflate_foreign(c, in, n, (uint8_t*)d_out, cap, flags, st, &done, &full))) return rc;
    }
    if (!done) {
        // ---- sequential fallback: one warp, whole stream ----
        if ((rc = c->one_off.ensure(64))) return rc;
        uint64_t* d = (uint64_t*)c->one_off.p;   // [0] in_off [1] in_len [2] out_off [3] out_cap [4] out_len [5] status
        const uint64_t h[7] = {0, (uint64_t)n, 0, (uint64_t)cap, 0, 0, 0};   // [6] = work counter
        CK(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, st));
        PROF_BEGIN(c, K_INFLATE_BATCH, st);
        inflate_batch_kernel<<<1, INF_THREADS, 0, st>>>(in, d, d + 1, (uint8_t*)d_out, d + 2, d + 3, d + 4, (int32_t*)(d + 5), 1, flags,
                                                        (unsigned long long*)(d + 6));
        LAUNCHED();
        PROF_END(c, st);
        uint64_t r[2] = {0, 0};
        CK(cudaMemcpyAsync(r, d + 4, 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        full = r[0];
        status = (int)(int32_t)(r[1] & 0xFFFFFFFFu);
    }
    const uint64_t written = full < cap ? full : cap;
    if (d_out_n) CK(cudaMemcpyAsync(d_out_n, &written, 8, cudaMemcpyHostToDevice, st));
    if (d_status) { int32_t s32 = status; CK(cudaMemcpyAsync(d_status, &s32, 4, cudaMemcpyHostToDevice, st)); }
    if (d_out_n || d_status) CK(cudaStreamSynchronize(st));
    if (h_out_n) *h_out_n = (size_t)written;
    if (h_full_n) *h_full_n = (size_t)full;
    return status;
}

// A window of a longer stream of this library's format (multi-GPU inflate: every rank takes a byte range of the joined
// stream plus some slack and decodes the chunks that START inside its range).
int b200_inflate_shard_dev(b200_ctx* c, const void* d_in, size_t n, size_t lo, size_t hi, int first_is_start, int ends_stream,
                           void* d_out, size_t cap, size_t* h_out_n, size_t* h_nchunks, size_t* h_next_start,
                           unsigned flags, void* stream_) {
    if (!c || (!d_in && n) || (!d_out && cap) || !h_out_n) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    bool valid = false;
    uint64_t total = 0, units = 0, nxt = n;
    if (hi > n) hi = n;
    if (lo > hi) return B200_E_ARG;
    const int rc = inflate_chunked(c, (const uint8_t*)d_in, n, first_is_start ? 0 : lo, hi, first_is_start != 0, ends_stream != 0, (uint8_t*)d_out, cap, flags, st,
                                   &valid, &total, &units, &nxt);
    if (rc) return rc;
    if (!valid) return B200_E_DATA;                       // not a run of this library's chunks
    if (total > cap) return B200_E_OUTPUT;
    *h_out_n = (size_t)total;
    if (h_nchunks) *h_nchunks = (size_t)units;
    if (h_next_start) *h_next_start = (size_t)nxt;
    return B200_OK;
}

// order[] = unit indices by decreasing key, in O(n): keys are bucketed by their top bits (a comparison sort of 40 000 units
// costs ~1.5 ms of host time per pass while the GPU waits; the queue only needs "long units first", not a total order)
static void order_longest_first(const std::vector<uint64_t>& key, std::vector<uint32_t>& order) {
    const size_t n = key.size();
    order.resize(n);
    uint64_t mx = 1;
    for (size_t k = 0; k < n; k++) if (key[k] != ~0ull && key[k] > mx) mx = key[k];
    int shift = 0;
    while ((mx >> shift) >= 4096) shift++;
    std::vector<uint32_t> head(4098, 0);
    auto bucket = [&](uint64_t v) -> uint32_t { return v == ~0ull ? 0u : 4096u - (uint32_t)(v >> shift); };   // 0: unknown length (first)
    for (size_t k = 0; k < n; k++) head[bucket(key[k]) + 1]++;
    for (size_t b = 1; b < head.size(); b++) head[b] += head[b - 1];
    for (size_t k = 0; k < n; k++) order[head[bucket(key[k])]++] = (uint32_t)k;
}

// Block-parallel inflate of a stream that is not made of this library's chunks (inflate_foreign.cuh).  *done = false:
// nothing was decided (too small, too few blocks found, an error, no memory): the caller decodes sequentially, which
// also produces the right error code.  Synchronizes `st` several times (candidate list, chain walk, verdict).
static int inflate_foreign(b200_ctx* c, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, unsigned flags, cudaStream_t st,
                           bool* done, uint64_t* full) {
    *done = false;
    if (n < c->foreign_min || n >= (1ull << 33)) return B200_OK;
    int rc = 0;
    (void)rc;
    // ---- F1: candidate block starts ----
    const uint64_t npieces = (n + FB_PIECE - 1) / FB_PIECE;
    if ((rc = c->f_cand.ensure(npieces * 8))) return B200_OK;
    PROF_BEGIN(c, K_F_FIND, st);
    foreign_find_blocks_kernel<<<(uint32_t)((npieces + FB_WARPS - 1) / FB_WARPS), FB_THREADS, 0, st>>>(
        in, n, npieces, (unsigned long long*)c->f_cand.p);
    LAUNCHED();
    PROF_END(c, st);
    std::vector<uint64_t> piece(npieces);
    CK(cudaMemcpyAsync(piece.data(), c->f_cand.p, npieces * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<uint64_t> starts;
    starts.reserve(npieces + 2);
    starts.push_back(0);                                    // the stream itself starts a block
    for (uint64_t p = 0; p < npieces; p++)
        if (piece[p] != ~0ull && piece[p] != 0) starts.push_back(piece[p]);
    if (starts.size() < 4) return B200_OK;                  // nothing to parallelise over

    // ---- F2: sizes of every candidate unit; chain walk; missing starts are added and counted in further rounds ----
    struct Unit { uint64_t start, stop; FUnitRes r; bool have; };
    std::vector<Unit> units(starts.size());
    for (size_t i = 0; i < starts.size(); i++) units[i] = Unit{starts[i], i + 1 < starts.size() ? starts[i + 1] : ~0ull, FUnitRes{}, false};
    std::vector<size_t> chain;
    for (int round = 0;; round++) {
        // count the units that have no result yet
        std::vector<uint64_t> hs, hp;
        std::vector<size_t> idx;
        for (size_t i = 0; i < units.size(); i++)
            if (!units[i].have) { hs.push_back(units[i].start); hp.push_back(units[i].stop); idx.push_back(i); }
        const uint64_t m = hs.size();
        if (m) {
            if ((rc = c->f_starts.ensure(m * 8)) || (rc = c->f_stops.ensure(m * 8)) || (rc = c->f_res.ensure(m * sizeof(FUnitRes))) ||
                (rc = c->f_order.ensure(m * 4)) || (rc = c->f_misc.ensure(256)))
                return B200_OK;
            // longest first: the compressed span is the estimate (the last unit's is unknown: first)
            std::vector<uint32_t> order;
            {
                std::vector<uint64_t> key(m);
                for (uint64_t k = 0; k < m; k++) key[k] = hp[k] == ~0ull ? ~0ull : hp[k] - hs[k];
                order_longest_first(key, order);
            }
            CK(cudaMemcpyAsync(c->f_starts.p, hs.data(), m * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(c->f_stops.p, hp.data(), m * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(c->f_order.p, order.data(), m * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(c->f_misc.p, 0, 256, st));
            PROF_BEGIN(c, K_F_COUNT, st);
            launch_foreign_decode<false>(c, st, in, n, (const uint64_t*)c->f_starts.p, (const uint64_t*)c->f_stops.p, m, (FUnitRes*)c->f_res.p, nullptr,
                                         nullptr, nullptr, nullptr, flags, (const uint32_t*)c->f_order.p, (unsigned long long*)c->f_misc.p + 16);
            LAUNCHED();
            PROF_END(c, st);
            std::vector<FUnitRes> hr(m);
            CK(cudaMemcpyAsync(hr.data(), c->f_res.p, m * sizeof(FUnitRes), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (uint64_t k = 0; k < m; k++) { units[idx[k]].r = hr[k]; units[idx[k]].have = true; }
        }
        // walk the chain from bit 0
        chain.clear();
        std::vector<uint64_t> missing;
        size_t u = 0;
        bool closed = false;
        for (;;) {
            const FUnitRes& r = units[u].r;
            if (r.status != ST_OK) return B200_OK;          // an error on the true path: the sequential decoder reports it
            chain.push_back(u);
            if (r.flags & FU_FINAL) { closed = true; break; }
            if (!(r.flags & FU_REACHED)) return B200_OK;
            // the next unit must start exactly where this one ended
            auto it = std::lower_bound(units.begin(), units.end(), r.end_bit, [](const Unit& a, uint64_t v) { return a.start < v; });
            if (it != units.end() && it->start == r.end_bit) { u = (size_t)(it - units.begin()); continue; }
            missing.push_back(r.end_bit);
            break;
        }
        if (closed) break;
        if (round >= 6) return B200_OK;
        // every unit (on the chain or not) whose end is nobody's start asks for a new unit: one round fixes them all
        for (size_t i = 0; i < units.size(); i++) {
            const FUnitRes& r = units[i].r;
            if (!units[i].have || r.status != ST_OK || (r.flags & FU_FINAL) || !(r.flags & FU_REACHED)) continue;
            auto it = std::lower_bound(units.begin(), units.end(), r.end_bit, [](const Unit& a, uint64_t v) { return a.start < v; });
            if (it == units.end() || it->start != r.end_bit) missing.push_back(r.end_bit);
        }
        std::sort(missing.begin(), missing.end());
        missing.erase(std::unique(missing.begin(), missing.end()), missing.end());
        std::vector<Unit> merged;
        merged.reserve(units.size() + missing.size());
        size_t a = 0, b = 0;
        while (a < units.size() || b < missing.size()) {
            if (b >= missing.size() || (a < units.size() && units[a].start < missing[b])) merged.push_back(units[a++]);
            else merged.push_back(Unit{missing[b++], 0, FUnitRes{}, false});
        }
        for (size_t i = 0; i < merged.size(); i++)
            if (!merged[i].have) merged[i].stop = i + 1 < merged.size() ? merged[i + 1].start : ~0ull;
        units.swap(merged);
    }

    // ---- layout ----
    const uint64_t nu = chain.size();
    std::vector<uint64_t> hs(nu), hp(nu), hbase(nu + 1), hops(nu + 1);
    uint64_t total = 0, total_ops = 0;
    for (uint64_t k = 0; k < nu; k++) {
        const Unit& U = units[chain[k]];
        hs[k] = U.start; hp[k] = U.r.end_bit;               // decode exactly what was counted
        hbase[k] = total; hops[k] = total_ops;
        total += U.r.out_len; total_ops += U.r.nops;
    }
    hbase[nu] = total; hops[nu] = total_ops;
    *full = total;
    if (total > cap) {
        // truncating caller: the sequential decoder handles it -- unless the caller only wants to learn the size
        // (the vector-returning host overloads probe with a guessed capacity and come back with the exact one)
        if (c->size_probe) *done = true;
        return B200_OK;
    }
    if (total == 0) { *done = true; return B200_OK; }
    if (c->f_sym.ensure(total * 2 + 64) || c->f_ops.ensure((total_ops + 64) * 8)) return B200_OK;     // no memory: sequential
    if ((rc = c->f_starts.ensure(nu * 8)) || (rc = c->f_stops.ensure(nu * 8)) || (rc = c->f_res.ensure(nu * sizeof(FUnitRes))) ||
        (rc = c->f_base.ensure((nu + 1) * 8)) || (rc = c->f_opsbase.ensure((nu + 1) * 8)) || (rc = c->f_misc.ensure(256)))
        return B200_OK;
    CK(cudaMemcpyAsync(c->f_starts.p, hs.data(), nu * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->f_stops.p, hp.data(), nu * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->f_base.p, hbase.data(), (nu + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->f_opsbase.p, hops.data(), (nu + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->f_misc.p, 0, 256, st));
    unsigned long long* d_counter = (unsigned long long*)c->f_misc.p;
    unsigned int* d_err = (unsigned int*)((unsigned long long*)c->f_misc.p + 1);
    uint16_t* S = (uint16_t*)c->f_sym.p;
    const uint64_t* d_base = (const uint64_t*)c->f_base.p;

    // ---- F3: decode with known offsets (longest units first) ----
    {
        std::vector<uint32_t> order;
        {
            std::vector<uint64_t> key(nu);
            for (uint64_t k = 0; k < nu; k++) key[k] = hbase[k + 1] - hbase[k];
            order_longest_first(key, order);
        }
        if ((rc = c->f_order.ensure(nu * 4))) return B200_OK;
        CK(cudaMemcpyAsync(c->f_order.p, order.data(), nu * 4, cudaMemcpyHostToDevice, st));
        PROF_BEGIN(c, K_F_EMIT, st);
        launch_foreign_decode<true>(c, st, in, n, (const uint64_t*)c->f_starts.p, (const uint64_t*)c->f_stops.p, nu, (FUnitRes*)c->f_res.p, d_base,
                                    (const uint64_t*)c->f_opsbase.p, S, (uint64_t*)c->f_ops.p, flags, (const uint32_t*)c->f_order.p, d_counter + 17);
        LAUNCHED();
        PROF_END(c, st);
        CK(cudaStreamSynchronize(st));            // `order` lives on this stack frame
    }
    // ---- F4: ops inside the symbol image (measured: running this on a side stream under the decode of the next part
    // of the units makes both slower -- the decode is bound by the length of its longest unit, not by throughput) ----
    {
        const uint64_t want = (nu + INF_WARPS - 1) / INF_WARPS;
        const uint64_t gmax = (uint64_t)(c->inf_grid / INF_MAX_CTAS_PER_SM) * 12;
        PROF_BEGIN(c, K_F_COPY, st);
        foreign_copy_kernel<<<(uint32_t)(want < gmax ? want : gmax), INF_THREADS, 0, st>>>(
            in, (const FUnitRes*)c->f_res.p, d_base, (const uint64_t*)c->f_opsbase.p, nu, S, (const uint64_t*)c->f_ops.p, d_counter + 8, d_err);
        LAUNCHED();
        PROF_END(c, st);
    }
    // ---- F5: window propagation, two levels ----
    uint32_t G = c->foreign_group;
    if (!G) { G = 1; while ((uint64_t)G * G < nu) G++; if (G < 8) G = 8; }       // ~sqrt(units): group chains and the chain of groups are equally long
    const uint64_t ngroups = (nu + G - 1) / G;
    if ((rc = c->f_gmap.ensure(ngroups * F_WINDOW * 2)) || (rc = c->f_gwin.ensure((ngroups + 1) * F_WINDOW))) return B200_OK;
    PROF_BEGIN(c, K_F_WINDOW, st);
    if (ngroups > 1) {
        foreign_window_kernel<0><<<(uint32_t)ngroups, FW_THREADS, FW_SMEM_BYTES, st>>>(S, d_base, nu, total, G, (uint16_t*)c->f_gmap.p, nullptr, nullptr, d_err);
        LAUNCHED();
        foreign_chain_kernel<<<1, FW_THREADS, FW_SMEM_BYTES, st>>>((const uint16_t*)c->f_gmap.p, (uint8_t*)c->f_gwin.p, ngroups, d_err);
        LAUNCHED();
    }
    foreign_window_kernel<1><<<(uint32_t)ngroups, FW_THREADS, FW_SMEM_BYTES, st>>>(S, d_base, nu, total, G, nullptr, (const uint8_t*)c->f_gwin.p, out, d_err);
    LAUNCHED();
    PROF_END(c, st);
    // ---- F6: the rest ----
    std::vector<FSlab> slabs;
    for (uint64_t k = 0; k < nu; k++) {
        const uint64_t len = hbase[k + 1] - hbase[k];
        if (len <= F_WINDOW) continue;
        for (uint64_t lo = hbase[k]; lo < hbase[k + 1] - F_WINDOW; lo += FR_SLAB) slabs.push_back(FSlab{k, lo});
    }
    if (!slabs.empty()) {
        if ((rc = c->f_slabs.ensure(slabs.size() * sizeof(FSlab)))) return B200_OK;
        CK(cudaMemcpyAsync(c->f_slabs.p, slabs.data(), slabs.size() * sizeof(FSlab), cudaMemcpyHostToDevice, st));
        PROF_BEGIN(c, K_F_RESOLVE, st);
        foreign_resolve_kernel<<<(uint32_t)slabs.size(), FR_THREADS, 0, st>>>(S, d_base, nu, total, (const FSlab*)c->f_slabs.p, out, d_err);
        LAUNCHED();
        PROF_END(c, st);
    }
    unsigned int herr = 1;
    std::vector<FUnitRes> emitted(nu);
    CK(cudaMemcpyAsync(emitted.data(), c->f_res.p, nu * sizeof(FUnitRes), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));            // also keeps hs / hp / hbase / hops / slabs alive until the copies are done
    if (herr) return B200_OK;                 // sequential decoder redoes the stream
    for (uint64_t k = 0; k < nu; k++) {       // the second decode must have produced exactly what the first one counted
        const FUnitRes& a = emitted[k];
        const FUnitRes& b = units[chain[k]].r;
        if (a.status != ST_OK || a.out_len != b.out_len || a.nops != b.nops || a.end_bit != b.end_bit) return B200_OK;
    }
    *done = true;
    return B200_OK;
}

int b200_adler32_dev(b200_ctx* c, const void* d_data, size_t n, uint32_t* h_out, uint32_t* d_out, void* stream_) {
    if (!c || (!d_data && n) || (!h_out && !d_out)) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    int rc;
    const uint64_t nblocks = (n + CHUNK - 1) / CHUNK;
    if ((rc = c->adler_parts.ensure((nblocks + 1) * sizeof(AdlerPart) + 16))) return rc;
    AdlerPart* parts = (AdlerPart*)c->adler_parts.p;
    uint32_t* d_res = d_out ? d_out : (uint32_t*)(parts + nblocks + 1);
    if (nblocks) {
        adler_partial_kernel<<<(uint32_t)nblocks, ADLER_THREADS, 0, st>>>((const uint8_t*)d_data, n, parts);
        LAUNCHED();
    }
    adler_fold_kernel<<<1, 32, 0, st>>>(parts, nblocks, n, d_res);
    LAUNCHED();
    if (h_out) {
        CK(cudaMemcpyAsync(h_out, d_res, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return B200_OK;
}

int b200_crc32_dev(b200_ctx* c, const void* d_data, size_t n, uint32_t* h_out, uint32_t* d_out, void* stream_) {
    if (!c || (!d_data && n) || (!h_out && !d_out)) return B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    ON_DEVICE(c);
    int rc;
    if ((rc = c->total.ensure(64))) return rc;
    uint32_t* d_res = d_out ? d_out : (uint32_t*)((uint64_t*)c->total.p + 6);
    if ((rc = checksum_dev(c, (const uint8_t*)d_data, n, 2, d_res, st))) return rc;
    if (h_out) {
        CK(cudaMemcpyAsync(h_out, d_res, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return B200_OK;
}

int b200_publish_dev(b200_ctx* c, void* h_pinned_dst, const void* d_src, size_t n_words, void* stream_) {
    if (!c || !h_pinned_dst || !d_src || n_words == 0 || n_words > 32) return B200_E_ARG;
    ON_DEVICE(c);
    // the address the DEVICE uses for this host memory (the same one under unified addressing for cudaHostAlloc memory; registered
    // memory may differ); pageable memory is refused
    void* dptr = nullptr;
    if (cudaHostGetDevicePointer(&dptr, h_pinned_dst, 0) != cudaSuccess || !dptr) { cudaGetLastError(); return B200_E_ARG; }
    publish_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>((volatile unsigned long long*)dptr, (const unsigned long long*)d_src,
                                                       (uint32_t)n_words, nullptr, 0);
    LAUNCHED();
    CK(cudaGetLastError());
    return B200_OK;
}

int b200_corpus_generate_dev(void* d_out, uint64_t seed, uint64_t first_chunk, uint64_t n_chunks, void* stream_) {
    if (!d_out && n_chunks) return B200_E_ARG;
    if (!n_chunks) return B200_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    corpus_text_kernel<<<(uint32_t)((n_chunks + 63) / 64), 64, 0, st>>>((uint8_t*)d_out, seed, first_chunk, n_chunks);
    LAUNCHED();
    const uint64_t words = n_chunks * (CHUNK / 8);
    corpus_flat_kernel<<<(uint32_t)((words + 255) / 256), 256, 0, st>>>((uint8_t*)d_out, seed, first_chunk, n_chunks);
    LAUNCHED();
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------
// Pageable caller memory.  cudaMemcpyAsync on pageable memory is a synchronous, single-threaded bounce through the
// driver's own staging buffer (6-12 GB/s) -- what a user of deflate::compress(char*, n, level) hands us is exactly
// that.  Large pageable buffers therefore travel through a ring of pinned 8 MiB buffers of our own: a few host threads
// copy user memory <-> ring in parallel, the DMA engines move ring <-> HBM asynchronously behind them.
static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
static int stage_threads(const b200_ctx* c) {
    if (c->stg_threads > 0) return c->stg_threads;
    const unsigned hw = std::thread::hardware_concurrency();
    return hw >= 16 ? 8 : hw >= 8 ? 4 : hw >= 4 ? 2 : 1;
}
static void par_memcpy(void* dst, const void* src, size_t len, int threads) {
    if (threads <= 1 || len < ((size_t)1 << 20)) { memcpy(dst, src, len); return; }
    std::vector<std::thread> pool;
    const size_t per = ((len + threads - 1) / threads + 4095) & ~(size_t)4095;
    for (int t = 1; t < threads; t++) {
        const size_t lo = (size_t)t * per;
        if (lo >= len) break;
        const size_t l = len - lo < per ? len - lo : per;
        pool.emplace_back([=]() { memcpy((uint8_t*)dst + lo, (const uint8_t*)src + lo, l); });
    }
    memcpy(dst, src, len < per ? len : per);
    for (auto& th : pool) th.join();
}
static int stage_init(b200_ctx* c) {
    for (int k = 0; k < b200_ctx::STG_N; k++) {
        if (!c->stg_in[k] && cudaHostAlloc(&c->stg_in[k], b200_ctx::STG_BYTES, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return B200_E_NOMEM; }
        if (!c->stg_out[k] && cudaHostAlloc(&c->stg_out[k], b200_ctx::STG_BYTES, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return B200_E_NOMEM; }
        if (!c->stg_in_ev[k]) CK(cudaEventCreateWithFlags(&c->stg_in_ev[k], cudaEventDisableTiming));
        if (!c->stg_out_ev[k]) CK(cudaEventCreateWithFlags(&c->stg_out_ev[k], cudaEventDisableTiming));
    }
    return B200_OK;
}
// host -> device.  Pinned source: one asynchronous copy.  Pageable source: staged; returns when the last piece has
// been HANDED to the DMA engine (the copies themselves complete in stream order, like any cudaMemcpyAsync).
static int copy_h2d(b200_ctx* c, void* d_dst, const void* h_src, size_t len, cudaStream_t st) {
    if (!len) return B200_OK;
    if (!c->stage_pageable || len < ((size_t)4 << 20) || !is_pageable(h_src)) {
        CK(cudaMemcpyAsync(d_dst, h_src, len, cudaMemcpyHostToDevice, st));
        return B200_OK;
    }
    int rc = stage_init(c);
    if (rc) return rc;
    const int T = stage_threads(c);
    size_t off = 0;
    for (int i = 0; off < len; i++) {
        const int slot = i % b200_ctx::STG_N;
        const size_t l = len - off < b200_ctx::STG_BYTES ? len - off : b200_ctx::STG_BYTES;
        if (c->stg_in_busy[slot]) CK(cudaEventSynchronize(c->stg_in_ev[slot]));       // the DMA that last read this slot is done
        par_memcpy(c->stg_in[slot], (const uint8_t*)h_src + off, l, T);
        CK(cudaMemcpyAsync((uint8_t*)d_dst + off, c->stg_in[slot], l, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(c->stg_in_ev[slot], st));
        c->stg_in_busy[slot] = true;
        off += l;
    }
    return B200_OK;
}
// device -> host.  Pinned destination: one asynchronous copy (*sync = false: the caller synchronises the stream).
// Pageable destination: staged and COMPLETE on return.
static int copy_d2h(b200_ctx* c, void* h_dst, const void* d_src, size_t len, cudaStream_t st) {
    if (!len) return B200_OK;
    if (!c->stage_pageable || len < ((size_t)4 << 20) || !is_pageable(h_dst)) {
        CK(cudaMemcpyAsync(h_dst, d_src, len, cudaMemcpyDeviceToHost, st));
        return B200_OK;
    }
    int rc = stage_init(c);
    if (rc) return rc;
    const int T = stage_threads(c);
    const size_t P = b200_ctx::STG_BYTES;
    const size_t npieces = (len + P - 1) / P;
    const int N = b200_ctx::STG_N;
    for (size_t i = 0; i < npieces + (size_t)(N - 1); i++) {
        if (i < npieces) {                                   // keep N - 1 device -> ring copies in flight
            const size_t off = i * P, l = len - off < P ? len - off : P;
            CK(cudaMemcpyAsync(c->stg_out[i % N], (const uint8_t*)d_src + off, l, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(c->stg_out_ev[i % N], st));
        }
        if (i >= (size_t)(N - 1)) {
            const size_t j = i - (N - 1), off = j * P, l = len - off < P ? len - off : P;
            CK(cudaEventSynchronize(c->stg_out_ev[j % N]));
            par_memcpy((uint8_t*)h_dst + off, c->stg_out[j % N], l, T);
        }
    }
    return B200_OK;
}

static int arena_ensure(b200_ctx* c, size_t need) {
    if (need <= c->arena_cap) return B200_OK;
    if (c->arena) cudaFreeHost(c->arena);
    c->arena = nullptr; c->arena_cap = 0;
    const size_t want = need + need / 8 + 4096;
    if (cudaHostAlloc(&c->arena, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return B200_E_NOMEM; }
    c->arena_cap = want;
    return B200_OK;
}

static int inflate_host(const uint8_t* in, size_t n, void* out, size_t cap, void** out_alloc, size_t* out_n,
                        size_t* full_n, unsigned flags, const uint8_t* adler_expect = nullptr, int tkind = 1, bool to_arena = false);

// ------------------------------------------------------------------------------------------------
// Host-buffer API
//
// Compress host memory into host memory.  The input is cut into slices (host_slice_chunks, 64 MiB by
// default); slice k+1's host->device copy, slice k's kernels and slice k-1's device->host copy run on
// three streams, chained by events, so with pinned host buffers the call is bound by the slower PCIe
// direction instead of the sum of copy + compute + copy.  Output offsets chain on the device (K3's
// carry), the host only learns each slice's end offset through a pinned mailbox.
static int compress_host(b200_ctx* c, const uint8_t* in, size_t n, int level, unsigned flags, uint8_t* out, size_t cap, size_t* out_n) {
    ON_DEVICE(c);
    int rc;
    const int frame = frame_mode(flags);
    IndexGuard guard(c, flags);
    const size_t bound = b200_deflate_bound(n);
    if ((rc = c->d_in.ensure(n + 64))) return rc;
    if ((rc = c->d_out.ensure(bound + 64))) return rc;
    const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
    if (nchunks == 0) {
        size_t cn = 0;
        rc = b200_deflate_compress_dev(c, c->d_in.p, 0, level, flags & (B200_F_ZLIB | B200_F_GZIP), c->d_out.p, c->d_out.cap, nullptr, &cn, nullptr, c->stream);
        if (rc) return rc;
        *out_n = cn;
        if (cn > cap) return B200_E_OUTPUT;
        CK(cudaMemcpyAsync(out, c->d_out.p, cn, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return B200_OK;
    }
    const uint64_t B = nchunks < c->host_slice_chunks ? nchunks : c->host_slice_chunks;
    // slice k = chunks [cuts[k], cuts[k + 1]).  What the call cannot hide behind the host -> device copies is the last slice's
    // kernels and its way back, so a long input ends with slices of B/2, B/4, B/4 chunks (64 MiB slices in between: smaller
    // ones starve the persistent matcher)
    std::vector<uint64_t> cuts;
    {
        const uint64_t tail = (nchunks >= 4 * B && B >= 4) ? B : 0;
        uint64_t pos = 0;
        while (pos < nchunks - tail) { cuts.push_back(pos); pos += (nchunks - tail - pos < B) ? nchunks - tail - pos : B; }
        if (tail) { cuts.push_back(pos); pos += B / 2; cuts.push_back(pos); pos += B / 4; cuts.push_back(pos); }
        cuts.push_back(nchunks);
    }
    const uint64_t nslices = cuts.size() - 1;
    if ((rc = c->tok.ensure(B * CHUNK * 4))) return rc;
    if ((rc = c->ntok.ensure(B * NSEG * 4))) return rc;
    if ((rc = c->hist.ensure(B * NSEG * NSYM * 2))) return rc;
    if ((rc = c->codes.ensure(B * 2 * NSYM * 4))) return rc;
    if ((rc = c->hdr.ensure(B * 2 * HDR_WORDS * 4))) return rc;
    if ((rc = c->desc.ensure(B * sizeof(BlockDesc)))) return rc;
    if ((rc = c->sizes.ensure(B * 4))) return rc;
    if ((rc = c->offsets.ensure((nchunks + 1) * 8))) return rc;
    if ((rc = c->total.ensure(64))) return rc;
    uint64_t* offs = (uint64_t*)c->offsets.p;
    uint64_t* d_total = (uint64_t*)c->total.p;
    const uint64_t* d_first_base = nullptr;
    if (frame) {
        // the header goes straight into the caller's buffer; on the device the stream starts behind a gap of the same size
        if (cap < kFrameLen[frame]) return B200_E_OUTPUT;
        memcpy(out, frame == 1 ? kZlibHeader : kGzipHeader, kFrameLen[frame]);
        CK(cudaMemcpyAsync(d_total + 2, &kFrameLen[frame], 8, cudaMemcpyHostToDevice, c->stream));
        d_first_base = d_total + 2;
    }
    if (c->mailbox_cap < nslices + 1) {
        if (c->mailbox) cudaFreeHost(c->mailbox);
        c->mailbox = nullptr; c->mailbox_cap = 0;
        CK(cudaHostAlloc((void**)&c->mailbox, (nslices + 16) * 8, cudaHostAllocDefault));
        c->mailbox_cap = nslices + 16;
    }
    while (c->events.size() < 2 * nslices) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->events.push_back(e);
    }
    uint8_t* d_in = (uint8_t*)c->d_in.p;
    uint8_t* d_out = (uint8_t*)c->d_out.p;
    uint64_t begin = kFrameLen[frame], drained = 0;
    int result = B200_OK;
    // enqueue every slice: H2D on s_in, kernels on stream (after the slice's H2D), mailbox write
    for (uint64_t k = 0; k < nslices; k++) {
        const uint64_t b0 = cuts[k];
        const uint32_t nb = (uint32_t)(cuts[k + 1] - b0);
        const size_t off = (size_t)b0 * CHUNK;
        const size_t len = (size_t)((b0 + nb == nchunks) ? n - off : (size_t)nb * CHUNK);
        if ((rc = copy_h2d(c, d_in + off, in + off, len, c->s_in))) return rc;          // pageable input: the host stages slice k
        CK(cudaEventRecord(c->events[2 * k], c->s_in));                                  // while the GPU works on slice k - 1
        CK(cudaStreamWaitEvent(c->stream, c->events[2 * k], 0));
        if ((rc = compress_batch(c, d_in + off, n - off, nb, b0, b0 + nb == nchunks, level, offs, d_total, d_out, c->stream, 3, nullptr,
                                 nullptr, d_first_base)))
            return rc;
        // the slice's end offset goes to the mailbox by a kernel, not by a copy that would queue behind the bulk copies
        publish_kernel<<<1, 32, 0, c->stream>>>((volatile unsigned long long*)&c->mailbox[k], (const unsigned long long*)(offs + b0 + nb), 1, nullptr, 0);
        LAUNCHED();
        CK(cudaEventRecord(c->events[2 * k + 1], c->stream));
        // slices that are finished already start their way back now instead of after the last slice was enqueued
        while (drained + 1 < k && cudaEventQuery(c->events[2 * drained + 1]) == cudaSuccess) {
            const uint64_t endo = c->mailbox[drained];
            if (endo > cap) result = B200_E_OUTPUT;
            if (result == B200_OK && endo > begin && (rc = copy_d2h(c, out + begin, d_out + begin, endo - begin, c->s_out))) return rc;
            begin = endo;
            drained++;
        }
    }
    if (frame) {
        // trailer: checksum of the INPUT, which is complete on the device once the last slice's kernels have run
        if ((rc = checksum_dev(c, d_in, n, frame, (uint32_t*)(d_total + 4), c->stream))) return rc;
        write_trailer_kernel<<<1, 1, 0, c->stream>>>(d_out, d_total, (const uint32_t*)(d_total + 4), (uint64_t)n, frame);
        LAUNCHED();
        publish_kernel<<<1, 32, 0, c->stream>>>((volatile unsigned long long*)&c->mailbox[nslices], (const unsigned long long*)d_total, 1, nullptr, 0);
        LAUNCHED();
    }
    // drain: as each slice finishes, copy its compressed bytes out on s_out
    for (uint64_t k = drained; k < nslices; k++) {
        CK(cudaEventSynchronize(c->events[2 * k + 1]));
        const uint64_t endo = c->mailbox[k];
        if (endo > cap) result = B200_E_OUTPUT;
        if (result == B200_OK && endo > begin && (rc = copy_d2h(c, out + begin, d_out + begin, endo - begin, c->s_out))) return rc;
        begin = endo;
    }
    CK(cudaStreamSynchronize(c->stream));
    if (frame) {
        const uint64_t endo = c->mailbox[nslices];
        if (endo > cap) result = B200_E_OUTPUT;
        if (result == B200_OK && endo > begin) CK(cudaMemcpyAsync(out + begin, d_out + begin, endo - begin, cudaMemcpyDeviceToHost, c->s_out));
        begin = endo;
    }
    CK(cudaStreamSynchronize(c->s_out));
    *out_n = (size_t)begin;
    return result;
}

int b200_deflate_compress_into_ex(const void* in, size_t n, int level, unsigned flags, void* out, size_t cap, size_t* out_n) {
    if ((!in && n) || !out_n || level < 0 || level > 3 || (!out && cap) || (flags & B200_F_NOT_LAST)) return B200_E_ARG;
    b200_ctx* c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    return compress_host(c, (const uint8_t*)in, n, level, flags, (uint8_t*)out, cap, out_n);
}

int b200_deflate_compress_into(const void* in, size_t n, int level, void* out, size_t cap, size_t* out_n) {
    return b200_deflate_compress_into_ex(in, n, level, 0, out, cap, out_n);
}

// Views: the result stays in the library's pinned arena (the DMA engine writes it there at full PCIe speed, no page
// faults) and the caller copies it once into wherever it must live -- the drop-in header builds its std::vector with
// one assign().  The default context is locked from a successful *_view call until b200_view_release().
int b200_deflate_compress_view(const void* in, size_t n, int level, unsigned flags, const void** view, size_t* out_n) {
    if (!view || !out_n || (!in && n) || level < 0 || level > 3 || (flags & B200_F_NOT_LAST)) return B200_E_ARG;
    *view = nullptr; *out_n = 0;
    b200_ctx* c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::unique_lock<std::mutex> lk(c->mu);
    {
        ON_DEVICE(c);
        if ((rc = arena_ensure(c, b200_deflate_bound(n) + 64))) return rc;
    }
    size_t cn = 0;
    rc = compress_host(c, (const uint8_t*)in, n, level, flags, (uint8_t*)c->arena, c->arena_cap, &cn);
    if (rc) return rc;
    *view = c->arena; *out_n = cn;
    c->view_locked = true;
    lk.release();
    return B200_OK;
}

int b200_inflate_view(const void* in, size_t n, unsigned flags, int framing, const void** view, size_t* out_n) {
    if (!view || !out_n || (!in && n) || framing < 0 || framing > 2) return B200_E_ARG;
    *view = nullptr; *out_n = 0;
    const uint8_t* p = (const uint8_t*)in;
    void* v = nullptr;
    int rc;
    if (framing == 1) {
        if (n < 2) return B200_E_OVERRUN;
        const size_t s = p[1] & 0x20 ? 6 : 2;
        if (s > n) return B200_E_OVERRUN;
        if (flags & B200_F_STRICT) {
            if (n < s + 4) return B200_E_OVERRUN;
            if ((p[0] & 15) != 8 || (((uint32_t)p[0] << 8) | p[1]) % 31 != 0) return B200_E_DATA;
            rc = inflate_host(p + s, n - s - 4, nullptr, 0, &v, out_n, nullptr, flags, p + n - 4, 1, true);
        } else {
            rc = inflate_host(p + s, n - s, nullptr, 0, &v, out_n, nullptr, flags, nullptr, 1, true);
        }
    } else {
        rc = inflate_host(p, n, nullptr, 0, &v, out_n, nullptr, flags, nullptr, 1, true);
    }
    if (rc) { *out_n = 0; return rc; }
    *view = v;
    return B200_OK;
}

// Touch every page of [p, p + n) from a few threads.  The std::vector a drop-in call returns is fresh memory: its first
// write faults in 4 KiB at a time, single-threaded, at ~3 GB/s -- slower than everything the GPU did.  The header reserves
// the vector, calls this (the faults, i.e. the kernel's page zeroing, then run on several cores at once) and only then
// assigns.  Transparent huge pages are asked for where the system offers them on request.  Writes one zero byte per page:
// only for memory whose contents do not matter yet.
void b200_host_prefault(void* p, size_t n) {
    if (!p || n < ((size_t)4 << 20)) return;
    uint8_t* b = (uint8_t*)p;
    const unsigned hw = std::thread::hardware_concurrency();
    const int T = hw >= 16 ? 8 : hw >= 8 ? 4 : hw >= 4 ? 2 : 1;
    const size_t page = 4096;
#ifdef MADV_HUGEPAGE
    {
        const uintptr_t lo = ((uintptr_t)b + (((size_t)2 << 20) - 1)) & ~(uintptr_t)(((size_t)2 << 20) - 1);
        const uintptr_t hi = ((uintptr_t)b + n) & ~(uintptr_t)(((size_t)2 << 20) - 1);
        if (hi > lo) madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
    }
#endif
    std::vector<std::thread> pool;
    const size_t per = ((n + T - 1) / T + page - 1) & ~(page - 1);
    auto touch = [b, n, page](size_t lo, size_t hi) {
        if (hi > n) hi = n;
        for (size_t o = lo; o < hi; o += page) ((volatile uint8_t*)b)[o] = 0;
    };
    for (int t = 1; t < T; t++) pool.emplace_back(touch, (size_t)t * per, (size_t)(t + 1) * per);
    touch(0, per);
    for (auto& th : pool) th.join();
    ((volatile uint8_t*)b)[n - 1] = 0;
}

void b200_view_release(void) {
    std::lock_guard<std::mutex> g(g_default_mu);
    if (g_default && g_default->view_locked) {
        g_default->view_locked = false;
        g_default->mu.unlock();
    }
}

int b200_deflate_compress(const void* in, size_t n, int level, void** out, size_t* out_n) {
    return b200_deflate_compress_ex(in, n, level, 0, out, out_n);
}

int b200_deflate_compress_ex(const void* in, size_t n, int level, unsigned flags, void** out, size_t* out_n) {
    if (!out || !out_n || (!in && n) || level < 0 || level > 3 || (flags & B200_F_NOT_LAST)) return B200_E_ARG;
    *out = nullptr; *out_n = 0;
    b200_ctx* c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    const size_t bound = b200_deflate_bound(n);
    uint8_t* buf = (uint8_t*)malloc(bound);          // untouched pages cost nothing; shrunk below
    if (!buf) return B200_E_NOMEM;
    size_t cn = 0;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        rc = compress_host(c, (const uint8_t*)in, n, level, flags, buf, bound, &cn);
    }
    if (rc) { free(buf); return rc; }
    void* shrunk = realloc(buf, cn ? cn : 1);
    *out = shrunk ? shrunk : buf;
    *out_n = cn;
    return B200_OK;
}

// Host buffers, long streams of this library's own format: the input travels in slices; as soon as a slice has
// landed, the chunks that are complete so far are found (find_sync from the last known chunk start), decoded and
// validated as a group, and their output starts its way back on a third stream -- H2D of slice k + 1, kernels of
// slice k and D2H of slice k - 1 overlap (PCIe is full duplex), instead of H2D, then kernels, then D2H.
// *handled = false: the stream did not validate (foreign, damaged, ...): nothing is reported, the caller takes the
// plain path, which also produces the proper error code.  The decoded bytes stay in c->d_out.
static int inflate_host_pipelined(b200_ctx* c, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n,
                                  size_t* full_n, unsigned flags, bool* handled, bool* in_resident) {
    *handled = false;
    *in_resident = false;
    const size_t S = c->host_inflate_slice & ~(size_t)15;
    const size_t first = (S / 8 > 65536 ? S / 8 : 65536) & ~(size_t)15;         // 16 MiB with the default slice of 128 MiB
    if (!S || n < 3 * first || !cap || c->inflate_warp_path) return B200_OK;
    int rc;
    // Slice ends.  The device -> host copy of the output is the long pole (1.6x the bytes of the input), so what counts is how
    // early it can start: the first slices are small (32, 64, 128 MiB ...: the first output is on its way back after ~1.5 ms
    // instead of ~7), the later ones have the full size (small groups of chunks run as partial waves of the decoder).
    std::vector<size_t> ends;
    {
        size_t pos = 0, step = first;
        while (pos < n) {
            const size_t cur = step < S ? step : S;
            pos = n - pos <= cur + cur / 2 ? n : pos + cur;      // no tiny last slice
            ends.push_back(pos);
            step <<= 1;
        }
    }
    const size_t nsl = ends.size();
    if ((rc = c->d_in.ensure(n + 64))) return rc;
    if ((rc = c->d_out.ensure(cap + 64))) return rc;
    if ((rc = c->result.ensure(64))) return rc;
    while (c->events.size() < nsl) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->events.push_back(e);
    }
    uint8_t* d_in = (uint8_t*)c->d_in.p;
    uint8_t* d_out = (uint8_t*)c->d_out.p;
    unsigned long long* d_result = (unsigned long long*)c->result.p;
    cudaStream_t st = c->stream;
    if (!c->ctl) CK(cudaHostAlloc((void**)&c->ctl, 16 * 8, cudaHostAllocDefault));
    volatile unsigned long long* ctl = c->ctl;
    // the input travels on its own host thread: with pageable caller memory every slice is staged through the pinned
    // ring by host threads (copy_h2d), which must not hold up the decode loop below
    std::atomic<size_t> arrived{0};
    std::atomic<int> feed_rc{B200_OK};
    std::thread feeder([&]() {
        cudaSetDevice(c->device);
        for (size_t k = 0; k < nsl; k++) {
            const size_t off = k ? ends[k - 1] : 0, len = ends[k] - off;
            int r = copy_h2d(c, d_in + off, in + off, len, c->s_in);
            if (r == B200_OK && cudaEventRecord(c->events[k], c->s_in) != cudaSuccess) r = B200_E_CUDA;
            if (r) { feed_rc.store(r); arrived.store(nsl, std::memory_order_release); return; }
            arrived.store(k + 1, std::memory_order_release);
        }
    });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{feeder};
    uint64_t start = 0, chunk0 = 0, total = 0;
    bool ok = true;
    for (size_t k = 0; k < nsl && ok; k++) {
        while (arrived.load(std::memory_order_acquire) <= k) std::this_thread::yield();     // events[k] has been recorded
        if (feed_rc.load()) return feed_rc.load();
        CK(cudaStreamWaitEvent(st, c->events[k], 0));
        const bool last = k + 1 == nsl;
        const uint64_t avail = ends[k];
        const uint64_t base = start & ~15ull, rel0 = start - base, region = avail - base;
        const uint8_t* rin = d_in + base;
        const uint64_t nwarps = (region + SYNC_REGION - 1) / SYNC_REGION;
        const uint64_t cand_cap = region / 64 + 1024;
        if ((rc = c->counts.ensure(nwarps * 4))) return rc;
        if ((rc = c->woffs.ensure((nwarps + 1) * 8))) return rc;
        if ((rc = c->sync_cache.ensure(nwarps * SYNC_CACHE * 4))) return rc;
        if ((rc = c->cand.ensure((cand_cap + 2) * 8))) return rc;
        uint64_t* cand = (uint64_t*)c->cand.p;
        const uint32_t g = (uint32_t)((nwarps * 32 + 255) / 256);
        PROF_BEGIN(c, K_FIND_SYNC, st);
        find_sync_kernel<false><<<g, 256, 0, st>>>(rin, region, (uint32_t*)c->counts.p, (uint32_t*)c->sync_cache.p, nullptr, nullptr, 0, rel0);
        LAUNCHED();
        PROF_END(c, st);
        scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)c->counts.p, (uint32_t)nwarps, nullptr, (uint64_t*)c->woffs.p,
                                                     (uint64_t*)d_result + 2);
        LAUNCHED();
        // (control words travel through kernels, not through the copy engines the slices are on: see publish_kernel)
        publish_kernel<<<1, 32, 0, st>>>(ctl, d_result + 2, 1, nullptr, 0);
        LAUNCHED();
        CK(cudaStreamSynchronize(st));
        const uint64_t nmark = ctl[0];
        if (nmark + 2 > cand_cap) { ok = false; break; }
        const uint64_t ncand = nmark + 1;
        const uint64_t units = last ? ncand : ncand - 1;      // the last candidate of a slice starts a chunk that is still arriving
        if (units == 0) continue;
        // cand[0] = where this slice's first chunk starts, the sentinel behind the last slice's candidates (the last chunk
        // ends at n), and the verdict words {valid = 1, total = 0}
        poke_kernel<<<1, 1, 0, st>>>((unsigned long long*)cand, rel0, last ? (unsigned long long*)cand + ncand : nullptr, region,
                                     d_result, 1ull, d_result + 1, 0ull);
        LAUNCHED();
        if (nmark) {
            PROF_BEGIN(c, K_FIND_SYNC, st);
            find_sync_kernel<true><<<g, 256, 0, st>>>(rin, region, (uint32_t*)c->counts.p, (uint32_t*)c->sync_cache.p,
                                                      (const uint64_t*)c->woffs.p, cand + 1, nmark, rel0);
            LAUNCHED();
            PROF_END(c, st);
        }
        const uint64_t o0 = chunk0 * CHUNK;
        if ((rc = inflate_chunks_two_pass(c, rin, region, cand, units, d_out + o0, cap > o0 ? cap - o0 : 0, flags, st))) return rc;
        PROF_BEGIN(c, K_VALIDATE, st);
        validate_units_kernel<<<(uint32_t)((units + 255) / 256), 256, 0, st>>>(cand, units, region, (const TpResult*)c->tpres.p, d_result,
                                                                              last ? 1 : 0);
        LAUNCHED();
        PROF_END(c, st);
        publish_kernel<<<1, 32, 0, st>>>(ctl + 4, d_result, 2, (const unsigned long long*)cand + units, 1);
        LAUNCHED();
        CK(cudaStreamSynchronize(st));
        const unsigned long long verdict[2] = {ctl[4], ctl[5]};
        const uint64_t next_rel = ctl[6];
        if (verdict[0] != 1) { ok = false; break; }
        const uint64_t lo = o0 < cap ? o0 : cap, hi = o0 + verdict[1] < cap ? o0 + verdict[1] : cap;
        if (hi > lo && (rc = copy_d2h(c, out + lo, d_out + lo, hi - lo, c->s_out))) return rc;
        total = o0 + verdict[1];
        chunk0 += units;
        start = base + next_rel;
    }
    feeder.join();
    CK(cudaStreamSynchronize(c->s_in));
    CK(cudaStreamSynchronize(c->s_out));
    if (feed_rc.load()) return feed_rc.load();
    *in_resident = true;                       // whatever happens next, the whole input sits in c->d_in
    if (!ok) return B200_OK;
    *handled = true;
    if (out_n) *out_n = (size_t)(total < cap ? total : cap);
    if (full_n) *full_n = (size_t)total;
    return B200_OK;
}

// adler_expect: if non-NULL, the trailer the decoded bytes must match (strict mode): tkind 1 = 4-byte big-endian
// Adler-32 (zlib), tkind 2 = CRC-32 + ISIZE, little endian (gzip)
// to_arena: *out_alloc receives a pointer into the context's pinned arena instead of a malloc'ed buffer, and the
// context stays locked until b200_view_release()
static int inflate_host(const uint8_t* in, size_t n, void* out, size_t cap, void** out_alloc, size_t* out_n,
                        size_t* full_n, unsigned flags, const uint8_t* adler_expect, int tkind, bool to_arena) {
    b200_ctx* c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::unique_lock<std::mutex> lk(c->mu);
    ON_DEVICE(c);
    size_t written = 0, full = 0;
    size_t dcap = cap;
    bool piped = false, in_resident = false;
    if (!out_alloc) {
        if ((rc = inflate_host_pipelined(c, in, n, (uint8_t*)out, cap, &written, &full, flags, &piped, &in_resident))) return rc;
        // strict zlib: the trailer covers the WHOLE decoded stream; a truncating caller buffer does not excuse it
        if (piped && adler_expect && full > cap) piped = false;
    }
    if (!piped) {
    if ((rc = c->d_in.ensure(n + 64))) return rc;
    if (n && !in_resident && (rc = copy_h2d(c, c->d_in.p, in, n, c->stream))) return rc;
    const bool need_full = out_alloc || adler_expect;      // the device buffer must hold every decoded byte
    if (out_alloc) dcap = n * 4 + 65536 > (size_t)1 << 20 ? n * 4 + 65536 : (size_t)1 << 20;   // first guess
    for (int attempt = 0; attempt < 3; attempt++) {
        if ((rc = c->d_out.ensure(dcap + 64))) return rc;
        c->size_probe = need_full && attempt == 0;
        rc = b200_inflate_dev(c, c->d_in.p, n, c->d_out.p, dcap, nullptr, &written, &full, nullptr, flags, c->stream);
        c->size_probe = false;
        if (!need_full || full <= dcap) break;
        dcap = full;                       // decoded size is now known exactly: one more pass
    }
    if (!out_alloc && written > cap) written = cap;        // the caller's buffer truncates (inflate.hpp:345)
    }   // !piped
    if (rc == B200_OK && adler_expect && full <= dcap) {
        // the whole decoded stream sits in d_out: check it there
        uint32_t got = 0;
        if (tkind == 1) {
            const int rc2 = b200_adler32_dev(c, c->d_out.p, full, &got, nullptr, c->stream);
            if (rc2) return rc2;
            const uint32_t want = ((uint32_t)adler_expect[0] << 24) | ((uint32_t)adler_expect[1] << 16) |
                                  ((uint32_t)adler_expect[2] << 8) | adler_expect[3];
            if (got != want) rc = B200_E_DATA;
        } else {
            const int rc2 = b200_crc32_dev(c, c->d_out.p, full, &got, nullptr, c->stream);
            if (rc2) return rc2;
            const uint32_t want = (uint32_t)adler_expect[0] | ((uint32_t)adler_expect[1] << 8) | ((uint32_t)adler_expect[2] << 16) |
                                  ((uint32_t)adler_expect[3] << 24);
            const uint32_t isize = (uint32_t)adler_expect[4] | ((uint32_t)adler_expect[5] << 8) | ((uint32_t)adler_expect[6] << 16) |
                                   ((uint32_t)adler_expect[7] << 24);
            if (got != want || isize != (uint32_t)full) rc = B200_E_DATA;
        }
    }
    if (out_alloc && to_arena) {
        if (rc == B200_OK || written) {
            int rc2 = arena_ensure(c, written + 64);
            if (rc2) return rc2;
            if (written) CK(cudaMemcpyAsync(c->arena, c->d_out.p, written, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        *out_alloc = c->arena;
        if (rc == B200_OK) { c->view_locked = true; lk.release(); }      // b200_view_release() unlocks
    } else if (out_alloc) {
        void* buf = malloc(written ? written : 1);
        if (!buf) return B200_E_NOMEM;
        if (written && cudaMemcpyAsync(buf, c->d_out.p, written, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { free(buf); return B200_E_CUDA; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { free(buf); return B200_E_CUDA; }
        *out_alloc = buf;
    } else if (written && !piped) {
        const int rc2 = copy_d2h(c, out, c->d_out.p, written, c->stream);      // (rc keeps the decoder's verdict)
        if (rc2) return rc2;
        CK(cudaStreamSynchronize(c->stream));
    }
    if (out_n) *out_n = written;
    if (full_n) *full_n = full;
    return rc;
}

int b200_inflate(const void* in, size_t n, void* out, size_t cap, size_t* out_n, size_t* full_n, unsigned flags) {
    if ((!in && n) || (!out && cap)) return B200_E_ARG;
    return inflate_host((const uint8_t*)in, n, out, cap, nullptr, out_n, full_n, flags);
}

int b200_inflate_alloc(const void* in, size_t n, void** out, size_t* out_n, unsigned flags) {
    if ((!in && n) || !out || !out_n) return B200_E_ARG;
    *out = nullptr; *out_n = 0;
    return inflate_host((const uint8_t*)in, n, nullptr, 0, out, out_n, nullptr, flags);
}

static size_t zlib_header_skip(const uint8_t* in, size_t n) {
    // RFC 1950: CMF, FLG; FLG bit 5 = FDICT -> 4 more bytes (DICTID).  (The reference always skips 2:
    // its FDICT test reads bit 26 of a byte value, inflate.hpp:329,355.)
    if (n < 2) return n;
    size_t skip = 2;
    if (in[1] & 0x20) skip = 6;
    return skip < n ? skip : n;
}

// strict mode: CM must be 8 (deflate), the header check bits must hold, and there must be room for the trailer
static int zlib_strict_check(const uint8_t* in, size_t n, size_t skip) {
    if (n < skip + 4) return B200_E_OVERRUN;
    if ((in[0] & 15) != 8 || (((uint32_t)in[0] << 8) | in[1]) % 31 != 0) return B200_E_DATA;
    return B200_OK;
}

int b200_inflate_zlib(const void* in, size_t n, void* out, size_t cap, size_t* out_n, size_t* full_n, unsigned flags) {
    if (!in || n < 2) return B200_E_OVERRUN;
    if ((!out && cap)) return B200_E_ARG;
    const uint8_t* p = (const uint8_t*)in;
    const size_t s = zlib_header_skip(p, n);
    if (flags & B200_F_STRICT) {
        const int rc = zlib_strict_check(p, n, s);
        if (rc) return rc;
        return inflate_host(p + s, n - s - 4, out, cap, nullptr, out_n, full_n, flags, p + n - 4);
    }
    return b200_inflate(p + s, n - s, out, cap, out_n, full_n, flags);
}

// RFC 1952 member header: fixed 10 bytes, then FEXTRA / FNAME / FCOMMENT / FHCRC as flagged.  Returns the offset of the
// raw stream, or 0 if this is not a gzip member / the header runs past the input.
static size_t gzip_header_skip(const uint8_t* in, size_t n) {
    if (n < 18 || in[0] != 0x1F || in[1] != 0x8B || in[2] != 8 || (in[3] & 0xE0)) return 0;
    const uint8_t flg = in[3];
    size_t p = 10;
    if (flg & 4) { if (p + 2 > n) return 0; p += 2 + ((size_t)in[p] | ((size_t)in[p + 1] << 8)); }
    if (flg & 8) { while (p < n && in[p]) p++; p++; }
    if (flg & 16) { while (p < n && in[p]) p++; p++; }
    if (flg & 2) p += 2;
    return p + 8 <= n ? p : 0;
}

// gzip (RFC 1952) around the raw stream: an addition next to decompressZlib (the reference has no gzip entry point).
// The first member is decoded; the CRC-32 / ISIZE trailer is ignored unless B200_F_STRICT is set, then the CRC-32 is
// computed on the GPU over the decoded bytes.
int b200_inflate_gzip(const void* in, size_t n, void* out, size_t cap, size_t* out_n, size_t* full_n, unsigned flags) {
    if (!in) return B200_E_ARG;
    if ((!out && cap)) return B200_E_ARG;
    const uint8_t* p = (const uint8_t*)in;
    const size_t s = gzip_header_skip(p, n);
    if (!s) return n < 18 ? B200_E_OVERRUN : B200_E_DATA;
    if (flags & B200_F_STRICT) return inflate_host(p + s, n - s - 8, out, cap, nullptr, out_n, full_n, flags, p + n - 8, 2);
    return b200_inflate(p + s, n - s, out, cap, out_n, full_n, flags);
}

int b200_inflate_gzip_alloc(const void* in, size_t n, void** out, size_t* out_n, unsigned flags) {
    if (!out || !out_n) return B200_E_ARG;
    *out = nullptr; *out_n = 0;
    if (!in) return B200_E_ARG;
    const uint8_t* p = (const uint8_t*)in;
    const size_t s = gzip_header_skip(p, n);
    if (!s) return n < 18 ? B200_E_OVERRUN : B200_E_DATA;
    if (flags & B200_F_STRICT) return inflate_host(p + s, n - s - 8, nullptr, 0, out, out_n, nullptr, flags, p + n - 8, 2);
    return b200_inflate_alloc(p + s, n - s, out, out_n, flags);
}

int b200_inflate_zlib_alloc(const void* in, size_t n, void** out, size_t* out_n, unsigned flags) {
    if (!in || n < 2) { if (out) *out = nullptr; if (out_n) *out_n = 0; return B200_E_OVERRUN; }
    if (!out || !out_n) return B200_E_ARG;
    const uint8_t* p = (const uint8_t*)in;
    const size_t s = zlib_header_skip(p, n);
    if (flags & B200_F_STRICT) {
        *out = nullptr; *out_n = 0;
        const int rc = zlib_strict_check(p, n, s);
        if (rc) return rc;
        return inflate_host(p + s, n - s - 4, nullptr, 0, out, out_n, nullptr, flags, p + n - 4);
    }
    return b200_inflate_alloc(p + s, n - s, out, out_n, flags);
}

// ------------------------------------------------------------------------------------------------
// File-path API: what deflate::compress(string, string, int) (reference include/deflate.hpp:755-777) and
// inflate::decompress(string, string) (include/inflate.hpp:390-408) do -- stream a file through the codec in slices with
// bounded memory -- as pinned, double-buffered host <-> device pipelines.  (The reference reads 32 KB at a time and its
// inflate side fails on multi-block files; these handle any size, any of the stream kinds the
// memory API handles.)
static int pin_ensure(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return B200_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    if (cudaHostAlloc(p, need, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return B200_E_NOMEM; }
    *cap = need;
    return B200_OK;
}
static bool read_fully(int fd, void* buf, size_t len, off_t off) {
    uint8_t* p = (uint8_t*)buf;
    while (len) {
        const ssize_t r = pread(fd, p, len, off);
        if (r <= 0) return false;
        p += r; len -= (size_t)r; off += r;
    }
    return true;
}
static bool write_fully(int fd, const void* buf, size_t len) {
    const uint8_t* p = (const uint8_t*)buf;
    while (len) {
        const ssize_t r = write(fd, p, len);
        if (r <= 0) return false;
        p += r; len -= (size_t)r;
    }
    return true;
}
struct FdCloser { int fd; ~FdCloser() { if (fd >= 0) close(fd); } };

int b200_deflate_compress_file(const char* in_path, const char* out_path, int level, unsigned flags, size_t* in_n, size_t* out_n) {
    if (!in_path || !out_path || level < 0 || level > 3 || (flags & B200_F_NOT_LAST)) return B200_E_ARG;
    b200_ctx* c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    ON_DEVICE(c);
    FdCloser fi{open(in_path, O_RDONLY)};
    if (fi.fd < 0) return B200_E_IO;
    struct stat sb;
    if (fstat(fi.fd, &sb) != 0 || !S_ISREG(sb.st_mode)) return B200_E_IO;
    FdCloser fo{open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644)};
    if (fo.fd < 0) return B200_E_IO;
    const uint64_t n = (uint64_t)sb.st_size;
    const size_t S = c->file_slice;
    const uint64_t nslices = n ? (n + S - 1) / S : 1;
    const size_t ocap = b200_deflate_bound(S);
    const int frame = frame_mode(flags);
    const unsigned cflags = flags & B200_F_NO_INDEX;
    uint64_t written = 0;
    if (frame) {
        if (!write_fully(fo.fd, frame == 1 ? kZlibHeader : kGzipHeader, kFrameLen[frame])) return B200_E_IO;
        written += kFrameLen[frame];
    }
    if ((rc = c->total.ensure(64))) return rc;
    uint32_t* d_sum = (uint32_t*)((uint64_t*)c->total.p + 5);       // running checksum of the input (framing)
    for (int k = 0; k < 2; k++) {
        if ((rc = pin_ensure(&c->pin_in[k], &c->pin_in_cap[k], S))) return rc;
        if ((rc = pin_ensure(&c->pin_out[k], &c->pin_out_cap[k], ocap))) return rc;
        if ((rc = c->file_in[k].ensure(S + 64)) || (rc = c->file_out[k].ensure(ocap + 64))) return rc;
        if (!c->file_ev[k]) CK(cudaEventCreateWithFlags(&c->file_ev[k], cudaEventDisableTiming));
    }
    if (c->mailbox_cap < 4) {
        if (c->mailbox) cudaFreeHost(c->mailbox);
        c->mailbox = nullptr; c->mailbox_cap = 0;
        CK(cudaHostAlloc((void**)&c->mailbox, 32 * 8, cudaHostAllocDefault));
        c->mailbox_cap = 32;
    }
    auto retire = [&](uint64_t j) -> int {           // slice j: wait for its kernels, fetch its bytes, append them to the file
        const int idx = (int)(j & 1);
        CK(cudaEventSynchronize(c->file_ev[idx]));
        const uint64_t sz = c->mailbox[idx];
        if (sz > ocap) return B200_E_OUTPUT;
        CK(cudaMemcpyAsync(c->pin_out[idx], c->file_out[idx].p, sz, cudaMemcpyDeviceToHost, c->s_out));
        CK(cudaStreamSynchronize(c->s_out));
        if (!write_fully(fo.fd, c->pin_out[idx], sz)) return B200_E_IO;
        written += sz;
        return B200_OK;
    };
    for (uint64_t k = 0; k < nslices; k++) {
        const int idx = (int)(k & 1);
        if (k >= 2 && (rc = retire(k - 2))) return rc;                 // frees pin_in / file_in / file_out / pin_out [idx]
        const uint64_t off = k * S;
        const size_t len = (size_t)(n - off < S ? n - off : S);
        if (len && !read_fully(fi.fd, c->pin_in[idx], len, (off_t)off)) return B200_E_IO;      // the GPU works on slice k - 1 meanwhile
        if (len) CK(cudaMemcpyAsync(c->file_in[idx].p, c->pin_in[idx], len, cudaMemcpyHostToDevice, c->stream));
        const bool last = k + 1 == nslices;
        if ((rc = b200_deflate_compress_dev(c, c->file_in[idx].p, len, level, cflags | (last ? 0u : B200_F_NOT_LAST), c->file_out[idx].p,
                                            c->file_out[idx].cap, nullptr, nullptr, nullptr, c->stream)))
            return rc;
        if (frame && (rc = checksum_dev(c, (const uint8_t*)c->file_in[idx].p, len, frame, d_sum, c->stream, k ? d_sum : nullptr))) return rc;
        CK(cudaMemcpyAsync(&c->mailbox[idx], c->total.p, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaEventRecord(c->file_ev[idx], c->stream));
    }
    if (nslices >= 2 && (rc = retire(nslices - 2))) return rc;
    if ((rc = retire(nslices - 1))) return rc;
    if (frame) {
        uint32_t sum = 0;
        CK(cudaMemcpyAsync(&sum, d_sum, 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        uint8_t t[8];
        if (frame == 1) { t[0] = (uint8_t)(sum >> 24); t[1] = (uint8_t)(sum >> 16); t[2] = (uint8_t)(sum >> 8); t[3] = (uint8_t)sum; }
        else for (int k = 0; k < 4; k++) { t[k] = (uint8_t)(sum >> (8 * k)); t[4 + k] = (uint8_t)((uint32_t)n >> (8 * k)); }
        if (!write_fully(fo.fd, t, frame == 1 ? 4 : 8)) return B200_E_IO;
        written += frame == 1 ? 4 : 8;
    }
    if (in_n) *in_n = (size_t)n;
    if (out_n) *out_n = (size_t)written;
    return B200_OK;
}

int b200_inflate_file(const char* in_path, const char* out_path, unsigned flags, size_t* in_n, size_t* out_n) {
    if (!in_path || !out_path) return B200_E_ARG;
    b200_ctx* c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    FdCloser fi{open(in_path, O_RDONLY)};
    if (fi.fd < 0) return B200_E_IO;
    struct stat sb;
    if (fstat(fi.fd, &sb) != 0 || !S_ISREG(sb.st_mode)) return B200_E_IO;
    const uint64_t n = (uint64_t)sb.st_size;
    if (in_n) *in_n = (size_t)n;
    const size_t S = c->file_slice;
    const size_t slack = CHUNK + 4096, lead = 16;
    uint64_t total = 0;
    bool windowed = n > 2 * S;
    FdCloser fo{-1};
    if (windowed) {
        std::lock_guard<std::mutex> lk(c->mu);
        ON_DEVICE(c);
        const uint64_t nw = (n + S - 1) / S;
        const size_t wcap = S + slack + lead + 64;
        for (int k = 0; k < 2; k++)
            if ((rc = pin_ensure(&c->pin_in[k], &c->pin_in_cap[k], wcap))) return rc;
        if ((rc = c->file_in[0].ensure(wcap))) return rc;
        auto window = [&](uint64_t k, uint64_t* ws, uint64_t* we) {
            const uint64_t lo = k * S, hi = lo + S < n ? lo + S : n;
            *ws = k == 0 ? 0 : (lo - lead) & ~15ull;
            *we = hi + slack < n ? hi + slack : n;
        };
        auto read_window = [&](uint64_t k) -> bool {
            uint64_t ws, we;
            window(k, &ws, &we);
            return read_fully(fi.fd, c->pin_in[k & 1], (size_t)(we - ws), (off_t)ws);
        };
        std::future<bool> reader = std::async(std::launch::async, read_window, (uint64_t)0);
        std::future<bool> writer;
        for (uint64_t k = 0; k < nw && windowed; k++) {
            const bool got = reader.get();
            if (k + 1 < nw) reader = std::async(std::launch::async, read_window, k + 1);       // the next window arrives while this one is decoded
            if (!got) { if (writer.valid()) writer.get(); if (reader.valid()) reader.get(); return B200_E_IO; }
            uint64_t ws, we;
            window(k, &ws, &we);
            const uint64_t lo = k * S, hi = lo + S < n ? lo + S : n;
            CK(cudaMemcpyAsync(c->file_in[0].p, c->pin_in[k & 1], we - ws, cudaMemcpyHostToDevice, c->stream));
            bool valid = false;
            uint64_t wtotal = 0, units = 0, nxt = 0;
            rc = inflate_chunked(c, (const uint8_t*)c->file_in[0].p, we - ws, k == 0 ? 0 : lo - ws, hi - ws, k == 0, we == n, nullptr, 0, flags,
                                 c->stream, &valid, &wtotal, &units, &nxt, &c->file_out[0]);
            if (rc == B200_OK && !valid) {
                // not (only) this library's chunks -- a foreign stream, or stored user data that contains the separator
                // pattern: start over with the whole file through the memory API
                windowed = false;
                break;
            }
            if (rc) { if (writer.valid()) writer.get(); if (reader.valid()) reader.get(); return rc; }
            if (writer.valid() && !writer.get()) { if (reader.valid()) reader.get(); return B200_E_IO; }      // pin_out[k & 1] was used by window k - 2 ... and the file stays in order
            if (wtotal) {
                if ((rc = pin_ensure(&c->pin_out[k & 1], &c->pin_out_cap[k & 1], wtotal))) { if (reader.valid()) reader.get(); return rc; }
                CK(cudaMemcpyAsync(c->pin_out[k & 1], c->file_out[0].p, wtotal, cudaMemcpyDeviceToHost, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                if (fo.fd < 0) { fo.fd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644); if (fo.fd < 0) { if (reader.valid()) reader.get(); return B200_E_IO; } }
                const void* src = c->pin_out[k & 1];
                const int fd = fo.fd;
                writer = std::async(std::launch::async, [fd, src, wtotal]() { return write_fully(fd, src, (size_t)wtotal); });
            }
            total += wtotal;
        }
        if (reader.valid()) reader.get();
        if (writer.valid() && !writer.get()) return B200_E_IO;
        if (windowed) {
            if (fo.fd < 0) { fo.fd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644); if (fo.fd < 0) return B200_E_IO; }
            if (out_n) *out_n = (size_t)total;
            return B200_OK;
        }
    }
    // small files and streams this library did not write: the whole stream through the memory API
    std::vector<uint8_t> in(n);
    if (n && !read_fully(fi.fd, in.data(), n, 0)) return B200_E_IO;
    void* out = nullptr;
    size_t on = 0;
    rc = inflate_host(in.data(), n, nullptr, 0, &out, &on, nullptr, flags);
    if (rc) { if (out) free(out); return rc; }
    if (fo.fd >= 0) close(fo.fd);
    fo.fd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    const bool ok = fo.fd >= 0 && write_fully(fo.fd, out, on);
    free(out);
    if (!ok) return B200_E_IO;
    if (out_n) *out_n = on;
    return B200_OK;
}

}  // extern "C"
