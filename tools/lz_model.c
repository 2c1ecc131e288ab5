/* Development tool (not product, not oracle): CPU model of the GPU LZ77 parse, used to explore the
 * ratio effect of matcher parameters before spending GPU time.  It emulates the warp-window semantics
 * of deflate.hpp_b200/csrc/lz77.cuh (32 positions probed at once, lookup before insert, greedy or
 * lazy selection over the window) and reports token counts and an entropy-coded size estimate.
 *   gcc -O2 -o /tmp/lz_model tools/lz_model.c -lm && /tmp/lz_model file [hash_bits seg warm minmatch ways lazy]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHUNK 65536
static int HB = 11, SEG = 8192, WARM = 4096, MINM = 4, WAYS = 1, LAZY = 0, NICE = 258, HASH3 = 0;

static uint32_t ld4(const uint8_t* d, uint32_t p) { uint32_t v; memcpy(&v, d + p, 4); return v; }
static uint32_t hashf(uint32_t w) { if (HASH3) w &= 0xFFFFFF; return (w * 0x9E3779B1u) >> (32 - HB); }

static double cost_bits(const uint32_t* h, int n) {
    double tot = 0, bits = 0;
    for (int i = 0; i < n; i++) tot += h[i];
    for (int i = 0; i < n; i++) if (h[i]) bits += h[i] * -log2(h[i] / tot);
    return bits;
}
static int len_sym(int len, int* ex) { int l = len - 3; if (l < 8) { *ex = 0; return l; } if (len == 258) { *ex = 0; return 28; } int k = 31 - __builtin_clz(l); *ex = k - 2; return 4 * (k - 1) + ((l >> (k - 2)) & 3); }
static int dist_sym(int d, int* ex) { d--; if (d < 4) { *ex = 0; return d; } int k = 31 - __builtin_clz(d); *ex = k - 1; return 2 * k + ((d >> (k - 1)) & 1); }

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    if (argc > 2) HB = atoi(argv[2]);
    if (argc > 3) SEG = atoi(argv[3]);
    if (argc > 4) WARM = atoi(argv[4]);
    if (argc > 5) MINM = atoi(argv[5]);
    if (argc > 6) WAYS = atoi(argv[6]);
    if (argc > 7) LAZY = atoi(argv[7]);
    if (argc > 8) HASH3 = atoi(argv[8]);
    FILE* f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* buf = malloc(n + 64); memset(buf + n, 0, 64);
    if (fread(buf, 1, n, f) != (size_t)n) return 2;
    double total_bits = 0; long ntok = 0, nmatch = 0, mbytes = 0;
    uint16_t* tab = malloc(sizeof(uint16_t) * (1 << HB) * WAYS);
    for (long c0 = 0; c0 < n; c0 += CHUNK) {
        uint32_t clen = n - c0 < CHUNK ? n - c0 : CHUNK;
        uint8_t d[CHUNK + 64]; memcpy(d, buf + c0, clen); memset(d + clen, 0, 64);
        uint32_t hl[288] = {0}, hd[32] = {0}; hl[256] = 1;
        double extra = 0;
        for (uint32_t s0 = 0; s0 < clen; s0 += SEG) {
            uint32_t s1 = s0 + SEG < clen ? s0 + SEG : clen;
            memset(tab, 0, sizeof(uint16_t) * (1 << HB) * WAYS);
            uint32_t ws = s0 > (uint32_t)WARM ? s0 - WARM : 0;
            for (uint32_t p = ws; p < s0; p++) { uint32_t h = hashf(ld4(d, p)); for (int w = WAYS - 1; w > 0; w--) tab[h * WAYS + w] = tab[h * WAYS + w - 1]; tab[h * WAYS] = p; }
            uint32_t pos = s0;
            while (pos < s1) {
                uint32_t len[33], dist[33];
                /* lookup for all 32 (+1 lookahead for lazy) lanes before any insert */
                for (int l = 0; l < 32; l++) {
                    uint32_t p = pos + l; len[l] = 0; dist[l] = 0;
                    if (p >= s1) continue;
                    uint32_t avail = s1 - p; if (avail < (uint32_t)MINM) continue;
                    uint32_t w4 = ld4(d, p), h = hashf(w4);
                    uint32_t maxl = avail < 258 ? avail : 258;
                    for (int w = 0; w < WAYS + 1; w++) {
                        uint32_t q;
                        if (w < WAYS) q = tab[h * WAYS + w]; else { if (p == 0) break; q = p - 1; }
                        if (q >= p || p - q > 32768) continue;
                        uint32_t l2 = 0; while (l2 < maxl && d[q + l2] == d[p + l2]) l2++;
                        if (l2 >= (uint32_t)MINM && l2 > len[l]) { len[l] = l2; dist[l] = p - q; }
                    }
                }
                for (int l = 0; l < 32; l++) { uint32_t p = pos + l; if (p < s1 && s1 - p >= 4) { uint32_t h = hashf(ld4(d, p)); for (int w = WAYS - 1; w > 0; w--) tab[h * WAYS + w] = tab[h * WAYS + w - 1]; tab[h * WAYS] = p; } }
                uint32_t valid = s1 - pos < 32 ? s1 - pos : 32, cur = 0;
                uint32_t limit = LAZY ? 31 : 32;   /* lazy: lane 31 is only a lookahead */
                if (valid < 32) limit = valid;
                while (cur < limit) {
                    if (len[cur] >= (uint32_t)MINM) {
                        if (LAZY && cur + 1 < valid && len[cur + 1] > len[cur]) { hl[d[pos + cur]]++; ntok++; cur++; continue; }
                        int ex; hl[257 + len_sym(len[cur], &ex)]++; extra += ex; hd[dist_sym(dist[cur], &ex)]++; extra += ex;
                        ntok++; nmatch++; mbytes += len[cur]; cur += len[cur];
                    } else { hl[d[pos + cur]]++; ntok++; cur++; }
                }
                pos += cur;
            }
        }
        double b = cost_bits(hl, 288) + cost_bits(hd, 32) + extra + 600;
        if (b > clen * 8.0 + 80) b = clen * 8.0 + 80;
        total_bits += b;
    }
    printf("HB=%d SEG=%d WARM=%d MINM=%d WAYS=%d LAZY=%d H3=%d: ratio %.4f  tokens/byte %.4f  match-frac %.3f avg-match %.2f\n", HB, SEG, WARM, MINM,
           WAYS, LAZY, HASH3, total_bits / 8 / n, (double)ntok / n, (double)nmatch / ntok, nmatch ? (double)mbytes / nmatch : 0.0);
    return 0;
}
