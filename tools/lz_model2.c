/* Development tool: CPU model of the two-phase fast matcher (chunk-wide table, tile-lagged inserts,
 * per-segment greedy parse).  gcc -O2 -o /tmp/lz_model2 tools/lz_model2.c -lm
 * usage: lz_model2 file hash_bits tile cap peer minmatch */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CHUNK 65536
#define SEG 8192
static uint32_t ld4(const uint8_t* d, uint32_t p) { uint32_t v; memcpy(&v, d + p, 4); return v; }
static double cost_bits(const uint32_t* h, int n) { double tot = 0, bits = 0; for (int i = 0; i < n; i++) tot += h[i]; for (int i = 0; i < n; i++) if (h[i]) bits += h[i] * -log2(h[i] / tot); return bits; }
static int len_sym(int len, int* ex) { int l = len - 3; if (l < 8) { *ex = 0; return l; } if (len == 258) { *ex = 0; return 28; } int k = 31 - __builtin_clz(l); *ex = k - 2; return 4 * (k - 1) + ((l >> (k - 2)) & 3); }
static int dist_sym(int d, int* ex) { d--; if (d < 4) { *ex = 0; return d; } int k = 31 - __builtin_clz(d); *ex = k - 1; return 2 * k + ((d >> (k - 1)) & 1); }
int main(int argc, char** argv) {
    int HB = atoi(argv[2]), TILE = atoi(argv[3]), CAP = atoi(argv[4]), PEER = atoi(argv[5]), MINM = atoi(argv[6]); int ALL = argc > 7 ? atoi(argv[7]) : 0;
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* buf = malloc(n + 64); memset(buf + n, 0, 64); if (fread(buf, 1, n, f) != (size_t)n) return 2;
    double total_bits = 0; long ntok = 0, nmatch = 0, mbytes = 0;
    uint32_t* tab = malloc(4u << HB);
    static uint32_t clen_[CHUNK], cdist[CHUNK];
    for (long c0 = 0; c0 < n; c0 += CHUNK) {
        uint32_t clen = n - c0 < CHUNK ? n - c0 : CHUNK;
        static uint8_t d[CHUNK + 64]; memcpy(d, buf + c0, clen); memset(d + clen, 0, 64);
        memset(tab, 0, 4u << HB);
        for (uint32_t t0 = 0; t0 < clen; t0 += TILE) {
            uint32_t t1 = t0 + TILE < clen ? t0 + TILE : clen;
            for (uint32_t p = t0; p < t1; p++) {
                clen_[p] = 0; cdist[p] = 0;
                if (p + 4 > clen) continue;
                uint32_t w4 = ld4(d, p), h = (w4 * 0x9E3779B1u) >> (32 - HB);
                uint32_t cands[8]; int nc = 0;
                cands[nc++] = tab[h];
                if (p > 0) cands[nc++] = p - 1;
                if (PEER >= 2) { for (uint32_t k = 2; k <= (uint32_t)PEER + 1 && k <= p; k++) cands[nc++] = p - k; }
                else if (PEER) { /* nearest earlier position in the same 32-lane group with the same hash */
                    for (uint32_t q = p; q-- > (p & ~31u);) { uint32_t h2 = (ld4(d, q) * 0x9E3779B1u) >> (32 - HB); if (h2 == h) { cands[nc++] = q; break; } }
                }
                uint32_t maxl = clen - p < 258 ? clen - p : 258;
                uint32_t bscore = 0;
                for (int k = 0; k < nc; k++) {
                    uint32_t q = cands[k]; if (q >= p || p - q > 32768) continue;
                    uint32_t l = 0; while (l < maxl && d[q + l] == d[p + l]) l++;
                    if (l < (uint32_t)MINM) continue;
                    uint32_t score = CAP ? (l < (uint32_t)CAP ? l : (uint32_t)CAP) : l;   /* phase B only sees min(len, CAP) */
                    if (score > bscore || (score == bscore && p - q < cdist[p])) { bscore = score; clen_[p] = l; cdist[p] = p - q; }
                    if (!ALL && k == 0 && clen_[p]) break;   /* table hit wins; others only as fallback */
                }
            }
            for (uint32_t p = t0; p < t1; p++) if (p + 4 <= clen) { uint32_t h = (ld4(d, p) * 0x9E3779B1u) >> (32 - HB); if (tab[h] < p) tab[h] = p; }
        }
        uint32_t hl[288] = {0}, hd[32] = {0}; hl[256] = 1; double extra = 0;
        for (uint32_t s0 = 0; s0 < clen; s0 += SEG) {
            uint32_t s1 = s0 + SEG < clen ? s0 + SEG : clen, p = s0;
            while (p < s1) {
                uint32_t l = clen_[p]; if (l > s1 - p) l = s1 - p;
                if (l >= (uint32_t)MINM) { int ex; hl[257 + len_sym(l, &ex)]++; extra += ex; hd[dist_sym(cdist[p], &ex)]++; extra += ex; ntok++; nmatch++; mbytes += l; p += l; }
                else { hl[d[p]]++; ntok++; p++; }
            }
        }
        double b = cost_bits(hl, 288) + cost_bits(hd, 32) + extra + 600; if (b > clen * 8.0 + 80) b = clen * 8.0 + 80; total_bits += b;
    }
    printf("HB=%d TILE=%d PEER=%d MINM=%d: ratio %.4f tokens/byte %.4f match-frac %.3f avg-match %.2f\n", HB, TILE, PEER, MINM, total_bits / 8 / n, (double)ntok / n, (double)nmatch / ntok, nmatch ? (double)mbytes / nmatch : 0.0);
    return 0;
}
