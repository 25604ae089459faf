// One-shot DEFLATE (RFC 1951) compressor for the BGZF members of the BAM writer (row f1): every member holds at most
// 65,280 payload bytes and is independent of the others, so there is no window to carry, positions fit 16 bits and the
// whole member is one block.  Greedy LZ77 with a single-probe hash of the next four bytes (a miss streak lengthens the
// step, which is what the quality strings need; the positions inside short matches are entered as well), one dynamic Huffman block with length-limited codes, or a stored block
// when that is smaller.  The reference writes its BAM through libStatGen's BGZF at zlib's default level
// (src/BwtMapper.cpp:2131-2143, misc/bam/bgzf.c); the records, not the compressed bytes, are what must match.
#include "fq_deflate.h"

#include <algorithm>
#include <cstring>
#include <memory>

namespace fqb {
namespace {

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Tables {
    uint8_t len_sym[259];          // match length -> length symbol - 257
    uint8_t dist_sym[512];         // distance - 1 (< 256) or 256 + ((distance - 1) >> 7) -> distance symbol
    Tables() {
        for (int s = 0; s < 29; ++s)
            for (int l = kLenBase[s]; l <= (s == 28 ? 258 : kLenBase[s + 1] - 1) && l <= 258; ++l) len_sym[l] = (uint8_t)s;
        len_sym[258] = 28;
        for (int s = 0; s < 30; ++s) {
            const int lo = kDistBase[s], hi = s == 29 ? 32768 : kDistBase[s + 1] - 1;
            for (int d = lo; d <= hi; ++d) {
                if (d - 1 < 256) dist_sym[d - 1] = (uint8_t)s;
                else dist_sym[256 + ((d - 1) >> 7)] = (uint8_t)s;
            }
        }
    }
};
const Tables kT;
inline int dist_symbol(unsigned d) { return d <= 256 ? kT.dist_sym[d - 1] : kT.dist_sym[256 + ((d - 1) >> 7)]; }

inline uint32_t load32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

constexpr int kHashBits = 14;
constexpr size_t kMaxIn = 65535;

struct Scratch {
    uint16_t head[1 << kHashBits];
    uint32_t tok[kMaxIn + 1];       // literal: the byte; match: 1 << 31 | length << 16 | distance (distance < 65536)
};

// Huffman code lengths of at most `limit` bits for the symbols with freq > 0 (at least two of them, see the caller).
void code_lengths(const uint32_t *freq, int n, int limit, uint8_t *len) {
    struct Sym { uint32_t f; uint16_t s; };
    Sym sym[288];
    int m = 0;
    for (int s = 0; s < n; ++s) { len[s] = 0; if (freq[s]) sym[m++] = Sym{freq[s], (uint16_t)s}; }
    std::sort(sym, sym + m, [](const Sym &a, const Sym &b) { return a.f < b.f || (a.f == b.f && a.s < b.s); });
    // two-queue merge: leaves 0..m-1 in weight order, internal nodes m..2m-2 are created in weight order as well
    uint64_t w[2 * 288]; uint16_t parent[2 * 288]; uint16_t depth[2 * 288];
    for (int i = 0; i < m; ++i) w[i] = sym[i].f;
    int leaf = 0, inner = m;
    for (int k = m; k < 2 * m - 1; ++k) {
        int pick[2];
        for (int t = 0; t < 2; ++t) pick[t] = (leaf < m && (inner >= k || w[leaf] <= w[inner])) ? leaf++ : inner++;
        w[k] = w[pick[0]] + w[pick[1]];
        parent[pick[0]] = parent[pick[1]] = (uint16_t)k;
    }
    depth[2 * m - 2] = 0;
    for (int v = 2 * m - 3; v >= 0; --v) depth[v] = (uint16_t)(depth[parent[v]] + 1);
    // fold the lengths above the limit into it and give back what the Kraft sum lost: drop one code of the limit length,
    // split the deepest shorter code into two; every round lowers the sum by one unit of 2^-limit
    int cnt[290] = {0};
    for (int i = 0; i < m; ++i) ++cnt[depth[i] > limit ? limit : depth[i]];
    uint64_t total = 0;
    for (int l = 1; l <= limit; ++l) total += (uint64_t)cnt[l] << (limit - l);
    while (total > ((uint64_t)1 << limit)) {
        --cnt[limit];
        for (int l = limit - 1; l > 0; --l) if (cnt[l]) { --cnt[l]; cnt[l + 1] += 2; break; }
        --total;
    }
    int i = 0;                                                  // rarest symbols take the longest codes
    for (int l = limit; l >= 1; --l) for (int c = cnt[l]; c > 0; --c) len[sym[i++].s] = (uint8_t)l;
}

// canonical codes, bit-reversed because DEFLATE sends Huffman codes most significant bit first in an LSB-first stream
void assign_codes(const uint8_t *len, int n, uint16_t *code) {
    int cnt[16] = {0};
    for (int s = 0; s < n; ++s) ++cnt[len[s]];
    cnt[0] = 0;
    uint32_t next[16]; uint32_t c = 0;
    for (int l = 1; l <= 15; ++l) { c = (c + (uint32_t)cnt[l - 1]) << 1; next[l] = c; }
    for (int s = 0; s < n; ++s) {
        const int l = len[s];
        if (!l) { code[s] = 0; continue; }
        uint32_t v = next[l]++, r = 0;
        for (int b = 0; b < l; ++b) { r = (r << 1) | (v & 1); v >>= 1; }
        code[s] = (uint16_t)r;
    }
}

struct BitWriter {
    uint8_t *p; uint64_t bb = 0; unsigned bc = 0;
    void put(uint32_t v, unsigned n) { bb |= (uint64_t)v << bc; bc += n; }
    void flush() { memcpy(p, &bb, 8); p += bc >> 3; bb >>= bc & ~7u; bc &= 7; }     // writes 8 bytes, keeps < 8 bits
    uint8_t *finish() { flush(); if (bc) { *p++ = (uint8_t)bb; bb = 0; bc = 0; } return p; }
};

size_t stored(const uint8_t *in, size_t n, uint8_t *out) {
    out[0] = 1;
    out[1] = (uint8_t)(n & 0xff); out[2] = (uint8_t)(n >> 8);
    out[3] = (uint8_t)(~n & 0xff); out[4] = (uint8_t)((~n >> 8) & 0xff);
    memcpy(out + 5, in, n);
    return n + 5;
}

}  // namespace

size_t deflate_fast(const uint8_t *in, size_t n, uint8_t *out, size_t cap) {
    if (n > kMaxIn || cap < deflate_fast_bound(n)) return 0;
    if (n == 0) { out[0] = 3; out[1] = 0; return 2; }              // an empty final block with the fixed code
    static thread_local std::unique_ptr<Scratch> S;                // one per compressing thread, freed when it ends
    if (!S) S.reset(new Scratch);
    memset(S->head, 0, sizeof(S->head));
    uint32_t lit_freq[288] = {0}, dist_freq[32] = {0};
    uint32_t *tok = S->tok;
    size_t n_tok = 0;
    uint64_t extra_bits = 0;

    // ---- LZ77
    size_t pos = 0, misses = 0;
    const size_t last = n >= 8 ? n - 8 : 0;                        // matches start where eight bytes can still be loaded
    while (pos < last) {
        const uint32_t here = load32(in + pos);
        const uint32_t h = (here * 2654435761u) >> (32 - kHashBits);
        const size_t cand = S->head[h];
        S->head[h] = (uint16_t)(pos + 1);
        if (cand && load32(in + cand - 1) == here && pos - (cand - 1) <= 32768) {
            const uint8_t *a = in + pos, *b = in + cand - 1;
            const size_t max_len = std::min<size_t>(258, n - pos);
            size_t len = 4;
            while (len + 8 <= max_len) {
                const uint64_t x = load64(a + len) ^ load64(b + len);
                if (x) { len += (size_t)__builtin_ctzll(x) >> 3; goto matched; }
                len += 8;
            }
            while (len < max_len && a[len] == b[len]) ++len;
        matched:
            const unsigned dist = (unsigned)(pos - (cand - 1));
            tok[n_tok++] = 0x80000000u | ((uint32_t)len << 16) | dist;
            const int ls = kT.len_sym[len], ds = dist_symbol(dist);
            ++lit_freq[257 + ls]; ++dist_freq[ds];
            extra_bits += kLenExtra[ls] + kDistExtra[ds];
            if (len <= 8) {                                         // short matches: every position inside stays findable
                const size_t e = std::min(pos + len, last);         // (7 % smaller members on FASTQ-like text for ~20 % of the speed)
                for (size_t q = pos + 1; q < e; ++q) S->head[(load32(in + q) * 2654435761u) >> (32 - kHashBits)] = (uint16_t)(q + 1);
                pos += len;
            } else {                                                // long ones: only their last position
                pos += len;
                if (pos < last) S->head[(load32(in + pos - 1) * 2654435761u) >> (32 - kHashBits)] = (uint16_t)pos;
            }
            misses = 0;
            continue;
        }
        const size_t step = 1 + (misses++ >> 5);                    // incompressible stretches are walked faster and faster
        const size_t stop = std::min(pos + step, last);
        do { ++lit_freq[in[pos]]; tok[n_tok++] = in[pos]; } while (++pos < stop);
    }
    for (; pos < n; ++pos) { ++lit_freq[in[pos]]; tok[n_tok++] = in[pos]; }
    lit_freq[256] = 1;

    // ---- codes.  Both alphabets get at least two used symbols so that the codes are complete (what every inflater accepts).
    { int used = 0; for (int s = 0; s < 30; ++s) used += dist_freq[s] != 0; for (int s = 0; used < 2; ++s) if (!dist_freq[s]) { dist_freq[s] = 1; ++used; } }
    { int used = 0; for (int s = 0; s < 286; ++s) used += lit_freq[s] != 0; for (int s = 0; used < 2; ++s) if (!lit_freq[s]) { lit_freq[s] = 1; ++used; } }
    uint8_t lit_len[288], dist_len[32];
    uint16_t lit_code[288], dist_code[32];
    code_lengths(lit_freq, 286, 15, lit_len);
    code_lengths(dist_freq, 30, 15, dist_len);
    int hlit = 286; while (hlit > 257 && !lit_len[hlit - 1]) --hlit;
    int hdist = 30; while (hdist > 1 && !dist_len[hdist - 1]) --hdist;

    // code lengths, run-length coded with the symbols 16 (repeat previous 3-6), 17 (zeros 3-10), 18 (zeros 11-138)
    uint8_t all[320];
    memcpy(all, lit_len, (size_t)hlit);
    memcpy(all + hlit, dist_len, (size_t)hdist);
    const int n_all = hlit + hdist;
    uint8_t rl_sym[320], rl_extra[320];
    int n_rl = 0;
    uint32_t pre_freq[19] = {0};
    for (int i = 0; i < n_all;) {
        const uint8_t v = all[i];
        int run = 1;
        while (i + run < n_all && all[i + run] == v) ++run;
        i += run;
        if (v == 0) {
            while (run >= 11) { const int r = std::min(run, 138); rl_sym[n_rl] = 18; rl_extra[n_rl++] = (uint8_t)(r - 11); run -= r; }
            if (run >= 3) { rl_sym[n_rl] = 17; rl_extra[n_rl++] = (uint8_t)(run - 3); run = 0; }
        } else {
            rl_sym[n_rl] = v; rl_extra[n_rl++] = 0; --run;
            while (run >= 3) { const int r = std::min(run, 6); rl_sym[n_rl] = 16; rl_extra[n_rl++] = (uint8_t)(r - 3); run -= r; }
        }
        while (run-- > 0) { rl_sym[n_rl] = v; rl_extra[n_rl++] = 0; }
    }
    for (int i = 0; i < n_rl; ++i) ++pre_freq[rl_sym[i]];
    { int used = 0; for (int s = 0; s < 19; ++s) used += pre_freq[s] != 0; for (int s = 0; used < 2; ++s) if (!pre_freq[s]) { pre_freq[s] = 1; ++used; } }
    uint8_t pre_len[19]; uint16_t pre_code[19];
    code_lengths(pre_freq, 19, 7, pre_len);
    int hclen = 19; while (hclen > 4 && !pre_len[kPreOrder[hclen - 1]]) --hclen;

    // ---- stored or dynamic, whichever is smaller
    uint64_t bits = 3 + 5 + 5 + 4 + 3 * (uint64_t)hclen + extra_bits;
    for (int i = 0; i < n_rl; ++i) bits += pre_len[rl_sym[i]] + (rl_sym[i] == 16 ? 2 : rl_sym[i] == 17 ? 3 : rl_sym[i] == 18 ? 7 : 0);
    for (int s = 0; s < 286; ++s) bits += (uint64_t)lit_freq[s] * lit_len[s];
    for (int s = 0; s < 30; ++s) bits += (uint64_t)dist_freq[s] * dist_len[s];
    // (the dummy symbols above carry a frequency of one each: the estimate is at most a few bits high)
    if ((bits + 7) / 8 >= n + 5) return stored(in, n, out);

    assign_codes(lit_len, 286, lit_code);
    assign_codes(dist_len, 30, dist_code);
    assign_codes(pre_len, 19, pre_code);
    BitWriter bw{out};
    bw.put(1, 1); bw.put(2, 2);
    bw.put((uint32_t)(hlit - 257), 5); bw.put((uint32_t)(hdist - 1), 5); bw.put((uint32_t)(hclen - 4), 4);
    bw.flush();
    for (int i = 0; i < hclen; ++i) { bw.put(pre_len[kPreOrder[i]], 3); bw.flush(); }
    for (int i = 0; i < n_rl; ++i) {
        const int s = rl_sym[i];
        bw.put(pre_code[s], pre_len[s]);
        if (s >= 16) bw.put(rl_extra[i], s == 16 ? 2 : s == 17 ? 3 : 7);
        bw.flush();
    }
    uint32_t lit_enc[256];                                         // code | length << 16: one load per literal
    for (int s = 0; s < 256; ++s) lit_enc[s] = lit_code[s] | ((uint32_t)lit_len[s] << 16);
    for (size_t i = 0; i < n_tok; ++i) {
        const uint32_t t = tok[i];
        if (!(t & 0x80000000u)) {                                   // literals: flush once 32 bits have gathered (< 32 + 15 held)
            const uint32_t e = lit_enc[t];
            bw.put(e & 0xffff, e >> 16);
            if (bw.bc >= 32) bw.flush();
        } else {                                                    // a match adds up to 48 bits: start from fewer than 8
            if (bw.bc >= 8) bw.flush();
            const unsigned len = (t >> 16) & 0x1ff, dist = t & 0xffff;
            const int ls = kT.len_sym[len], ds = dist_symbol(dist);
            bw.put(lit_code[257 + ls], lit_len[257 + ls]);
            bw.put(len - kLenBase[ls], kLenExtra[ls]);
            bw.put(dist_code[ds], dist_len[ds]);
            bw.put(dist - kDistBase[ds], kDistExtra[ds]);
            bw.flush();
        }
    }
    bw.flush();
    bw.put(lit_code[256], lit_len[256]);
    return (size_t)(bw.finish() - out);
}

}  // namespace fqb
