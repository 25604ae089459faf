// Per-thread device logic of the paired-end resolution stage (rows a6-a9):
// bwa_aln2seq_core with the glibc drand48 stream (libbwa/bwase.c:19-46), SA row ->
// position (src/BwtMapper.cpp:765-772), bwa_approx_mapQ (libbwa/bwase.c:102-111) and
// pairing (libbwa/bwape.c:119-213, bwape.h:55-82).  Same host/device duality as
// fq_device_core.cuh.
#pragma once
#include "fq_device_core.cuh"
#include "../../include/fastquick_b200.h"

namespace fqb {

// ---- glibc drand48: X <- (0x5DEECE66D X + 0xB) mod 2^48, result X / 2^48 -----------
constexpr uint64_t kLcgA = 0x5DEECE66Dull, kLcgC = 0xBull, kLcgMask = 0xFFFFFFFFFFFFull;
FQB_HD uint64_t lcg_seed(uint32_t seed) { return ((uint64_t)seed << 16) | 0x330Eull; }
// state after n further calls (closed-form jump-ahead by repeated squaring of the affine map)
FQB_HD uint64_t lcg_advance(uint64_t x, uint64_t n) {
    uint64_t a = kLcgA, c = kLcgC, A = 1, Cc = 0;
    while (n) {
        if (n & 1) { A = (A * a) & kLcgMask; Cc = (Cc * a + c) & kLcgMask; }
        c = ((a + 1) * c) & kLcgMask;
        a = (a * a) & kLcgMask;
        n >>= 1;
    }
    return (A * x + Cc) & kLcgMask;
}
FQB_HD uint64_t lcg_next(uint64_t x) { return (kLcgA * x + kLcgC) & kLcgMask; }
FQB_HD double lcg_double(uint64_t x) { return (double)x * (1.0 / 281474976710656.0); }

constexpr uint8_t kTypeNoMatch = 0, kTypeUnique = 1, kTypeRepeat = 2, kTypeMateSW = 3;
constexpr uint8_t kSamPaired = 1, kSamProper = 2, kSamRead1 = 64, kSamRead2 = 128;

// number of leading hits that share the best score (the only ones bwa_aln2seq_core samples)
FQB_HD int count_best(const Hit *aln, int n_aln) {
    int nb = 0;
    while (nb < n_aln && aln[nb].score <= aln[0].score) ++nb;
    return nb;
}

// bwa_aln2seq_core(set_main = 1).  x = generator state BEFORE this read's first call.
// Returns the number of drand48 calls consumed.  Fields of `row` other than the ones the
// reference writes are left alone (a read whose first draw is exactly 0.0 keeps zeros).
FQB_HD uint32_t se_choose(const Hit *aln, int n_aln, uint64_t x, fqb_read_t &row) {
    if (n_aln == 0) { row.type = kTypeNoMatch; row.c1 = row.c2 = 0; return 0; }
    uint32_t calls = 0, cnt = 0;
    const int best = aln[0].score;
    int i = 0;
    for (; i < n_aln; ++i) {
        const Hit &p = aln[i];
        if (p.score > best) break;
        const uint32_t width = p.l - p.k + 1;
        x = lcg_next(x); ++calls;
#if defined(__CUDA_ARCH__)
        const bool take = __dmul_rn(lcg_double(x), (double)(width + cnt)) > (double)cnt;
#else
        const bool take = lcg_double(x) * (double)(width + cnt) > (double)cnt;
#endif
        if (take) {
            row.n_mm = p.n_mm; row.n_gapo = p.n_gapo; row.n_gape = p.n_gape; row.strand = p.a; row.score = p.score;
            x = lcg_next(x); ++calls;
#if defined(__CUDA_ARCH__)
            row.sa = p.k + (uint32_t)__dmul_rn((double)width, lcg_double(x));
#else
            row.sa = p.k + (uint32_t)((double)width * lcg_double(x));
#endif
        }
        cnt += width;
    }
    row.c1 = cnt;
    for (; i < n_aln; ++i) cnt += aln[i].l - aln[i].k + 1;
    row.c2 = cnt - row.c1;
    row.type = row.c1 > 1 ? kTypeRepeat : kTypeUnique;
    return calls;
}

FQB_HD uint32_t hit_position(const DevBwt *bwt, int a, uint32_t sa_row, int len) {
    return a ? sa_lookup(bwt[0], sa_row) : bwt[1].seq_len - (sa_lookup(bwt[1], sa_row) + (uint32_t)len);
}

FQB_HD int approx_mapq(uint32_t c1, uint32_t c2, int n_mm, int max_diff, const int32_t *g_log_n) {
    if (c1 == 0) return 23;
    if (c1 > 1) return 0;
    if (n_mm == max_diff) return 25;
    if (c2 == 0) return 37;
    int n = c2 >= 255 ? 255 : (int)c2;
    return 23 < g_log_n[n] ? 0 : 23 - g_log_n[n];
}

// hash_64 (libbwa/bwape.h:41-52)
FQB_HD uint64_t hash64(uint64_t key) {
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}

// insert-size facts a batch needs on the device: isize_info_t + the libm-dependent
// penalty (int)(-4.343*log(.5*erfc(M_SQRT1_2*|l-avg|/std))+.499), tabulated on the host
// with glibc for every insert size l in [0, high_bayesian].
struct PairParams {
    uint32_t high, high_bayesian;
    int max_isize, s_mm;
    uint32_t max_occ;
    int n_multi, N_multi;
    const int32_t *penalty;    // [high_bayesian + 1] or null when high == 0
    const int32_t *g_log_n;    // [256]
    int sw_on;                 // popt->is_sw && ii.avg >= 0 (bwa_paired_sw runs for this batch)
};

// pairing() for one pair.  arr = packed hits (pos<<32 | aln_index<<1 | end), sorted ascending.
FQB_HD void pair_resolve(fqb_read_t *p0, fqb_read_t *p1, const Hit *aln0, const Hit *aln1, const uint64_t *arr, int n_arr,
                         const PairParams &pp) {
    const uint64_t NONE = ~0ull;
    int o_n = 0, subo_n = 0;
    uint32_t max_len = (uint32_t)(p0->full_len > p1->full_len ? p0->full_len : p1->full_len);
    uint64_t last00 = NONE, last01 = NONE, last10 = NONE, last11 = NONE;   // last_pos[end][0|1]
    uint64_t o_pos0 = 0, o_pos1 = 0, subo_score = NONE, o_score = NONE;
    for (int i = 0; i < n_arr; ++i) {
        const uint64_t x = arr[i];
        const int end = (int)(x & 1);
        const Hit &hx = (end ? aln1 : aln0)[(uint32_t)x >> 1];
        if (hx.a == 1) {
            // reverse-strand hit: try the last two forward hits of the other end, newest first
            for (int t = 1; t >= 0; --t) {
                const uint64_t u = end ? (t ? last01 : last00) : (t ? last11 : last10);
                const uint64_t v = x;
                const uint32_t vlen = (uint32_t)(end ? p1->len : p0->len);
                const uint32_t l = (uint32_t)((v >> 32) + vlen - (u >> 32));
                if (u != NONE && (v >> 32) > (u >> 32) && l >= max_len &&
                    ((pp.high && l <= pp.high_bayesian) || (pp.high == 0 && l <= (uint32_t)pp.max_isize))) {
                    const Hit &hu = ((u & 1) ? aln1 : aln0)[(uint32_t)u >> 1];
                    uint64_t s = (uint64_t)(hx.score + hu.score);
                    s *= 10;
                    if (pp.high) s += (uint64_t)pp.penalty[l];
                    s = s << 32 | (uint32_t)hash64((u >> 32 << 32) | (v >> 32));
                    if (s >> 32 == o_score >> 32) ++o_n;
                    else if (s >> 32 < o_score << 32) { subo_n += o_n; o_n = 1; }   // sic (libbwa/bwape.h:65)
                    else ++subo_n;
                    if (s < o_score) {
                        subo_score = o_score; o_score = s;
                        if (u & 1) o_pos1 = u; else o_pos0 = u;
                        if (v & 1) o_pos1 = v; else o_pos0 = v;
                    } else if (s < subo_score) subo_score = s;
                }
            }
        } else if (end) { last10 = last11; last11 = x; }
        else { last00 = last01; last01 = x; }
    }
    if (o_score == NONE) return;
    int mapQ_p = 0;
    if (o_n == 1) {
        if (subo_score == NONE) mapQ_p = 29;
        else if ((subo_score >> 32) - (o_score >> 32) > (uint64_t)(pp.s_mm * 10)) mapQ_p = 23;
        else {
            int n = subo_n > 255 ? 255 : subo_n;
            mapQ_p = (int)(((subo_score >> 32) - (o_score >> 32)) / 2) - pp.g_log_n[n];
            if (mapQ_p < 0) mapQ_p = 0;
        }
    }
    const Hit &r0 = ((o_pos0 & 1) ? aln1 : aln0)[(uint32_t)o_pos0 >> 1];
    const Hit &r1 = ((o_pos1 & 1) ? aln1 : aln0)[(uint32_t)o_pos1 >> 1];
    const bool same0 = p0->pos == (uint32_t)(o_pos0 >> 32) && p0->strand == r0.a;
    const bool same1 = p1->pos == (uint32_t)(o_pos1 >> 32) && p1->strand == r1.a;
    if (same0 && same1) {
        if (p0->mapQ > 0 && p1->mapQ > 0) {
            int mq = p0->mapQ + p1->mapQ;
            if (mq > 60) mq = 60;
            p0->mapQ = p1->mapQ = (uint8_t)mq;
        } else {
            if (p0->mapQ == 0) p0->mapQ = (uint8_t)((mapQ_p + 7 < p1->mapQ) ? mapQ_p + 7 : p1->mapQ);
            if (p1->mapQ == 0) p1->mapQ = (uint8_t)((mapQ_p + 7 < p0->mapQ) ? mapQ_p + 7 : p0->mapQ);
        }
    } else if (same0) {
        p1->seQ = 0; p1->mapQ = p0->mapQ;
        if (p1->mapQ > mapQ_p) p1->mapQ = (uint8_t)mapQ_p;
    } else if (same1) {
        p0->seQ = 0; p0->mapQ = p1->mapQ;
        if (p0->mapQ > mapQ_p) p0->mapQ = (uint8_t)mapQ_p;
    } else {
        p0->seQ = p1->seQ = 0;
        mapQ_p -= 20;
        if (mapQ_p < 0) mapQ_p = 0;
        p0->mapQ = p1->mapQ = (uint8_t)mapQ_p;
    }
    // __pairing_aux2
    p0->extra_flag |= kSamProper;
    if (p0->pos != (uint32_t)(o_pos0 >> 32) || p0->strand != r0.a) {
        p0->n_mm = r0.n_mm; p0->n_gapo = r0.n_gapo; p0->n_gape = r0.n_gape; p0->strand = r0.a; p0->score = r0.score;
        p0->pos = (uint32_t)(o_pos0 >> 32);
    }
    p1->extra_flag |= kSamProper;
    if (p1->pos != (uint32_t)(o_pos1 >> 32) || p1->strand != r1.a) {
        p1->n_mm = r1.n_mm; p1->n_gapo = r1.n_gapo; p1->n_gape = r1.n_gape; p1->strand = r1.a; p1->score = r1.score;
        p1->pos = (uint32_t)(o_pos1 >> 32);
    }
}

// Head of bwa_paired_sw's per-pair loop (libbwa/bwape.c:489-510): a read filtered by the k-mer test whose mate
// was not filtered is "expanded" (un-filtered) so that it can be rescued; returns true when the pair qualifies for
// mate rescue (unpaired and one end with mapQ >= SW_MIN_MAPQ = 17).
FQB_HD bool sw_candidate(fqb_read_t *p0, fqb_read_t *p1, const PairParams &pp) {
    if (!pp.sw_on) return false;
    if (p0->filtered) { if (p1->filtered) return false; p0->filtered = 0; }
    else if (p1->filtered) p1->filtered = 0;
    return (p0->mapQ >= 17 || p1->mapQ >= 17) && (p0->extra_flag & kSamProper) == 0;
}

constexpr int kPairArrCap = 48;   // packed hit positions a thread sorts in place; larger pairs take the scratch path

// PE pass of bwa_cal_pac_pos_pe for one pair (src/BwtMapper.cpp:789-886).  Returns false when
// the pair has more hit positions than kPairArrCap (caller reroutes it).
FQB_HD bool pair_one(const DevBwt *bwt, fqb_read_t *p0, fqb_read_t *p1, const Hit *aln0, int na0, const Hit *aln1, int na1,
                     const PairParams &pp, uint64_t *arr, int arr_cap) {
    const bool m0 = p0->type == kTypeUnique || p0->type == kTypeRepeat, m1 = p1->type == kTypeUnique || p1->type == kTypeRepeat;
    if (m0 && m1) {
        uint32_t occ0 = 0, occ1 = 0;
        for (int k = 0; k < na0; ++k) occ0 += aln0[k].l - aln0[k].k + 1;
        for (int k = 0; k < na1; ++k) occ1 += aln1[k].l - aln1[k].k + 1;
        if (!(occ0 > pp.max_occ || occ1 > pp.max_occ)) {
            if ((uint64_t)occ0 + occ1 > (uint64_t)arr_cap) return false;
            int n = 0;
            for (int j = 0; j < 2; ++j) {
                const Hit *al = j ? aln1 : aln0;
                const int na = j ? na1 : na0, len = j ? p1->len : p0->len;
                for (int k = 0; k < na; ++k)
                    for (uint32_t l = al[k].k; l <= al[k].l; ++l) {
                        uint64_t x = (uint64_t)hit_position(bwt, al[k].a, l, len) << 32 | (uint32_t)k << 1 | (uint32_t)j;
                        int q = n++;                      // insertion sort: n is 2 in the common case
                        while (q > 0 && arr[q - 1] > x) { arr[q] = arr[q - 1]; --q; }
                        arr[q] = x;
                    }
            }
            pair_resolve(p0, p1, aln0, aln1, arr, n, pp);
        }
    }
    // multi-hit counts: bwa_aln2seq_core(set_main = 0) (libbwa/bwase.c:47-95, src/BwtMapper.cpp:857-885)
    if (pp.N_multi || pp.n_multi) {
        for (int j = 0; j < 2; ++j) {
            fqb_read_t *p = j ? p1 : p0, *q = j ? p0 : p1;
            if (p->type == kTypeNoMatch) continue;
            const Hit *al = j ? aln1 : aln0;
            const int na = j ? na1 : na0;
            int lim = pp.n_multi;
            if (!(p->extra_flag & kSamProper) && q->type != kTypeNoMatch)
                lim = (p->c1 + p->c2 - 1 > (uint32_t)pp.N_multi) ? pp.n_multi : pp.N_multi;
            uint32_t n_occ = 0, others = 0;
            for (int k = 0; k < na; ++k) {
                n_occ += al[k].l - al[k].k + 1;
                others += al[k].l - al[k].k + 1 - ((p->sa >= al[k].k && p->sa <= al[k].l) ? 1u : 0u);
            }
            p->n_multi = (n_occ > (uint32_t)lim + 1) ? 0 : (uint8_t)(others < (uint32_t)lim ? others : (uint32_t)lim);
        }
    }
    return true;
}

}  // namespace fqb
