// Per-lane device logic of the align hot path: FM-index rank queries, SA
// resolution, bwt_cal_width, and the bounded best-first backtracking search
// (bwt_match_gap).  Everything here is a plain per-thread function over raw
// pointers so the same code can be instantiated by the CUDA kernels
// (fq_kernels.cu) and, for CPU-side CI only, by tests/emul (logic check without a
// GPU; never linked into the product library).
//
// Reference anchors: libbwa/bwt.h:89-222 (occ), libbwa/bwt.c:69-79 (bwt_sa),
// libbwa/bwtaln.c:73-97 (bwt_cal_width), libbwa/bwtgap.c:45-264 (stack + search).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FQB_HD __host__ __device__ __forceinline__
#else
#define FQB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define FQB_POPCLL(x) __popcll(x)
#define FQB_FFSLL(x) __ffsll((long long)(x))
#define FQB_LDG4(p) __ldg(p)
#else
#define FQB_POPCLL(x) __builtin_popcountll(x)
#define FQB_FFSLL(x) __builtin_ffsll((long long)(x))
#define FQB_LDG4(p) (*(p))
#endif

#if !defined(__CUDACC__)
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

namespace fqb {

// One 32-byte rank block in ONE load: sm_100a has 256-bit global loads (LDG.E.256), so a rank query costs a single
// request to a single L2 sector (the blocks are 32-byte aligned).  Host builds read the two halves.
FQB_HD void load_block(const uint4 *p, uint4 &cnt, uint4 &bases) {
#if defined(__CUDA_ARCH__)
    unsigned long long a, b, c, d;
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    cnt.x = (uint32_t)a; cnt.y = (uint32_t)(a >> 32); cnt.z = (uint32_t)b; cnt.w = (uint32_t)(b >> 32);
    bases.x = (uint32_t)c; bases.y = (uint32_t)(c >> 32); bases.z = (uint32_t)d; bases.w = (uint32_t)(d >> 32);
#else
    cnt = p[0]; bases = p[1];
#endif
}

// ---------------------------------------------------------------------------
// FM index, re-laid for the GPU: one 32-byte block (= one L2 sector) per 64 BWT
// symbols: uint4 {cumulative A,C,G,T counts before the block} + uint4 {the 64 bases as
// two bit planes, high bits in .x:.y and low bits in .z:.w, first base in the top bit}.  The reference keeps 48-byte
// blocks per 128 symbols (libbwa/bwt.h:34,56-62); values returned are identical.
struct DevBwt {
    const uint4 *blocks;      // 2 x uint4 per block
    const uint32_t *sa;       // every 32nd row, sa[0] = 0xffffffff
    uint32_t primary, seq_len;
    uint32_t L2[5];
    uint32_t n_blocks;
};

constexpr uint32_t kNoRow = 0xffffffffu;

// dynamic index into a 4-vector without forcing it into local memory
FQB_HD uint32_t pick4(const uint32_t v[4], uint32_t c) {
    uint32_t lo = (c & 1) ? v[1] : v[0], hi = (c & 1) ? v[3] : v[2];
    return (c & 2) ? hi : lo;
}

// counts of A,C,G,T among stored symbols [0, k] (after the primary shift); bwt_occ4.
// bases = two 64-bit planes (x:y = high bits of the 64 symbols, z:w = low bits, first symbol in the top bit)
FQB_HD void occ4_block(const uint4 cnt, const uint4 bases, uint32_t n /*1..64 symbols of this block*/, uint32_t out[4]) {
    const uint64_t hi = ((uint64_t)bases.x << 32) | bases.y, lo = ((uint64_t)bases.z << 32) | bases.w;
    const uint64_t m = ~0ull << (64u - n);
    const uint64_t h = hi & m, l = lo & m;
    const uint32_t H = FQB_POPCLL(h), L = FQB_POPCLL(l), T = FQB_POPCLL(h & l);
    out[3] = cnt.w + T;
    out[2] = cnt.z + H - T;
    out[1] = cnt.y + L - T;
    out[0] = cnt.x + n - H - L + T;
}

// count of ONE symbol c among the first n symbols of a block (bwt_occ): a single masked popc
FQB_HD uint32_t occ1_block(const uint4 cnt, const uint4 bases, uint32_t n, uint32_t c) {
    const uint64_t hi = ((uint64_t)bases.x << 32) | bases.y, lo = ((uint64_t)bases.z << 32) | bases.w;
    const uint64_t m = ~0ull << (64u - n);
    const uint64_t xh = (c & 2) ? hi : ~hi, xl = (c & 1) ? lo : ~lo;
    const uint32_t base = (c & 2) ? ((c & 1) ? cnt.w : cnt.z) : ((c & 1) ? cnt.y : cnt.x);
    return base + (uint32_t)FQB_POPCLL(xh & xl & m);
}
// bwt_2occ(bwt, k, l, c): both ranks of one symbol, one block fetch when k and l share a block
FQB_HD void occ1_pair(const DevBwt &b, uint32_t k, uint32_t l, uint32_t c, uint32_t &ok, uint32_t &ol) {
    uint32_t ll = l - (l >= b.primary);
    const uint4 *pl = b.blocks + 2 * (size_t)(ll >> 6);
    uint4 cc, w;
    if (k == kNoRow) { ok = 0; load_block(pl, cc, w); ol = occ1_block(cc, w, (ll & 63) + 1, c); return; }
    uint32_t kk = k - (k >= b.primary);
    const uint4 *pk = b.blocks + 2 * (size_t)(kk >> 6);
    load_block(pk, cc, w);
    ok = occ1_block(cc, w, (kk & 63) + 1, c);
    if ((ll >> 6) != (kk >> 6)) load_block(pl, cc, w);
    ol = occ1_block(cc, w, (ll & 63) + 1, c);
}

FQB_HD void occ4(const DevBwt &b, uint32_t k, uint32_t out[4]) {
    if (k == kNoRow) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    k -= (k >= b.primary);
    const uint4 *p = b.blocks + 2 * (size_t)(k >> 6);
    uint4 c_, w_;
    load_block(p, c_, w_);
    occ4_block(c_, w_, (k & 63) + 1, out);
}

// Algorithmic occ-block touches of one bwt_2occ/bwt_2occ4(k, l) call as SURVEY.md §8(d) counts
// them for the roofline: 1 when k == l or both fall in the same 128-symbol reference block, else 2.
FQB_HD uint32_t ref_block_touches(const DevBwt &b, uint32_t k, uint32_t l) {
    if (k == l) return 1;
    if (k == kNoRow || l == kNoRow) return 2;
    uint32_t kk = k - (k >= b.primary), ll = l - (l >= b.primary);
    return (kk >> 7) == (ll >> 7) ? 1u : 2u;
}

// bwt_2occ4(bwt, k, l): both rank vectors, one block fetch when k and l share a block
FQB_HD void occ4_pair(const DevBwt &b, uint32_t k, uint32_t l, uint32_t ck[4], uint32_t cl[4]) {
    if (k == kNoRow) { ck[0] = ck[1] = ck[2] = ck[3] = 0; occ4(b, l, cl); return; }
    uint32_t kk = k - (k >= b.primary), ll = l - (l >= b.primary);
    const uint4 *p = b.blocks + 2 * (size_t)(kk >> 6);
    uint4 c, w;
    load_block(p, c, w);
    occ4_block(c, w, (kk & 63) + 1, ck);
    if ((ll >> 6) != (kk >> 6)) {
        p = b.blocks + 2 * (size_t)(ll >> 6);
        load_block(p, c, w);
    }
    occ4_block(c, w, (ll & 63) + 1, cl);
}

// bwt_sa (libbwa/bwt.c:69-79) with bwt_invPsi (libbwa/bwt.h:66-70)
FQB_HD uint32_t sa_lookup(const DevBwt &b, uint32_t k) {
    uint32_t steps = 0;
    while (k & 31u) {
        ++steps;
        if (k == b.primary) { k = 0; continue; }
        uint32_t kk = k - (k > b.primary);          // stored index of row k's symbol
        const uint4 *p = b.blocks + 2 * (size_t)(kk >> 6);
        uint4 c, w;
        load_block(p, c, w);
        uint32_t j = kk & 63;
        const uint32_t sh = 31u - (j & 31u);
        uint32_t sym = (((j < 32 ? w.x : w.y) >> sh) & 1u) << 1 | (((j < 32 ? w.z : w.w) >> sh) & 1u);
        uint32_t cnt[4];
        occ4_block(c, w, j + 1, cnt);               // occ(k, sym): same block as the symbol itself
        k = pick4(b.L2, sym) + pick4(cnt, sym);
    }
    return steps + b.sa[k >> 5];
}

// ---------------------------------------------------------------------------
// width lower bounds, packed: w in the low 27 bits, min(bid, 31) above.
// (bid is only ever compared with m <= max_diff < 31; gap_shadow stores bid = 1.)
constexpr uint32_t kWidthBits = 27;
constexpr uint32_t kWidthMask = (1u << kWidthBits) - 1;
FQB_HD uint32_t pack_width(uint32_t w, int bid) { return w | ((uint32_t)(bid > 31 ? 31 : bid) << kWidthBits); }
FQB_HD uint32_t width_w(uint32_t p) { return p & kWidthMask; }
FQB_HD int width_bid(uint32_t p) { return (int)(p >> kWidthBits); }

// read bases: `fwd` holds nt4 codes in read orientation.  The reference searches
// seq[0] = reversed read and seq[1] = reverse complement (src/BwtMapper.cpp:580-587):
// seq[a][i] = a ? comp(fwd[len-1-i]) : fwd[len-1-i].
FQB_HD uint32_t read_sym(const uint8_t *fwd, int len, int a, int i) {
    uint32_t c = fwd[len - 1 - i];
    return (a && c < 4) ? 3 - c : c;
}

// bwt_cal_width over symbols [first, first+n) of strand-a sequence; writes n+1 packed entries
FQB_HD uint32_t cal_width(const DevBwt &b, const uint8_t *fwd, int len, int a, int first, int n, uint32_t *out) {
    uint32_t k = 0, l = b.seq_len, touches = 0;
    int bid = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t c = read_sym(fwd, len, a, first + i);
        if (c < 4) {
            uint32_t ok, ol;
            occ1_pair(b, k - 1, l, c, ok, ol);
            touches += ref_block_touches(b, k - 1, l);
            k = pick4(b.L2, c) + ok + 1;
            l = pick4(b.L2, c) + ol;
        }
        if (k > l || c > 3) { k = 0; l = b.seq_len; ++bid; }
        out[i] = pack_width(l - k + 1, bid);
    }
    out[n] = pack_width(0, bid + 1);
    return touches;
}

// ---------------------------------------------------------------------------
// search options (gap_opt_t fields used by bwt_match_gap) + per-read limits
struct SearchOpt {
    int s_mm, s_gapo, s_gape;
    int mode;
    int indel_end_skip, max_del_occ, max_entries;
    int max_gapo, max_gape;
    int max_seed_diff, seed_len;
    int max_top2;
    int n_buckets;
};
constexpr int kModeGapE = 0x01, kModeLogGap = 0x04, kModeNonStop = 0x10;
constexpr int kStateM = 0, kStateI = 1, kStateD = 2;

struct Hit { uint32_t k, l; int32_t score; uint8_t n_mm, n_gapo, n_gape, a; };   // == fqb_aln_t

// One stack entry = 16 bytes.
//   x = k, y = l,
//   z = i:10 | a:1 | state:2 | n_mm:6 | n_gapo:4 | n_gape:5
//   w = last_diff_pos:10 | prev:22   (prev = next-older entry of the same score bucket)
constexpr uint32_t kNoSlot = 0x3fffffu;
FQB_HD uint32_t pack_meta(int i, int a, int state, int mm, int go, int ge) {
    return (uint32_t)i | (uint32_t)a << 10 | (uint32_t)state << 11 | (uint32_t)mm << 13 | (uint32_t)go << 19 | (uint32_t)ge << 23;
}

enum LaneStatus { kLaneRunning = 0, kLaneDone = 1, kLaneOverflow = 2, kLaneHit = 3 };
enum LaneMode { kModePop = 0, kModeExpand = 1, kModeExact = 2 };

// One search lane = one read.  step() is a flat state machine: every call does at
// most ONE stack pop and at most ONE pair of rank queries, so the 32 lanes of a warp
// stay on the same instructions whatever their reads look like.
//   HeadT:     uint16_t when the arena holds < 65535 entries (fast pass), else uint32_t
//   kFreeList: recycle popped slots through a free list (overflow pass; the fast pass
//              only recycles the most recently popped slot and otherwise bump-allocates)
//   kVar:      memory-path variants of the device build (bump arena only; the results are the same whatever the value):
//              bit 0  pop staging: every pop from memory starts an asynchronous copy (cp.async, 16 bytes) of the entry the
//                     popped one chains to -- the next entry its score bucket will hand out -- into the lane's shared-memory
//                     slot; the next pop finds it there instead of waiting for the arena in L2 / HBM.  Entries of the bump
//                     arena are never rewritten within a read, so the staged copy is valid whenever its tag matches.
//              bit 2  stack entries stored with the evict-first (streaming) policy: two thirds of them are never read back
//              (bits 1 and 3 were evict-last priorities on the rank-block and width loads: measured, no gain, removed --
//              profiles/r02_search_memory_variants.md)
template <typename HeadT, bool kFreeList, int kVar = 0>
struct SearchLane {
    static constexpr bool kStage = (kVar & 1) && !kFreeList, kStreamSt = (kVar & 4) != 0;
    // wiring
    const DevBwt *bwt;         // [2]
    const SearchOpt *opt;
    const uint8_t *fwd;
    uint32_t *w[2];            // packed widths, len+1 each (mutated by gap_shadow)
    const uint32_t *sw[2];     // packed seed widths or nullptr
    uint4 *arena;
    HeadT *heads; int head_stride;
    Hit *out; int out_cap;
    uint32_t arena_cap;
    // read state
    int len, max_diff_opt, max_diff, best_score, best_cnt, n_aln, n_entries;
    uint64_t mask0, mask1;     // non-empty score buckets
    uint32_t top, free_head;
    // current entry
    uint32_t k, l;
    int i, a, state, n_mm, n_gapo, n_gape, ldp;
    int mode;
    bool have_cur, overflow;
    uint32_t hit_x;            // interval size of the hit whose gap_shadow is pending
#if !defined(__CUDACC__)
    // host instantiation (tests/emul): the staging slot is a member and the asynchronous copy an immediate one; st_stage_stale
    // counts staged entries that differ from the arena's when they are used (the invariant kStage rests on: must stay 0)
    uint4 stage_e; uint32_t stage_tag = 0x3fffffu; uint32_t st_stage_hit = 0, st_stage_miss = 0, st_stage_stale = 0;
#endif
#if defined(__CUDA_ARCH__)
    // kStage: shared-space address of this lane's staging slot {uint4 entry, u32 tag = the entry's arena slot}: the slots are the
    // first 32 x blockDim.x bytes of the block's dynamic shared memory (recomputed where needed, so that it holds no register)
    __device__ __forceinline__ static uint32_t stage_slot() {
        extern __shared__ __align__(16) unsigned char fqb_dyn_smem[];
        return (uint32_t)__cvta_generic_to_shared(fqb_dyn_smem) + threadIdx.x * 32u;
    }
#endif
    // statistics
    uint32_t n_pops, n_occ, n_blk;
#ifdef FQB_LANE_STATS
    uint32_t st_iter, st_mempop, st_skip, st_exact, st_expand, st_push, st_hit, st_adiff, st_gapok, st_am;
    uint32_t st_x[20]; int st_run;
#define FQB_STAT(x) (++(x))
#else
#define FQB_STAT(x) ((void)0)
#endif

    FQB_HD int score3(int mm, int go, int ge) const { return mm * opt->s_mm + go * opt->s_gapo + ge * opt->s_gape; }
    FQB_HD HeadT &head(int s) { return heads[(size_t)s * head_stride]; }
    FQB_HD uint32_t *wa() const { return a ? w[1] : w[0]; }
    FQB_HD const uint32_t *swa() const { return a ? sw[1] : sw[0]; }
    FQB_HD bool bucket_set(int s) const { return s < 64 ? (mask0 >> s) & 1 : (mask1 >> (s - 64)) & 1; }
    FQB_HD int diffs_left() const { return max_diff - (n_mm + n_gapo) - ((opt->mode & kModeGapE) ? n_gape : 0); }

    FQB_HD void store_entry(uint32_t s, uint32_t x, uint32_t y, uint32_t z, uint32_t w_) {
#if defined(__CUDA_ARCH__)
        if (kStreamSt) { asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(arena + s), "r"(x), "r"(y), "r"(z), "r"(w_) : "memory"); return; }
#endif
        arena[s] = make_uint4(x, y, z, w_);
    }
    FQB_HD uint32_t alloc_slot() {
        if (kFreeList && free_head != kNoSlot) { uint32_t s = free_head; free_head = arena[s].w & kNoSlot; return s; }
        if (top < arena_cap) return top++;
        overflow = true;
        return kNoSlot;
    }
    FQB_HD void release_slot(uint32_t s) {
        if (kFreeList) { arena[s].w = free_head; free_head = s; }
    }

    // gap_push (libbwa/bwtgap.c:45-64), split in two: emit() stores one entry and chains it behind
    // `prev`; the caller publishes the last entry of a group as the new head of its score bucket.
    FQB_HD uint32_t emit(uint32_t pk, uint32_t pl, uint32_t meta, int pldp, uint32_t prev) {
        uint32_t s = alloc_slot();
        ++n_entries;
        FQB_STAT(st_push);
        if (s == kNoSlot) return prev;
        store_entry(s, pk, pl, meta, (uint32_t)pldp << 22 | prev);
        return s;
    }
    FQB_HD uint32_t bucket_head(int sc) { return bucket_set(sc) ? (uint32_t)head(sc) : kNoSlot; }
    // All children of one expansion that land in the same score bucket, pushed in order with ONE allocation
    // (bump arena) -- slot numbers and chaining are exactly those of N consecutive gap_push calls.
    template <int N>
    FQB_HD void emit_group(int sc, const bool (&v)[N], const uint32_t (&ek)[N], const uint32_t (&el)[N], const uint32_t (&em)[N], const uint32_t (&ed)[N]) {
        if (kFreeList) {
            const uint32_t old = bucket_head(sc);
            uint32_t last = old;
            for (int j = 0; j < N; ++j) if (v[j]) last = emit(ek[j], el[j], em[j], (int)ed[j], last);
            if (last != old) publish(sc, last);
            return;
        }
        int cnt = 0;
        _Pragma("unroll")
        for (int j = 0; j < N; ++j) cnt += v[j] ? 1 : 0;
        if (cnt == 0) return;
        n_entries += cnt;
#ifdef FQB_LANE_STATS
        st_push += cnt;
        if (n_aln == 0) st_x[0] += cnt;                 // pushed before the first hit
        if (n_mm + n_gapo + n_gape == 0) st_x[1] += cnt;  // children of a score-0 (root path) parent
        if (n_aln > 0 && sc > best_score + opt->s_mm) st_x[2] += cnt;   // dead on arrival (post-hit, beyond the stop score)
        st_x[3] += 1;                                    // groups
#endif
        if (top + (uint32_t)cnt > arena_cap) { overflow = true; return; }
        uint32_t prev = bucket_head(sc), s = top;
        _Pragma("unroll")
        for (int j = 0; j < N; ++j)
            if (v[j]) { store_entry(s, ek[j], el[j], em[j], ed[j] << 22 | prev); prev = s; ++s; }
        top = s;
        publish(sc, prev);
    }
    FQB_HD void publish(int sc, uint32_t last) {
        head(sc) = (HeadT)last;
        if (sc < 64) mask0 |= 1ull << sc; else mask1 |= 1ull << (sc - 64);
    }
    FQB_HD void push(int pa, int pi, uint32_t pk, uint32_t pl, int mm, int go, int ge, int st, int pldp) {
        int sc = score3(mm, go, ge);
        uint32_t s = emit(pk, pl, pack_meta(pi, pa, st, mm, go, ge), pldp, bucket_head(sc));
        if (!overflow) publish(sc, s);
    }

    // gap_pop (libbwa/bwtgap.c:66-79); the exact-match child of the previous expansion is
    // always the next entry popped, so it never leaves registers.
#ifdef FQB_LANE_STATS
    void flush_run() { if (st_run) { st_x[6]++; st_x[7 + (st_run >= 64 ? 6 : st_run >= 32 ? 5 : st_run >= 16 ? 4 : st_run >= 8 ? 3 : st_run >= 4 ? 2 : st_run >= 2 ? 1 : 0)] += st_run; } st_run = 0; }
#endif
    FQB_HD void pop() {
#ifdef FQB_LANE_STATS
        if (!have_cur) flush_run();
#endif
        --n_entries;
        ++n_pops;
        if (have_cur) { have_cur = false; return; }
        FQB_STAT(st_mempop);
        int b = mask0 ? FQB_FFSLL(mask0) - 1 : 63 + FQB_FFSLL(mask1);
        uint32_t s = (uint32_t)head(b);
        uint4 e;
#if defined(__CUDA_ARCH__)
        if (kStage) {
            uint32_t tag;
            const uint32_t stage_sa = stage_slot();
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tag) : "r"(stage_sa + 16u) : "memory");
            if (tag == s) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "r"(stage_sa) : "memory");
            else e = arena[s];
        } else
#elif !defined(__CUDACC__)
        if (kStage) {
            if (stage_tag == s) {
                e = stage_e; ++st_stage_hit;
                const uint4 now = arena[s];
                if (now.x != e.x || now.y != e.y || now.z != e.z || now.w != e.w) ++st_stage_stale;
            } else { e = arena[s]; ++st_stage_miss; }
        } else
#endif
        e = arena[s];
        uint32_t prev = e.w & kNoSlot;
        if (prev == kNoSlot) { if (b < 64) mask0 &= ~(1ull << b); else mask1 &= ~(1ull << (b - 64)); }
        else {
            head(b) = (HeadT)prev;
#if defined(__CUDA_ARCH__)
            if (kStage) {
                const uint32_t stage_sa = stage_slot();
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage_sa), "l"(arena + prev) : "memory");
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage_sa + 16u), "r"(prev) : "memory");
            }
#elif !defined(__CUDACC__)
            if (kStage) { stage_e = arena[prev]; stage_tag = prev; }
#endif
#if defined(__CUDA_ARCH__) && defined(FQB_POP_PREFETCH)
            // the next pop from this bucket follows the chain: start pulling it towards the SM now (a whole step early)
            asm volatile(FQB_POP_PREFETCH " [%0];" ::"l"(arena + prev));
#endif
        }
        release_slot(s);
        k = e.x; l = e.y;
        i = e.z & 1023; a = (e.z >> 10) & 1; state = (e.z >> 11) & 3;
        n_mm = (e.z >> 13) & 63; n_gapo = (e.z >> 19) & 15; n_gape = (e.z >> 23) & 31;
        ldp = (int)(e.w >> 22);
    }

    // bwt_match_gap prologue (libbwa/bwtgap.c:104-128); n_ambig = bases > 3 among the first len
    FQB_HD LaneStatus begin(int read_len, int read_max_diff, int n_ambig) {
        len = read_len; max_diff_opt = max_diff = read_max_diff;
        best_score = score3(max_diff_opt + 1, opt->max_gapo + 1, opt->max_gape + 1);
        best_cnt = 0; n_aln = 0; n_entries = 0;
        mask0 = mask1 = 0; top = 0; free_head = kNoSlot;
        have_cur = false; mode = kModePop; overflow = false;
        n_pops = n_occ = n_blk = 0; hit_x = 0;
#if defined(__CUDA_ARCH__)
        if (kStage) asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage_slot() + 16u), "r"(kNoSlot) : "memory");   // nothing staged for this read yet
#elif !defined(__CUDACC__)
        stage_tag = kNoSlot;
#endif
        if (n_ambig > max_diff) return kLaneDone;
        push(0, len, 0, bwt[0].seq_len, 0, 0, 0, kStateM, 0);
        // second root (strand 1) is the top of bucket 0: keep it in registers
        k = 0; l = bwt[0].seq_len; i = len; a = 1; state = kStateM; n_mm = n_gapo = n_gape = 0; ldp = 0;
        have_cur = true; ++n_entries;
        return overflow ? kLaneOverflow : kLaneRunning;
    }

    // a hit: libbwa/bwtgap.c:163-199 minus gap_shadow, which the caller runs (warp-cooperatively
    // on the GPU) when kLaneHit is returned.
    FQB_HD LaneStatus on_hit() {
        int sc = score3(n_mm, n_gapo, n_gape);
        if (n_aln == 0) {
            best_score = sc;
            int best_diff = n_mm + n_gapo + ((opt->mode & kModeGapE) ? n_gape : 0);
            if (!(opt->mode & kModeNonStop)) max_diff = (best_diff + 1 > max_diff_opt) ? max_diff_opt : best_diff + 1;
        }
        if (sc == best_score) best_cnt += (int)(l - k + 1);
        else if (best_cnt > opt->max_top2) return kLaneDone;
        if (n_gapo) {
            int n_cmp = n_aln < out_cap ? n_aln : out_cap;
            for (int j = 0; j < n_cmp; ++j)
                if (out[j].k == k && out[j].l == l) return kLaneRunning;
        }
        if (n_aln < out_cap) {
            Hit h; h.k = k; h.l = l; h.score = sc;
            h.n_mm = (uint8_t)n_mm; h.n_gapo = (uint8_t)n_gapo; h.n_gape = (uint8_t)n_gape; h.a = (uint8_t)a;
            out[n_aln] = h;
        } else overflow = true;
        ++n_aln;
        hit_x = l - k + 1;
        return overflow ? kLaneOverflow : kLaneHit;
    }

    // gap_shadow (libbwa/bwtgap.c:81-91), serial form (host emulation; the kernel has a warp-wide one)
    FQB_HD void shadow_serial() {
        uint32_t x = hit_x, maxv = bwt[1 - a].seq_len;
        uint32_t *wa_ = wa();
        int jj = 0;
        for (int p = 0; p < ldp; ++p) {
            uint32_t v = wa_[p], ww = width_w(v);
            if (ww > x) wa_[p] = v - x;
            else if (ww == x) wa_[p] = pack_width(maxv - (uint32_t)(++jj), 1);
        }
    }

    FQB_HD LaneStatus step() {
        FQB_STAT(st_iter);
        if (mode == kModePop) {
            if (n_entries == 0 || n_entries > opt->max_entries) return kLaneDone;
            pop();
            if (!(opt->mode & kModeNonStop) && score3(n_mm, n_gapo, n_gape) > best_score + opt->s_mm) return kLaneDone;
            if (diffs_left() < 0) { FQB_STAT(st_skip); return kLaneRunning; }
            if (i == 0) { FQB_STAT(st_hit); return on_hit(); }
        }
        // ---- every load this step can need, issued together (one memory round trip instead of a chain of them):
        // the two rank blocks, the width bounds of positions i-1 and i-2, their seed counterparts, and two read symbols
        const DevBwt &b = a ? bwt[0] : bwt[1];
        const bool no_k = k == 0;                       // bwt_occ4(k - 1) with k - 1 == (bwtint_t)-1
        const uint32_t kk_ = no_k ? 0 : (k - 1) - ((k - 1) >= b.primary), ll_ = l - (l >= b.primary);
        const uint4 *pk = b.blocks + 2 * (size_t)(kk_ >> 6), *pl = b.blocks + 2 * (size_t)(ll_ >> 6);
        uint4 bk_c, bk_w;
        load_block(pk, bk_c, bk_w);
        const bool same_blk = (kk_ >> 6) == (ll_ >> 6);
        uint4 bl_c = bk_c, bl_w = bk_w;
        if (!same_blk) load_block(pl, bl_c, bl_w);
        const uint32_t c_here = fwd[len - i];                           // read_sym(i - 1), before complementing
        const uint32_t c_next = i >= 2 ? fwd[len - i + 1] : 4u;         // read_sym(i - 2)
        uint32_t w_hi = 0, w_lo = 0, s_hi = 0, s_lo = 0;
        const int ii = (i - 1) - (len - opt->seed_len);
        const bool seed_chk = sw[0] && ii > 0;
        if (mode != kModeExact) {
            const uint32_t *wp = wa();
            w_hi = wp[i - 1];
            if (i >= 2) w_lo = wp[i - 2];
            if (seed_chk) { const uint32_t *sp = swa(); s_lo = sp[ii - 1]; s_hi = sp[ii]; }
        }
        const uint32_t sym_here = (a && c_here < 4) ? 3 - c_here : c_here;
        const uint32_t sym_next = (a && c_next < 4) ? 3 - c_next : c_next;

        if (mode == kModePop) {
            const int m = diffs_left();
            if (m < width_bid(w_hi)) { FQB_STAT(st_skip); return kLaneRunning; }
            if (m == 0 && (state == kStateM || (opt->mode & kModeGapE) || n_gape == opt->max_gape)) {
                if (sym_here > 3) { return kLaneRunning; }   // bwt_match_exact_alt: N never matches
                mode = kModeExact;
            } else mode = kModeExpand;
        }
        uint32_t ck[4], cl[4];
        if (no_k) ck[0] = ck[1] = ck[2] = ck[3] = 0;
        else occ4_block(bk_c, bk_w, (kk_ & 63) + 1, ck);
        occ4_block(bl_c, bl_w, (ll_ & 63) + 1, cl);
        ++n_occ;
        n_blk += 1u + (uint32_t)(no_k || (kk_ >> 7) != (ll_ >> 7));     // == ref_block_touches(b, k - 1, l)

        if (mode == kModeExact) FQB_STAT(st_exact); else FQB_STAT(st_expand);
        if (mode == kModeExact) {                   // one step of bwt_match_exact_alt (libbwa/bwt.c:102-117)
            k = b.L2[sym_here] + pick4(ck, sym_here) + 1;
            l = b.L2[sym_here] + pick4(cl, sym_here);
            --i;
            if (k > l) { mode = kModePop; return kLaneRunning; }
            if (i == 0) { mode = kModePop; return on_hit(); }
            if (sym_next > 3) { mode = kModePop; }
            return kLaneRunning;
        }

        // expansion: libbwa/bwtgap.c:201-259
        mode = kModePop;
        int m = diffs_left();
        int m_seed = opt->max_seed_diff - (n_mm + n_gapo) - ((opt->mode & kModeGapE) ? n_gape : 0);
        --i;
        uint32_t occ = l - k + 1;
        bool allow_diff = true, allow_M = true;
        if (i > 0) {
            const uint32_t wl = w_lo, wh = w_hi;
            if (width_bid(wl) > m - 1) allow_diff = false;
            else if (width_bid(wl) == m - 1 && width_bid(wh) == m - 1 && width_w(wl) == width_w(wh)) allow_M = false;
            if (seed_chk) {
                const uint32_t sl = s_lo, sh = s_hi;
                if (width_bid(sl) > m_seed - 1) allow_diff = false;
                else if (width_bid(sl) == m_seed - 1 && width_bid(sh) == m_seed - 1 && width_w(sl) == width_w(sh)) allow_M = false;
            }
        }
        if (allow_diff) FQB_STAT(st_adiff);
#ifdef FQB_LANE_STATS
        if (k == l) st_x[4]++;                           // expansions on a one-row interval
        if (k == l && allow_diff) st_x[5]++;
        if (!allow_diff) { ++st_run; } else flush_run();
#endif
        if (allow_diff && allow_M) FQB_STAT(st_am);
        int gaps = n_gapo + n_gape;
        if (opt->mode & kModeLogGap) { int v = gaps, c = 0; while (v >>= 1) ++c; gaps = c / 2 + 1; }
        // child intervals for the four symbols
        uint32_t kk[4], ll[4];
        _Pragma("unroll")
        for (int j = 0; j < 4; ++j) { kk[j] = b.L2[j] + ck[j] + 1; ll[j] = b.L2[j] + cl[j]; }
        // gap children share one score bucket: insertion first, then deletions j = 0..3
        bool do_i = false, do_d = false;
        int go2 = n_gapo, ge2 = n_gape;
        if (allow_diff && i >= opt->indel_end_skip + gaps && len - i >= opt->indel_end_skip + gaps) {
            if (state == kStateM) { if (n_gapo < opt->max_gapo) { do_i = do_d = true; go2 = n_gapo + 1; } }
            else if (state == kStateI) { if (n_gape < opt->max_gape) { do_i = true; ge2 = n_gape + 1; } }
            else if (n_gape < opt->max_gape && (n_gape + n_gapo < max_diff || occ < (uint32_t)opt->max_del_occ)) { do_d = true; ge2 = n_gape + 1; }
        }
        if (do_i || do_d) {
            FQB_STAT(st_gapok);
            const uint32_t meta_i = pack_meta(i, a, kStateI, n_mm, go2, ge2), meta_d = pack_meta(i + 1, a, kStateD, n_mm, go2, ge2);
            const bool v[5] = {do_i, do_d && kk[0] <= ll[0], do_d && kk[1] <= ll[1], do_d && kk[2] <= ll[2], do_d && kk[3] <= ll[3]};
            const uint32_t ek[5] = {k, kk[0], kk[1], kk[2], kk[3]}, el[5] = {l, ll[0], ll[1], ll[2], ll[3]};
            const uint32_t em[5] = {meta_i, meta_d, meta_d, meta_d, meta_d};
            const uint32_t ed[5] = {(uint32_t)i, (uint32_t)i + 1, (uint32_t)i + 1, (uint32_t)i + 1, (uint32_t)i + 1};
            emit_group<5>(score3(n_mm, go2, ge2), v, ek, el, em, ed);
        }
        // mismatch children (one bucket) in the order (s+1)&3, (s+2)&3, (s+3)&3 [, s&3 when s is N]; the exact child is kept
        const uint32_t s = sym_here;
        bool keep = false;
        uint32_t nk = 0, nl = 0;
        {
            uint32_t ck4[4], cl4[4];
            _Pragma("unroll")
            for (int j = 1; j <= 4; ++j) { const uint32_t c = (s + j) & 3; ck4[j - 1] = pick4(kk, c); cl4[j - 1] = pick4(ll, c); }
            if (allow_diff && allow_M) {
                const uint32_t meta_m = pack_meta(i, a, kStateM, n_mm + 1, n_gapo, n_gape);
                const bool v[4] = {ck4[0] <= cl4[0], ck4[1] <= cl4[1], ck4[2] <= cl4[2], s > 3 && ck4[3] <= cl4[3]};
                const uint32_t em[4] = {meta_m, meta_m, meta_m, meta_m}, ed[4] = {(uint32_t)i, (uint32_t)i, (uint32_t)i, (uint32_t)i};
                emit_group<4>(score3(n_mm + 1, n_gapo, n_gape), v, ck4, cl4, em, ed);
            }
            if (s < 4 && ck4[3] <= cl4[3]) { keep = true; nk = ck4[3]; nl = cl4[3]; }
        }
        if (keep) {      // exact child: pushed last into the parent's own bucket, hence popped next;
            k = nk; l = nl; state = kStateM; have_cur = true; ++n_entries;   // inherits last_diff_pos
        }
       
        return overflow ? kLaneOverflow : kLaneRunning;
    }
};

}  // namespace fqb
