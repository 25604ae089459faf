// Per-thread device logic of the statistics rows (a12, a13):
// StatCollector::AddAlignment / ProcessPairStatus (src/StatCollector.cpp:950-1101, 623-948) as a
// per-pair classifier, and AddSingleAlignment's per-base walk (424-621, 342-422, 310-340).
#pragma once
#include "fq_device_dp.cuh"

namespace fqb {

struct ContigDev { int64_t offset; int32_t len; int32_t gstart; uint8_t is_xy; uint8_t pad[7]; };   // gstart = genome coordinate of offset

// per-pair outcome; the host formats the InsertSizeTable line from it (+ names, contig names, rows)
enum PairStatus : uint8_t { kStNone = 0, kStPropPair, kStPartialPair, kStNotPair, kStLowQual, kStFwdOnly, kStRevOnly };
struct PairStat {
    int32_t max_insert, max_insert2, actual_insert;
    int32_t seqid[2];
    uint16_t flag[2];        // SAM flags as ProcessPairStatus prints them
    uint8_t status;          // PairStatus; kStNone = no InsertSizeTable line
    uint8_t line_kind;       // 0 none, 1 first-only, 2 second-only, 3 both
    uint8_t retained;        // AddAlignment's return value (FSC.TotalRetained)
    uint8_t low_mapq;        // total_add_failed increment (FSC.TotalMAPQ)
    uint8_t both_filtered, both_unmapped;
    uint8_t add[2];          // AddSingleAlignment returned true for end e
    uint8_t demoted[2];      // bridge check demoted end e to NO_MATCH
};

// pos_end (libbwa/bwase.c:420-433)
FQB_HD int64_t pos_end(const fqb_read_t &p) {
    if (p.has_cigar) {
        int64_t x = p.pos;
        for (int j = 0; j < p.n_cigar; ++j) { int op = p.cigar[j] >> 14; if (op == 0 || op == 2) x += p.cigar[j] & 0x3fff; }
        return x;
    }
    return (int64_t)p.pos + p.len;
}
// bns_coor_pac2real's contig search (libbwa/bntseq.c:268-283)
FQB_HD int find_contig(const ContigDev *c, int n, int64_t pac) {
    int left = 0, mid = 0, right = n;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pac >= c[mid].offset) {
            if (mid == n - 1) break;
            if (pac < c[mid + 1].offset) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}
FQB_HD bool partial_align(const fqb_read_t &p) {       // IsPartialAlign: any soft clip
    if (!p.has_cigar) return false;
    for (int k = 0; k < p.n_cigar; ++k) if ((p.cigar[k] >> 14) == kOpS) return true;
    return false;
}

constexpr int kInsertLimit = 4096;

// per-contig counters of contigStatusTable: [0] overlapped [1] fully included [2] pair overlapped [3] fully included paired
struct StatAccum {
    uint32_t *contig_ctr;        // [n_contigs][4]
    uint32_t *contig_first;      // [n_contigs] first pair index that touched the contig (insertion order of the unordered_map)
    unsigned long long *isize_dist;   // [4096]
    unsigned long long *scalars;      // [0] NumPCRDup/2 candidates are resolved by the key table; [1] NumPairReads, [2] isize out of range
    unsigned long long *dup_keys; uint32_t dup_cap; unsigned long long *dup_count;   // open-addressing set of start<<32|end
};

// order = 2*pair + (0 when reached through the second read's contig name, 1 through the first read's): the
// reference inserts contigStatusTable[qname] before [pname] within one AddAlignment call
FQB_HD void touch_contig(const StatAccum &A, int seqid, uint32_t pair, int which) {
#if defined(__CUDA_ARCH__)
    atomicAdd(A.contig_ctr + 4 * seqid + which, 1u);
    atomicMin(A.contig_first + seqid, pair);
#else
    A.contig_ctr[4 * seqid + which] += 1;
    if (pair < A.contig_first[seqid]) A.contig_first[seqid] = pair;
#endif
}
// counter += v.  On the device the lanes of a warp that hit the same counter with the same increment elect one leader
// (most pairs bump the same few scalars and insert-size bins, so this removes almost all same-address atomics).
FQB_HD void bump64(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, (unsigned long long)(uintptr_t)p) & __match_any_sync(act, v);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(p, v * (unsigned long long)__popc(peers));
#else
    *p += v;
#endif
}
// duplicateTable.insert("seqid:start:end"): returns true if the key was already present
FQB_HD bool dup_insert(const StatAccum &A, uint32_t start, uint32_t end) {
    const unsigned long long key = ((unsigned long long)start << 32 | end) + 1;      // 0 = empty slot
    uint32_t h = (uint32_t)(hash64(key) % A.dup_cap);
    for (uint32_t probes = 0; probes < A.dup_cap; ++probes) {
#if defined(__CUDA_ARCH__)
        unsigned long long old = atomicCAS(A.dup_keys + h, 0ull, key);
#else
        unsigned long long old = A.dup_keys[h];
        if (old == 0) A.dup_keys[h] = key;
#endif
        if (old == 0) { bump64(A.dup_count, 1); return false; }
        if (old == key) return true;
        h = h + 1 == A.dup_cap ? 0 : h + 1;
    }
    bump64(A.scalars + 2, 1);          // table full: reported by fqb_stats_finish as a limit error, never silently dropped
    return false;
}

struct ClipInfo { int cl_left, cl_right; };
FQB_HD ClipInfo clips_of(const fqb_read_t &p) {
    ClipInfo c; c.cl_left = c.cl_right = 0;
    if (p.has_cigar) {
        if ((p.cigar[0] >> 14) == kOpS) c.cl_left = p.cigar[0] & 0x3fff;
        if ((p.cigar[p.n_cigar - 1] >> 14) == kOpS) c.cl_right = p.cigar[p.n_cigar - 1] & 0x3fff;
    }
    return c;
}
FQB_HD uint16_t sam_flag(const fqb_read_t &p) { return (uint16_t)(p.extra_flag | (p.type == kTypeNoMatch ? 4 : 0) | (p.strand ? 16 : 0)); }

// ProcessPairStatus (src/StatCollector.cpp:623-948).  type: 0 FirstOnly, 1 Both, 2 SecondOnly.  Returns 0 or 2.
FQB_HD int pair_status(const ContigDev *ctg, int n_ctg, const fqb_read_t &p, const fqb_read_t &q, int type, const StatAccum &A, PairStat &o) {
    int maxInsert = -1, maxInsert2 = -1;
    o.flag[0] = sam_flag(p); o.flag[1] = sam_flag(q);
    if (q.full_len == 0) o.flag[1] = 0;          // single-end input: the reference passes q = NULL (flag2 stays 0)
    if (p.full_len == 0) o.flag[0] = 0;
    o.actual_insert = -1;
    if (type != 1) {                                   // single end: e = the aligned read
        const fqb_read_t &e = type == 0 ? p : q;
        const int sid = find_contig(ctg, n_ctg, e.pos);
        o.seqid[type == 0 ? 0 : 1] = sid;
        o.line_kind = type == 0 ? 1 : 2;
        if (e.mapQ > 0) {
            const ClipInfo c = clips_of(e);
            const uint32_t left = e.pos - (uint32_t)c.cl_left;             // bwtint_t arithmetic
            if (e.strand) {
                if (ctg[sid].offset + ctg[sid].len >= (int64_t)(uint32_t)(left + (uint32_t)e.len))
                    maxInsert2 = (int)((int64_t)(uint32_t)(left + (uint32_t)e.len) - ctg[sid].offset);
                else { o.line_kind = 0; return 2; }
                o.status = kStRevOnly;
            } else {
                if ((int64_t)left >= ctg[sid].offset) maxInsert = (int)(ctg[sid].offset + ctg[sid].len - (int64_t)left);
                else { o.line_kind = 0; return 2; }
                o.status = kStFwdOnly;
            }
            o.max_insert = maxInsert; o.max_insert2 = maxInsert2;
            return 0;
        }
        o.status = kStLowQual; o.max_insert = -1; o.max_insert2 = -1;
        return 2;
    }
    const int sp = find_contig(ctg, n_ctg, p.pos), sq = find_contig(ctg, n_ctg, q.pos);
    o.seqid[0] = sp; o.seqid[1] = sq;
    o.line_kind = 3;
    const ClipInfo cp = clips_of(p), cq = clips_of(q);
    const uint32_t pl = p.pos - (uint32_t)cp.cl_left, ql = q.pos - (uint32_t)cq.cl_left;
    const bool fr = !p.strand && q.strand && p.pos < q.pos, rf = !q.strand && p.strand && q.pos < p.pos;
    if (fr) {
        maxInsert = ((int64_t)pl >= ctg[sp].offset) ? (int)(ctg[sp].offset + ctg[sp].len - (int64_t)pl) : -1;
        maxInsert2 = (ctg[sq].offset + ctg[sq].len >= (int64_t)(uint32_t)(ql + (uint32_t)q.len)) ? (int)((int64_t)(uint32_t)(ql + (uint32_t)q.len) - ctg[sq].offset) : -1;
    } else if (rf) {
        maxInsert = ((int64_t)ql >= ctg[sq].offset) ? (int)(ctg[sq].offset + ctg[sq].len - (int64_t)ql) : -1;
        maxInsert2 = (ctg[sp].offset + ctg[sp].len >= (int64_t)(uint32_t)(pl + (uint32_t)p.len)) ? (int)((int64_t)(uint32_t)(pl + (uint32_t)p.len) - ctg[sp].offset) : -1;
    } else {
        o.status = kStNotPair; o.max_insert = -1; o.max_insert2 = -1;
        return 0;
    }
    if (maxInsert >= kInsertLimit) maxInsert = kInsertLimit - 1;
    if (maxInsert2 >= kInsertLimit) maxInsert2 = kInsertLimit - 1;
    o.max_insert = maxInsert; o.max_insert2 = maxInsert2;
    if (sp != sq) {
        bump64(A.isize_dist + 0, 1);
        o.status = kStNotPair;
        return 0;
    }
    if (p.mapQ > 0 && q.mapQ > 0) {
        bool noClip = false;
        int start, end;
        if (fr) { start = (int)pl; end = (int)(ql + (uint32_t)q.len); noClip = cp.cl_left == 0 && cq.cl_right == 0; }
        else { start = (int)ql; end = (int)(pl + (uint32_t)p.len); noClip = cq.cl_left == 0 && cp.cl_right == 0; }
        const int actual = end - start;
        const bool prop = maxInsert != -1 && maxInsert2 != -1;
        o.status = prop ? kStPropPair : kStPartialPair;
        o.actual_insert = actual;
        if (actual >= 0 && actual < kInsertLimit) bump64(A.isize_dist + actual, 1);
        else bump64(A.scalars + 2, 1);                 // the reference indexes out of bounds here (src/StatCollector.cpp:913)
        if (prop && noClip) {
            if (dup_insert(A, (uint32_t)start, (uint32_t)end)) bump64(A.scalars + 0, 2);
            bump64(A.scalars + 1, 2);
        }
        return 0;
    }
    o.status = kStLowQual;
    return 2;
}

// AddSingleAlignment's accept test for FASTQuick's own alignments (contig names carry ':')
FQB_HD bool single_ok(const fqb_read_t &p) { return !(p.type == kTypeNoMatch || p.mapQ < 20); }

// StatCollector::AddAlignment (src/StatCollector.cpp:950-1101) + the per-pair counters of
// BwtMapper::PairEndMapper's main-thread loop (src/BwtMapper.cpp:2053-2085).  p and q may be demoted.
FQB_HD void classify_pair(const ContigDev *ctg, int n_ctg, fqb_read_t &p, fqb_read_t &q, uint32_t pair_index, int cal_dup, const StatAccum &A, PairStat &o) {
    const uint32_t pair = 2 * pair_index, pair_p = 2 * pair_index + 1;
    o.max_insert = o.max_insert2 = o.actual_insert = -1;
    o.seqid[0] = o.seqid[1] = -1; o.flag[0] = o.flag[1] = 0;
    o.status = kStNone; o.line_kind = 0; o.retained = 0; o.low_mapq = 0;
    o.both_filtered = o.both_unmapped = 0; o.add[0] = o.add[1] = 0; o.demoted[0] = o.demoted[1] = 0;
    if (p.filtered && q.filtered) { o.both_filtered = 1; return; }
    if (p.type == kTypeNoMatch && q.type == kTypeNoMatch) { o.both_unmapped = 1; return; }
    int seqid = 0, seqid2 = 0;
    if (p.type != kTypeNoMatch) {
        const int64_t j = pos_end(p) - p.pos;
        seqid = find_contig(ctg, n_ctg, p.pos);
        if ((int64_t)p.pos + j - ctg[seqid].offset > ctg[seqid].len) { p.type = kTypeNoMatch; o.demoted[0] = 1; }
    }
    if (q.type != kTypeNoMatch) {
        const int64_t j = pos_end(q) - q.pos;
        seqid2 = find_contig(ctg, n_ctg, q.pos);
        if ((int64_t)q.pos + j - ctg[seqid2].offset > ctg[seqid2].len) { q.type = kTypeNoMatch; o.demoted[1] = 1; }
    }
    const bool q_xy = ctg[seqid2].is_xy, p_xy = ctg[seqid].is_xy;
    if (p.type == kTypeNoMatch) {
        if (single_ok(q)) {
            o.add[1] = 1;
            if (q_xy) { touch_contig(A, seqid2, pair, 0); if (!partial_align(q)) touch_contig(A, seqid2, pair, 1); }
            pair_status(ctg, n_ctg, p, q, 2, A, o);
            o.low_mapq = 1; o.retained = 1;
            return;
        }
        o.low_mapq = 2; o.retained = 0;
        return;
    }
    if (q.type == kTypeNoMatch) {
        if (single_ok(p)) {
            o.add[0] = 1;
            if (p_xy) { touch_contig(A, seqid, pair_p, 0); if (!partial_align(p)) touch_contig(A, seqid, pair_p, 1); }
            pair_status(ctg, n_ctg, p, q, 0, A, o);
            o.low_mapq = 1; o.retained = 1;
            return;
        }
        o.low_mapq = 2; o.retained = 0;
        return;
    }
    // both ends aligned; the X/Y bookkeeping is keyed on the SECOND read's contig name (qname)
    const bool same = seqid == seqid2;      // pname == qname
    if (partial_align(p)) {
        if (q_xy) {
            touch_contig(A, seqid2, pair, 0);
            if (!partial_align(q)) touch_contig(A, seqid2, pair, 1);
            if (same) touch_contig(A, seqid2, pair, 2);
            touch_contig(A, seqid, pair_p, 0);
        }
    } else {
        if (q_xy) {
            touch_contig(A, seqid2, pair, 0);
            if (partial_align(q)) { if (same) touch_contig(A, seqid2, pair, 2); }
            else { touch_contig(A, seqid2, pair, 1); if (same) { touch_contig(A, seqid2, pair, 2); touch_contig(A, seqid2, pair, 3); } }
            touch_contig(A, seqid, pair_p, 0);
            touch_contig(A, seqid, pair_p, 1);
        }
    }
    const int rc = pair_status(ctg, n_ctg, p, q, 1, A, o);
    if (rc != 1 || cal_dup) {
        o.add[0] = single_ok(p); o.add[1] = single_ok(q);
        o.retained = (uint8_t)(o.add[0] + o.add[1]);
        o.low_mapq = (uint8_t)(2 - o.retained);
        return;
    }
    o.low_mapq = 2; o.retained = 0;
}

// ---- per-base walk -------------------------------------------------------------------------
// side tables over pac coordinates, built on the host from .SelectedSite.vcf/.gc/.dbSNP.subset.vcf and the contig names
constexpr uint32_t kSiteNone = 0x3fffffffu, kSiteMask = 0x3fffffffu, kSiteDbsnp = 0x80000000u, kSiteMarker = 0x40000000u;

struct PileupTuple { uint32_t marker; uint32_t key_hi; uint32_t key_lo; uint8_t base, qual, mapq, strand; int32_t cycle; };
// key = (global pair index, end, offset on read): arrival order of UpdateInfoVecAtMarker

struct BaseTables {
    const uint32_t *site;          // [l_pac]: site id | flags
    const int32_t *marker;         // [l_pac]: marker index (VCF order) or -1 -- only read where kSiteMarker is set
    uint32_t *depth, *q20, *q30;   // [n_sites]
    unsigned long long *emp;       // [4][256]: EmpRepDist, misEmpRepDist, EmpCycleDist, misEmpCycleDist
    PileupTuple *tuples; uint32_t *n_tuples; uint32_t tuple_cap;
};

}  // namespace fqb
