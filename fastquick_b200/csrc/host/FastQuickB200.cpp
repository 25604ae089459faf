#include "FastQuickB200.h"
#include <random>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/time.h>
#include <thread>
#include <zlib.h>

namespace fqb200 {

static double realtime() { struct timeval tp; gettimeofday(&tp, nullptr); return tp.tv_sec + tp.tv_usec * 1e-6; }
void notice(const char *fmt, ...) { va_list ap; va_start(ap, fmt); fprintf(stderr, "NOTICE - "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n"); va_end(ap); }
void warning(const char *fmt, ...) { va_list ap; va_start(ap, fmt); fprintf(stderr, "\aWARNING - \n"); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n"); va_end(ap); }
void error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); fprintf(stderr, "\nFATAL ERROR - \n"); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n"); va_end(ap); exit(EXIT_FAILURE); }

bool BwtIndexer::LoadIndex(std::string &NewRef) {
    IndexPrefix = NewRef;
    std::ifstream par(NewRef + ".param");
    if (!par) error("Open %s failed!", (NewRef + ".param").c_str());
    std::string k;
    par >> k >> RefPath;
    for (const char *ext : {".bwt", ".rbwt", ".sa", ".rsa", ".pac", ".ann", ".amb"})       // .rollhash is rebuilt on the device
        if (!std::ifstream(NewRef + ext)) error("Open %s failed!", (NewRef + ext).c_str());
    return true;
}

int StatCollector::RestoreVcfSites(const std::string &RefPath, const gap_opt_t *) {
    if (fqb_stats_open(h_, RefPath.c_str()) != FQB_OK) error("%s", fqb_last_error());
    return 0;
}
int StatCollector::ProcessCore(const std::string &statPrefix, const gap_opt_t *) {
    if (fqb_stats_finish(h_, statPrefix.c_str()) != FQB_OK) error("%s", fqb_last_error());
    return 0;
}

// ---- serial FASTQ reader, used only with --frac_samp < 1 (the draws follow the records in file order); every other run
// reads through the parallel feeder of the library (fqb_feeder_*, fq_feeder.cpp).  4-line records from gz, into fixed-stride
// pinned batches (kseq_read3_fpc's contract, libbwa/kseq.h:327-370); lines are located in place in the inflate buffer.
struct FastqReader {
    gzFile fp = nullptr;
    std::vector<char> buf; size_t pos = 0, end = 0;
    bool eof = false;
    bool open(const std::string &p) {
        fp = gzopen(p.c_str(), "rb");
        if (fp) { gzbuffer(fp, 1 << 20); buf.resize((size_t)8 << 20); pos = end = 0; eof = false; }
        return fp != nullptr;
    }
    void close() { if (fp) gzclose(fp); fp = nullptr; }
    // next line as [b, e) inside buf (without the terminator); false at end of file
    bool line(const char *&b, const char *&e) {
        for (;;) {
            const char *s = buf.data() + pos;
            const char *nl = pos < end ? (const char *)memchr(s, '\n', end - pos) : nullptr;
            if (nl) { b = s; e = nl; pos = (size_t)(nl - buf.data()) + 1; if (e > b && e[-1] == '\r') --e; return true; }
            if (eof) {
                if (pos == end) return false;
                b = s; e = buf.data() + end; pos = end; if (e > b && e[-1] == '\r') --e; return true;      // last line without a newline
            }
            // refill: keep the partial line at the front
            const size_t keep = end - pos;
            if (keep && pos) memmove(buf.data(), buf.data() + pos, keep);
            pos = 0; end = keep;
            if (end == buf.size()) buf.resize(buf.size() * 2);
            const int n = gzread(fp, buf.data() + end, (unsigned)(buf.size() - end));
            if (n <= 0) eof = true; else end += (size_t)n;
        }
    }
    // returns the number of records kept (<= n_max).  frac < 1: --frac_samp, the reference's per-record draw from a Mersenne
    // twister re-seeded with the IO round of the batch (src/BwtMapper.cpp:483-505; VerifyBamID/Random.cpp:155-198), so both
    // files of a pair drop the same records
    int fill(int n_max, int stride, uint8_t *bases, uint8_t *quals, int32_t *lens, char *names, int name_stride, double frac = 1.0, uint32_t seed = 0) {
        std::mt19937 mt(seed);
        std::string nm;
        int n = 0;
        while (n < n_max) {
            const double rand_num = ((double)mt() + 0.5) * (1.0 / 4294967296.0);
            const char *hb, *he, *sb, *se, *pb, *pe, *qb, *qe;
            if (!line(hb, he)) break;
            if (hb == he) continue;
            nm.assign(hb + 1, he);                     // the header may move when the buffer refills: copy it first
            if (!line(sb, se)) error("truncated FASTQ record");
            const size_t sl = (size_t)(se - sb);
            if ((int)sl > stride) error("read longer than %d bases: not supported", stride);
            const bool keep = !(rand_num > frac);
            if (keep) {
                memcpy(bases + (size_t)n * stride, sb, sl);
                memset(bases + (size_t)n * stride + sl, 'N', (size_t)stride - sl);
            }
            if (!line(pb, pe) || !line(qb, qe)) error("truncated FASTQ record");
            if ((size_t)(qe - qb) != sl) error("sequence and quality lengths differ in a FASTQ record");
            if (!keep) continue;
            memcpy(quals + (size_t)n * stride, qb, sl);
            memset(quals + (size_t)n * stride + sl, '!', (size_t)stride - sl);
            lens[n] = (int32_t)sl;
            const size_t l = nm.find_first_of(" \t");
            if (l != std::string::npos) nm.resize(l);
            if (nm.size() > 2 && nm[nm.size() - 2] == '/' && (nm.back() == '1' || nm.back() == '2')) nm.resize(nm.size() - 2);
            char *dst = names + (size_t)n * name_stride;
            const size_t nl = std::min(nm.size(), (size_t)name_stride - 1);
            memcpy(dst, nm.data(), nl);
            memset(dst + nl, 0, (size_t)name_stride - nl);
            ++n;
        }
        return n;
    }
};

BwtMapper::BwtMapper(BwtIndexer &BwtIndex, const std::string &FQList, const std::string &Fastq_1, const std::string &Fastq_2,
                     const std::string &Prefix, const std::string &RefPath, const pe_opt_t *popt, gap_opt_t *opt,
                     const std::string &targetRegionPath, const std::vector<int> &devices) : prefix_(Prefix) {
    if (devices.empty()) error("no device given");

    fqb_gap_opt_t g; fqb_gap_opt_default(&g);
    g.s_mm = opt->s_mm; g.s_gapo = opt->s_gapo; g.s_gape = opt->s_gape; g.mode = opt->mode;
    g.indel_end_skip = opt->indel_end_skip; g.max_del_occ = opt->max_del_occ; g.max_entries = opt->max_entries;
    g.fnr = opt->fnr; g.max_diff = opt->max_diff; g.max_gapo = opt->max_gapo; g.max_gape = opt->max_gape;
    g.max_seed_diff = opt->max_seed_diff; g.seed_len = opt->seed_len; g.max_top2 = opt->max_top2; g.trim_qual = opt->trim_qual;
    g.flank_len = opt->flank_len; g.flank_long_len = opt->flank_long_len; g.read_len = opt->read_len;
    g.kmer_thresh = BwtIndex.RollParam.thresh; g.is_il13 = (opt->mode & 0x200) ? 1 : 0;
    fqb_pe_opt_t p; fqb_pe_opt_default(&p);
    p.max_isize = popt->max_isize; p.force_isize = popt->force_isize; p.max_occ = popt->max_occ; p.n_multi = popt->n_multi;
    p.N_multi = popt->N_multi; p.type = popt->type; p.is_sw = popt->is_sw; p.ap_prior = popt->ap_prior;
    double t_tmp = realtime();
    hs_.assign(devices.size(), nullptr);
    {   // one engine per device, created side by side (each builds its k-mer tables on its own GPU)
        std::vector<std::thread> th;
        std::vector<std::string> errs(devices.size());
        for (size_t r = 0; r < devices.size(); ++r)
            th.emplace_back([&, r]() { if (fqb_create(RefPath.c_str(), &g, &p, devices[r], &hs_[r]) != FQB_OK) errs[r] = fqb_last_error(); });
        for (auto &t : th) t.join();
        for (auto &e : errs) if (!e.empty()) error("%s", e.c_str());
    }
    h_ = hs_[0];
    notice("Index on the device (FM index, SA, pac, k-mer tables)...%f sec", realtime() - t_tmp);
    collector.Attach(h_);
    for (fqb_handle *h : hs_)
        if (targetRegionPath != "Empty" && fqb_stats_set_target_region(h, targetRegionPath.c_str()) != FQB_OK) error("%s", fqb_last_error());
    t_tmp = realtime();
    collector.RestoreVcfSites(RefPath, opt);
    for (size_t r = 1; r < hs_.size(); ++r) if (fqb_stats_open(hs_[r], RefPath.c_str()) != FQB_OK) error("%s", fqb_last_error());
    notice("Restore Variant Site Info...%f sec", realtime() - t_tmp);
    if (hs_.size() > 1 && fqb_comm_init_local(hs_.data(), (int)hs_.size()) != FQB_OK) error("%s", fqb_last_error());
    bam_out_ = opt->out_bam != 0;
    if (bam_out_ && fqb_bam_open(h_, (Prefix + ".bam").c_str(), opt->RG.c_str()) != FQB_OK) error("%s", fqb_last_error());   // SetSamFileHeader + writeHeader
    for (size_t r = 1; bam_out_ && r < hs_.size(); ++r) if (fqb_bam_attach(hs_[r], h_) != FQB_OK) error("%s", fqb_last_error());
    auto run_pair = [&](const std::string &f1, const std::string &f2) {
        notice("Processing Pair End mapping\t%s\t%s", f1.c_str(), f2.c_str());
        double t0 = realtime();
        FileStatCollector FSC(f1.c_str(), f2.c_str());
        if (hs_.size() > 1) PairEndMapperSharded(f1, f2, opt, FSC); else PairEndMapper(f1, f2, opt, FSC);
        notice("Processed Pair End mapping in %f sec", realtime() - t0);
    };
    auto run_single = [&](const std::string &f1) {
        notice("Processing Single End mapping\t%s\n", f1.c_str());
        double t0 = realtime();
        FileStatCollector FSC(f1.c_str());
        if (hs_.size() > 1) error("single-end input runs on one device (--device)");
        SingleEndMapper(f1, opt, FSC);
        notice("Processed Single End mapping in %f sec", realtime() - t0);
    };
    if (FQList != "Empty") {
        notice("Open Fastq List ...");
        std::ifstream fin(FQList);
        if (!fin.is_open()) error("Open file %s failed", FQList.c_str());
        std::string line;
        while (std::getline(fin, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::string a, b;
            std::stringstream ss(line);
            ss >> a >> b;
            if (b.empty()) run_single(a); else run_pair(a, b);
        }
    } else if (Fastq_2 != "Empty") run_pair(Fastq_1, Fastq_2);
    else run_single(Fastq_1);
    if (bam_out_ && fqb_bam_close(h_) != FQB_OK) error("%s", fqb_last_error());
    double t1 = realtime();
    if (hs_.size() > 1) {
        // end of a sharded run: NCCL reduce of the accumulators onto rank 0 + exact-size sends of the pile-up entries and
        // duplicate keys (one host thread per device: the collectives of one communicator must be entered side by side),
        // then the ranks' InsertSizeTable batches are spliced back into file order
        std::vector<std::thread> th;
        std::vector<std::string> errs(hs_.size());
        for (size_t r = 0; r < hs_.size(); ++r)
            th.emplace_back([&, r]() {
                if (fqb_comm_merge_stats(hs_[r], nullptr) != FQB_OK) { errs[r] = fqb_last_error(); return; }
                if (r > 0 && fqb_stats_close_table(hs_[r]) != FQB_OK) errs[r] = fqb_last_error();
            });
        for (auto &t : th) t.join();
        for (auto &e : errs) if (!e.empty()) error("%s", e.c_str());
        std::vector<std::string> shard(hs_.size() - 1);
        std::vector<const char *> ptrs;
        for (size_t r = 1; r < hs_.size(); ++r) { shard[r - 1] = Prefix + ".shard" + std::to_string(r); ptrs.push_back(shard[r - 1].c_str()); }
        if (fqb_stats_merge_tables(h_, ptrs.data(), (int32_t)ptrs.size()) != FQB_OK) error("%s", fqb_last_error());
        for (auto &sp : shard) { remove((sp + ".InsertSizeTable").c_str()); remove((sp + ".InsertSizeTable.idx").c_str()); }
        remove((Prefix + ".InsertSizeTable.idx").c_str());
    }
    collector.ProcessCore(Prefix, opt);
    notice("Calculate distributions... %f sec", realtime() - t1);
}

BwtMapper::~BwtMapper() { for (fqb_handle *h : hs_) fqb_destroy(h); }

bool BwtMapper::PairEndMapper(const std::string &fq1, const std::string &fq2, const gap_opt_t *opt, FileStatCollector &FSC) {
    // --frac_samp draws once per record in file order (FastqReader::fill); everything else goes through the parallel feeder
    const bool sampled = opt->frac < 1.0;
    FastqReader r[2];
    fqb_feeder *fd[2] = {nullptr, nullptr};
    if (sampled) { if (!r[0].open(fq1) || !r[1].open(fq2)) error("Open fastq failed: %s / %s", fq1.c_str(), fq2.c_str()); }
    else if (fqb_feeder_open(fq1.c_str(), 0, &fd[0]) != FQB_OK || fqb_feeder_open(fq2.c_str(), 0, &fd[1]) != FQB_OK) error("Open fastq failed: %s", fqb_last_error());
    if (fqb_stats_begin_file(h_, prefix_.c_str(), fq1.c_str(), fq2.c_str()) != FQB_OK) error("%s", fqb_last_error());
    const int stride = opt->read_len < FQB_MAX_READ_LEN ? opt->read_len : FQB_MAX_READ_LEN, cap = FQB_BATCH_PAIRS, name_stride = 64;
    // Four pinned batches: batch N is collected (pairing .. statistics on the main stream) and its lines / records are
    // formatted on the host while the align stage of batch N+1 -- submitted before -- runs on the align stream, and the
    // feeder decodes batch N+2; the fourth buffer is the one FQB_ASYNC_EMIT's writer threads may still be reading
    // (the two IO workers of the reference, src/BwtMapper.cpp:1969-1982, become one feeder per end; its "map batch N while
    // batch N+1 is read" becomes fqb_submit_pairs(N+1) before fqb_collect_pairs(N))
    constexpr int kBufs = 4;
    struct Buf { uint8_t *b[2], *q[2]; int32_t *l[2]; char *nm[2]; int n[2]; } bufs[kBufs];
    for (auto &B : bufs)
        for (int e = 0; e < 2; ++e) {
            B.b[e] = (uint8_t *)fqb_host_alloc((size_t)cap * stride); B.q[e] = (uint8_t *)fqb_host_alloc((size_t)cap * stride);
            B.l[e] = (int32_t *)fqb_host_alloc((size_t)cap * 4); B.nm[e] = (char *)fqb_host_alloc((size_t)cap * name_stride);
            if (!B.b[e] || !B.q[e] || !B.l[e] || !B.nm[e]) error("pinned host allocation failed");
            B.n[e] = 0;
        }
    int n_loaded = 0;                     // IO rounds: the reference reads batches 0 and 1 with round 0, batch k >= 1 with round k - 1
    auto load = [&](Buf &B) {
        const uint32_t seed = n_loaded ? (uint32_t)(n_loaded - 1) : 0u;
        ++n_loaded;
        auto one = [&](int e) {
            if (sampled) { B.n[e] = r[e].fill(cap, stride, B.b[e], B.q[e], B.l[e], B.nm[e], name_stride, opt->frac, seed); return; }
            const int64_t n = fqb_feeder_fill(fd[e], cap, stride, B.b[e], B.q[e], B.l[e], B.nm[e], name_stride);
            if (n < 0) error("%s", fqb_last_error());
            B.n[e] = (int)n;
        };
        std::thread t0([&]() { one(0); });
        one(1);
        t0.join();
    };
    auto good = [](const Buf &B) { return B.n[0] > 0 && B.n[1] > 0; };
    auto submit = [&](const Buf &S) {     // upload + align stage, asynchronous; the buffers stay untouched until the batch has been emitted
        if (S.n[0] != S.n[1]) error("Abort, please make sure input pair of fastq files are in the same order!");
        if (fqb_submit_pairs(h_, S.n[0], stride, S.b[0], S.q[0], S.l[0], S.b[1], S.q[1], S.l[1], 0) != FQB_OK) error("%s", fqb_last_error());
    };
    int cur = 0;
    load(bufs[0]);
    if (good(bufs[0])) { submit(bufs[0]); load(bufs[1]); }
    while (good(bufs[cur])) {
        Buf &B = bufs[cur], &N1 = bufs[(cur + 1) % kBufs], &N2 = bufs[(cur + 2) % kBufs];
        if (good(N1)) submit(N1);
        std::thread next([&]() { if (good(N1)) load(N2); else N2.n[0] = N2.n[1] = 0; });      // IO(N+2) overlaps GPU(N, N+1) and the emission of N
        if (fqb_collect_pairs(h_, nullptr, nullptr) != FQB_OK) error("%s", fqb_last_error());   // includes StatCollector's accumulation
        if (fqb_stats_emit2(h_, B.nm[0], B.nm[1], name_stride) != FQB_OK) error("%s", fqb_last_error());
        if (bam_out_ && fqb_bam_emit2(h_, B.nm[0], B.nm[1], name_stride, B.b[0], B.q[0], B.b[1], B.q[1], stride) != FQB_OK) error("%s", fqb_last_error());
        FSC.NumRead += 2LL * B.n[0];
        if (FSC.NumRead % FQB_BATCH_PAIRS == 0) fprintf(stderr, "NOTICE - %lld sequences are processed.\n", FSC.NumRead);
        next.join();
        cur = (cur + 1) % kBufs;
    }
    notice("%lld sequences are loaded.", FSC.NumRead);
    {   // src/BwtMapper.cpp:2116-2122
        int64_t c[6];
        if (fqb_stats_file_counters(h_, c) != FQB_OK) error("%s", fqb_last_error());
        notice("%ld sequences are filtered.", (long)(c[0] * 2));
        notice("%ld sequences are unmapped.", (long)(c[1] * 2));
        notice("%ld sequences are discarded of low mapQ.", (long)c[2]);
        notice("%ld sequences are retained for QC.", (long)c[3]);
    }
    if (fqb_emit_sync(h_) != FQB_OK) error("%s", fqb_last_error());      // joins the writer threads before the buffers go
    for (auto &B : bufs) for (int e = 0; e < 2; ++e) { fqb_host_free(B.b[e]); fqb_host_free(B.q[e]); fqb_host_free(B.l[e]); fqb_host_free(B.nm[e]); }
    r[0].close(); r[1].close();
    fqb_feeder_close(fd[0]); fqb_feeder_close(fd[1]);
    return 0;
}

// PairEndMapper over several devices: global batch b of the file goes to device b % G.  One host thread deals the batches:
// fqb_submit_pairs (upload + align stage, asynchronous) runs up to two batches ahead per device, fqb_collect_pairs_sharded
// takes the batches in file order through pairing .. statistics with the hand-off inside the library, and the text / BAM
// emission of batch b (the only calls that wait for the device) happens while the later batches are being aligned.
bool BwtMapper::PairEndMapperSharded(const std::string &fq1, const std::string &fq2, const gap_opt_t *opt, FileStatCollector &FSC) {
    if (opt->frac < 1.0) error("--frac_samp < 1 runs on one device (--device)");
    const int G = (int)hs_.size();
    fqb_feeder *fd[2] = {nullptr, nullptr};
    if (fqb_feeder_open(fq1.c_str(), 0, &fd[0]) != FQB_OK || fqb_feeder_open(fq2.c_str(), 0, &fd[1]) != FQB_OK) error("Open fastq failed: %s", fqb_last_error());
    for (int r = 0; r < G; ++r) {
        const std::string pre = r == 0 ? prefix_ : prefix_ + ".shard" + std::to_string(r);
        if (fqb_stats_begin_file(hs_[r], pre.c_str(), fq1.c_str(), fq2.c_str()) != FQB_OK) error("%s", fqb_last_error());
    }
    const int stride = opt->read_len < FQB_MAX_READ_LEN ? opt->read_len : FQB_MAX_READ_LEN, cap = FQB_BATCH_PAIRS, name_stride = 64;
    const int K = 2 * G + 2;              // pinned batches: two in flight per device, one being emitted, one being read
    struct Buf { uint8_t *b[2], *q[2]; int32_t *l[2]; char *nm[2]; int n[2]; };
    std::vector<Buf> bufs((size_t)K);
    for (auto &B : bufs)
        for (int e = 0; e < 2; ++e) {
            B.b[e] = (uint8_t *)fqb_host_alloc((size_t)cap * stride); B.q[e] = (uint8_t *)fqb_host_alloc((size_t)cap * stride);
            B.l[e] = (int32_t *)fqb_host_alloc((size_t)cap * 4); B.nm[e] = (char *)fqb_host_alloc((size_t)cap * name_stride);
            if (!B.b[e] || !B.q[e] || !B.l[e] || !B.nm[e]) error("pinned host allocation failed");
            B.n[e] = 0;
        }
    int64_t b_load = 0, b_sub = 0, b_col = 0;      // batches read / submitted / collected so far
    bool eof = false;
    auto load = [&]() {
        Buf &B = bufs[(size_t)(b_load % K)];
        auto one = [&](int e) {
            const int64_t n = fqb_feeder_fill(fd[e], cap, stride, B.b[e], B.q[e], B.l[e], B.nm[e], name_stride);
            if (n < 0) error("%s", fqb_last_error());
            B.n[e] = (int)n;
        };
        std::thread t0([&]() { one(0); });
        one(1);
        t0.join();
        if (B.n[0] != B.n[1]) error("Abort, please make sure input pair of fastq files are in the same order!");
        if (B.n[0] == 0) eof = true; else ++b_load;
    };
    for (;;) {
        while (!eof && b_load < b_col + 2 * G + 1 && b_load - b_col < K - 1) load();
        if (b_col >= b_load) break;
        for (; b_sub < b_load && b_sub < b_col + 2 * G; ++b_sub) {
            Buf &S = bufs[(size_t)(b_sub % K)];
            if (fqb_submit_pairs(hs_[b_sub % G], S.n[0], stride, S.b[0], S.q[0], S.l[0], S.b[1], S.q[1], S.l[1], 0) != FQB_OK) error("%s", fqb_last_error());
        }
        Buf &B = bufs[(size_t)(b_col % K)];
        fqb_handle *h = hs_[b_col % G];
        const bool is_last = eof && b_col == b_load - 1;
        if (fqb_collect_pairs_sharded(h, nullptr, nullptr, (uint64_t)b_col, pair_base_, is_last ? 1 : 0) != FQB_OK) error("%s", fqb_last_error());
        if (fqb_stats_emit2(h, B.nm[0], B.nm[1], name_stride) != FQB_OK) error("%s", fqb_last_error());
        if (bam_out_ && fqb_bam_emit2(h, B.nm[0], B.nm[1], name_stride, B.b[0], B.q[0], B.b[1], B.q[1], stride) != FQB_OK) error("%s", fqb_last_error());
        pair_base_ += (uint64_t)B.n[0];
        FSC.NumRead += 2LL * B.n[0];
        if (FSC.NumRead % FQB_BATCH_PAIRS == 0) fprintf(stderr, "NOTICE - %lld sequences are processed.\n", FSC.NumRead);
        ++b_col;
    }
    notice("%lld sequences are loaded.", FSC.NumRead);
    {   // src/BwtMapper.cpp:2116-2122, summed over the devices
        int64_t c[6] = {0, 0, 0, 0, 0, 0}, one[6];
        for (fqb_handle *h : hs_) { if (fqb_stats_file_counters(h, one) != FQB_OK) error("%s", fqb_last_error()); for (int k = 0; k < 6; ++k) c[k] += one[k]; }
        notice("%ld sequences are filtered.", (long)(c[0] * 2));
        notice("%ld sequences are unmapped.", (long)(c[1] * 2));
        notice("%ld sequences are discarded of low mapQ.", (long)c[2]);
        notice("%ld sequences are retained for QC.", (long)c[3]);
    }
    for (fqb_handle *h : hs_) if (fqb_emit_sync(h) != FQB_OK) error("%s", fqb_last_error());
    for (auto &B : bufs) for (int e = 0; e < 2; ++e) { fqb_host_free(B.b[e]); fqb_host_free(B.q[e]); fqb_host_free(B.l[e]); fqb_host_free(B.nm[e]); }
    fqb_feeder_close(fd[0]); fqb_feeder_close(fd[1]);
    return 0;
}

// BwtMapper::SingleEndMapper (src/BwtMapper.cpp:1266-1407): same stages on batches that carry first reads only
bool BwtMapper::SingleEndMapper(const std::string &fq1, const gap_opt_t *opt, FileStatCollector &FSC) {
    fqb_feeder *fd = nullptr;
    if (fqb_feeder_open(fq1.c_str(), 0, &fd) != FQB_OK) error("Open fastq failed: %s", fqb_last_error());
    if (fqb_stats_begin_file(h_, prefix_.c_str(), fq1.c_str(), fq1.c_str()) != FQB_OK) error("%s", fqb_last_error());
    const int stride = opt->read_len < FQB_MAX_READ_LEN ? opt->read_len : FQB_MAX_READ_LEN, cap = FQB_BATCH_PAIRS, name_stride = 64;
    constexpr int kBufs = 4;               // see PairEndMapper
    struct Buf { uint8_t *b, *q; int32_t *l; char *nm; int n; } bufs[kBufs];
    for (auto &B : bufs) {
        B.b = (uint8_t *)fqb_host_alloc((size_t)cap * stride); B.q = (uint8_t *)fqb_host_alloc((size_t)cap * stride);
        B.l = (int32_t *)fqb_host_alloc((size_t)cap * 4); B.nm = (char *)fqb_host_alloc((size_t)cap * name_stride);
        if (!B.b || !B.q || !B.l || !B.nm) error("pinned host allocation failed");
        B.n = 0;
    }
    auto load = [&](Buf &B) {
        const int64_t n = fqb_feeder_fill(fd, cap, stride, B.b, B.q, B.l, B.nm, name_stride);
        if (n < 0) error("%s", fqb_last_error());
        B.n = (int)n;
    };
    int cur = 0;
    load(bufs[0]);
    if (bufs[0].n > 0) load(bufs[1]);
    while (bufs[cur].n > 0) {
        Buf &B = bufs[cur], &N1 = bufs[(cur + 1) % kBufs], &N2 = bufs[(cur + 2) % kBufs];
        if (N1.n > 0 && fqb_prefetch_pairs(h_, N1.n, stride, N1.b, N1.q, N1.l, nullptr, nullptr, nullptr) != FQB_OK) error("%s", fqb_last_error());
        std::thread next([&]() { if (N1.n > 0) load(N2); else N2.n = 0; });
        if (fqb_align_pairs(h_, B.n, stride, B.b, B.q, B.l, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) != FQB_OK) error("%s", fqb_last_error());
        if (fqb_stage_stats(h_) != FQB_OK) error("%s", fqb_last_error());
        if (fqb_stats_emit(h_, B.nm, name_stride) != FQB_OK) error("%s", fqb_last_error());
        if (bam_out_ && fqb_bam_emit(h_, B.nm, name_stride, B.b, B.q, nullptr, nullptr, stride) != FQB_OK) error("%s", fqb_last_error());
        FSC.NumRead += B.n;
        next.join();
        cur = (cur + 1) % kBufs;
    }
    notice("%lld sequences are loaded.", FSC.NumRead);
    {   // src/BwtMapper.cpp:1395-1397 (the reference prints them after every batch; here once per file)
        int64_t c[6];
        if (fqb_stats_file_counters(h_, c) != FQB_OK) error("%s", fqb_last_error());
        notice("%ld sequences are filtered.", (long)c[0]);
        notice("%ld sequences are retained for QC.", (long)c[3]);
    }
    if (fqb_emit_sync(h_) != FQB_OK) error("%s", fqb_last_error());
    for (auto &B : bufs) { fqb_host_free(B.b); fqb_host_free(B.q); fqb_host_free(B.l); fqb_host_free(B.nm); }
    fqb_feeder_close(fd);
    return 0;
}

// `FASTQuick align` option table (src/FASTQuick.cpp:177-306): --name value, booleans as bare flags
int runAlign(int argc, char **argv) {
    double t_real = realtime();
    gap_opt_t opt; pe_opt_t popt;
    std::string Fastq_1("Empty"), Fastq_2("Empty"), FaList("Empty"), BamIn("Empty"), Prefix("Empty"), IndexPrefix("Empty"), ReadGroup("@RG\tID:foo\tSM:bar");
    int kmer_thresh = 3, opte = -1;
    std::vector<int> devices{0};
    bool nonstop = false, il13 = false, loggap = false, sam_out = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) error("option %s needs a value", a.c_str()); return argv[++i]; };
        if (a == "--fq_list") FaList = val(); else if (a == "--fastq_1") Fastq_1 = val(); else if (a == "--fastq_2") Fastq_2 = val();
        else if (a == "--bam_in") BamIn = val(); else if (a == "--sam_out") sam_out = true;
        else if (a == "--out_prefix") Prefix = val(); else if (a == "--index_prefix") IndexPrefix = val();
        else if (a == "--kmer_thresh") kmer_thresh = atoi(val()); else if (a == "--n") opt.fnr = atof(val());
        else if (a == "--o") opt.max_gapo = atoi(val()); else if (a == "--e") opte = atoi(val()); else if (a == "--i") opt.indel_end_skip = atoi(val());
        else if (a == "--d") opt.max_del_occ = atoi(val()); else if (a == "--l") opt.seed_len = atoi(val()); else if (a == "--k") opt.max_seed_diff = atoi(val());
        else if (a == "--m") opt.max_entries = atoi(val()); else if (a == "--t") opt.n_threads = atoi(val()); else if (a == "--R") opt.max_top2 = atoi(val());
        else if (a == "--q") opt.trim_qual = atoi(val()); else if (a == "--RG") ReadGroup = val();
        else if (a == "--N") nonstop = true; else if (a == "--I") il13 = true; else if (a == "--L") loggap = true;
        else if (a == "--max_isize") popt.max_isize = atoi(val()); else if (a == "--max_occ") popt.max_occ = (unsigned)atoi(val());
        else if (a == "--is_sw") popt.is_sw = 1; else if (a == "--n_multi") popt.n_multi = atoi(val()); else if (a == "--N_multi") popt.N_multi = atoi(val());
        else if (a == "--ap_prior") popt.ap_prior = atof(val()); else if (a == "--force_isize") popt.force_isize = 1;
        else if (a == "--cal_dup") opt.cal_dup = 1; else if (a == "--frac_samp") opt.frac = atof(val());
        else if (a == "--device") devices.assign(1, atoi(val()));
        else if (a == "--devices") {                 // "0,1,2,3" or "0-7"
            devices.clear();
            std::string v = val(), tok;
            std::stringstream ss(v);
            while (std::getline(ss, tok, ',')) {
                const size_t dash = tok.find('-');
                if (dash != std::string::npos && dash > 0) { for (int d = atoi(tok.substr(0, dash).c_str()); d <= atoi(tok.substr(dash + 1).c_str()); ++d) devices.push_back(d); }
                else if (!tok.empty()) devices.push_back(atoi(tok.c_str()));
            }
            if (devices.empty()) error("--devices needs a list such as 0,1 or 0-7");
        }
        else error("unknown option %s", a.c_str());
    }
    if (opt.fnr >= 1.0) { opt.max_diff = (int)opt.fnr; opt.fnr = -1.0; }
    if (Prefix == "Empty") error("--out_prefix is required");
    if (IndexPrefix == "Empty") error("--index_prefix is required");
    if (BamIn != "Empty") error("Input alignments from Bam file is disabled.");
    if (opt.frac < 1.0 && Fastq_2 == "Empty" && FaList == "Empty") error("--frac_samp < 1 with single-end input is seeded from clock() in the reference and is not reproducible; not supported");
    if (sam_out) opt.out_bam = 0;
    opt.RG = ReadGroup;
    if (opte > 0) { opt.max_gape = opte; opt.mode &= ~0x01; }
    if (nonstop) { opt.mode |= 0x10; opt.max_top2 = 0x7fffffff; }
    if (il13) opt.mode |= 0x200;
    if (loggap) opt.mode |= 0x04;
    BwtIndexer Indexer(kmer_thresh);
    std::string NewRef = IndexPrefix + ".FASTQuick.fa", TargetRegionPath("Empty");
    {   // <index>.FASTQuick.fa.param (src/FASTQuick.cpp:365-467)
        std::ifstream par(NewRef + ".param");
        if (!par) error("Open %s failed!", (NewRef + ".param").c_str());
        std::string k, v;
        while (par >> k >> v) {
            if (k == "TARGET_REGION_PATH") TargetRegionPath = v;          // set at index time by --regionList
            else if (k == "NUM_VAR_LONG") opt.num_variant_long = (unsigned)atoi(v.c_str());
            else if (k == "NUM_VAR_SHORT") opt.num_variant_short = (unsigned)atoi(v.c_str());
            else if (k == "SHORT_FLANK_LENGTH") opt.flank_len = atoi(v.c_str());
            else if (k == "LONG_FLANK_LENGTH") opt.flank_long_len = atoi(v.c_str());
        }
    }
    double t_tmp = realtime();
    Indexer.LoadIndex(NewRef);
    notice("Load Index... %f sec", realtime() - t_tmp);
    t_tmp = realtime();
    if (TargetRegionPath != "Empty") notice("Read in target region from %s", TargetRegionPath.c_str());
    BwtMapper Mapper(Indexer, FaList, Fastq_1, Fastq_2, Prefix, NewRef, &popt, &opt, TargetRegionPath, devices);
    notice("Mapping... %f sec", realtime() - t_tmp);
    notice("Real time: %.3f sec", realtime() - t_real);
    return 0;
}

}  // namespace fqb200
