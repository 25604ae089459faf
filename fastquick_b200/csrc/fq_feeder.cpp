// Host feeder: FASTQ text -> fixed-stride batches in pinned memory, decoded in parallel.
// Replaces the reference's reader for this path: kseq_read3_fpc (libbwa/kseq.h:327-370) called record by record from
// bwa_read_seq_with_hash_dev (src/BwtMapper.cpp:476-613) on the two IO workers of PairEndMapper (:1969-1982).
//
//   producer thread   file -> text blocks of ~4 MiB, in order.  Three sources:
//                       BGZF (bgzip / bcl2fastq output: gzip members <= 64 KiB that carry their own size) - members
//                            are inflated independently by the worker pool, 64 at a time;
//                       any other gzip stream - inherently serial: the file is memory-mapped and decoded by this
//                            library's own inflate loop (fq_inflate.cpp), each member's CRC-32 and length checked;
//                            zlib's inflate() when the file cannot be mapped or FQB_GZIP_ZLIB is set;
//                       plain text.
//   fill()            walks the blocks, finds the record boundaries (memchr), cuts the text into runs of whole records
//                     and hands each run to the worker pool together with the batch slot of its first record;
//                     a record cut by a block boundary is stitched and parsed inline.
//   worker pool       parses runs into the batch: bases, qualities, lengths, names (what the GPU prep kernel and the
//                     writers consume); shared by all open feeders.
// The nt4 encoding, trimming and the k-mer filter of bwa_read_seq_with_hash_dev stay on the GPU (prep_kernel).
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fastquick_b200.h"
#include "fq_common.h"
#include "fq_inflate.h"

namespace fqb {
namespace {

constexpr size_t kBlockBytes = (size_t)4 << 20;
constexpr int kRunRecords = 4096;          // records per parse job
constexpr int kQueueDepth = 8;             // text blocks buffered ahead of fill()
constexpr int kBgzfGroup = 64;             // BGZF members per inflate job
constexpr int kBgzfWindow = 12;            // inflate jobs in flight

class WorkPool {
public:
    explicit WorkPool(unsigned n) {
        for (unsigned i = 0; i < n; ++i) th_.emplace_back([this]() { run(); });
    }
    ~WorkPool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void submit(std::function<void()> f) {
        { std::lock_guard<std::mutex> l(m_); q_.push_back(std::move(f)); }
        cv_.notify_one();
    }
    unsigned size() const { return (unsigned)th_.size(); }
private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&]() { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front()); q_.pop_front();
            }
            f();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_; std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    bool stop_ = false;
};

WorkPool &pool(int n_threads) {
    static std::mutex m;
    static std::unique_ptr<WorkPool> p;
    std::lock_guard<std::mutex> l(m);
    const unsigned want = n_threads > 0 ? (unsigned)n_threads : std::min(16u, std::max(2u, std::thread::hardware_concurrency()));
    if (!p) p.reset(new WorkPool(want));                         // sized by the first feeder, shared by all later ones
    return *p;
}

struct Latch {
    std::mutex m; std::condition_variable cv; int pending = 0;
    void add() { std::lock_guard<std::mutex> l(m); ++pending; }
    void done() { std::lock_guard<std::mutex> l(m); if (--pending == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> l(m); cv.wait(l, [&]() { return pending == 0; }); }
};

// Text blocks come from a free list: a fresh 4 MiB allocation costs its page faults (and a zero fill) every time, which
// on the serial gzip path is paid by the one thread that sets the pace.
class BufferPool {
public:
    static constexpr size_t kBufBytes = kBlockBytes + 32768 + 1024;       // every request of the feeder fits
    static constexpr size_t kKeep = 48;                                   // buffers kept for reuse (~200 MiB at most)
    static BufferPool &get() { static BufferPool *p = new BufferPool(); return *p; }
    char *take(size_t cap, bool &pooled) {
        pooled = cap <= kBufBytes;
        if (pooled) {
            std::lock_guard<std::mutex> l(m_);
            if (!free_.empty()) { char *p = free_.back(); free_.pop_back(); return p; }
        }
        char *p = (char *)malloc(pooled ? kBufBytes : cap);
        if (!p) throw std::bad_alloc();
        return p;
    }
    void give(char *p, bool pooled) {
        if (pooled) {
            std::lock_guard<std::mutex> l(m_);
            if (free_.size() < kKeep) { free_.push_back(p); return; }
        }
        free(p);
    }
private:
    std::mutex m_;
    std::vector<char *> free_;
};

struct Block {
    struct Text {                                                          // the buffer: data() .. data() + size()
        char *p = nullptr; size_t cap = 0;
        char *data() const { return p; }
        size_t size() const { return cap; }
    } text;
    size_t off = 0, n = 0;                                                 // the text is [off, off + n)
    bool pooled = false;
    Block() = default;
    Block(const Block &) = delete;
    Block &operator=(const Block &) = delete;
    ~Block() { if (text.p) BufferPool::get().give(text.p, pooled); }
};
using BlockPtr = std::shared_ptr<Block>;

struct Batch {
    int stride, name_stride;
    uint8_t *bases, *quals; int32_t *lens; char *names;
    // optional second form of the same rows for the upload (fqb_feeder_fill_packed): 2-bit bases, qualities with the not-ACGT flag
    int packed_stride = 0; uint8_t *packed = nullptr, *qflag = nullptr;
};
// nst_nt4_table (libbwa/bntseq.c:38-55); the packed rows store code & 3 and flag the codes above 3 in bit 7 of the quality
struct Nt4Table {
    uint8_t t[256];
    Nt4Table() { for (int c = 0; c < 256; ++c) t[c] = 4; t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; t['-'] = 5; }
};
static const Nt4Table kNt4;

// first failure of any job of this feeder; reported by fill()
struct Failure {
    std::mutex m; std::string msg; std::atomic<bool> set{false};
    void raise(const std::string &s) { std::lock_guard<std::mutex> l(m); if (!set.load()) { msg = s; set.store(true); } }
};

BlockPtr new_block(size_t cap) {
    BlockPtr b = std::make_shared<Block>();
    b->text.p = BufferPool::get().take(cap, b->pooled);
    b->text.cap = cap;
    return b;
}

// A whole gzip file (one or more members, zero padding behind the last one tolerated as gzip(1) does) held in memory
// -> text blocks of about block_bytes, through fqb::Inflater.  Each block starts with the last 32 KiB of the block
// before it (the deflate window), so a match never leaves its block.  Every member's CRC-32 and length are checked.
// With submit (the worker pool) the checksums are computed off the calling thread, one job per block, and folded with
// crc32_combine when the member ends.  false = corrupt input (err set) or push() refused a block (err empty).
bool gunzip_blocks(const uint8_t *p, const uint8_t *const end, size_t block_bytes, const std::function<bool(BlockPtr)> &push,
                   const std::function<void(std::function<void()>)> &submit, std::string &err) {
    constexpr size_t W = Inflater::kWindow;
    if (block_bytes < 2 * Inflater::kOutSlack) block_bytes = 2 * Inflater::kOutSlack;
    struct CrcPart { BlockPtr keep; const uint8_t *p; size_t n; uLong crc; };
    std::vector<std::shared_ptr<CrcPart>> parts;
    Latch sums;
    struct WaitAll { Latch &l; ~WaitAll() { l.wait(); } } wait_all{sums};       // no checksum job may outlive this frame
    std::unique_ptr<Inflater> inf(new Inflater());
    BlockPtr b = new_block(W + block_bytes);
    b->off = W;
    uint8_t *out = (uint8_t *)b->text.data() + W, *out_end = (uint8_t *)b->text.data() + b->text.size();
    auto hand_over = [&](const uint8_t *floor, const uint8_t **new_floor) -> bool {
        b->n = (size_t)(out - ((uint8_t *)b->text.data() + W));
        const size_t hist = std::min(W, (size_t)(out - floor));
        BlockPtr nb = new_block(W + block_bytes);
        nb->off = W;
        memcpy(nb->text.data() + W - hist, out - hist, hist);
        const bool ok = b->n == 0 || push(b);
        b = nb;
        out = (uint8_t *)b->text.data() + W; out_end = (uint8_t *)b->text.data() + b->text.size();
        *new_floor = out - hist;
        return ok;
    };
    while (p < end) {
        if (end - p < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xe0)) {
            for (const uint8_t *q = p; q < end; ++q) if (*q) { err = "corrupt gzip stream: no member header where one is due"; return false; }
            break;
        }
        const unsigned flg = p[3];
        const uint8_t *q = p + 10;
        if (flg & 4) q += 2 + ((size_t)q[0] | ((size_t)q[1] << 8));                       // FEXTRA (end - p >= 18 covers the length field)
        if (flg & 8) { while (q < end && *q) ++q; ++q; }                                  // FNAME
        if (flg & 16) { while (q < end && *q) ++q; ++q; }                                 // FCOMMENT
        if (flg & 2) q += 2;                                                              // FHCRC
        if (q >= end) { err = "corrupt gzip stream: truncated member header"; return false; }
        inf->reset(q, end);
        const uint8_t *floor = out, *chunk = out;
        uint64_t isize = 0;
        parts.clear();
        for (;;) {
            const Inflater::Status s = inf->run(out, out_end, floor);
            if (s == Inflater::kError) { err = std::string("corrupt gzip stream: ") + inf->error(); return false; }
            if (out > chunk) {
                auto part = std::make_shared<CrcPart>(CrcPart{b, chunk, (size_t)(out - chunk), 0});
                parts.push_back(part);
                if (submit) { sums.add(); Latch *l = &sums; submit([part, l]() { part->crc = crc32_z(crc32(0L, Z_NULL, 0), part->p, part->n); l->done(); }); }
                else part->crc = crc32_z(crc32(0L, Z_NULL, 0), part->p, part->n);
                isize += part->n;
            }
            if (s == Inflater::kStreamEnd) break;
            if (!hand_over(floor, &floor)) return false;
            chunk = out;
        }
        sums.wait();
        uLong crc = crc32(0L, Z_NULL, 0);
        for (auto &part : parts) crc = crc32_combine(crc, part->crc, (z_off_t)part->n);
        const uint8_t *t = inf->in_pos();
        if (end - t < 8) { err = "corrupt gzip stream: truncated member"; return false; }
        const uint32_t want_crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        const uint32_t want_len = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        if (want_crc != (uint32_t)crc || want_len != (uint32_t)isize) { err = "corrupt gzip stream: member checksum or length mismatch"; return false; }
        p = t + 8;
    }
    b->n = (size_t)(out - ((uint8_t *)b->text.data() + W));
    return b->n == 0 || push(b);
}

inline const char *line_end(const char *p, const char *end) {
    const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
    return nl ? nl : end;
}

// one record whose four lines are [p, end): header, bases, '+', qualities (kseq's contract: name up to the first blank,
// a trailing /1 or /2 removed as bwa_read_seq_with_hash_dev does, src/BwtMapper.cpp:598-603)
bool parse_record(const char *p, const char *end, const Batch &B, int slot, Failure &fail, const char **next) {
    const char *l[4], *e[4];
    for (int k = 0; k < 4; ++k) {
        l[k] = p;
        const char *nl = line_end(p, end);
        e[k] = (nl > p && nl[-1] == '\r') ? nl - 1 : nl;
        p = nl < end ? nl + 1 : end;
    }
    *next = p;
    if (l[0] == e[0] || l[0][0] != '@') { fail.raise("malformed FASTQ record: header line does not start with '@'"); return false; }
    const size_t sl = (size_t)(e[1] - l[1]);
    if ((int)sl > B.stride) { fail.raise("read longer than " + std::to_string(B.stride) + " bases: not supported"); return false; }
    if ((size_t)(e[3] - l[3]) != sl) { fail.raise("sequence and quality lengths differ in a FASTQ record"); return false; }
    uint8_t *b = B.bases + (size_t)slot * B.stride, *q = B.quals + (size_t)slot * B.stride;
    memcpy(b, l[1], sl); memset(b + sl, 'N', (size_t)B.stride - sl);
    memcpy(q, l[3], sl); memset(q + sl, '!', (size_t)B.stride - sl);
    B.lens[slot] = (int32_t)sl;
    if (B.packed) {
        uint32_t *w = reinterpret_cast<uint32_t *>(B.packed + (size_t)slot * B.packed_stride);
        uint8_t *qf = B.qflag + (size_t)slot * B.stride;
        uint32_t q_or = 0;
        for (int k = 0, j = 0; k < B.packed_stride / 4; ++k) {          // one word (16 bases) at a time, accumulated in a register
            uint32_t acc = 0;
            const int end = j + 16 < B.stride ? j + 16 : B.stride;
            for (int sh = 0; j < end; ++j, sh += 2) {
                const uint32_t c = kNt4.t[b[j]];
                acc |= (c & 3u) << sh;
                q_or |= q[j];
                qf[j] = (uint8_t)(q[j] | ((c & 4u) << 5));              // codes 4 and 5 -> bit 7
            }
            w[k] = acc;
        }
        if (q_or & 0x80u) { fail.raise("quality byte above 127 in a FASTQ record"); return false; }
    }
    const char *nb = l[0] + 1, *ne = nb;
    while (ne < e[0] && *ne != ' ' && *ne != '\t') ++ne;
    if (ne - nb > 2 && ne[-2] == '/' && (ne[-1] == '1' || ne[-1] == '2')) ne -= 2;
    char *dst = B.names + (size_t)slot * B.name_stride;
    const size_t nl = std::min((size_t)(ne - nb), (size_t)B.name_stride - 1);
    memcpy(dst, nb, nl); memset(dst + nl, 0, (size_t)B.name_stride - nl);
    return true;
}

}  // namespace

class Feeder {
public:
    ~Feeder() { close(); }
    bool open(const std::string &path, int n_threads, std::string &err) {
        fp_ = fopen(path.c_str(), "rb");
        if (!fp_) { err = "cannot open " + path; return false; }
        pool_ = &pool(n_threads);
        unsigned char head[18];
        const size_t got = fread(head, 1, sizeof(head), fp_);
        rewind(fp_);
        kind_ = 0;
        if (got >= 2 && head[0] == 0x1f && head[1] == 0x8b) {
            kind_ = 1;
            if (got >= 18 && (head[3] & 4) && head[12] == 'B' && head[13] == 'C' && head[14] == 2 && head[15] == 0) kind_ = 2;
        }
        stop_ = false; done_ = false;
        producer_ = std::thread([this]() { produce(); });
        return true;
    }
    void close() {
        if (producer_.joinable()) {
            { std::lock_guard<std::mutex> l(qm_); stop_ = true; }
            qcv_.notify_all();
            producer_.join();
        }
        if (fp_) fclose(fp_);
        fp_ = nullptr; cur_.reset(); queue_.clear(); carry_.clear();
    }
    int kind() const { return kind_; }

    // up to n_max records into the batch; 0 at end of file; -1 on a malformed input (message in err)
    int fill(int n_max, const Batch &B, std::string &err) {
        Latch latch;
        int n = 0;
        while (n < n_max && !fail_.set.load()) {
            if (!cur_ || pos_ == cur_->n) {
                cur_ = next_block();
                pos_ = 0;
                if (!cur_) {                                   // end of input: a last record without its final newline
                    if (!carry_.empty()) {
                        const char *nx;
                        int nl = 0; for (char c : carry_) nl += c == '\n';
                        if (nl < 3) fail_.raise("truncated FASTQ record");
                        else if (parse_record(carry_.data(), carry_.data() + carry_.size(), B, n, fail_, &nx)) ++n;
                        carry_.clear();
                    }
                    break;
                }
            }
            const char *base = cur_->text.data() + cur_->off, *end = base + cur_->n, *p = base + pos_;
            if (!carry_.empty()) {                              // finish the record the previous block cut
                int nl = 0; for (char c : carry_) nl += c == '\n';
                const char *q = p;
                while (nl < 4 && q < end) { const char *e = line_end(q, end); if (e == end) { q = end; break; } q = e + 1; ++nl; }
                carry_.append(p, q);
                pos_ = (size_t)(q - base);
                if (nl < 4) continue;                           // still incomplete: next block
                const char *nx;
                if (parse_record(carry_.data(), carry_.data() + carry_.size(), B, n, fail_, &nx)) ++n;
                carry_.clear();
                continue;
            }
            // runs of whole records
            while (n < n_max && p < end) {
                const char *run = p; int k = 0;
                const int want = std::min(kRunRecords, n_max - n);
                const char *rec = p;
                while (k < want) {
                    while (rec < end && (*rec == '\n' || *rec == '\r')) ++rec;     // blank lines between records
                    if (k == 0) run = rec;
                    const char *q = rec; int nl = 0;
                    while (nl < 4) { const char *e = line_end(q, end); if (e == end) break; q = e + 1; ++nl; }
                    if (nl < 4) break;
                    rec = q; ++k;
                }
                if (k > 0) {
                    BlockPtr keep = cur_;
                    const int first = n;
                    const char *run_end = rec;
                    latch.add();
                    pool_->submit([keep, run, run_end, first, k, B, this, &latch]() {
                        const char *q = run;
                        for (int i = 0; i < k && !fail_.set.load(); ++i) {
                            while (q < run_end && (*q == '\n' || *q == '\r')) ++q;
                            if (!parse_record(q, run_end, B, first + i, fail_, &q)) break;
                        }
                        latch.done();
                    });
                    n += k;
                    p = rec;
                }
                if (k < want) {                                 // the block ends inside a record (or in blank lines)
                    while (p < end && (*p == '\n' || *p == '\r')) ++p;
                    carry_.assign(p, end);
                    p = end;
                }
            }
            pos_ = (size_t)(p - base);
        }
        latch.wait();
        if (fail_.set.load()) { err = fail_.msg; return -1; }
        return n;
    }

private:
    BlockPtr next_block() {
        std::unique_lock<std::mutex> l(qm_);
        qcv_.wait(l, [&]() { return !queue_.empty() || done_; });
        if (queue_.empty()) return nullptr;
        BlockPtr b = queue_.front(); queue_.pop_front();
        l.unlock();
        qcv_.notify_all();
        return b;
    }
    bool push_block(BlockPtr b) {
        std::unique_lock<std::mutex> l(qm_);
        qcv_.wait(l, [&]() { return stop_ || (int)queue_.size() < kQueueDepth; });
        if (stop_) return false;
        queue_.push_back(std::move(b));
        l.unlock();
        qcv_.notify_all();
        return true;
    }
    void finish() {
        { std::lock_guard<std::mutex> l(qm_); done_ = true; }
        qcv_.notify_all();
    }

    void produce() {
        if (kind_ == 0) produce_text();
        else if (kind_ == 2) produce_bgzf();
        else if (getenv("FQB_GZIP_ZLIB") || !produce_gzip_mapped()) produce_gzip();
        if (getenv("FQB_FEEDER_DEBUG")) {                         // what the serial side of this file cost
            struct timespec ts;
            clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
            fprintf(stderr, "feeder: producer thread of a %s input used %.3f s of CPU\n", kind_ == 0 ? "text" : kind_ == 1 ? "gzip" : "BGZF", (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec);
        }
        finish();
    }
    void produce_text() {
        for (;;) {
            BlockPtr b = new_block(kBlockBytes);
            b->n = fread(b->text.data(), 1, b->text.size(), fp_);
            if (b->n == 0) return;
            if (!push_block(b)) return;
        }
    }
    // gzip stream from a memory-mapped file through gunzip_blocks().  false = the file could not be mapped and nothing
    // has been produced.
    bool produce_gzip_mapped() {
        struct stat st;
        if (fstat(fileno(fp_), &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 18) return false;
        const size_t size = (size_t)st.st_size;
        void *map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fileno(fp_), 0);
        if (map == MAP_FAILED) return false;
        madvise(map, size, MADV_SEQUENTIAL);
        std::string err;
        if (!gunzip_blocks((const uint8_t *)map, (const uint8_t *)map + size, kBlockBytes, [this](BlockPtr b) { return push_block(std::move(b)); },
                           [this](std::function<void()> f) { pool_->submit(std::move(f)); }, err) && !err.empty())
            fail_.raise(err);
        munmap(map, size);
        return true;
    }
    void produce_gzip() {
        z_stream zs; memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, 31) != Z_OK) { fail_.raise("inflateInit2 failed"); return; }
        std::vector<unsigned char> in((size_t)1 << 20);
        BlockPtr b = new_block(kBlockBytes);
        zs.next_out = (Bytef *)b->text.data(); zs.avail_out = (uInt)b->text.size();
        bool eof = false, ok = true;
        while (ok && !eof) {
            zs.avail_in = (uInt)fread(in.data(), 1, in.size(), fp_);
            zs.next_in = in.data();
            if (zs.avail_in == 0) break;
            while (zs.avail_in > 0) {
                const int rc = inflate(&zs, Z_NO_FLUSH);
                if (rc == Z_STREAM_END) {                       // next gzip member, if any
                    if (inflateReset(&zs) != Z_OK) { fail_.raise("inflateReset failed"); ok = false; break; }
                } else if (rc != Z_OK && rc != Z_BUF_ERROR) { fail_.raise("corrupt gzip stream"); ok = false; break; }
                if (zs.avail_out == 0) {
                    b->n = b->text.size();
                    if (!push_block(b)) { ok = false; break; }
                    b = new_block(kBlockBytes);
                    zs.next_out = (Bytef *)b->text.data(); zs.avail_out = (uInt)b->text.size();
                }
            }
        }
        b->n = b->text.size() - zs.avail_out;
        if (ok && b->n) push_block(b);
        inflateEnd(&zs);
    }
    struct Group {
        std::vector<unsigned char> comp;                        // whole members back to back
        std::vector<std::pair<uint32_t, uint32_t>> member;      // offset in comp, size
        BlockPtr out;
        std::mutex m; std::condition_variable cv; bool ready = false;
    };
    void produce_bgzf() {
        std::deque<std::shared_ptr<Group>> flight;
        std::vector<unsigned char> buf; size_t have = 0, at = 0;
        buf.resize((size_t)8 << 20);
        bool eof = false;
        auto drain_one = [&]() -> bool {
            std::shared_ptr<Group> g = flight.front(); flight.pop_front();
            { std::unique_lock<std::mutex> l(g->m); g->cv.wait(l, [&]() { return g->ready; }); }
            if (g->out->n == 0) return true;
            return push_block(g->out);
        };
        for (;;) {
            auto g = std::make_shared<Group>();
            while ((int)g->member.size() < kBgzfGroup) {
                if (have - at < 18 + 8 && !eof) {                // refill, keeping the unread tail
                    memmove(buf.data(), buf.data() + at, have - at); have -= at; at = 0;
                    const size_t r = fread(buf.data() + have, 1, buf.size() - have, fp_);
                    if (r == 0) eof = true;
                    have += r;
                }
                if (have - at == 0) break;
                if (have - at < 18) { fail_.raise("truncated BGZF member"); break; }
                const unsigned char *h = buf.data() + at;
                if (h[0] != 0x1f || h[1] != 0x8b || !(h[3] & 4) || h[12] != 'B' || h[13] != 'C') { fail_.raise("not a BGZF member"); break; }
                const size_t bsize = (size_t)(h[16] | (h[17] << 8)) + 1;
                {   // untrusted input: the member must at least hold its header, its extra field and CRC32 + ISIZE, and a BGZF
                    // member never inflates to more than 64 KiB (the inflate job sizes its output from the ISIZE fields)
                    const size_t xlen = (size_t)h[10] | ((size_t)h[11] << 8);
                    if (bsize < 12 + xlen + 8 || xlen < 6) { fail_.raise("corrupt BGZF member (BSIZE smaller than its own header)"); break; }
                }
                if (have - at < bsize) {
                    if (eof) { fail_.raise("truncated BGZF member"); break; }
                    memmove(buf.data(), buf.data() + at, have - at); have -= at; at = 0;
                    const size_t r = fread(buf.data() + have, 1, buf.size() - have, fp_);
                    if (r == 0) eof = true;
                    have += r;
                    continue;
                }
                g->member.emplace_back((uint32_t)g->comp.size(), (uint32_t)bsize);
                g->comp.insert(g->comp.end(), h, h + bsize);
                at += bsize;
            }
            if (fail_.set.load() || g->member.empty()) break;
            flight.push_back(g);
            Failure *fail = &fail_;
            pool_->submit([g, fail]() {
                size_t total = 0;
                bool sane = true;
                for (auto &mb : g->member) {
                    const unsigned char *t = g->comp.data() + mb.first + mb.second - 4;
                    const size_t isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);     // ISIZE
                    if (isize > 65536) sane = false;
                    total += isize;
                }
                if (!sane) {
                    fail->raise("corrupt BGZF member (ISIZE above 64 KiB)");
                    g->out = new_block(Inflater::kOutSlack); g->out->n = 0;
                    { std::lock_guard<std::mutex> l(g->m); g->ready = true; }
                    g->cv.notify_all();
                    return;
                }
                g->out = new_block(total + Inflater::kOutSlack);
                std::unique_ptr<Inflater> inf(new Inflater());
                uint8_t *const base = (uint8_t *)g->out->text.data(), *const out_end = base + g->out->text.size();
                bool ok = true;
                size_t o = 0;
                for (auto &mb : g->member) {                    // each member: raw deflate between the header and CRC32 + ISIZE
                    const unsigned char *h = g->comp.data() + mb.first, *t = h + mb.second - 8;
                    const size_t xlen = (size_t)h[10] | ((size_t)h[11] << 8);
                    if (12 + xlen + 8 > mb.second) { ok = false; break; }
                    const uint32_t want_crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
                    const uint32_t want_len = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
                    uint8_t *out = base + o;
                    inf->reset(h + 12 + xlen, t);
                    if (inf->run(out, out_end, base + o) != Inflater::kStreamEnd || (size_t)(out - (base + o)) != want_len ||
                        (uint32_t)crc32_z(crc32(0L, Z_NULL, 0), base + o, want_len) != want_crc) { ok = false; break; }
                    o += want_len;
                }
                if (!ok || o != total) fail->raise("corrupt BGZF member");
                g->out->n = ok ? o : 0;
                { std::lock_guard<std::mutex> l(g->m); g->ready = true; }
                g->cv.notify_all();
            });
            if ((int)flight.size() >= kBgzfWindow && !drain_one()) { while (!flight.empty()) { auto f = flight.front(); flight.pop_front(); std::unique_lock<std::mutex> l(f->m); f->cv.wait(l, [&]() { return f->ready; }); } return; }
        }
        while (!flight.empty()) if (!drain_one()) { while (!flight.empty()) { auto f = flight.front(); flight.pop_front(); std::unique_lock<std::mutex> l(f->m); f->cv.wait(l, [&]() { return f->ready; }); } return; }
    }

    FILE *fp_ = nullptr;
    int kind_ = 0;                                              // 0 text, 1 gzip stream, 2 BGZF
    WorkPool *pool_ = nullptr;
    std::thread producer_;
    std::mutex qm_; std::condition_variable qcv_;
    std::deque<BlockPtr> queue_;
    bool stop_ = false, done_ = false;
    BlockPtr cur_; size_t pos_ = 0;
    std::string carry_;
    Failure fail_;
};

}  // namespace fqb

struct fqb_feeder { fqb::Feeder f; };

extern "C" {

int fqb_feeder_open(const char *path, int n_threads, fqb_feeder **out) {
    if (!path || !out) { fqb::set_error("null argument"); return FQB_ERR_ARG; }
    fqb_feeder *f = new fqb_feeder();
    std::string err;
    if (!f->f.open(path, n_threads, err)) { fqb::set_error(err); delete f; return FQB_ERR_IO; }
    *out = f;
    return FQB_OK;
}

int fqb_feeder_format(const fqb_feeder *f) { return f ? f->f.kind() : -1; }

int64_t fqb_feeder_fill(fqb_feeder *f, int32_t n_max, int32_t stride, uint8_t *bases, uint8_t *quals, int32_t *lens, char *names, int32_t name_stride) {
    if (!f || !bases || !quals || !lens || !names || n_max < 0 || stride < 1 || name_stride < 2) { fqb::set_error("bad argument"); return -1; }
    fqb::Batch B{stride, name_stride, bases, quals, lens, names};
    std::string err;
    const int n = f->f.fill(n_max, B, err);
    if (n < 0) fqb::set_error(err);
    return n;
}

int64_t fqb_feeder_fill_packed(fqb_feeder *f, int32_t n_max, int32_t stride, uint8_t *bases, uint8_t *quals, int32_t *lens, char *names, int32_t name_stride,
                               int32_t packed_stride, uint8_t *packed, uint8_t *quals_flagged) {
    if (!f || !bases || !quals || !lens || !names || n_max < 0 || stride < 1 || name_stride < 2 || !packed || !quals_flagged ||
        packed_stride != ((stride + 63) / 64) * 16) { fqb::set_error("bad argument (packed_stride must be fqb_packed_stride(stride))"); return -1; }
    fqb::Batch B{stride, name_stride, bases, quals, lens, names};
    B.packed_stride = packed_stride; B.packed = packed; B.qflag = quals_flagged;
    std::string err;
    const int n = f->f.fill(n_max, B, err);
    if (n < 0) fqb::set_error(err);
    return n;
}

void fqb_feeder_close(fqb_feeder *f) { delete f; }

int fqb_gunzip(const uint8_t *gz, int64_t n_gz, uint8_t *out, int64_t cap, int32_t block_bytes, int64_t *n_out) {
    if (!gz || n_gz < 0 || (!out && cap > 0) || cap < 0 || !n_out) { fqb::set_error("bad argument"); return FQB_ERR_ARG; }
    int64_t n = 0;
    bool fits = true;
    std::string err;
    const bool ok = fqb::gunzip_blocks(gz, gz + n_gz, block_bytes > 0 ? (size_t)block_bytes : fqb::kBlockBytes, [&](fqb::BlockPtr b) {
        if (n + (int64_t)b->n > cap) { fits = false; return false; }
        memcpy(out + n, b->text.data() + b->off, b->n);
        n += (int64_t)b->n;
        return true;
    }, nullptr, err);
    *n_out = n;
    if (!ok) { fqb::set_error(fits ? err : "output buffer too small"); return fits ? FQB_ERR_IO : FQB_ERR_ARG; }
    return FQB_OK;
}

}  // extern "C"
