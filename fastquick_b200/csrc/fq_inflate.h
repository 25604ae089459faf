// Raw DEFLATE (RFC 1951) decoder for the FASTQ feeder (row f2).  A gzip stream is serial by construction, so for the
// usual `.fq.gz` the inflate loop IS the feeder's ceiling; this one decodes from a contiguous (memory-mapped) input with
// a 64-bit bit buffer refilled by one unaligned load, 11-bit / 8-bit first-level tables and word-wise match copies, and
// can stop and resume between any two symbols, which is what lets it fill the feeder's text blocks one after another.
// Every member's CRC-32 and length are checked by the caller (fq_feeder.cpp), so a decoding fault cannot pass silently.
#pragma once
#include <cstddef>
#include <cstdint>

namespace fqb {

class Inflater {
public:
    enum Status { kOutputFull, kStreamEnd, kError };
    static constexpr size_t kWindow = 32768;        // history a match may reach back into
    static constexpr size_t kOutSlack = 320;        // inside a Huffman block run() returns kOutputFull once fewer bytes than
                                                    // this are left (a match and its over-copy fit); buffers must be larger

    void reset(const uint8_t *in, const uint8_t *in_end);
    // Decodes into [out, out_end), advancing out.  floor = lowest address a match may copy from (the bytes between
    // floor and out must be the previously decoded output, at least kWindow of it once that much exists).
    Status run(uint8_t *&out, uint8_t *out_end, const uint8_t *floor);
    const uint8_t *in_pos() const { return in_ - (bc_ >> 3); }      // after kStreamEnd: first byte behind the stream
    const char *error() const { return err_; }

private:
    enum State { kHeader, kStored, kHuff, kDone, kFailed };
    enum Step { kBlockEnd, kNeedSpace, kInputLow, kFault };
    static constexpr int kLitBits = 11, kDistBits = 8;
    static constexpr int kLitCap = 2048 + 2560, kDistCap = 256 + 1280;

    bool read_header();
    bool build(const uint8_t *lens, int n, int root, uint32_t *tab, int cap, int kind);
    Step huff_fast(uint8_t *&out, uint8_t *out_end, const uint8_t *floor);
    Step huff_tail(uint8_t *&out, uint8_t *out_end, const uint8_t *floor);
    bool need(unsigned n);                          // byte-wise refill; false = fewer than n bits are left
    void drop(unsigned n) { bb_ >>= n; bc_ -= n; }
    Status fail(const char *m) { err_ = m; state_ = kFailed; return kError; }

    const uint8_t *in_ = nullptr, *in_end_ = nullptr;
    uint64_t bb_ = 0; unsigned bc_ = 0;
    State state_ = kHeader;
    bool last_ = false;
    uint32_t stored_left_ = 0;
    const char *err_ = "";
    uint32_t lit_[kLitCap], dist_[kDistCap];
};

}  // namespace fqb
