// Raw DEFLATE decoder of the FASTQ feeder (see fq_inflate.h).  Written from RFC 1951; the table layout (one 32-bit
// entry per prefix: value | flags | bits to drop) and the "load eight bytes, count whole ones" refill are the usual
// ones of table-driven decoders.  The reference reads its FASTQ through zlib's gzread (libbwa/kseq.h:327-370 via
// src/BwtMapper.cpp:476-613); what has to match is the decoded text, which the member CRC-32 pins.
#include "fq_inflate.h"

#include <cstring>

namespace fqb {
namespace {

// entry = value << 16 | flags << 8 | bits to drop.  Length / distance entries drop their extra bits together with the
// code and keep the code's own length in the low nibble of flags (value + the dropped bits shifted right by it is the
// length / distance); on a literal entry that nibble is 1 when the entry carries two bytes.
constexpr uint32_t kLit = 0x8000, kEob = 0x4000, kSub = 0x2000, kBad = 0x1000;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }      // little-endian hosts only (x86-64 / aarch64)

inline uint32_t symbol_entry(int sym, int kind) {
    if (kind == 0) return (uint32_t)sym << 16;
    if (kind == 1) {
        if (sym < 256) return ((uint32_t)sym << 16) | kLit;
        if (sym == 256) return kEob;
        if (sym < 286) return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8);
        return kBad;
    }
    if (sym < 30) return ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8);
    return kBad;
}

}  // namespace

void Inflater::reset(const uint8_t *in, const uint8_t *in_end) {
    in_ = in; in_end_ = in_end; bb_ = 0; bc_ = 0; state_ = kHeader; last_ = false; stored_left_ = 0; err_ = "";
}

bool Inflater::need(unsigned n) {
    while (bc_ < 56 && in_ < in_end_) { bb_ |= (uint64_t)*in_++ << bc_; bc_ += 8; }
    return bc_ >= n;
}

// Canonical Huffman code -> lookup table indexed by the next `root` bits (LSB first); longer codes go through one
// second-level table per first-level prefix.  kind: 0 code-length code, 1 literal/length, 2 distance.  An incomplete code
// is accepted (its unused prefixes fault when met); an over-subscribed one is refused.
bool Inflater::build(const uint8_t *lens, int n, int root, uint32_t *tab, int cap, int kind) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; }
    uint32_t next[16]; uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    const int size = 1 << root;
    uint16_t rev[288]; uint8_t submax[1 << kLitBits];
    memset(submax, 0, (size_t)size);
    for (int i = 0; i < size; ++i) tab[i] = kBad | 1;
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t c = next[l]++, r = 0;
        for (int b = 0; b < l; ++b) { r = (r << 1) | (c & 1); c >>= 1; }
        rev[s] = (uint16_t)r;
        if (l > root) { uint8_t &m = submax[r & (uint32_t)(size - 1)]; if (l > m) m = (uint8_t)l; }
    }
    int used = size;
    for (int p = 0; p < size; ++p) {
        if (!submax[p]) continue;
        const int sb = submax[p] - root;
        if (used + (1 << sb) > cap) return false;
        tab[p] = ((uint32_t)used << 16) | kSub | (uint32_t)sb;
        for (int k = 0; k < (1 << sb); ++k) tab[used + k] = kBad | 1;
        used += 1 << sb;
    }
    // final form of an entry whose code takes cl bits at this level: a length / distance entry drops its extra bits
    // together with the code (low byte = cl + extra) and keeps cl in the flag nibble to shift the code away
    auto finish = [kind](uint32_t e, int cl) -> uint32_t {
        if (kind == 0 || (e & (kLit | kEob | kBad))) return e | (uint32_t)cl;
        return (e & ~0xf00u) | ((uint32_t)cl << 8) | ((uint32_t)cl + ((e >> 8) & 15));
    };
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t e = symbol_entry(s, kind);
        if (l <= root) {
            for (int k = rev[s]; k < size; k += 1 << l) tab[k] = finish(e, l);
        } else {
            const uint32_t sub = tab[rev[s] & (uint32_t)(size - 1)];
            const int off = (int)(sub >> 16), sb = (int)(sub & 0xff);
            for (int k = rev[s] >> root; k < (1 << sb); k += 1 << (l - root)) tab[off + k] = finish(e, l - root);
        }
    }
    if (kind == 1) {
        // Two literals behind one lookup: where a literal's code leaves enough index bits for a second literal's whole
        // code, the entry carries both bytes (value = first | second << 8, flag bit 0x100) and the summed length.
        // The second entry is read at index i >> l1 < i, so walking down from the top only ever reads single entries.
        for (int i = size - 1; i > 0; --i) {
            const uint32_t e1 = tab[i];
            if (!(e1 & kLit)) continue;
            const int l1 = (int)(e1 & 0xff);
            const uint32_t e2 = tab[i >> l1];
            if (!(e2 & kLit) || (int)(e2 & 0xff) > root - l1) continue;
            tab[i] = (e1 & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8) | kLit | 0x100u | (uint32_t)(l1 + (int)(e2 & 0xff));
        }
    }
    return true;
}

bool Inflater::read_header() {
    if (!need(3)) { fail("truncated deflate stream"); return false; }
    last_ = bb_ & 1;
    const unsigned type = (unsigned)(bb_ >> 1) & 3;
    drop(3);
    if (type == 0) {
        drop(bc_ & 7);
        if (!need(32)) { fail("truncated deflate stream"); return false; }
        const uint32_t len = (uint32_t)bb_ & 0xffff, nlen = (uint32_t)(bb_ >> 16) & 0xffff;
        if ((len ^ 0xffff) != nlen) { fail("stored block length check failed"); return false; }
        drop(32);
        in_ -= bc_ >> 3; bb_ = 0; bc_ = 0;                   // whole bytes still in the bit buffer go back to the input
        stored_left_ = len; state_ = kStored;
        return true;
    }
    uint8_t lens[320];
    if (type == 1) {
        for (int i = 0; i < 144; ++i) lens[i] = 8;
        for (int i = 144; i < 256; ++i) lens[i] = 9;
        for (int i = 256; i < 280; ++i) lens[i] = 7;
        for (int i = 280; i < 288; ++i) lens[i] = 8;
        build(lens, 288, kLitBits, lit_, kLitCap, 1);
        for (int i = 0; i < 32; ++i) lens[i] = 5;
        build(lens, 32, kDistBits, dist_, kDistCap, 2);
        state_ = kHuff;
        return true;
    }
    if (type == 3) { fail("invalid deflate block type"); return false; }
    if (!need(14)) { fail("truncated deflate stream"); return false; }
    const int hlit = (int)(bb_ & 31) + 257, hdist = (int)((bb_ >> 5) & 31) + 1, hclen = (int)((bb_ >> 10) & 15) + 4;
    drop(14);
    if (hlit > 286 || hdist > 30) { fail("too many length or distance symbols"); return false; }
    uint8_t pre_lens[19] = {0};
    for (int i = 0; i < hclen; ++i) {
        if (!need(3)) { fail("truncated deflate stream"); return false; }
        pre_lens[kPreOrder[i]] = (uint8_t)(bb_ & 7);
        drop(3);
    }
    uint32_t pre[128];
    if (!build(pre_lens, 19, 7, pre, 128, 0)) { fail("invalid code-length code"); return false; }
    const int total = hlit + hdist;
    for (int i = 0; i < total;) {
        need(0);
        const uint32_t e = pre[bb_ & 127];
        if (e & kBad) { fail("invalid code-length symbol"); return false; }
        if ((e & 0xff) > bc_) { fail("truncated deflate stream"); return false; }
        drop(e & 0xff);
        const int sym = (int)(e >> 16);
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        const unsigned xb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
        if (!need(xb)) { fail("truncated deflate stream"); return false; }
        int rep = (int)(bb_ & ((1u << xb) - 1)) + (sym == 18 ? 11 : 3);
        drop(xb);
        uint8_t v = 0;
        if (sym == 16) { if (i == 0) { fail("length repeat with no previous length"); return false; } v = lens[i - 1]; }
        if (i + rep > total) { fail("length repeat runs past the code"); return false; }
        while (rep--) lens[i++] = v;
    }
    if (lens[256] == 0) { fail("block has no end-of-block code"); return false; }
    if (!build(lens, hlit, kLitBits, lit_, kLitCap, 1)) { fail("invalid literal/length code"); return false; }
    if (!build(lens + hlit, hdist, kDistBits, dist_, kDistCap, 2)) { fail("invalid distance code"); return false; }
    state_ = kHuff;
    return true;
}

// The symbol loop while at least 32 input bytes and kOutSlack output bytes are left (checked once per round, which
// refills at most three times and writes at most one match plus its over-copy).  The bit buffer is refilled with one
// 8-byte load; bits above the counted ones are the following input bytes, which a later refill ORs in again unchanged.
// The entry of the next symbol is looked up before a match is copied so that the two latencies overlap.
Inflater::Step Inflater::huff_fast(uint8_t *&out_ref, uint8_t *out_end, const uint8_t *floor) {
    const uint8_t *in = in_;
    const uint8_t *const in_stop = in_end_ - 32;
    uint8_t *const out_stop = out_end - kOutSlack;
    uint64_t bb = bb_; unsigned bc = bc_;
    uint8_t *out = out_ref;
    const uint32_t *const lit = lit_, *const dtab = dist_;
    Step result;
#define FQB_REFILL() do { bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56; } while (0)
#define FQB_LOOKUP() lit[bb & ((1u << kLitBits) - 1)]
#define FQB_PUT() do { const uint16_t v_ = (uint16_t)(e >> 16); memcpy(out, &v_, 2); out += 1 + ((e >> 8) & 1); } while (0)
#define FQB_LIT() do { bb >>= e & 0xff; bc -= e & 0xff; FQB_PUT(); } while (0)
#define FQB_DROP(n) do { const unsigned n_ = (n); bb >>= n_; bc -= n_; } while (0)
#define FQB_BOUNDS() if (out > out_stop) { result = kNeedSpace; break; } if (in > in_stop) { result = kInputLow; break; }
    if (out > out_stop) return kNeedSpace;
    if (in > in_stop) return kInputLow;
    FQB_REFILL();
    uint32_t e = FQB_LOOKUP();
    for (;;) {
        if (e & kLit) {                                       // up to three literal entries per refill (3 x 15 bits <= 56)
            FQB_LIT();
            e = FQB_LOOKUP();
            if (e & kLit) {
                FQB_LIT();
                e = FQB_LOOKUP();
                if (e & kLit) {
                    FQB_LIT();
                    FQB_BOUNDS();
                    FQB_REFILL();
                    e = FQB_LOOKUP();
                    continue;
                }
            }
            FQB_REFILL();
        }
        if (e & kSub) {
            FQB_DROP(kLitBits);
            e = lit[(e >> 16) + (bb & ((1u << (e & 0xff)) - 1))];
        }
        uint32_t saved = (uint32_t)bb;
        FQB_DROP(e & 0xff);
        if (e & (kLit | kEob | kBad)) {
            if (e & kLit) {                                   // a literal with a long code
                FQB_PUT();
                FQB_BOUNDS();
                FQB_REFILL();
                e = FQB_LOOKUP();
                continue;
            }
            if (e & kBad) { err_ = "invalid literal/length code in the stream"; result = kFault; } else result = kBlockEnd;
            break;
        }
        const unsigned len = (e >> 16) + ((saved & ((1u << (e & 0xff)) - 1)) >> ((e >> 8) & 15));
        e = dtab[bb & ((1u << kDistBits) - 1)];
        if (e & kSub) {
            FQB_DROP(kDistBits);
            e = dtab[(e >> 16) + (bb & ((1u << (e & 0xff)) - 1))];
        }
        saved = (uint32_t)bb;
        FQB_DROP(e & 0xff);
        if (e & kBad) { err_ = "invalid distance code in the stream"; result = kFault; break; }
        const size_t dist = (e >> 16) + (size_t)((saved & ((1u << (e & 0xff)) - 1)) >> ((e >> 8) & 15));
        if (dist > (size_t)(out - floor)) { err_ = "match distance reaches before the start of the output"; result = kFault; break; }
        FQB_REFILL();
        e = FQB_LOOKUP();
        const uint8_t *src = out - dist;
        uint8_t *const end = out + len;
        if (dist >= 8) {                                      // may write up to 7 bytes past end: inside kOutSlack
            memcpy(out, src, 8);
            if (len > 8) {
                memcpy(out + 8, src + 8, 8);
                if (len > 16) { out += 16; src += 16; do { memcpy(out, src, 8); out += 8; src += 8; } while (out < end); }
            }
        } else if (dist == 1) {
            memset(out, *src, len);
        } else {
            do { *out++ = *src++; } while (out < end);
        }
        out = end;
        FQB_BOUNDS();
    }
#undef FQB_REFILL
#undef FQB_LOOKUP
#undef FQB_PUT
#undef FQB_LIT
#undef FQB_DROP
#undef FQB_BOUNDS
    in_ = in; bb_ = bb; bc_ = bc; out_ref = out;
    return result;
}

// The same loop for the last bytes of the input: bytes are added one by one and every drop is checked against the bits
// that exist (past the end of the input the buffer reads as zeros).
Inflater::Step Inflater::huff_tail(uint8_t *&out_ref, uint8_t *out_end, const uint8_t *floor) {
    const uint8_t *in = in_;
    uint64_t bb = bb_; unsigned bc = bc_;
    uint8_t *out = out_ref;
    Step result;
#define FQB_FILL() while (bc < 56 && in < in_end_) { bb |= (uint64_t)*in++ << bc; bc += 8; }
#define FQB_DROP(n) do { const unsigned n_ = (n); if (n_ > bc) { err_ = "truncated deflate stream"; result = kFault; goto leave; } bb >>= n_; bc -= n_; } while (0)
    for (;;) {
        if ((size_t)(out_end - out) < kOutSlack) { result = kNeedSpace; break; }
        FQB_FILL();
        uint32_t e = lit_[bb & ((1u << kLitBits) - 1)];
        if (e & kSub) {
            FQB_DROP(kLitBits);
            e = lit_[(e >> 16) + (bb & ((1u << (e & 0xff)) - 1))];
        }
        uint32_t saved = (uint32_t)bb;
        FQB_DROP(e & 0xff);
        if (e & kLit) { *out++ = (uint8_t)(e >> 16); if (e & 0x100) *out++ = (uint8_t)(e >> 24); continue; }
        if (e & (kEob | kBad)) {
            if (e & kBad) { err_ = "invalid literal/length code in the stream"; result = kFault; } else result = kBlockEnd;
            break;
        }
        const unsigned len = (e >> 16) + ((saved & ((1u << (e & 0xff)) - 1)) >> ((e >> 8) & 15));
        FQB_FILL();
        e = dist_[bb & ((1u << kDistBits) - 1)];
        if (e & kSub) {
            FQB_DROP(kDistBits);
            e = dist_[(e >> 16) + (bb & ((1u << (e & 0xff)) - 1))];
        }
        saved = (uint32_t)bb;
        FQB_DROP(e & 0xff);
        if (e & kBad) { err_ = "invalid distance code in the stream"; result = kFault; break; }
        const size_t dist = (e >> 16) + (size_t)((saved & ((1u << (e & 0xff)) - 1)) >> ((e >> 8) & 15));
        if (dist > (size_t)(out - floor)) { err_ = "match distance reaches before the start of the output"; result = kFault; break; }
        for (unsigned k = 0; k < len; ++k) out[k] = out[(ptrdiff_t)k - (ptrdiff_t)dist];
        out += len;
    }
leave:
#undef FQB_FILL
#undef FQB_DROP
    in_ = in; bb_ = bb; bc_ = bc; out_ref = out;
    return result;
}

Inflater::Status Inflater::run(uint8_t *&out, uint8_t *out_end, const uint8_t *floor) {
    for (;;) {
        switch (state_) {
        case kHeader:
            if (!read_header()) return kError;
            break;
        case kStored: {
            if (stored_left_ == 0) { state_ = last_ ? kDone : kHeader; break; }
            const size_t room = (size_t)(out_end - out), avail = (size_t)(in_end_ - in_);
            if (room == 0) return kOutputFull;
            size_t n = stored_left_ < room ? stored_left_ : room;
            if (n > avail) n = avail;
            if (n == 0) return fail("truncated deflate stream");
            memcpy(out, in_, n);
            out += n; in_ += n; stored_left_ -= (uint32_t)n;
            break;
        }
        case kHuff: {
            const Step s = (in_end_ - in_ >= 32) ? huff_fast(out, out_end, floor) : huff_tail(out, out_end, floor);
            if (s == kBlockEnd) state_ = last_ ? kDone : kHeader;
            else if (s == kNeedSpace) return kOutputFull;
            else if (s == kFault) { state_ = kFailed; return kError; }
            break;                                            // kInputLow: next round takes the checked loop
        }
        case kDone:
            drop(bc_ & 7);
            return kStreamEnd;
        case kFailed:
            return kError;
        }
    }
}

}  // namespace fqb
