#include "fq_relayout.h"

namespace fqb {

void relayout_bwt(const HostBwt &b, std::vector<Block32> &out) {
    const uint32_t n = b.seq_len;
    const uint32_t n_blocks = n / 64 + 2;
    out.assign(n_blocks, Block32{{0, 0, 0, 0}, {0, 0, 0, 0}});
    uint32_t run[4] = {0, 0, 0, 0};
    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        Block32 &o = out[blk];
        for (int c = 0; c < 4; ++c) o.cnt[c] = run[c];
        for (uint32_t j = 0; j < 64; ++j) {
            uint32_t pos = blk * 64 + j;
            if (pos >= n) break;
            // stored symbol `pos` in the reference layout: word (pos/128)*12 + 4 + (pos%128)/16
            uint32_t word = b.bwt[(size_t)(pos / 128) * 12 + 4 + (pos % 128) / 16];
            uint32_t sym = (word >> ((15u - (pos & 15u)) << 1)) & 3u;
            // two bit planes of 64 bits each: bases[0..1] = high bits, bases[2..3] = low bits, symbol j at bit 31 - (j & 31)
            o.bases[(j >> 5)] |= (sym >> 1) << (31u - (j & 31u));
            o.bases[2 + (j >> 5)] |= (sym & 1u) << (31u - (j & 31u));
            ++run[sym];
        }
    }
}

}  // namespace fqb
