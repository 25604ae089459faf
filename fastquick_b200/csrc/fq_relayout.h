// Host-side conversion of the reference's on-disk FM index (48-byte blocks per
// 128 symbols, libbwa/bwt.h:34,56-62) into the 32-byte-per-64-symbols blocks the
// kernels read (fq_device_core.cuh: DevBwt).
#pragma once
#include <cstdint>
#include <vector>
#include "fq_index.h"

namespace fqb {
struct Block32 { uint32_t cnt[4]; uint32_t bases[4]; };
static_assert(sizeof(Block32) == 32, "one L2 sector per block");
void relayout_bwt(const HostBwt &b, std::vector<Block32> &out);
}  // namespace fqb
