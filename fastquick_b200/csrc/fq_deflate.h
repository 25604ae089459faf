// One-shot raw DEFLATE compressor for the BGZF members of the BAM writer (row f1); see fq_deflate.cpp.
#pragma once
#include <cstddef>
#include <cstdint>

namespace fqb {

// room deflate_fast needs for n input bytes: the stored form (n + 5) plus the bit writer's 8-byte stores
inline size_t deflate_fast_bound(size_t n) { return n + 32; }

// in[0..n), n <= 65535 -> one complete raw deflate stream (a single final block) in out; returns its size, 0 if n is too
// large or cap < deflate_fast_bound(n).  Never larger than n + 5.
size_t deflate_fast(const uint8_t *in, size_t n, uint8_t *out, size_t cap);

}  // namespace fqb
