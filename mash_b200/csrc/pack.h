// pack.h -- host feed path (see pack.cpp)
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace mashgpu {

struct PackSegment { const uint8_t *src; uint64_t pos, len; };   // a record placed at stream positions [pos, pos+len)
struct PackRun { uint64_t start, len; };                          // invalid positions [start, start+len)

// codes must hold ceil(stream_len / 32) uint64 words.  Segments sorted by pos, non-overlapping; positions not covered by
// any segment are separators.  `threads` host threads are used (>= 1).
void pack_stream(const PackSegment *segments, size_t n_segments, uint64_t stream_len, int preserve_case, int threads,
                 uint64_t *codes, std::vector<PackRun> &runs);

}  // namespace mashgpu
