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

// One contiguous chunk (reads separated by bytes outside the alphabet, the screen feed): codes as above and the invalid positions
// as a bit mask, 32 positions per word, instead of runs -- a separator every 150 bases would make a run per read.  Both arrays hold
// ceil(len / 32) words; positions past `len` in the last word are invalid.
void pack_chunk_mask(const uint8_t *src, uint64_t len, int preserve_case, int threads, uint64_t *codes, uint32_t *inval);

// Packer threads the process can really run: hardware threads capped by the container's CPU quota (cgroup v2 cpu.max) and by
// MASHGPU_PACK_THREADS.
int host_pack_threads();

}  // namespace mashgpu
