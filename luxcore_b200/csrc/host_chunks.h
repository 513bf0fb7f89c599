// host_chunks.h -- how a batch that arrives from host memory is cut into pieces for the copy / trace / copy pipeline
// (lrb_trace_host and the pipelined plugin sequence lrb_h2d -> lrb_trace -> lrb_d2h, device.cu).
//
// The pipeline is bound by the host->device copies (48 B per ray against 20 B back); what it adds to their time is the
// DRAIN: the trace and the read-back of the LAST piece, which nothing overlaps.  Uniform pieces of 1 Mi rays leave about
// 1 ms of a 16 ms batch there.  So the tail shrinks geometrically: full pieces while more than two of them remain, then
// every piece is half of what is left, down to `minChunk`; the last piece -- the drain -- is `minChunk` rays at most.
// Pure integer arithmetic, no CUDA: also compiled into the CPU tests (tests/test_host_chunks_cpu.py).
#ifndef LRB_HOST_CHUNKS_H
#define LRB_HOST_CHUNKS_H

#include <stdint.h>

#include <vector>

namespace lrb {

// Exclusive end of every piece, in units (rays), for n units.  Pieces tile [0, n) in order, none is empty.
inline void ChunkEnds(const uint64_t n, uint64_t chunk, uint64_t minChunk, const bool taper, std::vector<uint64_t> *ends) {
	ends->clear();
	if (chunk == 0) chunk = 1;
	if (minChunk == 0) minChunk = 1;
	if (minChunk > chunk) minChunk = chunk;
	uint64_t pos = 0;
	while (pos < n) {
		const uint64_t rem = n - pos;
		uint64_t c = chunk;
		if (taper && n > chunk && rem <= 2 * chunk) {
			c = (rem / 2 + 1023) & ~(uint64_t)1023;     // half of what is left, in whole kibi-rays
			if (c < minChunk) c = minChunk;
			if (c > chunk) c = chunk;
			if (rem <= minChunk) c = rem;
		}
		if (c > rem) c = rem;
		pos += c;
		ends->push_back(pos);
	}
}

}   // namespace lrb

#endif
