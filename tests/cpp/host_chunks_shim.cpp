// TEST SCAFFOLDING: exposes luxcore_b200/csrc/host_chunks.h (the chunk schedule of the host pipelines) to ctypes.
#include "host_chunks.h"

extern "C" uint32_t hc_chunk_ends(uint64_t n, uint64_t chunk, uint64_t minChunk, int taper, uint64_t *out, uint32_t cap) {
	std::vector<uint64_t> ends;
	lrb::ChunkEnds(n, chunk, minChunk, taper != 0, &ends);
	for (size_t i = 0; i < ends.size() && i < cap; ++i)
		out[i] = ends[i];
	return (uint32_t)ends.size();
}
