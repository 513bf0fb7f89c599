// device.cu -- implementation of the C ABI in include/luxrays_b200.h for sm_100a.
//
// Replaces, behind one flat interface, the reference's CUDADevice buffer/queue API
// (src/luxrays/devices/cudadevice.cpp:407-540), the BVHKernel / MBVHKernel upload paths
// (bvhaccelhw.cpp:38-237, mbvhaccelhw.cpp:41-306,308-466) and their kernel launches
// (bvhaccelhw.cpp:259-268, mbvhaccelhw.cpp:468-507).  No NVRTC, no cuew, no OptiX.

#include <cuda.h>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <stdexcept>
#include <string>
#include <vector>

#include "luxrays_b200.h"
#include "layout.h"
#include "relayout.h"
#include "host_chunks.h"
#include "trace_kernels.cuh"
#include "batch_kernels.cuh"
#include "build_kernels.cuh"
#include "relayout_kernels.cuh"

using namespace lrb;

static thread_local std::string g_lastError;

static int Fail(int code, const std::string &msg) {
	g_lastError = msg;
	return code;
}

#define LRB_CUDA(call)                                                                              \
	do {                                                                                            \
		const cudaError_t e_ = (call);                                                              \
		if (e_ != cudaSuccess) {                                                                    \
			const int code_ = (e_ == cudaErrorMemoryAllocation) ? LRB_ERR_OOM :                     \
					((e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? LRB_ERR_NO_DEVICE : LRB_ERR_CUDA); \
			return Fail(code_, std::string(#call) + " failed: " + cudaGetErrorString(e_) +        \
					" (" __FILE__ ":" + std::to_string(__LINE__) + ")");                            \
		}                                                                                           \
	} while (0)

struct lrb_device {
	int ordinal;
	cudaStream_t ownStream;
	cudaStream_t stream;            // the in-order queue in use (own or adopted)
	cudaStream_t copyInStream, copyOutStream;   // lrb_trace_host pipeline
	cudaDeviceProp prop;
	lrb_counters counters;
	std::mutex mtx;
	std::unordered_map<void *, size_t> allocs;  // lrb_alloc bookkeeping (GetUsedMemory parity)
	// options
	int blocksPerSM;                // 0 = from occupancy
	int persistent;                 // 1 = TracePersistent, 0 = TraceStatic
	int smemDepth;                  // shared-memory stack entries per thread
	int refillBelow;
	int instBias;                   // two-level scenes: see TraceArgs::instBias
	int triBias;
	int gatherStores;               // 1: lrb_trace_gather(n_chunks = 0) uses dual-destination stores instead of signalled DMA pushes
	int gatherChunkShift;           // log2(rays per signalled chunk)
	int gatherDefer;                // 1: lrb_trace_gather does not make the queue wait for its pushes (see lrb_gather_wait)
	cudaEvent_t gatherDone[2];      // end of the pushes of the last two lrb_trace_gather calls (alternating)
	bool gatherPending[2];
	unsigned gatherSeq;
	int wideStores;                 // bit 0: vector RayHit stores to the local buffer, bit 1: to the peer buffer
	int compact;                    // 1: lrb_trace first builds the dense list of live (non-masked) rays and traces through it
	uint32_t *compactIdx, *compactBlocks, *compactTotal;    // scratch of the compaction kernels
	size_t compactCap;
	int carveout;                   // preferred shared-memory carve-out of the trace kernels in percent (-1 = driver default)
	int l2Persist;                  // L2 persistence window over the scene's nodes + triangles: 0 never, 1 always (if it fits), 2 = when it fits the set-aside
	const void *l2WindowBase;       // window currently set on the queue (stream attribute), or NULL
	size_t l2WindowBytes;
	int prefetch;                   // L2 prefetch of the children pushed on the stack: 0 never, 1 always, 2 when the scene does not fit L2
	int prefetchMode;               // what the prefetching kernels fetch ahead: bit 0 pushed children -> L2, bit 1 nearest child -> L2, bit 2 nearest child -> L1
	int sortRays;                   // order the rays of a batch for coherence before tracing them: 0 never, 1 always, 2 when the scene does not fit L2
	int sortBitsPerAxis;            // origin-cell resolution of the sort key
	int sortMinRays;                // batches smaller than this are traced in index order
	// scratch of the ray-ordering pre-pass (keys / indices, double-buffered, + CUB temp storage)
	uint32_t *sortKeys[2], *sortVals[2];
	void *sortTemp;
	size_t sortCap, sortTempBytes;
	int hostChunk;                  // rays per chunk in lrb_trace_host
	int hostTaper;                  // 1: the last chunks shrink geometrically down to hostMinChunk rays (host_chunks.h), 0: uniform chunks
	int hostMinChunk;
	// The reference-facing call sequence (AllocBufferRW(src) / EnqueueTraceRayBuffer / EnqueueReadBuffer / FinishQueue =
	// lrb_h2d / lrb_trace / lrb_d2h / lrb_sync) pipelined behind its own interface: a large asynchronous upload is cut
	// into chunks on the copy-in stream, a trace whose ray buffer is exactly that upload follows it chunk by chunk, a
	// read of exactly that trace's RayHit buffer follows the trace chunk by chunk on the copy-out stream.  Anything
	// else first joins the queue with what is pending (JoinPending), so the in-order semantics of the queue hold.
	int pipeline;                   // option "pipeline": 1 (default) = as described, 0 = every call on the queue as it is
	struct Pending {
		const char *base;           // device range [base, base + bytes)
		size_t bytes;
		std::vector<uint64_t> ends;     // exclusive end of every chunk in BYTES of this range (host_chunks.h: the tail tapers)
		std::vector<cudaEvent_t> ev;    // one per chunk (borrowed from pipeEvents)
		bool active;
	} pendUpload[2], pendTraced;
	std::vector<cudaEvent_t> pipeEvents;
	size_t pipeEventNext;
	cudaEvent_t pipeJoin;
	bool copyOutBusy;               // the copy-out stream holds chunked reads the queue has not joined yet
	uint32_t *signalValues;         // lrb_gather_signal: table of step values in device memory
	// staging for lrb_trace_host
	void *stageRays, *stageHits;
	size_t stageRaysBytes, stageHitsBytes;
	std::vector<cudaEvent_t> events;
};

struct lrb_scene {
	lrb_device *dev;
	WideScene host;                 // kept for MBVH (Update); cleared for single-level scenes
	WideNode *dNodes;
	TriRecord *dTris;
	TriIds *dIds;
	void *dSlab;                    // one-level scenes: nodes + triangle records + ids in ONE allocation (the L2 persistence window covers it)
	size_t slabBytes;
	InstRecord *dInsts;
	float *dMinv;
	uint32_t *dMotionFirst, *dMotionLast;
	DevInterp *dInterps;
	size_t capNodes, capInsts;      // allocated element counts (Update re-uses them)
	uint32_t *dCounter;
	uint32_t *dSpillNode;
	float *dSpillT;
	size_t spillEntries;
	TraceStats *dStats;
	uint32_t *dWatermark, *dChunkFlag;  // signalled gather: per-warp watermarks / completion flags per chunk
	size_t chunkCap, watermarkCap;
	uint32_t epoch;
	SceneView view;
	lrb_scene_info info;
};

static int SetDev(lrb_device *dev) {
	if (!dev)
		return Fail(LRB_ERR_INVALID, "null device");
	LRB_CUDA(cudaSetDevice(dev->ordinal));
	return LRB_OK;
}

#define LRB_SETDEV(dev)                         \
	do {                                        \
		const int rc_ = SetDev(dev);            \
		if (rc_ != LRB_OK) return rc_;          \
	} while (0)

// ---- the plugin sequence, pipelined (see lrb_device::pipeline) ----------------------------------------------

static const size_t kPipeMinBytes = 32u << 20;     // smaller transfers are not worth cutting up

static int PipeEvent(lrb_device *dev, cudaEvent_t *out) {
	if (dev->pipeEvents.size() < 1024) {
		cudaEvent_t e;
		LRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		dev->pipeEvents.push_back(e);
		*out = e;
		return LRB_OK;
	}
	*out = dev->pipeEvents[dev->pipeEventNext++ % dev->pipeEvents.size()];     // (an event that old has long been consumed)
	return LRB_OK;
}

// Everything the side streams still hold becomes part of the device's in-order queue.
static int JoinPending(lrb_device *dev) {
	for (int k = 0; k < 2; ++k) {
		lrb_device::Pending &u = dev->pendUpload[k];
		if (u.active && !u.ev.empty())
			LRB_CUDA(cudaStreamWaitEvent(dev->stream, u.ev.back(), 0));
		u.active = false;
	}
	dev->pendTraced.active = false;     // its chunk events were recorded on the queue itself: nothing to wait for
	if (dev->copyOutBusy) {
		if (!dev->pipeJoin)
			LRB_CUDA(cudaEventCreateWithFlags(&dev->pipeJoin, cudaEventDisableTiming));
		LRB_CUDA(cudaEventRecord(dev->pipeJoin, dev->copyOutStream));
		LRB_CUDA(cudaStreamWaitEvent(dev->stream, dev->pipeJoin, 0));
		dev->copyOutBusy = false;
	dev->signalValues = nullptr;
	}
	return LRB_OK;
}

// Asynchronous upload in chunks on the copy-in stream, remembered as pending.
static int PipelinedUpload(lrb_device *dev, void *dst, const void *src, size_t bytes) {
	int slot = dev->pendUpload[0].active ? 1 : 0;
	if (dev->pendUpload[slot].active) {
		const int rc = JoinPending(dev);
		if (rc != LRB_OK) return rc;
		slot = 0;
	}
	lrb_device::Pending &u = dev->pendUpload[slot];
	// the copy-in stream starts behind everything queued so far (earlier users of dst)
	cudaEvent_t start;
	int rc = PipeEvent(dev, &start);
	if (rc != LRB_OK) return rc;
	LRB_CUDA(cudaEventRecord(start, dev->stream));
	LRB_CUDA(cudaStreamWaitEvent(dev->copyInStream, start, 0));
	// chunk boundaries in whole rays (48 B: a ray buffer cut this way is traced chunk by chunk, lrb_trace); a remainder
	// that is not a whole ray -- some other kind of buffer -- travels with the last chunk
	u.base = (const char *)dst; u.bytes = bytes; u.ev.clear();
	ChunkEnds(bytes / sizeof(lrb_ray), (uint64_t)dev->hostChunk, (uint64_t)dev->hostMinChunk, dev->hostTaper != 0, &u.ends);
	for (uint64_t &e : u.ends)
		e *= sizeof(lrb_ray);
	if (u.ends.empty())
		u.ends.push_back(bytes);
	u.ends.back() = bytes;
	size_t off = 0;
	for (size_t c = 0; c < u.ends.size(); ++c) {
		const size_t cnt = (size_t)u.ends[c] - off;
		LRB_CUDA(cudaMemcpyAsync((char *)dst + off, (const char *)src + off, cnt, cudaMemcpyHostToDevice, dev->copyInStream));
		cudaEvent_t e;
		if ((rc = PipeEvent(dev, &e)) != LRB_OK) return rc;
		LRB_CUDA(cudaEventRecord(e, dev->copyInStream));
		u.ev.push_back(e);
		off = (size_t)u.ends[c];
	}
	u.active = true;
	return LRB_OK;
}

extern "C" {

const char *lrb_last_error_string(void) { return g_lastError.c_str(); }

const char *lrb_version_string(void) { return "luxrays_b200 0.1 (sm_100a)"; }

int lrb_device_count(int *count) {
	if (!count)
		return Fail(LRB_ERR_INVALID, "null count");
	*count = 0;
	int n = 0;
	const cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return Fail(LRB_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
	}
	*count = n;
	return LRB_OK;
}

int lrb_device_create(int ordinal, lrb_device **out) {
	if (!out)
		return Fail(LRB_ERR_INVALID, "null out pointer");
	*out = nullptr;
	int n = 0;
	const int rc = lrb_device_count(&n);
	if (rc != LRB_OK)
		return rc;
	if (n == 0)
		return Fail(LRB_ERR_NO_DEVICE, "no CUDA device present; this library has no CPU fallback");
	if (ordinal < 0 || ordinal >= n)
		return Fail(LRB_ERR_INVALID, "CUDA ordinal out of range");
	LRB_CUDA(cudaSetDevice(ordinal));
	lrb_device *dev = new lrb_device();
	dev->ordinal = ordinal;
	memset(&dev->counters, 0, sizeof(dev->counters));
	dev->stageRays = dev->stageHits = nullptr;
	dev->stageRaysBytes = dev->stageHitsBytes = 0;
	cudaError_t e = cudaGetDeviceProperties(&dev->prop, ordinal);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dev->ownStream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dev->copyInStream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dev->copyOutStream, cudaStreamNonBlocking);
	if (e != cudaSuccess) {
		delete dev;
		return Fail(LRB_ERR_CUDA, std::string("device initialisation failed: ") + cudaGetErrorString(e));
	}
	dev->stream = dev->ownStream;
	dev->blocksPerSM = 0;
	dev->persistent = 1;
	dev->smemDepth = 16;
	dev->refillBelow = 24;
	dev->triBias = 8;
	dev->instBias = 8;
	dev->sortRays = 2;
	dev->prefetch = 0;      // prepared, not yet measured on a GPU: off
	dev->carveout = -1;
	dev->prefetchMode = 1;
	dev->l2Persist = 0;     // measured on B200 (profiles/r02_measure_ingest_c10.json): no gain alone, no protection against concurrent traffic
	dev->l2WindowBase = nullptr;
	dev->l2WindowBytes = 0;
	dev->compact = 0;
	dev->compactIdx = dev->compactBlocks = dev->compactTotal = nullptr;
	dev->compactCap = 0;
	dev->wideStores = 2;
	dev->gatherStores = 0;
	dev->gatherChunkShift = 19;
	dev->gatherDefer = 0;
	dev->gatherDone[0] = dev->gatherDone[1] = nullptr;
	dev->gatherPending[0] = dev->gatherPending[1] = false;
	dev->gatherSeq = 0;
	dev->sortBitsPerAxis = 5;
	dev->sortMinRays = 1 << 18;
	dev->sortKeys[0] = dev->sortKeys[1] = dev->sortVals[0] = dev->sortVals[1] = nullptr;
	dev->sortTemp = nullptr;
	dev->sortCap = dev->sortTempBytes = 0;
	dev->hostChunk = 1 << 20;
	dev->hostTaper = 1;
	dev->hostMinChunk = 1 << 16;
	dev->pipeline = 1;
	dev->pendUpload[0].active = dev->pendUpload[1].active = dev->pendTraced.active = false;
	dev->pipeEventNext = 0;
	dev->pipeJoin = nullptr;
	dev->copyOutBusy = false;
	*out = dev;
	return LRB_OK;
}

int lrb_device_destroy(lrb_device *dev) {
	if (!dev)
		return LRB_OK;
	LRB_SETDEV(dev);
	cudaStreamSynchronize(dev->copyInStream);
	cudaStreamSynchronize(dev->copyOutStream);
	cudaStreamSynchronize(dev->stream);
	for (size_t i = 0; i < dev->events.size(); ++i) cudaEventDestroy(dev->events[i]);
	for (size_t i = 0; i < dev->pipeEvents.size(); ++i) cudaEventDestroy(dev->pipeEvents[i]);
	if (dev->pipeJoin) cudaEventDestroy(dev->pipeJoin);
	cudaFree(dev->signalValues);
	for (int i = 0; i < 2; ++i) if (dev->gatherDone[i]) cudaEventDestroy(dev->gatherDone[i]);
	if (dev->stageRays) cudaFree(dev->stageRays);
	if (dev->stageHits) cudaFree(dev->stageHits);
	cudaFree(dev->sortKeys[0]); cudaFree(dev->sortKeys[1]); cudaFree(dev->sortVals[0]); cudaFree(dev->sortVals[1]);
	cudaFree(dev->sortTemp);
	cudaFree(dev->compactIdx); cudaFree(dev->compactBlocks); cudaFree(dev->compactTotal);
	cudaStreamDestroy(dev->copyInStream);
	cudaStreamDestroy(dev->copyOutStream);
	cudaStreamDestroy(dev->ownStream);
	delete dev;
	return LRB_OK;
}

int lrb_device_get_props(lrb_device *dev, lrb_device_props *out) {
	if (!dev || !out)
		return Fail(LRB_ERR_INVALID, "null argument");
	memset(out, 0, sizeof(*out));
	out->cuda_ordinal = dev->ordinal;
	out->cc_major = dev->prop.major;
	out->cc_minor = dev->prop.minor;
	out->sm_count = dev->prop.multiProcessorCount;
	out->l2_bytes = dev->prop.l2CacheSize;
	out->total_mem_bytes = dev->prop.totalGlobalMem;
	strncpy(out->name, dev->prop.name, sizeof(out->name) - 1);
	return LRB_OK;
}

int lrb_device_set_stream(lrb_device *dev, void *s) {
	if (!dev)
		return Fail(LRB_ERR_INVALID, "null device");
	{
		const int rcJoin = JoinPending(dev);    // what the side streams hold joins the queue that is being left
		if (rcJoin != LRB_OK) return rcJoin;
	}
	dev->stream = s ? (cudaStream_t)s : dev->ownStream;
	dev->l2WindowBase = nullptr;        // the access-policy window is an attribute of the queue: set again at the next launch
	dev->l2WindowBytes = 0;
	return LRB_OK;
}

int lrb_device_get_stream(lrb_device *dev, void **s) {
	if (!dev || !s)
		return Fail(LRB_ERR_INVALID, "null argument");
	*s = (void *)dev->stream;
	return LRB_OK;
}

int lrb_device_set_option(lrb_device *dev, const char *key, const char *value) {
	if (!dev || !key || !value)
		return Fail(LRB_ERR_INVALID, "null argument");
	const std::string k(key), v(value);
	if (k == "kernel") {
		if (v == "persistent") dev->persistent = 1;
		else if (v == "simple" || v == "static") dev->persistent = 0;
		else return Fail(LRB_ERR_INVALID, "kernel must be persistent|simple");
		return LRB_OK;
	}
	const int iv = atoi(value);
	if (k == "blocks_per_sm") {
		if (iv < 0 || iv > 32) return Fail(LRB_ERR_INVALID, "blocks_per_sm out of range");
		dev->blocksPerSM = iv;
	} else if (k == "smem_depth") {
		if (iv < 1 || iv > 128) return Fail(LRB_ERR_INVALID, "smem_depth out of range");
		dev->smemDepth = iv;
	} else if (k == "refill_below") {
		if (iv < 1 || iv > 32) return Fail(LRB_ERR_INVALID, "refill_below out of range");
		dev->refillBelow = iv;
	} else if (k == "tri_bias") {
		if (iv < 1 || iv > 64) return Fail(LRB_ERR_INVALID, "tri_bias out of range");
		dev->triBias = iv;
	} else if (k == "inst_bias") {
		if (iv < 0 || iv > 64) return Fail(LRB_ERR_INVALID, "inst_bias out of range");
		dev->instBias = iv;
	} else if (k == "gather_stores") {
		dev->gatherStores = iv ? 1 : 0;
	} else if (k == "gather_defer") {
		dev->gatherDefer = iv ? 1 : 0;
	} else if (k == "gather_chunk_shift") {
		if (iv < 12 || iv > 28) return Fail(LRB_ERR_INVALID, "gather_chunk_shift must be 12..28");
		dev->gatherChunkShift = iv;
	} else if (k == "wide_stores") {
		if (iv < 0 || iv > 3) return Fail(LRB_ERR_INVALID, "wide_stores must be 0..3");
		dev->wideStores = iv;
	} else if (k == "compact") {
		if (iv < 0 || iv > 2) return Fail(LRB_ERR_INVALID, "compact must be 0 (masked rays are skipped inside the trace kernel), 1 (compacted before it) or 2 (counted; compacted when fewer than 60 % are live)");
		dev->compact = iv;
	} else if (k == "pipeline") {
		if (iv < 0 || iv > 1) return Fail(LRB_ERR_INVALID, "pipeline must be 0 or 1");
		dev->pipeline = iv;
	} else if (k == "l2_persist") {
		if (iv < 0 || iv > 2) return Fail(LRB_ERR_INVALID, "l2_persist must be 0 (never), 1 (always, clipped to the set-aside) or 2 (scenes that fit the set-aside)");
		dev->l2Persist = iv;
	} else if (k == "carveout") {
		if (iv < -1 || iv > 100) return Fail(LRB_ERR_INVALID, "carveout must be -1 (default) or 0..100 percent of shared memory");
		dev->carveout = iv;
	} else if (k == "prefetch") {
		if (iv < 0 || iv > 2) return Fail(LRB_ERR_INVALID, "prefetch must be 0 (never), 1 (always) or 2 (scenes larger than L2)");
		dev->prefetch = iv;
	} else if (k == "prefetch_mode") {
		if (iv < 1 || iv > 7) return Fail(LRB_ERR_INVALID, "prefetch_mode is a bit mask 1..7 (1 pushed children -> L2, 2 nearest child -> L2, 4 nearest child -> L1)");
		dev->prefetchMode = iv;
	} else if (k == "sort_rays") {
		if (iv < 0 || iv > 2) return Fail(LRB_ERR_INVALID, "sort_rays must be 0 (never), 1 (always) or 2 (scenes larger than L2)");
		dev->sortRays = iv;
	} else if (k == "sort_bits") {
		if (iv < 1 || iv > 9) return Fail(LRB_ERR_INVALID, "sort_bits (per axis) must be 1..9");
		dev->sortBitsPerAxis = iv;
	} else if (k == "sort_min_rays") {
		if (iv < 0) return Fail(LRB_ERR_INVALID, "sort_min_rays out of range");
		dev->sortMinRays = iv;
	} else if (k == "host_chunk") {
		if (iv < 1024) return Fail(LRB_ERR_INVALID, "host_chunk too small");
		dev->hostChunk = iv;
	} else if (k == "host_taper") {
		if (iv < 0 || iv > 1) return Fail(LRB_ERR_INVALID, "host_taper must be 0 or 1");
		dev->hostTaper = iv;
	} else if (k == "host_min_chunk") {
		if (iv < 1024) return Fail(LRB_ERR_INVALID, "host_min_chunk too small");
		dev->hostMinChunk = iv;
	} else
		return Fail(LRB_ERR_INVALID, "unknown option: " + k);
	return LRB_OK;
}

// ---- memory + queue ---------------------------------------------------------------------------

int lrb_alloc(lrb_device *dev, size_t bytes, void **devptr) {
	if (!devptr)
		return Fail(LRB_ERR_INVALID, "null out pointer");
	*devptr = nullptr;
	LRB_SETDEV(dev);
	if (bytes == 0)
		return LRB_OK;
	LRB_CUDA(cudaMalloc(devptr, bytes));
	std::lock_guard<std::mutex> g(dev->mtx);
	dev->allocs[*devptr] = bytes;
	dev->counters.device_bytes_in_use += bytes;
	return LRB_OK;
}

int lrb_free(lrb_device *dev, void *devptr) {
	LRB_SETDEV(dev);
	if (!devptr)
		return LRB_OK;
	{
		std::lock_guard<std::mutex> g(dev->mtx);
		std::unordered_map<void *, size_t>::iterator it = dev->allocs.find(devptr);
		if (it == dev->allocs.end())
			return Fail(LRB_ERR_INVALID, "pointer was not allocated by lrb_alloc on this device");
		dev->counters.device_bytes_in_use -= std::min<uint64_t>(dev->counters.device_bytes_in_use, it->second);
		dev->allocs.erase(it);
	}
	{
		const int rc = JoinPending(dev);    // (cudaFree synchronises the device; the bookkeeping must not outlive the buffer)
		if (rc != LRB_OK) return rc;
	}
	LRB_CUDA(cudaFree(devptr));
	return LRB_OK;
}

int lrb_h2d(lrb_device *dev, void *dst, const void *src, size_t bytes, int blocking) {
	LRB_SETDEV(dev);
	if (bytes == 0)
		return LRB_OK;
	if (!dst || !src)
		return Fail(LRB_ERR_INVALID, "null pointer in h2d");
	dev->counters.h2d_bytes += bytes;
	if (dev->pipeline && !blocking && bytes >= kPipeMinBytes)
		return PipelinedUpload(dev, dst, src, bytes);
	int rc = JoinPending(dev);
	if (rc != LRB_OK) return rc;
	LRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, dev->stream));
	if (blocking)
		LRB_CUDA(cudaStreamSynchronize(dev->stream));
	return LRB_OK;
}

int lrb_d2h(lrb_device *dev, void *dst, const void *src, size_t bytes, int blocking) {
	LRB_SETDEV(dev);
	if (bytes == 0)
		return LRB_OK;
	if (!dst || !src)
		return Fail(LRB_ERR_INVALID, "null pointer in d2h");
	dev->counters.d2h_bytes += bytes;
	lrb_device::Pending &t = dev->pendTraced;
	if (t.active && (const char *)src == t.base && bytes == t.bytes) {
		// the RayHit buffer of the chunked trace: every chunk leaves as soon as it has been traced
		size_t off = 0;
		for (size_t c = 0; c < t.ends.size(); ++c) {
			const size_t cnt = (size_t)t.ends[c] - off;
			LRB_CUDA(cudaStreamWaitEvent(dev->copyOutStream, t.ev[c], 0));
			LRB_CUDA(cudaMemcpyAsync((char *)dst + off, (const char *)src + off, cnt, cudaMemcpyDeviceToHost, dev->copyOutStream));
			off = (size_t)t.ends[c];
		}
		t.active = false;
		dev->copyOutBusy = true;
		if (blocking) {
			const int rc = JoinPending(dev);
			if (rc != LRB_OK) return rc;
			LRB_CUDA(cudaStreamSynchronize(dev->stream));
		}
		return LRB_OK;
	}
	const int rc = JoinPending(dev);
	if (rc != LRB_OK) return rc;
	LRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream));
	if (blocking)
		LRB_CUDA(cudaStreamSynchronize(dev->stream));
	return LRB_OK;
}

int lrb_flush(lrb_device *dev) {
	LRB_SETDEV(dev);
	// CUDA submits eagerly; cuStreamQuery is what the reference uses to kick the queue
	const cudaError_t e = cudaStreamQuery(dev->stream);
	if (e != cudaSuccess && e != cudaErrorNotReady)
		LRB_CUDA(e);
	return LRB_OK;
}

int lrb_sync(lrb_device *dev) {
	LRB_SETDEV(dev);
	{
		const int rc = JoinPending(dev);
		if (rc != LRB_OK) return rc;
	}
	if (dev->gatherPending[0] || dev->gatherPending[1]) {      // deferred gather pushes (gather_defer)
		LRB_CUDA(cudaStreamSynchronize(dev->copyOutStream));
		dev->gatherPending[0] = dev->gatherPending[1] = false;
	}
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	return LRB_OK;
}

int lrb_get_counters(lrb_device *dev, lrb_counters *out) {
	if (!dev || !out)
		return Fail(LRB_ERR_INVALID, "null argument");
	*out = dev->counters;
	return LRB_OK;
}

int lrb_reset_counters(lrb_device *dev) {
	if (!dev)
		return Fail(LRB_ERR_INVALID, "null device");
	const uint64_t inUse = dev->counters.device_bytes_in_use;
	memset(&dev->counters, 0, sizeof(dev->counters));
	dev->counters.device_bytes_in_use = inUse;
	return LRB_OK;
}

// ---- bandwidth probe --------------------------------------------------------------------------

}   // extern "C"

// Eight independent 256-bit read-only loads (256 B) in flight per thread: 2 048 threads / SM x 256 B =
// 512 KB in flight per SM, enough to cover the L2 / HBM latency-bandwidth product (round 1's probe kept
// four 16-B loads in flight and under-measured L2).
__global__ void __launch_bounds__(256) ReadProbeKernel(const uint4 *__restrict__ src, size_t n, unsigned *sink) {
	unsigned acc = 0;
	const size_t n32 = n / 2;       // 32-byte units
	const char *base = reinterpret_cast<const char *>(src);
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + 7 * stride < n32; i += 8 * stride) {
		F8 v[8];
#pragma unroll
		for (int k = 0; k < 8; ++k)
			v[k] = Ld256(base + ((i + k * stride) << 5));
#pragma unroll
		for (int k = 0; k < 8; ++k)
			acc ^= __float_as_uint(v[k].v[0]) ^ __float_as_uint(v[k].v[1]) ^ __float_as_uint(v[k].v[2]) ^ __float_as_uint(v[k].v[3]) ^
					__float_as_uint(v[k].v[4]) ^ __float_as_uint(v[k].v[5]) ^ __float_as_uint(v[k].v[6]) ^ __float_as_uint(v[k].v[7]);
	}
	for (; i < n32; i += stride) {
		const F8 v = Ld256(base + (i << 5));
		acc ^= __float_as_uint(v.v[0]) ^ __float_as_uint(v.v[1]) ^ __float_as_uint(v.v[2]) ^ __float_as_uint(v.v[3]) ^
				__float_as_uint(v.v[4]) ^ __float_as_uint(v.v[5]) ^ __float_as_uint(v.v[6]) ^ __float_as_uint(v.v[7]);
	}
	if (acc == 0x9e3779b9u)
		*sink = acc;    // practically never: keeps the loads alive
}

extern "C" {

int lrb_measure_read_bandwidth(lrb_device *dev, size_t bytes, int iters, double *gbps) {
	if (!gbps || bytes < 4096 || iters < 1)
		return Fail(LRB_ERR_INVALID, "bad argument");
	LRB_SETDEV(dev);
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	void *buf = nullptr;
	unsigned *sink = nullptr;
	LRB_CUDA(cudaMalloc(&buf, bytes));
	LRB_CUDA(cudaMalloc((void **)&sink, 4));
	LRB_CUDA(cudaMemsetAsync(buf, 1, bytes, dev->stream));
	cudaEvent_t e0, e1;
	LRB_CUDA(cudaEventCreate(&e0));
	LRB_CUDA(cudaEventCreate(&e1));
	const size_t n = (bytes / 32) * 2;      // 16-byte units, whole 32-byte records
	const int grid = dev->prop.multiProcessorCount * 8;
	for (int w = 0; w < 3; ++w)
		ReadProbeKernel<<<grid, 256, 0, dev->stream>>>((const uint4 *)buf, n, sink);
	LRB_CUDA(cudaEventRecord(e0, dev->stream));
	for (int it = 0; it < iters; ++it)
		ReadProbeKernel<<<grid, 256, 0, dev->stream>>>((const uint4 *)buf, n, sink);
	LRB_CUDA(cudaEventRecord(e1, dev->stream));
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	float ms = 0.f;
	LRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(buf); cudaFree(sink);
	dev->counters.kernel_launches += iters + 3;
	*gbps = (double)n * 16.0 * iters / (ms * 1e-3) / 1e9;
	return LRB_OK;
}

// ---- scenes -----------------------------------------------------------------------------------

}   // extern "C"

// L2 persistence (the reference only asks for CU_FUNC_CACHE_PREFER_L1, cudadevice.cpp:162).
// A scene that is L2-resident competes for L2 with what streams through it: the 48-B rays and 20-B RayHits of its own
// batch and, on the gathering GPU of a multi-GPU run, the other ranks' RayHit slices arriving over NVLink.  An access-
// policy window on the queue marks the scene's slab "persisting" for the kernels launched there; the set-aside is sized
// once per device.  Clipped to the device's limits; a no-op for scenes that do not fit.
static int ClearL2Window(lrb_device *dev) {
	if (!dev->l2WindowBase)
		return LRB_OK;
	cudaStreamAttrValue attr;
	memset(&attr, 0, sizeof(attr));
	attr.accessPolicyWindow.num_bytes = 0;
	LRB_CUDA(cudaStreamSetAttribute(dev->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
	dev->l2WindowBase = nullptr;
	dev->l2WindowBytes = 0;
	return LRB_OK;
}

static int SetL2Window(lrb_device *dev, const void *base, size_t bytes) {
	const size_t maxPersist = (size_t)dev->prop.persistingL2CacheMaxSize, maxWindow = (size_t)dev->prop.accessPolicyMaxWindowSize;
	const bool fits = base && bytes && maxPersist && bytes <= maxPersist && bytes <= maxWindow;
	const bool want = dev->l2Persist == 1 ? (base && bytes && maxPersist && maxWindow) : (dev->l2Persist == 2 && fits);
	if (!want)
		return ClearL2Window(dev);
	if (dev->l2WindowBase == base && dev->l2WindowBytes == bytes)
		return LRB_OK;
	size_t limit = 0;
	LRB_CUDA(cudaDeviceGetLimit(&limit, cudaLimitPersistingL2CacheSize));
	const size_t need = std::min(bytes, maxPersist);
	if (limit < need)
		LRB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, need));
	cudaStreamAttrValue attr;
	memset(&attr, 0, sizeof(attr));
	attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
	attr.accessPolicyWindow.num_bytes = std::min(bytes, maxWindow);
	attr.accessPolicyWindow.hitRatio = fits ? 1.f : (float)((double)maxPersist / (double)bytes);
	attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	LRB_CUDA(cudaStreamSetAttribute(dev->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
	dev->l2WindowBase = base;
	dev->l2WindowBytes = bytes;
	return LRB_OK;
}

template <class V, class T = typename V::value_type> static int UploadArray(lrb_device *dev, const V &src, T **dst, size_t *cap, uint64_t *bytes) {
	const size_t n = src.size();
	if (cap && *dst && *cap >= n) {
		// re-use the allocation (Update path)
	} else {
		if (*dst) {
			LRB_CUDA(cudaFree(*dst));
			*dst = nullptr;
			if (bytes && cap) *bytes -= *cap * sizeof(T);      // re-grown (Update): the old allocation is gone
		}
		if (n) {
			LRB_CUDA(cudaMalloc((void **)dst, n * sizeof(T)));
			if (bytes) *bytes += n * sizeof(T);
		}
		if (cap) *cap = n;
	}
	if (n) {
		LRB_CUDA(cudaMemcpyAsync(*dst, src.data(), n * sizeof(T), cudaMemcpyHostToDevice, dev->stream));
		dev->counters.h2d_bytes += n * sizeof(T);
	}
	return LRB_OK;
}

static void FillView(lrb_scene *s) {
	SceneView &v = s->view;
	v.nodes = s->dNodes;
	v.tris = s->dTris;
	v.ids = s->dIds;
	v.insts = s->dInsts;
	v.minv = s->dMinv;
	v.motionFirst = s->dMotionFirst;
	v.motionLast = s->dMotionLast;
	v.interps = s->dInterps;
	FillRootOfView(s->host, &v);
	s->info.n_ref_nodes = s->host.nRefNodes;
	s->info.n_wide_nodes = (uint32_t)s->host.wide.size();
	s->info.n_triangles = (uint32_t)s->host.tris.size();
	s->info.n_instances = (uint32_t)s->host.insts.size();
	s->info.stack_need = s->host.stackNeed;
	s->info.two_level = v.twoLevel;
}

static int UploadScene(lrb_scene *s) {
	lrb_device *dev = s->dev;
	uint64_t bytes = 0;
	int rc;
	if ((rc = JoinPending(dev)) != LRB_OK) return rc;
	if (!s->host.twoLevel && !s->dNodes && !s->host.wide.empty()) {
		// one-level scenes never change: nodes, triangle records and ids share ONE allocation, so that a single
		// L2 access-policy window (SetL2Window) can keep all of it resident
		const size_t nb = (s->host.wide.size() * sizeof(WideNode) + 255) & ~(size_t)255;
		const size_t tb = (s->host.tris.size() * sizeof(TriRecord) + 255) & ~(size_t)255;
		const size_t ib = (s->host.ids.size() * sizeof(TriIds) + 255) & ~(size_t)255;
		LRB_CUDA(cudaMalloc(&s->dSlab, nb + tb + ib));
		s->slabBytes = nb + tb + ib;
		bytes += s->slabBytes;
		s->dNodes = reinterpret_cast<WideNode *>(s->dSlab);
		s->dTris = reinterpret_cast<TriRecord *>((char *)s->dSlab + nb);
		s->dIds = reinterpret_cast<TriIds *>((char *)s->dSlab + nb + tb);
		s->capNodes = s->host.wide.size();
		LRB_CUDA(cudaMemcpyAsync(s->dNodes, s->host.wide.data(), s->host.wide.size() * sizeof(WideNode), cudaMemcpyHostToDevice, dev->stream));
		if (!s->host.tris.empty()) {
			LRB_CUDA(cudaMemcpyAsync(s->dTris, s->host.tris.data(), s->host.tris.size() * sizeof(TriRecord), cudaMemcpyHostToDevice, dev->stream));
			LRB_CUDA(cudaMemcpyAsync(s->dIds, s->host.ids.data(), s->host.ids.size() * sizeof(TriIds), cudaMemcpyHostToDevice, dev->stream));
		}
		dev->counters.h2d_bytes += s->slabBytes;
	} else {
		if ((rc = UploadArray(dev, s->host.wide, &s->dNodes, &s->capNodes, &bytes)) != LRB_OK) return rc;
		if ((rc = UploadArray(dev, s->host.tris, &s->dTris, (size_t *)nullptr, &bytes)) != LRB_OK) return rc;
		if ((rc = UploadArray(dev, s->host.ids, &s->dIds, (size_t *)nullptr, &bytes)) != LRB_OK) return rc;
	}
	if ((rc = UploadArray(dev, s->host.insts, &s->dInsts, &s->capInsts, &bytes)) != LRB_OK) return rc;
	if ((rc = UploadArray(dev, s->host.minv, &s->dMinv, (size_t *)nullptr, &bytes)) != LRB_OK) return rc;
	if ((rc = UploadArray(dev, s->host.motionFirst, &s->dMotionFirst, (size_t *)nullptr, &bytes)) != LRB_OK) return rc;
	if ((rc = UploadArray(dev, s->host.motionLast, &s->dMotionLast, (size_t *)nullptr, &bytes)) != LRB_OK) return rc;
	if ((rc = UploadArray(dev, s->host.interps, &s->dInterps, (size_t *)nullptr, &bytes)) != LRB_OK) return rc;
	if (!s->dCounter) {
		LRB_CUDA(cudaMalloc((void **)&s->dCounter, 256));
		LRB_CUDA(cudaMalloc((void **)&s->dStats, sizeof(TraceStats)));
		bytes += 256 + sizeof(TraceStats);
	}
	// the host vectors are pageable: wait for the copies before they can be released/modified
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	s->info.device_bytes += bytes;
	{
		std::lock_guard<std::mutex> g(dev->mtx);
		dev->counters.device_bytes_in_use += bytes;
	}
	FillView(s);
	return LRB_OK;
}

static lrb_scene *NewScene(lrb_device *dev) {
	lrb_scene *s = new lrb_scene();
	s->dev = dev;
	s->dNodes = nullptr; s->dTris = nullptr; s->dIds = nullptr; s->dInsts = nullptr; s->dMinv = nullptr;
	s->dSlab = nullptr; s->slabBytes = 0;
	s->dMotionFirst = s->dMotionLast = nullptr; s->dInterps = nullptr;
	s->capNodes = s->capInsts = 0;
	s->dCounter = nullptr; s->dSpillNode = nullptr; s->dSpillT = nullptr; s->spillEntries = 0;
	s->dStats = nullptr;
	s->dWatermark = s->dChunkFlag = nullptr; s->chunkCap = 0; s->watermarkCap = 0; s->epoch = 0;
	memset(&s->view, 0, sizeof(s->view));
	memset(&s->info, 0, sizeof(s->info));
	return s;
}

extern "C" {

int lrb_scene_free(lrb_scene *s) {
	if (!s)
		return LRB_OK;
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	cudaStreamSynchronize(dev->stream);
	if (s->dSlab) {
		if (dev->l2WindowBase == s->dSlab)
			ClearL2Window(dev);
		cudaFree(s->dSlab);
	} else {
		cudaFree(s->dNodes); cudaFree(s->dTris); cudaFree(s->dIds);
	}
	cudaFree(s->dInsts); cudaFree(s->dMinv);
	cudaFree(s->dMotionFirst); cudaFree(s->dMotionLast); cudaFree(s->dInterps);
	cudaFree(s->dCounter); cudaFree(s->dSpillNode); cudaFree(s->dSpillT); cudaFree(s->dStats); cudaFree(s->dWatermark); cudaFree(s->dChunkFlag);
	{
		std::lock_guard<std::mutex> g(dev->mtx);
		dev->counters.device_bytes_in_use -= std::min<uint64_t>(dev->counters.device_bytes_in_use, s->info.device_bytes);
	}
	delete s;
	return LRB_OK;
}

int lrb_bvh_upload(lrb_device *dev, const lrb_bvh_node *nodes, uint32_t nNodes, const float *xyz, uint64_t nVerts,
		const uint32_t *meshVertexOffsets, uint32_t nMeshes, lrb_scene **out) {
	if (!out)
		return Fail(LRB_ERR_INVALID, "null out pointer");
	*out = nullptr;
	LRB_SETDEV(dev);
	lrb_scene *s = NewScene(dev);
	try {
		BuildWideBVH(nodes, nNodes, xyz, nVerts, meshVertexOffsets, nMeshes, &s->host);
	} catch (const std::bad_alloc &) {
		delete s;
		return Fail(LRB_ERR_OOM, "host out of memory during BVH re-layout");
	} catch (const std::exception &e) {
		delete s;
		return Fail(LRB_ERR_INVALID, e.what());
	}
	const int rc = UploadScene(s);
	if (rc != LRB_OK) {
		lrb_scene_free(s);
		return rc;
	}
	// single-level scenes never change: drop the host copy (the view / info keep the bookkeeping)
	RawVector<WideNode>().swap(s->host.wide);
	RawVector<TriRecord>().swap(s->host.tris);
	RawVector<TriIds>().swap(s->host.ids);
	*out = s;
	return LRB_OK;
}

int lrb_mbvh_upload(lrb_device *dev, const lrb_mbvh_desc *desc, lrb_scene **out) {
	if (!out || !desc)
		return Fail(LRB_ERR_INVALID, "null argument");
	*out = nullptr;
	LRB_SETDEV(dev);
	lrb_scene *s = NewScene(dev);
	try {
		BuildWideMBVH(*desc, &s->host);
	} catch (const std::bad_alloc &) {
		delete s;
		return Fail(LRB_ERR_OOM, "host out of memory during MBVH re-layout");
	} catch (const std::exception &e) {
		delete s;
		return Fail(LRB_ERR_INVALID, e.what());
	}
	const int rc = UploadScene(s);
	if (rc != LRB_OK) {
		lrb_scene_free(s);
		return rc;
	}
	// triangles are immutable under Update; the wide nodes / instances stay on the host
	RawVector<TriRecord>().swap(s->host.tris);
	RawVector<TriIds>().swap(s->host.ids);
	s->host.tris.resize(0);
	*out = s;
	return LRB_OK;
}

int lrb_mbvh_update(lrb_scene *s, const lrb_bvh_node *rootNodes, uint32_t nRootNodes, const float *minv, uint32_t nTransforms) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	const uint32_t nTris = s->info.n_triangles;
	try {
		UpdateWideMBVHRoot(rootNodes, nRootNodes, minv, nTransforms, &s->host);
	} catch (const std::exception &e) {
		return Fail(LRB_ERR_INVALID, e.what());
	}
	// in-order with earlier traces on the stream; only the root tail, instances and matrices move
	uint64_t bytes = 0;
	int rc;
	if ((rc = UploadArray(dev, s->host.wide, &s->dNodes, &s->capNodes, &bytes)) != LRB_OK) return rc;
	if ((rc = UploadArray(dev, s->host.insts, &s->dInsts, &s->capInsts, &bytes)) != LRB_OK) return rc;
	if (s->host.minv.size()) {
		LRB_CUDA(cudaMemcpyAsync(s->dMinv, s->host.minv.data(), s->host.minv.size() * sizeof(float), cudaMemcpyHostToDevice, dev->stream));
		dev->counters.h2d_bytes += s->host.minv.size() * sizeof(float);
	}
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	s->info.device_bytes += bytes;
	{
		std::lock_guard<std::mutex> g(dev->mtx);
		dev->counters.device_bytes_in_use += bytes;
	}
	FillView(s);
	s->info.n_triangles = nTris;
	return LRB_OK;
}

int lrb_scene_get_info(lrb_scene *s, lrb_scene_info *out) {
	if (!s || !out)
		return Fail(LRB_ERR_INVALID, "null argument");
	*out = s->info;
	return LRB_OK;
}

// ---- trace ------------------------------------------------------------------------------------

}   // extern "C"

template <class K> static int Occupancy(K kernel, int block, int smemBytes, int *blocksPerSM, int carveout = -1) {
	if (carveout >= 0)
		LRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
	if (smemBytes > 48 * 1024)
		LRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
	LRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocksPerSM, kernel, block, smemBytes));
	return LRB_OK;
}

static int EnsureSpill(lrb_scene *s, uint32_t residentDepth, int totalThreads) {
	const uint32_t need = s->info.stack_need;
	if (need <= residentDepth)
		return LRB_OK;
	const size_t entries = (size_t)(need - residentDepth) * (size_t)totalThreads;
	if (entries <= s->spillEntries)
		return LRB_OK;
	lrb_device *dev = s->dev;
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	if (s->dSpillNode) { cudaFree(s->dSpillNode); cudaFree(s->dSpillT); s->dSpillNode = nullptr; s->dSpillT = nullptr; }
	LRB_CUDA(cudaMalloc((void **)&s->dSpillNode, entries * sizeof(uint32_t)));
	LRB_CUDA(cudaMalloc((void **)&s->dSpillT, entries * sizeof(float)));
	s->info.device_bytes += (entries - s->spillEntries) * 8;
	{
		std::lock_guard<std::mutex> g(dev->mtx);        // lrb_scene_free takes the scene's whole device_bytes off again
		dev->counters.device_bytes_in_use += (entries - s->spillEntries) * 8;
	}
	s->spillEntries = entries;
	return LRB_OK;
}

typedef void (*PersistentKernel)(const TraceArgs);

// Ray-ordering pre-pass: keys from origin cell + direction octant, LSD radix sort of (key, index).
// Leaves the processing order in *perm (device memory owned by the device object).
static int SortRays(lrb_scene *s, const void *rays, uint32_t n, cudaStream_t stream, const uint32_t **perm) {
	lrb_device *dev = s->dev;
	const int bits = dev->sortBitsPerAxis;
	const int keyBits = 3 * bits + 3;
	cub::DoubleBuffer<uint32_t> keys(dev->sortKeys[0], dev->sortKeys[1]), vals(dev->sortVals[0], dev->sortVals[1]);
	if (dev->sortCap < n) {
		LRB_CUDA(cudaStreamSynchronize(stream));
		for (int i = 0; i < 2; ++i) {
			cudaFree(dev->sortKeys[i]); cudaFree(dev->sortVals[i]);
			dev->sortKeys[i] = dev->sortVals[i] = nullptr;
		}
		dev->sortCap = 0;
		for (int i = 0; i < 2; ++i) {
			LRB_CUDA(cudaMalloc((void **)&dev->sortKeys[i], (size_t)n * sizeof(uint32_t)));
			LRB_CUDA(cudaMalloc((void **)&dev->sortVals[i], (size_t)n * sizeof(uint32_t)));
		}
		dev->sortCap = n;
		keys = cub::DoubleBuffer<uint32_t>(dev->sortKeys[0], dev->sortKeys[1]);
		vals = cub::DoubleBuffer<uint32_t>(dev->sortVals[0], dev->sortVals[1]);
	}
	size_t tempBytes = 0;
	LRB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keys, vals, (int)n, 0, keyBits, stream));
	if (dev->sortTempBytes < tempBytes) {
		LRB_CUDA(cudaStreamSynchronize(stream));
		cudaFree(dev->sortTemp);
		dev->sortTemp = nullptr; dev->sortTempBytes = 0;
		LRB_CUDA(cudaMalloc(&dev->sortTemp, tempBytes));
		dev->sortTempBytes = tempBytes;
	}
	const float *rb = s->view.rootBox;
	const float cells = (float)(1u << bits);
	float sc[3];
	for (int k = 0; k < 3; ++k) {
		const float ext = rb[3 + k] - rb[k];
		sc[k] = (ext > 0.f && ext < 3e38f) ? cells / ext : 0.f;
	}
	RayKeyKernel<<<(n + 255) / 256, 256, 0, stream>>>((const lrb_ray *)rays, n, rb[0], rb[1], rb[2], sc[0], sc[1], sc[2],
			(uint32_t)bits, keys.Current(), vals.Current());
	LRB_CUDA(cudaGetLastError());
	LRB_CUDA(cub::DeviceRadixSort::SortPairs(dev->sortTemp, tempBytes, keys, vals, (int)n, 0, keyBits, stream));
	*perm = vals.Current();
	dev->counters.kernel_launches += 1 + (keyBits + 7) / 8 + 1;
	return LRB_OK;
}

static PersistentKernel PickPersistent(bool two, bool spill, bool signal, bool prefetch, bool anyhit = false) {
	if (anyhit) {       // shadow rays: no gather signalling, no prefetch twin
		if (two) return spill ? TracePersistent<true, true, false, false, true> : TracePersistent<true, false, false, false, true>;
		return spill ? TracePersistent<false, true, false, false, true> : TracePersistent<false, false, false, false, true>;
	}
	// the prefetching variant exists for one-level scenes with a spilling stack (large scenes are both)
	if (prefetch && !two && spill)
		return signal ? TracePersistent<false, true, true, true> : TracePersistent<false, true, false, true>;
	if (two) {
		if (spill) return signal ? TracePersistent<true, true, true> : TracePersistent<true, true, false>;
		return signal ? TracePersistent<true, false, true> : TracePersistent<true, false, false>;
	}
	if (spill) return signal ? TracePersistent<false, true, true> : TracePersistent<false, true, false>;
	return signal ? TracePersistent<false, false, true> : TracePersistent<false, false, false>;
}

// Driver entry points for stream-ordered memory operations, resolved at run time (the library does
// not link against libcuda, so that it also loads on a machine without a driver).
typedef CUresult (*PFN_StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*PFN_MemsetD32Async)(CUdeviceptr, unsigned int, size_t, CUstream);
static PFN_StreamWaitValue32 g_streamWaitValue32 = nullptr;
static PFN_MemsetD32Async g_memsetD32Async = nullptr;

static int ResolveDriverEntryPoints() {
	if (g_streamWaitValue32 && g_memsetD32Async)
		return LRB_OK;
	void *f0 = nullptr, *f1 = nullptr;
	cudaDriverEntryPointQueryResult q0, q1;
	LRB_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &f0, cudaEnableDefault, &q0));
	LRB_CUDA(cudaGetDriverEntryPoint("cuMemsetD32Async", &f1, cudaEnableDefault, &q1));
	if (!f0 || !f1 || q0 != cudaDriverEntryPointSuccess || q1 != cudaDriverEntryPointSuccess)
		return Fail(LRB_ERR_CUDA, "driver entry points for stream memory operations are not available");
	g_streamWaitValue32 = (PFN_StreamWaitValue32)f0;
	g_memsetD32Async = (PFN_MemsetD32Async)f1;
	return LRB_OK;
}

// Dense list of the live (non-masked) rays of a batch, in increasing index order (batch_kernels.cuh).  The
// list and its length stay on the device: *idx / *count are device pointers owned by the device object.
static int CompactRays(lrb_device *dev, const void *rays, uint32_t n, cudaStream_t stream, const uint32_t **idx, const uint32_t **count,
		bool autoMode = false) {
	const uint32_t nBlocks = (n + kCompactBlock - 1) / kCompactBlock;
	if (dev->compactCap < n) {
		LRB_CUDA(cudaStreamSynchronize(stream));
		cudaFree(dev->compactIdx); cudaFree(dev->compactBlocks); cudaFree(dev->compactTotal);
		dev->compactIdx = dev->compactBlocks = dev->compactTotal = nullptr;
		dev->compactCap = 0;
		LRB_CUDA(cudaMalloc((void **)&dev->compactIdx, (size_t)n * sizeof(uint32_t)));
		LRB_CUDA(cudaMalloc((void **)&dev->compactBlocks, (size_t)nBlocks * sizeof(uint32_t)));
		LRB_CUDA(cudaMalloc((void **)&dev->compactTotal, 64));
		dev->compactCap = n;
	}
	CompactCountKernel<<<nBlocks, 256, 0, stream>>>((const lrb_ray *)rays, n, dev->compactBlocks);
	CompactScanKernel<<<1, 1024, 0, stream>>>(dev->compactBlocks, nBlocks, dev->compactTotal);
	CompactScatterKernel<<<nBlocks, 256, 0, stream>>>((const lrb_ray *)rays, n, dev->compactBlocks, dev->compactIdx,
			autoMode ? dev->compactTotal : nullptr);
	LRB_CUDA(cudaGetLastError());
	dev->counters.kernel_launches += 3;
	*idx = dev->compactIdx;
	*count = dev->compactTotal;
	return LRB_OK;
}

struct TraceMode {
	bool anyhit;                    // shadow rays: the first accepted hit ends the ray
	const uint32_t *liveIdx;        // trace only the rays listed here (device pointer) ...
	const uint32_t *liveCountDev;   // ... as many as this device word says (or n when NULL)
	TraceMode() : anyhit(false), liveIdx(nullptr), liveCountDev(nullptr) { }
};

static int LaunchTrace(lrb_scene *s, const void *rays, void *hits, uint32_t n, bool stats, cudaStream_t stream,
		lrb_rayhit *hitsPeer = nullptr, bool signal = false, TraceMode mode = TraceMode()) {
	lrb_device *dev = s->dev;
	if (n == 0)
		return LRB_OK;
	if (!rays || (!hits && !hitsPeer && !stats))
		return Fail(LRB_ERR_INVALID, "null ray/hit buffer");
	if ((reinterpret_cast<uintptr_t>(rays) & 15u) != 0)
		return Fail(LRB_ERR_INVALID, "ray buffer must be 16-byte aligned");
	{
		const int rcJoin = JoinPending(dev);    // chunked uploads / reads of the pipelined plugin sequence (no-op when there are none)
		if (rcJoin != LRB_OK) return rcJoin;
	}

	TraceArgs a;
	memset(&a, 0, sizeof(a));
	a.sc = s->view;
	a.rays = (const lrb_ray *)rays;
	a.hits = (lrb_rayhit *)hits;
	a.hitsPeer = hitsPeer;
	// vector stores: off for the local buffer (the four-way switch costs more than it saves there),
	// optional for the peer buffer
	a.hitFlags = ((dev->wideStores & 1) && (reinterpret_cast<uintptr_t>(hits) & 15u) == 0 ? 1u : 0u) |
			((dev->wideStores & 2) && (reinterpret_cast<uintptr_t>(hitsPeer) & 15u) == 0 ? 2u : 0u);
	a.rayCount = n;
	a.counter = s->dCounter;
	a.stats = s->dStats;
	a.refillBelow = (uint32_t)dev->refillBelow;
	a.triBias = (uint32_t)dev->triBias;
	a.instBias = (uint32_t)dev->instBias;
	a.prefetchMode = (uint32_t)dev->prefetchMode;
	const bool two = s->view.twoLevel != 0;
	const int sm = dev->prop.multiProcessorCount;
	int rc;
	if (!mode.liveIdx && dev->compact && !signal && !stats && n >= 4096) {
		if ((rc = CompactRays(dev, rays, n, stream, &mode.liveIdx, &mode.liveCountDev, dev->compact == 2)) != LRB_OK) return rc;
		a.permAuto = dev->compact == 2 ? 1u : 0u;
	}
	a.perm = mode.liveIdx;
	a.rayCountDev = mode.liveCountDev;

	if (dev->persistent && !stats) {
		const int block = kTraceBlock;
		int depth = std::min<int>(dev->smemDepth, (int)std::max<uint32_t>(s->info.stack_need, 4u));
		const int smemBytes = (depth + 1) * block * 8 + (two ? 9 * block * 4 : 0);     // stack columns incl. the bottom entry (+ the world ray, SmemStack::stashRay)
		int bps = 0;
		const bool spill = s->info.stack_need > (uint32_t)depth;
		const size_t sceneBytes = (size_t)s->info.n_wide_nodes * sizeof(WideNode) + (size_t)s->info.n_triangles * sizeof(TriRecord);
		const bool bigScene = sceneBytes > (size_t)dev->prop.l2CacheSize;
		PersistentKernel kernel = PickPersistent(two, spill, signal, dev->prefetch == 1 || (dev->prefetch == 2 && bigScene), mode.anyhit);
		// shared-memory carve-out: enough for the resident blocks the kernel was compiled for (stack columns + 1 KB
		// reserved per block); measured: a larger L1 does not help this kernel (carveout sweep, profiles/r02_kitchen_sweeps.json)
		int carve = dev->carveout;
		if (carve < 0) {
			const long long want = (long long)(two ? LRB_MINBLOCKS_2L : LRB_MINBLOCKS_1L) * (smemBytes + 1024);
			const long long cap = (long long)dev->prop.sharedMemPerMultiprocessor;
			carve = (int)std::min<long long>(100, (want * 100 + cap - 1) / std::max<long long>(cap, 1));
		}
		if ((rc = Occupancy(kernel, block, smemBytes, &bps, carve)) != LRB_OK) return rc;
		if (bps < 1)
			return Fail(LRB_ERR_INTERNAL, "traversal kernel does not fit on an SM with the requested smem_depth");
		if (dev->blocksPerSM > 0) bps = std::min(bps, dev->blocksPerSM);
		// never launch more threads than rays
		long long grid = (long long)sm * bps;
		const long long maxUseful = ((long long)n + block - 1) / block;
		if (grid > maxUseful) grid = maxUseful;
		a.smemDepth = (uint32_t)depth;
		if ((rc = EnsureSpill(s, (uint32_t)depth, (int)grid * block)) != LRB_OK) return rc;
		a.spillNode = s->dSpillNode;
		a.spillT = s->dSpillT;
		LRB_CUDA(cudaMemsetAsync(s->dCounter, 0, 2 * sizeof(uint32_t), stream));
		if (stream == dev->stream && (rc = SetL2Window(dev, s->dSlab, s->slabBytes)) != LRB_OK) return rc;
		// optional coherence pre-pass
		// (measured on a 2 GB triangle soup: 614 -> 720 Mrays/s; on the L2-resident kitchen the sort costs what it gains)
		const bool wantSort = dev->sortRays == 1 || (dev->sortRays == 2 && bigScene);
		if (wantSort && !signal && !a.perm && s->view.rootHasBox && n >= (uint32_t)dev->sortMinRays) {
			if ((rc = SortRays(s, rays, n, stream, &a.perm)) != LRB_OK) return rc;
		}
		if (signal) {
			const size_t nChunks = ((size_t)n >> dev->gatherChunkShift) + 1;
			const size_t nWarps = (size_t)grid * block / 32;
			if (s->chunkCap < nChunks || s->watermarkCap < nWarps) {
				LRB_CUDA(cudaStreamSynchronize(stream));
				cudaFree(s->dWatermark); cudaFree(s->dChunkFlag);
				s->dWatermark = s->dChunkFlag = nullptr; s->chunkCap = 0; s->watermarkCap = 0;
				const size_t capC = std::max(nChunks, s->chunkCap), capW = std::max<size_t>(nWarps, 16384);
				LRB_CUDA(cudaMalloc((void **)&s->dWatermark, capW * sizeof(uint32_t)));
				LRB_CUDA(cudaMalloc((void **)&s->dChunkFlag, capC * sizeof(uint32_t)));
				LRB_CUDA(cudaMemset(s->dChunkFlag, 0, capC * sizeof(uint32_t)));
				s->chunkCap = capC;
				s->watermarkCap = capW;
				s->epoch = 0;
			}
			if (grid < 2)
				return Fail(LRB_ERR_INVALID, "signalled gather needs at least two blocks");
			LRB_CUDA(cudaMemsetAsync(s->dWatermark, 0, nWarps * sizeof(uint32_t), stream));
			a.watermark = s->dWatermark;
			a.chunkFlag = s->dChunkFlag;
			a.chunkShift = (uint32_t)dev->gatherChunkShift;
			a.epoch = ++s->epoch;
		}
		kernel<<<(unsigned)grid, block, smemBytes, stream>>>(a);
	} else {
		const int block = kTraceBlock;
		int bps = 0;
		if (stats) {
			if (two) rc = Occupancy(TraceStatic<true, true>, block, 0, &bps);
			else rc = Occupancy(TraceStatic<false, true>, block, 0, &bps);
		} else {
			if (mode.anyhit) rc = two ? Occupancy(TraceStatic<true, false, true>, block, 0, &bps) : Occupancy(TraceStatic<false, false, true>, block, 0, &bps);
			else if (two) rc = Occupancy(TraceStatic<true, false>, block, 0, &bps);
			else rc = Occupancy(TraceStatic<false, false>, block, 0, &bps);
		}
		if (rc != LRB_OK) return rc;
		if (bps < 1) bps = 1;
		if (dev->blocksPerSM > 0) bps = std::min(bps, dev->blocksPerSM);
		long long grid = (long long)sm * bps;
		const long long maxUseful = ((long long)n + block - 1) / block;
		if (grid > maxUseful) grid = maxUseful;
		if ((rc = EnsureSpill(s, 32u, (int)grid * block)) != LRB_OK) return rc;
		a.spillNode = s->dSpillNode;
		a.spillT = s->dSpillT;
		if (stats) {
			LRB_CUDA(cudaMemsetAsync(s->dStats, 0, sizeof(TraceStats), stream));
			if (two) TraceStatic<true, true><<<(unsigned)grid, block, 0, stream>>>(a);
			else TraceStatic<false, true><<<(unsigned)grid, block, 0, stream>>>(a);
		} else if (mode.anyhit) {
			if (two) TraceStatic<true, false, true><<<(unsigned)grid, block, 0, stream>>>(a);
			else TraceStatic<false, false, true><<<(unsigned)grid, block, 0, stream>>>(a);
		} else {
			if (two) TraceStatic<true, false><<<(unsigned)grid, block, 0, stream>>>(a);
			else TraceStatic<false, false><<<(unsigned)grid, block, 0, stream>>>(a);
		}
	}
	LRB_CUDA(cudaGetLastError());
	dev->counters.rays_traced += n;
	dev->counters.trace_launches += 1;
	dev->counters.kernel_launches += 1;
	return LRB_OK;
}

extern "C" {

int lrb_trace(lrb_scene *s, const void *rays, void *hits, uint32_t n) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	LRB_SETDEV(s->dev);
	lrb_device *dev = s->dev;
	for (int k = 0; k < 2; ++k) {
		lrb_device::Pending &u = dev->pendUpload[k];
		if (!u.active || (const char *)rays != u.base || (size_t)n * sizeof(lrb_ray) != u.bytes || !hits || n == 0)
			continue;
		// the ray buffer is arriving in chunks: trace every chunk as it lands.  Any other pending upload (a pre-loaded
		// RayHit buffer) must be complete before the first chunk writes.
		lrb_device::Pending &o = dev->pendUpload[1 - k];
		if (o.active && !o.ev.empty())
			LRB_CUDA(cudaStreamWaitEvent(dev->stream, o.ev.back(), 0));
		o.active = false;
		const std::vector<cudaEvent_t> landed = u.ev;
		const std::vector<uint64_t> rayEnds = u.ends;       // bytes of the ray buffer; u.bytes == n * 48, so every end is a whole ray
		u.active = false;       // (the launches below join whatever else is pending; the queue follows this upload chunk by chunk)
		std::vector<cudaEvent_t> traced;
		std::vector<uint64_t> hitEnds;
		uint32_t first = 0;
		for (size_t c = 0; c < rayEnds.size(); ++c) {
			const uint32_t end = (uint32_t)(rayEnds[c] / sizeof(lrb_ray));
			const uint32_t cnt = end - first;
			LRB_CUDA(cudaStreamWaitEvent(dev->stream, landed[c], 0));
			int rc = LaunchTrace(s, (const lrb_ray *)rays + first, (lrb_rayhit *)hits + first, cnt, false, dev->stream);
			if (rc != LRB_OK) {
				for (size_t r = c + 1; r < landed.size(); ++r)      // the rest of the upload still belongs in front of what follows
					cudaStreamWaitEvent(dev->stream, landed[r], 0);
				return rc;
			}
			cudaEvent_t e;
			if ((rc = PipeEvent(dev, &e)) != LRB_OK) return rc;
			LRB_CUDA(cudaEventRecord(e, dev->stream));
			traced.push_back(e);
			hitEnds.push_back((uint64_t)end * sizeof(lrb_rayhit));
			first = end;
		}
		lrb_device::Pending &t = dev->pendTraced;
		t.base = (const char *)hits; t.bytes = (size_t)n * sizeof(lrb_rayhit);
		t.ends.swap(hitEnds);
		t.ev.swap(traced);
		t.active = true;
		return LRB_OK;
	}
	return LaunchTrace(s, rays, hits, n, false, dev->stream);
}

int lrb_trace_anyhit(lrb_scene *s, const void *rays, void *hits, uint32_t n) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	LRB_SETDEV(s->dev);
	TraceMode mode;
	mode.anyhit = true;
	return LaunchTrace(s, rays, hits, n, false, s->dev->stream, nullptr, false, mode);
}

int lrb_compact_rays(lrb_device *dev, const void *rays, uint32_t n, const uint32_t **liveIdxDev, const uint32_t **liveCountDev, uint32_t *liveCountHost) {
	if (!liveIdxDev || !liveCountDev)
		return Fail(LRB_ERR_INVALID, "null out pointer");
	LRB_SETDEV(dev);
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	*liveIdxDev = nullptr; *liveCountDev = nullptr;
	if (liveCountHost) *liveCountHost = 0;
	if (n == 0)
		return LRB_OK;
	if (!rays)
		return Fail(LRB_ERR_INVALID, "null ray buffer");
	const int rc = CompactRays(dev, rays, n, dev->stream, liveIdxDev, liveCountDev);
	if (rc != LRB_OK)
		return rc;
	if (liveCountHost) {
		LRB_CUDA(cudaMemcpyAsync(liveCountHost, *liveCountDev, sizeof(uint32_t), cudaMemcpyDeviceToHost, dev->stream));
		LRB_CUDA(cudaStreamSynchronize(dev->stream));
		dev->counters.d2h_bytes += 4;
	}
	return LRB_OK;
}

int lrb_trace_indexed(lrb_scene *s, const void *rays, void *hits, uint32_t n, const uint32_t *liveIdxDev, const uint32_t *liveCountDev, int anyhit) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	LRB_SETDEV(s->dev);
	if (!liveIdxDev)
		return Fail(LRB_ERR_INVALID, "null index list");
	TraceMode mode;
	mode.anyhit = anyhit != 0;
	mode.liveIdx = liveIdxDev;
	mode.liveCountDev = liveCountDev;
	return LaunchTrace(s, rays, hits, n, false, s->dev->stream, nullptr, false, mode);
}

int lrb_advance_rays(lrb_scene *s, void *rays, void *hits, uint32_t n, const uint32_t *passMeshBitsDev, uint32_t nPassWords,
		const uint8_t *continueFlagsDev, uint32_t *nContinuingHost) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	if (nContinuingHost) *nContinuingHost = 0;
	if (n == 0)
		return LRB_OK;
	if (!rays || !hits)
		return Fail(LRB_ERR_INVALID, "null ray/hit buffer");
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	uint32_t *cnt = s->dCounter + 8;        // a word of the scene's counter block the trace kernels do not use
	LRB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t), dev->stream));
	AdvanceRaysKernel<<<(n + 255) / 256, 256, 0, dev->stream>>>((lrb_ray *)rays, (lrb_rayhit *)hits, n, passMeshBitsDev, nPassWords,
			continueFlagsDev, cnt);
	LRB_CUDA(cudaGetLastError());
	dev->counters.kernel_launches += 1;
	if (nContinuingHost) {
		LRB_CUDA(cudaMemcpyAsync(nContinuingHost, cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, dev->stream));
		LRB_CUDA(cudaStreamSynchronize(dev->stream));
		dev->counters.d2h_bytes += 4;
	}
	return LRB_OK;
}

int lrb_trace_passthrough(lrb_scene *s, void *rays, void *hits, uint32_t n, const uint32_t *passMeshBitsDev, uint32_t nPassWords,
		uint32_t maxRounds, uint32_t *roundsOut, uint64_t *raysTracedOut) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	if (roundsOut) *roundsOut = 0;
	if (raysTracedOut) *raysTracedOut = 0;
	if (n == 0)
		return LRB_OK;
	if (maxRounds == 0) maxRounds = 64;
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	uint32_t live = n, rounds = 0;
	uint64_t traced = 0;
	for (;;) {
		// after the first round most lanes are dead: trace through the dense list of the live ones
		int rc;
		if (rounds == 0)
			rc = LaunchTrace(s, rays, hits, n, false, dev->stream);
		else {
			TraceMode mode;
			if ((rc = CompactRays(dev, rays, n, dev->stream, &mode.liveIdx, &mode.liveCountDev)) != LRB_OK) return rc;
			rc = LaunchTrace(s, rays, hits, n, false, dev->stream, nullptr, false, mode);
		}
		if (rc != LRB_OK)
			return rc;
		traced += live;
		++rounds;
		if ((rc = lrb_advance_rays(s, rays, hits, n, passMeshBitsDev, nPassWords, nullptr, &live)) != LRB_OK)
			return rc;
		if (live == 0 || rounds >= maxRounds)
			break;
	}
	if (roundsOut) *roundsOut = rounds;
	if (raysTracedOut) *raysTracedOut = traced;
	return LRB_OK;
}

int lrb_film_reduce(lrb_device *dev, const float *const *tilesDev, uint32_t nTiles, float *dstDev, uint64_t first, uint64_t count) {
	LRB_SETDEV(dev);
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	if (count == 0)
		return LRB_OK;
	if (!tilesDev || !dstDev || nTiles == 0 || nTiles > (uint32_t)kMaxFilmTiles)
		return Fail(LRB_ERR_INVALID, "film reduce: 1..16 tile pointers and a destination are required");
	FilmTiles t;
	memset(&t, 0, sizeof(t));
	bool aligned = ((reinterpret_cast<uintptr_t>(dstDev) & 15u) == 0) && (first % 4 == 0) && (count % 4 == 0);
	for (uint32_t r = 0; r < nTiles; ++r) {
		if (!tilesDev[r])
			return Fail(LRB_ERR_INVALID, "film reduce: null tile pointer");
		t.tile[r] = tilesDev[r];
		aligned = aligned && (reinterpret_cast<uintptr_t>(tilesDev[r]) & 15u) == 0;
	}
	const uint64_t items = aligned ? count / 4 : count;
	const int grid = (int)std::min<uint64_t>((items + 255) / 256, (uint64_t)dev->prop.multiProcessorCount * 8);
	FilmReduceKernel<<<grid, 256, 0, dev->stream>>>(t, (int)nTiles, dstDev, first, count, aligned ? 1 : 0);
	LRB_CUDA(cudaGetLastError());
	dev->counters.kernel_launches += 1;
	return LRB_OK;
}

// ---- BVH construction on the device (build_kernels.cuh) ---------------------------------------------------

// The tree of lrb_build_bvh from leaf boxes that are already on the device (n >= 2): steps 1 - 7 of build_kernels.cuh.
// The array stays on the device (res->nodes); leaves carry their input index in triangleLeaf.v[0].
struct DeviceTree {
	DevBuf nodes;                   // lrb_bvh_node[total]
	uint32_t total;
	float sortMs, treeMs, emitMs;   // CUDA-event times of the stages
	uint32_t launches;
	DeviceTree() : total(0), sortMs(0.f), treeMs(0.f), emitMs(0.f), launches(0) {}
};

static int BuildTreeOnDevice(lrb_device *dev, const float *dLeafBoxes, const uint32_t n, const uint32_t treeType, const uint32_t quality, DeviceTree *res) {
	const uint32_t nInner = n - 1, nAll = 2 * n - 1;
	cudaStream_t st = dev->stream;
	cudaEvent_t ev[4];
	for (int i = 0; i < 4; ++i) LRB_CUDA(cudaEventCreate(&ev[i]));
	BuildEvents evGuard = { ev, 4 };
	uint32_t launches = 0;

	DevBuf dBounds, dKeys[2], dVals[2], dTemp, dLeft, dRight, dParent, dNodeBox, dSize, dArrived, dKept, dCounters;
	LRB_CUDA(cudaMalloc(&dBounds.p, 32));
	for (int k = 0; k < 2; ++k) {
		LRB_CUDA(cudaMalloc(&dKeys[k].p, (size_t)n * 8));
		LRB_CUDA(cudaMalloc(&dVals[k].p, (size_t)n * 4));
	}
	LRB_CUDA(cudaMalloc(&dLeft.p, (size_t)nAll * 4));
	LRB_CUDA(cudaMalloc(&dRight.p, (size_t)nAll * 4));
	LRB_CUDA(cudaMalloc(&dParent.p, (size_t)nAll * 4));
	LRB_CUDA(cudaMalloc(&dNodeBox.p, (size_t)nAll * 24));
	LRB_CUDA(cudaMalloc(&dSize.p, (size_t)nInner * 4));
	LRB_CUDA(cudaMalloc(&dArrived.p, (size_t)nInner * 4));
	LRB_CUDA(cudaMalloc(&dKept.p, (size_t)nInner));
	LRB_CUDA(cudaMalloc(&dCounters.p, 64));
	const float *dBoxes = dLeafBoxes;

	LRB_CUDA(cudaEventRecord(ev[0], st));

	// 1 + 2: centroid bounds, Morton codes, sort
	const uint32_t initBounds[6] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u };
	LRB_CUDA(cudaMemcpyAsync(dBounds.p, initBounds, sizeof(initBounds), cudaMemcpyHostToDevice, st));
	const int blocks = (int)((n + 255) / 256);
	CentroidBoundsKernel<<<std::min(blocks, dev->prop.multiProcessorCount * 8), 256, 0, st>>>(dBoxes, n, dBounds.as<uint32_t>());
	MortonKernel<<<blocks, 256, 0, st>>>(dBoxes, n, dBounds.as<uint32_t>(), dKeys[0].as<uint64_t>(), dVals[0].as<uint32_t>());
	launches += 2;
	cub::DoubleBuffer<uint64_t> keys(dKeys[0].as<uint64_t>(), dKeys[1].as<uint64_t>());
	cub::DoubleBuffer<uint32_t> vals(dVals[0].as<uint32_t>(), dVals[1].as<uint32_t>());
	size_t tempBytes = 0, selectBytes = 0;
	LRB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keys, vals, (int)n, 0, 63, st));
	if (quality == 1)
		LRB_CUDA(cub::DeviceSelect::Flagged(nullptr, selectBytes, (const uint32_t *)nullptr, (const uint8_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, st));
	LRB_CUDA(cudaMalloc(&dTemp.p, std::max<size_t>(std::max(tempBytes, selectBytes), 16)));
	LRB_CUDA(cub::DeviceRadixSort::SortPairs(dTemp.p, tempBytes, keys, vals, (int)n, 0, 63, st));
	GatherLeafBoxesKernel<<<blocks, 256, 0, st>>>(dBoxes, vals.Current(), n, dNodeBox.as<float>());
	++launches;
	LRB_CUDA(cudaEventRecord(ev[1], st));

	// 3: the binary tree
	uint32_t root = n;
	uint32_t *dCount = dCounters.as<uint32_t>();        // [0] inner nodes created (PLOC) / next frontier size, [1] compaction result
	if (quality == 0) {
		RadixTreeKernel<<<(int)((nInner + 255) / 256), 256, 0, st>>>(keys.Current(), (int)n, dLeft.as<uint32_t>(), dRight.as<uint32_t>(), dParent.as<uint32_t>());
		LRB_CUDA(cudaMemsetAsync(dArrived.p, 0, (size_t)nInner * 4, st));
		BottomUpKernel<true, false><<<blocks, 256, 0, st>>>(n, dLeft.as<uint32_t>(), dRight.as<uint32_t>(), dParent.as<uint32_t>(), nullptr,
				dNodeBox.as<float>(), nullptr, dArrived.as<uint32_t>());
		launches += 2;
	} else {
		// PLOC: cluster lists (two, alternating) + partner positions + keep flags
		DevBuf dClusters[2], dNN, dKeep;
		LRB_CUDA(cudaMalloc(&dClusters[0].p, (size_t)n * 4));
		LRB_CUDA(cudaMalloc(&dClusters[1].p, (size_t)n * 4));
		LRB_CUDA(cudaMalloc(&dNN.p, (size_t)n * 4));
		LRB_CUDA(cudaMalloc(&dKeep.p, (size_t)n));
		IotaKernel<<<blocks, 256, 0, st>>>(dClusters[0].as<uint32_t>(), n);      // the leaves in Morton order: ids 0 .. n-1
		++launches;
		LRB_CUDA(cudaMemsetAsync(dCount, 0, 64, st));
		LRB_CUDA(cudaMemsetAsync(dParent.p, 0xff, (size_t)nAll * 4, st));
		uint32_t m = n;
		int cur = 0;
		const int radius = 16;
		for (uint32_t iter = 0; m > 1; ++iter) {
			if (iter > 4u * 64u + n)
				return Fail(LRB_ERR_INTERNAL, "device builder: clustering does not converge");
			const int mb = (int)((m + 255) / 256);
			PlocNearestKernel<<<mb, 256, 0, st>>>(dClusters[cur].as<uint32_t>(), (int)m, dNodeBox.as<float>(), radius, dNN.as<int>());
			PlocMergeKernel<<<mb, 256, 0, st>>>(dClusters[cur].as<uint32_t>(), (int)m, dNN.as<int>(), n, dNodeBox.as<float>(), dLeft.as<uint32_t>(),
					dRight.as<uint32_t>(), dParent.as<uint32_t>(), dCount, dClusters[1 - cur].as<uint32_t>(), dKeep.as<uint8_t>());
			// compaction in place of the NEXT list: select into the current one (free now), then swap roles
			LRB_CUDA(cub::DeviceSelect::Flagged(dTemp.p, selectBytes, dClusters[1 - cur].as<uint32_t>(), dKeep.as<uint8_t>(), dClusters[cur].as<uint32_t>(),
					dCount + 1, (int)m, st));
			launches += 3;
			uint32_t mNew = 0;
			LRB_CUDA(cudaMemcpyAsync(&mNew, dCount + 1, 4, cudaMemcpyDeviceToHost, st));
			LRB_CUDA(cudaStreamSynchronize(st));
			if (mNew == 0 || mNew >= m)
				return Fail(LRB_ERR_INTERNAL, "device builder: clustering made no progress");
			m = mNew;
		}
		uint32_t created = 0;
		LRB_CUDA(cudaMemcpyAsync(&created, dCount, 4, cudaMemcpyDeviceToHost, st));
		LRB_CUDA(cudaMemcpyAsync(&root, dClusters[cur].p, 4, cudaMemcpyDeviceToHost, st));
		LRB_CUDA(cudaStreamSynchronize(st));
		if (created != nInner || root < n || root >= nAll)
			return Fail(LRB_ERR_INTERNAL, "device builder: inconsistent cluster tree");
	}
	LRB_CUDA(cudaEventRecord(ev[2], st));

	// 4: k-ary collapse over a frontier (two lists, alternating)
	{
		DevBuf dFrontier[2];
		LRB_CUDA(cudaMalloc(&dFrontier[0].p, (size_t)nInner * 4));
		LRB_CUDA(cudaMalloc(&dFrontier[1].p, (size_t)nInner * 4));
		LRB_CUDA(cudaMemsetAsync(dKept.p, 0, (size_t)nInner, st));
		LRB_CUDA(cudaMemcpyAsync(dFrontier[0].p, &root, 4, cudaMemcpyHostToDevice, st));
		uint32_t count = 1;
		int cur = 0;
		for (uint32_t level = 0; count > 0; ++level) {
			if (level > nInner)
				return Fail(LRB_ERR_INTERNAL, "device builder: collapse does not terminate");
			LRB_CUDA(cudaMemsetAsync(dCount, 0, 4, st));
			CollapseKernel<<<(int)((count + 127) / 128), 128, 0, st>>>(dFrontier[cur].as<uint32_t>(), count, n, treeType, dLeft.as<uint32_t>(),
					dRight.as<uint32_t>(), dNodeBox.as<float>(), dKept.as<uint8_t>(), dFrontier[1 - cur].as<uint32_t>(), dCount);
			++launches;
			LRB_CUDA(cudaMemcpyAsync(&count, dCount, 4, cudaMemcpyDeviceToHost, st));
			LRB_CUDA(cudaStreamSynchronize(st));
			cur = 1 - cur;
		}
	}

	// 5: sizes of the collapsed subtrees
	LRB_CUDA(cudaMemsetAsync(dArrived.p, 0, (size_t)nInner * 4, st));
	BottomUpKernel<false, true><<<blocks, 256, 0, st>>>(n, dLeft.as<uint32_t>(), dRight.as<uint32_t>(), dParent.as<uint32_t>(), dKept.as<uint8_t>(),
			dNodeBox.as<float>(), dSize.as<uint32_t>(), dArrived.as<uint32_t>());
	++launches;
	uint32_t total = 0;
	LRB_CUDA(cudaMemcpyAsync(&total, dSize.as<uint32_t>() + (root - n), 4, cudaMemcpyDeviceToHost, st));      // size of the root's subtree = nodes in the array
	LRB_CUDA(cudaStreamSynchronize(st));
	if (total < n + 1 || total > nAll)
		return Fail(LRB_ERR_INTERNAL, "device builder: inconsistent tree size");

	// 6 + 7: array indices and emission
	LRB_CUDA(cudaMalloc(&res->nodes.p, (size_t)total * sizeof(lrb_bvh_node)));
	EmitKernel<<<(int)((nAll + 255) / 256), 256, 0, st>>>(n, vals.Current(), dLeft.as<uint32_t>(), dRight.as<uint32_t>(), dParent.as<uint32_t>(),
			dKept.as<uint8_t>(), dSize.as<uint32_t>(), dNodeBox.as<float>(), res->nodes.as<lrb_bvh_node>());
	++launches;
	LRB_CUDA(cudaEventRecord(ev[3], st));
	LRB_CUDA(cudaStreamSynchronize(st));
	LRB_CUDA(cudaGetLastError());
	res->total = total;
	cudaEventElapsedTime(&res->sortMs, ev[0], ev[1]);
	cudaEventElapsedTime(&res->treeMs, ev[1], ev[2]);
	cudaEventElapsedTime(&res->emitMs, ev[2], ev[3]);
	res->launches = launches;
	return LRB_OK;
}

int lrb_build_bvh(lrb_device *dev, const float *leafBoxes, uint32_t nLeaves, uint32_t treeType, uint32_t quality, lrb_bvh_node *outNodes,
		uint32_t outCapacity, uint32_t *nNodes, lrb_build_timings *timings) {
	if (!leafBoxes || !outNodes || !nNodes)
		return Fail(LRB_ERR_INVALID, "null argument");
	if (treeType != 2 && treeType != 4 && treeType != 8)
		return Fail(LRB_ERR_INVALID, "tree type must be 2, 4 or 8 (bvhaccel.cpp:51)");
	if (quality > 1)
		return Fail(LRB_ERR_INVALID, "builder quality must be 0 (radix tree) or 1 (PLOC)");
	if (nLeaves == 0 || nLeaves >= 0x3fffffffu)
		return Fail(LRB_ERR_INVALID, "leaf count out of range");
	LRB_SETDEV(dev);
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	*nNodes = 0;
	if (timings) memset(timings, 0, sizeof(*timings));
	if (nLeaves == 1) {
		// the tree is its only leaf (bvhclassicbuild.cpp: a leaf list of one)
		if (outCapacity < 1)
			return Fail(LRB_ERR_INVALID, "output array too small");
		memset(outNodes, 0, sizeof(*outNodes));
		outNodes[0].triangleLeaf.v[0] = 0;
		outNodes[0].nodeData = 1u | 0x80000000u;
		*nNodes = 1;
		return LRB_OK;
	}
	const uint32_t n = nLeaves;
	cudaStream_t st = dev->stream;
	cudaEvent_t ev[4];
	for (int i = 0; i < 4; ++i) LRB_CUDA(cudaEventCreate(&ev[i]));
	BuildEvents evGuard = { ev, 4 };

	DevBuf dBoxes;
	LRB_CUDA(cudaMalloc(&dBoxes.p, (size_t)n * 24));
	LRB_CUDA(cudaEventRecord(ev[0], st));
	LRB_CUDA(cudaMemcpyAsync(dBoxes.p, leafBoxes, (size_t)n * 24, cudaMemcpyHostToDevice, st));
	dev->counters.h2d_bytes += (uint64_t)n * 24;
	LRB_CUDA(cudaEventRecord(ev[1], st));

	DeviceTree tree;
	const int rc = BuildTreeOnDevice(dev, dBoxes.as<float>(), n, treeType, quality, &tree);
	if (rc != LRB_OK)
		return rc;
	const uint32_t total = tree.total;
	if (total > outCapacity)
		return Fail(LRB_ERR_INVALID, "output array too small (2 * leaves - 1 nodes always suffice)");

	LRB_CUDA(cudaEventRecord(ev[2], st));
	LRB_CUDA(cudaMemcpyAsync(outNodes, tree.nodes.p, (size_t)total * sizeof(lrb_bvh_node), cudaMemcpyDeviceToHost, st));
	dev->counters.d2h_bytes += (uint64_t)total * sizeof(lrb_bvh_node);
	LRB_CUDA(cudaEventRecord(ev[3], st));
	LRB_CUDA(cudaStreamSynchronize(st));
	LRB_CUDA(cudaGetLastError());
	*nNodes = total;
	if (timings) {
		float ms;
		cudaEventElapsedTime(&ms, ev[0], ev[1]); timings->h2d_ms = ms;
		timings->sort_ms = tree.sortMs;
		timings->tree_ms = tree.treeMs;
		timings->emit_ms = tree.emitMs;
		cudaEventElapsedTime(&ms, ev[2], ev[3]); timings->d2h_ms = ms;
		timings->kernels = tree.launches;
	}
	return LRB_OK;
}

int lrb_build_lbvh(lrb_device *dev, const float *leafBoxes, uint32_t nLeaves, uint32_t treeType, lrb_bvh_node *outNodes,
		uint32_t outCapacity, uint32_t *nNodes, lrb_build_timings *timings) {
	return lrb_build_bvh(dev, leafBoxes, nLeaves, treeType, 0u, outNodes, outCapacity, nNodes, timings);
}

int lrb_bvh_build_scene(lrb_device *dev, const float *xyz, uint64_t nVerts, const uint32_t *meshVertexOffsets, const uint32_t *meshTriangleOffsets,
		uint32_t nMeshes, const uint32_t *triangles, uint32_t treeType, uint32_t quality, lrb_scene **out, lrb_bvh_node *outNodes, uint32_t outCapacity,
		uint32_t *nNodes, lrb_scene_build_timings *timings) {
	if (!out)
		return Fail(LRB_ERR_INVALID, "null out pointer");
	*out = nullptr;
	if (nNodes) *nNodes = 0;
	if (timings) memset(timings, 0, sizeof(*timings));
	if (!xyz || !meshVertexOffsets || !meshTriangleOffsets || nMeshes == 0)
		return Fail(LRB_ERR_INVALID, "scene build needs vertices, a mesh vertex-offset table and a mesh triangle-offset table");
	if (treeType != 2 && treeType != 4 && treeType != 8)
		return Fail(LRB_ERR_INVALID, "tree type must be 2, 4 or 8 (bvhaccel.cpp:51)");
	if (quality > 1)
		return Fail(LRB_ERR_INVALID, "builder quality must be 0 (radix tree) or 1 (PLOC)");
	if (meshTriangleOffsets[0] != 0)
		return Fail(LRB_ERR_INVALID, "mesh triangle offsets must start at 0");
	for (uint32_t m = 0; m < nMeshes; ++m) {
		if (meshTriangleOffsets[m + 1] < meshTriangleOffsets[m])
			return Fail(LRB_ERR_INVALID, "mesh triangle offsets must not decrease");
		if (meshVertexOffsets[m] > nVerts)
			return Fail(LRB_ERR_INVALID, "mesh vertex offset outside the vertex buffer");
	}
	const uint32_t nTris = meshTriangleOffsets[nMeshes];
	if (nTris >= 0x3fffffffu)
		return Fail(LRB_ERR_INVALID, "triangle count out of range");
	if (nTris && !triangles)
		return Fail(LRB_ERR_INVALID, "null triangle index array");
	LRB_SETDEV(dev);
	if (nTris <= 1) {
		// nothing to build: an empty array, or the tree that is its only leaf (bvhclassicbuild.cpp: a leaf list of one)
		lrb_bvh_node leaf;
		memset(&leaf, 0, sizeof(leaf));
		if (nTris == 1) {
			const uint32_t m = MeshOfTriangle(meshTriangleOffsets, nMeshes, 0u);
			for (int j = 0; j < 3; ++j) leaf.triangleLeaf.v[j] = triangles[j];
			leaf.triangleLeaf.meshIndex = m;
			leaf.triangleLeaf.triangleIndex = 0u - meshTriangleOffsets[m];
			leaf.nodeData = 1u | 0x80000000u;
			if (outNodes) {
				if (outCapacity < 1)
					return Fail(LRB_ERR_INVALID, "output array too small");
				outNodes[0] = leaf;
			}
		}
		if (nNodes) *nNodes = nTris;
		return lrb_bvh_upload(dev, &leaf, nTris, xyz, nVerts, meshVertexOffsets, nMeshes, out);
	}
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	cudaStream_t st = dev->stream;
	cudaEvent_t ev[6];
	for (int i = 0; i < 6; ++i) LRB_CUDA(cudaEventCreate(&ev[i]));
	BuildEvents evGuard = { ev, 6 };
	uint32_t launches = 0;

	// inputs
	DevBuf dXyz, dTriIdx, dMeshVertOff, dMeshTriOff, dBoxes, dErr;
	LRB_CUDA(cudaMalloc(&dXyz.p, std::max<size_t>((size_t)nVerts * 12, 16)));
	LRB_CUDA(cudaMalloc(&dTriIdx.p, (size_t)nTris * 12));
	LRB_CUDA(cudaMalloc(&dMeshVertOff.p, (size_t)nMeshes * 4));
	LRB_CUDA(cudaMalloc(&dMeshTriOff.p, ((size_t)nMeshes + 1) * 4));
	LRB_CUDA(cudaMalloc(&dBoxes.p, (size_t)nTris * 24));
	LRB_CUDA(cudaMalloc(&dErr.p, 64));       // [0] error code, [1] stack bound, [2..7] exact root box
	LRB_CUDA(cudaEventRecord(ev[0], st));
	LRB_CUDA(cudaMemcpyAsync(dXyz.p, xyz, (size_t)nVerts * 12, cudaMemcpyHostToDevice, st));
	LRB_CUDA(cudaMemcpyAsync(dTriIdx.p, triangles, (size_t)nTris * 12, cudaMemcpyHostToDevice, st));
	LRB_CUDA(cudaMemcpyAsync(dMeshVertOff.p, meshVertexOffsets, (size_t)nMeshes * 4, cudaMemcpyHostToDevice, st));
	LRB_CUDA(cudaMemcpyAsync(dMeshTriOff.p, meshTriangleOffsets, ((size_t)nMeshes + 1) * 4, cudaMemcpyHostToDevice, st));
	LRB_CUDA(cudaMemsetAsync(dErr.p, 0, 64, st));
	dev->counters.h2d_bytes += (uint64_t)nVerts * 12 + (uint64_t)nTris * 12 + (uint64_t)nMeshes * 8 + 4;
	LRB_CUDA(cudaEventRecord(ev[1], st));

	// 0: the triangles' build boxes
	const int triBlocks = (int)((nTris + 255) / 256);
	LeafBoxKernel<<<triBlocks, 256, 0, st>>>(dXyz.as<float>(), nVerts, dMeshVertOff.as<uint32_t>(), dMeshTriOff.as<uint32_t>(), nMeshes, dTriIdx.as<uint32_t>(),
			nTris, dBoxes.as<float>(), dErr.as<uint32_t>());
	++launches;
	LRB_CUDA(cudaEventRecord(ev[2], st));
	uint32_t errCode = 0;
	LRB_CUDA(cudaMemcpyAsync(&errCode, dErr.p, 4, cudaMemcpyDeviceToHost, st));
	LRB_CUDA(cudaStreamSynchronize(st));
	if (errCode != kRelayoutOk)
		return Fail(LRB_ERR_INVALID, RelayoutErrorString((int)errCode));

	// the tree
	DeviceTree tree;
	{
		const int rc = BuildTreeOnDevice(dev, dBoxes.as<float>(), nTris, treeType, quality, &tree);
		if (rc != LRB_OK)
			return rc;
	}
	launches += tree.launches;
	cudaFree(dBoxes.p);
	dBoxes.p = nullptr;
	const uint32_t total = tree.total;
	if (outNodes && total > outCapacity)
		return Fail(LRB_ERR_INVALID, "output array too small (2 * triangles - 1 nodes always suffice)");
	lrb_bvh_node *dNodes = tree.nodes.as<lrb_bvh_node>();
	const int nodeBlocks = (int)((total + 255) / 256);

	LRB_CUDA(cudaEventRecord(ev[3], st));
	// 1: leaf payload
	LeafPayloadKernel<<<nodeBlocks, 256, 0, st>>>(dNodes, total, dMeshTriOff.as<uint32_t>(), nMeshes, dTriIdx.as<uint32_t>());
	++launches;

	// 2: indices of the wide nodes and of the triangle records
	DevBuf dCounts, dScanned, dScanTemp, dWideOf;
	LRB_CUDA(cudaMalloc(&dCounts.p, (size_t)total * 8));
	LRB_CUDA(cudaMalloc(&dScanned.p, (size_t)total * 8));
	LRB_CUDA(cudaMalloc(&dWideOf.p, (size_t)total * 4));
	size_t scanBytes = 0;
	LRB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, dCounts.as<unsigned long long>(), dScanned.as<unsigned long long>(), (int)total, st));
	LRB_CUDA(cudaMalloc(&dScanTemp.p, std::max<size_t>(scanBytes, 16)));
	RelayoutCountKernel<<<nodeBlocks, 256, 0, st>>>(dNodes, total, dCounts.as<unsigned long long>());
	LRB_CUDA(cub::DeviceScan::ExclusiveSum(dScanTemp.p, scanBytes, dCounts.as<unsigned long long>(), dScanned.as<unsigned long long>(), (int)total, st));
	RelayoutIndexKernel<<<nodeBlocks, 256, 0, st>>>(dNodes, total, dScanned.as<unsigned long long>(), dWideOf.as<uint32_t>());
	launches += 2;
	unsigned long long lastScan = 0, lastCount = 0;
	LRB_CUDA(cudaMemcpyAsync(&lastScan, dScanned.as<unsigned long long>() + (total - 1), 8, cudaMemcpyDeviceToHost, st));
	LRB_CUDA(cudaMemcpyAsync(&lastCount, dCounts.as<unsigned long long>() + (total - 1), 8, cudaMemcpyDeviceToHost, st));
	LRB_CUDA(cudaStreamSynchronize(st));
	const unsigned long long sums = lastScan + lastCount;
	const uint64_t nLeafRecords = sums & 0xffffffffull, nWide64 = 1ull + (sums >> 32);
	if (nLeafRecords != nTris)
		return Fail(LRB_ERR_INTERNAL, "device re-layout: leaf count of the built tree differs from the triangle count");
	if (nWide64 >= kMaxRefIndex || nLeafRecords >= kMaxRefIndex)
		return Fail(LRB_ERR_INVALID, "too many wide nodes");
	const uint32_t nWide = (uint32_t)nWide64;
	cudaFree(dCounts.p); dCounts.p = nullptr;
	cudaFree(dScanned.p); dScanned.p = nullptr;

	// 3 + 4: the scene's arrays, written in place
	lrb_scene *s = NewScene(dev);
	const size_t nb = ((size_t)nWide * sizeof(WideNode) + 255) & ~(size_t)255;
	const size_t tb = ((size_t)nTris * sizeof(TriRecord) + 255) & ~(size_t)255;
	const size_t ib = ((size_t)nTris * sizeof(TriIds) + 255) & ~(size_t)255;
	DevBuf dParentOf, dBelow, dArrived;
	cudaError_t ce = cudaMalloc(&s->dSlab, nb + tb + ib);
	if (ce == cudaSuccess) ce = cudaMalloc((void **)&s->dCounter, 256);
	if (ce == cudaSuccess) ce = cudaMalloc((void **)&s->dStats, sizeof(TraceStats));
	if (ce == cudaSuccess) ce = cudaMalloc(&dParentOf.p, (size_t)nWide * 4);
	if (ce == cudaSuccess) ce = cudaMalloc(&dBelow.p, (size_t)nWide * 4);
	if (ce == cudaSuccess) ce = cudaMalloc(&dArrived.p, (size_t)nWide * 4);
	if (ce != cudaSuccess) {
		cudaGetLastError();
		lrb_scene_free(s);
		return Fail(ce == cudaErrorMemoryAllocation ? LRB_ERR_OOM : LRB_ERR_CUDA, std::string("device re-layout: ") + cudaGetErrorString(ce));
	}
	s->slabBytes = nb + tb + ib;
	s->dNodes = reinterpret_cast<WideNode *>(s->dSlab);
	s->dTris = reinterpret_cast<TriRecord *>((char *)s->dSlab + nb);
	s->dIds = reinterpret_cast<TriIds *>((char *)s->dSlab + nb + tb);
	s->capNodes = nWide;
	s->info.device_bytes = s->slabBytes + 256 + sizeof(TraceStats);
	{
		std::lock_guard<std::mutex> g(dev->mtx);
		dev->counters.device_bytes_in_use += s->info.device_bytes;
	}
	TriTreeView tv;
	tv.nodes = dNodes; tv.n = total; tv.xyz = dXyz.as<float>(); tv.nVerts = nVerts; tv.meshOff = dMeshVertOff.as<uint32_t>(); tv.nMeshes = nMeshes;
	uint32_t *dStatus = dErr.as<uint32_t>();
	cudaMemsetAsync(dBelow.p, 0, (size_t)nWide * 4, st);
	cudaMemsetAsync(dArrived.p, 0, (size_t)nWide * 4, st);
	cudaMemsetAsync(dParentOf.p, 0xff, (size_t)nWide * 4, st);
	RelayoutFillKernel<<<(int)((total + 127) / 128), 128, 0, st>>>(tv, dWideOf.as<uint32_t>(), s->dNodes, s->dTris, s->dIds, dParentOf.as<uint32_t>(),
			reinterpret_cast<float *>(dStatus + 2), dStatus);
	StackNeedKernel<<<(int)((nWide + 255) / 256), 256, 0, st>>>(s->dNodes, nWide, dParentOf.as<uint32_t>(), dBelow.as<uint32_t>(), dArrived.as<uint32_t>(), dStatus + 1);
	launches += 2;
	cudaEventRecord(ev[4], st);
	uint32_t status[8];
	memset(status, 0, sizeof(status));
	cudaMemcpyAsync(status, dStatus, sizeof(status), cudaMemcpyDeviceToHost, st);
	if (outNodes) {
		cudaMemcpyAsync(outNodes, dNodes, (size_t)total * sizeof(lrb_bvh_node), cudaMemcpyDeviceToHost, st);
		dev->counters.d2h_bytes += (uint64_t)total * sizeof(lrb_bvh_node);
	}
	cudaEventRecord(ev[5], st);
	ce = cudaStreamSynchronize(st);
	if (ce == cudaSuccess) ce = cudaGetLastError();
	if (ce != cudaSuccess) {
		lrb_scene_free(s);
		return Fail(LRB_ERR_CUDA, std::string("device re-layout: ") + cudaGetErrorString(ce));
	}
	if (status[0] != kRelayoutOk) {
		lrb_scene_free(s);
		return Fail(LRB_ERR_INVALID, RelayoutErrorString((int)status[0]));
	}
	if (status[1] < 2 * (kWideSlots - 1)) {     // entry node + root: the smallest tree already needs this much
		lrb_scene_free(s);
		return Fail(LRB_ERR_INTERNAL, "device re-layout: the stack bound was not computed");
	}
	dev->counters.kernel_launches += launches;

	// bookkeeping the host path takes from its WideScene (FillView / FillRootOfView)
	s->host.twoLevel = false;
	s->host.nRefNodes = total;
	s->host.rootWide = 0;
	s->host.stackNeed = status[1] + 1;
	memcpy(s->host.entryBox, status + 2, 24);
	SceneView &v = s->view;
	v.nodes = s->dNodes;
	v.tris = s->dTris;
	v.ids = s->dIds;
	v.nWide = nWide;
	v.rootWide = 0;
	v.twoLevel = 0;
	v.rootHasBox = 1;
	v.rootChild = 1;            // the root is the first inner node of the array: wide node 1, behind the entry node
	memcpy(v.rootBox, status + 2, 24);
	v.oneBits = 0x3f800000u;
	s->info.n_ref_nodes = total;
	s->info.n_wide_nodes = nWide;
	s->info.n_triangles = nTris;
	s->info.n_instances = 0;
	s->info.stack_need = s->host.stackNeed;
	s->info.two_level = 0;
	if (nNodes) *nNodes = total;
	if (timings) {
		float ms;
		cudaEventElapsedTime(&ms, ev[0], ev[1]); timings->h2d_ms = ms;
		cudaEventElapsedTime(&ms, ev[1], ev[2]); timings->leafbox_ms = ms;
		timings->sort_ms = tree.sortMs;
		timings->tree_ms = tree.treeMs;
		timings->emit_ms = tree.emitMs;
		cudaEventElapsedTime(&ms, ev[3], ev[4]); timings->relayout_ms = ms;
		cudaEventElapsedTime(&ms, ev[4], ev[5]); timings->d2h_ms = outNodes ? ms : 0.0;
		timings->kernels = launches;
	}
	*out = s;
	return LRB_OK;
}

int lrb_scene_download(lrb_scene *s, void *wide, void *tris, void *ids) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	if (s->view.twoLevel)
		return Fail(LRB_ERR_INVALID, "lrb_scene_download serves single-level scenes");
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	{
		const int rc = JoinPending(dev);
		if (rc != LRB_OK) return rc;
	}
	if (wide && s->info.n_wide_nodes)
		LRB_CUDA(cudaMemcpyAsync(wide, s->dNodes, (size_t)s->info.n_wide_nodes * sizeof(WideNode), cudaMemcpyDeviceToHost, dev->stream));
	if (tris && s->info.n_triangles)
		LRB_CUDA(cudaMemcpyAsync(tris, s->dTris, (size_t)s->info.n_triangles * sizeof(TriRecord), cudaMemcpyDeviceToHost, dev->stream));
	if (ids && s->info.n_triangles)
		LRB_CUDA(cudaMemcpyAsync(ids, s->dIds, (size_t)s->info.n_triangles * sizeof(TriIds), cudaMemcpyDeviceToHost, dev->stream));
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	return LRB_OK;
}

int lrb_scene_adopt(lrb_device *dev, lrb_scene *s) {
	if (!dev || !s)
		return Fail(LRB_ERR_INVALID, "null argument");
	lrb_device *old = s->dev;
	if (old == dev)
		return LRB_OK;
	if (old->ordinal != dev->ordinal)
		return Fail(LRB_ERR_INVALID, "a scene can only be handed to a device handle of the same CUDA device");
	LRB_SETDEV(old);
	{
		const int rc = JoinPending(old);
		if (rc != LRB_OK) return rc;
	}
	LRB_CUDA(cudaStreamSynchronize(old->stream));
	if (old->l2WindowBase && old->l2WindowBase == s->dSlab)
		ClearL2Window(old);
	{
		std::lock_guard<std::mutex> g(old->mtx);
		old->counters.device_bytes_in_use -= std::min<uint64_t>(old->counters.device_bytes_in_use, s->info.device_bytes);
	}
	{
		std::lock_guard<std::mutex> g(dev->mtx);
		dev->counters.device_bytes_in_use += s->info.device_bytes;
	}
	s->dev = dev;
	return LRB_OK;
}

int lrb_trace_stats(lrb_scene *s, const void *rays, void *hits, uint32_t n, lrb_trace_stats_t *out) {
	if (!s || !out)
		return Fail(LRB_ERR_INVALID, "null argument");
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	memset(out, 0, sizeof(*out));
	if (n == 0)
		return LRB_OK;
	const int rc = LaunchTrace(s, rays, hits, n, true, dev->stream);
	if (rc != LRB_OK)
		return rc;
	TraceStats st;
	LRB_CUDA(cudaMemcpyAsync(&st, s->dStats, sizeof(st), cudaMemcpyDeviceToHost, dev->stream));
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	out->rays = st.rays;
	out->wide_nodes = st.wideNodes;
	out->triangles = st.triangles;
	out->instances = st.instances;
	out->motion_samples = st.motionSamples;
	out->max_stack = st.maxStack;
	return LRB_OK;
}

int lrb_ipc_get_handle(lrb_device *dev, void *devptr, unsigned char handle[LRB_IPC_HANDLE_BYTES]) {
	if (!devptr || !handle)
		return Fail(LRB_ERR_INVALID, "null argument");
	LRB_SETDEV(dev);
	static_assert(sizeof(cudaIpcMemHandle_t) == LRB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
	cudaIpcMemHandle_t h;
	LRB_CUDA(cudaIpcGetMemHandle(&h, devptr));
	memcpy(handle, &h, sizeof(h));
	return LRB_OK;
}

int lrb_ipc_open_handle(lrb_device *dev, const unsigned char handle[LRB_IPC_HANDLE_BYTES], void **devptr) {
	if (!devptr || !handle)
		return Fail(LRB_ERR_INVALID, "null argument");
	*devptr = nullptr;
	LRB_SETDEV(dev);
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	// maps the exporter's allocation into this process and enables peer access from this device
	LRB_CUDA(cudaIpcOpenMemHandle(devptr, h, cudaIpcMemLazyEnablePeerAccess));
	return LRB_OK;
}

int lrb_ipc_close_handle(lrb_device *dev, void *devptr) {
	LRB_SETDEV(dev);
	if (devptr)
		LRB_CUDA(cudaIpcCloseMemHandle(devptr));
	return LRB_OK;
}

}   // extern "C"

// End of a gather's pushes on the copy stream.  Default: the queue waits for them (later work is ordered after
// the gather).  Deferred (device option gather_defer = 1): the queue does NOT wait -- the next trace may start
// while the tail of this gather's pushes is still on the wire; the pushes of call k are waited for at the start
// of call k + 2 (the caller alternates two RayHit buffers), by lrb_gather_wait, or by lrb_sync.
static int GatherBegin(lrb_device *dev) {
	if (!dev->gatherDefer)
		return LRB_OK;
	const int slot = dev->gatherSeq & 1u;
	if (dev->gatherPending[slot]) {
		LRB_CUDA(cudaStreamWaitEvent(dev->stream, dev->gatherDone[slot], 0));
		dev->gatherPending[slot] = false;
	}
	return LRB_OK;
}

static int GatherEnd(lrb_device *dev, cudaEvent_t fallback) {
	if (!dev->gatherDefer) {
		LRB_CUDA(cudaEventRecord(fallback, dev->copyOutStream));
		LRB_CUDA(cudaStreamWaitEvent(dev->stream, fallback, 0));
		return LRB_OK;
	}
	const int slot = dev->gatherSeq & 1u;
	if (!dev->gatherDone[slot])
		LRB_CUDA(cudaEventCreateWithFlags(&dev->gatherDone[slot], cudaEventDisableTiming));
	LRB_CUDA(cudaEventRecord(dev->gatherDone[slot], dev->copyOutStream));
	dev->gatherPending[slot] = true;
	++dev->gatherSeq;
	return LRB_OK;
}

extern "C" {

int lrb_gather_wait(lrb_device *dev, void *cudaStream, int which) {
	LRB_SETDEV(dev);
	cudaStream_t st = cudaStream ? (cudaStream_t)cudaStream : dev->stream;
	// which: 0 = the most recent lrb_trace_gather, 1 = the one before it, -1 = both (and forget them)
	for (int back = 0; back < 2; ++back) {
		if (which >= 0 && which != back)
			continue;
		const int slot = (dev->gatherSeq + 1u - (unsigned)back) & 1u;      // seq was incremented after the record
		if (dev->gatherDone[slot] && dev->gatherPending[slot]) {
			LRB_CUDA(cudaStreamWaitEvent(st, dev->gatherDone[slot], 0));
			if (which < 0 && st == dev->stream)
				dev->gatherPending[slot] = false;
		}
	}
	return LRB_OK;
}

// Completion signal of a gather WITHOUT a kernel: the value travels as a 4-byte copy-engine transfer on the push stream,
// behind this rank's pushes, into a flag word on the gathering GPU (peer-mapped like the gather buffer itself).
// (Round-2 timeline: the 4-byte ncclAllReduce used as the signal needs SMs; against a persistent trace kernel that
// fills every SM it can only run between two trace kernels, and cost 0.25 ms of every 5 ms step at >= 4 GPUs.)
int lrb_gather_signal(lrb_device *dev, void *flagDev, uint32_t value) {
	if (!dev || !flagDev)
		return Fail(LRB_ERR_INVALID, "null argument");
	LRB_SETDEV(dev);
	if (!dev->signalValues) {
		// a table of the values 0 .. 65535: the source of the 4-byte copies (the value must live in device memory)
		std::vector<uint32_t> v(65536);
		for (uint32_t i = 0; i < 65536u; ++i) v[i] = i;
		LRB_CUDA(cudaMalloc((void **)&dev->signalValues, v.size() * 4));
		LRB_CUDA(cudaMemcpy(dev->signalValues, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
	}
	if (value >= 65536u)
		return Fail(LRB_ERR_INVALID, "signal values are 16-bit step counters");
	// ordered behind the pushes of every gather issued so far (they run on the same stream) and behind the queue itself
	// (a gather without pushes -- the gathering rank's own -- still signals that its trace has been queued ... and run)
	if (!dev->pipeJoin)
		LRB_CUDA(cudaEventCreateWithFlags(&dev->pipeJoin, cudaEventDisableTiming));
	LRB_CUDA(cudaEventRecord(dev->pipeJoin, dev->stream));
	LRB_CUDA(cudaStreamWaitEvent(dev->copyOutStream, dev->pipeJoin, 0));
	LRB_CUDA(cudaMemcpyAsync(flagDev, dev->signalValues + value, 4, cudaMemcpyDefault, dev->copyOutStream));
	return LRB_OK;
}

// Makes a stream (NULL = the device's queue) wait until the 32-bit word at flagDev (memory of THIS device) is >= value.
int lrb_wait_value(lrb_device *dev, void *flagDev, uint32_t value, void *cudaStream) {
	if (!dev || !flagDev)
		return Fail(LRB_ERR_INVALID, "null argument");
	LRB_SETDEV(dev);
	const int rc = ResolveDriverEntryPoints();
	if (rc != LRB_OK)
		return rc;
	cudaStream_t st = cudaStream ? (cudaStream_t)cudaStream : dev->stream;
	if (g_streamWaitValue32((CUstream)st, (CUdeviceptr)(uintptr_t)flagDev, value, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
		return Fail(LRB_ERR_CUDA, "cuStreamWaitValue32 failed");
	return LRB_OK;
}

int lrb_trace_gather(lrb_scene *s, const void *rays, void *hits, uint32_t n, void *dst, uint32_t nChunks) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	if (n == 0)
		return LRB_OK;
	if (!rays || !dst || (!hits && nChunks != 0))
		return Fail(LRB_ERR_INVALID, "null buffer");
	if (nChunks == 0) {
		if (dst == hits)
			return LaunchTrace(s, rays, hits, n, false, dev->stream);
		// a batch of at most one block has no room for the detector warp of the signalled form
		if (dev->gatherStores || !hits || !dev->persistent || n <= (uint32_t)kTraceBlock) {
			// ONE kernel; every lane stores its RayHit into the local buffer (if any) and into the gather slice
			return LaunchTrace(s, rays, hits, n, false, dev->stream, (lrb_rayhit *)dst);
		}
		// ONE kernel + signalled pushes: the kernel raises a flag per completed chunk of ray indices; the
		// copy stream waits on each flag (stream-ordered, no host involvement) and pushes that chunk of
		// the RayHit buffer with the copy engine while the kernel keeps tracing.
		int rc = ResolveDriverEntryPoints();
		if (rc != LRB_OK)
			return rc;
		while (dev->events.size() < 2) {
			cudaEvent_t e;
			LRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			dev->events.push_back(e);
		}
		if ((rc = GatherBegin(dev)) != LRB_OK)
			return rc;
		// the copy stream starts after everything queued so far (previous readers of dst / hits)
		LRB_CUDA(cudaEventRecord(dev->events[0], dev->stream));
		LRB_CUDA(cudaStreamWaitEvent(dev->copyOutStream, dev->events[0], 0));
		if ((rc = LaunchTrace(s, rays, hits, n, false, dev->stream, nullptr, true)) != LRB_OK)
			return rc;
		const uint32_t epoch = s->epoch;
		const uint32_t chunkRays = 1u << dev->gatherChunkShift;
		const uint32_t nC = (n + chunkRays - 1) >> dev->gatherChunkShift;
		// safety net: once the kernel has finished every flag is raised, whatever happened inside it
		if (g_memsetD32Async((CUdeviceptr)(uintptr_t)s->dChunkFlag, epoch, nC, (CUstream)dev->stream) != CUDA_SUCCESS)
			return Fail(LRB_ERR_CUDA, "cuMemsetD32Async failed");
		for (uint32_t c = 0; c < nC; ++c) {
			const uint32_t first = c << dev->gatherChunkShift, cnt = std::min(chunkRays, n - first);
			if (g_streamWaitValue32((CUstream)dev->copyOutStream, (CUdeviceptr)(uintptr_t)(s->dChunkFlag + c), epoch, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
				return Fail(LRB_ERR_CUDA, "cuStreamWaitValue32 failed");
			LRB_CUDA(cudaMemcpyAsync((lrb_rayhit *)dst + first, (lrb_rayhit *)hits + first, (size_t)cnt * sizeof(lrb_rayhit),
					cudaMemcpyDefault, dev->copyOutStream));
		}
		return GatherEnd(dev, dev->events[1]);
	}
	if (nChunks > 1024) nChunks = 1024;
	// chunk boundaries on multiples of 4 rays keep every RayHit range 16-byte aligned (4 x 20 B)
	uint32_t per = ((n + nChunks - 1) / nChunks + 3u) & ~3u;
	const bool push = dst != hits;
	while (push && dev->events.size() < 2 * (size_t)nChunks + 2) {
		cudaEvent_t e;
		LRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		dev->events.push_back(e);
	}
	if (push) {
		// (deferred gathers: the pushes of the call before the previous one must have left this RayHit buffer)
		const int rc = GatherBegin(dev);
		if (rc != LRB_OK)
			return rc;
	}
	uint32_t c = 0;
	for (uint32_t first = 0; first < n; first += per, ++c) {
		const uint32_t cnt = std::min(per, n - first);
		const lrb_ray *r = (const lrb_ray *)rays + first;
		lrb_rayhit *h = (lrb_rayhit *)hits + first;
		const int rc = LaunchTrace(s, r, h, cnt, false, dev->stream);
		if (rc != LRB_OK)
			return rc;
		if (push) {
			cudaEvent_t traced = dev->events[2 * c];
			LRB_CUDA(cudaEventRecord(traced, dev->stream));
			LRB_CUDA(cudaStreamWaitEvent(dev->copyOutStream, traced, 0));
			LRB_CUDA(cudaMemcpyAsync((lrb_rayhit *)dst + first, h, (size_t)cnt * sizeof(lrb_rayhit), cudaMemcpyDefault, dev->copyOutStream));
		}
	}
	if (push)
		return GatherEnd(dev, dev->events[2 * (size_t)nChunks + 1]);
	return LRB_OK;
}

// Host buffers in, host buffers out.  The batch is cut into chunks; chunk k+1 is copied in and
// chunk k-1 copied out (separate streams, PCIe is full duplex) while chunk k is traced.
int lrb_trace_host(lrb_scene *s, const lrb_ray *rays, lrb_rayhit *hits, uint32_t n, int preloadHits) {
	if (!s)
		return Fail(LRB_ERR_INVALID, "null scene");
	lrb_device *dev = s->dev;
	LRB_SETDEV(dev);
	{
		const int rcJoin = JoinPending(dev);
		if (rcJoin != LRB_OK) return rcJoin;
	}
	if (n == 0)
		return LRB_OK;
	if (!rays || !hits)
		return Fail(LRB_ERR_INVALID, "null host buffer");
	const size_t rb = (size_t)n * sizeof(lrb_ray), hb = (size_t)n * sizeof(lrb_rayhit);
	if (dev->stageRaysBytes < rb) {
		LRB_CUDA(cudaStreamSynchronize(dev->stream));
		if (dev->stageRays) cudaFree(dev->stageRays);
		dev->stageRays = nullptr; dev->stageRaysBytes = 0;
		LRB_CUDA(cudaMalloc(&dev->stageRays, rb));
		dev->stageRaysBytes = rb;
	}
	if (dev->stageHitsBytes < hb) {
		LRB_CUDA(cudaStreamSynchronize(dev->stream));
		if (dev->stageHits) cudaFree(dev->stageHits);
		dev->stageHits = nullptr; dev->stageHitsBytes = 0;
		LRB_CUDA(cudaMalloc(&dev->stageHits, hb));
		dev->stageHitsBytes = hb;
	}
	const bool anyMasked = preloadHits != 0;
	if (!anyMasked)     // masked rays leave their record untouched: make what is read back for them deterministic
		LRB_CUDA(cudaMemsetAsync(dev->stageHits, 0, hb, dev->stream));

	std::vector<uint64_t> ends;
	ChunkEnds(n, (uint64_t)dev->hostChunk, (uint64_t)dev->hostMinChunk, dev->hostTaper != 0, &ends);
	const uint32_t nChunks = (uint32_t)ends.size();
	while (dev->events.size() < 3 * (size_t)nChunks + 1) {
		cudaEvent_t e;
		LRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		dev->events.push_back(e);
	}
	// everything queued so far on the main stream must be visible to the copy streams
	cudaEvent_t start = dev->events[3 * (size_t)nChunks];
	LRB_CUDA(cudaEventRecord(start, dev->stream));
	LRB_CUDA(cudaStreamWaitEvent(dev->copyInStream, start, 0));
	LRB_CUDA(cudaStreamWaitEvent(dev->copyOutStream, start, 0));
	for (uint32_t c = 0; c < nChunks; ++c) {
		const uint32_t first = c ? (uint32_t)ends[c - 1] : 0u, cnt = (uint32_t)ends[c] - first;
		lrb_ray *dR = (lrb_ray *)dev->stageRays + first;
		lrb_rayhit *dH = (lrb_rayhit *)dev->stageHits + first;
		cudaEvent_t in = dev->events[3 * c], done = dev->events[3 * c + 1];
		LRB_CUDA(cudaMemcpyAsync(dR, rays + first, (size_t)cnt * sizeof(lrb_ray), cudaMemcpyHostToDevice, dev->copyInStream));
		if (anyMasked)
			LRB_CUDA(cudaMemcpyAsync(dH, hits + first, (size_t)cnt * sizeof(lrb_rayhit), cudaMemcpyHostToDevice, dev->copyInStream));
		LRB_CUDA(cudaEventRecord(in, dev->copyInStream));
		LRB_CUDA(cudaStreamWaitEvent(dev->stream, in, 0));
		const int rc = LaunchTrace(s, dR, dH, cnt, false, dev->stream);
		if (rc != LRB_OK)
			return rc;
		LRB_CUDA(cudaEventRecord(done, dev->stream));
		LRB_CUDA(cudaStreamWaitEvent(dev->copyOutStream, done, 0));
		LRB_CUDA(cudaMemcpyAsync(hits + first, dH, (size_t)cnt * sizeof(lrb_rayhit), cudaMemcpyDeviceToHost, dev->copyOutStream));
	}
	cudaEvent_t out = dev->events[3 * (size_t)(nChunks - 1) + 2];
	LRB_CUDA(cudaEventRecord(out, dev->copyOutStream));
	LRB_CUDA(cudaStreamWaitEvent(dev->stream, out, 0));
	LRB_CUDA(cudaStreamSynchronize(dev->stream));
	dev->counters.h2d_bytes += rb + (anyMasked ? hb : 0);
	dev->counters.d2h_bytes += hb;
	return LRB_OK;
}

}   // extern "C"
