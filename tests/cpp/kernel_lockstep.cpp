// kernel_lockstep.cpp -- TEST SCAFFOLDING (never shipped).  Runs the product's REAL kernel source,
// luxcore_b200/csrc/trace_kernels.cuh, on the host: compiled by g++ against tests/cpp/fakecuda/
// cuda_runtime.h, one OS thread per lane, warp collectives as rendezvous.  What this covers and the
// per-ray emulation (wide_emulation.cpp) cannot: TracePersistent's own loop -- bulk re-fill with ballot
// + popc prefix, masked rays, the deferred RayHit stores, Resolve / instance-entry / phase votes as
// written in the kernel, the shared-memory stack with its global spill (SmemStack), TraceStatic's
// local stack.  The arithmetic is the host variant of traverse.h (same as the emulation), so results
// must equal the emulation's bit for bit.
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>       // the fake one (-I tests/cpp/fakecuda comes first)

thread_local FakeDim3 threadIdx, blockIdx;
FakeDim3 blockDim, gridDim;

namespace {
// one warp = 32 lane threads meeting at every collective
struct Warp {
	std::mutex m;
	std::condition_variable cv;
	int arrived = 0;
	unsigned long long generation = 0;
	uint32_t slot[32];
	uint32_t result[32];
	// every lane deposits `v`; the last one to arrive runs `combine` over the 32 deposits
	template <class F> uint32_t rendezvous(int lane, uint32_t v, F combine) {
		std::unique_lock<std::mutex> lk(m);
		slot[lane] = v;
		const unsigned long long gen = generation;
		if (++arrived == 32) {
			combine(slot, result);
			arrived = 0;
			++generation;
			cv.notify_all();
		} else
			cv.wait(lk, [&] { return generation != gen; });
		return result[lane];
	}
};
thread_local Warp *t_warp = nullptr;
thread_local int t_lane = 0;
}   // namespace

unsigned __ballot_sync(unsigned mask, int pred) {
	if (mask != 0xffffffffu) __builtin_trap();
	return t_warp->rendezvous(t_lane, pred ? 1u : 0u, [](const uint32_t *s, uint32_t *r) {
		uint32_t b = 0;
		for (int i = 0; i < 32; ++i) b |= (s[i] ? 1u : 0u) << i;
		for (int i = 0; i < 32; ++i) r[i] = b;
	});
}
// (the kernels only broadcast from a warp-uniform source lane)
static thread_local int t_srcLane = 0;
uint32_t __shfl_sync(unsigned mask, uint32_t v, int srcLane) {
	if (mask != 0xffffffffu) __builtin_trap();
	t_srcLane = srcLane;
	const int src = srcLane;
	return t_warp->rendezvous(t_lane, v, [src](const uint32_t *s, uint32_t *r) { for (int i = 0; i < 32; ++i) r[i] = s[src & 31]; });
}
uint32_t __reduce_min_sync(unsigned mask, uint32_t v) {
	if (mask != 0xffffffffu) __builtin_trap();
	return t_warp->rendezvous(t_lane, v, [](const uint32_t *s, uint32_t *r) {
		uint32_t m = s[0];
		for (int i = 1; i < 32; ++i) m = s[i] < m ? s[i] : m;
		for (int i = 0; i < 32; ++i) r[i] = m;
	});
}
void __nanosleep(unsigned ns) { std::this_thread::sleep_for(std::chrono::nanoseconds(ns < 200000u ? 200000u : ns)); }
void __syncwarp() {
	t_warp->rendezvous(t_lane, 0u, [](const uint32_t *, uint32_t *) { });
}

#include "trace_kernels.cuh"

namespace lrb {
__attribute__((aligned(16))) uint32_t smem[kTraceBlock * (2 * 65 + 9)];        // `extern __shared__ uint32_t smem[]` of the kernels: one block at a time
}

using namespace lrb;

namespace {
template <class KERNEL> void RunBlock(KERNEL kernel, const TraceArgs &a, int nWarps) {
	std::vector<Warp> warps(nWarps);
	std::vector<std::thread> threads;
	for (int t = 0; t < 32 * nWarps; ++t) {
		threads.emplace_back([&, t]() {
			threadIdx.x = (unsigned)t; threadIdx.y = threadIdx.z = 0;
			blockIdx.x = blockIdx.y = blockIdx.z = 0;
			t_warp = &warps[t / 32];
			t_lane = t % 32;
			kernel(a);
		});
	}
	for (auto &th : threads) th.join();
}
}   // namespace

extern "C" {

// Runs one block (nWarps <= 4 warps) of TracePersistent / TraceStatic over the batch.
//   view: SceneView of a re-laid-out scene (emu_scene_view of libwide_emulation.so)
//   kernel: 0 = TracePersistent, 1 = TraceStatic (stats6 receives its counters when non-null)
//   prefetch: bit 0 = the prefetching twin, bit 1 = the any-hit (shadow ray) kernels
int ks_trace_signal(const SceneView *view, const lrb_ray *rays, lrb_rayhit *hits, lrb_rayhit *hitsPeer, uint32_t n, int nWarps,
		uint32_t smemDepth, uint32_t stackNeed, uint32_t chunkShift, uint32_t epoch, uint32_t *chunkFlags, uint32_t *flagOrderOk);

int ks_trace(const SceneView *view, const lrb_ray *rays, lrb_rayhit *hits, uint32_t n, int kernel, int nWarps,
		uint32_t smemDepth, uint32_t stackNeed, uint32_t refillBelow, uint32_t triBias, uint32_t instBias, int prefetch,
		unsigned long long *stats6) {
	if (nWarps < 1 || nWarps > 4 || smemDepth < 1 || smemDepth > 64)
		return 1;
	gridDim.x = 1; gridDim.y = gridDim.z = 1;
	blockDim.x = kTraceBlock; blockDim.y = blockDim.z = 1;
	TraceArgs a;
	memset(&a, 0, sizeof(a));
	a.sc = *view;
	a.rays = rays;
	a.hits = hits;
	a.rayCount = n;
	uint32_t counter[2] = { 0, 0 };
	a.counter = counter;
	const uint32_t totalThreads = kTraceBlock;
	const uint32_t spillDepth = stackNeed + 4;
	std::vector<uint32_t> spillNode((size_t)spillDepth * totalThreads);
	std::vector<float> spillT((size_t)spillDepth * totalThreads);
	a.spillNode = spillNode.data();
	a.spillT = spillT.data();
	a.smemDepth = smemDepth;
	a.refillBelow = refillBelow;
	a.triBias = triBias;
	a.instBias = instBias;
	TraceStats st;
	memset(&st, 0, sizeof(st));
	a.stats = &st;
	const bool two = view->twoLevel != 0;
	const bool spill = stackNeed > smemDepth;
	const bool anyhit = (prefetch & 2) != 0;
	prefetch &= 1;
	if (anyhit) {
		if (kernel == 0) {
			if (two) {
				if (spill) RunBlock(TracePersistent<true, true, false, false, true>, a, nWarps);
				else RunBlock(TracePersistent<true, false, false, false, true>, a, nWarps);
			} else if (spill) RunBlock(TracePersistent<false, true, false, false, true>, a, nWarps);
			else RunBlock(TracePersistent<false, false, false, false, true>, a, nWarps);
		} else {
			if (nWarps != 4) return 2;
			if (two) RunBlock(TraceStatic<true, false, true>, a, 4);
			else RunBlock(TraceStatic<false, false, true>, a, 4);
		}
		return 0;
	}
	if (kernel == 0) {
		if (two) {
			if (spill) RunBlock(TracePersistent<true, true, false>, a, nWarps);
			else RunBlock(TracePersistent<true, false, false>, a, nWarps);
		} else if (spill) {
			if (prefetch) RunBlock(TracePersistent<false, true, false, true>, a, nWarps);
			else RunBlock(TracePersistent<false, true, false>, a, nWarps);
		} else
			RunBlock(TracePersistent<false, false, false>, a, nWarps);
	} else {
		// TraceStatic strides by gridDim.x * blockDim.x threads: every one of them must exist
		if (nWarps != 4) return 2;
		if (two) RunBlock(TraceStatic<true, true>, a, 4);
		else RunBlock(TraceStatic<false, true>, a, 4);
		if (stats6) {
			stats6[0] = st.rays; stats6[1] = st.wideNodes; stats6[2] = st.triangles;
			stats6[3] = st.instances; stats6[4] = st.motionSamples; stats6[5] = st.maxStack;
		}
	}
	return 0;
}

// SIGNAL kernels (the multi-GPU gather's trace): warp 0 is the detector, the other warps trace and publish
// watermarks; `hitsPeer` stands for the peer-mapped gather slice (dual stores).  A watcher thread plays the
// copy stream: whenever a chunk flag shows `epoch`, every RayHit of that chunk must already be final --
// it snapshots the chunk at that moment; the caller compares the snapshots with the final buffer.
int ks_trace_signal(const SceneView *view, const lrb_ray *rays, lrb_rayhit *hits, lrb_rayhit *hitsPeer, uint32_t n, int nWarps,
		uint32_t smemDepth, uint32_t stackNeed, uint32_t chunkShift, uint32_t epoch, uint32_t *chunkFlags, uint32_t *flagOrderOk) {
	if (nWarps != 4 || smemDepth < 1 || smemDepth > 64)
		return 1;       // the detector waits for a watermark from every warp of the block: all four must run
	gridDim.x = 1; gridDim.y = gridDim.z = 1;
	blockDim.x = kTraceBlock; blockDim.y = blockDim.z = 1;
	TraceArgs a;
	memset(&a, 0, sizeof(a));
	a.sc = *view;
	a.rays = rays;
	a.hits = hits;
	a.hitsPeer = hitsPeer;
	a.rayCount = n;
	uint32_t counter[2] = { 0, 0 };
	a.counter = counter;
	const uint32_t spillDepth = stackNeed + 4;
	std::vector<uint32_t> spillNode((size_t)spillDepth * kTraceBlock);
	std::vector<float> spillT((size_t)spillDepth * kTraceBlock);
	a.spillNode = spillNode.data();
	a.spillT = spillT.data();
	a.smemDepth = smemDepth;
	a.refillBelow = 24;
	a.triBias = 8;
	a.instBias = 8;
	std::vector<uint32_t> watermark(kTraceBlock / 32, 0u);
	a.watermark = watermark.data();
	a.chunkFlag = chunkFlags;
	a.chunkShift = chunkShift;
	a.epoch = epoch;
	const uint32_t chunkRays = 1u << chunkShift, nChunks = (n + chunkRays - 1) >> chunkShift;
	// the copy stream's stand-in
	std::vector<lrb_rayhit> snap(n);
	std::vector<char> taken(nChunks, 0);
	volatile bool stop = false;
	std::thread watcher([&]() {
		uint32_t left = nChunks;
		while (left && !stop) {
			for (uint32_t c = 0; c < nChunks; ++c) {
				if (taken[c]) continue;
				if (__atomic_load_n(chunkFlags + c, __ATOMIC_ACQUIRE) == epoch) {
					const uint32_t b = c << chunkShift, e = std::min(n, b + chunkRays);
					memcpy(&snap[b], (const void *)(hits + b), (size_t)(e - b) * sizeof(lrb_rayhit));
					taken[c] = 1;
					--left;
				}
			}
			std::this_thread::sleep_for(std::chrono::microseconds(200));
		}
	});
	const bool two = view->twoLevel != 0;
	const bool spill = stackNeed > smemDepth;
	if (two) {
		if (spill) RunBlock(TracePersistent<true, true, true>, a, nWarps);
		else RunBlock(TracePersistent<true, false, true>, a, nWarps);
	} else if (spill) RunBlock(TracePersistent<false, true, true>, a, nWarps);
	else RunBlock(TracePersistent<false, false, true>, a, nWarps);
	stop = true;
	watcher.join();
	// flags raised while the kernel ran: their snapshots must equal the final records
	uint32_t ok = 1, raised = 0;
	for (uint32_t c = 0; c < nChunks; ++c) {
		if (!taken[c]) continue;
		++raised;
		const uint32_t b = c << chunkShift, e = std::min(n, b + chunkRays);
		if (memcmp(&snap[b], hits + b, (size_t)(e - b) * sizeof(lrb_rayhit)) != 0) ok = 0;
	}
	flagOrderOk[0] = ok;
	flagOrderOk[1] = raised;
	return 0;
}

}
