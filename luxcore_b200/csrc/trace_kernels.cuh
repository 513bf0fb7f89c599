// trace_kernels.cuh -- sm_100a closest-hit kernels (replace bvh.cl / mbvh.cl
// Accelerator_Intersect_RayBuffer, include/luxrays/accelerators/bvh.cl:228-260, mbvh.cl:351-383).
//
// TracePersistent: persistent warps, one ray per lane.  A warp pulls ray indices from a global
// atomic counter; lanes whose ray finished are re-filled in bulk (one atomicAdd per warp, the
// lanes' slots assigned with a ballot + popc prefix) as soon as the number of live lanes drops
// under a threshold, so incoherent bounce rays do not leave a warp running with a handful of
// lanes.  Finished lanes keep their result in registers until that re-fill, where all of them
// store their RayHit together.  The traversal stack is a per-lane column in shared memory
// (bank = lane, conflict-free at any mix of depths); entries beyond the shared depth spill to a
// global scratch column.  Nodes/triangles are fetched with 256-bit read-only loads
// (ld.global.nc.v8.f32).
//
// TraceStatic: same per-ray code, static grid-stride assignment, stack in local memory.  Kept as
// the "simple" variant for A/B measurements and as the instrumented (STATS) kernel.
#ifndef LRB_TRACE_KERNELS_CUH
#define LRB_TRACE_KERNELS_CUH

#include <cuda_runtime.h>

#include "traverse.h"

namespace lrb {

static const int kTraceBlock = 128;     // threads per block of every trace kernel
// "auto" compaction threshold (measured crossover on kitchen bounce-2, profiles/r02_masked_batches.json: the list pays
// for its three passes once about half of the lanes are dead)
static const uint32_t kCompactAutoNum = 6, kCompactAutoDen = 10;
// Resident blocks per SM the persistent kernels are compiled for (register budget = 65 536 / (128 x blocks)).
// Measured on B200 (profiles/r02_*): one-level 8 -> 10 blocks (64 -> 48 registers, 88 bytes of spill stores in cold
// paths) 3 108 -> 3 197 Mrays/s on kitchen bounce-2; two-level 6 -> 8 blocks (80 -> 64 registers) 1 555 -> 1 752
// Mrays/s on lightinstances bounce-1.
#ifndef LRB_MINBLOCKS_1L
#define LRB_MINBLOCKS_1L 10
#endif
#ifndef LRB_MINBLOCKS_2L
#define LRB_MINBLOCKS_2L 8
#endif
#ifndef LRB_BOTH_PHASES
#define LRB_BOTH_PHASES 0
#endif

// ---- stacks ---------------------------------------------------------------------------------

// Shared-memory column per thread + global spill (compiled out when the scene's worst-case stack
// fits the shared depth, which the host knows at launch).  Entry i of thread t is the 8-byte word pair
// [2 * (i * kTraceBlock + t)] = (reference, entry distance): one 64-bit store per push, one 64-bit load
// per pop; a warp's 32 pairs are 256 contiguous bytes (two conflict-free wavefronts).
template <bool SPILL> struct SmemStack {
	uint32_t *base;         // &smem[2 * threadIdx.x]
	uint32_t *top;          // next free shared entry (== base + sp * 2 * kTraceBlock while sp <= depthSmem)
	uint32_t *limit;        // base + depthSmem * 2 * kTraceBlock
	uint32_t *gNode;        // spill columns [spillDepth][totalThreads] (the kernel's arguments as they are: uniform)
	float *gT;
	uint32_t gStride;       // totalThreads
	int over;               // SPILL: entries currently held in the global column

	// Entry 0 of every column is the permanent bottom entry (kStackBottom, -inf): see Resolve.
	__device__ __forceinline__ void init(uint32_t *smemBlock, int depthSmem) {
		base = smemBlock + 2 * threadIdx.x;
		st2(base, kStackBottom, -LRB_INF);
		base += 2 * kTraceBlock;
		top = base;
		limit = base + depthSmem * (2 * kTraceBlock);
		over = 0;
	}
	__device__ __forceinline__ void reset() { top = base; over = 0; }
	__device__ __forceinline__ void keepBottom() { top += 2 * kTraceBlock; }
	__device__ __forceinline__ bool room(int n) const { return !SPILL || top + n * (2 * kTraceBlock) <= limit; }
	__device__ __forceinline__ static void st2(uint32_t *p, uint32_t n, float t) {
		*reinterpret_cast<uint2 *>(p) = make_uint2(n, __float_as_uint(t));
	}
	__device__ __forceinline__ static void ld2(const uint32_t *p, uint32_t &n, float &t) {
		const uint2 v = *reinterpret_cast<const uint2 *>(p);
		n = v.x;
		t = __uint_as_float(v.y);
	}
	__device__ __forceinline__ void pushFast(uint32_t n, float t) {
		st2(top, n, t);
		top += 2 * kTraceBlock;
	}
	__device__ __forceinline__ void pushFastIf(bool p, uint32_t n, float t) {
		if (p)
			st2(top, n, t);
		top += p ? 2 * kTraceBlock : 0;
	}
	__device__ __forceinline__ void push(uint32_t n, float t) {
		if (!SPILL || top < limit) {
			pushFast(n, t);
			return;
		}
		const size_t o = spillSlot(over);
		gNode[o] = n;
		gT[o] = t;
		++over;
	}
	// Slot of this thread's spilled entry `i`.  The address arithmetic must stay inside the (rare) branch that uses
	// it: left alone, the compiler computes it speculatively in front of Resolve's loops, where every lane that pops
	// pays ~20 instructions for it (round-2 SASS).  Nothing that depends on a volatile asm can move above it.
	__device__ __forceinline__ size_t spillSlot(int i) const {
		uint32_t z = 0;
#if defined(__CUDA_ARCH__)
		asm volatile("" : "+r"(z));
#endif
		return (size_t)i * gStride + (blockIdx.x * kTraceBlock + threadIdx.x + z);
	}
	__device__ __forceinline__ void pop(uint32_t &n, float &t) {
		if (SPILL && over > 0) {
			--over;
			const size_t o = spillSlot(over);
			n = gNode[o];
			t = gT[o];
			return;
		}
		top -= 2 * kTraceBlock;
		ld2(top, n, t);
	}
	__device__ __forceinline__ bool slow() const { return SPILL && over > 0; }
	__device__ __forceinline__ void popFast(uint32_t &n, float &t) {
		top -= 2 * kTraceBlock;
		ld2(top, n, t);
	}
	// Speculative pop at the end of a phase (PopSpec): reads the top entry when `want` and the top lives in shared
	// memory; nothing is removed until dropIf.
	__device__ __forceinline__ bool peekIf(bool want, uint32_t &n, float &t) const {
		const bool can = want && (!SPILL || over == 0);     // (an empty column shows its bottom entry)
		if (can)
			ld2(top - 2 * kTraceBlock, n, t);
		return can;
	}
	__device__ __forceinline__ void dropIf(bool p) { top -= p ? 2 * kTraceBlock : 0; }
	__device__ __forceinline__ bool empty() const { return top == base; }
	__device__ __forceinline__ unsigned long long depth() const { return (unsigned long long)((top - base) / (2 * kTraceBlock) + over); }
	// two-level kernels: nine more words per thread behind the stack columns (the launch sizes shared memory for
	// them), word k of thread t at [(1 + depthSmem) * 2 * kTraceBlock + k * kTraceBlock + t]
	__device__ __forceinline__ uint32_t *stash() const { return limit - threadIdx.x; }
	__device__ __forceinline__ void stashRay(float ox, float oy, float oz, float dx, float dy, float dz, float ix, float iy, float iz) {
		uint32_t *w = stash();
		w[0] = __float_as_uint(ox); w[kTraceBlock] = __float_as_uint(oy); w[2 * kTraceBlock] = __float_as_uint(oz);
		w[3 * kTraceBlock] = __float_as_uint(dx); w[4 * kTraceBlock] = __float_as_uint(dy); w[5 * kTraceBlock] = __float_as_uint(dz);
		w[6 * kTraceBlock] = __float_as_uint(ix); w[7 * kTraceBlock] = __float_as_uint(iy); w[8 * kTraceBlock] = __float_as_uint(iz);
	}
	__device__ __forceinline__ void loadRay(float &ox, float &oy, float &oz, float &dx, float &dy, float &dz, float &ix, float &iy, float &iz) const {
		const uint32_t *w = stash();
		ox = __uint_as_float(w[0]); oy = __uint_as_float(w[kTraceBlock]); oz = __uint_as_float(w[2 * kTraceBlock]);
		dx = __uint_as_float(w[3 * kTraceBlock]); dy = __uint_as_float(w[4 * kTraceBlock]); dz = __uint_as_float(w[5 * kTraceBlock]);
		ix = __uint_as_float(w[6 * kTraceBlock]); iy = __uint_as_float(w[7 * kTraceBlock]); iz = __uint_as_float(w[8 * kTraceBlock]);
	}
};

// Local-memory stack + global spill.
template <int CAP> struct LocalStack {
	uint32_t node[CAP];
	float t0[CAP];
	uint32_t *gNode;
	float *gT;
	uint32_t gStride;
	int sp;

	__device__ __forceinline__ bool room(int) const { return false; }
	__device__ __forceinline__ void pushFast(uint32_t n, float t) { push(n, t); }
	__device__ __forceinline__ void pushFastIf(bool p, uint32_t n, float t) { if (p) push(n, t); }
	__device__ __forceinline__ void push(uint32_t n, float t) {
		if (sp < CAP) {
			node[sp] = n;
			t0[sp] = t;
		} else {
			const size_t o = (size_t)(sp - CAP) * gStride;
			gNode[o] = n;
			gT[o] = t;
		}
		++sp;
	}
	__device__ __forceinline__ void pop(uint32_t &n, float &t) {
		--sp;
		if (sp < CAP) {
			n = node[sp];
			t = t0[sp];
		} else {
			const size_t o = (size_t)(sp - CAP) * gStride;
			n = gNode[o];
			t = gT[o];
		}
	}
	__device__ __forceinline__ bool slow() const { return false; }
	__device__ __forceinline__ void popFast(uint32_t &n, float &t) {
		if (sp == 0) {
			n = kStackBottom;
			t = -LRB_INF;
		} else
			pop(n, t);
	}
	__device__ __forceinline__ void keepBottom() { }
	__device__ __forceinline__ bool peekIf(bool want, uint32_t &n, float &t) const {
		const bool can = want && sp > 0 && sp <= CAP;
		if (can) {
			n = node[sp - 1];
			t = t0[sp - 1];
		}
		return can;
	}
	__device__ __forceinline__ void dropIf(bool p) { sp -= p ? 1 : 0; }
	__device__ __forceinline__ bool empty() const { return sp == 0; }
	__device__ __forceinline__ unsigned long long depth() const { return (unsigned long long)sp; }
	float w[9];
	__device__ __forceinline__ void stashRay(float ox, float oy, float oz, float dx, float dy, float dz, float ix, float iy, float iz) {
		w[0] = ox; w[1] = oy; w[2] = oz; w[3] = dx; w[4] = dy; w[5] = dz; w[6] = ix; w[7] = iy; w[8] = iz;
	}
	__device__ __forceinline__ void loadRay(float &ox, float &oy, float &oz, float &dx, float &dy, float &dz, float &ix, float &iy, float &iz) const {
		ox = w[0]; oy = w[1]; oz = w[2]; dx = w[3]; dy = w[4]; dz = w[5]; ix = w[6]; iy = w[7]; iz = w[8];
	}
};

struct TraceArgs {
	SceneView sc;
	const lrb_ray *rays;
	lrb_rayhit *hits;
	uint32_t rayCount;
	uint32_t *counter;          // persistent kernel: next unassigned ray (zeroed before launch)
	const uint32_t *perm;       // optional processing order (ray indices sorted for coherence, or the dense list of
	                            // live rays left by the compaction kernels), or NULL
	const uint32_t *rayCountDev;    // optional: number of entries of perm to process, read on the device (compaction
	                            // leaves it there; no host round trip between the compaction and the trace)
	uint32_t permAuto;          // compaction in "auto" mode: the list is only used (and was only written) when fewer than
	                            // kCompactAutoNum / kCompactAutoDen of the rays are live; otherwise masked rays are skipped here
	uint32_t *spillNode;        // global stack spill, [spillDepth][totalThreads]
	float *spillT;
	uint32_t smemDepth;         // stack entries held in shared memory per thread
	uint32_t refillBelow;       // re-fill when fewer live lanes than this
	uint32_t triBias;           // triangle phase runs when nTri * triBias >= nNode * 4 (4 = plain majority)
	uint32_t instBias;          // two-level: instances are entered when nInst * instBias >= max(nNode * 4, nTri * triBias); 0 = at once
	uint32_t prefetchMode;      // PREFETCH kernels: see NodeStep
	TraceStats *stats;          // STATS kernels only
	// Multi-GPU gather fused into the trace: when set, every RayHit record is ALSO stored here -- this
	// rank's slice of the gather buffer on the destination GPU, peer-mapped over NVLink.  The five
	// 4-byte stores per ray are posted writes that ride along with the traversal (measured: no change
	// in kernel time at 57 GB/s of RayHit payload), so no copy pass / collective follows the kernel.
	lrb_rayhit *hitsPeer;
	uint32_t hitFlags;          // bit 0: hits is 16-byte aligned, bit 1: hitsPeer is 16-byte aligned
	// Chunk-completion signalling (SIGNAL kernels).  Ray indices are handed out in increasing order,
	// so "everything below index p is finished" is a prefix property: every warp publishes a
	// watermark -- the smallest ray index it still holds (or, holding none, the end of its last fetch,
	// below which it will never fetch again) -- and one detector warp keeps the minimum over all
	// watermarks.  When that prefix passes the end of a chunk of (1 << chunkShift) ray indices the
	// detector writes `epoch` into chunkFlag[chunk]; a copy stream waits on the flag
	// (cuStreamWaitValue32) and pushes that chunk of the RayHit buffer with the copy engine, in
	// full-size NVLink packets, while the kernel keeps tracing.  No per-ray atomics.
	uint32_t *watermark;        // one word per warp of the grid, zeroed before launch
	uint32_t *chunkFlag;
	uint32_t chunkShift;
	uint32_t epoch;
};

__device__ __forceinline__ void LoadRay(const lrb_ray *rays, uint32_t i, lrb_ray &r) {
	const float4 *p = reinterpret_cast<const float4 *>(rays + i);    // 48-B records, 16-B aligned
	const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
	r.o[0] = a.x; r.o[1] = a.y; r.o[2] = a.z;
	r.d[0] = a.w; r.d[1] = b.x; r.d[2] = b.y;
	r.mint = b.z; r.maxt = b.w;
	r.time = c.x; r.flags = __float_as_uint(c.y);
}

// 20-byte records are only 4-byte aligned, but with a 16-byte aligned buffer the alignment of record
// i is known from i & 3: two or three vector stores cover it instead of five scalar ones (fewer
// store instructions, and fewer, larger write packets when the buffer is peer memory over NVLink).
__device__ __forceinline__ void StoreHitTo(lrb_rayhit *hits, uint32_t i, const lrb_rayhit &h, const bool aligned16) {
	char *p = reinterpret_cast<char *>(hits + i);
	const uint32_t w0 = __float_as_uint(h.t), w1 = __float_as_uint(h.b1), w2 = __float_as_uint(h.b2), w3 = h.meshIndex, w4 = h.triangleIndex;
	if (!aligned16) {
		uint32_t *q = reinterpret_cast<uint32_t *>(p);
		q[0] = w0; q[1] = w1; q[2] = w2; q[3] = w3; q[4] = w4;
		return;
	}
	switch (i & 3u) {
	case 0:     // offset 0 mod 16
		*reinterpret_cast<uint4 *>(p) = make_uint4(w0, w1, w2, w3);
		*reinterpret_cast<uint32_t *>(p + 16) = w4;
		break;
	case 1:     // offset 4 mod 16
		*reinterpret_cast<uint32_t *>(p) = w0;
		*reinterpret_cast<uint2 *>(p + 4) = make_uint2(w1, w2);
		*reinterpret_cast<uint2 *>(p + 12) = make_uint2(w3, w4);
		break;
	case 2:     // offset 8 mod 16
		*reinterpret_cast<uint2 *>(p) = make_uint2(w0, w1);
		*reinterpret_cast<uint2 *>(p + 8) = make_uint2(w2, w3);
		*reinterpret_cast<uint32_t *>(p + 16) = w4;
		break;
	default:    // offset 12 mod 16
		*reinterpret_cast<uint32_t *>(p) = w0;
		*reinterpret_cast<uint4 *>(p + 4) = make_uint4(w1, w2, w3, w4);
		break;
	}
}

__device__ __forceinline__ void StoreHit(const TraceArgs &a, uint32_t i, const RayState &s, float rayMaxt) {
	lrb_rayhit h;
	WriteHit(a.sc, s, rayMaxt, &h);
	if (a.hits)
		StoreHitTo(a.hits, i, h, (a.hitFlags & 1u) != 0);
	if (a.hitsPeer)
		StoreHitTo(a.hitsPeer, i, h, (a.hitFlags & 2u) != 0);
}

// A masked ray's RayHit is left untouched (bvh.cl:242-244); the gather slice must still end up equal
// to the local buffer, so the record is forwarded as it is.
__device__ __forceinline__ void ForwardMaskedHit(const TraceArgs &a, uint32_t i) {
	if (a.hits && a.hitsPeer) {
		const uint32_t *src = reinterpret_cast<const uint32_t *>(a.hits + i);
		uint32_t *dst = reinterpret_cast<uint32_t *>(a.hitsPeer + i);
#pragma unroll
		for (int k = 0; k < 5; ++k)
			dst[k] = src[k];
	}
}

// ---- persistent, warp-cooperative kernel ----------------------------------------------------

// Watermark of one warp (SIGNAL kernels).  __syncwarp orders every lane's RayHit stores before the
// leader's fence; the fence makes them visible device-wide before the watermark moves.
__device__ __forceinline__ void PublishWatermark(const TraceArgs &a, const uint32_t lane, const uint32_t warpId,
		const bool holdsRay, const uint32_t rayIdx, const uint32_t lowBound, const bool exhausted) {
	uint32_t v = __reduce_min_sync(0xffffffffu, holdsRay ? rayIdx : 0xffffffffu);
	if (v == 0xffffffffu && !exhausted)
		v = lowBound;
	__syncwarp();
	if (lane == 0) {
		__threadfence();
#if defined(__CUDA_ARCH__)
		asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(a.watermark + warpId), "r"(v) : "memory");
#else
		*(volatile uint32_t *)(a.watermark + warpId) = v;       // host pass / host harness of the tests
#endif
	}
}

// Detector warp: minimum over all watermarks -> finished prefix -> chunk flags.
__device__ __forceinline__ void DetectorLoop(const TraceArgs &a, const uint32_t lane, const uint32_t nWarps) {
	const uint32_t chunkRays = 1u << a.chunkShift;
	const uint32_t nChunks = (a.rayCount + chunkRays - 1) >> a.chunkShift;
	uint32_t next = 0;
	// bounded (~10 s): the host raises every flag after the kernel anyway
#pragma unroll 1
	for (uint32_t spin = 0; spin < (1u << 20) && next < nChunks; ++spin) {
		uint32_t m = 0xffffffffu;
		for (uint32_t w = 1 + lane; w < nWarps; w += 32) {      // warp 0 is the detector itself
			uint32_t v;
#if defined(__CUDA_ARCH__)
			asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.watermark + w) : "memory");
#else
			v = *(volatile const uint32_t *)(a.watermark + w);
#endif
			m = min(m, v);
		}
		m = __reduce_min_sync(0xffffffffu, m);
		bool raised = false;
		while (next < nChunks && m >= min((next + 1) << a.chunkShift, a.rayCount)) {
			if (lane == 0) {
				__threadfence();
				atomicExch(a.chunkFlag + next, a.epoch);
			}
			++next;
			raised = true;
		}
		if (!raised)
			__nanosleep(8192);
	}
}

// ANYHIT (shadow / visibility rays, lrb_trace_anyhit): the first accepted hit finishes the ray -- its stack is
// dropped and the lane stores that hit at the next re-fill.  Hit / miss is the closest-hit kernel's (and the
// reference's Intersect's): a ray is a hit exactly when at least one triangle passes the test and its gate.
template <bool TWO_LEVEL, bool SPILL, bool SIGNAL, bool PREFETCH = false, bool ANYHIT = false>
__global__ void __launch_bounds__(kTraceBlock, TWO_LEVEL ? LRB_MINBLOCKS_2L : LRB_MINBLOCKS_1L) TracePersistent(const TraceArgs a) {
	extern __shared__ __align__(16) uint32_t smem[];
	uint32_t rayCount = a.rayCount;
	const uint32_t *perm = a.perm;
	if (a.rayCountDev) {
		const uint32_t live = __ldg(a.rayCountDev);
		if (a.permAuto && (unsigned long long)live * kCompactAutoDen > (unsigned long long)a.rayCount * kCompactAutoNum)
			perm = nullptr;         // mostly live: index order, masked rays skipped below
		else
			rayCount = live;
	}
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warpId = (blockIdx.x * kTraceBlock + threadIdx.x) >> 5;
	if (SIGNAL && warpId == 0) {
		DetectorLoop(a, lane, (gridDim.x * kTraceBlock) >> 5);
		return;
	}
	const uint32_t totalThreads = gridDim.x * kTraceBlock;
	const uint32_t gtid = blockIdx.x * kTraceBlock + threadIdx.x;

	SmemStack<SPILL> stk;
	stk.init(smem, (int)a.smemDepth);
	stk.gNode = a.spillNode;
	stk.gT = a.spillT;
	stk.gStride = totalThreads;

	RayState s;
	uint32_t rayIdx = 0;
	float rayMaxt = 0.f;
	// lane state: kIdle (nothing), kActive (traversing), kUnsaved (finished, RayHit not stored yet)
	enum { kIdle = 0, kActive = 1, kUnsaved = 2 };
	int state = kIdle;
	int exhausted = 0;
	uint32_t lowBound = 0;          // SIGNAL: this warp never fetches a ray index below this again
	uint32_t round = 0;

	for (;;) {
		// ---- finished lanes store their RayHit (all of them in the same instructions) ----
		if (state == kUnsaved) {
			StoreHit(a, rayIdx, s, rayMaxt);
			state = kIdle;
		}
		if (SIGNAL) {
			if (exhausted || (++round & 15u) == 0)
				PublishWatermark(a, lane, warpId, state == kActive, rayIdx, lowBound, exhausted != 0);
		}
		// ---- re-fill idle lanes ----
		const unsigned idle = __ballot_sync(0xffffffffu, state == kIdle);
		if (!exhausted && idle) {
			const int nIdle = __popc(idle);
			const int leader = __ffs(idle) - 1;
			uint32_t base = 0;
			if ((int)lane == leader)
				base = atomicAdd(a.counter, (uint32_t)nIdle);
			base = __shfl_sync(0xffffffffu, base, leader);
			if (state == kIdle) {
				const uint32_t slot = base + __popc(idle & ((1u << lane) - 1u));
				if (slot < rayCount) {
					const uint32_t idx = perm ? __ldg(perm + slot) : slot;
					lrb_ray r;
					LoadRay(a.rays, idx, r);
					// masked rays are skipped and their RayHit is left untouched (bvh.cl:242-244)
					if (!(r.flags & LRB_RAY_FLAGS_MASKED)) {
						rayIdx = idx;
						rayMaxt = r.maxt;
						if (InitRay(a.sc, r, s)) {
							stk.reset();
							state = kActive;
						} else
							StoreHit(a, idx, s, rayMaxt);   // empty scene: miss
					} else
						ForwardMaskedHit(a, idx);
				}
			}
			if (SIGNAL)
				lowBound = base + (uint32_t)nIdle;
			if (base + (uint32_t)nIdle >= rayCount)
				exhausted = 1;
		}
		if (__ballot_sync(0xffffffffu, state == kActive) == 0) {
			if (exhausted) {
				if (SIGNAL)
					PublishWatermark(a, lane, warpId, false, 0u, 0u, true);     // nothing left: +inf
				break;
			}
			continue;
		}

		// ---- traverse until too few lanes are alive ----
		// Every iteration first lets the lanes that ran out of work pop their stack until they hold an
		// entry that is still worth visiting (Resolve; an empty stack finishes the ray), then the warp
		// runs ONE of two branch-free phases, whichever has more lanes ready: a node phase (fetch a 64-B
		// node / four box tests / ordered push) or a triangle phase (one triangle per lane).  A lane
		// holding a triangle reference waits for a triangle phase; batching the two kinds of work keeps
		// lanes converged on incoherent rays instead of serialising them against each other.
		// (Round-2 ncu source view: Resolve's pop loop ran 2.5 trips per iteration at 3.8 of 32 lanes -- 22 % of the
		// issue slots; it is now 5 instructions per culled entry, see Resolve.  Measured alternatives, all slower on
		// incoherent rays: one converged pop attempt per iteration INSTEAD of the loop; pops only inside the phases --
		// lanes whose popped entry is culled then idle a phase; and LRB_POPSPEC speculative attempts at the end of
		// the phases in addition to the loop -- PopSpec, compiled out by default.)
		const int floorLanes = exhausted ? 1 : (int)a.refillBelow;
		int nLive;
		do {
			if (state == kActive && NeedsResolve<TWO_LEVEL>(s.cur)) {
				if (!Resolve<TWO_LEVEL, false>(a.sc, s, stk, nullptr))
					state = kUnsaved;
			}
			__syncwarp();
			LaneWork work = state == kActive ? WorkOf<TWO_LEVEL>(s.cur) : kWorkNone;
			int nTri = __popc(__ballot_sync(0xffffffffu, work == kWorkTri));
			int nNode = __popc(__ballot_sync(0xffffffffu, work == kWorkNode));
			int nInst = 0;
			if (TWO_LEVEL) {
				// Entering an instance is a third kind of work, and Resolve leaves it to this place (see
				// EnterInstance): lanes that reached an instance reference wait until they outweigh the lanes
				// with node / triangle work, then enter together; the entered lanes join this iteration's vote.
				nInst = __popc(__ballot_sync(0xffffffffu, work == kWorkInstance));
				if (VoteEnterInstances(nInst, nNode, nTri, a.instBias, a.triBias)) {
					if (work == kWorkInstance) {
						EnterInstance<false>(a.sc, s, stk, nullptr);
						work = WorkOf<TWO_LEVEL>(s.cur);
					}
					__syncwarp();
					nNode = __popc(__ballot_sync(0xffffffffu, work == kWorkNode));
					nInst = 0;
				}
			}
#if LRB_BOTH_PHASES
			// Experimental schedule (-DLRB_BOTH_PHASES=<triMin>): every iteration runs the node phase for the lanes that
			// hold a node and THEN the triangle phase for the lanes that hold a triangle -- those that waited and those
			// whose nearest child just turned out to be one -- when at least triMin of them are ready or no node phase
			// ran.  One Resolve / vote / loop overhead per pair of phases instead of per phase.
			if (nNode) {
				if (work == kWorkNode) {
					NodeStep<TWO_LEVEL, false, PREFETCH>(a.sc, s, stk, nullptr);
					work = (s.cur < kTagInstance && (s.cur & kTagTri)) ? kWorkTri : kWorkNone;
				}
				__syncwarp();
				nTri = __popc(__ballot_sync(0xffffffffu, work == kWorkTri));
			}
			if (nTri && (nTri >= LRB_BOTH_PHASES || !nNode)) {
				if (work == kWorkTri) {
					TriStep<TWO_LEVEL, false>(a.sc, s, nullptr);
					if (ANYHIT && s.hitRef != kNullIndex) {
						stk.reset();
						s.inInstance = false;
					}
				}
			}
			nLive = __popc(__ballot_sync(0xffffffffu, state == kActive));
			continue;
#endif
			if (VoteTrianglePhase(nTri, nNode, a.triBias)) {
				if (work == kWorkTri) {
					TriStep<TWO_LEVEL, false>(a.sc, s, nullptr);
					if (ANYHIT && s.hitRef != kNullIndex) {
						stk.reset();        // s.cur is already kNullIndex: the next Resolve finds the ray finished
						s.inInstance = false;
					}
					PopSpec<TWO_LEVEL>(s, stk);
				}
			} else {
				if (work == kWorkNode) {
					NodeStep<TWO_LEVEL, false, PREFETCH>(a.sc, s, stk, nullptr, a.prefetchMode);
					PopSpec<TWO_LEVEL>(s, stk);
				}
			}
			nLive = nTri + nNode + nInst;
		} while (nLive >= floorLanes);
	}
}

// ---- ray ordering ----------------------------------------------------------------------------

// Sort key of a ray for the optional coherence pre-pass: direction octant in the top bits, Morton
// code of the origin cell (bitsPerAxis bits per axis inside the scene's root box) below.  Rays with
// the same key start in the same cell and order the children of every node the same way.  Masked
// rays get the largest key.
__device__ __forceinline__ uint32_t SpreadBits3(uint32_t v) {      // 10 bits -> every third bit
	v = (v | (v << 16)) & 0x030000ffu;
	v = (v | (v << 8)) & 0x0300f00fu;
	v = (v | (v << 4)) & 0x030c30c3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

__global__ void __launch_bounds__(256) RayKeyKernel(const lrb_ray *__restrict__ rays, const uint32_t n,
		const float lox, const float loy, const float loz, const float sx, const float sy, const float sz,
		const uint32_t bitsPerAxis, uint32_t *__restrict__ keys, uint32_t *__restrict__ idx) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const float4 *p = reinterpret_cast<const float4 *>(rays + i);
	const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
	const float cells = (float)((1u << bitsPerAxis) - 1u);
	const uint32_t qx = (uint32_t)fminf(fmaxf((a.x - lox) * sx, 0.f), cells);
	const uint32_t qy = (uint32_t)fminf(fmaxf((a.y - loy) * sy, 0.f), cells);
	const uint32_t qz = (uint32_t)fminf(fmaxf((a.z - loz) * sz, 0.f), cells);
	const uint32_t morton = SpreadBits3(qx) | (SpreadBits3(qy) << 1) | (SpreadBits3(qz) << 2);
	const uint32_t octant = (a.w < 0.f ? 1u : 0u) | (b.x < 0.f ? 2u : 0u) | (b.y < 0.f ? 4u : 0u);
	uint32_t key = (octant << (3u * bitsPerAxis)) | morton;
	if (__float_as_uint(c.y) & LRB_RAY_FLAGS_MASKED)
		key = (8u << (3u * bitsPerAxis)) - 1u;
	keys[i] = key;
	idx[i] = i;
}

// ---- static grid-stride kernel (simple variant + instrumented variant) -----------------------

template <bool TWO_LEVEL, bool STATS, bool ANYHIT = false>
__global__ void __launch_bounds__(kTraceBlock) TraceStatic(const TraceArgs a) {
	const uint32_t totalThreads = gridDim.x * blockDim.x;
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
	LocalStack<32> stk;
	stk.gNode = a.spillNode ? a.spillNode + gtid : nullptr;
	stk.gT = a.spillT ? a.spillT + gtid : nullptr;
	stk.gStride = totalThreads;
	TraceStats local;
	local.rays = local.wideNodes = local.triangles = local.instances = local.motionSamples = local.maxStack = 0;
	unsigned long long nRays = 0;

	uint32_t rayCount = a.rayCount;
	const uint32_t *perm = a.perm;
	if (a.rayCountDev) {
		const uint32_t live = __ldg(a.rayCountDev);
		if (a.permAuto && (unsigned long long)live * kCompactAutoDen > (unsigned long long)a.rayCount * kCompactAutoNum)
			perm = nullptr;
		else
			rayCount = live;
	}
	for (uint32_t slot = gtid; slot < rayCount; slot += totalThreads) {
		const uint32_t i = perm ? __ldg(perm + slot) : slot;
		lrb_ray r;
		LoadRay(a.rays, i, r);
		if (r.flags & LRB_RAY_FLAGS_MASKED) {
			ForwardMaskedHit(a, i);
			continue;
		}
		++nRays;
		RayState s;
		stk.sp = 0;
		if (InitRay(a.sc, r, s)) {
			while (Step<TWO_LEVEL, STATS>(a.sc, s, stk, &local)) {
				if (ANYHIT && s.hitRef != kNullIndex)
					break;
			}
		}
		StoreHit(a, i, s, r.maxt);
	}
	if (STATS) {
		atomicAdd(&a.stats->wideNodes, local.wideNodes);
		atomicAdd(&a.stats->triangles, local.triangles);
		atomicAdd(&a.stats->instances, local.instances);
		atomicAdd(&a.stats->motionSamples, local.motionSamples);
		atomicMax(&a.stats->maxStack, local.maxStack);
		atomicAdd(&a.stats->rays, nRays);
	}
}

}   // namespace lrb

#endif
