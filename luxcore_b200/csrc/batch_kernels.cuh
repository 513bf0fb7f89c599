// batch_kernels.cuh -- sm_100a kernels that run BETWEEN two traces of a batch: the pass-through re-trace
// step of the reference's Scene::Intersect loop and the compaction of dead (masked) lanes.  Included by
// device.cu only (the trace kernels proper live in trace_kernels.cuh).
#ifndef LRB_BATCH_KERNELS_CUH
#define LRB_BATCH_KERNELS_CUH

#include <cuda_runtime.h>

#include "traverse.h"

namespace lrb {

// ---- between two traces: pass-through re-trace, dead-lane compaction --------------------------

// MachineEpsilon::E(float) (include/luxrays/core/epsilon.h:48-53,75-82, constants epsilon_types.cl:21-30):
// |NextFloat(v) - v| with NextFloat = bits + 0x80, clamped to [1e-5, 1e-1].
LRB_HD float MachineEpsilonE(const float v) {
	const float next = LRB_U2F(LRB_F2U(v) + 0x80u);
	const float e = fabsf(LRB_SUB(next, v));
	return e < 1e-5f ? 1e-5f : (e > 1e-1f ? 1e-1f : e);     // Clamp (utils.h:142-150): NaN passes through
}

// One round of the reference's pass-through loop (Scene::Intersect, src/slg/scene/scene.cpp:556-690; GPU twin
// include/slg/scene/scene_funcs.cl:21-150), for a whole batch, between two traces:
//   * a ray whose hit is "continue to trace" -- its mesh has its bit set in passMesh (camera-invisible object,
//     fully transparent material: scene.cpp:646-668) or the caller flagged the ray in continueFlags -- is
//     re-armed behind the hit:  ray.mint = hit.t + MachineEpsilon::E(hit.t)  (scene.cpp:675; ray.maxt is
//     never touched by the trace, so "ray->maxt = originalMaxT" is a no-op here); when that leaves no
//     interval (mint == t or mint >= maxt: "not enough numerical precision", scene.cpp:679-680) the ray
//     ends as a miss;
//   * every other ray is finished: RAY_FLAGS_MASKED is set, so the next trace leaves its RayHit untouched.
// *nContinuing counts the rays that are still armed (the caller's loop condition).
__global__ void __launch_bounds__(256) AdvanceRaysKernel(lrb_ray *__restrict__ rays, lrb_rayhit *__restrict__ hits, const uint32_t n,
		const uint32_t *__restrict__ passMesh, const uint32_t nPassWords, const uint8_t *__restrict__ continueFlags,
		uint32_t *__restrict__ nContinuing) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	bool armed = false;
	if (i < n) {
		lrb_ray *r = rays + i;
		const uint32_t flags = r->flags;
		if (!(flags & LRB_RAY_FLAGS_MASKED)) {
			lrb_rayhit *h = hits + i;
			const uint32_t mesh = h->meshIndex;
			bool cont = false;
			if (mesh != kNullIndex) {
				if (passMesh && (mesh >> 5) < nPassWords)
					cont = (passMesh[mesh >> 5] >> (mesh & 31u)) & 1u;
				if (continueFlags && continueFlags[i])
					cont = true;
			}
			if (cont) {
				const float t = h->t;
				const float mint = LRB_ADD(t, MachineEpsilonE(t));
				const float maxt = r->maxt;
				if (mint == t || mint >= maxt) {
					h->t = maxt;
					h->b1 = 0.f; h->b2 = 0.f;
					h->meshIndex = kNullIndex;
					h->triangleIndex = kNullIndex;
					r->flags = flags | LRB_RAY_FLAGS_MASKED;
				} else {
					r->mint = mint;
					armed = true;
				}
			} else
				r->flags = flags | LRB_RAY_FLAGS_MASKED;
		}
	}
	const unsigned m = __ballot_sync(0xffffffffu, armed);
	if ((threadIdx.x & 31u) == 0 && m)
		atomicAdd(nContinuing, (uint32_t)__popc(m));
}

// Dead-lane compaction between two launches (the reference re-launches a fixed-size Ray[taskCount] with dead
// lanes flagged RAY_FLAGS_MASKED, pathoclbase_kernels_micro.cl:34-106,1029): the indices of the live rays, in
// increasing order, as a dense list the persistent kernel consumes through TraceArgs::perm.  Three kernels:
// per-block live counts (ballot + popc), an exclusive scan of the block counts by one block, the scatter.
static const int kCompactBlock = 1024;      // rays per block of the count / scatter kernels (256 threads x 4)

__device__ __forceinline__ bool RayIsLive(const lrb_ray *rays, const uint32_t i) {
	return !(__ldg(&rays[i].flags) & LRB_RAY_FLAGS_MASKED);
}

__global__ void __launch_bounds__(256) CompactCountKernel(const lrb_ray *__restrict__ rays, const uint32_t n, uint32_t *__restrict__ blockCounts) {
	__shared__ uint32_t warpSum[8];
	const uint32_t base = blockIdx.x * kCompactBlock;
	uint32_t c = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t i = base + k * 256 + threadIdx.x;
		c += __popc(__ballot_sync(0xffffffffu, i < n && RayIsLive(rays, i)));
	}
	if ((threadIdx.x & 31u) == 0)
		warpSum[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < 8; ++w) t += warpSum[w];
		blockCounts[blockIdx.x] = t;
	}
}

// Exclusive scan of nBlocks counts in place (one 1024-thread block, sequential over 1024-wide tiles);
// total[0] receives the number of live rays.
__global__ void __launch_bounds__(1024) CompactScanKernel(uint32_t *__restrict__ blockCounts, const uint32_t nBlocks, uint32_t *__restrict__ total) {
	__shared__ uint32_t warpTot[32];
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	for (uint32_t tile = 0; tile < nBlocks; tile += 1024) {
		const uint32_t i = tile + threadIdx.x;
		const uint32_t v = i < nBlocks ? blockCounts[i] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
			if ((int)lane >= d) incl += o;
		}
		if (lane == 31) warpTot[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = warpTot[lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
				if ((int)lane >= d) w += o;
			}
			warpTot[lane] = w;      // inclusive over the warps
		}
		__syncthreads();
		const uint32_t before = carry + (warp ? warpTot[warp - 1] : 0u);
		if (i < nBlocks) blockCounts[i] = before + incl - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry = before + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) total[0] = carry;
}

// autoTotal (compaction in "auto" mode): the live count left by the scan; when most rays are live the list is not
// worth its scatter pass -- the trace kernels take the same decision from the same word (TraceArgs::permAuto).
__global__ void __launch_bounds__(256) CompactScatterKernel(const lrb_ray *__restrict__ rays, const uint32_t n,
		const uint32_t *__restrict__ blockOffsets, uint32_t *__restrict__ liveIdx, const uint32_t *__restrict__ autoTotal) {
	__shared__ uint32_t warpBase[33];
	if (autoTotal && (unsigned long long)autoTotal[0] * kCompactAutoDen > (unsigned long long)n * kCompactAutoNum)
		return;
	const uint32_t base = blockIdx.x * kCompactBlock;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	// ray k * 256 + threadIdx.x of the block: row k, warp `warp` -> 32 (row, warp) groups in index order
	unsigned m[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t i = base + k * 256 + threadIdx.x;
		m[k] = __ballot_sync(0xffffffffu, i < n && RayIsLive(rays, i));
		if (lane == 0) warpBase[k * 8 + warp + 1] = __popc(m[k]);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		warpBase[0] = blockOffsets[blockIdx.x];
		for (int g = 1; g <= 32; ++g) warpBase[g] += warpBase[g - 1];
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if ((m[k] >> lane) & 1u)
			liveIdx[warpBase[k * 8 + warp] + __popc(m[k] & ((1u << lane) - 1u))] = base + k * 256 + threadIdx.x;
	}
}

// ---- film merge over NVLink peer memory ------------------------------------------------------

// Sum of per-GPU film planes in device order, the arithmetic of the reference's merge of per-device films
// (PathOCLRenderEngine::MergeThreadFilms, src/slg/engines/pathocl/pathocl.cpp:184-201: film->Clear(), then
// Film::AddFilm of every device's film in device order; AddFilm adds pixel by pixel, channel by channel,
// src/slg/film/film.cpp:707-760):   dst[i] = (((0 + t_0[i]) + t_1[i]) + ...) + t_{n-1}[i]   in binary32.
// The tile pointers may be local or peer-mapped (CUDA IPC over NVLink), and so may dst: every rank runs this
// kernel over ITS slice of the film, pulling that slice of every other rank's planes with peer loads and
// storing the sums into the merged film on the gathering GPU -- a reduce-scatter and the gather of its result
// in one kernel per rank, no NCCL, no staging copy; the summation order does not depend on the rank count's
// schedule, so the result is the reference's bit for bit.
static const int kMaxFilmTiles = 16;
struct FilmTiles { const float *tile[kMaxFilmTiles]; };

__global__ void __launch_bounds__(256) FilmReduceKernel(const FilmTiles t, const int nTiles, float *__restrict__ dst,
		const unsigned long long first, const unsigned long long count, const int vec4) {
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (vec4) {
		// first and count are multiples of four floats and every pointer is 16-byte aligned
		const unsigned long long n4 = count >> 2, f4 = first >> 2;
		for (; i < n4; i += stride) {
			float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
			for (int r = 0; r < nTiles; ++r) {
				const float4 v = reinterpret_cast<const float4 *>(t.tile[r])[f4 + i];
				acc.x = LRB_ADD(acc.x, v.x); acc.y = LRB_ADD(acc.y, v.y); acc.z = LRB_ADD(acc.z, v.z); acc.w = LRB_ADD(acc.w, v.w);
			}
			reinterpret_cast<float4 *>(dst)[f4 + i] = acc;
		}
	} else {
		for (; i < count; i += stride) {
			float acc = 0.f;
			for (int r = 0; r < nTiles; ++r)
				acc = LRB_ADD(acc, t.tile[r][first + i]);
			dst[first + i] = acc;
		}
	}
}

}   // namespace lrb

#endif
