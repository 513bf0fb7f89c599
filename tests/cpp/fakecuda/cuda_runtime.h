// TEST SCAFFOLDING.  A stand-in for <cuda_runtime.h> that lets the REAL kernel source
// (luxcore_b200/csrc/trace_kernels.cuh) be compiled by g++ and executed on the host with one OS
// thread per lane: tests/cpp/kernel_lockstep.cpp.  Warp collectives (__ballot_sync, __shfl_sync,
// __syncwarp, __reduce_min_sync) are rendezvous of the 32 lane threads of a warp, so the kernel's
// loop -- re-fill, Resolve, votes, phases, stores: the part the per-ray emulation cannot see -- runs
// with its real control flow.  Only what the trace kernels use is provided.
#ifndef LRB_FAKE_CUDA_RUNTIME_H
#define LRB_FAKE_CUDA_RUNTIME_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__ __restrict

struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = { x, y, z, w }; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = { x, y }; return r; }

struct FakeDim3 { unsigned x, y, z; };
extern thread_local FakeDim3 threadIdx, blockIdx;
extern FakeDim3 blockDim, gridDim;

using std::max;
using std::min;

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
void __nanosleep(unsigned ns);      // harness: really sleeps, so that a spinning detector warp leaves the cores to the tracing lanes
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicExch(uint32_t *p, uint32_t v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
	unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { }
	return old;
}

// warp collectives: implemented by the harness (one rendezvous of the warp's lane threads each)
unsigned __ballot_sync(unsigned mask, int pred);
uint32_t __shfl_sync(unsigned mask, uint32_t v, int srcLane);
uint32_t __reduce_min_sync(unsigned mask, uint32_t v);
void __syncwarp();

#endif
