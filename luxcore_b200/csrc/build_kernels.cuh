// build_kernels.cuh -- BVH construction on the device (SURVEY.md 8f "GPU BVH builder").
//
// Replaces the reference's Embree Morton builder (BuildEmbreeBVHMorton, src/luxrays/core/bvh/
// bvhembreebuild.cpp:218-336 with rtcBVHBuilderMorton -- the fast, lower-quality flavour next to
// EMBREE_BINNED_SAH) by a linear BVH built entirely on the GPU, and emits what that function returns: the
// depth-first skip-list BVHArrayNode array (rules: bvhclassicbuild.cpp:181-220), with at most `treeType`
// children per node.  Steps, all data-parallel:
//   1. bounds of the leaf-box centroids (block reduction + atomics on ordered integer keys);
//   2. 63-bit Morton code of every centroid (21 bits per axis), radix-sorted with the leaf index (CUB);
//   3. binary radix tree over the sorted codes (Karras 2012: every inner node finds its key range and its split
//      independently; duplicate codes are told apart by their position);
//   4. depth of every inner node (walk up the parent links); a node is KEPT in the k-ary tree when its depth is a
//      multiple of log2(treeType) -- its children are then its descendants log2(treeType) levels down (or leaves
//      met earlier), at most treeType of them;
//   5. bottom-up: boxes (unions of the leaf boxes: exact float min / max) and array sizes of the collapsed
//      subtrees, one thread per leaf, the second thread to arrive at a node carries on (atomic flag);
//   6. array index of every node = sum over its ancestors of (1 if kept) + (size of the left sibling's subtree where
//      the path turns right): a second walk up;
//   7. emission: kept inner nodes write box + skip index, leaves write their input index (the host patches the
//      triangle / instance payload in, it owns those tables) + skip = index + 1 with bit 31.
// HBM-bound integer / pointer work: coalesced streams where the data allows it, no tensor cores.
#ifndef LRB_BUILD_KERNELS_CUH
#define LRB_BUILD_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "luxrays_b200.h"

namespace lrb {

// device scratch / events of one build: released on scope exit (error paths)
struct DevBuf {
	void *p;
	DevBuf() : p(nullptr) {}
	~DevBuf() { if (p) cudaFree(p); }
	template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct BuildEvents {
	cudaEvent_t *e;
	int n;
	~BuildEvents() { for (int i = 0; i < n; ++i) cudaEventDestroy(e[i]); }
};

static const uint32_t kLeafBit = 0x80000000u;   // child reference: leaf (sorted position) instead of inner node

// order-preserving float <-> uint map for atomicMin / atomicMax
__device__ __forceinline__ uint32_t OrderedKey(float f) {
	const uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float FromOrderedKey(uint32_t k) {
	const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#if defined(__CUDA_ARCH__)
	return __uint_as_float(u);
#else
	float f;
	memcpy(&f, &u, 4);
	return f;
#endif
}

// bounds[0..2] = min keys, bounds[3..5] = max keys of the centroids (x2: lo + hi, like the reference's builders)
__global__ void __launch_bounds__(256) CentroidBoundsKernel(const float *__restrict__ boxes, const uint32_t n, uint32_t *bounds) {
	float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float *b = boxes + 6 * (size_t)i;
		for (int a = 0; a < 3; ++a) {
			const float c = b[a] + b[3 + a];
			if (c == c && fabsf(c) < 3.0e38f) {      // NaN / inf boxes do not stretch the grid (they get code 0)
				lo[a] = fminf(lo[a], c);
				hi[a] = fmaxf(hi[a], c);
			}
		}
	}
	for (int a = 0; a < 3; ++a) {
		for (int o = 16; o > 0; o >>= 1) {
			lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
			hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
		}
		if ((threadIdx.x & 31) == 0) {
			atomicMin(bounds + a, OrderedKey(lo[a]));
			atomicMax(bounds + 3 + a, OrderedKey(hi[a]));
		}
	}
}

__device__ __forceinline__ uint64_t Spread21(uint64_t v) {      // 21 bits -> every third bit
	v &= 0x1fffffull;
	v = (v | (v << 32)) & 0x1f00000000ffffull;
	v = (v | (v << 16)) & 0x1f0000ff0000ffull;
	v = (v | (v << 8)) & 0x100f00f00f00f00full;
	v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
	v = (v | (v << 2)) & 0x1249249249249249ull;
	return v;
}

__global__ void __launch_bounds__(256) MortonKernel(const float *__restrict__ boxes, const uint32_t n, const uint32_t *__restrict__ bounds,
		uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const float *b = boxes + 6 * (size_t)i;
	uint64_t code = 0;
	for (int a = 0; a < 3; ++a) {
		const float lo = FromOrderedKey(bounds[a]), hi = FromOrderedKey(bounds[3 + a]);
		const float c = b[a] + b[3 + a];
		const float ext = hi - lo;
		float u = ext > 0.f ? (c - lo) / ext : 0.f;
		u = (u == u) ? fminf(fmaxf(u, 0.f), 1.f) : 0.f;
		const uint64_t q = (uint64_t)min(2097151.f, u * 2097152.f);
		code |= Spread21(q) << a;
	}
	keys[i] = code;
	vals[i] = i;
}

// number of leading bits two sorted positions share; ties in the code are broken by the position itself
__device__ __forceinline__ int Delta(const uint64_t *__restrict__ keys, const int n, const int i, const int j) {
	if (j < 0 || j >= n)
		return -1;
	const uint64_t a = keys[i], b = keys[j];
	return a == b ? 64 + __clz(i ^ j) : __clzll((long long)(a ^ b));
}

// Karras 2012, one thread per inner node (n - 1 of them; node 0 is the root).
__global__ void __launch_bounds__(256) RadixTreeKernel(const uint64_t *__restrict__ keys, const int n, uint32_t *__restrict__ left, uint32_t *__restrict__ right,
		uint32_t *__restrict__ parentOfInner, uint32_t *__restrict__ parentOfLeaf) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1)
		return;
	const int d = (Delta(keys, n, i, i + 1) - Delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	const int dMin = Delta(keys, n, i, i - d);
	int lMax = 2;
	while (Delta(keys, n, i, i + lMax * d) > dMin)
		lMax <<= 1;
	int l = 0;
	for (int t = lMax >> 1; t >= 1; t >>= 1)
		if (Delta(keys, n, i, i + (l + t) * d) > dMin)
			l += t;
	const int j = i + l * d;
	const int dNode = Delta(keys, n, i, j);
	int s = 0;
	for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
		if (Delta(keys, n, i, i + (s + t) * d) > dNode)
			s += t;
		if (t == 1)
			break;
	}
	const int split = i + s * d + min(d, 0);
	const int lo = min(i, j), hi = max(i, j);
	const uint32_t L = (lo == split) ? (kLeafBit | (uint32_t)split) : (uint32_t)split;
	const uint32_t R = (hi == split + 1) ? (kLeafBit | (uint32_t)(split + 1)) : (uint32_t)(split + 1);
	left[i] = L;
	right[i] = R;
	if (L & kLeafBit) parentOfLeaf[L & ~kLeafBit] = (uint32_t)i; else parentOfInner[L] = (uint32_t)i;
	if (R & kLeafBit) parentOfLeaf[R & ~kLeafBit] = (uint32_t)i; else parentOfInner[R] = (uint32_t)i;
	if (i == 0)
		parentOfInner[0] = 0xffffffffu;
}

__global__ void __launch_bounds__(256) DepthKernel(const uint32_t *__restrict__ parentOfInner, const int nInner, uint32_t *__restrict__ depth) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nInner)
		return;
	uint32_t d = 0;
	for (uint32_t p = parentOfInner[i]; p != 0xffffffffu; p = parentOfInner[p])
		++d;
	depth[i] = d;
}

// One thread per leaf walks up; the second arrival at an inner node combines its children.
//   box[i]  = union of the children's boxes
//   size[i] = array nodes of the collapsed subtree below (and including, when kept) inner node i
__global__ void __launch_bounds__(256) BottomUpKernel(const float *__restrict__ leafBoxes, const uint32_t *__restrict__ sortedLeaf, const int n,
		const uint32_t *__restrict__ left, const uint32_t *__restrict__ right, const uint32_t *__restrict__ parentOfInner,
		const uint32_t *__restrict__ parentOfLeaf, const uint32_t *__restrict__ depth, const uint32_t levelStep,
		float *box, uint32_t *size, uint32_t *arrived) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	uint32_t p = parentOfLeaf[i];
	while (p != 0xffffffffu) {
		__threadfence();
		if (atomicAdd(arrived + p, 1u) == 0u)
			return;             // the sibling subtree is not finished: its thread will carry on from here
		__threadfence();
		float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
		uint32_t sz = (depth[p] % levelStep == 0u) ? 1u : 0u;
		const uint32_t kids[2] = { left[p], right[p] };
		for (int k = 0; k < 2; ++k) {
			const uint32_t c = kids[k];
			const volatile float *b;
			if (c & kLeafBit) {
				b = leafBoxes + 6 * (size_t)sortedLeaf[c & ~kLeafBit];
				sz += 1u;
			} else {
				b = box + 6 * (size_t)c;
				sz += ((const volatile uint32_t *)size)[c];
			}
			for (int a = 0; a < 3; ++a) {
				// (the reference's Union: NaN coordinates never replace a bound)
				const float l = b[a], h = b[3 + a];
				lo[a] = l < lo[a] ? l : lo[a];
				hi[a] = h > hi[a] ? h : hi[a];
			}
		}
		for (int a = 0; a < 3; ++a) {
			box[6 * (size_t)p + a] = lo[a];
			box[6 * (size_t)p + 3 + a] = hi[a];
		}
		size[p] = sz;
		p = parentOfInner[p];
	}
}

// Array index of a node from its ancestors (see the file comment).  `self` = inner index, or kLeafBit | sorted position.
__device__ __forceinline__ uint32_t ArrayIndexOf(const uint32_t self, uint32_t p, const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
		const uint32_t *__restrict__ parentOfInner, const uint32_t *__restrict__ depth, const uint32_t *__restrict__ size, const uint32_t levelStep) {
	uint32_t idx = 0, child = self;
	while (p != 0xffffffffu) {
		if (depth[p] % levelStep == 0u)
			idx += 1u;
		if (right[p] == child) {
			const uint32_t l = left[p];
			idx += (l & kLeafBit) ? 1u : size[l];
		}
		child = p;
		p = parentOfInner[p];
	}
	return idx;
}

__global__ void __launch_bounds__(256) EmitInnerKernel(const int nInner, const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
		const uint32_t *__restrict__ parentOfInner, const uint32_t *__restrict__ depth, const uint32_t *__restrict__ size, const float *__restrict__ box,
		const uint32_t levelStep, lrb_bvh_node *__restrict__ out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nInner || depth[i] % levelStep != 0u)
		return;
	const uint32_t idx = ArrayIndexOf((uint32_t)i, parentOfInner[i], left, right, parentOfInner, depth, size, levelStep);
	lrb_bvh_node nd;
	for (int a = 0; a < 3; ++a) {
		nd.bvhNode.bboxMin[a] = box[6 * (size_t)i + a];
		nd.bvhNode.bboxMax[a] = box[6 * (size_t)i + 3 + a];
	}
	nd.nodeData = idx + size[i];
	nd.pad0 = 0;
	out[idx] = nd;
}

__global__ void __launch_bounds__(256) EmitLeafKernel(const int n, const uint32_t *__restrict__ sortedLeaf, const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
		const uint32_t *__restrict__ parentOfInner, const uint32_t *__restrict__ parentOfLeaf, const uint32_t *__restrict__ depth,
		const uint32_t *__restrict__ size, const uint32_t levelStep, lrb_bvh_node *__restrict__ out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t idx = ArrayIndexOf(kLeafBit | (uint32_t)i, parentOfLeaf[i], left, right, parentOfInner, depth, size, levelStep);
	lrb_bvh_node nd;
	memset(&nd, 0, sizeof(nd));
	nd.triangleLeaf.v[0] = sortedLeaf[i];       // input index of the leaf: the caller patches its payload in
	nd.nodeData = (idx + 1u) | 0x80000000u;
	out[idx] = nd;
}

}   // namespace lrb

#endif
