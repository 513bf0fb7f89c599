// build_kernels.cuh -- BVH construction on the device (SURVEY.md 8f "GPU BVH builder").
//
// Replaces the reference's Embree builders for trees built where the rays are traced (BuildEmbreeBVHMorton /
// BuildEmbreeBVHBinnedSAH, src/luxrays/core/bvh/bvhembreebuild.cpp:218-336: 5-17 s on the authors' machine, :289-322)
// and emits what those functions return: the depth-first skip-list BVHArrayNode array (rules: bvhclassicbuild.cpp:
// 181-220), with at most `treeType` children per node.  Everything is data-parallel; ids: leaves are 0 .. n-1 in
// Morton order, inner nodes n .. 2n-2.
//   1. bounds of the leaf-box centroids (warp reduction + atomics on ordered integer keys);
//   2. 63-bit Morton code of every centroid (21 bits per axis), radix-sorted with the leaf index (CUB);
//   3. the binary tree, one of
//        quality 0 (EMBREE_MORTON): Karras' radix tree over the sorted codes (every inner node finds its key range and
//          its split independently; duplicate codes are told apart by their position), boxes bottom-up;
//        quality 1: PLOC -- parallel locally-ordered clustering (Meister & Bittner 2018): every cluster looks `radius`
//          places to both sides in Morton order for the partner whose union with it has the smallest surface area;
//          mutual choices merge; the cluster list is compacted; repeat until one cluster is left.  Close to a full
//          SAH sweep in quality (kitchen, CPU prototype: 20.7 node visits per bounce ray against 31.9 for the radix
//          tree and 15.6 for the host builder with its re-insertion passes), a few hundred small launches in cost;
//   4. k-ary collapse, top-down over a frontier: a kept node adopts its two children and keeps opening the inner
//      child of largest surface area until it holds treeType children; the inner children it ends up with are the
//      next frontier;
//   5. bottom-up: array sizes of the collapsed subtrees (one thread per leaf, the second thread to arrive at a node
//      carries on; the same pass computes the boxes of a radix tree before step 4);
//   6. array index of every node = sum over its ancestors of (1 if kept) + (size of the left sibling's subtree where
//      the path turns right): a walk up the parent links;
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

static const uint32_t kNoNode = 0xffffffffu;

// order-preserving float <-> uint map for atomicMin / atomicMax
__device__ __forceinline__ uint32_t OrderedKey(float f) {
	const uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float FromOrderedKey(uint32_t k) {
	return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
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

__global__ void __launch_bounds__(256) IotaKernel(uint32_t *__restrict__ v, const uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		v[i] = i;
}

// node boxes of the leaves, in Morton order (node id = sorted position)
__global__ void __launch_bounds__(256) GatherLeafBoxesKernel(const float *__restrict__ leafBoxes, const uint32_t *__restrict__ sortedLeaf, const uint32_t n,
		float *__restrict__ nodeBox) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const float *b = leafBoxes + 6 * (size_t)sortedLeaf[i];
	float *o = nodeBox + 6 * (size_t)i;
	for (int a = 0; a < 6; ++a)
		o[a] = b[a];
}

// ---- quality 0: Karras' radix tree ------------------------------------------------------------------------

// number of leading bits two sorted positions share; ties in the code are broken by the position itself
__device__ __forceinline__ int Delta(const uint64_t *__restrict__ keys, const int n, const int i, const int j) {
	if (j < 0 || j >= n)
		return -1;
	const uint64_t a = keys[i], b = keys[j];
	return a == b ? 64 + __clz(i ^ j) : __clzll((long long)(a ^ b));
}

// One thread per inner node (inner node i has id n + i; inner node 0 is the root).
__global__ void __launch_bounds__(256) RadixTreeKernel(const uint64_t *__restrict__ keys, const int n, uint32_t *__restrict__ left, uint32_t *__restrict__ right,
		uint32_t *__restrict__ parent) {
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
	const uint32_t self = (uint32_t)(n + i);
	const uint32_t L = (lo == split) ? (uint32_t)split : (uint32_t)(n + split);
	const uint32_t R = (hi == split + 1) ? (uint32_t)(split + 1) : (uint32_t)(n + split + 1);
	left[self] = L;
	right[self] = R;
	parent[L] = self;
	parent[R] = self;
	if (i == 0)
		parent[self] = kNoNode;
}

// ---- quality 1: PLOC ----------------------------------------------------------------------------------------

__device__ __forceinline__ float HalfArea(const float lx, const float ly, const float lz, const float hx, const float hy, const float hz) {
	const float dx = fmaxf(hx - lx, 0.f), dy = fmaxf(hy - ly, 0.f), dz = fmaxf(hz - lz, 0.f);
	return dx * dy + dy * dz + dz * dx;
}

// nn[i] = position (in the cluster list) of the partner of cluster i: smallest surface area of the union among the
// clusters at most `radius` places away.  Equal areas (duplicated geometry) are told apart by a key that both ends of
// a pair compute alike, so that tied neighbourhoods still pair up at random instead of forming one long chain of
// one-sided choices (which would merge a single pair per iteration).
__device__ __forceinline__ uint32_t PairKey(const int i, const int j) {
	const uint32_t a = (uint32_t)min(i, j), b = (uint32_t)max(i, j);
	uint32_t h = a * 0x9E3779B1u ^ (b * 0x85EBCA6Bu + 0xC2B2AE35u);
	h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
	return h;
}

__global__ void __launch_bounds__(256) PlocNearestKernel(const uint32_t *__restrict__ clusters, const int m, const float *__restrict__ nodeBox, const int radius,
		int *__restrict__ nn) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m)
		return;
	const float *b = nodeBox + 6 * (size_t)clusters[i];
	const float lx = b[0], ly = b[1], lz = b[2], hx = b[3], hy = b[4], hz = b[5];
	float best = 0.f;
	uint32_t bestKey = 0;
	int arg = -1;
	for (int off = -radius; off <= radius; ++off) {
		const int j = i + off;
		if (off == 0 || j < 0 || j >= m)
			continue;
		const float *c = nodeBox + 6 * (size_t)clusters[j];
		float a = HalfArea(fminf(lx, c[0]), fminf(ly, c[1]), fminf(lz, c[2]), fmaxf(hx, c[3]), fmaxf(hy, c[4]), fmaxf(hz, c[5]));
		a = (a == a) ? a : 3.4e38f;
		const uint32_t key = PairKey(i, j);
		if (arg < 0 || a < best || (a == best && key < bestKey)) {
			best = a;
			bestKey = key;
			arg = j;
		}
	}
	nn[i] = arg;
}

// Mutual partners merge: the one with the smaller position creates the node (ids are handed out by an atomic counter:
// the ids vary from run to run, the tree does not), the other one leaves the list.
__global__ void __launch_bounds__(256) PlocMergeKernel(const uint32_t *__restrict__ clusters, const int m, const int *__restrict__ nn, const uint32_t n,
		float *nodeBox, uint32_t *__restrict__ left, uint32_t *__restrict__ right, uint32_t *__restrict__ parent,
		uint32_t *nextInner, uint32_t *__restrict__ clustersOut, uint8_t *__restrict__ keep) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m)
		return;
	const int j = nn[i];
	const bool mutual = j >= 0 && nn[j] == i;
	uint32_t self = clusters[i];
	if (mutual && i < j) {
		const uint32_t other = clusters[j];
		const uint32_t id = n + atomicAdd(nextInner, 1u);
		const float *a = nodeBox + 6 * (size_t)self, *b = nodeBox + 6 * (size_t)other;
		float *o = nodeBox + 6 * (size_t)id;
		for (int k = 0; k < 3; ++k) {
			// (the reference's Union: NaN coordinates never replace a bound)
			const float l0 = a[k], l1 = b[k], h0 = a[3 + k], h1 = b[3 + k];
			o[k] = l1 < l0 ? l1 : l0;
			o[3 + k] = h1 > h0 ? h1 : h0;
		}
		left[id] = self;
		right[id] = other;
		parent[self] = id;
		parent[other] = id;
		parent[id] = kNoNode;
		self = id;
	}
	clustersOut[i] = self;
	keep[i] = (mutual && i > j) ? 0 : 1;
}

// ---- k-ary collapse over a frontier ---------------------------------------------------------------------

__device__ __forceinline__ float OpenKey(const float *__restrict__ nodeBox, const uint32_t id, const uint32_t n) {
	if (id < n)
		return -1.f;            // leaves are never opened
	const float *b = nodeBox + 6 * (size_t)id;
	const float a = HalfArea(b[0], b[1], b[2], b[3], b[4], b[5]);
	return a == a ? a : 3.4e38f;        // NaN boxes: open them first
}

// One thread per kept node of the current level.
__global__ void __launch_bounds__(128) CollapseKernel(const uint32_t *__restrict__ frontier, const uint32_t count, const uint32_t n, const uint32_t treeType,
		const uint32_t *__restrict__ left, const uint32_t *__restrict__ right, const float *__restrict__ nodeBox, uint8_t *__restrict__ kept,
		uint32_t *__restrict__ nextFrontier, uint32_t *nextCount) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= count)
		return;
	const uint32_t node = frontier[t];
	kept[node - n] = 1;
	uint32_t kids[9];
	float key[9];
	uint32_t nk = 2;
	kids[0] = left[node];
	kids[1] = right[node];
	key[0] = OpenKey(nodeBox, kids[0], n);
	key[1] = OpenKey(nodeBox, kids[1], n);
	while (nk < treeType) {
		int best = -1;
		for (uint32_t k = 0; k < nk; ++k)
			if (kids[k] >= n && (best < 0 || key[k] > key[best]))
				best = (int)k;
		if (best < 0)
			break;
		// replace kid `best` by its two children, keeping the left-to-right order of the binary tree
		const uint32_t c = kids[best];
		for (int k = (int)nk; k > best + 1; --k) {
			kids[k] = kids[k - 1];
			key[k] = key[k - 1];
		}
		kids[best] = left[c];
		kids[best + 1] = right[c];
		key[best] = OpenKey(nodeBox, kids[best], n);
		key[best + 1] = OpenKey(nodeBox, kids[best + 1], n);
		++nk;
	}
	for (uint32_t k = 0; k < nk; ++k)
		if (kids[k] >= n)
			nextFrontier[atomicAdd(nextCount, 1u)] = kids[k];
}

// ---- bottom-up passes --------------------------------------------------------------------------------------

// One thread per leaf walks up; the second arrival at an inner node combines its children.
//   BOXES: nodeBox[id] = union of the children's boxes (radix tree)
//   SIZES: size[id - n] = array nodes of the collapsed subtree below (and including, when kept) inner node id
template <bool BOXES, bool SIZES>
__global__ void __launch_bounds__(256) BottomUpKernel(const uint32_t n, const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
		const uint32_t *__restrict__ parent, const uint8_t *__restrict__ kept, float *nodeBox, uint32_t *size, uint32_t *arrived) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	uint32_t p = parent[i];
	while (p != kNoNode) {
		__threadfence();
		if (atomicAdd(arrived + (p - n), 1u) == 0u)
			return;             // the sibling subtree is not finished: its thread will carry on from here
		__threadfence();
		const uint32_t kids[2] = { left[p], right[p] };
		if (BOXES) {
			float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
			for (int k = 0; k < 2; ++k) {
				const volatile float *b = nodeBox + 6 * (size_t)kids[k];
				for (int a = 0; a < 3; ++a) {
					const float l = b[a], h = b[3 + a];
					lo[a] = l < lo[a] ? l : lo[a];
					hi[a] = h > hi[a] ? h : hi[a];
				}
			}
			for (int a = 0; a < 3; ++a) {
				nodeBox[6 * (size_t)p + a] = lo[a];
				nodeBox[6 * (size_t)p + 3 + a] = hi[a];
			}
		}
		if (SIZES) {
			uint32_t sz = kept[p - n] ? 1u : 0u;
			for (int k = 0; k < 2; ++k)
				sz += kids[k] < n ? 1u : ((const volatile uint32_t *)size)[kids[k] - n];
			size[p - n] = sz;
		}
		p = parent[p];
	}
}

// Array index of a node from its ancestors (see the file comment).
__device__ __forceinline__ uint32_t ArrayIndexOf(const uint32_t self, const uint32_t n, const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
		const uint32_t *__restrict__ parent, const uint8_t *__restrict__ kept, const uint32_t *__restrict__ size) {
	uint32_t idx = 0, child = self;
	for (uint32_t p = parent[self]; p != kNoNode; p = parent[p]) {
		if (kept[p - n])
			idx += 1u;
		if (right[p] == child) {
			const uint32_t l = left[p];
			idx += l < n ? 1u : size[l - n];
		}
		child = p;
	}
	return idx;
}

// One thread per node id (leaves and inner nodes alike).
__global__ void __launch_bounds__(256) EmitKernel(const uint32_t n, const uint32_t *__restrict__ sortedLeaf, const uint32_t *__restrict__ left,
		const uint32_t *__restrict__ right, const uint32_t *__restrict__ parent, const uint8_t *__restrict__ kept, const uint32_t *__restrict__ size,
		const float *__restrict__ nodeBox, lrb_bvh_node *__restrict__ out) {
	const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= 2u * n - 1u)
		return;
	lrb_bvh_node nd;
	memset(&nd, 0, sizeof(nd));
	if (id < n) {
		const uint32_t idx = ArrayIndexOf(id, n, left, right, parent, kept, size);
		nd.triangleLeaf.v[0] = sortedLeaf[id];      // input index of the leaf: the caller patches its payload in
		nd.nodeData = (idx + 1u) | 0x80000000u;
		out[idx] = nd;
	} else if (kept[id - n]) {
		const uint32_t idx = ArrayIndexOf(id, n, left, right, parent, kept, size);
		for (int a = 0; a < 3; ++a) {
			nd.bvhNode.bboxMin[a] = nodeBox[6 * (size_t)id + a];
			nd.bvhNode.bboxMax[a] = nodeBox[6 * (size_t)id + 3 + a];
		}
		nd.nodeData = idx + size[id - n];
		out[idx] = nd;
	}
}

}   // namespace lrb

#endif
