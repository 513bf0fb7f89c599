// relayout_kernels.cuh -- the re-layout of layout.h on the device, for trees that were built there (lrb_bvh_build_scene).
//
// The host re-layout (relayout.cpp) serves any reference array that arrives through lrb_bvh_upload; for a 50 M-triangle
// scene it is 6 s of host work next to a 0.1 s device build, plus 2.4 GB of array down and 5.1 GB of lay-out up.  Here
// the BVHArrayNode array the builder kernels emitted (build_kernels.cuh) never leaves the device:
//   0. LeafBoxKernel        one thread per triangle: the box BVHAccel::Init hands its builders (bvhaccel.cpp:116-122);
//   -- the builder (BuildTreeOnDevice) --
//   1. LeafPayloadKernel    the builder's leaves carry the input triangle number: write v[3] / meshIndex / triangleIndex
//                           (bvhclassicbuild.cpp:196-214) so that the array is the one a host would download;
//   2. RelayoutCountKernel  per reference node: 1 leaf, or the wide nodes an inner node becomes (64-bit pair) ->
//      CUB exclusive sum  -> RelayoutIndexKernel: wide-node index of every inner node, TriRecord index of every leaf,
//      both in array order (the record index is the reference's tie-break order, layout.h);
//   3. RelayoutFillKernel   one thread per inner node: ConvertInnerNodeTri of relayout_shared.h -- the SAME function the
//                           host's second pass calls, so the bytes are the host's bytes;
//   4. StackNeedKernel      worst-case live stack entries, bottom-up over the wide nodes (the host's reverse sweep).
// HBM-bound gather / scatter of 32-B and 64-B records: one thread per record, no tensor cores.
#ifndef LRB_RELAYOUT_KERNELS_CUH
#define LRB_RELAYOUT_KERNELS_CUH

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "relayout_shared.h"

namespace lrb {

// mesh of global triangle g: meshTriOff[m] <= g < meshTriOff[m + 1] (empty meshes have equal offsets and are skipped)
LRB_RHD uint32_t MeshOfTriangle(const uint32_t *meshTriOff, const uint32_t nMeshes, const uint32_t g) {
	uint32_t lo = 0, hi = nMeshes;      // invariant: meshTriOff[lo] <= g < meshTriOff[hi]
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (meshTriOff[mid] <= g)
			lo = mid;
		else
			hi = mid;
	}
	return lo;
}

// ---- per-record bodies (host + device: the CPU test drives them with plain loops) ---------------------------------------

LRB_RHD int LeafBoxBody(const float *xyz, const uint64_t nVerts, const uint32_t *meshVertOff, const uint32_t *meshTriOff, const uint32_t nMeshes,
		const uint32_t *triIdx, const uint32_t g, float *box) {
	const uint32_t m = MeshOfTriangle(meshTriOff, nMeshes, g);
	const float *p[3];
	for (int j = 0; j < 3; ++j) {
		const uint64_t v = (uint64_t)triIdx[3 * (size_t)g + j] + meshVertOff[m];
		if (v >= nVerts)
			return kRelayoutBadVertex;
		p[j] = xyz + 3 * v;
	}
	TriBuildBoxOf(p[0], p[1], p[2], box, box + 3);
	return kRelayoutOk;
}

LRB_RHD void LeafPayloadBody(lrb_bvh_node *nd, const uint32_t *meshTriOff, const uint32_t nMeshes, const uint32_t *triIdx) {
	if (!RlIsLeaf(nd->nodeData))
		return;
	const uint32_t g = nd->triangleLeaf.v[0];
	const uint32_t m = MeshOfTriangle(meshTriOff, nMeshes, g);
	nd->triangleLeaf.v[0] = triIdx[3 * (size_t)g];
	nd->triangleLeaf.v[1] = triIdx[3 * (size_t)g + 1];
	nd->triangleLeaf.v[2] = triIdx[3 * (size_t)g + 2];
	nd->triangleLeaf.meshIndex = m;
	nd->triangleLeaf.triangleIndex = g - meshTriOff[m];
}

// low word: leaves, high word: wide nodes
LRB_RHD unsigned long long RelayoutCountBody(const lrb_bvh_node *nodes, const uint32_t i) {
	if (RlIsLeaf(nodes[i].nodeData))
		return 1ull;
	return (unsigned long long)WideNodesFor(CountKids(nodes, i)) << 32;
}

// wideStart = 0, the entry node is wide node 0, the first inner node (the root) gets wide node 1
LRB_RHD uint32_t RelayoutIndexBody(const lrb_bvh_node *nodes, const uint32_t i, const unsigned long long scanned) {
	return RlIsLeaf(nodes[i].nodeData) ? (uint32_t)(scanned & 0xffffffffull) : 1u + (uint32_t)(scanned >> 32);
}

// number of entries of wide node w that are wide nodes themselves (inner children + continuation)
LRB_RHD uint32_t InnerEntriesOf(const WideNode &w) {
	uint32_t c = w.next != kNullIndex ? 1u : 0u;
	const uint32_t nChild = NodeSlots(w);
	for (uint32_t k = 0; k < nChild; ++k)
		if (!(w.child[k] & (kTagTri | kTagInstance)))
			++c;
	return c;
}

#if defined(__CUDACC__) && !defined(LRB_FAKE_CUDA_RUNTIME_H)

__global__ void __launch_bounds__(256) LeafBoxKernel(const float *__restrict__ xyz, const uint64_t nVerts, const uint32_t *__restrict__ meshVertOff,
		const uint32_t *__restrict__ meshTriOff, const uint32_t nMeshes, const uint32_t *__restrict__ triIdx, const uint32_t nTris, float *__restrict__ boxes,
		uint32_t *err) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nTris)
		return;
	const int rc = LeafBoxBody(xyz, nVerts, meshVertOff, meshTriOff, nMeshes, triIdx, g, boxes + 6 * (size_t)g);
	if (rc != kRelayoutOk) {
		for (int k = 0; k < 6; ++k) boxes[6 * (size_t)g + k] = 0.f;
		atomicMax(err, (uint32_t)rc);
	}
}

__global__ void __launch_bounds__(256) LeafPayloadKernel(lrb_bvh_node *nodes, const uint32_t n, const uint32_t *__restrict__ meshTriOff, const uint32_t nMeshes,
		const uint32_t *__restrict__ triIdx) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		LeafPayloadBody(nodes + i, meshTriOff, nMeshes, triIdx);
}

__global__ void __launch_bounds__(256) RelayoutCountKernel(const lrb_bvh_node *__restrict__ nodes, const uint32_t n, unsigned long long *__restrict__ counts) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		counts[i] = RelayoutCountBody(nodes, i);
}

__global__ void __launch_bounds__(256) RelayoutIndexKernel(const lrb_bvh_node *__restrict__ nodes, const uint32_t n, const unsigned long long *__restrict__ scanned,
		uint32_t *__restrict__ wideOf) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		wideOf[i] = RelayoutIndexBody(nodes, i, scanned[i]);
}

// One thread per reference node; inner nodes write their wide node(s) and their triangle records.  Thread 0 also writes
// the entry node (wide node 0) and the exact root box.
__global__ void __launch_bounds__(128) RelayoutFillKernel(const TriTreeView in, const uint32_t *__restrict__ wideOf, WideNode *wide, TriRecord *tris, TriIds *ids,
		uint32_t *parentOf, float *entryBox, uint32_t *err) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= in.n)
		return;
	if (i == 0) {
		const int rc = MakeEntryNode(in.nodes[0], wideOf[0], &wide[0], entryBox);
		parentOf[0] = kNullIndex;
		parentOf[wideOf[0]] = 0u;
		if (rc != kRelayoutOk)
			atomicMax(err, (uint32_t)rc);
	}
	if (RlIsLeaf(in.nodes[i].nodeData))
		return;
	const int rc = ConvertInnerNodeTri(in, i, wideOf, wide, tris, ids, parentOf);
	if (rc != kRelayoutOk)
		atomicMax(err, (uint32_t)rc);
}

// D[w] = (slots of w, all four, + continuation) - 1 + max over its wide-node entries D[entry]  (relayout.cpp: the sweep at
// the end of ConvertTree).  One thread per wide node without wide-node entries starts; the last arrival at a node carries on.
// below[] and arrived[] are zeroed by the caller; result = D[0] (the entry node).
__global__ void __launch_bounds__(256) StackNeedKernel(const WideNode *__restrict__ wide, const uint32_t nWide, const uint32_t *__restrict__ parentOf,
		uint32_t *below, uint32_t *arrived, uint32_t *result) {
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nWide)
		return;
	if (InnerEntriesOf(wide[w]) != 0u)
		return;
	uint32_t cur = w;
	uint32_t D = StackNeedOfNode(wide[cur], 0u);
	for (uint32_t hops = 0; hops <= nWide; ++hops) {    // a walk up is shorter than the node count (bound against corrupt links)
		const uint32_t p = parentOf[cur];
		if (p >= nWide) {
			if (cur == 0u)
				*result = D;        // the entry node has no parent
			return;
		}
		atomicMax(below + p, D);
		__threadfence();
		const uint32_t seen = atomicAdd(arrived + p, 1u) + 1u;
		if (seen < InnerEntriesOf(wide[p]))
			return;             // another subtree below p is not finished: its thread carries on from here
		__threadfence();
		cur = p;
		D = StackNeedOfNode(wide[cur], atomicMax(below + cur, 0u));
	}
}

#endif  // __CUDACC__

}   // namespace lrb

#endif
