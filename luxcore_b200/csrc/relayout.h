// relayout.h -- host-side conversion of reference BVHArrayNode skip-list arrays into the
// WideNode / TriRecord / InstRecord layout of layout.h.  Pure C++ (no CUDA), so it is also
// compiled into the CPU-only unit tests.
#ifndef LRB_RELAYOUT_H
#define LRB_RELAYOUT_H

#include <stdint.h>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "luxrays_b200.h"
#include "layout.h"

namespace lrb {

// std::vector whose resize() leaves new elements of a trivial type UNINITIALISED instead of zero-filling them: the big
// arrays of a scene (50 M triangles: 5 GB) are written exactly once, by the threads of the re-layout's fill pass, and a
// serial zero-fill in front of that pass cost as much as the pass itself (and put every page on the resizing thread's
// NUMA node).  Every element must be written before it is read: the fill pass writes every wide node and record.
template <class T> struct DefaultInitAllocator : std::allocator<T> {
	template <class U> struct rebind { typedef DefaultInitAllocator<U> other; };
	DefaultInitAllocator() {}
	template <class U> DefaultInitAllocator(const DefaultInitAllocator<U> &) {}
	template <class U> void construct(U *p) { ::new ((void *)p) U; }
	template <class U, class A0, class... Args> void construct(U *p, A0 &&a0, Args &&... args) {
		::new ((void *)p) U(std::forward<A0>(a0), std::forward<Args>(args)...);
	}
};
template <class T> using RawVector = std::vector<T, DefaultInitAllocator<T> >;

struct WideScene {
	RawVector<WideNode> wide;
	RawVector<TriRecord> tris;
	RawVector<TriIds> ids;
	std::vector<InstRecord> insts;
	std::vector<DevInterp> interps;
	std::vector<uint32_t> motionFirst, motionLast;
	std::vector<float> minv;
	uint32_t rootWide;      // wide index of the (root tree's) root node
	uint32_t nRootWide;     // wide nodes belonging to the MBVH root tree (they are the LAST ones)
	uint32_t stackNeed;     // worst-case number of live stack entries
	uint32_t nRefNodes;
	bool twoLevel;
	float entryBox[6];      // exact box of the top-level tree's root (reference node 0): min xyz, max xyz
	std::vector<uint32_t> leafRootWide;   // per unique leaf: wide index of its root
	std::vector<uint32_t> leafStackNeed;
	std::vector<float> leafBox;           // per unique leaf: root box of its tree in instance space (min xyz, max xyz)

	WideScene() : rootWide(0), nRootWide(0), stackNeed(0), nRefNodes(0), twoLevel(false) { for (int i = 0; i < 6; ++i) entryBox[i] = 0.f; }
};

// Validates a reference array (skip indices in range and properly nested).  Returns false and
// fills `err` on a malformed tree; the device refuses such input instead of walking off the array.
bool ValidateTree(const lrb_bvh_node *nodes, uint32_t nNodes, std::string *err);

// Single-level BVH (what BVHKernel receives).  `xyz` = all meshes' vertices concatenated,
// meshVertexOffsets[m] = first vertex of mesh m.  Throws std::runtime_error on invalid input.
void BuildWideBVH(const lrb_bvh_node *nodes, uint32_t nNodes, const float *xyz, uint64_t nVerts,
		const uint32_t *meshVertexOffsets, uint32_t nMeshes, WideScene *out);

// Two-level MBVH (what MBVHKernel receives).
void BuildWideMBVH(const lrb_mbvh_desc &desc, WideScene *out);

// MBVHKernel::Update: replace the root tree (kept as the tail of `wide` / all of `insts`) and
// the inverse matrices.
void UpdateWideMBVHRoot(const lrb_bvh_node *rootNodes, uint32_t nRootNodes, const float *minv,
		uint32_t nTransforms, WideScene *scene);

// Fills nWide/rootWide/twoLevel/rootHasBox/rootChild/rootBox of a view from the host-side scene
// (pointers are left to the caller).
void FillRootOfView(const WideScene &scene, SceneView *view);

// Re-pack one 576-byte ocl::InterpolatedTransform into a DevInterp.
void PackInterp(const void *oclInterpolatedTransform, DevInterp *out);

}   // namespace lrb

#endif
