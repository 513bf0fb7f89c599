// relayout.h -- host-side conversion of reference BVHArrayNode skip-list arrays into the
// WideNode / TriRecord / InstRecord layout of layout.h.  Pure C++ (no CUDA), so it is also
// compiled into the CPU-only unit tests.
#ifndef LRB_RELAYOUT_H
#define LRB_RELAYOUT_H

#include <stdint.h>
#include <string>
#include <vector>

#include "luxrays_b200.h"
#include "layout.h"

namespace lrb {

struct WideScene {
	std::vector<WideNode> wide;
	std::vector<TriRecord> tris;
	std::vector<TriIds> ids;
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
