// luxrays/core/bvh/bvhbuild.h -- flattened BVH node type and the host-side builders (reference:
// include/luxrays/core/bvh/bvhbuild.h:32-86, bvhbuild_types.cl:21-42).
//
// Builders emit the reference's depth-first skip-list array of 32-byte BVHArrayNode records:
//   CLASSIC            restated from src/luxrays/core/bvh/bvhclassicbuild.cpp (bit-identical arrays)
//   EMBREE_BINNED_SAH  the reference calls Embree's rtcBuildBVH (bvhembreebuild.cpp:218-280), a third-party
//                      library; here this tree's own from-scratch binned-SAH builder with insertion-based
//                      optimisation and optimal k-ary collapse (luxcore_b200/host/bvhbuild.cpp);
//   EMBREE_MORTON      the fast builder (rtcBVHBuilderMorton there): here a radix tree built ON THE GPU
//   B200_PLOC          (extension) the GPU builder with a PLOC binary tree (luxcore_b200/csrc/build_kernels.cuh).
//   All emit the same array format under the same contract: one triangle per leaf, <= treeType children
//   per node; topology is ours and never changes a closest hit.
#ifndef _LUXRAYS_B200_BVHBUILD_H
#define _LUXRAYS_B200_BVHBUILD_H

#include <deque>
#include <vector>

#include "luxrays/luxrays.h"
#include "luxrays/core/trianglemesh.h"

namespace luxrays {

namespace ocl {
// bit-identical to ocl::BVHArrayNode
typedef struct {
	union {
		struct { float bboxMin[3]; float bboxMax[3]; } bvhNode;
		struct { unsigned int v[3]; unsigned int meshIndex, triangleIndex; } triangleLeaf;
		struct { unsigned int leafIndex; unsigned int transformIndex, motionIndex; unsigned int meshOffsetIndex; } bvhLeaf;
	};
	unsigned int nodeData;
	int pad0;
} BVHArrayNode;

typedef struct {
	unsigned int interpolatedTransformFirstIndex, interpolatedTransformLastIndex;
	unsigned int interpolatedInverseTransformFirstIndex, interpolatedInverseTransformLastIndex;
} MotionSystem;
}   // namespace ocl

static_assert(sizeof(ocl::BVHArrayNode) == 32, "BVHArrayNode must stay 32 bytes");

#define BVHNodeData_IsLeaf(nodeData) ((nodeData) & 0x80000000u)
#define BVHNodeData_GetSkipIndex(nodeData) ((nodeData) & 0x7fffffffu)

typedef struct {
	u_int treeType;
	int costSamples, isectCost, traversalCost;
	float emptyBonus;
} BVHParams;

// One build primitive: a triangle (BVH) or a whole mesh (MBVH root).
struct BVHTreeNode {
	BBox bbox;
	union {
		struct { u_int meshIndex, triangleIndex; } triangleLeaf;
		struct { u_int leafIndex; u_int transformIndex, motionIndex; u_int meshOffsetIndex; bool isMotionMesh; } bvhLeaf;
	};
	BVHTreeNode *leftChild;
	BVHTreeNode *rightSibling;
};

// meshes != NULL: leaves are triangles (vertex indices are copied from the mesh);
// meshes == NULL: leaves carry the bvhLeaf payload (MBVH root tree).
// The caller owns the returned array (delete[]).
extern ocl::BVHArrayNode *BuildBVH(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes,
		std::vector<BVHTreeNode *> &leafList);
extern ocl::BVHArrayNode *BuildEmbreeBVHBinnedSAH(const BVHParams &params, u_int *nNodes,
		const std::deque<const Mesh *> *meshes, std::vector<BVHTreeNode *> &leafList);
extern ocl::BVHArrayNode *BuildEmbreeBVHMorton(const BVHParams &params, u_int *nNodes,
		const std::deque<const Mesh *> *meshes, std::vector<BVHTreeNode *> &leafList);
// extension: builder type "B200_PLOC" -- the GPU builder with a PLOC binary tree (luxcore_b200/csrc/build_kernels.cuh)
extern ocl::BVHArrayNode *BuildB200BVHPloc(const BVHParams &params, u_int *nNodes,
		const std::deque<const Mesh *> *meshes, std::vector<BVHTreeNode *> &leafList);

// extension: the whole single-level accelerator made ON the device (C ABI lrb_bvh_build_scene: triangle boxes, tree, leaf
// payload and the device lay-out, none of it on the host).  quality 0 = EMBREE_MORTON's radix tree, 1 = B200_PLOC.  Fills
// *nodes / *nNodes with the reference array (downloaded once: BVHAccel::bvhTree stays what the reference's friends read) and
// *scene with the resident scene (an lrb_scene of CUDA device *ordinal) that BVHKernel adopts instead of uploading.
// Returns false, touching nothing, when there is no CUDA device (the caller then builds on the host).
extern bool BuildB200SceneOnDevice(const BVHParams &params, const u_int quality, const std::deque<const Mesh *> &meshes,
		ocl::BVHArrayNode **nodes, u_int *nNodes, void **scene, int *ordinal);
extern void FreeB200ResidentScene(void *scene);

}   // namespace luxrays

#endif
