// luxrays/accelerators/mbvhaccel.h -- two-level BVH for instanced / motion-blurred meshes
// (reference: include/luxrays/accelerators/mbvhaccel.h:33-82, src/luxrays/accelerators/mbvhaccel.cpp:40-250).
#ifndef _LUXRAYS_B200_MBVHACCEL_H
#define _LUXRAYS_B200_MBVHACCEL_H

#include "luxrays/accelerators/bvhaccel.h"

namespace luxrays {

class MBVHAccel : public Accelerator {
public:
	MBVHAccel(const Context *context);
	virtual ~MBVHAccel();

	virtual AcceleratorType GetType() const { return ACCEL_MBVH; }
	virtual bool HasNativeSupport(const IntersectionDevice &device) const;
	virtual bool HasHWSupport(const IntersectionDevice &device) const;
	virtual HardwareIntersectionKernel *NewHardwareIntersectionKernel(HardwareIntersectionDevice &device) const;

	virtual void Init(const std::deque<const Mesh *> &meshes, const u_longlong totalVertexCount, const u_longlong totalTriangleCount);
	virtual bool DoesSupportUpdate() const { return true; }
	virtual void Update();
	// throws: no CPU intersection code in this build (accelerator.h)
	virtual bool Intersect(const Ray *ray, RayHit *hit) const;

	u_int GetRootNodeCount() const { return nRootNodes; }
	const ocl::BVHArrayNode *GetRootNodes() const { return bvhRootTree; }
	size_t GetUniqueLeafCount() const { return uniqueLeafs.size(); }
	const BVHAccel *GetUniqueLeaf(size_t i) const { return uniqueLeafs[i]; }

	friend class MBVHKernel;

private:
	void UpdateRootBVH();

	BVHParams params;

	// per-mesh root primitives (kept so Update() can rebuild the root tree only)
	std::vector<BVHTreeNode> bvhLeafs;
	std::vector<BVHTreeNode *> bvhLeafsList;

	u_int nRootNodes;
	ocl::BVHArrayNode *bvhRootTree;

	std::vector<BVHAccel *> uniqueLeafs;
	std::vector<const Transform *> uniqueLeafsTransform;       // live pointers: Update() sees edits
	std::vector<const MotionSystem *> uniqueLeafsMotionSystem;

	const Context *ctx;
	std::deque<const Mesh *> meshes;
	bool initialized;
};

}   // namespace luxrays

#endif
