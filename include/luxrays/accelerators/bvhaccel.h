// luxrays/accelerators/bvhaccel.h -- single-level BVH accelerator (reference:
// include/luxrays/accelerators/bvhaccel.h:32-72, src/luxrays/accelerators/bvhaccel.cpp:35-168).
#ifndef _LUXRAYS_B200_BVHACCEL_H
#define _LUXRAYS_B200_BVHACCEL_H

#include "luxrays/core/accelerator.h"
#include "luxrays/core/bvh/bvhbuild.h"

namespace luxrays {

class Properties;

class BVHAccel : public Accelerator {
public:
	BVHAccel(const Context *context);
	virtual ~BVHAccel();

	virtual AcceleratorType GetType() const { return ACCEL_BVH; }
	virtual bool HasNativeSupport(const IntersectionDevice &device) const;
	virtual bool HasHWSupport(const IntersectionDevice &device) const;
	virtual HardwareIntersectionKernel *NewHardwareIntersectionKernel(HardwareIntersectionDevice &device) const;

	virtual void Init(const std::deque<const Mesh *> &meshes, const u_longlong totalVertexCount, const u_longlong totalTriangleCount);
	// throws: no CPU intersection code in this build (accelerator.h)
	virtual bool Intersect(const Ray *ray, RayHit *hit) const;

	static BVHParams ToBVHParams(const Properties &props);

	// read-only views for tests / tools
	u_int GetNodeCount() const { return nNodes; }
	const ocl::BVHArrayNode *GetNodes() const { return bvhTree; }

	friend class MBVHAccel;
	friend class BVHKernel;
	friend class MBVHKernel;

private:
	BVHParams params;
	u_int nNodes;
	ocl::BVHArrayNode *bvhTree;

	const Context *ctx;
	std::deque<const Mesh *> meshes;
	u_longlong totalVertexCount, totalTriangleCount;
	bool initialized;

	// B200 extension (builder types EMBREE_MORTON / B200_PLOC): the accelerator as the device built and laid it out, waiting
	// for the BVHKernel of the same CUDA device to adopt it (then NULL again).  Leaf BVHs of an MBVH never have one.
	bool allowResidentScene;
	mutable void *residentScene;
	mutable int residentOrdinal;
};

}   // namespace luxrays

#endif
