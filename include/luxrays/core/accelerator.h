// luxrays/core/accelerator.h -- the Accelerator plugin interface (reference:
// include/luxrays/core/accelerator.h:31-59).
#ifndef _LUXRAYS_B200_ACCELERATOR_H
#define _LUXRAYS_B200_ACCELERATOR_H

#include "luxrays/luxrays.h"
#include "luxrays/core/trianglemesh.h"

namespace luxrays {

typedef enum { ACCEL_AUTO, ACCEL_BVH, ACCEL_MBVH, ACCEL_EMBREE, ACCEL_OPTIX } AcceleratorType;

class Accelerator {
public:
	Accelerator() { }
	virtual ~Accelerator() { }

	virtual AcceleratorType GetType() const = 0;
	virtual bool HasNativeSupport(const IntersectionDevice &device) const = 0;
	virtual bool HasHWSupport(const IntersectionDevice &device) const = 0;
	virtual HardwareIntersectionKernel *NewHardwareIntersectionKernel(HardwareIntersectionDevice &device) const = 0;

	virtual void Init(const std::deque<const Mesh *> &meshes, const u_longlong totalVertexCount, const u_longlong totalTriangleCount) = 0;
	virtual bool DoesSupportUpdate() const { return false; }
	virtual void Update() { throw new std::runtime_error("Internal error in Accelerator::Update()"); }

	// Serial CPU interface of the reference.  The B200 build ships NO CPU intersection code: the
	// accelerators implement this by throwing (see bvhaccel.h); single rays go through
	// CUDAIntersectionDevice::TraceRay, which traces them on the GPU.
	virtual bool Intersect(const Ray *ray, RayHit *hit) const = 0;

	static std::string AcceleratorType2String(const AcceleratorType type);
	static AcceleratorType String2AcceleratorType(const std::string &type);
};

}   // namespace luxrays

#endif
