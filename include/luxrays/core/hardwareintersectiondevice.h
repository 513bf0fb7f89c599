// luxrays/core/hardwareintersectiondevice.h (reference: include/luxrays/core/hardwareintersectiondevice.h:32-72).
#ifndef _LUXRAYS_B200_HARDWAREINTERSECTIONDEVICE_H
#define _LUXRAYS_B200_HARDWAREINTERSECTIONDEVICE_H

#include "luxrays/core/intersectiondevice.h"
#include "luxrays/core/hardwaredevice.h"

namespace luxrays {

class HardwareIntersectionDevice : public IntersectionDevice, virtual public HardwareDevice {
public:
	virtual bool HasHWSupport() const { return true; }

	// data-parallel interface: rayBuff holds rayCount packed Ray, rayHitBuff rayCount packed RayHit
	virtual void EnqueueTraceRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount) {
		throw std::runtime_error("Called EnqueueTraceRayBuffer() on a device without parallel support");
	}

	friend class Context;

protected:
	virtual void Update() = 0;

	HardwareIntersectionDevice();
	virtual ~HardwareIntersectionDevice();
};

class HardwareIntersectionKernel {
public:
	HardwareIntersectionKernel(HardwareIntersectionDevice &dev) : device(dev) { }
	virtual ~HardwareIntersectionKernel() { }

	virtual void Update(const DataSet *newDataSet) = 0;
	virtual void EnqueueTraceRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount) = 0;

protected:
	HardwareIntersectionDevice &device;
};

}   // namespace luxrays

#endif
