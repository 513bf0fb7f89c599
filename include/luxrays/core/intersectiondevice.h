// luxrays/core/intersectiondevice.h (reference: include/luxrays/core/intersectiondevice.h:31-83).
#ifndef _LUXRAYS_B200_INTERSECTIONDEVICE_H
#define _LUXRAYS_B200_INTERSECTIONDEVICE_H

#include "luxrays/core/device.h"
#include "luxrays/core/accelerator.h"

namespace luxrays {

class IntersectionDevice : virtual public Device {
public:
	virtual bool HasHWSupport() const { return false; }
	const Accelerator *GetAccelerator() const { return accel; }

	// statistics: rays / wall time since Start (masked rays are counted, like the reference)
	virtual double GetTotalRaysCount() const { return (double)(statsTotalSerialRayCount + statsTotalDataParallelRayCount); }
	virtual double GetTotalPerformance() const {
		const double dt = WallClockTime() - statsStartTime;
		return (dt == 0.0) ? 1.0 : ((statsTotalSerialRayCount + statsTotalDataParallelRayCount) / dt);
	}
	virtual u_longlong GetSerialPerformance() const {
		const double dt = WallClockTime() - statsStartTime;
		return (dt == 0.0) ? 1 : (u_longlong)(statsTotalSerialRayCount / dt);
	}
	virtual u_longlong GetDataParallelPerformance() const {
		const double dt = WallClockTime() - statsStartTime;
		return (dt == 0.0) ? 1 : (u_longlong)(statsTotalDataParallelRayCount / dt);
	}
	virtual void ResetPerformaceStats() {
		statsStartTime = WallClockTime();
		statsTotalSerialRayCount = 0;
		statsTotalDataParallelRayCount = 0;
	}

	// serial interface (one ray)
	virtual bool TraceRay(const Ray *ray, RayHit *rayHit) {
		statsTotalSerialRayCount += 1;
		return accel->Intersect(ray, rayHit);
	}

	friend class Context;

protected:
	IntersectionDevice();
	virtual ~IntersectionDevice();

	virtual void SetDataSet(DataSet *newDataSet);
	virtual void Start();

	DataSet *dataSet;
	const Accelerator *accel;
	double statsStartTime;
	u_longlong statsTotalSerialRayCount, statsTotalDataParallelRayCount;
};

}   // namespace luxrays

#endif
