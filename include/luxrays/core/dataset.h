// luxrays/core/dataset.h -- mesh registry + accelerator cache (reference:
// include/luxrays/core/dataset.h:34-91, src/luxrays/core/dataset.cpp:40-176).
#ifndef _LUXRAYS_B200_DATASET_H
#define _LUXRAYS_B200_DATASET_H

#include <map>
#include <mutex>

#include "luxrays/luxrays.h"
#include "luxrays/core/accelerator.h"

namespace luxrays {

class DataSet {
public:
	DataSet(const Context *luxRaysContext);
	~DataSet();

	AcceleratorType GetAcceleratorType() const { return accelType; }
	void SetAcceleratorType(AcceleratorType type) { accelType = type; }

	bool GetInstanceSupport() const { return enableInstanceSupport; }
	bool RequiresInstanceSupport() const { return enableInstanceSupport && hasInstances; }
	bool HasInstances() const { return hasInstances; }
	bool GetMotionBlurSupport() const { return hasMotionBlur; }
	bool RequiresMotionBlurSupport() const { return enableMotionBlurSupport && hasMotionBlur; }
	bool HasMotionBlur() const { return hasMotionBlur; }

	TriangleMeshID Add(const Mesh *mesh);       // returns the meshIndex reported in RayHit
	void Preprocess();
	bool IsPreprocessed() const { return preprocessed; }
	void UpdateBBoxes();

	bool HasAccelerator(const AcceleratorType accelType) const;
	const Accelerator *GetAccelerator(const AcceleratorType accelType);     // built once, cached, thread-safe
	bool DoesAllAcceleratorsSupportUpdate() const;
	void UpdateAccelerators();

	const BBox &GetBBox() const { return bbox; }
	const BSphere &GetBSphere() const { return bsphere; }
	u_longlong GetTotalVertexCount() const { return totalVertexCount; }
	u_longlong GetTotalTriangleCount() const { return totalTriangleCount; }
	u_int GetDataSetID() const { return dataSetID; }
	bool IsEqual(const DataSet *dataSet) const;

	friend class Context;

private:
	u_int dataSetID;
	const Context *context;
	u_longlong totalVertexCount, totalTriangleCount;
	std::deque<const Mesh *> meshes;
	BBox bbox;
	BSphere bsphere;
	mutable std::mutex accelsMutex;
	std::map<AcceleratorType, Accelerator *> accels;
	AcceleratorType accelType;
	bool preprocessed;
	bool hasInstances, enableInstanceSupport;
	bool hasMotionBlur, enableMotionBlurSupport;
};

}   // namespace luxrays

#endif
