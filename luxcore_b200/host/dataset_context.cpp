// dataset_context.cpp -- DataSet and Context (reference: src/luxrays/core/dataset.cpp:40-176,
// src/luxrays/core/context.cpp:44-356).  Device list: only the B200 CUDA devices are created by
// this library (native / OpenCL devices belong to the reference build).
#include <atomic>

#include "luxrays/accelerators/bvhaccel.h"
#include "luxrays/accelerators/mbvhaccel.h"
#include "luxrays/core/context.h"
#include "luxrays/devices/cudaintersectiondevice.h"

namespace luxrays {

static std::atomic<u_int> nextDataSetID(0);

DataSet::DataSet(const Context *luxRaysContext) : dataSetID(nextDataSetID++), context(luxRaysContext),
		totalVertexCount(0), totalTriangleCount(0), preprocessed(false), hasInstances(false), hasMotionBlur(false) {
	const Properties &cfg = luxRaysContext->GetConfig();
	accelType = Accelerator::String2AcceleratorType(cfg.Get(Property("accelerator.type")("AUTO")).Get<std::string>());
	enableInstanceSupport = cfg.Get(Property("accelerator.instances.enable")(true)).Get<bool>();
	enableMotionBlurSupport = cfg.Get(Property("accelerator.motionblur.enable")(true)).Get<bool>();
}

DataSet::~DataSet() {
	for (std::map<AcceleratorType, Accelerator *>::iterator it = accels.begin(); it != accels.end(); ++it)
		delete it->second;
}

TriangleMeshID DataSet::Add(const Mesh *mesh) {
	const TriangleMeshID id = (TriangleMeshID)meshes.size();
	meshes.push_back(mesh);
	totalVertexCount += mesh->GetTotalVertexCount();
	totalTriangleCount += mesh->GetTotalTriangleCount();
	const MeshType t = mesh->GetType();
	if (t == TYPE_TRIANGLE_INSTANCE || t == TYPE_EXT_TRIANGLE_INSTANCE)
		hasInstances = true;
	else if (t == TYPE_TRIANGLE_MOTION || t == TYPE_EXT_TRIANGLE_MOTION)
		hasMotionBlur = true;
	return id;
}

void DataSet::Preprocess() {
	LR_LOG(context, "Preprocessing DataSet");
	LR_LOG(context, "Total vertex count: " << totalVertexCount);
	LR_LOG(context, "Total triangle count: " << totalTriangleCount);
	UpdateBBoxes();
	preprocessed = true;
}

void DataSet::UpdateBBoxes() {
	if (totalTriangleCount == 0)
		bbox = Union(Union(bbox, Point(-1.f, -1.f, -1.f)), Point(1.f, 1.f, 1.f));
	else
		for (size_t i = 0; i < meshes.size(); ++i)
			bbox = Union(bbox, meshes[i]->GetBBox());
	bsphere = bbox.BoundingSphere();    // dataset.cpp:104
}

bool DataSet::HasAccelerator(const AcceleratorType t) const {
	std::lock_guard<std::mutex> lock(accelsMutex);
	return accels.find(t) != accels.end();
}

const Accelerator *DataSet::GetAccelerator(const AcceleratorType t) {
	std::lock_guard<std::mutex> lock(accelsMutex);
	std::map<AcceleratorType, Accelerator *>::iterator it = accels.find(t);
	if (it != accels.end())
		return it->second;

	LR_LOG(context, "Adding DataSet accelerator: " << Accelerator::AcceleratorType2String(t));
	Accelerator *accel;
	switch (t) {
		case ACCEL_BVH: accel = new BVHAccel(context); break;
		case ACCEL_MBVH: accel = new MBVHAccel(context); break;
		case ACCEL_EMBREE:
			throw std::runtime_error("EMBREE is a CPU accelerator of the reference build; the B200 library provides BVH and MBVH");
		case ACCEL_OPTIX:
			throw std::runtime_error("OPTIX needs RT cores; the B200 library provides BVH and MBVH");
		default:
			throw std::runtime_error("Unknown AcceleratorType in DataSet::AddAccelerator()");
	}
	try {
		accel->Init(meshes, totalVertexCount, totalTriangleCount);
	} catch (...) {
		delete accel;
		throw;
	}
	accels[t] = accel;
	return accel;
}

bool DataSet::DoesAllAcceleratorsSupportUpdate() const {
	std::lock_guard<std::mutex> lock(accelsMutex);
	for (std::map<AcceleratorType, Accelerator *>::const_iterator it = accels.begin(); it != accels.end(); ++it)
		if (!it->second->DoesSupportUpdate())
			return false;
	return true;
}

void DataSet::UpdateAccelerators() {
	std::lock_guard<std::mutex> lock(accelsMutex);
	for (std::map<AcceleratorType, Accelerator *>::iterator it = accels.begin(); it != accels.end(); ++it) {
		if (!it->second->DoesSupportUpdate())
			throw std::runtime_error("DataSet::UpdateAccelerators(): accelerator " +
					Accelerator::AcceleratorType2String(it->first) + " does not support Update()");
		it->second->Update();
	}
}

bool DataSet::IsEqual(const DataSet *dataSet) const {
	return (dataSet != NULL) && (dataSetID == dataSet->dataSetID);
}

//------------------------------------------------------------------------------
// Context
//------------------------------------------------------------------------------

Context::Context(LuxRaysDebugHandler handler, const Properties &config) : cfg(config), debugHandler(handler),
		currentDataSet(nullptr), started(false), useOutOfCoreBuffers(false) {
	verbose = cfg.Get(Property("context.verbose")(true)).Get<bool>();
	Init();
	LR_LOG(this, "CUDA support: " << (isCudaAvilable ? "available" : "not available"));
	if (isCudaAvilable)
		CUDADeviceDescription::AddDeviceDescs(deviceDescriptions);
	for (size_t i = 0; i < deviceDescriptions.size(); ++i) {
		const DeviceDescription *d = deviceDescriptions[i];
		LR_LOG(this, "Device " << i << " name: " << d->GetName());
		LR_LOG(this, "Device " << i << " type: " << DeviceDescription::GetDeviceType(d->GetType()));
		LR_LOG(this, "Device " << i << " compute units: " << d->GetComputeUnits());
		LR_LOG(this, "Device " << i << " max allocable memory: " << d->GetMaxMemory() / (1024 * 1024) << "MBytes");
	}
}

Context::~Context() {
	if (started)
		Stop();
	for (size_t i = 0; i < devices.size(); ++i)
		delete devices[i];
	for (size_t i = 0; i < deviceDescriptions.size(); ++i)
		delete deviceDescriptions[i];
}

void Context::SetDataSet(DataSet *dataSet) {
	if (started)
		throw std::runtime_error("Context::SetDataSet() while the context is running");
	currentDataSet = dataSet;
	for (size_t i = 0; i < idevices.size(); ++i)
		idevices[i]->SetDataSet(currentDataSet);
}

void Context::UpdateDataSet() {
	if (!started)
		throw std::runtime_error("Context::UpdateDataSet() while the context is stopped");
	currentDataSet->UpdateAccelerators();
	for (size_t i = 0; i < idevices.size(); ++i) {
		HardwareIntersectionDevice *hd = dynamic_cast<HardwareIntersectionDevice *>(idevices[i]);
		if (hd)
			hd->Update();
	}
}

void Context::Start() {
	if (started)
		throw std::runtime_error("Context::Start() called twice");
	for (size_t i = 0; i < devices.size(); ++i) {
		devices[i]->PushThreadCurrentDevice();
		devices[i]->Start();
		devices[i]->PopThreadCurrentDevice();
	}
	started = true;
}

void Context::Interrupt() {
	for (size_t i = 0; i < devices.size(); ++i) {
		devices[i]->PushThreadCurrentDevice();
		devices[i]->Interrupt();
		devices[i]->PopThreadCurrentDevice();
	}
}

void Context::Stop() {
	if (!started)
		throw std::runtime_error("Context::Stop() on a stopped context");
	Interrupt();
	for (size_t i = 0; i < devices.size(); ++i) {
		devices[i]->PushThreadCurrentDevice();
		devices[i]->Stop();
		devices[i]->PopThreadCurrentDevice();
	}
	started = false;
}

std::vector<IntersectionDevice *> Context::AddIntersectionDevices(std::vector<DeviceDescription *> &descs) {
	if (started)
		throw std::runtime_error("Context::AddIntersectionDevices() while the context is running");
	LR_LOG(this, "Creating " << descs.size() << " intersection device(s)");
	std::vector<IntersectionDevice *> created;
	for (size_t i = 0; i < descs.size(); ++i) {
		if (!(descs[i]->GetType() & DEVICE_TYPE_CUDA_ALL))
			throw std::runtime_error("Unknown device type in Context::CreateIntersectionDevices(): " + std::to_string(descs[i]->GetType()));
		CUDAIntersectionDevice *d = new CUDAIntersectionDevice(this, static_cast<CUDADeviceDescription *>(descs[i]), idevices.size());
		idevices.push_back(d);
		devices.push_back(d);
		created.push_back(d);
	}
	return created;
}

std::vector<HardwareDevice *> Context::AddHardwareDevices(std::vector<DeviceDescription *> &descs) {
	if (started)
		throw std::runtime_error("Context::AddHardwareDevices() while the context is running");
	std::vector<HardwareDevice *> created;
	for (size_t i = 0; i < descs.size(); ++i) {
		if (!(descs[i]->GetType() & DEVICE_TYPE_CUDA_ALL))
			throw std::runtime_error("Unknown device type in Context::CreateHardwareDevices(): " + std::to_string(descs[i]->GetType()));
		CUDADevice *d = new CUDADevice(this, static_cast<CUDADeviceDescription *>(descs[i]), hdevices.size());
		hdevices.push_back(d);
		devices.push_back(d);
		created.push_back(d);
	}
	return created;
}

}   // namespace luxrays
