// devices.cpp -- Device hierarchy + the B200 CUDADevice / CUDAIntersectionDevice on top of the C ABI.
// Reference: src/luxrays/core/{device,intersectiondevice,hardwaredevice,hardwareintersectiondevice}.cpp,
// src/luxrays/devices/cudadevice.cpp:40-542, cudaintersectiondevice.cpp:31-86.
#include <sys/time.h>

#include "luxrays_b200.h"

#include "luxrays/core/context.h"
#include "luxrays/devices/cudaintersectiondevice.h"

namespace luxrays {

bool isCudaAvilable = false;

double WallClockTime() {
	struct timeval t;
	gettimeofday(&t, NULL);
	return t.tv_sec + t.tv_usec / 1000000.0;
}

void Init() {
	int n = 0;
	isCudaAvilable = (lrb_device_count(&n) == LRB_OK) && (n > 0);
}

static void Check(const int rc, const char *what) {
	if (rc != LRB_OK)
		throw std::runtime_error(std::string(what) + ": " + lrb_last_error_string());
}

//------------------------------------------------------------------------------
// DeviceDescription / Device
//------------------------------------------------------------------------------

void DeviceDescription::FilterOne(std::vector<DeviceDescription *> &descs) {
	// keep one device: a hardware one if present, else the first
	if (descs.empty())
		return;
	DeviceDescription *pick = descs[0];
	for (size_t i = 0; i < descs.size(); ++i)
		if (descs[i]->GetType() & DEVICE_TYPE_ALL_HARDWARE) { pick = descs[i]; break; }
	descs.assign(1, pick);
}

void DeviceDescription::Filter(const DeviceType type, std::vector<DeviceDescription *> &descs) {
	std::vector<DeviceDescription *> keep;
	for (size_t i = 0; i < descs.size(); ++i)
		if (descs[i]->GetType() & type)
			keep.push_back(descs[i]);
	descs.swap(keep);
}

std::string DeviceDescription::GetDeviceType(const DeviceType type) {
	switch (type) {
		case DEVICE_TYPE_ALL: return "ALL";
		case DEVICE_TYPE_NATIVE: return "NATIVE";
		case DEVICE_TYPE_OPENCL_ALL: return "OPENCL_ALL";
		case DEVICE_TYPE_OPENCL_DEFAULT: return "OPENCL_DEFAULT";
		case DEVICE_TYPE_OPENCL_CPU: return "OPENCL_CPU";
		case DEVICE_TYPE_OPENCL_GPU: return "OPENCL_GPU";
		case DEVICE_TYPE_OPENCL_UNKNOWN: return "OPENCL_UNKNOWN";
		case DEVICE_TYPE_CUDA_GPU: return "CUDA_GPU";
		default: return "UNKNOWN";
	}
}

Device::Device(const Context *context, const size_t index) : deviceContext(context), deviceIndex(index), started(false) { }
Device::~Device() { }
void Device::Start() { started = true; }
void Device::Interrupt() { }
void Device::Stop() { started = false; }

IntersectionDevice::IntersectionDevice() : dataSet(nullptr), accel(nullptr), statsStartTime(0.0),
		statsTotalSerialRayCount(0), statsTotalDataParallelRayCount(0) { }
IntersectionDevice::~IntersectionDevice() { }
void IntersectionDevice::SetDataSet(DataSet *newDataSet) { dataSet = newDataSet; }
void IntersectionDevice::Start() {
	Device::Start();
	statsStartTime = WallClockTime();
	statsTotalSerialRayCount = 0;
	statsTotalDataParallelRayCount = 0;
}

HardwareDevice::HardwareDevice() : usedMemory(0) { }
HardwareDevice::~HardwareDevice() {
	if (usedMemory != 0 && deviceContext)
		LR_LOG(deviceContext, "WARNING: there is a memory leak in LuxRays HardwareDevice " << deviceName << ": " << usedMemory << "bytes");
}

HardwareIntersectionDevice::HardwareIntersectionDevice() { }
HardwareIntersectionDevice::~HardwareIntersectionDevice() { }

//------------------------------------------------------------------------------
// CUDADeviceDescription
//------------------------------------------------------------------------------

CUDADeviceDescription::CUDADeviceDescription(const int cudaOrdinal, const std::string &deviceName, const int smCount,
		const size_t totalMem, const int ccMajor, const int ccMinor) :
		DeviceDescription(deviceName, DEVICE_TYPE_CUDA_GPU), ordinal(cudaOrdinal), computeUnits(smCount),
		major(ccMajor), minor(ccMinor), maxMemory(totalMem) { }

void CUDADeviceDescription::AddDeviceDescs(std::vector<DeviceDescription *> &descriptions) {
	int n = 0;
	if (lrb_device_count(&n) != LRB_OK)
		return;
	for (int i = 0; i < n; ++i) {
		lrb_device *d = nullptr;
		if (lrb_device_create(i, &d) != LRB_OK)
			continue;
		lrb_device_props p;
		if (lrb_device_get_props(d, &p) == LRB_OK)
			descriptions.push_back(new CUDADeviceDescription(i, p.name, p.sm_count, (size_t)p.total_mem_bytes, p.cc_major, p.cc_minor));
		lrb_device_destroy(d);
	}
}

//------------------------------------------------------------------------------
// CUDADevice
//------------------------------------------------------------------------------

CUDADevice::CUDADevice(const Context *context, CUDADeviceDescription *desc, const size_t devIndex) :
		Device(context, devIndex), deviceDesc(desc), handle(nullptr) {
	deviceName = (desc->GetName() + " CUDAIntersect").c_str();
	// the reference creates its CUDA context in the constructor; the C-ABI device is the equivalent
	Check(lrb_device_create(desc->GetCUDADeviceIndex(), &handle), "CUDADevice");
}

CUDADevice::~CUDADevice() {
	if (started)
		CUDADevice::Stop();
	lrb_device_destroy(handle);
	handle = nullptr;
}

// The runtime API binds the device per call inside the C ABI (cudaSetDevice), so there is no
// context stack to maintain; the calls are kept because callers bracket their work with them.
void CUDADevice::PushThreadCurrentDevice() { }
void CUDADevice::PopThreadCurrentDevice() { }

void CUDADevice::Start() { HardwareDevice::Start(); }
void CUDADevice::Stop() {
	if (handle)
		lrb_sync(handle);
	HardwareDevice::Stop();
}

static void NoRuntimeKernels() {
	throw std::runtime_error("The B200 device does not compile kernels at run time (no NVRTC/OpenCL-C path); "
			"only the intersection kernels built into libluxrays_b200.so are available");
}
void CUDADevice::CompileProgram(HardwareDeviceProgram **, const std::vector<std::string> &, const std::string &, const std::string &) { NoRuntimeKernels(); }
void CUDADevice::GetKernel(HardwareDeviceProgram *, HardwareDeviceKernel **, const std::string &) { NoRuntimeKernels(); }
u_int CUDADevice::GetKernelWorkGroupSize(HardwareDeviceKernel *) { NoRuntimeKernels(); return 0; }
void CUDADevice::SetKernelArg(HardwareDeviceKernel *, const u_int, const size_t, const void *) { NoRuntimeKernels(); }
void CUDADevice::EnqueueKernel(HardwareDeviceKernel *, const HardwareDeviceRange &, const HardwareDeviceRange &) { NoRuntimeKernels(); }

void CUDADevice::EnqueueReadBuffer(const HardwareDeviceBuffer *buff, const bool blocking, const size_t size, void *ptr) {
	const CUDADeviceBuffer *cb = dynamic_cast<const CUDADeviceBuffer *>(buff);
	if (!cb || cb->IsNull())
		throw std::runtime_error("Null buffer in CUDADevice::EnqueueReadBuffer()");
	if (size > cb->GetSize())
		throw std::runtime_error("Read past the end of the buffer in CUDADevice::EnqueueReadBuffer()");
	Check(lrb_d2h(handle, ptr, cb->GetDevicePointer(), size, blocking ? 1 : 0), "EnqueueReadBuffer");
}

void CUDADevice::EnqueueWriteBuffer(const HardwareDeviceBuffer *buff, const bool blocking, const size_t size, const void *ptr) {
	const CUDADeviceBuffer *cb = dynamic_cast<const CUDADeviceBuffer *>(buff);
	if (!cb || cb->IsNull())
		throw std::runtime_error("Null buffer in CUDADevice::EnqueueWriteBuffer()");
	if (size > cb->GetSize())
		throw std::runtime_error("Write past the end of the buffer in CUDADevice::EnqueueWriteBuffer()");
	Check(lrb_h2d(handle, cb->GetDevicePointer(), ptr, size, blocking ? 1 : 0), "EnqueueWriteBuffer");
}

void CUDADevice::FlushQueue() { Check(lrb_flush(handle), "FlushQueue"); }
void CUDADevice::FinishQueue() { Check(lrb_sync(handle), "FinishQueue"); }

void CUDADevice::AllocBuffer(HardwareDeviceBuffer **buff, const BufferType, void *src, const size_t size, const std::string &) {
	if (!*buff)
		*buff = new CUDADeviceBuffer();
	CUDADeviceBuffer *cb = dynamic_cast<CUDADeviceBuffer *>(*buff);
	if (!cb)
		throw std::runtime_error("Foreign buffer passed to CUDADevice::AllocBuffer()");

	if (size == 0) {
		// free, keep the (null) wrapper
		if (cb->ptr) {
			FreeMemory(cb->size);
			Check(lrb_free(handle, cb->ptr), "AllocBuffer/free");
			cb->ptr = nullptr;
			cb->size = 0;
		}
		return;
	}
	if (cb->ptr && cb->size != size) {
		FreeMemory(cb->size);
		Check(lrb_free(handle, cb->ptr), "AllocBuffer/realloc");
		cb->ptr = nullptr;
		cb->size = 0;
	}
	if (!cb->ptr) {
		Check(lrb_alloc(handle, size, &cb->ptr), "AllocBuffer");
		cb->size = size;
		AllocMemory(size);
	}
	if (src)
		Check(lrb_h2d(handle, cb->ptr, src, size, 0), "AllocBuffer/upload");
}

void CUDADevice::FreeBuffer(HardwareDeviceBuffer **buff) {
	if (!*buff)
		return;
	CUDADeviceBuffer *cb = dynamic_cast<CUDADeviceBuffer *>(*buff);
	if (!cb)
		throw std::runtime_error("Foreign buffer passed to CUDADevice::FreeBuffer()");
	if (cb->ptr) {
		FreeMemory(cb->size);
		Check(lrb_free(handle, cb->ptr), "FreeBuffer");
	}
	delete *buff;
	*buff = nullptr;
}

HardwareDeviceBuffer *CUDADevice::AdoptBuffer(void *devicePointer, const size_t size) const {
	CUDADeviceBuffer *b = new CUDADeviceBuffer();
	b->ptr = devicePointer;
	b->size = size;
	return b;
}

//------------------------------------------------------------------------------
// CUDAIntersectionDevice
//------------------------------------------------------------------------------

CUDAIntersectionDevice::CUDAIntersectionDevice(const Context *context, CUDADeviceDescription *desc, const size_t devIndex) :
		Device(context, devIndex), CUDADevice(context, desc, devIndex), HardwareIntersectionDevice(),
		kernel(nullptr), oneRay(nullptr), oneHit(nullptr) { }

CUDAIntersectionDevice::~CUDAIntersectionDevice() {
	if (started)
		CUDAIntersectionDevice::Stop();
}

// accelerator choice: explicit accelerator.type, else MBVH when instances / motion blur must be
// honoured, else BVH.  (The reference prefers OPTIX when an RT-core context exists; B200 has none.)
void CUDAIntersectionDevice::SetDataSet(DataSet *newDataSet) {
	IntersectionDevice::SetDataSet(newDataSet);
	if (!dataSet)
		return;
	AcceleratorType t = dataSet->GetAcceleratorType();
	if (t == ACCEL_AUTO || t == ACCEL_OPTIX || t == ACCEL_EMBREE)
		t = (dataSet->RequiresInstanceSupport() || dataSet->RequiresMotionBlurSupport()) ? ACCEL_MBVH : ACCEL_BVH;
	accel = dataSet->GetAccelerator(t);
}

extern lrb_scene *NativeSceneOf(HardwareIntersectionKernel *k);

lrb_scene *CUDAIntersectionDevice::GetNativeScene() const {
	return kernel ? NativeSceneOf(kernel) : nullptr;
}

void CUDAIntersectionDevice::Update() {
	kernel->Update(dataSet);
}

void CUDAIntersectionDevice::Start() {
	IntersectionDevice::Start();
	CUDADevice::Start();
	if (!accel)
		throw std::runtime_error("CUDAIntersectionDevice::Start() without a DataSet");
	kernel = accel->NewHardwareIntersectionKernel(*this);
}

void CUDAIntersectionDevice::Stop() {
	FreeBuffer(&oneRay);
	FreeBuffer(&oneHit);
	delete kernel;
	kernel = nullptr;
	CUDADevice::Stop();
}

void CUDAIntersectionDevice::EnqueueTraceRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount) {
	if (!kernel)
		throw std::runtime_error("EnqueueTraceRayBuffer() on a device that was not started");
	kernel->EnqueueTraceRayBuffer(rayBuff, rayHitBuff, rayCount);
	statsTotalDataParallelRayCount += rayCount;
}

static void *DevicePointerOf(HardwareDeviceBuffer *b, const char *what, const bool optional = false) {
	if (!b && optional)
		return nullptr;
	CUDADeviceBuffer *cb = dynamic_cast<CUDADeviceBuffer *>(b);
	if (!cb || cb->IsNull())
		throw std::runtime_error(std::string("Null or foreign buffer passed as ") + what);
	return cb->GetDevicePointer();
}

void CUDAIntersectionDevice::EnqueueTraceShadowRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount) {
	if (!kernel)
		throw std::runtime_error("EnqueueTraceShadowRayBuffer() on a device that was not started");
	if (rayCount == 0)
		return;
	Check(lrb_trace_anyhit(GetNativeScene(), DevicePointerOf(rayBuff, "ray buffer"), DevicePointerOf(rayHitBuff, "ray hit buffer"), rayCount),
			"shadow ray trace");
	statsTotalDataParallelRayCount += rayCount;
}

unsigned int CUDAIntersectionDevice::AdvancePassThroughRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff,
		const unsigned int rayCount, HardwareDeviceBuffer *passMeshBits, const unsigned int passMeshWords, HardwareDeviceBuffer *continueFlags) {
	if (!kernel)
		throw std::runtime_error("AdvancePassThroughRayBuffer() on a device that was not started");
	uint32_t n = 0;
	Check(lrb_advance_rays(GetNativeScene(), DevicePointerOf(rayBuff, "ray buffer"), DevicePointerOf(rayHitBuff, "ray hit buffer"), rayCount,
			(const uint32_t *)DevicePointerOf(passMeshBits, "pass-through mesh bits", true), passMeshWords,
			(const uint8_t *)DevicePointerOf(continueFlags, "continue flags", true), &n), "pass-through advance");
	return n;
}

bool CUDAIntersectionDevice::TraceRay(const Ray *ray, RayHit *rayHit) {
	if (!kernel)
		throw std::runtime_error("TraceRay() on a device that was not started");
	statsTotalSerialRayCount += 1;
	Ray r = *ray;
	r.flags = RAY_FLAGS_NONE;
	AllocBufferRW(&oneRay, &r, sizeof(Ray), "TraceRay ray");
	AllocBufferRW(&oneHit, nullptr, sizeof(RayHit), "TraceRay hit");
	kernel->EnqueueTraceRayBuffer(oneRay, oneHit, 1);
	EnqueueReadBuffer(oneHit, true, sizeof(RayHit), rayHit);
	return !rayHit->Miss();
}

}   // namespace luxrays
