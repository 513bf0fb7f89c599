// luxrays/devices/cudaintersectiondevice.h -- the B200 device behind the reference's CUDA device
// class names (reference: include/luxrays/devices/cudadevice.h:38-241, cudaintersectiondevice.h:31-60,
// src/luxrays/devices/cudadevice.cpp, cudaintersectiondevice.cpp).
//
// CUDADevice here is a thin C++ wrapper of the C ABI in include/luxrays_b200.h: no cuew, no NVRTC,
// no OptiX.  It registers as DEVICE_TYPE_CUDA_GPU, so Context / SLG code selecting CUDA devices
// picks it up unchanged.
#ifndef _LUXRAYS_B200_CUDAINTERSECTIONDEVICE_H
#define _LUXRAYS_B200_CUDAINTERSECTIONDEVICE_H

#include "luxrays/core/hardwareintersectiondevice.h"

struct lrb_device;
struct lrb_scene;

namespace luxrays {

class CUDADeviceDescription : public DeviceDescription {
public:
	CUDADeviceDescription(const int cudaOrdinal, const std::string &deviceName, const int smCount,
			const size_t totalMem, const int ccMajor, const int ccMinor);
	virtual ~CUDADeviceDescription() { }

	virtual int GetComputeUnits() const { return computeUnits; }
	virtual u_int GetNativeVectorWidthFloat() const { return 1; }
	virtual size_t GetMaxMemory() const { return maxMemory; }
	virtual size_t GetMaxMemoryAllocSize() const { return std::numeric_limits<size_t>::max(); }   // cudadevice.cpp:113-115
	virtual bool HasOutOfCoreMemorySupport() const { return false; }     // 180 GB of HBM3e: not needed

	int GetCUDADeviceIndex() const { return ordinal; }
	int GetCUDAComputeCapabilityMajor() const { return major; }
	int GetCUDAComputeCapabilityMinor() const { return minor; }

	static void AddDeviceDescs(std::vector<DeviceDescription *> &descriptions);

private:
	int ordinal, computeUnits, major, minor;
	size_t maxMemory;
};

class CUDADeviceBuffer : public HardwareDeviceBuffer {
public:
	CUDADeviceBuffer() : ptr(nullptr), size(0) { }
	virtual ~CUDADeviceBuffer() { }
	virtual bool IsNull() const { return ptr == nullptr; }
	virtual size_t GetSize() const { return size; }
	void *GetDevicePointer() const { return ptr; }

	friend class CUDADevice;
private:
	void *ptr;
	size_t size;
};

class CUDADevice : virtual public HardwareDevice {
public:
	CUDADevice(const Context *context, CUDADeviceDescription *desc, const size_t devIndex);
	virtual ~CUDADevice();

	virtual const DeviceDescription *GetDeviceDesc() const { return deviceDesc; }
	virtual void PushThreadCurrentDevice();
	virtual void PopThreadCurrentDevice();

	virtual void CompileProgram(HardwareDeviceProgram **program, const std::vector<std::string> &programParameters,
			const std::string &programSource, const std::string &programName);
	virtual void GetKernel(HardwareDeviceProgram *program, HardwareDeviceKernel **kernel, const std::string &kernelName);
	virtual u_int GetKernelWorkGroupSize(HardwareDeviceKernel *kernel);
	virtual void SetKernelArg(HardwareDeviceKernel *kernel, const u_int index, const size_t size, const void *arg);
	virtual void EnqueueKernel(HardwareDeviceKernel *kernel, const HardwareDeviceRange &globalSize, const HardwareDeviceRange &workGroupSize);

	virtual void EnqueueReadBuffer(const HardwareDeviceBuffer *buff, const bool blocking, const size_t size, void *ptr);
	virtual void EnqueueWriteBuffer(const HardwareDeviceBuffer *buff, const bool blocking, const size_t size, const void *ptr);
	virtual void FlushQueue();
	virtual void FinishQueue();

	virtual void AllocBuffer(HardwareDeviceBuffer **buff, const BufferType type, void *src, const size_t size, const std::string &desc = "");
	virtual void FreeBuffer(HardwareDeviceBuffer **buff);

	// the C-ABI handle
	lrb_device *GetNativeHandle() const { return handle; }
	// Extension: wrap device memory the application already owns on this GPU (e.g. its own ray
	// buffers) so it can be passed to EnqueueTraceRayBuffer without a copy.  The returned wrapper is
	// a view: delete it (not FreeBuffer) when done; the memory stays with its owner.
	HardwareDeviceBuffer *AdoptBuffer(void *devicePointer, const size_t size) const;

	friend class Context;

protected:
	virtual void Start();
	virtual void Stop();

	CUDADeviceDescription *deviceDesc;
	lrb_device *handle;
};

class CUDAIntersectionDevice : public CUDADevice, public HardwareIntersectionDevice {
public:
	CUDAIntersectionDevice(const Context *context, CUDADeviceDescription *desc, const size_t devIndex);
	virtual ~CUDAIntersectionDevice();

	virtual void SetDataSet(DataSet *newDataSet);
	virtual void Start();
	virtual void Stop();

	virtual void EnqueueTraceRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount);
	// Single ray: traced on the GPU as a batch of one (the reference calls the CPU
	// accel->Intersect here; this build has no CPU intersection code).
	virtual bool TraceRay(const Ray *ray, RayHit *rayHit);

	// Extensions (not in the reference, which traces shadow rays as closest-hit and runs the pass-through
	// loop of Scene::Intersect ray by ray, src/slg/scene/scene.cpp:556-690):
	//  * shadow rays: any hit ends the ray; RayHit::Miss() is what EnqueueTraceRayBuffer would report;
	//  * one round of the pass-through loop over a traced batch: rays that hit a mesh flagged in
	//    passMeshBits (bit m = dataset mesh m; HardwareDeviceBuffer of u_int words, may be NULL) or flagged
	//    by the caller in continueFlags (one byte per ray, may be NULL) get  mint = t + MachineEpsilon::E(t)
	//    (scene.cpp:675), all the others RAY_FLAGS_MASKED; returns how many rays continue (blocks).
	virtual void EnqueueTraceShadowRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount);
	virtual unsigned int AdvancePassThroughRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount,
			HardwareDeviceBuffer *passMeshBits, const unsigned int passMeshWords, HardwareDeviceBuffer *continueFlags);

	// the C-ABI scene of the running kernel (nullptr before Start)
	lrb_scene *GetNativeScene() const;

	friend class Context;

protected:
	virtual void Update();

private:
	HardwareIntersectionKernel *kernel;
	HardwareDeviceBuffer *oneRay, *oneHit;
};

}   // namespace luxrays

#endif
