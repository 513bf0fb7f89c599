// luxrays/core/device.h -- DeviceDescription / Device base (reference: include/luxrays/core/device.h:39-138).
#ifndef _LUXRAYS_B200_DEVICE_H
#define _LUXRAYS_B200_DEVICE_H

#include <limits>

#include "luxrays/luxrays.h"

namespace luxrays {

typedef enum {
	DEVICE_TYPE_NATIVE = 1 << 0,
	DEVICE_TYPE_OPENCL_DEFAULT = 1 << 1,
	DEVICE_TYPE_OPENCL_CPU = 1 << 2,
	DEVICE_TYPE_OPENCL_GPU = 1 << 3,
	DEVICE_TYPE_OPENCL_UNKNOWN = 1 << 4,
	DEVICE_TYPE_CUDA_GPU = 1 << 5,
	DEVICE_TYPE_OPENCL_ALL = DEVICE_TYPE_OPENCL_DEFAULT | DEVICE_TYPE_OPENCL_CPU | DEVICE_TYPE_OPENCL_GPU | DEVICE_TYPE_OPENCL_UNKNOWN,
	DEVICE_TYPE_CUDA_ALL = DEVICE_TYPE_CUDA_GPU,
	DEVICE_TYPE_ALL = DEVICE_TYPE_NATIVE | DEVICE_TYPE_OPENCL_ALL,
	DEVICE_TYPE_ALL_HARDWARE = DEVICE_TYPE_OPENCL_ALL | DEVICE_TYPE_CUDA_ALL,
	DEVICE_TYPE_ALL_INTERSECTION = DEVICE_TYPE_NATIVE | DEVICE_TYPE_OPENCL_ALL
} DeviceType;

class DeviceDescription {
public:
	DeviceDescription(const std::string &deviceName, const DeviceType deviceType) :
			name(deviceName), type(deviceType), forceWorkGroupSize(0) { }
	virtual ~DeviceDescription() { }

	const std::string &GetName() const { return name; }
	DeviceType GetType() const { return type; }
	virtual int GetComputeUnits() const { return 1; }
	virtual u_int GetNativeVectorWidthFloat() const { return 4; }
	virtual size_t GetMaxMemory() const { return std::numeric_limits<size_t>::max(); }
	virtual size_t GetMaxMemoryAllocSize() const { return std::numeric_limits<size_t>::max(); }
	virtual bool HasOutOfCoreMemorySupport() const { return false; }
	virtual u_int GetForceWorkGroupSize() const { return forceWorkGroupSize; }
	virtual void SetForceWorkGroupSize(const u_int size) { forceWorkGroupSize = size; }

	static void FilterOne(std::vector<DeviceDescription *> &deviceDescriptions);
	static void Filter(const DeviceType type, std::vector<DeviceDescription *> &deviceDescriptions);
	static std::string GetDeviceType(const DeviceType type);

protected:
	std::string name;
	DeviceType type;
	u_int forceWorkGroupSize;
};

class Device {
public:
	const std::string &GetName() const { return deviceName; }
	const Context *GetContext() const { return deviceContext; }
	virtual const DeviceDescription *GetDeviceDesc() const = 0;
	size_t GetDeviceIndex() const { return deviceIndex; }
	virtual bool IsRunning() const { return started; }

	// make this device the calling thread's current one (cuCtxPushCurrent in the reference,
	// cudaSetDevice behind the C ABI here)
	virtual void PushThreadCurrentDevice() { }
	virtual void PopThreadCurrentDevice() { }

	friend class Context;

protected:
	Device() : deviceContext(nullptr), deviceIndex(0), started(false) { }
	Device(const Context *context, const size_t index);
	virtual ~Device();

	virtual void Start();
	virtual void Interrupt();
	virtual void Stop();

	const Context *deviceContext;
	size_t deviceIndex;
	std::string deviceName;
	bool started;
};

}   // namespace luxrays

#endif
