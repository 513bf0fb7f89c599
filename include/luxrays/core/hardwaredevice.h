// luxrays/core/hardwaredevice.h -- buffer / queue abstraction of a GPU device (reference:
// include/luxrays/core/hardwaredevice.h:28-186).  The generic run-time kernel API of the reference
// (CompileProgram / GetKernel / SetKernelArg / EnqueueKernel) serves SLG's own OpenCL-C kernels
// through NVRTC; the B200 device compiles nothing at run time, so those entry points exist and
// throw (they are outside the intersection path, SURVEY.md 8 "out of scope").
#ifndef _LUXRAYS_B200_HARDWAREDEVICE_H
#define _LUXRAYS_B200_HARDWAREDEVICE_H

#include "luxrays/core/device.h"

namespace luxrays {

class HardwareDeviceRange {
public:
	HardwareDeviceRange(const size_t s0) { sizes[0] = s0; sizes[1] = 0; sizes[2] = 0; dimensions = 1; }
	HardwareDeviceRange(const size_t s0, const size_t s1) { sizes[0] = s0; sizes[1] = s1; sizes[2] = 0; dimensions = 2; }
	HardwareDeviceRange(const size_t s0, const size_t s1, const size_t s2) { sizes[0] = s0; sizes[1] = s1; sizes[2] = s2; dimensions = 3; }
	virtual ~HardwareDeviceRange() { }
	size_t sizes[3];
	u_int dimensions;
};

class HardwareDeviceKernel {
public:
	virtual ~HardwareDeviceKernel() { }
	virtual bool IsNull() const = 0;
protected:
	HardwareDeviceKernel() { }
};

class HardwareDeviceProgram {
public:
	virtual ~HardwareDeviceProgram() { }
	virtual bool IsNull() const = 0;
protected:
	HardwareDeviceProgram() { }
};

typedef enum {
	BUFFER_TYPE_NONE = 0,
	BUFFER_TYPE_READ_ONLY = 1 << 0,
	BUFFER_TYPE_READ_WRITE = 1 << 1,
	BUFFER_TYPE_OUT_OF_CORE = 1 << 2
} BufferType;

class HardwareDeviceBuffer {
public:
	virtual ~HardwareDeviceBuffer() { }
	virtual bool IsNull() const = 0;
	virtual size_t GetSize() const = 0;
protected:
	HardwareDeviceBuffer() { }
};

class HardwareDevice : virtual public Device {
public:
	void SetAdditionalCompileOpts(const std::vector<std::string> &opts) { additionalCompileOpts = opts; }
	const std::vector<std::string> &GetAdditionalCompileOpts() { return additionalCompileOpts; }

	virtual void CompileProgram(HardwareDeviceProgram **program, const std::vector<std::string> &programParameters,
			const std::string &programSource, const std::string &programName) = 0;
	virtual void GetKernel(HardwareDeviceProgram *program, HardwareDeviceKernel **kernel, const std::string &kernelName) = 0;
	virtual u_int GetKernelWorkGroupSize(HardwareDeviceKernel *kernel) = 0;
	virtual void SetKernelArg(HardwareDeviceKernel *kernel, const u_int index, const size_t size, const void *arg) = 0;
	virtual void EnqueueKernel(HardwareDeviceKernel *kernel, const HardwareDeviceRange &globalSize,
			const HardwareDeviceRange &workGroupSize) = 0;

	virtual void EnqueueReadBuffer(const HardwareDeviceBuffer *buff, const bool blocking, const size_t size, void *ptr) = 0;
	virtual void EnqueueWriteBuffer(const HardwareDeviceBuffer *buff, const bool blocking, const size_t size, const void *ptr) = 0;
	virtual void FlushQueue() = 0;
	virtual void FinishQueue() = 0;

	size_t GetUsedMemory() const { return usedMemory; }

	// *buff == nullptr: a wrapper is created.  Same size as the current allocation: it is re-used
	// (and re-filled from src when src != nullptr).  size == 0 frees.  (cudadevice.cpp:448-523)
	virtual void AllocBuffer(HardwareDeviceBuffer **buff, const BufferType type, void *src, const size_t size, const std::string &desc = "") = 0;
	virtual void AllocBufferRO(HardwareDeviceBuffer **buff, void *src, const size_t size, const std::string &desc = "") {
		AllocBuffer(buff, BUFFER_TYPE_READ_ONLY, src, size, desc);
	}
	virtual void AllocBufferRW(HardwareDeviceBuffer **buff, void *src, const size_t size, const std::string &desc = "") {
		AllocBuffer(buff, BUFFER_TYPE_READ_WRITE, src, size, desc);
	}
	virtual void FreeBuffer(HardwareDeviceBuffer **buff) = 0;

protected:
	HardwareDevice();
	virtual ~HardwareDevice();

	void AllocMemory(const size_t s) { usedMemory += s; }
	void FreeMemory(const size_t s) { usedMemory -= s; }

	std::vector<std::string> additionalCompileOpts;
	size_t usedMemory;
};

typedef HardwareDeviceBuffer *HardwareDeviceBufferPtr;

}   // namespace luxrays

#endif
