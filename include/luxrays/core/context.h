// luxrays/core/context.h -- device discovery / creation and DataSet lifecycle (reference:
// include/luxrays/core/context.h:48-203, src/luxrays/core/context.cpp:44-356).
// Configuration keys honoured: context.verbose, accelerator.type, accelerator.instances.enable,
// accelerator.motionblur.enable, accelerator.bvh.builder.type, accelerator.bvh.treetype,
// accelerator.bvh.costsamples / isectcost / travcost / emptybonus.
#ifndef _LUXRAYS_B200_CONTEXT_H
#define _LUXRAYS_B200_CONTEXT_H

#include <sstream>

#include "luxrays/luxrays.h"
#include "luxrays/core/dataset.h"
#include "luxrays/utils/properties.h"

namespace luxrays {

typedef void (*LuxRaysDebugHandler)(const char *msg);

#define LR_LOG(c, a) { if (c->HasDebugHandler() && c->IsVerbose()) { std::stringstream _LR_LOG_LOCAL_SS; _LR_LOG_LOCAL_SS << a; c->PrintDebugMsg(_LR_LOG_LOCAL_SS.str().c_str()); } }

class Context {
public:
	Context(LuxRaysDebugHandler handler = NULL, const Properties &config = Properties());
	~Context();

	const Properties &GetConfig() const { return cfg; }

	const std::vector<DeviceDescription *> &GetAvailableDeviceDescriptions() const { return deviceDescriptions; }
	const std::vector<IntersectionDevice *> &GetIntersectionDevices() const { return idevices; }
	const std::vector<HardwareDevice *> &GetHardwareDevices() const { return hdevices; }
	const std::vector<Device *> &GetDevices() const { return devices; }

	std::vector<IntersectionDevice *> AddIntersectionDevices(std::vector<DeviceDescription *> &deviceDescs);
	std::vector<HardwareDevice *> AddHardwareDevices(std::vector<DeviceDescription *> &deviceDescs);

	DataSet *GetCurrentDataSet() const { return currentDataSet; }
	void SetDataSet(DataSet *dataSet);
	void UpdateDataSet();

	bool GetUseOutOfCoreBuffers() const { return useOutOfCoreBuffers; }
	void SetUseOutOfCoreBuffers(const bool v) { useOutOfCoreBuffers = v; }

	void Start();
	void Interrupt();
	void Stop();
	bool IsRunning() const { return started; }

	bool HasDebugHandler() const { return debugHandler != NULL; }
	void PrintDebugMsg(const char *msg) const { if (debugHandler) debugHandler(msg); }
	void SetVerbose(const bool v) { verbose = v; }
	bool IsVerbose() const { return verbose; }

private:
	const Properties cfg;
	LuxRaysDebugHandler debugHandler;
	DataSet *currentDataSet;
	std::vector<DeviceDescription *> deviceDescriptions;
	std::vector<IntersectionDevice *> idevices;
	std::vector<HardwareDevice *> hdevices;
	std::vector<Device *> devices;
	bool started, verbose, useOutOfCoreBuffers;
};

}   // namespace luxrays

#endif
