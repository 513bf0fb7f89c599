// Stand-in for include/luxrays/core/context.h (the real Context needs every device class of
// LuxRays).  The accelerators only read their configuration from it and log through it.
#ifndef _LUXRAYS_CONTEXT_H
#define _LUXRAYS_CONTEXT_H
#include <sstream>
#include "luxrays/luxrays.h"
#include "luxrays/core/dataset.h"
#include "luxrays/utils/properties.h"
namespace luxrays {
typedef void (*LuxRaysDebugHandler)(const char *msg);
#define LR_LOG(c, a) { if (c->HasDebugHandler() && c->IsVerbose()) { std::stringstream _LR_LOG_LOCAL_SS; _LR_LOG_LOCAL_SS << a; c->PrintDebugMsg(_LR_LOG_LOCAL_SS.str().c_str()); } }
class Context {
public:
	Context(LuxRaysDebugHandler handler = NULL, const Properties &config = Properties()) : cfg(config) { }
	const Properties &GetConfig() const { return cfg; }
	bool HasDebugHandler() const { return false; }
	bool IsVerbose() const { return false; }
	void PrintDebugMsg(const char *) const { }
private:
	Properties cfg;
};
}
#endif
