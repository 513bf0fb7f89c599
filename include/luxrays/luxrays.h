// luxrays/luxrays.h -- umbrella declarations of the B200 host layer (reference:
// include/luxrays/luxrays.h:24-64, without Boost).
#ifndef _LUXRAYS_B200_LUXRAYS_H
#define _LUXRAYS_B200_LUXRAYS_H

#include <deque>
#include <string>
#include <vector>

#include "luxrays/core/geometry.h"

namespace luxrays {

class Accelerator;
class Context;
class DataSet;
class Device;
class DeviceDescription;
class HardwareDevice;
class HardwareDeviceBuffer;
class HardwareIntersectionDevice;
class HardwareIntersectionKernel;
class IntersectionDevice;
class Mesh;
class TriangleMesh;

typedef u_int TriangleMeshID;
typedef u_int TriangleID;

// src/luxrays/core/init.cpp:45-84 probes cuew/clew; here it only checks that the CUDA library is
// loadable.  Safe to call more than once.
extern void Init();
extern bool isCudaAvilable;     // (sic) spelling of the reference, include/luxrays/utils/cuda.h

double WallClockTime();

}   // namespace luxrays

#endif
