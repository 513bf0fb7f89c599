// Forwarder: CUDADevice / CUDADeviceDescription (reference: include/luxrays/devices/cudadevice.h).
#ifndef _LUXRAYS_B200_FWD_CUDADEVICE_H
#define _LUXRAYS_B200_FWD_CUDADEVICE_H
#include "luxrays/devices/cudaintersectiondevice.h"
#endif
