// Forwarder: MachineEpsilon (reference: include/luxrays/core/epsilon.h) lives in geometry.h here.
#ifndef _LUXRAYS_B200_FWD_EPSILON_H
#define _LUXRAYS_B200_FWD_EPSILON_H
#include "luxrays/core/geometry.h"
#endif
