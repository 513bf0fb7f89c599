// Forwarder: Min/Max/Clamp/Lerp/RoundUp/WallClockTime (reference: include/luxrays/utils/utils.h).
#ifndef _LUXRAYS_B200_FWD_UTILS_H
#define _LUXRAYS_B200_FWD_UTILS_H
#include "luxrays/luxrays.h"
#endif
