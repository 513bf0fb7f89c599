// Forwarder: the reference keeps this type in include/luxrays/core/geometry/vector.h; in the B200 host
// layer all geometry value types live in one header.
#ifndef _LUXRAYS_B200_FWD_VECTOR_H
#define _LUXRAYS_B200_FWD_VECTOR_H
#include "luxrays/core/geometry.h"
#endif
