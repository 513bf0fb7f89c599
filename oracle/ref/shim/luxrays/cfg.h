#pragma once
#define LUXRAYS_VERSION_MAJOR "2"
#define LUXRAYS_VERSION_MINOR "5"
