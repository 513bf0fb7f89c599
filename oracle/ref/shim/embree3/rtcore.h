// Declarations only: lets the reference's bvhbuild.h (whose inline templates mention the Embree
// builder API) parse.  The Embree builders are NOT available in oracle/_ref; calling any of these aborts.
#pragma once
#include <cstddef>
#include <cstdlib>
typedef struct RTCDeviceTy *RTCDevice;
typedef struct RTCBVHTy *RTCBVH;
typedef struct RTCThreadLocalAllocatorTy *RTCThreadLocalAllocator;
struct RTCBounds { float lower_x, lower_y, lower_z, align0, upper_x, upper_y, upper_z, align1; };
struct RTCBuildPrimitive { float lower_x, lower_y, lower_z; unsigned int geomID; float upper_x, upper_y, upper_z; unsigned int primID; };
enum RTCBuildQuality { RTC_BUILD_QUALITY_LOW = 0, RTC_BUILD_QUALITY_MEDIUM = 1, RTC_BUILD_QUALITY_HIGH = 2, RTC_BUILD_QUALITY_REFIT = 3 };
enum RTCBuildFlags { RTC_BUILD_FLAG_NONE = 0, RTC_BUILD_FLAG_DYNAMIC = 1 };
typedef void *(*RTCCreateNodeFunction)(RTCThreadLocalAllocator, unsigned int, void *);
typedef void (*RTCSetNodeChildrenFunction)(void *, void **, unsigned int, void *);
typedef void (*RTCSetNodeBoundsFunction)(void *, const RTCBounds **, unsigned int, void *);
typedef void *(*RTCCreateLeafFunction)(RTCThreadLocalAllocator, const RTCBuildPrimitive *, size_t, void *);
typedef void (*RTCSplitPrimitiveFunction)(const RTCBuildPrimitive *, unsigned int, float, RTCBounds *, RTCBounds *, void *);
typedef bool (*RTCProgressMonitorFunction)(void *, double);
struct RTCBuildArguments {
	size_t byteSize; RTCBuildQuality buildQuality; RTCBuildFlags buildFlags; unsigned int maxBranchingFactor, maxDepth, sahBlockSize,
	minLeafSize, maxLeafSize; float traversalCost, intersectionCost; RTCBVH bvh; RTCBuildPrimitive *primitives; size_t primitiveCount,
	primitiveArrayCapacity; RTCCreateNodeFunction createNode; RTCSetNodeChildrenFunction setNodeChildren; RTCSetNodeBoundsFunction setNodeBounds;
	RTCCreateLeafFunction createLeaf; RTCSplitPrimitiveFunction splitPrimitive; RTCProgressMonitorFunction buildProgress; void *userPtr;
};
inline RTCBuildArguments rtcDefaultBuildArguments() { abort(); }
inline RTCDevice rtcNewDevice(const char *) { abort(); }
inline RTCBVH rtcNewBVH(RTCDevice) { abort(); }
inline void rtcReleaseBVH(RTCBVH) { abort(); }
inline void rtcReleaseDevice(RTCDevice) { abort(); }
inline void *rtcThreadLocalAlloc(RTCThreadLocalAllocator, size_t, size_t) { abort(); }
inline void *rtcBuildBVH(const RTCBuildArguments *) { abort(); }
