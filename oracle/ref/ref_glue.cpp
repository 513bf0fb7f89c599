// ref_glue.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only; never shipped, never linked into the product).
//
// The reference's accelerators (bvhaccel.cpp, mbvhaccel.cpp) read a handful of configuration values
// through luxrays::Properties and declare three virtuals that live in the OpenCL/CUDA host files.
// properties.cpp needs the real Boost (regex, string algorithms, base64), so the few members the
// accelerators touch are defined here instead -- configuration plumbing only, no arithmetic of the
// intersection path.  Everything that computes (Triangle::Intersect, BBox::IntersectP, the CLASSIC
// builder, BVHAccel/MBVHAccel::Init/Intersect, Transform, Matrix4x4::Inverse, MotionSystem, Quaternion,
// MachineEpsilon, the mesh bounding boxes) is compiled from /root/reference unchanged.
#include <locale>
#include <stdexcept>
#include <string>

#include "luxrays/luxrays.h"
#include "luxrays/utils/properties.h"
#include "luxrays/utils/proputils.h"
#include "luxrays/core/bvh/bvhbuild.h"
#include "luxrays/accelerators/bvhaccel.h"
#include "luxrays/accelerators/mbvhaccel.h"

namespace luxrays {

std::locale cLocale("C");

// ---- PropertyValue / Property / Properties: store and return ---------------------------------

PropertyValue::PropertyValue() : dataType(NONE_VAL) { }
PropertyValue::PropertyValue(const PropertyValue &o) : dataType(NONE_VAL) { Copy(o, *this); }
PropertyValue::PropertyValue(const bool v) : dataType(BOOL_VAL) { data.boolVal = v; }
PropertyValue::PropertyValue(const int v) : dataType(INT_VAL) { data.intVal = v; }
PropertyValue::PropertyValue(const unsigned int v) : dataType(UINT_VAL) { data.uintVal = v; }
PropertyValue::PropertyValue(const float v) : dataType(FLOAT_VAL) { data.floatVal = v; }
PropertyValue::PropertyValue(const double v) : dataType(DOUBLE_VAL) { data.doubleVal = v; }
PropertyValue::PropertyValue(const long long v) : dataType(LONGLONG_VAL) { data.longlongVal = v; }
PropertyValue::PropertyValue(const unsigned long long v) : dataType(ULONGLONG_VAL) { data.ulonglongVal = v; }
PropertyValue::PropertyValue(const std::string &v) : dataType(STRING_VAL) { data.stringVal = new std::string(v); }
PropertyValue::~PropertyValue() {
	if (dataType == STRING_VAL)
		delete data.stringVal;
}
void PropertyValue::Copy(const PropertyValue &a, PropertyValue &b) {
	if (b.dataType == STRING_VAL)
		delete b.data.stringVal;
	b.dataType = a.dataType;
	if (a.dataType == STRING_VAL)
		b.data.stringVal = new std::string(*a.data.stringVal);
	else if (a.dataType == BLOB_VAL)
		throw std::runtime_error("Blob properties are not supported by the oracle/_ref glue");
	else
		b.data = a.data;
}
PropertyValue &PropertyValue::operator=(const PropertyValue &o) {
	if (this != &o)
		Copy(o, *this);
	return *this;
}
PropertyValue::DataType PropertyValue::GetValueType() const { return dataType; }

template <class T> static T Numeric(const PropertyValue::DataType t, const void *dp) {
	typedef union { bool b; int i; unsigned int u; float f; double d; long long l; unsigned long long ul; std::string *s; } U;
	const U &d = *static_cast<const U *>(dp);
	switch (t) {
		case PropertyValue::BOOL_VAL: return (T)d.b;
		case PropertyValue::INT_VAL: return (T)d.i;
		case PropertyValue::UINT_VAL: return (T)d.u;
		case PropertyValue::FLOAT_VAL: return (T)d.f;
		case PropertyValue::DOUBLE_VAL: return (T)d.d;
		case PropertyValue::LONGLONG_VAL: return (T)d.l;
		case PropertyValue::ULONGLONG_VAL: return (T)d.ul;
		case PropertyValue::STRING_VAL: return boost::lexical_cast<T>(*d.s);
		default: throw std::runtime_error("empty property value");
	}
}
template<> bool PropertyValue::Get<bool>() const { return Numeric<int>(dataType, &data) != 0; }
template<> int PropertyValue::Get<int>() const { return Numeric<int>(dataType, &data); }
template<> unsigned int PropertyValue::Get<unsigned int>() const { return Numeric<unsigned int>(dataType, &data); }
template<> float PropertyValue::Get<float>() const { return Numeric<float>(dataType, &data); }
template<> double PropertyValue::Get<double>() const { return Numeric<double>(dataType, &data); }
template<> std::string PropertyValue::Get<std::string>() const {
	if (dataType == STRING_VAL)
		return *data.stringVal;
	return std::to_string(Numeric<double>(dataType, &data));
}

Property::Property() : name("") { }
Property::Property(const std::string &propName) : name(propName) { }
Property::~Property() { }
template<> bool Property::Get<bool>() const { return Get<bool>(0); }
template<> int Property::Get<int>() const { return Get<int>(0); }
template<> unsigned int Property::Get<unsigned int>() const { return Get<unsigned int>(0); }
template<> float Property::Get<float>() const { return Get<float>(0); }
template<> double Property::Get<double>() const { return Get<double>(0); }
template<> std::string Property::Get<std::string>() const { return Get<std::string>(0); }
template<> Property &Property::Add<Matrix4x4>(const Matrix4x4 &m) {
	for (int i = 0; i < 4; ++i)        // column-major, like properties.cpp
		for (int j = 0; j < 4; ++j)
			values.push_back(PropertyValue(m.m[j][i]));
	return *this;
}

Properties &Properties::Set(const Property &prop) {
	if (props.find(prop.GetName()) == props.end())
		names.push_back(prop.GetName());
	props[prop.GetName()] = prop;
	return *this;
}
const Property &Properties::Get(const Property &defaultProp) const {
	std::map<std::string, Property>::const_iterator it = props.find(defaultProp.GetName());
	return (it == props.end()) ? defaultProp : it->second;
}

// ---- pieces of the accelerators that live in files outside the CPU path ------------------------

bool BVHAccel::HasNativeSupport(const IntersectionDevice &) const { return true; }
bool BVHAccel::HasHWSupport(const IntersectionDevice &) const { return false; }
HardwareIntersectionKernel *BVHAccel::NewHardwareIntersectionKernel(HardwareIntersectionDevice &) const {
	throw std::runtime_error("no hardware kernels in oracle/_ref");
}
bool MBVHAccel::HasNativeSupport(const IntersectionDevice &) const { return true; }
bool MBVHAccel::HasHWSupport(const IntersectionDevice &) const { return false; }
HardwareIntersectionKernel *MBVHAccel::NewHardwareIntersectionKernel(HardwareIntersectionDevice &) const {
	throw std::runtime_error("no hardware kernels in oracle/_ref");
}

// Intel Embree is not available: only accelerator.bvh.builder.type = CLASSIC can be used
luxrays::ocl::BVHArrayNode *BuildEmbreeBVHBinnedSAH(const BVHParams &, u_int *, const std::deque<const Mesh *> *, std::vector<BVHTreeNode *> &) {
	throw std::runtime_error("Embree builders are not available in oracle/_ref");
}
luxrays::ocl::BVHArrayNode *BuildEmbreeBVHMorton(const BVHParams &, u_int *, const std::deque<const Mesh *> *, std::vector<BVHTreeNode *> &) {
	throw std::runtime_error("Embree builders are not available in oracle/_ref");
}

}   // namespace luxrays
