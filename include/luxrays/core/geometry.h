// luxrays/core/geometry.h -- geometry value types of the LuxRays plugin surface, restated for the
// B200 host layer (no Boost).  Same names, same member layout and the same floating-point
// operation order as the reference so that data crossing the interface is bit-identical:
//   Point / Vector        include/luxrays/core/geometry/point.h, vector.h
//   BBox                  include/luxrays/core/geometry/bbox.h, src/luxrays/core/geometry/bbox.cpp:29-49
//   Ray / RayHit          include/luxrays/core/geometry/ray.h:35-103
//   Triangle              include/luxrays/core/geometry/triangle.h:37-53
//   Matrix4x4             include/luxrays/core/geometry/matrix4x4.h, src/.../matrix4x4.cpp:117-175
//   Transform             include/luxrays/core/geometry/transform.h:48-280
//   Quaternion            include/luxrays/core/geometry/quaternion.h, src/.../quaternion.cpp
//   MotionSystem          include/luxrays/core/geometry/motionsystem.h, src/.../motionsystem.cpp
//   MachineEpsilon        include/luxrays/core/epsilon.h:40-103
// The reference headers of those names are thin forwarders to this file in this tree.
//
// There is deliberately NO CPU Intersect() here (Triangle::Intersect / BBox::IntersectP): the
// product has no CPU intersection path; the CPU restatement lives in oracle/ for the tests.
#ifndef _LUXRAYS_B200_GEOMETRY_H
#define _LUXRAYS_B200_GEOMETRY_H

#include <cmath>
#include <cstring>
#include <limits>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace luxrays {

typedef unsigned int u_int;
typedef unsigned long long u_longlong;

#ifndef NULL_INDEX
#define NULL_INDEX (0xffffffffu)
#endif

template <class T> inline T Max(T a, T b) { return a > b ? a : b; }
template <class T> inline T Min(T a, T b) { return a < b ? a : b; }
template <class T> inline T Clamp(T val, T low, T high) { return val > low ? (val < high ? val : high) : low; }
template <class T> inline void Swap(T &a, T &b) { const T t = a; a = b; b = t; }
template <class T> inline T Lerp(float t, T v1, T v2) { return v1 + t * (v2 - v1); }
template <class T> inline T RoundUp(const T a, const T b) { const T r = a % b; return r == 0 ? a : a + b - r; }

//------------------------------------------------------------------------------
// Vector / Point
//------------------------------------------------------------------------------

class Vector {
public:
	Vector(float _x = 0.f, float _y = 0.f, float _z = 0.f) : x(_x), y(_y), z(_z) { }
	Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); }
	Vector &operator+=(const Vector &v) { x += v.x; y += v.y; z += v.z; return *this; }
	Vector operator-(const Vector &v) const { return Vector(x - v.x, y - v.y, z - v.z); }
	Vector &operator-=(const Vector &v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
	Vector operator-() const { return Vector(-x, -y, -z); }
	Vector operator*(float f) const { return Vector(f * x, f * y, f * z); }
	Vector &operator*=(float f) { x *= f; y *= f; z *= f; return *this; }
	Vector operator/(float f) const { const float inv = 1.f / f; return Vector(x * inv, y * inv, z * inv); }
	float operator[](int i) const { return (&x)[i]; }
	float &operator[](int i) { return (&x)[i]; }
	float LengthSquared() const { return x * x + y * y + z * z; }
	float Length() const { return sqrtf(LengthSquared()); }
	float x, y, z;
};

inline Vector operator*(float f, const Vector &v) { return v * f; }
inline float Dot(const Vector &a, const Vector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector Cross(const Vector &a, const Vector &b) {
	return Vector((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
inline Vector Normalize(const Vector &v) { return v / v.Length(); }

class Point {
public:
	Point(float _x = 0.f, float _y = 0.f, float _z = 0.f) : x(_x), y(_y), z(_z) { }
	Point operator+(const Vector &v) const { return Point(x + v.x, y + v.y, z + v.z); }
	Point &operator+=(const Vector &v) { x += v.x; y += v.y; z += v.z; return *this; }
	Point operator+(const Point &p) const { return Point(x + p.x, y + p.y, z + p.z); }
	Point &operator+=(const Point &p) { x += p.x; y += p.y; z += p.z; return *this; }
	Vector operator-(const Point &p) const { return Vector(x - p.x, y - p.y, z - p.z); }
	Point operator-(const Vector &v) const { return Point(x - v.x, y - v.y, z - v.z); }
	Point &operator-=(const Vector &v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
	Point operator*(float f) const { return Point(f * x, f * y, f * z); }
	Point operator/(float f) const { const float inv = 1.f / f; return Point(inv * x, inv * y, inv * z); }
	Point &operator/=(float f) { const float inv = 1.f / f; x *= inv; y *= inv; z *= inv; return *this; }
	float operator[](int i) const { return (&x)[i]; }
	float &operator[](int i) { return (&x)[i]; }
	bool operator==(const Point &p) const { return x == p.x && y == p.y && z == p.z; }
	float x, y, z;
};

inline std::ostream &operator<<(std::ostream &os, const Point &p) { return os << "Point[" << p.x << ", " << p.y << ", " << p.z << "]"; }
inline std::ostream &operator<<(std::ostream &os, const Vector &v) { return os << "Vector[" << v.x << ", " << v.y << ", " << v.z << "]"; }

//------------------------------------------------------------------------------
// BSphere (include/luxrays/core/geometry/bsphere.h:28-48)
//------------------------------------------------------------------------------

class BSphere {
public:
	BSphere() : center(0.f, 0.f, 0.f), rad(0.f) { }
	BSphere(const Point &c, const float r) : center(c), rad(r) { }

	Point center;
	float rad;
};

inline std::ostream &operator<<(std::ostream &os, const BSphere &s) { return os << "BSphere[" << s.center << ", " << s.rad << "]"; }

//------------------------------------------------------------------------------
// BBox
//------------------------------------------------------------------------------

class BBox {
public:
	BBox() {
		const float inf = std::numeric_limits<float>::infinity();
		pMin = Point(inf, inf, inf);
		pMax = Point(-inf, -inf, -inf);
	}
	BBox(const Point &p) : pMin(p), pMax(p) { }
	BBox(const Point &p1, const Point &p2) {
		pMin = Point(Min(p1.x, p2.x), Min(p1.y, p2.y), Min(p1.z, p2.z));
		pMax = Point(Max(p1.x, p2.x), Max(p1.y, p2.y), Max(p1.z, p2.z));
	}
	void Expand(const float delta) {
		pMin -= Vector(delta, delta, delta);
		pMax += Vector(delta, delta, delta);
	}
	float SurfaceArea() const {
		const Vector d = pMax - pMin;
		return 2.f * (d.x * d.y + d.y * d.z + d.z * d.x);
	}
	bool IsValid() const { return (pMin.x <= pMax.x) && (pMin.y <= pMax.y) && (pMin.z <= pMax.z); }
	bool Inside(const Point &pt) const {
		return pt.x >= pMin.x && pt.x <= pMax.x && pt.y >= pMin.y && pt.y <= pMax.y && pt.z >= pMin.z && pt.z <= pMax.z;
	}
	// bbox.cpp:65-75: centre of the box, radius to a corner (0 for an invalid box)
	BSphere BoundingSphere() const {
		const Point c = (pMin + pMax) * .5f;
		const float dx = c.x - pMax.x, dy = c.y - pMax.y, dz = c.z - pMax.z;
		const float rad = Inside(c) ? sqrtf(dx * dx + dy * dy + dz * dz) : 0.f;
		return BSphere(c, rad);
	}
	Point Center() const { return (pMin + pMax) * .5f; }

	Point pMin, pMax;
};

inline BBox Union(const BBox &b, const Point &p) {
	BBox r;
	r.pMin = Point(Min(b.pMin.x, p.x), Min(b.pMin.y, p.y), Min(b.pMin.z, p.z));
	r.pMax = Point(Max(b.pMax.x, p.x), Max(b.pMax.y, p.y), Max(b.pMax.z, p.z));
	return r;
}
inline BBox Union(const BBox &a, const BBox &b) {
	BBox r;
	r.pMin = Point(Min(a.pMin.x, b.pMin.x), Min(a.pMin.y, b.pMin.y), Min(a.pMin.z, b.pMin.z));
	r.pMax = Point(Max(a.pMax.x, b.pMax.x), Max(a.pMax.y, b.pMax.y), Max(a.pMax.z, b.pMax.z));
	return r;
}

//------------------------------------------------------------------------------
// MachineEpsilon
//------------------------------------------------------------------------------

class MachineEpsilon {
public:
	static void SetMin(const float v) { minEpsilon = v; }
	static float GetMin() { return minEpsilon; }
	static void SetMax(const float v) { maxEpsilon = v; }
	static float GetMax() { return maxEpsilon; }

	static float E(const float value) {
		union { float f; u_int i; } mf;
		mf.f = value;
		mf.i += 0x80u;      // DEFAULT_EPSILON_DISTANCE_FROM_VALUE
		return Clamp(fabsf(mf.f - value), minEpsilon, maxEpsilon);
	}
	static float E(const Vector &v) { return Max(E(v.x), Max(E(v.y), E(v.z))); }
	static float E(const Point &p) { return Max(E(p.x), Max(E(p.y), E(p.z))); }
	static float E(const BBox &bb) { return Max(E(bb.pMin), E(bb.pMax)); }

private:
	static float minEpsilon, maxEpsilon;
};

//------------------------------------------------------------------------------
// Ray / RayHit -- wire types, 48 and 20 bytes
//------------------------------------------------------------------------------

typedef enum { RAY_FLAGS_NONE = 0x00000000, RAY_FLAGS_MASKED = 0x00000001 } RayFlags;

class Ray {
public:
	Ray() : maxt(std::numeric_limits<float>::infinity()), time(0.f), flags(RAY_FLAGS_NONE) { mint = MachineEpsilon::E(1.f); }
	Ray(const Point &origin, const Vector &direction) : o(origin), d(direction),
			maxt(std::numeric_limits<float>::infinity()), time(0.f), flags(RAY_FLAGS_NONE) {
		mint = MachineEpsilon::E(origin);
	}
	Ray(const Point &origin, const Vector &direction, const float start,
			const float end = std::numeric_limits<float>::infinity(), const float t = 0.f) :
			o(origin), d(direction), mint(start), maxt(end), time(t), flags(RAY_FLAGS_NONE) { }

	Point operator()(float t) const { return o + d * t; }
	void Update(const Point &origin, const Vector &direction) {
		o = origin;
		d = direction;
		mint = MachineEpsilon::E(o);
		maxt = std::numeric_limits<float>::infinity();
	}

	Point o;
	Vector d;
	mutable float mint, maxt;
	float time;
	unsigned int flags;
	float pad[2];
};

class RayHit {
public:
	float t;
	float b1, b2;
	unsigned int meshIndex, triangleIndex;

	void SetMiss() { meshIndex = 0xffffffffu; }
	bool Miss() const { return meshIndex == 0xffffffffu; }
};

static_assert(sizeof(Ray) == 48, "luxrays::Ray must stay 48 bytes");
static_assert(sizeof(RayHit) == 20, "luxrays::RayHit must stay 20 bytes");

//------------------------------------------------------------------------------
// Triangle (indices only)
//------------------------------------------------------------------------------

class Triangle {
public:
	Triangle() { }
	Triangle(const unsigned int v0, const unsigned int v1, const unsigned int v2) { v[0] = v0; v[1] = v1; v[2] = v2; }
	BBox WorldBound(const Point *verts) const { return Union(BBox(verts[v[0]], verts[v[1]]), verts[v[2]]); }
	static float Area(const Point &p0, const Point &p1, const Point &p2) { return .5f * Cross(p1 - p0, p2 - p0).Length(); }
	float Area(const Point *verts) const { return Area(verts[v[0]], verts[v[1]], verts[v[2]]); }
	unsigned int v[3];
};

//------------------------------------------------------------------------------
// Matrix4x4 / Transform
//------------------------------------------------------------------------------

class Matrix4x4 {
public:
	Matrix4x4() {
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j)
				m[i][j] = (i == j) ? 1.f : 0.f;
	}
	Matrix4x4(const float mat[4][4]) { memcpy(m, mat, sizeof(m)); }
	explicit Matrix4x4(const float *mat16) { memcpy(m, mat16, sizeof(m)); }     // row-major
	Matrix4x4 Transpose() const;
	float Determinant() const;
	Matrix4x4 Inverse() const;      // throws std::runtime_error on a singular matrix
	Matrix4x4 operator*(const Matrix4x4 &b) const {
		Matrix4x4 r;
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j)
				r.m[i][j] = m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j] + m[i][2] * b.m[2][j] + m[i][3] * b.m[3][j];
		return r;
	}
	float m[4][4];
	static const Matrix4x4 MAT_IDENTITY;
};

inline Point operator*(const Matrix4x4 &m, const Point &pt) {
	const float x = pt.x, y = pt.y, z = pt.z;
	const Point pr(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z + m.m[0][3],
			m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z + m.m[1][3],
			m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z + m.m[2][3]);
	const float w = m.m[3][0] * x + m.m[3][1] * y + m.m[3][2] * z + m.m[3][3];
	return (w != 1.f) ? pr / w : pr;
}
inline Vector operator*(const Matrix4x4 &m, const Vector &v) {
	const float x = v.x, y = v.y, z = v.z;
	return Vector(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z,
			m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z,
			m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z);
}
inline BBox operator*(const Matrix4x4 &m, const BBox &b) {
	BBox r(m * b.pMin, m * b.pMax);
	r = Union(r, m * Point(b.pMax.x, b.pMin.y, b.pMin.z));
	r = Union(r, m * Point(b.pMin.x, b.pMax.y, b.pMin.z));
	r = Union(r, m * Point(b.pMin.x, b.pMin.y, b.pMax.z));
	r = Union(r, m * Point(b.pMax.x, b.pMax.y, b.pMin.z));
	r = Union(r, m * Point(b.pMax.x, b.pMin.y, b.pMax.z));
	r = Union(r, m * Point(b.pMin.x, b.pMax.y, b.pMax.z));
	return r;
}

class Transform;
class InvTransform {
public:
	const Transform &ref;
protected:
	InvTransform(const Transform &t) : ref(t) { }
	friend InvTransform Inverse(const Transform &t);
};

class Transform {
public:
	Transform() { }
	explicit Transform(const float mat[4][4]) : m(mat) { mInv = m.Inverse(); }
	explicit Transform(const Matrix4x4 &mat) : m(mat) { mInv = m.Inverse(); }
	Transform(const Matrix4x4 &mat, const Matrix4x4 &minv) : m(mat), mInv(minv) { }
	Transform(const InvTransform &t);
	Matrix4x4 GetMatrix() const { return m; }
	Transform operator*(const Transform &t2) const { return Transform(m * t2.m, t2.mInv * mInv); }
	bool SwapsHandedness() const;

	static const Transform TRANS_IDENTITY;
	Matrix4x4 m, mInv;
};

inline InvTransform Inverse(const Transform &t) { return InvTransform(t); }
inline Transform::Transform(const InvTransform &t) : m(t.ref.mInv), mInv(t.ref.m) { }

inline Point operator*(const Transform &t, const Point &p) { return t.m * p; }
inline Vector operator*(const Transform &t, const Vector &v) { return t.m * v; }
inline BBox operator*(const Transform &t, const BBox &b) { return t.m * b; }
inline Point operator*(const InvTransform &t, const Point &p) { return t.ref.mInv * p; }
inline Vector operator*(const InvTransform &t, const Vector &v) { return t.ref.mInv * v; }
inline Point &operator*=(Point &p, const Transform &t) { p = t.m * p; return p; }

Transform Translate(const Vector &delta);
Transform Scale(float x, float y, float z);
Transform RotateX(float angle);
Transform RotateY(float angle);
Transform RotateZ(float angle);

//------------------------------------------------------------------------------
// Quaternion
//------------------------------------------------------------------------------

class Quaternion {
public:
	Quaternion() : w(1.f), v(0.f, 0.f, 0.f) { }
	Quaternion(float _w, const Vector &_v) : w(_w), v(_v) { }
	explicit Quaternion(const Matrix4x4 &m);    // from a rotation matrix (quaternion.cpp:69-107)
	void ToMatrix(float m[4][4]) const;
	float w;
	Vector v;
};
inline Quaternion operator+(const Quaternion &a, const Quaternion &b) { return Quaternion(a.w + b.w, a.v + b.v); }
inline Quaternion operator*(float f, const Quaternion &q) { return Quaternion(q.w * f, q.v * f); }
inline float Dot(const Quaternion &a, const Quaternion &b) { return a.w * b.w + Dot(a.v, b.v); }
inline Quaternion Normalize(const Quaternion &q) { return (1.f / sqrtf(Dot(q, q))) * q; }
Quaternion Slerp(float t, const Quaternion &q1, const Quaternion &q2);

//------------------------------------------------------------------------------
// MotionSystem.  InterpolatedTransform keeps the exact 576-byte layout of
// ocl::InterpolatedTransform (motionsystem_types.cl:21-47): MBVHKernel uploads these objects
// byte-for-byte (mbvhaccelhw.cpp:160-166).
//------------------------------------------------------------------------------

class InterpolatedTransform {
public:
	InterpolatedTransform() : startTime(0.f), endTime(0.f) { InitFlags(); }
	InterpolatedTransform(float st, float et, const Transform &s, const Transform &e);

	Matrix4x4 Sample(const float time) const;
	BBox Bound(BBox ibox, const bool storingGlobal2Local) const;
	bool IsStatic() const { return !isActive; }

	class DecomposedTransform {
	public:
		DecomposedTransform() : Sx(0), Sy(0), Sz(0), Sxy(0), Sxz(0), Syz(0), Tx(0), Ty(0), Tz(0), Px(0), Py(0), Pz(0), Pw(0), Valid(false) { }
		explicit DecomposedTransform(const Matrix4x4 &m);
		float Sx, Sy, Sz;
		float Sxy, Sxz, Syz;
		Matrix4x4 R;
		float Tx, Ty, Tz;
		float Px, Py, Pz, Pw;
		bool Valid;
	};

	float startTime, endTime;
	Transform start, end;
	DecomposedTransform startT, endT;
	Quaternion startQ, endQ;
	int hasRotation, hasTranslation, hasScale;
	int hasTranslationX, hasTranslationY, hasTranslationZ;
	int hasScaleX, hasScaleY, hasScaleZ;
	int isActive;

private:
	void InitFlags() {
		hasRotation = hasTranslation = hasScale = 0;
		hasTranslationX = hasTranslationY = hasTranslationZ = 0;
		hasScaleX = hasScaleY = hasScaleZ = 0;
		isActive = 0;
	}
};
static_assert(sizeof(InterpolatedTransform) == 576, "InterpolatedTransform must match ocl::InterpolatedTransform");

class MotionSystem {
public:
	MotionSystem();
	explicit MotionSystem(const Transform &t);
	MotionSystem(const std::vector<float> &t, const std::vector<Transform> &transforms);

	bool IsStatic() const { return times.size() <= 1; }
	float StartTime() const { return times.front(); }
	float EndTime() const { return times.back(); }
	Matrix4x4 Sample(const float time) const;
	Matrix4x4 SampleInverse(const float time) const;
	BBox Bound(BBox ibox, const bool storingGlobal2Local) const;

	std::vector<float> times;
	std::vector<InterpolatedTransform> interpolatedTransforms;
	std::vector<InterpolatedTransform> interpolatedInverseTransforms;

private:
	void Init(const std::vector<float> &t, const std::vector<Transform> &transforms);
};

}   // namespace luxrays

#endif
