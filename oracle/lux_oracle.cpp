// lux_oracle.cpp -- CPU ORACLE for the LuxRays closest-hit path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a from-scratch CPU restatement of the reference's *native* (CPU) BVH / MBVH
// closest-hit algorithm.  It exists so that tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py have something to check (and time) the CUDA
// device against.  NOTHING in the product (luxcore_b200/, include/) may include, link or call
// it; the product fails loudly when its CUDA library is missing.
//
// PARITY PINNED AGAINST THE REFERENCE ITSELF.  The reference ships no golden vectors, known-answer
// tests or unit tests for Intersect, and its build system cannot run here
// (no Boost / Embree / cmake configuration).  But the files of THIS path compile from their own
// sources with no-op stand-ins for the Boost serialization / foreach / lexical_cast declarations
// their headers mention (oracle/ref/Makefile -> oracle/_ref/libluxrays_ref.so: triangle.h, bbox.cpp,
// bvhclassicbuild.cpp, bvhaccel.cpp, mbvhaccel.cpp, transform.cpp, matrix4x4.cpp, motionsystem.cpp,
// quaternion.cpp, epsilon.cpp, trianglemesh.cpp, unchanged, with the reference's CPU flags).
// tests/test_oracle_pinned_cpu.py compares this restatement with that library BIT FOR BIT:
// MachineEpsilon, Matrix4x4::Inverse, Triangle::Intersect, BBox::IntersectP, mesh bounding boxes,
// the CLASSIC builder's arrays (arity 2/4/8, cost samples), BVHAccel::Intersect and
// MBVHAccel::Intersect / Update (instances, motion blur) on the fixtures and on stress batches
// (grazing, axis-parallel, degenerate rays) -- and, where /root/reference is absent, with the
// vectors that library produced (tests/golden/ref_vectors.npz, tools/make_ref_vectors.py).
// Further pins: a topology-free brute-force closest hit over all triangles (orc_brute_*) and an
// independent second implementation of the CLASSIC builder in the product's host layer that must
// produce bit-identical node arrays.
//
// All citations are relative to /root/reference.  Arithmetic is IEEE binary32 with no FMA
// contraction and no fast-math (cmake/PlatformSpecific.cmake:265-273); see oracle/Makefile.
//
//   MachineEpsilon::E ............ include/luxrays/core/epsilon.h:48-86, epsilon_types.cl:21-30
//   Triangle::Intersect .......... include/luxrays/core/geometry/triangle.h:55-89
//   BBox::IntersectP ............. src/luxrays/core/geometry/bbox.cpp:147-165
//   BBox Union/Expand ............ src/luxrays/core/geometry/bbox.cpp:29-49, bbox.h:42-83
//   CLASSIC builder .............. src/luxrays/core/bvh/bvhclassicbuild.cpp:51-233
//   BVHAccel::Init / Intersect ... src/luxrays/accelerators/bvhaccel.cpp:72-168 / :170-260
//   MBVHAccel::Init / Update / Intersect  src/luxrays/accelerators/mbvhaccel.cpp:58-216 / :237-250 / :252-357
//   Transform / InvTransform ops . include/luxrays/core/geometry/transform.h:117-280
//   Matrix4x4::Inverse ........... src/luxrays/core/geometry/matrix4x4.cpp:117-175
//   Matrix4x4 * Point/BBox ....... include/luxrays/core/geometry/matrix4x4op.h:31-41,95-103
//   MotionSystem / InterpolatedTransform  src/luxrays/core/geometry/motionsystem.cpp:37-159,168-276,296-346
//   Quaternion ................... src/luxrays/core/geometry/quaternion.cpp:27-164, quaternion.h:63-88
//   Mesh bounding boxes .......... src/luxrays/core/trianglemesh.cpp:72-83,241-275

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace orc {

typedef uint32_t u32;
static const u32 NULL_INDEX = 0xffffffffu;

//------------------------------------------------------------------------------------------------
// Wire types (ray_types.cl:21-36, bvhbuild_types.cl:21-42)
//------------------------------------------------------------------------------------------------

struct Ray {            // 48 bytes
	float o[3], d[3];
	float mint, maxt, time;
	u32 flags;
	float pad[2];
};

struct RayHit {         // 20 bytes
	float t, b1, b2;
	u32 meshIndex, triangleIndex;
};

struct Node {           // 32 bytes, BVHArrayNode
	union {
		struct { float bmin[3], bmax[3]; } box;
		struct { u32 v[3], meshIndex, triangleIndex; } tri;
		struct { u32 leafIndex, transformIndex, motionIndex, meshOffsetIndex; } inst;
	};
	u32 nodeData;
	int pad0;
};

static_assert(sizeof(Ray) == 48, "Ray");
static_assert(sizeof(RayHit) == 20, "RayHit");
static_assert(sizeof(Node) == 32, "Node");

static inline bool IsLeaf(u32 nd) { return (nd & 0x80000000u) != 0; }
static inline u32 Skip(u32 nd) { return nd & 0x7fffffffu; }

//------------------------------------------------------------------------------------------------
// Small maths
//------------------------------------------------------------------------------------------------

struct V3 { float x, y, z; };

static inline V3 v3(float x, float y, float z) { V3 r = { x, y, z }; return r; }
static inline V3 sub(const V3 &a, const V3 &b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
// vector.h:151-153: x*x + y*y + z*z evaluated left to right
static inline float dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vector.h:159-163
static inline V3 cross(const V3 &a, const V3 &b) {
	return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
static inline float fminr(float a, float b) { return a < b ? a : b; }   // utils.h Min
static inline float fmaxr(float a, float b) { return a > b ? a : b; }   // utils.h Max

// epsilon.h:48-53,76-84.  min/max are run-time settable in the reference (scene.epsilon.min/max).
static float g_epsMin = 1e-5f, g_epsMax = 1e-1f;

static inline float EpsF(float value) {
	union { float f; u32 i; } mf;
	mf.f = value;
	mf.i += 0x80u;
	const float e = fabsf(mf.f - value);
	// Clamp(val, low, high) = val > low ? (val < high ? val : high) : low   (utils.h:146-148)
	return e > g_epsMin ? (e < g_epsMax ? e : g_epsMax) : g_epsMin;
}
static inline float EpsV(const V3 &p) { return fmaxr(EpsF(p.x), fmaxr(EpsF(p.y), EpsF(p.z))); }

struct Box {
	V3 lo, hi;
	Box() {
		const float inf = std::numeric_limits<float>::infinity();
		lo = v3(inf, inf, inf);
		hi = v3(-inf, -inf, -inf);
	}
};

static inline Box BoxOf2(const V3 &a, const V3 &b) {       // bbox.h:48-55
	Box r;
	r.lo = v3(fminr(a.x, b.x), fminr(a.y, b.y), fminr(a.z, b.z));
	r.hi = v3(fmaxr(a.x, b.x), fmaxr(a.y, b.y), fmaxr(a.z, b.z));
	return r;
}
static inline Box UnionP(const Box &b, const V3 &p) {      // bbox.cpp:29-38
	Box r;
	r.lo = v3(fminr(b.lo.x, p.x), fminr(b.lo.y, p.y), fminr(b.lo.z, p.z));
	r.hi = v3(fmaxr(b.hi.x, p.x), fmaxr(b.hi.y, p.y), fmaxr(b.hi.z, p.z));
	return r;
}
static inline Box UnionB(const Box &a, const Box &b) {     // bbox.cpp:40-49
	Box r;
	r.lo = v3(fminr(a.lo.x, b.lo.x), fminr(a.lo.y, b.lo.y), fminr(a.lo.z, b.lo.z));
	r.hi = v3(fmaxr(a.hi.x, b.hi.x), fmaxr(a.hi.y, b.hi.y), fmaxr(a.hi.z, b.hi.z));
	return r;
}
static inline float EpsB(const Box &b) { return fmaxr(EpsV(b.lo), EpsV(b.hi)); }
static inline void ExpandBox(Box &b, float d) {            // bbox.h:80-83
	b.lo = v3(b.lo.x - d, b.lo.y - d, b.lo.z - d);
	b.hi = v3(b.hi.x + d, b.hi.y + d, b.hi.z + d);
}
static inline float SurfaceArea(const Box &b) {            // bbox.h:97-100
	const V3 d = sub(b.hi, b.lo);
	return 2.f * (d.x * d.y + d.y * d.z + d.z * d.x);
}
static inline float Axis(const V3 &v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

//------------------------------------------------------------------------------------------------
// Triangle::Intersect (triangle.h:55-89) and BBox::IntersectP (bbox.cpp:147-165)
//------------------------------------------------------------------------------------------------

struct RayC {   // the mutable copy the reference walks with
	V3 o, d;
	float mint, maxt, time;
};

static inline bool TriangleIntersect(const RayC &ray, const V3 &p0, const V3 &p1, const V3 &p2,
		float *t, float *b1, float *b2) {
	const V3 e1 = sub(p1, p0);
	const V3 e2 = sub(p2, p0);
	const V3 s1 = cross(ray.d, e2);

	const float divisor = dot(s1, e1);
	if (divisor == 0.f)
		return false;
	const float invDivisor = 1.f / divisor;

	const V3 d = sub(ray.o, p0);
	*b1 = dot(d, s1) * invDivisor;
	if (*b1 < 0.f)
		return false;

	const V3 s2 = cross(d, e1);
	*b2 = dot(ray.d, s2) * invDivisor;
	if (*b2 < 0.f)
		return false;

	const float b0 = 1.f - *b1 - *b2;
	if (b0 < 0.f)
		return false;

	*t = dot(e2, s2) * invDivisor;
	if (*t < ray.mint || *t > ray.maxt)
		return false;
	return true;
}

static inline bool BoxIntersectP(const RayC &ray, const float *bmin, const float *bmax) {
	float t0 = ray.mint, t1 = ray.maxt;
	const float o[3] = { ray.o.x, ray.o.y, ray.o.z };
	const float d[3] = { ray.d.x, ray.d.y, ray.d.z };
	for (int i = 0; i < 3; ++i) {
		const float invRayDir = 1.f / d[i];
		float tNear = (bmin[i] - o[i]) * invRayDir;
		float tFar = (bmax[i] - o[i]) * invRayDir;
		if (tNear > tFar) { const float s = tNear; tNear = tFar; tFar = s; }
		t0 = tNear > t0 ? tNear : t0;
		t1 = tFar < t1 ? tFar : t1;
		if (t0 > t1) return false;
	}
	return true;
}

//------------------------------------------------------------------------------------------------
// Matrix4x4 / Transform (matrix4x4.cpp, matrix4x4op.h, transform.h)
//------------------------------------------------------------------------------------------------

struct M44 { float m[4][4]; };

static M44 Identity() {
	M44 r;
	memset(&r, 0, sizeof(r));
	r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.f;
	return r;
}

static M44 Transpose(const M44 &a) {
	M44 r;
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			r.m[i][j] = a.m[j][i];
	return r;
}

// matrix4x4.cpp:117-175 -- Gauss-Jordan with full pivoting; `>=` keeps the LAST largest pivot.
static bool Inverse(const M44 &src, M44 *out) {
	int indxc[4], indxr[4];
	int ipiv[4] = { 0, 0, 0, 0 };
	float minv[4][4];
	memcpy(minv, src.m, sizeof(minv));
	for (int i = 0; i < 4; ++i) {
		int irow = -1, icol = -1;
		float big = 0.f;
		for (int j = 0; j < 4; ++j) {
			if (ipiv[j] != 1) {
				for (int k = 0; k < 4; ++k) {
					if (ipiv[k] == 0) {
						if (fabsf(minv[j][k]) >= big) {
							big = fabsf(minv[j][k]);
							irow = j;
							icol = k;
						}
					} else if (ipiv[k] > 1)
						return false;
				}
			}
		}
		++ipiv[icol];
		if (irow != icol) {
			for (int k = 0; k < 4; ++k)
				std::swap(minv[irow][k], minv[icol][k]);
		}
		indxr[i] = irow;
		indxc[i] = icol;
		if (minv[icol][icol] == 0.f)
			return false;
		const float pivinv = 1.f / minv[icol][icol];
		minv[icol][icol] = 1.f;
		for (int j = 0; j < 4; ++j)
			minv[icol][j] *= pivinv;
		for (int j = 0; j < 4; ++j) {
			if (j != icol) {
				const float save = minv[j][icol];
				minv[j][icol] = 0;
				for (int k = 0; k < 4; ++k)
					minv[j][k] -= minv[icol][k] * save;
			}
		}
	}
	for (int j = 3; j >= 0; --j) {
		if (indxr[j] != indxc[j]) {
			for (int k = 0; k < 4; ++k)
				std::swap(minv[k][indxr[j]], minv[k][indxc[j]]);
		}
	}
	memcpy(out->m, minv, sizeof(minv));
	return true;
}

static M44 InverseOrThrow(const M44 &a) {
	M44 r;
	if (!Inverse(a, &r))
		throw std::runtime_error("Singular matrix in MatrixInvert");
	return r;
}

static M44 Mul(const M44 &a, const M44 &b) {   // matrix4x4.h operator*
	M44 r;
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] +
					a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
	return r;
}

// transform.h:117-130 / matrix4x4op.h:31-41: divide by w only when w != 1, via inv = 1/w
static inline V3 XfPoint(const M44 &m, const V3 &p) {
	const float x = p.x, y = p.y, z = p.z;
	const V3 pr = v3(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z + m.m[0][3],
			m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z + m.m[1][3],
			m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z + m.m[2][3]);
	const float w = m.m[3][0] * x + m.m[3][1] * y + m.m[3][2] * z + m.m[3][3];
	if (w != 1.f) {
		const float inv = 1.f / w;     // point.h:100-103
		return v3(inv * pr.x, inv * pr.y, inv * pr.z);
	}
	return pr;
}
// transform.h:160-167
static inline V3 XfVector(const M44 &m, const V3 &v) {
	const float x = v.x, y = v.y, z = v.z;
	return v3(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z,
			m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z,
			m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z);
}
// transform.h:247-262: mint/maxt/time copied, d NOT renormalised
static inline RayC XfRay(const M44 &m, const RayC &r) {
	RayC o;
	o.o = XfPoint(m, r.o);
	o.d = XfVector(m, r.d);
	o.mint = r.mint;
	o.maxt = r.maxt;
	o.time = r.time;
	return o;
}
// transform.h:271-280: union of the 8 transformed corners in this exact order
static Box XfBox(const M44 &m, const Box &b) {
	Box r = BoxOf2(XfPoint(m, b.lo), XfPoint(m, b.hi));
	r = UnionP(r, XfPoint(m, v3(b.hi.x, b.lo.y, b.lo.z)));
	r = UnionP(r, XfPoint(m, v3(b.lo.x, b.hi.y, b.lo.z)));
	r = UnionP(r, XfPoint(m, v3(b.lo.x, b.lo.y, b.hi.z)));
	r = UnionP(r, XfPoint(m, v3(b.hi.x, b.hi.y, b.lo.z)));
	r = UnionP(r, XfPoint(m, v3(b.hi.x, b.lo.y, b.hi.z)));
	r = UnionP(r, XfPoint(m, v3(b.lo.x, b.hi.y, b.hi.z)));
	return r;
}

struct Xform { M44 m, mInv; };   // transform.h:48-91

static Xform MakeXform(const M44 &m) {
	Xform t;
	t.m = m;
	t.mInv = InverseOrThrow(m);
	return t;
}
static Xform InvXform(const Xform &t) { Xform r; r.m = t.mInv; r.mInv = t.m; return r; }

//------------------------------------------------------------------------------------------------
// Quaternion (quaternion.cpp, quaternion.h)
//------------------------------------------------------------------------------------------------

struct Quat { float w; V3 v; };

static void OrthoNormalize(float m[4][4]) {     // quaternion.cpp:27-66
	float len, temp[3][3];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			temp[i][j] = m[i][j];

	len = sqrtf(temp[0][0] * temp[0][0] + temp[0][1] * temp[0][1] + temp[0][2] * temp[0][2]);
	len = (len == 0.f) ? 1.f : 1.f / len;
	temp[0][0] *= len; temp[0][1] *= len; temp[0][2] *= len;

	temp[2][0] = (temp[0][1] * temp[1][2] - temp[0][2] * temp[1][1]);
	temp[2][1] = (temp[0][2] * temp[1][0] - temp[0][0] * temp[1][2]);
	temp[2][2] = (temp[0][0] * temp[1][1] - temp[0][1] * temp[1][0]);

	len = sqrtf(temp[2][0] * temp[2][0] + temp[2][1] * temp[2][1] + temp[2][2] * temp[2][2]);
	len = (len == 0.f) ? 1.f : 1.f / len;
	temp[2][0] *= len; temp[2][1] *= len; temp[2][2] *= len;

	temp[1][0] = (temp[2][1] * temp[0][2] - temp[2][2] * temp[0][1]);
	temp[1][1] = (temp[2][2] * temp[0][0] - temp[2][0] * temp[0][2]);
	temp[1][2] = (temp[2][0] * temp[0][1] - temp[2][1] * temp[0][0]);

	len = sqrtf(temp[1][0] * temp[1][0] + temp[1][1] * temp[1][1] + temp[1][2] * temp[1][2]);
	len = (len == 0.f) ? 1.f : 1.f / len;
	temp[1][0] *= len; temp[1][1] *= len; temp[1][2] *= len;

	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			m[i][j] = temp[i][j];
}

static Quat QuatFromMatrix(const M44 &mat) {    // quaternion.cpp:69-107
	float o[4][4];
	memcpy(o, mat.m, sizeof(o));
	OrthoNormalize(o);
	Quat q;
	const float trace = o[0][0] + o[1][1] + o[2][2] + 1.f;
	if (trace > 1e-6f) {
		const float s = sqrtf(trace) * 2.f;
		q.v = v3((o[1][2] - o[2][1]) / s, (o[2][0] - o[0][2]) / s, (o[0][1] - o[1][0]) / s);
		q.w = 0.25f * s;
	} else if (o[0][0] > o[1][1] && o[0][0] > o[2][2]) {
		const float s = sqrtf(1.f + o[0][0] - o[1][1] - o[2][2]) * 2.f;
		q.v = v3(0.25f * s, (o[0][1] + o[1][0]) / s, (o[2][0] + o[0][2]) / s);
		q.w = (o[1][2] - o[2][1]) / s;
	} else if (o[1][1] > o[2][2]) {
		const float s = sqrtf(1.f + o[1][1] - o[0][0] - o[2][2]) * 2.f;
		q.v = v3((o[0][1] + o[1][0]) / s, 0.25f * s, (o[1][2] + o[2][1]) / s);
		q.w = (o[2][0] - o[0][2]) / s;
	} else {
		const float s = sqrtf(1.f + o[2][2] - o[0][0] - o[1][1]) * 2.f;
		q.v = v3((o[2][0] + o[0][2]) / s, (o[1][2] + o[2][1]) / s, 0.25f * s);
		q.w = (o[0][1] - o[1][0]) / s;
	}
	return q;
}

static inline float QDot(const Quat &a, const Quat &b) { return a.w * b.w + dot(a.v, b.v); }   // quaternion.h:80-82
static inline Quat QScale(float f, const Quat &q) {      // quaternion.h:69-71 (q.w * f, q.v * f)
	Quat r; r.w = q.w * f; r.v = v3(q.v.x * f, q.v.y * f, q.v.z * f); return r;
}
static inline Quat QAdd(const Quat &a, const Quat &b) {
	Quat r; r.w = a.w + b.w; r.v = v3(a.v.x + b.v.x, a.v.y + b.v.y, a.v.z + b.v.z); return r;
}
static inline Quat QNormalize(const Quat &q) { return QScale(1.f / sqrtf(QDot(q, q)), q); }   // quaternion.h:84-86

static void QuatToMatrix(const Quat &q, float m[4][4]) {   // quaternion.cpp:117-142
	const float xx = q.v.x * q.v.x, yy = q.v.y * q.v.y, zz = q.v.z * q.v.z;
	const float xy = q.v.x * q.v.y, xz = q.v.x * q.v.z, yz = q.v.y * q.v.z;
	const float xw = q.v.x * q.w, yw = q.v.y * q.w, zw = q.v.z * q.w;
	m[0][0] = 1.f - 2.f * (yy + zz);
	m[1][0] = 2.f * (xy - zw);
	m[2][0] = 2.f * (xz + yw);
	m[0][1] = 2.f * (xy + zw);
	m[1][1] = 1.f - 2.f * (xx + zz);
	m[2][1] = 2.f * (yz - xw);
	m[0][2] = 2.f * (xz - yw);
	m[1][2] = 2.f * (yz + xw);
	m[2][2] = 1.f - 2.f * (xx + yy);
	m[0][3] = m[1][3] = m[2][3] = 0.f;
	m[3][0] = m[3][1] = m[3][2] = 0.f;
	m[3][3] = 1.f;
}

static Quat Slerp(float t, const Quat &q1, const Quat &q2) {   // quaternion.cpp:144-164
	float cos_phi = QDot(q1, q2);
	const float sign = (cos_phi > 0.f) ? 1.f : -1.f;
	cos_phi *= sign;
	float f1, f2;
	if (1.f - cos_phi > 1e-6f) {
		const float phi = acosf(cos_phi);
		const float sin_phi = sinf(phi);
		f1 = sinf((1.f - t) * phi) / sin_phi;
		f2 = sinf(t * phi) / sin_phi;
	} else {
		f1 = 1.f - t;
		f2 = t;
	}
	return QAdd(QScale(f1, q1), QScale(sign * f2, q2));
}

//------------------------------------------------------------------------------------------------
// InterpolatedTransform / MotionSystem (motionsystem.cpp).  The struct below has the exact
// 576-byte layout of ocl::InterpolatedTransform (motionsystem_types.cl:21-47) so the tests can
// hand it unchanged to the C-ABI device.
//------------------------------------------------------------------------------------------------

struct Decomposed {     // 120 bytes
	float Sx, Sy, Sz;
	float Sxy, Sxz, Syz;
	M44 R;
	float Tx, Ty, Tz;
	float Px, Py, Pz, Pw;
	bool Valid;
};

struct InterpT {        // 576 bytes
	float startTime, endTime;
	Xform start, end;
	Decomposed startT, endT;
	Quat startQ, endQ;
	int hasRotation, hasTranslation, hasScale;
	int hasTranslationX, hasTranslationY, hasTranslationZ;
	int hasScaleX, hasScaleY, hasScaleZ;
	int isActive;
};
static_assert(sizeof(Decomposed) == 120, "Decomposed");
static_assert(sizeof(InterpT) == 576, "InterpT");

static float Det2x2(float a00, float a01, float a10, float a11) { return a00 * a11 - a01 * a10; }
static float Det3x3(float A[3][3]) {
	return A[0][0] * Det2x2(A[1][1], A[1][2], A[2][1], A[2][2]) -
			A[0][1] * Det2x2(A[1][0], A[1][2], A[2][0], A[2][2]) +
			A[0][2] * Det2x2(A[1][0], A[1][1], A[2][0], A[2][1]);
}
static float Determinant(const M44 &mm) {       // matrix4x4.cpp:80-115
	float result = 0, s = -1;
	float A[3][3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
			A[i][j] = mm.m[i][j + 1];
	int k = 0;
	while (true) {
		if (mm.m[3][k] != 0.f)
			result += s * mm.m[3][k] * Det3x3(A);
		if (k >= 3)
			break;
		s *= -1;
		for (int i = 0; i < 3; i++)
			A[i][k] = mm.m[i][k];
		k++;
	}
	return result;
}

static inline float Len(const V3 &v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
static inline V3 Scale(const V3 &v, float f) { return v3(v.x * f, v.y * f, v.z * f); }

static Decomposed Decompose(const M44 &mIn) {   // motionsystem.cpp:168-276 (unmatrix, Graphics Gems II)
	Decomposed D;
	memset(&D, 0, sizeof(D));
	D.R = mIn;
	D.Valid = false;
	M44 &R = D.R;
	if (R.m[3][3] == 0)
		return D;
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			R.m[i][j] /= R.m[3][3];   // NB: divides by the *current* m[3][3], which becomes 1 at the last element
	M44 pmat = R;
	for (int i = 0; i < 3; ++i)
		pmat.m[i][3] = 0.f;
	pmat.m[3][3] = 1.f;
	if (Determinant(pmat) == 0.f)
		return D;

	if (R.m[3][0] != 0.f || R.m[3][1] != 0.f || R.m[3][2] != 0.f) {
		const float prhs[4] = { R.m[3][0], R.m[3][1], R.m[3][2], R.m[3][3] };
		const M44 A = Transpose(InverseOrThrow(pmat));
		float psol[4];
		for (int i = 0; i < 4; ++i)
			psol[i] = A.m[i][0] * prhs[0] + A.m[i][1] * prhs[1] + A.m[i][2] * prhs[2] + A.m[i][3] * prhs[3];
		D.Px = psol[0]; D.Py = psol[1]; D.Pz = psol[2]; D.Pw = psol[3];
		R.m[3][0] = R.m[3][1] = R.m[3][2] = 0.f;
		R.m[3][3] = 1.f;
	}

	D.Tx = R.m[0][3];
	D.Ty = R.m[1][3];
	D.Tz = R.m[2][3];
	for (int i = 0; i < 3; ++i)
		R.m[i][3] = 0.f;

	V3 row[3];
	for (int i = 0; i < 3; ++i)
		row[i] = v3(R.m[i][0], R.m[i][1], R.m[i][2]);

	D.Sx = Len(row[0]);
	row[0] = Scale(row[0], 1.f / D.Sx);

	D.Sxy = dot(row[0], row[1]);
	row[1] = sub(row[1], Scale(row[0], D.Sxy));

	D.Sy = Len(row[1]);
	row[1] = Scale(row[1], 1.f / D.Sy);
	D.Sxy /= D.Sy;

	D.Sxz = dot(row[0], row[2]);
	row[2] = sub(row[2], Scale(row[0], D.Sxz));
	D.Syz = dot(row[1], row[2]);
	row[2] = sub(row[2], Scale(row[1], D.Syz));

	D.Sz = Len(row[2]);
	row[2] = Scale(row[2], 1.f / D.Sz);
	D.Sxz /= D.Sz;
	D.Syz /= D.Sz;

	if (dot(row[0], cross(row[1], row[2])) < 0.f) {
		D.Sx *= -1.f;
		D.Sy *= -1.f;
		D.Sz *= -1.f;
		for (int i = 0; i < 3; ++i)
			row[i] = Scale(row[i], -1.f);
	}
	for (int i = 0; i < 3; ++i) {
		R.m[i][0] = row[i].x;
		R.m[i][1] = row[i].y;
		R.m[i][2] = row[i].z;
	}
	D.Valid = true;
	return D;
}

static InterpT MakeInterp(float st, float et, const Xform &s, const Xform &e) {   // motionsystem.cpp:37-73
	InterpT it;
	memset(&it, 0, sizeof(it));
	it.startTime = st;
	it.endTime = et;
	it.start = s;
	it.end = e;
	// default-constructed members of the reference object
	it.startT.Valid = false;
	it.endT.Valid = false;
	it.startQ.w = 1.f; it.startQ.v = v3(0, 0, 0);
	it.endQ.w = 1.f; it.endQ.v = v3(0, 0, 0);
	if (st == et)
		return it;

	it.startT = Decompose(s.m);
	it.endT = Decompose(e.m);
	if (!it.startT.Valid)
		throw std::runtime_error("Singular start matrix in InterpolatedTransform, interpolation disabled");
	if (!it.endT.Valid)
		throw std::runtime_error("Singular end matrix in InterpolatedTransform, interpolation disabled");

	it.startQ = QNormalize(QuatFromMatrix(it.startT.R));
	it.endQ = QNormalize(QuatFromMatrix(it.endT.R));

	it.hasTranslationX = it.startT.Tx != it.endT.Tx;
	it.hasTranslationY = it.startT.Ty != it.endT.Ty;
	it.hasTranslationZ = it.startT.Tz != it.endT.Tz;
	it.hasTranslation = it.hasTranslationX || it.hasTranslationY || it.hasTranslationZ;

	it.hasScaleX = it.startT.Sx != it.endT.Sx;
	it.hasScaleY = it.startT.Sy != it.endT.Sy;
	it.hasScaleZ = it.startT.Sz != it.endT.Sz;
	it.hasScale = it.hasScaleX || it.hasScaleY || it.hasScaleZ;

	it.hasRotation = fabsf(QDot(it.startQ, it.endQ) - 1.f) >= 1e-6f;
	it.isActive = it.hasTranslation || it.hasScale || it.hasRotation;
	return it;
}

static inline float Lerp(float t, float v1, float v2) { return v1 + t * (v2 - v1); }   // utils.h:122-125

static M44 SampleInterp(const InterpT &it, float time) {   // motionsystem.cpp:89-159
	if (!it.isActive)
		return it.start.m;
	if (time <= it.startTime)
		return it.start.m;
	if (time >= it.endTime)
		return it.end.m;

	const float w = it.endTime - it.startTime;
	const float d = time - it.startTime;
	const float le = d / w;

	M44 r;
	if (it.hasTranslation && !(it.hasScale || it.hasRotation)) {
		r = it.start.m;
		if (it.hasTranslationX) r.m[0][3] = Lerp(le, it.startT.Tx, it.endT.Tx);
		if (it.hasTranslationY) r.m[1][3] = Lerp(le, it.startT.Ty, it.endT.Ty);
		if (it.hasTranslationZ) r.m[2][3] = Lerp(le, it.startT.Tz, it.endT.Tz);
		return r;
	}

	if (it.hasRotation) {
		const Quat q = Slerp(le, it.startQ, it.endQ);
		QuatToMatrix(q, r.m);
	} else
		r = it.startT.R;

	if (it.hasScale) {
		const float Sx = Lerp(le, it.startT.Sx, it.endT.Sx);
		const float Sy = Lerp(le, it.startT.Sy, it.endT.Sy);
		const float Sz = Lerp(le, it.startT.Sz, it.endT.Sz);
		for (int j = 0; j < 3; ++j) {
			r.m[0][j] = Sx * r.m[0][j];
			r.m[1][j] = Sy * r.m[1][j];
			r.m[2][j] = Sz * r.m[2][j];
		}
	} else {
		for (int j = 0; j < 3; ++j) {
			r.m[0][j] = it.startT.Sx * r.m[0][j];
			r.m[1][j] = it.startT.Sy * r.m[1][j];
			r.m[2][j] = it.startT.Sz * r.m[2][j];
		}
	}
	r.m[0][3] = it.hasTranslationX ? Lerp(le, it.startT.Tx, it.endT.Tx) : it.startT.Tx;
	r.m[1][3] = it.hasTranslationY ? Lerp(le, it.startT.Ty, it.endT.Ty) : it.startT.Ty;
	r.m[2][3] = it.hasTranslationZ ? Lerp(le, it.startT.Tz, it.endT.Tz) : it.startT.Tz;
	return r;
}

struct MotionSys {      // motionsystem.h:156-193
	std::vector<float> times;
	std::vector<InterpT> its;   // interpolatedTransforms (the inverse list is not used by traversal)

	void Init(const std::vector<float> &t, const std::vector<Xform> &x) {   // motionsystem.cpp:296-331
		times = t;
		its.clear();
		size_t prev = 0;
		for (size_t i = 0; i < times.size(); ++i) {
			its.push_back(MakeInterp(times[prev], times[i], x[prev], x[i]));
			prev = i;
		}
		its.push_back(MakeInterp(times[prev], times[prev], x[prev], x[prev]));
	}

	M44 Sample(float time) const {      // motionsystem.cpp:340-346
		size_t index = std::upper_bound(times.begin(), times.end(), time) - times.begin();
		index = std::min(index, times.size() - 1);
		return SampleInterp(its[index], time);
	}

	Box Bound(const Box &ibox, bool storingGlobal2Local) const {    // motionsystem.cpp:75-87,356-365
		Box result;
		for (size_t k = 0; k < its.size(); ++k) {
			Box tbox;
			const float N = 1024.f;
			for (float i = 0; i <= N; ++i) {
				const float t = Lerp(i / N, its[k].startTime, its[k].endTime);
				M44 m = SampleInterp(its[k], t);
				if (storingGlobal2Local)
					m = InverseOrThrow(m);
				tbox = UnionB(tbox, XfBox(m, ibox));
			}
			result = UnionB(result, tbox);
		}
		return result;
	}
};

//------------------------------------------------------------------------------------------------
// Meshes / DataSet (trianglemesh.h/.cpp, dataset.cpp:66-81)
//------------------------------------------------------------------------------------------------

enum MeshKind { MESH_PLAIN = 0, MESH_INSTANCE = 1, MESH_MOTION = 2 };

// A TriangleMesh: the vertex / triangle arrays (trianglemesh.h:79-150).  Instances and motion
// meshes point at one; several dataset entries may share it.
struct Shape {
	std::vector<V3> verts;
	std::vector<u32> tris;      // 3 per triangle
};

// One DataSet entry (what DataSet::Add received)
struct Mesh {
	MeshKind kind;
	const Shape *shape;         // TriangleMesh / GetTriangleMesh()
	Xform trans;                // instance local->world
	MotionSys motion;           // motion: stores world->local matrices (parseobjects.cpp:155-157)
};

struct Scene {
	std::vector<Shape *> shapes;
	std::vector<Mesh *> meshes;     // dataset order == meshIndex
	~Scene() {
		for (size_t i = 0; i < meshes.size(); ++i) delete meshes[i];
		for (size_t i = 0; i < shapes.size(); ++i) delete shapes[i];
	}

	const Shape &Base(const Mesh &m) const { return *m.shape; }
	u32 VertCount(const Mesh &m) const { return (u32)m.shape->verts.size(); }
	u32 TriCount(const Mesh &m) const { return (u32)(m.shape->tris.size() / 3); }

	// Mesh::GetVertex(Transform::TRANS_IDENTITY, i)  (trianglemesh.h:96,214-216,319-321)
	V3 GetVertex(const Mesh &m, u32 i) const {
		const V3 v = m.shape->verts[i];
		switch (m.kind) {
			case MESH_INSTANCE: return XfPoint(m.trans.m, v);
			case MESH_MOTION: return XfPoint(Identity(), v);   // local2World (= identity) * v
			default: return v;
		}
	}

	Box GetBBox(const Mesh &m) const {  // trianglemesh.cpp:72-83,241-248,268-275
		const Shape &b = *m.shape;
		Box bb;
		for (size_t i = 0; i < b.verts.size(); ++i)
			bb = UnionP(bb, b.verts[i]);
		switch (m.kind) {
			case MESH_INSTANCE: return XfBox(m.trans.m, bb);
			case MESH_MOTION: return m.motion.Bound(bb, true);
			default: return bb;
		}
	}
};

//------------------------------------------------------------------------------------------------
// CLASSIC builder (bvhclassicbuild.cpp)
//------------------------------------------------------------------------------------------------

struct Params { u32 treeType; int costSamples, isectCost, traversalCost; float emptyBonus; };

struct TreeNode {       // BVHTreeNode, bvhbuild.h:49-64
	Box bbox;
	u32 a, b, c, d;     // triangleLeaf{meshIndex,triangleIndex} or bvhLeaf{leafIndex,transformIndex,motionIndex,meshOffsetIndex}
	TreeNode *leftChild, *rightSibling;
};

static void FindBestSplit(const Params &params, std::vector<TreeNode *> &list, u32 begin, u32 end,
		float *splitValue, u32 *bestAxis) {     // :51-119
	if (end - begin == 2) {
		*splitValue = (list[begin]->bbox.hi.x + list[begin]->bbox.lo.x +
				list[end - 1]->bbox.hi.x + list[end - 1]->bbox.lo.x) / 2;
		*bestAxis = 0;
		return;
	}
	V3 mean2 = v3(0, 0, 0), var = v3(0, 0, 0);
	for (u32 i = begin; i < end; i++) {
		const Box &b = list[i]->bbox;
		mean2 = v3(mean2.x + (b.hi.x + b.lo.x), mean2.y + (b.hi.y + b.lo.y), mean2.z + (b.hi.z + b.lo.z));
	}
	{
		const float inv = 1.f / static_cast<float>(end - begin);    // Point::operator/=
		mean2 = v3(mean2.x * inv, mean2.y * inv, mean2.z * inv);
	}
	for (u32 i = begin; i < end; i++) {
		const Box &b = list[i]->bbox;
		V3 v = v3((b.hi.x + b.lo.x) - mean2.x, (b.hi.y + b.lo.y) - mean2.y, (b.hi.z + b.lo.z) - mean2.z);
		v.x *= v.x; v.y *= v.y; v.z *= v.z;
		var = v3(var.x + v.x, var.y + v.y, var.z + v.z);
	}
	if (var.x > var.y && var.x > var.z) *bestAxis = 0;
	else if (var.y > var.z) *bestAxis = 1;
	else *bestAxis = 2;

	if (params.costSamples > 1) {
		Box nodeBounds;
		for (u32 i = begin; i < end; i++)
			nodeBounds = UnionB(nodeBounds, list[i]->bbox);
		const V3 d = sub(nodeBounds.hi, nodeBounds.lo);
		const float invTotalSA = 1.f / SurfaceArea(nodeBounds);
		const int ax = (int)*bestAxis;
		const float increment = 2 * Axis(d, ax) / (params.costSamples + 1);
		float bestCost = INFINITY;
		for (float splitVal = 2 * Axis(nodeBounds.lo, ax) + increment; splitVal < 2 * Axis(nodeBounds.hi, ax); splitVal += increment) {
			int nBelow = 0, nAbove = 0;
			Box bbBelow, bbAbove;
			for (u32 j = begin; j < end; j++) {
				if ((Axis(list[j]->bbox.hi, ax) + Axis(list[j]->bbox.lo, ax)) < splitVal) {
					nBelow++;
					bbBelow = UnionB(bbBelow, list[j]->bbox);
				} else {
					nAbove++;
					bbAbove = UnionB(bbAbove, list[j]->bbox);
				}
			}
			const float pBelow = SurfaceArea(bbBelow) * invTotalSA;
			const float pAbove = SurfaceArea(bbAbove) * invTotalSA;
			const float eb = (nAbove == 0 || nBelow == 0) ? params.emptyBonus : 0.f;
			const float cost = params.traversalCost + params.isectCost * (1.f - eb) * (pBelow * nBelow + pAbove * nAbove);
			if (cost < bestCost) {
				bestCost = cost;
				*splitValue = splitVal;
			}
		}
	} else
		*splitValue = Axis(mean2, (int)*bestAxis);
}

static TreeNode *BuildTree(u32 *nNodes, const Params &params, std::vector<TreeNode *> &leafList,
		u32 begin, u32 end, u32 axis) {      // :121-175
	u32 splitAxis = axis;
	float splitValue = 0.f;

	*nNodes += 1;
	if (end - begin == 1) {
		TreeNode *node = new TreeNode(*leafList[begin]);
		return node;
	}

	TreeNode *parent = new TreeNode();
	parent->leftChild = NULL;
	parent->rightSibling = NULL;

	std::vector<u32> splits;
	splits.reserve(params.treeType + 1);
	splits.push_back(begin);
	splits.push_back(end);
	for (u32 i = 2; i <= params.treeType; i *= 2) {
		// NB: `j` is unsigned in the reference; j-- at j == 0 wraps and the loop's j += 2 brings it back to 1
		for (u32 j = 0, offset = 0; j + offset < i && splits.size() > j + 1; j += 2) {
			if (splits[j + 1] - splits[j] < 2) {
				j--;
				offset++;
				continue;
			}
			FindBestSplit(params, leafList, splits[j], splits[j + 1], &splitValue, &splitAxis);
			const u32 ax = splitAxis;
			const float sv = splitValue;
			std::vector<TreeNode *>::iterator it = std::partition(leafList.begin() + splits[j], leafList.begin() + splits[j + 1],
					[ax, sv](TreeNode *n) { return Axis(n->bbox.hi, (int)ax) + Axis(n->bbox.lo, (int)ax) < sv; });
			u32 middle = (u32)std::distance(leafList.begin(), it);
			middle = std::max(splits[j] + 1, std::min(splits[j + 1] - 1, middle));
			splits.insert(splits.begin() + j + 1, middle);
		}
	}

	TreeNode *child = BuildTree(nNodes, params, leafList, splits[0], splits[1], splitAxis);
	parent->leftChild = child;
	parent->bbox = child->bbox;
	TreeNode *lastChild = child;
	for (u32 i = 1; i < splits.size() - 1; i++) {
		child = BuildTree(nNodes, params, leafList, splits[i], splits[i + 1], splitAxis);
		lastChild->rightSibling = child;
		parent->bbox = UnionB(parent->bbox, child->bbox);
		lastChild = child;
	}
	return parent;
}

static void FreeTree(TreeNode *n) {
	while (n) {
		TreeNode *next = n->rightSibling;
		FreeTree(n->leftChild);
		delete n;
		n = next;
	}
}

// :181-220.  `scene` != NULL: BVH of triangles; NULL: BVH of BVHs (MBVH root).
static u32 Flatten(const Scene *scene, const std::vector<const Mesh *> *meshes, TreeNode *node, u32 offset, Node *out) {
	while (node) {
		Node *an = &out[offset];
		memset(an, 0, sizeof(Node));
		if (node->leftChild) {
			an->box.bmin[0] = node->bbox.lo.x; an->box.bmin[1] = node->bbox.lo.y; an->box.bmin[2] = node->bbox.lo.z;
			an->box.bmax[0] = node->bbox.hi.x; an->box.bmax[1] = node->bbox.hi.y; an->box.bmax[2] = node->bbox.hi.z;
			offset = Flatten(scene, meshes, node->leftChild, offset + 1, out);
			an->nodeData = offset;
		} else {
			if (meshes) {
				const Shape &m = scene->Base(*(*meshes)[node->a]);
				an->tri.v[0] = m.tris[3 * node->b + 0];
				an->tri.v[1] = m.tris[3 * node->b + 1];
				an->tri.v[2] = m.tris[3 * node->b + 2];
				an->tri.meshIndex = node->a;
				an->tri.triangleIndex = node->b;
			} else {
				an->inst.leafIndex = node->a;
				an->inst.transformIndex = node->b;
				an->inst.motionIndex = node->c;
				an->inst.meshOffsetIndex = node->d;
			}
			++offset;
			an->nodeData = offset | 0x80000000u;
		}
		node = node->rightSibling;
	}
	return offset;
}

static std::vector<Node> BuildClassic(const Params &params, const Scene *scene, const std::vector<const Mesh *> *meshes,
		std::vector<TreeNode *> &leafList) {    // :222-233
	u32 nNodes = 0;
	TreeNode *root = BuildTree(&nNodes, params, leafList, 0, (u32)leafList.size(), 2);
	std::vector<Node> arr(nNodes);
	Flatten(scene, meshes, root, 0, arr.data());
	FreeTree(root);
	return arr;
}

static Params MakeParams(int treeType, int costSamples, int isectCost, int travCost, float emptyBonus) {    // bvhaccel.cpp:49-70
	Params p;
	if (treeType <= 2) p.treeType = 2;
	else if (treeType <= 4) p.treeType = 4;
	else p.treeType = 8;
	p.costSamples = costSamples;
	p.isectCost = isectCost;
	p.traversalCost = travCost;
	p.emptyBonus = emptyBonus;
	return p;
}

//------------------------------------------------------------------------------------------------
// BVHAccel (bvhaccel.cpp)
//------------------------------------------------------------------------------------------------

struct Counters { uint64_t inner, leaf, instLeaf, motionLeaf; };

struct BVH {
	const Scene *scene;
	std::vector<const Mesh *> meshes;
	std::vector<Node> nodes;    // bvhTree / nNodes
	u32 totalTris;

	void InitLeafList(std::vector<TreeNode> &bvNodes, std::vector<TreeNode *> &bvList) const {  // :96-134
		u32 total = 0;
		for (size_t m = 0; m < meshes.size(); ++m)
			total += scene->TriCount(*meshes[m]);
		bvNodes.resize(total);
		bvList.resize(total);
		u32 idx = 0;
		for (size_t m = 0; m < meshes.size(); ++m) {
			const Mesh &mesh = *meshes[m];
			const Shape &base = scene->Base(mesh);
			const u32 tc = scene->TriCount(mesh);
			for (u32 i = 0; i < tc; ++i, ++idx) {
				TreeNode *node = &bvNodes[idx];
				node->bbox = UnionP(BoxOf2(scene->GetVertex(mesh, base.tris[3 * i + 0]),
						scene->GetVertex(mesh, base.tris[3 * i + 1])),
						scene->GetVertex(mesh, base.tris[3 * i + 2]));
				ExpandBox(node->bbox, EpsB(node->bbox));
				node->a = (u32)m;
				node->b = i;
				node->c = node->d = 0;
				node->leftChild = NULL;
				node->rightSibling = NULL;
				bvList[idx] = node;
			}
		}
	}

	void Init(const Params &params) {   // :72-168 with builder CLASSIC
		totalTris = 0;
		for (size_t m = 0; m < meshes.size(); ++m)
			totalTris += scene->TriCount(*meshes[m]);
		nodes.clear();
		if (totalTris == 0)
			return;
		std::vector<TreeNode> bvNodes;
		std::vector<TreeNode *> bvList;
		InitLeafList(bvNodes, bvList);
		nodes = BuildClassic(params, scene, &meshes, bvList);
	}

	// :170-260
	bool Intersect(const Ray *initialRay, RayHit *rayHit, Counters *cnt) const {
		rayHit->t = initialRay->maxt;
		rayHit->meshIndex = NULL_INDEX;
		if (nodes.empty())
			return false;

		RayC ray;
		ray.o = v3(initialRay->o[0], initialRay->o[1], initialRay->o[2]);
		ray.d = v3(initialRay->d[0], initialRay->d[1], initialRay->d[2]);
		ray.mint = initialRay->mint; ray.maxt = initialRay->maxt; ray.time = initialRay->time;

		const Node *tree = nodes.data();
		u32 cur = 0;
		const u32 stop = Skip(tree[0].nodeData);
		float t, b1, b2;
		while (cur < stop) {
			const Node &node = tree[cur];
			const u32 nd = node.nodeData;
			if (IsLeaf(nd)) {
				if (cnt) cnt->leaf++;
				const Mesh &mesh = *meshes[node.tri.meshIndex];
				const V3 p0 = scene->GetVertex(mesh, node.tri.v[0]);
				const V3 p1 = scene->GetVertex(mesh, node.tri.v[1]);
				const V3 p2 = scene->GetVertex(mesh, node.tri.v[2]);
				if (TriangleIntersect(ray, p0, p1, p2, &t, &b1, &b2)) {
					if (t < rayHit->t) {
						ray.maxt = t;
						rayHit->t = t;
						rayHit->b1 = b1;
						rayHit->b2 = b2;
						rayHit->meshIndex = node.tri.meshIndex;
						rayHit->triangleIndex = node.tri.triangleIndex;
					}
				}
				++cur;
			} else {
				if (cnt) cnt->inner++;
				if (BoxIntersectP(ray, node.box.bmin, node.box.bmax))
					++cur;
				else
					cur = nd;
			}
		}
		return rayHit->meshIndex != NULL_INDEX;
	}
};

//------------------------------------------------------------------------------------------------
// MBVHAccel (mbvhaccel.cpp)
//------------------------------------------------------------------------------------------------

struct MBVH {
	const Scene *scene;
	Params params;
	std::vector<BVH *> uniqueLeafs;
	std::vector<Mesh *> leafMeshes;                 // owners of the one-mesh lists of the leaves
	std::vector<const Xform *> leafTransforms;      // uniqueLeafsTransform (live pointers, :141-142)
	std::vector<const MotionSys *> leafMotions;     // uniqueLeafsMotionSystem
	std::vector<TreeNode> bvhLeafs;
	std::vector<TreeNode *> bvhLeafsList;
	std::vector<Node> root;                         // bvhRootTree / nRootNodes

	~MBVH() {
		for (size_t i = 0; i < uniqueLeafs.size(); ++i) delete uniqueLeafs[i];
		for (size_t i = 0; i < leafMeshes.size(); ++i) delete leafMeshes[i];
	}

	void Init() {   // :58-216
		root.clear();
		u32 totalTris = 0;
		for (size_t i = 0; i < scene->meshes.size(); ++i)
			totalTris += scene->TriCount(*scene->meshes[i]);
		if (totalTris == 0)
			return;

		const u32 nLeafs = (u32)scene->meshes.size();
		std::vector<u32> leafsIndex, leafsTransformIndex, leafsMotionIndex;
		std::map<const Shape *, u32> uniqueLeafIndexByMesh;
		for (u32 i = 0; i < nLeafs; ++i) {
			const Mesh *mesh = scene->meshes[i];
			const Shape *base = mesh->shape;
			std::map<const Shape *, u32>::iterator it = uniqueLeafIndexByMesh.find(base);
			// a plain mesh always gets a fresh leaf (:100-112); instances/motion share (:118-139,148-169)
			if (mesh->kind == MESH_PLAIN || it == uniqueLeafIndexByMesh.end()) {
				Mesh *plain = new Mesh();       // the TriangleMesh itself, as a one-mesh list
				plain->kind = MESH_PLAIN;
				plain->shape = base;
				leafMeshes.push_back(plain);
				BVH *leaf = new BVH();
				leaf->scene = scene;
				leaf->meshes.assign(1, plain);
				leaf->Init(params);
				const u32 ui = (u32)uniqueLeafs.size();
				uniqueLeafIndexByMesh[base] = ui;
				uniqueLeafs.push_back(leaf);
				leafsIndex.push_back(ui);
			} else
				leafsIndex.push_back(it->second);

			if (mesh->kind == MESH_INSTANCE) {
				leafsTransformIndex.push_back((u32)leafTransforms.size());
				leafTransforms.push_back(&mesh->trans);
				leafsMotionIndex.push_back(NULL_INDEX);
			} else if (mesh->kind == MESH_MOTION) {
				leafsMotionIndex.push_back((u32)leafMotions.size());
				leafMotions.push_back(&mesh->motion);
				leafsTransformIndex.push_back(NULL_INDEX);
			} else {
				leafsTransformIndex.push_back(NULL_INDEX);
				leafsMotionIndex.push_back(NULL_INDEX);
			}
		}

		bvhLeafs.resize(nLeafs);
		bvhLeafsList.assign(nLeafs, NULL);
		for (u32 i = 0; i < nLeafs; ++i) {
			TreeNode *l = &bvhLeafs[i];
			l->bbox = scene->GetBBox(*scene->meshes[i]);
			ExpandBox(l->bbox, EpsB(l->bbox));
			l->a = leafsIndex[i];
			l->b = leafsTransformIndex[i];
			l->c = leafsMotionIndex[i];
			l->d = i;
			l->leftChild = NULL;
			l->rightSibling = NULL;
			bvhLeafsList[i] = l;
		}
		UpdateRoot();
	}

	void UpdateRoot() { root = BuildClassic(params, scene, NULL, bvhLeafsList); }   // :218-235

	void Update() {     // :237-250 -- NB: boxes are NOT re-expanded here
		for (size_t i = 0; i < bvhLeafs.size(); ++i)
			bvhLeafs[i].bbox = scene->GetBBox(*scene->meshes[i]);
		UpdateRoot();
	}

	// :252-357
	bool Intersect(const Ray *ray0, RayHit *rayHit, Counters *cnt) const {
		rayHit->t = ray0->maxt;
		rayHit->meshIndex = NULL_INDEX;
		if (root.empty())
			return false;

		RayC ray;
		ray.o = v3(ray0->o[0], ray0->o[1], ray0->o[2]);
		ray.d = v3(ray0->d[0], ray0->d[1], ray0->d[2]);
		ray.mint = ray0->mint; ray.maxt = ray0->maxt; ray.time = ray0->time;

		bool insideLeafTree = false;
		u32 currentRootNode = 0;
		const u32 rootStopNode = Skip(root[0].nodeData);
		u32 currentNode = currentRootNode;
		u32 currentStopNode = rootStopNode;
		u32 currentMeshOffset = 0;
		const Node *currentTree = root.data();
		RayC currentRay = ray;

		for (;;) {
			if (currentNode >= currentStopNode) {
				if (insideLeafTree) {
					currentTree = root.data();
					currentNode = currentRootNode;
					currentStopNode = rootStopNode;
					currentRay = ray;
					currentRay.maxt = rayHit->t;
					insideLeafTree = false;
					if (currentNode >= currentStopNode)
						break;
				} else
					break;
			}

			const Node &node = currentTree[currentNode];
			const u32 nd = node.nodeData;
			if (IsLeaf(nd)) {
				if (insideLeafTree) {
					if (cnt) cnt->leaf++;
					const u32 absoluteMeshIndex = node.tri.meshIndex + currentMeshOffset;
					const std::vector<V3> &vertices = scene->meshes[absoluteMeshIndex]->shape->verts;
					const V3 &p0 = vertices[node.tri.v[0]];
					const V3 &p1 = vertices[node.tri.v[1]];
					const V3 &p2 = vertices[node.tri.v[2]];
					float t, b1, b2;
					if (TriangleIntersect(currentRay, p0, p1, p2, &t, &b1, &b2)) {
						if (t < rayHit->t) {
							currentRay.maxt = t;
							rayHit->t = t;
							rayHit->b1 = b1;
							rayHit->b2 = b2;
							rayHit->meshIndex = absoluteMeshIndex;
							rayHit->triangleIndex = node.tri.triangleIndex;
						}
					}
					++currentNode;
				} else {
					currentTree = uniqueLeafs[node.inst.leafIndex]->nodes.data();
					if (node.inst.transformIndex != NULL_INDEX) {
						if (cnt) cnt->instLeaf++;
						currentRay = XfRay(leafTransforms[node.inst.transformIndex]->mInv, ray);
					} else if (node.inst.motionIndex != NULL_INDEX) {
						if (cnt) cnt->motionLeaf++;
						currentRay = XfRay(leafMotions[node.inst.motionIndex]->Sample(ray.time), ray);
					} else
						currentRay = ray;
					currentRay.maxt = rayHit->t;
					currentMeshOffset = node.inst.meshOffsetIndex;
					currentRootNode = currentNode + 1;
					currentNode = 0;
					currentStopNode = Skip(currentTree[0].nodeData);
					insideLeafTree = true;
				}
			} else {
				if (cnt) cnt->inner++;
				if (BoxIntersectP(currentRay, node.box.bmin, node.box.bmax))
					++currentNode;
				else
					currentNode = nd;
			}
		}
		return rayHit->meshIndex != NULL_INDEX;
	}
};

//------------------------------------------------------------------------------------------------
// Topology-free pins: brute force over every triangle with the same triangle test and the same
// "strictly closer wins" rule, visiting triangles in a caller-chosen order.
//------------------------------------------------------------------------------------------------

// Single level: triangles in dataset order (mesh 0 tri 0, 1, ...).  Matches BVHAccel::Intersect
// exactly whenever no two candidate hits tie in t and the BVH culling is geometrically correct.
static void BruteBVH(const Scene &scene, const Ray *r, RayHit *hit, float *secondT) {
	hit->t = r->maxt;
	hit->meshIndex = NULL_INDEX;
	RayC ray;
	ray.o = v3(r->o[0], r->o[1], r->o[2]);
	ray.d = v3(r->d[0], r->d[1], r->d[2]);
	ray.mint = r->mint; ray.maxt = r->maxt; ray.time = r->time;
	float second = std::numeric_limits<float>::infinity();
	for (size_t m = 0; m < scene.meshes.size(); ++m) {
		const Mesh &mesh = *scene.meshes[m];
		const Shape &base = scene.Base(mesh);
		const u32 tc = scene.TriCount(mesh);
		for (u32 i = 0; i < tc; ++i) {
			float t, b1, b2;
			if (TriangleIntersect(ray, scene.GetVertex(mesh, base.tris[3 * i]), scene.GetVertex(mesh, base.tris[3 * i + 1]),
					scene.GetVertex(mesh, base.tris[3 * i + 2]), &t, &b1, &b2)) {
				if (t < hit->t) {
					if (hit->meshIndex != NULL_INDEX) second = fminr(second, hit->t);
					hit->t = t; hit->b1 = b1; hit->b2 = b2;
					hit->meshIndex = (u32)m; hit->triangleIndex = i;
				} else
					second = fminr(second, t);
			}
		}
	}
	if (secondT) *secondT = second;
}

// Two level: every mesh in dataset order, ray moved into the mesh's space exactly as MBVH does.
static void BruteMBVH(const Scene &scene, const Ray *r, RayHit *hit, float *secondT) {
	hit->t = r->maxt;
	hit->meshIndex = NULL_INDEX;
	RayC ray;
	ray.o = v3(r->o[0], r->o[1], r->o[2]);
	ray.d = v3(r->d[0], r->d[1], r->d[2]);
	ray.mint = r->mint; ray.maxt = r->maxt; ray.time = r->time;
	float second = std::numeric_limits<float>::infinity();
	for (size_t m = 0; m < scene.meshes.size(); ++m) {
		const Mesh &mesh = *scene.meshes[m];
		const Shape &base = scene.Base(mesh);
		RayC lr = ray;
		if (mesh.kind == MESH_INSTANCE) lr = XfRay(mesh.trans.mInv, ray);
		else if (mesh.kind == MESH_MOTION) lr = XfRay(mesh.motion.Sample(ray.time), ray);
		const u32 tc = scene.TriCount(mesh);
		for (u32 i = 0; i < tc; ++i) {
			float t, b1, b2;
			if (TriangleIntersect(lr, base.verts[base.tris[3 * i]], base.verts[base.tris[3 * i + 1]],
					base.verts[base.tris[3 * i + 2]], &t, &b1, &b2)) {
				if (t < hit->t) {
					if (hit->meshIndex != NULL_INDEX) second = fminr(second, hit->t);
					hit->t = t; hit->b1 = b1; hit->b2 = b2;
					hit->meshIndex = (u32)m; hit->triangleIndex = i;
				} else
					second = fminr(second, t);
			}
		}
	}
	if (secondT) *secondT = second;
}

template <class F> static void ParallelFor(uint64_t n, int nthreads, F f) {
	if (nthreads <= 1 || n < 1024) { f(0, n, 0); return; }
	std::vector<std::thread> th;
	const uint64_t chunk = (n + nthreads - 1) / nthreads;
	for (int i = 0; i < nthreads; ++i) {
		const uint64_t b = std::min<uint64_t>(n, chunk * i), e = std::min<uint64_t>(n, b + chunk);
		th.emplace_back([=]() { f(b, e, i); });
	}
	for (size_t i = 0; i < th.size(); ++i) th[i].join();
}

}   // namespace orc

//================================================================================================
// C API (ctypes / test drivers)
//================================================================================================

using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (const std::exception &e) { g_err = e.what(); return -1; } return 0;

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void orc_set_epsilon(float mn, float mx) { g_epsMin = mn; g_epsMax = mx; }
float orc_epsilon(float v) { return EpsF(v); }

int orc_matrix_inverse(const float *m16, float *out16) {
	M44 a, r;
	memcpy(a.m, m16, 64);
	if (!Inverse(a, &r)) return -1;
	memcpy(out16, r.m, 64);
	return 0;
}

int orc_triangle_intersect(const Ray *r, const float *p0, const float *p1, const float *p2, float *tb) {
	RayC ray;
	ray.o = v3(r->o[0], r->o[1], r->o[2]); ray.d = v3(r->d[0], r->d[1], r->d[2]);
	ray.mint = r->mint; ray.maxt = r->maxt; ray.time = r->time;
	return TriangleIntersect(ray, v3(p0[0], p0[1], p0[2]), v3(p1[0], p1[1], p1[2]), v3(p2[0], p2[1], p2[2]), &tb[0], &tb[1], &tb[2]) ? 1 : 0;
}

int orc_bbox_intersectp(const Ray *r, const float *bmin, const float *bmax) {
	RayC ray;
	ray.o = v3(r->o[0], r->o[1], r->o[2]); ray.d = v3(r->d[0], r->d[1], r->d[2]);
	ray.mint = r->mint; ray.maxt = r->maxt; ray.time = r->time;
	return BoxIntersectP(ray, bmin, bmax) ? 1 : 0;
}

//---- scene ----

void *orc_scene_create() { return new Scene(); }
void orc_scene_free(void *s) { delete (Scene *)s; }
int orc_scene_mesh_count(void *s) { return (int)((Scene *)s)->meshes.size(); }

// a TriangleMesh (geometry only; not yet part of the dataset)
int orc_scene_add_shape(void *s, const float *xyz, uint32_t nVerts, const uint32_t *tris, uint32_t nTris) {
	Scene *sc = (Scene *)s;
	Shape *m = new Shape();
	m->verts.resize(nVerts);
	for (uint32_t i = 0; i < nVerts; ++i) m->verts[i] = v3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
	m->tris.assign(tris, tris + 3 * (size_t)nTris);
	sc->shapes.push_back(m);
	return (int)sc->shapes.size() - 1;
}

static Mesh *NewMesh(Scene *sc, int shape, MeshKind kind) {
	if (shape < 0 || shape >= (int)sc->shapes.size())
		throw std::runtime_error("bad shape index");
	Mesh *m = new Mesh();
	m->kind = kind;
	m->shape = sc->shapes[shape];
	m->trans.m = Identity();
	m->trans.mInv = Identity();
	return m;
}

// DataSet::Add(TriangleMesh)
int orc_scene_add_plain(void *s, int shape) {
	Scene *sc = (Scene *)s;
	ORC_TRY
	sc->meshes.push_back(NewMesh(sc, shape, MESH_PLAIN));
	ORC_CATCH
}

// DataSet::Add(InstanceTriangleMesh(shape, Transform(m16))); m16 row-major local->world
int orc_scene_add_instance(void *s, int shape, const float *m16) {
	Scene *sc = (Scene *)s;
	ORC_TRY
	Mesh *m = NewMesh(sc, shape, MESH_INSTANCE);
	M44 a; memcpy(a.m, m16, 64);
	try { m->trans = MakeXform(a); } catch (...) { delete m; throw; }
	sc->meshes.push_back(m);
	ORC_CATCH
}

// DataSet::Add(MotionTriangleMesh(shape, MotionSystem(times, Transform(m16s[i])))).  The matrices
// are what the MotionSystem stores, i.e. what MBVH multiplies the ray by (world->local for scene
// objects).
int orc_scene_add_motion(void *s, int shape, uint32_t nKeys, const float *times, const float *m16s) {
	Scene *sc = (Scene *)s;
	ORC_TRY
	Mesh *m = NewMesh(sc, shape, MESH_MOTION);
	try {
		std::vector<float> t(times, times + nKeys);
		std::vector<Xform> x;
		for (uint32_t i = 0; i < nKeys; ++i) { M44 a; memcpy(a.m, m16s + 16 * i, 64); x.push_back(MakeXform(a)); }
		m->motion.Init(t, x);
	} catch (...) { delete m; throw; }
	sc->meshes.push_back(m);
	ORC_CATCH
}

// live edit of an instance transform (MBVHAccel keeps pointers, so Update() sees it)
int orc_scene_set_instance_transform(void *s, int mesh, const float *m16) {
	Scene *sc = (Scene *)s;
	ORC_TRY
	if (mesh < 0 || mesh >= (int)sc->meshes.size() || sc->meshes[mesh]->kind != MESH_INSTANCE)
		throw std::runtime_error("not an instance");
	M44 a; memcpy(a.m, m16, 64);
	sc->meshes[mesh]->trans = MakeXform(a);
	ORC_CATCH
}

int orc_scene_mesh_bbox(void *s, int mesh, float *out6) {
	Scene *sc = (Scene *)s;
	ORC_TRY
	const Box b = sc->GetBBox(*sc->meshes[mesh]);
	out6[0] = b.lo.x; out6[1] = b.lo.y; out6[2] = b.lo.z; out6[3] = b.hi.x; out6[4] = b.hi.y; out6[5] = b.hi.z;
	ORC_CATCH
}

// flattened (GetVertex) vertices of one mesh, as BVHKernel uploads them (bvhaccelhw.cpp:78-92)
int orc_scene_mesh_world_vertices(void *s, int mesh, float *outXYZ) {
	Scene *sc = (Scene *)s;
	const Mesh &m = *sc->meshes[mesh];
	const uint32_t n = sc->VertCount(m);
	for (uint32_t i = 0; i < n; ++i) {
		const V3 v = sc->GetVertex(m, i);
		outXYZ[3 * i] = v.x; outXYZ[3 * i + 1] = v.y; outXYZ[3 * i + 2] = v.z;
	}
	return (int)n;
}

//---- BVH ----

void *orc_bvh_build(void *s, int treeType, int costSamples, int isectCost, int travCost, float emptyBonus) {
	Scene *sc = (Scene *)s;
	try {
		BVH *b = new BVH();
		b->scene = sc;
		for (size_t i = 0; i < sc->meshes.size(); ++i) b->meshes.push_back(sc->meshes[i]);
		b->Init(MakeParams(treeType, costSamples, isectCost, travCost, emptyBonus));
		return b;
	} catch (const std::exception &e) { g_err = e.what(); return NULL; }
}

// wrap an externally built BVHArrayNode[] (e.g. from the product's host-layer builder) so the
// oracle's Intersect can walk the SAME tree the device was given
void *orc_bvh_from_nodes(void *s, const void *nodes, uint32_t nNodes) {
	Scene *sc = (Scene *)s;
	BVH *b = new BVH();
	b->scene = sc;
	for (size_t i = 0; i < sc->meshes.size(); ++i) b->meshes.push_back(sc->meshes[i]);
	b->nodes.assign((const Node *)nodes, (const Node *)nodes + nNodes);
	b->totalTris = 0;
	return b;
}

void orc_bvh_free(void *b) { delete (BVH *)b; }
uint32_t orc_bvh_node_count(void *b) { return (uint32_t)((BVH *)b)->nodes.size(); }
const void *orc_bvh_nodes(void *b) { return ((BVH *)b)->nodes.data(); }

// counters4: inner, leaf, instLeaf, motionLeaf visit totals (may be NULL)
int orc_bvh_intersect(void *bp, const Ray *rays, RayHit *hits, uint64_t n, int nthreads, uint64_t *counters4) {
	const BVH *b = (const BVH *)bp;
	std::vector<Counters> cs(std::max(1, nthreads));
	for (size_t i = 0; i < cs.size(); ++i) memset(&cs[i], 0, sizeof(Counters));
	const bool count = counters4 != NULL;
	ParallelFor(n, nthreads, [&](uint64_t lo, uint64_t hi, int tid) {
		Counters local; memset(&local, 0, sizeof(local));
		for (uint64_t i = lo; i < hi; ++i)
			b->Intersect(&rays[i], &hits[i], count ? &local : NULL);
		cs[tid] = local;
	});
	if (counters4) {
		counters4[0] = counters4[1] = counters4[2] = counters4[3] = 0;
		for (size_t i = 0; i < cs.size(); ++i) {
			counters4[0] += cs[i].inner; counters4[1] += cs[i].leaf;
			counters4[2] += cs[i].instLeaf; counters4[3] += cs[i].motionLeaf;
		}
	}
	return 0;
}

//---- MBVH ----

void *orc_mbvh_build(void *s, int treeType, int costSamples, int isectCost, int travCost, float emptyBonus) {
	Scene *sc = (Scene *)s;
	try {
		MBVH *m = new MBVH();
		m->scene = sc;
		m->params = MakeParams(treeType, costSamples, isectCost, travCost, emptyBonus);
		m->Init();
		return m;
	} catch (const std::exception &e) { g_err = e.what(); return NULL; }
}
void orc_mbvh_free(void *m) { delete (MBVH *)m; }
int orc_mbvh_update(void *m) {
	ORC_TRY
	((MBVH *)m)->Update();
	ORC_CATCH
}
uint32_t orc_mbvh_root_node_count(void *m) { return (uint32_t)((MBVH *)m)->root.size(); }
const void *orc_mbvh_root_nodes(void *m) { return ((MBVH *)m)->root.data(); }
// replace the root tree with an externally built one (same leaf payload convention)
void orc_mbvh_set_root_nodes(void *m, const void *nodes, uint32_t n) {
	((MBVH *)m)->root.assign((const Node *)nodes, (const Node *)nodes + n);
}
uint32_t orc_mbvh_leaf_count(void *m) { return (uint32_t)((MBVH *)m)->uniqueLeafs.size(); }
uint32_t orc_mbvh_leaf_node_count(void *m, uint32_t i) { return (uint32_t)((MBVH *)m)->uniqueLeafs[i]->nodes.size(); }
const void *orc_mbvh_leaf_nodes(void *m, uint32_t i) { return ((MBVH *)m)->uniqueLeafs[i]->nodes.data(); }
void orc_mbvh_set_leaf_nodes(void *m, uint32_t i, const void *nodes, uint32_t n) {
	((MBVH *)m)->uniqueLeafs[i]->nodes.assign((const Node *)nodes, (const Node *)nodes + n);
}
// index of the shape (TriangleMesh) unique leaf i was built over
int orc_mbvh_leaf_mesh(void *mp, uint32_t i) {
	MBVH *m = (MBVH *)mp;
	const Shape *base = m->uniqueLeafs[i]->meshes[0]->shape;
	for (size_t k = 0; k < m->scene->shapes.size(); ++k)
		if (m->scene->shapes[k] == base) return (int)k;
	return -1;
}
uint32_t orc_mbvh_transform_count(void *m) { return (uint32_t)((MBVH *)m)->leafTransforms.size(); }
// mInv of uniqueLeafsTransform[i], row-major (what MBVHKernel uploads, mbvhaccelhw.cpp:141-152)
void orc_mbvh_transform_minv(void *m, uint32_t i, float *out16) { memcpy(out16, ((MBVH *)m)->leafTransforms[i]->mInv.m, 64); }
uint32_t orc_mbvh_motion_count(void *m) { return (uint32_t)((MBVH *)m)->leafMotions.size(); }
uint32_t orc_mbvh_motion_interp_count(void *m, uint32_t i) { return (uint32_t)((MBVH *)m)->leafMotions[i]->its.size(); }
// 576-byte ocl::InterpolatedTransform records of motion system i
const void *orc_mbvh_motion_interps(void *m, uint32_t i) { return ((MBVH *)m)->leafMotions[i]->its.data(); }
int orc_motion_sample(void *mp, uint32_t i, float time, float *out16) {
	const M44 r = ((MBVH *)mp)->leafMotions[i]->Sample(time);
	memcpy(out16, r.m, 64);
	return 0;
}

int orc_mbvh_intersect(void *mp, const Ray *rays, RayHit *hits, uint64_t n, int nthreads, uint64_t *counters4) {
	const MBVH *m = (const MBVH *)mp;
	std::vector<Counters> cs(std::max(1, nthreads));
	for (size_t i = 0; i < cs.size(); ++i) memset(&cs[i], 0, sizeof(Counters));
	const bool count = counters4 != NULL;
	ParallelFor(n, nthreads, [&](uint64_t lo, uint64_t hi, int tid) {
		Counters local; memset(&local, 0, sizeof(local));
		for (uint64_t i = lo; i < hi; ++i)
			m->Intersect(&rays[i], &hits[i], count ? &local : NULL);
		cs[tid] = local;
	});
	if (counters4) {
		counters4[0] = counters4[1] = counters4[2] = counters4[3] = 0;
		for (size_t i = 0; i < cs.size(); ++i) {
			counters4[0] += cs[i].inner; counters4[1] += cs[i].leaf;
			counters4[2] += cs[i].instLeaf; counters4[3] += cs[i].motionLeaf;
		}
	}
	return 0;
}

//---- brute force ----

// twoLevel = 0: world-space flattened meshes (BVHAccel semantics); 1: per-mesh ray transform (MBVH)
int orc_brute(void *s, int twoLevel, const Ray *rays, RayHit *hits, float *secondT, uint64_t n, int nthreads) {
	const Scene *sc = (const Scene *)s;
	ParallelFor(n, nthreads, [&](uint64_t lo, uint64_t hi, int) {
		for (uint64_t i = lo; i < hi; ++i) {
			if (twoLevel) BruteMBVH(*sc, &rays[i], &hits[i], secondT ? &secondT[i] : NULL);
			else BruteBVH(*sc, &rays[i], &hits[i], secondT ? &secondT[i] : NULL);
		}
	});
	return 0;
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}   // extern "C"
