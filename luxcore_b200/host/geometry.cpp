// geometry.cpp -- out-of-line parts of include/luxrays/core/geometry.h.
// Restates (same operation order, so results are bit-identical with the reference build):
//   Matrix4x4::Inverse/Determinant/Transpose   src/luxrays/core/geometry/matrix4x4.cpp:63-175
//   Quaternion(Matrix4x4), ToMatrix, Slerp     src/luxrays/core/geometry/quaternion.cpp:27-164
//   InterpolatedTransform / DecomposedTransform / MotionSystem
//                                              src/luxrays/core/geometry/motionsystem.cpp:37-159,168-276,296-365
#include <algorithm>

#include "luxrays/core/geometry.h"

namespace luxrays {

float MachineEpsilon::minEpsilon = 1e-5f;   // DEFAULT_EPSILON_MIN
float MachineEpsilon::maxEpsilon = 1e-1f;   // DEFAULT_EPSILON_MAX

const Matrix4x4 Matrix4x4::MAT_IDENTITY = Matrix4x4();
const Transform Transform::TRANS_IDENTITY = Transform();

Matrix4x4 Matrix4x4::Transpose() const {
	Matrix4x4 r;
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			r.m[i][j] = m[j][i];
	return r;
}

static float Minor2(float a, float b, float c, float d) { return a * d - b * c; }

static float Det3(const float A[3][3]) {
	return A[0][0] * Minor2(A[1][1], A[1][2], A[2][1], A[2][2]) -
			A[0][1] * Minor2(A[1][0], A[1][2], A[2][0], A[2][2]) +
			A[0][2] * Minor2(A[1][0], A[1][1], A[2][0], A[2][1]);
}

// expansion along the last row, skipping zero entries
float Matrix4x4::Determinant() const {
	float A[3][3];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			A[i][j] = m[i][j + 1];
	float det = 0.f, sign = -1.f;
	for (int k = 0; k < 4; ++k) {
		if (m[3][k] != 0.f)
			det += sign * m[3][k] * Det3(A);
		if (k == 3)
			break;
		sign *= -1.f;
		for (int i = 0; i < 3; ++i)
			A[i][k] = m[i][k];
	}
	return det;
}

// Gauss-Jordan elimination with full pivoting.  The pivot search uses `>=`, i.e. among equal
// magnitudes the LAST candidate wins -- this decides the rounding of the result and therefore
// has to match the reference exactly.
Matrix4x4 Matrix4x4::Inverse() const {
	float a[4][4];
	memcpy(a, m, sizeof(a));
	int pivRow[4], pivCol[4];
	int used[4] = { 0, 0, 0, 0 };
	for (int step = 0; step < 4; ++step) {
		int row = -1, col = -1;
		float best = 0.f;
		for (int j = 0; j < 4; ++j) {
			if (used[j] == 1)
				continue;
			for (int k = 0; k < 4; ++k) {
				if (used[k] == 0) {
					if (fabsf(a[j][k]) >= best) {
						best = fabsf(a[j][k]);
						row = j;
						col = k;
					}
				} else if (used[k] > 1)
					throw std::runtime_error("Singular matrix in MatrixInvert");
			}
		}
		++used[col];
		if (row != col)
			for (int k = 0; k < 4; ++k)
				Swap(a[row][k], a[col][k]);
		pivRow[step] = row;
		pivCol[step] = col;
		if (a[col][col] == 0.f)
			throw std::runtime_error("Singular matrix in MatrixInvert");
		const float pinv = 1.f / a[col][col];
		a[col][col] = 1.f;
		for (int j = 0; j < 4; ++j)
			a[col][j] *= pinv;
		for (int j = 0; j < 4; ++j) {
			if (j == col)
				continue;
			const float f = a[j][col];
			a[j][col] = 0;
			for (int k = 0; k < 4; ++k)
				a[j][k] -= a[col][k] * f;
		}
	}
	for (int j = 3; j >= 0; --j) {
		if (pivRow[j] != pivCol[j])
			for (int k = 0; k < 4; ++k)
				Swap(a[k][pivRow[j]], a[k][pivCol[j]]);
	}
	return Matrix4x4(a);
}

bool Transform::SwapsHandedness() const {
	const float det = ((m.m[0][0] * (m.m[1][1] * m.m[2][2] - m.m[1][2] * m.m[2][1])) -
			(m.m[0][1] * (m.m[1][0] * m.m[2][2] - m.m[1][2] * m.m[2][0])) +
			(m.m[0][2] * (m.m[1][0] * m.m[2][1] - m.m[1][1] * m.m[2][0])));
	return det < 0.f;
}

static Matrix4x4 Rows(float a, float b, float c, float d, float e, float f, float g, float h,
		float i, float j, float k, float l, float mm, float n, float o, float p) {
	const float v[16] = { a, b, c, d, e, f, g, h, i, j, k, l, mm, n, o, p };
	return Matrix4x4(v);
}

Transform Translate(const Vector &d) {
	return Transform(Rows(1, 0, 0, d.x, 0, 1, 0, d.y, 0, 0, 1, d.z, 0, 0, 0, 1),
			Rows(1, 0, 0, -d.x, 0, 1, 0, -d.y, 0, 0, 1, -d.z, 0, 0, 0, 1));
}

Transform Scale(float x, float y, float z) {
	return Transform(Rows(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1),
			Rows(1.f / x, 0, 0, 0, 0, 1.f / y, 0, 0, 0, 0, 1.f / z, 0, 0, 0, 0, 1));
}

static const float kDegToRad = 3.14159265358979323846f / 180.f;

Transform RotateX(float angle) {
	const float s = sinf(angle * kDegToRad), c = cosf(angle * kDegToRad);
	const Matrix4x4 m = Rows(1, 0, 0, 0, 0, c, -s, 0, 0, s, c, 0, 0, 0, 0, 1);
	return Transform(m, m.Transpose());
}

Transform RotateY(float angle) {
	const float s = sinf(angle * kDegToRad), c = cosf(angle * kDegToRad);
	const Matrix4x4 m = Rows(c, 0, s, 0, 0, 1, 0, 0, -s, 0, c, 0, 0, 0, 0, 1);
	return Transform(m, m.Transpose());
}

Transform RotateZ(float angle) {
	const float s = sinf(angle * kDegToRad), c = cosf(angle * kDegToRad);
	const Matrix4x4 m = Rows(c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1);
	return Transform(m, m.Transpose());
}

//------------------------------------------------------------------------------
// Quaternion
//------------------------------------------------------------------------------

static inline float InvLenOrOne(float x, float y, float z) {
	const float len = sqrtf(x * x + y * y + z * z);
	return (len == 0.f) ? 1.f : 1.f / len;
}

// Gram-Schmidt in the reference's order: normalise x, z = x cross y, y = z cross x
static void OrthoNormalise3(float r[3][3]) {
	float s = InvLenOrOne(r[0][0], r[0][1], r[0][2]);
	r[0][0] *= s; r[0][1] *= s; r[0][2] *= s;

	r[2][0] = (r[0][1] * r[1][2] - r[0][2] * r[1][1]);
	r[2][1] = (r[0][2] * r[1][0] - r[0][0] * r[1][2]);
	r[2][2] = (r[0][0] * r[1][1] - r[0][1] * r[1][0]);
	s = InvLenOrOne(r[2][0], r[2][1], r[2][2]);
	r[2][0] *= s; r[2][1] *= s; r[2][2] *= s;

	r[1][0] = (r[2][1] * r[0][2] - r[2][2] * r[0][1]);
	r[1][1] = (r[2][2] * r[0][0] - r[2][0] * r[0][2]);
	r[1][2] = (r[2][0] * r[0][1] - r[2][1] * r[0][0]);
	s = InvLenOrOne(r[1][0], r[1][1], r[1][2]);
	r[1][0] *= s; r[1][1] *= s; r[1][2] *= s;
}

Quaternion::Quaternion(const Matrix4x4 &mat) {
	float o[3][3];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			o[i][j] = mat.m[i][j];
	OrthoNormalise3(o);

	const float trace = o[0][0] + o[1][1] + o[2][2] + 1.f;
	if (trace > 1e-6f) {
		const float s = sqrtf(trace) * 2.f;
		v = Vector((o[1][2] - o[2][1]) / s, (o[2][0] - o[0][2]) / s, (o[0][1] - o[1][0]) / s);
		w = 0.25f * s;
	} else if (o[0][0] > o[1][1] && o[0][0] > o[2][2]) {
		const float s = sqrtf(1.f + o[0][0] - o[1][1] - o[2][2]) * 2.f;
		v = Vector(0.25f * s, (o[0][1] + o[1][0]) / s, (o[2][0] + o[0][2]) / s);
		w = (o[1][2] - o[2][1]) / s;
	} else if (o[1][1] > o[2][2]) {
		const float s = sqrtf(1.f + o[1][1] - o[0][0] - o[2][2]) * 2.f;
		v = Vector((o[0][1] + o[1][0]) / s, 0.25f * s, (o[1][2] + o[2][1]) / s);
		w = (o[2][0] - o[0][2]) / s;
	} else {
		const float s = sqrtf(1.f + o[2][2] - o[0][0] - o[1][1]) * 2.f;
		v = Vector((o[2][0] + o[0][2]) / s, (o[1][2] + o[2][1]) / s, 0.25f * s);
		w = (o[0][1] - o[1][0]) / s;
	}
}

void Quaternion::ToMatrix(float m[4][4]) const {
	const float xx = v.x * v.x, yy = v.y * v.y, zz = v.z * v.z;
	const float xy = v.x * v.y, xz = v.x * v.z, yz = v.y * v.z;
	const float xw = v.x * w, yw = v.y * w, zw = v.z * w;
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

Quaternion Slerp(float t, const Quaternion &q1, const Quaternion &q2) {
	float cosPhi = Dot(q1, q2);
	const float sign = (cosPhi > 0.f) ? 1.f : -1.f;
	cosPhi *= sign;
	float f1, f2;
	if (1.f - cosPhi > 1e-6f) {
		const float phi = acosf(cosPhi);
		const float sinPhi = sinf(phi);
		f1 = sinf((1.f - t) * phi) / sinPhi;
		f2 = sinf(t * phi) / sinPhi;
	} else {
		f1 = 1.f - t;
		f2 = t;
	}
	return f1 * q1 + (sign * f2) * q2;
}

//------------------------------------------------------------------------------
// InterpolatedTransform
//------------------------------------------------------------------------------

// "unmatrix" (Graphics Gems II) as the reference applies it
InterpolatedTransform::DecomposedTransform::DecomposedTransform(const Matrix4x4 &src) :
		Sx(0), Sy(0), Sz(0), Sxy(0), Sxz(0), Syz(0), R(src), Tx(0), Ty(0), Tz(0), Px(0), Py(0), Pz(0), Pw(0), Valid(false) {
	if (R.m[3][3] == 0)
		return;
	// in-place normalisation by the (changing) last element, exactly like the reference loop
	for (u_int i = 0; i < 4; ++i)
		for (u_int j = 0; j < 4; ++j)
			R.m[i][j] /= R.m[3][3];

	Matrix4x4 upper(R);
	upper.m[0][3] = upper.m[1][3] = upper.m[2][3] = 0.f;
	upper.m[3][3] = 1.f;
	if (upper.Determinant() == 0.f)
		return;

	if (R.m[3][0] != 0.f || R.m[3][1] != 0.f || R.m[3][2] != 0.f) {
		const float rhs[4] = { R.m[3][0], R.m[3][1], R.m[3][2], R.m[3][3] };
		const Matrix4x4 A = upper.Inverse().Transpose();
		float sol[4];
		for (int i = 0; i < 4; ++i)
			sol[i] = A.m[i][0] * rhs[0] + A.m[i][1] * rhs[1] + A.m[i][2] * rhs[2] + A.m[i][3] * rhs[3];
		Px = sol[0]; Py = sol[1]; Pz = sol[2]; Pw = sol[3];
		R.m[3][0] = R.m[3][1] = R.m[3][2] = 0.f;
		R.m[3][3] = 1.f;
	}

	Tx = R.m[0][3]; Ty = R.m[1][3]; Tz = R.m[2][3];
	R.m[0][3] = R.m[1][3] = R.m[2][3] = 0.f;

	Vector row[3];
	for (u_int i = 0; i < 3; ++i)
		row[i] = Vector(R.m[i][0], R.m[i][1], R.m[i][2]);

	Sx = row[0].Length();
	row[0] *= 1.f / Sx;
	Sxy = Dot(row[0], row[1]);
	row[1] -= Sxy * row[0];
	Sy = row[1].Length();
	row[1] *= 1.f / Sy;
	Sxy /= Sy;
	Sxz = Dot(row[0], row[2]);
	row[2] -= Sxz * row[0];
	Syz = Dot(row[1], row[2]);
	row[2] -= Syz * row[1];
	Sz = row[2].Length();
	row[2] *= 1.f / Sz;
	Sxz /= Sz;
	Syz /= Sz;

	if (Dot(row[0], Cross(row[1], row[2])) < 0.f) {
		Sx *= -1.f; Sy *= -1.f; Sz *= -1.f;
		for (u_int i = 0; i < 3; ++i)
			row[i] *= -1.f;
	}
	for (u_int i = 0; i < 3; ++i) {
		R.m[i][0] = row[i].x; R.m[i][1] = row[i].y; R.m[i][2] = row[i].z;
	}
	Valid = true;
}

InterpolatedTransform::InterpolatedTransform(float st, float et, const Transform &s, const Transform &e) {
	// keep the padding bytes defined: the object is uploaded byte-for-byte
	memset(static_cast<void *>(this), 0, sizeof(*this));
	startTime = st;
	endTime = et;
	start = s;
	end = e;
	startT = DecomposedTransform();
	endT = DecomposedTransform();
	startQ = Quaternion();
	endQ = Quaternion();
	InitFlags();
	if (startTime == endTime)
		return;

	startT = DecomposedTransform(start.m);
	endT = DecomposedTransform(end.m);
	if (!startT.Valid)
		throw std::runtime_error("Singular start matrix in InterpolatedTransform, interpolation disabled");
	if (!endT.Valid)
		throw std::runtime_error("Singular end matrix in InterpolatedTransform, interpolation disabled");

	startQ = Normalize(Quaternion(startT.R));
	endQ = Normalize(Quaternion(endT.R));

	hasTranslationX = startT.Tx != endT.Tx;
	hasTranslationY = startT.Ty != endT.Ty;
	hasTranslationZ = startT.Tz != endT.Tz;
	hasTranslation = hasTranslationX || hasTranslationY || hasTranslationZ;
	hasScaleX = startT.Sx != endT.Sx;
	hasScaleY = startT.Sy != endT.Sy;
	hasScaleZ = startT.Sz != endT.Sz;
	hasScale = hasScaleX || hasScaleY || hasScaleZ;
	hasRotation = fabsf(Dot(startQ, endQ) - 1.f) >= 1e-6f;
	isActive = hasTranslation || hasScale || hasRotation;
}

Matrix4x4 InterpolatedTransform::Sample(const float time) const {
	if (!isActive || time <= startTime)
		return start.m;
	if (time >= endTime)
		return end.m;

	const float le = (time - startTime) / (endTime - startTime);
	float im[4][4];

	if (hasTranslation && !(hasScale || hasRotation)) {
		memcpy(im, start.m.m, sizeof(im));
		if (hasTranslationX) im[0][3] = Lerp(le, startT.Tx, endT.Tx);
		if (hasTranslationY) im[1][3] = Lerp(le, startT.Ty, endT.Ty);
		if (hasTranslationZ) im[2][3] = Lerp(le, startT.Tz, endT.Tz);
		return Matrix4x4(im);
	}

	if (hasRotation)
		Slerp(le, startQ, endQ).ToMatrix(im);
	else
		memcpy(im, startT.R.m, sizeof(im));

	const float Sx = hasScale ? Lerp(le, startT.Sx, endT.Sx) : startT.Sx;
	const float Sy = hasScale ? Lerp(le, startT.Sy, endT.Sy) : startT.Sy;
	const float Sz = hasScale ? Lerp(le, startT.Sz, endT.Sz) : startT.Sz;
	for (u_int j = 0; j < 3; ++j) {
		im[0][j] = Sx * im[0][j];
		im[1][j] = Sy * im[1][j];
		im[2][j] = Sz * im[2][j];
	}
	im[0][3] = hasTranslationX ? Lerp(le, startT.Tx, endT.Tx) : startT.Tx;
	im[1][3] = hasTranslationY ? Lerp(le, startT.Ty, endT.Ty) : startT.Ty;
	im[2][3] = hasTranslationZ ? Lerp(le, startT.Tz, endT.Tz) : startT.Tz;
	return Matrix4x4(im);
}

// union over 1025 samples of the interval
BBox InterpolatedTransform::Bound(BBox ibox, const bool storingGlobal2Local) const {
	BBox tbox;
	const float N = 1024.f;
	for (float i = 0; i <= N; ++i) {
		const float t = Lerp(i / N, startTime, endTime);
		Matrix4x4 mm = Sample(t);
		if (storingGlobal2Local)
			mm = mm.Inverse();
		tbox = Union(tbox, mm * ibox);
	}
	return tbox;
}

//------------------------------------------------------------------------------
// MotionSystem
//------------------------------------------------------------------------------

MotionSystem::MotionSystem() : times(1, 0.f),
		interpolatedTransforms(1, InterpolatedTransform(0.f, 0.f, Transform(), Transform())),
		interpolatedInverseTransforms(1, InterpolatedTransform(0.f, 0.f, Transform(), Transform())) {
}

MotionSystem::MotionSystem(const Transform &t) : times(1, 0.f),
		interpolatedTransforms(1, InterpolatedTransform(0.f, 0.f, t, t)),
		interpolatedInverseTransforms(1, InterpolatedTransform(0.f, 0.f, Transform(Inverse(t)), Transform(Inverse(t)))) {
}

MotionSystem::MotionSystem(const std::vector<float> &t, const std::vector<Transform> &transforms) {
	Init(t, transforms);
}

// one interpolated segment per knot (the first one degenerate) plus a trailing static one
void MotionSystem::Init(const std::vector<float> &t, const std::vector<Transform> &xf) {
	times = t;
	interpolatedTransforms.clear();
	interpolatedInverseTransforms.clear();
	interpolatedTransforms.reserve(times.size() + 1);
	interpolatedInverseTransforms.reserve(times.size() + 1);
	size_t prev = 0;
	for (size_t i = 0; i < times.size(); ++i) {
		interpolatedTransforms.push_back(InterpolatedTransform(times[prev], times[i], xf[prev], xf[i]));
		interpolatedInverseTransforms.push_back(InterpolatedTransform(times[prev], times[i],
				Transform(Inverse(xf[prev])), Transform(Inverse(xf[i]))));
		prev = i;
	}
	interpolatedTransforms.push_back(InterpolatedTransform(times[prev], times[prev], xf[prev], xf[prev]));
	interpolatedInverseTransforms.push_back(InterpolatedTransform(times[prev], times[prev],
			Transform(Inverse(xf[prev])), Transform(Inverse(xf[prev]))));
}

Matrix4x4 MotionSystem::Sample(const float time) const {
	size_t index = std::upper_bound(times.begin(), times.end(), time) - times.begin();
	index = Min(index, times.size() - 1);
	return interpolatedTransforms[index].Sample(time);
}

Matrix4x4 MotionSystem::SampleInverse(const float time) const {
	size_t index = std::upper_bound(times.begin(), times.end(), time) - times.begin();
	index = Min(index, times.size() - 1);
	return interpolatedInverseTransforms[index].Sample(time);
}

BBox MotionSystem::Bound(BBox ibox, const bool storingGlobal2Local) const {
	BBox result;
	for (size_t i = 0; i < interpolatedTransforms.size(); ++i)
		result = Union(result, interpolatedTransforms[i].Bound(ibox, storingGlobal2Local));
	return result;
}

}   // namespace luxrays
