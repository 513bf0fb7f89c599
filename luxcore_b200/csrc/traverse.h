// traverse.h -- per-ray closest-hit traversal over the wide layout (layout.h).
//
// One function body serves the CUDA kernels (trace_kernels.cuh) and, compiled for the host with
// the same no-FMA floating-point model, the CPU-only unit tests of the layout/traversal logic
// (tests/cpp/wide_emulation.cpp).  The host build is test scaffolding: the shipped library only
// contains the device instantiation.
//
// Arithmetic contract (parity with the reference's native CPU Intersect):
//   * triangle test: operation-for-operation Triangle::Intersect
//     (include/luxrays/core/geometry/triangle.h:55-89), products and sums kept separate (no FMA);
//   * box test: BBox::IntersectP (src/luxrays/core/geometry/bbox.cpp:147-165) with the same
//     (p - o) * (1/d) formulation, the same swap-on-`>` and the same NaN-ignoring selects; 1/d is
//     hoisted (it is the same IEEE value on every call);
//   * hit acceptance: strictly closer wins; on an exactly equal t the leaf that comes first in
//     the reference's depth-first array wins (bvhaccel.cpp:233, mbvhaccel.cpp:305);
//   * ray -> instance space: InvTransform * Ray (transform.h:144-155,190-197,259-262), divide by
//     w only when w != 1; motion: MotionSystem::Sample (motionsystem.cpp:340-346,89-159).
// The traversal ORDER differs from the reference (near-to-far, stack-based); order only changes
// which boxes are culled against an already-found closer hit, never the result.
#ifndef LRB_TRAVERSE_H
#define LRB_TRAVERSE_H

#include <math.h>
#include <stdint.h>

#include "luxrays_b200.h"
#include "layout.h"

#if defined(__CUDACC__)
#define LRB_HD __host__ __device__ __forceinline__
#else
#define LRB_HD inline
#endif

#if defined(__CUDA_ARCH__)
// explicit round-to-nearest intrinsics are never contracted into FMAs, whatever -fmad says
#define LRB_MUL(a, b) __fmul_rn((a), (b))
#define LRB_ADD(a, b) __fadd_rn((a), (b))
#define LRB_SUB(a, b) __fsub_rn((a), (b))
#define LRB_RCP(a) __frcp_rn(a)            /* IEEE 1/x, round to nearest */
#define LRB_DIV(a, b) __fdiv_rn((a), (b))
#define LRB_LDG4(p) __ldg(reinterpret_cast<const float4 *>(p))
#define LRB_LDGU4(p) __ldg(reinterpret_cast<const uint4 *>(p))
#define LRB_INF __int_as_float(0x7f800000)
#define LRB_FMAX(a, b) fmaxf((a), (b))
#define LRB_FMIN(a, b) fminf((a), (b))
#define LRB_F2U(x) __float_as_uint(x)
#define LRB_U2F(x) __uint_as_float(x)
#define LRB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
/* byte k of w placed in mantissa bits 8-15 of 1.0f: the float 1 + q * 2^-15 (one PRMT) */
#define LRB_QBYTE(w, k, one) lrb::QByte<(k)>((w), (one))
/* Slerp's sinf/acosf (quaternion.cpp:150-156): the reference gets glibc's results, which are the
 * correctly rounded float in all but vanishingly rare cases.  CUDA's sinf/acosf are 1-2 ulp
 * functions, enough to move b1/b2 of distant triangles past the 1e-5 tolerance, so the device
 * evaluates them in double precision and rounds once (motion leaves only; not on the hot path). */
#define LRB_SINF(x) ((float)sin((double)(x)))
#define LRB_ACOSF(x) ((float)acos((double)(x)))
#else
#define LRB_INF __builtin_huge_valf()
/* fmaxf/fminf semantics of the device (the non-NaN operand wins) */
#define LRB_FMAX(a, b) ((a) != (a) ? (b) : ((b) != (b) ? (a) : ((a) > (b) ? (a) : (b))))
#define LRB_FMIN(a, b) ((a) != (a) ? (b) : ((b) != (b) ? (a) : ((a) < (b) ? (a) : (b))))
#define LRB_F2U(x) lrb::HostF2U(x)
#define LRB_U2F(x) lrb::HostU2F(x)
#define LRB_FMA(a, b, c) fmaf((a), (b), (c))
#define LRB_QBYTE(w, k, one) lrb::HostU2F((one) | ((((w) >> (8 * (k))) & 0xffu) << 8))
#define LRB_SINF(x) sinf(x)
#define LRB_ACOSF(x) acosf(x)
#define LRB_MUL(a, b) ((a) * (b))
#define LRB_ADD(a, b) ((a) + (b))
#define LRB_SUB(a, b) ((a) - (b))
#define LRB_RCP(a) (1.f / (a))
#define LRB_DIV(a, b) ((a) / (b))
#endif

namespace lrb {

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
#endif

// 32 bytes fetched by ONE load instruction: Blackwell's 256-bit LDG (ld.global.nc.v8.f32,
// SASS LDG.E.ENL2.256.CONSTANT) halves the L1 data-pipe wavefronts of a divergent node fetch.
struct F8 { float v[8]; };

#if defined(__CUDA_ARCH__)
// The selector is the immediate and 1.0f sits in a register (a PRMT takes one immediate).
template <int K> __device__ __forceinline__ float QByte(const uint32_t w, const uint32_t one) {
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(one), "n"(0x7604 | (K << 4)));
	return __uint_as_float(r);
}
__device__ __forceinline__ F8 Ld256(const void *p) {
	F8 r;
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
		: "l"(p));
	return r;
}
// L2 prefetch of the record a child reference points at (scenes that do not fit L2; see NodeStep).
__device__ __forceinline__ void PrefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void PrefetchL1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
#else
static inline void PrefetchL2(const void *) { }
static inline void PrefetchL1(const void *) { }
static inline F8 Ld256(const void *p) { F8 r; __builtin_memcpy(&r, p, 32); return r; }
static inline uint32_t HostF2U(float x) { uint32_t u; __builtin_memcpy(&u, &x, 4); return u; }
static inline float HostU2F(uint32_t u) { float x; __builtin_memcpy(&x, &u, 4); return x; }
#endif

#if !defined(__CUDA_ARCH__)
static inline float4 HostLd4(const void *p) { float4 r; __builtin_memcpy(&r, p, 16); return r; }
static inline uint4 HostLdU4(const void *p) { uint4 r; __builtin_memcpy(&r, p, 16); return r; }
#undef LRB_LDG4
#undef LRB_LDGU4
#define LRB_LDG4(p) HostLd4(p)
#define LRB_LDGU4(p) HostLdU4(p)
#endif

struct TraceStats {
	unsigned long long rays, wideNodes, triangles, instances, motionSamples, maxStack;
};

// Per-ray traversal state.  Everything lives in registers on the device.
struct RayState {
	float ox, oy, oz;       // ray in the CURRENT space (world, or instance-local)
	float dx, dy, dz;
	float ix, iy, iz;       // 1/d of the current-space ray
	float mint, maxt;       // maxt == best t so far (ray.maxt = rayHit->t in the reference)
	float time;
	float b1, b2;
	uint32_t hitRef;                // TriRecord index of the current best hit (== its reference array order), or kNullIndex
	uint32_t hitMeshOffset;         // two-level: mesh offset of the instance the best hit lies in
	uint32_t bestInst;              // two-level: reference array order of that instance
	uint32_t curInstOrder, curMeshOffset;
	// What to process next: a wide-node index, kTagTri | triangle, kTagInstance | instance,
	// kStackSentinel, or kNullIndex when the stack must be popped (Resolve).
	uint32_t cur;
	bool inInstance;
};

LRB_HD float Dot3(float ax, float ay, float az, float bx, float by, float bz) {
	// vector.h:151-153: (x*x + y*y) + z*z
	return LRB_ADD(LRB_ADD(LRB_MUL(ax, bx), LRB_MUL(ay, by)), LRB_MUL(az, bz));
}

// Triangle::Intersect, triangle.h:55-89, evaluated without branches: every quantity is computed
// and the reference's chain of early-outs becomes one predicate.  `!(x < 0)` keeps the reference's
// behaviour for NaN (a NaN barycentric does not reject), and divisor == 0 rejects explicitly.
LRB_HD bool TriangleTest(const RayState &r, const float p0x, const float p0y, const float p0z,
		const float p1x, const float p1y, const float p1z, const float p2x, const float p2y, const float p2z,
		float *tOut, float *b1Out, float *b2Out) {
	const float e1x = LRB_SUB(p1x, p0x), e1y = LRB_SUB(p1y, p0y), e1z = LRB_SUB(p1z, p0z);
	const float e2x = LRB_SUB(p2x, p0x), e2y = LRB_SUB(p2y, p0y), e2z = LRB_SUB(p2z, p0z);
	// s1 = Cross(d, e2)   (vector.h:159-163)
	const float s1x = LRB_SUB(LRB_MUL(r.dy, e2z), LRB_MUL(r.dz, e2y));
	const float s1y = LRB_SUB(LRB_MUL(r.dz, e2x), LRB_MUL(r.dx, e2z));
	const float s1z = LRB_SUB(LRB_MUL(r.dx, e2y), LRB_MUL(r.dy, e2x));
	const float divisor = Dot3(s1x, s1y, s1z, e1x, e1y, e1z);
	const float invDivisor = LRB_RCP(divisor);

	const float ddx = LRB_SUB(r.ox, p0x), ddy = LRB_SUB(r.oy, p0y), ddz = LRB_SUB(r.oz, p0z);
	const float b1 = LRB_MUL(Dot3(ddx, ddy, ddz, s1x, s1y, s1z), invDivisor);
	// s2 = Cross(dd, e1)
	const float s2x = LRB_SUB(LRB_MUL(ddy, e1z), LRB_MUL(ddz, e1y));
	const float s2y = LRB_SUB(LRB_MUL(ddz, e1x), LRB_MUL(ddx, e1z));
	const float s2z = LRB_SUB(LRB_MUL(ddx, e1y), LRB_MUL(ddy, e1x));
	const float b2 = LRB_MUL(Dot3(r.dx, r.dy, r.dz, s2x, s2y, s2z), invDivisor);
	const float b0 = LRB_SUB(LRB_SUB(1.f, b1), b2);
	const float t = LRB_MUL(Dot3(e2x, e2y, e2z, s2x, s2y, s2z), invDivisor);

	const bool hit = (divisor != 0.f) & !(b1 < 0.f) & !(b2 < 0.f) & !(b0 < 0.f) & !(t < r.mint) & !(t > r.maxt);
	*tOut = t; *b1Out = b1; *b2Out = b2;
	return hit;
}

// BBox::IntersectP (bbox.cpp:147-165) for one child box, returning its entry distance, or +inf
// when the ray misses it / the slot is unused.
// The reference computes a = (lo - o) * inv, b = (hi - o) * inv and swaps when a > b.  lo <= hi, so
// for finite values the swap happens exactly when inv < 0: picking the near / far plane by the sign
// of inv yields the same two numbers with a select per plane instead of a compare + two selects.
// When a product is NaN (0 * inf: origin on a slab plane, direction parallel to it) the reference
// does not swap and its `>` / `<` updates ignore the NaN; here fmaxf / fminf ignore it as well, and
// in the sub-cases where the two differ this form only ever PASSES a box the reference rejects
// (never the opposite), which cannot change a result.
// (Used for the root box, which travels as exact floats in the kernel parameters.  Unused slots of a
// node hold an inverted box; a degenerate ray that "passes" one gets a kNullIndex reference, which
// Resolve drops.)
LRB_HD float ChildEntry(const RayState &s, const bool nx, const bool ny, const bool nz,
		const float lox, const float loy, const float loz, const float hix, const float hiy, const float hiz) {
	const float tnx = LRB_MUL(LRB_SUB(nx ? hix : lox, s.ox), s.ix);
	const float tny = LRB_MUL(LRB_SUB(ny ? hiy : loy, s.oy), s.iy);
	const float tnz = LRB_MUL(LRB_SUB(nz ? hiz : loz, s.oz), s.iz);
	const float tfx = LRB_MUL(LRB_SUB(nx ? lox : hix, s.ox), s.ix);
	const float tfy = LRB_MUL(LRB_SUB(ny ? loy : hiy, s.oy), s.iy);
	const float tfz = LRB_MUL(LRB_SUB(nz ? loz : hiz, s.oz), s.iz);
	const float t0 = LRB_FMAX(LRB_FMAX(LRB_FMAX(tnx, tny), tnz), s.mint);
	const float t1 = LRB_FMIN(LRB_FMIN(LRB_FMIN(tfx, tfy), tfz), s.maxt);
	return !(t0 > t1) ? t0 : LRB_INF;
}

LRB_HD void SetRay(RayState &s, float ox, float oy, float oz, float dx, float dy, float dz) {
	s.ox = ox; s.oy = oy; s.oz = oz;
	s.dx = dx; s.dy = dy; s.dz = dz;
	s.ix = LRB_RCP(dx); s.iy = LRB_RCP(dy); s.iz = LRB_RCP(dz);
}

// InvTransform * Ray with m = mInv (transform.h:144-155,190-197,259-262)
LRB_HD void TransformRay(RayState &s, const float *m, float ox, float oy, float oz, float dx, float dy, float dz) {
	const float4 r0 = LRB_LDG4(m), r1 = LRB_LDG4(m + 4), r2 = LRB_LDG4(m + 8), r3 = LRB_LDG4(m + 12);
	float px = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(r0.x, ox), LRB_MUL(r0.y, oy)), LRB_MUL(r0.z, oz)), r0.w);
	float py = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(r1.x, ox), LRB_MUL(r1.y, oy)), LRB_MUL(r1.z, oz)), r1.w);
	float pz = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(r2.x, ox), LRB_MUL(r2.y, oy)), LRB_MUL(r2.z, oz)), r2.w);
	const float w = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(r3.x, ox), LRB_MUL(r3.y, oy)), LRB_MUL(r3.z, oz)), r3.w);
	if (w != 1.f) {
		const float inv = LRB_RCP(w);
		px = LRB_MUL(inv, px); py = LRB_MUL(inv, py); pz = LRB_MUL(inv, pz);
	}
	const float vx = LRB_ADD(LRB_ADD(LRB_MUL(r0.x, dx), LRB_MUL(r0.y, dy)), LRB_MUL(r0.z, dz));
	const float vy = LRB_ADD(LRB_ADD(LRB_MUL(r1.x, dx), LRB_MUL(r1.y, dy)), LRB_MUL(r1.z, dz));
	const float vz = LRB_ADD(LRB_ADD(LRB_MUL(r2.x, dx), LRB_MUL(r2.y, dy)), LRB_MUL(r2.z, dz));
	SetRay(s, px, py, pz, vx, vy, vz);
}

LRB_HD float LerpF(float t, float v1, float v2) { return LRB_ADD(v1, LRB_MUL(t, LRB_SUB(v2, v1))); }   // utils.h:122-125

// MotionSystem::Sample -> 4x4 row-major matrix in m[16]
LRB_HD void MotionSample(const SceneView &sc, uint32_t motionIndex, float time, float *m) {
	const uint32_t first = sc.motionFirst[motionIndex], last = sc.motionLast[motionIndex];
	// upper_bound(times, time) clamped to times.size() - 1 (motionsystem.cpp:340-344); times[i] is
	// the endTime of interpolated transform first + i, and the trailing static entry is never picked.
	uint32_t index = last > first ? last - 1 : first;
	for (uint32_t i = first; i < last; ++i) {
		if (time < sc.interps[i].endTime) { index = i; break; }
	}
	const DevInterp &it = sc.interps[index];
	const uint32_t f = it.flags;
	const float *src = it.startM;
	bool direct = true;
	if (f & kItActive) {
		if (time <= it.startTime) src = it.startM;
		else if (time >= it.endTime) src = it.endM;
		else direct = false;
	}
	if (direct) {
		for (int i = 0; i < 16; ++i) m[i] = src[i];
		return;
	}
	const float w = LRB_SUB(it.endTime, it.startTime);
	const float d = LRB_SUB(time, it.startTime);
	const float le = LRB_DIV(d, w);

	if ((f & kItTranslation) && !(f & (kItScale | kItRotation))) {
		for (int i = 0; i < 16; ++i) m[i] = it.startM[i];
		if (f & kItTX) m[3] = LerpF(le, it.sT[0], it.eT[0]);
		if (f & kItTY) m[7] = LerpF(le, it.sT[1], it.eT[1]);
		if (f & kItTZ) m[11] = LerpF(le, it.sT[2], it.eT[2]);
		return;
	}

	if (f & kItRotation) {
		// Slerp (quaternion.cpp:144-164)
		const float q1w = it.sQ[0], q1x = it.sQ[1], q1y = it.sQ[2], q1z = it.sQ[3];
		const float q2w = it.eQ[0], q2x = it.eQ[1], q2y = it.eQ[2], q2z = it.eQ[3];
		float cosPhi = LRB_ADD(LRB_MUL(q1w, q2w), Dot3(q1x, q1y, q1z, q2x, q2y, q2z));
		const float sign = (cosPhi > 0.f) ? 1.f : -1.f;
		cosPhi = LRB_MUL(cosPhi, sign);
		float f1, f2;
		if (LRB_SUB(1.f, cosPhi) > 1e-6f) {
			const float phi = LRB_ACOSF(cosPhi);
			const float sinPhi = LRB_SINF(phi);
			f1 = LRB_DIV(LRB_SINF(LRB_MUL(LRB_SUB(1.f, le), phi)), sinPhi);
			f2 = LRB_DIV(LRB_SINF(LRB_MUL(le, phi)), sinPhi);
		} else {
			f1 = LRB_SUB(1.f, le);
			f2 = le;
		}
		const float g2 = LRB_MUL(sign, f2);
		const float qw = LRB_ADD(LRB_MUL(q1w, f1), LRB_MUL(q2w, g2));
		const float qx = LRB_ADD(LRB_MUL(q1x, f1), LRB_MUL(q2x, g2));
		const float qy = LRB_ADD(LRB_MUL(q1y, f1), LRB_MUL(q2y, g2));
		const float qz = LRB_ADD(LRB_MUL(q1z, f1), LRB_MUL(q2z, g2));
		// ToMatrix (quaternion.cpp:117-142)
		const float xx = LRB_MUL(qx, qx), yy = LRB_MUL(qy, qy), zz = LRB_MUL(qz, qz);
		const float xy = LRB_MUL(qx, qy), xz = LRB_MUL(qx, qz), yz = LRB_MUL(qy, qz);
		const float xw = LRB_MUL(qx, qw), yw = LRB_MUL(qy, qw), zw = LRB_MUL(qz, qw);
		m[0] = LRB_SUB(1.f, LRB_MUL(2.f, LRB_ADD(yy, zz)));
		m[4] = LRB_MUL(2.f, LRB_SUB(xy, zw));
		m[8] = LRB_MUL(2.f, LRB_ADD(xz, yw));
		m[1] = LRB_MUL(2.f, LRB_ADD(xy, zw));
		m[5] = LRB_SUB(1.f, LRB_MUL(2.f, LRB_ADD(xx, zz)));
		m[9] = LRB_MUL(2.f, LRB_SUB(yz, xw));
		m[2] = LRB_MUL(2.f, LRB_SUB(xz, yw));
		m[6] = LRB_MUL(2.f, LRB_ADD(yz, xw));
		m[10] = LRB_SUB(1.f, LRB_MUL(2.f, LRB_ADD(xx, yy)));
		m[3] = m[7] = m[11] = 0.f;
		m[12] = m[13] = m[14] = 0.f;
		m[15] = 1.f;
	} else {
		for (int i = 0; i < 16; ++i) m[i] = it.R[i];
	}

	float Sx = it.sS[0], Sy = it.sS[1], Sz = it.sS[2];
	if (f & kItScale) {
		Sx = LerpF(le, it.sS[0], it.eS[0]);
		Sy = LerpF(le, it.sS[1], it.eS[1]);
		Sz = LerpF(le, it.sS[2], it.eS[2]);
	}
	for (int j = 0; j < 3; ++j) {
		m[j] = LRB_MUL(Sx, m[j]);
		m[4 + j] = LRB_MUL(Sy, m[4 + j]);
		m[8 + j] = LRB_MUL(Sz, m[8 + j]);
	}
	m[3] = (f & kItTX) ? LerpF(le, it.sT[0], it.eT[0]) : it.sT[0];
	m[7] = (f & kItTY) ? LerpF(le, it.sT[1], it.eT[1]) : it.sT[1];
	m[11] = (f & kItTZ) ? LerpF(le, it.sT[2], it.eT[2]) : it.sT[2];
}

// Matrix4x4 * Ray with a matrix held in registers/local memory
LRB_HD void TransformRayLocal(RayState &s, const float *m, float ox, float oy, float oz, float dx, float dy, float dz) {
	float px = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(m[0], ox), LRB_MUL(m[1], oy)), LRB_MUL(m[2], oz)), m[3]);
	float py = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(m[4], ox), LRB_MUL(m[5], oy)), LRB_MUL(m[6], oz)), m[7]);
	float pz = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(m[8], ox), LRB_MUL(m[9], oy)), LRB_MUL(m[10], oz)), m[11]);
	const float w = LRB_ADD(LRB_ADD(LRB_ADD(LRB_MUL(m[12], ox), LRB_MUL(m[13], oy)), LRB_MUL(m[14], oz)), m[15]);
	if (w != 1.f) {
		const float inv = LRB_RCP(w);
		px = LRB_MUL(inv, px); py = LRB_MUL(inv, py); pz = LRB_MUL(inv, pz);
	}
	const float vx = LRB_ADD(LRB_ADD(LRB_MUL(m[0], dx), LRB_MUL(m[1], dy)), LRB_MUL(m[2], dz));
	const float vy = LRB_ADD(LRB_ADD(LRB_MUL(m[4], dx), LRB_MUL(m[5], dy)), LRB_MUL(m[6], dz));
	const float vz = LRB_ADD(LRB_ADD(LRB_MUL(m[8], dx), LRB_MUL(m[9], dy)), LRB_MUL(m[10], dz));
	SetRay(s, px, py, pz, vx, vy, vz);
}

// Motion-blurred instance: the ray in instance space at the ray's time.  On the device this is ONE out-of-line
// function: MotionSample carries double-precision sin / acos (see LRB_SINF) and a 16-float matrix, several hundred
// instructions that would otherwise be inlined into the hot loop of every two-level kernel (instruction-cache
// misses showed up as `no_instruction` stalls in the round-2 profile of lightinstances, which has no motion at all).
struct Ray6 { float ox, oy, oz, dx, dy, dz; };
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
Ray6 MotionRay(const uint32_t *motionFirst, const uint32_t *motionLast, const DevInterp *interps, const uint32_t motionIndex,
		const float time, const float ox, const float oy, const float oz, const float dx, const float dy, const float dz) {
	SceneView sc = SceneView();
	sc.motionFirst = motionFirst;
	sc.motionLast = motionLast;
	sc.interps = interps;
	float m[16];
	MotionSample(sc, motionIndex, time, m);
	RayState t;
	TransformRayLocal(t, m, ox, oy, oz, dx, dy, dz);
	Ray6 r = { t.ox, t.oy, t.oz, t.dx, t.dy, t.dz };
	return r;
}

// Initialise the state from a wire ray.  Returns false when there is nothing to traverse.
LRB_HD bool InitRay(const SceneView &sc, const lrb_ray &ray, RayState &s) {
	SetRay(s, ray.o[0], ray.o[1], ray.o[2], ray.d[0], ray.d[1], ray.d[2]);
	// A NaN mint never rejects anything in the reference (`t < mint` and BBox::IntersectP's `tNear > t0` are
	// both false): for the triangle test that is mint = -inf.  Kept as NaN it would poison the entry distance
	// of a slot whose three slabs are NaN as well (the whole-grid slot of a lone instance), and a NaN key reads
	// as "missed" (found by tools/fuzz_parity.py).
	s.mint = (ray.mint != ray.mint) ? -LRB_INF : ray.mint;
	s.maxt = ray.maxt;      // rayHit->t = ray->maxt
	s.time = ray.time;
	s.b1 = 0.f; s.b2 = 0.f;
	s.hitRef = kNullIndex;
	s.hitMeshOffset = 0;
	s.bestInst = 0;
	s.curInstOrder = 0; s.curMeshOffset = 0;
	s.inInstance = false;
	if (!sc.nWide) {
		s.cur = kNullIndex;
		return false;
	}
	if (sc.rootHasBox) {
		// root box test (bvhaccel.cpp:245-255 at currentNode == 0) straight from the parameters
		const float d = ChildEntry(s, s.ix < 0.f, s.iy < 0.f, s.iz < 0.f,
				sc.rootBox[0], sc.rootBox[1], sc.rootBox[2], sc.rootBox[3], sc.rootBox[4], sc.rootBox[5]);
		s.cur = (d < LRB_INF) ? sc.rootChild : kNullIndex;     // kNullIndex + empty stack => finished at the first Resolve
	} else
		s.cur = sc.rootWide;
	return true;
}

// True when s.cur is neither a wide node nor a triangle: the ray needs Resolve (and, for an instance
// reference, EnterInstance) before its next step.
template <bool TWO_LEVEL>
LRB_HD bool NeedsResolve(const uint32_t cur) {
	return TWO_LEVEL ? (cur >= kTagInstance) : (cur == kNullIndex);
}

// True for kTagInstance | index (not the sentinel, not kNullIndex): the ray is about to enter a leaf tree.
LRB_HD bool IsInstanceRef(const uint32_t cur) { return (cur >> 30) == 2u; }

// What a live ray holds after Resolve -- the kind of work it waits for.  One definition for the CUDA
// kernels and for the host model of their scheduling (tests/cpp/wide_emulation.cpp).
enum LaneWork { kWorkNone = 0, kWorkNode = 1, kWorkTri = 2, kWorkInstance = 3 };
template <bool TWO_LEVEL>
LRB_HD LaneWork WorkOf(const uint32_t cur) {
	if (TWO_LEVEL) {
		if (cur == kNullIndex || cur == kStackSentinel) return kWorkNone;     // (an empty leaf tree leaves kNullIndex behind)
		if (IsInstanceRef(cur)) return kWorkInstance;
	}
	return (cur & kTagTri) ? kWorkTri : kWorkNode;
}
// Phase votes of a warp.  Triangle phase when nTri * triBias >= nNode * 4 (4 = plain majority);
// instances are entered when the lanes waiting for it outweigh both other kinds (instBias 0: at once).
LRB_HD bool VoteTrianglePhase(const int nTri, const int nNode, const uint32_t triBias) { return nTri * (int)triBias >= nNode * 4; }
LRB_HD bool VoteEnterInstances(const int nInst, const int nNode, const int nTri, const uint32_t instBias, const uint32_t triBias) {
	const int a = nNode * 4, b = nTri * (int)triBias;
	return nInst > 0 && (instBias == 0 || nInst * (int)instBias >= (a > b ? a : b));
}

// Enters the leaf tree s.cur refers to (mbvhaccel.cpp:312-333): ray into instance space, sentinel on
// the stack, s.cur = root of the leaf tree -- or kNullIndex for an empty leaf tree (the next Resolve
// pops on).  Kept out of Resolve's pop loop: that loop runs a different number of trips on every lane,
// and this is its one expensive step (64-B matrix fetch, ~40 flops, three IEEE reciprocals); here all
// lanes of a warp that enter an instance in the same iteration do it in the same instructions.
template <bool STATS, class STACK>
LRB_HD void EnterInstance(const SceneView &sc, RayState &s, STACK &stk, TraceStats *stats) {
	const char *ip = reinterpret_cast<const char *>(&sc.insts[s.cur & kRefIndexMask]);
	const uint4 ir = LRB_LDGU4(ip);
	const uint4 ir2 = LRB_LDGU4(ip + 16);
	if (STATS) stats->instances++;
	if (ir.x == kNullIndex) {
		s.cur = kNullIndex;     // empty leaf tree
		return;
	}
	// s holds the world ray here (instances do not nest): keep it, with its 1/d, for the way back (Resolve), which
	// would otherwise reload it from the ray buffer and recompute three IEEE reciprocals inside its divergent pop loop
	stk.stashRay(s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, s.ix, s.iy, s.iz);
	if (ir.y != kNullIndex) {
		TransformRay(s, sc.minv + 16 * (size_t)ir.y, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz);
	} else if (ir.z != kNullIndex) {
		const Ray6 r = MotionRay(sc.motionFirst, sc.motionLast, sc.interps, ir.z, s.time, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz);
		if (STATS) stats->motionSamples++;
		SetRay(s, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
	}
	s.curMeshOffset = ir.w;
	s.curInstOrder = ir2.x;
	s.inInstance = true;
	stk.push(kStackSentinel, -LRB_INF);
	s.cur = ir.x;
}

// A popped reference that is neither a wide node nor a triangle (cur >= kTagInstance; not culled).  Returns true when
// the pop loop ends with it: an instance reference (the caller runs EnterInstance), or the bottom of the stack
// (*bottom set: the ray is finished).  The sentinel takes the ray back to world space (mbvhaccel.cpp:271-283, from
// the stash); it and the kNullIndex reference of an empty slot (only NaN / inf rays push those) are popped over.
template <bool TWO_LEVEL, class STACK>
LRB_HD bool RarePopped(RayState &s, STACK &stk, const uint32_t cur, bool *bottom) {
	if (cur == kStackBottom) {
		*bottom = true;
		return true;
	}
	if (TWO_LEVEL) {
		if (cur == kStackSentinel) {
			stk.loadRay(s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, s.ix, s.iy, s.iz);
			s.inInstance = false;
			return false;
		}
		return cur != kNullIndex;
	}
	return false;
}

// Turns s.cur into a wide-node, triangle or instance reference: pops the stack while there is nothing
// to do or the popped entry lies behind the best hit, and leaves leaf trees (sentinel).  An instance
// reference is returned as it is (EnterInstance follows).  Returns false when the ray is finished.
// An entry's distance was recorded at push time; a closer hit found since then culls the entry (same effect as
// running the box test now: t0 > min(maxt, tFar)).  Sentinel and bottom carry -inf and are never culled.
// STACK provides push(uint32_t ref, float t0) / pop(uint32_t&, float&) / depth() / room(n), a nine-float side slot
// stashRay(o, d, 1/d) / loadRay(...) (the world ray while the traversal is inside an instance), and
// slow() / popFast() / keepBottom(): while slow() the top entry lives in the global spill column and pop() must be
// used; afterwards (pops never make a stack slow again) popFast() takes entries without looking at the spill and
// returns (kStackBottom, -inf) from an empty stack -- the shared-memory columns keep that entry permanently
// under the first real one, so the divergent loop below needs no emptiness test (it runs a different number of
// trips on every lane and every instruction in it counts); keepBottom() undoes the pop of the bottom entry.
template <bool TWO_LEVEL, bool STATS, class STACK>
LRB_HD bool Resolve(const SceneView &sc, RayState &s, STACK &stk, TraceStats *stats) {
	uint32_t cur = s.cur;
	bool bottom = false;
	// On entry s.cur is kNullIndex (pop), or -- two-level -- an instance reference, which is returned as it is.
	// Single exit: lanes that finish and lanes that found work leave the loops together.
	if (!TWO_LEVEL || !IsInstanceRef(cur)) {
		bool found = false;
		while (stk.slow()) {
			float t0;
			stk.pop(cur, t0);
			if (t0 > s.maxt)
				continue;
			if (cur < kTagInstance || RarePopped<TWO_LEVEL>(s, stk, cur, &bottom)) {
				found = true;
				break;
			}
		}
		if (!found) {
			for (;;) {
				float t0;
				stk.popFast(cur, t0);
				if (t0 > s.maxt)
					continue;
				if (cur < kTagInstance)
					break;                  // a wide node or a triangle: the common case
				if (RarePopped<TWO_LEVEL>(s, stk, cur, &bottom))
					break;
			}
		}
		if (bottom) {
			stk.keepBottom();
			cur = kNullIndex;
		}
	}
	s.cur = cur;
	return !bottom;
}

// Speculative pop at the end of a phase: a lane that is left with nothing to do (triangle tested; node without a
// child on the ray) takes the top entry of its stack while the lanes of the phase are still converged.  An entry
// behind the best hit is dropped and leaves s.cur empty; the sentinel, the bottom of the stack, kNullIndex
// entries and spilled entries are left where they are: Resolve deals with all of those at the next iteration.  Purely a scheduling
// device -- the entries come off the stack in the same order with the same culling rule as in Resolve.
// STACK::peekIf(want, ref&, t0&) -> bool reads the top entry if there is one at hand; dropIf(bool) removes it.
#ifndef LRB_POPSPEC
#define LRB_POPSPEC 0       /* attempts per phase end; 0 = none (every pop happens in Resolve).  Measured on B200, kitchen
                             * bounce-2 (gpurun_out/r02_measure_kitchen_c5_*): 0 -> 5.03 ms, 1 -> 5.08 ms, 2 -> 5.22 ms per 16 Mi rays:
                             * what the attempts take out of Resolve's loop (65 -> 35 warp instructions per ray) they add to
                             * the phases, where every lane of the phase pays for them. */
#endif
template <bool TWO_LEVEL, class STACK>
LRB_HD void PopSpec(RayState &s, STACK &stk) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int attempt = 0; attempt < LRB_POPSPEC; ++attempt) {
		uint32_t c = kNullIndex;
		float t0 = 0.f;
		const bool can = stk.peekIf(s.cur == kNullIndex, c, t0);
		const bool take = can && c < kStackBottom;      // not the bottom, the sentinel or an empty slot's kNullIndex
		stk.dropIf(take);
		if (take)
			s.cur = (t0 > s.maxt) ? kNullIndex : c;
	}
}

// Tests the triangle s.cur refers to.
//   accept: strictly closer, or exactly as close as the current hit but earlier in the reference's
//   depth-first array (so a hit at exactly t == ray.maxt is rejected while nothing was hit yet,
//   like the reference's `t < rayHit->t` against the initial rayHit->t = maxt).
template <bool TWO_LEVEL, bool STATS>
LRB_HD void TriStep(const SceneView &sc, RayState &s, TraceStats *stats) {
	const uint32_t triIndex = s.cur & kRefIndexMask;
	const char *tp = reinterpret_cast<const char *>(sc.tris + triIndex);
	s.cur = kNullIndex;
	const F8 a = Ld256(tp), b = Ld256(tp + 32);    // p0 p1 p2.xy | p2.z gateLo gateHi order
	if (STATS) stats->triangles++;
	float t, b1, b2;
	const bool hit = TriangleTest(s, a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7], b.v[0], &t, &b1, &b2);
	const uint32_t instOrder = TWO_LEVEL ? s.curInstOrder : 0u;
	const bool closer = t < s.maxt;
	// (records are stored in reference order: the record index is the tie-break order, layout.h)
	const bool tieWin = (t == s.maxt) & (s.hitRef != kNullIndex) &
			((instOrder < s.bestInst) | ((instOrder == s.bestInst) & (triIndex < s.hitRef)));
	// the reference's gate (layout.h): the exact box of the triangle's parent with the reference's own
	// arithmetic, against the ray as it is BEFORE this hit shortens it.  A genuine hit always passes (the
	// hit point is inside the triangle's grown box).  Evaluated for every tested triangle, next to the
	// triangle test: no second fetch and no divergent side path for the accepted hits.
	const float ge = ChildEntry(s, s.ix < 0.f, s.iy < 0.f, s.iz < 0.f, b.v[1], b.v[2], b.v[3], b.v[4], b.v[5], b.v[6]);
	if (hit & (closer | tieWin) & (ge < LRB_INF)) {
		s.maxt = t;
		s.b1 = b1; s.b2 = b2;
		s.hitRef = triIndex;
		if (TWO_LEVEL) {
			s.hitMeshOffset = s.curMeshOffset;
			s.bestInst = instOrder;
		}
	}
}

// Entry distance of one slot box of a quantized node, or +inf when the ray misses it.
//   plane position = org + q * step  =>  t = (org + q * step - o) * inv = fma(1 + q * 2^-15, A, B)
//   with A = step * 2^15 * inv and B = (org - o) * inv - A  (per node and axis, see NodeStep);
//   `1 + q * 2^-15` is the float whose mantissa bits 8-15 hold the byte q (one PRMT).
// Rounding: B carries an error below ulp(A) = 2^-9 step, covered by the 1/64-step margin the
// planes were rounded outward with (relayout.cpp QuantizeNode); the remaining error is the same
// few ulp of |plane - o| * |inv| that the reference's own (p - o) * inv test has, far inside the
// >= 128 ulp / 1e-5 by which every build box is grown around its triangle (bvhaccel.cpp:116-122).
// Degenerate products (0 * inf, inf - inf) give NaN, which fmaxf / fminf ignore: the box passes.
template <int k>
LRB_HD float SlotEntry(const RayState &s, const uint32_t one, const uint32_t nqx, const uint32_t nqy, const uint32_t nqz,
		const uint32_t fqx, const uint32_t fqy, const uint32_t fqz,
		const float ax, const float ay, const float az, const float bx, const float by, const float bz) {
	const float tnx = LRB_FMA(LRB_QBYTE(nqx, k, one), ax, bx);
	const float tny = LRB_FMA(LRB_QBYTE(nqy, k, one), ay, by);
	const float tnz = LRB_FMA(LRB_QBYTE(nqz, k, one), az, bz);
	const float tfx = LRB_FMA(LRB_QBYTE(fqx, k, one), ax, bx);
	const float tfy = LRB_FMA(LRB_QBYTE(fqy, k, one), ay, by);
	const float tfz = LRB_FMA(LRB_QBYTE(fqz, k, one), az, bz);
	const float t0 = LRB_FMAX(LRB_FMAX(LRB_FMAX(tnx, tny), tnz), s.mint);
	const float t1 = LRB_FMIN(LRB_FMIN(LRB_FMIN(tfx, tfy), tfz), s.maxt);
	return !(t0 > t1) ? t0 : LRB_INF;
}

// Visits the wide node s.cur refers to: box-tests its four child slots (no branches), orders the
// children near-to-far, continues with the nearest and pushes the others with their entry
// distances.
//
// PREFETCH (scenes larger than L2, where every fetch of a deep node is a DRAM round trip the whole warp
// waits for): the records of the children that go on the stack are requested into L2 now; by the time
// one of them is popped its fetch hits L2.  Entries that are culled before they are popped cost
// bandwidth (plentiful: a latency-bound walk uses a small fraction of HBM), not time.
// pfMode (PREFETCH kernels; device option prefetch_mode): bit 0 = the pushed children into L2, bit 1 = the NEAREST child
// -- the record this lane fetches next, whatever happens, about a hundred issue slots from now -- into L2, bit 2 =
// the nearest child into L1.
template <bool TWO_LEVEL, bool STATS, bool PREFETCH = false, class STACK>
LRB_HD void NodeStep(const SceneView &sc, RayState &s, STACK &stk, TraceStats *stats, const uint32_t pfMode = 1u) {
	// ---- fetch the 64-byte node with two 256-bit loads ----
	const char *np = reinterpret_cast<const char *>(sc.nodes + s.cur);
	if (STATS) stats->wideNodes++;
	const F8 A = Ld256(np);          // org[3] exps child[4]
	const F8 B = Ld256(np + 32);     // qlo[3] qhi[3] next flags
	uint32_t c0 = LRB_F2U(A.v[4]), c1 = LRB_F2U(A.v[5]), c2 = LRB_F2U(A.v[6]), c3 = LRB_F2U(A.v[7]);
	const uint32_t exps = LRB_F2U(A.v[3]);
	const uint32_t next = LRB_F2U(B.v[6]);

	// per-axis decode constants.  A zero direction component has 1/d = +-inf, for which A = inf and
	// B = inf - inf = NaN: the slab would be ignored and an axis-parallel ray would walk every box in
	// its column (15 000 nodes per ray on the kitchen, half the tree on a 50 M-triangle soup).  With
	// the reciprocal clamped to +-2^80 the slab keeps deciding by position, like the reference's
	// (lo - o) * inf = +-inf: t = (plane - o) * 2^80 is astronomically negative / positive on either
	// side of the plane (>= 1e17 for planes one float spacing away) and 0 on it, and A = step * 2^15
	// * 2^80 stays finite for any grid step below 2^33.  Directions with |d| > 2^-80 are not affected.
	const float kBig = 1.2089258e24f;
	const float cix = LRB_FMIN(LRB_FMAX(s.ix, -kBig), kBig);
	const float ciy = LRB_FMIN(LRB_FMAX(s.iy, -kBig), kBig);
	const float ciz = LRB_FMIN(LRB_FMAX(s.iz, -kBig), kBig);
	const float ax = LRB_MUL(LRB_U2F((exps << 23) & 0x7f800000u), cix);
	const float ay = LRB_MUL(LRB_U2F((exps << 15) & 0x7f800000u), ciy);
	const float az = LRB_MUL(LRB_U2F((exps << 7) & 0x7f800000u), ciz);
	const float bx = LRB_SUB(LRB_MUL(LRB_SUB(A.v[0], s.ox), cix), ax);
	const float by = LRB_SUB(LRB_MUL(LRB_SUB(A.v[1], s.oy), ciy), ay);
	const float bz = LRB_SUB(LRB_MUL(LRB_SUB(A.v[2], s.oz), ciz), az);
	// near / far plane words by direction sign (see ChildEntry)
	const bool nx = s.ix < 0.f, ny = s.iy < 0.f, nz = s.iz < 0.f;
	const uint32_t qlx = LRB_F2U(B.v[0]), qly = LRB_F2U(B.v[1]), qlz = LRB_F2U(B.v[2]);
	const uint32_t qhx = LRB_F2U(B.v[3]), qhy = LRB_F2U(B.v[4]), qhz = LRB_F2U(B.v[5]);
	const uint32_t nqx = nx ? qhx : qlx, nqy = ny ? qhy : qly, nqz = nz ? qhz : qlz;
	const uint32_t fqx = nx ? qlx : qhx, fqy = ny ? qly : qhy, fqz = nz ? qlz : qhz;

	// ---- box tests of all four slots, near-to-far ordering ----
	// A box that passes with entry distance +inf (mint = maxt = +inf) cannot hold an acceptable hit:
	// the triangle test rejects t > maxt and a tie at +inf never wins; "not hit" is exact.
	const float kInf = LRB_INF;
	float d0 = SlotEntry<0>(s, sc.oneBits, nqx, nqy, nqz, fqx, fqy, fqz, ax, ay, az, bx, by, bz);
	float d1 = SlotEntry<1>(s, sc.oneBits, nqx, nqy, nqz, fqx, fqy, fqz, ax, ay, az, bx, by, bz);
	float d2 = SlotEntry<2>(s, sc.oneBits, nqx, nqy, nqz, fqx, fqy, fqz, ax, ay, az, bx, by, bz);
	float d3 = SlotEntry<3>(s, sc.oneBits, nqx, nqy, nqz, fqx, fqy, fqz, ax, ay, az, bx, by, bz);

	// sorting network on (distance, child), ascending, written with selects
#define LRB_CSWAP(da, ca, db, cb) { const bool sw_ = db < da; const float lo_ = sw_ ? db : da, hi_ = sw_ ? da : db; \
		const uint32_t cl_ = sw_ ? cb : ca, ch_ = sw_ ? ca : cb; da = lo_; db = hi_; ca = cl_; cb = ch_; }
	LRB_CSWAP(d0, c0, d1, c1)
	LRB_CSWAP(d2, c2, d3, c3)
	LRB_CSWAP(d0, c0, d2, c2)
	LRB_CSWAP(d1, c1, d3, c3)
	LRB_CSWAP(d1, c1, d2, c2)
#undef LRB_CSWAP

	if (PREFETCH && (pfMode & 6u)) {
		if (d0 < kInf && c0 < kTagInstance) {
			const char *p0 = ((c0 & kTagTri) ? reinterpret_cast<const char *>(sc.tris) : reinterpret_cast<const char *>(sc.nodes)) +
					((size_t)(c0 & kRefIndexMask) << 6);
			if (pfMode & 2u) PrefetchL2(p0);
			if (pfMode & 4u) PrefetchL1(p0);
		}
	}
	// push far-to-near, continue with the nearest.  A continuation node (reference nodes with more
	// than four children) is always visited.
	const bool h1 = d1 < kInf, h2 = d2 < kInf, h3 = d3 < kInf;
	if (stk.room(4)) {
		if (next != kNullIndex) stk.pushFast(next, -kInf);
		stk.pushFastIf(h3, c3, d3);
		stk.pushFastIf(h2, c2, d2);
		stk.pushFastIf(h1, c1, d1);
	} else {
		if (next != kNullIndex) stk.push(next, -kInf);
		if (h3) stk.push(c3, d3);
		if (h2) stk.push(c2, d2);
		if (h1) stk.push(c1, d1);
	}
	if (PREFETCH && (pfMode & 1u)) {
		// wide nodes and triangle records are both 64 bytes: one address computation serves either
		const char *nb = reinterpret_cast<const char *>(sc.nodes), *tb = reinterpret_cast<const char *>(sc.tris);
		if (h1 && c1 < kTagInstance) PrefetchL2(((c1 & kTagTri) ? tb : nb) + ((size_t)(c1 & kRefIndexMask) << 6));
		if (h2 && c2 < kTagInstance) PrefetchL2(((c2 & kTagTri) ? tb : nb) + ((size_t)(c2 & kRefIndexMask) << 6));
		if (h3 && c3 < kTagInstance) PrefetchL2(((c3 & kTagTri) ? tb : nb) + ((size_t)(c3 & kRefIndexMask) << 6));
	}
	s.cur = (d0 < kInf) ? c0 : kNullIndex;
	if (STATS) { const unsigned long long d = stk.depth(); if (d > stats->maxStack) stats->maxStack = d; }
}

// One unit of work for one ray: Resolve, then a node visit or a triangle test (static kernel, host
// emulation).  Returns false when the ray is finished.
template <bool TWO_LEVEL, bool STATS, class STACK>
LRB_HD bool Step(const SceneView &sc, RayState &s, STACK &stk, TraceStats *stats) {
	if (NeedsResolve<TWO_LEVEL>(s.cur)) {
		if (!Resolve<TWO_LEVEL, STATS>(sc, s, stk, stats))
			return false;
	}
	if (TWO_LEVEL && IsInstanceRef(s.cur)) {
		EnterInstance<STATS>(sc, s, stk, stats);
		if (s.cur == kNullIndex)
			return true;        // empty leaf tree: pop on at the next step
	}
	if (s.cur & kTagTri)
		TriStep<TWO_LEVEL, STATS>(sc, s, stats);
	else
		NodeStep<TWO_LEVEL, STATS>(sc, s, stk, stats);
	return true;
}

// Final RayHit.  Miss payload: t = ray.maxt, meshIndex = triangleIndex = NULL_INDEX (bvh.cl:219-223);
// b1/b2 are written as 0 (the reference leaves them unspecified).
LRB_HD void WriteHit(const SceneView &sc, const RayState &s, float rayMaxt, lrb_rayhit *hit) {
	if (s.hitRef == kNullIndex) {
		hit->t = rayMaxt;
		hit->b1 = 0.f; hit->b2 = 0.f;
		hit->meshIndex = kNullIndex;
		hit->triangleIndex = kNullIndex;
	} else {
		hit->t = s.maxt;
		hit->b1 = s.b1; hit->b2 = s.b2;
		const TriIds id = sc.ids[s.hitRef];     // (two-level: mesh index relative to the instance's leaf tree, which holds one mesh)
		hit->meshIndex = id.meshIndex + s.hitMeshOffset;
		hit->triangleIndex = id.triangleIndex;
	}
}

}   // namespace lrb

#endif
