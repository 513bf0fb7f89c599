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
#else
#define LRB_INF __builtin_huge_valf()
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
	uint32_t hitMesh, hitTri;
	uint32_t bestInst, bestTri;     // reference array order of the current best hit
	uint32_t curInstOrder, curMeshOffset;
	uint32_t cur;           // wide node to visit next, or kNullIndex when the stack must be popped
	bool inInstance;
};

LRB_HD float Dot3(float ax, float ay, float az, float bx, float by, float bz) {
	// vector.h:151-153: (x*x + y*y) + z*z
	return LRB_ADD(LRB_ADD(LRB_MUL(ax, bx), LRB_MUL(ay, by)), LRB_MUL(az, bz));
}

// Triangle::Intersect, triangle.h:55-89.  Returns true and t/b1/b2 when the triangle is hit
// within [mint, maxt].
LRB_HD bool TriangleTest(const RayState &r, const float4 a, const float4 b, const float p2z,
		float *tOut, float *b1Out, float *b2Out) {
	// TriRecord as three 16-B words: a = p0.xyz p1.x ; b = p1.yz p2.xy ; c = p2.z meshIndex triangleIndex order
	const float p0x = a.x, p0y = a.y, p0z = a.z;
	const float p1x = a.w, p1y = b.x, p1z = b.y;
	const float p2x = b.z, p2y = b.w;

	const float e1x = LRB_SUB(p1x, p0x), e1y = LRB_SUB(p1y, p0y), e1z = LRB_SUB(p1z, p0z);
	const float e2x = LRB_SUB(p2x, p0x), e2y = LRB_SUB(p2y, p0y), e2z = LRB_SUB(p2z, p0z);
	// s1 = Cross(d, e2)   (vector.h:159-163)
	const float s1x = LRB_SUB(LRB_MUL(r.dy, e2z), LRB_MUL(r.dz, e2y));
	const float s1y = LRB_SUB(LRB_MUL(r.dz, e2x), LRB_MUL(r.dx, e2z));
	const float s1z = LRB_SUB(LRB_MUL(r.dx, e2y), LRB_MUL(r.dy, e2x));

	const float divisor = Dot3(s1x, s1y, s1z, e1x, e1y, e1z);
	if (divisor == 0.f)
		return false;
	const float invDivisor = LRB_RCP(divisor);

	const float ddx = LRB_SUB(r.ox, p0x), ddy = LRB_SUB(r.oy, p0y), ddz = LRB_SUB(r.oz, p0z);
	const float b1 = LRB_MUL(Dot3(ddx, ddy, ddz, s1x, s1y, s1z), invDivisor);
	if (b1 < 0.f)
		return false;

	// s2 = Cross(dd, e1)
	const float s2x = LRB_SUB(LRB_MUL(ddy, e1z), LRB_MUL(ddz, e1y));
	const float s2y = LRB_SUB(LRB_MUL(ddz, e1x), LRB_MUL(ddx, e1z));
	const float s2z = LRB_SUB(LRB_MUL(ddx, e1y), LRB_MUL(ddy, e1x));
	const float b2 = LRB_MUL(Dot3(r.dx, r.dy, r.dz, s2x, s2y, s2z), invDivisor);
	if (b2 < 0.f)
		return false;

	const float b0 = LRB_SUB(LRB_SUB(1.f, b1), b2);
	if (b0 < 0.f)
		return false;

	const float t = LRB_MUL(Dot3(e2x, e2y, e2z, s2x, s2y, s2z), invDivisor);
	if (t < r.mint || t > r.maxt)
		return false;
	*tOut = t; *b1Out = b1; *b2Out = b2;
	return true;
}

// One slab of BBox::IntersectP (bbox.cpp:152-161) with the reference's exact select semantics.
LRB_HD void Slab(float lo, float hi, float o, float inv, float &t0, float &t1) {
	float tNear = LRB_MUL(LRB_SUB(lo, o), inv);
	float tFar = LRB_MUL(LRB_SUB(hi, o), inv);
	if (tNear > tFar) { const float s = tNear; tNear = tFar; tFar = s; }
	t0 = tNear > t0 ? tNear : t0;
	t1 = tFar < t1 ? tFar : t1;
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
			const float phi = acosf(cosPhi);
			const float sinPhi = sinf(phi);
			f1 = LRB_DIV(sinf(LRB_MUL(LRB_SUB(1.f, le), phi)), sinPhi);
			f2 = LRB_DIV(sinf(LRB_MUL(le, phi)), sinPhi);
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

// Initialise the state from a wire ray.  Returns false when there is nothing to traverse.
LRB_HD bool InitRay(const SceneView &sc, const lrb_ray &ray, RayState &s) {
	SetRay(s, ray.o[0], ray.o[1], ray.o[2], ray.d[0], ray.d[1], ray.d[2]);
	s.mint = ray.mint;
	s.maxt = ray.maxt;      // rayHit->t = ray->maxt
	s.time = ray.time;
	s.b1 = 0.f; s.b2 = 0.f;
	s.hitMesh = kNullIndex; s.hitTri = kNullIndex;
	// order 0 can never be beaten, so a hit at exactly t == ray.maxt is rejected as in the
	// reference (its `t < rayHit->t` fails against the initial rayHit->t = maxt)
	s.bestInst = 0; s.bestTri = 0;
	s.curInstOrder = 0; s.curMeshOffset = 0;
	s.inInstance = false;
	s.cur = sc.nWide ? sc.rootWide : kNullIndex;
	return sc.nWide != 0;
}

// One traversal step: visit s.cur (or pop).  Returns false when the ray is finished.
// STACK provides push(uint32_t node, float t0) / pop(uint32_t&, float&) / empty().
template <bool TWO_LEVEL, bool STATS, class STACK>
LRB_HD bool Step(const SceneView &sc, const lrb_ray &worldRay, RayState &s, STACK &stk, TraceStats *stats) {
	uint32_t cur = s.cur;
	if (cur == kNullIndex) {
		// pop until something is still worth visiting
		for (;;) {
			if (stk.empty())
				return false;
			float t0;
			stk.pop(cur, t0);
			if (TWO_LEVEL) {
				if (cur == kStackSentinel) {
					// leave the instance: back to the world-space ray (mbvhaccel.cpp:271-283)
					SetRay(s, worldRay.o[0], worldRay.o[1], worldRay.o[2], worldRay.d[0], worldRay.d[1], worldRay.d[2]);
					s.inInstance = false;
					continue;
				}
				if (cur & kTagInstance) {
					// enter a leaf tree (mbvhaccel.cpp:312-333)
					const uint4 ir = LRB_LDGU4(&sc.insts[cur & ~kTagInstance]);
					const uint4 ir2 = LRB_LDGU4(reinterpret_cast<const char *>(&sc.insts[cur & ~kTagInstance]) + 16);
					if (STATS) stats->instances++;
					if (ir.x == kNullIndex)
						continue;       // empty leaf tree
					if (ir.y != kNullIndex) {
						TransformRay(s, sc.minv + 16 * (size_t)ir.y, worldRay.o[0], worldRay.o[1], worldRay.o[2],
								worldRay.d[0], worldRay.d[1], worldRay.d[2]);
					} else if (ir.z != kNullIndex) {
						float m[16];
						MotionSample(sc, ir.z, s.time, m);
						if (STATS) stats->motionSamples++;
						TransformRayLocal(s, m, worldRay.o[0], worldRay.o[1], worldRay.o[2],
								worldRay.d[0], worldRay.d[1], worldRay.d[2]);
					}
					s.curMeshOffset = ir.w;
					s.curInstOrder = ir2.x;
					s.inInstance = true;
					stk.push(kStackSentinel, 0.f);
					cur = ir.x;
					break;
				}
			}
			// entry distance recorded at push time; a closer hit found since then culls the node
			// (same effect as running the box test now: t0 > min(maxt, tFar))
			if (t0 > s.maxt)
				continue;
			break;
		}
	}

	const WideNode *node = sc.nodes + cur;
	if (STATS) stats->wideNodes++;
	const float4 nminx = LRB_LDG4(node->bminx), nminy = LRB_LDG4(node->bminy), nminz = LRB_LDG4(node->bminz);
	const float4 nmaxx = LRB_LDG4(node->bmaxx), nmaxy = LRB_LDG4(node->bmaxy), nmaxz = LRB_LDG4(node->bmaxz);
	const uint4 kids = LRB_LDGU4(node->child);
	const uint4 meta = LRB_LDGU4(&node->leafBase);      // leafBase, counts, next, pad
	const uint32_t nInner = meta.y & 0xffu;
	const uint32_t nLeaf = meta.y >> 8;

	// ---- leaf children -------------------------------------------------------------------
	if (TWO_LEVEL && !s.inInstance) {
		// root tree: leaves are instances; they have no box of their own in the reference, so each
		// one is entered (mbvhaccel.cpp:312).  Defer them through the stack.
		for (uint32_t j = 0; j < nLeaf; ++j)
			stk.push(kTagInstance | (meta.x + j), 0.f);
	} else {
		for (uint32_t j = 0; j < nLeaf; ++j) {
			const char *tp = reinterpret_cast<const char *>(sc.tris + (meta.x + j));
			const float4 a = LRB_LDG4(tp), b = LRB_LDG4(tp + 16), cf = LRB_LDG4(tp + 32);
			const uint4 c = LRB_LDGU4(tp + 32);
			if (STATS) stats->triangles++;
			float t, b1, b2;
			if (TriangleTest(s, a, b, cf.x, &t, &b1, &b2)) {
				const uint32_t instOrder = TWO_LEVEL ? s.curInstOrder : 0u;
				const bool closer = t < s.maxt;
				const bool tieWin = (t == s.maxt) && (s.hitMesh != kNullIndex) &&
						(instOrder < s.bestInst || (instOrder == s.bestInst && c.w < s.bestTri));
				if (closer || tieWin) {
					s.maxt = t;
					s.b1 = b1; s.b2 = b2;
					s.hitMesh = TWO_LEVEL ? (c.y + s.curMeshOffset) : c.y;
					s.hitTri = c.z;
					s.bestInst = instOrder;
					s.bestTri = c.w;
				}
			}
		}
	}

	// ---- inner children: box tests, near-to-far ordering ------------------------------------
	float d0, d1, d2, d3;       // entry distances, +inf = not hit
	const float kInf = LRB_INF;
	{
		float t0 = s.mint, t1 = s.maxt;
		Slab(nminx.x, nmaxx.x, s.ox, s.ix, t0, t1); Slab(nminy.x, nmaxy.x, s.oy, s.iy, t0, t1); Slab(nminz.x, nmaxz.x, s.oz, s.iz, t0, t1);
		d0 = (nInner > 0 && !(t0 > t1)) ? t0 : kInf;
	}
	{
		float t0 = s.mint, t1 = s.maxt;
		Slab(nminx.y, nmaxx.y, s.ox, s.ix, t0, t1); Slab(nminy.y, nmaxy.y, s.oy, s.iy, t0, t1); Slab(nminz.y, nmaxz.y, s.oz, s.iz, t0, t1);
		d1 = (nInner > 1 && !(t0 > t1)) ? t0 : kInf;
	}
	{
		float t0 = s.mint, t1 = s.maxt;
		Slab(nminx.z, nmaxx.z, s.ox, s.ix, t0, t1); Slab(nminy.z, nmaxy.z, s.oy, s.iy, t0, t1); Slab(nminz.z, nmaxz.z, s.oz, s.iz, t0, t1);
		d2 = (nInner > 2 && !(t0 > t1)) ? t0 : kInf;
	}
	{
		float t0 = s.mint, t1 = s.maxt;
		Slab(nminx.w, nmaxx.w, s.ox, s.ix, t0, t1); Slab(nminy.w, nmaxy.w, s.oy, s.iy, t0, t1); Slab(nminz.w, nmaxz.w, s.oz, s.iz, t0, t1);
		d3 = (nInner > 3 && !(t0 > t1)) ? t0 : kInf;
	}
	// A box whose entry distance is +inf but which passed the test (mint = maxt = +inf) cannot
	// contain an acceptable hit closer than +inf; treating it as "not hit" is exact because the
	// triangle test rejects t > maxt and a tie at +inf never beats order 0.
	uint32_t c0 = kids.x, c1 = kids.y, c2 = kids.z, c3 = kids.w;

	// sorting network on (distance, child): ascending
#define LRB_CSWAP(da, ca, db, cb) { if (db < da) { const float td = da; da = db; db = td; const uint32_t tc = ca; ca = cb; cb = tc; } }
	LRB_CSWAP(d0, c0, d1, c1)
	LRB_CSWAP(d2, c2, d3, c3)
	LRB_CSWAP(d0, c0, d2, c2)
	LRB_CSWAP(d1, c1, d3, c3)
	LRB_CSWAP(d1, c1, d2, c2)
#undef LRB_CSWAP

	// continuation node (reference nodes with more than four inner children): always visited
	if (meta.z != kNullIndex)
		stk.push(meta.z, -kInf);
	// push far-to-near, keep the nearest
	if (d3 < kInf) stk.push(c3, d3);
	if (d2 < kInf) stk.push(c2, d2);
	if (d1 < kInf) stk.push(c1, d1);
	s.cur = (d0 < kInf) ? c0 : kNullIndex;
	if (STATS) { const unsigned long long d = stk.depth(); if (d > stats->maxStack) stats->maxStack = d; }
	return true;
}

// Final RayHit.  Miss payload: t = ray.maxt, meshIndex = triangleIndex = NULL_INDEX (bvh.cl:219-223);
// b1/b2 are written as 0 (the reference leaves them unspecified).
LRB_HD void WriteHit(const RayState &s, float rayMaxt, lrb_rayhit *hit) {
	if (s.hitMesh == kNullIndex) {
		hit->t = rayMaxt;
		hit->b1 = 0.f; hit->b2 = 0.f;
		hit->meshIndex = kNullIndex;
		hit->triangleIndex = kNullIndex;
	} else {
		hit->t = s.maxt;
		hit->b1 = s.b1; hit->b2 = s.b2;
		hit->meshIndex = s.hitMesh;
		hit->triangleIndex = s.hitTri;
	}
}

}   // namespace lrb

#endif
