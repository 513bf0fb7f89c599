// relayout_shared.h -- the per-node part of the re-layout (layout.h), written once for the host and for the device.
//
// relayout.cpp (host: any reference array that arrives through lrb_bvh_upload / lrb_mbvh_upload) and
// relayout_kernels.cuh (device: trees the GPU builder has just produced, lrb_bvh_build_scene) call the SAME functions
// for everything that decides a byte of a WideNode or a TriRecord: the triangle's build box, the slot order, the
// node grid, the gate.  Only IEEE double / float operations with one rounding each are used (no fused products whose
// contraction could differ: the library is built -fmad=false, the host tests -ffp-contract=off), so a scene laid out
// on the device is byte-identical to the host's lay-out of the same array (tests/test_relayout_device_cpu.py on the
// host build of these functions, tests/test_gpu_builder.py on the GPU).
// No exceptions, no allocation, no std:: -- error codes instead (RelayoutErrorString).
#ifndef LRB_RELAYOUT_SHARED_H
#define LRB_RELAYOUT_SHARED_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "luxrays_b200.h"
#include "layout.h"

#if defined(__CUDACC__)
#define LRB_RHD __host__ __device__ inline
#else
#define LRB_RHD inline
#endif

namespace lrb {

enum RelayoutError {
	kRelayoutOk = 0,
	kRelayoutBadMesh = 1,           // triangle leaf references a mesh outside the vertex-offset table
	kRelayoutBadVertex = 2,         // triangle leaf references a vertex outside the vertex buffer
	kRelayoutGridTooLarge = 3,      // box coordinates are too large for the node grid
	kRelayoutNonFiniteSlot = 4,     // internal error: a non-finite slot box reached the node grid
	kRelayoutTooManyKids = 5        // more children than the fixed-size path handles (the host falls back to its general path)
};

inline const char *RelayoutErrorString(int e) {
	switch (e) {
		case kRelayoutOk: return "ok";
		case kRelayoutBadMesh: return "triangle leaf references a mesh outside the vertex-offset table";
		case kRelayoutBadVertex: return "triangle leaf references a vertex outside the vertex buffer";
		case kRelayoutGridTooLarge: return "BVH box coordinates are too large for the node grid";
		case kRelayoutNonFiniteSlot: return "internal error: a non-finite slot box reached the node grid";
		case kRelayoutTooManyKids: return "internal error: node arity above the fixed-size re-layout path";
		default: return "unknown re-layout error";
	}
}

LRB_RHD bool RlIsLeaf(uint32_t nd) { return (nd & 0x80000000u) != 0; }
LRB_RHD uint32_t RlSkip(uint32_t nd) { return nd & 0x7fffffffu; }
LRB_RHD bool RlFinite(double v) { return fabs(v) < (double)__builtin_huge_val(); }     // false for NaN and +-inf
LRB_RHD bool RlFiniteF(float v) { return fabsf(v) < __builtin_huge_valf(); }
LRB_RHD double RlMin(double a, double b) { return b < a ? b : a; }                      // std::min / std::max semantics
LRB_RHD double RlMax(double a, double b) { return a < b ? b : a; }
LRB_RHD float RlMinF(float a, float b) { return b < a ? b : a; }
LRB_RHD float RlMaxF(float a, float b) { return a < b ? b : a; }

// MachineEpsilon::E (include/luxrays/core/epsilon.h:48-86) with the default clamp 1e-5 .. 1e-1.
LRB_RHD float EpsOf(float v) {
	uint32_t i;
	memcpy(&i, &v, 4);
	i += 0x80u;
	float f;
	memcpy(&f, &i, 4);
	const float e = fabsf(f - v);
	return e > 1e-5f ? (e < 1e-1f ? e : 1e-1f) : 1e-5f;
}

// Float boxes of the slots of one wide node, before they are put on the node's grid.
struct SlotBoxes {
	float lo[kWideSlots][3], hi[kWideSlots][3];
	bool whole[kWideSlots];         // MBVH root leaf: the slot covers the whole grid
	uint32_t child[kWideSlots];
	uint32_t n;
	bool hasOwn;                    // the node's own reference box (bounds every child, instances included)
	float ownLo[3], ownHi[3];
	LRB_RHD SlotBoxes() : n(0), hasOwn(false) {}
};

LRB_RHD uint32_t NodeSlots(const WideNode &w) { return w.exps >> 24; }

// Puts the slot boxes on the node's grid (layout.h): origin = min corner of everything the node
// bounds, power-of-two step per axis, lower planes rounded down and upper planes rounded up with a
// 1/64-step margin (the kernel's decode error is below 2^-9 step, traverse.h).
LRB_RHD int QuantizeNode(const SlotBoxes &b, uint32_t next, uint32_t flags, WideNode *w) {
	memset(w, 0, sizeof(*w));
	w->next = next;
	w->flags = flags;
	for (uint32_t k = 0; k < kWideSlots; ++k)
		w->child[k] = k < b.n ? b.child[k] : kNullIndex;
	uint32_t exps = b.n << 24;
	for (int a = 0; a < 3; ++a) {
		double lo = (double)__builtin_huge_val(), hi = -lo;
		if (b.hasOwn) { lo = b.ownLo[a]; hi = b.ownHi[a]; }
		for (uint32_t k = 0; k < b.n; ++k) {
			if (b.whole[k]) continue;
			lo = RlMin(lo, (double)b.lo[k][a]);
			hi = RlMax(hi, (double)b.hi[k][a]);
		}
		int eb;             // biased exponent byte of step * 2^kGridShift
		float org;
		uint32_t qlo = 0, qhi = 0;
		if (!(lo <= hi) || !RlFinite(lo) || !RlFinite(hi)) {
			// nothing finite to bound (a lone MBVH root leaf, or non-finite input): the axis never rejects.
			// Exponent byte 255 makes the decode scale +inf: A = +-inf, B = finite - A = -+inf, and every
			// plane distance fma(q, A, B) is inf - inf = NaN, which the slab test ignores -- for every ray,
			// with no overflow edge (a huge finite grid fails where A is finite but B overflows).
			org = 0.f;
			eb = 255;
			for (uint32_t k = 0; k < kWideSlots; ++k) {
				qlo |= (k < b.n ? 0u : 255u) << (8 * k);
				qhi |= (k < b.n ? 255u : 0u) << (8 * k);
			}
		} else {
			// smallest power-of-two step with 250 steps >= extent, but never finer than 4 ulp of the
			// largest coordinate (the grid origin is a float)
			const double ext = hi - lo;
			int E = -140, e2;
			if (ext > 0.0) {
				const double m = frexp(ext / 250.0, &e2);      // ext / 250 = m * 2^e2, m in [0.5, 1)
				E = (m == 0.5) ? e2 - 1 : e2;
			}
			const double mag = RlMax(fabs(lo), fabs(hi));
			if (mag > 0.0) {
				frexp(mag, &e2);
				E = E < e2 - 24 + 2 ? e2 - 24 + 2 : E;
			}
			E = E < 1 - kGridShift - 127 ? 1 - kGridShift - 127 : E;
			for (;; ++E) {
				eb = E + kGridShift + 127;
				if (eb > 254)
					return kRelayoutGridTooLarge;
				const double step = ldexp(1.0, E);
				// the origin sits 1.5 steps below the lowest plane: every plane keeps its outward margin
				// (no plane is clamped at 0 or 255), whole-grid slots extend past the node's own box
				org = (float)(lo - 1.5 * step);
				bool ok = RlFiniteF(org);
				qlo = qhi = 0;
				for (uint32_t k = 0; k < kWideSlots && ok; ++k) {
					uint32_t l = 255, h = 0;        // unused slot: inverted
					if (k < b.n) {
						if (b.whole[k]) {
							l = 0; h = 255;
						} else {
							const double xl = ((double)b.lo[k][a] - (double)org) / step, xh = ((double)b.hi[k][a] - (double)org) / step;
							const double fl = floor(xl - 1.0 / 64.0), ch = ceil(xh + 1.0 / 64.0);
							if (!(fl >= 0.0) || !(ch <= 255.0) || !(fl <= ch)) {
								ok = false;     // needs a coarser grid (or the child box is not finite)
								break;
							}
							l = (uint32_t)fl;
							h = (uint32_t)ch;
						}
					}
					qlo |= l << (8 * k);
					qhi |= h << (8 * k);
				}
				if (ok)
					break;
				bool finite = true;
				for (uint32_t k = 0; k < b.n; ++k)
					if (!b.whole[k] && (!RlFiniteF(b.lo[k][a]) || !RlFiniteF(b.hi[k][a]) || !(b.lo[k][a] <= b.hi[k][a])))
						finite = false;
				if (!finite)
					return kRelayoutNonFiniteSlot;
			}
		}
		w->org[a] = org;
		w->qlo[a] = qlo;
		w->qhi[a] = qhi;
		exps |= (uint32_t)eb << (8 * a);
	}
	w->exps = exps;
	return kRelayoutOk;
}

// The box BVHAccel::Init gives the builders for one triangle (bvhaccel.cpp:116-122): bounds of the
// three vertices, grown by MachineEpsilon::E of the bounds.
LRB_RHD void TriBuildBoxOf(const float *p0, const float *p1, const float *p2, float lo[3], float hi[3]) {
	float e = 0.f;
	for (int k = 0; k < 3; ++k) {
		lo[k] = RlMinF(RlMinF(p0[k], p1[k]), p2[k]);
		hi[k] = RlMaxF(RlMaxF(p0[k], p1[k]), p2[k]);
		e = RlMaxF(e, RlMaxF(EpsOf(lo[k]), EpsOf(hi[k])));
	}
	for (int k = 0; k < 3; ++k) {
		lo[k] -= e;
		hi[k] += e;
	}
}
LRB_RHD void TriBuildBox(const TriRecord &tr, float lo[3], float hi[3]) { TriBuildBoxOf(tr.p0, tr.p1, tr.p2, lo, hi); }

// Boxes the reference's builders never emit but its traversal tolerates: BBox::IntersectP swaps the two
// slab distances when they come out of order, so a box with min > max behaves like the sorted box, and
// every comparison with a NaN is false, so a NaN plane never rejects.  Same behaviour here: corners
// are sorted per axis, a box with a non-finite corner covers the whole grid of its node.
LRB_RHD void SanitizeSlot(SlotBoxes *b, uint32_t k) {
	for (int a = 0; a < 3; ++a) {
		if (!RlFiniteF(b->lo[k][a]) || !RlFiniteF(b->hi[k][a])) {
			b->whole[k] = true;
			return;
		}
		if (b->lo[k][a] > b->hi[k][a]) {
			const float t = b->lo[k][a];
			b->lo[k][a] = b->hi[k][a];
			b->hi[k][a] = t;
		}
	}
}

// What a triangle tree's re-layout reads: the reference array, the vertices of all meshes back to back, first vertex per mesh.
struct TriTreeView {
	const lrb_bvh_node *nodes;
	uint32_t n;
	const float *xyz;
	uint64_t nVerts;
	const uint32_t *meshOff;
	uint32_t nMeshes;
};

// Triangle record of reference leaf c (vertices gathered, gate open, order = c) and its ids.
LRB_RHD int FillTriOf(const TriTreeView &in, uint32_t c, TriRecord *tr, TriIds *ids) {
	const lrb_bvh_node &nd = in.nodes[c];
	const uint32_t mesh = nd.triangleLeaf.meshIndex;
	if (mesh >= in.nMeshes)
		return kRelayoutBadMesh;
	const float *p[3];
	for (int j = 0; j < 3; ++j) {
		const uint64_t g = (uint64_t)nd.triangleLeaf.v[j] + in.meshOff[mesh];
		if (g >= in.nVerts)
			return kRelayoutBadVertex;
		p[j] = in.xyz + 3 * g;
	}
	for (int k = 0; k < 3; ++k) {
		tr->p0[k] = p[0][k];
		tr->p1[k] = p[1][k];
		tr->p2[k] = p[2][k];
	}
	for (int k = 0; k < 3; ++k) {
		tr->gateLo[k] = -__builtin_huge_valf();
		tr->gateHi[k] = __builtin_huge_valf();
	}
	tr->order = c;
	if (ids) {
		ids->meshIndex = mesh;
		ids->triangleIndex = nd.triangleLeaf.triangleIndex;
	}
	return kRelayoutOk;
}

// Half the surface area of a box (slot order = ascending size; NaN boxes last).
LRB_RHD double BoxSizeKey(const float lo[3], const float hi[3]) {
	const double dx = fabs((double)hi[0] - lo[0]), dy = fabs((double)hi[1] - lo[1]), dz = fabs((double)hi[2] - lo[2]);
	const double a = dx * dy + dy * dz + dz * dx;
	return a == a ? a : (double)__builtin_huge_val();
}

// Number of children of inner reference node i (walk over the skip links) and of the wide nodes it becomes.
LRB_RHD uint32_t CountKids(const lrb_bvh_node *nodes, uint32_t i) {
	const uint32_t end = RlSkip(nodes[i].nodeData);
	uint32_t nKids = 0;
	for (uint32_t c = i + 1; c < end; c = RlSkip(nodes[c].nodeData))
		++nKids;
	return nKids;
}
LRB_RHD uint32_t WideNodesFor(uint32_t nKids) {
	const uint32_t w = (nKids + kWideSlots - 1) / kWideSlots;
	return w ? w : 1u;
}

static const uint32_t kFixedMaxKids = 8;    // accelerator.bvh.treetype is 2, 4 or 8 (bvhaccel.cpp:51)

// The wide node(s) and the triangle records of ONE inner node of a triangle tree -- the body of the re-layout's second
// pass.  wideOf[c] = wide-node index of an inner child / TriRecord index of a leaf child (both fixed by the first pass, in
// reference order).  `parentOf` (optional, device stack-need pass): parentOf[w] = the wide node whose visit pushes w.
// Slot order = ascending box size: the kernel sorts the children of a node by entry distance with a network that keeps
// the slot order of equal keys, and for a ray that STARTS inside several child boxes (every bounce ray does, near the
// root) all those keys equal ray.mint: visiting the smaller box first finds a near hit sooner and culls more of the
// rest (kitchen, bounce-2 rays: 16.2 -> 15.6 node visits and 5.4 -> 4.8 triangle tests per ray).  Order never changes
// a result.
LRB_RHD int ConvertInnerNodeTri(const TriTreeView &in, const uint32_t i, const uint32_t *wideOf, WideNode *wide, TriRecord *tris,
		TriIds *ids, uint32_t *parentOf) {
	const lrb_bvh_node *nodes = in.nodes;
	uint32_t kids[kFixedMaxKids];
	double keys[kFixedMaxKids];
	uint32_t nKids = 0;
	const uint32_t end = RlSkip(nodes[i].nodeData);
	for (uint32_t c = i + 1; c < end; c = RlSkip(nodes[c].nodeData)) {
		if (nKids == kFixedMaxKids)
			return kRelayoutTooManyKids;
		kids[nKids++] = c;
	}
	if (nKids > 1) {
		for (uint32_t k = 0; k < nKids; ++k) {
			const lrb_bvh_node &ch = nodes[kids[k]];
			float lo[3], hi[3];
			if (!RlIsLeaf(ch.nodeData)) {
				for (int a = 0; a < 3; ++a) { lo[a] = ch.bvhNode.bboxMin[a]; hi[a] = ch.bvhNode.bboxMax[a]; }
			} else {
				TriRecord tr;
				const int rc = FillTriOf(in, kids[k], &tr, nullptr);
				if (rc != kRelayoutOk)
					return rc;
				TriBuildBox(tr, lo, hi);
			}
			keys[k] = BoxSizeKey(lo, hi);
		}
		// stable insertion sort (== std::stable_sort on the keys)
		for (uint32_t k = 1; k < nKids; ++k) {
			const double key = keys[k];
			const uint32_t c = kids[k];
			uint32_t j = k;
			while (j > 0 && key < keys[j - 1]) {
				keys[j] = keys[j - 1];
				kids[j] = kids[j - 1];
				--j;
			}
			keys[j] = key;
			kids[j] = c;
		}
	}
	const uint32_t nW = WideNodesFor(nKids);
	const uint32_t w0 = wideOf[i];
	for (uint32_t j = 0; j < nW; ++j) {
		const uint32_t first = j * kWideSlots;
		const uint32_t left = nKids > first ? nKids - first : 0u;
		const uint32_t cnt = left < kWideSlots ? left : kWideSlots;
		SlotBoxes b;
		// the node's own box bounds every child
		b.hasOwn = true;
		for (int a = 0; a < 3; ++a) { b.ownLo[a] = nodes[i].bvhNode.bboxMin[a]; b.ownHi[a] = nodes[i].bvhNode.bboxMax[a]; }
		for (uint32_t s = 0; s < cnt; ++s) {
			const uint32_t c = kids[first + s];
			const lrb_bvh_node &ch = nodes[c];
			const uint32_t k = b.n++;
			b.whole[k] = false;
			if (!RlIsLeaf(ch.nodeData)) {
				for (int a = 0; a < 3; ++a) { b.lo[k][a] = ch.bvhNode.bboxMin[a]; b.hi[k][a] = ch.bvhNode.bboxMax[a]; }
				b.child[k] = wideOf[c];
				SanitizeSlot(&b, k);
				if (parentOf)
					parentOf[wideOf[c]] = w0 + j;
			} else {
				TriRecord tr;
				TriIds id;
				const int rc = FillTriOf(in, c, &tr, &id);
				if (rc != kRelayoutOk)
					return rc;
				TriBuildBox(tr, b.lo[k], b.hi[k]);
				SanitizeSlot(&b, k);
				// the reference's gate for this triangle: its parent's exact box
				for (int a = 0; a < 3; ++a) {
					tr.gateLo[a] = nodes[i].bvhNode.bboxMin[a];
					tr.gateHi[a] = nodes[i].bvhNode.bboxMax[a];
					if (tr.gateLo[a] > tr.gateHi[a]) {      // see SanitizeSlot: the reference's slab swap
						const float t = tr.gateLo[a];
						tr.gateLo[a] = tr.gateHi[a];
						tr.gateHi[a] = t;
					}
				}
				const uint32_t ti = wideOf[c];
				b.child[k] = kTagTri | ti;
				tris[ti] = tr;
				ids[ti] = id;
			}
		}
		const int rc = QuantizeNode(b, (j + 1 < nW) ? (w0 + j + 1) : kNullIndex, 0, &wide[w0 + j]);
		if (rc != kRelayoutOk)
			return rc;
		if (parentOf && j + 1 < nW)
			parentOf[w0 + j + 1] = w0 + j;
	}
	return kRelayoutOk;
}

// The one-child entry node in front of a tree whose root is an inner node: it carries the root's own box
// (bvhaccel.cpp:245-255 with currentNode == 0), which no parent holds.  entryBox = that box, corners sorted.
LRB_RHD int MakeEntryNode(const lrb_bvh_node &root, uint32_t rootWide, WideNode *w, float entryBox[6]) {
	SlotBoxes b;
	b.n = 1;
	b.whole[0] = false;
	for (int a = 0; a < 3; ++a) { b.lo[0][a] = root.bvhNode.bboxMin[a]; b.hi[0][a] = root.bvhNode.bboxMax[a]; }
	b.child[0] = rootWide;
	SanitizeSlot(&b, 0);
	const int rc = QuantizeNode(b, kNullIndex, kNodeEntry, w);
	if (entryBox) {
		for (int a = 0; a < 3; ++a) {
			entryBox[a] = root.bvhNode.bboxMin[a];
			entryBox[3 + a] = root.bvhNode.bboxMax[a];
			if (entryBox[a] > entryBox[3 + a]) {
				const float t = entryBox[a];
				entryBox[a] = entryBox[3 + a];
				entryBox[3 + a] = t;
			}
		}
	}
	return rc;
}

// Worst-case live stack entries below wide node w, given the largest such number among what its slots lead to
// (relayout.cpp, the sweep at the end of ConvertTree): ALL FOUR slots count, used or not, plus the continuation.
LRB_RHD uint32_t StackNeedOfNode(const WideNode &w, uint32_t below) {
	const uint32_t k = kWideSlots + (w.next != kNullIndex ? 1u : 0u);
	return (k - 1) + below;
}

}   // namespace lrb

#endif
