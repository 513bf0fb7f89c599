// relayout.cpp -- see relayout.h.  Converts the reference's depth-first skip-list arrays
// (layout rules: src/luxrays/core/bvh/bvhclassicbuild.cpp:181-220 -- a node's first child is
// index+1, siblings are reached through the skip index, a leaf's skip index is index+1) into
// wide nodes.  The host side replaces the per-node index rewriting the reference does before
// upload (bvhaccelhw.cpp:126-145, mbvhaccelhw.cpp:380-427).

#include "relayout.h"
#include "relayout_shared.h"  // the per-node functions shared with the device re-layout (relayout_kernels.cuh)
#include "traverse.h"       // MotionSample: the kernels' own MotionSystem::Sample, compiled for the host

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <cstdlib>
#include <exception>
#include <functional>
#include <stdexcept>
#include <thread>

namespace lrb {

static inline bool IsLeaf(uint32_t nd) { return (nd & 0x80000000u) != 0; }
static inline uint32_t Skip(uint32_t nd) { return nd & 0x7fffffffu; }

bool ValidateTree(const lrb_bvh_node *nodes, uint32_t n, std::string *err) {
	if (n == 0)
		return true;
	if (!nodes) { if (err) *err = "null node array"; return false; }
	if (n >= 0x7fffffffu) { if (err) *err = "node count does not fit 31 bits"; return false; }
	if (Skip(nodes[0].nodeData) != n) {
		if (err) *err = "root skip index != node count";
		return false;
	}
	// Every node's skip index must lie inside its parent's range.  Walk with an explicit stack
	// of enclosing range ends.
	std::vector<uint32_t> ends;
	ends.push_back(n);
	for (uint32_t i = 0; i < n; ++i) {
		while (!ends.empty() && ends.back() == i)
			ends.pop_back();
		if (ends.empty()) { if (err) *err = "node outside every subtree"; return false; }
		const uint32_t nd = nodes[i].nodeData;
		const uint32_t s = Skip(nd);
		if (IsLeaf(nd)) {
			if (s != i + 1) { if (err) *err = "leaf skip index != index + 1"; return false; }
		} else {
			if (s <= i + 1 || s > ends.back()) { if (err) *err = "inner skip index out of range"; return false; }
			ends.push_back(s);
		}
	}
	return true;
}

namespace {

struct TreeInput {
	const lrb_bvh_node *nodes;
	uint32_t n;
	bool instLeaves;                // MBVH root tree
	// triangle trees
	const float *xyz;
	uint64_t nVerts;
	const uint32_t *meshOff;
	uint32_t nMeshes;
	// root tree
	const std::vector<uint32_t> *leafRootWide;
	const std::vector<uint32_t> *leafStackNeed;
	uint32_t nTransforms, nMotions;
	const std::vector<float> *leafBox;      // 6 floats per unique leaf: its tree's root box (instance space)
	const float *minv;                      // 16 floats per transform
	const std::vector<uint32_t> *motionFirst, *motionLast;
	const std::vector<DevInterp> *interps;
};

static inline TriTreeView TriViewOf(const TreeInput &in) {
	TriTreeView v;
	v.nodes = in.nodes; v.n = in.n; v.xyz = in.xyz; v.nVerts = in.nVerts; v.meshOff = in.meshOff; v.nMeshes = in.nMeshes;
	return v;
}

static inline void ThrowIfRelayoutError(const int rc) {
	if (rc != kRelayoutOk)
		throw std::runtime_error(RelayoutErrorString(rc));
}

static void FillTri(const TreeInput &in, uint32_t c, TriRecord *tr, TriIds *ids = nullptr) {
	ThrowIfRelayoutError(FillTriOf(TriViewOf(in), c, tr, ids));
}

static void FillInst(const TreeInput &in, uint32_t c, InstRecord *ir) {
	const lrb_bvh_node &nd = in.nodes[c];
	if (nd.bvhLeaf.leafIndex >= in.leafRootWide->size())
		throw std::runtime_error("MBVH root leaf references a leaf tree that was not supplied");
	if (nd.bvhLeaf.transformIndex != kNullIndex && nd.bvhLeaf.transformIndex >= in.nTransforms)
		throw std::runtime_error("MBVH root leaf references a transform that was not supplied");
	if (nd.bvhLeaf.motionIndex != kNullIndex && nd.bvhLeaf.motionIndex >= in.nMotions)
		throw std::runtime_error("MBVH root leaf references a motion system that was not supplied");
	memset(ir, 0, sizeof(*ir));
	ir->rootWide = (*in.leafRootWide)[nd.bvhLeaf.leafIndex];
	ir->transformIndex = nd.bvhLeaf.transformIndex;
	ir->motionIndex = nd.bvhLeaf.motionIndex;
	ir->meshOffset = nd.bvhLeaf.meshOffsetIndex;
	ir->order = c;
	ir->pad[0] = (*in.leafStackNeed)[nd.bvhLeaf.leafIndex];   // host-side bookkeeping only
}

static const float kInfF = std::numeric_limits<float>::infinity();

// EpsOf, SlotBoxes, NodeSlots, QuantizeNode, TriBuildBox, SanitizeSlot: relayout_shared.h
static void QuantizeNodeOrThrow(const SlotBoxes &b, uint32_t next, uint32_t flags, WideNode *w) {
	ThrowIfRelayoutError(QuantizeNode(b, next, flags, w));
}

// 4x4 inverse in double precision (Gauss-Jordan with partial pivoting); false when singular.
static bool Invert4x4(const float *m, double out[16]) {
	double a[4][8];
	for (int r = 0; r < 4; ++r)
		for (int c = 0; c < 4; ++c) {
			a[r][c] = m[4 * r + c];
			a[r][4 + c] = (r == c) ? 1.0 : 0.0;
		}
	for (int col = 0; col < 4; ++col) {
		int piv = col;
		for (int r = col + 1; r < 4; ++r)
			if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
		if (!(fabs(a[piv][col]) > 1e-300) || !std::isfinite(a[piv][col]))
			return false;
		if (piv != col)
			for (int c = 0; c < 8; ++c) std::swap(a[piv][c], a[col][c]);
		const double inv = 1.0 / a[col][col];
		for (int c = 0; c < 8; ++c) a[col][c] *= inv;
		for (int r = 0; r < 4; ++r) {
			if (r == col) continue;
			const double f = a[r][col];
			if (f != 0.0)
				for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
		}
	}
	for (int r = 0; r < 4; ++r)
		for (int c = 0; c < 4; ++c) {
			out[4 * r + c] = a[r][4 + c];
			if (!std::isfinite(out[4 * r + c])) return false;
		}
	return true;
}

// World-space bounds of one instance: the leaf tree's root box (instance space) taken through the
// inverse of mInv, grown generously.  The reference enters every instance of a visited root node and
// then tests the same root box in instance space (mbvhaccel.cpp:312-333 -> bvhaccel.cpp:245-255); a ray
// that misses the world-space bounds of that box misses the box itself, so skipping the instance
// changes nothing.  Motion-blurred instances: MotionWorldBox below.  Returns false (=> the slot takes the
// whole grid) for singular or projective matrices and non-finite results.
static bool MotionWorldBox(const TreeInput &in, const lrb_bvh_node &nd, const float *lb, float lo[3], float hi[3]);

// Rounding of the reference's own ray transform.  The reference tests the leaf tree's root box with the
// float ray  mInv * o, mInv * d : each component carries an error of a few ulp of the terms it sums, i.e.
// ~ eps * |mInv| * (|o| + |translation|) in instance space, which is  eps * cond(M) * (|o| + |t|)  back in
// world space (cond = |M| |mInv|, infinity norms of the 3x3 parts).  A ray that misses the exact world
// bounds by less than that can still reach a silhouette triangle in the reference, so the world box is
// grown by it, taking for |o| and for the distance travelled four times the largest coordinate of the
// dataset's root box: rays that start farther than that from the origin are outside what this margin
// covers (supported range; stated in DESIGN.md "Parity").
static double SceneMagnitude(const TreeInput &in) {
	double m = 0.0;
	if (in.n && !IsLeaf(in.nodes[0].nodeData))
		for (int a = 0; a < 3; ++a) {
			const double lo = in.nodes[0].bvhNode.bboxMin[a], hi = in.nodes[0].bvhNode.bboxMax[a];
			if (std::isfinite(lo)) m = std::max(m, fabs(lo));
			if (std::isfinite(hi)) m = std::max(m, fabs(hi));
		}
	return m;
}

static double Norm3x3(const double *M) {
	double n = 0.0;
	for (int r = 0; r < 3; ++r)
		n = std::max(n, fabs(M[4 * r]) + fabs(M[4 * r + 1]) + fabs(M[4 * r + 2]));
	return n;
}

static double TransformRoundingMargin(const TreeInput &in, const double *M, const double *Minv, const double mag) {
	const double cond = std::min(Norm3x3(M) * Norm3x3(Minv), 1e12);
	return 16.0 * 5.9604644775390625e-08 * cond * (4.0 * SceneMagnitude(in) + mag);
}

static bool InstanceWorldBox(const TreeInput &in, const lrb_bvh_node &nd, float lo[3], float hi[3]) {
	if (!in.leafBox)
		return false;
	const float *lb = in.leafBox->data() + 6 * (size_t)nd.bvhLeaf.leafIndex;
	for (int k = 0; k < 6; ++k)
		if (!std::isfinite(lb[k])) return false;
	if (nd.bvhLeaf.motionIndex != kNullIndex)
		return MotionWorldBox(in, nd, lb, lo, hi);
	double M[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	double Minv[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	if (nd.bvhLeaf.transformIndex != kNullIndex) {
		if (!in.minv || !Invert4x4(in.minv + 16 * (size_t)nd.bvhLeaf.transformIndex, M))
			return false;
		for (int k = 0; k < 16; ++k) Minv[k] = in.minv[16 * (size_t)nd.bvhLeaf.transformIndex + k];
		if (fabs(M[12]) > 1e-12 || fabs(M[13]) > 1e-12 || fabs(M[14]) > 1e-12 || fabs(M[15] - 1.0) > 1e-9)
			return false;       // projective: the bounds of the corners do not bound the box
	}
	double wlo[3] = { 1e300, 1e300, 1e300 }, whi[3] = { -1e300, -1e300, -1e300 }, mag = 0.0;
	for (int corner = 0; corner < 8; ++corner) {
		const double p[3] = { lb[(corner & 1) ? 3 : 0], lb[(corner & 2) ? 4 : 1], lb[(corner & 4) ? 5 : 2] };
		for (int r = 0; r < 3; ++r) {
			const double v = M[4 * r] * p[0] + M[4 * r + 1] * p[1] + M[4 * r + 2] * p[2] + M[4 * r + 3];
			wlo[r] = std::min(wlo[r], v);
			whi[r] = std::max(whi[r], v);
			mag = std::max(mag, fabs(v));
		}
	}
	const double diag = std::max(std::max(whi[0] - wlo[0], whi[1] - wlo[1]), whi[2] - wlo[2]);
	const double grow = 1e-4 * diag + 4e-6 * mag + 1e-6 + TransformRoundingMargin(in, M, Minv, mag);
	for (int r = 0; r < 3; ++r) {
		lo[r] = (float)(wlo[r] - grow);
		hi[r] = (float)(whi[r] + grow);
		if (!std::isfinite(lo[r]) || !std::isfinite(hi[r])) return false;
		lo[r] = std::nextafter(lo[r], -kInfF);
		hi[r] = std::nextafter(hi[r], kInfF);
	}
	return true;
}

// World-space bounds of a motion-blurred instance over ALL times: the leaf tree's root box taken through
// the inverse of the sampled world->instance matrix (the kernels' own MotionSample) at 32 times per
// interpolation segment plus every segment boundary, grown by 1.25 x the largest displacement of a box
// corner between two consecutive samples -- a point of the moving box at a time between two samples is
// never farther from its sampled positions than the path it travels between them -- plus the margins
// of InstanceWorldBox.  (The reference's own root-tree boxes are unions over 1 025 samples with no margin,
// motionsystem.cpp:75-88.)  Times outside the motion system's range clamp to its first / last key.
static bool MotionWorldBox(const TreeInput &in, const lrb_bvh_node &nd, const float *lb, float lo[3], float hi[3]) {
	if (!in.motionFirst || !in.motionLast || !in.interps || nd.bvhLeaf.motionIndex >= in.motionFirst->size())
		return false;
	SceneView sv;
	memset(&sv, 0, sizeof(sv));
	sv.motionFirst = in.motionFirst->data();
	sv.motionLast = in.motionLast->data();
	sv.interps = in.interps->data();
	const uint32_t first = (*in.motionFirst)[nd.bvhLeaf.motionIndex], last = (*in.motionLast)[nd.bvhLeaf.motionIndex];
	std::vector<float> times;
	for (uint32_t i = first; i <= last && i < in.interps->size(); ++i) {
		const DevInterp &it = (*in.interps)[i];
		if (std::isfinite(it.startTime)) times.push_back(it.startTime);
		if (std::isfinite(it.endTime)) times.push_back(it.endTime);
	}
	if (times.empty())
		times.push_back(0.f);
	std::sort(times.begin(), times.end());
	times.erase(std::unique(times.begin(), times.end()), times.end());
	const size_t nKeys = times.size();
	const int perSegment = 32;
	for (size_t k = 0; k + 1 < nKeys; ++k)
		for (int j = 1; j < perSegment; ++j)
			times.push_back((float)((double)times[k] + ((double)times[k + 1] - (double)times[k]) * j / perSegment));
	std::sort(times.begin(), times.end());

	double wlo[3] = { 1e300, 1e300, 1e300 }, whi[3] = { -1e300, -1e300, -1e300 }, mag = 0.0, step = 0.0, rounding = 0.0;
	double prev[8][3];
	bool havePrev = false;
	for (size_t ti = 0; ti < times.size(); ++ti) {
		float m[16];
		MotionSample(sv, nd.bvhLeaf.motionIndex, times[ti], m);
		double M[16];
		if (!Invert4x4(m, M))
			return false;
		if (fabs(M[12]) > 1e-12 || fabs(M[13]) > 1e-12 || fabs(M[14]) > 1e-12 || fabs(M[15] - 1.0) > 1e-9)
			return false;
		{
			double md[16];
			for (int k = 0; k < 16; ++k) md[k] = m[k];
			rounding = std::max(rounding, TransformRoundingMargin(in, M, md, 0.0));
		}
		for (int corner = 0; corner < 8; ++corner) {
			const double p[3] = { lb[(corner & 1) ? 3 : 0], lb[(corner & 2) ? 4 : 1], lb[(corner & 4) ? 5 : 2] };
			double w[3], d2 = 0.0;
			for (int r = 0; r < 3; ++r) {
				w[r] = M[4 * r] * p[0] + M[4 * r + 1] * p[1] + M[4 * r + 2] * p[2] + M[4 * r + 3];
				if (!std::isfinite(w[r])) return false;
				wlo[r] = std::min(wlo[r], w[r]);
				whi[r] = std::max(whi[r], w[r]);
				mag = std::max(mag, fabs(w[r]));
				if (havePrev) d2 += (w[r] - prev[corner][r]) * (w[r] - prev[corner][r]);
				prev[corner][r] = w[r];
			}
			step = std::max(step, sqrt(d2));
		}
		havePrev = true;
	}
	const double diag = std::max(std::max(whi[0] - wlo[0], whi[1] - wlo[1]), whi[2] - wlo[2]);
	const double grow = 1.25 * step + 1e-4 * diag + 4e-6 * mag + 1e-6 + rounding + 16.0 * 5.9604644775390625e-08 * mag;
	for (int r = 0; r < 3; ++r) {
		lo[r] = (float)(wlo[r] - grow);
		hi[r] = (float)(whi[r] + grow);
		if (!std::isfinite(lo[r]) || !std::isfinite(hi[r])) return false;
		lo[r] = std::nextafter(lo[r], -kInfF);
		hi[r] = std::nextafter(hi[r], kInfF);
	}
	return true;
}

// Half the surface area of the box slot `c` will get (+inf for a slot that covers the whole grid).
static double SlotSizeKey(const TreeInput &in, uint32_t c) {
	const lrb_bvh_node &ch = in.nodes[c];
	float lo[3], hi[3];
	if (!IsLeaf(ch.nodeData)) {
		for (int a = 0; a < 3; ++a) { lo[a] = ch.bvhNode.bboxMin[a]; hi[a] = ch.bvhNode.bboxMax[a]; }
	} else if (in.instLeaves) {
		if (!InstanceWorldBox(in, ch, lo, hi))
			return std::numeric_limits<double>::infinity();
	} else {
		TriRecord tr;
		FillTri(in, c, &tr);
		TriBuildBox(tr, lo, hi);
	}
	const double dx = fabs((double)hi[0] - lo[0]), dy = fabs((double)hi[1] - lo[1]), dz = fabs((double)hi[2] - lo[2]);
	const double a = dx * dy + dy * dz + dz * dx;
	return a == a ? a : std::numeric_limits<double>::infinity();     // NaN boxes last
}

// Adds reference child `c` (inner node, triangle leaf or MBVH root leaf) as the next slot of `b`.  wideOf[c] is the
// wide-node index of an inner child and the (pre-assigned, reference-ordered) TriRecord index of a triangle leaf.
static void AddSlot(const TreeInput &in, const std::vector<uint32_t> &wideOf, uint32_t c, const lrb_bvh_node *parent,
		SlotBoxes *b, WideScene *out) {
	const lrb_bvh_node &ch = in.nodes[c];
	const uint32_t k = b->n++;
	b->whole[k] = false;
	if (!IsLeaf(ch.nodeData)) {
		for (int a = 0; a < 3; ++a) { b->lo[k][a] = ch.bvhNode.bboxMin[a]; b->hi[k][a] = ch.bvhNode.bboxMax[a]; }
		b->child[k] = wideOf[c];
		SanitizeSlot(b, k);
	} else if (in.instLeaves) {
		b->whole[k] = !InstanceWorldBox(in, ch, b->lo[k], b->hi[k]);
		InstRecord ir;
		FillInst(in, c, &ir);
		b->child[k] = kTagInstance | (uint32_t)out->insts.size();
		out->insts.push_back(ir);
	} else {
		TriRecord tr;
		TriIds ids;
		FillTri(in, c, &tr, &ids);
		TriBuildBox(tr, b->lo[k], b->hi[k]);
		SanitizeSlot(b, k);
		// the reference's gate for this triangle: its parent's exact box (none for a root that is a leaf)
		for (int a = 0; a < 3; ++a) {
			tr.gateLo[a] = parent ? parent->bvhNode.bboxMin[a] : -kInfF;
			tr.gateHi[a] = parent ? parent->bvhNode.bboxMax[a] : kInfF;
			if (tr.gateLo[a] > tr.gateHi[a])    // see SanitizeSlot: the reference's slab swap
				std::swap(tr.gateLo[a], tr.gateHi[a]);
		}
		const uint32_t ti = wideOf[c];
		b->child[k] = kTagTri | ti;
		out->tris[ti] = tr;
		out->ids[ti] = ids;
	}
}

// Appends the wide form of one reference tree to `out`.  Returns the wide index of its root
// (kNullIndex for an empty tree) and the tree's worst-case stack need.
// development switch LRB_RELAYOUT_VERBOSE=1: wall time of the phases of a conversion on stderr
struct PhaseTimer {
	bool on;
	std::chrono::steady_clock::time_point t;
	PhaseTimer() : on(getenv("LRB_RELAYOUT_VERBOSE") != nullptr), t(std::chrono::steady_clock::now()) {}
	void operator()(const char *what) {
		if (!on) return;
		const std::chrono::steady_clock::time_point n = std::chrono::steady_clock::now();
		fprintf(stderr, "luxrays_b200 re-layout: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
		t = n;
	}
};

static uint32_t ConvertTree(const TreeInput &in, WideScene *out, uint32_t *stackNeed) {
	*stackNeed = 0;
	if (in.n == 0)
		return kNullIndex;
	PhaseTimer phase;
	std::string err;
	if (!ValidateTree(in.nodes, in.n, &err))
		throw std::runtime_error("malformed BVHArrayNode array: " + err);

	phase("validate");
	const uint32_t wideStart = (uint32_t)out->wide.size();
	const lrb_bvh_node *nodes = in.nodes;
	std::vector<uint32_t> wideOf(in.n, kNullIndex);

	// The root itself is a leaf (one-triangle mesh / one-mesh dataset): wrap it in a node.
	if (IsLeaf(nodes[0].nodeData)) {
		SlotBoxes b;
		if (!in.instLeaves) {
			wideOf[0] = (uint32_t)out->tris.size();
			out->tris.resize(out->tris.size() + 1);
			out->ids.resize(out->tris.size());
		}
		AddSlot(in, wideOf, 0, nullptr, &b, out);
		WideNode w;
		QuantizeNodeOrThrow(b, kNullIndex, 0, &w);
		*stackNeed = kWideSlots - 1;        // the unused slots "pass" for a NaN ray (see the sweep at the end)
		if (in.instLeaves)
			*stackNeed += 1 + (*in.leafStackNeed)[nodes[0].bvhLeaf.leafIndex];
		out->wide.push_back(w);
		return wideStart;
	}

	// The reference tests the root's own box first (bvhaccel.cpp:245-255 with currentNode == 0);
	// no parent holds that box, so a one-child entry node carries it.
	// Pass 1: wide index of every inner reference node, in array order; triangle leaves get their TriRecord index,
	// also in array order (the record index is the reference's tie-break order, layout.h).
	uint32_t nWide = 1;
	uint64_t nLeafTotal = 0;
	const size_t triStart = out->tris.size();
	// Big triangle trees (a 50 M-triangle soup: 74 M reference nodes) run both passes on several threads: ranges of the
	// array are independent once every node knows its indices.  Instance trees append to `insts` and stay serial.
	unsigned nThreads = 1;
	if (!in.instLeaves && in.n >= 400000u) {
		nThreads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
		if (const char *env = getenv("LRB_RELAYOUT_THREADS"))
			nThreads = (unsigned)std::max(1, atoi(env));
	}
	// fn(first node, end node, thread number) over nThreads contiguous ranges of the array
	auto inRanges = [&](const std::function<void(uint32_t, uint32_t, unsigned)> &fn) {
		if (nThreads <= 1) {
			fn(0, in.n, 0);
			return;
		}
		std::vector<std::thread> pool;
		std::vector<std::exception_ptr> errors(nThreads);
		const uint64_t per = ((uint64_t)in.n + nThreads - 1) / nThreads;
		for (unsigned t = 0; t < nThreads; ++t) {
			const uint32_t i0 = (uint32_t)std::min<uint64_t>(in.n, per * t), i1 = (uint32_t)std::min<uint64_t>(in.n, per * (t + 1));
			pool.emplace_back([&, t, i0, i1]() {
				try {
					fn(i0, i1, t);
				} catch (...) {
					errors[t] = std::current_exception();
				}
			});
		}
		for (std::thread &th : pool)
			th.join();
		for (unsigned t = 0; t < nThreads; ++t)
			if (errors[t])
				std::rethrow_exception(errors[t]);
	};
	if (nThreads > 1) {
		// counts per range, exclusive prefix over the ranges, indices: the same numbers the serial loops below give
		std::vector<uint64_t> leaves(nThreads, 0), wides(nThreads, 0);
		inRanges([&](const uint32_t i0, const uint32_t i1, const unsigned t) {
			uint64_t nl = 0, nw = 0;
			for (uint32_t i = i0; i < i1; ++i) {
				if (IsLeaf(nodes[i].nodeData)) {
					++nl;
					continue;
				}
				const uint32_t w = WideNodesFor(CountKids(nodes, i));
				wideOf[i] = w;
				nw += w;
			}
			leaves[t] = nl;
			wides[t] = nw;
		});
		std::vector<uint64_t> leafBase(nThreads), wideBase(nThreads);
		uint64_t l = triStart, w = (uint64_t)wideStart + 1;
		for (unsigned t = 0; t < nThreads; ++t) {
			leafBase[t] = l; wideBase[t] = w;
			l += leaves[t]; w += wides[t];
		}
		nLeafTotal = l - triStart;
		if (l >= kMaxRefIndex)
			throw std::runtime_error("too many leaves");
		if (w >= kMaxRefIndex)
			throw std::runtime_error("too many wide nodes");
		nWide = (uint32_t)(w - wideStart);
		inRanges([&](const uint32_t i0, const uint32_t i1, const unsigned t) {
			uint64_t li = leafBase[t], wi = wideBase[t];
			for (uint32_t i = i0; i < i1; ++i) {
				if (IsLeaf(nodes[i].nodeData))
					wideOf[i] = (uint32_t)li++;
				else {
					const uint32_t cnt = wideOf[i];
					wideOf[i] = (uint32_t)wi;
					wi += cnt;
				}
			}
		});
	} else {
		if (!in.instLeaves) {
			uint64_t nTriLeaves = 0;
			for (uint32_t i = 0; i < in.n; ++i)
				if (IsLeaf(nodes[i].nodeData))
					wideOf[i] = (uint32_t)(triStart + nTriLeaves++);
			if (triStart + nTriLeaves >= kMaxRefIndex)
				throw std::runtime_error("too many leaves");
		}
		for (uint32_t i = 0; i < in.n; ++i) {
			if (IsLeaf(nodes[i].nodeData))
				continue;
			const uint32_t end = Skip(nodes[i].nodeData);
			uint32_t nKids = 0;
			for (uint32_t c = i + 1; c < end; c = Skip(nodes[c].nodeData)) {
				if (IsLeaf(nodes[c].nodeData)) ++nLeafTotal;
				++nKids;
			}
			wideOf[i] = wideStart + nWide;
			nWide += std::max<uint32_t>(1u, (nKids + kWideSlots - 1) / kWideSlots);
		}
	}
	if ((uint64_t)wideStart + nWide >= kMaxRefIndex)
		throw std::runtime_error("too many wide nodes");
	if ((in.instLeaves ? out->insts.size() : out->tris.size()) + nLeafTotal >= kMaxRefIndex)
		throw std::runtime_error("too many leaves");
	phase("pass 1 (indices)");
	out->wide.resize((size_t)wideStart + nWide);
	if (in.instLeaves)
		out->insts.reserve(out->insts.size() + nLeafTotal);
	else {
		out->tris.resize(triStart + nLeafTotal);
		out->ids.resize(triStart + nLeafTotal);
	}

	{
		SlotBoxes b;
		b.n = 1;
		b.whole[0] = false;
		for (int a = 0; a < 3; ++a) { b.lo[0][a] = nodes[0].bvhNode.bboxMin[a]; b.hi[0][a] = nodes[0].bvhNode.bboxMax[a]; }
		b.child[0] = wideOf[0];
		SanitizeSlot(&b, 0);
		QuantizeNodeOrThrow(b, kNullIndex, kNodeEntry, &out->wide[wideStart]);
		if (in.instLeaves || !out->twoLevel) {
			for (int a = 0; a < 3; ++a) {
				out->entryBox[a] = nodes[0].bvhNode.bboxMin[a];
				out->entryBox[3 + a] = nodes[0].bvhNode.bboxMax[a];
				if (out->entryBox[a] > out->entryBox[3 + a])    // see SanitizeSlot
					std::swap(out->entryBox[a], out->entryBox[3 + a]);
			}
		}
	}

	phase("allocate");
	// Pass 2: fill, children in reference order.  Every inner node writes its own wide node(s) and its own
	// triangle records (indices fixed in pass 1), so ranges of the array are filled concurrently for big triangle
	// trees (a 50 M-triangle soup: 74 M reference nodes); instance trees append to `insts` and stay serial.
	const TriTreeView triView = TriViewOf(in);
	auto fillRange = [&](const uint32_t i0, const uint32_t i1) {
		std::vector<uint32_t> kids;
		std::vector<std::pair<double, uint32_t> > keyed;
		for (uint32_t i = i0; i < i1; ++i) {
			if (IsLeaf(nodes[i].nodeData))
				continue;
			if (!in.instLeaves) {
				// triangle trees: the body shared with the device re-layout (relayout_shared.h); nodes of an arity
				// above the builders' maximum of 8 (a foreign array) take the general path below
				const int rc = ConvertInnerNodeTri(triView, i, wideOf.data(), out->wide.data(), out->tris.data(), out->ids.data(), nullptr);
				if (rc == kRelayoutOk)
					continue;
				if (rc != kRelayoutTooManyKids)
					ThrowIfRelayoutError(rc);
			}
			kids.clear();
			const uint32_t end = Skip(nodes[i].nodeData);
			for (uint32_t c = i + 1; c < end; c = Skip(nodes[c].nodeData))
				kids.push_back(c);
			// Slot order = ascending box size.  The kernel sorts the children of a node by entry distance with
			// a network that keeps the slot order of equal keys, and for a ray that STARTS inside several child
			// boxes (every bounce ray does, near the root) all those keys equal ray.mint: visiting the smaller
			// box first finds a near hit sooner and culls more of the rest (kitchen, bounce-2 rays: 16.2 -> 15.6
			// node visits and 5.4 -> 4.8 triangle tests per ray).  Order never changes a result.
			if (kids.size() > 1) {
				keyed.clear();
				for (uint32_t c : kids)
					keyed.push_back(std::make_pair(SlotSizeKey(in, c), c));
				std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<double, uint32_t> &a, const std::pair<double, uint32_t> &b) { return a.first < b.first; });
				for (size_t k = 0; k < kids.size(); ++k)
					kids[k] = keyed[k].second;
			}
			const uint32_t nW = std::max<uint32_t>(1u, ((uint32_t)kids.size() + kWideSlots - 1) / kWideSlots);
			for (uint32_t j = 0; j < nW; ++j) {
				const uint32_t first = j * kWideSlots;
				const uint32_t cnt = std::min<uint32_t>(kWideSlots, (uint32_t)kids.size() - std::min<uint32_t>((uint32_t)kids.size(), first));
				SlotBoxes b;
				// the node's own box bounds every child; MBVH root leaves (no box of their own) take it whole
				b.hasOwn = true;
				for (int a = 0; a < 3; ++a) { b.ownLo[a] = nodes[i].bvhNode.bboxMin[a]; b.ownHi[a] = nodes[i].bvhNode.bboxMax[a]; }
				for (uint32_t k = 0; k < cnt; ++k)
					AddSlot(in, wideOf, kids[first + k], &nodes[i], &b, out);
				QuantizeNodeOrThrow(b, (j + 1 < nW) ? (wideOf[i] + j + 1) : kNullIndex, 0, &out->wide[wideOf[i] + j]);
			}
		}
	};
	inRanges([&](const uint32_t i0, const uint32_t i1, unsigned) { fillRange(i0, i1); });

	phase("pass 2 (fill)");
	// Worst-case live stack entries.  Children always have larger wide indices than their parent
	// (depth-first pre-order), so one reverse sweep suffices.  Visiting w pushes every entry but the
	// one it continues with:  D[w] = (slots of w) - 1 + max over entries D[entry];  entering an
	// instance first pushes the sentinel.  ALL FOUR slots count, used or not: for a ray with a NaN
	// origin / direction (or a NaN time, through the motion matrices) every comparison of the slab test is
	// false, so every slot "passes" -- the unused ones too, whose kNullIndex references are pushed and
	// dropped when popped.  Counting only the used slots under-estimated the depth such a ray reaches
	// (binary trees: three pushes per level instead of one), and the kernels' spill buffers and the choice
	// of the non-spilling kernel are sized by this number (found by tools/fuzz_parity.py --lockstep under
	// AddressSanitizer).
	std::vector<uint32_t> D(nWide, 0);
	for (uint32_t r = nWide; r-- > 0;) {
		const WideNode &w = out->wide[wideStart + r];
		const uint32_t nChild = NodeSlots(w);
		uint32_t k = kWideSlots + (w.next != kNullIndex ? 1u : 0u);
		uint32_t below = 0;
		for (uint32_t c = 0; c < nChild; ++c) {
			const uint32_t ref = w.child[c];
			if (ref & kTagInstance)
				below = std::max(below, 1u + out->insts[ref & kRefIndexMask].pad[0]);
			else if (!(ref & kTagTri))
				below = std::max(below, D[ref - wideStart]);
		}
		if (w.next != kNullIndex)
			below = std::max(below, D[w.next - wideStart]);
		D[r] = (k > 0 ? k - 1 : 0) + below;
	}
	*stackNeed = D[0];
	phase("stack bound");
	return wideStart;
}

struct OclXform { float m[16], mInv[16]; };
struct OclDecomposed {
	float Sx, Sy, Sz, Sxy, Sxz, Syz;
	float R[16];
	float Tx, Ty, Tz, Px, Py, Pz, Pw;
	bool Valid;
};
struct OclInterp {
	float startTime, endTime;
	OclXform start, end;
	OclDecomposed startT, endT;
	float startQ[4], endQ[4];   // w, x, y, z (quaternion_types.cl)
	int hasRotation, hasTranslation, hasScale;
	int hasTranslationX, hasTranslationY, hasTranslationZ;
	int hasScaleX, hasScaleY, hasScaleZ;
	int isActive;
};
static_assert(sizeof(OclInterp) == LRB_INTERPOLATED_TRANSFORM_SIZE, "ocl::InterpolatedTransform layout");

}   // namespace

void FillRootOfView(const WideScene &w, SceneView *v) {
	v->nWide = (uint32_t)w.wide.size();
	v->rootWide = w.rootWide;
	v->twoLevel = w.twoLevel ? 1u : 0u;
	v->rootHasBox = 0;
	v->rootChild = 0;
	v->oneBits = 0x3f800000u;
	for (int i = 0; i < 6; ++i) v->rootBox[i] = 0.f;
	if (w.wide.empty() || w.rootWide == kNullIndex)
		return;
	const WideNode &e = w.wide[w.rootWide];
	// the one-child entry node ConvertTree puts in front of a tree whose root is an inner node: its
	// box (the reference's node 0, exact floats) is tested from the kernel parameters
	if (e.flags & kNodeEntry) {
		v->rootHasBox = 1;
		v->rootChild = e.child[0];
		for (int i = 0; i < 6; ++i) v->rootBox[i] = w.entryBox[i];
	}
}

void PackInterp(const void *src, DevInterp *d) {
	OclInterp it;
	memcpy(&it, src, sizeof(it));
	memset(d, 0, sizeof(*d));
	d->startTime = it.startTime;
	d->endTime = it.endTime;
	d->flags = (it.isActive ? kItActive : 0) | (it.hasRotation ? kItRotation : 0) |
			(it.hasTranslation ? kItTranslation : 0) | (it.hasScale ? kItScale : 0) |
			(it.hasTranslationX ? kItTX : 0) | (it.hasTranslationY ? kItTY : 0) | (it.hasTranslationZ ? kItTZ : 0);
	memcpy(d->startM, it.start.m, 64);
	memcpy(d->endM, it.end.m, 64);
	memcpy(d->R, it.startT.R, 64);
	d->sS[0] = it.startT.Sx; d->sS[1] = it.startT.Sy; d->sS[2] = it.startT.Sz;
	d->eS[0] = it.endT.Sx; d->eS[1] = it.endT.Sy; d->eS[2] = it.endT.Sz;
	d->sT[0] = it.startT.Tx; d->sT[1] = it.startT.Ty; d->sT[2] = it.startT.Tz;
	d->eT[0] = it.endT.Tx; d->eT[1] = it.endT.Ty; d->eT[2] = it.endT.Tz;
	memcpy(d->sQ, it.startQ, 16);
	memcpy(d->eQ, it.endQ, 16);
}

void BuildWideBVH(const lrb_bvh_node *nodes, uint32_t nNodes, const float *xyz, uint64_t nVerts,
		const uint32_t *meshVertexOffsets, uint32_t nMeshes, WideScene *out) {
	*out = WideScene();
	out->twoLevel = false;
	out->nRefNodes = nNodes;
	if (nNodes == 0)
		return;
	if (!xyz || !meshVertexOffsets || nMeshes == 0)
		throw std::runtime_error("BVH upload needs vertices and a mesh vertex-offset table");
	TreeInput in;
	memset(&in, 0, sizeof(in));
	in.nodes = nodes; in.n = nNodes; in.instLeaves = false;
	in.xyz = xyz; in.nVerts = nVerts; in.meshOff = meshVertexOffsets; in.nMeshes = nMeshes;
	uint32_t need = 0;
	out->rootWide = ConvertTree(in, out, &need);
	out->stackNeed = need + 1;
}

static void ConvertRoot(const lrb_bvh_node *rootNodes, uint32_t nRootNodes, uint32_t nTransforms,
		uint32_t nMotions, WideScene *out) {
	TreeInput in;
	memset(&in, 0, sizeof(in));
	in.nodes = rootNodes; in.n = nRootNodes; in.instLeaves = true;
	in.leafRootWide = &out->leafRootWide;
	in.leafStackNeed = &out->leafStackNeed;
	in.nTransforms = nTransforms;
	in.nMotions = nMotions;
	in.leafBox = &out->leafBox;
	in.minv = out->minv.empty() ? nullptr : out->minv.data();
	in.motionFirst = &out->motionFirst;
	in.motionLast = &out->motionLast;
	in.interps = &out->interps;
	const size_t before = out->wide.size();
	uint32_t need = 0;
	out->rootWide = ConvertTree(in, out, &need);
	out->nRootWide = (uint32_t)(out->wide.size() - before);
	out->stackNeed = need + 1;
}

void BuildWideMBVH(const lrb_mbvh_desc &d, WideScene *out) {
	*out = WideScene();
	out->twoLevel = true;
	out->nRefNodes = d.n_root_nodes;
	if (d.n_root_nodes == 0)
		return;
	if (d.n_leaves && (!d.leaf_nodes || !d.leaf_n_nodes || !d.leaf_vertices || !d.leaf_n_vertices))
		throw std::runtime_error("MBVH upload: leaf arrays missing");
	if (d.n_transforms && !d.transforms_minv)
		throw std::runtime_error("MBVH upload: transform array missing");
	if (d.n_motion_systems && (!d.motion_systems || !d.interpolated_transforms))
		throw std::runtime_error("MBVH upload: motion arrays missing");

	const uint32_t zeroOff = 0;
	for (uint32_t i = 0; i < d.n_leaves; ++i) {
		TreeInput in;
		memset(&in, 0, sizeof(in));
		in.nodes = d.leaf_nodes[i]; in.n = d.leaf_n_nodes[i]; in.instLeaves = false;
		in.xyz = d.leaf_vertices[i]; in.nVerts = d.leaf_n_vertices[i];
		in.meshOff = &zeroOff; in.nMeshes = 1;
		uint32_t need = 0;
		uint32_t leafRoot = ConvertTree(in, out, &need);
		// ConvertTree puts a one-child entry node in front of a tree whose root is an inner node: it holds
		// the root's own box, which the reference tests first (bvhaccel.cpp:245-255).  For a leaf tree that
		// test has already happened in world space -- the instance's slot in the root tree bounds this very
		// box (InstanceWorldBox) -- so instances enter at the real root node: one node visit less per
		// instance entry (lightinstances: 3.3-3.6 entries per ray).  Boxes only cull; a ray that would have
		// failed the instance-space root box now visits the root node and fails its children's boxes.
		if (leafRoot != kNullIndex && (out->wide[leafRoot].flags & kNodeEntry))
			leafRoot = out->wide[leafRoot].child[0];
		out->leafRootWide.push_back(leafRoot);
		out->leafStackNeed.push_back(need);
		// root box of the leaf tree in instance space: its node 0, or the lone triangle's build box
		float lb[6] = { kInfF, kInfF, kInfF, -kInfF, -kInfF, -kInfF };      // empty tree: never entered anyway
		if (in.n > 0 && !IsLeaf(in.nodes[0].nodeData)) {
			for (int a = 0; a < 3; ++a) { lb[a] = in.nodes[0].bvhNode.bboxMin[a]; lb[3 + a] = in.nodes[0].bvhNode.bboxMax[a]; }
		} else if (in.n > 0) {
			TriRecord tr;
			FillTri(in, 0, &tr);
			TriBuildBox(tr, lb, lb + 3);
		}
		out->leafBox.insert(out->leafBox.end(), lb, lb + 6);
		out->nRefNodes += d.leaf_n_nodes[i];
	}

	out->minv.assign(d.transforms_minv, d.transforms_minv + 16 * (size_t)d.n_transforms);
	for (uint32_t i = 0; i < d.n_motion_systems; ++i) {
		const lrb_motion_system &ms = d.motion_systems[i];
		if (ms.interpolatedTransformFirstIndex > ms.interpolatedTransformLastIndex ||
				ms.interpolatedTransformLastIndex >= d.n_interpolated_transforms)
			throw std::runtime_error("MBVH upload: motion system index range outside the interpolated-transform array");
		out->motionFirst.push_back(ms.interpolatedTransformFirstIndex);
		out->motionLast.push_back(ms.interpolatedTransformLastIndex);
	}
	out->interps.resize(d.n_interpolated_transforms);
	for (uint32_t i = 0; i < d.n_interpolated_transforms; ++i)
		PackInterp((const char *)d.interpolated_transforms + (size_t)i * LRB_INTERPOLATED_TRANSFORM_SIZE, &out->interps[i]);

	ConvertRoot(d.root_nodes, d.n_root_nodes, d.n_transforms, d.n_motion_systems, out);
}

void UpdateWideMBVHRoot(const lrb_bvh_node *rootNodes, uint32_t nRootNodes, const float *minv,
		uint32_t nTransforms, WideScene *scene) {
	if (!scene->twoLevel)
		throw std::runtime_error("Update is only supported by MBVH scenes");
	if (nTransforms != scene->minv.size() / 16)
		throw std::runtime_error("MBVH update: transform count changed");
	if (nRootNodes == 0)
		throw std::runtime_error("MBVH update: empty root tree");
	scene->wide.resize(scene->wide.size() - scene->nRootWide);
	scene->insts.clear();
	if (nTransforms)
		scene->minv.assign(minv, minv + 16 * (size_t)nTransforms);
	ConvertRoot(rootNodes, nRootNodes, nTransforms, (uint32_t)scene->motionFirst.size(), scene);
}

}   // namespace lrb
