// relayout.cpp -- see relayout.h.  Converts the reference's depth-first skip-list arrays
// (layout rules: src/luxrays/core/bvh/bvhclassicbuild.cpp:181-220 -- a node's first child is
// index+1, siblings are reached through the skip index, a leaf's skip index is index+1) into
// wide nodes.  The host side replaces the per-node index rewriting the reference does before
// upload (bvhaccelhw.cpp:126-145, mbvhaccelhw.cpp:380-427).

#include "relayout.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>

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
};

static void FillTri(const TreeInput &in, uint32_t c, TriRecord *tr) {
	const lrb_bvh_node &nd = in.nodes[c];
	const uint32_t mesh = nd.triangleLeaf.meshIndex;
	if (mesh >= in.nMeshes)
		throw std::runtime_error("triangle leaf references a mesh outside the vertex-offset table");
	const float *p[3];
	for (int j = 0; j < 3; ++j) {
		const uint64_t g = (uint64_t)nd.triangleLeaf.v[j] + in.meshOff[mesh];
		if (g >= in.nVerts)
			throw std::runtime_error("triangle leaf references a vertex outside the vertex buffer");
		p[j] = in.xyz + 3 * g;
	}
	for (int k = 0; k < 3; ++k) {
		tr->p0[k] = p[0][k];
		tr->p1[k] = p[1][k];
		tr->p2[k] = p[2][k];
	}
	tr->meshIndex = mesh;
	tr->triangleIndex = nd.triangleLeaf.triangleIndex;
	tr->order = c;
	tr->pad[0] = tr->pad[1] = tr->pad[2] = tr->pad[3] = 0;
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

// MachineEpsilon::E (include/luxrays/core/epsilon.h:48-86) with the default clamp 1e-5 .. 1e-1.
static inline float EpsOf(float v) {
	union { float f; uint32_t i; } mf;
	mf.f = v;
	mf.i += 0x80u;
	const float e = fabsf(mf.f - v);
	return e > 1e-5f ? (e < 1e-1f ? e : 1e-1f) : 1e-5f;
}

static const float kInfF = std::numeric_limits<float>::infinity();

static void ClearNode(WideNode *w) {
	memset(w, 0, sizeof(*w));
	for (uint32_t k = 0; k < kWideSlots; ++k) {
		w->lox[k] = w->loy[k] = w->loz[k] = kInfF;      // the empty box: no ray passes it
		w->hix[k] = w->hiy[k] = w->hiz[k] = -kInfF;
		w->child[k] = kNullIndex;
	}
	w->next = kNullIndex;
}

static void SetSlotBox(WideNode *w, uint32_t k, const float lo[3], const float hi[3]) {
	w->lox[k] = lo[0]; w->loy[k] = lo[1]; w->loz[k] = lo[2];
	w->hix[k] = hi[0]; w->hiy[k] = hi[1]; w->hiz[k] = hi[2];
}

// The box BVHAccel::Init gives the builders for one triangle (bvhaccel.cpp:116-122): bounds of the
// three vertices, grown by MachineEpsilon::E of the bounds.
static void TriBuildBox(const TriRecord &tr, float lo[3], float hi[3]) {
	float e = 0.f;
	for (int k = 0; k < 3; ++k) {
		lo[k] = std::min(std::min(tr.p0[k], tr.p1[k]), tr.p2[k]);
		hi[k] = std::max(std::max(tr.p0[k], tr.p1[k]), tr.p2[k]);
		e = std::max(e, std::max(EpsOf(lo[k]), EpsOf(hi[k])));
	}
	for (int k = 0; k < 3; ++k) {
		lo[k] -= e;
		hi[k] += e;
	}
}

// Fills slot k of `w` with reference child `c` (inner node, triangle leaf or MBVH root leaf).
static void FillSlot(const TreeInput &in, const std::vector<uint32_t> &wideOf, uint32_t c, WideNode *w, uint32_t k, WideScene *out) {
	const lrb_bvh_node &ch = in.nodes[c];
	if (!IsLeaf(ch.nodeData)) {
		SetSlotBox(w, k, ch.bvhNode.bboxMin, ch.bvhNode.bboxMax);
		w->child[k] = wideOf[c];
	} else if (in.instLeaves) {
		const float lo[3] = { -kInfF, -kInfF, -kInfF }, hi[3] = { kInfF, kInfF, kInfF };
		SetSlotBox(w, k, lo, hi);
		InstRecord ir;
		FillInst(in, c, &ir);
		w->child[k] = kTagInstance | (uint32_t)out->insts.size();
		out->insts.push_back(ir);
	} else {
		TriRecord tr;
		FillTri(in, c, &tr);
		float lo[3], hi[3];
		TriBuildBox(tr, lo, hi);
		SetSlotBox(w, k, lo, hi);
		w->child[k] = kTagTri | (uint32_t)out->tris.size();
		out->tris.push_back(tr);
	}
	w->nChild = k + 1;
}

// Appends the wide form of one reference tree to `out`.  Returns the wide index of its root
// (kNullIndex for an empty tree) and the tree's worst-case stack need.
static uint32_t ConvertTree(const TreeInput &in, WideScene *out, uint32_t *stackNeed) {
	*stackNeed = 0;
	if (in.n == 0)
		return kNullIndex;
	std::string err;
	if (!ValidateTree(in.nodes, in.n, &err))
		throw std::runtime_error("malformed BVHArrayNode array: " + err);

	const uint32_t wideStart = (uint32_t)out->wide.size();
	const lrb_bvh_node *nodes = in.nodes;
	std::vector<uint32_t> wideOf(in.n, kNullIndex);

	// The root itself is a leaf (one-triangle mesh / one-mesh dataset): wrap it in a node.
	if (IsLeaf(nodes[0].nodeData)) {
		WideNode w;
		ClearNode(&w);
		FillSlot(in, wideOf, 0, &w, 0, out);
		if (in.instLeaves)
			*stackNeed = 1 + (*in.leafStackNeed)[nodes[0].bvhLeaf.leafIndex];
		out->wide.push_back(w);
		return wideStart;
	}

	// The reference tests the root's own box first (bvhaccel.cpp:245-255 with currentNode == 0);
	// no parent holds that box, so a one-child entry node carries it.
	// Pass 1: wide index of every inner reference node, in array order.
	uint32_t nWide = 1;
	uint64_t nLeafTotal = 0;
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
	if ((uint64_t)wideStart + nWide >= kMaxRefIndex)
		throw std::runtime_error("too many wide nodes");
	if ((in.instLeaves ? out->insts.size() : out->tris.size()) + nLeafTotal >= kMaxRefIndex)
		throw std::runtime_error("too many leaves");
	out->wide.resize((size_t)wideStart + nWide);
	if (in.instLeaves)
		out->insts.reserve(out->insts.size() + nLeafTotal);
	else
		out->tris.reserve(out->tris.size() + nLeafTotal);

	{
		WideNode &e = out->wide[wideStart];
		ClearNode(&e);
		SetSlotBox(&e, 0, nodes[0].bvhNode.bboxMin, nodes[0].bvhNode.bboxMax);
		e.child[0] = wideOf[0];
		e.nChild = 1;
		e.flags = kNodeEntry;
	}

	// Pass 2: fill, children in reference order.
	std::vector<uint32_t> kids;
	for (uint32_t i = 0; i < in.n; ++i) {
		if (IsLeaf(nodes[i].nodeData))
			continue;
		kids.clear();
		const uint32_t end = Skip(nodes[i].nodeData);
		for (uint32_t c = i + 1; c < end; c = Skip(nodes[c].nodeData))
			kids.push_back(c);
		const uint32_t nW = std::max<uint32_t>(1u, ((uint32_t)kids.size() + kWideSlots - 1) / kWideSlots);
		for (uint32_t j = 0; j < nW; ++j) {
			WideNode &w = out->wide[wideOf[i] + j];
			ClearNode(&w);
			const uint32_t first = j * kWideSlots;
			const uint32_t cnt = std::min<uint32_t>(kWideSlots, (uint32_t)kids.size() - std::min<uint32_t>((uint32_t)kids.size(), first));
			for (uint32_t k = 0; k < cnt; ++k)
				FillSlot(in, wideOf, kids[first + k], &w, k, out);
			w.next = (j + 1 < nW) ? (wideOf[i] + j + 1) : kNullIndex;
		}
	}

	// Worst-case live stack entries.  Children always have larger wide indices than their parent
	// (depth-first pre-order), so one reverse sweep suffices.  Visiting w pushes every entry but the
	// one it continues with:  D[w] = (entries of w) - 1 + max over entries D[entry];  entering an
	// instance first pushes the sentinel.
	std::vector<uint32_t> D(nWide, 0);
	for (uint32_t r = nWide; r-- > 0;) {
		const WideNode &w = out->wide[wideStart + r];
		uint32_t k = w.nChild + (w.next != kNullIndex ? 1u : 0u);
		uint32_t below = 0;
		for (uint32_t c = 0; c < w.nChild; ++c) {
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
	for (int i = 0; i < 6; ++i) v->rootBox[i] = 0.f;
	if (w.wide.empty() || w.rootWide == kNullIndex)
		return;
	const WideNode &e = w.wide[w.rootWide];
	// the one-child entry node ConvertTree puts in front of a tree whose root is an inner node
	if (e.flags & kNodeEntry) {
		v->rootHasBox = 1;
		v->rootChild = e.child[0];
		v->rootBox[0] = e.lox[0]; v->rootBox[1] = e.loy[0]; v->rootBox[2] = e.loz[0];
		v->rootBox[3] = e.hix[0]; v->rootBox[4] = e.hiy[0]; v->rootBox[5] = e.hiz[0];
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
		out->leafRootWide.push_back(ConvertTree(in, out, &need));
		out->leafStackNeed.push_back(need);
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
