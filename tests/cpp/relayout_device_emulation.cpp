// relayout_device_emulation.cpp -- TEST SCAFFOLDING (never shipped).  Drives the per-record bodies of the DEVICE re-layout
// (luxcore_b200/csrc/relayout_kernels.cuh + relayout_shared.h: the functions the CUDA kernels of lrb_bvh_build_scene call,
// one thread per record) with plain loops on the host, in the order of the device pipeline: build boxes, leaf payload,
// count -> exclusive sum -> index, fill, stack bound.  tests/test_relayout_device_cpu.py holds the result to the bytes of
// the host re-layout (relayout.cpp BuildWideBVH through tests/cpp/wide_emulation.cpp) of the same array.
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "relayout_kernels.cuh"

using namespace lrb;

namespace {
std::string g_err;

struct Result {
	std::vector<WideNode> wide;
	std::vector<TriRecord> tris;
	std::vector<TriIds> ids;
	std::vector<lrb_bvh_node> nodes;    // the array with its leaf payload
	std::vector<float> boxes;           // 6 per triangle
	uint32_t stackNeed;
	float entryBox[6];
};
}

extern "C" {

const char *rde_last_error() { return g_err.c_str(); }

// builderNodes: the array as the builder kernels emit it (leaves carry the global triangle number in triangleLeaf.v[0]).
void *rde_run(const lrb_bvh_node *builderNodes, uint32_t n, const float *xyz, uint64_t nVerts, const uint32_t *meshVertOff, const uint32_t *meshTriOff,
		uint32_t nMeshes, const uint32_t *triIdx) {
	Result *r = new Result();
	const uint32_t nTris = meshTriOff[nMeshes];
	// 0: build boxes
	r->boxes.resize(6 * (size_t)nTris);
	for (uint32_t g = 0; g < nTris; ++g) {
		const int rc = LeafBoxBody(xyz, nVerts, meshVertOff, meshTriOff, nMeshes, triIdx, g, &r->boxes[6 * (size_t)g]);
		if (rc != kRelayoutOk) { g_err = RelayoutErrorString(rc); delete r; return nullptr; }
	}
	// 1: leaf payload
	r->nodes.assign(builderNodes, builderNodes + n);
	for (uint32_t i = 0; i < n; ++i)
		LeafPayloadBody(&r->nodes[i], meshTriOff, nMeshes, triIdx);
	// 2: count, exclusive sum, index
	std::vector<unsigned long long> counts(n), scanned(n);
	for (uint32_t i = 0; i < n; ++i)
		counts[i] = RelayoutCountBody(r->nodes.data(), i);
	unsigned long long run = 0;
	for (uint32_t i = 0; i < n; ++i) { scanned[i] = run; run += counts[i]; }
	std::vector<uint32_t> wideOf(n);
	for (uint32_t i = 0; i < n; ++i)
		wideOf[i] = RelayoutIndexBody(r->nodes.data(), i, scanned[i]);
	const uint32_t nWide = 1u + (uint32_t)(run >> 32);
	if ((uint32_t)(run & 0xffffffffull) != nTris) { g_err = "leaf count differs from the triangle count"; delete r; return nullptr; }
	// 3: fill (the kernel's thread i; thread 0 also writes the entry node)
	r->wide.resize(nWide);
	r->tris.resize(nTris);
	r->ids.resize(nTris);
	std::vector<uint32_t> parentOf(nWide, 0xdeadbeefu);
	TriTreeView tv;
	tv.nodes = r->nodes.data(); tv.n = n; tv.xyz = xyz; tv.nVerts = nVerts; tv.meshOff = meshVertOff; tv.nMeshes = nMeshes;
	for (uint32_t i = 0; i < n; ++i) {
		if (i == 0) {
			const int rc = MakeEntryNode(tv.nodes[0], wideOf[0], &r->wide[0], r->entryBox);
			parentOf[0] = kNullIndex;
			parentOf[wideOf[0]] = 0u;
			if (rc != kRelayoutOk) { g_err = RelayoutErrorString(rc); delete r; return nullptr; }
		}
		if (RlIsLeaf(tv.nodes[i].nodeData))
			continue;
		const int rc = ConvertInnerNodeTri(tv, i, wideOf.data(), r->wide.data(), r->tris.data(), r->ids.data(), parentOf.data());
		if (rc != kRelayoutOk) { g_err = RelayoutErrorString(rc); delete r; return nullptr; }
	}
	// 4: stack bound -- StackNeedKernel's walk, one "thread" after the other (threads in DESCENDING index order, so that
	// the order of arrival differs from the host's sweep)
	std::vector<uint32_t> below(nWide, 0u), arrived(nWide, 0u);
	r->stackNeed = 0xffffffffu;
	for (uint32_t t = nWide; t-- > 0;) {
		if (InnerEntriesOf(r->wide[t]) != 0u)
			continue;
		uint32_t cur = t, D = StackNeedOfNode(r->wide[cur], 0u);
		for (;;) {
			const uint32_t p = parentOf[cur];
			if (p == kNullIndex) { r->stackNeed = D; break; }
			if (p >= nWide) { g_err = "wide node without a parent"; delete r; return nullptr; }
			below[p] = below[p] < D ? D : below[p];
			if (++arrived[p] < InnerEntriesOf(r->wide[p]))
				break;
			cur = p;
			D = StackNeedOfNode(r->wide[cur], below[cur]);
		}
	}
	return r;
}

void rde_free(void *h) { delete static_cast<Result *>(h); }
void rde_info(void *h, uint32_t *out) {     // wide, tris, stack need (+ 1, as BuildWideBVH reports it)
	Result *r = static_cast<Result *>(h);
	out[0] = (uint32_t)r->wide.size(); out[1] = (uint32_t)r->tris.size(); out[2] = r->stackNeed + 1u;
}
void rde_copy_nodes(void *h, void *dst) { Result *r = static_cast<Result *>(h); memcpy(dst, r->wide.data(), r->wide.size() * sizeof(WideNode)); }
void rde_copy_tris(void *h, void *dst) { Result *r = static_cast<Result *>(h); memcpy(dst, r->tris.data(), r->tris.size() * sizeof(TriRecord)); }
void rde_copy_ids(void *h, void *dst) { Result *r = static_cast<Result *>(h); memcpy(dst, r->ids.data(), r->ids.size() * sizeof(TriIds)); }
void rde_copy_ref_nodes(void *h, void *dst) { Result *r = static_cast<Result *>(h); memcpy(dst, r->nodes.data(), r->nodes.size() * sizeof(lrb_bvh_node)); }
void rde_copy_boxes(void *h, void *dst) { Result *r = static_cast<Result *>(h); memcpy(dst, r->boxes.data(), r->boxes.size() * sizeof(float)); }
void rde_copy_entry_box(void *h, void *dst) { memcpy(dst, static_cast<Result *>(h)->entryBox, 24); }

}
