// wide_emulation.cpp -- TEST SCAFFOLDING (never shipped, never linked into the product).
// Compiles the product's re-layout (relayout.cpp) and the product's per-ray traversal body
// (traverse.h) for the HOST, with the reference's no-FMA floating-point model, so that the
// layout conversion, stack discipline, tie rule and two-level logic can be checked against the
// oracle on a machine without a GPU (pytest -m "not gpu").  The GPU tests run the same
// traverse.h body through the CUDA kernels.
#include <cstring>
#include <string>
#include <vector>

#include "luxrays_b200.h"
#include "layout.h"
#include "relayout.h"
#include "traverse.h"

using namespace lrb;

namespace {
struct HostStack {
	std::vector<uint32_t> n;
	std::vector<float> t;
	void push(uint32_t a, float b) { n.push_back(a); t.push_back(b); }
	bool room(int) const { return (n.size() & 1) == 0; }    // exercise both push paths
	void pushFast(uint32_t a, float b) { push(a, b); }
	void pushFastIf(bool p, uint32_t a, float b) { if (p) push(a, b); }
	void pop(uint32_t &a, float &b) { a = n.back(); b = t.back(); n.pop_back(); t.pop_back(); }
	bool empty() const { return n.empty(); }
	unsigned long long depth() const { return n.size(); }
};

static std::string g_err;

static SceneView View(const WideScene &w) {
	SceneView v;
	memset(&v, 0, sizeof(v));
	v.nodes = w.wide.data();
	v.tris = w.tris.data();
	v.gates = w.gates.data();
	v.insts = w.insts.data();
	v.minv = w.minv.data();
	v.motionFirst = w.motionFirst.data();
	v.motionLast = w.motionLast.data();
	v.interps = w.interps.data();
	FillRootOfView(w, &v);
	return v;
}

template <bool TWO> static void Run(const WideScene &w, const lrb_ray *rays, lrb_rayhit *hits, uint32_t n, unsigned long long *stats6) {
	const SceneView v = View(w);
	TraceStats st;
	memset(&st, 0, sizeof(st));
	HostStack stk;
	for (uint32_t i = 0; i < n; ++i) {
		if (rays[i].flags & LRB_RAY_FLAGS_MASKED)
			continue;
		st.rays++;
		RayState s;
		stk.n.clear(); stk.t.clear();
		if (InitRay(v, rays[i], s)) {
			while (Step<TWO, true>(v, rays[i], s, stk, &st)) { }
		}
		WriteHit(s, rays[i].maxt, &hits[i]);
	}
	if (stats6) {
		stats6[0] = st.rays; stats6[1] = st.wideNodes; stats6[2] = st.triangles;
		stats6[3] = st.instances; stats6[4] = st.motionSamples; stats6[5] = st.maxStack;
	}
}
}   // namespace

extern "C" {

const char *emu_last_error() { return g_err.c_str(); }

void *emu_bvh_create(const lrb_bvh_node *nodes, uint32_t nNodes, const float *xyz, uint64_t nVerts,
		const uint32_t *offs, uint32_t nMeshes) {
	WideScene *w = new WideScene();
	try {
		BuildWideBVH(nodes, nNodes, xyz, nVerts, offs, nMeshes, w);
	} catch (const std::exception &e) {
		g_err = e.what();
		delete w;
		return nullptr;
	}
	return w;
}

void *emu_mbvh_create(const lrb_mbvh_desc *d) {
	WideScene *w = new WideScene();
	try {
		BuildWideMBVH(*d, w);
	} catch (const std::exception &e) {
		g_err = e.what();
		delete w;
		return nullptr;
	}
	return w;
}

int emu_mbvh_update(void *wp, const lrb_bvh_node *root, uint32_t nRoot, const float *minv, uint32_t nT) {
	try {
		UpdateWideMBVHRoot(root, nRoot, minv, nT, (WideScene *)wp);
	} catch (const std::exception &e) {
		g_err = e.what();
		return 1;
	}
	return 0;
}

void emu_free(void *w) { delete (WideScene *)w; }

// info6: wide nodes, triangles, instances, stackNeed, rootWide, twoLevel
void emu_info(void *wp, uint32_t *info6) {
	const WideScene *w = (const WideScene *)wp;
	info6[0] = (uint32_t)w->wide.size(); info6[1] = (uint32_t)w->tris.size(); info6[2] = (uint32_t)w->insts.size();
	info6[3] = w->stackNeed; info6[4] = w->rootWide; info6[5] = w->twoLevel ? 1 : 0;
}

// raw copies of the re-laid-out arrays (layout checks from Python)
void emu_copy_nodes(void *wp, void *dst) {
	const WideScene *w = (const WideScene *)wp;
	memcpy(dst, w->wide.data(), w->wide.size() * sizeof(WideNode));
}
void emu_copy_tris(void *wp, void *dst) {
	const WideScene *w = (const WideScene *)wp;
	memcpy(dst, w->tris.data(), w->tris.size() * sizeof(TriRecord));
}
void emu_copy_gates(void *wp, void *dst) {
	const WideScene *w = (const WideScene *)wp;
	memcpy(dst, w->gates.data(), w->gates.size() * sizeof(TriGate));
}

void emu_trace(void *wp, const lrb_ray *rays, lrb_rayhit *hits, uint32_t n, unsigned long long *stats6) {
	const WideScene *w = (const WideScene *)wp;
	if (w->twoLevel) Run<true>(*w, rays, hits, n, stats6);
	else Run<false>(*w, rays, hits, n, stats6);
}

int emu_validate_tree(const lrb_bvh_node *nodes, uint32_t n) {
	std::string e;
	if (ValidateTree(nodes, n, &e)) return 0;
	g_err = e;
	return 1;
}

}
