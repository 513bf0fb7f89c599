// wide_emulation.cpp -- TEST SCAFFOLDING (never shipped, never linked into the product).
// Compiles the product's re-layout (relayout.cpp) and the product's per-ray traversal body
// (traverse.h) for the HOST, with the reference's no-FMA floating-point model, so that the
// layout conversion, stack discipline, tie rule and two-level logic can be checked against the
// oracle on a machine without a GPU (pytest -m "not gpu").  The GPU tests run the same
// traverse.h body through the CUDA kernels.
#include <algorithm>
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
	bool slow() const { return false; }
	void popFast(uint32_t &a, float &b) {
		if (n.empty()) { a = kStackBottom; b = -LRB_INF; }
		else pop(a, b);
	}
	void keepBottom() { }
	bool peekIf(bool want, uint32_t &a, float &b) const {
		if (!want || n.empty()) return false;
		a = n.back(); b = t.back();
		return true;
	}
	void dropIf(bool p) { if (p) { n.pop_back(); t.pop_back(); } }
	bool empty() const { return n.empty(); }
	unsigned long long depth() const { return n.size(); }
	float w[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
	void stashRay(float ox, float oy, float oz, float dx, float dy, float dz, float ix, float iy, float iz) {
		w[0] = ox; w[1] = oy; w[2] = oz; w[3] = dx; w[4] = dy; w[5] = dz; w[6] = ix; w[7] = iy; w[8] = iz;
	}
	void loadRay(float &ox, float &oy, float &oz, float &dx, float &dy, float &dz, float &ix, float &iy, float &iz) const {
		ox = w[0]; oy = w[1]; oz = w[2]; dx = w[3]; dy = w[4]; dz = w[5]; ix = w[6]; iy = w[7]; iz = w[8];
	}
};

static std::string g_err;

static SceneView View(const WideScene &w) {
	SceneView v;
	memset(&v, 0, sizeof(v));
	v.nodes = w.wide.data();
	v.tris = w.tris.data();
	v.ids = w.ids.data();
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
			while (Step<TWO, true>(v, s, stk, &st)) { }
		}
		WriteHit(v, s, rays[i].maxt, &hits[i]);
	}
	if (stats6) {
		stats6[0] = st.rays; stats6[1] = st.wideNodes; stats6[2] = st.triangles;
		stats6[3] = st.instances; stats6[4] = st.motionSamples; stats6[5] = st.maxStack;
	}
}

// ---- scheduling model of the persistent kernel ------------------------------------------------
// Replays TracePersistent's control flow (trace_kernels.cuh) for `nWarps` warps that share the ray
// counter, one outer-loop iteration per warp in turn: bulk re-fill when fewer than refillBelow lanes
// are alive, per-iteration Resolve, node / triangle phase vote with triBias.  Counts what the warp
// ISSUES (phases, pop-loop trips = the longest lane's), which is what an issue-bound kernel pays
// for; used to compare trees and policies without a GPU (tools/warp_model.py).
struct SimLane {
	RayState s;
	HostStack stk;
	uint32_t rayIdx;
	int state;      // 0 idle, 1 active, 2 unsaved
};
struct SimWarp {
	SimLane lane[32];
	bool exhausted;
	SimWarp() : exhausted(false) { for (int i = 0; i < 32; ++i) lane[i].state = 0; }
};

template <bool TWO> static void WarpSim(const WideScene &w, const lrb_ray *rays, lrb_rayhit *hits, uint32_t n,
		uint32_t nWarps, uint32_t refillBelow, uint32_t triBias, unsigned long long *out16) {
	const uint32_t instBias = (triBias >> 16) & 0x7fffu;      // packed by the Python wrapper: low 16 bits tri_bias, bits 16-30 inst_bias,
	const bool specPop = !(triBias >> 31);                    // bit 31: WITHOUT the speculative pop at the end of the phases
	triBias &= 0xffffu;
	// experimental policy (refillBelow bits 8-15 = triMin > 0): every iteration runs the node phase for all lanes that hold a
	// node and THEN the triangle phase for all lanes that hold a triangle (including those that just got one), the latter
	// only when at least triMin lanes are ready or no node phase ran
	const uint32_t triMin = (refillBelow >> 8) & 0xffu;
	refillBelow &= 0xffu;
	const SceneView v = View(w);
	std::vector<SimWarp> warps(nWarps);
	uint32_t counter = 0;
	unsigned long long outer = 0, inner = 0, nodePhases = 0, triPhases = 0, nodeLanes = 0, triLanes = 0,
			popTrips = 0, popLanes = 0, gatePhases = 0, gateLanes = 0, storePhases = 0, refills = 0, traced = 0, idlePhaseLanes = 0, slowPhases = 0, slowLanes = 0, instTrips = 0, instLanes = 0, leaveTrips = 0, leaveLanes = 0;
	const size_t smemDepth = 16;
	uint32_t live = nWarps;
	std::vector<char> done(nWarps, 0);
	while (live) {
		for (uint32_t wi = 0; wi < nWarps; ++wi) {
			if (done[wi]) continue;
			SimWarp &W = warps[wi];
			++outer;
			bool anyStore = false;
			for (int l = 0; l < 32; ++l) {
				SimLane &L = W.lane[l];
				if (L.state == 2) {
					WriteHit(v, L.s, rays[L.rayIdx].maxt, &hits[L.rayIdx]);
					L.state = 0;
					anyStore = true;
				}
			}
			if (anyStore) ++storePhases;
			int nIdle = 0;
			for (int l = 0; l < 32; ++l) nIdle += W.lane[l].state == 0;
			if (!W.exhausted && nIdle) {
				++refills;
				const uint32_t base = counter;
				counter += (uint32_t)nIdle;
				uint32_t k = 0;
				for (int l = 0; l < 32; ++l) {
					SimLane &L = W.lane[l];
					if (L.state != 0) continue;
					const uint32_t slot = base + k++;
					if (slot >= n) continue;
					if (rays[slot].flags & LRB_RAY_FLAGS_MASKED) continue;
					L.rayIdx = slot;
					++traced;
					L.stk.n.clear(); L.stk.t.clear();
					if (InitRay(v, rays[slot], L.s)) L.state = 1;
					else WriteHit(v, L.s, rays[slot].maxt, &hits[slot]);
				}
				if (base + (uint32_t)nIdle >= n) W.exhausted = true;
			}
			int nActive = 0;
			for (int l = 0; l < 32; ++l) nActive += W.lane[l].state == 1;
			if (nActive == 0) {
				if (W.exhausted) { done[wi] = 1; --live; }
				continue;
			}
			const int floorLanes = W.exhausted ? 1 : (int)refillBelow;
			int nLive;
			do {
				++inner;
				unsigned long long maxTrips = 0, maxInst = 0;
				bool anyLeave = false;
				for (int l = 0; l < 32; ++l) {
					SimLane &L = W.lane[l];
					if (L.state == 1 && NeedsResolve<TWO>(L.s.cur)) {
						const size_t before = L.stk.n.size();
						const bool wasInside = L.s.inInstance;
						if (!Resolve<TWO, false>(v, L.s, L.stk, nullptr))
							L.state = 2;
						if (TWO && wasInside && !L.s.inInstance) { ++leaveLanes; anyLeave = true; }
						// pops performed (an entry into an instance pushes the sentinel: count at least one trip)
						const size_t after = L.stk.n.size();
						unsigned long long trips = before > after ? (unsigned long long)(before - after) : 1ull;
						if (L.state == 2) trips += 1;
						popLanes += trips;
						if (trips > maxTrips) maxTrips = trips;
					}
				}
				popTrips += maxTrips;
				if (anyLeave) ++leaveTrips;
				// ---- from here on: the same statements as TracePersistent's loop body (trace_kernels.cuh), with
				// ballots replaced by loops over the lanes; the votes are the kernels' own functions (traverse.h)
				LaneWork work[32];
				int nTri = 0, nNode = 0, nInst = 0;
				for (int l = 0; l < 32; ++l) {
					work[l] = W.lane[l].state == 1 ? WorkOf<TWO>(W.lane[l].s.cur) : kWorkNone;
					nTri += work[l] == kWorkTri;
					nNode += work[l] == kWorkNode;
					if (TWO) nInst += work[l] == kWorkInstance;
				}
				if (TWO) {
					if (VoteEnterInstances(nInst, nNode, nTri, instBias, triBias)) {
						for (int l = 0; l < 32; ++l) {
							SimLane &L = W.lane[l];
							if (work[l] == kWorkInstance) {
								EnterInstance<false>(v, L.s, L.stk, nullptr);
								work[l] = WorkOf<TWO>(L.s.cur);
								++instLanes;
								maxInst = 1;
							}
						}
						nNode = 0;
						for (int l = 0; l < 32; ++l) nNode += work[l] == kWorkNode;
						nInst = 0;
					}
				}
				instTrips += maxInst;
				if (triMin) {
					bool ranNode = false;
					if (nNode) {
						ranNode = true;
						++nodePhases;
						nodeLanes += (unsigned long long)nNode;
						for (int l = 0; l < 32; ++l) {
							SimLane &L = W.lane[l];
							if (work[l] == kWorkNode) {
								NodeStep<TWO, false>(v, L.s, L.stk, nullptr);
								work[l] = (L.s.cur != kNullIndex && L.s.cur < kTagInstance && (L.s.cur & kTagTri)) ? kWorkTri : kWorkNone;
							}
						}
					}
					int nT = 0;
					for (int l = 0; l < 32; ++l) nT += work[l] == kWorkTri;
					if (nT && ((uint32_t)nT >= triMin || !ranNode)) {
						++triPhases;
						triLanes += (unsigned long long)nT;
						int accepted = 0;
						for (int l = 0; l < 32; ++l) {
							SimLane &L = W.lane[l];
							if (work[l] == kWorkTri) {
								const float before = L.s.maxt;
								const uint32_t bi = L.s.bestInst, hm = L.s.hitRef;
								TriStep<TWO, false>(v, L.s, nullptr);
								if (L.s.maxt != before || L.s.bestInst != bi || L.s.hitRef != hm) ++accepted;
							}
						}
						if (accepted) { ++gatePhases; gateLanes += (unsigned long long)accepted; }
					}
					// live lanes for the loop condition
					nTri = 0; nNode = 0;
					for (int l = 0; l < 32; ++l) nNode += W.lane[l].state == 1;
				} else
				if (VoteTrianglePhase(nTri, nNode, triBias)) {
					if (nTri) {
						++triPhases;
						triLanes += (unsigned long long)nTri;
						idlePhaseLanes += (unsigned long long)nNode;
						int accepted = 0;
						for (int l = 0; l < 32; ++l) {
							SimLane &L = W.lane[l];
							if (work[l] == kWorkTri) {
								const float before = L.s.maxt;
								const uint32_t bi = L.s.bestInst, hm = L.s.hitRef;
								TriStep<TWO, false>(v, L.s, nullptr);
								if (L.s.maxt != before || L.s.bestInst != bi || L.s.hitRef != hm) ++accepted;
								if (specPop) PopSpec<TWO>(L.s, L.stk);
							}
						}
						if (accepted) { ++gatePhases; gateLanes += (unsigned long long)accepted; }
					}
				} else {
					++nodePhases;
					nodeLanes += (unsigned long long)nNode;
					{
						int slow = 0;
						for (int l = 0; l < 32; ++l)
							if (work[l] == kWorkNode && W.lane[l].stk.n.size() + 4 > smemDepth) ++slow;
						if (slow) { ++slowPhases; slowLanes += (unsigned long long)slow; }
					}
					idlePhaseLanes += (unsigned long long)nTri;
					for (int l = 0; l < 32; ++l) {
						SimLane &L = W.lane[l];
						if (work[l] == kWorkNode) {
							NodeStep<TWO, false>(v, L.s, L.stk, nullptr);
							if (specPop) PopSpec<TWO>(L.s, L.stk);
						}
					}
				}
				nLive = nTri + nNode + nInst;
			} while (nLive >= floorLanes);
		}
	}
	out16[0] = traced; out16[1] = outer; out16[2] = inner; out16[3] = nodePhases; out16[4] = triPhases;
	out16[5] = nodeLanes; out16[6] = triLanes; out16[7] = popTrips; out16[8] = popLanes; out16[9] = gatePhases;
	out16[10] = gateLanes; out16[11] = storePhases; out16[12] = refills; out16[13] = idlePhaseLanes; out16[14] = slowPhases; out16[15] = slowLanes; out16[16] = instTrips; out16[17] = instLanes; out16[18] = leaveTrips; out16[19] = leaveLanes;
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
void emu_copy_ids(void *wp, void *dst) {
	const WideScene *w = (const WideScene *)wp;
	memcpy(dst, w->ids.data(), w->ids.size() * sizeof(TriIds));
}

void emu_trace(void *wp, const lrb_ray *rays, lrb_rayhit *hits, uint32_t n, unsigned long long *stats6) {
	const WideScene *w = (const WideScene *)wp;
	if (w->twoLevel) Run<true>(*w, rays, hits, n, stats6);
	else Run<false>(*w, rays, hits, n, stats6);
}

void emu_warp_sim(void *wp, const lrb_ray *rays, lrb_rayhit *hits, uint32_t n, uint32_t nWarps, uint32_t refillBelow,
		uint32_t triBias, unsigned long long *out16) {
	const WideScene *w = (const WideScene *)wp;
	if (w->twoLevel) WarpSim<true>(*w, rays, hits, n, nWarps, refillBelow, triBias, out16);
	else WarpSim<false>(*w, rays, hits, n, nWarps, refillBelow, triBias, out16);
}

// The SceneView of the re-laid-out scene (plain struct of host pointers into the handle's arrays), for
// the lockstep harness of the real kernel source (kernel_lockstep.cpp, a separate library).
uint32_t emu_scene_view_size() { return (uint32_t)sizeof(SceneView); }
void emu_scene_view(void *wp, void *out) {
	const SceneView v = View(*(const WideScene *)wp);
	memcpy(out, &v, sizeof(v));
}

int emu_validate_tree(const lrb_bvh_node *nodes, uint32_t n) {
	std::string e;
	if (ValidateTree(nodes, n, &e)) return 0;
	g_err = e;
	return 1;
}

}
