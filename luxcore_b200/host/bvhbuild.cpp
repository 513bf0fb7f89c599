// bvhbuild.cpp -- host-side BVH builders emitting the reference's BVHArrayNode skip-list array
// (array rules: src/luxrays/core/bvh/bvhclassicbuild.cpp:181-220 -- depth-first pre-order, first
// child at index+1, skip index = first node after the subtree, leaf skip = index+1 with bit 31 set).
//
//   BuildBVH (CLASSIC)       : restatement of bvhclassicbuild.cpp:51-233.  Emits the array directly
//                              during the recursion instead of building a pointer tree and
//                              flattening it; partitions, split values and box unions are the same
//                              operations in the same order, so the array is bit-identical.
//   BuildEmbreeBVHMorton     : GPU linear-BVH builder behind the C ABI (lrb_build_lbvh), see below.
//   BuildEmbreeBVHBinnedSAH  : the reference delegates these to Intel Embree 3.12.2 (not part of
//                              /root/reference).  Replaced by a from-scratch binned-SAH k-ary
//                              builder (16 bins, 3 axes, largest-area child split first, one
//                              primitive per leaf, at most treeType children per node), multi-
//                              threaded over sub-trees.  Closest-hit results do not depend on the
//                              topology.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <future>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>

#include "luxrays/core/bvh/bvhbuild.h"
#include "luxrays_b200.h"

namespace luxrays {

typedef ocl::BVHArrayNode Node;
typedef std::vector<BVHTreeNode *> LeafList;

static inline float Centroid2(const BVHTreeNode *n, const u_int axis) { return n->bbox.pMax[axis] + n->bbox.pMin[axis]; }

static void EmitLeaf(const std::deque<const Mesh *> *meshes, const BVHTreeNode *leaf, std::vector<Node> &out) {
	Node an;
	memset(&an, 0, sizeof(an));
	if (meshes) {
		const Triangle &tri = (*meshes)[leaf->triangleLeaf.meshIndex]->GetTriangles()[leaf->triangleLeaf.triangleIndex];
		an.triangleLeaf.v[0] = tri.v[0];
		an.triangleLeaf.v[1] = tri.v[1];
		an.triangleLeaf.v[2] = tri.v[2];
		an.triangleLeaf.meshIndex = leaf->triangleLeaf.meshIndex;
		an.triangleLeaf.triangleIndex = leaf->triangleLeaf.triangleIndex;
	} else {
		an.bvhLeaf.leafIndex = leaf->bvhLeaf.leafIndex;
		an.bvhLeaf.transformIndex = leaf->bvhLeaf.transformIndex;
		an.bvhLeaf.motionIndex = leaf->bvhLeaf.motionIndex;
		an.bvhLeaf.meshOffsetIndex = leaf->bvhLeaf.meshOffsetIndex;
	}
	an.nodeData = ((u_int)out.size() + 1) | 0x80000000u;
	out.push_back(an);
}

static inline void StoreBox(Node &n, const BBox &b) {
	n.bvhNode.bboxMin[0] = b.pMin.x; n.bvhNode.bboxMin[1] = b.pMin.y; n.bvhNode.bboxMin[2] = b.pMin.z;
	n.bvhNode.bboxMax[0] = b.pMax.x; n.bvhNode.bboxMax[1] = b.pMax.y; n.bvhNode.bboxMax[2] = b.pMax.z;
}

static Node *ToArray(const std::vector<Node> &v, u_int *nNodes) {
	*nNodes = (u_int)v.size();
	Node *arr = new Node[v.size()];
	if (!v.empty())
		memcpy(arr, v.data(), v.size() * sizeof(Node));
	return arr;
}

//------------------------------------------------------------------------------
// CLASSIC
//------------------------------------------------------------------------------

namespace {

struct ClassicBuilder {
	const BVHParams &params;
	const std::deque<const Mesh *> *meshes;
	LeafList &list;
	std::vector<Node> out;

	ClassicBuilder(const BVHParams &p, const std::deque<const Mesh *> *m, LeafList &l) : params(p), meshes(m), list(l) { }

	// Smits-style split choice: axis of largest centroid variance, split at the mean (or at the
	// cheapest of costSamples SAH probes).  `splitValue` is only written when a value is chosen,
	// like the reference's out-parameter.
	void ChooseSplit(const u_int begin, const u_int end, float *splitValue, u_int *axis) const {
		if (end - begin == 2) {
			*splitValue = (list[begin]->bbox.pMax[0] + list[begin]->bbox.pMin[0] +
					list[end - 1]->bbox.pMax[0] + list[end - 1]->bbox.pMin[0]) / 2;
			*axis = 0;
			return;
		}
		Point mean2(0, 0, 0), var(0, 0, 0);
		for (u_int i = begin; i < end; i++)
			mean2 += list[i]->bbox.pMax + list[i]->bbox.pMin;
		mean2 /= static_cast<float>(end - begin);
		for (u_int i = begin; i < end; i++) {
			Vector v = list[i]->bbox.pMax + list[i]->bbox.pMin - mean2;
			v.x *= v.x;
			v.y *= v.y;
			v.z *= v.z;
			var += v;
		}
		if (var.x > var.y && var.x > var.z) *axis = 0;
		else if (var.y > var.z) *axis = 1;
		else *axis = 2;

		if (params.costSamples > 1) {
			BBox bounds;
			for (u_int i = begin; i < end; i++)
				bounds = Union(bounds, list[i]->bbox);
			const Vector d = bounds.pMax - bounds.pMin;
			const float invTotalSA = 1.f / bounds.SurfaceArea();
			const float increment = 2 * d[*axis] / (params.costSamples + 1);
			float bestCost = INFINITY;
			for (float probe = 2 * bounds.pMin[*axis] + increment; probe < 2 * bounds.pMax[*axis]; probe += increment) {
				int nBelow = 0, nAbove = 0;
				BBox bbBelow, bbAbove;
				for (u_int j = begin; j < end; j++) {
					if (Centroid2(list[j], *axis) < probe) {
						nBelow++;
						bbBelow = Union(bbBelow, list[j]->bbox);
					} else {
						nAbove++;
						bbAbove = Union(bbAbove, list[j]->bbox);
					}
				}
				const float pBelow = bbBelow.SurfaceArea() * invTotalSA;
				const float pAbove = bbAbove.SurfaceArea() * invTotalSA;
				const float eb = (nAbove == 0 || nBelow == 0) ? params.emptyBonus : 0.f;
				const float cost = params.traversalCost + params.isectCost * (1.f - eb) * (pBelow * nBelow + pAbove * nAbove);
				if (cost < bestCost) {
					bestCost = cost;
					*splitValue = probe;
				}
			}
		} else
			*splitValue = mean2[*axis];
	}

	// Emits the subtree over list[begin, end) and returns its bounding box.
	BBox Emit(const u_int begin, const u_int end) {
		if (end - begin == 1) {
			EmitLeaf(meshes, list[begin], out);
			return list[begin]->bbox;
		}
		const size_t self = out.size();
		out.push_back(Node());

		// log2(treeType) rounds; each round splits every range of the previous round that still
		// holds two or more primitives (bvhclassicbuild.cpp:138-158 walks the same ranges in place)
		std::vector<u_int> cuts;
		cuts.push_back(begin);
		cuts.push_back(end);
		float splitValue = 0.f;
		u_int splitAxis = 0;
		for (u_int fan = 2; fan <= params.treeType; fan *= 2) {
			std::vector<u_int> next;
			next.push_back(cuts[0]);
			for (size_t r = 0; r + 1 < cuts.size(); ++r) {
				const u_int b = cuts[r], e = cuts[r + 1];
				if (e - b >= 2) {
					ChooseSplit(b, e, &splitValue, &splitAxis);
					const u_int ax = splitAxis;
					const float sv = splitValue;
					LeafList::iterator mid = std::partition(list.begin() + b, list.begin() + e,
							[ax, sv](BVHTreeNode *n) { return n->bbox.pMax[ax] + n->bbox.pMin[ax] < sv; });
					u_int middle = (u_int)(mid - list.begin());
					middle = Max(b + 1, Min(e - 1, middle));     // coincident boxes are still split
					next.push_back(middle);
				}
				next.push_back(e);
			}
			cuts.swap(next);
		}

		BBox bbox = Emit(cuts[0], cuts[1]);
		for (size_t r = 1; r + 1 < cuts.size(); ++r)
			bbox = Union(bbox, Emit(cuts[r], cuts[r + 1]));

		Node &n = out[self];
		memset(&n, 0, sizeof(n));
		StoreBox(n, bbox);
		n.nodeData = (u_int)out.size();
		return bbox;
	}
};

}   // namespace

Node *BuildBVH(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes, LeafList &leafList) {
	ClassicBuilder b(params, meshes, leafList);
	b.out.reserve(leafList.size() + leafList.size() / 2 + 1);
	if (!leafList.empty())
		b.Emit(0, (u_int)leafList.size());
	return ToArray(b.out, nNodes);
}

//------------------------------------------------------------------------------
// Binned SAH (stands in for the Embree builders)
//------------------------------------------------------------------------------

namespace {

struct SAHBuilder {
	static const int kBins = 16;
	static const u_int kParallelMin = 1u << 16;     // primitives below which a subtree is built inline

	const BVHParams &params;
	const std::deque<const Mesh *> *meshes;
	LeafList &list;
	std::atomic<int> tasksLeft;

	SAHBuilder(const BVHParams &p, const std::deque<const Mesh *> *m, LeafList &l) : params(p), meshes(m), list(l) {
		const int hw = (int)std::thread::hardware_concurrency();
		tasksLeft = (hw > 1) ? 4 * hw : 0;
	}

	struct Part { u_int b, e; BBox box; };

	BBox BoundsOf(const u_int b, const u_int e) const {
		BBox r;
		for (u_int i = b; i < e; ++i)
			r = Union(r, list[i]->bbox);
		return r;
	}

	// Splits [b, e) (e - b >= 2) into two non-empty parts; returns the cut.
	u_int Split(const u_int b, const u_int e) {
		BBox cb;    // bounds of doubled centroids
		for (u_int i = b; i < e; ++i)
			cb = Union(cb, Point(Centroid2(list[i], 0), Centroid2(list[i], 1), Centroid2(list[i], 2)));

		float bestCost = INFINITY;
		int bestAxis = -1, bestBin = -1;
		for (int axis = 0; axis < 3; ++axis) {
			const float lo = cb.pMin[axis], extent = cb.pMax[axis] - cb.pMin[axis];
			if (!(extent > 0.f))
				continue;
			const float scale = kBins / extent;
			u_int count[kBins];
			BBox box[kBins];
			for (int k = 0; k < kBins; ++k) count[k] = 0;
			for (u_int i = b; i < e; ++i) {
				int k = (int)((Centroid2(list[i], axis) - lo) * scale);
				k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
				count[k]++;
				box[k] = Union(box[k], list[i]->bbox);
			}
			// sweep from the right, then evaluate cuts from the left
			float rightArea[kBins];
			u_int rightCount[kBins];
			BBox acc;
			u_int n = 0;
			for (int k = kBins - 1; k > 0; --k) {
				acc = Union(acc, box[k]);
				n += count[k];
				rightArea[k] = n ? acc.SurfaceArea() : 0.f;
				rightCount[k] = n;
			}
			acc = BBox();
			n = 0;
			for (int k = 1; k < kBins; ++k) {
				acc = Union(acc, box[k - 1]);
				n += count[k - 1];
				if (n == 0 || rightCount[k] == 0)
					continue;
				const float cost = acc.SurfaceArea() * n + rightArea[k] * rightCount[k];
				if (cost < bestCost) {
					bestCost = cost;
					bestAxis = axis;
					bestBin = k;
				}
			}
		}

		if (bestAxis >= 0) {
			const float lo = cb.pMin[bestAxis];
			const float scale = kBins / (cb.pMax[bestAxis] - cb.pMin[bestAxis]);
			const int ax = bestAxis, cut = bestBin;
			LeafList::iterator mid = std::partition(list.begin() + b, list.begin() + e, [=](BVHTreeNode *n) {
				int k = (int)((Centroid2(n, ax) - lo) * scale);
				k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
				return k < cut;
			});
			const u_int m = (u_int)(mid - list.begin());
			if (m > b && m < e)
				return m;
		}
		// all centroids coincide (or numerical trouble): split the run in half
		return b + (e - b) / 2;
	}

	void Build(const u_int b, const u_int e, std::vector<Node> &out, BBox *boxOut) {
		if (e - b == 1) {
			EmitLeaf(meshes, list[b], out);
			*boxOut = list[b]->bbox;
			return;
		}
		const size_t self = out.size();
		out.push_back(Node());

		// grow the child set: always split the child with the largest surface area
		std::vector<Part> parts;
		Part whole = { b, e, BoundsOf(b, e) };
		parts.push_back(whole);
		while (parts.size() < params.treeType) {
			int pick = -1;
			float pickArea = -1.f;
			for (size_t i = 0; i < parts.size(); ++i) {
				if (parts[i].e - parts[i].b < 2)
					continue;
				const float a = parts[i].box.SurfaceArea();
				if (a > pickArea) { pickArea = a; pick = (int)i; }
			}
			if (pick < 0)
				break;
			const Part p = parts[pick];
			const u_int m = Split(p.b, p.e);
			Part l = { p.b, m, BoundsOf(p.b, m) }, r = { m, p.e, BoundsOf(m, p.e) };
			parts[pick] = l;
			parts.insert(parts.begin() + pick + 1, r);
		}

		// big children are built concurrently into their own arrays and spliced in afterwards
		std::vector<std::future<void> > jobs(parts.size());
		std::vector<std::vector<Node> > sub(parts.size());
		std::vector<BBox> subBox(parts.size());
		std::vector<char> async(parts.size(), 0);
		if (e - b >= kParallelMin) {
			for (size_t i = 0; i < parts.size(); ++i) {
				if (parts[i].e - parts[i].b < kParallelMin / 4)
					continue;
				if (tasksLeft.fetch_sub(1) <= 0) {
					tasksLeft.fetch_add(1);
					continue;
				}
				async[i] = 1;
				const Part p = parts[i];
				std::vector<Node> *dst = &sub[i];
				BBox *bx = &subBox[i];
				jobs[i] = std::async(std::launch::async, [this, p, dst, bx]() {
					dst->reserve((size_t)(p.e - p.b) * 3 / 2);
					Build(p.b, p.e, *dst, bx);
					tasksLeft.fetch_add(1);
				});
			}
		}
		BBox bbox;
		for (size_t i = 0; i < parts.size(); ++i) {
			if (async[i]) {
				jobs[i].get();
				const u_int base = (u_int)out.size();
				for (size_t k = 0; k < sub[i].size(); ++k) {
					Node n = sub[i][k];
					n.nodeData = (n.nodeData & 0x80000000u) | ((n.nodeData & 0x7fffffffu) + base);
					out.push_back(n);
				}
				std::vector<Node>().swap(sub[i]);
				bbox = Union(bbox, subBox[i]);
			} else {
				BBox cb;
				Build(parts[i].b, parts[i].e, out, &cb);
				bbox = Union(bbox, cb);
			}
		}

		Node &n = out[self];
		memset(&n, 0, sizeof(n));
		StoreBox(n, bbox);
		n.nodeData = (u_int)out.size();
		*boxOut = bbox;
	}
};

}   // namespace


//------------------------------------------------------------------------------
// Binary SAH tree -> insertion-based optimisation -> optimal k-ary collapse
//------------------------------------------------------------------------------
//
// What the traversal pays for is node visits: with one triangle per leaf every triangle carries its
// own box, so the expected number of triangle tests of a ray does not depend on the topology, and
// the expected number of node visits is  sum over inner nodes of area(node) / area(root).  This
// builder minimises that sum in three steps:
//   1. a binary tree by top-down SAH (32 bins; exact sweep for small ranges),
//   2. insertion-based optimisation of that tree (remove a subtree, re-insert it where the summed
//      area of the inner nodes grows least, branch-and-bound search; after Bittner, Hapala and
//      Havran, "Fast insertion-based optimization of bounding volume hierarchies", 2013),
//   3. the k-ary tree with the smallest summed inner-node area that can be obtained from the binary
//      tree by dissolving inner nodes into their parents (dynamic programme over "subtree of n as a
//      forest of at most i trees", after Ylitie, Karras and Laine 2017, section 3.1).
// The result is emitted in the reference's array format like every other builder here.

namespace {

// Measured (tools/tree_stats.py, emulated traversal of bounce rays): kitchen 86 k triangles, 19.1 -> 17.4
// node visits per ray from the collapse alone, -> 16.2 with the optimisation (first pass gives 4/5 of the
// gain, ~140 search steps per node and pass); a uniform random soup gains nothing from either (50.8 vs
// 50.6 at 1 M triangles).  So: optimisation for scenes up to 256 k primitives (seconds), collapse up to
// 4 M, the plain top-down k-ary builder beyond.
static const size_t kOptMaxPrims = 4u << 20;
static const size_t kReinsertMaxPrims = 1u << 18;
static const int kReinsertPasses = 3;
static const size_t kMaxStepsPerSearch = 2048;      // typical searches take ~140 steps; boxes that all overlap would visit everything
static const size_t kOptStepBudget = 256u << 20;    // search steps of the optimisation (deterministic bound)

struct OptBuilder {
	const BVHParams &params;
	const std::deque<const Mesh *> *meshes;
	LeafList &list;
	const u_int N;
	SAHBuilder splitter;

	// binary tree: ids [0, N-1) are inner nodes, id N-1+i is the leaf list[i]
	std::vector<BBox> box;
	std::vector<float> area;
	std::vector<int> kid0, kid1, parent;
	int root;

	OptBuilder(const BVHParams &p, const std::deque<const Mesh *> *m, LeafList &l) : params(p), meshes(m), list(l),
			N((u_int)l.size()), splitter(p, m, l), root(-1) { }

	bool IsLeafId(const int id) const { return id >= (int)N - 1; }

	// ---- 1. binary SAH tree.  The inner node that cuts [b, e) at m gets id m - 1 (unique), so
	// sub-ranges are built concurrently without any shared allocation.
	int BuildBinary(const u_int b, const u_int e, const int depthBudget) {
		if (e - b == 1) {
			const int id = (int)(N - 1 + b);
			box[id] = list[b]->bbox;
			return id;
		}
		const u_int m = splitter.Split(b, e);
		const int id = (int)m - 1;
		int l, r;
		if (depthBudget > 0 && e - b >= SAHBuilder::kParallelMin) {
			std::future<int> job = std::async(std::launch::async, [this, b, m, depthBudget]() { return BuildBinary(b, m, depthBudget - 1); });
			r = BuildBinary(m, e, depthBudget - 1);
			l = job.get();
		} else {
			l = BuildBinary(b, m, 0);
			r = BuildBinary(m, e, 0);
		}
		kid0[id] = l; kid1[id] = r;
		parent[l] = id; parent[r] = id;
		box[id] = Union(box[l], box[r]);
		return id;
	}

	// ---- 2. insertion-based optimisation
	std::vector<int> height;    // leaves 0, inner nodes 1 + max over children
	int maxHeight;              // no move may make the tree deeper than this (the traversal stack is sized by it)

	void Refit(int id) {
		while (id >= 0) {
			const BBox nb = Union(box[kid0[id]], box[kid1[id]]);
			const float na = nb.SurfaceArea();
			const int nh = 1 + std::max(height[kid0[id]], height[kid1[id]]);
			if (nh == height[id] && na == area[id] && nb.pMin.x == box[id].pMin.x && nb.pMin.y == box[id].pMin.y && nb.pMin.z == box[id].pMin.z &&
					nb.pMax.x == box[id].pMax.x && nb.pMax.y == box[id].pMax.y && nb.pMax.z == box[id].pMax.z)
				break;
			box[id] = nb;
			area[id] = na;
			height[id] = nh;
			id = parent[id];
		}
	}

	struct Cand { float induced; int id; int depth; };
	struct CandLess { bool operator()(const Cand &a, const Cand &b) const { return a.induced > b.induced; } };

	// Takes subtree n out of the tree (its parent p goes with it) and puts it back at the cheapest
	// position -- where it was unless another position is STRICTLY cheaper (equal costs must not
	// reshuffle the tree: a set of identical boxes would end up as one long chain) and keeps the tree
	// within maxHeight.  Returns the number of search steps.
	size_t Reinsert(const int n, std::vector<Cand> &heap) {
		const int p = parent[n];
		if (p < 0)
			return 0;
		const int g = parent[p];
		if (g < 0)
			return 0;       // children of the root stay
		const int s = (kid0[p] == n) ? kid1[p] : kid0[p];
		// unlink p (and n with it)
		if (kid0[g] == p) kid0[g] = s; else kid1[g] = s;
		parent[s] = g;
		Refit(g);

		const BBox nb = box[n];
		const float an = area[n];
		const int hn = height[n];
		// cost of the position it came from
		float best = Union(box[s], nb).SurfaceArea();
		for (int a = g; a >= 0; a = parent[a])
			best += Union(box[a], nb).SurfaceArea() - area[a];
		int bestAt = s;
		size_t steps = 0;
		heap.clear();
		Cand c0 = { 0.f, root, 0 };
		heap.push_back(c0);
		while (!heap.empty()) {
			std::pop_heap(heap.begin(), heap.end(), CandLess());
			const Cand c = heap.back();
			heap.pop_back();
			if (c.induced + an >= best || steps >= kMaxStepsPerSearch)
				break;          // every remaining candidate is at least as expensive (or: heavily overlapping input, give up)
			++steps;
			const float direct = Union(box[c.id], nb).SurfaceArea();
			const float total = c.induced + direct;
			if (total < best && c.depth + 1 + std::max(height[c.id], hn) <= maxHeight) {
				best = total;
				bestAt = c.id;
			}
			if (!IsLeafId(c.id)) {
				const float below = total - area[c.id];
				if (below + an < best) {
					Cand a = { below, kid0[c.id], c.depth + 1 }, b2 = { below, kid1[c.id], c.depth + 1 };
					heap.push_back(a); std::push_heap(heap.begin(), heap.end(), CandLess());
					heap.push_back(b2); std::push_heap(heap.begin(), heap.end(), CandLess());
				}
			}
		}

		// link p above bestAt, with children (bestAt, n)
		const int x = bestAt, px = parent[x];
		if (px < 0)
			root = p;
		else if (kid0[px] == x) kid0[px] = p; else kid1[px] = p;
		parent[p] = px;
		kid0[p] = x; kid1[p] = n;
		parent[x] = p; parent[n] = p;
		box[p] = Union(box[x], nb);
		area[p] = box[p].SurfaceArea();
		height[p] = 1 + std::max(height[x], hn);
		if (px >= 0)
			Refit(px);
		return steps;
	}

	double InnerAreaSum() const {
		double s = 0.0;
		for (u_int i = 0; i + 1 < N; ++i) s += area[i];
		return s;
	}

	void Optimise(const int maxPasses, const size_t stepBudget) {
		std::vector<int> order;
		std::vector<Cand> heap;
		size_t steps = 0;
		double before = InnerAreaSum();
		for (int pass = 0; pass < maxPasses && steps < stepBudget; ++pass) {
			// largest subtrees first: moving them changes the most, and later moves see their new places
			order.resize(2 * (size_t)N - 1);
			for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
			std::sort(order.begin(), order.end(), [this](int a, int b) { return area[a] > area[b] || (area[a] == area[b] && a < b); });
			for (size_t i = 0; i < order.size() && steps < stepBudget; ++i)
				steps += Reinsert(order[i], heap);
			const double now = InnerAreaSum();
			if (getenv("LRB_BVH_VERBOSE"))
				fprintf(stderr, "[bvhbuild] pass %d: %zu search steps so far, inner area %.6g -> %.6g\n", pass, steps, before, now);
			if (!(now < before * 0.9995))
				break;
			before = now;
		}
	}

	// ---- 3. optimal collapse.  cost[(k-1) * n + (i-1)]: smallest summed area of the k-ary inner
	// nodes needed for the subtree of binary inner node n presented as at most i sibling entries.
	std::vector<float> cost;
	u_int K;

	float CostOf(const int id, const u_int i) const { return IsLeafId(id) ? 0.f : cost[(size_t)(K - 1) * id + (i - 1)]; }

	// best way to hand j entries to the two children of n: entries given to kid0
	float Distribute(const int n, const u_int j, u_int *give0) const {
		float bestC = std::numeric_limits<float>::infinity();
		u_int bestA = 1;
		for (u_int a = 1; a + 1 <= j; ++a) {
			const float c = CostOf(kid0[n], a) + CostOf(kid1[n], j - a);
			if (c < bestC) { bestC = c; bestA = a; }
		}
		if (give0) *give0 = bestA;
		return bestC;
	}

	void SolveCollapse() {
		K = params.treeType < 2 ? 2 : params.treeType;
		cost.assign((size_t)(K - 1) * (N - 1), 0.f);
		// post-order over the inner nodes
		std::vector<int> stack, order;
		order.reserve(N - 1);
		stack.push_back(root);
		while (!stack.empty()) {
			const int id = stack.back();
			stack.pop_back();
			if (IsLeafId(id)) continue;
			order.push_back(id);
			stack.push_back(kid0[id]);
			stack.push_back(kid1[id]);
		}
		for (size_t r = order.size(); r-- > 0;) {
			const int n = order[r];
			float *c = &cost[(size_t)(K - 1) * n];
			c[0] = area[n] + Distribute(n, K, nullptr);        // n survives as a k-ary node
			for (u_int i = 2; i <= K - 1; ++i)
				c[i - 1] = std::min(c[i - 2], Distribute(n, i, nullptr));   // or dissolves into i entries of its parent
		}
	}

	// entries of the k-ary node that presents subtree `id` within a budget of `i` entries
	void Entries(const int id, u_int i, std::vector<int> &outIds) const {
		if (IsLeafId(id)) { outIds.push_back(id); return; }
		while (i > 1 && CostOf(id, i) == CostOf(id, i - 1)) --i;
		if (i == 1) { outIds.push_back(id); return; }
		u_int a;
		Distribute(id, i, &a);
		Entries(kid0[id], a, outIds);
		Entries(kid1[id], i - a, outIds);
	}

	void Emit(const int id, std::vector<Node> &out) const {
		if (IsLeafId(id)) {
			EmitLeaf(meshes, list[id - (int)(N - 1)], out);
			return;
		}
		const size_t self = out.size();
		out.push_back(Node());
		std::vector<int> kids;
		u_int a;
		Distribute(id, K, &a);
		Entries(kid0[id], a, kids);
		Entries(kid1[id], K - a, kids);
		for (size_t i = 0; i < kids.size(); ++i)
			Emit(kids[i], out);
		Node &n = out[self];
		memset(&n, 0, sizeof(n));
		StoreBox(n, box[id]);
		n.nodeData = (u_int)out.size();
	}

	void Run(std::vector<Node> &out, const int optimisePasses, const size_t stepBudget) {
		if (N == 0)
			return;
		if (N == 1) {
			EmitLeaf(meshes, list[0], out);
			return;
		}
		box.resize(2 * (size_t)N - 1);
		area.resize(2 * (size_t)N - 1);
		parent.assign(2 * (size_t)N - 1, -1);
		kid0.assign(N - 1, -1);
		kid1.assign(N - 1, -1);
		root = BuildBinary(0, N, 4);
		for (size_t i = 0; i < box.size(); ++i) area[i] = box[i].SurfaceArea();
		// heights bottom-up (children of inner node m - 1 lie on either side of cut m: explicit post-order)
		height.assign(box.size(), 0);
		{
			std::vector<int> stack, order;
			stack.push_back(root);
			while (!stack.empty()) {
				const int id = stack.back();
				stack.pop_back();
				if (IsLeafId(id)) continue;
				order.push_back(id);
				stack.push_back(kid0[id]);
				stack.push_back(kid1[id]);
			}
			for (size_t r = order.size(); r-- > 0;)
				height[order[r]] = 1 + std::max(height[kid0[order[r]]], height[kid1[order[r]]]);
		}
		// head-room for the optimisation: half as much again as the SAH tree's own height (not binding on the
		// reference scenes; it is what keeps nested / coincident geometry from turning into a chain)
		{
			const char *e = getenv("LRB_BVH_HEADROOM");
			const int pct = e ? atoi(e) : 50;
			maxHeight = height[root] + height[root] * pct / 100 + 4;
		}
		const bool verbose = getenv("LRB_BVH_VERBOSE") != nullptr;
		const double a0 = InnerAreaSum() / area[root];
		if (optimisePasses > 0 && N >= 4)
			Optimise(optimisePasses, stepBudget);
		SolveCollapse();
		if (verbose)
			fprintf(stderr, "[bvhbuild] %u prims: binary SAH %.3f -> optimised %.3f -> %u-ary %.3f (expected node visits of a long random ray)\n",
					N, a0, InnerAreaSum() / area[root], K, (double)CostOf(root, 1) / area[root]);
		if (verbose)
			fprintf(stderr, "[bvhbuild] binary height %d (limit %d)\n", height[root], maxHeight);
		Emit(root, out);
	}
};

}   // namespace

Node *BuildEmbreeBVHBinnedSAH(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes, LeafList &leafList) {
	std::vector<Node> out;
	out.reserve(leafList.size() + leafList.size() / 2 + 1);
	const char *env = getenv("LRB_BVH_OPT");
	const int mode = env ? atoi(env) : 2;
	const char *envP = getenv("LRB_BVH_OPT_PASSES");
	const int passes = envP ? atoi(envP) : (leafList.size() <= kReinsertMaxPrims ? kReinsertPasses : 0);
	if (mode > 0 && leafList.size() <= kOptMaxPrims) {
		OptBuilder b(params, meshes, leafList);
		b.Run(out, mode >= 2 ? passes : 0, kOptStepBudget);
	} else if (!leafList.empty()) {
		SAHBuilder b(params, meshes, leafList);
		BBox box;
		b.Build(0, (u_int)leafList.size(), out, &box);
	}
	return ToArray(out, nNodes);
}

// EMBREE_MORTON (bvhembreebuild.cpp:218-336 with rtcBVHBuilderMorton): the fast builder.  Here: a linear BVH built ON
// THE GPU by the C ABI's lrb_build_lbvh (luxcore_b200/csrc/build_kernels.cuh) from the leaf boxes; this function
// only gathers the boxes and writes the leaf payload (vertex indices / instance records, which live in host
// tables) into the array the device returns.  The device is the process-wide builder device (LRB_BUILDER_DEVICE,
// default CUDA ordinal 0), opened on first use.  Without any CUDA device the host SAH builder answers instead
// (stated on stderr): tree topology never changes a closest hit.
static lrb_device *BuilderDevice() {
	static std::mutex mtx;
	static lrb_device *dev = nullptr;
	static bool tried = false;
	std::lock_guard<std::mutex> lock(mtx);
	if (!tried) {
		tried = true;
		int count = 0;
		if (lrb_device_count(&count) == LRB_OK && count > 0) {
			const char *env = getenv("LRB_BUILDER_DEVICE");
			if (lrb_device_create(env ? atoi(env) : 0, &dev) != LRB_OK)
				dev = nullptr;
		}
	}
	return dev;
}

// The builder device is one in-order queue shared by the whole process: builds from several host threads (several
// DataSets preprocessed at once) take turns.  (lrb_last_error_string is thread-local: read under the same lock.)
static std::mutex &BuilderQueueMutex() {
	static std::mutex m;
	return m;
}

static Node *BuildOnDevice(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes, LeafList &leafList, const uint32_t quality) {
	lrb_device *dev = leafList.empty() ? nullptr : BuilderDevice();
	if (!dev) {
		if (!leafList.empty())
			fprintf(stderr, "luxrays_b200: no CUDA device for the GPU BVH builder, using the host SAH builder\n");
		return BuildEmbreeBVHBinnedSAH(params, nNodes, meshes, leafList);
	}
	const size_t n = leafList.size();
	std::vector<float> boxes(6 * n);
	for (size_t i = 0; i < n; ++i) {
		const BBox &b = leafList[i]->bbox;
		float *o = &boxes[6 * i];
		o[0] = b.pMin.x; o[1] = b.pMin.y; o[2] = b.pMin.z;
		o[3] = b.pMax.x; o[4] = b.pMax.y; o[5] = b.pMax.z;
	}
	const size_t cap = 2 * n;
	Node *arr = new Node[cap];
	uint32_t total = 0;
	lrb_build_timings tm;
	static_assert(sizeof(Node) == sizeof(lrb_bvh_node), "BVHArrayNode layout");
	{
		std::lock_guard<std::mutex> turn(BuilderQueueMutex());
		if (lrb_build_bvh(dev, boxes.data(), (uint32_t)n, params.treeType, quality, reinterpret_cast<lrb_bvh_node *>(arr), (uint32_t)cap, &total, &tm) != LRB_OK) {
			delete[] arr;
			throw std::runtime_error(std::string("GPU BVH builder failed: ") + lrb_last_error_string());
		}
	}
	// leaf payload (the device wrote the input index of every leaf into the first word)
	for (uint32_t i = 0; i < total; ++i) {
		Node &an = arr[i];
		if (!(an.nodeData & 0x80000000u))
			continue;
		const BVHTreeNode *leaf = leafList[an.triangleLeaf.v[0]];
		const u_int nodeData = an.nodeData;
		memset(&an, 0, sizeof(an));
		if (meshes) {
			const Triangle &tri = (*meshes)[leaf->triangleLeaf.meshIndex]->GetTriangles()[leaf->triangleLeaf.triangleIndex];
			an.triangleLeaf.v[0] = tri.v[0];
			an.triangleLeaf.v[1] = tri.v[1];
			an.triangleLeaf.v[2] = tri.v[2];
			an.triangleLeaf.meshIndex = leaf->triangleLeaf.meshIndex;
			an.triangleLeaf.triangleIndex = leaf->triangleLeaf.triangleIndex;
		} else {
			an.bvhLeaf.leafIndex = leaf->bvhLeaf.leafIndex;
			an.bvhLeaf.transformIndex = leaf->bvhLeaf.transformIndex;
			an.bvhLeaf.motionIndex = leaf->bvhLeaf.motionIndex;
			an.bvhLeaf.meshOffsetIndex = leaf->bvhLeaf.meshOffsetIndex;
		}
		an.nodeData = nodeData;
	}
	*nNodes = total;
	return arr;
}

// All meshes' vertices back to back in dataset order (world-space positions for instances when the BVH was built with
// instance support disabled), first vertex of every mesh: what BVHKernel hands lrb_bvh_upload (bvhaccelhw.cpp:68-92).
void GatherSceneVertices(const std::deque<const Mesh *> &meshes, std::vector<float> &xyz, std::vector<uint32_t> &offsets) {
	size_t total = 0;
	for (size_t m = 0; m < meshes.size(); ++m)
		total += meshes[m]->GetTotalVertexCount();
	xyz.clear();
	offsets.clear();
	xyz.reserve(3 * total);
	for (size_t m = 0; m < meshes.size(); ++m) {
		const Mesh *mesh = meshes[m];
		offsets.push_back((uint32_t)(xyz.size() / 3));
		const u_int n = mesh->GetTotalVertexCount();
		if (mesh->GetType() == TYPE_TRIANGLE || mesh->GetType() == TYPE_EXT_TRIANGLE) {
			const float *src = reinterpret_cast<const float *>(mesh->GetVertices());
			xyz.insert(xyz.end(), src, src + 3 * (size_t)n);
		} else {
			for (u_int i = 0; i < n; ++i) {
				const Point p = mesh->GetVertex(Transform::TRANS_IDENTITY, i);
				xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z);
			}
		}
	}
}

bool BuildB200SceneOnDevice(const BVHParams &params, const u_int quality, const std::deque<const Mesh *> &meshes,
		Node **nodes, u_int *nNodes, void **scene, int *ordinal) {
	lrb_device *dev = BuilderDevice();
	if (!dev)
		return false;
	std::vector<float> xyz;
	std::vector<uint32_t> vertOff, triOff;
	GatherSceneVertices(meshes, xyz, vertOff);
	// triangle indices: luxrays::Triangle is three u_int (include/luxrays/core/geometry/triangle.h:35-53)
	static_assert(sizeof(Triangle) == 3 * sizeof(uint32_t), "Triangle layout");
	size_t nTris = 0;
	triOff.push_back(0u);
	for (size_t m = 0; m < meshes.size(); ++m) {
		nTris += meshes[m]->GetTotalTriangleCount();
		triOff.push_back((uint32_t)nTris);
	}
	if (nTris >= 0x3fffffffu)
		throw std::runtime_error("GPU BVH builder: too many triangles");
	const uint32_t *tris;
	std::vector<uint32_t> triBuf;
	if (meshes.size() == 1)
		tris = reinterpret_cast<const uint32_t *>(meshes[0]->GetTriangles());       // one mesh: its own array, no copy
	else {
		triBuf.reserve(3 * nTris);
		for (size_t m = 0; m < meshes.size(); ++m) {
			const uint32_t *src = reinterpret_cast<const uint32_t *>(meshes[m]->GetTriangles());
			triBuf.insert(triBuf.end(), src, src + 3 * (size_t)meshes[m]->GetTotalTriangleCount());
		}
		tris = triBuf.data();
	}
	const size_t cap = 2 * nTris;
	Node *arr = new Node[cap];
	uint32_t total = 0;
	lrb_scene *sc = nullptr;
	static_assert(sizeof(Node) == sizeof(lrb_bvh_node), "BVHArrayNode layout");
	{
		std::lock_guard<std::mutex> turn(BuilderQueueMutex());
		if (lrb_bvh_build_scene(dev, xyz.data(), xyz.size() / 3, vertOff.data(), triOff.data(), (uint32_t)meshes.size(), tris, params.treeType, quality,
				&sc, reinterpret_cast<lrb_bvh_node *>(arr), (uint32_t)cap, &total, nullptr) != LRB_OK) {
			delete[] arr;
			throw std::runtime_error(std::string("GPU BVH builder failed: ") + lrb_last_error_string());
		}
	}
	lrb_device_props props;
	*ordinal = lrb_device_get_props(dev, &props) == LRB_OK ? props.cuda_ordinal : -1;
	*nodes = arr;
	*nNodes = total;
	*scene = sc;
	return true;
}

void FreeB200ResidentScene(void *scene) {
	lrb_scene_free(static_cast<lrb_scene *>(scene));
}

Node *BuildEmbreeBVHMorton(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes, LeafList &leafList) {
	return BuildOnDevice(params, nNodes, meshes, leafList, 0u);
}

// accelerator.bvh.builder.type = B200_PLOC (an extension of the reference's three names): the GPU builder with its
// PLOC binary tree -- SAH-class quality at GPU build speed, for scenes whose host build would take seconds.
Node *BuildB200BVHPloc(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes, LeafList &leafList) {
	return BuildOnDevice(params, nNodes, meshes, leafList, 1u);
}

}   // namespace luxrays
