// bvhbuild.cpp -- host-side BVH builders emitting the reference's BVHArrayNode skip-list array
// (array rules: src/luxrays/core/bvh/bvhclassicbuild.cpp:181-220 -- depth-first pre-order, first
// child at index+1, skip index = first node after the subtree, leaf skip = index+1 with bit 31 set).
//
//   BuildBVH (CLASSIC)       : restatement of bvhclassicbuild.cpp:51-233.  Emits the array directly
//                              during the recursion instead of building a pointer tree and
//                              flattening it; partitions, split values and box unions are the same
//                              operations in the same order, so the array is bit-identical.
//   BuildEmbreeBVHBinnedSAH /
//   BuildEmbreeBVHMorton     : the reference delegates these to Intel Embree 3.12.2 (not part of
//                              /root/reference).  Replaced by a from-scratch binned-SAH k-ary
//                              builder (16 bins, 3 axes, largest-area child split first, one
//                              primitive per leaf, at most treeType children per node), multi-
//                              threaded over sub-trees.  Closest-hit results do not depend on the
//                              topology.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <future>
#include <thread>

#include "luxrays/core/bvh/bvhbuild.h"

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

Node *BuildEmbreeBVHBinnedSAH(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes, LeafList &leafList) {
	SAHBuilder b(params, meshes, leafList);
	std::vector<Node> out;
	out.reserve(leafList.size() + leafList.size() / 2 + 1);
	if (!leafList.empty()) {
		BBox box;
		b.Build(0, (u_int)leafList.size(), out, &box);
	}
	return ToArray(out, nNodes);
}

Node *BuildEmbreeBVHMorton(const BVHParams &params, u_int *nNodes, const std::deque<const Mesh *> *meshes, LeafList &leafList) {
	// quality/speed trade-off of the Morton builder is not reproduced; same SAH builder
	return BuildEmbreeBVHBinnedSAH(params, nNodes, meshes, leafList);
}

}   // namespace luxrays
