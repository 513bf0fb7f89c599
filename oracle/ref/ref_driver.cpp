// ref_driver.cpp -- TEST INFRASTRUCTURE.  C entry points over the REFERENCE's own classes, compiled
// from the sources under /root/reference (see Makefile): luxrays::TriangleMesh / InstanceTriangleMesh /
// MotionTriangleMesh, BVHAccel, MBVHAccel, Triangle::Intersect, BBox::IntersectP, MachineEpsilon,
// Matrix4x4::Inverse, MotionSystem::Sample.  Used only to pin oracle/lux_oracle.cpp (tests/) and as the
// `reference` CPU baseline of bench.py.  Same entry-point shapes as the oracle's orc_* API.
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "luxrays/luxrays.h"
#include "luxrays/core/epsilon.h"
#include "luxrays/core/context.h"
#include "luxrays/core/trianglemesh.h"
#include "luxrays/core/geometry/triangle.h"
#include "luxrays/core/geometry/bbox.h"
#include "luxrays/core/geometry/matrix4x4.h"
#include "luxrays/core/geometry/transform.h"
#include "luxrays/core/geometry/motionsystem.h"
#include "luxrays/accelerators/bvhaccel.h"
#include "luxrays/accelerators/mbvhaccel.h"

namespace luxrays {
// BVHKernel / MBVHKernel are friends of the accelerators (bvhaccel.h:55-57, mbvhaccel.h:56): the
// same door the hardware kernels use to read the flattened arrays.
class BVHKernel {
public:
	static const ocl::BVHArrayNode *Nodes(const BVHAccel &a, u_int *n) { *n = a.nNodes; return a.bvhTree; }
	// Replaces the tree of an initialised accelerator with a caller-supplied array in the same format
	// (e.g. the product's binned-SAH tree), so that BVHAccel::Intersect can be timed / compared on it.
	static void SetNodes(BVHAccel &a, const ocl::BVHArrayNode *nodes, u_int n) {
		delete[] a.bvhTree;
		a.bvhTree = new ocl::BVHArrayNode[n];
		memcpy(a.bvhTree, nodes, (size_t)n * sizeof(ocl::BVHArrayNode));
		a.nNodes = n;
	}
};
class MBVHKernel {
public:
	static const ocl::BVHArrayNode *Root(const MBVHAccel &a, u_int *n) { *n = a.nRootNodes; return a.bvhRootTree; }
	static size_t LeafCount(const MBVHAccel &a) { return a.uniqueLeafs.size(); }
	static const ocl::BVHArrayNode *Leaf(const MBVHAccel &a, size_t i, u_int *n) { return BVHKernel::Nodes(*a.uniqueLeafs[i], n); }
	static size_t TransformCount(const MBVHAccel &a) { return a.uniqueLeafsTransform.size(); }
	static const Transform *Xform(const MBVHAccel &a, size_t i) { return a.uniqueLeafsTransform[i]; }
	static size_t MotionCount(const MBVHAccel &a) { return a.uniqueLeafsMotionSystem.size(); }
	static const MotionSystem *Motion(const MBVHAccel &a, size_t i) { return a.uniqueLeafsMotionSystem[i]; }
};
}

using namespace luxrays;

namespace {

std::string g_err;

struct RefScene {
	std::vector<TriangleMesh *> shapes;
	std::deque<const Mesh *> meshes;
	std::vector<Mesh *> owned;
	u_longlong totalVerts, totalTris;
	RefScene() : totalVerts(0), totalTris(0) { }
	~RefScene() {
		for (size_t i = 0; i < owned.size(); ++i) delete owned[i];
		for (size_t i = 0; i < shapes.size(); ++i) {
			// the reference's meshes do not own their buffers unless Delete() is called
			shapes[i]->Delete();
			delete shapes[i];
		}
	}
};

Matrix4x4 ToMatrix(const float *m16) {
	float m[4][4];
	memcpy(m, m16, 64);
	return Matrix4x4(m);
}

struct RefAccel {
	Context *ctx;
	Accelerator *accel;
	RefAccel() : ctx(NULL), accel(NULL) { }
	~RefAccel() { delete accel; delete ctx; }
};

Context *NewContext(int treeType, int costSamples, int isectCost, int travCost, float emptyBonus) {
	Properties cfg;
	cfg.Set(Property("accelerator.bvh.builder.type")("CLASSIC"));
	cfg.Set(Property("accelerator.bvh.treetype")(treeType));
	cfg.Set(Property("accelerator.bvh.costsamples")(costSamples));
	cfg.Set(Property("accelerator.bvh.isectcost")(isectCost));
	cfg.Set(Property("accelerator.bvh.travcost")(travCost));
	cfg.Set(Property("accelerator.bvh.emptybonus")(emptyBonus));
	return new Context(NULL, cfg);
}

template <class F> void ParallelFor(uint64_t n, int nthreads, F f) {
	if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
	if (nthreads <= 1 || n < 1024) { f(0, n); return; }
	std::vector<std::thread> th;
	const uint64_t per = (n + nthreads - 1) / nthreads;
	for (int t = 0; t < nthreads; ++t) {
		const uint64_t b = std::min<uint64_t>(n, t * per), e = std::min<uint64_t>(n, b + per);
		if (b < e) th.push_back(std::thread(f, b, e));
	}
	for (size_t i = 0; i < th.size(); ++i) th[i].join();
}

}   // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API const char *ref_last_error() { return g_err.c_str(); }
REF_API const char *ref_describe() {
	return "LuxRays reference sources compiled from /root/reference (Triangle::Intersect, BBox::IntersectP, CLASSIC builder, "
		"BVHAccel, MBVHAccel, Transform, Matrix4x4, MotionSystem, Quaternion, MachineEpsilon, TriangleMesh) with no-op Boost/Embree shims";
}
REF_API void ref_set_epsilon(float mn, float mx) { MachineEpsilon::SetMin(mn); MachineEpsilon::SetMax(mx); }
REF_API float ref_epsilon(float v) { return MachineEpsilon::E(v); }

REF_API int ref_matrix_inverse(const float *m16, float *out16) {
	try {
		const Matrix4x4 inv = ToMatrix(m16).Inverse();
		memcpy(out16, inv.m, 64);
		return 0;
	} catch (const std::exception &e) { g_err = e.what(); return 1; }
}

// tb = { t, b1, b2 }
REF_API int ref_triangle_intersect(const Ray *r, const float *p0, const float *p1, const float *p2, float *tb) {
	Point verts[3] = { Point(p0[0], p0[1], p0[2]), Point(p1[0], p1[1], p1[2]), Point(p2[0], p2[1], p2[2]) };
	const Triangle tri(0, 1, 2);
	float t = 0.f, b1 = 0.f, b2 = 0.f;
	const bool hit = tri.Intersect(*r, verts, &t, &b1, &b2);
	tb[0] = t; tb[1] = b1; tb[2] = b2;
	return hit ? 1 : 0;
}

REF_API int ref_bbox_intersectp(const Ray *r, const float *bmin, const float *bmax) {
	BBox b;      // the two-point constructor would sort the corners: set them as given
	b.pMin = Point(bmin[0], bmin[1], bmin[2]);
	b.pMax = Point(bmax[0], bmax[1], bmax[2]);
	return b.IntersectP(*r) ? 1 : 0;
}

REF_API void *ref_scene_create() { return new RefScene(); }
REF_API void ref_scene_free(void *s) { delete (RefScene *)s; }
REF_API int ref_scene_mesh_count(void *s) { return (int)((RefScene *)s)->meshes.size(); }

REF_API int ref_scene_add_shape(void *sp, const float *xyz, uint32_t nVerts, const uint32_t *tris, uint32_t nTris) {
	RefScene *s = (RefScene *)sp;
	try {
		Point *v = TriangleMesh::AllocVerticesBuffer(nVerts);
		Triangle *t = TriangleMesh::AllocTrianglesBuffer(nTris);
		for (uint32_t i = 0; i < nVerts; ++i) v[i] = Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
		for (uint32_t i = 0; i < nTris; ++i) t[i] = Triangle(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]);
		s->shapes.push_back(new TriangleMesh(nVerts, nTris, v, t));
		return (int)s->shapes.size() - 1;
	} catch (const std::exception &e) { g_err = e.what(); return -1; }
}

static int AddMesh(RefScene *s, const Mesh *m, Mesh *own) {
	s->meshes.push_back(m);
	if (own) s->owned.push_back(own);
	s->totalVerts += m->GetTotalVertexCount();
	s->totalTris += m->GetTotalTriangleCount();
	return (int)s->meshes.size() - 1;
}

REF_API int ref_scene_add_plain(void *sp, int shape) {
	RefScene *s = (RefScene *)sp;
	return AddMesh(s, s->shapes[shape], NULL);
}

REF_API int ref_scene_add_instance(void *sp, int shape, const float *m16) {
	RefScene *s = (RefScene *)sp;
	try {
		InstanceTriangleMesh *m = new InstanceTriangleMesh(s->shapes[shape], Transform(ToMatrix(m16)));
		return AddMesh(s, m, m);
	} catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// m16s: nKeys world->local matrices, as the scene parser stores them (parseobjects.cpp:155-157)
REF_API int ref_scene_add_motion(void *sp, int shape, uint32_t nKeys, const float *times, const float *m16s) {
	RefScene *s = (RefScene *)sp;
	try {
		std::vector<float> t(times, times + nKeys);
		std::vector<Transform> x;
		for (uint32_t i = 0; i < nKeys; ++i) x.push_back(Transform(ToMatrix(m16s + 16 * i)));
		MotionTriangleMesh *m = new MotionTriangleMesh(s->shapes[shape], MotionSystem(t, x));
		return AddMesh(s, m, m);
	} catch (const std::exception &e) { g_err = e.what(); return -1; }
}

REF_API int ref_scene_set_instance_transform(void *sp, int mesh, const float *m16) {
	RefScene *s = (RefScene *)sp;
	InstanceTriangleMesh *m = dynamic_cast<InstanceTriangleMesh *>(const_cast<Mesh *>(s->meshes[mesh]));
	if (!m) { g_err = "not an instance"; return 1; }
	m->SetTransformation(Transform(ToMatrix(m16)));
	return 0;
}

REF_API int ref_scene_mesh_bbox(void *sp, int mesh, float *out6) {
	const BBox b = ((RefScene *)sp)->meshes[mesh]->GetBBox();
	out6[0] = b.pMin.x; out6[1] = b.pMin.y; out6[2] = b.pMin.z;
	out6[3] = b.pMax.x; out6[4] = b.pMax.y; out6[5] = b.pMax.z;
	return 0;
}

REF_API void *ref_bvh_build(void *sp, int treeType, int costSamples, int isectCost, int travCost, float emptyBonus) {
	RefScene *s = (RefScene *)sp;
	RefAccel *a = new RefAccel();
	try {
		a->ctx = NewContext(treeType, costSamples, isectCost, travCost, emptyBonus);
		BVHAccel *bvh = new BVHAccel(a->ctx);
		a->accel = bvh;
		bvh->Init(s->meshes, s->totalVerts, s->totalTris);
		return a;
	} catch (const std::exception &e) { g_err = e.what(); delete a; return NULL; }
}
REF_API int ref_bvh_set_nodes(void *ap, const void *nodes, uint32_t n) {
	BVHAccel *bvh = dynamic_cast<BVHAccel *>(((RefAccel *)ap)->accel);
	if (!bvh || !nodes || !n) { g_err = "not a BVH accelerator / empty array"; return 1; }
	BVHKernel::SetNodes(*bvh, (const ocl::BVHArrayNode *)nodes, n);
	return 0;
}
REF_API void ref_accel_free(void *a) { delete (RefAccel *)a; }
REF_API uint32_t ref_bvh_node_count(void *ap) { u_int n; BVHKernel::Nodes(*(BVHAccel *)((RefAccel *)ap)->accel, &n); return n; }
REF_API const void *ref_bvh_nodes(void *ap) { u_int n; return BVHKernel::Nodes(*(BVHAccel *)((RefAccel *)ap)->accel, &n); }

// Accelerator::Intersect over a batch (the RayHit of every ray is initialised by Intersect itself)
REF_API int ref_accel_intersect(void *ap, const Ray *rays, RayHit *hits, uint64_t n, int nthreads) {
	const Accelerator *acc = ((RefAccel *)ap)->accel;
	try {
		ParallelFor(n, nthreads, [=](uint64_t b, uint64_t e) {
			for (uint64_t i = b; i < e; ++i) {
				// the CPU path leaves b1/b2/triangleIndex of a miss untouched: give them the same
				// defined content the oracle uses
				hits[i].b1 = 0.f; hits[i].b2 = 0.f; hits[i].triangleIndex = 0xffffffffu;
				acc->Intersect(&rays[i], &hits[i]);
			}
		});
		return 0;
	} catch (const std::exception &e) { g_err = e.what(); return 1; }
}

REF_API void *ref_mbvh_build(void *sp, int treeType, int costSamples, int isectCost, int travCost, float emptyBonus) {
	RefScene *s = (RefScene *)sp;
	RefAccel *a = new RefAccel();
	try {
		a->ctx = NewContext(treeType, costSamples, isectCost, travCost, emptyBonus);
		MBVHAccel *m = new MBVHAccel(a->ctx);
		a->accel = m;
		m->Init(s->meshes, s->totalVerts, s->totalTris);
		return a;
	} catch (const std::exception &e) { g_err = e.what(); delete a; return NULL; }
}
REF_API int ref_mbvh_update(void *ap) {
	try { ((RefAccel *)ap)->accel->Update(); return 0; } catch (const std::exception &e) { g_err = e.what(); return 1; }
}
REF_API uint32_t ref_mbvh_root_node_count(void *ap) { u_int n; MBVHKernel::Root(*(MBVHAccel *)((RefAccel *)ap)->accel, &n); return n; }
REF_API const void *ref_mbvh_root_nodes(void *ap) { u_int n; return MBVHKernel::Root(*(MBVHAccel *)((RefAccel *)ap)->accel, &n); }
REF_API uint32_t ref_mbvh_leaf_count(void *ap) { return (uint32_t)MBVHKernel::LeafCount(*(MBVHAccel *)((RefAccel *)ap)->accel); }
REF_API uint32_t ref_mbvh_leaf_node_count(void *ap, uint32_t i) { u_int n; MBVHKernel::Leaf(*(MBVHAccel *)((RefAccel *)ap)->accel, i, &n); return n; }
REF_API const void *ref_mbvh_leaf_nodes(void *ap, uint32_t i) { u_int n; return MBVHKernel::Leaf(*(MBVHAccel *)((RefAccel *)ap)->accel, i, &n); }
REF_API uint32_t ref_mbvh_transform_count(void *ap) { return (uint32_t)MBVHKernel::TransformCount(*(MBVHAccel *)((RefAccel *)ap)->accel); }
REF_API void ref_mbvh_transform_minv(void *ap, uint32_t i, float *out16) {
	memcpy(out16, MBVHKernel::Xform(*(MBVHAccel *)((RefAccel *)ap)->accel, i)->mInv.m, 64);
}
REF_API uint32_t ref_mbvh_motion_count(void *ap) { return (uint32_t)MBVHKernel::MotionCount(*(MBVHAccel *)((RefAccel *)ap)->accel); }
REF_API int ref_motion_sample(void *ap, uint32_t i, float time, float *out16) {
	try {
		const Matrix4x4 m = MBVHKernel::Motion(*(MBVHAccel *)((RefAccel *)ap)->accel, i)->Sample(time);
		memcpy(out16, m.m, 64);
		return 0;
	} catch (const std::exception &e) { g_err = e.what(); return 1; }
}
REF_API int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
