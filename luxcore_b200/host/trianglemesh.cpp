// trianglemesh.cpp -- mesh classes (reference: src/luxrays/core/trianglemesh.cpp:40-83,222-284).
#include "luxrays/core/trianglemesh.h"

namespace luxrays {

TriangleMesh::TriangleMesh(const u_int meshVertCount, const u_int meshTriCount, Point *meshVertices, Triangle *meshTris) {
	if (!meshVertices || !meshTris)
		throw std::runtime_error("luxrays::TriangleMesh() needs vertex and triangle buffers");
	// the sentinel written by AllocVerticesBuffer proves where the buffer came from
	const float *raw = reinterpret_cast<const float *>(meshVertices);
	if (raw[3 * (size_t)meshVertCount] != 1234.1234f)
		throw std::runtime_error("luxrays::TriangleMesh() used with a vertex buffer not allocated with luxrays::TriangleMesh::AllocVerticesBuffer()");
	vertCount = meshVertCount;
	triCount = meshTriCount;
	vertices = meshVertices;
	tris = meshTris;
	cachedBBoxValid = false;
}

BBox TriangleMesh::GetBBox() const {
	if (!cachedBBoxValid) {
		BBox bbox;
		for (u_int i = 0; i < vertCount; ++i)
			bbox = Union(bbox, vertices[i]);
		cachedBBox = bbox;
		cachedBBoxValid = true;
	}
	return cachedBBox;
}

void TriangleMesh::ApplyTransform(const Transform &trans) {
	appliedTrans = appliedTrans * trans;
	for (u_int i = 0; i < vertCount; ++i)
		vertices[i] *= trans;
	cachedBBoxValid = false;
}

InstanceTriangleMesh::InstanceTriangleMesh(TriangleMesh *m, const Transform &t) : trans(t), mesh(m), cachedBBoxValid(false) {
	if (!m)
		throw std::runtime_error("InstanceTriangleMesh needs a base mesh");
}

// bounding box of the 8 transformed corners of the base mesh's box
BBox InstanceTriangleMesh::GetBBox() const {
	if (!cachedBBoxValid) {
		cachedBBox = trans * mesh->GetBBox();
		cachedBBoxValid = true;
	}
	return cachedBBox;
}

MotionTriangleMesh::MotionTriangleMesh(TriangleMesh *m, const MotionSystem &ms) : motionSystem(ms), mesh(m), cachedBBoxValid(false) {
	if (!m)
		throw std::runtime_error("MotionTriangleMesh needs a base mesh");
}

// union over the sampled motion; the system stores global->local matrices
BBox MotionTriangleMesh::GetBBox() const {
	if (!cachedBBoxValid) {
		cachedBBox = motionSystem.Bound(mesh->GetBBox(), true);
		cachedBBoxValid = true;
	}
	return cachedBBox;
}

void MotionTriangleMesh::ApplyTransform(const Transform &t) {
	// MotionSystem::ApplyTransform (motionsystem.cpp:367-378): re-key every knot with knot * t
	std::vector<Transform> xf;
	const std::vector<InterpolatedTransform> &its = motionSystem.interpolatedTransforms;
	for (size_t i = 1; i + 1 < its.size(); ++i)
		xf.push_back(its[i].start * t);
	xf.push_back(its[its.size() - 2].end * t);
	motionSystem = MotionSystem(motionSystem.times, xf);
	cachedBBoxValid = false;
}

}   // namespace luxrays
