// luxrays/core/trianglemesh.h -- mesh classes handed to DataSet::Add (reference:
// include/luxrays/core/trianglemesh.h:48-391, src/luxrays/core/trianglemesh.cpp:40-284).
// Only what the intersection path touches is kept: type, vertices/triangles, counts, bounding
// box, GetVertex.  Area/Sample/serialization belong to the renderer and are out of scope.
#ifndef _LUXRAYS_B200_TRIANGLEMESH_H
#define _LUXRAYS_B200_TRIANGLEMESH_H

#include "luxrays/luxrays.h"

namespace luxrays {

typedef enum {
	TYPE_TRIANGLE, TYPE_TRIANGLE_INSTANCE, TYPE_TRIANGLE_MOTION,
	TYPE_EXT_TRIANGLE, TYPE_EXT_TRIANGLE_INSTANCE, TYPE_EXT_TRIANGLE_MOTION
} MeshType;

class Mesh {
public:
	Mesh() { }
	virtual ~Mesh() { }

	virtual MeshType GetType() const = 0;
	virtual BBox GetBBox() const = 0;
	virtual void GetLocal2World(const float time, Transform &local2World) const = 0;
	virtual Point GetVertex(const Transform &local2World, const u_int vertIndex) const = 0;
	virtual Point *GetVertices() const = 0;
	virtual Triangle *GetTriangles() const = 0;
	virtual u_int GetTotalVertexCount() const = 0;
	virtual u_int GetTotalTriangleCount() const = 0;
	virtual void ApplyTransform(const Transform &trans) = 0;
};

class TriangleMesh : virtual public Mesh {
public:
	// Vertices must come from AllocVerticesBuffer (it appends a sentinel float that the
	// constructor checks); ownership of both arrays stays with the application.
	TriangleMesh(const u_int meshVertCount, const u_int meshTriCount, Point *meshVertices, Triangle *meshTris);
	virtual ~TriangleMesh() { }
	void Delete() { delete[] reinterpret_cast<float *>(vertices); delete[] tris; }

	virtual MeshType GetType() const { return TYPE_TRIANGLE; }
	virtual BBox GetBBox() const;
	virtual void GetLocal2World(const float time, Transform &local2World) const { local2World = appliedTrans; }
	virtual Point GetVertex(const Transform &local2World, const u_int vertIndex) const { return vertices[vertIndex]; }
	virtual Point *GetVertices() const { return vertices; }
	virtual Triangle *GetTriangles() const { return tris; }
	virtual u_int GetTotalVertexCount() const { return vertCount; }
	virtual u_int GetTotalTriangleCount() const { return triCount; }
	virtual void ApplyTransform(const Transform &trans);

	static Point *AllocVerticesBuffer(const u_int meshVertCount) {
		float *buffer = new float[3 * (size_t)meshVertCount + 1];
		buffer[3 * (size_t)meshVertCount] = 1234.1234f;
		return reinterpret_cast<Point *>(buffer);
	}
	static Triangle *AllocTrianglesBuffer(const u_int meshTriCount) { return new Triangle[meshTriCount]; }

protected:
	u_int vertCount, triCount;
	Point *vertices;
	Triangle *tris;
	Transform appliedTrans;
	mutable BBox cachedBBox;
	mutable bool cachedBBoxValid;
};

class InstanceTriangleMesh : virtual public Mesh {
public:
	InstanceTriangleMesh(TriangleMesh *m, const Transform &t);
	virtual ~InstanceTriangleMesh() { }

	virtual MeshType GetType() const { return TYPE_TRIANGLE_INSTANCE; }
	virtual BBox GetBBox() const;
	virtual void GetLocal2World(const float time, Transform &local2World) const { local2World = trans; }
	virtual Point GetVertex(const Transform &local2World, const u_int vertIndex) const {
		return trans * mesh->GetVertex(local2World, vertIndex);
	}
	virtual Point *GetVertices() const { return mesh->GetVertices(); }
	virtual Triangle *GetTriangles() const { return mesh->GetTriangles(); }
	virtual u_int GetTotalVertexCount() const { return mesh->GetTotalVertexCount(); }
	virtual u_int GetTotalTriangleCount() const { return mesh->GetTotalTriangleCount(); }
	virtual void ApplyTransform(const Transform &t) { trans = trans * t; cachedBBoxValid = false; }

	const Transform &GetTransformation() const { return trans; }
	void SetTransformation(const Transform &t) { trans = t; cachedBBoxValid = false; }
	TriangleMesh *GetTriangleMesh() const { return mesh; }

protected:
	Transform trans;
	TriangleMesh *mesh;
	mutable BBox cachedBBox;
	mutable bool cachedBBoxValid;
};

class MotionTriangleMesh : virtual public Mesh {
public:
	MotionTriangleMesh(TriangleMesh *m, const MotionSystem &ms);
	virtual ~MotionTriangleMesh() { }

	virtual MeshType GetType() const { return TYPE_TRIANGLE_MOTION; }
	virtual BBox GetBBox() const;
	virtual void GetLocal2World(const float time, Transform &local2World) const {
		local2World = Transform(motionSystem.SampleInverse(time));
	}
	virtual Point GetVertex(const Transform &local2World, const u_int vertIndex) const {
		return local2World * mesh->GetVertex(local2World, vertIndex);
	}
	virtual Point *GetVertices() const { return mesh->GetVertices(); }
	virtual Triangle *GetTriangles() const { return mesh->GetTriangles(); }
	virtual u_int GetTotalVertexCount() const { return mesh->GetTotalVertexCount(); }
	virtual u_int GetTotalTriangleCount() const { return mesh->GetTotalTriangleCount(); }
	virtual void ApplyTransform(const Transform &t);

	TriangleMesh *GetTriangleMesh() const { return mesh; }
	const MotionSystem &GetMotionSystem() const { return motionSystem; }

protected:
	MotionSystem motionSystem;
	TriangleMesh *mesh;
	mutable BBox cachedBBox;
	mutable bool cachedBBoxValid;
};

}   // namespace luxrays

#endif
