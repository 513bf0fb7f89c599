// accelerators.cpp -- Accelerator helpers, BVHAccel and MBVHAccel host sides (build only).
// Reference: src/luxrays/core/accelerator.cpp:25-55, src/luxrays/accelerators/bvhaccel.cpp:35-168,
// src/luxrays/accelerators/mbvhaccel.cpp:40-250.  The CPU Intersect() of both classes is NOT
// provided (no CPU intersection code in the product); it throws.
#include <map>

#include "luxrays/accelerators/bvhaccel.h"
#include "luxrays/accelerators/mbvhaccel.h"
#include "luxrays/core/context.h"
#include "luxrays/core/intersectiondevice.h"

namespace luxrays {

std::string Accelerator::AcceleratorType2String(const AcceleratorType type) {
	switch (type) {
		case ACCEL_AUTO: return "AUTO";
		case ACCEL_BVH: return "BVH";
		case ACCEL_MBVH: return "MBVH";
		case ACCEL_EMBREE: return "EMBREE";
		case ACCEL_OPTIX: return "OPTIX";
		default: throw std::runtime_error("Unknown AcceleratorType in AcceleratorType2String()");
	}
}

AcceleratorType Accelerator::String2AcceleratorType(const std::string &type) {
	if (type == "AUTO") return ACCEL_AUTO;
	if (type == "BVH") return ACCEL_BVH;
	if (type == "MBVH") return ACCEL_MBVH;
	if (type == "EMBREE") return ACCEL_EMBREE;
	if (type == "OPTIX") return ACCEL_OPTIX;
	throw std::runtime_error("Unknown accelerator type: " + type);
}

static ocl::BVHArrayNode *RunBuilder(const Context *ctx, const char *who, const BVHParams &params, u_int *nNodes,
		const std::deque<const Mesh *> *meshes, std::vector<BVHTreeNode *> &list) {
	const std::string builderType = ctx->GetConfig().Get(Property("accelerator.bvh.builder.type")("EMBREE_BINNED_SAH")).Get<std::string>();
	LR_LOG(ctx, who << " builder: " << builderType);
	if (builderType == "CLASSIC")
		return BuildBVH(params, nNodes, meshes, list);
	if (builderType == "EMBREE_BINNED_SAH")
		return BuildEmbreeBVHBinnedSAH(params, nNodes, meshes, list);
	if (builderType == "EMBREE_MORTON")
		return BuildEmbreeBVHMorton(params, nNodes, meshes, list);
	if (builderType == "B200_PLOC")
		return BuildB200BVHPloc(params, nNodes, meshes, list);
	throw std::runtime_error(std::string("Unknown BVH builder type in ") + who + ": " + builderType);
}

//------------------------------------------------------------------------------
// BVHAccel
//------------------------------------------------------------------------------

BVHAccel::BVHAccel(const Context *context) : nNodes(0), bvhTree(nullptr), ctx(context),
		totalVertexCount(0), totalTriangleCount(0), initialized(false),
		allowResidentScene(true), residentScene(nullptr), residentOrdinal(-1) {
	params = ToBVHParams(ctx->GetConfig());
}

BVHAccel::~BVHAccel() {
	delete[] bvhTree;
	if (residentScene)
		FreeB200ResidentScene(residentScene);
}

BVHParams BVHAccel::ToBVHParams(const Properties &props) {
	const int treeType = props.Get(Property("accelerator.bvh.treetype")(4)).Get<int>();
	BVHParams p;
	p.treeType = (treeType <= 2) ? 2 : ((treeType <= 4) ? 4 : 8);
	p.costSamples = props.Get(Property("accelerator.bvh.costsamples")(0)).Get<int>();
	p.isectCost = props.Get(Property("accelerator.bvh.isectcost")(80)).Get<int>();
	p.traversalCost = props.Get(Property("accelerator.bvh.travcost")(10)).Get<int>();
	p.emptyBonus = props.Get(Property("accelerator.bvh.emptybonus")(.5f)).Get<float>();
	return p;
}

void BVHAccel::Init(const std::deque<const Mesh *> &ms, const u_longlong totVert, const u_longlong totTri) {
	meshes = ms;
	totalVertexCount = totVert;
	totalTriangleCount = totTri;
	if (totalTriangleCount == 0) {
		LR_LOG(ctx, "Empty BVH");
		nNodes = 0;
		bvhTree = nullptr;
		initialized = true;
		return;
	}

	// GPU builders: boxes, tree, leaf payload and the device lay-out are all made on the device (no BVHTreeNode list, no
	// host re-layout at upload); the reference array is downloaded once for bvhTree.  "accelerator.b200.resident" = false
	// keeps the two-step path (lrb_build_bvh from host boxes, lrb_bvh_upload of the array).
	{
		const std::string builderType = ctx->GetConfig().Get(Property("accelerator.bvh.builder.type")("EMBREE_BINNED_SAH")).Get<std::string>();
		const bool gpuBuilder = builderType == "EMBREE_MORTON" || builderType == "B200_PLOC";
		if (gpuBuilder && allowResidentScene && totalTriangleCount > 1 &&
				ctx->GetConfig().Get(Property("accelerator.b200.resident")(true)).Get<bool>()) {
			const double tb = WallClockTime();
			if (BuildB200SceneOnDevice(params, builderType == "B200_PLOC" ? 1u : 0u, meshes, &bvhTree, &nNodes, &residentScene, &residentOrdinal)) {
				LR_LOG(ctx, "BVH builder: " << builderType << " (scene resident on CUDA device " << residentOrdinal << ")");
				LR_LOG(ctx, "BVH build hierarchy time: " << int((WallClockTime() - tb) * 1000) << "ms");
				LR_LOG(ctx, "Total BVH memory usage: " << nNodes * sizeof(ocl::BVHArrayNode) / 1024 << "Kbytes");
				initialized = true;
				return;
			}
		}
	}

	const double t0 = WallClockTime();
	// one build primitive per triangle: its box, grown by MachineEpsilon so that rays grazing the
	// box still reach the triangle test
	std::vector<BVHTreeNode> prims(totalTriangleCount);
	std::vector<BVHTreeNode *> list(totalTriangleCount, nullptr);
	size_t base = 0;
	for (size_t m = 0; m < meshes.size(); ++m) {
		const Mesh *mesh = meshes[m];
		const Triangle *tris = mesh->GetTriangles();
		const u_int count = mesh->GetTotalTriangleCount();
		for (u_int i = 0; i < count; ++i) {
			BVHTreeNode &n = prims[base + i];
			n.bbox = Union(BBox(mesh->GetVertex(Transform::TRANS_IDENTITY, tris[i].v[0]),
					mesh->GetVertex(Transform::TRANS_IDENTITY, tris[i].v[1])),
					mesh->GetVertex(Transform::TRANS_IDENTITY, tris[i].v[2]));
			n.bbox.Expand(MachineEpsilon::E(n.bbox));
			n.triangleLeaf.meshIndex = (u_int)m;
			n.triangleLeaf.triangleIndex = i;
			n.leftChild = nullptr;
			n.rightSibling = nullptr;
			list[base + i] = &n;
		}
		base += count;
	}
	LR_LOG(ctx, "BVH Dataset preprocessing time: " << int((WallClockTime() - t0) * 1000) << "ms");

	const double t1 = WallClockTime();
	bvhTree = RunBuilder(ctx, "BVH", params, &nNodes, &meshes, list);
	LR_LOG(ctx, "BVH build hierarchy time: " << int((WallClockTime() - t1) * 1000) << "ms");
	LR_LOG(ctx, "Total BVH memory usage: " << nNodes * sizeof(ocl::BVHArrayNode) / 1024 << "Kbytes");
	initialized = true;
}

bool BVHAccel::Intersect(const Ray *, RayHit *) const {
	throw std::runtime_error("BVHAccel::Intersect(): the B200 build has no CPU intersection path; "
			"use CUDAIntersectionDevice::TraceRay / EnqueueTraceRayBuffer");
}

bool BVHAccel::HasNativeSupport(const IntersectionDevice &) const { return false; }
bool BVHAccel::HasHWSupport(const IntersectionDevice &device) const { return device.HasHWSupport(); }

//------------------------------------------------------------------------------
// MBVHAccel
//------------------------------------------------------------------------------

MBVHAccel::MBVHAccel(const Context *context) : nRootNodes(0), bvhRootTree(nullptr), ctx(context), initialized(false) {
	params = BVHAccel::ToBVHParams(ctx->GetConfig());
}

MBVHAccel::~MBVHAccel() {
	for (size_t i = 0; i < uniqueLeafs.size(); ++i)
		delete uniqueLeafs[i];
	delete[] bvhRootTree;
}

void MBVHAccel::Init(const std::deque<const Mesh *> &ms, const u_longlong, const u_longlong totalTriangleCount) {
	if (totalTriangleCount == 0) {
		LR_LOG(ctx, "Empty MBVH");
		nRootNodes = 0;
		bvhRootTree = nullptr;
		initialized = true;
		return;
	}
	meshes = ms;
	const double t0 = WallClockTime();
	const u_int nLeafs = (u_int)meshes.size();
	LR_LOG(ctx, "Building Multilevel Bounding Volume Hierarchy: " << nLeafs << " leafs");

	// One leaf BVH per distinct base mesh of the instances / motion meshes; every plain mesh gets
	// its own leaf BVH (even if the same TriangleMesh was added twice).
	std::map<const Mesh *, u_int> leafOfBase;
	bvhLeafs.resize(nLeafs);
	bvhLeafsList.assign(nLeafs, nullptr);
	for (u_int i = 0; i < nLeafs; ++i) {
		const Mesh *mesh = meshes[i];
		const Mesh *base = nullptr;
		u_int transformIndex = NULL_INDEX, motionIndex = NULL_INDEX;
		bool share = true;
		switch (mesh->GetType()) {
			case TYPE_TRIANGLE:
			case TYPE_EXT_TRIANGLE:
				base = mesh;
				share = false;
				break;
			case TYPE_TRIANGLE_INSTANCE:
			case TYPE_EXT_TRIANGLE_INSTANCE: {
				const InstanceTriangleMesh *itm = dynamic_cast<const InstanceTriangleMesh *>(mesh);
				base = itm->GetTriangleMesh();
				transformIndex = (u_int)uniqueLeafsTransform.size();
				uniqueLeafsTransform.push_back(&itm->GetTransformation());
				break;
			}
			case TYPE_TRIANGLE_MOTION:
			case TYPE_EXT_TRIANGLE_MOTION: {
				const MotionTriangleMesh *mtm = dynamic_cast<const MotionTriangleMesh *>(mesh);
				base = mtm->GetTriangleMesh();
				motionIndex = (u_int)uniqueLeafsMotionSystem.size();
				uniqueLeafsMotionSystem.push_back(&mtm->GetMotionSystem());
				break;
			}
			default:
				throw std::runtime_error("Unknown Mesh type in MBVHAccel::Init(): " + std::to_string(mesh->GetType()));
		}

		u_int leafIndex;
		std::map<const Mesh *, u_int>::iterator it = leafOfBase.find(base);
		if (!share || it == leafOfBase.end()) {
			BVHAccel *leaf = new BVHAccel(ctx);
			leaf->allowResidentScene = false;       // leaf trees travel as arrays (lrb_mbvh_upload)
			const std::deque<const Mesh *> one(1, base);
			leaf->Init(one, base->GetTotalVertexCount(), base->GetTotalTriangleCount());
			leafIndex = (u_int)uniqueLeafs.size();
			leafOfBase[base] = leafIndex;
			uniqueLeafs.push_back(leaf);
		} else
			leafIndex = it->second;

		// root primitive: the mesh's world-space box, epsilon-expanded
		BVHTreeNode &l = bvhLeafs[i];
		l.bbox = mesh->GetBBox();
		l.bbox.Expand(MachineEpsilon::E(l.bbox));
		l.bvhLeaf.leafIndex = leafIndex;
		l.bvhLeaf.transformIndex = transformIndex;
		l.bvhLeaf.motionIndex = motionIndex;
		l.bvhLeaf.meshOffsetIndex = i;
		l.bvhLeaf.isMotionMesh = (motionIndex != NULL_INDEX);
		l.leftChild = nullptr;
		l.rightSibling = nullptr;
		bvhLeafsList[i] = &l;
	}

	LR_LOG(ctx, "Building Multilevel Bounding Volume Hierarchy root tree");
	UpdateRootBVH();
	LR_LOG(ctx, "MBVH build time: " << int((WallClockTime() - t0) * 1000) << "ms");
	initialized = true;
}

void MBVHAccel::UpdateRootBVH() {
	delete[] bvhRootTree;
	bvhRootTree = nullptr;
	bvhRootTree = RunBuilder(ctx, "MBVH root tree", params, &nRootNodes, nullptr, bvhLeafsList);
}

// Scene edit: instance transforms changed in place (the accelerator holds pointers to them).
// Only the root tree is rebuilt; as in the reference the refreshed boxes are NOT re-expanded.
void MBVHAccel::Update() {
	for (size_t i = 0; i < meshes.size(); ++i)
		bvhLeafs[i].bbox = meshes[i]->GetBBox();
	UpdateRootBVH();
}

bool MBVHAccel::Intersect(const Ray *, RayHit *) const {
	throw std::runtime_error("MBVHAccel::Intersect(): the B200 build has no CPU intersection path; "
			"use CUDAIntersectionDevice::TraceRay / EnqueueTraceRayBuffer");
}

bool MBVHAccel::HasNativeSupport(const IntersectionDevice &) const { return false; }
bool MBVHAccel::HasHWSupport(const IntersectionDevice &device) const { return device.HasHWSupport(); }

}   // namespace luxrays
