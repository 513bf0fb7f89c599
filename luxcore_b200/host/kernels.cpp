// kernels.cpp -- BVHKernel / MBVHKernel: the HardwareIntersectionKernel objects the accelerators
// hand to a hardware device.  They replace src/luxrays/accelerators/bvhaccelhw.cpp:38-281 and
// mbvhaccelhw.cpp:41-512.  Where the reference pages vertices/nodes, rewrites indices, stitches
// OpenCL-C sources and compiles them with NVRTC, these classes pass the accelerator's own arrays
// to the C ABI (include/luxrays_b200.h), which re-lays them out for sm_100a and owns the kernels.
#include "luxrays_b200.h"

#include "luxrays/accelerators/bvhaccel.h"
#include "luxrays/accelerators/mbvhaccel.h"
#include "luxrays/core/context.h"
#include "luxrays/devices/cudaintersectiondevice.h"

namespace luxrays {

void GatherSceneVertices(const std::deque<const Mesh *> &meshes, std::vector<float> &xyz, std::vector<uint32_t> &offsets);    // bvhbuild.cpp

static void Check(const int rc, const char *what) {
	if (rc != LRB_OK)
		throw std::runtime_error(std::string(what) + ": " + lrb_last_error_string());
}

static lrb_device *NativeOf(HardwareIntersectionDevice &dev) {
	CUDADevice *cd = dynamic_cast<CUDADevice *>(&dev);
	if (!cd || !cd->GetNativeHandle())
		throw std::runtime_error("BVH/MBVH hardware kernels need a started CUDADevice (B200 C-ABI device)");
	return cd->GetNativeHandle();
}

static void *DevPtr(HardwareDeviceBuffer *b, const char *what) {
	CUDADeviceBuffer *cb = dynamic_cast<CUDADeviceBuffer *>(b);
	if (!cb || cb->IsNull())
		throw std::runtime_error(std::string("Null or foreign buffer passed as ") + what);
	return cb->GetDevicePointer();
}

// lets tools reach the C-ABI scene behind a kernel object (statistics, host-buffer pipeline)
class B200SceneOwner {
public:
	virtual ~B200SceneOwner() { }
	virtual lrb_scene *NativeScene() const = 0;
};

lrb_scene *NativeSceneOf(HardwareIntersectionKernel *k) {
	B200SceneOwner *o = dynamic_cast<B200SceneOwner *>(k);
	return o ? o->NativeScene() : nullptr;
}

//------------------------------------------------------------------------------
// BVHKernel
//------------------------------------------------------------------------------

class BVHKernel : public HardwareIntersectionKernel, public B200SceneOwner {
public:
	BVHKernel(HardwareIntersectionDevice &dev, const BVHAccel &bvh) : HardwareIntersectionKernel(dev), scene(nullptr) {
		lrb_device *nd = NativeOf(dev);
		// A GPU-built accelerator is already laid out on its device (BVHAccel::Init, lrb_bvh_build_scene): the kernel of that
		// CUDA device takes the scene over; any other device gets the array like a host-built one.
		// (One DataSet may serve several devices, each started from its own host thread -- the reference's model: exactly
		// one kernel takes the scene, atomically.)
		if (bvh.residentScene) {
			lrb_device_props props;
			if (lrb_device_get_props(nd, &props) == LRB_OK && props.cuda_ordinal == bvh.residentOrdinal) {
				void *taken = __atomic_exchange_n(&bvh.residentScene, (void *)nullptr, __ATOMIC_ACQ_REL);
				if (taken) {
					if (lrb_scene_adopt(nd, static_cast<lrb_scene *>(taken)) == LRB_OK) {
						scene = static_cast<lrb_scene *>(taken);
						return;
					}
					lrb_scene_free(static_cast<lrb_scene *>(taken));     // could not be handed over: upload the array instead
				}
			}
		}
		std::vector<float> xyz;
		std::vector<uint32_t> offsets;
		GatherSceneVertices(bvh.meshes, xyz, offsets);
		Check(lrb_bvh_upload(nd, reinterpret_cast<const lrb_bvh_node *>(bvh.bvhTree), bvh.nNodes,
				xyz.empty() ? nullptr : xyz.data(), xyz.size() / 3,
				offsets.empty() ? nullptr : offsets.data(), (uint32_t)offsets.size(), &scene), "BVHKernel upload");
	}
	virtual ~BVHKernel() { lrb_scene_free(scene); }

	virtual void Update(const DataSet *) { throw std::runtime_error("BVHAccel does not support Update()"); }
	virtual lrb_scene *NativeScene() const { return scene; }

	virtual void EnqueueTraceRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount) {
		if (rayCount == 0)
			return;
		Check(lrb_trace(scene, DevPtr(rayBuff, "ray buffer"), DevPtr(rayHitBuff, "ray hit buffer"), rayCount), "BVHKernel trace");
	}

	lrb_scene *scene;
};

HardwareIntersectionKernel *BVHAccel::NewHardwareIntersectionKernel(HardwareIntersectionDevice &device) const {
	return new BVHKernel(device, *this);
}

//------------------------------------------------------------------------------
// MBVHKernel
//------------------------------------------------------------------------------

class MBVHKernel : public HardwareIntersectionKernel, public B200SceneOwner {
public:
	MBVHKernel(HardwareIntersectionDevice &dev, const MBVHAccel &acc) : HardwareIntersectionKernel(dev), mbvh(acc), scene(nullptr) {
		lrb_device *nd = NativeOf(dev);
		const size_t nLeaves = mbvh.uniqueLeafs.size();
		std::vector<const lrb_bvh_node *> leafNodes(nLeaves);
		std::vector<uint32_t> leafNodeCount(nLeaves), leafVertCount(nLeaves);
		std::vector<const float *> leafVerts(nLeaves);
		for (size_t i = 0; i < nLeaves; ++i) {
			const BVHAccel *leaf = mbvh.uniqueLeafs[i];
			leafNodes[i] = reinterpret_cast<const lrb_bvh_node *>(leaf->bvhTree);
			leafNodeCount[i] = leaf->nNodes;
			// leaf BVHs are built over exactly one plain mesh; its LOCAL vertices are used
			leafVerts[i] = reinterpret_cast<const float *>(leaf->meshes[0]->GetVertices());
			leafVertCount[i] = leaf->meshes[0]->GetTotalVertexCount();
		}
		std::vector<float> minv;
		GatherInverseMatrices(minv);

		// motion systems -> ocl::MotionSystem index table + flat InterpolatedTransform array; the
		// inverse ranges are not needed for traversal (mbvhaccelhw.cpp:157-175)
		std::vector<lrb_motion_system> systems;
		std::vector<InterpolatedTransform> interps;
		for (size_t i = 0; i < mbvh.uniqueLeafsMotionSystem.size(); ++i) {
			const MotionSystem *ms = mbvh.uniqueLeafsMotionSystem[i];
			lrb_motion_system s;
			s.interpolatedTransformFirstIndex = (uint32_t)interps.size();
			interps.insert(interps.end(), ms->interpolatedTransforms.begin(), ms->interpolatedTransforms.end());
			s.interpolatedTransformLastIndex = (uint32_t)interps.size() - 1;
			s.interpolatedInverseTransformFirstIndex = NULL_INDEX;
			s.interpolatedInverseTransformLastIndex = NULL_INDEX;
			systems.push_back(s);
		}

		lrb_mbvh_desc d;
		memset(&d, 0, sizeof(d));
		d.root_nodes = reinterpret_cast<const lrb_bvh_node *>(mbvh.bvhRootTree);
		d.n_root_nodes = mbvh.nRootNodes;
		d.n_leaves = (uint32_t)nLeaves;
		d.leaf_nodes = leafNodes.data();
		d.leaf_n_nodes = leafNodeCount.data();
		d.leaf_vertices = leafVerts.data();
		d.leaf_n_vertices = leafVertCount.data();
		d.transforms_minv = minv.empty() ? nullptr : minv.data();
		d.n_transforms = (uint32_t)(minv.size() / 16);
		d.motion_systems = systems.empty() ? nullptr : systems.data();
		d.n_motion_systems = (uint32_t)systems.size();
		d.interpolated_transforms = interps.empty() ? nullptr : interps.data();
		d.n_interpolated_transforms = (uint32_t)interps.size();
		Check(lrb_mbvh_upload(nd, &d, &scene), "MBVHKernel upload");
	}
	virtual ~MBVHKernel() { lrb_scene_free(scene); }

	virtual lrb_scene *NativeScene() const { return scene; }

	// after MBVHAccel::Update(): new root tree, refreshed inverse instance matrices
	virtual void Update(const DataSet *) {
		std::vector<float> minv;
		GatherInverseMatrices(minv);
		Check(lrb_mbvh_update(scene, reinterpret_cast<const lrb_bvh_node *>(mbvh.bvhRootTree), mbvh.nRootNodes,
				minv.empty() ? nullptr : minv.data(), (uint32_t)(minv.size() / 16)), "MBVHKernel update");
	}

	virtual void EnqueueTraceRayBuffer(HardwareDeviceBuffer *rayBuff, HardwareDeviceBuffer *rayHitBuff, const unsigned int rayCount) {
		if (rayCount == 0)
			return;
		Check(lrb_trace(scene, DevPtr(rayBuff, "ray buffer"), DevPtr(rayHitBuff, "ray hit buffer"), rayCount), "MBVHKernel trace");
	}

private:
	void GatherInverseMatrices(std::vector<float> &minv) const {
		minv.clear();
		for (size_t i = 0; i < mbvh.uniqueLeafsTransform.size(); ++i) {
			const float *m = &mbvh.uniqueLeafsTransform[i]->mInv.m[0][0];
			minv.insert(minv.end(), m, m + 16);
		}
	}

	const MBVHAccel &mbvh;
	lrb_scene *scene;
};

HardwareIntersectionKernel *MBVHAccel::NewHardwareIntersectionKernel(HardwareIntersectionDevice &device) const {
	return new MBVHKernel(device, *this);
}

}   // namespace luxrays
