// shim.cpp -- flat C entry points over the C++ host layer, so that Python (tests, bench.py) can
// drive the exact call sequence a LuxCore application uses:
//   Context -> DataSet::Add/Preprocess -> Context::SetDataSet/Start -> AllocBufferRW ->
//   EnqueueTraceRayBuffer -> EnqueueReadBuffer -> FinishQueue -> Stop.
// Nothing here adds functionality; every function forwards to the luxrays:: classes.
#include <memory>
#include <string>

#include "luxrays_b200.h"

#include "luxrays/accelerators/bvhaccel.h"
#include "luxrays/accelerators/mbvhaccel.h"
#include "luxrays/core/context.h"
#include "luxrays/devices/cudaintersectiondevice.h"

using namespace luxrays;

namespace {

thread_local std::string g_err;

struct Session {
	std::unique_ptr<Context> ctx;
	std::unique_ptr<DataSet> dataSet;
	std::vector<TriangleMesh *> shapes;
	std::vector<Mesh *> meshes;     // dataset entries that are not plain shapes (instances, motion)
	std::vector<Mesh *> entries;    // dataset order
	const Accelerator *accel;
	CUDAIntersectionDevice *device;
	HardwareDeviceBuffer *rays, *hits;
	std::string log;

	Session() : accel(nullptr), device(nullptr), rays(nullptr), hits(nullptr) { }
	~Session() {
		if (device) {
			device->FreeBuffer(&rays);
			device->FreeBuffer(&hits);
		}
		ctx.reset();        // stops + deletes devices (kernels free their scenes first)
		dataSet.reset();
		for (size_t i = 0; i < meshes.size(); ++i) delete meshes[i];
		for (size_t i = 0; i < shapes.size(); ++i) { shapes[i]->Delete(); delete shapes[i]; }
	}
};

}   // namespace

#define LRH_TRY try {
#define LRH_CATCH } catch (const std::exception &e) { g_err = e.what(); return -1; } catch (...) { g_err = "unknown exception"; return -1; } return 0;

extern "C" {

#define LRH_API __attribute__((visibility("default")))

LRH_API const char *lrh_last_error() { return g_err.c_str(); }

LRH_API void *lrh_create(const char *configText) {
	try {
		Session *s = new Session();
		Properties cfg;
		cfg << Property("context.verbose")(false);
		if (configText)
			cfg.SetFromString(configText);
		s->ctx.reset(new Context(nullptr, cfg));
		s->dataSet.reset(new DataSet(s->ctx.get()));
		return s;
	} catch (const std::exception &e) {
		g_err = e.what();
		return nullptr;
	}
}

LRH_API void lrh_destroy(void *sp) { delete (Session *)sp; }

LRH_API int lrh_device_description_count(void *sp) { return (int)((Session *)sp)->ctx->GetAvailableDeviceDescriptions().size(); }

LRH_API int lrh_add_shape(void *sp, const float *xyz, uint32_t nVerts, const uint32_t *tris, uint32_t nTris) {
	Session *s = (Session *)sp;
	try {
		Point *v = TriangleMesh::AllocVerticesBuffer(nVerts);
		memcpy(reinterpret_cast<void *>(v), xyz, sizeof(float) * 3 * (size_t)nVerts);
		Triangle *t = TriangleMesh::AllocTrianglesBuffer(nTris);
		memcpy(reinterpret_cast<void *>(t), tris, sizeof(uint32_t) * 3 * (size_t)nTris);
		s->shapes.push_back(new TriangleMesh(nVerts, nTris, v, t));
		return (int)s->shapes.size() - 1;
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
}

static TriangleMesh *ShapeOf(Session *s, int shape) {
	if (shape < 0 || shape >= (int)s->shapes.size())
		throw std::runtime_error("bad shape index");
	return s->shapes[shape];
}

LRH_API int lrh_add_plain(void *sp, int shape) {
	Session *s = (Session *)sp;
	LRH_TRY
	TriangleMesh *m = ShapeOf(s, shape);
	s->entries.push_back(m);
	s->dataSet->Add(m);
	LRH_CATCH
}

// m16: row-major local->world
LRH_API int lrh_add_instance(void *sp, int shape, const float *m16) {
	Session *s = (Session *)sp;
	LRH_TRY
	InstanceTriangleMesh *m = new InstanceTriangleMesh(ShapeOf(s, shape), Transform(Matrix4x4(m16)));
	s->meshes.push_back(m);
	s->entries.push_back(m);
	s->dataSet->Add(m);
	LRH_CATCH
}

// matrices as the MotionSystem stores them (world->local for scene objects), row-major
LRH_API int lrh_add_motion(void *sp, int shape, uint32_t nKeys, const float *times, const float *m16s) {
	Session *s = (Session *)sp;
	LRH_TRY
	std::vector<float> t(times, times + nKeys);
	std::vector<Transform> x;
	for (uint32_t i = 0; i < nKeys; ++i)
		x.push_back(Transform(Matrix4x4(m16s + 16 * i)));
	MotionTriangleMesh *m = new MotionTriangleMesh(ShapeOf(s, shape), MotionSystem(t, x));
	s->meshes.push_back(m);
	s->entries.push_back(m);
	s->dataSet->Add(m);
	LRH_CATCH
}

LRH_API int lrh_preprocess(void *sp) {
	Session *s = (Session *)sp;
	LRH_TRY
	s->dataSet->Preprocess();
	LRH_CATCH
}

// Host-only accelerator build (no GPU needed).  type: "AUTO" | "BVH" | "MBVH".
// Returns 1 for BVH, 2 for MBVH, -1 on error.
LRH_API int lrh_build_accelerator(void *sp, const char *type) {
	Session *s = (Session *)sp;
	try {
		AcceleratorType t = Accelerator::String2AcceleratorType(type ? type : "AUTO");
		if (t == ACCEL_AUTO)
			t = (s->dataSet->RequiresInstanceSupport() || s->dataSet->RequiresMotionBlurSupport()) ? ACCEL_MBVH : ACCEL_BVH;
		s->dataSet->SetAcceleratorType(t);
		s->accel = s->dataSet->GetAccelerator(t);
		return t == ACCEL_BVH ? 1 : 2;
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
}

// ---- read-only views of what the builder produced (for parity tests against the oracle) ----

LRH_API uint32_t lrh_bvh_node_count(void *sp) {
	const BVHAccel *b = dynamic_cast<const BVHAccel *>(((Session *)sp)->accel);
	return b ? b->GetNodeCount() : 0;
}
LRH_API const void *lrh_bvh_nodes(void *sp) {
	const BVHAccel *b = dynamic_cast<const BVHAccel *>(((Session *)sp)->accel);
	return b ? b->GetNodes() : nullptr;
}
LRH_API uint32_t lrh_mbvh_root_node_count(void *sp) {
	const MBVHAccel *m = dynamic_cast<const MBVHAccel *>(((Session *)sp)->accel);
	return m ? m->GetRootNodeCount() : 0;
}
LRH_API const void *lrh_mbvh_root_nodes(void *sp) {
	const MBVHAccel *m = dynamic_cast<const MBVHAccel *>(((Session *)sp)->accel);
	return m ? m->GetRootNodes() : nullptr;
}
LRH_API uint32_t lrh_mbvh_leaf_count(void *sp) {
	const MBVHAccel *m = dynamic_cast<const MBVHAccel *>(((Session *)sp)->accel);
	return m ? (uint32_t)m->GetUniqueLeafCount() : 0;
}
LRH_API uint32_t lrh_mbvh_leaf_node_count(void *sp, uint32_t i) {
	const MBVHAccel *m = dynamic_cast<const MBVHAccel *>(((Session *)sp)->accel);
	return m ? m->GetUniqueLeaf(i)->GetNodeCount() : 0;
}
LRH_API const void *lrh_mbvh_leaf_nodes(void *sp, uint32_t i) {
	const MBVHAccel *m = dynamic_cast<const MBVHAccel *>(((Session *)sp)->accel);
	return m ? m->GetUniqueLeaf(i)->GetNodes() : nullptr;
}
LRH_API int lrh_mesh_bbox(void *sp, int mesh, float *out6) {
	Session *s = (Session *)sp;
	LRH_TRY
	const BBox b = s->entries.at(mesh)->GetBBox();
	out6[0] = b.pMin.x; out6[1] = b.pMin.y; out6[2] = b.pMin.z; out6[3] = b.pMax.x; out6[4] = b.pMax.y; out6[5] = b.pMax.z;
	LRH_CATCH
}

// DataSet::GetBBox / GetBSphere (dataset.h:60-61): out10 = min xyz, max xyz, centre xyz, radius
LRH_API int lrh_dataset_bounds(void *sp, float *out10) {
	Session *s = (Session *)sp;
	LRH_TRY
	if (!s->dataSet->IsPreprocessed())
		s->dataSet->Preprocess();
	const BBox &b = s->dataSet->GetBBox();
	const BSphere &bs = s->dataSet->GetBSphere();
	out10[0] = b.pMin.x; out10[1] = b.pMin.y; out10[2] = b.pMin.z; out10[3] = b.pMax.x; out10[4] = b.pMax.y; out10[5] = b.pMax.z;
	out10[6] = bs.center.x; out10[7] = bs.center.y; out10[8] = bs.center.z; out10[9] = bs.rad;
	LRH_CATCH
}

// ---- device lifecycle ----

// Context::AddIntersectionDevices(descs[deviceIndex]) + SetDataSet + Start
LRH_API int lrh_start(void *sp, int deviceIndex) {
	Session *s = (Session *)sp;
	LRH_TRY
	std::vector<DeviceDescription *> descs = s->ctx->GetAvailableDeviceDescriptions();
	DeviceDescription::Filter(DEVICE_TYPE_CUDA_GPU, descs);
	if (descs.empty())
		throw std::runtime_error("no CUDA device available (the B200 library has no CPU fallback)");
	if (deviceIndex < 0 || deviceIndex >= (int)descs.size())
		throw std::runtime_error("device index out of range");
	std::vector<DeviceDescription *> one(1, descs[deviceIndex]);
	std::vector<IntersectionDevice *> devs = s->ctx->AddIntersectionDevices(one);
	s->device = dynamic_cast<CUDAIntersectionDevice *>(devs[0]);
	if (!s->dataSet->IsPreprocessed())
		s->dataSet->Preprocess();
	s->ctx->SetDataSet(s->dataSet.get());
	s->accel = s->device->GetAccelerator();
	s->ctx->Start();
	LRH_CATCH
}

LRH_API int lrh_stop(void *sp) {
	Session *s = (Session *)sp;
	LRH_TRY
	if (s->device) {
		s->device->FreeBuffer(&s->rays);
		s->device->FreeBuffer(&s->hits);
	}
	s->ctx->Stop();
	LRH_CATCH
}

LRH_API void *lrh_native_device(void *sp) {
	Session *s = (Session *)sp;
	return s->device ? s->device->GetNativeHandle() : nullptr;
}

LRH_API void *lrh_native_scene(void *sp) {
	Session *s = (Session *)sp;
	return s->device ? s->device->GetNativeScene() : nullptr;
}

LRH_API int lrh_accelerator_type(void *sp) {
	Session *s = (Session *)sp;
	return s->accel ? (int)s->accel->GetType() : -1;
}

// The SURVEY.md 3.5 sequence with HOST buffers:
//   AllocBufferRW(rays, hostRays) / AllocBufferRW(hits) / EnqueueTraceRayBuffer /
//   EnqueueReadBuffer(non-blocking) / FinishQueue
LRH_API int lrh_trace_host(void *sp, const void *hostRays, void *hostHits, uint32_t n, int preloadHits) {
	Session *s = (Session *)sp;
	LRH_TRY
	if (!s->device)
		throw std::runtime_error("session not started");
	if (n == 0)
		return 0;
	s->device->PushThreadCurrentDevice();
	s->device->AllocBufferRW(&s->rays, const_cast<void *>(hostRays), (size_t)n * sizeof(Ray), "Ray");
	s->device->AllocBufferRW(&s->hits, preloadHits ? hostHits : nullptr, (size_t)n * sizeof(RayHit), "RayHit");
	s->device->EnqueueTraceRayBuffer(s->rays, s->hits, n);
	s->device->EnqueueReadBuffer(s->hits, false, (size_t)n * sizeof(RayHit), hostHits);
	s->device->FinishQueue();
	s->device->PopThreadCurrentDevice();
	LRH_CATCH
}

// EnqueueTraceRayBuffer on buffers that already live in HBM (allocated by the caller on the same
// device, e.g. torch tensors): wrapped as HardwareDeviceBuffer views, not copied.
LRH_API int lrh_trace_device(void *sp, void *raysDev, void *hitsDev, uint32_t n) {
	Session *s = (Session *)sp;
	LRH_TRY
	if (!s->device)
		throw std::runtime_error("session not started");
	HardwareDeviceBuffer *r = s->device->AdoptBuffer(raysDev, (size_t)n * sizeof(Ray));
	HardwareDeviceBuffer *h = s->device->AdoptBuffer(hitsDev, (size_t)n * sizeof(RayHit));
	try {
		s->device->EnqueueTraceRayBuffer(r, h, n);
	} catch (...) {
		delete r; delete h;
		throw;
	}
	delete r;
	delete h;
	LRH_CATCH
}

// shadow rays on caller-owned device memory (extension, see cudaintersectiondevice.h)
LRH_API int lrh_trace_device_shadow(void *sp, void *raysDev, void *hitsDev, uint32_t n) {
	Session *s = (Session *)sp;
	LRH_TRY
	if (!s->device)
		throw std::runtime_error("session not started");
	HardwareDeviceBuffer *r = s->device->AdoptBuffer(raysDev, (size_t)n * sizeof(Ray));
	HardwareDeviceBuffer *h = s->device->AdoptBuffer(hitsDev, (size_t)n * sizeof(RayHit));
	try {
		s->device->EnqueueTraceShadowRayBuffer(r, h, n);
	} catch (...) {
		delete r; delete h;
		throw;
	}
	delete r;
	delete h;
	LRH_CATCH
}

LRH_API int lrh_finish(void *sp) {
	Session *s = (Session *)sp;
	LRH_TRY
	if (s->device) s->device->FinishQueue();
	LRH_CATCH
}

// IntersectionDevice::TraceRay (one ray, traced on the GPU)
LRH_API int lrh_trace_ray(void *sp, const void *ray, void *hit) {
	Session *s = (Session *)sp;
	try {
		if (!s->device)
			throw std::runtime_error("session not started");
		return s->device->TraceRay((const Ray *)ray, (RayHit *)hit) ? 1 : 0;
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
}

// scene edit: InstanceTriangleMesh::SetTransformation + Context::UpdateDataSet
LRH_API int lrh_set_instance_transform(void *sp, int mesh, const float *m16) {
	Session *s = (Session *)sp;
	LRH_TRY
	InstanceTriangleMesh *m = dynamic_cast<InstanceTriangleMesh *>(s->entries.at(mesh));
	if (!m)
		throw std::runtime_error("mesh is not an instance");
	m->SetTransformation(Transform(Matrix4x4(m16)));
	LRH_CATCH
}

LRH_API int lrh_update(void *sp) {
	Session *s = (Session *)sp;
	LRH_TRY
	s->ctx->UpdateDataSet();
	LRH_CATCH
}

LRH_API double lrh_stats_total_rays(void *sp) {
	Session *s = (Session *)sp;
	return s->device ? s->device->GetTotalRaysCount() : 0.0;
}

LRH_API uint64_t lrh_used_memory(void *sp) {
	Session *s = (Session *)sp;
	return s->device ? (uint64_t)s->device->GetUsedMemory() : 0;
}

// accelerator.cpp string helpers + MachineEpsilon + Matrix inverse, exposed for unit tests
LRH_API float lrh_machine_epsilon(float v) { return MachineEpsilon::E(v); }
LRH_API int lrh_matrix_inverse(const float *m16, float *out16) {
	LRH_TRY
	const Matrix4x4 r = Matrix4x4(m16).Inverse();
	memcpy(out16, r.m, 64);
	LRH_CATCH
}

}   // extern "C"
