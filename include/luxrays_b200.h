/* luxrays_b200.h -- C ABI of the B200 (sm_100a) closest-hit intersection device.
 *
 * This is the drop-in boundary: a thin, torch-free, exception-free C
 * interface that replaces what the reference reaches through cuew/NVRTC inside
 *   - luxrays::CUDADevice          (src/luxrays/devices/cudadevice.cpp:205-211,407-540:
 *                                   Push/Pop current, Alloc/Free, Read/Write, Flush/Finish)
 *   - luxrays::BVHKernel           (src/luxrays/accelerators/bvhaccelhw.cpp:38-268: vertex + node
 *                                   upload, kernel launch)
 *   - luxrays::MBVHKernel          (src/luxrays/accelerators/mbvhaccelhw.cpp:41-306 upload,
 *                                   :308-466 UpdateBVHNodes/Update, :468-507 launch)
 *   - bvh.cl / mbvh.cl             (include/luxrays/accelerators/bvh.cl:228-260, mbvh.cl:351-383:
 *                                   Accelerator_Intersect_RayBuffer)
 * and, one step outwards, what it reaches through Embree and SLG's own loops:
 *   - BuildEmbreeBVHMorton / BinnedSAH (src/luxrays/core/bvh/bvhembreebuild.cpp:218-336) and, for trees built on
 *     the GPU, BVHAccel::Init itself (src/luxrays/accelerators/bvhaccel.cpp:72-168): lrb_build_bvh,
 *     lrb_bvh_build_scene
 *   - Scene::Intersect's pass-through loop and shadow rays (src/slg/scene/scene.cpp:556-690), the masked-ray
 *     contract of the path tracer (pathoclbase_kernels_micro.cl:34-106,1029): lrb_trace_anyhit, lrb_compact_rays,
 *     lrb_trace_indexed, lrb_advance_rays, lrb_trace_passthrough
 *   - the film merge of PathOCLRenderEngine (src/slg/engines/pathocl/pathocl.cpp:184-201): lrb_film_reduce
 *
 * The C++ host layer in luxcore_b200/host (namespace luxrays, same class names as the reference)
 * sits on top of this ABI; INTEGRATION.md shows the binding a LuxCore maintainer would add.
 *
 * Conventions
 *   - every function returns an int status: 0 = LRB_OK, anything else is an error whose text is
 *     available (per calling thread) from lrb_last_error_string(); no C++ exception crosses the ABI;
 *   - all device work (copies, re-layout kernels, traces) is enqueued IN ORDER on the device's
 *     stream; lrb_sync() is the only implicit synchronisation point besides `blocking` copies --
 *     the same contract as the reference's single in-order queue (cudadevice.cpp:407-442);
 *   - the structs below are bit-identical to the reference wire types, so the reference's own
 *     arrays are handed over unchanged.
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with
 *     LRB_ERR_NO_DEVICE.
 */
#ifndef LUXRAYS_B200_H
#define LUXRAYS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRB_API __attribute__((visibility("default")))

/* status codes */
#define LRB_OK 0
#define LRB_ERR_INVALID 1      /* bad argument */
#define LRB_ERR_CUDA 2         /* a CUDA runtime call failed */
#define LRB_ERR_NO_DEVICE 3    /* no usable CUDA device */
#define LRB_ERR_OOM 4
#define LRB_ERR_INTERNAL 5

#define LRB_NULL_INDEX 0xffffffffu
#define LRB_RAY_FLAGS_NONE 0u
#define LRB_RAY_FLAGS_MASKED 1u   /* include/luxrays/core/geometry/ray_types.cl:21-22 */

/* luxrays::Ray / ocl::Ray -- include/luxrays/core/geometry/ray.h:35-88, ray_types.cl:24-31. 48 B. */
typedef struct {
	float o[3];
	float d[3];
	float mint, maxt, time;
	uint32_t flags;
	float pad[2];
} lrb_ray;

/* luxrays::RayHit -- ray.h:95-103, ray_types.cl:33-36. 20 B. Miss <=> meshIndex == LRB_NULL_INDEX. */
typedef struct {
	float t, b1, b2;
	uint32_t meshIndex, triangleIndex;
} lrb_rayhit;

/* ocl::BVHArrayNode -- include/luxrays/core/bvh/bvhbuild_types.cl:21-42. 32 B.
 * nodeData: bit 31 = leaf flag, bits 0-30 = skip index (bvhbuild.h:40-41). */
typedef struct {
	union {
		struct { float bboxMin[3], bboxMax[3]; } bvhNode;
		struct { uint32_t v[3], meshIndex, triangleIndex; } triangleLeaf;
		struct { uint32_t leafIndex, transformIndex, motionIndex, meshOffsetIndex; } bvhLeaf;
	};
	uint32_t nodeData;
	int32_t pad0;
} lrb_bvh_node;

/* ocl::MotionSystem -- include/luxrays/core/geometry/motionsystem_types.cl:49-55. 16 B.
 * Indices into the flat interpolated-transform array; the inverse range is ignored. */
typedef struct {
	uint32_t interpolatedTransformFirstIndex;
	uint32_t interpolatedTransformLastIndex;
	uint32_t interpolatedInverseTransformFirstIndex;
	uint32_t interpolatedInverseTransformLastIndex;
} lrb_motion_system;

/* Size of one ocl::InterpolatedTransform record (motionsystem_types.cl:21-47); the device reads
 * the reference layout directly from the host array and re-packs what traversal needs. */
#define LRB_INTERPOLATED_TRANSFORM_SIZE 576

typedef struct lrb_device lrb_device;
typedef struct lrb_scene lrb_scene;

typedef struct {
	int cuda_ordinal;
	int cc_major, cc_minor;
	int sm_count;
	int l2_bytes;
	uint64_t total_mem_bytes;
	char name[128];
} lrb_device_props;

/* What lrb_mbvh_upload consumes: exactly the arrays MBVHKernel walks (mbvhaccelhw.cpp:63-175,308-440)
 * BEFORE its index rewriting -- i.e. the accelerator's own, tree-relative arrays. */
typedef struct {
	const lrb_bvh_node *root_nodes;        /* MBVHAccel::bvhRootTree */
	uint32_t n_root_nodes;                 /* MBVHAccel::nRootNodes */
	uint32_t n_leaves;                     /* uniqueLeafs.size() */
	const lrb_bvh_node *const *leaf_nodes; /* uniqueLeafs[i]->bvhTree */
	const uint32_t *leaf_n_nodes;          /* uniqueLeafs[i]->nNodes */
	const float *const *leaf_vertices;     /* uniqueLeafs[i]->meshes[0]->GetVertices(): local xyz, 12-B stride */
	const uint32_t *leaf_n_vertices;
	const float *transforms_minv;          /* uniqueLeafsTransform[i]->mInv, 16 floats row-major each */
	uint32_t n_transforms;
	const lrb_motion_system *motion_systems;
	uint32_t n_motion_systems;
	const void *interpolated_transforms;   /* n * LRB_INTERPOLATED_TRANSFORM_SIZE bytes */
	uint32_t n_interpolated_transforms;
} lrb_mbvh_desc;

typedef struct {
	uint32_t n_ref_nodes;          /* BVHArrayNode count received */
	uint32_t n_wide_nodes;         /* 64-B quantized wide nodes after re-layout */
	uint32_t n_triangles;          /* 64-B pre-gathered triangle records */
	uint32_t n_instances;          /* 32-B instance records (MBVH) */
	uint32_t stack_need;           /* worst-case traversal stack entries */
	uint32_t two_level;            /* 0 = BVH, 1 = MBVH */
	uint64_t device_bytes;         /* bytes of device memory owned by the scene */
} lrb_scene_info;

typedef struct {
	uint64_t rays_traced;          /* rays submitted (masked included, like cudaintersectiondevice.cpp:85) */
	uint64_t trace_launches;       /* traversal kernel launches */
	uint64_t kernel_launches;      /* every kernel launched by this library (trace + re-layout) */
	uint64_t h2d_bytes, d2h_bytes;
	uint64_t device_bytes_in_use;
} lrb_counters;

/* per-batch traversal statistics from the instrumented kernel (lrb_trace_stats) */
typedef struct {
	uint64_t rays;                 /* non-masked rays */
	uint64_t wide_nodes;           /* 64-B wide-node fetches */
	uint64_t triangles;            /* 64-B triangle record fetches */
	uint64_t instances;            /* instance entries (32-B record + 64-B matrix) */
	uint64_t motion_samples;       /* motion leaves entered */
	uint64_t max_stack;            /* deepest stack seen */
} lrb_trace_stats_t;

/* ---- lifecycle ------------------------------------------------------------------------- */
LRB_API int lrb_device_count(int *count);
LRB_API int lrb_device_create(int cuda_ordinal, lrb_device **out);       /* replaces CUDADevice ctor + Start */
LRB_API int lrb_device_destroy(lrb_device *dev);
LRB_API int lrb_device_get_props(lrb_device *dev, lrb_device_props *out);
/* Adopt a caller-owned cudaStream_t (e.g. the application's render stream) as the in-order queue;
 * NULL restores the device's own stream. */
LRB_API int lrb_device_set_stream(lrb_device *dev, void *cuda_stream);
LRB_API int lrb_device_get_stream(lrb_device *dev, void **cuda_stream);
/* Tunables (strings): "kernel" = "persistent"|"simple", "blocks_per_sm", "smem_depth",
 * "refill_below", "tri_bias", "inst_bias", "host_chunk", "host_taper", "host_min_chunk", "sort_rays" (0 never | 1 always | 2 = default: scenes larger than L2),
 * "sort_bits", "sort_min_rays", "gather_stores", "gather_chunk_shift", "gather_defer", "wide_stores",
 * "prefetch" (L2 prefetch of pushed children: 0 never = default | 1 always | 2 scenes larger than L2),
 * "compact" (0 = default: masked rays are skipped inside the kernel | 1 = lrb_trace compacts the live rays first |
 * 2 = counts them first and compacts when fewer than 60 % are live),
 * "carveout" (preferred shared-memory carve-out of the trace kernels in percent, -1 = driver default),
 * "l2_persist" (L2 access-policy window over a one-level scene: 0 = default never | 1 always | 2 when it fits the set-aside),
 * "pipeline" (1 = default: the h2d / trace / d2h sequence below is pipelined chunk by chunk | 0 = every call as it is). */
LRB_API int lrb_device_set_option(lrb_device *dev, const char *key, const char *value);

/* ---- memory + queue (cudadevice.cpp:407-540) ------------------------------------------- */
LRB_API int lrb_alloc(lrb_device *dev, size_t bytes, void **devptr);
LRB_API int lrb_free(lrb_device *dev, void *devptr);
/* One in-order queue, as in the reference.  Behind that contract the reference-facing sequence
 *     AllocBufferRW(&rays, hostRays) -> EnqueueTraceRayBuffer(rays, hits, n) -> EnqueueReadBuffer(hits, hostHits) -> FinishQueue
 *   = lrb_h2d(rays_dev, ..., blocking = 0) -> lrb_trace(scene, rays_dev, hits_dev, n) -> lrb_d2h(hostHits, hits_dev, ...) -> lrb_sync
 * is pipelined: a non-blocking upload of >= 32 MB travels in chunks of "host_chunk" rays on a copy stream (the last
 * chunks halve down to "host_min_chunk" rays unless "host_taper" is 0: the pipeline's drain is one small piece), a trace
 * whose ray buffer is exactly that upload follows it chunk by chunk, and a read of exactly that trace's RayHit buffer
 * follows the trace chunk by chunk on a second copy stream (PCIe is full duplex).  Any other call first joins the
 * queue with whatever is pending, so results and ordering are those of the plain sequence; as with the reference's
 * cuMemcpyHtoDAsync (cudadevice.cpp:485) a pinned source must stay valid until the queue has been finished. */
LRB_API int lrb_h2d(lrb_device *dev, void *dst_dev, const void *src_host, size_t bytes, int blocking);
LRB_API int lrb_d2h(lrb_device *dev, void *dst_host, const void *src_dev, size_t bytes, int blocking);
LRB_API int lrb_flush(lrb_device *dev);
LRB_API int lrb_sync(lrb_device *dev);                                   /* FinishQueue */

/* ---- scenes ---------------------------------------------------------------------------- */
/* Single-level BVH (replaces the BVHKernel ctor).  `nodes` is BVHAccel::bvhTree with its original,
 * mesh-relative vertex indices; `xyz` the concatenation of every mesh's vertices
 * (Mesh::GetVertex(TRANS_IDENTITY, i), 12-B stride) and mesh_vertex_offsets[m] the first vertex of
 * mesh m in it (bvhaccelhw.cpp:68-92,126-145).  n_nodes == 0 is the empty DataSet: every ray misses. */
LRB_API int lrb_bvh_upload(lrb_device *dev, const lrb_bvh_node *nodes, uint32_t n_nodes,
		const float *xyz, uint64_t n_vertices, const uint32_t *mesh_vertex_offsets, uint32_t n_meshes,
		lrb_scene **out);
/* Two-level MBVH (replaces the MBVHKernel ctor). */
LRB_API int lrb_mbvh_upload(lrb_device *dev, const lrb_mbvh_desc *desc, lrb_scene **out);
/* MBVHKernel::Update (mbvhaccelhw.cpp:442-466): new root tree + new inverse instance matrices;
 * leaf trees, vertices and motion systems are kept. */
LRB_API int lrb_mbvh_update(lrb_scene *scene, const lrb_bvh_node *root_nodes, uint32_t n_root_nodes,
		const float *transforms_minv, uint32_t n_transforms);
LRB_API int lrb_scene_free(lrb_scene *scene);
LRB_API int lrb_scene_get_info(lrb_scene *scene, lrb_scene_info *out);

/* ---- trace ----------------------------------------------------------------------------- */
/* Accelerator_Intersect_RayBuffer: rays_dev = ray_count packed lrb_ray, hits_dev = ray_count packed
 * lrb_rayhit, both DEVICE pointers (from lrb_alloc or any CUDA allocation of this device, peer-mapped
 * memory included).  Asynchronous.  Rays with flags & LRB_RAY_FLAGS_MASKED are skipped and their
 * RayHit is left untouched (bvh.cl:242-244). */
LRB_API int lrb_trace(lrb_scene *scene, const void *rays_dev, void *hits_dev, uint32_t ray_count);
/* Shadow / visibility rays (any-hit).  Same buffers and masked-ray rule as lrb_trace; a ray stops at the FIRST
 * triangle it is found to hit instead of the closest one.  Contract: RayHit::Miss() -- meshIndex ==
 * 0xffffffff -- is exactly what lrb_trace (and the reference's Intersect, which SLG also uses for shadow
 * rays: Scene::Intersect with SHADOW_RAY, src/slg/scene/scene.cpp:556-575) reports for that ray; for a hit
 * the record describes SOME triangle the ray hits inside [mint, maxt] (t / b1 / b2 bit-exact for that
 * triangle), not necessarily the nearest.  Misses carry the closest-hit miss payload. */
LRB_API int lrb_trace_anyhit(lrb_scene *scene, const void *rays_dev, void *hits_dev, uint32_t ray_count);

/* ---- between two traces of a batch (the path tracer's ray producer / consumer contract) ------------ */
/* Dead-lane compaction (the reference re-launches a fixed-size Ray[taskCount] with dead lanes flagged
 * RAY_FLAGS_MASKED, include/slg/engines/pathoclbase/kernels/pathoclbase_kernels_micro.cl:34-106,1029):
 * builds the dense, increasing list of the indices of the non-masked rays.  The list and its length stay
 * in device memory owned by the device object (valid until the next compaction on this device);
 * live_count_host, when non-NULL, also receives the length (this blocks).  Device option "compact" = 1
 * makes every lrb_trace do this first. */
LRB_API int lrb_compact_rays(lrb_device *dev, const void *rays_dev, uint32_t ray_count,
		const uint32_t **live_idx_dev, const uint32_t **live_count_dev, uint32_t *live_count_host);
/* Traces only the rays listed in live_idx_dev[0 .. *live_count_dev) (ray_count entries when
 * live_count_dev is NULL); RayHit records are written at the rays' own indices.  any_hit != 0: shadow rays. */
LRB_API int lrb_trace_indexed(lrb_scene *scene, const void *rays_dev, void *hits_dev, uint32_t ray_count,
		const uint32_t *live_idx_dev, const uint32_t *live_count_dev, int any_hit);
/* One round of the reference's pass-through loop (Scene::Intersect, src/slg/scene/scene.cpp:556-690; GPU twin
 * include/slg/scene/scene_funcs.cl:21-150) over a traced batch.  A ray "continues to trace" when it hit a
 * mesh whose bit is set in pass_mesh_bits_dev (bit m of word m / 32: camera-invisible objects, fully
 * transparent materials, scene.cpp:646-668) or when continue_flags_dev[i] != 0 (the caller's own material
 * decision); either pointer may be NULL.  Such a ray is re-armed behind its hit -- ray.mint = hit.t +
 * MachineEpsilon::E(hit.t), scene.cpp:675 -- unless that leaves no interval (scene.cpp:679-680), in which
 * case it ends as a miss.  Every other ray gets RAY_FLAGS_MASKED, so that the next lrb_trace leaves its final
 * RayHit untouched.  MODIFIES the ray buffer.  n_continuing_host (optional; blocks) receives the number of
 * re-armed rays. */
LRB_API int lrb_advance_rays(lrb_scene *scene, void *rays_dev, void *hits_dev, uint32_t ray_count,
		const uint32_t *pass_mesh_bits_dev, uint32_t n_pass_words, const uint8_t *continue_flags_dev,
		uint32_t *n_continuing_host);
/* The whole loop: trace, advance, compact, re-trace ... until no ray continues or max_rounds (0 = 64)
 * rounds were traced.  On return every RayHit holds the first hit on a mesh that is not pass-through (or a
 * miss); the ray buffer has been consumed (see lrb_advance_rays).  Blocks. */
LRB_API int lrb_trace_passthrough(lrb_scene *scene, void *rays_dev, void *hits_dev, uint32_t ray_count,
		const uint32_t *pass_mesh_bits_dev, uint32_t n_pass_words, uint32_t max_rounds,
		uint32_t *rounds_out, uint64_t *rays_traced_out);

/* Host-buffer path (end to end): rays are copied in, traced and the hits copied out in chunks, with
 * the copies of neighbouring chunks overlapping the trace of the current one; synchronises before
 * returning.  preload_hits != 0 first uploads the caller's hit buffer, so that the RayHit of masked
 * rays keeps its previous content (what AllocBufferRW(&hits, hostHits, ...) does in the reference
 * sequence); with 0 the RayHit of a masked ray reads back as all-zero bytes. */
LRB_API int lrb_trace_host(lrb_scene *scene, const lrb_ray *rays, lrb_rayhit *hits, uint32_t ray_count, int preload_hits);
/* Same trace through the instrumented kernel; blocks and fills `stats`. hits_dev may be NULL. */
LRB_API int lrb_trace_stats(lrb_scene *scene, const void *rays_dev, void *hits_dev, uint32_t ray_count,
		lrb_trace_stats_t *stats);

/* ---- multi-GPU: RayHit gather over NVLink ------------------------------------------------- */
/* The BVH is replicated and every GPU traces its own slice of the rays (one process per GPU); the
 * only exchange of the path is collecting the RayHit slices in one buffer.  That buffer lives on
 * one GPU, is exported with lrb_ipc_get_handle and opened (peer-mapped over NVLink) by the other
 * processes with lrb_ipc_open_handle.  Handles are the 64 opaque bytes of cudaIpcMemHandle_t; the
 * pointer must be the start of an allocation made by lrb_alloc. */
#define LRB_IPC_HANDLE_BYTES 64
LRB_API int lrb_ipc_get_handle(lrb_device *dev, void *devptr, unsigned char handle[LRB_IPC_HANDLE_BYTES]);
LRB_API int lrb_ipc_open_handle(lrb_device *dev, const unsigned char handle[LRB_IPC_HANDLE_BYTES], void **devptr);
LRB_API int lrb_ipc_close_handle(lrb_device *dev, void *devptr);
/* Trace + gather, fused.  gather_dst_dev is this rank's slice of the gather buffer (local or
 * peer-mapped memory); hits_dev keeps the local copy (may be NULL with n_chunks == 0 when only the
 * gathered copy is wanted); gather_dst_dev == hits_dev is a plain trace.
 *   n_chunks == 0 (default): ONE kernel launch.
 *     - signalled pushes (device option gather_stores = 0, the default; needs hits_dev, the persistent
 *       kernel and the driver entry points cuStreamWaitValue32 / cuMemsetD32Async): the kernel raises a
 *       flag per finished chunk of 2^gather_chunk_shift ray indices and the copy engine pushes that chunk
 *       of hits_dev into gather_dst_dev on a second stream while the kernel keeps tracing.  Batches of at
 *       most 128 rays (one block: no room for the detector warp) take the dual-store form below.
 *     - dual-destination stores (gather_stores = 1, or hits_dev == NULL, or kernel = simple): every lane
 *       stores its finished RayHit record into hits_dev and into gather_dst_dev (posted stores over
 *       NVLink).  No copy engine.
 *     In both forms the records of masked rays are forwarded from hits_dev, so the gathered slice equals
 *     the local buffer.  No NCCL.
 *   n_chunks >= 1: the batch is cut into n_chunks launches; each traced piece is pushed by the copy
 *     engine on a second stream while the next piece is traced.
 * Asynchronous: later work on the device's stream is ordered after the last store / push. */
LRB_API int lrb_trace_gather(lrb_scene *scene, const void *rays_dev, void *hits_dev, uint32_t ray_count,
		void *gather_dst_dev, uint32_t n_chunks);
/* Deferred completion (device option "gather_defer" = 1, signalled pushes only): lrb_trace_gather then does
 * NOT make the queue wait for its pushes, so the next batch is traced while the tail of this one's RayHit
 * records is still on the wire.  The caller alternates two hits_dev buffers; the pushes of call k are waited
 * for automatically before call k + 2 traces, by lrb_sync, or explicitly here: makes cuda_stream (NULL = the
 * device's queue) wait for the pushes of the most recent call (which = 0), the one before it (1) or both (-1).
 * A completion signal for the gathering rank belongs behind this wait. */
LRB_API int lrb_gather_wait(lrb_device *dev, void *cuda_stream, int which);
/* Completion signal without a kernel and without NCCL: lrb_gather_signal writes `value` (a 16-bit step counter) into the
 * 32-bit flag word flag_dev -- normally a word of the gathering GPU's memory, peer-mapped like the gather buffer -- as
 * a 4-byte copy-engine transfer ordered behind this rank's pushes and behind everything queued so far.  On the
 * gathering GPU lrb_wait_value makes a stream (NULL = the device's queue) wait until that word is >= value
 * (cuStreamWaitValue32): no SM is involved on either side, so the signal is not serialised against a persistent trace
 * kernel that occupies every SM (which is what happens to a 4-byte ncclAllReduce: +0.25 ms per 5 ms step at >= 4 GPUs). */
LRB_API int lrb_gather_signal(lrb_device *dev, void *flag_dev, uint32_t value);
LRB_API int lrb_wait_value(lrb_device *dev, void *flag_dev, uint32_t value, void *cuda_stream);

/* ---- BVH construction on the device ------------------------------------------------------------- */
/* Replaces BuildEmbreeBVHMorton (src/luxrays/core/bvh/bvhembreebuild.cpp:218-336 with rtcBVHBuilderMorton,
 * declared in include/luxrays/core/bvh/bvhbuild.h:68-69): a linear BVH (63-bit Morton codes, radix sort, Karras
 * radix tree, k-ary collapse by depth) built on the GPU from the n_leaves leaf boxes the accelerators hand their
 * builders (BVHAccel::Init bvhaccel.cpp:100-135, MBVHAccel root tree mbvhaccel.cpp:132-200): leaf_boxes = 6
 * floats per leaf (min xyz, max xyz), HOST memory.  out_nodes (HOST, capacity >= 2 * n_leaves - 1) receives the
 * depth-first skip-list array of bvhclassicbuild.cpp:181-220 with at most tree_type (2 / 4 / 8) children per node;
 * a leaf node carries the INDEX of its input leaf in triangleLeaf.v[0] (== bvhLeaf.leafIndex) and zeros elsewhere:
 * the caller, which owns the mesh / instance tables, writes the leaf payload in.  Synchronous. */
typedef struct {
	double h2d_ms, sort_ms, tree_ms, emit_ms, d2h_ms;   /* CUDA-event times of the stages */
	uint32_t kernels;                                   /* own kernels launched (CUB's passes not counted) */
} lrb_build_timings;
LRB_API int lrb_build_lbvh(lrb_device *dev, const float *leaf_boxes, uint32_t n_leaves, uint32_t tree_type,
		lrb_bvh_node *out_nodes, uint32_t out_capacity, uint32_t *n_nodes, lrb_build_timings *timings);
/* Same interface with a choice of binary tree under the k-ary collapse: quality 0 = the radix tree of lrb_build_lbvh
 * (EMBREE_MORTON's trade-off: fastest build, dearest traversal), quality 1 = PLOC, parallel locally-ordered
 * clustering (Meister & Bittner 2018) with search radius 16 -- the stand-in for BuildEmbreeBVHBinnedSAH
 * (bvhembreebuild.cpp:218-336) when the tree is to be built on the GPU (host layer: builder type "B200_PLOC"). */
LRB_API int lrb_build_bvh(lrb_device *dev, const float *leaf_boxes, uint32_t n_leaves, uint32_t tree_type, uint32_t quality,
		lrb_bvh_node *out_nodes, uint32_t out_capacity, uint32_t *n_nodes, lrb_build_timings *timings);

/* Triangles in, traceable scene out -- everything on the device.  Replaces BVHAccel::Init + the builder + BVHKernel's
 * upload (src/luxrays/accelerators/bvhaccel.cpp:72-168, src/luxrays/core/bvh/bvhembreebuild.cpp:218-336,
 * src/luxrays/accelerators/bvhaccelhw.cpp:38-257) for trees that are built where the rays are traced: the triangles'
 * build boxes (bvhaccel.cpp:116-122), the tree of lrb_build_bvh, the leaf payload (bvhclassicbuild.cpp:196-214) and the
 * re-layout of lrb_bvh_upload all run as kernels; the BVHArrayNode array never visits the host unless it is asked for.
 * The scene is byte-identical to lrb_bvh_upload of that array (same functions, compiled for both sides).
 *   xyz / n_verts / mesh_vertex_offsets / n_meshes : as for lrb_bvh_upload (HOST memory);
 *   mesh_triangle_offsets : n_meshes + 1 entries, first triangle of mesh m in `triangles`; the last entry is the total;
 *   triangles             : 3 mesh-local vertex indices per triangle (luxrays::Triangle::v), all meshes back to back;
 *   out_nodes (may be NULL): receives the reference array (what BVHAccel::bvhTree would hold), capacity in nodes;
 *   n_nodes (may be NULL) : number of nodes of that array.
 * Synchronous.  0 or 1 triangle: the (trivial) tree is made on the host and goes through lrb_bvh_upload. */
typedef struct {
	double h2d_ms;                      /* vertices + triangle indices to the device */
	double leafbox_ms;                  /* build boxes of the triangles */
	double sort_ms, tree_ms, emit_ms;   /* the builder's stages, as in lrb_build_timings */
	double relayout_ms;                 /* leaf payload, wide nodes + triangle records, stack bound */
	double d2h_ms;                      /* download of the reference array (0 when out_nodes == NULL) */
	uint32_t kernels;                   /* own kernels launched (CUB's passes not counted) */
} lrb_scene_build_timings;
LRB_API int lrb_bvh_build_scene(lrb_device *dev, const float *xyz, uint64_t n_verts, const uint32_t *mesh_vertex_offsets,
		const uint32_t *mesh_triangle_offsets, uint32_t n_meshes, const uint32_t *triangles, uint32_t tree_type, uint32_t quality,
		lrb_scene **scene, lrb_bvh_node *out_nodes, uint32_t out_capacity, uint32_t *n_nodes, lrb_scene_build_timings *timings);
/* Hands a scene to another lrb_device handle of the SAME CUDA device (the accelerator is built before the intersection
 * device is started: Context::SetDataSet / Start, context.cpp:173-232): the scene's memory accounting moves to `dev`
 * and its traces run on dev's queue from now on.  Fails, and changes nothing, when the CUDA ordinals differ. */
LRB_API int lrb_scene_adopt(lrb_device *dev, lrb_scene *scene);
/* Diagnostics: copies the laid-out arrays of a single-level scene back to the host -- n_wide_nodes records of 64 B,
 * n_triangles records of 64 B, n_triangles id pairs of 8 B (counts from lrb_scene_get_info; any pointer may be NULL).
 * The tests compare a scene laid out on the device with the host's lay-out of the same array byte for byte. */
LRB_API int lrb_scene_download(lrb_scene *scene, void *wide_nodes, void *tri_records, void *tri_ids);

/* ---- multi-GPU: film merge over NVLink -------------------------------------------------------- */
/* The sum of per-GPU film planes that replaces the host-side merge of per-device films
 * (PathOCLRenderEngine::MergeThreadFilms -> Film::AddFilm, src/slg/engines/pathocl/pathocl.cpp:184-201,
 * src/slg/film/film.cpp:707-760):  dst[i] = (((0 + tile_0[i]) + tile_1[i]) + ...) for i in [first, first + count),
 * binary32 adds in tile (= device) order, bit for bit what the reference's loop computes.  tiles_dev is a HOST
 * array of n_tiles (<= 16) DEVICE pointers to float planes of identical layout -- this GPU's own film and the
 * other ranks' films opened with lrb_ipc_open_handle -- and dst_dev the merged film (local or peer-mapped).
 * Every rank calls this for its own slice of the film: peer loads over NVLink pull that slice of every film,
 * the sums go straight into the merged film on the gathering GPU.  Asynchronous on the device's queue; the
 * caller orders it after every rank finished writing its film and signals completion afterwards. */
LRB_API int lrb_film_reduce(lrb_device *dev, const float *const *tiles_dev, uint32_t n_tiles, float *dst_dev,
		uint64_t first, uint64_t count);

/* ---- diagnostics ----------------------------------------------------------------------- */
LRB_API const char *lrb_last_error_string(void);
LRB_API int lrb_get_counters(lrb_device *dev, lrb_counters *out);
LRB_API int lrb_reset_counters(lrb_device *dev);
LRB_API const char *lrb_version_string(void);
/* Roofline denominator probe: streams `bytes` with 256-bit read-only loads (eight in flight per thread) `iters` times and reports
 * GB/s.  A buffer that fits in L2 (e.g. 32 MiB) measures L2 bandwidth, a multi-GiB one HBM.  Blocks. */
LRB_API int lrb_measure_read_bandwidth(lrb_device *dev, size_t bytes, int iters, double *gb_per_s);

#ifdef __cplusplus
}
#endif
#endif /* LUXRAYS_B200_H */
