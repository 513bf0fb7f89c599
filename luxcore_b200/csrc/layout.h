// layout.h -- device-side data layout of a re-laid-out LuxRays BVH / MBVH.
//
// The reference traverses a depth-first skip-list of 32-byte BVHArrayNode records, one box or
// one triangle per record, with 3 x 12-byte unaligned vertex gathers per leaf
// (include/luxrays/accelerators/bvh.cl:136-217).  On upload we keep every reference box VALUE
// and every triangle, but re-lay them out for 128-bit vector loads:
//
//   WideNode  (128 B, 128-B aligned = one L2 line / four 32-B sectors)
//     one per reference inner node (plus continuation nodes when a reference node has more than
//     four inner children, i.e. accelerator.bvh.treetype = 8).  Holds the boxes of up to four
//     INNER children in struct-of-arrays form (the boxes are copied bit-for-bit from the children's
//     own BVHArrayNode records) and their wide-node indices, so one fetch replaces up to five
//     dependent 32-B fetches of the reference walk.
//   TriRecord (64 B, 64-B aligned = two 256-bit loads): the three vertices pre-gathered next to
//     meshIndex / triangleIndex, one per reference triangle leaf; the leaf children of one node are
//     contiguous.
//   InstRecord (32 B): one per MBVH root leaf (bvhLeaf payload, bvhbuild_types.cl:33-37).
//
// Reference leaves carry no box of their own (bvhclassicbuild.cpp:196-214), so -- exactly like
// the reference -- every leaf child of a visited node is tested without a box pre-test.
//
// `order` fields record the position of the leaf in the reference's depth-first array.  The
// reference keeps the FIRST hit in array order among hits with exactly equal t (strict `t <
// rayHit->t`, bvhaccel.cpp:233); a traversal in any other order reproduces that choice by
// preferring the smaller order on an exact tie.
#ifndef LRB_LAYOUT_H
#define LRB_LAYOUT_H

#include <stdint.h>

namespace lrb {

static const uint32_t kNullIndex = 0xffffffffu;
static const uint32_t kWideSlots = 4;

// stack entry tags (two-level traversal)
static const uint32_t kTagInstance = 0x80000000u;   // entry = kTagInstance | instance record index
static const uint32_t kStackSentinel = 0xffffffffu; // pop => leave the current instance

struct __attribute__((aligned(128))) WideNode {
	float bminx[4], bminy[4], bminz[4];
	float bmaxx[4], bmaxy[4], bmaxz[4];
	uint32_t child[4];      // wide-node index of inner child k, k < nInner
	uint32_t leafBase;      // first TriRecord (or InstRecord, in an MBVH root tree) of this node
	uint32_t counts;        // bits 0-7 nInner, bits 8-31 nLeaf
	uint32_t next;          // continuation node holding further children, or kNullIndex
	uint32_t pad;
};

struct __attribute__((aligned(64))) TriRecord {
	float p0[3], p1[3], p2[3];
	uint32_t meshIndex, triangleIndex;
	uint32_t order;         // index of the leaf in its reference BVHArrayNode array
	uint32_t pad[4];        // 64 B: two 256-bit loads, never straddles a 128-B line
};

struct __attribute__((aligned(16))) InstRecord {
	uint32_t rootWide;      // wide-node index of the leaf tree's root (absolute)
	uint32_t transformIndex, motionIndex;   // at most one != kNullIndex
	uint32_t meshOffset;    // == dataset mesh index reported in RayHit
	uint32_t order;         // index of the root leaf in the reference root array
	uint32_t pad[3];
};

// What traversal needs from one 576-B ocl::InterpolatedTransform (motionsystem_types.cl:33-47).
struct __attribute__((aligned(16))) DevInterp {
	float startTime, endTime;
	uint32_t flags;
	uint32_t pad;
	float startM[16];       // start.m
	float endM[16];         // end.m
	float R[16];            // startT.R
	float sS[4], eS[4];     // startT/endT Sx,Sy,Sz
	float sT[4], eT[4];     // startT/endT Tx,Ty,Tz
	float sQ[4], eQ[4];     // startQ/endQ as w,x,y,z
};
enum {
	kItActive = 1, kItRotation = 2, kItTranslation = 4, kItScale = 8,
	kItTX = 16, kItTY = 32, kItTZ = 64
};

static_assert(sizeof(WideNode) == 128, "WideNode");
static_assert(sizeof(TriRecord) == 64, "TriRecord");
static_assert(sizeof(InstRecord) == 32, "InstRecord");
static_assert(sizeof(DevInterp) == 16 + 3 * 64 + 6 * 16, "DevInterp");

// pointers the traversal kernels read (all device pointers on the GPU)
struct SceneView {
	const WideNode *nodes;
	const TriRecord *tris;
	const InstRecord *insts;
	const float *minv;              // 16 floats per instance transform, row-major
	const uint32_t *motionFirst;    // per motion system: first / last DevInterp index
	const uint32_t *motionLast;
	const DevInterp *interps;
	uint32_t nWide;                 // 0 => empty scene, every ray misses
	uint32_t rootWide;              // wide index where traversal starts
	uint32_t twoLevel;
	// The top-level root box travels in the kernel parameters (constant bank): its test costs no
	// memory access and rays that miss the scene never touch a node.  rootHasBox == 0 when the
	// tree's root is itself a leaf (it has no box in the reference either).
	uint32_t rootHasBox;
	uint32_t rootChild;             // wide index of the real root node
	float rootBox[6];               // min xyz, max xyz of the reference's node 0
};

}   // namespace lrb

#endif
