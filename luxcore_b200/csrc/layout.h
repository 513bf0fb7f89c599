// layout.h -- device-side data layout of a re-laid-out LuxRays BVH / MBVH.
//
// The reference traverses a depth-first skip-list of 32-byte BVHArrayNode records, one box or
// one triangle per record, with 3 x 12-byte unaligned vertex gathers per leaf
// (include/luxrays/accelerators/bvh.cl:136-217).  On upload we keep every reference box VALUE
// and every triangle, but re-lay them out for 256-bit vector loads:
//
//   WideNode  (64 B, 64-B aligned = two 32-B sectors, two 256-bit loads)
//     one per reference inner node (plus continuation nodes when a reference node has more than
//     four children, i.e. accelerator.bvh.treetype = 8).  Holds one box and one reference per child:
//       * inner child    : its own box from its BVHArrayNode record; reference = wide-node index;
//       * triangle leaf  : the triangle's build box -- bounds of the three vertices grown by
//                          MachineEpsilon::E(bbox), the box BVHAccel::Init hands to the builders
//                          (bvhaccel.cpp:116-122); reference = kTagTri | TriRecord index;
//       * MBVH root leaf : world-space bounds of the instance (its leaf tree's root box through the
//                          inverse of mInv, grown; relayout.cpp InstanceWorldBox) -- the reference
//                          enters every instance of a visited root node (mbvhaccel.cpp:312) and then
//                          tests that root box in instance space; motion-blurred instances get the
//                          bounds of that box over all times (relayout.cpp MotionWorldBox);
//                          reference = kTagInstance | index;
//       * unused slot    : an inverted box (lo = 255, hi = 0) and kNullIndex.
//     The four boxes are stored on a per-node grid: origin = min corner of the union of the slot
//     boxes, one power-of-two step per axis, 8 bits per plane, lo rounded down and hi rounded up,
//     so every stored box CONTAINS the reference's float box (a box test here passes whenever the
//     reference's passes, up to the rounding discussed in traverse.h; boxes only cull, they never
//     decide a result).  One 64-B fetch replaces up to five dependent 32-B fetches of the reference
//     walk, and every child (node or triangle) is ordered near-to-far and culled by entry distance.
//   TriRecord (64 B, 64-B aligned = two 256-bit loads): the three vertices pre-gathered next to the
//     reference's gate for the triangle (exact box of its parent node, see below), one per reference
//     triangle leaf, stored in the order of the leaves in the reference array -- the record index IS the
//     tie-break order; meshIndex / triangleIndex live in a side table (TriIds) read once per ray.
//   InstRecord (32 B): one per MBVH root leaf (bvhLeaf payload, bvhbuild_types.cl:33-37).
//
// Reference leaves carry no box of their own in the array (bvhclassicbuild.cpp:196-214): the
// reference tests every leaf child of a visited node.  The leaf box used here only skips triangle
// tests that cannot succeed (a hit point lies on the triangle, i.e. >= 128 ulp / 1e-5 inside the
// grown box), so results are unchanged; see DESIGN.md "leaf boxes".
//
// The reference keeps the FIRST hit in array order among hits with exactly equal t (strict `t <
// rayHit->t`, bvhaccel.cpp:233); a traversal in any other order reproduces that choice by
// preferring, on an exact tie, the leaf that comes first in the reference's depth-first array:
// triangle records are stored in that order (per tree), instances carry an `order` field.
#ifndef LRB_LAYOUT_H
#define LRB_LAYOUT_H

#include <stdint.h>

namespace lrb {

static const uint32_t kNullIndex = 0xffffffffu;
static const uint32_t kWideSlots = 4;

// child / stack references: bits 31-30 give the kind
static const uint32_t kTagTri = 0x40000000u;        // kTagTri | TriRecord index
static const uint32_t kTagInstance = 0x80000000u;   // kTagInstance | InstRecord index
static const uint32_t kRefIndexMask = 0x3fffffffu;
static const uint32_t kStackSentinel = 0xfffffffeu; // pop => leave the current instance
static const uint32_t kStackBottom = 0xfffffffdu;   // pop => the stack is empty (permanent entry under a shared-memory column)
static const uint32_t kMaxRefIndex = 0x3ffffff0u;   // node / triangle / instance counts stay below this

enum { kNodeEntry = 1 };    // WideNode::flags: one-child entry node carrying a tree's root box

struct __attribute__((aligned(64))) WideNode {
	float org[3];           // grid origin
	uint32_t exps;          // bytes 0-2: biased exponent byte of (step_axis * 2^15); byte 3: used slots
	uint32_t child[4];      // reference of slot k (wide index, kTagTri|i, kTagInstance|i) or kNullIndex
	uint32_t qlo[3];        // x, y, z: byte k = lower plane of slot k in grid steps
	uint32_t qhi[3];        // x, y, z: byte k = upper plane of slot k in grid steps
	uint32_t next;          // continuation node holding further children, or kNullIndex
	uint32_t flags;
};
static const int kGridShift = 15;   // plane byte q sits in bits 8-15 of a float mantissa: 1 + q * 2^-15

// The reference tests a leaf triangle exactly when the boxes of all its inner ancestors pass, and
// its Triangle::Intersect can report a "hit" on a triangle the ray does not touch when the ray lies
// in the triangle's plane (divisor = rounding noise).  The traversal here uses boxes that CONTAIN
// the reference's, so it may reach such a triangle where the reference never does.  The gate is
// the exact box of the triangle's parent node (the last, and tightest, of the reference's gates --
// ancestor boxes are unions of their children's, so they pass whenever it does); a hit is accepted
// only if the reference's own box arithmetic passes it.  It travels in the triangle's own record:
// the test runs branch-free next to the triangle test, with no second, dependent fetch.
struct __attribute__((aligned(64))) TriRecord {
	float p0[3], p1[3], p2[3];
	float gateLo[3], gateHi[3];
	uint32_t order;         // index of the leaf in its reference BVHArrayNode array (layout checks; the kernels use the record index)
};

struct TriIds {
	uint32_t meshIndex, triangleIndex;
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

static_assert(sizeof(WideNode) == 64, "WideNode");
static_assert(sizeof(TriRecord) == 64, "TriRecord");
static_assert(sizeof(TriIds) == 8, "TriIds");
static_assert(sizeof(InstRecord) == 32, "InstRecord");
static_assert(sizeof(DevInterp) == 16 + 3 * 64 + 6 * 16, "DevInterp");

// pointers the traversal kernels read (all device pointers on the GPU)
struct SceneView {
	const WideNode *nodes;
	const TriRecord *tris;
	const TriIds *ids;              // one per TriRecord
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
	// 0x3f800000 (1.0f).  Read from the parameter bank so that the plane decode's PRMT keeps its byte
	// selector as the immediate operand (with a literal constant ptxas moves the selectors to registers).
	uint32_t oneBits;
};

}   // namespace lrb

#endif
