// b2g_types.cuh -- device-side data layout of one solver step.
//
// Wire format (what crosses PCIe, packed by the host side of the C-ABI from the reference's own arrays):
//     states   b2BodyState[bodyCount]      32 B each, copied as is
//     wireBody 2 float4 per body           {invMass, invInertia, force.x, force.y} {torque, linDamping, angDamping, gravityScale}
//     wire     6 float4 per contact slot   96 of the 112 bytes of b2ContactSim's 200 the solver reads (see WireRow; the
//                                          masses travel separately and only when they differ from the bodies'); colour c
//                                          occupies slots [colors[c].contactStart, +contactCount), starts are multiples of 32
//     joints   b2JointSim[jointCount]      252 B each padded to 256, prepared on the host; solved in place in a working copy
//
// Solver state, reached through a SolveView (generic pointers: global memory for the grid-barrier kernel, shared
// memory for the island-local kernel -- the stage code is the same):
//     vel[1+n]  float4 {v.x, v.y, w, flags-bits}        index 0 = static dummy (identity)
//     pos[1+n]  float4 {dp.x, dp.y, dq.c, dq.s}         index 0 = {0,0,1,0}
//     bodyK[n]  float4 {lvd.x, lvd.y, avd, linDamp} + angDamp[n]   (per-step body constants)
//     contact constraint = 10 float4 + int2 + int per slot, field-major: field f of slot s at f*cfStride+s, so a warp
//       reads 512 contiguous bytes per field (one thread per constraint).
//
// Splitting b2BodyState's two 16-byte halves into two arrays keeps the float4 gathers the north_star asks for and
// lets warm-start / restitution skip the position half.
#pragma once

#include "b2g_math.cuh"

#include "b2_gpu_solver.h"

namespace b2g
{

// float4 field groups of a contact constraint (replaces the reference's 8/4-wide b2ContactConstraintWide,
// src/contact_solver.c:1070-1100, by one lane per constraint)
enum ContactField
{
	CF_MASS = 0,	// invMassA, invIA, invMassB, invIB
	CF_NORMAL = 1,	// normal.x, normal.y, friction, tangentSpeed
	CF_ROLL = 2,	// rollingResistance, restitution, rollingMass, (unused)
	CF_SOFT = 3,	// biasRate, massScale, impulseScale, (unused)
	CF_ANCHOR1 = 4, // anchorA1.xy, anchorB1.xy
	CF_ANCHOR2 = 5, // anchorA2.xy, anchorB2.xy
	CF_PMASS = 6,	// normalMass1, tangentMass1, normalMass2, tangentMass2
	CF_BASE = 7,	// baseSeparation1, baseSeparation2, relativeVelocity1, relativeVelocity2
	CF_IMP1 = 8,	// (mutable) normalImpulse1, tangentImpulse1, totalNormalImpulse1, rollingImpulse
	CF_IMP2 = 9,	// (mutable) normalImpulse2, tangentImpulse2, totalNormalImpulse2, (unused)
	CF_COUNT = 10
};

// rows of a wire contact record (float4 each)
enum WireRow
{
	WR_HEAD = 0,	// indexA (int), indexB (int), meta (int: colour<<8 | hitEnable<<2 | pointCount), rollingImpulse
	WR_NORMAL = 1,	// normal.x, normal.y, friction, tangentSpeed
	WR_MATERIAL = 2, // rollingResistance, restitution, separation1, separation2
	WR_ANCHOR1 = 3, // anchorA1.xy, anchorB1.xy
	WR_ANCHOR2 = 4, // anchorA2.xy, anchorB2.xy
	WR_IMPULSE = 5, // normalImpulse1, tangentImpulse1, normalImpulse2, tangentImpulse2
	WR_COUNT = 6
};
// The seventh row -- invMassA, invIA, invMassB, invIB of b2ContactSim (src/contact.h:103-142) -- has a region of its own at
// the end of the input arena (StepParams::wireMass).  The reference copies these from the bodies when a contact enters the
// constraint graph (src/constraint_graph.c, b2AddContactToGraph), so they almost always equal what the body constants
// already carry; the pack pass checks that bit for bit and the region is only uploaded for a step in which some contact
// differs (a body whose mass changed under a persisting contact).

constexpr int kMetaPointMask = 3;
constexpr int kMetaHitEnable = 4;
constexpr int kMetaGroupRolling = 8;	   // some lane of this constraint's SIMD group of 4 has rolling resistance
constexpr int kMetaGroupRestitution = 16; // ... has restitution
constexpr int kMetaColorShift = 8;

// ---- resident mode (single worlds): the device keeps what does not change from step to step ------------------------------
// Every contact has a HOME: (graph colour, index in the colour's array) mapped into a table that keeps its layout from
// step to step (the colours' regions have spare room, see b2gBegin).  The pack pass compares every contact with a
// host-side shadow of what the device holds at its home and every body with what came back from the previous step.
// Per step and contact slot only a 16-byte LIGHT record is uploaded: { key, separation1, separation2, ref }.  A contact
// whose record is unchanged but for its separations -- the reference's recycled manifolds, src/physics_world.c:508-560 --
// is read from the resident TABLE (static rows, by home) and its warm-start impulses from the previous step's own output
// records (ref = its slot in that step); anything else -- a re-evaluated manifold, a contact that is new at its home
// after a swap-remove (src/constraint_graph.c:198-211) -- travels as a FULL record (ref = ~index into the step's full
// stream) and is copied into the table after the solve.
//   key  bits 0..27 home, bit 28/29 the SIMD-group bits (kMetaGroup*), bit 30: the contact's inverse masses are its
//        bodies' (nothing was written to the mass region for it), negative = dead slot
constexpr int kLightIdMask = ( 1 << 28 ) - 1;
constexpr int kLightBodyMass = 1 << 30;
constexpr int kLightGroupShift = 28; // key >> 28 & 3 -> kMetaGroupRolling | kMetaGroupRestitution after << 3
constexpr int kTableRows = 5;		 // WR_HEAD .. WR_ANCHOR2 of a contact, by home
constexpr int kLightJointQuads = 6;	 // { home, ref, -, - } + B2L_JOINT_RUN_MAX bytes of prepared run
constexpr int kDirtyBodyQuads = 5;	 // { body index, -, -, - }, the b2BodyState (2 quads), the packed constants (2 quads)

struct ColorRange
{
	int contactStart; // slot of the colour's first contact (multiple of 32)
	int contactCount;
	int jointStart; // index of the colour's first joint in the joint working array
	int jointCount;
};

constexpr int kMaxColors = 23;
constexpr int kJointStride = 256; // b2JointSim (252 B) padded to 16-byte multiples on the wire and in the working copy
// A plain revolute joint (no spring, motor or limit) as the island / cluster kernels keep it in shared memory when EVERY
// joint of the step is one (b2g_joint.cuh, LiteRevolute): 31 floats instead of 64.  An odd number of words, so that a
// warp's accesses to one field of 32 consecutive records hit 32 different banks.
constexpr int kLiteJointWords = 31;
constexpr int kStageTimerCount = 8;

// Per-contact output record copied back to the host and scattered into b2Manifold
// (what b2StoreImpulsesTask writes, src/contact_solver.c:2293-2303): 9 floats + the hit-event flag (:2305-2320).
constexpr int kImpulseFloats = 10;

// The arrays the stage functions work on.  Generic pointers: global memory or shared memory.
struct SolveView
{
	float4* vel;
	float4* pos;
	float4* bodyK;
	float* angDamp;
	float4* cf;	   // CF_COUNT field groups, field f of slot s at f*cfStride + s
	int cfStride;
	int2* cidx;	   // indexA+1, indexB+1 (0 = static dummy) in the view's own body numbering
	int* cmeta;	   // kMeta* bits
	uint8_t* joints; // b2JointSim working copy (indexA/indexB in the view's numbering, 0-based, -1 = static)
	int jointLite;	 // != 0: `joints` holds LiteRevolute records of kLiteJointWords floats instead (island / cluster kernels)
	int* anyRestitution; // set by prepare when a contact of the view has restitution != 0
	// Cluster mode (island kernel on a thread-block cluster): the body arrays are distributed over the blocks' shared
	// memory in runs of clusterRun bodies; body index i (1-based, 0 = this block's own static dummy) lives in block
	// (i-1) / clusterRun at slot (i-1) % clusterRun + 1 and is reached through distributed shared memory.  The division
	// is a multiply-high with clusterMagic = ceil(2^32 / clusterRun), exact for i < 65536.  clusterRun == 0: flat view.
	int clusterRun;
	unsigned clusterMagic;
	// != 0: body writes are st.async stores that report to the owner block's mbarrier at this shared::cta address
	// (b2g_cluster.cuh); 0: plain stores
	unsigned asyncBar;
};

struct StepParams
{
	// b2StepContext / b2World scalars (include/b2_gpu_solver.h b2GpuStepDesc)
	float dt, inv_dt, h, inv_h;
	int subStepCount;
	Soft contactSoft;
	Soft staticSoft;
	float restitutionThreshold;
	float maxLinearVelocity;
	float gravityX, gravityY;
	float contactSpeed;
	float contactHertz;
	float contactDampingRatio;
	float hitEventThreshold;
	float lengthUnitsPerMeter;
	int enableWarmStarting;
	int enableSoftening;

	int bodyCount;
	int colorCount;
	ColorRange colors[kMaxColors];
	ColorRange overflow;
	int contactSlots; // slots in use, colours + overflow, padded
	int jointCount;	  // joints in use, colours + overflow
	int jointWords;	  // uint32 words in jointBits

	// wire inputs (global memory)
	const uint8_t* rawStates;
	const float4* wireBody;
	const float4* wire;		  // WR_COUNT float4 per slot, AoS
	const uint8_t* rawJoints; // pristine prepared joints as uploaded
	const float4* wireMass;	  // [contactSlots] invMassA, invIA, invMassB, invIB of every contact -- valid when massFromBodies == 0
	int massFromBodies;		  // every contact's masses equal its bodies' (checked by the pack pass): read them from wireBody

	// resident mode (light != nullptr): `wire` is unused, rawStates / wireBody point at the resident body arrays
	const float4* light;		// [contactSlots] { key, separation1, separation2, ref }
	float4* table;				// [homes * kTableRows] static rows of the contacts, by home
	const float4* full;			// the step's full records (WR_COUNT rows each)
	const float* prevImpulses;	// the previous step's impulse records (kImpulseFloats per slot of THAT step)
	uint8_t* residentOut;		// [bodyCount] b2BodyState the next step starts from: this step's v, w, flags with the deltas reset
	float4* residentStates;		// writable aliases of rawStates / wireBody for the apply pass (dirty bodies)
	float4* residentBody;
	const float4* dirtyBodies;	// [dirtyBodyCapacity * kDirtyBodyQuads] bodies whose state or constants the host changed
	int dirtyBodyCapacity;
	// joints in resident mode (lightJoints != nullptr): b2gAssembleJointsKernel builds rawJoints from these before the step
	const float4* lightJoints;	// [jointCount * kLightJointQuads] { home, ref, -, - } + the joint's prepared run (b2lJointPreparedRun)
	float4* jointTable;			// [joint homes * kJointStride / 16] the records as assembled for the last step, by home
	const float4* fullJoints;	// the step's full joint records (kJointStride bytes each)
	const float* prevOutJoints; // the previous step's joint impulse records (B2L_JOINT_OUT_FLOATS per joint of THAT step)
	float4* jointAssembled;		// writable alias of rawJoints

	// solver state in global memory (the grid-barrier kernel's view)
	SolveView g;

	// outputs
	uint8_t* outStates; // b2BodyState[bodyCount]
	float* outImpulses; // kImpulseFloats per slot
	float* outJoints;	// B2L_JOINT_OUT_FLOATS per joint: the fields the stages wrote (b2lJointMutableRuns)
	uint32_t* jointBits;
	int* hasHitEvents;

	// island-local mode (b2g_island.cuh): islands are packed into bins, one thread block solves one bin entirely in
	// shared memory.  Scratch written by the partition kernel, read by the island kernel.
	int binCount;	   // 0 = island mode off for this step
	int clusterSize;   // thread blocks per bin (1 = one block per bin, no cluster)
	int resolveContacts; // the partition kernel fills binContactInfo; 0 (diagnostics): the island kernels chase head -> bodyLocal themselves
	int stageAllThreads; // diagnostics: every thread of a block walks the stage loops (B2GPU_STAGE_ALL=1)
	int clusterRun;	   // bodies per block in cluster mode
	unsigned clusterMagic; // ceil(2^32 / clusterRun)
	int capBodies;	   // per-BLOCK capacities the shared memory carve-up was sized for (a bin holds clusterSize times that)
	int capContacts;
	int capJoints;
	int liteJoints;	   // island / cluster kernels: every joint of the step is a plain revolute joint (checked by the scatter / partition kernel: binFail otherwise)
	int jointsSpilled; // cluster kernel: the joint records stay in the global working copy instead of shared memory
	int ownerLists;	   // one bin shared by a cluster: constraint lists per BLOCK, keyed by the owner of the first body
	int listCount;	   // number of constraint lists: binCount, or clusterSize with owner lists
	int leveliseContacts; // flat lists: jointless bins run their coloured contacts level by level (b2g_island.cuh)
	// The bins' plans (one block per bin, flat lists): a step that keeps its lists also writes down, per bin, the order it put its
	// constraints in -- sorted by colour, levelised -- and the offsets of the levels; a step that runs on the previous step's
	// lists reads that instead of sorting and levelising again (b2gIslandKernel).
	int planWrite, planRead;
	int* planStart;	   // [bins][2][kColorSlots] first contact / first joint of every level, totals in the last entry
	int4* planInfo;	   // [bins][capContacts] { wire slot, bin-local body A, B, SIMD-group bits } in the bin's final order
	int* planJoints;   // [bins][capJoints] joint index, final order
	int keepLists;	   // one block per bin, flat lists: the island kernel leaves the bins' counters as they are (the next step may run on the same lists, b2gEnqueueRun)
	int flatLists;	   // one block per bin: b2gScatterKernel appends to flat per-bin lists, the island kernel sorts them by colour
	int listCapContacts; // stride of the constraint lists (binCap*, or the per-block capacity with owner lists)
	int listCapJoints;
	int binCapBodies;  // strides of the per-bin lists: capacity of a whole bin (= per-block capacity * clusterSize)
	int binCapContacts;
	int binCapJoints;
	const int* bodyBin;	 // [bodyCount] bin of each body (wire arena)
	int* bodyLocal;		 // [bodyCount] 1-based index of the body inside its bin
	int* binBodyCount;	 // [binCount]
	int* binBodyList;	 // [binCount * binCapBodies] global body index
	int* binColorStart;	 // [binCount * kColorSlots] contact counts per colour (+ the overflow bucket)
	int* binJointStart;	 // same for joints
	int* binColorTotal;	 // [2 * kColorSlots] owner lists: contacts / joints of the whole bin per colour
	int* binColorOffset; // [binCount * kColorSlots] exclusive offsets of the colours in the bin's contact list (+ total)
	int* binJointOffset; // same for joints
	int2* contactBinRank; // [contactSlots] bin, rank within (bin, colour)
	int* slotGroupBits;	 // [contactSlots] kMetaGroup* bits in wire order
	int* binContactList; // [binCount * binCapContacts] wire slots, colour-major
	int4* binContactInfo; // same layout: { wire slot, bin-local body A, bin-local body B, SIMD-group bits }
	int2* jointBinRank;	 // [jointCount]
	int* binJointList;	 // [binCount * binCapJoints] joint index, colour-major
	int2* binJointBodies; // flat lists: the two bodies (wire indices, -1 = static) of every entry of binJointList
	int* islandFailed;	 // control block: set by the island kernels when they give up (binFail), read by the host
	int* binFail;		 // set when some bin does not fit its capacities: the grid-barrier kernel takes the step

	// grid-barrier kernel: joints per block it may keep in shared memory for the whole step (b2g_stages.cuh, JointCache)
	int gridJointCache;
	int gridJointsAllCached; // the capacity covers the fullest block: no coloured joint needs the global working copy

	// sync + profiling
	unsigned int* barrier;			 // [0] arrival counter, [1] exit counter
	unsigned long long* stageCycles; // kStageTimerCount clock64 accumulators, [8] barrier count, [9] total cycles
};

// Stage operations, in the order of b2SolverTask (reference src/solver.c:1055-1197)
enum StageOp
{
	OP_PREPARE = 0,			 // load bodies, prepare contacts (coloured + overflow), stage joints, clear event bits
	OP_INTEGRATE_VELOCITIES, // solver.c:66
	OP_OVERFLOW_WARM,		 // solver.c:1100-1101
	OP_WARM,				 // per colour
	OP_OVERFLOW_SOLVE,		 // solver.c:1119-1120 (useBias) / 1147-1148 (relax)
	OP_SOLVE,				 // per colour, useBias
	OP_INTEGRATE_POSITIONS,	 // solver.c:114
	OP_OVERFLOW_RELAX,
	OP_RELAX, // per colour, useBias = false
	OP_OVERFLOW_RESTITUTION,
	OP_RESTITUTION, // per colour
	OP_STORE		// store impulses (coloured + overflow) + write states back as AoS
};

} // namespace b2g
