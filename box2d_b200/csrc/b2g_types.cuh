// b2g_types.cuh -- device-side data layout of one solver step.
//
// Data layout in HBM (all buffers owned by b2GpuSolver, resident for the whole step):
//
//   raw inputs (uploaded as the reference's own AoS, read once by the prepare stage)
//     rawStates   b2BodyState[bodyCount]            32 B each
//     rawSims     b2BodySim[bodyCount]              96 B each
//     rawContacts b2ContactSim[contactSlots]       200 B each, colour c at slot colors[c].contactStart
//     joints      b2JointSim[jointSlots]           252 B each (working copy, solved in place)
//
//   solver state (SoA, what the 3*subSteps*colours hot stages touch)
//     vel[1+bodyCount]  float4 {v.x, v.y, w, flags-bits}        index 0 = static dummy (identity)
//     pos[1+bodyCount]  float4 {dp.x, dp.y, dq.c, dq.s}         index 0 = {0,0,1,0}
//     bodyK[bodyCount]  float4 {lvd.x, lvd.y, avd, linDamp} + angDamp[bodyCount]   (per-step body constants)
//     contact constraint = 10 float4 + 2 int2 per slot, field-major: field f of slot s at f*slotCapacity+s
//       so a warp reads 512 contiguous bytes per field (one thread per constraint).
//
// Splitting b2BodyState's two 16-byte halves into two arrays keeps the float4 gathers the north_star
// asks for and lets warm-start / restitution skip the position half.
#pragma once

#include "b2g_math.cuh"

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

struct ColorRange
{
	int contactStart; // slot of the colour's first contact (multiple of 32)
	int contactCount;
	int jointStart; // index of the colour's first joint in the joint working array
	int jointCount;
};

constexpr int kMaxColors = 23;
constexpr int kStageTimerCount = 8;

// Per-contact output record copied back to the host and scattered into b2Manifold
// (what b2StoreImpulsesTask writes, src/contact_solver.c:2293-2303): 9 floats.
constexpr int kImpulseFloats = 9;

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
	int slotCapacity; // field stride of the contact SoA
	int hitWords;	  // uint32 words in hitBits
	int jointWords;	  // uint32 words in jointBits

	// raw inputs
	const uint8_t* rawStates;
	const uint8_t* rawSims;
	const uint8_t* rawContacts;
	const uint8_t* rawJoints; // pristine prepared joints as uploaded

	// solver state
	float4* vel;
	float4* pos;
	float4* bodyK;
	float* angDamp;
	float4* cf; // CF_COUNT * slotCapacity
	int2* cidx; // indexA+1, indexB+1 (0 = static)
	int2* cmeta; // contactId, (simFlags & hitEvent) | pointCount
	uint8_t* joints; // b2JointSim working copy

	// outputs
	uint8_t* outStates; // b2BodyState[bodyCount]
	float* outImpulses; // kImpulseFloats per slot
	uint32_t* hitBits;
	uint32_t* jointBits;
	int* hasHitEvents;
	int* anyRestitution; // set by prepare when some contact has restitution != 0 (else the restitution stages are skipped)

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
