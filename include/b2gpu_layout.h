/*
 * b2gpu_layout.h -- byte layout of the reference structures that cross the C-ABI.
 *
 * These are LAYOUT FACTS about Box2D v3.2.0 (release build, single precision positions) measured from
 * the reference headers; they are the "wire format" of include/b2_gpu_solver.h.  The host seam
 * (box2d_b200/host/b2_gpu_seam.c) static-asserts every one of them against the real reference types, so
 * a reference upgrade that moves a field fails the build instead of silently corrupting the solve.
 *
 *   b2BodyState  src/body.h:153-168      b2BodySim   src/body.h:175-208
 *   b2ContactSim src/contact.h:103-142   b2Manifold  include/box2d/collision.h:572-586
 *   b2ManifoldPoint include/box2d/collision.h:527-568
 *   b2JointSim   src/joint.h:267-301 and the per-type unions src/joint.h:64-262
 */
#ifndef B2GPU_LAYOUT_H
#define B2GPU_LAYOUT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

/* ---- b2BodyState (32 B): two 16-byte halves ---------------------------------------------------------- */
#define B2L_STATE_SIZE 32
/* half 0: linearVelocity.xy, angularVelocity, flags(u32) ; half 1: deltaPosition.xy, deltaRotation.c,.s */

/* body flag bits used by the solver, src/body.h:14-65 */
#define B2L_FLAG_LOCK_LINEAR_X 0x00000001u
#define B2L_FLAG_LOCK_LINEAR_Y 0x00000002u
#define B2L_FLAG_LOCK_ANGULAR_Z 0x00000004u
#define B2L_FLAG_IS_SPEED_CAPPED 0x00000020u
#define B2L_FLAG_ALLOW_FAST_ROTATION 0x00000080u
#define B2L_FLAG_DYNAMIC 0x00000200u
/* b2_bodyTransientFlags = b2_isFast | b2_isSpeedCapped | b2_hadTimeOfImpact (src/body.h:64), cleared by b2FinalizeBodiesTask */
#define B2L_FLAG_TRANSIENT 0x00000068u

/* ---- b2BodySim (96 B) ------------------------------------------------------------------------------ */
#define B2L_SIM_SIZE 96
#define B2L_SIM_FORCE 48
#define B2L_SIM_TORQUE 56
#define B2L_SIM_INV_MASS 60
#define B2L_SIM_INV_INERTIA 64
#define B2L_SIM_LINEAR_DAMPING 76
#define B2L_SIM_ANGULAR_DAMPING 80
#define B2L_SIM_GRAVITY_SCALE 84

/* ---- b2ContactSim (200 B) -------------------------------------------------------------------------- */
#define B2L_CONTACT_SIZE 200
#define B2L_CONTACT_ID 0
#define B2L_CONTACT_INDEX_A 36
#define B2L_CONTACT_INDEX_B 40
#define B2L_CONTACT_INV_MASS_A 52
#define B2L_CONTACT_INV_I_A 56
#define B2L_CONTACT_INV_MASS_B 60
#define B2L_CONTACT_INV_I_B 64
#define B2L_CONTACT_MANIFOLD 68
#define B2L_CONTACT_FRICTION 172
#define B2L_CONTACT_RESTITUTION 176
#define B2L_CONTACT_ROLLING_RESISTANCE 180
#define B2L_CONTACT_TANGENT_SPEED 184
#define B2L_CONTACT_SIM_FLAGS 188
#define B2L_SIM_ENABLE_HIT_EVENT 0x00100000u /* b2_simEnableHitEvent, src/contact.h:43 */

/* b2Manifold (104 B), offsets relative to the manifold */
#define B2L_MANIFOLD_NORMAL 0
#define B2L_MANIFOLD_ROLLING_IMPULSE 8
#define B2L_MANIFOLD_POINTS 12
#define B2L_MANIFOLD_POINT_COUNT 100
/* b2ManifoldPoint (44 B), offsets relative to the point */
#define B2L_MP_SIZE 44
#define B2L_MP_ANCHOR_A 0
#define B2L_MP_ANCHOR_B 8
#define B2L_MP_SEPARATION 16
#define B2L_MP_NORMAL_IMPULSE 24
#define B2L_MP_TANGENT_IMPULSE 28
#define B2L_MP_TOTAL_NORMAL_IMPULSE 32
#define B2L_MP_NORMAL_VELOCITY 36

/* ---- b2JointSim (252 B) ---------------------------------------------------------------------------- */
#define B2L_JOINT_SIZE 252

/* b2JointType, include/box2d/types.h */
enum
{
	b2l_distanceJoint = 0,
	b2l_filterJoint = 1,
	b2l_motorJoint = 2,
	b2l_moverJoint = 3,
	b2l_pogoJoint = 4,
	b2l_prismaticJoint = 5,
	b2l_revoluteJoint = 6,
	b2l_weldJoint = 7,
	b2l_wheelJoint = 8
};

typedef struct b2lVec2
{
	float x, y;
} b2lVec2;

typedef struct b2lRot
{
	float c, s;
} b2lRot;

typedef struct b2lTransform
{
	b2lVec2 p;
	b2lRot q;
} b2lTransform;

typedef struct b2lSoft
{
	float biasRate, massScale, impulseScale;
} b2lSoft;

typedef struct b2lMat22
{
	b2lVec2 cx, cy;
} b2lMat22;

/* src/joint.h:64-84 */
typedef struct b2lPogo
{
	b2lVec2 normal;
	float restLength, hertz, dampingRatio, maxTensionForce, maxCompressionForce;
	float impulse;
	int indexA, indexB;
	b2lTransform frameA, frameB;
	b2lVec2 deltaCenter;
	float linearMass;
	float velocity;
} b2lPogo;

/* src/joint.h:86-115 */
typedef struct b2lDistance
{
	float length, hertz, dampingRatio, lowerSpringForce, upperSpringForce, minLength, maxLength;
	float maxMotorForce, motorSpeed;
	float impulse, lowerImpulse, upperImpulse, motorImpulse;
	int indexA, indexB;
	b2lVec2 anchorA, anchorB, deltaCenter;
	b2lSoft distanceSoftness;
	float axialMass;
	uint8_t enableSpring, enableLimit, enableMotor;
} b2lDistance;

/* src/joint.h:117-146 */
typedef struct b2lMotor
{
	b2lVec2 linearVelocity;
	float maxVelocityForce, angularVelocity, maxVelocityTorque;
	float linearHertz, linearDampingRatio, maxSpringForce;
	float angularHertz, angularDampingRatio, maxSpringTorque;
	b2lVec2 linearVelocityImpulse;
	float angularVelocityImpulse;
	b2lVec2 linearSpringImpulse;
	float angularSpringImpulse;
	b2lSoft linearSpring, angularSpring;
	int indexA, indexB;
	b2lTransform frameA, frameB;
	b2lVec2 deltaCenter;
	b2lMat22 linearMass;
	float angularMass;
} b2lMotor;

/* src/joint.h:148-160 */
typedef struct b2lMover
{
	b2lVec2 linearVelocity, maxVelocityForce, linearVelocityImpulse;
	int indexA, indexB;
	b2lTransform frameA, frameB;
	float linearMass;
} b2lMover;

/* src/joint.h:162-188 */
typedef struct b2lPrismatic
{
	b2lVec2 impulse;
	float springImpulse, motorImpulse, lowerImpulse, upperImpulse;
	float hertz, dampingRatio, targetTranslation, maxMotorForce, motorSpeed;
	float lowerTranslation, upperTranslation;
	int indexA, indexB;
	b2lTransform frameA, frameB;
	b2lVec2 deltaCenter;
	b2lSoft springSoftness;
	uint8_t enableSpring, enableLimit, enableMotor;
} b2lPrismatic;

/* src/joint.h:190-217 */
typedef struct b2lRevolute
{
	b2lVec2 linearImpulse;
	float springImpulse, motorImpulse, lowerImpulse, upperImpulse;
	float hertz, dampingRatio, targetAngle, maxMotorTorque, motorSpeed;
	float lowerAngle, upperAngle;
	int indexA, indexB;
	b2lTransform frameA, frameB;
	b2lVec2 deltaCenter;
	float axialMass;
	b2lSoft springSoftness;
	uint8_t enableSpring, enableMotor, enableLimit;
} b2lRevolute;

/* src/joint.h:219-237 */
typedef struct b2lWeld
{
	float linearHertz, linearDampingRatio, angularHertz, angularDampingRatio;
	b2lSoft linearSpring, angularSpring;
	b2lVec2 linearImpulse;
	float angularImpulse;
	int indexA, indexB;
	b2lTransform frameA, frameB;
	b2lVec2 deltaCenter;
	float axialMass;
} b2lWeld;

/* src/joint.h:239-265 */
typedef struct b2lWheel
{
	float perpImpulse, motorImpulse, springImpulse, lowerImpulse, upperImpulse;
	float maxMotorTorque, motorSpeed, lowerTranslation, upperTranslation, hertz, dampingRatio;
	int indexA, indexB;
	b2lTransform frameA, frameB;
	b2lVec2 deltaCenter;
	float perpMass, motorMass, axialMass;
	b2lSoft springSoftness;
	uint8_t enableSpring, enableMotor, enableLimit;
} b2lWheel;

/* src/joint.h:267-301 */
typedef struct b2lJointSim
{
	int jointId;
	int bodyIdA, bodyIdB;
	int type;
	b2lTransform localFrameA, localFrameB;
	float invMassA, invMassB, invIA, invIB;
	float constraintHertz, constraintDampingRatio;
	b2lSoft constraintSoftness;
	float forceThreshold, torqueThreshold;
	union
	{
		b2lDistance distance;
		b2lMotor motor;
		b2lMover mover;
		b2lPogo pogo;
		b2lRevolute revolute;
		b2lPrismatic prismatic;
		b2lWeld weld;
		b2lWheel wheel;
	} u;
} b2lJointSim;

/* What the solver stages WRITE in a joint sim (everything else is input): the accumulated impulses of every type
 * (src/*_joint.c: b2WarmStart*Joint / b2Solve*Joint), plus the motor joint's linearMass (src/motor_joint.c, recomputed
 * in the solve) and the pogo joint's velocity.  At most two runs of consecutive floats per type, at most
 * B2L_JOINT_OUT_FLOATS floats: this is the record the device sends back per joint. */
#define B2L_JOINT_OUT_FLOATS 12
#if defined( __CUDACC__ )
#define B2L_HD __host__ __device__
#else
#define B2L_HD
#endif
#include <stddef.h>
/* returns the number of runs; offsets are bytes from the start of b2JointSim */
static inline B2L_HD int b2lJointMutableRuns( int type, int offsets[2], int floats[2] )
{
	switch ( type )
	{
		case b2l_distanceJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.distance.impulse ), floats[0] = 4;
			return 1;
		case b2l_motorJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.motor.linearVelocityImpulse ), floats[0] = 6;
			offsets[1] = (int)offsetof( b2lJointSim, u.motor.linearMass ), floats[1] = 4;
			return 2;
		case b2l_moverJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.mover.linearVelocityImpulse ), floats[0] = 2;
			return 1;
		case b2l_pogoJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.pogo.impulse ), floats[0] = 1;
			offsets[1] = (int)offsetof( b2lJointSim, u.pogo.velocity ), floats[1] = 1;
			return 2;
		case b2l_prismaticJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.prismatic.impulse ), floats[0] = 6;
			return 1;
		case b2l_revoluteJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.revolute.linearImpulse ), floats[0] = 6;
			return 1;
		case b2l_weldJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.weld.linearImpulse ), floats[0] = 3;
			return 1;
		case b2l_wheelJoint:
			offsets[0] = (int)offsetof( b2lJointSim, u.wheel.perpImpulse ), floats[0] = 5;
			return 1;
		default:
			return 0; /* filter joint: nothing */
	}
}

/* What b2PrepareJoint (src/joint.c:1406 + the per-type prepare functions) REWRITES every step from the bodies' poses: the
 * body indices, the anchor frames / anchors in world orientation, deltaCenter and the effective masses that depend on them.
 * In every joint type these fields are one run of consecutive bytes that starts at indexA.  Everything outside this run and
 * outside the mutable runs above only changes when the application changes the joint (or the step parameters).  Returns the
 * length of the run in bytes (at most B2L_JOINT_RUN_MAX) and its offset from the start of b2JointSim; 0 for a filter joint. */
#define B2L_JOINT_RUN_MAX 80
static inline B2L_HD int b2lJointPreparedRun( int type, int* offset )
{
	switch ( type )
	{
		case b2l_distanceJoint: /* indexA .. axialMass (anchors, deltaCenter, distanceSoftness, axialMass) */
			*offset = (int)offsetof( b2lJointSim, u.distance.indexA );
			return (int)( offsetof( b2lJointSim, u.distance.axialMass ) + 4 - offsetof( b2lJointSim, u.distance.indexA ) );
		case b2l_motorJoint: /* indexA .. angularMass */
			*offset = (int)offsetof( b2lJointSim, u.motor.indexA );
			return (int)( offsetof( b2lJointSim, u.motor.angularMass ) + 4 - offsetof( b2lJointSim, u.motor.indexA ) );
		case b2l_moverJoint: /* indexA .. linearMass */
			*offset = (int)offsetof( b2lJointSim, u.mover.indexA );
			return (int)( offsetof( b2lJointSim, u.mover.linearMass ) + 4 - offsetof( b2lJointSim, u.mover.indexA ) );
		case b2l_pogoJoint: /* indexA .. linearMass (velocity behind it is the solver's own) */
			*offset = (int)offsetof( b2lJointSim, u.pogo.indexA );
			return (int)( offsetof( b2lJointSim, u.pogo.linearMass ) + 4 - offsetof( b2lJointSim, u.pogo.indexA ) );
		case b2l_prismaticJoint: /* indexA .. deltaCenter */
			*offset = (int)offsetof( b2lJointSim, u.prismatic.indexA );
			return (int)( offsetof( b2lJointSim, u.prismatic.deltaCenter ) + 8 - offsetof( b2lJointSim, u.prismatic.indexA ) );
		case b2l_revoluteJoint: /* indexA .. axialMass */
			*offset = (int)offsetof( b2lJointSim, u.revolute.indexA );
			return (int)( offsetof( b2lJointSim, u.revolute.axialMass ) + 4 - offsetof( b2lJointSim, u.revolute.indexA ) );
		case b2l_weldJoint: /* indexA .. axialMass */
			*offset = (int)offsetof( b2lJointSim, u.weld.indexA );
			return (int)( offsetof( b2lJointSim, u.weld.axialMass ) + 4 - offsetof( b2lJointSim, u.weld.indexA ) );
		case b2l_wheelJoint: /* indexA .. axialMass (perpMass, motorMass, axialMass) */
			*offset = (int)offsetof( b2lJointSim, u.wheel.indexA );
			return (int)( offsetof( b2lJointSim, u.wheel.axialMass ) + 4 - offsetof( b2lJointSim, u.wheel.indexA ) );
		default:
			*offset = 0;
			return 0;
	}
}

#if defined( __cplusplus )
static_assert( sizeof( b2lJointSim ) == B2L_JOINT_SIZE, "b2JointSim mirror size" );
static_assert( sizeof( b2lRevolute ) == 120 && sizeof( b2lWeld ) == 104 && sizeof( b2lPrismatic ) == 116, "joint mirrors" );
static_assert( sizeof( b2lWheel ) == 120 && sizeof( b2lDistance ) == 104 && sizeof( b2lMotor ) == 160, "joint mirrors" );
static_assert( sizeof( b2lMover ) == 68 && sizeof( b2lPogo ) == 88, "joint mirrors" );
#else
_Static_assert( sizeof( b2lJointSim ) == B2L_JOINT_SIZE, "b2JointSim mirror size" );
_Static_assert( sizeof( b2lRevolute ) == 120 && sizeof( b2lWeld ) == 104 && sizeof( b2lPrismatic ) == 116, "joint mirrors" );
_Static_assert( sizeof( b2lWheel ) == 120 && sizeof( b2lDistance ) == 104 && sizeof( b2lMotor ) == 160, "joint mirrors" );
_Static_assert( sizeof( b2lMover ) == 68 && sizeof( b2lPogo ) == 88, "joint mirrors" );
#endif

#ifdef __cplusplus
}
#endif

#endif /* B2GPU_LAYOUT_H */
