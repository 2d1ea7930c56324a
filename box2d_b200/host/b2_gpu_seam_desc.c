/*
 * b2_gpu_seam_desc.c -- layout checks + step descriptor + host joint prepare (no device dependency).
 *
 * Compiled together with the reference's own (unmodified) translation units; it includes the reference's
 * internal headers from where they lie (nothing of the reference is copied into this repository).
 * See b2_gpu_seam.h and INTEGRATION.md.
 */
#include "b2_gpu_seam.h"

#include "b2gpu_layout.h"

/* reference internals (include path: <reference>/src and <reference>/include) */
#include "bitset.h"
#include "body.h"
#include "constraint_graph.h"
#include "contact.h"
#include "core.h"
#include "id_pool.h"
#include "island.h"
#include "joint.h"
#include "parallel_for.h"
#include "physics_world.h"
#include "solver.h"
#include "solver_set.h"

#include "box2d/base.h"
#include "box2d/constants.h"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- the wire format must be the reference's layout, or the build fails ------------------------------- */
_Static_assert( sizeof( b2BodyState ) == B2L_STATE_SIZE, "b2BodyState" );
_Static_assert( offsetof( b2BodyState, linearVelocity ) == 0 && offsetof( b2BodyState, angularVelocity ) == 8, "b2BodyState" );
_Static_assert( offsetof( b2BodyState, flags ) == 12 && offsetof( b2BodyState, deltaPosition ) == 16, "b2BodyState" );
_Static_assert( offsetof( b2BodyState, deltaRotation ) == 24, "b2BodyState" );
_Static_assert( sizeof( b2BodySim ) == B2L_SIM_SIZE, "b2BodySim" );
_Static_assert( offsetof( b2BodySim, force ) == B2L_SIM_FORCE && offsetof( b2BodySim, torque ) == B2L_SIM_TORQUE, "b2BodySim" );
_Static_assert( offsetof( b2BodySim, invMass ) == B2L_SIM_INV_MASS, "b2BodySim" );
_Static_assert( offsetof( b2BodySim, invInertia ) == B2L_SIM_INV_INERTIA, "b2BodySim" );
_Static_assert( offsetof( b2BodySim, linearDamping ) == B2L_SIM_LINEAR_DAMPING, "b2BodySim" );
_Static_assert( offsetof( b2BodySim, angularDamping ) == B2L_SIM_ANGULAR_DAMPING, "b2BodySim" );
_Static_assert( offsetof( b2BodySim, gravityScale ) == B2L_SIM_GRAVITY_SCALE, "b2BodySim" );
_Static_assert( sizeof( b2ContactSim ) == B2L_CONTACT_SIZE, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, contactId ) == B2L_CONTACT_ID, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, bodySimIndexA ) == B2L_CONTACT_INDEX_A, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, bodySimIndexB ) == B2L_CONTACT_INDEX_B, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, invMassA ) == B2L_CONTACT_INV_MASS_A, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, invIA ) == B2L_CONTACT_INV_I_A, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, invMassB ) == B2L_CONTACT_INV_MASS_B, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, invIB ) == B2L_CONTACT_INV_I_B, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, manifold ) == B2L_CONTACT_MANIFOLD, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, friction ) == B2L_CONTACT_FRICTION, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, restitution ) == B2L_CONTACT_RESTITUTION, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, rollingResistance ) == B2L_CONTACT_ROLLING_RESISTANCE, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, tangentSpeed ) == B2L_CONTACT_TANGENT_SPEED, "b2ContactSim" );
_Static_assert( offsetof( b2ContactSim, simFlags ) == B2L_CONTACT_SIM_FLAGS, "b2ContactSim" );
_Static_assert( offsetof( b2Manifold, normal ) == B2L_MANIFOLD_NORMAL, "b2Manifold" );
_Static_assert( offsetof( b2Manifold, rollingImpulse ) == B2L_MANIFOLD_ROLLING_IMPULSE, "b2Manifold" );
_Static_assert( offsetof( b2Manifold, points ) == B2L_MANIFOLD_POINTS, "b2Manifold" );
_Static_assert( offsetof( b2Manifold, pointCount ) == B2L_MANIFOLD_POINT_COUNT, "b2Manifold" );
_Static_assert( sizeof( b2ManifoldPoint ) == B2L_MP_SIZE, "b2ManifoldPoint" );
_Static_assert( offsetof( b2ManifoldPoint, anchorA ) == B2L_MP_ANCHOR_A, "b2ManifoldPoint" );
_Static_assert( offsetof( b2ManifoldPoint, anchorB ) == B2L_MP_ANCHOR_B, "b2ManifoldPoint" );
_Static_assert( offsetof( b2ManifoldPoint, separation ) == B2L_MP_SEPARATION, "b2ManifoldPoint" );
_Static_assert( offsetof( b2ManifoldPoint, normalImpulse ) == B2L_MP_NORMAL_IMPULSE, "b2ManifoldPoint" );
_Static_assert( offsetof( b2ManifoldPoint, tangentImpulse ) == B2L_MP_TANGENT_IMPULSE, "b2ManifoldPoint" );
_Static_assert( offsetof( b2ManifoldPoint, totalNormalImpulse ) == B2L_MP_TOTAL_NORMAL_IMPULSE, "b2ManifoldPoint" );
_Static_assert( offsetof( b2ManifoldPoint, normalVelocity ) == B2L_MP_NORMAL_VELOCITY, "b2ManifoldPoint" );
_Static_assert( b2_simEnableHitEvent == B2L_SIM_ENABLE_HIT_EVENT, "sim flags" );
_Static_assert( b2_lockLinearX == B2L_FLAG_LOCK_LINEAR_X && b2_lockLinearY == B2L_FLAG_LOCK_LINEAR_Y, "body flags" );
_Static_assert( b2_lockAngularZ == B2L_FLAG_LOCK_ANGULAR_Z && b2_isSpeedCapped == B2L_FLAG_IS_SPEED_CAPPED, "body flags" );
_Static_assert( b2_allowFastRotation == B2L_FLAG_ALLOW_FAST_ROTATION && b2_dynamicFlag == B2L_FLAG_DYNAMIC, "body flags" );
_Static_assert( (unsigned)b2_bodyTransientFlags == B2L_FLAG_TRANSIENT, "body flags" );
_Static_assert( B2_GRAPH_COLOR_COUNT == B2GPU_GRAPH_COLOR_COUNT, "colour count" );
_Static_assert( sizeof( b2Softness ) == sizeof( b2GpuSoftness ), "b2Softness" );

/* b2JointSim and every per-type block */
_Static_assert( sizeof( b2JointSim ) == sizeof( b2lJointSim ), "b2JointSim" );
_Static_assert( offsetof( b2JointSim, type ) == offsetof( b2lJointSim, type ), "b2JointSim" );
_Static_assert( offsetof( b2JointSim, invMassA ) == offsetof( b2lJointSim, invMassA ), "b2JointSim" );
_Static_assert( offsetof( b2JointSim, invIB ) == offsetof( b2lJointSim, invIB ), "b2JointSim" );
_Static_assert( offsetof( b2JointSim, constraintSoftness ) == offsetof( b2lJointSim, constraintSoftness ), "b2JointSim" );
_Static_assert( offsetof( b2JointSim, forceThreshold ) == offsetof( b2lJointSim, forceThreshold ), "b2JointSim" );
_Static_assert( offsetof( b2JointSim, torqueThreshold ) == offsetof( b2lJointSim, torqueThreshold ), "b2JointSim" );
_Static_assert( offsetof( b2JointSim, revoluteJoint ) == offsetof( b2lJointSim, u ), "b2JointSim" );
_Static_assert( (int)b2_distanceJoint == (int)b2l_distanceJoint && (int)b2_filterJoint == (int)b2l_filterJoint, "joint types" );
_Static_assert( (int)b2_motorJoint == (int)b2l_motorJoint && (int)b2_moverJoint == (int)b2l_moverJoint, "joint types" );
_Static_assert( (int)b2_pogoJoint == (int)b2l_pogoJoint && (int)b2_prismaticJoint == (int)b2l_prismaticJoint, "joint types" );
_Static_assert( (int)b2_revoluteJoint == (int)b2l_revoluteJoint && (int)b2_weldJoint == (int)b2l_weldJoint, "joint types" );
_Static_assert( (int)b2_wheelJoint == (int)b2l_wheelJoint, "joint types" );

#define B2S_SAME( T, M, f ) _Static_assert( offsetof( T, f ) == offsetof( M, f ), #T "." #f )
_Static_assert( sizeof( b2RevoluteJoint ) == sizeof( b2lRevolute ), "revolute" );
B2S_SAME( b2RevoluteJoint, b2lRevolute, linearImpulse );
B2S_SAME( b2RevoluteJoint, b2lRevolute, springImpulse );
B2S_SAME( b2RevoluteJoint, b2lRevolute, upperImpulse );
B2S_SAME( b2RevoluteJoint, b2lRevolute, hertz );
B2S_SAME( b2RevoluteJoint, b2lRevolute, targetAngle );
B2S_SAME( b2RevoluteJoint, b2lRevolute, maxMotorTorque );
B2S_SAME( b2RevoluteJoint, b2lRevolute, motorSpeed );
B2S_SAME( b2RevoluteJoint, b2lRevolute, lowerAngle );
B2S_SAME( b2RevoluteJoint, b2lRevolute, upperAngle );
B2S_SAME( b2RevoluteJoint, b2lRevolute, indexA );
B2S_SAME( b2RevoluteJoint, b2lRevolute, frameA );
B2S_SAME( b2RevoluteJoint, b2lRevolute, frameB );
B2S_SAME( b2RevoluteJoint, b2lRevolute, deltaCenter );
B2S_SAME( b2RevoluteJoint, b2lRevolute, axialMass );
B2S_SAME( b2RevoluteJoint, b2lRevolute, springSoftness );
B2S_SAME( b2RevoluteJoint, b2lRevolute, enableSpring );
B2S_SAME( b2RevoluteJoint, b2lRevolute, enableMotor );
B2S_SAME( b2RevoluteJoint, b2lRevolute, enableLimit );
_Static_assert( sizeof( b2WeldJoint ) == sizeof( b2lWeld ), "weld" );
B2S_SAME( b2WeldJoint, b2lWeld, linearHertz );
B2S_SAME( b2WeldJoint, b2lWeld, angularHertz );
B2S_SAME( b2WeldJoint, b2lWeld, linearSpring );
B2S_SAME( b2WeldJoint, b2lWeld, angularSpring );
B2S_SAME( b2WeldJoint, b2lWeld, linearImpulse );
B2S_SAME( b2WeldJoint, b2lWeld, angularImpulse );
B2S_SAME( b2WeldJoint, b2lWeld, indexA );
B2S_SAME( b2WeldJoint, b2lWeld, frameA );
B2S_SAME( b2WeldJoint, b2lWeld, deltaCenter );
B2S_SAME( b2WeldJoint, b2lWeld, axialMass );
_Static_assert( sizeof( b2PrismaticJoint ) == sizeof( b2lPrismatic ), "prismatic" );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, impulse );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, springImpulse );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, motorImpulse );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, lowerImpulse );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, upperImpulse );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, targetTranslation );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, maxMotorForce );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, motorSpeed );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, lowerTranslation );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, upperTranslation );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, indexA );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, frameA );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, deltaCenter );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, springSoftness );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, enableSpring );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, enableLimit );
B2S_SAME( b2PrismaticJoint, b2lPrismatic, enableMotor );
_Static_assert( sizeof( b2WheelJoint ) == sizeof( b2lWheel ), "wheel" );
B2S_SAME( b2WheelJoint, b2lWheel, perpImpulse );
B2S_SAME( b2WheelJoint, b2lWheel, motorImpulse );
B2S_SAME( b2WheelJoint, b2lWheel, springImpulse );
B2S_SAME( b2WheelJoint, b2lWheel, lowerImpulse );
B2S_SAME( b2WheelJoint, b2lWheel, upperImpulse );
B2S_SAME( b2WheelJoint, b2lWheel, maxMotorTorque );
B2S_SAME( b2WheelJoint, b2lWheel, motorSpeed );
B2S_SAME( b2WheelJoint, b2lWheel, lowerTranslation );
B2S_SAME( b2WheelJoint, b2lWheel, upperTranslation );
B2S_SAME( b2WheelJoint, b2lWheel, indexA );
B2S_SAME( b2WheelJoint, b2lWheel, frameA );
B2S_SAME( b2WheelJoint, b2lWheel, deltaCenter );
B2S_SAME( b2WheelJoint, b2lWheel, perpMass );
B2S_SAME( b2WheelJoint, b2lWheel, motorMass );
B2S_SAME( b2WheelJoint, b2lWheel, axialMass );
B2S_SAME( b2WheelJoint, b2lWheel, springSoftness );
B2S_SAME( b2WheelJoint, b2lWheel, enableSpring );
B2S_SAME( b2WheelJoint, b2lWheel, enableMotor );
B2S_SAME( b2WheelJoint, b2lWheel, enableLimit );
_Static_assert( sizeof( b2DistanceJoint ) == sizeof( b2lDistance ), "distance" );
B2S_SAME( b2DistanceJoint, b2lDistance, length );
B2S_SAME( b2DistanceJoint, b2lDistance, lowerSpringForce );
B2S_SAME( b2DistanceJoint, b2lDistance, upperSpringForce );
B2S_SAME( b2DistanceJoint, b2lDistance, minLength );
B2S_SAME( b2DistanceJoint, b2lDistance, maxLength );
B2S_SAME( b2DistanceJoint, b2lDistance, maxMotorForce );
B2S_SAME( b2DistanceJoint, b2lDistance, motorSpeed );
B2S_SAME( b2DistanceJoint, b2lDistance, impulse );
B2S_SAME( b2DistanceJoint, b2lDistance, lowerImpulse );
B2S_SAME( b2DistanceJoint, b2lDistance, upperImpulse );
B2S_SAME( b2DistanceJoint, b2lDistance, motorImpulse );
B2S_SAME( b2DistanceJoint, b2lDistance, indexA );
B2S_SAME( b2DistanceJoint, b2lDistance, anchorA );
B2S_SAME( b2DistanceJoint, b2lDistance, anchorB );
B2S_SAME( b2DistanceJoint, b2lDistance, deltaCenter );
B2S_SAME( b2DistanceJoint, b2lDistance, distanceSoftness );
B2S_SAME( b2DistanceJoint, b2lDistance, axialMass );
B2S_SAME( b2DistanceJoint, b2lDistance, enableSpring );
B2S_SAME( b2DistanceJoint, b2lDistance, enableLimit );
B2S_SAME( b2DistanceJoint, b2lDistance, enableMotor );
_Static_assert( sizeof( b2MotorJoint ) == sizeof( b2lMotor ), "motor" );
B2S_SAME( b2MotorJoint, b2lMotor, linearVelocity );
B2S_SAME( b2MotorJoint, b2lMotor, maxVelocityForce );
B2S_SAME( b2MotorJoint, b2lMotor, angularVelocity );
B2S_SAME( b2MotorJoint, b2lMotor, maxVelocityTorque );
B2S_SAME( b2MotorJoint, b2lMotor, linearHertz );
B2S_SAME( b2MotorJoint, b2lMotor, maxSpringForce );
B2S_SAME( b2MotorJoint, b2lMotor, angularHertz );
B2S_SAME( b2MotorJoint, b2lMotor, maxSpringTorque );
B2S_SAME( b2MotorJoint, b2lMotor, linearVelocityImpulse );
B2S_SAME( b2MotorJoint, b2lMotor, angularVelocityImpulse );
B2S_SAME( b2MotorJoint, b2lMotor, linearSpringImpulse );
B2S_SAME( b2MotorJoint, b2lMotor, angularSpringImpulse );
B2S_SAME( b2MotorJoint, b2lMotor, linearSpring );
B2S_SAME( b2MotorJoint, b2lMotor, angularSpring );
B2S_SAME( b2MotorJoint, b2lMotor, indexA );
B2S_SAME( b2MotorJoint, b2lMotor, frameA );
B2S_SAME( b2MotorJoint, b2lMotor, deltaCenter );
B2S_SAME( b2MotorJoint, b2lMotor, linearMass );
B2S_SAME( b2MotorJoint, b2lMotor, angularMass );
_Static_assert( sizeof( b2MoverJoint ) == sizeof( b2lMover ), "mover" );
B2S_SAME( b2MoverJoint, b2lMover, linearVelocity );
B2S_SAME( b2MoverJoint, b2lMover, maxVelocityForce );
B2S_SAME( b2MoverJoint, b2lMover, linearVelocityImpulse );
B2S_SAME( b2MoverJoint, b2lMover, indexA );
B2S_SAME( b2MoverJoint, b2lMover, linearMass );
_Static_assert( sizeof( b2PogoJoint ) == sizeof( b2lPogo ), "pogo" );
B2S_SAME( b2PogoJoint, b2lPogo, normal );
B2S_SAME( b2PogoJoint, b2lPogo, restLength );
B2S_SAME( b2PogoJoint, b2lPogo, hertz );
B2S_SAME( b2PogoJoint, b2lPogo, dampingRatio );
B2S_SAME( b2PogoJoint, b2lPogo, maxTensionForce );
B2S_SAME( b2PogoJoint, b2lPogo, maxCompressionForce );
B2S_SAME( b2PogoJoint, b2lPogo, impulse );
B2S_SAME( b2PogoJoint, b2lPogo, indexA );
B2S_SAME( b2PogoJoint, b2lPogo, frameA );
B2S_SAME( b2PogoJoint, b2lPogo, frameB );
B2S_SAME( b2PogoJoint, b2lPogo, deltaCenter );
B2S_SAME( b2PogoJoint, b2lPogo, linearMass );
B2S_SAME( b2PogoJoint, b2lPogo, velocity );

/* ---- descriptor ----------------------------------------------------------------------------------------- */

void b2GpuSeam_BuildDesc( b2World* world, b2StepContext* context, b2GpuStepDesc* desc )
{
	memset( desc, 0, sizeof( *desc ) );

	desc->dt = context->dt;
	desc->inv_dt = context->inv_dt;
	desc->h = context->h;
	desc->inv_h = context->inv_h;
	desc->subStepCount = context->subStepCount;
	memcpy( &desc->contactSoftness, &context->contactSoftness, sizeof( b2Softness ) );
	memcpy( &desc->staticSoftness, &context->staticSoftness, sizeof( b2Softness ) );
	desc->restitutionThreshold = world->restitutionThreshold;
	desc->maxLinearVelocity = context->maxLinearVelocity;

	desc->gravity[0] = world->gravity.x;
	desc->gravity[1] = world->gravity.y;
	desc->contactSpeed = world->contactSpeed;
	desc->contactHertz = world->contactHertz;
	desc->contactDampingRatio = world->contactDampingRatio;
	desc->hitEventThreshold = world->hitEventThreshold;
	desc->lengthUnitsPerMeter = b2GetLengthUnitsPerMeter();
	desc->enableWarmStarting = world->enableWarmStarting ? 1 : 0;
	desc->enableContactSoftening = world->enableContactSoftening ? 1 : 0;

	b2SolverSet* awakeSet = world->solverSets.data + b2_awakeSet;
	desc->states = awakeSet->bodyStates.data;
	desc->sims = awakeSet->bodySims.data;
	desc->awakeBodyCount = awakeSet->bodySims.count;

	// Active colours in ascending colour index: the same walk as reference src/solver.c:1341-1367.
	b2GraphColor* colors = world->constraintGraph.colors;
	int c = 0;
	for ( int i = 0; i < B2_GRAPH_COLOR_COUNT - 1; ++i )
	{
		int contactCount = colors[i].contactSims.count;
		int jointCount = colors[i].jointSims.count;
		if ( contactCount + jointCount == 0 )
		{
			continue;
		}
		desc->colors[c].contactSims = colors[i].contactSims.data;
		desc->colors[c].contactCount = contactCount;
		desc->colors[c].jointSims = colors[i].jointSims.data;
		desc->colors[c].jointCount = jointCount;
		desc->colors[c].colorIndex = i;
		c += 1;
	}
	desc->activeColorCount = c;

	b2GraphColor* overflow = colors + B2_OVERFLOW_INDEX;
	desc->overflow.contactSims = overflow->contactSims.data;
	desc->overflow.contactCount = overflow->contactSims.count;
	desc->overflow.jointSims = overflow->jointSims.data;
	desc->overflow.jointCount = overflow->jointSims.count;
	desc->overflow.colorIndex = B2_OVERFLOW_INDEX;

	desc->contactIdCapacity = b2GetIdCapacity( &world->contactIdPool );
	desc->jointIdCapacity = b2GetIdCapacity( &world->jointIdPool );
}

/* ---- island hint ------------------------------------------------------------------------------------------- */

typedef struct b2SeamIslandTask
{
	b2World* world;
	const b2BodySim* sims;
	int* labels;
} b2SeamIslandTask;

// label = index of the body's island among the awake islands (b2Island::localIndex, src/island.h:49-74)
static void b2SeamIslandLabelsTask( int startIndex, int endIndex, int workerIndex, void* taskContext )
{
	(void)workerIndex;
	b2SeamIslandTask* task = taskContext;
	const b2Body* bodies = task->world->bodies.data;
	const b2Island* islands = task->world->islands.data;
	for ( int i = startIndex; i < endIndex; ++i )
	{
		int islandId = bodies[task->sims[i].bodyId].islandId;
		task->labels[i] = islandId == B2_NULL_INDEX ? -1 : islands[islandId].localIndex;
	}
}

void b2GpuSeam_FillIslands( b2World* world, b2GpuStepDesc* desc, int* labels, b2GpuIslandSize* sizes, bool parallel )
{
	b2SolverSet* awakeSet = world->solverSets.data + b2_awakeSet;
	if ( sizes != NULL )
	{
		// what the reference already keeps per island (src/island.h:64-73): its bodies, its touching contacts, its joints
		const b2Island* islands = world->islands.data;
		int islandCount = awakeSet->islandSims.count;
		for ( int i = 0; i < islandCount; ++i )
		{
			const b2Island* island = islands + awakeSet->islandSims.data[i].islandId;
			sizes[i].bodyCount = island->bodies.count;
			sizes[i].contactCount = island->contacts.count;
			sizes[i].jointCount = island->joints.count;
			sizes[i].reserved = 0;
		}
	}
	desc->islandSizes = sizes;
	desc->islandCount = awakeSet->islandSims.count;
	if ( labels == NULL )
	{
		return; // the caller fills the labels itself (the seam's team)
	}
	b2SeamIslandTask task = { world, awakeSet->bodySims.data, labels };
	int count = awakeSet->bodySims.count;
	if ( parallel )
	{
		b2ParallelFor( world, b2SeamIslandLabelsTask, count, 512, &task );
	}
	else
	{
		b2SeamIslandLabelsTask( 0, count, 0, &task );
	}
	desc->bodyIsland = labels;
	desc->islandCount = awakeSet->islandSims.count;
}

/* ---- host joint prepare --------------------------------------------------------------------------------- */

typedef struct b2SeamJointRange
{
	b2StepContext* context;
	b2JointSim* joints[B2_GRAPH_COLOR_COUNT];
	int starts[B2_GRAPH_COLOR_COUNT + 1];
	int colorCount;
} b2SeamJointRange;

// b2ParallelFor callback over the flat joint index range of all active colours.
static void b2SeamPrepareJointsTask( int startIndex, int endIndex, int workerIndex, void* taskContext )
{
	(void)workerIndex;
	b2SeamJointRange* range = taskContext;
	int color = 0;
	while ( range->starts[color + 1] <= startIndex )
	{
		color += 1;
	}
	for ( int i = startIndex; i < endIndex; ++i )
	{
		while ( range->starts[color + 1] <= i )
		{
			color += 1;
		}
		b2PrepareJoint( range->joints[color] + ( i - range->starts[color] ), range->context );
	}
}

void b2GpuSeam_PrepareJoints( b2World* world, b2StepContext* context )
{
	b2GraphColor* colors = world->constraintGraph.colors;
	b2SeamJointRange range;
	range.context = context;
	int c = 0;
	int total = 0;
	for ( int i = 0; i < B2_GRAPH_COLOR_COUNT - 1; ++i )
	{
		int jointCount = colors[i].jointSims.count;
		if ( jointCount == 0 )
		{
			continue;
		}
		range.joints[c] = colors[i].jointSims.data;
		range.starts[c] = total;
		total += jointCount;
		c += 1;
	}
	range.starts[c] = total;
	range.colorCount = c;

	if ( total > 0 )
	{
		b2ParallelFor( world, b2SeamPrepareJointsTask, total, 16, &range );
	}

	// The overflow colour is prepared serially by the reference too (src/solver.c:1077).
	b2PrepareJoints_Overflow( context );
}

