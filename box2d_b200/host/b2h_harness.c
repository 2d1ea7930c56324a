/*
 * b2h_harness.c -- scene + stepping harness over the reference's PUBLIC API (scene driver for the tests and bench.py;
 * it only calls include/box2d/box2d.h, so it behaves the same in front of the CPU solver and in front of the seam).
 *
 * The same file is linked into
 *    oracle/_ref/libbox2d_ref.so      the untouched reference (CPU solver)           -> the parity oracle
 *    oracle/_ref/libbox2d_refcap.so   the reference + capture hooks (b2h_capture.c)  -> golden fixtures
 *    box2d_b200/libbox2d_b200.so      the reference host + the B200 solver seam      -> the product
 * so a test can create the very same scene in two libraries and step them in lockstep, comparing
 * b2World_GetStateHash (include/box2d/box2d.h:235), the idiom of the reference's test/test_snapshot.c:258-283.
 *
 * Scenes: the reference's own closed-form benchmark builders (shared/benchmarks.c, shared/determinism.c: no RNG
 * anywhere) plus a few small scenes written here that reach the parts of the solver the benchmarks do not:
 * every joint type, restitution, rolling resistance, conveyor belts, kinematic bodies, contact softening,
 * hit events, motion locks, the overflow colour.
 */
#include "box2d/box2d.h"
#include "box2d/math_functions.h"
#include "box2d/types.h"

#include "benchmarks.h"
#include "determinism.h"

#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define B2H_API __attribute__( ( visibility( "default" ) ) )
#define B2H_MAX_WORLDS 8192 /* the library's own limit (B2_MAX_WORLDS) decides how many can really be created */

typedef struct b2hWorld
{
	b2WorldId worldId;
	int inUse;
	int stepIndex;
	int subStepCount;
	float timeStep;
	float ( *stepFcn )( b2WorldId, int );
	FallingHingeData hinges;
	int hasHinges;
	b2BodyId kinematicId; // moved every step in contact_zoo
	int hasKinematic;
	// the "mutator" scene: handles the per-step callback pokes through the public API
	b2BodyId mutBodies[64];
	b2ShapeId mutShapes[64];
	b2JointId mutJoints[8];
	int mutBodyCount, mutJointCount;
} b2hWorld;

static b2hWorld s_worlds[B2H_MAX_WORLDS];

/* Variant of the scene being created (b2h_create_variant): worlds of a batch differ by an offset derived from their index,
 * never from an RNG.  0 = the scene as the reference's benchmarks build it. */
static int s_variant = 0;

/* ---- small scenes ------------------------------------------------------------------------------------------ */

static void b2hGround( b2WorldId worldId, float halfWidth )
{
	b2BodyDef bodyDef = b2DefaultBodyDef();
	b2BodyId groundId = b2CreateBody( worldId, &bodyDef );
	b2ShapeDef shapeDef = b2DefaultShapeDef();
	b2Segment segment = { { -halfWidth, 0.0f }, { halfWidth, 0.0f } };
	b2CreateSegmentShape( groundId, &shapeDef, &segment );
}

/* one base-N pyramid of boxes on a static segment */
static void b2hCreatePyramid( b2WorldId worldId, int baseCount, float extent )
{
	b2World_EnableSleeping( worldId, false );
	b2hGround( worldId, 20.0f );

	b2BodyDef bodyDef = b2DefaultBodyDef();
	bodyDef.type = b2_dynamicBody;
	b2ShapeDef shapeDef = b2DefaultShapeDef();
	b2Polygon box = b2MakeSquare( extent );
	float centerX = -extent * baseCount;

	for ( int i = 0; i < baseCount; ++i )
	{
		float y = ( 2.0f * i + 1.0f ) * extent;
		for ( int j = i; j < baseCount; ++j )
		{
			float x = ( i + 1.0f ) * extent + 2.0f * ( j - i ) * extent + centerX - 0.5f;
			if ( s_variant != 0 )
			{
				// every world of a batch gets its own, slightly different pile
				x += 0.004f * (float)( ( s_variant * 7 + i * 3 + j * 5 ) % 11 - 5 ) + 0.0002f * (float)( s_variant % 97 );
				bodyDef.linearVelocity = (b2Vec2){ 0.05f * (float)( ( s_variant + i + 2 * j ) % 5 - 2 ), 0.0f };
			}
			bodyDef.position = (b2Pos){ x, y };
			b2BodyId bodyId = b2CreateBody( worldId, &bodyDef );
			b2CreatePolygonShape( bodyId, &shapeDef, &box );
		}
	}
}

static b2BodyId b2hBox( b2WorldId worldId, b2BodyType type, float x, float y, float hx, float hy, float angle )
{
	b2BodyDef bodyDef = b2DefaultBodyDef();
	bodyDef.type = type;
	bodyDef.position = (b2Pos){ x, y };
	bodyDef.rotation = b2MakeRot( angle );
	b2BodyId bodyId = b2CreateBody( worldId, &bodyDef );
	b2ShapeDef shapeDef = b2DefaultShapeDef();
	b2Polygon box = b2MakeBox( hx, hy );
	b2CreatePolygonShape( bodyId, &shapeDef, &box );
	return bodyId;
}

/* Every joint type, with springs, limits, motors and event thresholds, hanging off static and dynamic bodies. */
static void b2hCreateJointZoo( b2WorldId worldId )
{
	b2World_EnableSleeping( worldId, false );
	b2hGround( worldId, 40.0f );

	b2BodyDef groundDef = b2DefaultBodyDef();
	groundDef.position = (b2Pos){ 0.0f, 12.0f };
	b2BodyId anchorId = b2CreateBody( worldId, &groundDef );

	for ( int row = 0; row < 3; ++row )
	{
		float x0 = -16.0f;
		float y0 = 4.0f + 2.5f * row;
		float tilt = 0.15f * row;
		b2BodyId b[10];
		for ( int i = 0; i < 10; ++i )
		{
			b[i] = b2hBox( worldId, b2_dynamicBody, x0 + 1.6f * i, y0 + 0.1f * i, 0.5f, 0.25f + 0.02f * i, tilt - 0.05f * i );
		}

		/* revolute to the static anchor, spring + motor + limit */
		b2RevoluteJointDef rev = b2DefaultRevoluteJointDef();
		rev.base.bodyIdA = anchorId;
		rev.base.bodyIdB = b[0];
		rev.base.localFrameA.p = (b2Vec2){ x0 - 0.5f, y0 - 12.0f };
		rev.base.localFrameB.p = (b2Vec2){ -0.5f, 0.0f };
		rev.enableLimit = true;
		rev.lowerAngle = -0.3f * B2_PI;
		rev.upperAngle = 0.25f * B2_PI;
		rev.enableSpring = true;
		rev.hertz = 1.5f + row;
		rev.dampingRatio = 0.3f;
		rev.targetAngle = 0.2f;
		rev.enableMotor = true;
		rev.motorSpeed = 0.5f - row;
		rev.maxMotorTorque = 3.0f;
		b2JointId revId = b2CreateRevoluteJoint( worldId, &rev );
		b2Joint_SetForceThreshold( revId, 5.0f );
		b2Joint_SetTorqueThreshold( revId, 1.0f );

		/* plain revolute between dynamic bodies */
		b2RevoluteJointDef rev2 = b2DefaultRevoluteJointDef();
		rev2.base.bodyIdA = b[0];
		rev2.base.bodyIdB = b[1];
		rev2.base.localFrameA.p = (b2Vec2){ 0.5f, 0.0f };
		rev2.base.localFrameB.p = (b2Vec2){ -1.1f, 0.0f };
		rev2.base.localFrameB.q = b2MakeRot( 0.3f );
		b2CreateRevoluteJoint( worldId, &rev2 );

		/* distance: spring + limit + motor on row 0/2, rigid on row 1 */
		b2DistanceJointDef dist = b2DefaultDistanceJointDef();
		dist.base.bodyIdA = b[1];
		dist.base.bodyIdB = b[2];
		dist.base.localFrameA.p = (b2Vec2){ 0.4f, 0.1f };
		dist.base.localFrameB.p = (b2Vec2){ -0.4f, -0.1f };
		dist.length = 1.0f;
		if ( row != 1 )
		{
			dist.enableSpring = true;
			dist.hertz = 3.0f;
			dist.dampingRatio = 0.4f;
			dist.lowerSpringForce = -40.0f;
			dist.upperSpringForce = 50.0f;
			dist.enableLimit = true;
			dist.minLength = 0.6f;
			dist.maxLength = 1.4f;
			dist.enableMotor = row == 0;
			dist.motorSpeed = 0.3f;
			dist.maxMotorForce = 5.0f;
		}
		b2JointId distId = b2CreateDistanceJoint( worldId, &dist );
		b2Joint_SetForceThreshold( distId, 0.0f ); /* zero threshold: every awake joint reports */

		/* prismatic: spring + limit + motor */
		b2PrismaticJointDef pris = b2DefaultPrismaticJointDef();
		pris.base.bodyIdA = b[2];
		pris.base.bodyIdB = b[3];
		pris.base.localFrameA.p = (b2Vec2){ 0.6f, 0.0f };
		pris.base.localFrameA.q = b2MakeRot( 0.2f * row );
		pris.base.localFrameB.p = (b2Vec2){ -0.9f, 0.0f };
		pris.enableSpring = row != 2;
		pris.hertz = 2.0f;
		pris.dampingRatio = 0.5f;
		pris.targetTranslation = 0.1f;
		pris.enableLimit = true;
		pris.lowerTranslation = -0.25f;
		pris.upperTranslation = 0.35f;
		pris.enableMotor = row != 1;
		pris.motorSpeed = 0.2f;
		pris.maxMotorForce = 8.0f;
		b2JointId prisId = b2CreatePrismaticJoint( worldId, &pris );
		b2Joint_SetTorqueThreshold( prisId, 0.5f );

		/* wheel: spring + limit + motor */
		b2WheelJointDef wheel = b2DefaultWheelJointDef();
		wheel.base.bodyIdA = b[3];
		wheel.base.bodyIdB = b[4];
		wheel.base.localFrameA.p = (b2Vec2){ 0.7f, 0.0f };
		wheel.base.localFrameA.q = b2MakeRot( 0.5f * B2_PI );
		wheel.base.localFrameB.p = (b2Vec2){ -0.9f, 0.0f };
		wheel.enableSpring = true;
		wheel.hertz = 4.0f;
		wheel.dampingRatio = 0.7f;
		wheel.enableLimit = true;
		wheel.lowerTranslation = -0.3f;
		wheel.upperTranslation = 0.3f;
		wheel.enableMotor = true;
		wheel.motorSpeed = 1.0f + row;
		wheel.maxMotorTorque = 6.0f;
		b2JointId wheelId = b2CreateWheelJoint( worldId, &wheel );
		b2Joint_SetForceThreshold( wheelId, 2.0f );

		/* weld: soft on row 0, rigid otherwise */
		b2WeldJointDef weld = b2DefaultWeldJointDef();
		weld.base.bodyIdA = b[4];
		weld.base.bodyIdB = b[5];
		weld.base.localFrameA.p = (b2Vec2){ 0.8f, 0.0f };
		weld.base.localFrameB.p = (b2Vec2){ -0.8f, 0.0f };
		if ( row == 0 )
		{
			weld.linearHertz = 5.0f;
			weld.linearDampingRatio = 0.6f;
			weld.angularHertz = 4.0f;
			weld.angularDampingRatio = 0.5f;
		}
		b2JointId weldId = b2CreateWeldJoint( worldId, &weld );
		b2Joint_SetForceThreshold( weldId, 1.0f );
		b2Joint_SetTorqueThreshold( weldId, 1.0f );

		/* motor: velocity + springs */
		b2MotorJointDef motor = b2DefaultMotorJointDef();
		motor.base.bodyIdA = b[5];
		motor.base.bodyIdB = b[6];
		motor.base.localFrameA.p = (b2Vec2){ 1.6f, 0.0f };
		motor.linearVelocity = (b2Vec2){ 0.1f, -0.05f };
		motor.angularVelocity = 0.2f;
		motor.maxVelocityForce = row == 1 ? 0.0f : 10.0f;
		motor.maxVelocityTorque = 10.0f;
		motor.linearHertz = 2.0f;
		motor.linearDampingRatio = 0.5f;
		motor.angularHertz = row == 2 ? 0.0f : 2.0f;
		motor.angularDampingRatio = 0.5f;
		motor.maxSpringForce = 20.0f;
		motor.maxSpringTorque = 20.0f;
		b2JointId motorId = b2CreateMotorJoint( worldId, &motor );
		b2Joint_SetForceThreshold( motorId, 3.0f );

		/* mover */
		b2MoverJointDef mover = b2DefaultMoverJointDef();
		mover.base.bodyIdA = anchorId;
		mover.base.bodyIdB = b[7];
		mover.linearVelocity = (b2Vec2){ 0.5f - 0.5f * row, 0.25f };
		mover.maxVelocityForce = (b2Vec2){ 30.0f, row == 2 ? 0.0f : 15.0f };
		b2JointId moverId = b2CreateMoverJoint( worldId, &mover );
		b2Joint_SetForceThreshold( moverId, 1.0f );

		/* pogo */
		b2PogoJointDef pogo = b2DefaultPogoJointDef();
		pogo.base.bodyIdA = anchorId;
		pogo.base.bodyIdB = b[8];
		pogo.base.localFrameA.p = (b2Vec2){ x0 + 1.6f * 8.0f, y0 - 12.0f - 1.0f };
		pogo.normal = (b2Vec2){ 0.0f, 1.0f };
		pogo.hertz = row == 1 ? 0.0f : 3.0f;
		pogo.dampingRatio = 0.5f;
		pogo.restLength = 1.0f;
		pogo.maxTensionForce = 10.0f;
		pogo.maxCompressionForce = 200.0f;
		b2JointId pogoId = b2CreatePogoJoint( worldId, &pogo );
		b2Joint_SetForceThreshold( pogoId, 1.0f );

		/* filter joint: a no-op in the solver, still occupies a slot in its colour */
		b2FilterJointDef filter = b2DefaultFilterJointDef();
		filter.base.bodyIdA = b[8];
		filter.base.bodyIdB = b[9];
		b2CreateFilterJoint( worldId, &filter );

		/* weld with a fixed-rotation body: iA + iB paths with zero inertia */
		b2Body_SetMotionLocks( b[9], (b2MotionLocks){ false, false, true } );
		b2WeldJointDef weld2 = b2DefaultWeldJointDef();
		weld2.base.bodyIdA = b[7];
		weld2.base.bodyIdB = b[9];
		weld2.base.localFrameA.p = (b2Vec2){ 1.6f, 0.0f };
		weld2.base.localFrameB.p = (b2Vec2){ -1.6f, 0.0f };
		weld2.linearHertz = 2.0f;
		weld2.angularHertz = 2.0f;
		b2CreateWeldJoint( worldId, &weld2 );
	}
}

/* Contacts off the beaten path of the benchmark scenes. */
static void b2hCreateContactZoo( b2WorldId worldId, b2hWorld* w )
{
	b2World_EnableSleeping( worldId, false );
	b2World_SetHitEventThreshold( worldId, 0.5f );

	/* ground: a conveyor belt box + a bouncy slope */
	{
		b2BodyDef bodyDef = b2DefaultBodyDef();
		b2BodyId groundId = b2CreateBody( worldId, &bodyDef );
		b2ShapeDef shapeDef = b2DefaultShapeDef();
		shapeDef.material.tangentSpeed = 1.5f;
		shapeDef.material.friction = 0.8f;
		b2Polygon belt = b2MakeOffsetBox( 12.0f, 0.5f, (b2Vec2){ 0.0f, -0.5f }, b2Rot_identity );
		b2CreatePolygonShape( groundId, &shapeDef, &belt );

		b2ShapeDef bouncy = b2DefaultShapeDef();
		bouncy.material.restitution = 0.7f;
		b2Segment slope = { { 12.0f, 0.0f }, { 24.0f, 4.0f } };
		b2CreateSegmentShape( groundId, &bouncy, &slope );
		b2Segment wall = { { -12.0f, 0.0f }, { -12.0f, 10.0f } };
		b2CreateSegmentShape( groundId, &bouncy, &wall );
	}

	/* kinematic platform, moved by its velocity; several bodies ride on it */
	{
		b2BodyDef bodyDef = b2DefaultBodyDef();
		bodyDef.type = b2_kinematicBody;
		bodyDef.position = (b2Pos){ -6.0f, 3.0f };
		bodyDef.linearVelocity = (b2Vec2){ 0.6f, 0.0f };
		bodyDef.angularVelocity = 0.05f;
		w->kinematicId = b2CreateBody( worldId, &bodyDef );
		w->hasKinematic = 1;
		b2ShapeDef shapeDef = b2DefaultShapeDef();
		b2Polygon box = b2MakeBox( 3.0f, 0.25f );
		b2CreatePolygonShape( w->kinematicId, &shapeDef, &box );
	}

	b2ShapeDef shapeDef = b2DefaultShapeDef();
	shapeDef.enableHitEvents = true;

	for ( int i = 0; i < 48; ++i )
	{
		b2BodyDef bodyDef = b2DefaultBodyDef();
		bodyDef.type = b2_dynamicBody;
		float x = -9.0f + 0.45f * i;
		float y = 4.0f + 0.9f * ( i % 7 );
		bodyDef.position = (b2Pos){ x, y };
		bodyDef.rotation = b2MakeRot( 0.37f * i );
		bodyDef.linearVelocity = (b2Vec2){ 0.3f * ( i % 5 ) - 0.6f, -2.0f - 0.5f * ( i % 3 ) };
		bodyDef.angularVelocity = 0.4f * ( i % 4 ) - 0.6f;
		bodyDef.linearDamping = 0.05f * ( i % 3 );
		bodyDef.angularDamping = 0.1f * ( i % 2 );
		bodyDef.gravityScale = 1.0f + 0.1f * ( i % 4 );
		if ( i % 11 == 0 )
		{
			bodyDef.motionLocks = (b2MotionLocks){ false, false, true };
		}
		if ( i % 13 == 5 )
		{
			bodyDef.motionLocks = (b2MotionLocks){ true, false, false };
		}
		b2BodyId bodyId = b2CreateBody( worldId, &bodyDef );

		shapeDef.material.restitution = ( i % 3 == 0 ) ? 0.6f : 0.0f;
		shapeDef.material.rollingResistance = ( i % 4 == 1 ) ? 0.2f : 0.0f;
		shapeDef.material.friction = 0.2f + 0.1f * ( i % 6 );
		shapeDef.density = 0.5f + 0.5f * ( i % 5 );

		switch ( i % 3 )
		{
			case 0:
			{
				b2Circle circle = { { 0.0f, 0.0f }, 0.2f + 0.02f * ( i % 5 ) };
				b2CreateCircleShape( bodyId, &shapeDef, &circle );
			}
			break;
			case 1:
			{
				b2Capsule capsule = { { -0.2f, 0.0f }, { 0.2f, 0.05f }, 0.15f };
				b2CreateCapsuleShape( bodyId, &shapeDef, &capsule );
			}
			break;
			default:
			{
				b2Polygon box = b2MakeBox( 0.2f + 0.01f * ( i % 4 ), 0.15f );
				b2CreatePolygonShape( bodyId, &shapeDef, &box );
			}
			break;
		}
		if ( i % 9 == 4 )
		{
			b2Body_ApplyForceToCenter( bodyId, (b2Vec2){ 3.0f, 1.0f }, true );
			b2Body_ApplyTorque( bodyId, 0.5f, true );
		}
	}

	/* one very fast body: trips the linear and angular speed caps of integrate-positions */
	{
		b2BodyDef bodyDef = b2DefaultBodyDef();
		bodyDef.type = b2_dynamicBody;
		bodyDef.position = (b2Pos){ 0.0f, 30.0f };
		bodyDef.linearVelocity = (b2Vec2){ 3000.0f, -4500.0f };
		bodyDef.angularVelocity = 500.0f;
		b2BodyId bodyId = b2CreateBody( worldId, &bodyDef );
		b2Circle circle = { { 0.0f, 0.0f }, 0.25f };
		b2CreateCircleShape( bodyId, &shapeDef, &circle );
	}
}

/* One heavy dynamic tray carrying many small boxes, and a hub with many spokes: more constraints on one body
 * than there are dynamic colours, so the graph overflows (src/constraint_graph.c:66-133, :216). */
static void b2hCreateOverflow( b2WorldId worldId )
{
	b2World_EnableSleeping( worldId, false );
	b2hGround( worldId, 40.0f );

	b2BodyId trayId = b2hBox( worldId, b2_dynamicBody, 0.0f, 1.0f, 8.0f, 0.25f, 0.0f );
	(void)trayId;
	for ( int i = 0; i < 40; ++i )
	{
		b2BodyDef bodyDef = b2DefaultBodyDef();
		bodyDef.type = b2_dynamicBody;
		bodyDef.position = (b2Pos){ -7.6f + 0.39f * i, 1.45f };
		b2BodyId bodyId = b2CreateBody( worldId, &bodyDef );
		b2ShapeDef shapeDef = b2DefaultShapeDef();
		shapeDef.material.restitution = ( i % 5 == 0 ) ? 0.4f : 0.0f;
		shapeDef.material.rollingResistance = ( i % 7 == 0 ) ? 0.1f : 0.0f;
		if ( i % 2 == 0 )
		{
			b2Polygon box = b2MakeBox( 0.18f, 0.18f );
			b2CreatePolygonShape( bodyId, &shapeDef, &box );
		}
		else
		{
			b2Circle circle = { { 0.0f, 0.0f }, 0.18f };
			b2CreateCircleShape( bodyId, &shapeDef, &circle );
		}
	}

	b2BodyDef hubDef = b2DefaultBodyDef();
	hubDef.type = b2_dynamicBody;
	hubDef.position = (b2Pos){ 20.0f, 8.0f };
	b2BodyId hubId = b2CreateBody( worldId, &hubDef );
	b2ShapeDef hubShape = b2DefaultShapeDef();
	b2Circle hubCircle = { { 0.0f, 0.0f }, 0.5f };
	b2CreateCircleShape( hubId, &hubShape, &hubCircle );
	for ( int i = 0; i < 30; ++i )
	{
		float angle = 2.0f * B2_PI * i / 30.0f;
		b2CosSin cs = b2ComputeCosSin( angle );
		b2BodyDef bodyDef = b2DefaultBodyDef();
		bodyDef.type = b2_dynamicBody;
		bodyDef.position = (b2Pos){ 20.0f + 2.0f * cs.cosine, 8.0f + 2.0f * cs.sine };
		b2BodyId bodyId = b2CreateBody( worldId, &bodyDef );
		b2ShapeDef shapeDef = b2DefaultShapeDef();
		shapeDef.filter.groupIndex = -1;
		b2Circle circle = { { 0.0f, 0.0f }, 0.15f };
		b2CreateCircleShape( bodyId, &shapeDef, &circle );

		b2RevoluteJointDef rev = b2DefaultRevoluteJointDef();
		rev.base.bodyIdA = hubId;
		rev.base.bodyIdB = bodyId;
		rev.base.localFrameA.p = (b2Vec2){ 0.5f * cs.cosine, 0.5f * cs.sine };
		rev.base.localFrameB.p = (b2Vec2){ -1.5f * cs.cosine, -1.5f * cs.sine };
		rev.enableSpring = ( i % 2 ) == 0;
		rev.hertz = 2.0f;
		rev.dampingRatio = 0.2f;
		b2JointId jointId = b2CreateRevoluteJoint( worldId, &rev );
		b2Joint_SetForceThreshold( jointId, 0.5f );
	}
}

/* A base-9 pyramid next to a chain of boxes on revolute joints, and a per-step callback that changes the world BETWEEN steps
 * through the public API the way an application does: gravity, velocities, impulses and forces, masses, friction and
 * restitution, joint motors and springs, contact and joint tuning, warm starting, teleports, body types, disabling, destroying
 * and creating bodies, the sub-step count.  Whatever a solver keeps from one step to the next must notice every one of these. */
static void b2hCreateMutator( b2WorldId worldId, b2hWorld* w )
{
	b2World_EnableSleeping( worldId, false );
	b2hGround( worldId, 40.0f );
	b2ShapeDef shapeDef = b2DefaultShapeDef();
	shapeDef.enableHitEvents = true;
	b2Polygon box = b2MakeSquare( 0.5f );
	int n = 0;
	for ( int i = 0; i < 9; ++i )
	{
		for ( int j = i; j < 9; ++j )
		{
			b2BodyDef bodyDef = b2DefaultBodyDef();
			bodyDef.type = b2_dynamicBody;
			bodyDef.position = (b2Pos){ -12.0f + ( i + 1.0f ) * 0.5f + ( j - i ) * 1.0f, ( 2.0f * i + 1.0f ) * 0.5f };
			b2BodyId bodyId = b2CreateBody( worldId, &bodyDef );
			b2ShapeId shapeId = b2CreatePolygonShape( bodyId, &shapeDef, &box );
			if ( n < 48 )
			{
				w->mutBodies[n] = bodyId;
				w->mutShapes[n] = shapeId;
				n += 1;
			}
		}
	}
	// a chain hanging from a static anchor, dragging over the ground
	b2BodyDef anchorDef = b2DefaultBodyDef();
	anchorDef.position = (b2Pos){ 6.0f, 6.0f };
	b2BodyId previous = b2CreateBody( worldId, &anchorDef );
	for ( int i = 0; i < 8; ++i )
	{
		b2BodyId link = b2hBox( worldId, b2_dynamicBody, 6.5f + 1.0f * i, 6.0f, 0.5f, 0.125f, 0.0f );
		b2RevoluteJointDef rev = b2DefaultRevoluteJointDef();
		rev.base.bodyIdA = previous;
		rev.base.bodyIdB = link;
		rev.base.localFrameA.p = i == 0 ? (b2Vec2){ 0.0f, 0.0f } : (b2Vec2){ 0.5f, 0.0f };
		rev.base.localFrameB.p = (b2Vec2){ -0.5f, 0.0f };
		rev.maxMotorTorque = 20.0f;
		w->mutJoints[i] = b2CreateRevoluteJoint( worldId, &rev );
		w->mutBodies[n] = link;
		w->mutShapes[n] = (b2ShapeId){ 0 };
		n += 1;
		previous = link;
	}
	w->mutBodyCount = n;
	w->mutJointCount = 8;
}

static b2hWorld* b2hFind( b2WorldId worldId );

static float b2hStepMutator( b2WorldId worldId, int stepIndex )
{
	b2hWorld* w = b2hFind( worldId );
	if ( w == NULL )
	{
		return 0.0f;
	}
	b2BodyId* body = w->mutBodies;
	switch ( stepIndex )
	{
		case 8:
			b2World_SetGravity( worldId, (b2Vec2){ 1.5f, -6.0f } );
			break;
		case 14:
			b2Body_SetLinearVelocity( body[40], (b2Vec2){ 3.0f, 4.0f } );
			b2Body_ApplyAngularImpulse( body[20], 1.5f, true );
			b2Body_ApplyLinearImpulseToCenter( body[3], (b2Vec2){ -2.0f, 0.5f }, true );
			break;
		case 20:
		{
			// a body under persisting contacts gets three times its mass: the contacts remember the old one
			b2MassData massData = b2Body_GetMassData( body[10] );
			massData.mass *= 3.0f;
			massData.rotationalInertia *= 3.0f;
			b2Body_SetMassData( body[10], massData );
			b2Body_SetGravityScale( body[11], 0.25f );
			b2Body_SetLinearDamping( body[12], 0.8f );
			break;
		}
		case 26:
			for ( int i = 0; i < 12; ++i )
			{
				b2Shape_SetFriction( w->mutShapes[2 * i], 0.05f + 0.07f * i );
				b2Shape_SetRestitution( w->mutShapes[2 * i + 1], 0.1f * ( i % 5 ) );
			}
			break;
		case 32:
			b2RevoluteJoint_EnableMotor( w->mutJoints[0], true );
			b2RevoluteJoint_SetMotorSpeed( w->mutJoints[0], 1.5f );
			b2RevoluteJoint_EnableSpring( w->mutJoints[3], true );
			b2RevoluteJoint_SetSpringHertz( w->mutJoints[3], 3.0f );
			b2Joint_SetConstraintTuning( w->mutJoints[5], 30.0f, 1.0f );
			break;
		case 38:
			b2World_SetContactTuning( worldId, 20.0f, 5.0f, 2.0f );
			break;
		case 44:
			b2World_EnableWarmStarting( worldId, false );
			break;
		case 48:
			b2World_EnableWarmStarting( worldId, true );
			b2World_SetGravity( worldId, (b2Vec2){ 0.0f, -10.0f } );
			break;
		case 54:
			b2Body_SetTransform( body[44], (b2Pos){ -3.0f, 9.0f }, b2MakeRot( 0.4f ) );
			b2RevoluteJoint_SetMotorSpeed( w->mutJoints[0], -2.0f );
			break;
		case 60:
			b2Body_SetType( body[30], b2_kinematicBody );
			b2Body_SetLinearVelocity( body[30], (b2Vec2){ 0.5f, 0.0f } );
			b2Body_Disable( body[5] );
			break;
		case 66:
			b2Body_SetType( body[30], b2_dynamicBody );
			b2Body_Enable( body[5] );
			break;
		case 72:
			// swap-removes in the solver arrays and the colour arrays, ids handed out again
			b2DestroyBody( body[7] );
			body[7] = b2hBox( worldId, b2_dynamicBody, -5.0f, 12.0f, 0.5f, 0.5f, 0.2f );
			b2DestroyBody( body[22] );
			body[22] = b2hBox( worldId, b2_dynamicBody, -7.0f, 14.0f, 0.4f, 0.6f, -0.3f );
			break;
		case 78:
			b2World_SetMaximumLinearSpeed( worldId, 3.0f ); /* speed caps: the flag travels in the body state */
			b2Body_SetLinearVelocity( body[41], (b2Vec2){ 30.0f, 10.0f } );
			break;
		case 82:
			b2World_SetMaximumLinearSpeed( worldId, 400.0f );
			b2World_SetRestitutionThreshold( worldId, 0.2f );
			b2World_SetHitEventThreshold( worldId, 0.1f );
			break;
		case 88:
			w->subStepCount = 6;
			break;
		case 94:
			w->subStepCount = 4;
			b2World_EnableSleeping( worldId, true );
			break;
		case 97:
			// a step that does not advance time: the narrow phase runs (and may re-evaluate manifolds), the solver does not
			// (src/physics_world.c:912-950) -- whatever the next solve is told about recycled manifolds must not skip this
			b2Body_SetTransform( body[43], (b2Pos){ -2.0f, 7.0f }, b2MakeRot( -0.2f ) );
			b2World_Step( worldId, 0.0f, 4 );
			break;
		case 112:
			// the application puts an island to sleep itself (src/body.c:1575 -> b2TrySleepIsland) ...
			b2Body_SetAwake( body[15], false );
			break;
		case 116:
		{
			// ... asks for contact data and joint reactions in the middle of the run ...
			b2ContactData data[8];
			(void)b2Body_GetContactData( body[2], data, 8 );
			(void)b2Joint_GetConstraintForce( w->mutJoints[2] );
			break;
		}
		case 119:
			// ... wakes it again, sets off an explosion next to the pile
			b2Body_SetAwake( body[15], true );
			{
				b2ExplosionDef explosion = b2DefaultExplosionDef();
				explosion.position = (b2Pos){ 0.0f, 2.0f };
				explosion.radius = 3.0f;
				explosion.falloff = 1.0f;
				explosion.impulsePerLength = 2.0f;
				b2World_Explode( worldId, &explosion );
			}
			break;
		case 124:
			// ... and takes a joint away
			if ( w->mutJointCount > 1 )
			{
				b2DestroyJoint( w->mutJoints[1] );
			}
			break;
		default:
			break;
	}
	if ( 100 <= stepIndex && stepIndex < 110 )
	{
		b2Body_ApplyForceToCenter( body[1], (b2Vec2){ 40.0f, 0.0f }, true );
		b2Body_ApplyTorque( body[2], 15.0f, true );
	}
	return 0.0f;
}

/* ---- API ----------------------------------------------------------------------------------------------------- */

static b2hWorld* b2hFind( b2WorldId worldId )
{
	for ( int i = 0; i < B2H_MAX_WORLDS; ++i )
	{
		if ( s_worlds[i].inUse && s_worlds[i].worldId.index1 == worldId.index1 && s_worlds[i].worldId.generation == worldId.generation )
		{
			return s_worlds + i;
		}
	}
	return NULL;
}

B2H_API int b2h_create( const char* scene, int workerCount )
{
	int h = -1;
	for ( int i = 0; i < B2H_MAX_WORLDS; ++i )
	{
		if ( s_worlds[i].inUse == 0 )
		{
			h = i;
			break;
		}
	}
	if ( h < 0 )
	{
		return -1;
	}

	b2hWorld* w = s_worlds + h;
	memset( w, 0, sizeof( *w ) );

	b2WorldDef worldDef = b2DefaultWorldDef();
	worldDef.workerCount = workerCount;
	if ( strcmp( scene, "pyramid_soft" ) == 0 )
	{
		worldDef.enableContactSoftening = true;
	}

	w->worldId = b2CreateWorld( &worldDef );
	if ( b2World_IsValid( w->worldId ) == false )
	{
		return -1; /* the library's world array is full (B2_MAX_WORLDS) */
	}
	w->inUse = 1;
	w->timeStep = 1.0f / 60.0f;
	w->subStepCount = 4;

	if ( strcmp( scene, "large_pyramid" ) == 0 )
	{
		CreateLargePyramid( w->worldId );
	}
	else if ( strcmp( scene, "many_pyramids" ) == 0 )
	{
		CreateManyPyramids( w->worldId );
	}
	else if ( strcmp( scene, "joint_grid" ) == 0 )
	{
		CreateJointGrid( w->worldId );
	}
	else if ( strcmp( scene, "rain" ) == 0 )
	{
		CreateRain( w->worldId );
		w->stepFcn = StepRain;
	}
	else if ( strcmp( scene, "tumbler" ) == 0 )
	{
		CreateTumbler( w->worldId );
	}
	else if ( strcmp( scene, "spinner" ) == 0 )
	{
		CreateSpinner( w->worldId );
		w->stepFcn = StepSpinner;
	}
	else if ( strcmp( scene, "smash" ) == 0 )
	{
		CreateSmash( w->worldId );
	}
	else if ( strcmp( scene, "compounds" ) == 0 )
	{
		CreateCompounds( w->worldId );
	}
	else if ( strcmp( scene, "washer" ) == 0 )
	{
		CreateWasher( w->worldId );
	}
	else if ( strcmp( scene, "falling_hinges" ) == 0 )
	{
		w->hinges = CreateFallingHinges( w->worldId );
		w->hasHinges = 1;
	}
	else if ( strcmp( scene, "small_pyramid" ) == 0 || strcmp( scene, "pyramid_soft" ) == 0 )
	{
		b2hCreatePyramid( w->worldId, 10, 0.5f );
	}
	else if ( strcmp( scene, "pyramid_cold" ) == 0 )
	{
		b2hCreatePyramid( w->worldId, 6, 0.5f );
		b2World_EnableWarmStarting( w->worldId, false );
	}
	else if ( strcmp( scene, "joint_zoo" ) == 0 )
	{
		b2hCreateJointZoo( w->worldId );
	}
	else if ( strcmp( scene, "joint_zoo_cold" ) == 0 )
	{
		b2hCreateJointZoo( w->worldId );
		b2World_EnableWarmStarting( w->worldId, false );
	}
	else if ( strcmp( scene, "contact_zoo" ) == 0 )
	{
		b2hCreateContactZoo( w->worldId, w );
	}
	else if ( strcmp( scene, "overflow" ) == 0 )
	{
		b2hCreateOverflow( w->worldId );
	}
	else if ( strcmp( scene, "mutator" ) == 0 )
	{
		b2hCreateMutator( w->worldId, w );
		w->stepFcn = b2hStepMutator;
	}
	else
	{
		b2DestroyWorld( w->worldId );
		w->inUse = 0;
		return -2;
	}
	return h;
}

B2H_API int b2h_create_variant( const char* scene, int workerCount, int variant )
{
	s_variant = variant;
	int h = b2h_create( scene, workerCount );
	s_variant = 0;
	return h;
}

B2H_API void b2h_destroy( int h )
{
	b2hWorld* w = s_worlds + h;
	if ( w->inUse )
	{
		if ( w->hasHinges )
		{
			DestroyFallingHinges( &w->hinges );
		}
		b2DestroyWorld( w->worldId );
		w->inUse = 0;
	}
}

B2H_API void b2h_set_substeps( int h, int subStepCount )
{
	s_worlds[h].subStepCount = subStepCount;
}

/* Step n times exactly like benchmark/main.c:318-356: per-step scene callback first, then b2World_Step. */
B2H_API void b2h_step( int h, int n )
{
	b2hWorld* w = s_worlds + h;
	for ( int i = 0; i < n; ++i )
	{
		if ( w->stepFcn != NULL )
		{
			w->stepFcn( w->worldId, w->stepIndex );
		}
		if ( w->hasKinematic && w->stepIndex % 120 == 119 )
		{
			b2Vec2 v = b2Body_GetLinearVelocity( w->kinematicId );
			b2Body_SetLinearVelocity( w->kinematicId, (b2Vec2){ -v.x, v.y } );
		}
		b2World_Step( w->worldId, w->timeStep, w->subStepCount );
		if ( w->hasHinges )
		{
			UpdateFallingHinges( w->worldId, &w->hinges );
		}
		w->stepIndex += 1;
	}
}

/* Step many worlds concurrently, one thread per world, `n` steps each (include/box2d/box2d.h:31-32).  In front of the seam
 * with the worlds in a group (b2GpuSeam_CreateGroup) every round of steps is one batched solve; in front of the CPU solver
 * the worlds simply run side by side. */
typedef struct b2hStepJob
{
	int handle;
	int steps;
} b2hStepJob;

static void* b2hStepThread( void* arg )
{
	b2hStepJob* job = arg;
	b2h_step( job->handle, job->steps );
	return NULL;
}

B2H_API int b2h_step_many( const int* handles, int count, int steps )
{
	pthread_t* threads = malloc( (size_t)count * sizeof( pthread_t ) );
	b2hStepJob* jobs = malloc( (size_t)count * sizeof( b2hStepJob ) );
	pthread_attr_t attr;
	pthread_attr_init( &attr );
	pthread_attr_setstacksize( &attr, 1 << 20 );
	int started = 0;
	for ( int i = 0; i < count; ++i )
	{
		jobs[i].handle = handles[i];
		jobs[i].steps = steps;
		if ( pthread_create( threads + i, &attr, b2hStepThread, jobs + i ) != 0 )
		{
			break;
		}
		started += 1;
	}
	for ( int i = 0; i < started; ++i )
	{
		pthread_join( threads[i], NULL );
	}
	pthread_attr_destroy( &attr );
	free( threads );
	free( jobs );
	return started == count ? 0 : 1;
}

B2H_API uint64_t b2h_hash( int h )
{
	return b2World_GetStateHash( s_worlds[h].worldId );
}

/* Checksum of what b2Shape_GetContactData (include/box2d/box2d.h:841) reports for every shape of the world: the manifolds'
 * impulses as an application sees them.  Order independent (a sum of per-contact hashes). */
typedef struct b2hContactSum
{
	uint64_t sum;
	int contacts;
} b2hContactSum;

static bool b2hContactSumShape( b2ShapeId shapeId, void* context )
{
	b2hContactSum* acc = context;
	b2ContactData data[64];
	int count = b2Shape_GetContactData( shapeId, data, 64 );
	for ( int i = 0; i < count; ++i )
	{
		const b2Manifold* m = &data[i].manifold;
		uint64_t h = 1469598103934665603ull;
		float values[10] = { m->rollingImpulse };
		for ( int p = 0; p < m->pointCount && p < 2; ++p )
		{
			values[1 + 4 * p] = m->points[p].normalImpulse;
			values[2 + 4 * p] = m->points[p].tangentImpulse;
			values[3 + 4 * p] = m->points[p].totalNormalImpulse;
			values[4 + 4 * p] = m->points[p].normalVelocity;
		}
		values[9] = (float)m->pointCount;
		const uint8_t* bytes = (const uint8_t*)values;
		for ( size_t k = 0; k < sizeof( values ); ++k )
		{
			h = ( h ^ bytes[k] ) * 1099511628211ull;
		}
		acc->sum += h;
		acc->contacts += 1;
	}
	return true;
}

B2H_API uint64_t b2h_contact_checksum( int h, int* contacts )
{
	b2hContactSum acc = { 0, 0 };
	b2AABB everything = { { -1.0e6f, -1.0e6f }, { 1.0e6f, 1.0e6f } };
	b2World_OverlapAABB( s_worlds[h].worldId, (b2Pos){ 0.0f, 0.0f }, everything, b2DefaultQueryFilter(), b2hContactSumShape, &acc );
	if ( contacts != NULL )
	{
		*contacts = acc.contacts;
	}
	return acc.sum;
}

/* Reactions of the mutator scene's joints as the application sees them: b2Joint_GetConstraintForce / Torque
 * (include/box2d/box2d.h:964-967) of every joint, b2RevoluteJoint_GetMotorTorque of the first.  Returns the number of floats. */
B2H_API int b2h_mut_joint_reactions( int h, float* out, int capacity )
{
	b2hWorld* w = s_worlds + h;
	int n = 0;
	for ( int i = 0; i < w->mutJointCount && n + 3 <= capacity; ++i )
	{
		if ( b2Joint_IsValid( w->mutJoints[i] ) == false )
		{
			continue;
		}
		b2Vec2 force = b2Joint_GetConstraintForce( w->mutJoints[i] );
		out[n++] = force.x;
		out[n++] = force.y;
		out[n++] = b2Joint_GetConstraintTorque( w->mutJoints[i] );
	}
	if ( w->mutJointCount > 0 && n < capacity && b2Joint_IsValid( w->mutJoints[0] ) && b2Joint_GetType( w->mutJoints[0] ) == b2_revoluteJoint )
	{
		out[n++] = b2RevoluteJoint_GetMotorTorque( w->mutJoints[0] );
	}
	return n;
}

/* b2World_Snapshot / b2World_Restore (include/box2d/box2d.h:316,329) */
B2H_API int b2h_snapshot( int h, uint8_t* image, int capacity )
{
	return b2World_Snapshot( s_worlds[h].worldId, image, capacity );
}

B2H_API int b2h_restore( int h, const uint8_t* image, int size, int stepIndex )
{
	s_worlds[h].stepIndex = stepIndex;
	return b2World_Restore( s_worlds[h].worldId, image, size ) ? 1 : 0;
}

B2H_API int b2h_step_index( int h )
{
	return s_worlds[h].stepIndex;
}

B2H_API int b2h_world_index( int h )
{
	return (int)s_worlds[h].worldId.index1 - 1;
}

/* FallingHinges golden (reference test/test_determinism.c:22-23): sleepStep and transform hash once asleep. */
B2H_API int b2h_hinges_result( int h, int* sleepStep, uint32_t* hash )
{
	b2hWorld* w = s_worlds + h;
	if ( w->hasHinges == 0 )
	{
		return -1;
	}
	*sleepStep = w->hinges.sleepStep;
	*hash = w->hinges.hash;
	return w->hinges.hash != 0 ? 1 : 0;
}

/* out[0..22] = b2Profile as floats, in declaration order (include/box2d/types.h:526-551) */
B2H_API void b2h_profile( int h, float* out )
{
	b2Profile p = b2World_GetProfile( s_worlds[h].worldId );
	memcpy( out, &p, sizeof( p ) );
}

B2H_API int b2h_profile_float_count( void )
{
	return (int)( sizeof( b2Profile ) / sizeof( float ) );
}

/* out: bodyCount, shapeCount, contactCount, jointCount, islandCount, awakeBodyCount, then colorCounts[24] */
B2H_API void b2h_counters( int h, int* out )
{
	b2Counters c = b2World_GetCounters( s_worlds[h].worldId );
	out[0] = c.bodyCount;
	out[1] = c.shapeCount;
	out[2] = c.contactCount;
	out[3] = c.jointCount;
	out[4] = c.islandCount;
	out[5] = b2World_GetAwakeBodyCount( s_worlds[h].worldId );
	for ( int i = 0; i < 24; ++i )
	{
		out[6 + i] = c.colorCounts[i];
	}
}

/* Contacts of the last step whose manifold the narrow phase recycled (src/physics_world.c:508-560). */
B2H_API int b2h_recycled( int h )
{
	return b2World_GetCounters( s_worlds[h].worldId ).recycledContactCount;
}

/* Event counts of the last step: move, begin-touch, end-touch, hit, joint events. */
B2H_API void b2h_event_counts( int h, int* out )
{
	b2WorldId id = s_worlds[h].worldId;
	b2BodyEvents be = b2World_GetBodyEvents( id );
	b2ContactEvents ce = b2World_GetContactEvents( id );
	b2JointEvents je = b2World_GetJointEvents( id );
	out[0] = be.moveCount;
	out[1] = ce.beginCount;
	out[2] = ce.endCount;
	out[3] = ce.hitCount;
	out[4] = je.count;
}

/* Per awake body of the last step's move events: px py qc qs (4 floats), for the tolerance fallback check. */
B2H_API int b2h_move_transforms( int h, float* out, int maxBodies )
{
	b2BodyEvents be = b2World_GetBodyEvents( s_worlds[h].worldId );
	int n = be.moveCount < maxBodies ? be.moveCount : maxBodies;
	for ( int i = 0; i < n; ++i )
	{
		b2WorldTransform t = be.moveEvents[i].transform;
		out[4 * i + 0] = (float)t.p.x;
		out[4 * i + 1] = (float)t.p.y;
		out[4 * i + 2] = t.q.c;
		out[4 * i + 3] = t.q.s;
	}
	return n;
}

/* Body velocities in move-event order: vx vy w (3 floats) */
B2H_API int b2h_move_velocities( int h, float* out, int maxBodies )
{
	b2BodyEvents be = b2World_GetBodyEvents( s_worlds[h].worldId );
	int n = be.moveCount < maxBodies ? be.moveCount : maxBodies;
	for ( int i = 0; i < n; ++i )
	{
		b2BodyId id = be.moveEvents[i].bodyId;
		b2Vec2 v = b2Body_GetLinearVelocity( id );
		out[3 * i + 0] = v.x;
		out[3 * i + 1] = v.y;
		out[3 * i + 2] = b2Body_GetAngularVelocity( id );
	}
	return n;
}

/* Wall-clock benchmark loop like benchmark/main.c:336-356: returns total ms of n steps and accumulates the
 * solver's b2Profile.constraints (the hot path) into *constraintsMs and the whole step into *stepMs. */
B2H_API float b2h_bench( int h, int n, float* constraintsMs, float* stepMs, float* solverStageMs )
{
	b2hWorld* w = s_worlds + h;
	uint64_t ticks = b2GetTicks();
	float constraints = 0.0f, step = 0.0f;
	float stages[8] = { 0 };
	for ( int i = 0; i < n; ++i )
	{
		if ( w->stepFcn != NULL )
		{
			w->stepFcn( w->worldId, w->stepIndex );
		}
		b2World_Step( w->worldId, w->timeStep, w->subStepCount );
		w->stepIndex += 1;
		b2Profile p = b2World_GetProfile( w->worldId );
		constraints += p.constraints;
		step += p.step;
		stages[0] += p.prepareConstraints;
		stages[1] += p.integrateVelocities;
		stages[2] += p.warmStart;
		stages[3] += p.solveImpulses;
		stages[4] += p.integratePositions;
		stages[5] += p.relaxImpulses;
		stages[6] += p.applyRestitution;
		stages[7] += p.storeImpulses;
	}
	float ms = b2GetMilliseconds( ticks );
	*constraintsMs = constraints;
	*stepMs = step;
	if ( solverStageMs != NULL )
	{
		memcpy( solverStageMs, stages, sizeof( stages ) );
	}
	return ms;
}

B2H_API int b2h_version( void )
{
	b2Version v = b2GetVersion();
	return v.major * 10000 + v.minor * 100 + v.revision;
}
