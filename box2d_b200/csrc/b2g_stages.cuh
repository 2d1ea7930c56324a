// b2g_stages.cuh -- body stages and the per-stage work distribution shared by the persistent step kernel
// and the one-launch-per-stage debug path.
//
// Work distribution: items are cut into chunks of 32 (one warp each).  Chunk k belongs to block
// (k mod gridDim) and warp (k div gridDim) of that block, so consecutive chunks land on different SMs and a
// colour with n constraints occupies ceil(n/32) SMs with ONE warp each before any SM gets a second warp:
// the stages are latency bound (SURVEY.md section 7 "hard parts"), so spreading beats packing.  The mapping is
// static, i.e. the same thread owns the same constraint in every stage of the step.
#pragma once

#include "b2g_joint.cuh"

namespace b2g
{

constexpr float kMaxRotation = 0.25f * kPi; // B2_MAX_ROTATION, include/box2d/constants.h:50

// ---- body stages --------------------------------------------------------------------------------------------

// AoS -> SoA + per-step body constants.  The velocity increment of b2IntegrateVelocitiesTask
// (src/solver.c:94-102) depends only on per-step constants, so it is evaluated once here with the same
// operations and reused by every sub-step.
B2G_DEV void loadBody( const StepParams& P, int i )
{
	const uint8_t* s = P.rawStates + (size_t)i * B2L_STATE_SIZE;
	const float4* s4 = reinterpret_cast<const float4*>( s );
	float4 v = s4[0];
	float4 p = s4[1];
	__stcg( P.vel + i + 1, v );
	__stcg( P.pos + i + 1, p );

	const uint8_t* sim = P.rawSims + (size_t)i * B2L_SIM_SIZE;
	float invMass = rawF( sim, B2L_SIM_INV_MASS );
	float invInertia = rawF( sim, B2L_SIM_INV_INERTIA );
	V2 force = v2( rawF( sim, B2L_SIM_FORCE ), rawF( sim, B2L_SIM_FORCE + 4 ) );
	float torque = rawF( sim, B2L_SIM_TORQUE );
	float h = P.h;

	float linearDamping = 1.0f / ( 1.0f + h * rawF( sim, B2L_SIM_LINEAR_DAMPING ) );
	float angularDamping = 1.0f / ( 1.0f + h * rawF( sim, B2L_SIM_ANGULAR_DAMPING ) );

	// gravity scale will be zero for kinematic bodies
	float gravityScale = invMass > 0.0f ? rawF( sim, B2L_SIM_GRAVITY_SCALE ) : 0.0f;

	V2 linearVelocityDelta = add( mulSV( h * invMass, force ), mulSV( h * gravityScale, v2( P.gravityX, P.gravityY ) ) );
	float angularVelocityDelta = h * invInertia * torque;

	P.bodyK[i] = make_float4( linearVelocityDelta.x, linearVelocityDelta.y, angularVelocityDelta, linearDamping );
	P.angDamp[i] = angularDamping;
}

// b2IntegrateVelocitiesTask, src/solver.c:66-112
B2G_DEV void integrateVelocities( const StepParams& P, int i )
{
	float4 v = __ldcg( P.vel + i + 1 );
	float4 k = P.bodyK[i];
	float angularDamping = P.angDamp[i];

	v.x = k.x + k.w * v.x;
	v.y = k.y + k.w * v.y;
	v.z = k.z + angularDamping * v.z;
	__stcg( P.vel + i + 1, v );
}

// b2IntegratePositionsTask, src/solver.c:114-162
B2G_DEV void integratePositions( const StepParams& P, int i )
{
	float4 s = __ldcg( P.vel + i + 1 );
	float4 p = __ldcg( P.pos + i + 1 );
	uint32_t flags = __float_as_uint( s.w );

	float h = P.h;
	float maxLinearSpeed = P.maxLinearVelocity;
	float maxAngularSpeed = kMaxRotation * P.inv_dt;
	float maxLinearSpeedSquared = maxLinearSpeed * maxLinearSpeed;
	float maxAngularSpeedSquared = maxAngularSpeed * maxAngularSpeed;

	V2 v = v2( s.x, s.y );
	float w = s.z;

	// motion locks
	v.x = ( flags & B2L_FLAG_LOCK_LINEAR_X ) ? 0.0f : v.x;
	v.y = ( flags & B2L_FLAG_LOCK_LINEAR_Y ) ? 0.0f : v.y;
	w = ( flags & B2L_FLAG_LOCK_ANGULAR_Z ) ? 0.0f : w;

	if ( dot( v, v ) > maxLinearSpeedSquared )
	{
		float ratio = maxLinearSpeed / length( v );
		v = mulSV( ratio, v );
		flags |= B2L_FLAG_IS_SPEED_CAPPED;
	}

	if ( w * w > maxAngularSpeedSquared && ( flags & B2L_FLAG_ALLOW_FAST_ROTATION ) == 0 )
	{
		float ratio = maxAngularSpeed / absf_( w );
		w *= ratio;
		flags |= B2L_FLAG_IS_SPEED_CAPPED;
	}

	V2 dp = mulAdd( v2( p.x, p.y ), h, v );
	Rot dq;
	dq.c = p.z;
	dq.s = p.w;
	dq = integrateRotation( dq, h * w );

	__stcg( P.vel + i + 1, make_float4( v.x, v.y, w, __uint_as_float( flags ) ) );
	__stcg( P.pos + i + 1, make_float4( dp.x, dp.y, dq.c, dq.s ) );
}

// SoA -> the reference's AoS b2BodyState for the download
B2G_DEV void storeBody( const StepParams& P, int i )
{
	float4* out = reinterpret_cast<float4*>( P.outStates + (size_t)i * B2L_STATE_SIZE );
	out[0] = __ldcg( P.vel + i + 1 );
	out[1] = __ldcg( P.pos + i + 1 );
}

// ---- distribution -------------------------------------------------------------------------------------------

B2G_DEV int roundUp32( int n )
{
	return ( n + 31 ) & ~31;
}

// Calls f( itemIndex ) with ALL 32 lanes of the warp that owns each chunk (itemIndex may be >= itemCount for
// the tail lanes; the callee masks).  Needed because some stages take warp votes.
template <typename F> B2G_DEV void forEachItem( int itemCount, F f )
{
	int warpsPerBlock = (int)( blockDim.x >> 5 );
	int warp = (int)( threadIdx.x >> 5 );
	int lane = (int)( threadIdx.x & 31 );
	int stride = warpsPerBlock * (int)gridDim.x;
	for ( int chunk = warp * (int)gridDim.x + (int)blockIdx.x; chunk * 32 < itemCount; chunk += stride )
	{
		f( chunk * 32 + lane );
	}
}

// One colour stage: joints first (items [0, jointCount)), then contacts from the next multiple of 32 so that a
// warp never mixes the two kinds and a contact's colour-local index is congruent to its lane modulo 32.
// Joint blocks and contact blocks of one colour run concurrently in the reference too (src/solver.h:34-45).
template <typename FJ, typename FC> B2G_DEV void forEachInColor( const ColorRange& color, FJ joint, FC contact )
{
	int jointSpan = roundUp32( color.jointCount );
	int itemCount = jointSpan + color.contactCount;
	unsigned lane = threadIdx.x & 31u;
	forEachItem( itemCount, [&]( int t ) {
		if ( t < jointSpan )
		{
			if ( t < color.jointCount )
			{
				joint( color.jointStart + t );
			}
		}
		else
		{
			int local = t - jointSpan;
			contact( color.contactStart + local, local < color.contactCount, lane );
		}
	} );
}

B2G_DEV b2lJointSim* jointAt( const StepParams& P, int index )
{
	return reinterpret_cast<b2lJointSim*>( P.joints + (size_t)index * B2L_JOINT_SIZE );
}

B2G_DEV bool isLeadThread()
{
	return blockIdx.x == 0 && threadIdx.x == 0;
}

// ---- one stage ----------------------------------------------------------------------------------------------
B2G_DEV void runStage( const StepParams& P, int op, int colorIndex )
{
	switch ( op )
	{
		case OP_PREPARE:
		{
			if ( isLeadThread() )
			{
				__stcg( P.vel, make_float4( 0.0f, 0.0f, 0.0f, __uint_as_float( 0u ) ) );
				__stcg( P.pos, make_float4( 0.0f, 0.0f, 1.0f, 0.0f ) );
			}
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					loadBody( P, i );
				}
			} );
			// coloured contacts, flat over all colours (b2_stagePrepareContacts, src/solver.c:1068-1074)
			for ( int c = 0; c < P.colorCount; ++c )
			{
				ColorRange color = P.colors[c];
				forEachItem( color.contactCount, [&]( int i ) {
					if ( i < color.contactCount )
					{
						prepareContact( P, color.contactStart + i, true );
					}
				} );
			}
			// overflow contacts (b2PrepareContacts_Overflow, src/solver.c:1078): order free, they only read
			forEachItem( P.overflow.contactCount, [&]( int i ) {
				if ( i < P.overflow.contactCount )
				{
					prepareContact( P, P.overflow.contactStart + i, false );
				}
			} );
			// stage the host-prepared joints into the working copy + clear the event bit sets
			{
				int words = P.jointCount * ( B2L_JOINT_SIZE / 4 );
				const uint32_t* src = reinterpret_cast<const uint32_t*>( P.rawJoints );
				uint32_t* dst = reinterpret_cast<uint32_t*>( P.joints );
				forEachItem( words, [&]( int i ) {
					if ( i < words )
					{
						dst[i] = src[i];
					}
				} );
				forEachItem( P.hitWords, [&]( int i ) {
					if ( i < P.hitWords )
					{
						P.hitBits[i] = 0u;
					}
				} );
				forEachItem( P.jointWords, [&]( int i ) {
					if ( i < P.jointWords )
					{
						P.jointBits[i] = 0u;
					}
				} );
			}
		}
		break;

		case OP_INTEGRATE_VELOCITIES:
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					integrateVelocities( P, i );
				}
			} );
			break;

		case OP_INTEGRATE_POSITIONS:
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					integratePositions( P, i );
				}
			} );
			break;

		case OP_WARM:
			forEachInColor(
				P.colors[colorIndex], [&]( int j ) { warmStartJoint( P, jointAt( P, j ) ); },
				[&]( int slot, bool active, unsigned ) {
					if ( active )
					{
						warmStartContact( P, slot );
					}
				} );
			break;

		case OP_SOLVE:
			forEachInColor(
				P.colors[colorIndex],
				[&]( int j ) {
					b2lJointSim* joint = jointAt( P, j );
					solveJoint( P, joint, true );
					jointEventTest( P, joint );
				},
				[&]( int slot, bool active, unsigned lane ) { solveContact( P, slot, active, true, lane ); } );
			break;

		case OP_RELAX:
			forEachInColor(
				P.colors[colorIndex], [&]( int j ) { solveJoint( P, jointAt( P, j ), false ); },
				[&]( int slot, bool active, unsigned lane ) { solveContact( P, slot, active, false, lane ); } );
			break;

		case OP_RESTITUTION:
			forEachInColor(
				P.colors[colorIndex], [&]( int ) {},
				[&]( int slot, bool active, unsigned lane ) { restitutionContact( P, slot, active, lane ); } );
			break;

		// The overflow colour is solved by ONE thread, strictly in array order, joints before contacts
		// (src/solver.c:1100-1101, 1119-1120, 1147-1148, 1168).
		case OP_OVERFLOW_WARM:
			if ( isLeadThread() )
			{
				for ( int i = 0; i < P.overflow.jointCount; ++i )
				{
					warmStartJoint( P, jointAt( P, P.overflow.jointStart + i ) );
				}
				for ( int i = 0; i < P.overflow.contactCount; ++i )
				{
					warmStartContactOverflow( P, P.overflow.contactStart + i );
				}
			}
			break;

		case OP_OVERFLOW_SOLVE:
		case OP_OVERFLOW_RELAX:
			if ( isLeadThread() )
			{
				bool useBias = op == OP_OVERFLOW_SOLVE;
				for ( int i = 0; i < P.overflow.jointCount; ++i )
				{
					solveJoint( P, jointAt( P, P.overflow.jointStart + i ), useBias );
				}
				for ( int i = 0; i < P.overflow.contactCount; ++i )
				{
					solveContactOverflow( P, P.overflow.contactStart + i, useBias );
				}
			}
			break;

		case OP_OVERFLOW_RESTITUTION:
			if ( isLeadThread() )
			{
				for ( int i = 0; i < P.overflow.contactCount; ++i )
				{
					restitutionContactOverflow( P, P.overflow.contactStart + i );
				}
			}
			break;

		case OP_STORE:
		{
			for ( int c = 0; c < P.colorCount; ++c )
			{
				ColorRange color = P.colors[c];
				forEachItem( color.contactCount, [&]( int i ) {
					if ( i < color.contactCount )
					{
						storeContact( P, color.contactStart + i, true );
					}
				} );
			}
			forEachItem( P.overflow.contactCount, [&]( int i ) {
				if ( i < P.overflow.contactCount )
				{
					storeContact( P, P.overflow.contactStart + i, false );
				}
			} );
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					storeBody( P, i );
				}
			} );
		}
		break;

		default:
			break;
	}
}

} // namespace b2g
