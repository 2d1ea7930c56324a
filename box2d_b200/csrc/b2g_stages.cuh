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
constexpr int kBlockThreads = 256;

B2G_DEV bool isLeadThread()
{
	return blockIdx.x == 0 && threadIdx.x == 0;
}

// ---- grid barrier ---------------------------------------------------------------------------------------------
// Arrive = release-add at gpu scope by one thread after the block has synchronised; wait = acquire-load spin.
// The acquire makes the other blocks' body/constraint writes visible to every thread of this block after the
// trailing __syncthreads (PTX memory model: bar.sync and release/acquire chains compose by causality order).
B2G_DEV void gridBarrier( unsigned int* counter, unsigned int target )
{
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		asm volatile( "red.release.gpu.global.add.u32 [%0], 1;" ::"l"( counter ) : "memory" );
		unsigned int seen;
		do
		{
			asm volatile( "ld.acquire.gpu.global.u32 %0, [%1];" : "=r"( seen ) : "l"( counter ) : "memory" );
		}
		while ( seen < target );
	}
	__syncthreads();
}

struct StageClock
{
	long long last;
	long long acc[b2GpuStage_count];
	bool lead;

	B2G_DEV void start()
	{
		lead = isLeadThread();
		last = lead ? clock64() : 0;
#pragma unroll
		for ( int i = 0; i < b2GpuStage_count; ++i )
		{
			acc[i] = 0;
		}
	}

	// timers are compile-time constants, so acc[] stays in registers
	B2G_DEV void lap( int timer )
	{
		if ( lead )
		{
			long long now = clock64();
			acc[timer] += now - last;
			last = now;
		}
	}
};


// ---- body stages --------------------------------------------------------------------------------------------

// Per-step body constants.  The velocity increment of b2IntegrateVelocitiesTask (src/solver.c:94-102) depends only
// on per-step constants, so it is evaluated once here with the same operations and reused by every sub-step.
// `body` is the global body index (wire order), `local` the 1-based index in the view.
B2G_DEV void loadBody( const StepParams& P, const SolveView& V, int body, int local )
{
	const float4* s4 = reinterpret_cast<const float4*>( P.rawStates + (size_t)body * B2L_STATE_SIZE );
	V.vel[local] = s4[0];
	V.pos[local] = s4[1];

	float4 a = P.wireBody[2 * (size_t)body + 0]; // invMass, invInertia, force.x, force.y
	float4 b = P.wireBody[2 * (size_t)body + 1]; // torque, linearDamping, angularDamping, gravityScale
	float invMass = a.x;
	float invInertia = a.y;
	V2 force = v2( a.z, a.w );
	float torque = b.x;
	float h = P.h;

	float linearDamping = 1.0f / ( 1.0f + h * b.y );
	float angularDamping = 1.0f / ( 1.0f + h * b.z );

	// gravity scale will be zero for kinematic bodies
	float gravityScale = invMass > 0.0f ? b.w : 0.0f;

	V2 linearVelocityDelta = add( mulSV( h * invMass, force ), mulSV( h * gravityScale, v2( P.gravityX, P.gravityY ) ) );
	float angularVelocityDelta = h * invInertia * torque;

	V.bodyK[local - 1] = make_float4( linearVelocityDelta.x, linearVelocityDelta.y, angularVelocityDelta, linearDamping );
	V.angDamp[local - 1] = angularDamping;
}

// b2IntegrateVelocitiesTask, src/solver.c:66-112 (i is 0-based in the view)
B2G_DEV void integrateVelocities( const SolveView& V, int i )
{
	float4 v = V.vel[i + 1];
	float4 k = V.bodyK[i];
	float angularDamping = V.angDamp[i];

	v.x = k.x + k.w * v.x;
	v.y = k.y + k.w * v.y;
	v.z = k.z + angularDamping * v.z;
	V.vel[i + 1] = v;
}

// b2IntegratePositionsTask, src/solver.c:114-162
B2G_DEV void integratePositions( const StepParams& P, const SolveView& V, int i )
{
	float4 s = V.vel[i + 1];
	float4 p = V.pos[i + 1];
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

	V.vel[i + 1] = make_float4( v.x, v.y, w, __uint_as_float( flags ) );
	V.pos[i + 1] = make_float4( dp.x, dp.y, dq.c, dq.s );
}

// one 32-byte record in ONE store (sm_100: STG.256; the records are 32-byte aligned).  With direct outputs
// the store is a PCIe write: a whole sector per instruction instead of two half-written ones.
B2G_DEV void storeState( void* record, float4 a, float4 b )
{
	asm volatile( "st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"( record ), "f"( a.x ), "f"( a.y ), "f"( a.z ), "f"( a.w ),
				  "f"( b.x ), "f"( b.y ), "f"( b.z ), "f"( b.w )
				  : "memory" );
}

// view -> the reference's AoS b2BodyState for the download
B2G_DEV void storeBody( const StepParams& P, const SolveView& V, int body, int local )
{
	float4 v = V.vel[local];
	storeState( P.outStates + (size_t)body * B2L_STATE_SIZE, v, V.pos[local] );
	if ( P.residentOut != nullptr )
	{
		// what the host's state will be when the next step begins, unless somebody touches the body in between (the pack
		// pass checks): b2FinalizeBodiesTask resets the deltas and clears the transient flags (src/solver.c:611-612, :632)
		v.w = __uint_as_float( __float_as_uint( v.w ) & ~B2L_FLAG_TRANSIENT );
		storeState( P.residentOut + (size_t)body * B2L_STATE_SIZE, v, make_float4( 0.0f, 0.0f, 1.0f, 0.0f ) );
	}
}

// {v, w} of a body as the prepare stage needs it, straight from the wire states (0 for the static dummy)
B2G_DEV float4 wireVelocity( const StepParams& P, int index )
{
	if ( index < 0 )
	{
		return make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
	}
	return *reinterpret_cast<const float4*>( P.rawStates + (size_t)index * B2L_STATE_SIZE );
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

B2G_DEV b2lJointSim* jointAt( const SolveView& V, int index )
{
	return reinterpret_cast<b2lJointSim*>( V.joints + (size_t)index * kJointStride );
}

// ---- a joint of a view by slot, whatever the view keeps: the full 256-byte record or a LiteRevolute ------------------------
B2G_DEV size_t jointBytesOf( bool lite )
{
	return lite ? (size_t)kLiteJointWords * sizeof( float ) : (size_t)kJointStride;
}

// the two bodies of the joint in slot k (-1 = static); false for a filter joint (no solver data)
B2G_DEV bool jointSlotBodies( const SolveView& V, int k, int& a, int& b )
{
	if ( V.jointLite != 0 )
	{
		const float* lite = liteJointAt( V, k );
		a = __float_as_int( lite[LR_INDEX_A] );
		b = __float_as_int( lite[LR_INDEX_B] );
		return true;
	}
	const int* pair = jointIndexPair( jointAt( V, k ) );
	a = pair != nullptr ? pair[0] : -1;
	b = pair != nullptr ? pair[1] : -1;
	return pair != nullptr;
}

B2G_DEV void setJointSlotBodies( const SolveView& V, int k, int a, int b )
{
	if ( V.jointLite != 0 )
	{
		float* lite = liteJointAt( V, k );
		lite[LR_INDEX_A] = __int_as_float( a );
		lite[LR_INDEX_B] = __int_as_float( b );
		return;
	}
	int* pair = jointIndexPair( jointAt( V, k ) );
	if ( pair != nullptr )
	{
		pair[0] = a;
		pair[1] = b;
	}
}

// false for the one joint that writes no body in a pass: a pogo joint without a spring (src/pogo_joint.c:165,203)
B2G_DEV bool jointSlotWrites( const SolveView& V, int k )
{
	if ( V.jointLite != 0 )
	{
		return true;
	}
	const b2lJointSim* joint = jointAt( V, k );
	return !( joint->type == b2l_pogoJoint && joint->u.pogo.hertz == 0.0f );
}

B2G_DEV void warmStartJointSlot( const StepParams& P, const SolveView& V, const SolveView& geometry, int k )
{
	// (`geometry` names the records, `V` the body arrays the stage writes through: the cluster kernel's counted stores)
	if ( geometry.jointLite != 0 )
	{
		warmStartRevoluteLite( V, liteJointAt( geometry, k ) );
	}
	else
	{
		warmStartJoint( P, V, jointAt( geometry, k ) );
	}
}

B2G_DEV void solveJointSlot( const StepParams& P, const SolveView& V, const SolveView& geometry, int k, bool useBias, bool events )
{
	if ( geometry.jointLite != 0 )
	{
		float* lite = liteJointAt( geometry, k );
		solveRevoluteLite( V, lite, useBias );
		if ( events )
		{
			jointEventTestLite( P, lite );
		}
	}
	else
	{
		b2lJointSim* joint = jointAt( geometry, k );
		solveJoint( P, V, joint, useBias );
		if ( events )
		{
			jointEventTest( P, joint );
		}
	}
}

B2G_DEV void storeJointSlot( const StepParams& P, const SolveView& V, int jointIndex, int k )
{
	if ( V.jointLite != 0 )
	{
		storeJointImpulsesLite( P, jointIndex, liteJointAt( V, k ) );
	}
	else
	{
		storeJointImpulses( P, jointIndex, jointAt( V, k ) );
	}
}

// ---- joints of the grid-barrier kernel, cached by the block that solves them --------------------------------------------
// forEachInColor gives joint t of a colour to chunk t / 32, i.e. to block (t / 32) % gridDim -- the same thread in every
// stage of the step.  The persistent kernel therefore keeps the 256-byte records of ITS joints in shared memory for the
// whole step instead of paying the L2 round trips of their fields (several dependent ones: flags -> branch -> more fields)
// in each of the ~80 stages; the bodies stay in global memory, they are what the blocks share.  Joints beyond the
// capacity and the overflow colour's joints stay in the global working copy.
struct JointCache
{
	uint8_t* base; // capacity * kJointStride bytes of shared memory
	int capacity;
	int colorBase[kMaxColors + 1]; // first slot of every colour's joints of this block
};

// slots this block needs for a colour with `jointCount` joints (whole chunks of 32)
B2G_DEV int jointCacheSlots( int jointCount )
{
	int chunks = ( jointCount + 31 ) >> 5;
	int mine = (int)blockIdx.x < chunks ? ( chunks - (int)blockIdx.x + (int)gridDim.x - 1 ) / (int)gridDim.x : 0;
	return mine * 32;
}

// joint t of colour c (this thread owns it): its cached record, or the global working copy
B2G_DEV b2lJointSim* jointOfColor( const SolveView& V, const JointCache* cache, int c, const ColorRange& color, int t )
{
	if ( cache != nullptr )
	{
		int local = cache->colorBase[c] + ( ( t >> 5 ) / (int)gridDim.x ) * 32 + ( t & 31 );
		if ( local < cache->capacity )
		{
			return reinterpret_cast<b2lJointSim*>( cache->base + (size_t)local * kJointStride );
		}
	}
	return jointAt( V, color.jointStart + t );
}

// ---- one stage of the grid-barrier path (view = global memory, wire slot == constraint slot) -----------------------
B2G_DEV void prepareContactGlobal( const StepParams& P, int slot, bool inRange, bool wide, unsigned lane )
{
	float4 head = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
	if ( inRange )
	{
		head = wireHead( P, slot );
	}
	// dead slots (padding between the segments of a batch) have pointCount 0
	bool active = inRange && ( __float_as_int( head.z ) & kMetaPointMask ) != 0;
	int groupBits = wide ? __float_as_int( head.z ) & kMetaGroupMask : 0;
	(void)lane;
	if ( active )
	{
		int indexA = __float_as_int( head.x );
		int indexB = __float_as_int( head.y );
		prepareContact( P, P.g, slot, slot, indexA + 1, indexB + 1, wireVelocity( P, indexA ), wireVelocity( P, indexB ), wide,
						groupBits );
	}
	else if ( inRange )
	{
		// an all-zero constraint on the static dummy is a no-op in every stage, like the zeroed tail lanes of the
		// reference's last wide constraint (src/solver.c:1419-1425)
		const float4 zero = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
#pragma unroll
		for ( int f = 0; f < CF_COUNT; ++f )
		{
			storeField( P.g, f, slot, zero );
		}
		P.g.cidx[slot] = make_int2( 0, 0 );
		P.g.cmeta[slot] = 0;
	}
}

B2G_DEV void runStage( const StepParams& P, int op, int colorIndex, const JointCache* cache = nullptr )
{
	const SolveView& V = P.g;
	switch ( op )
	{
		case OP_PREPARE:
		{
			if ( isLeadThread() )
			{
				V.vel[0] = make_float4( 0.0f, 0.0f, 0.0f, __uint_as_float( 0u ) );
				V.pos[0] = make_float4( 0.0f, 0.0f, 1.0f, 0.0f );
			}
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					loadBody( P, V, i, i + 1 );
				}
			} );
			// coloured contacts, flat over all colours (b2_stagePrepareContacts, src/solver.c:1068-1074)
			unsigned lane = threadIdx.x & 31u;
			for ( int c = 0; c < P.colorCount; ++c )
			{
				ColorRange color = P.colors[c];
				forEachItem( color.contactCount, [&]( int i ) {
					prepareContactGlobal( P, color.contactStart + i, i < color.contactCount, true, lane );
				} );
			}
			// overflow contacts (b2PrepareContacts_Overflow, src/solver.c:1078): order free, they only read
			forEachItem( P.overflow.contactCount, [&]( int i ) {
				prepareContactGlobal( P, P.overflow.contactStart + i, i < P.overflow.contactCount, false, lane );
			} );
			// stage the host-prepared joints into the working copy + clear the joint event bit set
			{
				// (when every coloured joint is cached by its block, only the overflow colour's joints need the working copy)
				const bool allCached = cache != nullptr && cache->capacity >= cache->colorBase[P.colorCount] && P.gridJointsAllCached != 0;
				int firstWord = allCached ? P.overflow.jointStart * ( kJointStride / 4 ) : 0;
				int words = P.jointCount * ( kJointStride / 4 ) - firstWord;
				const uint32_t* src = reinterpret_cast<const uint32_t*>( P.rawJoints ) + firstWord;
				uint32_t* dst = reinterpret_cast<uint32_t*>( V.joints ) + firstWord;
				forEachItem( words, [&]( int i ) {
					if ( i < words )
					{
						dst[i] = src[i];
					}
				} );
				forEachItem( P.jointWords, [&]( int i ) {
					if ( i < P.jointWords )
					{
						P.jointBits[i] = 0u;
					}
				} );
				if ( cache != nullptr )
				{
					// a chunk of 32 joints is 8 KB in a row, in the wire arena and in the cache: the warp copies it as such
					const int lane = (int)( threadIdx.x & 31u );
					for ( int c = 0; c < P.colorCount; ++c )
					{
						ColorRange color = P.colors[c];
						forEachItem( color.jointCount, [&]( int t ) {
							int first = t - lane; // the chunk's first joint
							int count = color.jointCount - first < 32 ? color.jointCount - first : 32;
							int local = cache->colorBase[c] + ( ( first >> 5 ) / (int)gridDim.x ) * 32;
							count = local + count <= cache->capacity ? count : cache->capacity - local; // the rest is not cached
							const int quads = kJointStride / 16;
							const float4* from = reinterpret_cast<const float4*>( P.rawJoints + (size_t)( color.jointStart + first ) * kJointStride );
							float4* to = reinterpret_cast<float4*>( cache->base + (size_t)local * kJointStride );
							for ( int q = lane; q < count * quads; q += 32 )
							{
								// asynchronous copies: the chunks of all colours are in flight together
								unsigned target = (unsigned)__cvta_generic_to_shared( to + q );
								asm volatile( "cp.async.cg.shared.global [%0], [%1], 16;" ::"r"( target ), "l"( from + q ) : "memory" );
							}
						} );
					}
					asm volatile( "cp.async.wait_all;" ::: "memory" );
				}
			}
		}
		break;

		case OP_INTEGRATE_VELOCITIES:
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					integrateVelocities( V, i );
				}
			} );
			break;

		case OP_INTEGRATE_POSITIONS:
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					integratePositions( P, V, i );
				}
			} );
			break;

		case OP_WARM:
			forEachInColor(
				P.colors[colorIndex],
				[&]( int j ) { warmStartJoint( P, V, jointOfColor( V, cache, colorIndex, P.colors[colorIndex], j - P.colors[colorIndex].jointStart ) ); },
				[&]( int slot, bool active, unsigned ) {
					if ( active )
					{
						warmStartContact( V, slot );
					}
				} );
			break;

		case OP_SOLVE:
			forEachInColor(
				P.colors[colorIndex],
				[&]( int j ) {
					b2lJointSim* joint = jointOfColor( V, cache, colorIndex, P.colors[colorIndex], j - P.colors[colorIndex].jointStart );
					solveJoint( P, V, joint, true );
					jointEventTest( P, joint );
				},
				[&]( int slot, bool active, unsigned ) {
					if ( active )
					{
						solveContact( P, V, slot, true );
					}
				} );
			break;

		case OP_RELAX:
			forEachInColor(
				P.colors[colorIndex],
				[&]( int j ) { solveJoint( P, V, jointOfColor( V, cache, colorIndex, P.colors[colorIndex], j - P.colors[colorIndex].jointStart ), false ); },
				[&]( int slot, bool active, unsigned ) {
					if ( active )
					{
						solveContact( P, V, slot, false );
					}
				} );
			break;

		case OP_RESTITUTION:
			forEachInColor(
				P.colors[colorIndex], [&]( int ) {},
				[&]( int slot, bool active, unsigned ) {
					if ( active )
					{
						restitutionContact( P, V, slot );
					}
				} );
			break;

		// The overflow colour is solved by ONE thread, strictly in array order, joints before contacts
		// (src/solver.c:1100-1101, 1119-1120, 1147-1148, 1168).
		case OP_OVERFLOW_WARM:
			if ( isLeadThread() )
			{
				for ( int i = 0; i < P.overflow.jointCount; ++i )
				{
					warmStartJoint( P, V, jointAt( V, P.overflow.jointStart + i ) );
				}
				for ( int i = 0; i < P.overflow.contactCount; ++i )
				{
					warmStartContactOverflow( V, P.overflow.contactStart + i );
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
					solveJoint( P, V, jointAt( V, P.overflow.jointStart + i ), useBias );
				}
				for ( int i = 0; i < P.overflow.contactCount; ++i )
				{
					solveContactOverflow( P, V, P.overflow.contactStart + i, useBias );
				}
			}
			break;

		case OP_OVERFLOW_RESTITUTION:
			if ( isLeadThread() )
			{
				for ( int i = 0; i < P.overflow.contactCount; ++i )
				{
					restitutionContactOverflow( P, V, P.overflow.contactStart + i );
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
						storeContact( P, V, color.contactStart + i, color.contactStart + i, true );
					}
				} );
			}
			forEachItem( P.overflow.contactCount, [&]( int i ) {
				if ( i < P.overflow.contactCount )
				{
					storeContact( P, V, P.overflow.contactStart + i, P.overflow.contactStart + i, false );
				}
			} );
			forEachItem( P.bodyCount, [&]( int i ) {
				if ( i < P.bodyCount )
				{
					storeBody( P, V, i, i + 1 );
				}
			} );
			// joints colour by colour, by the thread that solved them (its record may live in the block's shared memory)
			for ( int c = 0; c < P.colorCount; ++c )
			{
				ColorRange color = P.colors[c];
				forEachItem( color.jointCount, [&]( int t ) {
					if ( t < color.jointCount )
					{
						storeJointImpulses( P, color.jointStart + t, jointOfColor( V, cache, c, color, t ) );
					}
				} );
			}
			forEachItem( P.overflow.jointCount, [&]( int t ) {
				if ( t < P.overflow.jointCount )
				{
					storeJointImpulses( P, P.overflow.jointStart + t, jointAt( V, P.overflow.jointStart + t ) );
				}
			} );
		}
		break;

		default:
			break;
	}
}

} // namespace b2g
