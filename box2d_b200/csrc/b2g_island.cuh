// b2g_island.cuh -- island-local execution: zero grid barriers.
//
// Constraints only couple bodies of the same simulation island (reference src/island.c: b2LinkContact / b2LinkJoint
// merge the islands of the two bodies, static bodies belong to none), so islands can be solved independently and the
// colour order only has to be respected INSIDE an island.  The host packs the awake islands into `binCount` bins
// (b2GpuStepDesc::bodyIsland -> bodyBin); here
//   * b2gScatterKernel appends every body and constraint to its bin's list in one flat pass (one block per bin; the
//     island kernel sorts its list by colour in shared memory), or -- bins shared by a cluster -- b2gPartitionKernel
//     buckets them by (bin, colour) with atomics and one grid barrier into colour-major lists, and
//   * b2gIslandKernel gives each bin to ONE thread block that keeps the bin's body state, contact constraints and
//     joints in shared memory for the whole step and separates colours with __syncthreads() (~20 ns) instead of a
//     grid-wide barrier (~1.2 us measured, tools/microbench/barrier_bench.cu).
// The order of constraints inside a (bin, colour) bucket comes from atomics and is not deterministic, but constraints
// of one colour touch disjoint dynamic bodies (src/constraint_graph.c:84-133), so every body sees exactly the same
// sequence of updates as in the reference: results are bit-identical (tests/test_gpu_lockstep.py).  The SIMD-group
// early-outs of the wide path depend on the reference's array order; they are evaluated in wire order by the
// scatter / partition kernel and travel with the constraint (kMetaGroup* bits).
// When a bin does not fit its shared-memory budget the scatter / partition kernel raises binFail and the grid-barrier
// kernel (b2g_grid.cuh) solves the step instead.  The overflow colour of a bin (strictly sequential in the reference) is
// levelised, or -- a deep chain -- walked by one warp whose lanes take turns (overflowChainWarp, b2g_contact.cuh).
#pragma once

#include "b2g_stages.cuh"

namespace b2g
{

constexpr int kIslandThreads = 512;
constexpr int kPartitionThreads = 1024; // one block per SM: few arrivals at the grid barrier, enough threads for one item each
constexpr int kColorSlots = kMaxColors + 2; // per-bin offsets: active colours, then the overflow bucket, then the total
constexpr int kMaxBinOverflow = 256;		   // overflow constraints (contacts or joints) one bin can order

B2G_DEV int jointBodyForBin( const uint8_t* record )
{
	// the per-type block starts with different fields, indexA/indexB sit at a type dependent offset
	b2lJointSim* joint = reinterpret_cast<b2lJointSim*>( const_cast<uint8_t*>( record ) );
	const int* pair = jointIndexPair( joint );
	if ( pair == nullptr )
	{
		return -1;
	}
	return pair[0] >= 0 ? pair[0] : pair[1];
}

// Warp-aggregated counter increment: lanes with the same key elect a leader that does ONE atomicAdd for the group and
// hands out consecutive values.  Must be called by all 32 lanes; inactive lanes pass active == false.
B2G_DEV int aggregatedAdd( int* counters, int key, bool active )
{
	unsigned peers = __match_any_sync( 0xffffffffu, active ? key : -1 - (int)( threadIdx.x & 31u ) );
	int result = 0;
	if ( active )
	{
		unsigned lane = threadIdx.x & 31u;
		int leader = __ffs( peers ) - 1;
		int base = 0;
		if ( (int)lane == leader )
		{
			base = atomicAdd( counters + key, __popc( peers ) );
		}
		base = __shfl_sync( peers, base, leader );
		result = base + __popc( peers & ( ( 1u << lane ) - 1u ) );
	}
	return result;
}

// ---- partition ---------------------------------------------------------------------------------------------------------
// counters (binBodyCount, binColorStart, binJointStart, binFail) are zeroed by the host before the launch
__global__ void __launch_bounds__( kPartitionThreads, 1 ) b2gPartitionKernel( const __grid_constant__ StepParams P )
{
	// the cluster kernel is launched as a programmatic dependent: its blocks may be set up while this grid is running
	asm volatile( "griddepcontrol.launch_dependents;" );
	const unsigned blocks = gridDim.x;

	// phase 1: local index of every body in its bin, (bin, rank) of every constraint, SIMD-group bits in wire order.
	// Every pass is flat over all items (no per-colour loop: each pass is a chain of dependent L2 round trips, so the
	// passes must not be serialised), and lanes that hit the same counter are aggregated into one atomic.
	forEachItem( P.jointWords, [&]( int i ) {
		if ( i < P.jointWords )
		{
			P.jointBits[i] = 0u;
		}
	} );
	// Owner lists (one bin shared by a cluster): the bodies keep their order -- neighbours in the awake set are neighbours
	// in the scene more often than not -- and every constraint goes to the list of the block that owns its first body,
	// so that most gathers and scatters of the cluster kernel stay in the block's own shared memory.
	const bool ownerLists = P.ownerLists != 0;
	if ( ownerLists )
	{
		forEachItem( P.bodyCount, [&]( int b ) {
			if ( b < P.bodyCount )
			{
				P.binBodyList[b] = b;
				P.bodyLocal[b] = b + 1;
			}
		} );
		if ( blockIdx.x == 0 && threadIdx.x == 0 )
		{
			P.binBodyCount[0] = P.bodyCount;
		}
	}
	forEachItem( ownerLists ? 0 : P.bodyCount, [&]( int b ) {
		bool active = b < P.bodyCount;
		int bin = active ? P.bodyBin[b] : -1;
		int local = aggregatedAdd( P.binBodyCount, bin, active );
		if ( active )
		{
			if ( local < P.binCapBodies )
			{
				P.binBodyList[(size_t)bin * P.binCapBodies + local] = b;
			}
			else
			{
				*P.binFail = 1;
			}
			P.bodyLocal[b] = local + 1;
		}
	} );
	// contacts: colour slots are 32-aligned, so a chunk of 32 slots belongs to exactly one colour (or to the overflow
	// colour, bucket index colorCount, whose constraints keep their array order: see the island kernel)
	forEachItem( P.contactSlots, [&]( int slot ) {
		int c = 0;
		while ( c < P.colorCount && ( c + 1 < P.colorCount ? P.colors[c + 1].contactStart : P.overflow.contactStart ) <= slot )
		{
			c += 1;
		}
		bool isOverflow = c == P.colorCount;
		ColorRange color = isOverflow ? P.overflow : P.colors[c];
		bool inRange = slot < color.contactStart + color.contactCount;
		float4 head = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
		if ( inRange )
		{
			head = wireHead( P, slot );
		}
		// dead slots (padding between the segments of a batch) have pointCount 0 and are not placed in any bin
		bool active = inRange && ( __float_as_int( head.z ) & kMetaPointMask ) != 0;
		int bits = isOverflow ? 0 : __float_as_int( head.z ) & kMetaGroupMask;
		int key = -1;
		int bin = 0;
		if ( active )
		{
			int indexA = __float_as_int( head.x );
			int indexB = __float_as_int( head.y );
			int first = indexA >= 0 ? indexA : indexB;
			// the list of the constraint: its bin, or -- owner lists -- the block that owns its first body (the overflow
			// colour is solved by the cluster's first block)
			bin = ownerLists ? ( isOverflow ? 0 : (int)__umulhi( (unsigned)first, P.clusterMagic ) ) : P.bodyBin[first];
			if ( indexA >= 0 && indexB >= 0 && P.bodyBin[indexA] != P.bodyBin[indexB] )
			{
				*P.binFail = 1; // wrong island hint: the two bodies of a constraint must share an island (and so a bin)
			}
			key = bin * kColorSlots + c;
		}
		int rank = aggregatedAdd( P.binColorStart, key, active );
		if ( active )
		{
			P.contactBinRank[slot] = make_int2( bin, rank );
			P.slotGroupBits[slot] = bits;
		}
	} );
	forEachItem( P.jointCount, [&]( int j ) {
		bool active = j < P.jointCount;
		int key = -1, bin = 0;
		if ( active )
		{
			int c = 0;
			while ( c < P.colorCount && ( c + 1 < P.colorCount ? P.colors[c + 1].jointStart : P.overflow.jointStart ) <= j )
			{
				c += 1;
			}
			int body = jointBodyForBin( P.rawJoints + (size_t)j * kJointStride );
			{
				b2lJointSim* record = reinterpret_cast<b2lJointSim*>( const_cast<uint8_t*>( P.rawJoints + (size_t)j * kJointStride ) );
				if ( P.liteJoints != 0 && !isLiteRevolute( record ) )
				{
					// the plan counted on plain revolute joints only (b2gPlanIslands): the grid-barrier kernel takes the step
					*P.binFail = 1;
				}
				const int* pair = jointIndexPair( record );
				if ( pair != nullptr && pair[0] >= 0 && pair[1] >= 0 && P.bodyBin[pair[0]] != P.bodyBin[pair[1]] )
				{
					*P.binFail = 1;
				}
			}
			// a filter joint has no solver data: park it in bin 0, it is a no-op in every stage
			bin = body < 0 ? 0 : ownerLists ? ( c == P.colorCount ? 0 : (int)__umulhi( (unsigned)body, P.clusterMagic ) ) : P.bodyBin[body];
			key = bin * kColorSlots + c;
		}
		int rank = aggregatedAdd( P.binJointStart, key, active );
		if ( active )
		{
			P.jointBinRank[j] = make_int2( bin, rank );
		}
	} );
	gridBarrier( P.barrier + 1, blocks );

	// phase 2: the counts are final.  One thread per bin turns them into the exclusive offsets the island kernels read
	// (colour-major layout of the bin's lists) and checks the capacities; the placement passes below do not wait for
	// that, every constraint sums the counts of the colours before its own (a handful of L2 hits).
	forEachItem( P.listCount, [&]( int bin ) {
		if ( bin < P.listCount )
		{
			const int* contactCount = P.binColorStart + (size_t)bin * kColorSlots;
			const int* jointCount = P.binJointStart + (size_t)bin * kColorSlots;
			int* contactOffset = P.binColorOffset + (size_t)bin * kColorSlots;
			int* jointOffset = P.binJointOffset + (size_t)bin * kColorSlots;
			int contacts = 0, joints = 0;
			// what the fullest block of the bin's cluster has to hold: a colour is dealt out evenly, the overflow
			// colour goes to the first block as a whole
			int blockContacts = 0, blockJoints = 0;
			const int share = P.clusterSize;
			for ( int c = 0; c <= P.colorCount; ++c ) // the bucket after the last colour is the overflow colour
			{
				int n = contactCount[c];
				contactOffset[c] = contacts;
				contacts += n;
				int m = jointCount[c];
				jointOffset[c] = joints;
				joints += m;
				blockContacts += c < P.colorCount && !ownerLists ? ( n + share - 1 ) / share : n;
				blockJoints += c < P.colorCount && !ownerLists ? ( m + share - 1 ) / share : m;
				if ( ownerLists && ( n | m ) != 0 )
				{
					// what the whole bin has of this colour (every block of the cluster must agree on skipping a colour)
					atomicAdd( P.binColorTotal + c, n );
					atomicAdd( P.binColorTotal + kColorSlots + c, m );
				}
				if ( c == P.colorCount && ( n > kMaxBinOverflow || m > kMaxBinOverflow ) )
				{
					*P.binFail = 1;
				}
			}
			for ( int c = P.colorCount + 1; c < kColorSlots; ++c )
			{
				contactOffset[c] = contacts;
				jointOffset[c] = joints;
			}
			if ( blockContacts > P.capContacts || blockJoints > P.capJoints )
			{
				*P.binFail = 1;
			}
		}
	} );

	auto offsetOf = [&]( const int* counts, int bin, int c ) -> int {
		const int* row = counts + (size_t)bin * kColorSlots;
		int offset = 0;
		for ( int k = 0; k < c; ++k )
		{
			offset += row[k];
		}
		return offset;
	};
	forEachItem( P.contactSlots, [&]( int slot ) {
		int c = 0;
		while ( c < P.colorCount && ( c + 1 < P.colorCount ? P.colors[c + 1].contactStart : P.overflow.contactStart ) <= slot )
		{
			c += 1;
		}
		ColorRange color = c == P.colorCount ? P.overflow : P.colors[c];
		bool inRange = slot < color.contactStart + color.contactCount;
		float4 head = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
		if ( inRange )
		{
			head = wireHead( P, slot );
		}
		if ( inRange && ( __float_as_int( head.z ) & kMetaPointMask ) != 0 )
		{
			int2 br = P.contactBinRank[slot];
			int dest = offsetOf( P.binColorStart, br.x, c ) + br.y;
			if ( dest < P.listCapContacts ) // a list that does not fit raised binFail above
			{
				// everything the island kernel needs to start preparing the contact without chasing pointers: the
				// wire slot, the bodies' indices inside the bin, the SIMD-group bits
				P.binContactList[(size_t)br.x * P.listCapContacts + dest] = slot;
				if ( P.resolveContacts != 0 )
				{
					int indexA = __float_as_int( head.x ), indexB = __float_as_int( head.y );
					int localA = indexA >= 0 ? P.bodyLocal[indexA] : 0;
					int localB = indexB >= 0 ? P.bodyLocal[indexB] : 0;
					P.binContactInfo[(size_t)br.x * P.listCapContacts + dest] = make_int4( slot, localA, localB, P.slotGroupBits[slot] );
				}
			}
		}
	} );
	forEachItem( P.jointCount, [&]( int j ) {
		if ( j < P.jointCount )
		{
			int c = 0;
			while ( c < P.colorCount && ( c + 1 < P.colorCount ? P.colors[c + 1].jointStart : P.overflow.jointStart ) <= j )
			{
				c += 1;
			}
			int2 br = P.jointBinRank[j];
			int dest = offsetOf( P.binJointStart, br.x, c ) + br.y;
			if ( dest < P.listCapJoints )
			{
				P.binJointList[(size_t)br.x * P.listCapJoints + dest] = j;
			}
		}
	} );
}

// ---- partition, one block per bin: single pass ------------------------------------------------------------------------------
// With one thread block per bin the colour-major order of a bin's lists need not be built here: every body and every
// constraint is appended to its bin's list in ONE flat pass (a warp-aggregated atomic on the bin's counter), tagged with
// its colour, and the island kernel sorts its own few hundred entries by colour in shared memory.  No grid barrier, no
// second pass over the constraints, no cooperative launch.  The counters of bin b: binBodyCount[b], contacts
// binColorStart[b * kColorSlots + {0: all, 1: overflow colour}], joints binJointStart[...] likewise.
constexpr int kFlatColorShift = 8;	// binContactInfo.w = SIMD-group bits | colour slot << 8
constexpr int kFlatJointShift = 26; // binJointList entry = joint index | colour slot << 26

__global__ void __launch_bounds__( 256 ) b2gScatterKernel( const __grid_constant__ StepParams P )
{
	// programmatic dependent launch: the island kernel's blocks may be set up on the SMs while this grid is still running;
	// they wait (griddepcontrol.wait) for this grid to complete before they read anything
	asm volatile( "griddepcontrol.launch_dependents;" );
	forEachItem( P.jointWords, [&]( int i ) {
		if ( i < P.jointWords )
		{
			P.jointBits[i] = 0u;
		}
	} );
	forEachItem( P.bodyCount, [&]( int b ) {
		bool active = b < P.bodyCount;
		int bin = active ? P.bodyBin[b] : -1;
		int local = aggregatedAdd( P.binBodyCount, bin, active );
		if ( active )
		{
			if ( local < P.binCapBodies )
			{
				P.binBodyList[(size_t)bin * P.binCapBodies + local] = b;
			}
			else
			{
				*P.binFail = 1;
			}
			P.bodyLocal[b] = local + 1;
		}
	} );
	forEachItem( P.contactSlots, [&]( int slot ) {
		int c = 0;
		while ( c < P.colorCount && ( c + 1 < P.colorCount ? P.colors[c + 1].contactStart : P.overflow.contactStart ) <= slot )
		{
			c += 1;
		}
		bool isOverflow = c == P.colorCount;
		ColorRange color = isOverflow ? P.overflow : P.colors[c];
		bool inRange = slot < color.contactStart + color.contactCount;
		float4 head = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
		if ( inRange )
		{
			head = wireHead( P, slot );
		}
		// dead slots (padding between the segments of a batch) have pointCount 0 and are not placed in any bin
		bool active = inRange && ( __float_as_int( head.z ) & kMetaPointMask ) != 0;
		int bits = isOverflow ? 0 : __float_as_int( head.z ) & kMetaGroupMask;
		int indexA = __float_as_int( head.x ), indexB = __float_as_int( head.y );
		int bin = active ? P.bodyBin[indexA >= 0 ? indexA : indexB] : -1;
		if ( active && indexA >= 0 && indexB >= 0 && P.bodyBin[indexB] != bin )
		{
			// the island hint is wrong (two bodies of one constraint in different islands): no partition, the grid-barrier
			// kernel solves the step
			*P.binFail = 1;
		}
		int position = aggregatedAdd( P.binColorStart, bin * kColorSlots, active );
		if ( active )
		{
			if ( position < P.binCapContacts )
			{
				P.binContactInfo[(size_t)bin * P.binCapContacts + position] = make_int4( slot, indexA, indexB, bits | ( c << kFlatColorShift ) );
			}
			else
			{
				*P.binFail = 1;
			}
			if ( isOverflow && atomicAdd( P.binColorStart + (size_t)bin * kColorSlots + 1, 1 ) >= kMaxBinOverflow )
			{
				*P.binFail = 1;
			}
		}
	} );
	forEachItem( P.jointCount, [&]( int j ) {
		bool active = j < P.jointCount;
		int bin = -1, c = 0;
		int2 bodies = make_int2( -1, -1 );
		if ( active )
		{
			while ( c < P.colorCount && ( c + 1 < P.colorCount ? P.colors[c + 1].jointStart : P.overflow.jointStart ) <= j )
			{
				c += 1;
			}
			const int* pair = jointIndexPair( reinterpret_cast<b2lJointSim*>( const_cast<uint8_t*>( P.rawJoints + (size_t)j * kJointStride ) ) );
			bodies = pair != nullptr ? make_int2( pair[0], pair[1] ) : make_int2( -1, -1 );
			int body = bodies.x >= 0 ? bodies.x : bodies.y;
			// a filter joint has no solver data: park it in bin 0, it is a no-op in every stage
			bin = body < 0 ? 0 : P.bodyBin[body];
			if ( bodies.x >= 0 && bodies.y >= 0 && P.bodyBin[bodies.y] != bin )
			{
				*P.binFail = 1; // wrong island hint, see the contacts above
			}
		}
		int position = aggregatedAdd( P.binJointStart, bin * kColorSlots, active );
		if ( active )
		{
			if ( position < P.binCapJoints )
			{
				P.binJointList[(size_t)bin * P.binCapJoints + position] = j | ( c << kFlatJointShift );
				P.binJointBodies[(size_t)bin * P.binCapJoints + position] = bodies; // for the island kernel's levelisation
			}
			else
			{
				*P.binFail = 1;
			}
			if ( c == P.colorCount && atomicAdd( P.binJointStart + (size_t)bin * kColorSlots + 1, 1 ) >= kMaxBinOverflow )
			{
				*P.binFail = 1;
			}
		}
	} );
}

// ---- the overflow colour of a bin, levelised ------------------------------------------------------------------------------
// The reference solves the overflow colour one constraint after the other (joints, then contacts, in array order:
// src/solver.c:1100-1101, src/joint.c:1576-1591, src/contact_solver.c:216-340).  Two constraints only interact through
// a DYNAMIC body they share (nothing else is written), so the sequence can be cut into levels: the level of a constraint
// is one more than the highest level among the earlier constraints it shares a dynamic body with.  Constraints of one
// level touch disjoint dynamic bodies and every pair that shares one keeps its order, hence solving the levels one after
// the other -- each level in parallel -- gives every body exactly the sequence of updates the reference gives it.
// Typical overflow sets (everything touching one kinematic or very crowded body) have one or two levels.
constexpr int kMaxOverflowItems = 2 * kMaxBinOverflow; // joints + contacts

struct alignas( 16 ) OverflowSchedule
{
	int bodyA[kMaxOverflowItems]; // dynamic bodies of item i (bin-local, 1-based), 0 = none
	int bodyB[kMaxOverflowItems];
	short level[kMaxOverflowItems];
	short order[kMaxOverflowItems];			// items sorted by (level, position in the sequence)
	short levelStart[kMaxOverflowItems + 2]; // first entry of `order` with a level >= L
	int levelCount;
	int changed;
};

// Called by all threads of the block that holds the bin's overflow constraints.  bodiesOf( i, a, b ) returns the dynamic
// bodies of the i-th constraint of the sequence.
template <typename BodiesOf> B2G_DEV void buildOverflowSchedule( OverflowSchedule& S, int itemCount, BodiesOf bodiesOf )
{
	for ( int i = (int)threadIdx.x; i < itemCount; i += (int)blockDim.x )
	{
		int a = 0, b = 0;
		bodiesOf( i, a, b );
		S.bodyA[i] = a;
		S.bodyB[i] = b;
		S.level[i] = 1;
	}
	if ( threadIdx.x == 0 )
	{
		S.levelCount = 0;
		S.levelStart[0] = (short)itemCount;
	}
	__syncthreads();
	// chaotic iteration towards the least fixed point: levels only grow, a sweep without a change ends it.  A deep chain
	// (every constraint touches the same dynamic body, e.g. the drum of the tumbler scene) gains nothing from levels and
	// would pay a barrier per constraint: give up after a few sweeps, the caller then solves the sequence on one thread.
	const int sweepLimit = itemCount / 4 + 2;
	// quick test first: a dynamic body shared by m constraints forces m levels
	if ( threadIdx.x == 0 )
	{
		S.changed = 0;
	}
	__syncthreads();
	for ( int i = (int)threadIdx.x; i < itemCount; i += (int)blockDim.x )
	{
		int a = S.bodyA[i], b = S.bodyB[i];
		int degreeA = 0, degreeB = 0;
		for ( int j = 0; j < itemCount; ++j )
		{
			int ja = S.bodyA[j], jb = S.bodyB[j];
			degreeA += ( a != 0 && ( a == ja || a == jb ) ) ? 1 : 0;
			degreeB += ( b != 0 && ( b == ja || b == jb ) ) ? 1 : 0;
		}
		atomicMax( &S.changed, degreeA > degreeB ? degreeA : degreeB );
	}
	__syncthreads();
	const bool hub = S.changed > sweepLimit;
	__syncthreads();
	for ( int sweep = hub ? sweepLimit + 1 : 0;; ++sweep )
	{
		if ( sweep > sweepLimit )
		{
			if ( threadIdx.x == 0 )
			{
				S.levelCount = -1;
			}
			__syncthreads();
			return;
		}
		if ( threadIdx.x == 0 )
		{
			S.changed = 0;
		}
		__syncthreads();
		for ( int i = (int)threadIdx.x; i < itemCount; i += (int)blockDim.x )
		{
			int a = S.bodyA[i], b = S.bodyB[i];
			int level = 1;
			for ( int j = 0; j < i; ++j )
			{
				int ja = S.bodyA[j], jb = S.bodyB[j];
				bool shares = ( a != 0 && ( a == ja || a == jb ) ) || ( b != 0 && ( b == ja || b == jb ) );
				int after = (int)S.level[j] + 1;
				level = shares && after > level ? after : level;
			}
			if ( level != (int)S.level[i] )
			{
				S.level[i] = (short)level;
				S.changed = 1;
			}
		}
		__syncthreads();
		bool done = S.changed == 0;
		__syncthreads();
		if ( done )
		{
			break;
		}
	}
	for ( int i = (int)threadIdx.x; i < itemCount; i += (int)blockDim.x )
	{
		int mine = S.level[i], rank = 0;
		for ( int j = 0; j < itemCount; ++j )
		{
			int other = S.level[j];
			rank += ( other < mine || ( other == mine && j < i ) ) ? 1 : 0;
		}
		S.order[rank] = (short)i;
		atomicMax( &S.levelCount, mine );
	}
	__syncthreads();
	for ( int level = (int)threadIdx.x; level <= S.levelCount + 1; level += (int)blockDim.x )
	{
		int below = 0;
		for ( int j = 0; j < itemCount; ++j )
		{
			below += (int)S.level[j] < level ? 1 : 0;
		}
		S.levelStart[level] = (short)below;
	}
	__syncthreads();
}

// One pass over the overflow colour: level by level, `sync` between the levels.  `holder` is false for the blocks of a
// cluster that do not hold the bin's overflow constraints (they only take part in the barriers).
template <typename FJ, typename FC, typename Sync>
B2G_DEV void overflowLevels( const OverflowSchedule& S, int levelCount, bool holder, int jointCount, int jointBegin, int contactBegin, FJ joint,
							 FC contact, Sync sync, int threads )
{
	if ( levelCount < 0 )
	{
		// deep chain: the sequence as it is, on one thread
		if ( holder && threadIdx.x == 0 )
		{
			int itemCount = (int)S.levelStart[0];
			for ( int item = 0; item < itemCount; ++item )
			{
				if ( item < jointCount )
				{
					joint( jointBegin + item );
				}
				else
				{
					contact( contactBegin + ( item - jointCount ) );
				}
			}
		}
		sync();
		return;
	}
	for ( int level = 1; level <= levelCount; ++level )
	{
		if ( holder )
		{
			int end = S.levelStart[level + 1];
			for ( int t = (int)S.levelStart[level] + (int)threadIdx.x; t < end; t += threads )
			{
				int item = S.order[t];
				if ( item < jointCount )
				{
					joint( jointBegin + item );
				}
				else
				{
					contact( contactBegin + ( item - jointCount ) );
				}
			}
		}
		sync();
	}
}

// ---- island kernel -------------------------------------------------------------------------------------------------------
template <typename F> B2G_DEV void forEachLocal( int itemCount, F f )
{
	for ( int i = (int)threadIdx.x; i < itemCount; i += (int)blockDim.x )
	{
		f( i );
	}
}

// the same for the first `threads` threads of the block only
template <typename F> B2G_DEV void forEachLocal( int itemCount, int threads, F f )
{
	for ( int i = (int)threadIdx.x; i < itemCount; i += threads )
	{
		f( i );
	}
}

// One colour of a bin: joints [r.x, r.y) are taken by the block's first threads, contacts [r.z, r.w) by its LAST threads
// (the last thread takes the first contact), so that joints and contacts of the colour run side by side in different
// warps as long as the block has a thread for each.  The header is one broadcast 16-byte shared load.
template <typename FJ, typename FC> B2G_DEV void forEachInLocalColor( int4 r, int threads, FJ joint, FC contact )
{
	for ( int k = r.x + (int)threadIdx.x; k < r.y; k += threads )
	{
		joint( k );
	}
	for ( int k = r.z + ( threads - 1 - (int)threadIdx.x ); k < r.w; k += threads )
	{
		contact( k );
	}
}

template <typename FJ, typename FC> B2G_DEV void forEachInLocalColor( int4 r, FJ joint, FC contact )
{
	forEachInLocalColor( r, (int)blockDim.x, joint, contact );
}

__global__ void __launch_bounds__( kIslandThreads, 1 ) b2gIslandKernel( const __grid_constant__ StepParams P )
{
	// launched as a programmatic dependent of b2gScatterKernel: everything below reads what that grid wrote (a no-op
	// after an ordinary launch)
	asm volatile( "griddepcontrol.wait;" ::: "memory" );
	if ( __ldcg( P.binFail ) != 0 )
	{
		if ( blockIdx.x == 0 && threadIdx.x == 0 )
		{
			*P.islandFailed = 1;
		}
		return; // some bin does not fit: the grid-barrier kernel solves this step
	}

	extern __shared__ __align__( 16 ) uint8_t smem[];
	__shared__ int colorStartC[kColorSlots];
	__shared__ int colorStartJ[kColorSlots];
	__shared__ int anyRestitution;
	__shared__ int overflowOrder[kMaxBinOverflow]; // scratch for ordering the bin's overflow constraints
	__shared__ int overflowOrderJ[kMaxBinOverflow]; // flat lists: the same for its overflow joints
	__shared__ int flatCursorC[kColorSlots], flatCursorJ[kColorSlots]; // flat lists: entries per colour, then the colours' cursors
	__shared__ int flatOverflowCount[2];
	__shared__ OverflowSchedule overflow;
	// the colours that are present in this bin, in order: { jointBegin, jointEnd, contactBegin, contactEnd }
	__shared__ __align__( 16 ) int4 passRange[kMaxColors];
	__shared__ int passCount;
	__shared__ int stageThreadCount; // threads that run the stage loops: enough for the bin's largest colour

	const int bin = (int)blockIdx.x;
	const int capB = P.capBodies, capC = P.capContacts, capJ = P.capJoints;

	// carve-up (all 16-byte aligned: capB, capC are multiples of 4)
	SolveView V;
	uint8_t* cursor = smem;
	V.vel = reinterpret_cast<float4*>( cursor );
	cursor += (size_t)( capB + 1 ) * sizeof( float4 );
	V.pos = reinterpret_cast<float4*>( cursor );
	cursor += (size_t)( capB + 1 ) * sizeof( float4 );
	V.bodyK = reinterpret_cast<float4*>( cursor );
	cursor += (size_t)capB * sizeof( float4 );
	V.cf = reinterpret_cast<float4*>( cursor );
	cursor += (size_t)CF_COUNT * capC * sizeof( float4 );
	V.cfStride = capC;
	V.joints = cursor;
	cursor += (size_t)capJ * kJointStride;
	V.cidx = reinterpret_cast<int2*>( cursor );
	cursor += (size_t)capC * sizeof( int2 );
	int* jointIndexOf = reinterpret_cast<int*>( cursor ); // [capJ] global joint index of every local joint
	cursor += (size_t)capJ * sizeof( int );
	V.angDamp = reinterpret_cast<float*>( cursor );
	cursor += (size_t)capB * sizeof( float );
	V.cmeta = reinterpret_cast<int*>( cursor );
	cursor += (size_t)capC * sizeof( int );
	int* wireSlot = reinterpret_cast<int*>( cursor );
	V.anyRestitution = &anyRestitution;
	V.clusterRun = 0;
	V.clusterMagic = 0;
	V.asyncBar = 0;

	const int bodyCount = P.binBodyCount[bin];
	const int* bodyList = P.binBodyList + (size_t)bin * capB;
	const int* contactList = P.binContactList + (size_t)bin * capC;
	const int* jointList = P.binJointList + (size_t)bin * capJ;

	StageClock clk;
	clk.start();
	long long begin = clk.last;

	// Flat lists (b2gScatterKernel): the bin's constraints arrive in arrival order, tagged with their colour; a counting
	// sort in shared memory puts them colour-major -- pass 1 counts (while the bodies are on their way), one thread turns
	// the counts into offsets, pass 2 hands out the slots.  The order inside a colour stays arbitrary (see the file
	// header); the overflow colour is ranked by wire slot / joint index, i.e. back into array order.
	// ... unless the step runs on the previous step's lists and that step wrote its order down (planRead): then the bin's
	// constraints are read in their final order, like the colour-major lists of the two-phase partition kernel
	const bool planned = P.planRead != 0;
	const bool flat = P.flatLists != 0 && !planned;
	const int4* contactInfo = planned ? P.planInfo + (size_t)bin * capC : P.binContactInfo + (size_t)bin * capC;
	constexpr int kFlatJointMask = ( 1 << kFlatJointShift ) - 1;
	int flatContacts = 0, flatJoints = 0;
	if ( flat )
	{
		if ( threadIdx.x < kColorSlots )
		{
			flatCursorC[threadIdx.x] = 0;
			flatCursorJ[threadIdx.x] = 0;
		}
		if ( threadIdx.x < 2 )
		{
			flatOverflowCount[threadIdx.x] = 0;
		}
		flatContacts = P.binColorStart[(size_t)bin * kColorSlots];
		flatJoints = P.binJointStart[(size_t)bin * kColorSlots];
		__syncthreads();
		forEachLocal( flatContacts, [&]( int k ) {
			int4 info = __ldg( contactInfo + k );
			int c = info.w >> kFlatColorShift;
			atomicAdd( &flatCursorC[c], 1 );
			// the wire record is read by the prepare pass, two block-wide barriers from here: start it on its way
			prefetchWire( P, info.x );
			if ( c == P.colorCount )
			{
				overflowOrder[atomicAdd( &flatOverflowCount[0], 1 )] = info.x;
			}
		} );
		forEachLocal( flatJoints, [&]( int k ) {
			int entry = __ldg( jointList + k );
			int c = entry >> kFlatJointShift;
			atomicAdd( &flatCursorJ[c], 1 );
			if ( c == P.colorCount )
			{
				overflowOrderJ[atomicAdd( &flatOverflowCount[1], 1 )] = entry & kFlatJointMask;
			}
		} );
	}
	else if ( threadIdx.x < kColorSlots )
	{
		const int* planStart = P.planStart + (size_t)bin * 2 * kColorSlots;
		colorStartC[threadIdx.x] = planned ? planStart[threadIdx.x] : P.binColorOffset[(size_t)bin * kColorSlots + threadIdx.x];
		colorStartJ[threadIdx.x] = planned ? planStart[kColorSlots + threadIdx.x] : P.binJointOffset[(size_t)bin * kColorSlots + threadIdx.x];
	}
	if ( threadIdx.x == 0 )
	{
		V.vel[0] = make_float4( 0.0f, 0.0f, 0.0f, __uint_as_float( 0u ) );
		V.pos[0] = make_float4( 0.0f, 0.0f, 1.0f, 0.0f );
		anyRestitution = 0;
	}
	forEachLocal( bodyCount, [&]( int i ) { loadBody( P, V, bodyList[i], i + 1 ); } );
	__syncthreads();
	// this block is the last reader of its bin's counters: leave them zeroed for the next step's partition kernel -- unless
	// the host may run the next step on these very lists (keepLists: a steady scene, see b2gEnqueueRun)
	if ( P.keepLists == 0 )
	{
		if ( threadIdx.x < kColorSlots )
		{
			P.binColorStart[(size_t)bin * kColorSlots + threadIdx.x] = 0;
			P.binJointStart[(size_t)bin * kColorSlots + threadIdx.x] = 0;
		}
		if ( threadIdx.x == 0 )
		{
			P.binBodyCount[bin] = 0;
		}
	}

	const int colorCount = P.colorCount;
	auto buildColourTable = [&]() {
	if ( threadIdx.x < 32 )
	{
		// one lane per colour slot: offsets of the colours (flat lists: exclusive scan of the counts), then the table of the
		// colours that are present in this bin
		const int lane = (int)threadIdx.x;
		int nC = 0, nJ = 0, startC = 0, startJ = 0;
		if ( flat )
		{
			nC = lane < kColorSlots ? flatCursorC[lane] : 0; // nothing was counted beyond the overflow bucket
			nJ = lane < kColorSlots ? flatCursorJ[lane] : 0;
			int sumC = nC, sumJ = nJ;
#pragma unroll
			for ( int d = 1; d < 32; d <<= 1 )
			{
				int upC = __shfl_up_sync( 0xffffffffu, sumC, d ), upJ = __shfl_up_sync( 0xffffffffu, sumJ, d );
				sumC += lane >= d ? upC : 0;
				sumJ += lane >= d ? upJ : 0;
			}
			startC = sumC - nC;
			startJ = sumJ - nJ;
			if ( lane < kColorSlots )
			{
				colorStartC[lane] = startC;
				colorStartJ[lane] = startJ;
				flatCursorC[lane] = startC;
				flatCursorJ[lane] = startJ;
			}
		}
		else if ( lane + 1 < kColorSlots )
		{
			startC = colorStartC[lane];
			startJ = colorStartJ[lane];
			nC = colorStartC[lane + 1] - startC;
			nJ = colorStartJ[lane + 1] - startJ;
		}
		const bool present = lane < colorCount && ( nC | nJ ) != 0; // a colour that is not present in this bin has nothing to order
		const unsigned presentMask = __ballot_sync( 0xffffffffu, present );
		if ( present )
		{
			passRange[__popc( presentMask & ( ( 1u << lane ) - 1u ) )] = make_int4( startJ, startJ + nJ, startC, startC + nC );
		}
		int widest = present ? roundUp32( nJ ) + roundUp32( nC ) : 32;
#pragma unroll
		for ( int d = 16; d > 0; d >>= 1 )
		{
			int other = __shfl_xor_sync( 0xffffffffu, widest, d );
			widest = other > widest ? other : widest;
		}
		if ( lane == 0 )
		{
			passCount = __popc( presentMask );
			if ( P.stageAllThreads != 2 )
			{
				// ... and for one body each in the two body stages of a sub-step (integrate-positions is a chain of an IEEE
				// square root and a division: a second round of it costs more than a few more warps at the barriers;
				// B2GPU_STAGE_ALL=2 sizes by the colours only)
				int bodies = roundUp32( bodyCount );
				widest = bodies > widest ? bodies : widest;
			}
			stageThreadCount = widest < (int)blockDim.x && P.stageAllThreads != 1 ? widest : (int)blockDim.x;
		}
	}
	};
	buildColourTable();
	if ( flat )
	{
		__syncthreads(); // the offsets
	}

	// Levels instead of colours.  The reference's colours are global: a colour stage of this bin must
	// wait for the previous one even when none of its constraints touches a body of that colour.  What bit-exactness needs is
	// only that every DYNAMIC BODY sees its contacts in colour order.  So the coloured contacts are levelised like the
	// overflow colour: walking the colours in order, level( contact ) = 1 + the highest level among the earlier contacts of
	// its dynamic bodies (kept per body).  Contacts of one level touch disjoint dynamic bodies, a body's contacts have
	// strictly increasing levels in colour order, hence solving level by level gives every body the reference's sequence
	// of updates -- in fewer stages (a base-10 pyramid: 7 levels for 8-9 colours).  From here on "colour c" of this bin
	// means level c + 1.
	constexpr int kOwnContacts = 3, kOwnJoints = 2; // constraints per thread kept in registers across the levelisation
	int ownLevel[kOwnContacts] = { 0, 0, 0 }, ownA[kOwnContacts] = { 0, 0, 0 }, ownB[kOwnContacts] = { 0, 0, 0 };
	int ownJointLevel[kOwnJoints] = { 0, 0 };
	const int2* jointBodies = P.binJointBodies + (size_t)bin * capJ;
	const bool levelise = flat && P.leveliseContacts != 0 && flatContacts <= kOwnContacts * (int)blockDim.x &&
						  flatJoints <= kOwnJoints * (int)blockDim.x && colorStartC[colorCount] + colorStartJ[colorCount] > 0 &&
						  (size_t)( bodyCount + 1 ) * sizeof( int ) <= (size_t)CF_COUNT * capC * sizeof( float4 );
	if ( levelise )
	{
		int* bodyLevel = reinterpret_cast<int*>( V.cf ); // scratch: the constraint fields are written by the prepare pass below
		for ( int i = (int)threadIdx.x; i <= bodyCount; i += (int)blockDim.x )
		{
			bodyLevel[i] = 0;
		}
		int ownColour[kOwnContacts];
#pragma unroll
		for ( int j = 0; j < kOwnContacts; ++j )
		{
			int k = (int)threadIdx.x + j * (int)blockDim.x;
			ownColour[j] = -1;
			if ( k < flatContacts )
			{
				int4 info = __ldg( contactInfo + k );
				int c = info.w >> kFlatColorShift;
				ownColour[j] = c < colorCount ? c : -1; // the overflow colour keeps its own bucket and order
				int localA = info.y >= 0 ? P.bodyLocal[info.y] : 0;
				int localB = info.z >= 0 ? P.bodyLocal[info.z] : 0;
				// only dynamic bodies order their contacts (nothing else is written)
				ownA[j] = ( __float_as_uint( V.vel[localA].w ) & B2L_FLAG_DYNAMIC ) != 0 ? localA : -localA;
				ownB[j] = ( __float_as_uint( V.vel[localB].w ) & B2L_FLAG_DYNAMIC ) != 0 ? localB : -localB;
			}
		}
		int ownJointColour[kOwnJoints], ownJA[kOwnJoints], ownJB[kOwnJoints];
#pragma unroll
		for ( int j = 0; j < kOwnJoints; ++j )
		{
			int k = (int)threadIdx.x + j * (int)blockDim.x;
			ownJointColour[j] = -1;
			ownJA[j] = ownJB[j] = 0;
			if ( k < flatJoints )
			{
				int c = __ldg( jointList + k ) >> kFlatJointShift;
				ownJointColour[j] = c < colorCount ? c : -1;
				int2 bodies = __ldg( jointBodies + k );
				int localA = bodies.x >= 0 ? P.bodyLocal[bodies.x] : 0;
				int localB = bodies.y >= 0 ? P.bodyLocal[bodies.y] : 0;
				ownJA[j] = ( __float_as_uint( V.vel[localA].w ) & B2L_FLAG_DYNAMIC ) != 0 ? localA : 0;
				ownJB[j] = ( __float_as_uint( V.vel[localB].w ) & B2L_FLAG_DYNAMIC ) != 0 ? localB : 0;
			}
		}
		const int overflowContacts = colorStartC[colorCount + 1] - colorStartC[colorCount];
		const int overflowJoints = colorStartJ[colorCount + 1] - colorStartJ[colorCount];
		__syncthreads();
		for ( int c = 0; c < colorCount; ++c )
		{
			if ( colorStartC[c + 1] == colorStartC[c] && colorStartJ[c + 1] == colorStartJ[c] )
			{
				continue; // colour not present in this bin (uniform for the block)
			}
#pragma unroll
			for ( int j = 0; j < kOwnJoints; ++j )
			{
				if ( ownJointColour[j] == c )
				{
					// joints and contacts of one colour touch disjoint dynamic bodies too (one graph colouring for both)
					int la = ownJA[j] > 0 ? bodyLevel[ownJA[j]] : 0;
					int lb = ownJB[j] > 0 ? bodyLevel[ownJB[j]] : 0;
					int level = 1 + ( la > lb ? la : lb );
					ownJointLevel[j] = level;
					if ( ownJA[j] > 0 )
					{
						bodyLevel[ownJA[j]] = level;
					}
					if ( ownJB[j] > 0 )
					{
						bodyLevel[ownJB[j]] = level;
					}
				}
			}
#pragma unroll
			for ( int j = 0; j < kOwnContacts; ++j )
			{
				if ( ownColour[j] == c )
				{
					// a dynamic body has at most one contact per colour: nobody else reads or writes these two entries now
					int la = ownA[j] > 0 ? bodyLevel[ownA[j]] : 0;
					int lb = ownB[j] > 0 ? bodyLevel[ownB[j]] : 0;
					int level = 1 + ( la > lb ? la : lb );
					ownLevel[j] = level;
					if ( ownA[j] > 0 )
					{
						bodyLevel[ownA[j]] = level;
					}
					if ( ownB[j] > 0 )
					{
						bodyLevel[ownB[j]] = level;
					}
				}
			}
			__syncthreads();
		}
		if ( threadIdx.x < kColorSlots )
		{
			flatCursorC[threadIdx.x] = (int)threadIdx.x == colorCount ? overflowContacts : 0;
			flatCursorJ[threadIdx.x] = (int)threadIdx.x == colorCount ? overflowJoints : 0;
		}
		__syncthreads();
#pragma unroll
		for ( int j = 0; j < kOwnContacts; ++j )
		{
			if ( ownLevel[j] > 0 )
			{
				atomicAdd( &flatCursorC[ownLevel[j] - 1], 1 );
			}
		}
#pragma unroll
		for ( int j = 0; j < kOwnJoints; ++j )
		{
			if ( ownJointLevel[j] > 0 )
			{
				atomicAdd( &flatCursorJ[ownJointLevel[j] - 1], 1 );
			}
		}
		__syncthreads();
		buildColourTable(); // the same table, of levels
		__syncthreads();
	}
	const int contactCount = colorStartC[kColorSlots - 1];
	const int jointCount = colorStartJ[kColorSlots - 1];
	// the overflow colour's constraints of this bin, solved by one thread in array order
	const int ovCb = colorStartC[colorCount], ovCe = colorStartC[colorCount + 1];
	const int ovJb = colorStartJ[colorCount], ovJe = colorStartJ[colorCount + 1];
	const bool hasOverflow = ( ovCe - ovCb ) + ( ovJe - ovJb ) > 0;

	// The atomics of the partition kernel scrambled the order inside a bucket.  That is harmless for a colour, but the
	// overflow colour is solved sequentially in ARRAY order (src/solver.c:1100-1101): rank-sort its part of the lists.
	auto orderedIndex = [&]( const int* list, int begin, int end, int k ) -> int {
		// returns the element of list[begin, end) whose rank is (k - begin) in ascending order
		return k < begin || k >= end ? list[k] : overflowOrder[k - begin];
	};
	// rank of `mine` among the n distinct values of `values`
	auto rankAmong = [&]( const int* values, int n, int mine ) -> int {
		int rank = 0;
		for ( int m = 0; m < n; ++m )
		{
			rank += values[m] < mine ? 1 : 0;
		}
		return rank;
	};
	if ( !flat && !planned && ovCe > ovCb )
	{
		for ( int k = ovCb + (int)threadIdx.x; k < ovCe; k += (int)blockDim.x )
		{
			int mine = contactList[k], rank = 0;
			for ( int m = ovCb; m < ovCe; ++m )
			{
				rank += contactList[m] < mine ? 1 : 0;
			}
			overflowOrder[rank] = mine;
		}
	}
	if ( !flat )
	{
		__syncthreads();
	}

	// prepare: contacts read the bodies' initial velocities from the view (identical to the wire states here)
	if ( flat )
	{
		auto prepareOwn = [&]( int k, int bucket, int localA, int localB ) {
			int4 info = __ldg( contactInfo + k );
			int c = info.w >> kFlatColorShift;
			bool wide = c != colorCount;
			int dest = wide ? atomicAdd( &flatCursorC[bucket >= 0 ? bucket : c], 1 ) : ovCb + rankAmong( overflowOrder, ovCe - ovCb, info.x );
			if ( localA < 0 )
			{
				localA = info.y >= 0 ? P.bodyLocal[info.y] : 0;
				localB = info.z >= 0 ? P.bodyLocal[info.z] : 0;
			}
			wireSlot[dest] = info.x;
			if ( P.planWrite != 0 )
			{
				P.planInfo[(size_t)bin * capC + dest] = make_int4( info.x, localA, localB, info.w & ( kMetaGroupRolling | kMetaGroupRestitution ) );
			}
			prepareContact( P, V, info.x, dest, localA, localB, V.vel[localA], V.vel[localB], wide,
							info.w & ( kMetaGroupRolling | kMetaGroupRestitution ) );
		};
		if ( P.planWrite != 0 && threadIdx.x < kColorSlots )
		{
			// (of the levels, if the bin was levelised: the table built last)
			int* planStart = P.planStart + (size_t)bin * 2 * kColorSlots;
			planStart[threadIdx.x] = colorStartC[threadIdx.x];
			planStart[kColorSlots + threadIdx.x] = colorStartJ[threadIdx.x];
		}
		if ( levelise )
		{
			// (the levelisation's scratch lives in the field array: all of it was read before the barrier above)
#pragma unroll
			for ( int j = 0; j < kOwnContacts; ++j )
			{
				int k = (int)threadIdx.x + j * (int)blockDim.x;
				if ( k < flatContacts )
				{
					int localA = ownA[j] < 0 ? -ownA[j] : ownA[j], localB = ownB[j] < 0 ? -ownB[j] : ownB[j];
					prepareOwn( k, ownLevel[j] - 1, localA, localB ); // level 0 = an overflow contact: bucket -1, unused
				}
			}
		}
		else
		{
			forEachLocal( flatContacts, [&]( int k ) { prepareOwn( k, -1, -1, -1 ); } );
		}
		auto placeJoint = [&]( int k, int bucket ) {
			int entry = __ldg( jointList + k );
			int c = entry >> kFlatJointShift, j = entry & kFlatJointMask;
			int dest = c != colorCount ? atomicAdd( &flatCursorJ[bucket >= 0 ? bucket : c], 1 ) : ovJb + rankAmong( overflowOrderJ, ovJe - ovJb, j );
			jointIndexOf[dest] = j;
			if ( P.planWrite != 0 )
			{
				P.planJoints[(size_t)bin * capJ + dest] = j;
			}
		};
		if ( levelise )
		{
#pragma unroll
			for ( int j = 0; j < kOwnJoints; ++j )
			{
				int k = (int)threadIdx.x + j * (int)blockDim.x;
				if ( k < flatJoints )
				{
					placeJoint( k, ownJointLevel[j] - 1 );
				}
			}
		}
		else
		{
			forEachLocal( flatJoints, [&]( int k ) { placeJoint( k, -1 ); } );
		}
	}
	else
	{
		forEachLocal( contactCount, [&]( int k ) {
			bool wide = k < ovCb || k >= ovCe;
			int slot, localA, localB, groupBits = 0;
			if ( planned || ( wide && P.resolveContacts != 0 ) )
			{
				int4 info = contactInfo[k]; // resolved by the partition kernel
				slot = info.x, localA = info.y, localB = info.z, groupBits = info.w;
			}
			else
			{
				slot = wide ? contactList[k] : overflowOrder[k - ovCb];
				groupBits = wide ? P.slotGroupBits[slot] : 0;
				float4 head = wireHead( P, slot );
				int indexA = __float_as_int( head.x );
				int indexB = __float_as_int( head.y );
				localA = indexA >= 0 ? P.bodyLocal[indexA] : 0;
				localB = indexB >= 0 ? P.bodyLocal[indexB] : 0;
			}
			wireSlot[k] = slot;
			prepareContact( P, V, slot, k, localA, localB, V.vel[localA], V.vel[localB], wide, groupBits );
		} );
	}
	__syncthreads();
	// joints: copy the prepared record into shared memory, renumber its bodies to the bin
	if ( planned )
	{
		forEachLocal( jointCount, [&]( int k ) { jointIndexOf[k] = P.planJoints[(size_t)bin * capJ + k]; } );
		__syncthreads();
	}
	else if ( !flat )
	{
		if ( ovJe > ovJb )
		{
			for ( int k = ovJb + (int)threadIdx.x; k < ovJe; k += (int)blockDim.x )
			{
				int mine = jointList[k], rank = 0;
				for ( int m = ovJb; m < ovJe; ++m )
				{
					rank += jointList[m] < mine ? 1 : 0;
				}
				overflowOrder[rank] = mine;
			}
		}
		__syncthreads();
		forEachLocal( jointCount, [&]( int k ) { jointIndexOf[k] = orderedIndex( jointList, ovJb, ovJe, k ); } );
		__syncthreads();
	}
	{
		const int quads = kJointStride / 16;
		for ( int t = (int)threadIdx.x; t < jointCount * quads; t += (int)blockDim.x )
		{
			int k = t / quads, q = t - k * quads;
			const float4* src = reinterpret_cast<const float4*>( P.rawJoints + (size_t)jointIndexOf[k] * kJointStride );
			reinterpret_cast<float4*>( V.joints + (size_t)k * kJointStride )[q] = src[q];
		}
	}
	__syncthreads();
	forEachLocal( jointCount, [&]( int k ) {
		int* pair = jointIndexPair( jointAt( V, k ) );
		if ( pair != nullptr )
		{
			int a = pair[0], b = pair[1];
			pair[0] = a >= 0 ? P.bodyLocal[a] - 1 : -1;
			pair[1] = b >= 0 ? P.bodyLocal[b] - 1 : -1;
		}
	} );
	__syncthreads();
	const int ovJoints = ovJe - ovJb;
	if ( hasOverflow )
	{
		buildOverflowSchedule( overflow, ovJoints + ( ovCe - ovCb ), [&]( int i, int& a, int& b ) {
			if ( i < ovJoints )
			{
				const int* pair = jointIndexPair( jointAt( V, ovJb + i ) );
				a = pair != nullptr ? pair[0] + 1 : 0; // bin-local, -1 = static
				b = pair != nullptr ? pair[1] + 1 : 0;
			}
			else
			{
				int2 idx = V.cidx[ovCb + ( i - ovJoints )];
				a = idx.x;
				b = idx.y;
			}
			a = ( __float_as_uint( V.vel[a].w ) & B2L_FLAG_DYNAMIC ) != 0 ? a : 0;
			b = ( __float_as_uint( V.vel[b].w ) & B2L_FLAG_DYNAMIC ) != 0 ? b : 0;
		} );
	}
	const int overflowLevelCount = hasOverflow ? overflow.levelCount : 0;
	// a deep chain of contacts is walked by one warp whose lanes take turns (overflowChainWarp) instead of by one thread
	const bool chainWalk = overflowLevelCount < 0 && ovJoints == 0;
	clk.lap( b2GpuStage_prepareConstraints );

	// The stage loops: a colour of a bin keeps only a few warps busy, and every other warp of the block would still walk
	// the loops and arrive at ~100 barriers.  Only the first `stageThreads` threads (whole warps, enough for the bin's
	// widest colour) take part; they meet at a named barrier of their own, the rest waits for the store phase.
	const int passes = passCount;
	const int stageThreads = stageThreadCount;
	auto blockSync = [&]() { asm volatile( "bar.sync 1, %0;" ::"r"( stageThreads ) : "memory" ); };
	// (Tried: the extra warps of the body stages waiting at a SECOND named barrier while the colours run: slower than
	// simply letting them walk the colour loops, 0.0607 -> 0.0623 ms on many_pyramids.)
	if ( (int)threadIdx.x < stageThreads )
	{
		for ( int subStep = 0; subStep < P.subStepCount; ++subStep )
		{
			forEachLocal( bodyCount, stageThreads, [&]( int i ) { integrateVelocities( V, i ); } );
			blockSync();
			clk.lap( b2GpuStage_integrateVelocities );

			if ( chainWalk )
			{
				if ( threadIdx.x < 32 )
				{
					overflowChainWarp<OV_WARM>( P, V, ovCb, ovCe );
				}
				blockSync();
			}
			else
			{
				overflowLevels(
					overflow, overflowLevelCount, true, ovJoints, ovJb, ovCb, [&]( int k ) { warmStartJoint( P, V, jointAt( V, k ) ); },
					[&]( int k ) { warmStartContactOverflow( V, k ); }, blockSync, stageThreads );
			}
			for ( int pass = 0; pass < passes; ++pass )
			{
				forEachInLocalColor(
					passRange[pass], stageThreads, [&]( int k ) { warmStartJoint( P, V, jointAt( V, k ) ); },
					[&]( int k ) { warmStartContact( V, k ); } );
				blockSync();
			}
			clk.lap( b2GpuStage_warmStart );

			if ( chainWalk )
			{
				if ( threadIdx.x < 32 )
				{
					overflowChainWarp<OV_SOLVE>( P, V, ovCb, ovCe );
				}
				blockSync();
			}
			else
			{
				overflowLevels(
					overflow, overflowLevelCount, true, ovJoints, ovJb, ovCb, [&]( int k ) { solveJoint( P, V, jointAt( V, k ), true ); },
					[&]( int k ) { solveContactOverflow( P, V, k, true ); }, blockSync, stageThreads );
			}
			for ( int pass = 0; pass < passes; ++pass )
			{
				forEachInLocalColor(
					passRange[pass], stageThreads,
					[&]( int k ) {
						b2lJointSim* joint = jointAt( V, k );
						solveJoint( P, V, joint, true );
						jointEventTest( P, joint );
					},
					[&]( int k ) { solveContact( P, V, k, true ); } );
				blockSync();
			}
			clk.lap( b2GpuStage_solveImpulses );

			forEachLocal( bodyCount, stageThreads, [&]( int i ) { integratePositions( P, V, i ); } );
			blockSync();
			clk.lap( b2GpuStage_integratePositions );

			if ( chainWalk )
			{
				if ( threadIdx.x < 32 )
				{
					overflowChainWarp<OV_RELAX>( P, V, ovCb, ovCe );
				}
				blockSync();
			}
			else
			{
				overflowLevels(
					overflow, overflowLevelCount, true, ovJoints, ovJb, ovCb, [&]( int k ) { solveJoint( P, V, jointAt( V, k ), false ); },
					[&]( int k ) { solveContactOverflow( P, V, k, false ); }, blockSync, stageThreads );
			}
			for ( int pass = 0; pass < passes; ++pass )
			{
				forEachInLocalColor(
					passRange[pass], stageThreads, [&]( int k ) { solveJoint( P, V, jointAt( V, k ), false ); },
					[&]( int k ) { solveContact( P, V, k, false ); } );
				blockSync();
			}
			clk.lap( b2GpuStage_relaxImpulses );
		}

		if ( anyRestitution != 0 )
		{
			if ( ovCe > ovCb && chainWalk )
			{
				if ( threadIdx.x < 32 )
				{
					overflowChainWarp<OV_RESTITUTION>( P, V, ovCb, ovCe );
				}
				blockSync();
			}
			else if ( ovCe > ovCb )
			{
				overflowLevels(
					overflow, overflowLevelCount, true, ovJoints, ovJb, ovCb, []( int ) {}, [&]( int k ) { restitutionContactOverflow( P, V, k ); },
					blockSync, stageThreads );
			}
			for ( int pass = 0; pass < passes; ++pass )
			{
				int4 r = passRange[pass];
				if ( r.z == r.w )
				{
					continue;
				}
				for ( int k = r.z + (int)threadIdx.x; k < r.w; k += stageThreads )
				{
					restitutionContact( P, V, k );
				}
				blockSync();
			}
		}
		clk.lap( b2GpuStage_applyRestitution );
	}
	__syncthreads();

	// store: impulses by wire slot, states by global body index, the joints' accumulated impulses by joint index
	forEachLocal( contactCount, [&]( int k ) { storeContact( P, V, k, wireSlot[k], k < ovCb || k >= ovCe ); } );
	forEachLocal( bodyCount, [&]( int i ) { storeBody( P, V, bodyList[i], i + 1 ); } );
	forEachLocal( jointCount, [&]( int k ) { storeJointImpulses( P, jointIndexOf[k], jointAt( V, k ) ); } );
	clk.lap( b2GpuStage_storeImpulses );

	if ( clk.lead )
	{
#pragma unroll
		for ( int i = 0; i < b2GpuStage_count; ++i )
		{
			P.stageCycles[i] = (unsigned long long)clk.acc[i];
		}
		P.stageCycles[8] = 0;
		P.stageCycles[9] = (unsigned long long)( clk.last - begin );
	}
}

// bytes of dynamic shared memory the island kernels need for the given per-block capacities; jointsResident = false:
// the joint records themselves stay in global memory (cluster kernel with spilled joints)
inline size_t islandSharedBytes( int capB, int capC, int capJ, bool jointsResident = true, bool liteJoints = false )
{
	size_t bytes = 0;
	bytes += 2 * (size_t)( capB + 1 ) * sizeof( float4 ); // vel, pos
	bytes += (size_t)capB * sizeof( float4 );			  // bodyK
	bytes += (size_t)CF_COUNT * capC * sizeof( float4 );  // contact fields
	bytes += jointsResident ? (size_t)capJ * ( liteJoints ? (size_t)kLiteJointWords * sizeof( float ) : (size_t)kJointStride ) : 0; // joints
	bytes += (size_t)capC * sizeof( int2 );				  // cidx
	bytes += (size_t)capJ * sizeof( int );				  // jointIndexOf
	bytes += (size_t)capB * sizeof( float );			  // angDamp
	bytes += 2 * (size_t)capC * sizeof( int );			  // cmeta, wireSlot
	return bytes;
}

} // namespace b2g
