// b2g_cluster.cuh -- island-local execution for islands that outgrow one thread block: a thread-block CLUSTER per bin.
//
// One big island (a 5000-body pyramid, the contents of a tumbler) cannot be split: every colour of it has to be
// finished before the next starts.  The grid-barrier kernel pays ~2 us per colour for that (1.2 us barrier + the L2
// round trips of the stage body).  Here up to 16 thread blocks of one cluster share the bin instead:
//   * the bin's bodies are dealt out in equal runs, a block keeps its run in shared memory and the other blocks of the
//     cluster read / write it through distributed shared memory (SolveView::clusterRun),
//   * every colour of the bin is dealt out evenly over the blocks, a block keeps its share of the constraints in its
//     own shared memory for the whole step,
//   * colours are separated by the hardware cluster barrier (~0.25 us measured, tools/microbench/barrier_bench.cu).
// The partition kernel is the same as for single-block bins (b2g_island.cuh), its lists are simply read in slices.
// The bin's overflow colour (sequential, array order) is solved by one thread of the cluster's first block.
#pragma once

#include "b2g_island.cuh"

namespace b2g
{

namespace cg = cooperative_groups;

// colour of local constraint k given the exclusive local offsets (localStart has slotCount + 1 entries)
B2G_DEV int colorOfLocal( const int* localStart, int slotCount, int k )
{
	int c = 0;
	while ( c + 1 < slotCount && localStart[c + 1] <= k )
	{
		c += 1;
	}
	return c;
}

__global__ void __launch_bounds__( kIslandThreads, 1 ) b2gClusterIslandKernel( const __grid_constant__ StepParams P )
{
	// launched as a programmatic dependent of the partition kernel: everything below reads what that grid wrote
	asm volatile( "griddepcontrol.wait;" ::: "memory" );
	if ( __ldcg( P.binFail ) != 0 )
	{
		if ( blockIdx.x == 0 && threadIdx.x == 0 )
		{
			*P.islandFailed = 1;
		}
		return; // some bin does not fit (uniform for the whole grid): the grid-barrier kernel solves this step
	}

	extern __shared__ __align__( 16 ) uint8_t smem[];
	// per colour slot (active colours, then the overflow bucket): first element of this block's share in the bin's
	// list, local slot of its first constraint (+ total at the end) and the bin-wide count (uniform for the cluster)
	__shared__ int listBeginC[kColorSlots], localStartC[kColorSlots + 1], binCountC[kColorSlots];
	__shared__ int listBeginJ[kColorSlots], localStartJ[kColorSlots + 1], binCountJ[kColorSlots];
	__shared__ int anyRestitution;
	__shared__ int clusterRestitution;
	// bytes of body writes this block receives in a pass over colour c (16 per dynamic body of mine that a contact of
	// the colour touches), counted by the writers during prepare
	__shared__ int expectBytes[kColorSlots];
	// ... counted here first, per owner block and colour, and sent to the owners in one remote atomic per (owner, colour):
	// thousands of constraints adding 16 at a time to the same few remote counters serialised the prologue (joint_grid:
	// 39 600 remote atomics on 96 addresses)
	__shared__ int expectLocal[16][kColorSlots];
	__shared__ __align__( 8 ) unsigned long long arrivalBar;
	__shared__ int overflowOrder[kMaxBinOverflow];
	__shared__ OverflowSchedule overflow; // of the bin's overflow colour (held by the cluster's first block)
	__shared__ int overflowCacheCount;	  // > 0: the bodies of a deep overflow chain are cached in the first block (see below)

	cg::cluster_group cluster = cg::this_cluster();
	const int share = P.clusterSize;
	const int rank = (int)cluster.block_rank();
	const int bin = (int)blockIdx.x / share;
	const int capB = P.capBodies, capC = P.capContacts, capJ = P.capJoints;
	const int colorCount = P.colorCount;
	const int slotCount = colorCount + 1; // + the overflow bucket

	// carve-up: identical to the single-block island kernel, with per-block capacities
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
	// joints: resident in shared memory, or -- spilled: a big jointed island whose records do not fit -- left in the
	// global working copy (L2); a record is only ever touched by the thread that owns the joint, the bodies stay in
	// distributed shared memory either way
	const bool jointsSpilled = P.jointsSpilled != 0;
	V.joints = jointsSpilled ? P.g.joints : cursor;
	V.jointLite = jointsSpilled ? 0 : P.liteJoints; // (the planner never combines the two)
	cursor += jointsSpilled ? 0 : (size_t)capJ * jointBytesOf( V.jointLite != 0 );
	V.cidx = reinterpret_cast<int2*>( cursor );
	cursor += (size_t)capC * sizeof( int2 );
	int* jointIndexOf = reinterpret_cast<int*>( cursor );
	cursor += (size_t)capJ * sizeof( int );
	V.angDamp = reinterpret_cast<float*>( cursor );
	cursor += (size_t)capB * sizeof( float );
	V.cmeta = reinterpret_cast<int*>( cursor );
	cursor += (size_t)capC * sizeof( int );
	int* wireSlot = reinterpret_cast<int*>( cursor );
	V.anyRestitution = &anyRestitution;
	V.clusterRun = P.clusterRun;
	V.clusterMagic = P.clusterMagic;
	V.asyncBar = 0;
	SolveView VA = V; // same view, body writes as counted st.async stores
	const unsigned barAddr = (unsigned)__cvta_generic_to_shared( &arrivalBar );
	VA.asyncBar = barAddr;
	unsigned barPhase = 0;
	// slot of local joint k in the view's joint array (spilled: the global working copy, indexed by the step's joint index)
	auto jointSlot = [&]( int k ) -> int { return jointsSpilled ? jointIndexOf[k] : k; };

	// this block's run of the bin's bodies
	const int binBodies = P.binBodyCount[bin];
	const int bodyBegin = min( binBodies, rank * P.clusterRun );
	const int bodyCount = min( binBodies - bodyBegin, P.clusterRun );
	const int* bodyList = P.binBodyList + (size_t)bin * P.binCapBodies + bodyBegin;
	// the constraint lists this block reads: its slice of the bin's lists, or -- owner lists -- its own list
	const bool ownerLists = P.ownerLists != 0;
	const int list = ownerLists ? rank : bin;
	const int* contactList = P.binContactList + (size_t)list * P.listCapContacts;
	const int4* contactInfo = P.binContactInfo + (size_t)list * P.listCapContacts;
	const int* jointList = P.binJointList + (size_t)list * P.listCapJoints;

	StageClock clk;
	clk.start();
	long long begin = clk.last;

	// the tables of this block's list: loaded by one thread per entry (a dependent global load each), combined by one
	__shared__ int rawStartC[kColorSlots], rawStartJ[kColorSlots], rawTotal[2 * kColorSlots];
	if ( threadIdx.x < kColorSlots )
	{
		rawStartC[threadIdx.x] = P.binColorOffset[(size_t)list * kColorSlots + threadIdx.x];
		rawStartJ[threadIdx.x] = P.binJointOffset[(size_t)list * kColorSlots + threadIdx.x];
		rawTotal[threadIdx.x] = ownerLists ? P.binColorTotal[threadIdx.x] : 0;
		rawTotal[kColorSlots + threadIdx.x] = ownerLists ? P.binColorTotal[kColorSlots + threadIdx.x] : 0;
	}
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		const int* startC = rawStartC;
		const int* startJ = rawStartJ;
		int localC = 0, localJ = 0;
		for ( int c = 0; c < slotCount; ++c )
		{
			bool isOverflow = c == colorCount;
			// even dealing: a slice of the bin's colour; owner lists: the whole colour of this block's own list
			bool whole = ownerLists || isOverflow;
			int s0 = startC[c], n = startC[c + 1] - s0;
			int lo = whole ? s0 : s0 + n * rank / share;
			int hi = whole ? ( ownerLists || rank == 0 ? s0 + n : s0 ) : s0 + n * ( rank + 1 ) / share;
			listBeginC[c] = lo;
			localStartC[c] = localC;
			binCountC[c] = ownerLists ? rawTotal[c] : n;
			localC += hi - lo;

			s0 = startJ[c], n = startJ[c + 1] - s0;
			lo = whole ? s0 : s0 + n * rank / share;
			hi = whole ? ( ownerLists || rank == 0 ? s0 + n : s0 ) : s0 + n * ( rank + 1 ) / share;
			listBeginJ[c] = lo;
			localStartJ[c] = localJ;
			binCountJ[c] = ownerLists ? rawTotal[kColorSlots + c] : n;
			localJ += hi - lo;
		}
		localStartC[slotCount] = localC;
		localStartJ[slotCount] = localJ;
		V.vel[0] = make_float4( 0.0f, 0.0f, 0.0f, __uint_as_float( 0u ) );
		V.pos[0] = make_float4( 0.0f, 0.0f, 1.0f, 0.0f );
		anyRestitution = 0;
		asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"( barAddr ) : "memory" );
		asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
	}
	if ( threadIdx.x < kColorSlots )
	{
		expectBytes[threadIdx.x] = 0;
	}
	for ( int t = (int)threadIdx.x; t < 16 * kColorSlots; t += (int)blockDim.x )
	{
		( &expectLocal[0][0] )[t] = 0;
	}
	forEachLocal( bodyCount, [&]( int i ) { loadBody( P, V, bodyList[i], i + 1 ); } );
	cluster.sync(); // every block of the cluster is running and its bodies are in place
	// every block has read the bin's counters: leave them zeroed for the next step's partition kernel -- unless the host may
	// run the next step on these very lists (keepLists, see b2gEnqueueRun)
	if ( P.keepLists == 0 && ( rank == 0 || ownerLists ) )
	{
		if ( threadIdx.x < kColorSlots )
		{
			P.binColorStart[(size_t)list * kColorSlots + threadIdx.x] = 0;
			P.binJointStart[(size_t)list * kColorSlots + threadIdx.x] = 0;
			if ( ownerLists && rank == 0 )
			{
				P.binColorTotal[threadIdx.x] = 0;
				P.binColorTotal[kColorSlots + threadIdx.x] = 0;
			}
		}
		if ( threadIdx.x == 0 && rank == 0 )
		{
			P.binBodyCount[bin] = 0;
		}
	}

	const int contactCount = localStartC[slotCount];
	const int jointCount = localStartJ[slotCount];
	// the bin's overflow colour lives in the first block; whether it exists is uniform for the cluster
	const bool hasOverflow = binCountC[colorCount] + binCountJ[colorCount] > 0;
	const int ovCb = localStartC[colorCount], ovCe = localStartC[slotCount];
	const int ovJb = localStartJ[colorCount], ovJe = localStartJ[slotCount];

	// overflow constraints in ARRAY order (src/solver.c:1100-1101): rank-sort this bucket of the bin's list
	auto sortOverflow = [&]( const int* list, int first, int count ) {
		for ( int k = (int)threadIdx.x; k < count; k += (int)blockDim.x )
		{
			int mine = list[first + k], order = 0;
			for ( int m = 0; m < count; ++m )
			{
				order += list[first + m] < mine ? 1 : 0;
			}
			overflowOrder[order] = mine;
		}
	};
	sortOverflow( contactList, listBeginC[colorCount], ovCe - ovCb );
	__syncthreads();

	forEachLocal( contactCount, [&]( int k ) {
		int c = colorOfLocal( localStartC, slotCount, k );
		bool wide = c < colorCount;
		int offset = k - localStartC[c];
		int slot, localA, localB, groupBits = 0;
		if ( wide && P.resolveContacts != 0 )
		{
			int4 info = contactInfo[listBeginC[c] + offset]; // resolved by the partition kernel
			slot = info.x, localA = info.y, localB = info.z, groupBits = info.w;
		}
		else
		{
			slot = wide ? contactList[listBeginC[c] + offset] : overflowOrder[offset];
			groupBits = wide ? P.slotGroupBits[slot] : 0;
			float4 head = wireHead( P, slot );
			int indexA = __float_as_int( head.x );
			int indexB = __float_as_int( head.y );
			localA = indexA >= 0 ? P.bodyLocal[indexA] : 0;
			localB = indexB >= 0 ? P.bodyLocal[indexB] : 0;
		}
		wireSlot[k] = slot;
		float4 sA = gatherVel( V, localA ), sB = gatherVel( V, localB );
		if ( wide )
		{
			// tell the owners of the two bodies what to expect from this contact in every pass over its colour
			if ( ( __float_as_uint( sA.w ) & B2L_FLAG_DYNAMIC ) != 0 )
			{
				atomicAdd( &expectLocal[clusterOwner( V, (unsigned)localA - 1u )][c], (int)sizeof( float4 ) );
			}
			if ( ( __float_as_uint( sB.w ) & B2L_FLAG_DYNAMIC ) != 0 )
			{
				atomicAdd( &expectLocal[clusterOwner( V, (unsigned)localB - 1u )][c], (int)sizeof( float4 ) );
			}
		}
		prepareContact( P, V, slot, k, localA, localB, sA, sB, wide, groupBits );
	} );
	__syncthreads();
	sortOverflow( jointList, listBeginJ[colorCount], ovJe - ovJb );
	__syncthreads();
	forEachLocal( jointCount, [&]( int k ) {
		int c = colorOfLocal( localStartJ, slotCount, k );
		int offset = k - localStartJ[c];
		jointIndexOf[k] = c < colorCount ? jointList[listBeginJ[c] + offset] : overflowOrder[offset];
	} );
	__syncthreads();
	if ( V.jointLite != 0 )
	{
		loadLiteRevolutes( P, V, jointCount, jointSlot, [&]( int k ) { return jointIndexOf[k]; } );
	}
	else
	{
		const int quads = kJointStride / 16;
		for ( int t = (int)threadIdx.x; t < jointCount * quads; t += (int)blockDim.x )
		{
			int k = t / quads, q = t - k * quads;
			const float4* src = reinterpret_cast<const float4*>( P.rawJoints + (size_t)jointIndexOf[k] * kJointStride );
			reinterpret_cast<float4*>( jointAt( V, jointSlot( k ) ) )[q] = src[q];
		}
	}
	__syncthreads();
	forEachLocal( jointCount, [&]( int k ) {
		int a, b;
		if ( jointSlotBodies( V, jointSlot( k ), a, b ) )
		{
			a = a >= 0 ? P.bodyLocal[a] - 1 : -1;
			b = b >= 0 ? P.bodyLocal[b] - 1 : -1;
			setJointSlotBodies( V, jointSlot( k ), a, b );
			// tell the owners of the two bodies what to expect from this joint in every pass over its colour: every joint
			// type writes both bodies back once per pass, except a pogo joint without a spring (src/pogo_joint.c:165,203)
			int c = colorOfLocal( localStartJ, slotCount, k );
			if ( c < colorCount && jointSlotWrites( V, jointSlot( k ) ) )
			{
				if ( a >= 0 && ( __float_as_uint( gatherVel( V, a + 1 ).w ) & B2L_FLAG_DYNAMIC ) != 0 )
				{
					atomicAdd( &expectLocal[clusterOwner( V, (unsigned)a )][c], (int)sizeof( float4 ) );
				}
				if ( b >= 0 && ( __float_as_uint( gatherVel( V, b + 1 ).w ) & B2L_FLAG_DYNAMIC ) != 0 )
				{
					atomicAdd( &expectLocal[clusterOwner( V, (unsigned)b )][c], (int)sizeof( float4 ) );
				}
			}
		}
	} );
	__syncthreads();
	for ( int t = (int)threadIdx.x; t < share * kColorSlots; t += (int)blockDim.x )
	{
		int owner = t / kColorSlots, c = t - owner * kColorSlots;
		int bytes = expectLocal[owner][c];
		if ( bytes != 0 )
		{
			atomicAdd( cluster.map_shared_rank( expectBytes, owner ) + c, bytes );
		}
	}
	const int ovJoints = ovJe - ovJb;
	if ( hasOverflow && rank == 0 )
	{
		__syncthreads();
		buildOverflowSchedule( overflow, ovJoints + ( ovCe - ovCb ), [&]( int i, int& a, int& b ) {
			if ( i < ovJoints )
			{
				jointSlotBodies( V, jointSlot( ovJb + i ), a, b ); // bin-local, -1 = static
				a += 1;
				b += 1;
			}
			else
			{
				int2 idx = V.cidx[ovCb + ( i - ovJoints )];
				a = idx.x;
				b = idx.y;
			}
			a = ( __float_as_uint( gatherVel( V, a ).w ) & B2L_FLAG_DYNAMIC ) != 0 ? a : 0;
			b = ( __float_as_uint( gatherVel( V, b ).w ) & B2L_FLAG_DYNAMIC ) != 0 ? b : 0;
		} );
	}
	// A deep overflow chain (every contact touches the same dynamic body: the drum of the tumbler scene) is walked by ONE
	// thread, and in a cluster each of its gathers and scatters is a distributed-shared-memory round trip (~215 cycles):
	// 0.42 us per contact.  The bodies of the chain are therefore cached in the first block for the duration of a pass --
	// filled and written back by all threads in parallel, the sequential walk in between only touches local shared
	// memory.  The cache lives in the part of the schedule that a deep chain does not use (bodyA .. order); the contacts'
	// body indices are replaced by cache slots (nothing else reads the indices of overflow contacts after this point).
	constexpr int kOverflowCacheSlots = ( 2 * kMaxOverflowItems * (int)sizeof( int ) + 2 * kMaxOverflowItems * (int)sizeof( short ) ) / 36;
	float4* const cacheVel = reinterpret_cast<float4*>( &overflow );
	float4* const cachePos = cacheVel + kOverflowCacheSlots;
	int* const cacheBody = reinterpret_cast<int*>( cachePos + kOverflowCacheSlots );
	static_assert( offsetof( OverflowSchedule, levelStart ) >= (size_t)kOverflowCacheSlots * 36, "the overflow cache overlaps the live part of the schedule" );
	if ( threadIdx.x == 0 )
	{
		overflowCacheCount = 0;
	}
	if ( hasOverflow && rank == 0 )
	{
		__syncthreads();
		const int chain = ovCe - ovCb, entries = 2 * chain;
		if ( overflow.levelCount < 0 && ovJoints == 0 && entries <= kMaxOverflowItems )
		{
			auto bodyOfEntry = [&]( int e ) -> int {
				int2 idx = V.cidx[ovCb + ( e >> 1 )];
				return ( e & 1 ) != 0 ? idx.y : idx.x;
			};
			// first occurrence of every body (0 = the static dummy keeps slot 0)
			short* firstFlag = overflow.levelStart + 1;
			int* slotOfFirst = reinterpret_cast<int*>( cacheVel ); // scratch until the cache is filled for the first time
			for ( int e = (int)threadIdx.x; e < entries; e += (int)blockDim.x )
			{
				int body = bodyOfEntry( e );
				bool first = body != 0;
				for ( int f = 0; f < e && first; ++f )
				{
					first = bodyOfEntry( f ) != body;
				}
				firstFlag[e] = first ? 1 : 0;
			}
			__syncthreads();
			for ( int e = (int)threadIdx.x; e < entries; e += (int)blockDim.x )
			{
				int before = 0;
				for ( int f = 0; f < e; ++f )
				{
					before += firstFlag[f];
				}
				slotOfFirst[e] = firstFlag[e] != 0 ? 1 + before : 0;
				if ( e == entries - 1 )
				{
					overflowCacheCount = before + firstFlag[e]; // distinct bodies
				}
			}
			__syncthreads();
			const int distinct = overflowCacheCount;
			int mySlot[( kMaxOverflowItems + kIslandThreads - 1 ) / kIslandThreads];
			int myBody[( kMaxOverflowItems + kIslandThreads - 1 ) / kIslandThreads];
			if ( distinct + 1 <= kOverflowCacheSlots )
			{
				int n = 0;
				for ( int e = (int)threadIdx.x; e < entries; e += (int)blockDim.x, ++n )
				{
					int body = bodyOfEntry( e ), slot = 0;
					for ( int f = 0; f <= e && body != 0 && slot == 0; ++f )
					{
						slot = bodyOfEntry( f ) == body ? slotOfFirst[f] : 0;
					}
					mySlot[n] = slot;
					myBody[n] = body;
				}
				__syncthreads(); // everybody has read the indices and the scratch
				n = 0;
				for ( int e = (int)threadIdx.x; e < entries; e += (int)blockDim.x, ++n )
				{
					int* pair = reinterpret_cast<int*>( V.cidx + ovCb + ( e >> 1 ) );
					pair[e & 1] = mySlot[n];
					if ( firstFlag[e] != 0 )
					{
						cacheBody[mySlot[n]] = myBody[n];
					}
				}
				if ( threadIdx.x == 0 )
				{
					cacheBody[0] = 0;
				}
				__syncthreads();
				if ( threadIdx.x == 0 )
				{
					cacheVel[0] = make_float4( 0.0f, 0.0f, 0.0f, __uint_as_float( 0u ) );
					cachePos[0] = make_float4( 0.0f, 0.0f, 1.0f, 0.0f );
				}
			}
			else if ( threadIdx.x == 0 )
			{
				overflowCacheCount = 0; // too many bodies: the chain is walked through distributed shared memory
			}
		}
	}
	cluster.sync();
	// the number of overflow levels is needed by every block of the cluster (they all take part in the barriers)
	const int overflowLevelCount = hasOverflow ? *cluster.map_shared_rank( &overflow.levelCount, 0 ) : 0;
	const int overflowCached = hasOverflow ? *cluster.map_shared_rank( &overflowCacheCount, 0 ) : 0;
	SolveView VO = V; // the first block's cache of a deep chain's bodies as a flat view
	VO.vel = cacheVel;
	VO.pos = cachePos;
	VO.clusterRun = 0;
	VO.clusterMagic = 0;
	auto clusterSync = [&]() { cluster.sync(); };
	// restitution is applied by the whole cluster or not at all
	if ( threadIdx.x < 32 )
	{
		int mine = (int)threadIdx.x < share ? *cluster.map_shared_rank( &anyRestitution, threadIdx.x ) : 0;
		unsigned any = __ballot_sync( 0xffffffffu, mine != 0 );
		if ( threadIdx.x == 0 )
		{
			clusterRestitution = any != 0u ? 1 : 0;
		}
	}
	__syncthreads();
	clk.lap( b2GpuStage_prepareConstraints );

	auto overflowPass = [&]( auto op, auto joint, auto contact ) {
		if ( overflowCached > 0 )
		{
			if ( rank == 0 )
			{
				for ( int slot = 1 + (int)threadIdx.x; slot <= overflowCached; slot += (int)blockDim.x )
				{
					cacheVel[slot] = gatherVel( V, cacheBody[slot] );
					cachePos[slot] = gatherPos( V, cacheBody[slot] );
				}
				__syncthreads();
				if ( threadIdx.x < 32 )
				{
					overflowChainWarp<decltype( op )::value>( P, VO, ovCb, ovCe ); // one warp, the lanes take turns
				}
				__syncthreads();
				for ( int slot = 1 + (int)threadIdx.x; slot <= overflowCached; slot += (int)blockDim.x )
				{
					scatterVel( V, cacheBody[slot], cacheVel[slot] ); // dynamic bodies only, like every scatter
				}
			}
			cluster.sync();
			return;
		}
		overflowLevels(
			overflow, overflowLevelCount, rank == 0, ovJoints, ovJb, ovCb, joint, [&]( int k ) { contact( V, k ); }, clusterSync, (int)blockDim.x );
	};
	// A pass over the colours.  A full cluster barrier (release / acquire) makes every writer fence its remote stores at
	// GPU scope (MEMBAR.ALL.GPU), which costs more than the colour itself.  The colours use counted stores instead: the
	// owner of a body waits until the bytes announced for the colour have landed in its shared memory (mbarrier
	// transaction count) and only then joins a barrier that carries no fence.
	auto colorPass = [&]( auto joint, auto contact ) {
		for ( int c = 0; c < colorCount; ++c )
		{
			if ( binCountJ[c] == 0 && binCountC[c] == 0 )
			{
				continue; // colour not present in this bin (uniform for the cluster)
			}
			if ( threadIdx.x == 0 )
			{
				int bytes = expectBytes[c];
				if ( bytes > 0 )
				{
					asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( barAddr ), "r"( bytes ) : "memory" );
				}
				else
				{
					asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( barAddr ) : "memory" );
				}
			}
			forEachInLocalColor(
				make_int4( localStartJ[c], localStartJ[c + 1], localStartC[c], localStartC[c + 1] ), [&]( int k ) { joint( VA, k ); },
				[&]( int k ) { contact( VA, k ); } );
			// ONE warp waits for the announced bytes (a try_wait with cluster-scope acquire carries an L1 invalidation: with all
			// 16 warps spinning, the CCTL of the idle ones was 30 % of the kernel's stall samples and competed with the warps
			// that still had constraints to solve); the others go straight to the cluster barrier, which cannot complete before
			// the waiting warp of every block has arrived
			unsigned done = threadIdx.x < 32 ? 0u : 1u;
			for ( int spin = 0; done == 0; ++spin )
			{
				asm volatile( "{\n"
							  ".reg .pred p;\n"
							  "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
							  "selp.u32 %0, 1, 0, p;\n"
							  "}\n"
							  : "=r"( done )
							  : "r"( barAddr ), "r"( barPhase )
							  : "memory" );
				if ( spin > ( 1 << 22 ) )
				{
					__trap(); // the announced bytes never arrived: a bug, fail loudly instead of hanging the GPU
				}
			}
			barPhase ^= 1u;
			__syncwarp();
			asm volatile( "barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory" );
		}
	};

	for ( int subStep = 0; subStep < P.subStepCount; ++subStep )
	{
		forEachLocal( bodyCount, [&]( int i ) { integrateVelocities( V, i ); } );
		cluster.sync();
		clk.lap( b2GpuStage_integrateVelocities );

		overflowPass( std::integral_constant<int, OV_WARM>{}, [&]( int k ) { warmStartJointSlot( P, V, V, jointSlot( k ) ); },
					  [&]( const SolveView& view, int k ) { warmStartContactOverflow( view, k ); } );
		colorPass( [&]( const SolveView& view, int k ) { warmStartJointSlot( P, view, V, jointSlot( k ) ); },
				   [&]( const SolveView& view, int k ) { warmStartContact( view, k ); } );
		clk.lap( b2GpuStage_warmStart );

		overflowPass( std::integral_constant<int, OV_SOLVE>{}, [&]( int k ) { solveJointSlot( P, V, V, jointSlot( k ), true, false ); },
					  [&]( const SolveView& view, int k ) { solveContactOverflow( P, view, k, true ); } );
		colorPass(
			[&]( const SolveView& view, int k ) { solveJointSlot( P, view, V, jointSlot( k ), true, true ); },
			[&]( const SolveView& view, int k ) { solveContact( P, view, k, true ); } );
		clk.lap( b2GpuStage_solveImpulses );

		forEachLocal( bodyCount, [&]( int i ) { integratePositions( P, V, i ); } );
		cluster.sync();
		clk.lap( b2GpuStage_integratePositions );

		overflowPass( std::integral_constant<int, OV_RELAX>{}, [&]( int k ) { solveJointSlot( P, V, V, jointSlot( k ), false, false ); },
					  [&]( const SolveView& view, int k ) { solveContactOverflow( P, view, k, false ); } );
		colorPass( [&]( const SolveView& view, int k ) { solveJointSlot( P, view, V, jointSlot( k ), false, false ); },
				   [&]( const SolveView& view, int k ) { solveContact( P, view, k, false ); } );
		clk.lap( b2GpuStage_relaxImpulses );
	}

	if ( clusterRestitution != 0 )
	{
		if ( binCountC[colorCount] > 0 )
		{
			overflowPass( std::integral_constant<int, OV_RESTITUTION>{}, []( int ) {},
						  [&]( const SolveView& view, int k ) { restitutionContactOverflow( P, view, k ); } );
		}
		for ( int c = 0; c < colorCount; ++c )
		{
			if ( binCountC[c] == 0 )
			{
				continue;
			}
			for ( int k = localStartC[c] + (int)threadIdx.x; k < localStartC[c + 1]; k += (int)blockDim.x )
			{
				restitutionContact( P, V, k );
			}
			cluster.sync();
		}
	}
	clk.lap( b2GpuStage_applyRestitution );

	// store: everything a block needs is in its own shared memory again
	forEachLocal( contactCount, [&]( int k ) { storeContact( P, V, k, wireSlot[k], k < ovCb || k >= ovCe ); } );
	forEachLocal( bodyCount, [&]( int i ) { storeBody( P, V, bodyList[i], i + 1 ); } );
	forEachLocal( jointCount, [&]( int k ) { storeJointSlot( P, V, jointIndexOf[k], jointSlot( k ) ); } );
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
	cluster.sync(); // no block leaves while a peer may still look at its shared memory
}

} // namespace b2g
