// b2g_island.cuh -- island-local execution: zero grid barriers.
//
// Constraints only couple bodies of the same simulation island (reference src/island.c: b2LinkContact / b2LinkJoint
// merge the islands of the two bodies, static bodies belong to none), so islands can be solved independently and the
// colour order only has to be respected INSIDE an island.  The host packs the awake islands into `binCount` bins
// (b2GpuStepDesc::bodyIsland -> bodyBin); here
//   * b2gPartitionKernel buckets bodies, contacts and joints by (bin, colour) with atomics (three grid barriers), and
//   * b2gIslandKernel gives each bin to ONE thread block that keeps the bin's body state, contact constraints and
//     joints in shared memory for the whole step and separates colours with __syncthreads() (~10 ns) instead of a
//     grid-wide barrier (~1.2 us measured, tools/microbench/barrier_bench.cu).
// The order of constraints inside a (bin, colour) bucket comes from atomics and is not deterministic, but constraints
// of one colour touch disjoint dynamic bodies (src/constraint_graph.c:84-133), so every body sees exactly the same
// sequence of updates as in the reference: results are bit-identical (tests/test_gpu_lockstep.py).  The SIMD-group
// early-outs of the wide path depend on the reference's array order; they are evaluated in wire order by the
// partition kernel and travel with the constraint (kMetaGroup* bits).
// When a bin does not fit its shared-memory budget the partition kernel raises binFail and the grid-barrier kernel
// (b2g_solver.cu) solves the step instead; the overflow colour (strictly sequential) also uses that kernel.
#pragma once

#include "b2g_stages.cuh"

namespace b2g
{

constexpr int kIslandThreads = 512;
constexpr int kColorSlots = kMaxColors + 1; // per-bin offsets: colours + total

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

// ---- partition ---------------------------------------------------------------------------------------------------------
// counters (binBodyCount, binColorStart, binJointStart, binFail) are zeroed by the host before the launch
__global__ void __launch_bounds__( kBlockThreads, 1 ) b2gPartitionKernel( const __grid_constant__ StepParams P )
{
	const unsigned lane = threadIdx.x & 31u;
	const unsigned blocks = gridDim.x;

	// phase 1: local index of every body in its bin, (bin, rank) of every constraint, SIMD-group bits in wire order
	forEachItem( P.jointWords, [&]( int i ) {
		if ( i < P.jointWords )
		{
			P.jointBits[i] = 0u;
		}
	} );
	forEachItem( P.bodyCount, [&]( int b ) {
		if ( b < P.bodyCount )
		{
			int bin = P.bodyBin[b];
			int local = atomicAdd( P.binBodyCount + bin, 1 );
			if ( local < P.capBodies )
			{
				P.binBodyList[(size_t)bin * P.capBodies + local] = b;
			}
			else
			{
				*P.binFail = 1;
			}
			P.bodyLocal[b] = local + 1;
		}
	} );
	for ( int c = 0; c < P.colorCount; ++c )
	{
		ColorRange color = P.colors[c];
		forEachItem( color.contactCount, [&]( int i ) {
			bool active = i < color.contactCount;
			int slot = color.contactStart + i;
			int bits = simdGroupBits( P, slot, active, lane );
			if ( active )
			{
				float4 head = P.wire[(size_t)slot * WR_COUNT + WR_HEAD];
				int indexA = __float_as_int( head.x );
				int indexB = __float_as_int( head.y );
				int bin = P.bodyBin[indexA >= 0 ? indexA : indexB];
				int rank = atomicAdd( P.binColorStart + (size_t)bin * kColorSlots + c, 1 );
				P.contactBinRank[slot] = make_int2( bin, rank );
				P.slotGroupBits[slot] = bits;
			}
		} );
		forEachItem( color.jointCount, [&]( int i ) {
			if ( i < color.jointCount )
			{
				int j = color.jointStart + i;
				int body = jointBodyForBin( P.rawJoints + (size_t)j * kJointStride );
				// a filter joint has no solver data: park it in bin 0, it is a no-op in every stage
				int bin = body >= 0 ? P.bodyBin[body] : 0;
				int rank = atomicAdd( P.binJointStart + (size_t)bin * kColorSlots + c, 1 );
				P.jointBinRank[j] = make_int2( bin, rank );
			}
		} );
	}
	gridBarrier( P.barrier + 1, blocks );

	// phase 2: per bin, counts -> exclusive offsets (colour-major layout of the bin's lists), capacity check
	forEachItem( P.binCount, [&]( int bin ) {
		if ( bin < P.binCount )
		{
			int* contactStart = P.binColorStart + (size_t)bin * kColorSlots;
			int* jointStart = P.binJointStart + (size_t)bin * kColorSlots;
			int contacts = 0, joints = 0;
			for ( int c = 0; c < P.colorCount; ++c )
			{
				int n = contactStart[c];
				contactStart[c] = contacts;
				contacts += n;
				int m = jointStart[c];
				jointStart[c] = joints;
				joints += m;
			}
			for ( int c = P.colorCount; c < kColorSlots; ++c )
			{
				contactStart[c] = contacts;
				jointStart[c] = joints;
			}
			if ( contacts > P.capContacts || joints > P.capJoints )
			{
				*P.binFail = 1;
			}
		}
	} );
	gridBarrier( P.barrier + 1, 2 * blocks );
	if ( __ldcg( P.binFail ) != 0 )
	{
		return;
	}

	// phase 3: place every constraint in its bin's list
	for ( int c = 0; c < P.colorCount; ++c )
	{
		ColorRange color = P.colors[c];
		forEachItem( color.contactCount, [&]( int i ) {
			if ( i < color.contactCount )
			{
				int slot = color.contactStart + i;
				int2 br = P.contactBinRank[slot];
				int dest = P.binColorStart[(size_t)br.x * kColorSlots + c] + br.y;
				P.binContactList[(size_t)br.x * P.capContacts + dest] = slot;
			}
		} );
		forEachItem( color.jointCount, [&]( int i ) {
			if ( i < color.jointCount )
			{
				int j = color.jointStart + i;
				int2 br = P.jointBinRank[j];
				int dest = P.binJointStart[(size_t)br.x * kColorSlots + c] + br.y;
				P.binJointList[(size_t)br.x * P.capJoints + dest] = j;
			}
		} );
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

// joints on the first warps, contacts from the next multiple of 32: a warp never mixes the two kinds
template <typename FJ, typename FC> B2G_DEV void forEachInLocalColor( int jointBegin, int jointEnd, int contactBegin, int contactEnd, FJ joint,
																	   FC contact )
{
	int jointCount = jointEnd - jointBegin;
	int jointSpan = roundUp32( jointCount );
	int itemCount = jointSpan + ( contactEnd - contactBegin );
	for ( int t = (int)threadIdx.x; t < itemCount; t += (int)blockDim.x )
	{
		if ( t < jointSpan )
		{
			if ( t < jointCount )
			{
				joint( jointBegin + t );
			}
		}
		else
		{
			contact( contactBegin + ( t - jointSpan ) );
		}
	}
}

__global__ void __launch_bounds__( kIslandThreads, 1 ) b2gIslandKernel( const __grid_constant__ StepParams P )
{
	if ( __ldcg( P.binFail ) != 0 )
	{
		return; // some bin does not fit: the grid-barrier kernel solves this step
	}

	extern __shared__ __align__( 16 ) uint8_t smem[];
	__shared__ int colorStartC[kColorSlots];
	__shared__ int colorStartJ[kColorSlots];
	__shared__ int anyRestitution;

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
	int2* jointGlobal = reinterpret_cast<int2*>( cursor ); // the joints' global body indices, restored at the end
	cursor += (size_t)capJ * sizeof( int2 );
	V.angDamp = reinterpret_cast<float*>( cursor );
	cursor += (size_t)capB * sizeof( float );
	V.cmeta = reinterpret_cast<int*>( cursor );
	cursor += (size_t)capC * sizeof( int );
	int* wireSlot = reinterpret_cast<int*>( cursor );
	V.anyRestitution = &anyRestitution;

	const int bodyCount = P.binBodyCount[bin];
	const int* bodyList = P.binBodyList + (size_t)bin * capB;
	const int* contactList = P.binContactList + (size_t)bin * capC;
	const int* jointList = P.binJointList + (size_t)bin * capJ;

	StageClock clk;
	clk.start();
	long long begin = clk.last;

	if ( threadIdx.x < kColorSlots )
	{
		colorStartC[threadIdx.x] = P.binColorStart[(size_t)bin * kColorSlots + threadIdx.x];
		colorStartJ[threadIdx.x] = P.binJointStart[(size_t)bin * kColorSlots + threadIdx.x];
	}
	if ( threadIdx.x == 0 )
	{
		V.vel[0] = make_float4( 0.0f, 0.0f, 0.0f, __uint_as_float( 0u ) );
		V.pos[0] = make_float4( 0.0f, 0.0f, 1.0f, 0.0f );
		anyRestitution = 0;
	}
	forEachLocal( bodyCount, [&]( int i ) { loadBody( P, V, bodyList[i], i + 1 ); } );
	__syncthreads();

	const int contactCount = colorStartC[kMaxColors];
	const int jointCount = colorStartJ[kMaxColors];
	const int colorCount = P.colorCount;

	// prepare: contacts read the bodies' initial velocities from the view (identical to the wire states here)
	forEachLocal( contactCount, [&]( int k ) {
		int slot = contactList[k];
		float4 head = P.wire[(size_t)slot * WR_COUNT + WR_HEAD];
		int indexA = __float_as_int( head.x );
		int indexB = __float_as_int( head.y );
		int localA = indexA >= 0 ? P.bodyLocal[indexA] : 0;
		int localB = indexB >= 0 ? P.bodyLocal[indexB] : 0;
		wireSlot[k] = slot;
		prepareContact( P, V, slot, k, localA, localB, V.vel[localA], V.vel[localB], true, P.slotGroupBits[slot] );
	} );
	// joints: copy the prepared record into shared memory, renumber its bodies to the bin
	{
		const int quads = kJointStride / 16;
		for ( int t = (int)threadIdx.x; t < jointCount * quads; t += (int)blockDim.x )
		{
			int k = t / quads, q = t - k * quads;
			const float4* src = reinterpret_cast<const float4*>( P.rawJoints + (size_t)jointList[k] * kJointStride );
			reinterpret_cast<float4*>( V.joints + (size_t)k * kJointStride )[q] = src[q];
		}
	}
	__syncthreads();
	forEachLocal( jointCount, [&]( int k ) {
		int* pair = jointIndexPair( jointAt( V, k ) );
		if ( pair != nullptr )
		{
			int a = pair[0], b = pair[1];
			jointGlobal[k] = make_int2( a, b );
			pair[0] = a >= 0 ? P.bodyLocal[a] - 1 : -1;
			pair[1] = b >= 0 ? P.bodyLocal[b] - 1 : -1;
		}
	} );
	__syncthreads();
	clk.lap( b2GpuStage_prepareConstraints );

	for ( int subStep = 0; subStep < P.subStepCount; ++subStep )
	{
		forEachLocal( bodyCount, [&]( int i ) { integrateVelocities( V, i ); } );
		__syncthreads();
		clk.lap( b2GpuStage_integrateVelocities );

		for ( int c = 0; c < colorCount; ++c )
		{
			int jb = colorStartJ[c], je = colorStartJ[c + 1], cb = colorStartC[c], ce = colorStartC[c + 1];
			if ( jb == je && cb == ce )
			{
				continue; // colour not present in this bin: nothing to order (uniform for the block)
			}
			forEachInLocalColor(
				jb, je, cb, ce, [&]( int k ) { warmStartJoint( P, V, jointAt( V, k ) ); }, [&]( int k ) { warmStartContact( V, k ); } );
			__syncthreads();
		}
		clk.lap( b2GpuStage_warmStart );

		for ( int c = 0; c < colorCount; ++c )
		{
			int jb = colorStartJ[c], je = colorStartJ[c + 1], cb = colorStartC[c], ce = colorStartC[c + 1];
			if ( jb == je && cb == ce )
			{
				continue;
			}
			forEachInLocalColor(
				jb, je, cb, ce,
				[&]( int k ) {
					b2lJointSim* joint = jointAt( V, k );
					solveJoint( P, V, joint, true );
					jointEventTest( P, joint );
				},
				[&]( int k ) { solveContact( P, V, k, true ); } );
			__syncthreads();
		}
		clk.lap( b2GpuStage_solveImpulses );

		forEachLocal( bodyCount, [&]( int i ) { integratePositions( P, V, i ); } );
		__syncthreads();
		clk.lap( b2GpuStage_integratePositions );

		for ( int c = 0; c < colorCount; ++c )
		{
			int jb = colorStartJ[c], je = colorStartJ[c + 1], cb = colorStartC[c], ce = colorStartC[c + 1];
			if ( jb == je && cb == ce )
			{
				continue;
			}
			forEachInLocalColor(
				jb, je, cb, ce, [&]( int k ) { solveJoint( P, V, jointAt( V, k ), false ); },
				[&]( int k ) { solveContact( P, V, k, false ); } );
			__syncthreads();
		}
		clk.lap( b2GpuStage_relaxImpulses );
	}

	if ( anyRestitution != 0 )
	{
		for ( int c = 0; c < colorCount; ++c )
		{
			int cb = colorStartC[c], ce = colorStartC[c + 1];
			if ( cb == ce )
			{
				continue;
			}
			for ( int k = cb + (int)threadIdx.x; k < ce; k += (int)blockDim.x )
			{
				restitutionContact( P, V, k );
			}
			__syncthreads();
		}
	}
	clk.lap( b2GpuStage_applyRestitution );

	// store: impulses by wire slot, states by global body index, joints with their global body indices restored
	forEachLocal( contactCount, [&]( int k ) { storeContact( P, V, k, wireSlot[k], true ); } );
	forEachLocal( bodyCount, [&]( int i ) { storeBody( P, V, bodyList[i], i + 1 ); } );
	forEachLocal( jointCount, [&]( int k ) {
		int* pair = jointIndexPair( jointAt( V, k ) );
		if ( pair != nullptr )
		{
			pair[0] = jointGlobal[k].x;
			pair[1] = jointGlobal[k].y;
		}
	} );
	__syncthreads();
	{
		const int quads = kJointStride / 16;
		for ( int t = (int)threadIdx.x; t < jointCount * quads; t += (int)blockDim.x )
		{
			int k = t / quads, q = t - k * quads;
			float4* dst = reinterpret_cast<float4*>( P.g.joints + (size_t)jointList[k] * kJointStride );
			dst[q] = reinterpret_cast<const float4*>( V.joints + (size_t)k * kJointStride )[q];
		}
	}
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

// bytes of dynamic shared memory the island kernel needs for the given capacities
inline size_t islandSharedBytes( int capB, int capC, int capJ )
{
	size_t bytes = 0;
	bytes += 2 * (size_t)( capB + 1 ) * sizeof( float4 ); // vel, pos
	bytes += (size_t)capB * sizeof( float4 );			  // bodyK
	bytes += (size_t)CF_COUNT * capC * sizeof( float4 );  // contact fields
	bytes += (size_t)capJ * kJointStride;				  // joints
	bytes += (size_t)capC * sizeof( int2 );				  // cidx
	bytes += (size_t)capJ * sizeof( int2 );				  // jointGlobal
	bytes += (size_t)capB * sizeof( float );			  // angDamp
	bytes += 2 * (size_t)capC * sizeof( int );			  // cmeta, wireSlot
	return bytes;
}

} // namespace b2g
