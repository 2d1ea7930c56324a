// b2g_grid.cuh -- the grid-barrier kernels.
//
// b2gStepKernel: ONE persistent cooperative kernel runs the whole b2SolverTask stage sequence (reference
// src/solver.c:1055-1197) with the solver state in HBM / L2 and a grid-wide barrier where the reference's orchestrator
// spins on stage->completionCount (src/solver.c:999-1005).  It takes the steps the island-local kernels cannot: no island
// hint, an island too big for a 16-block cluster, a bin that did not fit (binFail).
// b2gStageKernel: the same device functions, one launch per stage (mode 1: tests, per-stage profiling).
#pragma once

#include "b2g_stages.cuh"

namespace b2g
{

// The whole step.  Stage order and barrier placement = b2SolverTask (src/solver.c:1055-1197); the stage timers
// are the reference's b2Profile split (src/solver.c:1080,1097,1112,1132,1141,1159,1182,1191).
__global__ void __launch_bounds__( kBlockThreads, 1 ) b2gStepKernel( const __grid_constant__ StepParams P )
{
	if ( P.binCount > 0 && __ldcg( P.binFail ) == 0 )
	{
		return; // the island kernel solved this step
	}

	// this block's joints stay in shared memory for the whole step (JointCache, b2g_stages.cuh)
	extern __shared__ __align__( 16 ) uint8_t gridSmem[];
	__shared__ JointCache jointCache;
	if ( threadIdx.x == 0 )
	{
		jointCache.base = gridSmem;
		jointCache.capacity = P.gridJointCache;
		int slots = 0;
		for ( int c = 0; c < P.colorCount; ++c )
		{
			jointCache.colorBase[c] = slots;
			slots += jointCacheSlots( P.colors[c].jointCount );
		}
		jointCache.colorBase[P.colorCount] = slots;
	}
	__syncthreads();
	const JointCache* cache = P.gridJointCache > 0 ? &jointCache : nullptr;

	unsigned int epoch = 0;
	const unsigned int blocks = gridDim.x;
	auto sync = [&]() {
		epoch += 1;
		gridBarrier( P.barrier, epoch * blocks );
	};

	StageClock clk;
	clk.start();
	long long begin = clk.last;

	const bool hasOverflow = P.overflow.contactCount + P.overflow.jointCount > 0;
	const int colorCount = P.colorCount;

	runStage( P, OP_PREPARE, 0, cache );
	sync();
	clk.lap( b2GpuStage_prepareConstraints );

	for ( int subStep = 0; subStep < P.subStepCount; ++subStep )
	{
		runStage( P, OP_INTEGRATE_VELOCITIES, 0 );
		sync();
		clk.lap( b2GpuStage_integrateVelocities );

		if ( hasOverflow )
		{
			runStage( P, OP_OVERFLOW_WARM, 0 );
			sync();
		}
		for ( int c = 0; c < colorCount; ++c )
		{
			runStage( P, OP_WARM, c, cache );
			sync();
		}
		clk.lap( b2GpuStage_warmStart );

		if ( hasOverflow )
		{
			runStage( P, OP_OVERFLOW_SOLVE, 0 );
			sync();
		}
		for ( int c = 0; c < colorCount; ++c )
		{
			runStage( P, OP_SOLVE, c, cache );
			sync();
		}
		clk.lap( b2GpuStage_solveImpulses );

		runStage( P, OP_INTEGRATE_POSITIONS, 0 );
		sync();
		clk.lap( b2GpuStage_integratePositions );

		if ( hasOverflow )
		{
			runStage( P, OP_OVERFLOW_RELAX, 0 );
			sync();
		}
		for ( int c = 0; c < colorCount; ++c )
		{
			runStage( P, OP_RELAX, c, cache );
			sync();
		}
		clk.lap( b2GpuStage_relaxImpulses );
	}

	// Restitution: the reference skips every SIMD group whose lanes all have restitution 0
	// (src/contact_solver.c:2131, :432); when NO contact of the step has any, all groups skip, so the colour
	// stages and their barriers are skipped as a whole.
	if ( __ldcg( P.g.anyRestitution ) != 0 )
	{
		if ( hasOverflow )
		{
			runStage( P, OP_OVERFLOW_RESTITUTION, 0 );
			sync();
		}
		for ( int c = 0; c < colorCount; ++c )
		{
			runStage( P, OP_RESTITUTION, c );
			sync();
		}
	}
	clk.lap( b2GpuStage_applyRestitution );

	runStage( P, OP_STORE, 0, cache );
	clk.lap( b2GpuStage_storeImpulses );

	if ( clk.lead )
	{
#pragma unroll
		for ( int i = 0; i < b2GpuStage_count; ++i )
		{
			P.stageCycles[i] = (unsigned long long)clk.acc[i];
		}
		P.stageCycles[8] = epoch;
		P.stageCycles[9] = (unsigned long long)( clk.last - begin );
	}
}

// One stage per launch (mode 1): same device code, the stream orders the stages.
__global__ void __launch_bounds__( kBlockThreads, 1 ) b2gStageKernel( const __grid_constant__ StepParams P, int op, int colorIndex )
{
	runStage( P, op, colorIndex );
}

} // namespace b2g
