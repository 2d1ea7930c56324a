/*
 * b2_gpu_seam.c -- C17 host seam between the reference's b2Solve and the B200 solver's C-ABI.
 *
 * Compiled together with the reference's own (unmodified) translation units; it includes the reference's
 * internal headers from where they lie (nothing of the reference is copied into this repository).
 * See b2_gpu_seam.h and INTEGRATION.md.  The descriptor / layout checks live in b2_gpu_seam_desc.c.
 */
#include "b2_gpu_seam.h"

/* reference internals (include path: <reference>/src and <reference>/include) */
#include "bitset.h"
#include "core.h"
#include "id_pool.h"
#include "parallel_for.h"
#include "physics_world.h"
#include "solver.h"

#include "box2d/base.h"
#include "box2d/constants.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- per world-slot device solvers ---------------------------------------------------------------------- */

typedef struct b2SeamSlot
{
	b2GpuSolver* solver;
	int* islandLabels;
	int islandLabelCapacity;
	b2GpuSeamTotals totals;
	b2GpuStepResult lastResult;
	b2GpuStepDesc lastDesc;
} b2SeamSlot;

static b2SeamSlot s_slots[B2_MAX_WORLDS];
static int s_mode = -1;

static int b2SeamEnvInt( const char* name, int fallback )
{
	const char* v = getenv( name );
	return v != NULL && v[0] != 0 ? atoi( v ) : fallback;
}

static void b2SeamFatal( const char* what )
{
	fprintf( stderr, "box2d_b200: %s: %s\n", what, b2GpuGetLastError() );
	fflush( stderr );
	abort();
}

static b2SeamSlot* b2SeamGetSlot( b2World* world )
{
	b2SeamSlot* slot = s_slots + world->worldId;
	if ( slot->solver == NULL )
	{
		int device = b2SeamEnvInt( "B2GPU_DEVICE", 0 );
		slot->solver = b2GpuSolverCreate( device );
		if ( slot->solver == NULL )
		{
			// No CPU fallback by design (north_star): fail loudly.
			b2SeamFatal( "cannot create the device solver" );
		}
		int mode = s_mode >= 0 ? s_mode : b2SeamEnvInt( "B2GPU_MODE", 0 );
		b2GpuSolverSetMode( slot->solver, mode );
	}
	return slot;
}

void b2GpuSeam_SetMode( int mode )
{
	s_mode = mode;
	for ( int i = 0; i < B2_MAX_WORLDS; ++i )
	{
		if ( s_slots[i].solver != NULL )
		{
			b2GpuSolverSetMode( s_slots[i].solver, mode );
		}
	}
}

void b2GpuSeam_Shutdown( void )
{
	for ( int i = 0; i < B2_MAX_WORLDS; ++i )
	{
		if ( s_slots[i].solver != NULL )
		{
			b2GpuSolverDestroy( s_slots[i].solver );
			s_slots[i].solver = NULL;
			free( s_slots[i].islandLabels );
			s_slots[i].islandLabels = NULL;
			s_slots[i].islandLabelCapacity = 0;
		}
	}
}

const b2GpuStepResult* b2GpuSeam_GetLastResult( int worldIndex )
{
	return 0 <= worldIndex && worldIndex < B2_MAX_WORLDS ? &s_slots[worldIndex].lastResult : NULL;
}

void b2GpuSeam_GetTotals( int worldIndex, b2GpuSeamTotals* totals, int reset )
{
	if ( 0 <= worldIndex && worldIndex < B2_MAX_WORLDS )
	{
		*totals = s_slots[worldIndex].totals;
		if ( reset )
		{
			memset( &s_slots[worldIndex].totals, 0, sizeof( b2GpuSeamTotals ) );
		}
	}
}

const b2GpuStepDesc* b2GpuSeam_GetLastDesc( int worldIndex )
{
	return 0 <= worldIndex && worldIndex < B2_MAX_WORLDS ? &s_slots[worldIndex].lastDesc : NULL;
}

void b2GpuSeam_InstallPinnedAllocator( void )
{
	b2SetAllocator( b2GpuHostAlloc, b2GpuHostFree );
}

typedef struct b2SeamPackRange
{
	b2GpuSolver* solver;
	int offset;
} b2SeamPackRange;

static void b2SeamPackTask( int startIndex, int endIndex, int workerIndex, void* taskContext )
{
	(void)workerIndex;
	b2SeamPackRange* range = taskContext;
	b2GpuSolverPackRange( range->solver, range->offset + startIndex, range->offset + endIndex );
}

static void b2SeamUnpackTask( int startIndex, int endIndex, int workerIndex, void* taskContext )
{
	(void)workerIndex;
	b2GpuSolverUnpackRange( taskContext, startIndex, endIndex );
}

/* ---- the seam ------------------------------------------------------------------------------------------- */

void b2GpuSeam_SolveConstraints( b2World* world, b2StepContext* context )
{
	b2SeamSlot* slot = b2SeamGetSlot( world );

	// Same per-worker reset the reference does before fanning out (src/solver.c:1563-1570).
	int jointIdCapacity = b2GetIdCapacity( &world->jointIdPool );
	int contactIdCapacity = b2GetIdCapacity( &world->contactIdPool );
	for ( int i = 0; i < world->workerCount; ++i )
	{
		b2TaskContext* taskContext = world->taskContexts.data + i;
		b2SetBitCountAndClear( &taskContext->jointStateBitSet, jointIdCapacity );
		b2SetBitCountAndClear( &taskContext->hitEventBitSet, contactIdCapacity );
		taskContext->hasHitEvents = false;
	}

	// Joint preparation chases world->bodies -> solverSets -> bodySims (e.g. src/revolute_joint.c:221-266):
	// host-only structures, so it stays on the host.
	b2GpuSeam_PrepareJoints( world, context );

	b2GpuStepDesc* desc = &slot->lastDesc;
	b2GpuSeam_BuildDesc( world, context, desc );

	// island hint: lets the device solve islands independently in shared memory (no grid barriers)
	if ( slot->islandLabelCapacity < desc->awakeBodyCount )
	{
		free( slot->islandLabels );
		slot->islandLabelCapacity = desc->awakeBodyCount + desc->awakeBodyCount / 2 + 64;
		slot->islandLabels = malloc( (size_t)slot->islandLabelCapacity * sizeof( int ) );
	}
	b2GpuSeam_FillIslands( world, desc, slot->islandLabels, true );

	b2GpuStepResult* result = &slot->lastResult;
	memset( result, 0, sizeof( *result ) );
	b2TaskContext* taskContext0 = world->taskContexts.data + 0;
	result->hitEventBits = taskContext0->hitEventBitSet.bits;
	result->jointEventBits = taskContext0->jointStateBitSet.bits;

	// The two memory-bound host passes (wire packing, impulse write-back) run on the world's own workers.
	if ( b2GpuSolverBeginStep( slot->solver, desc, result ) != 0 )
	{
		b2SeamFatal( "b2GpuSolverBeginStep failed" );
	}
	// pack in a few chunks and start each chunk's upload as soon as it is packed: PCIe overlaps the packing
	{
		int itemCount = b2GpuSolverGetPackItemCount( slot->solver );
		int chunkCount = itemCount > 400000 ? 4 : 1; // measured: below that the extra parallel-for wake-ups cost more than the overlap wins
		int done = 0;
		for ( int chunk = 0; chunk < chunkCount; ++chunk )
		{
			int end = chunk + 1 == chunkCount ? itemCount : (int)( (long long)itemCount * ( chunk + 1 ) / chunkCount );
			b2SeamPackRange range = { slot->solver, done };
			b2ParallelFor( world, b2SeamPackTask, end - done, 512, &range );
			done = end;
			if ( chunk + 1 < chunkCount )
			{
				b2GpuSolverFlushPacked( slot->solver, done );
			}
		}
	}
	if ( b2GpuSolverSubmit( slot->solver ) != 0 || b2GpuSolverWait( slot->solver ) != 0 )
	{
		b2SeamFatal( "device solve failed" );
	}
	b2ParallelFor( world, b2SeamUnpackTask, b2GpuSolverGetUnpackItemCount( slot->solver ), 512, slot->solver );
	if ( b2GpuSolverEndStep( slot->solver, result ) != 0 )
	{
		b2SeamFatal( "b2GpuSolverEndStep failed" );
	}

	slot->totals.steps += 1;
	slot->totals.kernelMs += result->kernelMs;
	slot->totals.abiMs += result->totalMs;
	slot->totals.h2dBytes += (double)result->h2dBytes;
	slot->totals.d2hBytes += (double)result->d2hBytes;
	slot->totals.launches += result->kernelLaunches;
	slot->totals.gridBarriers += result->gridBarriers;
	for ( int i = 0; i < b2GpuStage_count; ++i )
	{
		slot->totals.stageMs[i] += result->stageMs[i];
	}

	// The event consumers at src/solver.c:1648-1820 read worker 0's sets after OR-ing the others in.
	taskContext0->hasHitEvents = result->hasHitEvents != 0;

	// Keep filling the reference's stage profile: device time per stage group.
	b2Profile* profile = &world->profile;
	profile->prepareConstraints += result->stageMs[b2GpuStage_prepareConstraints];
	profile->integrateVelocities += result->stageMs[b2GpuStage_integrateVelocities];
	profile->warmStart += result->stageMs[b2GpuStage_warmStart];
	profile->solveImpulses += result->stageMs[b2GpuStage_solveImpulses];
	profile->integratePositions += result->stageMs[b2GpuStage_integratePositions];
	profile->relaxImpulses += result->stageMs[b2GpuStage_relaxImpulses];
	profile->applyRestitution += result->stageMs[b2GpuStage_applyRestitution];
	profile->storeImpulses += result->stageMs[b2GpuStage_storeImpulses];
}
