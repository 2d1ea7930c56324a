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
#include "solver_set.h"

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
	b2GpuIslandSize* islandSizes;
	int islandSizeCapacity;
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
			free( s_slots[i].islandSizes );
			s_slots[i].islandSizes = NULL;
			s_slots[i].islandSizeCapacity = 0;
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

/* The two memory-bound host passes run on the world's own workers: workerCount - 1 tasks go through the world's task
 * callbacks (the same ones b2ParallelFor uses, src/parallel_for.c:108-132) and claim blocks of items from the device
 * library; the calling thread works along and pumps the PCIe transfers (b2GpuSolverPackWork / UnpackWork, pump = 1). */
static void b2SeamPackWorker( void* taskContext )
{
	b2GpuSolverPackWork( taskContext, 0 );
}

static void b2SeamUnpackWorker( void* taskContext )
{
	b2GpuSolverUnpackWork( taskContext, 0 );
}

static void b2SeamWorkOnItems( b2World* world, b2GpuSolver* solver, int itemCount, b2TaskCallback* worker,
							   int ( *work )( b2GpuSolver*, int ), const char* what )
{
	void* handles[B2_MAX_WORKERS];
	int helperCount = world->workerCount - 1;
	int useful = itemCount / 2048; /* a helper that wakes up for less than that only costs */
	helperCount = helperCount < useful ? helperCount : useful;
	int enqueued = 0;
	for ( int i = 0; i < helperCount && world->taskCount < B2_MAX_TASKS; ++i )
	{
		handles[enqueued++] = world->enqueueTaskFcn( worker, solver, world->userTaskContext );
		world->taskCount += 1;
	}
	int rc = work( solver, 1 );
	for ( int i = 0; i < enqueued; ++i )
	{
		if ( handles[i] != NULL )
		{
			world->finishTaskFcn( handles[i], world->userTaskContext );
		}
	}
	if ( rc != 0 )
	{
		b2SeamFatal( what );
	}
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
	b2SolverSet* awakeSet = world->solverSets.data + b2_awakeSet;
	if ( slot->islandSizeCapacity < awakeSet->islandSims.count )
	{
		free( slot->islandSizes );
		slot->islandSizeCapacity = awakeSet->islandSims.count + awakeSet->islandSims.count / 2 + 64;
		slot->islandSizes = malloc( (size_t)slot->islandSizeCapacity * sizeof( b2GpuIslandSize ) );
	}
	b2GpuSeam_FillIslands( world, desc, slot->islandLabels, slot->islandSizes, true );

	b2GpuStepResult* result = &slot->lastResult;
	memset( result, 0, sizeof( *result ) );
	b2TaskContext* taskContext0 = world->taskContexts.data + 0;
	result->hitEventBits = taskContext0->hitEventBitSet.bits;
	result->jointEventBits = taskContext0->jointStateBitSet.bits;

	if ( b2GpuSolverBeginStep( slot->solver, desc, result ) != 0 )
	{
		b2SeamFatal( "b2GpuSolverBeginStep failed" );
	}
	int itemCount = b2GpuSolverGetPackItemCount( slot->solver );
	b2SeamWorkOnItems( world, slot->solver, itemCount, b2SeamPackWorker, b2GpuSolverPackWork, "packing / upload failed" );
	if ( b2GpuSolverSubmit( slot->solver ) != 0 )
	{
		b2SeamFatal( "device solve failed" );
	}
	// the helpers are woken while the kernels run and unpack behind the download
	b2SeamWorkOnItems( world, slot->solver, itemCount, b2SeamUnpackWorker, b2GpuSolverUnpackWork, "device solve / download failed" );
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
