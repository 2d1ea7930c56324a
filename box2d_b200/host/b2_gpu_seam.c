/*
 * b2_gpu_seam.c -- C17 host seam between the reference's b2Solve and the B200 solver's C-ABI.
 *
 * Compiled together with the reference's own (unmodified) translation units; it includes the reference's
 * internal headers from where they lie (nothing of the reference is copied into this repository).
 * See b2_gpu_seam.h and INTEGRATION.md.  The descriptor / layout checks live in b2_gpu_seam_desc.c.
 */
#include "b2_gpu_seam.h"

/* reference internals (include path: <reference>/src and <reference>/include) */
#include "atomic.h"
#include "bitset.h"
#include "body.h"
#include "contact.h"
#include "constraint_graph.h"
#include "core.h"
#include "id_pool.h"
#include "island.h"
#include "joint.h"
#include "parallel_for.h"
#include "physics_world.h"
#include "solver.h"
#include "solver_set.h"
#include "world_snapshot.h"

#include "box2d/base.h"
#include "box2d/box2d.h"
#include "box2d/constants.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- per world-slot device solvers ---------------------------------------------------------------------- */

struct b2SeamGroup;

typedef struct b2SeamSlot
{
	struct b2SeamGroup* group; /* the world is stepped as a member of a group (b2GpuSeam_CreateGroup) */
	int groupMember;
	b2GpuSolver* solver;
	int* islandLabels;
	int islandLabelCapacity;
	b2GpuIslandSize* islandSizes;
	int islandSizeCapacity;
	b2GpuSeamTotals totals;
	b2GpuStepResult lastResult;
	b2GpuStepDesc lastDesc;
	uint16_t generation;  /* of the world the solver was created for (b2World::generation, bumped by b2DestroyWorld) */
	/* what the narrow phase recycled this step (b2GpuSeam_BeginCollide / b2GpuSeam_ContactRecycled) */
	b2GpuRecycledContact* recycled;
	int recycledCapacity;
	uint32_t recycledStamp;
	int collideStart[B2_GRAPH_COLOR_COUNT], collideCount[B2_GRAPH_COLOR_COUNT];
	bool colorTouched[B2_GRAPH_COLOR_COUNT]; /* a contact was added to / removed from the colour since the narrow phase began */
	int collideTotal;						 /* contacts of all colours (b2Collide's flat array goes on with the non-touching ones) */
	bool impulsesPending;					 /* b2GpuSolverDeferredPending when the narrow phase began */
	int collideCursor[B2_MAX_WORKERS][16];	 /* per narrow-phase worker (a cache line each): the colour of its previous contact */
	int collidesSinceSolve;
	bool islandsCaptured; /* b2GpuSeam_BeforeIslandSplit filled the hint of the step in flight */
	int capturedIslandCount;
	long long flushes; /* deferred impulses: times every pending manifold had to be materialized at once */
} b2SeamSlot;

static b2SeamSlot s_slots[B2_MAX_WORLDS];
static int s_mode = -1;

/* A group of worlds that meet at the seam (see b2_gpu_seam.h) */
typedef struct b2SeamGroup
{
	b2GpuSolver* solver; /* one solver for the whole group */
	int count;
	int* worldIndices;
	b2GpuStepDesc* descs;
	b2GpuStepResult* results;
	pthread_mutex_t mutex;
	pthread_cond_t arrivedAll;
	int arrived;
	unsigned round;
	int failed;
	int inUse;
} b2SeamGroup;

#define B2_SEAM_MAX_GROUPS 16
static b2SeamGroup s_groups[B2_SEAM_MAX_GROUPS];

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

static void b2SeamReleaseSlot( b2SeamSlot* slot )
{
	if ( slot->group != NULL )
	{
		return; /* a member of a live group: released with the group */
	}
	if ( slot->solver != NULL )
	{
		b2GpuSolverDestroy( slot->solver );
	}
	free( slot->islandLabels );
	free( slot->islandSizes );
	free( slot->recycled );
	memset( slot, 0, sizeof( *slot ) );
}

static b2SeamSlot* b2SeamGetSlot( b2World* world )
{
	b2SeamSlot* slot = s_slots + world->worldId;
	if ( slot->group != NULL )
	{
		return slot; /* the group owns the device solver */
	}
	if ( slot->solver != NULL && slot->generation != world->generation )
	{
		// the slot's previous world was destroyed (src/physics_world.c, b2DestroyWorld bumps the generation): its device
		// buffers, resident copies and planner state must not leak into the world that reuses the id
		b2SeamReleaseSlot( slot );
	}
	if ( slot->solver == NULL )
	{
		slot->generation = world->generation;
		int device = b2SeamEnvInt( "B2GPU_DEVICE", 0 );
		slot->solver = b2GpuSolverCreate( device );
		if ( slot->solver == NULL )
		{
			// No CPU fallback by design (north_star): fail loudly.
			b2SeamFatal( "cannot create the device solver" );
		}
		int mode = s_mode >= 0 ? s_mode : b2SeamEnvInt( "B2GPU_MODE", 0 );
		b2GpuSolverSetMode( slot->solver, mode );
		// the impulse records stay on the library's side until a manifold is read ("deferred impulses" below); B2GPU_DEFER=0:
		// every step writes them into the manifolds like b2StoreImpulsesTask does
		if ( b2GpuSolverSetDeferredImpulses( slot->solver, b2SeamEnvInt( "B2GPU_DEFER", 1 ) ) != 0 )
		{
			b2SeamFatal( "b2GpuSolverSetDeferredImpulses" );
		}
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
		b2SeamReleaseSlot( s_slots + i );
	}
}

void b2GpuSeam_ReleaseWorld( int worldIndex )
{
	if ( 0 <= worldIndex && worldIndex < B2_MAX_WORLDS )
	{
		b2SeamReleaseSlot( s_slots + worldIndex );
	}
}

int b2GpuSeam_HasSolver( int worldIndex )
{
	return 0 <= worldIndex && worldIndex < B2_MAX_WORLDS && s_slots[worldIndex].solver != NULL ? 1 + (int)s_slots[worldIndex].generation : 0;
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

int b2GpuSeam_GetResidentStats( int worldIndex, int* fullContacts, int* dirtyBodies, int* vouchedContacts )
{
	if ( worldIndex < 0 || worldIndex >= B2_MAX_WORLDS || s_slots[worldIndex].solver == NULL )
	{
		return 0;
	}
	return b2GpuSolverGetResidentStats( s_slots[worldIndex].solver, fullContacts, dirtyBodies, vouchedContacts, NULL );
}

const b2GpuStepDesc* b2GpuSeam_GetLastDesc( int worldIndex )
{
	return 0 <= worldIndex && worldIndex < B2_MAX_WORLDS ? &s_slots[worldIndex].lastDesc : NULL;
}

void b2GpuSeam_InstallPinnedAllocator( void )
{
	b2SetAllocator( b2GpuHostAlloc, b2GpuHostFree );
}

/* ---- the team: the world's workers are woken ONCE per step ------------------------------------------------------
 * The host side of a step has three parallel passes -- island labels, packing, unpacking -- with a few microseconds of
 * serial work between them (descriptor + layout, launches).  Waking the workers costs more than that (a semaphore post per
 * task, src/scheduler.c:158-176, and tens of microseconds until a parked thread runs), so workerCount - 1 helper tasks go
 * through the world's task callbacks once (the same ones b2ParallelFor uses, src/parallel_for.c:108-132) and walk the
 * passes together with the calling thread, spinning across the short gaps like the reference's own solver workers spin
 * for their next stage (src/solver.c:921-1008).  Work is claimed in blocks, so a helper that starts late -- or never -- only
 * means fewer hands.  The calling thread pumps the PCIe transfers (b2GpuSolverPackWork / UnpackWork, pump = 1). */
enum
{
	b2_seamLabels = 0,	   /* claiming blocks of island labels */
	b2_seamPackOpen = 1,   /* the step has begun: b2GpuSolverPackWork */
	b2_seamPackClosed = 2, /* everything is packed and on its way; the calling thread submits */
	b2_seamUnpackOpen = 3, /* submitted: b2GpuSolverUnpackWork */
	b2_seamDone = 4,
};

typedef struct b2SeamTeam
{
	b2World* world;
	b2GpuSolver* solver;
	const b2BodySim* sims;
	int* labels;
	int labelCount, labelBlocks;
	b2AtomicInt labelNext, labelDone;
	// host joint prepare (b2PrepareJoint, src/joint.c:1406): the colours' joint arrays laid end to end, overflow colour last
	b2StepContext* context;
	b2JointSim* jointArrays[B2_GRAPH_COLOR_COUNT];
	int jointStarts[B2_GRAPH_COLOR_COUNT + 1];
	int jointArrayCount, jointBlocks;
	b2AtomicInt jointNext, jointDone;
	b2AtomicInt phase;
	b2AtomicInt inPack, inUnpack;
	b2AtomicInt failed;
} b2SeamTeam;

#define B2_SEAM_LABEL_BLOCK 512
#define B2_SEAM_JOINT_BLOCK 32

static inline void b2SeamRelax( int* spins )
{
#if defined( B2_CPU_X86_X64 ) || defined( __x86_64__ )
	__builtin_ia32_pause();
#endif
	*spins += 1;
	if ( *spins > 4096 )
	{
		/* somebody this thread waits for is not running: give it the core */
		b2Yield();
		*spins = 0;
	}
}

static void b2SeamTeamLabels( b2SeamTeam* team )
{
	const b2Body* bodies = team->world->bodies.data;
	const b2Island* islands = team->world->islands.data;
	for ( ;; )
	{
		int block = b2AtomicFetchAddInt( &team->labelNext, 1 );
		if ( block >= team->labelBlocks )
		{
			break;
		}
		int begin = block * B2_SEAM_LABEL_BLOCK;
		int end = begin + B2_SEAM_LABEL_BLOCK < team->labelCount ? begin + B2_SEAM_LABEL_BLOCK : team->labelCount;
		// label = index of the body's island among the awake islands (b2Island::localIndex, src/island.h:49-74)
		// (two dependent look-ups per body, the first one all over world->bodies: the body records of the bodies a few
		// places ahead are asked for early)
		const int ahead = 12;
		for ( int i = begin; i < end; ++i )
		{
			if ( i + ahead < end )
			{
				__builtin_prefetch( &bodies[team->sims[i + ahead].bodyId].islandId, 0, 1 );
			}
			int islandId = bodies[team->sims[i].bodyId].islandId;
			team->labels[i] = islandId == B2_NULL_INDEX ? -1 : islands[islandId].localIndex;
		}
		b2AtomicFetchAddInt( &team->labelDone, 1 );
	}
}

// Joint preparation chases world->bodies -> solverSets -> bodySims (e.g. src/revolute_joint.c:221-266): host-only structures,
// so it stays on the host -- the reference's own b2PrepareJoint, joint by joint, in blocks the
// team claims (the reference runs it as its first solver stage, src/solver.c:1060-1077).
static void b2SeamTeamJoints( b2SeamTeam* team )
{
	for ( ;; )
	{
		int block = b2AtomicFetchAddInt( &team->jointNext, 1 );
		if ( block >= team->jointBlocks )
		{
			break;
		}
		int begin = block * B2_SEAM_JOINT_BLOCK;
		int total = team->jointStarts[team->jointArrayCount];
		int end = begin + B2_SEAM_JOINT_BLOCK < total ? begin + B2_SEAM_JOINT_BLOCK : total;
		int array = 0;
		for ( int i = begin; i < end; ++i )
		{
			while ( team->jointStarts[array + 1] <= i )
			{
				array += 1;
			}
			b2PrepareJoint( team->jointArrays[array] + ( i - team->jointStarts[array] ), team->context );
		}
		b2AtomicFetchAddInt( &team->jointDone, 1 );
	}
}

static void b2SeamTeamHelper( void* taskContext )
{
	b2SeamTeam* team = taskContext;
	int spins = 0;
	b2SeamTeamJoints( team );
	b2SeamTeamLabels( team );
	while ( b2AtomicLoadInt( &team->phase ) < b2_seamPackOpen )
	{
		b2SeamRelax( &spins );
	}
	b2AtomicFetchAddInt( &team->inPack, 1 );
	if ( b2AtomicLoadInt( &team->phase ) == b2_seamPackOpen && b2GpuSolverPackWork( team->solver, 0 ) != 0 )
	{
		b2AtomicStoreInt( &team->failed, 1 );
	}
	b2AtomicFetchAddInt( &team->inPack, -1 );
	while ( b2AtomicLoadInt( &team->phase ) < b2_seamUnpackOpen )
	{
		b2SeamRelax( &spins );
	}
	b2AtomicFetchAddInt( &team->inUnpack, 1 );
	if ( b2AtomicLoadInt( &team->phase ) == b2_seamUnpackOpen && b2GpuSolverUnpackWork( team->solver, 0 ) != 0 )
	{
		b2AtomicStoreInt( &team->failed, 1 );
	}
	b2AtomicFetchAddInt( &team->inUnpack, -1 );
}

/* ---- recycled manifolds: the narrow phase tells, the pack pass need not rediscover ------------------------------------
 * Called by the generated physics_world.c (tools/patch_collide.py).  The reference's contact recycling
 * (src/physics_world.c:508-560) leaves a b2ContactSim exactly as the previous step's solver left it, except for the two
 * separations it recomputes and the body indices / masses it refreshes -- which is what b2GpuStepDesc::recycled says. */
void b2GpuSeam_BeginCollide( b2World* world, b2StepContext* context, int contactCount )
{
	(void)context;
	b2SeamSlot* slot = b2SeamGetSlot( world );
	if ( slot->recycledCapacity < contactCount )
	{
		free( slot->recycled );
		slot->recycledCapacity = contactCount + contactCount / 2 + 256;
		slot->recycled = calloc( (size_t)slot->recycledCapacity, sizeof( b2GpuRecycledContact ) );
		slot->recycledStamp = 0;
	}
	slot->recycledStamp += 1;
	if ( slot->recycledStamp == 0 )
	{
		// wrapped: no entry of the past may look current
		memset( slot->recycled, 0, (size_t)slot->recycledCapacity * sizeof( b2GpuRecycledContact ) );
		slot->recycledStamp = 1;
	}
	// b2Collide lays the colours' arrays end to end, in colour order (src/physics_world.c:665-677): contact j of colour i is
	// contactIndex collideStart[i] + j of b2CollideTask
	int start = 0;
	for ( int i = 0; i < B2_GRAPH_COLOR_COUNT; ++i )
	{
		slot->collideStart[i] = start;
		slot->collideCount[i] = world->constraintGraph.colors[i].contactSims.count;
		slot->colorTouched[i] = false;
		start += slot->collideCount[i];
	}
	slot->collideTotal = start;
	memset( slot->collideCursor, 0, sizeof( slot->collideCursor ) );
	slot->impulsesPending = slot->solver != NULL && b2GpuSolverDeferredPending( slot->solver ) != 0;
	slot->collidesSinceSolve += 1;
	// the narrow phase's workers materialize what they re-evaluate (b2GpuSeam_ContactReevaluated): the tail of the previous
	// step's download is waited for here, once
	if ( slot->solver != NULL && b2GpuSolverDeferredSync( slot->solver ) != 0 )
	{
		b2SeamFatal( "download of the deferred impulses failed" );
	}
}

void b2GpuSeam_ContactRecycled( b2World* world, int contactIndex, const b2ContactSim* contactSim )
{
	b2SeamSlot* slot = s_slots + world->worldId;
	b2GpuRecycledContact* entry = slot->recycled + contactIndex;
	entry->contactId = contactSim->contactId;
	entry->separation[0] = contactSim->manifold.points[0].separation;
	entry->separation[1] = contactSim->manifold.points[1].separation;
	entry->indexA = contactSim->bodySimIndexA;
	entry->indexB = contactSim->bodySimIndexB;
	entry->stamp = slot->recycledStamp;
}

/* ---- deferred impulses: a manifold receives the solver's impulses when somebody is going to read them -----------------
 * The reference stores the impulses of every contact at the end of every step (b2StoreImpulsesTask,
 * src/contact_solver.c:2293-2320) -- and then reads them back in few places: b2UpdateContact when it re-evaluates a manifold
 * (src/contact.c:523: the old points' impulses are matched by feature id), the hit events of the step
 * (src/solver.c:1759-1766), the contact-data API (src/body.c:482, src/shape.c:1765, src/contact.c:83), the debug draw
 * (src/physics_world.c:1325), the snapshot / state hash (src/world_snapshot.c) and, implicitly, island sleep, which moves
 * the b2ContactSims out of the solver's sight (src/solver_set.c:318).  A RECYCLED manifold (src/physics_world.c:508-560)
 * is not read at all, and the device warm-starts it from its own previous output.  So the seam leaves the records in the
 * library's page-locked output arena (b2GpuSolverSetDeferredImpulses) and calls b2GpuSolverMaterializeContacts from
 * exactly those places: the generated physics_world.c before b2UpdateContact, and the functions below, which the build
 * puts in front of the reference's (tools/buildlib.py renames the reference's definitions to b2Ref_*; the reference's
 * sources are not touched).  The many_pyramids step loses its 3 MB download and the unpack pass over 58 000 manifolds from
 * the critical path. */
void b2GpuSeam_ContactReevaluated( b2World* world, int workerIndex, int contactIndex, b2ContactSim* contactSim )
{
	b2SeamSlot* slot = s_slots + world->worldId;
	if ( slot->impulsesPending == false || contactIndex >= slot->collideTotal )
	{
		// (what follows the colours' arrays in b2Collide's flat array are the awake set's non-touching contacts: never solved)
		return;
	}
	// contactIndex -> (graph colour, place): b2Collide laid the colours' arrays end to end (b2GpuSeam_BeginCollide) and a worker
	// walks a range of that array front to back, so the colour of its previous contact is where to start looking
	int* cursor = slot->collideCursor[workerIndex & ( B2_MAX_WORKERS - 1 )];
	int i = *cursor;
	if ( contactIndex < slot->collideStart[i] )
	{
		i = 0;
	}
	while ( i < B2_GRAPH_COLOR_COUNT - 1 && contactIndex >= slot->collideStart[i] + slot->collideCount[i] )
	{
		i += 1;
	}
	if ( i != *cursor )
	{
		*cursor = i;
	}
	int j = contactIndex - slot->collideStart[i];
	if ( 0 <= j && j < slot->collideCount[i] && b2GpuSolverMaterializeContacts( slot->solver, i, j, contactSim, 1, NULL ) < 0 )
	{
		b2SeamFatal( "b2GpuSolverMaterializeContacts" );
	}
}

/* A flush in the middle of a step (an island goes to sleep, hit events): the world's workers share it.  The colours' contact
 * arrays, then their joint arrays, laid end to end; b2ParallelFor deals out ranges of that. */
typedef struct b2SeamFlushJob
{
	b2World* world;
	b2GpuSolver* solver;
	b2GpuStepResult* result;
	int starts[2 * B2_GRAPH_COLOR_COUNT + 1];
	int failed;
} b2SeamFlushJob;

static void b2SeamFlushTask( int startIndex, int endIndex, int workerIndex, void* context )
{
	(void)workerIndex;
	b2SeamFlushJob* job = context;
	int segment = 0;
	while ( job->starts[segment + 1] <= startIndex )
	{
		segment += 1;
	}
	for ( int at = startIndex; at < endIndex; )
	{
		while ( job->starts[segment + 1] <= at )
		{
			segment += 1;
		}
		int upto = endIndex < job->starts[segment + 1] ? endIndex : job->starts[segment + 1];
		int first = at - job->starts[segment], count = upto - at;
		int rc;
		if ( segment < B2_GRAPH_COLOR_COUNT )
		{
			b2GraphColor* color = job->world->constraintGraph.colors + segment;
			rc = b2GpuSolverMaterializeContacts( job->solver, segment, first, color->contactSims.data + first, count, job->result );
		}
		else
		{
			b2GraphColor* color = job->world->constraintGraph.colors + ( segment - B2_GRAPH_COLOR_COUNT );
			rc = b2GpuSolverMaterializeJoints( job->solver, segment - B2_GRAPH_COLOR_COUNT, first, color->jointSims.data + first, count );
		}
		if ( rc < 0 )
		{
			job->failed = 1;
		}
		at = upto;
	}
}

/* every contact of the constraint graph that still waits for its impulses receives them */
static void b2SeamFlushImpulses( b2World* world, b2GpuStepResult* result )
{
	if ( world == NULL || world->worldId < 0 || world->worldId >= B2_MAX_WORLDS )
	{
		return;
	}
	b2SeamSlot* slot = s_slots + world->worldId;
	if ( slot->solver == NULL || slot->generation != world->generation || b2GpuSolverDeferredPending( slot->solver ) == 0 )
	{
		return;
	}
	if ( world->locked && world->workerCount > 1 )
	{
		// inside b2World_Step, on the stepping thread, between the reference's own parallel passes
		b2SeamFlushJob job = { world, slot->solver, result, { 0 }, 0 };
		for ( int i = 0; i < B2_GRAPH_COLOR_COUNT; ++i )
		{
			job.starts[i + 1] = job.starts[i] + world->constraintGraph.colors[i].contactSims.count;
		}
		for ( int i = 0; i < B2_GRAPH_COLOR_COUNT; ++i )
		{
			job.starts[B2_GRAPH_COLOR_COUNT + i + 1] = job.starts[B2_GRAPH_COLOR_COUNT + i] + world->constraintGraph.colors[i].jointSims.count;
		}
		int total = job.starts[2 * B2_GRAPH_COLOR_COUNT];
		if ( total >= 4096 && b2GpuSolverDeferredSync( slot->solver ) == 0 )
		{
			b2ParallelFor( world, b2SeamFlushTask, total, 512, &job );
			if ( job.failed )
			{
				b2SeamFatal( "b2GpuSolverMaterializeContacts" );
			}
			b2GpuSolverDeferredDone( slot->solver );
			slot->flushes += 1;
			return;
		}
	}
	for ( int i = 0; i < B2_GRAPH_COLOR_COUNT; ++i )
	{
		b2GraphColor* color = world->constraintGraph.colors + i;
		if ( color->contactSims.count > 0 &&
			 b2GpuSolverMaterializeContacts( slot->solver, i, 0, color->contactSims.data, color->contactSims.count, result ) < 0 )
		{
			b2SeamFatal( "b2GpuSolverMaterializeContacts" );
		}
		if ( color->jointSims.count > 0 && b2GpuSolverMaterializeJoints( slot->solver, i, 0, color->jointSims.data, color->jointSims.count ) < 0 )
		{
			b2SeamFatal( "b2GpuSolverMaterializeJoints" );
		}
	}
	b2GpuSolverDeferredDone( slot->solver );
	slot->flushes += 1;
}

void b2GpuSeam_FlushImpulses( int worldIndex )
{
	if ( 0 <= worldIndex && worldIndex < B2_MAX_WORLDS && s_slots[worldIndex].solver != NULL )
	{
		b2SeamFlushImpulses( b2GetWorld( worldIndex ), NULL );
	}
}

int b2GpuSeam_GetDeferredStats( int worldIndex, int* pending, long long* flushes )
{
	if ( worldIndex < 0 || worldIndex >= B2_MAX_WORLDS || s_slots[worldIndex].solver == NULL )
	{
		return 0;
	}
	*pending = b2GpuSolverDeferredPending( s_slots[worldIndex].solver );
	*flushes = s_slots[worldIndex].flushes;
	return 1;
}

/* The reference's readers of manifold impulses outside the narrow phase and the solver, each behind a flush.  b2Ref_* are the
 * reference's own definitions (renamed in the product's copies of its object files). */
int b2Ref_Body_GetContactData( b2BodyId bodyId, b2ContactData* contactData, int capacity );
int b2Ref_Shape_GetContactData( b2ShapeId shapeId, b2ContactData* contactData, int capacity );
b2ContactData b2Ref_Contact_GetData( b2ContactId contactId );
void b2Ref_World_Draw( b2WorldId worldId, b2DebugDraw* draw );
uint64_t b2Ref_World_GetStateHash( b2WorldId worldId );
int b2Ref_World_Snapshot( b2WorldId worldId, uint8_t* image, int capacity );
bool b2Ref_World_Restore( b2WorldId worldId, const uint8_t* image, int size );
void b2Ref_SerializeWorld( b2World* world, b2RecBuffer* buf );
uint64_t b2Ref_HashWorldStateDeep( b2World* world );
void b2Ref_TrySleepIsland( b2World* world, int islandId );

static void b2SeamFlushIndex( int worldIndex )
{
	if ( 0 <= worldIndex && worldIndex < B2_MAX_WORLDS && s_slots[worldIndex].solver != NULL )
	{
		b2World* world = b2GetWorld( worldIndex );
		if ( world->inUse && world->locked == false )
		{
			b2SeamFlushImpulses( world, NULL );
		}
	}
}

int b2Body_GetContactData( b2BodyId bodyId, b2ContactData* contactData, int capacity )
{
	b2SeamFlushIndex( bodyId.world0 );
	return b2Ref_Body_GetContactData( bodyId, contactData, capacity );
}

int b2Shape_GetContactData( b2ShapeId shapeId, b2ContactData* contactData, int capacity )
{
	b2SeamFlushIndex( shapeId.world0 );
	return b2Ref_Shape_GetContactData( shapeId, contactData, capacity );
}

b2ContactData b2Contact_GetData( b2ContactId contactId )
{
	b2SeamFlushIndex( contactId.world0 );
	return b2Ref_Contact_GetData( contactId );
}

void b2World_Draw( b2WorldId worldId, b2DebugDraw* draw )
{
	b2SeamFlushIndex( (int)worldId.index1 - 1 );
	b2Ref_World_Draw( worldId, draw );
}

uint64_t b2World_GetStateHash( b2WorldId worldId )
{
	b2SeamFlushIndex( (int)worldId.index1 - 1 );
	return b2Ref_World_GetStateHash( worldId );
}

int b2World_Snapshot( b2WorldId worldId, uint8_t* image, int capacity )
{
	b2SeamFlushIndex( (int)worldId.index1 - 1 );
	return b2Ref_World_Snapshot( worldId, image, capacity );
}

bool b2World_Restore( b2WorldId worldId, const uint8_t* image, int size )
{
	// the host's contacts are replaced wholesale: nothing of the previous step is owed to them any more, and the narrow
	// phase's word on what it recycles does not hold for the step that follows (the device has never seen this state;
	// the pack pass compares everything by value)
	int worldIndex = (int)worldId.index1 - 1;
	if ( 0 <= worldIndex && worldIndex < B2_MAX_WORLDS && s_slots[worldIndex].solver != NULL )
	{
		b2GpuSolverDeferredDone( s_slots[worldIndex].solver );
		s_slots[worldIndex].collidesSinceSolve = 2;
	}
	return b2Ref_World_Restore( worldId, image, size );
}

void b2SerializeWorld( b2World* world, b2RecBuffer* buf )
{
	if ( world->locked == false )
	{
		b2SeamFlushImpulses( world, NULL );
	}
	b2Ref_SerializeWorld( world, buf );
}

uint64_t b2HashWorldStateDeep( b2World* world )
{
	if ( world->locked == false )
	{
		b2SeamFlushImpulses( world, NULL );
	}
	return b2Ref_HashWorldStateDeep( world );
}

/* The two functions through which a contact enters or leaves a colour's array between the narrow phase and the solver
 * (b2UpdateContacts, src/physics_world.c:802,828; b2DestroyContact, src/contact.c:472; b2WakeSolverSet, src/solver_set.c:114):
 * a colour they have touched is no longer the array the narrow phase's entries were written against. */
void b2Ref_AddContactToGraph( b2World* world, b2ContactSim* contactSim, b2Contact* contact );
void b2Ref_RemoveContactFromGraph( b2World* world, int bodyIdA, int bodyIdB, int colorIndex, int localIndex );

void b2AddContactToGraph( b2World* world, b2ContactSim* contactSim, b2Contact* contact )
{
	b2Ref_AddContactToGraph( world, contactSim, contact );
	if ( 0 <= contact->colorIndex && contact->colorIndex < B2_GRAPH_COLOR_COUNT )
	{
		s_slots[world->worldId].colorTouched[contact->colorIndex] = true;
	}
}

void b2RemoveContactFromGraph( b2World* world, int bodyIdA, int bodyIdB, int colorIndex, int localIndex )
{
	if ( 0 <= colorIndex && colorIndex < B2_GRAPH_COLOR_COUNT )
	{
		b2SeamSlot* slot = s_slots + world->worldId;
		slot->colorTouched[colorIndex] = true;
		if ( slot->solver != NULL && slot->generation == world->generation && b2GpuSolverDeferredPending( slot->solver ) != 0 )
		{
			// deferred impulses are found by place: the colour's LAST contact is about to move into the place of the one that
			// leaves (b2Array_RemoveSwap, src/constraint_graph.c:198-211) -- it receives its impulses first; the record of the
			// one that leaves is void (it stopped touching, or is being destroyed)
			b2GraphColor* color = world->constraintGraph.colors + colorIndex;
			int last = color->contactSims.count - 1;
			if ( last >= 0 && last != localIndex &&
				 b2GpuSolverMaterializeContacts( slot->solver, colorIndex, last, color->contactSims.data + last, 1, NULL ) < 0 )
			{
				b2SeamFatal( "b2GpuSolverMaterializeContacts" );
			}
			b2GpuSolverDeferredForget( slot->solver, colorIndex, localIndex );
		}
	}
	b2Ref_RemoveContactFromGraph( world, bodyIdA, bodyIdB, colorIndex, localIndex );
}

/* The joints' accumulated impulses are deferred like the contacts' (b2GpuSolverMaterializeJoints).  Their readers and writers
 * on the host all fetch the b2JointSim first: the per-type API through b2GetJointSimCheckType (src/joint.c:143; 124 call sites
 * in src/*_joint.c), the generic reaction getters through b2GetJointSim inside src/joint.c.  A joint receives what is owed to
 * it before it is handed out, before it moves in its colour's array, and before it leaves the awake set. */
b2JointSim* b2Ref_GetJointSimCheckType( b2JointId jointId, b2JointType type );
b2Vec2 b2Ref_Joint_GetConstraintForce( b2JointId jointId );
float b2Ref_Joint_GetConstraintTorque( b2JointId jointId );
void b2Ref_RemoveJointFromGraph( b2World* world, int bodyIdA, int bodyIdB, int colorIndex, int localIndex );
void b2Ref_TransferJoint( b2World* world, b2SolverSet* targetSet, b2SolverSet* sourceSet, b2Joint* joint );

static void b2SeamMaterializeJointAt( b2World* world, int colorIndex, int localIndex )
{
	b2SeamSlot* slot = s_slots + world->worldId;
	if ( slot->solver == NULL || slot->generation != world->generation || b2GpuSolverDeferredPending( slot->solver ) == 0 )
	{
		return;
	}
	if ( colorIndex < 0 || colorIndex >= B2_GRAPH_COLOR_COUNT )
	{
		return;
	}
	b2GraphColor* color = world->constraintGraph.colors + colorIndex;
	if ( 0 <= localIndex && localIndex < color->jointSims.count &&
		 b2GpuSolverMaterializeJoints( slot->solver, colorIndex, localIndex, color->jointSims.data + localIndex, 1 ) < 0 )
	{
		b2SeamFatal( "b2GpuSolverMaterializeJoints" );
	}
}

static void b2SeamMaterializeJointId( b2JointId jointId )
{
	int worldIndex = jointId.world0;
	if ( worldIndex < 0 || worldIndex >= B2_MAX_WORLDS || s_slots[worldIndex].solver == NULL )
	{
		return;
	}
	b2World* world = b2GetWorld( worldIndex );
	int id = jointId.index1 - 1;
	if ( world->inUse == false || id < 0 || id >= world->joints.count )
	{
		return;
	}
	const b2Joint* joint = world->joints.data + id;
	if ( joint->setIndex == b2_awakeSet && joint->generation == jointId.generation )
	{
		b2SeamMaterializeJointAt( world, joint->colorIndex, joint->localIndex );
	}
}

b2JointSim* b2GetJointSimCheckType( b2JointId jointId, b2JointType type )
{
	b2SeamMaterializeJointId( jointId );
	return b2Ref_GetJointSimCheckType( jointId, type );
}

b2Vec2 b2Joint_GetConstraintForce( b2JointId jointId )
{
	b2SeamMaterializeJointId( jointId );
	return b2Ref_Joint_GetConstraintForce( jointId );
}

float b2Joint_GetConstraintTorque( b2JointId jointId )
{
	b2SeamMaterializeJointId( jointId );
	return b2Ref_Joint_GetConstraintTorque( jointId );
}

void b2RemoveJointFromGraph( b2World* world, int bodyIdA, int bodyIdB, int colorIndex, int localIndex )
{
	if ( 0 <= colorIndex && colorIndex < B2_GRAPH_COLOR_COUNT )
	{
		// the joint that leaves keeps its b2JointSim elsewhere (another set) or is destroyed; the colour's last joint is about to
		// move into its place (src/constraint_graph.c:312-324): both receive their impulses first, and the place is void
		b2SeamMaterializeJointAt( world, colorIndex, localIndex );
		b2SeamMaterializeJointAt( world, colorIndex, world->constraintGraph.colors[colorIndex].jointSims.count - 1 );
		b2SeamSlot* slot = s_slots + world->worldId;
		if ( slot->solver != NULL && slot->generation == world->generation )
		{
			b2GpuSolverDeferredForgetJoint( slot->solver, colorIndex, localIndex );
		}
	}
	b2Ref_RemoveJointFromGraph( world, bodyIdA, bodyIdB, colorIndex, localIndex );
}

void b2TransferJoint( b2World* world, b2SolverSet* targetSet, b2SolverSet* sourceSet, b2Joint* joint )
{
	// (copies the b2JointSim to the target set BEFORE it takes it out of the graph, src/solver_set.c:573-591)
	if ( sourceSet != targetSet && sourceSet->setIndex == b2_awakeSet )
	{
		b2SeamMaterializeJointAt( world, joint->colorIndex, joint->localIndex );
	}
	b2Ref_TransferJoint( world, targetSet, sourceSet, joint );
}

void b2TrySleepIsland( b2World* world, int islandId )
{
	// (called on the stepping thread: the tail of b2Solve, src/solver.c:2073, and b2Body_SetAwake, src/body.c:1575)
	b2Island* island = world->islands.data + islandId;
	if ( !( island->constraintRemoveCount > 0 && island->bodies.count > 1 ) )
	{
		// (otherwise the island stays awake: src/solver_set.c:160-164)
		b2SeamFlushImpulses( world, NULL );
	}
	b2Ref_TrySleepIsland( world, islandId );
}

/* ---- island capture ahead of a split ---------------------------------------------------------------------- */

static void b2SeamReserveIslands( b2SeamSlot* slot, b2World* world )
{
	b2SolverSet* awakeSet = world->solverSets.data + b2_awakeSet;
	int bodyCount = awakeSet->bodySims.count, islandCount = awakeSet->islandSims.count;
	if ( slot->islandLabelCapacity < bodyCount )
	{
		free( slot->islandLabels );
		slot->islandLabelCapacity = bodyCount + bodyCount / 2 + 64;
		slot->islandLabels = malloc( (size_t)slot->islandLabelCapacity * sizeof( int ) );
	}
	if ( slot->islandSizeCapacity < islandCount )
	{
		free( slot->islandSizes );
		slot->islandSizeCapacity = islandCount + islandCount / 2 + 64;
		slot->islandSizes = malloc( (size_t)slot->islandSizeCapacity * sizeof( b2GpuIslandSize ) );
	}
}

/* Called by the generated solver.c right before b2Solve may enqueue b2SplitIslandTask (src/solver.c:1476-1495).  That task
 * runs concurrently with the constraint solve and rewrites world->islands, the awake set's island sims and the bodies'
 * island ids (src/island.c, b2SplitIsland) -- everything the island hint is read from.  The reference's own solver never
 * looks at islands, the seam does: so when a split is pending the hint is taken HERE, before the task exists.  The labels
 * of the unsplit island are still a valid partition (a superset of what the split will produce). */
void b2GpuSeam_BeforeIslandSplit( b2World* world, b2StepContext* context )
{
	(void)context;
	b2SeamSlot* slot = b2SeamGetSlot( world );
	slot->islandsCaptured = false;
	if ( world->splitIslandId == B2_NULL_INDEX )
	{
		return;
	}
	b2SeamReserveIslands( slot, world );
	b2GpuStepDesc scratch;
	b2GpuSeam_FillIslands( world, &scratch, slot->islandLabels, slot->islandSizes, true );
	slot->capturedIslandCount = scratch.islandCount;
	slot->islandsCaptured = true;
}

/* ---- world-level batched step: groups ------------------------------------------------------------------------ */

int b2GpuSeam_CreateGroup( const int* worldIndices, int worldCount )
{
	if ( worldIndices == NULL || worldCount <= 0 )
	{
		return -1;
	}
	for ( int g = 0; g < B2_SEAM_MAX_GROUPS; ++g )
	{
		b2SeamGroup* group = s_groups + g;
		if ( group->inUse )
		{
			continue;
		}
		for ( int i = 0; i < worldCount; ++i )
		{
			int w = worldIndices[i];
			if ( w < 0 || w >= B2_MAX_WORLDS || s_slots[w].group != NULL )
			{
				return -1;
			}
		}
		memset( group, 0, sizeof( *group ) );
		group->solver = b2GpuSolverCreate( b2SeamEnvInt( "B2GPU_DEVICE", 0 ) );
		if ( group->solver == NULL )
		{
			b2SeamFatal( "cannot create the device solver of a group" );
		}
		group->count = worldCount;
		group->worldIndices = malloc( (size_t)worldCount * sizeof( int ) );
		group->descs = calloc( (size_t)worldCount, sizeof( b2GpuStepDesc ) );
		group->results = calloc( (size_t)worldCount, sizeof( b2GpuStepResult ) );
		pthread_mutex_init( &group->mutex, NULL );
		pthread_cond_init( &group->arrivedAll, NULL );
		group->inUse = 1;
		for ( int i = 0; i < worldCount; ++i )
		{
			int w = worldIndices[i];
			group->worldIndices[i] = w;
			b2SeamReleaseSlot( s_slots + w ); /* a solver the world had on its own */
			s_slots[w].group = group;
			s_slots[w].groupMember = i;
		}
		return g;
	}
	return -1;
}

void b2GpuSeam_DestroyGroup( int g )
{
	if ( g < 0 || g >= B2_SEAM_MAX_GROUPS || s_groups[g].inUse == 0 )
	{
		return;
	}
	b2SeamGroup* group = s_groups + g;
	for ( int i = 0; i < group->count; ++i )
	{
		b2SeamSlot* slot = s_slots + group->worldIndices[i];
		slot->group = NULL;
		b2SeamReleaseSlot( slot );
	}
	b2GpuSolverDestroy( group->solver );
	free( group->worldIndices );
	free( group->descs );
	free( group->results );
	pthread_mutex_destroy( &group->mutex );
	pthread_cond_destroy( &group->arrivedAll );
	memset( group, 0, sizeof( *group ) );
}

/* The seam of a world that is a member of a group: prepare on this world's thread, meet the others, let the last one solve. */
static void b2SeamSolveInGroup( b2World* world, b2StepContext* context, b2SeamSlot* slot )
{
	b2SeamGroup* group = slot->group;
	const int member = slot->groupMember;
	uint64_t seamTicks = b2GetTicks();

	b2GpuSeam_PrepareJoints( world, context );
	b2SeamReserveIslands( slot, world );
	b2GpuStepDesc* desc = group->descs + member;
	b2GpuSeam_BuildDesc( world, context, desc );
	if ( slot->islandsCaptured )
	{
		desc->bodyIsland = slot->islandLabels;
		desc->islandSizes = slot->islandSizes;
		desc->islandCount = slot->capturedIslandCount;
	}
	else
	{
		b2GpuSeam_FillIslands( world, desc, slot->islandLabels, slot->islandSizes, false );
	}
	slot->islandsCaptured = false;
	slot->collidesSinceSolve = 0; /* the batch runs in plain mode: no recycle hints */

	b2GpuStepResult* result = group->results + member;
	memset( result, 0, sizeof( *result ) );
	b2TaskContext* taskContext0 = world->taskContexts.data + 0;
	result->hitEventBits = taskContext0->hitEventBitSet.bits;
	result->jointEventBits = taskContext0->jointStateBitSet.bits;

	pthread_mutex_lock( &group->mutex );
	group->arrived += 1;
	if ( group->arrived == group->count )
	{
		// the last world to arrive solves the whole group; the library packs / unpacks on its own host threads while the
		// other worlds' threads sleep
		int rc = b2GpuSolverStepBatch( group->solver, group->descs, group->count, group->results );
		group->failed = rc != 0;
		group->arrived = 0;
		group->round += 1;
		pthread_cond_broadcast( &group->arrivedAll );
	}
	else
	{
		unsigned round = group->round;
		while ( group->round == round )
		{
			pthread_cond_wait( &group->arrivedAll, &group->mutex );
		}
	}
	int failed = group->failed;
	pthread_mutex_unlock( &group->mutex );
	if ( failed )
	{
		b2SeamFatal( "b2GpuSolverStepBatch failed" );
	}

	slot->lastResult = *result;
	slot->lastDesc = *desc;
	slot->totals.steps += 1;
	slot->totals.kernelMs += result->kernelMs;
	slot->totals.abiMs += result->totalMs;
	slot->totals.h2dBytes += (double)result->h2dBytes;
	slot->totals.d2hBytes += (double)result->d2hBytes;
	slot->totals.launches += result->kernelLaunches;
	slot->totals.seamMs += b2GetMilliseconds( seamTicks );
	taskContext0->hasHitEvents = result->hasHitEvents != 0;
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

/* ---- the seam ------------------------------------------------------------------------------------------- */

void b2GpuSeam_SolveConstraints( b2World* world, b2StepContext* context )
{
	b2SeamSlot* slot = b2SeamGetSlot( world );
	uint64_t seamTicks = b2GetTicks();

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

	if ( slot->group != NULL )
	{
		b2SeamSolveInGroup( world, context, slot );
		return;
	}

	if ( world->enableWarmStarting == false )
	{
		// b2Prepare*Joint is about to zero the joints' impulses on the host (e.g. src/revolute_joint.c:273-280): what the
		// previous step computed must be in the b2JointSims before that, or it would be written over the zeroes later
		b2SeamFlushImpulses( world, NULL );
	}

	b2SolverSet* awakeSet = world->solverSets.data + b2_awakeSet;
	const bool captured = slot->islandsCaptured;
	slot->islandsCaptured = false;
	if ( !captured )
	{
		b2SeamReserveIslands( slot, world );
	}

	// the team (see above): helpers for as much work as is worth a wake-up
	b2SeamTeam team;
	team.world = world;
	team.solver = slot->solver;
	team.sims = awakeSet->bodySims.data;
	team.labels = slot->islandLabels;
	team.labelCount = captured ? 0 : awakeSet->bodySims.count;
	team.labelBlocks = ( team.labelCount + B2_SEAM_LABEL_BLOCK - 1 ) / B2_SEAM_LABEL_BLOCK;
	b2AtomicStoreInt( &team.labelNext, 0 );
	b2AtomicStoreInt( &team.labelDone, 0 );
	team.context = context;
	team.jointArrayCount = 0;
	team.jointStarts[0] = 0;
	for ( int i = 0; i < B2_GRAPH_COLOR_COUNT; ++i ) // the overflow colour is the last one (B2_OVERFLOW_INDEX)
	{
		b2GraphColor* color = world->constraintGraph.colors + i;
		if ( color->jointSims.count > 0 )
		{
			team.jointArrays[team.jointArrayCount] = color->jointSims.data;
			team.jointStarts[team.jointArrayCount + 1] = team.jointStarts[team.jointArrayCount] + color->jointSims.count;
			team.jointArrayCount += 1;
		}
	}
	team.jointBlocks = ( team.jointStarts[team.jointArrayCount] + B2_SEAM_JOINT_BLOCK - 1 ) / B2_SEAM_JOINT_BLOCK;
	b2AtomicStoreInt( &team.jointNext, 0 );
	b2AtomicStoreInt( &team.jointDone, 0 );
	b2AtomicStoreInt( &team.phase, b2_seamLabels );
	b2AtomicStoreInt( &team.inPack, 0 );
	b2AtomicStoreInt( &team.inUnpack, 0 );
	b2AtomicStoreInt( &team.failed, 0 );

	int itemEstimate = awakeSet->bodySims.count;
	for ( int i = 0; i < B2_GRAPH_COLOR_COUNT; ++i )
	{
		itemEstimate += world->constraintGraph.colors[i].contactSims.count + world->constraintGraph.colors[i].jointSims.count;
	}
	void* handles[B2_MAX_WORKERS];
	int helperCount = world->workerCount - 1;
	static int itemsPerHelper = 0;
	if ( itemsPerHelper == 0 )
	{
		itemsPerHelper = b2SeamEnvInt( "B2GPU_SEAM_ITEMS_PER_HELPER", 1024 );
		itemsPerHelper = itemsPerHelper < 1 ? 1 : itemsPerHelper;
	}
	int useful = itemEstimate / itemsPerHelper; /* a helper that wakes up for less than that only costs */
	helperCount = helperCount < useful ? helperCount : useful;
	{
		static int maxHelpers = -2;
		if ( maxHelpers == -2 )
		{
			maxHelpers = b2SeamEnvInt( "B2GPU_SEAM_HELPERS", -1 );
		}
		helperCount = maxHelpers >= 0 && helperCount > maxHelpers ? maxHelpers : helperCount;
	}
	int enqueued = 0;
	for ( int i = 0; i < helperCount && world->taskCount < B2_MAX_TASKS; ++i )
	{
		handles[enqueued++] = world->enqueueTaskFcn( b2SeamTeamHelper, &team, world->userTaskContext );
		world->taskCount += 1;
	}

	// While the helpers wake up and start on the joints and the island labels, the calling thread lays the step out: the
	// descriptor and BeginStep read the arrays' addresses and counts and the islands' sizes, not the joints' contents or the labels
	// (those are read by the pack pass, which opens when the team's first pass is done).
	int spins = 0;
	b2GpuStepDesc* desc = &slot->lastDesc;
	b2GpuSeam_BuildDesc( world, context, desc );
	if ( captured )
	{
		desc->bodyIsland = slot->islandLabels;
		desc->islandSizes = slot->islandSizes;
		desc->islandCount = slot->capturedIslandCount;
	}
	else
	{
		// sizes here, labels by the team above
		b2GpuSeam_FillIslands( world, desc, NULL, slot->islandSizes, false );
		desc->bodyIsland = slot->islandLabels;
	}

	// The narrow phase's word on the manifolds it recycled.  It only holds when exactly one narrow phase ran since this
	// world's previous solve: a step with dt == 0 runs the narrow phase but not the solver (src/physics_world.c:912-950),
	// and what it re-evaluated the device has never seen.
	if ( slot->collidesSinceSolve == 1 && slot->recycled != NULL )
	{
		desc->recycled = slot->recycled;
		desc->recycledStamp = slot->recycledStamp;
		for ( int c = 0; c <= desc->activeColorCount; ++c )
		{
			const b2GpuColorDesc* color = c < desc->activeColorCount ? desc->colors + c : &desc->overflow;
			int i = color->colorIndex;
			desc->recycledStart[c] = slot->collideStart[i];
			desc->recycledCount[c] = slot->collideCount[i] < color->contactCount ? slot->collideCount[i] : color->contactCount;
			// the colour's array is what the narrow phase walked (nothing added, removed or moved since): entry j is contact j
			desc->recycledInPlace[c] = slot->colorTouched[i] ? 0 : 1;
		}
	}
	slot->collidesSinceSolve = 0;

	b2GpuStepResult* result = &slot->lastResult;
	memset( result, 0, sizeof( *result ) );
	b2TaskContext* taskContext0 = world->taskContexts.data + 0;
	result->hitEventBits = taskContext0->hitEventBitSet.bits;
	result->jointEventBits = taskContext0->jointStateBitSet.bits;

	slot->totals.beforeMs += b2GetMilliseconds( seamTicks );
	const char* failure = NULL;
	if ( b2GpuSolverBeginStep( slot->solver, desc, result ) != 0 )
	{
		failure = "b2GpuSolverBeginStep failed";
	}
	// island hint: lets the device solve islands independently in shared memory (no grid barriers)
	b2SeamTeamJoints( &team );
	b2SeamTeamLabels( &team );
	while ( b2AtomicLoadInt( &team.labelDone ) < team.labelBlocks || b2AtomicLoadInt( &team.jointDone ) < team.jointBlocks )
	{
		b2SeamRelax( &spins );
	}
	if ( failure == NULL )
	{
		b2AtomicStoreInt( &team.phase, b2_seamPackOpen );
		if ( b2GpuSolverPackWork( slot->solver, 1 ) != 0 )
		{
			failure = "packing / upload failed";
		}
	}
	// nobody may still be inside PackWork when Submit re-arms the block counter for the unpack pass
	b2AtomicStoreInt( &team.phase, b2_seamPackClosed );
	while ( b2AtomicLoadInt( &team.inPack ) != 0 )
	{
		b2SeamRelax( &spins );
	}
	if ( failure == NULL && b2GpuSolverSubmit( slot->solver ) != 0 )
	{
		failure = "device solve failed";
	}
	if ( failure == NULL )
	{
		// the helpers unpack behind the download
		b2AtomicStoreInt( &team.phase, b2_seamUnpackOpen );
		if ( b2GpuSolverUnpackWork( slot->solver, 1 ) != 0 )
		{
			failure = "device solve / download failed";
		}
	}
	b2AtomicStoreInt( &team.phase, b2_seamDone );
	while ( b2AtomicLoadInt( &team.inUnpack ) != 0 )
	{
		b2SeamRelax( &spins );
	}
	for ( int i = 0; i < enqueued; ++i )
	{
		if ( handles[i] != NULL )
		{
			world->finishTaskFcn( handles[i], world->userTaskContext );
		}
	}
	if ( failure == NULL && b2AtomicLoadInt( &team.failed ) != 0 )
	{
		failure = "a worker failed in packing / unpacking";
	}
	if ( failure == NULL && b2GpuSolverEndStep( slot->solver, result ) != 0 )
	{
		failure = "b2GpuSolverEndStep failed";
	}
	if ( failure != NULL )
	{
		b2SeamFatal( failure );
	}
	if ( result->hasHitEvents != 0 && b2GpuSolverDeferredPending( slot->solver ) != 0 )
	{
		// deferred impulses: some contact of the step reports a hit event (the device's flag).  The events are built from the
		// manifolds of the flagged contacts (src/solver.c:1745-1790), and the flags are in the records.
		b2SeamFlushImpulses( world, result );
	}

	slot->totals.seamMs += b2GetMilliseconds( seamTicks );
	slot->totals.packMs += result->uploadMs;
	slot->totals.waitMs += result->waitMs;
	slot->totals.unpackMs += result->scatterMs;
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
