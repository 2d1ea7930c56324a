/*
 * b2h_capture.c -- capture hooks for the oracle build of the reference (test infrastructure).
 *
 * tools/patch_solver.py --mode hook brackets the reference's own solve region (src/solver.c:1563-1605, kept
 * verbatim) with b2OracleHook_BeforeSolve / b2OracleHook_AfterSolve.  When armed, the hooks dump the exact
 * inputs the C-ABI of include/b2_gpu_solver.h would receive for this step and the outputs the reference's CPU
 * solver produced from them.  The dumps are the kernel-level golden vectors under tests/golden/.
 *
 * File format "B2CAP003" (little endian):
 *   char[8] magic | u32 descBytes | b2GpuStepDesc (raw, pointers meaningless)
 *   inputs : states[n*32] sims[n*96] islandLabels[n*4] { contacts[c*200] joints[j*252] } per active colour, then overflow
 *            (joints are PREPARED copies: b2PrepareJoint applied to a copy, src/joint.c:1406)
 *   outputs: states[n*32] { contacts[c*200] joints[j*252] } per active colour, then overflow
 *            u32 hitWords | u64[hitWords] | u32 jointWords | u64[jointWords] | i32 hasHitEvents
 */
#include "b2_gpu_seam.h"

#include "bitset.h"
#include "body.h"
#include "constraint_graph.h"
#include "contact.h"
#include "joint.h"
#include "physics_world.h"
#include "solver.h"
#include "solver_set.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define B2H_API __attribute__( ( visibility( "default" ) ) )

static struct
{
	int armed;
	int active;
	char path[1024];
	FILE* file;
	b2GpuStepDesc desc;
} s_capture;

/* Dump the next b2Solve of ANY world in this library into `path`. */
B2H_API void b2h_capture_arm( const char* path )
{
	snprintf( s_capture.path, sizeof( s_capture.path ), "%s", path );
	s_capture.armed = 1;
}

static void b2hWrite( const void* data, size_t bytes )
{
	if ( bytes > 0 )
	{
		fwrite( data, 1, bytes, s_capture.file );
	}
}

void b2OracleHook_BeforeSolve( b2World* world, b2StepContext* context )
{
	if ( s_capture.armed == 0 )
	{
		return;
	}
	s_capture.armed = 0;
	s_capture.file = fopen( s_capture.path, "wb" );
	if ( s_capture.file == NULL )
	{
		return;
	}
	s_capture.active = 1;

	b2GpuStepDesc* desc = &s_capture.desc;
	b2GpuSeam_BuildDesc( world, context, desc );
	int* labels = malloc( (size_t)( desc->awakeBodyCount + 1 ) * sizeof( int ) );
	b2GpuSeam_FillIslands( world, desc, labels, NULL, false );

	uint32_t descBytes = (uint32_t)sizeof( b2GpuStepDesc );
	b2hWrite( "B2CAP003", 8 );
	b2hWrite( &descBytes, 4 );
	b2hWrite( desc, sizeof( *desc ) );
	b2hWrite( desc->states, (size_t)desc->awakeBodyCount * sizeof( b2BodyState ) );
	b2hWrite( desc->sims, (size_t)desc->awakeBodyCount * sizeof( b2BodySim ) );
	b2hWrite( labels, (size_t)desc->awakeBodyCount * sizeof( int ) );
	free( labels );
	desc->bodyIsland = NULL;
	desc->islandSizes = NULL;

	for ( int c = 0; c <= desc->activeColorCount; ++c )
	{
		const b2GpuColorDesc* color = c < desc->activeColorCount ? desc->colors + c : &desc->overflow;
		b2hWrite( color->contactSims, (size_t)color->contactCount * sizeof( b2ContactSim ) );

		/* what the device receives: joints prepared on the host.  Prepare a COPY, the reference prepares the
		 * originals itself inside b2SolverTask right after this hook. */
		if ( color->jointCount > 0 )
		{
			size_t bytes = (size_t)color->jointCount * sizeof( b2JointSim );
			b2JointSim* copy = malloc( bytes );
			memcpy( copy, color->jointSims, bytes );
			for ( int i = 0; i < color->jointCount; ++i )
			{
				b2PrepareJoint( copy + i, context );
			}
			b2hWrite( copy, bytes );
			free( copy );
		}
	}
}

void b2OracleHook_AfterSolve( b2World* world, b2StepContext* context )
{
	(void)context;
	if ( s_capture.active == 0 )
	{
		return;
	}
	s_capture.active = 0;

	const b2GpuStepDesc* desc = &s_capture.desc;
	b2hWrite( desc->states, (size_t)desc->awakeBodyCount * sizeof( b2BodyState ) );
	for ( int c = 0; c <= desc->activeColorCount; ++c )
	{
		const b2GpuColorDesc* color = c < desc->activeColorCount ? desc->colors + c : &desc->overflow;
		b2hWrite( color->contactSims, (size_t)color->contactCount * sizeof( b2ContactSim ) );
		b2hWrite( color->jointSims, (size_t)color->jointCount * sizeof( b2JointSim ) );
	}

	/* union of the per-worker event bit sets, the way src/solver.c:1654-1658 and :1705-1720 merge them */
	uint32_t hitWords = (uint32_t)( ( desc->contactIdCapacity + 63 ) / 64 );
	uint32_t jointWords = (uint32_t)( ( desc->jointIdCapacity + 63 ) / 64 );
	uint64_t* hit = calloc( hitWords + 1, sizeof( uint64_t ) );
	uint64_t* joint = calloc( jointWords + 1, sizeof( uint64_t ) );
	int hasHitEvents = 0;
	for ( int i = 0; i < world->workerCount; ++i )
	{
		b2TaskContext* tc = world->taskContexts.data + i;
		for ( uint32_t k = 0; k < hitWords && k < tc->hitEventBitSet.blockCount; ++k )
		{
			hit[k] |= tc->hitEventBitSet.bits[k];
		}
		for ( uint32_t k = 0; k < jointWords && k < tc->jointStateBitSet.blockCount; ++k )
		{
			joint[k] |= tc->jointStateBitSet.bits[k];
		}
		hasHitEvents |= tc->hasHitEvents ? 1 : 0;
	}
	b2hWrite( &hitWords, 4 );
	b2hWrite( hit, hitWords * sizeof( uint64_t ) );
	b2hWrite( &jointWords, 4 );
	b2hWrite( joint, jointWords * sizeof( uint64_t ) );
	b2hWrite( &hasHitEvents, 4 );
	free( hit );
	free( joint );

	fclose( s_capture.file );
	s_capture.file = NULL;
}
