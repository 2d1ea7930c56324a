/*
 * b2_gpu_seam.h -- host side of the drop-in boundary (C17, compiled INTO the reference's library).
 *
 * The reference has no plugin API for its solver; the boundary is the internal seam inside b2Solve
 * between "Solver Setup" and "Update Transforms" (reference src/solver.c:1560-1616).  The generated
 * solver.c (tools/patch_solver.py --mode gpu) calls b2GpuSeam_SolveConstraints where the reference
 * enqueues its b2SolverTask workers (src/solver.c:1563-1605).  Everything else in b2Solve -- setup,
 * island split, finalize, events, refit, bullets, sleep -- is the reference's own code, unchanged.
 */
#pragma once

#include "b2_gpu_solver.h"

#include <stdbool.h>

typedef struct b2World b2World;
typedef struct b2StepContext b2StepContext;

#ifdef __cplusplus
extern "C"
{
#endif

/* Fill the C-ABI step descriptor from the reference's world + step context.  Joint sims must already be
 * prepared (b2PrepareJoint, src/joint.c:1406).  Shared by the product seam and the oracle's capture hook. */
void b2GpuSeam_BuildDesc( b2World* world, b2StepContext* stepContext, b2GpuStepDesc* desc );

/* Fill the optional island hint of the descriptor: labels[i] = index of awake body i's island among the awake islands,
 * sizes[k] = bodies / touching contacts / joints of awake island k (may be NULL).  `labels` must hold awakeBodyCount
 * ints, `sizes` one entry per awake island; both must stay valid for the duration of the solver call. */
void b2GpuSeam_FillIslands( b2World* world, b2GpuStepDesc* desc, int* labels, b2GpuIslandSize* sizes, bool parallel );

/* Prepare every awake joint on the host (b2ParallelFor over the flat joint range + the overflow colour),
 * i.e. the b2_stagePrepareJoints stage and b2PrepareJoints_Overflow (src/solver.c:1060-1077). */
void b2GpuSeam_PrepareJoints( b2World* world, b2StepContext* stepContext );

/* The seam itself: host joint prepare -> b2GpuSolverStep -> event bits + b2Profile.  Aborts (B2_ASSERT-style)
 * if the device solver fails: there is no CPU fallback. */
void b2GpuSeam_SolveConstraints( b2World* world, b2StepContext* stepContext );

/* Called right before b2Solve may start its island-split task: takes the island hint while nobody is rewriting the islands. */
void b2GpuSeam_BeforeIslandSplit( b2World* world, b2StepContext* stepContext );

/* Called by the generated physics_world.c (tools/patch_collide.py): the narrow phase begins / has recycled a manifold. */
typedef struct b2ContactSim b2ContactSim;
void b2GpuSeam_BeginCollide( b2World* world, b2StepContext* stepContext, int contactCount );
void b2GpuSeam_ContactRecycled( b2World* world, int contactIndex, const b2ContactSim* contactSim );

/* Deferred impulses (b2_gpu_seam.c): called by the generated physics_world.c before b2UpdateContact re-evaluates a manifold. */
void b2GpuSeam_ContactReevaluated( b2World* world, int workerIndex, int contactIndex, b2ContactSim* contactSim );
/* Writes every pending impulse record into its manifold now (what the interposed readers do); stats for tests. */
void b2GpuSeam_FlushImpulses( int worldIndex );
int b2GpuSeam_GetDeferredStats( int worldIndex, int* pending, long long* flushes );

/* Route the reference's allocations through page-locked memory (b2SetAllocator, include/box2d/base.h:86).
 * Call before creating any world. */
void b2GpuSeam_InstallPinnedAllocator( void );

/* World-level batched step (the RL-style workload).  A GROUP is a set of independent worlds that
 * the application steps concurrently -- one thread per world, every world with the same time step and sub-step count
 * (include/box2d/box2d.h:31-32: different worlds may be stepped from different threads).  Inside b2World_Step each world
 * runs its own broad phase, narrow phase and solver setup as always; at the seam the worlds of a group meet, and the last
 * one to arrive solves ALL of them with one b2GpuSolverStepBatch (one launch sequence, one pair of transfers) while the
 * others wait; then every world finalizes on its own thread.  Results are those of stepping each world alone (bit for
 * bit).  Every world of a group must be stepped in every round, or the others wait for it.  Worlds of a group should be
 * created with workerCount = 1 (their parallelism is the group).  Returns a group id >= 0, or -1. */
int b2GpuSeam_CreateGroup( const int* worldIndices, int worldCount );
void b2GpuSeam_DestroyGroup( int group );

/* Sums over the steps of a world slot since the last reset (for benchmarks). */
typedef struct b2GpuSeamTotals
{
	double kernelMs; /* CUDA-event time of the step's kernels */
	double abiMs;	 /* host wall time of the phased C-ABI calls */
	double h2dBytes, d2hBytes;
	double stageMs[b2GpuStage_count];
	long long steps, launches, gridBarriers;
	double seamMs;	 /* host wall time of the whole seam call (what b2Profile.constraints reports) */
	double packMs, waitMs, unpackMs; /* Begin..Submit, Submit..first output seen, ..EndStep of the phased C-ABI calls */
	double beforeMs; /* seam entry .. BeginStep: bit-set clears, host joint prepare, island labels and sizes, the descriptor */
} b2GpuSeamTotals;
void b2GpuSeam_GetTotals( int worldIndex, b2GpuSeamTotals* totals, int reset );

/* Last step's device-side result for a world id slot (for benchmarks / tests). */
const b2GpuStepResult* b2GpuSeam_GetLastResult( int worldIndex );
const b2GpuStepDesc* b2GpuSeam_GetLastDesc( int worldIndex );
/* b2GpuSolverGetResidentStats of the world's solver for its last step (0 = no solver or not a resident step). */
int b2GpuSeam_GetResidentStats( int worldIndex, int* fullContacts, int* dirtyBodies, int* vouchedContacts );

/* 0 = persistent cooperative kernel, 1 = one launch per stage.  Applies to solvers created afterwards and
 * to existing ones. */
void b2GpuSeam_SetMode( int mode );

/* Destroy all device solvers created by the seam. */
void b2GpuSeam_Shutdown( void );
/* Device solvers are created lazily per world id and follow the world's generation: a world that reuses a destroyed world's id
 * gets a fresh solver at its first step.  ReleaseWorld frees a destroyed world's device memory right away (optional).
 * HasSolver: 0 = the slot has no solver, else 1 + the generation it belongs to (tests). */
void b2GpuSeam_ReleaseWorld( int worldIndex );
int b2GpuSeam_HasSolver( int worldIndex );

#ifdef __cplusplus
}
#endif
