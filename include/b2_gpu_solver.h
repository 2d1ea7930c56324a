/*
 * b2_gpu_solver.h -- C-ABI of the B200-native Soft Step constraint solver.
 *
 * This is the drop-in boundary for ONE hot path of Box2D v3.2: the region of b2Solve between
 * "Solver Setup" and "Update Transforms" (reference src/solver.c:1560-1616), i.e. everything
 * b2SolverTask (src/solver.c:1010-1198) runs over the constraint-graph colours.  The reference has
 * no FFI for this path (it is an internal seam); the entry points below are what
 * a maintainer's solver.c would call instead of enqueueing b2SolverTask workers.  INTEGRATION.md
 * shows the reference-side patch.
 *
 * Conventions
 *  - plain C, plain pointers and sizes; all pointers in b2GpuStepDesc are HOST pointers that are only
 *    valid for the duration of the call (the reference may reallocate/reorder its arrays between
 *    steps: src/constraint_graph.c:198-211).
 *  - the arrays are the reference's own structures, bit for bit (release build, float precision):
 *        b2BodyState  32 B  src/body.h:153-168
 *        b2BodySim    96 B  src/body.h:175-208
 *        b2ContactSim 200 B src/contact.h:103-142  (b2Manifold include/box2d/collision.h:572-586)
 *        b2JointSim   252 B src/joint.h:267-301    (already prepared on the host by b2PrepareJoint,
 *                                                   src/joint.c:1406)
 *    The byte offsets this library relies on are listed in b2gpu_layout.h and static-asserted against
 *    the reference headers when the host seam is compiled (box2d_b200/host/b2_gpu_seam.c).
 *  - every function returns 0 on success, non-zero on failure; b2GpuGetLastError() gives the text.
 *    There is NO CPU fallback: if CUDA is unavailable the create call fails.
 */
#ifndef B2_GPU_SOLVER_H
#define B2_GPU_SOLVER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#if defined( _WIN32 )
#define B2GPU_API __declspec( dllexport )
#else
#define B2GPU_API __attribute__( ( visibility( "default" ) ) )
#endif

#define B2GPU_GRAPH_COLOR_COUNT 24 /* include/box2d/constants.h:29 B2_GRAPH_COLOR_COUNT */
#define B2GPU_MAX_ACTIVE_COLORS ( B2GPU_GRAPH_COLOR_COUNT - 1 )

/* b2Softness, src/solver.h:127-132 */
typedef struct b2GpuSoftness
{
	float biasRate;
	float massScale;
	float impulseScale;
} b2GpuSoftness;

/* One graph colour as the solver sees it: colors[i].contactSims / colors[i].jointSims
 * (src/constraint_graph.h:26-48). */
typedef struct b2GpuColorDesc
{
	void* contactSims; /* b2ContactSim[contactCount], in: manifold+material, out: manifold impulses */
	void* jointSims;   /* b2JointSim[jointCount], prepared; in/out (accumulated impulses live in place) */
	int contactCount;
	int jointCount;
	int colorIndex; /* index into b2ConstraintGraph::colors, informational */
	int reserved;
} b2GpuColorDesc;

/* size of one awake island, see b2GpuStepDesc::islandSizes */
typedef struct b2GpuIslandSize
{
	int bodyCount;
	int contactCount; /* touching contacts of the island, incl. those with static bodies */
	int jointCount;
	int reserved;
} b2GpuIslandSize;

/* What the narrow phase knows about a contact it RECYCLED this step (src/physics_world.c:508-560): the b2ContactSim is what
 * the previous step left behind -- the solver's own impulse outputs included -- except for the two separations, which are
 * recomputed, and the body indices / inverse masses, which are refreshed from the bodies.  See b2GpuStepDesc::recycled. */
typedef struct b2GpuRecycledContact
{
	uint32_t stamp; /* == b2GpuStepDesc::recycledStamp when the entry was written for this step */
	int contactId;	/* b2ContactSim::contactId of the contact the entry was written for */
	float separation[2]; /* b2ManifoldPoint::separation of the two points */
	int indexA, indexB;	 /* b2ContactSim::bodySimIndexA / B as the narrow phase refreshed them (src/physics_world.c:497-504) */
} b2GpuRecycledContact;

/* Everything b2SolverTask reads from b2StepContext (src/solver.h:155-237) and b2World
 * (src/physics_world.h:162-216). */
typedef struct b2GpuStepDesc
{
	/* b2StepContext scalars, filled by b2World_Step (src/physics_world.c:907-935) */
	float dt;
	float inv_dt;
	float h;
	float inv_h;
	int subStepCount;
	b2GpuSoftness contactSoftness;
	b2GpuSoftness staticSoftness;
	float restitutionThreshold;
	float maxLinearVelocity;

	/* b2World fields read inside the stages */
	float gravity[2];
	float contactSpeed;
	float contactHertz;
	float contactDampingRatio;
	float hitEventThreshold;
	float lengthUnitsPerMeter; /* b2GetLengthUnitsPerMeter(), read by src/prismatic_joint.c:513,560 */
	int enableWarmStarting;
	int enableContactSoftening;

	/* awake solver set (src/solver.c:1300-1301) */
	void* states;     /* b2BodyState[awakeBodyCount], in/out */
	const void* sims; /* b2BodySim[awakeBodyCount], in */
	int awakeBodyCount;

	/* active colours in ascending colour index (src/solver.c:1341-1367), then the overflow colour
	 * colors[B2_OVERFLOW_INDEX] (src/constraint_graph.h:20) */
	int activeColorCount;
	b2GpuColorDesc colors[B2GPU_MAX_ACTIVE_COLORS];
	b2GpuColorDesc overflow;

	/* bit-set sizing (src/solver.c:1563-1564) */
	int contactIdCapacity;
	int jointIdCapacity;

	/* Optional partition hint: for every awake body the index of its simulation island among the awake islands
	 * (b2Island::localIndex of world->islands[b2Body::islandId], src/island.h:49-74), -1 if the body has none.
	 * Constraints never couple dynamic bodies of different islands (src/island.c:194-330), which lets the device
	 * solve islands independently in shared memory with no grid-wide barriers.  NULL = no hint: the whole step is
	 * solved by the grid-barrier kernel.  The result does not depend on the hint (bit-identical either way). */
	const int* bodyIsland;
	int islandCount;
	int reserved0;

	/* Optional, with bodyIsland: how big every island is (b2Island::bodies / contacts / joints .count, src/island.h:64-73).
	 * With it the bins are packed by their real size instead of an estimate from the body counts, and nothing has to be
	 * counted on the step's critical path.  Counts may be slightly off (they are only used for sizing; the device checks).
	 * NULL = the library counts the bodies per island itself and estimates the rest. */
	const struct b2GpuIslandSize* islandSizes;

	/* Optional (resident mode): recycled[recycledStart[c] + i] describes contact i of colors[c] (c == activeColorCount: the
	 * overflow colour) for i < recycledCount[c].  An entry counts only when its stamp equals recycledStamp AND its contactId
	 * is that contact's; the caller then vouches that, since the previous b2GpuSolverStep of this solver, nothing the solver
	 * reads from the contact's b2ContactSim has changed but the separations given in the entry, the body indices (re-read
	 * by the library) and the inverse masses, which equal the bodies'.  The pack pass then skips re-reading and comparing the
	 * record.  Any other entry (stale stamp, another contact's id: the contact moved in its array after the narrow phase
	 * ran) is ignored and the contact is examined as usual.  NULL = no hint.  Results do not depend on the hint as long as
	 * the caller's claim is true. */
	const b2GpuRecycledContact* recycled;
	uint32_t recycledStamp;
	int recycledStart[B2GPU_MAX_ACTIVE_COLORS + 1];
	int recycledCount[B2GPU_MAX_ACTIVE_COLORS + 1];
	/* Optional, with recycled: recycledInPlace[c] != 0 = the caller also vouches that the contact array of colors[c] has not
	 * changed since the entries were written -- no contact added, removed or moved (src/constraint_graph.c:66-211) -- so entry
	 * i IS contact i and its indexA / indexB are the contact's: the pack pass takes a current entry without looking at the
	 * contact at all.  0 = it checks the contact's id and body indices itself (one cache line per contact). */
	int recycledInPlace[B2GPU_MAX_ACTIVE_COLORS + 1];
} b2GpuStepDesc;

/* Index of each per-stage timer, same split as b2Profile (include/box2d/types.h:526-551) filled by the
 * orchestrator at src/solver.c:1080,1097,1112,1132,1141,1159,1182,1191. */
enum
{
	b2GpuStage_prepareConstraints = 0,
	b2GpuStage_integrateVelocities = 1,
	b2GpuStage_warmStart = 2,
	b2GpuStage_solveImpulses = 3,
	b2GpuStage_integratePositions = 4,
	b2GpuStage_relaxImpulses = 5,
	b2GpuStage_applyRestitution = 6,
	b2GpuStage_storeImpulses = 7,
	b2GpuStage_count = 8
};

typedef struct b2GpuStepResult
{
	/* Caller-provided bit sets, b2BitSet::bits layout (src/bitset.h): bit i of word i/64.  Sized for
	 * contactIdCapacity / jointIdCapacity bits rounded up to 64.  The library ORs into them.  May be
	 * NULL when the corresponding capacity is 0. */
	uint64_t* hitEventBits;   /* src/contact_solver.c:2305-2320 */
	uint64_t* jointEventBits; /* src/joint.c:1663-1675 */
	int hasHitEvents;

	/* device time per stage group in milliseconds (CUDA globaltimer inside the step kernel) */
	float stageMs[b2GpuStage_count];
	float kernelMs;     /* CUDA-event time of all kernels of the step */
	float totalMs;      /* host wall time of the whole call, copies included */
	uint64_t h2dBytes;  /* bytes copied host -> device during the call */
	uint64_t d2hBytes;  /* bytes copied device -> host during the call */
	int kernelLaunches; /* kernels launched by the call */
	int gridBarriers;   /* grid-wide barriers executed inside the step kernel */
	/* host wall-clock split of b2GpuSolverStep, milliseconds */
	float uploadMs;   /* descriptor -> params, buffer growth, enqueue of the H2D copies */
	float waitMs;     /* launch + waiting for H2D, kernels and D2H to drain */
	float scatterMs;  /* writing the impulses back into the reference's manifolds + event bits */
	float h2dMs;      /* CUDA-event time of the H2D copies; only read when B2GPU_TRACE is set (a driver call per step), else 0 */
} b2GpuStepResult;

typedef struct b2GpuSolver b2GpuSolver;

/* Create a solver bound to one CUDA device; owns a stream and geometrically grown device buffers.  One
 * solver per world (different worlds may step concurrently, include/box2d/box2d.h:31-32).  Returns NULL on
 * failure (no device, no driver): there is no CPU fallback. */
B2GPU_API b2GpuSolver* b2GpuSolverCreate( int device );
B2GPU_API void b2GpuSolverDestroy( b2GpuSolver* solver );

/* The whole hot path for one world step: upload, all stages of src/solver.c:1055-1197 on the device,
 * download.  Replaces the worker enqueue + b2SolverTask at src/solver.c:1563-1605.  Synchronous. */
B2GPU_API int b2GpuSolverStep( b2GpuSolver* solver, const b2GpuStepDesc* desc, b2GpuStepResult* result );

/* The same step split in three so a benchmark can time the device part with inputs resident in HBM:
 * Upload copies the inputs, Run executes all stages from the resident inputs (it does not modify them, so
 * it can be repeated), Download copies the outputs of the last Run back into the desc's host arrays. */
B2GPU_API int b2GpuSolverUpload( b2GpuSolver* solver, const b2GpuStepDesc* desc );
B2GPU_API int b2GpuSolverRun( b2GpuSolver* solver, b2GpuStepResult* result );
B2GPU_API int b2GpuSolverDownload( b2GpuSolver* solver, const b2GpuStepDesc* desc, b2GpuStepResult* result );

/* The same step in phases, so the host side can use ITS OWN worker threads for the two memory-bound host passes
 * (the seam runs them through the reference's b2ParallelFor): Begin fixes the layout; PackRange converts a range of
 * items (first the awake bodies, then all contacts in colour order, overflow last) from the reference's arrays into
 * the page-locked wire buffer and may be called concurrently on disjoint ranges; Submit enqueues H2D + kernels + D2H
 * and returns immediately; Wait blocks; UnpackRange writes the impulses of a range of contacts back into the
 * reference's manifolds (+ hit-event bits) and may be called concurrently on disjoint ranges; End fills the result.
 * The desc's arrays must stay valid and unchanged from Begin to End. */
B2GPU_API int b2GpuSolverBeginStep( b2GpuSolver* solver, const b2GpuStepDesc* desc, b2GpuStepResult* result );
B2GPU_API int b2GpuSolverGetPackItemCount( const b2GpuSolver* solver );
B2GPU_API void b2GpuSolverPackRange( b2GpuSolver* solver, int begin, int end );
B2GPU_API int b2GpuSolverSubmit( b2GpuSolver* solver );
B2GPU_API int b2GpuSolverWait( b2GpuSolver* solver );
B2GPU_API int b2GpuSolverGetUnpackItemCount( const b2GpuSolver* solver );
B2GPU_API void b2GpuSolverUnpackRange( b2GpuSolver* solver, int begin, int end );
B2GPU_API int b2GpuSolverEndStep( b2GpuSolver* solver, b2GpuStepResult* result );

/* Pipelined form of the two host passes.  After Begin, every participating host thread calls PackWork once; a call
 * claims blocks of items until none are left.  Exactly one of the callers passes pump = 1: it also starts the upload
 * of every finished prefix of the wire buffer, so PCIe runs behind the packing instead of after it.  When all calls
 * have returned the caller of pump = 1 calls Submit (kernels + chunked download), then every participating thread
 * calls UnpackWork: blocks are unpacked as soon as their part of the download has arrived (the pump = 1 caller watches
 * the download events and handles the island-kernel fallback), so unpacking runs behind the download.  UnpackWork
 * replaces Wait + UnpackRange.  Both return 0 on success. */
B2GPU_API int b2GpuSolverPackWork( b2GpuSolver* solver, int pump );
B2GPU_API int b2GpuSolverUnpackWork( b2GpuSolver* solver, int pump );

/* Deferred contact impulses (resident mode of a single world, phased entry points).  What a step writes into the
 * manifolds -- b2StoreImpulsesTask, src/contact_solver.c:2293-2320 -- is read back by the host far less often than it is
 * written: a manifold the narrow phase RECYCLES (src/physics_world.c:508-560) is not looked at, and the device warm-starts
 * such a contact from its own previous output.  When enabled, EndStep returns once the body states and the joints'
 * outputs are in place; the impulse records follow behind the caller's back and stay in the library's page-locked output
 * arena, and the CALLER promises to call b2GpuSolverMaterializeContacts on a contact before anything reads its manifold's
 * impulses (normalImpulse, tangentImpulse, totalNormalImpulse, normalVelocity, rollingImpulse) or moves it out of the
 * awake contact arrays for good: the narrow phase before it re-evaluates the manifold (src/contact.c:523), island sleep
 * (src/solver_set.c:155), the contact-data and snapshot API, hit events (b2GpuStepResult::hasHitEvents is then set from
 * the device's flag and the bits are set by the materialize call).  The pack pass of the next step materializes whatever it
 * has to read in full itself.  Must be set before the first step (or between steps with nothing pending); off by default. */
B2GPU_API int b2GpuSolverSetDeferredImpulses( b2GpuSolver* solver, int enabled );
/* 1 while some manifolds may not have received the last step's impulses */
B2GPU_API int b2GpuSolverDeferredPending( const b2GpuSolver* solver );
/* Waits for the tail of the last step's download (otherwise the first materialize call does); 0 on success. */
B2GPU_API int b2GpuSolverDeferredSync( b2GpuSolver* solver );
/* Writes the pending impulses of `count` consecutive b2ContactSim -- the contacts at places firstIndex .. of the colour whose
 * b2GpuColorDesc::colorIndex was `colorIndex` (the overflow colour: B2GPU_GRAPH_COLOR_COUNT - 1) -- into their manifolds,
 * those that have some pending, and with `result` ORs their hit-event bits into result->hitEventBits.  A pending record is
 * found by PLACE: the contact must still be where it was when the step was solved (the library checks the contact id and
 * skips a contact that is not), so the caller materializes a contact BEFORE it moves it in its colour's array (the last
 * contact of a colour before a swap-remove, src/constraint_graph.c:198-211) and calls b2GpuSolverDeferredForget for a place
 * whose contact leaves the array.  May be called concurrently for different places.  Needs the colours' indices to be
 * ascending (what the reference produces); otherwise nothing is deferred.  Returns the number of manifolds written, -1 on a
 * device error. */
B2GPU_API int b2GpuSolverMaterializeContacts( b2GpuSolver* solver, int colorIndex, int firstIndex, void* contactSims, int count,
											   b2GpuStepResult* result );
/* The contact at this place is leaving its colour's array (it stopped touching, or is destroyed): its record is void. */
B2GPU_API void b2GpuSolverDeferredForget( b2GpuSolver* solver, int colorIndex, int index );
/* The same for the joints: with deferred impulses the fields the stages write in a b2JointSim (the accumulated impulses,
 * src/*_joint.c b2WarmStart* / b2Solve*) stay on the library's side too, and MaterializeJoints writes them into `count`
 * consecutive b2JointSim -- the joints at places firstIndex .. of the colour `colorIndex` -- before anything reads or
 * changes them (the per-type b2*Joint_Get* / Set* functions, b2Joint_GetConstraintForce / Torque), before a joint moves in
 * its colour's array (src/constraint_graph.c:299-325) or leaves the awake set.  The caller must not defer across a step
 * with warm starting disabled (b2Prepare*Joint then zeroes the impulses on the host, e.g. src/revolute_joint.c:273-280). */
B2GPU_API int b2GpuSolverMaterializeJoints( b2GpuSolver* solver, int colorIndex, int firstIndex, void* jointSims, int count );
B2GPU_API void b2GpuSolverDeferredForgetJoint( b2GpuSolver* solver, int colorIndex, int index );
/* Every manifold that matters has been materialized (or the host's contacts were replaced wholesale): nothing is pending. */
B2GPU_API void b2GpuSolverDeferredDone( b2GpuSolver* solver );

/* Batch of independent worlds (the RL-style workload): one launch solves all of them, one thread block
 * per world with the world's bodies and constraints resident in shared memory for all sub-steps.  Each
 * desc is a complete, independent world step.  Worlds that do not fit the per-block budget are solved
 * one after the other with the single-world kernel. */
B2GPU_API int b2GpuSolverUploadBatch( b2GpuSolver* solver, const b2GpuStepDesc* descs, int worldCount );
B2GPU_API int b2GpuSolverRunBatch( b2GpuSolver* solver, b2GpuStepResult* result );
B2GPU_API int b2GpuSolverDownloadBatch( b2GpuSolver* solver, const b2GpuStepDesc* descs, int worldCount,
										 b2GpuStepResult* results );
B2GPU_API int b2GpuSolverStepBatch( b2GpuSolver* solver, const b2GpuStepDesc* descs, int worldCount,
									 b2GpuStepResult* results );

/* Execution mode knobs (for tests and profiling).  mode 0 = one persistent cooperative kernel per step
 * (default), mode 1 = one kernel launch per stage (same device functions, stream ordered). */
B2GPU_API int b2GpuSolverSetMode( b2GpuSolver* solver, int mode );

/* Page-locked host memory for the reference's own arrays, to be installed with b2SetAllocator
 * (include/box2d/base.h:86) so uploads are true DMA.  Signatures match b2AllocFcn / b2FreeFcn. */
B2GPU_API void* b2GpuHostAlloc( size_t size, int alignment );
B2GPU_API void b2GpuHostFree( void* mem, size_t size );

/* Diagnostics */
/* Steps of this solver that ran on the bin lists the previous step left on the device (a steady scene: no contact travelled
 * in full, every body in the bin it was in, same layout and plan; the partition kernel is not launched). */
B2GPU_API int b2GpuSolverGetListReuseCount( const b2GpuSolver* solver );
B2GPU_API const char* b2GpuGetLastError( void );
B2GPU_API int b2GpuGetDeviceCount( void );
B2GPU_API int b2GpuGetVersion( void );
/* Totals since creation */
B2GPU_API uint64_t b2GpuSolverGetLaunchCount( const b2GpuSolver* solver );
/* How the last step was laid out for the island-local kernels: number of bins (0 = the step was planned for the
 * grid-barrier kernel) and thread blocks per bin (1, or the cluster size 2..16).  Returns binCount. */
B2GPU_API int b2GpuSolverGetIslandPlan( const b2GpuSolver* solver, int* binCount, int* blocksPerBin );
/* Resident mode (single worlds): the device keeps the contacts' static data, their impulses and the bodies across steps, and
 * the pack pass only uploads what differs from that -- a 16-byte record per contact whose manifold the narrow phase recycled
 * (src/physics_world.c:508-560), the full 96 bytes otherwise.  Returns 1 when the last step ran in resident mode and fills
 * how many contacts travelled as full records, how many bodies were re-uploaded and how many contacts were taken on the
 * caller's word (b2GpuStepDesc::recycled) without their record being read, and how many joints travelled as full 256-byte
 * records (the others as 96 bytes: the fields b2PrepareJoint rewrites every step); 0 (counts untouched) otherwise.  The
 * results do not depend on the mode (bit-identical); B2GPU_RESIDENT=0 turns it off.  Any of the pointers may be NULL. */
B2GPU_API int b2GpuSolverGetResidentStats( const b2GpuSolver* solver, int* fullContacts, int* dirtyBodies, int* vouchedContacts,
										  int* fullJoints );
/* Host utility for callers that have island labels but no island bookkeeping: fill sizes[desc->islandCount] from
 * desc->bodyIsland and the constraint arrays (one pass over the constraints).  Returns 0 on success. */
B2GPU_API int b2GpuCountIslandSizes( const b2GpuStepDesc* desc, b2GpuIslandSize* sizes );

#ifdef __cplusplus
}
#endif

#endif /* B2_GPU_SOLVER_H */
