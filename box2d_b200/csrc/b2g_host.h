// b2g_host.h -- internals shared by the host-side translation units of the device library (not part of the ABI).
#pragma once

#include "b2_gpu_solver.h"
#include "b2g_types.cuh"

#include <cuda_runtime.h>

#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

// error reporting (b2g_solver.cu): the text of the last failure of the calling thread, returned by b2GpuGetLastError
int b2gFail( const char* what, cudaError_t err );
int b2gFailMsg( const char* what );

#define B2G_CUDA( call )                                                                                                         \
	do                                                                                                                           \
	{                                                                                                                            \
		cudaError_t err_ = ( call );                                                                                             \
		if ( err_ != cudaSuccess )                                                                                               \
		{                                                                                                                        \
			return b2gFail( #call, err_ );                                                                                       \
		}                                                                                                                        \
	}                                                                                                                            \
	while ( 0 )

template <typename T> struct DeviceBuffer
{
	T* ptr = nullptr;
	size_t capacity = 0; // elements

	// grow geometrically, contents are not preserved
	cudaError_t reserve( size_t count )
	{
		if ( count <= capacity )
		{
			return cudaSuccess;
		}
		size_t newCapacity = capacity < 1024 ? 1024 : capacity;
		while ( newCapacity < count )
		{
			newCapacity += newCapacity / 2;
		}
		if ( ptr != nullptr )
		{
			cudaFree( ptr );
			ptr = nullptr;
			capacity = 0;
		}
		cudaError_t err = cudaMalloc( &ptr, newCapacity * sizeof( T ) );
		if ( err == cudaSuccess )
		{
			capacity = newCapacity;
		}
		return err;
	}

	void release()
	{
		if ( ptr != nullptr )
		{
			cudaFree( ptr );
		}
		ptr = nullptr;
		capacity = 0;
	}
};

// page-locked host memory, on huge pages where possible (b2g_alloc.cu)
int b2gHugePagesWanted(); // B2GPU_HUGE_PAGES: 0 none, 1 the library's own host arrays, 2 the page-locked blocks as well
void* b2gPinnedAlloc( size_t bytes, unsigned int flags );
void b2gPinnedFree( void* mem );

template <typename T> struct PinnedBuffer
{
	T* ptr = nullptr;
	size_t capacity = 0;

	cudaError_t reserve( size_t count )
	{
		if ( count <= capacity )
		{
			return cudaSuccess;
		}
		size_t newCapacity = capacity < 1024 ? 1024 : capacity;
		while ( newCapacity < count )
		{
			newCapacity += newCapacity / 2;
		}
		if ( ptr != nullptr )
		{
			b2gPinnedFree( ptr );
			ptr = nullptr;
			capacity = 0;
		}
		ptr = static_cast<T*>( b2gPinnedAlloc( newCapacity * sizeof( T ), cudaHostAllocDefault ) );
		if ( ptr == nullptr )
		{
			return cudaErrorMemoryAllocation;
		}
		capacity = newCapacity;
		return cudaSuccess;
	}

	void release()
	{
		if ( ptr != nullptr )
		{
			b2gPinnedFree( ptr );
		}
		ptr = nullptr;
		capacity = 0;
	}
};

// control block, zeroed before every run
struct ControlBlock
{
	unsigned int barrier[2];
	int hasHitEvents;
	int anyRestitution;
	unsigned long long stageCycles[10];
	int islandFailed; // a bin did not fit its block (binFail): the host reruns the step on the grid-barrier kernel
	int reserved;
};

// segments of the step (see "the step as segments" below)
struct b2gBodySeg
{
	uint8_t* states;
	const uint8_t* sims;
	const int* islands;
	const b2GpuIslandSize* islandSizes; // optional
	int islandCount;
	int islandBase; // first island of this world in the batch-wide numbering
	int count;
	int base;		  // first body of this world in the batch-wide numbering
	int jointBitBase; // first bit of this world in the joint-event bit set
	int jointWords;
};

struct b2gContactSeg
{
	uint8_t* sims;
	int count;
	int slotStart; // wire slot of the segment's first contact (multiple of 4)
	int world;
	bool wide; // false for the overflow colour
	int colorIndex;
	const b2GpuRecycledContact* hints; // optional (b2GpuStepDesc::recycled), entry i describes contact i
	int hintCount;
	uint32_t hintStamp;
	bool hintsInPlace; // b2GpuStepDesc::recycledInPlace: entry i is contact i, the contact need not be looked at
};

struct b2gJointSeg
{
	uint8_t* sims;
	int count;
	int jointStart;
	int world;
	int colorIndex;
	bool overflow;
};

// ---- resident mode: host-side shadows of what the device holds (b2g_types.cuh, b2g_resident.cuh) -----------------------
// One per home (graph colour, index in the colour's array): which contact lives there, the static rows of the record the
// device has in its table, the impulses it computed last (what the unpack pass wrote into the manifold).  A contact is
// "clean" when it was at the same home in the previous step and the record the pack pass would send equals this, bit for
// bit, separations aside.  Homes follow the colour arrays, so both host passes walk the shadows front to back.
struct b2gShadowContact
{
	uint32_t rows[17]; // WR_HEAD.xyz, WR_NORMAL, WR_MATERIAL.xy, WR_ANCHOR1, WR_ANCHOR2
};

// ... what the pack pass needs of a contact that comes with a valid b2GpuRecycledContact (nothing else is read then): who
// lives at the home, its bodies (they may move in the awake set while the manifold is recycled) and whether it has rolling
// resistance / restitution (kMetaGroup* bits of its own, OR-ed over its SIMD group)
struct b2gShadowHead
{
	int contactId;
	int indexA, indexB;
	int ownBits;
};

// ... and, in an array of their own (the unpack pass touches nothing else of the shadows): normalImpulse1, tangentImpulse1,
// normalImpulse2, tangentImpulse2, rollingImpulse
struct b2gShadowImpulses
{
	uint32_t values[5];
};

// ---- deferred contact impulses (b2GpuSolverSetDeferredImpulses) ----------------------------------------------------------
// A contact's record among the outputs the host has not written into the manifolds yet is found by PLACE: the contact sat at
// (graph colour, index in the colour's array) when the pending step was solved = at a home (see b2gShadowContact), the home's
// shadow head says which contact that was, and the record is at the colour's first slot of that step + index.  The caller
// materializes a contact before it moves it (include/b2_gpu_solver.h), so a pending contact is still where it was.
// consumedStamp[home] == the pending step's stamp: that record has been written into its manifold (or must not be).

constexpr int kHomeColors = B2GPU_GRAPH_COLOR_COUNT;

constexpr int kStreamChunk = 16; // records a pack block reserves at a time in the full / dirty-body streams

// The shadows are megabytes of host memory that the pack pass streams through on every step, next to the reference's
// own arrays.  On 4 KB pages that is a few thousand TLB misses per step, each a two-dimensional page walk when the
// host is a guest (every box this was measured on): with the address layout of the process as the dice, a third of the
// runs of many_pyramids packed in 0.055 - 0.09 ms instead of 0.050.  Large shadow arrays are therefore 2 MB aligned and
// advised to transparent huge pages (the default policy of the boxes is `madvise`); small ones stay plain malloc.
template <class T> struct b2gHugeAllocator
{
	using value_type = T;
	static constexpr size_t kHugePage = (size_t)2 << 20;
	b2gHugeAllocator() = default;
	template <class U> b2gHugeAllocator( const b2gHugeAllocator<U>& )
	{
	}
	T* allocate( size_t n )
	{
		size_t bytes = n * sizeof( T );
		void* mem = nullptr;
		if ( bytes >= kHugePage / 4 && b2gHugePagesWanted() >= 1 )
		{
			size_t size = ( bytes + kHugePage - 1 ) / kHugePage * kHugePage;
			mem = aligned_alloc( kHugePage, size );
			if ( mem != nullptr )
			{
				madvise( mem, size, MADV_HUGEPAGE ); // (advice: a refusal costs nothing but the TLB misses)
			}
		}
		else
		{
			mem = malloc( bytes > 0 ? bytes : 1 );
		}
		if ( mem == nullptr )
		{
			throw std::bad_alloc();
		}
		return static_cast<T*>( mem );
	}
	void deallocate( T* mem, size_t )
	{
		free( mem );
	}
	template <class U> bool operator==( const b2gHugeAllocator<U>& ) const
	{
		return true;
	}
	template <class U> bool operator!=( const b2gHugeAllocator<U>& ) const
	{
		return false;
	}
};
template <class T> using b2gHugeVector = std::vector<T, b2gHugeAllocator<T>>;

struct b2GpuSolver
{
	int device = 0;
	int smCount = 0;
	int gridBlocks = 0;
	int mode = 0;
	bool cooperative = false;
	cudaStream_t stream = nullptr;
	cudaEvent_t evStart = nullptr, evStop = nullptr, evUpload = nullptr;

	// device: one input arena (mirror of the host wire staging), one output arena, the SoA solver state
	DeviceBuffer<float4> wireAll, outAll, vel, pos, bodyK, cf;
	DeviceBuffer<float> angDamp;
	DeviceBuffer<int2> cidx;
	DeviceBuffer<int> cmeta;
	ControlBlock* control = nullptr;
	bool controlClean = false; // zeroed behind the previous step's download: b2gEnqueueRun need not

	// The bins' lists from one step to the next (one block per bin with flat lists, or a cluster per bin).  What b2gScatterKernel builds -- which bodies
	// and contacts every bin holds -- only depends on the bodies' bins, the contacts' slots and their body indices.  In a
	// steady scene none of that changes: no contact or joint travels in full (each sits at the home, with the bodies, it had before),
	// the layout and the plan are the same, every body is in the bin it was in.  The step then runs on the lists
	// the previous step left in device memory: no scatter kernel (many_pyramids: 6 of 58 us).  B2GPU_KEEP_LISTS=0 turns it off.
	bool keepListsEnabled = true;
	bool listsValid = false; // the device's lists and counters are those of the step described by listsOf
	struct ListsOf
	{
		int binCount, capBodies, capContacts, capJoints, bodyCount, contactSlots, colorCount, jointCount, jointWords;
		int clusterSize, ownerLists, listCount, clusterRun, listCapContacts, flatLists;
		b2g::ColorRange colors[b2g::kMaxColors];
		b2g::ColorRange overflow;
		const void* buffers[11];
	} listsOf = {};
	b2gHugeVector<int> prevBins;		   // bin of every awake body in the previous island-mode step
	int prevBinCount = 0;
	std::atomic<int> binsChanged{ 0 }; // pack pass: some body is in another bin than in the previous step
	int listsReused = 0;			   // statistics: steps that ran without the scatter kernel


	// island mode scratch (b2g_island.cuh)
	DeviceBuffer<int> binCounters; // [binBodyCount | binColorStart | binJointStart | binFail], zeroed every run
	DeviceBuffer<int> bodyLocal, binBodyList, slotGroupBits, binContactList, binJointList;
	DeviceBuffer<int2> binJointBodies;
	DeviceBuffer<int4> binContactInfo;
	DeviceBuffer<float4> jointWork;
	double islandHeadRoom = 1.3; // bins are sized for this many times the average bytes per bin
	int headRoomCooldown = 0;	 // steps to wait after a failure before lowering it again
	int countersBinCount = 0, countersListCount = 0;
	int ownerListsOff = 0;		 // steps during which owner lists stay off after a block's share did not fit
	bool ownerListsEnabled = true; // B2GPU_OWNER_LISTS=0 turns them off
	size_t downloadQuads = 32 * 1024; // chunk of the pipelined download (B2GPU_DOWNLOAD_KIB)
	bool gridJointCacheEnabled = true; // B2GPU_GRID_JOINT_CACHE=0: the grid-barrier kernel solves its joints in the global working copy
	int gridJointCacheMax = 0;		   // joints per block that fit the kernel's shared memory
	bool dependentLaunch = true;   // B2GPU_PDL=0: the island kernel is launched after the scatter kernel has drained
	bool leveliseEnabled = true;   // B2GPU_LEVELISE=0: jointless bins keep the reference's colours as their stages
	bool flatListsEnabled = true;  // B2GPU_FLAT_LISTS=0: two-phase partition kernel for one block per bin too
	bool countersClean = false; // the bin counters are all zero (the island kernels zero what they have read)
	DeviceBuffer<int2> contactBinRank, jointBinRank;
	DeviceBuffer<int> planStart, planJoints; // the bins' plans (b2g::StepParams::planWrite)
	DeviceBuffer<int4> planInfo;
	bool planValid = false; // ... written by the step that built the lists the device holds
	std::vector<int> islandBin;	 // host: bin of every awake island
	std::vector<int> islandBodies; // host: bodies per island
	std::vector<int> islandContacts, islandJoints; // host: with exact island sizes (b2GpuStepDesc::islandSizes)
	bool islandSizesExact = false;
	double exactHeadRoom = 1.05; // like islandHeadRoom, for bins packed by their real size
	double binSqueeze = 1.0;	 // extra bins (factor) the last plan needed before its fullest bin fit, see b2gPlanBins
	int binSqueezeAge = 0;
	size_t binCounterCount = 0;
	size_t islandSmemBytes = 0;
	size_t islandSmemBudget = 0;
	// B2GPU_SPILL_JOINTS=1: allow a cluster plan with the joint records left in global memory.  Off by default: measured
	// on joint_grid it loses to the grid-barrier kernel (0.47 vs 0.25 ms) -- 16 SMs cannot pull 19 800 records of 256 B
	// per stage through L2 as fast as 148 SMs can
	bool spillJointsEnabled = false;
	bool spillJointsForced = false; // B2GPU_SPILL_JOINTS=2 (testing): steps with joints take that plan first
	bool resolveContacts = true; // diagnostics: B2GPU_RESOLVE=0 makes the island kernels chase head -> bodyLocal themselves
	int stageAllThreads = false;
	bool testTightBins = false;			 // testing: B2GPU_TEST_TIGHT_BINS=1 makes every island step fail over to the grid kernel
	int clusterForce = 0;				 // testing: smallest cluster size the planner may use (B2GPU_CLUSTER_FORCE)
	int clusterBins[4] = { 0, 0, 0, 0 }; // resident clusters of 2, 4, 8, 16 blocks (0 = not available)
	int overflowContacts = 0, overflowJoints = 0; // overflow colour totals of the step (all worlds)
	bool islandMode = false;
	int islandsEnabled = 1;
	int maxSharedOptin = 0;

	// resident mode (single worlds): see b2gShadowContact
	bool liteJointsEnabled = true; // B2GPU_LITE_JOINTS=0: joints always keep their 256-byte records
	bool liteJointsSeen = false;  // every joint of the previous step was a plain revolute joint (pack pass)
	bool planLiteJoints = false;  // this step's plan counts on that
	std::atomic<int> heavyJoint{ 0 }; // pack pass: some joint of this step is not a plain revolute joint
	bool residentEnabled = true; // B2GPU_RESIDENT=0: every step uploads everything (plain wire)
	bool resident = false;		 // this step
	bool cacheValid = false;	 // the shadows describe what the device holds (false: the next resident step sends everything)
	bool cacheUsable = false;	 // ... and this step's pack pass may rely on them
	int parity = 0;				 // which of the double-buffered arrays this step WRITES (outAll, residentStates)
	b2gHugeVector<b2gShadowContact> shadowContacts; // by home
	b2gHugeVector<b2gShadowImpulses> shadowImpulses; // by home
	b2gHugeVector<b2gShadowHead> shadowHeads;		   // by home
	int homeBase[kHomeColors + 1] = { 0 };		   // first home of every graph colour (persistent layout with spare room)
	int homeCount[kHomeColors] = { 0 };			   // contacts the colour had in the previous resident step
	int homeSlot[kHomeColors] = { 0 };			   // ... and the slot its array started at
	int segHome[kHomeColors] = { 0 };			   // this step: home colour of every contact segment
	b2gHugeVector<float4> shadowStates, shadowBody;  // by awake index: what residentStates[in] / residentBody hold
	int shadowBodyCount = 0;
	DeviceBuffer<float4> table, residentStates[2], residentBody, outOther, fullStream, dirtyStream;
	PinnedBuffer<float4> hFull, hDirty;
	std::atomic<int> fullCursor{ 0 }, dirtyCursor{ 0 }; // records handed out in the two streams (whole chunks)
	int fullCapacity = 0, dirtyCapacity = 0;
	size_t prevOutImpulses = 0; // where the previous step's impulse records start in ITS output arena (quads)
	int homeTotal = 0;
	// the same for joints: homes by (graph colour, index in the colour's joint array), shadows = the complete 256-byte record
	// the device's table holds (the solver's outputs written in by the unpack pass)
	b2gHugeVector<uint8_t> shadowJoints; // [joint homes * kJointStride]
	int jointHomeBase[kHomeColors + 1] = { 0 }, jointHomeCount[kHomeColors] = { 0 }, jointHomeSlot[kHomeColors] = { 0 }, jointSegHome[kHomeColors] = { 0 };
	int jointHomeTotal = 0;
	DeviceBuffer<float4> jointTable, jointAssembled, fullJointStream;
	PinnedBuffer<float4> hFullJoints;
	std::atomic<int> fullJointCursor{ 0 };
	int fullJointCapacity = 0, fullJointSent = 0;
	std::atomic<int> fullJointCount{ 0 };
	size_t prevOutJoints = 0; // where the previous step's joint records start in ITS output arena (quads)
	int jointWireQuads = b2g::kJointStride / 16; // quads per joint in the input arena: 16, or kLightJointQuads
	int wireQuads = b2g::WR_COUNT; // quads per contact slot in the input arena: WR_COUNT, or 1 (light records)
	int fullSent = 0, dirtySent = 0;
	std::atomic<int> streamOverflow{ 0 };
	std::atomic<int> fullCount{ 0 }, dirtyCount{ 0 }, vouchedCount{ 0 }; // records actually written / contacts taken on the caller's word (statistics)

	// page-locked staging owned by the library.  The input staging is written with non-temporal stores: on the
	// target hosts a DMA read of lines that sit dirty in several cores' caches runs at ~6 GB/s instead of ~53 GB/s
	// (tools/microbench/h2d_bench.cu, profiles/).
	PinnedBuffer<float4> hWire;
	PinnedBuffer<float4> hOut;

	// Deferred contact impulses (resident mode, the phased entry points; see include/b2_gpu_solver.h).  The step's impulse
	// records come back LAST and behind the caller's back: EndStep returns once the body states and the joints' outputs are
	// unpacked, the records stay in the page-locked output arena (two arenas, swapped every step) and are written into a
	// manifold only when somebody is going to read it (b2GpuSolverMaterializeContacts: the narrow phase before it
	// re-evaluates a manifold, the pack pass before it reads a contact it has no word on, the caller's flush).
	bool deferEnabled = false;
	bool defer = false; // this step
	PinnedBuffer<float4> hOutOther; // the previous step's output arena
	bool deferPending = false;		// hOutOther holds records that some manifolds have not received
	uint32_t deferStamp = 0;		// of the pending step (0: never)
	uint32_t deferNewStamp = 1;		// of the step in flight
	b2gHugeVector<uint32_t> consumedStamp; // by home
	// ... and the joints' output records likewise: found by the joint's home, identified by the joint id in the home's shadow
	b2gHugeVector<uint32_t> consumedJointStamp; // by joint home
	bool deferJointsPending = false;
	bool jointHomesOrdered = true;
	const float* pendingJointRecords = nullptr; // B2L_JOINT_OUT_FLOATS per joint of the pending step, by its place among the step's joints
	bool homesOrdered = true;		// the homes' keys are the callers' graph colour indices (they came in ascending order)
	const float* pendingRecords = nullptr; // the pending step's impulse records, by wire slot
	const float* prevRecords = nullptr;	   // the previous resident step's (what the device warm-starts clean contacts from)
	cudaEvent_t evRecords = nullptr;	   // behind the last chunk of the download
	std::atomic<int> recordsSynced{ 1 };
	int deferWaitChunks = 0; // chunks of the download the unpack pass waits for (the rest holds the impulse records)
	std::atomic<int> materialized{ 0 }; // statistics: records written into manifolds since the last step began

	// Direct outputs (with deferred impulses; B2GPU_DIRECT_OUT=0 turns them off): the kernels store the body states straight
	// into the page-locked output arena (mapped into the device's address space: the stores travel over PCIe while other
	// blocks are still solving) and the step's last kernel copies the control block the same way and raises a flag the host
	// polls -- no copy-engine round trips between the end of the kernels and the host's finalize pass.
	bool directEnabled = true;
	bool direct = false; // this step
	int* hFlag = nullptr; // page-locked, written by b2gSignalKernel
	int flagSerial = 0;
	unsigned flagPolls = 0;
	size_t directEnd = 0; // quads at the front of the output arena that are there when the flag is (the body states)
	ControlBlock* hControl = nullptr;

	// arena layouts, in float4 units
	size_t inStates = 0, inBody = 0, inWire = 0, inJoints = 0, inBins = 0, inMass = 0, inTotal = 0;
	bool checkMasses = true; // this step's pack pass compares the contacts' masses with the bodies'
	std::atomic<int> massMismatch{ 0 }; // pack pass: some contact's masses differ from its bodies' -> the mass region is uploaded
	// the copies that end a step's upload (the last piece of the arena, the streams of full records) go out in ONE driver
	// call (cudaMemcpyBatchAsync): each call costs 5 - 10 us on the calling thread and these are on the critical path
	bool batching = false;
	bool batchEnabled = true; // B2GPU_BATCH_COPIES=0
	std::vector<void*> batchDst, batchSrc;
	std::vector<size_t> batchBytes;
	bool uploadStarted = false;		// evUpload recorded (first copy of the step)
	bool arenaSent = false;			// the whole input arena has been enqueued for upload
	std::vector<uint8_t> blockSent; // pack blocks whose part of the input arena has been enqueued for upload (b2gPumpUploads)
	int sendScan = 0;				// every block before this one has been sent
	int blockItems = 512;			// items per block of the pack / unpack passes of this step (b2gBegin)
	size_t sendThreshold = 0;		// a run of packed blocks goes out when it is this long (quads)
	size_t outStates = 0, outImpulses = 0, outJoints = 0, outBits = 0, outTotal = 0;

	// the step in flight
	std::vector<b2GpuStepDesc> descs;
	b2GpuStepResult* results = nullptr; // one per world, or NULL
	std::vector<b2gBodySeg> bodySegs;
	std::vector<b2gContactSeg> contactSegs;
	std::vector<b2gJointSeg> jointSegs;
	std::vector<int> bodyStart, contactStart, jointStart; // item prefix sums, one more entry than segments
	std::vector<int> binBodies, binContacts, binJoints;
	b2g::StepParams params;
	int jointTotal = 0;
	int contactTotal = 0;
	bool begun = false;
	bool uploaded = false;
	bool ran = false;

	uint64_t launchCount = 0;
	uint64_t lastH2D = 0;
	uint64_t lastD2H = 0;
	int lastLaunches = 0;
	float lastKernelMs = 0.0f;
	std::chrono::steady_clock::time_point tBegin, tSubmit, tWaited;

	// pipelined host passes (b2GpuSolverPackWork / b2GpuSolverUnpackWork): the items are dealt out in blocks, claimed in
	// increasing order; one caller (the pump) moves the finished prefix over PCIe while the others keep packing, and
	// publishes how much of the output arena has arrived while the others unpack behind it
	std::atomic<int> workNext{ 0 };
	// the CUDA-event time of the kernels is read by whoever runs out of unpack blocks first (b2GpuSolverUnpackWork)
	std::atomic<int> kernelsSeen{ 0 }, timerClaim{ 0 }, timerDone{ 0 };
	int workBlocks = 0;
	int workItems = 0;
	std::unique_ptr<std::atomic<unsigned char>[]> workDone;
	size_t workDoneCapacity = 0;
	int pumpPrefix = 0; // blocks [0, pumpPrefix) are packed (owned by whoever holds pumpBusy)
	std::atomic<int> pumpBusy{ 0 }; // a thread is enqueueing uploads
	std::atomic<size_t> arrivedQuads{ 0 };
	std::atomic<int> workFailed{ 0 };
	std::vector<cudaEvent_t> chunkEvents;
	std::vector<size_t> chunkEnd;
	int chunkCount = 0;
	int chunkNext = 0; // pump only
	cudaEvent_t evControl = nullptr;
	bool trace = false; // B2GPU_TRACE=1: print the timeline of the pipelined transfers at EndStep (stderr)
	std::vector<std::pair<float, size_t>> traceSends, traceArrivals, tracePump;
	float traceBegun = 0.0f, traceSubmit = 0.0f, traceControl = 0.0f;
	float traceMarks[8] = { 0 };
	bool controlSeen = false;
};

inline int b2gRoundUp32( int n )
{
	return ( n + 31 ) & ~31;
}

// ---- the step as segments ---------------------------------------------------------------------------------------------
// One step solves `worldCount` independent worlds (1 for b2GpuSolverStep, N for the batch API).  Their arrays are
// addressed through segments: a body segment per world, a contact / joint segment per (colour slot, world).  Colour
// slot c holds every world's c-th ACTIVE colour (the stage order only matters inside a world, and inside a world the
// active colours are visited in ascending order, src/solver.c:1341-1367), the last slot is the overflow colour.
// Item order for pack/unpack: all bodies world by world, all contacts in segment (= slot) order, all joints.
inline int b2gFindSegment( const std::vector<int>& starts, int flat )
{
	// starts has segmentCount + 1 entries; returns the segment that contains `flat`
	int lo = 0, hi = (int)starts.size() - 1;
	while ( hi - lo > 1 )
	{
		int mid = ( lo + hi ) >> 1;
		if ( starts[mid] <= flat )
		{
			lo = mid;
		}
		else
		{
			hi = mid;
		}
	}
	return lo;
}

// ---- blocks of host work --------------------------------------------------------------------------------------------
constexpr int kWorkBlockItems = 512; // largest block of the host passes
constexpr size_t kTransferQuads = 16 * 1024;	  // 256 KiB: the first piece of a pipelined upload; the pieces double up to
constexpr size_t kTransferQuadsMax = 256 * 1024; // 4 MiB
constexpr size_t kDownloadQuads = 32 * 1024; // 512 KiB: the unpack pass runs this far behind the download

// b2g_wire.cu
int b2gSendArena( b2GpuSolver* s, size_t uptoQuads );
void b2gFlushLines( const void* ptr, size_t bytes );
// b2g_wire.cu
int b2gMaterializePendingFromSegs( b2GpuSolver* s, b2GpuStepResult* results );
// b2g_solver.cu
int b2gPollControl( b2GpuSolver* s, bool* seen );
int b2gDeferSync( b2GpuSolver* s );
int b2gEnqueueDownload( b2GpuSolver* s );
int b2gRerunIfIslandsFailed( b2GpuSolver* s, bool download );

