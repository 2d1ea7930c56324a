// b2g_solver.cu -- host side of the device library: the C-ABI of include/b2_gpu_solver.h.
//
// Layout of a step (segments, arenas), the island plan, the wire packing / unpacking passes with their pipelined PCIe
// transfers, and the launches.  The kernels live in b2g_island.cuh (partition + one block per bin), b2g_cluster.cuh
// (one thread-block cluster per bin) and b2g_grid.cuh (grid-barrier kernels); DESIGN.md has the map.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
#include "b2g_host.h"

#include "b2g_cluster.cuh"
#include "b2g_grid.cuh"
#include "b2g_resident.cuh"

#include <cuda_runtime.h>

#include <immintrin.h>
#if defined( __x86_64__ )
#include <cpuid.h>
#endif

#include <atomic>
#include <chrono>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>


// =================================================================================================================
// Host side
// =================================================================================================================

// The text of the last failure: of the calling thread if it had one, else of whichever thread failed last (the host passes
// run on the caller's worker threads, the thread that asks is usually not the one that hit the error).
static thread_local std::string t_lastError;
static std::mutex g_errorMutex;
static std::string g_lastError;
static thread_local std::string t_errorReply;

static void b2gRemember( const std::string& text )
{
	t_lastError = text;
	std::lock_guard<std::mutex> lock( g_errorMutex );
	g_lastError = text;
}

int b2gFail( const char* what, cudaError_t err )
{
	b2gRemember( std::string( what ) + ": " + cudaGetErrorString( err ) );
	return 1;
}

int b2gFailMsg( const char* what )
{
	b2gRemember( what );
	return 1;
}

extern "C" int b2GpuGetVersion( void )
{
	return 101;
}

extern "C" const char* b2GpuGetLastError( void )
{
	if ( !t_lastError.empty() )
	{
		return t_lastError.c_str();
	}
	std::lock_guard<std::mutex> lock( g_errorMutex );
	t_errorReply = g_lastError;
	return t_errorReply.c_str();
}

extern "C" int b2GpuGetDeviceCount( void )
{
	int count = 0;
	if ( cudaGetDeviceCount( &count ) != cudaSuccess )
	{
		cudaGetLastError();
		return 0;
	}
	return count;
}

extern "C" b2GpuSolver* b2GpuSolverCreate( int device )
{
	int count = 0;
	cudaError_t err = cudaGetDeviceCount( &count );
	if ( err != cudaSuccess || count == 0 )
	{
		b2gFail( "b2GpuSolverCreate: no CUDA device (there is no CPU fallback)", err );
		cudaGetLastError();
		return nullptr;
	}
	if ( device < 0 || device >= count )
	{
		b2gFailMsg( "b2GpuSolverCreate: bad device index" );
		return nullptr;
	}
	if ( ( err = cudaSetDevice( device ) ) != cudaSuccess )
	{
		b2gFail( "cudaSetDevice", err );
		return nullptr;
	}

	b2GpuSolver* s = new b2GpuSolver();
	s->device = device;
	cudaDeviceProp prop;
	if ( ( err = cudaGetDeviceProperties( &prop, device ) ) != cudaSuccess )
	{
		b2gFail( "cudaGetDeviceProperties", err );
		delete s;
		return nullptr;
	}
	s->smCount = prop.multiProcessorCount;
	s->cooperative = prop.cooperativeLaunch != 0;

	int blocksPerSm = 0;
	err = cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSm, b2g::b2gStepKernel, b2g::kBlockThreads, 0 );
	if ( err != cudaSuccess || blocksPerSm < 1 )
	{
		b2gFail( "step kernel cannot be resident (built for sm_100a only)", err );
		delete s;
		return nullptr;
	}
	{
		// the grid-barrier kernel keeps its blocks' joints in shared memory (b2g_stages.cuh, JointCache)
		const char* cacheEnv = getenv( "B2GPU_GRID_JOINT_CACHE" );
		s->gridJointCacheEnabled = cacheEnv == nullptr || atoi( cacheEnv ) != 0;
		int bytes = (int)prop.sharedMemPerBlockOptin - 8 * 1024;
		bytes = bytes > 192 * 1024 ? 192 * 1024 : bytes;
		if ( bytes < 0 || cudaFuncSetAttribute( b2g::b2gStepKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes ) != cudaSuccess )
		{
			cudaGetLastError();
			bytes = 0;
		}
		s->gridJointCacheMax = bytes / b2g::kJointStride;
	}
	// one persistent block per SM: the grid barrier needs every block co-resident
	s->gridBlocks = s->smCount;
	const char* gridEnv = getenv( "B2GPU_GRID" );
	if ( gridEnv != nullptr && atoi( gridEnv ) > 0 && atoi( gridEnv ) <= s->smCount * blocksPerSm )
	{
		s->gridBlocks = atoi( gridEnv );
	}

	s->maxSharedOptin = (int)prop.sharedMemPerBlockOptin;
	const char* islandEnv = getenv( "B2GPU_ISLANDS" );
	s->islandsEnabled = islandEnv != nullptr ? atoi( islandEnv ) : 1;
	{
		cudaFuncAttributes attr;
		int dynamicMax = 0;
		cudaFuncAttributes clusterAttr;
		if ( cudaFuncGetAttributes( &attr, b2g::b2gIslandKernel ) == cudaSuccess &&
			 cudaFuncGetAttributes( &clusterAttr, b2g::b2gClusterIslandKernel ) == cudaSuccess )
		{
			// one budget for both island kernels: the planner does not care which of them runs
			size_t staticBytes = attr.sharedSizeBytes > clusterAttr.sharedSizeBytes ? attr.sharedSizeBytes : clusterAttr.sharedSizeBytes;
			dynamicMax = s->maxSharedOptin - (int)staticBytes - 256;
		}
		if ( dynamicMax <= 0 ||
			 cudaFuncSetAttribute( b2g::b2gIslandKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dynamicMax ) != cudaSuccess )
		{
			cudaGetLastError();
			s->islandsEnabled = 0;
			dynamicMax = 0;
		}
		s->islandSmemBudget = (size_t)dynamicMax;
		// clusters of 2, 4, 8, 16 blocks with the same carve-up: how many can be resident at once
		const char* traceEnv = getenv( "B2GPU_TRACE" );
		s->trace = traceEnv != nullptr && atoi( traceEnv ) != 0;
		const char* spillEnv = getenv( "B2GPU_SPILL_JOINTS" );
		s->spillJointsEnabled = spillEnv != nullptr && atoi( spillEnv ) != 0;
		s->spillJointsForced = spillEnv != nullptr && atoi( spillEnv ) == 2;
		const char* ownerEnv = getenv( "B2GPU_OWNER_LISTS" );
		s->ownerListsEnabled = ownerEnv == nullptr || atoi( ownerEnv ) != 0;
		// (Tried: write-combined memory for the upload staging buffer -- the device-clock upload time of many_pyramids stayed
		// at ~192 us, the link is the limit, not snoops.)
		const char* chunkEnv = getenv( "B2GPU_DOWNLOAD_KIB" );
		s->downloadQuads = chunkEnv != nullptr && atoi( chunkEnv ) >= 16 ? (size_t)atoi( chunkEnv ) * 64 : kDownloadQuads;
		const char* batchEnv = getenv( "B2GPU_BATCH_COPIES" );
		s->batchEnabled = batchEnv == nullptr || atoi( batchEnv ) != 0;
		const char* keepEnv = getenv( "B2GPU_KEEP_LISTS" );
		s->keepListsEnabled = keepEnv == nullptr || atoi( keepEnv ) != 0;
		const char* directEnv = getenv( "B2GPU_DIRECT_OUT" );
		s->directEnabled = directEnv == nullptr || atoi( directEnv ) != 0;
		const char* liteEnv = getenv( "B2GPU_LITE_JOINTS" );
		s->liteJointsEnabled = liteEnv == nullptr || atoi( liteEnv ) != 0;
		const char* residentEnv = getenv( "B2GPU_RESIDENT" );
		s->residentEnabled = residentEnv == nullptr || atoi( residentEnv ) != 0;
		const char* pdlEnv = getenv( "B2GPU_PDL" );
		s->dependentLaunch = pdlEnv == nullptr || atoi( pdlEnv ) != 0;
		const char* levelEnv = getenv( "B2GPU_LEVELISE" );
		s->leveliseEnabled = levelEnv == nullptr || atoi( levelEnv ) != 0;
		const char* flatEnv = getenv( "B2GPU_FLAT_LISTS" );
		s->flatListsEnabled = flatEnv == nullptr || atoi( flatEnv ) != 0;
		const char* resolveEnv = getenv( "B2GPU_RESOLVE" );
		s->resolveContacts = resolveEnv == nullptr || atoi( resolveEnv ) != 0;
		const char* stageEnv = getenv( "B2GPU_STAGE_ALL" );
		s->stageAllThreads = stageEnv != nullptr ? atoi( stageEnv ) : 0;
		const char* tightEnv = getenv( "B2GPU_TEST_TIGHT_BINS" );
		s->testTightBins = tightEnv != nullptr && atoi( tightEnv ) != 0;
		const char* forceEnv = getenv( "B2GPU_CLUSTER_FORCE" );
		s->clusterForce = forceEnv != nullptr ? atoi( forceEnv ) : 0;
		const char* clusterEnv = getenv( "B2GPU_CLUSTERS" );
		bool clustersEnabled = s->islandsEnabled != 0 && ( clusterEnv == nullptr || atoi( clusterEnv ) != 0 );
		if ( clustersEnabled && dynamicMax > 0 &&
			 cudaFuncSetAttribute( b2g::b2gClusterIslandKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dynamicMax ) == cudaSuccess &&
			 cudaFuncSetAttribute( b2g::b2gClusterIslandKernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1 ) == cudaSuccess )
		{
			for ( int k = 0; k < 4; ++k )
			{
				cudaLaunchConfig_t config = {};
				config.gridDim = dim3( (unsigned)( ( 2 << k ) * s->smCount ) );
				config.blockDim = dim3( b2g::kIslandThreads );
				config.dynamicSmemBytes = (size_t)dynamicMax;
				cudaLaunchAttribute attribute;
				attribute.id = cudaLaunchAttributeClusterDimension;
				attribute.val.clusterDim.x = (unsigned)( 2 << k );
				attribute.val.clusterDim.y = 1;
				attribute.val.clusterDim.z = 1;
				config.attrs = &attribute;
				config.numAttrs = 1;
				int clusters = 0;
				if ( cudaOccupancyMaxActiveClusters( &clusters, b2g::b2gClusterIslandKernel, &config ) == cudaSuccess )
				{
					s->clusterBins[k] = clusters;
				}
			}
		}
		cudaGetLastError();
	}

	bool ok = cudaStreamCreateWithFlags( &s->stream, cudaStreamNonBlocking ) == cudaSuccess;
	ok = ok && cudaEventCreate( &s->evStart ) == cudaSuccess;
	ok = ok && cudaEventCreate( &s->evStop ) == cudaSuccess;
	ok = ok && cudaEventCreate( &s->evUpload ) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags( &s->evControl, cudaEventDisableTiming ) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags( &s->evRecords, cudaEventDisableTiming ) == cudaSuccess;
	ok = ok && cudaMalloc( &s->control, sizeof( ControlBlock ) ) == cudaSuccess;
	ok = ok && cudaHostAlloc( &s->hControl, sizeof( ControlBlock ), cudaHostAllocDefault ) == cudaSuccess;
	ok = ok && cudaHostAlloc( &s->hFlag, 64, cudaHostAllocDefault ) == cudaSuccess;
	if ( ok )
	{
		*s->hFlag = 0;
	}
	if ( !ok )
	{
		b2gFail( "b2GpuSolverCreate: resource allocation", cudaGetLastError() );
		b2GpuSolverDestroy( s );
		return nullptr;
	}
	memset( &s->params, 0, sizeof( s->params ) );
	return s;
}

extern "C" void b2GpuSolverDestroy( b2GpuSolver* s )
{
	if ( s == nullptr )
	{
		return;
	}
	cudaSetDevice( s->device );
	if ( s->stream != nullptr )
	{
		cudaStreamSynchronize( s->stream );
	}
	s->wireAll.release();
	s->outAll.release();
	s->vel.release();
	s->pos.release();
	s->bodyK.release();
	s->cf.release();
	s->angDamp.release();
	s->cidx.release();
	s->cmeta.release();
	s->hWire.release();
	s->hOut.release();
	s->hOutOther.release();
	s->table.release();
	s->residentStates[0].release();
	s->residentStates[1].release();
	s->residentBody.release();
	s->outOther.release();
	s->fullStream.release();
	s->dirtyStream.release();
	s->hFull.release();
	s->hDirty.release();
	s->jointTable.release();
	s->jointAssembled.release();
	s->fullJointStream.release();
	s->hFullJoints.release();
	s->binCounters.release();
	s->bodyLocal.release();
	s->binBodyList.release();
	s->slotGroupBits.release();
	s->binContactList.release();
	s->binContactInfo.release();
	s->planInfo.release();
	s->planJoints.release();
	s->planStart.release();
	s->jointWork.release();
	s->binJointList.release();
	s->binJointBodies.release();
	s->contactBinRank.release();
	s->jointBinRank.release();
	if ( s->control != nullptr )
	{
		cudaFree( s->control );
	}
	if ( s->hControl != nullptr )
	{
		cudaFreeHost( s->hControl );
	}
	if ( s->hFlag != nullptr )
	{
		cudaFreeHost( s->hFlag );
	}
	for ( cudaEvent_t ev : { s->evStart, s->evStop, s->evUpload, s->evControl, s->evRecords } )
	{
		if ( ev != nullptr )
		{
			cudaEventDestroy( ev );
		}
	}
	for ( cudaEvent_t ev : s->chunkEvents )
	{
		cudaEventDestroy( ev );
	}
	if ( s->stream != nullptr )
	{
		cudaStreamDestroy( s->stream );
	}
	delete s;
}

extern "C" int b2GpuSolverSetMode( b2GpuSolver* s, int mode )
{
	if ( s == nullptr || mode < 0 || mode > 1 )
	{
		return b2gFailMsg( "b2GpuSolverSetMode: bad argument" );
	}
	s->mode = mode;
	return 0;
}

extern "C" uint64_t b2GpuSolverGetLaunchCount( const b2GpuSolver* s )
{
	return s != nullptr ? s->launchCount : 0;
}

extern "C" int b2GpuSolverGetIslandPlan( const b2GpuSolver* s, int* binCount, int* blocksPerBin )
{
	int bins = s != nullptr && s->islandMode ? s->params.binCount : 0;
	if ( binCount != nullptr )
	{
		*binCount = bins;
	}
	if ( blocksPerBin != nullptr )
	{
		*blocksPerBin = bins > 0 ? s->params.clusterSize : 0;
	}
	return bins;
}

// steps of this solver that ran on the previous step's bin lists (no scatter kernel)
extern "C" int b2GpuSolverGetListReuseCount( const b2GpuSolver* s )
{
	return s != nullptr ? s->listsReused : 0;
}

extern "C" int b2GpuSolverGetResidentStats( const b2GpuSolver* s, int* fullContacts, int* dirtyBodies, int* vouchedContacts, int* fullJoints )
{
	if ( s == nullptr || !s->resident )
	{
		return 0;
	}
	if ( fullContacts != nullptr )
	{
		*fullContacts = s->fullCount.load( std::memory_order_relaxed );
	}
	if ( dirtyBodies != nullptr )
	{
		*dirtyBodies = s->dirtyCount.load( std::memory_order_relaxed );
	}
	if ( vouchedContacts != nullptr )
	{
		*vouchedContacts = s->vouchedCount.load( std::memory_order_relaxed );
	}
	if ( fullJoints != nullptr )
	{
		*fullJoints = s->fullJointCount.load( std::memory_order_relaxed );
	}
	return 1;
}

struct b2gBinPlan
{
	int binCount, share, capB, capC, capJ;
	bool spillJoints;
};

// Bins for blocks-per-bin = share.  Island i goes to the bin its first body falls in when the islands are laid end to
// end and cut every `target` bodies, so a bin gets between target - (largest island) and target + (largest island)
// bodies.  `waves`: more bins than `binLimit` are allowed when the data does not fit (they run in waves).
static bool b2gPlanBinsOnce( b2GpuSolver* s, int islandCount, int share, int binLimitIn, bool waves, bool spillJoints, double squeeze,
							  b2gBinPlan* plan, bool* hopeless = nullptr )
{
	const b2g::StepParams& P = s->params;
	const int bodies = P.bodyCount;
	size_t budget = s->islandSmemBudget;
	const double bytesPerBody = 52.0, bytesPerContact = b2g::CF_COUNT * 16.0 + 8.0 + 8.0;
	const bool lite = s->planLiteJoints && !spillJoints && share > 1; // (the cluster kernel only: b2gIslandKernel keeps the full records)
	const double bytesPerJoint = ( spillJoints ? 0.0 : lite ? (double)b2g::kLiteJointWords * 4.0 : (double)b2g::kJointStride ) + 4.0;
	const bool exact = s->islandSizesExact;
	// bytes of shared memory an island needs: from its real size, or -- no sizes given -- from its bodies and the step's
	// average constraint density
	const double densityC = bodies > 0 ? (double)s->contactTotal / bodies : 0.0, densityJ = bodies > 0 ? (double)s->jointTotal / bodies : 0.0;
	auto islandBytes = [&]( int i ) -> double {
		double nb = s->islandBodies[(size_t)i];
		double nc = exact ? (double)s->islandContacts[(size_t)i] : nb * densityC;
		double nj = exact ? (double)s->islandJoints[(size_t)i] : nb * densityJ;
		return nb * bytesPerBody + nc * bytesPerContact + nj * bytesPerJoint;
	};
	double totalBytes = 0.0, largest = 0.0;
	for ( int i = 0; i < islandCount; ++i )
	{
		double bytes = islandBytes( i );
		totalBytes += bytes;
		largest = bytes > largest ? bytes : largest;
	}
	if ( largest > (double)budget * share )
	{
		// an island that is too big for a bin whatever the number of bins
		if ( hopeless != nullptr )
		{
			*hopeless = true;
		}
		return false;
	}
	int binLimit = binLimitIn < islandCount ? binLimitIn : islandCount;
	// head room between the average bin and the capacity (adaptive: raised when a bin did not fit, lowered slowly while
	// all is well): generous when the constraint density of the islands is only an estimate, small when the sizes are real
	double headRoom = ( exact ? s->exactHeadRoom : ( share > 1 ? 1.1 : s->islandHeadRoom ) ) * squeeze;
	int wanted = (int)( totalBytes * headRoom / ( (double)budget * share ) ) + 1;
	if ( wanted > binLimit )
	{
		if ( !waves )
		{
			return false;
		}
		// The bins run in waves of one block per SM and a wave takes the same time however full its blocks are (the step
		// is a chain of latency-bound stages): a few bins more than a whole number of waves cost a whole wave.  Give up
		// some of the head room if that saves one.
		int perWave = s->smCount;
		int waveCount = ( wanted + perWave - 1 ) / perWave;
		int tight = (int)( totalBytes * ( exact ? 1.02 : 1.05 ) * squeeze / (double)budget ) + 1;
		if ( waveCount > 1 && ( waveCount - 1 ) * perWave >= tight )
		{
			wanted = ( waveCount - 1 ) * perWave;
		}
		binLimit = wanted < islandCount ? wanted : islandCount;
	}
	// islands laid end to end and cut every `target` bytes
	double target = totalBytes / binLimit;
	target = target > 1.0 ? target : 1.0;
	s->islandBin.assign( (size_t)islandCount, 0 );
	std::vector<int>& binBodies = s->binBodies;
	binBodies.assign( (size_t)binLimit, 0 );
	s->binContacts.assign( (size_t)binLimit, 0 );
	s->binJoints.assign( (size_t)binLimit, 0 );
	int bin = 0, maxBin = 0, maxBinContacts = 0, maxBinJoints = 0;
	double before = 0.0;
	for ( int i = 0; i < islandCount; ++i )
	{
		int b = (int)( before / target );
		b = b < binLimit ? b : binLimit - 1;
		s->islandBin[i] = b;
		binBodies[b] += s->islandBodies[i];
		maxBin = binBodies[b] > maxBin ? binBodies[b] : maxBin;
		if ( exact )
		{
			s->binContacts[b] += s->islandContacts[i];
			s->binJoints[b] += s->islandJoints[i];
			maxBinContacts = s->binContacts[b] > maxBinContacts ? s->binContacts[b] : maxBinContacts;
			maxBinJoints = s->binJoints[b] > maxBinJoints ? s->binJoints[b] : maxBinJoints;
		}
		bin = b > bin ? b : bin;
		before += islandBytes( i );
	}
	int binCount = bin + 1;

	// capacities PER BLOCK: exact for bodies (a power of two per block when the bin is shared by a cluster); the rest
	// of the budget is split between contacts and joints in proportion to their estimated bytes, so a bin may hold
	// several times its fair share before binFail trips
	int capB = ( maxBin + 3 ) & ~3;
	if ( share > 1 )
	{
		// equal runs of the bin's bodies (a multiple of 4 keeps the carve-up 16-byte aligned); the owner of a body is
		// found with a multiply-high, exact below 65536 bodies per bin
		capB = ( ( maxBin + share - 1 ) / share + 3 ) & ~3;
		if ( maxBin >= 65536 )
		{
			return false;
		}
	}
	double fraction = (double)maxBin / (double)bodies / (double)share;
	// a block of a cluster holds ceil(n / share) of every colour, the first block the bin's overflow colour on top
	double slack = share > 1 ? (double)( b2g::kMaxColors + 1 ) : 0.0;
	double binC = exact ? (double)maxBinContacts / share : fraction * s->contactTotal;
	double binJ = exact ? (double)maxBinJoints / share : fraction * s->jointTotal;
	double needC = binC + slack + ( share > 1 ? (double)s->overflowContacts : 0.0 );
	double needJ = binJ + ( s->jointTotal > 0 ? slack : 0.0 ) + ( share > 1 ? (double)s->overflowJoints : 0.0 );
	needC = needC < (double)s->contactTotal ? needC : (double)s->contactTotal;
	needJ = needJ < (double)s->jointTotal ? needJ : (double)s->jointTotal;
	size_t fixed = b2g::islandSharedBytes( capB, 0, 0, !spillJoints, lite );
	const double margin = exact ? 1.02 : 1.1;
	if ( fixed + (size_t)( needC * margin * bytesPerContact + needJ * margin * bytesPerJoint ) + 4096 > budget )
	{
		return false;
	}
	double weightC = needC * bytesPerContact + 2048.0, weightJ = s->jointTotal > 0 ? needJ * bytesPerJoint + 2048.0 : 0.0;
	double spare = (double)( budget - fixed ) - 64.0;
	int capC = ( (int)( spare * weightC / ( weightC + weightJ ) / bytesPerContact ) ) & ~3;
	int capJ = s->jointTotal > 0 ? ( (int)( spare * weightJ / ( weightC + weightJ ) / bytesPerJoint ) ) & ~3 : 0;
	if ( binCount > s->smCount )
	{
		// many small bins (batches of worlds): do not hog the SM, several blocks should be co-resident
		int tightC = ( (int)( needC * 2.0 ) + 35 ) & ~3, tightJ = s->jointTotal > 0 ? ( (int)( needJ * 2.0 ) + 11 ) & ~3 : 0;
		capC = capC < tightC ? capC : tightC;
		capJ = capJ < tightJ ? capJ : tightJ;
	}
	// no point in exceeding what exists
	capC = capC > ( ( s->contactTotal + 3 ) & ~3 ) ? ( ( s->contactTotal + 3 ) & ~3 ) : capC;
	capJ = capJ > ( ( s->jointTotal + 3 ) & ~3 ) ? ( ( s->jointTotal + 3 ) & ~3 ) : capJ;
	capC = capC < 4 ? 4 : capC;
	for ( int guard = 0; guard < 64 && b2g::islandSharedBytes( capB, capC, capJ, !spillJoints, lite ) > budget; ++guard )
	{
		capC = ( capC - capC / 32 - 4 ) & ~3; // rounding slack: shave ~3 % until it fits
		capJ = capJ > 0 ? ( capJ - capJ / 32 - 4 ) & ~3 : 0;
		capC = capC < 4 ? 4 : capC;
		capJ = capJ < 0 ? 0 : capJ;
	}
	if ( b2g::islandSharedBytes( capB, capC, capJ, !spillJoints, lite ) > budget || capC < needC || capJ < needJ )
	{
		return false;
	}
	*plan = { binCount, share, capB, capC, capJ, spillJoints };
	return true;
}

// A cut every `target` bytes can leave a bin one island above the average; when the bins run in waves (a batch of worlds:
// islands that are a sizeable fraction of a block) the fullest bin may then miss the budget by a little.  More bins cure
// that: retry with a few per cent more until the plan fits.
static bool b2gPlanBins( b2GpuSolver* s, int islandCount, int share, int binLimitIn, bool waves, bool spillJoints, b2gBinPlan* plan )
{
	if ( !waves )
	{
		return b2gPlanBinsOnce( s, islandCount, share, binLimitIn, false, spillJoints, 1.0, plan );
	}
	// start from what worked last step; now and then try one notch less
	s->binSqueezeAge += 1;
	if ( s->binSqueezeAge >= 64 )
	{
		s->binSqueezeAge = 0;
		s->binSqueeze = s->binSqueeze / 1.01 > 1.0 ? s->binSqueeze / 1.01 : 1.0;
	}
	double squeeze = s->binSqueeze;
	for ( int attempt = 0; attempt < 40 && squeeze < 1.5; ++attempt )
	{
		bool hopeless = false;
		if ( b2gPlanBinsOnce( s, islandCount, share, binLimitIn, true, spillJoints, squeeze, plan, &hopeless ) )
		{
			s->binSqueeze = squeeze;
			return true;
		}
		if ( hopeless )
		{
			return false;
		}
		squeeze *= 1.01;
	}
	return false;
}

// ---- island mode planning (host) --------------------------------------------------------------------------------
// Pack the awake islands of all worlds into bins balanced by body count and size the island kernel's shared memory
// carve-up.  Island mode is used when every world brings the hint and the estimated bins fit; the device double-checks
// the exact sizes (binFail -> the grid-barrier kernel takes the step).
static int b2gPlanIslands( b2GpuSolver* s )
{
	b2g::StepParams& P = s->params;
	s->islandMode = false;
	P.binCount = 0;
	P.liteJoints = 0;
	// Plain revolute joints are kept as 27 words instead of 64 by the island / cluster kernels when EVERY joint of the step is
	// one (b2g_joint.cuh).  Whether that holds is only known once the pack pass has looked at the joints, so the plan goes by
	// what the previous step saw; the scatter / partition kernel checks, and a wrong guess costs a rerun on the grid kernel.
	s->planLiteJoints = s->liteJointsEnabled && s->liteJointsSeen && s->jointTotal > 0;
	int bodies = P.bodyCount;
	if ( s->islandsEnabled == 0 || s->mode != 0 || bodies == 0 )
	{
		return 0;
	}

	int islandCount = 0;
	for ( b2gBodySeg& seg : s->bodySegs )
	{
		if ( seg.islands == nullptr || seg.islandCount <= 0 )
		{
			return 0;
		}
		seg.islandBase = islandCount;
		islandCount += seg.islandCount;
	}
	s->islandBodies.assign( (size_t)islandCount, 0 );
	s->islandSizesExact = true;
	for ( const b2gBodySeg& seg : s->bodySegs )
	{
		s->islandSizesExact = s->islandSizesExact && seg.islandSizes != nullptr;
	}
	if ( s->islandSizesExact )
	{
		// the caller knows its islands (the seam reads b2Island::bodies/contacts/joints.count): nothing to count here
		s->islandContacts.assign( (size_t)islandCount, 0 );
		s->islandJoints.assign( (size_t)islandCount, 0 );
		for ( const b2gBodySeg& seg : s->bodySegs )
		{
			for ( int i = 0; i < seg.islandCount; ++i )
			{
				const b2GpuIslandSize& size = seg.islandSizes[i];
				if ( size.bodyCount < 0 || size.contactCount < 0 || size.jointCount < 0 )
				{
					return 0;
				}
				s->islandBodies[(size_t)( seg.islandBase + i )] = size.bodyCount;
				s->islandContacts[(size_t)( seg.islandBase + i )] = size.contactCount;
				s->islandJoints[(size_t)( seg.islandBase + i )] = size.jointCount;
			}
		}
	}
	for ( const b2gBodySeg& seg : s->bodySegs )
	{
		if ( s->islandSizesExact )
		{
			break;
		}
		if ( seg.islandCount == 1 )
		{
			// a world that is one island (the worlds of a batch usually are): nothing to count, nothing to look up
			s->islandBodies[(size_t)seg.islandBase] = seg.count;
			continue;
		}
		for ( int i = 0; i < seg.count; ++i )
		{
			int island = seg.islands[i];
			if ( island < 0 || island >= seg.islandCount )
			{
				return 0; // a body without an island: no partition guarantee, use the grid-barrier kernel
			}
			s->islandBodies[seg.islandBase + island] += 1;
		}
	}

	// One block per bin when that fits; otherwise clusters of 2..16 blocks per bin (b2g_cluster.cuh), the smallest that
	// holds the largest bin.  If nothing fits the grid-barrier kernel takes the step.
	if ( s->headRoomCooldown > 0 )
	{
		s->headRoomCooldown -= 1;
	}
	else
	{
		s->islandHeadRoom = s->islandHeadRoom * 0.995 > 1.2 ? s->islandHeadRoom * 0.995 : 1.2;
		s->exactHeadRoom = s->exactHeadRoom * 0.998 > 1.05 ? s->exactHeadRoom * 0.998 : 1.05;
	}
	b2gBinPlan plan;
	const bool residentFirst = !( s->spillJointsForced && s->jointTotal > 0 );
	bool planned = residentFirst && s->clusterForce <= 1 && b2gPlanBins( s, islandCount, 1, s->smCount, true, false, &plan );
	for ( int k = 0; !planned && residentFirst && k < 4; ++k )
	{
		int share = 2 << k;
		if ( s->clusterBins[k] > 0 && share >= s->clusterForce )
		{
			planned = b2gPlanBins( s, islandCount, share, s->clusterBins[k], false, false, &plan );
		}
	}
	// a jointed island too big for that (joint_grid: 19 800 joints x 256 B): the largest cluster with the joint records
	// left in global memory -- the bodies, which is what the blocks share, still live in distributed shared memory
	for ( int k = 3; !planned && k >= 0 && s->jointTotal > 0 && s->spillJointsEnabled; --k )
	{
		int share = 2 << k;
		if ( s->clusterBins[k] > 0 && share >= s->clusterForce )
		{
			planned = b2gPlanBins( s, islandCount, share, s->clusterBins[k], false, true, &plan );
		}
	}
	if ( !planned )
	{
		return 0;
	}
	if ( s->testTightBins )
	{
		plan.capC = 4; // testing: no bin fits, every step takes the rerun path
	}
	const int binCount = plan.binCount, capB = plan.capB, capC = plan.capC, capJ = plan.capJ;

	// Owner lists (b2g_island.cuh): one bin shared by a cluster, constraints go to the block that owns their first body.
	// The blocks' shares are then as uneven as the scene; if one does not fit, the step is rerun and even dealing is used
	// for a while.
	if ( s->ownerListsOff > 0 )
	{
		s->ownerListsOff -= 1;
	}
	const bool ownerLists = plan.share > 1 && binCount == 1 && s->ownerListsOff == 0 && s->ownerListsEnabled;
	const int listCount = ownerLists ? plan.share : binCount;
	size_t slots = (size_t)P.contactSlots;
	// zeroed every step: [binBodyCount][contact counts per list and colour][joint counts][binFail][bin-wide totals]
	s->binCounterCount = (size_t)binCount + (size_t)listCount * 2 * b2g::kColorSlots + 1 + 2 * b2g::kColorSlots;
	{
		const int* before = s->binCounters.ptr;
		B2G_CUDA( s->binCounters.reserve( s->binCounterCount + (size_t)listCount * 2 * b2g::kColorSlots ) );
		if ( s->binCounters.ptr != before || binCount != s->countersBinCount || listCount != s->countersListCount )
		{
			s->countersClean = false; // fresh memory, or the tables move (the offset tables behind them are not zero)
		}
		s->countersBinCount = binCount;
		s->countersListCount = listCount;
	}
	B2G_CUDA( s->bodyLocal.reserve( (size_t)bodies + 1 ) );
	const size_t share = (size_t)plan.share;
	B2G_CUDA( s->binBodyList.reserve( (size_t)binCount * capB * share + 1 ) );
	B2G_CUDA( s->slotGroupBits.reserve( slots + 1 ) );
	B2G_CUDA( s->contactBinRank.reserve( slots + 1 ) );
	B2G_CUDA( s->binContactList.reserve( (size_t)binCount * capC * share + 1 ) );
	B2G_CUDA( s->binContactInfo.reserve( (size_t)binCount * capC * share + 1 ) );
	if ( plan.share == 1 )
	{
		B2G_CUDA( s->planInfo.reserve( (size_t)binCount * capC + 1 ) );
		B2G_CUDA( s->planJoints.reserve( (size_t)binCount * capJ + 1 ) );
		B2G_CUDA( s->planStart.reserve( (size_t)binCount * 2 * b2g::kColorSlots + 1 ) );
	}
	B2G_CUDA( s->jointBinRank.reserve( (size_t)s->jointTotal + 1 ) );
	B2G_CUDA( s->binJointList.reserve( (size_t)binCount * capJ * share + 1 ) );
	B2G_CUDA( s->binJointBodies.reserve( (size_t)binCount * capJ * share + 1 ) );

	P.binCount = binCount;
	P.capBodies = capB;
	P.capContacts = capC;
	P.capJoints = capJ;
	P.clusterSize = plan.share;
	P.stageAllThreads = s->stageAllThreads;
	P.resolveContacts = s->resolveContacts ? 1 : 0;
	P.clusterRun = plan.share > 1 ? plan.capB : 0;
	P.clusterMagic = plan.share > 1 ? (unsigned)( ( ( 1ull << 32 ) + (unsigned)plan.capB - 1ull ) / (unsigned)plan.capB ) : 0u;
	P.binCapBodies = capB * plan.share;
	P.binCapContacts = capC * plan.share;
	P.binCapJoints = capJ * plan.share;
	P.bodyBin = reinterpret_cast<const int*>( s->wireAll.ptr + s->inBins );
	P.bodyLocal = s->bodyLocal.ptr;
	P.ownerLists = ownerLists ? 1 : 0;
	P.jointsSpilled = plan.spillJoints ? 1 : 0;
	P.listCount = listCount;
	P.leveliseContacts = s->leveliseEnabled ? 1 : 0;
	P.flatLists = plan.share == 1 && s->flatListsEnabled && s->resolveContacts && s->jointTotal < ( 1 << b2g::kFlatJointShift ) ? 1 : 0;
	P.listCapContacts = ownerLists ? capC : capC * plan.share;
	P.listCapJoints = ownerLists ? capJ : capJ * plan.share;
	P.binBodyCount = s->binCounters.ptr;
	P.binColorStart = s->binCounters.ptr + binCount;
	P.binJointStart = P.binColorStart + (size_t)listCount * b2g::kColorSlots;
	P.binFail = P.binJointStart + (size_t)listCount * b2g::kColorSlots;
	P.binColorTotal = P.binFail + 1;
	P.binColorOffset = P.binColorTotal + 2 * b2g::kColorSlots;
	P.binJointOffset = P.binColorOffset + (size_t)listCount * b2g::kColorSlots;
	P.binBodyList = s->binBodyList.ptr;
	P.contactBinRank = s->contactBinRank.ptr;
	P.slotGroupBits = s->slotGroupBits.ptr;
	P.binContactList = s->binContactList.ptr;
	P.binContactInfo = s->binContactInfo.ptr;
	P.planInfo = s->planInfo.ptr;
	P.planJoints = s->planJoints.ptr;
	P.planStart = s->planStart.ptr;
	P.jointBinRank = s->jointBinRank.ptr;
	P.binJointList = s->binJointList.ptr;
	P.binJointBodies = s->binJointBodies.ptr;
	const bool liteJoints = s->planLiteJoints && !plan.spillJoints && plan.share > 1;
	s->islandSmemBytes = b2g::islandSharedBytes( capB, capC, capJ, !plan.spillJoints, liteJoints );
	P.liteJoints = liteJoints ? 1 : 0;
	s->islandMode = true;
	return 0;
}

static int b2gBlocksFor( const b2GpuSolver* s, int itemCount )
{
	return ( itemCount + s->blockItems - 1 ) / s->blockItems;
}

static void b2gResetWork( b2GpuSolver* s, int itemCount, int blockCount )
{
	s->workItems = itemCount;
	s->workBlocks = blockCount;
	if ( (size_t)s->workBlocks > s->workDoneCapacity )
	{
		s->workDoneCapacity = (size_t)s->workBlocks * 2 + 64;
		s->workDone.reset( new std::atomic<unsigned char>[s->workDoneCapacity] );
	}
	for ( int i = 0; i < s->workBlocks; ++i )
	{
		s->workDone[i].store( 0, std::memory_order_relaxed );
	}
	s->pumpPrefix = 0;
	s->sendScan = 0;
	s->sendThreshold = kTransferQuads;
	s->blockSent.assign( (size_t)blockCount, 0 );
	s->workFailed.store( 0, std::memory_order_relaxed );
	s->kernelsSeen.store( 0, std::memory_order_relaxed );
	s->timerClaim.store( 0, std::memory_order_relaxed );
	s->timerDone.store( 0, std::memory_order_relaxed );
	s->workNext.store( 0, std::memory_order_release );
}

// ---- phase 1: layout --------------------------------------------------------------------------------------------
static bool b2gSameStepParams( const b2GpuStepDesc& a, const b2GpuStepDesc& b )
{
	return a.dt == b.dt && a.inv_dt == b.inv_dt && a.h == b.h && a.inv_h == b.inv_h && a.subStepCount == b.subStepCount &&
		   memcmp( &a.contactSoftness, &b.contactSoftness, sizeof( a.contactSoftness ) ) == 0 &&
		   memcmp( &a.staticSoftness, &b.staticSoftness, sizeof( a.staticSoftness ) ) == 0 &&
		   a.restitutionThreshold == b.restitutionThreshold && a.maxLinearVelocity == b.maxLinearVelocity &&
		   a.gravity[0] == b.gravity[0] && a.gravity[1] == b.gravity[1] && a.contactSpeed == b.contactSpeed &&
		   a.contactHertz == b.contactHertz && a.contactDampingRatio == b.contactDampingRatio &&
		   a.hitEventThreshold == b.hitEventThreshold && a.lengthUnitsPerMeter == b.lengthUnitsPerMeter &&
		   a.enableWarmStarting == b.enableWarmStarting && a.enableContactSoftening == b.enableContactSoftening;
}

#define B2G_MARK( i ) s->traceMarks[i] = std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count()
static int b2gBegin( b2GpuSolver* s, const b2GpuStepDesc* descs, int worldCount, b2GpuStepResult* results )
{
	if ( s == nullptr || descs == nullptr || worldCount <= 0 )
	{
		return b2gFailMsg( "b2GpuSolverBeginStep: bad argument" );
	}
	B2G_CUDA( cudaSetDevice( s->device ) );
	s->tBegin = std::chrono::steady_clock::now();
	s->begun = false;
	s->uploaded = false;
	s->ran = false;
	s->results = results;
	s->heavyJoint.store( 0, std::memory_order_relaxed );
	const b2GpuStepDesc* d = descs;

	int maxColors = 0;
	for ( int w = 0; w < worldCount; ++w )
	{
		const b2GpuStepDesc& dw = descs[w];
		if ( dw.activeColorCount < 0 || dw.activeColorCount > b2g::kMaxColors || dw.awakeBodyCount < 0 || dw.subStepCount < 0 )
		{
			return b2gFailMsg( "b2GpuSolverBeginStep: bad descriptor" );
		}
		if ( w > 0 && !b2gSameStepParams( descs[0], dw ) )
		{
			return b2gFailMsg( "b2GpuSolverStepBatch: all worlds of a batch must share the step parameters" );
		}
		maxColors = dw.activeColorCount > maxColors ? dw.activeColorCount : maxColors;
	}

	B2G_MARK( 0 );
	b2g::StepParams& P = s->params;
	memset( &P, 0, sizeof( P ) );
	P.dt = d->dt;
	P.inv_dt = d->inv_dt;
	P.h = d->h;
	P.inv_h = d->inv_h;
	P.subStepCount = d->subStepCount;
	P.contactSoft = { d->contactSoftness.biasRate, d->contactSoftness.massScale, d->contactSoftness.impulseScale };
	P.staticSoft = { d->staticSoftness.biasRate, d->staticSoftness.massScale, d->staticSoftness.impulseScale };
	P.restitutionThreshold = d->restitutionThreshold;
	P.maxLinearVelocity = d->maxLinearVelocity;
	P.gravityX = d->gravity[0];
	P.gravityY = d->gravity[1];
	P.contactSpeed = d->contactSpeed;
	P.contactHertz = d->contactHertz;
	P.contactDampingRatio = d->contactDampingRatio;
	P.hitEventThreshold = d->hitEventThreshold;
	P.lengthUnitsPerMeter = d->lengthUnitsPerMeter;
	P.enableWarmStarting = d->enableWarmStarting;
	P.enableSoftening = d->enableContactSoftening;
	P.colorCount = maxColors;

	// body segments
	s->bodySegs.clear();
	s->bodyStart.assign( 1, 0 );
	int bodies = 0, jointBitWords = 0;
	for ( int w = 0; w < worldCount; ++w )
	{
		const b2GpuStepDesc& dw = descs[w];
		b2gBodySeg seg;
		seg.states = static_cast<uint8_t*>( dw.states );
		seg.sims = static_cast<const uint8_t*>( dw.sims );
		seg.islands = dw.bodyIsland;
		seg.islandSizes = dw.islandSizes;
		seg.islandCount = dw.islandCount;
		seg.islandBase = 0;
		seg.count = dw.awakeBodyCount;
		seg.base = bodies;
		seg.jointBitBase = jointBitWords * 32;
		seg.jointWords = 2 * ( ( dw.jointIdCapacity + 63 ) / 64 );
		s->bodySegs.push_back( seg );
		bodies += dw.awakeBodyCount;
		jointBitWords += seg.jointWords;
		s->bodyStart.push_back( bodies );
	}
	P.bodyCount = bodies;
	P.jointWords = jointBitWords;

	// constraint segments in slot order; every colour slot starts on a multiple of 32 slots, every segment on a multiple
	// of 4 (the SIMD groups of the reference are per colour array); the gaps are dead slots (pointCount 0)
	B2G_MARK( 1 );
	s->contactStart.assign( 1, 0 );
	s->jointStart.assign( 1, 0 );
	int slotCount = maxColors + 1;
	// two light passes over the descriptors (they are large and there may be thousands): count the segments of every
	// colour slot, then place each segment at its final position -- slot by slot, world by world inside a slot
	int contactSegStart[b2g::kMaxColors + 2] = { 0 }, jointSegStart[b2g::kMaxColors + 2] = { 0 };
	for ( int w = 0; w < worldCount; ++w )
	{
		const b2GpuStepDesc& dw = descs[w];
		for ( int c = 0; c <= dw.activeColorCount; ++c )
		{
			bool isOverflow = c == dw.activeColorCount;
			const b2GpuColorDesc& color = isOverflow ? dw.overflow : dw.colors[c];
			int slotIndex = isOverflow ? slotCount - 1 : c;
			if ( color.contactCount < 0 || color.jointCount < 0 )
			{
				return b2gFailMsg( "b2GpuSolverBeginStep: negative count" );
			}
			contactSegStart[slotIndex + 1] += color.contactCount > 0 ? 1 : 0;
			jointSegStart[slotIndex + 1] += color.jointCount > 0 ? 1 : 0;
		}
	}
	for ( int c = 0; c < slotCount; ++c )
	{
		contactSegStart[c + 1] += contactSegStart[c];
		jointSegStart[c + 1] += jointSegStart[c];
	}
	s->contactSegs.resize( (size_t)contactSegStart[slotCount] );
	s->jointSegs.resize( (size_t)jointSegStart[slotCount] );
	{
		int contactCursor[b2g::kMaxColors + 2], jointCursor[b2g::kMaxColors + 2];
		memcpy( contactCursor, contactSegStart, sizeof( contactCursor ) );
		memcpy( jointCursor, jointSegStart, sizeof( jointCursor ) );
		for ( int w = 0; w < worldCount; ++w )
		{
			const b2GpuStepDesc& dw = descs[w];
			for ( int c = 0; c <= dw.activeColorCount; ++c )
			{
				bool isOverflow = c == dw.activeColorCount;
				const b2GpuColorDesc& color = isOverflow ? dw.overflow : dw.colors[c];
				int slotIndex = isOverflow ? slotCount - 1 : c;
				if ( color.contactCount > 0 )
				{
					b2gContactSeg& seg = s->contactSegs[(size_t)contactCursor[slotIndex]++];
					seg.sims = static_cast<uint8_t*>( color.contactSims );
					seg.count = color.contactCount;
					seg.slotStart = 0;
					seg.world = w;
					seg.wide = !isOverflow;
					seg.colorIndex = color.colorIndex;
					seg.hints = dw.recycled != nullptr && dw.recycledCount[c] > 0 ? dw.recycled + dw.recycledStart[c] : nullptr;
					seg.hintCount = seg.hints != nullptr ? dw.recycledCount[c] : 0;
					seg.hintStamp = dw.recycledStamp;
					seg.hintsInPlace = seg.hints != nullptr && dw.recycledInPlace[c] != 0;
				}
				if ( color.jointCount > 0 )
				{
					b2gJointSeg& seg = s->jointSegs[(size_t)jointCursor[slotIndex]++];
					seg.sims = static_cast<uint8_t*>( color.jointSims );
					seg.count = color.jointCount;
					seg.jointStart = 0;
					seg.world = w;
					seg.colorIndex = color.colorIndex;
					seg.overflow = isOverflow;
				}
			}
		}
	}
	s->contactStart.resize( s->contactSegs.size() + 1 );
	s->jointStart.resize( s->jointSegs.size() + 1 );
	int slot = 0, joint = 0, flat = 0;
	for ( int c = 0; c < slotCount; ++c )
	{
		bool isOverflow = c + 1 == slotCount;
		b2g::ColorRange& range = isOverflow ? P.overflow : P.colors[c];
		range.contactStart = slot;
		range.jointStart = joint;
		for ( int k = contactSegStart[c]; k < contactSegStart[c + 1]; ++k )
		{
			b2gContactSeg& seg = s->contactSegs[(size_t)k];
			seg.slotStart = ( slot + 3 ) & ~3;
			slot = seg.slotStart + seg.count;
			flat += seg.count;
			s->contactStart[(size_t)k + 1] = flat;
		}
		for ( int k = jointSegStart[c]; k < jointSegStart[c + 1]; ++k )
		{
			b2gJointSeg& seg = s->jointSegs[(size_t)k];
			seg.jointStart = joint;
			joint += seg.count;
			s->jointStart[(size_t)k + 1] = joint;
		}
		range.contactCount = slot - range.contactStart;
		range.jointCount = joint - range.jointStart;
		slot = b2gRoundUp32( slot );
	}
	B2G_MARK( 2 );
	s->contactTotal = flat;
	s->jointTotal = joint;
	s->overflowContacts = P.overflow.contactCount;
	s->overflowJoints = P.overflow.jointCount;
	P.contactSlots = slot;
	P.jointCount = joint;
	{
		// joints the fullest block (block 0) of the grid-barrier kernel would cache: whole chunks of 32 per colour
		int slots = 0;
		for ( int c = 0; c < P.colorCount; ++c )
		{
			int chunks = ( P.colors[c].jointCount + 31 ) / 32;
			slots += 32 * ( ( chunks + s->gridBlocks - 1 ) / s->gridBlocks );
		}
		P.gridJointCache = s->gridJointCacheEnabled ? ( slots < s->gridJointCacheMax ? slots : s->gridJointCacheMax ) : 0;
		P.gridJointsAllCached = P.gridJointCache > 0 && P.gridJointCache >= slots ? 1 : 0;
	}

	size_t nb = (size_t)bodies;
	const size_t jointQuads = b2g::kJointStride / 16;
	// Resident mode (single worlds, b2g_types.cuh): the arena carries a 16-byte light record per slot instead of the
	// 96-byte wire record and no body regions; full records and dirty bodies travel in two streams of their own.
	s->resident = s->residentEnabled && worldCount == 1 && slot < b2g::kLightIdMask / 4;
	if ( !s->resident )
	{
		s->cacheValid = false; // a plain step reuses the arenas the resident copies live in
	}
	s->wireQuads = s->resident ? 1 : b2g::WR_COUNT;
	s->jointWireQuads = s->resident ? b2g::kLightJointQuads : (int)jointQuads;
	const size_t bodyQuads = s->resident ? 0 : 2;
	// input arena: [contacts 6/slot or 1/slot][joints 16/joint][states 2/body][packed sims 2/body][bins 1/4 body].  The
	// pipelined pack pass (b2GpuSolverPackWork) works through it in address order -- constraints first, the three body
	// regions last -- so that a finished prefix of its blocks is a prefix of the arena and the upload can start with the
	// first megabyte of contacts.
	s->inWire = 0;
	s->inJoints = s->inWire + (size_t)s->wireQuads * slot;
	s->inStates = s->inJoints + (size_t)s->jointWireQuads * joint;
	s->inBody = s->inStates + bodyQuads * nb;
	s->inBins = s->inBody + bodyQuads * nb;
	s->inMass = s->inBins + ( nb + 3 ) / 4; // optional tail: one quad per contact slot, see b2g::WireRow
	s->inTotal = s->inMass + slot;
	// A batch packs colour slot by colour slot across all worlds: looking up the bodies of every contact there would stream
	// the worlds' b2BodySim arrays through the host's caches once per colour (measured: 8192 worlds, +12 % on a step that is
	// bound by host memory bandwidth).  The check is for single worlds, whose body array stays cached; batches upload the
	// masses as before.
	s->checkMasses = s->bodySegs.size() == 1;
	s->massMismatch.store( s->checkMasses ? 0 : 1, std::memory_order_relaxed );
	s->uploadStarted = false;
	s->arenaSent = false;
	// output arena: [states 2/body][impulse records][joint impulse records 3/joint][joint event bits]; with deferred
	// impulses the records come last: [states][joints][bits][impulse records], and the step ends without them
	size_t impulseQuads = ( (size_t)slot * b2g::kImpulseFloats + 3 ) / 4;
	s->defer = s->deferEnabled && s->resident;
	s->outStates = 0;
	if ( s->defer )
	{
		// (what the step waits for first: the states and the joint-event bits; then the records nobody waits for)
		s->outBits = s->outStates + 2 * nb;
		s->outJoints = s->outBits + ( (size_t)P.jointWords + 3 ) / 4;
		s->outImpulses = s->outJoints + (size_t)( B2L_JOINT_OUT_FLOATS / 4 ) * joint;
		s->outTotal = s->outImpulses + impulseQuads;
	}
	else
	{
		s->outImpulses = s->outStates + 2 * nb;
		s->outJoints = s->outImpulses + impulseQuads;
		s->outBits = s->outJoints + (size_t)( B2L_JOINT_OUT_FLOATS / 4 ) * joint;
		s->outTotal = s->outBits + ( (size_t)P.jointWords + 3 ) / 4;
	}

	B2G_CUDA( s->wireAll.reserve( s->inTotal + 1 ) );
	B2G_CUDA( s->hWire.reserve( s->inTotal + 1 ) );
	B2G_CUDA( s->outAll.reserve( s->outTotal + 1 ) );
	B2G_CUDA( s->hOut.reserve( s->outTotal + 1 ) );
	B2G_CUDA( s->jointWork.reserve( jointQuads * joint + 1 ) );
	B2G_CUDA( s->vel.reserve( nb + 1 ) );
	B2G_CUDA( s->pos.reserve( nb + 1 ) );
	B2G_CUDA( s->bodyK.reserve( nb + 1 ) );
	B2G_CUDA( s->angDamp.reserve( nb + 1 ) );
	B2G_CUDA( s->cidx.reserve( (size_t)slot + 1 ) );
	B2G_CUDA( s->cmeta.reserve( (size_t)slot + 1 ) );
	// the SoA field stride follows cidx's capacity so that all per-slot arrays grow together
	size_t slotCapacity = s->cidx.capacity;
	B2G_CUDA( s->cf.reserve( slotCapacity * b2g::CF_COUNT ) );
	if ( s->resident )
	{
		// the persistent copies: growing one of them loses its contents, the step then sends everything
		auto persistent = [&]( DeviceBuffer<float4>& buffer, size_t count ) -> cudaError_t {
			if ( count > buffer.capacity )
			{
				s->cacheValid = false;
			}
			return buffer.reserve( count );
		};
		// Homes: every graph colour owns a region of the table whose place does not change from step to step, so that a
		// contact keeps its home as long as it keeps its place in its colour's array.  The regions have spare room; when a
		// colour outgrows its region everything is laid out again (and sent again).  The key of a segment is its graph
		// colour index when the descriptor's indices are what the reference produces (ascending, overflow last), else its
		// ordinal.
		{
			bool ordered = true;
			int last = -1;
			for ( const b2gContactSeg& seg : s->contactSegs )
			{
				int index = seg.colorIndex;
				ordered = ordered && ( seg.wide ? index > last && index < kHomeColors - 1 : index == kHomeColors - 1 );
				last = seg.wide ? index : last;
			}
			bool fits = true;
			int ordinal = 0;
			for ( size_t k = 0; k < s->contactSegs.size(); ++k )
			{
				const b2gContactSeg& seg = s->contactSegs[k];
				int key = !seg.wide ? kHomeColors - 1 : ordered ? seg.colorIndex : ordinal++;
				s->segHome[k] = key;
				fits = fits && seg.count <= s->homeBase[key + 1] - s->homeBase[key];
			}
			if ( s->deferPending && ( !fits || ordered != s->homesOrdered || !s->deferEnabled ) )
			{
				// the homes are about to move (or to mean something else): whatever is pending is found by home, so it
				// goes into the manifolds now -- the step's own contact arrays say where every contact is
				if ( b2gMaterializePendingFromSegs( s, nullptr ) != 0 )
				{
					return 1;
				}
			}
			s->homesOrdered = ordered;
			if ( !fits )
			{
				int need[kHomeColors] = { 0 };
				for ( size_t k = 0; k < s->contactSegs.size(); ++k )
				{
					need[s->segHome[k]] = s->contactSegs[k].count;
				}
				int base = 0;
				for ( int key = 0; key < kHomeColors; ++key )
				{
					int have = s->homeBase[key + 1] - s->homeBase[key]; // (of the old layout: read before it is overwritten below)
					int room = need[key] > have ? need[key] + need[key] / 2 + 64 : have;
					need[key] = room;
				}
				for ( int key = 0; key < kHomeColors; ++key )
				{
					s->homeBase[key] = base;
					base += need[key];
					s->homeCount[key] = 0;
				}
				s->homeBase[kHomeColors] = base;
				s->cacheValid = false;
			}
			s->homeTotal = s->homeBase[kHomeColors];
		}
		B2G_CUDA( persistent( s->table, (size_t)( s->homeTotal + 1 ) * b2g::kTableRows ) );
		// joints: the same layout of homes
		{
			bool ordered = true;
			int last = -1;
			for ( const b2gJointSeg& seg : s->jointSegs )
			{
				int index = seg.colorIndex;
				ordered = ordered && ( !seg.overflow ? index > last && index < kHomeColors - 1 : index == kHomeColors - 1 );
				last = !seg.overflow ? index : last;
			}
			bool fits = true;
			int ordinal = 0;
			for ( size_t k = 0; k < s->jointSegs.size(); ++k )
			{
				const b2gJointSeg& seg = s->jointSegs[k];
				int key = seg.overflow ? kHomeColors - 1 : ordered ? seg.colorIndex : ordinal++;
				s->jointSegHome[k] = key;
				fits = fits && seg.count <= s->jointHomeBase[key + 1] - s->jointHomeBase[key];
			}
			if ( s->deferJointsPending && ( !fits || ordered != s->jointHomesOrdered ) )
			{
				// (the joints' homes are about to move: see the contacts' homes above)
				if ( b2gMaterializePendingFromSegs( s, nullptr ) != 0 )
				{
					return 1;
				}
			}
			s->jointHomesOrdered = ordered;
			if ( !fits )
			{
				int need[kHomeColors] = { 0 };
				for ( size_t k = 0; k < s->jointSegs.size(); ++k )
				{
					need[s->jointSegHome[k]] = s->jointSegs[k].count;
				}
				for ( int key = 0; key < kHomeColors; ++key )
				{
					int have = s->jointHomeBase[key + 1] - s->jointHomeBase[key];
					need[key] = need[key] > have ? need[key] + need[key] / 2 + 32 : have;
				}
				int base = 0;
				for ( int key = 0; key < kHomeColors; ++key )
				{
					s->jointHomeBase[key] = base;
					base += need[key];
					s->jointHomeCount[key] = 0;
				}
				s->jointHomeBase[kHomeColors] = base;
				s->cacheValid = false;
			}
			s->jointHomeTotal = s->jointHomeBase[kHomeColors];
		}
		B2G_CUDA( persistent( s->jointTable, (size_t)( s->jointHomeTotal + 1 ) * jointQuads ) );
		B2G_CUDA( s->jointAssembled.reserve( jointQuads * joint + 1 ) );
		s->fullJointCapacity = ( joint + kStreamChunk ) & ~( kStreamChunk - 1 );
		B2G_CUDA( persistent( s->residentStates[0], 2 * nb + 2 ) );
		B2G_CUDA( persistent( s->residentStates[1], 2 * nb + 2 ) );
		B2G_CUDA( persistent( s->residentBody, 2 * nb + 2 ) );
		s->fullCapacity = ( slot + kStreamChunk ) & ~( kStreamChunk - 1 );
		s->dirtyCapacity = ( bodies + kStreamChunk ) & ~( kStreamChunk - 1 );
		// every pack block may leave one chunk partly used
		int packBlocks = ( bodies + s->contactTotal + s->jointTotal ) / 128 + 2;
		s->fullCapacity += packBlocks * kStreamChunk;
		s->dirtyCapacity += packBlocks * kStreamChunk;
		s->fullJointCapacity += packBlocks * kStreamChunk;
		B2G_CUDA( s->fullJointStream.reserve( (size_t)s->fullJointCapacity * jointQuads ) );
		B2G_CUDA( s->hFullJoints.reserve( (size_t)s->fullJointCapacity * jointQuads ) );
		s->fullJointCursor.store( 0, std::memory_order_relaxed );
		s->fullJointCount.store( 0, std::memory_order_relaxed );
		s->fullJointSent = 0;
		if ( s->shadowJoints.size() < (size_t)( s->jointHomeTotal + 1 ) * b2g::kJointStride )
		{
			s->shadowJoints.resize( (size_t)( s->jointHomeTotal + 1 ) * b2g::kJointStride, 0 );
		}
		B2G_CUDA( s->fullStream.reserve( (size_t)s->fullCapacity * b2g::WR_COUNT ) );
		B2G_CUDA( s->hFull.reserve( (size_t)s->fullCapacity * b2g::WR_COUNT ) );
		B2G_CUDA( s->dirtyStream.reserve( (size_t)s->dirtyCapacity * b2g::kDirtyBodyQuads ) );
		B2G_CUDA( s->hDirty.reserve( (size_t)s->dirtyCapacity * b2g::kDirtyBodyQuads ) );
		s->fullCursor.store( 0, std::memory_order_relaxed );
		s->dirtyCursor.store( 0, std::memory_order_relaxed );
		s->fullSent = s->dirtySent = 0;
		s->streamOverflow.store( 0, std::memory_order_relaxed );
		s->fullCount.store( 0, std::memory_order_relaxed );
		s->dirtyCount.store( 0, std::memory_order_relaxed );
		s->vouchedCount.store( 0, std::memory_order_relaxed );
		if ( s->shadowContacts.size() < (size_t)s->homeTotal + 1 )
		{
			s->shadowContacts.resize( (size_t)s->homeTotal + 1, b2gShadowContact{} );
			s->shadowImpulses.resize( (size_t)s->homeTotal + 1, b2gShadowImpulses{} );
			s->shadowHeads.resize( (size_t)s->homeTotal + 1, b2gShadowHead{ -1, -1, -1, 0 } );
		}
		if ( s->shadowStates.size() < 2 * nb + 2 )
		{
			s->shadowStates.resize( 2 * nb + 2 + nb );
			s->shadowBody.resize( 2 * nb + 2 + nb );
		}
		if ( !s->cacheValid )
		{
			s->shadowBodyCount = 0;
		}
	}
	// the shadows are trusted by this step's pack pass only; they count again once the step has ended (b2gEnd)
	s->cacheUsable = s->resident && s->cacheValid;
	s->cacheValid = false;
	if ( s->defer )
	{
		if ( s->consumedStamp.size() < (size_t)s->homeTotal + 1 )
		{
			s->consumedStamp.resize( (size_t)s->homeTotal + 1, 0 );
		}
		if ( s->consumedJointStamp.size() < (size_t)s->jointHomeTotal + 1 )
		{
			s->consumedJointStamp.resize( (size_t)s->jointHomeTotal + 1, 0 );
		}
		s->deferNewStamp += 1;
		if ( s->deferNewStamp == 0 )
		{
			// wrapped: no mark of the past may look current (nothing is pending with stamp 0)
			if ( s->deferPending && b2gMaterializePendingFromSegs( s, nullptr ) != 0 )
			{
				return 1;
			}
			s->consumedStamp.assign( s->consumedStamp.size(), 0 );
			s->consumedJointStamp.assign( s->consumedJointStamp.size(), 0 );
			s->deferNewStamp = 1;
		}
		// the pack pass reads the previous step's records (a contact it has no word on: are its impulses still the device's?)
		if ( b2gDeferSync( s ) != 0 )
		{
			return 1;
		}
		s->prevRecords = s->cacheUsable && s->hOutOther.ptr != nullptr ? reinterpret_cast<const float*>( s->hOutOther.ptr + s->prevOutImpulses ) : nullptr;
		s->materialized.store( 0, std::memory_order_relaxed );
	}
	else if ( s->deferPending )
	{
		// a step that does not defer (a batch, a plain step) after one that did: the manifolds it is about to read receive what
		// is owed to them, found by place in the step's own arrays (single worlds; a batch's arrays are other worlds')
		if ( worldCount != 1 )
		{
			return b2gFailMsg( "b2GpuSolverStepBatch: the solver's previous single-world step left impulses deferred "
							   "(b2GpuSolverMaterializeContacts, b2GpuSolverDeferredDone)" );
		}
		if ( b2gMaterializePendingFromSegs( s, nullptr ) != 0 )
		{
			return 1;
		}
	}

	P.rawStates = reinterpret_cast<const uint8_t*>( s->resident ? s->residentStates[0].ptr : s->wireAll.ptr + s->inStates );
	P.wireBody = s->resident ? s->residentBody.ptr : s->wireAll.ptr + s->inBody;
	P.wireMass = s->wireAll.ptr + s->inMass;
	P.massFromBodies = 0; // decided when the packing is done (b2gEnqueueUpload)
	P.wire = s->wireAll.ptr + s->inWire;
	P.light = s->resident ? s->wireAll.ptr + s->inWire : nullptr;
	P.table = s->table.ptr;
	P.full = s->fullStream.ptr;
	P.prevImpulses = s->outOther.ptr != nullptr ? reinterpret_cast<const float*>( s->outOther.ptr + s->prevOutImpulses ) : nullptr;
	P.residentOut = s->resident ? reinterpret_cast<uint8_t*>( s->residentStates[1].ptr ) : nullptr;
	P.residentStates = s->residentStates[0].ptr;
	P.residentBody = s->residentBody.ptr;
	P.dirtyBodies = s->dirtyStream.ptr;
	P.dirtyBodyCapacity = 0; // decided when the packing is done
	P.rawJoints = reinterpret_cast<const uint8_t*>( s->resident ? s->jointAssembled.ptr : s->wireAll.ptr + s->inJoints );
	P.lightJoints = s->resident ? s->wireAll.ptr + s->inJoints : nullptr;
	P.jointTable = s->jointTable.ptr;
	P.fullJoints = s->fullJointStream.ptr;
	P.prevOutJoints = s->outOther.ptr != nullptr ? reinterpret_cast<const float*>( s->outOther.ptr + s->prevOutJoints ) : nullptr;
	P.jointAssembled = s->jointAssembled.ptr;
	P.g.vel = s->vel.ptr;
	P.g.pos = s->pos.ptr;
	P.g.bodyK = s->bodyK.ptr;
	P.g.angDamp = s->angDamp.ptr;
	P.g.cf = s->cf.ptr;
	P.g.cfStride = (int)slotCapacity;
	P.g.cidx = s->cidx.ptr;
	P.g.cmeta = s->cmeta.ptr;
	P.g.joints = reinterpret_cast<uint8_t*>( s->jointWork.ptr ); // working copy of the joint records (grid kernel, spilled joints)
	P.outJoints = reinterpret_cast<float*>( s->outAll.ptr + s->outJoints );
	P.outStates = reinterpret_cast<uint8_t*>( s->outAll.ptr + s->outStates );
	// (the joint-event bits, set with atomics in device memory, follow by DMA; the joints' output records are deferred like
	// the contacts')
	s->direct = s->defer && s->directEnabled;
	s->directEnd = 0;
	if ( s->direct )
	{
		// (page-locked host memory has the same address on the device: unified addressing)
		P.outStates = reinterpret_cast<uint8_t*>( s->hOut.ptr + s->outStates );
		s->directEnd = s->outBits; // the states are the arena's first region
	}
	P.outImpulses = reinterpret_cast<float*>( s->outAll.ptr + s->outImpulses );
	P.jointBits = reinterpret_cast<uint32_t*>( s->outAll.ptr + s->outBits );
	P.hasHitEvents = &s->control->hasHitEvents;
	P.g.anyRestitution = &s->control->anyRestitution;
	P.islandFailed = &s->control->islandFailed;
	P.g.clusterRun = 0;
	P.g.clusterMagic = 0;
	P.g.asyncBar = 0;
	P.barrier = s->control->barrier;
	P.stageCycles = s->control->stageCycles;

	B2G_MARK( 3 );
	_mm_sfence();
	B2G_MARK( 4 );

	if ( b2gPlanIslands( s ) != 0 )
	{
		return 1;
	}

	// the bodies' bins of the previous step, for the pack pass to compare with (the bins' lists may serve again, b2gEnqueueRun)
	s->binsChanged.store( 0, std::memory_order_relaxed );
	if ( s->islandMode )
	{
		if ( s->prevBins.size() < (size_t)P.bodyCount + 1 )
		{
			s->prevBins.resize( (size_t)P.bodyCount + (size_t)P.bodyCount / 2 + 64, -1 );
		}
		if ( s->prevBinCount != P.bodyCount )
		{
			s->binsChanged.store( 1, std::memory_order_relaxed );
			s->prevBinCount = P.bodyCount;
		}
	}
	else
	{
		s->prevBinCount = 0;
		s->listsValid = false;
	}

	// pack blocks: the constraints' blocks, then the bodies' blocks (b2gPackBlockRange)
	{
		// blocks of the host passes: ~128 per step so that a small step still spreads evenly over the host's workers and its
		// first bytes go out early, at most 512 items (a block is claimed with one atomic and ends with a fence)
		int items = P.bodyCount + s->contactTotal + s->jointTotal;
		int perBlock = ( items / 128 + 63 ) & ~63;
		s->blockItems = perBlock < 128 ? 128 : perBlock > kWorkBlockItems ? kWorkBlockItems : perBlock;
	}
	b2gResetWork( s, P.bodyCount + s->contactTotal + s->jointTotal, b2gBlocksFor( s, s->contactTotal + s->jointTotal ) + b2gBlocksFor( s, P.bodyCount ) );
	s->traceBegun = std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count();
	s->begun = true;
	return 0;
}

extern "C" int b2GpuSolverBeginStep( b2GpuSolver* s, const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	return b2gBegin( s, d, 1, r );
}

// pack items: [bodies][contacts, slot order, overflow last][joints, same order]
extern "C" int b2GpuSolverGetPackItemCount( const b2GpuSolver* s )
{
	return s != nullptr && s->begun ? s->params.bodyCount + s->contactTotal + s->jointTotal : 0;
}

// unpack items: the same index space
extern "C" int b2GpuSolverGetUnpackItemCount( const b2GpuSolver* s )
{
	return b2GpuSolverGetPackItemCount( s );
}

static int b2gEnqueueUpload( b2GpuSolver* s )
{
	if ( b2gSendArena( s, s->inTotal ) != 0 )
	{
		return 1;
	}
	const bool massSent = s->massMismatch.load( std::memory_order_acquire ) != 0;
	s->params.massFromBodies = massSent ? 0 : 1;
	s->lastH2D = ( massSent ? s->inTotal : s->inMass ) * sizeof( float4 );
	if ( s->resident )
	{
		s->params.dirtyBodyCapacity = s->dirtySent;
		s->lastH2D += ( (size_t)s->fullSent * b2g::WR_COUNT + (size_t)s->dirtySent * b2g::kDirtyBodyQuads +
						(size_t)s->fullJointSent * ( b2g::kJointStride / 16 ) ) * sizeof( float4 );
	}
	s->uploaded = true;
	return 0;
}

// resident mode, after the step's kernels: the static rows of the step's full records go into the table (b2g_resident.cuh)
static int b2gEnqueueCommit( b2GpuSolver* s )
{
	if ( !s->resident || s->fullSent == 0 )
	{
		return 0;
	}
	int rows = s->params.contactSlots * b2g::kTableRows;
	int blocks = ( rows + 255 ) / 256;
	blocks = blocks < 1 ? 1 : blocks > s->smCount * 8 ? s->smCount * 8 : blocks;
	b2g::b2gCommitKernel<<<blocks, 256, 0, s->stream>>>( s->params );
	cudaError_t err = cudaGetLastError();
	if ( err != cudaSuccess )
	{
		return b2gFail( "b2gCommitKernel launch", err );
	}
	s->lastLaunches += 1;
	s->launchCount += 1;
	return 0;
}

static int b2gLaunchStage( b2GpuSolver* s, int op, int color, int blocks )
{
	b2g::b2gStageKernel<<<blocks, b2g::kBlockThreads, 0, s->stream>>>( s->params, op, color );
	s->lastLaunches += 1;
	cudaError_t err = cudaGetLastError();
	if ( err != cudaSuccess )
	{
		return b2gFail( "b2gStageKernel launch", err );
	}
	return 0;
}

static int b2gRunStages( b2GpuSolver* s )
{
	const b2g::StepParams& P = s->params;
	int grid = s->gridBlocks;
	bool hasOverflow = P.overflow.contactCount + P.overflow.jointCount > 0;
#define B2G_STAGE( op, c, blocks )                                                                                               \
	if ( b2gLaunchStage( s, op, c, blocks ) != 0 )                                                                               \
	return 1
	B2G_STAGE( b2g::OP_PREPARE, 0, grid );
	for ( int sub = 0; sub < P.subStepCount; ++sub )
	{
		B2G_STAGE( b2g::OP_INTEGRATE_VELOCITIES, 0, grid );
		if ( hasOverflow )
		{
			B2G_STAGE( b2g::OP_OVERFLOW_WARM, 0, 1 );
		}
		for ( int c = 0; c < P.colorCount; ++c )
		{
			B2G_STAGE( b2g::OP_WARM, c, grid );
		}
		if ( hasOverflow )
		{
			B2G_STAGE( b2g::OP_OVERFLOW_SOLVE, 0, 1 );
		}
		for ( int c = 0; c < P.colorCount; ++c )
		{
			B2G_STAGE( b2g::OP_SOLVE, c, grid );
		}
		B2G_STAGE( b2g::OP_INTEGRATE_POSITIONS, 0, grid );
		if ( hasOverflow )
		{
			B2G_STAGE( b2g::OP_OVERFLOW_RELAX, 0, 1 );
		}
		for ( int c = 0; c < P.colorCount; ++c )
		{
			B2G_STAGE( b2g::OP_RELAX, c, grid );
		}
	}
	if ( hasOverflow )
	{
		B2G_STAGE( b2g::OP_OVERFLOW_RESTITUTION, 0, 1 );
	}
	for ( int c = 0; c < P.colorCount; ++c )
	{
		B2G_STAGE( b2g::OP_RESTITUTION, c, grid );
	}
	B2G_STAGE( b2g::OP_STORE, 0, grid );
#undef B2G_STAGE
	return 0;
}


static int b2gLaunchGridKernel( b2GpuSolver* s )
{
	void* args[] = { (void*)&s->params };
	cudaError_t err;
	if ( s->cooperative )
	{
		err = cudaLaunchCooperativeKernel( (const void*)b2g::b2gStepKernel, dim3( s->gridBlocks ), dim3( b2g::kBlockThreads ), args,
										   (size_t)s->params.gridJointCache * b2g::kJointStride, s->stream );
	}
	else
	{
		err = cudaLaunchKernel( (const void*)b2g::b2gStepKernel, dim3( s->gridBlocks ), dim3( b2g::kBlockThreads ), args,
								(size_t)s->params.gridJointCache * b2g::kJointStride, s->stream );
	}
	if ( err != cudaSuccess )
	{
		return b2gFail( "b2gStepKernel launch", err );
	}
	s->lastLaunches += 1;
	return 0;
}

// The island kernels give up when a bin turns out not to fit its block (rare: the planner sizes the bins from the body
// counts and cannot see how the constraints spread).  The flag comes back with the control block; the step is then run
// again on the grid-barrier kernel from the untouched inputs.  Called after the stream has been synchronised.
int b2gRerunIfIslandsFailed( b2GpuSolver* s, bool download )
{
	if ( !s->islandMode || s->mode != 0 || s->hControl->islandFailed == 0 )
	{
		return 0;
	}
	s->countersClean = false; // the island kernels returned before zeroing their counters
	s->listsValid = false;
	s->planValid = false;
	if ( s->params.ownerLists != 0 )
	{
		s->ownerListsOff = 512; // a block's share did not fit: deal the colours out evenly for a while
	}
	else if ( s->islandSizesExact )
	{
		s->exactHeadRoom = s->exactHeadRoom * 1.15 < 3.0 ? s->exactHeadRoom * 1.15 : 3.0;
		s->headRoomCooldown = 512;
	}
	else
	{
		s->islandHeadRoom = s->islandHeadRoom * 1.3 < 3.0 ? s->islandHeadRoom * 1.3 : 3.0;
		s->headRoomCooldown = 512;
	}
	B2G_CUDA( cudaMemsetAsync( s->control, 0, sizeof( ControlBlock ), s->stream ) );
	if ( b2gLaunchGridKernel( s ) != 0 )
	{
		return 1;
	}
	B2G_CUDA( cudaEventRecord( s->evStop, s->stream ) );
	s->launchCount += 1;
	if ( download )
	{
		if ( b2gEnqueueDownload( s ) != 0 )
		{
			return 1;
		}
	}
	else
	{
		B2G_CUDA( cudaMemcpyAsync( s->hControl, s->control, sizeof( ControlBlock ), cudaMemcpyDeviceToHost, s->stream ) );
	}
	B2G_CUDA( cudaStreamSynchronize( s->stream ) );
	return 0;
}

static int b2gEnqueueRun( b2GpuSolver* s )
{
	if ( !s->uploaded )
	{
		return b2gFailMsg( "b2GpuSolverRun: nothing uploaded" );
	}
	B2G_CUDA( cudaSetDevice( s->device ) );
	s->lastLaunches = 0;
	if ( !s->controlClean )
	{
		B2G_CUDA( cudaMemsetAsync( s->control, 0, sizeof( ControlBlock ), s->stream ) );
	}
	s->controlClean = false;
	B2G_CUDA( cudaEventRecord( s->evStart, s->stream ) );
	if ( s->resident && s->params.dirtyBodyCapacity > 0 )
	{
		// bodies the host touched since the last step overwrite their resident copies first (b2g_resident.cuh)
		int blocks = ( s->params.dirtyBodyCapacity + 255 ) / 256;
		blocks = blocks > s->smCount * 8 ? s->smCount * 8 : blocks;
		b2g::b2gApplyBodiesKernel<<<blocks, 256, 0, s->stream>>>( s->params );
		cudaError_t applyErr = cudaGetLastError();
		if ( applyErr != cudaSuccess )
		{
			return b2gFail( "b2gApplyBodiesKernel launch", applyErr );
		}
		s->lastLaunches += 1;
	}
	if ( s->resident && s->params.jointCount > 0 )
	{
		// joints: the complete records of this step from the table, the previous outputs and the uploaded runs
		int blocks = ( s->params.jointCount * ( b2g::kJointStride / 16 ) + 255 ) / 256;
		blocks = blocks > s->smCount * 8 ? s->smCount * 8 : blocks;
		b2g::b2gAssembleJointsKernel<<<blocks, 256, 0, s->stream>>>( s->params );
		cudaError_t assembleErr = cudaGetLastError();
		if ( assembleErr != cudaSuccess )
		{
			return b2gFail( "b2gAssembleJointsKernel launch", assembleErr );
		}
		s->lastLaunches += 1;
	}
	if ( s->mode == 0 )
	{
		void* args[] = { (void*)&s->params };
		cudaError_t err;
		if ( s->islandMode )
		{
			// partition -> island kernel; if a bin does not fit (binFail) the host reruns the step on the grid-barrier
			// kernel once the flag has come back (b2gRerunIfIslandsFailed)
			// A steady step (see b2GpuSolver::listsValid): no contact travelled in full -- every one of them sits at the home it
			// had, with the bodies it had (on the narrow phase's word or by comparison) --, no joints, one block per bin with flat lists.  Such a step leaves its lists behind, and runs on the previous step's if that was
			// one too and nothing the lists depend on has moved.
			b2GpuSolver::ListsOf now = {};
			{
				const b2g::StepParams& P = s->params;
				now.binCount = P.binCount, now.capBodies = P.capBodies, now.capContacts = P.capContacts, now.capJoints = P.capJoints;
				now.bodyCount = P.bodyCount, now.contactSlots = P.contactSlots, now.colorCount = P.colorCount, now.jointCount = P.jointCount;
				now.jointWords = P.jointWords;
				now.clusterSize = P.clusterSize, now.ownerLists = P.ownerLists, now.listCount = P.listCount, now.clusterRun = P.clusterRun;
				now.listCapContacts = P.listCapContacts, now.flatLists = P.flatLists;
				memcpy( now.colors, P.colors, sizeof( now.colors ) );
				now.overflow = P.overflow;
				now.buffers[0] = P.binBodyCount, now.buffers[1] = P.binBodyList, now.buffers[2] = P.bodyLocal, now.buffers[3] = P.binContactInfo;
				now.buffers[4] = P.binContactList, now.buffers[5] = P.bodyBin, now.buffers[6] = P.binJointList, now.buffers[7] = P.binJointBodies;
				now.buffers[8] = P.planInfo, now.buffers[9] = P.planJoints, now.buffers[10] = P.planStart;
			}
			const bool steady = s->keepListsEnabled && s->resident && s->cacheUsable && ( s->params.flatLists != 0 || s->params.clusterSize > 1 ) &&
								s->contactTotal + s->jointTotal > 0 && s->fullCount.load( std::memory_order_relaxed ) == 0 &&
								s->fullJointCount.load( std::memory_order_relaxed ) == 0;
			const bool reuse = steady && s->listsValid && s->binsChanged.load( std::memory_order_relaxed ) == 0 &&
							   memcmp( &now, &s->listsOf, sizeof( now ) ) == 0;
			// a steady step that builds its lists writes the bins' plans down; the steps that run on those lists read them
			const bool planCapable = s->params.flatLists != 0 && s->params.clusterSize == 1 && s->params.leveliseContacts != 0;
			s->params.planRead = reuse && s->planValid && planCapable ? 1 : 0;
			s->params.planWrite = steady && !reuse && planCapable ? 1 : 0;
			s->planValid = steady && ( reuse ? s->planValid : planCapable );
			s->params.keepLists = steady ? 1 : 0;
			s->listsValid = steady; // (unless the step fails: b2gRerunIfIslandsFailed)
			s->listsOf = now;
			// (the lists this run builds, or runs on, are those of the bins the pack pass has just seen: another run of the same
			// upload -- b2GpuSolverRun can be repeated -- finds nothing changed)
			s->binsChanged.store( 0, std::memory_order_relaxed );
			if ( reuse )
			{
				s->listsReused += 1;
				err = cudaSuccess;
				if ( s->params.jointWords > 0 )
				{
					// (the partition kernels clear the joint-event bits on their way)
					B2G_CUDA( cudaMemsetAsync( s->params.jointBits, 0, (size_t)s->params.jointWords * sizeof( uint32_t ), s->stream ) );
				}
			}
			else if ( !s->countersClean )
			{
				B2G_CUDA( cudaMemsetAsync( s->binCounters.ptr, 0, s->binCounters.capacity * sizeof( int ), s->stream ) );
			}
			s->countersClean = !steady; // the island kernels zero the counters they have read, unless they keep the lists
			if ( reuse )
			{
				// nothing to partition
			}
			else if ( s->params.flatLists != 0 )
			{
				// one flat pass, one item per thread
				const b2g::StepParams& P = s->params;
				int items = P.bodyCount > P.contactSlots ? P.bodyCount : P.contactSlots;
				items = items > P.jointCount ? items : P.jointCount;
				int blocks = ( items + 255 ) / 256;
				blocks = blocks < 1 ? 1 : blocks > s->smCount * 8 ? s->smCount * 8 : blocks;
				err = cudaLaunchKernel( (const void*)b2g::b2gScatterKernel, dim3( blocks ), dim3( 256 ), args, 0, s->stream );
			}
			else if ( s->cooperative )
			{
				err = cudaLaunchCooperativeKernel( (const void*)b2g::b2gPartitionKernel, dim3( s->gridBlocks ),
												   dim3( b2g::kPartitionThreads ), args, 0, s->stream );
			}
			else
			{
				err = cudaLaunchKernel( (const void*)b2g::b2gPartitionKernel, dim3( s->gridBlocks ), dim3( b2g::kPartitionThreads ), args, 0,
										s->stream );
			}
			if ( err != cudaSuccess )
			{
				return b2gFail( "b2gPartitionKernel launch", err );
			}
			if ( s->params.clusterSize > 1 )
			{
				cudaLaunchConfig_t config = {};
				config.gridDim = dim3( (unsigned)( s->params.binCount * s->params.clusterSize ) );
				config.blockDim = dim3( b2g::kIslandThreads );
				config.dynamicSmemBytes = s->islandSmemBytes;
				config.stream = s->stream;
				cudaLaunchAttribute attribute[2];
				attribute[0].id = cudaLaunchAttributeClusterDimension;
				attribute[0].val.clusterDim.x = (unsigned)s->params.clusterSize;
				attribute[0].val.clusterDim.y = 1;
				attribute[0].val.clusterDim.z = 1;
				attribute[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
				attribute[1].val.programmaticStreamSerializationAllowed = 1;
				config.attrs = attribute;
				config.numAttrs = s->dependentLaunch && !reuse ? 2 : 1; // (reuse: no partition kernel in front of it)
				err = cudaLaunchKernelEx( &config, b2g::b2gClusterIslandKernel, s->params );
			}
			else
			{
				cudaLaunchConfig_t config = {};
				config.gridDim = dim3( (unsigned)s->params.binCount );
				config.blockDim = dim3( b2g::kIslandThreads );
				config.dynamicSmemBytes = s->islandSmemBytes;
				config.stream = s->stream;
				cudaLaunchAttribute attribute;
				attribute.id = cudaLaunchAttributeProgrammaticStreamSerialization;
				attribute.val.programmaticStreamSerializationAllowed = 1;
				config.attrs = &attribute;
				const bool noPrimary = reuse; // (no scatter kernel in front of it)
				config.numAttrs = s->params.flatLists != 0 && s->dependentLaunch && !noPrimary ? 1 : 0;
				err = cudaLaunchKernelEx( &config, b2g::b2gIslandKernel, s->params );
			}
			if ( err != cudaSuccess )
			{
				return b2gFail( "b2gIslandKernel launch", err );
			}
			s->lastLaunches += reuse ? 1 : 2;
		}
		else if ( b2gLaunchGridKernel( s ) != 0 )
		{
			return 1;
		}
	}
	else
	{
		if ( b2gRunStages( s ) != 0 )
		{
			return 1;
		}
	}
	B2G_CUDA( cudaEventRecord( s->evStop, s->stream ) );
	s->launchCount += (uint64_t)s->lastLaunches;
	s->ran = true;
	return 0;
}

// Direct outputs: the last kernel of a step.  The body states are in the host's arena already (the solve kernels stored them
// there); the control block follows the same way, then the flag.  Stream order puts this kernel behind every store of the
// solve kernels; the fences order its own stores before the flag for an observer on the host.
__global__ void b2gSignalKernel( const int* control, int* hostControl, int words, volatile int* flag, int serial )
{
	for ( int i = (int)threadIdx.x; i < words; i += (int)blockDim.x )
	{
		hostControl[i] = control[i];
	}
	__threadfence_system();
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		__threadfence_system();
		*flag = serial;
	}
}

// has the control block of the step in flight come back?  (the host's poll, b2gPumpDownloads)
int b2gPollControl( b2GpuSolver* s, bool* seen )
{
	*seen = false;
	if ( s->direct )
	{
		if ( *const_cast<volatile int*>( s->hFlag ) == s->flagSerial )
		{
			std::atomic_thread_fence( std::memory_order_acquire );
			*seen = true;
			return 0;
		}
		// the flag never comes if a kernel faulted: look at the event behind the signal kernel now and then
		s->flagPolls += 1;
		if ( ( s->flagPolls & 1023 ) != 0 )
		{
			return 0;
		}
	}
	cudaError_t err = cudaEventQuery( s->evControl );
	if ( err == cudaErrorNotReady )
	{
		return 0;
	}
	if ( err != cudaSuccess )
	{
		return b2gFail( "device solve", err );
	}
	*seen = true;
	return 0;
}

int b2gEnqueueDownload( b2GpuSolver* s )
{
	// the control block first (it says whether the island kernels gave up), then the output arena in chunks with an
	// event each, so that unpacking can start behind the transfer
	cudaStream_t st = s->stream;
	if ( s->direct )
	{
		s->flagSerial += 1;
		s->flagPolls = 0;
		b2gSignalKernel<<<1, 32, 0, st>>>( reinterpret_cast<const int*>( s->control ), reinterpret_cast<int*>( s->hControl ),
										   (int)( sizeof( ControlBlock ) / sizeof( int ) ), s->hFlag, s->flagSerial );
		cudaError_t err = cudaGetLastError();
		if ( err != cudaSuccess )
		{
			return b2gFail( "b2gSignalKernel launch", err );
		}
		s->lastLaunches += 1;
		s->launchCount += 1;
	}
	else
	{
		B2G_CUDA( cudaMemcpyAsync( s->hControl, s->control, sizeof( ControlBlock ), cudaMemcpyDeviceToHost, st ) );
	}
	B2G_CUDA( cudaEventRecord( s->evControl, st ) );
	s->controlSeen = false;
	size_t total = s->outTotal;
	const size_t chunkQuads = s->downloadQuads;
	s->chunkEnd.clear();
	// deferred impulses: the arena's tail holds the records nobody waits for (one piece, one event)
	const size_t awaited = s->defer ? s->outJoints : total; // (deferred: the states and the joint-event bits)
	s->deferWaitChunks = 0;
	int made = 0;
	for ( size_t begin = s->direct ? s->directEnd : 0; begin < total; )
	{
		size_t end = begin + chunkQuads < total ? begin + chunkQuads : total;
		end = begin < awaited && end > awaited ? awaited : end;
		end = begin >= awaited ? total : end;
		B2G_CUDA( cudaMemcpyAsync( s->hOut.ptr + begin, s->outAll.ptr + begin, ( end - begin ) * sizeof( float4 ), cudaMemcpyDeviceToHost, st ) );
		if ( (int)s->chunkEvents.size() <= made )
		{
			cudaEvent_t ev = nullptr;
			B2G_CUDA( cudaEventCreateWithFlags( &ev, cudaEventDisableTiming ) );
			s->chunkEvents.push_back( ev );
		}
		B2G_CUDA( cudaEventRecord( s->chunkEvents[(size_t)made], st ) );
		s->chunkEnd.push_back( end );
		made += 1;
		if ( end == awaited )
		{
			s->deferWaitChunks = made;
		}
		begin = end;
	}
	const int chunks = made;
	s->deferWaitChunks = s->defer && awaited < total ? s->deferWaitChunks : chunks;
	if ( s->defer )
	{
		B2G_CUDA( cudaEventRecord( s->evRecords, st ) );
	}
	s->chunkCount = chunks;
	s->chunkNext = 0;
	s->arrivedQuads.store( 0, std::memory_order_release );
	s->lastD2H = total * sizeof( float4 ) + sizeof( ControlBlock );
	// behind the download, off the host's critical path: the table's update, and the control block zeroed for the next
	// step's kernels (one driver call less between the next Submit and its first launch)
	if ( b2gEnqueueCommit( s ) != 0 )
	{
		return 1;
	}
	B2G_CUDA( cudaMemsetAsync( s->control, 0, sizeof( ControlBlock ), s->stream ) );
	s->controlClean = true;
	return 0;
}

extern "C" int b2GpuSolverSubmit( b2GpuSolver* s )
{
	if ( s == nullptr || !s->begun )
	{
		return b2gFailMsg( "b2GpuSolverSubmit: no step begun" );
	}
	s->tSubmit = std::chrono::steady_clock::now();
	s->traceSubmit = std::chrono::duration<float, std::micro>( s->tSubmit - s->tBegin ).count();
	if ( b2gEnqueueUpload( s ) != 0 || b2gEnqueueRun( s ) != 0 || b2gEnqueueDownload( s ) != 0 )
	{
		return 1;
	}
	// deferred impulses: the unpack pass has the bodies and the joints to do (b2GpuSolverUnpackWork skips the contacts' items)
	const int unpackItems = b2GpuSolverGetUnpackItemCount( s ) - ( s->defer ? s->contactTotal + s->jointTotal : 0 );
	b2gResetWork( s, unpackItems, b2gBlocksFor( s, unpackItems ) );
	return 0;
}

// ---- deferred contact impulses -------------------------------------------------------------------------------------------
// the pending records have arrived (any thread; the first one waits for the tail of the previous step's download)
int b2gDeferSync( b2GpuSolver* s )
{
	if ( s->recordsSynced.load( std::memory_order_acquire ) == 0 )
	{
		cudaError_t err = cudaEventSynchronize( s->evRecords );
		if ( err != cudaSuccess )
		{
			return b2gFail( "download of the deferred impulse records", err );
		}
		s->recordsSynced.store( 1, std::memory_order_release );
	}
	return 0;
}

extern "C" int b2GpuSolverSetDeferredImpulses( b2GpuSolver* s, int enabled )
{
	if ( s == nullptr || s->begun )
	{
		return b2gFailMsg( "b2GpuSolverSetDeferredImpulses: no solver, or a step is in flight" );
	}
	if ( s->deferPending )
	{
		return b2gFailMsg( "b2GpuSolverSetDeferredImpulses: impulses are pending" );
	}
	if ( ( enabled != 0 ) != s->deferEnabled )
	{
		s->deferEnabled = enabled != 0;
		s->cacheValid = false; // the two modes keep the previous step's impulses in different places
	}
	return 0;
}

extern "C" int b2GpuSolverDeferredPending( const b2GpuSolver* s )
{
	return s != nullptr && s->deferPending ? 1 : 0;
}

extern "C" int b2GpuSolverDeferredSync( b2GpuSolver* s )
{
	if ( s == nullptr )
	{
		return b2gFailMsg( "b2GpuSolverDeferredSync: null solver" );
	}
	return s->deferPending ? b2gDeferSync( s ) : 0;
}

extern "C" void b2GpuSolverDeferredDone( b2GpuSolver* s )
{
	if ( s != nullptr )
	{
		s->deferPending = false;
		s->deferJointsPending = false;
	}
}

extern "C" int b2GpuSolverWait( b2GpuSolver* s )
{
	if ( s == nullptr )
	{
		return b2gFailMsg( "b2GpuSolverWait: null solver" );
	}
	B2G_CUDA( cudaStreamSynchronize( s->stream ) );
	if ( s->ran && b2gRerunIfIslandsFailed( s, true ) != 0 )
	{
		return 1;
	}
	s->controlSeen = true;
	s->chunkNext = s->chunkCount;
	s->arrivedQuads.store( s->outTotal, std::memory_order_release );
	s->tWaited = std::chrono::steady_clock::now();
	if ( s->ran )
	{
		B2G_CUDA( cudaEventElapsedTime( &s->lastKernelMs, s->evStart, s->evStop ) );
	}
	return 0;
}

static void b2gFillTimers( b2GpuSolver* s, b2GpuStepResult* r )
{
	// stage split from the in-kernel cycle counters, scaled to the CUDA-event kernel time
	const ControlBlock* c = s->hControl;
	unsigned long long total = 0;
	for ( int i = 0; i < b2GpuStage_count; ++i )
	{
		total += c->stageCycles[i];
	}
	for ( int i = 0; i < b2GpuStage_count; ++i )
	{
		r->stageMs[i] = total > 0 ? s->lastKernelMs * (float)( (double)c->stageCycles[i] / (double)total ) : 0.0f;
	}
	r->gridBarriers = (int)c->stageCycles[8];
}

static int b2gEnd( b2GpuSolver* s, b2GpuStepResult* results )
{
	if ( s == nullptr || !s->begun )
	{
		return b2gFailMsg( "b2GpuSolverEndStep: no step begun" );
	}
	auto ms = []( std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b ) {
		return std::chrono::duration<float, std::milli>( b - a ).count();
	};
	s->traceMarks[6] = std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count();
	if ( results != nullptr )
	{
		const uint32_t* bits = reinterpret_cast<const uint32_t*>( s->hOut.ptr + s->outBits );
		auto now = std::chrono::steady_clock::now();
		float h2dMs = 0.0f;
		static const bool eagerTimers = getenv( "B2GPU_EAGER_TIMERS" ) != nullptr && atoi( getenv( "B2GPU_EAGER_TIMERS" ) ) != 0; // (A/B)
		if ( s->trace || eagerTimers )
		{
			// (a driver call of 4 - 5 us on the way out of every step, for a number only the trace shows)
			cudaEventElapsedTime( &h2dMs, s->evUpload, s->evStart );
		}
		s->traceMarks[7] = std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count();
		for ( size_t w = 0; w < s->bodySegs.size(); ++w )
		{
			const b2gBodySeg& seg = s->bodySegs[w];
			b2GpuStepResult* r = results + w;
			// uint32 pairs are the little-endian halves of the reference's uint64 blocks (src/bitset.h)
			if ( r->jointEventBits != nullptr )
			{
				const uint32_t* mine = bits + seg.jointBitBase / 32;
				for ( int i = 0; i < seg.jointWords / 2; ++i )
				{
					uint64_t word = (uint64_t)mine[2 * i] | ( (uint64_t)mine[2 * i + 1] << 32 );
					r->jointEventBits[i] |= word;
				}
			}
			if ( s->defer && s->hControl->hasHitEvents != 0 )
			{
				// nobody has looked at the records: the caller materializes them (that sets the bits) before it reads the events
				r->hasHitEvents = 1;
			}
			r->kernelMs = s->lastKernelMs;
			r->kernelLaunches = s->lastLaunches;
			r->h2dBytes = s->lastH2D;
			r->d2hBytes = s->lastD2H;
			b2gFillTimers( s, r );
			r->uploadMs = ms( s->tBegin, s->tSubmit );	 // layout + packing
			r->waitMs = ms( s->tSubmit, s->tWaited );	 // H2D + kernels + D2H
			r->scatterMs = ms( s->tWaited, now );		 // unpack + event bits
			r->h2dMs = h2dMs;
			r->totalMs = ms( s->tBegin, now );
		}
		b2gFlushLines( bits, (size_t)s->params.jointWords * sizeof( uint32_t ) );
	}
	if ( s->trace )
	{
		fprintf( stderr, "[b2gpu] in %zu quads, out %zu quads | begin marks %.0f %.0f %.0f %.0f %.0f | begun %.0f | sends (us: upto):", s->inTotal, s->outTotal,
				 s->traceMarks[0], s->traceMarks[1], s->traceMarks[2], s->traceMarks[3], s->traceMarks[4], s->traceBegun );
		for ( auto& e : s->traceSends )
		{
			fprintf( stderr, " %.0f:%zu", e.first, e.second );
		}
		fprintf( stderr, " | pump saw (us: blocks of %d):", s->workBlocks );
		for ( auto& e : s->tracePump )
		{
			fprintf( stderr, " %.0f:%zu", e.first, e.second );
		}
		{
			float h2dMs = 0.0f;
			cudaEventElapsedTime( &h2dMs, s->evUpload, s->evStart );
			fprintf( stderr, " | first copy to kernels %.0f us (device clock)", h2dMs * 1000.0f );
		}
		fprintf( stderr, " | submit %.0f | kernels done %.0f | arrivals:", s->traceSubmit, s->traceControl );
		for ( auto& e : s->traceArrivals )
		{
			fprintf( stderr, " %.0f:%zu", e.first, e.second );
		}
		fprintf( stderr, " | pump out of blocks %.0f | EndStep %.0f | h2d timer read %.0f | results filled %.0f", s->traceMarks[5], s->traceMarks[6],
				 s->traceMarks[7], std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count() );
		fprintf( stderr, " | end %.0f | kernel %.0f us\n",
				 std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count(), s->lastKernelMs * 1000.0f );
	}
	s->traceSends.clear();
	s->tracePump.clear();
	s->traceArrivals.clear();
	s->begun = false;
	s->liteJointsSeen = s->jointTotal > 0 && s->heavyJoint.load( std::memory_order_relaxed ) == 0;
	bool deferredNow = false;
	if ( s->resident && s->ran && s->workFailed.load() == 0 )
	{
		// this step's outputs are the next step's resident inputs
		if ( s->defer )
		{
			// ... and its impulse records wait in the arena that was just filled (the tail may still be on its way)
			std::swap( s->hOut, s->hOutOther );
			s->pendingRecords = reinterpret_cast<const float*>( s->hOutOther.ptr + s->outImpulses );
			s->pendingJointRecords = reinterpret_cast<const float*>( s->hOutOther.ptr + s->outJoints );
			s->deferJointsPending = s->jointTotal > 0;
			s->deferPending = s->contactTotal > 0 || s->jointTotal > 0;
			s->deferStamp = s->deferNewStamp;
			s->recordsSynced.store( 0, std::memory_order_release );
			deferredNow = s->deferPending;
		}
		std::swap( s->outAll, s->outOther );
		std::swap( s->residentStates[0], s->residentStates[1] );
		s->prevOutImpulses = s->outImpulses;
		s->prevOutJoints = s->outJoints;
		for ( int key = 0; key < kHomeColors; ++key )
		{
			s->jointHomeCount[key] = 0;
		}
		for ( size_t k = 0; k < s->jointSegs.size(); ++k )
		{
			s->jointHomeCount[s->jointSegHome[k]] = s->jointSegs[k].count;
			s->jointHomeSlot[s->jointSegHome[k]] = s->jointSegs[k].jointStart;
		}
		s->shadowBodyCount = s->params.bodyCount;
		for ( int key = 0; key < kHomeColors; ++key )
		{
			s->homeCount[key] = 0;
		}
		for ( size_t k = 0; k < s->contactSegs.size(); ++k )
		{
			s->homeCount[s->segHome[k]] = s->contactSegs[k].count;
			s->homeSlot[s->segHome[k]] = s->contactSegs[k].slotStart;
		}
		s->cacheValid = true;
	}
	if ( deferredNow && ( !s->homesOrdered || ( s->jointTotal > 0 && !s->jointHomesOrdered ) ) )
	{
		// pending records are found by (graph colour, place); a caller whose colours do not come with ascending indices has
		// no such address: its manifolds are written now, like a step that does not defer
		return b2gMaterializePendingFromSegs( s, s->results );
	}
	return 0;
}

extern "C" int b2GpuSolverEndStep( b2GpuSolver* s, b2GpuStepResult* r )
{
	return b2gEnd( s, r );
}

// ---- the whole step, single host thread ----------------------------------------------------------------------------------
// The two host passes of the one-call entry points.  A world stepped through the seam uses the world's own workers
// (b2ParallelFor); a caller of b2GpuSolverStep / StepBatch has no task system to offer, so large steps are split over
// a few short-lived threads here (B2GPU_HOST_THREADS overrides the count, 1 = calling thread only).
static int b2gWorkWithThreads( b2GpuSolver* s, int ( *work )( b2GpuSolver*, int ) )
{
	static const int configured = []() {
		const char* env = getenv( "B2GPU_HOST_THREADS" );
		int n = env != nullptr ? atoi( env ) : (int)std::thread::hardware_concurrency();
		return n < 1 ? 1 : ( n > 32 ? 32 : n );
	}();
	int threads = s->workItems / 16384;
	threads = threads > configured ? configured : threads;
	std::vector<std::thread> helpers;
	for ( int t = 1; t < threads; ++t )
	{
		helpers.emplace_back( work, s, 0 );
	}
	int rc = work( s, 1 ); // the calling thread is the pump
	for ( std::thread& th : helpers )
	{
		th.join();
	}
	return rc != 0 || s->workFailed.load() != 0 ? 1 : 0;
}

static int b2gStepAll( b2GpuSolver* s, const b2GpuStepDesc* descs, int worldCount, b2GpuStepResult* results )
{
	if ( b2gBegin( s, descs, worldCount, results ) != 0 )
	{
		return 1;
	}
	if ( b2gWorkWithThreads( s, b2GpuSolverPackWork ) != 0 || b2GpuSolverSubmit( s ) != 0 ||
		 b2gWorkWithThreads( s, b2GpuSolverUnpackWork ) != 0 )
	{
		return 1;
	}
	return b2gEnd( s, results );
}

extern "C" int b2GpuSolverStep( b2GpuSolver* s, const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	return b2gStepAll( s, d, 1, r );
}

// ---- split for benchmarks: Upload (pack + H2D), Run (kernels only, repeatable), Download (D2H + unpack) -------------------
static int b2gUploadAll( b2GpuSolver* s, const b2GpuStepDesc* descs, int worldCount )
{
	if ( b2gBegin( s, descs, worldCount, nullptr ) != 0 )
	{
		return 1;
	}
	b2GpuSolverPackRange( s, 0, b2GpuSolverGetPackItemCount( s ) );
	s->tSubmit = std::chrono::steady_clock::now();
	if ( b2gEnqueueUpload( s ) != 0 )
	{
		return 1;
	}
	B2G_CUDA( cudaStreamSynchronize( s->stream ) );
	return 0;
}

extern "C" int b2GpuSolverUpload( b2GpuSolver* s, const b2GpuStepDesc* d )
{
	return b2gUploadAll( s, d, 1 );
}

extern "C" int b2GpuSolverRun( b2GpuSolver* s, b2GpuStepResult* r )
{
	if ( s == nullptr )
	{
		return b2gFailMsg( "b2GpuSolverRun: null solver" );
	}
	if ( b2gEnqueueRun( s ) != 0 )
	{
		return 1;
	}
	B2G_CUDA( cudaMemcpyAsync( s->hControl, s->control, sizeof( ControlBlock ), cudaMemcpyDeviceToHost, s->stream ) );
	B2G_CUDA( cudaStreamSynchronize( s->stream ) );
	if ( b2gRerunIfIslandsFailed( s, false ) != 0 )
	{
		return 1;
	}
	B2G_CUDA( cudaEventElapsedTime( &s->lastKernelMs, s->evStart, s->evStop ) );
	if ( r != nullptr )
	{
		r->kernelMs = s->lastKernelMs;
		r->kernelLaunches = s->lastLaunches;
		r->hasHitEvents = s->hControl->hasHitEvents;
		b2gFillTimers( s, r );
	}
	return 0;
}

static int b2gDownloadAll( b2GpuSolver* s, const b2GpuStepDesc* descs, int worldCount, b2GpuStepResult* results )
{
	if ( s == nullptr || descs == nullptr )
	{
		return b2gFailMsg( "b2GpuSolverDownload: null argument" );
	}
	if ( !s->ran || !s->begun || (int)s->bodySegs.size() != worldCount )
	{
		return b2gFailMsg( "b2GpuSolverDownload: nothing has run for these worlds" );
	}
	s->results = results;
	if ( b2gEnqueueDownload( s ) != 0 || b2GpuSolverWait( s ) != 0 )
	{
		return 1;
	}
	b2GpuSolverUnpackRange( s, 0, b2GpuSolverGetUnpackItemCount( s ) );
	return b2gEnd( s, results );
}

extern "C" int b2GpuSolverDownload( b2GpuSolver* s, const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	return b2gDownloadAll( s, d, 1, r );
}

// ---- batch of independent worlds -------------------------------------------------------------------------------------------
extern "C" int b2GpuSolverStepBatch( b2GpuSolver* s, const b2GpuStepDesc* descs, int worldCount, b2GpuStepResult* results )
{
	return b2gStepAll( s, descs, worldCount, results );
}

extern "C" int b2GpuSolverUploadBatch( b2GpuSolver* s, const b2GpuStepDesc* descs, int worldCount )
{
	return b2gUploadAll( s, descs, worldCount );
}

extern "C" int b2GpuSolverRunBatch( b2GpuSolver* s, b2GpuStepResult* r )
{
	return b2GpuSolverRun( s, r );
}

extern "C" int b2GpuSolverDownloadBatch( b2GpuSolver* s, const b2GpuStepDesc* descs, int worldCount, b2GpuStepResult* results )
{
	return b2gDownloadAll( s, descs, worldCount, results );
}

