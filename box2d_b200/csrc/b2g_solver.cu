// b2g_solver.cu -- kernels + the C-ABI of include/b2_gpu_solver.h.
//
// Single world: ONE persistent cooperative kernel per step (b2gStepKernel) runs the whole b2SolverTask
// stage sequence (reference src/solver.c:1055-1197) with a grid-wide barrier where the reference's
// orchestrator spins on stage->completionCount (src/solver.c:999-1005).  Body and constraint state stay
// resident on the device across all sub-steps; the host sees one launch.
// Debug/profiling path: the same device functions, one kernel launch per stage (mode 1).
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
#include "b2_gpu_solver.h"

#include "b2g_stages.cuh"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace b2g
{

constexpr int kBlockThreads = 256;

// ---- grid barrier ---------------------------------------------------------------------------------------------
// Arrive = release-add at gpu scope by one thread after the block has synchronised; wait = acquire-load spin.
// The acquire makes the other blocks' body/constraint writes visible to every thread of this block after the
// trailing __syncthreads (PTX memory model: bar.sync and release/acquire chains compose by causality order).
B2G_DEV void gridBarrier( unsigned int* counter, unsigned int target )
{
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		asm volatile( "red.release.gpu.global.add.u32 [%0], 1;" ::"l"( counter ) : "memory" );
		unsigned int seen;
		do
		{
			asm volatile( "ld.acquire.gpu.global.u32 %0, [%1];" : "=r"( seen ) : "l"( counter ) : "memory" );
		}
		while ( seen < target );
	}
	__syncthreads();
}

struct StageClock
{
	long long last;
	long long acc[b2GpuStage_count];
	bool lead;

	B2G_DEV void start()
	{
		lead = isLeadThread();
		last = lead ? clock64() : 0;
#pragma unroll
		for ( int i = 0; i < b2GpuStage_count; ++i )
		{
			acc[i] = 0;
		}
	}

	// timers are compile-time constants, so acc[] stays in registers
	B2G_DEV void lap( int timer )
	{
		if ( lead )
		{
			long long now = clock64();
			acc[timer] += now - last;
			last = now;
		}
	}
};

// The whole step.  Stage order and barrier placement = b2SolverTask (src/solver.c:1055-1197); the stage timers
// are the reference's b2Profile split (src/solver.c:1080,1097,1112,1132,1141,1159,1182,1191).
__global__ void __launch_bounds__( kBlockThreads, 1 ) b2gStepKernel( const __grid_constant__ StepParams P )
{
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

	runStage( P, OP_PREPARE, 0 );
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
			runStage( P, OP_WARM, c );
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
			runStage( P, OP_SOLVE, c );
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
			runStage( P, OP_RELAX, c );
			sync();
		}
		clk.lap( b2GpuStage_relaxImpulses );
	}

	// Restitution: the reference skips every SIMD group whose lanes all have restitution 0
	// (src/contact_solver.c:2131, :432); when NO contact of the step has any, all groups skip, so the colour
	// stages and their barriers are skipped as a whole.
	if ( __ldcg( P.anyRestitution ) != 0 )
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

	runStage( P, OP_STORE, 0 );
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

// =================================================================================================================
// Host side
// =================================================================================================================

static thread_local std::string t_lastError;

static int b2gFail( const char* what, cudaError_t err )
{
	t_lastError = std::string( what ) + ": " + cudaGetErrorString( err );
	return 1;
}

static int b2gFailMsg( const char* what )
{
	t_lastError = what;
	return 1;
}

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
			cudaFreeHost( ptr );
			ptr = nullptr;
			capacity = 0;
		}
		cudaError_t err = cudaHostAlloc( &ptr, newCapacity * sizeof( T ), cudaHostAllocDefault );
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
			cudaFreeHost( ptr );
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
};

struct b2GpuSolver
{
	int device = 0;
	int smCount = 0;
	int gridBlocks = 0;
	int mode = 0;
	bool cooperative = false;
	cudaStream_t stream = nullptr;
	cudaEvent_t evStart = nullptr, evStop = nullptr, evUpload = nullptr;

	DeviceBuffer<uint8_t> rawStates, rawSims, rawContacts, rawJoints, joints, outStates;
	DeviceBuffer<float4> vel, pos, bodyK, cf;
	DeviceBuffer<float> angDamp, outImpulses;
	DeviceBuffer<int2> cidx, cmeta;
	DeviceBuffer<uint32_t> hitBits, jointBits;
	ControlBlock* control = nullptr;

	PinnedBuffer<float> hImpulses;
	PinnedBuffer<uint32_t> hBits;
	ControlBlock* hControl = nullptr;

	b2g::StepParams params;
	bool uploaded = false;
	bool ran = false;
	uint64_t launchCount = 0;
	uint64_t lastH2D = 0;
	int lastLaunches = 0;
	float lastKernelMs = 0.0f;
};

static int b2gRoundUp32( int n )
{
	return ( n + 31 ) & ~31;
}

extern "C" int b2GpuGetVersion( void )
{
	return 100;
}

extern "C" const char* b2GpuGetLastError( void )
{
	return t_lastError.c_str();
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
	// one persistent block per SM: the grid barrier needs every block co-resident
	s->gridBlocks = s->smCount;
	const char* gridEnv = getenv( "B2GPU_GRID" );
	if ( gridEnv != nullptr && atoi( gridEnv ) > 0 && atoi( gridEnv ) <= s->smCount * blocksPerSm )
	{
		s->gridBlocks = atoi( gridEnv );
	}

	bool ok = cudaStreamCreateWithFlags( &s->stream, cudaStreamNonBlocking ) == cudaSuccess;
	ok = ok && cudaEventCreate( &s->evStart ) == cudaSuccess;
	ok = ok && cudaEventCreate( &s->evStop ) == cudaSuccess;
	ok = ok && cudaEventCreate( &s->evUpload ) == cudaSuccess;
	ok = ok && cudaMalloc( &s->control, sizeof( ControlBlock ) ) == cudaSuccess;
	ok = ok && cudaHostAlloc( &s->hControl, sizeof( ControlBlock ), cudaHostAllocDefault ) == cudaSuccess;
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
	s->rawStates.release();
	s->rawSims.release();
	s->rawContacts.release();
	s->rawJoints.release();
	s->joints.release();
	s->outStates.release();
	s->vel.release();
	s->pos.release();
	s->bodyK.release();
	s->cf.release();
	s->angDamp.release();
	s->outImpulses.release();
	s->cidx.release();
	s->cmeta.release();
	s->hitBits.release();
	s->jointBits.release();
	s->hImpulses.release();
	s->hBits.release();
	if ( s->control != nullptr )
	{
		cudaFree( s->control );
	}
	if ( s->hControl != nullptr )
	{
		cudaFreeHost( s->hControl );
	}
	if ( s->evStart != nullptr )
	{
		cudaEventDestroy( s->evStart );
	}
	if ( s->evStop != nullptr )
	{
		cudaEventDestroy( s->evStop );
	}
	if ( s->evUpload != nullptr )
	{
		cudaEventDestroy( s->evUpload );
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

// ---- upload ---------------------------------------------------------------------------------------------------
extern "C" int b2GpuSolverUpload( b2GpuSolver* s, const b2GpuStepDesc* d )
{
	if ( s == nullptr || d == nullptr )
	{
		return b2gFailMsg( "b2GpuSolverUpload: null argument" );
	}
	if ( d->activeColorCount < 0 || d->activeColorCount > b2g::kMaxColors || d->awakeBodyCount < 0 || d->subStepCount < 0 )
	{
		return b2gFailMsg( "b2GpuSolverUpload: bad descriptor" );
	}
	B2G_CUDA( cudaSetDevice( s->device ) );
	s->uploaded = false;
	s->ran = false;

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
	P.bodyCount = d->awakeBodyCount;
	P.colorCount = d->activeColorCount;

	// slot layout: every colour starts on a multiple of 32, overflow last
	int slot = 0, joint = 0;
	for ( int c = 0; c < d->activeColorCount; ++c )
	{
		const b2GpuColorDesc& color = d->colors[c];
		if ( color.contactCount < 0 || color.jointCount < 0 )
		{
			return b2gFailMsg( "b2GpuSolverUpload: negative count" );
		}
		P.colors[c].contactStart = slot;
		P.colors[c].contactCount = color.contactCount;
		P.colors[c].jointStart = joint;
		P.colors[c].jointCount = color.jointCount;
		slot += b2gRoundUp32( color.contactCount );
		joint += color.jointCount;
	}
	P.overflow.contactStart = slot;
	P.overflow.contactCount = d->overflow.contactCount;
	P.overflow.jointStart = joint;
	P.overflow.jointCount = d->overflow.jointCount;
	slot += b2gRoundUp32( d->overflow.contactCount );
	joint += d->overflow.jointCount;
	P.contactSlots = slot;
	P.jointCount = joint;
	P.hitWords = 2 * ( ( d->contactIdCapacity + 63 ) / 64 );
	P.jointWords = 2 * ( ( d->jointIdCapacity + 63 ) / 64 );

	size_t bodies = (size_t)P.bodyCount;
	B2G_CUDA( s->rawStates.reserve( bodies * B2L_STATE_SIZE + 16 ) );
	B2G_CUDA( s->rawSims.reserve( bodies * B2L_SIM_SIZE + 16 ) );
	B2G_CUDA( s->outStates.reserve( bodies * B2L_STATE_SIZE + 16 ) );
	B2G_CUDA( s->vel.reserve( bodies + 1 ) );
	B2G_CUDA( s->pos.reserve( bodies + 1 ) );
	B2G_CUDA( s->bodyK.reserve( bodies + 1 ) );
	B2G_CUDA( s->angDamp.reserve( bodies + 1 ) );
	B2G_CUDA( s->rawContacts.reserve( (size_t)slot * B2L_CONTACT_SIZE + 16 ) );
	B2G_CUDA( s->cidx.reserve( (size_t)slot + 1 ) );
	B2G_CUDA( s->cmeta.reserve( (size_t)slot + 1 ) );
	// the SoA field stride follows cidx's capacity so that all per-slot arrays grow together
	size_t slotCapacity = s->cidx.capacity;
	B2G_CUDA( s->cf.reserve( slotCapacity * b2g::CF_COUNT ) );
	B2G_CUDA( s->outImpulses.reserve( slotCapacity * b2g::kImpulseFloats ) );
	B2G_CUDA( s->rawJoints.reserve( (size_t)joint * B2L_JOINT_SIZE + 16 ) );
	B2G_CUDA( s->joints.reserve( (size_t)joint * B2L_JOINT_SIZE + 16 ) );
	B2G_CUDA( s->hitBits.reserve( (size_t)P.hitWords + 2 ) );
	B2G_CUDA( s->jointBits.reserve( (size_t)P.jointWords + 2 ) );
	B2G_CUDA( s->hImpulses.reserve( (size_t)slot * b2g::kImpulseFloats + 16 ) );
	B2G_CUDA( s->hBits.reserve( (size_t)P.hitWords + (size_t)P.jointWords + 4 ) );
	P.slotCapacity = (int)slotCapacity;

	P.rawStates = s->rawStates.ptr;
	P.rawSims = s->rawSims.ptr;
	P.rawContacts = s->rawContacts.ptr;
	P.rawJoints = s->rawJoints.ptr;
	P.vel = s->vel.ptr;
	P.pos = s->pos.ptr;
	P.bodyK = s->bodyK.ptr;
	P.angDamp = s->angDamp.ptr;
	P.cf = s->cf.ptr;
	P.cidx = s->cidx.ptr;
	P.cmeta = s->cmeta.ptr;
	P.joints = s->joints.ptr;
	P.outStates = s->outStates.ptr;
	P.outImpulses = s->outImpulses.ptr;
	P.hitBits = s->hitBits.ptr;
	P.jointBits = s->jointBits.ptr;
	P.hasHitEvents = &s->control->hasHitEvents;
	P.anyRestitution = &s->control->anyRestitution;
	P.barrier = s->control->barrier;
	P.stageCycles = s->control->stageCycles;

	// host -> device: the reference's own arrays, no host-side repacking
	uint64_t bytes = 0;
	cudaStream_t st = s->stream;
	B2G_CUDA( cudaEventRecord( s->evUpload, st ) );
	if ( bodies > 0 )
	{
		B2G_CUDA( cudaMemcpyAsync( s->rawStates.ptr, d->states, bodies * B2L_STATE_SIZE, cudaMemcpyHostToDevice, st ) );
		B2G_CUDA( cudaMemcpyAsync( s->rawSims.ptr, d->sims, bodies * B2L_SIM_SIZE, cudaMemcpyHostToDevice, st ) );
		bytes += bodies * ( B2L_STATE_SIZE + B2L_SIM_SIZE );
	}
	for ( int c = 0; c <= d->activeColorCount; ++c )
	{
		const b2GpuColorDesc& color = c < d->activeColorCount ? d->colors[c] : d->overflow;
		const b2g::ColorRange& range = c < d->activeColorCount ? P.colors[c] : P.overflow;
		if ( color.contactCount > 0 )
		{
			size_t n = (size_t)color.contactCount * B2L_CONTACT_SIZE;
			B2G_CUDA( cudaMemcpyAsync( s->rawContacts.ptr + (size_t)range.contactStart * B2L_CONTACT_SIZE, color.contactSims, n,
									   cudaMemcpyHostToDevice, st ) );
			bytes += n;
		}
		if ( color.jointCount > 0 )
		{
			size_t n = (size_t)color.jointCount * B2L_JOINT_SIZE;
			B2G_CUDA( cudaMemcpyAsync( s->rawJoints.ptr + (size_t)range.jointStart * B2L_JOINT_SIZE, color.jointSims, n,
									   cudaMemcpyHostToDevice, st ) );
			bytes += n;
		}
	}
	s->lastH2D = bytes;
	s->uploaded = true;
	return 0;
}

// ---- run -----------------------------------------------------------------------------------------------------
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

static int b2gEnqueueRun( b2GpuSolver* s )
{
	if ( !s->uploaded )
	{
		return b2gFailMsg( "b2GpuSolverRun: nothing uploaded" );
	}
	B2G_CUDA( cudaSetDevice( s->device ) );
	s->lastLaunches = 0;
	B2G_CUDA( cudaMemsetAsync( s->control, 0, sizeof( ControlBlock ), s->stream ) );
	B2G_CUDA( cudaEventRecord( s->evStart, s->stream ) );
	if ( s->mode == 0 )
	{
		void* args[] = { (void*)&s->params };
		cudaError_t err;
		if ( s->cooperative )
		{
			err = cudaLaunchCooperativeKernel( (const void*)b2g::b2gStepKernel, dim3( s->gridBlocks ), dim3( b2g::kBlockThreads ),
											   args, 0, s->stream );
		}
		else
		{
			err = cudaLaunchKernel( (const void*)b2g::b2gStepKernel, dim3( s->gridBlocks ), dim3( b2g::kBlockThreads ), args, 0,
									s->stream );
		}
		if ( err != cudaSuccess )
		{
			return b2gFail( "b2gStepKernel launch", err );
		}
		s->lastLaunches = 1;
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

// ---- download --------------------------------------------------------------------------------------------------
static int b2gEnqueueDownload( b2GpuSolver* s, const b2GpuStepDesc* d, uint64_t* bytesOut )
{
	const b2g::StepParams& P = s->params;
	cudaStream_t st = s->stream;
	uint64_t bytes = 0;
	size_t bodies = (size_t)P.bodyCount;
	if ( bodies > 0 )
	{
		B2G_CUDA( cudaMemcpyAsync( d->states, s->outStates.ptr, bodies * B2L_STATE_SIZE, cudaMemcpyDeviceToHost, st ) );
		bytes += bodies * B2L_STATE_SIZE;
	}
	if ( P.contactSlots > 0 )
	{
		size_t n = (size_t)P.contactSlots * b2g::kImpulseFloats * sizeof( float );
		B2G_CUDA( cudaMemcpyAsync( s->hImpulses.ptr, s->outImpulses.ptr, n, cudaMemcpyDeviceToHost, st ) );
		bytes += n;
	}
	for ( int c = 0; c <= d->activeColorCount; ++c )
	{
		const b2GpuColorDesc& color = c < d->activeColorCount ? d->colors[c] : d->overflow;
		const b2g::ColorRange& range = c < d->activeColorCount ? P.colors[c] : P.overflow;
		if ( color.jointCount > 0 )
		{
			size_t n = (size_t)color.jointCount * B2L_JOINT_SIZE;
			B2G_CUDA( cudaMemcpyAsync( color.jointSims, s->joints.ptr + (size_t)range.jointStart * B2L_JOINT_SIZE, n,
									   cudaMemcpyDeviceToHost, st ) );
			bytes += n;
		}
	}
	if ( P.hitWords > 0 )
	{
		B2G_CUDA( cudaMemcpyAsync( s->hBits.ptr, s->hitBits.ptr, (size_t)P.hitWords * 4, cudaMemcpyDeviceToHost, st ) );
		bytes += (uint64_t)P.hitWords * 4;
	}
	if ( P.jointWords > 0 )
	{
		B2G_CUDA( cudaMemcpyAsync( s->hBits.ptr + P.hitWords, s->jointBits.ptr, (size_t)P.jointWords * 4, cudaMemcpyDeviceToHost, st ) );
		bytes += (uint64_t)P.jointWords * 4;
	}
	B2G_CUDA( cudaMemcpyAsync( s->hControl, s->control, sizeof( ControlBlock ), cudaMemcpyDeviceToHost, st ) );
	bytes += sizeof( ControlBlock );
	*bytesOut = bytes;
	return 0;
}

// Scatter the packed impulse records into the reference's manifolds: what b2StoreImpulsesTask
// (src/contact_solver.c:2293-2303) and b2StoreImpulses_Overflow (:526-542) write.
static void b2gScatterImpulses( const b2GpuSolver* s, const b2GpuStepDesc* d )
{
	const b2g::StepParams& P = s->params;
	for ( int c = 0; c <= d->activeColorCount; ++c )
	{
		bool wide = c < d->activeColorCount;
		const b2GpuColorDesc& color = wide ? d->colors[c] : d->overflow;
		const b2g::ColorRange& range = wide ? P.colors[c] : P.overflow;
		uint8_t* sims = static_cast<uint8_t*>( color.contactSims );
		const float* records = s->hImpulses.ptr + (size_t)range.contactStart * b2g::kImpulseFloats;
		for ( int i = 0; i < color.contactCount; ++i )
		{
			uint8_t* manifold = sims + (size_t)i * B2L_CONTACT_SIZE + B2L_CONTACT_MANIFOLD;
			const float* rec = records + (size_t)i * b2g::kImpulseFloats;
			int pointCount = wide ? 2 : *reinterpret_cast<const int*>( manifold + B2L_MANIFOLD_POINT_COUNT );
			*reinterpret_cast<float*>( manifold + B2L_MANIFOLD_ROLLING_IMPULSE ) = rec[0];
			for ( int j = 0; j < pointCount; ++j )
			{
				uint8_t* mp = manifold + B2L_MANIFOLD_POINTS + j * B2L_MP_SIZE;
				const float* pr = rec + 1 + 4 * j;
				*reinterpret_cast<float*>( mp + B2L_MP_NORMAL_IMPULSE ) = pr[0];
				*reinterpret_cast<float*>( mp + B2L_MP_TANGENT_IMPULSE ) = pr[1];
				*reinterpret_cast<float*>( mp + B2L_MP_TOTAL_NORMAL_IMPULSE ) = pr[2];
				*reinterpret_cast<float*>( mp + B2L_MP_NORMAL_VELOCITY ) = pr[3];
			}
		}
	}
}

static void b2gFinishDownload( b2GpuSolver* s, const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	const b2g::StepParams& P = s->params;
	b2gScatterImpulses( s, d );
	if ( r != nullptr )
	{
		// uint32 pairs are the little-endian halves of the reference's uint64 blocks (src/bitset.h)
		if ( r->hitEventBits != nullptr )
		{
			const uint64_t* src = reinterpret_cast<const uint64_t*>( s->hBits.ptr );
			for ( int i = 0; i < P.hitWords / 2; ++i )
			{
				r->hitEventBits[i] |= src[i];
			}
		}
		if ( r->jointEventBits != nullptr )
		{
			for ( int i = 0; i < P.jointWords / 2; ++i )
			{
				uint64_t word = (uint64_t)s->hBits.ptr[P.hitWords + 2 * i] | ( (uint64_t)s->hBits.ptr[P.hitWords + 2 * i + 1] << 32 );
				r->jointEventBits[i] |= word;
			}
		}
		r->hasHitEvents = s->hControl->hasHitEvents;
	}
}

extern "C" int b2GpuSolverDownload( b2GpuSolver* s, const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	if ( s == nullptr || d == nullptr )
	{
		return b2gFailMsg( "b2GpuSolverDownload: null argument" );
	}
	if ( !s->ran )
	{
		return b2gFailMsg( "b2GpuSolverDownload: nothing has run" );
	}
	B2G_CUDA( cudaSetDevice( s->device ) );
	uint64_t bytes = 0;
	if ( b2gEnqueueDownload( s, d, &bytes ) != 0 )
	{
		return 1;
	}
	B2G_CUDA( cudaStreamSynchronize( s->stream ) );
	b2gFinishDownload( s, d, r );
	if ( r != nullptr )
	{
		r->d2hBytes = bytes;
	}
	return 0;
}

// ---- the whole step --------------------------------------------------------------------------------------------
extern "C" int b2GpuSolverStep( b2GpuSolver* s, const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	using clock = std::chrono::steady_clock;
	auto ms = []( clock::time_point a, clock::time_point b ) { return std::chrono::duration<float, std::milli>( b - a ).count(); };
	auto t0 = clock::now();
	if ( b2GpuSolverUpload( s, d ) != 0 )
	{
		return 1;
	}
	auto t1 = clock::now();
	if ( b2gEnqueueRun( s ) != 0 )
	{
		return 1;
	}
	uint64_t d2h = 0;
	if ( b2gEnqueueDownload( s, d, &d2h ) != 0 )
	{
		return 1;
	}
	B2G_CUDA( cudaStreamSynchronize( s->stream ) );
	auto t2 = clock::now();
	B2G_CUDA( cudaEventElapsedTime( &s->lastKernelMs, s->evStart, s->evStop ) );
	b2gFinishDownload( s, d, r );
	auto t3 = clock::now();
	if ( r != nullptr )
	{
		r->kernelMs = s->lastKernelMs;
		r->kernelLaunches = s->lastLaunches;
		r->h2dBytes = s->lastH2D;
		r->d2hBytes = d2h;
		b2gFillTimers( s, r );
		r->uploadMs = ms( t0, t1 );
		r->waitMs = ms( t1, t2 );
		r->scatterMs = ms( t2, t3 );
		r->h2dMs = 0.0f;
		cudaEventElapsedTime( &r->h2dMs, s->evUpload, s->evStart );
		r->totalMs = ms( t0, clock::now() );
	}
	return 0;
}

// ---- batch of independent worlds (implemented in b2g_batch.cu) ----------------------------------------------------
// see b2g_batch.cu

// =================================================================================================================
// Page-locked host allocator for b2SetAllocator (include/box2d/base.h:86)
// =================================================================================================================
namespace
{

struct PinnedPool
{
	static constexpr int kMinShift = 6;	 // 64 B: every block is at least cache-line aligned
	static constexpr int kMaxShift = 40;
	static constexpr size_t kSlabBytes = size_t( 32 ) << 20;

	std::mutex mutex;
	std::vector<void*> freeLists[kMaxShift + 1];
	std::vector<std::pair<char*, size_t>> slabs;
	char* cursor = nullptr;
	size_t remaining = 0;
	bool pinned = true;

	static int classOf( size_t size )
	{
		int shift = kMinShift;
		while ( ( size_t( 1 ) << shift ) < size )
		{
			shift += 1;
		}
		return shift;
	}

	char* newSlab( size_t bytes )
	{
		void* mem = nullptr;
		if ( pinned )
		{
			if ( cudaHostAlloc( &mem, bytes, cudaHostAllocPortable ) != cudaSuccess )
			{
				cudaGetLastError();
				pinned = false; // no driver: plain memory keeps the host library usable for CPU-only tests
				mem = nullptr;
			}
		}
		if ( mem == nullptr )
		{
			if ( posix_memalign( &mem, 4096, bytes ) != 0 )
			{
				return nullptr;
			}
		}
		slabs.emplace_back( static_cast<char*>( mem ), bytes );
		return static_cast<char*>( mem );
	}

	void* allocate( size_t size )
	{
		int shift = classOf( size );
		size_t bytes = size_t( 1 ) << shift;
		std::lock_guard<std::mutex> lock( mutex );
		std::vector<void*>& list = freeLists[shift];
		if ( !list.empty() )
		{
			void* mem = list.back();
			list.pop_back();
			return mem;
		}
		if ( bytes >= kSlabBytes / 4 )
		{
			return newSlab( bytes ); // big blocks get their own registration
		}
		if ( remaining < bytes )
		{
			cursor = newSlab( kSlabBytes );
			remaining = cursor != nullptr ? kSlabBytes : 0;
			if ( cursor == nullptr )
			{
				return nullptr;
			}
		}
		// keep natural alignment of the size class (up to 4 KiB)
		size_t align = bytes < 4096 ? bytes : 4096;
		size_t misalign = reinterpret_cast<uintptr_t>( cursor ) & ( align - 1 );
		if ( misalign != 0 )
		{
			size_t skip = align - misalign;
			if ( skip + bytes > remaining )
			{
				cursor = newSlab( kSlabBytes );
				remaining = cursor != nullptr ? kSlabBytes : 0;
				if ( cursor == nullptr )
				{
					return nullptr;
				}
			}
			else
			{
				cursor += skip;
				remaining -= skip;
			}
		}
		void* mem = cursor;
		cursor += bytes;
		remaining -= bytes;
		return mem;
	}

	void release( void* mem, size_t size )
	{
		if ( mem == nullptr )
		{
			return;
		}
		int shift = classOf( size );
		std::lock_guard<std::mutex> lock( mutex );
		freeLists[shift].push_back( mem );
	}
};

PinnedPool& pinnedPool()
{
	static PinnedPool* pool = new PinnedPool(); // intentionally leaked: outlives every world
	return *pool;
}

} // namespace

extern "C" void* b2GpuHostAlloc( size_t size, int alignment )
{
	(void)alignment; // blocks are aligned to min(size class, 4096) >= any alignment Box2D asks for (<= 64)
	return pinnedPool().allocate( size == 0 ? 1 : size );
}

extern "C" void b2GpuHostFree( void* mem, size_t size )
{
	pinnedPool().release( mem, size == 0 ? 1 : size );
}
