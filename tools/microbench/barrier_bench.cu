// barrier_bench.cu -- how fast can colour-to-colour synchronisation be on a B200?
// Measures ns per barrier for several grid-wide barrier designs and grid sizes, with and without a "stage-like"
// body (dependent index load -> gather -> a little math -> scatter), to choose the step kernel's barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu && ./barrier_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK( x )                                                                                                                  \
	do                                                                                                                           \
	{                                                                                                                            \
		cudaError_t e = ( x );                                                                                                   \
		if ( e != cudaSuccess )                                                                                                  \
		{                                                                                                                        \
			printf( "CUDA error %s at %s:%d\n", cudaGetErrorString( e ), __FILE__, __LINE__ );                                   \
			exit( 1 );                                                                                                           \
		}                                                                                                                        \
	}                                                                                                                            \
	while ( 0 )

struct Args
{
	unsigned* counter; // [0] arrivals
	unsigned* flags;   // one per block, 32-byte apart
	int* idx;
	float4* state;
	int items;
	int rounds;
	int body;
};

__device__ __forceinline__ void stageBody( const Args& a, int round )
{
	if ( a.body == 0 )
	{
		return;
	}
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if ( t < a.items )
	{
		int i = a.idx[( t + round * 7919 ) % a.items];
		float4 v = __ldcg( a.state + i );
		v.x = v.x * 1.0001f + v.y;
		v.y = v.y * 0.9999f + v.z;
		v.z = v.z + v.x * 0.5f;
		__stcg( a.state + i, v );
	}
}

// A: release-add + acquire-load spin on one counter
__device__ __forceinline__ void barrierA( unsigned* counter, unsigned target )
{
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		asm volatile( "red.release.gpu.global.add.u32 [%0], 1;" ::"l"( counter ) : "memory" );
		unsigned seen;
		do
		{
			asm volatile( "ld.acquire.gpu.global.u32 %0, [%1];" : "=r"( seen ) : "l"( counter ) : "memory" );
		}
		while ( seen < target );
	}
	__syncthreads();
}

// B: fence + relaxed atomic + relaxed spin + fence (cooperative-groups style)
__device__ __forceinline__ void barrierB( unsigned* counter, unsigned target )
{
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		__threadfence();
		atomicAdd( counter, 1u );
		while ( *( (volatile unsigned*)counter ) < target )
		{
		}
		__threadfence();
	}
	__syncthreads();
}

// C: one flag per block (no atomics); warp 0 polls every flag
__device__ __forceinline__ void barrierC( unsigned* flags, unsigned epoch )
{
	__syncthreads();
	if ( threadIdx.x < 32 )
	{
		if ( threadIdx.x == 0 )
		{
			asm volatile( "st.release.gpu.global.u32 [%0], %1;" ::"l"( flags + blockIdx.x * 8 ), "r"( epoch ) : "memory" );
		}
		for ( unsigned b = threadIdx.x; b < gridDim.x; b += 32 )
		{
			unsigned seen;
			do
			{
				asm volatile( "ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"( seen ) : "l"( flags + b * 8 ) : "memory" );
			}
			while ( seen < epoch );
		}
		__syncwarp();
		asm volatile( "fence.acq_rel.gpu;" ::: "memory" );
	}
	__syncthreads();
}

// D: relaxed spin, single acquire fence at the end (polls do not invalidate L1 each iteration)
__device__ __forceinline__ void barrierD( unsigned* counter, unsigned target )
{
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		asm volatile( "red.release.gpu.global.add.u32 [%0], 1;" ::"l"( counter ) : "memory" );
		unsigned seen;
		do
		{
			asm volatile( "ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"( seen ) : "l"( counter ) : "memory" );
		}
		while ( seen < target );
		asm volatile( "fence.acq_rel.gpu;" ::: "memory" );
	}
	__syncthreads();
}

template <int KIND> __global__ void __launch_bounds__( 256, 1 ) kernelGrid( Args a )
{
	for ( int r = 1; r <= a.rounds; ++r )
	{
		stageBody( a, r );
		if ( KIND == 0 )
			barrierA( a.counter, (unsigned)r * gridDim.x );
		else if ( KIND == 1 )
			barrierB( a.counter, (unsigned)r * gridDim.x );
		else if ( KIND == 2 )
			barrierC( a.flags, (unsigned)r );
		else if ( KIND == 3 )
			barrierD( a.counter, (unsigned)r * gridDim.x );
		else
			cg::this_grid().sync();
	}
}

// E: cluster barrier (+ global barrier between cluster leaders when there is more than one cluster)
__global__ void __launch_bounds__( 256, 1 ) kernelCluster( Args a, int clusters )
{
	cg::cluster_group cluster = cg::this_cluster();
	unsigned rank = cluster.block_rank();
	for ( int r = 1; r <= a.rounds; ++r )
	{
		stageBody( a, r );
		if ( clusters == 1 )
		{
			asm volatile( "barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory" );
		}
		else
		{
			asm volatile( "barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory" );
			if ( rank == 0 && threadIdx.x == 0 )
			{
				asm volatile( "red.release.gpu.global.add.u32 [%0], 1;" ::"l"( a.counter ) : "memory" );
				unsigned seen;
				do
				{
					asm volatile( "ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"( seen ) : "l"( a.counter ) : "memory" );
				}
				while ( seen < (unsigned)r * clusters );
				asm volatile( "fence.acq_rel.gpu;" ::: "memory" );
			}
			asm volatile( "barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory" );
		}
	}
}

static float timeKernel( void ( *launch )( Args, int, int ), Args a, int grid, int extra )
{
	cudaEvent_t e0, e1;
	CK( cudaEventCreate( &e0 ) );
	CK( cudaEventCreate( &e1 ) );
	float best = 1e30f;
	for ( int rep = 0; rep < 5; ++rep )
	{
		CK( cudaMemset( a.counter, 0, 64 ) );
		CK( cudaMemset( a.flags, 0, 148 * 32 + 64 ) );
		CK( cudaEventRecord( e0 ) );
		launch( a, grid, extra );
		CK( cudaEventRecord( e1 ) );
		CK( cudaEventSynchronize( e1 ) );
		CK( cudaGetLastError() );
		float ms;
		CK( cudaEventElapsedTime( &ms, e0, e1 ) );
		if ( ms < best )
		{
			best = ms;
		}
	}
	return best;
}

template <int KIND> static void launchGrid( Args a, int grid, int )
{
	void* args[] = { &a };
	CK( cudaLaunchCooperativeKernel( (const void*)kernelGrid<KIND>, dim3( grid ), dim3( 256 ), args, 0, 0 ) );
}

static void launchCluster( Args a, int grid, int clusterSize )
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3( grid );
	cfg.blockDim = dim3( 256 );
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = clusterSize;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	int clusters = grid / clusterSize;
	CK( cudaLaunchKernelEx( &cfg, kernelCluster, a, clusters ) );
}

int main()
{
	cudaDeviceProp prop;
	CK( cudaGetDeviceProperties( &prop, 0 ) );
	printf( "%s, %d SMs, %d MHz\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1000 );
	CK( cudaFuncSetAttribute( kernelCluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1 ) );

	Args a;
	a.rounds = 2000;
	a.items = 1 << 15;
	CK( cudaMalloc( &a.counter, 64 ) );
	CK( cudaMalloc( &a.flags, 148 * 32 + 64 ) );
	CK( cudaMalloc( &a.idx, a.items * sizeof( int ) ) );
	CK( cudaMalloc( &a.state, a.items * sizeof( float4 ) ) );
	std::vector<int> idx( a.items );
	for ( int i = 0; i < a.items; ++i )
	{
		idx[i] = (int)( ( (long long)i * 40503 ) % a.items );
	}
	CK( cudaMemcpy( a.idx, idx.data(), a.items * sizeof( int ), cudaMemcpyHostToDevice ) );
	CK( cudaMemset( a.state, 0, a.items * sizeof( float4 ) ) );

	const char* names[] = { "A release-add + acquire spin", "B fence+atomic+volatile spin+fence", "C per-block flags, warp polls",
							"D release-add + relaxed spin + fence", "cg grid.sync" };
	int grids[] = { 8, 16, 37, 74, 148 };
	for ( int body = 0; body <= 1; ++body )
	{
		a.body = body;
		printf( "\n== %s ==  (ns per barrier, %d rounds)\n", body ? "with stage-like body" : "empty stages", a.rounds );
		printf( "%-40s", "grid blocks:" );
		for ( int g : grids )
			printf( "%8d", g );
		printf( "\n" );
		for ( int k = 0; k < 5; ++k )
		{
			printf( "%-40s", names[k] );
			for ( int g : grids )
			{
				float ms = 0;
				switch ( k )
				{
					case 0: ms = timeKernel( launchGrid<0>, a, g, 0 ); break;
					case 1: ms = timeKernel( launchGrid<1>, a, g, 0 ); break;
					case 2: ms = timeKernel( launchGrid<2>, a, g, 0 ); break;
					case 3: ms = timeKernel( launchGrid<3>, a, g, 0 ); break;
					default: ms = timeKernel( launchGrid<4>, a, g, 0 ); break;
				}
				printf( "%8.0f", ms * 1e6f / a.rounds );
			}
			printf( "\n" );
		}
		int csizes[] = { 2, 4, 8, 16 };
		for ( int cs : csizes )
		{
			printf( "single cluster of %-22d", cs );
			float ms = timeKernel( launchCluster, a, cs, cs );
			printf( "%8.0f\n", ms * 1e6f / a.rounds );
		}
		for ( int cs : { 4, 8 } )
		{
			int g = ( 148 / cs ) * cs;
			printf( "clusters of %d x %d + leader barrier   ", cs, g / cs );
			float ms = timeKernel( launchCluster, a, g, cs );
			printf( "%8.0f\n", ms * 1e6f / a.rounds );
		}
	}
	return 0;
}
