// h2d_bench.cu -- why is a freshly CPU-written staging buffer slow to upload?  Compares H2D bandwidth of
// (a) an untouched pinned buffer, (b) a pinned buffer rewritten by N host threads right before the copy,
// (c) a write-combined pinned buffer rewritten the same way.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
#define CK( x ) do { cudaError_t e = ( x ); if ( e != cudaSuccess ) { printf( "%s\n", cudaGetErrorString( e ) ); return 1; } } while ( 0 )

static void fill( float* p, size_t n, int threads, float v )
{
	std::vector<std::thread> pool;
	for ( int t = 0; t < threads; ++t )
		pool.emplace_back( [=]() { size_t b = n * t / threads, e = n * ( t + 1 ) / threads; for ( size_t i = b; i < e; ++i ) p[i] = v + (float)i; } );
	for ( auto& th : pool ) th.join();
}

static void fillNT( float* p, size_t n, int threads, float v )
{
	std::vector<std::thread> pool;
	for ( int t = 0; t < threads; ++t )
		pool.emplace_back( [=]() { size_t b = ( n * t / threads ) & ~size_t( 15 ), e = ( n * ( t + 1 ) / threads ) & ~size_t( 15 );
			for ( size_t i = b; i < e; i += 4 ) _mm_stream_ps( p + i, _mm_set_ps( v + i + 3, v + i + 2, v + i + 1, v + i ) );
			_mm_sfence(); } );
	for ( auto& th : pool ) th.join();
}

static void fillFlush( float* p, size_t n, int threads, float v )
{
	std::vector<std::thread> pool;
	for ( int t = 0; t < threads; ++t )
		pool.emplace_back( [=]() { size_t b = ( n * t / threads ) & ~size_t( 15 ), e = ( n * ( t + 1 ) / threads ) & ~size_t( 15 );
			for ( size_t i = b; i < e; ++i ) p[i] = v + (float)i;
			for ( size_t i = b; i < e; i += 16 ) _mm_clflush( p + i );
			_mm_sfence(); } );
	for ( auto& th : pool ) th.join();
}

int main()
{
	const size_t bytes = 8u << 20;
	float *hDef, *hWc, *dev;
	CK( cudaHostAlloc( &hDef, bytes, cudaHostAllocDefault ) );
	CK( cudaHostAlloc( &hWc, bytes, cudaHostAllocWriteCombined ) );
	CK( cudaMalloc( &dev, bytes ) );
	cudaStream_t st;
	CK( cudaStreamCreateWithFlags( &st, cudaStreamNonBlocking ) );
	cudaEvent_t e0, e1;
	cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
	memset( hDef, 1, bytes );
	for ( int mode = 0; mode < 9; ++mode )
	{
		float best = 1e9f, fillMs = 0;
		for ( int rep = 0; rep < 10; ++rep )
		{
			float* src = ( mode == 3 || mode == 4 ) ? hWc : hDef;
			auto t0 = std::chrono::steady_clock::now();
			if ( mode == 1 ) fill( hDef, bytes / 4, 1, (float)rep );
			if ( mode == 2 ) fill( hDef, bytes / 4, 8, (float)rep );
			if ( mode == 3 ) fill( hWc, bytes / 4, 1, (float)rep );
			if ( mode == 4 ) fill( hWc, bytes / 4, 8, (float)rep );
			if ( mode == 5 ) fillNT( hDef, bytes / 4, 1, (float)rep );
			if ( mode == 6 ) fillNT( hDef, bytes / 4, 8, (float)rep );
			if ( mode == 7 ) fillFlush( hDef, bytes / 4, 8, (float)rep );
			if ( mode == 8 ) { fill( hDef, bytes / 4, 8, (float)rep ); std::this_thread::sleep_for( std::chrono::milliseconds( 5 ) ); }
			fillMs = std::chrono::duration<float, std::milli>( std::chrono::steady_clock::now() - t0 ).count();
			cudaEventRecord( e0, st );
			CK( cudaMemcpyAsync( dev, src, bytes, cudaMemcpyHostToDevice, st ) );
			cudaEventRecord( e1, st );
			CK( cudaStreamSynchronize( st ) );
			float ms; cudaEventElapsedTime( &ms, e0, e1 );
			if ( ms < best ) best = ms;
		}
		const char* names[] = { "pinned, untouched", "pinned, rewritten by 1 thread", "pinned, rewritten by 8 threads", "write-combined, 1 thread", "write-combined, 8 threads", "pinned, NT stores 1 thread", "pinned, NT stores 8 threads", "pinned, 8 threads + clflush", "pinned, 8 threads, copy 5 ms later" };
		printf( "%-34s H2D %.3f ms = %.1f GB/s   (host fill %.3f ms)\n", names[mode], best, bytes / best / 1e6, fillMs );
	}
	return 0;
}
