// d2h_bench.cu -- D2H bandwidth into a pinned buffer depending on what the CPU did to it before the copy.
#include <cuda_runtime.h>
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#define CK( x ) do { cudaError_t e = ( x ); if ( e != cudaSuccess ) { printf( "%s\n", cudaGetErrorString( e ) ); return 1; } } while ( 0 )
static volatile float g_sink;
template <typename F> static void par( int threads, size_t n, F f )
{
	std::vector<std::thread> pool;
	for ( int t = 0; t < threads; ++t ) pool.emplace_back( [=]() { f( n * t / threads & ~size_t( 15 ), n * ( t + 1 ) / threads & ~size_t( 15 ) ); } );
	for ( auto& th : pool ) th.join();
}
int main()
{
	const size_t bytes = 6u << 20, n = bytes / 4;
	float *h, *dev;
	CK( cudaHostAlloc( &h, bytes, cudaHostAllocDefault ) );
	CK( cudaMalloc( &dev, bytes ) );
	CK( cudaMemset( dev, 1, bytes ) );
	cudaStream_t st; CK( cudaStreamCreateWithFlags( &st, cudaStreamNonBlocking ) );
	cudaEvent_t e0, e1; cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
	const char* names[] = { "untouched", "read by 8 threads", "written by 8 threads", "read by 8 threads + clflushopt", "read with prefetchnta", "read by 8 threads, copy twice" };
	for ( int mode = 0; mode < 6; ++mode )
	{
		float best = 1e9f, second = 0, hostMs = 0;
		for ( int rep = 0; rep < 8; ++rep )
		{
			auto t0 = std::chrono::steady_clock::now();
			if ( mode == 1 || mode == 5 ) par( 8, n, [=]( size_t b, size_t e ) { float s = 0; for ( size_t i = b; i < e; ++i ) s += h[i]; g_sink = s; } );
			if ( mode == 2 ) par( 8, n, [=]( size_t b, size_t e ) { for ( size_t i = b; i < e; ++i ) h[i] = (float)i; } );
			if ( mode == 3 ) par( 8, n, [=]( size_t b, size_t e ) { float s = 0; for ( size_t i = b; i < e; ++i ) s += h[i]; g_sink = s; for ( size_t i = b; i < e; i += 16 ) _mm_clflushopt( h + i ); _mm_sfence(); } );
			if ( mode == 4 ) par( 8, n, [=]( size_t b, size_t e ) { float s = 0; for ( size_t i = b; i < e; i += 16 ) { _mm_prefetch( (const char*)( h + i + 64 ), _MM_HINT_NTA ); for ( int k = 0; k < 16; ++k ) s += h[i + k]; } g_sink = s; } );
			hostMs = std::chrono::duration<float, std::milli>( std::chrono::steady_clock::now() - t0 ).count();
			cudaEventRecord( e0, st );
			CK( cudaMemcpyAsync( h, dev, bytes, cudaMemcpyDeviceToHost, st ) );
			cudaEventRecord( e1, st );
			CK( cudaStreamSynchronize( st ) );
			float ms; cudaEventElapsedTime( &ms, e0, e1 );
			if ( ms < best ) best = ms;
			if ( mode == 5 )
			{
				cudaEventRecord( e0, st );
				CK( cudaMemcpyAsync( h, dev, bytes, cudaMemcpyDeviceToHost, st ) );
				cudaEventRecord( e1, st );
				CK( cudaStreamSynchronize( st ) );
				cudaEventElapsedTime( &second, e0, e1 );
			}
		}
		printf( "%-34s D2H %.3f ms = %.1f GB/s  (host pass %.3f ms) %s%.3f\n", names[mode], best, bytes / best / 1e6, hostMs, mode == 5 ? "second copy ms " : "", second );
	}
	return 0;
}
