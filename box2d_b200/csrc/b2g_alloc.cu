// b2g_alloc.cu -- page-locked host allocator for b2SetAllocator (reference include/box2d/base.h:86).
#include "b2_gpu_solver.h"
#include "b2g_host.h"

#include <cuda_runtime.h>
#include <sys/mman.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

// =================================================================================================================
// Page-locked host allocator for b2SetAllocator (include/box2d/base.h:86)
// =================================================================================================================
// Page-locked host memory on transparent huge pages, where the platform grants them (B2GPU_HUGE_PAGES: 2 = here and for
// the library's own arrays, the default; 1 = only the latter, b2gHugeVector; 0 = neither): 2 MB
// aligned anonymous memory advised MADV_HUGEPAGE and then registered with the driver, instead of cudaHostAlloc (whose
// pages are 4 KB ones).  The host passes over the reference's arrays and the staging arenas are streaming loops over
// tens of megabytes; on a guest every TLB miss of theirs is a two-dimensional page walk.  The device address of such a
// block must be its host address (the kernels store results through it, b2g_solver.cu "direct outputs"): where the
// driver says otherwise, or refuses the registration, the block comes from cudaHostAlloc as before.
namespace
{
std::mutex g_registeredMutex;
std::unordered_map<void*, size_t> g_registered; // blocks of b2gPinnedAlloc that are registrations (the others are cudaHostAlloc's)

} // namespace

int b2gHugePagesWanted()
{
	static const int wanted = getenv( "B2GPU_HUGE_PAGES" ) == nullptr ? 2 : atoi( getenv( "B2GPU_HUGE_PAGES" ) );
	return wanted;
}

void* b2gPinnedAlloc( size_t bytes, unsigned int flags )
{
	const size_t page = size_t( 2 ) << 20;
	if ( b2gHugePagesWanted() >= 2 && bytes >= page / 2 )
	{
		size_t size = ( bytes + page - 1 ) / page * page;
		void* mem = aligned_alloc( page, size );
		if ( mem != nullptr )
		{
			madvise( mem, size, MADV_HUGEPAGE );
			void* device = nullptr;
			if ( cudaHostRegister( mem, size, cudaHostRegisterPortable | cudaHostRegisterMapped ) == cudaSuccess )
			{
				if ( cudaHostGetDevicePointer( &device, mem, 0 ) == cudaSuccess && device == mem )
				{
					std::lock_guard<std::mutex> lock( g_registeredMutex );
					g_registered[mem] = size;
					return mem;
				}
				cudaHostUnregister( mem );
			}
			cudaGetLastError();
			free( mem );
		}
	}
	void* mem = nullptr;
	if ( cudaHostAlloc( &mem, bytes, flags ) != cudaSuccess )
	{
		cudaGetLastError();
		return nullptr;
	}
	return mem;
}

void b2gPinnedFree( void* mem )
{
	if ( mem == nullptr )
	{
		return;
	}
	{
		std::lock_guard<std::mutex> lock( g_registeredMutex );
		auto it = g_registered.find( mem );
		if ( it != g_registered.end() )
		{
			g_registered.erase( it );
			cudaHostUnregister( mem );
			free( mem );
			return;
		}
	}
	cudaFreeHost( mem );
}

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
	// blocks of kSlabBytes / 4 and more are registrations of their own, sized to whole 2 MiB pages (not to a power of two: a
	// 33 MB array would page-lock 64 MB) and given back when they are released -- Box2D's arrays grow geometrically, every
	// outgrown capacity would otherwise stay page-locked for good
	std::unordered_map<void*, std::pair<size_t, bool>> bigBlocks; // -> { bytes, pinned }
	bool warned = false;
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
			mem = b2gPinnedAlloc( bytes, cudaHostAllocPortable );
			if ( mem == nullptr )
			{
				pinned = false; // no driver: plain memory keeps the host library usable for CPU-only tests
				mem = nullptr;
				if ( !warned )
				{
					warned = true;
					fprintf( stderr, "box2d_b200: page-locked host memory is not available (%zu bytes asked); the reference's arrays "
									 "stay in pageable memory from here on (uploads still go through the library's own staging)\n", bytes );
				}
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

	void* allocateBig( size_t size )
	{
		const size_t page = size_t( 2 ) << 20;
		size_t bytes = ( size + page - 1 ) / page * page;
		void* mem = nullptr;
		bool locked = false;
		if ( pinned && ( mem = b2gPinnedAlloc( bytes, cudaHostAllocPortable ) ) != nullptr )
		{
			locked = true;
		}
		else
		{
			mem = nullptr;
			if ( posix_memalign( &mem, 4096, bytes ) != 0 )
			{
				return nullptr;
			}
		}
		bigBlocks[mem] = { bytes, locked };
		return mem;
	}

	void* allocate( size_t size )
	{
		int shift = classOf( size );
		size_t bytes = size_t( 1 ) << shift;
		std::lock_guard<std::mutex> lock( mutex );
		if ( size >= kSlabBytes / 4 )
		{
			return allocateBig( size );
		}
		std::vector<void*>& list = freeLists[shift];
		if ( !list.empty() )
		{
			void* mem = list.back();
			list.pop_back();
			return mem;
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
		auto big = bigBlocks.find( mem );
		if ( big != bigBlocks.end() )
		{
			if ( big->second.second )
			{
				b2gPinnedFree( mem );
			}
			else
			{
				free( mem );
			}
			bigBlocks.erase( big );
			return;
		}
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
