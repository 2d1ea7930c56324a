// b2g_alloc.cu -- page-locked host allocator for b2SetAllocator (reference include/box2d/base.h:86).
#include "b2_gpu_solver.h"

#include <cuda_runtime.h>

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
			if ( cudaHostAlloc( &mem, bytes, cudaHostAllocPortable ) != cudaSuccess )
			{
				cudaGetLastError();
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
		if ( pinned && cudaHostAlloc( &mem, bytes, cudaHostAllocPortable ) == cudaSuccess )
		{
			locked = true;
		}
		else
		{
			cudaGetLastError();
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
				cudaFreeHost( mem );
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
