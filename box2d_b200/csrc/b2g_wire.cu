// b2g_wire.cu -- the two memory-bound host passes of a step and their PCIe transfers.
//
// Pack: the reference's arrays -> the wire format in the page-locked input arena (non-temporal stores); unpack: the
// output arena -> the reference's arrays in place.  Both run on however many host threads the caller brings
// (b2GpuSolverPackWork / UnpackWork): blocks of items are claimed in order, one caller pumps the transfers.
#include "b2g_host.h"

#include "b2gpu_layout.h"

#include <immintrin.h>
#if defined( __x86_64__ )
#include <cpuid.h>
#endif

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static int* b2gJointIndexPair( b2lJointSim* joint )
{
	switch ( joint->type )
	{
		case b2l_distanceJoint:
			return &joint->u.distance.indexA;
		case b2l_motorJoint:
			return &joint->u.motor.indexA;
		case b2l_moverJoint:
			return &joint->u.mover.indexA;
		case b2l_pogoJoint:
			return &joint->u.pogo.indexA;
		case b2l_prismaticJoint:
			return &joint->u.prismatic.indexA;
		case b2l_revoluteJoint:
			return &joint->u.revolute.indexA;
		case b2l_weldJoint:
			return &joint->u.weld.indexA;
		case b2l_wheelJoint:
			return &joint->u.wheel.indexA;
		default:
			return nullptr;
	}
}


// ---- phase 2: pack the reference's arrays into the wire format (callable concurrently on disjoint ranges) ---------
static inline float b2gRdF( const uint8_t* p, int offset )
{
	float v;
	memcpy( &v, p + offset, 4 );
	return v;
}

static inline float b2gIntBits( int v )
{
	float f;
	memcpy( &f, &v, 4 );
	return f;
}

static inline int b2gRdI( const uint8_t* p, int offset )
{
	int v;
	memcpy( &v, p + offset, 4 );
	return v;
}

// non-temporal 16-byte stores: the staging buffer must not stay dirty in the CPU caches (see b2GpuSolver::hWire)
static inline void b2gStream4( float4* dst, float a, float b, float c, float d )
{
	_mm_stream_ps( reinterpret_cast<float*>( dst ), _mm_set_ps( d, c, b, a ) );
}

static inline void b2gStreamCopy( float4* dst, const uint8_t* src, int quads )
{
	for ( int q = 0; q < quads; ++q )
	{
		_mm_stream_si128( reinterpret_cast<__m128i*>( dst + q ), _mm_loadu_si128( reinterpret_cast<const __m128i*>( src + 16 * q ) ) );
	}
}

// A pack block's window into one of the two variable-length streams of resident mode (full contact records, dirty
// bodies): whole chunks of kStreamChunk records are reserved from the shared cursor, so blocks packed concurrently never
// contend for a record and a stream is one dense prefix [0, cursor) when the packing is done.
struct b2gStreamChunk
{
	std::atomic<int>* cursor;
	int capacity;
	int next, end;
	std::atomic<int>* failed;
	int taken = 0;
};

static inline int b2gStreamTake( b2gStreamChunk& chunk )
{
	if ( chunk.next == chunk.end )
	{
		chunk.next = chunk.cursor->fetch_add( kStreamChunk, std::memory_order_relaxed );
		chunk.end = chunk.next + kStreamChunk;
		if ( chunk.end > chunk.capacity )
		{
			// only a caller that packs in ranges far smaller than the blocks b2gBegin planned for can get here (the streams
			// hold every record plus a chunk per planned block): the step fails at Submit
			chunk.failed->store( 1, std::memory_order_relaxed );
			chunk.next = 0;
			chunk.end = kStreamChunk;
		}
	}
	chunk.taken += 1;
	return chunk.next++;
}

void b2gFlushLines( const void* ptr, size_t bytes );

// Deferred impulses: write the pending record `slot` into the manifold of `sim` -- what b2StoreImpulsesTask
// (src/contact_solver.c:2293-2320: both points of a coloured contact) and b2StoreImpulses_Overflow (:526-542) would have
// written at the end of the pending step.  Returns the record's hit-event flag.
static inline bool b2gMaterializeRecord( const b2GpuSolver* s, uint8_t* sim, int slot, bool wide )
{
	const float* rec = s->pendingRecords + (size_t)slot * b2g::kImpulseFloats;
	uint8_t* manifold = sim + B2L_CONTACT_MANIFOLD;
	int pointCount = wide ? 2 : *reinterpret_cast<const int*>( manifold + B2L_MANIFOLD_POINT_COUNT );
	memcpy( manifold + B2L_MANIFOLD_ROLLING_IMPULSE, rec + 0, 4 );
	for ( int j = 0; j < pointCount; ++j )
	{
		// normalImpulse, tangentImpulse, totalNormalImpulse, normalVelocity are contiguous (collision.h:549-561)
		memcpy( manifold + B2L_MANIFOLD_POINTS + j * B2L_MP_SIZE + B2L_MP_NORMAL_IMPULSE, rec + 1 + 4 * j, 16 );
	}
	// (The lines of the arena that are read here stay in this core's cache, and the arena is a DMA target again two steps
	// from now -- slower into cached lines, see b2gFlushLines.  Nobody waits for that transfer, and flushing record by record
	// throws out the line the neighbouring record is in: measured on the rain scene, +0.2 ms of narrow phase.)
	return rec[9] != 0.0f;
}

// The contact `sim` sits at place `index` of the colour with home key `key`: if that is where it was when the pending step was
// solved and its record has not been used yet, the record goes into its manifold.  Returns 1 (materialized) or 0; *hit = the
// record's hit-event flag.  Concurrent calls must be for different places.
static inline int b2gMaterializeAt( b2GpuSolver* s, int key, int index, uint8_t* sim, bool* hit )
{
	if ( index >= s->homeCount[key] )
	{
		return 0; // no such place in the pending step
	}
	const int home = s->homeBase[key] + index;
	if ( s->consumedStamp[(size_t)home] == s->deferStamp ||
		 s->shadowHeads[(size_t)home].contactId != *reinterpret_cast<const int*>( sim + B2L_CONTACT_ID ) )
	{
		return 0; // used, or another contact's place
	}
	*hit = b2gMaterializeRecord( s, sim, s->homeSlot[key] + index, key != kHomeColors - 1 );
	s->consumedStamp[(size_t)home] = s->deferStamp;
	return 1;
}

static inline void b2gNoteHit( b2GpuStepResult* result, const uint8_t* sim )
{
	if ( result != nullptr )
	{
		if ( result->hitEventBits != nullptr )
		{
			uint32_t id = (uint32_t) * reinterpret_cast<const int*>( sim + B2L_CONTACT_ID );
			__atomic_fetch_or( result->hitEventBits + ( id >> 6 ), (uint64_t)1 << ( id & 63u ), __ATOMIC_RELAXED );
		}
		__atomic_store_n( &result->hasHitEvents, 1, __ATOMIC_RELAXED );
	}
}

// home key of a caller's graph colour index (b2GpuColorDesc::colorIndex), or -1
static inline int b2gHomeKeyOf( const b2GpuSolver* s, int colorIndex )
{
	return s->homesOrdered && colorIndex >= 0 && colorIndex < kHomeColors ? colorIndex : -1;
}

extern "C" int b2GpuSolverMaterializeContacts( b2GpuSolver* s, int colorIndex, int firstIndex, void* contactSims, int count,
											   b2GpuStepResult* result )
{
	if ( s == nullptr || !s->deferPending || count <= 0 || firstIndex < 0 )
	{
		return 0;
	}
	const int key = b2gHomeKeyOf( s, colorIndex );
	if ( key < 0 )
	{
		return 0;
	}
	if ( b2gDeferSync( s ) != 0 )
	{
		return -1;
	}
	uint8_t* sims = static_cast<uint8_t*>( contactSims );
	int done = 0;
	for ( int i = 0; i < count; ++i )
	{
		uint8_t* sim = sims + (size_t)i * B2L_CONTACT_SIZE;
		bool hit = false;
		if ( b2gMaterializeAt( s, key, firstIndex + i, sim, &hit ) != 0 )
		{
			done += 1;
			if ( hit )
			{
				b2gNoteHit( result, sim );
			}
		}
	}
	if ( done > 0 && count > 1 )
	{
		// (statistics; not for the single contacts the narrow phase's workers materialize: one shared counter)
		s->materialized.fetch_add( done, std::memory_order_relaxed );
	}
	return done;
}

extern "C" void b2GpuSolverDeferredForget( b2GpuSolver* s, int colorIndex, int index )
{
	if ( s == nullptr || !s->deferPending || index < 0 )
	{
		return;
	}
	const int key = b2gHomeKeyOf( s, colorIndex );
	if ( key >= 0 && index < s->homeCount[key] )
	{
		s->consumedStamp[(size_t)( s->homeBase[key] + index )] = s->deferStamp;
	}
}

// The joint `sim` sits at place `index` of the colour with home key `key`: if that is where it was when the pending step was
// solved and its output record has not been used yet, the fields the stages wrote (b2lJointMutableRuns: the accumulated
// impulses, ...) go into the b2JointSim -- and into the shadow of what the device's table holds, like the unpack pass does.
static inline int b2gMaterializeJointAt( b2GpuSolver* s, int key, int index, uint8_t* sim )
{
	if ( index >= s->jointHomeCount[key] )
	{
		return 0;
	}
	const int home = s->jointHomeBase[key] + index;
	uint8_t* shadow = s->shadowJoints.data() + (size_t)home * b2g::kJointStride;
	if ( s->consumedJointStamp[(size_t)home] == s->deferStamp ||
		 *reinterpret_cast<const int*>( shadow + offsetof( b2lJointSim, jointId ) ) != *reinterpret_cast<const int*>( sim + offsetof( b2lJointSim, jointId ) ) )
	{
		return 0; // used, or another joint's place
	}
	const float* record = s->pendingJointRecords + (size_t)( s->jointHomeSlot[key] + index ) * B2L_JOINT_OUT_FLOATS;
	int offsets[2], floats[2];
	int runs = b2lJointMutableRuns( *reinterpret_cast<const int*>( sim + offsetof( b2lJointSim, type ) ), offsets, floats );
	for ( int r = 0; r < runs; ++r )
	{
		memcpy( sim + offsets[r], record, (size_t)floats[r] * sizeof( float ) );
		memcpy( shadow + offsets[r], record, (size_t)floats[r] * sizeof( float ) );
		record += floats[r];
	}
	s->consumedJointStamp[(size_t)home] = s->deferStamp;
	return 1;
}

extern "C" int b2GpuSolverMaterializeJoints( b2GpuSolver* s, int colorIndex, int firstIndex, void* jointSims, int count )
{
	if ( s == nullptr || !s->deferJointsPending || count <= 0 || firstIndex < 0 )
	{
		return 0;
	}
	const int key = s->jointHomesOrdered && colorIndex >= 0 && colorIndex < kHomeColors ? colorIndex : -1;
	if ( key < 0 )
	{
		return 0;
	}
	if ( b2gDeferSync( s ) != 0 )
	{
		return -1;
	}
	uint8_t* sims = static_cast<uint8_t*>( jointSims );
	int done = 0;
	for ( int i = 0; i < count; ++i )
	{
		done += b2gMaterializeJointAt( s, key, firstIndex + i, sims + (size_t)i * B2L_JOINT_SIZE );
	}
	return done;
}

extern "C" void b2GpuSolverDeferredForgetJoint( b2GpuSolver* s, int colorIndex, int index )
{
	if ( s == nullptr || !s->deferJointsPending || index < 0 || !s->jointHomesOrdered || colorIndex < 0 || colorIndex >= kHomeColors )
	{
		return;
	}
	if ( index < s->jointHomeCount[colorIndex] )
	{
		s->consumedJointStamp[(size_t)( s->jointHomeBase[colorIndex] + index )] = s->deferStamp;
	}
}

// every pending record goes into its manifold, the contacts found in the arrays of the step that is being laid out (or has
// just ended): s->contactSegs
int b2gMaterializePendingFromSegs( b2GpuSolver* s, b2GpuStepResult* results )
{
	if ( !s->deferPending )
	{
		return 0;
	}
	if ( b2gDeferSync( s ) != 0 )
	{
		return 1;
	}
	int done = 0;
	if ( s->deferJointsPending && s->jointHomesOrdered )
	{
		for ( const b2gJointSeg& seg : s->jointSegs )
		{
			const int key = seg.overflow ? kHomeColors - 1 : seg.colorIndex;
			if ( key < 0 || key >= kHomeColors || ( !seg.overflow && key == kHomeColors - 1 ) )
			{
				continue;
			}
			for ( int i = 0; i < seg.count; ++i )
			{
				b2gMaterializeJointAt( s, key, i, seg.sims + (size_t)i * B2L_JOINT_SIZE );
			}
		}
	}
	s->deferJointsPending = false;
	for ( const b2gContactSeg& seg : s->contactSegs )
	{
		const int key = seg.wide ? b2gHomeKeyOf( s, seg.colorIndex ) : kHomeColors - 1;
		if ( key < 0 || ( seg.wide && key == kHomeColors - 1 ) )
		{
			continue;
		}
		for ( int i = 0; i < seg.count; ++i )
		{
			bool hit = false;
			uint8_t* sim = seg.sims + (size_t)i * B2L_CONTACT_SIZE;
			if ( b2gMaterializeAt( s, key, i, sim, &hit ) != 0 )
			{
				done += 1;
				if ( hit )
				{
					b2gNoteHit( results, sim );
				}
			}
		}
	}
	s->materialized.fetch_add( done, std::memory_order_relaxed );
	s->deferPending = false;
	return 0;
}

static const bool kPollAll = getenv( "B2GPU_POLL_ALL" ) != nullptr; // (experiment)
static const int kPackPrefetch = []() {
	const char* v = getenv( "B2GPU_PACK_PREFETCH" );
	return v != nullptr ? atoi( v ) : 16;
}();

extern "C" void b2GpuSolverPackRange( b2GpuSolver* s, int begin, int end )
{
	int bodyCount = s->params.bodyCount;
	float4* base = s->hWire.ptr;

	// ---- bodies: the state as is + the 32 of b2BodySim's 96 bytes that integrate-velocities reads (src/solver.c:94-102).
	// Resident mode: only the bodies whose state or constants differ from what the device holds (the shadows) travel, as
	// records of the dirty-body stream.
	{
		float4* wireStates = base + s->inStates;
		float4* wireBody = base + s->inBody;
		int* wireBins = reinterpret_cast<int*>( base + s->inBins );
		const bool resident = s->resident;
		const int known = s->cacheUsable ? s->shadowBodyCount : 0;
		b2gStreamChunk dirty = { &s->dirtyCursor, s->dirtyCapacity, 0, 0, &s->streamOverflow };
		bool binMoved = false;
		int i = begin;
		int bodyEnd = end < bodyCount ? end : bodyCount;
		int w = i < bodyEnd ? b2gFindSegment( s->bodyStart, i ) : 0;
		while ( i < bodyEnd )
		{
			const b2gBodySeg& seg = s->bodySegs[w];
			int segEnd = seg.base + seg.count < bodyEnd ? seg.base + seg.count : bodyEnd;
			for ( ; i < segEnd; ++i )
			{
				int local = i - seg.base;
				if ( s->islandMode )
				{
					// a label out of range (a body without an island has no constraints) goes to the world's first island:
					// harmless for the result, and if it overfills a bin the device notices (binFail)
					int label = seg.islandCount == 1 ? 0 : seg.islands[local];
					label = label >= 0 && label < seg.islandCount ? label : 0;
					const int bin = s->islandBin[seg.islandBase + label];
					_mm_stream_si32( wireBins + i, bin );
					if ( s->prevBins[(size_t)i] != bin )
					{
						// (the bins' lists of the previous step cannot serve this one, b2gEnqueueRun)
						s->prevBins[(size_t)i] = bin;
						binMoved = true;
					}
				}
				const uint8_t* state = seg.states + (size_t)local * B2L_STATE_SIZE;
				const uint8_t* sim = seg.sims + (size_t)local * B2L_SIM_SIZE;
				alignas( 16 ) float constants[8] = { b2gRdF( sim, B2L_SIM_INV_MASS ),		 b2gRdF( sim, B2L_SIM_INV_INERTIA ),
													 b2gRdF( sim, B2L_SIM_FORCE ),			 b2gRdF( sim, B2L_SIM_FORCE + 4 ),
													 b2gRdF( sim, B2L_SIM_TORQUE ),			 b2gRdF( sim, B2L_SIM_LINEAR_DAMPING ),
													 b2gRdF( sim, B2L_SIM_ANGULAR_DAMPING ), b2gRdF( sim, B2L_SIM_GRAVITY_SCALE ) };
				if ( !resident )
				{
					b2gStreamCopy( wireStates + 2 * (size_t)i, state, 2 );
					b2gStreamCopy( wireBody + 2 * (size_t)i, reinterpret_cast<const uint8_t*>( constants ), 2 );
					continue;
				}
				float4* shadowState = s->shadowStates.data() + 2 * (size_t)i;
				float4* shadowBody = s->shadowBody.data() + 2 * (size_t)i;
				if ( i < known && memcmp( shadowState, state, B2L_STATE_SIZE ) == 0 && memcmp( shadowBody, constants, 32 ) == 0 )
				{
					continue; // the device has exactly this
				}
				memcpy( shadowBody, constants, 32 );
				float4* record = s->hDirty.ptr + (size_t)b2gStreamTake( dirty ) * b2g::kDirtyBodyQuads;
				b2gStream4( record, b2gIntBits( i ), 0.0f, 0.0f, 0.0f );
				b2gStreamCopy( record + 1, state, 2 );
				b2gStreamCopy( record + 3, reinterpret_cast<const uint8_t*>( constants ), 2 );
			}
			w += 1;
		}
		if ( dirty.taken > 0 )
		{
			s->dirtyCount.fetch_add( dirty.taken, std::memory_order_relaxed );
		}
		if ( binMoved )
		{
			s->binsChanged.store( 1, std::memory_order_relaxed );
		}
		// the unused tail of the last chunk: records the apply pass skips
		while ( dirty.next < dirty.end )
		{
			b2gStream4( s->hDirty.ptr + (size_t)dirty.next * b2g::kDirtyBodyQuads, b2gIntBits( -1 ), 0.0f, 0.0f, 0.0f );
			dirty.next += 1;
		}
	}

	// ---- contacts: the 112 of b2ContactSim's 200 bytes that prepare reads (src/contact_solver.c:1629-1785)
	{
		float4* wire = base + s->inWire;
		float4* wireMass = base + s->inMass;
		const bool resident = s->resident;
		const bool usable = s->cacheUsable;
		const bool pending = s->defer && s->deferPending;
		int materialized = 0;
		b2gStreamChunk full = { &s->fullCursor, s->fullCapacity, 0, 0, &s->streamOverflow };
		bool massDiffers = false;
		int vouchedTotal = 0;
		int flat = ( begin > bodyCount ? begin : bodyCount ) - bodyCount;
		int flatEnd = ( end - bodyCount < s->contactTotal ? end - bodyCount : s->contactTotal );
		int k = flat < flatEnd ? b2gFindSegment( s->contactStart, flat ) : 0;
		while ( flat < flatEnd )
		{
			const b2gContactSeg& seg = s->contactSegs[k];
			int segFlat = s->contactStart[k];
			int local = flat - segFlat;
			int localEnd = ( flatEnd < s->contactStart[k + 1] ? flatEnd : s->contactStart[k + 1] ) - segFlat;
			int bodyBase = s->bodySegs[seg.world].base;
			const uint8_t* worldSims = s->bodySegs[seg.world].sims;
			// resident mode: the segment's homes, how many of them were occupied in the previous step and where they were then
			const int homeKey = resident ? s->segHome[k] : 0;
			const int homeBase = resident ? s->homeBase[homeKey] : 0;
			const int homeCount = resident && usable ? s->homeCount[homeKey] : 0;
			const int homeSlot = resident ? s->homeSlot[homeKey] : 0;
			const int hintCount = resident && usable && seg.hints != nullptr ? ( seg.hintCount < homeCount ? seg.hintCount : homeCount ) : 0;
			// The contacts are walked in the reference's SIMD groups: aligned groups of 4 contacts of the colour's array.  The
			// reference skips rolling resistance / restitution for a whole register when all its lanes have none
			// (src/contact_solver.c:2021, :2131; 4 lanes in the default build), so every contact needs to know whether ANY
			// member of its group has some (x == 0 is false for NaN, like _mm_cmpeq_ps).  Members outside [local, localEnd)
			// belong to another pack block: they are only looked at.
			for ( int group = local & ~3; group < localEnd; group += 4 )
			{
				const int groupEnd = group + 4 < seg.count ? group + 4 : seg.count;
				bool vouched[4] = { false, false, false, false };
				int groupBits = 0;
				if ( kPackPrefetch > 0 && !seg.hintsInPlace && group + kPackPrefetch < seg.count )
				{
					// the first cache line of the contacts two groups ahead (a 200-byte stride of which one line is read: the
					// hardware prefetcher does not follow it), their hints and shadow heads
					for ( int j = 0; j < 4; ++j )
					{
						_mm_prefetch( reinterpret_cast<const char*>( seg.sims + (size_t)( group + kPackPrefetch + j ) * B2L_CONTACT_SIZE ), _MM_HINT_T0 );
					}
				}
				for ( int j = group; j < groupEnd; ++j )
				{
					const uint8_t* sim = seg.sims + (size_t)j * B2L_CONTACT_SIZE;
					int ownBits = -1;
					if ( j < hintCount )
					{
						// A contact the narrow phase recycled (b2GpuStepDesc::recycled): if it still is where it was in the
						// previous step, with the same bodies, the device has everything but the separations -- nothing of
						// its record is read beyond the first cache line.
						const b2GpuRecycledContact& hint = seg.hints[j];
						const b2gShadowHead& head = s->shadowHeads[(size_t)( homeBase + j )];
						// (entries written in place: the caller says entry j is contact j -- the contact itself is not touched)
						const int id = seg.hintsInPlace ? hint.contactId : b2gRdI( sim, B2L_CONTACT_ID );
						const int simIndexA = seg.hintsInPlace ? hint.indexA : b2gRdI( sim, B2L_CONTACT_INDEX_A );
						const int simIndexB = seg.hintsInPlace ? hint.indexB : b2gRdI( sim, B2L_CONTACT_INDEX_B );
						if ( hint.stamp == seg.hintStamp && hint.contactId == id && head.contactId == id && head.indexA == simIndexA &&
							 head.indexB == simIndexB )
						{
							vouched[j - group] = true;
							ownBits = head.ownBits;
						}
					}
					if ( ownBits < 0 )
					{
						ownBits = ( !( b2gRdF( sim, B2L_CONTACT_ROLLING_RESISTANCE ) == 0.0f ) ? b2g::kMetaGroupRolling : 0 ) |
								  ( !( b2gRdF( sim, B2L_CONTACT_RESTITUTION ) == 0.0f ) ? b2g::kMetaGroupRestitution : 0 );
					}
					groupBits |= seg.wide ? ownBits : 0;
				}
				for ( int i = group > local ? group : local; i < groupEnd && i < localEnd; ++i )
				{
					const int slot = seg.slotStart + i;
					if ( vouched[i - group] )
					{
						const b2GpuRecycledContact& hint = seg.hints[i];
						b2gStream4( wire + slot, b2gIntBits( ( homeBase + i ) | ( ( groupBits >> 3 ) << b2g::kLightGroupShift ) | b2g::kLightBodyMass ),
									hint.separation[0], hint.separation[1], b2gIntBits( homeSlot + i ) );
						vouchedTotal += 1;
						continue;
					}
					const uint8_t* sim = seg.sims + (size_t)i * B2L_CONTACT_SIZE;
					if ( pending )
					{
						// a contact nobody vouches for is read in full below: if the previous step's impulses never reached
						// its manifold (nothing has read it since), they do now
						bool hit = false;
						materialized += b2gMaterializeAt( s, homeKey, i, seg.sims + (size_t)i * B2L_CONTACT_SIZE, &hit );
					}
					const uint8_t* m = sim + B2L_CONTACT_MANIFOLD;
					const uint8_t* p0 = m + B2L_MANIFOLD_POINTS;
					const uint8_t* p1 = p0 + B2L_MP_SIZE;
					int pointCount = b2gRdI( m, B2L_MANIFOLD_POINT_COUNT );
					int hitEnable = ( (uint32_t)b2gRdI( sim, B2L_CONTACT_SIM_FLAGS ) & B2L_SIM_ENABLE_HIT_EVENT ) != 0 ? b2g::kMetaHitEnable : 0;
					int meta = ( seg.colorIndex << b2g::kMetaColorShift ) | hitEnable | ( pointCount & b2g::kMetaPointMask );
					int indexA = b2gRdI( sim, B2L_CONTACT_INDEX_A ), indexB = b2gRdI( sim, B2L_CONTACT_INDEX_B );
					if ( s->checkMasses )
					{
						// invMass, invInertia of the two bodies as the contact remembers them vs. as the bodies have them now
						// (adjacent floats in both structures; bitwise, so that "equal" means the device may use either)
						static const uint8_t zero[8] = { 0 };
						const uint8_t* bodyA = indexA >= 0 ? worldSims + (size_t)indexA * B2L_SIM_SIZE + B2L_SIM_INV_MASS : zero;
						const uint8_t* bodyB = indexB >= 0 ? worldSims + (size_t)indexB * B2L_SIM_SIZE + B2L_SIM_INV_MASS : zero;
						massDiffers = massDiffers || memcmp( sim + B2L_CONTACT_INV_MASS_A, bodyA, 8 ) != 0 ||
									  memcmp( sim + B2L_CONTACT_INV_MASS_B, bodyB, 8 ) != 0;
					}
					const int rawIndexA = indexA, rawIndexB = indexB;
					indexA = indexA >= 0 ? indexA + bodyBase : indexA;
					indexB = indexB >= 0 ? indexB + bodyBase : indexB;
					b2gStream4( wireMass + slot, b2gRdF( sim, B2L_CONTACT_INV_MASS_A ), b2gRdF( sim, B2L_CONTACT_INV_I_A ),
								b2gRdF( sim, B2L_CONTACT_INV_MASS_B ), b2gRdF( sim, B2L_CONTACT_INV_I_B ) );
					const float separation0 = b2gRdF( p0, B2L_MP_SEPARATION ), separation1 = b2gRdF( p1, B2L_MP_SEPARATION );
					const float rollingImpulse = b2gRdF( m, B2L_MANIFOLD_ROLLING_IMPULSE );
					if ( !resident )
					{
						float4* w = wire + (size_t)slot * b2g::WR_COUNT;
						b2gStream4( w + b2g::WR_HEAD, b2gIntBits( indexA ), b2gIntBits( indexB ), b2gIntBits( meta | groupBits ), rollingImpulse );
						b2gStream4( w + b2g::WR_NORMAL, b2gRdF( m, B2L_MANIFOLD_NORMAL ), b2gRdF( m, B2L_MANIFOLD_NORMAL + 4 ),
									b2gRdF( sim, B2L_CONTACT_FRICTION ), b2gRdF( sim, B2L_CONTACT_TANGENT_SPEED ) );
						b2gStream4( w + b2g::WR_MATERIAL, b2gRdF( sim, B2L_CONTACT_ROLLING_RESISTANCE ), b2gRdF( sim, B2L_CONTACT_RESTITUTION ),
									separation0, separation1 );
						b2gStream4( w + b2g::WR_ANCHOR1, b2gRdF( p0, B2L_MP_ANCHOR_A ), b2gRdF( p0, B2L_MP_ANCHOR_A + 4 ),
									b2gRdF( p0, B2L_MP_ANCHOR_B ), b2gRdF( p0, B2L_MP_ANCHOR_B + 4 ) );
						b2gStream4( w + b2g::WR_ANCHOR2, b2gRdF( p1, B2L_MP_ANCHOR_A ), b2gRdF( p1, B2L_MP_ANCHOR_A + 4 ),
									b2gRdF( p1, B2L_MP_ANCHOR_B ), b2gRdF( p1, B2L_MP_ANCHOR_B + 4 ) );
						b2gStream4( w + b2g::WR_IMPULSE, b2gRdF( p0, B2L_MP_NORMAL_IMPULSE ), b2gRdF( p0, B2L_MP_TANGENT_IMPULSE ),
									b2gRdF( p1, B2L_MP_NORMAL_IMPULSE ), b2gRdF( p1, B2L_MP_TANGENT_IMPULSE ) );
						continue;
					}
					// resident mode: the record the device would need, against the record it has
					alignas( 16 ) float rows[17] = {
						b2gIntBits( indexA ), b2gIntBits( indexB ), b2gIntBits( meta ),
						b2gRdF( m, B2L_MANIFOLD_NORMAL ), b2gRdF( m, B2L_MANIFOLD_NORMAL + 4 ), b2gRdF( sim, B2L_CONTACT_FRICTION ), b2gRdF( sim, B2L_CONTACT_TANGENT_SPEED ),
						b2gRdF( sim, B2L_CONTACT_ROLLING_RESISTANCE ), b2gRdF( sim, B2L_CONTACT_RESTITUTION ),
						b2gRdF( p0, B2L_MP_ANCHOR_A ), b2gRdF( p0, B2L_MP_ANCHOR_A + 4 ), b2gRdF( p0, B2L_MP_ANCHOR_B ), b2gRdF( p0, B2L_MP_ANCHOR_B + 4 ),
						b2gRdF( p1, B2L_MP_ANCHOR_A ), b2gRdF( p1, B2L_MP_ANCHOR_A + 4 ), b2gRdF( p1, B2L_MP_ANCHOR_B ), b2gRdF( p1, B2L_MP_ANCHOR_B + 4 ) };
					float impulses[5] = { b2gRdF( p0, B2L_MP_NORMAL_IMPULSE ), b2gRdF( p0, B2L_MP_TANGENT_IMPULSE ), b2gRdF( p1, B2L_MP_NORMAL_IMPULSE ),
										  b2gRdF( p1, B2L_MP_TANGENT_IMPULSE ), rollingImpulse };
					const int id = b2gRdI( sim, B2L_CONTACT_ID );
					const int home = homeBase + i;
					b2gShadowContact& shadow = s->shadowContacts[(size_t)home];
					b2gShadowImpulses& shadowImpulses = s->shadowImpulses[(size_t)home];
					b2gShadowHead& head = s->shadowHeads[(size_t)home];
					int ref = homeSlot + i; // its record among the previous step's outputs
					bool clean = i < homeCount && head.contactId == id && memcmp( shadow.rows, rows, sizeof( rows ) ) == 0;
					if ( clean && s->defer )
					{
						// the impulses in the manifold against what the device computed last (rollingImpulse, normalImpulse1,
						// tangentImpulse1, ..., normalImpulse2, tangentImpulse2: b2g::ImpulseRecord) -- the unpack pass keeps no shadow
						const float* rec = s->prevRecords + (size_t)ref * b2g::kImpulseFloats;
						clean = s->prevRecords != nullptr && memcmp( rec + 1, impulses + 0, 8 ) == 0 && memcmp( rec + 5, impulses + 2, 8 ) == 0 &&
								memcmp( rec + 0, impulses + 4, 4 ) == 0;
					}
					else if ( clean )
					{
						clean = memcmp( shadowImpulses.values, impulses, sizeof( impulses ) ) == 0;
					}
					if ( !clean )
					{
						int entry = b2gStreamTake( full );
						ref = ~entry;
						float4* w = s->hFull.ptr + (size_t)entry * b2g::WR_COUNT;
						b2gStream4( w + b2g::WR_HEAD, rows[0], rows[1], rows[2], rollingImpulse );
						b2gStream4( w + b2g::WR_NORMAL, rows[3], rows[4], rows[5], rows[6] );
						b2gStream4( w + b2g::WR_MATERIAL, rows[7], rows[8], separation0, separation1 );
						b2gStream4( w + b2g::WR_ANCHOR1, rows[9], rows[10], rows[11], rows[12] );
						b2gStream4( w + b2g::WR_ANCHOR2, rows[13], rows[14], rows[15], rows[16] );
						b2gStream4( w + b2g::WR_IMPULSE, impulses[0], impulses[1], impulses[2], impulses[3] );
						memcpy( shadow.rows, rows, sizeof( rows ) );
						memcpy( shadowImpulses.values, impulses, sizeof( impulses ) );
						head.contactId = id;
						head.indexA = rawIndexA;
						head.indexB = rawIndexB;
						head.ownBits = ( !( rows[7] == 0.0f ) ? b2g::kMetaGroupRolling : 0 ) | ( !( rows[8] == 0.0f ) ? b2g::kMetaGroupRestitution : 0 );
					}
					b2gStream4( wire + slot, b2gIntBits( home | ( ( groupBits >> 3 ) << b2g::kLightGroupShift ) ), separation0, separation1, b2gIntBits( ref ) );
				}
			}
			if ( localEnd == seg.count )
			{
				// dead slots between this segment and the next (segments start on multiples of 4 slots): a zero head
				// (pointCount 0) -- resident mode: a negative key -- is all the kernels look at
				int segEnd = seg.slotStart + seg.count;
				int next = (size_t)k + 1 < s->contactSegs.size() ? s->contactSegs[(size_t)k + 1].slotStart : segEnd;
				int limit = ( segEnd + 3 ) & ~3;
				next = next < limit ? next : limit;
				for ( int dead = segEnd; dead < next; ++dead )
				{
					if ( resident )
					{
						b2gStream4( wire + dead, b2gIntBits( -1 ), 0.0f, 0.0f, 0.0f );
					}
					else
					{
						_mm_stream_ps( reinterpret_cast<float*>( wire + (size_t)dead * b2g::WR_COUNT + b2g::WR_HEAD ), _mm_setzero_ps() );
					}
				}
			}
			flat = s->contactStart[k + 1];
			k += 1;
		}
		if ( massDiffers )
		{
			s->massMismatch.store( 1, std::memory_order_release );
		}
		if ( full.taken > 0 )
		{
			s->fullCount.fetch_add( full.taken, std::memory_order_relaxed );
		}
		if ( vouchedTotal > 0 )
		{
			s->vouchedCount.fetch_add( vouchedTotal, std::memory_order_relaxed );
		}
		if ( materialized > 0 )
		{
			s->materialized.fetch_add( materialized, std::memory_order_relaxed );
		}
	}

	// ---- joints: the prepared b2JointSim padded to 256 bytes; bodies renumbered to the batch, and the world's base in
	// the joint-event bit set stored in the padding (read by jointEventTest).  Resident mode: a joint whose record equals
	// what the device holds at its home -- outside the run of fields b2PrepareJoint rewrites every step -- travels as its
	// home, the place of its previous output record and that run (b2gAssembleJointsKernel, b2g_resident.cuh).
	{
		float4* wireJoints = base + s->inJoints;
		const bool resident = s->resident;
		b2gStreamChunk full = { &s->fullJointCursor, s->fullJointCapacity, 0, 0, &s->streamOverflow };
		bool heavy = false;
		bool jointMoved = false;
		int first = bodyCount + s->contactTotal;
		int flat = ( begin > first ? begin : first ) - first;
		int flatEnd = end - first;
		int k = flat < flatEnd ? b2gFindSegment( s->jointStart, flat ) : 0;
		while ( flat < flatEnd )
		{
			const b2gJointSeg& seg = s->jointSegs[k];
			int local = flat - s->jointStart[k];
			int localEnd = ( flatEnd < s->jointStart[k + 1] ? flatEnd : s->jointStart[k + 1] ) - s->jointStart[k];
			const b2gBodySeg& world = s->bodySegs[seg.world];
			const int homeKey = resident ? s->jointSegHome[k] : 0;
			const int homeBase = resident ? s->jointHomeBase[homeKey] : 0;
			const int homeCount = resident && s->cacheUsable ? s->jointHomeCount[homeKey] : 0;
			const int homeSlot = resident ? s->jointHomeSlot[homeKey] : 0;
			for ( int i = local; i < localEnd; ++i )
			{
				alignas( 16 ) uint8_t padded[b2g::kJointStride] = { 0 };
				memcpy( padded, seg.sims + (size_t)i * B2L_JOINT_SIZE, B2L_JOINT_SIZE );
				if ( world.base != 0 )
				{
					int* pair = b2gJointIndexPair( reinterpret_cast<b2lJointSim*>( padded ) );
					if ( pair != nullptr )
					{
						pair[0] = pair[0] >= 0 ? pair[0] + world.base : pair[0];
						pair[1] = pair[1] >= 0 ? pair[1] + world.base : pair[1];
					}
				}
				memcpy( padded + B2L_JOINT_SIZE, &world.jointBitBase, 4 );
				{
					// is every joint of the step a plain revolute joint?  (the next step's plan goes by it, b2gPlanIslands)
					const b2lJointSim* joint = reinterpret_cast<const b2lJointSim*>( padded );
					heavy = heavy || !( joint->type == b2l_revoluteJoint && joint->u.revolute.enableSpring == 0 &&
										joint->u.revolute.enableMotor == 0 && joint->u.revolute.enableLimit == 0 );
				}
				if ( !resident )
				{
					b2gStreamCopy( wireJoints + (size_t)( seg.jointStart + i ) * ( b2g::kJointStride / 16 ), padded, b2g::kJointStride / 16 );
					continue;
				}
				const int home = homeBase + i;
				uint8_t* shadow = s->shadowJoints.data() + (size_t)home * b2g::kJointStride;
				int runOffset = 0;
				const int runBytes = b2lJointPreparedRun( reinterpret_cast<const b2lJointSim*>( padded )->type, &runOffset );
				auto sameOutsideTheRun = [&]() {
					return i < homeCount && memcmp( shadow, padded, (size_t)runOffset ) == 0 &&
						   memcmp( shadow + runOffset + runBytes, padded + runOffset + runBytes, (size_t)( b2g::kJointStride - runOffset - runBytes ) ) == 0;
				};
				bool clean = sameOutsideTheRun();
				if ( !clean && s->defer && s->deferJointsPending &&
					 b2gMaterializeJointAt( s, homeKey, i, seg.sims + (size_t)i * B2L_JOINT_SIZE ) != 0 )
				{
					// the record travels in full: the previous step's outputs, which never reached this b2JointSim (nobody has read
					// it since), do now -- and the record is built again from what it holds then
					memcpy( padded, seg.sims + (size_t)i * B2L_JOINT_SIZE, B2L_JOINT_SIZE );
					if ( world.base != 0 )
					{
						int* pair = b2gJointIndexPair( reinterpret_cast<b2lJointSim*>( padded ) );
						if ( pair != nullptr )
						{
							pair[0] = pair[0] >= 0 ? pair[0] + world.base : pair[0];
							pair[1] = pair[1] >= 0 ? pair[1] + world.base : pair[1];
						}
					}
					clean = sameOutsideTheRun();
				}
				float4* light = wireJoints + (size_t)( seg.jointStart + i ) * b2g::kLightJointQuads;
				int ref = homeSlot + i; // its record among the previous step's outputs
				if ( clean )
				{
					// its bodies are part of the run that travels every step: the bins' lists of the previous step only serve
					// this one (b2gEnqueueRun) if they are the bodies the joint had then
					int* pair = b2gJointIndexPair( reinterpret_cast<b2lJointSim*>( padded ) );
					if ( pair != nullptr )
					{
						int* known = reinterpret_cast<int*>( shadow + ( reinterpret_cast<uint8_t*>( pair ) - padded ) );
						if ( known[0] != pair[0] || known[1] != pair[1] )
						{
							known[0] = pair[0];
							known[1] = pair[1];
							jointMoved = true;
						}
					}
					alignas( 16 ) uint8_t run[B2L_JOINT_RUN_MAX] = { 0 };
					memcpy( run, padded + runOffset, (size_t)runBytes );
					b2gStreamCopy( light + 1, run, B2L_JOINT_RUN_MAX / 16 );
				}
				else
				{
					int entry = b2gStreamTake( full );
					ref = ~entry;
					b2gStreamCopy( s->hFullJoints.ptr + (size_t)entry * ( b2g::kJointStride / 16 ), padded, b2g::kJointStride / 16 );
					memcpy( shadow, padded, b2g::kJointStride );
				}
				b2gStream4( light, b2gIntBits( home ), b2gIntBits( ref ), 0.0f, 0.0f );
			}
			flat = s->jointStart[k + 1];
			k += 1;
		}
		if ( full.taken > 0 )
		{
			s->fullJointCount.fetch_add( full.taken, std::memory_order_relaxed );
		}
		if ( heavy )
		{
			s->heavyJoint.store( 1, std::memory_order_relaxed );
		}
		if ( jointMoved )
		{
			s->binsChanged.store( 1, std::memory_order_relaxed );
		}
	}
	_mm_sfence();
}

// ---- phase 3: H2D + kernels + D2H, all asynchronous on the solver's stream -------------------------------------------
// quads of the input arena that hold items [0, itemEnd) (bodies, then contacts in slot order, then joints)
// The pack pass claims blocks in ARENA order: first the blocks of the constraints (items [bodyCount, itemCount)), then
// the blocks of the bodies (items [0, bodyCount)).
static void b2gPackBlockRange( const b2GpuSolver* s, int block, int* begin, int* end )
{
	int bodyCount = s->params.bodyCount;
	int restItems = s->workItems - bodyCount;
	int restBlocks = ( restItems + s->blockItems - 1 ) / s->blockItems;
	if ( block < restBlocks )
	{
		*begin = bodyCount + block * s->blockItems;
		*end = *begin + s->blockItems < s->workItems ? *begin + s->blockItems : s->workItems;
	}
	else
	{
		*begin = ( block - restBlocks ) * s->blockItems;
		*end = *begin + s->blockItems < bodyCount ? *begin + s->blockItems : bodyCount;
	}
}

// quads of the input arena that are complete once the first `blocksDone` blocks (in claim order) are packed
static size_t b2gPackedPrefix( const b2GpuSolver* s, int blocksDone )
{
	int restItems = s->contactTotal + s->jointTotal;
	int restBlocks = ( restItems + s->blockItems - 1 ) / s->blockItems;
	if ( blocksDone >= s->workBlocks )
	{
		return s->inTotal;
	}
	if ( blocksDone >= restBlocks )
	{
		return s->inStates; // the three body regions are interleaved by region, not by body: wait for all bodies
	}
	int flat = blocksDone * s->blockItems; // constraints [0, flat) are packed, flat < restItems
	if ( flat < s->contactTotal )
	{
		int k = b2gFindSegment( s->contactStart, flat );
		int slot = s->contactSegs[k].slotStart + ( flat - s->contactStart[k] );
		return s->inWire + (size_t)slot * s->wireQuads;
	}
	return s->inJoints + (size_t)( flat - s->contactTotal ) * s->jointWireQuads;
}

// one host -> device copy on the solver's stream, or -- while the step's last copies are being collected -- an entry of the batch
static int b2gCopyUp( b2GpuSolver* s, void* dst, const void* src, size_t bytes )
{
	if ( s->batching )
	{
		s->batchDst.push_back( dst );
		s->batchSrc.push_back( const_cast<void*>( src ) );
		s->batchBytes.push_back( bytes );
		return 0;
	}
	B2G_CUDA( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyHostToDevice, s->stream ) );
	return 0;
}

static int b2gFlushBatch( b2GpuSolver* s )
{
	s->batching = false;
	const size_t count = s->batchDst.size();
	int rc = 0;
	if ( count > 1 )
	{
		cudaMemcpyAttributes attributes = {};
		attributes.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
		size_t attributeIndex = 0, failed = 0;
		cudaError_t err = cudaMemcpyBatchAsync( s->batchDst.data(), s->batchSrc.data(), s->batchBytes.data(), count, &attributes, &attributeIndex, 1,
												&failed, s->stream );
		if ( err != cudaSuccess )
		{
			// (a driver without the batch call: one copy each, from now on)
			cudaGetLastError();
			s->batchEnabled = false;
			for ( size_t i = 0; i < count && rc == 0; ++i )
			{
				rc = b2gCopyUp( s, s->batchDst[i], s->batchSrc[i], s->batchBytes[i] );
			}
		}
	}
	else if ( count == 1 )
	{
		rc = b2gCopyUp( s, s->batchDst[0], s->batchSrc[0], s->batchBytes[0] );
	}
	s->batchDst.clear();
	s->batchSrc.clear();
	s->batchBytes.clear();
	return rc;
}

// enqueue the upload of quads [fromQuads, uptoQuads) of the input arena
static int b2gSendRange( b2GpuSolver* s, size_t fromQuads, size_t uptoQuads )
{
	if ( !s->uploadStarted )
	{
		B2G_CUDA( cudaEventRecord( s->evUpload, s->stream ) );
		s->uploadStarted = true;
	}
	if ( uptoQuads > fromQuads )
	{
		if ( s->trace )
		{
			s->traceSends.emplace_back( std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count(), uptoQuads );
		}
		return b2gCopyUp( s, s->wireAll.ptr + fromQuads, s->hWire.ptr + fromQuads, ( uptoQuads - fromQuads ) * sizeof( float4 ) );
	}
	return 0;
}

// resident mode: the two variable-length streams, once the packing is done (dense prefixes of their staging buffers)
static int b2gSendStreams( b2GpuSolver* s )
{
	if ( !s->resident )
	{
		return 0;
	}
	if ( s->streamOverflow.load( std::memory_order_relaxed ) != 0 )
	{
		return b2gFailMsg( "b2GpuSolverPackRange: packed in ranges too small for the planned streams" );
	}
	int full = s->fullCursor.load( std::memory_order_acquire ), dirty = s->dirtyCursor.load( std::memory_order_acquire );
	full = full < s->fullCapacity ? full : s->fullCapacity;
	dirty = dirty < s->dirtyCapacity ? dirty : s->dirtyCapacity;
	if ( full > 0 && b2gCopyUp( s, s->fullStream.ptr, s->hFull.ptr, (size_t)full * b2g::WR_COUNT * sizeof( float4 ) ) != 0 )
	{
		return 1;
	}
	if ( dirty > 0 && b2gCopyUp( s, s->dirtyStream.ptr, s->hDirty.ptr, (size_t)dirty * b2g::kDirtyBodyQuads * sizeof( float4 ) ) != 0 )
	{
		return 1;
	}
	int fullJoints = s->fullJointCursor.load( std::memory_order_acquire );
	fullJoints = fullJoints < s->fullJointCapacity ? fullJoints : s->fullJointCapacity;
	if ( fullJoints > 0 && b2gCopyUp( s, s->fullJointStream.ptr, s->hFullJoints.ptr, (size_t)fullJoints * b2g::kJointStride ) != 0 )
	{
		return 1;
	}
	s->fullSent = full;
	s->dirtySent = dirty;
	s->fullJointSent = fullJoints;
	return 0;
}

// the whole input arena in one piece (callers that packed it with b2GpuSolverPackRange)
int b2gSendArena( b2GpuSolver* s, size_t uptoQuads )
{
	if ( s->arenaSent )
	{
		return 0; // the pipelined pack pass has sent it piece by piece
	}
	s->arenaSent = true;
	(void)uptoQuads;
	if ( b2gSendRange( s, 0, s->inMass ) != 0 || b2gSendStreams( s ) != 0 )
	{
		return 1;
	}
	return s->massMismatch.load( std::memory_order_acquire ) != 0 ? b2gSendRange( s, s->inMass, s->inTotal ) : 0;
}

// ---- phase 4: unpack (callable concurrently on disjoint ranges) ---------------------------------------------------------
// Evict a consumed part of the D2H staging buffer from the CPU caches.  On the target hosts a DMA write into lines
// that are still cached by several cores runs at ~7 GB/s instead of ~54 GB/s (tools/microbench/d2h_bench.cu); flushing
// right after the unpack pass keeps the next step's download at full speed for ~0.03 ms of host work.
#if defined( __x86_64__ )
static bool b2gHasClflushopt()
{
	static int cached = -1;
	if ( cached < 0 )
	{
		unsigned a = 0, b = 0, c = 0, d = 0;
		cached = ( __get_cpuid_count( 7, 0, &a, &b, &c, &d ) != 0 && ( b & ( 1u << 23 ) ) != 0 ) ? 1 : 0;
	}
	return cached == 1;
}

__attribute__( ( target( "clflushopt" ) ) ) static void b2gFlushOpt( const char* p, const char* end )
{
	for ( ; p < end; p += 64 )
	{
		_mm_clflushopt( const_cast<char*>( p ) );
	}
}

void b2gFlushLines( const void* ptr, size_t bytes )
{
	if ( bytes == 0 )
	{
		return;
	}
	const char* p = reinterpret_cast<const char*>( reinterpret_cast<uintptr_t>( ptr ) & ~uintptr_t( 63 ) );
	const char* end = static_cast<const char*>( ptr ) + bytes;
	if ( b2gHasClflushopt() )
	{
		b2gFlushOpt( p, end );
	}
	else
	{
		for ( ; p < end; p += 64 )
		{
			_mm_clflush( p );
		}
	}
}
#else
void b2gFlushLines( const void*, size_t )
{
}
#endif

// Scatter the packed impulse records into the reference's manifolds: what b2StoreImpulsesTask
// (src/contact_solver.c:2293-2320) and b2StoreImpulses_Overflow (:526-542) write.
extern "C" void b2GpuSolverUnpackRange( b2GpuSolver* s, int begin, int end )
{
	if ( begin >= end )
	{
		return;
	}
	int bodyCount = s->params.bodyCount;
	const float4* base = s->hOut.ptr;

	// ---- body states
	{
		const float4* outStates = base + s->outStates;
		int i = begin;
		int bodyEnd = end < bodyCount ? end : bodyCount;
		int w = i < bodyEnd ? b2gFindSegment( s->bodyStart, i ) : 0;
		while ( i < bodyEnd )
		{
			const b2gBodySeg& seg = s->bodySegs[w];
			int segEnd = seg.base + seg.count < bodyEnd ? seg.base + seg.count : bodyEnd;
			if ( i < segEnd )
			{
				memcpy( seg.states + (size_t)( i - seg.base ) * B2L_STATE_SIZE, outStates + 2 * (size_t)i, (size_t)( segEnd - i ) * B2L_STATE_SIZE );
				if ( s->resident )
				{
					// what the device starts the next step from (storeBody, b2g_stages.cuh): this step's velocity and flags
					// with the transient flags cleared and the deltas reset, like the host's state after b2FinalizeBodiesTask
					for ( int j = i; j < segEnd; ++j )
					{
						float4 velocity = outStates[2 * (size_t)j];
						uint32_t flags;
						memcpy( &flags, &velocity.w, 4 );
						flags &= ~B2L_FLAG_TRANSIENT;
						memcpy( &velocity.w, &flags, 4 );
						s->shadowStates[2 * (size_t)j] = velocity;
						s->shadowStates[2 * (size_t)j + 1] = make_float4( 0.0f, 0.0f, 1.0f, 0.0f );
					}
				}
				b2gFlushLines( outStates + 2 * (size_t)i, (size_t)( segEnd - i ) * B2L_STATE_SIZE );
				i = segEnd;
			}
			w += 1;
		}
	}

	// ---- contact impulses (deferred impulses: not here -- b2GpuSolverMaterializeContacts, when somebody needs them)
	if ( !s->defer )
	{
		const float* allRecords = reinterpret_cast<const float*>( base + s->outImpulses );
		int flat = ( begin > bodyCount ? begin : bodyCount ) - bodyCount;
		int flatEnd = ( end - bodyCount < s->contactTotal ? end - bodyCount : s->contactTotal );
		int k = flat < flatEnd ? b2gFindSegment( s->contactStart, flat ) : 0;
		while ( flat < flatEnd )
		{
			const b2gContactSeg& seg = s->contactSegs[k];
			b2GpuStepResult* result = s->results != nullptr ? s->results + seg.world : nullptr;
			uint64_t* hitBits = result != nullptr ? result->hitEventBits : nullptr;
			int local = flat - s->contactStart[k];
			int localEnd = ( flatEnd < s->contactStart[k + 1] ? flatEnd : s->contactStart[k + 1] ) - s->contactStart[k];
			const float* records = allRecords + (size_t)seg.slotStart * b2g::kImpulseFloats;
			for ( int i = local; i < localEnd; ++i )
			{
				uint8_t* sim = seg.sims + (size_t)i * B2L_CONTACT_SIZE;
				uint8_t* manifold = sim + B2L_CONTACT_MANIFOLD;
				const float* rec = records + (size_t)i * b2g::kImpulseFloats;
				int pointCount = seg.wide ? 2 : b2gRdI( manifold, B2L_MANIFOLD_POINT_COUNT );
				memcpy( manifold + B2L_MANIFOLD_ROLLING_IMPULSE, rec + 0, 4 );
				for ( int j = 0; j < pointCount; ++j )
				{
					uint8_t* mp = manifold + B2L_MANIFOLD_POINTS + j * B2L_MP_SIZE;
					// normalImpulse, tangentImpulse, totalNormalImpulse, normalVelocity are contiguous (collision.h:549-561)
					memcpy( mp + B2L_MP_NORMAL_IMPULSE, rec + 1 + 4 * j, 16 );
				}
				if ( s->resident )
				{
					// the manifold now holds what the device holds: remember it for the next pack pass
					b2gShadowImpulses& shadow = s->shadowImpulses[(size_t)( s->homeBase[s->segHome[k]] + i )];
					memcpy( shadow.values + 4, rec + 0, 4 );
					for ( int j = 0; j < pointCount && j < 2; ++j )
					{
						memcpy( shadow.values + 2 * j, rec + 1 + 4 * j, 8 );
					}
				}
				if ( rec[9] != 0.0f && result != nullptr )
				{
					if ( hitBits != nullptr )
					{
						uint32_t id = (uint32_t)b2gRdI( sim, B2L_CONTACT_ID );
						__atomic_fetch_or( hitBits + ( id >> 6 ), (uint64_t)1 << ( id & 63u ), __ATOMIC_RELAXED );
					}
					__atomic_store_n( &result->hasHitEvents, 1, __ATOMIC_RELAXED );
				}
			}
			if ( local < localEnd )
			{
				b2gFlushLines( records + (size_t)local * b2g::kImpulseFloats, (size_t)( localEnd - local ) * b2g::kImpulseFloats * sizeof( float ) );
			}
			flat = s->contactStart[k + 1];
			k += 1;
		}
	}

	// ---- joints: the fields the stages wrote (b2lJointMutableRuns) go back into the reference's b2JointSim in place
	// (deferred: not here -- b2GpuSolverMaterializeJoints, when somebody needs them)
	if ( !s->defer )
	{
		const float* outJoints = reinterpret_cast<const float*>( base + s->outJoints );
		int first = bodyCount + s->contactTotal;
		int flat = ( begin > first ? begin : first ) - first;
		int flatEnd = end - first;
		int k = flat < flatEnd ? b2gFindSegment( s->jointStart, flat ) : 0;
		while ( flat < flatEnd )
		{
			const b2gJointSeg& seg = s->jointSegs[k];
			int local = flat - s->jointStart[k];
			int localEnd = ( flatEnd < s->jointStart[k + 1] ? flatEnd : s->jointStart[k + 1] ) - s->jointStart[k];
			for ( int i = local; i < localEnd; ++i )
			{
				uint8_t* sim = seg.sims + (size_t)i * B2L_JOINT_SIZE;
				const float* record = outJoints + (size_t)( seg.jointStart + i ) * B2L_JOINT_OUT_FLOATS;
				int offsets[2], floats[2];
				int runs = b2lJointMutableRuns( b2gRdI( sim, offsetof( b2lJointSim, type ) ), offsets, floats );
				uint8_t* shadow = s->resident ? s->shadowJoints.data() + (size_t)( s->jointHomeBase[s->jointSegHome[k]] + i ) * b2g::kJointStride : nullptr;
				for ( int r = 0; r < runs; ++r )
				{
					memcpy( sim + offsets[r], record, (size_t)floats[r] * sizeof( float ) );
					if ( shadow != nullptr )
					{
						// the device's table holds the same (b2gAssembleJointsKernel takes them from the output records)
						memcpy( shadow + offsets[r], record, (size_t)floats[r] * sizeof( float ) );
					}
					record += floats[r];
				}
			}
			if ( local < localEnd )
			{
				b2gFlushLines( outJoints + (size_t)( seg.jointStart + local ) * B2L_JOINT_OUT_FLOATS,
							   (size_t)( localEnd - local ) * B2L_JOINT_OUT_FLOATS * sizeof( float ) );
			}
			flat = s->jointStart[k + 1];
			k += 1;
		}
	}
}

// ---- pipelined host passes ------------------------------------------------------------------------------------------------
// b2GpuSolverPackWork / b2GpuSolverUnpackWork are called by ANY number of host threads at the same time (the world's
// workers); each call claims blocks of items until none are left.  Exactly one caller passes pump = 1: besides packing
// it starts the upload of every finished prefix of the arena (PCIe runs behind the packing instead of after it), and
// besides unpacking it watches the download events and tells the others how much of the output has arrived (unpacking
// runs behind the download).
// Uploads what has been packed.  The constraint blocks come first in the arena and block b fills the quads
// [b2gPackedPrefix( b ), b2gPackedPrefix( b + 1 )): every long enough run of finished, unsent blocks goes out as one
// copy -- runs, not just the finished PREFIX, so that one slow block (a worker that was descheduled in the middle of it)
// does not hold back everything behind it.  The three body regions are interleaved by region, not by body: they follow
// in one piece when all blocks are done (`everything`: also the runs that stayed short).
static int b2gPumpUploads( b2GpuSolver* s, bool everything )
{
	const int restItems = s->contactTotal + s->jointTotal;
	const int restBlocks = ( restItems + s->blockItems - 1 ) / s->blockItems;
	while ( s->pumpPrefix < s->workBlocks && s->workDone[s->pumpPrefix].load( std::memory_order_acquire ) != 0 )
	{
		s->pumpPrefix += 1;
	}
	const bool complete = s->pumpPrefix == s->workBlocks;
	if ( s->trace && s->tracePump.size() < 48 && ( s->tracePump.empty() || s->tracePump.back().second != (size_t)s->pumpPrefix ) )
	{
		s->tracePump.emplace_back( std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count(), (size_t)s->pumpPrefix );
	}
	if ( complete && !everything )
	{
		// everything is packed: what is left goes out in ONE copy, from the caller that ends the pass (a copy costs 5 - 10 us
		// of driver time on the calling thread, and the last ones are on the step's critical path)
		return 0;
	}
	while ( s->sendScan < restBlocks && s->blockSent[(size_t)s->sendScan] != 0 )
	{
		s->sendScan += 1;
	}
	// the end of what follows the constraints in the arena (the body regions; the masses' region only when some contact's
	// differ from its bodies', b2g::WireRow)
	const size_t arenaEnd = complete && everything ? ( s->massMismatch.load( std::memory_order_acquire ) != 0 ? s->inTotal : s->inMass ) : 0;
	bool tailSent = false;
	s->batching = complete && everything && s->batchEnabled; // the step's last copies: one driver call (b2gFlushBatch)
	for ( int i = s->sendScan; i < restBlocks; )
	{
		if ( s->blockSent[(size_t)i] != 0 || s->workDone[i].load( std::memory_order_acquire ) == 0 )
		{
			i += 1;
			continue;
		}
		int j = i + 1;
		while ( j < restBlocks && s->blockSent[(size_t)j] == 0 && s->workDone[j].load( std::memory_order_acquire ) != 0 )
		{
			j += 1;
		}
		size_t from = b2gPackedPrefix( s, i ), upto = b2gPackedPrefix( s, j );
		if ( complete && everything && j == restBlocks )
		{
			upto = arenaEnd; // the last run of constraints and the body regions behind it: one piece
			tailSent = true;
		}
		if ( upto - from >= s->sendThreshold || ( complete && everything ) )
		{
			if ( b2gSendRange( s, from, upto ) != 0 )
			{
				return 1;
			}
			// small pieces first (the link starts early), then larger ones (a copy costs a few microseconds of set-up; the
			// packing runs ahead of the link, so there is always more to send)
			s->sendThreshold = s->sendThreshold * 2 < kTransferQuadsMax ? s->sendThreshold * 2 : kTransferQuadsMax;
			for ( int k = i; k < j; ++k )
			{
				s->blockSent[(size_t)k] = 1;
			}
		}
		i = j;
	}
	if ( complete && everything )
	{
		s->arenaSent = true;
		if ( ( !tailSent && b2gSendRange( s, s->inStates, arenaEnd ) != 0 ) || b2gSendStreams( s ) != 0 )
		{
			s->batching = false;
			return 1;
		}
		return b2gFlushBatch( s );
	}
	return 0;
}

// Whoever has just finished a block looks after the uploads, unless somebody else is already doing that: the transfers do
// not wait for one particular thread (which may be in the middle of a block, or descheduled).
static int b2gTryPumpUploads( b2GpuSolver* s )
{
	if ( s->pumpBusy.exchange( 1, std::memory_order_acquire ) != 0 )
	{
		return 0;
	}
	int rc = b2gPumpUploads( s, false );
	s->pumpBusy.store( 0, std::memory_order_release );
	return rc;
}

extern "C" int b2GpuSolverPackWork( b2GpuSolver* s, int pump )
{
	if ( s == nullptr || !s->begun )
	{
		return b2gFailMsg( "b2GpuSolverPackWork: no step begun" );
	}
	cudaSetDevice( s->device );
	for ( ;; )
	{
		int block = s->workNext.fetch_add( 1, std::memory_order_acq_rel );
		if ( block >= s->workBlocks )
		{
			break;
		}
		int begin, end;
		b2gPackBlockRange( s, block, &begin, &end );
		b2GpuSolverPackRange( s, begin, end ); // ends with an sfence: the streaming stores are visible to the DMA engine
		s->workDone[block].store( 1, std::memory_order_release );
		if ( b2gTryPumpUploads( s ) != 0 )
		{
			s->workFailed.store( 1 );
			return 1;
		}
	}
	if ( pump != 0 )
	{
		// the others may still be packing the blocks they claimed; the last bytes go out from here
		for ( ;; )
		{
			if ( s->pumpBusy.exchange( 1, std::memory_order_acquire ) == 0 )
			{
				int rc = b2gPumpUploads( s, false );
				bool complete = s->pumpPrefix == s->workBlocks;
				if ( rc == 0 && complete )
				{
					rc = b2gPumpUploads( s, true );
				}
				s->pumpBusy.store( 0, std::memory_order_release );
				if ( rc != 0 )
				{
					s->workFailed.store( 1 );
					return 1;
				}
				if ( complete )
				{
					break;
				}
			}
			if ( s->workFailed.load( std::memory_order_relaxed ) != 0 )
			{
				return 1;
			}
			_mm_pause();
		}
	}
	return 0;
}

// quads of the output arena that must have arrived before items [0, itemEnd) can be unpacked
static size_t b2gOutPrefix( const b2GpuSolver* s, int itemEnd )
{
	int bodyCount = s->params.bodyCount;
	if ( itemEnd <= bodyCount )
	{
		return s->outStates + 2 * (size_t)itemEnd;
	}
	int flat = itemEnd - bodyCount;
	if ( flat <= s->contactTotal )
	{
		// the record of the last contact of the range
		int k = b2gFindSegment( s->contactStart, flat - 1 );
		int slot = s->contactSegs[k].slotStart + ( flat - 1 - s->contactStart[k] );
		return s->outImpulses + ( (size_t)( slot + 1 ) * b2g::kImpulseFloats + 3 ) / 4;
	}
	int joints = flat - s->contactTotal;
	joints = joints < s->jointTotal ? joints : s->jointTotal;
	return s->outJoints + (size_t)joints * ( B2L_JOINT_OUT_FLOATS / 4 );
}

static int b2gPumpDownloads( b2GpuSolver* s )
{
	if ( !s->controlSeen )
	{
		// kernels done?  (the control block is the first thing that comes back)
		bool seen = false;
		if ( b2gPollControl( s, &seen ) != 0 )
		{
			return 1;
		}
		if ( !seen )
		{
			return 0;
		}
		if ( s->ran && s->islandMode && s->hControl->islandFailed != 0 )
		{
			// rare: rerun on the grid-barrier kernel; that re-enqueues the downloads and waits for them
			B2G_CUDA( cudaStreamSynchronize( s->stream ) );
			if ( b2gRerunIfIslandsFailed( s, true ) != 0 )
			{
				return 1;
			}
			s->chunkNext = s->chunkCount;
			s->arrivedQuads.store( s->outTotal, std::memory_order_release );
		}
		s->controlSeen = true;
		s->kernelsSeen.store( 1, std::memory_order_release );
		if ( s->direct && s->arrivedQuads.load( std::memory_order_relaxed ) < s->directEnd )
		{
			s->arrivedQuads.store( s->directEnd, std::memory_order_release ); // the kernels stored the body states themselves
		}
		s->tWaited = std::chrono::steady_clock::now();
		s->traceControl = std::chrono::duration<float, std::micro>( s->tWaited - s->tBegin ).count();
	}
	while ( s->chunkNext < s->chunkCount )
	{
		cudaError_t err = cudaEventQuery( s->chunkEvents[(size_t)s->chunkNext] );
		if ( err == cudaErrorNotReady )
		{
			break;
		}
		if ( err != cudaSuccess )
		{
			return b2gFail( "download", err );
		}
		s->arrivedQuads.store( s->chunkEnd[(size_t)s->chunkNext], std::memory_order_release );
		if ( s->trace )
		{
			s->traceArrivals.emplace_back( std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count(),
										   s->chunkEnd[(size_t)s->chunkNext] );
		}
		s->chunkNext += 1;
	}
	return 0;
}

// (Tried: each thread unpacks the blocks it packed, so that the b2ContactSim lines it writes are still in its own cache.
// No measurable gain on the 16-core hosts, and the search for one's own blocks does not scale to the thousands of blocks
// of a batch; blocks are claimed in order.)
// Whoever waits for output looks after the download events, unless somebody else is already doing that (the pump = 1
// caller may be in the middle of a block: with a few large blocks per thread the others would wait for it to look up).
static int b2gTryPumpDownloads( b2GpuSolver* s )
{
	if ( s->pumpBusy.exchange( 1, std::memory_order_acquire ) != 0 )
	{
		return 0;
	}
	int rc = b2gPumpDownloads( s );
	s->pumpBusy.store( 0, std::memory_order_release );
	return rc;
}

extern "C" int b2GpuSolverUnpackWork( b2GpuSolver* s, int pump )
{
	if ( s == nullptr || !s->begun )
	{
		return b2gFailMsg( "b2GpuSolverUnpackWork: no step begun" );
	}
	cudaSetDevice( s->device );
	for ( ;; )
	{
		int block = s->workNext.fetch_add( 1, std::memory_order_acq_rel );
		if ( block >= s->workBlocks )
		{
			break;
		}
		int begin = block * s->blockItems;
		int end = begin + s->blockItems < s->workItems ? begin + s->blockItems : s->workItems;
		int begin2 = 0, end2 = 0; // deferred impulses: the items are the bodies and the joints, a block may hold some of both
		if ( s->defer )
		{
			// (the items are the bodies: the contacts' and the joints' records are deferred)
			const int bodyCount = s->params.bodyCount;
			begin = begin < bodyCount ? begin : bodyCount;
			end = end < bodyCount ? end : bodyCount;
		}
		size_t need = b2gOutPrefix( s, end2 > end ? end2 : end );
		unsigned spins = 0;
		while ( s->arrivedQuads.load( std::memory_order_acquire ) < need )
		{
			// Whoever waits looks after the downloads -- the pump = 1 caller all the time, the others now and then: a dozen
			// threads exchanging on the try-lock's cache line (and reading the host-mapped flag the device is about to
			// write) for the whole length of the kernels slows down the one poll that matters, and everything else on a
			// host that several ranks share.
			if ( ( pump != 0 || kPollAll || ( ++spins & 63u ) == 0u ) && b2gTryPumpDownloads( s ) != 0 )
			{
				s->workFailed.store( 1 );
				return 1;
			}
			if ( s->workFailed.load( std::memory_order_relaxed ) != 0 )
			{
				return 1;
			}
			_mm_pause();
		}
		b2GpuSolverUnpackRange( s, begin, end );
		b2GpuSolverUnpackRange( s, begin2, end2 );
	}
	// The CUDA-event time of the kernels (the stage split of b2Profile is scaled to it) is a driver call of 4 - 5 us: the
	// first caller that runs out of blocks after the kernels were seen to be done reads it while the others still unpack;
	// failing that, the pump = 1 caller reads it on its way out, as before.
	auto readKernelTimer = [&]() -> int {
		cudaError_t err = cudaEventElapsedTime( &s->lastKernelMs, s->evStart, s->evStop );
		s->timerDone.store( 1, std::memory_order_release );
		return err == cudaSuccess ? 0 : b2gFail( "kernel timer", err );
	};
	static const bool eagerTimers = getenv( "B2GPU_EAGER_TIMERS" ) != nullptr && atoi( getenv( "B2GPU_EAGER_TIMERS" ) ) != 0; // (A/B)
	if ( pump == 0 && !eagerTimers && s->ran && s->kernelsSeen.load( std::memory_order_acquire ) != 0 && s->timerClaim.exchange( 1, std::memory_order_acq_rel ) == 0 )
	{
		if ( readKernelTimer() != 0 )
		{
			s->workFailed.store( 1 );
			return 1;
		}
	}
	if ( pump != 0 )
	{
		s->traceMarks[5] = std::chrono::duration<float, std::micro>( std::chrono::steady_clock::now() - s->tBegin ).count();
		// the tail of the arena (joint event bits) is consumed by EndStep
		for ( ;; )
		{
			if ( s->pumpBusy.exchange( 1, std::memory_order_acquire ) == 0 )
			{
				int rc = b2gPumpDownloads( s );
				bool done = s->controlSeen && s->chunkNext >= s->deferWaitChunks; // (all chunks, unless the impulse records are deferred)
				s->pumpBusy.store( 0, std::memory_order_release );
				if ( rc != 0 )
				{
					s->workFailed.store( 1 );
					return 1;
				}
				if ( done )
				{
					break;
				}
			}
			if ( s->workFailed.load( std::memory_order_relaxed ) != 0 )
			{
				return 1;
			}
			_mm_pause();
		}
		if ( s->ran )
		{
			if ( s->timerClaim.exchange( 1, std::memory_order_acq_rel ) == 0 )
			{
				if ( readKernelTimer() != 0 )
				{
					return 1;
				}
			}
			else
			{
				while ( s->timerDone.load( std::memory_order_acquire ) == 0 && s->workFailed.load( std::memory_order_relaxed ) == 0 )
				{
					_mm_pause();
				}
			}
		}
	}
	return 0;
}


// ---- host utility: island sizes from labels ---------------------------------------------------------------------------------
extern "C" int b2GpuCountIslandSizes( const b2GpuStepDesc* d, b2GpuIslandSize* sizes )
{
	if ( d == nullptr || sizes == nullptr || d->bodyIsland == nullptr || d->islandCount <= 0 )
	{
		return b2gFailMsg( "b2GpuCountIslandSizes: no island labels" );
	}
	memset( sizes, 0, (size_t)d->islandCount * sizeof( b2GpuIslandSize ) );
	auto islandOf = [&]( int body ) -> int {
		int label = body >= 0 && body < d->awakeBodyCount ? d->bodyIsland[body] : -1;
		return label >= 0 && label < d->islandCount ? label : -1;
	};
	for ( int i = 0; i < d->awakeBodyCount; ++i )
	{
		int island = islandOf( i );
		if ( island >= 0 )
		{
			sizes[island].bodyCount += 1;
		}
	}
	for ( int c = 0; c <= d->activeColorCount; ++c )
	{
		const b2GpuColorDesc& color = c < d->activeColorCount ? d->colors[c] : d->overflow;
		const uint8_t* contacts = static_cast<const uint8_t*>( color.contactSims );
		for ( int i = 0; i < color.contactCount; ++i )
		{
			const uint8_t* sim = contacts + (size_t)i * B2L_CONTACT_SIZE;
			int a = b2gRdI( sim, B2L_CONTACT_INDEX_A ), b = b2gRdI( sim, B2L_CONTACT_INDEX_B );
			int island = islandOf( a >= 0 ? a : b );
			if ( island >= 0 )
			{
				sizes[island].contactCount += 1;
			}
		}
		uint8_t* joints = static_cast<uint8_t*>( color.jointSims );
		for ( int i = 0; i < color.jointCount; ++i )
		{
			const int* pair = b2gJointIndexPair( reinterpret_cast<b2lJointSim*>( joints + (size_t)i * B2L_JOINT_SIZE ) );
			int island = pair != nullptr ? islandOf( pair[0] >= 0 ? pair[0] : pair[1] ) : -1;
			if ( island >= 0 )
			{
				sizes[island].jointCount += 1;
			}
		}
	}
	return 0;
}
