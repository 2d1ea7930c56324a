// b2g_contact.cuh -- contact constraint stages, one thread per constraint.
//
// Two families, exactly as in the reference (they use DIFFERENT operation orders):
//   *  coloured contacts follow the wide path   src/contact_solver.c:1573-2331 (b2ContactConstraintWide)
//   *  overflow contacts follow the scalar path src/contact_solver.c:24-545    (b2ContactConstraint)
// Expression association is kept operator by operator (see b2g_math.cuh for the compile flags).
#pragma once

#include "b2g_types.cuh"

#include "b2gpu_layout.h"

#include <cooperative_groups.h>

namespace b2g
{

// ---- raw AoS readers --------------------------------------------------------------------------------------
B2G_DEV float rawF( const uint8_t* base, int offset )
{
	return *reinterpret_cast<const float*>( base + offset );
}

B2G_DEV int rawI( const uint8_t* base, int offset )
{
	return *reinterpret_cast<const int*>( base + offset );
}

// ---- body gather / scatter (replaces b2GatherBodies / b2ScatterBodies, src/contact_solver.c:1121-1562) -----
// Index 0 is the static dummy body: vel = 0, flags = 0, pos = identity; it is never written because the
// scatter is guarded by b2_dynamicFlag exactly like the reference (contact_solver.c:1531).
// Plain generic loads/stores: the view may live in shared memory (island-local kernel) or in global memory
// (grid-barrier kernel, where the barrier's release/acquire pair orders them across blocks).
// block of the cluster that owns body `linear` (0-based index inside the bin): linear / clusterRun
B2G_DEV unsigned clusterOwner( const SolveView& V, unsigned linear )
{
	return __umulhi( linear, V.clusterMagic );
}

// distributed shared memory address of slot `index` of a per-block body array (cluster mode)
template <typename T> B2G_DEV T* clusterSlot( const SolveView& V, T* localArray, int index )
{
	unsigned linear = (unsigned)index - 1u;
	unsigned owner = clusterOwner( V, linear );
	unsigned slot = linear - owner * (unsigned)V.clusterRun + 1u;
	T* remote = static_cast<T*>( __cluster_map_shared_rank( static_cast<void*>( localArray ), owner ) );
	return remote + slot;
}

B2G_DEV float4 gatherVel( const SolveView& V, int index )
{
	if ( V.clusterRun == 0 || index == 0 )
	{
		return V.vel[index];
	}
	return *clusterSlot( V, V.vel, index );
}

B2G_DEV float4 gatherPos( const SolveView& V, int index )
{
	if ( V.clusterRun == 0 || index == 0 )
	{
		return V.pos[index];
	}
	return *clusterSlot( V, V.pos, index );
}

B2G_DEV void scatterVel( const SolveView& V, int index, float4 v )
{
	if ( ( __float_as_uint( v.w ) & B2L_FLAG_DYNAMIC ) != 0 )
	{
		if ( V.clusterRun == 0 )
		{
			V.vel[index] = v;
		}
		else if ( V.asyncBar == 0 )
		{
			*clusterSlot( V, V.vel, index ) = v; // a dynamic body is never the static dummy
		}
		else
		{
			// the owner block counts the bytes that arrive for its bodies (mbarrier transaction count) instead of every
			// writer fencing its stores at GPU scope
			unsigned linear = (unsigned)index - 1u;
			unsigned owner = clusterOwner( V, linear );
			unsigned slot = linear - owner * (unsigned)V.clusterRun + 1u;
			unsigned local = (unsigned)__cvta_generic_to_shared( V.vel ) + slot * (unsigned)sizeof( float4 );
			unsigned remote, remoteBar;
			asm volatile( "mapa.shared::cluster.u32 %0, %1, %2;" : "=r"( remote ) : "r"( local ), "r"( owner ) );
			asm volatile( "mapa.shared::cluster.u32 %0, %1, %2;" : "=r"( remoteBar ) : "r"( V.asyncBar ), "r"( owner ) );
			asm volatile( "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
						  :
						  : "r"( remote ), "r"( __float_as_uint( v.x ) ), "r"( __float_as_uint( v.y ) ), "r"( __float_as_uint( v.z ) ),
							"r"( __float_as_uint( v.w ) ), "r"( remoteBar )
						  : "memory" );
		}
	}
}

B2G_DEV float4 loadField( const SolveView& V, int field, int slot )
{
	return V.cf[(size_t)field * V.cfStride + slot];
}

B2G_DEV void storeField( const SolveView& V, int field, int slot, float4 value )
{
	V.cf[(size_t)field * V.cfStride + slot] = value;
}

// ---- the wire record of a slot ------------------------------------------------------------------------------
// Plain mode: WR_COUNT rows per slot in P.wire.  Resident mode (P.light != nullptr, b2g_types.cuh): the slot's light
// record names the contact id and where the rest lives -- the resident table + the previous step's output record, or
// a full record of this step.

// WR_HEAD of a slot with the SIMD-group bits in its meta word; pointCount 0 = dead slot
B2G_DEV float4 wireHead( const StepParams& P, int slot )
{
	if ( P.light == nullptr )
	{
		return P.wire[(size_t)slot * WR_COUNT + WR_HEAD];
	}
	float4 L = P.light[slot];
	int key = __float_as_int( L.x );
	if ( key < 0 )
	{
		return make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
	}
	int ref = __float_as_int( L.w );
	float4 head = ref < 0 ? P.full[(size_t)( ~ref ) * WR_COUNT + WR_HEAD] : P.table[(size_t)( key & kLightIdMask ) * kTableRows + WR_HEAD];
	head.z = __int_as_float( __float_as_int( head.z ) | ( ( ( key >> kLightGroupShift ) & 3 ) << 3 ) );
	return head;
}

// starts the slot's record on its way into L2 (the prepare pass reads it a few barriers later)
B2G_DEV void prefetchWire( const StepParams& P, int slot )
{
	const uint8_t* record;
	int bytes = WR_COUNT * 16;
	if ( P.light == nullptr )
	{
		record = reinterpret_cast<const uint8_t*>( P.wire + (size_t)slot * WR_COUNT );
	}
	else
	{
		float4 L = P.light[slot];
		int key = __float_as_int( L.x ), ref = __float_as_int( L.w );
		if ( key < 0 )
		{
			return;
		}
		if ( ref < 0 )
		{
			record = reinterpret_cast<const uint8_t*>( P.full + (size_t)( ~ref ) * WR_COUNT );
		}
		else
		{
			record = reinterpret_cast<const uint8_t*>( P.table + (size_t)( key & kLightIdMask ) * kTableRows );
			bytes = kTableRows * 16;
			asm volatile( "prefetch.global.L2 [%0];" ::"l"( P.prevImpulses + (size_t)ref * kImpulseFloats ) );
		}
	}
	for ( int sector = 0; sector < bytes; sector += 32 )
	{
		asm volatile( "prefetch.global.L2 [%0];" ::"l"( record + sector ) );
	}
}

// ---- prepare ---------------------------------------------------------------------------------------------
// b2PrepareContactsTask (src/contact_solver.c:1573-1809) per lane; with wide == false it is
// b2PrepareContacts_Overflow (src/contact_solver.c:24-160): no contact-softening branch.
// Reads the wire record `wireSlot`, takes the bodies' current {v, w} (sA, sB; zero for a static body, like the
// reference's vA = 0 for B2_NULL_INDEX) and writes constraint slot `slot` of the view; localA/localB are the body
// indices in the view's numbering (0 = static).
B2G_DEV void prepareContact( const StepParams& P, const SolveView& V, int wireSlot, int slot, int localA, int localB, float4 sA,
							 float4 sB, bool wide, int groupBits )
{
	float4 head, nrm, mat, imp, anchor1, anchor2;
	bool massFromBodies = P.massFromBodies != 0;
	if ( P.light == nullptr )
	{
		const float4* w = P.wire + (size_t)wireSlot * WR_COUNT;
		head = w[WR_HEAD];
		nrm = w[WR_NORMAL];
		mat = w[WR_MATERIAL];
		anchor1 = w[WR_ANCHOR1];
		anchor2 = w[WR_ANCHOR2];
		imp = w[WR_IMPULSE];
	}
	else
	{
		float4 L = P.light[wireSlot];
		int key = __float_as_int( L.x ), ref = __float_as_int( L.w );
		const float4* w = ref < 0 ? P.full + (size_t)( ~ref ) * WR_COUNT : P.table + (size_t)( key & kLightIdMask ) * kTableRows;
		massFromBodies = massFromBodies || ( key & kLightBodyMass ) != 0;
		head = w[WR_HEAD];
		nrm = w[WR_NORMAL];
		mat = w[WR_MATERIAL];
		anchor1 = w[WR_ANCHOR1];
		anchor2 = w[WR_ANCHOR2];
		mat.z = L.y; // this step's separations
		mat.w = L.z;
		if ( ref < 0 )
		{
			imp = w[WR_IMPULSE];
		}
		else
		{
			// the device's own output of the previous step (storeContact): what the host would send back
			const float2* r = reinterpret_cast<const float2*>( P.prevImpulses + (size_t)ref * kImpulseFloats );
			float2 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3];
			head.w = r0.x;
			imp = make_float4( r0.y, r1.x, r2.y, r3.x );
		}
	}
	float4 mass;
	if ( massFromBodies )
	{
		// the contact's masses are its bodies' (the pack pass compared them bit for bit); a static body has none
		int indexA = __float_as_int( head.x ), indexB = __float_as_int( head.y );
		float4 a = indexA >= 0 ? P.wireBody[2 * (size_t)indexA] : make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
		float4 b = indexB >= 0 ? P.wireBody[2 * (size_t)indexB] : make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
		mass = make_float4( a.x, a.y, b.x, b.y );
	}
	else
	{
		mass = P.wireMass[wireSlot];
	}
	int meta = __float_as_int( head.z );
	int pointCount = meta & kMetaPointMask;

	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;

	V2 vA = v2( sA.x, sA.y );
	float wA = sA.z;
	V2 vB = v2( sB.x, sB.y );
	float wB = sB.z;

	float rollingMass;
	{
		float k = iA + iB;
		rollingMass = k > 0.0f ? 1.0f / k : 0.0f;
	}

	Soft soft = P.contactSoft;
	if ( localA == 0 || localB == 0 )
	{
		soft = P.staticSoft;
	}
	else if ( wide && P.enableSoftening != 0 )
	{
		// contact_solver.c:1685-1699
		float contactHertz = minf_( P.contactHertz, 0.125f * P.inv_h );
		float ratio = 1.0f;
		if ( mA < mB )
		{
			ratio = maxf_( 0.5f, mA / mB );
		}
		else if ( mB < mA )
		{
			ratio = maxf_( 0.5f, mB / mA );
		}
		soft = makeSoft( ratio * contactHertz, ratio * P.contactDampingRatio, P.h );
	}

	float warmStartScale = P.enableWarmStarting != 0 ? 1.0f : 0.0f;

	V2 normal = v2( nrm.x, nrm.y );
	V2 tangent = rightPerp( normal );
	float friction = nrm.z;
	float tangentSpeed = nrm.w;
	float rollingResistance = mat.x;
	float restitution = mat.y;
	float rollingImpulse = warmStartScale * head.w;

	float4 anchors[2] = { anchor1, anchor2 };
	float separation[2] = { mat.z, mat.w };
	float wireNormalImpulse[2] = { imp.x, imp.z };
	float wireTangentImpulse[2] = { imp.y, imp.w };
	float4 impulses[2];
	float normalMass[2], tangentMass[2], baseSeparation[2], relativeVelocity[2];

#pragma unroll
	for ( int j = 0; j < 2; ++j )
	{
		if ( j < pointCount )
		{
			V2 rA = v2( anchors[j].x, anchors[j].y );
			V2 rB = v2( anchors[j].z, anchors[j].w );

			baseSeparation[j] = separation[j] - dot( sub( rB, rA ), normal );

			impulses[j].x = warmStartScale * wireNormalImpulse[j];
			impulses[j].y = warmStartScale * wireTangentImpulse[j];
			impulses[j].z = 0.0f;

			float rnA = cross( rA, normal );
			float rnB = cross( rB, normal );
			float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
			normalMass[j] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;

			float rtA = cross( rA, tangent );
			float rtB = cross( rB, tangent );
			float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
			tangentMass[j] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;

			// relative velocity for restitution
			V2 vrA = add( vA, crossSV( wA, rA ) );
			V2 vrB = add( vB, crossSV( wB, rB ) );
			relativeVelocity[j] = dot( normal, sub( vrB, vrA ) );
		}
		else
		{
			// dummy data that has no effect (contact_solver.c:1786-1800)
			anchors[j] = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
			impulses[j] = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
			baseSeparation[j] = 0.0f;
			normalMass[j] = 0.0f;
			tangentMass[j] = 0.0f;
			relativeVelocity[j] = 0.0f;
		}
	}
	impulses[0].w = rollingImpulse;
	impulses[1].w = 0.0f;

	storeField( V, CF_MASS, slot, make_float4( mA, iA, mB, iB ) );
	storeField( V, CF_NORMAL, slot, make_float4( normal.x, normal.y, friction, tangentSpeed ) );
	storeField( V, CF_ROLL, slot, make_float4( rollingResistance, restitution, rollingMass, 0.0f ) );
	storeField( V, CF_SOFT, slot, make_float4( soft.biasRate, soft.massScale, soft.impulseScale, 0.0f ) );
	storeField( V, CF_ANCHOR1, slot, anchors[0] );
	storeField( V, CF_ANCHOR2, slot, anchors[1] );
	storeField( V, CF_PMASS, slot, make_float4( normalMass[0], tangentMass[0], normalMass[1], tangentMass[1] ) );
	storeField( V, CF_BASE, slot,
				make_float4( baseSeparation[0], baseSeparation[1], relativeVelocity[0], relativeVelocity[1] ) );
	storeField( V, CF_IMP1, slot, impulses[0] );
	storeField( V, CF_IMP2, slot, impulses[1] );

	if ( !( restitution == 0.0f ) )
	{
		*V.anyRestitution = 1;
	}

	// 0 for null (contact_solver.c:1644-1646)
	V.cidx[slot] = make_int2( localA, localB );
	V.cmeta[slot] = ( meta & ( kMetaPointMask | kMetaHitEnable ) ) | groupBits;
}

// The reference skips rolling resistance / restitution for a whole SIMD register when all its lanes have none
// (src/contact_solver.c:2021, :2131).  The default build is SSE2 with 4 lanes, so the test is over aligned groups of
// 4 consecutive constraints of the colour's array.  The pack pass evaluates it in array order (b2g_wire.cu); the result
// arrives in the meta word (kMetaGroup*, see wireHead) and travels with the constraint in cmeta.
constexpr int kMetaGroupMask = kMetaGroupRolling | kMetaGroupRestitution;

// ===========================================================================================================
// Wide path (coloured contacts)
// ===========================================================================================================

// b2WarmStartContactsTask, src/contact_solver.c:1811-1871
B2G_DEV void warmStartContact( const SolveView& V, int slot )
{
	int2 idx = V.cidx[slot];
	float4 bA = gatherVel( V, idx.x );
	float4 bB = gatherVel( V, idx.y );

	float4 mass = loadField( V, CF_MASS, slot );
	float invMassA = mass.x, invIA = mass.y, invMassB = mass.z, invIB = mass.w;
	float4 nrm = loadField( V, CF_NORMAL, slot );
	float nx = nrm.x, ny = nrm.y;
	float tangentX = ny;
	float tangentY = 0.0f - nx;

	float4 imp1 = loadField( V, CF_IMP1, slot );
	float4 imp2 = loadField( V, CF_IMP2, slot );

	{
		float4 a = loadField( V, CF_ANCHOR1, slot );
		float Px = imp1.x * nx + imp1.y * tangentX;
		float Py = imp1.x * ny + imp1.y * tangentY;
		bA.z = bA.z - invIA * ( a.x * Py - a.y * Px );
		bA.x = bA.x - invMassA * Px;
		bA.y = bA.y - invMassA * Py;
		bB.z = invIB * ( a.z * Py - a.w * Px ) + bB.z;
		bB.x = invMassB * Px + bB.x;
		bB.y = invMassB * Py + bB.y;
		imp1.z = imp1.z + imp1.x;
	}
	{
		float4 a = loadField( V, CF_ANCHOR2, slot );
		float Px = imp2.x * nx + imp2.y * tangentX;
		float Py = imp2.x * ny + imp2.y * tangentY;
		bA.z = bA.z - invIA * ( a.x * Py - a.y * Px );
		bA.x = bA.x - invMassA * Px;
		bA.y = bA.y - invMassA * Py;
		bB.z = invIB * ( a.z * Py - a.w * Px ) + bB.z;
		bB.x = invMassB * Px + bB.x;
		bB.y = invMassB * Py + bB.y;
		imp2.z = imp2.z + imp2.x;
	}

	bA.z = bA.z - invIA * imp1.w;
	bB.z = invIB * imp1.w + bB.z;

	storeField( V, CF_IMP1, slot, imp1 );
	storeField( V, CF_IMP2, slot, imp2 );
	scatterVel( V, idx.x, bA );
	scatterVel( V, idx.y, bB );
}

// One non-penetration row of b2SolveContactsTask (src/contact_solver.c:1909-1964 / 1966-2016)
B2G_DEV void solveNormalRow( float4& bA, float4& bB, float4 pA, float4 pB, float dpx, float dpy, float4 a, float nx, float ny,
							 float baseSeparation, float normalMass, float biasRate, float massScale, float impulseScale,
							 float inv_h, float negContactSpeed, float invMassA, float invIA, float invMassB, float invIB,
							 float& normalImpulse, float& totalImpulseOfPoint, float& totalNormalImpulse )
{
	// moving anchors for the current separation: rs = rotate(dq, r)
	float rsAx = pA.z * a.x - pA.w * a.y;
	float rsAy = pA.w * a.x + pA.z * a.y;
	float rsBx = pB.z * a.z - pB.w * a.w;
	float rsBy = pB.w * a.z + pB.z * a.w;

	float dsx = dpx + ( rsBx - rsAx );
	float dsy = dpy + ( rsBy - rsAy );
	float s = ( nx * dsx + ny * dsy ) + baseSeparation;

	bool mask = s > 0.0f;
	float specBias = s * inv_h;
	float softBias = maxf_( biasRate * s, negContactSpeed );
	float bias = mask ? specBias : softBias;
	float pointMassScale = mask ? 1.0f : massScale;
	float pointImpulseScale = mask ? 0.0f : impulseScale;

	// relative velocity at contact
	float dvx = ( bB.x - bB.z * a.w ) - ( bA.x - bA.z * a.y );
	float dvy = ( bB.y + bB.z * a.z ) - ( bA.y + bA.z * a.x );
	float vn = dvx * nx + dvy * ny;

	float negImpulse = normalMass * ( pointMassScale * vn + bias ) + pointImpulseScale * normalImpulse;

	float newImpulse = maxf_( normalImpulse - negImpulse, 0.0f );
	float impulse = newImpulse - normalImpulse;
	normalImpulse = newImpulse;
	totalImpulseOfPoint = totalImpulseOfPoint + impulse;
	totalNormalImpulse = totalNormalImpulse + newImpulse;

	float Px = impulse * nx;
	float Py = impulse * ny;

	bA.x = bA.x - invMassA * Px;
	bA.y = bA.y - invMassA * Py;
	bA.z = bA.z - invIA * ( a.x * Py - a.y * Px );

	bB.x = invMassB * Px + bB.x;
	bB.y = invMassB * Py + bB.y;
	bB.z = invIB * ( a.z * Py - a.w * Px ) + bB.z;
}

// One friction row (src/contact_solver.c:2036-2071 / 2073-2108)
B2G_DEV void solveFrictionRow( float4& bA, float4& bB, float4 a, float tangentX, float tangentY, float tangentSpeed,
							   float tangentMass, float friction, float normalImpulse, float invMassA, float invIA,
							   float invMassB, float invIB, float& tangentImpulse )
{
	float dvx = ( bB.x - bB.z * a.w ) - ( bA.x - bA.z * a.y );
	float dvy = ( bB.y + bB.z * a.z ) - ( bA.y + bA.z * a.x );
	float vt = dvx * tangentX + dvy * tangentY;
	vt = vt - tangentSpeed;

	float negImpulse = tangentMass * vt;

	float maxFriction = friction * normalImpulse;
	float newImpulse = tangentImpulse - negImpulse;
	newImpulse = maxf_( 0.0f - maxFriction, minf_( newImpulse, maxFriction ) );
	float impulse = newImpulse - tangentImpulse;
	tangentImpulse = newImpulse;

	float Px = impulse * tangentX;
	float Py = impulse * tangentY;

	bA.x = bA.x - invMassA * Px;
	bA.y = bA.y - invMassA * Py;
	bA.z = bA.z - invIA * ( a.x * Py - a.y * Px );

	bB.x = invMassB * Px + bB.x;
	bB.y = invMassB * Py + bB.y;
	bB.z = invIB * ( a.z * Py - a.w * Px ) + bB.z;
}

// b2SolveContactsTask, src/contact_solver.c:1873-2116.
B2G_DEV void solveContact( const StepParams& P, const SolveView& V, int slot, bool useBias )
{
	float4 roll = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
	bool groupHasRolling = false;
	if ( useBias == false )
	{
		groupHasRolling = ( V.cmeta[slot] & kMetaGroupRolling ) != 0;
		if ( groupHasRolling )
		{
			roll = loadField( V, CF_ROLL, slot );
		}
	}

	int2 idx = V.cidx[slot];
	float4 bA = gatherVel( V, idx.x );
	float4 bB = gatherVel( V, idx.y );
	float4 pA = gatherPos( V, idx.x );
	float4 pB = gatherPos( V, idx.y );

	float4 mass = loadField( V, CF_MASS, slot );
	float invMassA = mass.x, invIA = mass.y, invMassB = mass.z, invIB = mass.w;
	float4 nrm = loadField( V, CF_NORMAL, slot );
	float nx = nrm.x, ny = nrm.y;
	float4 soft = loadField( V, CF_SOFT, slot );
	float4 pmass = loadField( V, CF_PMASS, slot );
	float4 base = loadField( V, CF_BASE, slot );
	float4 a1 = loadField( V, CF_ANCHOR1, slot );
	float4 a2 = loadField( V, CF_ANCHOR2, slot );
	float4 imp1 = loadField( V, CF_IMP1, slot );
	float4 imp2 = loadField( V, CF_IMP2, slot );

	float biasRate, massScale, impulseScale;
	if ( useBias )
	{
		biasRate = soft.y * soft.x;
		massScale = soft.y;
		impulseScale = soft.z;
	}
	else
	{
		biasRate = 0.0f;
		massScale = 1.0f;
		impulseScale = 0.0f;
	}

	float totalNormalImpulse = 0.0f;
	float dpx = pB.x - pA.x;
	float dpy = pB.y - pA.y;
	float negContactSpeed = -P.contactSpeed;

	solveNormalRow( bA, bB, pA, pB, dpx, dpy, a1, nx, ny, base.x, pmass.x, biasRate, massScale, impulseScale, P.inv_h,
					negContactSpeed, invMassA, invIA, invMassB, invIB, imp1.x, imp1.z, totalNormalImpulse );
	solveNormalRow( bA, bB, pA, pB, dpx, dpy, a2, nx, ny, base.y, pmass.z, biasRate, massScale, impulseScale, P.inv_h,
					negContactSpeed, invMassA, invIA, invMassB, invIB, imp2.x, imp2.z, totalNormalImpulse );

	if ( useBias == false )
	{
		// rolling resistance
		if ( groupHasRolling )
		{
			float deltaLambda = roll.z * ( bA.z - bB.z );
			float lambda = imp1.w;
			float maxLambda = roll.x * totalNormalImpulse;
			// b2SymClampW of the default SSE2 build flips the sign bit (src/contact_solver.c:869-878), so the lower
			// bound is -maxLambda (-0 for +0), unlike the AVX2 wrapper's 0 - maxLambda (:641-645)
			float nb = -maxLambda;
			imp1.w = maxf_( nb, minf_( lambda + deltaLambda, maxLambda ) );
			deltaLambda = imp1.w - lambda;

			bA.z = bA.z - invIA * deltaLambda;
			bB.z = invIB * deltaLambda + bB.z;
		}

		float tangentX = ny;
		float tangentY = 0.0f - nx;
		solveFrictionRow( bA, bB, a1, tangentX, tangentY, nrm.w, pmass.y, nrm.z, imp1.x, invMassA, invIA, invMassB, invIB,
						  imp1.y );
		solveFrictionRow( bA, bB, a2, tangentX, tangentY, nrm.w, pmass.w, nrm.z, imp2.x, invMassA, invIA, invMassB, invIB,
						  imp2.y );
	}

	storeField( V, CF_IMP1, slot, imp1 );
	storeField( V, CF_IMP2, slot, imp2 );
	scatterVel( V, idx.x, bA );
	scatterVel( V, idx.y, bB );
}

// One row of b2ApplyRestitutionTask (src/contact_solver.c:2144-2181 / 2183-2221)
B2G_DEV void restitutionRow( float4& bA, float4& bB, float4 a, float nx, float ny, float restitution, bool restitutionIsZero,
							 float threshold, float relativeVelocity, float normalMass, float invMassA, float invIA,
							 float invMassB, float invIB, float& normalImpulse, float& totalImpulseOfPoint )
{
	// set effective mass to zero if restitution should not be applied
	bool mask1 = ( relativeVelocity + threshold ) > 0.0f;
	bool mask2 = totalImpulseOfPoint == 0.0f;
	float mass = ( mask1 || mask2 || restitutionIsZero ) ? 0.0f : normalMass;

	float dvx = ( bB.x - bB.z * a.w ) - ( bA.x - bA.z * a.y );
	float dvy = ( bB.y + bB.z * a.z ) - ( bA.y + bA.z * a.x );
	float vn = dvx * nx + dvy * ny;

	float negImpulse = mass * ( vn + restitution * relativeVelocity );

	float newImpulse = maxf_( normalImpulse - negImpulse, 0.0f );
	float deltaImpulse = newImpulse - normalImpulse;
	normalImpulse = newImpulse;
	totalImpulseOfPoint = totalImpulseOfPoint + deltaImpulse;

	float Px = deltaImpulse * nx;
	float Py = deltaImpulse * ny;

	bA.x = bA.x - invMassA * Px;
	bA.y = bA.y - invMassA * Py;
	bA.z = bA.z - invIA * ( a.x * Py - a.y * Px );

	bB.x = invMassB * Px + bB.x;
	bB.y = invMassB * Py + bB.y;
	bB.z = invIB * ( a.z * Py - a.w * Px ) + bB.z;
}

// b2ApplyRestitutionTask, src/contact_solver.c:2118-2228 (group-of-4 early out at :2131)
B2G_DEV void restitutionContact( const StepParams& P, const SolveView& V, int slot )
{
	if ( ( V.cmeta[slot] & kMetaGroupRestitution ) == 0 )
	{
		return;
	}
	float4 roll = loadField( V, CF_ROLL, slot );
	float restitution = roll.y;

	bool restitutionIsZero = restitution == 0.0f;

	int2 idx = V.cidx[slot];
	float4 bA = gatherVel( V, idx.x );
	float4 bB = gatherVel( V, idx.y );

	float4 mass = loadField( V, CF_MASS, slot );
	float4 nrm = loadField( V, CF_NORMAL, slot );
	float4 pmass = loadField( V, CF_PMASS, slot );
	float4 base = loadField( V, CF_BASE, slot );
	float4 a1 = loadField( V, CF_ANCHOR1, slot );
	float4 a2 = loadField( V, CF_ANCHOR2, slot );
	float4 imp1 = loadField( V, CF_IMP1, slot );
	float4 imp2 = loadField( V, CF_IMP2, slot );

	restitutionRow( bA, bB, a1, nrm.x, nrm.y, restitution, restitutionIsZero, P.restitutionThreshold, base.z, pmass.x, mass.x,
					mass.y, mass.z, mass.w, imp1.x, imp1.z );
	restitutionRow( bA, bB, a2, nrm.x, nrm.y, restitution, restitutionIsZero, P.restitutionThreshold, base.w, pmass.z, mass.x,
					mass.y, mass.z, mass.w, imp2.x, imp2.z );

	storeField( V, CF_IMP1, slot, imp1 );
	storeField( V, CF_IMP2, slot, imp2 );
	scatterVel( V, idx.x, bA );
	scatterVel( V, idx.y, bB );
}

// b2StoreImpulsesTask, src/contact_solver.c:2238-2331 (wide == true) and b2StoreImpulses_Overflow :516-545
// (wide == false: no hit-event test).  The 9 floats + hit flag go to the packed record of the constraint's wire slot;
// the host scatters them into b2Manifold and sets the hit bit by contactId (the overflow path only writes the first
// pointCount points, the host honours that).
B2G_DEV void storeContact( const StepParams& P, const SolveView& V, int slot, int wireSlot, bool wide )
{
	float4 imp1 = loadField( V, CF_IMP1, slot );
	float4 imp2 = loadField( V, CF_IMP2, slot );
	float4 base = loadField( V, CF_BASE, slot );

	float hitFlag = 0.0f;
	if ( wide )
	{
		int meta = V.cmeta[slot];
		if ( ( meta & kMetaHitEnable ) != 0 )
		{
			int pointCount = meta & kMetaPointMask;
			float negHitThreshold = -P.hitEventThreshold;
			bool hit = ( base.z < negHitThreshold && imp1.z > 0.0f );
			if ( pointCount > 1 )
			{
				hit = hit || ( base.w < negHitThreshold && imp2.z > 0.0f );
			}
			if ( hit )
			{
				hitFlag = 1.0f;
				*P.hasHitEvents = 1;
			}
		}
	}

	float2* out = reinterpret_cast<float2*>( P.outImpulses + (size_t)wireSlot * kImpulseFloats );
	out[0] = make_float2( imp1.w, imp1.x ); // rollingImpulse, normalImpulse1
	out[1] = make_float2( imp1.y, imp1.z ); // tangentImpulse1, totalNormalImpulse1
	out[2] = make_float2( base.z, imp2.x ); // normalVelocity1, normalImpulse2
	out[3] = make_float2( imp2.y, imp2.z ); // tangentImpulse2, totalNormalImpulse2
	out[4] = make_float2( base.w, hitFlag ); // normalVelocity2, hit event
}

// ===========================================================================================================
// Scalar path (overflow colour): strictly sequential in array order, one thread.
// ===========================================================================================================

// b2WarmStartContacts_Overflow, src/contact_solver.c:162-237
B2G_DEV void warmStartContactOverflow( const SolveView& V, int slot )
{
	int2 idx = V.cidx[slot];
	int pointCount = V.cmeta[slot] & kMetaPointMask;
	float4 sA = gatherVel( V, idx.x );
	float4 sB = gatherVel( V, idx.y );
	V2 vA = v2( sA.x, sA.y );
	float wA = sA.z;
	V2 vB = v2( sB.x, sB.y );
	float wB = sB.z;

	float4 mass = loadField( V, CF_MASS, slot );
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;
	float4 nrm = loadField( V, CF_NORMAL, slot );
	V2 normal = v2( nrm.x, nrm.y );
	V2 tangent = rightPerp( normal );

	float4 imp[2] = { loadField( V, CF_IMP1, slot ), loadField( V, CF_IMP2, slot ) };
	float4 anc[2] = { loadField( V, CF_ANCHOR1, slot ), loadField( V, CF_ANCHOR2, slot ) };

	for ( int j = 0; j < pointCount; ++j )
	{
		V2 rA = v2( anc[j].x, anc[j].y );
		V2 rB = v2( anc[j].z, anc[j].w );

		V2 Pv = add( mulSV( imp[j].x, normal ), mulSV( imp[j].y, tangent ) );
		imp[j].z += imp[j].x;

		wA -= iA * cross( rA, Pv );
		vA = mulAdd( vA, -mA, Pv );
		wB += iB * cross( rB, Pv );
		vB = mulAdd( vB, mB, Pv );
	}

	wA -= iA * imp[0].w;
	wB += iB * imp[0].w;

	storeField( V, CF_IMP1, slot, imp[0] );
	storeField( V, CF_IMP2, slot, imp[1] );
	scatterVel( V, idx.x, make_float4( vA.x, vA.y, wA, sA.w ) );
	scatterVel( V, idx.y, make_float4( vB.x, vB.y, wB, sB.w ) );
}

// b2SolveContacts_Overflow, src/contact_solver.c:239-408 (friction BEFORE rolling resistance)
B2G_DEV void solveContactOverflow( const StepParams& P, const SolveView& V, int slot, bool useBias )
{
	int2 idx = V.cidx[slot];
	int pointCount = V.cmeta[slot] & kMetaPointMask;

	float4 mass = loadField( V, CF_MASS, slot );
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;

	float4 sA = gatherVel( V, idx.x );
	float4 qA = gatherPos( V, idx.x );
	V2 vA = v2( sA.x, sA.y );
	float wA = sA.z;
	Rot dqA;
	dqA.c = qA.z;
	dqA.s = qA.w;

	float4 sB = gatherVel( V, idx.y );
	float4 qB = gatherPos( V, idx.y );
	V2 vB = v2( sB.x, sB.y );
	float wB = sB.z;
	Rot dqB;
	dqB.c = qB.z;
	dqB.s = qB.w;

	V2 dp = sub( v2( qB.x, qB.y ), v2( qA.x, qA.y ) );

	float4 nrm = loadField( V, CF_NORMAL, slot );
	V2 normal = v2( nrm.x, nrm.y );
	V2 tangent = rightPerp( normal );
	float friction = nrm.z;
	float tangentSpeed = nrm.w;
	float4 soft = loadField( V, CF_SOFT, slot );
	float4 roll = loadField( V, CF_ROLL, slot );
	float4 pmass = loadField( V, CF_PMASS, slot );
	float4 base = loadField( V, CF_BASE, slot );
	float normalMass[2] = { pmass.x, pmass.z };
	float tangentMass[2] = { pmass.y, pmass.w };
	float baseSeparation[2] = { base.x, base.y };

	float4 imp[2] = { loadField( V, CF_IMP1, slot ), loadField( V, CF_IMP2, slot ) };
	float4 anc[2] = { loadField( V, CF_ANCHOR1, slot ), loadField( V, CF_ANCHOR2, slot ) };

	float totalNormalImpulse = 0.0f;

	// non-penetration
	for ( int j = 0; j < pointCount; ++j )
	{
		V2 rA = v2( anc[j].x, anc[j].y );
		V2 rB = v2( anc[j].z, anc[j].w );

		V2 ds = add( dp, sub( rotate( dqB, rB ), rotate( dqA, rA ) ) );
		float s = baseSeparation[j] + dot( ds, normal );

		float velocityBias = 0.0f;
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( s > 0.0f )
		{
			velocityBias = s * P.inv_h;
		}
		else if ( useBias )
		{
			velocityBias = maxf_( soft.y * soft.x * s, -P.contactSpeed );
			massScale = soft.y;
			impulseScale = soft.z;
		}

		V2 vrA = add( vA, crossSV( wA, rA ) );
		V2 vrB = add( vB, crossSV( wB, rB ) );
		float vn = dot( sub( vrB, vrA ), normal );

		float impulse = -normalMass[j] * ( massScale * vn + velocityBias ) - impulseScale * imp[j].x;

		float newImpulse = maxf_( imp[j].x + impulse, 0.0f );
		impulse = newImpulse - imp[j].x;
		imp[j].x = newImpulse;
		imp[j].z += impulse;

		totalNormalImpulse += newImpulse;

		V2 Pv = mulSV( impulse, normal );
		vA = mulSub( vA, mA, Pv );
		wA -= iA * cross( rA, Pv );
		vB = mulAdd( vB, mB, Pv );
		wB += iB * cross( rB, Pv );
	}

	if ( useBias == false )
	{
		// friction
		for ( int j = 0; j < pointCount; ++j )
		{
			V2 rA = v2( anc[j].x, anc[j].y );
			V2 rB = v2( anc[j].z, anc[j].w );

			V2 vrB = add( vB, crossSV( wB, rB ) );
			V2 vrA = add( vA, crossSV( wA, rA ) );

			float vt = dot( sub( vrB, vrA ), tangent ) - tangentSpeed;

			float impulse = tangentMass[j] * ( -vt );

			float maxFriction = friction * imp[j].x;
			float newImpulse = clampf_( imp[j].y + impulse, -maxFriction, maxFriction );
			impulse = newImpulse - imp[j].y;
			imp[j].y = newImpulse;

			V2 Pv = mulSV( impulse, tangent );
			vA = mulSub( vA, mA, Pv );
			wA -= iA * cross( rA, Pv );
			vB = mulAdd( vB, mB, Pv );
			wB += iB * cross( rB, Pv );
		}

		// rolling resistance
		{
			float deltaLambda = -roll.z * ( wB - wA );
			float lambda = imp[0].w;
			float maxLambda = roll.x * totalNormalImpulse;
			imp[0].w = clampf_( lambda + deltaLambda, -maxLambda, maxLambda );
			deltaLambda = imp[0].w - lambda;

			wA -= iA * deltaLambda;
			wB += iB * deltaLambda;
		}
	}

	storeField( V, CF_IMP1, slot, imp[0] );
	storeField( V, CF_IMP2, slot, imp[1] );
	scatterVel( V, idx.x, make_float4( vA.x, vA.y, wA, sA.w ) );
	scatterVel( V, idx.y, make_float4( vB.x, vB.y, wB, sB.w ) );
}

// b2ApplyRestitution_Overflow, src/contact_solver.c:410-514
B2G_DEV void restitutionContactOverflow( const StepParams& P, const SolveView& V, int slot )
{
	float4 roll = loadField( V, CF_ROLL, slot );
	float restitution = roll.y;
	if ( restitution == 0.0f )
	{
		return;
	}

	int2 idx = V.cidx[slot];
	int pointCount = V.cmeta[slot] & kMetaPointMask;
	float4 mass = loadField( V, CF_MASS, slot );
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;

	float4 sA = gatherVel( V, idx.x );
	float4 sB = gatherVel( V, idx.y );
	V2 vA = v2( sA.x, sA.y );
	float wA = sA.z;
	V2 vB = v2( sB.x, sB.y );
	float wB = sB.z;

	float4 nrm = loadField( V, CF_NORMAL, slot );
	V2 normal = v2( nrm.x, nrm.y );
	float4 pmass = loadField( V, CF_PMASS, slot );
	float4 base = loadField( V, CF_BASE, slot );
	float normalMass[2] = { pmass.x, pmass.z };
	float relativeVelocity[2] = { base.z, base.w };
	float4 imp[2] = { loadField( V, CF_IMP1, slot ), loadField( V, CF_IMP2, slot ) };
	float4 anc[2] = { loadField( V, CF_ANCHOR1, slot ), loadField( V, CF_ANCHOR2, slot ) };
	float threshold = P.restitutionThreshold;

	for ( int j = 0; j < pointCount; ++j )
	{
		if ( relativeVelocity[j] > -threshold || imp[j].z == 0.0f )
		{
			continue;
		}

		V2 rA = v2( anc[j].x, anc[j].y );
		V2 rB = v2( anc[j].z, anc[j].w );

		V2 vrB = add( vB, crossSV( wB, rB ) );
		V2 vrA = add( vA, crossSV( wA, rA ) );
		float vn = dot( sub( vrB, vrA ), normal );

		float impulse = -normalMass[j] * ( vn + restitution * relativeVelocity[j] );

		float newImpulse = maxf_( imp[j].x + impulse, 0.0f );
		impulse = newImpulse - imp[j].x;
		imp[j].x = newImpulse;
		imp[j].z += impulse;

		V2 Pv = mulSV( impulse, normal );
		vA = mulSub( vA, mA, Pv );
		wA -= iA * cross( rA, Pv );
		vB = mulAdd( vB, mB, Pv );
		wB += iB * cross( rB, Pv );
	}

	storeField( V, CF_IMP1, slot, imp[0] );
	storeField( V, CF_IMP2, slot, imp[1] );
	scatterVel( V, idx.x, make_float4( vA.x, vA.y, wA, sA.w ) );
	scatterVel( V, idx.y, make_float4( vB.x, vB.y, wB, sB.w ) );
}

// ===========================================================================================================
// A deep overflow chain walked by ONE WARP (contacts [begin, end) of the view, in order).
// ===========================================================================================================
// The sequence of updates a body sees must be the reference's (one contact after the other, array order), but only the
// part of a contact that READS OR WRITES BODY VELOCITIES is sequential.  Lane L takes contact base + L: it loads the
// constraint and evaluates everything that does not depend on velocities (anchors, separation, bias and softness per
// point, the warm-start impulse vectors) together with the other 31 lanes; then the lanes take turns -- gather {v, w},
// the dependent chain, scatter -- and finally all store their impulses together.  Every expression is the one of the
// scalar functions above, evaluated in the same order; only the loads and the velocity-independent terms move.
enum OverflowOp
{
	OV_WARM = 0,
	OV_SOLVE = 1, // useBias
	OV_RELAX = 2,
	OV_RESTITUTION = 3
};

template <int Op> B2G_DEV void overflowChainWarp( const StepParams& P, const SolveView& V, int begin, int end )
{
	const int lane = (int)( threadIdx.x & 31u );
	constexpr bool useBias = Op == OV_SOLVE;
	for ( int base = begin; base < end; base += 32 )
	{
		const int k = base + lane;
		bool valid = k < end;
		const int slot = valid ? k : begin;

		// ---- together: the constraint and what does not depend on velocities
		int2 idx = V.cidx[slot];
		int pointCount = V.cmeta[slot] & kMetaPointMask;
		float4 mass = loadField( V, CF_MASS, slot );
		float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;
		float4 nrm = loadField( V, CF_NORMAL, slot );
		V2 normal = v2( nrm.x, nrm.y );
		V2 tangent = rightPerp( normal );
		float friction = nrm.z;
		float tangentSpeed = nrm.w;
		float4 roll = loadField( V, CF_ROLL, slot );
		float4 imp[2] = { loadField( V, CF_IMP1, slot ), loadField( V, CF_IMP2, slot ) };
		float4 anc[2] = { loadField( V, CF_ANCHOR1, slot ), loadField( V, CF_ANCHOR2, slot ) };
		float normalMass[2] = { 0.0f, 0.0f }, tangentMass[2] = { 0.0f, 0.0f };
		float velocityBias[2] = { 0.0f, 0.0f }, massScale[2] = { 1.0f, 1.0f }, impulseScale[2] = { 0.0f, 0.0f };
		float relativeVelocity[2] = { 0.0f, 0.0f };
		bool skipPoint[2] = { false, false };
		V2 warmP[2] = { v2( 0.0f, 0.0f ), v2( 0.0f, 0.0f ) };

		if ( Op == OV_WARM )
		{
#pragma unroll
			for ( int j = 0; j < 2; ++j )
			{
				if ( j >= pointCount )
				{
					continue;
				}
				warmP[j] = add( mulSV( imp[j].x, normal ), mulSV( imp[j].y, tangent ) );
				imp[j].z += imp[j].x;
			}
		}
		else if ( Op == OV_RESTITUTION )
		{
			float4 pmass = loadField( V, CF_PMASS, slot );
			float4 baseField = loadField( V, CF_BASE, slot );
			normalMass[0] = pmass.x, normalMass[1] = pmass.z;
			relativeVelocity[0] = baseField.z, relativeVelocity[1] = baseField.w;
			valid = valid && !( roll.y == 0.0f );
			float threshold = P.restitutionThreshold;
#pragma unroll
			for ( int j = 0; j < 2; ++j )
			{
				skipPoint[j] = relativeVelocity[j] > -threshold || imp[j].z == 0.0f;
			}
		}
		else
		{
			float4 soft = loadField( V, CF_SOFT, slot );
			float4 pmass = loadField( V, CF_PMASS, slot );
			float4 baseField = loadField( V, CF_BASE, slot );
			normalMass[0] = pmass.x, normalMass[1] = pmass.z;
			tangentMass[0] = pmass.y, tangentMass[1] = pmass.w;
			float baseSeparation[2] = { baseField.x, baseField.y };
			float4 qA = gatherPos( V, idx.x );
			float4 qB = gatherPos( V, idx.y );
			Rot dqA, dqB;
			dqA.c = qA.z, dqA.s = qA.w;
			dqB.c = qB.z, dqB.s = qB.w;
			V2 dp = sub( v2( qB.x, qB.y ), v2( qA.x, qA.y ) );
#pragma unroll
			for ( int j = 0; j < 2; ++j )
			{
				if ( j >= pointCount )
				{
					continue;
				}
				V2 rA = v2( anc[j].x, anc[j].y );
				V2 rB = v2( anc[j].z, anc[j].w );
				V2 ds = add( dp, sub( rotate( dqB, rB ), rotate( dqA, rA ) ) );
				float sep = baseSeparation[j] + dot( ds, normal );
				if ( sep > 0.0f )
				{
					velocityBias[j] = sep * P.inv_h;
				}
				else if ( useBias )
				{
					velocityBias[j] = maxf_( soft.y * soft.x * sep, -P.contactSpeed );
					massScale[j] = soft.y;
					impulseScale[j] = soft.z;
				}
			}
		}

		// ---- one after the other: the part that reads and writes body velocities
		const int turns = end - base < 32 ? end - base : 32;
		for ( int turn = 0; turn < turns; ++turn )
		{
			if ( lane == turn && valid )
			{
				float4 sA = gatherVel( V, idx.x );
				float4 sB = gatherVel( V, idx.y );
				V2 vA = v2( sA.x, sA.y );
				float wA = sA.z;
				V2 vB = v2( sB.x, sB.y );
				float wB = sB.z;

				if ( Op == OV_WARM )
				{
#pragma unroll
					for ( int j = 0; j < 2; ++j )
					{
						if ( j >= pointCount )
						{
							continue;
						}
						V2 rA = v2( anc[j].x, anc[j].y );
						V2 rB = v2( anc[j].z, anc[j].w );
						V2 Pv = warmP[j];
						wA -= iA * cross( rA, Pv );
						vA = mulAdd( vA, -mA, Pv );
						wB += iB * cross( rB, Pv );
						vB = mulAdd( vB, mB, Pv );
					}
					wA -= iA * imp[0].w;
					wB += iB * imp[0].w;
				}
				else if ( Op == OV_RESTITUTION )
				{
					float restitution = roll.y;
#pragma unroll
					for ( int j = 0; j < 2; ++j )
					{
						if ( j >= pointCount )
						{
							continue;
						}
						if ( skipPoint[j] )
						{
							continue;
						}
						V2 rA = v2( anc[j].x, anc[j].y );
						V2 rB = v2( anc[j].z, anc[j].w );
						V2 vrB = add( vB, crossSV( wB, rB ) );
						V2 vrA = add( vA, crossSV( wA, rA ) );
						float vn = dot( sub( vrB, vrA ), normal );
						float impulse = -normalMass[j] * ( vn + restitution * relativeVelocity[j] );
						float newImpulse = maxf_( imp[j].x + impulse, 0.0f );
						impulse = newImpulse - imp[j].x;
						imp[j].x = newImpulse;
						imp[j].z += impulse;
						V2 Pv = mulSV( impulse, normal );
						vA = mulSub( vA, mA, Pv );
						wA -= iA * cross( rA, Pv );
						vB = mulAdd( vB, mB, Pv );
						wB += iB * cross( rB, Pv );
					}
				}
				else
				{
					float totalNormalImpulse = 0.0f;
#pragma unroll
					for ( int j = 0; j < 2; ++j )
					{
						if ( j >= pointCount )
						{
							continue;
						}
						V2 rA = v2( anc[j].x, anc[j].y );
						V2 rB = v2( anc[j].z, anc[j].w );
						V2 vrA = add( vA, crossSV( wA, rA ) );
						V2 vrB = add( vB, crossSV( wB, rB ) );
						float vn = dot( sub( vrB, vrA ), normal );
						float impulse = -normalMass[j] * ( massScale[j] * vn + velocityBias[j] ) - impulseScale[j] * imp[j].x;
						float newImpulse = maxf_( imp[j].x + impulse, 0.0f );
						impulse = newImpulse - imp[j].x;
						imp[j].x = newImpulse;
						imp[j].z += impulse;
						totalNormalImpulse += newImpulse;
						V2 Pv = mulSV( impulse, normal );
						vA = mulSub( vA, mA, Pv );
						wA -= iA * cross( rA, Pv );
						vB = mulAdd( vB, mB, Pv );
						wB += iB * cross( rB, Pv );
					}
					if ( useBias == false )
					{
#pragma unroll
						for ( int j = 0; j < 2; ++j )
						{
							if ( j >= pointCount )
							{
								continue;
							}
							V2 rA = v2( anc[j].x, anc[j].y );
							V2 rB = v2( anc[j].z, anc[j].w );
							V2 vrB = add( vB, crossSV( wB, rB ) );
							V2 vrA = add( vA, crossSV( wA, rA ) );
							float vt = dot( sub( vrB, vrA ), tangent ) - tangentSpeed;
							float impulse = tangentMass[j] * ( -vt );
							float maxFriction = friction * imp[j].x;
							float newImpulse = clampf_( imp[j].y + impulse, -maxFriction, maxFriction );
							impulse = newImpulse - imp[j].y;
							imp[j].y = newImpulse;
							V2 Pv = mulSV( impulse, tangent );
							vA = mulSub( vA, mA, Pv );
							wA -= iA * cross( rA, Pv );
							vB = mulAdd( vB, mB, Pv );
							wB += iB * cross( rB, Pv );
						}
						{
							float deltaLambda = -roll.z * ( wB - wA );
							float lambda = imp[0].w;
							float maxLambda = roll.x * totalNormalImpulse;
							imp[0].w = clampf_( lambda + deltaLambda, -maxLambda, maxLambda );
							deltaLambda = imp[0].w - lambda;
							wA -= iA * deltaLambda;
							wB += iB * deltaLambda;
						}
					}
				}
				scatterVel( V, idx.x, make_float4( vA.x, vA.y, wA, sA.w ) );
				scatterVel( V, idx.y, make_float4( vB.x, vB.y, wB, sB.w ) );
			}
			__syncwarp();
		}

		// ---- together again
		if ( valid )
		{
			storeField( V, CF_IMP1, slot, imp[0] );
			storeField( V, CF_IMP2, slot, imp[1] );
		}
	}
}

} // namespace b2g
