// b2g_contact.cuh -- contact constraint stages, one thread per constraint.
//
// Two families, exactly as in the reference (they use DIFFERENT operation orders):
//   *  coloured contacts follow the wide path   src/contact_solver.c:1573-2331 (b2ContactConstraintWide)
//   *  overflow contacts follow the scalar path src/contact_solver.c:24-545    (b2ContactConstraint)
// Expression association is kept operator by operator (see b2g_math.cuh for the compile flags).
#pragma once

#include "b2g_types.cuh"

#include "b2gpu_layout.h"

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
B2G_DEV float4 gatherVel( const StepParams& P, int index )
{
	return __ldcg( P.vel + index );
}

B2G_DEV float4 gatherPos( const StepParams& P, int index )
{
	return __ldcg( P.pos + index );
}

B2G_DEV void scatterVel( const StepParams& P, int index, float4 v )
{
	if ( ( __float_as_uint( v.w ) & B2L_FLAG_DYNAMIC ) != 0 )
	{
		__stcg( P.vel + index, v );
	}
}

B2G_DEV float4 loadField( const StepParams& P, int field, int slot )
{
	return P.cf[(size_t)field * P.slotCapacity + slot];
}

B2G_DEV void storeField( const StepParams& P, int field, int slot, float4 value )
{
	P.cf[(size_t)field * P.slotCapacity + slot] = value;
}

// ---- prepare ---------------------------------------------------------------------------------------------
// b2PrepareContactsTask (src/contact_solver.c:1573-1809) per lane; with wide == false it is
// b2PrepareContacts_Overflow (src/contact_solver.c:24-160): no contact-softening branch.
B2G_DEV void prepareContact( const StepParams& P, int slot, bool wide )
{
	const uint8_t* sim = P.rawContacts + (size_t)slot * B2L_CONTACT_SIZE;
	const uint8_t* manifold = sim + B2L_CONTACT_MANIFOLD;

	int indexA = rawI( sim, B2L_CONTACT_INDEX_A );
	int indexB = rawI( sim, B2L_CONTACT_INDEX_B );

	float mA = rawF( sim, B2L_CONTACT_INV_MASS_A );
	float iA = rawF( sim, B2L_CONTACT_INV_I_A );
	float mB = rawF( sim, B2L_CONTACT_INV_MASS_B );
	float iB = rawF( sim, B2L_CONTACT_INV_I_B );

	V2 vA = v2( 0.0f, 0.0f );
	float wA = 0.0f;
	if ( indexA != -1 )
	{
		const uint8_t* s = P.rawStates + (size_t)indexA * B2L_STATE_SIZE;
		vA = v2( rawF( s, 0 ), rawF( s, 4 ) );
		wA = rawF( s, 8 );
	}
	V2 vB = v2( 0.0f, 0.0f );
	float wB = 0.0f;
	if ( indexB != -1 )
	{
		const uint8_t* s = P.rawStates + (size_t)indexB * B2L_STATE_SIZE;
		vB = v2( rawF( s, 0 ), rawF( s, 4 ) );
		wB = rawF( s, 8 );
	}

	float rollingMass;
	{
		float k = iA + iB;
		rollingMass = k > 0.0f ? 1.0f / k : 0.0f;
	}

	Soft soft = P.contactSoft;
	if ( indexA == -1 || indexB == -1 )
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

	V2 normal = v2( rawF( manifold, B2L_MANIFOLD_NORMAL ), rawF( manifold, B2L_MANIFOLD_NORMAL + 4 ) );
	V2 tangent = rightPerp( normal );
	float friction = rawF( sim, B2L_CONTACT_FRICTION );
	float restitution = rawF( sim, B2L_CONTACT_RESTITUTION );
	float rollingResistance = rawF( sim, B2L_CONTACT_ROLLING_RESISTANCE );
	float tangentSpeed = rawF( sim, B2L_CONTACT_TANGENT_SPEED );
	float rollingImpulse = warmStartScale * rawF( manifold, B2L_MANIFOLD_ROLLING_IMPULSE );
	int pointCount = rawI( manifold, B2L_MANIFOLD_POINT_COUNT );

	float4 anchors[2], impulses[2];
	float normalMass[2], tangentMass[2], baseSeparation[2], relativeVelocity[2];

#pragma unroll
	for ( int j = 0; j < 2; ++j )
	{
		if ( j < pointCount )
		{
			const uint8_t* mp = manifold + B2L_MANIFOLD_POINTS + j * B2L_MP_SIZE;
			V2 rA = v2( rawF( mp, B2L_MP_ANCHOR_A ), rawF( mp, B2L_MP_ANCHOR_A + 4 ) );
			V2 rB = v2( rawF( mp, B2L_MP_ANCHOR_B ), rawF( mp, B2L_MP_ANCHOR_B + 4 ) );
			anchors[j] = make_float4( rA.x, rA.y, rB.x, rB.y );

			baseSeparation[j] = rawF( mp, B2L_MP_SEPARATION ) - dot( sub( rB, rA ), normal );

			impulses[j].x = warmStartScale * rawF( mp, B2L_MP_NORMAL_IMPULSE );
			impulses[j].y = warmStartScale * rawF( mp, B2L_MP_TANGENT_IMPULSE );
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

	storeField( P, CF_MASS, slot, make_float4( mA, iA, mB, iB ) );
	storeField( P, CF_NORMAL, slot, make_float4( normal.x, normal.y, friction, tangentSpeed ) );
	storeField( P, CF_ROLL, slot, make_float4( rollingResistance, restitution, rollingMass, 0.0f ) );
	storeField( P, CF_SOFT, slot, make_float4( soft.biasRate, soft.massScale, soft.impulseScale, 0.0f ) );
	storeField( P, CF_ANCHOR1, slot, anchors[0] );
	storeField( P, CF_ANCHOR2, slot, anchors[1] );
	storeField( P, CF_PMASS, slot, make_float4( normalMass[0], tangentMass[0], normalMass[1], tangentMass[1] ) );
	storeField( P, CF_BASE, slot,
				make_float4( baseSeparation[0], baseSeparation[1], relativeVelocity[0], relativeVelocity[1] ) );
	storeField( P, CF_IMP1, slot, impulses[0] );
	storeField( P, CF_IMP2, slot, impulses[1] );

	if ( !( restitution == 0.0f ) )
	{
		*P.anyRestitution = 1;
	}

	// 0 for null (contact_solver.c:1644-1646)
	P.cidx[slot] = make_int2( indexA + 1, indexB + 1 );
	int simFlags = rawI( sim, B2L_CONTACT_SIM_FLAGS );
	P.cmeta[slot] = make_int2( rawI( sim, B2L_CONTACT_ID ), ( simFlags & (int)B2L_SIM_ENABLE_HIT_EVENT ) | pointCount );
}

// ===========================================================================================================
// Wide path (coloured contacts)
// ===========================================================================================================

// b2WarmStartContactsTask, src/contact_solver.c:1811-1871
B2G_DEV void warmStartContact( const StepParams& P, int slot )
{
	int2 idx = P.cidx[slot];
	float4 bA = gatherVel( P, idx.x );
	float4 bB = gatherVel( P, idx.y );

	float4 mass = loadField( P, CF_MASS, slot );
	float invMassA = mass.x, invIA = mass.y, invMassB = mass.z, invIB = mass.w;
	float4 nrm = loadField( P, CF_NORMAL, slot );
	float nx = nrm.x, ny = nrm.y;
	float tangentX = ny;
	float tangentY = 0.0f - nx;

	float4 imp1 = loadField( P, CF_IMP1, slot );
	float4 imp2 = loadField( P, CF_IMP2, slot );

	{
		float4 a = loadField( P, CF_ANCHOR1, slot );
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
		float4 a = loadField( P, CF_ANCHOR2, slot );
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

	storeField( P, CF_IMP1, slot, imp1 );
	storeField( P, CF_IMP2, slot, imp2 );
	scatterVel( P, idx.x, bA );
	scatterVel( P, idx.y, bB );
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

// b2SolveContactsTask, src/contact_solver.c:1873-2116.  `active` is false for the padding lanes of the last
// warp of a colour: they take part in the warp votes only.  The reference skips rolling resistance for a
// whole SIMD register when all its lanes have none (contact_solver.c:2021); the default build is SSE2 with
// 4 lanes, so the vote is taken over aligned groups of 4 consecutive constraints of the colour.
B2G_DEV void solveContact( const StepParams& P, int slot, bool active, bool useBias, unsigned lane )
{
	float4 roll = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
	if ( active && useBias == false )
	{
		roll = loadField( P, CF_ROLL, slot );
	}

	bool groupHasRolling = false;
	if ( useBias == false )
	{
		// x == 0 is false for NaN, like _mm_cmpeq_ps (ordered)
		unsigned nonZero = __ballot_sync( 0xffffffffu, !( roll.x == 0.0f ) );
		groupHasRolling = ( ( nonZero >> ( lane & ~3u ) ) & 0xFu ) != 0;
	}

	if ( active == false )
	{
		return;
	}

	int2 idx = P.cidx[slot];
	float4 bA = gatherVel( P, idx.x );
	float4 bB = gatherVel( P, idx.y );
	float4 pA = gatherPos( P, idx.x );
	float4 pB = gatherPos( P, idx.y );

	float4 mass = loadField( P, CF_MASS, slot );
	float invMassA = mass.x, invIA = mass.y, invMassB = mass.z, invIB = mass.w;
	float4 nrm = loadField( P, CF_NORMAL, slot );
	float nx = nrm.x, ny = nrm.y;
	float4 soft = loadField( P, CF_SOFT, slot );
	float4 pmass = loadField( P, CF_PMASS, slot );
	float4 base = loadField( P, CF_BASE, slot );
	float4 a1 = loadField( P, CF_ANCHOR1, slot );
	float4 a2 = loadField( P, CF_ANCHOR2, slot );
	float4 imp1 = loadField( P, CF_IMP1, slot );
	float4 imp2 = loadField( P, CF_IMP2, slot );

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

	storeField( P, CF_IMP1, slot, imp1 );
	storeField( P, CF_IMP2, slot, imp2 );
	scatterVel( P, idx.x, bA );
	scatterVel( P, idx.y, bB );
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
B2G_DEV void restitutionContact( const StepParams& P, int slot, bool active, unsigned lane )
{
	float4 roll = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
	if ( active )
	{
		roll = loadField( P, CF_ROLL, slot );
	}
	float restitution = roll.y;
	unsigned nonZero = __ballot_sync( 0xffffffffu, !( restitution == 0.0f ) );
	bool groupHasRestitution = ( ( nonZero >> ( lane & ~3u ) ) & 0xFu ) != 0;
	if ( active == false || groupHasRestitution == false )
	{
		return;
	}

	bool restitutionIsZero = restitution == 0.0f;

	int2 idx = P.cidx[slot];
	float4 bA = gatherVel( P, idx.x );
	float4 bB = gatherVel( P, idx.y );

	float4 mass = loadField( P, CF_MASS, slot );
	float4 nrm = loadField( P, CF_NORMAL, slot );
	float4 pmass = loadField( P, CF_PMASS, slot );
	float4 base = loadField( P, CF_BASE, slot );
	float4 a1 = loadField( P, CF_ANCHOR1, slot );
	float4 a2 = loadField( P, CF_ANCHOR2, slot );
	float4 imp1 = loadField( P, CF_IMP1, slot );
	float4 imp2 = loadField( P, CF_IMP2, slot );

	restitutionRow( bA, bB, a1, nrm.x, nrm.y, restitution, restitutionIsZero, P.restitutionThreshold, base.z, pmass.x, mass.x,
					mass.y, mass.z, mass.w, imp1.x, imp1.z );
	restitutionRow( bA, bB, a2, nrm.x, nrm.y, restitution, restitutionIsZero, P.restitutionThreshold, base.w, pmass.z, mass.x,
					mass.y, mass.z, mass.w, imp2.x, imp2.z );

	storeField( P, CF_IMP1, slot, imp1 );
	storeField( P, CF_IMP2, slot, imp2 );
	scatterVel( P, idx.x, bA );
	scatterVel( P, idx.y, bB );
}

// b2StoreImpulsesTask, src/contact_solver.c:2238-2331 (wide == true) and b2StoreImpulses_Overflow :516-545
// (wide == false: no hit-event test).  The 9 floats go to a packed record; the host scatters them into
// b2Manifold (the overflow path only writes the first pointCount points, the host honours that).
B2G_DEV void storeContact( const StepParams& P, int slot, bool wide )
{
	float4 imp1 = loadField( P, CF_IMP1, slot );
	float4 imp2 = loadField( P, CF_IMP2, slot );
	float4 base = loadField( P, CF_BASE, slot );

	float* out = P.outImpulses + (size_t)slot * kImpulseFloats;
	out[0] = imp1.w; // rollingImpulse
	out[1] = imp1.x; // normalImpulse
	out[2] = imp1.y; // tangentImpulse
	out[3] = imp1.z; // totalNormalImpulse
	out[4] = base.z; // normalVelocity
	out[5] = imp2.x;
	out[6] = imp2.y;
	out[7] = imp2.z;
	out[8] = base.w;

	if ( wide )
	{
		int2 meta = P.cmeta[slot];
		if ( ( meta.y & (int)B2L_SIM_ENABLE_HIT_EVENT ) != 0 )
		{
			int pointCount = meta.y & 3;
			float negHitThreshold = -P.hitEventThreshold;
			bool hit = ( base.z < negHitThreshold && imp1.z > 0.0f );
			if ( pointCount > 1 )
			{
				hit = hit || ( base.w < negHitThreshold && imp2.z > 0.0f );
			}
			if ( hit )
			{
				unsigned id = (unsigned)meta.x;
				atomicOr( P.hitBits + ( id >> 5 ), 1u << ( id & 31u ) );
				*P.hasHitEvents = 1;
			}
		}
	}
}

// ===========================================================================================================
// Scalar path (overflow colour): strictly sequential in array order, one thread.
// ===========================================================================================================

// b2WarmStartContacts_Overflow, src/contact_solver.c:162-237
B2G_DEV void warmStartContactOverflow( const StepParams& P, int slot )
{
	int2 idx = P.cidx[slot];
	int pointCount = P.cmeta[slot].y & 3;
	float4 sA = gatherVel( P, idx.x );
	float4 sB = gatherVel( P, idx.y );
	V2 vA = v2( sA.x, sA.y );
	float wA = sA.z;
	V2 vB = v2( sB.x, sB.y );
	float wB = sB.z;

	float4 mass = loadField( P, CF_MASS, slot );
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;
	float4 nrm = loadField( P, CF_NORMAL, slot );
	V2 normal = v2( nrm.x, nrm.y );
	V2 tangent = rightPerp( normal );

	float4 imp[2] = { loadField( P, CF_IMP1, slot ), loadField( P, CF_IMP2, slot ) };
	float4 anc[2] = { loadField( P, CF_ANCHOR1, slot ), loadField( P, CF_ANCHOR2, slot ) };

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

	storeField( P, CF_IMP1, slot, imp[0] );
	storeField( P, CF_IMP2, slot, imp[1] );
	scatterVel( P, idx.x, make_float4( vA.x, vA.y, wA, sA.w ) );
	scatterVel( P, idx.y, make_float4( vB.x, vB.y, wB, sB.w ) );
}

// b2SolveContacts_Overflow, src/contact_solver.c:239-408 (friction BEFORE rolling resistance)
B2G_DEV void solveContactOverflow( const StepParams& P, int slot, bool useBias )
{
	int2 idx = P.cidx[slot];
	int pointCount = P.cmeta[slot].y & 3;

	float4 mass = loadField( P, CF_MASS, slot );
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;

	float4 sA = gatherVel( P, idx.x );
	float4 qA = gatherPos( P, idx.x );
	V2 vA = v2( sA.x, sA.y );
	float wA = sA.z;
	Rot dqA;
	dqA.c = qA.z;
	dqA.s = qA.w;

	float4 sB = gatherVel( P, idx.y );
	float4 qB = gatherPos( P, idx.y );
	V2 vB = v2( sB.x, sB.y );
	float wB = sB.z;
	Rot dqB;
	dqB.c = qB.z;
	dqB.s = qB.w;

	V2 dp = sub( v2( qB.x, qB.y ), v2( qA.x, qA.y ) );

	float4 nrm = loadField( P, CF_NORMAL, slot );
	V2 normal = v2( nrm.x, nrm.y );
	V2 tangent = rightPerp( normal );
	float friction = nrm.z;
	float tangentSpeed = nrm.w;
	float4 soft = loadField( P, CF_SOFT, slot );
	float4 roll = loadField( P, CF_ROLL, slot );
	float4 pmass = loadField( P, CF_PMASS, slot );
	float4 base = loadField( P, CF_BASE, slot );
	float normalMass[2] = { pmass.x, pmass.z };
	float tangentMass[2] = { pmass.y, pmass.w };
	float baseSeparation[2] = { base.x, base.y };

	float4 imp[2] = { loadField( P, CF_IMP1, slot ), loadField( P, CF_IMP2, slot ) };
	float4 anc[2] = { loadField( P, CF_ANCHOR1, slot ), loadField( P, CF_ANCHOR2, slot ) };

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

	storeField( P, CF_IMP1, slot, imp[0] );
	storeField( P, CF_IMP2, slot, imp[1] );
	scatterVel( P, idx.x, make_float4( vA.x, vA.y, wA, sA.w ) );
	scatterVel( P, idx.y, make_float4( vB.x, vB.y, wB, sB.w ) );
}

// b2ApplyRestitution_Overflow, src/contact_solver.c:410-514
B2G_DEV void restitutionContactOverflow( const StepParams& P, int slot )
{
	float4 roll = loadField( P, CF_ROLL, slot );
	float restitution = roll.y;
	if ( restitution == 0.0f )
	{
		return;
	}

	int2 idx = P.cidx[slot];
	int pointCount = P.cmeta[slot].y & 3;
	float4 mass = loadField( P, CF_MASS, slot );
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;

	float4 sA = gatherVel( P, idx.x );
	float4 sB = gatherVel( P, idx.y );
	V2 vA = v2( sA.x, sA.y );
	float wA = sA.z;
	V2 vB = v2( sB.x, sB.y );
	float wB = sB.z;

	float4 nrm = loadField( P, CF_NORMAL, slot );
	V2 normal = v2( nrm.x, nrm.y );
	float4 pmass = loadField( P, CF_PMASS, slot );
	float4 base = loadField( P, CF_BASE, slot );
	float normalMass[2] = { pmass.x, pmass.z };
	float relativeVelocity[2] = { base.z, base.w };
	float4 imp[2] = { loadField( P, CF_IMP1, slot ), loadField( P, CF_IMP2, slot ) };
	float4 anc[2] = { loadField( P, CF_ANCHOR1, slot ), loadField( P, CF_ANCHOR2, slot ) };
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

	storeField( P, CF_IMP1, slot, imp[0] );
	storeField( P, CF_IMP2, slot, imp[1] );
	scatterVel( P, idx.x, make_float4( vA.x, vA.y, wA, sA.w ) );
	scatterVel( P, idx.y, make_float4( vB.x, vB.y, wB, sB.w ) );
}

} // namespace b2g
