// b2g_joint.cuh -- joint warm-start / solve / relax, one thread per joint, all eight joint types.
//
// The joint constants, accumulated impulses and host-prepared per-step data live in the reference's own
// b2JointSim record (252 B, mirrored by b2lJointSim in include/b2gpu_layout.h); the device solves it in place
// in a working copy, exactly like the reference solves it in place on the host (src/solver.h:73-74).
// Dispatch mirrors b2WarmStartJoint / b2SolveJoint (src/joint.c:1454-1540); every sub-constraint keeps the
// reference's order and association.
#pragma once

#include "b2g_contact.cuh"

namespace b2g
{

struct JointBodies
{
	int a, b;		  // 1-based indices into vel/pos, 0 = static dummy
	float4 vA, vB;	  // v.x v.y w flags
	float4 pA, pB;	  // dp.x dp.y dq.c dq.s
};

B2G_DEV V2 toV2( b2lVec2 v )
{
	return v2( v.x, v.y );
}

B2G_DEV Rot toRot( b2lRot q )
{
	Rot r;
	r.c = q.c;
	r.s = q.s;
	return r;
}

B2G_DEV Rot deltaRot( float4 p )
{
	Rot r;
	r.c = p.z;
	r.s = p.w;
	return r;
}

B2G_DEV JointBodies gatherJointBodies( const SolveView& V, int indexA, int indexB )
{
	JointBodies jb;
	jb.a = indexA + 1; // B2_NULL_INDEX (-1) -> dummy
	jb.b = indexB + 1;
	jb.vA = gatherVel( V, jb.a );
	jb.vB = gatherVel( V, jb.b );
	jb.pA = gatherPos( V, jb.a );
	jb.pB = gatherPos( V, jb.b );
	return jb;
}

B2G_DEV void scatterJointBodies( const SolveView& V, const JointBodies& jb, V2 vA, float wA, V2 vB, float wB )
{
	scatterVel( V, jb.a, make_float4( vA.x, vA.y, wA, jb.vA.w ) );
	scatterVel( V, jb.b, make_float4( vB.x, vB.y, wB, jb.vB.w ) );
}

// Warm start helper shared by the types whose warm start is "apply linear impulse L at anchors + angular
// impulse" written against the state directly (reference writes state->x -= ... under the dynamic flag).
B2G_DEV void applyWarmStart( const SolveView& V, const JointBodies& jb, float mA, float iA, float mB, float iB, V2 linear, float LA,
							 float LB )
{
	V2 vA = mulSub( v2( jb.vA.x, jb.vA.y ), mA, linear );
	float wA = jb.vA.z - iA * LA;
	V2 vB = mulAdd( v2( jb.vB.x, jb.vB.y ), mB, linear );
	float wB = jb.vB.z + iB * LB;
	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// ---- revolute (src/revolute_joint.c:283-500) ----------------------------------------------------------------
B2G_DEV void warmStartRevolute( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lRevolute* j = &base->u.revolute;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );
	V2 rA = rotate( deltaRot( jb.pA ), toV2( j->frameA.p ) );
	V2 rB = rotate( deltaRot( jb.pB ), toV2( j->frameB.p ) );
	float axialImpulse = j->springImpulse + j->motorImpulse + j->lowerImpulse - j->upperImpulse;
	V2 L = toV2( j->linearImpulse );
	applyWarmStart( V, jb, base->invMassA, base->invIA, base->invMassB, base->invIB, L, cross( rA, L ) + axialImpulse,
					cross( rB, L ) + axialImpulse );
}

B2G_DEV void solveRevolute( const StepParams& P, const SolveView& V, b2lJointSim* base, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB;
	float iA = base->invIA, iB = base->invIB;
	b2lRevolute* j = &base->u.revolute;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	Rot qA = mulRot( dqA, toRot( j->frameA.q ) );
	Rot qB = mulRot( dqB, toRot( j->frameB.q ) );
	Rot relQ = invMulRot( qA, qB );

	bool fixedRotation = ( iA + iB == 0.0f );
	Soft cs;
	cs.biasRate = base->constraintSoftness.biasRate;
	cs.massScale = base->constraintSoftness.massScale;
	cs.impulseScale = base->constraintSoftness.impulseScale;

	// spring
	if ( j->enableSpring && fixedRotation == false )
	{
		float jointAngle = rotAngle( relQ );
		float jointAngleDelta = unwindAngle( jointAngle - j->targetAngle );

		float C = jointAngleDelta;
		float bias = j->springSoftness.biasRate * C;
		float massScale = j->springSoftness.massScale;
		float impulseScale = j->springSoftness.impulseScale;

		float Cdot = wB - wA;
		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->springImpulse;
		j->springImpulse += impulse;

		wA -= iA * impulse;
		wB += iB * impulse;
	}

	// motor
	if ( j->enableMotor && fixedRotation == false )
	{
		float Cdot = wB - wA - j->motorSpeed;
		float impulse = -j->axialMass * Cdot;
		float oldImpulse = j->motorImpulse;
		float maxImpulse = P.h * j->maxMotorTorque;
		float newImpulse = clampf_( oldImpulse + impulse, -maxImpulse, maxImpulse );
		j->motorImpulse = newImpulse;
		impulse = newImpulse - oldImpulse;

		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if ( j->enableLimit && fixedRotation == false )
	{
		float jointAngle = rotAngle( relQ );

		// lower limit
		{
			float C = jointAngle - j->lowerAngle;
			float bias = 0.0f;
			float massScale = 1.0f;
			float impulseScale = 0.0f;
			if ( C > 0.0f )
			{
				bias = C * P.inv_h;
			}
			else if ( useBias )
			{
				bias = cs.biasRate * C;
				massScale = cs.massScale;
				impulseScale = cs.impulseScale;
			}

			float Cdot = wB - wA;
			float oldImpulse = j->lowerImpulse;
			float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * oldImpulse;
			float newImpulse = maxf_( oldImpulse + impulse, 0.0f );
			j->lowerImpulse = newImpulse;
			impulse = newImpulse - oldImpulse;

			wA -= iA * impulse;
			wB += iB * impulse;
		}

		// upper limit, signs flipped
		{
			float C = j->upperAngle - jointAngle;
			float bias = 0.0f;
			float massScale = 1.0f;
			float impulseScale = 0.0f;
			if ( C > 0.0f )
			{
				bias = C * P.inv_h;
			}
			else if ( useBias )
			{
				bias = cs.biasRate * C;
				massScale = cs.massScale;
				impulseScale = cs.impulseScale;
			}

			float Cdot = wA - wB;
			float oldImpulse = j->upperImpulse;
			float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * oldImpulse;
			float newImpulse = maxf_( oldImpulse + impulse, 0.0f );
			j->upperImpulse = newImpulse;
			impulse = newImpulse - oldImpulse;

			wA += iA * impulse;
			wB -= iB * impulse;
		}
	}

	// point to point
	{
		V2 rA = rotate( dqA, toV2( j->frameA.p ) );
		V2 rB = rotate( dqB, toV2( j->frameB.p ) );

		V2 Cdot = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );

		V2 bias = v2( 0.0f, 0.0f );
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( useBias )
		{
			V2 dcA = v2( jb.pA.x, jb.pA.y );
			V2 dcB = v2( jb.pB.x, jb.pB.y );
			V2 separation = add( add( sub( dcB, dcA ), sub( rB, rA ) ), toV2( j->deltaCenter ) );
			bias = mulSV( cs.biasRate, separation );
			massScale = cs.massScale;
			impulseScale = cs.impulseScale;
		}

		float k11 = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		float k12 = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		float k21 = k12;
		float k22 = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		V2 b = solve22( k11, k12, k21, k22, add( Cdot, bias ) );

		V2 impulse;
		impulse.x = -massScale * b.x - impulseScale * j->linearImpulse.x;
		impulse.y = -massScale * b.y - impulseScale * j->linearImpulse.y;
		j->linearImpulse.x += impulse.x;
		j->linearImpulse.y += impulse.y;

		vA = mulSub( vA, mA, impulse );
		wA -= iA * cross( rA, impulse );
		vB = mulAdd( vB, mB, impulse );
		wB += iB * cross( rB, impulse );
	}

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// ---- weld (src/weld_joint.c:222-453, non-block path: B2_WELD_BLOCK_SOLVE 0) ------------------------------------
B2G_DEV void warmStartWeld( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lWeld* j = &base->u.weld;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );
	V2 rA = rotate( deltaRot( jb.pA ), toV2( j->frameA.p ) );
	V2 rB = rotate( deltaRot( jb.pB ), toV2( j->frameB.p ) );
	V2 L = toV2( j->linearImpulse );
	applyWarmStart( V, jb, base->invMassA, base->invIA, base->invMassB, base->invIB, L, cross( rA, L ) + j->angularImpulse,
					cross( rB, L ) + j->angularImpulse );
}

B2G_DEV void solveWeld( const StepParams& P, const SolveView& V, b2lJointSim* base, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB;
	float iA = base->invIA, iB = base->invIB;
	b2lWeld* j = &base->u.weld;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	// angular constraint
	{
		Rot qA = mulRot( dqA, toRot( j->frameA.q ) );
		Rot qB = mulRot( dqB, toRot( j->frameB.q ) );
		Rot relQ = invMulRot( qA, qB );
		float jointAngle = rotAngle( relQ );

		float bias = 0.0f;
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( useBias || j->angularHertz > 0.0f )
		{
			float C = jointAngle;
			bias = j->angularSpring.biasRate * C;
			massScale = j->angularSpring.massScale;
			impulseScale = j->angularSpring.impulseScale;
		}

		float Cdot = wB - wA;
		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->angularImpulse;
		j->angularImpulse += impulse;

		wA -= iA * impulse;
		wB += iB * impulse;
	}

	// linear constraint
	{
		V2 rA = rotate( dqA, toV2( j->frameA.p ) );
		V2 rB = rotate( dqB, toV2( j->frameB.p ) );

		V2 bias = v2( 0.0f, 0.0f );
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( useBias || j->linearHertz > 0.0f )
		{
			V2 dcA = v2( jb.pA.x, jb.pA.y );
			V2 dcB = v2( jb.pB.x, jb.pB.y );
			V2 C = add( add( sub( dcB, dcA ), sub( rB, rA ) ), toV2( j->deltaCenter ) );

			bias = mulSV( j->linearSpring.biasRate, C );
			massScale = j->linearSpring.massScale;
			impulseScale = j->linearSpring.impulseScale;
		}

		V2 Cdot = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );

		float k11 = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		float k12 = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		float k21 = k12;
		float k22 = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		V2 b = solve22( k11, k12, k21, k22, add( Cdot, bias ) );

		V2 impulse;
		impulse.x = -massScale * b.x - impulseScale * j->linearImpulse.x;
		impulse.y = -massScale * b.y - impulseScale * j->linearImpulse.y;

		j->linearImpulse.x = j->linearImpulse.x + impulse.x;
		j->linearImpulse.y = j->linearImpulse.y + impulse.y;

		vA = mulSub( vA, mA, impulse );
		wA -= iA * cross( rA, impulse );
		vB = mulAdd( vB, mB, impulse );
		wB += iB * cross( rB, impulse );
	}

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// ---- prismatic (src/prismatic_joint.c:353-666) -----------------------------------------------------------------
B2G_DEV void warmStartPrismatic( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lPrismatic* j = &base->u.prismatic;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	V2 rA = rotate( dqA, toV2( j->frameA.p ) );
	V2 rB = rotate( dqB, toV2( j->frameB.p ) );

	V2 d = add( add( sub( v2( jb.pB.x, jb.pB.y ), v2( jb.pA.x, jb.pA.y ) ), toV2( j->deltaCenter ) ), sub( rB, rA ) );

	V2 axisA = rotate( toRot( j->frameA.q ), v2( 1.0f, 0.0f ) );
	axisA = rotate( dqA, axisA );

	float a1 = cross( add( rA, d ), axisA );
	float a2 = cross( rB, axisA );
	float axialImpulse = j->springImpulse + j->motorImpulse + j->lowerImpulse - j->upperImpulse;

	V2 perpA = leftPerp( axisA );
	float s1 = cross( add( rA, d ), perpA );
	float s2 = cross( rB, perpA );
	float perpImpulse = j->impulse.x;
	float angleImpulse = j->impulse.y;

	V2 Pv = add( mulSV( axialImpulse, axisA ), mulSV( perpImpulse, perpA ) );
	float LA = axialImpulse * a1 + perpImpulse * s1 + angleImpulse;
	float LB = axialImpulse * a2 + perpImpulse * s2 + angleImpulse;

	applyWarmStart( V, jb, base->invMassA, base->invIA, base->invMassB, base->invIB, Pv, LA, LB );
}

B2G_DEV void solvePrismatic( const StepParams& P, const SolveView& V, b2lJointSim* base, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB;
	float iA = base->invIA, iB = base->invIB;
	b2lPrismatic* j = &base->u.prismatic;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	Rot qA = mulRot( dqA, toRot( j->frameA.q ) );
	Rot qB = mulRot( dqB, toRot( j->frameB.q ) );
	Rot relQ = invMulRot( qA, qB );

	V2 rA = rotate( dqA, toV2( j->frameA.p ) );
	V2 rB = rotate( dqB, toV2( j->frameB.p ) );

	V2 d = add( add( sub( v2( jb.pB.x, jb.pB.y ), v2( jb.pA.x, jb.pA.y ) ), toV2( j->deltaCenter ) ), sub( rB, rA ) );

	V2 axisA = rotate( toRot( j->frameA.q ), v2( 1.0f, 0.0f ) );
	axisA = rotate( dqA, axisA );
	float translation = dot( axisA, d );

	float a1 = cross( add( rA, d ), axisA );
	float a2 = cross( rB, axisA );

	float k = mA + mB + iA * a1 * a1 + iB * a2 * a2;
	float axialMass = k > 0.0f ? 1.0f / k : 0.0f;

	Soft softness;
	softness.biasRate = base->constraintSoftness.biasRate;
	softness.massScale = base->constraintSoftness.massScale;
	softness.impulseScale = base->constraintSoftness.impulseScale;

	if ( j->enableSpring )
	{
		float C = translation - j->targetTranslation;
		float bias = j->springSoftness.biasRate * C;
		float massScale = j->springSoftness.massScale;
		float impulseScale = j->springSoftness.impulseScale;

		float Cdot = dot( axisA, sub( vB, vA ) ) + a2 * wB - a1 * wA;
		float deltaImpulse = -massScale * axialMass * ( Cdot + bias ) - impulseScale * j->springImpulse;
		j->springImpulse += deltaImpulse;

		V2 Pv = mulSV( deltaImpulse, axisA );
		float LA = deltaImpulse * a1;
		float LB = deltaImpulse * a2;

		vA = mulSub( vA, mA, Pv );
		wA -= iA * LA;
		vB = mulAdd( vB, mB, Pv );
		wB += iB * LB;
	}

	if ( j->enableMotor )
	{
		float Cdot = dot( axisA, sub( vB, vA ) ) + a2 * wB - a1 * wA;
		float impulse = axialMass * ( j->motorSpeed - Cdot );
		float oldImpulse = j->motorImpulse;
		float maxImpulse = P.h * j->maxMotorForce;
		float newImpulse = clampf_( oldImpulse + impulse, -maxImpulse, maxImpulse );
		j->motorImpulse = newImpulse;
		impulse = newImpulse - oldImpulse;

		V2 Pv = mulSV( impulse, axisA );
		float LA = impulse * a1;
		float LB = impulse * a2;

		vA = mulSub( vA, mA, Pv );
		wA -= iA * LA;
		vB = mulAdd( vB, mB, Pv );
		wB += iB * LB;
	}

	if ( j->enableLimit )
	{
		float speculativeDistance = 0.25f * ( j->upperTranslation - j->lowerTranslation );

		// lower limit
		{
			float C = translation - j->lowerTranslation;

			if ( C < speculativeDistance )
			{
				float bias = 0.0f;
				float massScale = 1.0f;
				float impulseScale = 0.0f;

				if ( C > 0.0f )
				{
					float safe = P.lengthUnitsPerMeter;
					bias = minf_( C, safe ) * P.inv_h;
				}
				else if ( useBias )
				{
					bias = softness.biasRate * C;
					massScale = softness.massScale;
					impulseScale = softness.impulseScale;
				}

				float oldImpulse = j->lowerImpulse;
				float Cdot = dot( axisA, sub( vB, vA ) ) + a2 * wB - a1 * wA;
				float deltaImpulse = -axialMass * massScale * ( Cdot + bias ) - impulseScale * oldImpulse;
				float newImpulse = maxf_( oldImpulse + deltaImpulse, 0.0f );
				j->lowerImpulse = newImpulse;
				deltaImpulse = newImpulse - oldImpulse;

				V2 Pv = mulSV( deltaImpulse, axisA );
				float LA = deltaImpulse * a1;
				float LB = deltaImpulse * a2;

				vA = mulSub( vA, mA, Pv );
				wA -= iA * LA;
				vB = mulAdd( vB, mB, Pv );
				wB += iB * LB;
			}
			else
			{
				j->lowerImpulse = 0.0f;
			}
		}

		// upper limit, signs flipped
		{
			float C = j->upperTranslation - translation;

			if ( C < speculativeDistance )
			{
				float bias = 0.0f;
				float massScale = 1.0f;
				float impulseScale = 0.0f;

				if ( C > 0.0f )
				{
					float safe = P.lengthUnitsPerMeter;
					bias = minf_( C, safe ) * P.inv_h;
				}
				else if ( useBias )
				{
					bias = softness.biasRate * C;
					massScale = softness.massScale;
					impulseScale = softness.impulseScale;
				}

				float oldImpulse = j->upperImpulse;
				float Cdot = dot( axisA, sub( vA, vB ) ) + a1 * wA - a2 * wB;
				float deltaImpulse = -axialMass * massScale * ( Cdot + bias ) - impulseScale * oldImpulse;
				float newImpulse = maxf_( oldImpulse + deltaImpulse, 0.0f );
				j->upperImpulse = newImpulse;
				deltaImpulse = newImpulse - oldImpulse;

				V2 Pv = mulSV( deltaImpulse, axisA );
				float LA = deltaImpulse * a1;
				float LB = deltaImpulse * a2;

				vA = mulAdd( vA, mA, Pv );
				wA += iA * LA;
				vB = mulSub( vB, mB, Pv );
				wB -= iB * LB;
			}
			else
			{
				j->upperImpulse = 0.0f;
			}
		}
	}

	// prismatic constraint: perpendicular + angle, 2x2 block
	{
		V2 perpA = leftPerp( axisA );

		float s1 = cross( add( d, rA ), perpA );
		float s2 = cross( rB, perpA );

		V2 Cdot;
		Cdot.x = dot( perpA, sub( vB, vA ) ) + s2 * wB - s1 * wA;
		Cdot.y = wB - wA;

		V2 bias = v2( 0.0f, 0.0f );
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( useBias )
		{
			V2 C;
			C.x = dot( perpA, d );
			C.y = rotAngle( relQ );

			bias = mulSV( softness.biasRate, C );
			massScale = softness.massScale;
			impulseScale = softness.impulseScale;
		}

		float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
		float k12 = iA * s1 + iB * s2;
		float k22 = iA + iB;
		if ( k22 == 0.0f )
		{
			k22 = 1.0f;
		}

		// K = { {k11, k12}, {k12, k22} } : cx = (k11,k12), cy = (k12,k22)
		V2 b = solve22( k11, k12, k12, k22, add( Cdot, bias ) );

		V2 deltaImpulse;
		deltaImpulse.x = -massScale * b.x - impulseScale * j->impulse.x;
		deltaImpulse.y = -massScale * b.y - impulseScale * j->impulse.y;

		j->impulse.x += deltaImpulse.x;
		j->impulse.y += deltaImpulse.y;

		V2 Pv = mulSV( deltaImpulse.x, perpA );
		float LA = deltaImpulse.x * s1 + deltaImpulse.y;
		float LB = deltaImpulse.x * s2 + deltaImpulse.y;

		vA = mulSub( vA, mA, Pv );
		wA -= iA * LA;
		vB = mulAdd( vB, mB, Pv );
		wB += iB * LB;
	}

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// ---- wheel (src/wheel_joint.c:280-523) -------------------------------------------------------------------------
B2G_DEV void warmStartWheel( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lWheel* j = &base->u.wheel;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	V2 rA = rotate( dqA, toV2( j->frameA.p ) );
	V2 rB = rotate( dqB, toV2( j->frameB.p ) );

	V2 d = add( add( sub( v2( jb.pB.x, jb.pB.y ), v2( jb.pA.x, jb.pA.y ) ), toV2( j->deltaCenter ) ), sub( rB, rA ) );
	V2 axisA = rotate( toRot( j->frameA.q ), v2( 1.0f, 0.0f ) );
	axisA = rotate( dqA, axisA );
	V2 perpA = leftPerp( axisA );

	float a1 = cross( add( d, rA ), axisA );
	float a2 = cross( rB, axisA );
	float s1 = cross( add( d, rA ), perpA );
	float s2 = cross( rB, perpA );

	float axialImpulse = j->springImpulse + j->lowerImpulse - j->upperImpulse;

	V2 Pv = add( mulSV( axialImpulse, axisA ), mulSV( j->perpImpulse, perpA ) );
	float LA = axialImpulse * a1 + j->perpImpulse * s1 + j->motorImpulse;
	float LB = axialImpulse * a2 + j->perpImpulse * s2 + j->motorImpulse;

	applyWarmStart( V, jb, base->invMassA, base->invIA, base->invMassB, base->invIB, Pv, LA, LB );
}

B2G_DEV void solveWheel( const StepParams& P, const SolveView& V, b2lJointSim* base, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB;
	float iA = base->invIA, iB = base->invIB;
	b2lWheel* j = &base->u.wheel;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	bool fixedRotation = ( iA + iB == 0.0f );

	V2 rA = rotate( dqA, toV2( j->frameA.p ) );
	V2 rB = rotate( dqB, toV2( j->frameB.p ) );

	V2 d = add( add( sub( v2( jb.pB.x, jb.pB.y ), v2( jb.pA.x, jb.pA.y ) ), toV2( j->deltaCenter ) ), sub( rB, rA ) );
	V2 axisA = rotate( toRot( j->frameA.q ), v2( 1.0f, 0.0f ) );
	axisA = rotate( dqA, axisA );
	float translation = dot( axisA, d );

	float a1 = cross( add( d, rA ), axisA );
	float a2 = cross( rB, axisA );

	Soft cs;
	cs.biasRate = base->constraintSoftness.biasRate;
	cs.massScale = base->constraintSoftness.massScale;
	cs.impulseScale = base->constraintSoftness.impulseScale;

	// motor
	if ( j->enableMotor && fixedRotation == false )
	{
		float Cdot = wB - wA - j->motorSpeed;
		float impulse = -j->motorMass * Cdot;
		float oldImpulse = j->motorImpulse;
		float maxImpulse = P.h * j->maxMotorTorque;
		float newImpulse = clampf_( oldImpulse + impulse, -maxImpulse, maxImpulse );
		j->motorImpulse = newImpulse;
		impulse = newImpulse - oldImpulse;

		wA -= iA * impulse;
		wB += iB * impulse;
	}

	// spring
	if ( j->enableSpring )
	{
		float C = translation;
		float bias = j->springSoftness.biasRate * C;
		float massScale = j->springSoftness.massScale;
		float impulseScale = j->springSoftness.impulseScale;

		float Cdot = dot( axisA, sub( vB, vA ) ) + a2 * wB - a1 * wA;
		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->springImpulse;
		j->springImpulse += impulse;

		V2 Pv = mulSV( impulse, axisA );
		float LA = impulse * a1;
		float LB = impulse * a2;

		vA = mulSub( vA, mA, Pv );
		wA -= iA * LA;
		vB = mulAdd( vB, mB, Pv );
		wB += iB * LB;
	}

	if ( j->enableLimit )
	{
		// lower limit
		{
			float C = translation - j->lowerTranslation;
			float bias = 0.0f;
			float massScale = 1.0f;
			float impulseScale = 0.0f;

			if ( C > 0.0f )
			{
				bias = C * P.inv_h;
			}
			else if ( useBias )
			{
				bias = cs.biasRate * C;
				massScale = cs.massScale;
				impulseScale = cs.impulseScale;
			}

			float Cdot = dot( axisA, sub( vB, vA ) ) + a2 * wB - a1 * wA;
			float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->lowerImpulse;
			float oldImpulse = j->lowerImpulse;
			float newImpulse = maxf_( oldImpulse + impulse, 0.0f );
			j->lowerImpulse = newImpulse;
			impulse = newImpulse - oldImpulse;

			V2 Pv = mulSV( impulse, axisA );
			float LA = impulse * a1;
			float LB = impulse * a2;

			vA = mulSub( vA, mA, Pv );
			wA -= iA * LA;
			vB = mulAdd( vB, mB, Pv );
			wB += iB * LB;
		}

		// upper limit, signs flipped
		{
			float C = j->upperTranslation - translation;
			float bias = 0.0f;
			float massScale = 1.0f;
			float impulseScale = 0.0f;

			if ( C > 0.0f )
			{
				bias = C * P.inv_h;
			}
			else if ( useBias )
			{
				bias = cs.biasRate * C;
				massScale = cs.massScale;
				impulseScale = cs.impulseScale;
			}

			float Cdot = dot( axisA, sub( vA, vB ) ) + a1 * wA - a2 * wB;
			float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->upperImpulse;
			float oldImpulse = j->upperImpulse;
			float newImpulse = maxf_( oldImpulse + impulse, 0.0f );
			j->upperImpulse = newImpulse;
			impulse = newImpulse - oldImpulse;

			V2 Pv = mulSV( impulse, axisA );
			float LA = impulse * a1;
			float LB = impulse * a2;

			vA = mulAdd( vA, mA, Pv );
			wA += iA * LA;
			vB = mulSub( vB, mB, Pv );
			wB -= iB * LB;
		}
	}

	// point to line
	{
		V2 perpA = leftPerp( axisA );

		float bias = 0.0f;
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( useBias )
		{
			float C = dot( perpA, d );
			bias = cs.biasRate * C;
			massScale = cs.massScale;
			impulseScale = cs.impulseScale;
		}

		float s1 = cross( add( d, rA ), perpA );
		float s2 = cross( rB, perpA );
		float Cdot = dot( perpA, sub( vB, vA ) ) + s2 * wB - s1 * wA;

		float impulse = -massScale * j->perpMass * ( Cdot + bias ) - impulseScale * j->perpImpulse;
		j->perpImpulse += impulse;

		V2 Pv = mulSV( impulse, perpA );
		float LA = impulse * s1;
		float LB = impulse * s2;

		vA = mulSub( vA, mA, Pv );
		wA -= iA * LA;
		vB = mulAdd( vB, mB, Pv );
		wB += iB * LB;
	}

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// ---- distance (src/distance_joint.c:315-542) -------------------------------------------------------------------
B2G_DEV void warmStartDistance( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lDistance* j = &base->u.distance;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 rA = rotate( deltaRot( jb.pA ), toV2( j->anchorA ) );
	V2 rB = rotate( deltaRot( jb.pB ), toV2( j->anchorB ) );

	V2 ds = add( sub( v2( jb.pB.x, jb.pB.y ), v2( jb.pA.x, jb.pA.y ) ), sub( rB, rA ) );
	V2 separation = add( toV2( j->deltaCenter ), ds );
	V2 axis = normalize( separation );

	float axialImpulse = j->impulse + j->lowerImpulse - j->upperImpulse + j->motorImpulse;
	V2 Pv = mulSV( axialImpulse, axis );

	applyWarmStart( V, jb, base->invMassA, base->invIA, base->invMassB, base->invIB, Pv, cross( rA, Pv ), cross( rB, Pv ) );
}

B2G_DEV void solveDistance( const StepParams& P, const SolveView& V, b2lJointSim* base, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB;
	float iA = base->invIA, iB = base->invIB;
	b2lDistance* j = &base->u.distance;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;

	V2 rA = rotate( deltaRot( jb.pA ), toV2( j->anchorA ) );
	V2 rB = rotate( deltaRot( jb.pB ), toV2( j->anchorB ) );

	V2 ds = add( sub( v2( jb.pB.x, jb.pB.y ), v2( jb.pA.x, jb.pA.y ) ), sub( rB, rA ) );
	V2 separation = add( toV2( j->deltaCenter ), ds );

	float len = length( separation );
	V2 axis = normalize( separation );

	Soft cs;
	cs.biasRate = base->constraintSoftness.biasRate;
	cs.massScale = base->constraintSoftness.massScale;
	cs.impulseScale = base->constraintSoftness.impulseScale;

	if ( j->enableSpring && ( j->minLength < j->maxLength || j->enableLimit == 0 ) )
	{
		// spring
		if ( j->hertz > 0.0f )
		{
			V2 vr = add( sub( vB, vA ), sub( crossSV( wB, rB ), crossSV( wA, rA ) ) );
			float cdot = dot( axis, vr );
			float c = len - j->length;
			float bias = j->distanceSoftness.biasRate * c;

			float m = j->distanceSoftness.massScale * j->axialMass;
			float oldImpulse = j->impulse;
			float impulse = -m * ( cdot + bias ) - j->distanceSoftness.impulseScale * oldImpulse;

			float h = P.h;
			float newImpulse = clampf_( oldImpulse + impulse, j->lowerSpringForce * h, j->upperSpringForce * h );
			j->impulse = newImpulse;
			impulse = newImpulse - oldImpulse;

			V2 Pv = mulSV( impulse, axis );
			vA = mulSub( vA, mA, Pv );
			wA -= iA * cross( rA, Pv );
			vB = mulAdd( vB, mB, Pv );
			wB += iB * cross( rB, Pv );
		}

		if ( j->enableMotor )
		{
			V2 vr = add( sub( vB, vA ), sub( crossSV( wB, rB ), crossSV( wA, rA ) ) );
			float Cdot = dot( axis, vr );
			float impulse = j->axialMass * ( j->motorSpeed - Cdot );
			float oldImpulse = j->motorImpulse;
			float maxImpulse = P.h * j->maxMotorForce;
			float newImpulse = clampf_( oldImpulse + impulse, -maxImpulse, maxImpulse );
			j->motorImpulse = newImpulse;
			impulse = newImpulse - oldImpulse;

			V2 Pv = mulSV( impulse, axis );
			vA = mulSub( vA, mA, Pv );
			wA -= iA * cross( rA, Pv );
			vB = mulAdd( vB, mB, Pv );
			wB += iB * cross( rB, Pv );
		}

		if ( j->enableLimit )
		{
			// lower limit
			{
				V2 vr = add( sub( vB, vA ), sub( crossSV( wB, rB ), crossSV( wA, rA ) ) );
				float cdot = dot( axis, vr );

				float c = len - j->minLength;

				float bias = 0.0f;
				float massScale = 1.0f;
				float impulseScale = 0.0f;
				if ( c > 0.0f )
				{
					bias = c * P.inv_h;
				}
				else if ( useBias )
				{
					bias = cs.biasRate * c;
					massScale = cs.massScale;
					impulseScale = cs.impulseScale;
				}

				float impulse = -massScale * j->axialMass * ( cdot + bias ) - impulseScale * j->lowerImpulse;
				float newImpulse = maxf_( 0.0f, j->lowerImpulse + impulse );
				impulse = newImpulse - j->lowerImpulse;
				j->lowerImpulse = newImpulse;

				V2 Pv = mulSV( impulse, axis );
				vA = mulSub( vA, mA, Pv );
				wA -= iA * cross( rA, Pv );
				vB = mulAdd( vB, mB, Pv );
				wB += iB * cross( rB, Pv );
			}

			// upper limit
			{
				V2 vr = add( sub( vA, vB ), sub( crossSV( wA, rA ), crossSV( wB, rB ) ) );
				float Cdot = dot( axis, vr );

				float C = j->maxLength - len;

				float bias = 0.0f;
				float massScale = 1.0f;
				float impulseScale = 0.0f;
				if ( C > 0.0f )
				{
					bias = C * P.inv_h;
				}
				else if ( useBias )
				{
					bias = cs.biasRate * C;
					massScale = cs.massScale;
					impulseScale = cs.impulseScale;
				}

				float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->upperImpulse;
				float newImpulse = maxf_( 0.0f, j->upperImpulse + impulse );
				impulse = newImpulse - j->upperImpulse;
				j->upperImpulse = newImpulse;

				V2 Pv = mulSV( -impulse, axis );
				vA = mulSub( vA, mA, Pv );
				wA -= iA * cross( rA, Pv );
				vB = mulAdd( vB, mB, Pv );
				wB += iB * cross( rB, Pv );
			}
		}
	}
	else
	{
		// rigid constraint
		V2 vr = add( sub( vB, vA ), sub( crossSV( wB, rB ), crossSV( wA, rA ) ) );
		float Cdot = dot( axis, vr );

		float C = len - j->length;

		float bias = 0.0f;
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if ( useBias )
		{
			bias = cs.biasRate * C;
			massScale = cs.massScale;
			impulseScale = cs.impulseScale;
		}

		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->impulse;
		j->impulse += impulse;

		V2 Pv = mulSV( impulse, axis );
		vA = mulSub( vA, mA, Pv );
		wA -= iA * cross( rA, Pv );
		vB = mulAdd( vB, mB, Pv );
		wB += iB * cross( rB, Pv );
	}

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// ---- motor (src/motor_joint.c:251-436) -------------------------------------------------------------------------
B2G_DEV void warmStartMotor( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lMotor* j = &base->u.motor;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 rA = rotate( deltaRot( jb.pA ), toV2( j->frameA.p ) );
	V2 rB = rotate( deltaRot( jb.pB ), toV2( j->frameB.p ) );

	V2 linearImpulse = add( toV2( j->linearVelocityImpulse ), toV2( j->linearSpringImpulse ) );
	float angularImpulse = j->angularVelocityImpulse + j->angularSpringImpulse;

	applyWarmStart( V, jb, base->invMassA, base->invIA, base->invMassB, base->invIB, linearImpulse,
					cross( rA, linearImpulse ) + angularImpulse, cross( rB, linearImpulse ) + angularImpulse );
}

B2G_DEV void solveMotor( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	float mA = base->invMassA, mB = base->invMassB;
	float iA = base->invIA, iB = base->invIB;
	b2lMotor* j = &base->u.motor;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	// angular spring
	if ( j->maxSpringTorque > 0.0f && j->angularHertz > 0.0f )
	{
		Rot qA = mulRot( dqA, toRot( j->frameA.q ) );
		Rot qB = mulRot( dqB, toRot( j->frameB.q ) );
		Rot relQ = invMulRot( qA, qB );

		float c = rotAngle( relQ );
		float bias = j->angularSpring.biasRate * c;
		float massScale = j->angularSpring.massScale;
		float impulseScale = j->angularSpring.impulseScale;

		float cdot = wB - wA;

		float maxImpulse = P.h * j->maxSpringTorque;
		float oldImpulse = j->angularSpringImpulse;
		float impulse = -massScale * j->angularMass * ( cdot + bias ) - impulseScale * oldImpulse;
		float newImpulse = clampf_( oldImpulse + impulse, -maxImpulse, maxImpulse );
		j->angularSpringImpulse = newImpulse;
		impulse = newImpulse - oldImpulse;

		wA -= iA * impulse;
		wB += iB * impulse;
	}

	// angular velocity
	if ( j->maxVelocityTorque > 0.0f )
	{
		float cdot = wB - wA - j->angularVelocity;
		float impulse = -j->angularMass * cdot;

		float maxImpulse = P.h * j->maxVelocityTorque;
		float oldImpulse = j->angularVelocityImpulse;
		float newImpulse = clampf_( oldImpulse + impulse, -maxImpulse, maxImpulse );
		j->angularVelocityImpulse = newImpulse;
		impulse = newImpulse - oldImpulse;

		wA -= iA * impulse;
		wB += iB * impulse;
	}

	V2 rA = rotate( dqA, toV2( j->frameA.p ) );
	V2 rB = rotate( dqB, toV2( j->frameB.p ) );

	// linear spring
	if ( j->maxSpringForce > 0.0f && j->linearHertz > 0.0f )
	{
		V2 dcA = v2( jb.pA.x, jb.pA.y );
		V2 dcB = v2( jb.pB.x, jb.pB.y );
		V2 c = add( add( sub( dcB, dcA ), sub( rB, rA ) ), toV2( j->deltaCenter ) );

		V2 bias = mulSV( j->linearSpring.biasRate, c );
		float massScale = j->linearSpring.massScale;
		float impulseScale = j->linearSpring.impulseScale;

		V2 cdot = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );
		cdot = add( cdot, bias );

		// kl.cx = (k11, k21), kl.cy = (k12, k22) with k12 == k21
		float k11 = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		float k21 = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		float k12 = k21;
		float k22 = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;

		// b2GetInverse22 (math_functions.h:730): a=cx.x b=cy.x c=cx.y d=cy.y
		{
			float a = k11, b = k12, cc = k21, d = k22;
			float det = a * d - b * cc;
			if ( det != 0.0f )
			{
				det = 1.0f / det;
			}
			j->linearMass.cx.x = det * d;
			j->linearMass.cx.y = -det * cc;
			j->linearMass.cy.x = -det * b;
			j->linearMass.cy.y = det * a;
		}

		// b2MulMV
		V2 b;
		b.x = j->linearMass.cx.x * cdot.x + j->linearMass.cy.x * cdot.y;
		b.y = j->linearMass.cx.y * cdot.x + j->linearMass.cy.y * cdot.y;

		V2 oldImpulse = toV2( j->linearSpringImpulse );
		V2 impulse;
		impulse.x = -massScale * b.x - impulseScale * oldImpulse.x;
		impulse.y = -massScale * b.y - impulseScale * oldImpulse.y;

		float maxImpulse = P.h * j->maxSpringForce;
		V2 accumulated = add( oldImpulse, impulse );

		if ( lengthSquared( accumulated ) > maxImpulse * maxImpulse )
		{
			accumulated = normalize( accumulated );
			accumulated.x *= maxImpulse;
			accumulated.y *= maxImpulse;
		}
		j->linearSpringImpulse.x = accumulated.x;
		j->linearSpringImpulse.y = accumulated.y;

		impulse = sub( accumulated, oldImpulse );

		vA = mulSub( vA, mA, impulse );
		wA -= iA * cross( rA, impulse );
		vB = mulAdd( vB, mB, impulse );
		wB += iB * cross( rB, impulse );
	}

	// linear velocity
	if ( j->maxVelocityForce > 0.0f )
	{
		V2 cdot = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );
		cdot = sub( cdot, toV2( j->linearVelocity ) );
		V2 b;
		b.x = j->linearMass.cx.x * cdot.x + j->linearMass.cy.x * cdot.y;
		b.y = j->linearMass.cx.y * cdot.x + j->linearMass.cy.y * cdot.y;
		V2 impulse = v2( -b.x, -b.y );

		V2 oldImpulse = toV2( j->linearVelocityImpulse );
		float maxImpulse = P.h * j->maxVelocityForce;
		V2 accumulated = add( oldImpulse, impulse );

		if ( lengthSquared( accumulated ) > maxImpulse * maxImpulse )
		{
			accumulated = normalize( accumulated );
			accumulated.x *= maxImpulse;
			accumulated.y *= maxImpulse;
		}
		j->linearVelocityImpulse.x = accumulated.x;
		j->linearVelocityImpulse.y = accumulated.y;

		impulse = sub( accumulated, oldImpulse );

		vA = mulSub( vA, mA, impulse );
		wA -= iA * cross( rA, impulse );
		vB = mulAdd( vB, mB, impulse );
		wB += iB * cross( rB, impulse );
	}

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// ---- mover (src/mover_joint.c:94-170) --------------------------------------------------------------------------
B2G_DEV void warmStartMover( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lMover* j = &base->u.mover;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );
	V2 L = toV2( j->linearVelocityImpulse );
	V2 vA = mulSub( v2( jb.vA.x, jb.vA.y ), base->invMassA, L );
	V2 vB = mulAdd( v2( jb.vB.x, jb.vB.y ), base->invMassB, L );
	scatterJointBodies( V, jb, vA, jb.vA.z, vB, jb.vB.z );
}

B2G_DEV void solveMover( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	float mA = base->invMassA, mB = base->invMassB;
	b2lMover* j = &base->u.mover;
	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	V2 vB = v2( jb.vB.x, jb.vB.y );

	if ( j->maxVelocityForce.x > 0.0f || j->maxVelocityForce.y > 0.0f )
	{
		V2 cdot = sub( vB, vA );
		cdot = sub( cdot, toV2( j->linearVelocity ) );
		V2 b = mulSV( j->linearMass, cdot );
		V2 impulse = v2( -b.x, -b.y );

		V2 oldImpulse = toV2( j->linearVelocityImpulse );
		V2 accumulated = add( oldImpulse, impulse );
		V2 maxImpulse = mulSV( P.h, toV2( j->maxVelocityForce ) );

		// b2Clamp( v, -max, max ) component-wise
		accumulated.x = clampf_( accumulated.x, -maxImpulse.x, maxImpulse.x );
		accumulated.y = clampf_( accumulated.y, -maxImpulse.y, maxImpulse.y );
		j->linearVelocityImpulse.x = accumulated.x;
		j->linearVelocityImpulse.y = accumulated.y;

		impulse = sub( accumulated, oldImpulse );
		vA = mulSub( vA, mA, impulse );
		vB = mulAdd( vB, mB, impulse );
	}
	else
	{
		j->linearVelocityImpulse.x = 0.0f;
		j->linearVelocityImpulse.y = 0.0f;
	}

	scatterJointBodies( V, jb, vA, jb.vA.z, vB, jb.vB.z );
}

// ---- pogo (src/pogo_joint.c:160-281) ---------------------------------------------------------------------------
B2G_DEV void warmStartPogo( const StepParams& P, const SolveView& V, b2lJointSim* base )
{
	b2lPogo* j = &base->u.pogo;
	if ( j->hertz == 0.0f )
	{
		j->impulse = 0.0f;
		return;
	}

	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );
	V2 rA = rotate( deltaRot( jb.pA ), toV2( j->frameA.p ) );
	V2 rB = rotate( deltaRot( jb.pB ), toV2( j->frameB.p ) );

	V2 linearImpulse = mulSV( j->impulse, toV2( j->normal ) );
	applyWarmStart( V, jb, base->invMassA, base->invIA, base->invMassB, base->invIB, linearImpulse, cross( rA, linearImpulse ),
					cross( rB, linearImpulse ) );
}

B2G_DEV void solvePogo( const StepParams& P, const SolveView& V, b2lJointSim* base, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB;
	float iA = base->invIA, iB = base->invIB;
	b2lPogo* j = &base->u.pogo;
	if ( j->hertz == 0.0f )
	{
		j->impulse = 0.0f;
		return;
	}

	JointBodies jb = gatherJointBodies( V, j->indexA, j->indexB );
	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;

	V2 rA = rotate( deltaRot( jb.pA ), toV2( j->frameA.p ) );
	V2 rB = rotate( deltaRot( jb.pB ), toV2( j->frameB.p ) );
	V2 normal = toV2( j->normal );

	float bias = 0.0f;
	if ( useBias )
	{
		V2 dcA = v2( jb.pA.x, jb.pA.y );
		V2 dcB = v2( jb.pB.x, jb.pB.y );
		V2 d = add( add( sub( dcB, dcA ), sub( rB, rA ) ), toV2( j->deltaCenter ) );

		V2 pogoAxis = rotate( toRot( j->frameB.q ), v2( 0.0f, 1.0f ) );
		float c = dot( pogoAxis, d ) - j->restLength;

		j->velocity = springDamper( j->hertz, j->dampingRatio, c, j->velocity, P.h );
		bias = -j->velocity;
	}

	V2 vr = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );
	float cdot = dot( normal, vr );

	float maxTensionImpulse = P.h * j->maxTensionForce;
	float maxCompressionImpulse = P.h * j->maxCompressionForce;
	float oldImpulse = j->impulse;
	float impulse = -j->linearMass * ( cdot + bias );
	float newImpulse = clampf_( oldImpulse + impulse, -maxTensionImpulse, maxCompressionImpulse );
	j->impulse = newImpulse;
	impulse = newImpulse - oldImpulse;

	V2 Pv = mulSV( impulse, normal );
	vA = mulSub( vA, mA, Pv );
	wA -= iA * cross( rA, Pv );
	vB = mulAdd( vB, mB, Pv );
	wB += iB * cross( rB, Pv );

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// indexA (indexB follows it) of the per-type block: rewritten to the island-local numbering by the island kernel
B2G_DEV int* jointIndexPair( b2lJointSim* joint )
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
			return nullptr; // filter joint: no solver data
	}
}

// What the device returns per joint: the fields the stages wrote (accumulated impulses, ...), B2L_JOINT_OUT_FLOATS floats.
// The reference solves joints in place in b2JointSim (no store stage, src/solver.h:73-74); the host copies these runs back.
B2G_DEV void storeJointImpulses( const StepParams& P, int jointIndex, const b2lJointSim* joint )
{
	int offsets[2], floats[2];
	int runs = b2lJointMutableRuns( joint->type, offsets, floats );
	float* out = P.outJoints + (size_t)jointIndex * B2L_JOINT_OUT_FLOATS;
	int n = 0;
	for ( int r = 0; r < runs; ++r )
	{
		const float* src = reinterpret_cast<const float*>( reinterpret_cast<const uint8_t*>( joint ) + offsets[r] );
		for ( int f = 0; f < floats[r]; ++f )
		{
			out[n++] = src[f];
		}
	}
}

// ---- plain revolute joints, compact ------------------------------------------------------------------------------
// A revolute joint without spring, motor and limit -- a hinge: chains, bridges, rag-doll-less rope, the joint_grid
// benchmark -- only ever runs the point-to-point part of b2SolveRevoluteJoint (src/revolute_joint.c:443-487) and the warm
// start (:283-315).  Of its 252-byte b2JointSim those two read 25 floats; when every joint of a step is such a hinge the
// cluster kernel keeps just those (31 words with what the event test and the output record need), so an island of 20 000
// joints fits the shared memory of one 16-block cluster instead of falling back to the grid-barrier kernel.  The
// arithmetic is the full path's, operation for operation (the disabled sub-constraints contribute nothing but constants).
enum LiteRevolute
{
	LR_INV_MASS_A = 0,
	LR_INV_MASS_B = 1,
	LR_INV_I_A = 2,
	LR_INV_I_B = 3,
	LR_BIAS_RATE = 4, // constraintSoftness
	LR_MASS_SCALE = 5,
	LR_IMPULSE_SCALE = 6,
	LR_JOINT_ID = 7,  // int
	LR_IMPULSE_X = 8, // linearImpulse (mutable)
	LR_IMPULSE_Y = 9,
	LR_INDEX_A = 10, // int, the view's numbering, -1 = static
	LR_INDEX_B = 11,
	LR_FRAME_A = 12, // p.x p.y q.c q.s
	LR_FRAME_B = 16,
	LR_DELTA_CENTER = 20,
	LR_FORCE_THRESHOLD = 22,
	LR_TORQUE_THRESHOLD = 23,
	// whatever the disabled sub-constraints accumulated while they were enabled still takes part in the warm start
	// (src/revolute_joint.c:296) and in the joint's reaction torque (src/joint.c:1040-1046); nothing changes it during the step
	LR_SPRING_IMPULSE = 24,
	LR_MOTOR_IMPULSE = 25,
	LR_LOWER_IMPULSE = 26,
	LR_UPPER_IMPULSE = 27,
	LR_BIT_BASE = 28, // int: the world's first bit in the joint-event bit set (the record's padding word)
	LR_RESERVED = 29  // .. 30: an odd number of words per record (b2g_types.cuh)
};
static_assert( LR_RESERVED + 2 == kLiteJointWords, "LiteRevolute layout" );

B2G_DEV bool isLiteRevolute( const b2lJointSim* joint )
{
	return joint->type == b2l_revoluteJoint && joint->u.revolute.enableSpring == 0 && joint->u.revolute.enableMotor == 0 &&
		   joint->u.revolute.enableLimit == 0;
}

B2G_DEV float* liteJointAt( const SolveView& V, int index )
{
	return reinterpret_cast<float*>( V.joints ) + (size_t)index * kLiteJointWords;
}

// word w of the 64-word joint record (kJointStride bytes) -> its place in a LiteRevolute, or -1
struct LiteRevoluteMap
{
	signed char liteOf[kJointStride / 4];
};

__host__ __device__ constexpr LiteRevoluteMap makeLiteRevoluteMap()
{
	LiteRevoluteMap m{};
	for ( int i = 0; i < kJointStride / 4; ++i )
	{
		m.liteOf[i] = -1;
	}
	m.liteOf[offsetof( b2lJointSim, invMassA ) / 4] = LR_INV_MASS_A;
	m.liteOf[offsetof( b2lJointSim, invMassB ) / 4] = LR_INV_MASS_B;
	m.liteOf[offsetof( b2lJointSim, invIA ) / 4] = LR_INV_I_A;
	m.liteOf[offsetof( b2lJointSim, invIB ) / 4] = LR_INV_I_B;
	m.liteOf[offsetof( b2lJointSim, constraintSoftness.biasRate ) / 4] = LR_BIAS_RATE;
	m.liteOf[offsetof( b2lJointSim, constraintSoftness.massScale ) / 4] = LR_MASS_SCALE;
	m.liteOf[offsetof( b2lJointSim, constraintSoftness.impulseScale ) / 4] = LR_IMPULSE_SCALE;
	m.liteOf[offsetof( b2lJointSim, jointId ) / 4] = LR_JOINT_ID;
	m.liteOf[offsetof( b2lJointSim, u.revolute.linearImpulse.x ) / 4] = LR_IMPULSE_X;
	m.liteOf[offsetof( b2lJointSim, u.revolute.linearImpulse.y ) / 4] = LR_IMPULSE_Y;
	m.liteOf[offsetof( b2lJointSim, u.revolute.indexA ) / 4] = LR_INDEX_A;
	m.liteOf[offsetof( b2lJointSim, u.revolute.indexB ) / 4] = LR_INDEX_B;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameA.p.x ) / 4] = LR_FRAME_A + 0;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameA.p.y ) / 4] = LR_FRAME_A + 1;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameA.q.c ) / 4] = LR_FRAME_A + 2;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameA.q.s ) / 4] = LR_FRAME_A + 3;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameB.p.x ) / 4] = LR_FRAME_B + 0;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameB.p.y ) / 4] = LR_FRAME_B + 1;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameB.q.c ) / 4] = LR_FRAME_B + 2;
	m.liteOf[offsetof( b2lJointSim, u.revolute.frameB.q.s ) / 4] = LR_FRAME_B + 3;
	m.liteOf[offsetof( b2lJointSim, u.revolute.deltaCenter.x ) / 4] = LR_DELTA_CENTER + 0;
	m.liteOf[offsetof( b2lJointSim, u.revolute.deltaCenter.y ) / 4] = LR_DELTA_CENTER + 1;
	m.liteOf[offsetof( b2lJointSim, forceThreshold ) / 4] = LR_FORCE_THRESHOLD;
	m.liteOf[offsetof( b2lJointSim, torqueThreshold ) / 4] = LR_TORQUE_THRESHOLD;
	m.liteOf[offsetof( b2lJointSim, u.revolute.springImpulse ) / 4] = LR_SPRING_IMPULSE;
	m.liteOf[offsetof( b2lJointSim, u.revolute.motorImpulse ) / 4] = LR_MOTOR_IMPULSE;
	m.liteOf[offsetof( b2lJointSim, u.revolute.lowerImpulse ) / 4] = LR_LOWER_IMPULSE;
	m.liteOf[offsetof( b2lJointSim, u.revolute.upperImpulse ) / 4] = LR_UPPER_IMPULSE;
	m.liteOf[B2L_JOINT_SIZE / 4] = LR_BIT_BASE;
	return m;
}

// The full records (global memory) -> the compact ones of a view, by 16 lanes per joint: lane q reads quad q of the 256-byte
// record -- one coalesced 256-byte request per joint instead of 25 scattered 4-byte loads per thread (measured on joint_grid:
// the scattered form took 41 of the step's 197 us) -- and puts the words of it that a LiteRevolute keeps in their places.
// slotOf( k ) = the view's slot of local joint k, recordOf( k ) = its index among the step's joints.  The body indices stay
// the step's (the caller renumbers them).
template <typename SlotOf, typename RecordOf>
B2G_DEV void loadLiteRevolutes( const StepParams& P, const SolveView& V, int jointCount, SlotOf slotOf, RecordOf recordOf )
{
	static constexpr LiteRevoluteMap map = makeLiteRevoluteMap();
	const int quad = (int)threadIdx.x & 15;
	for ( int k = (int)threadIdx.x >> 4; k < jointCount; k += (int)blockDim.x >> 4 )
	{
		const float4 q = reinterpret_cast<const float4*>( P.rawJoints + (size_t)recordOf( k ) * kJointStride )[quad];
		float* lite = liteJointAt( V, slotOf( k ) );
		const float words[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
		for ( int c = 0; c < 4; ++c )
		{
			const int place = map.liteOf[4 * quad + c];
			if ( place >= 0 )
			{
				lite[place] = words[c];
			}
		}
	}
}

// b2WarmStartRevoluteJoint, src/revolute_joint.c:283-315
B2G_DEV void warmStartRevoluteLite( const SolveView& V, const float* lite )
{
	JointBodies jb = gatherJointBodies( V, __float_as_int( lite[LR_INDEX_A] ), __float_as_int( lite[LR_INDEX_B] ) );
	V2 rA = rotate( deltaRot( jb.pA ), v2( lite[LR_FRAME_A + 0], lite[LR_FRAME_A + 1] ) );
	V2 rB = rotate( deltaRot( jb.pB ), v2( lite[LR_FRAME_B + 0], lite[LR_FRAME_B + 1] ) );
	float axialImpulse = lite[LR_SPRING_IMPULSE] + lite[LR_MOTOR_IMPULSE] + lite[LR_LOWER_IMPULSE] - lite[LR_UPPER_IMPULSE];
	V2 L = v2( lite[LR_IMPULSE_X], lite[LR_IMPULSE_Y] );
	applyWarmStart( V, jb, lite[LR_INV_MASS_A], lite[LR_INV_I_A], lite[LR_INV_MASS_B], lite[LR_INV_I_B], L, cross( rA, L ) + axialImpulse,
					cross( rB, L ) + axialImpulse );
}

// b2SolveRevoluteJoint with spring, motor and limit disabled: the point-to-point constraint, src/revolute_joint.c:443-499
B2G_DEV void solveRevoluteLite( const SolveView& V, float* lite, bool useBias )
{
	float mA = lite[LR_INV_MASS_A], mB = lite[LR_INV_MASS_B];
	float iA = lite[LR_INV_I_A], iB = lite[LR_INV_I_B];
	JointBodies jb = gatherJointBodies( V, __float_as_int( lite[LR_INDEX_A] ), __float_as_int( lite[LR_INDEX_B] ) );

	V2 vA = v2( jb.vA.x, jb.vA.y );
	float wA = jb.vA.z;
	V2 vB = v2( jb.vB.x, jb.vB.y );
	float wB = jb.vB.z;
	Rot dqA = deltaRot( jb.pA ), dqB = deltaRot( jb.pB );

	V2 rA = rotate( dqA, v2( lite[LR_FRAME_A + 0], lite[LR_FRAME_A + 1] ) );
	V2 rB = rotate( dqB, v2( lite[LR_FRAME_B + 0], lite[LR_FRAME_B + 1] ) );

	V2 Cdot = sub( add( vB, crossSV( wB, rB ) ), add( vA, crossSV( wA, rA ) ) );

	V2 bias = v2( 0.0f, 0.0f );
	float massScale = 1.0f;
	float impulseScale = 0.0f;
	if ( useBias )
	{
		V2 dcA = v2( jb.pA.x, jb.pA.y );
		V2 dcB = v2( jb.pB.x, jb.pB.y );
		V2 separation = add( add( sub( dcB, dcA ), sub( rB, rA ) ), v2( lite[LR_DELTA_CENTER + 0], lite[LR_DELTA_CENTER + 1] ) );
		bias = mulSV( lite[LR_BIAS_RATE], separation );
		massScale = lite[LR_MASS_SCALE];
		impulseScale = lite[LR_IMPULSE_SCALE];
	}

	float k11 = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
	float k12 = -rA.y * rA.x * iA - rB.y * rB.x * iB;
	float k21 = k12;
	float k22 = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
	V2 b = solve22( k11, k12, k21, k22, add( Cdot, bias ) );

	V2 impulse;
	impulse.x = -massScale * b.x - impulseScale * lite[LR_IMPULSE_X];
	impulse.y = -massScale * b.y - impulseScale * lite[LR_IMPULSE_Y];
	lite[LR_IMPULSE_X] += impulse.x;
	lite[LR_IMPULSE_Y] += impulse.y;

	vA = mulSub( vA, mA, impulse );
	wA -= iA * cross( rA, impulse );
	vB = mulAdd( vB, mB, impulse );
	wB += iB * cross( rB, impulse );

	scatterJointBodies( V, jb, vA, wA, vB, wB );
}

// jointEventTest for a LiteRevolute record
B2G_DEV void jointEventTestLite( const StepParams& P, const float* lite )
{
	float forceThreshold = lite[LR_FORCE_THRESHOLD], torqueThreshold = lite[LR_TORQUE_THRESHOLD];
	if ( !( forceThreshold < kHugeFloatMax || torqueThreshold < kHugeFloatMax ) )
	{
		return;
	}
	float linearImpulse = length( v2( lite[LR_IMPULSE_X], lite[LR_IMPULSE_Y] ) );
	float force = linearImpulse * P.inv_h;
	// b2GetJointReaction, src/joint.c:1040-1046
	float torque = absf_( lite[LR_MOTOR_IMPULSE] + lite[LR_LOWER_IMPULSE] - lite[LR_UPPER_IMPULSE] ) * P.inv_h;
	if ( force >= forceThreshold || torque >= torqueThreshold )
	{
		unsigned id = (unsigned)( __float_as_int( lite[LR_JOINT_ID] ) + __float_as_int( lite[LR_BIT_BASE] ) );
		atomicOr( P.jointBits + ( id >> 5 ), 1u << ( id & 31u ) );
	}
}

// the output record of a LiteRevolute joint: its linearImpulse, and the four accumulated impulses of the disabled
// sub-constraints exactly as they came in
B2G_DEV void storeJointImpulsesLite( const StepParams& P, int jointIndex, const float* lite )
{
	float* out = P.outJoints + (size_t)jointIndex * B2L_JOINT_OUT_FLOATS;
	out[0] = lite[LR_IMPULSE_X];
	out[1] = lite[LR_IMPULSE_Y];
	out[2] = lite[LR_SPRING_IMPULSE];
	out[3] = lite[LR_MOTOR_IMPULSE];
	out[4] = lite[LR_LOWER_IMPULSE];
	out[5] = lite[LR_UPPER_IMPULSE];
}

// ---- dispatch (src/joint.c:1454-1540) ----------------------------------------------------------------------------
B2G_DEV void warmStartJoint( const StepParams& P, const SolveView& V, b2lJointSim* joint )
{
	switch ( joint->type )
	{
		case b2l_distanceJoint:
			warmStartDistance( P, V, joint );
			break;
		case b2l_motorJoint:
			warmStartMotor( P, V, joint );
			break;
		case b2l_moverJoint:
			warmStartMover( P, V, joint );
			break;
		case b2l_pogoJoint:
			warmStartPogo( P, V, joint );
			break;
		case b2l_prismaticJoint:
			warmStartPrismatic( P, V, joint );
			break;
		case b2l_revoluteJoint:
			warmStartRevolute( P, V, joint );
			break;
		case b2l_weldJoint:
			warmStartWeld( P, V, joint );
			break;
		case b2l_wheelJoint:
			warmStartWheel( P, V, joint );
			break;
		default: // filter joint: nothing to solve
			break;
	}
}

B2G_DEV void solveJoint( const StepParams& P, const SolveView& V, b2lJointSim* joint, bool useBias )
{
	switch ( joint->type )
	{
		case b2l_distanceJoint:
			solveDistance( P, V, joint, useBias );
			break;
		case b2l_motorJoint:
			solveMotor( P, V, joint );
			break;
		case b2l_moverJoint:
			solveMover( P, V, joint );
			break;
		case b2l_pogoJoint:
			solvePogo( P, V, joint, useBias );
			break;
		case b2l_prismaticJoint:
			solvePrismatic( P, V, joint, useBias );
			break;
		case b2l_revoluteJoint:
			solveRevolute( P, V, joint, useBias );
			break;
		case b2l_weldJoint:
			solveWeld( P, V, joint, useBias );
			break;
		case b2l_wheelJoint:
			solveWheel( P, V, joint, useBias );
			break;
		default:
			break;
	}
}

// b2GetJointReaction (src/joint.c:993-1075) + the threshold test of b2SolveJointsTask (src/joint.c:1663-1675).
// Coloured joints only; overflow joints never raise events (src/joint.c:1576-1591).
B2G_DEV void jointEventTest( const StepParams& P, const b2lJointSim* sim )
{
	if ( !( sim->forceThreshold < kHugeFloatMax || sim->torqueThreshold < kHugeFloatMax ) )
	{
		return;
	}

	float linearImpulse = 0.0f;
	float angularImpulse = 0.0f;
	switch ( sim->type )
	{
		case b2l_distanceJoint:
		{
			const b2lDistance* j = &sim->u.distance;
			linearImpulse = absf_( j->impulse + j->lowerImpulse - j->upperImpulse + j->motorImpulse );
		}
		break;
		case b2l_motorJoint:
		{
			const b2lMotor* j = &sim->u.motor;
			linearImpulse = length( add( toV2( j->linearVelocityImpulse ), toV2( j->linearSpringImpulse ) ) );
			angularImpulse = absf_( j->angularVelocityImpulse + j->angularSpringImpulse );
		}
		break;
		case b2l_moverJoint:
			linearImpulse = length( toV2( sim->u.mover.linearVelocityImpulse ) );
			break;
		case b2l_pogoJoint:
			linearImpulse = absf_( sim->u.pogo.impulse );
			break;
		case b2l_prismaticJoint:
		{
			const b2lPrismatic* j = &sim->u.prismatic;
			float perpImpulse = j->impulse.x;
			float axialImpulse = j->motorImpulse + j->lowerImpulse - j->upperImpulse;
			linearImpulse = sqrtf( perpImpulse * perpImpulse + axialImpulse * axialImpulse );
			angularImpulse = absf_( j->impulse.y );
		}
		break;
		case b2l_revoluteJoint:
		{
			const b2lRevolute* j = &sim->u.revolute;
			linearImpulse = length( toV2( j->linearImpulse ) );
			angularImpulse = absf_( j->motorImpulse + j->lowerImpulse - j->upperImpulse );
		}
		break;
		case b2l_weldJoint:
			linearImpulse = length( toV2( sim->u.weld.linearImpulse ) );
			angularImpulse = absf_( sim->u.weld.angularImpulse );
			break;
		case b2l_wheelJoint:
		{
			const b2lWheel* j = &sim->u.wheel;
			float perpImpulse = j->perpImpulse;
			float axialImpulse = j->springImpulse + j->lowerImpulse - j->upperImpulse;
			linearImpulse = sqrtf( perpImpulse * perpImpulse + axialImpulse * axialImpulse );
			angularImpulse = absf_( j->motorImpulse );
		}
		break;
		default:
			break;
	}

	float force = linearImpulse * P.inv_h;
	float torque = angularImpulse * P.inv_h;
	if ( force >= sim->forceThreshold || torque >= sim->torqueThreshold )
	{
		// the 4 padding bytes behind the 252-byte record carry the world's first bit in the joint-event set (0 for a
		// single world, see b2GpuSolverPackRange)
		int bitBase = *reinterpret_cast<const int*>( reinterpret_cast<const uint8_t*>( sim ) + B2L_JOINT_SIZE );
		unsigned id = (unsigned)( sim->jointId + bitBase );
		atomicOr( P.jointBits + ( id >> 5 ), 1u << ( id & 31u ) );
	}
}

} // namespace b2g
