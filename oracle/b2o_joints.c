/*
 * b2o_joints.c -- CPU oracle: joint warm-start / solve / relax for all joint types (TEST INFRASTRUCTURE).
 *
 * A plain-C, one-joint-at-a-time restatement of the reference's joint solvers working on the reference's own
 * b2JointSim records (mirrored by b2lJointSim, include/b2gpu_layout.h) and b2BodyState array.  Each function cites
 * the reference function it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use this code; the product never does.
 */
#include "b2o_solver.h"

static const o_state o_identity_state = { { 0.0f, 0.0f }, 0.0f, 0u, { 0.0f, 0.0f }, { 1.0f, 0.0f } };

static inline o_vec2 jv( b2lVec2 v )
{
	return o_v( v.x, v.y );
}

static inline o_rot jr( b2lRot q )
{
	o_rot r = { q.c, q.s };
	return r;
}

static inline o_soft js( b2lSoft s )
{
	o_soft r = { s.biasRate, s.massScale, s.impulseScale };
	return r;
}

static inline o_state* state_of( const o_ctx* ctx, int index, o_state* dummy )
{
	return index == -1 ? dummy : ctx->states + index;
}

static inline void write_back( o_state* sA, o_vec2 vA, float wA, o_state* sB, o_vec2 vB, float wB )
{
	if ( sA->flags & B2L_FLAG_DYNAMIC )
	{
		sA->v = vA;
		sA->w = wA;
	}
	if ( sB->flags & B2L_FLAG_DYNAMIC )
	{
		sB->v = vB;
		sB->w = wB;
	}
}

/* common warm start: apply linear impulse P and angular impulses LA/LB under the dynamic flag */
static inline void warm_apply( o_state* sA, o_state* sB, float mA, float iA, float mB, float iB, o_vec2 P, float LA, float LB )
{
	if ( sA->flags & B2L_FLAG_DYNAMIC )
	{
		sA->v = o_mul_sub( sA->v, mA, P );
		sA->w -= iA * LA;
	}
	if ( sB->flags & B2L_FLAG_DYNAMIC )
	{
		sB->v = o_mul_add( sB->v, mB, P );
		sB->w += iB * LB;
	}
}

/* ---- revolute: src/revolute_joint.c:283 (warm start), :317 (solve) ------------------------------------------------ */
static void warm_revolute( b2lJointSim* base, const o_ctx* ctx )
{
	b2lRevolute* j = &base->u.revolute;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );
	float axialImpulse = j->springImpulse + j->motorImpulse + j->lowerImpulse - j->upperImpulse;
	o_vec2 L = jv( j->linearImpulse );
	warm_apply( sA, sB, base->invMassA, base->invIA, base->invMassB, base->invIB, L, o_cross( rA, L ) + axialImpulse,
				o_cross( rB, L ) + axialImpulse );
}

static void solve_revolute( b2lJointSim* base, const o_ctx* ctx, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB, iA = base->invIA, iB = base->invIB;
	b2lRevolute* j = &base->u.revolute;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;

	o_rot qA = o_mul_rot( sA->dq, jr( j->frameA.q ) );
	o_rot qB = o_mul_rot( sB->dq, jr( j->frameB.q ) );
	o_rot relQ = o_inv_mul_rot( qA, qB );
	bool fixedRotation = ( iA + iB == 0.0f );
	o_soft cs = js( base->constraintSoftness );

	if ( j->enableSpring && fixedRotation == false )
	{
		float C = o_unwind_angle( o_rot_angle( relQ ) - j->targetAngle );
		float bias = j->springSoftness.biasRate * C;
		float massScale = j->springSoftness.massScale;
		float impulseScale = j->springSoftness.impulseScale;
		float Cdot = wB - wA;
		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->springImpulse;
		j->springImpulse += impulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if ( j->enableMotor && fixedRotation == false )
	{
		float Cdot = wB - wA - j->motorSpeed;
		float impulse = -j->axialMass * Cdot;
		float oldImpulse = j->motorImpulse;
		float maxImpulse = ctx->h * j->maxMotorTorque;
		j->motorImpulse = o_clamp( oldImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j->motorImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if ( j->enableLimit && fixedRotation == false )
	{
		float jointAngle = o_rot_angle( relQ );
		{
			float C = jointAngle - j->lowerAngle;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if ( C > 0.0f )
			{
				bias = C * ctx->inv_h;
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
			j->lowerImpulse = o_max( oldImpulse + impulse, 0.0f );
			impulse = j->lowerImpulse - oldImpulse;
			wA -= iA * impulse;
			wB += iB * impulse;
		}
		{
			float C = j->upperAngle - jointAngle;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if ( C > 0.0f )
			{
				bias = C * ctx->inv_h;
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
			j->upperImpulse = o_max( oldImpulse + impulse, 0.0f );
			impulse = j->upperImpulse - oldImpulse;
			wA += iA * impulse;
			wB -= iB * impulse;
		}
	}

	{
		o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
		o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );
		o_vec2 Cdot = o_sub( o_add( vB, o_cross_sv( wB, rB ) ), o_add( vA, o_cross_sv( wA, rA ) ) );
		o_vec2 bias = o_v( 0.0f, 0.0f );
		float massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			o_vec2 separation = o_add( o_add( o_sub( sB->dp, sA->dp ), o_sub( rB, rA ) ), jv( j->deltaCenter ) );
			bias = o_mul_sv( cs.biasRate, separation );
			massScale = cs.massScale;
			impulseScale = cs.impulseScale;
		}
		float k11 = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		float k12 = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		float k22 = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		o_vec2 b = o_solve22( k11, k12, k12, k22, o_add( Cdot, bias ) );
		o_vec2 impulse;
		impulse.x = -massScale * b.x - impulseScale * j->linearImpulse.x;
		impulse.y = -massScale * b.y - impulseScale * j->linearImpulse.y;
		j->linearImpulse.x += impulse.x;
		j->linearImpulse.y += impulse.y;
		vA = o_mul_sub( vA, mA, impulse );
		wA -= iA * o_cross( rA, impulse );
		vB = o_mul_add( vB, mB, impulse );
		wB += iB * o_cross( rB, impulse );
	}
	write_back( sA, vA, wA, sB, vB, wB );
}

/* ---- weld: src/weld_joint.c:222, :253 (B2_WELD_BLOCK_SOLVE 0) ------------------------------------------------------ */
static void warm_weld( b2lJointSim* base, const o_ctx* ctx )
{
	b2lWeld* j = &base->u.weld;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );
	o_vec2 L = jv( j->linearImpulse );
	warm_apply( sA, sB, base->invMassA, base->invIA, base->invMassB, base->invIB, L, o_cross( rA, L ) + j->angularImpulse,
				o_cross( rB, L ) + j->angularImpulse );
}

static void solve_weld( b2lJointSim* base, const o_ctx* ctx, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB, iA = base->invIA, iB = base->invIB;
	b2lWeld* j = &base->u.weld;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;

	{
		o_rot qA = o_mul_rot( sA->dq, jr( j->frameA.q ) );
		o_rot qB = o_mul_rot( sB->dq, jr( j->frameB.q ) );
		float jointAngle = o_rot_angle( o_inv_mul_rot( qA, qB ) );
		float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias || j->angularHertz > 0.0f )
		{
			bias = j->angularSpring.biasRate * jointAngle;
			massScale = j->angularSpring.massScale;
			impulseScale = j->angularSpring.impulseScale;
		}
		float Cdot = wB - wA;
		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->angularImpulse;
		j->angularImpulse += impulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}
	{
		o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
		o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );
		o_vec2 bias = o_v( 0.0f, 0.0f );
		float massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias || j->linearHertz > 0.0f )
		{
			o_vec2 C = o_add( o_add( o_sub( sB->dp, sA->dp ), o_sub( rB, rA ) ), jv( j->deltaCenter ) );
			bias = o_mul_sv( j->linearSpring.biasRate, C );
			massScale = j->linearSpring.massScale;
			impulseScale = j->linearSpring.impulseScale;
		}
		o_vec2 Cdot = o_sub( o_add( vB, o_cross_sv( wB, rB ) ), o_add( vA, o_cross_sv( wA, rA ) ) );
		float k11 = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		float k12 = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		float k22 = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		o_vec2 b = o_solve22( k11, k12, k12, k22, o_add( Cdot, bias ) );
		o_vec2 impulse = { -massScale * b.x - impulseScale * j->linearImpulse.x,
						   -massScale * b.y - impulseScale * j->linearImpulse.y };
		j->linearImpulse.x = j->linearImpulse.x + impulse.x;
		j->linearImpulse.y = j->linearImpulse.y + impulse.y;
		vA = o_mul_sub( vA, mA, impulse );
		wA -= iA * o_cross( rA, impulse );
		vB = o_mul_add( vB, mB, impulse );
		wB += iB * o_cross( rB, impulse );
	}
	write_back( sA, vA, wA, sB, vB, wB );
}

/* ---- prismatic: src/prismatic_joint.c:353, :407 -------------------------------------------------------------------- */
typedef struct
{
	o_vec2 rA, rB, d, axisA;
	float a1, a2;
} o_slider;

static o_slider slider_frame( const o_state* sA, const o_state* sB, b2lTransform frameA, b2lTransform frameB, b2lVec2 deltaCenter,
							  bool dFirst )
{
	o_slider f;
	f.rA = o_rotate( sA->dq, jv( frameA.p ) );
	f.rB = o_rotate( sB->dq, jv( frameB.p ) );
	f.d = o_add( o_add( o_sub( sB->dp, sA->dp ), jv( deltaCenter ) ), o_sub( f.rB, f.rA ) );
	f.axisA = o_rotate( jr( frameA.q ), o_v( 1.0f, 0.0f ) );
	f.axisA = o_rotate( sA->dq, f.axisA );
	/* prismatic uses cross(rA + d, axis), wheel uses cross(d + rA, axis): the sums commute bit-exactly */
	(void)dFirst;
	f.a1 = o_cross( o_add( f.rA, f.d ), f.axisA );
	f.a2 = o_cross( f.rB, f.axisA );
	return f;
}

static void warm_prismatic( b2lJointSim* base, const o_ctx* ctx )
{
	b2lPrismatic* j = &base->u.prismatic;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_slider f = slider_frame( sA, sB, j->frameA, j->frameB, j->deltaCenter, false );
	float axialImpulse = j->springImpulse + j->motorImpulse + j->lowerImpulse - j->upperImpulse;
	o_vec2 perpA = o_left_perp( f.axisA );
	float s1 = o_cross( o_add( f.rA, f.d ), perpA );
	float s2 = o_cross( f.rB, perpA );
	float perpImpulse = j->impulse.x;
	float angleImpulse = j->impulse.y;
	o_vec2 P = o_add( o_mul_sv( axialImpulse, f.axisA ), o_mul_sv( perpImpulse, perpA ) );
	float LA = axialImpulse * f.a1 + perpImpulse * s1 + angleImpulse;
	float LB = axialImpulse * f.a2 + perpImpulse * s2 + angleImpulse;
	warm_apply( sA, sB, base->invMassA, base->invIA, base->invMassB, base->invIB, P, LA, LB );
}

/* one axial row shared by the spring/limit rows of the prismatic and wheel joints:
 * v -= m P, w -= i L with P = impulse * axis, LA = impulse * a1, LB = impulse * a2 (sign +1) or flipped (sign -1) */
static inline void axial_apply( o_vec2* vA, float* wA, o_vec2* vB, float* wB, float mA, float iA, float mB, float iB, o_vec2 axis,
								float a1, float a2, float impulse, bool flipped )
{
	o_vec2 P = o_mul_sv( impulse, axis );
	float LA = impulse * a1;
	float LB = impulse * a2;
	if ( flipped == false )
	{
		*vA = o_mul_sub( *vA, mA, P );
		*wA -= iA * LA;
		*vB = o_mul_add( *vB, mB, P );
		*wB += iB * LB;
	}
	else
	{
		*vA = o_mul_add( *vA, mA, P );
		*wA += iA * LA;
		*vB = o_mul_sub( *vB, mB, P );
		*wB -= iB * LB;
	}
}

static void solve_prismatic( b2lJointSim* base, const o_ctx* ctx, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB, iA = base->invIA, iB = base->invIB;
	b2lPrismatic* j = &base->u.prismatic;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;

	o_rot qA = o_mul_rot( sA->dq, jr( j->frameA.q ) );
	o_rot qB = o_mul_rot( sB->dq, jr( j->frameB.q ) );
	o_rot relQ = o_inv_mul_rot( qA, qB );

	o_slider f = slider_frame( sA, sB, j->frameA, j->frameB, j->deltaCenter, false );
	float translation = o_dot( f.axisA, f.d );
	float a1 = f.a1, a2 = f.a2;
	o_vec2 axisA = f.axisA;

	float k = mA + mB + iA * a1 * a1 + iB * a2 * a2;
	float axialMass = k > 0.0f ? 1.0f / k : 0.0f;
	o_soft softness = js( base->constraintSoftness );

	if ( j->enableSpring )
	{
		float C = translation - j->targetTranslation;
		float bias = j->springSoftness.biasRate * C;
		float massScale = j->springSoftness.massScale;
		float impulseScale = j->springSoftness.impulseScale;
		float Cdot = o_dot( axisA, o_sub( vB, vA ) ) + a2 * wB - a1 * wA;
		float deltaImpulse = -massScale * axialMass * ( Cdot + bias ) - impulseScale * j->springImpulse;
		j->springImpulse += deltaImpulse;
		axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, axisA, a1, a2, deltaImpulse, false );
	}

	if ( j->enableMotor )
	{
		float Cdot = o_dot( axisA, o_sub( vB, vA ) ) + a2 * wB - a1 * wA;
		float impulse = axialMass * ( j->motorSpeed - Cdot );
		float oldImpulse = j->motorImpulse;
		float maxImpulse = ctx->h * j->maxMotorForce;
		j->motorImpulse = o_clamp( oldImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j->motorImpulse - oldImpulse;
		axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, axisA, a1, a2, impulse, false );
	}

	if ( j->enableLimit )
	{
		float speculativeDistance = 0.25f * ( j->upperTranslation - j->lowerTranslation );
		{
			float C = translation - j->lowerTranslation;
			if ( C < speculativeDistance )
			{
				float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
				if ( C > 0.0f )
				{
					bias = o_min( C, ctx->lengthUnitsPerMeter ) * ctx->inv_h;
				}
				else if ( useBias )
				{
					bias = softness.biasRate * C;
					massScale = softness.massScale;
					impulseScale = softness.impulseScale;
				}
				float oldImpulse = j->lowerImpulse;
				float Cdot = o_dot( axisA, o_sub( vB, vA ) ) + a2 * wB - a1 * wA;
				float deltaImpulse = -axialMass * massScale * ( Cdot + bias ) - impulseScale * oldImpulse;
				j->lowerImpulse = o_max( oldImpulse + deltaImpulse, 0.0f );
				deltaImpulse = j->lowerImpulse - oldImpulse;
				axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, axisA, a1, a2, deltaImpulse, false );
			}
			else
			{
				j->lowerImpulse = 0.0f;
			}
		}
		{
			float C = j->upperTranslation - translation;
			if ( C < speculativeDistance )
			{
				float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
				if ( C > 0.0f )
				{
					bias = o_min( C, ctx->lengthUnitsPerMeter ) * ctx->inv_h;
				}
				else if ( useBias )
				{
					bias = softness.biasRate * C;
					massScale = softness.massScale;
					impulseScale = softness.impulseScale;
				}
				float oldImpulse = j->upperImpulse;
				float Cdot = o_dot( axisA, o_sub( vA, vB ) ) + a1 * wA - a2 * wB;
				float deltaImpulse = -axialMass * massScale * ( Cdot + bias ) - impulseScale * oldImpulse;
				j->upperImpulse = o_max( oldImpulse + deltaImpulse, 0.0f );
				deltaImpulse = j->upperImpulse - oldImpulse;
				axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, axisA, a1, a2, deltaImpulse, true );
			}
			else
			{
				j->upperImpulse = 0.0f;
			}
		}
	}

	{
		o_vec2 perpA = o_left_perp( axisA );
		float s1 = o_cross( o_add( f.d, f.rA ), perpA );
		float s2 = o_cross( f.rB, perpA );
		o_vec2 Cdot;
		Cdot.x = o_dot( perpA, o_sub( vB, vA ) ) + s2 * wB - s1 * wA;
		Cdot.y = wB - wA;
		o_vec2 bias = o_v( 0.0f, 0.0f );
		float massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			o_vec2 C;
			C.x = o_dot( perpA, f.d );
			C.y = o_rot_angle( relQ );
			bias = o_mul_sv( softness.biasRate, C );
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
		o_vec2 b = o_solve22( k11, k12, k12, k22, o_add( Cdot, bias ) );
		o_vec2 deltaImpulse;
		deltaImpulse.x = -massScale * b.x - impulseScale * j->impulse.x;
		deltaImpulse.y = -massScale * b.y - impulseScale * j->impulse.y;
		j->impulse.x += deltaImpulse.x;
		j->impulse.y += deltaImpulse.y;
		o_vec2 P = o_mul_sv( deltaImpulse.x, perpA );
		float LA = deltaImpulse.x * s1 + deltaImpulse.y;
		float LB = deltaImpulse.x * s2 + deltaImpulse.y;
		vA = o_mul_sub( vA, mA, P );
		wA -= iA * LA;
		vB = o_mul_add( vB, mB, P );
		wB += iB * LB;
	}
	write_back( sA, vA, wA, sB, vB, wB );
}

/* ---- wheel: src/wheel_joint.c:280, :329 ----------------------------------------------------------------------------- */
static void warm_wheel( b2lJointSim* base, const o_ctx* ctx )
{
	b2lWheel* j = &base->u.wheel;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_slider f = slider_frame( sA, sB, j->frameA, j->frameB, j->deltaCenter, true );
	o_vec2 perpA = o_left_perp( f.axisA );
	float s1 = o_cross( o_add( f.d, f.rA ), perpA );
	float s2 = o_cross( f.rB, perpA );
	float axialImpulse = j->springImpulse + j->lowerImpulse - j->upperImpulse;
	o_vec2 P = o_add( o_mul_sv( axialImpulse, f.axisA ), o_mul_sv( j->perpImpulse, perpA ) );
	float LA = axialImpulse * f.a1 + j->perpImpulse * s1 + j->motorImpulse;
	float LB = axialImpulse * f.a2 + j->perpImpulse * s2 + j->motorImpulse;
	warm_apply( sA, sB, base->invMassA, base->invIA, base->invMassB, base->invIB, P, LA, LB );
}

static void solve_wheel( b2lJointSim* base, const o_ctx* ctx, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB, iA = base->invIA, iB = base->invIB;
	b2lWheel* j = &base->u.wheel;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;
	bool fixedRotation = ( iA + iB == 0.0f );

	o_slider f = slider_frame( sA, sB, j->frameA, j->frameB, j->deltaCenter, true );
	o_vec2 axisA = f.axisA;
	float translation = o_dot( axisA, f.d );
	float a1 = f.a1, a2 = f.a2;
	o_soft cs = js( base->constraintSoftness );

	if ( j->enableMotor && fixedRotation == false )
	{
		float Cdot = wB - wA - j->motorSpeed;
		float impulse = -j->motorMass * Cdot;
		float oldImpulse = j->motorImpulse;
		float maxImpulse = ctx->h * j->maxMotorTorque;
		j->motorImpulse = o_clamp( oldImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j->motorImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if ( j->enableSpring )
	{
		float C = translation;
		float bias = j->springSoftness.biasRate * C;
		float massScale = j->springSoftness.massScale;
		float impulseScale = j->springSoftness.impulseScale;
		float Cdot = o_dot( axisA, o_sub( vB, vA ) ) + a2 * wB - a1 * wA;
		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->springImpulse;
		j->springImpulse += impulse;
		axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, axisA, a1, a2, impulse, false );
	}

	if ( j->enableLimit )
	{
		{
			float C = translation - j->lowerTranslation;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if ( C > 0.0f )
			{
				bias = C * ctx->inv_h;
			}
			else if ( useBias )
			{
				bias = cs.biasRate * C;
				massScale = cs.massScale;
				impulseScale = cs.impulseScale;
			}
			float Cdot = o_dot( axisA, o_sub( vB, vA ) ) + a2 * wB - a1 * wA;
			float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->lowerImpulse;
			float oldImpulse = j->lowerImpulse;
			j->lowerImpulse = o_max( oldImpulse + impulse, 0.0f );
			impulse = j->lowerImpulse - oldImpulse;
			axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, axisA, a1, a2, impulse, false );
		}
		{
			float C = j->upperTranslation - translation;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if ( C > 0.0f )
			{
				bias = C * ctx->inv_h;
			}
			else if ( useBias )
			{
				bias = cs.biasRate * C;
				massScale = cs.massScale;
				impulseScale = cs.impulseScale;
			}
			float Cdot = o_dot( axisA, o_sub( vA, vB ) ) + a1 * wA - a2 * wB;
			float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->upperImpulse;
			float oldImpulse = j->upperImpulse;
			j->upperImpulse = o_max( oldImpulse + impulse, 0.0f );
			impulse = j->upperImpulse - oldImpulse;
			axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, axisA, a1, a2, impulse, true );
		}
	}

	{
		o_vec2 perpA = o_left_perp( axisA );
		float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			float C = o_dot( perpA, f.d );
			bias = cs.biasRate * C;
			massScale = cs.massScale;
			impulseScale = cs.impulseScale;
		}
		float s1 = o_cross( o_add( f.d, f.rA ), perpA );
		float s2 = o_cross( f.rB, perpA );
		float Cdot = o_dot( perpA, o_sub( vB, vA ) ) + s2 * wB - s1 * wA;
		float impulse = -massScale * j->perpMass * ( Cdot + bias ) - impulseScale * j->perpImpulse;
		j->perpImpulse += impulse;
		axial_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, perpA, s1, s2, impulse, false );
	}
	write_back( sA, vA, wA, sB, vB, wB );
}

/* ---- distance: src/distance_joint.c:315, :354 ------------------------------------------------------------------------ */
static inline void point_apply( o_vec2* vA, float* wA, o_vec2* vB, float* wB, float mA, float iA, float mB, float iB, o_vec2 rA,
								o_vec2 rB, o_vec2 P )
{
	*vA = o_mul_sub( *vA, mA, P );
	*wA -= iA * o_cross( rA, P );
	*vB = o_mul_add( *vB, mB, P );
	*wB += iB * o_cross( rB, P );
}

static void warm_distance( b2lJointSim* base, const o_ctx* ctx )
{
	b2lDistance* j = &base->u.distance;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 rA = o_rotate( sA->dq, jv( j->anchorA ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->anchorB ) );
	o_vec2 ds = o_add( o_sub( sB->dp, sA->dp ), o_sub( rB, rA ) );
	o_vec2 axis = o_normalize( o_add( jv( j->deltaCenter ), ds ) );
	float axialImpulse = j->impulse + j->lowerImpulse - j->upperImpulse + j->motorImpulse;
	o_vec2 P = o_mul_sv( axialImpulse, axis );
	warm_apply( sA, sB, base->invMassA, base->invIA, base->invMassB, base->invIB, P, o_cross( rA, P ), o_cross( rB, P ) );
}

static void solve_distance( b2lJointSim* base, const o_ctx* ctx, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB, iA = base->invIA, iB = base->invIB;
	b2lDistance* j = &base->u.distance;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;

	o_vec2 rA = o_rotate( sA->dq, jv( j->anchorA ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->anchorB ) );
	o_vec2 ds = o_add( o_sub( sB->dp, sA->dp ), o_sub( rB, rA ) );
	o_vec2 separation = o_add( jv( j->deltaCenter ), ds );
	float length = o_length( separation );
	o_vec2 axis = o_normalize( separation );
	o_soft cs = js( base->constraintSoftness );

	if ( j->enableSpring && ( j->minLength < j->maxLength || j->enableLimit == 0 ) )
	{
		if ( j->hertz > 0.0f )
		{
			o_vec2 vr = o_add( o_sub( vB, vA ), o_sub( o_cross_sv( wB, rB ), o_cross_sv( wA, rA ) ) );
			float cdot = o_dot( axis, vr );
			float c = length - j->length;
			float bias = j->distanceSoftness.biasRate * c;
			float m = j->distanceSoftness.massScale * j->axialMass;
			float oldImpulse = j->impulse;
			float impulse = -m * ( cdot + bias ) - j->distanceSoftness.impulseScale * oldImpulse;
			float h = ctx->h;
			j->impulse = o_clamp( oldImpulse + impulse, j->lowerSpringForce * h, j->upperSpringForce * h );
			impulse = j->impulse - oldImpulse;
			point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, o_mul_sv( impulse, axis ) );
		}
		if ( j->enableMotor )
		{
			o_vec2 vr = o_add( o_sub( vB, vA ), o_sub( o_cross_sv( wB, rB ), o_cross_sv( wA, rA ) ) );
			float Cdot = o_dot( axis, vr );
			float impulse = j->axialMass * ( j->motorSpeed - Cdot );
			float oldImpulse = j->motorImpulse;
			float maxImpulse = ctx->h * j->maxMotorForce;
			j->motorImpulse = o_clamp( oldImpulse + impulse, -maxImpulse, maxImpulse );
			impulse = j->motorImpulse - oldImpulse;
			point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, o_mul_sv( impulse, axis ) );
		}
		if ( j->enableLimit )
		{
			{
				o_vec2 vr = o_add( o_sub( vB, vA ), o_sub( o_cross_sv( wB, rB ), o_cross_sv( wA, rA ) ) );
				float cdot = o_dot( axis, vr );
				float c = length - j->minLength;
				float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
				if ( c > 0.0f )
				{
					bias = c * ctx->inv_h;
				}
				else if ( useBias )
				{
					bias = cs.biasRate * c;
					massScale = cs.massScale;
					impulseScale = cs.impulseScale;
				}
				float impulse = -massScale * j->axialMass * ( cdot + bias ) - impulseScale * j->lowerImpulse;
				float newImpulse = o_max( 0.0f, j->lowerImpulse + impulse );
				impulse = newImpulse - j->lowerImpulse;
				j->lowerImpulse = newImpulse;
				point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, o_mul_sv( impulse, axis ) );
			}
			{
				o_vec2 vr = o_add( o_sub( vA, vB ), o_sub( o_cross_sv( wA, rA ), o_cross_sv( wB, rB ) ) );
				float Cdot = o_dot( axis, vr );
				float C = j->maxLength - length;
				float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
				if ( C > 0.0f )
				{
					bias = C * ctx->inv_h;
				}
				else if ( useBias )
				{
					bias = cs.biasRate * C;
					massScale = cs.massScale;
					impulseScale = cs.impulseScale;
				}
				float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->upperImpulse;
				float newImpulse = o_max( 0.0f, j->upperImpulse + impulse );
				impulse = newImpulse - j->upperImpulse;
				j->upperImpulse = newImpulse;
				point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, o_mul_sv( -impulse, axis ) );
			}
		}
	}
	else
	{
		o_vec2 vr = o_add( o_sub( vB, vA ), o_sub( o_cross_sv( wB, rB ), o_cross_sv( wA, rA ) ) );
		float Cdot = o_dot( axis, vr );
		float C = length - j->length;
		float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if ( useBias )
		{
			bias = cs.biasRate * C;
			massScale = cs.massScale;
			impulseScale = cs.impulseScale;
		}
		float impulse = -massScale * j->axialMass * ( Cdot + bias ) - impulseScale * j->impulse;
		j->impulse += impulse;
		point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, o_mul_sv( impulse, axis ) );
	}
	write_back( sA, vA, wA, sB, vB, wB );
}

/* ---- motor: src/motor_joint.c:251, :287 -------------------------------------------------------------------------------- */
static void warm_motor( b2lJointSim* base, const o_ctx* ctx )
{
	b2lMotor* j = &base->u.motor;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );
	o_vec2 linearImpulse = o_add( jv( j->linearVelocityImpulse ), jv( j->linearSpringImpulse ) );
	float angularImpulse = j->angularVelocityImpulse + j->angularSpringImpulse;
	warm_apply( sA, sB, base->invMassA, base->invIA, base->invMassB, base->invIB, linearImpulse,
				o_cross( rA, linearImpulse ) + angularImpulse, o_cross( rB, linearImpulse ) + angularImpulse );
}

static inline o_vec2 mat_mul( b2lMat22 A, o_vec2 v ) /* b2MulMV math_functions.h:720 */
{
	return o_v( A.cx.x * v.x + A.cy.x * v.y, A.cx.y * v.x + A.cy.y * v.y );
}

static void solve_motor( b2lJointSim* base, const o_ctx* ctx )
{
	float mA = base->invMassA, mB = base->invMassB, iA = base->invIA, iB = base->invIB;
	b2lMotor* j = &base->u.motor;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;

	if ( j->maxSpringTorque > 0.0f && j->angularHertz > 0.0f )
	{
		o_rot qA = o_mul_rot( sA->dq, jr( j->frameA.q ) );
		o_rot qB = o_mul_rot( sB->dq, jr( j->frameB.q ) );
		float c = o_rot_angle( o_inv_mul_rot( qA, qB ) );
		float bias = j->angularSpring.biasRate * c;
		float massScale = j->angularSpring.massScale;
		float impulseScale = j->angularSpring.impulseScale;
		float cdot = wB - wA;
		float maxImpulse = ctx->h * j->maxSpringTorque;
		float oldImpulse = j->angularSpringImpulse;
		float impulse = -massScale * j->angularMass * ( cdot + bias ) - impulseScale * oldImpulse;
		j->angularSpringImpulse = o_clamp( oldImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j->angularSpringImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if ( j->maxVelocityTorque > 0.0f )
	{
		float cdot = wB - wA - j->angularVelocity;
		float impulse = -j->angularMass * cdot;
		float maxImpulse = ctx->h * j->maxVelocityTorque;
		float oldImpulse = j->angularVelocityImpulse;
		j->angularVelocityImpulse = o_clamp( oldImpulse + impulse, -maxImpulse, maxImpulse );
		impulse = j->angularVelocityImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );

	if ( j->maxSpringForce > 0.0f && j->linearHertz > 0.0f )
	{
		o_vec2 c = o_add( o_add( o_sub( sB->dp, sA->dp ), o_sub( rB, rA ) ), jv( j->deltaCenter ) );
		o_vec2 bias = o_mul_sv( j->linearSpring.biasRate, c );
		float massScale = j->linearSpring.massScale;
		float impulseScale = j->linearSpring.impulseScale;
		o_vec2 cdot = o_sub( o_add( vB, o_cross_sv( wB, rB ) ), o_add( vA, o_cross_sv( wA, rA ) ) );
		cdot = o_add( cdot, bias );

		/* kl and its inverse (b2GetInverse22 math_functions.h:730); the inverse is written back into the joint */
		float kxx = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		float kxy = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		float kyy = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		{
			float a = kxx, b = kxy, cc = kxy, d = kyy;
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
		o_vec2 b = mat_mul( j->linearMass, cdot );
		o_vec2 oldImpulse = jv( j->linearSpringImpulse );
		o_vec2 impulse = { -massScale * b.x - impulseScale * oldImpulse.x, -massScale * b.y - impulseScale * oldImpulse.y };
		float maxImpulse = ctx->h * j->maxSpringForce;
		o_vec2 total = o_add( oldImpulse, impulse );
		if ( o_length_sq( total ) > maxImpulse * maxImpulse )
		{
			total = o_normalize( total );
			total.x *= maxImpulse;
			total.y *= maxImpulse;
		}
		j->linearSpringImpulse.x = total.x;
		j->linearSpringImpulse.y = total.y;
		impulse = o_sub( total, oldImpulse );
		point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, impulse );
	}

	if ( j->maxVelocityForce > 0.0f )
	{
		o_vec2 cdot = o_sub( o_add( vB, o_cross_sv( wB, rB ) ), o_add( vA, o_cross_sv( wA, rA ) ) );
		cdot = o_sub( cdot, jv( j->linearVelocity ) );
		o_vec2 b = mat_mul( j->linearMass, cdot );
		o_vec2 impulse = { -b.x, -b.y };
		o_vec2 oldImpulse = jv( j->linearVelocityImpulse );
		float maxImpulse = ctx->h * j->maxVelocityForce;
		o_vec2 total = o_add( oldImpulse, impulse );
		if ( o_length_sq( total ) > maxImpulse * maxImpulse )
		{
			total = o_normalize( total );
			total.x *= maxImpulse;
			total.y *= maxImpulse;
		}
		j->linearVelocityImpulse.x = total.x;
		j->linearVelocityImpulse.y = total.y;
		impulse = o_sub( total, oldImpulse );
		point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, impulse );
	}
	write_back( sA, vA, wA, sB, vB, wB );
}

/* ---- mover: src/mover_joint.c:94, :120 ----------------------------------------------------------------------------------- */
static void warm_mover( b2lJointSim* base, const o_ctx* ctx )
{
	b2lMover* j = &base->u.mover;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	if ( sA->flags & B2L_FLAG_DYNAMIC )
	{
		sA->v = o_mul_sub( sA->v, base->invMassA, jv( j->linearVelocityImpulse ) );
	}
	if ( sB->flags & B2L_FLAG_DYNAMIC )
	{
		sB->v = o_mul_add( sB->v, base->invMassB, jv( j->linearVelocityImpulse ) );
	}
}

static void solve_mover( b2lJointSim* base, const o_ctx* ctx )
{
	float mA = base->invMassA, mB = base->invMassB;
	b2lMover* j = &base->u.mover;
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;

	if ( j->maxVelocityForce.x > 0.0f || j->maxVelocityForce.y > 0.0f )
	{
		o_vec2 cdot = o_sub( o_sub( vB, vA ), jv( j->linearVelocity ) );
		o_vec2 b = o_mul_sv( j->linearMass, cdot );
		o_vec2 impulse = { -b.x, -b.y };
		o_vec2 oldImpulse = jv( j->linearVelocityImpulse );
		o_vec2 total = o_add( oldImpulse, impulse );
		o_vec2 maxImpulse = o_mul_sv( ctx->h, jv( j->maxVelocityForce ) );
		total.x = o_clamp( total.x, -maxImpulse.x, maxImpulse.x );
		total.y = o_clamp( total.y, -maxImpulse.y, maxImpulse.y );
		j->linearVelocityImpulse.x = total.x;
		j->linearVelocityImpulse.y = total.y;
		impulse = o_sub( total, oldImpulse );
		vA = o_mul_sub( vA, mA, impulse );
		vB = o_mul_add( vB, mB, impulse );
	}
	else
	{
		j->linearVelocityImpulse.x = 0.0f;
		j->linearVelocityImpulse.y = 0.0f;
	}
	if ( sA->flags & B2L_FLAG_DYNAMIC )
	{
		sA->v = vA;
	}
	if ( sB->flags & B2L_FLAG_DYNAMIC )
	{
		sB->v = vB;
	}
}

/* ---- pogo: src/pogo_joint.c:160, :200 -------------------------------------------------------------------------------------- */
static void warm_pogo( b2lJointSim* base, const o_ctx* ctx )
{
	b2lPogo* j = &base->u.pogo;
	if ( j->hertz == 0.0f )
	{
		j->impulse = 0.0f;
		return;
	}
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );
	o_vec2 L = o_mul_sv( j->impulse, jv( j->normal ) );
	warm_apply( sA, sB, base->invMassA, base->invIA, base->invMassB, base->invIB, L, o_cross( rA, L ), o_cross( rB, L ) );
}

static void solve_pogo( b2lJointSim* base, const o_ctx* ctx, bool useBias )
{
	float mA = base->invMassA, mB = base->invMassB, iA = base->invIA, iB = base->invIB;
	b2lPogo* j = &base->u.pogo;
	if ( j->hertz == 0.0f )
	{
		j->impulse = 0.0f;
		return;
	}
	o_state dummy = o_identity_state;
	o_state* sA = state_of( ctx, j->indexA, &dummy );
	o_state* sB = state_of( ctx, j->indexB, &dummy );
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;
	o_vec2 rA = o_rotate( sA->dq, jv( j->frameA.p ) );
	o_vec2 rB = o_rotate( sB->dq, jv( j->frameB.p ) );
	o_vec2 normal = jv( j->normal );

	float bias = 0.0f;
	if ( useBias )
	{
		o_vec2 d = o_add( o_add( o_sub( sB->dp, sA->dp ), o_sub( rB, rA ) ), jv( j->deltaCenter ) );
		o_vec2 pogoAxis = o_rotate( jr( j->frameB.q ), o_v( 0.0f, 1.0f ) );
		float c = o_dot( pogoAxis, d ) - j->restLength;
		j->velocity = o_spring_damper( j->hertz, j->dampingRatio, c, j->velocity, ctx->h );
		bias = -j->velocity;
	}

	o_vec2 vr = o_sub( o_add( vB, o_cross_sv( wB, rB ) ), o_add( vA, o_cross_sv( wA, rA ) ) );
	float cdot = o_dot( normal, vr );
	float maxTensionImpulse = ctx->h * j->maxTensionForce;
	float maxCompressionImpulse = ctx->h * j->maxCompressionForce;
	float oldImpulse = j->impulse;
	float impulse = -j->linearMass * ( cdot + bias );
	j->impulse = o_clamp( oldImpulse + impulse, -maxTensionImpulse, maxCompressionImpulse );
	impulse = j->impulse - oldImpulse;
	point_apply( &vA, &wA, &vB, &wB, mA, iA, mB, iB, rA, rB, o_mul_sv( impulse, normal ) );
	write_back( sA, vA, wA, sB, vB, wB );
}

/* ---- dispatch: src/joint.c:1454-1540 ------------------------------------------------------------------------------------------ */
void b2o_warm_start_joint( b2lJointSim* joint, const o_ctx* ctx )
{
	switch ( joint->type )
	{
		case b2l_distanceJoint: warm_distance( joint, ctx ); break;
		case b2l_motorJoint: warm_motor( joint, ctx ); break;
		case b2l_moverJoint: warm_mover( joint, ctx ); break;
		case b2l_pogoJoint: warm_pogo( joint, ctx ); break;
		case b2l_prismaticJoint: warm_prismatic( joint, ctx ); break;
		case b2l_revoluteJoint: warm_revolute( joint, ctx ); break;
		case b2l_weldJoint: warm_weld( joint, ctx ); break;
		case b2l_wheelJoint: warm_wheel( joint, ctx ); break;
		default: break; /* filter joint */
	}
}

void b2o_solve_joint( b2lJointSim* joint, const o_ctx* ctx, bool useBias )
{
	switch ( joint->type )
	{
		case b2l_distanceJoint: solve_distance( joint, ctx, useBias ); break;
		case b2l_motorJoint: solve_motor( joint, ctx ); break;
		case b2l_moverJoint: solve_mover( joint, ctx ); break;
		case b2l_pogoJoint: solve_pogo( joint, ctx, useBias ); break;
		case b2l_prismaticJoint: solve_prismatic( joint, ctx, useBias ); break;
		case b2l_revoluteJoint: solve_revolute( joint, ctx, useBias ); break;
		case b2l_weldJoint: solve_weld( joint, ctx, useBias ); break;
		case b2l_wheelJoint: solve_wheel( joint, ctx, useBias ); break;
		default: break;
	}
}

/* b2GetJointReaction src/joint.c:993-1075 */
void b2o_joint_reaction( const b2lJointSim* sim, float invTimeStep, float* force, float* torque )
{
	float linearImpulse = 0.0f, angularImpulse = 0.0f;
	switch ( sim->type )
	{
		case b2l_distanceJoint:
			linearImpulse = o_abs( sim->u.distance.impulse + sim->u.distance.lowerImpulse - sim->u.distance.upperImpulse +
								   sim->u.distance.motorImpulse );
			break;
		case b2l_motorJoint:
			linearImpulse = o_length( o_add( jv( sim->u.motor.linearVelocityImpulse ), jv( sim->u.motor.linearSpringImpulse ) ) );
			angularImpulse = o_abs( sim->u.motor.angularVelocityImpulse + sim->u.motor.angularSpringImpulse );
			break;
		case b2l_moverJoint:
			linearImpulse = o_length( jv( sim->u.mover.linearVelocityImpulse ) );
			break;
		case b2l_pogoJoint:
			linearImpulse = o_abs( sim->u.pogo.impulse );
			break;
		case b2l_prismaticJoint:
		{
			float perpImpulse = sim->u.prismatic.impulse.x;
			float axialImpulse = sim->u.prismatic.motorImpulse + sim->u.prismatic.lowerImpulse - sim->u.prismatic.upperImpulse;
			linearImpulse = sqrtf( perpImpulse * perpImpulse + axialImpulse * axialImpulse );
			angularImpulse = o_abs( sim->u.prismatic.impulse.y );
		}
		break;
		case b2l_revoluteJoint:
			linearImpulse = o_length( jv( sim->u.revolute.linearImpulse ) );
			angularImpulse = o_abs( sim->u.revolute.motorImpulse + sim->u.revolute.lowerImpulse - sim->u.revolute.upperImpulse );
			break;
		case b2l_weldJoint:
			linearImpulse = o_length( jv( sim->u.weld.linearImpulse ) );
			angularImpulse = o_abs( sim->u.weld.angularImpulse );
			break;
		case b2l_wheelJoint:
		{
			float perpImpulse = sim->u.wheel.perpImpulse;
			float axialImpulse = sim->u.wheel.springImpulse + sim->u.wheel.lowerImpulse - sim->u.wheel.upperImpulse;
			linearImpulse = sqrtf( perpImpulse * perpImpulse + axialImpulse * axialImpulse );
			angularImpulse = o_abs( sim->u.wheel.motorImpulse );
		}
		break;
		default:
			break;
	}
	*force = linearImpulse * invTimeStep;
	*torque = angularImpulse * invTimeStep;
}
