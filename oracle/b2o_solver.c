/*
 * b2o_solver.c -- CPU oracle: stage sequence, body integration and contact constraints (TEST INFRASTRUCTURE).
 * See b2o_solver.h for the rules that govern this directory.
 *
 * Coloured contacts follow the reference's WIDE path per lane (src/contact_solver.c:1573-2331, default SSE2 build,
 * 4 lanes: the two "all lanes zero" early-outs are evaluated over aligned groups of 4 constraints of a colour);
 * overflow contacts follow the SCALAR path (src/contact_solver.c:24-545).  Stage order: src/solver.c:1055-1197.
 */
#include "b2o_solver.h"

#include <stdlib.h>
#include <string.h>

_Static_assert( sizeof( o_state ) == B2L_STATE_SIZE, "o_state" );

#define O_SIMD_WIDTH 4 /* B2_SIMD_WIDTH of the default x86-64 build, src/core.h:50-75 */

/* One prepared contact constraint (the per-lane content of b2ContactConstraintWide, src/contact_solver.c:1070-1100,
 * and of b2ContactConstraint, src/contact_solver.h:22-39) */
typedef struct
{
	o_vec2 anchorA, anchorB;
	float baseSeparation, relativeVelocity;
	float normalImpulse, tangentImpulse, totalNormalImpulse;
	float normalMass, tangentMass;
} o_point;

typedef struct
{
	int indexA, indexB; /* 0-based, -1 = static */
	float invMassA, invIA, invMassB, invIB;
	o_vec2 normal;
	float friction, restitution, tangentSpeed, rollingResistance, rollingMass, rollingImpulse;
	o_soft softness;
	o_point points[2];
	int pointCount;
} o_contact;

static const o_state o_identity = { { 0.0f, 0.0f }, 0.0f, 0u, { 0.0f, 0.0f }, { 1.0f, 0.0f } };

static inline float rd_f( const uint8_t* p, int offset )
{
	float v;
	memcpy( &v, p + offset, 4 );
	return v;
}

static inline int rd_i( const uint8_t* p, int offset )
{
	int v;
	memcpy( &v, p + offset, 4 );
	return v;
}

static inline void wr_f( uint8_t* p, int offset, float v )
{
	memcpy( p + offset, &v, 4 );
}

/* ---- prepare: b2PrepareContactsTask :1573 (wide) / b2PrepareContacts_Overflow :24 ---------------------------------- */
static void prepare_contact( o_contact* c, const uint8_t* sim, const o_state* states, const b2GpuStepDesc* d, bool wide )
{
	const uint8_t* m = sim + B2L_CONTACT_MANIFOLD;
	int indexA = rd_i( sim, B2L_CONTACT_INDEX_A );
	int indexB = rd_i( sim, B2L_CONTACT_INDEX_B );
	c->indexA = indexA;
	c->indexB = indexB;

	float mA = rd_f( sim, B2L_CONTACT_INV_MASS_A ), iA = rd_f( sim, B2L_CONTACT_INV_I_A );
	float mB = rd_f( sim, B2L_CONTACT_INV_MASS_B ), iB = rd_f( sim, B2L_CONTACT_INV_I_B );
	c->invMassA = mA;
	c->invIA = iA;
	c->invMassB = mB;
	c->invIB = iB;

	o_vec2 vA = { 0.0f, 0.0f }, vB = { 0.0f, 0.0f };
	float wA = 0.0f, wB = 0.0f;
	if ( indexA != -1 )
	{
		vA = states[indexA].v;
		wA = states[indexA].w;
	}
	if ( indexB != -1 )
	{
		vB = states[indexB].v;
		wB = states[indexB].w;
	}

	{
		float k = iA + iB;
		c->rollingMass = k > 0.0f ? 1.0f / k : 0.0f;
	}

	o_soft soft = { d->contactSoftness.biasRate, d->contactSoftness.massScale, d->contactSoftness.impulseScale };
	if ( indexA == -1 || indexB == -1 )
	{
		o_soft s = { d->staticSoftness.biasRate, d->staticSoftness.massScale, d->staticSoftness.impulseScale };
		soft = s;
	}
	else if ( wide && d->enableContactSoftening )
	{
		float contactHertz = o_min( d->contactHertz, 0.125f * d->inv_h );
		float ratio = 1.0f;
		if ( mA < mB )
		{
			ratio = o_max( 0.5f, mA / mB );
		}
		else if ( mB < mA )
		{
			ratio = o_max( 0.5f, mB / mA );
		}
		soft = o_make_soft( ratio * contactHertz, ratio * d->contactDampingRatio, d->h );
	}
	c->softness = soft;

	float warmStartScale = d->enableWarmStarting ? 1.0f : 0.0f;
	o_vec2 normal = { rd_f( m, B2L_MANIFOLD_NORMAL ), rd_f( m, B2L_MANIFOLD_NORMAL + 4 ) };
	c->normal = normal;
	c->friction = rd_f( sim, B2L_CONTACT_FRICTION );
	c->restitution = rd_f( sim, B2L_CONTACT_RESTITUTION );
	c->rollingResistance = rd_f( sim, B2L_CONTACT_ROLLING_RESISTANCE );
	c->tangentSpeed = rd_f( sim, B2L_CONTACT_TANGENT_SPEED );
	c->rollingImpulse = warmStartScale * rd_f( m, B2L_MANIFOLD_ROLLING_IMPULSE );
	c->pointCount = rd_i( m, B2L_MANIFOLD_POINT_COUNT );

	o_vec2 tangent = o_right_perp( normal );
	for ( int j = 0; j < 2; ++j )
	{
		o_point* cp = c->points + j;
		if ( j >= c->pointCount )
		{
			/* dummy data that has no effect (wide path, :1786-1800) */
			memset( cp, 0, sizeof( *cp ) );
			continue;
		}
		const uint8_t* mp = m + B2L_MANIFOLD_POINTS + j * B2L_MP_SIZE;
		o_vec2 rA = { rd_f( mp, B2L_MP_ANCHOR_A ), rd_f( mp, B2L_MP_ANCHOR_A + 4 ) };
		o_vec2 rB = { rd_f( mp, B2L_MP_ANCHOR_B ), rd_f( mp, B2L_MP_ANCHOR_B + 4 ) };
		cp->anchorA = rA;
		cp->anchorB = rB;
		cp->baseSeparation = rd_f( mp, B2L_MP_SEPARATION ) - o_dot( o_sub( rB, rA ), normal );
		cp->normalImpulse = warmStartScale * rd_f( mp, B2L_MP_NORMAL_IMPULSE );
		cp->tangentImpulse = warmStartScale * rd_f( mp, B2L_MP_TANGENT_IMPULSE );
		cp->totalNormalImpulse = 0.0f;

		float rnA = o_cross( rA, normal );
		float rnB = o_cross( rB, normal );
		float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
		cp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;

		float rtA = o_cross( rA, tangent );
		float rtB = o_cross( rB, tangent );
		float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
		cp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;

		o_vec2 vrA = o_add( vA, o_cross_sv( wA, rA ) );
		o_vec2 vrB = o_add( vB, o_cross_sv( wB, rB ) );
		cp->relativeVelocity = o_dot( normal, o_sub( vrB, vrA ) );
	}
}

/* ---- wide path helpers: expression order of the b2FloatW code, one lane ---------------------------------------------- */
typedef struct
{
	float vx, vy, w;
} o_vel;

static inline const o_state* gather( const o_state* states, int index )
{
	return index == -1 ? &o_identity : states + index; /* b2GatherBodies :1490 */
}

static inline void scatter( o_state* states, int index, o_vel b )
{
	if ( index != -1 && ( states[index].flags & B2L_FLAG_DYNAMIC ) != 0 ) /* b2ScatterBodies :1521 */
	{
		states[index].v.x = b.vx;
		states[index].v.y = b.vy;
		states[index].w = b.w;
	}
}

static inline void apply_wide( o_vel* bA, o_vel* bB, const o_contact* c, o_vec2 rA, o_vec2 rB, float Px, float Py )
{
	bA->vx = bA->vx - c->invMassA * Px;
	bA->vy = bA->vy - c->invMassA * Py;
	bA->w = bA->w - c->invIA * ( rA.x * Py - rA.y * Px );
	bB->vx = bB->vx + c->invMassB * Px;
	bB->vy = bB->vy + c->invMassB * Py;
	bB->w = bB->w + c->invIB * ( rB.x * Py - rB.y * Px );
}

/* b2WarmStartContactsTask :1811 */
static void warm_start_wide( o_contact* c, o_state* states )
{
	const o_state* sA = gather( states, c->indexA );
	const o_state* sB = gather( states, c->indexB );
	o_vel bA = { sA->v.x, sA->v.y, sA->w }, bB = { sB->v.x, sB->v.y, sB->w };
	float tangentX = c->normal.y;
	float tangentY = 0.0f - c->normal.x;
	for ( int j = 0; j < 2; ++j )
	{
		o_point* cp = c->points + j;
		float Px = cp->normalImpulse * c->normal.x + cp->tangentImpulse * tangentX;
		float Py = cp->normalImpulse * c->normal.y + cp->tangentImpulse * tangentY;
		/* :1835-1840 updates w before v; the operands are independent so the bits are the same */
		apply_wide( &bA, &bB, c, cp->anchorA, cp->anchorB, Px, Py );
		cp->totalNormalImpulse = cp->totalNormalImpulse + cp->normalImpulse;
	}
	bA.w = bA.w - c->invIA * c->rollingImpulse;
	bB.w = bB.w + c->invIB * c->rollingImpulse;
	scatter( states, c->indexA, bA );
	scatter( states, c->indexB, bB );
}

/* b2SolveContactsTask :1873 */
static void solve_wide( o_contact* c, o_state* states, const b2GpuStepDesc* d, bool useBias, bool groupHasRolling )
{
	const o_state* sA = gather( states, c->indexA );
	const o_state* sB = gather( states, c->indexB );
	o_vel bA = { sA->v.x, sA->v.y, sA->w }, bB = { sB->v.x, sB->v.y, sB->w };
	o_rot dqA = sA->dq, dqB = sB->dq;

	float biasRate, massScale, impulseScale;
	if ( useBias )
	{
		biasRate = c->softness.massScale * c->softness.biasRate;
		massScale = c->softness.massScale;
		impulseScale = c->softness.impulseScale;
	}
	else
	{
		biasRate = 0.0f;
		massScale = 1.0f;
		impulseScale = 0.0f;
	}

	float totalNormalImpulse = 0.0f;
	float dpx = sB->dp.x - sA->dp.x;
	float dpy = sB->dp.y - sA->dp.y;
	float negContactSpeed = -d->contactSpeed;
	float nx = c->normal.x, ny = c->normal.y;

	for ( int j = 0; j < 2; ++j )
	{
		o_point* cp = c->points + j;
		o_vec2 rA = cp->anchorA, rB = cp->anchorB;
		o_vec2 rsA = { dqA.c * rA.x - dqA.s * rA.y, dqA.s * rA.x + dqA.c * rA.y };
		o_vec2 rsB = { dqB.c * rB.x - dqB.s * rB.y, dqB.s * rB.x + dqB.c * rB.y };
		float dsx = dpx + ( rsB.x - rsA.x );
		float dsy = dpy + ( rsB.y - rsA.y );
		float s = ( nx * dsx + ny * dsy ) + cp->baseSeparation;

		bool mask = s > 0.0f;
		float specBias = s * d->inv_h;
		float softBias = o_max( biasRate * s, negContactSpeed );
		float bias = mask ? specBias : softBias;
		float pointMassScale = mask ? 1.0f : massScale;
		float pointImpulseScale = mask ? 0.0f : impulseScale;

		float dvx = ( bB.vx - bB.w * rB.y ) - ( bA.vx - bA.w * rA.y );
		float dvy = ( bB.vy + bB.w * rB.x ) - ( bA.vy + bA.w * rA.x );
		float vn = dvx * nx + dvy * ny;

		float negImpulse = cp->normalMass * ( pointMassScale * vn + bias ) + pointImpulseScale * cp->normalImpulse;
		float newImpulse = o_max( cp->normalImpulse - negImpulse, 0.0f );
		float impulse = newImpulse - cp->normalImpulse;
		cp->normalImpulse = newImpulse;
		cp->totalNormalImpulse = cp->totalNormalImpulse + impulse;
		totalNormalImpulse = totalNormalImpulse + newImpulse;

		apply_wide( &bA, &bB, c, rA, rB, impulse * nx, impulse * ny );
	}

	if ( useBias == false )
	{
		if ( groupHasRolling )
		{
			float deltaLambda = c->rollingMass * ( bA.w - bB.w );
			float lambda = c->rollingImpulse;
			float maxLambda = c->rollingResistance * totalNormalImpulse;
			/* b2SymClampW, SSE2 flavour: lower bound by sign flip (:869-878) */
			c->rollingImpulse = o_max( -maxLambda, o_min( lambda + deltaLambda, maxLambda ) );
			deltaLambda = c->rollingImpulse - lambda;
			bA.w = bA.w - c->invIA * deltaLambda;
			bB.w = bB.w + c->invIB * deltaLambda;
		}

		float tangentX = ny;
		float tangentY = 0.0f - nx;
		for ( int j = 0; j < 2; ++j )
		{
			o_point* cp = c->points + j;
			o_vec2 rA = cp->anchorA, rB = cp->anchorB;
			float dvx = ( bB.vx - bB.w * rB.y ) - ( bA.vx - bA.w * rA.y );
			float dvy = ( bB.vy + bB.w * rB.x ) - ( bA.vy + bA.w * rA.x );
			float vt = dvx * tangentX + dvy * tangentY;
			vt = vt - c->tangentSpeed;
			float negImpulse = cp->tangentMass * vt;
			float maxFriction = c->friction * cp->normalImpulse;
			float newImpulse = cp->tangentImpulse - negImpulse;
			newImpulse = o_max( 0.0f - maxFriction, o_min( newImpulse, maxFriction ) );
			float impulse = newImpulse - cp->tangentImpulse;
			cp->tangentImpulse = newImpulse;
			apply_wide( &bA, &bB, c, rA, rB, impulse * tangentX, impulse * tangentY );
		}
	}
	scatter( states, c->indexA, bA );
	scatter( states, c->indexB, bB );
}

/* b2ApplyRestitutionTask :2118 (called only for groups where some lane has restitution) */
static void restitution_wide( o_contact* c, o_state* states, float threshold )
{
	bool restitutionIsZero = c->restitution == 0.0f;
	const o_state* sA = gather( states, c->indexA );
	const o_state* sB = gather( states, c->indexB );
	o_vel bA = { sA->v.x, sA->v.y, sA->w }, bB = { sB->v.x, sB->v.y, sB->w };
	float nx = c->normal.x, ny = c->normal.y;
	for ( int j = 0; j < 2; ++j )
	{
		o_point* cp = c->points + j;
		bool mask1 = ( cp->relativeVelocity + threshold ) > 0.0f;
		bool mask2 = cp->totalNormalImpulse == 0.0f;
		float mass = ( mask1 || mask2 || restitutionIsZero ) ? 0.0f : cp->normalMass;
		o_vec2 rA = cp->anchorA, rB = cp->anchorB;
		float dvx = ( bB.vx - bB.w * rB.y ) - ( bA.vx - bA.w * rA.y );
		float dvy = ( bB.vy + bB.w * rB.x ) - ( bA.vy + bA.w * rA.x );
		float vn = dvx * nx + dvy * ny;
		float negImpulse = mass * ( vn + c->restitution * cp->relativeVelocity );
		float newImpulse = o_max( cp->normalImpulse - negImpulse, 0.0f );
		float deltaImpulse = newImpulse - cp->normalImpulse;
		cp->normalImpulse = newImpulse;
		cp->totalNormalImpulse = cp->totalNormalImpulse + deltaImpulse;
		apply_wide( &bA, &bB, c, rA, rB, deltaImpulse * nx, deltaImpulse * ny );
	}
	scatter( states, c->indexA, bA );
	scatter( states, c->indexB, bB );
}

/* ---- scalar path (overflow colour) -------------------------------------------------------------------------------------- */
static inline void scalar_apply( o_vec2* vA, float* wA, o_vec2* vB, float* wB, const o_contact* c, o_vec2 rA, o_vec2 rB, o_vec2 P )
{
	*vA = o_mul_sub( *vA, c->invMassA, P );
	*wA -= c->invIA * o_cross( rA, P );
	*vB = o_mul_add( *vB, c->invMassB, P );
	*wB += c->invIB * o_cross( rB, P );
}

static inline void scalar_store( o_state* sA, o_vec2 vA, float wA, o_state* sB, o_vec2 vB, float wB )
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

/* b2WarmStartContacts_Overflow :162 */
static void warm_start_overflow( o_contact* c, o_state* states )
{
	o_state dummy = o_identity;
	o_state* sA = c->indexA == -1 ? &dummy : states + c->indexA;
	o_state* sB = c->indexB == -1 ? &dummy : states + c->indexB;
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;
	o_vec2 tangent = o_right_perp( c->normal );
	for ( int j = 0; j < c->pointCount; ++j )
	{
		o_point* cp = c->points + j;
		o_vec2 P = o_add( o_mul_sv( cp->normalImpulse, c->normal ), o_mul_sv( cp->tangentImpulse, tangent ) );
		cp->totalNormalImpulse += cp->normalImpulse;
		wA -= c->invIA * o_cross( cp->anchorA, P );
		vA = o_mul_add( vA, -c->invMassA, P );
		wB += c->invIB * o_cross( cp->anchorB, P );
		vB = o_mul_add( vB, c->invMassB, P );
	}
	wA -= c->invIA * c->rollingImpulse;
	wB += c->invIB * c->rollingImpulse;
	scalar_store( sA, vA, wA, sB, vB, wB );
}

/* b2SolveContacts_Overflow :239 (friction BEFORE rolling resistance) */
static void solve_overflow( o_contact* c, o_state* states, const b2GpuStepDesc* d, bool useBias )
{
	o_state dummy = o_identity;
	o_state* sA = c->indexA == -1 ? &dummy : states + c->indexA;
	o_state* sB = c->indexB == -1 ? &dummy : states + c->indexB;
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;
	o_rot dqA = sA->dq, dqB = sB->dq;
	o_vec2 dp = o_sub( sB->dp, sA->dp );
	o_vec2 normal = c->normal;
	o_vec2 tangent = o_right_perp( normal );
	float totalNormalImpulse = 0.0f;

	for ( int j = 0; j < c->pointCount; ++j )
	{
		o_point* cp = c->points + j;
		o_vec2 rA = cp->anchorA, rB = cp->anchorB;
		o_vec2 ds = o_add( dp, o_sub( o_rotate( dqB, rB ), o_rotate( dqA, rA ) ) );
		float s = cp->baseSeparation + o_dot( ds, normal );
		float velocityBias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if ( s > 0.0f )
		{
			velocityBias = s * d->inv_h;
		}
		else if ( useBias )
		{
			velocityBias = o_max( c->softness.massScale * c->softness.biasRate * s, -d->contactSpeed );
			massScale = c->softness.massScale;
			impulseScale = c->softness.impulseScale;
		}
		o_vec2 vrA = o_add( vA, o_cross_sv( wA, rA ) );
		o_vec2 vrB = o_add( vB, o_cross_sv( wB, rB ) );
		float vn = o_dot( o_sub( vrB, vrA ), normal );
		float impulse = -cp->normalMass * ( massScale * vn + velocityBias ) - impulseScale * cp->normalImpulse;
		float newImpulse = o_max( cp->normalImpulse + impulse, 0.0f );
		impulse = newImpulse - cp->normalImpulse;
		cp->normalImpulse = newImpulse;
		cp->totalNormalImpulse += impulse;
		totalNormalImpulse += newImpulse;
		scalar_apply( &vA, &wA, &vB, &wB, c, rA, rB, o_mul_sv( impulse, normal ) );
	}

	if ( useBias == false )
	{
		for ( int j = 0; j < c->pointCount; ++j )
		{
			o_point* cp = c->points + j;
			o_vec2 rA = cp->anchorA, rB = cp->anchorB;
			o_vec2 vrB = o_add( vB, o_cross_sv( wB, rB ) );
			o_vec2 vrA = o_add( vA, o_cross_sv( wA, rA ) );
			float vt = o_dot( o_sub( vrB, vrA ), tangent ) - c->tangentSpeed;
			float impulse = cp->tangentMass * ( -vt );
			float maxFriction = c->friction * cp->normalImpulse;
			float newImpulse = o_clamp( cp->tangentImpulse + impulse, -maxFriction, maxFriction );
			impulse = newImpulse - cp->tangentImpulse;
			cp->tangentImpulse = newImpulse;
			scalar_apply( &vA, &wA, &vB, &wB, c, rA, rB, o_mul_sv( impulse, tangent ) );
		}
		{
			float deltaLambda = -c->rollingMass * ( wB - wA );
			float lambda = c->rollingImpulse;
			float maxLambda = c->rollingResistance * totalNormalImpulse;
			c->rollingImpulse = o_clamp( lambda + deltaLambda, -maxLambda, maxLambda );
			deltaLambda = c->rollingImpulse - lambda;
			wA -= c->invIA * deltaLambda;
			wB += c->invIB * deltaLambda;
		}
	}
	scalar_store( sA, vA, wA, sB, vB, wB );
}

/* b2ApplyRestitution_Overflow :410 */
static void restitution_overflow( o_contact* c, o_state* states, float threshold )
{
	if ( c->restitution == 0.0f )
	{
		return;
	}
	o_state dummy = o_identity;
	o_state* sA = c->indexA == -1 ? &dummy : states + c->indexA;
	o_state* sB = c->indexB == -1 ? &dummy : states + c->indexB;
	o_vec2 vA = sA->v, vB = sB->v;
	float wA = sA->w, wB = sB->w;
	for ( int j = 0; j < c->pointCount; ++j )
	{
		o_point* cp = c->points + j;
		if ( cp->relativeVelocity > -threshold || cp->totalNormalImpulse == 0.0f )
		{
			continue;
		}
		o_vec2 rA = cp->anchorA, rB = cp->anchorB;
		o_vec2 vrB = o_add( vB, o_cross_sv( wB, rB ) );
		o_vec2 vrA = o_add( vA, o_cross_sv( wA, rA ) );
		float vn = o_dot( o_sub( vrB, vrA ), c->normal );
		float impulse = -cp->normalMass * ( vn + c->restitution * cp->relativeVelocity );
		float newImpulse = o_max( cp->normalImpulse + impulse, 0.0f );
		impulse = newImpulse - cp->normalImpulse;
		cp->normalImpulse = newImpulse;
		cp->totalNormalImpulse += impulse;
		scalar_apply( &vA, &wA, &vB, &wB, c, rA, rB, o_mul_sv( impulse, c->normal ) );
	}
	scalar_store( sA, vA, wA, sB, vB, wB );
}

/* ---- store: b2StoreImpulsesTask :2238 / b2StoreImpulses_Overflow :516 --------------------------------------------------- */
static void store_contact( const o_contact* c, uint8_t* sim, bool wide, const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	uint8_t* m = sim + B2L_CONTACT_MANIFOLD;
	wr_f( m, B2L_MANIFOLD_ROLLING_IMPULSE, c->rollingImpulse );
	int writeCount = wide ? 2 : c->pointCount;
	for ( int j = 0; j < writeCount; ++j )
	{
		uint8_t* mp = m + B2L_MANIFOLD_POINTS + j * B2L_MP_SIZE;
		wr_f( mp, B2L_MP_NORMAL_IMPULSE, c->points[j].normalImpulse );
		wr_f( mp, B2L_MP_TANGENT_IMPULSE, c->points[j].tangentImpulse );
		wr_f( mp, B2L_MP_TOTAL_NORMAL_IMPULSE, c->points[j].totalNormalImpulse );
		wr_f( mp, B2L_MP_NORMAL_VELOCITY, c->points[j].relativeVelocity );
	}
	if ( wide && ( (uint32_t)rd_i( sim, B2L_CONTACT_SIM_FLAGS ) & B2L_SIM_ENABLE_HIT_EVENT ) != 0 )
	{
		float negHitThreshold = -d->hitEventThreshold;
		for ( int k = 0; k < c->pointCount; ++k )
		{
			if ( c->points[k].relativeVelocity < negHitThreshold && c->points[k].totalNormalImpulse > 0.0f )
			{
				uint32_t id = (uint32_t)rd_i( sim, B2L_CONTACT_ID );
				if ( r != NULL && r->hitEventBits != NULL )
				{
					r->hitEventBits[id / 64] |= (uint64_t)1 << ( id % 64 );
				}
				if ( r != NULL )
				{
					r->hasHitEvents = 1;
				}
				break;
			}
		}
	}
}

/* ---- body stages: src/solver.c:66-162 -------------------------------------------------------------------------------------- */
static void integrate_velocities( o_state* states, const uint8_t* sims, int count, const b2GpuStepDesc* d )
{
	o_vec2 gravity = { d->gravity[0], d->gravity[1] };
	float h = d->h;
	for ( int i = 0; i < count; ++i )
	{
		const uint8_t* sim = sims + (size_t)i * B2L_SIM_SIZE;
		o_state* state = states + i;
		o_vec2 v = state->v;
		float w = state->w;
		float invMass = rd_f( sim, B2L_SIM_INV_MASS );
		float linearDamping = 1.0f / ( 1.0f + h * rd_f( sim, B2L_SIM_LINEAR_DAMPING ) );
		float angularDamping = 1.0f / ( 1.0f + h * rd_f( sim, B2L_SIM_ANGULAR_DAMPING ) );
		float gravityScale = invMass > 0.0f ? rd_f( sim, B2L_SIM_GRAVITY_SCALE ) : 0.0f;
		o_vec2 force = { rd_f( sim, B2L_SIM_FORCE ), rd_f( sim, B2L_SIM_FORCE + 4 ) };
		o_vec2 linearVelocityDelta = o_add( o_mul_sv( h * invMass, force ), o_mul_sv( h * gravityScale, gravity ) );
		float angularVelocityDelta = h * rd_f( sim, B2L_SIM_INV_INERTIA ) * rd_f( sim, B2L_SIM_TORQUE );
		v = o_mul_add( linearVelocityDelta, linearDamping, v );
		w = angularVelocityDelta + angularDamping * w;
		state->v = v;
		state->w = w;
	}
}

static void integrate_positions( o_state* states, int count, const b2GpuStepDesc* d )
{
	float h = d->h;
	float maxLinearSpeed = d->maxLinearVelocity;
	float maxAngularSpeed = ( 0.25f * O_PI ) * d->inv_dt; /* B2_MAX_ROTATION * inv_dt */
	float maxLinearSpeedSquared = maxLinearSpeed * maxLinearSpeed;
	float maxAngularSpeedSquared = maxAngularSpeed * maxAngularSpeed;
	for ( int i = 0; i < count; ++i )
	{
		o_state* state = states + i;
		o_vec2 v = state->v;
		float w = state->w;
		v.x = ( state->flags & B2L_FLAG_LOCK_LINEAR_X ) ? 0.0f : v.x;
		v.y = ( state->flags & B2L_FLAG_LOCK_LINEAR_Y ) ? 0.0f : v.y;
		w = ( state->flags & B2L_FLAG_LOCK_ANGULAR_Z ) ? 0.0f : w;
		if ( o_dot( v, v ) > maxLinearSpeedSquared )
		{
			float ratio = maxLinearSpeed / o_length( v );
			v = o_mul_sv( ratio, v );
			state->flags |= B2L_FLAG_IS_SPEED_CAPPED;
		}
		if ( w * w > maxAngularSpeedSquared && ( state->flags & B2L_FLAG_ALLOW_FAST_ROTATION ) == 0 )
		{
			float ratio = maxAngularSpeed / o_abs( w );
			w *= ratio;
			state->flags |= B2L_FLAG_IS_SPEED_CAPPED;
		}
		state->v = v;
		state->w = w;
		state->dp = o_mul_add( state->dp, h, state->v );
		state->dq = o_integrate_rotation( state->dq, h * state->w );
	}
}

/* ---- the step: src/solver.c:1055-1197 -------------------------------------------------------------------------------------------- */
typedef struct
{
	o_contact* contacts;
	int contactCount;
	b2lJointSim* joints;
	int jointCount;
	uint8_t* rawContacts;
} o_color;

static bool group_any( const o_color* color, int index, bool rolling )
{
	/* b2AllZeroW over the SIMD group of this lane (:2021, :2131); dead tail lanes are zero */
	int base = index & ~( O_SIMD_WIDTH - 1 );
	for ( int k = base; k < base + O_SIMD_WIDTH && k < color->contactCount; ++k )
	{
		float v = rolling ? color->contacts[k].rollingResistance : color->contacts[k].restitution;
		if ( !( v == 0.0f ) )
		{
			return true;
		}
	}
	return false;
}

int b2OracleSolverStep( const b2GpuStepDesc* d, b2GpuStepResult* r )
{
	o_state* states = (o_state*)d->states;
	const uint8_t* sims = (const uint8_t*)d->sims;
	int bodyCount = d->awakeBodyCount;
	int colorCount = d->activeColorCount;

	o_ctx ctx = { states, d->h, d->inv_h, d->inv_dt, d->lengthUnitsPerMeter };

	o_color colors[B2GPU_GRAPH_COLOR_COUNT];
	for ( int c = 0; c <= colorCount; ++c )
	{
		const b2GpuColorDesc* cd = c < colorCount ? d->colors + c : &d->overflow;
		colors[c].contactCount = cd->contactCount;
		colors[c].jointCount = cd->jointCount;
		colors[c].joints = (b2lJointSim*)cd->jointSims;
		colors[c].rawContacts = (uint8_t*)cd->contactSims;
		colors[c].contacts = cd->contactCount > 0 ? malloc( (size_t)cd->contactCount * sizeof( o_contact ) ) : NULL;
	}
	o_color* overflow = colors + colorCount;

	if ( r != NULL )
	{
		r->hasHitEvents = 0;
	}

	/* prepare (joints arrive prepared by the host, like the product's seam) */
	for ( int c = 0; c <= colorCount; ++c )
	{
		for ( int i = 0; i < colors[c].contactCount; ++i )
		{
			prepare_contact( colors[c].contacts + i, colors[c].rawContacts + (size_t)i * B2L_CONTACT_SIZE, states, d, c < colorCount );
		}
	}

	for ( int subStep = 0; subStep < d->subStepCount; ++subStep )
	{
		integrate_velocities( states, sims, bodyCount, d );

		/* warm start: overflow joints, overflow contacts, then colours ascending (joints and contacts of a colour
		 * are body-disjoint, src/constraint_graph.c:214-215, so their relative order is free) */
		for ( int i = 0; i < overflow->jointCount; ++i )
		{
			b2o_warm_start_joint( overflow->joints + i, &ctx );
		}
		for ( int i = 0; i < overflow->contactCount; ++i )
		{
			warm_start_overflow( overflow->contacts + i, states );
		}
		for ( int c = 0; c < colorCount; ++c )
		{
			for ( int i = 0; i < colors[c].jointCount; ++i )
			{
				b2o_warm_start_joint( colors[c].joints + i, &ctx );
			}
			for ( int i = 0; i < colors[c].contactCount; ++i )
			{
				warm_start_wide( colors[c].contacts + i, states );
			}
		}

		for ( int pass = 0; pass < 2; ++pass )
		{
			bool useBias = pass == 0;
			if ( pass == 1 )
			{
				integrate_positions( states, bodyCount, d );
			}
			for ( int i = 0; i < overflow->jointCount; ++i )
			{
				b2o_solve_joint( overflow->joints + i, &ctx, useBias );
			}
			for ( int i = 0; i < overflow->contactCount; ++i )
			{
				solve_overflow( overflow->contacts + i, states, d, useBias );
			}
			for ( int c = 0; c < colorCount; ++c )
			{
				for ( int i = 0; i < colors[c].jointCount; ++i )
				{
					b2lJointSim* joint = colors[c].joints + i;
					b2o_solve_joint( joint, &ctx, useBias );
					/* b2SolveJointsTask src/joint.c:1663-1675: coloured joints only */
					if ( useBias && ( joint->forceThreshold < FLT_MAX || joint->torqueThreshold < FLT_MAX ) )
					{
						float force, torque;
						b2o_joint_reaction( joint, d->inv_h, &force, &torque );
						if ( ( force >= joint->forceThreshold || torque >= joint->torqueThreshold ) && r != NULL &&
							 r->jointEventBits != NULL )
						{
							uint32_t id = (uint32_t)joint->jointId;
							r->jointEventBits[id / 64] |= (uint64_t)1 << ( id % 64 );
						}
					}
				}
				for ( int i = 0; i < colors[c].contactCount; ++i )
				{
					bool rolling = useBias ? false : group_any( colors + c, i, true );
					solve_wide( colors[c].contacts + i, states, d, useBias, rolling );
				}
			}
		}
	}

	/* restitution */
	for ( int i = 0; i < overflow->contactCount; ++i )
	{
		restitution_overflow( overflow->contacts + i, states, d->restitutionThreshold );
	}
	for ( int c = 0; c < colorCount; ++c )
	{
		for ( int i = 0; i < colors[c].contactCount; ++i )
		{
			if ( group_any( colors + c, i, false ) )
			{
				restitution_wide( colors[c].contacts + i, states, d->restitutionThreshold );
			}
		}
	}

	/* store impulses */
	for ( int c = 0; c <= colorCount; ++c )
	{
		for ( int i = 0; i < colors[c].contactCount; ++i )
		{
			store_contact( colors[c].contacts + i, colors[c].rawContacts + (size_t)i * B2L_CONTACT_SIZE, c < colorCount, d, r );
		}
		free( colors[c].contacts );
	}
	return 0;
}
