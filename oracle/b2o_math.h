/*
 * b2o_math.h -- scalar math of the CPU oracle (TEST INFRASTRUCTURE, never linked into the product).
 *
 * Plain-C restatement of the inlines the Soft Step solver uses from the reference's
 * include/box2d/math_functions.h and src/math_functions.c, cited per function.  Compile with
 * -ffp-contract=off (no FMA) like the reference (CMakeLists.txt:50-63).
 */
#ifndef B2O_MATH_H
#define B2O_MATH_H

#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stdint.h>

typedef struct
{
	float x, y;
} o_vec2;

typedef struct
{
	float c, s;
} o_rot;

typedef struct
{
	float biasRate, massScale, impulseScale;
} o_soft;

#define O_PI 3.14159265359f /* B2_PI math_functions.h:18 */

/* math_functions.h:170-191: second operand wins on ties/NaN, same as SSE2 MINPS/MAXPS */
static inline float o_min( float a, float b )
{
	return a < b ? a : b;
}

static inline float o_max( float a, float b )
{
	return a > b ? a : b;
}

static inline float o_abs( float a )
{
	return a < 0 ? -a : a;
}

static inline float o_clamp( float a, float lo, float hi )
{
	return a < lo ? lo : ( a > hi ? hi : a );
}

static inline o_vec2 o_v( float x, float y )
{
	o_vec2 r = { x, y };
	return r;
}

static inline float o_dot( o_vec2 a, o_vec2 b ) /* :204 */
{
	return a.x * b.x + a.y * b.y;
}

static inline float o_cross( o_vec2 a, o_vec2 b ) /* :210 */
{
	return a.x * b.y - a.y * b.x;
}

static inline o_vec2 o_cross_sv( float s, o_vec2 v ) /* :222 */
{
	return o_v( -s * v.y, s * v.x );
}

static inline o_vec2 o_left_perp( o_vec2 v ) /* :228 */
{
	return o_v( -v.y, v.x );
}

static inline o_vec2 o_right_perp( o_vec2 v ) /* :234 */
{
	return o_v( v.y, -v.x );
}

static inline o_vec2 o_add( o_vec2 a, o_vec2 b )
{
	return o_v( a.x + b.x, a.y + b.y );
}

static inline o_vec2 o_sub( o_vec2 a, o_vec2 b )
{
	return o_v( a.x - b.x, a.y - b.y );
}

static inline o_vec2 o_mul_sv( float s, o_vec2 v ) /* :271 */
{
	return o_v( s * v.x, s * v.y );
}

static inline o_vec2 o_mul_add( o_vec2 a, float s, o_vec2 b ) /* :277 */
{
	return o_v( a.x + s * b.x, a.y + s * b.y );
}

static inline o_vec2 o_mul_sub( o_vec2 a, float s, o_vec2 b ) /* :283 */
{
	return o_v( a.x - s * b.x, a.y - s * b.y );
}

static inline float o_length( o_vec2 v ) /* :325 */
{
	return sqrtf( v.x * v.x + v.y * v.y );
}

static inline float o_length_sq( o_vec2 v ) /* :402 */
{
	return v.x * v.x + v.y * v.y;
}

static inline o_vec2 o_normalize( o_vec2 a ) /* :340 */
{
	float lengthSquared = a.x * a.x + a.y * a.y;
	if ( lengthSquared > 1000.0f * FLT_MIN )
	{
		float s = 1.0f / sqrtf( lengthSquared );
		return o_v( s * a.x, s * a.y );
	}
	return o_v( 0.0f, 0.0f );
}

static inline o_vec2 o_rotate( o_rot q, o_vec2 v ) /* :558 */
{
	return o_v( q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y );
}

static inline o_rot o_mul_rot( o_rot q, o_rot r ) /* :503 */
{
	o_rot qr;
	qr.s = q.s * r.c + q.c * r.s;
	qr.c = q.c * r.c - q.s * r.s;
	return qr;
}

static inline o_rot o_inv_mul_rot( o_rot a, o_rot b ) /* :517 */
{
	o_rot r;
	r.s = a.c * b.s - a.s * b.c;
	r.c = a.c * b.c + a.s * b.s;
	return r;
}

static inline o_rot o_integrate_rotation( o_rot q1, float deltaAngle ) /* :388 */
{
	o_rot q2 = { q1.c - deltaAngle * q1.s, q1.s + deltaAngle * q1.c };
	float mag = sqrtf( q2.s * q2.s + q2.c * q2.c );
	float invMag = mag > 0.0f ? 1.0f / mag : 0.0f;
	o_rot qn = { q2.c * invMag, q2.s * invMag };
	return qn;
}

/* b2Atan2, src/math_functions.c:96-136 */
static inline float o_atan2( float y, float x )
{
	if ( x == 0.0f && y == 0.0f )
	{
		return 0.0f;
	}
	float ax = o_abs( x );
	float ay = o_abs( y );
	float mx = o_max( ay, ax );
	float mn = o_min( ay, ax );
	float a = mn / mx;
	float s = a * a;
	float c = s * a;
	float q = s * s;
	float r = 0.024840285f * q + 0.18681418f;
	float t = -0.094097948f * q - 0.33213072f;
	r = r * s + t;
	r = r * c + a;
	if ( ay > ax )
	{
		r = 1.57079637f - r;
	}
	if ( x < 0 )
	{
		r = 3.14159274f - r;
	}
	if ( y < 0 )
	{
		r = -r;
	}
	return r;
}

static inline float o_rot_angle( o_rot q ) /* :483 */
{
	return o_atan2( q.s, q.c );
}

/* b2UnwindAngle math_functions.h:540-555 (binary64 inside) */
static inline float o_unwind_angle( float radians )
{
	float x = o_clamp( radians, -1.0e6f, 1.0e6f );
	double twoPi = 2.0f * O_PI;
	double roundToNearest = 6755399441055744.0;
	double a = x;
	double k = ( a / twoPi + roundToNearest ) - roundToNearest;
	return (float)( a - k * twoPi );
}

/* b2Solve22 math_functions.h:748, A = [a11 a12; a21 a22] */
static inline o_vec2 o_solve22( float a11, float a12, float a21, float a22, o_vec2 b )
{
	float det = a11 * a22 - a12 * a21;
	if ( det != 0.0f )
	{
		det = 1.0f / det;
	}
	return o_v( det * ( a22 * b.x - a12 * b.y ), det * ( a11 * b.y - a21 * b.x ) );
}

/* b2MakeSoft src/solver.h:239-281 */
static inline o_soft o_make_soft( float hertz, float zeta, float h )
{
	o_soft r = { 0.0f, 0.0f, 0.0f };
	if ( hertz == 0.0f )
	{
		return r;
	}
	float omega = 2.0f * O_PI * hertz;
	float a1 = 2.0f * zeta + h * omega;
	float a2 = h * omega * a1;
	float a3 = 1.0f / ( 1.0f + a2 );
	r.biasRate = omega / a1;
	r.massScale = a2 * a3;
	r.impulseScale = a3;
	return r;
}

/* b2SpringDamper math_functions.h:832 */
static inline float o_spring_damper( float hertz, float dampingRatio, float position, float velocity, float timeStep )
{
	float omega = 2.0f * O_PI * hertz;
	float omegaH = omega * timeStep;
	return ( velocity - omega * omegaH * position ) / ( 1.0f + 2.0f * dampingRatio * omegaH + omegaH * omegaH );
}

#endif
