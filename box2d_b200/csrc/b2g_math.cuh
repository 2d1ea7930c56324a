// b2g_math.cuh -- device math for the Soft Step kernels.
//
// Every helper restates one inline of the reference's include/box2d/math_functions.h (cited per function)
// with the SAME association of floating point operations: the device code is compiled with
// -fmad=false -prec-div=true -prec-sqrt=true -ftz=false so each C operator is one IEEE binary32 operation,
// exactly like the reference's -ffp-contract=off host build (reference CMakeLists.txt:50-63).
// min/max/clamp are explicit ternaries, NOT fminf/fmaxf: the reference's scalar helpers
// (math_functions.h:170-191) and its SSE2 MINPS/MAXPS wrappers (src/contact_solver.c:858-866) both
// return the SECOND operand on equality or NaN.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define B2G_DEV __device__ __forceinline__

namespace b2g
{

struct V2
{
	float x, y;
};

struct Rot
{
	float c, s;
};

struct Soft
{
	float biasRate, massScale, impulseScale;
};

constexpr float kPi = 3.14159265359f; // B2_PI, math_functions.h:18
constexpr float kHugeFloatMax = 3.402823466e+38F; // FLT_MAX
constexpr float kFltMin = 1.175494351e-38F;		  // FLT_MIN

B2G_DEV float minf_( float a, float b )
{
	return a < b ? a : b; // b2MinFloat :171 / MINPS
}

B2G_DEV float maxf_( float a, float b )
{
	return a > b ? a : b; // b2MaxFloat :177 / MAXPS
}

B2G_DEV float absf_( float a )
{
	return a < 0 ? -a : a; // b2AbsFloat :183
}

B2G_DEV float clampf_( float a, float lo, float hi )
{
	return a < lo ? lo : ( a > hi ? hi : a ); // b2ClampFloat :189
}

B2G_DEV V2 v2( float x, float y )
{
	V2 r;
	r.x = x;
	r.y = y;
	return r;
}

B2G_DEV float dot( V2 a, V2 b )
{
	return a.x * b.x + a.y * b.y; // b2Dot :204
}

B2G_DEV float cross( V2 a, V2 b )
{
	return a.x * b.y - a.y * b.x; // b2Cross :210
}

B2G_DEV V2 crossSV( float s, V2 v )
{
	return v2( -s * v.y, s * v.x ); // b2CrossSV :222
}

B2G_DEV V2 leftPerp( V2 v )
{
	return v2( -v.y, v.x ); // b2LeftPerp :228
}

B2G_DEV V2 rightPerp( V2 v )
{
	return v2( v.y, -v.x ); // b2RightPerp :234
}

B2G_DEV V2 add( V2 a, V2 b )
{
	return v2( a.x + b.x, a.y + b.y );
}

B2G_DEV V2 sub( V2 a, V2 b )
{
	return v2( a.x - b.x, a.y - b.y );
}

B2G_DEV V2 neg( V2 a )
{
	return v2( -a.x, -a.y );
}

B2G_DEV V2 mulSV( float s, V2 v )
{
	return v2( s * v.x, s * v.y ); // b2MulSV :271
}

B2G_DEV V2 mulAdd( V2 a, float s, V2 b )
{
	return v2( a.x + s * b.x, a.y + s * b.y ); // b2MulAdd :277
}

B2G_DEV V2 mulSub( V2 a, float s, V2 b )
{
	return v2( a.x - s * b.x, a.y - s * b.y ); // b2MulSub :283
}

B2G_DEV float length( V2 v )
{
	return sqrtf( v.x * v.x + v.y * v.y ); // b2Length :325
}

B2G_DEV float lengthSquared( V2 v )
{
	return v.x * v.x + v.y * v.y; // b2LengthSquared :402
}

B2G_DEV V2 normalize( V2 a )
{
	// b2Normalize :340
	float lengthSq = a.x * a.x + a.y * a.y;
	if ( lengthSq > 1000.0f * kFltMin )
	{
		float s = 1.0f / sqrtf( lengthSq );
		return v2( s * a.x, s * a.y );
	}
	return v2( 0.0f, 0.0f );
}

B2G_DEV V2 rotate( Rot q, V2 v )
{
	return v2( q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y ); // b2RotateVector :558
}

B2G_DEV Rot mulRot( Rot q, Rot r )
{
	// b2MulRot :503
	Rot qr;
	qr.s = q.s * r.c + q.c * r.s;
	qr.c = q.c * r.c - q.s * r.s;
	return qr;
}

B2G_DEV Rot invMulRot( Rot a, Rot b )
{
	// b2InvMulRot :517
	Rot r;
	r.s = a.c * b.s - a.s * b.c;
	r.c = a.c * b.c + a.s * b.s;
	return r;
}

B2G_DEV Rot integrateRotation( Rot q1, float deltaAngle )
{
	// b2IntegrateRotation :388
	Rot q2;
	q2.c = q1.c - deltaAngle * q1.s;
	q2.s = q1.s + deltaAngle * q1.c;
	float mag = sqrtf( q2.s * q2.s + q2.c * q2.c );
	float invMag = mag > 0.0f ? 1.0f / mag : 0.0f;
	Rot qn;
	qn.c = q2.c * invMag;
	qn.s = q2.s * invMag;
	return qn;
}

B2G_DEV float atan2_( float y, float x )
{
	// b2Atan2, reference src/math_functions.c:96-136 (fixed minimax polynomial)
	if ( x == 0.0f && y == 0.0f )
	{
		return 0.0f;
	}
	float ax = absf_( x );
	float ay = absf_( y );
	float mx = maxf_( ay, ax );
	float mn = minf_( ay, ax );
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

B2G_DEV float rotAngle( Rot q )
{
	return atan2_( q.s, q.c ); // b2Rot_GetAngle :483
}

B2G_DEV float unwindAngle( float radians )
{
	// b2UnwindAngle :540 -- the rounding trick runs in binary64
	float x = clampf_( radians, -1.0e6f, 1.0e6f );
	double twoPi = (double)( 2.0f * kPi );
	double roundToNearest = 6755399441055744.0;
	double a = (double)x;
	double k = __dsub_rn( __dadd_rn( __ddiv_rn( a, twoPi ), roundToNearest ), roundToNearest );
	return (float)__dsub_rn( a, __dmul_rn( k, twoPi ) );
}

B2G_DEV V2 solve22( float a11, float a12, float a21, float a22, V2 b )
{
	// b2Solve22 :748, A = [cx cy] with cx=(a11,a21), cy=(a12,a22)
	float det = a11 * a22 - a12 * a21;
	if ( det != 0.0f )
	{
		det = 1.0f / det;
	}
	return v2( det * ( a22 * b.x - a12 * b.y ), det * ( a11 * b.y - a21 * b.x ) );
}

B2G_DEV Soft makeSoft( float hertz, float zeta, float h )
{
	// b2MakeSoft, reference src/solver.h:239-281
	Soft r;
	if ( hertz == 0.0f )
	{
		r.biasRate = 0.0f;
		r.massScale = 0.0f;
		r.impulseScale = 0.0f;
		return r;
	}
	float omega = 2.0f * kPi * hertz;
	float a1 = 2.0f * zeta + h * omega;
	float a2 = h * omega * a1;
	float a3 = 1.0f / ( 1.0f + a2 );
	r.biasRate = omega / a1;
	r.massScale = a2 * a3;
	r.impulseScale = a3;
	return r;
}

B2G_DEV float springDamper( float hertz, float dampingRatio, float position, float velocity, float timeStep )
{
	// b2SpringDamper :832
	float omega = 2.0f * kPi * hertz;
	float omegaH = omega * timeStep;
	return ( velocity - omega * omegaH * position ) / ( 1.0f + 2.0f * dampingRatio * omegaH + omegaH * omegaH );
}

} // namespace b2g
