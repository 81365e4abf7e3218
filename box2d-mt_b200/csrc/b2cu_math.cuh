// b2cu_math.cuh -- fp32 vector/rotation/transform math and tuning constants for the device kernels.
//
// Restates Box2D/Common/b2Math.h (b2Vec2 :36-125, b2Rot :281-331, b2Mul/b2MulT :404-570) and the constants of
// Box2D/Common/b2Settings.h:48-135.  The whole library is compiled with -fmad=false and the default
// -prec-div=true -prec-sqrt=true, and every expression keeps the reference's operand order, so each fp32
// result is the same IEEE-754 value the reference computes on x86-64 (SSE, no FMA).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define B2CU_MAX_FLOAT 3.402823466e+38F
#define B2CU_EPSILON 1.192092896e-07F
#define B2CU_PI 3.14159265359f

#define B2CU_MAX_MANIFOLD_POINTS 2
#define B2CU_MAX_POLY_VERTS 8
#define B2CU_AABB_EXTENSION 0.1f
#define B2CU_AABB_MULTIPLIER 2.0f
#define B2CU_LINEAR_SLOP 0.005f
#define B2CU_ANGULAR_SLOP (2.0f / 180.0f * B2CU_PI)
#define B2CU_POLYGON_RADIUS (2.0f * B2CU_LINEAR_SLOP)
#define B2CU_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * B2CU_PI)
#define B2CU_MAX_SUB_STEPS 8
#define B2CU_VELOCITY_THRESHOLD 1.0f
#define B2CU_MAX_LINEAR_CORRECTION 0.2f
#define B2CU_MAX_TRANSLATION 2.0f
#define B2CU_MAX_TRANSLATION_SQUARED (B2CU_MAX_TRANSLATION * B2CU_MAX_TRANSLATION)
#define B2CU_MAX_ROTATION (0.5f * B2CU_PI)
#define B2CU_MAX_ROTATION_SQUARED (B2CU_MAX_ROTATION * B2CU_MAX_ROTATION)
#define B2CU_BAUMGARTE 0.2f
#define B2CU_TIME_TO_SLEEP 0.5f
#define B2CU_LINEAR_SLEEP_TOLERANCE 0.01f
#define B2CU_ANGULAR_SLEEP_TOLERANCE (2.0f / 180.0f * B2CU_PI)

namespace b2cu
{

struct Vec2
{
	float x, y;
};

struct Rot
{
	float s, c;
};

struct Xf
{
	Vec2 p;
	Rot q;
};

__host__ __device__ __forceinline__ Vec2 V(float x, float y)
{
	Vec2 v;
	v.x = x;
	v.y = y;
	return v;
}

__host__ __device__ __forceinline__ Vec2 operator+(Vec2 a, Vec2 b) { return V(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ Vec2 operator-(Vec2 a, Vec2 b) { return V(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ Vec2 operator-(Vec2 a) { return V(-a.x, -a.y); }
__host__ __device__ __forceinline__ Vec2 operator*(float s, Vec2 a) { return V(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ float Dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
__host__ __device__ __forceinline__ float Cross(Vec2 a, Vec2 b) { return a.x * b.y - a.y * b.x; }
// cross(v, s) and cross(s, v), b2Math.h:398-409
__host__ __device__ __forceinline__ Vec2 CrossVS(Vec2 a, float s) { return V(s * a.y, -s * a.x); }
__host__ __device__ __forceinline__ Vec2 CrossSV(float s, Vec2 a) { return V(-s * a.y, s * a.x); }
__host__ __device__ __forceinline__ float DistanceSquared(Vec2 a, Vec2 b)
{
	Vec2 c = a - b;
	return Dot(c, c);
}
__host__ __device__ __forceinline__ float Length(Vec2 a) { return sqrtf(a.x * a.x + a.y * a.y); }

// b2Vec2::Normalize, b2Math.h:96-108
__host__ __device__ __forceinline__ Vec2 Normalized(Vec2 v)
{
	float length = Length(v);
	if (length < B2CU_EPSILON)
	{
		return v;
	}
	float invLength = 1.0f / length;
	return V(v.x * invLength, v.y * invLength);
}

__host__ __device__ __forceinline__ Vec2 Mul(Rot q, Vec2 v) { return V(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
__host__ __device__ __forceinline__ Vec2 MulT(Rot q, Vec2 v) { return V(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }

__host__ __device__ __forceinline__ Rot MulTRot(Rot q, Rot r)
{
	Rot qr;
	qr.s = q.c * r.s - q.s * r.c;
	qr.c = q.c * r.c + q.s * r.s;
	return qr;
}

__host__ __device__ __forceinline__ Vec2 Mul(const Xf& T, Vec2 v)
{
	float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
	float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
	return V(x, y);
}

__host__ __device__ __forceinline__ Vec2 MulT(const Xf& T, Vec2 v)
{
	float px = v.x - T.p.x;
	float py = v.y - T.p.y;
	float x = (T.q.c * px + T.q.s * py);
	float y = (-T.q.s * px + T.q.c * py);
	return V(x, y);
}

// b2MulT(b2Transform, b2Transform), b2Math.h:561-570
__host__ __device__ __forceinline__ Xf MulTXf(const Xf& A, const Xf& B)
{
	Xf C;
	C.q = MulTRot(A.q, B.q);
	C.p = MulT(A.q, B.p - A.p);
	return C;
}

__host__ __device__ __forceinline__ float Min(float a, float b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ float Max(float a, float b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ float Clamp(float a, float lo, float hi) { return Max(lo, Min(a, hi)); }
__host__ __device__ __forceinline__ float Abs(float a) { return a > 0.0f ? a : -a; }

__host__ __device__ __forceinline__ Xf MakeXf(float4 v)
{
	Xf t;
	t.p.x = v.x;
	t.p.y = v.y;
	t.q.s = v.z;
	t.q.c = v.w;
	return t;
}

#ifdef __CUDACC__
// fp32 sin/cos for b2Rot::Set (b2Math.h:289-299).  Same algorithm, operation for operation, as the oracle's
// oracle/b2o_math.c: Cody-Waite reduction and fdlibm kernel polynomials in binary64, rounded once to fp32.
// Only IEEE add/sub/mul and conversions are used (the *_rn intrinsics are never contracted into FMAs).
__device__ __forceinline__ Rot SinCos(float x)
{
	const double INV_PIO2 = __longlong_as_double(0x3FE45F306DC9C883ll);
	const double PIO2_1 = __longlong_as_double(0x3FF921FB54400000ll);
	const double PIO2_2 = __longlong_as_double(0x3DD0B4611A600000ll);
	const double PIO2_3 = __longlong_as_double(0x3BA3198A2E037073ll);
	const double MAGIC = 6755399441055744.0;
	const double S1 = __longlong_as_double(0xBFC5555555555549ll);
	const double S2 = __longlong_as_double(0x3F8111111110F8A6ll);
	const double S3 = __longlong_as_double(0xBF2A01A019C161D5ll);
	const double S4 = __longlong_as_double(0x3EC71DE357B1FE7Dll);
	const double S5 = __longlong_as_double(0xBE5AE5E68A2B9CEBll);
	const double S6 = __longlong_as_double(0x3DE5D93A5ACFD57Cll);
	const double C1 = __longlong_as_double(0x3FA555555555554Cll);
	const double C2 = __longlong_as_double(0xBF56C16C16C15177ll);
	const double C3 = __longlong_as_double(0x3EFA01A019CB1590ll);
	const double C4 = __longlong_as_double(0xBE927E4F809C52ADll);
	const double C5 = __longlong_as_double(0x3E21EE9EBDB4B1C4ll);
	const double C6 = __longlong_as_double(0xBDA8FAE9BE8838D4ll);

	double xd = (double)x;
	double t = __dadd_rn(__dmul_rn(xd, INV_PIO2), MAGIC);
	double k = __dsub_rn(t, MAGIC);
	long long n = (long long)k;

	double r = __dsub_rn(xd, __dmul_rn(k, PIO2_1));
	r = __dsub_rn(r, __dmul_rn(k, PIO2_2));
	r = __dsub_rn(r, __dmul_rn(k, PIO2_3));

	double z = __dmul_rn(r, r);

	double ps = __dadd_rn(S5, __dmul_rn(z, S6));
	ps = __dadd_rn(S4, __dmul_rn(z, ps));
	ps = __dadd_rn(S3, __dmul_rn(z, ps));
	ps = __dadd_rn(S2, __dmul_rn(z, ps));
	ps = __dadd_rn(S1, __dmul_rn(z, ps));
	double sr = __dadd_rn(r, __dmul_rn(__dmul_rn(r, z), ps));

	double pc = __dadd_rn(C5, __dmul_rn(z, C6));
	pc = __dadd_rn(C4, __dmul_rn(z, pc));
	pc = __dadd_rn(C3, __dmul_rn(z, pc));
	pc = __dadd_rn(C2, __dmul_rn(z, pc));
	pc = __dadd_rn(C1, __dmul_rn(z, pc));
	double cr = __dadd_rn(__dsub_rn(1.0, __dmul_rn(0.5, z)), __dmul_rn(__dmul_rn(z, z), pc));

	double s, c;
	switch ((int)(n & 3))
	{
	case 0: s = sr; c = cr; break;
	case 1: s = cr; c = -sr; break;
	case 2: s = -sr; c = -cr; break;
	default: s = -cr; c = sr; break;
	}

	Rot q;
	q.s = __double2float_rn(s);
	q.c = __double2float_rn(c);
	return q;
}
#endif

} // namespace b2cu
