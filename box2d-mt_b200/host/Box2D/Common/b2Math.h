// Host-side fp32 vector math with the reference's public names (reference: Box2D/Common/b2Math.h).
// Only what the host API and user code need: b2Vec2, b2Rot, b2Transform, b2Sweep, b2Mat22 and the free
// functions over them.  Expressions keep the reference's operand order so that host-computed quantities (mass
// data, AABBs, hull normals) are bit-identical to the reference's.  b2Rot uses the same correctly-rounded
// sincos as the device (b2SinCos), not libm.
#ifndef B2_MATH_H
#define B2_MATH_H

#include <cmath>
#include "Box2D/Common/b2Settings.h"

inline bool b2IsValid(float32 x) { return std::isfinite(x); }
#define b2Sqrt(x) sqrtf(x)
#define b2Atan2(y, x) atan2f(y, x)

/// sin and cos of an fp32 angle, identical on host and device (see b2cu_math.cuh SinCos)
void b2SinCos(float32 angle, float32* s, float32* c);

struct b2Vec2
{
	b2Vec2() {}
	b2Vec2(float32 xIn, float32 yIn) : x(xIn), y(yIn) {}
	void SetZero() { x = 0.0f; y = 0.0f; }
	void Set(float32 x_, float32 y_) { x = x_; y = y_; }
	b2Vec2 operator-() const { return b2Vec2(-x, -y); }
	float32 operator()(int32 i) const { return (&x)[i]; }
	float32& operator()(int32 i) { return (&x)[i]; }
	void operator+=(const b2Vec2& v) { x += v.x; y += v.y; }
	void operator-=(const b2Vec2& v) { x -= v.x; y -= v.y; }
	void operator*=(float32 a) { x *= a; y *= a; }
	float32 Length() const { return b2Sqrt(x * x + y * y); }
	float32 LengthSquared() const { return x * x + y * y; }
	float32 Normalize()
	{
		float32 length = Length();
		if (length < b2_epsilon) return 0.0f;
		float32 inv = 1.0f / length;
		x *= inv;
		y *= inv;
		return length;
	}
	bool IsValid() const { return b2IsValid(x) && b2IsValid(y); }
	b2Vec2 Skew() const { return b2Vec2(-y, x); }
	float32 x, y;
};

struct b2Vec3
{
	b2Vec3() {}
	b2Vec3(float32 xIn, float32 yIn, float32 zIn) : x(xIn), y(yIn), z(zIn) {}
	void SetZero() { x = y = z = 0.0f; }
	void Set(float32 x_, float32 y_, float32 z_) { x = x_; y = y_; z = z_; }
	float32 x, y, z;
};

struct b2Mat22
{
	b2Mat22() {}
	b2Mat22(const b2Vec2& c1, const b2Vec2& c2) : ex(c1), ey(c2) {}
	void SetZero() { ex.SetZero(); ey.SetZero(); }
	void SetIdentity() { ex.Set(1.0f, 0.0f); ey.Set(0.0f, 1.0f); }
	b2Vec2 ex, ey;
};

struct b2Rot
{
	b2Rot() {}
	explicit b2Rot(float32 angle) { b2SinCos(angle, &s, &c); }
	void Set(float32 angle) { b2SinCos(angle, &s, &c); }
	void SetIdentity() { s = 0.0f; c = 1.0f; }
	float32 GetAngle() const { return b2Atan2(s, c); }
	b2Vec2 GetXAxis() const { return b2Vec2(c, s); }
	b2Vec2 GetYAxis() const { return b2Vec2(-s, c); }
	float32 s, c;
};

struct b2Transform
{
	b2Transform() {}
	b2Transform(const b2Vec2& position, const b2Rot& rotation) : p(position), q(rotation) {}
	void SetIdentity() { p.SetZero(); q.SetIdentity(); }
	void Set(const b2Vec2& position, float32 angle) { p = position; q.Set(angle); }
	b2Vec2 p;
	b2Rot q;
};

struct b2Sweep
{
	void GetTransform(b2Transform* xf, float32 beta) const;
	b2Vec2 localCenter;
	b2Vec2 c0, c;
	float32 a0, a;
	float32 alpha0;
};

extern const b2Vec2 b2Vec2_zero;

inline float32 b2Dot(const b2Vec2& a, const b2Vec2& b) { return a.x * b.x + a.y * b.y; }
inline float32 b2Cross(const b2Vec2& a, const b2Vec2& b) { return a.x * b.y - a.y * b.x; }
inline b2Vec2 b2Cross(const b2Vec2& a, float32 s) { return b2Vec2(s * a.y, -s * a.x); }
inline b2Vec2 b2Cross(float32 s, const b2Vec2& a) { return b2Vec2(-s * a.y, s * a.x); }
inline b2Vec2 operator+(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(a.x + b.x, a.y + b.y); }
inline b2Vec2 operator-(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(a.x - b.x, a.y - b.y); }
inline b2Vec2 operator*(float32 s, const b2Vec2& a) { return b2Vec2(s * a.x, s * a.y); }
inline bool operator==(const b2Vec2& a, const b2Vec2& b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(const b2Vec2& a, const b2Vec2& b) { return a.x != b.x || a.y != b.y; }
inline float32 b2Distance(const b2Vec2& a, const b2Vec2& b) { return (a - b).Length(); }
inline float32 b2DistanceSquared(const b2Vec2& a, const b2Vec2& b)
{
	b2Vec2 c = a - b;
	return b2Dot(c, c);
}
inline b2Vec2 b2Mul(const b2Rot& q, const b2Vec2& v) { return b2Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline b2Vec2 b2MulT(const b2Rot& q, const b2Vec2& v) { return b2Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
inline b2Vec2 b2Mul(const b2Transform& T, const b2Vec2& v)
{
	float32 x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
	float32 y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
	return b2Vec2(x, y);
}
inline b2Vec2 b2MulT(const b2Transform& T, const b2Vec2& v)
{
	float32 px = v.x - T.p.x;
	float32 py = v.y - T.p.y;
	return b2Vec2(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
inline b2Vec2 b2Mul(const b2Mat22& A, const b2Vec2& v)
{
	return b2Vec2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y);
}

template <typename T> inline T b2Abs(T a) { return a > T(0) ? a : -a; }
template <typename T> inline T b2Min(T a, T b) { return a < b ? a : b; }
template <typename T> inline T b2Max(T a, T b) { return a > b ? a : b; }
template <typename T> inline T b2Clamp(T a, T low, T high) { return b2Max(low, b2Min(a, high)); }
template <typename T> inline void b2Swap(T& a, T& b) { T t = a; a = b; b = t; }
inline b2Vec2 b2Min(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(b2Min(a.x, b.x), b2Min(a.y, b.y)); }
inline b2Vec2 b2Max(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(b2Max(a.x, b.x), b2Max(a.y, b.y)); }
inline b2Vec2 b2Abs(const b2Vec2& a) { return b2Vec2(b2Abs(a.x), b2Abs(a.y)); }

inline void b2Sweep::GetTransform(b2Transform* xf, float32 beta) const
{
	xf->p = (1.0f - beta) * c0 + beta * c;
	float32 angle = (1.0f - beta) * a0 + beta * a;
	xf->q.Set(angle);
	xf->p -= b2Mul(xf->q, localCenter);
}

#endif
