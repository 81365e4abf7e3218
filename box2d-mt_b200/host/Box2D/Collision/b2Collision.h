// Collision data types seen by user code (reference: Box2D/Collision/b2Collision.h:38-286): contact feature
// ids, manifolds, world manifolds, AABBs.  Layouts equal the reference's (b2Manifold is 64 bytes); the
// manifold evaluation itself runs on the device (box2d-mt_b200/csrc/b2cu_collide.cuh).
#ifndef B2_COLLISION_H
#define B2_COLLISION_H

#include "Box2D/Common/b2Math.h"

struct b2ContactFeature
{
	enum Type { e_vertex = 0, e_face = 1 };
	uint8 indexA, indexB, typeA, typeB;
};

union b2ContactID
{
	b2ContactFeature cf;
	uint32 key;
};

struct b2ManifoldPoint
{
	b2Vec2 localPoint;
	float32 normalImpulse;
	float32 tangentImpulse;
	b2ContactID id;
};

struct b2Manifold
{
	enum Type { e_circles, e_faceA, e_faceB };
	b2ManifoldPoint points[b2_maxManifoldPoints];
	b2Vec2 localNormal;
	b2Vec2 localPoint;
	Type type;
	int32 pointCount;
};

struct b2WorldManifold
{
	/// Same evaluation as the solver's (reference: Box2D/Collision/b2Collision.cpp:22-86)
	void Initialize(const b2Manifold* manifold, const b2Transform& xfA, float32 radiusA, const b2Transform& xfB,
	                float32 radiusB);
	b2Vec2 normal;
	b2Vec2 points[b2_maxManifoldPoints];
	float32 separations[b2_maxManifoldPoints];
};

/// segment p1 -> p1 + maxFraction * (p2 - p1), and where / with which surface normal it hits
struct b2RayCastInput
{
	b2Vec2 p1, p2;
	float32 maxFraction;
};
struct b2RayCastOutput
{
	b2Vec2 normal;
	float32 fraction;
};

struct b2AABB
{
	bool IsValid() const
	{
		b2Vec2 d = upperBound - lowerBound;
		return d.x >= 0.0f && d.y >= 0.0f && lowerBound.IsValid() && upperBound.IsValid();
	}
	b2Vec2 GetCenter() const { return 0.5f * (lowerBound + upperBound); }
	b2Vec2 GetExtents() const { return 0.5f * (upperBound - lowerBound); }
	float32 GetPerimeter() const { return 2.0f * ((upperBound.x - lowerBound.x) + (upperBound.y - lowerBound.y)); }
	void Combine(const b2AABB& a)
	{
		lowerBound = b2Min(lowerBound, a.lowerBound);
		upperBound = b2Max(upperBound, a.upperBound);
	}
	void Combine(const b2AABB& a, const b2AABB& b)
	{
		lowerBound = b2Min(a.lowerBound, b.lowerBound);
		upperBound = b2Max(a.upperBound, b.upperBound);
	}
	bool Contains(const b2AABB& a) const
	{
		return lowerBound.x <= a.lowerBound.x && lowerBound.y <= a.lowerBound.y && a.upperBound.x <= upperBound.x &&
		       a.upperBound.y <= upperBound.y;
	}
	b2Vec2 lowerBound, upperBound;
};

inline bool b2TestOverlap(const b2AABB& a, const b2AABB& b)
{
	b2Vec2 d1 = b.lowerBound - a.upperBound;
	b2Vec2 d2 = a.lowerBound - b.upperBound;
	if (d1.x > 0.0f || d1.y > 0.0f) return false;
	if (d2.x > 0.0f || d2.y > 0.0f) return false;
	return true;
}

#endif
