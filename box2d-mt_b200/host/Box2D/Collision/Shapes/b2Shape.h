// Shape base class (reference: Box2D/Collision/Shapes/b2Shape.h:26-104).  Shapes are host-side value objects:
// user code builds them, b2Body::CreateFixture clones them, and their geometry is uploaded to the device
// shape table (a chain as one edge record per segment).
#ifndef B2_SHAPE_H
#define B2_SHAPE_H

#include "Box2D/Collision/b2Collision.h"

/// result of b2Shape::ComputeMass: mass, centroid in shape coordinates, rotational inertia about the shape origin
struct b2MassData
{
	float32 mass, I;
	b2Vec2 center;
};

class b2Shape
{
public:
	enum Type { e_circle = 0, e_edge = 1, e_polygon = 2, e_chain = 3, e_typeCount = 4 };

	Type m_type;
	float32 m_radius; // circle radius, or the skin of polygons / edges (b2_polygonRadius)

	Type GetType() const { return m_type; }

	// geometry queries every concrete shape implements
	virtual int32 GetChildCount() const = 0;
	virtual void ComputeAABB(b2AABB* aabb, const b2Transform& xf, int32 childIndex) const = 0;
	virtual void ComputeMass(b2MassData* massData, float32 density) const = 0;
	virtual bool TestPoint(const b2Transform& xf, const b2Vec2& p) const = 0;
	virtual bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf, int32 childIndex) const = 0;

	/// heap copy owned by the caller (the reference clones into its block allocator)
	virtual b2Shape* Clone() const = 0;
	virtual ~b2Shape() {}
};

#endif
