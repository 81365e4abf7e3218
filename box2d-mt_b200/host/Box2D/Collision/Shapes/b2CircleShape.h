// reference: Box2D/Collision/Shapes/b2CircleShape.h, b2CircleShape.cpp:83-101
#ifndef B2_CIRCLE_SHAPE_H
#define B2_CIRCLE_SHAPE_H

#include "Box2D/Collision/Shapes/b2Shape.h"

class b2CircleShape : public b2Shape
{
public:
	b2CircleShape()
	{
		m_type = e_circle;
		m_radius = 0.0f;
		m_p.SetZero();
	}
	b2Shape* Clone() const override { return new b2CircleShape(*this); }
	int32 GetChildCount() const override { return 1; }
	bool TestPoint(const b2Transform& xf, const b2Vec2& p) const override;
	bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf, int32 childIndex) const override;
	void ComputeAABB(b2AABB* aabb, const b2Transform& xf, int32 childIndex) const override;
	void ComputeMass(b2MassData* massData, float32 density) const override;

	b2Vec2 m_p;
};

#endif
