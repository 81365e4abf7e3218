// Chain of line segments with two-sided collision (reference: Box2D/Collision/Shapes/b2ChainShape.h,
// b2ChainShape.cpp:23-189).  On the device a chain does not exist: every segment of a chain fixture becomes one
// proxy whose geometry record is the child edge with its two ghost vertices (b2EdgeShape layout), which is exactly
// how the reference collides chains (b2ChainAndCircleContact / b2ChainAndPolygonContact fetch the child edge and
// call the edge manifold functions).
#ifndef B2_CHAIN_SHAPE_H
#define B2_CHAIN_SHAPE_H

#include <vector>

#include "Box2D/Collision/Shapes/b2Shape.h"

class b2EdgeShape;

class b2ChainShape : public b2Shape
{
public:
	b2ChainShape() : m_hasPrevVertex(false), m_hasNextVertex(false)
	{
		m_type = e_chain;
		m_radius = b2_polygonRadius;
		m_prevVertex.SetZero();
		m_nextVertex.SetZero();
	}

	void Clear() { m_points.clear(); }
	/// closed loop: `count` vertices, the last one connects back to the first
	void CreateLoop(const b2Vec2* vertices, int32 count);
	/// open chain with isolated end vertices (see SetPrevVertex / SetNextVertex)
	void CreateChain(const b2Vec2* vertices, int32 count);
	/// ghost vertices that connect an open chain to a neighbouring shape
	void SetPrevVertex(const b2Vec2& prevVertex)
	{
		m_prevVertex = prevVertex;
		m_hasPrevVertex = true;
	}
	void SetNextVertex(const b2Vec2& nextVertex)
	{
		m_nextVertex = nextVertex;
		m_hasNextVertex = true;
	}
	/// segment `index` as an edge shape with its neighbours as ghost vertices
	void GetChildEdge(b2EdgeShape* edge, int32 index) const;

	b2Shape* Clone() const override { return new b2ChainShape(*this); }
	int32 GetChildCount() const override { return m_points.empty() ? 0 : (int32)m_points.size() - 1; }
	bool TestPoint(const b2Transform&, const b2Vec2&) const override { return false; }
	bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf, int32 childIndex) const override;
	void ComputeAABB(b2AABB* aabb, const b2Transform& xf, int32 childIndex) const override;
	void ComputeMass(b2MassData* massData, float32 density) const override;

	int32 GetVertexCount() const { return (int32)m_points.size(); }
	const b2Vec2& GetVertex(int32 index) const { return m_points[index]; }

	b2Vec2 m_prevVertex, m_nextVertex;
	bool m_hasPrevVertex, m_hasNextVertex;

private:
	std::vector<b2Vec2> m_points;
};

#endif
