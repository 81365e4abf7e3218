// Shape geometry on the host: hull construction, AABBs and mass data.  These values are uploaded to the device
// (shape table, initial fat AABBs, inverse mass / inertia), so each routine keeps the fp32 operation order of
// the reference function it replaces and the results are bit-identical:
//   circle   Box2D/Collision/Shapes/b2CircleShape.cpp:40-101
//   edge     Box2D/Collision/Shapes/b2EdgeShape.cpp:116-140
//   polygon  Box2D/Collision/Shapes/b2PolygonShape.cpp:28-440
#include "Box2D/Collision/Shapes/b2CircleShape.h"
#include "Box2D/Collision/Shapes/b2ChainShape.h"
#include "Box2D/Collision/Shapes/b2EdgeShape.h"
#include "Box2D/Collision/Shapes/b2PolygonShape.h"

// ---- circle ----------------------------------------------------------------------------------------------

bool b2CircleShape::TestPoint(const b2Transform& xf, const b2Vec2& p) const
{
	b2Vec2 center = xf.p + b2Mul(xf.q, m_p);
	b2Vec2 d = p - center;
	return b2Dot(d, d) <= m_radius * m_radius;
}

void b2CircleShape::ComputeAABB(b2AABB* aabb, const b2Transform& xf, int32) const
{
	b2Vec2 p = xf.p + b2Mul(xf.q, m_p);
	aabb->lowerBound.Set(p.x - m_radius, p.y - m_radius);
	aabb->upperBound.Set(p.x + m_radius, p.y + m_radius);
}

void b2CircleShape::ComputeMass(b2MassData* massData, float32 density) const
{
	massData->mass = density * b2_pi * m_radius * m_radius;
	massData->center = m_p;
	// about the local origin
	massData->I = massData->mass * (0.5f * m_radius * m_radius + b2Dot(m_p, m_p));
}

// ---- edge ------------------------------------------------------------------------------------------------

void b2EdgeShape::ComputeAABB(b2AABB* aabb, const b2Transform& xf, int32) const
{
	b2Vec2 v1 = b2Mul(xf, m_vertex1);
	b2Vec2 v2 = b2Mul(xf, m_vertex2);
	b2Vec2 lower = b2Min(v1, v2);
	b2Vec2 upper = b2Max(v1, v2);
	b2Vec2 r(m_radius, m_radius);
	aabb->lowerBound = lower - r;
	aabb->upperBound = upper + r;
}

void b2EdgeShape::ComputeMass(b2MassData* massData, float32) const
{
	massData->mass = 0.0f;
	massData->center = 0.5f * (m_vertex1 + m_vertex2);
	massData->I = 0.0f;
}

// ---- polygon ---------------------------------------------------------------------------------------------

void b2PolygonShape::SetAsBox(float32 hx, float32 hy)
{
	m_count = 4;
	m_vertices[0].Set(-hx, -hy);
	m_vertices[1].Set(hx, -hy);
	m_vertices[2].Set(hx, hy);
	m_vertices[3].Set(-hx, hy);
	m_normals[0].Set(0.0f, -1.0f);
	m_normals[1].Set(1.0f, 0.0f);
	m_normals[2].Set(0.0f, 1.0f);
	m_normals[3].Set(-1.0f, 0.0f);
	m_centroid.SetZero();
}

void b2PolygonShape::SetAsBox(float32 hx, float32 hy, const b2Vec2& center, float32 angle)
{
	SetAsBox(hx, hy);
	m_centroid = center;
	b2Transform xf;
	xf.p = center;
	xf.q.Set(angle);
	for (int32 i = 0; i < m_count; ++i)
	{
		m_vertices[i] = b2Mul(xf, m_vertices[i]);
		m_normals[i] = b2Mul(xf.q, m_normals[i]);
	}
}

// area-weighted centroid over the triangle fan about the origin (reference :74-118)
static b2Vec2 PolygonCentroid(const b2Vec2* vs, int32 count)
{
	b2Vec2 c(0.0f, 0.0f);
	float32 area = 0.0f;
	const b2Vec2 origin(0.0f, 0.0f);
	const float32 inv3 = 1.0f / 3.0f;
	for (int32 i = 0; i < count; ++i)
	{
		b2Vec2 p1 = origin;
		b2Vec2 p2 = vs[i];
		b2Vec2 p3 = i + 1 < count ? vs[i + 1] : vs[0];
		b2Vec2 e1 = p2 - p1;
		b2Vec2 e2 = p3 - p1;
		float32 D = b2Cross(e1, e2);
		float32 triangleArea = 0.5f * D;
		area += triangleArea;
		c += triangleArea * inv3 * (p1 + p2 + p3);
	}
	c *= 1.0f / area;
	return c;
}

void b2PolygonShape::Set(const b2Vec2* points, int32 count)
{
	if (count < 3)
	{
		SetAsBox(1.0f, 1.0f);
		return;
	}
	int32 n = b2Min(count, (int32)b2_maxPolygonVertices);

	// weld points closer than half the linear slop
	b2Vec2 ps[b2_maxPolygonVertices];
	int32 kept = 0;
	const float32 weldSqr = (0.5f * b2_linearSlop) * (0.5f * b2_linearSlop);
	for (int32 i = 0; i < n; ++i)
	{
		bool unique = true;
		for (int32 j = 0; j < kept; ++j)
		{
			if (b2DistanceSquared(points[i], ps[j]) < weldSqr)
			{
				unique = false;
				break;
			}
		}
		if (unique) ps[kept++] = points[i];
	}
	n = kept;
	if (n < 3)
	{
		SetAsBox(1.0f, 1.0f);
		return;
	}

	// gift wrapping from the right-most (then lowest) point, counter-clockwise
	int32 start = 0;
	for (int32 i = 1; i < n; ++i)
	{
		if (ps[i].x > ps[start].x || (ps[i].x == ps[start].x && ps[i].y < ps[start].y)) start = i;
	}
	int32 hull[b2_maxPolygonVertices];
	int32 m = 0;
	int32 current = start;
	for (;;)
	{
		hull[m] = current;
		int32 candidate = 0;
		for (int32 j = 1; j < n; ++j)
		{
			if (candidate == current)
			{
				candidate = j;
				continue;
			}
			b2Vec2 r = ps[candidate] - ps[hull[m]];
			b2Vec2 v = ps[j] - ps[hull[m]];
			float32 c = b2Cross(r, v);
			if (c < 0.0f) candidate = j;
			// collinear: keep the farther point
			if (c == 0.0f && v.LengthSquared() > r.LengthSquared()) candidate = j;
		}
		++m;
		current = candidate;
		if (candidate == start) break;
	}
	if (m < 3)
	{
		SetAsBox(1.0f, 1.0f);
		return;
	}

	m_count = m;
	for (int32 i = 0; i < m; ++i) m_vertices[i] = ps[hull[i]];
	for (int32 i = 0; i < m; ++i)
	{
		int32 i2 = i + 1 < m ? i + 1 : 0;
		b2Vec2 edge = m_vertices[i2] - m_vertices[i];
		m_normals[i] = b2Cross(edge, 1.0f);
		m_normals[i].Normalize();
	}
	m_centroid = PolygonCentroid(m_vertices, m);
}

bool b2PolygonShape::TestPoint(const b2Transform& xf, const b2Vec2& p) const
{
	b2Vec2 local = b2MulT(xf.q, p - xf.p);
	for (int32 i = 0; i < m_count; ++i)
	{
		if (b2Dot(m_normals[i], local - m_vertices[i]) > 0.0f) return false;
	}
	return true;
}

void b2PolygonShape::ComputeAABB(b2AABB* aabb, const b2Transform& xf, int32) const
{
	b2Vec2 lower = b2Mul(xf, m_vertices[0]);
	b2Vec2 upper = lower;
	for (int32 i = 1; i < m_count; ++i)
	{
		b2Vec2 v = b2Mul(xf, m_vertices[i]);
		lower = b2Min(lower, v);
		upper = b2Max(upper, v);
	}
	b2Vec2 r(m_radius, m_radius);
	aabb->lowerBound = lower - r;
	aabb->upperBound = upper + r;
}

// mass, centre and inertia by a triangle fan about the vertex average (reference :359-440)
void b2PolygonShape::ComputeMass(b2MassData* massData, float32 density) const
{
	b2Vec2 center(0.0f, 0.0f);
	float32 area = 0.0f;
	float32 I = 0.0f;

	b2Vec2 s(0.0f, 0.0f);
	for (int32 i = 0; i < m_count; ++i) s += m_vertices[i];
	s *= 1.0f / m_count;

	const float32 k_inv3 = 1.0f / 3.0f;
	for (int32 i = 0; i < m_count; ++i)
	{
		b2Vec2 e1 = m_vertices[i] - s;
		b2Vec2 e2 = i + 1 < m_count ? m_vertices[i + 1] - s : m_vertices[0] - s;
		float32 D = b2Cross(e1, e2);
		float32 triangleArea = 0.5f * D;
		area += triangleArea;
		center += triangleArea * k_inv3 * (e1 + e2);

		float32 ex1 = e1.x, ey1 = e1.y;
		float32 ex2 = e2.x, ey2 = e2.y;
		float32 intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
		float32 inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
		I += (0.25f * k_inv3 * D) * (intx2 + inty2);
	}

	massData->mass = density * area;
	center *= 1.0f / area;
	massData->center = center + s;
	massData->I = density * I;
	// shift from the fan origin s to the centre of mass, then to the body origin
	massData->I += massData->mass * (b2Dot(massData->center, massData->center) - b2Dot(center, center));
}

// ---- b2ChainShape (reference b2ChainShape.cpp) ------------------------------------------------------------

void b2ChainShape::CreateLoop(const b2Vec2* vertices, int32 count)
{
	b2Assert(m_points.empty() && count >= 3);
	m_points.assign(vertices, vertices + count);
	m_points.push_back(vertices[0]); // closing vertex
	// the loop's own neighbours are the ghost vertices of its first and last segment
	m_prevVertex = m_points[m_points.size() - 2];
	m_nextVertex = m_points[1];
	m_hasPrevVertex = m_hasNextVertex = true;
}

void b2ChainShape::CreateChain(const b2Vec2* vertices, int32 count)
{
	b2Assert(m_points.empty() && count >= 2);
	m_points.assign(vertices, vertices + count);
	m_hasPrevVertex = m_hasNextVertex = false;
	m_prevVertex.SetZero();
	m_nextVertex.SetZero();
}

void b2ChainShape::GetChildEdge(b2EdgeShape* edge, int32 index) const
{
	const int32 last = (int32)m_points.size() - 2; // index of the last segment
	b2Assert(0 <= index && index <= last);
	edge->m_type = b2Shape::e_edge;
	edge->m_radius = m_radius;
	edge->m_vertex1 = m_points[index];
	edge->m_vertex2 = m_points[index + 1];
	edge->m_hasVertex0 = index > 0 ? true : m_hasPrevVertex;
	edge->m_vertex0 = index > 0 ? m_points[index - 1] : m_prevVertex;
	edge->m_hasVertex3 = index < last ? true : m_hasNextVertex;
	edge->m_vertex3 = index < last ? m_points[index + 2] : m_nextVertex;
}

void b2ChainShape::ComputeAABB(b2AABB* aabb, const b2Transform& xf, int32 childIndex) const
{
	b2Assert(childIndex + 1 < (int32)m_points.size());
	int32 i2 = childIndex + 1;
	if (i2 == (int32)m_points.size()) i2 = 0;
	b2Vec2 v1 = b2Mul(xf, m_points[childIndex]);
	b2Vec2 v2 = b2Mul(xf, m_points[i2]);
	aabb->lowerBound = b2Min(v1, v2);
	aabb->upperBound = b2Max(v1, v2);
}

void b2ChainShape::ComputeMass(b2MassData* massData, float32) const
{
	massData->mass = 0.0f;
	massData->center.SetZero();
	massData->I = 0.0f;
}

// ---- ray casts (reference b2CircleShape.cpp:46-81, b2EdgeShape.cpp:54-114, b2PolygonShape.cpp:268-338) -------------

bool b2CircleShape::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf, int32) const
{
	// |s + a r|^2 = radius^2 with s from the centre to p1 and r = p2 - p1; smaller root
	b2Vec2 centre = xf.p + b2Mul(xf.q, m_p);
	b2Vec2 s = input.p1 - centre;
	float32 b = b2Dot(s, s) - m_radius * m_radius;
	b2Vec2 r = input.p2 - input.p1;
	float32 c = b2Dot(s, r);
	float32 rr = b2Dot(r, r);
	float32 sigma = c * c - rr * b;
	if (sigma < 0.0f || rr < b2_epsilon) return false;
	float32 a = -(c + b2Sqrt(sigma));
	if (0.0f <= a && a <= input.maxFraction * rr)
	{
		a /= rr;
		output->fraction = a;
		output->normal = s + a * r;
		output->normal.Normalize();
		return true;
	}
	return false;
}

bool b2EdgeShape::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf, int32) const
{
	// in the edge's frame: hit the carrier line, then check that the hit lies between the end points
	b2Vec2 p1 = b2MulT(xf.q, input.p1 - xf.p);
	b2Vec2 p2 = b2MulT(xf.q, input.p2 - xf.p);
	b2Vec2 d = p2 - p1;
	b2Vec2 e = m_vertex2 - m_vertex1;
	b2Vec2 normal(e.y, -e.x);
	normal.Normalize();
	float32 numerator = b2Dot(normal, m_vertex1 - p1);
	float32 denominator = b2Dot(normal, d);
	if (denominator == 0.0f) return false;
	float32 t = numerator / denominator;
	if (t < 0.0f || input.maxFraction < t) return false;
	b2Vec2 q = p1 + t * d;
	float32 rr = b2Dot(e, e);
	if (rr == 0.0f) return false;
	float32 along = b2Dot(q - m_vertex1, e) / rr;
	if (along < 0.0f || 1.0f < along) return false;
	output->fraction = t;
	output->normal = numerator > 0.0f ? -b2Mul(xf.q, normal) : b2Mul(xf.q, normal);
	return true;
}

bool b2PolygonShape::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf, int32) const
{
	// clip the parameter interval [0, maxFraction] against every face half-plane, in the polygon's frame
	b2Vec2 p1 = b2MulT(xf.q, input.p1 - xf.p);
	b2Vec2 p2 = b2MulT(xf.q, input.p2 - xf.p);
	b2Vec2 d = p2 - p1;
	float32 lower = 0.0f, upper = input.maxFraction;
	int32 entryFace = -1;
	for (int32 i = 0; i < m_count; ++i)
	{
		float32 numerator = b2Dot(m_normals[i], m_vertices[i] - p1);
		float32 denominator = b2Dot(m_normals[i], d);
		if (denominator == 0.0f)
		{
			if (numerator < 0.0f) return false; // parallel and outside
		}
		else if (denominator < 0.0f && numerator < lower * denominator)
		{
			lower = numerator / denominator; // entering through this face
			entryFace = i;
		}
		else if (denominator > 0.0f && numerator < upper * denominator)
		{
			upper = numerator / denominator; // leaving through this face
		}
		if (upper < lower) return false;
	}
	if (entryFace < 0) return false;
	output->fraction = lower;
	output->normal = b2Mul(xf.q, m_normals[entryFace]);
	return true;
}

bool b2ChainShape::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf, int32 childIndex) const
{
	// the segment as a plain edge (reference b2ChainShape.cpp:153-171)
	b2EdgeShape edge;
	int32 i2 = childIndex + 1;
	if (i2 == (int32)m_points.size()) i2 = 0;
	edge.m_vertex1 = m_points[childIndex];
	edge.m_vertex2 = m_points[i2];
	return edge.RayCast(output, input, xf, 0);
}

