// b2cu_collide.cuh -- device manifold functions (one thread evaluates one shape pair).
//
// Restates the reference narrow-phase, keeping every fp32 operation in the reference's order so that point
// counts, feature ids and manifold geometry come out bit-identical:
//   CollidePolygons          Box2D/Collision/b2CollidePolygon.cpp:23-239
//   CollideCircles           Box2D/Collision/b2CollideCircle.cpp:23-49
//   CollidePolygonAndCircle  Box2D/Collision/b2CollideCircle.cpp:51-154
//   CollideEdgeAndCircle     Box2D/Collision/b2CollideEdge.cpp:27-152
//   CollideEdgeAndPolygon    Box2D/Collision/b2CollideEdge.cpp:230-698 (b2EPCollider)
//   ClipSegmentToLine        Box2D/Collision/b2Collision.cpp:201-231
// Shape geometry is read straight from the (small, L1/L2-resident) b2cuShape table with read-only loads; no
// per-thread vertex arrays are kept, so nothing spills to local memory.
#pragma once

#include "b2cu_math.cuh"
#include "../../include/b2cuda.h"

namespace b2cu
{

struct Manifold
{
	Vec2 localNormal;
	Vec2 localPoint;
	Vec2 lp[2];
	float ni[2];
	float ti[2];
	uint32_t id[2];
	int type;
	int pointCount;
};

struct ClipVertex
{
	Vec2 v;
	uint32_t id;
};

// b2ContactFeature, Box2D/Collision/b2Collision.h:38-52
#define B2CU_CF_VERTEX 0u
#define B2CU_CF_FACE 1u
__device__ __forceinline__ uint32_t MakeId(uint32_t indexA, uint32_t indexB, uint32_t typeA, uint32_t typeB)
{
	return (indexA & 0xFFu) | ((indexB & 0xFFu) << 8) | (typeA << 16) | (typeB << 24);
}
__device__ __forceinline__ uint32_t FlipId(uint32_t id)
{
	return ((id >> 8) & 0xFFu) | ((id & 0xFFu) << 8) | (((id >> 24) & 0xFFu) << 16) | (((id >> 16) & 0xFFu) << 24);
}

__device__ __forceinline__ Vec2 ShapeV(const b2cuShape* __restrict__ s, int i)
{
	float2 t = __ldg(reinterpret_cast<const float2*>(&s->v[i][0]));
	return V(t.x, t.y);
}
__device__ __forceinline__ Vec2 ShapeN(const b2cuShape* __restrict__ s, int i)
{
	float2 t = __ldg(reinterpret_cast<const float2*>(&s->n[i][0]));
	return V(t.x, t.y);
}

__device__ __forceinline__ int ClipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], Vec2 normal, float offset,
                                                 int vertexIndexA)
{
	int numOut = 0;
	float distance0 = Dot(normal, vIn[0].v) - offset;
	float distance1 = Dot(normal, vIn[1].v) - offset;

	if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
	if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];

	if (distance0 * distance1 < 0.0f)
	{
		float interp = distance0 / (distance0 - distance1);
		vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
		vOut[numOut].id = MakeId((uint32_t)vertexIndexA, (vIn[0].id >> 8) & 0xFFu, B2CU_CF_VERTEX, B2CU_CF_FACE);
		++numOut;
	}
	return numOut;
}

__device__ __forceinline__ float FindMaxSeparation(int* edgeIndex, const b2cuShape* __restrict__ poly1, const Xf& xf1,
                                                   const b2cuShape* __restrict__ poly2, const Xf& xf2)
{
	int count1 = poly1->count;
	int count2 = poly2->count;
	Xf xf = MulTXf(xf2, xf1);

	// the vertices of poly2 are read once into registers: the n1 x n2 loop is then pure arithmetic
	Vec2 v2s[B2CU_MAX_POLYGON_VERTICES];
#pragma unroll
	for (int j = 0; j < B2CU_MAX_POLYGON_VERTICES; ++j) v2s[j] = ShapeV(poly2, j);

	int bestIndex = 0;
	float maxSeparation = -B2CU_MAX_FLOAT;
	for (int i = 0; i < count1; ++i)
	{
		Vec2 n = Mul(xf.q, ShapeN(poly1, i));
		Vec2 v1 = Mul(xf, ShapeV(poly1, i));

		float si = B2CU_MAX_FLOAT;
#pragma unroll
		for (int j = 0; j < B2CU_MAX_POLYGON_VERTICES; ++j)
		{
			if (j < count2)
			{
				float sij = Dot(n, v2s[j] - v1);
				if (sij < si)
				{
					si = sij;
				}
			}
		}

		if (si > maxSeparation)
		{
			maxSeparation = si;
			bestIndex = i;
		}
	}

	*edgeIndex = bestIndex;
	return maxSeparation;
}

__device__ __forceinline__ void CollidePolygons(Manifold* m, const b2cuShape* __restrict__ polyA, const Xf& xfA,
                                                const b2cuShape* __restrict__ polyB, const Xf& xfB)
{
	m->pointCount = 0;
	float totalRadius = polyA->radius + polyB->radius;

	int edgeA = 0;
	float separationA = FindMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
	if (separationA > totalRadius) return;

	int edgeB = 0;
	float separationB = FindMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
	if (separationB > totalRadius) return;

	const b2cuShape* poly1;
	const b2cuShape* poly2;
	Xf xf1, xf2;
	int edge1;
	bool flip;
	const float k_tol = 0.1f * B2CU_LINEAR_SLOP;

	if (separationB > separationA + k_tol)
	{
		poly1 = polyB;
		poly2 = polyA;
		xf1 = xfB;
		xf2 = xfA;
		edge1 = edgeB;
		m->type = B2CU_MANIFOLD_FACE_B;
		flip = true;
	}
	else
	{
		poly1 = polyA;
		poly2 = polyB;
		xf1 = xfA;
		xf2 = xfB;
		edge1 = edgeA;
		m->type = B2CU_MANIFOLD_FACE_A;
		flip = false;
	}

	// incident edge (b2FindIncidentEdge)
	ClipVertex incidentEdge[2];
	{
		int count2 = poly2->count;
		Vec2 normal1 = MulT(xf2.q, Mul(xf1.q, ShapeN(poly1, edge1)));
		int index = 0;
		float minDot = B2CU_MAX_FLOAT;
		for (int i = 0; i < count2; ++i)
		{
			float dot = Dot(normal1, ShapeN(poly2, i));
			if (dot < minDot)
			{
				minDot = dot;
				index = i;
			}
		}
		int i1 = index;
		int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
		incidentEdge[0].v = Mul(xf2, ShapeV(poly2, i1));
		incidentEdge[0].id = MakeId((uint32_t)edge1, (uint32_t)i1, B2CU_CF_FACE, B2CU_CF_VERTEX);
		incidentEdge[1].v = Mul(xf2, ShapeV(poly2, i2));
		incidentEdge[1].id = MakeId((uint32_t)edge1, (uint32_t)i2, B2CU_CF_FACE, B2CU_CF_VERTEX);
	}

	int count1 = poly1->count;
	int iv1 = edge1;
	int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;

	Vec2 v11 = ShapeV(poly1, iv1);
	Vec2 v12 = ShapeV(poly1, iv2);

	Vec2 localTangent = Normalized(v12 - v11);
	Vec2 localNormal = CrossVS(localTangent, 1.0f);
	Vec2 planePoint = 0.5f * (v11 + v12);

	Vec2 tangent = Mul(xf1.q, localTangent);
	Vec2 normal = CrossVS(tangent, 1.0f);

	v11 = Mul(xf1, v11);
	v12 = Mul(xf1, v12);

	float frontOffset = Dot(normal, v11);
	float sideOffset1 = -Dot(tangent, v11) + totalRadius;
	float sideOffset2 = Dot(tangent, v12) + totalRadius;

	ClipVertex clipPoints1[2];
	ClipVertex clipPoints2[2];
	int np = ClipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
	if (np < 2) return;
	np = ClipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
	if (np < 2) return;

	m->localNormal = localNormal;
	m->localPoint = planePoint;

	int pointCount = 0;
	for (int i = 0; i < B2CU_MAX_MANIFOLD_POINTS; ++i)
	{
		float separation = Dot(normal, clipPoints2[i].v) - frontOffset;
		if (separation <= totalRadius)
		{
			m->lp[pointCount] = MulT(xf2, clipPoints2[i].v);
			m->id[pointCount] = flip ? FlipId(clipPoints2[i].id) : clipPoints2[i].id;
			++pointCount;
		}
	}
	m->pointCount = pointCount;
}

__device__ __forceinline__ void CollideCircles(Manifold* m, const b2cuShape* __restrict__ circleA, const Xf& xfA,
                                               const b2cuShape* __restrict__ circleB, const Xf& xfB)
{
	m->pointCount = 0;

	Vec2 cpA = ShapeV(circleA, 0);
	Vec2 cpB = ShapeV(circleB, 0);
	Vec2 pA = Mul(xfA, cpA);
	Vec2 pB = Mul(xfB, cpB);

	Vec2 d = pB - pA;
	float distSqr = Dot(d, d);
	float rA = circleA->radius, rB = circleB->radius;
	float radius = rA + rB;
	if (distSqr > radius * radius)
	{
		return;
	}

	m->type = B2CU_MANIFOLD_CIRCLES;
	m->localPoint = cpA;
	m->localNormal = V(0.0f, 0.0f);
	m->pointCount = 1;
	m->lp[0] = cpB;
	m->id[0] = 0;
}

__device__ __forceinline__ void CollidePolygonAndCircle(Manifold* m, const b2cuShape* __restrict__ polygonA,
                                                        const Xf& xfA, const b2cuShape* __restrict__ circleB,
                                                        const Xf& xfB)
{
	m->pointCount = 0;

	Vec2 cpB = ShapeV(circleB, 0);
	Vec2 c = Mul(xfB, cpB);
	Vec2 cLocal = MulT(xfA, c);

	int normalIndex = 0;
	float separation = -B2CU_MAX_FLOAT;
	float radius = polygonA->radius + circleB->radius;
	int vertexCount = polygonA->count;

	// all vertices and normals of the record are requested at once (unused slots are zero): the loop below would
	// otherwise wait for two dependent L1 loads per vertex, and half of all contacts of a mixed pile come through here
	Vec2 vs[B2CU_MAX_POLYGON_VERTICES], ns[B2CU_MAX_POLYGON_VERTICES];
#pragma unroll
	for (int i = 0; i < B2CU_MAX_POLYGON_VERTICES; ++i)
	{
		vs[i] = ShapeV(polygonA, i);
		ns[i] = ShapeN(polygonA, i);
	}
	bool outside = false;
#pragma unroll
	for (int i = 0; i < B2CU_MAX_POLYGON_VERTICES; ++i)
	{
		if (i < vertexCount && !outside)
		{
			float s = Dot(ns[i], cLocal - vs[i]);
			if (s > radius)
			{
				outside = true;
			}
			else if (s > separation)
			{
				separation = s;
				normalIndex = i;
			}
		}
	}
	if (outside) return;

	int vertIndex1 = normalIndex;
	int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
	Vec2 v1 = ShapeV(polygonA, vertIndex1);
	Vec2 v2 = ShapeV(polygonA, vertIndex2);

	if (separation < B2CU_EPSILON)
	{
		m->pointCount = 1;
		m->type = B2CU_MANIFOLD_FACE_A;
		m->localNormal = ShapeN(polygonA, normalIndex);
		m->localPoint = 0.5f * (v1 + v2);
		m->lp[0] = cpB;
		m->id[0] = 0;
		return;
	}

	float u1 = Dot(cLocal - v1, v2 - v1);
	float u2 = Dot(cLocal - v2, v1 - v2);
	if (u1 <= 0.0f)
	{
		if (DistanceSquared(cLocal, v1) > radius * radius)
		{
			return;
		}
		m->pointCount = 1;
		m->type = B2CU_MANIFOLD_FACE_A;
		m->localNormal = Normalized(cLocal - v1);
		m->localPoint = v1;
		m->lp[0] = cpB;
		m->id[0] = 0;
	}
	else if (u2 <= 0.0f)
	{
		if (DistanceSquared(cLocal, v2) > radius * radius)
		{
			return;
		}
		m->pointCount = 1;
		m->type = B2CU_MANIFOLD_FACE_A;
		m->localNormal = Normalized(cLocal - v2);
		m->localPoint = v2;
		m->lp[0] = cpB;
		m->id[0] = 0;
	}
	else
	{
		Vec2 faceCenter = 0.5f * (v1 + v2);
		float s = Dot(cLocal - faceCenter, ShapeN(polygonA, vertIndex1));
		if (s > radius)
		{
			return;
		}
		m->pointCount = 1;
		m->type = B2CU_MANIFOLD_FACE_A;
		m->localNormal = ShapeN(polygonA, vertIndex1);
		m->localPoint = faceCenter;
		m->lp[0] = cpB;
		m->id[0] = 0;
	}
}

__device__ __forceinline__ void CollideEdgeAndCircle(Manifold* m, const b2cuShape* __restrict__ edgeA, const Xf& xfA,
                                                     const b2cuShape* __restrict__ circleB, const Xf& xfB)
{
	m->pointCount = 0;

	Vec2 cpB = ShapeV(circleB, 0);
	Vec2 Q = MulT(xfA, Mul(xfB, cpB));

	Vec2 A = ShapeV(edgeA, 0), B = ShapeV(edgeA, 1);
	Vec2 e = B - A;

	float u = Dot(e, B - Q);
	float v = Dot(e, Q - A);

	float radius = edgeA->radius + circleB->radius;
	uint32_t eflags = edgeA->flags;

	// Region A
	if (v <= 0.0f)
	{
		Vec2 P = A;
		Vec2 d = Q - P;
		float dd = Dot(d, d);
		if (dd > radius * radius)
		{
			return;
		}
		if (eflags & B2CU_EDGE_HAS_VERTEX0)
		{
			Vec2 A1 = ShapeV(edgeA, 2);
			Vec2 B1 = A;
			Vec2 e1 = B1 - A1;
			float u1 = Dot(e1, B1 - Q);
			if (u1 > 0.0f)
			{
				return;
			}
		}
		m->pointCount = 1;
		m->type = B2CU_MANIFOLD_CIRCLES;
		m->localNormal = V(0.0f, 0.0f);
		m->localPoint = P;
		m->id[0] = MakeId(0, 0, B2CU_CF_VERTEX, B2CU_CF_VERTEX);
		m->lp[0] = cpB;
		return;
	}

	// Region B
	if (u <= 0.0f)
	{
		Vec2 P = B;
		Vec2 d = Q - P;
		float dd = Dot(d, d);
		if (dd > radius * radius)
		{
			return;
		}
		if (eflags & B2CU_EDGE_HAS_VERTEX3)
		{
			Vec2 B2 = ShapeV(edgeA, 3);
			Vec2 A2 = B;
			Vec2 e2 = B2 - A2;
			float v2 = Dot(e2, Q - A2);
			if (v2 > 0.0f)
			{
				return;
			}
		}
		m->pointCount = 1;
		m->type = B2CU_MANIFOLD_CIRCLES;
		m->localNormal = V(0.0f, 0.0f);
		m->localPoint = P;
		m->id[0] = MakeId(1, 0, B2CU_CF_VERTEX, B2CU_CF_VERTEX);
		m->lp[0] = cpB;
		return;
	}

	// Region AB
	float den = Dot(e, e);
	Vec2 P = (1.0f / den) * (u * A + v * B);
	Vec2 d = Q - P;
	float dd = Dot(d, d);
	if (dd > radius * radius)
	{
		return;
	}

	Vec2 n = V(-e.y, e.x);
	if (Dot(n, Q - A) < 0.0f)
	{
		n = V(-n.x, -n.y);
	}
	n = Normalized(n);

	m->pointCount = 1;
	m->type = B2CU_MANIFOLD_FACE_A;
	m->localNormal = n;
	m->localPoint = A;
	m->id[0] = MakeId(0, 0, B2CU_CF_FACE, B2CU_CF_VERTEX);
	m->lp[0] = cpB;
}

// b2EPCollider::Collide.  Polygon B in frame A is recomputed from the shape table where b2EPCollider keeps a
// b2TempPolygon copy; the values are the same.
__device__ __forceinline__ void CollideEdgeAndPolygon(Manifold* m, const b2cuShape* __restrict__ edgeA, const Xf& xfA,
                                                      const b2cuShape* __restrict__ polygonB, const Xf& xfB)
{
	Xf xf = MulTXf(xfA, xfB);
	Vec2 centroidB = Mul(xf, V(polygonB->centroid[0], polygonB->centroid[1]));

	Vec2 v0 = ShapeV(edgeA, 2);
	Vec2 v1 = ShapeV(edgeA, 0);
	Vec2 v2 = ShapeV(edgeA, 1);
	Vec2 v3 = ShapeV(edgeA, 3);
	bool hasVertex0 = (edgeA->flags & B2CU_EDGE_HAS_VERTEX0) != 0;
	bool hasVertex3 = (edgeA->flags & B2CU_EDGE_HAS_VERTEX3) != 0;

	Vec2 edge1 = Normalized(v2 - v1);
	Vec2 normal1 = V(edge1.y, -edge1.x);
	float offset1 = Dot(normal1, centroidB - v1);
	float offset0 = 0.0f, offset2 = 0.0f;
	bool convex1 = false, convex2 = false;
	Vec2 normal0 = V(0.0f, 0.0f), normal2 = V(0.0f, 0.0f);

	if (hasVertex0)
	{
		Vec2 edge0 = Normalized(v1 - v0);
		normal0 = V(edge0.y, -edge0.x);
		convex1 = Cross(edge0, edge1) >= 0.0f;
		offset0 = Dot(normal0, centroidB - v0);
	}
	if (hasVertex3)
	{
		Vec2 edge2 = Normalized(v3 - v2);
		normal2 = V(edge2.y, -edge2.x);
		convex2 = Cross(edge1, edge2) > 0.0f;
		offset2 = Dot(normal2, centroidB - v2);
	}

	bool front;
	Vec2 normal, lowerLimit, upperLimit;
	if (hasVertex0 && hasVertex3)
	{
		if (convex1 && convex2)
		{
			front = offset0 >= 0.0f || offset1 >= 0.0f || offset2 >= 0.0f;
			if (front) { normal = normal1; lowerLimit = normal0; upperLimit = normal2; }
			else { normal = -normal1; lowerLimit = -normal1; upperLimit = -normal1; }
		}
		else if (convex1)
		{
			front = offset0 >= 0.0f || (offset1 >= 0.0f && offset2 >= 0.0f);
			if (front) { normal = normal1; lowerLimit = normal0; upperLimit = normal1; }
			else { normal = -normal1; lowerLimit = -normal2; upperLimit = -normal1; }
		}
		else if (convex2)
		{
			front = offset2 >= 0.0f || (offset0 >= 0.0f && offset1 >= 0.0f);
			if (front) { normal = normal1; lowerLimit = normal1; upperLimit = normal2; }
			else { normal = -normal1; lowerLimit = -normal1; upperLimit = -normal0; }
		}
		else
		{
			front = offset0 >= 0.0f && offset1 >= 0.0f && offset2 >= 0.0f;
			if (front) { normal = normal1; lowerLimit = normal1; upperLimit = normal1; }
			else { normal = -normal1; lowerLimit = -normal2; upperLimit = -normal0; }
		}
	}
	else if (hasVertex0)
	{
		if (convex1)
		{
			front = offset0 >= 0.0f || offset1 >= 0.0f;
			if (front) { normal = normal1; lowerLimit = normal0; upperLimit = -normal1; }
			else { normal = -normal1; lowerLimit = normal1; upperLimit = -normal1; }
		}
		else
		{
			front = offset0 >= 0.0f && offset1 >= 0.0f;
			if (front) { normal = normal1; lowerLimit = normal1; upperLimit = -normal1; }
			else { normal = -normal1; lowerLimit = normal1; upperLimit = -normal0; }
		}
	}
	else if (hasVertex3)
	{
		if (convex2)
		{
			front = offset1 >= 0.0f || offset2 >= 0.0f;
			if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = normal2; }
			else { normal = -normal1; lowerLimit = -normal1; upperLimit = normal1; }
		}
		else
		{
			front = offset1 >= 0.0f && offset2 >= 0.0f;
			if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = normal1; }
			else { normal = -normal1; lowerLimit = -normal2; upperLimit = normal1; }
		}
	}
	else
	{
		front = offset1 >= 0.0f;
		if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = -normal1; }
		else { normal = -normal1; lowerLimit = normal1; upperLimit = normal1; }
	}

	int countB = polygonB->count;
	float radius = polygonB->radius + edgeA->radius;

	m->pointCount = 0;

	// ComputeEdgeSeparation
	float edgeSeparation = B2CU_MAX_FLOAT;
	for (int i = 0; i < countB; ++i)
	{
		float s = Dot(normal, Mul(xf, ShapeV(polygonB, i)) - v1);
		if (s < edgeSeparation)
		{
			edgeSeparation = s;
		}
	}
	// the edge axis type is always e_edgeA, so the "unknown" early-out of the reference never triggers
	if (edgeSeparation > radius)
	{
		return;
	}

	// ComputePolygonSeparation
	int polyType = 0; // 0 unknown, 2 edgeB
	int polyIndex = -1;
	float polySeparation = -B2CU_MAX_FLOAT;
	{
		Vec2 perp = V(-normal.y, normal.x);
		for (int i = 0; i < countB; ++i)
		{
			Vec2 n = -Mul(xf.q, ShapeN(polygonB, i));
			Vec2 vi = Mul(xf, ShapeV(polygonB, i));

			float s1 = Dot(n, vi - v1);
			float s2 = Dot(n, vi - v2);
			float s = Min(s1, s2);

			if (s > radius)
			{
				polyType = 2;
				polyIndex = i;
				polySeparation = s;
				break;
			}

			if (Dot(n, perp) >= 0.0f)
			{
				if (Dot(n - upperLimit, normal) < -B2CU_ANGULAR_SLOP)
				{
					continue;
				}
			}
			else
			{
				if (Dot(n - lowerLimit, normal) < -B2CU_ANGULAR_SLOP)
				{
					continue;
				}
			}

			if (s > polySeparation)
			{
				polyType = 2;
				polyIndex = i;
				polySeparation = s;
			}
		}
	}
	if (polyType != 0 && polySeparation > radius)
	{
		return;
	}

	const float k_relativeTol = 0.98f;
	const float k_absoluteTol = 0.001f;

	bool primaryIsEdgeA;
	if (polyType == 0)
	{
		primaryIsEdgeA = true;
	}
	else if (polySeparation > k_relativeTol * edgeSeparation + k_absoluteTol)
	{
		primaryIsEdgeA = false;
	}
	else
	{
		primaryIsEdgeA = true;
	}

	ClipVertex ie[2];
	int rf_i1, rf_i2;
	Vec2 rf_v1, rf_v2, rf_normal;
	if (primaryIsEdgeA)
	{
		m->type = B2CU_MANIFOLD_FACE_A;

		int bestIndex = 0;
		float bestValue = Dot(normal, Mul(xf.q, ShapeN(polygonB, 0)));
		for (int i = 1; i < countB; ++i)
		{
			float value = Dot(normal, Mul(xf.q, ShapeN(polygonB, i)));
			if (value < bestValue)
			{
				bestValue = value;
				bestIndex = i;
			}
		}

		int i1 = bestIndex;
		int i2 = i1 + 1 < countB ? i1 + 1 : 0;

		ie[0].v = Mul(xf, ShapeV(polygonB, i1));
		ie[0].id = MakeId(0, (uint32_t)i1, B2CU_CF_FACE, B2CU_CF_VERTEX);
		ie[1].v = Mul(xf, ShapeV(polygonB, i2));
		ie[1].id = MakeId(0, (uint32_t)i2, B2CU_CF_FACE, B2CU_CF_VERTEX);

		if (front)
		{
			rf_i1 = 0;
			rf_i2 = 1;
			rf_v1 = v1;
			rf_v2 = v2;
			rf_normal = normal1;
		}
		else
		{
			rf_i1 = 1;
			rf_i2 = 0;
			rf_v1 = v2;
			rf_v2 = v1;
			rf_normal = -normal1;
		}
	}
	else
	{
		m->type = B2CU_MANIFOLD_FACE_B;

		ie[0].v = v1;
		ie[0].id = MakeId(0, (uint32_t)polyIndex, B2CU_CF_VERTEX, B2CU_CF_FACE);
		ie[1].v = v2;
		ie[1].id = MakeId(0, (uint32_t)polyIndex, B2CU_CF_VERTEX, B2CU_CF_FACE);

		rf_i1 = polyIndex;
		rf_i2 = rf_i1 + 1 < countB ? rf_i1 + 1 : 0;
		rf_v1 = Mul(xf, ShapeV(polygonB, rf_i1));
		rf_v2 = Mul(xf, ShapeV(polygonB, rf_i2));
		rf_normal = Mul(xf.q, ShapeN(polygonB, rf_i1));
	}

	Vec2 sideNormal1 = V(rf_normal.y, -rf_normal.x);
	Vec2 sideNormal2 = -sideNormal1;
	float sideOffset1 = Dot(sideNormal1, rf_v1);
	float sideOffset2 = Dot(sideNormal2, rf_v2);

	ClipVertex clipPoints1[2];
	ClipVertex clipPoints2[2];
	int np = ClipSegmentToLine(clipPoints1, ie, sideNormal1, sideOffset1, rf_i1);
	if (np < B2CU_MAX_MANIFOLD_POINTS)
	{
		return;
	}
	np = ClipSegmentToLine(clipPoints2, clipPoints1, sideNormal2, sideOffset2, rf_i2);
	if (np < B2CU_MAX_MANIFOLD_POINTS)
	{
		return;
	}

	if (primaryIsEdgeA)
	{
		m->localNormal = rf_normal;
		m->localPoint = rf_v1;
	}
	else
	{
		m->localNormal = ShapeN(polygonB, rf_i1);
		m->localPoint = ShapeV(polygonB, rf_i1);
	}

	int pointCount = 0;
	for (int i = 0; i < B2CU_MAX_MANIFOLD_POINTS; ++i)
	{
		float separation = Dot(rf_normal, clipPoints2[i].v - rf_v1);
		if (separation <= radius)
		{
			if (primaryIsEdgeA)
			{
				m->lp[pointCount] = MulT(xf, clipPoints2[i].v);
				m->id[pointCount] = clipPoints2[i].id;
			}
			else
			{
				m->lp[pointCount] = clipPoints2[i].v;
				m->id[pointCount] = FlipId(clipPoints2[i].id);
			}
			++pointCount;
		}
	}
	m->pointCount = pointCount;
}

// Type dispatch of b2Contact::Evaluate (Box2D/Dynamics/Contacts/*.cpp:44-52).  Shape A is the primary type.
__device__ __forceinline__ void Evaluate(Manifold* m, const b2cuShape* __restrict__ sA, const Xf& xfA,
                                         const b2cuShape* __restrict__ sB, const Xf& xfB)
{
	int tA = sA->type, tB = sB->type;
	if (tA == B2CU_SHAPE_POLYGON)
	{
		if (tB == B2CU_SHAPE_POLYGON)
		{
			CollidePolygons(m, sA, xfA, sB, xfB);
		}
		else
		{
			CollidePolygonAndCircle(m, sA, xfA, sB, xfB);
		}
	}
	else if (tA == B2CU_SHAPE_CIRCLE)
	{
		CollideCircles(m, sA, xfA, sB, xfB);
	}
	else
	{
		if (tB == B2CU_SHAPE_CIRCLE)
		{
			CollideEdgeAndCircle(m, sA, xfA, sB, xfB);
		}
		else
		{
			CollideEdgeAndPolygon(m, sA, xfA, sB, xfB);
		}
	}
}

// true when (typeA, typeB) must be swapped to reach the primary order (b2Contact.cpp:44-50, :82-93)
__host__ __device__ __forceinline__ bool NeedsSwap(int typeA, int typeB)
{
	// primary pairs: (circle,circle) (polygon,circle) (polygon,polygon) (edge,circle) (edge,polygon)
	if (typeA == B2CU_SHAPE_CIRCLE && typeB != B2CU_SHAPE_CIRCLE) return true;
	if (typeA == B2CU_SHAPE_POLYGON && typeB == B2CU_SHAPE_EDGE) return true;
	return false;
}

} // namespace b2cu
