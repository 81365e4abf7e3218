// b2cu_gjk.cuh -- closest points of two convex shapes (GJK), the routine behind b2TestOverlap (sensor contacts) and
// b2TimeOfImpact.  Restates Box2D/Collision/b2Distance.cpp:30-603 of the reference with the same arithmetic in the
// same order (fp32, no FMA), on the device's geometry records: a proxy is (record, vertex count, radius) -- one vertex
// for a circle, two for an edge or a chain segment, m_count for a polygon -- and reads its vertices through the
// read-only path instead of copying them.
#pragma once

#include "b2cu_collide.cuh"

namespace b2cu
{

struct GjkProxy
{
	const b2cuShape* shape;
	int count;
	float radius;
};

__device__ __forceinline__ GjkProxy MakeGjkProxy(const b2cuShape* s)
{
	GjkProxy p;
	p.shape = s;
	p.radius = s->radius;
	int type = s->type;
	p.count = type == B2CU_SHAPE_CIRCLE ? 1 : (type == B2CU_SHAPE_EDGE ? 2 : s->count);
	return p;
}

// b2DistanceProxy::GetSupport (b2Distance.h:118-133): first vertex with the largest projection
__device__ __forceinline__ int GjkSupport(const GjkProxy& p, Vec2 dir)
{
	int best = 0;
	float bestValue = Dot(ShapeV(p.shape, 0), dir);
	for (int i = 1; i < p.count; ++i)
	{
		float value = Dot(ShapeV(p.shape, i), dir);
		if (value > bestValue)
		{
			best = i;
			bestValue = value;
		}
	}
	return best;
}

// b2SimplexCache (b2Distance.h:66-73)
struct GjkCache
{
	float metric;
	int count;
	int indexA[3], indexB[3];
};

struct GjkVertex
{
	Vec2 wA, wB, w; // support points and their difference
	float a;        // barycentric weight of the closest point
	int iA, iB;
};

struct GjkSimplex
{
	GjkVertex v[3];
	int count;
};

__device__ __forceinline__ float GjkMetric(const GjkSimplex& s)
{
	if (s.count == 2) return Length(s.v[0].w - s.v[1].w);
	if (s.count == 3) return Cross(s.v[1].w - s.v[0].w, s.v[2].w - s.v[0].w);
	return 0.0f;
}

__device__ __forceinline__ void GjkSetVertex(GjkVertex& o, const GjkProxy& A, const Xf& xfA, int iA, const GjkProxy& B,
                                             const Xf& xfB, int iB)
{
	o.iA = iA;
	o.iB = iB;
	o.wA = Mul(xfA, ShapeV(A.shape, iA));
	o.wB = Mul(xfB, ShapeV(B.shape, iB));
	o.w = o.wB - o.wA;
}

// line segment case: barycentric weights of the point of [w1, w2] closest to the origin (b2Simplex::Solve2)
__device__ __forceinline__ void GjkSolve2(GjkSimplex& s)
{
	Vec2 w1 = s.v[0].w, w2 = s.v[1].w;
	Vec2 e12 = w2 - w1;
	float d12_2 = -Dot(w1, e12);
	if (d12_2 <= 0.0f)
	{
		s.v[0].a = 1.0f;
		s.count = 1;
		return;
	}
	float d12_1 = Dot(w2, e12);
	if (d12_1 <= 0.0f)
	{
		s.v[1].a = 1.0f;
		s.count = 1;
		s.v[0] = s.v[1];
		return;
	}
	float inv = 1.0f / (d12_1 + d12_2);
	s.v[0].a = d12_1 * inv;
	s.v[1].a = d12_2 * inv;
	s.count = 2;
}

// triangle case: vertex regions, edge regions, interior (b2Simplex::Solve3)
__device__ __forceinline__ void GjkSolve3(GjkSimplex& s)
{
	Vec2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;

	Vec2 e12 = w2 - w1;
	float d12_1 = Dot(w2, e12);
	float d12_2 = -Dot(w1, e12);

	Vec2 e13 = w3 - w1;
	float d13_1 = Dot(w3, e13);
	float d13_2 = -Dot(w1, e13);

	Vec2 e23 = w3 - w2;
	float d23_1 = Dot(w3, e23);
	float d23_2 = -Dot(w2, e23);

	float n123 = Cross(e12, e13);
	float d123_1 = n123 * Cross(w2, w3);
	float d123_2 = n123 * Cross(w3, w1);
	float d123_3 = n123 * Cross(w1, w2);

	if (d12_2 <= 0.0f && d13_2 <= 0.0f)
	{
		s.v[0].a = 1.0f;
		s.count = 1;
		return;
	}
	if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f)
	{
		float inv = 1.0f / (d12_1 + d12_2);
		s.v[0].a = d12_1 * inv;
		s.v[1].a = d12_2 * inv;
		s.count = 2;
		return;
	}
	if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f)
	{
		float inv = 1.0f / (d13_1 + d13_2);
		s.v[0].a = d13_1 * inv;
		s.v[2].a = d13_2 * inv;
		s.count = 2;
		s.v[1] = s.v[2];
		return;
	}
	if (d12_1 <= 0.0f && d23_2 <= 0.0f)
	{
		s.v[1].a = 1.0f;
		s.count = 1;
		s.v[0] = s.v[1];
		return;
	}
	if (d13_1 <= 0.0f && d23_1 <= 0.0f)
	{
		s.v[2].a = 1.0f;
		s.count = 1;
		s.v[0] = s.v[2];
		return;
	}
	if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f)
	{
		float inv = 1.0f / (d23_1 + d23_2);
		s.v[1].a = d23_1 * inv;
		s.v[2].a = d23_2 * inv;
		s.count = 2;
		s.v[0] = s.v[2];
		return;
	}
	float inv = 1.0f / (d123_1 + d123_2 + d123_3);
	s.v[0].a = d123_1 * inv;
	s.v[1].a = d123_2 * inv;
	s.v[2].a = d123_3 * inv;
	s.count = 3;
}

struct GjkOutput
{
	Vec2 pointA, pointB;
	float distance;
	int iterations;
};

// b2Distance (b2Distance.cpp:452-603).  `cache` carries the simplex from one call to the next (count = 0: cold start).
__device__ __forceinline__ void GjkDistance(GjkOutput* out, GjkCache* cache, const GjkProxy& A, const Xf& xfA,
                                            const GjkProxy& B, const Xf& xfB, bool useRadii)
{
	GjkSimplex s;
	// ---- b2Simplex::ReadCache ----
	s.count = cache->count;
	for (int i = 0; i < s.count; ++i)
	{
		GjkSetVertex(s.v[i], A, xfA, cache->indexA[i], B, xfB, cache->indexB[i]);
		s.v[i].a = 0.0f;
	}
	if (s.count > 1)
	{
		float metric1 = cache->metric;
		float metric2 = GjkMetric(s);
		if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < B2CU_EPSILON) s.count = 0;
	}
	if (s.count == 0)
	{
		GjkSetVertex(s.v[0], A, xfA, 0, B, xfB, 0);
		s.v[0].a = 1.0f;
		s.count = 1;
	}

	const int kMaxIters = 20;
	int saveA[3], saveB[3];
	int saveCount = 0;
	int iter = 0;
	while (iter < kMaxIters)
	{
		saveCount = s.count;
		for (int i = 0; i < saveCount; ++i)
		{
			saveA[i] = s.v[i].iA;
			saveB[i] = s.v[i].iB;
		}

		if (s.count == 2) GjkSolve2(s);
		else if (s.count == 3) GjkSolve3(s);

		// a full triangle contains the origin: the shapes overlap
		if (s.count == 3) break;

		// b2Simplex::GetSearchDirection
		Vec2 dir;
		if (s.count == 1)
		{
			dir = -s.v[0].w;
		}
		else
		{
			Vec2 e12 = s.v[1].w - s.v[0].w;
			float sgn = Cross(e12, -s.v[0].w);
			dir = sgn > 0.0f ? CrossSV(1.0f, e12) : CrossVS(e12, 1.0f);
		}
		// the origin lies on the simplex (numerically): no direction to search in
		if (Dot(dir, dir) < B2CU_EPSILON * B2CU_EPSILON) break;

		GjkVertex& nv = s.v[s.count];
		int iA = GjkSupport(A, MulT(xfA.q, -dir));
		int iB = GjkSupport(B, MulT(xfB.q, dir));
		GjkSetVertex(nv, A, xfA, iA, B, xfB, iB);
		++iter;

		// the same support pair again: no progress possible
		bool duplicate = false;
		for (int i = 0; i < saveCount; ++i)
		{
			if (iA == saveA[i] && iB == saveB[i])
			{
				duplicate = true;
				break;
			}
		}
		if (duplicate) break;
		++s.count;
	}

	// ---- b2Simplex::GetWitnessPoints ----
	if (s.count == 1)
	{
		out->pointA = s.v[0].wA;
		out->pointB = s.v[0].wB;
	}
	else if (s.count == 2)
	{
		out->pointA = s.v[0].a * s.v[0].wA + s.v[1].a * s.v[1].wA;
		out->pointB = s.v[0].a * s.v[0].wB + s.v[1].a * s.v[1].wB;
	}
	else
	{
		out->pointA = s.v[0].a * s.v[0].wA + s.v[1].a * s.v[1].wA + s.v[2].a * s.v[2].wA;
		out->pointB = out->pointA;
	}
	out->distance = Length(out->pointA - out->pointB);
	out->iterations = iter;

	// ---- b2Simplex::WriteCache ----
	cache->metric = GjkMetric(s);
	cache->count = s.count;
	for (int i = 0; i < s.count; ++i)
	{
		cache->indexA[i] = s.v[i].iA;
		cache->indexB[i] = s.v[i].iB;
	}

	if (useRadii)
	{
		float rA = A.radius, rB = B.radius;
		if (out->distance > rA + rB && out->distance > B2CU_EPSILON)
		{
			// shapes still apart: move the witness points to the surfaces
			out->distance -= rA + rB;
			Vec2 normal = Normalized(out->pointB - out->pointA);
			out->pointA = out->pointA + rA * normal;
			out->pointB = out->pointB - rB * normal;
		}
		else
		{
			// overlapping once the radii are counted: meet in the middle
			Vec2 p = 0.5f * (out->pointA + out->pointB);
			out->pointA = p;
			out->pointB = p;
			out->distance = 0.0f;
		}
	}
}

// b2TestOverlap (b2Collision.cpp:233-252)
__device__ __forceinline__ bool TestOverlap(const b2cuShape* sA, const Xf& xfA, const b2cuShape* sB, const Xf& xfB)
{
	GjkCache cache;
	cache.count = 0;
	cache.metric = 0.0f;
	GjkOutput out;
	GjkDistance(&out, &cache, MakeGjkProxy(sA), xfA, MakeGjkProxy(sB), xfB, true);
	return out.distance < 10.0f * B2CU_EPSILON;
}

} // namespace b2cu
