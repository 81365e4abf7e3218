// b2cu_toi.cuh -- time of impact of two swept convex shapes by conservative advancement (b2TimeOfImpact), the routine
// the continuous-collision sub-stepper is built on.  Restates Box2D/Collision/b2TimeOfImpact.cpp:36-497 and
// b2Sweep::GetTransform / Normalize (Box2D/Common/b2Math.h:679-705) of the reference with the same fp32 arithmetic
// in the same order.
#pragma once

#include "b2cu_gjk.cuh"

namespace b2cu
{

// b2Sweep: motion of a body over the step; positions are of the centre of mass
struct Sweep
{
	Vec2 localCenter;
	Vec2 c0, c;
	float a0, a;
	float alpha0;
};

// interpolated transform at beta in [0, 1] (b2Sweep::GetTransform)
__device__ __forceinline__ Xf SweepTransform(const Sweep& s, float beta)
{
	Xf xf;
	xf.p = (1.0f - beta) * s.c0 + beta * s.c;
	float angle = (1.0f - beta) * s.a0 + beta * s.a;
	xf.q = SinCos(angle);
	xf.p = xf.p - Mul(xf.q, s.localCenter);
	return xf;
}

// b2Sweep::Normalize: bring a0 into [0, 2 pi)
__device__ __forceinline__ void SweepNormalize(Sweep& s)
{
	float twoPi = 2.0f * B2CU_PI;
	float d = twoPi * floorf(s.a0 / twoPi);
	s.a0 -= d;
	s.a -= d;
}

// b2Sweep::Advance (b2Math.h:689-697)
__device__ __forceinline__ void SweepAdvance(Sweep& s, float alpha)
{
	float beta = (alpha - s.alpha0) / (1.0f - s.alpha0);
	s.c0 = s.c0 + beta * (s.c - s.c0);
	s.a0 += beta * (s.a - s.a0);
	s.alpha0 = alpha;
}

enum ToiState
{
	TOI_UNKNOWN = 0,
	TOI_FAILED = 1,
	TOI_OVERLAPPED = 2,
	TOI_TOUCHING = 3,
	TOI_SEPARATED = 4
};

// b2SeparationFunction: signed separation along an axis fixed by the closest features at t1
struct SeparationFunction
{
	GjkProxy A, B;
	Sweep sweepA, sweepB;
	int type; // 0 points, 1 face of A, 2 face of B
	Vec2 localPoint, axis;
};

__device__ __forceinline__ void SeparationInit(SeparationFunction& f, const GjkCache& cache, const GjkProxy& A,
                                               const Sweep& sweepA, const GjkProxy& B, const Sweep& sweepB, float t1)
{
	f.A = A;
	f.B = B;
	f.sweepA = sweepA;
	f.sweepB = sweepB;
	Xf xfA = SweepTransform(sweepA, t1), xfB = SweepTransform(sweepB, t1);
	if (cache.count == 1)
	{
		f.type = 0;
		Vec2 pointA = Mul(xfA, ShapeV(A.shape, cache.indexA[0]));
		Vec2 pointB = Mul(xfB, ShapeV(B.shape, cache.indexB[0]));
		f.axis = Normalized(pointB - pointA);
		f.localPoint = V(0.0f, 0.0f);
	}
	else if (cache.indexA[0] == cache.indexA[1])
	{
		// two points on B, one on A: B's face
		f.type = 2;
		Vec2 b1 = ShapeV(B.shape, cache.indexB[0]), b2 = ShapeV(B.shape, cache.indexB[1]);
		f.axis = Normalized(CrossVS(b2 - b1, 1.0f));
		Vec2 normal = Mul(xfB.q, f.axis);
		f.localPoint = 0.5f * (b1 + b2);
		Vec2 pointB = Mul(xfB, f.localPoint);
		Vec2 pointA = Mul(xfA, ShapeV(A.shape, cache.indexA[0]));
		if (Dot(pointA - pointB, normal) < 0.0f) f.axis = -f.axis;
	}
	else
	{
		f.type = 1;
		Vec2 a1 = ShapeV(A.shape, cache.indexA[0]), a2 = ShapeV(A.shape, cache.indexA[1]);
		f.axis = Normalized(CrossVS(a2 - a1, 1.0f));
		Vec2 normal = Mul(xfA.q, f.axis);
		f.localPoint = 0.5f * (a1 + a2);
		Vec2 pointA = Mul(xfA, f.localPoint);
		Vec2 pointB = Mul(xfB, ShapeV(B.shape, cache.indexB[0]));
		if (Dot(pointB - pointA, normal) < 0.0f) f.axis = -f.axis;
	}
}

// deepest pair of features along the axis at time t
__device__ __forceinline__ float SeparationFindMin(const SeparationFunction& f, int* indexA, int* indexB, float t)
{
	Xf xfA = SweepTransform(f.sweepA, t), xfB = SweepTransform(f.sweepB, t);
	if (f.type == 0)
	{
		*indexA = GjkSupport(f.A, MulT(xfA.q, f.axis));
		*indexB = GjkSupport(f.B, MulT(xfB.q, -f.axis));
		Vec2 pointA = Mul(xfA, ShapeV(f.A.shape, *indexA));
		Vec2 pointB = Mul(xfB, ShapeV(f.B.shape, *indexB));
		return Dot(pointB - pointA, f.axis);
	}
	if (f.type == 1)
	{
		Vec2 normal = Mul(xfA.q, f.axis);
		Vec2 pointA = Mul(xfA, f.localPoint);
		*indexA = -1;
		*indexB = GjkSupport(f.B, MulT(xfB.q, -normal));
		Vec2 pointB = Mul(xfB, ShapeV(f.B.shape, *indexB));
		return Dot(pointB - pointA, normal);
	}
	Vec2 normal = Mul(xfB.q, f.axis);
	Vec2 pointB = Mul(xfB, f.localPoint);
	*indexB = -1;
	*indexA = GjkSupport(f.A, MulT(xfA.q, -normal));
	Vec2 pointA = Mul(xfA, ShapeV(f.A.shape, *indexA));
	return Dot(pointA - pointB, normal);
}

// separation of a fixed pair of features at time t
__device__ __forceinline__ float SeparationEvaluate(const SeparationFunction& f, int indexA, int indexB, float t)
{
	Xf xfA = SweepTransform(f.sweepA, t), xfB = SweepTransform(f.sweepB, t);
	if (f.type == 0)
	{
		Vec2 pointA = Mul(xfA, ShapeV(f.A.shape, indexA));
		Vec2 pointB = Mul(xfB, ShapeV(f.B.shape, indexB));
		return Dot(pointB - pointA, f.axis);
	}
	if (f.type == 1)
	{
		Vec2 normal = Mul(xfA.q, f.axis);
		Vec2 pointA = Mul(xfA, f.localPoint);
		Vec2 pointB = Mul(xfB, ShapeV(f.B.shape, indexB));
		return Dot(pointB - pointA, normal);
	}
	Vec2 normal = Mul(xfB.q, f.axis);
	Vec2 pointB = Mul(xfB, f.localPoint);
	Vec2 pointA = Mul(xfA, ShapeV(f.A.shape, indexA));
	return Dot(pointA - pointB, normal);
}

// b2TimeOfImpact (b2TimeOfImpact.cpp:256-497): first time in [0, tMax] at which the shapes come within the target
// separation.  Returns the b2TOIOutput state, *tOut the time.
__device__ __forceinline__ int TimeOfImpact(float* tOut, const GjkProxy& A, Sweep sweepA, const GjkProxy& B, Sweep sweepB,
                                            float tMax)
{
	int state = TOI_UNKNOWN;
	float tResult = tMax;
	SweepNormalize(sweepA);
	SweepNormalize(sweepB);

	float totalRadius = A.radius + B.radius;
	float target = Max(B2CU_LINEAR_SLOP, totalRadius - 3.0f * B2CU_LINEAR_SLOP);
	float tolerance = 0.25f * B2CU_LINEAR_SLOP;

	float t1 = 0.0f;
	const int kMaxIterations = 20;
	int iter = 0;
	GjkCache cache;
	cache.count = 0;
	cache.metric = 0.0f;

	// outer loop: new separating axis from the closest features at t1
	for (;;)
	{
		Xf xfA = SweepTransform(sweepA, t1), xfB = SweepTransform(sweepB, t1);
		GjkOutput dist;
		GjkDistance(&dist, &cache, A, xfA, B, xfB, false);
		if (dist.distance <= 0.0f)
		{
			state = TOI_OVERLAPPED;
			tResult = 0.0f;
			break;
		}
		if (dist.distance < target + tolerance)
		{
			state = TOI_TOUCHING;
			tResult = t1;
			break;
		}

		SeparationFunction fcn;
		SeparationInit(fcn, cache, A, sweepA, B, sweepB, t1);

		// inner loop: push t2 back until the deepest point at t2 is at the target separation
		bool done = false;
		float t2 = tMax;
		int pushBackIter = 0;
		for (;;)
		{
			int indexA, indexB;
			float s2 = SeparationFindMin(fcn, &indexA, &indexB, t2);
			if (s2 > target + tolerance)
			{
				state = TOI_SEPARATED;
				tResult = tMax;
				done = true;
				break;
			}
			if (s2 > target - tolerance)
			{
				t1 = t2; // advance the sweeps
				break;
			}
			float s1 = SeparationEvaluate(fcn, indexA, indexB, t1);
			if (s1 < target - tolerance)
			{
				state = TOI_FAILED;
				tResult = t1;
				done = true;
				break;
			}
			if (s1 <= target + tolerance)
			{
				state = TOI_TOUCHING;
				tResult = t1;
				done = true;
				break;
			}
			// 1-D root of s(t) = target on [t1, t2]: secant and bisection steps alternate
			int rootIterCount = 0;
			float a1 = t1, a2 = t2;
			for (;;)
			{
				float t;
				if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
				else t = 0.5f * (a1 + a2);
				++rootIterCount;
				float s = SeparationEvaluate(fcn, indexA, indexB, t);
				if (fabsf(s - target) < tolerance)
				{
					t2 = t;
					break;
				}
				if (s > target)
				{
					a1 = t;
					s1 = s;
				}
				else
				{
					a2 = t;
					s2 = s;
				}
				if (rootIterCount == 50) break;
			}
			++pushBackIter;
			if (pushBackIter == B2CU_MAX_POLYGON_VERTICES) break;
		}
		++iter;
		if (done) break;
		if (iter == kMaxIterations)
		{
			state = TOI_FAILED; // root finder got stuck
			tResult = t1;
			break;
		}
	}
	*tOut = tResult;
	return state;
}

} // namespace b2cu
