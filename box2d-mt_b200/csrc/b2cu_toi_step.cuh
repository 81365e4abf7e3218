// b2cu_toi_step.cuh -- the continuous-collision part of a step: b2World::SolveTOI (Box2D/Dynamics/b2World.cpp:
// 1026-1093) with its FindMinToiContact passes (:1525-1611, ComputeToi :366-447) and the time-of-impact events
// (b2World::StepSolveTOI :851-1024, b2Island::SolveTOI b2Island.cpp:398-530, b2ContactSolver::SolveTOIPositionConstraints
// b2ContactSolver.cpp:755-843).  Included by b2cu_kernels.cuh; the loop over the events is driven by world.cu.
//
// The reference processes the events one after the other on the user thread; so does this file, event by event, but
// every part of an event that touches more than a handful of items is a kernel over the whole world:
//
//   FindMinToiContact   ToiSelectKernel (who is eligible, whose cached time still holds), ToiComputeKernel
//                       (b2TimeOfImpact of the others, one thread each), ToiMinKeyKernel (ToiLessThan tie-break)
//   StepSolveTOI        ToiEventPrepareKernel  the contact lists of the two bodies: a scan of the contact set, ordered
//                                              afterwards by creation stamp (newest first, as OnContactCreate links them)
//                       ToiEventKernel         ONE CTA: advance the two bodies, update the contact, sort the lists
//                                              (shared memory), evaluate the listed contacts' manifolds in parallel,
//                                              settle on one byte per entry what the walk visits before the 32-contact /
//                                              64-body caps end it (one thread, flags only), commit the visited contacts
//                                              and the joined bodies in parallel, solve the island (rows and bodies set
//                                              up by all threads, the sequential impulses by one), write the bodies back
//                       ToiAfterEventKernel    SynchronizeFixtures of the displaced bodies + the flag reset of their
//                                              contacts (b2World.cpp:995-1013)
//                       ToiFindPairsKernel     FindNewContacts for the handful of moved proxies: every proxy of the
//                                              world against each of them (16 B per proxy: a million proxies are a few
//                                              microseconds of HBM time), then the usual insertion into the contact set
//   ClearPostSolveTOI   ToiClearKernel (b2World.cpp:1467-1504)
#pragma once

namespace b2cu
{

#define B2CU_TOI_THREADS 256
#define B2CU_TOI_MAX_CONTACTS 32  // b2_maxTOIContacts, b2Settings.h:95
#define B2CU_TOI_MAX_BODIES 64    // b2_toiBodyCapacity, b2World.cpp:41
#define B2CU_TOI_BAUMGARTE 0.75f  // b2_toiBaugarte, b2Settings.h:123

// ---------------------------------------------------------------------------------------------------------
// FindMinToiContact
// ---------------------------------------------------------------------------------------------------------

// the filters of b2FindMinToiContactTask / FindMinToiContact (b2World.cpp:317-326, :1590-1598) on a candidate
__device__ __forceinline__ bool ToiEligible(const DeviceArrays& d, int i, uint32_t f)
{
	if (f & B2CU_CONTACT_DEAD) return false;
	if ((f & (B2CU_CONTACT_TOI_CANDIDATE | B2CU_CONTACT_ENABLED)) != (B2CU_CONTACT_TOI_CANDIDATE | B2CU_CONTACT_ENABLED))
		return false;
	if (d.c.toiCount[i] > B2CU_MAX_SUB_STEPS) return false;
	int4 pr = d.c.proxies[i];
	return IsAwakeNonStatic(d.bflags[pr.z]) || IsAwakeNonStatic(d.bflags[pr.w]);
}

// One pass over the contact set: candidates whose cached time of impact is still valid (e_toiFlag) enter the minimum
// directly, the others are listed for ToiComputeKernel.
__global__ void __launch_bounds__(256) ToiSelectKernel(DeviceArrays d, int contactCount, int* __restrict__ work)
{
	GridDependencyWait();
	int eligible = 0;
	B2CU_GRID_STRIDE(i, contactCount)
	{
		uint32_t f = d.c.flags[i];
		if (!ToiEligible(d, i, f)) continue;
		++eligible;
		if (f & B2CU_CONTACT_TOI)
		{
			// alpha is in [0, 1]: its bit pattern orders like the value
			atomicMin(reinterpret_cast<unsigned int*>(d.counters + CNT_TOI_MIN_ALPHA), __float_as_uint(d.c.mix[i].w));
		}
		else
		{
			work[atomicAdd(&d.counters[CNT_TOI_WORK], 1)] = i;
		}
	}
	for (int dlt = 16; dlt > 0; dlt >>= 1) eligible += __shfl_down_sync(0xffffffffu, eligible, dlt);
	if ((threadIdx.x & 31) == 0 && eligible) atomicAdd(&d.counters[CNT_TOI], eligible);
}

// b2World::ComputeToi (b2World.cpp:366-447) for the listed contacts.  The two sweeps are put on the same interval
// first: the one that lags is advanced (c0 / a0 / alpha0 only, b2Sweep::Advance).  After a complete step every sweep
// starts at 0 and nothing is advanced; after an event the contacts to recompute are those of the displaced bodies, which
// all sit at the event's alpha, so whichever thread advances a neighbour writes the same 16 bytes.
__global__ void __launch_bounds__(64) ToiComputeKernel(DeviceArrays d, const int* __restrict__ work)
{
	GridDependencyWait();
	const int count = d.counters[CNT_TOI_WORK];
	B2CU_GRID_STRIDE(k, count)
	{
		int i = work[k];
		int4 pr = d.c.proxies[i];
		Sweep sA = LoadSweep(d, pr.z), sB = LoadSweep(d, pr.w);
		float alpha0 = sA.alpha0;
		if (sA.alpha0 < sB.alpha0)
		{
			alpha0 = sB.alpha0;
			SweepAdvance(sA, alpha0);
			d.pos0[pr.z] = make_float4(sA.c0.x, sA.c0.y, sA.a0, sA.alpha0);
		}
		else if (sB.alpha0 < sA.alpha0)
		{
			alpha0 = sA.alpha0;
			SweepAdvance(sB, alpha0);
			d.pos0[pr.w] = make_float4(sB.c0.x, sB.c0.y, sB.a0, sB.alpha0);
		}
		float t;
		int state = TimeOfImpact(&t, MakeGjkProxy(d.shapes + d.pshape[pr.x]), sA, MakeGjkProxy(d.shapes + d.pshape[pr.y]),
		                         sB, 1.0f);
		float alpha = state == TOI_TOUCHING ? Min(alpha0 + (1.0f - alpha0) * t, 1.0f) : 1.0f;
		float4 mix = d.c.mix[i];
		mix.w = alpha;
		d.c.mix[i] = mix;
		d.c.flags[i] |= B2CU_CONTACT_TOI;
		atomicMin(reinterpret_cast<unsigned int*>(d.counters + CNT_TOI_MIN_ALPHA), __float_as_uint(alpha));
	}
}

// the winner among equal alphas: the smallest contact key (b2Contact::ToiLessThan, b2Contact.cpp:326-334)
__global__ void __launch_bounds__(256) ToiMinKeyAllKernel(DeviceArrays d, int contactCount)
{
	GridDependencyWait();
	const unsigned int best = *reinterpret_cast<const unsigned int*>(d.counters + CNT_TOI_MIN_ALPHA);
	if (best == 0xFFFFFFFFu) return;
	B2CU_GRID_STRIDE(i, contactCount)
	{
		uint32_t f = d.c.flags[i];
		if (!(f & B2CU_CONTACT_TOI) || !ToiEligible(d, i, f)) continue;
		if (__float_as_uint(d.c.mix[i].w) == best)
			atomicMin(reinterpret_cast<unsigned long long*>(d.counters + CNT_TOI_MIN_KEY), (unsigned long long)d.c.key[i]);
	}
}

// b2World::ClearPostSolveTOI (b2World.cpp:1467-1504): every contact forgets its time of impact and sub-step count,
// every body (static ones too) goes back to alpha0 = 0
__global__ void __launch_bounds__(256) ToiClearKernel(DeviceArrays d, int contactCount, int bodyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, (contactCount > bodyCount ? contactCount : bodyCount))
	{
		if (i < contactCount)
		{
			uint32_t f = d.c.flags[i];
			if (f & (B2CU_CONTACT_TOI | B2CU_CONTACT_ISLAND)) d.c.flags[i] = f & ~(uint32_t)(B2CU_CONTACT_TOI | B2CU_CONTACT_ISLAND);
			if (f & B2CU_CONTACT_TOI_CANDIDATE)
			{
				// only candidates are ever evaluated or sub-stepped
				if (d.c.toiCount[i] != 0) d.c.toiCount[i] = 0;
				float4 mix = d.c.mix[i];
				if (mix.w != 1.0f)
				{
					mix.w = 1.0f;
					d.c.mix[i] = mix;
				}
			}
		}
		if (i < bodyCount)
		{
			uint32_t bf = d.bflags[i];
			if (bf & B2CU_BODY_ISLAND) d.bflags[i] = bf & ~(uint32_t)B2CU_BODY_ISLAND;
			float4 p0 = d.pos0[i];
			if (p0.w != 0.0f)
			{
				p0.w = 0.0f;
				d.pos0[i] = p0;
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// StepSolveTOI
// ---------------------------------------------------------------------------------------------------------

// slot of a live contact by key: the sorted main region, then the sorted tail; -1 if there is none
__device__ __forceinline__ int FindContactSlot(const DeviceArrays& d, int contactCount, int mainCount, uint64_t key)
{
	int i = LowerBound64(d.c.key, mainCount, key);
	if (i < mainCount && d.c.key[i] == key && !(d.c.flags[i] & B2CU_CONTACT_DEAD)) return i;
	if (contactCount > mainCount)
	{
		i = mainCount + LowerBound64(d.c.key + mainCount, contactCount - mainCount, key);
		if (i < contactCount && d.c.key[i] == key && !(d.c.flags[i] & B2CU_CONTACT_DEAD)) return i;
	}
	return -1;
}

#define B2CU_TOI_LIST_SIDE (1ull << 63)
// state of one entry of the two contact lists during the island search of ToiEventKernel
#define B2CU_TOI_SORT_KEYS 4096
#define B2CU_TOI_WALK_TOUCHING 1 // its manifold at the time of impact has points
#define B2CU_TOI_WALK_SKIP 2     // already in the island (the event's own contact)
#define B2CU_TOI_WALK_VISITED 4  // the walk reaches it before the island is full
#define B2CU_TOI_WALK_BEGIN 8    // its update raises BeginContact
#define B2CU_TOI_WALK_END 16     // ... EndContact
__device__ __forceinline__ uint64_t ToiListKey(int side, uint32_t stamp, int slot)
{
	return ((uint64_t)side << 63) | ((uint64_t)(0xFFFFFFFFu - stamp) << 31) | (uint64_t)(0x7FFFFFFF - slot);
}
__device__ __forceinline__ int ToiListSlot(uint64_t k) { return 0x7FFFFFFF - (int)(k & 0x7FFFFFFFull); }

// The contact lists of the event's two bodies (b2Body::m_contactList), as far as the island search looks at them
// (b2World.cpp:899-931): only a dynamic body's list is walked; a contact with a dynamic body is skipped unless one of
// the two is a bullet; sensors are skipped.  A body's list is in reverse creation order (OnContactCreate links at the
// head, b2ContactManager.cpp:530-556): batches by stamp, and inside a batch by key -- which, inside the main region and
// inside the tail of the contact set, is the slot order (all contacts of the tail are younger than all of the main region).
__global__ void __launch_bounds__(256) ToiEventPrepareKernel(DeviceArrays d, int contactCount, int mainCount, uint64_t minKey,
                                                             int capacity)
{
	GridDependencyWait();
	__shared__ int sh[3];
	if (threadIdx.x == 0)
	{
		int i0 = FindContactSlot(d, contactCount, mainCount, minKey);
		sh[0] = i0;
		if (i0 >= 0)
		{
			int4 pr = d.c.proxies[i0];
			sh[1] = pr.z;
			sh[2] = pr.w;
		}
	}
	__syncthreads();
	const int i0 = sh[0];
	if (i0 < 0) return;
	const int body[2] = {sh[1], sh[2]};
	const uint32_t bf[2] = {d.bflags[body[0]], d.bflags[body[1]]};
	B2CU_GRID_STRIDE(i, contactCount)
	{
		if (i == i0) continue;
		uint32_t f = d.c.flags[i];
		if (f & (B2CU_CONTACT_DEAD | B2CU_CONTACT_SENSOR)) continue;
		int4 pr = d.c.proxies[i];
		for (int side = 0; side < 2; ++side)
		{
			if (!IsDynamic(bf[side])) continue;
			int other;
			if (pr.z == body[side]) other = pr.w;
			else if (pr.w == body[side]) other = pr.z;
			else continue;
			uint32_t fo = d.bflags[other];
			if (IsDynamic(fo) && !(bf[side] & B2CU_BODY_BULLET) && !(fo & B2CU_BODY_BULLET)) continue;
			int slot = atomicAdd(&d.counters[CNT_TOI_LIST], 1);
			if (slot < capacity) d.toiListKeys[slot] = ToiListKey(side, d.c.stamp[i], i);
			else d.counters[CNT_ERROR] = 1;
		}
	}
}

// b2Body::SetAwake(true) (b2Body.h:690-718): in this branch of Box2D the sleep timer is reset whether or not the body slept
__device__ __forceinline__ void ToiSetAwake(const DeviceArrays& d, int b)
{
	d.bflags[b] |= B2CU_BODY_AWAKE;
	float4 f = d.force[b];
	if (f.w != 0.0f)
	{
		f.w = 0.0f;
		d.force[b] = f;
	}
}

// the same from several threads at once (the parallel part of the island search): flag bits by atomic OR, the timer by a
// store of the one value every writer agrees on
__device__ __forceinline__ void ToiSetAwakeShared(const DeviceArrays& d, int b)
{
	atomicOr(&d.bflags[b], (uint32_t)B2CU_BODY_AWAKE);
	float* timer = reinterpret_cast<float*>(&d.force[b]) + 3;
	if (*timer != 0.0f) *timer = 0.0f;
}

// b2Body::SynchronizeTransform (b2Body.h:958-962)
__device__ __forceinline__ void ToiSyncTransform(const DeviceArrays& d, int b)
{
	float4 p = d.pos[b], ms = d.mass[b];
	Rot q = SinCos(p.z);
	Vec2 o = V(p.x, p.y) - Mul(q, V(ms.z, ms.w));
	d.xf[b] = make_float4(o.x, o.y, q.s, q.c);
}

// where b2Body::Advance(alpha) (b2Body.h:964-972) puts a body: the new sweep start, which is also its position
struct ToiAdvanced
{
	float4 pos0; // c0, a0, alpha0
	Xf xf;
};
__device__ __forceinline__ ToiAdvanced ToiAdvanceOf(const DeviceArrays& d, int b, float alpha)
{
	Sweep s = LoadSweep(d, b);
	SweepAdvance(s, alpha);
	ToiAdvanced r;
	r.pos0 = make_float4(s.c0.x, s.c0.y, s.a0, s.alpha0);
	r.xf.q = SinCos(s.a0);
	r.xf.p = s.c0 - Mul(r.xf.q, s.localCenter);
	return r;
}
__device__ __forceinline__ void ToiAdvanceBody(const DeviceArrays& d, int b, float alpha)
{
	ToiAdvanced r = ToiAdvanceOf(d, b, alpha);
	float4 p = d.pos[b];
	d.pos0[b] = r.pos0;
	d.pos[b] = make_float4(r.pos0.x, r.pos0.y, r.pos0.z, p.w);
	d.xf[b] = make_float4(r.xf.p.x, r.xf.p.y, r.xf.q.s, r.xf.q.c);
}

__device__ __forceinline__ void ToiAppendEvent(const DeviceArrays& d, int kind, uint64_t key, int capacity)
{
	int slot = atomicAdd(&d.counters[CNT_TOI_EVENTS], 1);
	if (slot < capacity)
	{
		d.toiEventKeys[slot] = key;
		d.toiEventKinds[slot] = kind;
	}
	else
	{
		d.counters[CNT_ERROR] = 1;
	}
}

// manifold of contact i as stored (what b2Contact::Evaluate starts from: the collide functions leave part of it alone)
__device__ __forceinline__ void ToiLoadManifold(const DeviceArrays& d, int i, Manifold& m)
{
	float4 o0 = d.c.m0[i], o1 = d.c.m1[i], o2 = d.c.m2[i];
	uint4 o3 = d.c.m3[i];
	m.localNormal = V(o0.x, o0.y);
	m.localPoint = V(o0.z, o0.w);
	m.lp[0] = V(o1.x, o1.y);
	m.lp[1] = V(o2.x, o2.y);
	m.ni[0] = m.ni[1] = m.ti[0] = m.ti[1] = 0.0f;
	m.id[0] = o3.x;
	m.id[1] = o3.y;
	m.type = (int)o3.z;
	m.pointCount = 0;
}

// The rest of b2Contact::Update (single-threaded flavour, b2Contact.cpp:205-281) once the new manifold geometry is known:
// impulses carried over by feature id, touching flag, both bodies woken when touching changed, begin / end events in
// call order.  Returns whether the contact touches now.
// kind of the event the update raises: -1 none, else B2CU_EVENT_BEGIN / B2CU_EVENT_END.  `shared`: other threads may be
// updating other contacts of the same bodies at the same time; `joins`: the contact enters the island being built.
__device__ __forceinline__ bool ToiCommitUpdateCore(const DeviceArrays& d, int i, int bA, int bB, Manifold& m, bool shared,
                                                    bool joins, int* eventKind)
{
	uint32_t flags = d.c.flags[i] | B2CU_CONTACT_ENABLED;
	const bool wasTouching = (flags & B2CU_CONTACT_TOUCHING) != 0;
	float4 o1 = d.c.m1[i], o2 = d.c.m2[i];
	uint4 o3 = d.c.m3[i];
	const int oldCount = (int)o3.w;
	const bool touching = m.pointCount > 0;
	for (int k = 0; k < m.pointCount; ++k)
	{
		float ni = 0.0f, ti = 0.0f;
		uint32_t id2 = m.id[k];
		if (oldCount > 0 && o3.x == id2)
		{
			ni = o1.z;
			ti = o1.w;
		}
		else if (oldCount > 1 && o3.y == id2)
		{
			ni = o2.z;
			ti = o2.w;
		}
		m.ni[k] = ni;
		m.ti[k] = ti;
	}
	for (int k = m.pointCount; k < 2; ++k)
	{
		m.ni[k] = k == 0 ? o1.z : o2.z;
		m.ti[k] = k == 0 ? o1.w : o2.w;
	}
	if (touching != wasTouching)
	{
		if (shared)
		{
			ToiSetAwakeShared(d, bA);
			ToiSetAwakeShared(d, bB);
		}
		else
		{
			ToiSetAwake(d, bA);
			ToiSetAwake(d, bB);
		}
	}
	if (touching) flags |= B2CU_CONTACT_TOUCHING;
	else flags &= ~(uint32_t)B2CU_CONTACT_TOUCHING;
	if (joins && touching) flags |= B2CU_CONTACT_ISLAND;
	*eventKind = touching == wasTouching ? -1 : touching ? B2CU_EVENT_BEGIN : B2CU_EVENT_END;
	d.c.flags[i] = flags;
	d.c.m0[i] = make_float4(m.localNormal.x, m.localNormal.y, m.localPoint.x, m.localPoint.y);
	d.c.m1[i] = make_float4(m.lp[0].x, m.lp[0].y, m.ni[0], m.ti[0]);
	d.c.m2[i] = make_float4(m.lp[1].x, m.lp[1].y, m.ni[1], m.ti[1]);
	d.c.m3[i] = make_uint4(m.id[0], m.id[1], (uint32_t)m.type, (uint32_t)m.pointCount);
	return touching;
}

__device__ __forceinline__ bool ToiCommitUpdate(const DeviceArrays& d, int i, int bA, int bB, Manifold& m, int capacity)
{
	int kind;
	const bool touching = ToiCommitUpdateCore(d, i, bA, bB, m, false, false, &kind);
	if (kind >= 0) ToiAppendEvent(d, kind, d.c.key[i], capacity);
	return touching;
}

// rows of the island's contact solver (b2ContactVelocityConstraint / b2ContactPositionConstraint, b2ContactSolver.cpp:
// 32-45, b2ContactSolver.h:30-58), kept in shared memory
struct ToiConstraint
{
	int indexA, indexB;
	int pointCount;      // of the manifold (position constraints)
	int velocityPoints;  // after the block solver's conditioning test (velocity constraints)
	int type;
	float mA, iA, mB, iB;
	float friction, restitution, tangentSpeed;
	float radiusA, radiusB;
	Vec2 localCenterA, localCenterB;
	Vec2 localNormal, localPoint;
	Vec2 localPoints[2];
	Vec2 normal;
	Vec2 rA[2], rB[2];
	float normalMass[2], tangentMass[2], velocityBias[2];
	float normalImpulse[2], tangentImpulse[2];
	float k11, k12, k22;     // K
	float nm[4];             // normalMass matrix: ex.x, ey.x, ex.y, ey.y
};

struct ToiIsland
{
	int bodyCount, contactCount;
	int bodies[B2CU_TOI_MAX_BODIES];
	int contacts[B2CU_TOI_MAX_CONTACTS];
	float cx[B2CU_TOI_MAX_BODIES], cy[B2CU_TOI_MAX_BODIES], a[B2CU_TOI_MAX_BODIES];
	float vx[B2CU_TOI_MAX_BODIES], vy[B2CU_TOI_MAX_BODIES], w[B2CU_TOI_MAX_BODIES];
	float qs[B2CU_TOI_MAX_BODIES], qc[B2CU_TOI_MAX_BODIES]; // sin / cos of a[]: of the bodies the position solver cannot move
	ToiConstraint rows[B2CU_TOI_MAX_CONTACTS];
};

__device__ __forceinline__ int ToiIslandIndex(const ToiIsland& is, int body)
{
	for (int k = 0; k < is.bodyCount; ++k)
		if (is.bodies[k] == body) return k;
	return -1;
}

// b2PositionSolverManifold::Initialize (b2ContactSolver.cpp:620-673)
__device__ __forceinline__ void ToiPositionManifold(const ToiConstraint& pc, const Xf& xfA, const Xf& xfB, int index,
                                                    Vec2* normal, Vec2* point, float* separation)
{
	if (pc.type == B2CU_MANIFOLD_CIRCLES)
	{
		Vec2 pointA = Mul(xfA, pc.localPoint);
		Vec2 pointB = Mul(xfB, pc.localPoints[0]);
		*normal = Normalized(pointB - pointA);
		*point = 0.5f * (pointA + pointB);
		*separation = Dot(pointB - pointA, *normal) - pc.radiusA - pc.radiusB;
	}
	else if (pc.type == B2CU_MANIFOLD_FACE_A)
	{
		*normal = Mul(xfA.q, pc.localNormal);
		Vec2 planePoint = Mul(xfA, pc.localPoint);
		Vec2 clipPoint = Mul(xfB, pc.localPoints[index]);
		*separation = Dot(clipPoint - planePoint, *normal) - pc.radiusA - pc.radiusB;
		*point = clipPoint;
	}
	else
	{
		*normal = Mul(xfB.q, pc.localNormal);
		Vec2 planePoint = Mul(xfB, pc.localPoint);
		Vec2 clipPoint = Mul(xfA, pc.localPoints[index]);
		*separation = Dot(clipPoint - planePoint, *normal) - pc.radiusA - pc.radiusB;
		*point = clipPoint;
		*normal = -(*normal);
	}
}

// b2ContactSolver::SolveTOIPositionConstraints (b2ContactSolver.cpp:755-843): only the two bodies of the event have mass
__device__ __forceinline__ bool ToiSolvePositions(ToiIsland& is, int toiIndexA, int toiIndexB)
{
	float minSeparation = 0.0f;
	for (int i = 0; i < is.contactCount; ++i)
	{
		const ToiConstraint& pc = is.rows[i];
		const int indexA = pc.indexA, indexB = pc.indexB;
		float mA = 0.0f, iA = 0.0f;
		if (indexA == toiIndexA || indexA == toiIndexB)
		{
			mA = pc.mA;
			iA = pc.iA;
		}
		float mB = 0.0f, iB = 0.0f;
		if (indexB == toiIndexA || indexB == toiIndexB)
		{
			mB = pc.mB;
			iB = pc.iB;
		}
		Vec2 cA = V(is.cx[indexA], is.cy[indexA]);
		float aA = is.a[indexA];
		Vec2 cB = V(is.cx[indexB], is.cy[indexB]);
		float aB = is.a[indexB];
		const bool movesA = indexA == toiIndexA || indexA == toiIndexB, movesB = indexB == toiIndexA || indexB == toiIndexB;
		for (int j = 0; j < pc.pointCount; ++j)
		{
			Xf xfA, xfB;
			// a body without mass here keeps its angle through all iterations: its rotation was evaluated once
			if (movesA) xfA.q = SinCos(aA);
			else
			{
				xfA.q.s = is.qs[indexA];
				xfA.q.c = is.qc[indexA];
			}
			if (movesB) xfB.q = SinCos(aB);
			else
			{
				xfB.q.s = is.qs[indexB];
				xfB.q.c = is.qc[indexB];
			}
			xfA.p = cA - Mul(xfA.q, pc.localCenterA);
			xfB.p = cB - Mul(xfB.q, pc.localCenterB);
			Vec2 normal, point;
			float separation;
			ToiPositionManifold(pc, xfA, xfB, j, &normal, &point, &separation);
			Vec2 rA = point - cA;
			Vec2 rB = point - cB;
			minSeparation = Min(minSeparation, separation);
			float C = Clamp(B2CU_TOI_BAUMGARTE * (separation + B2CU_LINEAR_SLOP), -B2CU_MAX_LINEAR_CORRECTION, 0.0f);
			float rnA = Cross(rA, normal);
			float rnB = Cross(rB, normal);
			float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
			float impulse = K > 0.0f ? -C / K : 0.0f;
			Vec2 P = impulse * normal;
			cA = cA - mA * P;
			aA -= iA * Cross(rA, P);
			cB = cB + mB * P;
			aB += iB * Cross(rB, P);
		}
		is.cx[indexA] = cA.x;
		is.cy[indexA] = cA.y;
		is.a[indexA] = aA;
		is.cx[indexB] = cB.x;
		is.cy[indexB] = cB.y;
		is.a[indexB] = aB;
	}
	return minSeparation >= -1.5f * B2CU_LINEAR_SLOP;
}

// b2ContactSolver::InitializeVelocityConstraints (b2ContactSolver.cpp:142-251) with b2WorldManifold::Initialize
// (b2Collision.cpp:22-86) for one row
__device__ __forceinline__ void ToiInitVelocityRow(const ToiIsland& is, ToiConstraint& vc)
{
	const float mA = vc.mA, iA = vc.iA, mB = vc.mB, iB = vc.iB;
	Vec2 cA = V(is.cx[vc.indexA], is.cy[vc.indexA]);
	float aA = is.a[vc.indexA];
	Vec2 vA = V(is.vx[vc.indexA], is.vy[vc.indexA]);
	float wA = is.w[vc.indexA];
	Vec2 cB = V(is.cx[vc.indexB], is.cy[vc.indexB]);
	float aB = is.a[vc.indexB];
	Vec2 vB = V(is.vx[vc.indexB], is.vy[vc.indexB]);
	float wB = is.w[vc.indexB];

	Xf xfA, xfB;
	xfA.q = SinCos(aA);
	xfB.q = SinCos(aB);
	xfA.p = cA - Mul(xfA.q, vc.localCenterA);
	xfB.p = cB - Mul(xfB.q, vc.localCenterB);

	const int pointCount = vc.pointCount;
	Vec2 normal = V(1.0f, 0.0f);
	Vec2 wp[2] = {V(0.0f, 0.0f), V(0.0f, 0.0f)};
	if (vc.type == B2CU_MANIFOLD_CIRCLES)
	{
		Vec2 pointA = Mul(xfA, vc.localPoint);
		Vec2 pointB = Mul(xfB, vc.localPoints[0]);
		if (DistanceSquared(pointA, pointB) > B2CU_EPSILON * B2CU_EPSILON)
		{
			normal = Normalized(pointB - pointA);
		}
		Vec2 ccA = pointA + vc.radiusA * normal;
		Vec2 ccB = pointB - vc.radiusB * normal;
		wp[0] = 0.5f * (ccA + ccB);
	}
	else if (vc.type == B2CU_MANIFOLD_FACE_A)
	{
		normal = Mul(xfA.q, vc.localNormal);
		Vec2 planePoint = Mul(xfA, vc.localPoint);
		for (int j = 0; j < pointCount; ++j)
		{
			Vec2 clipPoint = Mul(xfB, vc.localPoints[j]);
			Vec2 ccA = clipPoint + (vc.radiusA - Dot(clipPoint - planePoint, normal)) * normal;
			Vec2 ccB = clipPoint - vc.radiusB * normal;
			wp[j] = 0.5f * (ccA + ccB);
		}
	}
	else
	{
		normal = Mul(xfB.q, vc.localNormal);
		Vec2 planePoint = Mul(xfB, vc.localPoint);
		for (int j = 0; j < pointCount; ++j)
		{
			Vec2 clipPoint = Mul(xfA, vc.localPoints[j]);
			Vec2 ccB = clipPoint + (vc.radiusB - Dot(clipPoint - planePoint, normal)) * normal;
			Vec2 ccA = clipPoint - vc.radiusA * normal;
			wp[j] = 0.5f * (ccA + ccB);
		}
		normal = -normal;
	}
	vc.normal = normal;
	vc.velocityPoints = pointCount;
	for (int j = 0; j < pointCount; ++j)
	{
		Vec2 rA = wp[j] - cA;
		Vec2 rB = wp[j] - cB;
		vc.rA[j] = rA;
		vc.rB[j] = rB;
		float rnA = Cross(rA, normal);
		float rnB = Cross(rB, normal);
		float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
		vc.normalMass[j] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
		Vec2 tangent = CrossVS(normal, 1.0f);
		float rtA = Cross(rA, tangent);
		float rtB = Cross(rB, tangent);
		float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
		vc.tangentMass[j] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
		vc.velocityBias[j] = 0.0f;
		float vRel = Dot(normal, vB + CrossSV(wB, rB) - vA - CrossSV(wA, rA));
		if (vRel < -B2CU_VELOCITY_THRESHOLD)
		{
			vc.velocityBias[j] = -vc.restitution * vRel;
		}
	}
	if (pointCount == 2)
	{
		float rn1A = Cross(vc.rA[0], normal);
		float rn1B = Cross(vc.rB[0], normal);
		float rn2A = Cross(vc.rA[1], normal);
		float rn2B = Cross(vc.rB[1], normal);
		float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
		float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
		float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
		const float k_maxConditionNumber = 1000.0f;
		if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12))
		{
			vc.k11 = k11;
			vc.k12 = k12;
			vc.k22 = k22;
			// b2Mat22::GetInverse (b2Math.h:205-216)
			float a = k11, b = k12, c = k12, dd = k22;
			float det = a * dd - b * c;
			if (det != 0.0f)
			{
				det = 1.0f / det;
			}
			vc.nm[0] = det * dd;
			vc.nm[1] = -det * b;
			vc.nm[2] = -det * c;
			vc.nm[3] = det * a;
		}
		else
		{
			vc.velocityPoints = 1;
		}
	}
}

// b2ContactSolver::SolveVelocityConstraints (b2ContactSolver.cpp:293-603) for one row of the island
__device__ __forceinline__ void ToiSolveVelocityRow(ToiIsland& is, ToiConstraint& vc)
{
	const float mA = vc.mA, iA = vc.iA, mB = vc.mB, iB = vc.iB;
	const int pointCount = vc.velocityPoints;
	Vec2 vA = V(is.vx[vc.indexA], is.vy[vc.indexA]);
	float wA = is.w[vc.indexA];
	Vec2 vB = V(is.vx[vc.indexB], is.vy[vc.indexB]);
	float wB = is.w[vc.indexB];
	const Vec2 normal = vc.normal;
	const Vec2 tangent = CrossVS(normal, 1.0f);
	const float friction = vc.friction;

	for (int j = 0; j < pointCount; ++j)
	{
		Vec2 rA = vc.rA[j], rB = vc.rB[j];
		Vec2 dv = vB + CrossSV(wB, rB) - vA - CrossSV(wA, rA);
		float vt = Dot(dv, tangent) - vc.tangentSpeed;
		float lambda = vc.tangentMass[j] * (-vt);
		float maxFriction = friction * vc.normalImpulse[j];
		float newImpulse = Clamp(vc.tangentImpulse[j] + lambda, -maxFriction, maxFriction);
		lambda = newImpulse - vc.tangentImpulse[j];
		vc.tangentImpulse[j] = newImpulse;
		Vec2 P = lambda * tangent;
		vA = vA - mA * P;
		wA -= iA * Cross(rA, P);
		vB = vB + mB * P;
		wB += iB * Cross(rB, P);
	}

	if (pointCount == 1)
	{
		Vec2 rA = vc.rA[0], rB = vc.rB[0];
		Vec2 dv = vB + CrossSV(wB, rB) - vA - CrossSV(wA, rA);
		float vn = Dot(dv, normal);
		float lambda = -vc.normalMass[0] * (vn - vc.velocityBias[0]);
		float newImpulse = Max(vc.normalImpulse[0] + lambda, 0.0f);
		lambda = newImpulse - vc.normalImpulse[0];
		vc.normalImpulse[0] = newImpulse;
		Vec2 P = lambda * normal;
		vA = vA - mA * P;
		wA -= iA * Cross(rA, P);
		vB = vB + mB * P;
		wB += iB * Cross(rB, P);
	}
	else if (pointCount == 2)
	{
		Vec2 rA1 = vc.rA[0], rB1 = vc.rB[0];
		Vec2 rA2 = vc.rA[1], rB2 = vc.rB[1];
		Vec2 a = V(vc.normalImpulse[0], vc.normalImpulse[1]);
		Vec2 dv1 = vB + CrossSV(wB, rB1) - vA - CrossSV(wA, rA1);
		Vec2 dv2 = vB + CrossSV(wB, rB2) - vA - CrossSV(wA, rA2);
		float vn1 = Dot(dv1, normal);
		float vn2 = Dot(dv2, normal);
		Vec2 b;
		b.x = vn1 - vc.velocityBias[0];
		b.y = vn2 - vc.velocityBias[1];
		b.x -= vc.k11 * a.x + vc.k12 * a.y;
		b.y -= vc.k12 * a.x + vc.k22 * a.y;

		Vec2 x;
		bool found = false;
		x.x = -(vc.nm[0] * b.x + vc.nm[1] * b.y);
		x.y = -(vc.nm[2] * b.x + vc.nm[3] * b.y);
		if (x.x >= 0.0f && x.y >= 0.0f)
		{
			found = true;
		}
		if (!found)
		{
			x.x = -vc.normalMass[0] * b.x;
			x.y = 0.0f;
			vn2 = vc.k12 * x.x + b.y;
			if (x.x >= 0.0f && vn2 >= 0.0f) found = true;
		}
		if (!found)
		{
			x.x = 0.0f;
			x.y = -vc.normalMass[1] * b.y;
			vn1 = vc.k12 * x.y + b.x;
			if (x.y >= 0.0f && vn1 >= 0.0f) found = true;
		}
		if (!found)
		{
			x.x = 0.0f;
			x.y = 0.0f;
			vn1 = b.x;
			vn2 = b.y;
			if (vn1 >= 0.0f && vn2 >= 0.0f) found = true;
		}
		if (found)
		{
			Vec2 dd = x - a;
			Vec2 P1 = dd.x * normal;
			Vec2 P2 = dd.y * normal;
			vA = vA - mA * (P1 + P2);
			wA -= iA * (Cross(rA1, P1) + Cross(rA2, P2));
			vB = vB + mB * (P1 + P2);
			wB += iB * (Cross(rB1, P1) + Cross(rB2, P2));
			vc.normalImpulse[0] = x.x;
			vc.normalImpulse[1] = x.y;
		}
	}

	is.vx[vc.indexA] = vA.x;
	is.vy[vc.indexA] = vA.y;
	is.w[vc.indexA] = wA;
	is.vx[vc.indexB] = vB.x;
	is.vy[vc.indexB] = vB.y;
	is.w[vc.indexB] = wB;
}

// b2ContactSolver::b2ContactSolver (b2ContactSolver.cpp:47-133) for contact slot i as row `row`; no warm starting in a
// time-of-impact island (b2World.cpp:984)
__device__ __forceinline__ void ToiMakeRow(const DeviceArrays& d, const ToiIsland& is, int i, ToiConstraint& r)
{
	int4 pr = d.c.proxies[i];
	float4 m0 = d.c.m0[i], m1 = d.c.m1[i], m2 = d.c.m2[i], mix = d.c.mix[i];
	uint4 m3 = d.c.m3[i];
	float4 msA = d.mass[pr.z], msB = d.mass[pr.w];
	r.indexA = ToiIslandIndex(is, pr.z);
	r.indexB = ToiIslandIndex(is, pr.w);
	r.pointCount = (int)m3.w;
	r.velocityPoints = r.pointCount;
	r.type = (int)m3.z;
	r.mA = msA.x;
	r.iA = msA.y;
	r.mB = msB.x;
	r.iB = msB.y;
	r.friction = mix.x;
	r.restitution = mix.y;
	r.tangentSpeed = mix.z;
	r.radiusA = d.pradius[pr.x];
	r.radiusB = d.pradius[pr.y];
	r.localCenterA = V(msA.z, msA.w);
	r.localCenterB = V(msB.z, msB.w);
	r.localNormal = V(m0.x, m0.y);
	r.localPoint = V(m0.z, m0.w);
	r.localPoints[0] = V(m1.x, m1.y);
	r.localPoints[1] = V(m2.x, m2.y);
	r.normal = V(0.0f, 0.0f);
	for (int j = 0; j < 2; ++j)
	{
		r.rA[j] = r.rB[j] = V(0.0f, 0.0f);
		r.normalMass[j] = r.tangentMass[j] = r.velocityBias[j] = 0.0f;
		r.normalImpulse[j] = r.tangentImpulse[j] = 0.0f;
	}
	r.k11 = r.k12 = r.k22 = 0.0f;
	r.nm[0] = r.nm[1] = r.nm[2] = r.nm[3] = 0.0f;
}

// One time-of-impact event.  One CTA; thread 0 does what the reference does in sequence, the other threads help where
// the work is independent per contact.
__global__ void __launch_bounds__(B2CU_TOI_THREADS) ToiEventKernel(DeviceArrays d, int contactCount, int mainCount,
                                                                   uint64_t minKey, float minAlpha, float dt,
                                                                   int velocityIterations, int capacity)
{
	GridDependencyWait();
	__shared__ ToiIsland is;
	__shared__ int shI0, shSolid, shBodyA, shBodyB;
	const int tid = threadIdx.x;
	const int listCount = min(d.counters[CNT_TOI_LIST], capacity);
#ifdef B2CU_TOI_PROFILE
	long long stamps[6];
#define B2CU_TOI_STAMP(k) stamps[k] = clock64()
#else
#define B2CU_TOI_STAMP(k)
#endif
	// sorted list keys and, per list entry, B2CU_TOI_WALK_* (long lists: in global memory; the candidate list of the
	// pass, listA, is spent by now)
	__shared__ uint64_t shKeys[B2CU_TOI_SORT_KEYS];
	__shared__ __align__(8) uint8_t shWalk[B2CU_TOI_SORT_KEYS];
	__shared__ int shSideStart, shWalkEnd[2];

	if (tid == 0)
	{
		shSolid = 0;
		shSideStart = 0x7FFFFFFF;
		d.toiScratch[B2CU_TOI_SCR_SOLID] = 0;
		d.toiScratch[B2CU_TOI_SCR_BODY_COUNT] = 0;
		const int i0 = FindContactSlot(d, contactCount, mainCount, minKey);
		shI0 = i0;
		if (i0 >= 0)
		{
			int4 pr = d.c.proxies[i0];
			const int bA = pr.z, bB = pr.w;
			shBodyA = bA;
			shBodyB = bB;
			// b2World.cpp:859-868: keep the sweeps, advance both bodies to the time of impact, update the contact there
			const float4 backupPos0A = d.pos0[bA], backupPosA = d.pos[bA];
			const float4 backupPos0B = d.pos0[bB], backupPosB = d.pos[bB];
			ToiAdvanceBody(d, bA, minAlpha);
			ToiAdvanceBody(d, bB, minAlpha);
			Manifold m;
			ToiLoadManifold(d, i0, m);
			Evaluate(&m, d.shapes + d.pshape[pr.x], MakeXf(d.xf[bA]), d.shapes + d.pshape[pr.y], MakeXf(d.xf[bB]));
			const bool touching = ToiCommitUpdate(d, i0, bA, bB, m, capacity);
			d.c.flags[i0] &= ~(uint32_t)B2CU_CONTACT_TOI;
			d.c.toiCount[i0] += 1;
			if (!touching)
			{
				// not solid: switch the contact off for the rest of the step and put the bodies back (:871-880)
				d.c.flags[i0] &= ~(uint32_t)B2CU_CONTACT_ENABLED;
				d.pos0[bA] = backupPos0A;
				d.pos[bA] = backupPosA;
				d.pos0[bB] = backupPos0B;
				d.pos[bB] = backupPosB;
				ToiSyncTransform(d, bA);
				ToiSyncTransform(d, bB);
			}
			else
			{
				ToiSetAwake(d, bA);
				ToiSetAwake(d, bB);
				is.bodyCount = 2;
				is.bodies[0] = bA;
				is.bodies[1] = bB;
				is.contactCount = 1;
				is.contacts[0] = i0;
				d.bflags[bA] |= B2CU_BODY_ISLAND;
				d.bflags[bB] |= B2CU_BODY_ISLAND;
				d.c.flags[i0] |= B2CU_CONTACT_ISLAND;
				shSolid = 1;
			}
		}
	}
	__syncthreads();
	if (shI0 < 0 || !shSolid) return;
	const int bodyA = shBodyA, bodyB = shBodyB;
	B2CU_TOI_STAMP(0);

	// ---- the two contact lists in list order: sort of the (unique) keys.  Up to B2CU_TOI_SORT_KEYS entries in shared
	// memory (bitonic network); longer lists by counting ranks in global memory ----
	const bool inShared = listCount <= B2CU_TOI_SORT_KEYS;
	if (inShared)
	{
		int padded = 2;
		while (padded < listCount) padded <<= 1;
		for (int e = tid; e < padded; e += B2CU_TOI_THREADS) shKeys[e] = e < listCount ? d.toiListKeys[e] : ~0ull;
		__syncthreads();
		for (int k = 2; k <= padded; k <<= 1)
		{
			for (int j = k >> 1; j > 0; j >>= 1)
			{
				for (int t = tid; t < (padded >> 1); t += B2CU_TOI_THREADS)
				{
					const int lo = 2 * t - (t & (j - 1)); // the element of the pair with bit j clear
					const int hi = lo + j;
					const uint64_t a = shKeys[lo], b = shKeys[hi];
					const bool ascending = (lo & k) == 0;
					if ((a > b) == ascending)
					{
						shKeys[lo] = b;
						shKeys[hi] = a;
					}
				}
				__syncthreads();
			}
		}
	}
	else
	{
		for (int e = tid; e < listCount; e += B2CU_TOI_THREADS)
		{
			const uint64_t key = d.toiListKeys[e];
			int rank = 0;
			for (int k = 0; k < listCount; ++k) rank += d.toiListKeys[k] < key ? 1 : 0;
			d.toiListSorted[rank] = key;
		}
		__syncthreads();
	}
	const uint64_t* sorted = inShared ? shKeys : d.toiListSorted;
	uint8_t* walk = inShared ? shWalk : reinterpret_cast<uint8_t*>(d.listA);
	B2CU_TOI_STAMP(1);

	// ---- manifolds of the listed contacts with the other body advanced to the time of impact (b2World.cpp:933-941).
	// Independent of the order of the walk: a body that is already in the island sits at that time already, and
	// advancing a sweep to the time it is at changes nothing.  Results wait in cAlt rows until the walk commits them. ----
	for (int e = tid; e < listCount; e += B2CU_TOI_THREADS)
	{
		const uint64_t key = sorted[e];
		const int i = ToiListSlot(key);
		const int body = (key & B2CU_TOI_LIST_SIDE) ? bodyB : bodyA;
		int4 pr = d.c.proxies[i];
		const int other = pr.z == body ? pr.w : pr.z;
		Xf xfBody = MakeXf(d.xf[body]);
		Xf xfOther = (other == bodyA || other == bodyB) ? MakeXf(d.xf[other]) : ToiAdvanceOf(d, other, minAlpha).xf;
		Manifold m;
		ToiLoadManifold(d, i, m);
		Evaluate(&m, d.shapes + d.pshape[pr.x], pr.z == body ? xfBody : xfOther, d.shapes + d.pshape[pr.y],
		         pr.z == body ? xfOther : xfBody);
		d.cAlt.m0[e] = make_float4(m.localNormal.x, m.localNormal.y, m.localPoint.x, m.localPoint.y);
		d.cAlt.m1[e] = make_float4(m.lp[0].x, m.lp[0].y, 0.0f, 0.0f);
		d.cAlt.m2[e] = make_float4(m.lp[1].x, m.lp[1].y, 0.0f, 0.0f);
		d.cAlt.m3[e] = make_uint4(m.id[0], m.id[1], (uint32_t)m.type, (uint32_t)m.pointCount);
		walk[e] = (uint8_t)((m.pointCount > 0 ? B2CU_TOI_WALK_TOUCHING : 0) |
		                    ((d.c.flags[i] & B2CU_CONTACT_ISLAND) ? B2CU_TOI_WALK_SKIP : 0));
		// where the second body's list begins
		if ((key & B2CU_TOI_LIST_SIDE) && (e == 0 || !(sorted[e - 1] & B2CU_TOI_LIST_SIDE))) shSideStart = e;
	}
	__syncthreads();
	B2CU_TOI_STAMP(2);

	// ---- the walk (b2World.cpp:897-970), in three parts.  Which contacts the walk looks at, which of them join the
	// island and which bodies come with them depends on the list order (the island stops growing at 32 contacts / 64
	// bodies), but only through the touching states that are already known: one thread settles that on the flags ... ----
	if (tid == 0)
	{
		const int sideStart = min(shSideStart, listCount);
		for (int side = 0; side < 2; ++side)
		{
			int e = side ? sideStart : 0;
			const int end = side ? listCount : sideStart;
			for (; e < end; ++e)
			{
				// a full island ends this body's loop (:905-909, :916-919)
				if (is.bodyCount == B2CU_TOI_MAX_BODIES || is.contactCount == B2CU_TOI_MAX_CONTACTS) break;
				if (inShared && (e & 7) == 0 && e + 8 <= end)
				{
					// eight entries at once while none of them touches or is to be skipped: they are visited, nothing else
					uint64_t* group = reinterpret_cast<uint64_t*>(shWalk + e);
					const uint64_t g = *group;
					if ((g & 0x0303030303030303ull) == 0)
					{
						*group = g | 0x0404040404040404ull;
						e += 7;
						continue;
					}
				}
				int info = walk[e];
				if (info & B2CU_TOI_WALK_SKIP) continue;
				info |= B2CU_TOI_WALK_VISITED;
				walk[e] = (uint8_t)info;
				if (!(info & B2CU_TOI_WALK_TOUCHING)) continue;
				const int i = ToiListSlot(sorted[e]);
				is.contacts[is.contactCount++] = i;
				const int body = side ? bodyB : bodyA;
				int4 pr = d.c.proxies[i];
				const int other = pr.z == body ? pr.w : pr.z;
				bool inIsland = false;
				for (int k = 0; k < is.bodyCount; ++k) inIsland |= is.bodies[k] == other;
				if (!inIsland) is.bodies[is.bodyCount++] = other;
			}
			shWalkEnd[side] = e;
		}
	}
	__syncthreads();
	B2CU_TOI_STAMP(3);
	// ... then every visited contact is updated (b2Contact::Update with the other body advanced; a contact that does not
	// touch leaves that body where it was), and every body that joined is advanced, flagged and woken (:933-968) ...
	for (int side = 0; side < 2; ++side)
	{
		const int begin = side ? min(shSideStart, listCount) : 0;
		for (int e = begin + tid; e < shWalkEnd[side]; e += B2CU_TOI_THREADS)
		{
			int info = walk[e];
			if (!(info & B2CU_TOI_WALK_VISITED)) continue;
			const int i = ToiListSlot(sorted[e]);
			int4 pr = d.c.proxies[i];
			Manifold m;
			{
				float4 g0 = d.cAlt.m0[e], g1 = d.cAlt.m1[e], g2 = d.cAlt.m2[e];
				uint4 g3 = d.cAlt.m3[e];
				m.localNormal = V(g0.x, g0.y);
				m.localPoint = V(g0.z, g0.w);
				m.lp[0] = V(g1.x, g1.y);
				m.lp[1] = V(g2.x, g2.y);
				m.id[0] = g3.x;
				m.id[1] = g3.y;
				m.type = (int)g3.z;
				m.pointCount = (int)g3.w;
				m.ni[0] = m.ni[1] = m.ti[0] = m.ti[1] = 0.0f;
			}
			int kind;
			ToiCommitUpdateCore(d, i, pr.z, pr.w, m, true, true, &kind);
			if (kind >= 0) walk[e] = (uint8_t)(info | (kind == B2CU_EVENT_BEGIN ? B2CU_TOI_WALK_BEGIN : B2CU_TOI_WALK_END));
		}
	}
	for (int k = 2 + tid; k < is.bodyCount; k += B2CU_TOI_THREADS)
	{
		const int other = is.bodies[k];
		ToiAdvanceBody(d, other, minAlpha);
		const uint32_t fo = atomicOr(&d.bflags[other], (uint32_t)B2CU_BODY_ISLAND);
		if (!IsStatic(fo)) ToiSetAwakeShared(d, other);
	}
	__syncthreads();
	B2CU_TOI_STAMP(4);

	// ... and the begin / end events are raised in the order of the walk
	if (tid == 0)
	{
		for (int side = 0; side < 2; ++side)
			for (int e = side ? min(shSideStart, listCount) : 0; e < shWalkEnd[side]; ++e)
			{
				const int info = walk[e];
				if (info & (B2CU_TOI_WALK_BEGIN | B2CU_TOI_WALK_END))
					ToiAppendEvent(d, (info & B2CU_TOI_WALK_BEGIN) ? B2CU_EVENT_BEGIN : B2CU_EVENT_END,
					               d.c.key[ToiListSlot(sorted[e])], capacity);
			}
	}

	// ---- b2Island::SolveTOI (b2Island.cpp:398-530): bodies and rows are set up by all threads, the sequential
	// impulses are one thread's ----
	for (int k = tid; k < is.bodyCount; k += B2CU_TOI_THREADS)
	{
		const int b = is.bodies[k];
		float4 p = d.pos[b], v = d.vel[b];
		is.cx[k] = p.x;
		is.cy[k] = p.y;
		is.a[k] = p.z;
		is.vx[k] = v.x;
		is.vy[k] = v.y;
		is.w[k] = v.z;
		Rot q = SinCos(p.z);
		is.qs[k] = q.s;
		is.qc[k] = q.c;
	}
	__syncthreads();
	for (int k = tid; k < is.contactCount; k += B2CU_TOI_THREADS) ToiMakeRow(d, is, is.contacts[k], is.rows[k]);
	__syncthreads();
	if (tid == 0)
	{
		const int toiIndexA = 0, toiIndexB = 1;
		for (int it = 0; it < 20; ++it)
		{
			if (ToiSolvePositions(is, toiIndexA, toiIndexB)) break;
		}
		// leap of faith to the new safe state (:470-474)
		{
			float4 p0 = d.pos0[bodyA];
			d.pos0[bodyA] = make_float4(is.cx[toiIndexA], is.cy[toiIndexA], is.a[toiIndexA], p0.w);
			p0 = d.pos0[bodyB];
			d.pos0[bodyB] = make_float4(is.cx[toiIndexB], is.cy[toiIndexB], is.a[toiIndexB], p0.w);
		}
	}
	__syncthreads();
	for (int k = tid; k < is.contactCount; k += B2CU_TOI_THREADS) ToiInitVelocityRow(is, is.rows[k]);
	__syncthreads();
	if (tid == 0)
	{
		for (int it = 0; it < velocityIterations; ++it)
		{
			for (int k = 0; k < is.contactCount; ++k) ToiSolveVelocityRow(is, is.rows[k]);
		}
	}
	__syncthreads();
	// the impulses of a time-of-impact solve are not stored for warm starting (:488-489)
	const float h = (1.0f - minAlpha) * dt;
	for (int k = tid; k < is.bodyCount; k += B2CU_TOI_THREADS)
	{
		Vec2 c = V(is.cx[k], is.cy[k]);
		float a = is.a[k];
		Vec2 v = V(is.vx[k], is.vy[k]);
		float w = is.w[k];
		Vec2 translation = h * v;
		if (Dot(translation, translation) > B2CU_MAX_TRANSLATION_SQUARED)
		{
			float ratio = B2CU_MAX_TRANSLATION / Length(translation);
			v = V(v.x * ratio, v.y * ratio);
		}
		float rotation = h * w;
		if (rotation * rotation > B2CU_MAX_ROTATION_SQUARED)
		{
			float ratio = B2CU_MAX_ROTATION / Abs(rotation);
			w *= ratio;
		}
		c = c + h * v;
		a += h * w;
		const int b = is.bodies[k];
		float4 p = d.pos[b], v4 = d.vel[b];
		d.pos[b] = make_float4(c.x, c.y, a, p.w);
		d.vel[b] = make_float4(v.x, v.y, w, v4.w);
		ToiSyncTransform(d, b);
		d.toiScratch[B2CU_TOI_SCR_BODIES + k] = b;
	}
	if (tid == 0)
	{
		d.toiScratch[B2CU_TOI_SCR_BODY_COUNT] = is.bodyCount;
		d.toiScratch[B2CU_TOI_SCR_SOLID] = 1;
#ifdef B2CU_TOI_PROFILE
		stamps[5] = clock64();
		printf("toi event: list %d island %d/%d  sort %lld spec %lld scan %lld commit %lld solve %lld cycles\n", listCount, is.contactCount,
		       is.bodyCount, stamps[1] - stamps[0], stamps[2] - stamps[1], stamps[3] - stamps[2], stamps[4] - stamps[3], stamps[5] - stamps[4]);
#endif
	}
}

// After the island solve (b2World.cpp:995-1013): the displaced dynamic bodies synchronise their fixtures (b2Body::
// SynchronizeFixtures, b2Body.cpp:475-489; same fat-box rule as SyncProxiesKernel) and all their contacts lose
// e_islandFlag and e_toiFlag, so that the next FindMinToiContact recomputes them.
__global__ void __launch_bounds__(256) ToiAfterEventKernel(DeviceArrays d, int proxyCount, int contactCount)
{
	GridDependencyWait();
	if (!d.toiScratch[B2CU_TOI_SCR_SOLID]) return;
	B2CU_GRID_STRIDE(t, (proxyCount > contactCount ? proxyCount : contactCount))
	{
		if (t < proxyCount)
		{
			const int p = t;
			const int b = d.pbody[p];
			const uint32_t bf = d.bflags[b];
			uint32_t g = d.pgroup[p];
			if (IsDynamic(bf) && (bf & B2CU_BODY_ISLAND) && !((g >> 16) & B2CU_PROXY_INACTIVE))
			{
				float4 p0 = d.pos0[b];
				float4 ms = d.mass[b];
				Xf xf1;
				xf1.q = SinCos(p0.z);
				xf1.p = V(p0.x, p0.y) - Mul(xf1.q, V(ms.z, ms.w));
				Xf xf2 = MakeXf(d.xf[b]);
				const b2cuShape* s = d.shapes + d.pshape[p];
				float4 a1 = ComputeAABB(s, xf1);
				float4 a2 = ComputeAABB(s, xf2);
				float4 ab = make_float4(Min(a1.x, a2.x), Min(a1.y, a2.y), Max(a1.z, a2.z), Max(a1.w, a2.w));
				d.aabb[p] = ab;
				float4 fat = d.fat[p];
				bool contains = fat.x <= ab.x && fat.y <= ab.y && ab.z <= fat.z && ab.w <= fat.w;
				if (!contains)
				{
					Vec2 disp = xf2.p - xf1.p;
					float4 nb = make_float4(ab.x - B2CU_AABB_EXTENSION, ab.y - B2CU_AABB_EXTENSION, ab.z + B2CU_AABB_EXTENSION,
					                        ab.w + B2CU_AABB_EXTENSION);
					Vec2 dd = B2CU_AABB_MULTIPLIER * disp;
					if (dd.x < 0.0f) nb.x += dd.x;
					else nb.z += dd.x;
					if (dd.y < 0.0f) nb.y += dd.y;
					else nb.w += dd.y;
					d.fat[p] = nb;
					d.pgroup[p] = g | ((uint32_t)B2CU_PROXY_MOVED << 16);
					d.movedList[atomicAdd(&d.counters[CNT_SCRATCH], 1)] = p;
				}
			}
		}
		if (t < contactCount)
		{
			const int i = t;
			uint32_t f = d.c.flags[i];
			if (!(f & B2CU_CONTACT_DEAD) && (f & (B2CU_CONTACT_ISLAND | B2CU_CONTACT_TOI)))
			{
				int4 pr = d.c.proxies[i];
				uint32_t fA = d.bflags[pr.z], fB = d.bflags[pr.w];
				if ((IsDynamic(fA) && (fA & B2CU_BODY_ISLAND)) || (IsDynamic(fB) && (fB & B2CU_BODY_ISLAND)))
					d.c.flags[i] = f & ~(uint32_t)(B2CU_CONTACT_ISLAND | B2CU_CONTACT_TOI);
			}
		}
	}
}

// FindNewContacts for the proxies the event moved (b2World.cpp:1015-1022, single-threaded b2BroadPhase::UpdatePairs):
// each proxy of the world tests its fat box against the boxes of the moved ones (staged in shared memory).  A pair of
// two moved proxies is reported by the one with the larger id.  Also takes the island flag off the event's bodies.
#define B2CU_TOI_MOVED_TILE 128
__global__ void __launch_bounds__(256) ToiFindPairsKernel(DeviceArrays d, int proxyCount, int2 contactCounts, int pairCapacity)
{
	GridDependencyWait();
	__shared__ float4 shBox[B2CU_TOI_MOVED_TILE];
	__shared__ int shId[B2CU_TOI_MOVED_TILE];
	if (blockIdx.x == 0)
	{
		const int nb = d.toiScratch[B2CU_TOI_SCR_BODY_COUNT];
		for (int k = threadIdx.x; k < nb; k += blockDim.x)
		{
			const int b = d.toiScratch[B2CU_TOI_SCR_BODIES + k];
			d.bflags[b] &= ~(uint32_t)B2CU_BODY_ISLAND;
		}
	}
	const int moved = d.counters[CNT_SCRATCH];
	if (moved == 0) return;
	for (int base = 0; base < moved; base += B2CU_TOI_MOVED_TILE)
	{
		const int n = min(B2CU_TOI_MOVED_TILE, moved - base);
		__syncthreads();
		if (threadIdx.x < n)
		{
			int q = d.movedList[base + threadIdx.x];
			shId[threadIdx.x] = q;
			shBox[threadIdx.x] = d.fat[q];
		}
		__syncthreads();
		B2CU_GRID_STRIDE(r, proxyCount)
		{
			const uint32_t flagsR = d.pgroup[r] >> 16;
			if (flagsR & B2CU_PROXY_INACTIVE) continue;
			const float4 fr = d.fat[r];
			const bool movedR = (flagsR & B2CU_PROXY_MOVED) != 0;
			for (int k = 0; k < n; ++k)
			{
				const int q = shId[k];
				if (q == r || (movedR && r < q)) continue;
				if (!AabbOverlap(shBox[k], fr)) continue;
				TryAddPair(d, q, r, contactCounts, pairCapacity);
			}
		}
	}
}

} // namespace b2cu
