// b2cu_kernels.cuh -- the phase kernels of b2cuStep.  One work item per thread, grid-stride, SoA float4 traffic.
// Included only by world.cu.  Each kernel names the reference code it replaces.
#pragma once

#include <cooperative_groups.h>

#include "b2cu_collide.cuh"
#include "b2cu_gjk.cuh"
#include "b2cu_toi.cuh"
#include "b2cu_world.cuh"
#include "b2cu_joints.cuh"

namespace b2cu
{

#define B2CU_GRID_STRIDE(i, n) \
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += gridDim.x * blockDim.x)

#define B2CU_EV_BEGIN 1
#define B2CU_EV_END 2
#define B2CU_EV_DESTROY 4
#define B2CU_EV_DESTROY_TOUCHING 8
// internal contact flag: destroyed, slot not yet reclaimed (the contact set is compacted lazily)
#define B2CU_CONTACT_DEAD 0x0100u
// internal contact flag: one of the two fixtures is a sensor (b2Fixture::IsSensor): overlap test instead of a manifold,
// never solid
#define B2CU_CONTACT_SENSOR 0x0200u

__device__ __forceinline__ bool IsStatic(uint32_t bf) { return (bf & B2CU_BODY_TYPE_MASK) == B2CU_STATIC_BODY; }
__device__ __forceinline__ bool IsDynamic(uint32_t bf) { return (bf & B2CU_BODY_TYPE_MASK) == B2CU_DYNAMIC_BODY; }
// dynamic and simulated by this shard (not a halo copy of a neighbour's body)
__device__ __forceinline__ bool IsOwnedDynamic(uint32_t bf) { return IsDynamic(bf) && !(bf & B2CU_BODY_GHOST); }
__device__ __forceinline__ bool IsAwakeNonStatic(uint32_t bf) { return (bf & B2CU_BODY_AWAKE) && !IsStatic(bf); }

// ordered-int encoding of a float so that signed integer min == float min
__device__ __forceinline__ int FloatToOrdered(float f)
{
	int i = __float_as_int(f);
	return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float OrderedToFloat(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// atomicMin(&base[root], value) with the lanes of a warp that share `root` combined first: a pile is one island,
// so without this every constraint of a colour would hit the same address
__device__ __forceinline__ void AtomicMinByRoot(int* base, int root, int value)
{
	unsigned active = __activemask();
	unsigned peers = __match_any_sync(active, root);
	int m = __reduce_min_sync(peers, value);
	if ((int)(threadIdx.x & 31) == __ffs((int)peers) - 1) atomicMin(&base[root], m);
}

// b2TestOverlap(b2AABB, b2AABB), Box2D/Collision/b2Collision.h:273-286
__device__ __forceinline__ bool AabbOverlap(float4 a, float4 b)
{
	float d1x = b.x - a.z, d1y = b.y - a.w;
	float d2x = a.x - b.z, d2y = a.y - b.w;
	if (d1x > 0.0f || d1y > 0.0f) return false;
	if (d2x > 0.0f || d2y > 0.0f) return false;
	return true;
}

// b2ContactFilter::ShouldCollide, Box2D/Dynamics/b2WorldCallbacks.cpp:24-38
__device__ __forceinline__ bool DefaultFilter(uint32_t filterA, uint32_t groupA, uint32_t filterB, uint32_t groupB)
{
	int16_t gA = (int16_t)(groupA & 0xFFFFu), gB = (int16_t)(groupB & 0xFFFFu);
	if (gA == gB && gA != 0)
	{
		return gA > 0;
	}
	uint32_t catA = filterA & 0xFFFFu, maskA = filterA >> 16;
	uint32_t catB = filterB & 0xFFFFu, maskB = filterB >> 16;
	return (maskA & catB) != 0 && (catA & maskB) != 0;
}

__device__ __forceinline__ int LowerBound64(const uint64_t* __restrict__ keys, int n, uint64_t key)
{
	int lo = 0, hi = n;
	while (lo < hi)
	{
		int mid = (lo + hi) >> 1;
		if (keys[mid] < key) lo = mid + 1;
		else hi = mid;
	}
	return lo;
}

// b2Body::ShouldCollide, joint half (b2Body.cpp:437-446): bodies connected by a joint that does not allow it never collide
__device__ __forceinline__ bool JointBlocks(const DeviceArrays& d, int bodyA, int bodyB)
{
	if (d.jointPairCount == 0) return false;
	uint32_t lo = (uint32_t)(bodyA < bodyB ? bodyA : bodyB), hi = (uint32_t)(bodyA < bodyB ? bodyB : bodyA);
	uint64_t key = ((uint64_t)lo << 32) | hi;
	int k = LowerBound64(d.jointPairKeys, d.jointPairCount, key);
	return k < d.jointPairCount && d.jointPairKeys[k] == key;
}

// ---------------------------------------------------------------------------------------------------------
// Collide: b2ContactManager::Collide (Box2D/Dynamics/b2ContactManager.cpp:177-230) fused with
// b2Contact::UpdateImpl (Box2D/Dynamics/Contacts/b2Contact.cpp:173-298).  One thread per contact.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void AppendKey(const DeviceArrays& d, int counter, uint64_t* list, uint64_t key, int capacity)
{
	int slot = atomicAdd(&d.counters[counter], 1);
	if (slot < capacity) list[slot] = key;
}

// b2Contact::Update (b2Contact.cpp:163-246) of live contact i: evaluate the manifold, carry the warm-start impulses
// over by feature id, raise begin/end events and wake requests.  Returns 1 when the contact is touching afterwards.
__device__ __forceinline__ int UpdateContact(const DeviceArrays& d, int i, int4 pr, int bA, int bB, uint32_t flags,
                                             uint4 m3, const b2cuShape* sA, const b2cuShape* sB, int capacity)
{
	int touchingNow = 0;
	// ---- b2Contact::Update ----
	float4 o1 = d.c.m1[i], o2 = d.c.m2[i];
	int oldCount = (int)m3.w;
	flags |= B2CU_CONTACT_ENABLED;
	bool wasTouching = (flags & B2CU_CONTACT_TOUCHING) != 0;

	Manifold m;
	float4 o0 = d.c.m0[i];
	m.localNormal = V(o0.x, o0.y);
	m.localPoint = V(o0.z, o0.w);
	m.lp[0] = V(o1.x, o1.y);
	m.lp[1] = V(o2.x, o2.y);
	m.id[0] = m3.x;
	m.id[1] = m3.y;
	m.type = (int)m3.z;
	m.pointCount = 0;

	Xf xfA = MakeXf(d.xf[bA]);
	Xf xfB = MakeXf(d.xf[bB]);
	const bool sensor = (flags & B2CU_CONTACT_SENSOR) != 0;
	bool touching;
	if (sensor)
	{
		// sensors do not generate manifolds (b2Contact.cpp:193-202): overlap by distance, point count zeroed, the rest of
		// the manifold left as it is
		touching = TestOverlap(sA, xfA, sB, xfB);
		m.pointCount = 0;
	}
	else
	{
		Evaluate(&m, sA, xfA, sB, xfB);
		touching = m.pointCount > 0;
	}

	// warm-start transfer by feature id
	for (int k = 0; k < m.pointCount; ++k)
	{
		float ni = 0.0f, ti = 0.0f;
		uint32_t id2 = m.id[k];
		if (oldCount > 0 && m3.x == id2)
		{
			ni = o1.z;
			ti = o1.w;
		}
		else if (oldCount > 1 && m3.y == id2)
		{
			ni = o2.z;
			ti = o2.w;
		}
		m.ni[k] = ni;
		m.ti[k] = ti;
	}
	for (int k = m.pointCount; k < 2; ++k)
	{
		// points beyond pointCount keep their previous impulses, as the untouched b2ManifoldPoint would
		m.ni[k] = k == 0 ? o1.z : o2.z;
		m.ti[k] = k == 0 ? o1.w : o2.w;
	}

	int ev = 0;
	if (touching != wasTouching)
	{
		// b2ContactManager::ConsumeAwakes wakes m_nodeB.other (= fixture A's body) only (:472-486); a sensor contact
		// wakes nobody (the awake request sits in the non-sensor branch of Update, b2Contact.cpp:229-240)
		if (!sensor) d.wake[bA] = 1;
		ev = touching ? B2CU_EV_BEGIN : B2CU_EV_END;
	}
	if (touching)
	{
		flags |= B2CU_CONTACT_TOUCHING;
		touchingNow = 1;
	}
	else
	{
		flags &= ~B2CU_CONTACT_TOUCHING;
	}
	// deferred Begin/End buffers (b2ContactManagerPerThreadData::m_beginContacts / m_endContacts); sorted by key
	// afterwards, which is what b2ThreadDataSorter does for the reference (b2ContactManager.cpp:388-433)
	if (ev == B2CU_EV_BEGIN) AppendKey(d, CNT_BEGIN, d.beginKeys, d.c.key[i], capacity);
	else if (ev == B2CU_EV_END) AppendKey(d, CNT_END, d.endKeys, d.c.key[i], capacity);

	d.c.flags[i] = flags;
	if (sensor)
	{
		d.c.m3[i] = make_uint4(m3.x, m3.y, m3.z, 0u);
		return touchingNow;
	}
	d.c.m0[i] = make_float4(m.localNormal.x, m.localNormal.y, m.localPoint.x, m.localPoint.y);
	d.c.m1[i] = make_float4(m.lp[0].x, m.lp[0].y, m.ni[0], m.ti[0]);
	d.c.m2[i] = make_float4(m.lp[1].x, m.lp[1].y, m.ni[1], m.ti[1]);
	d.c.m3[i] = make_uint4(m.id[0], m.id[1], (uint32_t)m.type, (uint32_t)m.pointCount);
	return touchingNow;
}

__global__ void __launch_bounds__(256) CollideKernel(DeviceArrays d, int contactCount, int mainCount, int capacity, int* __restrict__ heavyList)
{
	GridDependencyWait();
	int touchingCount = 0;
	B2CU_GRID_STRIDE(i, contactCount)
	{
		if (d.c.flags[i] & B2CU_CONTACT_DEAD) continue;
		int4 pr = d.c.proxies[i];
		int bA = pr.z, bB = pr.w;
		uint32_t fbA = d.bflags[bA], fbB = d.bflags[bB];
		uint32_t flags = d.c.flags[i];
		uint4 m3 = d.c.m3[i];
		bool destroy = false;

		if (flags & B2CU_CONTACT_FILTER)
		{
			bool should = (IsOwnedDynamic(fbA) || IsOwnedDynamic(fbB)) && !JointBlocks(d, bA, bB);
			if (should)
			{
				// under a caller's pair filter an existing contact is kept: that filter is only consulted for new pairs
				should = d.customFilter || DefaultFilter(d.pfilter[pr.x], d.pgroup[pr.x], d.pfilter[pr.y], d.pgroup[pr.y]);
			}
			if (!should)
			{
				destroy = true;
			}
			else
			{
				flags &= ~B2CU_CONTACT_FILTER;
			}
		}

		if (!destroy)
		{
			bool active = IsAwakeNonStatic(fbA) || IsAwakeNonStatic(fbB);
			if (!active)
			{
				d.c.flags[i] = flags | B2CU_CONTACT_INACTIVE;
				if (flags & B2CU_CONTACT_TOUCHING) ++touchingCount;
				continue;
			}
			flags &= ~B2CU_CONTACT_INACTIVE;
			if (!AabbOverlap(d.fat[pr.x], d.fat[pr.y]))
			{
				destroy = true;
			}
		}

		if (destroy)
		{
			int ev = B2CU_EV_DESTROY;
			if (flags & B2CU_CONTACT_TOUCHING) ev |= B2CU_EV_DESTROY_TOUCHING;
			// b2Contact::Destroy wakes both bodies if the manifold had points (b2Contact.cpp:107-113)
			if ((int)m3.w > 0)
			{
				d.wake[bA] = 1;
				d.wake[bB] = 1;
			}
			d.c.flags[i] = flags | B2CU_CONTACT_DEAD;
			atomicAdd(&d.counters[i < mainCount ? CNT_DESTROY : CNT_DESTROY_B], 1);
			if (ev & B2CU_EV_DESTROY_TOUCHING) AppendKey(d, CNT_DESTROY_END, d.destroyEndKeys, d.c.key[i], capacity);
			continue;
		}

		const b2cuShape* sA = d.shapes + d.pshape[pr.x];
		const b2cuShape* sB = d.shapes + d.pshape[pr.y];
		// polygon-polygon and edge-polygon manifolds cost several times the others; running them in the same
		// warps leaves most lanes idle, so they are queued for a second, dense pass (CollideHeavyKernel)
		bool heavy = (sB->type == B2CU_SHAPE_POLYGON && sA->type != B2CU_SHAPE_CIRCLE) || (flags & B2CU_CONTACT_SENSOR);
		if (heavy)
		{
			d.c.flags[i] = flags;
			unsigned peers = __activemask();
			int lane = threadIdx.x & 31;
			int leader = __ffs(peers) - 1;
			int base = 0;
			if (lane == leader) base = atomicAdd(&d.counters[CNT_HEAVY], __popc(peers));
			base = __shfl_sync(peers, base, leader);
			heavyList[base + __popc(peers & ((1u << lane) - 1u))] = i;
			continue;
		}
		touchingCount += UpdateContact(d, i, pr, bA, bB, flags, m3, sA, sB, capacity);
	}
	for (int dlt = 16; dlt > 0; dlt >>= 1) touchingCount += __shfl_down_sync(0xffffffffu, touchingCount, dlt);
	if ((threadIdx.x & 31) == 0 && touchingCount) atomicAdd(&d.counters[CNT_TOUCHING], touchingCount);
}

// second pass of Collide over the queued polygon-polygon / edge-polygon contacts
__global__ void __launch_bounds__(256) CollideHeavyKernel(DeviceArrays d, const int* __restrict__ heavyList, int capacity)
{
	GridDependencyWait();
	int touchingCount = 0;
	int n = d.counters[CNT_HEAVY];
	B2CU_GRID_STRIDE(k, n)
	{
		int i = heavyList[k];
		int4 pr = d.c.proxies[i];
		int bA = pr.z, bB = pr.w;
		touchingCount += UpdateContact(d, i, pr, bA, bB, d.c.flags[i], d.c.m3[i], d.shapes + d.pshape[pr.x],
		                               d.shapes + d.pshape[pr.y], capacity);
	}
	for (int dlt = 16; dlt > 0; dlt >>= 1) touchingCount += __shfl_down_sync(0xffffffffu, touchingCount, dlt);
	if ((threadIdx.x & 31) == 0 && touchingCount) atomicAdd(&d.counters[CNT_TOUCHING], touchingCount);
}

// ---------------------------------------------------------------------------------------------------------
// Host transport of the body records: the caller's mirror is an array of b2cuBody (26 words, include/b2cuda.h),
// the device keeps bodies as float4 columns.  The conversion runs here, through shared memory so that both the
// column side and the record side are coalesced, and the PCIe copy is one contiguous transfer.
// ---------------------------------------------------------------------------------------------------------
#define B2CU_BODY_WORDS 26
static_assert(sizeof(b2cuBody) == B2CU_BODY_WORDS * 4, "b2cuBody layout");

__global__ void __launch_bounds__(256) PackBodiesKernel(DeviceArrays d, int first, int count, float* __restrict__ out)
{
	GridDependencyWait();
	__shared__ float sh[256 * B2CU_BODY_WORDS];
	for (int tile = blockIdx.x; tile * 256 < count; tile += gridDim.x)
	{
		int r = tile * 256 + threadIdx.x;
		if (r < count)
		{
			int b = first + r;
			float4 xf = d.xf[b], pos = d.pos[b], pos0 = d.pos0[b], vel = d.vel[b], mass = d.mass[b];
			float4 force = d.force[b], damp = d.damp[b];
			float* s = sh + threadIdx.x * B2CU_BODY_WORDS;
			s[0] = xf.x; s[1] = xf.y; s[2] = xf.z; s[3] = xf.w;
			s[4] = pos.x; s[5] = pos.y; s[6] = pos.z;
			s[7] = pos0.x; s[8] = pos0.y; s[9] = pos0.z; s[10] = pos0.w;
			s[11] = mass.z; s[12] = mass.w;
			s[13] = vel.x; s[14] = vel.y; s[15] = vel.z;
			s[16] = force.x; s[17] = force.y; s[18] = force.z;
			s[19] = mass.x; s[20] = mass.y;
			s[21] = damp.x; s[22] = damp.y; s[23] = damp.z;
			s[24] = force.w;
			s[25] = __uint_as_float(d.bflags[b]);
		}
		__syncthreads();
		int n = min(256, count - tile * 256) * B2CU_BODY_WORDS;
		float* o = out + (size_t)tile * 256 * B2CU_BODY_WORDS;
		for (int i = threadIdx.x; i < n; i += 256) o[i] = sh[i];
		__syncthreads();
	}
}

#define B2CU_STATE_WORDS 12
static_assert(sizeof(b2cuBodyState) == B2CU_STATE_WORDS * 4, "b2cuBodyState layout");

// device -> host direction: only the fields a step changes and callers read (b2cuBodyState, 48 bytes)
__global__ void __launch_bounds__(256) PackBodyStatesKernel(DeviceArrays d, int first, int count, float* __restrict__ out)
{
	GridDependencyWait();
	__shared__ float sh[256 * (B2CU_STATE_WORDS + 1)]; // +1: rows of a multiple of four words would collide in the banks
	for (int tile = blockIdx.x; tile * 256 < count; tile += gridDim.x)
	{
		int r = tile * 256 + threadIdx.x;
		if (r < count)
		{
			int b = first + r;
			float4 xf = d.xf[b], pos = d.pos[b], vel = d.vel[b];
			float* s = sh + threadIdx.x * (B2CU_STATE_WORDS + 1);
			s[0] = xf.x; s[1] = xf.y; s[2] = xf.z; s[3] = xf.w;
			s[4] = pos.x; s[5] = pos.y; s[6] = pos.z;
			s[7] = vel.x; s[8] = vel.y; s[9] = vel.z;
			s[10] = d.force[b].w;
			s[11] = __uint_as_float(d.bflags[b]);
		}
		__syncthreads();
		int n = min(256, count - tile * 256) * B2CU_STATE_WORDS;
		float* o = out + (size_t)tile * 256 * B2CU_STATE_WORDS;
		for (int i = threadIdx.x; i < n; i += 256) o[i] = sh[(i / B2CU_STATE_WORDS) * (B2CU_STATE_WORDS + 1) + (i % B2CU_STATE_WORDS)];
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) UnpackBodiesKernel(DeviceArrays d, int first, int count, const float* __restrict__ in)
{
	GridDependencyWait();
	__shared__ float sh[256 * B2CU_BODY_WORDS];
	for (int tile = blockIdx.x; tile * 256 < count; tile += gridDim.x)
	{
		int n = min(256, count - tile * 256) * B2CU_BODY_WORDS;
		const float* src = in + (size_t)tile * 256 * B2CU_BODY_WORDS;
		for (int i = threadIdx.x; i < n; i += 256) sh[i] = src[i];
		__syncthreads();
		int r = tile * 256 + threadIdx.x;
		if (r < count)
		{
			int b = first + r;
			const float* s = sh + threadIdx.x * B2CU_BODY_WORDS;
			d.xf[b] = make_float4(s[0], s[1], s[2], s[3]);
			d.pos[b] = make_float4(s[4], s[5], s[6], 0.0f);
			d.pos0[b] = make_float4(s[7], s[8], s[9], s[10]);
			d.mass[b] = make_float4(s[19], s[20], s[11], s[12]);
			d.vel[b] = make_float4(s[13], s[14], s[15], 0.0f);
			d.force[b] = make_float4(s[16], s[17], s[18], s[24]);
			d.damp[b] = make_float4(s[21], s[22], s[23], 0.0f);
			const uint32_t flags = __float_as_uint(s[25]);
			// bit 0: a body type changed (the joint colouring depends on it); bit 1: type or bullet flag changed (the
			// time-of-impact candidacy of the body's contacts depends on them, b2Contact::IsToiCandidate)
			const uint32_t diff = flags ^ d.bflags[b];
			if (diff & (B2CU_BODY_TYPE_MASK | B2CU_BODY_BULLET))
				atomicOr(&d.counters[CNT_BODY_TYPE_CHANGED], ((diff & B2CU_BODY_TYPE_MASK) ? 1 : 0) | 2);
			d.bflags[b] = flags;
		}
		__syncthreads();
	}
}

// b2cuSetBodyForces: (fx, fy, torque) rows into the force column; the sleep timer in its fourth lane stays
__global__ void SetBodyForcesKernel(DeviceArrays d, int first, int count, const float* __restrict__ in)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(r, count)
	{
		float4 f = d.force[first + r];
		f.x = in[3 * r];
		f.y = in[3 * r + 1];
		f.z = in[3 * r + 2];
		d.force[first + r] = f;
	}
}

// pradius[p] = m_radius of the proxy's shape (b2Shape.h:93), read by the constraint initialisation
__global__ void FillProxyRadiusKernel(DeviceArrays d, int proxyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(p, proxyCount) { d.pradius[p] = d.shapes[d.pshape[p]].radius; }
}

// ---- PreSolve hook (b2cuSetPreSolveHook) ----
// contacts the narrow phase has just updated and found touching: what b2Contact::Update reports to PreSolve
__global__ void PreSolveSelectKernel(DeviceArrays d, int contactCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, contactCount)
	{
		uint32_t f = d.c.flags[i];
		d.cSelect[i] = ((f & B2CU_CONTACT_TOUCHING) && !(f & (B2CU_CONTACT_DEAD | B2CU_CONTACT_INACTIVE | B2CU_CONTACT_SENSOR))) ? 1 : 0;
	}
}
// records of the selected contacts + their manifolds of the previous step (saved in cAlt before Collide)
__global__ void PreSolveGatherKernel(DeviceArrays d, const int* __restrict__ list, int n, b2cuContact* __restrict__ out,
                                     b2cuManifold* __restrict__ oldOut)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(j, n)
	{
		int i = list[j];
		int4 pr = d.c.proxies[i];
		float4 m0 = d.c.m0[i], m1 = d.c.m1[i], m2 = d.c.m2[i], mix = d.c.mix[i];
		uint4 m3 = d.c.m3[i];
		b2cuContact o;
		o.proxyA = pr.x;
		o.proxyB = pr.y;
		o.flags = d.c.flags[i] & 0xFFu;
		o.friction = mix.x;
		o.restitution = mix.y;
		o.tangentSpeed = mix.z;
		o.toiCount = d.c.toiCount[i];
		o.toi = mix.w;
		o.manifold.localNormal[0] = m0.x; o.manifold.localNormal[1] = m0.y;
		o.manifold.localPoint[0] = m0.z; o.manifold.localPoint[1] = m0.w;
		o.manifold.points[0].localPoint[0] = m1.x; o.manifold.points[0].localPoint[1] = m1.y;
		o.manifold.points[0].normalImpulse = m1.z; o.manifold.points[0].tangentImpulse = m1.w;
		o.manifold.points[1].localPoint[0] = m2.x; o.manifold.points[1].localPoint[1] = m2.y;
		o.manifold.points[1].normalImpulse = m2.z; o.manifold.points[1].tangentImpulse = m2.w;
		o.manifold.id[0] = m3.x; o.manifold.id[1] = m3.y;
		o.manifold.type = (int32_t)m3.z;
		o.manifold.pointCount = (int32_t)m3.w;
		o.stamp = d.c.stamp[i];
		o.reserved = 0u;
		out[j] = o;
		float4 p0 = d.cAlt.m0[i], p1 = d.cAlt.m1[i], p2 = d.cAlt.m2[i];
		uint4 p3 = d.cAlt.m3[i];
		b2cuManifold q;
		q.localNormal[0] = p0.x; q.localNormal[1] = p0.y;
		q.localPoint[0] = p0.z; q.localPoint[1] = p0.w;
		q.points[0].localPoint[0] = p1.x; q.points[0].localPoint[1] = p1.y;
		q.points[0].normalImpulse = p1.z; q.points[0].tangentImpulse = p1.w;
		q.points[1].localPoint[0] = p2.x; q.points[1].localPoint[1] = p2.y;
		q.points[1].normalImpulse = p2.z; q.points[1].tangentImpulse = p2.w;
		q.id[0] = p3.x; q.id[1] = p3.y;
		q.type = (int32_t)p3.z;
		q.pointCount = (int32_t)p3.w;
		oldOut[j] = q;
	}
}
// b2Contact::SetEnabled(false) for a list of keys
__global__ void DisableContactsKernel(DeviceArrays d, int contactCount, int mainCount, const uint64_t* __restrict__ keys, int n)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(j, n)
	{
		uint64_t key = keys[j];
		int i = LowerBound64(d.c.key, mainCount, key);
		bool found = i < mainCount && d.c.key[i] == key && !(d.c.flags[i] & B2CU_CONTACT_DEAD);
		if (!found && contactCount > mainCount)
		{
			i = mainCount + LowerBound64(d.c.key + mainCount, contactCount - mainCount, key);
			found = i < contactCount && d.c.key[i] == key && !(d.c.flags[i] & B2CU_CONTACT_DEAD);
		}
		if (found) d.c.flags[i] &= ~(uint32_t)B2CU_CONTACT_ENABLED;
	}
}

// b2Fixture::Refilter (b2Fixture.cpp:197-210): the contacts of a refiltered proxy get e_filterFlag
__global__ void FlagFilterContactsKernel(DeviceArrays d, int contactCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, contactCount)
	{
		uint32_t f = d.c.flags[i];
		if (f & B2CU_CONTACT_DEAD) continue;
		int4 pr = d.c.proxies[i];
		if (((d.pgroup[pr.x] | d.pgroup[pr.y]) >> 16) & B2CU_PROXY_REFILTER) d.c.flags[i] = f | B2CU_CONTACT_FILTER;
	}
}
// The start-box shortcut of the pair search assumes that a pair whose boxes overlapped before has been judged before and
// would be judged the same again.  A pair of bodies that a joint kept apart until it was destroyed is the exception:
// the reference finds that pair at the next move of either proxy (b2BroadPhase::UpdatePairs -> AddPair), so it must
// go through the full test.
__device__ __forceinline__ bool JointFreed(const DeviceArrays& d, int bodyA, int bodyB)
{
	uint32_t lo = (uint32_t)(bodyA < bodyB ? bodyA : bodyB), hi = (uint32_t)(bodyA < bodyB ? bodyB : bodyA);
	uint64_t key = ((uint64_t)lo << 32) | hi;
	int k = LowerBound64(d.jointFreedKeys, d.jointFreedCount, key);
	return k < d.jointFreedCount && d.jointFreedKeys[k] == key;
}

// contacts between two bodies that a joint keeps from colliding get e_filterFlag; Collide then removes them
__global__ void FlagJointContactsKernel(DeviceArrays d, int contactCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, contactCount)
	{
		uint32_t f = d.c.flags[i];
		if (f & B2CU_CONTACT_DEAD) continue;
		int4 pr = d.c.proxies[i];
		if (JointBlocks(d, pr.z, pr.w)) d.c.flags[i] = f | B2CU_CONTACT_FILTER;
	}
}
__global__ void ClearProxyFlagKernel(DeviceArrays d, int proxyCount, uint32_t flag)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(p, proxyCount) { d.pgroup[p] &= ~(flag << 16); }
}

// contacts carry the bodies of their two proxies next to the proxy ids (one gather level less in every contact
// kernel); this fills them in after the caller has uploaded contacts or proxies
__global__ void FillContactBodiesKernel(DeviceArrays d, int contactCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, contactCount)
	{
		int4 pr = d.c.proxies[i];
		d.c.proxies[i] = make_int4(pr.x, pr.y, d.pbody[pr.x], d.pbody[pr.y]);
		// sensor-ness and TOI candidacy follow the fixtures (b2Fixture::SetSensor / SetThickShape recalculate them,
		// b2Fixture.cpp:222-262): same rule as at creation (RebuildNewKernel)
		uint32_t f = d.c.flags[i] & ~(B2CU_CONTACT_SENSOR | (uint32_t)B2CU_CONTACT_TOI_CANDIDATE);
		uint32_t pf = (d.pgroup[pr.x] | d.pgroup[pr.y]) >> 16;
		if (pf & B2CU_PROXY_SENSOR)
		{
			f |= B2CU_CONTACT_SENSOR;
		}
		else
		{
			uint32_t fbA = d.bflags[d.pbody[pr.x]], fbB = d.bflags[d.pbody[pr.y]];
			bool includesNonDynamic = !IsDynamic(fbA) || !IsDynamic(fbB);
			if (((fbA | fbB) & B2CU_BODY_BULLET) || (includesNonDynamic && !(pf & B2CU_PROXY_THICK)))
				f |= B2CU_CONTACT_TOI_CANDIDATE;
		}
		d.c.flags[i] = f;
	}
}

// b2Body::SetAwake(true) for every body flagged by Collide / contact creation / contact destruction
// (Box2D/Dynamics/b2Body.h:690-718: sets e_awakeFlag and resets m_sleepTime).
__global__ void ApplyWakeKernel(DeviceArrays d, int bodyCount, int* __restrict__ patch)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		if (d.wake[b])
		{
			d.wake[b] = 0;
			uint32_t bf = d.bflags[b];
			float4 f = d.force[b];
			// patch != nullptr: the body records are already on their way to the caller's mirror (b2cuSetBodyMirror);
			// the bodies this changes are listed so that the mirror can be corrected afterwards
			if (patch && (!(bf & B2CU_BODY_AWAKE) || f.w != 0.0f)) patch[atomicAdd(&d.counters[CNT_WAKE_PATCH], 1)] = b;
			d.bflags[b] = bf | B2CU_BODY_AWAKE;
			f.w = 0.0f;
			d.force[b] = f;
		}
	}
}

// out[offset + j] = key[list[j]] for j < *count
__global__ void GatherKeysKernel(const uint64_t* __restrict__ key, const int* __restrict__ list,
                                 const int* __restrict__ count, const int* __restrict__ offset,
                                 uint64_t* __restrict__ out, int capacity)
{
	GridDependencyWait();
	int n = *count;
	int off = offset ? *offset : 0;
	B2CU_GRID_STRIDE(j, n)
	{
		if (off + j < capacity) out[off + j] = key[list[j]];
	}
}

__global__ void CountMaskKernel(const uint32_t* __restrict__ flags, uint32_t mask, int n, int* counter)
{
	GridDependencyWait();
	int local = 0;
	B2CU_GRID_STRIDE(i, n)
	{
		if (flags[i] & mask) ++local;
	}
	for (int dlt = 16; dlt > 0; dlt >>= 1) local += __shfl_down_sync(0xffffffffu, local, dlt);
	if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, local);
}

// ---------------------------------------------------------------------------------------------------------
// Islands: parallel union-find replaces the serial DFS of b2World::Solve (Box2D/Dynamics/b2World.cpp:1200-1371).
// Non-static bodies joined by a touching, enabled, solid contact share an island; static bodies never merge
// islands (:1236-1241).  The island label is the smallest body id of the island, whatever the thread order.
// ---------------------------------------------------------------------------------------------------------
__global__ void SolveInitBodiesKernel(DeviceArrays d, int bodyCount, int positionIterations)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		d.island[b] = b;
		d.islandAwake[b] = 0;
		d.islandMinSleep[b] = 0x7F7FFFFF;
		d.colourMask[b] = 0u;
		d.colourClaim[b] = 0ull;
		for (int k = 0; k < positionIterations; ++k)
		{
			d.islandMinSep[k * bodyCount + b] = 0x7F7FFFFF;
		}
	}
}

__device__ __forceinline__ int UfFind(int* parent, int x)
{
	int p = parent[x];
	while (p != x)
	{
		int gp = parent[p];
		if (gp != p) parent[x] = gp; // path halving; benign race (only ever points further up)
		x = p;
		p = gp;
	}
	return x;
}

// root of x without touching the forest: for the pass that writes the final labels.  Path halving there would race
// with those writes: a halving store of an ancestor, issued by a thread that read the pointers earlier, could land
// AFTER the owner of that node has stored its root, leaving a non-root label behind (two labels for one island).
__device__ __forceinline__ int UfFindReadOnly(const int* parent, int x)
{
	int p = parent[x];
	while (p != x)
	{
		x = p;
		p = parent[x];
	}
	return x;
}

__device__ __forceinline__ void UfUnion(int* parent, int a, int b)
{
	while (true)
	{
		a = UfFind(parent, a);
		b = UfFind(parent, b);
		if (a == b) return;
		if (a < b)
		{
			int t = a;
			a = b;
			b = t;
		}
		// hook the larger root under the smaller one
		int old = atomicCAS(&parent[a], a, b);
		if (old == a) return;
	}
}

__device__ __forceinline__ bool IsSolidTouching(const DeviceArrays& d, int i)
{
	uint32_t f = d.c.flags[i];
	if ((f & (B2CU_CONTACT_TOUCHING | B2CU_CONTACT_ENABLED)) != (B2CU_CONTACT_TOUCHING | B2CU_CONTACT_ENABLED)) return false;
	if (f & (B2CU_CONTACT_DEAD | B2CU_CONTACT_SENSOR)) return false;
	return true;
}

// Two sweeps (the idea of sampling-based connected components): the first unites over every `sample`-th contact only,
// which already connects almost everything a pile connects; IslandCompressKernel then points every body at its root,
// and the second sweep finds for most of the remaining contacts, after one hop each, that both bodies share a root.
// phase 0: contacts with i % sample == 0; phase 1: the others; sample == 1 with phase 0: everything in one sweep.
__global__ void IslandUnionKernel(DeviceArrays d, int contactCount, int sample, int phase)
{
	GridDependencyWait();
	const int chunk = (contactCount + 31) / 32;
	B2CU_GRID_STRIDE(g, (sample < 0 ? chunk * 32 : contactCount))
	{
		// sample < 0 (experiment): the lanes of a warp take contacts from 32 distant parts of the (key-sorted) set, so
		// that they do not all unite into the same body
		const int i = sample < 0 ? (g & 31) * chunk + (g >> 5) : g;
		if (i >= contactCount) continue;
		if (sample > 1 && ((i % sample == 0) != (phase == 0))) continue;
		if (!IsSolidTouching(d, i)) continue;
		int4 pr = d.c.proxies[i];
		int bA = pr.z, bB = pr.w;
		if (IsStatic(d.bflags[bA]) || IsStatic(d.bflags[bB])) continue;
		UfUnion(d.island, bA, bB);
	}
}

// joints connect islands like touching contacts do (b2World.cpp:1286-1320): both bodies active, neither static
__global__ void JointUnionKernel(DeviceArrays d, int jointCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(j, jointCount)
	{
		int bA = d.joints[j].bodyA, bB = d.joints[j].bodyB;
		uint32_t fA = d.bflags[bA], fB = d.bflags[bB];
		if (IsStatic(fA) || IsStatic(fB)) continue;
		if (!(fA & B2CU_BODY_ACTIVE) || !(fB & B2CU_BODY_ACTIVE)) continue;
		UfUnion(d.island, bA, bB);
	}
}

// between the two sweeps: every body straight under its current root (no unions run meanwhile, so roots are fixed;
// a concurrent reader sees either the old ancestor or the root, both on its path)
__global__ void IslandCompressKernel(DeviceArrays d, int bodyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount) { d.island[b] = UfFindReadOnly(d.island, b); }
}

__global__ void IslandFlattenKernel(DeviceArrays d, int bodyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		int r = UfFind(d.island, b);
		uint32_t bf = d.bflags[b];
		// seeds: awake, active, non-static (b2World.cpp:1207-1219)
		if (IsAwakeNonStatic(bf) && (bf & B2CU_BODY_ACTIVE)) d.islandAwake[r] = 1;
	}
}

__global__ void IslandMarkKernel(DeviceArrays d, int bodyCount)
{
	GridDependencyWait();
	int local = 0;
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		// read-only find: the only stores of this kernel are final labels (a root), so a concurrent traversal through
		// a node that has already been labelled simply jumps to the root
		int r = UfFindReadOnly(d.island, b);
		d.island[b] = r;
		uint32_t bf = d.bflags[b];
		if (!IsStatic(bf) && d.islandAwake[r])
		{
			// the traversal wakes every body it reaches without resetting its sleep timer (:1243-1244)
			d.bflags[b] = bf | B2CU_BODY_ISLAND | B2CU_BODY_AWAKE;
			++local;
		}
	}
	for (int dlt = 16; dlt > 0; dlt >>= 1) local += __shfl_down_sync(0xffffffffu, local, dlt);
	if ((threadIdx.x & 31) == 0 && local) atomicAdd(&d.counters[CNT_ISLAND_BODIES], local);
}

// which contacts go to the solver: touching, enabled, solid, attached to a body of an awake island
__global__ void SelectConstraintsKernel(DeviceArrays d, int contactCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, contactCount)
	{
		int sel = 0;
		if (IsSolidTouching(d, i))
		{
			int4 pr = d.c.proxies[i];
			uint32_t fA = d.bflags[pr.z], fB = d.bflags[pr.w];
			if ((!IsStatic(fA) && (fA & B2CU_BODY_ISLAND)) || (!IsStatic(fB) && (fB & B2CU_BODY_ISLAND))) sel = 1;
		}
		d.cSelect[i] = sel;
		if (!sel) d.c.colour[i] = B2CU_COLOUR_NONE;
	}
}

// ---------------------------------------------------------------------------------------------------------
// Graph colouring of the constraint graph (bodies = vertices, constraints = edges; only dynamic bodies
// constrain a colour).  New in this design: it is what lets b2ContactSolver's sequential Gauss-Seidel
// (Box2D/Dynamics/Contacts/b2ContactSolver.cpp:293-603) run one colour at a time in parallel.  Colours persist
// across steps, so only new constraints are coloured.  Deterministic: ties are broken by contact index.
// ---------------------------------------------------------------------------------------------------------
// Colour classes of a sharded world: constraints between own bodies take colours [0, crossBase), constraints that
// touch a ghost body take colours [crossBase, 32); the two classes run in separate phases with a halo exchange
// between them.  Unsharded worlds have crossBase = 32 (one class).  No free colour: overflow list of the class
// (colour 32 / 33), solved serially.
__device__ __forceinline__ uint32_t ColourClassMask(bool cross, int crossBase)
{
	uint32_t low = crossBase >= 32 ? 0xFFFFFFFFu : ((1u << crossBase) - 1u);
	return cross ? ~low : low;
}

__global__ void __launch_bounds__(256) ColourPrepareKernel(DeviceArrays d, const int* __restrict__ list, int* uncoloured,
                                                           int crossBase)
{
	GridDependencyWait();
	__shared__ int hist[B2CU_MAX_COLOURS];
	if (threadIdx.x < B2CU_MAX_COLOURS) hist[threadIdx.x] = 0;
	__syncthreads();
	int n = d.counters[CNT_CONSTRAINT];
	B2CU_GRID_STRIDE(j, n)
	{
		int i = list[j];
		int c = d.c.colour[i];
		int4 pr = d.c.proxies[i];
		int bA = pr.z, bB = pr.w;
		uint32_t fA = d.bflags[bA], fB = d.bflags[bB];
		bool cross = ((fA | fB) & B2CU_BODY_GHOST) != 0;
		if (c >= 0 && c < B2CU_MAX_COLOURS && ((ColourClassMask(cross, crossBase) >> c) & 1u))
		{
			if (IsDynamic(fA)) atomicOr(&d.colourMask[bA], 1u << c);
			if (IsDynamic(fB)) atomicOr(&d.colourMask[bB], 1u << c);
			atomicAdd(&hist[c], 1);
		}
		else
		{
			d.c.colour[i] = B2CU_COLOUR_NONE;
			int slot = atomicAdd(&d.counters[CNT_UNCOLOURED], 1);
			uncoloured[slot] = i;
		}
	}
	__syncthreads();
	if (threadIdx.x < B2CU_MAX_COLOURS && hist[threadIdx.x]) atomicAdd(&d.colourCount[threadIdx.x], hist[threadIdx.x]);
}

__device__ __forceinline__ uint32_t ColourFreeMask(const DeviceArrays& d, int bA, int bB, uint32_t fA, uint32_t fB,
                                                  int crossBase)
{
	uint32_t used = 0u;
	if (IsDynamic(fA)) used |= d.colourMask[bA];
	if (IsDynamic(fB)) used |= d.colourMask[bB];
	bool cross = ((fA | fB) & B2CU_BODY_GHOST) != 0;
	return ~used & ColourClassMask(cross, crossBase);
}

__global__ void ColourProposeKernel(DeviceArrays d, const int* __restrict__ list, int counterIndex, uint32_t round,
                                    int crossBase)
{
	GridDependencyWait();
	int n = d.counters[counterIndex];
	B2CU_GRID_STRIDE(j, n)
	{
		int i = list[j];
		int4 pr = d.c.proxies[i];
		int bA = pr.z, bB = pr.w;
		uint32_t fA = d.bflags[bA], fB = d.bflags[bB];
		uint32_t freeMask = ColourFreeMask(d, bA, bB, fA, fB, crossBase);
		if (freeMask == 0u)
		{
			int overflow = ((fA | fB) & B2CU_BODY_GHOST) ? B2CU_COLOUR_OVERFLOW + 1 : B2CU_COLOUR_OVERFLOW;
			d.c.colour[i] = overflow;
			atomicAdd(&d.colourCount[overflow], 1);
			continue;
		}
		unsigned long long claim = ((unsigned long long)round << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
		if (IsDynamic(fA)) atomicMax(&d.colourClaim[bA], claim);
		if (IsDynamic(fB)) atomicMax(&d.colourClaim[bB], claim);
	}
}

__global__ void ColourCommitKernel(DeviceArrays d, const int* __restrict__ list, int counterIndex, int* next,
                                   int nextCounterIndex, uint32_t round, int crossBase)
{
	GridDependencyWait();
	int n = d.counters[counterIndex];
	B2CU_GRID_STRIDE(j, n)
	{
		int i = list[j];
		if (d.c.colour[i] >= B2CU_COLOUR_OVERFLOW) continue;
		int4 pr = d.c.proxies[i];
		int bA = pr.z, bB = pr.w;
		uint32_t fA = d.bflags[bA], fB = d.bflags[bB];
		bool dynA = IsDynamic(fA), dynB = IsDynamic(fB);
		unsigned long long claim = ((unsigned long long)round << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
		bool win = (!dynA || d.colourClaim[bA] == claim) && (!dynB || d.colourClaim[bB] == claim);
		if (win)
		{
			uint32_t freeMask = ColourFreeMask(d, bA, bB, fA, fB, crossBase);
			int c = __ffs((int)freeMask) - 1;
			d.c.colour[i] = c;
			atomicAdd(&d.colourCount[c], 1);
			// this constraint is the only winner on each of its dynamic bodies this round
			if (dynA) d.colourMask[bA] |= 1u << c;
			if (dynB) d.colourMask[bB] |= 1u << c;
		}
		else
		{
			int slot = atomicAdd(&d.counters[nextCounterIndex], 1);
			next[slot] = i;
		}
	}
}

// Solver order = (colour, class, contact index).  The order inside a colour is free (its constraints share no dynamic
// body); putting like constraints next to each other makes the warps of the solver kernels uniform: the one-point and
// the two-point (block solver) paths of SolveVelocityConstraints, and the three manifold types of the position solver,
// no longer run in the same warp, and a constraint's index tells whether it has a second point to load.
// class: 0 circles manifold, 1 one-point face manifold, 2 two-point face A, 3 two-point face B.
#define B2CU_ORDER_COLOUR_SHIFT 34
__global__ void ColourKeysKernel(DeviceArrays d, const int* __restrict__ list)
{
	GridDependencyWait();
	int n = d.counters[CNT_CONSTRAINT];
	B2CU_GRID_STRIDE(j, n)
	{
		int i = list[j];
		uint4 m3 = d.c.m3[i];
		uint32_t cls = m3.z == B2CU_MANIFOLD_CIRCLES ? 0u : m3.w < 2u ? 1u : m3.z == B2CU_MANIFOLD_FACE_A ? 2u : 3u;
		d.orderKeys[j] = ((uint64_t)(((uint32_t)d.c.colour[i] << 2) | cls) << 32) | (uint32_t)i;
	}
}

// ---------------------------------------------------------------------------------------------------------
// b2Island::Solve, body part 1 (Box2D/Dynamics/b2Island.cpp:192-230): integrate velocities, damping, store
// c0/a0.  Streaming kernel over bodies.
// ---------------------------------------------------------------------------------------------------------
__global__ void IntegrateVelocitiesKernel(DeviceArrays d, int bodyCount, float h, float2 gravity, int flowBase)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		uint32_t bf = d.bflags[b];
		if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) continue;

		float4 p = d.pos[b];
		float4 p0 = d.pos0[b];
		p0.x = p.x;
		p0.y = p.y;
		p0.z = p.z;
		d.pos0[b] = p0;

		if (IsDynamic(bf))
		{
			float4 v = d.vel[b];
			float4 ms = d.mass[b];
			float4 f = d.force[b];
			float4 dm = d.damp[b];
			Vec2 lin = V(v.x, v.y);
			float w = v.z;
			Vec2 acc = dm.z * V(gravity.x, gravity.y) + ms.x * V(f.x, f.y);
			lin = lin + h * acc;
			w = w + h * ms.y * f.z;
			float ld = 1.0f / (1.0f + h * dm.x);
			float ad = 1.0f / (1.0f + h * dm.y);
			lin = V(lin.x * ld, lin.y * ld);
			w = w * ad;
			// .w = 0: the update counter of the dataflow solver (b2cu_solver_flow.cuh) starts here
			d.vel[b] = make_float4(lin.x, lin.y, w, __int_as_float(flowBase));
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// b2ContactSolver constructor + InitializeVelocityConstraints (b2ContactSolver.cpp:47-251) with
// b2WorldManifold::Initialize (Box2D/Collision/b2Collision.cpp:22-86).  One thread per constraint, written
// straight into the colour-sorted SoA.
// ---------------------------------------------------------------------------------------------------------
// cSelect[i] = 1 + position of contact i in the solver order (0: not a constraint)
__global__ void __launch_bounds__(256) ConstraintSlotKernel(DeviceArrays d)
{
	GridDependencyWait();
	int n = d.counters[CNT_CONSTRAINT];
	B2CU_GRID_STRIDE(k, n) { d.cSelect[(int)(uint32_t)(d.orderKeys[k] & 0xFFFFFFFFull)] = k + 1; }
}

// Runs in CONTACT order (coalesced manifold reads, neighbouring contacts share their bodies) and writes row k of the
// solver arrays: within one colour the solver order is the contact order, so the rows written by neighbouring
// threads are neighbours too and the partial sectors merge in L2.
#ifndef B2CU_INIT_BLOCKS
#define B2CU_INIT_BLOCKS 3
#endif
// the rows are written once here and read next by the solver kernels, 400 MB later: streaming stores keep them from
// pushing the body rows this kernel gathers out of L2
#ifndef B2CU_ROW_STORE_PLAIN
#define B2CU_ROW_STORE(p, v) __stcs(p, v)
#else
#define B2CU_ROW_STORE(p, v) (*(p) = (v))
#endif
__global__ void __launch_bounds__(256, B2CU_INIT_BLOCKS) InitConstraintsKernel(DeviceArrays d, const int* __restrict__ list,
                                                                               float dtRatio, int warmStarting)
{
	GridDependencyWait();
	// list = the constraint contacts in ascending contact order (the compaction of cSelect)
	int n = d.counters[CNT_CONSTRAINT];
	B2CU_GRID_STRIDE(j, n)
	{
		int i = list[j];
		int k = d.cSelect[i] - 1;
		int4 pr = d.c.proxies[i];
		int bA = pr.z, bB = pr.w;
		float radiusA = d.pradius[pr.x];
		float radiusB = d.pradius[pr.y];

		float4 m0 = d.c.m0[i], m1 = d.c.m1[i], m2 = d.c.m2[i];
		uint4 m3 = d.c.m3[i];
		float4 mix = d.c.mix[i];
		int type = (int)m3.z;
		int pointCount = (int)m3.w;

		float4 msA = d.mass[bA], msB = d.mass[bB];
		float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
		float4 pA = d.pos[bA], pB = d.pos[bB];
		float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
		Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
		float aA = pA.z, aB = pB.z;
		Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
		float wA = vA4.z, wB = vB4.z;
		Vec2 localCenterA = V(msA.z, msA.w), localCenterB = V(msB.z, msB.w);

		Xf xfA, xfB;
		xfA.q = SinCos(aA);
		xfB.q = SinCos(aB);
		xfA.p = cA - Mul(xfA.q, localCenterA);
		xfB.p = cB - Mul(xfB.q, localCenterB);

		Vec2 localNormal = V(m0.x, m0.y), localPoint = V(m0.z, m0.w);
		Vec2 lp[2] = {V(m1.x, m1.y), V(m2.x, m2.y)};

		// b2WorldManifold::Initialize
		Vec2 normal = V(1.0f, 0.0f);
		Vec2 wp[2] = {V(0.0f, 0.0f), V(0.0f, 0.0f)};
		if (type == B2CU_MANIFOLD_CIRCLES)
		{
			Vec2 pointA = Mul(xfA, localPoint);
			Vec2 pointB = Mul(xfB, lp[0]);
			if (DistanceSquared(pointA, pointB) > B2CU_EPSILON * B2CU_EPSILON)
			{
				normal = Normalized(pointB - pointA);
			}
			Vec2 ccA = pointA + radiusA * normal;
			Vec2 ccB = pointB - radiusB * normal;
			wp[0] = 0.5f * (ccA + ccB);
		}
		else if (type == B2CU_MANIFOLD_FACE_A)
		{
			normal = Mul(xfA.q, localNormal);
			Vec2 planePoint = Mul(xfA, localPoint);
			for (int j = 0; j < pointCount; ++j)
			{
				Vec2 clipPoint = Mul(xfB, lp[j]);
				Vec2 ccA = clipPoint + (radiusA - Dot(clipPoint - planePoint, normal)) * normal;
				Vec2 ccB = clipPoint - radiusB * normal;
				wp[j] = 0.5f * (ccA + ccB);
			}
		}
		else
		{
			normal = Mul(xfB.q, localNormal);
			Vec2 planePoint = Mul(xfB, localPoint);
			for (int j = 0; j < pointCount; ++j)
			{
				Vec2 clipPoint = Mul(xfA, lp[j]);
				Vec2 ccB = clipPoint + (radiusB - Dot(clipPoint - planePoint, normal)) * normal;
				Vec2 ccA = clipPoint - radiusA * normal;
				wp[j] = 0.5f * (ccA + ccB);
			}
			normal = -normal;
		}

		float friction = mix.x, restitution = mix.y, tangentSpeed = mix.z;

		float4 pa[2], pb[2];
		float imp[4];
		float oldImp[4] = {m1.z, m1.w, m2.z, m2.w};
		Vec2 rAs[2], rBs[2];
		for (int j = 0; j < 2; ++j)
		{
			pa[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			pb[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			imp[2 * j] = 0.0f;
			imp[2 * j + 1] = 0.0f;
			rAs[j] = V(0.0f, 0.0f);
			rBs[j] = V(0.0f, 0.0f);
		}
		for (int j = 0; j < pointCount; ++j)
		{
			if (warmStarting)
			{
				imp[2 * j] = dtRatio * oldImp[2 * j];
				imp[2 * j + 1] = dtRatio * oldImp[2 * j + 1];
			}
			Vec2 rA = wp[j] - cA;
			Vec2 rB = wp[j] - cB;
			rAs[j] = rA;
			rBs[j] = rB;

			float rnA = Cross(rA, normal);
			float rnB = Cross(rB, normal);
			float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
			float normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;

			Vec2 tangent = CrossVS(normal, 1.0f);
			float rtA = Cross(rA, tangent);
			float rtB = Cross(rB, tangent);
			float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
			float tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;

			float velocityBias = 0.0f;
			float vRel = Dot(normal, vB + CrossSV(wB, rB) - vA - CrossSV(wA, rA));
			if (vRel < -B2CU_VELOCITY_THRESHOLD)
			{
				velocityBias = -restitution * vRel;
			}
			pa[j] = make_float4(rA.x, rA.y, rB.x, rB.y);
			pb[j] = make_float4(normalMass, tangentMass, velocityBias, 0.0f);
		}

		int solvePoints = pointCount;
		float4 K = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		float4 NM = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		if (pointCount == 2)
		{
			float rn1A = Cross(rAs[0], normal);
			float rn1B = Cross(rBs[0], normal);
			float rn2A = Cross(rAs[1], normal);
			float rn2B = Cross(rBs[1], normal);

			float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
			float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
			float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;

			const float k_maxConditionNumber = 1000.0f;
			if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12))
			{
				K = make_float4(k11, k12, k22, 0.0f);
				// b2Mat22::GetInverse (b2Math.h:205-216) of [[k11 k12][k12 k22]]
				float a = k11, b = k12, c = k12, dd = k22;
				float det = a * dd - b * c;
				if (det != 0.0f)
				{
					det = 1.0f / det;
				}
				NM = make_float4(det * dd, -det * b, -det * c, det * a); // ex.x, ey.x, ex.y, ey.y
			}
			else
			{
				solvePoints = 1;
			}
		}

		uint32_t bfA = d.bflags[bA];
		int root = d.island[IsStatic(bfA) ? bB : bA];

		B2CU_ROW_STORE(&d.solverKeys[k], d.c.key[i]);
		B2CU_ROW_STORE(&d.sBody[k], make_int4(bA, bB, i, solvePoints | (pointCount << 8)));
		B2CU_ROW_STORE(&d.sMass[k], make_float4(mA, iA, mB, iB));
		B2CU_ROW_STORE(&d.sNormal[k], make_float4(normal.x, normal.y, friction, tangentSpeed));
		B2CU_ROW_STORE(&d.sP0a[k], pa[0]);
		// second point: only rA / rB and the velocity bias are kept; its masses and the block solver's K and inverse are
		// recomputed by the solver from rA, rB, the normal and the body masses (the same expressions on the same values:
		// the same bits), which takes 48 bytes per iteration off every two-point constraint
		B2CU_ROW_STORE(&d.sP0b[k], make_float4(pb[0].x, pb[0].y, pb[0].z, pb[1].z));
		B2CU_ROW_STORE(&d.sImp[k], make_float4(imp[0], imp[1], imp[2], imp[3]));
		if (pointCount == 2)
		{
			B2CU_ROW_STORE(&d.sP1a[k], pa[1]);
			// first two-point row of the colour (rows are ordered one-point first inside a colour, ColourKeysKernel)
			const int colour = d.c.colour[i];
			const unsigned peers = __match_any_sync(__activemask(), colour);
			const int first = __reduce_min_sync(peers, k);
			if ((int)(threadIdx.x & 31) == __ffs((int)peers) - 1) atomicMin(&d.colourTwoStart[colour], first);
		}
		(void)K;
		(void)NM;
		B2CU_ROW_STORE(&d.sLocal[k], m0);
		B2CU_ROW_STORE(&d.sLocalP[k], make_float4(lp[0].x, lp[0].y, lp[1].x, lp[1].y));
		B2CU_ROW_STORE(&d.sCenters[k], make_float4(localCenterA.x, localCenterA.y, localCenterB.x, localCenterB.y));
		B2CU_ROW_STORE(&d.sRadius[k], make_float4(radiusA, radiusB, __int_as_float(type), __int_as_float(root)));
	}
}

// b2ContactSolver::WarmStart (b2ContactSolver.cpp:253-291), constraints [begin, begin+count) of one colour
// the per-constraint rows every velocity pass needs; loaded ahead of the colour barrier by the persistent solver
struct VelPre
{
	int4 sb;
	float4 ms, nf, imp, p0a, p0b, p1a;
};
// two: the row is in the two-point part of its colour (k >= colourTwoStart[colour]): its second point is loaded in the
// same breath as the rest instead of after the point count has arrived
__device__ __forceinline__ VelPre LoadVelPre(const DeviceArrays& d, int k, bool two)
{
	VelPre p;
	p.sb = __ldcs(&d.sBody[k]);
	p.ms = __ldcs(&d.sMass[k]);
	p.nf = __ldcs(&d.sNormal[k]);
	p.imp = __ldcs(&d.sImp[k]);
	p.p0a = __ldcs(&d.sP0a[k]);
	p.p0b = __ldcs(&d.sP0b[k]);
	p.p1a = two ? __ldcs(&d.sP1a[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
	return p;
}
// without the hint: decided by the row itself (one dependent load more for two-point rows)
__device__ __forceinline__ VelPre LoadVelPre(const DeviceArrays& d, int k)
{
	VelPre p = LoadVelPre(d, k, false);
	if ((p.sb.w >> 8) == 2) p.p1a = __ldcs(&d.sP1a[k]);
	return p;
}

// the arithmetic of one constraint's warm start on the two bodies' velocity rows (x, y = v, z = w; .w is not touched)
__device__ __forceinline__ void WarmStartCore(const DeviceArrays& d, int k, const VelPre& pre, float4& vA4, float4& vB4)
{
	int4 sb = pre.sb;
	float4 ms = pre.ms;
	float4 nf = pre.nf;
	float4 imp = pre.imp;
	int pointCount = sb.w & 0xFF;
	float mA = ms.x, iA = ms.y, mB = ms.z, iB = ms.w;

	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	Vec2 normal = V(nf.x, nf.y);
	Vec2 tangent = CrossVS(normal, 1.0f);

	for (int j = 0; j < pointCount; ++j)
	{
		float4 r = j == 0 ? pre.p0a : pre.p1a;
		float ni = j == 0 ? imp.x : imp.z;
		float ti = j == 0 ? imp.y : imp.w;
		Vec2 rA = V(r.x, r.y), rB = V(r.z, r.w);
		Vec2 P = ni * normal + ti * tangent;
		wA -= iA * Cross(rA, P);
		vA = vA - mA * P;
		wB += iB * Cross(rB, P);
		vB = vB + mB * P;
	}
	vA4.x = vA.x; vA4.y = vA.y; vA4.z = wA;
	vB4.x = vB.x; vB4.y = vB.y; vB4.z = wB;
}

__device__ __forceinline__ void WarmStartPre(const DeviceArrays& d, int k, const VelPre& pre)
{
	const int4 sb = pre.sb;
	const float4 ms = pre.ms;
	float4 vA4 = d.vel[sb.x], vB4 = d.vel[sb.y];
	WarmStartCore(d, k, pre, vA4, vB4);
	if (ms.x != 0.0f || ms.y != 0.0f) d.vel[sb.x] = vA4;
	if (ms.z != 0.0f || ms.w != 0.0f) d.vel[sb.y] = vB4;
}

__device__ __forceinline__ void WarmStartOne(const DeviceArrays& d, int k) { WarmStartPre(d, k, LoadVelPre(d, k)); }

__global__ void __launch_bounds__(256) WarmStartKernel(DeviceArrays d, int begin, int count)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(t, count) { WarmStartOne(d, begin + t); }
}

// b2ContactSolver::SolveVelocityConstraints (b2ContactSolver.cpp:293-603) for one constraint
__device__ __forceinline__ void SolveVelocityCore(const DeviceArrays& d, int k, const VelPre& pre, float4& vA4, float4& vB4)
{
	int4 sb = pre.sb;
	float4 ms = pre.ms;
	float4 nf = pre.nf;
	float4 imp = pre.imp;
	float4 p0a = pre.p0a, p0b = pre.p0b;
	int pointCount = sb.w & 0xFF;
	float mA = ms.x, iA = ms.y, mB = ms.z, iB = ms.w;

	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	Vec2 normal = V(nf.x, nf.y);
	Vec2 tangent = CrossVS(normal, 1.0f);
	float friction = nf.z, tangentSpeed = nf.w;

	// second point: rA / rB from the row, bias from p0b.w, the two masses recomputed as InitializeVelocityConstraints
	// computes them (b2ContactSolver.cpp:194-210)
	const float4 p1a = pre.p1a;
	float4 p1b = make_float4(0.f, 0.f, p0b.w, 0.f);
	if (pointCount == 2)
	{
		Vec2 rA = V(p1a.x, p1a.y), rB = V(p1a.z, p1a.w);
		float rnA = Cross(rA, normal);
		float rnB = Cross(rB, normal);
		float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
		p1b.x = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
		float rtA = Cross(rA, tangent);
		float rtB = Cross(rB, tangent);
		float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
		p1b.y = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
	}

	// tangent constraints first
	for (int j = 0; j < pointCount; ++j)
	{
		float4 r = j == 0 ? p0a : p1a;
		float4 q = j == 0 ? p0b : p1b;
		Vec2 rA = V(r.x, r.y), rB = V(r.z, r.w);
		float normalImpulse = j == 0 ? imp.x : imp.z;
		float tangentImpulse = j == 0 ? imp.y : imp.w;

		Vec2 dv = vB + CrossSV(wB, rB) - vA - CrossSV(wA, rA);
		float vt = Dot(dv, tangent) - tangentSpeed;
		float lambda = q.y * (-vt);

		float maxFriction = friction * normalImpulse;
		float newImpulse = Clamp(tangentImpulse + lambda, -maxFriction, maxFriction);
		lambda = newImpulse - tangentImpulse;
		if (j == 0) imp.y = newImpulse;
		else imp.w = newImpulse;

		Vec2 P = lambda * tangent;
		vA = vA - mA * P;
		wA -= iA * Cross(rA, P);
		vB = vB + mB * P;
		wB += iB * Cross(rB, P);
	}

	if (pointCount == 1)
	{
		Vec2 rA = V(p0a.x, p0a.y), rB = V(p0a.z, p0a.w);
		Vec2 dv = vB + CrossSV(wB, rB) - vA - CrossSV(wA, rA);
		float vn = Dot(dv, normal);
		float lambda = -p0b.x * (vn - p0b.z);

		float newImpulse = Max(imp.x + lambda, 0.0f);
		lambda = newImpulse - imp.x;
		imp.x = newImpulse;

		Vec2 P = lambda * normal;
		vA = vA - mA * P;
		wA -= iA * Cross(rA, P);
		vB = vB + mB * P;
		wB += iB * Cross(rB, P);
	}
	else
	{
		// block solver, total enumeration of the 2x2 LCP (b2ContactSolver.cpp:375-596)
		Vec2 rA1 = V(p0a.x, p0a.y), rB1 = V(p0a.z, p0a.w);
		Vec2 rA2 = V(p1a.x, p1a.y), rB2 = V(p1a.z, p1a.w);
		// K and its inverse as InitializeVelocityConstraints prepares them (b2ContactSolver.cpp:214-244)
		float4 Kq, NM;
		{
			float rn1A = Cross(rA1, normal);
			float rn1B = Cross(rB1, normal);
			float rn2A = Cross(rA2, normal);
			float rn2B = Cross(rB2, normal);
			float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
			float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
			float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
			Kq = make_float4(k11, k12, k22, 0.0f);
			float a = k11, b = k12, c = k12, dd = k22;
			float det = a * dd - b * c;
			if (det != 0.0f)
			{
				det = 1.0f / det;
			}
			NM = make_float4(det * dd, -det * b, -det * c, det * a);
		}

		Vec2 a = V(imp.x, imp.z);

		Vec2 dv1 = vB + CrossSV(wB, rB1) - vA - CrossSV(wA, rA1);
		Vec2 dv2 = vB + CrossSV(wB, rB2) - vA - CrossSV(wA, rA2);
		float vn1 = Dot(dv1, normal);
		float vn2 = Dot(dv2, normal);

		Vec2 b;
		b.x = vn1 - p0b.z;
		b.y = vn2 - p1b.z;
		// b -= K * a, K = [ex=(k11,k12), ey=(k12,k22)]
		b.x -= Kq.x * a.x + Kq.y * a.y;
		b.y -= Kq.y * a.x + Kq.z * a.y;

		Vec2 x;
		bool found = false;
		// case 1: x = -normalMass * b
		x.x = -(NM.x * b.x + NM.y * b.y);
		x.y = -(NM.z * b.x + NM.w * b.y);
		if (x.x >= 0.0f && x.y >= 0.0f)
		{
			found = true;
		}
		if (!found)
		{
			// case 2: vn1 = 0, x2 = 0
			x.x = -p0b.x * b.x;
			x.y = 0.0f;
			vn2 = Kq.y * x.x + b.y;
			if (x.x >= 0.0f && vn2 >= 0.0f) found = true;
		}
		if (!found)
		{
			// case 3: vn2 = 0, x1 = 0
			x.x = 0.0f;
			x.y = -p1b.x * b.y;
			vn1 = Kq.y * x.y + b.x;
			if (x.y >= 0.0f && vn1 >= 0.0f) found = true;
		}
		if (!found)
		{
			// case 4: x1 = 0, x2 = 0
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
			imp.x = x.x;
			imp.z = x.y;
		}
		// no solution: give up, as the reference does (:593-594)
	}

	__stcs(&d.sImp[k], imp);
	vA4.x = vA.x; vA4.y = vA.y; vA4.z = wA;
	vB4.x = vB.x; vB4.y = vB.y; vB4.z = wB;
}


__device__ __forceinline__ void SolveVelocityPre(const DeviceArrays& d, int k, const VelPre& pre)
{
	const int4 sb = pre.sb;
	const float4 ms = pre.ms;
	float4 vA4 = d.vel[sb.x], vB4 = d.vel[sb.y];
	SolveVelocityCore(d, k, pre, vA4, vB4);
	if (ms.x != 0.0f || ms.y != 0.0f) d.vel[sb.x] = vA4;
	if (ms.z != 0.0f || ms.w != 0.0f) d.vel[sb.y] = vB4;
}

__device__ __forceinline__ void SolveVelocityOne(const DeviceArrays& d, int k) { SolveVelocityPre(d, k, LoadVelPre(d, k)); }

__global__ void __launch_bounds__(256) SolveVelocityKernel(DeviceArrays d, int begin, int count)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(t, count) { SolveVelocityOne(d, begin + t); }
}

// b2ContactSolver::StoreImpulses (b2ContactSolver.cpp:605-618)
__global__ void StoreImpulsesKernel(DeviceArrays d)
{
	GridDependencyWait();
	int n = d.counters[CNT_CONSTRAINT];
	B2CU_GRID_STRIDE(k, n)
	{
		int4 sb = d.sBody[k];
		float4 imp = d.sImp[k];
		int i = sb.z;
		int pointCount = sb.w & 0xFF;
		float4 m1 = d.c.m1[i];
		m1.z = imp.x;
		m1.w = imp.y;
		d.c.m1[i] = m1;
		if (pointCount == 2)
		{
			float4 m2 = d.c.m2[i];
			m2.z = imp.z;
			m2.w = imp.w;
			d.c.m2[i] = m2;
		}
	}
}

// b2Island::Solve, body part 2 (b2Island.cpp:283-313): clamp and integrate positions
__global__ void IntegratePositionsKernel(DeviceArrays d, int bodyCount, float h)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		uint32_t bf = d.bflags[b];
		if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) continue;
		float4 p = d.pos[b];
		float4 v4 = d.vel[b];
		Vec2 c = V(p.x, p.y);
		float a = p.z;
		Vec2 v = V(v4.x, v4.y);
		float w = v4.z;

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

		d.pos[b] = make_float4(c.x, c.y, a, p.w);
		d.vel[b] = make_float4(v.x, v.y, w, v4.w);
	}
}

// b2ContactSolver::SolvePositionConstraints (b2ContactSolver.cpp:676-752) for one constraint.
// Returns the smallest separation seen; islands stop iterating once theirs is >= -3*linearSlop
// (b2Island.cpp:318-335), which is tracked per island root in islandMinSep[iteration][root].
struct PosPre
{
	int4 sb;
	float4 ms, loc, lps, cen, rad;
};
__device__ __forceinline__ PosPre LoadPosPre(const DeviceArrays& d, int k)
{
	PosPre p;
	p.sb = __ldcs(&d.sBody[k]);
	p.ms = __ldcs(&d.sMass[k]);
	p.loc = __ldcs(&d.sLocal[k]);
	p.lps = __ldcs(&d.sLocalP[k]);
	p.cen = __ldcs(&d.sCenters[k]);
	p.rad = __ldcs(&d.sRadius[k]);
	return p;
}

// the arithmetic of one constraint's position correction on the two bodies' position rows (x, y = c, z = a)
__device__ __forceinline__ float SolvePositionCore(const PosPre& pre, float4& pA4, float4& pB4)
{
	int4 sb = pre.sb;
	float4 ms = pre.ms;
	float4 loc = pre.loc;
	float4 lps = pre.lps;
	float4 cen = pre.cen;
	float4 rad = pre.rad;
	int pointCount = sb.w >> 8;
	int type = __float_as_int(rad.z);
	float mA = ms.x, iA = ms.y, mB = ms.z, iB = ms.w;
	Vec2 localCenterA = V(cen.x, cen.y), localCenterB = V(cen.z, cen.w);
	Vec2 localNormal = V(loc.x, loc.y), localPoint = V(loc.z, loc.w);

	Vec2 cA = V(pA4.x, pA4.y), cB = V(pB4.x, pB4.y);
	float aA = pA4.z, aB = pB4.z;
	float minSeparation = B2CU_MAX_FLOAT;

	for (int j = 0; j < pointCount; ++j)
	{
		Xf xfA, xfB;
		xfA.q = SinCos(aA);
		xfB.q = SinCos(aB);
		xfA.p = cA - Mul(xfA.q, localCenterA);
		xfB.p = cB - Mul(xfB.q, localCenterB);

		Vec2 normal, point;
		float separation;
		Vec2 lpj = j == 0 ? V(lps.x, lps.y) : V(lps.z, lps.w);
		if (type == B2CU_MANIFOLD_CIRCLES)
		{
			Vec2 pointA = Mul(xfA, localPoint);
			Vec2 pointB = Mul(xfB, V(lps.x, lps.y));
			normal = Normalized(pointB - pointA);
			point = 0.5f * (pointA + pointB);
			separation = Dot(pointB - pointA, normal) - rad.x - rad.y;
		}
		else if (type == B2CU_MANIFOLD_FACE_A)
		{
			normal = Mul(xfA.q, localNormal);
			Vec2 planePoint = Mul(xfA, localPoint);
			Vec2 clipPoint = Mul(xfB, lpj);
			separation = Dot(clipPoint - planePoint, normal) - rad.x - rad.y;
			point = clipPoint;
		}
		else
		{
			normal = Mul(xfB.q, localNormal);
			Vec2 planePoint = Mul(xfB, localPoint);
			Vec2 clipPoint = Mul(xfA, lpj);
			separation = Dot(clipPoint - planePoint, normal) - rad.x - rad.y;
			point = clipPoint;
			normal = -normal;
		}

		Vec2 rA = point - cA;
		Vec2 rB = point - cB;

		minSeparation = Min(minSeparation, separation);

		float C = Clamp(B2CU_BAUMGARTE * (separation + B2CU_LINEAR_SLOP), -B2CU_MAX_LINEAR_CORRECTION, 0.0f);

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

	pA4.x = cA.x; pA4.y = cA.y; pA4.z = aA;
	pB4.x = cB.x; pB4.y = cB.y; pB4.z = aB;
	return minSeparation;
}


__device__ __forceinline__ float SolvePositionPre(const DeviceArrays& d, const PosPre& pre)
{
	const int4 sb = pre.sb;
	const float4 ms = pre.ms;
	float4 pA4 = d.pos[sb.x], pB4 = d.pos[sb.y];
	float minSeparation = SolvePositionCore(pre, pA4, pB4);
	if (ms.x != 0.0f || ms.y != 0.0f) d.pos[sb.x] = pA4;
	if (ms.z != 0.0f || ms.w != 0.0f) d.pos[sb.y] = pB4;
	return minSeparation;
}

__device__ __forceinline__ float SolvePositionOne(const DeviceArrays& d, int k) { return SolvePositionPre(d, LoadPosPre(d, k)); }

__device__ __forceinline__ bool IslandDone(const DeviceArrays& d, int iteration, int root, int bodyCount)
{
	if (iteration == 0) return false;
	float prev = OrderedToFloat(d.islandMinSep[(iteration - 1) * bodyCount + root]);
	return prev >= -3.0f * B2CU_LINEAR_SLOP;
}

__global__ void __launch_bounds__(256) SolvePositionKernel(DeviceArrays d, int begin, int count, int iteration,
                                                           int bodyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(t, count)
	{
		int k = begin + t;
		int root = __float_as_int(__ldcs(&d.sRadius[k]).w);
		if (IslandDone(d, iteration, root, bodyCount)) continue;
		float minSep = SolvePositionOne(d, k);
		AtomicMinByRoot(d.islandMinSep + (size_t)iteration * bodyCount, root, FloatToOrdered(minSep));
	}
}

// overflow constraints (no free colour): solved one after the other by a single thread, in key order
__global__ void OverflowWarmStartKernel(DeviceArrays d, int begin, int count)
{
	GridDependencyWait();
	if (blockIdx.x == 0 && threadIdx.x == 0)
		for (int t = 0; t < count; ++t) WarmStartOne(d, begin + t);
}
__global__ void OverflowSolveVelocityKernel(DeviceArrays d, int begin, int count)
{
	GridDependencyWait();
	if (blockIdx.x == 0 && threadIdx.x == 0)
		for (int t = 0; t < count; ++t) SolveVelocityOne(d, begin + t);
}
__global__ void OverflowSolvePositionKernel(DeviceArrays d, int begin, int count, int iteration, int bodyCount)
{
	GridDependencyWait();
	if (blockIdx.x == 0 && threadIdx.x == 0)
		for (int t = 0; t < count; ++t)
		{
			int k = begin + t;
			int root = __float_as_int(d.sRadius[k].w);
			if (IslandDone(d, iteration, root, bodyCount)) continue;
			float minSep = SolvePositionOne(d, k);
			atomicMin(&d.islandMinSep[iteration * bodyCount + root], FloatToOrdered(minSep));
		}
}

// ---------------------------------------------------------------------------------------------------------
// The whole constraint solve of a step in ONE persistent cooperative kernel: warm start, velocity iterations,
// impulse store, position integration and position iterations, one grid-wide barrier per colour phase instead
// of one kernel launch per colour phase.  A pile is one island of ~12 colours, i.e. ~150 dependent phases of
// ~200k constraints each: each phase moves only ~25 MB, so launch latency, not bandwidth, bounded the
// multi-launch version.  Same per-constraint code (WarmStartOne / SolveVelocityOne / SolvePositionOne), same
// order, same results.
// ---------------------------------------------------------------------------------------------------------
// Halo exchange state of a sharded world (b2cuShardConfigure / Connect).  Mailboxes are plain device buffers;
// the ones of the neighbours are mapped peer memory (NVLink): a push is stores + __threadfence_system + a
// sequence flag, a receive is a spin on the local flag.
struct ShardState
{
	int rankCount;   // 1: unsharded, no exchange
	int ghostCount, exportCount;
	const int* ghostIds;   // bodies of this world that are copies of the upper neighbour's bodies
	const int* exportIds;  // own bodies that are ghosts in the lower neighbour
	float4* fromUpper;     // local mailbox written by the upper neighbour (owner -> ghost data)
	float4* fromLower;     // local mailbox written by the lower neighbour (ghost -> owner data)
	unsigned* flagFromUpper;
	unsigned* flagFromLower;
	float4* lowerFromUpper;   // lower neighbour's fromUpper mailbox (peer memory), nullptr on rank 0
	unsigned* lowerFlagFromUpper;
	float4* upperFromLower;   // upper neighbour's fromLower mailbox (peer memory), nullptr on the last rank
	unsigned* upperFlagFromLower;
	unsigned seq;             // sequence number of the next exchange
	int* stuck;               // set when a wait for a neighbour gave up (the step then fails instead of hanging the GPU)
};

enum SolverOpType
{
	OP_PARALLEL = 0,   // one colour: constraints [start, start+count) in parallel
	OP_SERIAL = 1,     // overflow list: one thread, in order
	OP_PUSH_DOWN = 2,  // halo: owners -> ghost copies (to the lower neighbour), then receive from the upper one
	OP_PUSH_UP = 3     // halo: ghost copies -> owners (to the upper neighbour), then receive from the lower one
};

#define B2CU_MAX_SOLVER_OPS (B2CU_MAX_COLOURS + 6)

struct SolverPlan
{
	int opCount;
	int opType[B2CU_MAX_SOLVER_OPS];
	int opStart[B2CU_MAX_SOLVER_OPS];
	int opSize[B2CU_MAX_SOLVER_OPS];
	int opColour[B2CU_MAX_SOLVER_OPS]; // colour of a parallel op (the dataflow kernels derive the bodies' update ranks from it)
	int constraintCount;
	int bodyCount;
	int velocityIterations;
	int positionIterations;
	int warmStarting;
	float h;
	ShardState shard;
	// joints: colour classes of DeviceArrays::jointOrder, solved before the contacts in a velocity iteration and after
	// them in a position iteration (b2Island.cpp:259-273, :323-327, :363-380)
	int jointOpCount;
	int jointOpStart[B2CU_MAX_JOINT_OPS], jointOpSize[B2CU_MAX_JOINT_OPS], jointOpSerial[B2CU_MAX_JOINT_OPS];
	float dtRatio;
	int ovStart, ovCount;  // dataflow kernels: rows of the serial overflow list (unsharded worlds)
	const int* ovRank;     // [2 * ovCount]: overflow rows before row j on its body A / body B
	const int* ovDeg;      // per body: its overflow rows
	int flowBase;          // dataflow kernels: the versions of this step start at this value
	unsigned* softBarrier; // counter of GridSync (zeroed before the launch); nullptr: cooperative launch
	int flowPrefetch;   // dataflow kernels: L2 prefetch of the next round's rows (B2CU_FLOW_PREFETCH)
	int debugSkipStore; // timing experiments only (B2CU_DEBUG_SKIP_STORE): leave the impulse store out
};

// Grid-wide barrier of the persistent solver kernels.  Normally cooperative_groups' (the kernel is a cooperative
// launch).  A sharded world whose neighbours share its device cannot use cooperative launches -- the driver runs
// cooperative grids of one device one after the other, and shards wait for each other INSIDE their kernels -- so there
// the kernels are ordinary launches sized to stay co-resident (b2cuShardConfigure's gridFraction) and synchronise on a
// counter of their own: every CTA adds one, thread 0 of each spins until the count reaches the round's target.
struct GridSync
{
	unsigned* counter; // nullptr: cooperative launch
	unsigned blocks;
	unsigned round;
	int* stuck;
	__device__ __forceinline__ void sync()
	{
		if (counter == nullptr)
		{
			cooperative_groups::this_grid().sync();
			return;
		}
		__syncthreads();
		if (threadIdx.x == 0)
		{
			__threadfence();
			const unsigned target = (round + 1u) * blocks;
			atomicAdd(counter, 1u);
			long long spins = 0;
			while (*reinterpret_cast<volatile unsigned*>(counter) < target)
			{
				if (*reinterpret_cast<volatile int*>(stuck) != 0) break;
				if (++spins > (1ll << 25))
				{
					*reinterpret_cast<volatile int*>(stuck) = 3;
					break;
				}
			}
			__threadfence();
		}
		++round;
		__syncthreads();
	}
};

enum { JOINT_INIT = 0, JOINT_VELOCITY = 1, JOINT_POSITION = 2 };

__device__ __forceinline__ void JointRunOne(const DeviceArrays& d, const SolverPlan& plan, int mode, int iteration, int j)
{
	if (mode == JOINT_INIT)
	{
		JointInitOne(d, j, plan.dtRatio, plan.warmStarting, plan.h);
	}
	else if (mode == JOINT_VELOCITY)
	{
		JointSolveVelocityOne(d, j, plan.h);
	}
	else
	{
		JointRow r = d.jointRows[j];
		if (!r.solved || IslandDone(d, iteration, r.root, plan.bodyCount)) return;
		// a joint out of tolerance keeps its island iterating, like a contact deeper than 3 slops
		if (!JointSolvePositionOne(d, j, r))
			atomicMin(&d.islandMinSep[(size_t)iteration * plan.bodyCount + r.root], FloatToOrdered(-B2CU_MAX_FLOAT));
	}
}

// all joint classes of one iteration; every thread of the grid takes part in the barriers
__device__ __forceinline__ void JointRunOps(GridSync& grid, const DeviceArrays& d, const SolverPlan& plan,
                                            int mode, int iteration, int tid, int stride)
{
	for (int jo = 0; jo < plan.jointOpCount; ++jo)
	{
		const int begin = plan.jointOpStart[jo], n = plan.jointOpSize[jo];
		if (!plan.jointOpSerial[jo])
		{
			for (int t = tid; t < n; t += stride) JointRunOne(d, plan, mode, iteration, d.jointOrder[begin + t]);
		}
		else if (tid == 0)
		{
			for (int t = 0; t < n; ++t) JointRunOne(d, plan, mode, iteration, d.jointOrder[begin + t]);
		}
		grid.sync();
	}
}

__device__ __forceinline__ void IntegratePositionOne(const DeviceArrays& d, int b, float h)
{
	uint32_t bf = d.bflags[b];
	if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) return;
	float4 p = d.pos[b];
	float4 v4 = d.vel[b];
	Vec2 c = V(p.x, p.y);
	float a = p.z;
	Vec2 v = V(v4.x, v4.y);
	float w = v4.z;

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

	d.pos[b] = make_float4(c.x, c.y, a, 0.0f); // .w = 0: update counter of the dataflow position solver
	d.vel[b] = make_float4(v.x, v.y, w, v4.w);
}

__device__ __forceinline__ void StoreImpulseOne(const DeviceArrays& d, int k)
{
	int4 sb = d.sBody[k];
	float4 imp = d.sImp[k];
	int i = sb.z;
	int pointCount = sb.w & 0xFF;
	float4 m1 = d.c.m1[i];
	m1.z = imp.x;
	m1.w = imp.y;
	d.c.m1[i] = m1;
	if (pointCount == 2)
	{
		float4 m2 = d.c.m2[i];
		m2.z = imp.z;
		m2.w = imp.w;
		d.c.m2[i] = m2;
	}
}

// field[ids[k]] -> peer mailbox, flag; then wait for the neighbour's push and scatter it into field
__device__ __forceinline__ void HaloExchange(GridSync& grid, const ShardState& sh, float4* field,
                                             bool down, unsigned seq, int tid, int stride)
{
	// send
	const int* sendIds = down ? sh.exportIds : sh.ghostIds;
	const int sendCount = down ? sh.exportCount : sh.ghostCount;
	float4* sendBox = down ? sh.lowerFromUpper : sh.upperFromLower;
	unsigned* sendFlag = down ? sh.lowerFlagFromUpper : sh.upperFlagFromLower;
	if (sendBox != nullptr)
	{
		for (int k = tid; k < sendCount; k += stride) sendBox[k] = field[sendIds[k]];
		__threadfence_system();
	}
	grid.sync();
	if (sendBox != nullptr && tid == 0)
	{
		*reinterpret_cast<volatile unsigned*>(sendFlag) = seq;
		__threadfence_system();
	}
	// receive: every CTA watches the local flag itself (one thread each, an L2-resident word), so no barrier is needed
	// between "the neighbour's push has arrived" and the scatter
	const int* recvIds = down ? sh.ghostIds : sh.exportIds;
	const int recvCount = down ? sh.ghostCount : sh.exportCount;
	const float4* recvBox = down ? sh.fromUpper : sh.fromLower;
	const unsigned* recvFlag = down ? sh.flagFromUpper : sh.flagFromLower;
	const bool hasPeer = down ? (sh.upperFromLower != nullptr) : (sh.lowerFromUpper != nullptr);
	if (hasPeer)
	{
		if (threadIdx.x == 0)
		{
			// bounded: a neighbour that never arrives (its kernel not co-resident, its process gone) must not hang the GPU
			long long spins = 0;
			while (*reinterpret_cast<const volatile unsigned*>(recvFlag) < seq)
			{
				if (*reinterpret_cast<volatile int*>(sh.stuck) != 0) break;
				if (++spins > (1ll << 25))
				{
					*reinterpret_cast<volatile int*>(sh.stuck) = 2;
					break;
				}
			}
			__threadfence_system();
		}
		__syncthreads();
	}
	if (hasPeer)
	{
		for (int k = tid; k < recvCount; k += stride) field[recvIds[k]] = __ldcv(&recvBox[k]);
	}
	grid.sync();
}

#ifndef B2CU_SOLVER_THREADS
#define B2CU_SOLVER_THREADS 256
#endif
#ifndef B2CU_VEL_BLOCKS
#define B2CU_VEL_BLOCKS 3
#endif
#ifndef B2CU_POS_BLOCKS
#define B2CU_POS_BLOCKS 3
#endif

// first parallel op at or after (pass, op) in the flattened (pass, op) sequence; returns false at the end
__device__ __forceinline__ bool NextParallelOp(const SolverPlan& plan, int passEnd, int& pass, int& op)
{
	while (pass <= passEnd)
	{
		for (; op < plan.opCount; ++op)
			if (plan.opType[op] == OP_PARALLEL) return true;
		op = 0;
		++pass;
	}
	return false;
}

// velocity half: warm start, velocity iterations, impulse store, position integration.
// Software-pipelined across the colour barriers: the rows of a thread's first constraint of the NEXT colour are
// loaded before the grid barrier (they do not depend on other threads; only the body velocities do), so the
// DRAM latency of every phase hides behind the barrier.
// JOINTS: the world has joints (a separate instance, so that the joint code costs the contact-only instance nothing).
template <bool JOINTS>
__global__ void __launch_bounds__(B2CU_SOLVER_THREADS, JOINTS ? 2 : B2CU_VEL_BLOCKS) SolverVelocityPersistentKernel(DeviceArrays d,
                                                                                                                 SolverPlan plan)
{
	GridDependencyWait();
	GridSync grid = {plan.softBarrier, gridDim.x, 0u, plan.shard.stuck};
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	unsigned seq = plan.shard.seq;
	// joints are initialised in pass 0, after the contacts' warm start, whether or not warm starting is on
	const int firstPass = (plan.warmStarting || (JOINTS && plan.jointOpCount > 0)) ? 0 : 1;
	const int lastPass = plan.velocityIterations;

	VelPre pre;
	bool havePre = false;
	{
		int np = firstPass, no = 0;
		if (NextParallelOp(plan, lastPass, np, no) && tid < plan.opSize[no])
		{
			pre = LoadVelPre(d, plan.opStart[no] + tid, plan.opStart[no] + tid >= d.colourTwoStart[plan.opColour[no]]);
			havePre = true;
		}
	}

	// pass 0 = warm start, passes 1..vIters = velocity iterations
	for (int pass = firstPass; pass <= lastPass; ++pass)
	{
		if (JOINTS && pass >= 1) JointRunOps(grid, d, plan, JOINT_VELOCITY, 0, tid, stride);
		const int opCount = (JOINTS && pass == 0 && !plan.warmStarting) ? 0 : plan.opCount;
		for (int op = 0; op < opCount; ++op)
		{
			const int type = plan.opType[op], begin = plan.opStart[op], n = plan.opSize[op];
			if (type == OP_PARALLEL)
			{
				// rows at or behind twoStart have a second point (ColourKeysKernel's order): loaded with the rest of the row
				const int twoStart = d.colourTwoStart[plan.opColour[op]];
				if (tid < n)
				{
					if (!havePre) pre = LoadVelPre(d, begin + tid, begin + tid >= twoStart);
					if (pass == 0) WarmStartPre(d, begin + tid, pre);
					else SolveVelocityPre(d, begin + tid, pre);
				}
				havePre = false;
				for (int t = tid + stride; t < n; t += stride)
				{
					const VelPre more = LoadVelPre(d, begin + t, begin + t >= twoStart);
					if (pass == 0) WarmStartPre(d, begin + t, more);
					else SolveVelocityPre(d, begin + t, more);
				}
				// prefetch for the next colour
				int np = pass, no = op + 1;
				if (NextParallelOp(plan, lastPass, np, no) && tid < plan.opSize[no])
				{
					// the impulses of that constraint were last written in an earlier phase, except when the next
					// colour is this very one (a single colour): then they are loaded after the barrier
					if (!(plan.opStart[no] == begin))
					{
						pre = LoadVelPre(d, plan.opStart[no] + tid, plan.opStart[no] + tid >= d.colourTwoStart[plan.opColour[no]]);
						havePre = true;
					}
				}
				grid.sync();
			}
			else if (type == OP_SERIAL)
			{
				if (tid == 0)
				{
					for (int t = 0; t < n; ++t)
					{
						if (pass == 0) WarmStartOne(d, begin + t);
						else SolveVelocityOne(d, begin + t);
					}
				}
				grid.sync();
			}
			else
			{
				HaloExchange(grid, plan.shard, d.vel, type == OP_PUSH_DOWN, seq++, tid, stride);
			}
		}
		if (JOINTS && pass == 0) JointRunOps(grid, d, plan, JOINT_INIT, 0, tid, stride);
	}

	// b2ContactSolver::StoreImpulses, then b2Island::Solve position integration (independent of each other)
	for (int k = tid; k < plan.constraintCount; k += stride) StoreImpulseOne(d, k);
	for (int b = tid; b < plan.bodyCount; b += stride) IntegratePositionOne(d, b, plan.h);
}

// position half: position iterations with the per-island early exit; same software pipelining
template <bool JOINTS>
__global__ void __launch_bounds__(B2CU_SOLVER_THREADS, JOINTS ? 2 : B2CU_POS_BLOCKS) SolverPositionPersistentKernel(DeviceArrays d,
                                                                                                                 SolverPlan plan)
{
	GridDependencyWait();
	GridSync grid = {plan.softBarrier, gridDim.x, 0u, plan.shard.stuck};
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	unsigned seq = plan.shard.seq;
	const int lastPass = plan.positionIterations - 1;

	PosPre pre;
	bool havePre = false;
	{
		int np = 0, no = 0;
		if (NextParallelOp(plan, lastPass, np, no) && tid < plan.opSize[no])
		{
			pre = LoadPosPre(d, plan.opStart[no] + tid);
			havePre = true;
		}
	}

	for (int it = 0; it <= lastPass; ++it)
	{
		for (int op = 0; op < plan.opCount; ++op)
		{
			const int type = plan.opType[op], begin = plan.opStart[op], n = plan.opSize[op];
			if (type == OP_PARALLEL)
			{
				for (int t = tid; t < n; t += stride)
				{
					if (!(t == tid && havePre)) pre = LoadPosPre(d, begin + t);
					int root = __float_as_int(pre.rad.w);
					if (IslandDone(d, it, root, plan.bodyCount)) continue;
					float minSep = SolvePositionPre(d, pre);
					AtomicMinByRoot(d.islandMinSep + (size_t)it * plan.bodyCount, root, FloatToOrdered(minSep));
				}
				havePre = false;
				int np = it, no = op + 1;
				if (NextParallelOp(plan, lastPass, np, no) && tid < plan.opSize[no])
				{
					// position rows are constants of the step: always safe to load ahead
					pre = LoadPosPre(d, plan.opStart[no] + tid);
					havePre = true;
				}
				grid.sync();
			}
			else if (type == OP_SERIAL)
			{
				if (tid == 0)
				{
					for (int t = 0; t < n; ++t)
					{
						int k = begin + t;
						int root = __float_as_int(d.sRadius[k].w);
						if (IslandDone(d, it, root, plan.bodyCount)) continue;
						float minSep = SolvePositionOne(d, k);
						atomicMin(&d.islandMinSep[(size_t)it * plan.bodyCount + root], FloatToOrdered(minSep));
					}
				}
				grid.sync();
			}
			else
			{
				HaloExchange(grid, plan.shard, d.pos, type == OP_PUSH_DOWN, seq++, tid, stride);
			}
		}
		if (JOINTS) JointRunOps(grid, d, plan, JOINT_POSITION, it, tid, stride);
	}
}

// ---- step-start halo sync: the owner's full body state (5 float4 rows per body) to the ghost copies ----------
#define B2CU_GHOST_ROWS 5
__global__ void GhostSendKernel(DeviceArrays d, ShardState sh)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(k, sh.exportCount)
	{
		int b = sh.exportIds[k];
		// the sync rows live behind the per-iteration exchange rows of the mailbox: the first push of the solver may
		// land before the receiver has read the sync rows
		float4* out = sh.lowerFromUpper + (size_t)sh.exportCount + (size_t)k * B2CU_GHOST_ROWS;
		out[0] = d.xf[b];
		out[1] = d.pos[b];
		out[2] = d.pos0[b];
		out[3] = d.vel[b];
		out[4] = d.force[b];
	}
	__threadfence_system();
}
__global__ void ShardSignalKernel(unsigned* flag, unsigned seq)
{
	GridDependencyWait();
	*reinterpret_cast<volatile unsigned*>(flag) = seq;
	__threadfence_system();
}
__global__ void ShardWaitKernel(const unsigned* flag, unsigned seq, int* stuck)
{
	GridDependencyWait();
	long long spins = 0;
	while (*reinterpret_cast<const volatile unsigned*>(flag) < seq)
	{
		if (++spins > (1ll << 25))
		{
			*stuck = 2;
			break;
		}
	}
	__threadfence_system();
}
__global__ void GhostApplyKernel(DeviceArrays d, ShardState sh)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(k, sh.ghostCount)
	{
		int b = sh.ghostIds[k];
		const float4* in = sh.fromUpper + (size_t)sh.ghostCount + (size_t)k * B2CU_GHOST_ROWS;
		d.xf[b] = __ldcv(&in[0]);
		d.pos[b] = __ldcv(&in[1]);
		d.pos0[b] = __ldcv(&in[2]);
		d.vel[b] = __ldcv(&in[3]);
		d.force[b] = __ldcv(&in[4]);
	}
}

// ---------------------------------------------------------------------------------------------------------
// b2Island::Solve, body part 3 (b2Island.cpp:338-395): copy back, SynchronizeTransform, sleep bookkeeping.
// ---------------------------------------------------------------------------------------------------------
__global__ void FinalizeBodiesKernel(DeviceArrays d, int bodyCount, float h, int allowSleep)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		uint32_t bf = d.bflags[b];
		if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) continue;
		float4 p = d.pos[b];
		float4 ms = d.mass[b];
		Rot q = SinCos(p.z);
		Vec2 o = V(p.x, p.y) - Mul(q, V(ms.z, ms.w));
		d.xf[b] = make_float4(o.x, o.y, q.s, q.c);

		if (allowSleep)
		{
			float4 v = d.vel[b];
			float4 f = d.force[b];
			const float linTolSqr = B2CU_LINEAR_SLEEP_TOLERANCE * B2CU_LINEAR_SLEEP_TOLERANCE;
			const float angTolSqr = B2CU_ANGULAR_SLEEP_TOLERANCE * B2CU_ANGULAR_SLEEP_TOLERANCE;
			float sleepTime;
			if ((bf & B2CU_BODY_AUTOSLEEP) == 0 || v.z * v.z > angTolSqr || Dot(V(v.x, v.y), V(v.x, v.y)) > linTolSqr)
			{
				sleepTime = 0.0f;
			}
			else
			{
				sleepTime = f.w + h;
			}
			f.w = sleepTime;
			d.force[b] = f;
			AtomicMinByRoot(d.islandMinSleep, d.island[b], __float_as_int(sleepTime));
		}
	}
}

__global__ void SleepIslandsKernel(DeviceArrays d, int bodyCount, int positionIterations)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		uint32_t bf = d.bflags[b];
		if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) continue;
		int root = d.island[b];
		float minSleepTime = __int_as_float(d.islandMinSleep[root]);
		bool positionSolved = positionIterations > 0 && IslandDone(d, positionIterations, root, bodyCount);
		if (minSleepTime >= B2CU_TIME_TO_SLEEP && positionSolved)
		{
			// b2Body::SetAwake(false), Box2D/Dynamics/b2Body.h:702-711
			d.bflags[b] = bf & ~B2CU_BODY_AWAKE;
			float4 v = d.vel[b];
			d.vel[b] = make_float4(0.0f, 0.0f, 0.0f, v.w);
			d.force[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// SynchronizeFixtures + MoveProxy: b2ContactManager::SynchronizeFixtures (b2ContactManager.cpp:315-364),
// shape ComputeAABB (b2PolygonShape.cpp:340-357, b2CircleShape.cpp:83-90, b2EdgeShape.cpp:116-129) and the
// fat-AABB rule of b2DynamicTree::MoveProxy (Box2D/Collision/b2DynamicTree.cpp:130-174).  One thread per proxy.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ComputeAABB(const b2cuShape* __restrict__ s, const Xf& xf)
{
	int type = s->type;
	float r = s->radius;
	if (type == B2CU_SHAPE_CIRCLE)
	{
		Vec2 p = xf.p + Mul(xf.q, ShapeV(s, 0));
		return make_float4(p.x - r, p.y - r, p.x + r, p.y + r);
	}
	if (type == B2CU_SHAPE_EDGE)
	{
		Vec2 v1 = Mul(xf, ShapeV(s, 0));
		Vec2 v2 = Mul(xf, ShapeV(s, 1));
		Vec2 lower = V(Min(v1.x, v2.x), Min(v1.y, v2.y));
		Vec2 upper = V(Max(v1.x, v2.x), Max(v1.y, v2.y));
		if (s->flags & B2CU_EDGE_CHAIN_CHILD) r = 0.0f; // b2ChainShape::ComputeAABB adds no margin
		return make_float4(lower.x - r, lower.y - r, upper.x + r, upper.y + r);
	}
	Vec2 lower = Mul(xf, ShapeV(s, 0));
	Vec2 upper = lower;
	int count = s->count;
	for (int i = 1; i < count; ++i)
	{
		Vec2 v = Mul(xf, ShapeV(s, i));
		lower = V(Min(lower.x, v.x), Min(lower.y, v.y));
		upper = V(Max(upper.x, v.x), Max(upper.y, v.y));
	}
	return make_float4(lower.x - r, lower.y - r, upper.x + r, upper.y + r);
}

__global__ void __launch_bounds__(256) SyncProxiesKernel(DeviceArrays d, int proxyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(p, proxyCount)
	{
		int b = d.pbody[p];
		uint32_t bf = d.bflags[b];
		if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) continue;

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
			// the box the proxy had when the step began, for the shortcut of QueryProxy
			d.fatPrev[p] = fat;
			uint32_t g = d.pgroup[p];
			if (!(g & ((uint32_t)B2CU_PROXY_MOVED << 16))) g |= (uint32_t)B2CU_PROXY_MOVED_SYNC << 16;
			d.pgroup[p] = g | ((uint32_t)B2CU_PROXY_MOVED << 16);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// Broad-phase pair finding.  Replaces b2BroadPhase::UpdatePairs + b2DynamicTree::Query (Box2D/Collision/
// b2BroadPhase.h:211-267, b2DynamicTree.h:168-201) and b2ContactManager::AddPair (b2ContactManager.cpp:237-312).
//
// A hierarchical hashed grid over the fat AABBs is rebuilt every step.  Level k has square cells of size
// S_k = S_0 * 2^k; a proxy lives on the finest level whose cell is larger than its fat AABB and is registered
// once, in the cell of its lower corner, in one hash table shared by all levels.  A pair is always examined
// from its finer-level member (or, on equal levels, from the moved one / the smaller id), which only has to
// look at <= 3x3 cells per level.  Every overlap of fat AABBs involving a moved proxy is found exactly, so the
// pair SET equals the dynamic tree's, whatever the tree shape would have been.
// ---------------------------------------------------------------------------------------------------------
#define B2CU_GRID_LEVELS 24

struct GridParams
{
	float cell0;     // S_0
	float invCell0;  // 1 / S_0 (S_k and 1/S_k are exact power-of-two multiples)
	uint32_t mask;   // hash table size - 1
};

__device__ __forceinline__ int CellCoord(float x, float invCell) { return (int)floorf(x * invCell); }
__device__ __forceinline__ uint32_t CellHash(int level, int ix, int iy, uint32_t mask)
{
	return (((uint32_t)ix * 0x9E3779B1u) ^ ((uint32_t)iy * 0x85EBCA77u) ^ ((uint32_t)level * 0xC2B2AE3Du)) & mask;
}
// finest level whose cell holds the AABB with a 1/64 margin (the margin absorbs the rounding of the cell
// coordinates, so that an overlapping proxy is never more than one cell below the query's lower cell)
__device__ __forceinline__ int ProxyLevel(float4 fat, float cell0)
{
	float ext = Max(fat.z - fat.x, fat.w - fat.y);
	float lim = cell0 * 0.984375f;
	int k = 0;
	while (!(ext <= lim) && k < B2CU_GRID_LEVELS)
	{
		lim *= 2.0f;
		++k;
	}
	return k; // == B2CU_GRID_LEVELS: larger than the coarsest cell ("huge")
}
__device__ __forceinline__ float LevelInvCell(float invCell0, int level) { return ldexpf(invCell0, -level); }

__device__ __forceinline__ bool IsMovedProxy(const DeviceArrays& d, int p)
{
	return ((d.pgroup[p] >> 16) & B2CU_PROXY_MOVED) != 0;
}

// Broad-phase invariant (b2ContactManager::AddPair + b2ContactManager::Collide, b2ContactManager.cpp:117-262): when a
// step begins, every pair of overlapping fat boxes that passes the filters has a contact, and Collide only destroys
// the contacts whose fat boxes do not overlap.  So a pair whose boxes ALREADY overlapped at the beginning of the step
// needs no lookup at all.  The box at the beginning of the step is fatPrev for a proxy SyncProxiesKernel moved
// (MOVED_SYNC), fat for one that did not move; a proxy the caller touched (MOVED without MOVED_SYNC: new fixture,
// SetTransform, Refilter, SetType) has no such history and always takes the full check.
__device__ __forceinline__ bool StartBox(const DeviceArrays& d, int p, uint32_t proxyFlags, float4 fat, float4* box)
{
	if (proxyFlags & B2CU_PROXY_MOVED_SYNC)
	{
		*box = d.fatPrev[p];
		return true;
	}
	*box = fat;
	return !(proxyFlags & B2CU_PROXY_MOVED);
}

// the counters and level statistics of one pair search, zeroed by one launch instead of five memset nodes
__global__ void BroadphaseResetKernel(DeviceArrays d)
{
	GridDependencyWait();
	const int t = threadIdx.x;
	if (t < 64) d.levelInfo[t] = 0;
	if (t == 64) d.counters[CNT_MOVED] = 0;
	if (t == 65) d.counters[CNT_SCRATCH] = 0;
	if (t == 66) d.counters[CNT_LARGE] = 0;
	if (t == 67) d.counters[CNT_NEW_PAIRS] = 0;
}

// levelInfo[l] = proxies on level l, levelInfo[LEVELS+1+l] = moved proxies on level l (l == LEVELS: huge)
__global__ void __launch_bounds__(256) GridCountKernel(DeviceArrays d, int proxyCount, GridParams g)
{
	GridDependencyWait();
	__shared__ int sh[2 * (B2CU_GRID_LEVELS + 1)];
	if (threadIdx.x < 2 * (B2CU_GRID_LEVELS + 1)) sh[threadIdx.x] = 0;
	__syncthreads();
	B2CU_GRID_STRIDE(p, proxyCount)
	{
		float4 fat = d.fat[p];
		if ((d.pgroup[p] >> 16) & B2CU_PROXY_INACTIVE)
		{
			// proxy of an inactive body: in no cell and never queried (GridFillKernel skips negative cells)
			d.cellOfProxy[p] = -2;
			continue;
		}
		bool moved = IsMovedProxy(d, p);
		int level = ProxyLevel(fat, g.cell0);
		atomicAdd(&sh[level], 1);
		if (moved)
		{
			atomicAdd(&sh[B2CU_GRID_LEVELS + 1 + level], 1);
			d.movedList[atomicAdd(&d.counters[CNT_SCRATCH], 1)] = p;
		}
		if (level == B2CU_GRID_LEVELS)
		{
			d.cellOfProxy[p] = -1;
			d.largeList[atomicAdd(&d.counters[CNT_LARGE], 1)] = p;
		}
		else
		{
			float inv = LevelInvCell(g.invCell0, level);
			uint32_t h = CellHash(level, CellCoord(fat.x, inv), CellCoord(fat.y, inv), g.mask);
			d.cellOfProxy[p] = (int)h;
			atomicAdd(&d.cellCount[h], 1);
		}
	}
	__syncthreads();
	if (threadIdx.x < 2 * (B2CU_GRID_LEVELS + 1) && sh[threadIdx.x])
	{
		atomicAdd(&d.levelInfo[threadIdx.x], sh[threadIdx.x]);
		if (threadIdx.x > B2CU_GRID_LEVELS) atomicAdd(&d.counters[CNT_MOVED], sh[threadIdx.x]);
	}
}

__global__ void GridFillKernel(DeviceArrays d, int proxyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(p, proxyCount)
	{
		int h = d.cellOfProxy[p];
		if (h < 0) continue;
		// the counts of GridCountKernel are used up as cursors (a cell is filled from its end; the order inside a cell
		// is immaterial, the pair search produces a set), so nothing has to be zeroed between the two kernels and the
		// array is clean for the next step's count
		int slot = d.cellStart[h] + atomicSub(&d.cellCount[h], 1) - 1;
		// the entry carries the fat box: a query reads its candidates as one contiguous run instead of one gather each
		d.cellItems[slot] = p;
		d.cellBoxes[slot] = d.fat[p];
	}
}

// b2ContactManager::AddPair: same body, existing contact, b2Body::ShouldCollide, default b2ContactFilter,
// and the contact-class table (no edge-edge contact, b2Contact.cpp:44-50)
__device__ __forceinline__ void TryAddPair(const DeviceArrays& d, int q, int p, int2 counts /* (all slots, main region) */,
                                           int pairCapacity)
{
	int a = q < p ? q : p, b = q < p ? p : q;
	int bodyA = d.pbody[a], bodyB = d.pbody[b];
	if (bodyA == bodyB) return;
	uint64_t key = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
	// existing contact?  the keys whose low id is `a` are the contiguous range [lowStart[a], lowStart[a+1])
	{
		int lo = d.lowStart[a], hi = d.lowStart[a + 1];
		while (lo < hi)
		{
			int mid = (lo + hi) >> 1;
			if (d.c.key[mid] < key) lo = mid + 1;
			else hi = mid;
		}
		if (lo < d.lowStart[a + 1] && d.c.key[lo] == key && !(d.c.flags[lo] & B2CU_CONTACT_DEAD)) return;
		// recent contacts sit in a small sorted tail behind the main region until the next compaction
		const int nTail = counts.x - counts.y;
		if (nTail > 0)
		{
			int j = counts.y + LowerBound64(d.c.key + counts.y, nTail, key);
			if (j < counts.x && d.c.key[j] == key && !(d.c.flags[j] & B2CU_CONTACT_DEAD)) return;
		}
	}
	// b2Body::ShouldCollide: at least one dynamic body; in a sharded world it must be one this shard owns
	if (!IsOwnedDynamic(d.bflags[bodyA]) && !IsOwnedDynamic(d.bflags[bodyB])) return;
	if (JointBlocks(d, bodyA, bodyB)) return;
	if (!d.customFilter && !DefaultFilter(d.pfilter[a], d.pgroup[a], d.pfilter[b], d.pgroup[b])) return;
	if (d.shapes[d.pshape[a]].type == B2CU_SHAPE_EDGE && d.shapes[d.pshape[b]].type == B2CU_SHAPE_EDGE) return;
	int slot = atomicAdd(&d.counters[CNT_NEW_PAIRS], 1);
	if (slot < pairCapacity) d.newKeys[slot] = key;
	else d.counters[CNT_ERROR] = 1;
}

// Pair search of one proxy p:
//   its own level       if p moved:  every overlapping r, except a moved r with a smaller id (r reports that pair)
//   each coarser level  every overlapping r such that p or r moved (r never looks down at p's level)
//   the huge list       like a coarser level
// so each pair with a moved member is examined exactly once.
// The cells a proxy must examine are dealt round-robin to the `lanes` threads that share the proxy (lane = 0..lanes-1):
// one proxy is a chain of several hundred dependent loads, far too long for one thread when only the moved proxies
// (a tenth of the world) are in flight.
__device__ __forceinline__ void QueryProxy(const DeviceArrays& d, int p, bool movedP, const GridParams& g,
                                           const int* shCount, const int* shMoved, int2 contactCount, int pairCapacity,
                                           int lane, int lanes)
{
	float4 fp = d.fat[p];
	int levelP = ProxyLevel(fp, g.cell0);
	const int nHuge = shCount[B2CU_GRID_LEVELS];
	int turn = 0;
	float4 startP;
	const bool knownP = StartBox(d, p, d.pgroup[p] >> 16, fp, &startP);

	for (int level = levelP; level < B2CU_GRID_LEVELS; ++level)
	{
		if (shCount[level] == 0) continue;
		bool same = level == levelP;
		if (!movedP && (same || shMoved[level] == 0)) continue;
		float inv = LevelInvCell(g.invCell0, level);
		int x0 = CellCoord(fp.x, inv) - 1, x1 = CellCoord(fp.z, inv);
		int y0 = CellCoord(fp.y, inv) - 1, y1 = CellCoord(fp.w, inv);
		for (int cy = y0; cy <= y1; ++cy)
		{
			for (int cx = x0; cx <= x1; ++cx)
			{
				if (turn++ % lanes != lane) continue;
				uint32_t h = CellHash(level, cx, cy, g.mask);
				int start = d.cellStart[h];
				int end = d.cellStart[h + 1];
				for (int s = start; s < end; ++s)
				{
					int r = d.cellItems[s];
					if (r == p) continue;
					float4 fr = d.cellBoxes[s];
					// a bucket can hold several cells and levels: take r only when it is registered in THIS cell
					if (CellCoord(fr.x, inv) != cx || CellCoord(fr.y, inv) != cy) continue;
					if (ProxyLevel(fr, g.cell0) != level) continue;
					if (!AabbOverlap(fp, fr)) continue;
					uint32_t flagsR = d.pgroup[r] >> 16;
					bool movedR = (flagsR & B2CU_PROXY_MOVED) != 0;
					if (same)
					{
						if (movedR && r < p) continue;
					}
					else if (!movedP && !movedR)
					{
						continue;
					}
					float4 startR;
					if (knownP && StartBox(d, r, flagsR, fr, &startR) && AabbOverlap(startP, startR) &&
					    !(d.jointFreedCount != 0 && JointFreed(d, d.pbody[p], d.pbody[r])))
						continue;
					TryAddPair(d, p, r, contactCount, pairCapacity);
				}
			}
		}
	}

	if (nHuge > 0 && (movedP || shMoved[B2CU_GRID_LEVELS] > 0))
	{
		bool hugeP = levelP == B2CU_GRID_LEVELS;
		for (int s = lane; s < nHuge; s += lanes)
		{
			int r = d.largeList[s];
			if (r == p) continue;
			float4 fr = d.fat[r];
			if (!AabbOverlap(fp, fr)) continue;
			uint32_t flagsR = d.pgroup[r] >> 16;
			bool movedR = (flagsR & B2CU_PROXY_MOVED) != 0;
			if (hugeP)
			{
				if (!movedP || (movedR && r < p)) continue;
			}
			else if (!movedP && !movedR)
			{
				continue;
			}
			float4 startR;
			if (knownP && StartBox(d, r, flagsR, fr, &startR) && AabbOverlap(startP, startR) &&
					    !(d.jointFreedCount != 0 && JointFreed(d, d.pbody[p], d.pbody[r])))
						continue;
			TryAddPair(d, p, r, contactCount, pairCapacity);
		}
	}
}

// dense pass over the compacted list of moved proxies (all lanes of a warp have work)
__global__ void __launch_bounds__(128) QueryMovedKernel(DeviceArrays d, GridParams g, int2 contactCount, int pairCapacity)
{
	GridDependencyWait();
	__shared__ int shCount[B2CU_GRID_LEVELS + 1];
	__shared__ int shMoved[B2CU_GRID_LEVELS + 1];
	if (threadIdx.x <= B2CU_GRID_LEVELS)
	{
		shCount[threadIdx.x] = d.levelInfo[threadIdx.x];
		shMoved[threadIdx.x] = d.levelInfo[B2CU_GRID_LEVELS + 1 + threadIdx.x];
	}
	__syncthreads();
	// threads per moved proxy: as many as the grid can give (a power of two up to a warp).  One proxy is a chain of
	// a few hundred dependent loads; with few moved proxies the kernel time is that chain, so it is split.
	const int moved = d.counters[CNT_SCRATCH];
	const int threads = gridDim.x * blockDim.x;
	int lanes = 1;
	// (a thread per proxy as long as the moved proxies alone fill a quarter of the grid: with 130k movers of a 1M-body
	// pile two lanes per proxy measured 3 % slower than one, four lanes 20 %)
	while (lanes < 32 && moved * lanes * 4 <= threads) lanes *= 2;
	const int n = moved * lanes;
	B2CU_GRID_STRIDE(t, n)
	{
		QueryProxy(d, d.movedList[t / lanes], true, g, shCount, shMoved, contactCount, pairCapacity, t % lanes, lanes);
	}
}

// proxies that did not move only have to look UP, at coarser levels that contain a moved proxy (a moved wall, a
// bullet): usually there is none and the kernel returns at once
__global__ void __launch_bounds__(128) QueryUnmovedKernel(DeviceArrays d, int proxyCount, GridParams g, int2 contactCount,
                                                          int pairCapacity)
{
	GridDependencyWait();
	__shared__ int shCount[B2CU_GRID_LEVELS + 1];
	__shared__ int shMoved[B2CU_GRID_LEVELS + 1];
	__shared__ int lowestMoved;
	if (threadIdx.x <= B2CU_GRID_LEVELS)
	{
		shCount[threadIdx.x] = d.levelInfo[threadIdx.x];
		shMoved[threadIdx.x] = d.levelInfo[B2CU_GRID_LEVELS + 1 + threadIdx.x];
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		int lm = B2CU_GRID_LEVELS + 1;
		for (int l = B2CU_GRID_LEVELS; l >= 0; --l)
			if (shMoved[l] > 0) lm = l;
		lowestMoved = lm;
	}
	__syncthreads();
	// a moved proxy on level L is only interesting to unmoved proxies on levels below L
	if (lowestMoved == 0 || lowestMoved > B2CU_GRID_LEVELS)
	{
		bool anyAbove = false;
		for (int l = 1; l <= B2CU_GRID_LEVELS; ++l) anyAbove |= shMoved[l] > 0;
		if (!anyAbove) return;
	}
	B2CU_GRID_STRIDE(p, proxyCount)
	{
		if ((d.pgroup[p] >> 16) & (B2CU_PROXY_MOVED | B2CU_PROXY_INACTIVE)) continue;
		QueryProxy(d, p, false, g, shCount, shMoved, contactCount, pairCapacity, 0, 1);
	}
}

// lowStart[p] = index of the first contact whose low proxy id is >= p, for p in [0, proxyCount]
__global__ void BuildLowStartKernel(DeviceArrays d, int contactCount, int proxyCount)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(p, proxyCount + 1) { d.lowStart[p] = LowerBound64(d.c.key, contactCount, (uint64_t)(uint32_t)p << 32); }
}

__global__ void IotaKernel(int* out, int n)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, n) { out[i] = i; }
}

// over the move buffer only (movedList holds every flagged proxy, GridCountKernel)
__global__ void ClearMovedKernel(DeviceArrays d)
{
	GridDependencyWait();
	const int n = d.counters[CNT_SCRATCH];
	B2CU_GRID_STRIDE(t, n)
	{
		int p = d.movedList[t];
		d.pgroup[p] &= ~((uint32_t)(B2CU_PROXY_MOVED | B2CU_PROXY_MOVED_SYNC | B2CU_PROXY_NEW) << 16);
	}
}

// ---------------------------------------------------------------------------------------------------------
// Contact set rebuild: drops contacts destroyed by Collide and merges the new (sorted) pairs, keeping the set
// in key order.  Replaces FinishFindNewContacts / OnContactCreate / b2Contact::Create / Destroy bookkeeping
// (b2ContactManager.cpp:120-172, 366-386, 507-564; b2Contact.cpp:72-157).
// ---------------------------------------------------------------------------------------------------------
// Merge step of the lazily compacted contact set: the live contacts of c[srcBegin, srcBegin+srcCount) go to
// cAlt[dstBegin + (rank among the live contacts of that range) + (number of live keys of the OTHER sorted run that
// are smaller)].  otherRank == nullptr: the other run has no dead entries.
__global__ void MergeMoveKernel(DeviceArrays d, int srcBegin, int srcCount, const int* __restrict__ srcRank,
                                const uint64_t* __restrict__ otherKeys, int otherCount, const int* __restrict__ otherRank,
                                int otherLive, int dstBegin)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(t, srcCount)
	{
		int i = srcBegin + t;
		if (d.c.flags[i] & B2CU_CONTACT_DEAD) continue;
		uint64_t key = d.c.key[i];
		int lb = LowerBound64(otherKeys, otherCount, key);
		int less = otherRank ? (lb < otherCount ? otherRank[lb] : otherLive) : lb;
		int dest = dstBegin + srcRank[t] + less;
		d.cAlt.key[dest] = key;
		d.cAlt.proxies[dest] = d.c.proxies[i];
		d.cAlt.flags[dest] = d.c.flags[i];
		d.cAlt.m0[dest] = d.c.m0[i];
		d.cAlt.m1[dest] = d.c.m1[i];
		d.cAlt.m2[dest] = d.c.m2[i];
		d.cAlt.m3[dest] = d.c.m3[i];
		d.cAlt.mix[dest] = d.c.mix[i];
		d.cAlt.toiCount[dest] = d.c.toiCount[i];
		d.cAlt.colour[dest] = d.c.colour[i];
		d.cAlt.stamp[dest] = d.c.stamp[i];
	}
}

// new contacts (sorted keys) are created in cAlt, merged with the live contacts of the tail region
// c[tailBegin, tailBegin+tailCount) (rank tailRank, tailLive of them)
__global__ void RebuildNewKernel(DeviceArrays d, int tailBegin, int tailCount, const int* __restrict__ tailRank,
                                 int tailLive, int newCount, int dstBegin, uint32_t stamp)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(j, newCount)
	{
		uint64_t key = d.newKeys[j];
		int lb = LowerBound64(d.c.key + tailBegin, tailCount, key);
		int keptBefore = lb < tailCount ? tailRank[lb] : tailLive;
		int dest = dstBegin + j + keptBefore;

		int a = (int)(key >> 32), b = (int)(key & 0xFFFFFFFFull);
		int typeA = d.shapes[d.pshape[a]].type, typeB = d.shapes[d.pshape[b]].type;
		int pA = a, pB = b;
		if (NeedsSwap(typeA, typeB))
		{
			pA = b;
			pB = a;
		}
		int bodyA = d.pbody[pA], bodyB = d.pbody[pB];
		uint32_t fbA = d.bflags[bodyA], fbB = d.bflags[bodyB];
		uint32_t flA = d.pgroup[pA] >> 16, flB = d.pgroup[pB] >> 16;
		bool sensor = ((flA | flB) & B2CU_PROXY_SENSOR) != 0;

		uint32_t flags = B2CU_CONTACT_ENABLED | (sensor ? B2CU_CONTACT_SENSOR : 0u);
		// b2Contact::IsToiCandidate, b2Contact.cpp:300-324
		if (!sensor)
		{
			if ((fbA | fbB) & B2CU_BODY_BULLET)
			{
				flags |= B2CU_CONTACT_TOI_CANDIDATE;
			}
			else
			{
				bool includesNonDynamic = !IsDynamic(fbA) || !IsDynamic(fbB);
				bool neitherThick = ((flA | flB) & B2CU_PROXY_THICK) == 0;
				if (includesNonDynamic && neitherThick) flags |= B2CU_CONTACT_TOI_CANDIDATE;
			}
			// OnContactCreate wakes both bodies (b2ContactManager.cpp:524-529)
			d.wake[bodyA] = 1;
			d.wake[bodyB] = 1;
		}

		float2 matA = d.pmat[pA], matB = d.pmat[pB];
		float friction = sqrtf(matA.x * matB.x);                  // b2MixFriction, b2Contact.h:40-43
		float restitution = matA.y > matB.y ? matA.y : matB.y;      // b2MixRestitution, b2Contact.h:47-50

		d.cAlt.key[dest] = key;
		d.cAlt.proxies[dest] = make_int4(pA, pB, bodyA, bodyB);
		d.cAlt.flags[dest] = flags;
		d.cAlt.m0[dest] = make_float4(0.f, 0.f, 0.f, 0.f);
		d.cAlt.m1[dest] = make_float4(0.f, 0.f, 0.f, 0.f);
		d.cAlt.m2[dest] = make_float4(0.f, 0.f, 0.f, 0.f);
		d.cAlt.m3[dest] = make_uint4(0u, 0u, 0u, 0u);
		d.cAlt.mix[dest] = make_float4(friction, restitution, 0.0f, 1.0f);
		d.cAlt.toiCount[dest] = 0;
		d.cAlt.colour[dest] = B2CU_COLOUR_NONE;
		d.cAlt.stamp[dest] = stamp;
	}
}

// TOI eligibility (b2World.cpp:317-341 filters + b2Contact::IsMinToiCandidate, b2Contact.h:404-419)
__global__ void ToiFlagsKernel(DeviceArrays d, int contactCount, int* flagsOut)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, contactCount)
	{
		uint32_t f = d.c.flags[i];
		int ok = 0;
		if (!(f & B2CU_CONTACT_DEAD) && (f & B2CU_CONTACT_TOI_CANDIDATE) && (f & B2CU_CONTACT_ENABLED) && d.c.toiCount[i] <= B2CU_MAX_SUB_STEPS)
		{
			int4 pr = d.c.proxies[i];
			uint32_t fA = d.bflags[pr.z], fB = d.bflags[pr.w];
			if (IsAwakeNonStatic(fA) || IsAwakeNonStatic(fB)) ok = 1;
		}
		flagsOut[i] = ok;
	}
}

// b2Body::m_sweep of a body, as the step left it: c0 / a0 = where the solver picked the body up, c / a = where it put it
__device__ __forceinline__ Sweep LoadSweep(const DeviceArrays& d, int body)
{
	float4 p0 = d.pos0[body], p = d.pos[body], m = d.mass[body];
	Sweep s;
	s.localCenter = V(m.z, m.w);
	s.c0 = V(p0.x, p0.y);
	s.a0 = p0.z;
	s.alpha0 = p0.w;
	s.c = V(p.x, p.y);
	s.a = p.z;
	return s;
}

// First pass of b2World::SolveTOI over the eligible contacts (b2FindMinToiContactTask, b2World.cpp:298-351, and
// b2World::ComputeToi :404-447): time of impact of every candidate from its bodies' sweeps.  After a complete step
// every sweep starts at alpha0 = 0, so no sweep has to be advanced and the contacts are independent.
__global__ void ToiFirstPassKernel(DeviceArrays d, const int* __restrict__ list, const int* __restrict__ countPtr,
                                   float* __restrict__ alphaOut)
{
	GridDependencyWait();
	const int count = *countPtr;
	B2CU_GRID_STRIDE(k, count)
	{
		int i = list[k];
		int4 pr = d.c.proxies[i];
		Sweep sA = LoadSweep(d, pr.z), sB = LoadSweep(d, pr.w);
		float alpha0 = sA.alpha0;
		float t;
		int state = TimeOfImpact(&t, MakeGjkProxy(d.shapes + d.pshape[pr.x]), sA, MakeGjkProxy(d.shapes + d.pshape[pr.y]),
		                         sB, 1.0f);
		float alpha = state == TOI_TOUCHING ? Min(alpha0 + (1.0f - alpha0) * t, 1.0f) : 1.0f;
		alphaOut[k] = alpha;
		// alpha is in [0, 1]: its bit pattern orders like the value
		atomicMin(reinterpret_cast<unsigned int*>(d.counters + CNT_TOI_MIN_ALPHA), __float_as_uint(alpha));
	}
}

// ... and the winner among equal alphas: the smallest contact key (b2Contact::ToiLessThan, b2Contact.cpp:326-334)
__global__ void ToiMinKeyKernel(DeviceArrays d, const int* __restrict__ list, const int* __restrict__ countPtr,
                                const float* __restrict__ alpha)
{
	GridDependencyWait();
	const int count = *countPtr;
	const unsigned int best = *reinterpret_cast<const unsigned int*>(d.counters + CNT_TOI_MIN_ALPHA);
	B2CU_GRID_STRIDE(k, count)
	{
		if (__float_as_uint(alpha[k]) == best)
			atomicMin(reinterpret_cast<unsigned long long*>(d.counters + CNT_TOI_MIN_KEY), (unsigned long long)d.c.key[list[k]]);
	}
}

// is b2Contact::IsToiCandidate (b2Contact.cpp:300-324) satisfiable by any pair: a bullet body, or a non-dynamic
// body carrying a fixture that is not thick-shape
__global__ void ToiPossibleKernel(DeviceArrays d, int bodyCount, int proxyCount)
{
	GridDependencyWait();
	bool any = false;
	B2CU_GRID_STRIDE(i, bodyCount > proxyCount ? bodyCount : proxyCount)
	{
		if (i < bodyCount && (d.bflags[i] & B2CU_BODY_BULLET)) any = true;
		if (i < proxyCount)
		{
			uint32_t pf = d.pgroup[i] >> 16;
			if (!(pf & (B2CU_PROXY_THICK | B2CU_PROXY_SENSOR)) && !IsDynamic(d.bflags[d.pbody[i]])) any = true;
		}
	}
	if (__any_sync(__activemask(), any) && any) d.counters[CNT_STICKY_TOI] = 1;
}

// end of step: clear island flags (b2World::ClearPostSolve, b2World.cpp:1433-1465), clear forces
// (b2World::ClearForces :1506-1523), count awake bodies
__global__ void EndStepBodiesKernel(DeviceArrays d, int bodyCount, int clearForces)
{
	GridDependencyWait();
	int local = 0;
	B2CU_GRID_STRIDE(b, bodyCount)
	{
		uint32_t bf = d.bflags[b];
		if (bf & B2CU_BODY_ISLAND) d.bflags[b] = bf & ~B2CU_BODY_ISLAND;
		if (clearForces && !IsStatic(bf))
		{
			float4 f = d.force[b];
			f.x = 0.0f;
			f.y = 0.0f;
			f.z = 0.0f;
			d.force[b] = f;
		}
		if (IsAwakeNonStatic(bf)) ++local;
	}
	for (int dlt = 16; dlt > 0; dlt >>= 1) local += __shfl_down_sync(0xffffffffu, local, dlt);
	if ((threadIdx.x & 31) == 0 && local) atomicAdd(&d.counters[CNT_AWAKE_BODIES], local);
}

// ---- world queries: one pass over the fat boxes (16 bytes per proxy: a million proxies are 16 MB, a few microseconds
// of HBM time), then a stable compaction -- no tree to walk ----
__global__ void QueryAabbSelectKernel(DeviceArrays d, int proxyCount, float4 box, int* __restrict__ flags)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(p, proxyCount)
	{
		flags[p] = (AabbOverlap(d.fat[p], box) && !((d.pgroup[p] >> 16) & B2CU_PROXY_INACTIVE)) ? 1 : 0;
	}
}

// b2DynamicTree::RayCast's node test (b2DynamicTree.h:203-287) applied to every proxy box: the box of the segment must
// overlap, and the segment's line must not separate the box (|dot(v, p1 - c)| - dot(|v|, h) <= 0 with v normal to the ray)
__global__ void RayCastSelectKernel(DeviceArrays d, int proxyCount, float2 p1, float2 p2, int* __restrict__ flags)
{
	GridDependencyWait();
	Vec2 a = V(p1.x, p1.y), b = V(p2.x, p2.y);
	Vec2 r = Normalized(b - a);
	Vec2 v = CrossSV(1.0f, r);
	Vec2 absV = V(fabsf(v.x), fabsf(v.y));
	float4 seg = make_float4(Min(a.x, b.x), Min(a.y, b.y), Max(a.x, b.x), Max(a.y, b.y));
	B2CU_GRID_STRIDE(p, proxyCount)
	{
		float4 f = d.fat[p];
		int hit = 0;
		if (AabbOverlap(f, seg) && !((d.pgroup[p] >> 16) & B2CU_PROXY_INACTIVE))
		{
			Vec2 c = V(0.5f * (f.x + f.z), 0.5f * (f.y + f.w));
			Vec2 h = V(0.5f * (f.z - f.x), 0.5f * (f.w - f.y));
			float separation = fabsf(Dot(v, a - c)) - Dot(absV, h);
			hit = separation > 0.0f ? 0 : 1;
		}
		flags[p] = hit;
	}
}

// stand-alone batched b2Distance (b2cuDistancePairs)
__global__ void DistancePairsKernel(const b2cuShape* __restrict__ shapes, int pairCount, const int* __restrict__ shapeA,
                                    const float4* __restrict__ xfA, const int* __restrict__ shapeB,
                                    const float4* __restrict__ xfB, int useRadii, b2cuDistanceResult* __restrict__ out)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, pairCount)
	{
		GjkCache cache;
		cache.count = 0;
		cache.metric = 0.0f;
		GjkOutput g;
		GjkDistance(&g, &cache, MakeGjkProxy(shapes + shapeA[i]), MakeXf(xfA[i]), MakeGjkProxy(shapes + shapeB[i]),
		            MakeXf(xfB[i]), useRadii != 0);
		b2cuDistanceResult o;
		o.distance = g.distance;
		o.pointA[0] = g.pointA.x;
		o.pointA[1] = g.pointA.y;
		o.pointB[0] = g.pointB.x;
		o.pointB[1] = g.pointB.y;
		o.iterations = g.iterations;
		out[i] = o;
	}
}

// stand-alone batched b2TimeOfImpact (b2cuTimeOfImpactPairs)
__device__ __forceinline__ Sweep MakeSweep(const b2cuSweep& r)
{
	Sweep s;
	s.localCenter = V(r.localCenter[0], r.localCenter[1]);
	s.c0 = V(r.c0[0], r.c0[1]);
	s.c = V(r.c[0], r.c[1]);
	s.a0 = r.a0;
	s.a = r.a;
	s.alpha0 = r.alpha0;
	return s;
}

__global__ void TimeOfImpactPairsKernel(const b2cuShape* __restrict__ shapes, int pairCount, const int* __restrict__ shapeA,
                                        const b2cuSweep* __restrict__ sweepA, const int* __restrict__ shapeB,
                                        const b2cuSweep* __restrict__ sweepB, const float* __restrict__ tMax,
                                        b2cuToiResult* __restrict__ out)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, pairCount)
	{
		float t;
		int state = TimeOfImpact(&t, MakeGjkProxy(shapes + shapeA[i]), MakeSweep(sweepA[i]),
		                         MakeGjkProxy(shapes + shapeB[i]), MakeSweep(sweepB[i]), tMax[i]);
		b2cuToiResult o;
		o.state = state;
		o.t = t;
		out[i] = o;
	}
}

// stand-alone batched narrow phase (b2cuCollidePairs)
__global__ void CollidePairsKernel(const b2cuShape* __restrict__ shapes, int pairCount, const int* __restrict__ shapeA,
                                   const float4* __restrict__ xfA, const int* __restrict__ shapeB,
                                   const float4* __restrict__ xfB, b2cuManifold* __restrict__ out)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, pairCount)
	{
		Manifold m;
		m.localNormal = V(0.f, 0.f);
		m.localPoint = V(0.f, 0.f);
		m.lp[0] = m.lp[1] = V(0.f, 0.f);
		m.ni[0] = m.ni[1] = m.ti[0] = m.ti[1] = 0.0f;
		m.id[0] = m.id[1] = 0u;
		m.type = 0;
		m.pointCount = 0;
		Evaluate(&m, shapes + shapeA[i], MakeXf(xfA[i]), shapes + shapeB[i], MakeXf(xfB[i]));
		b2cuManifold o;
		o.localNormal[0] = m.localNormal.x;
		o.localNormal[1] = m.localNormal.y;
		o.localPoint[0] = m.localPoint.x;
		o.localPoint[1] = m.localPoint.y;
		for (int k = 0; k < 2; ++k)
		{
			o.points[k].localPoint[0] = m.lp[k].x;
			o.points[k].localPoint[1] = m.lp[k].y;
			o.points[k].normalImpulse = 0.0f;
			o.points[k].tangentImpulse = 0.0f;
			o.id[k] = m.id[k];
		}
		o.type = m.type;
		o.pointCount = m.pointCount;
		out[i] = o;
	}
}

// b2cuGetContactsByKey: one thread per requested key, AoS records out
__global__ void GatherContactsByKeyKernel(DeviceArrays d, int contactCount, int mainCount,
                                          const uint64_t* __restrict__ keys, int n, b2cuContact* __restrict__ out)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(j, n)
	{
		uint64_t key = keys[j];
		int i = LowerBound64(d.c.key, mainCount, key);
		bool found = i < mainCount && d.c.key[i] == key && !(d.c.flags[i] & B2CU_CONTACT_DEAD);
		if (!found && contactCount > mainCount)
		{
			i = mainCount + LowerBound64(d.c.key + mainCount, contactCount - mainCount, key);
			found = i < contactCount && d.c.key[i] == key && !(d.c.flags[i] & B2CU_CONTACT_DEAD);
		}
		b2cuContact o;
		if (found)
		{
			int4 pr = d.c.proxies[i];
			float4 m0 = d.c.m0[i], m1 = d.c.m1[i], m2 = d.c.m2[i], mix = d.c.mix[i];
			uint4 m3 = d.c.m3[i];
			uint32_t fA = d.bflags[pr.z], fB = d.bflags[pr.w];
			uint32_t f = d.c.flags[i] & 0xFFu & ~(uint32_t)B2CU_CONTACT_INACTIVE;
			if (!IsAwakeNonStatic(fA) && !IsAwakeNonStatic(fB)) f |= B2CU_CONTACT_INACTIVE;
			o.proxyA = pr.x;
			o.proxyB = pr.y;
			o.flags = f;
			o.friction = mix.x;
			o.restitution = mix.y;
			o.tangentSpeed = mix.z;
			o.toiCount = d.c.toiCount[i];
			o.toi = mix.w;
			o.manifold.localNormal[0] = m0.x; o.manifold.localNormal[1] = m0.y;
			o.manifold.localPoint[0] = m0.z; o.manifold.localPoint[1] = m0.w;
			o.manifold.points[0].localPoint[0] = m1.x; o.manifold.points[0].localPoint[1] = m1.y;
			o.manifold.points[0].normalImpulse = m1.z; o.manifold.points[0].tangentImpulse = m1.w;
			o.manifold.points[1].localPoint[0] = m2.x; o.manifold.points[1].localPoint[1] = m2.y;
			o.manifold.points[1].normalImpulse = m2.z; o.manifold.points[1].tangentImpulse = m2.w;
			o.manifold.id[0] = m3.x; o.manifold.id[1] = m3.y;
			o.manifold.type = (int32_t)m3.z;
			o.manifold.pointCount = (int32_t)m3.w;
			o.stamp = d.c.stamp[i];
			o.reserved = 0u;
		}
		else
		{
			o.stamp = 0u;
			o.reserved = 0u;
			int a = (int)(key >> 32), b = (int)(key & 0xFFFFFFFFull);
			bool swap = NeedsSwap(d.shapes[d.pshape[a]].type, d.shapes[d.pshape[b]].type);
			o.proxyA = swap ? b : a;
			o.proxyB = swap ? a : b;
			o.flags = 0u;
			float2 matA = d.pmat[a], matB = d.pmat[b];
			o.friction = sqrtf(matA.x * matB.x);
			o.restitution = matA.y > matB.y ? matA.y : matB.y;
			o.tangentSpeed = 0.0f;
			o.toiCount = 0;
			o.toi = 1.0f;
			o.manifold.localNormal[0] = o.manifold.localNormal[1] = 0.0f;
			o.manifold.localPoint[0] = o.manifold.localPoint[1] = 0.0f;
			for (int k = 0; k < 2; ++k)
			{
				o.manifold.points[k].localPoint[0] = o.manifold.points[k].localPoint[1] = 0.0f;
				o.manifold.points[k].normalImpulse = o.manifold.points[k].tangentImpulse = 0.0f;
				o.manifold.id[k] = 0u;
			}
			o.manifold.type = 0;
			o.manifold.pointCount = 0;
		}
		out[j] = o;
	}
}

__global__ void SinCosKernel(int n, const float* __restrict__ x, float* __restrict__ s, float* __restrict__ c)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(i, n)
	{
		Rot q = SinCos(x[i]);
		s[i] = q.s;
		c[i] = q.c;
	}
}

} // namespace b2cu

#include "b2cu_solver_flow.cuh"
#include "b2cu_toi_step.cuh"
