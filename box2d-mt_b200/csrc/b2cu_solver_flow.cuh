// b2cu_solver_flow.cuh -- the contact solver of a step as a DATAFLOW over the bodies instead of colour phases behind
// grid barriers.  Included by b2cu_kernels.cuh.
//
// The coloured Gauss-Seidel fixes, for every dynamic body, the order in which its constraints touch it: colour by
// colour, pass after pass (a body has at most one constraint per colour).  That order is all that the result depends
// on.  The barrier kernels (SolverVelocityPersistentKernel / SolverPositionPersistentKernel) enforce it with one grid
// barrier per colour and pass -- 99 + 33 barriers for a pile in 11 colours, each costing a few microseconds of idle
// machine (the measured floor was 0.39 ms of the 1.13 ms velocity kernel).  Here every body row carries a VERSION in
// its spare fourth component: the number of constraint updates applied to it so far.  A constraint of colour c in pass
// p knows which version of each of its two bodies it must see,
//        expected = p * deg(body) + rank_c(body),   deg = popcount(colourMask), rank_c = popcount(colourMask below bit c),
// waits for exactly those (a 16-byte single-copy-atomic load of state + version: LDG.E.128.STRONG.GPU), solves, and
// publishes the new state with version + 1 in one 16-byte store.  No barrier, no fence: data and flag are one word.
// Threads take the constraints in the same (pass, colour, index) order as before, so whoever holds the globally
// earliest unsolved constraint can always proceed: no deadlock as long as the grid is co-resident (cooperative launch).
// Waiting is done in a warp-uniform polling loop (a lane that is ready solves inside the loop): lanes of one warp may
// depend on each other across a colour boundary, so no lane may sit at a reconvergence point while others spin.
//
// Same arithmetic (SolveVelocityCore / WarmStartCore / SolvePositionCore), same order per body: bit-identical results.
// Used for worlds without joints, unsharded, with no serial overflow list; everything else keeps the barrier kernels.
#pragma once

namespace b2cu
{

__device__ __forceinline__ float4 LoadRow128(const float4* p)
{
	float4 v;
	asm volatile("{\n\t.reg .b128 r;\n\tld.relaxed.gpu.global.b128 r, [%4];\n\tmov.b128 {%0,%1,%2,%3}, r;\n\t}"
	             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
	             : "l"(p)
	             : "memory");
	return v;
}
__device__ __forceinline__ void StoreRow128(float4* p, float4 v)
{
	asm volatile("{\n\t.reg .b128 r;\n\tmov.b128 r, {%1,%2,%3,%4};\n\tst.relaxed.gpu.global.b128 [%0], r;\n\t}" ::"l"(p), "f"(v.x),
	             "f"(v.y), "f"(v.z), "f"(v.w)
	             : "memory");
}

// b2Island::Solve position integration (b2Island.cpp:283-313) of body b from the velocity row just read; the rows are
// written back with version 0 for the position iterations
__device__ __forceinline__ void IntegratePositionRow(const DeviceArrays& d, int b, float h, float4 v4)
{
	float4 p = d.pos[b];
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
	d.pos[b] = make_float4(c.x, c.y, a, 0.0f);
	d.vel[b] = make_float4(v.x, v.y, w, 0.0f);
}

__device__ __forceinline__ void PrefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// row index this thread handles in the round after (op, base), or -1: the next slice of the same colour, else the first
// slice of the next colour (of the next pass after the last colour)
__device__ __forceinline__ int FlowNextRow(const SolverPlan& plan, int op, int base, int tid, int stride, bool lastPass)
{
	if (base + stride < plan.opSize[op])
	{
		const int t = base + stride + tid;
		return t < plan.opSize[op] ? plan.opStart[op] + t : -1;
	}
	int next = op + 1;
	if (next == plan.opCount)
	{
		if (lastPass) return -1;
		next = 0;
	}
	return tid < plan.opSize[next] ? plan.opStart[next] + tid : -1;
}

// polls before a wait is declared stuck (seconds of wall time): the step then fails loudly instead of hanging the GPU
#define B2CU_FLOW_SPIN_LIMIT (1 << 22)
// true when this wait must be given up: it has polled too long, or some other wait already has (checked now and then)
__device__ __forceinline__ bool FlowStuck(const DeviceArrays& d, int spins)
{
	if (spins > B2CU_FLOW_SPIN_LIMIT)
	{
		d.counters[CNT_FLOW_STUCK] = 1;
		return true;
	}
	if ((spins & 4095) == 0 && *reinterpret_cast<volatile int*>(&d.counters[CNT_FLOW_STUCK]) != 0) return true;
	return false;
}

// version a constraint of colour `colour` must find on a body in its `passIndex`-th pass
__device__ __forceinline__ int FlowExpected(uint32_t mask, int colour, int passIndex)
{
	return passIndex * __popc(mask) + __popc(mask & ((1u << colour) - 1u));
}

// warm start + velocity iterations + impulse store + position integration
#ifndef B2CU_FLOW_VEL_BLOCKS
#define B2CU_FLOW_VEL_BLOCKS 4
#endif
__global__ void __launch_bounds__(B2CU_SOLVER_THREADS, B2CU_FLOW_VEL_BLOCKS) SolverVelocityFlowKernel(DeviceArrays d, SolverPlan plan)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	const int firstPass = plan.warmStarting ? 0 : 1;
	const int lastPass = plan.velocityIterations;
	int passIndex = 0;
	for (int pass = firstPass; pass <= lastPass; ++pass, ++passIndex)
	{
		for (int op = 0; op < plan.opCount; ++op)
		{
			const int begin = plan.opStart[op], n = plan.opSize[op], colour = plan.opColour[op];
			const int twoStart = d.colourTwoStart[colour];
			for (int base = 0; base < n; base += stride) // warp-uniform trip count
			{
				const int t = base + tid;
				bool done = t >= n;
				VelPre pre;
				int expA = 0, expB = 0;
				bool dynA = false, dynB = false;
				if (!done)
				{
					pre = LoadVelPre(d, begin + t, begin + t >= twoStart);
					dynA = pre.ms.x != 0.0f || pre.ms.y != 0.0f;
					dynB = pre.ms.z != 0.0f || pre.ms.w != 0.0f;
					if (dynA) expA = FlowExpected(d.colourMask[pre.sb.x], colour, passIndex);
					if (dynB) expB = FlowExpected(d.colourMask[pre.sb.y], colour, passIndex);
				}
				if (plan.flowPrefetch)
				{
					// the rows of this thread's next constraint on their way into L2 while this one is solved
					const int kn = FlowNextRow(plan, op, base, tid, stride, pass == lastPass);
					if (kn >= 0)
					{
						PrefetchL2(&d.sBody[kn]);
						PrefetchL2(&d.sMass[kn]);
						PrefetchL2(&d.sNormal[kn]);
						PrefetchL2(&d.sImp[kn]);
						PrefetchL2(&d.sP0a[kn]);
						PrefetchL2(&d.sP0b[kn]);
					}
				}
				int spins = 0;
				for (;;)
				{
					if (!done)
					{
						float4 vA = LoadRow128(&d.vel[pre.sb.x]);
						float4 vB = LoadRow128(&d.vel[pre.sb.y]);
						bool ready = (!dynA || __float_as_int(vA.w) == expA) && (!dynB || __float_as_int(vB.w) == expB);
						if (!ready && FlowStuck(d, ++spins)) ready = true;
						if (ready)
						{
							if (pass == 0) WarmStartCore(d, begin + t, pre, vA, vB);
							else SolveVelocityCore(d, begin + t, pre, vA, vB);
							if (dynA)
							{
								vA.w = __int_as_float(expA + 1);
								StoreRow128(&d.vel[pre.sb.x], vA);
							}
							if (dynB)
							{
								vB.w = __int_as_float(expB + 1);
								StoreRow128(&d.vel[pre.sb.y], vB);
							}
							done = true;
						}
					}
					if (__all_sync(0xffffffffu, done)) break;
				}
			}
		}
	}
	// b2ContactSolver::StoreImpulses: every thread for the rows it solved (nobody else knows that they are final)
	for (int op = 0; op < plan.opCount && !plan.debugSkipStore; ++op)
	{
		const int begin = plan.opStart[op], n = plan.opSize[op];
		for (int t = tid; t < n; t += stride) StoreImpulseOne(d, begin + t);
	}
	// b2Island::Solve position integration: a body is ready when all its updates have arrived
	for (int base = 0; base < plan.bodyCount; base += stride)
	{
		const int b = base + tid;
		bool done = b >= plan.bodyCount;
		int expected = 0;
		if (!done)
		{
			const uint32_t bf = d.bflags[b];
			if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) done = true;
			else expected = IsDynamic(bf) ? passIndex * __popc(d.colourMask[b]) : 0;
		}
		int spins = 0;
		for (;;)
		{
			if (!done)
			{
				float4 v = LoadRow128(&d.vel[b]);
				bool ready = !IsDynamic(d.bflags[b]) || __float_as_int(v.w) == expected;
				if (!ready && FlowStuck(d, ++spins)) ready = true;
				if (ready)
				{
					IntegratePositionRow(d, b, plan.h, v);
					done = true;
				}
			}
			if (__all_sync(0xffffffffu, done)) break;
		}
	}
}

// position iterations.  An island stops iterating once the smallest separation of its previous iteration is within
// tolerance (b2Island.cpp:318-335): that is a property of the whole island, so iterations stay separated by a grid
// barrier (3 per step); inside an iteration the colours flow.
__global__ void __launch_bounds__(B2CU_SOLVER_THREADS, B2CU_POS_BLOCKS) SolverPositionFlowKernel(DeviceArrays d, SolverPlan plan)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	for (int it = 0; it < plan.positionIterations; ++it)
	{
		for (int op = 0; op < plan.opCount; ++op)
		{
			const int begin = plan.opStart[op], n = plan.opSize[op], colour = plan.opColour[op];
			for (int base = 0; base < n; base += stride)
			{
				const int t = base + tid;
				bool done = t >= n;
				PosPre pre;
				int expA = 0, expB = 0, root = 0;
				bool dynA = false, dynB = false;
				if (!done)
				{
					pre = LoadPosPre(d, begin + t);
					root = __float_as_int(pre.rad.w);
					// an island that is done was done in every later iteration too: its bodies' versions stand still and
					// nobody waits for them (all constraints of a dynamic body belong to its island)
					if (IslandDone(d, it, root, plan.bodyCount)) done = true;
				}
				if (!done)
				{
					dynA = pre.ms.x != 0.0f || pre.ms.y != 0.0f;
					dynB = pre.ms.z != 0.0f || pre.ms.w != 0.0f;
					if (dynA) expA = FlowExpected(d.colourMask[pre.sb.x], colour, it);
					if (dynB) expB = FlowExpected(d.colourMask[pre.sb.y], colour, it);
				}
				if (plan.flowPrefetch)
				{
					const int kn = FlowNextRow(plan, op, base, tid, stride, it + 1 == plan.positionIterations);
					if (kn >= 0)
					{
						PrefetchL2(&d.sBody[kn]);
						PrefetchL2(&d.sMass[kn]);
						PrefetchL2(&d.sLocal[kn]);
						PrefetchL2(&d.sLocalP[kn]);
						PrefetchL2(&d.sCenters[kn]);
						PrefetchL2(&d.sRadius[kn]);
					}
				}
				int spins = 0;
				float minSep = 0.0f;
				bool solved = false;
				for (;;)
				{
					if (!done)
					{
						float4 pA = LoadRow128(&d.pos[pre.sb.x]);
						float4 pB = LoadRow128(&d.pos[pre.sb.y]);
						bool ready = (!dynA || __float_as_int(pA.w) == expA) && (!dynB || __float_as_int(pB.w) == expB);
						if (!ready && FlowStuck(d, ++spins)) ready = true;
						if (ready)
						{
							minSep = SolvePositionCore(pre, pA, pB);
							if (dynA)
							{
								pA.w = __int_as_float(expA + 1);
								StoreRow128(&d.pos[pre.sb.x], pA);
							}
							if (dynB)
							{
								pB.w = __int_as_float(expB + 1);
								StoreRow128(&d.pos[pre.sb.y], pB);
							}
							done = true;
							solved = true;
						}
					}
					if (__all_sync(0xffffffffu, done)) break;
				}
				if (solved) AtomicMinByRoot(d.islandMinSep + (size_t)it * plan.bodyCount, root, FloatToOrdered(minSep));
			}
		}
		if (it + 1 < plan.positionIterations) grid.sync();
	}
}

} // namespace b2cu
