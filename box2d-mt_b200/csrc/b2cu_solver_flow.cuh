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
__device__ __forceinline__ void IntegratePositionRow(const DeviceArrays& d, int b, float h, float4 v4, int base)
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
	d.pos[b] = make_float4(c.x, c.y, a, __int_as_float(base));
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

// ---- sharded worlds: the halo bodies' rows travel between the shards inside the same dataflow -----------------------
// A halo body (a GHOST here = an EXPORT of the upper neighbour) is updated first by its owner's constraints (colours
// 0-15 there), then by the lower shard's cross constraints (colours 16-31 there); its colour mask is the union of the
// two sides' masks (exchanged once per step, HaloMaskKernels), so both sides number its updates alike.  Whoever applies
// the LAST update of its side in a pass also stores the row -- state and version, one 16-byte system-scope store over
// NVLink -- into the neighbour's mailbox; a constraint that waits for a halo body accepts the expected version from the
// local row or from the mailbox, whichever shows it.  No flags, no fences, no barriers: the transfer of a boundary
// body overlaps everything that does not depend on it.
#define B2CU_HALO_EXPORT 0x40000000
__device__ __forceinline__ float4 LoadRowSys(const float4* p)
{
	float4 v;
	asm volatile("{\n\t.reg .b128 r;\n\tld.relaxed.sys.global.b128 r, [%4];\n\tmov.b128 {%0,%1,%2,%3}, r;\n\t}"
	             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
	             : "l"(p)
	             : "memory");
	return v;
}
__device__ __forceinline__ void StoreRowSys(float4* p, float4 v)
{
	asm volatile("{\n\t.reg .b128 r;\n\tmov.b128 r, {%1,%2,%3,%4};\n\tst.relaxed.sys.global.b128 [%0], r;\n\t}" ::"l"(p), "f"(v.x),
	             "f"(v.y), "f"(v.z), "f"(v.w)
	             : "memory");
}
// mailbox rows of the dataflow: [6n, 7n) velocity rows, [7n, 8n) position rows of a box that holds n halo bodies
__device__ __forceinline__ const float4* HaloIn(const ShardState& sh, int slot, int region)
{
	const int k = slot & ~B2CU_HALO_EXPORT;
	return (slot & B2CU_HALO_EXPORT) ? sh.fromLower + (size_t)(6 + region) * sh.exportCount + k
	                                 : sh.fromUpper + (size_t)(6 + region) * sh.ghostCount + k;
}
__device__ __forceinline__ float4* HaloOut(const ShardState& sh, int slot, int region)
{
	const int k = slot & ~B2CU_HALO_EXPORT;
	return (slot & B2CU_HALO_EXPORT) ? sh.lowerFromUpper + (size_t)(6 + region) * sh.exportCount + k
	                                 : sh.upperFromLower + (size_t)(6 + region) * sh.ghostCount + k;
}
// is the update of colour `colour` the last one this shard applies to the halo body in a pass?
__device__ __forceinline__ bool HaloLastOfSide(uint32_t mask, int colour, int slot)
{
	const uint32_t side = (slot & B2CU_HALO_EXPORT) ? (mask & 0xFFFFu) : mask; // an export's own colours are the low half
	return (side >> (colour + 1)) == 0u;
}
// the row of body b at version `expected`, from the local array or (halo bodies) from the neighbour's last push
__device__ __forceinline__ bool FlowAcquire(const ShardState& sh, const float4* rows, int b, int slot, int region, int expected,
                                            float4* out)
{
	float4 v = LoadRow128(&rows[b]);
	if (__float_as_int(v.w) == expected)
	{
		*out = v;
		return true;
	}
	if (slot >= 0)
	{
		const float4* in = HaloIn(sh, slot, region);
		if (in != nullptr)
		{
			v = LoadRowSys(in);
			if (__float_as_int(v.w) == expected)
			{
				*out = v;
				return true;
			}
		}
	}
	*out = v;
	return false;
}

// the neighbour's last push of a halo body, if it is the expected version
__device__ __forceinline__ bool FlowAcquireMail(const ShardState& sh, int slot, int region, int expected, float4* out)
{
	const float4 v = LoadRowSys(HaloIn(sh, slot, region));
	if (__float_as_int(v.w) != expected) return false;
	*out = v;
	return true;
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
__device__ __forceinline__ int FlowExpected(uint32_t mask, int colour, int passIndex, int extraDeg = 0)
{
	return passIndex * (__popc(mask) + extraDeg) + __popc(mask & ((1u << colour) - 1u));
}

// Constraints that found no free colour (a body with more contacts than colours: the bullet box of Add Pair ploughing
// through a thousand circles) come after the colours in every pass, in list order.  For the dataflow they are simply
// further updates of their bodies: overflow row j is update number popcount(mask) + (overflow rows before j on that
// body) of the pass, and a body's degree grows by its overflow rows.  FlowOverflowPrepKernel counts both.
// keys[2j + side] = body << 32 | (2j + side) for the dynamic bodies of overflow row j (others: ~0, sorted to the end)
__global__ void FlowOverflowKeysKernel(DeviceArrays d, int ovStart, int ovCount, uint64_t* __restrict__ keys)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(j, ovCount)
	{
		const int4 sb = d.sBody[ovStart + j];
		const float4 ms = d.sMass[ovStart + j];
		const bool dynA = ms.x != 0.0f || ms.y != 0.0f, dynB = ms.z != 0.0f || ms.w != 0.0f;
		keys[2 * j] = dynA ? (((uint64_t)(uint32_t)sb.x << 32) | (uint32_t)(2 * j)) : ~0ull;
		keys[2 * j + 1] = dynB ? (((uint64_t)(uint32_t)sb.y << 32) | (uint32_t)(2 * j + 1)) : ~0ull;
	}
}
// after the sort: the position of a key inside its body's run is the rank of that row on that body; the run length is
// the body's extra degree
__global__ void FlowOverflowRanksKernel(const uint64_t* __restrict__ keys, int n, int* __restrict__ ovRank, int* __restrict__ ovDeg)
{
	GridDependencyWait();
	B2CU_GRID_STRIDE(p, n)
	{
		const uint64_t key = keys[p];
		if (key == ~0ull) continue;
		const uint64_t body = key >> 32;
		const int first = LowerBound64(keys, n, body << 32);
		ovRank[(int)(uint32_t)(key & 0xFFFFFFFFull)] = p - first;
		if (p == first) ovDeg[(int)body] = LowerBound64(keys, n, (body + 1) << 32) - first;
	}
}

// the serial overflow list of one velocity pass (kept out of line: its registers must not cost the main loop anything)
// warm start + velocity iterations + impulse store + position integration
#ifndef B2CU_FLOW_VEL_BLOCKS
#define B2CU_FLOW_VEL_BLOCKS 4
#endif
template <bool SHARD, bool OVERFLOW>
__global__ void __launch_bounds__(B2CU_SOLVER_THREADS, (SHARD || OVERFLOW) ? 3 : B2CU_FLOW_VEL_BLOCKS) SolverVelocityFlowKernel(const __grid_constant__ DeviceArrays d, const __grid_constant__ SolverPlan plan)
{
	GridDependencyWait();
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	const int firstPass = plan.warmStarting ? 0 : 1;
	const int lastPass = plan.velocityIterations;
	const int base0 = plan.flowBase; // versions of this step start here (stale mailbox rows of earlier steps never match)
	int passIndex = 0;
	for (int pass = firstPass; pass <= lastPass; ++pass, ++passIndex)
	{
		for (int op = 0; op < plan.opCount; ++op)
		{
			// the overflow list (constraints that found no free colour) is one more op: its rows carry explicit ranks,
			// and the rows of one body form a chain through the same versions (sharded plans also carry halo ops: skipped)
			const bool overflowOp = OVERFLOW && plan.opType[op] == OP_SERIAL;
			if (plan.opType[op] != OP_PARALLEL && !overflowOp) continue;
			const int begin = plan.opStart[op], n = plan.opSize[op], colour = overflowOp ? 0 : plan.opColour[op];
			const int twoStart = overflowOp ? 0x7FFFFFFF : d.colourTwoStart[colour];
			for (int base = 0; base < n; base += stride) // warp-uniform trip count
			{
				const int t = base + tid;
				bool done = t >= n;
				VelPre pre;
				int expA = 0, expB = 0, slotA = -1, slotB = -1;
				uint32_t maskA = 0u, maskB = 0u;
				bool dynA = false, dynB = false;
				if (!done)
				{
					pre = overflowOp ? LoadVelPre(d, begin + t) : LoadVelPre(d, begin + t, begin + t >= twoStart);
					dynA = pre.ms.x != 0.0f || pre.ms.y != 0.0f;
					dynB = pre.ms.z != 0.0f || pre.ms.w != 0.0f;
					if (dynA)
					{
						maskA = d.colourMask[pre.sb.x];
						expA = base0 + (overflowOp ? passIndex * (__popc(maskA) + plan.ovDeg[pre.sb.x]) + __popc(maskA) + plan.ovRank[2 * t]
						                           : FlowExpected(maskA, colour, passIndex, OVERFLOW ? plan.ovDeg[pre.sb.x] : 0));
						if (SHARD) slotA = d.haloSlot[pre.sb.x];
					}
					if (dynB)
					{
						maskB = d.colourMask[pre.sb.y];
						expB = base0 + (overflowOp ? passIndex * (__popc(maskB) + plan.ovDeg[pre.sb.y]) + __popc(maskB) + plan.ovRank[2 * t + 1]
						                           : FlowExpected(maskB, colour, passIndex, OVERFLOW ? plan.ovDeg[pre.sb.y] : 0));
						if (SHARD) slotB = d.haloSlot[pre.sb.y];
					}
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
						// both rows in flight before either is looked at
						float4 vA = LoadRow128(&d.vel[pre.sb.x]);
						float4 vB = LoadRow128(&d.vel[pre.sb.y]);
						bool readyA = !dynA || __float_as_int(vA.w) == expA;
						bool readyB = !dynB || __float_as_int(vB.w) == expB;
						if (SHARD)
						{
							if (!readyA && slotA >= 0) readyA = FlowAcquireMail(plan.shard, slotA, 0, expA, &vA);
							if (!readyB && slotB >= 0) readyB = FlowAcquireMail(plan.shard, slotB, 0, expB, &vB);
						}
						bool ready = readyA && readyB;
						if (!ready && FlowStuck(d, ++spins)) ready = true;
						if (ready)
						{
							if (pass == 0) WarmStartCore(d, begin + t, pre, vA, vB);
							else SolveVelocityCore(d, begin + t, pre, vA, vB);
							if (dynA)
							{
								vA.w = __int_as_float(expA + 1);
								StoreRow128(&d.vel[pre.sb.x], vA);
								if (SHARD && slotA >= 0 && HaloLastOfSide(maskA, colour, slotA)) StoreRowSys(HaloOut(plan.shard, slotA, 0), vA);
							}
							if (dynB)
							{
								vB.w = __int_as_float(expB + 1);
								StoreRow128(&d.vel[pre.sb.y], vB);
								if (SHARD && slotB >= 0 && HaloLastOfSide(maskB, colour, slotB)) StoreRowSys(HaloOut(plan.shard, slotB, 0), vB);
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
		if (plan.opType[op] != OP_PARALLEL && !(OVERFLOW && plan.opType[op] == OP_SERIAL)) continue;
		const int begin = plan.opStart[op], n = plan.opSize[op];
		for (int t = tid; t < n; t += stride) StoreImpulseOne(d, begin + t);
	}
	// b2Island::Solve position integration: a body is ready when all its updates have arrived
	for (int base = 0; base < plan.bodyCount; base += stride)
	{
		const int b = base + tid;
		bool done = b >= plan.bodyCount;
		int expected = 0, slot = -1;
		bool dynamic = false;
		if (!done)
		{
			const uint32_t bf = d.bflags[b];
			if (IsStatic(bf) || !(bf & B2CU_BODY_ISLAND)) done = true;
			else
			{
				dynamic = IsDynamic(bf);
				expected = base0 + (dynamic ? passIndex * (__popc(d.colourMask[b]) + (OVERFLOW ? plan.ovDeg[b] : 0)) : 0);
				if (SHARD && dynamic) slot = d.haloSlot[b];
			}
		}
		int spins = 0;
		for (;;)
		{
			if (!done)
			{
				float4 v;
				bool ready = true;
				if (dynamic) ready = FlowAcquire(plan.shard, d.vel, b, slot, 0, expected, &v);
				else v = LoadRow128(&d.vel[b]);
				if (!ready && FlowStuck(d, ++spins)) ready = true;
				if (ready)
				{
					IntegratePositionRow(d, b, plan.h, v, plan.flowBase);
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
template <bool SHARD, bool OVERFLOW>
__global__ void __launch_bounds__(B2CU_SOLVER_THREADS, B2CU_POS_BLOCKS) SolverPositionFlowKernel(const __grid_constant__ DeviceArrays d, const __grid_constant__ SolverPlan plan)
{
	GridDependencyWait();
	GridSync grid = {plan.softBarrier, gridDim.x, 0u, plan.shard.stuck};
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	const int base0 = plan.flowBase;
	for (int it = 0; it < plan.positionIterations; ++it)
	{
		for (int op = 0; op < plan.opCount; ++op)
		{
			const bool overflowOp = OVERFLOW && plan.opType[op] == OP_SERIAL;
			if (plan.opType[op] != OP_PARALLEL && !overflowOp) continue;
			const int begin = plan.opStart[op], n = plan.opSize[op], colour = overflowOp ? 0 : plan.opColour[op];
			for (int base = 0; base < n; base += stride)
			{
				const int t = base + tid;
				bool done = t >= n;
				bool skip = false; // island done: no correction, but the halo bodies' versions still advance (see below)
				PosPre pre;
				int expA = 0, expB = 0, root = 0, slotA = -1, slotB = -1;
				uint32_t maskA = 0u, maskB = 0u;
				bool dynA = false, dynB = false;
				if (!done)
				{
					pre = LoadPosPre(d, begin + t);
					root = __float_as_int(pre.rad.w);
					dynA = pre.ms.x != 0.0f || pre.ms.y != 0.0f;
					dynB = pre.ms.z != 0.0f || pre.ms.w != 0.0f;
					if (SHARD)
					{
						if (dynA) slotA = d.haloSlot[pre.sb.x];
						if (dynB) slotB = d.haloSlot[pre.sb.y];
					}
					// An island that is done was done in every later iteration too: its bodies' versions stand still and
					// nobody of this shard waits for them (all constraints of a dynamic body belong to its island).  The
					// NEIGHBOUR does wait for a halo body (its island is another one and may still iterate): a skipped
					// constraint passes a halo body's row on unchanged, with the version it would have written.
					if (IslandDone(d, it, root, plan.bodyCount))
					{
						skip = true;
						dynA = dynA && slotA >= 0;
						dynB = dynB && slotB >= 0;
						if (!dynA && !dynB) done = true;
					}
				}
				if (!done)
				{
					if (dynA)
					{
						maskA = d.colourMask[pre.sb.x];
						expA = base0 + (overflowOp ? it * (__popc(maskA) + plan.ovDeg[pre.sb.x]) + __popc(maskA) + plan.ovRank[2 * t]
						                           : FlowExpected(maskA, colour, it, OVERFLOW ? plan.ovDeg[pre.sb.x] : 0));
					}
					if (dynB)
					{
						maskB = d.colourMask[pre.sb.y];
						expB = base0 + (overflowOp ? it * (__popc(maskB) + plan.ovDeg[pre.sb.y]) + __popc(maskB) + plan.ovRank[2 * t + 1]
						                           : FlowExpected(maskB, colour, it, OVERFLOW ? plan.ovDeg[pre.sb.y] : 0));
					}
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
						bool readyA = !dynA || __float_as_int(pA.w) == expA;
						bool readyB = !dynB || __float_as_int(pB.w) == expB;
						if (SHARD)
						{
							if (!readyA && slotA >= 0) readyA = FlowAcquireMail(plan.shard, slotA, 1, expA, &pA);
							if (!readyB && slotB >= 0) readyB = FlowAcquireMail(plan.shard, slotB, 1, expB, &pB);
						}
						bool ready = readyA && readyB;
						if (!ready && FlowStuck(d, ++spins)) ready = true;
						if (ready)
						{
							if (!skip)
							{
								minSep = SolvePositionCore(pre, pA, pB);
								solved = true;
							}
							if (dynA)
							{
								pA.w = __int_as_float(expA + 1);
								StoreRow128(&d.pos[pre.sb.x], pA);
								if (SHARD && slotA >= 0 && HaloLastOfSide(maskA, colour, slotA)) StoreRowSys(HaloOut(plan.shard, slotA, 1), pA);
							}
							if (dynB)
							{
								pB.w = __int_as_float(expB + 1);
								StoreRow128(&d.pos[pre.sb.y], pB);
								if (SHARD && slotB >= 0 && HaloLastOfSide(maskB, colour, slotB)) StoreRowSys(HaloOut(plan.shard, slotB, 1), pB);
							}
							done = true;
						}
					}
					if (__all_sync(0xffffffffu, done)) break;
				}
				if (solved) AtomicMinByRoot(d.islandMinSep + (size_t)it * plan.bodyCount, root, FloatToOrdered(minSep));
			}
		}
		if (it + 1 < plan.positionIterations) grid.sync();
	}
	if (SHARD)
	{
		// the halo bodies end the step with the row of whoever updated them last: the local one or the neighbour's push
		const int nHalo = plan.shard.ghostCount + plan.shard.exportCount;
		for (int base = 0; base < nHalo; base += stride)
		{
			const int k = base + tid;
			bool done = k >= nHalo;
			int b = 0, slot = -1, expected = 0;
			if (!done)
			{
				b = k < plan.shard.ghostCount ? plan.shard.ghostIds[k] : plan.shard.exportIds[k - plan.shard.ghostCount];
				slot = d.haloSlot[b];
				const uint32_t bf = d.bflags[b];
				if (!IsDynamic(bf) || !(bf & B2CU_BODY_ISLAND)) done = true;
				else expected = base0 + plan.positionIterations * __popc(d.colourMask[b]);
			}
			int spins = 0;
			for (;;)
			{
				if (!done)
				{
					float4 p;
					bool ready = FlowAcquire(plan.shard, d.pos, b, slot, 1, expected, &p);
					if (!ready && FlowStuck(d, ++spins)) ready = true;
					if (ready)
					{
						StoreRow128(&d.pos[b], p);
						done = true;
					}
				}
				if (__all_sync(0xffffffffu, done)) break;
			}
		}
	}
}

// Once per step, after the colouring: the halo bodies' colour masks of the two sides are united (own colours 0-15 on
// the owner's side, cross colours 16-31 on the lower shard's side), so that both sides count a body's updates alike.
__global__ void HaloMaskSendKernel(DeviceArrays d, ShardState sh)
{
	GridDependencyWait();
	const int n = sh.ghostCount > sh.exportCount ? sh.ghostCount : sh.exportCount;
	B2CU_GRID_STRIDE(k, n)
	{
		if (k < sh.exportCount && sh.lowerFromUpper != nullptr)
			sh.lowerFromUpper[k] = make_float4(__uint_as_float(d.colourMask[sh.exportIds[k]] & 0xFFFFu), 0.f, 0.f, 0.f);
		if (k < sh.ghostCount && sh.upperFromLower != nullptr)
			sh.upperFromLower[k] = make_float4(__uint_as_float(d.colourMask[sh.ghostIds[k]] & 0xFFFF0000u), 0.f, 0.f, 0.f);
	}
	__threadfence_system();
}
__global__ void HaloMaskApplyKernel(DeviceArrays d, ShardState sh)
{
	GridDependencyWait();
	const int n = sh.ghostCount > sh.exportCount ? sh.ghostCount : sh.exportCount;
	B2CU_GRID_STRIDE(k, n)
	{
		if (k < sh.ghostCount) d.colourMask[sh.ghostIds[k]] |= __float_as_uint(__ldcv(&sh.fromUpper[k]).x) & 0xFFFFu;
		if (k < sh.exportCount) d.colourMask[sh.exportIds[k]] |= __float_as_uint(__ldcv(&sh.fromLower[k]).x) & 0xFFFF0000u;
	}
}

} // namespace b2cu
