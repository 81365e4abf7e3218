// world.cu -- C ABI of libb2cuda.so (include/b2cuda.h) and the step driver.
//
// b2cuStep replaces b2World::Step (Box2D/Dynamics/b2World.cpp:1613-1710) and keeps its phase order:
//   [FindNewContacts if new fixtures] -> Collide -> Solve (islands, solver, SynchronizeFixtures,
//   FindNewContacts) -> TOI candidate compaction -> ClearForces.
#include <cstdio>
#include "b2cu_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <cstring>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include <mutex>
#include <unordered_set>

using namespace b2cu;

namespace
{

const int kBlock = 256;
int g_smCount = 148;

inline int GridFor(int n, int block = kBlock)
{
	int blocks = (n + block - 1) / block;
	int cap = g_smCount * 8;
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	return blocks;
}

// QueryMovedKernel: 128-thread blocks, up to 16 resident blocks per SM; the kernel shares the threads out
inline int QueryGrid(int proxyCount)
{
	(void)proxyCount;
	return g_smCount * 16;
}

int SetError(b2cuWorld* w, int code, const char* fmt, ...)
{
	if (w)
	{
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(w->lastError, sizeof(w->lastError), fmt, ap);
		va_end(ap);
	}
	return code;
}

#define CUDA_TRY(w, expr)                                                                                   \
	do                                                                                                      \
	{                                                                                                       \
		cudaError_t _e = (expr);                                                                            \
		if (_e != cudaSuccess)                                                                              \
			return SetError(w, B2CU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
			                __LINE__);                                                                      \
	} while (0)

// B2CU_TRACE=1: a CUDA event after every kernel launch; b2cuStep prints per-kernel device time (including the gap
// before the kernel) for the step.  Developer aid only.
struct TraceState
{
	bool enabled = false;
	std::vector<cudaEvent_t> events;
	std::vector<const char*> names;
	size_t used = 0;
};
TraceState g_trace;

b2cuWorld* g_traceWorld = nullptr;
void TraceMark(b2cuWorld* w, const char* name);
void PrimTraceHook(const char* name, cudaStream_t) { if (g_traceWorld) TraceMark(g_traceWorld, name); }

void TraceMark(b2cuWorld* w, const char* name)
{
	if (!g_trace.enabled) return;
	if (g_trace.used == g_trace.events.size())
	{
		cudaEvent_t e;
		cudaEventCreate(&e);
		g_trace.events.push_back(e);
		g_trace.names.push_back(name);
	}
	g_trace.names[g_trace.used] = name;
	cudaEventRecord(g_trace.events[g_trace.used++], w->stream);
}

#define LAUNCH(w, kernel, grid, block, ...)                 \
	do                                                      \
	{                                                       \
		LaunchPdl(kernel, dim3(grid), dim3(block), (w)->stream, __VA_ARGS__); \
		++(w)->launches;                                    \
		TraceMark(w, #kernel);                              \
	} while (0)

enum CapKind
{
	CAP_BODY,
	CAP_PROXY,
	CAP_SHAPE,
	CAP_CONTACT,
	CAP_GRID,
	CAP_PROXY1,
	CAP_FIXED
};

struct ArrayDesc
{
	void** ptr;
	size_t elemSize;
	CapKind kind;
	size_t fixedCount;
};

template <typename T>
ArrayDesc Desc(T** p, CapKind kind, size_t fixedCount = 0)
{
	ArrayDesc a;
	a.ptr = reinterpret_cast<void**>(p);
	a.elemSize = sizeof(T);
	a.kind = kind;
	a.fixedCount = fixedCount;
	return a;
}

void ContactSetDescs(ContactSet* c, std::vector<ArrayDesc>& v)
{
	v.push_back(Desc(&c->key, CAP_CONTACT));
	v.push_back(Desc(&c->proxies, CAP_CONTACT));
	v.push_back(Desc(&c->flags, CAP_CONTACT));
	v.push_back(Desc(&c->m0, CAP_CONTACT));
	v.push_back(Desc(&c->m1, CAP_CONTACT));
	v.push_back(Desc(&c->m2, CAP_CONTACT));
	v.push_back(Desc(&c->m3, CAP_CONTACT));
	v.push_back(Desc(&c->mix, CAP_CONTACT));
	v.push_back(Desc(&c->toiCount, CAP_CONTACT));
	v.push_back(Desc(&c->colour, CAP_CONTACT));
	v.push_back(Desc(&c->stamp, CAP_CONTACT));
}

std::vector<ArrayDesc> AllArrays(b2cuWorld* w)
{
	DeviceArrays* d = &w->d;
	std::vector<ArrayDesc> v;
	v.push_back(Desc(&d->xf, CAP_BODY));
	v.push_back(Desc(&d->pos, CAP_BODY));
	v.push_back(Desc(&d->pos0, CAP_BODY));
	v.push_back(Desc(&d->vel, CAP_BODY));
	v.push_back(Desc(&d->mass, CAP_BODY));
	v.push_back(Desc(&d->force, CAP_BODY));
	v.push_back(Desc(&d->damp, CAP_BODY));
	v.push_back(Desc(&d->bflags, CAP_BODY));
	v.push_back(Desc(&d->wake, CAP_BODY));
	v.push_back(Desc(&d->wakePatch, CAP_BODY));
	v.push_back(Desc(&d->island, CAP_BODY));
	v.push_back(Desc(&d->islandAwake, CAP_BODY));
	v.push_back(Desc(&d->islandMinSleep, CAP_BODY));
	v.push_back(Desc(&d->colourMask, CAP_BODY));
	v.push_back(Desc(&d->haloSlot, CAP_BODY));
	v.push_back(Desc(&d->colourClaim, CAP_BODY));
	v.push_back(Desc(&d->shapes, CAP_SHAPE));
	v.push_back(Desc(&d->fat, CAP_PROXY));
	v.push_back(Desc(&d->fatPrev, CAP_PROXY));
	v.push_back(Desc(&d->aabb, CAP_PROXY));
	v.push_back(Desc(&d->pbody, CAP_PROXY));
	v.push_back(Desc(&d->pshape, CAP_PROXY));
	v.push_back(Desc(&d->pfilter, CAP_PROXY));
	v.push_back(Desc(&d->pgroup, CAP_PROXY));
	v.push_back(Desc(&d->pmat, CAP_PROXY));
	v.push_back(Desc(&d->pfixture, CAP_PROXY));
	v.push_back(Desc(&d->pradius, CAP_PROXY));
	v.push_back(Desc(&d->lowStart, CAP_PROXY1));
	ContactSetDescs(&d->c, v);
	ContactSetDescs(&d->cAlt, v);
	v.push_back(Desc(&d->cSelect, CAP_CONTACT));
	v.push_back(Desc(&d->listA, CAP_CONTACT));
	v.push_back(Desc(&d->listB, CAP_CONTACT));
	v.push_back(Desc(&d->beginKeys, CAP_CONTACT));
	v.push_back(Desc(&d->endKeys, CAP_CONTACT));
	v.push_back(Desc(&d->destroyEndKeys, CAP_CONTACT));
	v.push_back(Desc(&d->newKeys, CAP_CONTACT));
	v.push_back(Desc(&d->orderKeys, CAP_CONTACT));
	v.push_back(Desc(&d->solverKeys, CAP_CONTACT));
	v.push_back(Desc(&d->listC, CAP_CONTACT));
	v.push_back(Desc(&d->toiKeys, CAP_CONTACT));
	v.push_back(Desc(&d->toiListKeys, CAP_CONTACT));
	v.push_back(Desc(&d->toiListSorted, CAP_CONTACT));
	v.push_back(Desc(&d->toiEventKeys, CAP_CONTACT));
	v.push_back(Desc(&d->toiEventKinds, CAP_CONTACT));
	v.push_back(Desc(&d->toiScratch, CAP_FIXED, B2CU_TOI_SCRATCH_INTS));
	v.push_back(Desc(&d->movedList, CAP_PROXY));
	v.push_back(Desc(&d->largeList, CAP_PROXY));
	v.push_back(Desc(&d->levelInfo, CAP_FIXED, 64));
	v.push_back(Desc(&d->colourCount, CAP_FIXED, B2CU_MAX_COLOURS + 2));
	v.push_back(Desc(&d->colourTwoStart, CAP_FIXED, B2CU_MAX_COLOURS + 2));
	v.push_back(Desc(&d->cellCount, CAP_GRID));
	v.push_back(Desc(&d->cellStart, CAP_GRID));
	v.push_back(Desc(&d->cellItems, CAP_PROXY));
	v.push_back(Desc(&d->cellBoxes, CAP_PROXY));
	v.push_back(Desc(&d->cellOfProxy, CAP_PROXY));
	v.push_back(Desc(&d->sBody, CAP_CONTACT));
	v.push_back(Desc(&d->sMass, CAP_CONTACT));
	v.push_back(Desc(&d->sNormal, CAP_CONTACT));
	v.push_back(Desc(&d->sP0a, CAP_CONTACT));
	v.push_back(Desc(&d->sP0b, CAP_CONTACT));
	v.push_back(Desc(&d->sP1a, CAP_CONTACT));
	v.push_back(Desc(&d->sImp, CAP_CONTACT));
	v.push_back(Desc(&d->sLocal, CAP_CONTACT));
	v.push_back(Desc(&d->sLocalP, CAP_CONTACT));
	v.push_back(Desc(&d->sCenters, CAP_CONTACT));
	v.push_back(Desc(&d->sRadius, CAP_CONTACT));
	v.push_back(Desc(&d->counters, CAP_FIXED, CNT_COUNT));
	return v;
}

size_t CapOf(const b2cuWorld* w, const ArrayDesc& a)
{
	switch (a.kind)
	{
	case CAP_BODY: return (size_t)w->bodyCapacity;
	case CAP_PROXY: return (size_t)w->proxyCapacity;
	case CAP_SHAPE: return (size_t)w->shapeCapacity;
	case CAP_CONTACT: return (size_t)w->contactCapacity;
	case CAP_GRID: return (size_t)w->gridSize + 1;
	case CAP_PROXY1: return (size_t)w->proxyCapacity + 1;
	default: return a.fixedCount;
	}
}

int NextPow2(int x)
{
	int p = 1;
	while (p < x) p <<= 1;
	return p;
}

// (Re)allocate every array of the given kind for new capacities, preserving contents.
int Reserve(b2cuWorld* w, int bodyCap, int proxyCap, int shapeCap, int contactCap)
{
	b2cuWorld old = *w;
	bool first = w->d.counters == nullptr;
	w->bodyCapacity = std::max(bodyCap, w->bodyCapacity);
	w->proxyCapacity = std::max(proxyCap, w->proxyCapacity);
	w->shapeCapacity = std::max(shapeCap, w->shapeCapacity);
	w->contactCapacity = std::max(contactCap, w->contactCapacity);
	w->gridSize = NextPow2(std::max(1024, 2 * w->proxyCapacity));

	std::vector<ArrayDesc> arrays = AllArrays(w);
	for (size_t k = 0; k < arrays.size(); ++k)
	{
		ArrayDesc& a = arrays[k];
		size_t newCap = CapOf(w, a);
		size_t oldCap = first ? 0 : CapOf(&old, a);
		if (!first && newCap == oldCap) continue;
		void* fresh = nullptr;
		CUDA_TRY(w, cudaMalloc(&fresh, newCap * a.elemSize));
		CUDA_TRY(w, cudaMemsetAsync(fresh, 0, newCap * a.elemSize, w->stream));
		if (!first && *a.ptr && oldCap > 0)
		{
			CUDA_TRY(w, cudaMemcpyAsync(fresh, *a.ptr, std::min(oldCap, newCap) * a.elemSize, cudaMemcpyDeviceToDevice,
			                            w->stream));
			CUDA_TRY(w, cudaStreamSynchronize(w->stream));
			cudaFree(*a.ptr);
		}
		*a.ptr = fresh;
	}

	// islandMinSep: [positionIterationsCapacity][bodyCapacity]
	if (first || old.bodyCapacity != w->bodyCapacity)
	{
		cudaFree(w->d.islandMinSep);
		CUDA_TRY(w, cudaMalloc(&w->d.islandMinSep,
		                       sizeof(int) * (size_t)w->positionIterationsCapacity * (size_t)w->bodyCapacity));
	}
	int primCap = std::max(std::max(w->contactCapacity, w->gridSize + 1), std::max(w->bodyCapacity, w->proxyCapacity));
	if (first || primCap > w->prims.capacity)
	{
		CUDA_TRY(w, PrimScratchAlloc(&w->prims, primCap));
	}
	return B2CU_OK;
}

int SyncCheck(b2cuWorld* w)
{
	CUDA_TRY(w, cudaStreamSynchronize(w->stream));
	CUDA_TRY(w, cudaGetLastError());
	return B2CU_OK;
}

// read all device counters (+ colour starts) into pinned host memory
int ReadCounters(b2cuWorld* w)
{
	CUDA_TRY(w, cudaMemcpyAsync(w->hostCounters, w->d.counters, sizeof(int) * CNT_COUNT, cudaMemcpyDeviceToHost,
	                            w->stream));
	CUDA_TRY(w, cudaMemcpyAsync(w->hostCounters + CNT_COUNT, w->d.colourCount, sizeof(int) * (B2CU_MAX_COLOURS + 2),
	                            cudaMemcpyDeviceToHost, w->stream));
	return SyncCheck(w);
}

int ZeroCounter(b2cuWorld* w, int index)
{
	CUDA_TRY(w, cudaMemsetAsync(w->d.counters + index, 0, sizeof(int), w->stream));
	return B2CU_OK;
}

float ChooseCellSize(const std::vector<float>& extents)
{
	// finest grid level = twice the median extent: the typical proxy stays on level 0 even when a fast step
	// stretches its fat box, so QueryUnmovedKernel (needed when a coarser-level proxy moves) stays idle
	if (extents.empty()) return 1.0f;
	std::vector<float> e(extents);
	size_t mid = e.size() / 2;
	std::nth_element(e.begin(), e.begin() + mid, e.end());
	float m = e[mid];
	if (!(m > 1e-3f)) m = 1e-3f;
	return 2.0f * m;
}

// Pin one body array (16 B x bodies: vel during the velocity iterations, pos during the position iterations) in
// L2 with an access-policy window: the constraint stream (hundreds of MB per iteration, loaded with evict-first
// hints) then cannot evict the state every constraint gathers and scatters.
void SetL2Window(b2cuWorld* w, const void* base, size_t bytes)
{
	if (w->l2WindowMax == 0) return;
	cudaStreamAttrValue attr;
	memset(&attr, 0, sizeof(attr));
	attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
	attr.accessPolicyWindow.num_bytes = std::min(bytes, w->l2WindowMax);
	attr.accessPolicyWindow.hitRatio = 1.0f;
	attr.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
	attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
	cudaStreamSetAttribute(w->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
}

// ---- sharding -------------------------------------------------------------------------------------------
const size_t kMailboxHeader = 256;
size_t MailboxFromUpperBytes(int ghostCount) { return (((size_t)ghostCount * 8 * sizeof(float4)) + 255) & ~(size_t)255; }
size_t MailboxFromLowerBytes(int exportCount) { return (((size_t)exportCount * 8 * sizeof(float4)) + 255) & ~(size_t)255; }

ShardState MakeShardState(b2cuWorld* w)
{
	ShardState sh;
	memset(&sh, 0, sizeof(sh));
	sh.rankCount = w->shardCount > 0 ? w->shardCount : 1;
	sh.stuck = w->d.counters + CNT_FLOW_STUCK;
	if (w->shardCount <= 1 || w->mailbox == nullptr) return sh;
	sh.ghostCount = w->ghostCount;
	sh.exportCount = w->exportCount;
	sh.ghostIds = w->ghostIds;
	sh.exportIds = w->exportIds;
	sh.flagFromUpper = reinterpret_cast<unsigned*>(w->mailbox);
	sh.flagFromLower = reinterpret_cast<unsigned*>(w->mailbox + 64);
	sh.fromUpper = reinterpret_cast<float4*>(w->mailbox + kMailboxHeader);
	sh.fromLower = reinterpret_cast<float4*>(w->mailbox + kMailboxHeader + MailboxFromUpperBytes(w->ghostCount));
	if (w->peerLower)
	{
		sh.lowerFlagFromUpper = reinterpret_cast<unsigned*>(w->peerLower);
		sh.lowerFromUpper = reinterpret_cast<float4*>(w->peerLower + kMailboxHeader);
	}
	if (w->peerUpper)
	{
		sh.upperFlagFromLower = reinterpret_cast<unsigned*>(w->peerUpper + 64);
		sh.upperFromLower =
			reinterpret_cast<float4*>(w->peerUpper + kMailboxHeader + MailboxFromUpperBytes(w->peerUpperGhostCountOfUpper));
	}
	sh.seq = w->shardSeq;
	return sh;
}

// step-start halo sync: owners push the full state of their export bodies to the ghost copies below them
int ShardSyncGhosts(b2cuWorld* w)
{
	if (w->shardCount <= 1) return B2CU_OK;
	ShardState sh = MakeShardState(w);
	const unsigned seq = w->shardSeq++;
	if (sh.lowerFromUpper != nullptr)
	{
		if (w->exportCount > 0) LAUNCH(w, GhostSendKernel, GridFor(w->exportCount), kBlock, w->d, sh);
		LAUNCH(w, ShardSignalKernel, 1, 1, sh.lowerFlagFromUpper, seq);
	}
	if (sh.upperFromLower != nullptr)
	{
		LAUNCH(w, ShardWaitKernel, 1, 1, sh.flagFromUpper, seq, w->d.counters + CNT_FLOW_STUCK);
		if (w->ghostCount > 0) LAUNCH(w, GhostApplyKernel, GridFor(w->ghostCount), kBlock, w->d, sh);
	}
	if (w->shardFlow)
	{
		// and the other way round: the upper neighbour may not begin its step (and push rows of it into this shard's
		// mailbox) before this shard has finished with the rows of the last one
		if (sh.upperFromLower != nullptr) LAUNCH(w, ShardSignalKernel, 1, 1, sh.upperFlagFromLower, seq);
		if (sh.lowerFromUpper != nullptr) LAUNCH(w, ShardWaitKernel, 1, 1, sh.flagFromLower, seq, w->d.counters + CNT_FLOW_STUCK);
	}
	return B2CU_OK;
}

// ascending sort of contact keys (min proxy << 32 | max proxy): LSD over the two id fields
void SortKeys(b2cuWorld* w, uint64_t* keys, int n)
{
	if (n <= 1) return;
	if (n <= B2CU_SMALL_SORT_MAX)
	{
		SortSmall64(&w->prims, keys, n, w->stream);
		return;
	}
	int bits = 1;
	while ((1 << bits) < std::max(2, w->proxyCount)) ++bits;
	RadixSort64(&w->prims, keys, n, 0, bits, w->stream);
	RadixSort64(&w->prims, keys, n, 32, 32 + bits, w->stream);
}

// index of the sorted key array by low proxy id (existence checks of the broad-phase)
int RebuildLowStart(b2cuWorld* w)
{
	LAUNCH(w, BuildLowStartKernel, GridFor(w->proxyCount + 1), kBlock, w->d, w->mainCount, w->proxyCount);
	return B2CU_OK;
}

// ---- broad-phase pair finding + contact set rebuild ------------------------------------------------------

// Finds new pairs for the proxies flagged MOVED, then rebuilds the contact set (drop destroyed, merge new).
// copy rows [begin, begin+count) of every contact array from one set to the other
int CopyContactRange(b2cuWorld* w, ContactSet& dst, const ContactSet& src, int begin, int count)
{
	if (count <= 0) return B2CU_OK;
#define COPY_ROWS(field)                                                                                         \
	CUDA_TRY(w, cudaMemcpyAsync(dst.field + begin, src.field + begin, sizeof(*dst.field) * (size_t)count,          \
	                            cudaMemcpyDeviceToDevice, w->stream))
	COPY_ROWS(key);
	COPY_ROWS(proxies);
	COPY_ROWS(flags);
	COPY_ROWS(m0);
	COPY_ROWS(m1);
	COPY_ROWS(m2);
	COPY_ROWS(m3);
	COPY_ROWS(mix);
	COPY_ROWS(toiCount);
	COPY_ROWS(colour);
	COPY_ROWS(stamp);
#undef COPY_ROWS
	return B2CU_OK;
}

int RebuildContactSet(b2cuWorld* w, int newCount, int destroyedMain, int destroyedTail, int* newCountOut, int* destroyedOut);

// Finds the new pairs of the proxies flagged MOVED and updates the contact set.
//
// The set is compacted lazily.  Slots [0, mainCount) are the big sorted main region, slots [mainCount, contactCount)
// a small sorted tail.  Contacts destroyed by Collide only get the DEAD flag.  Every step the tail is rebuilt (its
// live contacts merged with the new pairs: cost proportional to the churn, not to the set); the main region is
// merged with the tail, and its dead slots reclaimed, only when the tail or the dead slots exceed a few percent of
// it.  A settled million-body pile changes ~0.3% of its 5M contacts per step, so the full 1 GB merge runs every
// ~15 steps instead of every step.
int FindNewContactsAndRebuild(b2cuWorld* w, int* newCountOut, int* destroyedOut, int* movedOut)
{
	DeviceArrays& d = w->d;
	const int np = w->proxyCount;
	const int nc = w->contactCount;
	const int nMain = w->mainCount;
	const int nTail = nc - nMain;
	int rc;

	LAUNCH(w, BroadphaseResetKernel, 1, 128, d);

	GridParams grid;
	grid.cell0 = w->cellSize;
	grid.invCell0 = 1.0f / w->cellSize;
	grid.mask = (uint32_t)w->gridSize - 1u;
	const int2 counts = make_int2(nc, nMain);
	if (np > 0)
	{
		// cellCount is all zero here: Reserve clears it and GridFillKernel leaves it so (w->gridDirty: after a failed step)
		if (w->gridDirty) CUDA_TRY(w, cudaMemsetAsync(d.cellCount, 0, sizeof(int) * (w->gridSize + 1), w->stream));
		w->gridDirty = true;
		LAUNCH(w, GridCountKernel, GridFor(np), kBlock, d, np, grid);
		// one entry more than there are cells: cellStart[h + 1] closes the last cell's run
		ExclusiveScan(&w->prims, d.cellCount, d.cellStart, w->gridSize + 1, nullptr, w->stream);
		LAUNCH(w, GridFillKernel, GridFor(np), kBlock, d, np);
		w->gridDirty = false;
		LAUNCH(w, QueryMovedKernel, QueryGrid(np), 128, d, grid, counts, w->contactCapacity);
		LAUNCH(w, QueryUnmovedKernel, GridFor(np, 128), 128, d, np, grid, counts, w->contactCapacity);
	}

	if ((rc = ReadCounters(w))) return rc;
	const int destroyedMain = w->hostCounters[CNT_DESTROY];
	const int destroyedTail = w->hostCounters[CNT_DESTROY_B];
	const int tailLive = nTail - destroyedTail;
	if (w->hostCounters[CNT_ERROR] || nMain + tailLive + w->hostCounters[CNT_NEW_PAIRS] > w->contactCapacity)
	{
		// the pair buffer (or the set) would overflow: grow geometrically and repeat the queries; the pair counter
		// kept counting past the capacity, so the size needed is known
		int needed = nMain + tailLive + w->hostCounters[CNT_NEW_PAIRS];
		int grown = std::max(needed + needed / 4, 2 * w->contactCapacity);
		if ((rc = Reserve(w, w->bodyCapacity, w->proxyCapacity, w->shapeCapacity, grown))) return rc;
		DeviceArrays& dd = w->d;
		if ((rc = ZeroCounter(w, CNT_NEW_PAIRS))) return rc;
		if ((rc = ZeroCounter(w, CNT_ERROR))) return rc;
		LAUNCH(w, QueryMovedKernel, QueryGrid(np), 128, dd, grid, counts, w->contactCapacity);
		LAUNCH(w, QueryUnmovedKernel, GridFor(np, 128), 128, dd, np, grid, counts, w->contactCapacity);
		if ((rc = ReadCounters(w))) return rc;
		if (w->hostCounters[CNT_ERROR])
			return SetError(w, B2CU_ERR_CAPACITY, "new-pair buffer overflow (contact capacity %d)", w->contactCapacity);
	}
	if (np > 0) LAUNCH(w, ClearMovedKernel, GridFor(np), kBlock, w->d);
	*movedOut = w->hostCounters[CNT_MOVED];
	return RebuildContactSet(w, w->hostCounters[CNT_NEW_PAIRS], destroyedMain, destroyedTail, newCountOut, destroyedOut);
}

// Second half of FindNewContacts: the new pairs (unsorted keys in newKeys, `newCount` of them) become contacts and the
// destroyed ones leave the set.  Also used after every time-of-impact event (no destroyed contacts there).
int RebuildContactSet(b2cuWorld* w, int newCount, int destroyedMain, int destroyedTail, int* newCountOut, int* destroyedOut)
{
	DeviceArrays& d = w->d;
	const int nc = w->contactCount;
	const int nMain = w->mainCount;
	const int nTail = nc - nMain;
	const int tailLive = nTail - destroyedTail;
	int rc;
	if (w->pairFilter != nullptr && newCount > 0)
	{
		// the caller's b2ContactFilter decides about the candidate pairs before any contact exists (AddPair)
		std::vector<uint64_t> cand((size_t)newCount);
		std::vector<uint8_t> keep((size_t)newCount, 1);
		CUDA_TRY(w, cudaMemcpyAsync(cand.data(), w->d.newKeys, sizeof(uint64_t) * (size_t)newCount, cudaMemcpyDeviceToHost,
		                            w->stream));
		if ((rc = SyncCheck(w))) return rc;
		if (w->pairFilter(w->pairFilterUser, cand.data(), newCount, keep.data()) != 0)
			return SetError(w, B2CU_ERR_ARGUMENT, "the pair filter callback failed");
		int kept = 0;
		for (int i = 0; i < newCount; ++i)
			if (keep[i]) cand[kept++] = cand[i];
		if (kept != newCount)
		{
			if (kept > 0)
				CUDA_TRY(w, cudaMemcpyAsync(w->d.newKeys, cand.data(), sizeof(uint64_t) * (size_t)kept, cudaMemcpyHostToDevice,
				                            w->stream));
			if ((rc = SyncCheck(w))) return rc; // cand goes out of scope
			newCount = kept;
		}
	}
	*newCountOut = newCount;
	*destroyedOut = destroyedMain + destroyedTail;
	w->deadMain += destroyedMain;
	// every batch of new contacts gets the next creation stamp (b2cuContact::stamp)
	const uint32_t stamp = w->contactBatch;
	if (newCount > 0) ++w->contactBatch;

	// ---- 1. tail' = live(tail) merged with the new pairs (built in cAlt, copied back) ----
	int newTail = nTail;
	if (newCount > 0 || destroyedTail > 0)
	{
		newTail = tailLive + newCount;
		SortKeys(w, d.newKeys, newCount);
		if (nTail > 0)
		{
			ExclusiveScanNotMask(&w->prims, d.c.flags + nMain, B2CU_CONTACT_DEAD, d.listA, nTail, nullptr, w->stream);
			LAUNCH(w, MergeMoveKernel, GridFor(nTail), kBlock, d, nMain, nTail, d.listA, d.newKeys, newCount,
			       (const int*)nullptr, newCount, nMain);
		}
		if (newCount > 0)
		{
			LAUNCH(w, RebuildNewKernel, GridFor(newCount), kBlock, d, nMain, nTail, d.listA, tailLive, newCount, nMain, stamp);
			// the only body writer after the mirror copy has started: let the pack kernel finish reading first
			if (w->mirrorInFlight) CUDA_TRY(w, cudaStreamWaitEvent(w->stream, w->evPacked, 0));
			LAUNCH(w, ApplyWakeKernel, GridFor(w->bodyCount), kBlock, d, w->bodyCount,
			       w->mirrorInFlight ? d.wakePatch : (int*)nullptr);
		}
		if ((rc = CopyContactRange(w, d.c, d.cAlt, nMain, newTail))) return rc;
		w->contactCount = nMain + newTail;
	}

	// ---- 2. compaction of the main region when the tail or the dead slots have grown ----
	const int tailLimit = std::max(w->compactMin, nMain / 16);
	const int deadLimit = std::max(w->compactMin, nMain / 8);
	if (newTail > 0 && (newTail > tailLimit || w->deadMain > deadLimit || w->compactNow))
	{
		const int mainLive = nMain - w->deadMain;
		if (nMain > 0)
		{
			ExclusiveScanNotMask(&w->prims, d.c.flags, B2CU_CONTACT_DEAD, d.listA, nMain, nullptr, w->stream);
			LAUNCH(w, MergeMoveKernel, GridFor(nMain), kBlock, d, 0, nMain, d.listA, d.c.key + nMain, newTail,
			       (const int*)nullptr, newTail, 0);
		}
		// the tail has no dead entries (just rebuilt): its rank is the identity
		LAUNCH(w, IotaKernel, GridFor(newTail), kBlock, d.listB, newTail);
		LAUNCH(w, MergeMoveKernel, GridFor(newTail), kBlock, d, nMain, newTail, d.listB, d.c.key, nMain, d.listA, mainLive, 0);
		std::swap(d.c, d.cAlt);
		w->mainCount = mainLive + newTail;
		w->contactCount = w->mainCount;
		w->deadMain = 0;
		w->compactNow = false;
		if ((rc = RebuildLowStart(w))) return rc;
	}
	else if (newTail == 0 && (w->deadMain > deadLimit || (w->compactNow && w->deadMain > 0)))
	{
		// nothing to merge, only dead slots to reclaim
		const int mainLive = nMain - w->deadMain;
		ExclusiveScanNotMask(&w->prims, d.c.flags, B2CU_CONTACT_DEAD, d.listA, nMain, nullptr, w->stream);
		LAUNCH(w, MergeMoveKernel, GridFor(nMain), kBlock, d, 0, nMain, d.listA, d.newKeys, 0, (const int*)nullptr, 0, 0);
		std::swap(d.c, d.cAlt);
		w->mainCount = mainLive;
		w->contactCount = mainLive;
		w->deadMain = 0;
		w->compactNow = false;
		if ((rc = RebuildLowStart(w))) return rc;
	}
	return B2CU_OK;
}


// ---- continuous collision: b2World::SolveTOI (b2World.cpp:1026-1093) --------------------------------------------
// One FindMinToiContact pass: the earliest (alpha, key) among the eligible candidates lands in the counters.
int ToiFindMin(b2cuWorld* w)
{
	DeviceArrays& d = w->d;
	const int nc = w->contactCount;
	int rc;
	if ((rc = ZeroCounter(w, CNT_TOI))) return rc;
	if ((rc = ZeroCounter(w, CNT_TOI_WORK))) return rc;
	CUDA_TRY(w, cudaMemsetAsync(d.counters + CNT_TOI_MIN_ALPHA, 0xFF, sizeof(int) * 3, w->stream));
	LAUNCH(w, ToiSelectKernel, GridFor(nc), kBlock, d, nc, d.listA);
	// one candidate is a long, divergent root search: small blocks, so that a thousand candidates already spread over
	// most SMs (the count of the last pass sizes the grid; the kernel strides over whatever there is)
	const int toiBlock = 64;
	const int toiGrid = std::max(16, std::min(8 * g_smCount, (2 * std::max(w->toiCount, 512) + toiBlock - 1) / toiBlock));
	LAUNCH(w, ToiComputeKernel, toiGrid, toiBlock, d, (const int*)d.listA);
	LAUNCH(w, ToiMinKeyAllKernel, GridFor(nc), kBlock, d, nc);
	return ReadCounters(w);
}

// The sub-step loop.  Host-driven, one event at a time in the reference's order; all arithmetic and
// all state stay on the device, the host only reads the winner of each pass and the number of new pairs of each event.
int SolveTOI(b2cuWorld* w, float dt, int velocityIterations)
{
	DeviceArrays& d = w->d;
	int rc;
	w->toiCount = 0;
	w->toiMinAlpha = 1.0f;
	w->toiMinKey = ~0ull;
	w->toiEventPending = 0;
	w->toiSubSteps = 0;
	w->toiEventCount = 0;
	w->toiNewContacts = 0;
	if (w->contactCount == 0 || w->hostCounters[CNT_STICKY_TOI] == 0)
	{
		w->stepComplete = true;
		return B2CU_OK;
	}
	const bool subStepping = (w->params.flags & B2CU_WORLD_SUB_STEPPING) != 0;
	// flags are cleared at the end only if some contact was evaluated (or a previous call left them behind)
	bool clearPostSolve = !w->stepComplete;
	bool first = true;
	for (;;)
	{
		if ((rc = ToiFindMin(w))) return rc;
		uint32_t bits;
		memcpy(&bits, &w->hostCounters[CNT_TOI_MIN_ALPHA], sizeof(bits));
		const bool have = bits != 0xFFFFFFFFu;
		float minAlpha = 1.0f;
		uint64_t minKey = ~0ull;
		if (have)
		{
			memcpy(&minAlpha, &bits, sizeof(float));
			memcpy(&minKey, &w->hostCounters[CNT_TOI_MIN_KEY], sizeof(uint64_t));
			clearPostSolve = true;
		}
		if (first)
		{
			w->toiCount = w->hostCounters[CNT_TOI];
			w->toiMinAlpha = minAlpha;
			w->toiMinKey = minKey;
			first = false;
		}
		if (!have || 1.0f - 10.0f * B2CU_EPSILON < minAlpha)
		{
			// no more events (b2World.cpp:1069-1078)
			w->stepComplete = true;
			if (clearPostSolve) LAUNCH(w, ToiClearKernel, GridFor(std::max(w->contactCount, w->bodyCount)), kBlock, d, w->contactCount, w->bodyCount);
			break;
		}
		if (w->shardCount > 1)
			return SetError(w, B2CU_ERR_UNSUPPORTED, "time-of-impact sub-steps are not supported in a sharded world "
			                                         "(switch continuous physics off or keep the world on one GPU)");

		// ---- StepSolveTOI (b2World.cpp:851-1024) ----
		const int nc = w->contactCount;
		const int np = w->proxyCount;
		if ((rc = ZeroCounter(w, CNT_TOI_LIST))) return rc;
		if ((rc = ZeroCounter(w, CNT_SCRATCH))) return rc;
		if ((rc = ZeroCounter(w, CNT_NEW_PAIRS))) return rc;
		LAUNCH(w, ToiEventPrepareKernel, GridFor(nc), kBlock, d, nc, w->mainCount, minKey, w->contactCapacity);
		LAUNCH(w, ToiEventKernel, 1, B2CU_TOI_THREADS, d, nc, w->mainCount, minKey, minAlpha, dt, velocityIterations,
		       w->contactCapacity);
		LAUNCH(w, ToiAfterEventKernel, GridFor(std::max(np, nc)), kBlock, d, np, nc);
		const int2 counts = make_int2(nc, w->mainCount);
		LAUNCH(w, ToiFindPairsKernel, GridFor(np), kBlock, d, np, counts, w->contactCapacity);
		if ((rc = ReadCounters(w))) return rc;
		if (w->hostCounters[CNT_ERROR] || nc + w->hostCounters[CNT_NEW_PAIRS] > w->contactCapacity)
		{
			if (w->hostCounters[CNT_TOI_LIST] > w->contactCapacity || w->hostCounters[CNT_TOI_EVENTS] > w->contactCapacity)
				return SetError(w, B2CU_ERR_CAPACITY, "time-of-impact event: list %d / events %d exceed the contact capacity %d",
				                w->hostCounters[CNT_TOI_LIST], w->hostCounters[CNT_TOI_EVENTS], w->contactCapacity);
			// the pair buffer would overflow: grow and repeat the pair search (the moved flags are still set)
			int needed = nc + w->hostCounters[CNT_NEW_PAIRS];
			int grown = std::max(needed + needed / 4, 2 * w->contactCapacity);
			if ((rc = Reserve(w, w->bodyCapacity, w->proxyCapacity, w->shapeCapacity, grown))) return rc;
			if ((rc = ZeroCounter(w, CNT_NEW_PAIRS))) return rc;
			if ((rc = ZeroCounter(w, CNT_ERROR))) return rc;
			LAUNCH(w, ToiFindPairsKernel, GridFor(np), kBlock, w->d, np, counts, w->contactCapacity);
			if ((rc = ReadCounters(w))) return rc;
			if (w->hostCounters[CNT_ERROR])
				return SetError(w, B2CU_ERR_CAPACITY, "new-pair buffer overflow (contact capacity %d)", w->contactCapacity);
		}
		if (w->hostCounters[CNT_SCRATCH] > 0) LAUNCH(w, ClearMovedKernel, GridFor(w->hostCounters[CNT_SCRATCH]), kBlock, w->d);
		int created = 0, destroyed = 0;
		if (w->hostCounters[CNT_NEW_PAIRS] > 0 &&
		    (rc = RebuildContactSet(w, w->hostCounters[CNT_NEW_PAIRS], 0, 0, &created, &destroyed)))
			return rc;
		w->toiNewContacts += created;
		++w->toiSubSteps;
		if (subStepping)
		{
			w->stepComplete = false;
			w->toiEventPending = 1;
			break;
		}
	}
	w->toiEventCount = std::min(w->hostCounters[CNT_TOI_EVENTS], w->contactCapacity);
	return B2CU_OK;
}

int CheckRange(b2cuWorld* w, int first, int count, int limit, const void* p)
{
	if (!w || !p || first < 0 || count < 0 || first + count > limit)
		return SetError(w, B2CU_ERR_ARGUMENT, "range [%d,%d) outside [0,%d)", first, first + count, limit);
	return B2CU_OK;
}

template <typename T>
int Upload(b2cuWorld* w, T* dst, int first, const std::vector<T>& src)
{
	if (src.empty()) return B2CU_OK;
	CUDA_TRY(w, cudaMemcpyAsync(dst + first, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, w->stream));
	return B2CU_OK;
}

template <typename T>
int Download(b2cuWorld* w, const T* src, int first, std::vector<T>& dst)
{
	if (dst.empty()) return B2CU_OK;
	CUDA_TRY(w, cudaMemcpyAsync(dst.data(), src + first, sizeof(T) * dst.size(), cudaMemcpyDeviceToHost, w->stream));
	return B2CU_OK;
}

} // namespace

// =========================================================================================================
// C ABI
// =========================================================================================================

extern "C" {

int b2cuGetDeviceCount(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

const char* b2cuVersion(void) { return "b2cuda 0.1 (sm_100a)"; }

int b2cuCreateWorld(const b2cuWorldDef* def, b2cuWorld** out)
{
	if (!def || !out) return B2CU_ERR_ARGUMENT;
	*out = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return B2CU_ERR_NO_DEVICE;
	if (def->device < 0 || def->device >= n) return B2CU_ERR_ARGUMENT;
	if (cudaSetDevice(def->device) != cudaSuccess) return B2CU_ERR_CUDA;

	b2cuWorld* w = new b2cuWorld();
	memset(w, 0, sizeof(*w));
	new (&w->prims) PrimScratch();
	w->device = def->device;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, def->device) == cudaSuccess)
	{
		g_smCount = prop.multiProcessorCount;
		// developer knob: DRAM->L2 fetch granularity (32/64/128 B); the step is dominated by 16-byte gathers
		const char* fg = getenv("B2CU_L2_FETCH");
		if (fg && atoi(fg) > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg));
		const char* l2env = getenv("B2CU_L2_WINDOW");
		if (l2env && atoi(l2env) > 0 && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0)
		{
			size_t want = std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)atoi(l2env) << 20);
			if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
				w->l2WindowMax = std::min(want, (size_t)prop.accessPolicyMaxWindowSize);
		}
	}
	if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess)
	{
		delete w;
		return B2CU_ERR_CUDA;
	}
	for (int i = 0; i < 10; ++i) cudaEventCreate(&w->ev[i]);
	cudaStreamCreateWithFlags(&w->copyStream, cudaStreamNonBlocking);
	cudaEventCreateWithFlags(&w->evBodiesFinal, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&w->evPacked, cudaEventDisableTiming);
	cudaMallocHost(&w->hostCounters, sizeof(int) * (CNT_COUNT + B2CU_MAX_COLOURS + 2));
	w->params.gravity = make_float2(def->gravity[0], def->gravity[1]);
	w->params.flags = def->flags;
	w->params.invDt0 = 0.0f;
	w->positionIterationsCapacity = 4;
	w->cellSize = 1.0f;
	w->toiCheckDirty = true;
	w->stepComplete = true;
	w->contactBatch = 1;
	{
		// persistent cooperative solver: as many CTAs as can be co-resident
		int coop = 0, perSm = 0;
		cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, def->device);
		int perSmPos = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, SolverVelocityPersistentKernel<false>, B2CU_SOLVER_THREADS, 0);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmPos, SolverPositionPersistentKernel<false>, B2CU_SOLVER_THREADS, 0);
		int perSmJ = 0, perSmPosJ = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmJ, SolverVelocityPersistentKernel<true>, B2CU_SOLVER_THREADS, 0);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmPosJ, SolverPositionPersistentKernel<true>, B2CU_SOLVER_THREADS, 0);
		w->persistentGridJoints = g_smCount * std::max(perSmJ, 0);
		w->persistentGridPositionJoints = g_smCount * std::max(perSmPosJ, 0);
		{
			const char* pdl = getenv("B2CU_PDL");
			g_pdl = !(pdl && atoi(pdl) == 0);
		}
		const char* pe = getenv("B2CU_PERSISTENT");
		w->persistentSolver = coop != 0 && perSm > 0 && perSmPos > 0 && !(pe && atoi(pe) == 0);
		const char* pb = getenv("B2CU_PERSISTENT_BLOCKS");
		if (pb && atoi(pb) > 0)
		{
			perSm = std::min(perSm, atoi(pb));
			perSmPos = std::min(perSmPos, atoi(pb));
		}
		w->persistentGrid = g_smCount * perSm;
		w->persistentGridMax = w->persistentGrid;
		{
			int perSmFlow = 0, perSmFlowPos = 0;
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmFlow, SolverVelocityFlowKernel<false, false>, B2CU_SOLVER_THREADS, 0);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmFlowPos, SolverPositionFlowKernel<false, false>, B2CU_SOLVER_THREADS, 0);
			int perSmFlowS = 0, perSmFlowPosS = 0;
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmFlowS, SolverVelocityFlowKernel<true, false>, B2CU_SOLVER_THREADS, 0);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmFlowPosS, SolverPositionFlowKernel<true, false>, B2CU_SOLVER_THREADS, 0);
			w->flowGrid = coop != 0 ? g_smCount * perSmFlow : 0;
			w->flowGridPosition = coop != 0 ? g_smCount * perSmFlowPos : 0;
			w->flowGridMax = coop != 0 ? g_smCount * perSmFlowS : 0;          // sharded instances
			w->flowGridPositionMax = coop != 0 ? g_smCount * perSmFlowPosS : 0;
		}
		w->persistentGridPosition = g_smCount * perSmPos;
		w->persistentGridPositionMax = w->persistentGridPosition;
		w->shardCount = 1;
		w->shardSeq = 1;
	}
	{
		// smallest tail / dead-slot count that triggers a compaction of the contact set (tests lower it)
		const char* cm = getenv("B2CU_COMPACT_MIN");
		w->compactMin = cm && atoi(cm) > 0 ? atoi(cm) : 8192;
	}
	{
		const char* t = getenv("B2CU_TRACE");
		g_trace.enabled = t && atoi(t) > 0;
	}
	int rc = Reserve(w, std::max(def->bodyCapacity, 64), std::max(def->proxyCapacity, 64),
	                 std::max(def->shapeCapacity, 16), std::max(def->contactCapacity, 256));
	if (rc != B2CU_OK)
	{
		fprintf(stderr, "b2cuCreateWorld: %s\n", w->lastError);
		b2cuDestroyWorld(w);
		return rc;
	}
	*out = w;
	return B2CU_OK;
}

void b2cuDestroyWorld(b2cuWorld* w)
{
	if (w && w->softBarrierCounter) cudaFree(w->softBarrierCounter);
	if (!w) return;
	cudaSetDevice(w->device);
	cudaStreamSynchronize(w->stream);
	std::vector<ArrayDesc> arrays = AllArrays(w);
	for (size_t k = 0; k < arrays.size(); ++k) cudaFree(*arrays[k].ptr);
	cudaFree(w->d.islandMinSep);
	cudaFree(w->d.joints);
	cudaFree(w->d.jointRows);
	cudaFree(w->d.jointOrder);
	cudaFree((void*)w->d.jointPairKeys);
	cudaFree((void*)w->d.jointFreedKeys);
	free(w->jointPairsHost);
	free(w->jointFreedHost);
	cudaFree(w->bodyStage);
	cudaFree(w->queryScratch);
	if (w->queryHost) cudaFreeHost(w->queryHost);
	if (w->hostPatch) cudaFreeHost(w->hostPatch);
	free(w->eventOrder);
	if (w->copyStream) cudaStreamDestroy(w->copyStream);
	if (w->evBodiesFinal) cudaEventDestroy(w->evBodiesFinal);
	if (w->evPacked) cudaEventDestroy(w->evPacked);
	if (w->peerLower && w->peerLowerIpc) cudaIpcCloseMemHandle(w->peerLower);
	if (w->peerUpper && w->peerUpperIpc) cudaIpcCloseMemHandle(w->peerUpper);
	cudaFree(w->ghostIds);
	cudaFree(w->exportIds);
	cudaFree(w->mailbox);
	PrimScratchFree(&w->prims);
	cudaFreeHost(w->hostCounters);
	for (int i = 0; i < 10; ++i) cudaEventDestroy(w->ev[i]);
	cudaStreamDestroy(w->stream);
	delete w;
}

const char* b2cuGetLastError(const b2cuWorld* w) { return w ? w->lastError : "null world"; }

int b2cuSetWorldParams(b2cuWorld* w, const float gravity[2], uint32_t flags)
{
	if (!w || !gravity) return B2CU_ERR_ARGUMENT;
	w->params.gravity = make_float2(gravity[0], gravity[1]);
	w->params.flags = flags;
	return B2CU_OK;
}

int b2cuSetInvDt0(b2cuWorld* w, float invDt0)
{
	if (!w) return B2CU_ERR_ARGUMENT;
	w->params.invDt0 = invDt0;
	return B2CU_OK;
}

int b2cuSetCounts(b2cuWorld* w, int32_t bodyCount, int32_t shapeCount, int32_t proxyCount)
{
	if (!w || bodyCount < 0 || shapeCount < 0 || proxyCount < 0) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	if (bodyCount > w->bodyCapacity || shapeCount > w->shapeCapacity || proxyCount > w->proxyCapacity)
	{
		int bc = bodyCount > w->bodyCapacity ? std::max(bodyCount, 2 * w->bodyCapacity) : w->bodyCapacity;
		int sc = shapeCount > w->shapeCapacity ? std::max(shapeCount, 2 * w->shapeCapacity) : w->shapeCapacity;
		int pc = proxyCount > w->proxyCapacity ? std::max(proxyCount, 2 * w->proxyCapacity) : w->proxyCapacity;
		int cc = std::max(w->contactCapacity, 8 * pc);
		int rc = Reserve(w, bc, pc, sc, cc);
		if (rc) return rc;
	}
	w->bodyCount = bodyCount;
	w->shapeCount = shapeCount;
	w->proxyCount = proxyCount;
	return RebuildLowStart(w);
}

// device staging for the record <-> column conversion of the bodies
static int EnsureBodyStage(b2cuWorld* w)
{
	if (w->bodyStage && w->bodyStageCapacity >= w->bodyCapacity) return B2CU_OK;
	if (w->bodyStage)
	{
		CUDA_TRY(w, cudaStreamSynchronize(w->stream));
		cudaFree(w->bodyStage);
		w->bodyStage = nullptr;
	}
	CUDA_TRY(w, cudaMalloc(&w->bodyStage, sizeof(b2cuBody) * (size_t)w->bodyCapacity));
	w->bodyStageCapacity = w->bodyCapacity;
	return B2CU_OK;
}

int b2cuSetBodies(b2cuWorld* w, int32_t first, int32_t count, const b2cuBody* bodies)
{
	int rc = CheckRange(w, first, count, w ? w->bodyCount : 0, bodies);
	if (rc) return rc;
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	if ((rc = EnsureBodyStage(w))) return rc;
	float* stage = w->bodyStage + (size_t)first * B2CU_BODY_WORDS;
	CUDA_TRY(w, cudaMemcpyAsync(stage, bodies, sizeof(b2cuBody) * (size_t)count, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(w, UnpackBodiesKernel, GridFor(count), kBlock, w->d, first, count, (const float*)stage);
	w->toiCheckDirty = true;
	// what depends on the bodies' types and bullet flags is redone only if one of them really changed
	CUDA_TRY(w, cudaMemcpyAsync(&w->hostCounters[CNT_BODY_TYPE_CHANGED], w->d.counters + CNT_BODY_TYPE_CHANGED, sizeof(int),
	                            cudaMemcpyDeviceToHost, w->stream));
	if ((rc = SyncCheck(w))) return rc;
	const int changed = w->hostCounters[CNT_BODY_TYPE_CHANGED];
	if (changed)
	{
		if ((changed & 1) && w->d.jointCount > 0) w->jointColourDirty = true;
		// b2Body::SetBullet / SetType -> b2ContactManager::RecalculateToiCandidacy (b2ContactManager.cpp:566-640)
		if (changed & 2) w->contactBodiesDirty = true;
		w->hostCounters[CNT_BODY_TYPE_CHANGED] = 0;
		return ZeroCounter(w, CNT_BODY_TYPE_CHANGED);
	}
	return B2CU_OK;
}

int b2cuGetBodies(b2cuWorld* w, int32_t first, int32_t count, b2cuBody* bodies)
{
	int rc = CheckRange(w, first, count, w ? w->bodyCount : 0, bodies);
	if (rc) return rc;
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	if ((rc = EnsureBodyStage(w))) return rc;
	float* stage = w->bodyStage + (size_t)first * B2CU_BODY_WORDS;
	LAUNCH(w, PackBodiesKernel, GridFor(count), kBlock, w->d, first, count, stage);
	CUDA_TRY(w, cudaMemcpyAsync(bodies, stage, sizeof(b2cuBody) * (size_t)count, cudaMemcpyDeviceToHost, w->stream));
	return SyncCheck(w);
}

// page-locked host memory for the caller's mirrors: transfers from ordinary (pageable) memory are staged by the
// driver at a fraction of the PCIe rate
namespace
{
std::mutex g_pinnedMutex;
std::unordered_set<void*> g_pinned;
}

void* b2cuHostAlloc(size_t bytes)
{
	if (bytes == 0) bytes = 1;
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess && p)
	{
		std::lock_guard<std::mutex> lock(g_pinnedMutex);
		g_pinned.insert(p);
		return p;
	}
	cudaGetLastError();
	return malloc(bytes);
}

void b2cuHostFree(void* p)
{
	if (p == nullptr) return;
	bool pinned = false;
	{
		std::lock_guard<std::mutex> lock(g_pinnedMutex);
		pinned = g_pinned.erase(p) > 0;
	}
	if (pinned) cudaFreeHost(p);
	else free(p);
}

int b2cuSetShapes(b2cuWorld* w, int32_t first, int32_t count, const b2cuShape* shapes)
{
	int rc = CheckRange(w, first, count, w ? w->shapeCount : 0, shapes);
	if (rc) return rc;
	cudaSetDevice(w->device);
	for (int i = 0; i < count; ++i)
	{
		if (shapes[i].type < B2CU_SHAPE_CIRCLE || shapes[i].type > B2CU_SHAPE_POLYGON)
			return SetError(w, B2CU_ERR_UNSUPPORTED, "shape %d: type %d is outside the GPU path (chains unsupported)",
			                first + i, shapes[i].type);
	}
	if (count > 0)
	{
		CUDA_TRY(w, cudaMemcpyAsync(w->d.shapes + first, shapes, sizeof(b2cuShape) * count, cudaMemcpyHostToDevice,
		                            w->stream));
		w->contactBodiesDirty = true; // pradius is derived from the shape table
	}
	return SyncCheck(w);
}

int b2cuSetProxies(b2cuWorld* w, int32_t first, int32_t count, const b2cuProxy* proxies)
{
	int rc = CheckRange(w, first, count, w ? w->proxyCount : 0, proxies);
	if (rc) return rc;
	cudaSetDevice(w->device);
	std::vector<float4> fat(count), aabb(count);
	std::vector<int> body(count), shape(count), fixture(count);
	std::vector<uint32_t> filter(count), group(count);
	std::vector<float2> mat(count);
	std::vector<float> extents;
	extents.reserve(count);
	bool anyMoved = false;
	for (int i = 0; i < count; ++i)
	{
		const b2cuProxy& p = proxies[i];
		if (p.body < 0 || p.body >= w->bodyCount || p.shape < 0 || p.shape >= w->shapeCount)
			return SetError(w, B2CU_ERR_ARGUMENT, "proxy %d: body %d / shape %d out of range", first + i, p.body, p.shape);
		fat[i] = make_float4(p.fat[0], p.fat[1], p.fat[2], p.fat[3]);
		aabb[i] = make_float4(p.aabb[0], p.aabb[1], p.aabb[2], p.aabb[3]);
		body[i] = p.body;
		shape[i] = p.shape;
		fixture[i] = p.fixture;
		filter[i] = (uint32_t)p.categoryBits | ((uint32_t)p.maskBits << 16);
		group[i] = (uint32_t)(uint16_t)p.groupIndex | ((uint32_t)(p.flags & B2CU_PROXY_PUBLIC_FLAGS) << 16);
		mat[i] = make_float2(p.friction, p.restitution);
		extents.push_back(std::max(p.fat[2] - p.fat[0], p.fat[3] - p.fat[1]));
		// the move buffer is only processed ahead of the step when a fixture is new (e_newFixture)
		if ((p.flags & B2CU_PROXY_MOVED) && (p.flags & B2CU_PROXY_NEW)) anyMoved = true;
		if (p.flags & B2CU_PROXY_REFILTER) w->refilterPending = true;
	}
	DeviceArrays& d = w->d;
	if ((rc = Upload(w, d.fat, first, fat)) || (rc = Upload(w, d.aabb, first, aabb)) ||
	    (rc = Upload(w, d.pbody, first, body)) || (rc = Upload(w, d.pshape, first, shape)) ||
	    (rc = Upload(w, d.pfixture, first, fixture)) || (rc = Upload(w, d.pfilter, first, filter)) ||
	    (rc = Upload(w, d.pgroup, first, group)) || (rc = Upload(w, d.pmat, first, mat)))
		return rc;
	if (anyMoved) w->newProxies = true;
	w->toiCheckDirty = true;
	w->contactBodiesDirty = true;
	if (first == 0 && count == w->proxyCount && count > 0)
	{
		// the whole table: the finest grid level follows the typical proxy
		w->cellSize = ChooseCellSize(extents);
		w->cellChosenCount = count;
	}
	else if (w->proxyCount >= 2 * std::max(1, w->cellChosenCount))
	{
		// the world has doubled since the cell size was chosen (a program that creates bodies as it runs): choose again
		// from all the boxes.  An upload of a few rows (a ground fixture whose friction changed) never moves the cell size.
		std::vector<float4> all((size_t)w->proxyCount);
		if ((rc = Download(w, d.fat, 0, all)) || (rc = SyncCheck(w))) return rc;
		std::vector<float> ext(all.size());
		for (size_t i = 0; i < all.size(); ++i) ext[i] = std::max(all[i].z - all[i].x, all[i].w - all[i].y);
		w->cellSize = ChooseCellSize(ext);
		w->cellChosenCount = w->proxyCount;
	}
	return SyncCheck(w);
}

int b2cuGetProxies(b2cuWorld* w, int32_t first, int32_t count, b2cuProxy* proxies)
{
	int rc = CheckRange(w, first, count, w ? w->proxyCount : 0, proxies);
	if (rc) return rc;
	cudaSetDevice(w->device);
	std::vector<float4> fat(count), aabb(count);
	std::vector<int> body(count), shape(count), fixture(count);
	std::vector<uint32_t> filter(count), group(count);
	std::vector<float2> mat(count);
	DeviceArrays& d = w->d;
	if ((rc = Download(w, d.fat, first, fat)) || (rc = Download(w, d.aabb, first, aabb)) ||
	    (rc = Download(w, d.pbody, first, body)) || (rc = Download(w, d.pshape, first, shape)) ||
	    (rc = Download(w, d.pfixture, first, fixture)) || (rc = Download(w, d.pfilter, first, filter)) ||
	    (rc = Download(w, d.pgroup, first, group)) || (rc = Download(w, d.pmat, first, mat)))
		return rc;
	if ((rc = SyncCheck(w))) return rc;
	for (int i = 0; i < count; ++i)
	{
		b2cuProxy& p = proxies[i];
		p.aabb[0] = aabb[i].x; p.aabb[1] = aabb[i].y; p.aabb[2] = aabb[i].z; p.aabb[3] = aabb[i].w;
		p.fat[0] = fat[i].x; p.fat[1] = fat[i].y; p.fat[2] = fat[i].z; p.fat[3] = fat[i].w;
		p.body = body[i];
		p.shape = shape[i];
		p.friction = mat[i].x;
		p.restitution = mat[i].y;
		p.categoryBits = (uint16_t)(filter[i] & 0xFFFFu);
		p.maskBits = (uint16_t)(filter[i] >> 16);
		p.groupIndex = (int16_t)(group[i] & 0xFFFFu);
		p.flags = (uint16_t)((group[i] >> 16) & B2CU_PROXY_PUBLIC_FLAGS);
		p.fixture = fixture[i];
		p.child = 0;
	}
	return B2CU_OK;
}

// ---- joints ------------------------------------------------------------------------------------------------------------

// Colour classes of the joint table: joints in id order take the lowest class in which neither of their dynamic bodies
// has a joint yet (non-dynamic bodies are not changed by an impulse, so they may be shared); what does not fit the
// parallel classes is solved by one thread, in id order.  The solve order is (class, id).
static int ColourJoints(b2cuWorld* w)
{
	const int nj = w->d.jointCount;
	w->jointOpCount = 0;
	w->jointColourDirty = false;
	if (nj == 0) return B2CU_OK;
	std::vector<b2cuJoint> joints((size_t)nj);
	std::vector<uint32_t> bflags((size_t)w->bodyCount);
	CUDA_TRY(w, cudaMemcpyAsync(joints.data(), w->d.joints, sizeof(b2cuJoint) * (size_t)nj, cudaMemcpyDeviceToHost, w->stream));
	CUDA_TRY(w, cudaMemcpyAsync(bflags.data(), w->d.bflags, sizeof(uint32_t) * (size_t)w->bodyCount, cudaMemcpyDeviceToHost,
	                            w->stream));
	int rc = SyncCheck(w);
	if (rc) return rc;
	std::unordered_map<int, uint32_t> used; // dynamic body -> classes taken
	std::vector<std::vector<int> > classes(B2CU_MAX_JOINT_COLOURS + 1);
	for (int j = 0; j < nj; ++j)
	{
		uint32_t mask = 0;
		// a gear joint also writes the first bodies of its two joints
		const int nBodies = joints[j].type == B2CU_JOINT_GEAR ? 4 : 2;
		const int bodies[4] = {joints[j].bodyA, joints[j].bodyB, joints[j].limitState, joints[j].reserved};
		bool dynamic[4];
		for (int k = 0; k < nBodies; ++k)
		{
			dynamic[k] = (bflags[bodies[k]] & B2CU_BODY_TYPE_MASK) == B2CU_DYNAMIC_BODY;
			if (k == 0 && joints[j].type == B2CU_JOINT_MOUSE) dynamic[k] = false; // a mouse joint writes body B only
			if (dynamic[k]) mask |= used[bodies[k]];
		}
		int colour = 0;
		while (colour < B2CU_MAX_JOINT_COLOURS && (mask & (1u << colour))) ++colour;
		classes[colour].push_back(j);
		if (colour < B2CU_MAX_JOINT_COLOURS)
			for (int k = 0; k < nBodies; ++k)
				if (dynamic[k]) used[bodies[k]] |= 1u << colour;
	}
	std::vector<int> order;
	order.reserve((size_t)nj);
	for (int c = 0; c <= B2CU_MAX_JOINT_COLOURS; ++c)
	{
		if (classes[c].empty()) continue;
		int op = w->jointOpCount++;
		w->jointOpStart[op] = (int)order.size();
		w->jointOpSize[op] = (int)classes[c].size();
		w->jointOpSerial[op] = c == B2CU_MAX_JOINT_COLOURS ? 1 : 0;
		order.insert(order.end(), classes[c].begin(), classes[c].end());
	}
	CUDA_TRY(w, cudaMemcpyAsync(w->d.jointOrder, order.data(), sizeof(int) * (size_t)nj, cudaMemcpyHostToDevice, w->stream));
	return SyncCheck(w);
}

int b2cuSetJoints(b2cuWorld* w, int32_t count, const b2cuJoint* joints)
{
	if (!w || count < 0 || (count > 0 && !joints)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	if (count > 0 && w->shardCount > 1) return SetError(w, B2CU_ERR_UNSUPPORTED, "joints in a sharded world");
	std::vector<uint64_t> pairs;
	for (int32_t j = 0; j < count; ++j)
	{
		const b2cuJoint& jt = joints[j];
		if (jt.type < B2CU_JOINT_REVOLUTE || jt.type > B2CU_JOINT_MOTOR)
			return SetError(w, B2CU_ERR_UNSUPPORTED, "joint %d: unknown type %d", j, jt.type);
		if (jt.type == B2CU_JOINT_GEAR &&
		    (jt.limitState < 0 || jt.limitState >= w->bodyCount || jt.reserved < 0 || jt.reserved >= w->bodyCount))
			return SetError(w, B2CU_ERR_ARGUMENT, "gear joint %d: bodies C, D = %d, %d", j, jt.limitState, jt.reserved);
		if (jt.bodyA < 0 || jt.bodyA >= w->bodyCount || jt.bodyB < 0 || jt.bodyB >= w->bodyCount || jt.bodyA == jt.bodyB)
			return SetError(w, B2CU_ERR_ARGUMENT, "joint %d: bodies %d, %d", j, jt.bodyA, jt.bodyB);
		if (!(jt.flags & B2CU_JOINT_COLLIDE_CONNECTED))
		{
			uint32_t lo = (uint32_t)std::min(jt.bodyA, jt.bodyB), hi = (uint32_t)std::max(jt.bodyA, jt.bodyB);
			pairs.push_back(((uint64_t)lo << 32) | hi);
		}
	}
	std::sort(pairs.begin(), pairs.end());
	pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
	// body pairs that stop being kept apart: the pair search must look at them afresh (JointFreed)
	{
		std::vector<uint64_t> freed(w->jointFreedHost, w->jointFreedHost + w->jointFreedHostCount);
		for (int k = 0; k < w->jointPairsHostCount; ++k)
			if (!std::binary_search(pairs.begin(), pairs.end(), w->jointPairsHost[k])) freed.push_back(w->jointPairsHost[k]);
		std::sort(freed.begin(), freed.end());
		freed.erase(std::unique(freed.begin(), freed.end()), freed.end());
		std::vector<uint64_t> kept;
		for (size_t k = 0; k < freed.size(); ++k)
			if (!std::binary_search(pairs.begin(), pairs.end(), freed[k])) kept.push_back(freed[k]);
		if ((int)kept.size() > w->jointFreedCapacity)
		{
			int cap = std::max((int)kept.size(), std::max(64, 2 * w->jointFreedCapacity));
			cudaFree((void*)w->d.jointFreedKeys);
			w->d.jointFreedKeys = nullptr;
			w->d.jointFreedCount = 0;
			free(w->jointFreedHost);
			w->jointFreedHost = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)cap);
			w->jointFreedHostCount = 0;
			w->jointFreedCapacity = 0;
			uint64_t* keys = nullptr;
			CUDA_TRY(w, cudaMalloc(&keys, sizeof(uint64_t) * (size_t)cap));
			w->d.jointFreedKeys = keys;
			w->jointFreedCapacity = cap;
		}
		if (!kept.empty())
		{
			memcpy(w->jointFreedHost, kept.data(), sizeof(uint64_t) * kept.size());
			CUDA_TRY(w, cudaMemcpyAsync((void*)w->d.jointFreedKeys, kept.data(), sizeof(uint64_t) * kept.size(),
			                            cudaMemcpyHostToDevice, w->stream));
			CUDA_TRY(w, cudaStreamSynchronize(w->stream));
		}
		w->jointFreedHostCount = (int)kept.size();
		w->d.jointFreedCount = (int)kept.size();
		free(w->jointPairsHost);
		w->jointPairsHost = pairs.empty() ? nullptr : (uint64_t*)malloc(sizeof(uint64_t) * pairs.size());
		if (!pairs.empty()) memcpy(w->jointPairsHost, pairs.data(), sizeof(uint64_t) * pairs.size());
		w->jointPairsHostCount = (int)pairs.size();
	}
	if (count > w->jointCapacity)
	{
		int cap = std::max(count, std::max(64, 2 * w->jointCapacity));
		cudaFree(w->d.joints);
		cudaFree(w->d.jointRows);
		cudaFree(w->d.jointOrder);
		cudaFree((void*)w->d.jointPairKeys);
		w->d.joints = nullptr;
		w->d.jointRows = nullptr;
		w->d.jointOrder = nullptr;
		w->d.jointPairKeys = nullptr;
		w->jointCapacity = 0;
		w->d.jointCount = 0;
		w->d.jointPairCount = 0;
		uint64_t* keys = nullptr;
		CUDA_TRY(w, cudaMalloc(&w->d.joints, sizeof(b2cuJoint) * (size_t)cap));
		CUDA_TRY(w, cudaMalloc(&w->d.jointRows, sizeof(JointRow) * (size_t)cap));
		CUDA_TRY(w, cudaMalloc(&w->d.jointOrder, sizeof(int) * (size_t)cap));
		CUDA_TRY(w, cudaMalloc(&keys, sizeof(uint64_t) * (size_t)cap));
		w->d.jointPairKeys = keys;
		w->jointCapacity = cap;
	}
	if (count > 0)
	{
		CUDA_TRY(w, cudaMemcpyAsync(w->d.joints, joints, sizeof(b2cuJoint) * (size_t)count, cudaMemcpyHostToDevice, w->stream));
		CUDA_TRY(w, cudaMemsetAsync(w->d.jointRows, 0, sizeof(JointRow) * (size_t)count, w->stream));
		if (!pairs.empty())
			CUDA_TRY(w, cudaMemcpyAsync((void*)w->d.jointPairKeys, pairs.data(), sizeof(uint64_t) * pairs.size(),
			                            cudaMemcpyHostToDevice, w->stream));
	}
	w->d.jointCount = count;
	w->d.jointPairCount = (int)pairs.size();
	w->jointFilterPending = true;
	int rc = SyncCheck(w);
	if (rc) return rc;
	if ((rc = ZeroCounter(w, CNT_BODY_TYPE_CHANGED))) return rc;
	return ColourJoints(w);
}

int b2cuGetJointCount(b2cuWorld* w, int32_t* count)
{
	if (!w || !count) return B2CU_ERR_ARGUMENT;
	*count = w->d.jointCount;
	return B2CU_OK;
}

int b2cuGetJoints(b2cuWorld* w, int32_t first, int32_t count, b2cuJoint* joints)
{
	int rc = CheckRange(w, first, count, w ? w->d.jointCount : 0, joints);
	if (rc) return rc;
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	CUDA_TRY(w, cudaMemcpyAsync(joints, w->d.joints + first, sizeof(b2cuJoint) * (size_t)count, cudaMemcpyDeviceToHost,
	                            w->stream));
	return SyncCheck(w);
}

int b2cuGetJointOrder(b2cuWorld* w, int32_t capacity, int32_t* jointIds, int32_t* count)
{
	if (!w || capacity < 0 || (capacity > 0 && !jointIds)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	if (w->d.jointCount > 0 && w->jointColourDirty)
	{
		int rc = ColourJoints(w);
		if (rc) return rc;
	}
	if (count) *count = w->d.jointCount;
	int m = std::min(capacity, w->d.jointCount);
	if (m <= 0) return B2CU_OK;
	CUDA_TRY(w, cudaMemcpyAsync(jointIds, w->d.jointOrder, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, w->stream));
	return SyncCheck(w);
}

int b2cuSetContacts(b2cuWorld* w, int32_t count, const b2cuContact* contacts)
{
	if (!w || count < 0 || (count > 0 && !contacts)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	if (count > w->contactCapacity)
	{
		int rc = Reserve(w, w->bodyCapacity, w->proxyCapacity, w->shapeCapacity, std::max(count, 2 * w->contactCapacity));
		if (rc) return rc;
	}
	std::vector<int> order(count);
	std::vector<uint64_t> keys(count);
	for (int i = 0; i < count; ++i)
	{
		order[i] = i;
		const b2cuContact& c = contacts[i];
		if (c.proxyA < 0 || c.proxyA >= w->proxyCount || c.proxyB < 0 || c.proxyB >= w->proxyCount || c.proxyA == c.proxyB)
			return SetError(w, B2CU_ERR_ARGUMENT, "contact %d: proxies %d,%d out of range", i, c.proxyA, c.proxyB);
		uint32_t lo = (uint32_t)std::min(c.proxyA, c.proxyB), hi = (uint32_t)std::max(c.proxyA, c.proxyB);
		keys[i] = ((uint64_t)lo << 32) | hi;
	}
	std::sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] < keys[b]; });
	std::vector<uint64_t> key(count);
	std::vector<int4> proxies(count);
	std::vector<uint32_t> flags(count);
	std::vector<float4> m0(count), m1(count), m2(count), mix(count);
	std::vector<uint4> m3(count);
	std::vector<int> toiCount(count), colour(count, B2CU_COLOUR_NONE);
	std::vector<uint32_t> stamp(count);
	uint32_t maxStamp = 0;
	for (int j = 0; j < count; ++j)
	{
		const b2cuContact& c = contacts[order[j]];
		stamp[j] = c.stamp;
		maxStamp = std::max(maxStamp, c.stamp);
		const b2cuManifold& m = c.manifold;
		key[j] = keys[order[j]];
		if (j > 0 && key[j] == key[j - 1]) return SetError(w, B2CU_ERR_ARGUMENT, "duplicate contact key");
		proxies[j] = make_int4(c.proxyA, c.proxyB, 0, 0);
		flags[j] = c.flags & 0xFFu & ~(uint32_t)B2CU_CONTACT_ISLAND; // the bits above are the device's own
		m0[j] = make_float4(m.localNormal[0], m.localNormal[1], m.localPoint[0], m.localPoint[1]);
		m1[j] = make_float4(m.points[0].localPoint[0], m.points[0].localPoint[1], m.points[0].normalImpulse,
		                    m.points[0].tangentImpulse);
		m2[j] = make_float4(m.points[1].localPoint[0], m.points[1].localPoint[1], m.points[1].normalImpulse,
		                    m.points[1].tangentImpulse);
		m3[j] = make_uint4(m.id[0], m.id[1], (uint32_t)m.type, (uint32_t)m.pointCount);
		mix[j] = make_float4(c.friction, c.restitution, c.tangentSpeed, c.toi);
		toiCount[j] = c.toiCount;
	}
	DeviceArrays& d = w->d;
	int rc;
	if ((rc = Upload(w, d.c.key, 0, key)) || (rc = Upload(w, d.c.proxies, 0, proxies)) ||
	    (rc = Upload(w, d.c.flags, 0, flags)) || (rc = Upload(w, d.c.m0, 0, m0)) || (rc = Upload(w, d.c.m1, 0, m1)) ||
	    (rc = Upload(w, d.c.m2, 0, m2)) || (rc = Upload(w, d.c.m3, 0, m3)) || (rc = Upload(w, d.c.mix, 0, mix)) ||
	    (rc = Upload(w, d.c.toiCount, 0, toiCount)) || (rc = Upload(w, d.c.colour, 0, colour)) ||
	    (rc = Upload(w, d.c.stamp, 0, stamp)))
		return rc;
	// contacts created from now on are younger than all of these
	w->contactBatch = maxStamp + 1u;
	w->contactCount = count;
	w->mainCount = count;
	w->deadMain = 0;
	if (count > 0) LAUNCH(w, FillContactBodiesKernel, GridFor(count), kBlock, d, count);
	if ((rc = RebuildLowStart(w))) return rc;
	return SyncCheck(w);
}

int b2cuGetContactCount(b2cuWorld* w, int32_t* count)
{
	if (!w || !count) return B2CU_ERR_ARGUMENT;
	*count = w->contactCount - w->deadMain;
	return B2CU_OK;
}

int b2cuGetContacts(b2cuWorld* w, int32_t capacity, b2cuContact* contacts, int32_t* countOut)
{
	if (!w || capacity < 0 || (capacity > 0 && !contacts)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	const int slots = w->contactCount;
	const int live = slots - w->deadMain;
	if (countOut) *countOut = live;
	if (capacity == 0 || slots == 0) return B2CU_OK;
	std::vector<uint64_t> key(slots);
	std::vector<int4> proxies(slots);
	std::vector<uint32_t> flags(slots);
	std::vector<float4> m0(slots), m1(slots), m2(slots), mix(slots);
	std::vector<uint4> m3(slots);
	std::vector<int> toiCount(slots);
	std::vector<uint32_t> stamp(slots);
	std::vector<int> pbody(w->proxyCount);
	std::vector<uint32_t> bflags(w->bodyCount);
	DeviceArrays& d = w->d;
	int rc;
	if ((rc = Download(w, d.pbody, 0, pbody)) || (rc = Download(w, d.bflags, 0, bflags))) return rc;
	if ((rc = Download(w, d.c.key, 0, key)) || (rc = Download(w, d.c.proxies, 0, proxies)) ||
	    (rc = Download(w, d.c.flags, 0, flags)) || (rc = Download(w, d.c.m0, 0, m0)) || (rc = Download(w, d.c.m1, 0, m1)) ||
	    (rc = Download(w, d.c.m2, 0, m2)) || (rc = Download(w, d.c.m3, 0, m3)) || (rc = Download(w, d.c.mix, 0, mix)) ||
	    (rc = Download(w, d.c.toiCount, 0, toiCount)) || (rc = Download(w, d.c.stamp, 0, stamp)))
		return rc;
	if ((rc = SyncCheck(w))) return rc;
	// live slots in key order: the main region and the tail are each sorted; this read-out does not touch the
	// device set (compacting here would change contact indices and with them the colouring of later steps)
	std::vector<int> order;
	order.reserve(live);
	for (int i = 0; i < slots; ++i)
		if (!(flags[i] & B2CU_CONTACT_DEAD)) order.push_back(i);
	size_t mid = 0;
	while (mid < order.size() && order[mid] < w->mainCount) ++mid;
	std::inplace_merge(order.begin(), order.begin() + mid, order.end(), [&](int a, int b) { return key[a] < key[b]; });
	const int count = std::min(capacity, (int)order.size());
	for (int out = 0; out < count; ++out)
	{
		const int j = order[out];
		b2cuContact& c = contacts[out];
		c.proxyA = proxies[j].x;
		c.proxyB = proxies[j].y;
		// e_inactiveFlag is not stored on the device: it is a function of the bodies' awake state
		// (b2ContactManager::IsContactActive, b2ContactManager.cpp:94-107)
		{
			uint32_t fA = bflags[pbody[proxies[j].x]], fB = bflags[pbody[proxies[j].y]];
			bool activeA = (fA & B2CU_BODY_AWAKE) && (fA & B2CU_BODY_TYPE_MASK) != B2CU_STATIC_BODY;
			bool activeB = (fB & B2CU_BODY_AWAKE) && (fB & B2CU_BODY_TYPE_MASK) != B2CU_STATIC_BODY;
			uint32_t f = flags[j] & ~(uint32_t)(B2CU_CONTACT_INACTIVE | B2CU_CONTACT_DEAD | B2CU_CONTACT_SENSOR);
			if (!activeA && !activeB) f |= B2CU_CONTACT_INACTIVE;
			c.flags = f;
		}
		c.friction = mix[j].x;
		c.restitution = mix[j].y;
		c.tangentSpeed = mix[j].z;
		c.toi = mix[j].w;
		c.toiCount = toiCount[j];
		c.stamp = stamp[j];
		c.reserved = 0u;
		b2cuManifold& m = c.manifold;
		m.localNormal[0] = m0[j].x; m.localNormal[1] = m0[j].y;
		m.localPoint[0] = m0[j].z; m.localPoint[1] = m0[j].w;
		m.points[0].localPoint[0] = m1[j].x; m.points[0].localPoint[1] = m1[j].y;
		m.points[0].normalImpulse = m1[j].z; m.points[0].tangentImpulse = m1[j].w;
		m.points[1].localPoint[0] = m2[j].x; m.points[1].localPoint[1] = m2[j].y;
		m.points[1].normalImpulse = m2[j].z; m.points[1].tangentImpulse = m2[j].w;
		m.id[0] = m3[j].x; m.id[1] = m3[j].y;
		m.type = (int32_t)m3[j].z;
		m.pointCount = (int32_t)m3[j].w;
	}
	return B2CU_OK;
}

// ---------------------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------------------
// ---- body mirror: the step's own device -> host copy of the body records ----
static int StartMirrorCopy(b2cuWorld* w)
{
	const int n = std::min(w->bodyMirrorCount, w->bodyCount);
	if (w->bodyMirror == nullptr || n <= 0) return B2CU_OK;
	int rc = EnsureBodyStage(w);
	if (rc) return rc;
	CUDA_TRY(w, cudaEventRecord(w->evBodiesFinal, w->stream));
	CUDA_TRY(w, cudaStreamWaitEvent(w->copyStream, w->evBodiesFinal, 0));
	PackBodyStatesKernel<<<GridFor(n), kBlock, 0, w->copyStream>>>(w->d, 0, n, w->bodyStage);
	++w->launches;
	CUDA_TRY(w, cudaEventRecord(w->evPacked, w->copyStream));
	CUDA_TRY(w, cudaMemcpyAsync(w->bodyMirror, w->bodyStage, sizeof(b2cuBodyState) * (size_t)n, cudaMemcpyDeviceToHost,
	                            w->copyStream));
	w->mirrorInFlight = true;
	return B2CU_OK;
}

static int FinishMirrorCopy(b2cuWorld* w)
{
	if (!w->mirrorInFlight) return B2CU_OK;
	w->mirrorInFlight = false;
	CUDA_TRY(w, cudaStreamSynchronize(w->copyStream));
	// bodies woken by the contacts created after the copy had started (the new-contact path of b2ContactManager wakes both bodies,
	// b2ContactManager.cpp:520-529): same change on the caller's side
	const int patches = w->hostCounters[CNT_WAKE_PATCH];
	if (patches > 0)
	{
		if (patches > w->hostPatchCapacity)
		{
			if (w->hostPatch) cudaFreeHost(w->hostPatch);
			w->hostPatch = nullptr;
			w->hostPatchCapacity = 0;
			CUDA_TRY(w, cudaMallocHost(&w->hostPatch, sizeof(int) * (size_t)w->bodyCapacity));
			w->hostPatchCapacity = w->bodyCapacity;
		}
		CUDA_TRY(w, cudaMemcpyAsync(w->hostPatch, w->d.wakePatch, sizeof(int) * (size_t)patches, cudaMemcpyDeviceToHost,
		                            w->stream));
		CUDA_TRY(w, cudaStreamSynchronize(w->stream));
		const int n = std::min(w->bodyMirrorCount, w->bodyCount);
		for (int k = 0; k < patches; ++k)
		{
			int b = w->hostPatch[k];
			if (b < 0 || b >= n) continue;
			w->bodyMirror[b].flags |= B2CU_BODY_AWAKE;
			w->bodyMirror[b].sleepTime = 0.0f;
		}
	}
	return B2CU_OK;
}

static int EnsureQueryScratch(b2cuWorld* w, size_t bytes);
static int IssueEventRecords(b2cuWorld* w, bool* issued);

int b2cuSetPreSolveHook(b2cuWorld* w, b2cuPreSolveFn fn, void* user)
{
	if (!w) return B2CU_ERR_ARGUMENT;
	w->preSolveHook = fn;
	w->preSolveUser = user;
	return B2CU_OK;
}

int b2cuGetPreSolveContacts(b2cuWorld* w, int32_t capacity, b2cuContact* records, b2cuManifold* oldManifolds, int32_t* count)
{
	if (!w || capacity < 0 || (capacity > 0 && (!records || !oldManifolds))) return B2CU_ERR_ARGUMENT;
	if (!w->inPreSolve) return SetError(w, B2CU_ERR_ARGUMENT, "b2cuGetPreSolveContacts is only valid inside the pre-solve hook");
	cudaSetDevice(w->device);
	DeviceArrays& d = w->d;
	const int nc = w->contactCount;
	int rc;
	if (w->preSolveCount < 0)
	{
		int n = 0;
		if (nc > 0)
		{
			LAUNCH(w, PreSolveSelectKernel, GridFor(nc), kBlock, d, nc);
			CompactFlags(&w->prims, d.cSelect, nc, d.listA, d.counters + CNT_SCRATCH, w->stream);
			CUDA_TRY(w, cudaMemcpyAsync(&w->hostCounters[CNT_SCRATCH], d.counters + CNT_SCRATCH, sizeof(int),
			                            cudaMemcpyDeviceToHost, w->stream));
			if ((rc = SyncCheck(w))) return rc;
			n = w->hostCounters[CNT_SCRATCH];
		}
		w->preSolveCount = n;
	}
	const int n = w->preSolveCount;
	if (count) *count = n;
	const int m = std::min(n, capacity);
	if (m <= 0) return B2CU_OK;
	const size_t recBytes = (sizeof(b2cuContact) * (size_t)m + 255) & ~(size_t)255;
	const size_t total = recBytes + sizeof(b2cuManifold) * (size_t)m;
	if ((rc = EnsureQueryScratch(w, total))) return rc;
	b2cuContact* dRec = reinterpret_cast<b2cuContact*>(w->queryScratch);
	b2cuManifold* dOld = reinterpret_cast<b2cuManifold*>(static_cast<char*>(w->queryScratch) + recBytes);
	LAUNCH(w, PreSolveGatherKernel, GridFor(m), kBlock, d, (const int*)d.listA, m, dRec, dOld);
	CUDA_TRY(w, cudaMemcpyAsync(records, dRec, sizeof(b2cuContact) * (size_t)m, cudaMemcpyDeviceToHost, w->stream));
	CUDA_TRY(w, cudaMemcpyAsync(oldManifolds, dOld, sizeof(b2cuManifold) * (size_t)m, cudaMemcpyDeviceToHost, w->stream));
	return SyncCheck(w);
}

int b2cuDisableContacts(b2cuWorld* w, int32_t count, const b2cuContactKey* keys)
{
	if (!w || count < 0 || (count > 0 && !keys)) return B2CU_ERR_ARGUMENT;
	if (!w->inPreSolve) return SetError(w, B2CU_ERR_ARGUMENT, "b2cuDisableContacts is only valid inside the pre-solve hook");
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	int rc;
	// the key list goes behind whatever the gather call left in the scratch (it has been copied out already)
	if ((rc = EnsureQueryScratch(w, sizeof(uint64_t) * (size_t)count))) return rc;
	uint64_t* dKeys = reinterpret_cast<uint64_t*>(w->queryScratch);
	CUDA_TRY(w, cudaMemcpyAsync(dKeys, keys, sizeof(uint64_t) * (size_t)count, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(w, DisableContactsKernel, GridFor(count), kBlock, w->d, w->contactCount, w->mainCount, (const uint64_t*)dKeys, count);
	return SyncCheck(w);
}

int b2cuSetPairFilter(b2cuWorld* w, b2cuPairFilterFn fn, void* user)
{
	if (!w) return B2CU_ERR_ARGUMENT;
	w->pairFilter = fn;
	w->pairFilterUser = user;
	w->d.customFilter = fn != nullptr ? 1 : 0;
	return B2CU_OK;
}

int b2cuGetBodyStates(b2cuWorld* w, int32_t first, int32_t count, b2cuBodyState* states)
{
	int rc = CheckRange(w, first, count, w ? w->bodyCount : 0, states);
	if (rc) return rc;
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	if ((rc = EnsureBodyStage(w))) return rc;
	float* stage = w->bodyStage + (size_t)first * B2CU_STATE_WORDS;
	LAUNCH(w, PackBodyStatesKernel, GridFor(count), kBlock, w->d, first, count, stage);
	CUDA_TRY(w, cudaMemcpyAsync(states, stage, sizeof(b2cuBodyState) * (size_t)count, cudaMemcpyDeviceToHost, w->stream));
	return SyncCheck(w);
}

int b2cuGetBodySweepStarts(b2cuWorld* w, int32_t first, int32_t count, b2cuSweepStart* starts)
{
	int rc = CheckRange(w, first, count, w ? w->bodyCount : 0, starts);
	if (rc) return rc;
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	// the column is stored as (c0.x, c0.y, a0, alpha0) rows already
	static_assert(sizeof(b2cuSweepStart) == sizeof(float4), "b2cuSweepStart layout");
	CUDA_TRY(w, cudaMemcpyAsync(starts, w->d.pos0 + first, sizeof(float4) * (size_t)count, cudaMemcpyDeviceToHost, w->stream));
	return SyncCheck(w);
}

int b2cuSetEventPrefetch(b2cuWorld* w, int32_t on)
{
	if (!w) return B2CU_ERR_ARGUMENT;
	w->eventPrefetch = on != 0;
	return B2CU_OK;
}

int b2cuSetBodyForces(b2cuWorld* w, int32_t first, int32_t count, const float* forces)
{
	int rc = CheckRange(w, first, count, w ? w->bodyCount : 0, forces);
	if (rc) return rc;
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	if ((rc = EnsureBodyStage(w))) return rc;
	float* stage = w->bodyStage + (size_t)first * 3;
	CUDA_TRY(w, cudaMemcpyAsync(stage, forces, sizeof(float) * 3 * (size_t)count, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(w, SetBodyForcesKernel, GridFor(count), kBlock, w->d, first, count, (const float*)stage);
	return SyncCheck(w);
}

int b2cuSetBodyMirror(b2cuWorld* w, b2cuBodyState* mirror, int32_t count)
{
	if (!w || count < 0 || (count > 0 && !mirror)) return B2CU_ERR_ARGUMENT;
	w->bodyMirror = count > 0 ? mirror : nullptr;
	w->bodyMirrorCount = count;
	return B2CU_OK;
}

int b2cuStep(b2cuWorld* w, float dt, int32_t velocityIterations, int32_t positionIterations, b2cuStepInfo* info)
{
	if (!w || velocityIterations < 0 || positionIterations < 0) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	DeviceArrays& d = w->d;
	int rc;
	w->launches = 0;
	g_primLaunches = 0;
	b2cuStepInfo out;
	memset(&out, 0, sizeof(out));
	const int nb = w->bodyCount;
	const int np = w->proxyCount;

	if (positionIterations > w->positionIterationsCapacity)
	{
		w->positionIterationsCapacity = positionIterations;
		cudaFree(d.islandMinSep);
		CUDA_TRY(w, cudaMalloc(&d.islandMinSep, sizeof(int) * (size_t)positionIterations * (size_t)w->bodyCapacity));
	}

	CUDA_TRY(w, cudaMemsetAsync(d.counters, 0, sizeof(int) * CNT_STICKY_TOI, w->stream));
	w->mirrorInFlight = false;
	w->eventCacheValid = false;
	cudaEventRecord(w->ev[0], w->stream);
	{
		const char* t = getenv("B2CU_TRACE");
		g_trace.enabled = t && atoi(t) > 0;
	}
	g_trace.used = 0;
	g_traceWorld = w;
	g_primTraceHook = g_trace.enabled ? PrimTraceHook : nullptr;
	TraceMark(w, "(start)");

	if (w->d.jointCount > 0 && w->jointColourDirty && (rc = ColourJoints(w))) return rc;

	if (w->toiCheckDirty)
	{
		// can any contact be a TOI candidate at all (b2Contact::IsToiCandidate)?  Worlds whose static geometry is
		// all thick-shape and that have no bullets skip the TOI eligibility pass entirely.
		CUDA_TRY(w, cudaMemsetAsync(d.counters + CNT_STICKY_TOI, 0, sizeof(int), w->stream));
		LAUNCH(w, ToiPossibleKernel, GridFor(std::max(nb, np)), kBlock, d, nb, np);
		w->toiCheckDirty = false;
	}

	if (w->refilterPending)
	{
		if (w->contactCount > 0) LAUNCH(w, FlagFilterContactsKernel, GridFor(w->contactCount), kBlock, d, w->contactCount);
		if (np > 0) LAUNCH(w, ClearProxyFlagKernel, GridFor(np), kBlock, d, np, (uint32_t)B2CU_PROXY_REFILTER);
		w->refilterPending = false;
	}
	if (w->contactBodiesDirty)
	{
		if (np > 0) LAUNCH(w, FillProxyRadiusKernel, GridFor(np), kBlock, d, np);
		if (w->contactCount > 0) LAUNCH(w, FillContactBodiesKernel, GridFor(w->contactCount), kBlock, d, w->contactCount);
		w->contactBodiesDirty = false;
	}
	if (w->jointFilterPending)
	{
		// b2World::CreateJoint (b2World.cpp:710-726): contacts between bodies the new joints keep apart are re-filtered
		if (w->contactCount > 0 && d.jointPairCount > 0)
			LAUNCH(w, FlagJointContactsKernel, GridFor(w->contactCount), kBlock, d, w->contactCount);
		w->jointFilterPending = false;
	}

	if (dt > 0.0f && (rc = ShardSyncGhosts(w))) return rc;

	int newContacts = 0, destroyed = 0, moved = 0;

	// ---- new fixtures: find their contacts first (b2World.cpp:1628-1639) ----
	if (w->newProxies)
	{
		int n0 = 0, d0 = 0, m0 = 0;
		if ((rc = FindNewContactsAndRebuild(w, &n0, &d0, &m0))) return rc;
		newContacts += n0;
		moved += m0;
		w->newProxies = false;
	}
	cudaEventRecord(w->ev[1], w->stream);

	// ---- Collide (b2World.cpp:1120-1141) ----
	int nc = w->contactCount;
	w->beginCount = w->endCount = 0;
	if (nc > 0)
	{
		// narrow phase; begin/end events are appended to the deferred buffers and sorted at the end of the step
		if (w->preSolveHook != nullptr)
		{
			// the manifolds of the previous step, for PreSolve (b2Contact::Update keeps `oldManifold` the same way)
			CUDA_TRY(w, cudaMemcpyAsync(d.cAlt.m0, d.c.m0, sizeof(float4) * (size_t)nc, cudaMemcpyDeviceToDevice, w->stream));
			CUDA_TRY(w, cudaMemcpyAsync(d.cAlt.m1, d.c.m1, sizeof(float4) * (size_t)nc, cudaMemcpyDeviceToDevice, w->stream));
			CUDA_TRY(w, cudaMemcpyAsync(d.cAlt.m2, d.c.m2, sizeof(float4) * (size_t)nc, cudaMemcpyDeviceToDevice, w->stream));
			CUDA_TRY(w, cudaMemcpyAsync(d.cAlt.m3, d.c.m3, sizeof(uint4) * (size_t)nc, cudaMemcpyDeviceToDevice, w->stream));
		}
		LAUNCH(w, CollideKernel, GridFor(nc), kBlock, d, nc, w->mainCount, w->contactCapacity, d.listA);
		LAUNCH(w, CollideHeavyKernel, GridFor(nc), kBlock, d, (const int*)d.listA, w->contactCapacity);
		LAUNCH(w, ApplyWakeKernel, GridFor(nb), kBlock, d, nb, (int*)nullptr);
		if (w->preSolveHook != nullptr)
		{
			// FinishCollide's PreSolve calls: between the narrow phase and the solver, on the calling thread
			if ((rc = SyncCheck(w))) return rc;
			w->inPreSolve = true;
			w->preSolveCount = -1;
			int hookRc = w->preSolveHook(w->preSolveUser, w);
			w->inPreSolve = false;
			if (hookRc != 0) return SetError(w, B2CU_ERR_ARGUMENT, "the pre-solve hook failed (%d)", hookRc);
		}
	}
	cudaEventRecord(w->ev[2], w->stream);

	const bool allowSleep = (w->params.flags & B2CU_WORLD_ALLOW_SLEEP) != 0;
	const bool warmStarting = (w->params.flags & B2CU_WORLD_WARM_STARTING) != 0;
	const float invDt = dt > 0.0f ? 1.0f / dt : 0.0f;
	const float dtRatio = w->params.invDt0 * dt;

	w->constraintCount = 0;
	w->colourCount = 0;
	w->overflowCount = 0;
	for (int c = 0; c < B2CU_MAX_COLOURS + 2; ++c) w->colourCounts[c] = 0;

	// b2World.cpp:1670: a sub-stepping world that stopped between two time-of-impact events does not solve again
	if (dt > 0.0f && w->stepComplete)
	{
		// ---- islands (serial DFS of b2World::Solve -> union-find) ----
		LAUNCH(w, SolveInitBodiesKernel, GridFor(nb), kBlock, d, nb, positionIterations);
		if (nc > 0)
		{
			static const int sample = []() {
				const char* e = getenv("B2CU_UNION_SAMPLE");
				return e && atoi(e) != 0 ? atoi(e) : 2;
			}();
			if (sample > 1 && nc >= 4096)
			{
				LAUNCH(w, IslandUnionKernel, GridFor(nc), kBlock, d, nc, sample, 0);
				LAUNCH(w, IslandCompressKernel, GridFor(nb), kBlock, d, nb);
				LAUNCH(w, IslandUnionKernel, GridFor(nc), kBlock, d, nc, sample, 1);
			}
			else
			{
				LAUNCH(w, IslandUnionKernel, GridFor(nc), kBlock, d, nc, sample < 0 ? -1 : 1, 0);
			}
		}
		if (w->d.jointCount > 0) LAUNCH(w, JointUnionKernel, GridFor(w->d.jointCount), kBlock, d, w->d.jointCount);
		LAUNCH(w, IslandFlattenKernel, GridFor(nb), kBlock, d, nb);
		LAUNCH(w, IslandMarkKernel, GridFor(nb), kBlock, d, nb);
		cudaEventRecord(w->ev[3], w->stream);

		// ---- constraint selection + colouring ----
		int* colourStart = w->colourStarts;
		const int crossBase = w->shardCount > 1 ? 16 : 32;
		int nConstraints = 0;
		if (nc > 0)
		{
			LAUNCH(w, SelectConstraintsKernel, GridFor(nc), kBlock, d, nc);
			CompactFlags(&w->prims, d.cSelect, nc, d.listA, d.counters + CNT_CONSTRAINT, w->stream);
			CUDA_TRY(w, cudaMemsetAsync(d.colourCount, 0, sizeof(int) * (B2CU_MAX_COLOURS + 2), w->stream));
			int* cur = d.listB;
			int* next = d.listC;
			LAUNCH(w, ColourPrepareKernel, GridFor(nc), kBlock, d, d.listA, cur, crossBase);
			// a fixed number of colouring rounds is queued without asking the device how many constraints are left
			// (in steady state only the few new constraints are uncoloured and 2-3 rounds finish them); one
			// read-back then gives the constraint count, the per-colour counts and what is left
			const int kBlindRounds = 4;
			uint32_t round = 1;
			int counter = CNT_UNCOLOURED;
			for (int r = 0; r < kBlindRounds; ++r)
			{
				int g = GridFor(std::max(1024, nc / 64));
				LAUNCH(w, ColourProposeKernel, g, kBlock, d, cur, counter, round, crossBase);
				LAUNCH(w, ColourCommitKernel, g, kBlock, d, cur, counter, next, counter + 1, round, crossBase);
				std::swap(cur, next);
				++counter;
				++round;
			}
			if ((rc = ReadCounters(w))) return rc;
			nConstraints = w->hostCounters[CNT_CONSTRAINT];
			int remaining = w->hostCounters[counter];
			while (remaining > 0)
			{
				// rare: more rounds, now one read-back per round; the two spare counters are ping-ponged
				int nextCounter = counter == CNT_UNCOLOURED_LAST ? CNT_UNCOLOURED_LAST - 1 : counter + 1;
				if ((rc = ZeroCounter(w, nextCounter))) return rc;
				LAUNCH(w, ColourProposeKernel, GridFor(remaining), kBlock, d, cur, counter, round, crossBase);
				LAUNCH(w, ColourCommitKernel, GridFor(remaining), kBlock, d, cur, counter, next, nextCounter, round, crossBase);
				if ((rc = ReadCounters(w))) return rc;
				remaining = w->hostCounters[nextCounter];
				std::swap(cur, next);
				counter = nextCounter;
				++round;
				if (round > 100000u) return SetError(w, B2CU_ERR_CUDA, "colouring did not converge");
			}
			if (nConstraints > 0)
			{
				int at = 0;
				for (int c = 0; c < B2CU_MAX_COLOURS + 2; ++c)
				{
					w->colourCounts[c] = w->hostCounters[CNT_COUNT + c];
					colourStart[c] = at;
					at += w->colourCounts[c];
					if (c < B2CU_MAX_COLOURS && w->colourCounts[c] > 0) ++w->colourCount;
				}
				colourStart[B2CU_MAX_COLOURS + 2] = at;
				if (at != nConstraints)
					return SetError(w, B2CU_ERR_CUDA, "internal: colour counts %d != constraints %d", at, nConstraints);
				w->overflowCount = w->colourCounts[B2CU_MAX_COLOURS] + w->colourCounts[B2CU_MAX_COLOURS + 1];
				LAUNCH(w, ColourKeysKernel, GridFor(nConstraints), kBlock, d, d.listA);
				RadixSort64(&w->prims, d.orderKeys, nConstraints, 32, 40, w->stream); // (colour << 2 | class) < 256
				CUDA_TRY(w, cudaMemsetAsync(d.colourTwoStart, 0x7F, sizeof(int) * (B2CU_MAX_COLOURS + 2), w->stream));
			}
		}
		w->constraintCount = nConstraints;
		cudaEventRecord(w->ev[4], w->stream);

		// ---- b2Island::Solve ----
		w->flowBase = (int)((w->flowEpoch++ & 0xFFFu) << 19);
		LAUNCH(w, IntegrateVelocitiesKernel, GridFor(nb), kBlock, d, nb, dt, w->params.gravity, w->flowBase);
		if (nConstraints > 0)
		{
			LAUNCH(w, ConstraintSlotKernel, GridFor(nConstraints), kBlock, d);
			LAUNCH(w, InitConstraintsKernel, GridFor(nConstraints), kBlock, d, (const int*)d.listA, dtRatio, warmStarting ? 1 : 0);
			if (warmStarting && !w->persistentSolver)
			{
				for (int c = 0; c < B2CU_MAX_COLOURS; ++c)
				{
					if (w->colourCounts[c] > 0)
						LAUNCH(w, WarmStartKernel, GridFor(w->colourCounts[c]), kBlock, d, colourStart[c], w->colourCounts[c]);
				}
				if (w->colourCounts[B2CU_MAX_COLOURS] > 0)
					LAUNCH(w, OverflowWarmStartKernel, 1, 32, d, colourStart[B2CU_MAX_COLOURS], w->colourCounts[B2CU_MAX_COLOURS]);
			}
		}
		cudaEventRecord(w->ev[5], w->stream);
		const int nJoints = w->d.jointCount;
		if (nJoints > 0 && (!w->persistentSolver || w->persistentGridJoints <= 0 || w->persistentGridPositionJoints <= 0))
			return SetError(w, B2CU_ERR_UNSUPPORTED, "joints need the persistent cooperative solver");
		if ((nConstraints > 0 || w->shardCount > 1 || nJoints > 0) && w->persistentSolver)
		{
			// one persistent cooperative kernel for warm start + velocity + store + integrate + position; in a
			// sharded world it also carries the halo exchanges, so it runs even without constraints
			SolverPlan plan;
			memset(&plan, 0, sizeof(plan));
			int nOps = 0;
			auto addSegments = [&](int first, int last, int overflowColour) {
				for (int c = first; c < last; ++c)
				{
					if (nConstraints > 0 && w->colourCounts[c] > 0)
					{
						plan.opType[nOps] = OP_PARALLEL;
						plan.opStart[nOps] = colourStart[c];
						plan.opSize[nOps] = w->colourCounts[c];
						plan.opColour[nOps] = c;
						++nOps;
					}
				}
				if (nConstraints > 0 && w->colourCounts[overflowColour] > 0)
				{
					plan.opType[nOps] = OP_SERIAL;
					plan.opStart[nOps] = colourStart[overflowColour];
					plan.opSize[nOps] = w->colourCounts[overflowColour];
					++nOps;
				}
			};
			addSegments(0, crossBase, B2CU_MAX_COLOURS);
			if (w->shardCount > 1) plan.opType[nOps++] = OP_PUSH_DOWN;
			addSegments(crossBase, B2CU_MAX_COLOURS, B2CU_MAX_COLOURS + 1);
			if (w->shardCount > 1) plan.opType[nOps++] = OP_PUSH_UP;
			plan.opCount = nOps;
			plan.constraintCount = nConstraints;
			plan.bodyCount = nb;
			plan.velocityIterations = velocityIterations;
			plan.positionIterations = positionIterations;
			plan.warmStarting = warmStarting ? 1 : 0;
			plan.h = dt;
			plan.shard = MakeShardState(w);
			plan.dtRatio = dtRatio;
			{
				static const int skip = []() { const char* e = getenv("B2CU_DEBUG_SKIP_STORE"); return e ? atoi(e) : 0; }();
				plan.debugSkipStore = skip;
				static const int prefetch = []() { const char* e = getenv("B2CU_FLOW_PREFETCH"); return e ? atoi(e) : 0; }();
				plan.flowPrefetch = prefetch;
			}
			plan.jointOpCount = nJoints > 0 ? w->jointOpCount : 0;
			for (int jo = 0; jo < plan.jointOpCount; ++jo)
			{
				plan.jointOpStart[jo] = w->jointOpStart[jo];
				plan.jointOpSize[jo] = w->jointOpSize[jo];
				plan.jointOpSerial[jo] = w->jointOpSerial[jo];
			}
			void* args[2] = {(void*)&d, (void*)&plan};
			// dataflow instance (b2cu_solver_flow.cuh): no colour barriers.  Needs every constraint coloured (no serial
			// list), no joints, no halo exchange
			static const bool flowEnabled = []() {
				const char* e = getenv("B2CU_FLOW");
				return !(e && atoi(e) == 0);
			}();
			const bool sharded = w->shardCount > 1;
			if (sharded && w->shardFlow && (nJoints > 0 || w->overflowCount > 0))
				return SetError(w, B2CU_ERR_UNSUPPORTED, "a sharded world on the dataflow solver cannot have joints or a body with more "
				                                         "constraints than colours (%d overflow constraints); B2CU_SHARD_FLOW=0 on every "
				                                         "shard selects the barrier kernels", w->overflowCount);
			const bool flow = sharded ? (w->shardFlow && w->flowGridMax > 0 && w->flowGridPositionMax > 0)
			                          : (flowEnabled && nJoints == 0 && nConstraints > 0 && w->flowGrid > 0 && w->flowGridPosition > 0);
			plan.ovStart = plan.ovCount = 0;
			plan.ovRank = nullptr;
			plan.ovDeg = nullptr;
			if (flow && !sharded && w->overflowCount > 0)
			{
				// the overflow list inside the dataflow: rank of every row on its two bodies (its place in the body's chain)
				// and the bodies' extra degree, by sorting (body, row) keys
				plan.ovStart = colourStart[B2CU_MAX_COLOURS];
				plan.ovCount = w->colourCounts[B2CU_MAX_COLOURS];
				plan.ovRank = d.listC;
				plan.ovDeg = d.islandAwake; // free once the islands are marked
				const int nKeys = 2 * plan.ovCount;
				if (nKeys > w->contactCapacity)
					return SetError(w, B2CU_ERR_CAPACITY, "overflow list of %d constraints exceeds the scratch (contact capacity %d)",
					                plan.ovCount, w->contactCapacity);
				CUDA_TRY(w, cudaMemsetAsync(d.islandAwake, 0, sizeof(int) * (size_t)nb, w->stream));
				LAUNCH(w, FlowOverflowKeysKernel, GridFor(plan.ovCount), kBlock, d, plan.ovStart, plan.ovCount, d.toiListKeys);
				{
					int bitsRow = 1, bitsBody = 1;
					while ((1 << bitsRow) < std::max(2, nKeys)) ++bitsRow;
					while ((1 << bitsBody) < std::max(2, nb)) ++bitsBody;
					RadixSort64(&w->prims, d.toiListKeys, nKeys, 0, bitsRow, w->stream);
					RadixSort64(&w->prims, d.toiListKeys, nKeys, 32, 32 + bitsBody + 1, w->stream); // +1: the ~0 keys sort last
				}
				LAUNCH(w, FlowOverflowRanksKernel, GridFor(nKeys), kBlock, (const uint64_t*)d.toiListKeys, nKeys, d.listC, d.islandAwake);
			}
			const bool ov = flow && !sharded && w->overflowCount > 0;
			const void* velocityKernel = flow ? (sharded ? (const void*)SolverVelocityFlowKernel<true, false>
			                                             : ov ? (const void*)SolverVelocityFlowKernel<false, true>
			                                                  : (const void*)SolverVelocityFlowKernel<false, false>)
			                                  : nJoints > 0 ? (const void*)SolverVelocityPersistentKernel<true>
			                                                : (const void*)SolverVelocityPersistentKernel<false>;
			const void* positionKernel = flow ? (sharded ? (const void*)SolverPositionFlowKernel<true, false>
			                                             : ov ? (const void*)SolverPositionFlowKernel<false, true>
			                                                  : (const void*)SolverPositionFlowKernel<false, false>)
			                                  : nJoints > 0 ? (const void*)SolverPositionPersistentKernel<true>
			                                                : (const void*)SolverPositionPersistentKernel<false>;
			plan.flowBase = w->flowBase;
			if (flow && sharded)
			{
				// unite the halo bodies' colour masks of the two sides (b2cu_solver_flow.cuh); rows [0, n) of the mailboxes
				ShardState sh = plan.shard;
				const unsigned seq = w->shardSeq++;
				const int nHalo = std::max(w->ghostCount, w->exportCount);
				if (nHalo > 0) LAUNCH(w, HaloMaskSendKernel, GridFor(nHalo), kBlock, d, sh);
				if (sh.lowerFromUpper != nullptr) LAUNCH(w, ShardSignalKernel, 1, 1, sh.lowerFlagFromUpper, seq);
				if (sh.upperFromLower != nullptr) LAUNCH(w, ShardSignalKernel, 1, 1, sh.upperFlagFromLower, seq);
				if (sh.upperFromLower != nullptr) LAUNCH(w, ShardWaitKernel, 1, 1, sh.flagFromUpper, seq, w->d.counters + CNT_FLOW_STUCK);
				if (sh.lowerFromUpper != nullptr) LAUNCH(w, ShardWaitKernel, 1, 1, sh.flagFromLower, seq, w->d.counters + CNT_FLOW_STUCK);
				if (nHalo > 0) LAUNCH(w, HaloMaskApplyKernel, GridFor(nHalo), kBlock, d, sh);
				plan.shard.seq = w->shardSeq;
			}
			int velocityGrid = nJoints > 0 ? std::min(w->persistentGrid, w->persistentGridJoints) : w->persistentGrid;
			int positionGrid =
			    nJoints > 0 ? std::min(w->persistentGridPosition, w->persistentGridPositionJoints) : w->persistentGridPosition;
			if (flow)
			{
				// no barriers to keep cheap: the more threads, the fewer constraints each takes in sequence
				velocityGrid = sharded ? w->shardFlowGrid : ov ? std::min(w->flowGrid, w->flowGridMax) : w->flowGrid;
				positionGrid = sharded ? w->shardFlowGridPosition : w->flowGridPosition;
			}
			else if (w->shardCount == 1)
			{
				// a world whose largest colour class does not fill the co-resident grid is bound by the grid barrier, and
				// the barrier is cheaper with fewer CTAs: launch only as many as that class can occupy
				int largest = 0;
				for (int op = 0; op < nOps; ++op) largest = std::max(largest, plan.opSize[op]);
				for (int jo = 0; jo < plan.jointOpCount; ++jo) largest = std::max(largest, plan.jointOpSize[jo]);
				static const int minGrid = []() {
					const char* e = getenv("B2CU_MIN_SOLVER_GRID");
					return e && atoi(e) > 0 ? atoi(e) : 0;
				}();
				const int floorGrid = minGrid > 0 ? minGrid : std::max(1, g_smCount); // below one CTA per SM nothing more is gained
				const int want = std::max(floorGrid, (largest + B2CU_SOLVER_THREADS - 1) / B2CU_SOLVER_THREADS);
				velocityGrid = std::min(velocityGrid, want);
				positionGrid = std::min(positionGrid, want);
			}
			// shards that share their device with a neighbour: ordinary launches + the kernels' own grid barrier (GridSync)
			const bool soft = w->shardSoftBarrier && w->shardCount > 1;
			auto launchSolver = [&](const void* kernel, int grid) -> cudaError_t {
				if (!soft) return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(B2CU_SOLVER_THREADS), args, 0, w->stream);
				cudaError_t e = cudaMemsetAsync(w->softBarrierCounter, 0, sizeof(unsigned), w->stream);
				if (e != cudaSuccess) return e;
				return cudaLaunchKernel(kernel, dim3(grid), dim3(B2CU_SOLVER_THREADS), args, 0, w->stream);
			};
			plan.softBarrier = soft ? w->softBarrierCounter : nullptr;
			CUDA_TRY(w, launchSolver(velocityKernel, velocityGrid));
			++w->launches;
			TraceMark(w, flow ? "SolverVelocityFlowKernel" : "SolverVelocityPersistentKernel");
			if (plan.debugSkipStore == 2 && flow) LAUNCH(w, StoreImpulsesKernel, GridFor(nConstraints), kBlock, d);
			if (w->shardCount > 1 && !flow) w->shardSeq += 2u * (unsigned)((warmStarting ? 1 : 0) + velocityIterations);
			plan.shard.seq = w->shardSeq;
			cudaEventRecord(w->ev[6], w->stream);
			if (positionIterations > 0)
			{
				CUDA_TRY(w, launchSolver(positionKernel, positionGrid));
				++w->launches;
				TraceMark(w, flow ? "SolverPositionFlowKernel" : "SolverPositionPersistentKernel");
				if (w->shardCount > 1 && !flow) w->shardSeq += 2u * (unsigned)positionIterations;
			}
		}
		else
		{
		if (nConstraints > 0)
		{
			SetL2Window(w, d.vel, sizeof(float4) * (size_t)nb);
			for (int it = 0; it < velocityIterations; ++it)
			{
				for (int c = 0; c < B2CU_MAX_COLOURS; ++c)
				{
					if (w->colourCounts[c] > 0)
						LAUNCH(w, SolveVelocityKernel, GridFor(w->colourCounts[c]), kBlock, d, colourStart[c],
						       w->colourCounts[c]);
				}
				if (w->colourCounts[B2CU_MAX_COLOURS] > 0)
					LAUNCH(w, OverflowSolveVelocityKernel, 1, 32, d, colourStart[B2CU_MAX_COLOURS], w->colourCounts[B2CU_MAX_COLOURS]);
			}
			LAUNCH(w, StoreImpulsesKernel, GridFor(nConstraints), kBlock, d);
		}
		cudaEventRecord(w->ev[6], w->stream);
		LAUNCH(w, IntegratePositionsKernel, GridFor(nb), kBlock, d, nb, dt);
		if (nConstraints > 0)
		{
			SetL2Window(w, d.pos, sizeof(float4) * (size_t)nb);
			for (int it = 0; it < positionIterations; ++it)
			{
				for (int c = 0; c < B2CU_MAX_COLOURS; ++c)
				{
					if (w->colourCounts[c] > 0)
						LAUNCH(w, SolvePositionKernel, GridFor(w->colourCounts[c]), kBlock, d, colourStart[c],
						       w->colourCounts[c], it, nb);
				}
				if (w->colourCounts[B2CU_MAX_COLOURS] > 0)
					LAUNCH(w, OverflowSolvePositionKernel, 1, 32, d, colourStart[B2CU_MAX_COLOURS], w->colourCounts[B2CU_MAX_COLOURS], it, nb);
			}
		}
		}
		if (nConstraints > 0) SetL2Window(w, nullptr, 0);
		LAUNCH(w, FinalizeBodiesKernel, GridFor(nb), kBlock, d, nb, dt, allowSleep ? 1 : 0);
		if (allowSleep) LAUNCH(w, SleepIslandsKernel, GridFor(nb), kBlock, d, nb, positionIterations);
		cudaEventRecord(w->ev[7], w->stream);

		// ---- SynchronizeFixtures + FindNewContacts (b2World.cpp:1410-1427) ----
		if (np > 0) LAUNCH(w, SyncProxiesKernel, GridFor(np), kBlock, d, np);
		// ClearPostSolve + ClearForces (b2World.cpp:1430, :1688-1691); done here so that the read-back of the
		// broad-phase also carries the awake-body count
		LAUNCH(w, EndStepBodiesKernel, GridFor(nb), kBlock, d, nb, (w->params.flags & B2CU_WORLD_CLEAR_FORCES) ? 1 : 0);
		if ((rc = StartMirrorCopy(w))) return rc;
		int n1 = 0, d1 = 0, m1 = 0;
		if ((rc = FindNewContactsAndRebuild(w, &n1, &d1, &m1))) return rc;
		newContacts += n1;
		destroyed += d1;
		moved += m1;
		if (w->hostCounters[CNT_FLOW_STUCK])
			return SetError(w, B2CU_ERR_CUDA,
			                "a bounded wait inside a solver kernel gave up (code %d: 1 = dependency wait of the dataflow solver, 2 = halo "
			                "push of the neighbouring shard, 3 = grid barrier of a co-resident shard); the step is invalid",
			                w->hostCounters[CNT_FLOW_STUCK]);
	}
	else
	{
		for (int k = 3; k < 8; ++k) cudaEventRecord(w->ev[k], w->stream);
		LAUNCH(w, EndStepBodiesKernel, GridFor(nb), kBlock, d, nb, (w->params.flags & B2CU_WORLD_CLEAR_FORCES) ? 1 : 0);
		if ((rc = StartMirrorCopy(w))) return rc;
		int n1 = 0, d1 = 0, m1 = 0;
		if ((rc = FindNewContactsAndRebuild(w, &n1, &d1, &m1))) return rc;
		destroyed += d1;
	}
	if (dt > 0.0f) w->params.invDt0 = invDt;
	cudaEvent_t evBroad = w->ev[8];
	cudaEventRecord(evBroad, w->stream);

	// ---- deferred event buffers: counts only; b2cuGetEvents puts them in callback order.  hostCounters is current:
	// the broad-phase has just read it.
	{
		int nBegin = std::min(w->hostCounters[CNT_BEGIN], w->contactCapacity);
		int nEnd = std::min(w->hostCounters[CNT_END], w->contactCapacity);
		int nDestroyEnd = std::min(w->hostCounters[CNT_DESTROY_END], w->contactCapacity);
		w->beginCount = nBegin;
		w->endUpdateCount = nEnd;
		w->endCount = nEnd + nDestroyEnd;
	}

	// ---- continuous collision: b2World::SolveTOI (b2World.cpp:1677-1682) ----
	if ((w->params.flags & B2CU_WORLD_CONTINUOUS) && dt > 0.0f)
	{
		if ((rc = SolveTOI(w, dt, velocityIterations))) return rc;
	}
	else
	{
		w->toiCount = 0;
		w->toiMinAlpha = 1.0f;
		w->toiMinKey = ~0ull;
		w->toiEventPending = 0;
		w->toiSubSteps = w->toiEventCount = w->toiNewContacts = 0;
	}
	cudaEvent_t evEnd = w->ev[9];
	cudaEventRecord(evEnd, w->stream);

	// a listener will want the contacts of the step's events: gather them and start their way to the host now, behind
	// the step's own work, instead of in a round trip of its own afterwards (events of time-of-impact sub-steps refer
	// to contacts as they are at the very end and are fetched on request)
	w->eventCachePending = false;
	if (w->eventPrefetch && (w->beginCount > 0 || w->endCount > 0))
	{
		if ((rc = IssueEventRecords(w, &w->eventCachePending))) return rc;
	}

	if (w->mirrorInFlight)
	{
		if ((rc = ReadCounters(w))) return rc;
	}
	else
	{
		if ((rc = SyncCheck(w))) return rc;
	}
	if (w->eventCachePending)
	{
		// the step stream has been synchronised above: the records are on the host
		w->eventCacheValid = true;
		w->eventCachePending = false;
	}
	if ((rc = FinishMirrorCopy(w))) return rc;
	if (w->toiSubSteps > 0 && w->bodyMirror != nullptr && std::min(w->bodyMirrorCount, w->bodyCount) > 0)
	{
		// the sub-steps moved bodies after the mirror copy of the step had been taken: take it again
		if ((rc = b2cuGetBodyStates(w, 0, std::min(w->bodyMirrorCount, w->bodyCount), w->bodyMirror))) return rc;
	}

	if (g_trace.enabled && g_trace.used > 1)
	{
		std::vector<std::pair<std::string, std::pair<float, int> > > rows;
		for (size_t i = 1; i < g_trace.used; ++i)
		{
			float t = 0.0f;
			cudaEventElapsedTime(&t, g_trace.events[i - 1], g_trace.events[i]);
			size_t k = 0;
			for (; k < rows.size(); ++k)
				if (rows[k].first == g_trace.names[i]) break;
			if (k == rows.size()) rows.push_back(std::make_pair(std::string(g_trace.names[i]), std::make_pair(0.0f, 0)));
			rows[k].second.first += t;
			rows[k].second.second += 1;
		}
		std::sort(rows.begin(), rows.end(), [](const std::pair<std::string, std::pair<float, int> >& a,
		                                       const std::pair<std::string, std::pair<float, int> >& b) {
			return a.second.first > b.second.first;
		});
		fprintf(stderr, "[b2cu colours]");
		for (int c = 0; c < B2CU_MAX_COLOURS + 2; ++c)
			if (w->colourCounts[c] > 0) fprintf(stderr, " %d:%d", c, w->colourCounts[c]);
		fprintf(stderr, "\n");
		fprintf(stderr, "[b2cu trace] step: %zu marks\n", g_trace.used);
		for (size_t k = 0; k < rows.size(); ++k)
			fprintf(stderr, "[b2cu trace] %-32s n=%4d %9.3f ms\n", rows[k].first.c_str(), rows[k].second.second,
			        rows[k].second.first);
	}

	float ms = 0.0f;
	cudaEventElapsedTime(&ms, w->ev[0], evEnd); out.step = ms;
	cudaEventElapsedTime(&ms, w->ev[1], w->ev[2]); out.collide = ms;
	cudaEventElapsedTime(&ms, w->ev[2], w->ev[7]); out.solve = ms;
	cudaEventElapsedTime(&ms, w->ev[2], w->ev[4]); out.solveTraversal = ms;
	cudaEventElapsedTime(&ms, w->ev[4], w->ev[5]); out.solveInit = ms;
	cudaEventElapsedTime(&ms, w->ev[5], w->ev[6]); out.solveVelocity = ms;
	cudaEventElapsedTime(&ms, w->ev[6], w->ev[7]); out.solvePosition = ms;
	cudaEventElapsedTime(&ms, w->ev[7], evBroad); out.broadphase = ms;
	out.broadphaseFindContacts = ms;
	cudaEventElapsedTime(&ms, evBroad, evEnd); out.solveTOI = ms;
	cudaEventElapsedTime(&ms, w->ev[0], w->ev[1]); out.broadphase += ms;

	out.bodyCount = nb;
	out.proxyCount = np;
	out.contactCount = w->contactCount - w->deadMain;
	out.touchingCount = w->hostCounters[CNT_TOUCHING];
	out.constraintCount = w->constraintCount;
	out.colourCount = w->colourCount;
	out.overflowCount = w->overflowCount;
	out.islandBodyCount = w->hostCounters[CNT_ISLAND_BODIES];
	out.awakeBodyCount = w->hostCounters[CNT_AWAKE_BODIES];
	out.moveCount = moved;
	out.newContactCount = newContacts;
	out.destroyedContactCount = destroyed;
	out.beginCount = w->beginCount;
	out.endCount = w->endCount;
	out.toiCandidateCount = w->toiCount;
	out.toiEventPending = w->toiEventPending;
	out.toiMinKey = w->toiMinKey;
	out.toiMinAlpha = w->toiMinAlpha;
	out.toiSubSteps = w->toiSubSteps;
	out.toiEventCount = w->toiEventCount;
	out.toiNewContactCount = w->toiNewContacts;
	out.kernelLaunches = w->launches + g_primLaunches;
	if (info) *info = out;
	return B2CU_OK;
}

int b2cuGetEvents(b2cuWorld* w, int32_t kind, int32_t capacity, b2cuContactKey* keys, int32_t* count)
{
	if (!w || capacity < 0 || (capacity > 0 && !keys)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	const int n = kind == B2CU_EVENT_BEGIN ? w->beginCount : w->endCount;
	if (count) *count = n;
	if (capacity == 0 || n == 0) return B2CU_OK;
	// The device appends events in arrival order; the deferred-callback order (ascending key, EndContact calls made
	// by Destroy after the sorted ends, b2ContactManager.cpp:388-439) is established here, when somebody asks: the
	// lists are short and a step that nobody listens to pays nothing for the ordering.
	std::vector<uint64_t> sorted(n);
	if (kind == B2CU_EVENT_BEGIN)
	{
		CUDA_TRY(w, cudaMemcpyAsync(sorted.data(), w->d.beginKeys, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, w->stream));
		int rc = SyncCheck(w);
		if (rc) return rc;
		std::sort(sorted.begin(), sorted.end());
	}
	else
	{
		const int nEnd = w->endUpdateCount, nDestroy = n - nEnd;
		if (nEnd > 0)
			CUDA_TRY(w, cudaMemcpyAsync(sorted.data(), w->d.endKeys, sizeof(uint64_t) * nEnd, cudaMemcpyDeviceToHost, w->stream));
		if (nDestroy > 0)
			CUDA_TRY(w, cudaMemcpyAsync(sorted.data() + nEnd, w->d.destroyEndKeys, sizeof(uint64_t) * nDestroy,
			                            cudaMemcpyDeviceToHost, w->stream));
		int rc = SyncCheck(w);
		if (rc) return rc;
		std::sort(sorted.begin(), sorted.begin() + nEnd);
		std::sort(sorted.begin() + nEnd, sorted.end());
	}
	memcpy(keys, sorted.data(), sizeof(uint64_t) * std::min(n, capacity));
	return B2CU_OK;
}

int b2cuGetToiEvents(b2cuWorld* w, int32_t capacity, b2cuContactKey* keys, int32_t* kinds, b2cuContact* records,
                     int32_t* count)
{
	if (!w || capacity < 0 || (capacity > 0 && (!keys || !kinds))) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	const int n = w->toiEventCount;
	if (count) *count = n;
	const int m = std::min(n, capacity);
	if (m <= 0) return B2CU_OK;
	// already in call order: the sub-steps run one after the other and one thread appends
	CUDA_TRY(w, cudaMemcpyAsync(keys, w->d.toiEventKeys, sizeof(uint64_t) * (size_t)m, cudaMemcpyDeviceToHost, w->stream));
	CUDA_TRY(w, cudaMemcpyAsync(kinds, w->d.toiEventKinds, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, w->stream));
	int rc = SyncCheck(w);
	if (rc || records == nullptr) return rc;
	return b2cuGetContactsByKey(w, m, keys, records);
}

// grow-only scratch pair (device + page-locked host) used by the query entry points
static int EnsureQueryScratch(b2cuWorld* w, size_t bytes)
{
	if (bytes > w->queryScratchBytes)
	{
		CUDA_TRY(w, cudaStreamSynchronize(w->stream));
		cudaFree(w->queryScratch);
		w->queryScratch = nullptr;
		w->queryScratchBytes = 0;
		size_t grown = std::max(bytes + bytes / 2, (size_t)1 << 20);
		CUDA_TRY(w, cudaMalloc(&w->queryScratch, grown));
		w->queryScratchBytes = grown;
	}
	return B2CU_OK;
}

static int EnsureQueryHost(b2cuWorld* w, size_t bytes)
{
	if (bytes > w->queryHostBytes)
	{
		if (w->queryHost) cudaFreeHost(w->queryHost);
		w->queryHost = nullptr;
		w->queryHostBytes = 0;
		size_t grown = std::max(bytes + bytes / 2, (size_t)1 << 20);
		CUDA_TRY(w, cudaMallocHost(&w->queryHost, grown));
		w->queryHostBytes = grown;
	}
	return B2CU_OK;
}

// keys and contact records of ALL events of the last step, fetched in one round trip on the first request and
// kept in the page-locked scratch: [begin | end (Update) | end (Destroy)] keys, then the records in the same order
// device half: gather + copy started on the step stream, no synchronisation
static int IssueEventRecords(b2cuWorld* w, bool* issued)
{
	*issued = false;
	const int nB = w->beginCount, nE = w->endUpdateCount, nD = w->endCount - w->endUpdateCount;
	const int n = nB + nE + nD;
	if (n == 0) return B2CU_OK;
	const size_t keyBytes = (sizeof(uint64_t) * (size_t)n + 255) & ~(size_t)255;
	const size_t total = keyBytes + sizeof(b2cuContact) * (size_t)n;
	int rc;
	if ((rc = EnsureQueryScratch(w, total)) || (rc = EnsureQueryHost(w, total))) return rc;
	uint64_t* dKeys = reinterpret_cast<uint64_t*>(w->queryScratch);
	b2cuContact* dOut = reinterpret_cast<b2cuContact*>(static_cast<char*>(w->queryScratch) + keyBytes);
	if (nB > 0)
		CUDA_TRY(w, cudaMemcpyAsync(dKeys, w->d.beginKeys, sizeof(uint64_t) * nB, cudaMemcpyDeviceToDevice, w->stream));
	if (nE > 0)
		CUDA_TRY(w, cudaMemcpyAsync(dKeys + nB, w->d.endKeys, sizeof(uint64_t) * nE, cudaMemcpyDeviceToDevice, w->stream));
	if (nD > 0)
		CUDA_TRY(w, cudaMemcpyAsync(dKeys + nB + nE, w->d.destroyEndKeys, sizeof(uint64_t) * nD, cudaMemcpyDeviceToDevice,
		                            w->stream));
	LAUNCH(w, GatherContactsByKeyKernel, GridFor(n), kBlock, w->d, w->contactCount, w->mainCount, (const uint64_t*)dKeys, n, dOut);
	CUDA_TRY(w, cudaMemcpyAsync(w->queryHost, w->queryScratch, total, cudaMemcpyDeviceToHost, w->stream));
	w->eventCacheKeyBytes = keyBytes;
	*issued = true;
	return B2CU_OK;
}

static int FetchEventRecords(b2cuWorld* w)
{
	if (w->eventCacheValid) return B2CU_OK;
	bool issued = false;
	int rc = IssueEventRecords(w, &issued);
	if (rc) return rc;
	if (issued && (rc = SyncCheck(w))) return rc;
	w->eventCacheValid = true;
	return B2CU_OK;
}

// (key, arrival index) pairs by key: the keys of one step's events are unique.  A few thousand events per step make a
// comparison sort the largest item of the host's event path; an LSD radix sort over the 11-bit digits that actually
// differ between the keys (two or three for worlds up to a million proxies) is several times cheaper.
static void SortEventOrder(std::pair<uint64_t, int>* a, std::pair<uint64_t, int>* tmp, int n)
{
	if (n < 2) return;
	if (n < 64)
	{
		std::sort(a, a + n);
		return;
	}
	uint64_t all0 = ~0ull, all1 = 0;
	for (int i = 0; i < n; ++i)
	{
		all0 &= a[i].first;
		all1 |= a[i].first;
	}
	const uint64_t varying = all0 ^ all1;
	std::pair<uint64_t, int>* src = a;
	std::pair<uint64_t, int>* dst = tmp;
	for (int shift = 0; shift < 64; shift += 11)
	{
		if (((varying >> shift) & 0x7FFull) == 0) continue;
		uint32_t count[2048];
		memset(count, 0, sizeof(count));
		for (int i = 0; i < n; ++i) ++count[(src[i].first >> shift) & 0x7FFull];
		uint32_t sum = 0;
		for (int k = 0; k < 2048; ++k)
		{
			uint32_t c = count[k];
			count[k] = sum;
			sum += c;
		}
		for (int i = 0; i < n; ++i) dst[count[(src[i].first >> shift) & 0x7FFull]++] = src[i];
		std::swap(src, dst);
	}
	if (src != a) memcpy(a, src, sizeof(a[0]) * (size_t)n);
}

int b2cuGetEventContacts(b2cuWorld* w, int32_t kind, int32_t capacity, b2cuContactKey* keys, b2cuContact* records,
                         int32_t* count)
{
	if (!w || capacity < 0 || (capacity > 0 && (!keys || !records))) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	const int n = kind == B2CU_EVENT_BEGIN ? w->beginCount : w->endCount;
	if (count) *count = n;
	if (capacity == 0 || n == 0) return B2CU_OK;
	// the records are gathered on the device in arrival order, keys and records come back together, and the callback
	// order of b2cuGetEvents is established on the host by sorting an index
	int rc = FetchEventRecords(w);
	if (rc) return rc;
	const int offset = kind == B2CU_EVENT_BEGIN ? 0 : w->beginCount;
	const int firstPart = kind == B2CU_EVENT_BEGIN ? n : w->endUpdateCount;
	const uint64_t* hKeys = reinterpret_cast<const uint64_t*>(w->queryHost) + offset;
	const b2cuContact* hOut =
	    reinterpret_cast<const b2cuContact*>(static_cast<const char*>(w->queryHost) + w->eventCacheKeyBytes) + offset;
	typedef std::pair<uint64_t, int> KeyIndex;
	if ((size_t)n > w->eventOrderCapacity)
	{
		free(w->eventOrder);
		w->eventOrderCapacity = (size_t)n + (size_t)n / 2 + 1024;
		w->eventOrder = malloc(sizeof(KeyIndex) * 2 * w->eventOrderCapacity);
		if (!w->eventOrder)
		{
			w->eventOrderCapacity = 0;
			return SetError(w, B2CU_ERR_CAPACITY, "b2cuGetEventContacts: no host memory for %d events", n);
		}
	}
	KeyIndex* order = static_cast<KeyIndex*>(w->eventOrder);
	KeyIndex* spare = order + w->eventOrderCapacity;
	for (int i = 0; i < n; ++i) order[(size_t)i] = std::make_pair(hKeys[i], i);
	SortEventOrder(order, spare, firstPart);
	SortEventOrder(order + firstPart, spare, n - firstPart);
	const int m = std::min(n, capacity);
	for (int j = 0; j < m; ++j)
	{
		keys[j] = order[(size_t)j].first;
		records[j] = hOut[order[(size_t)j].second];
	}
	return B2CU_OK;
}

int b2cuGetContactsByKey(b2cuWorld* w, int32_t count, const b2cuContactKey* keys, b2cuContact* out)
{
	if (!w || count < 0 || (count > 0 && (!keys || !out))) return B2CU_ERR_ARGUMENT;
	if (count == 0) return B2CU_OK;
	cudaSetDevice(w->device);
	for (int i = 0; i < count; ++i)
	{
		int a = (int)(keys[i] >> 32), b = (int)(keys[i] & 0xFFFFFFFFull);
		if (a < 0 || a >= w->proxyCount || b < 0 || b >= w->proxyCount)
			return SetError(w, B2CU_ERR_ARGUMENT, "key %d: proxies %d,%d out of range", i, a, b);
	}
	// keys and records go through a grow-only device scratch (no allocation per call)
	const size_t keyBytes = (sizeof(uint64_t) * (size_t)count + 255) & ~(size_t)255;
	const size_t need = keyBytes + sizeof(b2cuContact) * (size_t)count;
	{
		int rcs = EnsureQueryScratch(w, need);
		if (rcs) return rcs;
	}
	uint64_t* dKeys = reinterpret_cast<uint64_t*>(w->queryScratch);
	b2cuContact* dOut = reinterpret_cast<b2cuContact*>(static_cast<char*>(w->queryScratch) + keyBytes);
	cudaError_t e = cudaMemcpyAsync(dKeys, keys, sizeof(uint64_t) * count, cudaMemcpyHostToDevice, w->stream);
	if (e == cudaSuccess)
	{
		GatherContactsByKeyKernel<<<GridFor(count), kBlock, 0, w->stream>>>(w->d, w->contactCount, w->mainCount, dKeys, count,
		                                                                  dOut);
		e = cudaMemcpyAsync(out, dOut, sizeof(b2cuContact) * count, cudaMemcpyDeviceToHost, w->stream);
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
	if (e != cudaSuccess) return SetError(w, B2CU_ERR_CUDA, "b2cuGetContactsByKey: %s", cudaGetErrorString(e));
	return B2CU_OK;
}

int b2cuGetSolverOrder(b2cuWorld* w, int32_t capacity, b2cuContactKey* keys, int32_t* colour, int32_t* count)
{
	if (!w || capacity < 0) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	int n = w->constraintCount;
	if (count) *count = n;
	int m = std::min(n, capacity);
	if (m <= 0) return B2CU_OK;
	// storage order is colour-sorted (0..31, 32, 33); the solver ran own colours, own overflow (32), cross colours,
	// cross overflow (33): report that order
	std::vector<uint64_t> stored(n), order(n);
	CUDA_TRY(w, cudaMemcpyAsync(stored.data(), w->d.solverKeys, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, w->stream));
	CUDA_TRY(w, cudaMemcpyAsync(order.data(), w->d.orderKeys, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, w->stream));
	int rc = SyncCheck(w);
	if (rc) return rc;
	const int crossBase = w->shardCount > 1 ? 16 : 32;
	std::vector<int> sequence;
	for (int c = 0; c < crossBase; ++c) sequence.push_back(c);
	sequence.push_back(B2CU_MAX_COLOURS);
	for (int c = crossBase; c < B2CU_MAX_COLOURS; ++c) sequence.push_back(c);
	sequence.push_back(B2CU_MAX_COLOURS + 1);
	int at = 0;
	for (size_t q = 0; q < sequence.size(); ++q)
	{
		int c = sequence[q];
		for (int k = w->colourStarts[c]; k < w->colourStarts[c] + w->colourCounts[c]; ++k, ++at)
		{
			if (at >= m) break;
			if (keys) keys[at] = stored[k];
			if (colour) colour[at] = (int32_t)(order[k] >> B2CU_ORDER_COLOUR_SHIFT);
		}
	}
	return B2CU_OK;
}

int b2cuGetIslandLabels(b2cuWorld* w, int32_t first, int32_t count, int32_t* labels)
{
	int rc = CheckRange(w, first, count, w ? w->bodyCount : 0, labels);
	if (rc) return rc;
	cudaSetDevice(w->device);
	if (count > 0)
	{
		CUDA_TRY(w, cudaMemcpyAsync(labels, w->d.island + first, sizeof(int) * count, cudaMemcpyDeviceToHost, w->stream));
	}
	return SyncCheck(w);
}

int b2cuGetToiCandidates(b2cuWorld* w, int32_t capacity, b2cuContactKey* keys, int32_t* count)
{
	if (!w || capacity < 0 || (capacity > 0 && !keys)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	// evaluated when asked for (the step itself never needs the list in key order)
	DeviceArrays& d = w->d;
	const int nc = w->contactCount;
	int n = 0;
	int rc;
	if (nc > 0)
	{
		LAUNCH(w, ToiFlagsKernel, GridFor(nc), kBlock, d, nc, d.cSelect);
		CompactFlags(&w->prims, d.cSelect, nc, d.listA, d.counters + CNT_SCRATCH, w->stream);
		LAUNCH(w, GatherKeysKernel, GridFor(nc), kBlock, d.c.key, d.listA, d.counters + CNT_SCRATCH, (const int*)nullptr,
		       d.toiKeys, w->contactCapacity);
		CUDA_TRY(w, cudaMemcpyAsync(&w->hostCounters[CNT_SCRATCH], d.counters + CNT_SCRATCH, sizeof(int),
		                            cudaMemcpyDeviceToHost, w->stream));
		if ((rc = SyncCheck(w))) return rc;
		n = w->hostCounters[CNT_SCRATCH];
	}
	if (count) *count = n;
	int m = std::min(n, capacity);
	if (m > 0)
	{
		// main region and tail are each in key order
		std::vector<uint64_t> all((size_t)n);
		CUDA_TRY(w, cudaMemcpyAsync(all.data(), d.toiKeys, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, w->stream));
		if ((rc = SyncCheck(w))) return rc;
		std::sort(all.begin(), all.end());
		memcpy(keys, all.data(), sizeof(uint64_t) * (size_t)m);
	}
	return B2CU_OK;
}

int b2cuShardConfigure(b2cuWorld* w, int32_t rank, int32_t rankCount, int32_t ghostCount, const int32_t* ghostBodies,
                       int32_t exportCount, const int32_t* exportBodies, float gridFraction)
{
	if (!w || rankCount < 1 || rank < 0 || rank >= rankCount || ghostCount < 0 || exportCount < 0)
		return B2CU_ERR_ARGUMENT;
	if ((ghostCount > 0 && !ghostBodies) || (exportCount > 0 && !exportBodies)) return B2CU_ERR_ARGUMENT;
	if (rankCount > 1 && !w->persistentSolver)
		return SetError(w, B2CU_ERR_UNSUPPORTED, "sharding needs the persistent cooperative solver");
	cudaSetDevice(w->device);
	for (int i = 0; i < ghostCount; ++i)
		if (ghostBodies[i] < 0 || ghostBodies[i] >= w->bodyCount)
			return SetError(w, B2CU_ERR_ARGUMENT, "ghost body %d out of range", ghostBodies[i]);
	for (int i = 0; i < exportCount; ++i)
		if (exportBodies[i] < 0 || exportBodies[i] >= w->bodyCount)
			return SetError(w, B2CU_ERR_ARGUMENT, "export body %d out of range", exportBodies[i]);
	cudaFree(w->ghostIds);
	cudaFree(w->exportIds);
	cudaFree(w->mailbox);
	w->ghostIds = w->exportIds = nullptr;
	w->mailbox = nullptr;
	w->shardRank = rank;
	w->shardCount = rankCount;
	w->ghostCount = ghostCount;
	w->exportCount = exportCount;
	w->shardSeq = 1;
	CUDA_TRY(w, cudaMalloc(&w->ghostIds, sizeof(int) * std::max(1, ghostCount)));
	CUDA_TRY(w, cudaMalloc(&w->exportIds, sizeof(int) * std::max(1, exportCount)));
	if (ghostCount) CUDA_TRY(w, cudaMemcpy(w->ghostIds, ghostBodies, sizeof(int) * ghostCount, cudaMemcpyHostToDevice));
	if (exportCount) CUDA_TRY(w, cudaMemcpy(w->exportIds, exportBodies, sizeof(int) * exportCount, cudaMemcpyHostToDevice));
	w->mailboxBytes = kMailboxHeader + MailboxFromUpperBytes(ghostCount) + MailboxFromLowerBytes(exportCount);
	CUDA_TRY(w, cudaMalloc(&w->mailbox, w->mailboxBytes));
	CUDA_TRY(w, cudaMemset(w->mailbox, 0, w->mailboxBytes));
	const float fraction = (gridFraction > 0.0f && gridFraction < 1.0f) ? gridFraction : 1.0f;
	// several shards on one device (tests): their persistent kernels must be co-resident
	w->persistentGrid = std::max(1, (int)(w->persistentGridMax * fraction));
	w->persistentGridPosition = std::max(1, (int)(w->persistentGridPositionMax * fraction));
	w->shardFlowGrid = std::max(1, (int)(w->flowGridMax * fraction));
	w->shardFlowGridPosition = std::max(1, (int)(w->flowGridPositionMax * fraction));
	{
		const char* e = getenv("B2CU_SHARD_FLOW");
		w->shardFlow = rankCount > 1 && !(e && atoi(e) == 0);
	}
	// slot of every halo body in the mailboxes of the dataflow solver
	{
		std::vector<int> slots((size_t)std::max(1, w->bodyCapacity), -1);
		for (int i = 0; i < ghostCount; ++i) slots[(size_t)ghostBodies[i]] = i;
		for (int i = 0; i < exportCount; ++i) slots[(size_t)exportBodies[i]] = i | 0x40000000;
		CUDA_TRY(w, cudaMemcpy(w->d.haloSlot, slots.data(), sizeof(int) * (size_t)w->bodyCapacity, cudaMemcpyHostToDevice));
	}
	return B2CU_OK;
}

int b2cuShardGetLink(b2cuWorld* w, b2cuShardLink* link)
{
	if (!w || !link || !w->mailbox) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	memset(link, 0, sizeof(*link));
	cudaIpcMemHandle_t h;
	CUDA_TRY(w, cudaIpcGetMemHandle(&h, w->mailbox));
	static_assert(sizeof(h) <= sizeof(link->ipcHandle), "ipc handle size");
	memcpy(link->ipcHandle, &h, sizeof(h));
	link->localPointer = (uint64_t)(uintptr_t)w->mailbox;
	link->processId = (int32_t)getpid();
	link->device = w->device;
	link->ghostCount = w->ghostCount;
	link->exportCount = w->exportCount;
	link->rank = w->shardRank;
	link->rankCount = w->shardCount;
	return B2CU_OK;
}

static int OpenPeer(b2cuWorld* w, const b2cuShardLink* link, unsigned char** out, bool* ipc)
{
	*out = nullptr;
	*ipc = false;
	if (link->processId == (int32_t)getpid())
	{
		if (link->device != w->device)
		{
			int can = 0;
			cudaDeviceCanAccessPeer(&can, w->device, link->device);
			if (!can) return SetError(w, B2CU_ERR_UNSUPPORTED, "device %d cannot access device %d", w->device, link->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(link->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
				return SetError(w, B2CU_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
			cudaGetLastError();
		}
		*out = reinterpret_cast<unsigned char*>((uintptr_t)link->localPointer);
		return B2CU_OK;
	}
	cudaIpcMemHandle_t h;
	memcpy(&h, link->ipcHandle, sizeof(h));
	void* p = nullptr;
	CUDA_TRY(w, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	*out = static_cast<unsigned char*>(p);
	*ipc = true;
	return B2CU_OK;
}

int b2cuShardConnect(b2cuWorld* w, const b2cuShardLink* lower, const b2cuShardLink* upper)
{
	if (!w || !w->mailbox) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	int rc;
	if (lower)
	{
		if (lower->rank != w->shardRank - 1 || lower->ghostCount != w->exportCount)
			return SetError(w, B2CU_ERR_ARGUMENT, "lower link: rank %d ghosts %d, expected rank %d ghosts %d", lower->rank,
			                lower->ghostCount, w->shardRank - 1, w->exportCount);
		if ((rc = OpenPeer(w, lower, &w->peerLower, &w->peerLowerIpc))) return rc;
	}
	if (upper)
	{
		if (upper->rank != w->shardRank + 1 || upper->exportCount != w->ghostCount)
			return SetError(w, B2CU_ERR_ARGUMENT, "upper link: rank %d exports %d, expected rank %d exports %d",
			                upper->rank, upper->exportCount, w->shardRank + 1, w->ghostCount);
		if ((rc = OpenPeer(w, upper, &w->peerUpper, &w->peerUpperIpc))) return rc;
		w->peerUpperGhostCountOfUpper = upper->ghostCount;
	}
	// a neighbour on this very device: cooperative launches of one device run one after the other, but the shards wait
	// for each other inside their solver kernels, so these kernels become ordinary launches with their own grid barrier
	w->shardSoftBarrier = (lower && lower->device == w->device) || (upper && upper->device == w->device);
	{
		const char* e = getenv("B2CU_SOFT_BARRIER");
		if (e) w->shardSoftBarrier = atoi(e) != 0;
	}
	if (w->shardSoftBarrier && w->softBarrierCounter == nullptr)
		CUDA_TRY(w, cudaMalloc(&w->softBarrierCounter, sizeof(unsigned)));
	return B2CU_OK;
}

int b2cuCollidePairs(int32_t device, int32_t shapeCount, const b2cuShape* shapes, int32_t pairCount,
                     const int32_t* shapeA, const float* xfA, const int32_t* shapeB, const float* xfB,
                     b2cuManifold* manifolds)
{
	if (shapeCount <= 0 || pairCount < 0 || !shapes || !shapeA || !shapeB || !xfA || !xfB || !manifolds)
		return B2CU_ERR_ARGUMENT;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return B2CU_ERR_NO_DEVICE;
	if (cudaSetDevice(device) != cudaSuccess) return B2CU_ERR_CUDA;
	if (pairCount == 0) return B2CU_OK;
	b2cuShape* dShapes = nullptr;
	int *dA = nullptr, *dB = nullptr;
	float4 *dXa = nullptr, *dXb = nullptr;
	b2cuManifold* dOut = nullptr;
	cudaError_t e = cudaSuccess;
	if (e == cudaSuccess) e = cudaMalloc(&dShapes, sizeof(b2cuShape) * shapeCount);
	if (e == cudaSuccess) e = cudaMalloc(&dA, sizeof(int) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dB, sizeof(int) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dXa, sizeof(float4) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dXb, sizeof(float4) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dOut, sizeof(b2cuManifold) * pairCount);
	if (e == cudaSuccess) e = cudaMemcpy(dShapes, shapes, sizeof(b2cuShape) * shapeCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dA, shapeA, sizeof(int) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dB, shapeB, sizeof(int) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dXa, xfA, sizeof(float4) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dXb, xfB, sizeof(float4) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
	{
		CollidePairsKernel<<<GridFor(pairCount), kBlock>>>(dShapes, pairCount, dA, dXa, dB, dXb, dOut);
		e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess) e = cudaMemcpy(manifolds, dOut, sizeof(b2cuManifold) * pairCount, cudaMemcpyDeviceToHost);
	cudaFree(dShapes);
	cudaFree(dA);
	cudaFree(dB);
	cudaFree(dXa);
	cudaFree(dXb);
	cudaFree(dOut);
	return e == cudaSuccess ? B2CU_OK : B2CU_ERR_CUDA;
}

// shared tail of the two query entry points: flags (cellOfProxy, free between steps) -> ascending id list (movedList)
static int FinishProxyQuery(b2cuWorld* w, int32_t capacity, int32_t* proxyIds, int32_t* count)
{
	DeviceArrays& d = w->d;
	const int np = w->proxyCount;
	CompactFlags(&w->prims, d.cellOfProxy, np, d.movedList, d.counters + CNT_SCRATCH, w->stream);
	CUDA_TRY(w, cudaMemcpyAsync(&w->hostCounters[CNT_SCRATCH], d.counters + CNT_SCRATCH, sizeof(int), cudaMemcpyDeviceToHost,
	                            w->stream));
	int rc = SyncCheck(w);
	if (rc) return rc;
	const int n = w->hostCounters[CNT_SCRATCH];
	if (count) *count = n;
	const int m = std::min(n, capacity);
	if (m > 0)
	{
		CUDA_TRY(w, cudaMemcpyAsync(proxyIds, d.movedList, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, w->stream));
		rc = SyncCheck(w);
	}
	return rc;
}

int b2cuQueryAABB(b2cuWorld* w, const float aabb[4], int32_t capacity, int32_t* proxyIds, int32_t* count)
{
	if (!w || !aabb || capacity < 0 || (capacity > 0 && !proxyIds)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	if (count) *count = 0;
	if (w->proxyCount == 0) return B2CU_OK;
	LAUNCH(w, QueryAabbSelectKernel, GridFor(w->proxyCount), kBlock, w->d, w->proxyCount,
	       make_float4(aabb[0], aabb[1], aabb[2], aabb[3]), w->d.cellOfProxy);
	return FinishProxyQuery(w, capacity, proxyIds, count);
}

int b2cuRayCastCandidates(b2cuWorld* w, const float p1[2], const float p2[2], int32_t capacity, int32_t* proxyIds,
                          int32_t* count)
{
	if (!w || !p1 || !p2 || capacity < 0 || (capacity > 0 && !proxyIds)) return B2CU_ERR_ARGUMENT;
	cudaSetDevice(w->device);
	if (count) *count = 0;
	if (w->proxyCount == 0) return B2CU_OK;
	LAUNCH(w, RayCastSelectKernel, GridFor(w->proxyCount), kBlock, w->d, w->proxyCount, make_float2(p1[0], p1[1]),
	       make_float2(p2[0], p2[1]), w->d.cellOfProxy);
	return FinishProxyQuery(w, capacity, proxyIds, count);
}

int b2cuDistancePairs(int32_t device, int32_t shapeCount, const b2cuShape* shapes, int32_t pairCount,
                      const int32_t* shapeA, const float* xfA, const int32_t* shapeB, const float* xfB, int32_t useRadii,
                      b2cuDistanceResult* results)
{
	if (shapeCount <= 0 || pairCount < 0 || !shapes || !shapeA || !shapeB || !xfA || !xfB || !results)
		return B2CU_ERR_ARGUMENT;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return B2CU_ERR_NO_DEVICE;
	if (cudaSetDevice(device) != cudaSuccess) return B2CU_ERR_CUDA;
	if (pairCount == 0) return B2CU_OK;
	b2cuShape* dShapes = nullptr;
	int *dA = nullptr, *dB = nullptr;
	float4 *dXa = nullptr, *dXb = nullptr;
	b2cuDistanceResult* dOut = nullptr;
	cudaError_t e = cudaSuccess;
	if (e == cudaSuccess) e = cudaMalloc(&dShapes, sizeof(b2cuShape) * shapeCount);
	if (e == cudaSuccess) e = cudaMalloc(&dA, sizeof(int) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dB, sizeof(int) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dXa, sizeof(float4) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dXb, sizeof(float4) * pairCount);
	if (e == cudaSuccess) e = cudaMalloc(&dOut, sizeof(b2cuDistanceResult) * pairCount);
	if (e == cudaSuccess) e = cudaMemcpy(dShapes, shapes, sizeof(b2cuShape) * shapeCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dA, shapeA, sizeof(int) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dB, shapeB, sizeof(int) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dXa, xfA, sizeof(float4) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dXb, xfB, sizeof(float4) * pairCount, cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
	{
		DistancePairsKernel<<<GridFor(pairCount), kBlock>>>(dShapes, pairCount, dA, dXa, dB, dXb, useRadii, dOut);
		e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess) e = cudaMemcpy(results, dOut, sizeof(b2cuDistanceResult) * pairCount, cudaMemcpyDeviceToHost);
	cudaFree(dShapes);
	cudaFree(dA);
	cudaFree(dB);
	cudaFree(dXa);
	cudaFree(dXb);
	cudaFree(dOut);
	return e == cudaSuccess ? B2CU_OK : B2CU_ERR_CUDA;
}

// device copy of a host array for the stand-alone batched entries; freed by the caller
static cudaError_t UploadArray(void* dstPointer, const void* src, size_t bytes, cudaError_t e)
{
	if (e != cudaSuccess) return e;
	void** dst = (void**)dstPointer;
	e = cudaMalloc(dst, bytes);
	if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
	return e;
}

int b2cuTimeOfImpactPairs(int32_t device, int32_t shapeCount, const b2cuShape* shapes, int32_t pairCount,
                          const int32_t* shapeA, const b2cuSweep* sweepA, const int32_t* shapeB, const b2cuSweep* sweepB,
                          const float* tMax, b2cuToiResult* results)
{
	if (shapeCount <= 0 || pairCount < 0 || !shapes || !shapeA || !shapeB || !sweepA || !sweepB || !tMax || !results)
		return B2CU_ERR_ARGUMENT;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return B2CU_ERR_NO_DEVICE;
	if (cudaSetDevice(device) != cudaSuccess) return B2CU_ERR_CUDA;
	if (pairCount == 0) return B2CU_OK;
	b2cuShape* dShapes = nullptr;
	int32_t *dA = nullptr, *dB = nullptr;
	b2cuSweep *dSa = nullptr, *dSb = nullptr;
	float* dT = nullptr;
	b2cuToiResult* dOut = nullptr;
	cudaError_t e = cudaSuccess;
	e = UploadArray(&dShapes, shapes, sizeof(b2cuShape) * shapeCount, e);
	e = UploadArray(&dA, shapeA, sizeof(int32_t) * pairCount, e);
	e = UploadArray(&dB, shapeB, sizeof(int32_t) * pairCount, e);
	e = UploadArray(&dSa, sweepA, sizeof(b2cuSweep) * pairCount, e);
	e = UploadArray(&dSb, sweepB, sizeof(b2cuSweep) * pairCount, e);
	e = UploadArray(&dT, tMax, sizeof(float) * pairCount, e);
	if (e == cudaSuccess) e = cudaMalloc(&dOut, sizeof(b2cuToiResult) * pairCount);
	if (e == cudaSuccess)
	{
		TimeOfImpactPairsKernel<<<GridFor(pairCount), kBlock>>>(dShapes, pairCount, dA, dSa, dB, dSb, dT, dOut);
		e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess) e = cudaMemcpy(results, dOut, sizeof(b2cuToiResult) * pairCount, cudaMemcpyDeviceToHost);
	cudaFree(dShapes);
	cudaFree(dA);
	cudaFree(dB);
	cudaFree(dSa);
	cudaFree(dSb);
	cudaFree(dT);
	cudaFree(dOut);
	return e == cudaSuccess ? B2CU_OK : B2CU_ERR_CUDA;
}

int b2cuSinCos(int32_t device, int32_t count, const float* angles, float* sinOut, float* cosOut)
{
	if (count < 0 || !angles || !sinOut || !cosOut) return B2CU_ERR_ARGUMENT;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return B2CU_ERR_NO_DEVICE;
	if (cudaSetDevice(device) != cudaSuccess) return B2CU_ERR_CUDA;
	if (count == 0) return B2CU_OK;
	float *dx = nullptr, *ds = nullptr, *dc = nullptr;
	cudaError_t e = cudaMalloc(&dx, sizeof(float) * count);
	if (e == cudaSuccess) e = cudaMalloc(&ds, sizeof(float) * count);
	if (e == cudaSuccess) e = cudaMalloc(&dc, sizeof(float) * count);
	if (e == cudaSuccess) e = cudaMemcpy(dx, angles, sizeof(float) * count, cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
	{
		SinCosKernel<<<GridFor(count), kBlock>>>(count, dx, ds, dc);
		e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess) e = cudaMemcpy(sinOut, ds, sizeof(float) * count, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(cosOut, dc, sizeof(float) * count, cudaMemcpyDeviceToHost);
	cudaFree(dx);
	cudaFree(ds);
	cudaFree(dc);
	return e == cudaSuccess ? B2CU_OK : B2CU_ERR_CUDA;
}

} // extern "C"
