// prims.cu -- exclusive scan, stable compaction, LSD radix sort (see b2cu_prims.cuh).
#include "b2cu_prims.cuh"

namespace b2cu
{

int g_primLaunches = 0;
bool g_pdl = true;
void (*g_primTraceHook)(const char* name, cudaStream_t stream) = nullptr;
#define PRIM_MARK(name)                                  \
	do                                                   \
	{                                                    \
		++g_primLaunches;                                \
		if (g_primTraceHook) g_primTraceHook(name, stream); \
	} while (0)

static const int SCAN_BLOCK = 256;
static const int SCAN_ITEMS = 4;
static const int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

static const int RADIX_BLOCK = 256;
static const int RADIX_ITEMS = 8;
static const int RADIX_TILE = RADIX_BLOCK * RADIX_ITEMS;

struct IntLoader
{
	const int* p;
	__device__ __forceinline__ int operator()(int i) const { return p[i]; }
};
struct NonZeroLoader
{
	const int* p;
	__device__ __forceinline__ int operator()(int i) const { return p[i] != 0 ? 1 : 0; }
};
struct NotMaskLoader
{
	const uint32_t* p;
	uint32_t mask;
	__device__ __forceinline__ int operator()(int i) const { return (p[i] & mask) == 0 ? 1 : 0; }
};
struct MaskLoader
{
	const uint32_t* p;
	uint32_t mask;
	__device__ __forceinline__ int operator()(int i) const { return (p[i] & mask) != 0 ? 1 : 0; }
};

template <typename Loader>
__global__ void __launch_bounds__(SCAN_BLOCK) ScanTilesKernel(Loader load, int* __restrict__ out,
                                                              int* __restrict__ tileSums, int n, int* total)
{
	GridDependencyWait();
	__shared__ int warpSums[SCAN_BLOCK / 32];
	const int tile = blockIdx.x;
	const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;

	int v[SCAN_ITEMS];
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
	{
		v[k] = (base + k < n) ? load(base + k) : 0;
	}
	int tsum = 0;
	int pre[SCAN_ITEMS];
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
	{
		pre[k] = tsum;
		tsum += v[k];
	}

	int inc = tsum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		int t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += t;
	}
	if (lane == 31) warpSums[warp] = inc;
	__syncthreads();
	if (warp == 0)
	{
		int w = lane < SCAN_BLOCK / 32 ? warpSums[lane] : 0;
		int winc = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			int t = __shfl_up_sync(0xffffffffu, winc, d);
			if (lane >= d) winc += t;
		}
		if (lane < SCAN_BLOCK / 32) warpSums[lane] = winc - w; // exclusive warp offsets
		if (lane == SCAN_BLOCK / 32 - 1)
		{
			if (tileSums) tileSums[tile] = winc;
			if (total && gridDim.x == 1) *total = winc;
		}
	}
	__syncthreads();
	const int threadOffset = (inc - tsum) + warpSums[warp];
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
	{
		if (base + k < n) out[base + k] = threadOffset + pre[k];
	}
}

// Single-pass scan (decoupled look-back): a tile publishes its aggregate, then the inclusive prefix once it knows
// its predecessors'.  Tile ids are handed out by an atomic ticket, so a tile only ever waits for tiles that are
// already running.  state word = status << 32 | value, written with one 64-bit store.
#define SCAN_STATE_AGGREGATE 1ull
#define SCAN_STATE_INCLUSIVE 2ull

template <typename Loader>
__global__ void __launch_bounds__(SCAN_BLOCK) ScanLookbackKernel(Loader load, int* __restrict__ out,
                                                                 unsigned long long* state, int* ticket, int n, int* total)
{
	GridDependencyWait();
	__shared__ int warpSums[SCAN_BLOCK / 32];
	__shared__ int shTile, shExclusive;
	if (threadIdx.x == 0) shTile = atomicAdd(ticket, 1);
	__syncthreads();
	const int tile = shTile;
	const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;

	int v[SCAN_ITEMS];
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < n) ? load(base + k) : 0;
	int tsum = 0;
	int pre[SCAN_ITEMS];
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
	{
		pre[k] = tsum;
		tsum += v[k];
	}
	int inc = tsum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		int t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += t;
	}
	if (lane == 31) warpSums[warp] = inc;
	__syncthreads();
	if (warp == 0)
	{
		int w = lane < SCAN_BLOCK / 32 ? warpSums[lane] : 0;
		int winc = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			int t = __shfl_up_sync(0xffffffffu, winc, d);
			if (lane >= d) winc += t;
		}
		if (lane < SCAN_BLOCK / 32) warpSums[lane] = winc - w; // exclusive warp offsets
		// whole warp 0: publish the aggregate, then look back 32 predecessors at a time
		const int aggregate = __shfl_sync(0xffffffffu, winc, SCAN_BLOCK / 32 - 1);
		volatile unsigned long long* vs = state;
		int exclusive = 0;
		if (tile == 0)
		{
			if (lane == 0) vs[0] = (SCAN_STATE_INCLUSIVE << 32) | (unsigned)aggregate;
		}
		else
		{
			if (lane == 0) vs[tile] = (SCAN_STATE_AGGREGATE << 32) | (unsigned)aggregate;
			for (int j0 = tile - 1;; j0 -= 32)
			{
				const int j = j0 - lane;
				unsigned long long st = (SCAN_STATE_INCLUSIVE << 32); // before tile 0: prefix 0
				if (j >= 0)
				{
					do
					{
						st = vs[j];
					} while ((st >> 32) == 0ull);
				}
				const unsigned inclusiveLanes = __ballot_sync(0xffffffffu, (st >> 32) == SCAN_STATE_INCLUSIVE);
				const int stop = inclusiveLanes ? __ffs((int)inclusiveLanes) - 1 : 31;
				int val = lane <= stop ? (int)(unsigned)(st & 0xFFFFFFFFull) : 0;
#pragma unroll
				for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
				exclusive += val;
				if (inclusiveLanes) break;
			}
			if (lane == 0) vs[tile] = (SCAN_STATE_INCLUSIVE << 32) | (unsigned)(exclusive + aggregate);
		}
		if (lane == 0)
		{
			shExclusive = exclusive;
			if (total && tile == gridDim.x - 1) *total = exclusive + aggregate;
		}
	}
	__syncthreads();
	const int threadOffset = shExclusive + (inc - tsum) + warpSums[warp];
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
	{
		if (base + k < n) out[base + k] = threadOffset + pre[k];
	}
}

__global__ void SetIntKernel(int* p, int v) {
	GridDependencyWait(); *p = v; }

cudaError_t PrimScratchAlloc(PrimScratch* s, int capacity)
{
	PrimScratchFree(s);
	if (capacity < 1024) capacity = 1024;
	s->capacity = capacity;
	s->radixBlocks = (capacity + RADIX_TILE - 1) / RADIX_TILE;
	int histSize = 256 * s->radixBlocks;
	int scanCap = capacity > histSize ? capacity : histSize;
	cudaError_t e;
	if ((e = cudaMalloc(&s->scanState, sizeof(unsigned long long) * (scanCap / SCAN_TILE + 4))) != cudaSuccess) return e;
	if ((e = cudaMalloc(&s->radixHist, sizeof(int) * histSize)) != cudaSuccess) return e;
	if ((e = cudaMalloc(&s->radixAlt, sizeof(uint64_t) * capacity)) != cudaSuccess) return e;
	if ((e = cudaMalloc(&s->compactPos, sizeof(int) * capacity)) != cudaSuccess) return e;
	return cudaSuccess;
}

void PrimScratchFree(PrimScratch* s)
{
	cudaFree(s->scanState);
	cudaFree(s->radixHist);
	cudaFree(s->radixAlt);
	cudaFree(s->compactPos);
	*s = PrimScratch();
}

template <typename Loader>
static void ScanImpl(PrimScratch* s, Loader load, int* out, int n, int* total, cudaStream_t stream)
{
	if (n <= 0)
	{
		if (total)
		{
			LaunchPdl(SetIntKernel, dim3(1), dim3(1), stream, total, 0);
			PRIM_MARK("SetInt");
		}
		return;
	}
	int tiles1 = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (tiles1 == 1)
	{
		LaunchPdl(ScanTilesKernel<Loader>, dim3(1), dim3(SCAN_BLOCK), stream, load, out, (int*)nullptr, n, total);
		PRIM_MARK("ScanTiles");
		return;
	}
	// scanState: [ticket (8 bytes) | one state word per tile]
	cudaMemsetAsync(s->scanState, 0, sizeof(unsigned long long) * (size_t)(tiles1 + 1), stream);
	LaunchPdl(ScanLookbackKernel<Loader>, dim3(tiles1), dim3(SCAN_BLOCK), stream, load, out, s->scanState + 1, reinterpret_cast<int*>(s->scanState), n,
	                                                      total);
	PRIM_MARK("ScanLookback");
}

void ExclusiveScanNotMask(PrimScratch* s, const uint32_t* flags, uint32_t mask, int* out, int n, int* total,
                          cudaStream_t stream)
{
	NotMaskLoader l{flags, mask};
	ScanImpl(s, l, out, n, total, stream);
}

void ExclusiveScan(PrimScratch* s, const int* in, int* out, int n, int* total, cudaStream_t stream)
{
	IntLoader l{in};
	ScanImpl(s, l, out, n, total, stream);
}

template <typename Loader>
__global__ void CompactScatterKernel(Loader load, const int* __restrict__ pos, int n, int* __restrict__ outIdx)
{
	GridDependencyWait();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && load(i))
	{
		outIdx[pos[i]] = i;
	}
}

void CompactFlags(PrimScratch* s, const int* flags, int n, int* outIdx, int* outCount, cudaStream_t stream)
{
	NonZeroLoader l{flags};
	ScanImpl(s, l, s->compactPos, n, outCount, stream);
	if (n > 0)
	{
		LaunchPdl(CompactScatterKernel<NonZeroLoader>, dim3((n + 255) / 256), dim3(256), stream, l, s->compactPos, n, outIdx);
		PRIM_MARK("CompactScatter");
	}
}

void CompactMask(PrimScratch* s, const uint32_t* flags, uint32_t mask, int n, int* outIdx, int* outCount,
                 cudaStream_t stream)
{
	MaskLoader l{flags, mask};
	ScanImpl(s, l, s->compactPos, n, outCount, stream);
	if (n > 0)
	{
		LaunchPdl(CompactScatterKernel<MaskLoader>, dim3((n + 255) / 256), dim3(256), stream, l, s->compactPos, n, outIdx);
		PRIM_MARK("CompactScatter");
	}
}

// ---- radix sort -----------------------------------------------------------------------------------------

__global__ void __launch_bounds__(RADIX_BLOCK) RadixHistogramKernel(const uint64_t* __restrict__ keys, int n, int shift,
                                                                    int* __restrict__ hist, int numBlocks)
{
	GridDependencyWait();
	__shared__ int sh[256];
	sh[threadIdx.x] = 0;
	__syncthreads();
	const int base = blockIdx.x * RADIX_TILE;
#pragma unroll
	for (int r = 0; r < RADIX_ITEMS; ++r)
	{
		int i = base + r * RADIX_BLOCK + threadIdx.x;
		if (i < n)
		{
			int d = (int)((keys[i] >> shift) & 0xFFull);
			atomicAdd(&sh[d], 1);
		}
	}
	__syncthreads();
	hist[threadIdx.x * numBlocks + blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RADIX_BLOCK) RadixScatterKernel(const uint64_t* __restrict__ keys,
                                                                  uint64_t* __restrict__ out, int n, int shift,
                                                                  const int* __restrict__ offsets, int numBlocks)
{
	GridDependencyWait();
	__shared__ int warpCount[RADIX_BLOCK / 32][256];
	__shared__ int running[256];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const unsigned ltMask = (1u << lane) - 1u;

	running[threadIdx.x] = offsets[threadIdx.x * numBlocks + blockIdx.x];
	const int base = blockIdx.x * RADIX_TILE;

	for (int r = 0; r < RADIX_ITEMS; ++r)
	{
		for (int w = 0; w < RADIX_BLOCK / 32; ++w)
		{
			warpCount[w][threadIdx.x] = 0;
		}
		__syncthreads();

		int i = base + r * RADIX_BLOCK + threadIdx.x;
		bool valid = i < n;
		uint64_t key = valid ? keys[i] : 0ull;
		unsigned d = valid ? (unsigned)((key >> shift) & 0xFFull) : 0xFFFFFFFFu;
		unsigned peers = __match_any_sync(0xffffffffu, d);
		int rankInWarp = __popc(peers & ltMask);
		if (valid && rankInWarp == 0)
		{
			warpCount[warp][d] = __popc(peers);
		}
		__syncthreads();

		if (valid)
		{
			int pos = running[d] + rankInWarp;
			for (int w = 0; w < warp; ++w)
			{
				pos += warpCount[w][d];
			}
			out[pos] = key;
		}
		__syncthreads();

		int add = 0;
		for (int w = 0; w < RADIX_BLOCK / 32; ++w)
		{
			add += warpCount[w][threadIdx.x];
		}
		running[threadIdx.x] += add;
		__syncthreads();
	}
}

// Sort of small / medium key lists (event buffers, new pairs: hundreds to ~10^5 keys) without the launch train
// of the LSD radix sort: every CTA sorts one tile of 4096 keys with a bitonic network in shared memory, then
// log2(tiles) merge rounds place each key by binary search in the sibling run.  Full 64-bit ascending order.
static const int SORT_TILE = 4096;

__global__ void __launch_bounds__(1024) TileSort64Kernel(uint64_t* __restrict__ keys, int n)
{
	GridDependencyWait();
	__shared__ uint64_t sh[SORT_TILE];
	const int base = blockIdx.x * SORT_TILE;
	const int count = min(SORT_TILE, n - base);
	int m = 1;
	while (m < count) m <<= 1;
	for (int i = threadIdx.x; i < m; i += blockDim.x) sh[i] = i < count ? keys[base + i] : 0xFFFFFFFFFFFFFFFFull;
	__syncthreads();
	for (int k = 2; k <= m; k <<= 1)
	{
		for (int j = k >> 1; j > 0; j >>= 1)
		{
			for (int i = threadIdx.x; i < m; i += blockDim.x)
			{
				int ixj = i ^ j;
				if (ixj > i)
				{
					uint64_t a = sh[i], b = sh[ixj];
					bool up = (i & k) == 0;
					if ((a > b) == up)
					{
						sh[i] = b;
						sh[ixj] = a;
					}
				}
			}
			__syncthreads();
		}
	}
	for (int i = threadIdx.x; i < count; i += blockDim.x) keys[base + i] = sh[i];
}

// runs of length `run` are sorted; merge run pairs (2r, 2r+1) from src into dst
__global__ void MergeRuns64Kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int n, int run)
{
	GridDependencyWait();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int r = i / run;
	int pairBase = (r & ~1) * run;
	int sibBase = (r ^ 1) * run;
	uint64_t key = src[i];
	int sibCount = sibBase < n ? min(run, n - sibBase) : 0;
	// left run: count sibling keys strictly smaller; right run: count sibling keys smaller or equal (stable)
	int lo = 0, hi = sibCount;
	const bool left = (r & 1) == 0;
	while (lo < hi)
	{
		int mid = (lo + hi) >> 1;
		uint64_t v = src[sibBase + mid];
		bool goRight = left ? (v < key) : (v <= key);
		if (goRight) lo = mid + 1;
		else hi = mid;
	}
	dst[pairBase + (i - r * run) + lo] = key;
}

void SortSmall64(PrimScratch* s, uint64_t* keys, int n, cudaStream_t stream)
{
	if (n <= 1) return;
	int tiles = (n + SORT_TILE - 1) / SORT_TILE;
	LaunchPdl(TileSort64Kernel, dim3(tiles), dim3(1024), stream, keys, n);
	PRIM_MARK("TileSort64");
	uint64_t* src = keys;
	uint64_t* dst = s->radixAlt;
	for (int run = SORT_TILE; run < n; run <<= 1)
	{
		LaunchPdl(MergeRuns64Kernel, dim3((n + 255) / 256), dim3(256), stream, src, dst, n, run);
		PRIM_MARK("MergeRuns64");
		uint64_t* t = src;
		src = dst;
		dst = t;
	}
	if (src != keys) cudaMemcpyAsync(keys, src, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, stream);
}

void RadixSort64(PrimScratch* s, uint64_t* keys, int n, int beginBit, int endBit, cudaStream_t stream)
{
	if (n <= 1)
	{
		return;
	}
	int numBlocks = (n + RADIX_TILE - 1) / RADIX_TILE;
	uint64_t* src = keys;
	uint64_t* dst = s->radixAlt;
	for (int shift = beginBit; shift < endBit; shift += 8)
	{
		LaunchPdl(RadixHistogramKernel, dim3(numBlocks), dim3(RADIX_BLOCK), stream, src, n, shift, s->radixHist, numBlocks);
		PRIM_MARK("RadixHistogram");
		ExclusiveScan(s, s->radixHist, s->radixHist, 256 * numBlocks, nullptr, stream);
		LaunchPdl(RadixScatterKernel, dim3(numBlocks), dim3(RADIX_BLOCK), stream, src, dst, n, shift, s->radixHist, numBlocks);
		PRIM_MARK("RadixScatter");
		uint64_t* t = src;
		src = dst;
		dst = t;
	}
	if (src != keys)
	{
		cudaMemcpyAsync(keys, src, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, stream);
	}
}

} // namespace b2cu
