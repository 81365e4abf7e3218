// b2cu_prims.cuh -- data-parallel primitives used by every phase: exclusive scan, stable stream compaction,
// LSD radix sort of 64-bit keys.  They replace the reference's per-thread buffers + b2ThreadDataSorter
// (Box2D/MT/b2ThreadDataSorter.h:262-380): every list the step produces (events, new pairs, constraints) is
// emitted in a deterministic total order.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b2cu
{

// Programmatic dependent launch: the step is ~50 short kernels in a row on one stream, and between two of them the GPU
// idles for ~3 us (drain, launch, ramp).  Kernels launched with the programmatic-serialization attribute (LaunchPdl) are
// scheduled while their predecessor is still running and wait HERE, first thing, for it to complete with its memory
// operations visible (measured on B200: 6.2 -> 3.3 us per dependent launch of a short kernel).  In a kernel launched the
// ordinary way the instruction does nothing.  Every kernel executes it unconditionally: a kernel that returned without
// it could complete before its predecessor and let ITS successor start too early.
#ifdef __CUDACC__
__device__ __forceinline__ void GridDependencyWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern bool g_pdl; // B2CU_PDL=0 turns the attribute off (ordinary stream order)
template <typename... Params, typename... Args>
inline cudaError_t LaunchPdl(void (*kernel)(Params...), dim3 grid, dim3 block, cudaStream_t stream, Args&&... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = 0;
	cfg.stream = stream;
	cudaLaunchAttribute attribute[1];
	attribute[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attribute[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
	cfg.attrs = attribute;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}
#endif

struct PrimScratch
{
	unsigned long long* scanState = nullptr; // ticket + one look-back word per tile
	int* radixHist = nullptr;  // 256 * numBlocks
	uint64_t* radixAlt = nullptr;
	int* compactPos = nullptr; // capacity
	int capacity = 0;
	int radixBlocks = 0;
};

// number of kernels launched by the primitives since the last reset (for b2cuStepInfo.kernelLaunches)
extern int g_primLaunches;
// optional hook called after every primitive kernel launch (tracing)
extern void (*g_primTraceHook)(const char* name, cudaStream_t stream);

cudaError_t PrimScratchAlloc(PrimScratch* s, int capacity);
void PrimScratchFree(PrimScratch* s);

// out[i] = sum(in[0..i)), in and out may alias; *total (device) = sum of all, if total != nullptr
void ExclusiveScan(PrimScratch* s, const int* in, int* out, int n, int* total, cudaStream_t stream);

// out[i] = number of j < i with (flags[j] & mask) == 0
void ExclusiveScanNotMask(PrimScratch* s, const uint32_t* flags, uint32_t mask, int* out, int n, int* total,
                          cudaStream_t stream);

// outIdx = ascending indices i with flags[i] != 0; *outCount (device) = how many
void CompactFlags(PrimScratch* s, const int* flags, int n, int* outIdx, int* outCount, cudaStream_t stream);
// same with a bit mask test: (flags[i] & mask) != 0
void CompactMask(PrimScratch* s, const uint32_t* flags, uint32_t mask, int n, int* outIdx, int* outCount,
                 cudaStream_t stream);

// stable ascending sort on bits [beginBit, endBit) of the keys; result ends up in `keys`
void RadixSort64(PrimScratch* s, uint64_t* keys, int n, int beginBit, int endBit, cudaStream_t stream);

// full 64-bit ascending sort for small / medium lists: tile bitonic sort + merge rounds (few launches)
void SortSmall64(PrimScratch* s, uint64_t* keys, int n, cudaStream_t stream);
#define B2CU_SMALL_SORT_MAX (1 << 18)

} // namespace b2cu
