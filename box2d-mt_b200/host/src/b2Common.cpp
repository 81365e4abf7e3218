// Settings, logging, host sincos, range partitioning.
#include "Box2D/Common/b2Math.h"
#include "Box2D/MT/b2Task.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

b2Version b2_version = {2, 3, 2};
b2Version b2_mtVersion = {0, 1, 0};
const b2Vec2 b2Vec2_zero(0.0f, 0.0f);

void* b2Alloc(int32 size) { return malloc((size_t)size); }
void b2Free(void* mem) { free(mem); }

void b2Log(const char* string, ...)
{
	va_list args;
	va_start(args, string);
	vprintf(string, args);
	va_end(args);
}

namespace
{
inline double Bits(uint64 u)
{
	double d;
	memcpy(&d, &u, sizeof d);
	return d;
}
} // namespace

// Host twin of the device SinCos (box2d-mt_b200/csrc/b2cu_math.cuh): Cody-Waite reduction by pi/2 and the
// fdlibm kernel polynomials in binary64, one rounding to fp32.  This file is compiled with -ffp-contract=off so
// every line is one IEEE operation, as on the device; host transforms (SetTransform, CreateBody) therefore
// match the device's bit for bit.  Replaces the sinf/cosf of the reference's b2Rot (Box2D/Common/b2Math.h:289-299).
void b2SinCos(float32 angle, float32* sOut, float32* cOut)
{
	const double INV_PIO2 = Bits(0x3FE45F306DC9C883ull);
	const double PIO2_1 = Bits(0x3FF921FB54400000ull);
	const double PIO2_2 = Bits(0x3DD0B4611A600000ull);
	const double PIO2_3 = Bits(0x3BA3198A2E037073ull);
	const double MAGIC = 6755399441055744.0;
	const double S[6] = {Bits(0xBFC5555555555549ull), Bits(0x3F8111111110F8A6ull), Bits(0xBF2A01A019C161D5ull),
	                     Bits(0x3EC71DE357B1FE7Dull), Bits(0xBE5AE5E68A2B9CEBull), Bits(0x3DE5D93A5ACFD57Cull)};
	const double C[6] = {Bits(0x3FA555555555554Cull), Bits(0xBF56C16C16C15177ull), Bits(0x3EFA01A019CB1590ull),
	                     Bits(0xBE927E4F809C52ADull), Bits(0x3E21EE9EBDB4B1C4ull), Bits(0xBDA8FAE9BE8838D4ull)};

	double x = (double)angle;
	double t = x * INV_PIO2 + MAGIC;
	double k = t - MAGIC;
	long long n = (long long)k;
	double r = x - k * PIO2_1;
	r = r - k * PIO2_2;
	r = r - k * PIO2_3;
	double z = r * r;

	double ps = S[4] + z * S[5];
	double pc = C[4] + z * C[5];
	for (int i = 3; i >= 0; --i)
	{
		ps = S[i] + z * ps;
		pc = C[i] + z * pc;
	}
	double sr = r + (r * z) * ps;
	double cr = (1.0 - 0.5 * z) + (z * z) * pc;

	double s, c;
	switch ((int)(n & 3))
	{
	case 0: s = sr; c = cr; break;
	case 1: s = cr; c = -sr; break;
	case 2: s = -sr; c = -cr; break;
	default: s = -cr; c = sr; break;
	}
	*sOut = (float32)s;
	*cOut = (float32)c;
}

// reference: Box2D/MT/b2Task.cpp:22-71 -- at most maxOutputRanges contiguous ranges of at least
// minElementsPerRange elements; the remainder is spread one element at a time over the first ranges.
void b2PartitionRange(uint32 begin, uint32 end, uint32 maxOutputRanges, uint32 minElementsPerRange,
                      b2PartitionedRange& output)
{
	uint32 total = end - begin;
	if (maxOutputRanges > b2_maxRangeSubTasks) maxOutputRanges = b2_maxRangeSubTasks;
	if (minElementsPerRange == 0) minElementsPerRange = 1;
	uint32 ranges = total / minElementsPerRange;
	if (ranges > maxOutputRanges) ranges = maxOutputRanges;
	if (ranges == 0) ranges = 1;
	uint32 base = total / ranges;
	uint32 extra = total % ranges;
	uint32 at = begin;
	for (uint32 i = 0; i < ranges; ++i)
	{
		uint32 n = base + (i < extra ? 1u : 0u);
		output.ranges[i].begin = at;
		output.ranges[i].end = at + n;
		at += n;
	}
	output.count = ranges;
}
