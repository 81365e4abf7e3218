/*
 * oracle/b2o_math.c -- TEST INFRASTRUCTURE (CPU oracle), see b2o_math.h.
 * Restates the device sincos (box2d-mt_b200/csrc/b2cu_math.cuh, b2cu_sincosf) in plain C.
 * Must be compiled with -ffp-contract=off (no fused multiply-add): every operation below is one
 * IEEE-754 binary64 operation, in this order.
 */
#include "b2o_math.h"

#include <stdint.h>
#include <string.h>

static double b2o_bits(uint64_t u)
{
	double d;
	memcpy(&d, &u, sizeof d);
	return d;
}

void b2o_sincosf(float x, float* sOut, float* cOut)
{
	/* fdlibm constants, given as bit patterns so that no decimal parsing is involved */
	const double INV_PIO2 = b2o_bits(0x3FE45F306DC9C883ull); /* 2/pi */
	const double PIO2_1 = b2o_bits(0x3FF921FB54400000ull);   /* first 33 bits of pi/2 */
	const double PIO2_2 = b2o_bits(0x3DD0B4611A600000ull);   /* next 33 bits */
	const double PIO2_3 = b2o_bits(0x3BA3198A2E037073ull);   /* the rest */
	const double MAGIC = 6755399441055744.0;                 /* 1.5 * 2^52 */
	const double S1 = b2o_bits(0xBFC5555555555549ull);
	const double S2 = b2o_bits(0x3F8111111110F8A6ull);
	const double S3 = b2o_bits(0xBF2A01A019C161D5ull);
	const double S4 = b2o_bits(0x3EC71DE357B1FE7Dull);
	const double S5 = b2o_bits(0xBE5AE5E68A2B9CEBull);
	const double S6 = b2o_bits(0x3DE5D93A5ACFD57Cull);
	const double C1 = b2o_bits(0x3FA555555555554Cull);
	const double C2 = b2o_bits(0xBF56C16C16C15177ull);
	const double C3 = b2o_bits(0x3EFA01A019CB1590ull);
	const double C4 = b2o_bits(0xBE927E4F809C52ADull);
	const double C5 = b2o_bits(0x3E21EE9EBDB4B1C4ull);
	const double C6 = b2o_bits(0xBDA8FAE9BE8838D4ull);

	double xd = (double)x;
	double t = xd * INV_PIO2 + MAGIC;
	double k = t - MAGIC;
	long long n = (long long)k;

	double r = xd - k * PIO2_1;
	r = r - k * PIO2_2;
	r = r - k * PIO2_3;

	double z = r * r;

	double ps = S5 + z * S6;
	ps = S4 + z * ps;
	ps = S3 + z * ps;
	ps = S2 + z * ps;
	ps = S1 + z * ps;
	double sr = r + (r * z) * ps;

	double pc = C5 + z * C6;
	pc = C4 + z * pc;
	pc = C3 + z * pc;
	pc = C2 + z * pc;
	pc = C1 + z * pc;
	double cr = (1.0 - 0.5 * z) + (z * z) * pc;

	double s, c;
	switch ((int)(n & 3))
	{
	case 0: s = sr; c = cr; break;
	case 1: s = cr; c = -sr; break;
	case 2: s = -sr; c = -cr; break;
	default: s = -cr; c = sr; break;
	}

	*sOut = (float)s;
	*cOut = (float)c;
}
