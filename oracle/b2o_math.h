/*
 * oracle/b2o_math.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * b2o_sincosf restates, in plain C, the fp32 sin/cos that the device kernels use for b2Rot::Set
 * (reference: Box2D/Common/b2Math.h:289-299, which calls libm sinf/cosf).  glibc's sinf and CUDA's sinf
 * differ in the last bit for some arguments, and a 1-ulp difference can flip a fat-AABB containment test
 * (Box2D/Collision/b2DynamicTree.cpp:136-139) and so change the contact set.  Both sides therefore use this
 * one algorithm, built only from IEEE-754 double +,-,* and conversions, which round identically under gcc
 * (-ffp-contract=off) and nvcc (__dmul_rn/__dadd_rn):
 *
 *   t = x * 2/pi + 1.5*2^52 ; k = t - 1.5*2^52 ; n = (int)k                  (round-to-nearest integer)
 *   r = ((x - k*PIO2_1) - k*PIO2_2) - k*PIO2_3                              (Cody-Waite, 33+33+53 bits)
 *   sin r, cos r by the fdlibm kernel polynomials (degree 13 / 14 in r)
 *   rotate by quadrant n & 3, round the double result to float.
 *
 * The compiled reference in oracle/_ref is linked with sinf/cosf/sincosf defined in terms of this function
 * (see oracle/ref_harness.cpp), so oracle and device agree bit for bit on every transform.
 */
#ifndef B2O_MATH_H
#define B2O_MATH_H

#ifdef __cplusplus
extern "C" {
#endif

void b2o_sincosf(float x, float* s, float* c);

#ifdef __cplusplus
}
#endif

#endif
