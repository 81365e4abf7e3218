// reference: Box2D/Dynamics/b2TimeStep.h:25-40 (b2Profile).  The 13 fields are filled from CUDA-event timings
// of the device phases (b2cuStepInfo); `locking` is always 0 (there is no thread pool to lock).
#ifndef B2_TIME_STEP_H
#define B2_TIME_STEP_H

#include "Box2D/Common/b2Math.h"

struct b2Profile
{
	float32 step;
	float32 collide;
	float32 solve;
	float32 solveTraversal;
	float32 solveInit;
	float32 solveVelocity;
	float32 solvePosition;
	float32 solveTOI;
	float32 broadphase;
	float32 broadphaseSyncFixtures;
	float32 broadphaseFindContacts;
	float32 locking;
	float32 reserved;
};

#endif
