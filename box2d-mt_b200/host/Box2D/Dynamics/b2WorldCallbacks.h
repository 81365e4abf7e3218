// Listener / filter interfaces (reference: Box2D/Dynamics/b2WorldCallbacks.h:36-174).
//
// GPU-path semantics:
//  * Begin/EndContactImmediate are invoked on the user thread with threadId 0, from the device's compacted
//    event lists, right after the device step; returning true queues the deferred BeginContact/EndContact,
//    which then run in ascending proxy-id key order, begins before ends (b2ContactManager.cpp:420-433).
//  * PostSolveImmediate / PostSolve are invoked after the device step for every contact the solver handled when
//    b2CudaStepOptions::reportPostSolve is set (Immediate first, deferred ones in key order).
//  * PreSolveImmediate / PreSolve are invoked between the narrow phase and the solver of the device step when
//    b2CudaStepOptions::reportPreSolve is set; b2Contact::SetEnabled(false) there is honoured.
//  * The default b2ContactFilter rule runs on the device; a user subclass is called on the stepping thread for the
//    candidate pairs of every pair search (include/b2cuda.h, b2cuSetPairFilter).
#ifndef B2_WORLD_CALLBACKS_H
#define B2_WORLD_CALLBACKS_H

#include "Box2D/Common/b2Settings.h"

struct b2Manifold;
class b2Contact;
class b2Fixture;

class b2DestructionListener
{
public:
	virtual ~b2DestructionListener() {}
	virtual void SayGoodbye(b2Fixture* fixture) = 0;
};

class b2ContactFilter
{
public:
	virtual ~b2ContactFilter() {}
	/// default rule: group index wins when equal and non-zero, else category/mask test (reference
	/// b2WorldCallbacks.cpp:24-38).  The device evaluates exactly this rule.
	virtual bool ShouldCollide(b2Fixture* fixtureA, b2Fixture* fixtureB, uint32 threadId);
};

struct b2ContactImpulse
{
	float32 normalImpulses[b2_maxManifoldPoints];
	float32 tangentImpulses[b2_maxManifoldPoints];
	int32 count;
};

class b2ContactListener
{
public:
	virtual ~b2ContactListener() {}

	virtual void BeginContact(b2Contact* contact) { B2_NOT_USED(contact); }
	virtual void EndContact(b2Contact* contact) { B2_NOT_USED(contact); }
	virtual void PreSolve(b2Contact* contact, const b2Manifold* oldManifold)
	{
		B2_NOT_USED(contact);
		B2_NOT_USED(oldManifold);
	}
	virtual void PostSolve(b2Contact* contact, const b2ContactImpulse* impulse)
	{
		B2_NOT_USED(contact);
		B2_NOT_USED(impulse);
	}

	/// return true to also receive the deferred callback
	virtual bool BeginContactImmediate(b2Contact* contact, uint32 threadId) = 0;
	virtual bool EndContactImmediate(b2Contact* contact, uint32 threadId) = 0;
	virtual bool PreSolveImmediate(b2Contact* contact, const b2Manifold* oldManifold, uint32 threadId) = 0;
	virtual bool PostSolveImmediate(b2Contact* contact, const b2ContactImpulse* impulse, uint32 threadId) = 0;
};

#endif
