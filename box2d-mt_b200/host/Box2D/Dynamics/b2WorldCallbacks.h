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
struct b2Vec2;
class b2Contact;
class b2Fixture;

/// What PostSolve reports: the impulses the solver accumulated on each manifold point during the step.
struct b2ContactImpulse
{
	float32 normalImpulses[b2_maxManifoldPoints], tangentImpulses[b2_maxManifoldPoints];
	int32 count;
};

/// Contact events.  The four *Immediate methods are pure: each decides, per contact, whether the matching deferred
/// method is called as well (return true).  On the GPU path all of them run on the stepping thread, threadId 0.
class b2ContactListener
{
public:
	virtual bool BeginContactImmediate(b2Contact* contact, uint32 threadId) = 0;
	virtual bool EndContactImmediate(b2Contact* contact, uint32 threadId) = 0;
	virtual bool PreSolveImmediate(b2Contact* contact, const b2Manifold* oldManifold, uint32 threadId) = 0;
	virtual bool PostSolveImmediate(b2Contact* contact, const b2ContactImpulse* impulse, uint32 threadId) = 0;

	virtual void BeginContact(b2Contact*) {}
	virtual void EndContact(b2Contact*) {}
	virtual void PreSolve(b2Contact*, const b2Manifold* /*oldManifold*/) {}
	virtual void PostSolve(b2Contact*, const b2ContactImpulse*) {}

	virtual ~b2ContactListener() {}
};

/// Decides whether two fixtures whose fat boxes begin to overlap get a contact.  The base class is the default rule
/// (group index wins when equal and non-zero, else the category / mask test; reference b2WorldCallbacks.cpp:24-38),
/// which the device evaluates itself; a subclass is called on the stepping thread.
class b2ContactFilter
{
public:
	virtual bool ShouldCollide(b2Fixture* fixtureA, b2Fixture* fixtureB, uint32 threadId);
	virtual ~b2ContactFilter() {}
};

/// b2World::QueryAABB reports every fixture whose fat box overlaps the query box; return false to stop.
class b2QueryCallback
{
public:
	virtual bool ReportFixture(b2Fixture* fixture) = 0;
	virtual ~b2QueryCallback() {}
};

/// b2World::RayCast reports every fixture the segment hits, in no particular order.  The return value steers the
/// cast: -1 ignore this fixture, 0 stop, a fraction clips the segment there (closest hit: return `fraction`),
/// 1 go on unclipped.
class b2RayCastCallback
{
public:
	virtual float32 ReportFixture(b2Fixture* fixture, const b2Vec2& point, const b2Vec2& normal, float32 fraction) = 0;
	virtual ~b2RayCastCallback() {}
};

/// Told about fixtures that disappear implicitly (their body is destroyed).
class b2Joint;

class b2DestructionListener
{
public:
	/// a joint is about to be destroyed because one of its bodies is (reference b2WorldCallbacks.h:46-48)
	virtual void SayGoodbye(b2Joint* joint) { (void)joint; }
	virtual void SayGoodbye(b2Fixture* fixture) = 0;
	virtual ~b2DestructionListener() {}
};

#endif
