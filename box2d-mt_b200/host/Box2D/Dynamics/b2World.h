// b2World (reference: Box2D/Dynamics/b2World.h:43-469, b2World.cpp).  Owns bodies and fixtures as host handles
// over struct-of-arrays state (the records of include/b2cuda.h) that is mirrored to the device; Step delegates
// to the executor (b2CudaStepExecutor).  Debug draw is outside this version of the
// GPU path and is not declared.
#ifndef B2_WORLD_H
#define B2_WORLD_H

#include <string>
#include <unordered_map>
#include <vector>

#include "Box2D/Dynamics/b2Body.h"
#include "Box2D/Dynamics/b2TimeStep.h"
#include "Box2D/Dynamics/b2WorldCallbacks.h"
#include "Box2D/Dynamics/Contacts/b2Contact.h"
#include "Box2D/Dynamics/Joints/b2Joint.h"
#include "Box2D/MT/b2TaskExecutor.h"
#include "b2cuda.h"

/// Allocator of the host mirrors: page-locked memory from the device library, so that the body records move
/// between the mirror and the device in a single DMA transfer.
template <typename T>
struct b2MirrorAllocator
{
	typedef T value_type;
	b2MirrorAllocator() {}
	template <typename U>
	b2MirrorAllocator(const b2MirrorAllocator<U>&) {}
	T* allocate(size_t n) { return static_cast<T*>(b2cuHostAlloc(n * sizeof(T))); }
	void deallocate(T* p, size_t) { b2cuHostFree(p); }
	template <typename U>
	bool operator==(const b2MirrorAllocator<U>&) const { return true; }
	template <typename U>
	bool operator!=(const b2MirrorAllocator<U>&) const { return false; }
};
typedef std::vector<b2cuBodyState, b2MirrorAllocator<b2cuBodyState> > b2BodyStateArray;

/// The fields of b2cuBody that only the host writes (the device clears the forces at the end of a step, as
/// b2World::ClearForces does): they never travel device -> host, so they are kept apart from the b2cuBodyState
/// records that every step copies back.
struct b2BodyProps
{
	float lcx, lcy;
	float fx, fy, torque;
	float invMass, invI;
	float linearDamping, angularDamping, gravityScale;
};

/// One body's record seen as a whole (the field names of b2cuBody) over the two arrays.
struct b2BodyView
{
	float &px, &py, &qs, &qc, &cx, &cy, &a, &c0x, &c0y, &a0, &alpha0, &vx, &vy, &w, &sleepTime;
	uint32_t& flags;
	float &lcx, &lcy, &fx, &fy, &torque, &invMass, &invI, &linearDamping, &angularDamping, &gravityScale;
	b2BodyView(b2cuBodyState& s, b2cuSweepStart& z, b2BodyProps& p)
		: px(s.px), py(s.py), qs(s.qs), qc(s.qc), cx(s.cx), cy(s.cy), a(s.a), c0x(z.c0x), c0y(z.c0y), a0(z.a0),
		  alpha0(z.alpha0), vx(s.vx), vy(s.vy), w(s.w), sleepTime(s.sleepTime), flags(s.flags), lcx(p.lcx), lcy(p.lcy),
		  fx(p.fx), fy(p.fy), torque(p.torque), invMass(p.invMass), invI(p.invI), linearDamping(p.linearDamping),
		  angularDamping(p.angularDamping), gravityScale(p.gravityScale)
	{
	}
	void ToRecord(b2cuBody* out) const
	{
		out->px = px; out->py = py; out->qs = qs; out->qc = qc;
		out->cx = cx; out->cy = cy; out->a = a;
		out->c0x = c0x; out->c0y = c0y; out->a0 = a0; out->alpha0 = alpha0;
		out->lcx = lcx; out->lcy = lcy;
		out->vx = vx; out->vy = vy; out->w = w;
		out->fx = fx; out->fy = fy; out->torque = torque;
		out->invMass = invMass; out->invI = invI;
		out->linearDamping = linearDamping; out->angularDamping = angularDamping; out->gravityScale = gravityScale;
		out->sleepTime = sleepTime;
		out->flags = flags;
	}
};
typedef std::vector<b2cuProxy, b2MirrorAllocator<b2cuProxy> > b2ProxyStateArray;

class b2CudaStepExecutor;
class b2CudaShardedWorld;

class b2World
{
public:
	explicit b2World(const b2Vec2& gravity);
	~b2World();

	void SetDestructionListener(b2DestructionListener* listener) { m_destructionListener = listener; }
	void SetContactFilter(b2ContactFilter* filter) { m_contactFilter = filter; }
	void SetContactListener(b2ContactListener* listener) { m_contactListener = listener; }

	b2Body* CreateBody(const b2BodyDef* def);
	void DestroyBody(b2Body* body);

	/// reference: b2World.h:87-95.  All eleven joint types; an unknown type returns nullptr and sets
	/// GetLastStepStatus() to B2CU_ERR_UNSUPPORTED.  Not while the world is locked.
	b2Joint* CreateJoint(const b2JointDef* def);
	void DestroyJoint(b2Joint* joint);
	b2Joint* GetJointList() { return m_jointList; }
	const b2Joint* GetJointList() const { return m_jointList; }

	/// reference: b2World.h:105-108.  Fails loudly (assert + GetLastStepStatus() != 0) if the executor cannot run
	/// the step on a GPU.
	void Step(float32 timeStep, int32 velocityIterations, int32 positionIterations, b2TaskExecutor& executor);

	void ClearForces();

	b2Body* GetBodyList() { return m_bodyList; }
	const b2Body* GetBodyList() const { return m_bodyList; }
	/// snapshot of the device contact set in key order (valid until the next Step)
	b2Contact* GetContactList();

	void SetAllowSleeping(bool flag);
	bool GetAllowSleeping() const { return m_allowSleep; }
	void SetWarmStarting(bool flag) { m_warmStarting = flag; }
	bool GetWarmStarting() const { return m_warmStarting; }
	void SetContinuousPhysics(bool flag) { m_continuousPhysics = flag; }
	bool GetContinuousPhysics() const { return m_continuousPhysics; }
	void SetSubStepping(bool flag) { m_subStepping = flag; }
	bool GetSubStepping() const { return m_subStepping; }

	int32 GetProxyCount() const { return (int32)m_proxies.size(); }
	int32 GetBodyCount() const { return m_bodyCount; }
	int32 GetJointCount() const { return (int32)m_joints.size(); }
	int32 GetContactCount() const { return m_contactCount; }

	/// every fixture whose fat box overlaps `aabb` (reference b2World.cpp:1752-1758); between steps only
	void QueryAABB(b2QueryCallback* callback, const b2AABB& aabb);
	/// every fixture hit by the segment, with the callback's clipping protocol (reference b2World.cpp:1760-1795)
	void RayCast(b2RayCastCallback* callback, const b2Vec2& point1, const b2Vec2& point2);
	/// move the origin of the world: every position has newOrigin subtracted (reference b2World.cpp:2084-2103)
	void ShiftOrigin(const b2Vec2& newOrigin);
	/// C++ code that rebuilds this world, through b2Log (reference b2World.cpp:2107-2164)
	void Dump();
	void SetGravity(const b2Vec2& gravity) { m_gravity = gravity; }
	b2Vec2 GetGravity() const { return m_gravity; }
	bool IsLocked() const { return m_locked; }
	void SetAutoClearForces(bool flag) { m_clearForces = flag; }
	bool GetAutoClearForces() const { return m_clearForces; }
	const b2Profile& GetProfile() const { return m_profile; }

	/// 0 if the last Step ran, else the b2cuStatus that made it fail
	int32 GetLastStepStatus() const { return m_lastStatus; }

	// ---- struct-of-arrays access (bulk readers such as renderers; rows are dense body / proxy ids) ----
	const b2cuBody* GetBodyStates() const;
	const b2cuProxy* GetProxyStates() const;

private:
	friend class b2Body;
	friend class b2Fixture;
	friend class b2Contact;
	friend class b2Joint;
	friend class b2CudaStepExecutor;
	friend class b2CudaShardedWorld;

	// host mirror of the device state
	b2BodyStateArray m_states;            // device-written part of the bodies (the step's body mirror)
	std::vector<b2cuSweepStart> m_sweepStarts; // sweep starts of the bodies (see RefreshSweepStarts)
	mutable bool m_sweepStartsStale;
	std::vector<b2BodyProps> m_props;     // host-written part
	std::vector<int32> m_forced;          // bodies with a non-zero force/torque on the host side
	std::vector<b2cuBody, b2MirrorAllocator<b2cuBody> > m_uploadRows; // page-locked staging of the dirty rows
	mutable std::vector<b2cuBody> m_records; // whole records, assembled on request (GetBodyStates)
	b2BodyView BodyView(int32 i) { return b2BodyView(m_states[i], m_sweepStarts[i], m_props[i]); }
	/// m_sweep.c0 / a0 / alpha0 of the bodies: device-owned, not part of the step's mirror; current only after this call.
	/// Every host code path that reads or writes them (SetTransform, SetType, mass data, ShiftOrigin, row uploads) makes it.
	void RefreshSweepStarts() const;
	std::vector<b2Body*> m_bodies;
	b2ProxyStateArray m_proxies;
	std::vector<b2Fixture*> m_fixtures;
	std::vector<b2cuShape> m_shapes;
	std::unordered_map<std::string, int32> m_shapeLookup;

	int32 InternShape(const b2Shape* shape, bool chainChild = false);
	void MarkBodyDirty(int32 index);
	void MarkBodyForced(int32 index); // dirty + remembered: the step clears forces on the device, the host follows
	void MarkProxyDirty(int32 index);
	void RefreshBodies() const;    // device -> host mirror if stale
	void RefreshProxies() const;
	void RefreshContacts();
	void RefreshJoints() const;    // device -> joint objects (accumulated impulses) if stale
	void InvalidateSnapshots();
	static int PreSolveThunk(void* user, b2cuWorld* device);
	static int PairFilterThunk(void* user, const b2cuContactKey* keys, int32_t count, uint8_t* keep);
	void DestroyContactsOfBody(int32 bodyIndex);
	void ProxyQuery(const b2AABB* box, const b2Vec2* p1, const b2Vec2* p2, std::vector<int32>& ids);
	void DestroyFixtureInternal(b2Body* body, b2Fixture* fixture);
	void RemoveProxies(const std::vector<int32>& proxyIds, const std::vector<int32>& bodyIds);
	void DispatchEvents(b2cuWorld* device);
	void DispatchPostSolve(b2cuWorld* device);

	// driven by b2CudaStepExecutor::StepWorld
	int32 UploadDirty(b2cuWorld* device);
	int32 AfterDeviceStep(b2cuWorld* device, const b2cuStepInfo& info, bool downloadBodies, bool dispatchEvents,
	                      float32* hostMs = nullptr, bool reportPostSolve = false);
	void MakeContact(b2Contact* c, const b2cuContact& rec);

	b2cuWorld* m_device;          // owned by the executor that last stepped this world
	b2CudaStepExecutor* m_owner;
	bool m_fullUpload;            // everything (including the contact set) must be re-sent
	int32 m_bodiesUploaded, m_proxiesUploaded, m_shapesUploaded;
	int32 m_bodyDirtyLo, m_bodyDirtyHi, m_proxyDirtyLo, m_proxyDirtyHi;
	int32 m_forceDirtyLo, m_forceDirtyHi; // rows whose force / torque alone changed (b2Body::ApplyForce on awake bodies)
	std::vector<float> m_forceRows;       // staging of b2cuSetBodyForces
	// event dispatch scratch, kept between steps (growing only: no allocation, no construction per step)
	std::vector<b2cuContactKey> m_eventKeys[2];
	std::vector<b2cuContact> m_eventRecs[2];
	std::vector<b2Contact> m_eventContacts[2];
	std::vector<char> m_eventDeferred[2];
	mutable bool m_bodiesStale, m_proxiesStale;
	bool m_contactsStale;
	std::vector<b2cuContact> m_contactRecords;
	std::vector<b2Contact> m_contacts;
	std::vector<b2ContactEdge*> m_contactHeads;

	std::vector<b2Joint*> m_joints;   // creation order = rows of the device joint table
	b2Joint* m_jointList;
	bool m_jointsDirty;               // the table must be sent again before the next step
	mutable bool m_jointsStale;       // the device holds newer impulses than the joint objects
	b2Body* m_bodyList;
	int32 m_bodyCount;
	int32 m_contactCount;
	b2Vec2 m_gravity;
	bool m_allowSleep, m_warmStarting, m_continuousPhysics, m_subStepping, m_clearForces, m_locked, m_newFixture;
	float32 m_inv_dt0;
	b2DestructionListener* m_destructionListener;
	b2ContactFilter* m_contactFilter;
	b2ContactListener* m_contactListener;
	b2Profile m_profile;
	int32 m_lastStatus;
};

#endif
