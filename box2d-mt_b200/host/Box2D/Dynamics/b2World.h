// b2World (reference: Box2D/Dynamics/b2World.h:43-469, b2World.cpp).  Owns bodies and fixtures as host handles
// over struct-of-arrays state (the records of include/b2cuda.h) that is mirrored to the device; Step delegates
// to the executor (b2CudaStepExecutor).  Joints, chain shapes, sensors, world queries and debug draw are outside
// this version of the GPU path and are not declared, so that their use fails at compile time
// instead of being silently ignored.
#ifndef B2_WORLD_H
#define B2_WORLD_H

#include <string>
#include <unordered_map>
#include <vector>

#include "Box2D/Dynamics/b2Body.h"
#include "Box2D/Dynamics/b2TimeStep.h"
#include "Box2D/Dynamics/b2WorldCallbacks.h"
#include "Box2D/Dynamics/Contacts/b2Contact.h"
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
typedef std::vector<b2cuBody, b2MirrorAllocator<b2cuBody> > b2BodyStateArray;
typedef std::vector<b2cuProxy, b2MirrorAllocator<b2cuProxy> > b2ProxyStateArray;

class b2CudaStepExecutor;

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
	int32 GetJointCount() const { return 0; }
	int32 GetContactCount() const { return m_contactCount; }

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
	friend class b2CudaStepExecutor;

	// host mirror of the device state
	b2BodyStateArray m_states;
	std::vector<b2Body*> m_bodies;
	b2ProxyStateArray m_proxies;
	std::vector<b2Fixture*> m_fixtures;
	std::vector<b2cuShape> m_shapes;
	std::unordered_map<std::string, int32> m_shapeLookup;

	int32 InternShape(const b2Shape* shape);
	void MarkBodyDirty(int32 index);
	void MarkProxyDirty(int32 index);
	void RefreshBodies() const;    // device -> host mirror if stale
	void RefreshProxies() const;
	void RefreshContacts();
	void InvalidateSnapshots();
	void DestroyFixtureInternal(b2Body* body, b2Fixture* fixture);
	void RemoveProxies(const std::vector<int32>& proxyIds, const std::vector<int32>& bodyIds);
	void DispatchEvents(b2cuWorld* device);

	// driven by b2CudaStepExecutor::StepWorld
	int32 UploadDirty(b2cuWorld* device);
	int32 AfterDeviceStep(b2cuWorld* device, const b2cuStepInfo& info, bool downloadBodies, bool dispatchEvents,
	                      float32* hostMs = nullptr);
	void MakeContact(b2Contact* c, const b2cuContact& rec);

	b2cuWorld* m_device;          // owned by the executor that last stepped this world
	b2CudaStepExecutor* m_owner;
	bool m_fullUpload;            // everything (including the contact set) must be re-sent
	int32 m_bodiesUploaded, m_proxiesUploaded, m_shapesUploaded;
	int32 m_bodyDirtyLo, m_bodyDirtyHi, m_proxyDirtyLo, m_proxyDirtyHi;
	mutable bool m_bodiesStale, m_proxiesStale;
	bool m_contactsStale;
	std::vector<b2cuContact> m_contactRecords;
	std::vector<b2Contact> m_contacts;
	std::vector<b2ContactEdge*> m_contactHeads;

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
