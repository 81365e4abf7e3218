// Contacts (reference: Box2D/Dynamics/Contacts/b2Contact.h:36-428).  Contacts live on the device; a b2Contact
// is a host snapshot of one device contact record, materialised for listener callbacks and for
// b2World::GetContactList().  Snapshots are valid until the next Step.
#ifndef B2_CONTACT_H
#define B2_CONTACT_H

#include "Box2D/Collision/b2Collision.h"
#include "Box2D/Dynamics/b2Fixture.h"

class b2Body;
class b2World;

/// friction mixing law (reference b2Contact.h:40-43)
inline float32 b2MixFriction(float32 friction1, float32 friction2) { return b2Sqrt(friction1 * friction2); }
/// restitution mixing law (reference b2Contact.h:47-50)
inline float32 b2MixRestitution(float32 restitution1, float32 restitution2)
{
	return restitution1 > restitution2 ? restitution1 : restitution2;
}

class b2Contact
{
public:
	b2Manifold* GetManifold() { return &m_manifold; }
	const b2Manifold* GetManifold() const { return &m_manifold; }
	void GetWorldManifold(b2WorldManifold* worldManifold) const;
	bool IsTouching() const { return (m_flags & e_touchingFlag) != 0; }
	bool IsEnabled() const { return (m_flags & e_enabledFlag) != 0; }
	/// Meaningful inside b2ContactListener::PreSolve: switches the contact off for the current step
	/// (reference b2Contact.h:113, :299-309); needs b2CudaStepOptions::reportPreSolve
	void SetEnabled(bool flag)
	{
		if (flag) m_flags |= e_enabledFlag;
		else m_flags &= ~(uint32)e_enabledFlag;
	}
	b2Contact* GetNext() { return m_next; }
	const b2Contact* GetNext() const { return m_next; }
	b2Fixture* GetFixtureA() { return m_fixtureA; }
	const b2Fixture* GetFixtureA() const { return m_fixtureA; }
	int32 GetChildIndexA() const { return m_indexA; }
	b2Fixture* GetFixtureB() { return m_fixtureB; }
	const b2Fixture* GetFixtureB() const { return m_fixtureB; }
	int32 GetChildIndexB() const { return m_indexB; }
	float32 GetFriction() const { return m_friction; }
	float32 GetRestitution() const { return m_restitution; }
	float32 GetTangentSpeed() const { return m_tangentSpeed; }
	/// (min proxy id << 32) | max proxy id: the deterministic ordering key of deferred callbacks
	uint64 GetKey() const { return m_key; }

	enum
	{
		e_islandFlag = 0x0001,
		e_touchingFlag = 0x0002,
		e_enabledFlag = 0x0004,
		e_filterFlag = 0x0008,
		e_bulletHitFlag = 0x0010,
		e_toiFlag = 0x0020,
		e_toiCandidateFlag = 0x0040,
		e_inactiveFlag = 0x0080
	};

private:
	friend class b2World;
	uint32 m_flags;
	uint64 m_key;
	b2Fixture* m_fixtureA;
	b2Fixture* m_fixtureB;
	int32 m_indexA, m_indexB; // child (chain segment) of each fixture
	b2Manifold m_manifold;
	float32 m_friction, m_restitution, m_tangentSpeed;
	b2Contact* m_next;
	b2ContactEdge m_nodeA, m_nodeB;
};

#endif
