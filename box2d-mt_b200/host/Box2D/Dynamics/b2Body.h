// Rigid bodies (reference: Box2D/Dynamics/b2Body.h:36-963).  A b2Body is a stable host handle onto row
// m_index of the world's struct-of-arrays body state, which is mirrored to the device.  Accessors read the
// host mirror (refreshed after every device step); mutators edit the mirror and mark the row for upload.
#ifndef B2_BODY_H
#define B2_BODY_H

#include "Box2D/Dynamics/b2Fixture.h"

class b2Contact;
class b2World;

enum b2BodyType { b2_staticBody = 0, b2_kinematicBody, b2_dynamicBody };

struct b2BodyDef
{
	b2BodyDef()
		: type(b2_staticBody), position(0.0f, 0.0f), angle(0.0f), linearVelocity(0.0f, 0.0f), angularVelocity(0.0f),
		  linearDamping(0.0f), angularDamping(0.0f), allowSleep(true), awake(true), fixedRotation(false), bullet(false),
		  active(true), userData(nullptr), gravityScale(1.0f)
	{
	}
	b2BodyType type;
	b2Vec2 position;
	float32 angle;
	b2Vec2 linearVelocity;
	float32 angularVelocity;
	float32 linearDamping;
	float32 angularDamping;
	bool allowSleep;
	bool awake;
	bool fixedRotation;
	bool bullet;
	bool active;
	void* userData;
	float32 gravityScale;
};

/// adjacency record of the contact graph (reference: Box2D/Dynamics/Contacts/b2Contact.h:79-90); valid until
/// the next Step
struct b2JointEdge;

struct b2ContactEdge
{
	b2Body* other;
	b2Contact* contact;
	b2ContactEdge* prev;
	b2ContactEdge* next;
};

class b2Body
{
public:
	b2Fixture* CreateFixture(const b2FixtureDef* def);
	b2Fixture* CreateFixture(const b2Shape* shape, float32 density);
	void DestroyFixture(b2Fixture* fixture);

	void SetTransform(const b2Vec2& position, float32 angle);
	const b2Transform& GetTransform() const;
	const b2Vec2& GetPosition() const;
	float32 GetAngle() const;
	const b2Vec2& GetWorldCenter() const;
	const b2Vec2& GetLocalCenter() const;
	void SetLinearVelocity(const b2Vec2& v);
	const b2Vec2& GetLinearVelocity() const;
	void SetAngularVelocity(float32 omega);
	float32 GetAngularVelocity() const;
	void ApplyForce(const b2Vec2& force, const b2Vec2& point, bool wake);
	void ApplyForceToCenter(const b2Vec2& force, bool wake);
	void ApplyTorque(float32 torque, bool wake);
	void ApplyLinearImpulse(const b2Vec2& impulse, const b2Vec2& point, bool wake);
	void ApplyLinearImpulseToCenter(const b2Vec2& impulse, bool wake);
	void ApplyAngularImpulse(float32 impulse, bool wake);
	float32 GetMass() const { return m_mass; }
	float32 GetInertia() const;
	void GetMassData(b2MassData* data) const;
	void SetMassData(const b2MassData* data);
	void ResetMassData();
	b2Vec2 GetWorldPoint(const b2Vec2& localPoint) const { return b2Mul(GetTransform(), localPoint); }
	b2Vec2 GetWorldVector(const b2Vec2& localVector) const { return b2Mul(GetTransform().q, localVector); }
	b2Vec2 GetLocalPoint(const b2Vec2& worldPoint) const { return b2MulT(GetTransform(), worldPoint); }
	b2Vec2 GetLocalVector(const b2Vec2& worldVector) const { return b2MulT(GetTransform().q, worldVector); }
	b2Vec2 GetLinearVelocityFromWorldPoint(const b2Vec2& worldPoint) const;
	b2Vec2 GetLinearVelocityFromLocalPoint(const b2Vec2& localPoint) const;
	float32 GetLinearDamping() const;
	void SetLinearDamping(float32 linearDamping);
	float32 GetAngularDamping() const;
	void SetAngularDamping(float32 angularDamping);
	float32 GetGravityScale() const;
	void SetGravityScale(float32 scale);
	void SetType(b2BodyType type);
	b2BodyType GetType() const;
	void SetBullet(bool flag);
	bool IsBullet() const;
	void SetSleepingAllowed(bool flag);
	bool IsSleepingAllowed() const;
	void SetAwake(bool flag);
	bool IsAwake() const;
	void SetActive(bool flag);
	bool IsActive() const;
	void SetFixedRotation(bool flag);
	bool IsFixedRotation() const;
	b2Fixture* GetFixtureList() { return m_fixtureList; }
	const b2Fixture* GetFixtureList() const { return m_fixtureList; }
	/// contacts attached to this body (built from the device contact set on first use after a step)
	b2ContactEdge* GetContactList();
	/// joints attached to this body (reference b2Body.h:369-373)
	b2JointEdge* GetJointList() { return m_jointList; }
	const b2JointEdge* GetJointList() const { return m_jointList; }
	b2Body* GetNext() { return m_next; }
	const b2Body* GetNext() const { return m_next; }
	void* GetUserData() const { return m_userData; }
	void SetUserData(void* data) { m_userData = data; }
	b2World* GetWorld() { return m_world; }
	const b2World* GetWorld() const { return m_world; }
	/// dense body id (row of the device body arrays)
	int32 GetIndex() const { return m_index; }

private:
	friend class b2World;
	friend class b2Fixture;
	friend class b2Contact;
	friend class b2Joint;
	friend class b2CudaShardedWorld;
	b2Body() {}
	~b2Body() {}
	void SynchronizeProxies(const b2Transform& xf1, const b2Transform& xf2);

	b2World* m_world;
	int32 m_index;
	float32 m_mass, m_I;
	b2Fixture* m_fixtureList;
	int32 m_fixtureCount;
	b2JointEdge* m_jointList;
	b2Body* m_prev;
	b2Body* m_next;
	void* m_userData;
};

#endif
