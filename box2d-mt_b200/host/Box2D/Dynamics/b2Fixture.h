// Fixtures (reference: Box2D/Dynamics/b2Fixture.h:33-367).  A fixture is a host handle: its shape is cloned on
// the host, its geometry lives in the device shape table, and its proxy (fat AABB, filter, material) lives in
// the device proxy arrays at m_proxyIndex.
#ifndef B2_FIXTURE_H
#define B2_FIXTURE_H

#include "Box2D/Collision/Shapes/b2Shape.h"

class b2Body;
class b2World;

struct b2Filter
{
	b2Filter() : categoryBits(0x0001), maskBits(0xFFFF), groupIndex(0) {}
	uint16 categoryBits;
	uint16 maskBits;
	int16 groupIndex;
};

struct b2FixtureDef
{
	b2FixtureDef()
		: shape(nullptr), userData(nullptr), friction(0.2f), restitution(0.0f), density(0.0f), isSensor(false),
		  thickShape(false)
	{
	}
	const b2Shape* shape;
	void* userData;
	float32 friction;
	float32 restitution;
	float32 density;
	bool isSensor;
	/// not prone to tunnelling: only generates TOI events against bullets (reference b2Fixture.h:90-93)
	bool thickShape;
	b2Filter filter;
};

class b2Fixture
{
public:
	b2Shape::Type GetType() const { return m_shape->GetType(); }
	b2Shape* GetShape() { return m_shape; }
	const b2Shape* GetShape() const { return m_shape; }
	bool IsSensor() const { return m_isSensor; }
	/// a sensor reports overlaps (Begin/EndContact) but never collides (reference b2Fixture.cpp:222-239)
	void SetSensor(bool sensor);
	void SetFilterData(const b2Filter& filter);
	const b2Filter& GetFilterData() const { return m_filter; }
	void Refilter();
	b2Body* GetBody() { return m_body; }
	const b2Body* GetBody() const { return m_body; }
	b2Fixture* GetNext() { return m_next; }
	const b2Fixture* GetNext() const { return m_next; }
	void* GetUserData() const { return m_userData; }
	void SetUserData(void* data) { m_userData = data; }
	bool TestPoint(const b2Vec2& p) const;
	void GetMassData(b2MassData* massData) const { m_shape->ComputeMass(massData, m_density); }
	void SetDensity(float32 density) { m_density = density; }
	float32 GetDensity() const { return m_density; }
	float32 GetFriction() const { return m_friction; }
	void SetFriction(float32 friction);
	float32 GetRestitution() const { return m_restitution; }
	void SetRestitution(float32 restitution);
	/// swept tight AABB of the (single) child, as last synchronised by the device
	const b2AABB& GetAABB(int32 childIndex) const;
	bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, int32 childIndex) const;
	void SetThickShape(bool flag);
	bool IsThickShape() const { return m_thickShape; }
	/// dense proxy id on the device (-1 while the body is inactive)
	int32 GetProxyIndex() const { return m_proxyIndex; }
	int32 GetProxyCount() const { return m_proxyCount; }

private:
	friend class b2Body;
	friend class b2World;
	friend class b2CudaShardedWorld;
	b2Fixture() {}
	~b2Fixture() { delete m_shape; }

	b2Body* m_body;
	b2Fixture* m_next;
	b2Shape* m_shape;
	float32 m_density, m_friction, m_restitution;
	b2Filter m_filter;
	bool m_isSensor, m_thickShape;
	void* m_userData;
	int32 m_proxyIndex; // first proxy of the fixture; a chain has one proxy per segment, consecutive
	int32 m_proxyCount;
};

#endif
