// b2Body / b2Fixture: host handles over the struct-of-arrays state of b2World.
// Semantics follow Box2D/Dynamics/b2Body.cpp and b2Fixture.cpp of the reference; storage does not (state lives
// in b2World::m_states / m_proxies, the records that are mirrored to the device).
#include "Box2D/Dynamics/b2Body.h"
#include "Box2D/Collision/Shapes/b2ChainShape.h"
#include "Box2D/Collision/Shapes/b2EdgeShape.h"
#include "Box2D/Dynamics/b2World.h"

namespace
{
inline const b2Vec2& AsVec2(const float& x) { return reinterpret_cast<const b2Vec2&>(x); }
inline b2Vec2& AsVec2(float& x) { return reinterpret_cast<b2Vec2&>(x); }
} // namespace

#define B2_STATE() (m_world->BodyView(m_index))

// ---- accessors -------------------------------------------------------------------------------------------

const b2Transform& b2Body::GetTransform() const
{
	m_world->RefreshBodies();
	// b2cuBody starts with px, py, qs, qc: the layout of b2Transform
	return reinterpret_cast<const b2Transform&>(B2_STATE().px);
}
const b2Vec2& b2Body::GetPosition() const { return GetTransform().p; }
float32 b2Body::GetAngle() const
{
	m_world->RefreshBodies();
	return B2_STATE().a;
}
const b2Vec2& b2Body::GetWorldCenter() const
{
	m_world->RefreshBodies();
	return AsVec2(B2_STATE().cx);
}
const b2Vec2& b2Body::GetLocalCenter() const { return AsVec2(B2_STATE().lcx); }
const b2Vec2& b2Body::GetLinearVelocity() const
{
	m_world->RefreshBodies();
	return AsVec2(B2_STATE().vx);
}
float32 b2Body::GetAngularVelocity() const
{
	m_world->RefreshBodies();
	return B2_STATE().w;
}
float32 b2Body::GetInertia() const
{
	b2BodyView s = B2_STATE();
	return m_I + m_mass * b2Dot(AsVec2(s.lcx), AsVec2(s.lcx));
}
void b2Body::GetMassData(b2MassData* data) const
{
	data->mass = m_mass;
	data->I = GetInertia();
	data->center = GetLocalCenter();
}
b2Vec2 b2Body::GetLinearVelocityFromWorldPoint(const b2Vec2& worldPoint) const
{
	return GetLinearVelocity() + b2Cross(GetAngularVelocity(), worldPoint - GetWorldCenter());
}
b2Vec2 b2Body::GetLinearVelocityFromLocalPoint(const b2Vec2& localPoint) const
{
	return GetLinearVelocityFromWorldPoint(GetWorldPoint(localPoint));
}
float32 b2Body::GetLinearDamping() const { return B2_STATE().linearDamping; }
float32 b2Body::GetAngularDamping() const { return B2_STATE().angularDamping; }
float32 b2Body::GetGravityScale() const { return B2_STATE().gravityScale; }
b2BodyType b2Body::GetType() const { return (b2BodyType)(B2_STATE().flags & B2CU_BODY_TYPE_MASK); }
bool b2Body::IsBullet() const { return (B2_STATE().flags & B2CU_BODY_BULLET) != 0; }
bool b2Body::IsSleepingAllowed() const { return (B2_STATE().flags & B2CU_BODY_AUTOSLEEP) != 0; }
bool b2Body::IsAwake() const
{
	m_world->RefreshBodies();
	return (B2_STATE().flags & B2CU_BODY_AWAKE) != 0;
}
bool b2Body::IsActive() const { return (B2_STATE().flags & B2CU_BODY_ACTIVE) != 0; }
bool b2Body::IsFixedRotation() const { return (B2_STATE().flags & B2CU_BODY_FIXED_ROTATION) != 0; }

// ---- mutators (edit the mirror, mark the row for upload) -----------------------------------------------------

void b2Body::SetLinearVelocity(const b2Vec2& v)
{
	if (GetType() == b2_staticBody) return;
	if (b2Dot(v, v) > 0.0f) SetAwake(true);
	m_world->RefreshBodies();
	AsVec2(B2_STATE().vx) = v;
	m_world->MarkBodyDirty(m_index);
}

void b2Body::SetAngularVelocity(float32 omega)
{
	if (GetType() == b2_staticBody) return;
	if (omega * omega > 0.0f) SetAwake(true);
	m_world->RefreshBodies();
	B2_STATE().w = omega;
	m_world->MarkBodyDirty(m_index);
}

void b2Body::ApplyForce(const b2Vec2& force, const b2Vec2& point, bool wake)
{
	m_world->RefreshBodies();
	if (GetType() != b2_dynamicBody) return;
	if (wake && !IsAwake()) SetAwake(true);
	if (!IsAwake()) return;
	b2BodyView s = B2_STATE();
	AsVec2(s.fx) += force;
	s.torque += b2Cross(point - AsVec2(s.cx), force);
	m_world->MarkBodyForced(m_index);
}

void b2Body::ApplyForceToCenter(const b2Vec2& force, bool wake)
{
	if (GetType() != b2_dynamicBody) return;
	if (wake && !IsAwake()) SetAwake(true);
	if (!IsAwake()) return;
	AsVec2(B2_STATE().fx) += force;
	m_world->MarkBodyForced(m_index);
}

void b2Body::ApplyTorque(float32 torque, bool wake)
{
	if (GetType() != b2_dynamicBody) return;
	if (wake && !IsAwake()) SetAwake(true);
	if (!IsAwake()) return;
	B2_STATE().torque += torque;
	m_world->MarkBodyForced(m_index);
}

void b2Body::ApplyLinearImpulse(const b2Vec2& impulse, const b2Vec2& point, bool wake)
{
	m_world->RefreshBodies();
	if (GetType() != b2_dynamicBody) return;
	if (wake && !IsAwake()) SetAwake(true);
	if (!IsAwake()) return;
	b2BodyView s = B2_STATE();
	AsVec2(s.vx) += s.invMass * impulse;
	s.w += s.invI * b2Cross(point - AsVec2(s.cx), impulse);
	m_world->MarkBodyDirty(m_index);
}

void b2Body::ApplyLinearImpulseToCenter(const b2Vec2& impulse, bool wake)
{
	// the device owns position, velocity, sleep time and flags between steps: edit the current row, not a stale one
	m_world->RefreshBodies();
	if (GetType() != b2_dynamicBody) return;
	if (wake && !IsAwake()) SetAwake(true);
	if (!IsAwake()) return;
	b2BodyView s = B2_STATE();
	AsVec2(s.vx) += s.invMass * impulse;
	m_world->MarkBodyDirty(m_index);
}

void b2Body::ApplyAngularImpulse(float32 impulse, bool wake)
{
	// the device owns position, velocity, sleep time and flags between steps: edit the current row, not a stale one
	m_world->RefreshBodies();
	if (GetType() != b2_dynamicBody) return;
	if (wake && !IsAwake()) SetAwake(true);
	if (!IsAwake()) return;
	b2BodyView s = B2_STATE();
	s.w += s.invI * impulse;
	m_world->MarkBodyDirty(m_index);
}

void b2Body::SetLinearDamping(float32 d)
{
	// the device owns position, velocity, sleep time and flags between steps: edit the current row, not a stale one
	m_world->RefreshBodies();
	B2_STATE().linearDamping = d;
	m_world->MarkBodyDirty(m_index);
}
void b2Body::SetAngularDamping(float32 d)
{
	// the device owns position, velocity, sleep time and flags between steps: edit the current row, not a stale one
	m_world->RefreshBodies();
	B2_STATE().angularDamping = d;
	m_world->MarkBodyDirty(m_index);
}
void b2Body::SetGravityScale(float32 scale)
{
	// the device owns position, velocity, sleep time and flags between steps: edit the current row, not a stale one
	m_world->RefreshBodies();
	B2_STATE().gravityScale = scale;
	m_world->MarkBodyDirty(m_index);
}

// reference b2Body.h:690-718: waking resets the sleep timer; sleeping also zeroes velocity and forces.  The
// e_inactiveFlag of the attached contacts is a function of the awake flags on the device, so there is nothing
// to recalculate here.
void b2Body::SetAwake(bool flag)
{
	if (m_world->IsLocked()) return;
	m_world->RefreshBodies();
	b2BodyView s = B2_STATE();
	if (flag)
	{
		s.flags |= B2CU_BODY_AWAKE;
		s.sleepTime = 0.0f;
	}
	else
	{
		s.flags &= ~(uint32)B2CU_BODY_AWAKE;
		s.sleepTime = 0.0f;
		s.vx = s.vy = s.w = 0.0f;
		s.fx = s.fy = s.torque = 0.0f;
	}
	m_world->MarkBodyDirty(m_index);
}

void b2Body::SetSleepingAllowed(bool flag)
{
	// the device owns position, velocity, sleep time and flags between steps: edit the current row, not a stale one
	m_world->RefreshBodies();
	if (flag)
	{
		B2_STATE().flags |= B2CU_BODY_AUTOSLEEP;
	}
	else
	{
		B2_STATE().flags &= ~(uint32)B2CU_BODY_AUTOSLEEP;
		SetAwake(true);
	}
	m_world->MarkBodyDirty(m_index);
}

void b2Body::SetBullet(bool flag)
{
	// the device re-evaluates the time-of-impact candidacy of the body's contacts when the uploaded row changes the
	// flag (b2ContactManager::RecalculateToiCandidacy, reference b2Body.h / b2ContactManager.cpp:566-640)
	m_world->RefreshBodies();
	if (flag) B2_STATE().flags |= B2CU_BODY_BULLET;
	else B2_STATE().flags &= ~(uint32)B2CU_BODY_BULLET;
	m_world->MarkBodyDirty(m_index);
}

// reference b2Body.cpp:496-544.  The proxies keep their slots (ids) and are only flagged out of the broad-phase, so that
// a body that comes back has the ids it had: the fixture A / fixture B roles of its future contacts follow the ids.
void b2Body::SetActive(bool flag)
{
	if (m_world->IsLocked() || flag == IsActive()) return;
	m_world->RefreshBodies();
	m_world->RefreshProxies();
	b2BodyView s = B2_STATE();
	if (flag)
	{
		s.flags |= B2CU_BODY_ACTIVE;
		// b2Fixture::CreateProxies at the current transform: tight box, fat box = tight box + margin, in the move buffer
		// (no e_newFixture: the pairs are found by the pair search at the END of the next step)
		const b2Transform& xf = reinterpret_cast<const b2Transform&>(s.px);
		for (b2Fixture* f = m_fixtureList; f; f = f->m_next)
		{
			for (int32 child = 0; child < f->m_proxyCount; ++child)
			{
				b2cuProxy& p = m_world->m_proxies[f->m_proxyIndex + child];
				b2AABB aabb;
				f->m_shape->ComputeAABB(&aabb, xf, child);
				p.aabb[0] = aabb.lowerBound.x;
				p.aabb[1] = aabb.lowerBound.y;
				p.aabb[2] = aabb.upperBound.x;
				p.aabb[3] = aabb.upperBound.y;
				p.fat[0] = aabb.lowerBound.x - b2_aabbExtension;
				p.fat[1] = aabb.lowerBound.y - b2_aabbExtension;
				p.fat[2] = aabb.upperBound.x + b2_aabbExtension;
				p.fat[3] = aabb.upperBound.y + b2_aabbExtension;
				p.flags = (uint16)((p.flags & ~(uint16)B2CU_PROXY_INACTIVE) | B2CU_PROXY_MOVED);
				m_world->MarkProxyDirty(f->m_proxyIndex + child);
			}
		}
	}
	else
	{
		s.flags &= ~(uint32)B2CU_BODY_ACTIVE;
		for (b2Fixture* f = m_fixtureList; f; f = f->m_next)
		{
			for (int32 child = 0; child < f->m_proxyCount; ++child)
			{
				b2cuProxy& p = m_world->m_proxies[f->m_proxyIndex + child];
				p.flags = (uint16)((p.flags & ~(uint16)(B2CU_PROXY_MOVED | B2CU_PROXY_NEW)) | B2CU_PROXY_INACTIVE);
				m_world->MarkProxyDirty(f->m_proxyIndex + child);
			}
		}
		// destroy the attached contacts (EndContact for the touching ones)
		m_world->DestroyContactsOfBody(m_index);
	}
	m_world->MarkBodyDirty(m_index);
}

void b2Body::SetFixedRotation(bool flag)
{
	if (flag == IsFixedRotation()) return;
	if (flag) B2_STATE().flags |= B2CU_BODY_FIXED_ROTATION;
	else B2_STATE().flags &= ~(uint32)B2CU_BODY_FIXED_ROTATION;
	m_world->RefreshBodies();
	B2_STATE().w = 0.0f;
	ResetMassData();
}

// reference b2Body.cpp:118-188
void b2Body::SetType(b2BodyType type)
{
	if (m_world->IsLocked() || type == GetType()) return;
	m_world->RefreshBodies();
	m_world->RefreshSweepStarts();
	{
		b2BodyView s = B2_STATE();
		s.flags = (s.flags & ~(uint32)B2CU_BODY_TYPE_MASK) | (uint32)type;
	}
	ResetMassData();
	if (type == b2_staticBody)
	{
		b2BodyView s = B2_STATE();
		s.vx = s.vy = s.w = 0.0f;
		s.a0 = s.a;
		s.c0x = s.cx;
		s.c0y = s.cy;
		// SynchronizeFixtures with xf1 from the (now equal) sweep start
		b2Transform xf1;
		xf1.q.Set(s.a0);
		xf1.p = b2Vec2(s.c0x, s.c0y) - b2Mul(xf1.q, AsVec2(s.lcx));
		SynchronizeProxies(xf1, GetTransform());
	}
	SetAwake(true);
	{
		b2BodyView s = B2_STATE();
		s.fx = s.fy = s.torque = 0.0f;
	}
	m_world->MarkBodyDirty(m_index);

	// delete the attached contacts, then touch the proxies so that new contacts are created when appropriate
	m_world->DestroyContactsOfBody(m_index);
	m_world->RefreshProxies();
	for (b2Fixture* f = m_fixtureList; f; f = f->m_next)
	{
		if (f->m_proxyIndex < 0) continue;
		for (int32 child = 0; child < f->m_proxyCount; ++child)
		{
			m_world->m_proxies[f->m_proxyIndex + child].flags |= B2CU_PROXY_MOVED;
			m_world->MarkProxyDirty(f->m_proxyIndex + child);
		}
	}
}

// reference b2Body.cpp:449-473
void b2Body::SetTransform(const b2Vec2& position, float32 angle)
{
	if (m_world->IsLocked()) return;
	m_world->RefreshBodies();
	m_world->RefreshSweepStarts();
	b2BodyView s = B2_STATE();
	b2Transform xf;
	xf.q.Set(angle);
	xf.p = position;
	s.px = xf.p.x;
	s.py = xf.p.y;
	s.qs = xf.q.s;
	s.qc = xf.q.c;
	b2Vec2 c = b2Mul(xf, AsVec2(s.lcx));
	s.cx = s.c0x = c.x;
	s.cy = s.c0y = c.y;
	s.a = s.a0 = angle;
	m_world->MarkBodyDirty(m_index);
	SynchronizeProxies(xf, xf);
}

// b2Fixture::Synchronize + b2BroadPhase::MoveProxy (reference b2Fixture.cpp:153-176, b2DynamicTree.cpp:130-174)
void b2Body::SynchronizeProxies(const b2Transform& xf1, const b2Transform& xf2)
{
	if (m_fixtureList == nullptr) return;
	m_world->RefreshProxies();
	for (b2Fixture* f = m_fixtureList; f; f = f->m_next)
	{
		if (f->m_proxyIndex < 0) continue;
		for (int32 child = 0; child < f->m_proxyCount; ++child)
		{
		b2cuProxy& p = m_world->m_proxies[f->m_proxyIndex + child];
		b2AABB a1, a2, ab;
		f->m_shape->ComputeAABB(&a1, xf1, child);
		f->m_shape->ComputeAABB(&a2, xf2, child);
		ab.Combine(a1, a2);
		p.aabb[0] = ab.lowerBound.x;
		p.aabb[1] = ab.lowerBound.y;
		p.aabb[2] = ab.upperBound.x;
		p.aabb[3] = ab.upperBound.y;
		b2AABB fat;
		fat.lowerBound.Set(p.fat[0], p.fat[1]);
		fat.upperBound.Set(p.fat[2], p.fat[3]);
		if (!fat.Contains(ab))
		{
			b2Vec2 d = b2_aabbMultiplier * (xf2.p - xf1.p);
			b2Vec2 r(b2_aabbExtension, b2_aabbExtension);
			b2AABB b;
			b.lowerBound = ab.lowerBound - r;
			b.upperBound = ab.upperBound + r;
			if (d.x < 0.0f) b.lowerBound.x += d.x;
			else b.upperBound.x += d.x;
			if (d.y < 0.0f) b.lowerBound.y += d.y;
			else b.upperBound.y += d.y;
			p.fat[0] = b.lowerBound.x;
			p.fat[1] = b.lowerBound.y;
			p.fat[2] = b.upperBound.x;
			p.fat[3] = b.upperBound.y;
			p.flags |= B2CU_PROXY_MOVED; // buffered move: its pairs are found at the end of the next step
		}
		m_world->MarkProxyDirty(f->m_proxyIndex + child);
		}
	}
}

// reference b2Body.cpp:306-385
void b2Body::ResetMassData()
{
	m_world->RefreshBodies();
	m_world->RefreshSweepStarts();
	if (m_jointList)
	{
		// a mouse joint's row carries this body's mass
		m_world->RefreshJoints();
		m_world->m_jointsDirty = true;
	}
	b2BodyView s = B2_STATE();
	m_mass = 0.0f;
	s.invMass = 0.0f;
	m_I = 0.0f;
	s.invI = 0.0f;
	s.lcx = s.lcy = 0.0f;
	m_world->MarkBodyDirty(m_index);

	b2Transform xf = reinterpret_cast<const b2Transform&>(s.px);
	if (GetType() != b2_dynamicBody)
	{
		s.c0x = s.cx = xf.p.x;
		s.c0y = s.cy = xf.p.y;
		s.a0 = s.a;
		return;
	}

	// fixtures newest first, the order of the reference's intrusive list: the float sums depend on it
	b2Vec2 localCenter = b2Vec2_zero;
	for (b2Fixture* f = m_fixtureList; f; f = f->m_next)
	{
		if (f->m_density == 0.0f) continue;
		b2MassData md;
		f->GetMassData(&md);
		m_mass += md.mass;
		localCenter += md.mass * md.center;
		m_I += md.I;
	}

	if (m_mass > 0.0f)
	{
		s.invMass = 1.0f / m_mass;
		localCenter *= s.invMass;
	}
	else
	{
		m_mass = 1.0f;
		s.invMass = 1.0f;
	}

	if (m_I > 0.0f && (s.flags & B2CU_BODY_FIXED_ROTATION) == 0)
	{
		m_I -= m_mass * b2Dot(localCenter, localCenter);
		s.invI = 1.0f / m_I;
	}
	else
	{
		m_I = 0.0f;
		s.invI = 0.0f;
	}

	b2Vec2 oldCenter(s.cx, s.cy);
	s.lcx = localCenter.x;
	s.lcy = localCenter.y;
	b2Vec2 c = b2Mul(xf, localCenter);
	s.c0x = s.cx = c.x;
	s.c0y = s.cy = c.y;
	AsVec2(s.vx) += b2Cross(s.w, c - oldCenter);
}

// reference b2Body.cpp:387-426
void b2Body::SetMassData(const b2MassData* massData)
{
	if (m_world->IsLocked() || GetType() != b2_dynamicBody) return;
	m_world->RefreshBodies();
	m_world->RefreshSweepStarts();
	b2BodyView s = B2_STATE();
	s.invMass = 0.0f;
	m_I = 0.0f;
	s.invI = 0.0f;
	m_mass = massData->mass;
	if (m_mass <= 0.0f) m_mass = 1.0f;
	s.invMass = 1.0f / m_mass;
	if (massData->I > 0.0f && (s.flags & B2CU_BODY_FIXED_ROTATION) == 0)
	{
		m_I = massData->I - m_mass * b2Dot(massData->center, massData->center);
		s.invI = 1.0f / m_I;
	}
	b2Vec2 oldCenter(s.cx, s.cy);
	s.lcx = massData->center.x;
	s.lcy = massData->center.y;
	b2Vec2 c = b2Mul(reinterpret_cast<const b2Transform&>(s.px), massData->center);
	s.c0x = s.cx = c.x;
	s.c0y = s.cy = c.y;
	AsVec2(s.vx) += b2Cross(s.w, c - oldCenter);
	m_world->MarkBodyDirty(m_index);
}

// ---- fixtures --------------------------------------------------------------------------------------------

// b2Body::CreateFixture + b2Fixture::Create + CreateProxies + b2DynamicTree::CreateProxy
// (reference b2Body.cpp:165-207, b2Fixture.cpp:43-143, b2DynamicTree.cpp:105-120)
b2Fixture* b2Body::CreateFixture(const b2FixtureDef* def)
{
	if (m_world->IsLocked()) return nullptr;
	b2Assert(def->shape != nullptr);
	m_world->RefreshBodies();

	b2Fixture* f = new b2Fixture;
	f->m_body = this;
	f->m_shape = def->shape->Clone();
	f->m_density = def->density;
	f->m_friction = def->friction;
	f->m_restitution = def->restitution;
	f->m_filter = def->filter;
	f->m_isSensor = def->isSensor;
	f->m_thickShape = def->thickShape;
	f->m_userData = def->userData;

	// one proxy per child (b2Fixture::CreateProxies, reference b2Fixture.cpp:122-137): a chain has one per segment, and
	// its device geometry is that segment as an edge with ghost vertices
	b2BodyView s = B2_STATE();
	const b2Transform& xf = reinterpret_cast<const b2Transform&>(s.px);
	const bool chain = f->m_shape->GetType() == b2Shape::e_chain;
	f->m_proxyIndex = (int32)m_world->m_proxies.size();
	f->m_proxyCount = f->m_shape->GetChildCount();
	for (int32 child = 0; child < f->m_proxyCount; ++child)
	{
		b2AABB aabb;
		f->m_shape->ComputeAABB(&aabb, xf, child);
		b2cuProxy p;
		p.aabb[0] = aabb.lowerBound.x;
		p.aabb[1] = aabb.lowerBound.y;
		p.aabb[2] = aabb.upperBound.x;
		p.aabb[3] = aabb.upperBound.y;
		p.fat[0] = aabb.lowerBound.x - b2_aabbExtension;
		p.fat[1] = aabb.lowerBound.y - b2_aabbExtension;
		p.fat[2] = aabb.upperBound.x + b2_aabbExtension;
		p.fat[3] = aabb.upperBound.y + b2_aabbExtension;
		p.body = m_index;
		if (chain)
		{
			b2EdgeShape edge;
			static_cast<const b2ChainShape*>(f->m_shape)->GetChildEdge(&edge, child);
			p.shape = m_world->InternShape(&edge, true);
		}
		else
		{
			p.shape = m_world->InternShape(f->m_shape);
		}
		p.friction = f->m_friction;
		p.restitution = f->m_restitution;
		p.categoryBits = f->m_filter.categoryBits;
		p.maskBits = f->m_filter.maskBits;
		p.groupIndex = f->m_filter.groupIndex;
		// a fixture of an inactive body gets its proxy slots but stays out of the broad-phase (b2Body.cpp:216-220)
		p.flags = (uint16)((f->m_isSensor ? B2CU_PROXY_SENSOR : 0) | (f->m_thickShape ? B2CU_PROXY_THICK : 0) |
		                   (IsActive() ? (B2CU_PROXY_MOVED | B2CU_PROXY_NEW) : B2CU_PROXY_INACTIVE));
		p.fixture = f->m_proxyIndex;
		p.child = child;
		m_world->m_proxies.push_back(p);
		m_world->m_fixtures.push_back(f);
	}

	f->m_next = m_fixtureList;
	m_fixtureList = f;
	++m_fixtureCount;

	if (f->m_density > 0.0f) ResetMassData();
	m_world->m_newFixture = true;
	return f;
}

b2Fixture* b2Body::CreateFixture(const b2Shape* shape, float32 density)
{
	b2FixtureDef def;
	def.shape = shape;
	def.density = density;
	return CreateFixture(&def);
}

void b2Body::DestroyFixture(b2Fixture* fixture)
{
	if (fixture == nullptr || m_world->IsLocked()) return;
	m_world->DestroyFixtureInternal(this, fixture);
}

b2ContactEdge* b2Body::GetContactList()
{
	m_world->RefreshContacts();
	return m_index < (int32)m_world->m_contactHeads.size() ? m_world->m_contactHeads[m_index] : nullptr;
}

void b2Fixture::SetFilterData(const b2Filter& filter)
{
	m_filter = filter;
	Refilter();
}

// reference b2Fixture.cpp:187-220: the attached contacts are flagged for re-filtering (done on the device from the
// REFILTER bit, before the next Collide) and the proxy is touched (buffered move)
void b2Fixture::Refilter()
{
	if (m_body == nullptr || m_proxyIndex < 0) return;
	b2World* w = m_body->m_world;
	w->RefreshProxies();
	for (int32 child = 0; child < m_proxyCount; ++child)
	{
		b2cuProxy& p = w->m_proxies[m_proxyIndex + child];
		p.categoryBits = m_filter.categoryBits;
		p.maskBits = m_filter.maskBits;
		p.groupIndex = m_filter.groupIndex;
		// e_filterFlag on the fixture's contacts + TouchProxy (reference b2Fixture.cpp:187-220)
		p.flags |= B2CU_PROXY_MOVED | B2CU_PROXY_REFILTER;
		w->MarkProxyDirty(m_proxyIndex + child);
	}
}

void b2Fixture::SetFriction(float32 friction)
{
	m_friction = friction;
	if (m_proxyIndex < 0) return;
	b2World* w = m_body->m_world;
	w->RefreshProxies();
	for (int32 child = 0; child < m_proxyCount; ++child)
	{
		w->m_proxies[m_proxyIndex + child].friction = friction;
		w->MarkProxyDirty(m_proxyIndex + child);
	}
}

void b2Fixture::SetRestitution(float32 restitution)
{
	m_restitution = restitution;
	if (m_proxyIndex < 0) return;
	b2World* w = m_body->m_world;
	w->RefreshProxies();
	for (int32 child = 0; child < m_proxyCount; ++child)
	{
		w->m_proxies[m_proxyIndex + child].restitution = restitution;
		w->MarkProxyDirty(m_proxyIndex + child);
	}
}

void b2Fixture::SetSensor(bool sensor)
{
	if (sensor == m_isSensor) return;
	m_body->SetAwake(true);
	m_isSensor = sensor;
	if (m_proxyIndex < 0) return;
	b2World* w = m_body->m_world;
	w->RefreshProxies();
	for (int32 child = 0; child < m_proxyCount; ++child)
	{
		b2cuProxy& p = w->m_proxies[m_proxyIndex + child];
		if (sensor) p.flags |= B2CU_PROXY_SENSOR;
		else p.flags &= ~(uint16)B2CU_PROXY_SENSOR;
		w->MarkProxyDirty(m_proxyIndex + child);
	}
}

void b2Fixture::SetThickShape(bool flag)
{
	m_thickShape = flag;
	if (m_proxyIndex < 0) return;
	b2World* w = m_body->m_world;
	w->RefreshProxies();
	for (int32 child = 0; child < m_proxyCount; ++child)
	{
		b2cuProxy& p = w->m_proxies[m_proxyIndex + child];
		if (flag) p.flags |= B2CU_PROXY_THICK;
		else p.flags &= ~(uint16)B2CU_PROXY_THICK;
		w->MarkProxyDirty(m_proxyIndex + child);
	}
}

bool b2Fixture::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, int32 childIndex) const
{
	return m_shape->RayCast(output, input, m_body->GetTransform(), childIndex);
}

bool b2Fixture::TestPoint(const b2Vec2& p) const { return m_shape->TestPoint(m_body->GetTransform(), p); }

const b2AABB& b2Fixture::GetAABB(int32 childIndex) const
{
	b2Assert(0 <= childIndex && childIndex < m_proxyCount);
	b2World* w = m_body->m_world;
	w->RefreshProxies();
	return reinterpret_cast<const b2AABB&>(w->m_proxies[m_proxyIndex + childIndex].aabb[0]);
}
