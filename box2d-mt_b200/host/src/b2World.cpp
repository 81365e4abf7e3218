// b2World: body / fixture bookkeeping on the host, state mirrored to the device, Step delegated to the executor.
// reference: Box2D/Dynamics/b2World.cpp (CreateBody :532-559, DestroyBody :561-640, Step :1613-1710) and
// b2ContactManager.cpp (callback dispatch :388-439).
#include "Box2D/Dynamics/b2World.h"
#include "Box2D/Dynamics/Joints/b2RevoluteJoint.h"
#include "Box2D/Dynamics/Joints/b2DistanceJoint.h"
#include "Box2D/Dynamics/Joints/b2WeldJoint.h"
#include "Box2D/Dynamics/Joints/b2PrismaticJoint.h"
#include "Box2D/Dynamics/Joints/b2WheelJoint.h"
#include "Box2D/Dynamics/Joints/b2RopeJoint.h"
#include "Box2D/Dynamics/Joints/b2FrictionJoint.h"
#include "Box2D/Dynamics/Joints/b2MotorJoint.h"
#include "Box2D/Dynamics/Joints/b2PulleyJoint.h"
#include "Box2D/Dynamics/Joints/b2MouseJoint.h"
#include "Box2D/Dynamics/Joints/b2GearJoint.h"

#include <chrono>
#include "Box2D/Collision/Shapes/b2CircleShape.h"
#include "Box2D/Collision/Shapes/b2EdgeShape.h"
#include "Box2D/Collision/Shapes/b2PolygonShape.h"
#include "Box2D/MT/b2CudaStepExecutor.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

bool b2ContactFilter::ShouldCollide(b2Fixture* fixtureA, b2Fixture* fixtureB, uint32 threadId)
{
	B2_NOT_USED(threadId);
	const b2Filter& fa = fixtureA->GetFilterData();
	const b2Filter& fb = fixtureB->GetFilterData();
	if (fa.groupIndex == fb.groupIndex && fa.groupIndex != 0) return fa.groupIndex > 0;
	return (fa.maskBits & fb.categoryBits) != 0 && (fa.categoryBits & fb.maskBits) != 0;
}

// reference: Box2D/Collision/b2Collision.cpp:22-86
void b2WorldManifold::Initialize(const b2Manifold* manifold, const b2Transform& xfA, float32 radiusA,
                                 const b2Transform& xfB, float32 radiusB)
{
	if (manifold->pointCount == 0) return;
	switch (manifold->type)
	{
	case b2Manifold::e_circles:
	{
		normal.Set(1.0f, 0.0f);
		b2Vec2 pointA = b2Mul(xfA, manifold->localPoint);
		b2Vec2 pointB = b2Mul(xfB, manifold->points[0].localPoint);
		if (b2DistanceSquared(pointA, pointB) > b2_epsilon * b2_epsilon)
		{
			normal = pointB - pointA;
			normal.Normalize();
		}
		b2Vec2 cA = pointA + radiusA * normal;
		b2Vec2 cB = pointB - radiusB * normal;
		points[0] = 0.5f * (cA + cB);
		separations[0] = b2Dot(cB - cA, normal);
		break;
	}
	case b2Manifold::e_faceA:
	{
		normal = b2Mul(xfA.q, manifold->localNormal);
		b2Vec2 planePoint = b2Mul(xfA, manifold->localPoint);
		for (int32 i = 0; i < manifold->pointCount; ++i)
		{
			b2Vec2 clipPoint = b2Mul(xfB, manifold->points[i].localPoint);
			b2Vec2 cA = clipPoint + (radiusA - b2Dot(clipPoint - planePoint, normal)) * normal;
			b2Vec2 cB = clipPoint - radiusB * normal;
			points[i] = 0.5f * (cA + cB);
			separations[i] = b2Dot(cB - cA, normal);
		}
		break;
	}
	case b2Manifold::e_faceB:
	{
		normal = b2Mul(xfB.q, manifold->localNormal);
		b2Vec2 planePoint = b2Mul(xfB, manifold->localPoint);
		for (int32 i = 0; i < manifold->pointCount; ++i)
		{
			b2Vec2 clipPoint = b2Mul(xfA, manifold->points[i].localPoint);
			b2Vec2 cB = clipPoint + (radiusB - b2Dot(clipPoint - planePoint, normal)) * normal;
			b2Vec2 cA = clipPoint - radiusA * normal;
			points[i] = 0.5f * (cA + cB);
			separations[i] = b2Dot(cA - cB, normal);
		}
		normal = -normal;
		break;
	}
	}
}

void b2Contact::GetWorldManifold(b2WorldManifold* worldManifold) const
{
	const b2Body* bodyA = m_fixtureA->GetBody();
	const b2Body* bodyB = m_fixtureB->GetBody();
	worldManifold->Initialize(&m_manifold, bodyA->GetTransform(), m_fixtureA->GetShape()->m_radius,
	                          bodyB->GetTransform(), m_fixtureB->GetShape()->m_radius);
}

// ---- construction ---------------------------------------------------------------------------------------

b2World::b2World(const b2Vec2& gravity)
	: m_sweepStartsStale(false), m_device(nullptr), m_owner(nullptr), m_fullUpload(false), m_bodiesUploaded(0), m_proxiesUploaded(0),
	  m_shapesUploaded(0), m_bodyDirtyLo(INT32_MAX), m_bodyDirtyHi(-1), m_proxyDirtyLo(INT32_MAX), m_proxyDirtyHi(-1),
	  m_forceDirtyLo(INT32_MAX), m_forceDirtyHi(-1), m_bodiesStale(false), m_proxiesStale(false), m_contactsStale(true), m_jointList(nullptr), m_jointsDirty(false),
	  m_jointsStale(false), m_bodyList(nullptr), m_bodyCount(0),
	  m_contactCount(0), m_gravity(gravity), m_allowSleep(true), m_warmStarting(true), m_continuousPhysics(true),
	  m_subStepping(false), m_clearForces(true), m_locked(false), m_newFixture(false), m_inv_dt0(0.0f),
	  m_destructionListener(nullptr), m_contactFilter(nullptr), m_contactListener(nullptr), m_lastStatus(0)
{
	memset(&m_profile, 0, sizeof(m_profile));
}

b2World::~b2World()
{
	if (m_owner) m_owner->DetachWorld(this);
	// m_fixtures has one entry per PROXY: a chain fixture appears once per segment
	for (size_t i = 0; i < m_fixtures.size(); ++i)
		if (i == 0 || m_fixtures[i] != m_fixtures[i - 1]) delete m_fixtures[i];
	for (size_t i = 0; i < m_joints.size(); ++i) delete m_joints[i];
	for (size_t i = 0; i < m_bodies.size(); ++i) delete m_bodies[i];
}

// reference b2World.cpp:532-559 + b2Body::b2Body (b2Body.cpp:26-111)
b2Body* b2World::CreateBody(const b2BodyDef* def)
{
	if (IsLocked()) return nullptr;
	RefreshBodies();

	b2Body* b = new b2Body;
	b->m_world = this;
	b->m_index = (int32)m_states.size();
	b->m_fixtureList = nullptr;
	b->m_fixtureCount = 0;
	b->m_jointList = nullptr;
	b->m_userData = def->userData;
	b->m_I = 0.0f;

	b2cuBody s;
	memset(&s, 0, sizeof(s));
	uint32 flags = (uint32)def->type;
	if (def->bullet) flags |= B2CU_BODY_BULLET;
	if (def->fixedRotation) flags |= B2CU_BODY_FIXED_ROTATION;
	if (def->allowSleep) flags |= B2CU_BODY_AUTOSLEEP;
	if (def->awake) flags |= B2CU_BODY_AWAKE;
	if (def->active) flags |= B2CU_BODY_ACTIVE;
	s.flags = flags;
	b2Rot q(def->angle);
	s.px = def->position.x;
	s.py = def->position.y;
	s.qs = q.s;
	s.qc = q.c;
	s.cx = s.c0x = def->position.x;
	s.cy = s.c0y = def->position.y;
	s.a = s.a0 = def->angle;
	s.alpha0 = 0.0f;
	s.vx = def->linearVelocity.x;
	s.vy = def->linearVelocity.y;
	s.w = def->angularVelocity;
	s.linearDamping = def->linearDamping;
	s.angularDamping = def->angularDamping;
	s.gravityScale = def->gravityScale;
	if (def->type == b2_dynamicBody)
	{
		b->m_mass = 1.0f;
		s.invMass = 1.0f;
	}
	else
	{
		b->m_mass = 0.0f;
		s.invMass = 0.0f;
	}
	{
		b2cuBodyState st;
		b2cuSweepStart z;
		b2BodyProps pr;
		st.px = s.px; st.py = s.py; st.qs = s.qs; st.qc = s.qc;
		st.cx = s.cx; st.cy = s.cy; st.a = s.a;
		z.c0x = s.c0x; z.c0y = s.c0y; z.a0 = s.a0; z.alpha0 = s.alpha0;
		st.vx = s.vx; st.vy = s.vy; st.w = s.w;
		st.sleepTime = s.sleepTime;
		st.flags = s.flags;
		pr.lcx = s.lcx; pr.lcy = s.lcy;
		pr.fx = s.fx; pr.fy = s.fy; pr.torque = s.torque;
		pr.invMass = s.invMass; pr.invI = s.invI;
		pr.linearDamping = s.linearDamping; pr.angularDamping = s.angularDamping; pr.gravityScale = s.gravityScale;
		m_states.push_back(st);
		m_sweepStarts.push_back(z);
		m_props.push_back(pr);
	}
	m_bodies.push_back(b);

	// newest first, as the reference's body list
	b->m_prev = nullptr;
	b->m_next = m_bodyList;
	if (m_bodyList) m_bodyList->m_prev = b;
	m_bodyList = b;
	++m_bodyCount;
	return b;
}

int32 b2World::InternShape(const b2Shape* shape, bool chainChild)
{
	b2cuShape r;
	memset(&r, 0, sizeof(r));
	r.radius = shape->m_radius;
	switch (shape->GetType())
	{
	case b2Shape::e_circle:
	{
		const b2CircleShape* c = static_cast<const b2CircleShape*>(shape);
		r.type = B2CU_SHAPE_CIRCLE;
		r.count = 1;
		r.v[0][0] = c->m_p.x;
		r.v[0][1] = c->m_p.y;
		break;
	}
	case b2Shape::e_edge:
	{
		const b2EdgeShape* e = static_cast<const b2EdgeShape*>(shape);
		r.type = B2CU_SHAPE_EDGE;
		if (chainChild) r.flags |= B2CU_EDGE_CHAIN_CHILD;
		r.count = 2;
		r.v[0][0] = e->m_vertex1.x; r.v[0][1] = e->m_vertex1.y;
		r.v[1][0] = e->m_vertex2.x; r.v[1][1] = e->m_vertex2.y;
		r.v[2][0] = e->m_vertex0.x; r.v[2][1] = e->m_vertex0.y;
		r.v[3][0] = e->m_vertex3.x; r.v[3][1] = e->m_vertex3.y;
		r.flags |= (e->m_hasVertex0 ? B2CU_EDGE_HAS_VERTEX0 : 0) | (e->m_hasVertex3 ? B2CU_EDGE_HAS_VERTEX3 : 0);
		break;
	}
	default:
	{
		const b2PolygonShape* p = static_cast<const b2PolygonShape*>(shape);
		r.type = B2CU_SHAPE_POLYGON;
		r.count = p->m_count;
		for (int32 i = 0; i < p->m_count; ++i)
		{
			r.v[i][0] = p->m_vertices[i].x; r.v[i][1] = p->m_vertices[i].y;
			r.n[i][0] = p->m_normals[i].x; r.n[i][1] = p->m_normals[i].y;
		}
		r.centroid[0] = p->m_centroid.x;
		r.centroid[1] = p->m_centroid.y;
		break;
	}
	}
	std::string key(reinterpret_cast<const char*>(&r), sizeof(r));
	auto it = m_shapeLookup.find(key);
	if (it != m_shapeLookup.end()) return it->second;
	int32 index = (int32)m_shapes.size();
	m_shapes.push_back(r);
	m_shapeLookup.emplace(std::move(key), index);
	return index;
}

void b2World::MarkBodyDirty(int32 index)
{
	m_bodyDirtyLo = std::min(m_bodyDirtyLo, index);
	m_bodyDirtyHi = std::max(m_bodyDirtyHi, index);
}

void b2World::MarkBodyForced(int32 index)
{
	// only m_force / m_torque of the row changed: it travels as 12 bytes (b2cuSetBodyForces), not as a whole record
	m_forceDirtyLo = std::min(m_forceDirtyLo, index);
	m_forceDirtyHi = std::max(m_forceDirtyHi, index);
	m_forced.push_back(index);
}

void b2World::MarkProxyDirty(int32 index)
{
	m_proxyDirtyLo = std::min(m_proxyDirtyLo, index);
	m_proxyDirtyHi = std::max(m_proxyDirtyHi, index);
}

void b2World::SetAllowSleeping(bool flag)
{
	if (flag == m_allowSleep) return;
	m_allowSleep = flag;
	if (!flag)
	{
		for (b2Body* b = m_bodyList; b; b = b->m_next) b->SetAwake(true);
	}
}

void b2World::ClearForces()
{
	RefreshBodies();
	for (size_t i = 0; i < m_props.size(); ++i)
	{
		m_props[i].fx = m_props[i].fy = m_props[i].torque = 0.0f;
	}
	m_forced.clear();
	if (!m_states.empty())
	{
		MarkBodyDirty(0);
		MarkBodyDirty((int32)m_states.size() - 1);
	}
}

const b2cuBody* b2World::GetBodyStates() const
{
	RefreshBodies();
	RefreshSweepStarts();
	b2World* self = const_cast<b2World*>(this);
	m_records.resize(m_states.size());
	for (size_t i = 0; i < m_states.size(); ++i) self->BodyView((int32)i).ToRecord(&m_records[i]);
	return m_records.data();
}

const b2cuProxy* b2World::GetProxyStates() const
{
	RefreshProxies();
	return m_proxies.data();
}

// ---- device <-> host mirror ------------------------------------------------------------------------------

void b2World::RefreshBodies() const
{
	if (!m_bodiesStale || m_device == nullptr) return;
	b2World* self = const_cast<b2World*>(this);
	int32 n = std::min(m_bodiesUploaded, (int32)m_states.size());
	if (n > 0) b2cuGetBodyStates(m_device, 0, n, self->m_states.data());
	m_bodiesStale = false;
}

void b2World::RefreshSweepStarts() const
{
	if (!m_sweepStartsStale || m_device == nullptr) return;
	b2World* self = const_cast<b2World*>(this);
	int32 n = std::min(m_bodiesUploaded, (int32)m_sweepStarts.size());
	if (n > 0) b2cuGetBodySweepStarts(m_device, 0, n, self->m_sweepStarts.data());
	m_sweepStartsStale = false;
}

void b2World::RefreshJoints() const
{
	if (!m_jointsStale || m_device == nullptr) return;
	m_jointsStale = false;
	int32 n = 0;
	if (b2cuGetJointCount(m_device, &n) != B2CU_OK || n != (int32)m_joints.size() || n == 0) return;
	std::vector<b2cuJoint> rows((size_t)n);
	if (b2cuGetJoints(m_device, 0, n, rows.data()) != B2CU_OK) return;
	for (int32 i = 0; i < n; ++i) m_joints[i]->ReadRecord(rows[i]);
}

// reference b2World.cpp:659-733
b2Joint* b2World::CreateJoint(const b2JointDef* def)
{
	if (IsLocked()) return nullptr;
	if (def->type <= e_unknownJoint || def->type > e_motorJoint)
	{
		m_lastStatus = B2CU_ERR_UNSUPPORTED;
		return nullptr;
	}
	RefreshJoints();
	RefreshBodies(); // constructors that read the bodies' transforms (gear, mouse) need the current ones
	b2Joint* j;
	if (def->type == e_revoluteJoint) j = new b2RevoluteJoint(static_cast<const b2RevoluteJointDef*>(def));
	else if (def->type == e_distanceJoint) j = new b2DistanceJoint(static_cast<const b2DistanceJointDef*>(def));
	else if (def->type == e_prismaticJoint) j = new b2PrismaticJoint(static_cast<const b2PrismaticJointDef*>(def));
	else if (def->type == e_wheelJoint) j = new b2WheelJoint(static_cast<const b2WheelJointDef*>(def));
	else if (def->type == e_ropeJoint) j = new b2RopeJoint(static_cast<const b2RopeJointDef*>(def));
	else if (def->type == e_frictionJoint) j = new b2FrictionJoint(static_cast<const b2FrictionJointDef*>(def));
	else if (def->type == e_motorJoint) j = new b2MotorJoint(static_cast<const b2MotorJointDef*>(def));
	else if (def->type == e_pulleyJoint) j = new b2PulleyJoint(static_cast<const b2PulleyJointDef*>(def));
	else if (def->type == e_mouseJoint) j = new b2MouseJoint(static_cast<const b2MouseJointDef*>(def));
	else if (def->type == e_gearJoint) j = new b2GearJoint(static_cast<const b2GearJointDef*>(def));
	else j = new b2WeldJoint(static_cast<const b2WeldJointDef*>(def));
	j->m_world = this;
	j->m_index = (int32)m_joints.size();
	m_joints.push_back(j);

	j->m_prev = nullptr;
	j->m_next = m_jointList;
	if (m_jointList) m_jointList->m_prev = j;
	m_jointList = j;

	b2JointEdge* edges[2] = {&j->m_edgeA, &j->m_edgeB};
	b2Body* ends[2] = {j->m_bodyA, j->m_bodyB};
	for (int32 k = 0; k < 2; ++k)
	{
		edges[k]->joint = j;
		edges[k]->other = ends[1 - k];
		edges[k]->prev = nullptr;
		edges[k]->next = ends[k]->m_jointList;
		if (ends[k]->m_jointList) ends[k]->m_jointList->prev = edges[k];
		ends[k]->m_jointList = edges[k];
	}
	// contacts the joint forbids are re-filtered by the device when the table arrives (b2cuSetJoints)
	m_jointsDirty = true;
	return j;
}

// reference b2World.cpp:735-841
void b2World::DestroyJoint(b2Joint* j)
{
	if (IsLocked() || j == nullptr) return;
	RefreshJoints();
	if (j->m_prev) j->m_prev->m_next = j->m_next;
	if (j->m_next) j->m_next->m_prev = j->m_prev;
	if (j == m_jointList) m_jointList = j->m_next;

	b2JointEdge* edges[2] = {&j->m_edgeA, &j->m_edgeB};
	b2Body* ends[2] = {j->m_bodyA, j->m_bodyB};
	for (int32 k = 0; k < 2; ++k)
	{
		// wake the bodies: whatever the joint held up may now fall
		ends[k]->SetAwake(true);
		if (edges[k]->prev) edges[k]->prev->next = edges[k]->next;
		if (edges[k]->next) edges[k]->next->prev = edges[k]->prev;
		if (edges[k] == ends[k]->m_jointList) ends[k]->m_jointList = edges[k]->next;
	}
	m_joints.erase(m_joints.begin() + j->m_index);
	for (size_t i = 0; i < m_joints.size(); ++i) m_joints[i]->m_index = (int32)i;
	delete j;
	m_jointsDirty = true;
}

void b2World::RefreshProxies() const
{
	if (!m_proxiesStale || m_device == nullptr) return;
	b2World* self = const_cast<b2World*>(this);
	int32 n = std::min(m_proxiesUploaded, (int32)m_proxies.size());
	if (n > 0) b2cuGetProxies(m_device, 0, n, self->m_proxies.data());
	// the device does not keep the child index: it follows from the fixture's first proxy
	for (int32 i = 0; i < n; ++i)
	{
		self->m_proxies[i].child = i - m_fixtures[i]->m_proxyIndex;
		self->m_proxies[i].fixture = m_fixtures[i]->m_proxyIndex;
	}
	m_proxiesStale = false;
}

void b2World::InvalidateSnapshots()
{
	m_contactsStale = true;
	m_contacts.clear();
	m_contactHeads.clear();
}

void b2World::MakeContact(b2Contact* c, const b2cuContact& rec)
{
	c->m_flags = rec.flags;
	uint32 lo = (uint32)std::min(rec.proxyA, rec.proxyB), hi = (uint32)std::max(rec.proxyA, rec.proxyB);
	c->m_key = ((uint64)lo << 32) | hi;
	c->m_fixtureA = m_fixtures[rec.proxyA];
	c->m_fixtureB = m_fixtures[rec.proxyB];
	c->m_indexA = rec.proxyA - c->m_fixtureA->m_proxyIndex;
	c->m_indexB = rec.proxyB - c->m_fixtureB->m_proxyIndex;
	const b2cuManifold& m = rec.manifold;
	c->m_manifold.localNormal.Set(m.localNormal[0], m.localNormal[1]);
	c->m_manifold.localPoint.Set(m.localPoint[0], m.localPoint[1]);
	for (int32 i = 0; i < 2; ++i)
	{
		c->m_manifold.points[i].localPoint.Set(m.points[i].localPoint[0], m.points[i].localPoint[1]);
		c->m_manifold.points[i].normalImpulse = m.points[i].normalImpulse;
		c->m_manifold.points[i].tangentImpulse = m.points[i].tangentImpulse;
		c->m_manifold.points[i].id.key = m.id[i];
	}
	c->m_manifold.type = (b2Manifold::Type)m.type;
	c->m_manifold.pointCount = m.pointCount;
	c->m_friction = rec.friction;
	c->m_restitution = rec.restitution;
	c->m_tangentSpeed = rec.tangentSpeed;
	c->m_next = nullptr;
	c->m_nodeA.contact = c;
	c->m_nodeA.other = c->m_fixtureB->GetBody();
	c->m_nodeA.prev = c->m_nodeA.next = nullptr;
	c->m_nodeB.contact = c;
	c->m_nodeB.other = c->m_fixtureA->GetBody();
	c->m_nodeB.prev = c->m_nodeB.next = nullptr;
}

// Download the device contact set (key order) and materialise b2Contact objects + per-body edge lists.  While a
// full upload is pending (after a destruction) the host records are the authoritative set.
void b2World::RefreshContacts()
{
	if (!m_contactsStale) return;
	m_contactsStale = false;
	m_contacts.clear();
	m_contactHeads.assign(m_bodies.size(), nullptr);
	if (m_device != nullptr && !m_fullUpload)
	{
		int32 n = 0;
		b2cuGetContactCount(m_device, &n);
		m_contactRecords.resize(n);
		if (n > 0) b2cuGetContacts(m_device, n, m_contactRecords.data(), &n);
	}
	else if (m_device == nullptr && !m_fullUpload)
	{
		m_contactRecords.clear();
	}
	const int32 n = (int32)m_contactRecords.size();
	if (n == 0) return;
	m_contacts.resize(n);
	for (int32 i = 0; i < n; ++i)
	{
		b2Contact* c = &m_contacts[i];
		MakeContact(c, m_contactRecords[i]);
		c->m_next = i + 1 < n ? &m_contacts[i + 1] : nullptr;
	}
	for (int32 i = 0; i < n; ++i)
	{
		b2Contact* c = &m_contacts[i];
		int32 ia = c->m_fixtureA->GetBody()->m_index, ib = c->m_fixtureB->GetBody()->m_index;
		c->m_nodeA.next = m_contactHeads[ia];
		if (m_contactHeads[ia]) m_contactHeads[ia]->prev = &c->m_nodeA;
		m_contactHeads[ia] = &c->m_nodeA;
		c->m_nodeB.next = m_contactHeads[ib];
		if (m_contactHeads[ib]) m_contactHeads[ib]->prev = &c->m_nodeB;
		m_contactHeads[ib] = &c->m_nodeB;
	}
}

b2Contact* b2World::GetContactList()
{
	RefreshContacts();
	return m_contacts.empty() ? nullptr : &m_contacts[0];
}

// ---- destruction: compact the dense arrays and remap the contact set --------------------------------------

void b2World::RemoveProxies(const std::vector<int32>& proxyIds, const std::vector<int32>& bodyIds)
{
	RefreshBodies();
	RefreshSweepStarts();
	RefreshProxies();
	// contact records as they are on the device
	m_contactsStale = true;
	RefreshContacts();

	std::vector<int32> proxyMap(m_proxies.size()), bodyMap(m_states.size());
	std::vector<char> deadProxy(m_proxies.size(), 0), deadBody(m_states.size(), 0);
	for (size_t i = 0; i < proxyIds.size(); ++i) deadProxy[proxyIds[i]] = 1;
	for (size_t i = 0; i < bodyIds.size(); ++i) deadBody[bodyIds[i]] = 1;

	// contacts of the removed proxies end (b2ContactManager::Destroy, b2ContactManager.cpp:120-172): EndContact if
	// touching, both bodies woken if the manifold had points
	std::vector<b2cuContact> kept;
	kept.reserve(m_contactRecords.size());
	for (size_t i = 0; i < m_contactRecords.size(); ++i)
	{
		const b2cuContact& rec = m_contactRecords[i];
		if (deadProxy[rec.proxyA] || deadProxy[rec.proxyB])
		{
			b2Contact* c = &m_contacts[i];
			if (m_contactListener && c->IsTouching() && m_contactListener->EndContactImmediate(c, 0))
			{
				m_contactListener->EndContact(c);
			}
			if (rec.manifold.pointCount > 0)
			{
				c->m_fixtureA->GetBody()->SetAwake(true);
				c->m_fixtureB->GetBody()->SetAwake(true);
			}
			continue;
		}
		kept.push_back(rec);
	}

	int32 np = 0;
	for (size_t i = 0; i < m_proxies.size(); ++i)
	{
		proxyMap[i] = deadProxy[i] ? -1 : np;
		if (!deadProxy[i])
		{
			m_proxies[np] = m_proxies[i];
			m_fixtures[np] = m_fixtures[i];
			if (np == 0 || m_fixtures[np] != m_fixtures[np - 1]) m_fixtures[np]->m_proxyIndex = np;
			++np;
		}
	}
	m_proxies.resize(np);
	m_fixtures.resize(np);

	int32 nb = 0;
	for (size_t i = 0; i < m_states.size(); ++i)
	{
		bodyMap[i] = deadBody[i] ? -1 : nb;
		if (!deadBody[i])
		{
			m_states[nb] = m_states[i];
			m_sweepStarts[nb] = m_sweepStarts[i];
			m_props[nb] = m_props[i];
			m_bodies[nb] = m_bodies[i];
			m_bodies[nb]->m_index = nb;
			++nb;
		}
	}
	m_states.resize(nb);
	m_sweepStarts.resize(nb);
	m_props.resize(nb);
	m_bodies.resize(nb);
	{
		size_t kept = 0;
		for (size_t i = 0; i < m_forced.size(); ++i)
			if (bodyMap[m_forced[i]] >= 0) m_forced[kept++] = bodyMap[m_forced[i]];
		m_forced.resize(kept);
	}

	for (int32 i = 0; i < np; ++i)
	{
		m_proxies[i].body = bodyMap[m_proxies[i].body];
		m_proxies[i].fixture = m_fixtures[i]->m_proxyIndex;
	}
	for (size_t i = 0; i < kept.size(); ++i)
	{
		kept[i].proxyA = proxyMap[kept[i].proxyA];
		kept[i].proxyB = proxyMap[kept[i].proxyB];
	}
	m_contactRecords.swap(kept);
	m_contactCount = (int32)m_contactRecords.size();
	m_contacts.clear();
	m_contactHeads.clear();
	m_contactsStale = true; // snapshots are rebuilt from m_contactRecords, which is authoritative until the upload
	m_fullUpload = true;
}

// b2Body::SetType destroys every contact attached to the body (reference b2Body.cpp:168-176): EndContact for the
// touching ones, both bodies woken if the manifold had points (b2ContactManager::Destroy).  The surviving records go
// back to the device with the next step (full upload: rare operation, simplest exact path).
void b2World::DestroyContactsOfBody(int32 bodyIndex)
{
	if (m_device == nullptr && !m_fullUpload) return; // never stepped: there are no contacts yet
	RefreshBodies();
	RefreshProxies();
	m_contactsStale = true;
	RefreshContacts();
	std::vector<b2cuContact> kept;
	kept.reserve(m_contactRecords.size());
	for (size_t i = 0; i < m_contactRecords.size(); ++i)
	{
		const b2cuContact& rec = m_contactRecords[i];
		if (m_proxies[rec.proxyA].body == bodyIndex || m_proxies[rec.proxyB].body == bodyIndex)
		{
			b2Contact* c = &m_contacts[i];
			if (m_contactListener && c->IsTouching() && m_contactListener->EndContactImmediate(c, 0))
			{
				m_contactListener->EndContact(c);
			}
			if (rec.manifold.pointCount > 0)
			{
				c->m_fixtureA->GetBody()->SetAwake(true);
				c->m_fixtureB->GetBody()->SetAwake(true);
			}
			continue;
		}
		kept.push_back(rec);
	}
	m_contactRecords.swap(kept);
	m_contactCount = (int32)m_contactRecords.size();
	m_contacts.clear();
	m_contactHeads.clear();
	m_contactsStale = true;
	m_fullUpload = true;
}

void b2World::DestroyFixtureInternal(b2Body* body, b2Fixture* fixture)
{
	b2Fixture** link = &body->m_fixtureList;
	while (*link && *link != fixture) link = &(*link)->m_next;
	if (*link == nullptr) return;
	*link = fixture->m_next;
	--body->m_fixtureCount;
	std::vector<int32> proxies, bodies;
	for (int32 k = 0; k < fixture->m_proxyCount; ++k) proxies.push_back(fixture->m_proxyIndex + k);
	RemoveProxies(proxies, bodies);
	delete fixture;
	body->ResetMassData();
}

// reference b2World.cpp:561-640
void b2World::DestroyBody(b2Body* b)
{
	if (IsLocked() || b == nullptr) return;
	// the joints attached to the body go first (reference b2World.cpp:594-610)
	while (b->m_jointList)
	{
		b2Joint* doomedJoint = b->m_jointList->joint;
		if (m_destructionListener) m_destructionListener->SayGoodbye(doomedJoint);
		DestroyJoint(doomedJoint);
	}
	if (!m_joints.empty())
	{
		// body rows are renumbered below: the joint table is rebuilt from the objects
		RefreshJoints();
		m_jointsDirty = true;
	}
	std::vector<int32> proxies, bodies(1, b->m_index);
	for (b2Fixture* f = b->m_fixtureList; f; f = f->m_next)
	{
		if (m_destructionListener) m_destructionListener->SayGoodbye(f);
		for (int32 k = 0; k < f->m_proxyCount; ++k) proxies.push_back(f->m_proxyIndex + k);
	}
	std::vector<b2Fixture*> doomed;
	for (b2Fixture* f = b->m_fixtureList; f; f = f->m_next) doomed.push_back(f);
	RemoveProxies(proxies, bodies);
	for (size_t i = 0; i < doomed.size(); ++i) delete doomed[i];

	if (b->m_prev) b->m_prev->m_next = b->m_next;
	if (b->m_next) b->m_next->m_prev = b->m_prev;
	if (b == m_bodyList) m_bodyList = b->m_next;
	--m_bodyCount;
	delete b;
}

// ---- queries ---------------------------------------------------------------------------------------------

// Candidate proxies of a box or segment query, ascending ids.  With a device copy the fat boxes are scanned there
// (after the pending edits have been uploaded); a world that has never been stepped has no device copy yet and is
// scanned here.
void b2World::ProxyQuery(const b2AABB* box, const b2Vec2* p1, const b2Vec2* p2, std::vector<int32>& ids)
{
	ids.clear();
	if (m_device != nullptr)
	{
		if (UploadDirty(m_device) != B2CU_OK) return;
		int32 n = 0;
		float a[4] = {0, 0, 0, 0}, q1[2] = {0, 0}, q2[2] = {0, 0};
		if (box)
		{
			a[0] = box->lowerBound.x; a[1] = box->lowerBound.y; a[2] = box->upperBound.x; a[3] = box->upperBound.y;
			if (b2cuQueryAABB(m_device, a, 0, nullptr, &n) != B2CU_OK || n == 0) return;
			ids.resize((size_t)n);
			b2cuQueryAABB(m_device, a, n, ids.data(), &n);
		}
		else
		{
			q1[0] = p1->x; q1[1] = p1->y; q2[0] = p2->x; q2[1] = p2->y;
			if (b2cuRayCastCandidates(m_device, q1, q2, 0, nullptr, &n) != B2CU_OK || n == 0) return;
			ids.resize((size_t)n);
			b2cuRayCastCandidates(m_device, q1, q2, n, ids.data(), &n);
		}
		return;
	}
	b2AABB seg;
	if (!box)
	{
		seg.lowerBound = b2Min(*p1, *p2);
		seg.upperBound = b2Max(*p1, *p2);
	}
	for (size_t i = 0; i < m_proxies.size(); ++i)
	{
		const b2AABB& fat = reinterpret_cast<const b2AABB&>(m_proxies[i].fat[0]);
		if (b2TestOverlap(fat, box ? *box : seg)) ids.push_back((int32)i);
	}
}

void b2World::QueryAABB(b2QueryCallback* callback, const b2AABB& aabb)
{
	if (IsLocked() || callback == nullptr) return;
	std::vector<int32> ids;
	ProxyQuery(&aabb, nullptr, nullptr, ids);
	for (size_t i = 0; i < ids.size(); ++i)
		if (!callback->ReportFixture(m_fixtures[ids[i]])) break;
}

void b2World::RayCast(b2RayCastCallback* callback, const b2Vec2& point1, const b2Vec2& point2)
{
	if (IsLocked() || callback == nullptr) return;
	std::vector<int32> ids;
	ProxyQuery(nullptr, &point1, &point2, ids);
	RefreshBodies();
	b2RayCastInput input;
	input.p1 = point1;
	input.p2 = point2;
	input.maxFraction = 1.0f;
	for (size_t i = 0; i < ids.size(); ++i)
	{
		b2Fixture* fixture = m_fixtures[ids[i]];
		b2RayCastOutput output;
		if (!fixture->RayCast(&output, input, ids[i] - fixture->m_proxyIndex)) continue;
		float32 fraction = output.fraction;
		b2Vec2 point = (1.0f - fraction) * input.p1 + fraction * input.p2;
		float32 value = callback->ReportFixture(fixture, point, output.normal, fraction);
		if (value == 0.0f) return;           // the client has terminated the cast
		if (value > 0.0f) input.maxFraction = value; // clip (value < 0: ignore this fixture and go on)
	}
}

void b2World::ShiftOrigin(const b2Vec2& newOrigin)
{
	if (IsLocked()) return;
	RefreshBodies();
	RefreshSweepStarts();
	RefreshProxies();
	for (size_t i = 0; i < m_states.size(); ++i)
	{
		b2cuBodyState& s = m_states[i];
		s.px -= newOrigin.x; s.py -= newOrigin.y;
		s.cx -= newOrigin.x; s.cy -= newOrigin.y;
		m_sweepStarts[i].c0x -= newOrigin.x; m_sweepStarts[i].c0y -= newOrigin.y;
	}
	for (size_t i = 0; i < m_proxies.size(); ++i)
	{
		b2cuProxy& p = m_proxies[i];
		p.aabb[0] -= newOrigin.x; p.aabb[1] -= newOrigin.y; p.aabb[2] -= newOrigin.x; p.aabb[3] -= newOrigin.y;
		p.fat[0] -= newOrigin.x; p.fat[1] -= newOrigin.y; p.fat[2] -= newOrigin.x; p.fat[3] -= newOrigin.y;
	}
	if (!m_states.empty())
	{
		MarkBodyDirty(0);
		MarkBodyDirty((int32)m_states.size() - 1);
	}
	if (!m_proxies.empty())
	{
		MarkProxyDirty(0);
		MarkProxyDirty((int32)m_proxies.size() - 1);
	}
	// joints that hold world points (reference b2World.cpp:2096-2099)
	for (size_t i = 0; i < m_joints.size(); ++i) m_joints[i]->ShiftOrigin(newOrigin);
}

// ---- step ------------------------------------------------------------------------------------------------

void b2World::Step(float32 timeStep, int32 velocityIterations, int32 positionIterations, b2TaskExecutor& executor)
{
	m_locked = true;
	bool ran = executor.StepWorld(*this, timeStep, velocityIterations, positionIterations);
	m_locked = false;
	if (!ran)
	{
		// no CPU fallback: an executor that cannot run the step on the device is a hard error
		if (m_lastStatus == 0) m_lastStatus = B2CU_ERR_NO_DEVICE;
		fprintf(stderr, "b2World::Step: the executor did not run the step on a GPU (status %d); this library has no "
		                "CPU step path\n", m_lastStatus);
		b2Assert(false);
	}
}

int32 b2World::UploadDirty(b2cuWorld* device)
{
	int32 rc;
	const int32 nb = (int32)m_states.size(), np = (int32)m_proxies.size(), ns = (int32)m_shapes.size();
	uint32 flags = (m_allowSleep ? B2CU_WORLD_ALLOW_SLEEP : 0) | (m_warmStarting ? B2CU_WORLD_WARM_STARTING : 0) |
	               (m_continuousPhysics ? B2CU_WORLD_CONTINUOUS : 0) | (m_subStepping ? B2CU_WORLD_SUB_STEPPING : 0) |
	               (m_clearForces ? B2CU_WORLD_CLEAR_FORCES : 0);
	float g[2] = {m_gravity.x, m_gravity.y};
	if ((rc = b2cuSetWorldParams(device, g, flags))) return rc;

	if (m_fullUpload)
	{
		m_bodiesUploaded = m_proxiesUploaded = m_shapesUploaded = 0;
	}
	if (nb != m_bodiesUploaded || np != m_proxiesUploaded || ns != m_shapesUploaded)
	{
		if ((rc = b2cuSetCounts(device, nb, ns, np))) return rc;
	}
	if (ns > m_shapesUploaded)
	{
		if ((rc = b2cuSetShapes(device, m_shapesUploaded, ns - m_shapesUploaded, m_shapes.data() + m_shapesUploaded)))
			return rc;
	}
	// new rows (tail) and edited rows (dirty range)
	int32 lo = std::min(m_bodyDirtyLo, m_bodiesUploaded), hi = std::max(m_bodyDirtyHi, nb - 1);
	if (nb > m_bodiesUploaded || m_bodyDirtyHi >= 0)
	{
		if (nb == m_bodiesUploaded) hi = m_bodyDirtyHi;
		if (lo <= hi)
		{
			// whole records of the range, assembled from the parts of the host mirror; the sweep starts are the device's
			RefreshSweepStarts();
			if (m_uploadRows.size() < (size_t)(hi - lo + 1)) m_uploadRows.resize((size_t)(hi - lo + 1));
			for (int32 i = lo; i <= hi; ++i) BodyView(i).ToRecord(&m_uploadRows[(size_t)(i - lo)]);
			if ((rc = b2cuSetBodies(device, lo, hi - lo + 1, m_uploadRows.data()))) return rc;
		}
	}
	if (m_forceDirtyHi >= 0)
	{
		// applied forces of rows that were not uploaded as a whole above
		const int32 flo = m_forceDirtyLo, fhi = std::min(m_forceDirtyHi, nb - 1);
		if (flo <= fhi)
		{
			const size_t n = (size_t)(fhi - flo + 1);
			if (m_forceRows.size() < 3 * n) m_forceRows.resize(3 * n);
			for (size_t i = 0; i < n; ++i)
			{
				const b2BodyProps& pr = m_props[(size_t)flo + i];
				m_forceRows[3 * i] = pr.fx;
				m_forceRows[3 * i + 1] = pr.fy;
				m_forceRows[3 * i + 2] = pr.torque;
			}
			if ((rc = b2cuSetBodyForces(device, flo, (int32)n, m_forceRows.data()))) return rc;
		}
	}
	lo = std::min(m_proxyDirtyLo, m_proxiesUploaded);
	hi = std::max(m_proxyDirtyHi, np - 1);
	if (np > m_proxiesUploaded || m_proxyDirtyHi >= 0)
	{
		if (np == m_proxiesUploaded) hi = m_proxyDirtyHi;
		if (lo <= hi && (rc = b2cuSetProxies(device, lo, hi - lo + 1, m_proxies.data() + lo))) return rc;
		// the MOVED flag has been handed to the device's move buffer
		for (int32 i = lo; i <= hi; ++i) m_proxies[i].flags &= ~(uint16)(B2CU_PROXY_MOVED | B2CU_PROXY_NEW | B2CU_PROXY_REFILTER);
	}
	if (m_jointsDirty || (m_fullUpload && !m_joints.empty()))
	{
		// the joint table as a whole.  The objects must hold the CURRENT accumulated impulses and limit states: a full
		// upload can be requested by code that never looked at the joints (DestroyFixture, SetType on another body)
		if (device == m_device) RefreshJoints();
		std::vector<b2cuJoint> rows(m_joints.size());
		for (size_t i = 0; i < m_joints.size(); ++i) m_joints[i]->WriteRecord(&rows[i]);
		if ((rc = b2cuSetJoints(device, (int32)rows.size(), rows.empty() ? nullptr : rows.data()))) return rc;
		m_jointsDirty = false;
		m_jointsStale = false;
	}
	if (m_fullUpload)
	{
		if ((rc = b2cuSetContacts(device, (int32)m_contactRecords.size(), m_contactRecords.data()))) return rc;
		if ((rc = b2cuSetInvDt0(device, m_inv_dt0))) return rc;
		m_fullUpload = false;
	}
	m_bodiesUploaded = nb;
	m_proxiesUploaded = np;
	m_shapesUploaded = ns;
	m_bodyDirtyLo = m_proxyDirtyLo = m_forceDirtyLo = INT32_MAX;
	m_bodyDirtyHi = m_proxyDirtyHi = m_forceDirtyHi = -1;
	m_newFixture = false;
	return 0;
}

// b2cuPairFilterFn for a user b2ContactFilter: the same call b2ContactManager::AddPair makes (b2ContactManager.cpp:
// 280-285), fixture A = the fixture with the lower proxy id
int b2World::PairFilterThunk(void* user, const b2cuContactKey* keys, int32_t count, uint8_t* keep)
{
	b2World* self = static_cast<b2World*>(user);
	if (self->m_contactFilter == nullptr) return 0;
	const size_t n = self->m_fixtures.size();
	for (int32_t i = 0; i < count; ++i)
	{
		size_t a = (size_t)(keys[i] >> 32), b = (size_t)(keys[i] & 0xFFFFFFFFull);
		if (a >= n || b >= n) return 1;
		keep[i] = self->m_contactFilter->ShouldCollide(self->m_fixtures[a], self->m_fixtures[b], 0) ? 1 : 0;
	}
	return 0;
}

// FinishCollide's callback order (b2ContactManager.cpp:420-433): Immediate callbacks first, then the deferred
// BeginContact calls in key order, then the deferred EndContact calls in key order.
void b2World::DispatchEvents(b2cuWorld* device)
{
	static const bool timing = getenv("B2H_EVENT_TIMING") != nullptr;
	typedef std::chrono::steady_clock Clock;
	Clock::time_point t0 = Clock::now();
	double queryMs = 0.0, makeMs = 0.0;
	std::vector<b2cuContactKey>* keys = m_eventKeys;
	std::vector<b2cuContact>* recs = m_eventRecs;
	std::vector<b2Contact>* contacts = m_eventContacts;
	std::vector<char>* deferred = m_eventDeferred;
	int32 counts[2] = {0, 0};
	for (int32 kind = 0; kind < 2; ++kind)
	{
		int32 n = 0;
		b2cuGetEventContacts(device, kind, 0, nullptr, nullptr, &n);
		counts[kind] = n;
		if ((int32)keys[kind].size() < n)
		{
			keys[kind].resize(n);
			recs[kind].resize(n);
			contacts[kind].resize(n);
			deferred[kind].resize(n);
		}
		if (n == 0) continue;
		Clock::time_point ta = Clock::now();
		b2cuGetEventContacts(device, kind, n, keys[kind].data(), recs[kind].data(), &n);
		Clock::time_point tb = Clock::now();
		for (int32 i = 0; i < n; ++i)
		{
			// the fixture table is far larger than the caches and the events hit it at random: ask ahead
			if (i + 8 < n)
			{
				__builtin_prefetch(&m_fixtures[recs[kind][i + 8].proxyA]);
				__builtin_prefetch(&m_fixtures[recs[kind][i + 8].proxyB]);
			}
			// ... and, once those table entries have arrived, for the fixture objects they point to
			if (i + 4 < n)
			{
				__builtin_prefetch(m_fixtures[recs[kind][i + 4].proxyA]);
				__builtin_prefetch(m_fixtures[recs[kind][i + 4].proxyB]);
			}
			MakeContact(&contacts[kind][i], recs[kind][i]);
		}
		Clock::time_point tc = Clock::now();
		queryMs += std::chrono::duration<double, std::milli>(tb - ta).count();
		makeMs += std::chrono::duration<double, std::milli>(tc - tb).count();
	}
	if (timing)
		fprintf(stderr, "[b2h events] begin %d end %d: query %.3f ms, make %.3f ms, alloc+rest %.3f ms\n", counts[0],
		        counts[1], queryMs, makeMs,
		        std::chrono::duration<double, std::milli>(Clock::now() - t0).count() - queryMs - makeMs);
	for (int32 i = 0; i < counts[0]; ++i)
		deferred[0][i] = m_contactListener->BeginContactImmediate(&contacts[0][i], 0) ? 1 : 0;
	for (int32 i = 0; i < counts[1]; ++i)
		deferred[1][i] = m_contactListener->EndContactImmediate(&contacts[1][i], 0) ? 1 : 0;
	for (int32 i = 0; i < counts[0]; ++i)
		if (deferred[0][i]) m_contactListener->BeginContact(&contacts[0][i]);
	for (int32 i = 0; i < counts[1]; ++i)
		if (deferred[1][i]) m_contactListener->EndContact(&contacts[1][i]);

	// Calls made from inside the time-of-impact sub-steps (b2Contact::Update from b2World::StepSolveTOI, reference
	// b2World.cpp:866 and :936): single-threaded there, so the Immediate call and the deferred call come back to back,
	// event by event, after all callbacks of the discrete part of the step.
	int32 nToi = 0;
	b2cuGetToiEvents(device, 0, nullptr, nullptr, nullptr, &nToi);
	if (nToi > 0)
	{
		std::vector<b2cuContactKey> toiKeys((size_t)nToi);
		std::vector<int32_t> toiKinds((size_t)nToi);
		std::vector<b2cuContact> toiRecs((size_t)nToi);
		if (b2cuGetToiEvents(device, nToi, toiKeys.data(), toiKinds.data(), toiRecs.data(), &nToi) != B2CU_OK) return;
		for (int32 i = 0; i < nToi; ++i)
		{
			b2Contact c;
			MakeContact(&c, toiRecs[(size_t)i]);
			if (toiKinds[(size_t)i] == B2CU_EVENT_BEGIN)
			{
				if (m_contactListener->BeginContactImmediate(&c, 0)) m_contactListener->BeginContact(&c);
			}
			else
			{
				if (m_contactListener->EndContactImmediate(&c, 0)) m_contactListener->EndContact(&c);
			}
		}
	}
}

// b2cuPreSolveFn: the PreSolve calls of b2Contact::Update / b2ContactManager::FinishCollide, made from inside the
// device step.  Immediate calls first, deferred ones in key order; contacts the callbacks disabled are switched off
// on the device before the solver runs.
int b2World::PreSolveThunk(void* user, b2cuWorld* device)
{
	b2World* self = static_cast<b2World*>(user);
	if (self->m_contactListener == nullptr) return 0;
	int32 n = 0;
	if (b2cuGetPreSolveContacts(device, 0, nullptr, nullptr, &n) != B2CU_OK) return 1;
	if (n <= 0) return 0;
	std::vector<b2cuContact> recs((size_t)n);
	std::vector<b2cuManifold> olds((size_t)n);
	if (b2cuGetPreSolveContacts(device, n, recs.data(), olds.data(), &n) != B2CU_OK) return 1;
	std::vector<b2Contact> contacts((size_t)n);
	std::vector<b2Manifold> oldManifolds((size_t)n);
	std::vector<char> deferred((size_t)n, 0);
	for (int32 i = 0; i < n; ++i)
	{
		self->MakeContact(&contacts[i], recs[i]);
		const b2cuManifold& m = olds[i];
		b2Manifold& o = oldManifolds[i];
		o.localNormal.Set(m.localNormal[0], m.localNormal[1]);
		o.localPoint.Set(m.localPoint[0], m.localPoint[1]);
		for (int32 j = 0; j < b2_maxManifoldPoints; ++j)
		{
			o.points[j].localPoint.Set(m.points[j].localPoint[0], m.points[j].localPoint[1]);
			o.points[j].normalImpulse = m.points[j].normalImpulse;
			o.points[j].tangentImpulse = m.points[j].tangentImpulse;
			o.points[j].id.key = m.id[j];
		}
		o.type = (b2Manifold::Type)m.type;
		o.pointCount = m.pointCount;
		deferred[i] = self->m_contactListener->PreSolveImmediate(&contacts[i], &o, 0) ? 1 : 0;
	}
	for (int32 i = 0; i < n; ++i)
		if (deferred[i]) self->m_contactListener->PreSolve(&contacts[i], &oldManifolds[i]);
	std::vector<b2cuContactKey> off;
	for (int32 i = 0; i < n; ++i)
		if (!contacts[i].IsEnabled()) off.push_back(contacts[i].GetKey());
	if (!off.empty() && b2cuDisableContacts(device, (int32)off.size(), off.data()) != B2CU_OK) return 1;
	return 0;
}

// b2Island::Report (reference b2Island.cpp:533-570) after the fact: the accumulated impulses of the step are what
// StoreImpulses left in the manifolds of the solved contacts
void b2World::DispatchPostSolve(b2cuWorld* device)
{
	int32 n = 0;
	if (b2cuGetSolverOrder(device, 0, nullptr, nullptr, &n) != B2CU_OK || n <= 0) return;
	std::vector<b2cuContactKey> keys((size_t)n);
	if (b2cuGetSolverOrder(device, n, keys.data(), nullptr, &n) != B2CU_OK) return;
	// deferred PostSolve calls run in contact-key order (b2DeferredPostSolveLessThan, b2ContactManager.cpp:84-87)
	std::sort(keys.begin(), keys.end());
	std::vector<b2cuContact> recs((size_t)n);
	if (b2cuGetContactsByKey(device, n, keys.data(), recs.data()) != B2CU_OK) return;
	std::vector<b2Contact> contacts((size_t)n);
	std::vector<b2ContactImpulse> impulses((size_t)n);
	std::vector<char> deferred((size_t)n, 0);
	for (int32 i = 0; i < n; ++i)
	{
		MakeContact(&contacts[i], recs[i]);
		b2ContactImpulse& imp = impulses[i];
		imp.count = recs[i].manifold.pointCount;
		for (int32 j = 0; j < b2_maxManifoldPoints; ++j)
		{
			imp.normalImpulses[j] = j < imp.count ? recs[i].manifold.points[j].normalImpulse : 0.0f;
			imp.tangentImpulses[j] = j < imp.count ? recs[i].manifold.points[j].tangentImpulse : 0.0f;
		}
		deferred[i] = m_contactListener->PostSolveImmediate(&contacts[i], &imp, 0) ? 1 : 0;
	}
	for (int32 i = 0; i < n; ++i)
		if (deferred[i]) m_contactListener->PostSolve(&contacts[i], &impulses[i]);
}

int32 b2World::AfterDeviceStep(b2cuWorld* device, const b2cuStepInfo& info, bool downloadBodies, bool dispatchEvents,
                               float32* hostMs, bool reportPostSolve)
{
	typedef std::chrono::steady_clock Clock;
	Clock::time_point t0 = Clock::now();
	m_contactCount = info.contactCount;
	memcpy(&m_profile, &info, sizeof(b2Profile)); // the first 13 floats of b2cuStepInfo are the b2Profile fields
	// downloadBodies: b2cuStep has already written the records into m_states (b2cuSetBodyMirror)
	m_bodiesStale = !downloadBodies;
	m_sweepStartsStale = true;
	if (!m_forced.empty())
	{
		// the device cleared the forces (b2World::ClearForces at the end of Step, b2World.cpp:1688-1691) or, without
		// auto-clear, zeroed those of the bodies it put to sleep (b2Body::SetAwake(false), b2Body.h:704-711)
		if (!m_clearForces) RefreshBodies();
		size_t kept = 0;
		for (size_t i = 0; i < m_forced.size(); ++i)
		{
			int32 b = m_forced[i];
			if (m_clearForces || !(m_states[b].flags & B2CU_BODY_AWAKE))
				m_props[b].fx = m_props[b].fy = m_props[b].torque = 0.0f;
			else
				m_forced[kept++] = b;
		}
		m_forced.resize(kept);
	}
	m_proxiesStale = true;
	if (!m_joints.empty()) m_jointsStale = true;
	InvalidateSnapshots();
	Clock::time_point t1 = Clock::now();
	if (hostMs) hostMs[0] = std::chrono::duration<float, std::milli>(t1 - t0).count();
	if (dispatchEvents && m_contactListener && (info.beginCount > 0 || info.endCount > 0 || info.toiEventCount > 0))
	{
		// callbacks may read bodies: make sure the mirror is current
		RefreshBodies();
		bool wasLocked = m_locked;
		m_locked = true;
		DispatchEvents(device);
		m_locked = wasLocked;
	}
	if (reportPostSolve && m_contactListener && info.constraintCount > 0)
	{
		RefreshBodies();
		bool wasLocked = m_locked;
		m_locked = true;
		DispatchPostSolve(device);
		m_locked = wasLocked;
	}
	if (hostMs) hostMs[1] = std::chrono::duration<float, std::milli>(Clock::now() - t1).count();
	return 0;
}
