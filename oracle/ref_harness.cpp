/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * Headless driver around the UNMODIFIED reference sources (compiled from /root/reference by
 * oracle/build_ref.py; nothing from the reference is copied into this repository).  This file is
 * compiled with -fno-access-control so that it can read the reference's private state (proxy ids,
 * fat AABBs, contact flags) and call its private phase drivers, as SURVEY.md 8c describes.
 *
 * Two ways of stepping:
 *   b2ref_step          -- the reference's own b2World::Step + b2ThreadPoolTaskExecutor.
 *   b2ref_step_ordered  -- the same phase sequence (Box2D/Dynamics/b2World.cpp:1613-1710), but the island
 *                          traversal of b2World::Solve (:1200-1371) is re-driven here so that every island's
 *                          contact array can be put into a caller-supplied order before the reference's own
 *                          b2Island::Solve runs on it.  Sequential Gauss-Seidel in colour order is what a
 *                          coloured parallel Gauss-Seidel computes, so this is the oracle for the GPU solver.
 *
 * sinf/cosf/sincosf are defined here (in terms of oracle/b2o_math.c) and the library is linked with
 * -Bsymbolic-functions, so every b2Rot::Set inside the reference uses the same sincos as the device.
 */
#include "ref_harness.h"
#include "b2o_math.h"

#include "Box2D/Box2D.h"
#include "Box2D/Dynamics/b2Island.h"
#include "Box2D/Dynamics/Contacts/b2ContactSolver.h"
#include "Box2D/MT/b2ThreadPool.h"

#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <vector>

#ifndef B2REF_STOCK_LIBM
extern "C" {
float sinf(float x) noexcept
{
	float s, c;
	b2o_sincosf(x, &s, &c);
	return s;
}
float cosf(float x) noexcept
{
	float s, c;
	b2o_sincosf(x, &s, &c);
	return c;
}
void sincosf(float x, float* s, float* c) noexcept
{
	b2o_sincosf(x, s, c);
}
}
#endif

namespace
{

inline int32 FixtureIndex(const b2Fixture* f)
{
	return (int32)(intptr_t)f->GetUserData();
}

// Dense proxy ids: one per (fixture, child) in creation order -- a chain fixture has one per segment.  Filled by
// b2ref_build; the contact keys, the exports and the event lists all use these ids.
std::vector<int32>* g_proxyBaseOf(const b2World* world);
inline int32 ProxyIndex(const b2Fixture* f, int32 child)
{
	return (*g_proxyBaseOf(f->GetBody()->GetWorld()))[FixtureIndex(f)] + child;
}

inline uint64_t MakeKey(int32 a, int32 b)
{
	uint32_t lo = (uint32_t)std::min(a, b), hi = (uint32_t)std::max(a, b);
	return ((uint64_t)lo << 32) | hi;
}

inline uint64_t ContactKey(const b2Contact* c)
{
	return MakeKey(ProxyIndex(c->GetFixtureA(), c->GetChildIndexA()), ProxyIndex(c->GetFixtureB(), c->GetChildIndexB()));
}

class RecordingListener : public b2ContactListener
{
public:
	bool BeginContactImmediate(b2Contact*, uint32) override { return true; }
	bool EndContactImmediate(b2Contact*, uint32) override { return true; }
	bool PreSolveImmediate(b2Contact*, const b2Manifold*, uint32) override { return preSolveModulus > 0; }
	void PreSolve(b2Contact* c, const b2Manifold* oldManifold) override
	{
		// test rule: contacts whose key is a multiple of the modulus are switched off every step
		uint64_t key = ContactKey(c);
		if (key % (uint64_t)preSolveModulus == 0) c->SetEnabled(false);
		preSolveDigest += key * 0x9E3779B97F4A7C15ull + (uint64_t)oldManifold->pointCount * 7u +
		                  (uint64_t)c->GetManifold()->pointCount;
		++preSolveCount;
	}
	bool PostSolveImmediate(b2Contact*, const b2ContactImpulse*, uint32) override { return recordPostSolve; }
	void PostSolve(b2Contact* c, const b2ContactImpulse* impulse) override
	{
		// order-independent digest of (key, count, impulses): the deferred calls of the reference and of the GPU path come in
		// the same key order, but a sum of bit patterns does not care
		uint64_t d = ContactKey(c) * 0x9E3779B97F4A7C15ull + (uint64_t)impulse->count;
		for (int32 j = 0; j < impulse->count; ++j)
		{
			uint32_t n, t;
			memcpy(&n, &impulse->normalImpulses[j], 4);
			memcpy(&t, &impulse->tangentImpulses[j], 4);
			d += ((uint64_t)n << 32 | t) * (uint64_t)(2 * j + 3);
		}
		postSolveDigest += d;
		++postSolveCount;
	}
	void BeginContact(b2Contact* c) override { begins.push_back(ContactKey(c)); }
	void EndContact(b2Contact* c) override { ends.push_back(ContactKey(c)); }

	std::vector<uint64_t> begins, ends;
	int preSolveModulus = 0;
	uint64_t preSolveDigest = 0;
	int64_t preSolveCount = 0;
	bool recordPostSolve = false;
	uint64_t postSolveDigest = 0;
	int64_t postSolveCount = 0;
};

} // namespace

struct b2refWorld
{
	b2World* world;
	b2ThreadPoolTaskExecutor* executor;
	RecordingListener listener;
	std::vector<b2Body*> bodies;
	std::vector<b2Fixture*> fixtures;
	std::vector<int32> proxyBase;                          // first proxy id of each fixture
	std::vector<std::pair<b2Fixture*, int32> > proxies;    // (fixture, child) of each proxy id
	std::vector<b2Joint*> joints;                          // b2ref_set_joints, table order
	std::unordered_map<const b2Joint*, int32> jointRank;   // position in the caller's solve order
	std::unordered_map<const b2Joint*, b2Vec2> rawAxis;    // b2PrismaticJointDef::localAxisA as given (the joint keeps it normalised)
};

namespace
{
std::unordered_map<const b2World*, b2refWorld*> g_worlds;
std::vector<int32>* g_proxyBaseOf(const b2World* world) { return &g_worlds[world]->proxyBase; }
} // namespace

extern "C" {

b2refWorld* b2ref_create(float gx, float gy, uint32_t worldFlags, int32_t threads)
{
	b2refWorld* w = new b2refWorld;
	w->world = new b2World(b2Vec2(gx, gy));
	w->world->SetAllowSleeping((worldFlags & B2CU_WORLD_ALLOW_SLEEP) != 0);
	w->world->SetWarmStarting((worldFlags & B2CU_WORLD_WARM_STARTING) != 0);
	w->world->SetContinuousPhysics((worldFlags & B2CU_WORLD_CONTINUOUS) != 0);
	w->world->SetSubStepping((worldFlags & B2CU_WORLD_SUB_STEPPING) != 0);
	w->world->SetAutoClearForces((worldFlags & B2CU_WORLD_CLEAR_FORCES) != 0);
	w->world->SetContactListener(&w->listener);
	g_worlds[w->world] = w;
	b2ThreadPoolOptions options;
	options.totalThreadCount = threads;
	w->executor = new b2ThreadPoolTaskExecutor(options);
	return w;
}

void b2ref_destroy(b2refWorld* w)
{
	if (w == nullptr)
	{
		return;
	}
	g_worlds.erase(w->world);
	delete w->world;
	delete w->executor;
	delete w;
}

static void MakeShape(const b2refShapeDef& sd, b2CircleShape& circle, b2EdgeShape& edge, b2PolygonShape& poly,
                      b2ChainShape& chain, const b2Shape** out)
{
	switch (sd.kind)
	{
	case B2REF_SHAPE_CHAIN:
	{
		b2Vec2 vs[b2_maxPolygonVertices];
		for (int32 i = 0; i < sd.count; ++i)
		{
			vs[i].Set(sd.v[i][0], sd.v[i][1]);
		}
		if (sd.flags & 1) chain.CreateLoop(vs, sd.count);
		else chain.CreateChain(vs, sd.count);
		*out = &chain;
		break;
	}
	case B2REF_SHAPE_CIRCLE:
		circle.m_radius = sd.radius;
		circle.m_p.Set(sd.v[0][0], sd.v[0][1]);
		*out = &circle;
		break;
	case B2REF_SHAPE_EDGE:
		edge.Set(b2Vec2(sd.v[0][0], sd.v[0][1]), b2Vec2(sd.v[1][0], sd.v[1][1]));
		if (sd.flags & B2CU_EDGE_HAS_VERTEX0)
		{
			edge.m_vertex0.Set(sd.v[2][0], sd.v[2][1]);
			edge.m_hasVertex0 = true;
		}
		if (sd.flags & B2CU_EDGE_HAS_VERTEX3)
		{
			edge.m_vertex3.Set(sd.v[3][0], sd.v[3][1]);
			edge.m_hasVertex3 = true;
		}
		*out = &edge;
		break;
	case B2REF_SHAPE_POLYGON:
	{
		b2Vec2 vs[b2_maxPolygonVertices];
		for (int32 i = 0; i < sd.count; ++i)
		{
			vs[i].Set(sd.v[i][0], sd.v[i][1]);
		}
		poly.Set(vs, sd.count);
		*out = &poly;
		break;
	}
	case B2REF_SHAPE_BOX:
		if (sd.flags & 1)
		{
			poly.SetAsBox(sd.v[0][0], sd.v[0][1], b2Vec2(sd.v[1][0], sd.v[1][1]), sd.v[2][0]);
		}
		else
		{
			poly.SetAsBox(sd.v[0][0], sd.v[0][1]);
		}
		*out = &poly;
		break;
	default: /* B2REF_SHAPE_RAW */
		poly.m_count = sd.count;
		for (int32 i = 0; i < sd.count; ++i)
		{
			poly.m_vertices[i].Set(sd.v[i][0], sd.v[i][1]);
			poly.m_normals[i].Set(sd.n[i][0], sd.n[i][1]);
		}
		poly.m_centroid.Set(sd.centroid[0], sd.centroid[1]);
		poly.m_radius = sd.radius;
		*out = &poly;
		break;
	}
}

int b2ref_build(b2refWorld* w, int32_t bodyCount, const b2refBodyDef* bodies, int32_t shapeCount,
                const b2refShapeDef* shapes, int32_t fixtureCount, const b2refFixtureDef* fixtures)
{
	int32 f = 0;
	for (int32 i = 0; i < bodyCount; ++i)
	{
		const b2refBodyDef& d = bodies[i];
		b2BodyDef bd;
		bd.type = (b2BodyType)d.type;
		bd.position.Set(d.px, d.py);
		bd.angle = d.angle;
		bd.linearVelocity.Set(d.vx, d.vy);
		bd.angularVelocity = d.w;
		bd.linearDamping = d.linearDamping;
		bd.angularDamping = d.angularDamping;
		bd.gravityScale = d.gravityScale;
		bd.allowSleep = (d.flags & B2REF_BODY_ALLOW_SLEEP) != 0;
		bd.awake = (d.flags & B2REF_BODY_AWAKE) != 0;
		bd.fixedRotation = (d.flags & B2REF_BODY_FIXED_ROTATION) != 0;
		bd.bullet = (d.flags & B2REF_BODY_BULLET) != 0;
		bd.active = (d.flags & B2REF_BODY_ACTIVE) != 0;
		bd.userData = (void*)(intptr_t)w->bodies.size();
		b2Body* body = w->world->CreateBody(&bd);
		w->bodies.push_back(body);

		while (f < fixtureCount && fixtures[f].body == i)
		{
			const b2refFixtureDef& fd = fixtures[f];
			if (fd.shape < 0 || fd.shape >= shapeCount)
			{
				return -1;
			}
			b2CircleShape circle;
			b2EdgeShape edge;
			b2PolygonShape poly;
			b2ChainShape chain;
			const b2Shape* shape = nullptr;
			MakeShape(shapes[fd.shape], circle, edge, poly, chain, &shape);

			b2FixtureDef def;
			def.shape = shape;
			def.density = fd.density;
			def.friction = fd.friction;
			def.restitution = fd.restitution;
			def.isSensor = (fd.flags & B2CU_PROXY_SENSOR) != 0;
			def.thickShape = (fd.flags & B2CU_PROXY_THICK) != 0;
			def.filter.categoryBits = fd.categoryBits;
			def.filter.maskBits = fd.maskBits;
			def.filter.groupIndex = fd.groupIndex;
			def.userData = (void*)(intptr_t)w->fixtures.size();
			b2Fixture* fixture = body->CreateFixture(&def);
			w->fixtures.push_back(fixture);
			w->proxyBase.push_back((int32)w->proxies.size());
			for (int32 child = 0; child < fixture->GetShape()->GetChildCount(); ++child)
			{
				w->proxies.push_back(std::make_pair(fixture, child));
			}
			++f;
		}
	}
	return f == fixtureCount ? 0 : -2;
}

void b2ref_step(b2refWorld* w, float dt, int32_t velocityIterations, int32_t positionIterations)
{
	w->listener.begins.clear();
	w->listener.ends.clear();
	w->world->Step(dt, velocityIterations, positionIterations, *w->executor);
}

/* Island traversal + ordered solve.  Follows the semantics of b2World::Solve (b2World.cpp:1166-1431):
 * seeds are awake, active, non-static bodies not yet in an island; the search crosses enabled, touching,
 * non-sensor contacts; static bodies join an island but are not crossed and may join several islands. */
static int SolveOrdered(b2refWorld* rw, const b2TimeStep& step, const std::unordered_map<uint64_t, int32>& rank)
{
	b2World* world = rw->world;
	b2ContactManager& cm = world->m_contactManager;
	b2ContactManagerPerThreadData& td = cm.m_perThreadData[0];
	int unranked = 0;

	world->SetMtLock(b2World::e_mtLocked | b2World::e_mtSolveLocked);

	std::vector<b2Body*> islandBodies;
	std::vector<b2Contact*> islandContacts;
	std::vector<b2Joint*> islandJoints;
	std::vector<b2Body*> pending;
	std::vector<b2Velocity> velocities;
	std::vector<b2Position> positions;

	for (uint32 s = 0; s < world->m_nonStaticBodies.size(); ++s)
	{
		b2Body* seed = world->m_nonStaticBodies[s];
		if ((seed->m_flags & b2Body::e_islandFlag) || !seed->IsAwake() || !seed->IsActive())
		{
			continue;
		}

		islandBodies.clear();
		islandContacts.clear();
		islandJoints.clear();
		pending.clear();
		pending.push_back(seed);
		seed->m_flags |= b2Body::e_islandFlag;

		while (!pending.empty())
		{
			b2Body* b = pending.back();
			pending.pop_back();
			islandBodies.push_back(b);

			if (b->GetType() == b2_staticBody)
			{
				continue;
			}

			b->m_flags |= b2Body::e_awakeFlag;

			for (b2ContactEdge* ce = b->m_contactList; ce; ce = ce->next)
			{
				b2Contact* contact = ce->contact;
				if (contact->m_flags & b2Contact::e_islandFlag)
				{
					continue;
				}
				contact->m_flags &= ~b2Contact::e_inactiveFlag;
				if (!contact->IsEnabled() || !contact->IsTouching())
				{
					continue;
				}
				if (contact->m_fixtureA->m_isSensor || contact->m_fixtureB->m_isSensor)
				{
					continue;
				}
				islandContacts.push_back(contact);
				contact->m_flags |= b2Contact::e_islandFlag;

				b2Body* other = ce->other;
				if (other->m_flags & b2Body::e_islandFlag)
				{
					continue;
				}
				pending.push_back(other);
				other->m_flags |= b2Body::e_islandFlag;
			}
			/* joints (b2World.cpp:1291-1318) */
			for (b2JointEdge* je = b->m_jointList; je; je = je->next)
			{
				if (je->joint->m_islandFlag)
				{
					continue;
				}
				b2Body* other = je->other;
				if (!other->IsActive())
				{
					continue;
				}
				islandJoints.push_back(je->joint);
				je->joint->m_islandFlag = true;
				if (other->m_flags & b2Body::e_islandFlag)
				{
					continue;
				}
				pending.push_back(other);
				other->m_flags |= b2Body::e_islandFlag;
			}
		}

		for (size_t j = 0; j < islandBodies.size(); ++j)
		{
			if (islandBodies[j]->GetType() == b2_staticBody)
			{
				islandBodies[j]->m_flags &= ~b2Body::e_islandFlag;
			}
		}

		/* the one deliberate difference from b2World::Solve: the caller's contact order */
		struct Ranked
		{
			int32 r;
			uint64_t key;
			b2Contact* c;
		};
		std::vector<Ranked> ranked(islandContacts.size());
		for (size_t j = 0; j < islandContacts.size(); ++j)
		{
			uint64_t key = ContactKey(islandContacts[j]);
			auto it = rank.find(key);
			int32 r;
			if (it == rank.end())
			{
				r = INT32_MAX;
				++unranked;
			}
			else
			{
				r = it->second;
			}
			ranked[j] = {r, key, islandContacts[j]};
		}
		std::sort(ranked.begin(), ranked.end(), [](const Ranked& a, const Ranked& b) {
			return a.r != b.r ? a.r < b.r : a.key < b.key;
		});
		for (size_t j = 0; j < ranked.size(); ++j)
		{
			islandContacts[j] = ranked[j].c;
		}

		velocities.resize(islandBodies.size());
		positions.resize(islandBodies.size());
		/* ... and the caller's joint order (joints it did not rank keep their discovery order, behind the ranked ones) */
		std::stable_sort(islandJoints.begin(), islandJoints.end(), [rw](const b2Joint* a, const b2Joint* b) {
			auto ia = rw->jointRank.find(a), ib = rw->jointRank.find(b);
			int32 ra = ia == rw->jointRank.end() ? INT32_MAX : ia->second;
			int32 rb = ib == rw->jointRank.end() ? INT32_MAX : ib->second;
			return ra < rb;
		});
		b2Island island((int32)islandBodies.size(), (int32)islandContacts.size(), (int32)islandJoints.size(),
		                islandBodies.data(), islandContacts.data(), islandJoints.data(), velocities.data(), positions.data());
		island.Solve(&td.m_profile, step, world->m_gravity, &world->m_stackAllocator, cm.m_contactListener, 0,
		             world->m_allowSleep, td.m_postSolves);
	}

	world->SetMtLock(0);
	return unranked;
}

int b2ref_step_ordered(b2refWorld* rw, float dt, int32_t velocityIterations, int32_t positionIterations,
                       int32_t orderCount, const uint64_t* keys)
{
	rw->listener.begins.clear();
	rw->listener.ends.clear();

	std::unordered_map<uint64_t, int32> rank;
	rank.reserve((size_t)orderCount * 2 + 1);
	for (int32 i = 0; i < orderCount; ++i)
	{
		rank[keys[i]] = i;
	}

	b2World* world = rw->world;
	b2TaskExecutor& executor = *rw->executor;

	memset(&world->m_profile, 0, sizeof(world->m_profile));
	memset(&world->m_contactManager.m_perThreadData[0].m_profile, 0, sizeof(b2Profile));

	b2TaskGroup* group = executor.AcquireTaskGroup();

	if (world->m_flags & b2World::e_newFixture)
	{
		world->FindNewContacts(executor, group);
		world->m_flags &= ~b2World::e_newFixture;
	}

	world->m_flags |= b2World::e_locked;

	world->Collide(executor, group);

	b2TimeStep step;
	step.dt = dt;
	step.velocityIterations = velocityIterations;
	step.positionIterations = positionIterations;
	step.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
	step.dtRatio = world->m_inv_dt0 * dt;
	step.warmStarting = world->m_warmStarting;

	int unranked = 0;
	if (world->m_stepComplete && step.dt > 0.0f)
	{
		unranked = SolveOrdered(rw, step, rank);
		world->m_contactManager.FinishSolve(executor, group, world->m_stackAllocator);
		world->SynchronizeFixtures(executor, group);
		world->FindNewContacts(executor, group);
		world->ClearPostSolve(executor, group);
	}

	if (world->m_continuousPhysics && step.dt > 0.0f)
	{
		world->SolveTOI(executor, group, step);
	}

	if (step.dt > 0.0f)
	{
		world->m_inv_dt0 = step.inv_dt;
	}

	if (world->m_flags & b2World::e_clearForces)
	{
		world->ClearForces(executor, group);
	}

	world->m_flags &= ~b2World::e_locked;
	executor.ReleaseTaskGroup(group);
	return unranked;
}

void b2ref_counts(b2refWorld* w, int32_t* bodyCount, int32_t* fixtureCount, int32_t* contactCount)
{
	*bodyCount = (int32_t)w->bodies.size();
	*fixtureCount = (int32_t)w->proxies.size(); /* proxies: one per (fixture, child) */
	*contactCount = w->world->GetContactCount();
}

float b2ref_inv_dt0(b2refWorld* w)
{
	return w->world->m_inv_dt0;
}

void b2ref_export_bodies(b2refWorld* w, b2cuBody* out)
{
	for (size_t i = 0; i < w->bodies.size(); ++i)
	{
		const b2Body* b = w->bodies[i];
		b2cuBody& o = out[i];
		o.px = b->m_xf.p.x;
		o.py = b->m_xf.p.y;
		o.qs = b->m_xf.q.s;
		o.qc = b->m_xf.q.c;
		o.cx = b->m_sweep.c.x;
		o.cy = b->m_sweep.c.y;
		o.a = b->m_sweep.a;
		o.c0x = b->m_sweep.c0.x;
		o.c0y = b->m_sweep.c0.y;
		o.a0 = b->m_sweep.a0;
		o.alpha0 = b->m_sweep.alpha0;
		o.lcx = b->m_sweep.localCenter.x;
		o.lcy = b->m_sweep.localCenter.y;
		o.vx = b->m_linearVelocity.x;
		o.vy = b->m_linearVelocity.y;
		o.w = b->m_angularVelocity;
		o.fx = b->m_force.x;
		o.fy = b->m_force.y;
		o.torque = b->m_torque;
		o.invMass = b->m_invMass;
		o.invI = b->m_invI;
		o.linearDamping = b->m_linearDamping;
		o.angularDamping = b->m_angularDamping;
		o.gravityScale = b->m_gravityScale;
		o.sleepTime = b->m_sleepTime;
		o.flags = (uint32_t)b->m_type | ((uint32_t)b->m_flags << 2);
	}
}

static void ExportShape(const b2Shape* shape, b2cuShape* o)
{
	memset(o, 0, sizeof(*o));
	o->radius = shape->m_radius;
	switch (shape->GetType())
	{
	case b2Shape::e_circle:
	{
		const b2CircleShape* s = (const b2CircleShape*)shape;
		o->type = B2CU_SHAPE_CIRCLE;
		o->count = 1;
		o->v[0][0] = s->m_p.x;
		o->v[0][1] = s->m_p.y;
		break;
	}
	case b2Shape::e_edge:
	{
		const b2EdgeShape* s = (const b2EdgeShape*)shape;
		o->type = B2CU_SHAPE_EDGE;
		o->count = 2;
		o->v[0][0] = s->m_vertex1.x;
		o->v[0][1] = s->m_vertex1.y;
		o->v[1][0] = s->m_vertex2.x;
		o->v[1][1] = s->m_vertex2.y;
		o->v[2][0] = s->m_vertex0.x;
		o->v[2][1] = s->m_vertex0.y;
		o->v[3][0] = s->m_vertex3.x;
		o->v[3][1] = s->m_vertex3.y;
		o->flags = (s->m_hasVertex0 ? B2CU_EDGE_HAS_VERTEX0 : 0) | (s->m_hasVertex3 ? B2CU_EDGE_HAS_VERTEX3 : 0);
		break;
	}
	case b2Shape::e_polygon:
	{
		const b2PolygonShape* s = (const b2PolygonShape*)shape;
		o->type = B2CU_SHAPE_POLYGON;
		o->count = s->m_count;
		for (int32 i = 0; i < s->m_count; ++i)
		{
			o->v[i][0] = s->m_vertices[i].x;
			o->v[i][1] = s->m_vertices[i].y;
			o->n[i][0] = s->m_normals[i].x;
			o->n[i][1] = s->m_normals[i].y;
		}
		o->centroid[0] = s->m_centroid.x;
		o->centroid[1] = s->m_centroid.y;
		break;
	}
	default:
		o->type = -1;
		break;
	}
}

void b2ref_export_shapes(b2refWorld* w, b2cuShape* out)
{
	for (size_t i = 0; i < w->proxies.size(); ++i)
	{
		const b2Shape* shape = w->proxies[i].first->GetShape();
		if (shape->GetType() == b2Shape::e_chain)
		{
			// the geometry of a chain proxy is its segment as an edge with ghost vertices (what b2ChainAnd*Contact collide)
			b2EdgeShape edge;
			((const b2ChainShape*)shape)->GetChildEdge(&edge, w->proxies[i].second);
			ExportShape(&edge, out + i);
			out[i].flags |= B2CU_EDGE_CHAIN_CHILD;
		}
		else
		{
			ExportShape(shape, out + i);
		}
	}
}

void b2ref_export_proxies(b2refWorld* w, b2cuProxy* out, int32_t* treeProxyIds)
{
	const b2BroadPhase& bp = w->world->m_contactManager.m_broadPhase;

	std::vector<uint8_t> moved(bp.m_tree.m_nodeCapacity, 0);
	for (uint32 i = 0; i < bp.m_moveBuffer.size(); ++i)
	{
		int32 id = bp.m_moveBuffer[i];
		if (id != b2BroadPhase::e_nullProxy)
		{
			moved[id] = 1;
		}
	}

	for (size_t i = 0; i < w->proxies.size(); ++i)
	{
		const b2Fixture* f = w->proxies[i].first;
		const int32 child = w->proxies[i].second;
		b2cuProxy& o = out[i];
		memset(&o, 0, sizeof(o));
		o.body = (int32_t)(intptr_t)f->GetBody()->GetUserData();
		o.shape = (int32_t)i;
		o.friction = f->m_friction;
		o.restitution = f->m_restitution;
		o.categoryBits = f->m_filter.categoryBits;
		o.maskBits = f->m_filter.maskBits;
		o.groupIndex = f->m_filter.groupIndex;
		o.flags = (uint16_t)((f->m_isSensor ? B2CU_PROXY_SENSOR : 0) | (f->IsThickShape() ? B2CU_PROXY_THICK : 0));
		o.fixture = w->proxyBase[FixtureIndex(f)];
		o.child = child;
		int32 treeId = -1;
		if (f->m_proxyCount <= child) o.flags |= B2CU_PROXY_INACTIVE; /* body inactive: its proxies are gone */
		if (f->m_proxyCount > child)
		{
			const b2FixtureProxy& p = f->m_proxies[child];
			treeId = p.proxyId;
			o.aabb[0] = p.aabb.lowerBound.x;
			o.aabb[1] = p.aabb.lowerBound.y;
			o.aabb[2] = p.aabb.upperBound.x;
			o.aabb[3] = p.aabb.upperBound.y;
			const b2AABB& fat = bp.GetFatAABB(p.proxyId);
			o.fat[0] = fat.lowerBound.x;
			o.fat[1] = fat.lowerBound.y;
			o.fat[2] = fat.upperBound.x;
			o.fat[3] = fat.upperBound.y;
			if (moved[p.proxyId])
			{
				o.flags |= B2CU_PROXY_MOVED;
				// e_newFixture is a world flag: the whole move buffer is processed ahead of the next step
				if (w->world->m_flags & b2World::e_newFixture) o.flags |= B2CU_PROXY_NEW;
			}
		}
		if (treeProxyIds)
		{
			treeProxyIds[i] = treeId;
		}
	}
}

static void ExportManifold(const b2Manifold& m, b2cuManifold* o)
{
	o->localNormal[0] = m.localNormal.x;
	o->localNormal[1] = m.localNormal.y;
	o->localPoint[0] = m.localPoint.x;
	o->localPoint[1] = m.localPoint.y;
	for (int32 i = 0; i < 2; ++i)
	{
		o->points[i].localPoint[0] = m.points[i].localPoint.x;
		o->points[i].localPoint[1] = m.points[i].localPoint.y;
		o->points[i].normalImpulse = m.points[i].normalImpulse;
		o->points[i].tangentImpulse = m.points[i].tangentImpulse;
		o->id[i] = m.points[i].id.key;
	}
	o->type = (int32_t)m.type;
	o->pointCount = m.pointCount;
}

int b2ref_export_contacts(b2refWorld* w, int32_t capacity, b2cuContact* out)
{
	std::vector<std::pair<uint64_t, const b2Contact*>> sorted;
	// b2ContactManager::AddToContactList links a new contact at the head of the world list, exactly as OnContactCreate
	// does with the two bodies' lists (b2ContactManager.cpp:530-556, :715-725): the position counted from the TAIL is a
	// creation rank that orders every body's contact list (newest first) -- the b2cuContact::stamp
	std::unordered_map<const b2Contact*, uint32_t> rankFromHead;
	for (const b2Contact* c = w->world->GetContactList(); c; c = c->GetNext())
	{
		rankFromHead[c] = (uint32_t)sorted.size();
		sorted.push_back(std::make_pair(ContactKey(c), c));
	}
	std::sort(sorted.begin(), sorted.end());
	int32_t n = (int32_t)sorted.size();
	for (int32_t i = 0; i < n && i < capacity; ++i)
	{
		const b2Contact* c = sorted[i].second;
		b2cuContact& o = out[i];
		memset(&o, 0, sizeof(o));
		o.proxyA = ProxyIndex(c->GetFixtureA(), c->GetChildIndexA());
		o.proxyB = ProxyIndex(c->GetFixtureB(), c->GetChildIndexB());
		o.flags = c->m_flags;
		o.friction = c->m_friction;
		o.restitution = c->m_restitution;
		o.tangentSpeed = c->m_tangentSpeed;
		o.toiCount = c->m_toiCount;
		o.toi = c->m_toi;
		ExportManifold(c->m_manifold, &o.manifold);
		o.stamp = (uint32_t)n - 1u - rankFromHead[c];
	}
	return n;
}

int b2ref_events(b2refWorld* w, int32_t kind, int32_t capacity, uint64_t* keys)
{
	const std::vector<uint64_t>& v = kind == B2CU_EVENT_BEGIN ? w->listener.begins : w->listener.ends;
	int32_t n = (int32_t)v.size();
	for (int32_t i = 0; i < n && i < capacity; ++i)
	{
		keys[i] = v[i];
	}
	return n;
}

int b2ref_toi_candidates(b2refWorld* w, int32_t capacity, uint64_t* keys)
{
	b2ContactManager& cm = w->world->m_contactManager;
	std::vector<uint64_t> v;
	for (uint32 i = 0; i < cm.m_toiCount; ++i)
	{
		v.push_back(ContactKey(cm.m_contacts[i]));
	}
	std::sort(v.begin(), v.end());
	for (size_t i = 0; i < v.size() && (int32_t)i < capacity; ++i)
	{
		keys[i] = v[i];
	}
	return (int)v.size();
}

/* First pass of b2World::SolveTOI on the state the last Step left (valid when that step ran with continuous physics
 * off, so that no sub-step has happened yet): the sequential b2World::FindMinToiContact (b2World.cpp:1582-1611), then
 * the time-of-impact cache it filled is cleared again as ClearPostSolveTOI would.  Returns 1 and the winner's key /
 * alpha when there is a candidate. */
int32_t b2ref_first_toi(b2refWorld* w, uint64_t* key, float* alpha)
{
	b2Contact* minContact = nullptr;
	float32 minAlpha = 1.0f;
	w->world->FindMinToiContact(&minContact, &minAlpha);
	b2ContactManager& cm = w->world->m_contactManager;
	for (uint32 i = 0; i < cm.m_toiCount; ++i)
	{
		cm.m_contacts[i]->m_flags &= ~b2Contact::e_toiFlag;
		cm.m_contacts[i]->m_toi = 1.0f;
	}
	*alpha = minAlpha;
	*key = minContact ? ContactKey(minContact) : ~0ull;
	return minContact != nullptr;
}

/* Joints from b2cuJoint records (revolute only), created in table order; replaces the joints made by an earlier call. */
int32_t b2ref_set_joints(b2refWorld* w, int32_t count, const b2cuJoint* joints)
{
	for (size_t i = 0; i < w->joints.size(); ++i)
	{
		w->world->DestroyJoint(w->joints[i]);
	}
	w->joints.clear();
	w->jointRank.clear();
	for (int32_t i = 0; i < count; ++i)
	{
		const b2cuJoint& j = joints[i];
		b2Joint* joint = nullptr;
		b2Body* bodyA = w->bodies[j.bodyA];
		b2Body* bodyB = w->bodies[j.bodyB];
		const bool collideConnected = (j.flags & B2CU_JOINT_COLLIDE_CONNECTED) != 0;
		if (j.type == B2CU_JOINT_REVOLUTE)
		{
			b2RevoluteJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.referenceAngle = j.referenceAngle;
			def.enableLimit = (j.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;
			def.lowerAngle = j.lowerAngle;
			def.upperAngle = j.upperAngle;
			def.enableMotor = (j.flags & B2CU_JOINT_ENABLE_MOTOR) != 0;
			def.motorSpeed = j.motorSpeed;
			def.maxMotorTorque = j.maxMotorTorque;
			b2RevoluteJoint* rj = (b2RevoluteJoint*)w->world->CreateJoint(&def);
			rj->m_impulse.Set(j.impulse[0], j.impulse[1], j.impulse[2]);
			rj->m_motorImpulse = j.motorImpulse;
			rj->m_limitState = (b2LimitState)j.limitState;
			joint = rj;
		}
		else if (j.type == B2CU_JOINT_PRISMATIC)
		{
			b2PrismaticJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.localAxisA.Set(j.axis[0], j.axis[1]);
			def.referenceAngle = j.referenceAngle;
			def.enableLimit = (j.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;
			def.lowerTranslation = j.lowerAngle;
			def.upperTranslation = j.upperAngle;
			def.enableMotor = (j.flags & B2CU_JOINT_ENABLE_MOTOR) != 0;
			def.motorSpeed = j.motorSpeed;
			def.maxMotorForce = j.maxMotorTorque;
			b2PrismaticJoint* pj = (b2PrismaticJoint*)w->world->CreateJoint(&def);
			pj->m_impulse.Set(j.impulse[0], j.impulse[1], j.impulse[2]);
			pj->m_motorImpulse = j.motorImpulse;
			pj->m_limitState = (b2LimitState)j.limitState;
			pj->m_axis.Set(j.lastSolve[0], j.lastSolve[1]);
			pj->m_perp.Set(j.lastSolve[2], j.lastSolve[3]);
			w->rawAxis[pj] = def.localAxisA;
			joint = pj;
		}
		else if (j.type == B2CU_JOINT_WHEEL)
		{
			b2WheelJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.localAxisA.Set(j.axis[0], j.axis[1]);
			def.enableMotor = (j.flags & B2CU_JOINT_ENABLE_MOTOR) != 0;
			def.maxMotorTorque = j.maxMotorTorque;
			def.motorSpeed = j.motorSpeed;
			def.frequencyHz = j.frequencyHz;
			def.dampingRatio = j.dampingRatio;
			b2WheelJoint* wj = (b2WheelJoint*)w->world->CreateJoint(&def);
			wj->m_impulse = j.impulse[0];
			wj->m_springImpulse = j.impulse[1];
			wj->m_motorImpulse = j.motorImpulse;
			wj->m_ax.Set(j.lastSolve[0], j.lastSolve[1]);
			wj->m_ay.Set(j.lastSolve[2], j.lastSolve[3]);
			wj->m_sAx = j.work[0];
			wj->m_sBx = j.work[1];
			joint = wj;
		}
		else if (j.type == B2CU_JOINT_ROPE)
		{
			b2RopeJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.maxLength = j.length;
			b2RopeJoint* rj = (b2RopeJoint*)w->world->CreateJoint(&def);
			rj->m_impulse = j.impulse[0];
			rj->m_state = (b2LimitState)j.limitState;
			rj->m_u.Set(j.lastSolve[0], j.lastSolve[1]);
			joint = rj;
		}
		else if (j.type == B2CU_JOINT_FRICTION)
		{
			b2FrictionJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.maxForce = j.length;
			def.maxTorque = j.maxMotorTorque;
			b2FrictionJoint* fj = (b2FrictionJoint*)w->world->CreateJoint(&def);
			fj->m_linearImpulse.Set(j.impulse[0], j.impulse[1]);
			fj->m_angularImpulse = j.impulse[2];
			joint = fj;
		}
		else if (j.type == B2CU_JOINT_MOTOR)
		{
			b2MotorJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.linearOffset.Set(j.axis[0], j.axis[1]);
			def.angularOffset = j.referenceAngle;
			def.maxForce = j.length;
			def.maxTorque = j.maxMotorTorque;
			def.correctionFactor = j.dampingRatio;
			b2MotorJoint* mj = (b2MotorJoint*)w->world->CreateJoint(&def);
			mj->m_linearImpulse.Set(j.impulse[0], j.impulse[1]);
			mj->m_angularImpulse = j.impulse[2];
			joint = mj;
		}
		else if (j.type == B2CU_JOINT_PULLEY)
		{
			b2PulleyJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.groundAnchorA.Set(j.axis[0], j.axis[1]);
			def.groundAnchorB.Set(j.lowerAngle, j.upperAngle);
			def.lengthA = j.length;
			def.lengthB = j.referenceAngle;
			def.ratio = j.motorSpeed;
			b2PulleyJoint* pj = (b2PulleyJoint*)w->world->CreateJoint(&def);
			pj->m_impulse = j.impulse[0];
			pj->m_uB.Set(j.lastSolve[0], j.lastSolve[1]);
			joint = pj;
		}
		else if (j.type == B2CU_JOINT_MOUSE)
		{
			b2MouseJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.target.Set(j.axis[0], j.axis[1]);
			def.maxForce = j.length;
			def.frequencyHz = j.frequencyHz;
			def.dampingRatio = j.dampingRatio;
			b2MouseJoint* mj = (b2MouseJoint*)w->world->CreateJoint(&def);
			/* the constructor derives the grabbed point from the target; the record states it */
			mj->m_localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			mj->m_impulse.Set(j.impulse[0], j.impulse[1]);
			joint = mj;
		}
		else if (j.type == B2CU_JOINT_GEAR)
		{
			/* the two joints it couples are earlier rows of the table */
			int32_t i1 = (int32_t)j.frequencyHz, i2 = (int32_t)j.dampingRatio;
			if (i1 < 0 || i2 < 0 || i1 >= i || i2 >= i)
			{
				return -3;
			}
			b2GearJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.joint1 = w->joints[i1];
			def.joint2 = w->joints[i2];
			def.ratio = j.motorSpeed;
			b2GearJoint* gj = (b2GearJoint*)w->world->CreateJoint(&def);
			/* the constructor measures the constant on the current transforms; a row that already has one states it */
			if (j.length != 0.0f || j.impulse[0] != 0.0f)
			{
				gj->m_constant = j.length;
			}
			gj->m_impulse = j.impulse[0];
			gj->m_JvAC.Set(j.lastSolve[0], j.lastSolve[1]);
			gj->m_JwA = j.lastSolve[2];
			joint = gj;
		}
		else if (j.type == B2CU_JOINT_DISTANCE)
		{
			b2DistanceJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.length = j.length;
			def.frequencyHz = j.frequencyHz;
			def.dampingRatio = j.dampingRatio;
			b2DistanceJoint* dj = (b2DistanceJoint*)w->world->CreateJoint(&def);
			dj->m_impulse = j.impulse[0];
			dj->m_u.Set(j.lastSolve[0], j.lastSolve[1]);
			joint = dj;
		}
		else if (j.type == B2CU_JOINT_WELD)
		{
			b2WeldJointDef def;
			def.bodyA = bodyA;
			def.bodyB = bodyB;
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(j.localAnchorA[0], j.localAnchorA[1]);
			def.localAnchorB.Set(j.localAnchorB[0], j.localAnchorB[1]);
			def.referenceAngle = j.referenceAngle;
			def.frequencyHz = j.frequencyHz;
			def.dampingRatio = j.dampingRatio;
			b2WeldJoint* wj = (b2WeldJoint*)w->world->CreateJoint(&def);
			wj->m_impulse.Set(j.impulse[0], j.impulse[1], j.impulse[2]);
			joint = wj;
		}
		else
		{
			return -1;
		}
		w->joints.push_back(joint);
		w->jointRank[joint] = i;
	}
	return 0;
}

/* the order in which b2ref_step_ordered solves the joints of an island: ids[k] is solved k-th */
void b2ref_set_joint_order(b2refWorld* w, int32_t count, const int32_t* ids)
{
	w->jointRank.clear();
	for (int32_t k = 0; k < count; ++k)
	{
		w->jointRank[w->joints[ids[k]]] = k;
	}
}

void b2ref_joint_set_motor(b2refWorld* w, int32_t joint, int32_t enable, float speed, float maxTorque)
{
	if (w->joints[joint]->GetType() == e_prismaticJoint)
	{
		b2PrismaticJoint* pj = (b2PrismaticJoint*)w->joints[joint];
		pj->EnableMotor(enable != 0);
		pj->SetMotorSpeed(speed);
		pj->SetMaxMotorForce(maxTorque);
		return;
	}
	b2RevoluteJoint* j = (b2RevoluteJoint*)w->joints[joint];
	j->EnableMotor(enable != 0);
	j->SetMotorSpeed(speed);
	j->SetMaxMotorTorque(maxTorque);
}

void b2ref_joint_set_limits(b2refWorld* w, int32_t joint, int32_t enable, float lower, float upper)
{
	if (w->joints[joint]->GetType() == e_prismaticJoint)
	{
		b2PrismaticJoint* pj = (b2PrismaticJoint*)w->joints[joint];
		pj->EnableLimit(enable != 0);
		pj->SetLimits(lower, upper);
		return;
	}
	b2RevoluteJoint* j = (b2RevoluteJoint*)w->joints[joint];
	j->EnableLimit(enable != 0);
	j->SetLimits(lower, upper);
}

void b2ref_joint_set_spring(b2refWorld* w, int32_t joint, float length, float frequencyHz, float dampingRatio)
{
	b2Joint* base = w->joints[joint];
	if (base->GetType() == e_distanceJoint)
	{
		b2DistanceJoint* j = (b2DistanceJoint*)base;
		j->SetLength(length);
		j->SetFrequency(frequencyHz);
		j->SetDampingRatio(dampingRatio);
	}
	else if (base->GetType() == e_weldJoint)
	{
		b2WeldJoint* j = (b2WeldJoint*)base;
		j->SetFrequency(frequencyHz);
		j->SetDampingRatio(dampingRatio);
	}
}

/* b2MouseJoint::SetTarget / b2MotorJoint::SetLinearOffset */
void b2ref_joint_set_target(b2refWorld* w, int32_t joint, float x, float y)
{
	b2Joint* base = w->joints[joint];
	if (base->GetType() == e_mouseJoint) ((b2MouseJoint*)base)->SetTarget(b2Vec2(x, y));
	else if (base->GetType() == e_motorJoint) ((b2MotorJoint*)base)->SetLinearOffset(b2Vec2(x, y));
}

void b2ref_destroy_joint(b2refWorld* w, int32_t joint)
{
	w->jointRank.erase(w->joints[joint]);
	w->rawAxis.erase(w->joints[joint]);
	w->world->DestroyJoint(w->joints[joint]);
	w->joints.erase(w->joints.begin() + joint);
}

/* per joint: reaction force x, y, reaction torque, motor torque (at inv_dt), joint angle, joint speed */
void b2ref_joint_readings(b2refWorld* w, float inv_dt, float* out6)
{
	for (size_t i = 0; i < w->joints.size(); ++i)
	{
		const b2Joint* base = w->joints[i];
		b2Vec2 f = base->GetReactionForce(inv_dt);
		float* o = out6 + 6 * i;
		o[0] = f.x;
		o[1] = f.y;
		o[2] = base->GetReactionTorque(inv_dt);
		o[3] = o[4] = o[5] = 0.0f;
		if (base->GetType() == e_revoluteJoint)
		{
			const b2RevoluteJoint* j = (const b2RevoluteJoint*)base;
			o[3] = j->GetMotorTorque(inv_dt);
			o[4] = j->GetJointAngle();
			o[5] = j->GetJointSpeed();
		}
		else if (base->GetType() == e_prismaticJoint)
		{
			const b2PrismaticJoint* j = (const b2PrismaticJoint*)base;
			o[3] = j->GetMotorForce(inv_dt);
			o[4] = j->GetJointTranslation();
			o[5] = j->GetJointSpeed();
		}
		else if (base->GetType() == e_wheelJoint)
		{
			const b2WheelJoint* j = (const b2WheelJoint*)base;
			o[3] = j->GetMotorTorque(inv_dt);
			o[4] = j->GetJointTranslation();
			o[5] = j->GetJointLinearSpeed();
		}
	}
}

void b2ref_export_joints(b2refWorld* w, b2cuJoint* out)
{
	for (size_t i = 0; i < w->joints.size(); ++i)
	{
		const b2Joint* base = w->joints[i];
		b2cuJoint& o = out[i];
		memset(&o, 0, sizeof(o));
		for (size_t b = 0; b < w->bodies.size(); ++b)
		{
			if (w->bodies[b] == base->m_bodyA) o.bodyA = (int32_t)b;
			if (w->bodies[b] == base->m_bodyB) o.bodyB = (int32_t)b;
		}
		o.flags = base->m_collideConnected ? B2CU_JOINT_COLLIDE_CONNECTED : 0;
		if (base->GetType() == e_revoluteJoint)
		{
			const b2RevoluteJoint* j = (const b2RevoluteJoint*)base;
			o.type = B2CU_JOINT_REVOLUTE;
			o.flags |= (j->m_enableLimit ? B2CU_JOINT_ENABLE_LIMIT : 0) | (j->m_enableMotor ? B2CU_JOINT_ENABLE_MOTOR : 0);
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.referenceAngle = j->m_referenceAngle;
			o.lowerAngle = j->m_lowerAngle;
			o.upperAngle = j->m_upperAngle;
			o.maxMotorTorque = j->m_maxMotorTorque;
			o.motorSpeed = j->m_motorSpeed;
			o.impulse[0] = j->m_impulse.x;
			o.impulse[1] = j->m_impulse.y;
			o.impulse[2] = j->m_impulse.z;
			o.motorImpulse = j->m_motorImpulse;
			o.limitState = (int32_t)j->m_limitState;
		}
		else if (base->GetType() == e_prismaticJoint)
		{
			const b2PrismaticJoint* j = (const b2PrismaticJoint*)base;
			o.type = B2CU_JOINT_PRISMATIC;
			o.flags |= (j->m_enableLimit ? B2CU_JOINT_ENABLE_LIMIT : 0) | (j->m_enableMotor ? B2CU_JOINT_ENABLE_MOTOR : 0);
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			b2Vec2 raw = w->rawAxis[j];
			o.axis[0] = raw.x;
			o.axis[1] = raw.y;
			o.referenceAngle = j->m_referenceAngle;
			o.lowerAngle = j->m_lowerTranslation;
			o.upperAngle = j->m_upperTranslation;
			o.maxMotorTorque = j->m_maxMotorForce;
			o.motorSpeed = j->m_motorSpeed;
			o.impulse[0] = j->m_impulse.x;
			o.impulse[1] = j->m_impulse.y;
			o.impulse[2] = j->m_impulse.z;
			o.motorImpulse = j->m_motorImpulse;
			o.limitState = (int32_t)j->m_limitState;
			o.lastSolve[0] = j->m_axis.x;
			o.lastSolve[1] = j->m_axis.y;
			o.lastSolve[2] = j->m_perp.x;
			o.lastSolve[3] = j->m_perp.y;
		}
		else if (base->GetType() == e_wheelJoint)
		{
			const b2WheelJoint* j = (const b2WheelJoint*)base;
			o.type = B2CU_JOINT_WHEEL;
			o.flags |= j->m_enableMotor ? B2CU_JOINT_ENABLE_MOTOR : 0;
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.axis[0] = j->m_localXAxisA.x;
			o.axis[1] = j->m_localXAxisA.y;
			o.maxMotorTorque = j->m_maxMotorTorque;
			o.motorSpeed = j->m_motorSpeed;
			o.frequencyHz = j->m_frequencyHz;
			o.dampingRatio = j->m_dampingRatio;
			o.impulse[0] = j->m_impulse;
			o.impulse[1] = j->m_springImpulse;
			o.motorImpulse = j->m_motorImpulse;
			o.lastSolve[0] = j->m_ax.x;
			o.lastSolve[1] = j->m_ax.y;
			o.lastSolve[2] = j->m_ay.x;
			o.lastSolve[3] = j->m_ay.y;
			o.work[0] = j->m_sAx;
			o.work[1] = j->m_sBx;
		}
		else if (base->GetType() == e_ropeJoint)
		{
			const b2RopeJoint* j = (const b2RopeJoint*)base;
			o.type = B2CU_JOINT_ROPE;
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.length = j->m_maxLength;
			o.impulse[0] = j->m_impulse;
			o.limitState = (int32_t)j->m_state;
			o.lastSolve[0] = j->m_u.x;
			o.lastSolve[1] = j->m_u.y;
		}
		else if (base->GetType() == e_frictionJoint)
		{
			const b2FrictionJoint* j = (const b2FrictionJoint*)base;
			o.type = B2CU_JOINT_FRICTION;
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.length = j->m_maxForce;
			o.maxMotorTorque = j->m_maxTorque;
			o.impulse[0] = j->m_linearImpulse.x;
			o.impulse[1] = j->m_linearImpulse.y;
			o.impulse[2] = j->m_angularImpulse;
		}
		else if (base->GetType() == e_motorJoint)
		{
			const b2MotorJoint* j = (const b2MotorJoint*)base;
			o.type = B2CU_JOINT_MOTOR;
			o.axis[0] = j->m_linearOffset.x;
			o.axis[1] = j->m_linearOffset.y;
			o.referenceAngle = j->m_angularOffset;
			o.length = j->m_maxForce;
			o.maxMotorTorque = j->m_maxTorque;
			o.dampingRatio = j->m_correctionFactor;
			o.impulse[0] = j->m_linearImpulse.x;
			o.impulse[1] = j->m_linearImpulse.y;
			o.impulse[2] = j->m_angularImpulse;
		}
		else if (base->GetType() == e_pulleyJoint)
		{
			const b2PulleyJoint* j = (const b2PulleyJoint*)base;
			o.type = B2CU_JOINT_PULLEY;
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.axis[0] = j->m_groundAnchorA.x;
			o.axis[1] = j->m_groundAnchorA.y;
			o.lowerAngle = j->m_groundAnchorB.x;
			o.upperAngle = j->m_groundAnchorB.y;
			o.length = j->m_lengthA;
			o.referenceAngle = j->m_lengthB;
			o.motorSpeed = j->m_ratio;
			o.impulse[0] = j->m_impulse;
			o.lastSolve[0] = j->m_uB.x;
			o.lastSolve[1] = j->m_uB.y;
		}
		else if (base->GetType() == e_mouseJoint)
		{
			const b2MouseJoint* j = (const b2MouseJoint*)base;
			o.type = B2CU_JOINT_MOUSE;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.axis[0] = j->m_targetA.x;
			o.axis[1] = j->m_targetA.y;
			o.length = j->m_maxForce;
			o.frequencyHz = j->m_frequencyHz;
			o.dampingRatio = j->m_dampingRatio;
			o.maxMotorTorque = j->m_bodyB->GetMass();
			o.impulse[0] = j->m_impulse.x;
			o.impulse[1] = j->m_impulse.y;
		}
		else if (base->GetType() == e_gearJoint)
		{
			const b2GearJoint* j = (const b2GearJoint*)base;
			o.type = B2CU_JOINT_GEAR;
			o.flags |= (j->m_typeA == e_prismaticJoint ? B2CU_JOINT_GEAR_PRISMATIC_1 : 0) |
			           (j->m_typeB == e_prismaticJoint ? B2CU_JOINT_GEAR_PRISMATIC_2 : 0);
			for (size_t b = 0; b < w->bodies.size(); ++b)
			{
				if (w->bodies[b] == j->m_bodyC) o.limitState = (int32_t)b;
				if (w->bodies[b] == j->m_bodyD) o.reserved = (int32_t)b;
			}
			for (size_t k = 0; k < w->joints.size(); ++k)
			{
				if (w->joints[k] == j->m_joint1) o.frequencyHz = (float)k;
				if (w->joints[k] == j->m_joint2) o.dampingRatio = (float)k;
			}
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.axis[0] = j->m_localAnchorC.x;
			o.axis[1] = j->m_localAnchorC.y;
			o.lowerAngle = j->m_localAnchorD.x;
			o.upperAngle = j->m_localAnchorD.y;
			o.work[0] = j->m_localAxisC.x;
			o.work[1] = j->m_localAxisC.y;
			o.work[2] = j->m_localAxisD.x;
			o.work[3] = j->m_localAxisD.y;
			o.referenceAngle = j->m_referenceAngleA;
			o.maxMotorTorque = j->m_referenceAngleB;
			o.motorSpeed = j->m_ratio;
			o.length = j->m_constant;
			o.impulse[0] = j->m_impulse;
			o.lastSolve[0] = j->m_JvAC.x;
			o.lastSolve[1] = j->m_JvAC.y;
			o.lastSolve[2] = j->m_JwA;
		}
		else if (base->GetType() == e_distanceJoint)
		{
			const b2DistanceJoint* j = (const b2DistanceJoint*)base;
			o.type = B2CU_JOINT_DISTANCE;
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.length = j->m_length;
			o.frequencyHz = j->m_frequencyHz;
			o.dampingRatio = j->m_dampingRatio;
			o.impulse[0] = j->m_impulse;
			o.lastSolve[0] = j->m_u.x;
			o.lastSolve[1] = j->m_u.y;
		}
		else
		{
			const b2WeldJoint* j = (const b2WeldJoint*)base;
			o.type = B2CU_JOINT_WELD;
			o.localAnchorA[0] = j->m_localAnchorA.x;
			o.localAnchorA[1] = j->m_localAnchorA.y;
			o.localAnchorB[0] = j->m_localAnchorB.x;
			o.localAnchorB[1] = j->m_localAnchorB.y;
			o.referenceAngle = j->m_referenceAngle;
			o.frequencyHz = j->m_frequencyHz;
			o.dampingRatio = j->m_dampingRatio;
			o.impulse[0] = j->m_impulse.x;
			o.impulse[1] = j->m_impulse.y;
			o.impulse[2] = j->m_impulse.z;
		}
	}
}

void b2ref_profile(b2refWorld* w, float* out13)
{
	const b2Profile& p = w->world->GetProfile();
	memcpy(out13, &p, 13 * sizeof(float));
}

void b2ref_set_transform(b2refWorld* w, int32_t body, float x, float y, float angle)
{
	w->bodies[body]->SetTransform(b2Vec2(x, y), angle);
}

void b2ref_set_filter(b2refWorld* w, int32_t fixture, uint16_t categoryBits, uint16_t maskBits, int16_t groupIndex)
{
	b2Filter f;
	f.categoryBits = categoryBits;
	f.maskBits = maskBits;
	f.groupIndex = groupIndex;
	w->fixtures[fixture]->SetFilterData(f); // calls Refilter()
}

// Test filter (parity of b2cuSetPairFilter): rejects the pairs whose fixture indices sum to a multiple of `modulus`,
// everything else collides -- the category / mask / group rule is NOT applied, as with any user subclass.
namespace
{
class ModuloFilter : public b2ContactFilter
{
public:
	explicit ModuloFilter(int32 m) : modulus(m) {}
	bool ShouldCollide(b2Fixture* fixtureA, b2Fixture* fixtureB, uint32 threadId) override
	{
		B2_NOT_USED(threadId);
		int32 a = (int32)(intptr_t)fixtureA->GetUserData(), b = (int32)(intptr_t)fixtureB->GetUserData();
		return (a + b) % modulus != 0;
	}
	int32 modulus;
};
} // namespace

void b2ref_set_modulo_filter(b2refWorld* w, int32_t modulus)
{
	// owned by the process for the lifetime of the test (a handful of objects)
	w->world->SetContactFilter(modulus > 0 ? new ModuloFilter(modulus) : nullptr);
}

void b2ref_set_pre_solve_rule(b2refWorld* w, int32_t modulus) { w->listener.preSolveModulus = modulus; }
void b2ref_pre_solve_digest(b2refWorld* w, uint64_t* digest, int64_t* count)
{
	*digest = w->listener.preSolveDigest;
	*count = w->listener.preSolveCount;
}
void b2ref_record_post_solve(b2refWorld* w, int32_t on) { w->listener.recordPostSolve = on != 0; }
void b2ref_post_solve_digest(b2refWorld* w, uint64_t* digest, int64_t* count)
{
	*digest = w->listener.postSolveDigest;
	*count = w->listener.postSolveCount;
}

void b2ref_set_active(b2refWorld* w, int32_t body, int32_t on) { w->bodies[body]->SetActive(on != 0); }
void b2ref_set_type(b2refWorld* w, int32_t body, int32_t type) { w->bodies[body]->SetType((b2BodyType)type); }

void b2ref_set_velocity(b2refWorld* w, int32_t body, float vx, float vy, float angw)
{
	w->bodies[body]->SetLinearVelocity(b2Vec2(vx, vy));
	w->bodies[body]->SetAngularVelocity(angw);
}

void b2ref_apply_force(b2refWorld* w, int32_t body, float fx, float fy, float torque)
{
	w->bodies[body]->ApplyForceToCenter(b2Vec2(fx, fy), true);
	w->bodies[body]->ApplyTorque(torque, true);
}

void b2ref_set_awake(b2refWorld* w, int32_t body, int32_t awake)
{
	w->bodies[body]->SetAwake(awake != 0);
}

void b2ref_set_body_param(b2refWorld* w, int32_t body, int32_t which, float value)
{
	b2Body* b = w->bodies[body];
	switch (which)
	{
	case 0: b->SetLinearDamping(value); break;
	case 1: b->SetAngularDamping(value); break;
	case 2: b->SetGravityScale(value); break;
	case 3: b->SetBullet(value != 0.0f); break;
	case 4: b->SetSleepingAllowed(value != 0.0f); break;
	default: break;
	}
}

/* b2Body::DestroyFixture of the LAST fixture created (so that no proxy id of the harness changes) */
void b2ref_destroy_last_fixture(b2refWorld* w)
{
	b2Fixture* f = w->fixtures.back();
	f->GetBody()->DestroyFixture(f);
	w->fixtures.pop_back();
	w->proxyBase.pop_back();
	while (!w->proxies.empty() && w->proxies.back().first == f) w->proxies.pop_back();
}

uint32_t b2ref_hash(b2refWorld* w)
{
	uint32_t h = 2166136261u;
	for (const b2Body* b = w->world->GetBodyList(); b; b = b->GetNext())
	{
		float v[3] = {b->GetPosition().x, b->GetPosition().y, b->GetAngle()};
		const unsigned char* p = (const unsigned char*)v;
		for (size_t i = 0; i < sizeof(v); ++i)
		{
			h ^= p[i];
			h *= 16777619u;
		}
	}
	return h;
}

static void ImportShape(const b2cuShape* s, b2CircleShape& circle, b2EdgeShape& edge, b2PolygonShape& poly)
{
	switch (s->type)
	{
	case B2CU_SHAPE_CIRCLE:
		circle.m_radius = s->radius;
		circle.m_p.Set(s->v[0][0], s->v[0][1]);
		break;
	case B2CU_SHAPE_EDGE:
		edge.m_radius = s->radius;
		edge.m_vertex1.Set(s->v[0][0], s->v[0][1]);
		edge.m_vertex2.Set(s->v[1][0], s->v[1][1]);
		edge.m_vertex0.Set(s->v[2][0], s->v[2][1]);
		edge.m_vertex3.Set(s->v[3][0], s->v[3][1]);
		edge.m_hasVertex0 = (s->flags & B2CU_EDGE_HAS_VERTEX0) != 0;
		edge.m_hasVertex3 = (s->flags & B2CU_EDGE_HAS_VERTEX3) != 0;
		break;
	default:
		poly.m_radius = s->radius;
		poly.m_count = s->count;
		for (int32 i = 0; i < s->count; ++i)
		{
			poly.m_vertices[i].Set(s->v[i][0], s->v[i][1]);
			poly.m_normals[i].Set(s->n[i][0], s->n[i][1]);
		}
		poly.m_centroid.Set(s->centroid[0], s->centroid[1]);
		break;
	}
}

void b2ref_collide(const b2cuShape* shapeA, const float xfA[4], const b2cuShape* shapeB, const float xfB[4],
                   b2cuManifold* out)
{
	b2CircleShape circleA, circleB;
	b2EdgeShape edgeA, edgeB;
	b2PolygonShape polyA, polyB;
	ImportShape(shapeA, circleA, edgeA, polyA);
	ImportShape(shapeB, circleB, edgeB, polyB);
	b2Transform tA, tB;
	tA.p.Set(xfA[0], xfA[1]);
	tA.q.s = xfA[2];
	tA.q.c = xfA[3];
	tB.p.Set(xfB[0], xfB[1]);
	tB.q.s = xfB[2];
	tB.q.c = xfB[3];

	b2Manifold m;
	memset(&m, 0, sizeof(m));
	if (shapeA->type == B2CU_SHAPE_POLYGON && shapeB->type == B2CU_SHAPE_POLYGON)
	{
		b2CollidePolygons(&m, &polyA, tA, &polyB, tB);
	}
	else if (shapeA->type == B2CU_SHAPE_POLYGON && shapeB->type == B2CU_SHAPE_CIRCLE)
	{
		b2CollidePolygonAndCircle(&m, &polyA, tA, &circleB, tB);
	}
	else if (shapeA->type == B2CU_SHAPE_CIRCLE && shapeB->type == B2CU_SHAPE_CIRCLE)
	{
		b2CollideCircles(&m, &circleA, tA, &circleB, tB);
	}
	else if (shapeA->type == B2CU_SHAPE_EDGE && shapeB->type == B2CU_SHAPE_CIRCLE)
	{
		b2CollideEdgeAndCircle(&m, &edgeA, tA, &circleB, tB);
	}
	else if (shapeA->type == B2CU_SHAPE_EDGE && shapeB->type == B2CU_SHAPE_POLYGON)
	{
		b2CollideEdgeAndPolygon(&m, &edgeA, tA, &polyB, tB);
	}
	ExportManifold(m, out);
}

void b2ref_sincos(float x, float* s, float* c)
{
	b2Rot r(x);
	*s = r.s;
	*c = r.c;
}

namespace
{
struct CollectQuery : public b2QueryCallback
{
	std::vector<int32> ids;
	bool ReportFixture(b2Fixture* fixture) override
	{
		// the reference reports a fixture once per overlapping child proxy; record fixtures (first proxy id)
		ids.push_back(ProxyIndex(fixture, 0));
		return true;
	}
};
struct ClosestRay : public b2RayCastCallback
{
	b2Fixture* fixture = nullptr;
	b2Vec2 point, normal;
	float32 fraction = 1.0f;
	float32 ReportFixture(b2Fixture* f, const b2Vec2& p, const b2Vec2& n, float32 fr) override
	{
		fixture = f;
		point = p;
		normal = n;
		fraction = fr;
		return fr; // clip to the closest hit so far
	}
};
} // namespace

/* b2World::QueryAABB: first proxy id of every reported fixture (with repeats for multi-child fixtures), sorted */
int32_t b2ref_query_aabb(b2refWorld* w, const float aabb[4], int32_t capacity, int32_t* out)
{
	CollectQuery q;
	b2AABB box;
	box.lowerBound.Set(aabb[0], aabb[1]);
	box.upperBound.Set(aabb[2], aabb[3]);
	w->world->QueryAABB(&q, box);
	std::sort(q.ids.begin(), q.ids.end());
	for (int32_t i = 0; i < (int32_t)q.ids.size() && i < capacity; ++i) out[i] = q.ids[i];
	return (int32_t)q.ids.size();
}

/* b2World::RayCast with a closest-hit callback: returns the first proxy id of the hit fixture or -1; out = point.xy, normal.xy, fraction */
int32_t b2ref_ray_cast_closest(b2refWorld* w, const float p1[2], const float p2[2], float out[5])
{
	ClosestRay r;
	w->world->RayCast(&r, b2Vec2(p1[0], p1[1]), b2Vec2(p2[0], p2[1]));
	if (r.fixture == nullptr) return -1;
	out[0] = r.point.x;
	out[1] = r.point.y;
	out[2] = r.normal.x;
	out[3] = r.normal.y;
	out[4] = r.fraction;
	return ProxyIndex(r.fixture, 0);
}

void b2ref_distance(const b2cuShape* shapeA, const float xfA[4], const b2cuShape* shapeB, const float xfB[4],
                    int32_t useRadii, b2cuDistanceResult* out)
{
	b2CircleShape circleA, circleB;
	b2EdgeShape edgeA, edgeB;
	b2PolygonShape polyA, polyB;
	ImportShape(shapeA, circleA, edgeA, polyA);
	ImportShape(shapeB, circleB, edgeB, polyB);
	const b2Shape* sA = shapeA->type == B2CU_SHAPE_CIRCLE ? (const b2Shape*)&circleA
	                    : (shapeA->type == B2CU_SHAPE_EDGE ? (const b2Shape*)&edgeA : (const b2Shape*)&polyA);
	const b2Shape* sB = shapeB->type == B2CU_SHAPE_CIRCLE ? (const b2Shape*)&circleB
	                    : (shapeB->type == B2CU_SHAPE_EDGE ? (const b2Shape*)&edgeB : (const b2Shape*)&polyB);
	b2DistanceInput input;
	input.proxyA.Set(sA, 0);
	input.proxyB.Set(sB, 0);
	input.transformA.p.Set(xfA[0], xfA[1]);
	input.transformA.q.s = xfA[2];
	input.transformA.q.c = xfA[3];
	input.transformB.p.Set(xfB[0], xfB[1]);
	input.transformB.q.s = xfB[2];
	input.transformB.q.c = xfB[3];
	input.useRadii = useRadii != 0;
	b2SimplexCache cache;
	cache.count = 0;
	b2DistanceOutput output;
	b2Distance(&output, &cache, &input);
	out->distance = output.distance;
	out->pointA[0] = output.pointA.x;
	out->pointA[1] = output.pointA.y;
	out->pointB[0] = output.pointB.x;
	out->pointB[1] = output.pointB.y;
	out->iterations = output.iterations;
}

/* the reference's b2TimeOfImpact (Collision/b2TimeOfImpact.cpp:256-497) on geometry records and b2Sweep values */
static void ImportSweep(const b2cuSweep* r, b2Sweep& s)
{
	s.localCenter.Set(r->localCenter[0], r->localCenter[1]);
	s.c0.Set(r->c0[0], r->c0[1]);
	s.c.Set(r->c[0], r->c[1]);
	s.a0 = r->a0;
	s.a = r->a;
	s.alpha0 = r->alpha0;
}

void b2ref_time_of_impact(const b2cuShape* shapeA, const b2cuSweep* sweepA, const b2cuShape* shapeB,
                          const b2cuSweep* sweepB, float tMax, b2cuToiResult* out)
{
	b2CircleShape circleA, circleB;
	b2EdgeShape edgeA, edgeB;
	b2PolygonShape polyA, polyB;
	ImportShape(shapeA, circleA, edgeA, polyA);
	ImportShape(shapeB, circleB, edgeB, polyB);
	const b2Shape* sA = shapeA->type == B2CU_SHAPE_CIRCLE ? (const b2Shape*)&circleA
	                    : (shapeA->type == B2CU_SHAPE_EDGE ? (const b2Shape*)&edgeA : (const b2Shape*)&polyA);
	const b2Shape* sB = shapeB->type == B2CU_SHAPE_CIRCLE ? (const b2Shape*)&circleB
	                    : (shapeB->type == B2CU_SHAPE_EDGE ? (const b2Shape*)&edgeB : (const b2Shape*)&polyB);
	b2TOIInput input;
	input.proxyA.Set(sA, 0);
	input.proxyB.Set(sB, 0);
	ImportSweep(sweepA, input.sweepA);
	ImportSweep(sweepB, input.sweepB);
	input.tMax = tMax;
	b2TOIOutput output;
	b2TimeOfImpact(&output, &input);
	out->state = (int32_t)output.state;
	out->t = output.t;
}

} // extern "C"
