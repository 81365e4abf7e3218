// Flat C binding over the C++ host API (b2World / b2Body / b2Fixture / b2CudaStepExecutor) for Python tests and
// bench.py: scenes are built with the same calls user code makes (CreateBody, CreateFixture), stepped with
// b2World::Step(dt, vIters, pIters, b2CudaStepExecutor&), and read back through the public accessors.
#include "Box2D/Box2D.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{

struct BodyDefRec
{
	int32 type;
	float px, py, angle, vx, vy, w, linearDamping, angularDamping, gravityScale;
	uint32 flags; // 1 allowSleep, 2 awake, 4 fixedRotation, 8 bullet, 16 active
};

struct ShapeDefRec
{
	int32 kind; // 0 circle, 1 edge, 2 polygon (Set), 3 box (SetAsBox)
	int32 count;
	float radius;
	uint32 flags;
	float v[8][2];
	float n[8][2];
	float centroid[2];
};

struct FixtureDefRec
{
	int32 body, shape;
	float density, friction, restitution;
	uint32 flags;
	uint16 categoryBits, maskBits;
	int16 groupIndex;
	uint16 pad;
};

class Recorder : public b2ContactListener
{
public:
	bool BeginContactImmediate(b2Contact*, uint32) override { return true; }
	bool EndContactImmediate(b2Contact*, uint32) override { return true; }
	bool PreSolveImmediate(b2Contact*, const b2Manifold*, uint32) override { return preSolveModulus > 0; }
	void PreSolve(b2Contact* c, const b2Manifold* oldManifold) override
	{
		// the same test rule and digest as the oracle harness (ref_harness.cpp RecordingListener::PreSolve)
		uint64 key = c->GetKey();
		if (key % (uint64)preSolveModulus == 0) c->SetEnabled(false);
		preSolveDigest += key * 0x9E3779B97F4A7C15ull + (uint64)oldManifold->pointCount * 7u +
		                  (uint64)c->GetManifold()->pointCount;
		++preSolveCount;
	}
	int preSolveModulus = 0;
	uint64 preSolveDigest = 0;
	long long preSolveCount = 0;
	bool PostSolveImmediate(b2Contact*, const b2ContactImpulse*, uint32) override { return recordPostSolve; }
	void PostSolve(b2Contact* c, const b2ContactImpulse* impulse) override
	{
		// same digest as the oracle harness (oracle side: ref_harness.cpp RecordingListener::PostSolve)
		uint64 d = c->GetKey() * 0x9E3779B97F4A7C15ull + (uint64)impulse->count;
		for (int32 j = 0; j < impulse->count; ++j)
		{
			uint32 n, t;
			memcpy(&n, &impulse->normalImpulses[j], 4);
			memcpy(&t, &impulse->tangentImpulses[j], 4);
			d += ((uint64)n << 32 | t) * (uint64)(2 * j + 3);
		}
		postSolveDigest += d;
		++postSolveCount;
	}
	bool recordPostSolve = false;
	uint64 postSolveDigest = 0;
	long long postSolveCount = 0;
	void BeginContact(b2Contact* c) override
	{
		begins.push_back(c->GetKey());
		if (c->IsTouching()) ++beginTouching;
	}
	void EndContact(b2Contact* c) override { ends.push_back(c->GetKey()); }
	std::vector<uint64> begins, ends;
	int beginTouching = 0;
};

/// test filter: pairs whose fixture indices (creation order) sum to a multiple of `modulus` never collide
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

struct Host
{
	b2World* world;
	b2CudaStepExecutor* executor;
	Recorder recorder;
	std::vector<b2Body*> bodies;
	std::vector<b2Fixture*> fixtures;
	std::vector<b2Joint*> joints;
};

} // namespace

extern "C" {

#define B2H_API __attribute__((visibility("default")))

B2H_API void* b2h_create(float gx, float gy, uint32 worldFlags, int32 device, int32 downloadBodies, int32 events)
{
	Host* h = new Host;
	h->world = new b2World(b2Vec2(gx, gy));
	h->world->SetAllowSleeping((worldFlags & B2CU_WORLD_ALLOW_SLEEP) != 0);
	h->world->SetWarmStarting((worldFlags & B2CU_WORLD_WARM_STARTING) != 0);
	h->world->SetContinuousPhysics((worldFlags & B2CU_WORLD_CONTINUOUS) != 0);
	h->world->SetSubStepping((worldFlags & B2CU_WORLD_SUB_STEPPING) != 0);
	h->world->SetAutoClearForces((worldFlags & B2CU_WORLD_CLEAR_FORCES) != 0);
	if (events) h->world->SetContactListener(&h->recorder);
	b2CudaStepOptions opt;
	opt.device = device;
	opt.downloadBodies = downloadBodies != 0;
	opt.dispatchEvents = events != 0;
	h->executor = new b2CudaStepExecutor(opt);
	return h;
}

B2H_API void b2h_set_options(void* p, int32 downloadBodies, int32 events)
{
	Host* h = static_cast<Host*>(p);
	b2CudaStepOptions opt = h->executor->GetOptions();
	opt.downloadBodies = downloadBodies != 0;
	opt.dispatchEvents = events != 0;
	h->executor->SetOptions(opt);
	h->world->SetContactListener(events ? &h->recorder : nullptr);
}

/// apply a force to the centre of `count` consecutive bodies starting at `first` (the per-step user input of
/// the end-to-end benchmark)
B2H_API void b2h_apply_force_range(void* p, int32 first, int32 count, float fx, float fy)
{
	Host* h = static_cast<Host*>(p);
	for (int32 i = first; i < first + count && i < (int32)h->bodies.size(); ++i)
	{
		if (h->bodies[i]) h->bodies[i]->ApplyForceToCenter(b2Vec2(fx, fy), true);
	}
}

B2H_API int b2h_shard_configure(void* p, int32 rank, int32 rankCount, const int32* ghosts, int32 ghostCount,
                                const int32* exports, int32 exportCount, float gridFraction)
{
	Host* h = static_cast<Host*>(p);
	std::vector<b2Body*> g(ghostCount), e(exportCount);
	for (int32 i = 0; i < ghostCount; ++i) g[i] = h->bodies[ghosts[i]];
	for (int32 i = 0; i < exportCount; ++i) e[i] = h->bodies[exports[i]];
	return h->executor->ConfigureShard(*h->world, rank, rankCount, g.data(), ghostCount, e.data(), exportCount, gridFraction);
}
B2H_API int b2h_shard_link(void* p, b2cuShardLink* link)
{
	Host* h = static_cast<Host*>(p);
	return h->executor->GetShardLink(*h->world, link);
}
B2H_API int b2h_shard_connect(void* p, const b2cuShardLink* lower, const b2cuShardLink* upper)
{
	Host* h = static_cast<Host*>(p);
	return h->executor->ConnectShard(*h->world, lower, upper);
}

B2H_API void b2h_destroy(void* p)
{
	Host* h = static_cast<Host*>(p);
	if (!h) return;
	delete h->world;
	delete h->executor;
	delete h;
}

B2H_API int b2h_build(void* p, int32 bodyCount, const BodyDefRec* bodies, int32 shapeCount, const ShapeDefRec* shapes,
                      int32 fixtureCount, const FixtureDefRec* fixtures)
{
	Host* h = static_cast<Host*>(p);
	int32 f = 0;
	for (int32 i = 0; i < bodyCount; ++i)
	{
		const BodyDefRec& d = bodies[i];
		b2BodyDef bd;
		bd.type = (b2BodyType)d.type;
		bd.position.Set(d.px, d.py);
		bd.angle = d.angle;
		bd.linearVelocity.Set(d.vx, d.vy);
		bd.angularVelocity = d.w;
		bd.linearDamping = d.linearDamping;
		bd.angularDamping = d.angularDamping;
		bd.gravityScale = d.gravityScale;
		bd.allowSleep = (d.flags & 1) != 0;
		bd.awake = (d.flags & 2) != 0;
		bd.fixedRotation = (d.flags & 4) != 0;
		bd.bullet = (d.flags & 8) != 0;
		bd.active = (d.flags & 16) != 0;
		b2Body* body = h->world->CreateBody(&bd);
		h->bodies.push_back(body);
		while (f < fixtureCount && fixtures[f].body == i)
		{
			const FixtureDefRec& fd = fixtures[f];
			if (fd.shape < 0 || fd.shape >= shapeCount) return -1;
			const ShapeDefRec& sd = shapes[fd.shape];
			b2CircleShape circle;
			b2EdgeShape edge;
			b2PolygonShape poly;
			b2ChainShape chain;
			const b2Shape* shape = nullptr;
			switch (sd.kind)
			{
			case 5: // chain of v[0..count): closed loop if flags & 1
			{
				b2Vec2 vs[b2_maxPolygonVertices];
				for (int32 k = 0; k < sd.count; ++k) vs[k].Set(sd.v[k][0], sd.v[k][1]);
				if (sd.flags & 1) chain.CreateLoop(vs, sd.count);
				else chain.CreateChain(vs, sd.count);
				shape = &chain;
				break;
			}
			case 0:
				circle.m_radius = sd.radius;
				circle.m_p.Set(sd.v[0][0], sd.v[0][1]);
				shape = &circle;
				break;
			case 1:
				edge.Set(b2Vec2(sd.v[0][0], sd.v[0][1]), b2Vec2(sd.v[1][0], sd.v[1][1]));
				if (sd.flags & 1)
				{
					edge.m_vertex0.Set(sd.v[2][0], sd.v[2][1]);
					edge.m_hasVertex0 = true;
				}
				if (sd.flags & 2)
				{
					edge.m_vertex3.Set(sd.v[3][0], sd.v[3][1]);
					edge.m_hasVertex3 = true;
				}
				shape = &edge;
				break;
			case 2:
			{
				b2Vec2 vs[b2_maxPolygonVertices];
				for (int32 k = 0; k < sd.count; ++k) vs[k].Set(sd.v[k][0], sd.v[k][1]);
				poly.Set(vs, sd.count);
				shape = &poly;
				break;
			}
			default:
				if (sd.flags & 1) poly.SetAsBox(sd.v[0][0], sd.v[0][1], b2Vec2(sd.v[1][0], sd.v[1][1]), sd.v[2][0]);
				else poly.SetAsBox(sd.v[0][0], sd.v[0][1]);
				shape = &poly;
				break;
			}
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
			def.userData = (void*)(intptr_t)h->fixtures.size(); // fixture index in creation order
			h->fixtures.push_back(body->CreateFixture(&def));
			++f;
		}
	}
	return f == fixtureCount ? 0 : -2;
}

B2H_API int b2h_step(void* p, float dt, int32 velocityIterations, int32 positionIterations)
{
	Host* h = static_cast<Host*>(p);
	h->recorder.begins.clear();
	h->recorder.ends.clear();
	h->world->Step(dt, velocityIterations, positionIterations, *h->executor);
	return h->world->GetLastStepStatus();
}

B2H_API const char* b2h_last_error(void* p) { return static_cast<Host*>(p)->executor->GetLastError(); }

B2H_API void b2h_counts(void* p, int32* bodies, int32* fixtures, int32* contacts)
{
	Host* h = static_cast<Host*>(p);
	*bodies = h->world->GetBodyCount();
	*fixtures = h->world->GetProxyCount();
	*contacts = h->world->GetContactCount();
}

/// body state rows (dense body id order) as seen through the host mirror
B2H_API void b2h_get_bodies(void* p, b2cuBody* out)
{
	Host* h = static_cast<Host*>(p);
	memcpy(out, h->world->GetBodyStates(), sizeof(b2cuBody) * (size_t)h->world->GetBodyCount());
}

B2H_API void b2h_get_proxies(void* p, b2cuProxy* out)
{
	Host* h = static_cast<Host*>(p);
	memcpy(out, h->world->GetProxyStates(), sizeof(b2cuProxy) * (size_t)h->world->GetProxyCount());
}

/// positions / angles / awake flags through the per-body accessors (what TestMT.cpp:91-110 compares)
B2H_API void b2h_get_transforms(void* p, float* xya, int32* awake)
{
	Host* h = static_cast<Host*>(p);
	for (size_t i = 0; i < h->bodies.size(); ++i)
	{
		const b2Body* b = h->bodies[i];
		xya[3 * i + 0] = b->GetPosition().x;
		xya[3 * i + 1] = b->GetPosition().y;
		xya[3 * i + 2] = b->GetAngle();
		awake[i] = b->IsAwake() ? 1 : 0;
	}
}

/// sum of GetPosition().y over a body range, read through the per-body accessors (a cheap consumer of the mirror)
B2H_API double b2h_sum_y(void* p, int32 first, int32 count)
{
	Host* h = static_cast<Host*>(p);
	double sum = 0.0;
	for (int32 i = first; i < first + count && i < (int32)h->bodies.size(); ++i)
	{
		if (h->bodies[i]) sum += h->bodies[i]->GetPosition().y;
	}
	return sum;
}

/// mass data computed on the host by CreateFixture / ResetMassData
B2H_API void b2h_get_mass(void* p, float* massInertiaCenter)
{
	Host* h = static_cast<Host*>(p);
	for (size_t i = 0; i < h->bodies.size(); ++i)
	{
		b2MassData md;
		h->bodies[i]->GetMassData(&md);
		massInertiaCenter[4 * i + 0] = md.mass;
		massInertiaCenter[4 * i + 1] = md.I;
		massInertiaCenter[4 * i + 2] = md.center.x;
		massInertiaCenter[4 * i + 3] = md.center.y;
	}
}

/// contact snapshot through b2World::GetContactList(): keys, touching flags, point counts
B2H_API int b2h_get_contacts(void* p, int32 capacity, uint64* keys, int32* touching, int32* pointCount)
{
	Host* h = static_cast<Host*>(p);
	int32 n = 0;
	for (b2Contact* c = h->world->GetContactList(); c; c = c->GetNext())
	{
		if (n < capacity)
		{
			keys[n] = c->GetKey();
			touching[n] = c->IsTouching() ? 1 : 0;
			pointCount[n] = c->GetManifold()->pointCount;
		}
		++n;
	}
	return n;
}

B2H_API int b2h_events(void* p, int32 kind, int32 capacity, uint64* keys)
{
	Host* h = static_cast<Host*>(p);
	const std::vector<uint64>& v = kind == 0 ? h->recorder.begins : h->recorder.ends;
	for (size_t i = 0; i < v.size() && (int32)i < capacity; ++i) keys[i] = v[i];
	return (int)v.size();
}

/// solver order of the last step (keys), straight from the device handle
B2H_API int b2h_solver_order(void* p, int32 capacity, uint64* keys)
{
	Host* h = static_cast<Host*>(p);
	b2cuWorld* dev = h->executor->GetDeviceWorld(h->world);
	if (!dev) return 0;
	int32 n = 0;
	b2cuGetSolverOrder(dev, capacity, keys, nullptr, &n);
	return n;
}

// ---- joints: b2World::CreateJoint / DestroyJoint and the b2RevoluteJoint accessors, from / to b2cuJoint rows ----
B2H_API int b2h_create_joints(void* p, int32 count, const b2cuJoint* rows)
{
	Host* h = static_cast<Host*>(p);
	for (int32 i = 0; i < count; ++i)
	{
		const b2cuJoint& r = rows[i];
		b2Joint* j = nullptr;
		const bool collideConnected = (r.flags & B2CU_JOINT_COLLIDE_CONNECTED) != 0;
		if (r.type == B2CU_JOINT_REVOLUTE)
		{
			b2RevoluteJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.referenceAngle = r.referenceAngle;
			def.enableLimit = (r.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;
			def.lowerAngle = r.lowerAngle;
			def.upperAngle = r.upperAngle;
			def.enableMotor = (r.flags & B2CU_JOINT_ENABLE_MOTOR) != 0;
			def.motorSpeed = r.motorSpeed;
			def.maxMotorTorque = r.maxMotorTorque;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_PRISMATIC)
		{
			b2PrismaticJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.localAxisA.Set(r.axis[0], r.axis[1]);
			def.referenceAngle = r.referenceAngle;
			def.enableLimit = (r.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;
			def.lowerTranslation = r.lowerAngle;
			def.upperTranslation = r.upperAngle;
			def.enableMotor = (r.flags & B2CU_JOINT_ENABLE_MOTOR) != 0;
			def.motorSpeed = r.motorSpeed;
			def.maxMotorForce = r.maxMotorTorque;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_WHEEL)
		{
			b2WheelJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.localAxisA.Set(r.axis[0], r.axis[1]);
			def.enableMotor = (r.flags & B2CU_JOINT_ENABLE_MOTOR) != 0;
			def.maxMotorTorque = r.maxMotorTorque;
			def.motorSpeed = r.motorSpeed;
			def.frequencyHz = r.frequencyHz;
			def.dampingRatio = r.dampingRatio;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_ROPE)
		{
			b2RopeJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.maxLength = r.length;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_FRICTION)
		{
			b2FrictionJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.maxForce = r.length;
			def.maxTorque = r.maxMotorTorque;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_MOTOR)
		{
			b2MotorJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.linearOffset.Set(r.axis[0], r.axis[1]);
			def.angularOffset = r.referenceAngle;
			def.maxForce = r.length;
			def.maxTorque = r.maxMotorTorque;
			def.correctionFactor = r.dampingRatio;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_PULLEY)
		{
			b2PulleyJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.groundAnchorA.Set(r.axis[0], r.axis[1]);
			def.groundAnchorB.Set(r.lowerAngle, r.upperAngle);
			def.lengthA = r.length;
			def.lengthB = r.referenceAngle;
			def.ratio = r.motorSpeed;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_MOUSE)
		{
			b2MouseJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.target.Set(r.axis[0], r.axis[1]);
			def.maxForce = r.length;
			def.frequencyHz = r.frequencyHz;
			def.dampingRatio = r.dampingRatio;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_GEAR)
		{
			int32 i1 = (int32)r.frequencyHz, i2 = (int32)r.dampingRatio; // rows of the joints it couples
			if (i1 < 0 || i2 < 0 || i1 >= (int32)h->joints.size() || i2 >= (int32)h->joints.size()) return -3;
			b2GearJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.joint1 = h->joints[i1];
			def.joint2 = h->joints[i2];
			def.ratio = r.motorSpeed;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_DISTANCE)
		{
			b2DistanceJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.length = r.length;
			def.frequencyHz = r.frequencyHz;
			def.dampingRatio = r.dampingRatio;
			j = h->world->CreateJoint(&def);
		}
		else if (r.type == B2CU_JOINT_WELD)
		{
			b2WeldJointDef def;
			def.bodyA = h->bodies[r.bodyA];
			def.bodyB = h->bodies[r.bodyB];
			def.collideConnected = collideConnected;
			def.localAnchorA.Set(r.localAnchorA[0], r.localAnchorA[1]);
			def.localAnchorB.Set(r.localAnchorB[0], r.localAnchorB[1]);
			def.referenceAngle = r.referenceAngle;
			def.frequencyHz = r.frequencyHz;
			def.dampingRatio = r.dampingRatio;
			j = h->world->CreateJoint(&def);
		}
		else
		{
			return -1;
		}
		if (j == nullptr) return -2;
		h->joints.push_back(j);
	}
	return 0;
}

B2H_API void b2h_destroy_joint(void* p, int32 joint)
{
	Host* h = static_cast<Host*>(p);
	h->world->DestroyJoint(h->joints[joint]);
	h->joints.erase(h->joints.begin() + joint);
}

/// per joint: reaction force x, y, reaction torque, motor torque (all at inv_dt), joint angle, joint speed
B2H_API void b2h_joint_readings(void* p, float inv_dt, float* out6)
{
	Host* h = static_cast<Host*>(p);
	for (size_t i = 0; i < h->joints.size(); ++i)
	{
		const b2Joint* base = h->joints[i];
		b2Vec2 f = base->GetReactionForce(inv_dt);
		float* o = out6 + 6 * i;
		o[0] = f.x;
		o[1] = f.y;
		o[2] = base->GetReactionTorque(inv_dt);
		o[3] = o[4] = o[5] = 0.0f;
		if (base->GetType() == e_revoluteJoint)
		{
			const b2RevoluteJoint* j = static_cast<const b2RevoluteJoint*>(base);
			o[3] = j->GetMotorTorque(inv_dt);
			o[4] = j->GetJointAngle();
			o[5] = j->GetJointSpeed();
		}
		else if (base->GetType() == e_prismaticJoint)
		{
			const b2PrismaticJoint* j = static_cast<const b2PrismaticJoint*>(base);
			o[3] = j->GetMotorForce(inv_dt);
			o[4] = j->GetJointTranslation();
			o[5] = j->GetJointSpeed();
		}
		else if (base->GetType() == e_wheelJoint)
		{
			const b2WheelJoint* j = static_cast<const b2WheelJoint*>(base);
			o[3] = j->GetMotorTorque(inv_dt);
			o[4] = j->GetJointTranslation();
			o[5] = j->GetJointLinearSpeed();
		}
	}
}

B2H_API void b2h_joint_set_motor(void* p, int32 joint, int32 enable, float speed, float maxTorque)
{
	if (static_cast<Host*>(p)->joints[joint]->GetType() == e_prismaticJoint)
	{
		b2PrismaticJoint* pj = static_cast<b2PrismaticJoint*>(static_cast<Host*>(p)->joints[joint]);
		pj->EnableMotor(enable != 0);
		pj->SetMotorSpeed(speed);
		pj->SetMaxMotorForce(maxTorque);
		return;
	}
	b2RevoluteJoint* j = static_cast<b2RevoluteJoint*>(static_cast<Host*>(p)->joints[joint]);
	j->EnableMotor(enable != 0);
	j->SetMotorSpeed(speed);
	j->SetMaxMotorTorque(maxTorque);
}

B2H_API void b2h_joint_set_limits(void* p, int32 joint, int32 enable, float lower, float upper)
{
	if (static_cast<Host*>(p)->joints[joint]->GetType() == e_prismaticJoint)
	{
		b2PrismaticJoint* pj = static_cast<b2PrismaticJoint*>(static_cast<Host*>(p)->joints[joint]);
		pj->EnableLimit(enable != 0);
		pj->SetLimits(lower, upper);
		return;
	}
	b2RevoluteJoint* j = static_cast<b2RevoluteJoint*>(static_cast<Host*>(p)->joints[joint]);
	j->EnableLimit(enable != 0);
	j->SetLimits(lower, upper);
}

/// b2DistanceJoint::SetLength / SetFrequency / SetDampingRatio, b2WeldJoint::SetFrequency / SetDampingRatio
B2H_API void b2h_joint_set_spring(void* p, int32 joint, float length, float frequencyHz, float dampingRatio)
{
	b2Joint* base = static_cast<Host*>(p)->joints[joint];
	if (base->GetType() == e_distanceJoint)
	{
		b2DistanceJoint* j = static_cast<b2DistanceJoint*>(base);
		j->SetLength(length);
		j->SetFrequency(frequencyHz);
		j->SetDampingRatio(dampingRatio);
	}
	else if (base->GetType() == e_weldJoint)
	{
		b2WeldJoint* j = static_cast<b2WeldJoint*>(base);
		j->SetFrequency(frequencyHz);
		j->SetDampingRatio(dampingRatio);
	}
}

/// b2MouseJoint::SetTarget / b2MotorJoint::SetLinearOffset
B2H_API void b2h_joint_set_target(void* p, int32 joint, float x, float y)
{
	b2Joint* base = static_cast<Host*>(p)->joints[joint];
	if (base->GetType() == e_mouseJoint) static_cast<b2MouseJoint*>(base)->SetTarget(b2Vec2(x, y));
	else if (base->GetType() == e_motorJoint) static_cast<b2MotorJoint*>(base)->SetLinearOffset(b2Vec2(x, y));
}

B2H_API int b2h_joint_count(void* p) { return static_cast<Host*>(p)->world->GetJointCount(); }

/// the order in which the device solves the joints (valid after a step)
B2H_API int b2h_joint_order(void* p, int32 capacity, int32* ids)
{
	Host* h = static_cast<Host*>(p);
	b2cuWorld* dev = h->executor->GetDeviceWorld(h->world);
	if (!dev) return 0;
	int32 n = 0;
	b2cuGetJointOrder(dev, capacity, ids, &n);
	return n;
}

B2H_API void b2h_profile(void* p, float* out13) { memcpy(out13, &static_cast<Host*>(p)->world->GetProfile(), 13 * sizeof(float)); }

/// the C-ABI handle behind the world (nullptr before the first step), for the diagnostic entry points of b2cuda.h
B2H_API void* b2h_device_handle(void* p)
{
	Host* h = static_cast<Host*>(p);
	return h->executor->GetDeviceWorld(h->world);
}

B2H_API void b2h_step_info(void* p, b2cuStepInfo* out) { *out = static_cast<Host*>(p)->executor->GetLastStepInfo(); }

B2H_API void b2h_host_timings(void* p, float* out4) { memcpy(out4, static_cast<Host*>(p)->executor->GetLastHostTimings(), 4 * sizeof(float)); }

/// FNV-1a over (x, y, angle) of every body in GetBodyList() order: the trajectory hash of SURVEY.md 8c
B2H_API uint32 b2h_hash(void* p)
{
	Host* h = static_cast<Host*>(p);
	uint32 hash = 2166136261u;
	for (const b2Body* b = h->world->GetBodyList(); b; b = b->GetNext())
	{
		float v[3] = {b->GetPosition().x, b->GetPosition().y, b->GetAngle()};
		const unsigned char* bytes = reinterpret_cast<const unsigned char*>(v);
		for (size_t i = 0; i < sizeof(v); ++i)
		{
			hash ^= bytes[i];
			hash *= 16777619u;
		}
	}
	return hash;
}

B2H_API void b2h_set_transform(void* p, int32 body, float x, float y, float angle)
{
	static_cast<Host*>(p)->bodies[body]->SetTransform(b2Vec2(x, y), angle);
}
B2H_API void b2h_set_modulo_filter(void* p, int32 modulus)
{
	Host* h = static_cast<Host*>(p);
	h->world->SetContactFilter(modulus > 0 ? new ModuloFilter(modulus) : nullptr);
}
B2H_API void b2h_set_pre_solve_rule(void* p, int32 modulus)
{
	Host* h = static_cast<Host*>(p);
	h->recorder.preSolveModulus = modulus;
	b2CudaStepOptions opt = h->executor->GetOptions();
	opt.reportPreSolve = modulus > 0;
	h->executor->SetOptions(opt);
}
B2H_API void b2h_pre_solve_digest(void* p, uint64* digest, long long* count)
{
	Host* h = static_cast<Host*>(p);
	*digest = h->recorder.preSolveDigest;
	*count = h->recorder.preSolveCount;
}
B2H_API void b2h_record_post_solve(void* p, int32 on)
{
	Host* h = static_cast<Host*>(p);
	h->recorder.recordPostSolve = on != 0;
	b2CudaStepOptions opt = h->executor->GetOptions();
	opt.reportPostSolve = on != 0;
	h->executor->SetOptions(opt);
}
B2H_API void b2h_post_solve_digest(void* p, uint64* digest, long long* count)
{
	Host* h = static_cast<Host*>(p);
	*digest = h->recorder.postSolveDigest;
	*count = h->recorder.postSolveCount;
}
namespace
{
struct CollectQuery : public b2QueryCallback
{
	std::vector<int32> ids;
	bool ReportFixture(b2Fixture* fixture) override
	{
		ids.push_back(fixture->GetProxyIndex());
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
		return fr;
	}
};
} // namespace

B2H_API int32 b2h_query_aabb(void* p, const float* aabb, int32 capacity, int32* out)
{
	Host* h = static_cast<Host*>(p);
	CollectQuery q;
	b2AABB box;
	box.lowerBound.Set(aabb[0], aabb[1]);
	box.upperBound.Set(aabb[2], aabb[3]);
	h->world->QueryAABB(&q, box);
	std::sort(q.ids.begin(), q.ids.end());
	for (int32 i = 0; i < (int32)q.ids.size() && i < capacity; ++i) out[i] = q.ids[i];
	return (int32)q.ids.size();
}
B2H_API int32 b2h_ray_cast_closest(void* p, const float* p1, const float* p2, float* out)
{
	Host* h = static_cast<Host*>(p);
	ClosestRay r;
	h->world->RayCast(&r, b2Vec2(p1[0], p1[1]), b2Vec2(p2[0], p2[1]));
	if (r.fixture == nullptr) return -1;
	out[0] = r.point.x;
	out[1] = r.point.y;
	out[2] = r.normal.x;
	out[3] = r.normal.y;
	out[4] = r.fraction;
	return r.fixture->GetProxyIndex();
}
B2H_API void b2h_shift_origin(void* p, float x, float y) { static_cast<Host*>(p)->world->ShiftOrigin(b2Vec2(x, y)); }
B2H_API void b2h_set_active(void* p, int32 body, int32 on) { static_cast<Host*>(p)->bodies[body]->SetActive(on != 0); }
B2H_API void b2h_set_type(void* p, int32 body, int32 type) { static_cast<Host*>(p)->bodies[body]->SetType((b2BodyType)type); }
B2H_API void b2h_set_filter(void* p, int32 fixture, uint16 categoryBits, uint16 maskBits, int16 groupIndex)
{
	b2Filter f;
	f.categoryBits = categoryBits;
	f.maskBits = maskBits;
	f.groupIndex = groupIndex;
	static_cast<Host*>(p)->fixtures[fixture]->SetFilterData(f);
}
B2H_API void b2h_set_velocity(void* p, int32 body, float vx, float vy, float w)
{
	Host* h = static_cast<Host*>(p);
	h->bodies[body]->SetLinearVelocity(b2Vec2(vx, vy));
	h->bodies[body]->SetAngularVelocity(w);
}
B2H_API void b2h_apply_force(void* p, int32 body, float fx, float fy, float torque)
{
	Host* h = static_cast<Host*>(p);
	h->bodies[body]->ApplyForceToCenter(b2Vec2(fx, fy), true);
	h->bodies[body]->ApplyTorque(torque, true);
}
B2H_API void b2h_set_awake(void* p, int32 body, int32 awake) { static_cast<Host*>(p)->bodies[body]->SetAwake(awake != 0); }
/// b2Body::SetLinearDamping (0) / SetAngularDamping (1) / SetGravityScale (2) / SetBullet (3) / SetSleepingAllowed (4)
B2H_API void b2h_set_body_param(void* p, int32 body, int32 which, float value)
{
	b2Body* b = static_cast<Host*>(p)->bodies[body];
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
/// b2Body::DestroyFixture of fixture `fixture` (creation order)
B2H_API void b2h_destroy_fixture(void* p, int32 fixture)
{
	Host* h = static_cast<Host*>(p);
	b2Fixture* f = h->fixtures[fixture];
	if (f == nullptr) return;
	f->GetBody()->DestroyFixture(f);
	h->fixtures[fixture] = nullptr;
}
B2H_API void b2h_destroy_body(void* p, int32 body)
{
	Host* h = static_cast<Host*>(p);
	if (h->bodies[body] == nullptr) return;
	h->world->DestroyBody(h->bodies[body]);
	h->bodies[body] = nullptr;
	// DestroyBody took the body's joints along: the handle table follows the world's joint list (rows = GetIndex())
	h->joints.assign((size_t)h->world->GetJointCount(), nullptr);
	for (b2Joint* j = h->world->GetJointList(); j; j = j->GetNext()) h->joints[(size_t)j->GetIndex()] = j;
}

// ---- b2CudaShardedWorld (Box2D/MT/b2CudaShardedWorld.h) ----

/// planner only, no device: the strip `rank` of `shardCount` of this host's scene.  bodyIds[capacity] receives the scene
/// body index of every strip body, ghostLocal / exportLocal the strip-local indices of the halo lists, counts =
/// {bodies, ghosts, exports, fixtures}, bounds[shardCount + 1] the strip boundaries.  Returns 0, or -1 if capacity is short.
B2H_API int b2h_plan_strip(void* p, int32 shardCount, int32 rank, float margin, int32 capacity, int32* bodyIds,
                           int32* ghostLocal, int32* exportLocal, int32* counts, double* bounds)
{
	Host* h = static_cast<Host*>(p);
	std::vector<float64> b;
	b2CudaShardedWorld::ComputeBounds(*h->world, shardCount, b);
	for (size_t i = 0; i < b.size(); ++i) bounds[i] = b[i];
	b2World* strip = b2CudaShardedWorld::MakeStripWorld(*h->world);
	b2ShardStrip info;
	b2CudaShardedWorld::BuildStrip(*h->world, b, rank, margin, *strip, info);
	counts[0] = (int32)info.bodies.size();
	counts[1] = (int32)info.ghosts.size();
	counts[2] = (int32)info.exports.size();
	counts[3] = strip->GetProxyCount();
	int rc = 0;
	if (counts[0] > capacity) rc = -1;
	else
	{
		for (size_t i = 0; i < info.globalIds.size(); ++i) bodyIds[i] = info.globalIds[i];
		for (size_t i = 0; i < info.ghosts.size(); ++i) ghostLocal[i] = info.ghosts[i]->GetIndex();
		for (size_t i = 0; i < info.exports.size(); ++i) exportLocal[i] = info.exports[i]->GetIndex();
	}
	delete strip;
	return rc;
}

B2H_API void* b2h_sharded_create(void* p, int32 shardCount, float margin, const int32* devices, float gridFraction)
{
	Host* h = static_cast<Host*>(p);
	return new b2CudaShardedWorld(*h->world, shardCount, margin, devices, gridFraction);
}
B2H_API int b2h_sharded_status(void* s) { return static_cast<b2CudaShardedWorld*>(s)->GetLastStatus(); }
B2H_API const char* b2h_sharded_error(void* s) { return static_cast<b2CudaShardedWorld*>(s)->GetLastError(); }
B2H_API int b2h_sharded_step(void* s, float dt, int32 velocityIterations, int32 positionIterations)
{
	b2CudaShardedWorld* w = static_cast<b2CudaShardedWorld*>(s);
	w->Step(dt, velocityIterations, positionIterations);
	return w->GetLastStatus();
}
/// copies the stepped state back into the scene the sharded world was made from
B2H_API void b2h_sharded_gather(void* s, void* p) { static_cast<b2CudaShardedWorld*>(s)->Gather(*static_cast<Host*>(p)->world); }
/// (x, y, angle) of the bodies of strip `rank` in strip order; returns their number
B2H_API int b2h_sharded_strip_transforms(void* s, int32 rank, int32 capacity, float* xya)
{
	b2CudaShardedWorld* w = static_cast<b2CudaShardedWorld*>(s);
	const b2ShardStrip& info = w->GetStripInfo(rank);
	int32 n = (int32)info.bodies.size();
	for (int32 i = 0; i < n && i < capacity; ++i)
	{
		const b2Body* b = info.bodies[(size_t)i];
		xya[3 * i + 0] = b->GetPosition().x;
		xya[3 * i + 1] = b->GetPosition().y;
		xya[3 * i + 2] = b->GetAngle();
	}
	return n;
}
/// Rebalance(): bounds = NULL (equal population at the current positions) or shardCount + 1 values.  Returns the status.
B2H_API int b2h_sharded_rebalance(void* s, const double* bounds)
{
	b2CudaShardedWorld* w = static_cast<b2CudaShardedWorld*>(s);
	w->Rebalance(bounds);
	return w->GetLastStatus();
}
B2H_API void b2h_sharded_set_transport(void* s, int32 downloadBodies, int32 events)
{
	static_cast<b2CudaShardedWorld*>(s)->SetTransport(downloadBodies != 0, events != 0);
}
B2H_API void b2h_sharded_set_rebalance_interval(void* s, int32 steps) { static_cast<b2CudaShardedWorld*>(s)->SetRebalanceInterval(steps); }
B2H_API int b2h_sharded_lost_contacts(void* s) { return static_cast<b2CudaShardedWorld*>(s)->GetLostContacts(); }
B2H_API void b2h_sharded_bounds(void* s, double* out)
{
	const std::vector<float64>& b = static_cast<b2CudaShardedWorld*>(s)->GetBounds();
	for (size_t i = 0; i < b.size(); ++i) out[i] = b[i];
}
/// counts = {bodies, ghosts, exports, proxies}; ids / ghostLocal / exportLocal may be NULL (counts only)
B2H_API void b2h_sharded_strip_plan(void* s, int32 rank, int32* counts, int32* ids, int32* ghostLocal, int32* exportLocal)
{
	const b2ShardStrip& info = static_cast<b2CudaShardedWorld*>(s)->GetStripInfo(rank);
	counts[0] = (int32)info.bodies.size();
	counts[1] = (int32)info.ghosts.size();
	counts[2] = (int32)info.exports.size();
	counts[3] = (int32)info.proxyGlobal.size();
	if (ids) for (size_t i = 0; i < info.globalIds.size(); ++i) ids[i] = info.globalIds[i];
	if (ghostLocal) for (size_t i = 0; i < info.ghosts.size(); ++i) ghostLocal[i] = info.ghosts[i]->GetIndex();
	if (exportLocal) for (size_t i = 0; i < info.exports.size(); ++i) exportLocal[i] = info.exports[i]->GetIndex();
}
/// whole body records of a strip, in strip order
B2H_API void b2h_sharded_strip_bodies(void* s, int32 rank, b2cuBody* out)
{
	b2World& w = static_cast<b2CudaShardedWorld*>(s)->GetStrip(rank);
	memcpy(out, w.GetBodyStates(), sizeof(b2cuBody) * (size_t)w.GetBodyCount());
}
namespace
{
uint64 GlobalKey(const b2ShardStrip& info, uint64 key)
{
	uint64 a = (uint64)info.proxyGlobal[(size_t)(key >> 32)], b = (uint64)info.proxyGlobal[(size_t)(key & 0xFFFFFFFFull)];
	return (std::min(a, b) << 32) | std::max(a, b);
}
}
/// the strip's constraints of the last step in solve order, keys over SCENE proxy ids, with their colours
B2H_API int b2h_sharded_solver_order(void* s, int32 rank, int32 capacity, uint64* keys, int32* colour)
{
	b2CudaShardedWorld* w = static_cast<b2CudaShardedWorld*>(s);
	b2cuWorld* device = w->GetExecutor(rank).GetDeviceWorld(&w->GetStrip(rank));
	int32 n = 0;
	if (device == nullptr || b2cuGetSolverOrder(device, capacity, keys, colour, &n) != B2CU_OK) return -1;
	const b2ShardStrip& info = w->GetStripInfo(rank);
	for (int32 i = 0; i < n && i < capacity; ++i) keys[i] = GlobalKey(info, keys[i]);
	return n;
}
/// the strip's contact set as keys over SCENE proxy ids (unsorted)
B2H_API int b2h_sharded_contact_keys(void* s, int32 rank, int32 capacity, uint64* keys)
{
	b2CudaShardedWorld* w = static_cast<b2CudaShardedWorld*>(s);
	b2cuWorld* device = w->GetExecutor(rank).GetDeviceWorld(&w->GetStrip(rank));
	int32 n = 0;
	if (device == nullptr || b2cuGetContactCount(device, &n) != B2CU_OK) return -1;
	if (n > capacity) return n;
	std::vector<b2cuContact> recs((size_t)n);
	if (n > 0 && b2cuGetContacts(device, n, recs.data(), &n) != B2CU_OK) return -1;
	const b2ShardStrip& info = w->GetStripInfo(rank);
	for (int32 i = 0; i < n; ++i)
		keys[i] = GlobalKey(info, ((uint64)(uint32)std::min(recs[(size_t)i].proxyA, recs[(size_t)i].proxyB) << 32) |
		                              (uint64)(uint32)std::max(recs[(size_t)i].proxyA, recs[(size_t)i].proxyB));
	return n;
}
B2H_API void b2h_sharded_destroy(void* s) { delete static_cast<b2CudaShardedWorld*>(s); }

} // extern "C"
