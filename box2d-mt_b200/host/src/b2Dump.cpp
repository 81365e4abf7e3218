// b2World::Dump (reference: b2World.cpp:2107-2164 and the Dump methods of b2Body, b2Fixture and the joints): writes, through
// b2Log, C++ code that rebuilds the world with the public API (`m_world`, `bodies[]`, `joints[]` as in the reference's
// output, so a dump pastes into a Testbed test).  Everything is printed from the flat records the world keeps (body
// rows, joint rows) through small field tables instead of per-class methods.
#include "Box2D/Box2D.h"

#include <cstdio>
#include <cstring>

namespace
{

// a float literal that round-trips the fp32 value exactly (9 significant digits) and is valid C++ ("12" -> "12.0f")
struct Literal
{
	char text[40];
	explicit Literal(float32 v)
	{
		snprintf(text, sizeof text - 4, "%.9g", (double)v);
		bool plain = true;
		for (const char* c = text; *c; ++c)
			if (*c == '.' || *c == 'e' || *c == 'n' || *c == 'i') plain = false;
		size_t n = strlen(text);
		if (plain)
		{
			text[n++] = '.';
			text[n++] = '0';
		}
		text[n++] = 'f';
		text[n] = 0;
	}
};
void Float(const char* indent, const char* lhs, float32 v) { b2Log("%s%s = %s;\n", indent, lhs, Literal(v).text); }
void Vec(const char* indent, const char* lhs, float32 x, float32 y)
{
	b2Log("%s%s.Set(%s, %s);\n", indent, lhs, Literal(x).text, Literal(y).text);
}
void Flag(const char* indent, const char* lhs, bool v) { b2Log("%s%s = %s;\n", indent, lhs, v ? "true" : "false"); }

void DumpShape(const b2Shape* shape)
{
	const char* in = "    ";
	switch (shape->GetType())
	{
	case b2Shape::e_circle:
	{
		const b2CircleShape* s = static_cast<const b2CircleShape*>(shape);
		b2Log("%sb2CircleShape shape;\n", in);
		Float(in, "shape.m_radius", s->m_radius);
		Vec(in, "shape.m_p", s->m_p.x, s->m_p.y);
		break;
	}
	case b2Shape::e_edge:
	{
		const b2EdgeShape* s = static_cast<const b2EdgeShape*>(shape);
		b2Log("%sb2EdgeShape shape;\n", in);
		Float(in, "shape.m_radius", s->m_radius);
		Vec(in, "shape.m_vertex0", s->m_vertex0.x, s->m_vertex0.y);
		Vec(in, "shape.m_vertex1", s->m_vertex1.x, s->m_vertex1.y);
		Vec(in, "shape.m_vertex2", s->m_vertex2.x, s->m_vertex2.y);
		Vec(in, "shape.m_vertex3", s->m_vertex3.x, s->m_vertex3.y);
		Flag(in, "shape.m_hasVertex0", s->m_hasVertex0);
		Flag(in, "shape.m_hasVertex3", s->m_hasVertex3);
		break;
	}
	case b2Shape::e_polygon:
	{
		const b2PolygonShape* s = static_cast<const b2PolygonShape*>(shape);
		b2Log("%sb2PolygonShape shape;\n%sb2Vec2 vs[%d];\n", in, in, (int)b2_maxPolygonVertices);
		for (int32 i = 0; i < s->GetVertexCount(); ++i)
		{
			char lhs[32];
			snprintf(lhs, sizeof lhs, "vs[%d]", (int)i);
			Vec(in, lhs, s->GetVertex(i).x, s->GetVertex(i).y);
		}
		b2Log("%sshape.Set(vs, %d);\n", in, (int)s->GetVertexCount());
		break;
	}
	case b2Shape::e_chain:
	{
		const b2ChainShape* s = static_cast<const b2ChainShape*>(shape);
		b2Log("%sb2ChainShape shape;\n%sb2Vec2 vs[%d];\n", in, in, (int)s->GetVertexCount());
		for (int32 i = 0; i < s->GetVertexCount(); ++i)
		{
			char lhs[32];
			snprintf(lhs, sizeof lhs, "vs[%d]", (int)i);
			Vec(in, lhs, s->GetVertex(i).x, s->GetVertex(i).y);
		}
		b2Log("%sshape.CreateChain(vs, %d);\n", in, (int)s->GetVertexCount());
		Vec(in, "shape.m_prevVertex", s->m_prevVertex.x, s->m_prevVertex.y);
		Vec(in, "shape.m_nextVertex", s->m_nextVertex.x, s->m_nextVertex.y);
		Flag(in, "shape.m_hasPrevVertex", s->m_hasPrevVertex);
		Flag(in, "shape.m_hasNextVertex", s->m_hasNextVertex);
		break;
	}
	default:
		break;
	}
}

// how the fields of a b2cuJoint row map to the members of each joint definition (include/b2cuda.h, "Fields by type")
enum FieldKind { F_VEC2, F_FLOAT, F_LIMIT_FLAG, F_MOTOR_FLAG };
struct Field
{
	const char* name;
	FieldKind kind;
	size_t offset; // of the first float in b2cuJoint (unused for the flags)
};
#define ROW(member) offsetof(b2cuJoint, member)
struct DefTable
{
	int32 type;
	const char* defType;
	Field fields[10];
};
const DefTable kDefs[] = {
	{B2CU_JOINT_REVOLUTE, "b2RevoluteJointDef", {{"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"referenceAngle", F_FLOAT, ROW(referenceAngle)}, {"enableLimit", F_LIMIT_FLAG, 0}, {"lowerAngle", F_FLOAT, ROW(lowerAngle)}, {"upperAngle", F_FLOAT, ROW(upperAngle)}, {"enableMotor", F_MOTOR_FLAG, 0}, {"motorSpeed", F_FLOAT, ROW(motorSpeed)}, {"maxMotorTorque", F_FLOAT, ROW(maxMotorTorque)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_PRISMATIC, "b2PrismaticJointDef", {{"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"localAxisA", F_VEC2, ROW(axis)}, {"referenceAngle", F_FLOAT, ROW(referenceAngle)}, {"enableLimit", F_LIMIT_FLAG, 0}, {"lowerTranslation", F_FLOAT, ROW(lowerAngle)}, {"upperTranslation", F_FLOAT, ROW(upperAngle)}, {"enableMotor", F_MOTOR_FLAG, 0}, {"motorSpeed", F_FLOAT, ROW(motorSpeed)}, {"maxMotorForce", F_FLOAT, ROW(maxMotorTorque)}}},
	{B2CU_JOINT_DISTANCE, "b2DistanceJointDef", {{"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"length", F_FLOAT, ROW(length)}, {"frequencyHz", F_FLOAT, ROW(frequencyHz)}, {"dampingRatio", F_FLOAT, ROW(dampingRatio)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_PULLEY, "b2PulleyJointDef", {{"groundAnchorA", F_VEC2, ROW(axis)}, {"groundAnchorB", F_VEC2, ROW(lowerAngle)}, {"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"lengthA", F_FLOAT, ROW(length)}, {"lengthB", F_FLOAT, ROW(referenceAngle)}, {"ratio", F_FLOAT, ROW(motorSpeed)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_MOUSE, "b2MouseJointDef", {{"target", F_VEC2, ROW(axis)}, {"maxForce", F_FLOAT, ROW(length)}, {"frequencyHz", F_FLOAT, ROW(frequencyHz)}, {"dampingRatio", F_FLOAT, ROW(dampingRatio)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_WHEEL, "b2WheelJointDef", {{"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"localAxisA", F_VEC2, ROW(axis)}, {"enableMotor", F_MOTOR_FLAG, 0}, {"motorSpeed", F_FLOAT, ROW(motorSpeed)}, {"maxMotorTorque", F_FLOAT, ROW(maxMotorTorque)}, {"frequencyHz", F_FLOAT, ROW(frequencyHz)}, {"dampingRatio", F_FLOAT, ROW(dampingRatio)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_WELD, "b2WeldJointDef", {{"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"referenceAngle", F_FLOAT, ROW(referenceAngle)}, {"frequencyHz", F_FLOAT, ROW(frequencyHz)}, {"dampingRatio", F_FLOAT, ROW(dampingRatio)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_FRICTION, "b2FrictionJointDef", {{"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"maxForce", F_FLOAT, ROW(length)}, {"maxTorque", F_FLOAT, ROW(maxMotorTorque)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_ROPE, "b2RopeJointDef", {{"localAnchorA", F_VEC2, ROW(localAnchorA)}, {"localAnchorB", F_VEC2, ROW(localAnchorB)}, {"maxLength", F_FLOAT, ROW(length)}, {nullptr, F_FLOAT, 0}}},
	{B2CU_JOINT_MOTOR, "b2MotorJointDef", {{"linearOffset", F_VEC2, ROW(axis)}, {"angularOffset", F_FLOAT, ROW(referenceAngle)}, {"maxForce", F_FLOAT, ROW(length)}, {"maxTorque", F_FLOAT, ROW(maxMotorTorque)}, {"correctionFactor", F_FLOAT, ROW(dampingRatio)}, {nullptr, F_FLOAT, 0}}},
};
#undef ROW

void DumpJointRow(const b2cuJoint& r, int32 index)
{
	const char* in = "  ";
	b2Log("{\n");
	if (r.type == B2CU_JOINT_GEAR)
	{
		b2Log("%sb2GearJointDef jd;\n", in);
		b2Log("%sjd.bodyA = bodies[%d];\n%sjd.bodyB = bodies[%d];\n", in, (int)r.bodyA, in, (int)r.bodyB);
		Flag(in, "jd.collideConnected", (r.flags & B2CU_JOINT_COLLIDE_CONNECTED) != 0);
		b2Log("%sjd.joint1 = joints[%d];\n%sjd.joint2 = joints[%d];\n", in, (int)r.frequencyHz, in, (int)r.dampingRatio);
		Float(in, "jd.ratio", r.motorSpeed);
	}
	else
	{
		for (size_t t = 0; t < sizeof(kDefs) / sizeof(kDefs[0]); ++t)
		{
			if (kDefs[t].type != r.type) continue;
			b2Log("%s%s jd;\n", in, kDefs[t].defType);
			b2Log("%sjd.bodyA = bodies[%d];\n%sjd.bodyB = bodies[%d];\n", in, (int)r.bodyA, in, (int)r.bodyB);
			Flag(in, "jd.collideConnected", (r.flags & B2CU_JOINT_COLLIDE_CONNECTED) != 0);
			for (int32 k = 0; k < 10 && kDefs[t].fields[k].name; ++k)
			{
				const Field& f = kDefs[t].fields[k];
				char lhs[48];
				snprintf(lhs, sizeof lhs, "jd.%s", f.name);
				const float32* v = reinterpret_cast<const float32*>(reinterpret_cast<const char*>(&r) + f.offset);
				if (f.kind == F_VEC2) Vec(in, lhs, v[0], v[1]);
				else if (f.kind == F_FLOAT) Float(in, lhs, v[0]);
				else if (f.kind == F_LIMIT_FLAG) Flag(in, lhs, (r.flags & B2CU_JOINT_ENABLE_LIMIT) != 0);
				else Flag(in, lhs, (r.flags & B2CU_JOINT_ENABLE_MOTOR) != 0);
			}
		}
	}
	b2Log("%sjoints[%d] = m_world->CreateJoint(&jd);\n}\n", in, (int)index);
}

} // namespace

void b2World::Dump()
{
	if (IsLocked()) return;
	RefreshBodies();
	RefreshJoints();

	Vec("", "b2Vec2 g; g", m_gravity.x, m_gravity.y);
	b2Log("m_world->SetGravity(g);\n");
	b2Log("b2Body** bodies = (b2Body**)b2Alloc(%d * sizeof(b2Body*));\n", (int)m_bodies.size());
	b2Log("b2Joint** joints = (b2Joint**)b2Alloc(%d * sizeof(b2Joint*));\n", (int)m_joints.size());

	for (size_t i = 0; i < m_bodies.size(); ++i)
	{
		const b2Body* b = m_bodies[i];
		const char* in = "  ";
		b2Log("{\n%sb2BodyDef bd;\n", in);
		b2Log("%sbd.type = b2BodyType(%d);\n", in, (int)b->GetType());
		Vec(in, "bd.position", b->GetPosition().x, b->GetPosition().y);
		Float(in, "bd.angle", b->GetAngle());
		Vec(in, "bd.linearVelocity", b->GetLinearVelocity().x, b->GetLinearVelocity().y);
		Float(in, "bd.angularVelocity", b->GetAngularVelocity());
		Float(in, "bd.linearDamping", b->GetLinearDamping());
		Float(in, "bd.angularDamping", b->GetAngularDamping());
		Flag(in, "bd.allowSleep", b->IsSleepingAllowed());
		Flag(in, "bd.awake", b->IsAwake());
		Flag(in, "bd.fixedRotation", b->IsFixedRotation());
		Flag(in, "bd.bullet", b->IsBullet());
		Flag(in, "bd.active", b->IsActive());
		Float(in, "bd.gravityScale", b->GetGravityScale());
		b2Log("%sbodies[%d] = m_world->CreateBody(&bd);\n", in, (int)i);
		// the fixture list is newest first; print oldest first so that the rebuilt world assigns the same proxy ids
		std::vector<const b2Fixture*> fixtures;
		for (const b2Fixture* f = b->GetFixtureList(); f; f = f->GetNext()) fixtures.push_back(f);
		for (size_t k = fixtures.size(); k-- > 0;)
		{
			const b2Fixture* f = fixtures[k];
			const char* in2 = "    ";
			b2Log("  {\n%sb2FixtureDef fd;\n", in2);
			Float(in2, "fd.friction", f->GetFriction());
			Float(in2, "fd.restitution", f->GetRestitution());
			Float(in2, "fd.density", f->GetDensity());
			Flag(in2, "fd.isSensor", f->IsSensor());
			Flag(in2, "fd.thickShape", f->IsThickShape());
			b2Log("%sfd.filter.categoryBits = uint16(%d);\n", in2, (int)f->GetFilterData().categoryBits);
			b2Log("%sfd.filter.maskBits = uint16(%d);\n", in2, (int)f->GetFilterData().maskBits);
			b2Log("%sfd.filter.groupIndex = int16(%d);\n", in2, (int)f->GetFilterData().groupIndex);
			DumpShape(f->GetShape());
			b2Log("%sfd.shape = &shape;\n%sbodies[%d]->CreateFixture(&fd);\n  }\n", in2, in2, (int)i);
		}
		b2Log("}\n");
	}

	// joints in table order: a gear joint always comes after the two joints it couples
	for (size_t i = 0; i < m_joints.size(); ++i)
	{
		b2cuJoint row;
		m_joints[i]->WriteRecord(&row);
		DumpJointRow(row, (int32)i);
	}

	b2Log("b2Free(joints);\nb2Free(bodies);\njoints = nullptr;\nbodies = nullptr;\n");
}
