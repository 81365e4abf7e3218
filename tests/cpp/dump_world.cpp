// Builds a small world with every shape class and several joint types (gear included) and prints b2World::Dump().
// With -DREBUILD the world is instead rebuilt from a previous dump (dump.inc) and dumped again: the two texts must agree.
#include <cstdio>

#include "Box2D/Box2D.h"

int main()
{
	b2World world(b2Vec2(0.0f, -10.0f));
	b2World* m_world = &world;
#ifdef REBUILD
#include "dump.inc"
#else
	b2BodyDef bd;
	b2Body* ground = world.CreateBody(&bd);
	b2EdgeShape edge;
	edge.Set(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f, 0.0f));
	ground->CreateFixture(&edge, 0.0f);
	b2Vec2 hill[4] = {b2Vec2(-5.0f, 0.0f), b2Vec2(-2.0f, 1.5f), b2Vec2(2.0f, 1.25f), b2Vec2(5.0f, 0.0f)};
	b2ChainShape chain;
	chain.CreateChain(hill, 4);
	ground->CreateFixture(&chain, 0.0f);

	bd.type = b2_dynamicBody;
	bd.position.Set(-3.0f, 12.0f);
	bd.angularVelocity = 1.5f;
	b2Body* wheelA = world.CreateBody(&bd);
	b2CircleShape disc;
	disc.m_radius = 1.0f;
	b2FixtureDef fd;
	fd.shape = &disc;
	fd.density = 5.0f;
	fd.friction = 0.3f;
	fd.filter.categoryBits = 0x0002;
	fd.filter.maskBits = 0xFFFD;
	wheelA->CreateFixture(&fd);

	bd.position.Set(0.0f, 12.0f);
	bd.angularVelocity = 0.0f;
	bd.bullet = true;
	b2Body* wheelB = world.CreateBody(&bd);
	disc.m_radius = 2.0f;
	b2FixtureDef sensor;
	sensor.shape = &disc;
	sensor.isSensor = true;
	wheelB->CreateFixture(&disc, 5.0f);
	wheelB->CreateFixture(&sensor);

	bd.bullet = false;
	bd.fixedRotation = true;
	bd.position.Set(2.5f, 12.0f);
	bd.angle = 0.25f;
	b2Body* rack = world.CreateBody(&bd);
	b2PolygonShape box;
	b2Vec2 corners[4] = {b2Vec2(-0.5f, -5.0f), b2Vec2(0.5f, -5.0f), b2Vec2(0.5f, 5.0f), b2Vec2(-0.5f, 5.0f)};
	box.Set(corners, 4); // through the hull, like the rebuilt one: the vertex order is canonical from the start
	b2FixtureDef thick;
	thick.shape = &box;
	thick.density = 5.0f;
	thick.thickShape = true;
	rack->CreateFixture(&thick);

	b2RevoluteJointDef hingeA;
	hingeA.Initialize(ground, wheelA, wheelA->GetPosition());
	hingeA.enableMotor = true;
	hingeA.motorSpeed = 1.0f;
	hingeA.maxMotorTorque = 100.0f;
	b2Joint* j1 = world.CreateJoint(&hingeA);
	b2RevoluteJointDef hingeB;
	hingeB.Initialize(ground, wheelB, wheelB->GetPosition());
	hingeB.enableLimit = true;
	hingeB.lowerAngle = -2.0f;
	hingeB.upperAngle = 3.0f;
	b2Joint* j2 = world.CreateJoint(&hingeB);
	b2PrismaticJointDef slide;
	slide.Initialize(ground, rack, rack->GetPosition(), b2Vec2(0.0f, 2.0f));
	slide.enableLimit = true;
	slide.lowerTranslation = -5.0f;
	slide.upperTranslation = 5.0f;
	b2Joint* j3 = world.CreateJoint(&slide);
	b2GearJointDef gearAB;
	gearAB.bodyA = wheelA;
	gearAB.bodyB = wheelB;
	gearAB.joint1 = j1;
	gearAB.joint2 = j2;
	gearAB.ratio = 2.0f;
	world.CreateJoint(&gearAB);
	b2GearJointDef gearBC;
	gearBC.bodyA = wheelB;
	gearBC.bodyB = rack;
	gearBC.joint1 = j2;
	gearBC.joint2 = j3;
	gearBC.ratio = -0.5f;
	world.CreateJoint(&gearBC);
	b2DistanceJointDef rod;
	rod.Initialize(wheelA, rack, wheelA->GetPosition(), rack->GetPosition());
	rod.frequencyHz = 3.0f;
	rod.dampingRatio = 0.25f;
	rod.collideConnected = true;
	world.CreateJoint(&rod);
	b2PulleyJointDef pulley;
	pulley.Initialize(wheelA, wheelB, b2Vec2(-3.0f, 20.0f), b2Vec2(0.0f, 20.0f), wheelA->GetPosition(), wheelB->GetPosition(), 1.5f);
	world.CreateJoint(&pulley);
	b2MotorJointDef servo;
	servo.Initialize(ground, rack);
	servo.maxForce = 50.0f;
	servo.correctionFactor = 0.5f;
	world.CreateJoint(&servo);
#endif
	fprintf(stderr, "%d bodies %d joints %d proxies\n", (int)world.GetBodyCount(), (int)world.GetJointCount(), (int)world.GetProxyCount());
	world.Dump();
	return 0;
}
