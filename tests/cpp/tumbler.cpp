// The Testbed's Tumbler (Testbed/Tests/Tumbler.h:24-98) as a user program against this library's host API: a dynamic
// container turned by the motor of a revolute joint, one small box dropped in per step.  Exercises b2World::CreateJoint,
// b2RevoluteJointDef and CreateBody between steps.  Prints the joint's readings and where the boxes are in the drum's frame.
#include <cstdio>

#include "Box2D/Box2D.h"

int main()
{
	b2CudaStepExecutor executor;
	b2World world(b2Vec2(0.0f, -10.0f));

	b2BodyDef groundDef;
	b2Body* ground = world.CreateBody(&groundDef);

	b2BodyDef drumDef;
	drumDef.type = b2_dynamicBody;
	drumDef.allowSleep = false;
	drumDef.position.Set(0.0f, 10.0f);
	b2Body* drum = world.CreateBody(&drumDef);
	const float32 walls[4][4] = {{0.5f, 10.0f, 10.0f, 0.0f}, {0.5f, 10.0f, -10.0f, 0.0f}, {10.0f, 0.5f, 0.0f, 10.0f}, {10.0f, 0.5f, 0.0f, -10.0f}};
	for (int32 k = 0; k < 4; ++k)
	{
		b2PolygonShape wall;
		wall.SetAsBox(walls[k][0], walls[k][1], b2Vec2(walls[k][2], walls[k][3]), 0.0f);
		drum->CreateFixture(&wall, 5.0f);
	}

	b2RevoluteJointDef hinge;
	hinge.bodyA = ground;
	hinge.bodyB = drum;
	hinge.localAnchorA.Set(0.0f, 10.0f);
	hinge.localAnchorB.Set(0.0f, 0.0f);
	hinge.motorSpeed = 0.05f * b2_pi;
	hinge.maxMotorTorque = 1e8f;
	hinge.enableMotor = true;
	b2RevoluteJoint* joint = static_cast<b2RevoluteJoint*>(world.CreateJoint(&hinge));
	if (joint == nullptr || world.GetJointCount() != 1 || drum->GetJointList() == nullptr) return 2;

	const int32 boxCount = 200, stepCount = 400;
	b2Body* boxes[boxCount];
	int32 made = 0;
	const float32 timeStep = 1.0f / 60.0f;
	for (int32 i = 0; i < stepCount; ++i)
	{
		if (made < boxCount)
		{
			b2BodyDef boxDef;
			boxDef.type = b2_dynamicBody;
			boxDef.position.Set(0.0f, 10.0f);
			boxes[made] = world.CreateBody(&boxDef);
			b2PolygonShape box;
			box.SetAsBox(0.125f, 0.125f);
			boxes[made]->CreateFixture(&box, 1.0f);
			++made;
		}
		world.Step(timeStep, 8, 3, executor);
		if (world.GetLastStepStatus() != 0)
		{
			fprintf(stderr, "step failed: %s\n", executor.GetLastError());
			return 1;
		}
	}
	float32 lo[2] = {1e9f, 1e9f}, hi[2] = {-1e9f, -1e9f};
	for (int32 k = 0; k < made; ++k)
	{
		b2Vec2 p = drum->GetLocalPoint(boxes[k]->GetPosition()); // in the drum's frame
		lo[0] = b2Min(lo[0], p.x);
		lo[1] = b2Min(lo[1], p.y);
		hi[0] = b2Max(hi[0], p.x);
		hi[1] = b2Max(hi[1], p.y);
	}
	printf("%.6f %.6f %.3f %.4f %.4f %.4f %.4f\n", joint->GetJointAngle(), joint->GetJointSpeed(), joint->GetMotorTorque(60.0f),
	       lo[0], lo[1], hi[0], hi[1]);
	return 0;
}
