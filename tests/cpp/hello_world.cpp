// The reference's HelloWorld program (HelloWorld/HelloWorld.cpp:29-106) written against this library's host API.
// The only change a user makes is the executor type: b2ThreadPoolTaskExecutor -> b2CudaStepExecutor.
#include <cstdio>

#include "Box2D/Box2D.h"

int main()
{
	b2CudaStepExecutor executor;

	b2World world(b2Vec2(0.0f, -10.0f));

	b2BodyDef groundBodyDef;
	groundBodyDef.position.Set(0.0f, -10.0f);
	b2Body* groundBody = world.CreateBody(&groundBodyDef);
	b2PolygonShape groundBox;
	groundBox.SetAsBox(50.0f, 10.0f);
	groundBody->CreateFixture(&groundBox, 0.0f);

	b2BodyDef bodyDef;
	bodyDef.type = b2_dynamicBody;
	bodyDef.position.Set(0.0f, 4.0f);
	b2Body* body = world.CreateBody(&bodyDef);
	b2PolygonShape dynamicBox;
	dynamicBox.SetAsBox(1.0f, 1.0f);
	b2FixtureDef fixtureDef;
	fixtureDef.shape = &dynamicBox;
	fixtureDef.density = 1.0f;
	fixtureDef.friction = 0.3f;
	body->CreateFixture(&fixtureDef);

	const float32 timeStep = 1.0f / 60.0f;
	for (int32 i = 0; i < 60; ++i)
	{
		world.Step(timeStep, 6, 2, executor);
		if (world.GetLastStepStatus() != 0)
		{
			fprintf(stderr, "step failed: %s\n", executor.GetLastError());
			return 1;
		}
		b2Vec2 position = body->GetPosition();
		float32 angle = body->GetAngle();
		printf("%4.2f %4.2f %4.2f\n", position.x, position.y, angle);
	}
	return 0;
}
