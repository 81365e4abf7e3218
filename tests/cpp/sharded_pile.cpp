// One scene on several GPUs as a user program: a pile of boxes and discs in a wide container is built as an ordinary
// b2World, handed to b2CudaShardedWorld, stepped, re-balanced in the middle (bodies and contacts change GPUs), stepped
// on and gathered back.  usage: sharded_pile GPUS [COLUMNS ROWS STEPS]
// Prints: bodies, strips, lost contacts, lowest body, fastest body, and a hash of the gathered positions.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "Box2D/Box2D.h"

int main(int argc, char** argv)
{
	const int32 gpus = argc > 1 ? atoi(argv[1]) : 2;
	const int32 columns = argc > 2 ? atoi(argv[2]) : 60;
	const int32 rows = argc > 3 ? atoi(argv[3]) : 10;
	const int32 steps = argc > 4 ? atoi(argv[4]) : 200;

	b2World scene(b2Vec2(0.0f, -10.0f));
	scene.SetAllowSleeping(false);
	scene.SetContinuousPhysics(false); // a sharded scene refuses time-of-impact events

	const float32 spacing = 0.6f, width = spacing * columns;
	{
		b2BodyDef bd;
		b2Body* container = scene.CreateBody(&bd);
		b2PolygonShape wall;
		b2FixtureDef fd;
		fd.shape = &wall;
		fd.thickShape = true;
		wall.SetAsBox(0.5f * width + 2.0f, 1.0f, b2Vec2(0.5f * width, -1.0f), 0.0f);
		container->CreateFixture(&fd);
		wall.SetAsBox(1.0f, 50.0f, b2Vec2(-2.0f, 50.0f), 0.0f);
		container->CreateFixture(&fd);
		wall.SetAsBox(1.0f, 50.0f, b2Vec2(width + 2.0f, 50.0f), 0.0f);
		container->CreateFixture(&fd);
	}
	uint32 seed = 12345u;
	for (int32 c = 0; c < columns; ++c)
		for (int32 r = 0; r < rows; ++r)
		{
			seed = seed * 1664525u + 1013904223u;
			const float32 jitter = ((seed >> 8) & 0xFFFF) / 65535.0f * 0.1f - 0.05f;
			b2BodyDef bd;
			bd.type = b2_dynamicBody;
			bd.position.Set(spacing * (c + 0.5f) + jitter, 0.5f + spacing * r);
			b2Body* body = scene.CreateBody(&bd);
			if ((c + r) & 1)
			{
				b2CircleShape disc;
				disc.m_radius = 0.25f;
				body->CreateFixture(&disc, 1.0f);
			}
			else
			{
				b2PolygonShape box;
				box.SetAsBox(0.22f, 0.22f);
				body->CreateFixture(&box, 1.0f);
			}
		}

	b2CudaShardedWorld sharded(scene, gpus, 2.0f);
	if (sharded.GetLastStatus() != 0)
	{
		fprintf(stderr, "sharding failed (%d): %s\n", sharded.GetLastStatus(), sharded.GetLastError());
		return 2;
	}
	sharded.SetTransport(false, false); // nothing is read back per step; Gather fetches what it needs
	for (int32 i = 0; i < steps; ++i)
	{
		if (i == steps / 2 && !sharded.Rebalance())
		{
			fprintf(stderr, "rebalance failed (%d): %s\n", sharded.GetLastStatus(), sharded.GetLastError());
			return 3;
		}
		if (!sharded.Step(1.0f / 60.0f, 8, 3))
		{
			fprintf(stderr, "step %d failed (%d): %s\n", i, sharded.GetLastStatus(), sharded.GetLastError());
			return 4;
		}
	}
	sharded.Gather(scene);

	float32 lowest = 1e9f, fastest = 0.0f;
	uint32 hash = 2166136261u;
	int32 bodies = 0;
	for (const b2Body* b = scene.GetBodyList(); b; b = b->GetNext())
	{
		if (b->GetType() != b2_dynamicBody) continue;
		++bodies;
		lowest = b2Min(lowest, b->GetPosition().y);
		fastest = b2Max(fastest, b->GetLinearVelocity().Length());
		float32 v[3] = {b->GetPosition().x, b->GetPosition().y, b->GetAngle()};
		unsigned char bytes[sizeof(v)];
		memcpy(bytes, v, sizeof(v));
		for (size_t k = 0; k < sizeof(v); ++k) hash = (hash ^ bytes[k]) * 16777619u;
	}
	int32 held = 0;
	for (int32 r = 0; r < sharded.GetShardCount(); ++r) held += sharded.GetStrip(r).GetBodyCount();
	printf("%d %d %d %d %.6f %.6f %08x\n", bodies, sharded.GetShardCount(), held, sharded.GetLostContacts(), lowest, fastest, hash);
	return 0;
}
