/*
 * oracle/ref_harness.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * C ABI over the UNMODIFIED reference (Box2D-MT compiled from /root/reference into
 * oracle/_ref/libb2ref.so by oracle/build_ref.py).  It builds reference worlds from flat scene
 * arrays, steps them with the reference's own b2ThreadPoolTaskExecutor, and exports the internal
 * state in the product's C-ABI record formats (include/b2cuda.h) so that tests can compare the
 * two sides field by field.
 */
#ifndef B2_REF_HARNESS_H
#define B2_REF_HARNESS_H

#include <stdint.h>
#include "b2cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Scene description: what user code would pass to b2World::CreateBody / b2Body::CreateFixture. */

enum
{
	B2REF_BODY_ALLOW_SLEEP = 1,
	B2REF_BODY_AWAKE = 2,
	B2REF_BODY_FIXED_ROTATION = 4,
	B2REF_BODY_BULLET = 8,
	B2REF_BODY_ACTIVE = 16
};

typedef struct b2refBodyDef
{
	int32_t type; /* b2BodyType */
	float px, py, angle;
	float vx, vy, w;
	float linearDamping, angularDamping, gravityScale;
	uint32_t flags;
} b2refBodyDef;

/* kind: how the shape is constructed, so that the reference's own constructors compute hull, normals, centroid */
enum
{
	B2REF_SHAPE_CIRCLE = 0,  /* radius, p = v[0] */
	B2REF_SHAPE_EDGE = 1,    /* v1 = v[0], v2 = v[1], v0 = v[2], v3 = v[3], flags = hasVertex0 | hasVertex3<<1 */
	B2REF_SHAPE_POLYGON = 2, /* b2PolygonShape::Set(v, count) */
	B2REF_SHAPE_BOX = 3,     /* SetAsBox(v[0].x, v[0].y) or, if flags&1, SetAsBox(hx, hy, center = v[1], angle = v[2].x) */
	B2REF_SHAPE_RAW = 4,     /* polygon given with explicit vertices AND normals (n[]) and centroid, copied as is */
	B2REF_SHAPE_CHAIN = 5    /* b2ChainShape of v[0..count): CreateLoop if flags&1 else CreateChain */
};

typedef struct b2refShapeDef
{
	int32_t kind;
	int32_t count;
	float radius;
	uint32_t flags;
	float v[8][2];
	float n[8][2];
	float centroid[2];
} b2refShapeDef;

typedef struct b2refFixtureDef
{
	int32_t body;
	int32_t shape;
	float density, friction, restitution;
	uint32_t flags; /* B2CU_PROXY_SENSOR | B2CU_PROXY_THICK */
	uint16_t categoryBits, maskBits;
	int16_t groupIndex;
	uint16_t pad;
} b2refFixtureDef;

typedef struct b2refWorld b2refWorld;

b2refWorld* b2ref_create(float gx, float gy, uint32_t worldFlags /* B2CU_WORLD_* */, int32_t threads);
void b2ref_destroy(b2refWorld* w);

/* Creates bodies in order; fixtures must be sorted by body and are created right after their body,
 * i.e. the usual "create body, add its fixtures, next body" order.  Returns 0 on success. */
int b2ref_build(b2refWorld* w, int32_t bodyCount, const b2refBodyDef* bodies, int32_t shapeCount,
                const b2refShapeDef* shapes, int32_t fixtureCount, const b2refFixtureDef* fixtures);

/* b2World::Step through b2ThreadPoolTaskExecutor. */
void b2ref_step(b2refWorld* w, float dt, int32_t velocityIterations, int32_t positionIterations);

/* Same phases as b2World::Step, but each island's contact array is ordered by the caller's ranks before
 * b2Island::Solve (the "permuted-order oracle").  keys are (min<<32|max) of fixture
 * indices.  Returns the number of island contacts that had no rank (0 when the solver sets agree). */
int b2ref_step_ordered(b2refWorld* w, float dt, int32_t velocityIterations, int32_t positionIterations,
                       int32_t orderCount, const uint64_t* keys);

void b2ref_counts(b2refWorld* w, int32_t* bodyCount, int32_t* fixtureCount, int32_t* contactCount);
float b2ref_inv_dt0(b2refWorld* w);

void b2ref_export_bodies(b2refWorld* w, b2cuBody* out);
/* one geometry record per fixture */
void b2ref_export_shapes(b2refWorld* w, b2cuShape* out);
/* one proxy per fixture (chains unsupported); shape index = fixture index; treeProxyIds optional */
void b2ref_export_proxies(b2refWorld* w, b2cuProxy* out, int32_t* treeProxyIds);
/* contacts sorted by key */
int b2ref_export_contacts(b2refWorld* w, int32_t capacity, b2cuContact* out);
/* Begin/End deferred events recorded during the last step, in callback order (keys of fixture indices) */
int b2ref_events(b2refWorld* w, int32_t kind, int32_t capacity, uint64_t* keys);
/* TOI candidate contacts (front partition of m_contacts), as sorted keys */
int b2ref_toi_candidates(b2refWorld* w, int32_t capacity, uint64_t* keys);
/* last step b2Profile (13 floats) */
void b2ref_profile(b2refWorld* w, float* out13);

/* mutators used by tests */
void b2ref_set_transform(b2refWorld* w, int32_t body, float x, float y, float angle);
void b2ref_set_type(b2refWorld* w, int32_t body, int32_t type); /* b2Body::SetType */
void b2ref_set_active(b2refWorld* w, int32_t body, int32_t on); /* b2Body::SetActive */
/* PostSolve recording: running digest over (contact key, impulse count, impulses) of every PostSolve call */
/* PreSolve test rule: the listener disables every contact whose key is a multiple of `modulus` (0: off) */
void b2ref_set_pre_solve_rule(b2refWorld* w, int32_t modulus);
void b2ref_pre_solve_digest(b2refWorld* w, uint64_t* digest, int64_t* count);
void b2ref_record_post_solve(b2refWorld* w, int32_t on);
void b2ref_post_solve_digest(b2refWorld* w, uint64_t* digest, int64_t* count);
/* installs a b2ContactFilter subclass: pairs whose fixture indices sum to a multiple of `modulus` never collide */
void b2ref_set_modulo_filter(b2refWorld* w, int32_t modulus);
/* b2Fixture::SetFilterData (which refilters) on fixture `fixture` (creation order) */
void b2ref_set_filter(b2refWorld* w, int32_t fixture, uint16_t categoryBits, uint16_t maskBits, int16_t groupIndex);
void b2ref_set_velocity(b2refWorld* w, int32_t body, float vx, float vy, float angw);
void b2ref_apply_force(b2refWorld* w, int32_t body, float fx, float fy, float torque);
void b2ref_set_awake(b2refWorld* w, int32_t body, int32_t awake);
/* b2Body::SetLinearDamping (0) / SetAngularDamping (1) / SetGravityScale (2) / SetBullet (3) / SetSleepingAllowed (4) */
void b2ref_set_body_param(b2refWorld* w, int32_t body, int32_t which, float value);
void b2ref_destroy_last_fixture(b2refWorld* w);

/* FNV-1a 32 over raw bytes of (pos.x, pos.y, angle) for all bodies in GetBodyList() order */
uint32_t b2ref_hash(b2refWorld* w);

/* Stand-alone manifold functions of the reference (b2Collide*.cpp) on b2cuShape records. */
void b2ref_collide(const b2cuShape* shapeA, const float xfA[4], const b2cuShape* shapeB, const float xfB[4],
                   b2cuManifold* out);
/* world queries of the reference (b2World::QueryAABB / RayCast), see ref_harness.cpp */
int32_t b2ref_query_aabb(b2refWorld* w, const float aabb[4], int32_t capacity, int32_t* out);
int32_t b2ref_ray_cast_closest(b2refWorld* w, const float p1[2], const float p2[2], float out[5]);
/* the reference's b2Distance with a cold simplex cache, on geometry records (same conventions as b2ref_collide) */
void b2ref_distance(const b2cuShape* shapeA, const float xfA[4], const b2cuShape* shapeB, const float xfB[4],
                    int32_t useRadii, b2cuDistanceResult* out);

/* joints (revolute) from b2cuJoint records; the solve order of b2ref_step_ordered; state export */
int32_t b2ref_set_joints(b2refWorld* w, int32_t count, const b2cuJoint* joints);
void b2ref_set_joint_order(b2refWorld* w, int32_t count, const int32_t* ids);
void b2ref_export_joints(b2refWorld* w, b2cuJoint* out);
void b2ref_joint_set_motor(b2refWorld* w, int32_t joint, int32_t enable, float speed, float maxTorque);
void b2ref_joint_set_limits(b2refWorld* w, int32_t joint, int32_t enable, float lower, float upper);
void b2ref_joint_set_spring(b2refWorld* w, int32_t joint, float length, float frequencyHz, float dampingRatio);
void b2ref_joint_set_target(b2refWorld* w, int32_t joint, float x, float y);
void b2ref_destroy_joint(b2refWorld* w, int32_t joint);
void b2ref_joint_readings(b2refWorld* w, float inv_dt, float* out6);
/* first pass of b2World::SolveTOI on the current state (see ref_harness.cpp) */
int32_t b2ref_first_toi(b2refWorld* w, uint64_t* key, float* alpha);
/* the reference's b2TimeOfImpact on geometry records and sweeps */
void b2ref_time_of_impact(const b2cuShape* shapeA, const b2cuSweep* sweepA, const b2cuShape* shapeB,
                          const b2cuSweep* sweepB, float tMax, b2cuToiResult* out);

/* The interposed sin/cos the reference build actually calls (checks that interposition works). */
void b2ref_sincos(float x, float* s, float* c);

#ifdef __cplusplus
}
#endif

#endif
