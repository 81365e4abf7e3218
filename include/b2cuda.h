/*
 * b2cuda.h -- C ABI of libb2cuda.so: the B200 (sm_100a) implementation of the
 * Box2D-MT b2World::Step hot path.
 *
 * This is the drop-in boundary.  Everything above it is host
 * C++ that keeps the reference API (b2World, b2Body, b2Fixture, b2ContactListener,
 * b2TaskExecutor); everything below it is hand-written CUDA.  Only plain pointers
 * and sizes cross it: no C++ types, no torch types.  The caller owns every host
 * buffer, the library owns every device buffer.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * the reference tree, Box2D/...):
 *
 *   b2cuCreateWorld / b2cuDestroyWorld   b2World::b2World / ~b2World       Dynamics/b2World.cpp:444-520
 *   b2cuSetWorldParams                   b2World::SetGravity, SetAllowSleeping, SetWarmStarting,
 *                                        SetContinuousPhysics, SetAutoClearForces
 *                                                                          Dynamics/b2World.h:133-228
 *   b2cuSetBodies / b2cuGetBodies        b2Body state (m_xf, m_sweep, velocities, forces, mass, flags)
 *                                                                          Dynamics/b2Body.h:471-508
 *   b2cuSetShapes                        b2PolygonShape / b2CircleShape / b2EdgeShape geometry
 *                                                                          Collision/Shapes/b2{Polygon,Circle,Edge}Shape.h
 *   b2cuSetProxies / b2cuGetProxies      b2Fixture + b2FixtureProxy + tree-leaf fat AABB
 *                                                                          Dynamics/b2Fixture.h:100-106, Collision/b2DynamicTree.h:36
 *   b2cuSetContacts / b2cuGetContacts    b2Contact persistent state (m_flags, m_manifold, mixes, TOI)
 *                                                                          Dynamics/Contacts/b2Contact.h:231-259
 *   b2cuSetJoints / b2cuGetJoints /      b2World::CreateJoint / DestroyJoint and the eleven b2Joint classes: parameters,
 *   b2cuGetJointCount / GetJointOrder    accumulated impulses, limit states   Dynamics/b2World.cpp:659-841, Dynamics/Joints/
 *   b2cuStep                             b2World::Step(dt, velocityIterations, positionIterations, executor)
 *                                                                          Dynamics/b2World.cpp:1613-1710
 *   b2cuGetContactsByKey                 b2Contact objects handed to listener callbacks (manifold, flags, mixes),
 *                                        fetched for a list of keys only       Dynamics/Contacts/b2Contact.h:95-176
 *   b2cuGetEvents                        deferred BeginContact/EndContact buffers, sorted by proxy-id key
 *                                                                          Dynamics/b2ContactManager.cpp:388-439
 *   b2cuGetEventContacts                 the two above in one round trip: what BeginContact / EndContact receive
 *                                                                          Dynamics/b2WorldCallbacks.h:84-107
 *   b2cuGetBodyStates /                  b2Island::Solve writing the new state into the b2Body objects
 *   b2cuSetBodyMirror
 *                                                                          Dynamics/b2Island.cpp:339-348
 *   b2cuSetBodyForces                    b2Body::ApplyForce / ApplyTorque on awake bodies                    Dynamics/b2Body.h:740-790
 *   b2cuHostAlloc / b2cuHostFree         the world's own allocation of its bodies (b2BlockAllocator)
 *                                                                          Common/b2BlockAllocator.cpp:93-170
 *   b2cuSetPreSolveHook /                b2ContactListener::PreSolve between Collide and Solve, b2Contact::SetEnabled
 *   b2cuGetPreSolveContacts /                                               Dynamics/Contacts/b2Contact.cpp:283-297,
 *   b2cuDisableContacts                                                     Dynamics/b2ContactManager.cpp:430-433
 *   b2cuSetPairFilter                    b2ContactFilter::ShouldCollide of a user subclass, called from AddPair
 *                                                                          Dynamics/b2WorldCallbacks.h:52-63, b2ContactManager.cpp:280-285
 *   b2cuGetSolverOrder                   (new) the colour-ordered constraint list the coloured Gauss-Seidel
 *                                        used this step; feeds the permuted-order oracle
 *   b2cuGetToiCandidates                 TOI-eligible front partition of b2ContactManager::m_contacts
 *                                                                          Dynamics/b2ContactManager.cpp:659-713, b2World.cpp:317-341
 *   b2cuGetToiEvents                     BeginContact / EndContact calls made from inside b2World::SolveTOI
 *                                                                          Dynamics/b2World.cpp:851-1024, Contacts/b2Contact.cpp:247-281
 *   b2cuShardConfigure / GetLink /       (new) spatial sharding of one large world over the GPUs of a box: halo
 *   Connect                              bodies + per-iteration halo exchange over NVLink peer memory, SURVEY.md 8e
 *   b2cuQueryAABB / RayCastCandidates    b2World::QueryAABB / RayCast (tree queries on the fat boxes)
 *                                                                          Dynamics/b2World.cpp:1752-1795
 *   b2cuDistancePairs                    b2Distance (GJK), batched; b2TestOverlap = its distance under 10 epsilon
 *                                                                          Collision/b2Distance.cpp:452-603, b2Collision.cpp:233-252
 *   b2cuTimeOfImpactPairs                b2TimeOfImpact (conservative advancement), batched
 *                                                                          Collision/b2TimeOfImpact.cpp:256-497
 *   b2cuCollidePairs                     b2CollidePolygons / b2CollideCircles / b2CollidePolygonAndCircle /
 *                                        b2CollideEdgeAndCircle / b2CollideEdgeAndPolygon, batched
 *                                                                          Collision/b2CollidePolygon.cpp, b2CollideCircle.cpp, b2CollideEdge.cpp
 *
 * All functions return B2CU_OK (0) or a negative b2cuStatus; b2cuGetLastError gives
 * the message.  There is no CPU fallback anywhere behind this interface.
 */
#ifndef B2CUDA_H
#define B2CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2CU_API __attribute__((visibility("default")))

typedef struct b2cuWorld b2cuWorld;

typedef enum b2cuStatus
{
	B2CU_OK = 0,
	B2CU_ERR_CUDA = -1,        /* a CUDA runtime call failed */
	B2CU_ERR_CAPACITY = -2,    /* a device buffer overflowed its capacity (grow and retry) */
	B2CU_ERR_ARGUMENT = -3,    /* bad argument */
	B2CU_ERR_UNSUPPORTED = -4, /* feature outside the GPU path (a joint type that is not solved, joints in a sharded world) */
	B2CU_ERR_NO_DEVICE = -5    /* no CUDA device: there is no CPU fallback */
} b2cuStatus;

/* ---- body ------------------------------------------------------------------ */

/* b2BodyType, Dynamics/b2Body.h:36-46 */
enum { B2CU_STATIC_BODY = 0, B2CU_KINEMATIC_BODY = 1, B2CU_DYNAMIC_BODY = 2 };

/* b2cuBody.flags: bits 0-1 = body type, then the b2Body m_flags (Dynamics/b2Body.h:441-449) */
enum
{
	B2CU_BODY_TYPE_MASK = 0x0003,
	B2CU_BODY_ISLAND = 0x0004,
	B2CU_BODY_AWAKE = 0x0008,
	B2CU_BODY_AUTOSLEEP = 0x0010,
	B2CU_BODY_BULLET = 0x0020,
	B2CU_BODY_FIXED_ROTATION = 0x0040,
	B2CU_BODY_ACTIVE = 0x0080,
	B2CU_BODY_GHOST = 0x0100   /* (new) halo copy of a body owned by the neighbouring shard, see b2cuShardConfigure */
};

typedef struct b2cuBody
{
	float px, py, qs, qc;          /* m_xf: origin position, sin, cos */
	float cx, cy, a;               /* m_sweep.c, m_sweep.a */
	float c0x, c0y, a0, alpha0;    /* m_sweep.c0, a0, alpha0 */
	float lcx, lcy;                /* m_sweep.localCenter */
	float vx, vy, w;               /* m_linearVelocity, m_angularVelocity */
	float fx, fy, torque;          /* m_force, m_torque */
	float invMass, invI;
	float linearDamping, angularDamping, gravityScale;
	float sleepTime;
	uint32_t flags;
} b2cuBody;

/* The part of b2cuBody that a step changes and callers read back (48 bytes): what b2Island::Solve writes into the
 * b2Body objects (Dynamics/b2Island.cpp:339-348: m_xf, m_sweep.c / a, velocities) plus sleep time and flags.  This is the
 * record of the per-step body mirror (b2cuSetBodyMirror).  The rest of b2cuBody only ever travels host -> device (local
 * centre, mass, damping, gravity scale, applied forces) or is internal to the step: */
typedef struct b2cuBodyState
{
	float px, py, qs, qc;          /* m_xf: origin position, sin, cos */
	float cx, cy, a;               /* m_sweep.c, m_sweep.a */
	float vx, vy, w;               /* m_linearVelocity, m_angularVelocity */
	float sleepTime;
	uint32_t flags;
} b2cuBodyState;

/* ... the start of the body's sweep, m_sweep.c0 / a0 / alpha0 (Common/b2Math.h:382-410): where the last solve picked
 * the body up.  The step needs it (SynchronizeFixtures, time of impact); no accessor of b2Body returns it, so it stays
 * on the device and is fetched only when the caller is about to re-upload whole b2cuBody rows. */
typedef struct b2cuSweepStart
{
	float c0x, c0y, a0, alpha0;
} b2cuSweepStart;

/* ---- shape geometry ---------------------------------------------------------- */

/* b2Shape::Type, Collision/Shapes/b2Shape.h:48-55 (chain is outside the GPU path) */
enum { B2CU_SHAPE_CIRCLE = 0, B2CU_SHAPE_EDGE = 1, B2CU_SHAPE_POLYGON = 2 };

/* edge records: ghost vertices present; CHAIN_CHILD: the edge is a segment of a b2ChainShape, whose child AABB has no
 * radius margin (Collision/Shapes/b2ChainShape.cpp:173-189 against b2EdgeShape.cpp:116-129) */
enum { B2CU_EDGE_HAS_VERTEX0 = 1, B2CU_EDGE_HAS_VERTEX3 = 2, B2CU_EDGE_CHAIN_CHILD = 4 };

#define B2CU_MAX_POLYGON_VERTICES 8

/* One geometry record, 160 bytes.
 *   circle : v[0] = m_p
 *   edge   : v[0] = m_vertex1, v[1] = m_vertex2, v[2] = m_vertex0, v[3] = m_vertex3, flags = hasVertex bits;
 *            a chain shape is one edge record per segment (flags |= B2CU_EDGE_CHAIN_CHILD), one proxy each
 *   polygon: v[i] = m_vertices[i], n[i] = m_normals[i], centroid = m_centroid, count = m_count          */
typedef struct b2cuShape
{
	int32_t type;
	int32_t count;
	float radius;
	uint32_t flags;
	float v[B2CU_MAX_POLYGON_VERTICES][2];
	float n[B2CU_MAX_POLYGON_VERTICES][2];
	float centroid[2];
	float pad[2];
} b2cuShape;

/* ---- proxy = fixture child ---------------------------------------------------- */

enum
{
	B2CU_PROXY_SENSOR = 0x0001,
	B2CU_PROXY_THICK = 0x0002,
	B2CU_PROXY_MOVED = 0x0004,  /* in the broad-phase move buffer (b2BroadPhase::BufferMove, TouchProxy): its new pairs are
	                               found by the next FindNewContacts, i.e. at the END of the next step */
	B2CU_PROXY_NEW = 0x0010,    /* with MOVED: the fixture is new (b2World::e_newFixture, set by b2Body::CreateFixture): the
	                               next step finds the pairs of the whole move buffer FIRST (Dynamics/b2World.cpp:1628-1639) */
	B2CU_PROXY_INACTIVE = 0x0040, /* the body is inactive (b2Body::SetActive(false) destroys its proxies, Dynamics/b2Body.cpp:
	                                 496-544): the slot keeps its id but takes no part in the broad-phase or in queries */
	B2CU_PROXY_REFILTER = 0x0020 /* with MOVED: b2Fixture::Refilter (Dynamics/b2Fixture.cpp:187-220): the contacts of this
	                               proxy are re-checked against the filters by the next Collide (e_filterFlag) */
};

typedef struct b2cuProxy
{
	float aabb[4];      /* b2FixtureProxy::aabb, swept tight AABB (lower.xy, upper.xy) */
	float fat[4];       /* tree-leaf fat AABB */
	int32_t body;       /* dense body index */
	int32_t shape;      /* index into the geometry table */
	float friction, restitution;
	uint16_t categoryBits, maskBits;
	int16_t groupIndex;
	uint16_t flags;
	int32_t fixture;    /* caller's fixture id (opaque to the device) */
	int32_t child;      /* child index (always 0: chains are outside the GPU path) */
} b2cuProxy;

/* ---- contact ------------------------------------------------------------------ */

/* b2Contact m_flags, Dynamics/Contacts/b2Contact.h:178-203 */
enum
{
	B2CU_CONTACT_ISLAND = 0x0001,
	B2CU_CONTACT_TOUCHING = 0x0002,
	B2CU_CONTACT_ENABLED = 0x0004,
	B2CU_CONTACT_FILTER = 0x0008,
	B2CU_CONTACT_BULLET_HIT = 0x0010,
	B2CU_CONTACT_TOI = 0x0020,
	B2CU_CONTACT_TOI_CANDIDATE = 0x0040,
	B2CU_CONTACT_INACTIVE = 0x0080
};

/* b2Manifold::Type, Collision/b2Collision.h:93-98 */
enum { B2CU_MANIFOLD_CIRCLES = 0, B2CU_MANIFOLD_FACE_A = 1, B2CU_MANIFOLD_FACE_B = 2 };

/* b2Manifold, 64 bytes (Collision/b2Collision.h:76-107) */
typedef struct b2cuManifold
{
	float localNormal[2];
	float localPoint[2];
	struct
	{
		float localPoint[2];
		float normalImpulse;
		float tangentImpulse;
	} points[2];
	uint32_t id[2];        /* b2ContactID::key: indexA | indexB<<8 | typeA<<16 | typeB<<24 */
	int32_t type;
	int32_t pointCount;
} b2cuManifold;

typedef struct b2cuContact
{
	int32_t proxyA, proxyB;  /* fixture A / fixture B side, after the primary-type swap (b2Contact.cpp:82-93) */
	uint32_t flags;
	float friction, restitution, tangentSpeed;
	int32_t toiCount;
	float toi;
	b2cuManifold manifold;
	/* creation order: the reference links a new contact at the HEAD of both bodies' contact lists (b2ContactManager::
	 * OnContactCreate, Dynamics/b2ContactManager.cpp:530-556) and b2World::StepSolveTOI walks those lists (b2World.cpp:
	 * 899-970), so which 32 contacts join a time-of-impact island, and in which order they are solved, follows the order
	 * of creation.  A body's list is its contacts by (stamp descending, key descending): every batch of new contacts
	 * (one b2ContactManager::FindNewContacts call, created in key order) gets the next stamp.  Round-trips through
	 * b2cuGetContacts / b2cuSetContacts; any values with that meaning do (e.g. the rank in b2World::GetContactList). */
	uint32_t stamp;
	uint32_t reserved;
} b2cuContact;

/* ---- world -------------------------------------------------------------------- */

enum
{
	B2CU_WORLD_ALLOW_SLEEP = 0x0001,
	B2CU_WORLD_WARM_STARTING = 0x0002,
	B2CU_WORLD_CONTINUOUS = 0x0004,
	B2CU_WORLD_SUB_STEPPING = 0x0008,
	B2CU_WORLD_CLEAR_FORCES = 0x0010
};

typedef struct b2cuWorldDef
{
	int32_t device;           /* CUDA device ordinal */
	float gravity[2];
	uint32_t flags;           /* B2CU_WORLD_* (b2World defaults: all but SUB_STEPPING, b2World.cpp:460-474) */
	int32_t bodyCapacity;     /* initial capacities; buffers grow geometrically */
	int32_t proxyCapacity;
	int32_t shapeCapacity;
	int32_t contactCapacity;
} b2cuWorldDef;

/* contact key = (min(proxyA,proxyB) << 32) | max(proxyA,proxyB): the b2ContactProxyIds ordering key
 * (Dynamics/Contacts/b2Contact.h:65-77) over dense proxy ids assigned in fixture creation order. */
typedef uint64_t b2cuContactKey;

/* Per-step report.  The first 13 floats mirror b2Profile (Dynamics/b2TimeStep.h:25-40), in ms,
 * measured with CUDA events on the step stream. */
typedef struct b2cuStepInfo
{
	float step, collide, solve, solveTraversal, solveInit, solveVelocity, solvePosition;
	float solveTOI, broadphase, broadphaseSyncFixtures, broadphaseFindContacts, locking, reserved;
	int32_t bodyCount, proxyCount, contactCount;
	int32_t touchingCount;     /* touching solid contacts after Collide */
	int32_t constraintCount;   /* contacts handed to the solver (touching, enabled, in an awake island) */
	int32_t colourCount;       /* colours used by the coloured Gauss-Seidel this step */
	int32_t overflowCount;     /* constraints that found no colour and were solved serially */
	int32_t islandBodyCount;   /* non-static bodies that were in an awake island */
	int32_t awakeBodyCount;    /* non-static bodies awake at the end of the step */
	int32_t moveCount;         /* proxies whose fat AABB moved */
	int32_t newContactCount, destroyedContactCount;
	int32_t beginCount, endCount;
	int32_t toiCandidateCount;
	int32_t kernelLaunches;    /* kernels launched by this step */
	/* b2World::SolveTOI (Dynamics/b2World.cpp:1026-1092) on a continuous world.  toiMinKey / toiMinAlpha: the result of
	 * its FIRST FindMinToiContact pass (:1525-1580), the candidate with the earliest time of impact.  toiSubSteps: how
	 * many time-of-impact events (b2World::StepSolveTOI, :851-1024) the step then executed.  toiEventPending = 1 only
	 * on a sub-stepping world (b2World::SetSubStepping) whose step stopped after one event with more to come
	 * (m_stepComplete == false): the next b2cuStep continues with the events and skips the solve. */
	int32_t toiEventPending;
	b2cuContactKey toiMinKey;  /* ~0 when there is no candidate */
	float toiMinAlpha;         /* 1 when there is no candidate */
	int32_t toiSubSteps;
	int32_t toiEventCount;     /* BeginContact / EndContact events raised inside the sub-steps (b2cuGetToiEvents) */
	int32_t toiNewContactCount; /* contacts created by the sub-steps' FindNewContacts */
} b2cuStepInfo;

enum { B2CU_EVENT_BEGIN = 0, B2CU_EVENT_END = 1 };

B2CU_API int b2cuGetDeviceCount(void);
B2CU_API const char* b2cuVersion(void);

B2CU_API int b2cuCreateWorld(const b2cuWorldDef* def, b2cuWorld** out);
B2CU_API void b2cuDestroyWorld(b2cuWorld* w);
B2CU_API const char* b2cuGetLastError(const b2cuWorld* w);

B2CU_API int b2cuSetWorldParams(b2cuWorld* w, const float gravity[2], uint32_t flags);
/* m_inv_dt0 (b2World.h:340): previous step's 1/dt, scales warm-start impulses */
B2CU_API int b2cuSetInvDt0(b2cuWorld* w, float invDt0);

/* Resize the live element counts (new elements must then be filled with the Set calls). */
B2CU_API int b2cuSetCounts(b2cuWorld* w, int32_t bodyCount, int32_t shapeCount, int32_t proxyCount);

/* Page-locked host memory for the caller's record arrays (body / proxy / contact mirrors).  The reference keeps its
 * bodies in host memory it allocates itself (b2BlockAllocator, Common/b2BlockAllocator.cpp:93-170); a mirror that
 * lives in memory from b2cuHostAlloc moves over PCIe in one DMA transfer instead of being staged by the driver.
 * Falls back to ordinary memory when no CUDA device is present.  Any pointer is accepted by the Set / Get calls. */
B2CU_API void* b2cuHostAlloc(size_t bytes);
B2CU_API void b2cuHostFree(void* p);

B2CU_API int b2cuSetBodies(b2cuWorld* w, int32_t first, int32_t count, const b2cuBody* bodies);
B2CU_API int b2cuGetBodies(b2cuWorld* w, int32_t first, int32_t count, b2cuBody* bodies);
/* The device -> host direction of the bodies: only the b2cuBodyState part (what a step changes). */
B2CU_API int b2cuGetBodyStates(b2cuWorld* w, int32_t first, int32_t count, b2cuBodyState* states);
B2CU_API int b2cuGetBodySweepStarts(b2cuWorld* w, int32_t first, int32_t count, b2cuSweepStart* starts);
/* b2World::Step leaves the new transforms and velocities in the b2Body objects (b2Island::Solve, Dynamics/b2Island.cpp:
 * 339-348).  With a mirror registered, every b2cuStep does the same for the caller's records of bodies [0, count):
 * the copy starts as soon as the solver has finished with the bodies and overlaps the broad-phase part of the step;
 * when b2cuStep returns the mirror is current (no b2cuGetBodyStates needed).  `mirror` should come from
 * b2cuHostAlloc; NULL / 0 removes it.  The mirror must stay valid until it is replaced or the world destroyed. */
B2CU_API int b2cuSetBodyMirror(b2cuWorld* w, b2cuBodyState* mirror, int32_t count);
/* b2Body::ApplyForce / ApplyForceToCenter / ApplyTorque on bodies that are awake only add to m_force / m_torque
 * (Dynamics/b2Body.h:740-790): the per-step user input of a running simulation.  forces = (fx, fy, torque) per body,
 * 12 bytes instead of the 104 of a whole b2cuBody row; nothing else of the bodies is touched. */
B2CU_API int b2cuSetBodyForces(b2cuWorld* w, int32_t first, int32_t count, const float* forces);
/* With prefetch on, b2cuStep itself gathers the contact records of the step's begin / end events and starts their
 * copy to the host before it returns, so that b2cuGetEventContacts finds them there (a listener is installed:
 * the callbacks will ask for them anyway).  Off by default. */
B2CU_API int b2cuSetEventPrefetch(b2cuWorld* w, int32_t on);
B2CU_API int b2cuSetShapes(b2cuWorld* w, int32_t first, int32_t count, const b2cuShape* shapes);
B2CU_API int b2cuSetProxies(b2cuWorld* w, int32_t first, int32_t count, const b2cuProxy* proxies);
B2CU_API int b2cuGetProxies(b2cuWorld* w, int32_t first, int32_t count, b2cuProxy* proxies);

/* Replace the whole contact set (teacher forcing, checkpoint restore).  Any order; sorted on device. */
B2CU_API int b2cuSetContacts(b2cuWorld* w, int32_t count, const b2cuContact* contacts);
B2CU_API int b2cuGetContactCount(b2cuWorld* w, int32_t* count);

/* Joints (Dynamics/Joints/b2Joint.h:28-226).  This version solves
 *   revolute joints   b2RevoluteJoint.cpp:64-400   point constraint, angular limit, motor
 *   distance joints   b2DistanceJoint.cpp:63-222   rigid or spring-damper rod between two anchors
 *   weld joints       b2WeldJoint.cpp:59-308       point + angle, rigid or with a soft angle
 *   prismatic joints  b2PrismaticJoint.cpp:127-478 slider along an axis of body A, translation limit, motor
 *   wheel joints      b2WheelJoint.cpp:78-318      point on a line of body A with a suspension spring, rotational motor
 *   rope joints       b2RopeJoint.cpp:47-195       maximum distance between two anchors
 *   friction joints   b2FrictionJoint.cpp:58-190   bounded linear and angular friction between two bodies
 *   motor joints      b2MotorJoint.cpp:66-200      drives body B to an offset from body A with bounded force / torque
 *   pulley joints     b2PulleyJoint.cpp:74-264     lengthA + ratio * lengthB constant over two ground anchors
 *   mouse joints      b2MouseJoint.cpp:96-190      soft bounded pull of a point of body B towards a world target
 *   gear joints       b2GearJoint.cpp:131-390      couples the coordinates of two revolute / prismatic joints (four bodies)
 * as rows of the coloured solver: inside every velocity iteration the joints run before the contacts, inside every
 * position iteration after them, as b2Island::Solve orders them (Dynamics/b2Island.cpp:259-273, :323-327, :363-380),
 * with warm starting.  A joint links the islands of its two bodies (b2World.cpp:1286-1320) and, unless
 * COLLIDE_CONNECTED, keeps them from colliding (b2Body::ShouldCollide, b2Body.cpp:428-449).  `type` uses b2JointType's
 * values: all eleven types are solved.
 * Fields by type: revolute   referenceAngle, lowerAngle, upperAngle, maxMotorTorque, motorSpeed, flags LIMIT / MOTOR
 *                 prismatic  axis (b2PrismaticJointDef::localAxisA as given; normalised as the constructor does),
 *                            referenceAngle, lowerAngle / upperAngle (= lower / upper translation), maxMotorTorque
 *                            (= maxMotorForce), motorSpeed, flags LIMIT / MOTOR
 *                 distance   length, frequencyHz, dampingRatio
 *                 weld       referenceAngle, frequencyHz, dampingRatio
 *                 wheel      axis (localAxisA, used as given), maxMotorTorque, motorSpeed, frequencyHz, dampingRatio,
 *                            flag MOTOR; impulse[0] = m_impulse, impulse[1] = m_springImpulse
 *                 rope       length (= maxLength); limitState = m_state
 *                 friction   length (= maxForce), maxMotorTorque (= maxTorque); impulse[0..1] linear, impulse[2] angular
 *                 motor      axis (= linearOffset), referenceAngle (= angularOffset), length (= maxForce),
 *                            maxMotorTorque (= maxTorque), dampingRatio (= correctionFactor); impulses as friction
 *                 pulley     axis (= groundAnchorA), lowerAngle / upperAngle (= groundAnchorB.x / .y), length (= lengthA),
 *                            referenceAngle (= lengthB), motorSpeed (= ratio)
 *                 mouse      axis (= target), length (= maxForce), frequencyHz, dampingRatio, maxMotorTorque (= the mass of
 *                            body B, b2Body::GetMass(): the reference reads the mass, not its inverse); localAnchorB is
 *                            the grabbed point; body A takes no part in the solve
 *                 gear       bodyA / bodyB = second bodies of joint 1 / 2, limitState / reserved = their first bodies (C /
 *                            D, body ids), flags GEAR_PRISMATIC_1 / _2 (else revolute), localAnchorA / B as in joint 1 / 2,
 *                            axis (= localAnchorC), lowerAngle / upperAngle (= localAnchorD), work[0..1] (= localAxisC),
 *                            work[2..3] (= localAxisD), referenceAngle (= referenceAngleA), maxMotorTorque (= referenceAngleB),
 *                            motorSpeed (= ratio), length (= the constant coordinateA + ratio * coordinateB the
 *                            constructor computes), frequencyHz / dampingRatio (= table ids of joint 1 / 2, for the
 *                            host side only); lastSolve = m_JvAC, m_JwA
 * impulse / motorImpulse / limitState are the joint's persistent solver state (m_impulse -- a scalar in impulse[0] for
 * the distance joint --, m_motorImpulse, m_limitState) and round-trip through Get / Set.  lastSolve is written by the
 * step with the world-space directions of its solve, which GetReactionForce needs: the distance and rope joints' m_u
 * in [0..1], the pulley joint's m_uB in [0..1], the prismatic joint's m_axis in [0..1] and m_perp in [2..3], the wheel
 * joint's m_ax and m_ay likewise.
 * work is solver scratch that the reference carries from step to step (wheel: m_sAx, m_sBx). */
enum
{
	B2CU_JOINT_REVOLUTE = 1,
	B2CU_JOINT_PRISMATIC = 2,
	B2CU_JOINT_DISTANCE = 3,
	B2CU_JOINT_PULLEY = 4,
	B2CU_JOINT_MOUSE = 5,
	B2CU_JOINT_GEAR = 6,
	B2CU_JOINT_WHEEL = 7,
	B2CU_JOINT_WELD = 8,
	B2CU_JOINT_FRICTION = 9,
	B2CU_JOINT_ROPE = 10,
	B2CU_JOINT_MOTOR = 11
};
enum
{
	B2CU_JOINT_COLLIDE_CONNECTED = 1,
	B2CU_JOINT_ENABLE_LIMIT = 2,
	B2CU_JOINT_ENABLE_MOTOR = 4,
	B2CU_JOINT_GEAR_PRISMATIC_1 = 0x100, /* gear: joint 1 / joint 2 is a prismatic joint (else revolute) */
	B2CU_JOINT_GEAR_PRISMATIC_2 = 0x200
};
enum { B2CU_LIMIT_INACTIVE = 0, B2CU_LIMIT_AT_LOWER = 1, B2CU_LIMIT_AT_UPPER = 2, B2CU_LIMIT_EQUAL = 3 };
typedef struct b2cuJoint
{
	int32_t type;
	int32_t bodyA, bodyB;
	uint32_t flags;
	float localAnchorA[2], localAnchorB[2];
	float referenceAngle, lowerAngle, upperAngle;
	float maxMotorTorque, motorSpeed;
	float length, frequencyHz, dampingRatio;
	float axis[2];
	float lastSolve[4];
	float work[4];
	float impulse[3];
	float motorImpulse;
	int32_t limitState;
	int32_t reserved;
} b2cuJoint;
/* Replace the world's joint table (b2World::CreateJoint / DestroyJoint, b2World.cpp:659-841, applied as a whole: the
 * caller keeps the list).  Joint ids are indices into this table. */
B2CU_API int b2cuSetJoints(b2cuWorld* w, int32_t count, const b2cuJoint* joints);
B2CU_API int b2cuGetJointCount(b2cuWorld* w, int32_t* count);
B2CU_API int b2cuGetJoints(b2cuWorld* w, int32_t first, int32_t count, b2cuJoint* joints);
/* The order in which the joints are solved inside an iteration (colour classes of joints that share no dynamic body,
 * ascending id inside a class): what a sequential Gauss-Seidel must follow to reproduce the step bit for bit. */
B2CU_API int b2cuGetJointOrder(b2cuWorld* w, int32_t capacity, int32_t* jointIds, int32_t* count);
/* Contacts in key order. */
B2CU_API int b2cuGetContacts(b2cuWorld* w, int32_t capacity, b2cuContact* contacts, int32_t* count);

/* Records of the contacts with the given keys, out[i] for keys[i].  A key that is no longer in the contact set
 * (e.g. the contact of an EndContact event that was destroyed) yields a record with the proxies of the key in
 * primary-type order, flags 0 and an empty manifold. */
B2CU_API int b2cuGetContactsByKey(b2cuWorld* w, int32_t count, const b2cuContactKey* keys, b2cuContact* out);
/* b2cuGetEvents and b2cuGetContactsByKey in one round trip: the events of `kind` in callback order together with the
 * record of each event's contact -- what b2ContactListener::BeginContact / EndContact receive as b2Contact*
 * (Dynamics/b2WorldCallbacks.h:84-107). */
B2CU_API int b2cuGetEventContacts(b2cuWorld* w, int32_t kind, int32_t capacity, b2cuContactKey* keys,
                                  b2cuContact* records, int32_t* count);

/* b2ContactFilter::ShouldCollide of a user subclass (Dynamics/b2WorldCallbacks.h:52-63), which the reference calls
 * from b2ContactManager::AddPair when two fat boxes begin to overlap (Dynamics/b2ContactManager.cpp:280-285).  With a
 * pair filter set, the device no longer applies the default category / mask / group rule: after every pair search of
 * b2cuStep the candidate pairs (same-body, existing-contact and b2Body::ShouldCollide checks already done) are handed
 * to `fn` on the calling thread, in one batch, before any contact is created or any body woken; fn sets keep[i] = 0 to
 * reject pair i.  A non-zero return value of fn aborts the step with B2CU_ERR_ARGUMENT.  NULL restores the default. */
typedef int (*b2cuPairFilterFn)(void* user, const b2cuContactKey* keys, int32_t count, uint8_t* keep);
B2CU_API int b2cuSetPairFilter(b2cuWorld* w, b2cuPairFilterFn fn, void* user);

/* b2ContactListener::PreSolve (Dynamics/b2WorldCallbacks.h:109-121): the reference calls it from b2Contact::Update for
 * every touching contact it has just updated, with the manifold of the previous step, and user code may switch the
 * contact off for this step there (b2Contact::SetEnabled(false); Dynamics/Contacts/b2Contact.cpp:283-297, deferred
 * calls in b2ContactManager::FinishCollide, b2ContactManager.cpp:430-433).  With a hook set, b2cuStep calls fn on the
 * calling thread between its narrow phase and its solver; inside fn the caller may use b2cuGetPreSolveContacts (the
 * touching contacts of this step in key order: new record + previous manifold) and b2cuDisableContacts (clears
 * e_enabledFlag until the next update: the contact is left out of islands and solver).  A non-zero return value of fn
 * aborts the step.  Costs a device round trip per step; NULL removes the hook. */
typedef int (*b2cuPreSolveFn)(void* user, b2cuWorld* w);
B2CU_API int b2cuSetPreSolveHook(b2cuWorld* w, b2cuPreSolveFn fn, void* user);
B2CU_API int b2cuGetPreSolveContacts(b2cuWorld* w, int32_t capacity, b2cuContact* records, b2cuManifold* oldManifolds,
                                     int32_t* count);
B2CU_API int b2cuDisableContacts(b2cuWorld* w, int32_t count, const b2cuContactKey* keys);

B2CU_API int b2cuStep(b2cuWorld* w, float dt, int32_t velocityIterations, int32_t positionIterations,
                      b2cuStepInfo* info);

/* Begin/End touch events of the last step as contact keys, sorted ascending (deferred-callback order). */
B2CU_API int b2cuGetEvents(b2cuWorld* w, int32_t kind, int32_t capacity, b2cuContactKey* keys, int32_t* count);

/* Begin / End touch events raised INSIDE the time-of-impact sub-steps of the last step, in the order in which the
 * reference makes those calls (b2Contact::Update called from b2World::StepSolveTOI, single-threaded: Immediate callback
 * and deferred callback back to back, Dynamics/Contacts/b2Contact.cpp:247-281), i.e. after every callback of the
 * discrete part of the step.  kinds[i] is B2CU_EVENT_BEGIN or B2CU_EVENT_END; records[i] (optional) is the contact as it
 * is at the end of the step.  b2cuGetEvents / b2cuGetEventContacts list only the events of the discrete part. */
B2CU_API int b2cuGetToiEvents(b2cuWorld* w, int32_t capacity, b2cuContactKey* keys, int32_t* kinds,
                              b2cuContact* records, int32_t* count);

/* Solver order of the last step: keys[i] is the i-th constraint, colour[i] its colour (overflow = colourCount). */
B2CU_API int b2cuGetSolverOrder(b2cuWorld* w, int32_t capacity, b2cuContactKey* keys, int32_t* colour,
                                int32_t* count);

/* Island label (smallest body index of the island) per body after the last step; -1 = not in an island. */
B2CU_API int b2cuGetIslandLabels(b2cuWorld* w, int32_t first, int32_t count, int32_t* labels);

/* TOI-eligible contacts of the current state (candidate flag, active, enabled, toiCount <= b2_maxSubSteps), key order. */
B2CU_API int b2cuGetToiCandidates(b2cuWorld* w, int32_t capacity, b2cuContactKey* keys, int32_t* count);

/* ---- spatial sharding (multi-GPU) ------------------------------------------------------------------------
 * A large world is cut into strips along x, one b2cuWorld per GPU.  Shard s holds its own bodies, every static
 * body, and GHOST copies (B2CU_BODY_GHOST) of the bodies of shard s+1 that lie within a margin of the common
 * boundary.  A contact between an own body and a ghost is owned by shard s (the lower one); ghost-ghost and
 * ghost-static pairs never become contacts.  Inside the solver every iteration runs
 *     own constraints (all colours) -> halo push owner->ghost -> cross constraints -> halo push ghost->owner
 * for velocities and then for positions, which is one sequential Gauss-Seidel order over the whole world.  The
 * pushes are stores into the neighbour's mailbox through NVLink peer memory from inside the persistent solver
 * kernel, followed by a system-scope fence and a sequence flag; there is no host round trip and no NCCL call on
 * the data path.  ghostBodies[k] of shard s and exportBodies[k] of shard s+1 are the same body. */
typedef struct b2cuShardLink
{
	unsigned char ipcHandle[64];   /* cudaIpcMemHandle_t of the mailbox allocation (other processes) */
	uint64_t localPointer;         /* the same allocation as a device pointer (same process) */
	int32_t processId;
	int32_t device;
	int32_t ghostCount, exportCount;
	int32_t rank, rankCount;
} b2cuShardLink;

B2CU_API int b2cuShardConfigure(b2cuWorld* w, int32_t rank, int32_t rankCount, int32_t ghostCount,
                                const int32_t* ghostBodies, int32_t exportCount, const int32_t* exportBodies,
                                float gridFraction /* share of the SMs the persistent solver may occupy, (0,1] */);
B2CU_API int b2cuShardGetLink(b2cuWorld* w, b2cuShardLink* link);
/* lower / upper: the links of shards rank-1 / rank+1 (NULL at the ends) */
B2CU_API int b2cuShardConnect(b2cuWorld* w, const b2cuShardLink* lower, const b2cuShardLink* upper);

/* Batched stand-alone narrow phase: manifold i = collide(shapes[shapeA[i]], xfA[i], shapes[shapeB[i]], xfB[i]).
 * xf = (p.x, p.y, sin, cos).  Shape A must be the primary type (polygon/edge before circle, edge before polygon). */
B2CU_API int b2cuCollidePairs(int32_t device, int32_t shapeCount, const b2cuShape* shapes, int32_t pairCount,
                              const int32_t* shapeA, const float* xfA, const int32_t* shapeB, const float* xfB,
                              b2cuManifold* manifolds);

/* World queries between steps, on the fat boxes the broad-phase keeps (as the reference's tree queries do):
 *   b2cuQueryAABB          b2World::QueryAABB  -> b2BroadPhase::Query    Dynamics/b2World.cpp:1752-1758
 *   b2cuRayCastCandidates  b2World::RayCast    -> b2DynamicTree::RayCast Collision/b2DynamicTree.h:203-287: the proxies
 *                          whose fat box the segment p1 -> p2 crosses (segment box overlap + the separating-axis test
 *                          of the tree walk); the exact per-shape ray cast and the callback protocol are the caller's
 * Both return ascending proxy ids; `count` receives the total even when it exceeds `capacity`. */
B2CU_API int b2cuQueryAABB(b2cuWorld* w, const float aabb[4], int32_t capacity, int32_t* proxyIds, int32_t* count);
B2CU_API int b2cuRayCastCandidates(b2cuWorld* w, const float p1[2], const float p2[2], int32_t capacity,
                                   int32_t* proxyIds, int32_t* count);

/* Batched stand-alone b2Distance (Collision/b2Distance.cpp:452-603), cold simplex cache: closest points and distance
 * of shapes[shapeA[i]] at xfA[i] and shapes[shapeB[i]] at xfB[i] (any two of circle / edge / polygon).  With useRadii
 * the result is what b2TestOverlap thresholds (distance < 10 * b2_epsilon, Collision/b2Collision.cpp:233-252). */
typedef struct b2cuDistanceResult
{
	float distance;
	float pointA[2], pointB[2];
	int32_t iterations;
} b2cuDistanceResult;
B2CU_API int b2cuDistancePairs(int32_t device, int32_t shapeCount, const b2cuShape* shapes, int32_t pairCount,
                               const int32_t* shapeA, const float* xfA, const int32_t* shapeB, const float* xfB,
                               int32_t useRadii, b2cuDistanceResult* results);

/* Batched stand-alone b2TimeOfImpact (Collision/b2TimeOfImpact.cpp:256-497), the conservative-advancement root finder
 * the continuous-collision pass (b2World::SolveTOI, Dynamics/b2World.cpp:1085-1390) is built on: first time in
 * [0, tMax] at which shapes[shapeA[i]] moving along sweepA[i] and shapes[shapeB[i]] along sweepB[i] come within the
 * target separation.  b2cuSweep restates b2Sweep (Common/b2Math.h:382-410), b2cuToiResult restates b2TOIOutput
 * (Collision/b2TimeOfImpact.h:38-52) with the same state values. */
typedef struct b2cuSweep
{
	float localCenter[2];
	float c0[2], c[2];
	float a0, a;
	float alpha0;
} b2cuSweep;
enum
{
	B2CU_TOI_UNKNOWN = 0,
	B2CU_TOI_FAILED = 1,
	B2CU_TOI_OVERLAPPED = 2,
	B2CU_TOI_TOUCHING = 3,
	B2CU_TOI_SEPARATED = 4
};
typedef struct b2cuToiResult
{
	int32_t state;
	float t;
} b2cuToiResult;
B2CU_API int b2cuTimeOfImpactPairs(int32_t device, int32_t shapeCount, const b2cuShape* shapes, int32_t pairCount,
                                   const int32_t* shapeA, const b2cuSweep* sweepA, const int32_t* shapeB,
                                   const b2cuSweep* sweepB, const float* tMax, b2cuToiResult* results);

/* sin/cos used by every transform on the device (one correctly-rounded-in-practice fp32 sincos shared with the
 * oracle so that fat-AABB decisions are reproducible); evaluated on the device. */
B2CU_API int b2cuSinCos(int32_t device, int32_t count, const float* angles, float* sinOut, float* cosOut);

#ifdef __cplusplus
}
#endif

#endif /* B2CUDA_H */
