// b2cu_world.cuh -- device-resident world state (struct-of-arrays) shared by the phase kernels.
//
// Layout follows SURVEY.md 7.2: everything a phase streams is a float4/uint4 array indexed by a dense id, so
// that a warp's loads are 512-byte coalesced transactions; body arrays (16 B x N_b each) stay L2-resident
// while constraint data streams from HBM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2cuda.h"
#include "b2cu_prims.cuh"

namespace b2cu
{

#define B2CU_MAX_COLOURS 32
#define B2CU_COLOUR_NONE (-1)
#define B2CU_COLOUR_OVERFLOW B2CU_MAX_COLOURS

// device-side counters, one int each (index into DeviceArrays::counters)
enum Counter
{
	CNT_BEGIN = 0,        // begin-touch events
	CNT_END,              // end-touch events from Update
	CNT_DESTROY,          // contacts destroyed in Collide (main region)
	CNT_DESTROY_B,        // contacts destroyed in Collide (tail region)
	CNT_DESTROY_END,      // of those, touching ones (EndContact from Destroy)
	CNT_TOUCHING,         // touching contacts after Collide
	CNT_CONSTRAINT,       // contacts handed to the solver
	CNT_UNCOLOURED,       // constraints still without a colour before colouring round r: CNT_UNCOLOURED + r
	CNT_UNCOLOURED_LAST = CNT_UNCOLOURED + 7,
	CNT_OVERFLOW,         // constraints with no free colour
	CNT_MOVED,            // proxies in the move buffer
	CNT_LARGE,            // proxies too large for the grid
	CNT_LARGE_MOVED,
	CNT_NEW_PAIRS,        // new contacts found by the broad-phase
	CNT_KEEP,             // contacts surviving Collide
	CNT_ISLAND_BODIES,
	CNT_AWAKE_BODIES,
	CNT_TOI,
	CNT_ERROR,            // != 0: a buffer overflowed
	CNT_SCRATCH,
	CNT_WAKE_PATCH,       // bodies woken after the mirror copy of the step had started
	CNT_HEAVY,            // contacts queued for the dense polygon pass of Collide
	CNT_TOI_WORK,         // FindMinToiContact pass: contacts whose time of impact has to be (re)computed
	CNT_TOI_LIST,         // time-of-impact event: entries of the two bodies' contact lists
	CNT_TOI_EVENTS,       // begin / end touch events raised inside the sub-steps of this step
	CNT_FLOW_STUCK,       // != 0: a dependency wait of the dataflow solver timed out (the step fails)
	CNT_STICKY_TOI,       // not cleared per step: a TOI-candidate contact has existed
	CNT_TOI_MIN_ALPHA,    // first TOI pass: float bits of the smallest alpha (0xFFFFFFFF: no candidate)
	CNT_TOI_MIN_KEY,      // 64 bits (two slots): smallest contact key among the candidates at that alpha
	CNT_TOI_MIN_KEY_HI,
	CNT_BODY_TYPE_CHANGED, // sticky: an uploaded body row changed the body's type (the joint colouring depends on it)
	CNT_COUNT
};

struct ContactSet
{
	uint64_t* key;      // (min proxy << 32) | max proxy, ascending
	int4* proxies;      // (proxy of fixture A, proxy of fixture B) after the primary-type swap, then their bodies
	uint32_t* flags;    // B2CU_CONTACT_*
	float4* m0;         // localNormal.xy, localPoint.xy
	float4* m1;         // point0: localPoint.xy, normalImpulse, tangentImpulse
	float4* m2;         // point1
	uint4* m3;          // id0, id1, type, pointCount
	float4* mix;        // friction, restitution, tangentSpeed, toi
	int* toiCount;
	int* colour;        // colour kept from the previous step (B2CU_COLOUR_NONE if it was not a constraint)
	uint32_t* stamp;    // creation batch (b2cuContact::stamp): orders the bodies' contact lists for the TOI islands
};

// device-only proxy flag (next to the public B2CU_PROXY_* bits): moved by SyncProxiesKernel in this step
#define B2CU_PROXY_MOVED_SYNC 0x8
#define B2CU_PROXY_PUBLIC_FLAGS 0x77

static_assert(CNT_TOI_MIN_KEY % 2 == 0, "64-bit atomics on the key slot need 8-byte alignment");

// DeviceArrays::toiScratch layout (ints)
#define B2CU_TOI_SCR_SOLID 0       // 1: the event's contact was solid and an island was solved
#define B2CU_TOI_SCR_BODY_COUNT 1  // bodies of the island, then their ids
#define B2CU_TOI_SCR_BODIES 2
#define B2CU_TOI_SCRATCH_INTS 80

#define B2CU_MAX_JOINT_COLOURS 8
#define B2CU_MAX_JOINT_OPS (B2CU_MAX_JOINT_COLOURS + 1)

struct JointRow;

struct DeviceArrays
{
	// ---- bodies (index = dense body id, creation order) ----
	float4* xf;      // p.x, p.y, sin, cos
	float4* pos;     // c.x, c.y, a, -
	float4* pos0;    // c0.x, c0.y, a0, alpha0
	float4* vel;     // v.x, v.y, w, -
	float4* mass;    // invMass, invI, localCenter.x, localCenter.y
	float4* force;   // f.x, f.y, torque, sleepTime
	float4* damp;    // linearDamping, angularDamping, gravityScale, -
	uint32_t* bflags;
	int* wake;            // wake request flags written by Collide / contact creation
	int* wakePatch;       // see ApplyWakeKernel
	int* island;          // union-find parent, then island label (root = smallest body id)
	int* islandAwake;     // per root: any awake member
	int* islandMinSleep;  // per root: min sleepTime (float bits, non-negative)
	int* islandMinSep;    // [positionIterations][root]: min separation of the iteration (ordered-int float)
	uint32_t* colourMask; // per body: colours used by its constraints
	int* haloSlot;        // per body: -1, or its slot in the halo lists (| B2CU_HALO_EXPORT for an export), sharded worlds
	unsigned long long* colourClaim; // per body: (round << 32) | (~contact index), max wins

	// ---- shape geometry table ----
	b2cuShape* shapes;

	// ---- proxies (index = dense proxy id, fixture creation order) ----
	float4* fat;
	float4* fatPrev;    // fat box at the beginning of the step, valid for proxies flagged B2CU_PROXY_MOVED_SYNC
	float4* aabb;
	int* pbody;
	int* pshape;
	uint32_t* pfilter;  // categoryBits | maskBits << 16
	uint32_t* pgroup;   // (uint16)groupIndex | flags << 16
	float2* pmat;       // friction, restitution
	int* pfixture;      // caller's fixture id
	float* pradius;     // radius of the proxy's shape (copy of shapes[pshape].radius)
	int* lowStart;      // first contact whose key has this proxy as its low id (-1: none); rebuilt with the set

	// ---- contacts ----
	ContactSet c;       // live set
	ContactSet cAlt;    // rebuild target
	int* cSelect;       // per contact: 1 = solver constraint this step

	// ---- per-step lists ----
	int* listA;             // scratch index lists, contact capacity each
	int* listB;
	uint64_t* beginKeys;
	uint64_t* endKeys;
	uint64_t* destroyEndKeys; // EndContact of contacts destroyed while touching (reported after the sorted ends)
	uint64_t* newKeys;      // new pair keys (unsorted, then sorted)
	uint64_t* orderKeys;    // colour << 32 | contact index, sorted = solver order
	uint64_t* solverKeys;   // contact key of the k-th constraint of the solver order (for b2cuGetSolverOrder)
	int* listC;
	int* movedList;         // proxy capacity
	int* largeList;         // proxies larger than the coarsest grid cell
	int* levelInfo;         // per grid level: proxy count, then moved count
	int* colourCount;       // [B2CU_MAX_COLOURS + 1]
	int* colourTwoStart;    // [B2CU_MAX_COLOURS + 2]: first row of a colour whose manifold has two points (rows are ordered one-point first)
	uint64_t* toiKeys;
	// time-of-impact sub-steps (b2cu_toi_step.cuh)
	uint64_t* toiListKeys;   // contact capacity: (side, ~stamp, ~index) of the contacts on the two event bodies' lists
	uint64_t* toiListSorted; // the same in list order
	uint64_t* toiEventKeys;  // contact capacity: keys of the begin / end events of the sub-steps, in call order
	int* toiEventKinds;
	int* toiScratch;         // B2CU_TOI_SCRATCH_INTS: state handed from one kernel of an event to the next

	// ---- broad-phase grid ----
	int* cellCount;   // hash table, size gridSize (+1)
	int* cellStart;
	int* cellItems;   // proxy capacity
	float4* cellBoxes; // fat box of each entry of cellItems
	int* cellOfProxy; // proxy capacity (hash bucket or -1 for large)

	// ---- solver constraints (index = position in solver order) ----
	int4* sBody;      // bodyA, bodyB, contact index, pointCount (after block-solver demotion) | original << 8
	float4* sMass;    // mA, iA, mB, iB
	float4* sNormal;  // normal.xy, friction, tangentSpeed
	float4* sP0a;     // rA.xy, rB.xy
	float4* sP0b;     // normalMass, tangentMass, velocityBias of point 0, velocityBias of point 1
	float4* sP1a;     // second point: rA.xy, rB.xy (two-point rows only)
	float4* sImp;     // normalImpulse0, tangentImpulse0, normalImpulse1, tangentImpulse1
	float4* sLocal;   // localNormal.xy, localPoint.xy
	float4* sLocalP;  // localPoints[0].xy, localPoints[1].xy
	float4* sCenters; // localCenterA.xy, localCenterB.xy
	float4* sRadius;  // radiusA, radiusB, manifold type (int bits), island root of the constraint (int bits)

	int* counters;    // CNT_COUNT ints
	int customFilter; // != 0: a pair filter of the caller replaces the default filter rule (b2cuSetPairFilter)

	// ---- joints (index = joint id, the caller's table order) ----
	b2cuJoint* joints;       // persistent records (impulses, limit state)
	JointRow* jointRows;     // per-step rows
	int* jointOrder;         // joint ids by (colour class, id): the order of the solve
	const uint64_t* jointPairKeys; // sorted (min body << 32 | max body) of the joints that forbid collision
	const uint64_t* jointFreedKeys; // sorted body pairs that a joint used to keep apart (see JointFreed)
	int jointCount;
	int jointPairCount;
	int jointFreedCount;
};

struct WorldParams
{
	float2 gravity;
	uint32_t flags;
	float invDt0;
};

} // namespace b2cu

struct b2cuWorld
{
	int device;
	cudaStream_t stream;
	b2cu::WorldParams params;
	b2cu::DeviceArrays d;
	b2cu::PrimScratch prims;

	int bodyCapacity, proxyCapacity, shapeCapacity, contactCapacity;
	int bodyCount, proxyCount, shapeCount;
	int contactCount;    // contact slots in use: sorted main region [0, mainCount) + sorted tail [mainCount, contactCount)
	int mainCount;
	int deadMain;        // destroyed contacts still occupying slots of the main region
	int compactMin;      // smallest tail / dead-slot count that triggers a compaction
	bool compactNow;     // force the compaction of the contact set at the end of the next step
	int gridSize;        // hash table size (power of two)
	float cellSize;
	int cellChosenCount; // proxies in the world when cellSize was chosen
	bool newProxies;     // e_newFixture: run FindNewContacts at the start of the next step
	bool toiCheckDirty;  // bodies / proxies changed: re-evaluate whether TOI candidates are possible
	int positionIterationsCapacity;
	size_t l2WindowMax;  // 0: no persisting-L2 support
	bool persistentSolver; // solve all colour phases in one cooperative kernel
	int persistentGrid;
	int persistentGridMax;
	int persistentGridPosition, persistentGridPositionMax;
	int flowGrid, flowGridPosition; // co-resident grids of the dataflow solver kernels (0: unavailable)

	// spatial sharding (b2cuShardConfigure / Connect)
	int shardRank, shardCount;
	int ghostCount, exportCount;
	int* ghostIds;        // device
	int* exportIds;       // device
	unsigned char* mailbox; // device: flags + fromUpper + fromLower
	size_t mailboxBytes;
	unsigned char* peerLower; // mapped mailbox of rank-1 / rank+1 (nullptr if none)
	unsigned char* peerUpper;
	bool peerLowerIpc, peerUpperIpc;
	int peerUpperGhostCountOfUpper; // the upper neighbour's own ghostCount (layout of its mailbox)
	unsigned shardSeq;     // sequence number of the next halo exchange
	bool shardFlow;        // the shards use the dataflow solver (halo rows travel inside it) instead of the barrier kernels
	uint32_t flowEpoch;    // steps taken: the dataflow versions of a step start at (flowEpoch & 0xFFF) << 19
	int flowGridMax, flowGridPositionMax;
	int shardFlowGrid, shardFlowGridPosition;
	int flowBase;
	bool shardSoftBarrier; // a neighbouring shard lives on this device: see GridSync
	unsigned* softBarrierCounter;

	// last step
	int endUpdateCount;  // EndContact events from Update (the rest of endCount come from Destroy)
	int beginCount, endCount, constraintCount, colourCount, overflowCount, toiCount;
	int colourCounts[B2CU_MAX_COLOURS + 2]; // [32] own-class overflow, [33] cross-class overflow
	int colourStarts[B2CU_MAX_COLOURS + 3];

	int* hostCounters;   // pinned, CNT_COUNT + colour counts
	float toiMinAlpha;       // first TOI pass of the last step (b2cuStepInfo)
	uint64_t toiMinKey;
	int toiEventPending;
	int toiSubSteps, toiEventCount, toiNewContacts; // sub-steps of the last step
	bool stepComplete;       // b2World::m_stepComplete: false while a sub-stepping world is between two events
	uint32_t contactBatch;   // stamp of the next batch of new contacts
	b2cuPreSolveFn preSolveHook;
	void* preSolveUser;
	bool inPreSolve;         // b2cuStep is inside the hook: the pre-solve entry points are valid
	int preSolveCount;       // -1: the touching list has not been compacted yet in this hook
	b2cuPairFilterFn pairFilter;
	void* pairFilterUser;
	bool refilterPending;    // a proxy was uploaded with B2CU_PROXY_REFILTER
	bool contactBodiesDirty; // contacts or proxies were uploaded: refresh the body half of ContactSet::proxies
	float* bodyStage;    // device staging of b2cuBody records for b2cuGetBodies / b2cuSetBodies (lazy)
	int bodyStageCapacity;
	// body mirror (b2cuSetBodyMirror): every step copies the body records there, overlapping the broad-phase
	b2cuBodyState* bodyMirror;
	int bodyMirrorCount;
	bool mirrorInFlight;
	cudaStream_t copyStream;
	cudaEvent_t evBodiesFinal, evPacked;
	int* hostPatch;      // pinned, body capacity
	int hostPatchCapacity;
	void* queryScratch;  // device scratch of b2cuGetContactsByKey (grow-only)
	size_t queryScratchBytes;
	void* queryHost;     // page-locked host side of the same
	size_t queryHostBytes;
	bool eventPrefetch;        // b2cuSetEventPrefetch
	bool gridDirty;            // cellCount may hold counts of an interrupted grid build
	bool eventCachePending;    // the copy of the event records into queryHost was started by the step itself
	void* eventOrder;         // host scratch of b2cuGetEventContacts: 2 x eventOrderCapacity (key, index) pairs
	size_t eventOrderCapacity;
	bool eventCacheValid;      // queryHost holds the keys + records of the last step's events
	size_t eventCacheKeyBytes;
	// joints (b2cuSetJoints)
	int jointCapacity;
	uint64_t* jointPairsHost;  // host copies of the two pair lists (malloc)
	int jointPairsHostCount;
	uint64_t* jointFreedHost;
	int jointFreedHostCount, jointFreedCapacity;
	bool jointFilterPending; // the joint table changed: flag the contacts it forbids for filtering
	bool jointColourDirty;   // bodies were uploaded since the joints were coloured (a body type may have changed)
	int jointOpCount;        // colour classes of the joint order: parallel ones, then the serial overflow
	int jointOpStart[B2CU_MAX_JOINT_OPS], jointOpSize[B2CU_MAX_JOINT_OPS], jointOpSerial[B2CU_MAX_JOINT_OPS];
	int persistentGridJoints, persistentGridPositionJoints;
	cudaEvent_t ev[10];
	int launches;
	char lastError[512];
};
