// b2CudaShardedWorld: one large scene stepped on several GPUs of one box.
//
// The scene is built once, as an ordinary b2World, through the usual calls (CreateBody / CreateFixture).  The
// planner cuts it into x-strips of equal dynamic-body population; strip r is a b2World of its own that holds every
// non-dynamic body, the dynamic bodies whose x lies in [bounds[r], bounds[r+1]) and GHOST copies of the dynamic
// bodies of strip r+1 within `margin` of the common boundary (the same bodies are strip r+1's EXPORT list, in the same
// order).  Each strip is stepped by its own b2CudaStepExecutor on its own device; the boundary bodies' rows travel
// between neighbouring strips inside the solver kernels, through NVLink peer memory (include/b2cuda.h, b2cuShard*).
//
//     b2World scene(gravity);   ... CreateBody / CreateFixture ...
//     b2CudaShardedWorld sharded(scene, 4);            // 4 GPUs: devices 0..3
//     for (;;) { sharded.Step(dt, 8, 3); }             // every strip runs b2World::Step concurrently
//     sharded.Gather();                                // scene's bodies now carry the stepped state
//
// One process per GPU (torchrun / MPI style) uses the static planner functions instead: every rank computes the same
// bounds from the same scene, builds its own strip with BuildStrip and links it with ConfigureShard / GetShardLink /
// ConnectShard of b2CudaStepExecutor, exchanging the b2cuShardLink records over its own transport.
//
// Bodies migrate between strips by re-planning: Rebalance() reads the whole state back (bodies with sweeps and sleep
// timers, fat boxes, the contact set with its manifolds and accumulated impulses), cuts new strips of equal
// population at the bodies' current positions and uploads them, so the simulation goes on exactly where it was.  Call
// it every few hundred steps, or when bodies have travelled a good part of `margin` (a body that drifts further than
// `margin` past its strip's boundary between two calls is no longer seen by the neighbour).
//
// Listeners and filters are per strip (GetStrip(r).SetContactListener(...)); a strip's callbacks are made on the host
// thread that steps it (strip 0: the caller's thread), concurrently with the other strips'.  Bodies are created and
// destroyed in the scene before it is sharded, not in the strips.  Rebalance() makes new strip worlds: pointers to strip
// bodies and per-strip listeners do not survive it.
//
// Limits (the device layer's, include/b2cuda.h): no joints, time-of-impact events are refused.
#ifndef B2_CUDA_SHARDED_WORLD_H
#define B2_CUDA_SHARDED_WORLD_H

#include <vector>

#include "Box2D/Dynamics/b2World.h"
#include "Box2D/MT/b2CudaStepExecutor.h"

/// what BuildStrip reports about the strip it made
struct b2ShardStrip
{
	std::vector<b2Body*> bodies;        ///< the strip's bodies in creation order
	std::vector<int32> globalIds;       ///< index of each one in the scene (creation order there)
	std::vector<int32> proxyGlobal;     ///< scene proxy id of each of the strip's proxies
	std::vector<b2Body*> ghosts;        ///< strip bodies owned by strip rank+1
	std::vector<b2Body*> exports;       ///< strip bodies that strip rank-1 holds as ghosts
};

class b2CudaShardedWorld
{
public:
	/// Cuts `scene` into `shardCount` strips on devices[0..shardCount) (default: device r for strip r) and uploads them.
	/// Check GetLastStatus() afterwards: there is no CPU fallback.
	b2CudaShardedWorld(const b2World& scene, int32 shardCount, float32 margin = 2.0f, const int32* devices = nullptr,
	                   float32 gridFraction = 1.0f);
	~b2CudaShardedWorld();

	/// b2World::Step of the whole scene.  False (and GetLastStatus() != 0) if any strip failed.
	bool Step(float32 timeStep, int32 velocityIterations, int32 positionIterations);
	/// copies transform, velocities and the awake flag of every dynamic body from the strip that owns it into `scene`
	/// (the world given to the constructor, or any world with the same bodies in the same creation order)
	void Gather(b2World& scene) const;
	/// Re-plans the strips at the bodies' current positions (equal population again; or the given shardCount + 1
	/// boundaries) and carries the whole state over: migration of bodies and contacts between GPUs.  Between steps only.
	bool Rebalance(const float64* bounds = nullptr);
	/// Rebalance() by itself every `steps` calls of Step (0, the default: never)
	void SetRebalanceInterval(int32 steps) { m_rebalanceEvery = steps; }
	/// the executors' per-step host transport (b2CudaStepOptions::downloadBodies / dispatchEvents), all strips alike
	void SetTransport(bool downloadBodies, bool dispatchEvents);
	/// contacts Rebalance could not place because one of their bodies had left the halo of its neighbour (0 in a healthy run)
	int32 GetLostContacts() const { return m_lostContacts; }

	int32 GetShardCount() const { return (int32)m_strips.size(); }
	b2World& GetStrip(int32 rank) { return *m_worlds[rank]; }
	b2CudaStepExecutor& GetExecutor(int32 rank) { return *m_executors[rank]; }
	const b2ShardStrip& GetStripInfo(int32 rank) const { return m_strips[rank]; }
	const std::vector<float64>& GetBounds() const { return m_bounds; }
	/// rank of the strip that owns scene body `index` (creation order); -1 for non-dynamic bodies (every strip has them)
	int32 GetOwner(int32 sceneBodyIndex) const { return m_owner[sceneBodyIndex]; }
	int32 GetLastStatus() const { return m_status; }
	const char* GetLastError() const { return m_error; }

	// ---- the planner, usable on its own ----
	/// bodies of `scene` in creation order
	static void BodiesInCreationOrder(const b2World& scene, std::vector<const b2Body*>& out);
	/// strip boundaries (shardCount + 1 values, the outer two infinite): equal dynamic-body population
	static void ComputeBounds(const b2World& scene, int32 shardCount, std::vector<float64>& bounds);
	/// Fills the empty world `strip` (made with MakeStripWorld or by hand) with strip `rank` of `scene`.
	static void BuildStrip(const b2World& scene, const std::vector<float64>& bounds, int32 rank, float32 margin,
	                       b2World& strip, b2ShardStrip& info, std::vector<int32>* owner = nullptr);
	/// an empty world with the scene's gravity and flags
	static b2World* MakeStripWorld(const b2World& scene);

private:
	b2CudaShardedWorld(const b2CudaShardedWorld&);
	b2CudaShardedWorld& operator=(const b2CudaShardedWorld&);
	struct Workers;
	void Fail(int32 status, const char* what);
	int32 Link(std::vector<b2World*>& worlds, std::vector<b2CudaStepExecutor*>& executors, std::vector<b2ShardStrip>& strips);

	std::vector<b2World*> m_worlds;
	std::vector<b2CudaStepExecutor*> m_executors;
	std::vector<b2ShardStrip> m_strips;
	std::vector<float64> m_bounds;
	std::vector<int32> m_owner;
	std::vector<int32> m_localIndex; // of a scene body in its owning strip's `bodies`
	std::vector<int32> m_devices;
	float32 m_margin, m_gridFraction;
	int32 m_lostContacts;
	int32 m_rebalanceEvery, m_stepsSinceRebalance;
	bool m_downloadBodies, m_dispatchEvents;
	Workers* m_workers;
	int32 m_status;
	char m_error[512];
};

#endif
