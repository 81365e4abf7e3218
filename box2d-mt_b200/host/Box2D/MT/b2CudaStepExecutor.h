// b2CudaStepExecutor: the executor that runs b2World::Step on a B200 through the C ABI of libb2cuda.so
// (include/b2cuda.h).  Takes the place of b2ThreadPoolTaskExecutor (reference: Box2D/MT/b2ThreadPool.h:134-169)
// in user code:
//
//     b2CudaStepExecutor executor;                    // was: b2ThreadPoolTaskExecutor executor;
//     world.Step(dt, velocityIterations, positionIterations, executor);
//
// One executor can serve several worlds, one step at a time (as the reference's, b2ThreadPool.cpp:336-341).
#ifndef B2_CUDA_STEP_EXECUTOR_H
#define B2_CUDA_STEP_EXECUTOR_H

#include "Box2D/MT/b2TaskExecutor.h"

struct b2cuWorld;
struct b2cuStepInfo;
struct b2cuShardLink;
class b2Body;

struct b2CudaStepOptions
{
	b2CudaStepOptions() : device(0), downloadBodies(true), dispatchEvents(true), reportPostSolve(false), reportPreSolve(false) {}
	/// CUDA device ordinal
	int32 device;
	/// refresh the host body mirror (transform, sweep, velocities, awake) after every step; when false the
	/// mirror is refreshed lazily by the first accessor that needs it
	bool downloadBodies;
	/// fetch begin/end touch events after every step and invoke the contact listener
	bool dispatchEvents;
	/// b2ContactListener::PostSolveImmediate / PostSolve for every contact the solver handled (b2Island::Report,
	/// reference b2Island.cpp:533-570), after the device step, with the accumulated impulses of the step.  Off by
	/// default: it brings the record of every solved contact to the host each step.
	bool reportPostSolve;
	/// b2ContactListener::PreSolveImmediate / PreSolve for every touching contact, between the narrow phase and the
	/// solver of the device step (reference b2Contact.cpp:283-297, b2ContactManager.cpp:430-433), with the manifold of
	/// the previous step; b2Contact::SetEnabled(false) there keeps the contact out of this step's solve.  Off by
	/// default: one device round trip and the records of all touching contacts per step.
	bool reportPreSolve;
};

class b2CudaStepExecutor : public b2TaskExecutor
{
public:
	explicit b2CudaStepExecutor(const b2CudaStepOptions& options = b2CudaStepOptions());
	~b2CudaStepExecutor() override;

	/// user tasks run inline on the calling thread
	uint32 GetThreadCount() const override { return 1; }
	void SubmitTask(b2TaskGroup* taskGroup, b2Task* task) override;
	b2TaskGroup* AcquireTaskGroup() override { return &m_group; }

	bool StepWorld(b2World& world, float32 timeStep, int32 velocityIterations, int32 positionIterations) override;

	const b2CudaStepOptions& GetOptions() const { return m_options; }
	/// downloadBodies / dispatchEvents may be changed between steps (the device ordinal may not)
	void SetOptions(const b2CudaStepOptions& options)
	{
		m_options.downloadBodies = options.downloadBodies;
		m_options.dispatchEvents = options.dispatchEvents;
		m_options.reportPostSolve = options.reportPostSolve;
		m_options.reportPreSolve = options.reportPreSolve;
	}
	/// status of the last StepWorld (0 = ok, else a b2cuStatus) and its message
	int32 GetLastStatus() const { return m_status; }
	const char* GetLastError() const { return m_error; }
	/// counters and per-phase device timings of the last step
	const b2cuStepInfo& GetLastStepInfo() const;
	/// wall-clock milliseconds of the last StepWorld on the calling thread: [0] upload of the dirty records,
	/// [1] b2cuStep, [2] download of the body mirror, [3] event download + listener dispatch
	const float32* GetLastHostTimings() const { return m_hostMs; }
	/// the C-ABI handle of a world this executor has stepped (nullptr before its first step): for callers that
	/// want the bulk / diagnostic entry points of include/b2cuda.h
	b2cuWorld* GetDeviceWorld(b2World* world) const;
	// ---- spatial sharding of one large scene over several GPUs (include/b2cuda.h, b2cuShard*) ----
	/// This world is strip `rank` of `rankCount`.  `ghosts` are its bodies that are copies of bodies owned by strip
	/// rank+1, `exports` its own bodies that strip rank-1 holds as ghosts (same order on both sides).  Uploads the
	/// world to the device if it is not there yet.  Returns a b2cuStatus.
	int32 ConfigureShard(b2World& world, int32 rank, int32 rankCount, b2Body* const* ghosts, int32 ghostCount,
	                     b2Body* const* exports, int32 exportCount, float32 gridFraction = 1.0f);
	int32 GetShardLink(b2World& world, b2cuShardLink* link);
	int32 ConnectShard(b2World& world, const b2cuShardLink* lower, const b2cuShardLink* upper);

	/// release the device copy of a world (called by ~b2World)
	void DetachWorld(b2World* world);

private:
	b2cuWorld* EnsureDevice(b2World& world);

	b2CudaStepOptions m_options;
	float32 m_hostMs[4];
	b2TaskGroup m_group;
	int32 m_status;
	char m_error[512];
	void* m_impl;
};

#endif
