// The executor plug-in point (reference: Box2D/MT/b2TaskExecutor.h:27-79), verbatim in its seven task methods,
// plus ONE addition: StepWorld.  The reference's step builds task closures over host pointers
// (Box2D/Dynamics/b2World.cpp:44-360), which a GPU cannot run, so b2World::Step hands the whole step to the
// executor; an executor that cannot run it returns false and Step fails loudly (there is no CPU step path in
// this library).
#ifndef B2_TASK_EXECUTOR_H
#define B2_TASK_EXECUTOR_H

#include "Box2D/MT/b2Task.h"

class b2World;

class b2TaskExecutor
{
public:
	/// THE ADDITION: run one whole world step.  An executor that cannot returns false (the default).
	virtual bool StepWorld(b2World& /*world*/, float32 /*timeStep*/, int32 /*velocityIterations*/, int32 /*positionIterations*/)
	{
		return false;
	}

	/// how many threads execute tasks, 1 .. b2_maxThreads
	virtual uint32 GetThreadCount() const = 0;

	// ---- task submission, as in the reference; the defaults describe a single-threaded executor ----
	virtual void SubmitTask(b2TaskGroup* /*taskGroup*/, b2Task* /*task*/) {}
	virtual void SubmitTasks(b2TaskGroup* taskGroup, b2Task** tasks, uint32 count)
	{
		for (uint32 k = 0; k != count; ++k) SubmitTask(taskGroup, tasks[k]);
	}
	virtual void Wait(b2TaskGroup* /*taskGroup*/, const b2ThreadContext& /*ctx*/) {}
	virtual b2TaskGroup* AcquireTaskGroup() { return nullptr; }
	virtual void ReleaseTaskGroup(b2TaskGroup* /*taskGroup*/) {}

	/// split [begin, end) into sub-ranges for range tasks; the default keeps it whole
	virtual void PartitionRange(b2Task::Type /*type*/, uint32 begin, uint32 end, b2PartitionedRange& output)
	{
		output.count = 1;
		output.ranges[0].begin = begin;
		output.ranges[0].end = end;
	}

	virtual ~b2TaskExecutor() {}
};

#endif
