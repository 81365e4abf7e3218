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
	virtual ~b2TaskExecutor() {}

	/// number of threads that can execute tasks; between 1 and b2_maxThreads
	virtual uint32 GetThreadCount() const = 0;

	virtual void SubmitTask(b2TaskGroup* taskGroup, b2Task* task)
	{
		B2_NOT_USED(taskGroup);
		B2_NOT_USED(task);
	}

	virtual void Wait(b2TaskGroup* taskGroup, const b2ThreadContext& ctx)
	{
		B2_NOT_USED(taskGroup);
		B2_NOT_USED(ctx);
	}

	virtual void SubmitTasks(b2TaskGroup* taskGroup, b2Task** tasks, uint32 count)
	{
		for (uint32 i = 0; i < count; ++i) SubmitTask(taskGroup, tasks[i]);
	}

	virtual b2TaskGroup* AcquireTaskGroup() { return nullptr; }

	virtual void ReleaseTaskGroup(b2TaskGroup* taskGroup) { B2_NOT_USED(taskGroup); }

	virtual void PartitionRange(b2Task::Type type, uint32 begin, uint32 end, b2PartitionedRange& output)
	{
		B2_NOT_USED(type);
		output.ranges[0].begin = begin;
		output.ranges[0].end = end;
		output.count = 1;
	}

	/// Run one whole world step.  Return false if this executor cannot.
	virtual bool StepWorld(b2World& world, float32 timeStep, int32 velocityIterations, int32 positionIterations)
	{
		B2_NOT_USED(world);
		B2_NOT_USED(timeStep);
		B2_NOT_USED(velocityIterations);
		B2_NOT_USED(positionIterations);
		return false;
	}
};

#endif
