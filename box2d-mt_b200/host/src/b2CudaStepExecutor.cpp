// b2CudaStepExecutor: runs b2World::Step on the device through the C ABI (include/b2cuda.h).
#include "Box2D/MT/b2CudaStepExecutor.h"

#include <chrono>
#include "Box2D/Dynamics/b2World.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

namespace
{
struct Impl
{
	std::map<b2World*, b2cuWorld*> worlds;
	b2cuStepInfo info;
};
} // namespace

b2CudaStepExecutor::b2CudaStepExecutor(const b2CudaStepOptions& options) : m_options(options), m_hostMs(), m_status(0)
{
	m_error[0] = 0;
	Impl* impl = new Impl;
	memset(&impl->info, 0, sizeof(impl->info));
	m_impl = impl;
}

b2CudaStepExecutor::~b2CudaStepExecutor()
{
	Impl* impl = static_cast<Impl*>(m_impl);
	for (auto it = impl->worlds.begin(); it != impl->worlds.end(); ++it)
	{
		// the world keeps its host mirror; a later step with another executor re-uploads everything
		it->first->RefreshBodies();
		it->first->RefreshSweepStarts();
		it->first->RefreshProxies();
		it->first->RefreshJoints();
		it->first->m_contactsStale = true;
		it->first->RefreshContacts();
		it->first->m_contactsStale = false;
		it->first->m_fullUpload = true;
		it->first->m_device = nullptr;
		it->first->m_owner = nullptr;
		b2cuDestroyWorld(it->second);
	}
	delete impl;
}

b2cuWorld* b2CudaStepExecutor::GetDeviceWorld(b2World* world) const
{
	Impl* impl = static_cast<Impl*>(m_impl);
	auto it = impl->worlds.find(world);
	return it == impl->worlds.end() ? nullptr : it->second;
}

void b2CudaStepExecutor::DetachWorld(b2World* world)
{
	Impl* impl = static_cast<Impl*>(m_impl);
	auto it = impl->worlds.find(world);
	if (it == impl->worlds.end()) return;
	b2cuDestroyWorld(it->second);
	impl->worlds.erase(it);
}

void b2CudaStepExecutor::SubmitTask(b2TaskGroup* taskGroup, b2Task* task)
{
	B2_NOT_USED(taskGroup);
	b2ThreadContext ctx;
	ctx.stack = nullptr;
	ctx.threadId = 0;
	task->Execute(ctx);
}

const b2cuStepInfo& b2CudaStepExecutor::GetLastStepInfo() const { return static_cast<Impl*>(m_impl)->info; }

b2cuWorld* b2CudaStepExecutor::EnsureDevice(b2World& world)
{
	Impl* impl = static_cast<Impl*>(m_impl);
	auto it = impl->worlds.find(&world);
	if (it != impl->worlds.end()) return it->second;
	b2cuWorld* device = nullptr;
	b2cuWorldDef def;
	memset(&def, 0, sizeof(def));
	def.device = m_options.device;
	def.gravity[0] = world.m_gravity.x;
	def.gravity[1] = world.m_gravity.y;
	def.bodyCapacity = (int32)world.m_states.size();
	def.proxyCapacity = (int32)world.m_proxies.size();
	def.shapeCapacity = (int32)world.m_shapes.size();
	def.contactCapacity = 8 * (int32)world.m_proxies.size();
	int rc = b2cuCreateWorld(&def, &device);
	if (rc != B2CU_OK)
	{
		m_status = rc;
		snprintf(m_error, sizeof(m_error), "b2cuCreateWorld failed with status %d (no CUDA device?)", rc);
		world.m_lastStatus = rc;
		return nullptr;
	}
	impl->worlds[&world] = device;
	world.m_device = device;
	world.m_owner = this;
	if (world.m_bodiesUploaded > 0) world.m_fullUpload = true;
	return device;
}

int32 b2CudaStepExecutor::ConfigureShard(b2World& world, int32 rank, int32 rankCount, b2Body* const* ghosts,
                                         int32 ghostCount, b2Body* const* exports, int32 exportCount, float32 gridFraction)
{
	std::vector<int32> g(ghostCount), e(exportCount);
	for (int32 i = 0; i < ghostCount; ++i)
	{
		g[i] = ghosts[i]->GetIndex();
		world.m_states[g[i]].flags |= B2CU_BODY_GHOST;
		world.MarkBodyDirty(g[i]);
	}
	for (int32 i = 0; i < exportCount; ++i) e[i] = exports[i]->GetIndex();
	b2cuWorld* device = EnsureDevice(world);
	if (device == nullptr) return m_status;
	int rc = world.UploadDirty(device);
	if (rc == B2CU_OK)
		rc = b2cuShardConfigure(device, rank, rankCount, ghostCount, g.data(), exportCount, e.data(), gridFraction);
	if (rc != B2CU_OK) snprintf(m_error, sizeof(m_error), "%s", b2cuGetLastError(device));
	return m_status = rc;
}

int32 b2CudaStepExecutor::GetShardLink(b2World& world, b2cuShardLink* link)
{
	b2cuWorld* device = GetDeviceWorld(&world);
	if (device == nullptr) return B2CU_ERR_ARGUMENT;
	return b2cuShardGetLink(device, link);
}

int32 b2CudaStepExecutor::ConnectShard(b2World& world, const b2cuShardLink* lower, const b2cuShardLink* upper)
{
	b2cuWorld* device = GetDeviceWorld(&world);
	if (device == nullptr) return B2CU_ERR_ARGUMENT;
	int rc = b2cuShardConnect(device, lower, upper);
	if (rc != B2CU_OK) snprintf(m_error, sizeof(m_error), "%s", b2cuGetLastError(device));
	return rc;
}

bool b2CudaStepExecutor::StepWorld(b2World& world, float32 timeStep, int32 velocityIterations, int32 positionIterations)
{
	Impl* impl = static_cast<Impl*>(m_impl);
	m_status = 0;
	m_error[0] = 0;

	if (world.m_owner != nullptr && world.m_owner != this)
	{
		// last stepped by another executor: take the world over (its state comes back through the host mirror)
		world.RefreshBodies();
		world.RefreshSweepStarts();
		world.RefreshProxies();
		world.RefreshJoints();
		world.m_contactsStale = true;
		world.RefreshContacts();
		world.m_contactsStale = false;
		world.m_owner->DetachWorld(&world);
		world.m_device = nullptr;
		world.m_fullUpload = true;
	}

	b2cuWorld* device = EnsureDevice(world);
	if (device == nullptr) return false;

	typedef std::chrono::steady_clock Clock;
	Clock::time_point t0 = Clock::now();
	int rc = world.UploadDirty(device);
	Clock::time_point t1 = Clock::now();
	// with downloadBodies the step itself copies the body records into the host mirror (overlapped with its
	// broad-phase part); the mirror is current when b2cuStep returns
	if (rc == B2CU_OK)
	{
		if (m_options.downloadBodies && !world.m_states.empty())
			rc = b2cuSetBodyMirror(device, world.m_states.data(), (int32)world.m_states.size());
		else
			rc = b2cuSetBodyMirror(device, nullptr, 0);
	}
	// a user b2ContactFilter is consulted for the new pairs of the step, on this thread (b2cuSetPairFilter)
	if (rc == B2CU_OK)
		rc = b2cuSetPairFilter(device, world.m_contactFilter ? &b2World::PairFilterThunk : nullptr, &world);
	if (rc == B2CU_OK)
		rc = b2cuSetPreSolveHook(device, (m_options.reportPreSolve && world.m_contactListener) ? &b2World::PreSolveThunk : nullptr,
		                         &world);
	// a listener will be handed the contacts of the step's events: let the step bring their records along
	if (rc == B2CU_OK) rc = b2cuSetEventPrefetch(device, (m_options.dispatchEvents && world.m_contactListener) ? 1 : 0);
	if (rc == B2CU_OK) rc = b2cuStep(device, timeStep, velocityIterations, positionIterations, &impl->info);
	Clock::time_point t2 = Clock::now();
	m_hostMs[0] = std::chrono::duration<float, std::milli>(t1 - t0).count();
	m_hostMs[1] = std::chrono::duration<float, std::milli>(t2 - t1).count();
	m_hostMs[2] = m_hostMs[3] = 0.0f;
	if (rc != B2CU_OK)
	{
		m_status = rc;
		snprintf(m_error, sizeof(m_error), "%s", b2cuGetLastError(device));
		world.m_lastStatus = rc;
		fprintf(stderr, "b2CudaStepExecutor: step failed (%d): %s\n", rc, m_error);
		return false;
	}
	if (timeStep > 0.0f) world.m_inv_dt0 = 1.0f / timeStep;
	world.m_lastStatus = 0;
	world.AfterDeviceStep(device, impl->info, m_options.downloadBodies, m_options.dispatchEvents, m_hostMs + 2,
	                      m_options.reportPostSolve);
	return true;
}
