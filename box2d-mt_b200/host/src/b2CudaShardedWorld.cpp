// b2CudaShardedWorld: strip planner + concurrent stepping of the strips (see the header).  The planner follows the
// same rules as python/b2shard.py (which plans from flat scene arrays) so that both produce the same strips.
#include "Box2D/MT/b2CudaShardedWorld.h"

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>

#include "b2cuda.h"

// ---------------------------------------------------------------------------------------------------------
// planner
// ---------------------------------------------------------------------------------------------------------

void b2CudaShardedWorld::BodiesInCreationOrder(const b2World& scene, std::vector<const b2Body*>& out)
{
	out.assign((size_t)scene.GetBodyCount(), nullptr);
	// the body list links new bodies at the head (as the reference's, b2World.cpp:120-127): walk it backwards
	size_t k = out.size();
	for (const b2Body* b = scene.GetBodyList(); b != nullptr && k > 0; b = b->GetNext()) out[--k] = b;
}

void b2CudaShardedWorld::ComputeBounds(const b2World& scene, int32 shardCount, std::vector<float64>& bounds)
{
	std::vector<float64> xs;
	for (const b2Body* b = scene.GetBodyList(); b != nullptr; b = b->GetNext())
		if (b->GetType() == b2_dynamicBody) xs.push_back((float64)b->GetPosition().x);
	std::sort(xs.begin(), xs.end());
	bounds.assign((size_t)shardCount + 1, 0.0);
	bounds[0] = -std::numeric_limits<float64>::infinity();
	bounds[(size_t)shardCount] = std::numeric_limits<float64>::infinity();
	for (int32 r = 1; r < shardCount; ++r)
		bounds[(size_t)r] = xs.empty() ? 0.0 : xs[(xs.size() * (size_t)r) / (size_t)shardCount];
}

b2World* b2CudaShardedWorld::MakeStripWorld(const b2World& scene)
{
	b2World* w = new b2World(scene.GetGravity());
	w->SetAllowSleeping(scene.GetAllowSleeping());
	w->SetWarmStarting(scene.GetWarmStarting());
	w->SetContinuousPhysics(scene.GetContinuousPhysics());
	w->SetSubStepping(scene.GetSubStepping());
	w->SetAutoClearForces(scene.GetAutoClearForces());
	return w;
}

namespace
{

// strip that owns x: bounds[r] <= x < bounds[r+1]
int32 OwnerOf(const std::vector<float64>& bounds, float64 x)
{
	const int32 count = (int32)bounds.size() - 1;
	int32 r = (int32)(std::upper_bound(bounds.begin(), bounds.end(), x) - bounds.begin()) - 1;
	return std::min(std::max(r, 0), count - 1);
}

// proxyGlobal (optional) receives, for every proxy the copy gets, the id of the source's proxy mapped through
// sourceProxyGlobal (nullptr: the source's own proxy ids)
b2Body* CloneBody(const b2Body* src, b2World& into, std::vector<int32>* proxyGlobal = nullptr,
                  const std::vector<int32>* sourceProxyGlobal = nullptr)
{
	b2BodyDef bd;
	bd.type = src->GetType();
	bd.position = src->GetPosition();
	bd.angle = src->GetAngle();
	bd.linearVelocity = src->GetLinearVelocity();
	bd.angularVelocity = src->GetAngularVelocity();
	bd.linearDamping = src->GetLinearDamping();
	bd.angularDamping = src->GetAngularDamping();
	bd.gravityScale = src->GetGravityScale();
	bd.allowSleep = src->IsSleepingAllowed();
	bd.awake = src->IsAwake();
	bd.fixedRotation = src->IsFixedRotation();
	bd.bullet = src->IsBullet();
	bd.active = src->IsActive();
	bd.userData = src->GetUserData();
	b2Body* body = into.CreateBody(&bd);
	// the fixture list links new fixtures at the head (b2Body.cpp:166-215 of the reference): recreate oldest first
	std::vector<const b2Fixture*> fixtures;
	for (const b2Fixture* f = src->GetFixtureList(); f != nullptr; f = f->GetNext()) fixtures.push_back(f);
	for (size_t k = fixtures.size(); k-- > 0;)
	{
		const b2Fixture* f = fixtures[k];
		b2FixtureDef fd;
		fd.shape = f->GetShape();
		fd.userData = f->GetUserData();
		fd.friction = f->GetFriction();
		fd.restitution = f->GetRestitution();
		fd.density = f->GetDensity();
		fd.isSensor = f->IsSensor();
		fd.thickShape = f->IsThickShape();
		fd.filter = f->GetFilterData();
		body->CreateFixture(&fd);
		if (proxyGlobal)
			for (int32 c = 0; c < f->GetProxyCount(); ++c)
			{
				const int32 sp = f->GetProxyIndex() + c;
				proxyGlobal->push_back(sourceProxyGlobal ? (*sourceProxyGlobal)[(size_t)sp] : sp);
			}
	}
	return body;
}

} // namespace

void b2CudaShardedWorld::BuildStrip(const b2World& scene, const std::vector<float64>& bounds, int32 rank, float32 margin,
                                    b2World& strip, b2ShardStrip& info, std::vector<int32>* ownerOut)
{
	const int32 count = (int32)bounds.size() - 1;
	std::vector<const b2Body*> all;
	BodiesInCreationOrder(scene, all);
	info.bodies.clear();
	info.globalIds.clear();
	info.proxyGlobal.clear();
	info.ghosts.clear();
	info.exports.clear();
	if (ownerOut) ownerOut->assign(all.size(), -1);
	// local bodies keep the scene's relative order, so that the fixture A / fixture B roles of a contact (lower proxy
	// id first, b2ContactManager::AddPair) are the same in the strip and in the whole world
	for (size_t i = 0; i < all.size(); ++i)
	{
		const b2Body* src = all[i];
		const bool dynamic = src->GetType() == b2_dynamicBody;
		const float64 x = (float64)src->GetPosition().x;
		const int32 owner = dynamic ? OwnerOf(bounds, x) : -1;
		if (ownerOut) (*ownerOut)[i] = owner;
		const bool own = !dynamic || owner == rank;
		const bool ghost = dynamic && rank + 1 < count && owner == rank + 1 && x < bounds[(size_t)rank + 1] + (float64)margin;
		if (!own && !ghost) continue;
		b2Body* body = CloneBody(src, strip, &info.proxyGlobal);
		info.bodies.push_back(body);
		info.globalIds.push_back((int32)i);
		if (ghost) info.ghosts.push_back(body);
		if (dynamic && own && rank > 0 && x < bounds[(size_t)rank] + (float64)margin) info.exports.push_back(body);
	}
}

// ---------------------------------------------------------------------------------------------------------
// one host thread per strip: the strips of one step must be in flight together (their solver kernels wait for
// each other's boundary rows)
// ---------------------------------------------------------------------------------------------------------

struct b2CudaShardedWorld::Workers
{
	std::mutex mutex;
	std::condition_variable wake, done;
	std::vector<std::thread> threads;
	uint64 generation = 0;
	int32 pending = 0;
	bool quit = false;
	float32 dt = 0.0f;
	int32 velocityIterations = 0, positionIterations = 0;
	std::vector<int32> status;
};

b2CudaShardedWorld::b2CudaShardedWorld(const b2World& scene, int32 shardCount, float32 margin, const int32* devices,
                                       float32 gridFraction)
	: m_margin(margin), m_gridFraction(gridFraction), m_lostContacts(0), m_rebalanceEvery(0), m_stepsSinceRebalance(0),
	  m_downloadBodies(true), m_dispatchEvents(true), m_workers(nullptr), m_status(0)
{
	m_error[0] = 0;
	if (shardCount < 1)
	{
		Fail(B2CU_ERR_ARGUMENT, "shardCount < 1");
		return;
	}
	if (scene.GetJointCount() > 0)
	{
		Fail(B2CU_ERR_UNSUPPORTED, "joints in a sharded world");
		return;
	}
	for (int32 r = 0; r < shardCount; ++r) m_devices.push_back(devices ? devices[r] : r);
	ComputeBounds(scene, shardCount, m_bounds);
	m_strips.resize((size_t)shardCount);
	for (int32 r = 0; r < shardCount; ++r)
	{
		b2CudaStepOptions opt;
		opt.device = m_devices[(size_t)r];
		m_worlds.push_back(MakeStripWorld(scene));
		m_executors.push_back(new b2CudaStepExecutor(opt));
		BuildStrip(scene, m_bounds, r, margin, *m_worlds[(size_t)r], m_strips[(size_t)r], r == 0 ? &m_owner : nullptr);
	}
	m_localIndex.assign(m_owner.size(), -1);
	for (int32 r = 0; r < shardCount; ++r)
	{
		const b2ShardStrip& s = m_strips[(size_t)r];
		for (size_t k = 0; k < s.globalIds.size(); ++k)
		{
			const size_t g = (size_t)s.globalIds[k];
			if (m_owner[g] == r || (m_owner[g] < 0 && r == 0)) m_localIndex[g] = (int32)k;
		}
	}
	if (Link(m_worlds, m_executors, m_strips) != B2CU_OK) return;
	// strip 0 is stepped by the calling thread
	m_workers = new Workers;
	m_workers->status.assign((size_t)shardCount, 0);
	for (int32 r = 1; r < shardCount; ++r)
	{
		m_workers->threads.push_back(std::thread([this, r]() {
			Workers& k = *m_workers;
			uint64 seen = 0;
			for (;;)
			{
				float32 dt;
				int32 vi, pi;
				{
					std::unique_lock<std::mutex> lock(k.mutex);
					k.wake.wait(lock, [&]() { return k.quit || k.generation != seen; });
					if (k.quit) return;
					seen = k.generation;
					dt = k.dt;
					vi = k.velocityIterations;
					pi = k.positionIterations;
				}
				m_worlds[(size_t)r]->Step(dt, vi, pi, *m_executors[(size_t)r]);
				{
					std::lock_guard<std::mutex> lock(k.mutex);
					k.status[(size_t)r] = m_worlds[(size_t)r]->GetLastStepStatus();
					if (--k.pending == 0) k.done.notify_one();
				}
			}
		}));
	}
}

// upload + b2cuShardConfigure of every strip, then the links between neighbours
int32 b2CudaShardedWorld::Link(std::vector<b2World*>& worlds, std::vector<b2CudaStepExecutor*>& executors,
                               std::vector<b2ShardStrip>& strips)
{
	const int32 shardCount = (int32)worlds.size();
	for (int32 r = 0; r + 1 < shardCount; ++r)
		if (strips[(size_t)r].ghosts.size() != strips[(size_t)r + 1].exports.size())
		{
			Fail(B2CU_ERR_ARGUMENT, "ghost / export lists of neighbouring strips differ in length");
			return m_status;
		}
	if (shardCount < 2) return B2CU_OK;
	std::vector<b2cuShardLink> links((size_t)shardCount);
	for (int32 r = 0; r < shardCount; ++r)
	{
		b2ShardStrip& s = strips[(size_t)r];
		int32 rc = executors[(size_t)r]->ConfigureShard(*worlds[(size_t)r], r, shardCount, s.ghosts.data(), (int32)s.ghosts.size(),
		                                                s.exports.data(), (int32)s.exports.size(), m_gridFraction);
		if (rc == B2CU_OK) rc = executors[(size_t)r]->GetShardLink(*worlds[(size_t)r], &links[(size_t)r]);
		if (rc != B2CU_OK)
		{
			Fail(rc, executors[(size_t)r]->GetLastError());
			return rc;
		}
	}
	for (int32 r = 0; r < shardCount; ++r)
	{
		int32 rc = executors[(size_t)r]->ConnectShard(*worlds[(size_t)r], r > 0 ? &links[(size_t)r - 1] : nullptr,
		                                              r + 1 < shardCount ? &links[(size_t)r + 1] : nullptr);
		if (rc != B2CU_OK)
		{
			Fail(rc, executors[(size_t)r]->GetLastError());
			return rc;
		}
	}
	return B2CU_OK;
}

b2CudaShardedWorld::~b2CudaShardedWorld()
{
	if (m_workers)
	{
		{
			std::lock_guard<std::mutex> lock(m_workers->mutex);
			m_workers->quit = true;
		}
		m_workers->wake.notify_all();
		for (size_t i = 0; i < m_workers->threads.size(); ++i) m_workers->threads[i].join();
		delete m_workers;
	}
	for (size_t r = 0; r < m_worlds.size(); ++r) delete m_worlds[r];
	for (size_t r = 0; r < m_executors.size(); ++r) delete m_executors[r];
}

void b2CudaShardedWorld::Fail(int32 status, const char* what)
{
	m_status = status;
	snprintf(m_error, sizeof(m_error), "%s", what ? what : "");
}

void b2CudaShardedWorld::SetTransport(bool downloadBodies, bool dispatchEvents)
{
	m_downloadBodies = downloadBodies;
	m_dispatchEvents = dispatchEvents;
	for (size_t r = 0; r < m_executors.size(); ++r)
	{
		b2CudaStepOptions opt = m_executors[r]->GetOptions();
		opt.downloadBodies = downloadBodies;
		opt.dispatchEvents = dispatchEvents;
		m_executors[r]->SetOptions(opt);
	}
}

bool b2CudaShardedWorld::Step(float32 timeStep, int32 velocityIterations, int32 positionIterations)
{
	if (m_workers == nullptr) return false;
	if (m_rebalanceEvery > 0 && m_stepsSinceRebalance >= m_rebalanceEvery && !Rebalance()) return false;
	++m_stepsSinceRebalance;
	Workers& k = *m_workers;
	const int32 n = (int32)m_worlds.size();
	{
		std::lock_guard<std::mutex> lock(k.mutex);
		k.dt = timeStep;
		k.velocityIterations = velocityIterations;
		k.positionIterations = positionIterations;
		k.pending = n - 1;
		++k.generation;
	}
	k.wake.notify_all();
	m_worlds[0]->Step(timeStep, velocityIterations, positionIterations, *m_executors[0]);
	k.status[0] = m_worlds[0]->GetLastStepStatus();
	{
		std::unique_lock<std::mutex> lock(k.mutex);
		k.done.wait(lock, [&]() { return k.pending == 0; });
	}
	m_status = 0;
	for (int32 r = 0; r < n; ++r)
		if (k.status[(size_t)r] != 0 && m_status == 0)
		{
			m_status = k.status[(size_t)r];
			snprintf(m_error, sizeof(m_error), "strip %d: %.480s", r, m_executors[(size_t)r]->GetLastError());
		}
	return m_status == 0;
}

void b2CudaShardedWorld::Gather(b2World& scene) const
{
	std::vector<b2Body*> all((size_t)scene.GetBodyCount(), nullptr);
	size_t k = all.size();
	for (b2Body* b = scene.GetBodyList(); b != nullptr && k > 0; b = b->GetNext()) all[--k] = b;
	for (size_t i = 0; i < all.size() && i < m_owner.size(); ++i)
	{
		b2Body* to = all[i];
		if (to->GetType() == b2_staticBody || m_localIndex[i] < 0) continue;
		// kinematic bodies live in every strip and move alike: strip 0's copy speaks for them
		const b2Body* from = m_strips[(size_t)std::max(m_owner[i], 0)].bodies[(size_t)m_localIndex[i]];
		to->SetTransform(from->GetPosition(), from->GetAngle());
		to->SetLinearVelocity(from->GetLinearVelocity());
		to->SetAngularVelocity(from->GetAngularVelocity());
		to->SetAwake(from->IsAwake());
	}
}

// ---------------------------------------------------------------------------------------------------------
// Migration by re-planning.  Everything a strip's device world carries from one step to the next
// is read back into the host mirrors of the old strips, new strips are cut at the bodies' current positions, and the
// rows are copied across: body records with sweep starts and sleep timers, fat boxes, and the contact set -- every
// pair of the whole scene lives in exactly one strip (the strip that owns one of its bodies; the lower one for a pair
// across a boundary), so the union of the strips' contact sets is the scene's, and each contact moves to the strip
// that holds it under the new plan.
// ---------------------------------------------------------------------------------------------------------
bool b2CudaShardedWorld::Rebalance(const float64* newBounds)
{
	if (m_workers == nullptr) return false;
	const int32 n = (int32)m_worlds.size();
	const size_t bodyTotal = m_owner.size();
	for (int32 r = 0; r < n; ++r)
	{
		b2World& w = *m_worlds[(size_t)r];
		w.RefreshBodies();
		w.RefreshSweepStarts();
		w.RefreshProxies();
		w.m_contactsStale = true;
		w.RefreshContacts();
	}
	// where every scene body is now, by its owner's copy (non-dynamic bodies: strip 0's copy; all strips move them alike)
	std::vector<const b2Body*> current(bodyTotal, nullptr);
	std::vector<int32> holder(bodyTotal, 0);
	for (size_t g = 0; g < bodyTotal; ++g)
	{
		holder[g] = std::max(m_owner[g], 0);
		current[g] = m_strips[(size_t)holder[g]].bodies[(size_t)m_localIndex[g]];
	}
	std::vector<float64> bounds((size_t)n + 1);
	if (newBounds) bounds.assign(newBounds, newBounds + n + 1);
	else
	{
		std::vector<float64> xs;
		for (size_t g = 0; g < bodyTotal; ++g)
			if (current[g]->GetType() == b2_dynamicBody) xs.push_back((float64)current[g]->GetPosition().x);
		std::sort(xs.begin(), xs.end());
		bounds[0] = -std::numeric_limits<float64>::infinity();
		bounds[(size_t)n] = std::numeric_limits<float64>::infinity();
		for (int32 r = 1; r < n; ++r) bounds[(size_t)r] = xs.empty() ? 0.0 : xs[(xs.size() * (size_t)r) / (size_t)n];
	}
	std::vector<int32> owner(bodyTotal, -1);
	for (size_t g = 0; g < bodyTotal; ++g)
		if (current[g]->GetType() == b2_dynamicBody) owner[g] = OwnerOf(bounds, (float64)current[g]->GetPosition().x);

	size_t proxyTotal = 0;
	for (int32 r = 0; r < n; ++r)
		for (size_t k = 0; k < m_strips[(size_t)r].proxyGlobal.size(); ++k)
			proxyTotal = std::max(proxyTotal, (size_t)m_strips[(size_t)r].proxyGlobal[k] + 1);

	std::vector<b2World*> worlds;
	std::vector<b2CudaStepExecutor*> executors;
	std::vector<b2ShardStrip> strips((size_t)n);
	std::vector<int32> localIndex(bodyTotal, -1);
	std::vector<std::vector<int32> > proxyLocal((size_t)n); // scene proxy id -> strip proxy id
	std::vector<int32> proxyBody(proxyTotal, -1);           // scene proxy id -> scene body id
	for (int32 r = 0; r < n; ++r)
	{
		b2CudaStepOptions opt;
		opt.device = m_devices[(size_t)r];
		opt.downloadBodies = m_downloadBodies;
		opt.dispatchEvents = m_dispatchEvents;
		b2World* nw = MakeStripWorld(*m_worlds[0]);
		worlds.push_back(nw);
		executors.push_back(new b2CudaStepExecutor(opt));
		b2ShardStrip& info = strips[(size_t)r];
		for (size_t g = 0; g < bodyTotal; ++g)
		{
			const b2Body* src = current[g];
			const bool dynamic = owner[g] >= 0;
			const float64 x = (float64)src->GetPosition().x;
			const bool own = !dynamic || owner[g] == r;
			const bool ghost = dynamic && r + 1 < n && owner[g] == r + 1 && x < bounds[(size_t)r + 1] + (float64)m_margin;
			if (!own && !ghost) continue;
			const b2World& ow = *m_worlds[(size_t)holder[g]];
			const size_t firstProxy = info.proxyGlobal.size();
			b2Body* body = CloneBody(src, *nw, &info.proxyGlobal, &m_strips[(size_t)holder[g]].proxyGlobal);
			// the rows as they are, not as CreateBody / CreateFixture derive them from a definition
			const size_t i = (size_t)src->m_index, j = (size_t)body->m_index;
			nw->m_states[j] = ow.m_states[i];
			nw->m_states[j].flags &= ~(uint32)B2CU_BODY_GHOST;
			nw->m_sweepStarts[j] = ow.m_sweepStarts[i];
			nw->m_props[j] = ow.m_props[i];
			body->m_mass = src->m_mass;
			body->m_I = src->m_I;
			if (ow.m_props[i].fx != 0.0f || ow.m_props[i].fy != 0.0f || ow.m_props[i].torque != 0.0f) nw->m_forced.push_back((int32)j);
			{
				// proxies in creation order on both sides
				std::vector<const b2Fixture*> of;
				for (const b2Fixture* f = src->GetFixtureList(); f != nullptr; f = f->GetNext()) of.push_back(f);
				size_t np = firstProxy;
				for (size_t k = of.size(); k-- > 0;)
					for (int32 c = 0; c < of[k]->m_proxyCount; ++c, ++np)
					{
						const b2cuProxy& from = ow.m_proxies[(size_t)(of[k]->m_proxyIndex + c)];
						b2cuProxy& to = nw->m_proxies[np];
						memcpy(to.aabb, from.aabb, sizeof(to.aabb));
						memcpy(to.fat, from.fat, sizeof(to.fat));
						to.flags = from.flags;
					}
			}
			localIndex[g] = (own && (dynamic || r == 0)) ? (int32)info.bodies.size() : localIndex[g];
			info.bodies.push_back(body);
			info.globalIds.push_back((int32)g);
			if (ghost) info.ghosts.push_back(body);
			if (dynamic && own && r > 0 && x < bounds[(size_t)r] + (float64)m_margin) info.exports.push_back(body);
			for (size_t q = firstProxy; q < info.proxyGlobal.size(); ++q) proxyBody[(size_t)info.proxyGlobal[q]] = (int32)g;
		}
		proxyLocal[(size_t)r].assign(proxyTotal, -1);
		for (size_t q = 0; q < info.proxyGlobal.size(); ++q) proxyLocal[(size_t)r][(size_t)info.proxyGlobal[q]] = (int32)q;
		nw->m_inv_dt0 = m_worlds[0]->m_inv_dt0;
		nw->m_contactRecords.clear();
		nw->m_fullUpload = true; // the contact set travels with the first upload
	}
	// contacts
	int32 lost = 0;
	for (int32 r = 0; r < n; ++r)
	{
		const b2World& ow = *m_worlds[(size_t)r];
		const std::vector<int32>& pg = m_strips[(size_t)r].proxyGlobal;
		for (size_t k = 0; k < ow.m_contactRecords.size(); ++k)
		{
			b2cuContact rec = ow.m_contactRecords[k];
			const int32 ga = pg[(size_t)rec.proxyA], gb = pg[(size_t)rec.proxyB];
			const int32 oa = owner[(size_t)proxyBody[(size_t)ga]], ob = owner[(size_t)proxyBody[(size_t)gb]];
			const int32 to = oa < 0 ? ob : ob < 0 ? oa : std::min(oa, ob);
			if (to < 0) continue;
			const int32 la = proxyLocal[(size_t)to][(size_t)ga], lb = proxyLocal[(size_t)to][(size_t)gb];
			if (la < 0 || lb < 0)
			{
				++lost;
				continue;
			}
			rec.proxyA = la;
			rec.proxyB = lb;
			worlds[(size_t)to]->m_contactRecords.push_back(rec);
		}
	}
	m_lostContacts += lost;
	if (Link(worlds, executors, strips) != B2CU_OK)
	{
		for (int32 r = 0; r < n; ++r)
		{
			delete worlds[(size_t)r];
			delete executors[(size_t)r];
		}
		return false;
	}
	for (int32 r = 0; r < n; ++r)
	{
		delete m_worlds[(size_t)r];
		delete m_executors[(size_t)r];
	}
	m_worlds.swap(worlds);
	m_executors.swap(executors);
	m_strips.swap(strips);
	m_bounds.swap(bounds);
	m_owner.swap(owner);
	m_localIndex.swap(localIndex);
	m_stepsSinceRebalance = 0;
	return true;
}
