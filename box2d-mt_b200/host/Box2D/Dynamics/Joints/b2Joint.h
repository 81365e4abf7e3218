// Joint base class (reference: Box2D/Dynamics/Joints/b2Joint.h:28-226).  A joint is a host handle: its parameters and
// its persistent solver state (accumulated impulses) live in the object, travel to the device as one b2cuJoint row of
// the world's joint table, and come back after a step when somebody asks for them.  The solve itself is the device's
// (csrc/b2cu_joints.cuh).  This version of the GPU path solves all eleven joint types.
#ifndef B2_JOINT_H
#define B2_JOINT_H

#include "Box2D/Common/b2Math.h"
#include "b2cuda.h"

class b2Body;
class b2Joint;
class b2World;

enum b2JointType
{
	e_unknownJoint,
	e_revoluteJoint,
	e_prismaticJoint,
	e_distanceJoint,
	e_pulleyJoint,
	e_mouseJoint,
	e_gearJoint,
	e_wheelJoint,
	e_weldJoint,
	e_frictionJoint,
	e_ropeJoint,
	e_motorJoint
};

enum b2LimitState
{
	e_inactiveLimit = B2CU_LIMIT_INACTIVE,
	e_atLowerLimit = B2CU_LIMIT_AT_LOWER,
	e_atUpperLimit = B2CU_LIMIT_AT_UPPER,
	e_equalLimits = B2CU_LIMIT_EQUAL
};

/// node of a body's joint list: the joint and the body at its other end
struct b2JointEdge
{
	b2Body* other;
	b2Joint* joint;
	b2JointEdge* prev;
	b2JointEdge* next;
};

struct b2JointDef
{
	b2JointDef() : type(e_unknownJoint), userData(nullptr), bodyA(nullptr), bodyB(nullptr), collideConnected(false) {}

	b2JointType type;
	void* userData;
	b2Body* bodyA;
	b2Body* bodyB;
	bool collideConnected; ///< may the two bodies collide with each other?
};

class b2Joint
{
public:
	b2JointType GetType() const { return m_type; }
	b2Body* GetBodyA() { return m_bodyA; }
	b2Body* GetBodyB() { return m_bodyB; }

	/// anchor points in world coordinates
	virtual b2Vec2 GetAnchorA() const = 0;
	virtual b2Vec2 GetAnchorB() const = 0;
	/// reaction on body B at the anchor, from the impulses of the last step (N, N*m)
	virtual b2Vec2 GetReactionForce(float32 inv_dt) const = 0;
	virtual float32 GetReactionTorque(float32 inv_dt) const = 0;

	b2Joint* GetNext() { return m_next; }
	const b2Joint* GetNext() const { return m_next; }
	void* GetUserData() const { return m_userData; }
	void SetUserData(void* data) { m_userData = data; }
	/// both bodies active?
	bool IsActive() const;
	bool GetCollideConnected() const { return m_collideConnected; }
	virtual void ShiftOrigin(const b2Vec2& newOrigin) { B2_NOT_USED(newOrigin); }
	/// row of the world's joint table (dense, creation order)
	int32 GetIndex() const { return m_index; }

protected:
	friend class b2World;
	friend class b2Body;

	explicit b2Joint(const b2JointDef* def);
	virtual ~b2Joint() {}

	/// the joint as a row of the device table / the persistent solver state of a row back into the joint
	virtual void WriteRecord(b2cuJoint* out) const = 0;
	virtual void ReadRecord(const b2cuJoint& in) = 0;
	void Touch();         ///< a parameter changed: fetch the device's state first, then mark the table for upload
	void Refresh() const; ///< make the persistent state current

	b2JointType m_type;
	b2Joint* m_prev;
	b2Joint* m_next;
	b2JointEdge m_edgeA, m_edgeB;
	b2Body* m_bodyA;
	b2Body* m_bodyB;
	b2World* m_world;
	int32 m_index;
	bool m_collideConnected;
	void* m_userData;
};

#endif
