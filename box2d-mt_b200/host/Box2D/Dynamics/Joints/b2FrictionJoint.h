// Friction joint (reference: Box2D/Dynamics/Joints/b2FrictionJoint.h:25-119): bounded linear and angular friction
// between two bodies, for top-down games.
#ifndef B2_FRICTION_JOINT_H
#define B2_FRICTION_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2FrictionJointDef : public b2JointDef
{
	b2FrictionJointDef() : maxForce(0.0f), maxTorque(0.0f)
	{
		type = e_frictionJoint;
		localAnchorA.Set(0.0f, 0.0f);
		localAnchorB.Set(0.0f, 0.0f);
	}

	void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor);

	b2Vec2 localAnchorA, localAnchorB;
	float32 maxForce;  ///< N
	float32 maxTorque; ///< N*m
};

class b2FrictionJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
	const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
	void SetMaxForce(float32 force);
	float32 GetMaxForce() const { return m_maxForce; }
	void SetMaxTorque(float32 torque);
	float32 GetMaxTorque() const { return m_maxTorque; }

protected:
	friend class b2World;
	explicit b2FrictionJoint(const b2FrictionJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_localAnchorA, m_localAnchorB;
	float32 m_maxForce, m_maxTorque;
	b2Vec2 m_linearImpulse;
	float32 m_angularImpulse;
};

#endif
