// Revolute joint (reference: Box2D/Dynamics/Joints/b2RevoluteJoint.h:25-204): two bodies share an anchor point and
// rotate freely about it, optionally between two angles and optionally driven by a motor.
#ifndef B2_REVOLUTE_JOINT_H
#define B2_REVOLUTE_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2RevoluteJointDef : public b2JointDef
{
	b2RevoluteJointDef()
		: referenceAngle(0.0f), enableLimit(false), lowerAngle(0.0f), upperAngle(0.0f), enableMotor(false), motorSpeed(0.0f),
		  maxMotorTorque(0.0f)
	{
		type = e_revoluteJoint;
		localAnchorA.Set(0.0f, 0.0f);
		localAnchorB.Set(0.0f, 0.0f);
	}

	/// bodies, local anchors and reference angle from the bodies' current transforms and a world anchor
	void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor);

	b2Vec2 localAnchorA, localAnchorB;
	float32 referenceAngle; ///< bodyB angle minus bodyA angle at which the joint angle reads zero
	bool enableLimit;
	float32 lowerAngle, upperAngle;
	bool enableMotor;
	float32 motorSpeed;     ///< rad/s
	float32 maxMotorTorque; ///< N*m
};

class b2RevoluteJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
	const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
	float32 GetReferenceAngle() const { return m_referenceAngle; }
	float32 GetJointAngle() const;
	float32 GetJointSpeed() const;

	bool IsLimitEnabled() const { return m_enableLimit; }
	void EnableLimit(bool flag);
	float32 GetLowerLimit() const { return m_lowerAngle; }
	float32 GetUpperLimit() const { return m_upperAngle; }
	void SetLimits(float32 lower, float32 upper);

	bool IsMotorEnabled() const { return m_enableMotor; }
	void EnableMotor(bool flag);
	void SetMotorSpeed(float32 speed);
	float32 GetMotorSpeed() const { return m_motorSpeed; }
	void SetMaxMotorTorque(float32 torque);
	float32 GetMaxMotorTorque() const { return m_maxMotorTorque; }

	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;
	float32 GetMotorTorque(float32 inv_dt) const;

protected:
	friend class b2World;
	explicit b2RevoluteJoint(const b2RevoluteJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;
	void WakeBodies();

	b2Vec2 m_localAnchorA, m_localAnchorB;
	float32 m_referenceAngle;
	bool m_enableLimit, m_enableMotor;
	float32 m_lowerAngle, m_upperAngle;
	float32 m_motorSpeed, m_maxMotorTorque;
	// persistent solver state
	b2Vec3 m_impulse;
	float32 m_motorImpulse;
	b2LimitState m_limitState;
};

#endif
