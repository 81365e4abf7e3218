// Prismatic joint (reference: Box2D/Dynamics/Joints/b2PrismaticJoint.h:25-196): body B slides along an axis fixed in
// body A and cannot rotate against it; optionally between two translations and optionally driven by a motor.
#ifndef B2_PRISMATIC_JOINT_H
#define B2_PRISMATIC_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2PrismaticJointDef : public b2JointDef
{
	b2PrismaticJointDef()
		: referenceAngle(0.0f), enableLimit(false), lowerTranslation(0.0f), upperTranslation(0.0f), enableMotor(false),
		  maxMotorForce(0.0f), motorSpeed(0.0f)
	{
		type = e_prismaticJoint;
		localAnchorA.Set(0.0f, 0.0f);
		localAnchorB.Set(0.0f, 0.0f);
		localAxisA.Set(1.0f, 0.0f);
	}

	/// bodies, local anchors, local axis and reference angle from a world anchor and a world axis
	void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor, const b2Vec2& axis);

	b2Vec2 localAnchorA, localAnchorB;
	b2Vec2 localAxisA;      ///< translation axis in body A's frame (normalised by the joint)
	float32 referenceAngle; ///< bodyB angle minus bodyA angle that the joint holds
	bool enableLimit;
	float32 lowerTranslation, upperTranslation;
	bool enableMotor;
	float32 maxMotorForce;  ///< N
	float32 motorSpeed;     ///< m/s
};

class b2PrismaticJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
	const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
	const b2Vec2& GetLocalAxisA() const { return m_localXAxisA; }
	float32 GetReferenceAngle() const { return m_referenceAngle; }
	float32 GetJointTranslation() const;
	float32 GetJointSpeed() const;

	bool IsLimitEnabled() const { return m_enableLimit; }
	void EnableLimit(bool flag);
	float32 GetLowerLimit() const { return m_lowerTranslation; }
	float32 GetUpperLimit() const { return m_upperTranslation; }
	void SetLimits(float32 lower, float32 upper);

	bool IsMotorEnabled() const { return m_enableMotor; }
	void EnableMotor(bool flag);
	void SetMotorSpeed(float32 speed);
	float32 GetMotorSpeed() const { return m_motorSpeed; }
	void SetMaxMotorForce(float32 force);
	float32 GetMaxMotorForce() const { return m_maxMotorForce; }
	float32 GetMotorForce(float32 inv_dt) const;

protected:
	friend class b2World;
	explicit b2PrismaticJoint(const b2PrismaticJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;
	void WakeBodies();

	b2Vec2 m_localAnchorA, m_localAnchorB;
	b2Vec2 m_localAxisGiven;  // the definition's axis: what the device row carries (it normalises like the constructor)
	b2Vec2 m_localXAxisA;     // normalised
	float32 m_referenceAngle;
	bool m_enableLimit, m_enableMotor;
	float32 m_lowerTranslation, m_upperTranslation;
	float32 m_maxMotorForce, m_motorSpeed;
	// persistent solver state; m_axis / m_perp are the world directions of the last solve
	b2Vec3 m_impulse;
	float32 m_motorImpulse;
	b2LimitState m_limitState;
	b2Vec2 m_axis, m_perp;
};

#endif
