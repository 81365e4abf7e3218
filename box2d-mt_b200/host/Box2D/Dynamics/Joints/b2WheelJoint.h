// Wheel joint (reference: Box2D/Dynamics/Joints/b2WheelJoint.h:25-214): a point of body B stays on a line fixed in body
// A, a spring along the line gives suspension, a motor can turn body B.  Made for vehicle wheels.
#ifndef B2_WHEEL_JOINT_H
#define B2_WHEEL_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2WheelJointDef : public b2JointDef
{
	b2WheelJointDef() : enableMotor(false), maxMotorTorque(0.0f), motorSpeed(0.0f), frequencyHz(2.0f), dampingRatio(0.7f)
	{
		type = e_wheelJoint;
		localAnchorA.Set(0.0f, 0.0f);
		localAnchorB.Set(0.0f, 0.0f);
		localAxisA.Set(1.0f, 0.0f);
	}

	void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor, const b2Vec2& axis);

	b2Vec2 localAnchorA, localAnchorB;
	b2Vec2 localAxisA; ///< suspension axis in body A's frame
	bool enableMotor;
	float32 maxMotorTorque, motorSpeed;
	float32 frequencyHz, dampingRatio; ///< suspension spring
};

class b2WheelJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
	const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
	const b2Vec2& GetLocalAxisA() const { return m_localXAxisA; }
	float32 GetJointTranslation() const;
	float32 GetJointLinearSpeed() const;
	float32 GetJointAngle() const;
	float32 GetJointAngularSpeed() const;

	bool IsMotorEnabled() const { return m_enableMotor; }
	void EnableMotor(bool flag);
	void SetMotorSpeed(float32 speed);
	float32 GetMotorSpeed() const { return m_motorSpeed; }
	void SetMaxMotorTorque(float32 torque);
	float32 GetMaxMotorTorque() const { return m_maxMotorTorque; }
	float32 GetMotorTorque(float32 inv_dt) const;
	void SetSpringFrequencyHz(float32 hz);
	float32 GetSpringFrequencyHz() const { return m_frequencyHz; }
	void SetSpringDampingRatio(float32 ratio);
	float32 GetSpringDampingRatio() const { return m_dampingRatio; }

protected:
	friend class b2World;
	explicit b2WheelJoint(const b2WheelJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_localAnchorA, m_localAnchorB, m_localXAxisA;
	bool m_enableMotor;
	float32 m_maxMotorTorque, m_motorSpeed, m_frequencyHz, m_dampingRatio;
	// persistent solver state, as the device keeps it
	float32 m_impulse, m_springImpulse, m_motorImpulse;
	b2Vec2 m_ax, m_ay;
	float32 m_sAx, m_sBx;
};

#endif
