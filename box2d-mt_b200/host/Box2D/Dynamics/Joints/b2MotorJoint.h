// Motor joint (reference: Box2D/Dynamics/Joints/b2MotorJoint.h:25-133): drives body B towards a position and angle
// relative to body A with bounded force and torque; typically body A is the ground.
#ifndef B2_MOTOR_JOINT_H
#define B2_MOTOR_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2MotorJointDef : public b2JointDef
{
	b2MotorJointDef() : angularOffset(0.0f), maxForce(1.0f), maxTorque(1.0f), correctionFactor(0.3f)
	{
		type = e_motorJoint;
		linearOffset.Set(0.0f, 0.0f);
	}

	/// offsets from the bodies' current transforms
	void Initialize(b2Body* bodyA, b2Body* bodyB);

	b2Vec2 linearOffset;      ///< position of body B in body A's frame
	float32 angularOffset;    ///< bodyB angle minus bodyA angle
	float32 maxForce, maxTorque;
	float32 correctionFactor; ///< 0..1, fraction of the error removed per step
};

class b2MotorJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	void SetLinearOffset(const b2Vec2& linearOffset);
	const b2Vec2& GetLinearOffset() const { return m_linearOffset; }
	void SetAngularOffset(float32 angularOffset);
	float32 GetAngularOffset() const { return m_angularOffset; }
	void SetMaxForce(float32 force);
	float32 GetMaxForce() const { return m_maxForce; }
	void SetMaxTorque(float32 torque);
	float32 GetMaxTorque() const { return m_maxTorque; }
	void SetCorrectionFactor(float32 factor);
	float32 GetCorrectionFactor() const { return m_correctionFactor; }

protected:
	friend class b2World;
	explicit b2MotorJoint(const b2MotorJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_linearOffset;
	float32 m_angularOffset, m_maxForce, m_maxTorque, m_correctionFactor;
	b2Vec2 m_linearImpulse;
	float32 m_angularImpulse;
};

#endif
