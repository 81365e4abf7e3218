// Weld joint (reference: Box2D/Dynamics/Joints/b2WeldJoint.h:25-126): glues two bodies together at an anchor; with a
// frequency the angle between them becomes a torsion spring.
#ifndef B2_WELD_JOINT_H
#define B2_WELD_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2WeldJointDef : public b2JointDef
{
	b2WeldJointDef() : referenceAngle(0.0f), frequencyHz(0.0f), dampingRatio(0.0f)
	{
		type = e_weldJoint;
		localAnchorA.Set(0.0f, 0.0f);
		localAnchorB.Set(0.0f, 0.0f);
	}

	/// bodies, local anchors and reference angle from the bodies' current transforms and a world anchor
	void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor);

	b2Vec2 localAnchorA, localAnchorB;
	float32 referenceAngle; ///< bodyB angle minus bodyA angle that the joint holds
	float32 frequencyHz;    ///< torsion spring frequency; 0 = rigid
	float32 dampingRatio;
};

class b2WeldJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
	const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
	float32 GetReferenceAngle() const { return m_referenceAngle; }
	void SetFrequency(float32 hz);
	float32 GetFrequency() const { return m_frequencyHz; }
	void SetDampingRatio(float32 ratio);
	float32 GetDampingRatio() const { return m_dampingRatio; }

protected:
	friend class b2World;
	explicit b2WeldJoint(const b2WeldJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_localAnchorA, m_localAnchorB;
	float32 m_referenceAngle, m_frequencyHz, m_dampingRatio;
	b2Vec3 m_impulse; // persistent solver state
};

#endif
