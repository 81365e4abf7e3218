// Distance joint (reference: Box2D/Dynamics/Joints/b2DistanceJoint.h:25-169): keeps two anchor points a fixed distance
// apart, like a massless rod; with a frequency it becomes a spring-damper.
#ifndef B2_DISTANCE_JOINT_H
#define B2_DISTANCE_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2DistanceJointDef : public b2JointDef
{
	b2DistanceJointDef() : length(1.0f), frequencyHz(0.0f), dampingRatio(0.0f)
	{
		type = e_distanceJoint;
		localAnchorA.Set(0.0f, 0.0f);
		localAnchorB.Set(0.0f, 0.0f);
	}

	/// bodies, local anchors and length from two world anchors
	void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchorA, const b2Vec2& anchorB);

	b2Vec2 localAnchorA, localAnchorB;
	float32 length;       ///< rest length between the anchors
	float32 frequencyHz;  ///< mass-spring-damper frequency; 0 = rigid
	float32 dampingRatio; ///< 0 = none, 1 = critical
};

class b2DistanceJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
	const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
	/// changing the length does not wake the bodies, as in the reference
	void SetLength(float32 length);
	float32 GetLength() const { return m_length; }
	void SetFrequency(float32 hz);
	float32 GetFrequency() const { return m_frequencyHz; }
	void SetDampingRatio(float32 ratio);
	float32 GetDampingRatio() const { return m_dampingRatio; }

protected:
	friend class b2World;
	explicit b2DistanceJoint(const b2DistanceJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_localAnchorA, m_localAnchorB;
	float32 m_length, m_frequencyHz, m_dampingRatio;
	// persistent solver state; m_u is the unit vector of the last solve
	float32 m_impulse;
	b2Vec2 m_u;
};

#endif
