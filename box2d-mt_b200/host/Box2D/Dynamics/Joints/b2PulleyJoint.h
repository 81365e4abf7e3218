// Pulley joint (reference: Box2D/Dynamics/Joints/b2PulleyJoint.h:25-152): two bodies hang from two fixed ground anchors
// on one rope; lengthA + ratio * lengthB stays constant, so a ratio other than 1 makes a block and tackle.
#ifndef B2_PULLEY_JOINT_H
#define B2_PULLEY_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

const float32 b2_minPulleyLength = 2.0f;

struct b2PulleyJointDef : public b2JointDef
{
	b2PulleyJointDef() : lengthA(0.0f), lengthB(0.0f), ratio(1.0f)
	{
		type = e_pulleyJoint;
		groundAnchorA.Set(-1.0f, 1.0f);
		groundAnchorB.Set(1.0f, 1.0f);
		localAnchorA.Set(-1.0f, 0.0f);
		localAnchorB.Set(1.0f, 0.0f);
		collideConnected = true;
	}

	/// everything from world anchors; the rope lengths are the current distances
	void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& groundAnchorA, const b2Vec2& groundAnchorB,
	                const b2Vec2& anchorA, const b2Vec2& anchorB, float32 ratio);

	b2Vec2 groundAnchorA, groundAnchorB; ///< world coordinates, never move
	b2Vec2 localAnchorA, localAnchorB;
	float32 lengthA, lengthB;
	float32 ratio;
};

class b2PulleyJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	b2Vec2 GetGroundAnchorA() const { return m_groundAnchorA; }
	b2Vec2 GetGroundAnchorB() const { return m_groundAnchorB; }
	float32 GetLengthA() const { return m_lengthA; }
	float32 GetLengthB() const { return m_lengthB; }
	float32 GetRatio() const { return m_ratio; }
	float32 GetCurrentLengthA() const;
	float32 GetCurrentLengthB() const;
	void ShiftOrigin(const b2Vec2& newOrigin) override;

protected:
	friend class b2World;
	explicit b2PulleyJoint(const b2PulleyJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_groundAnchorA, m_groundAnchorB, m_localAnchorA, m_localAnchorB;
	float32 m_lengthA, m_lengthB, m_ratio;
	float32 m_impulse;
	b2Vec2 m_uB;
};

#endif
