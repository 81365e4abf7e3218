// Rope joint (reference: Box2D/Dynamics/Joints/b2RopeJoint.h:25-116): the two anchors may come closer but never be
// further apart than maxLength.
#ifndef B2_ROPE_JOINT_H
#define B2_ROPE_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2RopeJointDef : public b2JointDef
{
	b2RopeJointDef() : maxLength(0.0f)
	{
		type = e_ropeJoint;
		localAnchorA.Set(-1.0f, 0.0f);
		localAnchorB.Set(1.0f, 0.0f);
	}

	b2Vec2 localAnchorA, localAnchorB;
	float32 maxLength; ///< must exceed b2_linearSlop
};

class b2RopeJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
	const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
	void SetMaxLength(float32 length);
	float32 GetMaxLength() const { return m_maxLength; }
	b2LimitState GetLimitState() const;

protected:
	friend class b2World;
	explicit b2RopeJoint(const b2RopeJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_localAnchorA, m_localAnchorB;
	float32 m_maxLength;
	float32 m_impulse;
	b2LimitState m_state;
	b2Vec2 m_u;
};

#endif
