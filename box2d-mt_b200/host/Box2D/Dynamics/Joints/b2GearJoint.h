// Gear joint (reference: Box2D/Dynamics/Joints/b2GearJoint.h:25-125): ties the coordinates of two other joints, each a
// revolute or a prismatic joint, together: coordinate1 + ratio * coordinate2 = constant.  The two joints must outlive
// the gear joint (destroy it first).
#ifndef B2_GEAR_JOINT_H
#define B2_GEAR_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2GearJointDef : public b2JointDef
{
	b2GearJointDef() : joint1(nullptr), joint2(nullptr), ratio(1.0f) { type = e_gearJoint; }

	b2Joint* joint1; ///< revolute or prismatic
	b2Joint* joint2; ///< revolute or prismatic
	float32 ratio;
};

class b2GearJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	b2Joint* GetJoint1() { return m_joint1; }
	b2Joint* GetJoint2() { return m_joint2; }
	void SetRatio(float32 ratio);
	float32 GetRatio() const { return m_ratio; }

protected:
	friend class b2World;
	explicit b2GearJoint(const b2GearJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Joint* m_joint1;
	b2Joint* m_joint2;
	b2JointType m_typeA, m_typeB;
	// A and B are the second bodies of joint 1 and 2 (the base class's m_bodyA / m_bodyB), C and D their first bodies
	b2Body* m_bodyC;
	b2Body* m_bodyD;
	b2Vec2 m_localAnchorA, m_localAnchorB, m_localAnchorC, m_localAnchorD;
	b2Vec2 m_localAxisC, m_localAxisD;
	float32 m_referenceAngleA, m_referenceAngleB;
	float32 m_constant, m_ratio;
	// persistent solver state and the Jacobian terms of the last solve
	float32 m_impulse;
	b2Vec2 m_JvAC;
	float32 m_JwA;
};

#endif
