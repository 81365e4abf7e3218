// Mouse joint (reference: Box2D/Dynamics/Joints/b2MouseJoint.h:25-129): pulls a point of body B towards a world target
// with a soft, force-limited constraint; made for dragging bodies with a pointer.  Body A is only a formality.
#ifndef B2_MOUSE_JOINT_H
#define B2_MOUSE_JOINT_H

#include "Box2D/Dynamics/Joints/b2Joint.h"

struct b2MouseJointDef : public b2JointDef
{
	b2MouseJointDef() : maxForce(0.0f), frequencyHz(5.0f), dampingRatio(0.7f)
	{
		type = e_mouseJoint;
		target.Set(0.0f, 0.0f);
	}

	b2Vec2 target;        ///< world point; the grabbed point of body B is wherever the target is at creation
	float32 maxForce;     ///< usually a multiple of the body's weight
	float32 frequencyHz, dampingRatio;
};

class b2MouseJoint : public b2Joint
{
public:
	b2Vec2 GetAnchorA() const override;
	b2Vec2 GetAnchorB() const override;
	b2Vec2 GetReactionForce(float32 inv_dt) const override;
	float32 GetReactionTorque(float32 inv_dt) const override;

	void SetTarget(const b2Vec2& target);
	const b2Vec2& GetTarget() const { return m_targetA; }
	void SetMaxForce(float32 force);
	float32 GetMaxForce() const { return m_maxForce; }
	void SetFrequency(float32 hz);
	float32 GetFrequency() const { return m_frequencyHz; }
	void SetDampingRatio(float32 ratio);
	float32 GetDampingRatio() const { return m_dampingRatio; }
	void ShiftOrigin(const b2Vec2& newOrigin) override;

protected:
	friend class b2World;
	explicit b2MouseJoint(const b2MouseJointDef* def);
	void WriteRecord(b2cuJoint* out) const override;
	void ReadRecord(const b2cuJoint& in) override;

	b2Vec2 m_localAnchorB, m_targetA;
	float32 m_maxForce, m_frequencyHz, m_dampingRatio;
	b2Vec2 m_impulse;
};

#endif
