// Host side of the joints (reference: Box2D/Dynamics/Joints/b2Joint.cpp, b2RevoluteJoint.cpp:36-62, :379-512, b2DistanceJoint.cpp,
// b2WeldJoint.cpp; the
// world's part is b2World::CreateJoint / DestroyJoint, b2World.cpp:659-841).  Nothing is solved here: the joint objects
// hold parameters and the persistent impulses, and exchange them with the device's joint table.
#include "Box2D/Dynamics/Joints/b2RevoluteJoint.h"
#include "Box2D/Dynamics/Joints/b2DistanceJoint.h"
#include "Box2D/Dynamics/Joints/b2WeldJoint.h"
#include "Box2D/Dynamics/Joints/b2PrismaticJoint.h"
#include "Box2D/Dynamics/Joints/b2WheelJoint.h"
#include "Box2D/Dynamics/Joints/b2RopeJoint.h"
#include "Box2D/Dynamics/Joints/b2FrictionJoint.h"
#include "Box2D/Dynamics/Joints/b2MotorJoint.h"
#include "Box2D/Dynamics/Joints/b2PulleyJoint.h"
#include "Box2D/Dynamics/Joints/b2MouseJoint.h"
#include "Box2D/Dynamics/Joints/b2GearJoint.h"
#include "Box2D/Dynamics/b2Body.h"
#include "Box2D/Dynamics/b2World.h"

b2Joint::b2Joint(const b2JointDef* def)
	: m_type(def->type), m_prev(nullptr), m_next(nullptr), m_bodyA(def->bodyA), m_bodyB(def->bodyB), m_world(nullptr),
	  m_index(-1), m_collideConnected(def->collideConnected), m_userData(def->userData)
{
	b2Assert(def->bodyA != def->bodyB);
	m_edgeA.joint = m_edgeB.joint = nullptr;
	m_edgeA.other = m_edgeB.other = nullptr;
	m_edgeA.prev = m_edgeA.next = m_edgeB.prev = m_edgeB.next = nullptr;
}

bool b2Joint::IsActive() const { return m_bodyA->IsActive() && m_bodyB->IsActive(); }

void b2Joint::Refresh() const { m_world->RefreshJoints(); }

void b2Joint::Touch()
{
	m_world->RefreshJoints();
	m_world->m_jointsDirty = true;
}

// ---- revolute ----------------------------------------------------------------------------------------------------------

void b2RevoluteJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor)
{
	bodyA = bA;
	bodyB = bB;
	localAnchorA = bA->GetLocalPoint(anchor);
	localAnchorB = bB->GetLocalPoint(anchor);
	referenceAngle = bB->GetAngle() - bA->GetAngle();
}

b2RevoluteJoint::b2RevoluteJoint(const b2RevoluteJointDef* def)
	: b2Joint(def), m_localAnchorA(def->localAnchorA), m_localAnchorB(def->localAnchorB),
	  m_referenceAngle(def->referenceAngle), m_enableLimit(def->enableLimit), m_enableMotor(def->enableMotor),
	  m_lowerAngle(def->lowerAngle), m_upperAngle(def->upperAngle), m_motorSpeed(def->motorSpeed),
	  m_maxMotorTorque(def->maxMotorTorque), m_impulse(0.0f, 0.0f, 0.0f), m_motorImpulse(0.0f), m_limitState(e_inactiveLimit)
{
}

void b2RevoluteJoint::WriteRecord(b2cuJoint* out) const
{
	b2cuJoint r = b2cuJoint();
	r.type = B2CU_JOINT_REVOLUTE;
	r.bodyA = m_bodyA->GetIndex();
	r.bodyB = m_bodyB->GetIndex();
	r.flags = (m_collideConnected ? B2CU_JOINT_COLLIDE_CONNECTED : 0u) | (m_enableLimit ? B2CU_JOINT_ENABLE_LIMIT : 0u) |
	          (m_enableMotor ? B2CU_JOINT_ENABLE_MOTOR : 0u);
	r.localAnchorA[0] = m_localAnchorA.x;
	r.localAnchorA[1] = m_localAnchorA.y;
	r.localAnchorB[0] = m_localAnchorB.x;
	r.localAnchorB[1] = m_localAnchorB.y;
	r.referenceAngle = m_referenceAngle;
	r.lowerAngle = m_lowerAngle;
	r.upperAngle = m_upperAngle;
	r.maxMotorTorque = m_maxMotorTorque;
	r.motorSpeed = m_motorSpeed;
	r.impulse[0] = m_impulse.x;
	r.impulse[1] = m_impulse.y;
	r.impulse[2] = m_impulse.z;
	r.motorImpulse = m_motorImpulse;
	r.limitState = (int32_t)m_limitState;
	*out = r;
}

void b2RevoluteJoint::ReadRecord(const b2cuJoint& in)
{
	m_impulse.Set(in.impulse[0], in.impulse[1], in.impulse[2]);
	m_motorImpulse = in.motorImpulse;
	m_limitState = (b2LimitState)in.limitState;
}

b2Vec2 b2RevoluteJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2RevoluteJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

float32 b2RevoluteJoint::GetJointAngle() const { return m_bodyB->GetAngle() - m_bodyA->GetAngle() - m_referenceAngle; }
float32 b2RevoluteJoint::GetJointSpeed() const { return m_bodyB->GetAngularVelocity() - m_bodyA->GetAngularVelocity(); }

b2Vec2 b2RevoluteJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	return b2Vec2(inv_dt * m_impulse.x, inv_dt * m_impulse.y);
}

float32 b2RevoluteJoint::GetReactionTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_impulse.z;
}

float32 b2RevoluteJoint::GetMotorTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_motorImpulse;
}

void b2RevoluteJoint::WakeBodies()
{
	m_bodyA->SetAwake(true);
	m_bodyB->SetAwake(true);
}

void b2RevoluteJoint::EnableMotor(bool flag)
{
	if (flag == m_enableMotor) return;
	Touch();
	WakeBodies();
	m_enableMotor = flag;
}

void b2RevoluteJoint::SetMotorSpeed(float32 speed)
{
	if (speed == m_motorSpeed) return;
	Touch();
	WakeBodies();
	m_motorSpeed = speed;
}

void b2RevoluteJoint::SetMaxMotorTorque(float32 torque)
{
	if (torque == m_maxMotorTorque) return;
	Touch();
	WakeBodies();
	m_maxMotorTorque = torque;
}

void b2RevoluteJoint::EnableLimit(bool flag)
{
	if (flag == m_enableLimit) return;
	Touch();
	WakeBodies();
	m_enableLimit = flag;
	m_impulse.z = 0.0f;
}

void b2RevoluteJoint::SetLimits(float32 lower, float32 upper)
{
	b2Assert(lower <= upper);
	if (lower == m_lowerAngle && upper == m_upperAngle) return;
	Touch();
	WakeBodies();
	m_impulse.z = 0.0f;
	m_lowerAngle = lower;
	m_upperAngle = upper;
}

// ---- distance (reference b2DistanceJoint.cpp:40-61, :224-260) -----------------------------------------------------------

void b2DistanceJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchorA, const b2Vec2& anchorB)
{
	bodyA = bA;
	bodyB = bB;
	localAnchorA = bA->GetLocalPoint(anchorA);
	localAnchorB = bB->GetLocalPoint(anchorB);
	length = (anchorB - anchorA).Length();
}

b2DistanceJoint::b2DistanceJoint(const b2DistanceJointDef* def)
	: b2Joint(def), m_localAnchorA(def->localAnchorA), m_localAnchorB(def->localAnchorB), m_length(def->length),
	  m_frequencyHz(def->frequencyHz), m_dampingRatio(def->dampingRatio), m_impulse(0.0f), m_u(0.0f, 0.0f)
{
}

static void WriteCommon(b2cuJoint* r, int32 type, const b2Body* bodyA, const b2Body* bodyB, bool collideConnected,
                        const b2Vec2& anchorA, const b2Vec2& anchorB)
{
	*r = b2cuJoint();
	r->type = type;
	r->bodyA = bodyA->GetIndex();
	r->bodyB = bodyB->GetIndex();
	r->flags = collideConnected ? B2CU_JOINT_COLLIDE_CONNECTED : 0u;
	r->localAnchorA[0] = anchorA.x;
	r->localAnchorA[1] = anchorA.y;
	r->localAnchorB[0] = anchorB.x;
	r->localAnchorB[1] = anchorB.y;
}

void b2DistanceJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_DISTANCE, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->length = m_length;
	out->frequencyHz = m_frequencyHz;
	out->dampingRatio = m_dampingRatio;
	out->impulse[0] = m_impulse;
	out->lastSolve[0] = m_u.x;
	out->lastSolve[1] = m_u.y;
}

void b2DistanceJoint::ReadRecord(const b2cuJoint& in)
{
	m_impulse = in.impulse[0];
	m_u.Set(in.lastSolve[0], in.lastSolve[1]);
}

b2Vec2 b2DistanceJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2DistanceJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2DistanceJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	float32 scale = inv_dt * m_impulse;
	return b2Vec2(scale * m_u.x, scale * m_u.y);
}

float32 b2DistanceJoint::GetReactionTorque(float32 inv_dt) const
{
	B2_NOT_USED(inv_dt);
	return 0.0f;
}

void b2DistanceJoint::SetLength(float32 length)
{
	if (length == m_length) return;
	Touch();
	m_length = length;
}

void b2DistanceJoint::SetFrequency(float32 hz)
{
	if (hz == m_frequencyHz) return;
	Touch();
	m_frequencyHz = hz;
}

void b2DistanceJoint::SetDampingRatio(float32 ratio)
{
	if (ratio == m_dampingRatio) return;
	Touch();
	m_dampingRatio = ratio;
}

// ---- weld (reference b2WeldJoint.cpp:37-57, :310-332) -------------------------------------------------------------------

void b2WeldJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor)
{
	bodyA = bA;
	bodyB = bB;
	localAnchorA = bA->GetLocalPoint(anchor);
	localAnchorB = bB->GetLocalPoint(anchor);
	referenceAngle = bB->GetAngle() - bA->GetAngle();
}

b2WeldJoint::b2WeldJoint(const b2WeldJointDef* def)
	: b2Joint(def), m_localAnchorA(def->localAnchorA), m_localAnchorB(def->localAnchorB),
	  m_referenceAngle(def->referenceAngle), m_frequencyHz(def->frequencyHz), m_dampingRatio(def->dampingRatio),
	  m_impulse(0.0f, 0.0f, 0.0f)
{
}

void b2WeldJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_WELD, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->referenceAngle = m_referenceAngle;
	out->frequencyHz = m_frequencyHz;
	out->dampingRatio = m_dampingRatio;
	out->impulse[0] = m_impulse.x;
	out->impulse[1] = m_impulse.y;
	out->impulse[2] = m_impulse.z;
}

void b2WeldJoint::ReadRecord(const b2cuJoint& in) { m_impulse.Set(in.impulse[0], in.impulse[1], in.impulse[2]); }

b2Vec2 b2WeldJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2WeldJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2WeldJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	return b2Vec2(inv_dt * m_impulse.x, inv_dt * m_impulse.y);
}

float32 b2WeldJoint::GetReactionTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_impulse.z;
}

void b2WeldJoint::SetFrequency(float32 hz)
{
	if (hz == m_frequencyHz) return;
	Touch();
	m_frequencyHz = hz;
}

void b2WeldJoint::SetDampingRatio(float32 ratio)
{
	if (ratio == m_dampingRatio) return;
	Touch();
	m_dampingRatio = ratio;
}

// ---- prismatic (reference b2PrismaticJoint.cpp:91-125, :490-635) ---------------------------------------------------------

void b2PrismaticJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor, const b2Vec2& axis)
{
	bodyA = bA;
	bodyB = bB;
	localAnchorA = bA->GetLocalPoint(anchor);
	localAnchorB = bB->GetLocalPoint(anchor);
	localAxisA = bA->GetLocalVector(axis);
	referenceAngle = bB->GetAngle() - bA->GetAngle();
}

b2PrismaticJoint::b2PrismaticJoint(const b2PrismaticJointDef* def)
	: b2Joint(def), m_localAnchorA(def->localAnchorA), m_localAnchorB(def->localAnchorB), m_localAxisGiven(def->localAxisA),
	  m_localXAxisA(def->localAxisA), m_referenceAngle(def->referenceAngle), m_enableLimit(def->enableLimit),
	  m_enableMotor(def->enableMotor), m_lowerTranslation(def->lowerTranslation), m_upperTranslation(def->upperTranslation),
	  m_maxMotorForce(def->maxMotorForce), m_motorSpeed(def->motorSpeed), m_impulse(0.0f, 0.0f, 0.0f), m_motorImpulse(0.0f),
	  m_limitState(e_inactiveLimit), m_axis(0.0f, 0.0f), m_perp(0.0f, 0.0f)
{
	m_localXAxisA.Normalize();
}

void b2PrismaticJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_PRISMATIC, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->flags |= (m_enableLimit ? B2CU_JOINT_ENABLE_LIMIT : 0u) | (m_enableMotor ? B2CU_JOINT_ENABLE_MOTOR : 0u);
	out->axis[0] = m_localAxisGiven.x;
	out->axis[1] = m_localAxisGiven.y;
	out->referenceAngle = m_referenceAngle;
	out->lowerAngle = m_lowerTranslation;
	out->upperAngle = m_upperTranslation;
	out->maxMotorTorque = m_maxMotorForce;
	out->motorSpeed = m_motorSpeed;
	out->impulse[0] = m_impulse.x;
	out->impulse[1] = m_impulse.y;
	out->impulse[2] = m_impulse.z;
	out->motorImpulse = m_motorImpulse;
	out->limitState = (int32_t)m_limitState;
	out->lastSolve[0] = m_axis.x;
	out->lastSolve[1] = m_axis.y;
	out->lastSolve[2] = m_perp.x;
	out->lastSolve[3] = m_perp.y;
}

void b2PrismaticJoint::ReadRecord(const b2cuJoint& in)
{
	m_impulse.Set(in.impulse[0], in.impulse[1], in.impulse[2]);
	m_motorImpulse = in.motorImpulse;
	m_limitState = (b2LimitState)in.limitState;
	m_axis.Set(in.lastSolve[0], in.lastSolve[1]);
	m_perp.Set(in.lastSolve[2], in.lastSolve[3]);
}

b2Vec2 b2PrismaticJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2PrismaticJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2PrismaticJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	return inv_dt * (m_impulse.x * m_perp + (m_motorImpulse + m_impulse.z) * m_axis);
}

float32 b2PrismaticJoint::GetReactionTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_impulse.y;
}

float32 b2PrismaticJoint::GetMotorForce(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_motorImpulse;
}

float32 b2PrismaticJoint::GetJointTranslation() const
{
	b2Vec2 d = m_bodyB->GetWorldPoint(m_localAnchorB) - m_bodyA->GetWorldPoint(m_localAnchorA);
	b2Vec2 axis = m_bodyA->GetWorldVector(m_localXAxisA);
	return b2Dot(d, axis);
}

float32 b2PrismaticJoint::GetJointSpeed() const
{
	const b2Rot& qA = m_bodyA->GetTransform().q;
	const b2Rot& qB = m_bodyB->GetTransform().q;
	b2Vec2 rA = b2Mul(qA, m_localAnchorA - m_bodyA->GetLocalCenter());
	b2Vec2 rB = b2Mul(qB, m_localAnchorB - m_bodyB->GetLocalCenter());
	b2Vec2 p1 = m_bodyA->GetWorldCenter() + rA;
	b2Vec2 p2 = m_bodyB->GetWorldCenter() + rB;
	b2Vec2 d = p2 - p1;
	b2Vec2 axis = b2Mul(qA, m_localXAxisA);
	b2Vec2 vA = m_bodyA->GetLinearVelocity(), vB = m_bodyB->GetLinearVelocity();
	float32 wA = m_bodyA->GetAngularVelocity(), wB = m_bodyB->GetAngularVelocity();
	return b2Dot(d, b2Cross(wA, axis)) + b2Dot(axis, vB + b2Cross(wB, rB) - vA - b2Cross(wA, rA));
}

void b2PrismaticJoint::WakeBodies()
{
	m_bodyA->SetAwake(true);
	m_bodyB->SetAwake(true);
}

void b2PrismaticJoint::EnableLimit(bool flag)
{
	if (flag == m_enableLimit) return;
	Touch();
	WakeBodies();
	m_enableLimit = flag;
	m_impulse.z = 0.0f;
}

void b2PrismaticJoint::SetLimits(float32 lower, float32 upper)
{
	b2Assert(lower <= upper);
	if (lower == m_lowerTranslation && upper == m_upperTranslation) return;
	Touch();
	WakeBodies();
	m_lowerTranslation = lower;
	m_upperTranslation = upper;
	m_impulse.z = 0.0f;
}

void b2PrismaticJoint::EnableMotor(bool flag)
{
	if (flag == m_enableMotor) return;
	Touch();
	WakeBodies();
	m_enableMotor = flag;
}

void b2PrismaticJoint::SetMotorSpeed(float32 speed)
{
	if (speed == m_motorSpeed) return;
	Touch();
	WakeBodies();
	m_motorSpeed = speed;
}

void b2PrismaticJoint::SetMaxMotorForce(float32 force)
{
	if (force == m_maxMotorForce) return;
	Touch();
	WakeBodies();
	m_maxMotorForce = force;
}

// ---- wheel (reference b2WheelJoint.cpp:40-76, :320-470) ------------------------------------------------------------------

void b2WheelJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor, const b2Vec2& axis)
{
	bodyA = bA;
	bodyB = bB;
	localAnchorA = bA->GetLocalPoint(anchor);
	localAnchorB = bB->GetLocalPoint(anchor);
	localAxisA = bA->GetLocalVector(axis);
}

b2WheelJoint::b2WheelJoint(const b2WheelJointDef* def)
	: b2Joint(def), m_localAnchorA(def->localAnchorA), m_localAnchorB(def->localAnchorB), m_localXAxisA(def->localAxisA),
	  m_enableMotor(def->enableMotor), m_maxMotorTorque(def->maxMotorTorque), m_motorSpeed(def->motorSpeed),
	  m_frequencyHz(def->frequencyHz), m_dampingRatio(def->dampingRatio), m_impulse(0.0f), m_springImpulse(0.0f),
	  m_motorImpulse(0.0f), m_ax(0.0f, 0.0f), m_ay(0.0f, 0.0f), m_sAx(0.0f), m_sBx(0.0f)
{
}

void b2WheelJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_WHEEL, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->flags |= m_enableMotor ? B2CU_JOINT_ENABLE_MOTOR : 0u;
	out->axis[0] = m_localXAxisA.x;
	out->axis[1] = m_localXAxisA.y;
	out->maxMotorTorque = m_maxMotorTorque;
	out->motorSpeed = m_motorSpeed;
	out->frequencyHz = m_frequencyHz;
	out->dampingRatio = m_dampingRatio;
	out->impulse[0] = m_impulse;
	out->impulse[1] = m_springImpulse;
	out->motorImpulse = m_motorImpulse;
	out->lastSolve[0] = m_ax.x;
	out->lastSolve[1] = m_ax.y;
	out->lastSolve[2] = m_ay.x;
	out->lastSolve[3] = m_ay.y;
	out->work[0] = m_sAx;
	out->work[1] = m_sBx;
}

void b2WheelJoint::ReadRecord(const b2cuJoint& in)
{
	m_impulse = in.impulse[0];
	m_springImpulse = in.impulse[1];
	m_motorImpulse = in.motorImpulse;
	m_ax.Set(in.lastSolve[0], in.lastSolve[1]);
	m_ay.Set(in.lastSolve[2], in.lastSolve[3]);
	m_sAx = in.work[0];
	m_sBx = in.work[1];
}

b2Vec2 b2WheelJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2WheelJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2WheelJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	return inv_dt * (m_impulse * m_ay + m_springImpulse * m_ax);
}

float32 b2WheelJoint::GetReactionTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_motorImpulse;
}

float32 b2WheelJoint::GetMotorTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_motorImpulse;
}

float32 b2WheelJoint::GetJointTranslation() const
{
	b2Vec2 d = m_bodyB->GetWorldPoint(m_localAnchorB) - m_bodyA->GetWorldPoint(m_localAnchorA);
	return b2Dot(d, m_bodyA->GetWorldVector(m_localXAxisA));
}

float32 b2WheelJoint::GetJointLinearSpeed() const
{
	const b2Rot& qA = m_bodyA->GetTransform().q;
	const b2Rot& qB = m_bodyB->GetTransform().q;
	b2Vec2 rA = b2Mul(qA, m_localAnchorA - m_bodyA->GetLocalCenter());
	b2Vec2 rB = b2Mul(qB, m_localAnchorB - m_bodyB->GetLocalCenter());
	b2Vec2 p1 = m_bodyA->GetWorldCenter() + rA;
	b2Vec2 p2 = m_bodyB->GetWorldCenter() + rB;
	b2Vec2 d = p2 - p1;
	b2Vec2 axis = b2Mul(qA, m_localXAxisA);
	b2Vec2 vA = m_bodyA->GetLinearVelocity(), vB = m_bodyB->GetLinearVelocity();
	float32 wA = m_bodyA->GetAngularVelocity(), wB = m_bodyB->GetAngularVelocity();
	return b2Dot(d, b2Cross(wA, axis)) + b2Dot(axis, vB + b2Cross(wB, rB) - vA - b2Cross(wA, rA));
}

float32 b2WheelJoint::GetJointAngle() const { return m_bodyB->GetAngle() - m_bodyA->GetAngle(); }
float32 b2WheelJoint::GetJointAngularSpeed() const { return m_bodyB->GetAngularVelocity() - m_bodyA->GetAngularVelocity(); }

void b2WheelJoint::EnableMotor(bool flag)
{
	if (flag == m_enableMotor) return;
	Touch();
	m_bodyA->SetAwake(true);
	m_bodyB->SetAwake(true);
	m_enableMotor = flag;
}

void b2WheelJoint::SetMotorSpeed(float32 speed)
{
	if (speed == m_motorSpeed) return;
	Touch();
	m_bodyA->SetAwake(true);
	m_bodyB->SetAwake(true);
	m_motorSpeed = speed;
}

void b2WheelJoint::SetMaxMotorTorque(float32 torque)
{
	if (torque == m_maxMotorTorque) return;
	Touch();
	m_bodyA->SetAwake(true);
	m_bodyB->SetAwake(true);
	m_maxMotorTorque = torque;
}

void b2WheelJoint::SetSpringFrequencyHz(float32 hz)
{
	if (hz == m_frequencyHz) return;
	Touch();
	m_frequencyHz = hz;
}

void b2WheelJoint::SetSpringDampingRatio(float32 ratio)
{
	if (ratio == m_dampingRatio) return;
	Touch();
	m_dampingRatio = ratio;
}

// ---- rope (reference b2RopeJoint.cpp:34-45, :197-228) --------------------------------------------------------------------

b2RopeJoint::b2RopeJoint(const b2RopeJointDef* def)
	: b2Joint(def), m_localAnchorA(def->localAnchorA), m_localAnchorB(def->localAnchorB), m_maxLength(def->maxLength),
	  m_impulse(0.0f), m_state(e_inactiveLimit), m_u(0.0f, 0.0f)
{
}

void b2RopeJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_ROPE, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->length = m_maxLength;
	out->impulse[0] = m_impulse;
	out->limitState = (int32_t)m_state;
	out->lastSolve[0] = m_u.x;
	out->lastSolve[1] = m_u.y;
}

void b2RopeJoint::ReadRecord(const b2cuJoint& in)
{
	m_impulse = in.impulse[0];
	m_state = (b2LimitState)in.limitState;
	m_u.Set(in.lastSolve[0], in.lastSolve[1]);
}

b2Vec2 b2RopeJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2RopeJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2RopeJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	float32 scale = inv_dt * m_impulse;
	return b2Vec2(scale * m_u.x, scale * m_u.y);
}

float32 b2RopeJoint::GetReactionTorque(float32 inv_dt) const
{
	B2_NOT_USED(inv_dt);
	return 0.0f;
}

b2LimitState b2RopeJoint::GetLimitState() const
{
	Refresh();
	return m_state;
}

void b2RopeJoint::SetMaxLength(float32 length)
{
	if (length == m_maxLength) return;
	Touch();
	m_maxLength = length;
}

// ---- friction (reference b2FrictionJoint.cpp:36-56, :192-236) ------------------------------------------------------------

void b2FrictionJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor)
{
	bodyA = bA;
	bodyB = bB;
	localAnchorA = bA->GetLocalPoint(anchor);
	localAnchorB = bB->GetLocalPoint(anchor);
}

b2FrictionJoint::b2FrictionJoint(const b2FrictionJointDef* def)
	: b2Joint(def), m_localAnchorA(def->localAnchorA), m_localAnchorB(def->localAnchorB), m_maxForce(def->maxForce),
	  m_maxTorque(def->maxTorque), m_linearImpulse(0.0f, 0.0f), m_angularImpulse(0.0f)
{
}

void b2FrictionJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_FRICTION, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->length = m_maxForce;
	out->maxMotorTorque = m_maxTorque;
	out->impulse[0] = m_linearImpulse.x;
	out->impulse[1] = m_linearImpulse.y;
	out->impulse[2] = m_angularImpulse;
}

void b2FrictionJoint::ReadRecord(const b2cuJoint& in)
{
	m_linearImpulse.Set(in.impulse[0], in.impulse[1]);
	m_angularImpulse = in.impulse[2];
}

b2Vec2 b2FrictionJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2FrictionJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2FrictionJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_linearImpulse;
}

float32 b2FrictionJoint::GetReactionTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_angularImpulse;
}

void b2FrictionJoint::SetMaxForce(float32 force)
{
	b2Assert(b2IsValid(force) && force >= 0.0f);
	if (force == m_maxForce) return;
	Touch();
	m_maxForce = force;
}

void b2FrictionJoint::SetMaxTorque(float32 torque)
{
	b2Assert(b2IsValid(torque) && torque >= 0.0f);
	if (torque == m_maxTorque) return;
	Touch();
	m_maxTorque = torque;
}

// ---- motor (reference b2MotorJoint.cpp:39-64, :202-300) ------------------------------------------------------------------

void b2MotorJointDef::Initialize(b2Body* bA, b2Body* bB)
{
	bodyA = bA;
	bodyB = bB;
	linearOffset = bA->GetLocalPoint(bB->GetPosition());
	angularOffset = bB->GetAngle() - bA->GetAngle();
}

b2MotorJoint::b2MotorJoint(const b2MotorJointDef* def)
	: b2Joint(def), m_linearOffset(def->linearOffset), m_angularOffset(def->angularOffset), m_maxForce(def->maxForce),
	  m_maxTorque(def->maxTorque), m_correctionFactor(def->correctionFactor), m_linearImpulse(0.0f, 0.0f), m_angularImpulse(0.0f)
{
}

void b2MotorJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_MOTOR, m_bodyA, m_bodyB, m_collideConnected, b2Vec2(0.0f, 0.0f), b2Vec2(0.0f, 0.0f));
	out->axis[0] = m_linearOffset.x;
	out->axis[1] = m_linearOffset.y;
	out->referenceAngle = m_angularOffset;
	out->length = m_maxForce;
	out->maxMotorTorque = m_maxTorque;
	out->dampingRatio = m_correctionFactor;
	out->impulse[0] = m_linearImpulse.x;
	out->impulse[1] = m_linearImpulse.y;
	out->impulse[2] = m_angularImpulse;
}

void b2MotorJoint::ReadRecord(const b2cuJoint& in)
{
	m_linearImpulse.Set(in.impulse[0], in.impulse[1]);
	m_angularImpulse = in.impulse[2];
}

b2Vec2 b2MotorJoint::GetAnchorA() const { return m_bodyA->GetPosition(); }
b2Vec2 b2MotorJoint::GetAnchorB() const { return m_bodyB->GetPosition(); }

b2Vec2 b2MotorJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_linearImpulse;
}

float32 b2MotorJoint::GetReactionTorque(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_angularImpulse;
}

void b2MotorJoint::SetLinearOffset(const b2Vec2& linearOffset)
{
	if (linearOffset.x == m_linearOffset.x && linearOffset.y == m_linearOffset.y) return;
	Touch();
	m_bodyA->SetAwake(true);
	m_bodyB->SetAwake(true);
	m_linearOffset = linearOffset;
}

void b2MotorJoint::SetAngularOffset(float32 angularOffset)
{
	if (angularOffset == m_angularOffset) return;
	Touch();
	m_bodyA->SetAwake(true);
	m_bodyB->SetAwake(true);
	m_angularOffset = angularOffset;
}

void b2MotorJoint::SetMaxForce(float32 force)
{
	b2Assert(b2IsValid(force) && force >= 0.0f);
	if (force == m_maxForce) return;
	Touch();
	m_maxForce = force;
}

void b2MotorJoint::SetMaxTorque(float32 torque)
{
	b2Assert(b2IsValid(torque) && torque >= 0.0f);
	if (torque == m_maxTorque) return;
	Touch();
	m_maxTorque = torque;
}

void b2MotorJoint::SetCorrectionFactor(float32 factor)
{
	b2Assert(b2IsValid(factor) && 0.0f <= factor && factor <= 1.0f);
	if (factor == m_correctionFactor) return;
	Touch();
	m_correctionFactor = factor;
}

// ---- pulley (reference b2PulleyJoint.cpp:36-72, :266-340) ----------------------------------------------------------------

void b2PulleyJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& groundA, const b2Vec2& groundB, const b2Vec2& anchorA,
                                  const b2Vec2& anchorB, float32 r)
{
	bodyA = bA;
	bodyB = bB;
	groundAnchorA = groundA;
	groundAnchorB = groundB;
	localAnchorA = bA->GetLocalPoint(anchorA);
	localAnchorB = bB->GetLocalPoint(anchorB);
	lengthA = (anchorA - groundA).Length();
	lengthB = (anchorB - groundB).Length();
	ratio = r;
	b2Assert(ratio > b2_epsilon);
}

b2PulleyJoint::b2PulleyJoint(const b2PulleyJointDef* def)
	: b2Joint(def), m_groundAnchorA(def->groundAnchorA), m_groundAnchorB(def->groundAnchorB), m_localAnchorA(def->localAnchorA),
	  m_localAnchorB(def->localAnchorB), m_lengthA(def->lengthA), m_lengthB(def->lengthB), m_ratio(def->ratio), m_impulse(0.0f),
	  m_uB(0.0f, 0.0f)
{
	b2Assert(def->ratio != 0.0f);
}

void b2PulleyJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_PULLEY, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->axis[0] = m_groundAnchorA.x;
	out->axis[1] = m_groundAnchorA.y;
	out->lowerAngle = m_groundAnchorB.x;
	out->upperAngle = m_groundAnchorB.y;
	out->length = m_lengthA;
	out->referenceAngle = m_lengthB;
	out->motorSpeed = m_ratio;
	out->impulse[0] = m_impulse;
	out->lastSolve[0] = m_uB.x;
	out->lastSolve[1] = m_uB.y;
}

void b2PulleyJoint::ReadRecord(const b2cuJoint& in)
{
	m_impulse = in.impulse[0];
	m_uB.Set(in.lastSolve[0], in.lastSolve[1]);
}

b2Vec2 b2PulleyJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2PulleyJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2PulleyJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	b2Vec2 P = m_impulse * m_uB;
	return inv_dt * P;
}

float32 b2PulleyJoint::GetReactionTorque(float32 inv_dt) const
{
	B2_NOT_USED(inv_dt);
	return 0.0f;
}

float32 b2PulleyJoint::GetCurrentLengthA() const { return (m_bodyA->GetWorldPoint(m_localAnchorA) - m_groundAnchorA).Length(); }
float32 b2PulleyJoint::GetCurrentLengthB() const { return (m_bodyB->GetWorldPoint(m_localAnchorB) - m_groundAnchorB).Length(); }

void b2PulleyJoint::ShiftOrigin(const b2Vec2& newOrigin)
{
	Touch();
	m_groundAnchorA -= newOrigin;
	m_groundAnchorB -= newOrigin;
}

// ---- mouse (reference b2MouseJoint.cpp:32-94, :192-215) ------------------------------------------------------------------

b2MouseJoint::b2MouseJoint(const b2MouseJointDef* def)
	: b2Joint(def), m_targetA(def->target), m_maxForce(def->maxForce), m_frequencyHz(def->frequencyHz),
	  m_dampingRatio(def->dampingRatio), m_impulse(0.0f, 0.0f)
{
	b2Assert(def->target.IsValid());
	m_localAnchorB = b2MulT(m_bodyB->GetTransform(), m_targetA);
}

void b2MouseJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_MOUSE, m_bodyA, m_bodyB, m_collideConnected, b2Vec2(0.0f, 0.0f), m_localAnchorB);
	out->axis[0] = m_targetA.x;
	out->axis[1] = m_targetA.y;
	out->length = m_maxForce;
	out->frequencyHz = m_frequencyHz;
	out->dampingRatio = m_dampingRatio;
	out->maxMotorTorque = m_bodyB->GetMass(); // the solve uses the mass itself, not the stored inverse
	out->impulse[0] = m_impulse.x;
	out->impulse[1] = m_impulse.y;
}

void b2MouseJoint::ReadRecord(const b2cuJoint& in) { m_impulse.Set(in.impulse[0], in.impulse[1]); }

b2Vec2 b2MouseJoint::GetAnchorA() const { return m_targetA; }
b2Vec2 b2MouseJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2MouseJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	return inv_dt * m_impulse;
}

float32 b2MouseJoint::GetReactionTorque(float32 inv_dt) const { return inv_dt * 0.0f; }

void b2MouseJoint::SetTarget(const b2Vec2& target)
{
	if (target.x == m_targetA.x && target.y == m_targetA.y) return;
	Touch();
	m_bodyB->SetAwake(true);
	m_targetA = target;
}

void b2MouseJoint::SetMaxForce(float32 force)
{
	if (force == m_maxForce) return;
	Touch();
	m_maxForce = force;
}

void b2MouseJoint::SetFrequency(float32 hz)
{
	if (hz == m_frequencyHz) return;
	Touch();
	m_frequencyHz = hz;
}

void b2MouseJoint::SetDampingRatio(float32 ratio)
{
	if (ratio == m_dampingRatio) return;
	Touch();
	m_dampingRatio = ratio;
}

void b2MouseJoint::ShiftOrigin(const b2Vec2& newOrigin)
{
	Touch();
	m_targetA -= newOrigin;
}

// ---- gear (reference b2GearJoint.cpp:45-129, :392-420) -------------------------------------------------------------------

// one side of the gear: anchors, axis and reference angle of a revolute / prismatic joint, and its current coordinate
static float32 GearSide(b2Joint* joint, b2JointType type, b2Body* bodyFirst, b2Body* bodySecond, b2Vec2* anchorFirst,
                        b2Vec2* anchorSecond, b2Vec2* axisFirst, float32* referenceAngle)
{
	const b2Transform& xfSecond = bodySecond->GetTransform();
	const b2Transform& xfFirst = bodyFirst->GetTransform();
	if (type == e_revoluteJoint)
	{
		b2RevoluteJoint* revolute = static_cast<b2RevoluteJoint*>(joint);
		*anchorFirst = revolute->GetLocalAnchorA();
		*anchorSecond = revolute->GetLocalAnchorB();
		*referenceAngle = revolute->GetReferenceAngle();
		axisFirst->SetZero();
		return bodySecond->GetAngle() - bodyFirst->GetAngle() - *referenceAngle;
	}
	b2PrismaticJoint* prismatic = static_cast<b2PrismaticJoint*>(joint);
	*anchorFirst = prismatic->GetLocalAnchorA();
	*anchorSecond = prismatic->GetLocalAnchorB();
	*referenceAngle = prismatic->GetReferenceAngle();
	*axisFirst = prismatic->GetLocalAxisA();
	b2Vec2 pFirst = *anchorFirst;
	b2Vec2 pSecond = b2MulT(xfFirst.q, b2Mul(xfSecond.q, *anchorSecond) + (xfSecond.p - xfFirst.p));
	return b2Dot(pSecond - pFirst, *axisFirst);
}

b2GearJoint::b2GearJoint(const b2GearJointDef* def)
	: b2Joint(def), m_joint1(def->joint1), m_joint2(def->joint2), m_typeA(def->joint1->GetType()), m_typeB(def->joint2->GetType()),
	  m_ratio(def->ratio), m_impulse(0.0f), m_JvAC(0.0f, 0.0f), m_JwA(0.0f)
{
	b2Assert(m_typeA == e_revoluteJoint || m_typeA == e_prismaticJoint);
	b2Assert(m_typeB == e_revoluteJoint || m_typeB == e_prismaticJoint);
	m_bodyC = m_joint1->GetBodyA();
	m_bodyA = m_joint1->GetBodyB();
	float32 coordinateA = GearSide(m_joint1, m_typeA, m_bodyC, m_bodyA, &m_localAnchorC, &m_localAnchorA, &m_localAxisC, &m_referenceAngleA);
	m_bodyD = m_joint2->GetBodyA();
	m_bodyB = m_joint2->GetBodyB();
	float32 coordinateB = GearSide(m_joint2, m_typeB, m_bodyD, m_bodyB, &m_localAnchorD, &m_localAnchorB, &m_localAxisD, &m_referenceAngleB);
	m_constant = coordinateA + m_ratio * coordinateB;
}

void b2GearJoint::WriteRecord(b2cuJoint* out) const
{
	WriteCommon(out, B2CU_JOINT_GEAR, m_bodyA, m_bodyB, m_collideConnected, m_localAnchorA, m_localAnchorB);
	out->flags |= (m_typeA == e_prismaticJoint ? B2CU_JOINT_GEAR_PRISMATIC_1 : 0u) | (m_typeB == e_prismaticJoint ? B2CU_JOINT_GEAR_PRISMATIC_2 : 0u);
	out->limitState = m_bodyC->GetIndex();
	out->reserved = m_bodyD->GetIndex();
	out->axis[0] = m_localAnchorC.x;
	out->axis[1] = m_localAnchorC.y;
	out->lowerAngle = m_localAnchorD.x;
	out->upperAngle = m_localAnchorD.y;
	out->work[0] = m_localAxisC.x;
	out->work[1] = m_localAxisC.y;
	out->work[2] = m_localAxisD.x;
	out->work[3] = m_localAxisD.y;
	out->referenceAngle = m_referenceAngleA;
	out->maxMotorTorque = m_referenceAngleB;
	out->motorSpeed = m_ratio;
	out->length = m_constant;
	out->frequencyHz = (float32)m_joint1->GetIndex();
	out->dampingRatio = (float32)m_joint2->GetIndex();
	out->impulse[0] = m_impulse;
	out->lastSolve[0] = m_JvAC.x;
	out->lastSolve[1] = m_JvAC.y;
	out->lastSolve[2] = m_JwA;
}

void b2GearJoint::ReadRecord(const b2cuJoint& in)
{
	m_impulse = in.impulse[0];
	m_JvAC.Set(in.lastSolve[0], in.lastSolve[1]);
	m_JwA = in.lastSolve[2];
}

b2Vec2 b2GearJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2GearJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }

b2Vec2 b2GearJoint::GetReactionForce(float32 inv_dt) const
{
	Refresh();
	b2Vec2 P = m_impulse * m_JvAC;
	return inv_dt * P;
}

float32 b2GearJoint::GetReactionTorque(float32 inv_dt) const
{
	Refresh();
	float32 L = m_impulse * m_JwA;
	return inv_dt * L;
}

void b2GearJoint::SetRatio(float32 ratio)
{
	b2Assert(b2IsValid(ratio));
	if (ratio == m_ratio) return;
	Touch();
	m_ratio = ratio;
}
