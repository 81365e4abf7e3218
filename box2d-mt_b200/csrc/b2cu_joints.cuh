// b2cu_joints.cuh -- joints as rows of the coloured solver.  Restates InitVelocityConstraints / SolveVelocityConstraints
// / SolvePositionConstraints of b2RevoluteJoint (Box2D/Dynamics/Joints/b2RevoluteJoint.cpp:64-400), b2DistanceJoint
// (b2DistanceJoint.cpp:63-222), b2WeldJoint (b2WeldJoint.cpp:59-308) and b2PrismaticJoint
// (b2PrismaticJoint.cpp:100-478), b2WheelJoint (b2WheelJoint.cpp:78-318), b2RopeJoint (b2RopeJoint.cpp:47-195),
// b2FrictionJoint (b2FrictionJoint.cpp:58-190), b2MotorJoint (b2MotorJoint.cpp:66-200), b2PulleyJoint
// (b2PulleyJoint.cpp:74-264), b2MouseJoint (b2MouseJoint.cpp:96-190) and b2GearJoint (b2GearJoint.cpp:131-390), and the
// small linear solves they use
// (b2Mat33::Solve33 / Solve22 / GetInverse22 / GetSymInverse33, Box2D/Common/b2Math.cpp:25-94; b2Mat22::Solve,
// b2Math.h:221-233) with the same fp32 arithmetic in the same order.  One thread solves one joint; joints of one colour class share no
// dynamic body, so a class is solved in parallel and the classes one after the other.
#pragma once

#include "b2cu_world.cuh"

namespace b2cu
{

struct Vec3
{
	float x, y, z;
};

__device__ __forceinline__ Vec3 V3(float x, float y, float z)
{
	Vec3 v;
	v.x = x;
	v.y = y;
	v.z = z;
	return v;
}
__device__ __forceinline__ float Dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3 Cross3(Vec3 a, Vec3 b)
{
	return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// per-step rows of a joint (b2RevoluteJoint's "solver temp" members)
struct JointRow
{
	Vec2 rA, rB;
	Vec2 localCenterA, localCenterB;
	float invMassA, invMassB, invIA, invIB;
	Vec3 ex, ey, ez; // m_mass (revolute: K; weld: its inverse)
	float motorMass; // revolute: motor mass; distance: m_mass
	Vec2 u;          // distance: unit vector from anchor A to anchor B; prismatic: m_localXAxisA (normalised)
	Vec2 axis, perp; // prismatic: m_axis, m_perp
	float a1, a2, s1, s2; // prismatic Jacobian terms
	float gamma, bias; // soft constraint terms (distance, weld)
	int root;   // island of the joint (position early exit)
	int solved; // in an awake island this step
};

__device__ __forceinline__ Vec3 Solve33(const JointRow& r, Vec3 b)
{
	float det = Dot3(r.ex, Cross3(r.ey, r.ez));
	if (det != 0.0f) det = 1.0f / det;
	Vec3 x;
	x.x = det * Dot3(b, Cross3(r.ey, r.ez));
	x.y = det * Dot3(r.ex, Cross3(b, r.ez));
	x.z = det * Dot3(r.ex, Cross3(r.ey, b));
	return x;
}

__device__ __forceinline__ Vec2 Solve22(float a11, float a12, float a21, float a22, Vec2 b)
{
	float det = a11 * a22 - a12 * a21;
	if (det != 0.0f) det = 1.0f / det;
	Vec2 x;
	x.x = det * (a22 * b.x - a12 * b.y);
	x.y = det * (a11 * b.y - a21 * b.x);
	return x;
}

__device__ __forceinline__ void StoreVelocity(const DeviceArrays& d, int body, float invMass, float invI, Vec2 v, float w,
                                              float keep)
{
	// a body without mass is not changed by an impulse; several joints of a class may share it
	if (invMass != 0.0f || invI != 0.0f) d.vel[body] = make_float4(v.x, v.y, w, keep);
}

// Is the joint part of this step's solve, and in which island?  The island search adds a joint from an island body when
// the body across is active (b2World.cpp:1286-1320).  Also loads the body constants every joint type starts from.
__device__ __forceinline__ bool JointPrepare(const DeviceArrays& d, const b2cuJoint& jt, JointRow& r)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	const uint32_t fA = d.bflags[bA], fB = d.bflags[bB];
	const bool staticA = (fA & B2CU_BODY_TYPE_MASK) == B2CU_STATIC_BODY, staticB = (fB & B2CU_BODY_TYPE_MASK) == B2CU_STATIC_BODY;
	const bool islandA = !staticA && (fA & B2CU_BODY_ISLAND), islandB = !staticB && (fB & B2CU_BODY_ISLAND);
	r.solved = (islandA || islandB) && (fA & B2CU_BODY_ACTIVE) && (fB & B2CU_BODY_ACTIVE) ? 1 : 0;
	r.root = islandA ? d.island[bA] : (islandB ? d.island[bB] : 0);
	if (!r.solved) return false;
	float4 massA = d.mass[bA], massB = d.mass[bB];
	r.localCenterA = V(massA.z, massA.w);
	r.localCenterB = V(massB.z, massB.w);
	r.invMassA = massA.x;
	r.invMassB = massB.x;
	r.invIA = massA.y;
	r.invIB = massB.y;
	r.u = r.axis = r.perp = V(0.0f, 0.0f);
	r.a1 = r.a2 = r.s1 = r.s2 = 0.0f;
	r.gamma = r.bias = r.motorMass = 0.0f;
	r.ex = r.ey = r.ez = V3(0.0f, 0.0f, 0.0f);
	return true;
}

// K of the point + angle constraint shared by the revolute and the weld joint
__device__ __forceinline__ void PointAngleK(JointRow& r, Vec2 rA, Vec2 rB)
{
	float mA = r.invMassA, mB = r.invMassB, iA = r.invIA, iB = r.invIB;
	r.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
	r.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
	r.ez.x = -rA.y * iA - rB.y * iB;
	r.ex.y = r.ey.x;
	r.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
	r.ez.y = rA.x * iA + rB.x * iB;
	r.ex.z = r.ez.x;
	r.ey.z = r.ez.y;
	r.ez.z = iA + iB;
}

// b2RevoluteJoint::InitVelocityConstraints (:64-183)
__device__ __forceinline__ void RevoluteInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	float aA = pA.z, aB = pB.z;
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	r.rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	r.rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);

	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;
	bool fixedRotation = (iA + iB == 0.0f);

	PointAngleK(r, r.rA, r.rB);

	r.motorMass = iA + iB;
	if (r.motorMass > 0.0f) r.motorMass = 1.0f / r.motorMass;

	const bool enableMotor = (jt.flags & B2CU_JOINT_ENABLE_MOTOR) != 0, enableLimit = (jt.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;
	if (!enableMotor || fixedRotation) jt.motorImpulse = 0.0f;

	if (enableLimit && !fixedRotation)
	{
		float jointAngle = aB - aA - jt.referenceAngle;
		if (Abs(jt.upperAngle - jt.lowerAngle) < 2.0f * B2CU_ANGULAR_SLOP)
		{
			jt.limitState = B2CU_LIMIT_EQUAL;
		}
		else if (jointAngle <= jt.lowerAngle)
		{
			if (jt.limitState != B2CU_LIMIT_AT_LOWER) jt.impulse[2] = 0.0f;
			jt.limitState = B2CU_LIMIT_AT_LOWER;
		}
		else if (jointAngle >= jt.upperAngle)
		{
			if (jt.limitState != B2CU_LIMIT_AT_UPPER) jt.impulse[2] = 0.0f;
			jt.limitState = B2CU_LIMIT_AT_UPPER;
		}
		else
		{
			jt.limitState = B2CU_LIMIT_INACTIVE;
			jt.impulse[2] = 0.0f;
		}
	}
	else
	{
		jt.limitState = B2CU_LIMIT_INACTIVE;
	}

	if (warmStarting)
	{
		// scale the impulses of the last step to this step's length
		jt.impulse[0] *= dtRatio;
		jt.impulse[1] *= dtRatio;
		jt.impulse[2] *= dtRatio;
		jt.motorImpulse *= dtRatio;

		Vec2 P = V(jt.impulse[0], jt.impulse[1]);
		vA = vA - mA * P;
		wA -= iA * (Cross(r.rA, P) + jt.motorImpulse + jt.impulse[2]);
		vB = vB + mB * P;
		wB += iB * (Cross(r.rB, P) + jt.motorImpulse + jt.impulse[2]);
	}
	else
	{
		jt.impulse[0] = jt.impulse[1] = jt.impulse[2] = 0.0f;
		jt.motorImpulse = 0.0f;
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

// b2RevoluteJoint::SolveVelocityConstraints (:185-294); h = the step's dt
__device__ __forceinline__ void RevoluteSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, b2cuJoint jt, float h)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;
	bool fixedRotation = (iA + iB == 0.0f);
	const bool enableMotor = (jt.flags & B2CU_JOINT_ENABLE_MOTOR) != 0, enableLimit = (jt.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;

	// motor
	if (enableMotor && jt.limitState != B2CU_LIMIT_EQUAL && !fixedRotation)
	{
		float Cdot = wB - wA - jt.motorSpeed;
		float impulse = -r.motorMass * Cdot;
		float oldImpulse = jt.motorImpulse;
		float maxImpulse = h * jt.maxMotorTorque;
		jt.motorImpulse = Clamp(jt.motorImpulse + impulse, -maxImpulse, maxImpulse);
		impulse = jt.motorImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if (enableLimit && jt.limitState != B2CU_LIMIT_INACTIVE && !fixedRotation)
	{
		// point constraint and angular limit together, 3x3
		Vec2 Cdot1 = vB + CrossSV(wB, r.rB) - vA - CrossSV(wA, r.rA);
		float Cdot2 = wB - wA;
		Vec3 s = Solve33(r, V3(Cdot1.x, Cdot1.y, Cdot2));
		Vec3 impulse = V3(-s.x, -s.y, -s.z);

		if (jt.limitState == B2CU_LIMIT_EQUAL)
		{
			jt.impulse[0] += impulse.x;
			jt.impulse[1] += impulse.y;
			jt.impulse[2] += impulse.z;
		}
		else
		{
			float newImpulse = jt.impulse[2] + impulse.z;
			bool release = jt.limitState == B2CU_LIMIT_AT_LOWER ? newImpulse < 0.0f : newImpulse > 0.0f;
			if (release)
			{
				// the limit would pull: drop its impulse and solve the point constraint alone
				Vec2 rhs = -Cdot1 + jt.impulse[2] * V(r.ez.x, r.ez.y);
				Vec2 reduced = Solve22(r.ex.x, r.ey.x, r.ex.y, r.ey.y, rhs);
				impulse.x = reduced.x;
				impulse.y = reduced.y;
				impulse.z = -jt.impulse[2];
				jt.impulse[0] += reduced.x;
				jt.impulse[1] += reduced.y;
				jt.impulse[2] = 0.0f;
			}
			else
			{
				jt.impulse[0] += impulse.x;
				jt.impulse[1] += impulse.y;
				jt.impulse[2] += impulse.z;
			}
		}

		Vec2 P = V(impulse.x, impulse.y);
		vA = vA - mA * P;
		wA -= iA * (Cross(r.rA, P) + impulse.z);
		vB = vB + mB * P;
		wB += iB * (Cross(r.rB, P) + impulse.z);
	}
	else
	{
		// point constraint, 2x2
		Vec2 Cdot = vB + CrossSV(wB, r.rB) - vA - CrossSV(wA, r.rA);
		Vec2 impulse = Solve22(r.ex.x, r.ey.x, r.ex.y, r.ey.y, -Cdot);
		jt.impulse[0] += impulse.x;
		jt.impulse[1] += impulse.y;
		vA = vA - mA * impulse;
		wA -= iA * Cross(r.rA, impulse);
		vB = vB + mB * impulse;
		wB += iB * Cross(r.rB, impulse);
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = jt.impulse[0];
	d.joints[j].impulse[1] = jt.impulse[1];
	d.joints[j].impulse[2] = jt.impulse[2];
	d.joints[j].motorImpulse = jt.motorImpulse;
}

// b2RevoluteJoint::SolvePositionConstraints (:296-377); returns "within tolerance"
__device__ __forceinline__ bool RevoluteSolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;

	float angularError = 0.0f, positionError = 0.0f;
	bool fixedRotation = (r.invIA + r.invIB == 0.0f);
	const bool enableLimit = (jt.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;

	if (enableLimit && jt.limitState != B2CU_LIMIT_INACTIVE && !fixedRotation)
	{
		float angle = aB - aA - jt.referenceAngle;
		float limitImpulse = 0.0f;
		if (jt.limitState == B2CU_LIMIT_EQUAL)
		{
			float C = Clamp(angle - jt.lowerAngle, -B2CU_MAX_ANGULAR_CORRECTION, B2CU_MAX_ANGULAR_CORRECTION);
			limitImpulse = -r.motorMass * C;
			angularError = Abs(C);
		}
		else if (jt.limitState == B2CU_LIMIT_AT_LOWER)
		{
			float C = angle - jt.lowerAngle;
			angularError = -C;
			C = Clamp(C + B2CU_ANGULAR_SLOP, -B2CU_MAX_ANGULAR_CORRECTION, 0.0f);
			limitImpulse = -r.motorMass * C;
		}
		else if (jt.limitState == B2CU_LIMIT_AT_UPPER)
		{
			float C = angle - jt.upperAngle;
			angularError = C;
			C = Clamp(C - B2CU_ANGULAR_SLOP, 0.0f, B2CU_MAX_ANGULAR_CORRECTION);
			limitImpulse = -r.motorMass * C;
		}
		aA -= r.invIA * limitImpulse;
		aB += r.invIB * limitImpulse;
	}

	{
		Rot qA = SinCos(aA), qB = SinCos(aB);
		Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
		Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);

		Vec2 C = cB + rB - cA - rA;
		positionError = Length(C);

		float mA = r.invMassA, mB = r.invMassB;
		float iA = r.invIA, iB = r.invIB;

		float exx = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
		float exy = -iA * rA.x * rA.y - iB * rB.x * rB.y;
		float eyx = exy;
		float eyy = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;

		Vec2 s = Solve22(exx, eyx, exy, eyy, C);
		Vec2 impulse = -s;

		cA = cA - mA * impulse;
		aA -= iA * Cross(rA, impulse);
		cB = cB + mB * impulse;
		aB += iB * Cross(rB, impulse);
	}

	if (r.invMassA != 0.0f || r.invIA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (r.invMassB != 0.0f || r.invIB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	return positionError <= B2CU_LINEAR_SLOP && angularError <= B2CU_ANGULAR_SLOP;
}

// ---- distance joint (b2DistanceJoint.cpp:63-222) ------------------------------------------------------------------------

__device__ __forceinline__ void DistanceInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting,
                                             float h)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	Rot qA = SinCos(pA.z), qB = SinCos(pB.z);
	r.rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	r.rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	r.u = cB + r.rB - cA - r.rA;

	// anchors on top of each other: no direction
	float length = Length(r.u);
	if (length > B2CU_LINEAR_SLOP) r.u = (1.0f / length) * r.u;
	else r.u = V(0.0f, 0.0f);

	float crAu = Cross(r.rA, r.u);
	float crBu = Cross(r.rB, r.u);
	float invMass = r.invMassA + r.invIA * crAu * crAu + r.invMassB + r.invIB * crBu * crBu;
	float mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;

	if (jt.frequencyHz > 0.0f)
	{
		// spring and damper folded into the constraint
		float C = length - jt.length;
		float omega = 2.0f * B2CU_PI * jt.frequencyHz;
		float dd = 2.0f * mass * jt.dampingRatio * omega;
		float k = mass * omega * omega;
		r.gamma = h * (dd + h * k);
		r.gamma = r.gamma != 0.0f ? 1.0f / r.gamma : 0.0f;
		r.bias = C * h * k * r.gamma;
		invMass += r.gamma;
		mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
	}
	else
	{
		r.gamma = 0.0f;
		r.bias = 0.0f;
	}
	r.motorMass = mass;

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		Vec2 P = jt.impulse[0] * r.u;
		vA = vA - r.invMassA * P;
		wA -= r.invIA * Cross(r.rA, P);
		vB = vB + r.invMassB * P;
		wB += r.invIB * Cross(r.rB, P);
	}
	else
	{
		jt.impulse[0] = 0.0f;
	}

	StoreVelocity(d, bA, r.invMassA, r.invIA, vA, wA, vA4.w);
	StoreVelocity(d, bB, r.invMassB, r.invIB, vB, wB, vB4.w);
	jt.lastSolve[0] = r.u.x; // b2DistanceJoint::GetReactionForce reads m_u
	jt.lastSolve[1] = r.u.y;
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void DistanceSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	Vec2 vpA = vA + CrossSV(wA, r.rA);
	Vec2 vpB = vB + CrossSV(wB, r.rB);
	float Cdot = Dot(r.u, vpB - vpA);

	float impulse = -r.motorMass * (Cdot + r.bias + r.gamma * jt.impulse[0]);
	float total = jt.impulse[0] + impulse;

	Vec2 P = impulse * r.u;
	vA = vA - r.invMassA * P;
	wA -= r.invIA * Cross(r.rA, P);
	vB = vB + r.invMassB * P;
	wB += r.invIB * Cross(r.rB, P);

	StoreVelocity(d, bA, r.invMassA, r.invIA, vA, wA, vA4.w);
	StoreVelocity(d, bB, r.invMassB, r.invIB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = total;
}

__device__ __forceinline__ bool DistanceSolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	// a soft distance constraint is not corrected in position
	if (jt.frequencyHz > 0.0f) return true;
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	Vec2 u = cB + rB - cA - rA;

	// b2Vec2::Normalize
	float length = Length(u);
	if (length < B2CU_EPSILON)
	{
		length = 0.0f;
	}
	else
	{
		float inv = 1.0f / length;
		u = V(u.x * inv, u.y * inv);
	}
	float C = length - jt.length;
	C = Clamp(C, -B2CU_MAX_LINEAR_CORRECTION, B2CU_MAX_LINEAR_CORRECTION);

	float impulse = -r.motorMass * C;
	Vec2 P = impulse * u;

	cA = cA - r.invMassA * P;
	aA -= r.invIA * Cross(rA, P);
	cB = cB + r.invMassB * P;
	aB += r.invIB * Cross(rB, P);

	if (r.invMassA != 0.0f || r.invIA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (r.invMassB != 0.0f || r.invIB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	return Abs(C) < B2CU_LINEAR_SLOP;
}

// ---- weld joint (b2WeldJoint.cpp:59-308) --------------------------------------------------------------------------------

// b2Mat33::GetInverse22 / GetSymInverse33 of the K in r.ex/ey/ez, in place
__device__ __forceinline__ void InvertK22(JointRow& r)
{
	float a = r.ex.x, b = r.ey.x, c = r.ex.y, dd = r.ey.y;
	float det = a * dd - b * c;
	if (det != 0.0f) det = 1.0f / det;
	r.ex = V3(det * dd, -det * c, 0.0f);
	r.ey = V3(-det * b, det * a, 0.0f);
	r.ez = V3(0.0f, 0.0f, 0.0f);
}

__device__ __forceinline__ void InvertKSym33(JointRow& r)
{
	float det = Dot3(r.ex, Cross3(r.ey, r.ez));
	if (det != 0.0f) det = 1.0f / det;
	float a11 = r.ex.x, a12 = r.ey.x, a13 = r.ez.x;
	float a22 = r.ey.y, a23 = r.ez.y;
	float a33 = r.ez.z;
	Vec3 ex, ey, ez;
	ex.x = det * (a22 * a33 - a23 * a23);
	ex.y = det * (a13 * a23 - a12 * a33);
	ex.z = det * (a12 * a23 - a13 * a22);
	ey.x = ex.y;
	ey.y = det * (a11 * a33 - a13 * a13);
	ey.z = det * (a13 * a12 - a11 * a23);
	ez.x = ex.z;
	ez.y = ey.z;
	ez.z = det * (a11 * a22 - a12 * a12);
	r.ex = ex;
	r.ey = ey;
	r.ez = ez;
}

__device__ __forceinline__ void WeldInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting,
                                         float h)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	float aA = pA.z, aB = pB.z;
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	r.rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	r.rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);

	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;
	PointAngleK(r, r.rA, r.rB);

	if (jt.frequencyHz > 0.0f)
	{
		InvertK22(r);
		float invM = iA + iB;
		float m = invM > 0.0f ? 1.0f / invM : 0.0f;
		float C = aB - aA - jt.referenceAngle;
		float omega = 2.0f * B2CU_PI * jt.frequencyHz;
		float dd = 2.0f * m * jt.dampingRatio * omega;
		float k = m * omega * omega;
		r.gamma = h * (dd + h * k);
		r.gamma = r.gamma != 0.0f ? 1.0f / r.gamma : 0.0f;
		r.bias = C * h * k * r.gamma;
		invM += r.gamma;
		r.ez.z = invM != 0.0f ? 1.0f / invM : 0.0f;
	}
	else if (r.ez.z == 0.0f)
	{
		InvertK22(r);
		r.gamma = 0.0f;
		r.bias = 0.0f;
	}
	else
	{
		InvertKSym33(r);
		r.gamma = 0.0f;
		r.bias = 0.0f;
	}

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		jt.impulse[1] *= dtRatio;
		jt.impulse[2] *= dtRatio;
		Vec2 P = V(jt.impulse[0], jt.impulse[1]);
		vA = vA - mA * P;
		wA -= iA * (Cross(r.rA, P) + jt.impulse[2]);
		vB = vB + mB * P;
		wB += iB * (Cross(r.rB, P) + jt.impulse[2]);
	}
	else
	{
		jt.impulse[0] = jt.impulse[1] = jt.impulse[2] = 0.0f;
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void WeldSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, b2cuJoint jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;

	if (jt.frequencyHz > 0.0f)
	{
		// soft angle first, then the point
		float Cdot2 = wB - wA;
		float impulse2 = -r.ez.z * (Cdot2 + r.bias + r.gamma * jt.impulse[2]);
		jt.impulse[2] += impulse2;
		wA -= iA * impulse2;
		wB += iB * impulse2;

		Vec2 Cdot1 = vB + CrossSV(wB, r.rB) - vA - CrossSV(wA, r.rA);
		Vec2 m22 = V(r.ex.x * Cdot1.x + r.ey.x * Cdot1.y, r.ex.y * Cdot1.x + r.ey.y * Cdot1.y);
		Vec2 impulse1 = -m22;
		jt.impulse[0] += impulse1.x;
		jt.impulse[1] += impulse1.y;

		Vec2 P = impulse1;
		vA = vA - mA * P;
		wA -= iA * Cross(r.rA, P);
		vB = vB + mB * P;
		wB += iB * Cross(r.rB, P);
	}
	else
	{
		Vec2 Cdot1 = vB + CrossSV(wB, r.rB) - vA - CrossSV(wA, r.rA);
		float Cdot2 = wB - wA;
		// b2Mul(m_mass, Cdot) = Cdot.x * ex + Cdot.y * ey + Cdot.z * ez
		Vec3 mv;
		mv.x = Cdot1.x * r.ex.x + Cdot1.y * r.ey.x + Cdot2 * r.ez.x;
		mv.y = Cdot1.x * r.ex.y + Cdot1.y * r.ey.y + Cdot2 * r.ez.y;
		mv.z = Cdot1.x * r.ex.z + Cdot1.y * r.ey.z + Cdot2 * r.ez.z;
		Vec3 impulse = V3(-mv.x, -mv.y, -mv.z);
		jt.impulse[0] += impulse.x;
		jt.impulse[1] += impulse.y;
		jt.impulse[2] += impulse.z;

		Vec2 P = V(impulse.x, impulse.y);
		vA = vA - mA * P;
		wA -= iA * (Cross(r.rA, P) + impulse.z);
		vB = vB + mB * P;
		wB += iB * (Cross(r.rB, P) + impulse.z);
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = jt.impulse[0];
	d.joints[j].impulse[1] = jt.impulse[1];
	d.joints[j].impulse[2] = jt.impulse[2];
}

__device__ __forceinline__ bool WeldSolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);

	float positionError, angularError;
	JointRow K = r;
	PointAngleK(K, rA, rB);

	if (jt.frequencyHz > 0.0f)
	{
		Vec2 C1 = cB + rB - cA - rA;
		positionError = Length(C1);
		angularError = 0.0f;
		Vec2 P = -Solve22(K.ex.x, K.ey.x, K.ex.y, K.ey.y, C1);
		cA = cA - mA * P;
		aA -= iA * Cross(rA, P);
		cB = cB + mB * P;
		aB += iB * Cross(rB, P);
	}
	else
	{
		Vec2 C1 = cB + rB - cA - rA;
		float C2 = aB - aA - jt.referenceAngle;
		positionError = Length(C1);
		angularError = Abs(C2);

		Vec3 impulse;
		if (K.ez.z > 0.0f)
		{
			Vec3 s = Solve33(K, V3(C1.x, C1.y, C2));
			impulse = V3(-s.x, -s.y, -s.z);
		}
		else
		{
			Vec2 s = -Solve22(K.ex.x, K.ey.x, K.ex.y, K.ey.y, C1);
			impulse = V3(s.x, s.y, 0.0f);
		}
		Vec2 P = V(impulse.x, impulse.y);
		cA = cA - mA * P;
		aA -= iA * (Cross(rA, P) + impulse.z);
		cB = cB + mB * P;
		aB += iB * (Cross(rB, P) + impulse.z);
	}

	if (mA != 0.0f || iA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (mB != 0.0f || iB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	return positionError <= B2CU_LINEAR_SLOP && angularError <= B2CU_ANGULAR_SLOP;
}

// ---- prismatic joint (b2PrismaticJoint.cpp:100-478) ---------------------------------------------------------------------

// b2PrismaticJoint::b2PrismaticJoint (:100-125): the local axis of the definition, normalised (b2Vec2::Normalize)
__device__ __forceinline__ Vec2 PrismaticLocalAxis(const b2cuJoint& jt)
{
	Vec2 a = V(jt.axis[0], jt.axis[1]);
	float length = Length(a);
	if (length < B2CU_EPSILON) return a;
	float inv = 1.0f / length;
	return V(a.x * inv, a.y * inv);
}

__device__ __forceinline__ void PrismaticInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	Rot qA = SinCos(pA.z), qB = SinCos(pB.z);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	Vec2 dd = (cB - cA) + rB - rA;

	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;

	r.u = PrismaticLocalAxis(jt);
	Vec2 localY = CrossSV(1.0f, r.u);

	// motor Jacobian and effective mass
	r.axis = Mul(qA, r.u);
	r.a1 = Cross(dd + rA, r.axis);
	r.a2 = Cross(rB, r.axis);
	r.motorMass = mA + mB + iA * r.a1 * r.a1 + iB * r.a2 * r.a2;
	if (r.motorMass > 0.0f) r.motorMass = 1.0f / r.motorMass;

	// the prismatic constraint: no motion along perp, no relative rotation
	r.perp = Mul(qA, localY);
	r.s1 = Cross(dd + rA, r.perp);
	r.s2 = Cross(rB, r.perp);
	{
		float k11 = mA + mB + iA * r.s1 * r.s1 + iB * r.s2 * r.s2;
		float k12 = iA * r.s1 + iB * r.s2;
		float k13 = iA * r.s1 * r.a1 + iB * r.s2 * r.a2;
		float k22 = iA + iB;
		if (k22 == 0.0f) k22 = 1.0f; // bodies with fixed rotation
		float k23 = iA * r.a1 + iB * r.a2;
		float k33 = mA + mB + iA * r.a1 * r.a1 + iB * r.a2 * r.a2;
		r.ex = V3(k11, k12, k13);
		r.ey = V3(k12, k22, k23);
		r.ez = V3(k13, k23, k33);
	}

	const bool enableMotor = (jt.flags & B2CU_JOINT_ENABLE_MOTOR) != 0, enableLimit = (jt.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;
	if (enableLimit)
	{
		float jointTranslation = Dot(r.axis, dd);
		if (Abs(jt.upperAngle - jt.lowerAngle) < 2.0f * B2CU_LINEAR_SLOP)
		{
			jt.limitState = B2CU_LIMIT_EQUAL;
		}
		else if (jointTranslation <= jt.lowerAngle)
		{
			if (jt.limitState != B2CU_LIMIT_AT_LOWER)
			{
				jt.limitState = B2CU_LIMIT_AT_LOWER;
				jt.impulse[2] = 0.0f;
			}
		}
		else if (jointTranslation >= jt.upperAngle)
		{
			if (jt.limitState != B2CU_LIMIT_AT_UPPER)
			{
				jt.limitState = B2CU_LIMIT_AT_UPPER;
				jt.impulse[2] = 0.0f;
			}
		}
		else
		{
			jt.limitState = B2CU_LIMIT_INACTIVE;
			jt.impulse[2] = 0.0f;
		}
	}
	else
	{
		jt.limitState = B2CU_LIMIT_INACTIVE;
		jt.impulse[2] = 0.0f;
	}

	if (!enableMotor) jt.motorImpulse = 0.0f;

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		jt.impulse[1] *= dtRatio;
		jt.impulse[2] *= dtRatio;
		jt.motorImpulse *= dtRatio;

		Vec2 P = jt.impulse[0] * r.perp + (jt.motorImpulse + jt.impulse[2]) * r.axis;
		float LA = jt.impulse[0] * r.s1 + jt.impulse[1] + (jt.motorImpulse + jt.impulse[2]) * r.a1;
		float LB = jt.impulse[0] * r.s2 + jt.impulse[1] + (jt.motorImpulse + jt.impulse[2]) * r.a2;
		vA = vA - mA * P;
		wA -= iA * LA;
		vB = vB + mB * P;
		wB += iB * LB;
	}
	else
	{
		jt.impulse[0] = jt.impulse[1] = jt.impulse[2] = 0.0f;
		jt.motorImpulse = 0.0f;
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	jt.lastSolve[0] = r.axis.x; // b2PrismaticJoint::GetReactionForce reads m_axis and m_perp
	jt.lastSolve[1] = r.axis.y;
	jt.lastSolve[2] = r.perp.x;
	jt.lastSolve[3] = r.perp.y;
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void PrismaticSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, b2cuJoint jt, float h)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;
	const bool enableMotor = (jt.flags & B2CU_JOINT_ENABLE_MOTOR) != 0, enableLimit = (jt.flags & B2CU_JOINT_ENABLE_LIMIT) != 0;

	// linear motor
	if (enableMotor && jt.limitState != B2CU_LIMIT_EQUAL)
	{
		float Cdot = Dot(r.axis, vB - vA) + r.a2 * wB - r.a1 * wA;
		float impulse = r.motorMass * (jt.motorSpeed - Cdot);
		float oldImpulse = jt.motorImpulse;
		float maxImpulse = h * jt.maxMotorTorque;
		jt.motorImpulse = Clamp(jt.motorImpulse + impulse, -maxImpulse, maxImpulse);
		impulse = jt.motorImpulse - oldImpulse;

		Vec2 P = impulse * r.axis;
		float LA = impulse * r.a1;
		float LB = impulse * r.a2;
		vA = vA - mA * P;
		wA -= iA * LA;
		vB = vB + mB * P;
		wB += iB * LB;
	}

	Vec2 Cdot1;
	Cdot1.x = Dot(r.perp, vB - vA) + r.s2 * wB - r.s1 * wA;
	Cdot1.y = wB - wA;

	if (enableLimit && jt.limitState != B2CU_LIMIT_INACTIVE)
	{
		// prismatic constraint and limit together, 3x3
		float Cdot2 = Dot(r.axis, vB - vA) + r.a2 * wB - r.a1 * wA;
		Vec3 f1 = V3(jt.impulse[0], jt.impulse[1], jt.impulse[2]);
		Vec3 df = Solve33(r, V3(-Cdot1.x, -Cdot1.y, -Cdot2));
		Vec3 imp = V3(f1.x + df.x, f1.y + df.y, f1.z + df.z);

		if (jt.limitState == B2CU_LIMIT_AT_LOWER) imp.z = Max(imp.z, 0.0f);
		else if (jt.limitState == B2CU_LIMIT_AT_UPPER) imp.z = Min(imp.z, 0.0f);

		// with the limit impulse clamped, re-solve the other two rows
		Vec2 b = -Cdot1 - (imp.z - f1.z) * V(r.ez.x, r.ez.y);
		Vec2 f2r = Solve22(r.ex.x, r.ey.x, r.ex.y, r.ey.y, b) + V(f1.x, f1.y);
		imp.x = f2r.x;
		imp.y = f2r.y;

		df = V3(imp.x - f1.x, imp.y - f1.y, imp.z - f1.z);
		jt.impulse[0] = imp.x;
		jt.impulse[1] = imp.y;
		jt.impulse[2] = imp.z;

		Vec2 P = df.x * r.perp + df.z * r.axis;
		float LA = df.x * r.s1 + df.y + df.z * r.a1;
		float LB = df.x * r.s2 + df.y + df.z * r.a2;
		vA = vA - mA * P;
		wA -= iA * LA;
		vB = vB + mB * P;
		wB += iB * LB;
	}
	else
	{
		Vec2 df = Solve22(r.ex.x, r.ey.x, r.ex.y, r.ey.y, -Cdot1);
		jt.impulse[0] += df.x;
		jt.impulse[1] += df.y;

		Vec2 P = df.x * r.perp;
		float LA = df.x * r.s1 + df.y;
		float LB = df.x * r.s2 + df.y;
		vA = vA - mA * P;
		wA -= iA * LA;
		vB = vB + mB * P;
		wB += iB * LB;
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = jt.impulse[0];
	d.joints[j].impulse[1] = jt.impulse[1];
	d.joints[j].impulse[2] = jt.impulse[2];
	d.joints[j].motorImpulse = jt.motorImpulse;
}

__device__ __forceinline__ bool PrismaticSolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	Vec2 dd = cB + rB - cA - rA;

	Vec2 axis = Mul(qA, r.u);
	float a1 = Cross(dd + rA, axis);
	float a2 = Cross(rB, axis);
	Vec2 perp = Mul(qA, CrossSV(1.0f, r.u));
	float s1 = Cross(dd + rA, perp);
	float s2 = Cross(rB, perp);

	Vec3 impulse;
	Vec2 C1;
	C1.x = Dot(perp, dd);
	C1.y = aB - aA - jt.referenceAngle;

	float linearError = Abs(C1.x);
	float angularError = Abs(C1.y);

	bool active = false;
	float C2 = 0.0f;
	if (jt.flags & B2CU_JOINT_ENABLE_LIMIT)
	{
		float translation = Dot(axis, dd);
		if (Abs(jt.upperAngle - jt.lowerAngle) < 2.0f * B2CU_LINEAR_SLOP)
		{
			C2 = Clamp(translation, -B2CU_MAX_LINEAR_CORRECTION, B2CU_MAX_LINEAR_CORRECTION);
			linearError = Max(linearError, Abs(translation));
			active = true;
		}
		else if (translation <= jt.lowerAngle)
		{
			C2 = Clamp(translation - jt.lowerAngle + B2CU_LINEAR_SLOP, -B2CU_MAX_LINEAR_CORRECTION, 0.0f);
			linearError = Max(linearError, jt.lowerAngle - translation);
			active = true;
		}
		else if (translation >= jt.upperAngle)
		{
			C2 = Clamp(translation - jt.upperAngle - B2CU_LINEAR_SLOP, 0.0f, B2CU_MAX_LINEAR_CORRECTION);
			linearError = Max(linearError, translation - jt.upperAngle);
			active = true;
		}
	}

	if (active)
	{
		float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
		float k12 = iA * s1 + iB * s2;
		float k13 = iA * s1 * a1 + iB * s2 * a2;
		float k22 = iA + iB;
		if (k22 == 0.0f) k22 = 1.0f;
		float k23 = iA * a1 + iB * a2;
		float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
		JointRow K = r;
		K.ex = V3(k11, k12, k13);
		K.ey = V3(k12, k22, k23);
		K.ez = V3(k13, k23, k33);
		impulse = Solve33(K, V3(-C1.x, -C1.y, -C2));
	}
	else
	{
		float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
		float k12 = iA * s1 + iB * s2;
		float k22 = iA + iB;
		if (k22 == 0.0f) k22 = 1.0f;
		Vec2 impulse1 = Solve22(k11, k12, k12, k22, -C1);
		impulse = V3(impulse1.x, impulse1.y, 0.0f);
	}

	Vec2 P = impulse.x * perp + impulse.z * axis;
	float LA = impulse.x * s1 + impulse.y + impulse.z * a1;
	float LB = impulse.x * s2 + impulse.y + impulse.z * a2;

	cA = cA - mA * P;
	aA -= iA * LA;
	cB = cB + mB * P;
	aB += iB * LB;

	if (mA != 0.0f || iA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (mB != 0.0f || iB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	return linearError <= B2CU_LINEAR_SLOP && angularError <= B2CU_ANGULAR_SLOP;
}

// ---- wheel joint (b2WheelJoint.cpp:78-318) ------------------------------------------------------------------------------
// Row use: axis = m_ax, perp = m_ay, a1 = m_sAx, a2 = m_sBx, s1 = m_sAy, s2 = m_sBy, motorMass = m_motorMass, gamma / bias,
// ex.x = m_mass, ex.y = m_springMass.  With the spring off the reference leaves m_ax, m_sAx and m_sBx at whatever they were
// and still runs the spring row with them (its mass is 0, so only signs of zero are at stake): they persist in
// lastSolve[0..1] and work[0..1].

__device__ __forceinline__ void WheelInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting,
                                          float h)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;

	Rot qA = SinCos(pA.z), qB = SinCos(pB.z);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	Vec2 dd = cB + rB - cA - rA;
	Vec2 localX = V(jt.axis[0], jt.axis[1]); // not normalised by the reference either
	Vec2 localY = CrossSV(1.0f, localX);

	// point on line
	r.perp = Mul(qA, localY);
	r.s1 = Cross(dd + rA, r.perp);
	r.s2 = Cross(rB, r.perp);
	float mass = mA + mB + iA * r.s1 * r.s1 + iB * r.s2 * r.s2;
	if (mass > 0.0f) mass = 1.0f / mass;

	// spring
	float springMass = 0.0f;
	r.bias = 0.0f;
	r.gamma = 0.0f;
	r.axis = V(jt.lastSolve[0], jt.lastSolve[1]);
	r.a1 = jt.work[0];
	r.a2 = jt.work[1];
	if (jt.frequencyHz > 0.0f)
	{
		r.axis = Mul(qA, localX);
		r.a1 = Cross(dd + rA, r.axis);
		r.a2 = Cross(rB, r.axis);
		float invMass = mA + mB + iA * r.a1 * r.a1 + iB * r.a2 * r.a2;
		if (invMass > 0.0f)
		{
			springMass = 1.0f / invMass;
			float C = Dot(dd, r.axis);
			float omega = 2.0f * B2CU_PI * jt.frequencyHz;
			float damp = 2.0f * springMass * jt.dampingRatio * omega;
			float k = springMass * omega * omega;
			r.gamma = h * (damp + h * k);
			if (r.gamma > 0.0f) r.gamma = 1.0f / r.gamma;
			r.bias = C * h * k * r.gamma;
			springMass = invMass + r.gamma;
			if (springMass > 0.0f) springMass = 1.0f / springMass;
		}
	}
	else
	{
		jt.impulse[1] = 0.0f;
	}

	// motor
	if (jt.flags & B2CU_JOINT_ENABLE_MOTOR)
	{
		r.motorMass = iA + iB;
		if (r.motorMass > 0.0f) r.motorMass = 1.0f / r.motorMass;
	}
	else
	{
		r.motorMass = 0.0f;
		jt.motorImpulse = 0.0f;
	}
	r.ex = V3(mass, springMass, 0.0f);

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		jt.impulse[1] *= dtRatio;
		jt.motorImpulse *= dtRatio;
		Vec2 P = jt.impulse[0] * r.perp + jt.impulse[1] * r.axis;
		float LA = jt.impulse[0] * r.s1 + jt.impulse[1] * r.a1 + jt.motorImpulse;
		float LB = jt.impulse[0] * r.s2 + jt.impulse[1] * r.a2 + jt.motorImpulse;
		vA = vA - mA * P;
		wA -= iA * LA;
		vB = vB + mB * P;
		wB += iB * LB;
	}
	else
	{
		jt.impulse[0] = 0.0f;
		jt.impulse[1] = 0.0f;
		jt.motorImpulse = 0.0f;
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	jt.lastSolve[0] = r.axis.x;
	jt.lastSolve[1] = r.axis.y;
	jt.lastSolve[2] = r.perp.x;
	jt.lastSolve[3] = r.perp.y;
	jt.work[0] = r.a1;
	jt.work[1] = r.a2;
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void WheelSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, b2cuJoint jt, float h)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;

	// spring
	{
		float Cdot = Dot(r.axis, vB - vA) + r.a2 * wB - r.a1 * wA;
		float impulse = -r.ex.y * (Cdot + r.bias + r.gamma * jt.impulse[1]);
		jt.impulse[1] += impulse;
		Vec2 P = impulse * r.axis;
		float LA = impulse * r.a1;
		float LB = impulse * r.a2;
		vA = vA - mA * P;
		wA -= iA * LA;
		vB = vB + mB * P;
		wB += iB * LB;
	}
	// motor
	{
		float Cdot = wB - wA - jt.motorSpeed;
		float impulse = -r.motorMass * Cdot;
		float oldImpulse = jt.motorImpulse;
		float maxImpulse = h * jt.maxMotorTorque;
		jt.motorImpulse = Clamp(jt.motorImpulse + impulse, -maxImpulse, maxImpulse);
		impulse = jt.motorImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}
	// point on line
	{
		float Cdot = Dot(r.perp, vB - vA) + r.s2 * wB - r.s1 * wA;
		float impulse = -r.ex.x * Cdot;
		jt.impulse[0] += impulse;
		Vec2 P = impulse * r.perp;
		float LA = impulse * r.s1;
		float LB = impulse * r.s2;
		vA = vA - mA * P;
		wA -= iA * LA;
		vB = vB + mB * P;
		wB += iB * LB;
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = jt.impulse[0];
	d.joints[j].impulse[1] = jt.impulse[1];
	d.joints[j].motorImpulse = jt.motorImpulse;
}

__device__ __forceinline__ bool WheelSolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	Vec2 dd = (cB - cA) + rB - rA;
	Vec2 ay = Mul(qA, CrossSV(1.0f, V(jt.axis[0], jt.axis[1])));

	float sAy = Cross(dd + rA, ay);
	float sBy = Cross(rB, ay);
	float C = Dot(dd, ay);

	// the reference uses the Jacobian terms of the velocity solve here (m_sAy, m_sBy), not the fresh ones
	float k = r.invMassA + r.invMassB + r.invIA * r.s1 * r.s1 + r.invIB * r.s2 * r.s2;
	float impulse = k != 0.0f ? -C / k : 0.0f;

	Vec2 P = impulse * ay;
	float LA = impulse * sAy;
	float LB = impulse * sBy;
	cA = cA - r.invMassA * P;
	aA -= r.invIA * LA;
	cB = cB + r.invMassB * P;
	aB += r.invIB * LB;

	if (r.invMassA != 0.0f || r.invIA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (r.invMassB != 0.0f || r.invIB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	return Abs(C) <= B2CU_LINEAR_SLOP;
}

// ---- rope joint (b2RopeJoint.cpp:47-195) --------------------------------------------------------------------------------
// Row use: u = m_u, motorMass = m_mass, bias = m_length.  limitState = m_state.

__device__ __forceinline__ void RopeInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	Rot qA = SinCos(pA.z), qB = SinCos(pB.z);
	r.rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	r.rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	r.u = cB + r.rB - cA - r.rA;
	r.bias = Length(r.u);

	float C = r.bias - jt.length;
	jt.limitState = C > 0.0f ? B2CU_LIMIT_AT_UPPER : B2CU_LIMIT_INACTIVE;

	if (r.bias > B2CU_LINEAR_SLOP)
	{
		r.u = (1.0f / r.bias) * r.u;
	}
	else
	{
		// slack and degenerate: nothing to apply, the velocities stay as they are
		r.u = V(0.0f, 0.0f);
		r.motorMass = 0.0f;
		jt.impulse[0] = 0.0f;
		jt.lastSolve[0] = jt.lastSolve[1] = 0.0f;
		d.jointRows[j] = r;
		d.joints[j] = jt;
		return;
	}

	float crA = Cross(r.rA, r.u);
	float crB = Cross(r.rB, r.u);
	float invMass = r.invMassA + r.invIA * crA * crA + r.invMassB + r.invIB * crB * crB;
	r.motorMass = invMass != 0.0f ? 1.0f / invMass : 0.0f;

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		Vec2 P = jt.impulse[0] * r.u;
		vA = vA - r.invMassA * P;
		wA -= r.invIA * Cross(r.rA, P);
		vB = vB + r.invMassB * P;
		wB += r.invIB * Cross(r.rB, P);
	}
	else
	{
		jt.impulse[0] = 0.0f;
	}

	StoreVelocity(d, bA, r.invMassA, r.invIA, vA, wA, vA4.w);
	StoreVelocity(d, bB, r.invMassB, r.invIB, vB, wB, vB4.w);
	jt.lastSolve[0] = r.u.x;
	jt.lastSolve[1] = r.u.y;
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void RopeSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, const b2cuJoint& jt, float h)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;

	Vec2 vpA = vA + CrossSV(wA, r.rA);
	Vec2 vpB = vB + CrossSV(wB, r.rB);
	float C = r.bias - jt.length;
	float Cdot = Dot(r.u, vpB - vpA);
	// the rope is still slack: let it run out exactly this step
	if (C < 0.0f) Cdot += (1.0f / h) * C;

	float impulse = -r.motorMass * Cdot;
	float oldImpulse = jt.impulse[0];
	float total = Min(0.0f, oldImpulse + impulse);
	impulse = total - oldImpulse;

	Vec2 P = impulse * r.u;
	vA = vA - r.invMassA * P;
	wA -= r.invIA * Cross(r.rA, P);
	vB = vB + r.invMassB * P;
	wB += r.invIB * Cross(r.rB, P);

	StoreVelocity(d, bA, r.invMassA, r.invIA, vA, wA, vA4.w);
	StoreVelocity(d, bB, r.invMassB, r.invIB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = total;
}

__device__ __forceinline__ bool RopeSolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	Vec2 u = cB + rB - cA - rA;

	float length = Length(u);
	if (length < B2CU_EPSILON)
	{
		length = 0.0f;
	}
	else
	{
		float inv = 1.0f / length;
		u = V(u.x * inv, u.y * inv);
	}
	float C = length - jt.length;
	C = Clamp(C, 0.0f, B2CU_MAX_LINEAR_CORRECTION);

	float impulse = -r.motorMass * C;
	Vec2 P = impulse * u;
	cA = cA - r.invMassA * P;
	aA -= r.invIA * Cross(rA, P);
	cB = cB + r.invMassB * P;
	aB += r.invIB * Cross(rB, P);

	if (r.invMassA != 0.0f || r.invIA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (r.invMassB != 0.0f || r.invIB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	return length - jt.length < B2CU_LINEAR_SLOP;
}

// ---- friction joint (b2FrictionJoint.cpp:58-190) and motor joint (b2MotorJoint.cpp:66-200) -------------------------------
// Both apply a bounded linear and a bounded angular impulse; the motor joint drives towards an offset instead of rest.
// Row use: ex.x, ex.y, ey.x, ey.y = m_linearMass (ex, ey columns), motorMass = m_angularMass, u = m_linearError,
// bias = m_angularError.  Record: length = maxForce, maxMotorTorque = maxTorque; motor joint: axis = linearOffset,
// referenceAngle = angularOffset, dampingRatio = correctionFactor.  impulse[0..1] = m_linearImpulse, impulse[2] = m_angularImpulse.

__device__ __forceinline__ void FrictionMotorInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio,
                                                  int warmStarting, bool motor)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	if (motor)
	{
		r.rA = Mul(qA, V(jt.axis[0], jt.axis[1]) - r.localCenterA);
		r.rB = Mul(qB, -r.localCenterB);
	}
	else
	{
		r.rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
		r.rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	}

	float kxx = mA + mB + iA * r.rA.y * r.rA.y + iB * r.rB.y * r.rB.y;
	float kxy = -iA * r.rA.x * r.rA.y - iB * r.rB.x * r.rB.y;
	float kyy = mA + mB + iA * r.rA.x * r.rA.x + iB * r.rB.x * r.rB.x;
	{
		// b2Mat22::GetInverse of ex = (kxx, kxy), ey = (kxy, kyy)
		float a = kxx, b = kxy, c = kxy, dd = kyy;
		float det = a * dd - b * c;
		if (det != 0.0f) det = 1.0f / det;
		r.ex = V3(det * dd, -det * c, 0.0f);
		r.ey = V3(-det * b, det * a, 0.0f);
	}
	r.motorMass = iA + iB;
	if (r.motorMass > 0.0f) r.motorMass = 1.0f / r.motorMass;

	if (motor)
	{
		r.u = cB + r.rB - cA - r.rA;
		r.bias = aB - aA - jt.referenceAngle;
	}

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		jt.impulse[1] *= dtRatio;
		jt.impulse[2] *= dtRatio;
		Vec2 P = V(jt.impulse[0], jt.impulse[1]);
		vA = vA - mA * P;
		wA -= iA * (Cross(r.rA, P) + jt.impulse[2]);
		vB = vB + mB * P;
		wB += iB * (Cross(r.rB, P) + jt.impulse[2]);
	}
	else
	{
		jt.impulse[0] = jt.impulse[1] = jt.impulse[2] = 0.0f;
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void FrictionMotorSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, b2cuJoint jt, float h,
                                                           bool motor)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	float mA = r.invMassA, mB = r.invMassB;
	float iA = r.invIA, iB = r.invIB;
	const float inv_h = 1.0f / h;

	// angular part
	{
		float Cdot = wB - wA;
		if (motor) Cdot = wB - wA + inv_h * jt.dampingRatio * r.bias;
		float impulse = -r.motorMass * Cdot;
		float oldImpulse = jt.impulse[2];
		float maxImpulse = h * jt.maxMotorTorque;
		jt.impulse[2] = Clamp(oldImpulse + impulse, -maxImpulse, maxImpulse);
		impulse = jt.impulse[2] - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}
	// linear part
	{
		Vec2 Cdot = vB + CrossSV(wB, r.rB) - vA - CrossSV(wA, r.rA);
		if (motor) Cdot = Cdot + (inv_h * jt.dampingRatio) * r.u;
		Vec2 mv = V(r.ex.x * Cdot.x + r.ey.x * Cdot.y, r.ex.y * Cdot.x + r.ey.y * Cdot.y);
		Vec2 impulse = -mv;
		Vec2 oldImpulse = V(jt.impulse[0], jt.impulse[1]);
		Vec2 total = oldImpulse + impulse;
		float maxImpulse = h * jt.length;
		if (Dot(total, total) > maxImpulse * maxImpulse)
		{
			// b2Vec2::Normalize, then scale
			float length = Length(total);
			if (!(length < B2CU_EPSILON))
			{
				float inv = 1.0f / length;
				total = V(total.x * inv, total.y * inv);
			}
			total = V(total.x * maxImpulse, total.y * maxImpulse);
		}
		impulse = total - oldImpulse;
		jt.impulse[0] = total.x;
		jt.impulse[1] = total.y;
		vA = vA - mA * impulse;
		wA -= iA * Cross(r.rA, impulse);
		vB = vB + mB * impulse;
		wB += iB * Cross(r.rB, impulse);
	}

	StoreVelocity(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity(d, bB, mB, iB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = jt.impulse[0];
	d.joints[j].impulse[1] = jt.impulse[1];
	d.joints[j].impulse[2] = jt.impulse[2];
}

// ---- pulley joint (b2PulleyJoint.cpp:74-264) ----------------------------------------------------------------------------
// Record: axis = groundAnchorA, (lowerAngle, upperAngle) = groundAnchorB, length = lengthA, referenceAngle = lengthB,
// motorSpeed = ratio.  Row use: u = m_uA, axis = m_uB, motorMass = m_mass.

__device__ __forceinline__ Vec2 PulleyDirection(Vec2 u)
{
	float length = Length(u);
	if (length > 10.0f * B2CU_LINEAR_SLOP) return (1.0f / length) * u;
	return V(0.0f, 0.0f);
}

__device__ __forceinline__ void PulleyInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	const float ratio = jt.motorSpeed;

	Rot qA = SinCos(pA.z), qB = SinCos(pB.z);
	r.rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	r.rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	r.u = PulleyDirection(cA + r.rA - V(jt.axis[0], jt.axis[1]));
	r.axis = PulleyDirection(cB + r.rB - V(jt.lowerAngle, jt.upperAngle));

	float ruA = Cross(r.rA, r.u);
	float ruB = Cross(r.rB, r.axis);
	float mA = r.invMassA + r.invIA * ruA * ruA;
	float mB = r.invMassB + r.invIB * ruB * ruB;
	r.motorMass = mA + ratio * ratio * mB;
	if (r.motorMass > 0.0f) r.motorMass = 1.0f / r.motorMass;

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		Vec2 PA = -(jt.impulse[0]) * r.u;
		Vec2 PB = (-ratio * jt.impulse[0]) * r.axis;
		vA = vA + r.invMassA * PA;
		wA += r.invIA * Cross(r.rA, PA);
		vB = vB + r.invMassB * PB;
		wB += r.invIB * Cross(r.rB, PB);
	}
	else
	{
		jt.impulse[0] = 0.0f;
	}

	StoreVelocity(d, bA, r.invMassA, r.invIA, vA, wA, vA4.w);
	StoreVelocity(d, bB, r.invMassB, r.invIB, vB, wB, vB4.w);
	jt.lastSolve[0] = r.axis.x; // b2PulleyJoint::GetReactionForce reads m_uB
	jt.lastSolve[1] = r.axis.y;
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void PulleySolveVelocity(const DeviceArrays& d, int j, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y);
	float wA = vA4.z, wB = vB4.z;
	const float ratio = jt.motorSpeed;

	Vec2 vpA = vA + CrossSV(wA, r.rA);
	Vec2 vpB = vB + CrossSV(wB, r.rB);
	float Cdot = -Dot(r.u, vpA) - ratio * Dot(r.axis, vpB);
	float impulse = -r.motorMass * Cdot;
	float total = jt.impulse[0] + impulse;

	Vec2 PA = -impulse * r.u;
	Vec2 PB = -ratio * impulse * r.axis;
	vA = vA + r.invMassA * PA;
	wA += r.invIA * Cross(r.rA, PA);
	vB = vB + r.invMassB * PB;
	wB += r.invIB * Cross(r.rB, PB);

	StoreVelocity(d, bA, r.invMassA, r.invIA, vA, wA, vA4.w);
	StoreVelocity(d, bB, r.invMassB, r.invIB, vB, wB, vB4.w);
	d.joints[j].impulse[0] = total;
}

__device__ __forceinline__ bool PulleySolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB;
	float4 pA = d.pos[bA], pB = d.pos[bB];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y);
	float aA = pA.z, aB = pB.z;
	const float ratio = jt.motorSpeed;

	Rot qA = SinCos(aA), qB = SinCos(aB);
	Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
	Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	Vec2 uA = cA + rA - V(jt.axis[0], jt.axis[1]);
	Vec2 uB = cB + rB - V(jt.lowerAngle, jt.upperAngle);
	float lengthA = Length(uA);
	float lengthB = Length(uB);
	uA = lengthA > 10.0f * B2CU_LINEAR_SLOP ? (1.0f / lengthA) * uA : V(0.0f, 0.0f);
	uB = lengthB > 10.0f * B2CU_LINEAR_SLOP ? (1.0f / lengthB) * uB : V(0.0f, 0.0f);

	float ruA = Cross(rA, uA);
	float ruB = Cross(rB, uB);
	float mA = r.invMassA + r.invIA * ruA * ruA;
	float mB = r.invMassB + r.invIB * ruB * ruB;
	float mass = mA + ratio * ratio * mB;
	if (mass > 0.0f) mass = 1.0f / mass;

	// m_constant = lengthA + ratio * lengthB of the definition
	float constant = jt.length + ratio * jt.referenceAngle;
	float C = constant - lengthA - ratio * lengthB;
	float linearError = Abs(C);
	float impulse = -mass * C;

	Vec2 PA = -impulse * uA;
	Vec2 PB = -ratio * impulse * uB;
	cA = cA + r.invMassA * PA;
	aA += r.invIA * Cross(rA, PA);
	cB = cB + r.invMassB * PB;
	aB += r.invIB * Cross(rB, PB);

	if (r.invMassA != 0.0f || r.invIA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (r.invMassB != 0.0f || r.invIB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	return linearError < B2CU_LINEAR_SLOP;
}

// ---- mouse joint (b2MouseJoint.cpp:96-190) --------------------------------------------------------------------------------
// Only body B is constrained (towards a world target).  Record: axis = target, length = maxForce, frequencyHz,
// dampingRatio, maxMotorTorque = b2Body::GetMass() of body B (the reference uses the mass, not its stored inverse).
// Row use: ex / ey = m_mass (inverse of K), u = m_C, gamma.

__device__ __forceinline__ void MouseInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting,
                                          float h)
{
	const int bB = jt.bodyB;
	float4 pB = d.pos[bB];
	float4 vB4 = d.vel[bB];
	Vec2 cB = V(pB.x, pB.y);
	Vec2 vB = V(vB4.x, vB4.y);
	float wB = vB4.z;
	Rot qB = SinCos(pB.z);

	float mass = jt.maxMotorTorque;
	float omega = 2.0f * B2CU_PI * jt.frequencyHz;
	float dd = 2.0f * mass * jt.dampingRatio * omega;
	float k = mass * (omega * omega);
	r.gamma = h * (dd + h * k);
	if (r.gamma != 0.0f) r.gamma = 1.0f / r.gamma;
	float beta = h * k * r.gamma;

	r.rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
	{
		float kxx = r.invMassB + r.invIB * r.rB.y * r.rB.y + r.gamma;
		float kxy = -r.invIB * r.rB.x * r.rB.y;
		float kyy = r.invMassB + r.invIB * r.rB.x * r.rB.x + r.gamma;
		float a = kxx, b = kxy, c = kxy, e = kyy;
		float det = a * e - b * c;
		if (det != 0.0f) det = 1.0f / det;
		r.ex = V3(det * e, -det * c, 0.0f);
		r.ey = V3(-det * b, det * a, 0.0f);
	}
	r.u = cB + r.rB - V(jt.axis[0], jt.axis[1]);
	r.u = V(r.u.x * beta, r.u.y * beta);

	wB *= 0.98f; // the reference's own damping fudge

	if (warmStarting)
	{
		jt.impulse[0] *= dtRatio;
		jt.impulse[1] *= dtRatio;
		Vec2 P = V(jt.impulse[0], jt.impulse[1]);
		vB = vB + r.invMassB * P;
		wB += r.invIB * Cross(r.rB, P);
	}
	else
	{
		jt.impulse[0] = jt.impulse[1] = 0.0f;
	}

	// body B's velocity is stored unconditionally (the 0.98 factor applies even to a body without mass)
	d.vel[bB] = make_float4(vB.x, vB.y, wB, vB4.w);
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void MouseSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, const b2cuJoint& jt, float h)
{
	const int bB = jt.bodyB;
	float4 vB4 = d.vel[bB];
	Vec2 vB = V(vB4.x, vB4.y);
	float wB = vB4.z;

	Vec2 Cdot = vB + CrossSV(wB, r.rB);
	Vec2 old = V(jt.impulse[0], jt.impulse[1]);
	Vec2 rhs = -(Cdot + r.u + r.gamma * old);
	Vec2 impulse = V(r.ex.x * rhs.x + r.ey.x * rhs.y, r.ex.y * rhs.x + r.ey.y * rhs.y);
	Vec2 total = old + impulse;
	float maxImpulse = h * jt.length;
	if (Dot(total, total) > maxImpulse * maxImpulse)
	{
		float scale = maxImpulse / Length(total);
		total = V(total.x * scale, total.y * scale);
	}
	impulse = total - old;

	vB = vB + r.invMassB * impulse;
	wB += r.invIB * Cross(r.rB, impulse);

	d.vel[bB] = make_float4(vB.x, vB.y, wB, vB4.w);
	d.joints[j].impulse[0] = total.x;
	d.joints[j].impulse[1] = total.y;
}

// ---- gear joint (b2GearJoint.cpp:131-390) ---------------------------------------------------------------------------------
// Couples the coordinates of two other joints (each revolute or prismatic): coordinateA + ratio * coordinateB = constant.
// Four bodies take part: A and B are the second bodies of joint 1 and 2, C and D their first bodies.
// Record: limitState = body C, reserved = body D, flags GEAR_PRISMATIC_1 / _2, axis = localAnchorC, (lowerAngle,
// upperAngle) = localAnchorD, work[0..1] = localAxisC, work[2..3] = localAxisD, referenceAngle = referenceAngleA,
// maxMotorTorque = referenceAngleB, motorSpeed = ratio, length = constant.
// Row use: ex = (lcC, mC), ey = (lcD, mD), ez = (iC, iD, -), axis = JvAC, perp = JvBD, a1 = JwA, a2 = JwB, s1 = JwC, s2 = JwD,
// motorMass = m_mass.

__device__ __forceinline__ void StoreVelocity4(const DeviceArrays& d, int body, float invMass, float invI, Vec2 v, float w, float keep)
{
	if (invMass != 0.0f || invI != 0.0f) d.vel[body] = make_float4(v.x, v.y, w, keep);
}

__device__ __forceinline__ void GearInit(const DeviceArrays& d, int j, b2cuJoint jt, JointRow r, float dtRatio, int warmStarting)
{
	const int bA = jt.bodyA, bB = jt.bodyB, bC = jt.limitState, bD = jt.reserved;
	float4 massC = d.mass[bC], massD = d.mass[bD];
	Vec2 lcC = V(massC.z, massC.w), lcD = V(massD.z, massD.w);
	float mA = r.invMassA, mB = r.invMassB, mC = massC.x, mD = massD.x;
	float iA = r.invIA, iB = r.invIB, iC = massC.y, iD = massD.y;
	r.ex = V3(lcC.x, lcC.y, mC);
	r.ey = V3(lcD.x, lcD.y, mD);
	r.ez = V3(iC, iD, 0.0f);
	const float ratio = jt.motorSpeed;

	float4 vA4 = d.vel[bA], vB4 = d.vel[bB], vC4 = d.vel[bC], vD4 = d.vel[bD];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y), vC = V(vC4.x, vC4.y), vD = V(vD4.x, vD4.y);
	float wA = vA4.z, wB = vB4.z, wC = vC4.z, wD = vD4.z;
	Rot qA = SinCos(d.pos[bA].z), qB = SinCos(d.pos[bB].z), qC = SinCos(d.pos[bC].z), qD = SinCos(d.pos[bD].z);

	float mass = 0.0f;
	if (!(jt.flags & B2CU_JOINT_GEAR_PRISMATIC_1))
	{
		r.axis = V(0.0f, 0.0f);
		r.a1 = 1.0f;
		r.s1 = 1.0f;
		mass += iA + iC;
	}
	else
	{
		Vec2 u = Mul(qC, V(jt.work[0], jt.work[1]));
		Vec2 rC = Mul(qC, V(jt.axis[0], jt.axis[1]) - lcC);
		Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
		r.axis = u;
		r.s1 = Cross(rC, u);
		r.a1 = Cross(rA, u);
		mass += mC + mA + iC * r.s1 * r.s1 + iA * r.a1 * r.a1;
	}
	if (!(jt.flags & B2CU_JOINT_GEAR_PRISMATIC_2))
	{
		r.perp = V(0.0f, 0.0f);
		r.a2 = ratio;
		r.s2 = ratio;
		mass += ratio * ratio * (iB + iD);
	}
	else
	{
		Vec2 u = Mul(qD, V(jt.work[2], jt.work[3]));
		Vec2 rD = Mul(qD, V(jt.lowerAngle, jt.upperAngle) - lcD);
		Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
		r.perp = ratio * u;
		r.s2 = ratio * Cross(rD, u);
		r.a2 = ratio * Cross(rB, u);
		mass += ratio * ratio * (mD + mB) + iD * r.s2 * r.s2 + iB * r.a2 * r.a2;
	}
	r.motorMass = mass > 0.0f ? 1.0f / mass : 0.0f;

	if (warmStarting)
	{
		// note: the reference does not scale the gear impulse by dtRatio
		float imp = jt.impulse[0];
		vA = vA + (mA * imp) * r.axis;
		wA += iA * imp * r.a1;
		vB = vB + (mB * imp) * r.perp;
		wB += iB * imp * r.a2;
		vC = vC - (mC * imp) * r.axis;
		wC -= iC * imp * r.s1;
		vD = vD - (mD * imp) * r.perp;
		wD -= iD * imp * r.s2;
	}
	else
	{
		jt.impulse[0] = 0.0f;
	}

	StoreVelocity4(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity4(d, bB, mB, iB, vB, wB, vB4.w);
	StoreVelocity4(d, bC, mC, iC, vC, wC, vC4.w);
	StoreVelocity4(d, bD, mD, iD, vD, wD, vD4.w);
	jt.lastSolve[0] = r.axis.x; // GetReactionForce / GetReactionTorque read m_JvAC and m_JwA
	jt.lastSolve[1] = r.axis.y;
	jt.lastSolve[2] = r.a1;
	d.jointRows[j] = r;
	d.joints[j] = jt;
}

__device__ __forceinline__ void GearSolveVelocity(const DeviceArrays& d, int j, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB, bC = jt.limitState, bD = jt.reserved;
	float mA = r.invMassA, mB = r.invMassB, mC = r.ex.z, mD = r.ey.z;
	float iA = r.invIA, iB = r.invIB, iC = r.ez.x, iD = r.ez.y;
	float4 vA4 = d.vel[bA], vB4 = d.vel[bB], vC4 = d.vel[bC], vD4 = d.vel[bD];
	Vec2 vA = V(vA4.x, vA4.y), vB = V(vB4.x, vB4.y), vC = V(vC4.x, vC4.y), vD = V(vD4.x, vD4.y);
	float wA = vA4.z, wB = vB4.z, wC = vC4.z, wD = vD4.z;

	float Cdot = Dot(r.axis, vA - vC) + Dot(r.perp, vB - vD);
	Cdot += (r.a1 * wA - r.s1 * wC) + (r.a2 * wB - r.s2 * wD);
	float impulse = -r.motorMass * Cdot;
	float total = jt.impulse[0] + impulse;

	vA = vA + (mA * impulse) * r.axis;
	wA += iA * impulse * r.a1;
	vB = vB + (mB * impulse) * r.perp;
	wB += iB * impulse * r.a2;
	vC = vC - (mC * impulse) * r.axis;
	wC -= iC * impulse * r.s1;
	vD = vD - (mD * impulse) * r.perp;
	wD -= iD * impulse * r.s2;

	StoreVelocity4(d, bA, mA, iA, vA, wA, vA4.w);
	StoreVelocity4(d, bB, mB, iB, vB, wB, vB4.w);
	StoreVelocity4(d, bC, mC, iC, vC, wC, vC4.w);
	StoreVelocity4(d, bD, mD, iD, vD, wD, vD4.w);
	d.joints[j].impulse[0] = total;
}

__device__ __forceinline__ bool GearSolvePosition(const DeviceArrays& d, const JointRow& r, const b2cuJoint& jt)
{
	const int bA = jt.bodyA, bB = jt.bodyB, bC = jt.limitState, bD = jt.reserved;
	float mA = r.invMassA, mB = r.invMassB, mC = r.ex.z, mD = r.ey.z;
	float iA = r.invIA, iB = r.invIB, iC = r.ez.x, iD = r.ez.y;
	Vec2 lcC = V(r.ex.x, r.ex.y), lcD = V(r.ey.x, r.ey.y);
	const float ratio = jt.motorSpeed;
	float4 pA = d.pos[bA], pB = d.pos[bB], pC = d.pos[bC], pD = d.pos[bD];
	Vec2 cA = V(pA.x, pA.y), cB = V(pB.x, pB.y), cC = V(pC.x, pC.y), cD = V(pD.x, pD.y);
	float aA = pA.z, aB = pB.z, aC = pC.z, aD = pD.z;
	Rot qA = SinCos(aA), qB = SinCos(aB), qC = SinCos(aC), qD = SinCos(aD);

	float coordinateA, coordinateB;
	Vec2 JvAC, JvBD;
	float JwA, JwB, JwC, JwD;
	float mass = 0.0f;

	if (!(jt.flags & B2CU_JOINT_GEAR_PRISMATIC_1))
	{
		JvAC = V(0.0f, 0.0f);
		JwA = 1.0f;
		JwC = 1.0f;
		mass += iA + iC;
		coordinateA = aA - aC - jt.referenceAngle;
	}
	else
	{
		Vec2 localAxisC = V(jt.work[0], jt.work[1]);
		Vec2 u = Mul(qC, localAxisC);
		Vec2 rC = Mul(qC, V(jt.axis[0], jt.axis[1]) - lcC);
		Vec2 rA = Mul(qA, V(jt.localAnchorA[0], jt.localAnchorA[1]) - r.localCenterA);
		JvAC = u;
		JwC = Cross(rC, u);
		JwA = Cross(rA, u);
		mass += mC + mA + iC * JwC * JwC + iA * JwA * JwA;
		Vec2 pc = V(jt.axis[0], jt.axis[1]) - lcC;
		Vec2 pa = MulT(qC, rA + (cA - cC));
		coordinateA = Dot(pa - pc, localAxisC);
	}
	if (!(jt.flags & B2CU_JOINT_GEAR_PRISMATIC_2))
	{
		JvBD = V(0.0f, 0.0f);
		JwB = ratio;
		JwD = ratio;
		mass += ratio * ratio * (iB + iD);
		coordinateB = aB - aD - jt.maxMotorTorque;
	}
	else
	{
		Vec2 localAxisD = V(jt.work[2], jt.work[3]);
		Vec2 u = Mul(qD, localAxisD);
		Vec2 rD = Mul(qD, V(jt.lowerAngle, jt.upperAngle) - lcD);
		Vec2 rB = Mul(qB, V(jt.localAnchorB[0], jt.localAnchorB[1]) - r.localCenterB);
		JvBD = ratio * u;
		JwD = ratio * Cross(rD, u);
		JwB = ratio * Cross(rB, u);
		mass += ratio * ratio * (mD + mB) + iD * JwD * JwD + iB * JwB * JwB;
		Vec2 pd = V(jt.lowerAngle, jt.upperAngle) - lcD;
		Vec2 pb = MulT(qD, rB + (cB - cD));
		coordinateB = Dot(pb - pd, localAxisD);
	}

	float C = (coordinateA + ratio * coordinateB) - jt.length;
	float impulse = 0.0f;
	if (mass > 0.0f) impulse = -C / mass;

	cA = cA + mA * impulse * JvAC;
	aA += iA * impulse * JwA;
	cB = cB + mB * impulse * JvBD;
	aB += iB * impulse * JwB;
	cC = cC - mC * impulse * JvAC;
	aC -= iC * impulse * JwC;
	cD = cD - mD * impulse * JvBD;
	aD -= iD * impulse * JwD;

	if (mA != 0.0f || iA != 0.0f) d.pos[bA] = make_float4(cA.x, cA.y, aA, pA.w);
	if (mB != 0.0f || iB != 0.0f) d.pos[bB] = make_float4(cB.x, cB.y, aB, pB.w);
	if (mC != 0.0f || iC != 0.0f) d.pos[bC] = make_float4(cC.x, cC.y, aC, pC.w);
	if (mD != 0.0f || iD != 0.0f) d.pos[bD] = make_float4(cD.x, cD.y, aD, pD.w);
	// the reference never measures the gear's error: it always reports "solved"
	return true;
}

// ---- dispatch by joint type -------------------------------------------------------------------------------------------

__device__ __forceinline__ void JointInitOne(const DeviceArrays& d, int j, float dtRatio, int warmStarting, float h)
{
	b2cuJoint jt = d.joints[j];
	JointRow r;
	if (!JointPrepare(d, jt, r))
	{
		d.jointRows[j].solved = 0;
		return;
	}
	if (jt.type == B2CU_JOINT_REVOLUTE) RevoluteInit(d, j, jt, r, dtRatio, warmStarting);
	else if (jt.type == B2CU_JOINT_PRISMATIC) PrismaticInit(d, j, jt, r, dtRatio, warmStarting);
	else if (jt.type == B2CU_JOINT_DISTANCE) DistanceInit(d, j, jt, r, dtRatio, warmStarting, h);
	else if (jt.type == B2CU_JOINT_WHEEL) WheelInit(d, j, jt, r, dtRatio, warmStarting, h);
	else if (jt.type == B2CU_JOINT_ROPE) RopeInit(d, j, jt, r, dtRatio, warmStarting);
	else if (jt.type == B2CU_JOINT_FRICTION) FrictionMotorInit(d, j, jt, r, dtRatio, warmStarting, false);
	else if (jt.type == B2CU_JOINT_MOTOR) FrictionMotorInit(d, j, jt, r, dtRatio, warmStarting, true);
	else if (jt.type == B2CU_JOINT_PULLEY) PulleyInit(d, j, jt, r, dtRatio, warmStarting);
	else if (jt.type == B2CU_JOINT_MOUSE) MouseInit(d, j, jt, r, dtRatio, warmStarting, h);
	else if (jt.type == B2CU_JOINT_GEAR) GearInit(d, j, jt, r, dtRatio, warmStarting);
	else WeldInit(d, j, jt, r, dtRatio, warmStarting, h);
}

__device__ __forceinline__ void JointSolveVelocityOne(const DeviceArrays& d, int j, float h)
{
	JointRow r = d.jointRows[j];
	if (!r.solved) return;
	b2cuJoint jt = d.joints[j];
	if (jt.type == B2CU_JOINT_REVOLUTE) RevoluteSolveVelocity(d, j, r, jt, h);
	else if (jt.type == B2CU_JOINT_PRISMATIC) PrismaticSolveVelocity(d, j, r, jt, h);
	else if (jt.type == B2CU_JOINT_DISTANCE) DistanceSolveVelocity(d, j, r, jt);
	else if (jt.type == B2CU_JOINT_WHEEL) WheelSolveVelocity(d, j, r, jt, h);
	else if (jt.type == B2CU_JOINT_ROPE) RopeSolveVelocity(d, j, r, jt, h);
	else if (jt.type == B2CU_JOINT_FRICTION) FrictionMotorSolveVelocity(d, j, r, jt, h, false);
	else if (jt.type == B2CU_JOINT_MOTOR) FrictionMotorSolveVelocity(d, j, r, jt, h, true);
	else if (jt.type == B2CU_JOINT_PULLEY) PulleySolveVelocity(d, j, r, jt);
	else if (jt.type == B2CU_JOINT_MOUSE) MouseSolveVelocity(d, j, r, jt, h);
	else if (jt.type == B2CU_JOINT_GEAR) GearSolveVelocity(d, j, r, jt);
	else WeldSolveVelocity(d, j, r, jt);
}

__device__ __forceinline__ bool JointSolvePositionOne(const DeviceArrays& d, int j, const JointRow& r)
{
	b2cuJoint jt = d.joints[j];
	if (jt.type == B2CU_JOINT_REVOLUTE) return RevoluteSolvePosition(d, r, jt);
	if (jt.type == B2CU_JOINT_PRISMATIC) return PrismaticSolvePosition(d, r, jt);
	if (jt.type == B2CU_JOINT_DISTANCE) return DistanceSolvePosition(d, r, jt);
	if (jt.type == B2CU_JOINT_WHEEL) return WheelSolvePosition(d, r, jt);
	if (jt.type == B2CU_JOINT_ROPE) return RopeSolvePosition(d, r, jt);
	if (jt.type == B2CU_JOINT_FRICTION || jt.type == B2CU_JOINT_MOTOR || jt.type == B2CU_JOINT_MOUSE) return true;
	if (jt.type == B2CU_JOINT_PULLEY) return PulleySolvePosition(d, r, jt);
	if (jt.type == B2CU_JOINT_GEAR) return GearSolvePosition(d, r, jt);
	return WeldSolvePosition(d, r, jt);
}

} // namespace b2cu
