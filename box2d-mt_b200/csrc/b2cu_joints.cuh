// b2cu_joints.cuh -- joints as rows of the coloured solver.  Restates b2RevoluteJoint::InitVelocityConstraints /
// SolveVelocityConstraints / SolvePositionConstraints (Box2D/Dynamics/Joints/b2RevoluteJoint.cpp:64-400) and the small
// linear solves they use (b2Mat33::Solve33 / Solve22, Box2D/Common/b2Math.cpp:25-53; b2Mat22::Solve, b2Math.h:221-233)
// with the same fp32 arithmetic in the same order.  One thread solves one joint; joints of one colour class share no
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
	Vec3 ex, ey, ez; // m_mass
	float motorMass;
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

// b2RevoluteJoint::InitVelocityConstraints (:64-183)
__device__ __forceinline__ void JointInitOne(const DeviceArrays& d, int j, float dtRatio, int warmStarting)
{
	b2cuJoint jt = d.joints[j];
	JointRow r;
	const int bA = jt.bodyA, bB = jt.bodyB;
	const uint32_t fA = d.bflags[bA], fB = d.bflags[bB];
	// the island search adds a joint from an island body when the body across is active (b2World.cpp:1286-1320)
	const bool staticA = (fA & B2CU_BODY_TYPE_MASK) == B2CU_STATIC_BODY, staticB = (fB & B2CU_BODY_TYPE_MASK) == B2CU_STATIC_BODY;
	const bool islandA = !staticA && (fA & B2CU_BODY_ISLAND), islandB = !staticB && (fB & B2CU_BODY_ISLAND);
	r.solved = (islandA || islandB) && (fA & B2CU_BODY_ACTIVE) && (fB & B2CU_BODY_ACTIVE) ? 1 : 0;
	r.root = islandA ? d.island[bA] : (islandB ? d.island[bB] : 0);
	if (!r.solved)
	{
		d.jointRows[j].solved = 0;
		return;
	}
	float4 massA = d.mass[bA], massB = d.mass[bB];
	r.localCenterA = V(massA.z, massA.w);
	r.localCenterB = V(massB.z, massB.w);
	r.invMassA = massA.x;
	r.invMassB = massB.x;
	r.invIA = massA.y;
	r.invIB = massB.y;

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

	r.ex.x = mA + mB + r.rA.y * r.rA.y * iA + r.rB.y * r.rB.y * iB;
	r.ey.x = -r.rA.y * r.rA.x * iA - r.rB.y * r.rB.x * iB;
	r.ez.x = -r.rA.y * iA - r.rB.y * iB;
	r.ex.y = r.ey.x;
	r.ey.y = mA + mB + r.rA.x * r.rA.x * iA + r.rB.x * r.rB.x * iB;
	r.ez.y = r.rA.x * iA + r.rB.x * iB;
	r.ex.z = r.ez.x;
	r.ey.z = r.ez.y;
	r.ez.z = iA + iB;

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
__device__ __forceinline__ void JointSolveVelocityOne(const DeviceArrays& d, int j, float h)
{
	JointRow r = d.jointRows[j];
	if (!r.solved) return;
	b2cuJoint jt = d.joints[j];
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
__device__ __forceinline__ bool JointSolvePositionOne(const DeviceArrays& d, int j, const JointRow& r)
{
	b2cuJoint jt = d.joints[j];
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

} // namespace b2cu
