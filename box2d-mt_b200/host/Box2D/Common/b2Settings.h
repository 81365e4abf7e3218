// Host-side mirror of the reference's settings header (reference: Box2D/Common/b2Settings.h:25-190).
// Types, tuning constants and version info that user code sees through "Box2D/Box2D.h".  The numerical
// constants must equal the reference's; the device restates them in b2cu_math.cuh.
#ifndef B2_SETTINGS_H
#define B2_SETTINGS_H

#include <cassert>
#include <cfloat>
#include <cstddef>
#include <cstdint>

#define B2_NOT_USED(x) ((void)(x))
#define b2Assert(A) assert(A)

typedef std::int8_t int8;
typedef std::int16_t int16;
typedef std::int32_t int32;
typedef std::uint8_t uint8;
typedef std::uint16_t uint16;
typedef std::uint32_t uint32;
typedef std::uint64_t uint64;
typedef float float32;
typedef double float64;

// Tuning constants: typed constants instead of the reference's macros (same names, same values, usable in the same
// expressions and as array bounds).
constexpr float32 b2_maxFloat = FLT_MAX, b2_epsilon = FLT_EPSILON, b2_pi = 3.14159265359f;

constexpr int32 b2_maxManifoldPoints = 2, b2_maxPolygonVertices = 8, b2_maxSubSteps = 8, b2_maxTOIContacts = 32;
constexpr float32 b2_aabbExtension = 0.1f, b2_aabbMultiplier = 2.0f;                      // fat AABB rule
constexpr float32 b2_linearSlop = 0.005f, b2_angularSlop = 2.0f / 180.0f * b2_pi;
constexpr float32 b2_polygonRadius = 2.0f * b2_linearSlop;

constexpr float32 b2_velocityThreshold = 1.0f;                                            // restitution cut-off
constexpr float32 b2_maxLinearCorrection = 0.2f, b2_maxAngularCorrection = 8.0f / 180.0f * b2_pi;
constexpr float32 b2_maxTranslation = 2.0f, b2_maxTranslationSquared = b2_maxTranslation * b2_maxTranslation;
constexpr float32 b2_maxRotation = 0.5f * b2_pi, b2_maxRotationSquared = b2_maxRotation * b2_maxRotation;
constexpr float32 b2_baumgarte = 0.2f, b2_toiBaugarte = 0.75f;

constexpr float32 b2_timeToSleep = 0.5f;                                                  // island sleep rule
constexpr float32 b2_linearSleepTolerance = 0.01f, b2_angularSleepTolerance = 2.0f / 180.0f * b2_pi;

// limits of the task layer, kept so that user task code compiles unchanged (reference :162-174)
constexpr int32 b2_cacheLineSize = 64, b2_maxThreads = 8, b2_maxRangeSubTasks = b2_maxThreads;
constexpr int32 b2_maxWorldStepTaskGroups = 1;

struct b2Version
{
	int32 major, minor, revision;
};

extern b2Version b2_version;      // Box2D version the API mirrors (2.3.2)
extern b2Version b2_mtVersion;    // Box2D-MT layer version (0.1.0)

void b2Log(const char* string, ...);

/// user-overridable allocation of the reference (b2Settings.h:176-183); the code b2World::Dump writes uses them
void* b2Alloc(int32 size);
void b2Free(void* mem);

#endif
