// Minimal stand-in for the slice of GLM that the SPH engine API surface needs.
//
// The reference engine exposes glm::vec3 in its public signatures
// (reference src/sph_engine.h:133-134, src/particle.h:19-21) but does not vendor
// GLM.  This header provides just enough of the namespace for (a) our host-side
// SPHEngine shell and (b) compiling the reference translation units as the parity
// oracle.  If a real GLM is on the include path ahead of include/compat it is
// picked up instead; the arithmetic below is component-wise IEEE fp32 evaluated
// left to right, which is what GLM's scalar (non-SIMD) code path does.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

namespace glm {

struct vec2 {
    float x, y;
    vec2() : x(0.0f), y(0.0f) {}
    explicit vec2(float s) : x(s), y(s) {}
    template <typename A, typename B>
    vec2(A a, B b) : x(static_cast<float>(a)), y(static_cast<float>(b)) {}
};

struct vec3 {
    union { float x; float r; };
    union { float y; float g; };
    union { float z; float b; };

    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    vec3(A a, B b_, C c) : x(static_cast<float>(a)), y(static_cast<float>(b_)), z(static_cast<float>(c)) {}

    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }

    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
};

inline ivec3 operator+(const ivec3& a, const ivec3& b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline float length(const vec2& a) { return std::sqrt(dot(a, a)); }

}  // namespace glm
