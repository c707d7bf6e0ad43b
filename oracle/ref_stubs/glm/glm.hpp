// TEST INFRASTRUCTURE — a minimal stand-in for the parts of glm that apps/lbmMultiRes/util.h touches.
//
// The reference's D3Q27 code (apps/lbmMultiRes/{lattice,collide,stream,util}.h) is compiled UNMODIFIED by
// oracle/Makefile.ref27; util.h includes <glm/glm.hpp> for one signed-distance helper (sdfJetfighter) that the LBM
// kernels never call.  glm is not available offline, so this header provides just enough of its vocabulary (vec2, vec3,
// ivec3, mat2, mat3, a dozen free functions, function swizzles) for that helper to compile.  Nothing here is on any
// numerical path of the oracle or of the product.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define GLMS_HD __host__ __device__
#else
#define GLMS_HD
#endif

namespace glm {

struct vec2
{
    float x, y;
    GLMS_HD vec2() : x(0), y(0) {}
    GLMS_HD explicit vec2(float s) : x(s), y(s) {}
    GLMS_HD vec2(float a, float b) : x(a), y(b) {}
    GLMS_HD float&       operator[](int i) { return i == 0 ? x : y; }
    GLMS_HD const float& operator[](int i) const { return i == 0 ? x : y; }
    GLMS_HD vec2&        operator-=(const vec2& o) { x -= o.x; y -= o.y; return *this; }
};
GLMS_HD inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
GLMS_HD inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
GLMS_HD inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }

struct mat2
{
    float m[4];
    GLMS_HD mat2(float a, float b, float c, float d) : m{a, b, c, d} {}
};
GLMS_HD inline vec2 operator*(vec2 v, const mat2& M) { return vec2(v.x * M.m[0] + v.y * M.m[1], v.x * M.m[2] + v.y * M.m[3]); }

struct vec3
{
    float x, y, z;
    GLMS_HD vec3() : x(0), y(0), z(0) {}
    GLMS_HD explicit vec3(float s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    GLMS_HD vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
    GLMS_HD float&       operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    GLMS_HD const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    GLMS_HD vec3&        operator*=(double s) { x = (float)(x * s); y = (float)(y * s); z = (float)(z * s); return *this; }
    // function swizzles (GLM_FORCE_SWIZZLE without language extensions: by value)
    GLMS_HD vec2 xy() const { return vec2(x, y); }
    GLMS_HD vec2 xz() const { return vec2(x, z); }
    GLMS_HD vec3 xyz() const { return *this; }
};
GLMS_HD inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS_HD inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS_HD inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
GLMS_HD inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }

struct ivec3
{
    int x, y, z;
    GLMS_HD ivec3() : x(0), y(0), z(0) {}
    GLMS_HD ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
};

struct mat3
{
    float m[9];
    GLMS_HD mat3(float a, float b, float c, float d, float e, float f, float g, float h, float i) : m{a, b, c, d, e, f, g, h, i} {}
};
GLMS_HD inline vec3 operator*(vec3 v, const mat3& M)
{
    return vec3(v.x * M.m[0] + v.y * M.m[1] + v.z * M.m[2], v.x * M.m[3] + v.y * M.m[4] + v.z * M.m[5],
                v.x * M.m[6] + v.y * M.m[7] + v.z * M.m[8]);
}

template <typename T>
GLMS_HD constexpr T pi() { return (T)3.14159265358979323846; }
GLMS_HD inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
GLMS_HD inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GLMS_HD inline float length(vec2 a) { return sqrtf(dot(a, a)); }
GLMS_HD inline float length(vec3 a) { return sqrtf(dot(a, a)); }
GLMS_HD inline vec3  normalize(vec3 a) { return a * (1.0f / length(a)); }
GLMS_HD inline vec3  abs(vec3 a) { return vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
GLMS_HD inline vec3  max(vec3 a, float s) { return vec3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }
GLMS_HD inline vec2  max(vec2 a, float s) { return vec2(fmaxf(a.x, s), fmaxf(a.y, s)); }
GLMS_HD inline vec2  max(vec2 a, vec2 b) { return vec2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
GLMS_HD inline float clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
GLMS_HD inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
GLMS_HD inline float step(float edge, float v) { return v < edge ? 0.0f : 1.0f; }
GLMS_HD inline float mod(float a, float b) { return a - b * floorf(a / b); }
GLMS_HD inline float atan(float y, float x) { return atan2f(y, x); }
GLMS_HD inline float cos(float a) { return cosf(a); }
GLMS_HD inline float sin(float a) { return sinf(a); }

}  // namespace glm
