// TEST INFRASTRUCTURE — CPU oracle for the pedoni per-timestep pedestrian update.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// use anything under oracle/. The product (pedoni_b200/) never includes or links this.
//
// vec2.hpp: the subset of glam 0.29.2 `Vec2` / `IVec2` arithmetic the reference's hot path calls
// (glam is a crates.io dependency, Cargo.lock:525-526, NOT vendored under /root/reference, so its
// published formulations are restated here; call sites: sfm.rs:108,113,131-149,190,197-199,223,
// 251-253; neighbor_grid.rs:15,27; util.rs:47-49,93-101,107-108; field.rs:25,236,243,250,256).
//
// Build with -ffp-contract=off -fno-fast-math: rustc never contracts a*b+c into an FMA.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace oracle {

struct Vec2 {
    float x, y;
};

inline Vec2 vec2(float x, float y) { return Vec2{x, y}; }
inline Vec2 operator+(Vec2 a, Vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2 operator-(Vec2 a, Vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline Vec2 operator-(Vec2 a) { return {-a.x, -a.y}; }
inline Vec2 operator*(Vec2 a, float s) { return {a.x * s, a.y * s}; }
inline Vec2 operator*(float s, Vec2 a) { return {s * a.x, s * a.y}; }
inline Vec2 operator/(Vec2 a, float s) { return {a.x / s, a.y / s}; }  // true per-component divide
inline Vec2& operator+=(Vec2& a, Vec2 b) {
    a.x += b.x;
    a.y += b.y;
    return a;
}

inline float dot(Vec2 a, Vec2 b) { return (a.x * b.x) + (a.y * b.y); }
inline float length_squared(Vec2 a) { return dot(a, a); }
inline float length(Vec2 a) { return std::sqrt(dot(a, a)); }
inline float length_recip(Vec2 a) { return 1.0f / length(a); }
// glam: `self.mul(self.length_recip())`, no zero guard in release builds -> NaN for a zero vector.
inline Vec2 normalize(Vec2 a) { return a * length_recip(a); }
inline Vec2 normalize_or_zero(Vec2 a) {
    float rcp = length_recip(a);
    if (std::isfinite(rcp) && rcp > 0.0f) return a * rcp;
    return {0.0f, 0.0f};
}
// glam: compares squared lengths, then `max * (self / sqrt(length_sq))`.
inline Vec2 clamp_length_max(Vec2 a, float max) {
    float length_sq = length_squared(a);
    if (length_sq > max * max) return max * (a / std::sqrt(length_sq));
    return a;
}
inline Vec2 floor(Vec2 a) { return {std::floor(a.x), std::floor(a.y)}; }
inline Vec2 ceil(Vec2 a) { return {std::ceil(a.x), std::ceil(a.y)}; }

// Rust `f32 as i32`: truncate toward zero, saturate, NaN -> 0.
inline int32_t f32_as_i32(float v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.0f) return std::numeric_limits<int32_t>::max();
    if (v <= -2147483648.0f) return std::numeric_limits<int32_t>::min();
    return static_cast<int32_t>(v);
}
// Rust `f32 as usize` (64-bit): saturating at 0 below, NaN -> 0.
inline uint64_t f32_as_usize(float v) {
    if (std::isnan(v) || v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return std::numeric_limits<uint64_t>::max();
    return static_cast<uint64_t>(v);
}

struct IVec2 {
    int32_t x, y;
};
inline IVec2 as_ivec2(Vec2 a) { return {f32_as_i32(a.x), f32_as_i32(a.y)}; }

}  // namespace oracle
