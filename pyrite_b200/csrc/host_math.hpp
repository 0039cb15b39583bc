// Host-side f32 linear algebra for the scene builder.  The operation order of every function is
// fixed (no FMA contraction: the host code is compiled with -ffp-contract=off) so that baked
// geometry is reproducible; the formulas follow cgmath 0.17 / collision 0.20 as used by the
// reference at the cited call sites.
#pragma once
#include <cmath>
#include <cstdint>

namespace pyr {
namespace host {

struct V2 { float x = 0, y = 0; };
struct V3 {
    float x = 0, y = 0, z = 0;
    V3() = default;
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float at(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct V4 { float x = 0, y = 0, z = 0, w = 0; };

inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 scale(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(V3 a) { return sqrtf(dot3(a, a)); }
inline V3 unit(V3 a) { return scale(a, 1.0f / length(a)); }  // cgmath normalize = v * (1/|v|)

struct Q4 { float s = 1, x = 0, y = 0, z = 0; };  // scalar-first quaternion
inline V3 qrotate(Q4 q, V3 v) {  // cgmath Quaternion * Vector3 (shapes/mod.rs:565)
    V3 qv{q.x, q.y, q.z};
    V3 tmp = add(cross3(qv, v), scale(v, q.s));
    return add(scale(cross3(qv, tmp), 2.0f), v);
}

// 3x3 given by its columns; Matrix3 -> Quaternion (world.rs:100,357-367, shapes/mod.rs:546,580)
inline Q4 quat_from_columns(V3 c0, V3 c1, V3 c2) {
    const float m00 = c0.x, m01 = c0.y, m02 = c0.z, m10 = c1.x, m11 = c1.y, m12 = c1.z, m20 = c2.x, m21 = c2.y, m22 = c2.z;
    float trace = m00 + m11 + m22;
    if (trace >= 0.0f) {
        float s = sqrtf(1.0f + trace);
        float w = 0.5f * s;
        s = 0.5f / s;
        return {w, (m12 - m21) * s, (m20 - m02) * s, (m01 - m10) * s};
    }
    if (m00 > m11 && m00 > m22) {
        float s = sqrtf((m00 - m11 - m22) + 1.0f);
        float x = 0.5f * s;
        s = 0.5f / s;
        return {(m12 - m21) * s, x, (m10 + m01) * s, (m02 + m20) * s};
    }
    if (m11 > m22) {
        float s = sqrtf((m11 - m00 - m22) + 1.0f);
        float y = 0.5f * s;
        s = 0.5f / s;
        return {(m20 - m02) * s, (m10 + m01) * s, y, (m21 + m12) * s};
    }
    float s = sqrtf((m22 - m00 - m11) + 1.0f);
    float z = 0.5f * s;
    s = 0.5f / s;
    return {(m01 - m10) * s, (m02 + m20) * s, (m21 + m12) * s, z};
}

// math.rs:98-123
inline V3 ortho(V3 v) {
    const float E = 0.0001f;
    V3 u;
    if (fabsf(v.x) < E) u = {1, 0, 0};
    else if (fabsf(v.y) < E) u = {0, 1, 0};
    else if (fabsf(v.z) < E) u = {0, 0, 1};
    else u = {-v.y, v.x, 0.0f};
    return cross3(v, u);
}
inline void basis(V3 x, V3& y, V3& z) {
    z = unit(ortho(x));
    y = unit(cross3(z, x));
}

// 4x4, column-major: e[col*4 + row]
struct M4 {
    float e[16];
    static M4 identity() {
        M4 m{};
        for (int i = 0; i < 16; ++i) m.e[i] = (i % 5 == 0) ? 1.0f : 0.0f;
        return m;
    }
    float at(int col, int row) const { return e[col * 4 + row]; }
};
inline V4 mul(const M4& m, V4 v) {  // c0*x + c1*y + c2*z + c3*w, left to right
    V4 r;
    r.x = ((m.e[0] * v.x + m.e[4] * v.y) + m.e[8] * v.z) + m.e[12] * v.w;
    r.y = ((m.e[1] * v.x + m.e[5] * v.y) + m.e[9] * v.z) + m.e[13] * v.w;
    r.z = ((m.e[2] * v.x + m.e[6] * v.y) + m.e[10] * v.z) + m.e[14] * v.w;
    r.w = ((m.e[3] * v.x + m.e[7] * v.y) + m.e[11] * v.z) + m.e[15] * v.w;
    return r;
}
inline V3 xform_point(const M4& m, V3 p) {  // Transform::transform_point: divide by w via reciprocal
    V4 h = mul(m, V4{p.x, p.y, p.z, 1.0f});
    float inv = 1.0f / h.w;
    return {h.x * inv, h.y * inv, h.z * inv};
}
inline V3 xform_vector(const M4& m, V3 v) {
    V4 h = mul(m, V4{v.x, v.y, v.z, 0.0f});
    return {h.x, h.y, h.z};
}
// Matrix4::look_at (right-handed), project/mod.rs:263
inline M4 look_at_rh(V3 eye, V3 center, V3 up) {
    V3 f = unit(sub(center, eye));
    V3 s = unit(cross3(f, up));
    V3 u = cross3(s, f);
    M4 m{};
    m.e[0] = s.x; m.e[1] = u.x; m.e[2] = -f.x; m.e[3] = 0;
    m.e[4] = s.y; m.e[5] = u.y; m.e[6] = -f.y; m.e[7] = 0;
    m.e[8] = s.z; m.e[9] = u.z; m.e[10] = -f.z; m.e[11] = 0;
    m.e[12] = -dot3(eye, s); m.e[13] = -dot3(eye, u); m.e[14] = dot3(eye, f); m.e[15] = 1;
    return m;
}
// adjugate / determinant inverse, determinant expanded along the first row
inline bool invert4(const M4& a, M4& out) {
    auto minor3 = [&](int skip_col, int skip_row) {
        float t[3][3];
        int cc = 0;
        for (int col = 0; col < 4; ++col) {
            if (col == skip_col) continue;
            int rr = 0;
            for (int row = 0; row < 4; ++row) {
                if (row == skip_row) continue;
                t[cc][rr++] = a.at(col, row);
            }
            ++cc;
        }
        return t[0][0] * (t[1][1] * t[2][2] - t[2][1] * t[1][2]) - t[1][0] * (t[0][1] * t[2][2] - t[2][1] * t[0][2]) +
               t[2][0] * (t[0][1] * t[1][2] - t[1][1] * t[0][2]);
    };
    float det = 0.0f;
    for (int col = 0; col < 4; ++col) det += a.at(col, 0) * (minor3(col, 0) * ((col & 1) ? -1.0f : 1.0f));
    if (det == 0.0f || det != det) return false;
    float inv_det = 1.0f / det;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) out.e[col * 4 + row] = (minor3(row, col) * (((row + col) & 1) ? -1.0f : 1.0f)) * inv_det;
    return true;
}

struct Box {
    V3 lo, hi;
    static Box of(V3 a, V3 b) {
        return {{a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z},
                {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}};
    }
    Box with(V3 p) const { return of_minmax(lo, hi, p); }
    static Box of_minmax(V3 lo, V3 hi, V3 p) {
        return {{lo.x < p.x ? lo.x : p.x, lo.y < p.y ? lo.y : p.y, lo.z < p.z ? lo.z : p.z},
                {hi.x > p.x ? hi.x : p.x, hi.y > p.y ? hi.y : p.y, hi.z > p.z ? hi.z : p.z}};
    }
    Box merged(const Box& o) const { return with(o.lo).with(o.hi); }  // Aabb3::union = grow(min).grow(max)
    V3 extent() const { return sub(hi, lo); }
    V3 middle() const { V3 d = extent(); return {lo.x + d.x / 2.0f, lo.y + d.y / 2.0f, lo.z + d.z / 2.0f}; }
    float area() const { V3 d = extent(); return 2.0f * ((d.x * d.y) + (d.x * d.z) + (d.y * d.z)); }
};

inline uint64_t sat_usize(float v) {  // Rust `as usize`
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return ~0ull;
    return (uint64_t)v;
}

}  // namespace host
}  // namespace pyr
