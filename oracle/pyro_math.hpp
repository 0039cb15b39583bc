// ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, the smoke check
// and bench.py's cpu_baseline / --impl reference legs may build or load anything in oracle/.
//
// CPU restatement of the arithmetic pyrite's render path takes from un-vendored crates
// (cgmath 0.17.0, collision 0.20.1, rand 0.8.5, rand_xorshift 0.3.0; versions from the
// reference's Cargo.lock).  None of those sources are under /root/reference and the
// reference ships no tests or golden vectors, so every definition here is written from
// the crates' published algorithms: PARITY UNPINNED.
// Call sites in the reference are cited next to each function.
//
// Build with -O2 -ffp-contract=off -fno-fast-math: the reference never fuses mul+add.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace pyro {

constexpr float DIST_EPSILON = 0.0001f;  // math.rs:4
constexpr float PI = 3.14159265358979323846f;
constexpr float FRAC_1_PI = 0.318309886183790671537767526745028724f;
constexpr float INF = std::numeric_limits<float>::infinity();

// The reference's f32::sin / cos / acos / atan2 / exp / powf are glibc's float functions (Rust's std calls the platform
// libm).  That is the default here.  -DPYRO_LIBM_DOUBLE builds the variant liboracle_dbl.so in which the SHADING-side calls
// (not the distance estimators) evaluate in double and round once - the very expression the device code uses (core.cuh
// m_sin ...), so that "GPU vs oracle" differences caused by last-bit libm differences can be told apart from real ones:
// tests/test_gpu_parity.py compares the GPU with both variants and tools/first_divergence.py prints the first differing value.
#ifdef PYRO_LIBM_DOUBLE
inline float m_sin(float x) { return (float)std::sin((double)x); }
inline float m_cos(float x) { return (float)std::cos((double)x); }
inline float m_acos(float x) { return (float)std::acos((double)x); }
inline float m_atan2(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float m_exp(float x) { return (float)std::exp((double)x); }
inline float m_pow(float x, float y) { return (float)std::pow((double)x, (double)y); }
constexpr int LIBM_MODE = 1;
#else
inline float m_sin(float x) { return sinf(x); }
inline float m_cos(float x) { return cosf(x); }
inline float m_acos(float x) { return acosf(x); }
inline float m_atan2(float y, float x) { return atan2f(y, x); }
inline float m_exp(float x) { return expf(x); }
inline float m_pow(float x, float y) { return powf(x, y); }
constexpr int LIBM_MODE = 0;
#endif

// f32::min / f32::max (IEEE minNum/maxNum: a NaN operand yields the other one).
inline float fmin_(float a, float b) { return fminf(a, b); }
inline float fmax_(float a, float b) { return fmaxf(a, b); }

// Rust `f32 as usize`: saturating, NaN -> 0.
inline size_t f32_as_usize(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}
// Rust `f64 as u16` (saturating) - expressions.rs:303
inline uint16_t f64_as_u16(double v) {
    if (!(v > 0.0)) return 0;
    if (v >= 65535.0) return 65535;
    return (uint16_t)v;
}
inline int64_t f32_as_isize(float v) {
    if (v != v) return 0;
    if (v >= 9223372036854775808.0f) return INT64_MAX;
    if (v <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)v;
}

struct Vec2 {
    float x = 0, y = 0;
};

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// cgmath: dot = mul_element_wise().sum() = (x*x' + y*y') + z*z'
inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float magnitude2(Vec3 a) { return dot(a, a); }
inline float magnitude(Vec3 a) { return sqrtf(dot(a, a)); }
// cgmath InnerSpace::normalize_to: v * (m / |v|); normalize = normalize_to(1)
inline Vec3 normalize_to(Vec3 a, float m) { return a * (m / magnitude(a)); }
inline Vec3 normalize(Vec3 a) { return normalize_to(a, 1.0f); }

struct Vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    Vec4() = default;
    Vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
};
inline Vec4 operator+(Vec4 a, Vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline Vec4 operator-(Vec4 a, Vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline Vec4 operator*(Vec4 a, Vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline Vec4 operator/(Vec4 a, Vec4 b) { return {a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
inline Vec4 operator*(Vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }

// cgmath Quaternion{s, v}; Quaternion::new(w, xi, yj, zk)
struct Quat {
    float s = 1, x = 0, y = 0, z = 0;
    Quat() = default;
    Quat(float s_, float x_, float y_, float z_) : s(s_), x(x_), y(y_), z(z_) {}
    Vec3 v() const { return {x, y, z}; }
};
inline Quat operator+(Quat a, Quat b) { return {a.s + b.s, a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Quat operator*(Quat a, float f) { return {a.s * f, a.x * f, a.y * f, a.z * f}; }
// Hamilton product, cgmath's scalar (non-SIMD) expansion order.
inline Quat operator*(Quat l, Quat r) {
    return {l.s * r.s - l.x * r.x - l.y * r.y - l.z * r.z,
            l.s * r.x + l.x * r.s + l.y * r.z - l.z * r.y,
            l.s * r.y + l.y * r.s + l.z * r.x - l.x * r.z,
            l.s * r.z + l.z * r.s + l.x * r.y - l.y * r.x};
}
inline float magnitude(Quat q) { return sqrtf(q.s * q.s + dot(q.v(), q.v())); }
inline Quat normalize(Quat q) { return q * (1.0f / magnitude(q)); }
inline Quat conjugate(Quat q) { return {q.s, -q.x, -q.y, -q.z}; }
// Quaternion * Vector3 (shapes/mod.rs:565,569)
inline Vec3 rotate(Quat q, Vec3 v) {
    Vec3 tmp = cross(q.v(), v) + v * q.s;
    return cross(q.v(), tmp) * 2.0f + v;
}

// Column-major 3x3, m[c][r] like cgmath.
struct Mat3 {
    float m[3][3];
    static Mat3 from_cols(Vec3 c0, Vec3 c1, Vec3 c2) {
        Mat3 r;
        r.m[0][0] = c0.x; r.m[0][1] = c0.y; r.m[0][2] = c0.z;
        r.m[1][0] = c1.x; r.m[1][1] = c1.y; r.m[1][2] = c1.z;
        r.m[2][0] = c2.x; r.m[2][1] = c2.y; r.m[2][2] = c2.z;
        return r;
    }
    Vec3 row(int r) const { return {m[0][r], m[1][r], m[2][r]}; }
    Vec3 col(int c) const { return {m[c][0], m[c][1], m[c][2]}; }
    float determinant() const {
        return m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) -
               m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) +
               m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    }
};
inline Mat3 operator*(const Mat3& l, const Mat3& r) {
    Mat3 o;
    for (int c = 0; c < 3; ++c)
        for (int rr = 0; rr < 3; ++rr) o.m[c][rr] = dot(l.row(rr), r.col(c));
    return o;
}
// Matrix3::from_angle_x / from_angle_y (shapes/mod.rs:357-358)
inline Mat3 mat3_from_angle_x(float theta) {
    float s = m_sin(theta), c = m_cos(theta);
    return Mat3::from_cols({1, 0, 0}, {0, c, s}, {0, -s, c});
}
inline Mat3 mat3_from_angle_y(float theta) {
    float s = m_sin(theta), c = m_cos(theta);
    return Mat3::from_cols({c, 0, -s}, {0, 1, 0}, {s, 0, c});
}
// Quaternion::from(Matrix3) - Shoemake, as in cgmath 0.17 (world.rs:100,357-367; shapes/mod.rs:546,580)
inline Quat quat_from_mat3(const Mat3& a) {
    const float(*m)[3] = a.m;
    float trace = m[0][0] + m[1][1] + m[2][2];
    const float half = 0.5f;
    if (trace >= 0.0f) {
        float s = sqrtf(1.0f + trace);
        float w = half * s;
        s = half / s;
        return {w, (m[1][2] - m[2][1]) * s, (m[2][0] - m[0][2]) * s, (m[0][1] - m[1][0]) * s};
    } else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) {
        float s = sqrtf((m[0][0] - m[1][1] - m[2][2]) + 1.0f);
        float x = half * s;
        s = half / s;
        return {(m[1][2] - m[2][1]) * s, x, (m[1][0] + m[0][1]) * s, (m[0][2] + m[2][0]) * s};
    } else if (m[1][1] > m[2][2]) {
        float s = sqrtf((m[1][1] - m[0][0] - m[2][2]) + 1.0f);
        float y = half * s;
        s = half / s;
        return {(m[2][0] - m[0][2]) * s, (m[1][0] + m[0][1]) * s, y, (m[2][1] + m[1][2]) * s};
    } else {
        float s = sqrtf((m[2][2] - m[0][0] - m[1][1]) + 1.0f);
        float z = half * s;
        s = half / s;
        return {(m[0][1] - m[1][0]) * s, (m[0][2] + m[2][0]) * s, (m[2][1] + m[1][2]) * s, z};
    }
}

// Column-major 4x4, c[col] like cgmath.
struct Mat4 {
    Vec4 c[4];
    static Mat4 identity() {
        Mat4 r;
        r.c[0] = {1, 0, 0, 0}; r.c[1] = {0, 1, 0, 0}; r.c[2] = {0, 0, 1, 0}; r.c[3] = {0, 0, 0, 1};
        return r;
    }
    float at(int col, int row) const {
        const Vec4& v = c[col];
        return row == 0 ? v.x : row == 1 ? v.y : row == 2 ? v.z : v.w;
    }
};
// M * Vector4 = c0*x + c1*y + c2*z + c3*w (left to right)
inline Vec4 mul(const Mat4& m, Vec4 v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }
// Transform::transform_point: from_homogeneous(M * (p,1)) = xyz * (1/w)
inline Vec3 transform_point(const Mat4& m, Vec3 p) {
    Vec4 h = mul(m, {p.x, p.y, p.z, 1.0f});
    float inv = 1.0f / h.w;
    return {h.x * inv, h.y * inv, h.z * inv};
}
inline Vec3 transform_vector(const Mat4& m, Vec3 v) {
    Vec4 h = mul(m, {v.x, v.y, v.z, 0.0f});
    return {h.x, h.y, h.z};
}
// Matrix4::look_at(eye, center, up) (right-handed) - project/mod.rs:263
inline Mat4 look_at(Vec3 eye, Vec3 center, Vec3 up) {
    Vec3 f = normalize(center - eye);
    Vec3 s = normalize(cross(f, up));
    Vec3 u = cross(s, f);
    Mat4 r;
    r.c[0] = {s.x, u.x, -f.x, 0};
    r.c[1] = {s.y, u.y, -f.y, 0};
    r.c[2] = {s.z, u.z, -f.z, 0};
    r.c[3] = {-dot(eye, s), -dot(eye, u), dot(eye, f), 1};
    return r;
}
// General 4x4 inverse by cofactors of the transpose, inv = adj / det (project/mod.rs:264, cameras.rs:112).
// cgmath's exact expression order for `determinant` is not recalled; DESIGN.md §2 fixes this
// formulation (Laplace expansion along the first row) for both oracle and product.
inline bool invert(const Mat4& a, Mat4& out) {
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
    for (int col = 0; col < 4; ++col) {
        float cof = minor3(col, 0) * ((col & 1) ? -1.0f : 1.0f);
        det += a.at(col, 0) * cof;
    }
    if (det == 0.0f || det != det) return false;
    float inv_det = 1.0f / det;
    float o[4][4];
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            // inverse[col][row] = cofactor(a, row<-col swap) / det
            float cof = minor3(row, col) * (((row + col) & 1) ? -1.0f : 1.0f);
            o[col][row] = cof * inv_det;
        }
    for (int col = 0; col < 4; ++col) out.c[col] = {o[col][0], o[col][1], o[col][2], o[col][3]};
    return true;
}

struct Ray {
    Vec3 origin, direction;
};
// collision Ray::transform (cameras.rs:94)
inline Ray transform_ray(const Mat4& m, Ray r) { return {transform_point(m, r.origin), transform_vector(m, r.direction)}; }

// collision Aabb3
struct Aabb {
    Vec3 min, max;
    static Aabb from_points(Vec3 a, Vec3 b) {
        return {{a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z},
                {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}};
    }
    Aabb grow(Vec3 p) const {
        return {{min.x < p.x ? min.x : p.x, min.y < p.y ? min.y : p.y, min.z < p.z ? min.z : p.z},
                {max.x > p.x ? max.x : p.x, max.y > p.y ? max.y : p.y, max.z > p.z ? max.z : p.z}};
    }
    Aabb unite(const Aabb& o) const { return grow(o.min).grow(o.max); }
    Vec3 dim() const { return max - min; }
    Vec3 center() const { return min + dim() / 2.0f; }
    float surface_area() const {
        Vec3 d = dim();
        return 2.0f * ((d.x * d.y) + (d.x * d.z) + (d.y * d.z));
    }
};

// rand_xorshift 0.3.0 XorShiftRng + the rand 0.8.5 distributions pyrite draws from.
struct XorShift {
    uint32_t x, y, z, w;
    uint32_t next_u32() {
        uint32_t t = x ^ (x << 11);
        x = y; y = z; z = w;
        w = w ^ (w >> 19) ^ (t ^ (t >> 8));
        return w;
    }
    uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return (hi << 32) | lo;
    }
    // Standard: f32 in [0,1) from the top 24 bits
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    // UniformFloat::sample_single, 23 mantissa bits, retry with a shrunken scale on res >= high
    float gen_range_f32(float low, float high) {
        float scale = high - low;
        for (;;) {
            uint32_t bits = (next_u32() >> 9) | 0x3f800000u;
            float v12;
            memcpy(&v12, &bits, 4);
            float v01 = v12 - 1.0f;
            float res = v01 * scale + low;
            if (res < high) return res;
            scale = nextafterf(scale, -INF);
        }
    }
    // UniformInt<usize>::sample_single: 64-bit widening multiply with a conservative rejection zone
    uint64_t gen_range_usize(uint64_t n) {
        uint64_t range = n;
        uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            uint64_t v = next_u64();
            unsigned __int128 m = (unsigned __int128)v * range;
            uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
            if (lo <= zone) return hi;
        }
    }
    // SliceRandom::choose -> gen_index -> UniformInt<u32>::sample_single (one u32 per attempt)
    uint32_t gen_index_u32(uint32_t n) {
        uint32_t range = n;
        uint32_t zone = (range << __builtin_clz(range)) - 1;
        for (;;) {
            uint32_t v = next_u32();
            uint64_t m = (uint64_t)v * range;
            uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
            if (lo <= zone) return hi;
        }
    }
};

// Per-path-sample stream key -> Xorshift128 seed (repo-defined; the reference seeds each tile
// from the OS, simple.rs:26-28, so only distribution and draw order matter).  splitmix64 over
// (seed, tile, sample); identical in pyrite_b200/csrc (device) - stated in DESIGN.md §5.
inline uint64_t splitmix64(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline XorShift keyed_rng(uint64_t seed, uint64_t tile, uint64_t sample) {
    uint64_t s = seed ^ (tile * 0xD1342543DE82EF95ull) ^ (sample * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull);
    uint64_t a = splitmix64(s), b = splitmix64(s);
    XorShift r{(uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32)};
    if ((r.x | r.y | r.z | r.w) == 0) r.w = 0x113ba7bbu;
    return r;
}

}  // namespace pyro
