// Device-side building blocks of the render path: f32 vector math with the reference's operation
// order (no FMA contraction: this file is compiled with -fmad=false), the Xorshift128 stream and
// the rand 0.8 distributions, spectra / textures / the expression VM, the ray-shape tests and the
// BVH walk.  Every function is __host__ __device__ so that tests/host_emu.cpp can run the very
// same code on the CPU next to the oracle; the product only ever calls it from kernels.
#pragma once
#include <math.h>

#include "device_types.h"

namespace pyr {

#define PYR_PI 3.14159265358979323846f
#define PYR_FRAC_1_PI 0.318309886183790671537767526745028724f
#if defined(__CUDA_ARCH__)
#define PYR_INF __int_as_float(0x7f800000)
#else
#define PYR_INF __builtin_inff()
#endif

// ---------------------------------------------------------------- 32-byte records
// Path headers, rays, hits and pending lights are 32-byte aligned records that one thread moves whole.  sm_100
// has 256-bit global loads / stores (LDG.E.256 / STG.E.256): one request and one full 32-byte sector per record
// instead of two half-sector 128-bit accesses.  `_stream` marks data that is written once and read once by the
// next kernel (evict-first).
struct alignas(32) Vec8 { float v[8]; };
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ Vec8 ld256(const void* p) {
    Vec8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
#ifndef PYR_STREAM_LD
#define PYR_STREAM_LD 1   // A/B on a B200: shade stage -2.2 % on C2, -1 % on C5 against ld.global.cs
#endif
__device__ __forceinline__ Vec8 ld256_stream(const void* p) {
    Vec8 r;
#if PYR_STREAM_LD == 1   // do not allocate the line in L1 at all: L1 stays with the scene tables and the thread-local lines
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ Vec8 ld256_readonly(const void* p) {  // scene data: the non-coherent path, like __ldg
    Vec8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(void* p, const Vec8& r) {
    asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]), "l"(p) : "memory");
}
__device__ __forceinline__ void st256_stream(void* p, const Vec8& r) {
    asm volatile("st.global.cs.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]), "l"(p) : "memory");
}
#else
inline Vec8 ld256(const void* p) { Vec8 r; __builtin_memcpy(&r, p, 32); return r; }
inline Vec8 ld256_stream(const void* p) { return ld256(p); }
inline Vec8 ld256_readonly(const void* p) { return ld256(p); }
inline void st256(void* p, const Vec8& r) { __builtin_memcpy(p, &r, 32); }
inline void st256_stream(void* p, const Vec8& r) { st256(p, r); }
#endif
template <class T> PYR_HD T from_vec8(const Vec8& v) {
    static_assert(sizeof(T) == 32, "32-byte record");
    T r;
    __builtin_memcpy(&r, &v, 32);
    return r;
}
template <class T> PYR_HD Vec8 to_vec8(const T& r) {
    static_assert(sizeof(T) == 32, "32-byte record");
    Vec8 v;
    __builtin_memcpy(&v, &r, 32);
    return v;
}
template <class T> PYR_HD T load_record(const T* p) { return from_vec8<T>(ld256(p)); }
template <class T> PYR_HD T load_record_stream(const T* p) { return from_vec8<T>(ld256_stream(p)); }
template <class T> PYR_HD void store_record(T* p, const T& r) { st256(p, to_vec8(r)); }
template <class T> PYR_HD void store_record_stream(T* p, const T& r) { st256_stream(p, to_vec8(r)); }

// ---------------------------------------------------------------- libm
// The reference's f32::sin / cos / acos / atan2 / exp / powf resolve to glibc's float functions, which
// are correctly rounded in all but rare cases; CUDA's float versions are 1-2 ULP.  On the device the
// shading-side calls are therefore evaluated in double and rounded once, which reproduces glibc's
// result for almost every argument and keeps films comparable pixel by pixel.  (The distance
// estimators keep the float functions: they are the FP32/SFU-bound inner loop of sphere tracing.)
#if defined(__CUDA_ARCH__)
PYR_HD float m_sin(float x) { return (float)sin((double)x); }
PYR_HD float m_cos(float x) { return (float)cos((double)x); }
PYR_HD void m_sincos(float x, float& s, float& c) { double ds, dc; sincos((double)x, &ds, &dc); s = (float)ds; c = (float)dc; }  // one range reduction for both
PYR_HD float m_acos(float x) { return (float)acos((double)x); }
PYR_HD float m_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
PYR_HD float m_exp(float x) { return (float)exp((double)x); }
PYR_HD float m_pow(float x, float y) { return (float)pow((double)x, (double)y); }
#else
PYR_HD float m_sin(float x) { return sinf(x); }
PYR_HD float m_cos(float x) { return cosf(x); }
PYR_HD void m_sincos(float x, float& s, float& c) { s = sinf(x); c = cosf(x); }
PYR_HD float m_acos(float x) { return acosf(x); }
PYR_HD float m_atan2(float y, float x) { return atan2f(y, x); }
PYR_HD float m_exp(float x) { return expf(x); }
PYR_HD float m_pow(float x, float y) { return powf(x, y); }
#endif

// ---------------------------------------------------------------- vectors (cgmath 0.17 order)
struct v3 { float x, y, z; };
PYR_HD v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
PYR_HD v3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
PYR_HD v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PYR_HD v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PYR_HD v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
PYR_HD v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PYR_HD v3 operator/(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
PYR_HD float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
PYR_HD v3 cross(v3 a, v3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PYR_HD float length2(v3 a) { return dot(a, a); }
PYR_HD float length(v3 a) { return sqrtf(dot(a, a)); }
PYR_HD v3 normalize_to(v3 a, float m) { return a * (m / length(a)); }  // cgmath InnerSpace::normalize_to
PYR_HD v3 normalize(v3 a) { return normalize_to(a, 1.0f); }

PYR_HD f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
PYR_HD f4 add4(f4 a, f4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
PYR_HD f4 sub4(f4 a, f4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
PYR_HD f4 mul4(f4 a, f4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
PYR_HD f4 div4(f4 a, f4 b) { return mk4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
PYR_HD f4 scale4(f4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }

// quaternions are f4 = (s, x, y, z)
PYR_HD f4 qmul(f4 l, f4 r) {  // Hamilton product, cgmath's scalar expansion
    return mk4(l.x * r.x - l.y * r.y - l.z * r.z - l.w * r.w, l.x * r.y + l.y * r.x + l.z * r.w - l.w * r.z,
               l.x * r.z + l.z * r.x + l.w * r.y - l.y * r.w, l.x * r.w + l.w * r.x + l.y * r.z - l.z * r.y);
}
PYR_HD float qlength(f4 q) { return sqrtf(q.x * q.x + dot(mk3(q.y, q.z, q.w), mk3(q.y, q.z, q.w))); }
PYR_HD f4 qnormalize(f4 q) { return scale4(q, 1.0f / qlength(q)); }
PYR_HD f4 qconj(f4 q) { return mk4(q.x, -q.y, -q.z, -q.w); }
PYR_HD v3 qrotate(f4 q, v3 v) {  // Quaternion * Vector3 (shapes/mod.rs:565,569)
    v3 qv = mk3(q.y, q.z, q.w);
    v3 tmp = cross(qv, v) + v * q.x;
    return cross(qv, tmp) * 2.0f + v;
}
// Quaternion::from(Matrix3) given the three columns (cgmath 0.17, Shoemake)
PYR_HD f4 quat_from_cols(v3 c0, v3 c1, v3 c2) {
    const float m00 = c0.x, m01 = c0.y, m02 = c0.z, m10 = c1.x, m11 = c1.y, m12 = c1.z, m20 = c2.x, m21 = c2.y, m22 = c2.z;
    float trace = m00 + m11 + m22;
    if (trace >= 0.0f) {
        float s = sqrtf(1.0f + trace);
        float w = 0.5f * s;
        s = 0.5f / s;
        return mk4(w, (m12 - m21) * s, (m20 - m02) * s, (m01 - m10) * s);
    }
    if (m00 > m11 && m00 > m22) {
        float s = sqrtf((m00 - m11 - m22) + 1.0f);
        float x = 0.5f * s;
        s = 0.5f / s;
        return mk4((m12 - m21) * s, x, (m10 + m01) * s, (m02 + m20) * s);
    }
    if (m11 > m22) {
        float s = sqrtf((m11 - m00 - m22) + 1.0f);
        float y = 0.5f * s;
        s = 0.5f / s;
        return mk4((m20 - m02) * s, (m10 + m01) * s, y, (m21 + m12) * s);
    }
    float s = sqrtf((m22 - m00 - m11) + 1.0f);
    float z = 0.5f * s;
    s = 0.5f / s;
    return mk4((m01 - m10) * s, (m02 + m20) * s, (m21 + m12) * s, z);
}

// column-major 4x4 (m[col*4+row]): M * (v, w), summed left to right
PYR_HD f4 mat_mul(const float* m, f4 v) {
    return mk4(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w, ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w,
               ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w, ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w);
}
PYR_HD v3 transform_point(const float* m, v3 p) {
    f4 h = mat_mul(m, mk4(p.x, p.y, p.z, 1.0f));
    float inv = 1.0f / h.w;
    return mk3(h.x * inv, h.y * inv, h.z * inv);
}
PYR_HD v3 transform_vector(const float* m, v3 v) {
    f4 h = mat_mul(m, mk4(v.x, v.y, v.z, 0.0f));
    return mk3(h.x, h.y, h.z);
}

// Rust `f32 as usize` / `as isize`: saturating, NaN -> 0
PYR_HD uint64_t f32_as_usize(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return ~0ull;
    return (uint64_t)v;
}
PYR_HD int64_t f32_as_isize(float v) {
    if (v != v) return 0;
    if (v >= 9223372036854775808.0f) return 0x7fffffffffffffffll;
    if (v <= -9223372036854775808.0f) return (int64_t)0x8000000000000000ull;
    return (int64_t)v;
}
PYR_HD uint32_t f_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; __builtin_memcpy(&u, &f, 4); return u;
#endif
}
PYR_HD float bits_f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}

// ---------------------------------------------------------------- RNG (rand_xorshift 0.3 / rand 0.8.5)
struct Rng {
    uint32_t x, y, z, w;
    PYR_HD uint32_t next_u32() {
        uint32_t t = x ^ (x << 11);
        x = y; y = z; z = w;
        w = w ^ (w >> 19) ^ (t ^ (t >> 8));
        return w;
    }
    PYR_HD uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return (hi << 32) | lo;
    }
    PYR_HD float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    PYR_HD float gen_range_f32(float low, float high) {
        float scale = high - low;
        for (;;) {
            float v12 = bits_f((next_u32() >> 9) | 0x3f800000u);
            float v01 = v12 - 1.0f;
            float res = v01 * scale + low;
            if (res < high) return res;
            scale = bits_f(f_bits(scale) - 1u);  // next float towards -inf of a positive scale
        }
    }
    PYR_HD uint64_t gen_range_usize(uint64_t range) {
#if defined(__CUDA_ARCH__)
        int lz = __clzll((long long)range);
#else
        int lz = __builtin_clzll(range);
#endif
        uint64_t zone = (range << lz) - 1;
        for (;;) {
            uint64_t v = next_u64();
#if defined(__CUDA_ARCH__)
            uint64_t hi = __umul64hi(v, range);
            uint64_t lo = v * range;
#else
            unsigned __int128 m = (unsigned __int128)v * range;
            uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
#endif
            if (lo <= zone) return hi;
        }
    }
    PYR_HD uint32_t gen_index_u32(uint32_t range) {
#if defined(__CUDA_ARCH__)
        int lz = __clz((int)range);
#else
        int lz = __builtin_clz(range);
#endif
        uint32_t zone = (range << lz) - 1;
        for (;;) {
            uint32_t v = next_u32();
            uint64_t m = (uint64_t)v * range;
            uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
            if (lo <= zone) return hi;
        }
    }
};
PYR_HD uint64_t splitmix64(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Per-path-sample stream key: the reference seeds one stream per tile from the OS
// (simple.rs:26-28); here each path sample (tile, i) owns a stream so that the wavefront is
// order-independent.  The distributions and the draw order are the reference's.
PYR_HD Rng keyed_rng(uint64_t seed, uint64_t tile, uint64_t sample) {
    uint64_t s = seed ^ (tile * 0xD1342543DE82EF95ull) ^ (sample * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull);
    uint64_t a = splitmix64(s), b = splitmix64(s);
    Rng r;
    r.x = (uint32_t)a; r.y = (uint32_t)(a >> 32); r.z = (uint32_t)b; r.w = (uint32_t)(b >> 32);
    if ((r.x | r.y | r.z | r.w) == 0) r.w = 0x113ba7bbu;
    return r;
}

// ---------------------------------------------------------------- math.rs helpers
PYR_HD v3 ortho(v3 v) {  // math.rs:98-113
    v3 unit;
    if (fabsf(v.x) < DIST_EPSILON) unit = mk3(1, 0, 0);
    else if (fabsf(v.y) < DIST_EPSILON) unit = mk3(0, 1, 0);
    else if (fabsf(v.z) < DIST_EPSILON) unit = mk3(0, 0, 1);
    else unit = mk3(-v.y, v.x, 0.0f);
    return cross(v, unit);
}
PYR_HD void basis(v3 x, v3& y, v3& z) {  // math.rs:119-123
    z = normalize(ortho(x));
    y = normalize(cross(z, x));
}
PYR_HD v3 sample_cone(Rng& rng, v3 direction, float cos_half) {  // math.rs:125-137
    v3 o1 = normalize(ortho(direction));
    v3 o2 = normalize(cross(direction, o1));
    float r1 = PYR_PI * 2.0f * rng.gen_f32();
    float r2 = cos_half + (1.0f - cos_half) * rng.gen_f32();
    float oneminus = sqrtf(1.0f - r2 * r2);
    float s1, c1;
    m_sincos(r1, s1, c1);
    return (o1 * c1 * oneminus + o2 * s1 * oneminus) + direction * r2;
}
PYR_HD float solid_angle(float cos_half) { return cos_half >= 1.0f ? 0.0f : 2.0f * PYR_PI * (1.0f - cos_half); }  // math.rs:139-145
PYR_HD v3 sample_sphere(Rng& rng) {  // math.rs:147-153
    float u = rng.gen_f32();
    float v = rng.gen_f32();
    float theta = 2.0f * PYR_PI * u;
    float phi = m_acos(2.0f * v - 1.0f);
    float st, ct, sp, cp;
    m_sincos(theta, st, ct);
    m_sincos(phi, sp, cp);
    return mk3(sp * ct, sp * st, cp);
}
PYR_HD v3 sample_hemisphere(Rng& rng, v3 direction) {  // math.rs:155-164
    v3 s = sample_sphere(rng);
    v3 x = normalize_to(ortho(direction), s.x);
    v3 y = normalize_to(cross(x, direction), s.y);
    v3 z = normalize_to(direction, fabsf(s.z));
    return (x + y) + z;
}
PYR_HD float schlick(float ref_index1, float ref_index2, v3 normal, v3 incident) {  // math.rs:75-96
    float cos_psi = -dot(normal, incident);
    float r0 = (ref_index1 - ref_index2) / (ref_index1 + ref_index2);
    if (ref_index1 > ref_index2) {
        float n = ref_index1 / ref_index2;
        float sin_t2 = n * n * (1.0f - cos_psi * cos_psi);
        if (sin_t2 > 1.0f) return 1.0f;
        cos_psi = sqrtf(1.0f - sin_t2);
    }
    float inv_cos = 1.0f - cos_psi;
    return r0 * r0 + (1.0f - r0 * r0) * inv_cos * inv_cos * inv_cos * inv_cos * inv_cos;
}
PYR_HD float fresnel(float ior, float env_ior, v3 normal, v3 incident) {  // math.rs:167-175
    if (dot(incident, normal) < 0.0f) return schlick(env_ior, ior, normal, incident);
    return schlick(ior, env_ior, -normal, incident);
}
PYR_HD float blackbody(float wavelength, float temperature) {  // math.rs:177-182
    wavelength = wavelength * 1.0e-9f;
    float a2 = wavelength * wavelength;
    float a4 = a2 * a2;
    float p5 = 1.0f / (wavelength * a4);
    float power_term = 3.74183e-16f * p5;
    return power_term / (m_exp(1.4388e-2f / (wavelength * temperature)) - 1.0f);
}

// ---------------------------------------------------------------- spectra and textures
// project/spectra.rs:30-58 `Spectrum::Array::get`
PYR_HD float array_get(const float* pts, uint32_t n, uint32_t stride, float lo, float hi, float w) {
    if (n == 0) return 0.0f;
    if (w <= lo) return pts[0];
    if (w >= hi) return pts[(size_t)(n - 1) * stride];
    float normalized = (w - lo) / (hi - lo);
    float float_index = normalized * ((float)n - 1.0f);
    float min_float_index = truncf(float_index);
    size_t min_index = (size_t)f32_as_usize(min_float_index);
    size_t max_index = min_index + 1;
    float min_value = pts[min_index * stride], max_value = pts[max_index * stride];
    float mix = float_index - min_float_index;
    return min_value * (1.0f - mix) + max_value * mix;
}
// math.rs:21-72 `Interpolated::get`: (x, y) pairs, 0 outside and at the end points
PYR_HD float curve_get(const float* pts, uint32_t n, float input) {
    if (n == 0) return 0.0f;
    uint32_t lo = 0, hi = n - 1;
    if (pts[2 * lo] >= input) return 0.0f;
    if (pts[2 * hi] <= input) return 0.0f;
    while (hi > lo + 1) {
        uint32_t check = (hi + lo) / 2;
        float x = pts[2 * check];
        if (x == input) return pts[2 * check + 1];
        if (x > input) hi = check; else lo = check;
    }
    float min_x = pts[2 * lo], min_y = pts[2 * lo + 1], max_x = pts[2 * hi], max_y = pts[2 * hi + 1];
    if (input < min_x) return 0.0f;
    if (input > max_x) return 0.0f;
    return min_y + (max_y - min_y) * ((input - min_x) / (max_x - min_x));
}
PYR_HD float spectrum_get(const SceneView& sc, uint32_t id, float w) {
    const SpectrumRec s = sc.spectra[id];
    const float* data = sc.spectrum_data + s.offset;
    return s.is_curve ? curve_get(data, s.n, w) : array_get(data, s.n, 1, s.lo, s.hi, w);
}
PYR_HD float cubic_interpolate(float v1, float v2, float v3_, float v4, float pos) {  // texture.rs:324-334
    float a = (v4 - v3_) - (v1 - v2);
    float b = (v1 - v2) - a;
    float c = v3_ - v1;
    float d = v2;
    return d + (c + (b + a * pos) * pos) * pos;
}
// texture.rs:88-148 `Texture::get_color`: wrap-around 4x4 bicubic
PYR_HD void texture_get(const SceneView& sc, uint32_t tex_index, float px, float py, float* out) {
    const TextureRec t = sc.textures[tex_index];
    const float* data = sc.texels + t.offset;
    int64_t width = t.width, height = t.height;
    int channels = (int)t.channels;
    float x = px * (float)width - 0.5f;
    float x_floor = floorf(x);
    int64_t x2 = ((f32_as_isize(x_floor) % width) + width) % width;
    int64_t x1 = x2 == 0 ? width - 1 : x2 - 1;
    int64_t x3 = x2 == width - 1 ? 0 : x2 + 1;
    int64_t x4 = x3 == width - 1 ? 0 : x3 + 1;
    float y = (1.0f - py) * (float)height - 0.5f;
    float y_floor = floorf(y);
    int64_t y2 = ((f32_as_isize(y_floor) % height) + height) % height;
    int64_t y1 = y2 == 0 ? height - 1 : y2 - 1;
    int64_t y3 = y2 == height - 1 ? 0 : y2 + 1;
    int64_t y4 = y3 == height - 1 ? 0 : y3 + 1;
    const int64_t xs[4] = {x1, x2, x3, x4}, ys[4] = {y1, y2, y3, y4};
    float fx = x - x_floor, fy = y - y_floor;
    for (int c = 0; c < channels; ++c) {
        float rows[4];
        for (int r = 0; r < 4; ++r) {
            float v[4];
            for (int k = 0; k < 4; ++k) v[k] = data[(size_t)(xs[k] + ys[r] * width) * channels + c];
            rows[r] = cubic_interpolate(v[0], v[1], v[2], v[3], fx);
        }
        out[c] = cubic_interpolate(rows[0], rows[1], rows[2], rows[3], fy);
    }
}

// ---------------------------------------------------------------- expression VM (program/execution_context.rs:69-283)
struct VmInputs {
    float wavelength;
    v3 normal, incident;
    float tex[2];
};
PYR_HD float burns_rgb_spectrum(const SceneView& sc, f4 rgb, float wavelength) {  // execution_context.rs:140-152
    float r = array_get(sc.burns + 0, sc.burns_t.n, 3, sc.burns_t.lo, sc.burns_t.hi, wavelength);
    float g = array_get(sc.burns + 1, sc.burns_t.n, 3, sc.burns_t.lo, sc.burns_t.hi, wavelength);
    float b = array_get(sc.burns + 2, sc.burns_t.n, 3, sc.burns_t.lo, sc.burns_t.hi, wavelength);
    float rr = rgb.x * r, gg = rgb.y * g, bb = rgb.z * b;
    return (rr + gg) + bb;
}
// The VM register file.  In the kernels it lives in dynamic shared memory, [register][thread] (blocks of
// PYR_BLOCK threads), so that its dynamically indexed accesses stay on chip instead of going through
// thread-local memory; on the host it is a plain array.
#if defined(__CUDA_ARCH__)
constexpr int PYR_BLOCK = 128;
extern __shared__ float4 pyr_dyn_smem[];  // [0, VM_REGS * PYR_BLOCK): VM registers; the wave kernels stage visibility rays after them
struct RegFile {
    f4* base;
    __device__ __forceinline__ f4& operator[](uint32_t i) const { return base[i * PYR_BLOCK]; }
};
#define PYR_REGFILE(R) RegFile R{reinterpret_cast<f4*>(pyr_dyn_smem) + threadIdx.x}
#else
struct RegFile {
    f4* base;
    f4& operator[](uint32_t i) const { return base[i]; }
};
#define PYR_REGFILE(R) f4 R##_storage[VM_REGS]; RegFile R{R##_storage}
#endif

// One instruction as eight 32-bit words (two 16-byte loads on the device).
struct InstrWords { uint32_t head, regs, deps, resource, v[4]; };
PYR_HD InstrWords fetch_instr(const Instr* p) {
    InstrWords w;
#if defined(__CUDA_ARCH__)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    w.head = a.x; w.regs = a.y; w.deps = a.z; w.resource = a.w;
    w.v[0] = b.x; w.v[1] = b.y; w.v[2] = b.z; w.v[3] = b.w;
#else
    __builtin_memcpy(&w, p, sizeof(w));
#endif
    return w;
}
// Runs `count` instructions starting at `code` into R (program/execution_context.rs:69-283).
PYR_HD_NOINLINE void vm_execute(const SceneView& sc, const Instr* code, uint32_t count, const VmInputs& in, RegFile R) {
    for (uint32_t pc = 0; pc < count; ++pc) {
        const InstrWords I = fetch_instr(code + pc);
        const uint32_t op = I.head & 0xffu, vtype = (I.head >> 8) & 0xffu;
#define PYR_NUM(k) (((I.regs >> (8 * (k))) & 1u) ? R[I.v[k] & (VM_REGS - 1)].x : bits_f(I.v[k]))
#define PYR_REG(k) R[I.v[k] & (VM_REGS - 1)]
        f4 out = mk4(0, 0, 0, 0);
        switch (op) {
            case OP_NUMBER: out.x = PYR_NUM(0); break;
            case OP_VECTOR: out = mk4(PYR_NUM(0), PYR_NUM(1), PYR_NUM(2), PYR_NUM(3)); break;
            case OP_RGB: out = mk4(PYR_NUM(0), PYR_NUM(1), PYR_NUM(2), 1.0f); break;
            case OP_SPECTRUM: out.x = spectrum_get(sc, I.resource, in.wavelength); break;
            case OP_COLOR_TEXTURE: {
                float c[4];
                texture_get(sc, I.resource, in.tex[0], in.tex[1], c);
                out = mk4(c[0], c[1], c[2], c[3]);
                break;
            }
            case OP_MONO_TEXTURE: {
                float c[4];
                texture_get(sc, sc.n_color_textures + I.resource, in.tex[0], in.tex[1], c);
                out.x = c[0];
                break;
            }
            case OP_RGB_SPECTRUM: out.x = burns_rgb_spectrum(sc, PYR_REG(0), in.wavelength); break;
            case OP_FRESNEL: out.x = fresnel(PYR_NUM(0), PYR_NUM(1), in.normal, in.incident); break;
            case OP_BLACKBODY: out.x = blackbody(in.wavelength, PYR_NUM(0)); break;
            case OP_NUM_TO_RGB: { float n = PYR_NUM(0); out = mk4(n, n, n, 1.0f); break; }
            case OP_NUM_TO_VEC: { float n = PYR_NUM(0); out = mk4(n, n, n, n); break; }
            case OP_RGB_TO_VEC: {  // execution_context.rs:183-193
                f4 c = PYR_REG(0);
                out = mk4((c.x * 2.0f) - 1.0f, (c.y * 2.0f) - 1.0f, (c.z * 2.0f) - 1.0f, (c.w * 2.0f) - 1.0f);
                break;
            }
            case OP_BINARY: {  // execution_context.rs:228-268
                const f4 l = PYR_REG(0), r = PYR_REG(1);
                switch ((I.head >> 16) & 0xffu) {
                    case 0: out = add4(l, r); break;
                    case 1: out = sub4(l, r); break;
                    case 2: out = mul4(l, r); break;
                    default: out = div4(l, r); break;
                }
                if (vtype == VT_NUMBER) { out.y = 0; out.z = 0; out.w = 0; }
                break;
            }
            case OP_MIX: {  // execution_context.rs:195-227
                float amount = fmaxf(fminf(PYR_NUM(0), 1.0f), 0.0f);
                const f4 l = PYR_REG(1), r = PYR_REG(2);
                if (vtype == VT_NUMBER) out.x = l.x * (1.0f - amount) + r.x * amount;
                else out = add4(l, scale4(sub4(r, l), amount));
                break;
            }
            case OP_CLAMP: out.x = fmaxf(fminf(PYR_NUM(0), PYR_NUM(2)), PYR_NUM(1)); break;
            default: break;
        }
#undef PYR_NUM
#undef PYR_REG
        R[(I.head >> 24) & (VM_REGS - 1)] = out;
    }
}
// ExecutionContext::run for T = f32 (execution_context.rs:29-56).  `rerun`: the previous call on
// the same R was the same program with the same inputs except the wavelength, so only the
// wavelength-dependent instructions run again (MemoizedContext, execution_context.rs:310-342).
PYR_HD float run_program(const SceneView& sc, const ProgramRec& p, const VmInputs& in, RegFile R, bool rerun) {
    if (p.is_constant) return p.value;
    if (rerun) vm_execute(sc, sc.code + p.wl_offset, p.wl_count, in, R);
    else vm_execute(sc, sc.code + p.code_offset, p.n_instr, in, R);
    return R[p.out_reg & (VM_REGS - 1)].x;
}
PYR_HD float run_number(const SceneView& sc, int32_t program, const VmInputs& in, RegFile R, bool rerun = false) {
    const ProgramRec p = sc.programs[program];
    return run_program(sc, p, in, R, rerun);
}
// ... and T = Vector (the compiler already appended the output conversion, compiler.rs:532-567)
PYR_HD f4 run_vector(const SceneView& sc, int32_t program, const VmInputs& in, RegFile R) {
    const ProgramRec p = sc.programs[program];
    if (p.is_constant) return mk4(p.value, p.value, p.value, p.value);
    vm_execute(sc, sc.code + p.code_offset, p.n_instr, in, R);
    return R[p.out_reg & (VM_REGS - 1)];
}
PYR_HD bool program_reads_wavelength(const SceneView& sc, int32_t program) {
    const ProgramRec p = sc.programs[program];
    return !p.is_constant && (p.reads & IN_WAVELENGTH);
}

// ---------------------------------------------------------------- ray / shape tests
// math.rs:184-207 with the reciprocal direction hoisted (same values: 1/dir is a pure function)
PYR_HD bool slab_test(v3 lo, v3 hi, v3 o, v3 inv, float& dist) {
    float t1 = (lo.x - o.x) * inv.x;
    float t2 = (hi.x - o.x) * inv.x;
    float tmin = fminf(t1, t2), tmax = fmaxf(t1, t2);
    t1 = (lo.y - o.y) * inv.y;
    t2 = (hi.y - o.y) * inv.y;
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
    t1 = (lo.z - o.z) * inv.z;
    t2 = (hi.z - o.z) * inv.z;
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
    if (tmax >= tmin && tmax >= 0.0f) { dist = fmaxf(tmin, 0.0f); return true; }
    return false;
}
// Moeller-Trumbore, shapes/mod.rs:75-119
PYR_HD bool triangle_test(v3 v1, v3 e1, v3 e2, v3 o, v3 d, float& dist, float& u_out, float& v_out) {
    v3 p = cross(d, e2);
    float det = dot(e1, p);
    if (det > -DIST_EPSILON && det < DIST_EPSILON) return false;
    float inv_det = 1.0f / det;
    v3 t = o - v1;
    float u = dot(t, p) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    v3 q = cross(t, e1);
    float v = dot(d, q) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    float ds = dot(e2, q) * inv_det;
    if (ds > DIST_EPSILON) { dist = ds; u_out = u; v_out = v; return true; }
    return false;
}
// collision::Sphere x Ray (near root only); distance = |hit - origin| (shapes/mod.rs:57-74)
PYR_HD bool sphere_test(v3 center, float radius, v3 o, v3 d, float& dist, v3& point) {
    v3 l = center - o;
    float tca = dot(l, d);
    if (tca < 0.0f) return false;
    float d2 = dot(l, l) - tca * tca;
    if (d2 > radius * radius) return false;
    float thc = sqrtf(radius * radius - d2);
    point = o + d * (tca - thc);
    dist = length(point - o);
    return true;
}
// collision::Plane x Ray: t = -(d + o.n) / (dir.n), None iff t < 0 (shapes/mod.rs:441-452)
PYR_HD bool plane_test(const PlaneRec& pl, v3 o, v3 d, float& dist, v3& point) {
    v3 n = ld3(pl.n);
    float t = -(pl.d + dot(o, n)) / dot(d, n);
    if (t < 0.0f) return false;
    point = o + d * t;
    dist = length(point - o);
    return true;
}

// ---------------------------------------------------------------- distance estimators (shapes/distance_estimators.rs)
PYR_HD f4 bicomplex_mul(f4 a, f4 b) {  // :96-107
    float x1 = a.x, x2 = b.x, y1 = a.y, y2 = b.y, z1 = a.z, z2 = b.z, w1 = a.w, w2 = b.w;
    return mk4(x1 * x2 - y1 * y2 - z1 * z2 + w1 * w2, x1 * y2 + y1 * x2 - z1 * w2 - w1 * z2, x1 * z2 - y1 * w2 + z1 * x2 - w1 * y2,
               x1 * w2 + y1 * z2 + z1 * y2 + w1 * x2);
}
PYR_HD float estimate_distance(const MarchedRec& m, v3 point, uint32_t& iters) {
    if (m.estimator == 0) {  // Mandelbulb::get :12-42
        v3 z = point;
        float r = 0.0f, dr = 1.0f;
        float dc = m.has_constant ? 0.0f : 1.0f;
        v3 c = m.has_constant ? ld3(m.mb_constant) : point;
        for (uint32_t i = 0; i < m.iterations; ++i) {
            r = length(z);
            if (r > m.threshold) break;
            ++iters;
            float theta = acosf(z.z / r);
            float phi = atan2f(z.y, z.x);
            dr = powf(r, m.power - 1.0f) * m.power * dr + dc;
            float zr = powf(r, m.power);
            theta *= m.power;
            phi *= m.power;
            z = mk3(zr * sinf(theta) * cosf(phi), zr * sinf(phi) * sinf(theta), zr * cosf(theta));
            z = z + c;
        }
        return 0.5f * logf(r) * r / dr;
    }
    // QuaternionJulia::get :52-70
    f4 z = mk4(point.x, point.y, point.z, m.slice_plane);
    float r = 0.0f;
    f4 dz = mk4(1.0f, 0.0f, 0.0f, 0.0f);
    for (uint32_t i = 0; i < m.iterations; ++i) {
        r = qlength(z);
        if (r > m.threshold) break;
        ++iters;
        if (m.variant == 0) { dz = scale4(qmul(dz, z), 2.0f); z = qmul(z, z); }
        else if (m.variant == 1) { dz = scale4(qmul(qmul(dz, z), z), 3.0f); z = qmul(qmul(z, z), z); }
        else { dz = scale4(bicomplex_mul(bicomplex_mul(dz, z), z), 2.0f); z = bicomplex_mul(z, z); }
        z = add4(z, m.constant);
    }
    return 0.5f * logf(r) * r / qlength(dz);
}
// BoundingVolume::intersect (shapes/mod.rs:591-682); Box widens t_max
PYR_HD bool bounds_test(const MarchedRec& m, v3 o, v3 d, float& t_min_out, float& t_max_out) {
    if (m.bounds_type == 0) {
        v3 lo = ld3(m.ba), hi = ld3(m.bb);
        v3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float t_min, t_max;
        if (inv.x < 0.0f) { t_min = (hi.x - o.x) * inv.x; t_max = (lo.x - o.x) * inv.x; }
        else { t_min = (lo.x - o.x) * inv.x; t_max = (hi.x - o.x) * inv.x; }
        float ty_min, ty_max;
        if (inv.y < 0.0f) { ty_min = (hi.y - o.y) * inv.y; ty_max = (lo.y - o.y) * inv.y; }
        else { ty_min = (lo.y - o.y) * inv.y; ty_max = (hi.y - o.y) * inv.y; }
        if (t_min > ty_max || ty_min > t_max) return false;
        if (ty_min > t_min) t_min = ty_min;
        if (ty_max > t_max) t_max = ty_max;
        float tz_min, tz_max;
        if (inv.z < 0.0f) { tz_min = (hi.z - o.z) * inv.z; tz_max = (lo.z - o.z) * inv.z; }
        else { tz_min = (lo.z - o.z) * inv.z; tz_max = (hi.z - o.z) * inv.z; }
        if (t_min > tz_max || tz_min > t_max) return false;
        if (tz_min > t_min) t_min = tz_min;
        if (tz_max > t_max) t_max = tz_max;
        t_min = fmaxf(t_min, 0.0f);
        if (t_min < t_max) { t_min_out = t_min; t_max_out = t_max; return true; }
        return false;
    }
    v3 c = ld3(m.ba);
    v3 l = c - o;
    float tca = dot(l, d);
    if (tca < 0.0f) return false;
    float d2 = dot(l, l) - tca * tca;
    if (d2 > m.bradius * m.bradius) return false;
    float thc = sqrtf(m.bradius * m.bradius - d2);
    t_min_out = tca - thc;
    t_max_out = tca + thc;
    return true;
}
PYR_HD v3 bounds_center(const MarchedRec& m) { return m.bounds_type == 0 ? (ld3(m.ba) + ld3(m.bb)) * 0.5f : ld3(m.ba); }
// Shape::RayMarched branch of ray_intersect (shapes/mod.rs:120-154)
PYR_HD bool march_test(const MarchedRec& m, v3 o, v3 d, float& dist, uint32_t& evals, uint32_t& iters) {
    float lo, hi;
    if (!bounds_test(m, o, d, lo, hi)) return false;
    v3 origin = o + (-bounds_center(m));
    float total = lo;
    while (total < hi) {
        v3 p = origin + d * total;
        ++evals;
        float distance = estimate_distance(m, p, iters);
        // a step of >= EPSILON that is below half an ulp of `total` would leave it unchanged for ever (the reference's loop,
        // shapes/mod.rs:127-135, never returns on such a ray): it ends the march like a step below EPSILON
        const bool stuck = total + distance == total;
        total += distance;
        if (distance < DIST_EPSILON || total > hi || stuck) break;
    }
    if (total <= hi) { dist = total; return true; }
    return false;
}

// ---------------------------------------------------------------- World::intersect (world.rs:273-299)
struct TraceStats { uint32_t nodes, leaves, de_evals, de_iters, fetches; };

PYR_HD v3 prim_v1(const Prim& p) { return mk3(p.a.x, p.a.y, p.a.z); }
PYR_HD v3 prim_e1(const Prim& p) { return mk3(p.a.w, p.b.x, p.b.y); }
PYR_HD v3 prim_e2(const Prim& p) { return mk3(p.b.z, p.b.w, p.c.x); }
PYR_HD uint32_t prim_kind(const Prim& p) { return f_bits(p.c.y); }
PYR_HD uint32_t prim_object(const Prim& p) { return f_bits(p.c.z); }
PYR_HD uint32_t prim_material(const Prim& p) { return f_bits(p.c.w); }

// Whether a hit at distance t ends a visibility ray (modes 1, 2 of `Ray`).
PYR_HD bool occludes(uint32_t mode, float t, float limit) { return mode == 1 ? (t * t < limit) : (t < limit); }

// The walk visits the same boxes with the same slab arithmetic as the reference's pre-order walk
// (spatial/bvh.rs:206-229) but nearest child first, with a stack.  Result = the lexicographic
// minimum (t, plane-before-leaf, rank) over the leaves whose boxes are hit - the reference's
// answer whenever no leaf lies closer than its own box's entry distance by rounding.
// 16-byte loads through the read-only path on the device
PYR_HD Node4 fetch_node(const Node4* p) {
#if defined(__CUDA_ARCH__)
    const Vec8 a = ld256_readonly(p), b = ld256_readonly(reinterpret_cast<const Vec8*>(p) + 1), c = ld256_readonly(reinterpret_cast<const Vec8*>(p) + 2),
               d = ld256_readonly(reinterpret_cast<const Vec8*>(p) + 3);
    Node4 n;
    n.lo_x = mk4(a.v[0], a.v[1], a.v[2], a.v[3]); n.lo_y = mk4(a.v[4], a.v[5], a.v[6], a.v[7]); n.lo_z = mk4(b.v[0], b.v[1], b.v[2], b.v[3]);
    n.hi_x = mk4(b.v[4], b.v[5], b.v[6], b.v[7]); n.hi_y = mk4(c.v[0], c.v[1], c.v[2], c.v[3]); n.hi_z = mk4(c.v[4], c.v[5], c.v[6], c.v[7]);
    n.child[0] = __float_as_int(d.v[0]); n.child[1] = __float_as_int(d.v[1]); n.child[2] = __float_as_int(d.v[2]); n.child[3] = __float_as_int(d.v[3]);
    return n;
#else
    return *p;
#endif
}
PYR_HD Prim fetch_prim(const Prim* p) {
#if defined(__CUDA_ARCH__)
    const float4* q = reinterpret_cast<const float4*>(p);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    Prim r;
    r.a = mk4(a.x, a.y, a.z, a.w); r.b = mk4(b.x, b.y, b.z, b.w); r.c = mk4(c.x, c.y, c.z, c.w);
    return r;
#else
    return *p;
#endif
}
// Traversal stack in thread-local storage (host emulation, fallback).  An entry is the child code and
// the entry distance of its box, so that a subtree made irrelevant by a closer hit is dropped at pop
// time without fetching its node.
struct LocalStack {
    // Slots below FAST_DEPTH may be written speculatively (Traversal::node_step stores all four children and only moves the
    // stack pointer past the real ones).  Here every slot is alike; the kernels' SharedStack keeps exactly these in shared
    // memory, and the host emulation takes the same code path as the device with the same depth.
    static constexpr int FAST_DEPTH = 20;
    PYR_HD void put_fast(int i, int code, float dist) { codes[i] = code; dists[i] = dist; }
    PYR_HD int code_fast(int i) const { return codes[i]; }
    PYR_HD float dist_fast(int i) const { return dists[i]; }
    int codes[BVH_STACK];
    float dists[BVH_STACK];
    float m;  // max(hit distance, entry distance of its leaf's box) of the current closest hit (Traversal::leaf_step)
    PYR_HD void set_best_m(float v) { m = v; }
    PYR_HD float best_m() const { return m; }
    PYR_HD void put(int i, int code, float dist) { codes[i] = code; dists[i] = dist; }
    PYR_HD int code(int i) const { return codes[i]; }
    PYR_HD float dist(int i) const { return dists[i]; }
};

// World::intersect for one ray as an explicit state machine: begin() tests the planes and the root
// box, every step() processes one interior node or one leaf.  The kernels interleave the steps of 32
// rays per warp and refill finished lanes; trace_ray() below simply runs it to completion.
constexpr float CULL_SLACK = 1.000002f;  // ~16 ulps

template <bool STATS>
struct Traversal {
    v3 o, d, inv;
    uint32_t mode;
    float limit, bound, closest;
    float cull;  // closest widened by a few ulps: boxes are only skipped beyond it (see node_step)
    float t, u, v;
    uint32_t rank, kind;
    int sp, cur;       // cur: Node4 index (>= 0) or leaf code (< 0) to process next
    bool done;
    uint32_t vn, vl, vf, de_evals, de_iters;  // boxes tested, leaves tested, nodes fetched (STATS)
    uint32_t march_mask;  // ray-marched shapes whose leaf was reached: sphere-traced after the walk (resolve_marched)

    template <class Stack>
    PYR_HD void begin(const SceneView& sc, const Ray& ray, Stack& stack) {
        (void)stack;
        o = ld3(ray.o); d = ld3(ray.d);
        mode = ray.mode; limit = ray.limit;
        closest = PYR_INF; cull = PYR_INF;
        // visibility rays: nothing at or beyond `bound` can occlude
        bound = PYR_INF;
        if (mode == 1) bound = limit > 0.0f ? sqrtf(limit) : 0.0f;
        else if (mode == 2) bound = limit;
        // (a visibility ray never takes a closest hit - it ends at its first occluder - so its `cull` is free to carry `bound`:
        // one comparison per box serves both kinds of ray)
        cull = bound;
        t = PYR_INF; u = 0; v = 0; rank = 0xFFFFFFFFu; kind = KIND_MISS;
        vn = 0; vl = 0; vf = 0; de_evals = 0; de_iters = 0; march_mask = 0;
        sp = 0; cur = 0; done = true;
        for (uint32_t i = 0; i < sc.n_planes; ++i) {
            float pt; v3 p;
            if (plane_test(sc.planes[i], o, d, pt, p) && pt > DIST_EPSILON && pt < closest) {
                if (mode != 0) {  // visibility rays only report occluders
                    if (occludes(mode, pt, limit)) { t = pt; rank = i; kind = KIND_PLANE; return; }
                    continue;
                }
                closest = pt; cull = pt * CULL_SLACK; t = pt; rank = i; kind = KIND_PLANE;
            }
        }
        if (sc.n_prims == 0) return;
        inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float dr;
        if (STATS) ++vn;
        if (!slab_test(ld3(sc.root_lo), ld3(sc.root_hi), o, inv, dr) || dr > cull) return;
        cur = sc.root;
        stack.put(0, cur, dr);  // slot `sp` always holds the entry distance of the box of `cur` (read by leaf_step)
        done = false;
    }

    PYR_HD bool at_node() const { return cur >= 0; }
    // pop, dropping entries whose box now starts beyond the closest hit
    template <class Stack>
    PYR_HD void pop(Stack& stack) {
        for (;;) {
            if (sp == 0) { done = true; return; }
            --sp;
            if (Stack::FAST_DEPTH > 0 && sp < Stack::FAST_DEPTH) {  // the rest of the stack is in the fast region: a loop without the region test
                for (;;) {
                    if (!(stack.dist_fast(sp) > cull)) { cur = stack.code_fast(sp); return; }
                    if (sp == 0) { done = true; return; }
                    --sp;
                }
            }
            const float entry_dist = stack.dist(sp);
            if (!(entry_dist > cull)) { cur = stack.code(sp); return; }
        }
    }
    // one 4-wide node (cur >= 0): up to four boxes with the reference's slab arithmetic; descend into the
    // nearest hit child and defer the others, farthest first, so that they pop nearest first
    template <class Stack>
    PYR_HD void node_step(const SceneView& sc, Stack& stack) {
        const Node4 nd = fetch_node(sc.nodes + cur);
        const float lox[4] = {nd.lo_x.x, nd.lo_x.y, nd.lo_x.z, nd.lo_x.w}, loy[4] = {nd.lo_y.x, nd.lo_y.y, nd.lo_y.z, nd.lo_y.w};
        const float loz[4] = {nd.lo_z.x, nd.lo_z.y, nd.lo_z.z, nd.lo_z.w}, hix[4] = {nd.hi_x.x, nd.hi_x.y, nd.hi_x.z, nd.hi_x.w};
        const float hiy[4] = {nd.hi_y.x, nd.hi_y.y, nd.hi_y.z, nd.hi_y.w}, hiz[4] = {nd.hi_z.x, nd.hi_z.y, nd.hi_z.z, nd.hi_z.w};
        float dist[4];
        int code[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float dk;
            bool h = slab_test(mk3(lox[k], loy[k], loz[k]), mk3(hix[k], hiy[k], hiz[k]), o, inv, dk);
            code[k] = nd.child[k];
            // A box is skipped only when it starts beyond the closest hit by more than rounding: a leaf that ties with the
            // current hit (a ray through a shared edge) can have a box entry a few ulps beyond its own hit distance, and the
            // reference, walking in pre-order, would have tested it first (World::intersect's tie rule, world.rs:288-296).
            h = h && code[k] != NODE4_EMPTY && !(dk > cull);
            if (STATS) vn += code[k] != NODE4_EMPTY ? 1u : 0u;
            dist[k] = h ? dk : PYR_INF;
            if (!h) code[k] = NODE4_EMPTY;
        }
        if (STATS) ++vf;
        // sorting network on (dist, code), ascending; misses carry +inf and sink to the end
#define PYR_CSWAP(a, b) { const bool sw = dist[b] < dist[a]; const float da = sw ? dist[b] : dist[a], db = sw ? dist[a] : dist[b]; \
                          const int ca = sw ? code[b] : code[a], cb = sw ? code[a] : code[b]; dist[a] = da; dist[b] = db; code[a] = ca; code[b] = cb; }
        PYR_CSWAP(0, 1) PYR_CSWAP(2, 3) PYR_CSWAP(0, 2) PYR_CSWAP(1, 3) PYR_CSWAP(1, 2)
#undef PYR_CSWAP
        if (Stack::FAST_DEPTH > 0 && sp + 4 <= Stack::FAST_DEPTH) {
            // branch-free: every entry is stored, the stack pointer only moves past the real ones (misses sorted to the end, so what
            // a skipped store leaves behind lies above the top); the last store puts the entry distance of `cur`'s box where
            // leaf_step expects it
            stack.put_fast(sp, code[3], dist[3]); sp += code[3] != NODE4_EMPTY ? 1 : 0;
            stack.put_fast(sp, code[2], dist[2]); sp += code[2] != NODE4_EMPTY ? 1 : 0;
            stack.put_fast(sp, code[1], dist[1]); sp += code[1] != NODE4_EMPTY ? 1 : 0;
            stack.put_fast(sp, code[0], dist[0]);
            if (code[0] != NODE4_EMPTY) cur = code[0]; else pop(stack);
            return;
        }
        if (code[3] != NODE4_EMPTY) stack.put(sp++, code[3], dist[3]);
        if (code[2] != NODE4_EMPTY) stack.put(sp++, code[2], dist[2]);
        if (code[1] != NODE4_EMPTY) stack.put(sp++, code[1], dist[1]);
        if (code[0] != NODE4_EMPTY) {
            cur = code[0];
            if (cur < 0) stack.put(sp, cur, dist[0]);  // a leaf reached without a pop: leave its box's entry distance where a pop would
        } else pop(stack);
    }
    // one leaf (cur < 0): the primitive test of Shape::ray_intersect and World::intersect's acceptance rule
    template <class Stack>
    PYR_HD void leaf_step(const SceneView& sc, Stack& stack) {
        const uint32_t r = (uint32_t)~cur;
        const Prim pr = fetch_prim(sc.prims + r);
        const uint32_t k = prim_kind(pr);
        if (STATS) ++vl;
        float ht = 0, hu = 0, hv = 0;
        bool ok;
        if (k == KIND_TRIANGLE) ok = triangle_test(prim_v1(pr), prim_e1(pr), prim_e2(pr), o, d, ht, hu, hv);
        else if (k == KIND_SPHERE) { v3 p; ok = sphere_test(prim_v1(pr), pr.a.w, o, d, ht, p); }
        else {
            // Sphere tracing is hundreds of distance-estimator evaluations: it is not run here, one lane at a time,
            // but after the walk (resolve_marched / k_march), where every lane marches.  The result is the same
            // lexicographic minimum (t, rank); only the culling by `closest` is less tight meanwhile.
            march_mask |= 1u << (f_bits(pr.a.x) & 31u);
            ok = false;
        }
        if (ok && ht > DIST_EPSILON) {
            if (mode != 0) {
                if (occludes(mode, ht, limit)) { t = ht; u = hu; v = hv; rank = r; kind = k; done = true; return; }
            } else {
                // The reference folds the leaves in pre-order (= rank order) and takes a leaf only if BOTH its own box's entry
                // distance and its hit distance lie below the closest hit so far (bvh.rs:213 `d >= closest` skips it otherwise,
                // world.rs:288-296 strict `<`).  With m = max(hit distance, box entry distance) that is "m < closest": a later leaf
                // beats an earlier hit only with m below its distance - also when rounding puts a hit a few ulps in FRONT of its own
                // box (a triangle lying in a face of its box, e.g. coplanar with a plane of the scene).  This walk meets the
                // candidates in another order, so it applies the same rule pairwise, ordered by rank (planes come before every leaf).
                const float m_new = fmaxf(ht, stack.dist(sp));
                bool take;
                if (kind == KIND_MISS) take = true;
                else if (kind == KIND_PLANE || rank < r) take = m_new < closest;   // the current hit comes first in the reference's order
                else take = !(stack.best_m() < ht);                               // the new one comes first: the current one survives only if it would have been taken after it
                if (take) { closest = ht; cull = ht * CULL_SLACK; stack.set_best_m(m_new); t = ht; u = hu; v = hv; rank = r; kind = k; }
            }
        }
        pop(stack);
    }
    template <class Stack>
    PYR_HD void step(const SceneView& sc, Stack& stack) {
        if (at_node()) node_step(sc, stack); else leaf_step(sc, stack);
    }

    // Shape::RayMarched leaves reached by the walk (shapes/mod.rs:120-154), merged with World::intersect's rule
    PYR_HD void resolve_marched(const SceneView& sc) {
        uint32_t m = march_mask;
        march_mask = 0;
        while (m) {
#if defined(__CUDA_ARCH__)
            const int i = __ffs((int)m) - 1;
#else
            const int i = __builtin_ffs((int)m) - 1;
#endif
            m &= m - 1u;
            const MarchedRec& mr = sc.marched[i];
            float ht = 0.0f;
            // (a march that runs off to +inf "hits" there: World::intersect's `distance < closest` drops it, closest starting at +inf)
            if (march_test(mr, o, d, ht, de_evals, de_iters) && ht > DIST_EPSILON && ht < PYR_INF) {
                if (mode != 0) {
                    if (occludes(mode, ht, limit)) { t = ht; u = 0; v = 0; rank = mr.rank; kind = KIND_RAY_MARCHED; return; }
                } else if (ht < closest || (ht == closest && kind != KIND_PLANE && mr.rank < rank)) {
                    closest = ht; cull = ht * CULL_SLACK; t = ht; u = 0; v = 0; rank = mr.rank; kind = KIND_RAY_MARCHED;
                }
            }
        }
    }
    // resume from a stored result (k_march): the ray and what the walk found so far
    PYR_HD void resume(const Ray& ray, float t_, float u_, float v_, uint32_t rank_, uint32_t kind_, uint32_t mask) {
        o = ld3(ray.o); d = ld3(ray.d);
        mode = ray.mode; limit = ray.limit;
        t = t_; u = u_; v = v_; rank = rank_; kind = kind_;
        closest = kind_ == KIND_MISS ? PYR_INF : t_;
        cull = closest * CULL_SLACK;
        bound = PYR_INF;
        vn = 0; vl = 0; vf = 0; de_evals = 0; de_iters = 0;
        sp = 0; cur = 0; done = true;
        march_mask = mask;
    }

    PYR_HD void finish(Hit& hit, TraceStats* stats) const {
        hit.t = t; hit.u = u; hit.v = v; hit.rank = rank; hit.kind = kind; hit.nodes = vn; hit.leaves = vl; hit.pad = 0;
        if (STATS && stats) { stats->nodes += vn; stats->leaves += vl; stats->de_evals += de_evals; stats->de_iters += de_iters; stats->fetches += vf; }
    }
};

template <bool STATS, class Stack>
PYR_HD void trace_ray(const SceneView& sc, const Ray& ray, Hit& hit, TraceStats* stats, Stack& stack) {
    Traversal<STATS> tr;
    tr.begin(sc, ray, stack);
    while (!tr.done) tr.step(sc, stack);
    if (tr.march_mask && !(tr.mode != 0 && tr.kind != KIND_MISS)) tr.resolve_marched(sc);
    tr.finish(hit, stats);
}

template <bool STATS>
PYR_HD void trace_ray(const SceneView& sc, const Ray& ray, Hit& hit, TraceStats* stats) {
    LocalStack stack;
    trace_ray<STATS>(sc, ray, hit, stats, stack);
}

}  // namespace pyr
