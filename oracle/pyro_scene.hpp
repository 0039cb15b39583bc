// ORACLE - TEST INFRASTRUCTURE ONLY (see pyro_math.hpp header).  PARITY UNPINNED.
//
// Scene build + geometry of the reference, restated on the CPU:
//   expression evaluation   program/compiler.rs, program/execution_context.rs, project/expressions.rs
//   spectra / textures      project/spectra.rs, math.rs:17-73, texture.rs:88-148,303-334
//   materials               materials/mod.rs, materials/{diffuse,mirror,refractive}.rs
//   shapes                  shapes/mod.rs, shapes/distance_estimators.rs
//   BVH                     spatial/bvh.rs
//   world, lamps, camera    world.rs, lamp.rs, cameras.rs
#pragma once
#include <cstdlib>
#include <algorithm>
#include <array>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "pyro_ir.hpp"

namespace pyro {

// ---------------------------------------------------------------- constant evaluation
// project/expressions.rs:203-258 (`Evaluate` for ComplexExpression) with T = f32 / Vector.
struct ConstEval {
    const Project& P;
    float number(const Expr& e) const {
        if (e.tag == 0) return (float)e.number;  // expressions.rs:279-281
        const ExprNode& n = P.exprs[e.id];
        switch (n.type) {
            case E_VECTOR: throw std::runtime_error("expected a number, but found a vector");
            case E_RGB: throw std::runtime_error("expected a number, but found an RGB color");
            case E_BINARY: {
                float l = number(n.a), r = number(n.b);
                switch (n.op) { case OP_ADD: return l + r; case OP_SUB: return l - r; case OP_MUL: return l * r; default: return l / r; }
            }
            case E_MIX: {
                float amount = number(n.a), l = number(n.b), r = number(n.c);
                amount = fmax_(fmin_(amount, 1.0f), 0.0f);  // expressions.rs:291-294
                return l * (1.0f - amount) + r * amount;
            }
            case E_CLAMP: {
                float v = number(n.a), lo = number(n.b), hi = number(n.c);
                return fmax_(fmin_(v, hi), lo);
            }
            case E_FRESNEL: throw std::runtime_error("cannot evaluate Fresnel functions as constants");
            case E_BLACKBODY: throw std::runtime_error("cannot evaluate black-body functions as constants");
            case E_SPECTRUM: throw std::runtime_error("cannot evaluate spectra as constants");
            default: throw std::runtime_error("cannot evaluate textures as constants");
        }
    }
    Vec4 vector(const Expr& e) const {
        if (e.tag == 0) { float v = (float)e.number; return {v, v, v, v}; }  // expressions.rs:328-335
        const ExprNode& n = P.exprs[e.id];
        switch (n.type) {
            case E_VECTOR: return {number(n.a), number(n.b), number(n.c), number(n.d)};
            case E_RGB: number(n.a); number(n.b); number(n.c); throw std::runtime_error("expected a vector, but found an RGB color");
            case E_BINARY: {
                Vec4 l = vector(n.a), r = vector(n.b);
                switch (n.op) { case OP_ADD: return l + r; case OP_SUB: return l - r; case OP_MUL: return l * r; default: return l / r; }
            }
            case E_MIX: {
                float amount = number(n.a);
                Vec4 l = vector(n.b), r = vector(n.c);
                amount = fmax_(fmin_(amount, 1.0f), 0.0f);
                return l + (r - l) * amount;  // Vector4::lerp (expressions.rs:345-347)
            }
            case E_CLAMP: throw std::runtime_error("vectors cannot be clamped");
            case E_FRESNEL: throw std::runtime_error("cannot evaluate Fresnel functions as constants");
            case E_BLACKBODY: throw std::runtime_error("cannot evaluate black-body functions as constants");
            case E_SPECTRUM: throw std::runtime_error("cannot evaluate spectra as constants");
            default: throw std::runtime_error("cannot evaluate textures as constants");
        }
    }
    Vec3 vec3(const Expr& e) const { Vec4 v = vector(e); return {v.x, v.y, v.z}; }
    uint16_t u16(const Expr& e) const {
        if (e.tag == 0) return f64_as_u16(e.number);  // expressions.rs:301-304
        return (uint16_t)number(e);                   // complex u16 expressions do not occur in practice
    }
};

// ---------------------------------------------------------------- spectra and textures
// math.rs:21-72 `Interpolated::get` (Curve): binary search + lerp, 0 outside and AT the end points.
inline float curve_get(const std::vector<std::pair<float, float>>& pts, float input) {
    if (pts.empty()) return 0.0f;
    size_t min = 0, max = pts.size() - 1;
    if (pts[min].first >= input) return 0.0f;
    if (pts[max].first <= input) return 0.0f;
    while (max > min + 1) {
        size_t check = (max + min) / 2;
        if (pts[check].first == input) return pts[check].second;
        if (pts[check].first > input) max = check; else min = check;
    }
    float min_x = pts[min].first, min_y = pts[min].second, max_x = pts[max].first, max_y = pts[max].second;
    if (input < min_x) return 0.0f;
    if (input > max_x) return 0.0f;
    return min_y + (max_y - min_y) * ((input - min_x) / (max_x - min_x));
}
// project/spectra.rs:30-58 `Spectrum::Array::get`: clamp to the end values, lerp between samples.
inline float array_get(const float* pts, size_t n, size_t stride, float min, float max, float w) {
    if (n == 0) return 0.0f;
    if (w <= min) return pts[0];
    if (w >= max) return pts[(n - 1) * stride];
    float normalized = (w - min) / (max - min);
    float float_index = normalized * ((float)n - 1.0f);
    float min_float_index = truncf(float_index);
    size_t min_index = f32_as_usize(min_float_index);
    size_t max_index = min_index + 1;
    float min_value = pts[min_index * stride], max_value = pts[max_index * stride];
    float mix = float_index - min_float_index;
    return min_value * (1.0f - mix) + max_value * mix;
}
inline float spectrum_get(const SpectrumData& s, float w) {
    return s.is_curve ? curve_get(s.curve, w) : array_get(s.points.data(), s.points.size(), 1, s.min, s.max, w);
}

// texture.rs:324-334
inline float cubic_interpolate(float v1, float v2, float v3, float v4, float pos) {
    float a = (v4 - v3) - (v1 - v2);
    float b = (v1 - v2) - a;
    float c = v3 - v1;
    float d = v2;
    return d + (c + (b + a * pos) * pos) * pos;
}
// texture.rs:88-148 `Texture::get_color`: wrap-around 4x4 bicubic; channels = 4 (LinSrgba) or 1 (LinLuma).
inline void texture_get(const TextureData& t, int channels, float px, float py, float* out) {
    int64_t width = t.width, height = t.height;
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
            for (int k = 0; k < 4; ++k) v[k] = t.data[(size_t)(xs[k] + ys[r] * width) * channels + c];
            rows[r] = cubic_interpolate(v[0], v[1], v[2], v[3], fx);
        }
        out[c] = cubic_interpolate(rows[0], rows[1], rows[2], rows[3], fy);
    }
}

// ---------------------------------------------------------------- math.rs helpers
// math.rs:75-96
inline float schlick(float ref_index1, float ref_index2, Vec3 normal, Vec3 incident) {
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
// math.rs:167-175
inline float fresnel(float ior, float env_ior, Vec3 normal, Vec3 incident) {
    if (dot(incident, normal) < 0.0f) return schlick(env_ior, ior, normal, incident);
    return schlick(ior, env_ior, -normal, incident);
}
// math.rs:177-182; powi(-5) = 1 / (a * (a^2)^2) as compiler-rt's __powisf2 expands it
inline float blackbody(float wavelength, float temperature) {
    wavelength = wavelength * 1.0e-9f;
    float a2 = wavelength * wavelength;
    float a4 = a2 * a2;
    float p5 = 1.0f / (wavelength * a4);
    float power_term = 3.74183e-16f * p5;
    return power_term / (m_exp(1.4388e-2f / (wavelength * temperature)) - 1.0f);
}
// math.rs:98-113
inline Vec3 ortho(Vec3 v) {
    Vec3 unit;
    if (fabsf(v.x) < DIST_EPSILON) unit = {1, 0, 0};
    else if (fabsf(v.y) < DIST_EPSILON) unit = {0, 1, 0};
    else if (fabsf(v.z) < DIST_EPSILON) unit = {0, 0, 1};
    else unit = {-v.y, v.x, 0.0f};
    return cross(v, unit);
}
// math.rs:119-123
inline void basis(Vec3 x, Vec3& y, Vec3& z) {
    z = normalize(ortho(x));
    y = normalize(cross(z, x));
}
// math.rs:125-137
inline Vec3 sample_cone(XorShift& rng, Vec3 direction, float cos_half) {
    Vec3 o1 = normalize(ortho(direction));
    Vec3 o2 = normalize(cross(direction, o1));
    float r1 = PI * 2.0f * rng.gen_f32();
    float r2 = cos_half + (1.0f - cos_half) * rng.gen_f32();
    float oneminus = sqrtf(1.0f - r2 * r2);
    return (o1 * m_cos(r1) * oneminus + o2 * m_sin(r1) * oneminus) + direction * r2;
}
// math.rs:139-145
inline float solid_angle(float cos_half) { return cos_half >= 1.0f ? 0.0f : 2.0f * PI * (1.0f - cos_half); }
// math.rs:147-153
inline Vec3 sample_sphere(XorShift& rng) {
    float u = rng.gen_f32();
    float v = rng.gen_f32();
    float theta = 2.0f * PI * u;
    float phi = m_acos(2.0f * v - 1.0f);
    return {m_sin(phi) * m_cos(theta), m_sin(phi) * m_sin(theta), m_cos(phi)};
}
// math.rs:155-164
inline Vec3 sample_hemisphere(XorShift& rng, Vec3 direction) {
    Vec3 s = sample_sphere(rng);
    Vec3 x = normalize_to(ortho(direction), s.x);
    Vec3 y = normalize_to(cross(x, direction), s.y);
    Vec3 z = normalize_to(direction, fabsf(s.z));
    return (x + y) + z;
}
// math.rs:184-207; returns false for None
inline bool aabb_intersection_distance(const Aabb& aabb, const Ray& ray, float& dist) {
    Vec3 inv{1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z};
    float t1 = (aabb.min.x - ray.origin.x) * inv.x;
    float t2 = (aabb.max.x - ray.origin.x) * inv.x;
    float tmin = fmin_(t1, t2), tmax = fmax_(t1, t2);
    t1 = (aabb.min.y - ray.origin.y) * inv.y;
    t2 = (aabb.max.y - ray.origin.y) * inv.y;
    tmin = fmax_(tmin, fmin_(t1, t2));
    tmax = fmin_(tmax, fmax_(t1, t2));
    t1 = (aabb.min.z - ray.origin.z) * inv.z;
    t2 = (aabb.max.z - ray.origin.z) * inv.z;
    tmin = fmax_(tmin, fmin_(t1, t2));
    tmax = fmin_(tmax, fmax_(t1, t2));
    if (tmax >= tmin && tmax >= 0.0f) { dist = fmax_(tmin, 0.0f); return true; }
    return false;
}

// ---------------------------------------------------------------- programs
// The reference compiles an expression graph into straight-line register bytecode
// (program/compiler.rs:48-586) and interprets it (program/execution_context.rs:69-283).  All
// instructions are pure, so a recursive evaluation of the graph with the compiler's coercion
// rules yields the same f32 results; the memoised re-run (execution_context.rs:310-342) only
// skips instructions whose inputs did not change and therefore also returns the same values.
struct ProgramInputs {
    float wavelength = 0;
    Vec3 normal, incident;
    Vec2 texture;
};
enum InputMask : uint32_t { IN_WAVELENGTH = 1, IN_NORMAL = 16, IN_INCIDENT = 32, IN_TEXTURE = 64 };  // program/mod.rs:150-158
enum ValueKind { V_NUMBER, V_VECTOR, V_RGB };
struct Value {
    ValueKind kind = V_NUMBER;
    Vec4 v;  // number in x; rgb = (r,g,b,alpha)
};

struct Resources {
    const Project* P = nullptr;
};

struct Program {
    bool present = false;
    bool constant = true;
    float value = 0.0f;
    uint32_t root = 0;
    bool vector_output = false;
    uint32_t reads = 0;  // union of InputMask bits of every instruction that would be emitted
};

struct ProgramEval {
    const Project& P;
    const ProgramInputs& in;

    // RgbSpectrumValue (execution_context.rs:140-152): rgb.color * RGB.get(wavelength), summed r+g+b
    float rgb_spectrum(Vec4 rgb) const {
        size_t n = P.burns.size() / 3;
        float r = array_get(P.burns.data() + 0, n, 3, P.burns_min, P.burns_max, in.wavelength);
        float g = array_get(P.burns.data() + 1, n, 3, P.burns_min, P.burns_max, in.wavelength);
        float b = array_get(P.burns.data() + 2, n, 3, P.burns_min, P.burns_max, in.wavelength);
        float rr = rgb.x * r, gg = rgb.y * g, bb = rgb.z * b;
        return (rr + gg) + bb;
    }
    // try_get_number_value (compiler.rs:638-680)
    float number(const Expr& e) const {
        if (e.tag == 0) return (float)e.number;
        Value v = eval(e.id);
        if (v.kind == V_NUMBER) return v.v.x;
        if (v.kind == V_RGB) return rgb_spectrum(v.v);
        throw std::runtime_error("cannot use a vector as a number");
    }
    static Vec4 rgb_to_vector(Vec4 c) {  // execution_context.rs:183-193
        return {(c.x * 2.0f) - 1.0f, (c.y * 2.0f) - 1.0f, (c.z * 2.0f) - 1.0f, (c.w * 2.0f) - 1.0f};
    }
    // convert_operands (compiler.rs:682-968)
    static void promote(Value& l, Value& r) {
        auto to_rgb = [](Value& v) { float n = v.v.x; v.kind = V_RGB; v.v = {n, n, n, 1.0f}; };
        auto to_vec = [](Value& v) {
            if (v.kind == V_NUMBER) { float n = v.v.x; v.v = {n, n, n, n}; } else v.v = rgb_to_vector(v.v);
            v.kind = V_VECTOR;
        };
        if (l.kind == r.kind) return;
        if (l.kind == V_VECTOR) { to_vec(r); return; }
        if (r.kind == V_VECTOR) { to_vec(l); return; }
        if (l.kind == V_NUMBER) to_rgb(l); else to_rgb(r);
    }
    Value operand(const Expr& e) const {  // try_get_register (compiler.rs:612-636)
        if (e.tag == 0) { Value v; v.kind = V_NUMBER; v.v.x = (float)e.number; return v; }
        return eval(e.id);
    }
    Value eval(uint32_t id) const {
        const ExprNode& n = P.exprs[id];
        Value out;
        switch (n.type) {
            case E_VECTOR: {
                float x = number(n.a), y = number(n.b), z = number(n.c), w = number(n.d);
                out.kind = V_VECTOR; out.v = {x, y, z, w};
                return out;
            }
            case E_RGB: {
                float r = number(n.a), g = number(n.b), b = number(n.c);
                out.kind = V_RGB; out.v = {r, g, b, 1.0f};
                return out;
            }
            case E_FRESNEL: {
                float ior = number(n.a), env = number(n.b);
                out.v.x = fresnel(ior, env, in.normal, in.incident);
                return out;
            }
            case E_BLACKBODY: {
                float t = number(n.a);
                out.v.x = blackbody(in.wavelength, t);
                return out;
            }
            case E_SPECTRUM: out.v.x = spectrum_get(P.spectra[n.resource], in.wavelength); return out;
            case E_COLOR_TEXTURE: {
                float c[4];
                texture_get(P.color_textures[n.resource], 4, in.texture.x, in.texture.y, c);
                out.kind = V_RGB; out.v = {c[0], c[1], c[2], c[3]};
                return out;
            }
            case E_MONO_TEXTURE: {
                float c[1];
                texture_get(P.mono_textures[n.resource], 1, in.texture.x, in.texture.y, c);
                out.v.x = c[0];
                return out;
            }
            case E_MIX: {  // execution_context.rs:195-227
                float amount = number(n.a);
                Value l = operand(n.b), r = operand(n.c);
                promote(l, r);
                amount = fmax_(fmin_(amount, 1.0f), 0.0f);
                out.kind = l.kind;
                if (l.kind == V_NUMBER) out.v.x = l.v.x * (1.0f - amount) + r.v.x * amount;
                else out.v = l.v + (r.v - l.v) * amount;  // Vector4::lerp / palette Mix (incl. alpha)
                return out;
            }
            case E_BINARY: {  // execution_context.rs:228-268
                Value l = operand(n.a), r = operand(n.b);
                promote(l, r);
                out.kind = l.kind;
                switch (n.op) {
                    case OP_ADD: out.v = l.v + r.v; break;
                    case OP_SUB: out.v = l.v - r.v; break;
                    case OP_MUL: out.v = l.v * r.v; break;
                    default: out.v = l.v / r.v; break;
                }
                if (l.kind == V_NUMBER) { out.v.y = out.v.z = out.v.w = 0; }
                return out;
            }
            case E_CLAMP: {
                float v = number(n.a), lo = number(n.b), hi = number(n.c);
                out.v.x = fmax_(fmin_(v, hi), lo);
                return out;
            }
        }
        throw std::runtime_error("bad expression node");
    }
};

// Static analysis the compiler performs: which inputs instructions read, and type errors.
struct ProgramCompiler {
    Project& P;
    // returns the value kind of node `id`; accumulates read inputs
    ValueKind kind_of(uint32_t id, uint32_t& reads, uint32_t allowed, int depth = 0) const {
        if (depth > 4096) throw std::runtime_error("expression graph too deep (cycle?)");
        const ExprNode& n = P.exprs[id];
        auto num = [&](const Expr& e) {
            if (e.tag == 0) return;
            ValueKind k = kind_of(e.id, reads, allowed, depth + 1);
            if (k == V_VECTOR) throw std::runtime_error("cannot use a vector as a number");
            if (k == V_RGB) need(IN_WAVELENGTH, reads, allowed);
        };
        auto opnd = [&](const Expr& e) { return e.tag == 0 ? V_NUMBER : kind_of(e.id, reads, allowed, depth + 1); };
        auto promote = [](ValueKind l, ValueKind r) {
            if (l == r) return l;
            if (l == V_VECTOR || r == V_VECTOR) return V_VECTOR;
            return V_RGB;
        };
        switch (n.type) {
            case E_VECTOR: num(n.a); num(n.b); num(n.c); num(n.d); return V_VECTOR;
            case E_RGB: num(n.a); num(n.b); num(n.c); return V_RGB;
            case E_FRESNEL: need(IN_NORMAL, reads, allowed); need(IN_INCIDENT, reads, allowed); num(n.a); num(n.b); return V_NUMBER;
            case E_BLACKBODY: need(IN_WAVELENGTH, reads, allowed); num(n.a); return V_NUMBER;
            case E_SPECTRUM: need(IN_WAVELENGTH, reads, allowed); return V_NUMBER;
            case E_COLOR_TEXTURE: need(IN_TEXTURE, reads, allowed); return V_RGB;
            case E_MONO_TEXTURE: need(IN_TEXTURE, reads, allowed); return V_NUMBER;
            case E_MIX: { num(n.a); ValueKind l = opnd(n.b), r = opnd(n.c); return promote(l, r); }
            case E_BINARY: { ValueKind l = opnd(n.a), r = opnd(n.b); return promote(l, r); }
            case E_CLAMP: num(n.a); num(n.b); num(n.c); return V_NUMBER;
        }
        throw std::runtime_error("bad expression node");
    }
    static void need(uint32_t bit, uint32_t& reads, uint32_t allowed) {
        if (!(allowed & bit)) {
            // tracer.rs:60-70, main.rs:502-517: the input is not available for this program type
            switch (bit) {
                case IN_WAVELENGTH: throw std::runtime_error("the wavelength is not available during normal mapping");
                case IN_NORMAL: throw std::runtime_error("the surface normal cannot be used while sampling a constant spectrum");
                case IN_INCIDENT: throw std::runtime_error("the incident vector cannot be used while sampling a constant spectrum");
                default: throw std::runtime_error("texture coordinates cannot be used while sampling a constant spectrum");
            }
        }
        reads |= bit;
    }
    // ProgramCompiler::compile (compiler.rs:48-586)
    Program compile(const Expr& e, bool vector_output, uint32_t allowed) const {
        Program p;
        p.present = true;
        p.vector_output = vector_output;
        if (e.tag == 0) { p.constant = true; p.value = (float)e.number; return p; }
        p.constant = false;
        p.root = e.id;
        ValueKind k = kind_of(e.id, p.reads, allowed);
        if (!vector_output) {
            if (k == V_VECTOR) throw std::runtime_error("cannot use a vector as a number");
            if (k == V_RGB) need(IN_WAVELENGTH, p.reads, allowed);
        }
        return p;
    }
};

// ExecutionContext::run for T = f32 (execution_context.rs:29-56)
inline float run_number(const Project& P, const Program& p, const ProgramInputs& in) {
    if (p.constant) return p.value;
    ProgramEval ev{P, in};
    Value v = ev.eval(p.root);
    if (v.kind == V_NUMBER) return v.v.x;
    if (v.kind == V_RGB) return ev.rgb_spectrum(v.v);
    throw std::runtime_error("cannot use a vector as a number");
}
// ... and for T = Vector (compiler.rs:532-567 output conversions)
inline Vec4 run_vector(const Project& P, const Program& p, const ProgramInputs& in) {
    if (p.constant) return {p.value, p.value, p.value, p.value};
    ProgramEval ev{P, in};
    Value v = ev.eval(p.root);
    if (v.kind == V_VECTOR) return v.v;
    if (v.kind == V_NUMBER) return {v.v.x, v.v.x, v.v.x, v.v.x};
    return ProgramEval::rgb_to_vector(v.v);
}

// ---------------------------------------------------------------- materials (materials/mod.rs)
enum BsdfType { B_EMISSIVE, B_DIFFUSE, B_MIRROR, B_REFRACTIVE };
struct RefractiveProps { float ior = 1, env_ior = 1, dispersion = 0, env_dispersion = 0; };
struct Component {  // materials/mod.rs:230-235
    float selection_compensation = 1.0f;
    Program probability;  // present == Some
    Program color;
    BsdfType bsdf = B_DIFFUSE;
    RefractiveProps props;
};
struct Material {  // materials/mod.rs:26-30, 83-87
    std::vector<Component> components, emissive;
    Program normal_map;  // present == Some
    bool is_emissive() const { return !emissive.empty(); }
};

// expressions.rs:20-63 (constant folding in f64, then new nodes)
inline Expr insert_sub(Project& P, Expr l, Expr r) {
    if (l.tag == 0 && r.tag == 0) return Expr::num(l.number - r.number);
    ExprNode n; n.type = E_BINARY; n.op = OP_SUB; n.a = l; n.b = r;
    P.exprs.push_back(n);
    return Expr::node((uint32_t)P.exprs.size() - 1);
}
inline Expr insert_mul(Project& P, Expr l, Expr r) {
    if (l.tag == 0 && r.tag == 0) return Expr::num(l.number * r.number);
    ExprNode n; n.type = E_BINARY; n.op = OP_MUL; n.a = l; n.b = r;
    P.exprs.push_back(n);
    return Expr::node((uint32_t)P.exprs.size() - 1);
}
inline Expr insert_clamp(Project& P, Expr v, Expr lo, Expr hi) {
    if (v.tag == 0 && lo.tag == 0 && hi.tag == 0) return Expr::num(std::fmax(std::fmin(v.number, hi.number), lo.number));
    ExprNode n; n.type = E_CLAMP; n.a = v; n.b = lo; n.c = hi;
    P.exprs.push_back(n);
    return Expr::node((uint32_t)P.exprs.size() - 1);
}

constexpr uint32_t ALLOW_RENDER = IN_WAVELENGTH | IN_NORMAL | IN_INCIDENT | IN_TEXTURE;  // RenderContext, ProbabilityInput
constexpr uint32_t ALLOW_NORMAL_MAP = IN_NORMAL | IN_INCIDENT | IN_TEXTURE;              // NormalInput (tracer.rs:57-70)
constexpr uint32_t ALLOW_SPECTRUM = IN_WAVELENGTH;                                       // SpectrumSamplingInput (main.rs:463-518)

// Material::from_project + SurfaceMaterial::from_project (materials/mod.rs:33-46, 89-228)
inline Material material_from_project(Project& P, const MaterialRef& ref) {
    ProgramCompiler pc{P};
    ConstEval ce{P};
    struct Entry { uint32_t material; bool has_prob; Expr prob; };
    std::vector<Entry> stack{{ref.surface, false, Expr()}};
    Material m;
    auto prob_program = [&](const Entry& e) {
        Program p;
        if (e.has_prob) p = pc.compile(e.prob, false, ALLOW_RENDER);
        return p;
    };
    while (!stack.empty()) {
        Entry entry = stack.back();
        stack.pop_back();
        const MatNode node = P.mats.at(entry.material);
        switch (node.type) {
            case M_EMISSIVE: case M_DIFFUSE: case M_MIRROR: {
                Component c;
                c.probability = prob_program(entry);
                c.color = pc.compile(node.color, false, ALLOW_RENDER);
                c.bsdf = node.type == M_EMISSIVE ? B_EMISSIVE : node.type == M_DIFFUSE ? B_DIFFUSE : B_MIRROR;
                m.components.push_back(c);
                if (node.type == M_EMISSIVE) m.emissive.push_back(c);
                break;
            }
            case M_REFRACTIVE: {
                Component c;
                c.probability = prob_program(entry);
                c.color = pc.compile(node.color, false, ALLOW_RENDER);
                c.bsdf = B_REFRACTIVE;
                c.props.ior = ce.number(node.ior);
                c.props.env_ior = node.env_ior.present ? ce.number(node.env_ior.e) : 1.0f;
                c.props.dispersion = node.dispersion.present ? ce.number(node.dispersion.e) : 0.0f;
                c.props.env_dispersion = node.env_dispersion.present ? ce.number(node.env_dispersion.e) : 0.0f;
                m.components.push_back(c);
                break;
            }
            case M_MIX: {  // materials/mod.rs:176-195; `amount` belongs to lhs
                Expr amount = insert_clamp(P, node.amount, Expr::num(0.0), Expr::num(1.0));
                Expr lhs_probability = entry.has_prob ? insert_mul(P, entry.prob, amount) : amount;
                stack.push_back({node.lhs, true, lhs_probability});
                stack.push_back({node.rhs, true, insert_sub(P, Expr::num(1.0), lhs_probability)});
                break;
            }
            case M_ADD:
                stack.push_back({node.lhs, entry.has_prob, entry.prob});
                stack.push_back({node.rhs, entry.has_prob, entry.prob});
                break;
            default: throw std::runtime_error("bad material node");
        }
    }
    for (auto& c : m.components) c.selection_compensation = (float)m.components.size();
    for (auto& c : m.emissive) c.selection_compensation = (float)m.emissive.size();
    if (ref.normal_map.present) m.normal_map = pc.compile(ref.normal_map.e, true, ALLOW_NORMAL_MAP);
    return m;
}

// MaterialComponent::get_probability (materials/mod.rs:237-249) + ProbabilityInput::wavelength_used
// (materials/mod.rs:263-270): the flag is set iff an executed instruction reads the wavelength,
// and `run` executes every instruction, so it equals the static "program reads wavelength" bit.
inline float component_probability(const Project& P, const Component& c, const ProgramInputs& in, bool& wavelength_used) {
    wavelength_used = false;
    if (c.probability.present) {
        wavelength_used = !c.probability.constant && (c.probability.reads & IN_WAVELENGTH);
        return run_number(P, c.probability, in) * c.selection_compensation;
    }
    return c.selection_compensation;
}

// ---------------------------------------------------------------- Normal (shapes/mod.rs:531-584)
struct Normal {
    Vec3 vector;
    Quat from_space;
    static Normal from_vector(Vec3 v) {
        Vec3 x, y;
        basis(v, x, y);
        return {v, quat_from_mat3(Mat3::from_cols(x, y, v))};
    }
    static Normal on_triangle(const Normal& n1, const Normal& n2, const Normal& n3, float u, float v) {
        float w = 1.0f - (u + v);
        Vec3 vec = (n1.vector * w + n2.vector * u) + n3.vector * v;
        Quat q = (n1.from_space * w + n2.from_space * u) + n3.from_space * v;
        return {normalize(vec), normalize(q)};
    }
    Vec3 from_space_v(Vec3 v) const { return rotate(from_space, v); }
    Vec3 into_space(Vec3 v) const { return rotate(conjugate(from_space), v); }
    Normal transform(const Mat4& t) const {
        Vec3 vec = normalize(transform_vector(t, vector));
        Vec3 x = normalize(transform_vector(t, from_space_v({1, 0, 0})));
        Vec3 y = normalize(transform_vector(t, from_space_v({0, 1, 0})));
        return {vec, quat_from_mat3(Mat3::from_cols(x, y, vec))};
    }
};

// ---------------------------------------------------------------- distance estimators
// per-thread work counters for the FP32 roofline of the sphere-tracing stage
inline thread_local uint64_t tl_de_evals = 0, tl_de_iters = 0;

struct Estimator {  // shapes/distance_estimators.rs
    uint32_t type = 0;  // 0 mandelbulb, 1 quaternion julia
    uint16_t iterations = 0;
    float threshold = 0, power = 0, slice_plane = 0;
    bool has_constant = false;
    Vec3 mb_constant;
    Quat constant;
    uint32_t variant = 0;

    static Quat bicomplex_mul(Quat a, Quat b) {  // distance_estimators.rs:96-107
        float x1 = a.s, x2 = b.s, y1 = a.x, y2 = b.x, z1 = a.y, z2 = b.y, w1 = a.z, w2 = b.z;
        float x = x1 * x2 - y1 * y2 - z1 * z2 + w1 * w2;
        float y = x1 * y2 + y1 * x2 - z1 * w2 - w1 * z2;
        float z = x1 * z2 - y1 * w2 + z1 * x2 - w1 * y2;
        float w = x1 * w2 + y1 * z2 + z1 * y2 + w1 * x2;
        return {x, y, z, w};
    }
    Quat qpow(Quat z) const {  // :78-84
        switch (variant) { case 0: return z * z; case 1: return (z * z) * z; default: return bicomplex_mul(z, z); }
    }
    Quat qpow_prim(Quat z, Quat dz) const {  // :86-92
        switch (variant) {
            case 0: return (dz * z) * 2.0f;
            case 1: return ((dz * z) * z) * 3.0f;
            default: return bicomplex_mul(bicomplex_mul(dz, z), z) * 2.0f;
        }
    }
    float get(Vec3 point, uint64_t* iter_count = nullptr) const {
        uint64_t it = 0;
        float result;
        if (type == 0) {  // Mandelbulb::get :12-42
            Vec3 z = point;
            float r = 0.0f, dr = 1.0f;
            float dc = has_constant ? 0.0f : 1.0f;
            for (uint16_t i = 0; i < iterations; ++i) {
                r = magnitude(z);
                if (r > threshold) break;
                ++it;
                float theta = acosf(z.z / r);
                float phi = atan2f(z.y, z.x);
                dr = powf(r, power - 1.0f) * power * dr + dc;
                float zr = powf(r, power);
                theta *= power;
                phi *= power;
                z = Vec3(zr * sinf(theta) * cosf(phi), zr * sinf(phi) * sinf(theta), zr * cosf(theta));
                z = z + (has_constant ? mb_constant : point);
            }
            result = 0.5f * logf(r) * r / dr;
        } else {  // QuaternionJulia::get :52-70
            Quat z(point.x, point.y, point.z, slice_plane);
            float r = 0.0f;
            Quat dz(1.0f, 0.0f, 0.0f, 0.0f);
            for (uint16_t i = 0; i < iterations; ++i) {
                r = magnitude(z);
                if (r > threshold) break;
                ++it;
                dz = qpow_prim(z, dz);
                z = qpow(z) + constant;
            }
            result = 0.5f * logf(r) * r / magnitude(dz);
        }
        if (iter_count) *iter_count += it;
        return result;
    }
};

// BoundingVolume (shapes/mod.rs:586-702); Box widens t_max
struct BoundingVolume {
    uint32_t type = 0;  // 0 box, 1 sphere
    Vec3 a, b;          // box min/max, or sphere centre in `a`
    float radius = 0;
    bool intersect(const Ray& ray, float& t_min_out, float& t_max_out) const {
        if (type == 0) {
            Vec3 min = a, max = b;
            Vec3 inv{1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z};
            float t_min, t_max;
            if (inv.x < 0.0f) { t_min = (max.x - ray.origin.x) * inv.x; t_max = (min.x - ray.origin.x) * inv.x; }
            else { t_min = (min.x - ray.origin.x) * inv.x; t_max = (max.x - ray.origin.x) * inv.x; }
            float ty_min, ty_max;
            if (inv.y < 0.0f) { ty_min = (max.y - ray.origin.y) * inv.y; ty_max = (min.y - ray.origin.y) * inv.y; }
            else { ty_min = (min.y - ray.origin.y) * inv.y; ty_max = (max.y - ray.origin.y) * inv.y; }
            if (t_min > ty_max || ty_min > t_max) return false;
            if (ty_min > t_min) t_min = ty_min;
            if (ty_max > t_max) t_max = ty_max;
            float tz_min, tz_max;
            if (inv.z < 0.0f) { tz_min = (max.z - ray.origin.z) * inv.z; tz_max = (min.z - ray.origin.z) * inv.z; }
            else { tz_min = (min.z - ray.origin.z) * inv.z; tz_max = (max.z - ray.origin.z) * inv.z; }
            if (t_min > tz_max || tz_min > t_max) return false;
            if (tz_min > t_min) t_min = tz_min;
            if (tz_max > t_max) t_max = tz_max;
            t_min = fmax_(t_min, 0.0f);
            if (t_min < t_max) { t_min_out = t_min; t_max_out = t_max; return true; }
            return false;
        }
        Vec3 l = a - ray.origin;
        float tca = dot(l, ray.direction);
        if (tca < 0.0f) return false;
        float d2 = dot(l, l) - tca * tca;
        if (d2 > radius * radius) return false;
        float thc = sqrtf(radius * radius - d2);
        t_min_out = tca - thc;
        t_max_out = tca + thc;
        return true;
    }
    Vec3 center() const { return type == 0 ? (a + b) * 0.5f : a; }
    Aabb aabb() const {
        if (type == 0) return Aabb::from_points(a, b);
        return Aabb::from_points({a.x - radius, a.y - radius, a.z - radius}, {a.x + radius, a.y + radius, a.z + radius});
    }
};

// ---------------------------------------------------------------- shapes (shapes/mod.rs)
struct Vertex { Vec3 position; Normal normal; Vec2 texture; };
enum ShapeKind : uint32_t { K_MISS = 0, K_PLANE = 1, K_TRIANGLE = 2, K_SPHERE = 3, K_RAY_MARCHED = 4 };

struct SurfaceData { Normal normal; Vec2 texture; };

struct Shape;
struct PlaneShape;
struct SurfacePoint {  // shapes/mod.rs:478-524
    Vec3 position;
    ShapeKind kind = K_MISS;
    const Shape* shape = nullptr;
    const PlaneShape* plane = nullptr;
    float u = 0, v = 0;
    Vec3 offset_position;
};
struct Intersection { float distance = 0; SurfacePoint surface_point; };

struct Shape {
    ShapeKind kind = K_TRIANGLE;
    uint32_t id = 0;  // insertion index in World::from_project's `objects`
    uint32_t material = 0;
    // sphere
    Vec3 position; float radius = 0; Vec2 texture_scale{1, 1};
    // triangle
    Vertex v1, v2, v3; Vec3 edge1, edge2;
    // ray marched
    Estimator estimator; BoundingVolume bounds;

    // Shape::ray_intersect (shapes/mod.rs:55-156)
    bool ray_intersect(const Ray& ray, Intersection& out) const {
        switch (kind) {
            case K_SPHERE: {  // collision::Sphere::intersection, near root only
                Vec3 l = position - ray.origin;
                float tca = dot(l, ray.direction);
                if (tca < 0.0f) return false;
                float d2 = dot(l, l) - tca * tca;
                if (d2 > radius * radius) return false;
                float thc = sqrtf(radius * radius - d2);
                Vec3 p = ray.origin + ray.direction * (tca - thc);
                out.distance = magnitude(p - ray.origin);
                out.surface_point = SurfacePoint{p, K_SPHERE, this, nullptr, 0, 0, {}};
                return true;
            }
            case K_TRIANGLE: {  // Moeller-Trumbore, :75-119
                Vec3 p = cross(ray.direction, edge2);
                float det = dot(edge1, p);
                if (det > -DIST_EPSILON && det < DIST_EPSILON) return false;
                float inv_det = 1.0f / det;
                Vec3 t = ray.origin - v1.position;
                float u = dot(t, p) * inv_det;
                if (u < 0.0f || u > 1.0f) return false;
                Vec3 q = cross(t, edge1);
                float v = dot(ray.direction, q) * inv_det;
                if (v < 0.0f || u + v > 1.0f) return false;
                float dist = dot(edge2, q) * inv_det;
                if (dist > DIST_EPSILON) {
                    out.distance = dist;
                    out.surface_point = SurfacePoint{ray.origin + ray.direction * dist, K_TRIANGLE, this, nullptr, u, v, {}};
                    return true;
                }
                return false;
            }
            case K_RAY_MARCHED: {  // :120-154
                float min, max;
                if (!bounds.intersect(ray, min, max)) return false;
                Vec3 origin = ray.origin + (-bounds.center());
                float total_distance = min;
                while (total_distance < max) {
                    Vec3 p = origin + ray.direction * total_distance;
                    ++tl_de_evals;
                    float distance = estimator.get(p, &tl_de_iters);
                    // DEVIATION from shapes/mod.rs:127-135 ("a march that cannot end"): far from the origin a step of
                    // >= EPSILON can be smaller than half an ulp of total_distance, the sum does not change and the reference's loop
                    // spins for ever (found on a ray leaving the floor plane 2100 units out: total 2101.72, step 1.12e-4).  Such a
                    // step is treated like one below EPSILON: the march has converged as far as f32 can tell.
                    const bool stuck = total_distance + distance == total_distance;
                    total_distance += distance;
                    if (distance < DIST_EPSILON || total_distance > max || stuck) break;
                }
                if (total_distance <= max) {
                    Vec3 offset_position = origin + ray.direction * (total_distance - DIST_EPSILON);
                    Vec3 p = ray.origin + ray.direction * total_distance;
                    out.distance = total_distance;
                    out.surface_point = SurfacePoint{p, K_RAY_MARCHED, this, nullptr, 0, 0, offset_position};
                    return true;
                }
                return false;
            }
            default: return false;
        }
    }
    // Bounded for Shape (:408-432)
    Aabb aabb() const {
        switch (kind) {
            case K_SPHERE:
                return Aabb::from_points({position.x - radius, position.y - radius, position.z - radius},
                                         {position.x + radius, position.y + radius, position.z + radius});
            case K_TRIANGLE: return Aabb::from_points(v1.position, v2.position).grow(v3.position);
            default: return bounds.aabb();
        }
    }
    // :166-207
    bool sample_point(XorShift& rng, SurfacePoint& sp) const {
        if (kind == K_SPHERE) {
            Vec3 s = sample_sphere(rng);
            sp = SurfacePoint{position + s * radius, K_SPHERE, this, nullptr, 0, 0, {}};
            return true;
        }
        if (kind == K_TRIANGLE) {
            float u = rng.gen_f32();
            float v = rng.gen_f32();
            Vec3 a = v2.position - v1.position;
            Vec3 b = v3.position - v1.position;
            if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
            sp = SurfacePoint{(v1.position + a * u) + b * v, K_TRIANGLE, this, nullptr, u, v, {}};
            return true;
        }
        return false;
    }
    // :209-251
    bool sample_towards(XorShift& rng, Vec3 target, Intersection& out) const {
        if (kind == K_SPHERE) {
            float r = fmax_(radius - DIST_EPSILON, 0.0f);
            Vec3 dir = position - target;
            float dist2 = magnitude2(dir);
            if (dist2 > r * r) {
                float cos_theta_max = sqrtf(fmax_(1.0f - (r * r) / dist2, 0.0f));
                Vec3 ray_dir = sample_cone(rng, normalize(dir), cos_theta_max);
                if (ray_intersect(Ray{target, ray_dir}, out)) return true;
                out.distance = 0.0f;  // "cheat"
                out.surface_point = SurfacePoint{target, K_SPHERE, this, nullptr, 0, 0, {}};
                return true;
            }
        }
        SurfacePoint sp;
        if (!sample_point(rng, sp)) return false;
        out.distance = magnitude(sp.position - target);
        out.surface_point = sp;
        return true;
    }
    // :253-271
    bool solid_angle_towards(Vec3 target, float& a) const {
        if (kind != K_SPHERE) return false;
        float dist2 = magnitude2(position - target);
        if (dist2 > radius * radius) {
            float cos_theta_max = sqrtf(fmax_(1.0f - (radius * radius) / dist2, 0.0f));
            a = solid_angle(cos_theta_max);
            return true;
        }
        return false;
    }
    // :273-288
    float surface_area() const {
        if (kind == K_SPHERE) return radius * radius * 4.0f * PI;
        if (kind == K_TRIANGLE) return 0.5f * magnitude(cross(v2.position - v1.position, v3.position - v1.position));
        return INF;
    }
    void scale(float s) {  // :290-316
        if (kind == K_SPHERE) { radius *= s; position = position * s; }
        else if (kind == K_TRIANGLE) {
            v1.position = v1.position * s; v2.position = v2.position * s; v3.position = v3.position * s;
            edge1 = v2.position - v1.position; edge2 = v3.position - v1.position;
        }
    }
    void transform(const Mat4& t) {  // :318-344
        if (kind == K_SPHERE) position = transform_point(t, position);
        else if (kind == K_TRIANGLE) {
            v1.normal = v1.normal.transform(t); v2.normal = v2.normal.transform(t); v3.normal = v3.normal.transform(t);
            v1.position = transform_point(t, v1.position);
            v2.position = transform_point(t, v2.position);
            v3.position = transform_point(t, v3.position);
            edge1 = v2.position - v1.position; edge2 = v3.position - v1.position;
        }
    }
    SurfaceData surface_data(const SurfacePoint& sp) const {
        if (kind == K_SPHERE) {  // :346-372
            Vec3 normal = normalize(sp.position - position);
            float latitude = m_acos(normal.y);
            float longitude = m_atan2(normal.x, normal.z);
            Mat3 rotation = mat3_from_angle_y(longitude) * mat3_from_angle_x(latitude - PI * 0.5f);
            Vec2 tc{longitude * FRAC_1_PI * 0.5f, 1.0f - (latitude * FRAC_1_PI)};
            return {Normal{normal, quat_from_mat3(rotation)}, Vec2{tc.x / texture_scale.x, tc.y / texture_scale.y}};
        }
        if (kind == K_TRIANGLE) {  // :374-385
            Normal n = Normal::on_triangle(v1.normal, v2.normal, v3.normal, sp.u, sp.v);
            float w = 1.0f - (sp.u + sp.v);
            Vec2 tex{(v1.texture.x * w + v2.texture.x * sp.u) + v3.texture.x * sp.v,
                     (v1.texture.y * w + v2.texture.y * sp.u) + v3.texture.y * sp.v};
            return {n, tex};
        }
        // ray marched :387-405 (six extra DE evaluations)
        Vec3 p = sp.offset_position;
        const float E = DIST_EPSILON;
        auto de = [&](Vec3 q) { ++tl_de_evals; return estimator.get(q, &tl_de_iters); };
        Vec3 n{de(p + Vec3(E, 0, 0)) - de(p + Vec3(-E, 0, 0)), de(p + Vec3(0, E, 0)) - de(p + Vec3(0, -E, 0)),
               de(p + Vec3(0, 0, E)) - de(p + Vec3(0, 0, -E))};
        return {Normal::from_vector(normalize(n)), Vec2{0, 0}};
    }
};

struct PlaneShape {  // shapes/mod.rs:434-470
    uint32_t id = 0;
    Vec3 n; float d = 0;  // collision::Plane::from_point_normal(p, n) = {n, d: p.n}
    Normal normal;
    Vec2 texture_scale{1, 1};
    uint32_t material = 0;
    bool ray_intersect(const Ray& ray, Intersection& out) const {
        // collision Plane x Ray: t = -(d + o.n) / (dir.n); None iff t < 0
        float t = -(d + dot(ray.origin, n)) / dot(ray.direction, n);
        if (t < 0.0f) return false;
        Vec3 p = ray.origin + ray.direction * t;
        out.distance = magnitude(p - ray.origin);
        out.surface_point = SurfacePoint{p, K_PLANE, nullptr, this, 0, 0, {}};
        return true;
    }
    SurfaceData surface_data(Vec3 position) const {
        Vec3 ns = normal.into_space(position);
        return {normal, Vec2{ns.x / texture_scale.x, ns.y / texture_scale.y}};
    }
};

// ---------------------------------------------------------------- BVH (spatial/bvh.rs)
struct FlatBvhNode {
    Aabb bounding_box;
    size_t subtree_size = 0;  // 0 for leaves
    const Shape* item = nullptr;
};

struct Hull {  // bvh.rs:318-370
    Aabb aabbs, centroids;
    static Hull make(const Aabb& a) { Vec3 c = a.center(); return {a, Aabb::from_points(c, c)}; }
    Hull expand(const Aabb& a) const { return {aabbs.unite(a), centroids.grow(a.center())}; }
    Hull join(const Hull& o) const { return {aabbs.unite(o.aabbs), centroids.unite(o.centroids)}; }
    void largest_axis(float& width, int& axis) const {
        Vec3 d = centroids.dim();
        if (d.y > d.x) { width = d.y; axis = 1; } else { width = d.x; axis = 0; }
        if (d.z > width) { width = d.z; axis = 2; }
    }
};

struct Bvh {
    std::vector<FlatBvhNode> nodes;

    // Bvh::new (bvh.rs:13-155) followed by BvhNode::flatten (:250-275)
    static Bvh build(const std::vector<const Shape*>& items) {
        Bvh out;
        if (items.empty()) return out;
        struct TreeNode { Aabb box; size_t subtree_size; int first, second; const Shape* item; };
        std::vector<TreeNode> pool;
        struct Entry { bool join; std::vector<const Shape*> items; Hull hull; Aabb bounding_box; };
        std::vector<Entry> stack;
        std::vector<int> nodes;  // the reference's `nodes` value stack, as indices into pool
        Hull hull = Hull::make(items[0]->aabb());
        for (auto* it : items) hull = hull.expand(it->aabb());
        stack.push_back(Entry{false, items, hull, {}});
        while (!stack.empty()) {
            Entry entry = std::move(stack.back());
            stack.pop_back();
            if (entry.join) {
                int first = nodes.back(); nodes.pop_back();
                int second = nodes.back(); nodes.pop_back();
                size_t sz = pool[first].subtree_size + pool[second].subtree_size + 2;
                pool.push_back(TreeNode{entry.bounding_box, sz, first, second, nullptr});
                nodes.push_back((int)pool.size() - 1);
                continue;
            }
            if (entry.items.size() == 1) {
                pool.push_back(TreeNode{entry.hull.aabbs, 0, -1, -1, entry.items[0]});
                nodes.push_back((int)pool.size() - 1);
                continue;
            }
            float split_axis_width; int split_axis;
            entry.hull.largest_axis(split_axis_width, split_axis);
            std::vector<const Shape*> first_items, second_items;
            Hull first_hull, second_hull;
            if (split_axis_width < DIST_EPSILON) {
                size_t half = entry.items.size() / 2;
                first_items.assign(entry.items.begin(), entry.items.begin() + half);
                second_items.assign(entry.items.begin() + half, entry.items.end());
                first_hull = Hull::make(first_items[0]->aabb());
                for (auto* it : first_items) first_hull = first_hull.expand(it->aabb());
                second_hull = Hull::make(second_items[0]->aabb());
                for (auto* it : second_items) second_hull = second_hull.expand(it->aabb());
            } else {
                constexpr int BUCKETS = 6;
                std::vector<const Shape*> bucket_items[BUCKETS];
                Hull bucket_hull[BUCKETS];
                bool used[BUCKETS] = {false, false, false, false, false, false};
                float min_bound = entry.hull.centroids.min[split_axis];
                for (auto* it : entry.items) {
                    Aabb bb = it->aabb();
                    float position = bb.center()[split_axis];
                    float float_index = (float)BUCKETS * (position - min_bound) / split_axis_width;
                    size_t index = std::min<size_t>(f32_as_usize(float_index), BUCKETS - 1);
                    if (used[index]) { bucket_items[index].push_back(it); bucket_hull[index] = bucket_hull[index].expand(bb); }
                    else { used[index] = true; bucket_items[index].push_back(it); bucket_hull[index] = Hull::make(bb); }
                }
                auto stats = [&](int from, int to, size_t& count, float& area) {  // get_bucket_stats :167-182
                    count = 0; bool any = false; Aabb acc{};
                    for (int i = from; i < to; ++i) {
                        if (!used[i]) continue;
                        acc = any ? acc.unite(bucket_hull[i].aabbs) : bucket_hull[i].aabbs;
                        any = true;
                        count += bucket_items[i].size();
                    }
                    area = any ? acc.surface_area() : 0.0f;
                };
                float min_cost = INF; int min_cost_split = 0;
                float hull_area = entry.hull.aabbs.surface_area();
                for (int index = 1; index < BUCKETS; ++index) {
                    size_t c1, c2; float a1, a2;
                    stats(0, index, c1, a1);
                    stats(index, BUCKETS, c2, a2);
                    float cost = (a1 * (float)c1 + a2 * (float)c2) / hull_area;
                    if (cost < min_cost) { min_cost_split = index; min_cost = cost; }
                }
                auto merge = [&](int from, int to, std::vector<const Shape*>& out_items, Hull& out_hull) {  // merge_buckets :184-199
                    bool any = false;
                    for (int i = from; i < to; ++i) {
                        if (!used[i]) continue;
                        out_hull = any ? bucket_hull[i].join(out_hull) : bucket_hull[i];
                        any = true;
                        out_items.insert(out_items.end(), bucket_items[i].begin(), bucket_items[i].end());
                    }
                    return any;
                };
                if (!merge(0, min_cost_split, first_items, first_hull)) throw std::runtime_error("there should be a first items hull");
                if (!merge(min_cost_split, BUCKETS, second_items, second_hull)) throw std::runtime_error("there should be a second items hull");
            }
            Aabb bb = entry.hull.aabbs;
            stack.push_back(Entry{true, {}, {}, bb});
            stack.push_back(Entry{false, std::move(second_items), second_hull, {}});
            stack.push_back(Entry{false, std::move(first_items), first_hull, {}});
        }
        // flatten: pre-order, `first` before `second` (bvh.rs:250-275).  NB `first` is the node
        // popped first at the Join, i.e. the subtree built from `second_items` (bvh.rs:39-50).
        std::vector<int> fstack{nodes.back()};
        while (!fstack.empty()) {
            int n = fstack.back();
            fstack.pop_back();
            const TreeNode& t = pool[n];
            if (!t.item) { fstack.push_back(t.second); fstack.push_back(t.first); }
            out.nodes.push_back(FlatBvhNode{t.box, t.item ? 0 : t.subtree_size, t.item});
        }
        return out;
    }
};

// ---------------------------------------------------------------- lamps (lamp.rs)
enum LampKind { L_DIRECTIONAL, L_POINT, L_SHAPE };
struct Lamp {
    LampKind kind = L_POINT;
    Vec3 direction; float width = 0;  // directional
    Vec3 position;                    // point
    Program color;                    // directional / point
    const Shape* shape = nullptr;     // shape
};
struct LampSurface {  // lamp.rs:121-128
    bool physical = false;
    Vec3 normal; Vec2 texture; uint32_t material = 0;
    Program color;
};
struct LampSample { Vec3 direction; bool has_sq_distance = false; float sq_distance = 0; LampSurface surface; float weight = 0; };
struct RaySample { Ray ray; LampSurface surface; float weight = 0; };

// ---------------------------------------------------------------- camera (cameras.rs)
struct Camera {
    Mat4 transform;
    float view_plane = 1, focus_distance = 1, aperture = 0;
    // cameras.rs:70-97
    Ray ray_towards(Vec2 target, XorShift& rng) const {
        float focus_x = target.x / view_plane * focus_distance;
        float focus_y = target.y / view_plane * focus_distance;
        Vec3 tgt{focus_x, -focus_y, -focus_distance};
        Vec3 origin{0, 0, 0}, direction = tgt;
        if (aperture > 0.0f) {
            float sqrt_r = sqrtf(aperture * rng.gen_f32());
            float psi = PI * 2.0f * rng.gen_f32();
            origin = {sqrt_r * m_cos(psi), sqrt_r * m_sin(psi), 0.0f};
            direction = tgt - origin;
        }
        return transform_ray(transform, Ray{origin, normalize(direction)});
    }
};

// ---------------------------------------------------------------- world (world.rs)
struct TraceCounters { uint64_t rays = 0, nodes = 0, leaves = 0; };

struct World {
    Project P;  // owns expression nodes (material flattening appends to them)
    Program sky;
    std::vector<Lamp> lights;
    std::vector<PlaneShape> planes;
    std::vector<std::unique_ptr<Shape>> objects;
    std::vector<Material> materials;
    Bvh bvh;

    // World::intersect (world.rs:273-299)
    bool intersect(const Ray& ray, Intersection& result, TraceCounters* tc = nullptr) const {
        bool found = false;
        float closest = INF;
        for (auto& plane : planes) {
            Intersection i;
            if (plane.ray_intersect(ray, i) && i.distance > DIST_EPSILON && i.distance < closest) {
                closest = i.distance; result = i; found = true;
            }
        }
        // Intersections::next (bvh.rs:206-229) inlined: pre-order walk with subtree skips
        const size_t n = bvh.nodes.size();
        uint64_t vn = 0, vt = 0;
        for (size_t idx = 0; idx < n;) {
            const FlatBvhNode& node = bvh.nodes[idx];
            ++vn;
            float d;
            if (aabb_intersection_distance(node.bounding_box, ray, d)) {
                if (d >= closest) { idx += node.subtree_size + 1; continue; }
                if (node.item) {
                    ++vt;
                    Intersection i;
                    if (node.item->ray_intersect(ray, i) && i.distance > DIST_EPSILON && i.distance < closest) {
                        closest = i.distance; result = i; found = true;
                    }
                }
                idx += 1;
            } else {
                idx += node.subtree_size + 1;
            }
        }
        if (tc) { tc->rays += 1; tc->nodes += vn; tc->leaves += vt; }
        return found;
    }
    // world.rs:301-305 (panics on an empty lamp list: gen_range(0..0))
    const Lamp* pick_lamp(XorShift& rng, float& probability) const {
        if (lights.empty()) throw std::runtime_error("cannot sample empty range");
        size_t i = rng.gen_range_usize(lights.size());
        probability = 1.0f / (float)lights.size();
        return &lights[i];
    }
    const Material& material_of(const SurfacePoint& sp) const {
        return materials[sp.kind == K_PLANE ? sp.plane->material : sp.shape->material];
    }
    SurfaceData surface_data(const SurfacePoint& sp) const {  // shapes/mod.rs:484-495
        return sp.kind == K_PLANE ? sp.plane->surface_data(sp.position) : sp.shape->surface_data(sp);
    }
};

// make_triangle (world.rs:308-374)
inline Shape make_triangle(const MeshData& obj, const int32_t* idx, uint32_t material) {
    auto pos = [&](int i) { return Vec3(obj.position[3 * i], obj.position[3 * i + 1], obj.position[3 * i + 2]); };
    auto nor = [&](int i) { return Vec3(obj.normal[3 * i], obj.normal[3 * i + 1], obj.normal[3 * i + 2]); };
    auto tex = [&](int i) { return i < 0 ? Vec2{0, 0} : Vec2{obj.texture[2 * i], obj.texture[2 * i + 1]}; };
    Vec3 v1 = pos(idx[0]), v2 = pos(idx[3]), v3 = pos(idx[6]);
    Vec3 n1, n2, n3;
    if (idx[2] >= 0 && idx[5] >= 0 && idx[8] >= 0) { n1 = nor(idx[2]); n2 = nor(idx[5]); n3 = nor(idx[8]); }
    else { n1 = n2 = n3 = normalize(cross(v2 - v1, v3 - v1)); }
    Vec2 t1 = tex(idx[1]), t2 = tex(idx[4]), t3 = tex(idx[7]);
    Vec3 dp1 = v2 - v1, dp2 = v3 - v1;
    Vec2 dt1{t2.x - t1.x, t2.y - t1.y}, dt2{t3.x - t1.x, t3.y - t1.y};
    float r = 1.0f / (dt1.x * dt2.y - dt1.y * dt2.x);
    Vec3 tangent = (dp1 * dt2.y - dp2 * dt1.y) * r;
    Vec3 bitangent = (dp2 * dt1.x - dp1 * dt2.x) * r;
    Shape s;
    s.kind = K_TRIANGLE;
    s.material = material;
    s.v1 = Vertex{v1, Normal{n1, quat_from_mat3(Mat3::from_cols(tangent, bitangent, n1))}, t1};
    s.v2 = Vertex{v2, Normal{n2, quat_from_mat3(Mat3::from_cols(tangent, bitangent, n2))}, t2};
    s.v3 = Vertex{v3, Normal{n3, quat_from_mat3(Mat3::from_cols(tangent, bitangent, n3))}, t3};
    s.edge1 = dp1;
    s.edge2 = dp2;
    return s;
}

// Transform::evaluate (project/mod.rs:254-268)
inline Mat4 eval_look_at(const Project& P, const LookAt& l) {
    ConstEval ce{P};
    Vec3 from = ce.vec3(l.from), to = ce.vec3(l.to);
    Vec3 up = l.up.present ? ce.vec3(l.up.e) : Vec3(0, 1, 0);
    Mat4 inv;
    if (!invert(look_at(from, to, up), inv)) throw std::runtime_error("could not invert view matrix");
    return inv;
}

// Camera::from_project (cameras.rs:30-55)
inline Camera camera_from_project(const Project& P) {
    ConstEval ce{P};
    Camera c;
    float fov = ce.number(P.fov);
    float fov_radians = (fov * 0.5f) * (PI / 180.0f);  // cgmath Deg -> Rad
    c.view_plane = cosf(fov_radians) / sinf(fov_radians);
    c.transform = eval_look_at(P, P.cam_transform);
    c.focus_distance = P.focus_distance.present ? ce.number(P.focus_distance.e) : 1.0f;
    c.aperture = P.aperture.present ? ce.number(P.aperture.e) : 0.0f;
    return c;
}

// World::from_project (world.rs:39-271)
inline std::unique_ptr<World> world_from_project(Project project) {
    auto W = std::make_unique<World>();
    W->P = std::move(project);
    Project& P = W->P;
    ProgramCompiler pc{P};
    W->sky = pc.compile(P.sky.present ? P.sky.e : Expr::num(0.0), false, ALLOW_RENDER);
    std::vector<size_t> shape_lights;  // indices into objects, resolved to pointers after the loop
    struct PendingLight { bool is_shape; size_t index; Lamp lamp; };
    std::vector<PendingLight> pending;
    const std::vector<WorldObject> objects = P.objects;
    for (size_t i = 0; i < objects.size(); ++i) {
        const WorldObject& o = objects[i];
        ConstEval ce{P};
        switch (o.type) {
            case O_SPHERE: {
                W->materials.push_back(material_from_project(P, o.material));
                bool emissive = W->materials.back().is_emissive();
                auto s = std::make_unique<Shape>();
                s->kind = K_SPHERE;
                if (o.texture_scale.present) { Vec4 ts = ConstEval{P}.vector(o.texture_scale.e); s->texture_scale = {ts.x, ts.y}; }
                s->position = ConstEval{P}.vec3(o.position);
                s->radius = ConstEval{P}.number(o.radius);
                s->material = (uint32_t)W->materials.size() - 1;
                s->id = (uint32_t)W->objects.size();
                if (emissive) pending.push_back({true, W->objects.size(), {}});
                W->objects.push_back(std::move(s));
                break;
            }
            case O_PLANE: {
                W->materials.push_back(material_from_project(P, o.material));
                PlaneShape pl;
                Vec3 normal = normalize(ConstEval{P}.vec3(o.normal));
                Vec3 binormal, tangent;
                basis(normal, binormal, tangent);
                if (o.texture_scale.present) { Vec4 ts = ConstEval{P}.vector(o.texture_scale.e); pl.texture_scale = {ts.x, ts.y}; }
                Vec3 origin = ConstEval{P}.vec3(o.origin);
                pl.n = normal;
                pl.d = dot(origin, normal);
                // SURVEY.md §9 Q1 is an inference (collision 0.20.1 is not vendored): `d = p.n` with `t = -(d + o.n)/(dir.n)` puts the
                // surface through -origin.  PYRO_Q1_LITERAL=1 builds the other reading (surface through +origin) so that
                // tools/reference_images.py can show which one the reference's own renders agree with.  Never set in parity tests.
                if (const char* q1 = getenv("PYRO_Q1_LITERAL")) if (q1[0] == '1') pl.d = -pl.d;
                pl.normal = Normal{normal, quat_from_mat3(Mat3::from_cols(binormal, tangent, normal))};
                pl.material = (uint32_t)W->materials.size() - 1;
                pl.id = (uint32_t)W->planes.size();
                W->planes.push_back(pl);
                break;
            }
            case O_RAY_MARCHED: {
                W->materials.push_back(material_from_project(P, o.material));
                auto s = std::make_unique<Shape>();
                s->kind = K_RAY_MARCHED;
                ConstEval c2{P};
                if (o.bounds_type == 0) { s->bounds.type = 0; s->bounds.a = c2.vec3(o.bmin); s->bounds.b = c2.vec3(o.bmax); }
                else { s->bounds.type = 1; s->bounds.a = c2.vec3(o.bpos); s->bounds.radius = c2.number(o.bradius); }
                Estimator& e = s->estimator;
                e.type = o.estimator;
                e.iterations = c2.u16(o.iterations);
                e.threshold = c2.number(o.threshold);
                if (o.estimator == 0) {
                    e.power = c2.number(o.power);
                    e.has_constant = o.mb_constant.present;
                    if (e.has_constant) e.mb_constant = c2.vec3(o.mb_constant.e);
                } else {
                    Vec4 q = c2.vector(o.constant);
                    e.constant = Quat(q.x, q.y, q.z, q.w);  // expressions.rs:450-454
                    e.slice_plane = c2.number(o.slice_plane);
                    e.variant = o.variant;
                }
                s->material = (uint32_t)W->materials.size() - 1;
                s->id = (uint32_t)W->objects.size();
                W->objects.push_back(std::move(s));
                break;
            }
            case O_MESH: {
                const MeshData& obj = P.meshes.at(o.mesh);
                auto mesh_materials = o.materials;
                for (auto& object : obj.objects) {
                    auto it = std::find_if(mesh_materials.begin(), mesh_materials.end(), [&](auto& kv) { return kv.first == object.name; });
                    if (it == mesh_materials.end())
                        throw std::runtime_error("objects[" + std::to_string(i) + "]: missing material for '" + object.name + "'");
                    MaterialRef ref = it->second;
                    mesh_materials.erase(it);
                    W->materials.push_back(material_from_project(P, ref));
                    uint32_t mat = (uint32_t)W->materials.size() - 1;
                    bool emissive = W->materials.back().is_emissive();
                    Mat4 transform = o.has_transform ? eval_look_at(P, o.transform) : Mat4::identity();
                    float scale = o.scale.present ? ConstEval{P}.number(o.scale.e) : 1.0f;
                    for (size_t t = 0; t < object.tris.size() / 9; ++t) {
                        auto s = std::make_unique<Shape>(make_triangle(obj, &object.tris[9 * t], mat));
                        s->scale(scale);
                        s->transform(transform);
                        s->id = (uint32_t)W->objects.size();
                        if (emissive) pending.push_back({true, W->objects.size(), {}});
                        W->objects.push_back(std::move(s));
                    }
                }
                break;
            }
            case O_DIRECTIONAL_LIGHT: {
                Lamp l;
                l.kind = L_DIRECTIONAL;
                l.direction = ce.vec3(o.direction);
                l.width = ce.number(o.width);
                l.color = pc.compile(o.color, false, ALLOW_RENDER);
                pending.push_back({false, 0, l});
                break;
            }
            case O_POINT_LIGHT: {
                Lamp l;
                l.kind = L_POINT;
                l.position = ce.vec3(o.position);
                l.color = pc.compile(o.color, false, ALLOW_RENDER);
                pending.push_back({false, 0, l});
                break;
            }
        }
    }
    for (auto& pl : pending) {
        if (pl.is_shape) { Lamp l; l.kind = L_SHAPE; l.shape = W->objects[pl.index].get(); W->lights.push_back(l); }
        else W->lights.push_back(pl.lamp);
    }
    std::vector<const Shape*> items;
    for (auto& s : W->objects) items.push_back(s.get());
    W->bvh = Bvh::build(items);
    return W;
}

// Lamp::sample (lamp.rs:23-82)
inline LampSample lamp_sample(const World& W, const Lamp& lamp, XorShift& rng, Vec3 target) {
    LampSample s;
    switch (lamp.kind) {
        case L_DIRECTIONAL:
            s.direction = lamp.width > 0.0f ? sample_cone(rng, lamp.direction, lamp.width) : lamp.direction;
            s.surface.color = lamp.color;
            s.weight = 1.0f;
            return s;
        case L_POINT: {
            Vec3 v = lamp.position - target;
            float distance = magnitude2(v);
            s.direction = normalize(v);
            s.has_sq_distance = true;
            s.sq_distance = distance;
            s.surface.color = lamp.color;
            s.weight = 4.0f * PI / distance;
            return s;
        }
        default: {
            Intersection hit;
            if (!lamp.shape->sample_towards(rng, target, hit)) throw std::runtime_error("trying to use infinite shape in direct lighting");
            Vec3 v = hit.surface_point.position - target;
            float sq_distance = hit.distance * hit.distance;
            Vec3 direction = normalize(v);
            SurfaceData sd = W.surface_data(hit.surface_point);
            float weight;
            if (!lamp.shape->solid_angle_towards(target, weight)) {
                float cos_in = fabsf(dot(sd.normal.vector, -direction));
                weight = cos_in * lamp.shape->surface_area() / sq_distance;
            }
            s.direction = direction;
            s.has_sq_distance = true;
            s.sq_distance = sq_distance;
            s.surface.physical = true;
            s.surface.normal = sd.normal.vector;
            s.surface.texture = sd.texture;
            s.surface.material = lamp.shape->material;
            s.weight = weight;
            return s;
        }
    }
}

// Lamp::sample_ray (lamp.rs:84-113); false == None
inline bool lamp_sample_ray(const World& W, const Lamp& lamp, XorShift& rng, RaySample& out) {
    switch (lamp.kind) {
        case L_DIRECTIONAL: return false;
        case L_POINT: {
            Vec3 direction = sample_sphere(rng);
            out.ray = Ray{lamp.position, direction};
            out.surface = LampSurface{};
            out.surface.color = lamp.color;
            out.weight = 4.0f * PI;
            return true;
        }
        default: {
            SurfacePoint sp;
            if (!lamp.shape->sample_point(rng, sp)) throw std::runtime_error("trying to use infinite shape as lamp");
            SurfaceData sd = W.surface_data(sp);
            Vec3 direction = sample_hemisphere(rng, sd.normal.vector);
            out.ray = Ray{sp.position, direction};
            out.surface = LampSurface{};
            out.surface.physical = true;
            out.surface.normal = sd.normal.vector;
            out.surface.texture = sd.texture;
            out.surface.material = lamp.shape->material;
            out.weight = lamp.shape->surface_area();
            return true;
        }
    }
}

}  // namespace pyro
