// ORACLE - TEST INFRASTRUCTURE ONLY (see pyro_math.hpp header).
//
// Reader for the project IR (pyrite_b200/project.py): the typed
// `project::Project` + `ProjectData{nodes, meshes, spectra, textures}` the reference holds
// after `load_project` (pyrite/src/project/mod.rs:29-93, 95-269).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pyro_math.hpp"

namespace pyro {

struct Expr {  // project/expressions.rs:65-71
    uint32_t tag = 0;  // 0 = Number(f64), 1 = Complex(node id)
    double number = 0.0;
    uint32_t id = 0;
    static Expr num(double v) { Expr e; e.tag = 0; e.number = v; return e; }
    static Expr node(uint32_t i) { Expr e; e.tag = 1; e.id = i; return e; }
};
struct OptExpr {
    bool present = false;
    Expr e;
};

enum ExprType : uint32_t { E_VECTOR, E_RGB, E_BINARY, E_MIX, E_CLAMP, E_FRESNEL, E_BLACKBODY, E_SPECTRUM, E_COLOR_TEXTURE, E_MONO_TEXTURE };
enum BinOp : uint32_t { OP_ADD, OP_SUB, OP_MUL, OP_DIV };

struct ExprNode {  // project/expressions.rs:152-201
    uint32_t type = 0;
    uint32_t op = 0;        // BINARY
    Expr a, b, c, d;        // VECTOR x,y,z,w | RGB r,g,b | BINARY lhs,rhs | MIX amount,lhs,rhs | CLAMP value,min,max | FRESNEL ior,env | BLACKBODY temperature
    uint32_t resource = 0;  // SPECTRUM / *_TEXTURE id
};

enum MatType : uint32_t { M_EMISSIVE, M_DIFFUSE, M_MIRROR, M_REFRACTIVE, M_MIX, M_ADD };
struct MatNode {  // project/materials.rs:7-34
    uint32_t type = 0;
    Expr color, ior, amount;
    OptExpr dispersion, env_ior, env_dispersion;
    uint32_t lhs = 0, rhs = 0;
};

struct SpectrumData {  // project/spectra.rs:13-24
    bool is_curve = false;
    float min = 0, max = 0;
    std::vector<float> points;                    // Array
    std::vector<std::pair<float, float>> curve;   // Curve
};

struct TextureData {
    uint32_t width = 0, height = 0;
    std::vector<float> data;  // RGBA (color) or luma (mono), row-major, row 0 = top
};

struct MeshObject {
    std::string name;
    std::vector<int32_t> tris;  // 9 per triangle: (v,t,n) x 3, -1 = None
};
struct MeshData {
    std::vector<float> position, texture, normal;
    std::vector<MeshObject> objects;
};

struct MaterialRef {  // project/mod.rs:240-244
    uint32_t surface = 0;
    OptExpr normal_map;
};
struct LookAt {
    Expr from, to;
    OptExpr up;
};

enum ObjType : uint32_t { O_SPHERE, O_PLANE, O_RAY_MARCHED, O_MESH, O_DIRECTIONAL_LIGHT, O_POINT_LIGHT };
struct WorldObject {  // project/mod.rs:169-203
    uint32_t type = 0;
    Expr position, radius, origin, normal, direction, width, color;
    OptExpr texture_scale, scale;
    MaterialRef material;
    // ray marched
    uint32_t estimator = 0;  // 0 mandelbulb, 1 quaternion julia
    Expr iterations, threshold, power, constant, slice_plane;
    OptExpr mb_constant;
    uint32_t variant = 0;    // 0 regular 1 cubic 2 bicomplex
    uint32_t bounds_type = 0;  // 0 box 1 sphere
    Expr bmin, bmax, bpos, bradius;
    // mesh
    uint32_t mesh = 0;
    std::vector<std::pair<std::string, MaterialRef>> materials;
    bool has_transform = false;
    LookAt transform;
};

struct OptU32 {
    bool present = false;
    uint32_t v = 0;
    uint32_t or_(uint32_t d) const { return present ? v : d; }
};

struct Project {
    std::vector<ExprNode> exprs;
    std::vector<MatNode> mats;
    std::vector<SpectrumData> spectra;
    std::vector<TextureData> color_textures, mono_textures;
    std::vector<MeshData> meshes;
    // constant resources (reference: build.rs -> rgb.rs / xyz.rs / light_source.rs)
    float burns_min = 0, burns_max = 0;
    std::vector<float> burns;  // r,g,b interleaved
    float xyz_min = 0, xyz_max = 0;
    std::vector<float> xyz;    // x,y,z interleaved
    float illum_min = 0, illum_max = 0;
    std::vector<float> d65;
    // image (project/mod.rs:111-118)
    uint32_t width = 0, height = 0;
    OptExpr filter, white;
    // renderer (project/mod.rs:131-161)
    uint32_t renderer_type = 0, pixel_samples = 0;
    OptU32 threads, bounces, light_samples, spectrum_samples, spectrum_resolution, tile_size, light_bounces;
    // camera (project/mod.rs:120-129)
    LookAt cam_transform;
    Expr fov;
    OptExpr focus_distance, aperture;
    // world
    OptExpr sky;
    std::vector<WorldObject> objects;
};

class Reader {
    const uint8_t* p_;
    const uint8_t* end_;

  public:
    Reader(const void* data, size_t n) : p_((const uint8_t*)data), end_((const uint8_t*)data + n) {}
    void need(size_t n) {
        if ((size_t)(end_ - p_) < n) throw std::runtime_error("project IR truncated");
    }
    uint32_t u32() { need(4); uint32_t v; memcpy(&v, p_, 4); p_ += 4; return v; }
    float f32() { need(4); float v; memcpy(&v, p_, 4); p_ += 4; return v; }
    double f64() { need(8); double v; memcpy(&v, p_, 8); p_ += 8; return v; }
    std::string str() {
        uint32_t n = u32();
        size_t padded = (n + 3u) & ~3u;
        need(padded);
        std::string s((const char*)p_, n);
        p_ += padded;
        return s;
    }
    template <class T>
    void arr(std::vector<T>& out, size_t count) {
        need(count * sizeof(T));
        out.resize(count);
        if (count) memcpy(out.data(), p_, count * sizeof(T));
        p_ += count * sizeof(T);
    }
    Expr expr() {
        uint32_t tag = u32();
        if (tag == 0) return Expr::num(f64());
        uint32_t id = u32();
        u32();
        return Expr::node(id);
    }
    OptExpr opt_expr() {
        OptExpr o;
        o.present = u32() != 0;
        if (o.present) o.e = expr();
        return o;
    }
    OptU32 opt_u32() {
        OptU32 o;
        o.present = u32() != 0;
        o.v = u32();
        return o;
    }
    MaterialRef material() {
        MaterialRef m;
        m.surface = u32();
        m.normal_map = opt_expr();
        return m;
    }
    LookAt look_at() {
        LookAt l;
        l.from = expr();
        l.to = expr();
        l.up = opt_expr();
        return l;
    }
    bool done() const { return p_ == end_; }
};

inline Project parse_project_ir(const void* data, size_t n) {
    Reader r(data, n);
    Project P;
    if (r.u32() != 0x52495950u) throw std::runtime_error("not a project IR blob (bad magic)");
    if (r.u32() != 1u) throw std::runtime_error("unsupported project IR version");
    uint32_t ne = r.u32();
    P.exprs.resize(ne);
    for (auto& e : P.exprs) {
        e.type = r.u32();
        switch (e.type) {
            case E_VECTOR: e.a = r.expr(); e.b = r.expr(); e.c = r.expr(); e.d = r.expr(); break;
            case E_RGB: e.a = r.expr(); e.b = r.expr(); e.c = r.expr(); break;
            case E_BINARY: e.op = r.u32(); e.a = r.expr(); e.b = r.expr(); break;
            case E_MIX: e.a = r.expr(); e.b = r.expr(); e.c = r.expr(); break;
            case E_CLAMP: e.a = r.expr(); e.b = r.expr(); e.c = r.expr(); break;
            case E_FRESNEL: e.a = r.expr(); e.b = r.expr(); break;
            case E_BLACKBODY: e.a = r.expr(); break;
            case E_SPECTRUM: case E_COLOR_TEXTURE: case E_MONO_TEXTURE: e.resource = r.u32(); break;
            default: throw std::runtime_error("unknown expression node type in IR");
        }
    }
    uint32_t nm = r.u32();
    P.mats.resize(nm);
    for (auto& m : P.mats) {
        m.type = r.u32();
        switch (m.type) {
            case M_EMISSIVE: case M_DIFFUSE: case M_MIRROR: m.color = r.expr(); break;
            case M_REFRACTIVE:
                m.color = r.expr(); m.ior = r.expr();
                m.dispersion = r.opt_expr(); m.env_ior = r.opt_expr(); m.env_dispersion = r.opt_expr();
                break;
            case M_MIX: m.lhs = r.u32(); m.rhs = r.u32(); m.amount = r.expr(); break;
            case M_ADD: m.lhs = r.u32(); m.rhs = r.u32(); break;
            default: throw std::runtime_error("unknown material node type in IR");
        }
    }
    uint32_t ns = r.u32();
    P.spectra.resize(ns);
    for (auto& s : P.spectra) {
        uint32_t kind = r.u32();
        if (kind == 0) {
            s.min = r.f32(); s.max = r.f32();
            r.arr(s.points, r.u32());
        } else if (kind == 1) {
            s.is_curve = true;
            uint32_t n2 = r.u32();
            std::vector<float> flat;
            r.arr(flat, (size_t)n2 * 2);
            for (uint32_t i = 0; i < n2; ++i) s.curve.emplace_back(flat[2 * i], flat[2 * i + 1]);
        } else throw std::runtime_error("unknown spectrum kind in IR");
    }
    for (int pass = 0; pass < 2; ++pass) {
        auto& list = pass == 0 ? P.color_textures : P.mono_textures;
        list.resize(r.u32());
        for (auto& t : list) {
            t.width = r.u32(); t.height = r.u32();
            r.arr(t.data, (size_t)t.width * t.height * (pass == 0 ? 4 : 1));
        }
    }
    P.meshes.resize(r.u32());
    for (auto& m : P.meshes) {
        r.arr(m.position, (size_t)r.u32() * 3);
        r.arr(m.texture, (size_t)r.u32() * 2);
        r.arr(m.normal, (size_t)r.u32() * 3);
        m.objects.resize(r.u32());
        for (auto& o : m.objects) {
            o.name = r.str();
            r.arr(o.tris, (size_t)r.u32() * 9);
        }
    }
    P.burns_min = r.f32(); P.burns_max = r.f32(); r.arr(P.burns, (size_t)r.u32() * 3);
    P.xyz_min = r.f32(); P.xyz_max = r.f32(); r.arr(P.xyz, (size_t)r.u32() * 3);
    P.illum_min = r.f32(); P.illum_max = r.f32(); r.arr(P.d65, r.u32());

    P.width = r.u32(); P.height = r.u32();
    P.filter = r.opt_expr(); P.white = r.opt_expr();
    P.renderer_type = r.u32(); P.pixel_samples = r.u32();
    P.threads = r.opt_u32(); P.bounces = r.opt_u32(); P.light_samples = r.opt_u32(); P.spectrum_samples = r.opt_u32();
    P.spectrum_resolution = r.opt_u32(); P.tile_size = r.opt_u32(); P.light_bounces = r.opt_u32();
    P.cam_transform = r.look_at();
    P.fov = r.expr(); P.focus_distance = r.opt_expr(); P.aperture = r.opt_expr();
    P.sky = r.opt_expr();
    P.objects.resize(r.u32());
    for (auto& o : P.objects) {
        o.type = r.u32();
        switch (o.type) {
            case O_SPHERE:
                o.position = r.expr(); o.radius = r.expr(); o.texture_scale = r.opt_expr(); o.material = r.material();
                break;
            case O_PLANE:
                o.origin = r.expr(); o.normal = r.expr(); o.texture_scale = r.opt_expr(); o.material = r.material();
                break;
            case O_RAY_MARCHED:
                o.estimator = r.u32();
                if (o.estimator == 0) {
                    o.iterations = r.expr(); o.threshold = r.expr(); o.power = r.expr(); o.mb_constant = r.opt_expr();
                } else if (o.estimator == 1) {
                    o.iterations = r.expr(); o.threshold = r.expr(); o.constant = r.expr(); o.slice_plane = r.expr();
                    o.variant = r.u32();
                } else throw std::runtime_error("unknown estimator in IR");
                o.bounds_type = r.u32();
                if (o.bounds_type == 0) { o.bmin = r.expr(); o.bmax = r.expr(); }
                else if (o.bounds_type == 1) { o.bpos = r.expr(); o.bradius = r.expr(); }
                else throw std::runtime_error("unknown bounds in IR");
                o.material = r.material();
                break;
            case O_MESH: {
                o.mesh = r.u32();
                uint32_t k = r.u32();
                for (uint32_t i = 0; i < k; ++i) {
                    std::string name = r.str();
                    o.materials.emplace_back(name, r.material());
                }
                o.scale = r.opt_expr();
                o.has_transform = r.u32() != 0;
                if (o.has_transform) o.transform = r.look_at();
                break;
            }
            case O_DIRECTIONAL_LIGHT: o.direction = r.expr(); o.width = r.expr(); o.color = r.expr(); break;
            case O_POINT_LIGHT: o.position = r.expr(); o.color = r.expr(); break;
            default: throw std::runtime_error("unknown world object type in IR");
        }
    }
    if (!r.done()) throw std::runtime_error("trailing bytes in project IR");
    return P;
}

}  // namespace pyro
