// Project IR decoder for the product's scene builder.  The IR is what the reference holds after
// `load_project` (pyrite/src/project/mod.rs:29-93): the typed `Project` plus the node, mesh,
// spectrum and texture tables.  Layout: pyrite_b200/project.py and DESIGN.md §3.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pyr {
namespace ir {

struct BuildError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Ex {  // project/expressions.rs:65-71 `Expression`
    bool is_node = false;
    double number = 0.0;
    uint32_t node = 0;
    static Ex constant(double v) { Ex e; e.number = v; return e; }
    static Ex ref(uint32_t n) { Ex e; e.is_node = true; e.node = n; return e; }
};
struct MaybeEx { bool some = false; Ex ex; };
struct MaybeU32 { bool some = false; uint32_t value = 0; uint32_t unwrap_or(uint32_t d) const { return some ? value : d; } };

enum NodeKind : uint32_t { N_VECTOR, N_RGB, N_BINARY, N_MIX, N_CLAMP, N_FRESNEL, N_BLACKBODY, N_SPECTRUM, N_COLOR_TEXTURE, N_MONO_TEXTURE };
struct ExprNode {  // project/expressions.rs:152-201 `ComplexExpression`
    uint32_t kind = 0, op = 0, resource = 0;
    Ex arg[4];
};
enum SurfaceKind : uint32_t { S_EMISSIVE, S_DIFFUSE, S_MIRROR, S_REFRACTIVE, S_MIX, S_ADD };
struct SurfaceNode {  // project/materials.rs:7-34
    uint32_t kind = 0, lhs = 0, rhs = 0;
    Ex color, ior, amount;
    MaybeEx dispersion, env_ior, env_dispersion;
};
struct SpectrumTable { bool curve = false; float lo = 0, hi = 0; std::vector<float> values; };  // curve: x,y interleaved
struct TextureTable { uint32_t width = 0, height = 0, channels = 0; std::vector<float> texels; };
struct MeshObject { std::string name; std::vector<int32_t> corners; };  // 9 ints per triangle
struct MeshTable { std::vector<float> positions, uvs, normals; std::vector<MeshObject> objects; };
struct MaterialUse { uint32_t surface = 0; MaybeEx normal_map; };
struct LookAtUse { Ex from, to; MaybeEx up; };

enum ObjectKind : uint32_t { OBJ_SPHERE, OBJ_PLANE, OBJ_RAY_MARCHED, OBJ_MESH, OBJ_DIRECTIONAL_LIGHT, OBJ_POINT_LIGHT };
struct SceneObject {  // project/mod.rs:169-203 `WorldObject`
    uint32_t kind = 0;
    Ex a, b, c;            // sphere: position, radius | plane: origin, normal | lights: direction/position, width, color
    MaybeEx texture_scale, mesh_scale;
    MaterialUse material;
    uint32_t estimator = 0, variant = 0, bounds_kind = 0, mesh = 0;
    Ex iterations, threshold, power, julia_constant, slice_plane, bound_a, bound_b;
    MaybeEx bulb_constant;
    std::vector<std::pair<std::string, MaterialUse>> mesh_materials;
    bool has_transform = false;
    LookAtUse transform;
};

struct Document {
    std::vector<ExprNode> nodes;
    std::vector<SurfaceNode> surfaces;
    std::vector<SpectrumTable> spectra;
    std::vector<TextureTable> color_textures, mono_textures;
    std::vector<MeshTable> meshes;
    float burns_lo = 0, burns_hi = 0, xyz_lo = 0, xyz_hi = 0, illum_lo = 0, illum_hi = 0;
    std::vector<float> burns, xyz, d65;
    uint32_t width = 0, height = 0;
    MaybeEx filter, white;
    uint32_t renderer_kind = 0, pixel_samples = 0;
    MaybeU32 threads, bounces, light_samples, spectrum_samples, spectrum_resolution, tile_size, light_bounces;
    LookAtUse camera_transform;
    Ex fov;
    MaybeEx focus_distance, aperture, sky;
    std::vector<SceneObject> objects;
};

class Cursor {
    const unsigned char* at_;
    size_t left_;

  public:
    Cursor(const void* p, size_t n) : at_((const unsigned char*)p), left_(n) {}
    const unsigned char* take(size_t n) {
        if (n > left_) throw BuildError("project IR is truncated");
        const unsigned char* p = at_;
        at_ += n;
        left_ -= n;
        return p;
    }
    template <class T> T scalar() { T v; memcpy(&v, take(sizeof(T)), sizeof(T)); return v; }
    uint32_t u32() { return scalar<uint32_t>(); }
    float f32() { return scalar<float>(); }
    std::string text() {
        uint32_t n = u32();
        const unsigned char* p = take(((size_t)n + 3) & ~(size_t)3);
        return std::string((const char*)p, n);
    }
    // a count read from the blob is checked against the bytes that are left BEFORE anything is allocated:
    // every record of `min_bytes` or more must still fit (a corrupt count cannot make the decoder zero-fill gigabytes)
    size_t count(size_t min_bytes) {
        const size_t n = u32();
        if (min_bytes && n > left_ / min_bytes) throw BuildError("project IR is truncated (a table is longer than the blob)");
        return n;
    }
    template <class T> std::vector<T> block(size_t count) {
        if (count > left_ / sizeof(T)) throw BuildError("project IR is truncated");
        std::vector<T> v(count);
        if (count) memcpy(v.data(), take(count * sizeof(T)), count * sizeof(T));
        return v;
    }
    Ex ex() {
        if (u32() == 0) return Ex::constant(scalar<double>());
        uint32_t id = u32();
        u32();
        return Ex::ref(id);
    }
    MaybeEx maybe_ex() { MaybeEx m; m.some = u32() != 0; if (m.some) m.ex = ex(); return m; }
    MaybeU32 maybe_u32() { MaybeU32 m; m.some = u32() != 0; m.value = u32(); return m; }
    MaterialUse material() { MaterialUse m; m.surface = u32(); m.normal_map = maybe_ex(); return m; }
    LookAtUse look_at() { LookAtUse l; l.from = ex(); l.to = ex(); l.up = maybe_ex(); return l; }
    bool empty() const { return left_ == 0; }
};

inline Document decode(const void* data, size_t bytes) {
    Cursor c(data, bytes);
    Document d;
    if (c.u32() != 0x52495950u) throw BuildError("not a pyrite project IR blob");
    if (c.u32() != 1u) throw BuildError("unsupported project IR version");
    static const int arity[] = {4, 3, 2, 3, 3, 2, 1, 0, 0, 0};
    d.nodes.resize(c.count(4));
    for (auto& n : d.nodes) {
        n.kind = c.u32();
        if (n.kind > N_MONO_TEXTURE) throw BuildError("unknown expression node in IR");
        if (n.kind == N_BINARY) n.op = c.u32();
        if (n.kind >= N_SPECTRUM) n.resource = c.u32();
        for (int i = 0; i < arity[n.kind]; ++i) n.arg[i] = c.ex();
    }
    d.surfaces.resize(c.count(4));
    for (auto& s : d.surfaces) {
        s.kind = c.u32();
        switch (s.kind) {
            case S_EMISSIVE: case S_DIFFUSE: case S_MIRROR: s.color = c.ex(); break;
            case S_REFRACTIVE:
                s.color = c.ex(); s.ior = c.ex(); s.dispersion = c.maybe_ex(); s.env_ior = c.maybe_ex(); s.env_dispersion = c.maybe_ex();
                break;
            case S_MIX: s.lhs = c.u32(); s.rhs = c.u32(); s.amount = c.ex(); break;
            case S_ADD: s.lhs = c.u32(); s.rhs = c.u32(); break;
            default: throw BuildError("unknown surface material node in IR");
        }
    }
    d.spectra.resize(c.count(8));
    for (auto& s : d.spectra) {
        uint32_t kind = c.u32();
        if (kind == 0) { s.lo = c.f32(); s.hi = c.f32(); s.values = c.block<float>(c.u32()); }
        else if (kind == 1) { s.curve = true; s.values = c.block<float>((size_t)c.u32() * 2); }
        else throw BuildError("unknown spectrum kind in IR");
    }
    d.color_textures.resize(c.count(8));
    for (auto& t : d.color_textures) { t.width = c.u32(); t.height = c.u32(); t.channels = 4; t.texels = c.block<float>((size_t)t.width * t.height * 4); }
    d.mono_textures.resize(c.count(8));
    for (auto& t : d.mono_textures) { t.width = c.u32(); t.height = c.u32(); t.channels = 1; t.texels = c.block<float>((size_t)t.width * t.height); }
    d.meshes.resize(c.count(16));
    for (auto& m : d.meshes) {
        m.positions = c.block<float>((size_t)c.u32() * 3);
        m.uvs = c.block<float>((size_t)c.u32() * 2);
        m.normals = c.block<float>((size_t)c.u32() * 3);
        m.objects.resize(c.count(8));
        for (auto& o : m.objects) { o.name = c.text(); o.corners = c.block<int32_t>((size_t)c.u32() * 9); }
    }
    d.burns_lo = c.f32(); d.burns_hi = c.f32(); d.burns = c.block<float>((size_t)c.u32() * 3);
    d.xyz_lo = c.f32(); d.xyz_hi = c.f32(); d.xyz = c.block<float>((size_t)c.u32() * 3);
    d.illum_lo = c.f32(); d.illum_hi = c.f32(); d.d65 = c.block<float>(c.u32());

    d.width = c.u32(); d.height = c.u32();
    d.filter = c.maybe_ex(); d.white = c.maybe_ex();
    d.renderer_kind = c.u32(); d.pixel_samples = c.u32();
    d.threads = c.maybe_u32(); d.bounces = c.maybe_u32(); d.light_samples = c.maybe_u32(); d.spectrum_samples = c.maybe_u32();
    d.spectrum_resolution = c.maybe_u32(); d.tile_size = c.maybe_u32(); d.light_bounces = c.maybe_u32();
    d.camera_transform = c.look_at();
    d.fov = c.ex(); d.focus_distance = c.maybe_ex(); d.aperture = c.maybe_ex();
    d.sky = c.maybe_ex();
    d.objects.resize(c.count(4));
    for (auto& o : d.objects) {
        o.kind = c.u32();
        switch (o.kind) {
            case OBJ_SPHERE: case OBJ_PLANE:
                o.a = c.ex(); o.b = c.ex(); o.texture_scale = c.maybe_ex(); o.material = c.material();
                break;
            case OBJ_RAY_MARCHED:
                o.estimator = c.u32();
                if (o.estimator == 0) { o.iterations = c.ex(); o.threshold = c.ex(); o.power = c.ex(); o.bulb_constant = c.maybe_ex(); }
                else if (o.estimator == 1) { o.iterations = c.ex(); o.threshold = c.ex(); o.julia_constant = c.ex(); o.slice_plane = c.ex(); o.variant = c.u32(); }
                else throw BuildError("unknown distance estimator in IR");
                o.bounds_kind = c.u32();
                if (o.bounds_kind > 1) throw BuildError("unknown bounding volume in IR");
                o.bound_a = c.ex(); o.bound_b = c.ex();
                o.material = c.material();
                break;
            case OBJ_MESH: {
                o.mesh = c.u32();
                uint32_t n = (uint32_t)c.count(8);
                for (uint32_t i = 0; i < n; ++i) { std::string name = c.text(); o.mesh_materials.emplace_back(name, c.material()); }
                o.mesh_scale = c.maybe_ex();
                o.has_transform = c.u32() != 0;
                if (o.has_transform) o.transform = c.look_at();
                break;
            }
            case OBJ_DIRECTIONAL_LIGHT: o.a = c.ex(); o.b = c.ex(); o.c = c.ex(); break;
            case OBJ_POINT_LIGHT: o.a = c.ex(); o.c = c.ex(); break;
            default: throw BuildError("unknown world object in IR");
        }
    }
    if (!c.empty()) throw BuildError("trailing bytes after project IR");
    return d;
}

}  // namespace ir
}  // namespace pyr
