// Host scene builder (see scene_build.hpp).  Compiled with -ffp-contract=off.
#include "scene_build.hpp"

#include <algorithm>
#include <thread>
#include <string>
#include <mutex>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

#include "host_math.hpp"

namespace pyr {
using namespace host;
using ir::BuildError;
using ir::Document;
using ir::Ex;
using ir::MaybeEx;

namespace {

inline float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline f4 pack4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline f4 quat4(Q4 q) { return pack4(q.s, q.x, q.y, q.z); }

// ------------------------------------------------------------------ constant folding
// `Evaluate` for ComplexExpression with T = f32 / Vector (project/expressions.rs:203-347).
struct Folder {
    const Document& doc;
    float scalar(const Ex& e, int depth = 0) const {
        if (!e.is_node) return (float)e.number;
        if (depth > 1024) throw BuildError("expression graph is cyclic");
        const ir::ExprNode& n = doc.nodes.at(e.node);
        switch (n.kind) {
            case ir::N_VECTOR: throw BuildError("expected a number, but found a vector");
            case ir::N_RGB: throw BuildError("expected a number, but found an RGB color");
            case ir::N_BINARY: {
                float l = scalar(n.arg[0], depth + 1), r = scalar(n.arg[1], depth + 1);
                return n.op == 0 ? l + r : n.op == 1 ? l - r : n.op == 2 ? l * r : l / r;
            }
            case ir::N_MIX: {
                float a = scalar(n.arg[0], depth + 1), l = scalar(n.arg[1], depth + 1), r = scalar(n.arg[2], depth + 1);
                a = fmaxf(fminf(a, 1.0f), 0.0f);
                return l * (1.0f - a) + r * a;
            }
            case ir::N_CLAMP: {
                float v = scalar(n.arg[0], depth + 1), lo = scalar(n.arg[1], depth + 1), hi = scalar(n.arg[2], depth + 1);
                return fmaxf(fminf(v, hi), lo);
            }
            case ir::N_FRESNEL: throw BuildError("cannot evaluate Fresnel functions as constants");
            case ir::N_BLACKBODY: throw BuildError("cannot evaluate black-body functions as constants");
            case ir::N_SPECTRUM: throw BuildError("cannot evaluate spectra as constants");
            default: throw BuildError("cannot evaluate textures as constants");
        }
    }
    V4 quad(const Ex& e, int depth = 0) const {
        if (!e.is_node) { float v = (float)e.number; return V4{v, v, v, v}; }
        if (depth > 1024) throw BuildError("expression graph is cyclic");
        const ir::ExprNode& n = doc.nodes.at(e.node);
        switch (n.kind) {
            case ir::N_VECTOR: return V4{scalar(n.arg[0]), scalar(n.arg[1]), scalar(n.arg[2]), scalar(n.arg[3])};
            case ir::N_RGB: throw BuildError("expected a vector, but found an RGB color");
            case ir::N_BINARY: {
                V4 l = quad(n.arg[0], depth + 1), r = quad(n.arg[1], depth + 1);
                switch (n.op) {
                    case 0: return V4{l.x + r.x, l.y + r.y, l.z + r.z, l.w + r.w};
                    case 1: return V4{l.x - r.x, l.y - r.y, l.z - r.z, l.w - r.w};
                    case 2: return V4{l.x * r.x, l.y * r.y, l.z * r.z, l.w * r.w};
                    default: return V4{l.x / r.x, l.y / r.y, l.z / r.z, l.w / r.w};
                }
            }
            case ir::N_MIX: {
                float a = fmaxf(fminf(scalar(n.arg[0], depth + 1), 1.0f), 0.0f);
                V4 l = quad(n.arg[1], depth + 1), r = quad(n.arg[2], depth + 1);
                return V4{l.x + (r.x - l.x) * a, l.y + (r.y - l.y) * a, l.z + (r.z - l.z) * a, l.w + (r.w - l.w) * a};
            }
            case ir::N_CLAMP: throw BuildError("vectors cannot be clamped");
            case ir::N_FRESNEL: throw BuildError("cannot evaluate Fresnel functions as constants");
            case ir::N_BLACKBODY: throw BuildError("cannot evaluate black-body functions as constants");
            case ir::N_SPECTRUM: throw BuildError("cannot evaluate spectra as constants");
            default: throw BuildError("cannot evaluate textures as constants");
        }
    }
    V3 triple(const Ex& e) const { V4 v = quad(e); return V3{v.x, v.y, v.z}; }
};

// ------------------------------------------------------------------ expression -> bytecode
// The reference allocates one fresh register per value (compiler.rs RegisterCounter); we do the
// same over a single 16-entry file, which is what keeps the memoised re-run valid on the device.
enum Lane { LANE_NUMBER = VT_NUMBER, LANE_VECTOR = VT_VECTOR, LANE_RGB = VT_RGB };
constexpr uint32_t ALLOW_RENDER = IN_WAVELENGTH | IN_NORMAL | IN_INCIDENT | IN_TEXTURE;
constexpr uint32_t ALLOW_NORMAL_MAP = IN_NORMAL | IN_INCIDENT | IN_TEXTURE;
constexpr uint32_t ALLOW_SPECTRUM_ONLY = IN_WAVELENGTH;

struct Emitter {
    Document& doc;
    BakedScene& out;
    uint32_t allowed = 0;
    std::vector<Instr> body;
    uint32_t next_reg = 0;
    uint32_t reads = 0;
    struct Slot { bool done = false; uint8_t reg = 0; uint8_t lane = 0; uint8_t deps = 0; };
    std::vector<Slot> slots;

    struct Operand { bool is_const; float value; uint8_t reg, lane, deps; };

    uint8_t fresh() {
        if (next_reg >= (uint32_t)VM_REGS) throw BuildError("expression needs more than 16 registers");
        return (uint8_t)next_reg++;
    }
    void require(uint32_t bit) {
        if (!(allowed & bit)) {
            if (bit == IN_WAVELENGTH) throw BuildError("the wavelength is not available in this kind of expression");
            if (bit == IN_NORMAL) throw BuildError("the surface normal is not available in this kind of expression");
            if (bit == IN_INCIDENT) throw BuildError("the incident vector is not available in this kind of expression");
            throw BuildError("texture coordinates are not available in this kind of expression");
        }
    }
    Instr blank(uint8_t op, uint8_t reg, uint8_t deps) {
        Instr i;
        memset(&i, 0, sizeof(i));
        i.op = op; i.out = reg; i.deps = deps;
        return i;
    }
    void set_operand(Instr& i, int k, const Operand& o) {
        if (o.is_const) { i.is_reg[k] = 0; i.v[k].f = o.value; }
        else { i.is_reg[k] = 1; i.v[k].u = o.reg; }
    }
    void push(const Instr& i) { body.push_back(i); reads |= i.deps; }

    Operand visit(const Ex& e, int depth) {
        if (!e.is_node) return Operand{true, (float)e.number, 0, LANE_NUMBER, 0};
        if (depth > 1024) throw BuildError("expression graph is cyclic");
        if (e.node >= doc.nodes.size()) throw BuildError("expression id out of range");
        if (slots.size() < doc.nodes.size()) slots.resize(doc.nodes.size());
        if (!slots[e.node].done) {
            compile_node(e.node, depth);
        }
        const Slot& s = slots[e.node];
        return Operand{false, 0.0f, s.reg, s.lane, s.deps};
    }
    // try_get_number_value (compiler.rs:638-680): rgb values are sampled at the wavelength
    Operand as_number(const Ex& e, int depth) {
        Operand o = visit(e, depth);
        if (o.is_const || o.lane == LANE_NUMBER) return o;
        if (o.lane == LANE_VECTOR) throw BuildError("cannot use a vector as a number");
        require(IN_WAVELENGTH);
        uint8_t deps = o.deps | IN_WAVELENGTH;
        Instr i = blank(OP_RGB_SPECTRUM, fresh(), deps);
        i.is_reg[0] = 1; i.v[0].u = o.reg;
        push(i);
        return Operand{false, 0.0f, i.out, LANE_NUMBER, deps};
    }
    // the register-or-constant -> register coercions of convert_operands (compiler.rs:682-968)
    Operand to_lane(const Operand& o, uint8_t lane) {
        if (!o.is_const && o.lane == lane) return o;
        Instr i;
        if (o.is_const) {
            if (lane == LANE_NUMBER) { i = blank(OP_NUMBER, fresh(), 0); i.v[0].f = o.value; }
            else if (lane == LANE_RGB) { i = blank(OP_RGB, fresh(), 0); for (int k = 0; k < 3; ++k) i.v[k].f = o.value; }
            else { i = blank(OP_VECTOR, fresh(), 0); for (int k = 0; k < 4; ++k) i.v[k].f = o.value; }
        } else if (o.lane == LANE_NUMBER) {
            i = blank(lane == LANE_RGB ? OP_NUM_TO_RGB : OP_NUM_TO_VEC, fresh(), o.deps);
            i.is_reg[0] = 1; i.v[0].u = o.reg;
        } else if (o.lane == LANE_RGB && lane == LANE_VECTOR) {
            i = blank(OP_RGB_TO_VEC, fresh(), o.deps);
            i.is_reg[0] = 1; i.v[0].u = o.reg;
        } else {
            throw BuildError("cannot convert a vector into a color or a number");
        }
        push(i);
        return Operand{false, 0.0f, i.out, lane, i.deps};
    }
    static uint8_t common_lane(const Operand& l, const Operand& r) {
        uint8_t a = l.is_const ? (uint8_t)LANE_NUMBER : l.lane, b = r.is_const ? (uint8_t)LANE_NUMBER : r.lane;
        if (a == b) return a;
        if (a == LANE_VECTOR || b == LANE_VECTOR) return LANE_VECTOR;
        return LANE_RGB;
    }
    void compile_node(uint32_t id, int depth) {
        const ir::ExprNode n = doc.nodes[id];
        Slot s;
        switch (n.kind) {
            case ir::N_VECTOR: case ir::N_RGB: {
                const int count = n.kind == ir::N_VECTOR ? 4 : 3;
                Operand o[4];
                uint8_t deps = 0;
                for (int k = 0; k < count; ++k) { o[k] = as_number(n.arg[k], depth + 1); deps |= o[k].deps; }
                Instr i = blank(n.kind == ir::N_VECTOR ? OP_VECTOR : OP_RGB, fresh(), deps);
                for (int k = 0; k < count; ++k) set_operand(i, k, o[k]);
                push(i);
                s = Slot{true, i.out, (uint8_t)(n.kind == ir::N_VECTOR ? LANE_VECTOR : LANE_RGB), deps};
                break;
            }
            case ir::N_FRESNEL: {
                require(IN_NORMAL); require(IN_INCIDENT);
                Operand a = as_number(n.arg[0], depth + 1), b = as_number(n.arg[1], depth + 1);
                uint8_t deps = a.deps | b.deps | IN_NORMAL | IN_INCIDENT;
                Instr i = blank(OP_FRESNEL, fresh(), deps);
                set_operand(i, 0, a); set_operand(i, 1, b);
                push(i);
                s = Slot{true, i.out, LANE_NUMBER, deps};
                break;
            }
            case ir::N_BLACKBODY: {
                require(IN_WAVELENGTH);
                Operand a = as_number(n.arg[0], depth + 1);
                uint8_t deps = a.deps | IN_WAVELENGTH;
                Instr i = blank(OP_BLACKBODY, fresh(), deps);
                set_operand(i, 0, a);
                push(i);
                s = Slot{true, i.out, LANE_NUMBER, deps};
                break;
            }
            case ir::N_SPECTRUM: {
                require(IN_WAVELENGTH);
                if (n.resource >= doc.spectra.size()) throw BuildError("spectrum id out of range");
                Instr i = blank(OP_SPECTRUM, fresh(), IN_WAVELENGTH);
                i.resource = n.resource;
                push(i);
                s = Slot{true, i.out, LANE_NUMBER, IN_WAVELENGTH};
                break;
            }
            case ir::N_COLOR_TEXTURE: case ir::N_MONO_TEXTURE: {
                require(IN_TEXTURE);
                const bool color = n.kind == ir::N_COLOR_TEXTURE;
                if (n.resource >= (color ? doc.color_textures.size() : doc.mono_textures.size())) throw BuildError("texture id out of range");
                Instr i = blank(color ? OP_COLOR_TEXTURE : OP_MONO_TEXTURE, fresh(), IN_TEXTURE);
                i.resource = n.resource;
                push(i);
                s = Slot{true, i.out, (uint8_t)(color ? LANE_RGB : LANE_NUMBER), IN_TEXTURE};
                break;
            }
            case ir::N_MIX: {
                Operand amount = as_number(n.arg[0], depth + 1);
                Operand l = visit(n.arg[1], depth + 1), r = visit(n.arg[2], depth + 1);
                uint8_t lane = common_lane(l, r);
                l = to_lane(l, lane); r = to_lane(r, lane);
                uint8_t deps = amount.deps | l.deps | r.deps;
                Instr i = blank(OP_MIX, fresh(), deps);
                i.vtype = lane;
                set_operand(i, 0, amount); set_operand(i, 1, l); set_operand(i, 2, r);
                push(i);
                s = Slot{true, i.out, lane, deps};
                break;
            }
            case ir::N_BINARY: {
                Operand l = visit(n.arg[0], depth + 1), r = visit(n.arg[1], depth + 1);
                uint8_t lane = common_lane(l, r);
                l = to_lane(l, lane); r = to_lane(r, lane);
                uint8_t deps = l.deps | r.deps;
                Instr i = blank(OP_BINARY, fresh(), deps);
                i.vtype = lane; i.binop = (uint8_t)n.op;
                set_operand(i, 0, l); set_operand(i, 1, r);
                push(i);
                s = Slot{true, i.out, lane, deps};
                break;
            }
            case ir::N_CLAMP: {
                Operand v = as_number(n.arg[0], depth + 1), lo = as_number(n.arg[1], depth + 1), hi = as_number(n.arg[2], depth + 1);
                uint8_t deps = v.deps | lo.deps | hi.deps;
                Instr i = blank(OP_CLAMP, fresh(), deps);
                set_operand(i, 0, v); set_operand(i, 1, lo); set_operand(i, 2, hi);
                push(i);
                s = Slot{true, i.out, LANE_NUMBER, deps};
                break;
            }
            default: throw BuildError("unknown expression node");
        }
        if (slots.size() < doc.nodes.size()) slots.resize(doc.nodes.size());
        slots[id] = s;
    }
};

// ProgramCompiler::compile; returns the index of the new ProgramRec
int32_t compile_program(Document& doc, BakedScene& out, const Ex& root, bool vector_output, uint32_t allowed) {
    ProgramRec rec;
    memset(&rec, 0, sizeof(rec));
    if (!root.is_node) {
        rec.is_constant = 1;
        rec.value = (float)root.number;
    } else {
        Emitter em{doc, out, allowed, {}, 0, 0, {}};
        Emitter::Operand o = em.visit(root, 0);
        if (vector_output) o = em.to_lane(o, LANE_VECTOR);
        else {
            if (o.lane == LANE_VECTOR) throw BuildError("cannot use a vector as a number");
            if (o.lane == LANE_RGB) {
                em.require(IN_WAVELENGTH);
                Instr i = em.blank(OP_RGB_SPECTRUM, em.fresh(), o.deps | IN_WAVELENGTH);
                i.is_reg[0] = 1; i.v[0].u = o.reg;
                em.push(i);
                o = Emitter::Operand{false, 0.0f, i.out, LANE_NUMBER, i.deps};
            }
        }
        rec.code_offset = (uint32_t)out.code.size();
        rec.n_instr = (uint32_t)em.body.size();
        rec.out_reg = o.reg;
        rec.reads = em.reads;
        out.code.insert(out.code.end(), em.body.begin(), em.body.end());
        rec.wl_offset = (uint32_t)out.code.size();
        for (const Instr& i : em.body)
            if (i.deps & IN_WAVELENGTH) out.code.push_back(i);
        rec.wl_count = (uint32_t)out.code.size() - rec.wl_offset;
    }
    out.programs.push_back(rec);
    return (int32_t)out.programs.size() - 1;
}

// ------------------------------------------------------------------ materials
// project/expressions.rs:20-63: new nodes, folded in f64 when every operand is a literal
Ex add_node(Document& doc, uint32_t kind, uint32_t op, Ex a, Ex b, Ex c = Ex()) {
    ir::ExprNode n;
    n.kind = kind; n.op = op; n.arg[0] = a; n.arg[1] = b; n.arg[2] = c;
    doc.nodes.push_back(n);
    return Ex::ref((uint32_t)doc.nodes.size() - 1);
}
Ex make_product(Document& doc, Ex l, Ex r) { return (!l.is_node && !r.is_node) ? Ex::constant(l.number * r.number) : add_node(doc, ir::N_BINARY, 2, l, r); }
Ex make_difference(Document& doc, Ex l, Ex r) { return (!l.is_node && !r.is_node) ? Ex::constant(l.number - r.number) : add_node(doc, ir::N_BINARY, 1, l, r); }
Ex make_clamp(Document& doc, Ex v, Ex lo, Ex hi) {
    if (!v.is_node && !lo.is_node && !hi.is_node) return Ex::constant(std::fmax(std::fmin(v.number, hi.number), lo.number));
    return add_node(doc, ir::N_CLAMP, 0, v, lo, hi);
}

// Material::from_project (materials/mod.rs:33-46) + SurfaceMaterial::from_project (:89-228)
uint32_t bake_material(Document& doc, BakedScene& out, const ir::MaterialUse& use) {
    // `depth` guards against a cyclic mix / add graph in a hand-made IR blob (the Lua loader cannot produce one);
    // the worklist is depth-first, so a cycle reaches the limit after ~1024 steps instead of growing for ever
    struct Work { uint32_t surface; bool weighted; Ex weight; uint32_t depth; };
    std::vector<Work> todo{{use.surface, false, Ex(), 0u}};
    std::vector<ComponentRec> all, glowing;
    Folder fold{doc};
    while (!todo.empty()) {
        Work w = todo.back();
        todo.pop_back();
        if (w.surface >= doc.surfaces.size()) throw BuildError("surface material id out of range");
        if (w.depth > 1024) throw BuildError("surface material graph is cyclic");
        const ir::SurfaceNode sn = doc.surfaces[w.surface];
        if (sn.kind == ir::S_MIX) {  // `amount` weights lhs (materials/mod.rs:176-195)
            Ex amount = make_clamp(doc, sn.amount, Ex::constant(0.0), Ex::constant(1.0));
            Ex lhs_weight = w.weighted ? make_product(doc, w.weight, amount) : amount;
            todo.push_back({sn.lhs, true, lhs_weight, w.depth + 1});
            todo.push_back({sn.rhs, true, make_difference(doc, Ex::constant(1.0), lhs_weight), w.depth + 1});
            continue;
        }
        if (sn.kind == ir::S_ADD) {
            todo.push_back({sn.lhs, w.weighted, w.weight, w.depth + 1});
            todo.push_back({sn.rhs, w.weighted, w.weight, w.depth + 1});
            continue;
        }
        ComponentRec c;
        memset(&c, 0, sizeof(c));
        c.probability_program = w.weighted ? compile_program(doc, out, w.weight, false, ALLOW_RENDER) : -1;
        c.color_program = compile_program(doc, out, sn.color, false, ALLOW_RENDER);
        c.ior = 1.0f; c.env_ior = 1.0f;
        switch (sn.kind) {
            case ir::S_EMISSIVE: c.bsdf = BSDF_EMISSIVE; break;
            case ir::S_DIFFUSE: c.bsdf = BSDF_DIFFUSE; break;
            case ir::S_MIRROR: c.bsdf = BSDF_MIRROR; break;
            case ir::S_REFRACTIVE:
                c.bsdf = BSDF_REFRACTIVE;
                c.ior = fold.scalar(sn.ior);
                c.env_ior = sn.env_ior.some ? fold.scalar(sn.env_ior.ex) : 1.0f;
                c.dispersion = sn.dispersion.some ? fold.scalar(sn.dispersion.ex) : 0.0f;
                c.env_dispersion = sn.env_dispersion.some ? fold.scalar(sn.env_dispersion.ex) : 0.0f;
                break;
            default: throw BuildError("unknown surface material node");
        }
        all.push_back(c);
        if (sn.kind == ir::S_EMISSIVE) glowing.push_back(c);
    }
    MaterialRec m;
    memset(&m, 0, sizeof(m));
    m.comp_offset = (uint32_t)out.components.size();
    m.n_components = (uint32_t)all.size();
    for (auto& c : all) { c.selection_compensation = (float)all.size(); out.components.push_back(c); }
    m.emissive_offset = (uint32_t)out.components.size();
    m.n_emissive = (uint32_t)glowing.size();
    for (auto& c : glowing) { c.selection_compensation = (float)glowing.size(); out.components.push_back(c); }
    m.normal_map_program = use.normal_map.some ? compile_program(doc, out, use.normal_map.ex, true, ALLOW_NORMAL_MAP) : -1;
    out.materials.push_back(m);
    return (uint32_t)out.materials.size() - 1;
}

// ------------------------------------------------------------------ geometry being assembled
struct Corner { V3 p, n; Q4 frame; V2 t; };
struct Item {  // one BVH item = one entry of World::from_project's `objects`
    uint32_t kind = KIND_TRIANGLE, object_id = 0, material = 0;
    Corner c[3];                 // triangle
    V3 centre; float radius = 0; V2 tex_scale{1, 1};  // sphere (tex_scale defaults to (1, 1))
    int32_t marched = -1;        // ray-marched: index into BakedScene::marched
    Box box;
};

// Normal::transform (shapes/mod.rs:571-583)
void transform_corner_normal(Corner& c, const M4& t) {
    V3 n = unit(xform_vector(t, c.n));
    V3 x = unit(xform_vector(t, qrotate(c.frame, V3{1, 0, 0})));
    V3 y = unit(xform_vector(t, qrotate(c.frame, V3{0, 1, 0})));
    c.n = n;
    c.frame = quat_from_columns(x, y, n);
}

// make_triangle (world.rs:308-374)
Item triangle_item(const ir::MeshTable& mesh, const int32_t* k, uint32_t material) {
    auto position = [&](int32_t i) {
        if (i < 0 || (size_t)i * 3 + 2 >= mesh.positions.size()) throw BuildError("mesh position index out of range");
        return V3{mesh.positions[3 * i], mesh.positions[3 * i + 1], mesh.positions[3 * i + 2]};
    };
    auto normal = [&](int32_t i) {
        if ((size_t)i * 3 + 2 >= mesh.normals.size()) throw BuildError("mesh normal index out of range");
        return V3{mesh.normals[3 * i], mesh.normals[3 * i + 1], mesh.normals[3 * i + 2]};
    };
    auto uv = [&](int32_t i) {
        if (i < 0) return V2{0, 0};
        if ((size_t)i * 2 + 1 >= mesh.uvs.size()) throw BuildError("mesh texture index out of range");
        return V2{mesh.uvs[2 * i], mesh.uvs[2 * i + 1]};
    };
    Item it;
    it.kind = KIND_TRIANGLE;
    it.material = material;
    V3 p[3] = {position(k[0]), position(k[3]), position(k[6])};
    V3 n[3];
    if (k[2] >= 0 && k[5] >= 0 && k[8] >= 0) { n[0] = normal(k[2]); n[1] = normal(k[5]); n[2] = normal(k[8]); }
    else { n[0] = n[1] = n[2] = unit(cross3(sub(p[1], p[0]), sub(p[2], p[0]))); }
    V2 t[3] = {uv(k[1]), uv(k[4]), uv(k[7])};
    V3 dp1 = sub(p[1], p[0]), dp2 = sub(p[2], p[0]);
    V2 dt1{t[1].x - t[0].x, t[1].y - t[0].y}, dt2{t[2].x - t[0].x, t[2].y - t[0].y};
    float r = 1.0f / (dt1.x * dt2.y - dt1.y * dt2.x);
    V3 tangent = scale(sub(scale(dp1, dt2.y), scale(dp2, dt1.y)), r);
    V3 bitangent = scale(sub(scale(dp2, dt1.x), scale(dp1, dt2.x)), r);
    for (int i = 0; i < 3; ++i) it.c[i] = Corner{p[i], n[i], quat_from_columns(tangent, bitangent, n[i]), t[i]};
    return it;
}

Box item_box(const Item& it, const BakedScene& out) {  // Bounded for Shape (shapes/mod.rs:408-432)
    if (it.kind == KIND_TRIANGLE) return Box::of(it.c[0].p, it.c[1].p).with(it.c[2].p);
    if (it.kind == KIND_SPHERE) {
        float r = it.radius;
        return Box::of(V3{it.centre.x - r, it.centre.y - r, it.centre.z - r}, V3{it.centre.x + r, it.centre.y + r, it.centre.z + r});
    }
    const MarchedRec& m = out.marched[it.marched];
    V3 a{m.ba[0], m.ba[1], m.ba[2]}, b{m.bb[0], m.bb[1], m.bb[2]};
    if (m.bounds_type == 0) return Box::of(a, b);
    float r = m.bradius;
    return Box::of(V3{a.x - r, a.y - r, a.z - r}, V3{a.x + r, a.y + r, a.z + r});
}

// fn(from, to) over [0, n) on a few threads (sequentially for small n); the first exception is rethrown as a BuildError
template <class F> void parallel_ranges(size_t n, F&& fn) {
    unsigned threads = std::min(std::thread::hardware_concurrency(), 32u);
    if (const char* e = getenv("PYR_BUILD_THREADS")) threads = (unsigned)atoi(e);
    if (threads <= 1 || n < 32768) { fn(0, n); return; }
    std::string error;
    std::mutex error_mutex;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            try { fn(n * t / threads, n * (t + 1) / threads); }
            catch (const std::exception& e) { std::lock_guard<std::mutex> g(error_mutex); if (error.empty()) error = e.what(); }
        });
    for (auto& t : pool) t.join();
    if (!error.empty()) throw BuildError(error);
}

// ------------------------------------------------------------------ BVH (spatial/bvh.rs:13-155, 250-275, 318-370)
struct Hull2 {
    Box all, centres;
    static Hull2 around(const Box& b) { V3 c = b.middle(); return {b, Box::of(c, c)}; }
    Hull2 plus(const Box& b) const { return {all.merged(b), centres.with(b.middle())}; }
    Hull2 joined(const Hull2& o) const { return {all.merged(o.all), centres.merged(o.centres)}; }
};

struct TreeBuilder {
    const RawVector<Item>& items;
    BakedScene& out;
    std::vector<uint32_t> order;  // item index per rank
    struct Interior { Box box[2]; int32_t child[2]; };  // same layout as BvhInterior (bvh_build.hpp): the GPU builder's records are taken as they are
    static_assert(sizeof(Box) == 24, "Box is six floats");
    std::vector<Interior> interiors;
    int max_depth = 0;

    // Builds the subtree over `ids`, emitting leaves in the reference's flattened pre-order, and
    // returns its child code.  The reference's `first` child (visited first) is the subtree of the
    // SECOND item group: the Join pops it from the node stack first (bvh.rs:39-50).
    //
    // The tree is a pure function of the item list (every split only looks at its own items), so subtrees are independent:
    // the top of the tree is expanded here until the pending groups are small (`grain` items), those groups are built by a
    // pool of threads into private builders, and the pieces are stitched together in depth-first order - the leaf ranks
    // (the pre-order the tie rule of World::intersect is defined on) come out exactly as from a sequential build.
    struct Frame { std::vector<uint32_t> ids; Hull2 hull; int32_t parent; int slot; int depth; };
    int32_t run(std::vector<uint32_t> root_ids, Hull2 root_hull) {
        const size_t total = root_ids.size();
        unsigned threads = std::min(std::thread::hardware_concurrency(), 32u);
        if (const char* e = getenv("PYR_BUILD_THREADS")) threads = (unsigned)atoi(e);
        if (threads <= 1 || total < 65536) return run_sequential(Frame{std::move(root_ids), root_hull, -1, 0, 0}, true);
        // 1. the top of the tree, sequentially; groups of at most `grain` items become tasks (in depth-first order)
        const size_t grain = std::max<size_t>(total / (threads * 8), 1024);
        std::vector<Frame> tasks;
        std::vector<Frame> stack;
        stack.push_back(Frame{std::move(root_ids), root_hull, -1, 0, 0});
        int32_t root_code = 0;
        bool root_is_task = false;
        while (!stack.empty()) {
            Frame f = std::move(stack.back());
            stack.pop_back();
            if (f.ids.size() <= grain) {
                if (f.parent < 0) root_is_task = true;
                tasks.push_back(std::move(f));
                continue;
            }
            if (f.depth > max_depth) max_depth = f.depth;
            std::vector<uint32_t> group_a, group_b;
            Hull2 hull_a{}, hull_b{};
            split(f.ids, f.hull, group_a, hull_a, group_b, hull_b);
            const int32_t code = (int32_t)interiors.size();
            Interior in;
            in.box[0] = hull_b.all; in.box[1] = hull_a.all;
            in.child[0] = in.child[1] = 0;
            interiors.push_back(in);
            stack.push_back(Frame{std::move(group_a), hull_a, code, 1, f.depth + 1});
            stack.push_back(Frame{std::move(group_b), hull_b, code, 0, f.depth + 1});
            if (f.parent < 0) root_code = code; else interiors[f.parent].child[f.slot] = code;
        }
        // 2. the tasks, in parallel, each into its own builder (local leaf ranks and interior indices)
        std::vector<TreeBuilder> parts;
        parts.reserve(tasks.size());
        for (size_t i = 0; i < tasks.size(); ++i) parts.push_back(TreeBuilder{items, out, {}, {}});
        std::vector<int32_t> part_root(tasks.size(), 0);
        std::atomic<size_t> next{0};
        std::string error;
        std::mutex error_mutex;
        auto worker = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= tasks.size()) return;
                try {
                    Frame f = std::move(tasks[i]);
                    const int32_t parent = f.parent;
                    const int slot = f.slot;
                    f.parent = -1;
                    part_root[i] = parts[i].run_sequential(std::move(f), false);
                    tasks[i].parent = parent; tasks[i].slot = slot;
                } catch (const std::exception& e) {
                    std::lock_guard<std::mutex> g(error_mutex);
                    if (error.empty()) error = e.what();
                }
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
        if (!error.empty()) throw BuildError(error);
        // 3. stitch in depth-first (task) order: leaf ranks and interior indices shift by what came before
        for (size_t i = 0; i < tasks.size(); ++i) {
            TreeBuilder& p = parts[i];
            const int32_t rank_base = (int32_t)order.size(), interior_base = (int32_t)interiors.size();
            auto remap = [&](int32_t code) { return code < 0 ? ~(~code + rank_base) : code + interior_base; };
            order.insert(order.end(), p.order.begin(), p.order.end());
            for (Interior in : p.interiors) { in.child[0] = remap(in.child[0]); in.child[1] = remap(in.child[1]); interiors.push_back(in); }
            if (p.max_depth > max_depth) max_depth = p.max_depth;
            const int32_t code = remap(part_root[i]);
            if (tasks[i].parent < 0) root_code = code; else interiors[tasks[i].parent].child[tasks[i].slot] = code;
        }
        (void)root_is_task;
        return root_code;
    }
    int32_t run_sequential(Frame root, bool) {
        std::vector<Frame> stack;
        stack.push_back(std::move(root));
        int32_t root_code = 0;
        while (!stack.empty()) {
            Frame f = std::move(stack.back());
            stack.pop_back();
            if (f.depth > max_depth) max_depth = f.depth;
            int32_t code;
            if (f.ids.size() == 1) {
                code = ~(int32_t)order.size();
                order.push_back(f.ids[0]);
                // the leaf's box is the item's own box (hull.aabbs of a one-item hull)
            } else {
                std::vector<uint32_t> group_a, group_b;
                Hull2 hull_a{}, hull_b{};
                split(f.ids, f.hull, group_a, hull_a, group_b, hull_b);
                code = (int32_t)interiors.size();
                Interior in;
                in.box[0] = hull_b.all; in.box[1] = hull_a.all;
                in.child[0] = in.child[1] = 0;
                interiors.push_back(in);
                // depth-first: the B group (child 0) completely before the A group (child 1)
                stack.push_back(Frame{std::move(group_a), hull_a, code, 1, f.depth + 1});
                stack.push_back(Frame{std::move(group_b), hull_b, code, 0, f.depth + 1});
            }
            if (f.parent < 0) root_code = code; else interiors[f.parent].child[f.slot] = code;
        }
        return root_code;
    }

    void split(const std::vector<uint32_t>& ids, const Hull2& hull, std::vector<uint32_t>& group_a, Hull2& hull_a,
               std::vector<uint32_t>& group_b, Hull2& hull_b) const {
        V3 d = hull.centres.extent();
        float width; int axis;
        if (d.y > d.x) { width = d.y; axis = 1; } else { width = d.x; axis = 0; }
        if (d.z > width) { width = d.z; axis = 2; }
        if (width < DIST_EPSILON) {  // all centres coincide: halve the list
            size_t half = ids.size() / 2;
            group_a.assign(ids.begin(), ids.begin() + half);
            group_b.assign(ids.begin() + half, ids.end());
            hull_a = Hull2::around(items[group_a[0]].box);
            for (uint32_t i : group_a) hull_a = hull_a.plus(items[i].box);
            hull_b = Hull2::around(items[group_b[0]].box);
            for (uint32_t i : group_b) hull_b = hull_b.plus(items[i].box);
            return;
        }
        constexpr int SLOTS = 6;
        std::vector<uint32_t> members[SLOTS];
        Hull2 hulls[SLOTS];
        bool filled[SLOTS] = {};
        const float low = hull.centres.lo.at(axis);
        for (uint32_t i : ids) {
            const Box& b = items[i].box;
            float where = b.middle().at(axis);
            float fi = (float)SLOTS * (where - low) / width;
            size_t s = (size_t)std::min<uint64_t>(sat_usize(fi), SLOTS - 1);
            hulls[s] = filled[s] ? hulls[s].plus(b) : Hull2::around(b);
            filled[s] = true;
            members[s].push_back(i);
        }
        auto tally = [&](int from, int to, size_t& count, float& area) {
            count = 0;
            bool any = false;
            Box acc{};
            for (int s = from; s < to; ++s) {
                if (!filled[s]) continue;
                acc = any ? acc.merged(hulls[s].all) : hulls[s].all;
                any = true;
                count += members[s].size();
            }
            area = any ? acc.area() : 0.0f;
        };
        float best = std::numeric_limits<float>::infinity();
        int cut = 0;
        const float whole = hull.all.area();
        for (int s = 1; s < SLOTS; ++s) {
            size_t n1, n2; float a1, a2;
            tally(0, s, n1, a1);
            tally(s, SLOTS, n2, a2);
            float cost = (a1 * (float)n1 + a2 * (float)n2) / whole;
            if (cost < best) { best = cost; cut = s; }
        }
        auto gather = [&](int from, int to, std::vector<uint32_t>& ids_out, Hull2& hull_out) {
            bool any = false;
            for (int s = from; s < to; ++s) {
                if (!filled[s]) continue;
                hull_out = any ? hulls[s].joined(hull_out) : hulls[s];
                any = true;
                ids_out.insert(ids_out.end(), members[s].begin(), members[s].end());
            }
            if (!any) throw BuildError("BVH split produced an empty side");
        };
        gather(0, cut, group_a, hull_a);
        gather(cut, SLOTS, group_b, hull_b);
    }
};

inline uint32_t float_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

}  // namespace

std::vector<TileRec> make_tiles(uint32_t width, uint32_t height, uint32_t tile_size) {
    if (tile_size == 0) throw BuildError("tile_size must be positive");
    uint32_t nx = width / tile_size, ny = height / tile_size;
    if (nx * tile_size < width) ++nx;
    if (ny * tile_size < height) ++ny;
    const float fw = (float)width, fh = (float)height;
    const float longest = fmaxf(fw, fh);
    std::vector<TileRec> tiles;
    tiles.reserve((size_t)nx * ny);
    for (uint32_t y = 0; y < ny; ++y)
        for (uint32_t x = 0; x < nx; ++x) {
            uint32_t sx = x * tile_size, sy = y * tile_size;
            uint32_t w = std::min(width - sx, tile_size), h = std::min(height - sy, tile_size);
            TileRec t;
            t.from[0] = ((float)sx + (-fw * 0.5f)) / (longest * 0.5f);
            t.from[1] = ((float)sy + (-fh * 0.5f)) / (longest * 0.5f);
            t.size[0] = (float)w / (longest * 0.5f);
            t.size[1] = (float)h / (longest * 0.5f);
            t.width = w; t.height = h; t.index = y * nx + x; t.pad = 0;
            tiles.push_back(t);
        }
    return tiles;
}

namespace {
struct BuildTimer {  // PYR_BUILD_TIMING=1: phase times of the scene build on stderr
    bool on = getenv("PYR_BUILD_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[scene build] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
        t = now;
    }
};
}  // namespace

BakedScene build_scene(const Document& input, const BvhBuildFn* bvh_builder) {
    BuildTimer timer;
    Document doc = input;  // material flattening appends expression nodes
    timer.lap("copy of the document");
    BakedScene out;
    Folder fold{doc};
    SceneView& sv = out.view;
    memset(&sv, 0, sizeof(sv));

    // ---- constant tables, spectra, textures
    out.burns = doc.burns; out.xyz = doc.xyz; out.d65 = doc.d65;
    sv.burns_t = TableRec{doc.burns_lo, doc.burns_hi, (uint32_t)(doc.burns.size() / 3), 0};
    sv.xyz_t = TableRec{doc.xyz_lo, doc.xyz_hi, (uint32_t)(doc.xyz.size() / 3), 0};
    sv.d65_t = TableRec{doc.illum_lo, doc.illum_hi, (uint32_t)doc.d65.size(), 0};
    for (const auto& s : doc.spectra) {
        SpectrumRec r;
        memset(&r, 0, sizeof(r));
        r.is_curve = s.curve ? 1u : 0u;
        r.lo = s.lo; r.hi = s.hi;
        r.offset = (uint32_t)out.spectrum_data.size();
        r.n = (uint32_t)(s.curve ? s.values.size() / 2 : s.values.size());
        out.spectrum_data.insert(out.spectrum_data.end(), s.values.begin(), s.values.end());
        out.spectra.push_back(r);
    }
    auto add_texture = [&](const ir::TextureTable& t) {
        if (t.width == 0 || t.height == 0) throw BuildError("empty texture");
        TextureRec r;
        memset(&r, 0, sizeof(r));
        r.width = t.width; r.height = t.height; r.channels = t.channels;
        r.offset = out.texels.size();
        out.texels.insert(out.texels.end(), t.texels.begin(), t.texels.end());
        out.textures.push_back(r);
    };
    for (const auto& t : doc.color_textures) add_texture(t);
    for (const auto& t : doc.mono_textures) add_texture(t);
    sv.n_color_textures = (uint32_t)doc.color_textures.size();

    // ---- renderer (renderer/mod.rs:31-75)
    if (doc.renderer_kind > 1) throw BuildError("the photon mapping renderer is outside this library's scope");
    RendererRec& R = sv.renderer;
    R.algorithm = doc.renderer_kind;
    R.bounces = doc.bounces.unwrap_or(8);
    R.pixel_samples = doc.pixel_samples;
    R.light_samples = doc.light_samples.unwrap_or(4);
    R.spectrum_samples = doc.spectrum_samples.unwrap_or(10);
    R.spectrum_bins = doc.spectrum_resolution.unwrap_or(64);
    R.tile_size = doc.tile_size.unwrap_or(32);
    R.light_bounces = doc.light_bounces.unwrap_or(8);
    R.span_lo = 380.0f; R.span_hi = 780.0f;  // renderer/mod.rs:16
    R.width = doc.width; R.height = doc.height;
    if (doc.width == 0 || doc.height == 0) throw BuildError("image size must be positive");
    if (R.spectrum_samples == 0 || R.spectrum_samples > (uint32_t)MAX_SPECTRUM_SAMPLES)
        throw BuildError("spectrum_samples must be between 1 and 16");
    if (R.light_samples > (uint32_t)MAX_LIGHT_SAMPLES) throw BuildError("light_samples must be at most 8");
    if (R.spectrum_bins == 0) throw BuildError("spectrum_resolution must be positive");

    // ---- film (film.rs:21-45, 203-224)
    FilmRec& F = sv.film;
    F.width = doc.width; F.height = doc.height; F.bins = R.spectrum_bins;
    F.wavelength_start = R.span_lo;
    F.wavelength_width = R.span_hi - R.span_lo;
    F.grains_per_wavelength = (float)F.bins / F.wavelength_width;
    if (doc.width >= doc.height) { F.horizontal = 1; F.ar_size = (float)doc.width; F.ar_ratio = (float)doc.height / (float)doc.width; }
    else { F.horizontal = 0; F.ar_size = (float)doc.height; F.ar_ratio = (float)doc.width / (float)doc.height; }

    // ---- camera (cameras.rs:30-55, project/mod.rs:254-268)
    auto placement = [&](const ir::LookAtUse& l) {
        V3 from = fold.triple(l.from), to = fold.triple(l.to);
        V3 up = l.up.some ? fold.triple(l.up.ex) : V3{0, 1, 0};
        M4 inv;
        if (!invert4(look_at_rh(from, to, up), inv)) throw BuildError("could not invert view matrix");
        return inv;
    };
    {
        CameraRec& C = sv.camera;
        float fov = fold.scalar(doc.fov);
        float half = (fov * 0.5f) * (3.14159265358979323846f / 180.0f);
        C.view_plane = cosf(half) / sinf(half);
        M4 m = placement(doc.camera_transform);
        memcpy(C.m, m.e, sizeof(C.m));
        M4 back;
        C.inv_ok = invert4(m, back) ? 1u : 0u;
        if (C.inv_ok) memcpy(C.inv, back.e, sizeof(C.inv));
        C.focus_distance = doc.focus_distance.some ? fold.scalar(doc.focus_distance.ex) : 1.0f;
        C.aperture = doc.aperture.some ? fold.scalar(doc.aperture.ex) : 0.0f;
    }

    // ---- programs that do not belong to a material
    sv.sky_program = compile_program(doc, out, doc.sky.some ? doc.sky.ex : Ex::constant(0.0), false, ALLOW_RENDER);
    sv.filter_program = doc.filter.some ? compile_program(doc, out, doc.filter.ex, false, ALLOW_SPECTRUM_ONLY) : -1;
    sv.white_program = doc.white.some ? compile_program(doc, out, doc.white.ex, false, ALLOW_SPECTRUM_ONLY) : -1;

    // ---- world (world.rs:39-271)
    RawVector<Item> items;  // (resize() leaves new entries to the threads that fill them)
    struct LampSeed { bool from_item; uint32_t item; LampRec rec; };
    std::vector<LampSeed> lamp_seeds;
    const std::vector<ir::SceneObject> objects = doc.objects;
    for (size_t oi = 0; oi < objects.size(); ++oi) {
        const ir::SceneObject& o = objects[oi];
        switch (o.kind) {
            case ir::OBJ_SPHERE: {
                uint32_t mat = bake_material(doc, out, o.material);
                Item it;
                it.kind = KIND_SPHERE;
                if (o.texture_scale.some) { V4 ts = fold.quad(o.texture_scale.ex); it.tex_scale = V2{ts.x, ts.y}; }
                it.centre = fold.triple(o.a);
                it.radius = fold.scalar(o.b);
                it.material = mat;
                it.object_id = (uint32_t)items.size();
                if (out.materials[mat].n_emissive) lamp_seeds.push_back({true, it.object_id, LampRec{}});
                items.push_back(it);
                break;
            }
            case ir::OBJ_PLANE: {
                uint32_t mat = bake_material(doc, out, o.material);
                PlaneRec p;
                memset(&p, 0, sizeof(p));
                V3 n = unit(fold.triple(o.b));
                V3 binormal, tangent;
                basis(n, binormal, tangent);
                p.texture_scale[0] = p.texture_scale[1] = 1.0f;
                if (o.texture_scale.some) { V4 ts = fold.quad(o.texture_scale.ex); p.texture_scale[0] = ts.x; p.texture_scale[1] = ts.y; }
                V3 origin = fold.triple(o.a);
                p.n[0] = n.x; p.n[1] = n.y; p.n[2] = n.z;
                p.d = dot3(origin, n);  // collision::Plane::from_point_normal
                p.from_space = quat4(quat_from_columns(binormal, tangent, n));
                p.material = mat;
                out.planes.push_back(p);
                break;
            }
            case ir::OBJ_RAY_MARCHED: {
                uint32_t mat = bake_material(doc, out, o.material);
                MarchedRec m;
                memset(&m, 0, sizeof(m));
                m.bounds_type = o.bounds_kind;
                V3 a = fold.triple(o.bound_a);
                m.ba[0] = a.x; m.ba[1] = a.y; m.ba[2] = a.z;
                if (o.bounds_kind == 0) { V3 b = fold.triple(o.bound_b); m.bb[0] = b.x; m.bb[1] = b.y; m.bb[2] = b.z; }
                else m.bradius = fold.scalar(o.bound_b);
                m.estimator = o.estimator;
                {   // expressions.rs:301-304: literal f64 `as u16`
                    double it = o.iterations.is_node ? (double)fold.scalar(o.iterations) : o.iterations.number;
                    m.iterations = !(it > 0.0) ? 0u : (it >= 65535.0 ? 65535u : (uint32_t)it);
                }
                m.threshold = fold.scalar(o.threshold);
                if (o.estimator == 0) {
                    m.power = fold.scalar(o.power);
                    m.has_constant = o.bulb_constant.some ? 1u : 0u;
                    if (o.bulb_constant.some) { V3 c = fold.triple(o.bulb_constant.ex); m.mb_constant[0] = c.x; m.mb_constant[1] = c.y; m.mb_constant[2] = c.z; }
                } else {
                    V4 q = fold.quad(o.julia_constant);
                    m.constant = pack4(q.x, q.y, q.z, q.w);  // Quaternion::new(x, y, z, w) (expressions.rs:450-454)
                    m.slice_plane = fold.scalar(o.slice_plane);
                    m.variant = o.variant;
                }
                m.material = mat;
                m.object_id = (uint32_t)items.size();
                Item it;
                it.kind = KIND_RAY_MARCHED;
                it.material = mat;
                it.object_id = m.object_id;
                it.marched = (int32_t)out.marched.size();
                if (out.marched.size() >= (size_t)MAX_RAY_MARCHED) throw BuildError("too many ray-marched shapes (at most 32)");
                out.marched.push_back(m);
                items.push_back(it);
                break;
            }
            case ir::OBJ_MESH: {
                if (o.mesh >= doc.meshes.size()) throw BuildError("mesh id out of range");
                const ir::MeshTable& mesh = doc.meshes[o.mesh];
                auto remaining = o.mesh_materials;
                for (const auto& part : mesh.objects) {
                    auto found = std::find_if(remaining.begin(), remaining.end(), [&](const auto& kv) { return kv.first == part.name; });
                    if (found == remaining.end())
                        throw BuildError("objects[" + std::to_string(oi) + "]: missing material for '" + part.name + "'");
                    ir::MaterialUse use = found->second;
                    remaining.erase(found);
                    uint32_t mat = bake_material(doc, out, use);
                    const bool glows = out.materials[mat].n_emissive != 0;
                    const M4 place = o.has_transform ? placement(o.transform) : M4::identity();
                    const float factor = o.mesh_scale.some ? fold.scalar(o.mesh_scale.ex) : 1.0f;
                    const size_t count = part.corners.size() / 9;
                    const size_t first = items.size();
                    items.resize(first + count);
                    // every triangle is independent and its place in `items` is known: big meshes are converted by a few threads
                    parallel_ranges(count, [&](size_t from, size_t to) {
                        for (size_t t = from; t < to; ++t) {
                            Item it = triangle_item(mesh, &part.corners[9 * t], mat);
                            for (auto& c : it.c) c.p = scale(c.p, factor);                                  // Shape::scale (shapes/mod.rs:290-316)
                            for (auto& c : it.c) transform_corner_normal(c, place);                          // Shape::transform (:318-344)
                            for (auto& c : it.c) c.p = xform_point(place, c.p);
                            it.object_id = (uint32_t)(first + t);
                            items[first + t] = it;
                        }
                    });
                    if (glows) for (size_t t = 0; t < count; ++t) lamp_seeds.push_back({true, (uint32_t)(first + t), LampRec{}});
                }
                break;
            }
            case ir::OBJ_DIRECTIONAL_LIGHT: case ir::OBJ_POINT_LIGHT: {
                LampRec l;
                memset(&l, 0, sizeof(l));
                V3 v = fold.triple(o.a);
                l.v[0] = v.x; l.v[1] = v.y; l.v[2] = v.z;
                if (o.kind == ir::OBJ_DIRECTIONAL_LIGHT) { l.kind = LAMP_DIRECTIONAL; l.width = fold.scalar(o.b); }
                else l.kind = LAMP_POINT;
                l.color_program = compile_program(doc, out, o.c, false, ALLOW_RENDER);
                lamp_seeds.push_back({false, 0, l});
                break;
            }
            default: throw BuildError("unknown world object");
        }
    }
    out.n_objects = (uint32_t)items.size();
    parallel_ranges(items.size(), [&](size_t from, size_t to) { for (size_t i = from; i < to; ++i) items[i].box = item_box(items[i], out); });

    timer.lap("materials, items, boxes");
    // ---- BVH + rank order
    bool any_normal_map = false;
    for (const auto& m : out.materials) any_normal_map = any_normal_map || m.normal_map_program >= 0;
    TreeBuilder tb{items, out, {}, {}};
    BvhTree gpu_tree;  // holds the interior records when `bvh_builder` made the tree (they are read where they are)
    if (!items.empty()) {
        // the hull of all items: minima and maxima, so partial hulls of ranges combine exactly
        Hull2 hull = Hull2::around(items[0].box);
        {
            std::mutex hull_mutex;
            parallel_ranges(items.size(), [&](size_t from, size_t to) {
                if (from >= to) return;
                Hull2 part = Hull2::around(items[from].box);
                for (size_t i = from; i < to; ++i) part = part.plus(items[i].box);
                std::lock_guard<std::mutex> g(hull_mutex);
                hull = hull.joined(part);
            });
        }
        if (bvh_builder && *bvh_builder && items.size() >= 2) {
            // the level-synchronous build (bvh_build_core.hpp; on the GPU: bvh_build.cu) - the same tree as TreeBuilder's
            std::vector<float> boxes(items.size() * 6);
            parallel_ranges(items.size(), [&](size_t from, size_t to) {
                for (size_t i = from; i < to; ++i) {
                    const Box& b = items[i].box;
                    float* v = &boxes[6 * i];
                    v[0] = b.lo.x; v[1] = b.lo.y; v[2] = b.lo.z; v[3] = b.hi.x; v[4] = b.hi.y; v[5] = b.hi.z;
                }
            });
            const float hull12[12] = {hull.all.lo.x, hull.all.lo.y, hull.all.lo.z, hull.all.hi.x, hull.all.hi.y, hull.all.hi.z,
                                      hull.centres.lo.x, hull.centres.lo.y, hull.centres.lo.z, hull.centres.hi.x, hull.centres.hi.y, hull.centres.hi.z};
            timer.lap("  hull + box array");
            BvhTree tree;
            (*bvh_builder)(boxes.data(), items.size(), hull12, tree);
            timer.lap("  builder");
            if (tree.order.size() != items.size() || tree.n_interiors + 1 != items.size() || !tree.interiors) throw BuildError("the BVH builder returned a tree of the wrong size");
            tb.order = std::move(tree.order);
            gpu_tree = std::move(tree);
            tb.max_depth = tree.max_depth;
            sv.root = tree.root;
        } else {
            std::vector<uint32_t> all(items.size());
            for (uint32_t i = 0; i < all.size(); ++i) all[i] = i;
            sv.root = tb.run(std::move(all), hull);
        }
        // the traversal defers at most one child per level
        if (tb.max_depth >= BVH_STACK) throw BuildError("the BVH is deeper than the traversal stack (" + std::to_string(tb.max_depth) + " levels)");
        out.bvh_depth = (uint32_t)tb.max_depth;
        sv.root_lo[0] = hull.all.lo.x; sv.root_lo[1] = hull.all.lo.y; sv.root_lo[2] = hull.all.lo.z;
        sv.root_hi[0] = hull.all.hi.x; sv.root_hi[1] = hull.all.hi.y; sv.root_hi[2] = hull.all.hi.z;
    }
    // per-rank outputs: sized without zero-filling (RawVector), every record is written in full by the parallel loop below
    out.rank_of_object.assign(items.size(), 0);
    out.prims.resize(items.size());
    out.tri_shade.resize(items.size());
    if (any_normal_map) out.tri_frames.resize(items.size());
    timer.lap("BVH build");
    parallel_ranges(tb.order.size(), [&](size_t rank_from, size_t rank_to) {
    for (uint32_t rank = (uint32_t)rank_from; rank < (uint32_t)rank_to; ++rank) {
        const Item& it = items[tb.order[rank]];
        out.rank_of_object[it.object_id] = rank;
        Prim p;
        memset(&p, 0, sizeof(p));
        TriShade ts;
        memset(&ts, 0, sizeof(ts));
        if (it.kind == KIND_TRIANGLE) {
            V3 v1 = it.c[0].p, e1 = sub(it.c[1].p, it.c[0].p), e2 = sub(it.c[2].p, it.c[0].p);
            p.a = pack4(v1.x, v1.y, v1.z, e1.x);
            p.b = pack4(e1.y, e1.z, e2.x, e2.y);
            p.c.x = e2.z;
            for (int k = 0; k < 3; ++k) {
                float* n = k == 0 ? ts.n1 : (k == 1 ? ts.n2 : ts.n3);
                float* t = k == 0 ? ts.t1 : (k == 1 ? ts.t2 : ts.t3);
                n[0] = it.c[k].n.x; n[1] = it.c[k].n.y; n[2] = it.c[k].n.z;
                t[0] = it.c[k].t.x; t[1] = it.c[k].t.y;
            }
            ts.material = it.material;
            if (any_normal_map) out.tri_frames[rank] = TriFrames{quat4(it.c[0].frame), quat4(it.c[1].frame), quat4(it.c[2].frame)};
        } else if (it.kind == KIND_SPHERE) {
            p.a = pack4(it.centre.x, it.centre.y, it.centre.z, it.radius);
            p.b.x = it.tex_scale.x; p.b.y = it.tex_scale.y;
        } else {
            p.a.x = bits_to_float((uint32_t)it.marched);
            out.marched[it.marched].rank = rank;
        }
        p.c.y = bits_to_float(it.kind);
        p.c.z = bits_to_float(it.object_id);
        p.c.w = bits_to_float(it.material);
        out.prims[rank] = p;
        out.tri_shade[rank] = ts;
    }
    });
    timer.lap("primitive records");
    // fold the binary tree into 4-wide nodes: a Node4 per binary node that is the root or a grandchild-level entry
    static_assert(sizeof(TreeBuilder::Interior) == sizeof(BvhInterior) && sizeof(Box) == 24, "interior records of the two builders have one layout");
    const size_t n_interiors = gpu_tree.interiors ? gpu_tree.n_interiors : tb.interiors.size();
    const TreeBuilder::Interior* interiors = gpu_tree.interiors ? reinterpret_cast<const TreeBuilder::Interior*>(gpu_tree.interiors.get()) : tb.interiors.data();
    if (n_interiors) {
        // 1. which binary nodes become Node4s and under which index: a sequential walk (the numbering is the order of discovery, which is
        //    the nodes' order in memory), touching child codes only
        std::vector<int32_t> node4_of(n_interiors, -1);
        std::vector<uint32_t> todo;      // binary interior indices whose Node4 is still to be walked
        std::vector<uint32_t> binary_of; // Node4 index -> binary interior index
        binary_of.reserve(n_interiors / 2 + 1);
        auto node4_for = [&](uint32_t binary) {
            if (node4_of[binary] < 0) {
                node4_of[binary] = (int32_t)binary_of.size();
                binary_of.push_back(binary);
                todo.push_back(binary);
            }
            return node4_of[binary];
        };
        sv.root = node4_for((uint32_t)sv.root);  // the root is interior here (sv.root >= 0)
        while (!todo.empty()) {
            const uint32_t b = todo.back();
            todo.pop_back();
            for (int side = 0; side < 2; ++side) {
                const int32_t c = interiors[b].child[side];
                if (c < 0) continue;
                for (int g = 0; g < 2; ++g) {
                    const int32_t gc = interiors[c].child[g];
                    if (gc >= 0) node4_for((uint32_t)gc);
                }
            }
        }
        // 2. the records: every Node4 is a function of its binary node, its children and the numbering - filled by several threads
        out.nodes.resize(binary_of.size());
        parallel_ranges(binary_of.size(), [&](size_t from, size_t to) {
            for (size_t at = from; at < to; ++at) {
                const uint32_t b = binary_of[at];
                struct Entry { int32_t code; Box box; };
                Entry entries[4];
                int n = 0;
                for (int side = 0; side < 2; ++side) {
                    const int32_t c = interiors[b].child[side];
                    if (c < 0) { entries[n++] = Entry{c, interiors[b].box[side]}; continue; }
                    for (int g = 0; g < 2; ++g) {
                        const int32_t gc = interiors[c].child[g];
                        entries[n++] = Entry{gc < 0 ? gc : node4_of[gc], interiors[c].box[g]};
                    }
                }
                Node4 nd;
                memset(&nd, 0, sizeof(nd));
                const float inf = std::numeric_limits<float>::infinity();
                float* comps[6] = {&nd.lo_x.x, &nd.lo_y.x, &nd.lo_z.x, &nd.hi_x.x, &nd.hi_y.x, &nd.hi_z.x};
                for (int k = 0; k < 4; ++k) {
                    if (k < n) {
                        const Box& bx = entries[k].box;
                        const float v[6] = {bx.lo.x, bx.lo.y, bx.lo.z, bx.hi.x, bx.hi.y, bx.hi.z};
                        for (int c = 0; c < 6; ++c) comps[c][k] = v[c];
                        nd.child[k] = entries[k].code;
                    } else {
                        for (int c = 0; c < 6; ++c) comps[c][k] = c < 3 ? inf : -inf;  // (an empty slot is recognised by its child code)
                        nd.child[k] = NODE4_EMPTY;
                    }
                }
                out.nodes[at] = nd;
            }
        });
    }

    timer.lap("Node4 fold");
    // ---- lamps, in the order the reference collects them (world.rs:75-83, 250-262)
    for (const auto& seed : lamp_seeds) {
        LampRec l = seed.rec;
        if (seed.from_item) {
            memset(&l, 0, sizeof(l));
            l.kind = LAMP_SHAPE;
            l.color_program = -1;
            l.rank = out.rank_of_object[seed.item];
        }
        out.lamps.push_back(l);
    }

    out.tiles = make_tiles(doc.width, doc.height, R.tile_size);

    sv.n_nodes = (uint32_t)out.nodes.size();
    sv.n_prims = (uint32_t)out.prims.size();
    sv.n_planes = (uint32_t)out.planes.size();
    sv.n_marched = (uint32_t)out.marched.size();
    sv.n_lamps = (uint32_t)out.lamps.size();
    sv.n_tiles = (uint32_t)out.tiles.size();
    sv.vm_regs = 1;
    for (const ProgramRec& pr : out.programs)
        if (!pr.is_constant)
            for (uint32_t i = 0; i < pr.n_instr; ++i) sv.vm_regs = std::max<uint32_t>(sv.vm_regs, (uint32_t)out.code[pr.code_offset + i].out + 1);
    (void)float_bits;
    return out;
}

}  // namespace pyr
