// TEST INFRASTRUCTURE: runs the product's scene builder and the __host__ __device__ stage
// functions of pyrite_b200/csrc/{core,shading,bdpt}.cuh on the CPU, one path sample at a time, so
// that the host logic (IR decode, material flattening, bytecode compiler, BVH build, wavefront
// state machine) can be checked against the oracle in the GPU-less container.  Not shipped and not
// loaded by the product; built by tests/conftest.py into tests/_build/libhostemu.so.
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../pyrite_b200/csrc/bvh_build_core.hpp"
#include "../pyrite_b200/csrc/scene_build.hpp"
#include "../pyrite_b200/csrc/bdpt.cuh"

using namespace pyr;

namespace {
struct Emu {
    BakedScene scene;
    SceneView view;
    std::vector<float> film;
    uint64_t rays = 0, nodes = 0, leaves = 0;
};
thread_local std::string g_error;

// The level-synchronous BVH build of bvh_build_core.hpp, phase by phase like the kernels of bvh_build.cu (assign, split, rank,
// scatter) but with plain loops: the same per-item / per-node functions the GPU runs, checked here against the depth-first builder.
void level_sync_build(const float* boxes6, size_t n_items, const float* hull12, BvhTree& out) {
    using namespace bvhb;
    const uint32_t n = (uint32_t)n_items;
    std::vector<uint32_t> ids(n), ids_next(n), node_of(n, 0u), node_next(n), rank_at(n, 0u), pre(n);
    std::vector<uint8_t> bucket(n);
    for (uint32_t i = 0; i < n; ++i) ids[i] = i;
    std::vector<LevelNode> nodes(1), next;
    nodes[0].start = 0; nodes[0].count = n; nodes[0].rank_base = 0; nodes[0].interior = 0;
    for (int k = 0; k < 3; ++k) {
        nodes[0].hull.lo[k] = hull12[k]; nodes[0].hull.hi[k] = hull12[3 + k];
        nodes[0].hull.c_lo[k] = hull12[6 + k]; nodes[0].hull.c_hi[k] = hull12[9 + k];
    }
    out.allocate_interiors(n - 1);
    int level = 0;
    while (!nodes.empty()) {
        std::vector<BucketStats> stats(nodes.size());
        memset(stats.data(), 0, stats.size() * sizeof(BucketStats));
        for (uint32_t p = 0; p < n; ++p) {  // k_bvh_assign
            const uint32_t j = node_of[p];
            if (j == NO_NODE) { bucket[p] = (uint8_t)BUCKET_NONE; continue; }
            const float* box = boxes6 + 6 * (size_t)ids[p];
            const uint32_t b = item_bucket(nodes[j], p, box);
            uint32_t key[12];
            item_keys(box, key);
            bucket[p] = (uint8_t)b;
            stats[j].count[b] += 1;
            for (int k = 0; k < 12; ++k) stats[j].key[b][k] = std::max(stats[j].key[b][k], key[k]);
        }
        std::vector<Split> splits(nodes.size());
        next.clear();
        for (size_t j = 0; j < nodes.size(); ++j) {  // k_bvh_split
            const LevelNode& nd = nodes[j];
            const SplitChoice c = choose_split(nd, stats[j]);
            Split& sp = splits[j];
            for (int s = 0; s < BUCKETS; ++s) sp.offset[s] = c.offset[s];
            sp.cut = c.cut;
            if (c.n_a == 0 || c.n_b == 0) throw ir::BuildError("BVH split produced an empty side");
            LevelNode a, b;
            a.start = nd.start; a.count = c.n_a; a.rank_base = nd.rank_base + c.n_b; a.interior = nd.interior + c.n_b; a.hull = c.hull_a;
            b.start = nd.start + c.n_a; b.count = c.n_b; b.rank_base = nd.rank_base; b.interior = nd.interior + 1; b.hull = c.hull_b;
            BvhInterior in;
            for (int k = 0; k < 3; ++k) {
                in.box[0][k] = b.hull.lo[k]; in.box[0][3 + k] = b.hull.hi[k];
                in.box[1][k] = a.hull.lo[k]; in.box[1][3 + k] = a.hull.hi[k];
            }
            in.child[0] = b.count == 1 ? ~(int32_t)b.rank_base : (int32_t)b.interior;
            in.child[1] = a.count == 1 ? ~(int32_t)a.rank_base : (int32_t)a.interior;
            out.interiors[nd.interior] = in;
            sp.child[0] = sp.child[1] = NO_NODE;
            if (a.count == 1) rank_at[a.start] = a.rank_base; else { sp.child[0] = (uint32_t)next.size(); next.push_back(a); }
            if (b.count == 1) rank_at[b.start] = b.rank_base; else { sp.child[1] = (uint32_t)next.size(); next.push_back(b); }
        }
        uint32_t seen[BUCKETS] = {0, 0, 0, 0, 0, 0};  // k_bvh_tile_scan + k_bvh_rank: items of the same bucket at earlier positions
        std::vector<uint32_t> node_prefix(nodes.size() * 8, 0u);
        for (uint32_t p = 0; p < n; ++p) {
            const uint32_t b = bucket[p];
            if (b == BUCKET_NONE) continue;
            const uint32_t j = node_of[p];
            if (nodes[j].start == p) for (int k = 0; k < BUCKETS; ++k) node_prefix[j * 8 + k] = seen[k];
            pre[p] = seen[b]++;
        }
        for (uint32_t p = 0; p < n; ++p) {  // k_bvh_scatter
            const uint32_t b = bucket[p];
            if (b == BUCKET_NONE) { ids_next[p] = ids[p]; node_next[p] = NO_NODE; continue; }
            const uint32_t j = node_of[p];
            const uint32_t to = nodes[j].start + splits[j].offset[b] + (pre[p] - node_prefix[j * 8 + b]);
            ids_next[to] = ids[p];
            node_next[to] = splits[j].child[b < splits[j].cut ? 0 : 1];
        }
        ids.swap(ids_next); node_of.swap(node_next); nodes.swap(next);
        ++level;
    }
    out.order.resize(n);
    for (uint32_t p = 0; p < n; ++p) out.order[rank_at[p]] = ids[p];
    out.root = 0;
    out.max_depth = level;
}

struct HostAdd {
    float* film;
    void operator()(uint64_t index, float increment, float weight) const { film[2 * index] += increment; film[2 * index + 1] += weight; }
};
}  // namespace

extern "C" {

const char* emu_last_error() { return g_error.c_str(); }

// A digest of the built BVH (4-wide nodes with -0 box coordinates read as +0, leaf order), as pyr_bvh_digest computes it
void emu_bvh_digest(void* h, uint64_t* out2) {
    Emu* e = (Emu*)h;
    auto fnv = [](uint64_t x, uint32_t w) { for (int i = 0; i < 4; ++i) { x ^= (w >> (8 * i)) & 0xffu; x *= 1099511628211ull; } return x; };
    uint64_t h_nodes = 14695981039346656037ull, h_order = 14695981039346656037ull;
    for (const Node4& nd : e->scene.nodes) {
        uint32_t w[32];
        memcpy(w, &nd, sizeof(w));
        for (int i = 0; i < 24; ++i) if (w[i] == 0x80000000u) w[i] = 0;
        for (int i = 0; i < 28; ++i) h_nodes = fnv(h_nodes, w[i]);
    }
    std::vector<uint32_t> order(e->scene.n_objects);
    for (uint32_t obj = 0; obj < e->scene.n_objects; ++obj) order[e->scene.rank_of_object[obj]] = obj;
    for (uint32_t o : order) h_order = fnv(h_order, o);
    out2[0] = h_nodes; out2[1] = h_order;
}

int emu_load_with(const void* ir_blob, size_t bytes, int level_sync_bvh, void** out);
int emu_load(const void* ir_blob, size_t bytes, void** out) { return emu_load_with(ir_blob, bytes, 0, out); }
// level_sync_bvh != 0: the BVH comes from the level-synchronous build (the algorithm of the GPU builder) instead of the depth-first one
int emu_load_with(const void* ir_blob, size_t bytes, int level_sync_bvh, void** out) {
    try {
        auto e = std::make_unique<Emu>();
        const BvhBuildFn level_sync = level_sync_build;
        e->scene = build_scene(ir::decode(ir_blob, bytes), level_sync_bvh ? &level_sync : nullptr);
        BakedScene& b = e->scene;
        SceneView v = b.view;
        v.nodes = b.nodes.data(); v.prims = b.prims.data(); v.tri_shade = b.tri_shade.data(); v.tri_frames = b.tri_frames.data();
        v.planes = b.planes.data(); v.marched = b.marched.data(); v.materials = b.materials.data(); v.components = b.components.data();
        v.programs = b.programs.data(); v.code = b.code.data(); v.spectra = b.spectra.data(); v.spectrum_data = b.spectrum_data.data();
        v.textures = b.textures.data(); v.texels = b.texels.data(); v.lamps = b.lamps.data(); v.tiles = b.tiles.data();
        v.burns = b.burns.data(); v.xyz = b.xyz.data(); v.d65 = b.d65.data();
        e->view = v;
        e->film.assign((size_t)v.film.width * v.film.height * v.film.bins * 2, 0.0f);
        *out = e.release();
        return 0;
    } catch (const std::exception& ex) {
        g_error = ex.what();
        return 1;
    }
}
void emu_free(void* h) { delete (Emu*)h; }

// sizes: n_objects, n_planes, n_lamps, n_nodes, n_materials, n_programs, n_instr, n_tiles
void emu_info(void* h, uint32_t* out) {
    Emu* e = (Emu*)h;
    out[0] = e->scene.n_objects; out[1] = e->view.n_planes; out[2] = e->view.n_lamps; out[3] = e->view.n_nodes;
    out[4] = (uint32_t)e->scene.materials.size(); out[5] = (uint32_t)e->scene.programs.size(); out[6] = (uint32_t)e->scene.code.size();
    out[7] = e->view.n_tiles;
    out[8] = e->scene.bvh_depth;
}
void emu_leaf_order(void* h, uint32_t* object_ids) {
    Emu* e = (Emu*)h;
    for (uint32_t obj = 0; obj < e->scene.n_objects; ++obj) object_ids[e->scene.rank_of_object[obj]] = obj;
}

struct AbiRay { float o[3], pad0, d[3], pad1; };
struct AbiHit { uint32_t prim_id, kind; float t, u, v; };
void emu_trace(void* h, const AbiRay* rays, size_t n, AbiHit* hits, uint64_t* stats3) {
    Emu* e = (Emu*)h;
    TraceStats st{0, 0, 0, 0, 0};
    uint64_t nodes = 0, leaves = 0;
    for (size_t i = 0; i < n; ++i) {
        Ray r = make_ray(mk3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), mk3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), 0, 0.0f);
        Hit hit;
        st = TraceStats{0, 0, 0, 0, 0};
        trace_ray<true>(e->view, r, hit, &st);
        nodes += st.nodes; leaves += st.leaves;
        AbiHit o;
        o.kind = hit.kind; o.t = hit.t; o.u = hit.u; o.v = hit.v;
        o.prim_id = hit.kind == KIND_MISS ? 0xFFFFFFFFu : (hit.kind == KIND_PLANE ? hit.rank : prim_object(e->view.prims[hit.rank]));
        hits[i] = o;
    }
    if (stats3) { stats3[0] = n; stats3[1] = nodes; stats3[2] = leaves; }
}

// The wavefront state machine, run depth-first for one path sample at a time.
int emu_render(void* h, uint64_t seed, uint32_t spp_override, uint32_t offset, uint32_t stride, int reset) {
    Emu* e = (Emu*)h;
    try {
        const SceneView& sc = e->view;
        if (reset) std::fill(e->film.begin(), e->film.end(), 0.0f);
        HostAdd add{e->film.data()};
        const uint32_t spp = spp_override ? spp_override : sc.renderer.pixel_samples;
        if (stride == 0) stride = 1;
        std::vector<Ray> rays(1 + BDPT_STAGE);
        std::vector<Hit> hits(rays.size());
        std::vector<uint32_t> kinds(rays.size());
        auto ps = std::make_unique<PathState>();
        std::vector<float> spectral(3 * MAX_SPECTRUM_SAMPLES);
        ps->wl.base = spectral.data();
        ps->bright.base = spectral.data() + MAX_SPECTRUM_SAMPLES;
        ps->refl.base = spectral.data() + 2 * MAX_SPECTRUM_SAMPLES;
        std::vector<PendingLight> pend(MAX_LIGHT_SAMPLES);
        BidirState bd{};
        ps->pend = pend.data();
        ps->bd = &bd;
        std::vector<LightVertex> lv(sc.renderer.light_bounces + 1);
        std::vector<CamVertex> cv(sc.renderer.bounces > 0 ? sc.renderer.bounces : 1);
        std::vector<float> scratch(2 * MAX_SPECTRUM_SAMPLES);
        BidirCtx cx{lv.data(), cv.data(), {scratch.data()}, {scratch.data() + MAX_SPECTRUM_SAMPLES}};
        for (uint32_t t = 0; t < sc.n_tiles; ++t) {
            const uint64_t iterations = (uint64_t)sc.tiles[t].width * sc.tiles[t].height * spp;
            for (uint64_t i = offset; i < iterations; i += stride) {
                ShadeOut out;
                out.stage_base = 0;
                PathCounters pc{0, 0};
                if (sc.renderer.algorithm == 0) {
                    generate_simple(sc, seed, t, i, *ps, out.main);
                    out.has_main = 1; out.n_shadow = 0; out.alive = 1;
                    while (out.alive) {
                        uint32_t n = 0;
                        if (out.has_main) rays[n++] = out.main; else n = 1;
                        for (uint32_t j = 0; j < out.n_shadow; ++j) rays[n++] = out.get_shadow(j);
                        for (uint32_t j = out.has_main ? 0 : 1; j < n; ++j) { trace_ray<false>(sc, rays[j], hits[j], nullptr); kinds[j] = hits[j].kind; ++e->rays; }
                        shade_simple(sc, *ps, rays.data(), hits.data(), rays.data() + 1, kinds.data() + 1, out, add, pc);
                    }
                } else {
                    BidirOut bo;
                    generate_bidirectional(sc, seed, t, i, *ps, cx, bo);
                    while (bo.alive) {
                        if (bo.has_main) { rays[0] = bo.main; trace_ray<false>(sc, rays[0], hits[0], nullptr); ++e->rays; }
                        for (uint32_t j = 0; j < bo.n_shadow; ++j) { rays[1 + j] = bo.shadow[j]; trace_ray<false>(sc, rays[1 + j], hits[1 + j], nullptr); kinds[1 + j] = hits[1 + j].kind; ++e->rays; }
                        shade_bidirectional(sc, *ps, cx, rays.data(), hits.data(), rays.data() + 1, kinds.data() + 1, bo, add, pc);
                    }
                }
            }
        }
        return 0;
    } catch (const std::exception& ex) {
        g_error = ex.what();
        return 1;
    }
}
// tools/first_divergence.py --backend emu: the product's depth-first diagnostic (shading.cuh debug_path_simple) on the CPU
void emu_debug_path(void* h, uint64_t seed, uint32_t tile, uint64_t sample, uint32_t max_bounces, uint32_t* records, uint32_t* n_bounces,
                    float* exposed, uint32_t* n_exposed, float* position2) {
    Emu* e = (Emu*)h;
    std::vector<Ray> rays(1 + MAX_LIGHT_SAMPLES);
    std::vector<Hit> hits(rays.size());
    std::vector<uint32_t> kinds(rays.size());
    std::vector<PendingLight> pend(MAX_LIGHT_SAMPLES);
    std::vector<float> spectral(3 * MAX_SPECTRUM_SAMPLES);
    PathState ps;
    ps.wl.base = spectral.data(); ps.bright.base = spectral.data() + MAX_SPECTRUM_SAMPLES; ps.refl.base = spectral.data() + 2 * MAX_SPECTRUM_SAMPLES;
    ps.pend = pend.data();
    ps.bd = nullptr;
    ShadeOut out;
    out.stage_base = 0;
    uint32_t counts[2] = {0, 0};
    debug_path_simple(e->view, seed, tile, sample, max_bounces, records, counts, exposed, position2, rays.data(), hits.data(), kinds.data(), ps, out);
    *n_bounces = counts[0]; *n_exposed = counts[1];
}

uint64_t emu_rays(void* h) { return ((Emu*)h)->rays; }
void emu_film(void* h, float* out) { Emu* e = (Emu*)h; memcpy(out, e->film.data(), e->film.size() * sizeof(float)); }
void emu_set_film(void* h, const float* in) { Emu* e = (Emu*)h; memcpy(e->film.data(), in, e->film.size() * sizeof(float)); }

void emu_expose(void* h, const float* positions, const float* samples, size_t n) {
    Emu* e = (Emu*)h;
    HostAdd add{e->film.data()};
    for (size_t i = 0; i < n; ++i)
        film_expose(e->view.film, positions[2 * i], positions[2 * i + 1], samples[3 * i], samples[3 * i + 1], samples[3 * i + 2], add);
}

void emu_develop(void* h, float step_size, float* xyz_out, uint8_t* srgb_out) {
    Emu* e = (Emu*)h;
    const SceneView& sc = e->view;
    DevelopParams dp{0, 0, step_size, 0};
    white_scan(sc, dp.white_max, dp.d65_max);
    const uint64_t pixels = (uint64_t)sc.film.width * sc.film.height;
    for (uint64_t p = 0; p < pixels; ++p) {
        float xyz[3] = {0, 0, 0};
        uint8_t rgb[3] = {0, 0, 0};
        if (p + 1 < pixels) { pixel_to_xyz(sc, dp, e->film.data(), p, xyz); xyz_to_srgb8(xyz, rgb); }
        if (xyz_out) for (int c = 0; c < 3; ++c) xyz_out[3 * p + c] = xyz[c];
        if (srgb_out) for (int c = 0; c < 3; ++c) srgb_out[3 * p + c] = rgb[c];
    }
}

// evaluate program `index` (ExecutionContext::run): inputs = wavelength, normal[3], incident[3], tex[2]; out[4]
void emu_run_program(void* h, int32_t index, const float* inputs, float* out4) {
    Emu* e = (Emu*)h;
    VmInputs in;
    in.wavelength = inputs[0];
    in.normal = mk3(inputs[1], inputs[2], inputs[3]);
    in.incident = mk3(inputs[4], inputs[5], inputs[6]);
    in.tex[0] = inputs[7]; in.tex[1] = inputs[8];
    PYR_REGFILE(R);
    f4 v = run_vector(e->view, index, in, R);
    const ProgramRec p = e->view.programs[index];
    if (p.is_constant) { out4[0] = p.value; out4[1] = out4[2] = out4[3] = p.value; return; }
    out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
}

void emu_camera_sample(void* h, uint64_t seed, uint32_t tile, uint64_t sample, float* position2, AbiRay* ray, float* wavelengths, uint32_t* hero) {
    Emu* e = (Emu*)h;
    const SceneView& sc = e->view;
    const TileRec t = sc.tiles[tile];
    Rng rng = keyed_rng(seed, t.index, sample);
    float ox = t.size[0] * rng.gen_f32();
    float oy = t.size[1] * rng.gen_f32();
    position2[0] = t.from[0] + ox; position2[1] = t.from[1] + oy;
    v3 o, d;
    if (sc.renderer.algorithm == 0) { camera_ray(sc.camera, position2[0], position2[1], rng, o, d); *hero = sample_wavelengths(sc, rng, wavelengths); }
    else { *hero = sample_wavelengths(sc, rng, wavelengths); camera_ray(sc.camera, position2[0], position2[1], rng, o, d); }
    *ray = AbiRay{{o.x, o.y, o.z}, 0.0f, {d.x, d.y, d.z}, 0.0f};
}

}  // extern "C"
