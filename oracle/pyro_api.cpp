// ORACLE - TEST INFRASTRUCTURE ONLY (see pyro_math.hpp header).  PARITY UNPINNED.
//
// C entry points of the CPU oracle (`liboracle.so`), loaded with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs - never by
// the product (pyrite_b200/).  The functions mirror include/pyrite_b200.h one for one
// (pyr_* -> pyro_*) so the parity tests drive both libraries with the same calls.
#include <chrono>
#include <cstdio>
#include <string>

#include "pyro_render.hpp"

using namespace pyro;

namespace {
struct Handle {
    std::unique_ptr<World> world;
    Camera camera;
    RendererParams R;
    std::unique_ptr<Film> film;
    RenderCounters counters;
    double last_render_seconds = 0;
};
thread_local std::string g_error;
}  // namespace

extern "C" {

struct pyro_ray { float o[3]; float pad0; float d[3]; float pad1; };
struct pyro_hit { uint32_t prim_id; uint32_t kind; float t; float u; float v; };
struct pyro_info_t {
    uint32_t width, height, bins, algorithm;
    uint32_t pixel_samples, bounces, light_samples, spectrum_samples, light_bounces, tile_size;
    uint32_t n_objects, n_planes, n_lights, n_bvh_nodes, n_materials, threads;
};
struct pyro_counters_t { uint64_t rays, nodes, leaves, path_samples, de_evals, de_iters; };
struct pyro_render_opts {
    uint64_t seed;
    int32_t rng_mode, eager_emissive_draw;
    uint32_t spp_override, sample_offset, sample_stride;
    int32_t threads, cas_attempts, reset_film;
};

const char* pyro_last_error() { return g_error.c_str(); }
int pyro_libm_mode() { return LIBM_MODE; }  // 0 = glibc float functions (the reference), 1 = double, rounded once (the device's expression)

int pyro_load(const void* ir, size_t bytes, void** out) {
    try {
        Project P = parse_project_ir(ir, bytes);
        auto h = std::make_unique<Handle>();
        h->camera = camera_from_project(P);
        h->R = RendererParams::from_project(P);
        h->world = world_from_project(std::move(P));
        const Project& Q = h->world->P;
        h->film = std::make_unique<Film>(Q.width, Q.height, h->R.spectrum_bins, h->R.span_lo, h->R.span_hi);
        *out = h.release();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

void pyro_free(void* handle) { delete (Handle*)handle; }

int pyro_info(void* handle, pyro_info_t* o) {
    Handle* h = (Handle*)handle;
    const Project& P = h->world->P;
    *o = pyro_info_t{P.width, P.height, h->R.spectrum_bins, h->R.algorithm, h->R.pixel_samples, h->R.bounces, h->R.light_samples,
                     h->R.spectrum_samples, h->R.light_bounces, h->R.tile_size, (uint32_t)h->world->objects.size(),
                     (uint32_t)h->world->planes.size(), (uint32_t)h->world->lights.size(), (uint32_t)h->world->bvh.nodes.size(),
                     (uint32_t)h->world->materials.size(), h->R.threads};
    return 0;
}

// World::intersect on a batch (world.rs:273-299).  prim_id: planes and BVH objects have separate
// index spaces; `kind` tells them apart.
int pyro_trace(void* handle, const pyro_ray* rays, size_t n, pyro_hit* hits, int threads, pyro_counters_t* counters) {
    Handle* h = (Handle*)handle;
    try {
        const World& W = *h->world;
        if (threads < 1) threads = 1;
        std::vector<TraceCounters> tcs(threads);
        auto work = [&](int t) {
            size_t lo = n * t / threads, hi = n * (t + 1) / threads;
            for (size_t i = lo; i < hi; ++i) {
                Ray r{{rays[i].o[0], rays[i].o[1], rays[i].o[2]}, {rays[i].d[0], rays[i].d[1], rays[i].d[2]}};
                Intersection isect;
                pyro_hit out{0xFFFFFFFFu, K_MISS, INF, 0, 0};
                if (W.intersect(r, isect, &tcs[t])) {
                    const SurfacePoint& sp = isect.surface_point;
                    out.kind = sp.kind;
                    out.prim_id = sp.kind == K_PLANE ? sp.plane->id : sp.shape->id;
                    out.t = isect.distance;
                    out.u = sp.u;
                    out.v = sp.v;
                }
                hits[i] = out;
            }
        };
        if (threads == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
            for (auto& t : pool) t.join();
        }
        if (counters) {
            *counters = pyro_counters_t{};
            for (auto& tc : tcs) { counters->rays += tc.rays; counters->nodes += tc.nodes; counters->leaves += tc.leaves; }
        }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// Parity ray batches: kind 0 = primary camera rays, 1 = uniform-hemisphere
// secondary rays leaving the primary hits, 2 = shadow rays from the primary hits towards lamp
// samples.  Rays whose primary ray misses are regenerated, so every batch has exactly n rays.
int pyro_gen_rays(void* handle, int kind, size_t n, uint64_t seed, pyro_ray* out) {
    Handle* h = (Handle*)handle;
    try {
        const World& W = *h->world;
        const Project& P = W.P;
        std::vector<Tile> tiles = make_tiles(P.width, P.height, h->R.tile_size);
        XorShift rng{0x193a6754u, 0xa8a7d469u, 0x97830e05u, 0x113ba7bbu};
        rng.x ^= (uint32_t)seed; rng.y ^= (uint32_t)(seed >> 32); rng.z += (uint32_t)kind;
        for (int i = 0; i < 16; ++i) rng.next_u32();
        size_t produced = 0, attempts = 0;
        while (produced < n) {
            if (++attempts > 64 * n + 1024) throw std::runtime_error("could not generate enough rays (camera sees nothing?)");
            const Tile& tile = tiles[rng.gen_range_usize(tiles.size())];
            Vec2 position = tile.sample_point(rng);
            Ray ray = h->camera.ray_towards(position, rng);
            if (kind != 0) {
                Intersection isect;
                if (!W.intersect(ray, isect)) continue;
                SurfaceData sd = W.surface_data(isect.surface_point);
                Vec3 normal = sd.normal.vector;
                if (dot(ray.direction, normal) >= 0.0f) normal = -normal;
                if (kind == 1) {
                    ray = Ray{isect.surface_point.position, sample_hemisphere(rng, normal)};
                } else {
                    if (W.lights.empty()) throw std::runtime_error("scene has no lamps for shadow rays");
                    float p;
                    const Lamp* lamp = W.pick_lamp(rng, p);
                    LampSample ls = lamp_sample(W, *lamp, rng, isect.surface_point.position);
                    ray = Ray{isect.surface_point.position, ls.direction};
                }
                if (!(ray.direction.x == ray.direction.x)) continue;
            }
            out[produced++] = pyro_ray{{ray.origin.x, ray.origin.y, ray.origin.z}, 0.0f, {ray.direction.x, ray.direction.y, ray.direction.z}, 0.0f};
        }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// Renderer::render (renderer/mod.rs:77-111) into the handle's film.
int pyro_render(void* handle, const pyro_render_opts* o) {
    Handle* h = (Handle*)handle;
    try {
        const Project& P = h->world->P;
        if (o->reset_film) h->film = std::make_unique<Film>(P.width, P.height, h->R.spectrum_bins, h->R.span_lo, h->R.span_hi);
        RenderOptions opt;
        opt.seed = o->seed; opt.rng_mode = o->rng_mode; opt.eager_emissive_draw = o->eager_emissive_draw != 0;
        opt.spp_override = o->spp_override; opt.sample_offset = o->sample_offset; opt.sample_stride = o->sample_stride ? o->sample_stride : 1;
        opt.threads = o->threads; opt.cas_attempts = o->cas_attempts;
        RenderState st{*h->world, h->camera, h->R, *h->film, opt, h->counters};
        auto t0 = std::chrono::steady_clock::now();
        render(st);
        h->last_render_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// Exactly ONE path sample (tile, sample) of the project's integrator into the film, on its keyed stream: replays a sample the
// product flagged (tools/find_nonfinite.py).
int pyro_render_sample(void* handle, uint64_t seed, uint32_t tile, uint32_t sample, int reset_film) {
    Handle* h = (Handle*)handle;
    try {
        const Project& P = h->world->P;
        if (reset_film) h->film = std::make_unique<Film>(P.width, P.height, h->R.spectrum_bins, h->R.span_lo, h->R.span_hi);
        RenderOptions opt;
        opt.seed = seed; opt.rng_mode = 1; opt.eager_emissive_draw = true;
        opt.spp_override = 0xFFFFFFFFu;   // `sample` < area * spp for any sample index
        opt.sample_offset = sample; opt.sample_stride = 1; opt.sample_limit = 1; opt.threads = 1; opt.cas_attempts = 0; opt.only_tile = tile;
        RenderState st{*h->world, h->camera, h->R, *h->film, opt, h->counters};
        render(st);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

double pyro_last_render_seconds(void* handle) { return ((Handle*)handle)->last_render_seconds; }

int pyro_counters(void* handle, pyro_counters_t* c, int reset) {
    Handle* h = (Handle*)handle;
    *c = pyro_counters_t{h->counters.rays.load(), h->counters.nodes.load(), h->counters.leaves.load(),
                         h->counters.path_samples.load(), h->counters.de_evals.load(), h->counters.de_iters.load()};
    if (reset) { h->counters.rays = 0; h->counters.nodes = 0; h->counters.leaves = 0; h->counters.path_samples = 0; h->counters.de_evals = 0; h->counters.de_iters = 0; }
    return 0;
}

// film as W*H*bins*(accumulator, weight) f32 pairs
int pyro_film_download(void* handle, float* out) {
    Handle* h = (Handle*)handle;
    size_t n = h->film->width * h->film->height * h->film->grains_per_pixel;
    for (size_t i = 0; i < n; ++i) Film::unpack(h->film->grains[i].load(std::memory_order_relaxed), out[2 * i], out[2 * i + 1]);
    return 0;
}
int pyro_film_upload(void* handle, const float* in) {
    Handle* h = (Handle*)handle;
    size_t n = h->film->width * h->film->height * h->film->grains_per_pixel;
    for (size_t i = 0; i < n; ++i) h->film->grains[i].store(Film::pack(in[2 * i], in[2 * i + 1]), std::memory_order_relaxed);
    return 0;
}

// Film::expose on a batch (film.rs:89-95): positions (x,y) view coords, samples (brightness, wavelength, weight)
int pyro_film_expose(void* handle, const float* positions, const float* samples, size_t n) {
    Handle* h = (Handle*)handle;
    try {
        h->film->cas_attempts = 0;
        for (size_t i = 0; i < n; ++i)
            h->film->expose(Vec2{positions[2 * i], positions[2 * i + 1]}, Sample{samples[3 * i], samples[3 * i + 1], samples[3 * i + 2]});
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// main.rs:313-327: XYZ (3 f32 / pixel) and sRGB8 (3 u8 / pixel); the last pixel is left black (film.rs:299)
int pyro_film_develop(void* handle, float step_size, float* xyz_out, uint8_t* srgb_out, int threads) {
    Handle* h = (Handle*)handle;
    try {
        const Film& film = *h->film;
        Developer dev(*h->world, h->R.span_lo, h->R.span_hi);
        size_t pixels = film.width * film.height;
        if (threads < 1) threads = 1;
        auto work = [&](int t) {
            for (size_t p = pixels * t / threads; p < pixels * (t + 1) / threads; ++p) {
                float xyz[3] = {0, 0, 0};
                uint8_t rgb[3] = {0, 0, 0};
                bool yielded = (p + 1) * film.grains_per_pixel < pixels * film.grains_per_pixel;  // `end < len`
                if (yielded) { pixel_to_xyz(film, dev, p, step_size, xyz); xyz_to_srgb8(xyz, rgb); }
                if (xyz_out) { xyz_out[3 * p] = xyz[0]; xyz_out[3 * p + 1] = xyz[1]; xyz_out[3 * p + 2] = xyz[2]; }
                if (srgb_out) { srgb_out[3 * p] = rgb[0]; srgb_out[3 * p + 1] = rgb[1]; srgb_out[3 * p + 2] = rgb[2]; }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (auto& t : pool) t.join();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// tools/first_divergence.py: one path sample of the camera-to-light integrator with per-bounce records.
// records: max_bounces x 20 words {kind, prim_id, t, u, v, incident[3], position[3], normal[3], out[3], n_direct, rng_w, pad};
// exposed: up to 16 x (brightness, wavelength); position2: film position.  The layout equals pyr_debug_path's.
int pyro_debug_path(void* handle, uint64_t seed, int eager_emissive_draw, uint32_t tile_index, uint64_t sample, uint32_t max_bounces,
                    uint32_t* records, uint32_t* n_bounces, float* exposed, uint32_t* n_exposed, float* position2) {
    Handle* h = (Handle*)handle;
    try {
        const Project& P = h->world->P;
        std::vector<Tile> tiles = make_tiles(P.width, P.height, h->R.tile_size);
        const Tile* tile = nullptr;
        for (auto& t : tiles) if (t.index == tile_index) tile = &t;
        if (!tile) throw std::runtime_error("tile index out of range");
        RenderOptions opt;
        opt.seed = seed; opt.eager_emissive_draw = eager_emissive_draw != 0;
        std::vector<DebugBounce> bounces;
        std::vector<Sample> samples;
        Vec2 pos;
        debug_path_simple(*h->world, h->camera, h->R, *h->film, opt, *tile, sample, bounces, pos, samples);
        static_assert(sizeof(DebugBounce) == 20 * 4, "20 words per record");
        *n_bounces = (uint32_t)bounces.size();
        for (size_t b = 0; b < bounces.size() && b < max_bounces; ++b) memcpy(records + 20 * b, &bounces[b], sizeof(DebugBounce));
        *n_exposed = (uint32_t)samples.size();
        for (size_t k = 0; k < samples.size() && k < 16; ++k) { exposed[2 * k] = samples[k].brightness; exposed[2 * k + 1] = samples[k].wavelength; }
        position2[0] = pos.x; position2[1] = pos.y;
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// leaf pre-order rank of every BVH object (tie rule of World::intersect)
int pyro_bvh_leaf_order(void* handle, uint32_t* object_ids_in_preorder) {
    Handle* h = (Handle*)handle;
    size_t k = 0;
    for (auto& n : h->world->bvh.nodes)
        if (n.item) object_ids_in_preorder[k++] = n.item->id;
    return 0;
}

// camera seam: Tile::sample_point + Camera::ray_towards + wavelengths + hero pick for path sample (tile, i)
int pyro_camera_sample(void* handle, uint64_t seed, uint32_t tile_index, uint64_t sample, float* position2, pyro_ray* ray,
                       float* wavelengths, uint32_t* hero) {
    Handle* h = (Handle*)handle;
    try {
        const Project& P = h->world->P;
        std::vector<Tile> tiles = make_tiles(P.width, P.height, h->R.tile_size);
        const Tile* tile = nullptr;
        for (auto& t : tiles) if (t.index == tile_index) tile = &t;
        if (!tile) throw std::runtime_error("tile index out of range");
        XorShift rng = keyed_rng(seed, tile_index, sample);
        Vec2 pos = tile->sample_point(rng);
        Ray r;
        std::vector<float> wl;
        if (h->R.algorithm == 0) { r = h->camera.ray_towards(pos, rng); h->film->sample_many_wavelengths(rng, h->R.spectrum_samples, wl); }
        else { h->film->sample_many_wavelengths(rng, h->R.spectrum_samples, wl); }
        size_t pick = rng.gen_range_usize(wl.size());
        if (h->R.algorithm != 0) r = h->camera.ray_towards(pos, rng);
        position2[0] = pos.x; position2[1] = pos.y;
        *ray = pyro_ray{{r.origin.x, r.origin.y, r.origin.z}, 0.0f, {r.direction.x, r.direction.y, r.direction.z}, 0.0f};
        for (size_t i = 0; i < wl.size(); ++i) wavelengths[i] = wl[i];
        *hero = (uint32_t)pick;
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

}  // extern "C"
