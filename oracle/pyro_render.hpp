// ORACLE - TEST INFRASTRUCTURE ONLY (see pyro_math.hpp header).  PARITY UNPINNED.
//
// The reference's render inner loop restated on the CPU:
//   film                    film.rs
//   tracer                  tracer.rs (trace / trace_direct / trace_directional)
//   BSDFs                   materials/{diffuse,mirror,refractive}.rs
//   integrators             renderer/{algorithm,simple,bidirectional,mod}.rs, utils.rs:5-13
//   develop                 main.rs:190-238, 313-418
#pragma once
#include <atomic>
#include <mutex>
#include <thread>

#include "pyro_scene.hpp"

namespace pyro {

// ---------------------------------------------------------------- film (film.rs)
struct Sample { float brightness = 0, wavelength = 0, weight = 0; };  // film.rs:275-280

struct Film {
    size_t width = 0, height = 0, grains_per_pixel = 0;
    float wavelength_start = 0, wavelength_width = 0, grains_per_wavelength = 0;
    // AspectRatio (film.rs:203-224)
    float ar_size = 0, ar_ratio = 0;
    bool horizontal = true;
    int cas_attempts = 5;  // film.rs:149-161 drops the sample after 5 failed CAS; <=0 means never drop
    std::unique_ptr<std::atomic<uint64_t>[]> grains;  // {accumulator f32, weight f32} packed

    Film(size_t w, size_t h, size_t bins, float span_lo, float span_hi) : width(w), height(h), grains_per_pixel(bins) {
        wavelength_start = span_lo;
        wavelength_width = span_hi - span_lo;
        grains_per_wavelength = (float)bins / wavelength_width;
        if (w >= h) { ar_size = (float)w; ar_ratio = (float)h / (float)w; horizontal = true; }
        else { ar_size = (float)h; ar_ratio = (float)w / (float)h; horizontal = false; }
        size_t n = w * h * bins;
        grains.reset(new std::atomic<uint64_t>[n]);
        for (size_t i = 0; i < n; ++i) grains[i].store(0, std::memory_order_relaxed);
    }
    static uint64_t pack(float acc, float wt) { uint32_t a, b; memcpy(&a, &acc, 4); memcpy(&b, &wt, 4); return (uint64_t)a | ((uint64_t)b << 32); }
    static void unpack(uint64_t v, float& acc, float& wt) { uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32); memcpy(&acc, &a, 4); memcpy(&wt, &b, 4); }
    // film.rs:226-246
    bool to_pixel(Vec2 p, size_t& px, size_t& py) const {
        if (horizontal) { if (!(fabsf(p.y) <= ar_ratio)) return false; }
        else if (!(fabsf(p.x) <= ar_ratio)) return false;
        float x = horizontal ? p.x + 1.0f : p.x + ar_ratio;
        float y = horizontal ? p.y + ar_ratio : p.y + 1.0f;
        px = f32_as_usize(ar_size * x * 0.5f);
        py = f32_as_usize(ar_size * y * 0.5f);
        return true;
    }
    size_t wavelength_to_grain(float w) const { return f32_as_usize((w - wavelength_start) * grains_per_wavelength); }
    // Film::expose (film.rs:89-95) -> Grain::expose/increment (:128-162)
    void expose(Vec2 position, const Sample& s) {
        size_t grain = wavelength_to_grain(s.wavelength);
        size_t px, py;
        if (!to_pixel(position, px, py)) return;
        if (px >= width || py >= height) return;
        if (grain >= grains_per_pixel) throw std::runtime_error("grain index out of bounds");  // slice index panic
        std::atomic<uint64_t>& g = grains[(px + py * width) * grains_per_pixel + grain];
        float increment = s.brightness * s.weight, weight = s.weight;
        uint64_t cur = g.load(std::memory_order_relaxed);
        for (int attempts = 0; cas_attempts <= 0 || attempts < cas_attempts; ++attempts) {
            float acc, wt;
            unpack(cur, acc, wt);
            if (g.compare_exchange_strong(cur, pack(acc + increment, wt + weight), std::memory_order_relaxed)) break;
        }
    }
    float develop(size_t index) const {  // film.rs:132-143
        float acc, wt;
        unpack(grains[index].load(std::memory_order_relaxed), acc, wt);
        return wt > 0.0f ? acc / wt : 0.0f;
    }
    // Film::sample_many_wavelengths (film.rs:68-83)
    void sample_many_wavelengths(XorShift& rng, size_t amount, std::vector<float>& out) const {
        float step_size = wavelength_width / (float)amount;
        float from = wavelength_start;
        for (size_t i = 0; i < amount; ++i) {
            float to = from + step_size;
            out.push_back(rng.gen_range_f32(from, to));
            from = to;
        }
    }
};

// ---------------------------------------------------------------- renderer params (renderer/mod.rs:16, 63-75)
struct RendererParams {
    uint32_t algorithm = 0;  // 0 simple, 1 bidirectional
    uint32_t threads = 1, bounces = 8, pixel_samples = 1, light_samples = 4, spectrum_samples = 10;
    uint32_t spectrum_bins = 64, tile_size = 32, light_bounces = 8;
    float span_lo = 380.0f, span_hi = 780.0f;
    static RendererParams from_project(const Project& P) {
        RendererParams r;
        r.algorithm = P.renderer_type;
        unsigned hc = std::thread::hardware_concurrency();
        r.threads = P.threads.or_(hc ? hc : 1);
        r.bounces = P.bounces.or_(8);
        r.pixel_samples = P.pixel_samples;
        r.light_samples = P.light_samples.or_(4);
        r.spectrum_samples = P.spectrum_samples.or_(10);
        r.spectrum_bins = P.spectrum_resolution.or_(64);
        r.tile_size = P.tile_size.or_(32);
        r.light_bounces = P.light_bounces.or_(8);
        return r;
    }
};

// ---------------------------------------------------------------- tracer records (tracer.rs:157-200)
enum BounceTy { BT_DIFFUSE, BT_SPECULAR, BT_EMISSION };
struct DirectLight {
    bool dispersed = false;
    Program color;
    Vec3 incident, normal;
    Vec2 texture;
    float probability = 0;
};
struct Bounce {
    BounceTy ty = BT_SPECULAR;
    Vec3 out;  // BounceType::Diffuse(brdf, out)
    bool dispersed = false;
    Program color;
    Vec3 incident, position, normal;
    Vec2 texture;
    float probability = 0;
    std::vector<DirectLight> direct_light;
    // materials/diffuse.rs:27-29 `lambertian`, called as brdf(incident, normal, out) (tracer.rs:176-182)
    float brdf() const { return ty == BT_DIFFUSE ? 2.0f * fabsf(dot(out, normal)) : 1.0f; }
    float brdf_with(Vec3 incident_, Vec3 normal_) const { (void)incident_; return ty == BT_DIFFUSE ? 2.0f * fabsf(dot(out, normal_)) : 1.0f; }
};
inline float lambertian(Vec3 /*ray_in*/, Vec3 ray_out, Vec3 normal) { return 2.0f * fabsf(dot(normal, ray_out)); }

struct RenderCounters {
    std::atomic<uint64_t> rays{0}, nodes{0}, leaves{0}, path_samples{0}, de_evals{0}, de_iters{0};
};

// What tools/first_divergence.py compares per bounce of one path sample (filled by `trace` when TraceCtx::debug is set).
struct DebugBounce {
    uint32_t kind = 0, prim_id = 0xFFFFFFFFu;   // closest hit of the incoming ray (K_MISS: none)
    float t = 0, u = 0, v = 0;
    float incident[3] = {0, 0, 0}, position[3] = {0, 0, 0}, normal[3] = {0, 0, 0}, out[3] = {0, 0, 0};
    uint32_t n_direct = 0;   // direct-light samples drawn at this bounce
    uint32_t rng_w = 0;      // Xorshift128 `w` after everything this bounce drew
    uint32_t pad = 0;
};
struct TraceCtx {
    const World& W;
    XorShift& rng;
    TraceCounters tc;
    bool eager_emissive_draw = false;  // RNG stream mapping used by the wavefront pipeline
    std::vector<DebugBounce>* debug = nullptr;
};

// ---------------------------------------------------------------- BSDF scatter (materials/*.rs)
struct Scattering {
    bool emitted = false;
    Vec3 out_direction;
    float probability = 1.0f;
    bool dispersed = false;
    bool has_brdf = false;
};
inline Scattering scatter(const Component& c, Vec3 in_direction, Vec3 normal, float wavelength, XorShift& rng) {
    Scattering s;
    switch (c.bsdf) {
        case B_EMISSIVE: s.emitted = true; return s;
        case B_DIFFUSE: {  // diffuse.rs:8-25
            Vec3 n = dot(in_direction, normal) < 0.0f ? normal : -normal;
            s.out_direction = sample_hemisphere(rng, n);
            s.has_brdf = true;
            return s;
        }
        case B_MIRROR: {  // mirror.rs:5-21
            Vec3 n = dot(in_direction, normal) < 0.0f ? normal : -normal;
            float perp = dot(in_direction, n) * 2.0f;
            n = n * perp;
            s.out_direction = in_direction - n;
            return s;
        }
        default: {  // refractive.rs:6-91
            const RefractiveProps& p = c.props;
            s.dispersed = p.dispersion != 0.0f || p.env_dispersion != 0.0f;
            float ior = p.ior, env_ior = p.env_ior;
            if (s.dispersed) {
                float wl = wavelength * 0.001f;
                ior = p.ior + p.dispersion / (wl * wl);
                env_ior = p.env_ior + p.env_dispersion / (wl * wl);
            }
            Vec3 nl = dot(normal, in_direction) < 0.0f ? normal : -normal;
            Vec3 reflected = in_direction - (normal * 2.0f * dot(normal, in_direction));
            bool into = dot(normal, nl) > 0.0f;
            float nnt = into ? env_ior / ior : ior / env_ior;
            float ddn = dot(in_direction, nl);
            float cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
            if (cos2t < 0.0f) { s.out_direction = reflected; s.probability = 1.0f; return s; }
            float sgn = (into ? 1.0f : -1.0f) * (ddn * nnt + sqrtf(cos2t));
            Vec3 tdir = normalize(in_direction * nnt - normal * sgn);
            float a = ior - env_ior, b = ior + env_ior;
            float r0 = a * a / (b * b);
            float cc = 1.0f - (into ? -ddn : dot(tdir, normal));
            float re = r0 + (1.0f - r0) * cc * cc * cc * cc * cc;
            float tr = 1.0f - re;
            float pp = 0.25f + 0.5f * re;
            float rp = re / pp;
            float tp = tr / (1.0f - pp);
            if (rng.gen_f32() < pp) { s.out_direction = reflected; s.probability = rp; }
            else { s.out_direction = tdir; s.probability = tp; }
            return s;
        }
    }
}

// Material::apply_normal_map (materials/mod.rs:68-81)
inline Vec3 apply_normal_map(const World& W, const Material& m, const Normal& normal, Vec3 incident, Vec2 texture) {
    if (m.normal_map.present) {
        ProgramInputs in;
        in.normal = normal.vector; in.incident = incident; in.texture = texture;
        Vec4 v = run_vector(W.P, m.normal_map, in);
        return normalize(normal.from_space_v({v.x, v.y, v.z}));
    }
    return normal.vector;
}

// trace_directional (tracer.rs:444-459)
inline bool trace_directional(const World& W, Vec3 ray, Program& out) {
    for (auto& l : W.lights)
        if (l.kind == L_DIRECTIONAL && dot(l.direction, ray) >= l.width) { out = l.color; return true; }
    return false;
}

// trace_direct (tracer.rs:347-442)
inline void trace_direct(TraceCtx& cx, size_t samples, float wavelength, Vec3 ray_in, Vec3 position, Vec3 normal, std::vector<DirectLight>& out) {
    const World& W = cx.W;
    float lamp_probability;
    const Lamp* lamp = W.pick_lamp(cx.rng, lamp_probability);
    if (dot(ray_in, normal) >= 0.0f) normal = -normal;
    float probability = 1.0f / ((float)samples * 2.0f * PI * lamp_probability);
    for (size_t k = 0; k < samples; ++k) {
        LampSample ls = lamp_sample(W, *lamp, cx.rng, position);
        Ray ray_out{position, ls.direction};
        float cos_out = fmax_(dot(normal, ray_out.direction), 0.0f);
        if (!(cos_out > 0.0f)) continue;
        // in wavefront stream mode the emissive-component draw happens before
        // (and regardless of) the visibility result; the reference draws it only if unblocked.
        uint32_t eager_index = 0;
        if (cx.eager_emissive_draw && ls.surface.physical)
            eager_index = cx.rng.gen_index_u32((uint32_t)W.materials[ls.surface.material].emissive.size());
        Intersection hit;
        bool has_hit = W.intersect(ray_out, hit, &cx.tc);
        bool blocked;
        if (has_hit && ls.has_sq_distance) blocked = !(hit.distance * hit.distance >= ls.sq_distance - DIST_EPSILON);
        else if (!has_hit) blocked = false;
        else blocked = true;
        if (blocked) continue;
        DirectLight dl;
        float material_probability = 1.0f;
        if (ls.surface.physical) {
            const Material& m = W.materials[ls.surface.material];
            if (m.emissive.empty()) throw std::runtime_error("the material is not emissive");
            uint32_t ci = cx.eager_emissive_draw ? eager_index : cx.rng.gen_index_u32((uint32_t)m.emissive.size());
            const Component& component = m.emissive[ci];
            ProgramInputs in;
            in.wavelength = wavelength; in.normal = ls.surface.normal; in.incident = ray_out.direction; in.texture = ls.surface.texture;
            bool used;
            material_probability = component_probability(W.P, component, in, used);
            dl.color = component.color; dl.dispersed = used; dl.normal = ls.surface.normal; dl.texture = ls.surface.texture;
        } else {
            dl.color = ls.surface.color; dl.dispersed = false; dl.normal = -ray_out.direction; dl.texture = Vec2{0, 0};
        }
        float scale = ls.weight * probability * lambertian(ray_in, normal, ray_out.direction);
        dl.incident = ray_out.direction;
        dl.probability = scale * material_probability;
        out.push_back(dl);
    }
}

// trace (tracer.rs:208-345)
inline void trace(std::vector<Bounce>& path, TraceCtx& cx, Ray ray, float wavelength, uint32_t bounces, size_t light_samples) {
    const World& W = cx.W;
    bool sample_light = true;
    int light_sample_events = 0;
    for (uint32_t b = 0; b < bounces; ++b) {
        Intersection isect;
        if (W.intersect(ray, isect, &cx.tc)) {
            const Material& material = W.material_of(isect.surface_point);
            SurfaceData sd = W.surface_data(isect.surface_point);
            Vec3 normal = apply_normal_map(W, material, sd.normal, ray.direction, sd.texture);
            Vec3 position = isect.surface_point.position;
            if (material.components.empty()) throw std::runtime_error("there should be at least one component");
            const Component& component = material.components[cx.rng.gen_index_u32((uint32_t)material.components.size())];
            ProgramInputs pin;
            pin.wavelength = wavelength; pin.normal = normal; pin.incident = ray.direction; pin.texture = sd.texture;
            bool normal_dispersed;
            float component_prob = component_probability(W.P, component, pin, normal_dispersed);
            Scattering sc = scatter(component, ray.direction, normal, wavelength, cx.rng);
            if (!sc.emitted) {
                Bounce bounce;
                if (light_sample_events < 2) {
                    sample_light = !sc.has_brdf || light_samples == 0;
                    if (sc.has_brdf) {
                        light_sample_events += 1;
                        trace_direct(cx, light_samples, wavelength, ray.direction, position, normal, bounce.direct_light);
                    }
                } else {
                    sample_light = true;
                }
                bounce.ty = sc.has_brdf ? BT_DIFFUSE : BT_SPECULAR;
                bounce.out = sc.out_direction;
                bounce.dispersed = sc.dispersed || normal_dispersed;
                bounce.color = component.color;
                bounce.incident = ray.direction;
                bounce.position = position;
                bounce.normal = normal;
                bounce.texture = sd.texture;
                bounce.probability = sc.probability * component_prob;
                if (cx.debug) {
                    DebugBounce db;
                    const SurfacePoint& sp = isect.surface_point;
                    db.kind = sp.kind; db.prim_id = sp.kind == K_PLANE ? sp.plane->id : sp.shape->id; db.t = isect.distance; db.u = sp.u; db.v = sp.v;
                    db.incident[0] = ray.direction.x; db.incident[1] = ray.direction.y; db.incident[2] = ray.direction.z;
                    db.position[0] = position.x; db.position[1] = position.y; db.position[2] = position.z;
                    db.normal[0] = normal.x; db.normal[1] = normal.y; db.normal[2] = normal.z;
                    db.out[0] = sc.out_direction.x; db.out[1] = sc.out_direction.y; db.out[2] = sc.out_direction.z;
                    db.n_direct = (uint32_t)bounce.direct_light.size(); db.rng_w = cx.rng.w;
                    cx.debug->push_back(db);
                }
                ray = Ray{position, sc.out_direction};
                path.push_back(std::move(bounce));
            } else {
                if (cx.debug) {
                    DebugBounce db;
                    const SurfacePoint& sp = isect.surface_point;
                    db.kind = sp.kind; db.prim_id = sp.kind == K_PLANE ? sp.plane->id : sp.shape->id; db.t = isect.distance; db.u = sp.u; db.v = sp.v;
                    db.incident[0] = ray.direction.x; db.incident[1] = ray.direction.y; db.incident[2] = ray.direction.z;
                    db.position[0] = position.x; db.position[1] = position.y; db.position[2] = position.z;
                    db.normal[0] = normal.x; db.normal[1] = normal.y; db.normal[2] = normal.z;
                    db.rng_w = cx.rng.w;
                    cx.debug->push_back(db);
                }
                if (sample_light) {
                    Bounce bounce;
                    bounce.ty = BT_EMISSION;
                    bounce.dispersed = normal_dispersed;
                    bounce.color = component.color;
                    bounce.incident = ray.direction;
                    bounce.position = position;
                    bounce.normal = normal;
                    bounce.texture = sd.texture;
                    bounce.probability = component_prob;
                    path.push_back(std::move(bounce));
                }
                break;
            }
        } else {
            if (cx.debug) {
                DebugBounce db;
                db.incident[0] = ray.direction.x; db.incident[1] = ray.direction.y; db.incident[2] = ray.direction.z;
                db.rng_w = cx.rng.w;
                cx.debug->push_back(db);
            }
            Program color = W.sky;
            if (sample_light) trace_directional(W, ray.direction, color);
            Bounce bounce;
            bounce.ty = BT_EMISSION;
            bounce.dispersed = false;
            bounce.color = color;
            bounce.incident = ray.direction;
            bounce.position = ray.direction * INF;
            bounce.normal = -ray.direction;
            bounce.texture = Vec2{0, 0};
            bounce.probability = 1.0f;
            path.push_back(std::move(bounce));
            break;
        }
    }
}

// ---------------------------------------------------------------- contribute (renderer/algorithm.rs:14-100)
struct WSample { Sample s; float reflectance = 1.0f; };  // (Sample, f32)

inline void contribute(const World& W, const Bounce& b, WSample& main_sample, WSample* additional, size_t n_additional) {
    ProgramInputs in;
    in.incident = b.incident; in.normal = b.normal; in.texture = b.texture;
    if (b.ty == BT_EMISSION) {
        in.wavelength = main_sample.s.wavelength;
        main_sample.s.brightness += run_number(W.P, b.color, in) * b.probability * main_sample.reflectance;
        for (size_t i = 0; i < n_additional; ++i) {
            in.wavelength = additional[i].s.wavelength;
            additional[i].s.brightness += run_number(W.P, b.color, in) * b.probability * additional[i].reflectance;
        }
    } else {
        in.wavelength = main_sample.s.wavelength;
        main_sample.reflectance *= run_number(W.P, b.color, in) * b.probability;
        for (size_t i = 0; i < n_additional; ++i) {
            in.wavelength = additional[i].s.wavelength;
            additional[i].reflectance *= run_number(W.P, b.color, in) * b.probability;
        }
        for (auto& d : b.direct_light) {
            ProgramInputs li;
            li.incident = d.incident; li.normal = d.normal; li.texture = d.texture;
            li.wavelength = main_sample.s.wavelength;
            main_sample.s.brightness += run_number(W.P, d.color, li) * d.probability * main_sample.reflectance;
            if (!d.dispersed) {
                for (size_t i = 0; i < n_additional; ++i) {
                    li.wavelength = additional[i].s.wavelength;
                    additional[i].s.brightness += run_number(W.P, d.color, li) * d.probability * additional[i].reflectance;
                }
            }
        }
        float brdf = b.brdf();
        main_sample.reflectance *= brdf;
        for (size_t i = 0; i < n_additional; ++i) additional[i].reflectance *= brdf;
    }
}

// ---------------------------------------------------------------- tiles (renderer/algorithm.rs:102-188, cameras.rs:57-68)
struct Tile {
    Vec2 from, size;  // view-space area
    size_t width = 0, height = 0;
    size_t index = 0;  // row-major tile index before the centre-out sort (RNG key)
    size_t area() const { return width * height; }
    Vec2 sample_point(XorShift& rng) const {
        float ox = size.x * rng.gen_f32();
        float oy = size.y * rng.gen_f32();
        return Vec2{from.x + ox, from.y + oy};
    }
    float center_mag2() const {
        float cx = from.x + size.x / 2.0f, cy = from.y + size.y / 2.0f;
        return cx * cx + cy * cy;
    }
};
inline std::vector<Tile> make_tiles(size_t film_width, size_t film_height, size_t tile_size) {
    size_t tiles_x = film_width / tile_size;
    if (tiles_x * tile_size < film_width) tiles_x += 1;
    size_t tiles_y = film_height / tile_size;
    if (tiles_y * tile_size < film_height) tiles_y += 1;
    std::vector<Tile> tiles;
    float fw = (float)film_width, fh = (float)film_height;
    float max_dimension = fmax_(fw, fh);
    for (size_t y = 0; y < tiles_y; ++y)
        for (size_t x = 0; x < tiles_x; ++x) {
            size_t sx = x * tile_size, sy = y * tile_size;
            size_t w = std::min(film_width - sx, tile_size), h = std::min(film_height - sy, tile_size);
            Tile t;
            // Camera::to_view_area
            t.from = Vec2{((float)sx + (-fw * 0.5f)) / (max_dimension * 0.5f), ((float)sy + (-fh * 0.5f)) / (max_dimension * 0.5f)};
            t.size = Vec2{(float)w / (max_dimension * 0.5f), (float)h / (max_dimension * 0.5f)};
            t.width = w; t.height = h;
            t.index = y * tiles_x + x;
            tiles.push_back(t);
        }
    std::stable_sort(tiles.begin(), tiles.end(), [](const Tile& a, const Tile& b) { return a.center_mag2() < b.center_mag2(); });
    return tiles;
}

// ---------------------------------------------------------------- render options
struct RenderOptions {
    uint64_t seed = 1;
    // 0: one Xorshift128 stream per tile (the reference's structure, simple.rs:41-47)
    // 1: one keyed stream per path sample (tile, i) - what the GPU wavefront uses
    int rng_mode = 1;
    bool eager_emissive_draw = true;
    uint32_t spp_override = 0;     // 0 = project's pixel_samples
    uint32_t sample_offset = 0;    // first per-tile sample index (multi-rank sharding by sample pass)
    uint32_t sample_stride = 1;    // samples i = offset, offset+stride, ... (< area*spp)
    int threads = 0;               // 0 = renderer.threads
    int cas_attempts = 5;
    int64_t only_tile = -1;        // >= 0: render this tile only (tools/find_nonfinite.py replays single path samples)
    uint64_t sample_limit = 0;     // > 0: at most this many samples per tile
};

struct RenderState {
    const World& W;
    const Camera& camera;
    const RendererParams& R;
    Film& film;
    const RenderOptions& opt;
    RenderCounters& counters;
};

// renderer/simple.rs:58-141 `render_tile`
inline void render_tile_simple(RenderState& st, const Tile& tile) {
    const RendererParams& R = st.R;
    std::vector<WSample> additional;
    std::vector<float> wavelengths;
    std::vector<Bounce> path;
    uint32_t spp = st.opt.spp_override ? st.opt.spp_override : R.pixel_samples;
    uint64_t iterations = (uint64_t)tile.area() * spp;
    XorShift rng = keyed_rng(st.opt.seed, tile.index, ~0ull);
    TraceCtx cx{st.W, rng, {}, st.opt.eager_emissive_draw};
    uint64_t done = 0;
    for (uint64_t i = st.opt.sample_offset; i < iterations; i += st.opt.sample_stride) {
        if (st.opt.sample_limit && done >= st.opt.sample_limit) break;
        if (st.opt.rng_mode == 1) rng = keyed_rng(st.opt.seed, tile.index, i);
        additional.clear(); path.clear(); wavelengths.clear();
        Vec2 position = tile.sample_point(rng);
        Ray ray = st.camera.ray_towards(position, rng);
        st.film.sample_many_wavelengths(rng, R.spectrum_samples, wavelengths);
        for (float w : wavelengths) additional.push_back(WSample{Sample{0.0f, w, 1.0f}, 1.0f});
        size_t pick = rng.gen_range_usize(additional.size());
        WSample main_sample = additional[pick];  // swap_remove
        additional[pick] = additional.back();
        additional.pop_back();
        float wavelength = main_sample.s.wavelength;
        trace(path, cx, ray, wavelength, R.bounces, R.light_samples);
        bool use_additional = true;
        for (auto& bounce : path) {
            use_additional = !bounce.dispersed && use_additional;
            contribute(st.W, bounce, main_sample, additional.data(), use_additional ? additional.size() : 0);
        }
        st.film.expose(position, main_sample.s);
        if (use_additional)
            for (auto& a : additional) st.film.expose(position, a.s);
        ++done;
    }
    st.counters.rays += cx.tc.rays; st.counters.nodes += cx.tc.nodes; st.counters.leaves += cx.tc.leaves;
    st.counters.path_samples += done;
}

// One `render_tile` iteration of the camera-to-light integrator (simple.rs:87-139) for path sample (tile, i) on its keyed
// stream, with the per-bounce records of `trace` and the exposed (brightness, wavelength) pairs: tools/first_divergence.py.
inline void debug_path_simple(const World& W, const Camera& camera, const RendererParams& R, const Film& film, const RenderOptions& opt,
                              const Tile& tile, uint64_t i, std::vector<DebugBounce>& bounces, Vec2& position_out, std::vector<Sample>& exposed) {
    XorShift rng = keyed_rng(opt.seed, tile.index, i);
    TraceCtx cx{W, rng, {}, opt.eager_emissive_draw, &bounces};
    std::vector<WSample> additional;
    std::vector<float> wavelengths;
    std::vector<Bounce> path;
    Vec2 position = tile.sample_point(rng);
    Ray ray = camera.ray_towards(position, rng);
    film.sample_many_wavelengths(rng, R.spectrum_samples, wavelengths);
    for (float w : wavelengths) additional.push_back(WSample{Sample{0.0f, w, 1.0f}, 1.0f});
    size_t pick = rng.gen_range_usize(additional.size());
    WSample main_sample = additional[pick];
    additional[pick] = additional.back();
    additional.pop_back();
    trace(path, cx, ray, main_sample.s.wavelength, R.bounces, R.light_samples);
    bool use_additional = true;
    for (auto& bounce : path) {
        use_additional = !bounce.dispersed && use_additional;
        contribute(W, bounce, main_sample, additional.data(), use_additional ? additional.size() : 0);
    }
    position_out = position;
    exposed.push_back(main_sample.s);
    if (use_additional)
        for (auto& a : additional) exposed.push_back(a.s);
}

// Camera::is_visible (cameras.rs:99-158)
inline bool camera_is_visible(const Camera& cam, Vec3 target, TraceCtx& cx, Vec2& out_pos, Ray& out_ray) {
    Mat4 inv_transform;
    if (!invert(cam.transform, inv_transform)) return false;
    Vec3 local_target = transform_point(inv_transform, target);
    if (local_target.z >= 0.0f) return false;
    Vec3 origin{0, 0, 0};
    if (cam.aperture > 0.0f) {
        float sqrt_r = sqrtf(cam.aperture * cx.rng.gen_f32());
        float psi = PI * 2.0f * cx.rng.gen_f32();
        origin = {sqrt_r * m_cos(psi), sqrt_r * m_sin(psi), 0.0f};
    }
    Vec3 world_origin = transform_point(cam.transform, origin);
    Vec3 direction = target - world_origin;
    float distance = magnitude(direction);
    Ray ray{world_origin, direction / distance};
    Intersection hit;
    if (cx.W.intersect(ray, hit, &cx.tc) && hit.distance < distance - DIST_EPSILON) return false;
    local_target.z += cam.focus_distance;
    float dist = local_target.z;
    local_target = local_target - (origin * dist) / cam.focus_distance;
    local_target.z -= cam.focus_distance;
    Vec3 view_plane_target = (-local_target) / local_target.z;
    float focus_x = view_plane_target.x, focus_y = -view_plane_target.y;
    out_pos = Vec2{focus_x * cam.view_plane, focus_y * cam.view_plane};
    out_ray = ray;
    return true;
}

// connect_paths (renderer/bidirectional.rs:310-398)
inline void connect_paths(TraceCtx& cx, const Bounce& bounce, const WSample& main_in, const std::vector<WSample>& additional_in,
                          const std::vector<Bounce>& path, bool use_additional_in, std::vector<Sample>& contributions) {
    if (bounce.ty != BT_DIFFUSE) return;
    for (size_t i = 0; i < path.size(); ++i) {
        const Bounce& lamp_bounce = path[i];
        if (lamp_bounce.ty == BT_SPECULAR) continue;
        Vec3 from = bounce.position, to = lamp_bounce.position;
        Vec3 direction = to - from;
        float sq_distance = magnitude2(direction);
        float distance = sqrtf(sq_distance);
        Ray ray{from, direction / distance};
        if (dot(bounce.normal, ray.direction) <= 0.0f) continue;
        if (dot(lamp_bounce.normal, -ray.direction) <= 0.0f) continue;
        Intersection hit;
        if (cx.W.intersect(ray, hit, &cx.tc) && hit.distance < distance - DIST_EPSILON) continue;
        float cos_out = fabsf(dot(bounce.normal, ray.direction));
        float cos_in = fabsf(dot(lamp_bounce.normal, -ray.direction));
        float brdf_out = lambertian(bounce.incident, bounce.normal, ray.direction) / bounce.brdf();
        float scale = cos_in * cos_out * brdf_out / (2.0f * PI * sq_distance);
        float brdf_in = lamp_bounce.brdf_with(-ray.direction, lamp_bounce.normal) / lamp_bounce.brdf();
        bool use_additional = use_additional_in;
        std::vector<WSample> additional = additional_in;
        for (auto& a : additional) a.reflectance = a.reflectance * scale;
        WSample main_sample = main_in;
        main_sample.reflectance *= scale;
        for (size_t k = i; k < path.size(); ++k) {
            use_additional = !path[k].dispersed && use_additional;
            size_t n_add = use_additional ? additional.size() : 0;
            contribute(cx.W, path[k], main_sample, additional.data(), n_add);
            if (k == i) {
                main_sample.reflectance *= brdf_in;
                for (size_t a = 0; a < n_add; ++a) additional[a].reflectance *= brdf_in;
            }
        }
        contributions.push_back(main_sample.s);
        if (use_additional)
            for (auto& a : additional) contributions.push_back(a.s);
    }
}

// renderer/bidirectional.rs:73-308 `render_tile`
inline void render_tile_bidirectional(RenderState& st, const Tile& tile) {
    const RendererParams& R = st.R;
    const World& W = st.W;
    std::vector<Bounce> lamp_path, camera_path;
    std::vector<WSample> additional;
    std::vector<float> wavelengths;
    std::vector<Sample> contributions;
    uint32_t spp = st.opt.spp_override ? st.opt.spp_override : R.pixel_samples;
    uint64_t iterations = (uint64_t)tile.area() * spp;
    XorShift rng = keyed_rng(st.opt.seed, tile.index, ~0ull);
    TraceCtx cx{W, rng, {}, st.opt.eager_emissive_draw};
    uint64_t done = 0;
    for (uint64_t it = st.opt.sample_offset; it < iterations; it += st.opt.sample_stride) {
        if (st.opt.sample_limit && done >= st.opt.sample_limit) break;
        if (st.opt.rng_mode == 1) rng = keyed_rng(st.opt.seed, tile.index, it);
        lamp_path.clear(); camera_path.clear(); additional.clear(); wavelengths.clear();
        Vec2 position = tile.sample_point(rng);
        st.film.sample_many_wavelengths(rng, R.spectrum_samples, wavelengths);
        for (float w : wavelengths) additional.push_back(WSample{Sample{0.0f, w, 1.0f}, 1.0f});
        size_t pick = rng.gen_range_usize(additional.size());
        WSample main_sample = additional[pick];
        additional[pick] = additional.back();
        additional.pop_back();
        float wavelength = main_sample.s.wavelength;

        Ray camera_ray = st.camera.ray_towards(position, rng);
        float lamp_probability;
        const Lamp* lamp = W.pick_lamp(rng, lamp_probability);
        RaySample lamp_sample;
        if (lamp_sample_ray(W, *lamp, rng, lamp_sample)) {
            Ray ray = lamp_sample.ray;
            Program color; float material_probability; bool dispersed; Vec3 normal; Vec2 texture;
            if (lamp_sample.surface.physical) {
                const Material& m = W.materials[lamp_sample.surface.material];
                if (m.emissive.empty()) throw std::runtime_error("the material is not emissive");
                const Component& component = m.emissive[rng.gen_index_u32((uint32_t)m.emissive.size())];
                ProgramInputs in;
                in.wavelength = wavelength; in.normal = lamp_sample.surface.normal; in.incident = -ray.direction; in.texture = lamp_sample.surface.texture;
                material_probability = component_probability(W.P, component, in, dispersed);
                color = component.color; normal = lamp_sample.surface.normal; texture = lamp_sample.surface.texture;
            } else {
                color = lamp_sample.surface.color; material_probability = 1.0f; dispersed = false; normal = ray.direction; texture = Vec2{0, 0};
            }
            ray.origin = ray.origin + normal * DIST_EPSILON;
            Bounce first;
            first.ty = BT_EMISSION; first.dispersed = dispersed; first.color = color; first.incident = Vec3(0, 0, 0);
            first.position = ray.origin; first.normal = normal; first.texture = texture;
            first.probability = lamp_sample.weight / (lamp_probability * material_probability);
            lamp_path.push_back(first);
            trace(lamp_path, cx, ray, wavelength, R.light_bounces, 0);
            // utils::pairs (utils.rs:5-13) visits positions 0 .. len-2 EXCLUSIVE: the last pair is skipped
            if (lamp_path.size() >= 2)
                for (size_t pos = 0; pos + 2 < lamp_path.size(); ++pos) {
                    Bounce& to = lamp_path[pos];
                    Bounce& from = lamp_path[pos + 1];
                    to.incident = -from.incident;
                    if (from.ty == BT_DIFFUSE) from.out = from.incident;
                }
            if (lamp_path.size() > 1 && lamp_path.back().ty == BT_EMISSION) lamp_path.pop_back();
            std::reverse(lamp_path.begin(), lamp_path.end());
        }

        trace(camera_path, cx, camera_ray, wavelength, R.bounces, R.light_samples);

        float total = (float)(camera_path.size() * lamp_path.size());
        float weight = 1.0f / total;
        bool use_additional = true;
        for (auto& bounce : camera_path) {
            use_additional = !bounce.dispersed && use_additional;
            contribute(W, bounce, main_sample, additional.data(), use_additional ? additional.size() : 0);
            contributions.clear();
            connect_paths(cx, bounce, main_sample, additional, lamp_path, use_additional, contributions);
            for (auto& c : contributions) { c.weight = weight; st.film.expose(position, c); }
        }
        st.film.expose(position, main_sample.s);
        if (use_additional)
            for (auto& a : additional) st.film.expose(position, a.s);

        weight = 1.0f / (float)lamp_path.size();
        for (size_t i = 0; i < lamp_path.size(); ++i) {
            const Bounce& bounce = lamp_path[i];
            if (bounce.ty != BT_DIFFUSE) continue;
            Vec2 hit_pos; Ray hit_ray;
            if (!camera_is_visible(st.camera, bounce.position, cx, hit_pos, hit_ray)) continue;
            if (!(hit_pos.x > -1.0f && hit_pos.x < 1.0f && hit_pos.y > -1.0f && hit_pos.y < 1.0f)) continue;
            float sq_distance = magnitude2(hit_ray.origin - bounce.position);
            float scale = 1.0f / sq_distance;
            float brdf_in = bounce.brdf_with(-hit_ray.direction, bounce.normal) / bounce.brdf();
            main_sample.s.brightness = 0.0f; main_sample.s.weight = weight; main_sample.reflectance = scale;
            use_additional = true;
            for (auto& a : additional) { a.s.brightness = 0.0f; a.s.weight = weight; a.reflectance = scale; }
            for (size_t k = i; k < lamp_path.size(); ++k) {
                use_additional = !lamp_path[k].dispersed && use_additional;
                size_t n_add = use_additional ? additional.size() : 0;
                contribute(W, lamp_path[k], main_sample, additional.data(), n_add);
                if (k == i) {
                    main_sample.reflectance *= brdf_in;
                    for (size_t a = 0; a < n_add; ++a) additional[a].reflectance *= brdf_in;
                }
            }
            st.film.expose(hit_pos, main_sample.s);
            if (use_additional)
                for (auto& a : additional) st.film.expose(hit_pos, a.s);
        }
        ++done;
    }
    st.counters.rays += cx.tc.rays; st.counters.nodes += cx.tc.nodes; st.counters.leaves += cx.tc.leaves;
    st.counters.path_samples += done;
}

// Renderer::render -> TaskRunner::run_tasks (renderer/mod.rs:77-111, 125-189): `threads` workers pull tiles
inline void render(RenderState& st) {
    std::vector<Tile> tiles = make_tiles(st.film.width, st.film.height, st.R.tile_size);
    int threads = st.opt.threads > 0 ? st.opt.threads : (int)st.R.threads;
    if (threads < 1) threads = 1;
    st.film.cas_attempts = st.opt.cas_attempts;
    std::atomic<size_t> next{0};
    std::mutex err_mutex;
    std::string error;
    auto worker = [&]() {
        tl_de_evals = 0; tl_de_iters = 0;
        try {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= tiles.size()) break;
                if (st.opt.only_tile >= 0 && (int64_t)tiles[i].index != st.opt.only_tile) continue;
                if (st.R.algorithm == 0) render_tile_simple(st, tiles[i]);
                else if (st.R.algorithm == 1) render_tile_bidirectional(st, tiles[i]);
                else throw std::runtime_error("photon mapping is out of scope (SURVEY.md §2 row 21)");
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> g(err_mutex);
            if (error.empty()) error = e.what();
            next.store(tiles.size());
        }
        st.counters.de_evals += tl_de_evals; st.counters.de_iters += tl_de_iters;
    };
    if (threads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    if (!error.empty()) throw std::runtime_error(error);
}

// ---------------------------------------------------------------- develop (main.rs:190-238, 313-418)
struct Developer {
    const World& W;
    const Project& P;
    Program filter, white;
    float white_max = 0, d65_max = 0;

    float d65(float w) const { return array_get(P.d65.data(), P.d65.size(), 1, P.illum_min, P.illum_max, w); }
    Developer(const World& w, float span_lo, float span_hi) : W(w), P(w.P) {
        ProgramCompiler pc{const_cast<Project&>(P)};
        if (P.filter.present) filter = pc.compile(P.filter.e, false, ALLOW_SPECTRUM);
        if (P.white.present) {
            white = pc.compile(P.white.e, false, ALLOW_SPECTRUM);
            float wavelength = span_lo;  // main.rs:206-214
            while (wavelength < span_hi) {
                ProgramInputs in; in.wavelength = wavelength;
                white_max = fmax_(white_max, run_number(P, white, in));
                d65_max = fmax_(d65_max, d65(wavelength));
                wavelength += 1.0f;
            }
        }
    }
    // spectrum_get closure (main.rs:224-238)
    float adjust(float intensity, float wavelength) const {
        ProgramInputs in; in.wavelength = wavelength;
        float filtered = filter.present ? intensity * run_number(P, filter, in) : intensity;
        if (white.present) {
            float white_intensity = run_number(P, white, in) / white_max;
            float neutral = filtered / fmax_(white_intensity, 0.000001f);
            return neutral * (d65(wavelength) / d65_max);
        }
        return filtered;
    }
};

// film.rs:321-337 `Spectrum::get` for one pixel's grains
inline float film_spectrum_get(const Film& film, size_t pixel, float w) {
    float min = film.wavelength_start, max = film.wavelength_start + film.wavelength_width;
    if (w < min) return 0.0f;
    if (w > max) return 0.0f;
    float normalized = (w - min) / (max - min);
    float float_index = normalized * (float)film.grains_per_pixel;
    size_t index = std::min(f32_as_usize(floorf(float_index)), film.grains_per_pixel - 1);
    return film.develop(pixel * film.grains_per_pixel + index);
}

// spectrum_to_xyz / spectrum_to_tristimulus (main.rs:352-418)
inline void pixel_to_xyz(const Film& film, const Developer& dev, size_t pixel, float step_size, float out[3]) {
    const Project& P = dev.P;
    size_t n = P.xyz.size() / 3;
    auto resp = [&](float w, float r[3]) {
        for (int c = 0; c < 3; ++c) r[c] = array_get(P.xyz.data() + c, n, 3, P.xyz_min, P.xyz_max, w);
    };
    float min = film.wavelength_start, max = film.wavelength_start + film.wavelength_width;
    float sum[3] = {0, 0, 0};
    float weight = 0.0f;
    float wl_min = min;
    float spectrum_min = dev.adjust(film_spectrum_get(film, pixel, wl_min), wl_min);
    float start_resp[3];
    resp(wl_min, start_resp);
    while (wl_min < max) {
        float wl_max = wl_min + step_size;
        float spectrum_max = dev.adjust(film_spectrum_get(film, pixel, wl_max), wl_max);
        float end_resp[3];
        resp(wl_max, end_resp);
        float w = wl_max - wl_min;
        for (int c = 0; c < 3; ++c) sum[c] += ((start_resp[c] * spectrum_min + end_resp[c] * spectrum_max) * 0.5f) * w;
        weight += w;
        wl_min = wl_max;
        spectrum_min = spectrum_max;
        for (int c = 0; c < 3; ++c) start_resp[c] = end_resp[c];
    }
    for (int c = 0; c < 3; ++c) out[c] = (weight == 0.0f ? sum[c] : sum[c] / weight) * 3.444f;
}

// palette: LinSrgb::from_color(Xyz<D65>) then into_encoding::<Srgb<u8>> (main.rs:323)
inline void xyz_to_srgb8(const float xyz[3], uint8_t out[3]) {
    float x = xyz[0], y = xyz[1], z = xyz[2];
    float lin[3] = {(3.2404542f * x + -1.5371385f * y) + -0.4985314f * z, (-0.9692660f * x + 1.8760108f * y) + 0.0415560f * z,
                    (0.0556434f * x + -0.2040259f * y) + 1.0572252f * z};
    for (int c = 0; c < 3; ++c) {
        float v = lin[c];
        if (!(v > 0.0f)) v = 0.0f;
        if (v > 1.0f) v = 1.0f;
        float e = v <= 0.0031308f ? 12.92f * v : 1.055f * m_pow(v, 1.0f / 2.4f) - 0.055f;
        float s = e * 255.0f + 0.5f;
        out[c] = (uint8_t)(s < 0.0f ? 0.0f : (s > 255.0f ? 255.0f : s));
    }
}

}  // namespace pyro
