// Per-path-sample stage logic of the wavefront pipeline: camera / wavelength sampling, surface
// data, BSDF scattering, next-event estimation, the `contribute` fold and film exposure.
// __host__ __device__ like core.cuh (the CPU unit tests run the same code); kernels.cu wraps each
// stage in a kernel and owns the queues.
#pragma once
#include "core.cuh"

namespace pyr {

// ---------------------------------------------------------------- path state (one per path sample in flight)
struct alignas(32) PendingLight {   // 32 B; tracer.rs:193-200 `DirectLight`, waiting for its visibility ray
    int32_t color_program;
    uint32_t dispersed;
    float normal[3];
    float tex[2];
    float probability;
};
enum : uint32_t { PS_USE_ADDITIONAL = 1u, PS_SAMPLE_LIGHT = 2u, PS_HAS_MAIN = 4u, PS_PENDING_FOLD = 8u, PS_ALIVE = 16u };

// What is stored per path slot in HBM: a 64-byte header and the three per-wavelength arrays, 256 B.
struct alignas(32) PathHeader {
    Rng rng;
    float pos[2];              // film position in view coordinates (Tile::sample_point)
    uint32_t tile, bounce;
    // the four words k_bin reads (one 16-byte load, bytes 32..47 of the record)
    uint32_t flags, n_pending, ray_base;
    uint32_t shadow_base;      // index of this path's first visibility ray in the shadow region
    uint32_t light_events;
    float pending_brdf;
    uint32_t pad[2];
};
struct alignas(32) PathCore : PathHeader {
    // wl[S] | bright[S] | refl[S], packed with stride S = spectrum_samples: wl[0] is the hero wavelength, then the
    // additional ones (simple.rs:105-107); bright = Sample::brightness; refl = the running reflectance of
    // renderer/algorithm.rs:14-100.  ceil(3 S / 8) 32-byte chunks are moved per pass.
    float spectral[3 * MAX_SPECTRUM_SAMPLES];
};
// bidirectional integrator only (bdpt.cuh)
struct alignas(32) BidirState {   // 80 B of state, padded to three 32-byte chunks
    uint32_t phase, n_light, n_cam, n_cam_stored, lamp_bounces, conn_cam, conn_light, conn_next;
    float cam_o[3], cam_d[3];
    uint32_t cam_store_pending;
    uint32_t conn_next_cam;   // connect phase: the camera vertex the next batch starts at (with conn_next = its first lamp vertex)
    Rng rng_saved;
    uint32_t pad3[4];
};
// A per-wavelength array of the thread.  In the kernels it lives in dynamic shared memory, [wavelength][thread]
// (dynamically indexed thread-local arrays would otherwise go through local memory); on the host it is plain memory.
struct SpecArray {
    float* base;
#if defined(__CUDA_ARCH__)
    __device__ __forceinline__ float& operator[](uint32_t k) const { return base[k * PYR_BLOCK]; }
#else
    float& operator[](uint32_t k) const { return base[k]; }
#endif
};
// What the stage functions see: the thread's copy of the header, its per-wavelength arrays, and where the colder
// records live.
struct PathState : PathHeader {
    SpecArray wl, bright, refl;
    PendingLight* pend;        // MAX_LIGHT_SAMPLES records per path (HBM)
    BidirState* bd;
};

// What one shade step asks to be traced next.  The visibility rays are staged on chip in the kernels (dynamic shared
// memory behind the VM registers): the light samples of one next-event estimation leave from the same surface point, so
// the staging area holds that origin once and a (direction, limit) quad per ray - (1 + light_samples) float4 per thread.
PYR_HD uint32_t stage_quads(uint32_t light_samples) { return 1u + (light_samples ? light_samples : 1u); }
struct ShadeOut {
    uint32_t alive, has_main, n_shadow, pad;
    Ray main;
#if defined(__CUDA_ARCH__)
    uint32_t stage_base;  // first float4 of the staging area = vm_regs * PYR_BLOCK
    __device__ __forceinline__ void put_shadow(uint32_t j, const Ray& r) {
        float4* s = pyr_dyn_smem + stage_base + threadIdx.x;
        if (j == 0) s[0] = make_float4(r.o[0], r.o[1], r.o[2], __uint_as_float(r.mode));  // shared by the rays of this event
        s[(1 + j) * PYR_BLOCK] = make_float4(r.d[0], r.d[1], r.d[2], r.limit);
    }
    __device__ __forceinline__ Ray get_shadow(uint32_t j) const {
        const float4* s = pyr_dyn_smem + stage_base + threadIdx.x;
        const float4 a = s[0], b = s[(1 + j) * PYR_BLOCK];
        Ray r;
        r.o[0] = a.x; r.o[1] = a.y; r.o[2] = a.z; r.mode = __float_as_uint(a.w);
        r.d[0] = b.x; r.d[1] = b.y; r.d[2] = b.z; r.limit = b.w;
        return r;
    }
#else
    uint32_t stage_base;
    Ray shadow_[MAX_LIGHT_SAMPLES];
    void put_shadow(uint32_t j, const Ray& r) { shadow_[j] = r; }
    Ray get_shadow(uint32_t j) const { return shadow_[j]; }
#endif
};

// ---------------------------------------------------------------- film (film.rs)
// AspectRatio::contains + to_pixel (film.rs:226-246)
PYR_HD bool film_to_pixel(const FilmRec& f, float x, float y, uint64_t& px, uint64_t& py) {
    if (f.horizontal) { if (!(fabsf(y) <= f.ar_ratio)) return false; }
    else if (!(fabsf(x) <= f.ar_ratio)) return false;
    float fx = f.horizontal ? x + 1.0f : x + f.ar_ratio;
    float fy = f.horizontal ? y + f.ar_ratio : y + 1.0f;
    px = f32_as_usize(f.ar_size * fx * 0.5f);
    py = f32_as_usize(f.ar_size * fy * 0.5f);
    return true;
}
// Film::expose (film.rs:89-95) -> Grain::expose (:128-162).  `add(index, increment, weight)` does
// the accumulation (an atomic add on the device; the reference's CAS may drop samples, ours never does).
template <class Add>
PYR_HD void film_expose(const FilmRec& f, float x, float y, float brightness, float wavelength, float weight, Add& add) {
    uint64_t grain = f32_as_usize((wavelength - f.wavelength_start) * f.grains_per_wavelength);
    uint64_t px, py;
    if (!film_to_pixel(f, x, y, px, py)) return;
    if (px >= f.width || py >= f.height) return;
    if (grain >= f.bins) return;  // the reference would panic on the slice index (film.rs:93)
    add((px + py * (uint64_t)f.width) * f.bins + grain, brightness * weight, weight);
}

// Several exposures of the same bin at once: `increment` = sum of brightness * weight, `weight` = sum of weights.
template <class Add>
PYR_HD void film_expose_sum(const FilmRec& f, float x, float y, float increment, float wavelength, float weight, Add& add) {
    uint64_t grain = f32_as_usize((wavelength - f.wavelength_start) * f.grains_per_wavelength);
    uint64_t px, py;
    if (!film_to_pixel(f, x, y, px, py)) return;
    if (px >= f.width || py >= f.height) return;
    if (grain >= f.bins) return;
    add((px + py * (uint64_t)f.width) * f.bins + grain, increment, weight);
}

// ---------------------------------------------------------------- camera (cameras.rs:70-97)
PYR_HD void camera_ray(const CameraRec& cam, float tx, float ty, Rng& rng, v3& origin_out, v3& dir_out) {
    float focus_x = tx / cam.view_plane * cam.focus_distance;
    float focus_y = ty / cam.view_plane * cam.focus_distance;
    v3 target = mk3(focus_x, -focus_y, -cam.focus_distance);
    v3 origin = mk3(0, 0, 0), direction = target;
    if (cam.aperture > 0.0f) {
        float sqrt_r = sqrtf(cam.aperture * rng.gen_f32());
        float psi = PYR_PI * 2.0f * rng.gen_f32();
        float sn, cs;
        m_sincos(psi, sn, cs);
        origin = mk3(sqrt_r * cs, sqrt_r * sn, 0.0f);
        direction = target - origin;
    }
    origin_out = transform_point(cam.m, origin);
    dir_out = transform_vector(cam.m, normalize(direction));
}
// Film::sample_many_wavelengths (film.rs:68-83) + the hero pick / swap_remove (simple.rs:105-107)
template <class W>
PYR_HD uint32_t sample_wavelengths(const SceneView& sc, Rng& rng, W wl) {
    const uint32_t S = sc.renderer.spectrum_samples;
    float step_size = sc.film.wavelength_width / (float)S;
    float from = sc.film.wavelength_start;
    for (uint32_t i = 0; i < S; ++i) {
        float to = from + step_size;
        wl[i] = rng.gen_range_f32(from, to);
        from = to;
    }
    return (uint32_t)rng.gen_range_usize(S);
}
template <class W>
PYR_HD void hero_first(W wl, uint32_t S, uint32_t pick) {
    float hero = wl[pick];
    wl[pick] = wl[S - 1];             // swap_remove
    for (uint32_t i = S - 1; i > 0; --i) wl[i] = wl[i - 1];
    wl[0] = hero;
}
PYR_HD Ray make_ray(v3 o, v3 d, uint32_t mode, float limit) {
    Ray r;
    r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z; r.mode = mode;
    r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z; r.limit = limit;
    return r;
}
// Start of one `render_tile` iteration (simple.rs:87-107): path sample `sample` of tile `tile`.
PYR_HD void generate_simple(const SceneView& sc, uint64_t seed, uint32_t tile, uint64_t sample, PathState& ps, Ray& ray) {
    const TileRec t = sc.tiles[tile];
    Rng rng = keyed_rng(seed, t.index, sample);
    float ox = t.size[0] * rng.gen_f32();
    float oy = t.size[1] * rng.gen_f32();
    ps.pos[0] = t.from[0] + ox;
    ps.pos[1] = t.from[1] + oy;
    v3 o, d;
    camera_ray(sc.camera, ps.pos[0], ps.pos[1], rng, o, d);
    const uint32_t S = sc.renderer.spectrum_samples;
    uint32_t pick = sample_wavelengths(sc, rng, ps.wl);
    hero_first(ps.wl, S, pick);
    for (uint32_t k = 0; k < S; ++k) { ps.bright[k] = 0.0f; ps.refl[k] = 1.0f; }
    ps.rng = rng;
    ps.tile = tile;
    ps.flags = PS_USE_ADDITIONAL | PS_SAMPLE_LIGHT | PS_HAS_MAIN;
    ps.bounce = 0; ps.light_events = 0; ps.n_pending = 0; ps.ray_base = 0; ps.pending_brdf = 1.0f;
    ray = make_ray(o, d, 0, 0.0f);
}

// ---------------------------------------------------------------- surface data (shapes/mod.rs:346-405, 454-469, 484-495)
struct Surface {
    v3 position, normal;
    f4 frame;  // Normal::from_space, only valid when `want_frame` was set
    float tex[2];
    uint32_t material;
};
PYR_HD void triangle_surface(const SceneView& sc, uint32_t rank, float u, float v, Surface& s) {
    const TriShade ts = sc.tri_shade[rank];
    s.material = ts.material;
    const bool want_frame = sc.tri_frames != nullptr && sc.materials[ts.material].normal_map_program >= 0;
    float w = 1.0f - (u + v);
    s.normal = normalize((ld3(ts.n1) * w + ld3(ts.n2) * u) + ld3(ts.n3) * v);
    s.tex[0] = (ts.t1[0] * w + ts.t2[0] * u) + ts.t3[0] * v;
    s.tex[1] = (ts.t1[1] * w + ts.t2[1] * u) + ts.t3[1] * v;
    if (want_frame) {
        const TriFrames tf = sc.tri_frames[rank];
        s.frame = qnormalize(add4(add4(scale4(tf.q1, w), scale4(tf.q2, u)), scale4(tf.q3, v)));
    }
}
PYR_HD void sphere_surface(const Prim& pr, v3 position, bool want_frame, Surface& s) {
    v3 normal = normalize(position - prim_v1(pr));
    float latitude = m_acos(normal.y);
    float longitude = m_atan2(normal.x, normal.z);
    s.normal = normal;
    float tx = longitude * PYR_FRAC_1_PI * 0.5f, ty = 1.0f - (latitude * PYR_FRAC_1_PI);
    s.tex[0] = tx / pr.b.x;
    s.tex[1] = ty / pr.b.y;
    if (want_frame) {
        // Matrix3::from_angle_y(longitude) * Matrix3::from_angle_x(latitude - pi/2)
        float sy, cy, sx, cx;
        m_sincos(longitude, sy, cy);
        float a = latitude - PYR_PI * 0.5f;
        m_sincos(a, sx, cx);
        v3 yc0 = mk3(cy, 0, -sy), yc1 = mk3(0, 1, 0), yc2 = mk3(sy, 0, cy);     // columns of Ry
        v3 xc0 = mk3(1, 0, 0), xc1 = mk3(0, cx, sx), xc2 = mk3(0, -sx, cx);     // columns of Rx
        v3 r0 = mk3(yc0.x, yc1.x, yc2.x), r1 = mk3(yc0.y, yc1.y, yc2.y), r2 = mk3(yc0.z, yc1.z, yc2.z);  // rows of Ry
        v3 c0 = mk3(dot(r0, xc0), dot(r1, xc0), dot(r2, xc0));
        v3 c1 = mk3(dot(r0, xc1), dot(r1, xc1), dot(r2, xc1));
        v3 c2 = mk3(dot(r0, xc2), dot(r1, xc2), dot(r2, xc2));
        s.frame = quat_from_cols(c0, c1, c2);
    }
}
PYR_HD void plane_surface(const PlaneRec& pl, v3 position, Surface& s) {
    s.normal = ld3(pl.n);
    s.frame = pl.from_space;
    v3 ns = qrotate(qconj(pl.from_space), position);
    s.tex[0] = ns.x / pl.texture_scale[0];
    s.tex[1] = ns.y / pl.texture_scale[1];
    s.material = pl.material;
}
PYR_HD void marched_surface(const MarchedRec& m, v3 offset_position, bool want_frame, Surface& s, uint32_t& evals, uint32_t& iters) {
    const float E = DIST_EPSILON;
    v3 p = offset_position;
    float xp = estimate_distance(m, p + mk3(E, 0, 0), iters), xn = estimate_distance(m, p + mk3(-E, 0, 0), iters);
    float yp = estimate_distance(m, p + mk3(0, E, 0), iters), yn = estimate_distance(m, p + mk3(0, -E, 0), iters);
    float zp = estimate_distance(m, p + mk3(0, 0, E), iters), zn = estimate_distance(m, p + mk3(0, 0, -E), iters);
    evals += 6;
    s.normal = normalize(mk3(xp - xn, yp - yn, zp - zn));
    s.tex[0] = 0; s.tex[1] = 0;
    if (want_frame) {
        v3 x, y;
        basis(s.normal, x, y);
        s.frame = quat_from_cols(x, y, s.normal);
    }
}
// SurfacePoint::get_surface_data + get_material for a closest-hit record
PYR_HD void hit_surface(const SceneView& sc, v3 o, v3 d, const Hit& h, Surface& s, uint32_t& evals, uint32_t& iters) {
    if (h.kind == KIND_PLANE) {
        const PlaneRec pl = sc.planes[h.rank];
        float t; v3 p;
        plane_test(pl, o, d, t, p);
        s.position = p;
        plane_surface(pl, p, s);
        return;
    }
    if (h.kind == KIND_TRIANGLE) {   // the triangle's shading record carries its material: no Prim fetch
        s.position = o + d * h.t;
        triangle_surface(sc, h.rank, h.u, h.v, s);
        return;
    }
    const Prim pr = sc.prims[h.rank];
    s.material = prim_material(pr);
    bool want_frame = sc.materials[s.material].normal_map_program >= 0;
    if (h.kind == KIND_SPHERE) {
        float t; v3 p;
        sphere_test(prim_v1(pr), pr.a.w, o, d, t, p);
        s.position = p;
        sphere_surface(pr, p, want_frame, s);
    } else {
        const MarchedRec m = sc.marched[f_bits(pr.a.x)];
        s.position = o + d * h.t;
        v3 origin = o + (-bounds_center(m));
        marched_surface(m, origin + d * (h.t - DIST_EPSILON), want_frame, s, evals, iters);
    }
}
// Material::apply_normal_map (materials/mod.rs:68-81)
PYR_HD v3 apply_normal_map(const SceneView& sc, const Surface& s, v3 incident, RegFile R) {
    int32_t prog = sc.materials[s.material].normal_map_program;
    if (prog < 0) return s.normal;
    VmInputs in;
    in.wavelength = 0.0f; in.normal = s.normal; in.incident = incident; in.tex[0] = s.tex[0]; in.tex[1] = s.tex[1];
    f4 v = run_vector(sc, prog, in, R);
    return normalize(qrotate(s.frame, mk3(v.x, v.y, v.z)));
}
// MaterialComponent::get_probability (materials/mod.rs:237-249); `used` = ProbabilityInput::wavelength_used
PYR_HD float component_probability(const SceneView& sc, const ComponentRec& c, const VmInputs& in, RegFile R, bool& used) {
    used = false;
    if (c.probability_program >= 0) {
        used = program_reads_wavelength(sc, c.probability_program);
        return run_number(sc, c.probability_program, in, R) * c.selection_compensation;
    }
    return c.selection_compensation;
}

// ---------------------------------------------------------------- BSDF scatter (materials/{diffuse,mirror,refractive}.rs)
struct Scatter { bool emitted, dispersed, has_brdf; v3 out; float probability; };
PYR_HD Scatter scatter(const ComponentRec& c, v3 in_direction, v3 normal, float wavelength, Rng& rng) {
    Scatter s;
    s.emitted = false; s.dispersed = false; s.has_brdf = false; s.probability = 1.0f; s.out = mk3(0, 0, 0);
    switch (c.bsdf) {
        case BSDF_EMISSIVE: s.emitted = true; return s;
        case BSDF_DIFFUSE: {
            v3 n = dot(in_direction, normal) < 0.0f ? normal : -normal;
            s.out = sample_hemisphere(rng, n);
            s.has_brdf = true;
            return s;
        }
        case BSDF_MIRROR: {
            v3 n = dot(in_direction, normal) < 0.0f ? normal : -normal;
            float perp = dot(in_direction, n) * 2.0f;
            n = n * perp;
            s.out = in_direction - n;
            return s;
        }
        default: {
            s.dispersed = c.dispersion != 0.0f || c.env_dispersion != 0.0f;
            float ior = c.ior, env_ior = c.env_ior;
            if (s.dispersed) {
                float wl = wavelength * 0.001f;
                ior = c.ior + c.dispersion / (wl * wl);
                env_ior = c.env_ior + c.env_dispersion / (wl * wl);
            }
            v3 nl = dot(normal, in_direction) < 0.0f ? normal : -normal;
            v3 reflected = in_direction - (normal * 2.0f * dot(normal, in_direction));
            bool into = dot(normal, nl) > 0.0f;
            float nnt = into ? env_ior / ior : ior / env_ior;
            float ddn = dot(in_direction, nl);
            float cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
            if (cos2t < 0.0f) { s.out = reflected; s.probability = 1.0f; return s; }
            float sgn = (into ? 1.0f : -1.0f) * (ddn * nnt + sqrtf(cos2t));
            v3 tdir = normalize(in_direction * nnt - normal * sgn);
            float a = ior - env_ior, b = ior + env_ior;
            float r0 = a * a / (b * b);
            float cc = 1.0f - (into ? -ddn : dot(tdir, normal));
            float re = r0 + (1.0f - r0) * cc * cc * cc * cc * cc;
            float tr = 1.0f - re;
            float pp = 0.25f + 0.5f * re;
            float rp = re / pp;
            float tp = tr / (1.0f - pp);
            if (rng.gen_f32() < pp) { s.out = reflected; s.probability = rp; }
            else { s.out = tdir; s.probability = tp; }
            return s;
        }
    }
}

// ---------------------------------------------------------------- lamps (lamp.rs:23-113, shapes/mod.rs:166-288)
struct LampSurface { bool physical; v3 normal; float tex[2]; uint32_t material; int32_t color_program; };
struct LampSample { v3 direction; bool has_sq; float sq_distance; LampSurface surface; float weight; };

PYR_HD float prim_surface_area(const SceneView& sc, const Prim& pr, uint32_t rank) {
    if (prim_kind(pr) == KIND_SPHERE) return pr.a.w * pr.a.w * 4.0f * PYR_PI;
    (void)sc; (void)rank;
    return 0.5f * length(cross(prim_e1(pr), prim_e2(pr)));  // 0.5 * |cross(v2 - v1, v3 - v1)|, shapes/mod.rs:279-283
}
// Shape::sample_point (shapes/mod.rs:166-207): position + what get_surface_data needs
PYR_HD void prim_sample_point(const Prim& pr, Rng& rng, v3& position, float& u_out, float& v_out) {
    if (prim_kind(pr) == KIND_SPHERE) {
        v3 s = sample_sphere(rng);
        position = prim_v1(pr) + s * pr.a.w;
        u_out = 0; v_out = 0;
        return;
    }
    float u = rng.gen_f32();
    float v = rng.gen_f32();
    if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
    position = (prim_v1(pr) + prim_e1(pr) * u) + prim_e2(pr) * v;
    u_out = u; v_out = v;
}
PYR_HD void prim_point_surface(const SceneView& sc, const Prim& pr, uint32_t rank, v3 position, float u, float v, Surface& s) {
    s.position = position;
    s.material = prim_material(pr);
    if (prim_kind(pr) == KIND_SPHERE) sphere_surface(pr, position, false, s);
    else triangle_surface(sc, rank, u, v, s);
}
// Lamp::sample (lamp.rs:23-82)
PYR_HD LampSample lamp_sample(const SceneView& sc, const LampRec& lamp, Rng& rng, v3 target) {
    LampSample out;
    out.surface.physical = false; out.surface.normal = mk3(0, 0, 0); out.surface.tex[0] = 0; out.surface.tex[1] = 0;
    out.surface.material = 0; out.surface.color_program = lamp.color_program;
    out.has_sq = false; out.sq_distance = 0.0f;
    if (lamp.kind == LAMP_DIRECTIONAL) {
        v3 dir = ld3(lamp.v);
        out.direction = lamp.width > 0.0f ? sample_cone(rng, dir, lamp.width) : dir;
        out.weight = 1.0f;
        return out;
    }
    if (lamp.kind == LAMP_POINT) {
        v3 v = ld3(lamp.v) - target;
        float distance = length2(v);
        out.direction = normalize(v);
        out.has_sq = true; out.sq_distance = distance;
        out.weight = 4.0f * PYR_PI / distance;
        return out;
    }
    const Prim pr = sc.prims[lamp.rank];
    const bool is_sphere = prim_kind(pr) == KIND_SPHERE;
    // Shape::sample_towards (shapes/mod.rs:209-251)
    float hit_distance = 0.0f, u = 0.0f, v = 0.0f;
    v3 position = target;
    bool done = false;
    if (is_sphere) {
        float r = fmaxf(pr.a.w - DIST_EPSILON, 0.0f);
        v3 dir = prim_v1(pr) - target;
        float dist2 = length2(dir);
        if (dist2 > r * r) {
            float cos_theta_max = sqrtf(fmaxf(1.0f - (r * r) / dist2, 0.0f));
            v3 ray_dir = sample_cone(rng, normalize(dir), cos_theta_max);
            float t; v3 p;
            if (sphere_test(prim_v1(pr), pr.a.w, target, ray_dir, t, p)) { hit_distance = t; position = p; }
            else { hit_distance = 0.0f; position = target; }  // the reference's "cheat" branch
            done = true;
        }
    }
    if (!done) {
        prim_sample_point(pr, rng, position, u, v);
        hit_distance = length(position - target);
    }
    v3 vec = position - target;
    float sq_distance = hit_distance * hit_distance;
    v3 direction = normalize(vec);
    Surface s;
    prim_point_surface(sc, pr, lamp.rank, position, u, v, s);
    // Shape::solid_angle_towards (shapes/mod.rs:253-271), else cos_in * area / d^2
    float weight;
    bool have = false;
    if (is_sphere) {
        float dist2 = length2(prim_v1(pr) - target);
        if (dist2 > pr.a.w * pr.a.w) {
            float cos_theta_max = sqrtf(fmaxf(1.0f - (pr.a.w * pr.a.w) / dist2, 0.0f));
            weight = solid_angle(cos_theta_max);
            have = true;
        }
    }
    if (!have) {
        float cos_in = fabsf(dot(s.normal, -direction));
        weight = cos_in * prim_surface_area(sc, pr, lamp.rank) / sq_distance;
    }
    out.direction = direction;
    out.has_sq = true; out.sq_distance = sq_distance;
    out.surface.physical = true; out.surface.normal = s.normal; out.surface.tex[0] = s.tex[0]; out.surface.tex[1] = s.tex[1];
    out.surface.material = s.material; out.surface.color_program = -1;
    out.weight = weight;
    return out;
}

// ---------------------------------------------------------------- the contribute fold (renderer/algorithm.rs:14-100)
// values[k] = color(wl[k]) for k < n: the program record is fetched once and the memoised re-run
// is used for k > 0
template <class W, class Sink>
PYR_HD void eval_spectral_each(const SceneView& sc, int32_t color, const VmInputs& base, W wl, uint32_t n, RegFile R, Sink&& sink) {
    const ProgramRec p = sc.programs[color];
    if (p.is_constant) { for (uint32_t k = 0; k < n; ++k) sink(k, p.value); return; }
    VmInputs in = base;
    for (uint32_t k = 0; k < n; ++k) {
        in.wavelength = wl[k];
        sink(k, run_program(sc, p, in, R, k > 0));
    }
}
template <class W>
PYR_HD void eval_spectral(const SceneView& sc, int32_t color, const VmInputs& base, W wl, uint32_t n, float* values, RegFile R) {
    eval_spectral_each(sc, color, base, wl, n, R, [&](uint32_t k, float v) { values[k] = v; });
}
// brightness[k] += color(wl[k]) * probability * reflectance[k] for k < n
PYR_HD void add_emission(const SceneView& sc, PathState& ps, uint32_t n, int32_t color, v3 incident, v3 normal, const float* tex,
                         float probability, RegFile R) {
    VmInputs in;
    in.wavelength = 0.0f; in.incident = incident; in.normal = normal; in.tex[0] = tex[0]; in.tex[1] = tex[1];
    eval_spectral_each(sc, color, in, ps.wl, n, R, [&](uint32_t k, float c) { ps.bright[k] += c * probability * ps.refl[k]; });
}
PYR_HD void mul_reflectance(const SceneView& sc, PathState& ps, uint32_t n, int32_t color, v3 incident, v3 normal, const float* tex,
                            float probability, RegFile R) {
    VmInputs in;
    in.wavelength = 0.0f; in.incident = incident; in.normal = normal; in.tex[0] = tex[0]; in.tex[1] = tex[1];
    eval_spectral_each(sc, color, in, ps.wl, n, R, [&](uint32_t k, float c) { ps.refl[k] *= c * probability; });
}

struct PathCounters { uint32_t de_evals, de_iters; };

// trace_direct (tracer.rs:347-442): draws the lamp samples of one next-event estimation and emits
// their visibility rays; the `DirectLight` records wait in `ps.pend` until the rays are traced.
// Draw order is the reference's, except that the emissive-component pick of a shape lamp happens
// before (and regardless of) the visibility result (the oracle has the same switch).
PYR_HD void next_event(const SceneView& sc, PathState& ps, float wavelength, v3 ray_in, v3 position, v3 normal, ShadeOut& out, RegFile R) {
    const uint32_t samples = sc.renderer.light_samples;
    if (sc.n_lamps == 0) return;  // unreachable: pyr_render refuses such scenes (the reference would panic here)
    const uint32_t lamp_index = (uint32_t)ps.rng.gen_range_usize(sc.n_lamps);  // World::pick_lamp (world.rs:301-305)
    const float lamp_probability = 1.0f / (float)sc.n_lamps;
    const LampRec lamp = sc.lamps[lamp_index];
    if (dot(ray_in, normal) >= 0.0f) normal = -normal;
    float probability = 1.0f / ((float)samples * 2.0f * PYR_PI * lamp_probability);
    for (uint32_t k = 0; k < samples; ++k) {
        LampSample ls = lamp_sample(sc, lamp, ps.rng, position);
        float cos_out = fmaxf(dot(normal, ls.direction), 0.0f);
        if (!(cos_out > 0.0f)) continue;
        PendingLight pl;
        float material_probability = 1.0f;
        if (ls.surface.physical) {
            const MaterialRec m = sc.materials[ls.surface.material];
            uint32_t ci = ps.rng.gen_index_u32(m.n_emissive);
            const ComponentRec comp = sc.components[m.emissive_offset + ci];
            VmInputs in;
            in.wavelength = wavelength; in.normal = ls.surface.normal; in.incident = ls.direction;
            in.tex[0] = ls.surface.tex[0]; in.tex[1] = ls.surface.tex[1];
            bool used;
            material_probability = component_probability(sc, comp, in, R, used);
            pl.color_program = comp.color_program; pl.dispersed = used ? 1u : 0u;
            pl.normal[0] = ls.surface.normal.x; pl.normal[1] = ls.surface.normal.y; pl.normal[2] = ls.surface.normal.z;
            pl.tex[0] = ls.surface.tex[0]; pl.tex[1] = ls.surface.tex[1];
        } else {
            pl.color_program = ls.surface.color_program; pl.dispersed = 0u;
            pl.normal[0] = -ls.direction.x; pl.normal[1] = -ls.direction.y; pl.normal[2] = -ls.direction.z;
            pl.tex[0] = 0; pl.tex[1] = 0;
        }
        float scale = ls.weight * probability * (2.0f * fabsf(dot(ls.direction, normal)));  // lambertian(ray_in, normal, ray_out)
        pl.probability = scale * material_probability;
        uint32_t j = out.n_shadow++;
        store_record_stream(ps.pend + j, pl);
        // blocked <=> a hit with t > eps and t^2 < sq_distance - eps (tracer.rs:381-389); lamps without a
        // distance (directional) are blocked by any hit
        out.put_shadow(j, make_ray(position, ls.direction, 1, ls.has_sq ? ls.sq_distance - DIST_EPSILON : PYR_INF));
    }
}

// trace_directional (tracer.rs:444-459)
PYR_HD int32_t directional_color(const SceneView& sc, v3 ray, int32_t fallback) {
    for (uint32_t i = 0; i < sc.n_lamps; ++i) {
        const LampRec l = sc.lamps[i];
        if (l.kind == LAMP_DIRECTIONAL && dot(ld3(l.v), ray) >= l.width) return l.color_program;
    }
    return fallback;
}

// The end of a `render_tile` iteration (simple.rs:127-139): expose the hero sample and, unless a
// dispersive event was seen, the additional wavelengths.
template <class Add>
PYR_HD void expose_path(const SceneView& sc, const PathState& ps, Add& add) {
    film_expose(sc.film, ps.pos[0], ps.pos[1], ps.bright[0], ps.wl[0], 1.0f, add);
    if (ps.flags & PS_USE_ADDITIONAL)
        for (uint32_t k = 1; k < sc.renderer.spectrum_samples; ++k) film_expose(sc.film, ps.pos[0], ps.pos[1], ps.bright[k], ps.wl[k], 1.0f, add);
}

// One wavefront step of the camera-to-light integrator for one path sample: fold the previous
// bounce's direct light (its visibility rays are now traced), then process the closest hit of the
// path ray exactly as one iteration of `trace`'s loop (tracer.rs:221-343) followed by
// `contribute` for that bounce (algorithm.rs:14-100).  `main_ray` / `main_hit`: the path ray of the
// finished trace pass (valid if PS_HAS_MAIN); `shadow_rays` / `shadow_kinds`: its `n_pending`
// visibility rays.
//
// `hooks` lets the bidirectional integrator observe the camera subpath it shares with this one
// (bidirectional.rs:205 calls the same `trace`): every pushed `Bounce`, and the moment the
// `contribute` of a bounce is complete.  Returns false when the path has ended.
struct NoHooks {
    PYR_HD void contribute_done(PathState&) {}
    PYR_HD void pushed_emission(PathState&) {}
    PYR_HD void pushed_surface(PathState&, bool, v3, v3, float) {}
};
template <class Hooks>
PYR_HD bool camera_step(const SceneView& sc, PathState& ps, const Ray* main_ray, const Hit* main_hit, const Ray* shadow_rays,
                        const uint32_t* shadow_kinds, ShadeOut& out, PathCounters& pc, Hooks& hooks) {
    PYR_REGFILE(R);
    const uint32_t S = sc.renderer.spectrum_samples;
    out.alive = 0; out.has_main = 0; out.n_shadow = 0;

    if (ps.flags & PS_PENDING_FOLD) {
        const uint32_t n = (ps.flags & PS_USE_ADDITIONAL) ? S : 1u;
        // The light samples of one event come from one lamp; when its colour program reads nothing but
        // the wavelength, colour(wl[k]) is the same for all of them and is evaluated once.
        int32_t cached_program = -1;
        uint32_t cached_n = 0;
#if defined(__CUDA_ARCH__)
        // [wavelength][thread] floats behind the staged visibility rays
        float* const c_base = reinterpret_cast<float*>(pyr_dyn_smem + out.stage_base + stage_quads(sc.renderer.light_samples) * PYR_BLOCK) + threadIdx.x;
#define PYR_C(k) c_base[(k) * PYR_BLOCK]
#else
        float c_local[MAX_SPECTRUM_SAMPLES];
#define PYR_C(k) c_local[k]
#endif
        for (uint32_t j = 0; j < ps.n_pending; ++j) {
            if (shadow_kinds[j] != KIND_MISS) continue;  // blocked
            const PendingLight pl = load_record_stream(ps.pend + j);
            const uint32_t m = pl.dispersed ? 1u : n;
            if (pl.color_program != cached_program || m > cached_n) {
                const ProgramRec p = sc.programs[pl.color_program];
                const bool wavelength_only = p.is_constant || !(p.reads & (IN_NORMAL | IN_INCIDENT | IN_TEXTURE));
                VmInputs in;
                in.wavelength = 0.0f; in.normal = ld3(pl.normal); in.tex[0] = pl.tex[0]; in.tex[1] = pl.tex[1];
                in.incident = wavelength_only ? mk3(0, 0, 0) : ld3(load_record_stream(shadow_rays + j).d);  // the ray is only fetched when the colour looks at it
                eval_spectral_each(sc, pl.color_program, in, ps.wl, m, R, [&](uint32_t k, float v) { PYR_C(k) = v; });
                cached_program = wavelength_only ? pl.color_program : -1;
                cached_n = m;
            }
            for (uint32_t k = 0; k < m; ++k) ps.bright[k] += PYR_C(k) * pl.probability * ps.refl[k];
        }
#undef PYR_C
        for (uint32_t k = 0; k < n; ++k) ps.refl[k] *= ps.pending_brdf;
        ps.flags &= ~PS_PENDING_FOLD;
        ps.n_pending = 0;
        hooks.contribute_done(ps);
    }
    if (!(ps.flags & PS_HAS_MAIN)) return false;

    const Ray incoming = load_record_stream(main_ray);
    const v3 o = ld3(incoming.o), d = ld3(incoming.d);
    const Hit h = load_record(main_hit);
    const float wavelength = ps.wl[0];
    if (h.kind == KIND_MISS) {  // tracer.rs:322-342
        int32_t color = sc.sky_program;
        if (ps.flags & PS_SAMPLE_LIGHT) color = directional_color(sc, d, color);
        const float tex0[2] = {0.0f, 0.0f};
        add_emission(sc, ps, (ps.flags & PS_USE_ADDITIONAL) ? S : 1u, color, d, -d, tex0, 1.0f, R);
        hooks.pushed_emission(ps);
        return false;
    }
    Surface s;
    hit_surface(sc, o, d, h, s, pc.de_evals, pc.de_iters);
    const v3 normal = apply_normal_map(sc, s, d, R);
    const MaterialRec mat = sc.materials[s.material];
    const ComponentRec comp = sc.components[mat.comp_offset + ps.rng.gen_index_u32(mat.n_components)];  // choose_component
    VmInputs pin;
    pin.wavelength = wavelength; pin.normal = normal; pin.incident = d; pin.tex[0] = s.tex[0]; pin.tex[1] = s.tex[1];
    bool normal_dispersed;
    const float component_prob = component_probability(sc, comp, pin, R, normal_dispersed);
    const Scatter sct = scatter(comp, d, normal, wavelength, ps.rng);
    if (sct.emitted) {
        if (ps.flags & PS_SAMPLE_LIGHT) {
            if (normal_dispersed) ps.flags &= ~PS_USE_ADDITIONAL;
            add_emission(sc, ps, (ps.flags & PS_USE_ADDITIONAL) ? S : 1u, comp.color_program, d, normal, s.tex, component_prob, R);
            hooks.pushed_emission(ps);
        }
        return false;
    }
    if (ps.light_events < 2) {
        if (!sct.has_brdf || sc.renderer.light_samples == 0) ps.flags |= PS_SAMPLE_LIGHT; else ps.flags &= ~PS_SAMPLE_LIGHT;
        if (sct.has_brdf) {
            ps.light_events += 1;
            next_event(sc, ps, wavelength, d, s.position, normal, out, R);
        }
    } else {
        ps.flags |= PS_SAMPLE_LIGHT;
    }
    if (sct.dispersed || normal_dispersed) ps.flags &= ~PS_USE_ADDITIONAL;
    mul_reflectance(sc, ps, (ps.flags & PS_USE_ADDITIONAL) ? S : 1u, comp.color_program, d, normal, s.tex, sct.probability * component_prob, R);
    ps.pending_brdf = sct.has_brdf ? 2.0f * fabsf(dot(sct.out, normal)) : 1.0f;  // materials/diffuse.rs:27-29
    ps.flags |= PS_PENDING_FOLD;
    ps.n_pending = out.n_shadow;
    ps.bounce += 1;
    hooks.pushed_surface(ps, sct.has_brdf, s.position, normal, ps.pending_brdf);
    if (ps.bounce < sc.renderer.bounces) {
        ps.flags |= PS_HAS_MAIN;
        out.has_main = 1;
        out.main = make_ray(s.position, sct.out, 0, 0.0f);
    } else {
        ps.flags &= ~PS_HAS_MAIN;
    }
    if (out.has_main || out.n_shadow) { out.alive = 1; return true; }
    // out of bounces with nothing pending: this bounce's `contribute` is complete now
    for (uint32_t k = 0; k < ((ps.flags & PS_USE_ADDITIONAL) ? S : 1u); ++k) ps.refl[k] *= ps.pending_brdf;
    ps.flags &= ~PS_PENDING_FOLD;
    hooks.contribute_done(ps);
    return false;
}

// The camera-to-light integrator (renderer/simple.rs:78-140): when the path ends, expose it.
template <class Add>
PYR_HD void shade_simple(const SceneView& sc, PathState& ps, const Ray* main_ray, const Hit* main_hit, const Ray* shadow_rays,
                         const uint32_t* shadow_kinds, ShadeOut& out, Add& add, PathCounters& pc) {
    NoHooks hooks;
    if (!camera_step(sc, ps, main_ray, main_hit, shadow_rays, shadow_kinds, out, pc, hooks)) expose_path(sc, ps, add);
}


// ---------------------------------------------------------------- diagnostic: one path sample, depth-first
// tools/first_divergence.py: one `render_tile` iteration (simple.rs:87-139) for path sample (tile, sample) with the very
// stage functions the wavefront uses (generate_simple, trace_ray, camera_step), recording per bounce what the oracle's
// checker records for the same sample - 20 words: kind, prim_id, t, u, v, incident[3], position[3], normal[3], out[3], visibility
// rays cast, Xorshift `w` after the bounce, 1 if a surface bounce was pushed - and the exposed (brightness, wavelength)
// pairs.  `rays` / `hits` / `kinds` hold 1 + MAX_LIGHT_SAMPLES records (global memory on the device: camera_step moves
// them with ld/st.global); `ps` comes with its per-wavelength arrays and pending-light storage bound.
struct DebugHooks {
    v3 position, normal;
    bool surface;
    PYR_HD void contribute_done(PathState&) {}
    PYR_HD void pushed_emission(PathState&) {}
    PYR_HD void pushed_surface(PathState&, bool, v3 p, v3 n, float) { position = p; normal = n; surface = true; }
};
PYR_HD void debug_path_simple(const SceneView& sc, uint64_t seed, uint32_t tile, uint64_t sample, uint32_t max_bounces, uint32_t* records, uint32_t* counts,
                              float* exposed, float* position2, Ray* rays, Hit* hits, uint32_t* kinds, PathState& ps, ShadeOut& out) {
    generate_simple(sc, seed, tile, sample, ps, out.main);
    ps.flags |= PS_ALIVE;
    out.has_main = 1; out.n_shadow = 0; out.alive = 1;
    position2[0] = ps.pos[0]; position2[1] = ps.pos[1];
    PathCounters pc;
    pc.de_evals = 0; pc.de_iters = 0;
    uint32_t n_bounces = 0;
    bool more = true;
    while (more) {
        const bool had_main = out.has_main != 0;
        if (had_main) { Hit h; trace_ray<false>(sc, out.main, h, nullptr); rays[0] = out.main; hits[0] = h; }
        for (uint32_t j = 0; j < out.n_shadow; ++j) {
            const Ray r = out.get_shadow(j);
            Hit h;
            trace_ray<false>(sc, r, h, nullptr);
            rays[1 + j] = r; hits[1 + j] = h; kinds[1 + j] = h.kind;
        }
        DebugHooks hooks;
        hooks.surface = false;
        hooks.position = mk3(0, 0, 0); hooks.normal = mk3(0, 0, 0);
        more = camera_step(sc, ps, rays, hits, rays + 1, kinds + 1, out, pc, hooks);
        if (had_main) {
            if (n_bounces < max_bounces) {
                uint32_t* r = records + 20 * n_bounces;
                const Hit h = hits[0];
                r[0] = h.kind;
                r[1] = h.kind == KIND_MISS ? 0xFFFFFFFFu : (h.kind == KIND_PLANE ? h.rank : prim_object(sc.prims[h.rank]));
                r[2] = f_bits(h.kind == KIND_MISS ? 0.0f : h.t); r[3] = f_bits(h.kind == KIND_MISS ? 0.0f : h.u); r[4] = f_bits(h.kind == KIND_MISS ? 0.0f : h.v);
                r[5] = f_bits(rays[0].d[0]); r[6] = f_bits(rays[0].d[1]); r[7] = f_bits(rays[0].d[2]);
                r[8] = f_bits(hooks.position.x); r[9] = f_bits(hooks.position.y); r[10] = f_bits(hooks.position.z);
                r[11] = f_bits(hooks.normal.x); r[12] = f_bits(hooks.normal.y); r[13] = f_bits(hooks.normal.z);
                const bool next = out.has_main != 0;
                r[14] = f_bits(next ? out.main.d[0] : 0.0f); r[15] = f_bits(next ? out.main.d[1] : 0.0f); r[16] = f_bits(next ? out.main.d[2] : 0.0f);
                r[17] = out.n_shadow; r[18] = ps.rng.w; r[19] = hooks.surface ? 1u : 0u;
            }
            ++n_bounces;
        }
    }
    const uint32_t n = (ps.flags & PS_USE_ADDITIONAL) ? sc.renderer.spectrum_samples : 1u;
    for (uint32_t k = 0; k < n; ++k) { exposed[2 * k] = ps.bright[k]; exposed[2 * k + 1] = ps.wl[k]; }
    counts[0] = n_bounces; counts[1] = n;
}

// ---------------------------------------------------------------- develop (main.rs:190-238, 313-418)
struct DevelopParams { float white_max, d65_max, step_size; uint32_t pad; };

PYR_HD float d65_get(const SceneView& sc, float w) { return array_get(sc.d65, sc.d65_t.n, 1, sc.d65_t.lo, sc.d65_t.hi, w); }
// the white-balance scan of main.rs:206-214
PYR_HD void white_scan(const SceneView& sc, float& white_max, float& d65_max) {
    PYR_REGFILE(R);
    white_max = 0.0f; d65_max = 0.0f;
    if (sc.white_program < 0) return;
    float wavelength = sc.renderer.span_lo;
    VmInputs in;
    in.normal = mk3(0, 0, 0); in.incident = mk3(0, 0, 0); in.tex[0] = 0; in.tex[1] = 0;
    while (wavelength < sc.renderer.span_hi) {
        in.wavelength = wavelength;
        white_max = fmaxf(white_max, run_number(sc, sc.white_program, in, R));
        d65_max = fmaxf(d65_max, d65_get(sc, wavelength));
        wavelength += 1.0f;
    }
}
// the spectrum_get closure of main.rs:224-238
PYR_HD float develop_adjust(const SceneView& sc, const DevelopParams& dp, float intensity, float wavelength, RegFile R) {
    VmInputs in;
    in.wavelength = wavelength; in.normal = mk3(0, 0, 0); in.incident = mk3(0, 0, 0); in.tex[0] = 0; in.tex[1] = 0;
    float filtered = sc.filter_program >= 0 ? intensity * run_number(sc, sc.filter_program, in, R) : intensity;
    if (sc.white_program >= 0) {
        float white_intensity = run_number(sc, sc.white_program, in, R) / dp.white_max;
        float neutral = filtered / fmaxf(white_intensity, 0.000001f);
        return neutral * (d65_get(sc, wavelength) / dp.d65_max);
    }
    return filtered;
}
// film.rs:321-337 `Spectrum::get` + Grain::develop (:132-143) for one pixel; film = (accumulator, weight) pairs
PYR_HD float film_spectrum_get(const SceneView& sc, const float* film, uint64_t pixel, float w) {
    float lo = sc.film.wavelength_start, hi = sc.film.wavelength_start + sc.film.wavelength_width;
    if (w < lo) return 0.0f;
    if (w > hi) return 0.0f;
    float normalized = (w - lo) / (hi - lo);
    float float_index = normalized * (float)sc.film.bins;
    uint64_t index = f32_as_usize(floorf(float_index));
    if (index > sc.film.bins - 1) index = sc.film.bins - 1;
    const float* g = film + 2 * (pixel * sc.film.bins + index);
    return g[1] > 0.0f ? g[0] / g[1] : 0.0f;
}
// spectrum_to_xyz / spectrum_to_tristimulus (main.rs:352-418), x 3.444 (main.rs:322)
PYR_HD void pixel_to_xyz(const SceneView& sc, const DevelopParams& dp, const float* film, uint64_t pixel, float* out) {
    PYR_REGFILE(R);
    float lo = sc.film.wavelength_start, hi = sc.film.wavelength_start + sc.film.wavelength_width;
    float sum[3] = {0, 0, 0};
    float weight = 0.0f;
    float wl_min = lo;
    float spectrum_min = develop_adjust(sc, dp, film_spectrum_get(sc, film, pixel, wl_min), wl_min, R);
    float start_resp[3];
    for (int c = 0; c < 3; ++c) start_resp[c] = array_get(sc.xyz + c, sc.xyz_t.n, 3, sc.xyz_t.lo, sc.xyz_t.hi, wl_min);
    while (wl_min < hi) {
        float wl_max = wl_min + dp.step_size;
        float spectrum_max = develop_adjust(sc, dp, film_spectrum_get(sc, film, pixel, wl_max), wl_max, R);
        float end_resp[3];
        for (int c = 0; c < 3; ++c) end_resp[c] = array_get(sc.xyz + c, sc.xyz_t.n, 3, sc.xyz_t.lo, sc.xyz_t.hi, wl_max);
        float w = wl_max - wl_min;
        for (int c = 0; c < 3; ++c) sum[c] += ((start_resp[c] * spectrum_min + end_resp[c] * spectrum_max) * 0.5f) * w;
        weight += w;
        wl_min = wl_max;
        spectrum_min = spectrum_max;
        for (int c = 0; c < 3; ++c) start_resp[c] = end_resp[c];
    }
    for (int c = 0; c < 3; ++c) out[c] = (weight == 0.0f ? sum[c] : sum[c] / weight) * 3.444f;
}
// palette LinSrgb::from_color(Xyz<D65>) + into_encoding::<Srgb<u8>> (main.rs:323)
PYR_HD void xyz_to_srgb8(const float* xyz, uint8_t* out) {
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

}  // namespace pyr
