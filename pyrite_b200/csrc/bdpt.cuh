// Bidirectional integrator (pyrite/src/renderer/bidirectional.rs:73-398) as a per-path-sample state
// machine for the wavefront pipeline.  One `render_tile` iteration runs through four phases, each
// wavefront iteration tracing the rays the phase asked for:
//
//   PH_LAMP     the lamp subpath: Lamp::sample_ray (lamp.rs:84-113), the emission vertex
//               (bidirectional.rs:128-174) and `trace(.., light_bounces, light_samples = 0)`, one path ray
//               per iteration; vertices are stored in HBM (LightVertex, light_bounces + 1 per path)
//   PH_FINISH   the end of the lamp subpath (fix-up, reversal, colour programs at the path's wavelengths, tail folds): a step of
//               its own, without rays - lamp paths end after different numbers of bounces, and inside the lamp kernel the few
//               finishing lanes of a warp held the others up (4.7 of 32 lanes active in the steady state)
//   PH_CAMERA   the camera subpath: the same `trace` as the camera-to-light integrator (camera_step),
//               folding `contribute` as it goes and storing the state at every diffuse vertex (CamVertex)
//   PH_CONNECT  connect_paths (bidirectional.rs:310-398): for each stored camera vertex, visibility rays to
//               the non-specular lamp vertices, in chunks of at most BDPT_STAGE rays per iteration; the
//               weight 1 / (len(camera_path) * len(lamp_path)) is only known once the camera path has ended,
//               which is why connections are made after it
//   PH_SPLAT    light tracing (bidirectional.rs:253-306): Camera::is_visible (cameras.rs:99-158) for every
//               diffuse lamp vertex, exposing at the projected film position
//
// The RNG draw order of the reference is kept: sample point, wavelengths, hero pick, camera ray, lamp
// pick, lamp ray, emissive component, lamp subpath, camera subpath, then the lens samples of is_visible.
#pragma once
#include "shading.cuh"

namespace pyr {

constexpr int BDPT_STAGE = 16;  // visibility rays staged per path per iteration

enum : uint32_t { PH_LAMP = 0, PH_CAMERA = 1, PH_CONNECT = 2, PH_SPLAT = 3, PH_FINISH = 4 };
enum : uint32_t { VT_DIFFUSE = 0, VT_SPECULAR = 1, VT_EMISSION = 2 };

// One stored vertex of the lamp subpath (tracer.rs:157-167 `Bounce`; lamp paths carry no direct light), plus
// its colour program evaluated at the path's wavelengths: connect_paths and light tracing re-fold the tail of
// the lamp path for every connection (bidirectional.rs:373-389, 276-292), always with the same colours. 144 B.
struct alignas(32) LightVertexHead {   // 96 B: what a thread keeps in registers; the first 32 B are what the staging loops read
    float position[3];
    uint32_t type;
    float normal[3];
    int32_t color_program;
    float incident[3];
    float probability;
    float out[3];         // BounceType::Diffuse(_, out)
    uint32_t dispersed;
    float tex[2];
    uint32_t pad[6];
};
struct alignas(32) LightVertex : LightVertexHead {
    // The fold of `contribute` over lamp_path[i..] as seen by a connection that STARTS at this vertex, per wavelength:
    // connect_paths (bidirectional.rs:373-389) and light tracing (:276-292) re-fold that tail for every connection, always
    // with the same colours, and the fold is linear in the reflectance it starts from -
    //     brightness_out[k] = brightness_in[k] + reflectance_in[k] * fold[k]
    // - so finish_lamp_path evaluates every vertex' colour program once at the path's wavelengths and stores the tail folds
    // (one backward pass); a connection is then ONE 64-byte load instead of a walk over the tail.
    float fold[MAX_SPECTRUM_SAMPLES];
};
constexpr uint32_t VT_TAIL_DISPERSED = 0x100u;  // in `type` after finish_lamp_path: some vertex of lamp_path[i..] is dispersive
// position, type and normal of a lamp vertex: one 32-byte load
struct VertexGeometry { v3 position; uint32_t type; bool tail_dispersed; v3 normal; };
PYR_HD VertexGeometry vertex_geometry(const Vec8& q) {
    VertexGeometry g;
    g.position = mk3(q.v[0], q.v[1], q.v[2]);
    g.type = f_bits(q.v[3]) & 0xffu;
    g.tail_dispersed = (f_bits(q.v[3]) & VT_TAIL_DISPERSED) != 0;
    g.normal = mk3(q.v[4], q.v[5], q.v[6]);
    return g;
}
// A diffuse camera-subpath vertex with the sample state right after its `contribute`. 160 B.
struct alignas(32) CamVertexHead {
    float position[3];
    float brdf;           // bounce.ty.brdf(..) = 2 |out . normal|
    float normal[3];
    uint32_t use_additional;
};
struct alignas(32) CamVertex : CamVertexHead {
    float bright[MAX_SPECTRUM_SAMPLES];
    float refl[MAX_SPECTRUM_SAMPLES];
};

// What one step asks to be traced next.  The visibility rays are not materialised per thread on the device (16 x 32 B of
// thread-local memory): a step only COUNTS them (`n_shadow`, `shadow_kind`), the kernel reserves queue space for the block and
// then writes them straight into the queue (write_staged), re-deriving them from the stored vertices.  On the host (CPU
// emulation) they land in `shadow`.
enum : uint32_t { SH_NONE = 0, SH_NEE = 1, SH_CONNECT = 2, SH_SPLAT = 3 };
struct BidirOut {
    uint32_t alive, has_main, n_shadow, shadow_kind;
    Ray main;
#if !defined(__CUDA_ARCH__)
    Ray shadow[BDPT_STAGE];
#endif
};
#if defined(__CUDA_ARCH__)
#define PYR_STAGE_RAYS(out) static_cast<Ray*>(nullptr)
#else
#define PYR_STAGE_RAYS(out) (out).shadow
#endif

struct BidirCtx {
    LightVertex* lv;   // this path's lamp vertices
    CamVertex* cv;     // this path's stored camera vertices
    SpecArray bright, refl;  // detached sample state of one connection / light-traced sample (shared memory in the kernel)
};

PYR_HD void st3(float* dst, v3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
PYR_HD float vertex_brdf(const LightVertexHead& v) {  // BounceType::brdf: lambertian(_, normal, out) = 2 |out . normal|
    return v.type == VT_DIFFUSE ? 2.0f * fabsf(dot(ld3(v.out), ld3(v.normal))) : 1.0f;
}

PYR_HD LightVertexHead load_head(const LightVertex* p) {
    static_assert(sizeof(LightVertexHead) == 96, "three chunks");
    const Vec8 a = ld256(p), b = ld256(reinterpret_cast<const Vec8*>(p) + 1), c = ld256(reinterpret_cast<const Vec8*>(p) + 2);
    LightVertexHead h;
    __builtin_memcpy(&h, &a, 32);
    __builtin_memcpy(reinterpret_cast<char*>(&h) + 32, &b, 32);
    __builtin_memcpy(reinterpret_cast<char*>(&h) + 64, &c, 32);
    return h;
}
PYR_HD void store_head(LightVertex* p, const LightVertexHead& h) {
    Vec8 a, b, c;
    __builtin_memcpy(&a, &h, 32);
    __builtin_memcpy(&b, reinterpret_cast<const char*>(&h) + 32, 32);
    __builtin_memcpy(&c, reinterpret_cast<const char*>(&h) + 64, 32);
    st256(p, a); st256(reinterpret_cast<Vec8*>(p) + 1, b); st256(reinterpret_cast<Vec8*>(p) + 2, c);
}
// Camera::ray_towards inputs for the lens sample of Camera::is_visible (cameras.rs:122-131)
PYR_HD v3 lens_origin(const CameraRec& cam, Rng& rng) {
    if (cam.aperture > 0.0f) {
        float sqrt_r = sqrtf(cam.aperture * rng.gen_f32());
        float psi = PYR_PI * 2.0f * rng.gen_f32();
        float sn, cs;
        m_sincos(psi, sn, cs);
        return mk3(sqrt_r * cs, sqrt_r * sn, 0.0f);
    }
    return mk3(0, 0, 0);
}

struct CameraHooks {
    BidirCtx cx;
    uint32_t n_spec;  // spectrum_samples
    PYR_HD void contribute_done(PathState& ps) {
        if (!ps.bd->cam_store_pending) return;
        CamVertex& c = cx.cv[ps.bd->n_cam_stored - 1];
        c.use_additional = (ps.flags & PS_USE_ADDITIONAL) ? 1u : 0u;
        Vec8 b[2], r[2];
#pragma unroll
        for (uint32_t k = 0; k < MAX_SPECTRUM_SAMPLES; ++k) {
            b[k >> 3].v[k & 7] = k < n_spec ? ps.bright[k] : 0.0f;
            r[k >> 3].v[k & 7] = k < n_spec ? ps.refl[k] : 0.0f;
        }
        st256(c.bright, b[0]); st256(c.refl, r[0]);
        if (n_spec > 8) { st256(c.bright + 8, b[1]); st256(c.refl + 8, r[1]); }
        ps.bd->cam_store_pending = 0;
    }
    PYR_HD void pushed_emission(PathState& ps) { ps.bd->n_cam += 1; }
    PYR_HD void pushed_surface(PathState& ps, bool diffuse, v3 position, v3 normal, float brdf) {
        ps.bd->n_cam += 1;
        if (!diffuse) return;  // connect_paths returns at once for specular bounces (:320-323)
        CamVertex& c = cx.cv[ps.bd->n_cam_stored++];
        st3(c.position, position); st3(c.normal, normal);
        c.brdf = brdf;
        ps.bd->cam_store_pending = 1;
    }
};

// -------------------------------------------------------------------------------- phase transitions
PYR_HD void begin_camera(PathState& ps, BidirOut& out) {
    ps.bd->phase = PH_CAMERA;
    ps.flags |= PS_HAS_MAIN | PS_SAMPLE_LIGHT | PS_USE_ADDITIONAL;
    ps.flags &= ~PS_PENDING_FOLD;
    ps.bounce = 0; ps.light_events = 0; ps.n_pending = 0; ps.pending_brdf = 1.0f;
    out.has_main = 1;
    out.main = make_ray(ld3(ps.bd->cam_o), ld3(ps.bd->cam_d), 0, 0.0f);
    out.alive = 1;
}

// The end of the lamp subpath: utils::pairs fix-up (skips the last pair, utils.rs:5-13), drop a trailing
// emission vertex, reverse (bidirectional.rs:187-202); then the tail folds (see LightVertex::fold).
//
// `contribute` (renderer/algorithm.rs:14-100) on a tail lamp_path[i..], started with (brightness b, reflectance r):
//     emission vertex k:  b += (c_k p_k) r                      other vertex k:  r *= c_k p_k;  r *= brdf_k
//     and after the FIRST vertex of the tail:  r *= brdf_in  (= brdf_i / brdf_i, the reference's quirk, :365-372)
// With T_i = the fold of lamp_path[i..] without that first-vertex factor (T_n = 0):
//     emission:  T_i = c_i p_i + T_{i+1}                          fold_i = c_i p_i + brdf_in_i T_{i+1}
//     other:     T_i = ((c_i p_i) brdf_i) T_{i+1}                 fold_i = (((c_i p_i) brdf_i) brdf_in_i) T_{i+1}
// (same products as the reference, associated from the tail end instead of from the connection: float rounding differs
// in the last bits only, far inside the film tolerance).  Additional wavelengths are exposed only if no vertex of the
// tail is dispersive (`use_additional`), which the VT_TAIL_DISPERSED bit records.
PYR_HD void finish_lamp_path(const SceneView& sc, PathState& ps, const BidirCtx& cx) {
    LightVertex* lv = cx.lv;
    uint32_t n = ps.bd->n_light;
    if (n >= 2)
        for (uint32_t pos = 0; pos + 2 < n; ++pos) {
            LightVertex& to = lv[pos];
            LightVertex& from = lv[pos + 1];
            to.incident[0] = -from.incident[0]; to.incident[1] = -from.incident[1]; to.incident[2] = -from.incident[2];
            if (from.type == VT_DIFFUSE) { from.out[0] = from.incident[0]; from.out[1] = from.incident[1]; from.out[2] = from.incident[2]; }
        }
    if (n > 1 && lv[n - 1].type == VT_EMISSION) n -= 1;
    for (uint32_t i = 0; i < n / 2; ++i) {  // only the heads move
        const LightVertexHead a = load_head(lv + i), b = load_head(lv + n - 1 - i);
        store_head(lv + i, b);
        store_head(lv + n - 1 - i, a);
    }
    ps.bd->n_light = n;
    PYR_REGFILE(R);
    const uint32_t S = sc.renderer.spectrum_samples;
    const SpecArray T = cx.bright, c = cx.refl;  // the detached per-wavelength scratch arrays (unused until the connect phase)
    for (uint32_t k = 0; k < S; ++k) T[k] = 0.0f;
    bool tail_dispersed = false;
    for (uint32_t i = n; i-- > 0;) {
        const LightVertexHead v = load_head(lv + i);
        VmInputs in;
        in.wavelength = 0.0f; in.incident = ld3(v.incident); in.normal = ld3(v.normal); in.tex[0] = v.tex[0]; in.tex[1] = v.tex[1];
        eval_spectral_each(sc, v.color_program, in, ps.wl, S, R, [&](uint32_t k, float value) { c[k] = value; });
        const float brdf = vertex_brdf(v), brdf_in = brdf / brdf;
        tail_dispersed = tail_dispersed || v.dispersed != 0;
        Vec8 out[2];
#pragma unroll
        for (uint32_t k = 0; k < MAX_SPECTRUM_SAMPLES; ++k) {
            float fold = 0.0f;
            if (k < S) {
                const float cp = c[k] * v.probability, rest = T[k];
                if (v.type == VT_EMISSION) { fold = cp + brdf_in * rest; T[k] = cp + rest; }
                else { fold = ((cp * brdf) * brdf_in) * rest; T[k] = (cp * brdf) * rest; }
            }
            out[k >> 3].v[k & 7] = fold;
        }
        st256(lv[i].fold, out[0]);
        if (S > 8) st256(lv[i].fold + 8, out[1]);
        if (tail_dispersed) lv[i].type = v.type | VT_TAIL_DISPERSED;
    }
}

// The connections of connect_paths (bidirectional.rs:310-398): (camera vertex, lamp vertex) pairs in the reference's order
// (camera vertices outermost), walked from the cursor (cam, light) until BDPT_STAGE of them have been found or the pairs
// are exhausted.  new_cam(index, vertex) is called when the walk enters a camera vertex, f(j, i, geometry of lamp vertex i,
// origin, unit direction, distance, squared distance) for the j-th connection.  Returns their number and leaves the cursor
// behind the last pair looked at.  The same walk counts the visibility rays, writes them, and evaluates them one iteration
// later, so nothing has to be remembered per ray - and one step serves as many camera vertices as fit the batch.
template <class NewCam, class F>
PYR_HD uint32_t for_each_connection(const PathState& ps, const BidirCtx& cx, uint32_t& cam, uint32_t& light, NewCam&& new_cam, F&& f) {
    uint32_t n = 0;
    const uint32_t n_light = ps.bd->n_light, n_cam = ps.bd->n_cam_stored;
    while (cam < n_cam && n < (uint32_t)BDPT_STAGE) {
        const CamVertex& c = cx.cv[cam];
        const Vec8 head = ld256(&c);   // position, brdf, normal, use_additional
        const v3 from = mk3(head.v[0], head.v[1], head.v[2]), cn = mk3(head.v[4], head.v[5], head.v[6]);
        new_cam(cam, c, head);
        uint32_t i = light;
        Vec8 fetched = ld256(cx.lv + (i < n_light ? i : 0u));
        for (; i < n_light && n < (uint32_t)BDPT_STAGE; ++i) {
            const VertexGeometry v = vertex_geometry(fetched);
            if (i + 1 < n_light) fetched = ld256(cx.lv + i + 1);  // the next vertex is on its way while this one is examined
            if (v.type == VT_SPECULAR) continue;
            v3 direction = v.position - from;
            float sq_distance = length2(direction);
            float distance = sqrtf(sq_distance);
            v3 dir = direction / distance;
            if (dot(cn, dir) <= 0.0f) continue;
            if (dot(v.normal, -dir) <= 0.0f) continue;
            f(n, i, v, from, cn, dir, distance, sq_distance);
            ++n;
        }
        if (i >= n_light) { cam += 1; light = 0; }
        else { light = i; break; }  // the batch is full: the next one goes on with this camera vertex
    }
    return n;
}
PYR_HD uint32_t stage_connections(const PathState& ps, const BidirCtx& cx, uint32_t& cam, uint32_t& light, Ray* rays) {
    return for_each_connection(ps, cx, cam, light, [](uint32_t, const CamVertex&, const Vec8&) {},
                               [&](uint32_t j, uint32_t, const VertexGeometry&, v3 from, v3, v3 dir, float distance, float) {
        if (rays) rays[j] = make_ray(from, dir, 2, distance - DIST_EPSILON);  // blocked <=> a hit closer than distance - eps
    });
}

// Stage the next batch of connections from the cursor (conn_cam, conn_light); false when every pair is done.
PYR_HD bool advance_connect(PathState& ps, const BidirCtx& cx, BidirOut& out) {
    uint32_t cam = ps.bd->conn_cam, light = ps.bd->conn_light;
    const uint32_t n = stage_connections(ps, cx, cam, light, PYR_STAGE_RAYS(out));
    if (!n) return false;
    ps.bd->conn_next_cam = cam; ps.bd->conn_next = light;
    ps.n_pending = n;
    out.n_shadow = n; out.shadow_kind = SH_CONNECT; out.alive = 1;
    return true;
}

// Camera::is_visible up to the visibility ray, for the diffuse lamp vertices from `from_light` (cameras.rs:99-142): f(j, i,
// geometry of lamp vertex i, its camera-space position, lens sample, world-space lens point, unit direction, distance).  `rng` advances exactly as the
// reference's does (one lens sample per candidate vertex).
template <class F>
PYR_HD uint32_t for_each_splat(const SceneView& sc, const PathState& ps, const BidirCtx& cx, Rng& rng, uint32_t from_light, uint32_t& next, F&& f) {
    uint32_t n = 0, i = from_light;
    if (!sc.camera.inv_ok) { next = ps.bd->n_light; return 0; }
    const uint32_t n_light = ps.bd->n_light;
    Vec8 fetched = ld256(cx.lv + (i < n_light ? i : 0u));
    for (; i < n_light && n < (uint32_t)BDPT_STAGE; ++i) {
        const VertexGeometry v = vertex_geometry(fetched);
        if (i + 1 < n_light) fetched = ld256(cx.lv + i + 1);
        if (v.type != VT_DIFFUSE) continue;
        const v3 target = v.position;
        v3 local_target = transform_point(sc.camera.inv, target);
        if (local_target.z >= 0.0f) continue;
        const v3 origin = lens_origin(sc.camera, rng);
        const v3 world_origin = transform_point(sc.camera.m, origin);
        const v3 direction = target - world_origin;
        const float distance = length(direction);
        f(n, i, v, local_target, origin, world_origin, direction / distance, distance);
        ++n;
    }
    next = i;
    return n;
}
PYR_HD uint32_t stage_visibility(const SceneView& sc, const PathState& ps, const BidirCtx& cx, Rng& rng, uint32_t from_light, Ray* rays, uint32_t& next) {
    return for_each_splat(sc, ps, cx, rng, from_light, next, [&](uint32_t j, uint32_t, const VertexGeometry&, v3, v3, v3 world_origin, v3 dir, float distance) {
        if (rays) rays[j] = make_ray(world_origin, dir, 2, distance - DIST_EPSILON);
    });
}

PYR_HD bool advance_splat(const SceneView& sc, PathState& ps, const BidirCtx& cx, BidirOut& out) {
    while (ps.bd->conn_light < ps.bd->n_light) {
        ps.bd->rng_saved = ps.rng;
        uint32_t next;
        uint32_t n = stage_visibility(sc, ps, cx, ps.rng, ps.bd->conn_light, PYR_STAGE_RAYS(out), next);
        if (n) { ps.bd->conn_next = next; ps.n_pending = n; out.n_shadow = n; out.shadow_kind = SH_SPLAT; out.alive = 1; return true; }
        ps.bd->conn_light = next;
    }
    return false;
}

// The staged visibility rays of a connect / splat step, written where the kernel reserved room for them (`dst`, in the ray
// queue): the same walk as the counting pass (the lens samples replay from the saved RNG state).
PYR_HD void write_staged(const SceneView& sc, const PathState& ps, const BidirCtx& cx, uint32_t shadow_kind, Ray* dst) {
    uint32_t next;
    if (shadow_kind == SH_CONNECT) { uint32_t cam = ps.bd->conn_cam, light = ps.bd->conn_light; stage_connections(ps, cx, cam, light, dst); }
    else if (shadow_kind == SH_SPLAT) { Rng replay = ps.bd->rng_saved; stage_visibility(sc, ps, cx, replay, ps.bd->conn_light, dst, next); }
}

// After the camera path has ended: expose it; the connect phase starts with the NEXT iteration (no ray now).  The first batch
// of connections is staged by the connect kernel, where every lane walks lamp vertices, not here, where only the few lanes whose
// camera path just ended would (35 % of the camera kernel's stall samples sat in that walk in the steady state).
template <class Add>
PYR_HD void end_camera_path(const SceneView& sc, PathState& ps, const BidirCtx& cx, BidirOut& out, Add& add) {
    (void)cx;
    expose_path(sc, ps, add);  // bidirectional.rs:245-251
    ps.bd->conn_cam = 0; ps.bd->conn_light = 0; ps.bd->conn_next_cam = 0; ps.bd->conn_next = 0;
    ps.n_pending = 0;
    if (ps.bd->n_light > 0) {
        ps.bd->phase = PH_CONNECT;
        out.alive = 1;   // n_shadow = 0: the connect step finds nothing to evaluate and stages the first batch
        return;
    }
    out.alive = 0;
}

// -------------------------------------------------------------------------------- generation
// The start of one `render_tile` iteration (bidirectional.rs:105-176).
PYR_HD void generate_bidirectional(const SceneView& sc, uint64_t seed, uint32_t tile, uint64_t sample, PathState& ps, const BidirCtx& cx, BidirOut& out) {
    PYR_REGFILE(R);
    const TileRec t = sc.tiles[tile];
    Rng rng = keyed_rng(seed, t.index, sample);
    float ox = t.size[0] * rng.gen_f32();
    float oy = t.size[1] * rng.gen_f32();
    ps.pos[0] = t.from[0] + ox;
    ps.pos[1] = t.from[1] + oy;
    const uint32_t S = sc.renderer.spectrum_samples;
    uint32_t pick = sample_wavelengths(sc, rng, ps.wl);
    hero_first(ps.wl, S, pick);
    for (uint32_t k = 0; k < S; ++k) { ps.bright[k] = 0.0f; ps.refl[k] = 1.0f; }
    const float wavelength = ps.wl[0];
    v3 co, cd;
    camera_ray(sc.camera, ps.pos[0], ps.pos[1], rng, co, cd);
    st3(ps.bd->cam_o, co); st3(ps.bd->cam_d, cd);
    ps.tile = tile;
    ps.flags = 0;
    ps.bd->n_light = 0; ps.bd->n_cam = 0; ps.bd->n_cam_stored = 0; ps.bd->lamp_bounces = 0; ps.bd->conn_cam = 0; ps.bd->conn_light = 0; ps.bd->conn_next = 0; ps.bd->conn_next_cam = 0;
    ps.bd->cam_store_pending = 0; ps.light_events = 0; ps.n_pending = 0; ps.bounce = 0;
    out.alive = 0; out.has_main = 0; out.n_shadow = 0; out.shadow_kind = SH_NONE;

    // World::pick_lamp + Lamp::sample_ray
    const uint32_t lamp_index = (uint32_t)rng.gen_range_usize(sc.n_lamps);
    const float lamp_probability = 1.0f / (float)sc.n_lamps;
    const LampRec lamp = sc.lamps[lamp_index];
    bool have = false;
    v3 origin = mk3(0, 0, 0), direction = mk3(0, 0, 0), normal = mk3(0, 0, 0);
    float weight = 0.0f, material_probability = 1.0f, tex[2] = {0.0f, 0.0f};
    int32_t color = -1;
    bool dispersed = false;
    if (lamp.kind == LAMP_POINT) {
        direction = sample_sphere(rng);
        origin = ld3(lamp.v);
        weight = 4.0f * PYR_PI;
        color = lamp.color_program;
        normal = direction;
        have = true;
    } else if (lamp.kind == LAMP_SHAPE) {
        const Prim pr = sc.prims[lamp.rank];
        float u, v;
        prim_sample_point(pr, rng, origin, u, v);
        Surface s;
        prim_point_surface(sc, pr, lamp.rank, origin, u, v, s);
        direction = sample_hemisphere(rng, s.normal);
        weight = prim_surface_area(sc, pr, lamp.rank);
        const MaterialRec m = sc.materials[s.material];
        const ComponentRec comp = sc.components[m.emissive_offset + rng.gen_index_u32(m.n_emissive)];  // choose_emissive
        VmInputs in;
        in.wavelength = wavelength; in.normal = s.normal; in.incident = -direction; in.tex[0] = s.tex[0]; in.tex[1] = s.tex[1];
        material_probability = component_probability(sc, comp, in, R, dispersed);
        color = comp.color_program;
        normal = s.normal;
        tex[0] = s.tex[0]; tex[1] = s.tex[1];
        have = true;
    }
    ps.rng = rng;
    if (!have) {  // a directional lamp has no sample_ray (lamp.rs:86): empty lamp path
        begin_camera(ps, out);
        return;
    }
    origin = origin + normal * DIST_EPSILON;
    LightVertexHead first;
    st3(first.position, origin); first.type = VT_EMISSION;
    st3(first.normal, normal); first.color_program = color;
    first.incident[0] = first.incident[1] = first.incident[2] = 0.0f;
    first.probability = weight / (lamp_probability * material_probability);
    first.out[0] = first.out[1] = first.out[2] = 0.0f;
    first.dispersed = dispersed ? 1u : 0u;
    first.tex[0] = tex[0]; first.tex[1] = tex[1];
    for (uint32_t& w : first.pad) w = 0;
    store_head(cx.lv, first);
    ps.bd->n_light = 1;
    if (sc.renderer.light_bounces == 0) {
        finish_lamp_path(sc, ps, cx);
        begin_camera(ps, out);
        return;
    }
    ps.bd->phase = PH_LAMP;
    out.has_main = 1;
    out.main = make_ray(origin, direction, 0, 0.0f);
    out.alive = 1;
}

// -------------------------------------------------------------------------------- the per-iteration step
// One iteration of `trace(&mut lamp_path, .., light_bounces, 0)` (tracer.rs:221-343 with light_samples = 0:
// sample_light stays true, trace_direct still draws its lamp pick for the first two diffuse events).
PYR_HD bool lamp_step(const SceneView& sc, PathState& ps, const BidirCtx& cx, const Ray& ray, const Hit& h, BidirOut& out, PathCounters& pc) {
    PYR_REGFILE(R);
    const v3 o = ld3(ray.o), d = ld3(ray.d);
    const float wavelength = ps.wl[0];
    LightVertexHead nv;
    for (uint32_t& w : nv.pad) w = 0;
    st3(nv.incident, d);
    nv.out[0] = nv.out[1] = nv.out[2] = 0.0f;
    if (h.kind == KIND_MISS) {
        nv.type = VT_EMISSION; nv.dispersed = 0;
        nv.color_program = directional_color(sc, d, sc.sky_program);
        st3(nv.position, d * PYR_INF); st3(nv.normal, -d);
        nv.tex[0] = nv.tex[1] = 0.0f; nv.probability = 1.0f;
        store_head(cx.lv + ps.bd->n_light++, nv);
        return false;
    }
    Surface s;
    hit_surface(sc, o, d, h, s, pc.de_evals, pc.de_iters);
    const v3 normal = apply_normal_map(sc, s, d, R);
    const MaterialRec mat = sc.materials[s.material];
    const ComponentRec comp = sc.components[mat.comp_offset + ps.rng.gen_index_u32(mat.n_components)];
    VmInputs pin;
    pin.wavelength = wavelength; pin.normal = normal; pin.incident = d; pin.tex[0] = s.tex[0]; pin.tex[1] = s.tex[1];
    bool normal_dispersed;
    const float component_prob = component_probability(sc, comp, pin, R, normal_dispersed);
    const Scatter sct = scatter(comp, d, normal, wavelength, ps.rng);
    st3(nv.position, s.position); st3(nv.normal, normal);
    nv.tex[0] = s.tex[0]; nv.tex[1] = s.tex[1];
    nv.color_program = comp.color_program;
    if (sct.emitted) {
        nv.type = VT_EMISSION; nv.dispersed = normal_dispersed ? 1u : 0u; nv.probability = component_prob;
        store_head(cx.lv + ps.bd->n_light++, nv);
        return false;
    }
    if (ps.light_events < 2 && sct.has_brdf) {
        ps.light_events += 1;
        (void)ps.rng.gen_range_usize(sc.n_lamps);  // trace_direct's World::pick_lamp, with zero samples
    }
    nv.type = sct.has_brdf ? VT_DIFFUSE : VT_SPECULAR;
    st3(nv.out, sct.out);
    nv.dispersed = (sct.dispersed || normal_dispersed) ? 1u : 0u;
    nv.probability = sct.probability * component_prob;
    store_head(cx.lv + ps.bd->n_light++, nv);
    ps.bd->lamp_bounces += 1;
    if (ps.bd->lamp_bounces < sc.renderer.light_bounces) {
        out.has_main = 1;
        out.main = make_ray(s.position, sct.out, 0, 0.0f);
        out.alive = 1;
        return true;
    }
    return false;
}

// ---- one wavefront step per phase.  The kernels run one specialised kernel per phase over the slots sorted by phase (each with
// the registers ITS phase needs); shade_bidirectional below dispatches for the CPU emulation.
// PH_LAMP: one iteration of the lamp subpath; when it ends the sample goes through PH_FINISH (no ray this iteration).
PYR_HD void shade_bd_lamp(const SceneView& sc, PathState& ps, const BidirCtx& cx, const Ray* main_ray, const Hit* main_hit, BidirOut& out, PathCounters& pc) {
    out.alive = 0; out.has_main = 0; out.n_shadow = 0; out.shadow_kind = SH_NONE;
    if (lamp_step(sc, ps, cx, load_record_stream(main_ray), load_record(main_hit), out, pc)) return;
    ps.bd->phase = PH_FINISH;
    out.alive = 1;
}
// PH_FINISH: fix-up, reversal, colours and tail folds of the finished lamp subpath; then the camera subpath starts.
PYR_HD void shade_bd_finish(const SceneView& sc, PathState& ps, const BidirCtx& cx, BidirOut& out) {
    out.alive = 0; out.has_main = 0; out.n_shadow = 0; out.shadow_kind = SH_NONE;
    finish_lamp_path(sc, ps, cx);
    ps.light_events = 0;
    begin_camera(ps, out);
}
// PH_CAMERA: one iteration of the camera subpath (the camera-to-light integrator's step with the vertex hooks); `so` carries its
// next-event visibility rays (SH_NEE).  When the path ends: expose, then connections, then light tracing.
template <class Add>
PYR_HD void shade_bd_camera(const SceneView& sc, PathState& ps, const BidirCtx& cx, const Ray* main_ray, const Hit* main_hit, const Ray* shadow_rays,
                            const uint32_t* shadow_kinds, ShadeOut& so, BidirOut& out, Add& add, PathCounters& pc) {
    out.alive = 0; out.has_main = 0; out.n_shadow = 0; out.shadow_kind = SH_NONE;
    CameraHooks hooks{cx, sc.renderer.spectrum_samples};
    const bool more = camera_step(sc, ps, main_ray, main_hit, shadow_rays, shadow_kinds, so, pc, hooks);
    if (more) {
        out.alive = 1; out.has_main = so.has_main; out.main = so.main; out.n_shadow = so.n_shadow; out.shadow_kind = so.n_shadow ? SH_NEE : SH_NONE;
        return;
    }
    end_camera_path(sc, ps, cx, out, add);
}
// PH_CONNECT: evaluate the connections whose visibility rays were just traced (bidirectional.rs:310-398), stage the next ones.
// Every connection of a path sample is exposed at the sample's film position with the sample's wavelengths and the same
// weight 1 / (len(camera_path) * len(lamp_path)) (bidirectional.rs:217-218), so the step sums them per wavelength and makes
// ONE film update per wavelength instead of one per connection (the film is additive: (sum of value * weight, sum of weight)).
template <class Add>
PYR_HD void shade_bd_connect(const SceneView& sc, PathState& ps, const BidirCtx& cx, const uint32_t* shadow_kinds, BidirOut& out, Add& add) {
    out.alive = 0; out.has_main = 0; out.n_shadow = 0; out.shadow_kind = SH_NONE;
    const uint32_t S = sc.renderer.spectrum_samples;
    const float weight = 1.0f / (float)(ps.bd->n_cam * ps.bd->n_light);
    // per-wavelength scratch: the camera vertex' sample state (detached arrays) and the step's sums (the sample's own brightness /
    // reflectance arrays are dead once the camera path has been exposed)
    const SpecArray bright = cx.bright, refl = cx.refl, sum = ps.bright, count = ps.refl;
    for (uint32_t k = 0; k < S; ++k) { sum[k] = 0.0f; count[k] = 0.0f; }
    float brdf = 1.0f;
    bool cam_additional = false;
    uint32_t cam = ps.bd->conn_cam, light = ps.bd->conn_light;
    if (ps.n_pending)   // (a sample that has just entered the phase has nothing staged yet)
    for_each_connection(ps, cx, cam, light,
        [&](uint32_t, const CamVertex& stored, const Vec8& head) {
            // the sample state right after this camera vertex' `contribute`: fetched once per vertex, not once per connection
            brdf = head.v[3];
            cam_additional = f_bits(head.v[7]) != 0;
            static_assert(MAX_SPECTRUM_SAMPLES == 16, "two chunks per array");
            Vec8 b[2], r[2];
            b[0] = ld256(stored.bright); r[0] = ld256(stored.refl);
            if (S > 8) { b[1] = ld256(stored.bright + 8); r[1] = ld256(stored.refl + 8); }
#pragma unroll
            for (uint32_t k = 0; k < MAX_SPECTRUM_SAMPLES; ++k)
                if (k < S) { bright[k] = b[k >> 3].v[k & 7]; refl[k] = r[k >> 3].v[k & 7]; }
        },
        [&](uint32_t j, uint32_t i, const VertexGeometry& v, v3, v3 cn, v3 dir, float, float sq_distance) {
            if (shadow_kinds[j] != KIND_MISS) return;
            float cos_out = fabsf(dot(cn, dir));
            float cos_in = fabsf(dot(v.normal, -dir));
            float brdf_out = (2.0f * fabsf(dot(dir, cn))) / brdf;
            float scale = cos_in * cos_out * brdf_out / (2.0f * PYR_PI * sq_distance);
            Vec8 f[2];
            f[0] = ld256(cx.lv[i].fold);
            if (S > 8) f[1] = ld256(cx.lv[i].fold + 8);
            const uint32_t n = (cam_additional && !v.tail_dispersed) ? S : 1u;
            // brightness + (reflectance * scale) * fold(lamp_path[i..])
#pragma unroll
            for (uint32_t k = 0; k < MAX_SPECTRUM_SAMPLES; ++k)
                if (k < n) { sum[k] += bright[k] + (refl[k] * scale) * f[k >> 3].v[k & 7]; count[k] += 1.0f; }
        });
    for (uint32_t k = 0; k < S; ++k)
        if (count[k] > 0.0f) film_expose_sum(sc.film, ps.pos[0], ps.pos[1], sum[k] * weight, ps.wl[k], count[k] * weight, add);
    ps.bd->conn_cam = ps.bd->conn_next_cam;
    ps.bd->conn_light = ps.bd->conn_next;
    if (advance_connect(ps, cx, out)) return;
    ps.bd->phase = PH_SPLAT;
    ps.bd->conn_light = 0;
    if (advance_splat(sc, ps, cx, out)) return;
    out.alive = 0;
}
// PH_SPLAT: evaluate the light-traced samples whose visibility rays were just traced (bidirectional.rs:253-306)
template <class Add>
PYR_HD void shade_bd_splat(const SceneView& sc, PathState& ps, const BidirCtx& cx, const uint32_t* shadow_kinds, BidirOut& out, Add& add) {
    out.alive = 0; out.has_main = 0; out.n_shadow = 0; out.shadow_kind = SH_NONE;
    const uint32_t S = sc.renderer.spectrum_samples;
    Rng replay = ps.bd->rng_saved;
    const float weight = 1.0f / (float)ps.bd->n_light;
    uint32_t next;
    for_each_splat(sc, ps, cx, replay, ps.bd->conn_light, next, [&](uint32_t j, uint32_t i, const VertexGeometry& v, v3 local_target, v3 origin, v3 world_origin, v3, float) {
        if (shadow_kinds[j] != KIND_MISS) return;
        const v3 target = v.position;
        local_target.z += sc.camera.focus_distance;
        const float dist = local_target.z;
        local_target = local_target - (origin * dist) / sc.camera.focus_distance;
        local_target.z -= sc.camera.focus_distance;
        const v3 view_plane_target = (-local_target) / local_target.z;
        const float px = view_plane_target.x * sc.camera.view_plane, py = (-view_plane_target.y) * sc.camera.view_plane;
        if (!(px > -1.0f && px < 1.0f && py > -1.0f && py < 1.0f)) return;
        const float sq_distance = length2(world_origin - target);
        const float scale = 1.0f / sq_distance;
        Vec8 f[2];
        f[0] = ld256(cx.lv[i].fold);
        if (S > 8) f[1] = ld256(cx.lv[i].fold + 8);
        // brightness 0, reflectance `scale`: the sample is scale * fold(lamp_path[i..])
        film_expose(sc.film, px, py, scale * f[0].v[0], ps.wl[0], weight, add);
        if (!v.tail_dispersed) {
#pragma unroll
            for (uint32_t k = 1; k < MAX_SPECTRUM_SAMPLES; ++k)
                if (k < S) film_expose(sc.film, px, py, scale * f[k >> 3].v[k & 7], ps.wl[k], weight, add);
        }
    });
    ps.bd->conn_light = ps.bd->conn_next;
    if (advance_splat(sc, ps, cx, out)) return;
    out.alive = 0;
}

#if !defined(__CUDA_ARCH__)
// The CPU emulation's dispatcher (tests/host_emu.cpp): one step of whatever phase the sample is in.
template <class Add>
inline void shade_bidirectional(const SceneView& sc, PathState& ps, const BidirCtx& cx, const Ray* main_ray, const Hit* main_hit,
                                const Ray* shadow_rays, const uint32_t* shadow_kinds, BidirOut& out, Add& add, PathCounters& pc) {
    switch (ps.bd->phase) {
        case PH_LAMP: shade_bd_lamp(sc, ps, cx, main_ray, main_hit, out, pc); return;
        case PH_CAMERA: {
            ShadeOut so;
            so.stage_base = 0;
            shade_bd_camera(sc, ps, cx, main_ray, main_hit, shadow_rays, shadow_kinds, so, out, add, pc);
            if (out.shadow_kind == SH_NEE) for (uint32_t j = 0; j < out.n_shadow; ++j) out.shadow[j] = so.get_shadow(j);
            return;
        }
        case PH_CONNECT: shade_bd_connect(sc, ps, cx, shadow_kinds, out, add); return;
        case PH_FINISH: shade_bd_finish(sc, ps, cx, out); return;
        default: shade_bd_splat(sc, ps, cx, shadow_kinds, out, add); return;
    }
}
#endif

}  // namespace pyr
