/* pyrite_b200 - C ABI of the B200-native render path.
 *
 * The reference (Ogeon/pyrite, a Rust binary crate) has no FFI or plugin interface: the render
 * path is the internal call
 *     Renderer::render(&self, film, task_runner, on_status, camera, world, resources)
 *                                                   (pyrite/src/renderer/mod.rs:77-111)
 * made from one place, pyrite/src/main.rs:245-305, after `parse_project` (main.rs:111-134)
 * has built Camera / Renderer / World from the decoded project, and followed by the film
 * development loop (main.rs:313-327).  This header is the seam a Rust `-sys` crate would bind
 * to replace exactly that path (INTEGRATION.md shows the binding).  Every entry point names
 * the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer; the library
 * owns device memory and streams.  Every call returns pyr_status (0 = ok) and never throws or
 * aborts across the ABI; the message is available from pyr_last_error().  A context is bound
 * to ONE CUDA device and is driven from one host thread at a time (the reference drives
 * Renderer::render from a single scoped thread, main.rs:243-245).  One process per GPU.
 */
#ifndef PYRITE_B200_H
#define PYRITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t pyr_status;
#define PYR_OK 0
#define PYR_ERR_INVALID 1  /* bad argument / malformed project IR / scene build error */
#define PYR_ERR_CUDA 2     /* CUDA runtime failure (no device, out of memory, launch error) */
#define PYR_ERR_STATE 3    /* call made in the wrong state (e.g. render before project load) */

typedef struct pyr_ctx pyr_ctx;

/* collision::Ray3<f32> {origin, direction} (the argument of World::intersect, world.rs:273). 32 B. */
typedef struct pyr_ray {
    float o[3];
    float pad0;
    float d[3];
    float pad1;
} pyr_ray;

/* What World::intersect returns, flattened: Option<Intersection{distance, SurfacePoint}>
 * (world.rs:273-299, shapes/mod.rs:472-524).  20 B.
 *   kind     0 miss, 1 plane, 2 triangle, 3 sphere, 4 ray-marched shape
 *   prim_id  planes: index in World::planes; others: insertion index in World::from_project's
 *            `objects` list (world.rs:75,182,229); 0xFFFFFFFF on a miss
 *   t        Intersection::distance (+inf on a miss);  u, v  triangle barycentrics, else 0 */
typedef struct pyr_hit {
    uint32_t prim_id;
    uint32_t kind;
    float t;
    float u;
    float v;
} pyr_hit;

#define PYR_KIND_MISS 0u
#define PYR_KIND_PLANE 1u
#define PYR_KIND_TRIANGLE 2u
#define PYR_KIND_SPHERE 3u
#define PYR_KIND_RAY_MARCHED 4u

/* The resolved `Renderer` (renderer/mod.rs:18-28, defaults :63-75) plus image size. */
typedef struct pyr_project_info {
    uint32_t width, height, bins, algorithm; /* algorithm: 0 simple, 1 bidirectional */
    uint32_t pixel_samples, bounces, light_samples, spectrum_samples, light_bounces, tile_size;
    uint32_t n_objects, n_planes, n_lights, n_bvh_nodes, n_materials, n_ray_marched;
} pyr_project_info;

/* Render controls that the reference takes from the OS / thread pool and that a host must now
 * choose.  All-zero is valid and means: seed 0, the project's pixel_samples, all samples. */
typedef struct pyr_render_params {
    uint64_t seed;          /* per-path-sample Xorshift128 streams are keyed by (seed, tile, sample) */
    uint32_t spp_override;  /* 0 = the project's renderer.pixel_samples */
    uint32_t sample_offset; /* this context renders samples offset, offset+stride, ... of every tile: */
    uint32_t sample_stride; /*   sample-pass sharding over GPUs (0 is read as 1) */
    uint32_t reset_film;    /* non-zero: clear the film first; zero: keep accumulating (film is additive) */
    uint32_t pool_paths;    /* paths in flight; 0 = library default */
    uint32_t flags;         /* PYR_RENDER_* */
    uint32_t tile_filter;   /* 0 = every tile; t + 1 = only tile t (diagnostics: replaying the samples of one tile) */
    uint32_t reserved;      /* must be 0 */
} pyr_render_params;
#define PYR_RENDER_STATS 1u  /* also count visited BVH nodes / tested leaves (slower) */
#define PYR_RENDER_TIMING 2u /* bracket every trace / shade launch with CUDA events (pyr_counters.*_seconds) */

/* Work done since the last reset.  rays = World::intersect calls (path segments + shadow rays +
 * BDPT connection / visibility rays); path_samples = render_tile loop iterations. */
typedef struct pyr_counters {
    uint64_t rays, path_samples, nodes_visited, leaves_tested, de_evals, de_iterations;
    uint64_t wavefront_iterations, kernel_launches;
    double render_seconds; /* device time of the last pyr_render / pyr_trace_device (CUDA events) */
    /* PYR_RENDER_TIMING: summed device time and launch count of the traversal kernel and of the
     * shade/regenerate kernel since the last reset */
    double trace_seconds, shade_seconds;
    uint64_t trace_launches, shade_launches;
    uint64_t node_fetches; /* 128-byte BVH nodes fetched (stats mode); nodes_visited counts the boxes tested */
    uint64_t path_rays;    /* the part of `rays` that are path segments (closest-hit); the rest are visibility rays */
    uint64_t march_iterations, julia_iterations; /* stats mode: estimator iterations run by the sphere-tracing kernel (de_iterations also counts
                                                   the normal estimation of the shade stage), and the quaternion-Julia part of them */
} pyr_counters;

/* renderer::Progress{progress: u8, message} (renderer/mod.rs:229-232); invoked on the calling
 * thread between wavefront iterations.  Return non-zero to cancel the render. */
typedef int (*pyr_progress_cb)(uint8_t progress, const char* message, void* user);

/* Create a context on CUDA device `device` (replaces ExecutionContext::new / Film::new set-up,
 * main.rs:190-204).  Fails with PYR_ERR_CUDA when no usable device exists: there is no CPU path. */
pyr_status pyr_init(int32_t device, pyr_ctx** out);
void pyr_shutdown(pyr_ctx* ctx);
/* Launch all further work of `ctx` on the caller's CUDA stream (a cudaStream_t; NULL = back to the
 * context's own stream) so that the caller's events and collectives order with it. */
pyr_status pyr_stream_set(pyr_ctx* ctx, void* cuda_stream);

/* Message for the last failing call on `ctx` (ctx == NULL: last pyr_init failure).  Replaces the
 * reference's error prints (main.rs:68-71,104). */
const char* pyr_last_error(const pyr_ctx* ctx);

/* parse_project (main.rs:111-134): Camera::from_project (cameras.rs:30-55),
 * Renderer::from_project (renderer/mod.rs:31-75), World::from_project incl. Bvh::new
 * (world.rs:39-271, spatial/bvh.rs:13-155), ProgramCompiler::compile (program/compiler.rs:48-586)
 * - then uploads the baked scene to the device.  `ir` is the project IR written by
 * pyrite_b200.project.serialize_project (= ProjectData after load_project, project/mod.rs:29-93). */
pyr_status pyr_project_load(pyr_ctx* ctx, const void* ir, size_t bytes);
pyr_status pyr_project_info_get(const pyr_ctx* ctx, pyr_project_info* out);

/* World::intersect (world.rs:273-299) over a batch.  Host buffers; copies are part of the call. */
pyr_status pyr_trace(pyr_ctx* ctx, const pyr_ray* rays, size_t n, pyr_hit* hits_out);
/* Same, on buffers already resident in device memory (`repeat` >= 1 back-to-back launches); the
 * device time of the launches is reported in pyr_counters.render_seconds.  `d_rays` must be
 * 32-byte aligned (rays are fetched with 256-bit loads); cudaMalloc / torch allocations are. */
pyr_status pyr_trace_device(pyr_ctx* ctx, const void* d_rays, size_t n, void* d_hits, uint32_t repeat);
/* pyr_trace that also counts the BVH boxes tested and leaves tested into pyr_counters (slower). */
pyr_status pyr_trace_stats(pyr_ctx* ctx, const pyr_ray* rays, size_t n, pyr_hit* hits_out);
/* The flattened BVH's leaf pre-order (spatial/bvh.rs:250-275): object id of every leaf rank,
 * n_objects entries.  World::intersect's tie rule (earlier leaf wins, world.rs:288-296) is defined on it. */
pyr_status pyr_bvh_leaf_order(pyr_ctx* ctx, uint32_t* object_ids_out);
/* Test hook: out4[0] = 1 when the BVH of the loaded project was built on the GPU (bvh_build.cu; meshes of 16384 triangles or more,
 * PYR_BVH_BUILD=host|gpu overrides), [1] a digest of the 4-wide nodes, [2] a digest of the leaf pre-order, [3] microseconds the
 * last pyr_project_load took.  The GPU build and the host build of Bvh::new (spatial/bvh.rs:13-155) give the same digests. */
pyr_status pyr_bvh_digest(pyr_ctx* ctx, uint64_t* out4);

/* Renderer::render (renderer/mod.rs:77-111 -> simple.rs:17-141 / bidirectional.rs:31-308):
 * runs the wavefront pipeline into the context's film. */
pyr_status pyr_render(pyr_ctx* ctx, const pyr_render_params* params, pyr_progress_cb cb, void* user);

/* Film::expose (film.rs:89-95) over a batch: positions[n][2] view coordinates,
 * samples[n][3] = (brightness, wavelength, weight). */
pyr_status pyr_film_expose(pyr_ctx* ctx, const float* positions, const float* samples, size_t n);
pyr_status pyr_film_clear(pyr_ctx* ctx);
/* The film as W*H*bins (accumulator, weight) f32 pairs (film.rs:165-185), row-major, bin-minor. */
pyr_status pyr_film_download(pyr_ctx* ctx, float* acc_weight_out);
pyr_status pyr_film_upload(pyr_ctx* ctx, const float* acc_weight_in);
/* Device address and byte size of the film, for the one NCCL reduction of the multi-GPU path. */
pyr_status pyr_film_device_ptr(pyr_ctx* ctx, void** d_ptr, size_t* bytes);

/* Multi-GPU (no counterpart in the reference, which renders on the host's threads, renderer/mod.rs:77-111): one process
 * and one context per GPU, rank r of n renders path samples r, r+n, ... of every tile (pyr_render_params.sample_offset /
 * sample_stride) and the films are summed once, over NCCL, before developing on the root - the film is additive, a bin's
 * value is the ratio of its two sums (film.rs:132-185).  The library owns the communicator; the host only has to carry the
 * 128-byte id from rank 0 to the other ranks (a file, a socket, MPI, torch.distributed ...).
 *   pyr_comm_unique_id  rank 0: a fresh id (ncclGetUniqueId)
 *   pyr_comm_init       every rank, collectively (ncclCommInitRank on the context's device)
 *   pyr_comm_init_async the same, but returns at once: the communicator is set up on a thread and a stream of the library's
 *                       own while the host goes on to pyr_render (on an 8-GPU box the set-up takes ~3 s, a sixth of the
 *                       1024-spp target job); pyr_film_reduce / pyr_comm_destroy wait for it and report its failure.  The
 *                       set-up and a render's start-up allocations block each other in the driver, so the thread begins its
 *                       NCCL calls when the next pyr_render has allocated its buffers and queued its first iterations (or when
 *                       something needs the communicator, or after 5 s; PYR_COMM_GATE=0: at once)
 *   pyr_film_reduce     every rank, collectively: film := sum over ranks, on `root` (root < 0: on every rank)
 * NCCL (libnccl.so.2) is bound when the first of these is called; without it they fail with PYR_ERR_STATE. */
#define PYR_COMM_ID_BYTES 128
pyr_status pyr_comm_unique_id(uint8_t* id_out /* PYR_COMM_ID_BYTES */);
pyr_status pyr_comm_init(pyr_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id /* PYR_COMM_ID_BYTES */);
pyr_status pyr_comm_init_async(pyr_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id /* PYR_COMM_ID_BYTES */);
pyr_status pyr_film_reduce(pyr_ctx* ctx, int32_t root);
pyr_status pyr_comm_destroy(pyr_ctx* ctx);

/* The develop loop (main.rs:313-327: developed_pixels -> spectrum_to_xyz(step) -> x3.444 ->
 * LinSrgb::from_color -> into_encoding).  xyz_out[W*H*3] f32 and/or srgb_out[W*H*3] u8, either
 * may be NULL.  step_size 2.0 is the final image, 30.0 the preview (main.rs:274-279). */
pyr_status pyr_film_develop(pyr_ctx* ctx, float step_size, float* xyz_out, uint8_t* srgb_out);

/* Camera seam: Tile::sample_point + Camera::ray_towards + Film::sample_many_wavelengths + hero pick
 * (simple.rs:87-107 / bidirectional.rs:105-124) for path sample (tile, sample): position_out[2],
 * one ray, wavelengths_out[spectrum_samples] in stratum order, index of the hero wavelength. */
pyr_status pyr_camera_sample(pyr_ctx* ctx, uint64_t seed, uint32_t tile, uint64_t sample, float* position_out, pyr_ray* ray_out,
                             float* wavelengths_out, uint32_t* hero_out);

/* Diagnostic seam (tools/first_divergence.py): ONE `render_tile` iteration of the camera-to-light integrator
 * (simple.rs:87-139) for path sample (tile, sample), run depth-first by one GPU thread with the same stage functions the
 * wavefront kernels use, with per-bounce records of `trace` (tracer.rs:221-343).  records_out: max_bounces x 20 words
 * {kind, prim_id, t, u, v, incident[3], position[3], normal[3], out[3], visibility rays cast, Xorshift `w` after the
 * bounce, 1 if a surface bounce was pushed}; exposed_out: 16 x (brightness, wavelength) of which n_exposed are valid. */
pyr_status pyr_debug_path(pyr_ctx* ctx, uint64_t seed, uint32_t tile, uint64_t sample, uint32_t max_bounces, uint32_t* records_out,
                          uint32_t* n_bounces_out, float* exposed_out, uint32_t* n_exposed_out, float* position_out);

pyr_status pyr_counters_get(pyr_ctx* ctx, pyr_counters* out, int32_t reset);

/* Library / device identification for reports: "pyrite_b200 <version>; sm_100a; <device name>; <SMs> SMs". */
const char* pyr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PYRITE_B200_H */
