/* A plain C host of include/pyrite_b200.h - no Python, no ctypes: what a Rust `-sys` crate's generated bindings call
 * (INTEGRATION.md).  It loads a project IR blob from a file, runs the render path through the C ABI
 *     pyr_init -> pyr_project_load -> pyr_trace -> pyr_render -> pyr_film_develop
 * and checks the results against the oracle (TEST INFRASTRUCTURE: liboracle.so, loaded with dlopen, given the same blob
 * and the same seeds): hit ids bit-exact, distances within 1e-5 relative, film mean luminance within 1e-3 and per-pixel
 * RMSE/mean within 1e-2.  Exit code 0 = parity, 1 = mismatch, 2 = could not run (no GPU, missing file).
 *
 *     gcc -std=c11 -O1 -I include -o tests/_build/abi_smoke tests/abi_smoke.c -ldl -lm
 *     tests/_build/abi_smoke pyrite_b200/libpyrite_b200.so oracle/liboracle.so project.ir
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pyrite_b200.h"

/* the oracle's C entry points (oracle/pyro_api.cpp) */
typedef struct { uint32_t width, height, bins, algorithm, pixel_samples, bounces, light_samples, spectrum_samples, light_bounces, tile_size,
                 n_objects, n_planes, n_lights, n_bvh_nodes, n_materials, threads; } oracle_info;
typedef struct { uint64_t seed; int32_t rng_mode, eager_emissive_draw; uint32_t spp_override, sample_offset, sample_stride; int32_t threads, cas_attempts, reset_film; } oracle_render_opts;

#define LOAD(lib, type, name)                                                      \
    type name = (type)dlsym(lib, #name);                                           \
    if (!name) { fprintf(stderr, "missing symbol %s: %s\n", #name, dlerror()); return 2; }

typedef pyr_status (*fn_init)(int32_t, pyr_ctx**);
typedef void (*fn_shutdown)(pyr_ctx*);
typedef const char* (*fn_last_error)(const pyr_ctx*);
typedef pyr_status (*fn_project_load)(pyr_ctx*, const void*, size_t);
typedef pyr_status (*fn_project_info)(const pyr_ctx*, pyr_project_info*);
typedef pyr_status (*fn_trace)(pyr_ctx*, const pyr_ray*, size_t, pyr_hit*);
typedef pyr_status (*fn_render)(pyr_ctx*, const pyr_render_params*, pyr_progress_cb, void*);
typedef pyr_status (*fn_develop)(pyr_ctx*, float, float*, uint8_t*);
typedef pyr_status (*fn_counters)(pyr_ctx*, pyr_counters*, int32_t);
typedef const char* (*fn_version)(void);

static int progress_calls = 0;
static int on_progress(uint8_t progress, const char* message, void* user) {
    (void)progress; (void)message; (void)user;
    ++progress_calls;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s libpyrite_b200.so liboracle.so project.ir\n", argv[0]); return 2; }
    void* product = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!product) { fprintf(stderr, "cannot load %s: %s\n", argv[1], dlerror()); return 2; }
    void* oracle = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    if (!oracle) { fprintf(stderr, "cannot load %s: %s\n", argv[2], dlerror()); return 2; }
    FILE* f = fopen(argv[3], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[3]); return 2; }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* ir = malloc((size_t)bytes);
    if (fread(ir, 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);

    LOAD(product, fn_init, pyr_init)
    LOAD(product, fn_shutdown, pyr_shutdown)
    LOAD(product, fn_last_error, pyr_last_error)
    LOAD(product, fn_project_load, pyr_project_load)
    LOAD(product, fn_project_info, pyr_project_info_get)
    LOAD(product, fn_trace, pyr_trace)
    LOAD(product, fn_render, pyr_render)
    LOAD(product, fn_develop, pyr_film_develop)
    LOAD(product, fn_counters, pyr_counters_get)
    LOAD(product, fn_version, pyr_version)
    typedef int (*o_load)(const void*, size_t, void**);
    typedef int (*o_info)(void*, oracle_info*);
    typedef int (*o_gen)(void*, int, size_t, uint64_t, pyr_ray*);
    typedef int (*o_trace)(void*, const pyr_ray*, size_t, pyr_hit*, int, void*);
    typedef int (*o_render)(void*, const oracle_render_opts*);
    typedef int (*o_develop)(void*, float, float*, uint8_t*, int);
    LOAD(oracle, o_load, pyro_load)
    LOAD(oracle, o_info, pyro_info)
    LOAD(oracle, o_gen, pyro_gen_rays)
    LOAD(oracle, o_trace, pyro_trace)
    LOAD(oracle, o_render, pyro_render)
    LOAD(oracle, o_develop, pyro_film_develop)

    pyr_ctx* ctx = NULL;
    pyr_status st = pyr_init(0, &ctx);
    if (st != PYR_OK) {   /* no CPU path exists: without a usable device the library says so and the host gives up */
        fprintf(stderr, "pyr_init failed with status %d: %s\n", st, pyr_last_error(NULL));
        return 2;
    }
    printf("%s\n", pyr_version());
    if (pyr_project_load(ctx, ir, (size_t)bytes) != PYR_OK) { fprintf(stderr, "pyr_project_load: %s\n", pyr_last_error(ctx)); return 1; }
    pyr_project_info info;
    if (pyr_project_info_get(ctx, &info) != PYR_OK) return 1;
    void* o = NULL;
    if (pyro_load(ir, (size_t)bytes, &o) != 0) { fprintf(stderr, "the oracle rejected the project\n"); return 2; }
    oracle_info oi;
    pyro_info(o, &oi);
    if (oi.width != info.width || oi.height != info.height || oi.n_objects != info.n_objects || oi.n_bvh_nodes != info.n_bvh_nodes) {
        fprintf(stderr, "project info differs from the oracle's\n");
        return 1;
    }

    /* World::intersect on three oracle-generated ray batches (camera, bounce and shadow rays) */
    const size_t n = 50000;
    pyr_ray* rays = (pyr_ray*)malloc(n * sizeof(pyr_ray));
    pyr_hit* got = (pyr_hit*)malloc(n * sizeof(pyr_hit));
    pyr_hit* want = (pyr_hit*)malloc(n * sizeof(pyr_hit));
    size_t id_mismatches = 0;
    double worst_t = 0.0;
    for (int kind = 0; kind < 3; ++kind) {
        if (pyro_gen_rays(o, kind, n, 900 + (uint64_t)kind, rays) != 0) { fprintf(stderr, "the oracle could not generate rays\n"); return 2; }
        pyro_trace(o, rays, n, want, 4, NULL);
        if (pyr_trace(ctx, rays, n, got) != PYR_OK) { fprintf(stderr, "pyr_trace: %s\n", pyr_last_error(ctx)); return 1; }
        for (size_t i = 0; i < n; ++i) {
            if (got[i].kind != want[i].kind || got[i].prim_id != want[i].prim_id) { ++id_mismatches; continue; }
            if (want[i].kind != PYR_KIND_MISS) {
                double rel = fabs((double)got[i].t - (double)want[i].t) / fabs((double)want[i].t);
                if (rel > worst_t) worst_t = rel;
            }
        }
    }
    printf("pyr_trace: %zu rays, %zu id mismatches, worst relative distance error %.1e\n", 3 * n, id_mismatches, worst_t);

    /* Renderer::render + the develop loop on identical per-path streams */
    pyr_render_params params;
    memset(&params, 0, sizeof(params));
    params.seed = 11;
    params.reset_film = 1;
    if (pyr_render(ctx, &params, on_progress, NULL) != PYR_OK) { fprintf(stderr, "pyr_render: %s\n", pyr_last_error(ctx)); return 1; }
    oracle_render_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.seed = 11; opts.rng_mode = 1; opts.eager_emissive_draw = 1; opts.sample_stride = 1; opts.threads = 4; opts.reset_film = 1;
    pyro_render(o, &opts);
    const size_t pixels = (size_t)info.width * info.height;
    float* xyz_g = (float*)malloc(pixels * 3 * sizeof(float));
    float* xyz_o = (float*)malloc(pixels * 3 * sizeof(float));
    uint8_t* srgb_g = (uint8_t*)malloc(pixels * 3);
    uint8_t* srgb_o = (uint8_t*)malloc(pixels * 3);
    if (pyr_film_develop(ctx, 2.0f, xyz_g, srgb_g) != PYR_OK) { fprintf(stderr, "pyr_film_develop: %s\n", pyr_last_error(ctx)); return 1; }
    pyro_film_develop(o, 2.0f, xyz_o, srgb_o, 4);
    double mean_g = 0, mean_o = 0, sq = 0;
    size_t srgb_off = 0;
    for (size_t p = 0; p < pixels; ++p) {
        const double a = xyz_g[3 * p + 1], b = xyz_o[3 * p + 1];
        mean_g += a; mean_o += b; sq += (a - b) * (a - b);
        for (int c = 0; c < 3; ++c) if (abs((int)srgb_g[3 * p + c] - (int)srgb_o[3 * p + c]) > 1) ++srgb_off;
    }
    mean_g /= (double)pixels; mean_o /= (double)pixels;
    const double mean_rel = fabs(mean_g - mean_o) / mean_o, rmse_rel = sqrt(sq / (double)pixels) / mean_o;
    pyr_counters counters;
    pyr_counters_get(ctx, &counters, 0);
    printf("pyr_render: %llu path samples, %llu rays, %llu kernel launches, %d progress callbacks; mean luminance %.6f (oracle %.6f, rel %.1e), "
           "RMSE/mean %.1e, sRGB bytes off by more than 1: %zu\n",
           (unsigned long long)counters.path_samples, (unsigned long long)counters.rays, (unsigned long long)counters.kernel_launches, progress_calls,
           mean_g, mean_o, mean_rel, rmse_rel, srgb_off);
    pyr_shutdown(ctx);
    const int ok = id_mismatches == 0 && worst_t <= 1e-5 && mean_rel <= 1e-3 && rmse_rel <= 1e-2 && counters.kernel_launches > 0 && progress_calls > 0;
    printf(ok ? "ABI SMOKE OK\n" : "ABI SMOKE FAILED\n");
    return ok ? 0 : 1;
}
