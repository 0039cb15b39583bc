// Launchers of the sm_100a kernels (kernels.cu), called by the C ABI (abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bvh_build.hpp"
#include "device_types.h"

namespace pyr {

struct PathCore;    // shading.cuh
struct PendingLight;
struct BidirState;
struct LightVertex; // bdpt.cuh
struct CamVertex;

// Device-side work counters (one instance per context).
struct DeviceCounters {
    unsigned long long rays, path_samples, nodes_visited, leaves_tested, de_evals, de_iterations, node_fetches, path_rays;
    unsigned long long march_overflow;
    unsigned long long march_iterations, julia_iterations;  // stats mode: estimator iterations run by k_march, and the quaternion-Julia part of them  // sphere-tracing candidates that did not fit their queue (must stay 0: pyr_render fails otherwise)
};

// Everything one wavefront iteration needs besides the scene.
struct WaveArgs {
    PathCore* paths;
    PendingLight* pend;        // pool * MAX_LIGHT_SAMPLES
    BidirState* bidir;         // bidirectional only
    uint32_t pool;
    uint32_t grid_paths;       // upper bound of the live-slot count of this pass (sizes the k_bin / shade grids)
    const Ray* rays_in;        // rays traced in the previous iteration (read by the shade stage)
    const Hit* hits_in;
    const uint32_t* shadow_kinds_in;  // hit kind per visibility ray of the previous pass (KIND_MISS = unblocked)
    Ray* rays_out;             // rays emitted by this iteration: path rays in [0, pool), visibility rays from `shadow_offset`
    uint32_t* count_out;       // [0] path rays, [1] visibility rays emitted (device counters, pre-zeroed)
    uint32_t shadow_offset;    // first index of the visibility-ray region in rays / hits
    uint32_t* trace_cursor;    // work cursor of the trace kernel (reset here)
    unsigned long long* next_sample;  // next global path-sample index to start
    unsigned long long total_samples;
    const unsigned long long* tile_first;  // n_tiles + 1 prefix sums of per-tile sample counts
    uint64_t seed;
    uint32_t sample_offset, sample_stride;
    float* film;               // (accumulator, weight) pairs
    DeviceCounters* counters;
    LightVertex* light_vertices;  // bidirectional: pool * light_stride lamp-subpath vertices
    CamVertex* cam_vertices;      // bidirectional: pool * cam_stride stored camera-subpath vertices
    uint32_t light_stride, cam_stride;
    uint32_t ray_capacity;
    // live slots sorted by (shading state, spatial cluster) (k_bin_*): bin_first[k] = start of key k's run in bin_list,
    // bin_first[NUM_KEYS] = number of entries
    const uint32_t* bin_first;
    const uint32_t* bin_list;
    // slots that can still do something: written by the shade kernel (slots alive after it), read by the next k_bin;
    // once every sample has started the list shrinks with the paths still in flight
    uint32_t* live_list;
    const uint32_t* live_count_in;
    uint32_t* live_count_out;
    // bidirectional: slots whose sample ended in this iteration's phase kernels (the GEN kernel restarts them in the same iteration)
    uint32_t* died_list;
    uint32_t* died_count;
};
#ifndef PYR_BIN_CLUSTERS
#define PYR_BIN_CLUSTERS 16
#endif
#ifndef PYR_BIN_STATES
#define PYR_BIN_STATES 48
#endif
constexpr uint32_t BIN_STATES = PYR_BIN_STATES, BIN_CLUSTERS = PYR_BIN_CLUSTERS, NUM_KEYS = BIN_STATES * BIN_CLUSTERS;
struct BinBuffers {
    uint32_t* count;  // [NUM_KEYS], zero between iterations (k_bin_scan clears it)
    uint32_t* first;  // [NUM_KEYS + 1]
    uint32_t* fill;   // [NUM_KEYS]
    uint16_t* keys;   // [pool], by live-list position
    uint32_t* list;   // [pool]
};
void launch_bin(const WaveArgs& a, const BinBuffers& b, uint32_t cluster_shift, int bidirectional, cudaStream_t s);

struct TraceArgs {
    const Ray* rays;
    Hit* hits;               // closest-hit records of the path rays
    uint32_t* shadow_kinds;  // hit kind per visibility ray
    const uint32_t* count;   // device-resident ray counts: [0] path rays, [1] visibility rays
    uint32_t shadow_offset;
    uint32_t* cursor;        // dynamic work cursor in 32-ray packets (zero on entry)
    DeviceCounters* counters;
    int stats;
    uint32_t refill_min, steps;  // lane-refill threshold and traversal steps between refill checks (tuning)
    // sphere tracing (stage 3): one work item per (ray, ray-marched shape) whose leaf the walk reached, in one queue per
    // distance-estimator type so that the lanes of a warp run the same estimator
    uint2* march_queue[2];       // (ray index, shape index); [0] Mandelbulb, [1] quaternion Julia
    uint32_t* march_count;       // two device counters (zeroed before the launch)
    uint32_t march_capacity[2];  // entries per queue: rays x ray-marched shapes of that estimator type (an exact bound)
    unsigned long long* march_key;  // per path ray: (distance bits, tie rank) of the best hit so far, merged with atomicMin
};
struct TraceTuning { uint32_t refill_min, steps; };
TraceTuning trace_tuning();

// the wavefront
void launch_pool_reset(PathCore* paths, uint32_t pool, uint32_t* live_list, uint32_t* live_count, cudaStream_t s);
void launch_wave_simple(const SceneView& sc, const WaveArgs& a, cudaStream_t s);
void launch_wave_bidirectional(const SceneView& sc, const WaveArgs& a, int sm_count, cudaStream_t s);
int wave_bidirectional_launches();
void launch_trace(const SceneView& sc, const TraceArgs& a, int grid_blocks, cudaStream_t s);
void launch_march(const SceneView& sc, const TraceArgs& a, int grid_blocks, cudaStream_t s);

// ABI seams
void launch_trace_batch(const SceneView& sc, const void* rays32, size_t n, void* hits20, uint32_t* cursor, DeviceCounters* counters, int stats,
                        int grid_blocks, cudaStream_t s);
void launch_film_expose(const SceneView& sc, float* film, const float* positions, const float* samples, size_t n, cudaStream_t s);
void launch_white_scan(const SceneView& sc, float* develop_params, cudaStream_t s);
void launch_develop(const SceneView& sc, const float* film, const float* develop_params, float step_size, float* xyz, uint8_t* srgb, cudaStream_t s);
void launch_camera_sample(const SceneView& sc, uint64_t seed, uint32_t tile, uint64_t sample, float* out /* 2 + 8 + 16 + 1 floats */, cudaStream_t s);

void launch_debug_path(const SceneView& sc, uint64_t seed, uint32_t tile, uint64_t sample, uint32_t max_bounces, uint32_t* records /* 20 words each */,
                       uint32_t* counts /* bounces, exposed */, float* exposed /* 16 x (brightness, wavelength) */, float* position2, void* scratch /* debug_path_scratch_bytes() */,
                       cudaStream_t s);
size_t debug_path_scratch_bytes();

size_t path_state_bytes();
size_t pending_light_bytes();
size_t bidir_state_bytes();
size_t light_vertex_bytes();
size_t cam_vertex_bytes();
int bdpt_stage_rays();
int trace_blocks_per_sm();

// bvh_build.cu: Bvh::new (spatial/bvh.rs:13-155) level by level on the GPU; the tree of the depth-first host builder.
// `device_scratch(bytes)` returns a device block of that size that stays valid for the call (about 350 bytes per item).
void gpu_bvh_build(const float* boxes6, size_t n, const float* root_hull12, BvhTree& out, cudaStream_t stream, const std::function<void*(size_t)>& device_scratch);

}  // namespace pyr
