// sm_100a kernels of the wavefront render path.  Compiled with -fmad=false: the reference never
// contracts a*b+c and hit/primitive IDs must come out bit-exact.
//
//   k_bin_*         a counting sort of the live path slots by (shading state, BVH-subtree cluster of the hit)
//   k_wave_simple   stage 1 + 4 + 5 + 6: regenerate finished path slots (camera ray, wavelengths),
//   k_wave_bidirectional   shade the closest hits of the previous trace pass (surface data, expression VM,
//                   BSDF scatter, next-event estimation, the `contribute` fold; lamp subpaths, connections and
//                   light tracing for BDPT), expose finished samples on the film, and append the next rays to
//                   the ray queue with warp-level prefix sums + one atomic per block
//   k_trace         stage 2: World::intersect for every queued ray (planes, 4-wide BVH walk, triangle /
//                   sphere tests); persistent warps hand ray indices to lanes as they finish
//   k_march         stage 3: sphere tracing of the ray-marched candidates the walk found, flattened to
//                   distance-estimator iterations; k_march_apply merges the winners into the hit records
//   k_develop       stage 6 tail: spectral film -> CIE XYZ -> sRGB
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.hpp"
#include "shading.cuh"
#include "bdpt.cuh"

namespace pyr {

namespace {

constexpr int WAVE_THREADS = 128;
#ifndef WAVE_MIN_BLOCKS
#define WAVE_MIN_BLOCKS 3
#endif
constexpr int TRACE_THREADS = 128;
constexpr unsigned FULL = 0xffffffffu;
// dynamic shared memory, sized by what the scene needs: VM registers (+ visibility-ray staging in the wave kernels)
inline size_t vm_smem(const SceneView& sc) { return (size_t)sc.vm_regs * WAVE_THREADS * sizeof(float4); }
inline size_t wave_smem(const SceneView& sc) {
    return vm_smem(sc) + (size_t)stage_quads(sc.renderer.light_samples) * WAVE_THREADS * sizeof(float4) +
           (size_t)4 * sc.renderer.spectrum_samples * WAVE_THREADS * sizeof(float);  // colour cache + wl / bright / refl
}

struct FilmAdd {
    float* film;
    __device__ __forceinline__ void operator()(uint64_t index, float increment, float weight) const {
        atomicAdd(reinterpret_cast<float2*>(film) + index, make_float2(increment, weight));
    }
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// exclusive prefix sum of `n` over the warp; returns the warp total in `total`
__device__ __forceinline__ uint32_t warp_exclusive_scan(uint32_t n, uint32_t& total) {
    uint32_t inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (lane_id() >= (uint32_t)d) inc += t;
    }
    total = __shfl_sync(FULL, inc, 31);
    return inc - n;
}

// Hand out global path-sample indices to the lanes that want one.
__device__ __forceinline__ bool claim_sample(unsigned long long* next, unsigned long long total, bool want, unsigned long long& index) {
    unsigned mask = __ballot_sync(FULL, want);
    if (!mask) return false;
    unsigned long long base = 0;
    int leader = __ffs(mask) - 1;
    if ((int)lane_id() == leader) base = atomicAdd(next, (unsigned long long)__popc(mask));
    base = __shfl_sync(FULL, base, leader);
    if (!want) return false;
    index = base + __popc(mask & ((1u << lane_id()) - 1u));
    return index < total;
}

// global sample index -> (tile, per-tile sample number); tile_first has n_tiles + 1 entries
__device__ __forceinline__ void locate_sample(const unsigned long long* tile_first, uint32_t n_tiles, unsigned long long g, uint32_t& tile,
                                              unsigned long long& k) {
    uint32_t lo = 0, hi = n_tiles;  // tile_first[lo] <= g < tile_first[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tile_first[mid] <= g) lo = mid; else hi = mid;
    }
    tile = lo;
    k = g - tile_first[lo];
}

// Path state and the ray / hit queues are touched once per iteration: they go through the caches with the
// streaming hint (ld/st.global.cs) so that they do not evict the BVH and the thread-local lines from L2.
__device__ __forceinline__ void store_ray(Ray* dst, const Ray& r) { store_record_stream(dst, r); }
// Dynamic shared memory of the shade kernels, in float4 units: VM registers [vm_regs][thread]; staged visibility
// rays [1 + light_samples][thread] (shared origin + one quad per ray); then floats: the fold's colour cache [S][thread] and the path's wl / bright / refl
// arrays [3 * S][thread].
// LEAN: the layout of kernels that run no expression program and stage no rays (the bidirectional connect / splat phases):
// only the per-wavelength arrays, from the start of the dynamic shared memory - less shared memory, more resident blocks.
template <bool LEAN = false>
__device__ __forceinline__ float* spectral_base(const SceneView& sc) {
#if defined(__CUDA_ARCH__)
    if (LEAN) return reinterpret_cast<float*>(pyr_dyn_smem) + threadIdx.x;
    float* floats = reinterpret_cast<float*>(pyr_dyn_smem + (sc.vm_regs + stage_quads(sc.renderer.light_samples)) * PYR_BLOCK);
    return floats + sc.renderer.spectrum_samples * PYR_BLOCK + threadIdx.x;
#else
    return nullptr;
#endif
}
template <bool LEAN = false>
__device__ __forceinline__ void bind_spectral(const SceneView& sc, PathState& ps) {
    float* base = spectral_base<LEAN>(sc);
    const uint32_t S = sc.renderer.spectrum_samples;
    ps.wl.base = base;
    ps.bright.base = base + S * WAVE_THREADS;
    ps.refl.base = base + 2u * S * WAVE_THREADS;
}
// header -> registers, per-wavelength arrays -> shared memory; 32-byte chunks
__device__ __forceinline__ void load_core(const SceneView& sc, PathState& ps, const PathCore* src) {
    static_assert(sizeof(PathHeader) == 64, "two chunks");
    const Vec8* s = reinterpret_cast<const Vec8*>(src);
    Vec8* h = reinterpret_cast<Vec8*>(static_cast<PathHeader*>(&ps));
    h[0] = ld256_stream(s);
    h[1] = ld256_stream(s + 1);
    const uint32_t n = 3u * sc.renderer.spectrum_samples;
    float* spectral = ps.wl.base;  // wl | bright | refl are contiguous in shared memory, stride S like the record
    for (uint32_t c = 0; c * 8 < n; ++c) {
        const Vec8 v = ld256_stream(s + 2 + c);
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j)
            if (c * 8 + j < n) spectral[(c * 8 + j) * WAVE_THREADS] = v.v[j];
    }
}
__device__ __forceinline__ void store_core(const SceneView& sc, PathCore* dst, const PathState& ps) {
    Vec8* d = reinterpret_cast<Vec8*>(dst);
    const Vec8* h = reinterpret_cast<const Vec8*>(static_cast<const PathHeader*>(&ps));
    st256_stream(d, h[0]);
    st256_stream(d + 1, h[1]);
    const uint32_t n = 3u * sc.renderer.spectrum_samples;
    const float* spectral = ps.wl.base;
    for (uint32_t c = 0; c * 8 < n; ++c) {
        Vec8 v;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) v.v[j] = c * 8 + j < n ? spectral[(c * 8 + j) * WAVE_THREADS] : 0.0f;
        st256_stream(d + 2 + c, v);
    }
}
__device__ __forceinline__ Ray load_ray(const Ray* src) { return load_record_stream(src); }

__global__ void k_pool_reset(PathCore* paths, uint32_t pool, uint32_t* live_list, uint32_t* live_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pool) { paths[i].flags = 0; live_list[i] = i; }  // every slot may start a path
    if (i == 0) *live_count = pool;
}

// ------------------------------------------------------------------------------------------
// Shading-state key of a path slot: threads of a warp that share it take the same way through the
// shade stage (regenerate / miss / surface hit, and the number of unblocked light samples to fold).
__device__ __forceinline__ uint32_t unblocked_count(const uint32_t* kinds, uint32_t at, uint32_t n) {
    uint32_t c = 0;
    for (uint32_t j = 0; j < n; ++j) c += __ldg(&kinds[at + j]) == KIND_MISS ? 1u : 0u;
    return c;
}
// Shading-state key x spatial cluster of a live slot.  The cluster is the top bits of the hit primitive's rank: ranks are
// the BVH's leaf pre-order, so equal top bits mean the same subtree, i.e. the same region of the scene - the next path ray
// and the visibility rays of neighbouring threads then start close together (and, for the light samples, head for the same
// lamp), which keeps the lanes of a traversal warp on the same nodes, and their surface / material records share cache lines.
__device__ __forceinline__ uint32_t bin_key(const PathCore* paths, const BidirState* bidir, uint32_t slot, const Hit* hits, const uint32_t* shadow_kinds,
                                            uint32_t cluster_shift) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(paths + slot) + 2);  // flags, n_pending, ray_base, shadow_base
    const uint32_t flags = h.x, n_pending = h.y, ray_base = h.z, shadow_base = h.w;
    if (!(flags & PS_ALIVE)) return 0;
    uint32_t phase = PH_CAMERA, n_light = 0;
    if (bidir) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(bidir + slot));  // phase, n_light
        phase = b.x; n_light = b.y;
    }
    uint32_t state, cluster = 0;
    if (phase == PH_CAMERA || phase == PH_LAMP) {
        uint32_t hc = 0;
        if (phase == PH_LAMP || (flags & PS_HAS_MAIN)) {
            const uint32_t kind = __ldg(&hits[ray_base].kind);
            hc = kind == KIND_MISS ? 1u : (kind == KIND_PLANE ? 2u : (kind == KIND_RAY_MARCHED ? 4u : 3u));
            if (hc == 3u) cluster = min(__ldg(&hits[ray_base].rank) >> cluster_shift, BIN_CLUSTERS - 1u);
            else if (hc == 2u) cluster = __ldg(&hits[ray_base].rank) & (BIN_CLUSTERS - 1u);
        }
        if (phase == PH_CAMERA) {
            const uint32_t lit = (flags & PS_PENDING_FOLD) ? min(unblocked_count(shadow_kinds, shadow_base, n_pending), 4u) : 0u;
            state = 1u + hc * 5u + lit;   // 1 .. 25
        } else {
            state = 26u + hc;             // 27 .. 30
        }
    } else if (phase == PH_FINISH) {
        state = 31u;                                       // a finished lamp subpath waiting for its fix-up / colours / folds
        cluster = min(n_light, BIN_CLUSTERS - 1u);         // equal path lengths together: the loops run over the lamp vertices
    } else {
        const uint32_t lit = min(unblocked_count(shadow_kinds, shadow_base, n_pending), 14u) >> 1;
        state = (phase == PH_CONNECT ? 32u : 40u) + lit;  // 32 .. 47
        cluster = min(n_light, BIN_CLUSTERS - 1u);         // connection / splat loops run over the lamp path: equal lengths together
    }
    return state * BIN_CLUSTERS + cluster;
}

// Counting sort of the live slots by key, in three small kernels per iteration:
//   k_bin_keys     key per live-list entry (2 bytes) + per-block histogram in shared memory, flushed with one atomic per
//                  non-empty key and block
//   k_bin_scan     one block: exclusive scan of the key counts -> first[]; clears the counters for the next iteration
//   k_bin_scatter  per-block histogram again (from the stored keys), one atomic per non-empty key and block reserves the
//                  block's range in each key's run, then the slot ids are scattered into ONE pool-sized list
constexpr int BIN_THREADS = 256, BIN_ITEMS = 8;
__device__ __forceinline__ uint32_t live_slot(const uint32_t* live_list, uint32_t n_live, uint32_t pool, uint32_t index) {
    // while every slot is alive the list is a permutation of [0, pool): walk the slots in order instead (coalesced)
    return n_live == pool ? index : live_list[index];
}
__global__ void __launch_bounds__(BIN_THREADS) k_bin_keys(const PathCore* paths, const BidirState* bidir, uint32_t pool, const Hit* hits, const uint32_t* shadow_kinds,
                                                          uint32_t cluster_shift, uint32_t* bin_count, uint16_t* keys, const uint32_t* live_list,
                                                          const uint32_t* live_count) {
    extern __shared__ uint32_t s_bins[];  // [NUM_KEYS]
    uint32_t* const s_count = s_bins;
    for (uint32_t k = threadIdx.x; k < NUM_KEYS; k += BIN_THREADS) s_count[k] = 0;
    __syncthreads();
    const uint32_t n_live = *live_count;
    const uint32_t first = blockIdx.x * (BIN_THREADS * BIN_ITEMS);
#pragma unroll 2
    for (int i = 0; i < BIN_ITEMS; ++i) {
        const uint32_t index = first + i * BIN_THREADS + threadIdx.x;
        if (index < n_live) {
            const uint32_t key = bin_key(paths, bidir, live_slot(live_list, n_live, pool, index), hits, shadow_kinds, cluster_shift);
            keys[index] = (uint16_t)key;
            atomicAdd(&s_count[key], 1u);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < NUM_KEYS; k += BIN_THREADS) {
        const uint32_t c = s_count[k];
        if (c) atomicAdd(&bin_count[k], c);
    }
}
__global__ void __launch_bounds__(1024) k_bin_scan(uint32_t* bin_count, uint32_t* bin_first, uint32_t* bin_fill, uint32_t* queue_counts, uint32_t* live_count_out,
                                                   uint32_t* died_count) {
    // also clears what the shade pass of this iteration counts into: its ray-queue counters, its live-slot count and its died-slot count
    if (threadIdx.x == 0) { queue_counts[0] = 0; queue_counts[1] = 0; *live_count_out = 0; *died_count = 0; }
    constexpr int PER = (NUM_KEYS + 1023) / 1024;
    __shared__ uint32_t s_warp[32];
    uint32_t c[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { c[j] = threadIdx.x * PER + j < NUM_KEYS ? bin_count[threadIdx.x * PER + j] : 0u; sum += c[j]; }
    uint32_t warp_total;
    uint32_t before = warp_exclusive_scan(sum, warp_total);
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = warp_total;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t t;
        const uint32_t w = warp_exclusive_scan(s_warp[threadIdx.x], t);
        s_warp[threadIdx.x] = w;
        if (threadIdx.x == 31) bin_first[NUM_KEYS] = t;
    }
    __syncthreads();
    before += s_warp[threadIdx.x >> 5];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (threadIdx.x * PER + j >= NUM_KEYS) break;
        bin_first[threadIdx.x * PER + j] = before;
        before += c[j];
        bin_count[threadIdx.x * PER + j] = 0;
        bin_fill[threadIdx.x * PER + j] = 0;
    }
}
__global__ void __launch_bounds__(BIN_THREADS) k_bin_scatter(uint32_t pool, const uint16_t* keys, const uint32_t* bin_first, uint32_t* bin_fill, uint32_t* bin_list,
                                                             const uint32_t* live_list, const uint32_t* live_count) {
    extern __shared__ uint32_t s_bins[];  // [2][NUM_KEYS]
    uint32_t* const s_count = s_bins;
    uint32_t* const s_base = s_bins + NUM_KEYS;
    for (uint32_t k = threadIdx.x; k < NUM_KEYS; k += BIN_THREADS) s_count[k] = 0;
    __syncthreads();
    const uint32_t n_live = *live_count;
    const uint32_t first = blockIdx.x * (BIN_THREADS * BIN_ITEMS);
    uint32_t key[BIN_ITEMS], local[BIN_ITEMS];
#pragma unroll
    for (int i = 0; i < BIN_ITEMS; ++i) {
        const uint32_t index = first + i * BIN_THREADS + threadIdx.x;
        key[i] = 0xFFFFFFFFu;
        if (index < n_live) {
            key[i] = keys[index];
            local[i] = atomicAdd(&s_count[key[i]], 1u);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < NUM_KEYS; k += BIN_THREADS) {
        const uint32_t c = s_count[k];
        if (c) s_base[k] = bin_first[k] + atomicAdd(&bin_fill[k], c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < BIN_ITEMS; ++i) {
        const uint32_t index = first + i * BIN_THREADS + threadIdx.x;
        if (key[i] != 0xFFFFFFFFu) bin_list[s_base[key[i]] + local[i]] = live_slot(live_list, n_live, pool, index);
    }
}

// Queue space for one block of a shade kernel: path rays, visibility rays and the live-slot list (the slots still
// alive after this pass are the domain of the next k_bin).  Warp totals meet in shared memory and one thread
// issues the block's atomics back to back, so a block pays one L2 round trip instead of three per warp.
struct Reservation { uint32_t main_at, shadow_at, live_at; };
__device__ __forceinline__ Reservation block_reserve(const WaveArgs& a, uint32_t n_main, uint32_t n_shadow, bool alive) {
    constexpr int WARPS = WAVE_THREADS / 32;
    __shared__ uint32_t s_total[3][WARPS];
    __shared__ uint32_t s_base[3];
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t total_main, total_shadow;
    Reservation r;
    r.main_at = warp_exclusive_scan(n_main, total_main);
    r.shadow_at = warp_exclusive_scan(n_shadow, total_shadow);
    const unsigned live_mask = __ballot_sync(FULL, alive);
    r.live_at = __popc(live_mask & ((1u << lane_id()) - 1u));
    if (lane_id() == 0) { s_total[0][warp] = total_main; s_total[1][warp] = total_shadow; s_total[2][warp] = __popc(live_mask); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t[3] = {0, 0, 0};
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { t[0] += s_total[0][w]; t[1] += s_total[1][w]; t[2] += s_total[2][w]; }
        // count_out[0] / [1] are adjacent and 8-byte aligned: one 64-bit add reserves both regions
        unsigned long long both = 0;
        uint32_t live = 0;
        if (t[0] | t[1]) both = atomicAdd(reinterpret_cast<unsigned long long*>(a.count_out), (unsigned long long)t[0] | ((unsigned long long)t[1] << 32));
        if (t[2]) live = atomicAdd(a.live_count_out, t[2]);
        s_base[0] = (uint32_t)both; s_base[1] = (uint32_t)(both >> 32); s_base[2] = live;
    }
    __syncthreads();
    uint32_t before[3] = {s_base[0], s_base[1], s_base[2]};
#pragma unroll
    for (int w = 0; w < WARPS; ++w)
        if ((uint32_t)w < warp) { before[0] += s_total[0][w]; before[1] += s_total[1][w]; before[2] += s_total[2][w]; }
    r.main_at += before[0]; r.shadow_at += before[1]; r.live_at += before[2];
    return r;
}

// thread -> slot through the sorted list (a permutation of the live list); the run of key 0 holds the slots without a live path
__device__ __forceinline__ uint32_t binned_slot(const WaveArgs& a, uint32_t g, bool& valid, bool& dead) {
    valid = g < __ldg(a.bin_first + NUM_KEYS);
    dead = g < __ldg(a.bin_first + 1);
    return valid ? a.bin_list[g] : 0u;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WAVE_THREADS, WAVE_MIN_BLOCKS) k_wave_simple(const __grid_constant__ SceneView sc, const __grid_constant__ WaveArgs a) {
    const uint32_t g_thread = blockIdx.x * blockDim.x + threadIdx.x;
    if (g_thread == 0) *a.trace_cursor = 0;
    bool valid, dead;
    // blocks beyond the sorted list have nothing to do (the grid is sized by a bound of the live count that can be stale)
    if (blockIdx.x * blockDim.x >= __ldg(a.bin_first + NUM_KEYS)) return;
    const uint32_t slot = binned_slot(a, g_thread, valid, dead);
    // a whole warp of dead slots with no samples left to start has nothing to do (long-tailed scenes)
    const bool idle = __all_sync(FULL, !valid || dead) && *a.next_sample >= a.total_samples;
    PathState ps;
    bind_spectral(sc, ps);
    ps.flags = 0;
    uint32_t flags_in = 0;
    ShadeOut out;
    out.alive = 0; out.has_main = 0; out.n_shadow = 0;
    out.stage_base = sc.vm_regs * WAVE_THREADS;
    PathCounters pc;
    pc.de_evals = 0; pc.de_iters = 0;
    bool alive = false;
    if (!idle) {
        if (valid && !dead) load_core(sc, ps, a.paths + slot);
        ps.pend = a.pend + (size_t)(valid ? slot : 0) * MAX_LIGHT_SAMPLES;
        ps.bd = nullptr;
        flags_in = ps.flags;
        FilmAdd add{a.film};
        alive = valid && (ps.flags & PS_ALIVE);
        if (alive) {
            shade_simple(sc, ps, a.rays_in + ps.ray_base, a.hits_in + ps.ray_base, a.rays_in + a.shadow_offset + ps.shadow_base,
                         a.shadow_kinds_in + ps.shadow_base, out, add, pc);
            alive = out.alive != 0;
            if (!alive) ps.flags = 0;
        }
        unsigned long long g = 0;
        if (claim_sample(a.next_sample, a.total_samples, valid && !alive, g)) {
            uint32_t tile; unsigned long long k;
            locate_sample(a.tile_first, sc.n_tiles, g, tile, k);
            generate_simple(sc, a.seed, tile, (uint64_t)a.sample_offset + k * a.sample_stride, ps, out.main);
            ps.flags |= PS_ALIVE;
            out.has_main = 1; out.n_shadow = 0; out.alive = 1;
            alive = true;
        }
    }
    // path rays and visibility rays go to separate queue regions so that trace packets are homogeneous
    const uint32_t n_main = alive ? out.has_main : 0u, n_shadow = alive ? out.n_shadow : 0u;
    const Reservation at = block_reserve(a, n_main, n_shadow, valid && alive);
    if (n_main) { ps.ray_base = at.main_at; store_ray(a.rays_out + at.main_at, out.main); }
    if (n_shadow) {
        ps.shadow_base = at.shadow_at;
        for (uint32_t j = 0; j < n_shadow; ++j) store_ray(a.rays_out + a.shadow_offset + at.shadow_at + j, out.get_shadow(j));
    }
    if (valid && (alive || (flags_in & PS_ALIVE))) store_core(sc, a.paths + slot, ps);
    if (valid && alive) a.live_list[at.live_at] = slot;
    if (pc.de_evals) { atomicAdd(&a.counters->de_evals, (unsigned long long)pc.de_evals); atomicAdd(&a.counters->de_iterations, (unsigned long long)pc.de_iters); }
}

// ------------------------------------------------------------------------------------------
// Traversal stack: the first SHARED_STACK entries live in shared memory ([entry][thread], no bank
// conflicts), deeper entries (rare) in local memory.
#ifndef PYR_SHARED_STACK
#define PYR_SHARED_STACK 20
#endif
constexpr int SHARED_STACK = PYR_SHARED_STACK;
struct SharedStack {
    int2* entries;  // &smem[threadIdx.x]: (child code, entry distance bits), one 8-byte access per push / pop
    static constexpr int FAST_DEPTH = SHARED_STACK;  // slots below it need no region test (Traversal::node_step / pop)
    __device__ __forceinline__ void put_fast(int i, int code, float dist) { entries[i * TRACE_THREADS] = make_int2(code, __float_as_int(dist)); }
    __device__ __forceinline__ int code_fast(int i) const { return entries[i * TRACE_THREADS].x; }
    __device__ __forceinline__ float dist_fast(int i) const { return __int_as_float(entries[i * TRACE_THREADS].y); }
    int2 deep[BVH_STACK - SHARED_STACK];
    __device__ __forceinline__ void put(int i, int code, float dist) {
        const int2 e = make_int2(code, __float_as_int(dist));
        if (i < SHARED_STACK) entries[i * TRACE_THREADS] = e; else deep[i - SHARED_STACK] = e;
    }
    __device__ __forceinline__ int2 at(int i) const { return i < SHARED_STACK ? entries[i * TRACE_THREADS] : deep[i - SHARED_STACK]; }
    __device__ __forceinline__ int code(int i) const { return at(i).x; }
    __device__ __forceinline__ float dist(int i) const { return __int_as_float(at(i).y); }
    // one more row behind the stack: max(hit distance, box entry distance) of the current closest hit (written when a hit is
    // taken, read only when a later candidate precedes it in the reference's order) - kept out of the registers of the walk
    __device__ __forceinline__ void set_best_m(float v) { entries[SHARED_STACK * TRACE_THREADS].y = __float_as_int(v); }
    __device__ __forceinline__ float best_m() const { return __int_as_float(entries[SHARED_STACK * TRACE_THREADS].y); }
};

// Persistent warps with lane-level refill: a warp reserves 32 ray indices with one atomicAdd and hands
// them to its lanes as they finish, so that lanes whose rays end early (misses, occluded visibility
// rays) do not idle until the longest ray of a packet is done.  Ray index space: [0, n_main) path rays,
// then n_shadow visibility rays stored from `shadow_offset`.
template <bool STATS, class Emit>
__device__ __forceinline__ void trace_persistent(const SceneView& sc, const Ray* rays, uint32_t n_main, uint32_t n_shadow, uint32_t shadow_offset,
                                                 uint32_t* cursor, DeviceCounters* counters, bool closest_only, uint32_t refill_min, uint32_t steps,
                                                 Emit& emit) {
    __shared__ int2 smem_stack[(SHARED_STACK + 1) * TRACE_THREADS];
    SharedStack stack;
    stack.entries = smem_stack + threadIdx.x;
    const uint32_t total = n_main + n_shadow;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters->rays, (unsigned long long)total);
        atomicAdd(&counters->path_rays, (unsigned long long)n_main);
    }
    unsigned long long nodes = 0, leaves = 0, evals = 0, iters = 0, fetches = 0;
    Traversal<STATS> tr;
    tr.done = true;
    bool has_ray = false;
    uint32_t ray_at = 0;
    uint32_t pool_base = 0, pool_left = 0;  // warp-uniform: reserved, not yet assigned ray indices
    bool exhausted = false;                 // warp-uniform
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !has_ray);
        if ((uint32_t)__popc(idle) >= refill_min || (idle == FULL)) {
            if (pool_left == 0 && !exhausted) {
                uint32_t base = 0;
                if (lane_id() == 0) base = atomicAdd(cursor, 32u);
                base = __shfl_sync(FULL, base, 0);
                if (base >= total) exhausted = true;
                else { pool_base = base; pool_left = min(32u, total - base); }
            }
            if (pool_left) {
                const uint32_t my = __popc(idle & ((1u << lane_id()) - 1u));
                const uint32_t taken = min((uint32_t)__popc(idle), pool_left);
                if (!has_ray && my < taken) {
                    const uint32_t idx = pool_base + my;
                    ray_at = idx < n_main ? idx : shadow_offset + (idx - n_main);
                    Ray r = load_ray(rays + ray_at);
                    if (closest_only) { r.mode = 0; r.limit = 0.0f; }
                    tr.begin(sc, r, stack);
                    has_ray = true;
                }
                pool_base += taken;
                pool_left -= taken;
            } else if (idle == FULL && exhausted) {
                break;
            } else if (exhausted) {
                refill_min = 33;  // nothing left to hand out: stop checking until the warp drains
            }
        }
        // interior nodes first, `steps` of them; then the lanes that stand at a leaf test it together (lanes
        // that reach a leaf early wait, which costs less than running the leaf code for one lane at a time)
        for (uint32_t k = 0; k < steps; ++k)
            if (has_ray && !tr.done && tr.at_node()) tr.node_step(sc, stack);
        if (has_ray && !tr.done && !tr.at_node()) tr.leaf_step(sc, stack);
        if (has_ray && tr.done) {
            emit(ray_at, tr);
            if (STATS) { nodes += tr.vn; leaves += tr.vl; evals += tr.de_evals; iters += tr.de_iters; fetches += tr.vf; }
            has_ray = false;
        }
    }
    if (STATS) {
        for (int d = 16; d; d >>= 1) {
            nodes += __shfl_down_sync(FULL, nodes, d); leaves += __shfl_down_sync(FULL, leaves, d);
            evals += __shfl_down_sync(FULL, evals, d); iters += __shfl_down_sync(FULL, iters, d);
            fetches += __shfl_down_sync(FULL, fetches, d);
        }
        if (lane_id() == 0) {
            atomicAdd(&counters->nodes_visited, nodes); atomicAdd(&counters->leaves_tested, leaves);
            atomicAdd(&counters->de_evals, evals); atomicAdd(&counters->de_iterations, iters);
            atomicAdd(&counters->node_fetches, fetches);
        }
    }
}

// (distance, tie rank) packed so that an unsigned 64-bit minimum is World::intersect's choice: smaller distance first
// (positive floats order like their bit patterns), then planes before BVH leaves, then the earlier leaf (lower rank)
__device__ __forceinline__ unsigned long long pack_hit(float t, uint32_t kind, uint32_t rank) {
    if (kind == KIND_MISS) return ~0ull;
    const uint32_t tie = kind == KIND_PLANE ? rank : (0x80000000u | rank);
    return ((unsigned long long)__float_as_uint(t) << 32) | tie;
}

struct EmitHit {
    Hit* hits;
    uint32_t* shadow_kinds;
    uint32_t shadow_offset;
    uint2* march_queue0;
    uint2* march_queue1;
    uint32_t* march_count;
    uint32_t march_capacity[2];
    unsigned long long* march_key;
    const MarchedRec* marched;
    DeviceCounters* counters;
    template <class T>
    __device__ __forceinline__ void operator()(uint32_t at, const T& tr) const {
        // rays that reached ray-marched leaves (and are not decided yet) go on to the sphere-tracing kernels
        if (tr.march_mask && !(tr.mode != 0 && tr.kind != KIND_MISS)) {
            uint32_t m = tr.march_mask;
            while (m) {
                const uint32_t shape = (uint32_t)(__ffs((int)m) - 1);
                m &= m - 1u;
                const uint32_t type = marched[shape].estimator ? 1u : 0u;
                // one atomicAdd per queue for the lanes that are here together
                const unsigned peers = __match_any_sync(__activemask(), type);
                const int leader = __ffs((int)peers) - 1;
                uint32_t base = 0;
                if ((int)lane_id() == leader) base = atomicAdd(march_count + type, (uint32_t)__popc(peers));
                base = __shfl_sync(peers, base, leader);
                const uint32_t pos = base + __popc(peers & ((1u << lane_id()) - 1u));
                // the queues hold rays x shapes-of-this-type entries, so this cannot overflow; if it ever does the render fails loudly
                if (pos < march_capacity[type]) (type ? march_queue1 : march_queue0)[pos] = make_uint2(at, shape);
                else atomicAdd(&counters->march_overflow, 1ull);
            }
            if (at < shadow_offset) march_key[at] = pack_hit(tr.t, tr.kind, tr.rank);
        }
        if (at >= shadow_offset) { shadow_kinds[at - shadow_offset] = tr.kind; return; }  // visibility rays only report blocked / unblocked
        Hit h;
        h.t = tr.t; h.u = tr.u; h.v = tr.v; h.rank = tr.rank; h.kind = tr.kind; h.nodes = 0; h.leaves = 0; h.pad = 0;
        store_record(hits + at, h);
    }
};

template <bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace(const __grid_constant__ SceneView sc, const __grid_constant__ TraceArgs a) {
    EmitHit emit{a.hits, a.shadow_kinds, a.shadow_offset, a.march_queue[0], a.march_queue[1], a.march_count, {a.march_capacity[0], a.march_capacity[1]}, a.march_key, sc.marched, a.counters};
    trace_persistent<STATS>(sc, a.rays, a.count[0], a.count[1], a.shadow_offset, a.cursor, a.counters, false, a.refill_min, a.steps, emit);
}

// Stage 3: sphere tracing (Shape::RayMarched, shapes/mod.rs:120-154 + shapes/distance_estimators.rs:12-70) for the
// (ray, shape) pairs whose leaf the walk reached.  Persistent warps with lane-level refill, one queue per estimator
// type, and the march flattened down to the estimator's inner iteration: one loop trip is ONE iteration of
// `Mandelbulb::get` / `QuaternionJulia::get` for every lane, whatever march step or ray the lane is at.  March
// lengths (a handful to thousands of steps) and escape times (2 to `iterations` trips) vary per ray, so any coarser
// unit leaves most lanes waiting.  Results merge per ray with a 64-bit atomicMin on (distance, tie rank).
template <bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS) k_march(const __grid_constant__ SceneView sc, const __grid_constant__ TraceArgs a, uint32_t* cursors) {
    // every warp works on one estimator type at a time (half of them start with each) and moves to the other queue
    // when its own is empty, so that both queues drain together and their long tails overlap
    uint32_t TYPE = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) & 1u;
    bool switched = false;
    uint32_t n = min(a.march_count[TYPE], a.march_capacity[TYPE]);
    const uint2* queue = a.march_queue[TYPE];
    uint32_t* cursor = cursors + TYPE;
    unsigned long long evals = 0, iters = 0, julia_iters = 0;
    bool active = false, in_de = false;
    uint32_t at = 0, shape = 0, mode = 0, it = 0;
    float limit = 0.0f, total = 0.0f, hi = 0.0f, r = 0.0f, dr = 1.0f;
    v3 dir = mk3(0, 0, 0), origin = mk3(0, 0, 0), point = mk3(0, 0, 0);
    f4 z = mk4(0, 0, 0, 0), dz = mk4(1, 0, 0, 0);
    uint32_t pool_base = 0, pool_left = 0;
    bool exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle) {
            if (pool_left == 0 && !exhausted) {
                uint32_t base = 0;
                if (lane_id() == 0) base = atomicAdd(cursor, 32u);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) exhausted = true;
                else { pool_base = base; pool_left = min(32u, n - base); }
            }
            if (pool_left) {
                const uint32_t my = __popc(idle & ((1u << lane_id()) - 1u));
                const uint32_t taken = min((uint32_t)__popc(idle), pool_left);
                if (!active && my < taken) {
                    const uint2 item = queue[pool_base + my];
                    at = item.x; shape = item.y;
                    const Ray ray = load_ray(a.rays + at);
                    const MarchedRec& mr = sc.marched[shape];
                    const v3 o = ld3(ray.o);
                    dir = ld3(ray.d); mode = ray.mode; limit = ray.limit;
                    float lo;
                    if (bounds_test(mr, o, dir, lo, hi)) {       // Shape::RayMarched branch of ray_intersect
                        // nothing behind the walk's closest hit can win: the march only moves forward from `lo`
                        const float best = at < a.shadow_offset ? __uint_as_float((uint32_t)(a.march_key[at] >> 32)) : PYR_INF;
                        if (!(lo > best)) {                      // (`best` is NaN-patterned for a miss: the test passes)
                            origin = o + (-bounds_center(mr));
                            total = lo;
                            if (total < hi) { active = true; in_de = false; }
                            else if (total <= hi && total > DIST_EPSILON && total < PYR_INF) {  // the loop body never runs
                                if (mode == 0) atomicMin(a.march_key + at, pack_hit(total, KIND_RAY_MARCHED, mr.rank));
                                else if (occludes(mode, total, limit)) a.shadow_kinds[at - a.shadow_offset] = KIND_RAY_MARCHED;
                            }
                        }
                    }
                }
                pool_base += taken;
                pool_left -= taken;
            } else if (idle == FULL && exhausted) {
                if (switched) break;
                switched = true;
                TYPE ^= 1u;
                n = min(a.march_count[TYPE], a.march_capacity[TYPE]);
                queue = a.march_queue[TYPE];
                cursor = cursors + TYPE;
                exhausted = false;
                continue;
            }
        }
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            if (!active) continue;
            const MarchedRec& mr = sc.marched[shape];
            bool de_done = false;
            if (!in_de) {  // start one distance estimate at the current march position
                point = origin + dir * total;
                if (STATS) ++evals;
                it = 0; r = 0.0f; in_de = true;
                if (TYPE == 0) { z = mk4(point.x, point.y, point.z, 0.0f); dr = 1.0f; }
                else { z = mk4(point.x, point.y, point.z, mr.slice_plane); dz = mk4(1.0f, 0.0f, 0.0f, 0.0f); }
            }
            if (it >= mr.iterations) de_done = true;
            else if (TYPE == 0) {  // one trip of Mandelbulb::get's loop (distance_estimators.rs:20-37)
                const v3 zz = mk3(z.x, z.y, z.z);
                r = length(zz);
                if (r > mr.threshold) de_done = true;
                else {
                    if (STATS) ++iters;
                    float theta = acosf(zz.z / r);
                    float phi = atan2f(zz.y, zz.x);
                    const float dc = mr.has_constant ? 0.0f : 1.0f;
                    dr = powf(r, mr.power - 1.0f) * mr.power * dr + dc;
                    const float zr = powf(r, mr.power);
                    theta *= mr.power;
                    phi *= mr.power;
                    v3 nz = mk3(zr * sinf(theta) * cosf(phi), zr * sinf(phi) * sinf(theta), zr * cosf(theta));
                    nz = nz + (mr.has_constant ? ld3(mr.mb_constant) : point);
                    z = mk4(nz.x, nz.y, nz.z, 0.0f);
                    ++it;
                }
            } else {               // one trip of QuaternionJulia::get's loop (distance_estimators.rs:57-66)
                r = qlength(z);
                if (r > mr.threshold) de_done = true;
                else {
                    if (STATS) { ++iters; ++julia_iters; }
                    if (mr.variant == 0) { dz = scale4(qmul(dz, z), 2.0f); z = qmul(z, z); }
                    else if (mr.variant == 1) { dz = scale4(qmul(qmul(dz, z), z), 3.0f); z = qmul(qmul(z, z), z); }
                    else { dz = scale4(bicomplex_mul(bicomplex_mul(dz, z), z), 2.0f); z = bicomplex_mul(z, z); }
                    z = add4(z, mr.constant);
                    ++it;
                }
            }
            if (!de_done) continue;
            // the estimate is complete: one step of the march loop (shapes/mod.rs:127-135)
            const float distance = TYPE == 0 ? 0.5f * logf(r) * r / dr : 0.5f * logf(r) * r / qlength(dz);
            in_de = false;
            const bool stuck = total + distance == total;   // see march_test (core.cuh): the reference would spin for ever here
            total += distance;
            if (distance < DIST_EPSILON || total > hi || !(total < hi) || stuck) {
                active = false;
                // World::intersect only takes a hit with `distance < closest`, and closest starts at +inf (world.rs:273-299): a march that
                // runs off to infinity inside an unbounded exit distance (a direction component of exactly 0) "hits" at +inf and is dropped
                if (total <= hi && total > DIST_EPSILON && total < PYR_INF) {
                    if (mode == 0) atomicMin(a.march_key + at, pack_hit(total, KIND_RAY_MARCHED, mr.rank));
                    else if (occludes(mode, total, limit)) a.shadow_kinds[at - a.shadow_offset] = KIND_RAY_MARCHED;
                }
            }
        }
    }
    if (STATS) {
        for (int d = 16; d; d >>= 1) {
            evals += __shfl_down_sync(FULL, evals, d); iters += __shfl_down_sync(FULL, iters, d); julia_iters += __shfl_down_sync(FULL, julia_iters, d);
        }
        if (lane_id() == 0) { atomicAdd(&a.counters->de_evals, evals); atomicAdd(&a.counters->de_iterations, iters); atomicAdd(&a.counters->march_iterations, iters);
                              atomicAdd(&a.counters->julia_iterations, julia_iters); }
    }
}

// Write the merged result back into the hit record of every path ray a marched shape won.
__global__ void __launch_bounds__(256) k_march_apply(const TraceArgs a, const MarchedRec* marched) {
    for (int type = 0; type < 2; ++type) {
        const uint32_t n = min(a.march_count[type], a.march_capacity[type]);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const uint2 item = a.march_queue[type][i];
            if (item.x >= a.shadow_offset) continue;
            const unsigned long long key = a.march_key[item.x];
            const uint32_t tie = (uint32_t)key;
            if (key == ~0ull || !(tie & 0x80000000u) || (tie & 0x7fffffffu) != marched[item.y].rank) continue;  // this shape did not win
            float4* dst = reinterpret_cast<float4*>(a.hits + item.x);
            dst[0] = make_float4(__uint_as_float((uint32_t)(key >> 32)), 0.0f, 0.0f, __uint_as_float(tie & 0x7fffffffu));
            dst[1] = make_float4(__uint_as_float((uint32_t)KIND_RAY_MARCHED), 0.0f, 0.0f, 0.0f);
        }
    }
}

// pyr_trace seam: pyr_ray (32 B) in, pyr_hit (20 B) out; always closest-hit mode
struct AbiHit { uint32_t prim_id, kind; float t, u, v; };
struct EmitAbiHit {
    AbiHit* hits;
    const Prim* prims;
    const SceneView* scene;
    template <class T>
    __device__ __forceinline__ void operator()(uint32_t at, const T& tr_in) const {
        T tr = tr_in;
        if (tr.march_mask) tr.resolve_marched(*scene);  // the parity seam resolves sphere tracing in place
        AbiHit o;
        o.kind = tr.kind; o.t = tr.t; o.u = tr.u; o.v = tr.v;
        if (tr.kind == KIND_MISS) o.prim_id = 0xFFFFFFFFu;
        else if (tr.kind == KIND_PLANE) o.prim_id = tr.rank;
        else o.prim_id = prim_object(prims[tr.rank]);
        hits[at] = o;
    }
};
template <bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_batch(const __grid_constant__ SceneView sc, const Ray* rays, uint32_t n, AbiHit* hits, uint32_t* cursor,
                                                               DeviceCounters* counters, uint32_t refill_min, uint32_t steps) {
    EmitAbiHit emit{hits, sc.prims, &sc};
    trace_persistent<STATS>(sc, rays, n, 0u, 0u, cursor, counters, true, refill_min, steps, emit);
}

// ------------------------------------------------------------------------------------------
__global__ void k_film_expose(const SceneView sc, float* film, const float* positions, const float* samples, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FilmAdd add{film};
    film_expose(sc.film, positions[2 * i], positions[2 * i + 1], samples[3 * i], samples[3 * i + 1], samples[3 * i + 2], add);
}

__global__ void k_white_scan(const SceneView sc, float* params) {
    float white_max, d65_max;
    white_scan(sc, white_max, d65_max);
    params[0] = white_max;
    params[1] = d65_max;
}

__global__ void k_develop(const SceneView sc, const float* film, const float* params, float step_size, float* xyz_out, uint8_t* srgb_out) {
    const uint64_t pixels = (uint64_t)sc.film.width * sc.film.height;
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pixels) return;
    float xyz[3] = {0.0f, 0.0f, 0.0f};
    uint8_t rgb[3] = {0, 0, 0};
    // DevelopedPixels::next stops at `end < len` (film.rs:299): the last pixel stays black
    if (p + 1 < pixels) {
        DevelopParams dp;
        dp.white_max = params[0]; dp.d65_max = params[1]; dp.step_size = step_size; dp.pad = 0;
        pixel_to_xyz(sc, dp, film, p, xyz);
        xyz_to_srgb8(xyz, rgb);
    }
    if (xyz_out) { xyz_out[3 * p] = xyz[0]; xyz_out[3 * p + 1] = xyz[1]; xyz_out[3 * p + 2] = xyz[2]; }
    if (srgb_out) { srgb_out[3 * p] = rgb[0]; srgb_out[3 * p + 1] = rgb[1]; srgb_out[3 * p + 2] = rgb[2]; }
}

// out: position[2], ray o[3] pad d[3] pad (8 floats), wavelengths[16], hero index (as float bits)
__global__ void k_camera_sample(const SceneView sc, uint64_t seed, uint32_t tile, uint64_t sample, float* out) {
    const TileRec t = sc.tiles[tile];
    Rng rng = keyed_rng(seed, t.index, sample);
    float ox = t.size[0] * rng.gen_f32();
    float oy = t.size[1] * rng.gen_f32();
    float px = t.from[0] + ox, py = t.from[1] + oy;
    float wl[MAX_SPECTRUM_SAMPLES];
    v3 o, d;
    uint32_t pick;
    if (sc.renderer.algorithm == 0) {  // simple.rs:87-107
        camera_ray(sc.camera, px, py, rng, o, d);
        pick = sample_wavelengths(sc, rng, wl);
    } else {                            // bidirectional.rs:105-124
        pick = sample_wavelengths(sc, rng, wl);
        camera_ray(sc.camera, px, py, rng, o, d);
    }
    out[0] = px; out[1] = py;
    out[2] = o.x; out[3] = o.y; out[4] = o.z; out[5] = 0.0f; out[6] = d.x; out[7] = d.y; out[8] = d.z; out[9] = 0.0f;
    for (uint32_t k = 0; k < sc.renderer.spectrum_samples; ++k) out[10 + k] = wl[k];
    out[10 + MAX_SPECTRUM_SAMPLES] = __uint_as_float(pick);
}

// tools/first_divergence.py: ONE path sample of the camera-to-light integrator run depth-first by a single thread
// (debug_path_simple in shading.cuh).  A diagnostic seam; nothing on the render path calls it.
__global__ void k_debug_path(const __grid_constant__ SceneView sc, uint64_t seed, uint32_t tile, uint64_t sample, uint32_t max_bounces, uint32_t* records,
                             uint32_t* counts, float* exposed, float* position2, Ray* rays, Hit* hits, PendingLight* pend, uint32_t* kinds) {
    // rays / hits / pend / kinds: global scratch (the stage functions move these records with ld/st.global)
    PathState ps;
    bind_spectral(sc, ps);
    ps.pend = pend;
    ps.bd = nullptr;
    ShadeOut out;
    out.stage_base = sc.vm_regs * WAVE_THREADS;
    debug_path_simple(sc, seed, tile, sample, max_bounces, records, counts, exposed, position2, rays, hits, kinds, ps, out);
}

}  // namespace

#include "bdpt_kernels.inl"

void launch_pool_reset(PathCore* paths, uint32_t pool, uint32_t* live_list, uint32_t* live_count, cudaStream_t s) {
    if (pool) k_pool_reset<<<(pool + 255) / 256, 256, 0, s>>>(paths, pool, live_list, live_count);
}
void launch_bin(const WaveArgs& a, const BinBuffers& b, uint32_t cluster_shift, int bidirectional, cudaStream_t s) {
    const unsigned blocks = (a.grid_paths + BIN_THREADS * BIN_ITEMS - 1) / (BIN_THREADS * BIN_ITEMS);
    cudaFuncSetAttribute(k_bin_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NUM_KEYS * sizeof(uint32_t)));
    k_bin_keys<<<blocks, BIN_THREADS, NUM_KEYS * sizeof(uint32_t), s>>>(a.paths, bidirectional ? a.bidir : nullptr, a.pool, a.hits_in, a.shadow_kinds_in, cluster_shift, b.count, b.keys,
                                              a.live_list, a.live_count_in);
    k_bin_scan<<<1, 1024, 0, s>>>(b.count, b.first, b.fill, a.count_out, a.live_count_out, a.died_count);
    cudaFuncSetAttribute(k_bin_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * NUM_KEYS * sizeof(uint32_t)));
    k_bin_scatter<<<blocks, BIN_THREADS, 2 * NUM_KEYS * sizeof(uint32_t), s>>>(a.pool, b.keys, b.first, b.fill, b.list, a.live_list, a.live_count_in);
}
void launch_wave_simple(const SceneView& sc, const WaveArgs& a, cudaStream_t s) {
    cudaFuncSetAttribute(k_wave_simple, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem(sc));
    k_wave_simple<<<(a.grid_paths + WAVE_THREADS - 1) / WAVE_THREADS, WAVE_THREADS, wave_smem(sc), s>>>(sc, a);
}
TraceTuning trace_tuning() {
    static TraceTuning t = [] {
        TraceTuning v{8u, 2u};
        if (const char* e = getenv("PYR_TRACE_REFILL")) v.refill_min = (uint32_t)atoi(e);
        if (const char* e = getenv("PYR_TRACE_STEPS")) v.steps = (uint32_t)atoi(e);
        if (v.refill_min < 1) v.refill_min = 1;
        if (v.refill_min > 32) v.refill_min = 32;
        if (v.steps < 1) v.steps = 1;
        return v;
    }();
    return t;
}
void launch_trace(const SceneView& sc, const TraceArgs& a_in, int grid_blocks, cudaStream_t s) {
    TraceArgs a = a_in;
    const TraceTuning t = trace_tuning();
    a.refill_min = t.refill_min; a.steps = t.steps;
    if (a.stats) k_trace<true><<<grid_blocks, TRACE_THREADS, 0, s>>>(sc, a);
    else k_trace<false><<<grid_blocks, TRACE_THREADS, 0, s>>>(sc, a);
}
void launch_march(const SceneView& sc, const TraceArgs& a_in, int grid_blocks, cudaStream_t s) {
    TraceArgs a = a_in;
    uint32_t* cursors = a.march_count + 2;  // two work cursors behind the two queue counters (zeroed with them)
    if (a.stats) k_march<true><<<grid_blocks, TRACE_THREADS, 0, s>>>(sc, a, cursors);
    else k_march<false><<<grid_blocks, TRACE_THREADS, 0, s>>>(sc, a, cursors);
    k_march_apply<<<grid_blocks, 256, 0, s>>>(a, sc.marched);
}
void launch_trace_batch(const SceneView& sc, const void* rays32, size_t n, void* hits20, uint32_t* cursor, DeviceCounters* counters, int stats,
                        int grid_blocks, cudaStream_t s) {
    const TraceTuning t = trace_tuning();
    if (stats) k_trace_batch<true><<<grid_blocks, TRACE_THREADS, 0, s>>>(sc, (const Ray*)rays32, (uint32_t)n, (AbiHit*)hits20, cursor, counters, t.refill_min, t.steps);
    else k_trace_batch<false><<<grid_blocks, TRACE_THREADS, 0, s>>>(sc, (const Ray*)rays32, (uint32_t)n, (AbiHit*)hits20, cursor, counters, t.refill_min, t.steps);
}
void launch_film_expose(const SceneView& sc, float* film, const float* positions, const float* samples, size_t n, cudaStream_t s) {
    if (n) k_film_expose<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc, film, positions, samples, n);
}
void launch_white_scan(const SceneView& sc, float* develop_params, cudaStream_t s) { k_white_scan<<<1, 1, vm_smem(sc), s>>>(sc, develop_params); }
void launch_develop(const SceneView& sc, const float* film, const float* develop_params, float step_size, float* xyz, uint8_t* srgb, cudaStream_t s) {
    const uint64_t pixels = (uint64_t)sc.film.width * sc.film.height;
    k_develop<<<(unsigned)((pixels + 127) / 128), 128, vm_smem(sc), s>>>(sc, film, develop_params, step_size, xyz, srgb);
}
void launch_camera_sample(const SceneView& sc, uint64_t seed, uint32_t tile, uint64_t sample, float* out, cudaStream_t s) {
    k_camera_sample<<<1, 1, 0, s>>>(sc, seed, tile, sample, out);
}

void launch_debug_path(const SceneView& sc, uint64_t seed, uint32_t tile, uint64_t sample, uint32_t max_bounces, uint32_t* records, uint32_t* counts,
                       float* exposed, float* position2, void* scratch, cudaStream_t s) {
    // scratch: (1 + MAX_LIGHT_SAMPLES) rays, as many hits, MAX_LIGHT_SAMPLES pending lights, 1 + MAX_LIGHT_SAMPLES kinds (32-byte aligned)
    Ray* rays = (Ray*)scratch;
    Hit* hits = (Hit*)(rays + 1 + MAX_LIGHT_SAMPLES);
    PendingLight* pend = (PendingLight*)(hits + 1 + MAX_LIGHT_SAMPLES);
    uint32_t* kinds = (uint32_t*)(pend + MAX_LIGHT_SAMPLES);
    cudaFuncSetAttribute(k_debug_path, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem(sc));
    k_debug_path<<<1, 1, wave_smem(sc), s>>>(sc, seed, tile, sample, max_bounces, records, counts, exposed, position2, rays, hits, pend, kinds);
}
size_t debug_path_scratch_bytes() { return (size_t)(2 * (1 + MAX_LIGHT_SAMPLES) + MAX_LIGHT_SAMPLES) * 32 + (1 + MAX_LIGHT_SAMPLES) * sizeof(uint32_t); }

size_t path_state_bytes() { return sizeof(PathCore); }
size_t pending_light_bytes() { return sizeof(PendingLight); }
size_t bidir_state_bytes() { return sizeof(BidirState); }
size_t light_vertex_bytes() { return sizeof(LightVertex); }
int trace_blocks_per_sm() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<false>, TRACE_THREADS, 0);
    return n > 0 ? n : 1;
}

}  // namespace pyr
