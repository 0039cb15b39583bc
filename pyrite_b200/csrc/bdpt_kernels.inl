// included by kernels.cu inside namespace pyr: the bidirectional wavefront kernels.
//
// One specialised kernel per phase of a path sample (bdpt.cuh) instead of one kernel holding the union of all phases'
// registers: the slots are already sorted by phase (k_bin_*), so every kernel walks ONE contiguous run of the sorted list
// with a grid-stride loop (grids are sized by the SM count, not by the pool), each compiled with the registers its phase
// needs.  Launch order per wavefront iteration:  LAMP, FINISH, CAMERA, CONNECT, SPLAT (slots that end a sample are appended to the
// `died` list), then GEN, which starts new samples in the dead slots of the sorted list and in the slots that just died -
// so a slot is never idle for an iteration.  All of them append to the same ray queue / live list.
namespace {

#ifndef BD_CONNECT_BLOCKS
#define BD_CONNECT_BLOCKS 4
#endif
enum : int { BD_GEN = 0, BD_LAMP = 1, BD_CAMERA = 2, BD_CONNECT = 3, BD_SPLAT = 4, BD_FINISH = 5 };

// Queue space for one block: path rays, visibility rays, the live list and the list of slots whose sample just ended.
struct ReservationBd { uint32_t main_at, shadow_at, live_at, died_at; };
__device__ __forceinline__ ReservationBd block_reserve_bd(const WaveArgs& a, uint32_t n_main, uint32_t n_shadow, bool alive, bool died) {
    constexpr int WARPS = WAVE_THREADS / 32;
    __shared__ uint32_t s_total[4][WARPS];
    __shared__ uint32_t s_base[4];
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t total_main, total_shadow;
    ReservationBd r;
    r.main_at = warp_exclusive_scan(n_main, total_main);
    r.shadow_at = warp_exclusive_scan(n_shadow, total_shadow);
    const unsigned live_mask = __ballot_sync(FULL, alive), died_mask = __ballot_sync(FULL, died);
    const unsigned below = (1u << lane_id()) - 1u;
    r.live_at = __popc(live_mask & below);
    r.died_at = __popc(died_mask & below);
    __syncthreads();  // the previous trip of the grid-stride loop has finished reading s_total / s_base
    if (lane_id() == 0) { s_total[0][warp] = total_main; s_total[1][warp] = total_shadow; s_total[2][warp] = __popc(live_mask); s_total[3][warp] = __popc(died_mask); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t[4] = {0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { t[0] += s_total[0][w]; t[1] += s_total[1][w]; t[2] += s_total[2][w]; t[3] += s_total[3][w]; }
        unsigned long long both = 0;
        uint32_t live = 0, died_base = 0;
        if (t[0] | t[1]) both = atomicAdd(reinterpret_cast<unsigned long long*>(a.count_out), (unsigned long long)t[0] | ((unsigned long long)t[1] << 32));
        if (t[2]) live = atomicAdd(a.live_count_out, t[2]);
        if (t[3]) died_base = atomicAdd(a.died_count, t[3]);
        s_base[0] = (uint32_t)both; s_base[1] = (uint32_t)(both >> 32); s_base[2] = live; s_base[3] = died_base;
    }
    __syncthreads();
    uint32_t before[4] = {s_base[0], s_base[1], s_base[2], s_base[3]};
#pragma unroll
    for (int w = 0; w < WARPS; ++w)
        if ((uint32_t)w < warp) { before[0] += s_total[0][w]; before[1] += s_total[1][w]; before[2] += s_total[2][w]; before[3] += s_total[3][w]; }
    r.main_at += before[0]; r.shadow_at += before[1]; r.live_at += before[2]; r.died_at += before[3];
    return r;
}

__device__ __forceinline__ void load_bidir(BidirState& bd, const BidirState* src_) {
    static_assert(sizeof(BidirState) == 96, "three chunks");
    const Vec8* src = reinterpret_cast<const Vec8*>(src_);
    const Vec8 q0 = ld256_stream(src), q1 = ld256_stream(src + 1), q2 = ld256_stream(src + 2);
    __builtin_memcpy(&bd, &q0, 32);
    __builtin_memcpy(reinterpret_cast<char*>(&bd) + 32, &q1, 32);
    __builtin_memcpy(reinterpret_cast<char*>(&bd) + 64, &q2, 32);
}
__device__ __forceinline__ void store_bidir(BidirState* dst_, const BidirState& bd) {
    Vec8 q[3];
    __builtin_memcpy(q, &bd, 96);
    Vec8* dst = reinterpret_cast<Vec8*>(dst_);
    st256_stream(dst, q[0]); st256_stream(dst + 1, q[1]); st256_stream(dst + 2, q[2]);
}

// first / one-past-last position of a phase's run in the sorted list (keys are state * BIN_CLUSTERS + cluster, see bin_key)
template <int PHASE> __device__ __forceinline__ void phase_run(const WaveArgs& a, uint32_t& lo, uint32_t& hi) {
    constexpr uint32_t S0 = PHASE == BD_GEN ? 0u : PHASE == BD_CAMERA ? 1u : PHASE == BD_LAMP ? 26u : PHASE == BD_FINISH ? 31u : PHASE == BD_CONNECT ? 32u : 40u;
    constexpr uint32_t S1 = PHASE == BD_GEN ? 1u : PHASE == BD_CAMERA ? 26u : PHASE == BD_LAMP ? 31u : PHASE == BD_FINISH ? 32u : PHASE == BD_CONNECT ? 40u : BIN_STATES;
    lo = __ldg(a.bin_first + S0 * BIN_CLUSTERS);
    hi = __ldg(a.bin_first + S1 * BIN_CLUSTERS);
}

// resident blocks per SM the compiler plans for: the camera step needs its 168 registers (3 blocks, like k_wave_simple),
// the other phases fit 128 (4 blocks)
template <int PHASE>
__global__ void __launch_bounds__(WAVE_THREADS, PHASE == BD_CAMERA ? 3 : (PHASE == BD_CONNECT || PHASE == BD_SPLAT) ? BD_CONNECT_BLOCKS : 4) k_wave_bd(const __grid_constant__ SceneView sc, const __grid_constant__ WaveArgs a) {
    if (PHASE == BD_LAMP && blockIdx.x == 0 && threadIdx.x == 0) *a.trace_cursor = 0;  // the first kernel after k_bin_*
    uint32_t lo, hi;
    phase_run<PHASE>(a, lo, hi);
    // GEN also takes the slots whose sample ended in this iteration's phase kernels
    const uint32_t n_sorted = hi - lo, n = PHASE == BD_GEN ? n_sorted + *a.died_count : n_sorted;
    FilmAdd add{a.film};
    PathCounters pc;
    pc.de_evals = 0; pc.de_iters = 0;
    for (uint32_t base = blockIdx.x * WAVE_THREADS; base < n; base += gridDim.x * WAVE_THREADS) {
        const uint32_t g = base + threadIdx.x;
        const bool valid = g < n;
        uint32_t slot = 0;
        if (valid) slot = (PHASE != BD_GEN || g < n_sorted) ? a.bin_list[lo + g] : a.died_list[g - n_sorted];
        PathState ps;
        bind_spectral<PHASE == BD_CONNECT || PHASE == BD_SPLAT>(sc, ps);
        ps.flags = 0;
        BidirState bd;
        ps.bd = &bd;
        ps.pend = a.pend + (size_t)slot * MAX_LIGHT_SAMPLES;
        BidirCtx cx;
        cx.lv = a.light_vertices + (size_t)slot * a.light_stride;
        cx.cv = a.cam_vertices + (size_t)slot * a.cam_stride;
        cx.bright.base = ps.refl.base + sc.renderer.spectrum_samples * WAVE_THREADS;  // two more [S][thread] arrays behind wl | bright | refl
        cx.refl.base = cx.bright.base + sc.renderer.spectrum_samples * WAVE_THREADS;
        BidirOut out;
        out.alive = 0; out.has_main = 0; out.n_shadow = 0; out.shadow_kind = SH_NONE;
        ShadeOut so;
        so.alive = 0; so.has_main = 0; so.n_shadow = 0;
        so.stage_base = sc.vm_regs * WAVE_THREADS;
        bool alive = false, touched = false;
        if (PHASE == BD_GEN) {
            unsigned long long gs = 0;
            if (claim_sample(a.next_sample, a.total_samples, valid, gs)) {
                uint32_t tile; unsigned long long k;
                locate_sample(a.tile_first, sc.n_tiles, gs, tile, k);
                for (uint32_t w = 0; w < sizeof(BidirState) / 4; ++w) reinterpret_cast<uint32_t*>(&bd)[w] = 0;
                generate_bidirectional(sc, a.seed, tile, (uint64_t)a.sample_offset + k * a.sample_stride, ps, cx, out);
                alive = true; touched = true;
            }
        } else if (valid) {
            load_core(sc, ps, a.paths + slot);
            load_bidir(bd, a.bidir + slot);
            touched = true;
            const Ray* main_ray = a.rays_in + ps.ray_base;
            const Hit* main_hit = a.hits_in + ps.ray_base;
            const uint32_t* kinds = a.shadow_kinds_in + ps.shadow_base;
            if (PHASE == BD_LAMP) shade_bd_lamp(sc, ps, cx, main_ray, main_hit, out, pc);
            else if (PHASE == BD_FINISH) shade_bd_finish(sc, ps, cx, out);
            else if (PHASE == BD_CAMERA) shade_bd_camera(sc, ps, cx, main_ray, main_hit, a.rays_in + a.shadow_offset + ps.shadow_base, kinds, so, out, add, pc);
            else if (PHASE == BD_CONNECT) shade_bd_connect(sc, ps, cx, kinds, out, add);
            else shade_bd_splat(sc, ps, cx, kinds, out, add);
            alive = out.alive != 0;
        }
        if (alive) ps.flags |= PS_ALIVE; else ps.flags = 0;
        if (alive && bd.phase >= PH_CONNECT) ps.n_pending = out.n_shadow;  // k_bin counts the unblocked ones
        // path rays and visibility rays go to separate queue regions so that trace packets are homogeneous
        const uint32_t n_main = alive ? out.has_main : 0u, n_shadow = alive ? out.n_shadow : 0u;
        const ReservationBd at = block_reserve_bd(a, n_main, n_shadow, alive, PHASE != BD_GEN && valid && !alive);
        if (n_main) { ps.ray_base = at.main_at; store_ray(a.rays_out + at.main_at, out.main); }
        if (n_shadow) {
            ps.shadow_base = at.shadow_at;
            Ray* dst = a.rays_out + a.shadow_offset + at.shadow_at;
            if (PHASE == BD_CAMERA && out.shadow_kind == SH_NEE) { for (uint32_t j = 0; j < n_shadow; ++j) store_ray(dst + j, so.get_shadow(j)); }
            else write_staged(sc, ps, cx, out.shadow_kind, dst);
        }
        if (touched) { store_core(sc, a.paths + slot, ps); store_bidir(a.bidir + slot, bd); }
        if (alive) a.live_list[at.live_at] = slot;
        else if (PHASE != BD_GEN && valid) a.died_list[at.died_at] = slot;
    }
    if (pc.de_evals) { atomicAdd(&a.counters->de_evals, (unsigned long long)pc.de_evals); atomicAdd(&a.counters->de_iterations, (unsigned long long)pc.de_iters); }
}

template <int PHASE>
void launch_bd_phase(const SceneView& sc, const WaveArgs& a, size_t smem, int sm_count, cudaStream_t s) {
    static int per_sm = 0;  // resident blocks per SM of this phase's kernel (registers / shared memory decide)
    static size_t per_sm_smem = ~(size_t)0;
    cudaFuncSetAttribute(k_wave_bd<PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (per_sm == 0 || per_sm_smem != smem) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wave_bd<PHASE>, WAVE_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        per_sm_smem = smem;
    }
    const unsigned needed = (a.grid_paths + WAVE_THREADS - 1) / WAVE_THREADS;
    const unsigned grid = std::min<unsigned>(needed, (unsigned)(sm_count * per_sm));
    k_wave_bd<PHASE><<<grid ? grid : 1u, WAVE_THREADS, smem, s>>>(sc, a);
}

}  // namespace

inline size_t bidir_smem(const SceneView& sc) { return wave_smem(sc) + (size_t)2 * sc.renderer.spectrum_samples * WAVE_THREADS * sizeof(float); }
// connect / splat: wl | bright | refl | detached bright | detached refl, nothing else (bind_spectral<LEAN>)
inline size_t bidir_lean_smem(const SceneView& sc) { return (size_t)5 * sc.renderer.spectrum_samples * WAVE_THREADS * sizeof(float); }
void launch_wave_bidirectional(const SceneView& sc, const WaveArgs& a, int sm_count, cudaStream_t s) {
    const size_t smem = bidir_smem(sc), lean = bidir_lean_smem(sc);
    launch_bd_phase<BD_LAMP>(sc, a, smem, sm_count, s);
    launch_bd_phase<BD_FINISH>(sc, a, smem, sm_count, s);
    launch_bd_phase<BD_CAMERA>(sc, a, smem, sm_count, s);
    launch_bd_phase<BD_CONNECT>(sc, a, lean, sm_count, s);
    launch_bd_phase<BD_SPLAT>(sc, a, lean, sm_count, s);
    launch_bd_phase<BD_GEN>(sc, a, smem, sm_count, s);
}
int wave_bidirectional_launches() { return 6; }
size_t cam_vertex_bytes() { return sizeof(CamVertex); }
int bdpt_stage_rays() { return BDPT_STAGE; }
