// included by kernels.cu inside namespace pyr: the bidirectional wavefront kernel
namespace {

__global__ void __launch_bounds__(WAVE_THREADS) k_wave_bidirectional(const __grid_constant__ SceneView sc, const __grid_constant__ WaveArgs a) {
    const uint32_t g_thread = blockIdx.x * blockDim.x + threadIdx.x;
    if (g_thread == 0) *a.trace_cursor = 0;
    bool valid, dead;
    // blocks beyond the sorted list have nothing to do (the grid is sized by a bound of the live count that can be stale)
    if (blockIdx.x * blockDim.x >= __ldg(a.bin_first + NUM_KEYS)) return;
    const uint32_t slot = binned_slot(a, g_thread, valid, dead);
    const bool idle = __all_sync(FULL, !valid || dead) && *a.next_sample >= a.total_samples;
    const uint32_t s = valid ? slot : 0;
    PathState ps;
    bind_spectral(sc, ps);
    ps.flags = 0;
    BidirState bd;
    ps.bd = &bd;
    uint32_t flags_in = 0;
    BidirOut out;
    out.alive = 0; out.has_main = 0; out.n_shadow = 0;
    PathCounters pc;
    pc.de_evals = 0; pc.de_iters = 0;
    bool alive = false;
    if (!idle) {
        load_core(sc, ps, a.paths + s);
        ps.pend = a.pend + (size_t)s * MAX_LIGHT_SAMPLES;
        {
            static_assert(sizeof(BidirState) == 96, "three chunks");
            const Vec8* src = reinterpret_cast<const Vec8*>(a.bidir + s);
            const Vec8 q0 = ld256_stream(src), q1 = ld256_stream(src + 1), q2 = ld256_stream(src + 2);
            __builtin_memcpy(&bd, &q0, 32);
            __builtin_memcpy(reinterpret_cast<char*>(&bd) + 32, &q1, 32);
            __builtin_memcpy(reinterpret_cast<char*>(&bd) + 64, &q2, 32);
        }
        flags_in = ps.flags;
        BidirCtx cx;
        cx.lv = a.light_vertices + (size_t)s * a.light_stride;
        cx.cv = a.cam_vertices + (size_t)s * a.cam_stride;
        cx.bright.base = ps.refl.base + sc.renderer.spectrum_samples * WAVE_THREADS;  // two more [S][thread] arrays behind wl | bright | refl
        cx.refl.base = cx.bright.base + sc.renderer.spectrum_samples * WAVE_THREADS;
        FilmAdd add{a.film};
        alive = valid && (ps.flags & PS_ALIVE);
        if (alive) {
            shade_bidirectional(sc, ps, cx, a.rays_in + ps.ray_base, a.hits_in + ps.ray_base, a.rays_in + a.shadow_offset + ps.shadow_base,
                                a.shadow_kinds_in + ps.shadow_base, out, add, pc);
            alive = out.alive != 0;
            if (alive) ps.flags |= PS_ALIVE; else ps.flags = 0;
            if (alive && bd.phase >= PH_CONNECT) ps.n_pending = out.n_shadow;  // k_bin counts the unblocked ones
        }
        unsigned long long g = 0;
        if (claim_sample(a.next_sample, a.total_samples, valid && !alive, g)) {
            uint32_t tile; unsigned long long k;
            locate_sample(a.tile_first, sc.n_tiles, g, tile, k);
            generate_bidirectional(sc, a.seed, tile, (uint64_t)a.sample_offset + k * a.sample_stride, ps, cx, out);
            ps.flags |= PS_ALIVE;
            alive = true;
        }
    }
    const uint32_t n_main = alive ? out.has_main : 0u, n_shadow = alive ? out.n_shadow : 0u;
    const Reservation at = block_reserve(a, n_main, n_shadow, valid && alive);
    if (n_main) { ps.ray_base = at.main_at; store_ray(a.rays_out + at.main_at, out.main); }
    if (n_shadow) {
        ps.shadow_base = at.shadow_at;
        for (uint32_t j = 0; j < n_shadow; ++j) store_ray(a.rays_out + a.shadow_offset + at.shadow_at + j, out.shadow[j]);
    }
    if (valid && (alive || (flags_in & PS_ALIVE))) {
        store_core(sc, a.paths + slot, ps);
        Vec8 q[3];
        __builtin_memcpy(q, &bd, 96);
        Vec8* dst = reinterpret_cast<Vec8*>(a.bidir + slot);
        st256_stream(dst, q[0]); st256_stream(dst + 1, q[1]); st256_stream(dst + 2, q[2]);
    }
    if (valid && alive) a.live_list[at.live_at] = slot;
    if (pc.de_evals) { atomicAdd(&a.counters->de_evals, (unsigned long long)pc.de_evals); atomicAdd(&a.counters->de_iterations, (unsigned long long)pc.de_iters); }
}

}  // namespace

inline size_t bidir_smem(const SceneView& sc) { return wave_smem(sc) + (size_t)2 * sc.renderer.spectrum_samples * WAVE_THREADS * sizeof(float); }
void launch_wave_bidirectional(const SceneView& sc, const WaveArgs& a, cudaStream_t s) {
    cudaFuncSetAttribute(k_wave_bidirectional, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bidir_smem(sc));
    k_wave_bidirectional<<<(a.grid_paths + WAVE_THREADS - 1) / WAVE_THREADS, WAVE_THREADS, bidir_smem(sc), s>>>(sc, a);
}
size_t cam_vertex_bytes() { return sizeof(CamVertex); }
int bdpt_stage_rays() { return BDPT_STAGE; }
