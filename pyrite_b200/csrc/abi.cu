// C ABI of the B200 render path (include/pyrite_b200.h): context management, scene upload, the
// wavefront driver loop and the film / trace / camera seams.  No CPU fallback exists: every entry
// point that computes launches kernels on the context's CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pyrite_b200.h"
#include "kernels.hpp"
#include "scene_build.hpp"

namespace {

using namespace pyr;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

std::string g_init_error;
struct StateError : std::runtime_error { using std::runtime_error::runtime_error; };

// NCCL is bound at the first pyr_comm_* call, not at load time: a single-GPU host needs no NCCL at all, and a host that
// already carries one (PyTorch bundles its own libnccl.so.2) must not get a second copy - dlopen by soname returns the
// one that is loaded.  Only the stable core API is used (present in every NCCL 2.x).
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.Reduce = (decltype(a.Reduce))dlsym(a.handle, "ncclReduce");
        a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        a.GetVersion = (decltype(a.GetVersion))dlsym(a.handle, "ncclGetVersion");
        return a;
    }();
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Reduce || !api.AllReduce || !api.GetErrorString)
        throw StateError("NCCL is not available (libnccl.so.2 could not be loaded): the multi-GPU film reduction needs it");
    return api;
}
#define NC(call)                                                                                              \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess) throw CudaError(std::string(#call) + ": " + nccl().GetErrorString(r_));     \
    } while (0)

struct DeviceBuffer {
    void* p = nullptr;
    size_t bytes = 0;
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    void ensure(size_t n) {
        if (n <= bytes && p) return;
        release();
        if (n == 0) n = 16;
        CU(cudaMalloc(&p, n));
        bytes = n;
    }
    template <class T> T* as() const { return (T*)p; }
};

template <class T, class A> void upload(DeviceBuffer& b, const std::vector<T, A>& v, cudaStream_t s) {
    b.ensure(v.size() * sizeof(T));
    if (!v.empty()) CU(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

}  // namespace

// One wavefront: a pool of path slots with its queues, sort buffers and counters.  A render drives one lane on the context's
// stream, or two lanes of half the pool each on two streams of their own: the shade kernels wait on memory latency with few
// warps per SM and leave most issue slots idle, the traversal kernel is issue-bound - two independent wavefronts let the GPU
// run the one lane's traversal under the other lane's shading.  Path samples are handed out from ONE counter and every
// sample owns its RNG stream, so the film does not depend on which lane renders a sample.
struct Lane {
    cudaStream_t own_stream = nullptr;  // two-lane renders
    cudaEvent_t batch_done = nullptr;
    unsigned long long* pinned = nullptr;  // [0] ray counts, [1] next sample, [2] live-slot count
    DeviceBuffer paths, pend, bidir, rays[2], hits, shadow_kinds, march_queue[2], march_key, light_vertices, cam_vertices, bin_count, bin_first, bin_fill,
        bin_keys, bin_list, live_list, died_list, scalars;
    uint32_t pool = 0;
    uint32_t march_capacity[2] = {0, 0};
    // state of the running render
    cudaStream_t stream = nullptr;
    int cur = 0;
    uint32_t grid_paths = 0;
    bool finished = false;
    // scalars: [0..1] counts A {path rays, visibility rays}, [2..3] counts B, [4] trace cursor, [8..9] march counts, [10..11] march cursors,
    // [12..13] live-slot counts, [14] died-slot count
    uint32_t* count(int i) const { return scalars.as<uint32_t>() + 2 * i; }
    uint32_t* cursor() const { return scalars.as<uint32_t>() + 4; }
    uint32_t* march_count() const { return scalars.as<uint32_t>() + 8; }
    uint32_t* live_count(int i) const { return scalars.as<uint32_t>() + 12 + i; }
    uint32_t* died_count() const { return scalars.as<uint32_t>() + 14; }
    void release() {
        DeviceBuffer* all[] = {&paths, &pend, &bidir, &rays[0], &rays[1], &hits, &shadow_kinds, &march_queue[0], &march_queue[1], &march_key, &light_vertices,
                               &cam_vertices, &bin_count, &bin_first, &bin_fill, &bin_keys, &bin_list, &live_list, &died_list, &scalars};
        for (DeviceBuffer* b : all) b->release();
        pool = 0;
    }
};
constexpr int MAX_LANES = 2;

struct pyr_ctx {
    int device = 0;
    int sm_count = 0;
    size_t device_memory = 0;
    cudaStream_t stream = nullptr;      // the stream work is launched on
    cudaStream_t own_stream = nullptr;  // created by pyr_init
    std::vector<cudaEvent_t> timing_events;
    std::string error;
    bool loaded = false;
    bool bvh_built_on_gpu = false;
    double load_seconds = 0.0;  // IR decode + scene build + upload of the last pyr_project_load
    BakedScene scene;
    SceneView view{};
    DeviceBuffer nodes, prims, tri_shade, tri_frames, planes, marched, materials, components, programs, code, spectra, spectrum_data,
        textures, texels, lamps, tiles, burns, xyz, d65;
    DeviceBuffer film, develop_params, counters, scalars, tile_first;
    Lane lanes[MAX_LANES];
    cudaEvent_t ev_fork = nullptr;
    uint32_t shadow_per_path = 1;
    DeviceBuffer scratch_a, scratch_b, bvh_scratch;
    bool develop_params_valid = false;
    pyr_counters host_counters{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ncclComm_t comm = nullptr;   // pyr_comm_init / pyr_comm_init_async
    int comm_ranks = 0, comm_rank = 0;
    // The communicator is set up on a thread of its own, on a stream of its own, so that a render can run meanwhile
    // (ncclCommInitRank + NCCL's connection set-up at the first collective take seconds on an 8-GPU box).  Everything that
    // uses `comm` joins the thread first (comm_wait).
    std::thread comm_thread;
    // The thread starts its NCCL calls when the gate opens: pyr_render opens it once its buffers are allocated and its first
    // wavefront iterations are queued (ncclCommInitRank and a render's start-up allocations block each other in the driver: on an
    // 8-GPU box the render started 1.7 s late with the set-up running from the beginning), anything that needs the communicator
    // opens it at once, and it opens by itself after 5 s.
    std::mutex comm_gate_mutex;
    std::condition_variable comm_gate_cv;
    bool comm_gate_open = true;
    std::string comm_error;       // written by the thread, read after the join
    cudaStream_t comm_stream = nullptr;
    DeviceBuffer comm_scratch;
    unsigned long long* pinned = nullptr;  // [3] sphere-tracing overflow counter

    size_t film_floats() const { return (size_t)view.film.width * view.film.height * view.film.bins * 2; }
    // shared scalars: [4] work cursor of the pyr_trace* seams, [6..7] next_sample (u64): ONE counter hands out the path samples of a render
    uint32_t* cursor() const { return scalars.as<uint32_t>() + 4; }
    unsigned long long* next_sample() const { return (unsigned long long*)(scalars.as<uint32_t>() + 6); }
};

namespace {

pyr_status fail(pyr_ctx* ctx, pyr_status code, const std::string& msg) {
    if (ctx) ctx->error = msg; else g_init_error = msg;
    return code;
}

template <class F> pyr_status guarded(pyr_ctx* ctx, F&& body) {
    if (!ctx) return fail(nullptr, PYR_ERR_INVALID, "null context");
    try {
        CU(cudaSetDevice(ctx->device));
        body();
        return PYR_OK;
    } catch (const CudaError& e) {
        return fail(ctx, PYR_ERR_CUDA, e.what());
    } catch (const StateError& e) {
        return fail(ctx, PYR_ERR_STATE, e.what());
    } catch (const ir::BuildError& e) {
        return fail(ctx, PYR_ERR_INVALID, e.what());
    } catch (const std::bad_alloc&) {
        return fail(ctx, PYR_ERR_INVALID, "out of host memory");
    } catch (const std::exception& e) {
        return fail(ctx, PYR_ERR_INVALID, e.what());
    }
}

void open_comm_gate(pyr_ctx* ctx) {
    { std::lock_guard<std::mutex> g(ctx->comm_gate_mutex); ctx->comm_gate_open = true; }
    ctx->comm_gate_cv.notify_all();
}

void need_project(pyr_ctx* ctx) {
    if (!ctx->loaded) throw StateError("no project is loaded");
}

// Argument checks shared by pyr_trace, pyr_trace_stats and pyr_trace_device.  Returns false for an empty batch.
bool check_trace_args(pyr_ctx* ctx, const void* rays, size_t n, const void* hits, bool device_buffers) {
    need_project(ctx);
    if (n > 0xFFFFFFF0ull) throw ir::BuildError("ray batch too large (at most 2^32 - 16 rays per call)");
    if (n == 0) return false;
    if (!rays || !hits) throw ir::BuildError("null ray or hit buffer");
    if (device_buffers) {
        if ((uintptr_t)rays & 31u) throw ir::BuildError("the device ray buffer must be 32-byte aligned");
        if ((uintptr_t)hits & 3u) throw ir::BuildError("the device hit buffer must be 4-byte aligned");
    }
    return true;
}

void ensure_develop_params(pyr_ctx* ctx) {
    if (ctx->develop_params_valid) return;
    ctx->develop_params.ensure(4 * sizeof(float));
    launch_white_scan(ctx->view, ctx->develop_params.as<float>(), ctx->stream);
    CU(cudaGetLastError());
    ctx->develop_params_valid = true;
}

// ray-marched shapes per distance-estimator type: each (ray, shape) pair queues at most once, so rays x shapes-of-a-type
// entries per queue can never overflow
void marched_per_type(const pyr_ctx* ctx, size_t out[2]) {
    out[0] = out[1] = 0;
    for (const MarchedRec& m : ctx->scene.marched) out[m.estimator ? 1 : 0] += 1;
}

void ensure_pool(pyr_ctx* ctx, Lane& ln, uint32_t pool) {
    if (pool == ln.pool && ln.paths.p) return;
    const RendererRec& R = ctx->view.renderer;
    const bool bidir = R.algorithm == 1;
    ctx->shadow_per_path = bidir ? (uint32_t)bdpt_stage_rays() : std::max<uint32_t>(R.light_samples, 1);
    const size_t ray_cap = (size_t)pool * (1 + ctx->shadow_per_path);
    ln.scalars.ensure(16 * sizeof(uint32_t));
    ln.paths.ensure((size_t)pool * path_state_bytes());
    ln.pend.ensure((size_t)pool * MAX_LIGHT_SAMPLES * pending_light_bytes());
    ln.bin_count.ensure(NUM_KEYS * sizeof(uint32_t));
    ln.bin_first.ensure((NUM_KEYS + 1) * sizeof(uint32_t));
    ln.bin_fill.ensure(NUM_KEYS * sizeof(uint32_t));
    ln.bin_keys.ensure((size_t)pool * sizeof(uint16_t));
    ln.bin_list.ensure((size_t)pool * sizeof(uint32_t));
    ln.live_list.ensure((size_t)pool * sizeof(uint32_t));
    ln.died_list.ensure(bidir ? (size_t)pool * sizeof(uint32_t) : 16);
    if (bidir) ln.bidir.ensure((size_t)pool * bidir_state_bytes());
    ln.rays[0].ensure(ray_cap * sizeof(Ray));
    ln.rays[1].ensure(ray_cap * sizeof(Ray));
    ln.hits.ensure((size_t)pool * sizeof(Hit));
    ln.shadow_kinds.ensure((size_t)pool * ctx->shadow_per_path * sizeof(uint32_t));
    if (ctx->view.n_marched) {
        size_t per_type[2];
        marched_per_type(ctx, per_type);
        ln.march_queue[0].ensure(ray_cap * per_type[0] * sizeof(uint2));
        ln.march_queue[1].ensure(ray_cap * per_type[1] * sizeof(uint2));
        ln.march_capacity[0] = (uint32_t)std::min<size_t>(ray_cap * per_type[0], 0xFFFFFFFFull);
        ln.march_capacity[1] = (uint32_t)std::min<size_t>(ray_cap * per_type[1], 0xFFFFFFFFull);
        ln.march_key.ensure((size_t)pool * sizeof(unsigned long long));
    }
    if (bidir) {
        ln.light_vertices.ensure((size_t)pool * (R.light_bounces + 1) * light_vertex_bytes());
        ln.cam_vertices.ensure((size_t)pool * std::max<uint32_t>(R.bounces, 1) * cam_vertex_bytes());
    }
    ln.pool = pool;
}

// Paths in flight.  Every wavefront iteration pays fixed costs (the tail of the persistent traversal kernel, launch
// ramps, three small launches), so large pools pay: C2 runs 11 % faster with 2^24 paths in flight than with 2^21.
// The library default is 2^24, reduced so that the pool's buffers stay within a sixth of the device memory (30 GB of
// a B200's 180 GB).
size_t pool_bytes_per_path(const pyr_ctx* ctx) {
    const RendererRec& R = ctx->view.renderer;
    const bool bidir = R.algorithm == 1;
    const size_t shadow = bidir ? (size_t)bdpt_stage_rays() : std::max<uint32_t>(R.light_samples, 1);
    size_t b = path_state_bytes() + MAX_LIGHT_SAMPLES * pending_light_bytes() + (bidir ? 3 : 2) * sizeof(uint32_t) + sizeof(uint16_t) +
               (1 + shadow) * 2 * sizeof(Ray) + sizeof(Hit) + shadow * sizeof(uint32_t);
    if (ctx->view.n_marched) b += (1 + shadow) * (size_t)ctx->view.n_marched * sizeof(uint2) + sizeof(unsigned long long);
    if (bidir) b += bidir_state_bytes() + (size_t)(R.light_bounces + 1) * light_vertex_bytes() + (size_t)std::max<uint32_t>(R.bounces, 1) * cam_vertex_bytes();
    return b;
}
uint32_t default_pool(const pyr_ctx* ctx) {
    size_t budget = ctx->device_memory / 6;
    if (budget == 0) budget = (size_t)30 << 30;
    const size_t pool = budget / pool_bytes_per_path(ctx);
    return (uint32_t)std::min<size_t>(std::max<size_t>(pool, 4096), (size_t)1 << 24);
}

}  // namespace

extern "C" {

const char* pyr_version(void) {
    static std::string text;
    text = "pyrite_b200 0.1; sm_100a";
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
        text += std::string("; ") + prop.name + "; " + std::to_string(prop.multiProcessorCount) + " SMs";
    else
        text += "; no CUDA device";
    return text.c_str();
}

const char* pyr_last_error(const pyr_ctx* ctx) { return ctx ? ctx->error.c_str() : g_init_error.c_str(); }

pyr_status pyr_init(int32_t device, pyr_ctx** out) {
    if (!out) return fail(nullptr, PYR_ERR_INVALID, "null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, PYR_ERR_CUDA, std::string("no usable CUDA device (there is no CPU path): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) return fail(nullptr, PYR_ERR_INVALID, "device index out of range");
    std::unique_ptr<pyr_ctx> ctx(new pyr_ctx);
    ctx->device = device;
    try {
        CU(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        ctx->device_memory = prop.totalGlobalMem;
        CU(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        CU(cudaEventCreate(&ctx->ev0));
        CU(cudaEventCreate(&ctx->ev1));
        CU(cudaMallocHost((void**)&ctx->pinned, 4 * sizeof(unsigned long long)));
        CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        for (Lane& ln : ctx->lanes) {
            CU(cudaStreamCreateWithFlags(&ln.own_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ln.batch_done, cudaEventDisableTiming));
            CU(cudaMallocHost((void**)&ln.pinned, 4 * sizeof(unsigned long long)));
        }
        ctx->counters.ensure(sizeof(DeviceCounters));
        CU(cudaMemsetAsync(ctx->counters.p, 0, sizeof(DeviceCounters), ctx->stream));
        ctx->scalars.ensure(16 * sizeof(uint32_t));
        CU(cudaMemsetAsync(ctx->scalars.p, 0, 16 * sizeof(uint32_t), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    } catch (const std::exception& ex) {
        return fail(nullptr, PYR_ERR_CUDA, ex.what());
    }
    *out = ctx.release();
    return PYR_OK;
}

void pyr_shutdown(pyr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    open_comm_gate(ctx);
    if (ctx->comm_thread.joinable()) ctx->comm_thread.join();
    if (ctx->comm) { try { nccl().CommDestroy(ctx->comm); } catch (...) {} ctx->comm = nullptr; }
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    ctx->comm_scratch.release();
    DeviceBuffer* all[] = {&ctx->nodes, &ctx->prims, &ctx->tri_shade, &ctx->tri_frames, &ctx->planes, &ctx->marched, &ctx->materials,
                           &ctx->components, &ctx->programs, &ctx->code, &ctx->spectra, &ctx->spectrum_data, &ctx->textures, &ctx->texels,
                           &ctx->lamps, &ctx->tiles, &ctx->burns, &ctx->xyz, &ctx->d65, &ctx->film, &ctx->develop_params, &ctx->counters,
                           &ctx->scalars, &ctx->tile_first, &ctx->scratch_a, &ctx->scratch_b, &ctx->bvh_scratch};
    for (Lane& ln : ctx->lanes) {
        ln.release();
        if (ln.pinned) cudaFreeHost(ln.pinned);
        if (ln.batch_done) cudaEventDestroy(ln.batch_done);
        if (ln.own_stream) cudaStreamDestroy(ln.own_stream);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (DeviceBuffer* b : all) b->release();
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

pyr_status pyr_stream_set(pyr_ctx* ctx, void* cuda_stream) {
    return guarded(ctx, [&] {
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    });
}

pyr_status pyr_project_load(pyr_ctx* ctx, const void* ir_blob, size_t bytes) {
    return guarded(ctx, [&] {
        if (!ir_blob) throw ir::BuildError("null project IR");
        const auto t_load = std::chrono::steady_clock::now();
        ir::Document doc = ir::decode(ir_blob, bytes);
        cudaStream_t s = ctx->stream;
        // The BVH of a big scene is built on the GPU (bvh_build.cu: the reference's tree, level by level); small ones, where the
        // launches would cost more than the work, on the host.  PYR_BVH_BUILD=host / gpu forces one of them.
        // (its scratch block stays with the context for the next load: releasing it cost 30 - 240 ms of cudaFree)
        const BvhBuildFn gpu_builder = [s, ctx](const float* boxes6, size_t n, const float* hull12, BvhTree& tree) {
            gpu_bvh_build(boxes6, n, hull12, tree, s, [ctx](size_t bytes) { ctx->bvh_scratch.ensure(bytes); return ctx->bvh_scratch.p; });
        };
        size_t n_bvh_items = 0;
        for (const auto& o : doc.objects)
            if (o.kind == ir::OBJ_MESH && o.mesh < doc.meshes.size())
                for (const auto& part : doc.meshes[o.mesh].objects) n_bvh_items += part.corners.size() / 9;
        bool on_gpu = n_bvh_items >= 16384;
        if (const char* e = getenv("PYR_BVH_BUILD")) on_gpu = std::string(e) == "gpu";
        BakedScene baked = build_scene(doc, on_gpu ? &gpu_builder : nullptr);
        ctx->bvh_built_on_gpu = on_gpu && baked.n_objects >= 2;
        // From here on the old project's buffers are freed / overwritten in place: the context holds NO project until the
        // last upload has completed, so a failure half-way (out of device memory on a bigger scene) leaves it in the
        // "nothing loaded" state - the next pyr_render / pyr_trace / pyr_film_* returns PYR_ERR_STATE instead of touching
        // freed memory.
        CU(cudaStreamSynchronize(s));
        ctx->loaded = false;
        for (Lane& ln : ctx->lanes) ln.pool = 0;
        ctx->develop_params_valid = false;
        upload(ctx->nodes, baked.nodes, s);
        upload(ctx->prims, baked.prims, s);
        upload(ctx->tri_shade, baked.tri_shade, s);
        upload(ctx->tri_frames, baked.tri_frames, s);
        upload(ctx->planes, baked.planes, s);
        upload(ctx->marched, baked.marched, s);
        upload(ctx->materials, baked.materials, s);
        upload(ctx->components, baked.components, s);
        upload(ctx->programs, baked.programs, s);
        upload(ctx->code, baked.code, s);
        upload(ctx->spectra, baked.spectra, s);
        upload(ctx->spectrum_data, baked.spectrum_data, s);
        upload(ctx->textures, baked.textures, s);
        upload(ctx->texels, baked.texels, s);
        upload(ctx->lamps, baked.lamps, s);
        upload(ctx->tiles, baked.tiles, s);
        upload(ctx->burns, baked.burns, s);
        upload(ctx->xyz, baked.xyz, s);
        upload(ctx->d65, baked.d65, s);
        SceneView v = baked.view;
        v.nodes = ctx->nodes.as<Node4>(); v.prims = ctx->prims.as<Prim>(); v.tri_shade = ctx->tri_shade.as<TriShade>();
        v.tri_frames = ctx->tri_frames.as<TriFrames>(); v.planes = ctx->planes.as<PlaneRec>(); v.marched = ctx->marched.as<MarchedRec>();
        v.materials = ctx->materials.as<MaterialRec>(); v.components = ctx->components.as<ComponentRec>();
        v.programs = ctx->programs.as<ProgramRec>(); v.code = ctx->code.as<Instr>(); v.spectra = ctx->spectra.as<SpectrumRec>();
        v.spectrum_data = ctx->spectrum_data.as<float>(); v.textures = ctx->textures.as<TextureRec>(); v.texels = ctx->texels.as<float>();
        v.lamps = ctx->lamps.as<LampRec>(); v.tiles = ctx->tiles.as<TileRec>(); v.burns = ctx->burns.as<float>();
        v.xyz = ctx->xyz.as<float>(); v.d65 = ctx->d65.as<float>();
        ctx->view = v;
        ctx->scene = std::move(baked);
        ctx->film.ensure(ctx->film_floats() * sizeof(float));
        CU(cudaMemsetAsync(ctx->film.p, 0, ctx->film_floats() * sizeof(float), s));
        ctx->develop_params_valid = false;
        for (Lane& ln : ctx->lanes) ln.pool = 0;
        CU(cudaStreamSynchronize(s));
        ctx->loaded = true;
        ctx->load_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
    });
}

pyr_status pyr_project_info_get(const pyr_ctx* ctx, pyr_project_info* out) {
    if (!ctx || !out) return PYR_ERR_INVALID;
    if (!ctx->loaded) return PYR_ERR_STATE;
    const RendererRec& r = ctx->view.renderer;
    *out = pyr_project_info{r.width, r.height, r.spectrum_bins, r.algorithm, r.pixel_samples, r.bounces, r.light_samples, r.spectrum_samples,
                            r.light_bounces, r.tile_size, ctx->scene.n_objects, ctx->view.n_planes, ctx->view.n_lamps,
                            ctx->scene.n_objects ? 2 * ctx->scene.n_objects - 1 : 0, (uint32_t)ctx->scene.materials.size(), ctx->view.n_marched};
    return PYR_OK;
}

pyr_status pyr_trace_device(pyr_ctx* ctx, const void* d_rays, size_t n, void* d_hits, uint32_t repeat) {
    return guarded(ctx, [&] {
        if (!check_trace_args(ctx, d_rays, n, d_hits, true)) return;
        const int blocks = ctx->sm_count * trace_blocks_per_sm();
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        for (uint32_t r = 0; r < (repeat ? repeat : 1u); ++r) {
            CU(cudaMemsetAsync(ctx->cursor(), 0, sizeof(uint32_t), ctx->stream));
            launch_trace_batch(ctx->view, d_rays, n, d_hits, ctx->cursor(), ctx->counters.as<DeviceCounters>(), 0, blocks, ctx->stream);
            CU(cudaGetLastError());
            ctx->host_counters.kernel_launches += 1;
        }
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->host_counters.render_seconds = ms * 1e-3;
    });
}

pyr_status pyr_trace_stats(pyr_ctx* ctx, const pyr_ray* rays, size_t n, pyr_hit* hits_out) {
    return guarded(ctx, [&] {
        if (!check_trace_args(ctx, rays, n, hits_out, false)) return;
        ctx->scratch_a.ensure(n * sizeof(pyr_ray));
        ctx->scratch_b.ensure(n * sizeof(pyr_hit));
        CU(cudaMemcpyAsync(ctx->scratch_a.p, rays, n * sizeof(pyr_ray), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->cursor(), 0, sizeof(uint32_t), ctx->stream));
        launch_trace_batch(ctx->view, ctx->scratch_a.p, n, ctx->scratch_b.p, ctx->cursor(), ctx->counters.as<DeviceCounters>(), 1,
                           ctx->sm_count * trace_blocks_per_sm(), ctx->stream);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hits_out, ctx->scratch_b.p, n * sizeof(pyr_hit), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->host_counters.kernel_launches += 1;
    });
}

pyr_status pyr_trace(pyr_ctx* ctx, const pyr_ray* rays, size_t n, pyr_hit* hits_out) {
    return guarded(ctx, [&] {
        if (!check_trace_args(ctx, rays, n, hits_out, false)) return;
        ctx->scratch_a.ensure(n * sizeof(pyr_ray));
        ctx->scratch_b.ensure(n * sizeof(pyr_hit));
        CU(cudaMemcpyAsync(ctx->scratch_a.p, rays, n * sizeof(pyr_ray), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->cursor(), 0, sizeof(uint32_t), ctx->stream));
        launch_trace_batch(ctx->view, ctx->scratch_a.p, n, ctx->scratch_b.p, ctx->cursor(), ctx->counters.as<DeviceCounters>(), 0,
                           ctx->sm_count * trace_blocks_per_sm(), ctx->stream);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hits_out, ctx->scratch_b.p, n * sizeof(pyr_hit), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->host_counters.kernel_launches += 1;
    });
}

pyr_status pyr_render(pyr_ctx* ctx, const pyr_render_params* params, pyr_progress_cb cb, void* user) {
    return guarded(ctx, [&] {
        need_project(ctx);
        pyr_render_params p{};
        if (params) p = *params;
        const RendererRec& R = ctx->view.renderer;
        const uint32_t spp = p.spp_override ? p.spp_override : R.pixel_samples;
        const uint32_t stride = p.sample_stride ? p.sample_stride : 1u;
        const uint32_t offset = p.sample_offset;
        if (p.tile_filter > ctx->scene.tiles.size()) throw ir::BuildError("tile_filter out of range");
        if (ctx->view.n_lamps == 0) {
            // World::pick_lamp panics on an empty lamp list (`gen_range(0..0)`, world.rs:301-305); the reference reaches it
            // from trace_direct at the first diffuse bounce (even with light_samples = 0) and from every bidirectional sample
            bool any_diffuse = false;
            for (const ComponentRec& c : ctx->scene.components) any_diffuse = any_diffuse || c.bsdf == BSDF_DIFFUSE;
            if (R.algorithm == 1 || any_diffuse)
                throw ir::BuildError("the scene has no lamps: World::pick_lamp would panic (world.rs:301-305)");
        }
        cudaStream_t s = ctx->stream;

        // per-tile sample counts for this shard: i = offset, offset + stride, ... < area * spp
        std::vector<unsigned long long> first(ctx->scene.tiles.size() + 1, 0);
        for (size_t t = 0; t < ctx->scene.tiles.size(); ++t) {
            unsigned long long iterations = (unsigned long long)ctx->scene.tiles[t].width * ctx->scene.tiles[t].height * spp;
            unsigned long long mine = offset < iterations ? (iterations - offset + stride - 1) / stride : 0;
            if (p.tile_filter && p.tile_filter != t + 1) mine = 0;
            first[t + 1] = first[t] + mine;
        }
        const unsigned long long total = first.back();
        upload(ctx->tile_first, first, s);

        const int stats = (p.flags & PYR_RENDER_STATS) ? 1 : 0;
        const bool timing = (p.flags & PYR_RENDER_TIMING) != 0;
        uint32_t pool = p.pool_paths ? p.pool_paths : default_pool(ctx);
        if ((unsigned long long)pool > total) pool = (uint32_t)std::max<unsigned long long>(total, 1);
        // Two wavefronts of half the pool on two streams when the job is big enough to fill both (see Lane); the per-kernel
        // timing and statistics modes run one wavefront so that their numbers are those of a kernel running alone.
        int n_lanes = (!timing && !stats && pool >= (1u << 20)) ? 2 : 1;
        if (const char* e = getenv("PYR_LANES")) n_lanes = std::max(1, std::min(MAX_LANES, atoi(e)));
        if (timing || stats) n_lanes = 1;
        const uint32_t lane_pool = ((pool + n_lanes - 1) / n_lanes + 127u) & ~127u;
        for (int l = 0; l < n_lanes; ++l) ensure_pool(ctx, ctx->lanes[l], lane_pool);
        if (p.reset_film) CU(cudaMemsetAsync(ctx->film.p, 0, ctx->film_floats() * sizeof(float), s));
        CU(cudaMemsetAsync(ctx->scalars.p, 0, 16 * sizeof(uint32_t), s));

        const int trace_blocks = ctx->sm_count * trace_blocks_per_sm();
        const int BATCH = 4;
        if (timing)
            while (ctx->timing_events.size() < (size_t)BATCH * 3) {
                cudaEvent_t e;
                CU(cudaEventCreate(&e));
                ctx->timing_events.push_back(e);
            }
        uint32_t cluster_shift = 0;  // the hit primitive's rank >> shift = one of (at most) BIN_CLUSTERS subtrees of the BVH
        while ((ctx->view.n_prims >> cluster_shift) > BIN_CLUSTERS) ++cluster_shift;
        CU(cudaEventRecord(ctx->ev0, s));
        if (n_lanes > 1) CU(cudaEventRecord(ctx->ev_fork, s));
        for (int l = 0; l < n_lanes; ++l) {
            Lane& ln = ctx->lanes[l];
            ln.stream = n_lanes > 1 ? ln.own_stream : s;
            if (n_lanes > 1) CU(cudaStreamWaitEvent(ln.stream, ctx->ev_fork, 0));
            ln.cur = 0;
            ln.grid_paths = lane_pool;  // once every sample has been started the live-slot count only falls: the last value read bounds the grids
            ln.finished = false;
            CU(cudaMemsetAsync(ln.scalars.p, 0, 16 * sizeof(uint32_t), ln.stream));
            CU(cudaMemsetAsync(ln.bin_count.p, 0, NUM_KEYS * sizeof(uint32_t), ln.stream));
            launch_pool_reset(ln.paths.as<PathCore>(), lane_pool, ln.live_list.as<uint32_t>(), ln.live_count(0), ln.stream);
        }
        unsigned long long iterations = 0, launches = (unsigned long long)n_lanes;

        // one wavefront iteration of a lane: sort the live slots, shade / regenerate, trace (and sphere-trace) what they asked for
        auto enqueue_batch = [&](Lane& ln) {
            cudaStream_t ls = ln.stream;
            const BinBuffers bins{ln.bin_count.as<uint32_t>(), ln.bin_first.as<uint32_t>(), ln.bin_fill.as<uint32_t>(), ln.bin_keys.as<uint16_t>(),
                                  ln.bin_list.as<uint32_t>()};
            for (int b = 0; b < BATCH; ++b) {
                const int cur = ln.cur, nxt = cur ^ 1;
                WaveArgs a{};
                a.paths = ln.paths.as<PathCore>();
                a.pend = ln.pend.as<PendingLight>();
                a.bidir = ln.bidir.as<BidirState>();
                a.pool = lane_pool;
                a.grid_paths = std::max<uint32_t>(ln.grid_paths, 1);
                a.rays_in = ln.rays[cur].as<Ray>();
                a.hits_in = ln.hits.as<Hit>();
                a.shadow_kinds_in = ln.shadow_kinds.as<uint32_t>();
                a.rays_out = ln.rays[nxt].as<Ray>();
                a.count_out = ln.count(nxt);
                a.shadow_offset = lane_pool;
                a.trace_cursor = ln.cursor();
                a.next_sample = ctx->next_sample();
                a.total_samples = total;
                a.tile_first = ctx->tile_first.as<unsigned long long>();
                a.seed = p.seed;
                a.sample_offset = offset;
                a.sample_stride = stride;
                a.film = ctx->film.as<float>();
                a.counters = ctx->counters.as<DeviceCounters>();
                a.light_vertices = ln.light_vertices.as<LightVertex>();
                a.cam_vertices = ln.cam_vertices.as<CamVertex>();
                a.light_stride = R.light_bounces + 1;
                a.cam_stride = std::max<uint32_t>(R.bounces, 1);
                a.ray_capacity = (uint32_t)(ln.rays[0].bytes / sizeof(Ray));
                if (timing) CU(cudaEventRecord(ctx->timing_events[3 * b], ls));
                a.live_list = ln.live_list.as<uint32_t>();
                a.live_count_in = ln.live_count(cur);
                a.live_count_out = ln.live_count(nxt);
                a.died_list = ln.died_list.as<uint32_t>();
                a.died_count = ln.died_count();
                a.bin_first = ln.bin_first.as<uint32_t>();
                a.bin_list = ln.bin_list.as<uint32_t>();
                launch_bin(a, bins, cluster_shift, R.algorithm == 1, ls);
                if (R.algorithm == 0) launch_wave_simple(ctx->view, a, ls); else launch_wave_bidirectional(ctx->view, a, ctx->sm_count, ls);
                if (timing) CU(cudaEventRecord(ctx->timing_events[3 * b + 1], ls));
                TraceArgs t{};
                t.rays = ln.rays[nxt].as<Ray>();
                t.hits = ln.hits.as<Hit>();
                t.shadow_kinds = ln.shadow_kinds.as<uint32_t>();
                t.count = ln.count(nxt);
                t.shadow_offset = lane_pool;
                t.cursor = ln.cursor();
                t.counters = ctx->counters.as<DeviceCounters>();
                t.stats = stats;
                t.march_queue[0] = ln.march_queue[0].as<uint2>();
                t.march_queue[1] = ln.march_queue[1].as<uint2>();
                t.march_count = ln.march_count();
                t.march_capacity[0] = ln.march_capacity[0];
                t.march_capacity[1] = ln.march_capacity[1];
                t.march_key = ln.march_key.as<unsigned long long>();
                if (ctx->view.n_marched) CU(cudaMemsetAsync(ln.march_count(), 0, 4 * sizeof(uint32_t), ls));
                launch_trace(ctx->view, t, trace_blocks, ls);
                if (ctx->view.n_marched) { launch_march(ctx->view, t, trace_blocks, ls); launches += 2; }
                if (timing) CU(cudaEventRecord(ctx->timing_events[3 * b + 2], ls));
                ln.cur = nxt;
                ++iterations;
                launches += R.algorithm == 0 ? 5 : 4 + wave_bidirectional_launches();
            }
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(&ln.pinned[0], ln.count(ln.cur), 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ls));
            CU(cudaMemcpyAsync(&ln.pinned[1], ctx->next_sample(), sizeof(unsigned long long), cudaMemcpyDeviceToHost, ls));
            CU(cudaMemcpyAsync(&ln.pinned[2], ln.live_count(ln.cur), sizeof(uint32_t), cudaMemcpyDeviceToHost, ls));
            CU(cudaEventRecord(ln.batch_done, ls));
        };

        for (int l = 0; l < n_lanes; ++l) enqueue_batch(ctx->lanes[l]);
        open_comm_gate(ctx);   // a communicator waiting to be set up (pyr_comm_init_async) starts now, under the queued iterations
        bool cancelled = false;
        unsigned long long started = 0;
        int active = n_lanes;
        while (active > 0 && !cancelled) {
            for (int l = 0; l < n_lanes && !cancelled; ++l) {
                Lane& ln = ctx->lanes[l];
                if (ln.finished) continue;
                CU(cudaEventSynchronize(ln.batch_done));   // the other lane's batch keeps the GPU busy meanwhile
                if (timing)
                    for (int b = 0; b < BATCH; ++b) {
                        float shade_ms = 0, trace_ms = 0;
                        CU(cudaEventElapsedTime(&shade_ms, ctx->timing_events[3 * b], ctx->timing_events[3 * b + 1]));
                        CU(cudaEventElapsedTime(&trace_ms, ctx->timing_events[3 * b + 1], ctx->timing_events[3 * b + 2]));
                        ctx->host_counters.shade_seconds += shade_ms * 1e-3;
                        ctx->host_counters.trace_seconds += trace_ms * 1e-3;
                        ctx->host_counters.shade_launches += 1;
                        ctx->host_counters.trace_launches += 1;
                    }
                const uint32_t pending = (uint32_t)(ln.pinned[0] & 0xffffffffull) + (uint32_t)(ln.pinned[0] >> 32);
                started = std::max(started, std::min<unsigned long long>(ln.pinned[1], total));
                if (pending == 0 && started >= total) { ln.finished = true; --active; continue; }
                if (started >= total) ln.grid_paths = (uint32_t)(ln.pinned[2] & 0xffffffffull);
                if (cb) {
                    uint8_t progress = total ? (uint8_t)((started * 100ull) / total) : 100;
                    if (cb(progress, "rendering", user)) { cancelled = true; break; }
                }
                enqueue_batch(ln);
            }
        }
        for (int l = 0; l < n_lanes; ++l) {   // join the lanes' streams into the context's stream
            Lane& ln = ctx->lanes[l];
            if (n_lanes > 1) { CU(cudaEventRecord(ln.batch_done, ln.stream)); CU(cudaStreamWaitEvent(s, ln.batch_done, 0)); }
        }
        CU(cudaEventRecord(ctx->ev1, s));
        if (ctx->view.n_marched)
            CU(cudaMemcpyAsync(&ctx->pinned[3], &ctx->counters.as<DeviceCounters>()->march_overflow, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->host_counters.render_seconds = ms * 1e-3;
        if (ctx->view.n_marched && ctx->pinned[3]) {
            CU(cudaMemsetAsync(&ctx->counters.as<DeviceCounters>()->march_overflow, 0, sizeof(unsigned long long), s));
            throw StateError("the sphere-tracing queue overflowed: " + std::to_string(ctx->pinned[3]) + " candidates were dropped, the film is incomplete");
        }
        ctx->host_counters.wavefront_iterations += iterations;
        ctx->host_counters.kernel_launches += launches;
        ctx->host_counters.path_samples += cancelled ? started : total;
        if (cb && !cancelled) cb(100, "done", user);
        if (cancelled) throw StateError("render cancelled by the progress callback");
    });
}

pyr_status pyr_film_expose(pyr_ctx* ctx, const float* positions, const float* samples, size_t n) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (n == 0) return;
        if (!positions || !samples) throw ir::BuildError("null sample buffers");
        ctx->scratch_a.ensure(n * 2 * sizeof(float));
        ctx->scratch_b.ensure(n * 3 * sizeof(float));
        CU(cudaMemcpyAsync(ctx->scratch_a.p, positions, n * 2 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->scratch_b.p, samples, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        launch_film_expose(ctx->view, ctx->film.as<float>(), ctx->scratch_a.as<float>(), ctx->scratch_b.as<float>(), n, ctx->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->host_counters.kernel_launches += 1;
    });
}

pyr_status pyr_film_clear(pyr_ctx* ctx) {
    return guarded(ctx, [&] {
        need_project(ctx);
        CU(cudaMemsetAsync(ctx->film.p, 0, ctx->film_floats() * sizeof(float), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    });
}

pyr_status pyr_film_download(pyr_ctx* ctx, float* out) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (!out) throw ir::BuildError("null film buffer");
        CU(cudaMemcpyAsync(out, ctx->film.p, ctx->film_floats() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    });
}

pyr_status pyr_film_upload(pyr_ctx* ctx, const float* in) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (!in) throw ir::BuildError("null film buffer");
        CU(cudaMemcpyAsync(ctx->film.p, in, ctx->film_floats() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    });
}

pyr_status pyr_film_device_ptr(pyr_ctx* ctx, void** d_ptr, size_t* bytes) {
    return guarded(ctx, [&] {
        need_project(ctx);
        CU(cudaStreamSynchronize(ctx->stream));
        if (d_ptr) *d_ptr = ctx->film.p;
        if (bytes) *bytes = ctx->film_floats() * sizeof(float);
    });
}

// ---- the one collective of the path: summing the films of the ranks (NCCL over NVLink)
pyr_status pyr_comm_unique_id(uint8_t* id_out) {
    if (!id_out) return fail(nullptr, PYR_ERR_INVALID, "null id buffer");
    try {
        static_assert(sizeof(ncclUniqueId) == PYR_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
        ncclUniqueId id;
        NC(nccl().GetUniqueId(&id));
        memcpy(id_out, &id, sizeof(id));
        return PYR_OK;
    } catch (const StateError& e) {
        return fail(nullptr, PYR_ERR_STATE, e.what());
    } catch (const std::exception& e) {
        return fail(nullptr, PYR_ERR_CUDA, e.what());
    }
}

namespace {
// Joins the communicator set-up started by pyr_comm_init_async and reports its failure, if any.
void comm_wait(pyr_ctx* ctx) {
    open_comm_gate(ctx);
    if (ctx->comm_thread.joinable()) ctx->comm_thread.join();
    if (!ctx->comm_error.empty()) {
        const std::string e = ctx->comm_error;
        ctx->comm_error.clear();
        throw CudaError("communicator set-up failed: " + e);
    }
}
}  // namespace

pyr_status pyr_comm_init_async(pyr_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id) {
    return guarded(ctx, [&] {
        if (!id) throw ir::BuildError("null communicator id");
        if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw ir::BuildError("bad rank / rank count");
        comm_wait(ctx);
        if (ctx->comm) { NC(nccl().CommDestroy(ctx->comm)); ctx->comm = nullptr; ctx->comm_ranks = 0; }
        nccl();  // bind NCCL here, so that a missing library is reported by this call
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        if (!ctx->comm_stream) CU(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
        ctx->comm_scratch.ensure(1024);
        CU(cudaMemsetAsync(ctx->comm_scratch.p, 0, 1024, ctx->comm_stream));
        CU(cudaStreamSynchronize(ctx->comm_stream));
        ctx->comm_ranks = n_ranks;
        ctx->comm_rank = rank;
        {
            const char* e = getenv("PYR_COMM_GATE");   // PYR_COMM_GATE=0: start the set-up at once (A/B)
            std::lock_guard<std::mutex> g(ctx->comm_gate_mutex);
            ctx->comm_gate_open = e && atoi(e) == 0;
        }
        ctx->comm_thread = std::thread([ctx, uid, n_ranks, rank] {
            try {
                {
                    std::unique_lock<std::mutex> lk(ctx->comm_gate_mutex);
                    ctx->comm_gate_cv.wait_for(lk, std::chrono::seconds(5), [ctx] { return ctx->comm_gate_open; });
                }
                CU(cudaSetDevice(ctx->device));
                NC(nccl().CommInitRank(&ctx->comm, n_ranks, uid, rank));
                // NCCL connects the ranks at the first collective: do that here, on a few bytes and on a stream that no render
                // uses, so that pyr_film_reduce costs what moving the film costs
                NC(nccl().AllReduce(ctx->comm_scratch.p, ctx->comm_scratch.p, 256, ncclFloat32, ncclSum, ctx->comm, ctx->comm_stream));
                CU(cudaStreamSynchronize(ctx->comm_stream));
            } catch (const std::exception& e) {
                ctx->comm_error = e.what();
                if (ctx->comm_error.empty()) ctx->comm_error = "unknown error";
            }
        });
    });
}

pyr_status pyr_comm_init(pyr_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id) {
    const pyr_status st = pyr_comm_init_async(ctx, n_ranks, rank, id);
    if (st != PYR_OK) return st;
    return guarded(ctx, [&] {
        try { comm_wait(ctx); } catch (...) { ctx->comm = nullptr; ctx->comm_ranks = 0; throw; }
    });
}

pyr_status pyr_comm_destroy(pyr_ctx* ctx) {
    return guarded(ctx, [&] {
        try { comm_wait(ctx); } catch (...) { ctx->comm = nullptr; ctx->comm_ranks = 0; throw; }
        if (!ctx->comm) return;
        CU(cudaStreamSynchronize(ctx->stream));
        NC(nccl().CommDestroy(ctx->comm));
        ctx->comm = nullptr;
        ctx->comm_ranks = 0;
    });
}

pyr_status pyr_film_reduce(pyr_ctx* ctx, int32_t root) {
    return guarded(ctx, [&] {
        need_project(ctx);
        comm_wait(ctx);
        if (!ctx->comm) throw StateError("pyr_comm_init has not been called on this context");
        if (root >= ctx->comm_ranks) throw ir::BuildError("root rank out of range");
        // (accumulator, weight) pairs are summed BEFORE developing: a bin's value is the ratio of the two sums (film.rs:132-143)
        float* film = ctx->film.as<float>();
        if (root < 0) NC(nccl().AllReduce(film, film, ctx->film_floats(), ncclFloat32, ncclSum, ctx->comm, ctx->stream));
        else NC(nccl().Reduce(film, film, ctx->film_floats(), ncclFloat32, ncclSum, root, ctx->comm, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    });
}

pyr_status pyr_film_develop(pyr_ctx* ctx, float step_size, float* xyz_out, uint8_t* srgb_out) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (!(step_size > 0.0f)) throw ir::BuildError("step_size must be positive");
        const size_t pixels = (size_t)ctx->view.film.width * ctx->view.film.height;
        ensure_develop_params(ctx);
        ctx->scratch_a.ensure(pixels * 3 * sizeof(float));
        ctx->scratch_b.ensure(pixels * 3);
        launch_develop(ctx->view, ctx->film.as<float>(), ctx->develop_params.as<float>(), step_size, ctx->scratch_a.as<float>(),
                       ctx->scratch_b.as<uint8_t>(), ctx->stream);
        CU(cudaGetLastError());
        if (xyz_out) CU(cudaMemcpyAsync(xyz_out, ctx->scratch_a.p, pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (srgb_out) CU(cudaMemcpyAsync(srgb_out, ctx->scratch_b.p, pixels * 3, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->host_counters.kernel_launches += 1;
    });
}

pyr_status pyr_camera_sample(pyr_ctx* ctx, uint64_t seed, uint32_t tile, uint64_t sample, float* position_out, pyr_ray* ray_out,
                             float* wavelengths_out, uint32_t* hero_out) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (tile >= ctx->view.n_tiles) throw ir::BuildError("tile index out of range");
        const size_t floats = 2 + 8 + MAX_SPECTRUM_SAMPLES + 1;
        ctx->scratch_a.ensure(floats * sizeof(float));
        launch_camera_sample(ctx->view, seed, tile, sample, ctx->scratch_a.as<float>(), ctx->stream);
        CU(cudaGetLastError());
        float host[2 + 8 + MAX_SPECTRUM_SAMPLES + 1];
        CU(cudaMemcpyAsync(host, ctx->scratch_a.p, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (position_out) { position_out[0] = host[0]; position_out[1] = host[1]; }
        if (ray_out) memcpy(ray_out, host + 2, sizeof(pyr_ray));
        if (wavelengths_out) for (uint32_t k = 0; k < ctx->view.renderer.spectrum_samples; ++k) wavelengths_out[k] = host[10 + k];
        if (hero_out) memcpy(hero_out, host + 10 + MAX_SPECTRUM_SAMPLES, sizeof(uint32_t));
    });
}

pyr_status pyr_debug_path(pyr_ctx* ctx, uint64_t seed, uint32_t tile, uint64_t sample, uint32_t max_bounces, uint32_t* records_out,
                          uint32_t* n_bounces_out, float* exposed_out, uint32_t* n_exposed_out, float* position_out) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (tile >= ctx->view.n_tiles) throw ir::BuildError("tile index out of range");
        if (!records_out || !n_bounces_out || !exposed_out || !n_exposed_out || !position_out) throw ir::BuildError("null output buffer");
        const size_t words = (size_t)20 * max_bounces + 2 + 2 * MAX_SPECTRUM_SAMPLES + 2;
        ctx->scratch_a.ensure(words * sizeof(uint32_t));
        uint32_t* d = ctx->scratch_a.as<uint32_t>();
        uint32_t* d_counts = d + (size_t)20 * max_bounces;
        float* d_exposed = reinterpret_cast<float*>(d_counts + 2);
        float* d_position = d_exposed + 2 * MAX_SPECTRUM_SAMPLES;
        CU(cudaMemsetAsync(d, 0, words * sizeof(uint32_t), ctx->stream));
        ctx->scratch_b.ensure(debug_path_scratch_bytes());
        launch_debug_path(ctx->view, seed, tile, sample, max_bounces, d, d_counts, d_exposed, d_position, ctx->scratch_b.p, ctx->stream);
        CU(cudaGetLastError());
        std::vector<uint32_t> host(words);
        CU(cudaMemcpyAsync(host.data(), d, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        memcpy(records_out, host.data(), (size_t)20 * max_bounces * sizeof(uint32_t));
        const uint32_t* counts = host.data() + (size_t)20 * max_bounces;
        *n_bounces_out = counts[0];
        *n_exposed_out = counts[1];
        memcpy(exposed_out, counts + 2, 2 * MAX_SPECTRUM_SAMPLES * sizeof(float));
        memcpy(position_out, counts + 2 + 2 * MAX_SPECTRUM_SAMPLES, 2 * sizeof(float));
        ctx->host_counters.kernel_launches += 1;
    });
}

pyr_status pyr_counters_get(pyr_ctx* ctx, pyr_counters* out, int32_t reset) {
    return guarded(ctx, [&] {
        DeviceCounters dc;
        CU(cudaMemcpyAsync(&dc, ctx->counters.p, sizeof(dc), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (out) {
            *out = ctx->host_counters;
            out->rays = dc.rays;
            out->nodes_visited = dc.nodes_visited;
            out->leaves_tested = dc.leaves_tested;
            out->de_evals = dc.de_evals;
            out->de_iterations = dc.de_iterations;
            out->march_iterations = dc.march_iterations;
            out->julia_iterations = dc.julia_iterations;
            out->node_fetches = dc.node_fetches;
            out->path_rays = dc.path_rays;
        }
        if (reset) {
            CU(cudaMemsetAsync(ctx->counters.p, 0, sizeof(DeviceCounters), ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            double secs = ctx->host_counters.render_seconds;
            ctx->host_counters = pyr_counters{};
            ctx->host_counters.render_seconds = secs;
        }
    });
}

// BVH leaf pre-order: object id of every rank (test hook for the tie rule of World::intersect)
pyr_status pyr_bvh_leaf_order(pyr_ctx* ctx, uint32_t* object_ids_out) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (!object_ids_out && ctx->scene.n_objects) throw ir::BuildError("null output buffer");
        for (uint32_t obj = 0; obj < ctx->scene.n_objects; ++obj) object_ids_out[ctx->scene.rank_of_object[obj]] = obj;
    });
}

// How the BVH of the loaded project was built, and a digest of it (test hook: the GPU build and the host build must agree)
pyr_status pyr_bvh_digest(pyr_ctx* ctx, uint64_t* out4) {
    return guarded(ctx, [&] {
        need_project(ctx);
        if (!out4) throw ir::BuildError("null output buffer");
        auto fnv = [](uint64_t h, uint32_t w) { for (int i = 0; i < 4; ++i) { h ^= (w >> (8 * i)) & 0xffu; h *= 1099511628211ull; } return h; };
        uint64_t h_nodes = 14695981039346656037ull, h_order = 14695981039346656037ull;
        for (const Node4& nd : ctx->scene.nodes) {
            uint32_t w[32];
            memcpy(w, &nd, sizeof(w));
            for (int i = 0; i < 24; ++i) if (w[i] == 0x80000000u) w[i] = 0;  // a box coordinate of -0 and one of +0 are the same box
            for (int i = 0; i < 28; ++i) h_nodes = fnv(h_nodes, w[i]);
        }
        std::vector<uint32_t> order(ctx->scene.n_objects);
        for (uint32_t obj = 0; obj < ctx->scene.n_objects; ++obj) order[ctx->scene.rank_of_object[obj]] = obj;
        for (uint32_t o : order) h_order = fnv(h_order, o);
        out4[0] = ctx->bvh_built_on_gpu ? 1 : 0;
        out4[1] = h_nodes;
        out4[2] = h_order;
        out4[3] = (uint64_t)(ctx->load_seconds * 1e6);
    });
}

}  // extern "C"
