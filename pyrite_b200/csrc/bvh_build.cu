// Bvh::new (pyrite/src/spatial/bvh.rs:13-155) on the GPU: the level-synchronous build of bvh_build_core.hpp as sm_100a kernels.
// One level of the tree = five launches over the item positions / the level's nodes:
//   k_bvh_assign   item -> bucket of its node; per (node, bucket) count and hull by warp-aggregated atomics; per-tile bucket counts
//   k_bvh_split    node -> cut, children (next level's nodes or leaves), the interior record
//   k_bvh_tile_scan / k_bvh_rank   stable rank of every item among the items of its bucket (prefix counts)
//   k_bvh_scatter  item -> its position in the next level (bucket-major, otherwise in the old order = merge_buckets)
// The result is the tree the depth-first host builder (scene_build.cpp TreeBuilder) makes, bit for bit; 871,200 triangles take
// a few milliseconds instead of a few hundred on the host's cores, which is what every rank of a multi-GPU job pays before it
// can render.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "bvh_build.hpp"
#include "bvh_build_core.hpp"
#include "kernels.hpp"
#include "project_ir.hpp"

namespace pyr {
namespace {

using namespace bvhb;

constexpr int TILE = 1024;          // item positions per block
constexpr unsigned FULL_MASK = 0xffffffffu;

#define BV(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("GPU BVH build: " #call ": ") + cudaGetErrorString(e_)); \
    } while (0)

struct Box6 { float v[6]; };

__global__ void __launch_bounds__(TILE) k_bvh_assign(const Box6* __restrict__ boxes, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ node_of_pos,
                                                     const LevelNode* __restrict__ nodes, uint32_t n, uint8_t* __restrict__ bucket, BucketStats* stats,
                                                     uint32_t* __restrict__ tile_hist) {
    __shared__ uint32_t s_hist[BUCKETS];
    if (threadIdx.x < BUCKETS) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t p = blockIdx.x * TILE + threadIdx.x;
    const uint32_t j = p < n ? node_of_pos[p] : NO_NODE;
    uint32_t b = BUCKET_NONE;
    if (j != NO_NODE) {
        const LevelNode nd = nodes[j];
        const Box6 box = boxes[ids[p]];
        b = item_bucket(nd, p, box.v);
        uint32_t key[12];
        item_keys(box.v, key);
        // the lanes of the warp that add to the same (node, bucket) record combine first: one atomic per record, key and warp
        const unsigned peers = __match_any_sync(__activemask(), (j << 3) | b);
        const int leader = __ffs((int)peers) - 1;
        const bool lead = (int)(threadIdx.x & 31) == leader;
        uint32_t* rec = reinterpret_cast<uint32_t*>(stats + j);
        if (lead) atomicAdd(rec + b, (uint32_t)__popc(peers));
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const uint32_t m = __reduce_max_sync(peers, key[k]);
            if (lead) atomicMax(rec + BUCKETS + b * 12 + k, m);
        }
    }
    if (p < n) bucket[p] = (uint8_t)b;
    for (int k = 0; k < BUCKETS; ++k) {
        const unsigned m = __ballot_sync(FULL_MASK, b == (uint32_t)k);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_hist[k], (uint32_t)__popc(m));
    }
    __syncthreads();
    if (threadIdx.x < BUCKETS) tile_hist[blockIdx.x * 8 + threadIdx.x] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(128) k_bvh_split(const LevelNode* __restrict__ nodes, uint32_t n_active, const BucketStats* __restrict__ stats, Split* __restrict__ splits,
                                                   LevelNode* __restrict__ next_nodes, uint32_t* next_count, uint32_t* error, BvhInterior* __restrict__ interiors,
                                                   uint32_t* __restrict__ rank_at_pos) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_active) return;
    const LevelNode nd = nodes[j];
    const BucketStats st = stats[j];
    const SplitChoice c = choose_split(nd, st);
    Split sp;
    for (int s = 0; s < BUCKETS; ++s) sp.offset[s] = c.offset[s];
    sp.cut = c.cut;
    sp.bad = (c.n_a == 0 || c.n_b == 0) ? 1u : 0u;
    sp.pad[0] = sp.pad[1] = 0;
    sp.child[0] = sp.child[1] = NO_NODE;
    if (sp.bad) {
        atomicExch(error, 1u);
        splits[j] = sp;
        return;
    }
    // group A (the reference's first items) is the SECOND child of the flattened tree, group B the first (bvh.rs:39-50)
    LevelNode a, b;
    a.start = nd.start; a.count = c.n_a; a.rank_base = nd.rank_base + c.n_b; a.interior = nd.interior + c.n_b; a.hull = c.hull_a;
    b.start = nd.start + c.n_a; b.count = c.n_b; b.rank_base = nd.rank_base; b.interior = nd.interior + 1; b.hull = c.hull_b;
    BvhInterior in;
    for (int k = 0; k < 3; ++k) {
        in.box[0][k] = b.hull.lo[k]; in.box[0][3 + k] = b.hull.hi[k];
        in.box[1][k] = a.hull.lo[k]; in.box[1][3 + k] = a.hull.hi[k];
    }
    in.child[0] = b.count == 1 ? ~(int32_t)b.rank_base : (int32_t)b.interior;
    in.child[1] = a.count == 1 ? ~(int32_t)a.rank_base : (int32_t)a.interior;
    interiors[nd.interior] = in;
    if (a.count == 1) rank_at_pos[a.start] = a.rank_base;
    else { const uint32_t at = atomicAdd(next_count, 1u); next_nodes[at] = a; sp.child[0] = at; }
    if (b.count == 1) rank_at_pos[b.start] = b.rank_base;
    else { const uint32_t at = atomicAdd(next_count, 1u); next_nodes[at] = b; sp.child[1] = at; }
    splits[j] = sp;
}

// exclusive prefix sums of the per-tile bucket counts, in place (one block; tile counts are small next to the item count)
__global__ void __launch_bounds__(1024) k_bvh_tile_scan(uint32_t* tile_hist, uint32_t tiles) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < BUCKETS; ++k) {
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < tiles; base += 1024) {
            const uint32_t t = base + threadIdx.x;
            const uint32_t v = t < tiles ? tile_hist[t * 8 + k] : 0u;
            uint32_t x = v;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL_MASK, x, d); if ((int)lane >= d) x += y; }
            if (lane == 31) s_warp[warp] = x;
            __syncthreads();
            if (warp == 0) {
                uint32_t w = s_warp[lane];
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL_MASK, w, d); if ((int)lane >= d) w += y; }
                s_warp[lane] = w;  // inclusive over warps
            }
            __syncthreads();
            const uint32_t before = s_carry + (warp ? s_warp[warp - 1] : 0u) + (x - v);
            if (t < tiles) tile_hist[t * 8 + k] = before;
            __syncthreads();
            if (threadIdx.x == 1023) s_carry = before + v;
            __syncthreads();
        }
    }
}

// pre[p] = number of earlier positions (over the whole array) holding an item of the same bucket; the thread at the first
// position of a node also records those counts for all buckets (node_prefix), so that "earlier in the same node" is a difference
__global__ void __launch_bounds__(TILE) k_bvh_rank(const uint8_t* __restrict__ bucket, const uint32_t* __restrict__ node_of_pos, const LevelNode* __restrict__ nodes, uint32_t n,
                                                   const uint32_t* __restrict__ tile_prefix, uint32_t* __restrict__ pre, uint32_t* __restrict__ node_prefix) {
    __shared__ uint32_t s_warp[BUCKETS][32];
    const uint32_t p = blockIdx.x * TILE + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t b = p < n ? bucket[p] : BUCKET_NONE;
    uint32_t below[BUCKETS];
#pragma unroll
    for (int k = 0; k < BUCKETS; ++k) {
        const unsigned m = __ballot_sync(FULL_MASK, b == (uint32_t)k);
        below[k] = (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[k][warp] = (uint32_t)__popc(m);
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < BUCKETS; ++k) {
            const uint32_t v = s_warp[k][lane];
            uint32_t x = v;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL_MASK, x, d); if ((int)lane >= d) x += y; }
            s_warp[k][lane] = x - v;  // exclusive over the warps of the tile
        }
    }
    __syncthreads();
    if (b == BUCKET_NONE) return;
    const uint32_t j = node_of_pos[p];
    const bool first_of_node = nodes[j].start == p;
#pragma unroll
    for (int k = 0; k < BUCKETS; ++k) {
        const uint32_t s = tile_prefix[blockIdx.x * 8 + k] + s_warp[k][warp] + below[k];
        if (b == (uint32_t)k) pre[p] = s;
        if (first_of_node) node_prefix[j * 8 + k] = s;
    }
}

__global__ void __launch_bounds__(TILE) k_bvh_scatter(const uint8_t* __restrict__ bucket, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ node_of_pos,
                                                      const LevelNode* __restrict__ nodes, const Split* __restrict__ splits, const uint32_t* __restrict__ pre,
                                                      const uint32_t* __restrict__ node_prefix, uint32_t n, uint32_t* __restrict__ ids_next, uint32_t* __restrict__ node_next) {
    const uint32_t p = blockIdx.x * TILE + threadIdx.x;
    if (p >= n) return;
    const uint32_t b = bucket[p];
    if (b == BUCKET_NONE) { ids_next[p] = ids[p]; node_next[p] = NO_NODE; return; }
    const uint32_t j = node_of_pos[p];
    const Split sp = splits[j];
    const uint32_t to = nodes[j].start + sp.offset[b] + (pre[p] - node_prefix[j * 8 + b]);
    ids_next[to] = ids[p];
    node_next[to] = sp.child[b < sp.cut ? 0 : 1];
}

__global__ void k_bvh_order(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ rank_at_pos, uint32_t n, uint32_t* __restrict__ order) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) order[rank_at_pos[p]] = ids[p];
}

__global__ void k_bvh_init(uint32_t* ids, uint32_t* node_of_pos, uint32_t n) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) { ids[p] = p; node_of_pos[p] = 0u; }
}

struct Scratch {  // one device block from the caller (kept by the context: cudaFree of it cost 30 - 240 ms on a B200), carved up
    char* base = nullptr;
    size_t used = 0;
    template <class T> size_t plan(size_t elements) {  // first pass: sizes -> offsets
        const size_t at = used;
        used += (elements * sizeof(T) + 255) / 256 * 256;
        return at;
    }
    template <class T> T* at(size_t offset) const { return reinterpret_cast<T*>(base + offset); }
};

}  // namespace

void gpu_bvh_build(const float* boxes6, size_t n_items, const float* root_hull12, BvhTree& out, cudaStream_t stream, const std::function<void*(size_t)>& device_scratch) {
    if (n_items < 2) throw std::runtime_error("GPU BVH build: needs at least two items");
    if (n_items > 0x7fffffffull) throw ir::BuildError("too many BVH items");
    const bool timing = getenv("PYR_BUILD_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gpu bvh build] %-24s %.4f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    const uint32_t n = (uint32_t)n_items;
    const uint32_t tiles = (n + TILE - 1) / TILE;
    const size_t node_cap = (size_t)n / 2 + 1;  // a level has at most n / 2 nodes of two items or more
    Scratch s;
    const size_t o_boxes = s.plan<Box6>(n), o_ids0 = s.plan<uint32_t>(n), o_ids1 = s.plan<uint32_t>(n), o_node0 = s.plan<uint32_t>(n), o_node1 = s.plan<uint32_t>(n);
    const size_t o_bucket = s.plan<uint8_t>(n), o_pre = s.plan<uint32_t>(n), o_rank_at = s.plan<uint32_t>(n), o_order = s.plan<uint32_t>(n);
    const size_t o_nodes0 = s.plan<LevelNode>(node_cap), o_nodes1 = s.plan<LevelNode>(node_cap), o_stats = s.plan<BucketStats>(node_cap);
    const size_t o_splits = s.plan<Split>(node_cap), o_node_prefix = s.plan<uint32_t>(node_cap * 8), o_tile_hist = s.plan<uint32_t>((size_t)tiles * 8);
    const size_t o_interiors = s.plan<BvhInterior>((size_t)n - 1), o_flags = s.plan<uint32_t>(4);
    s.base = static_cast<char*>(device_scratch(s.used));
    if (!s.base) throw std::runtime_error("GPU BVH build: no scratch memory");
    Box6* d_boxes = s.at<Box6>(o_boxes);
    uint32_t* d_ids[2] = {s.at<uint32_t>(o_ids0), s.at<uint32_t>(o_ids1)};
    uint32_t* d_node[2] = {s.at<uint32_t>(o_node0), s.at<uint32_t>(o_node1)};
    uint8_t* d_bucket = s.at<uint8_t>(o_bucket);
    uint32_t* d_pre = s.at<uint32_t>(o_pre);
    uint32_t* d_rank_at = s.at<uint32_t>(o_rank_at);
    uint32_t* d_order = s.at<uint32_t>(o_order);
    LevelNode* d_nodes[2] = {s.at<LevelNode>(o_nodes0), s.at<LevelNode>(o_nodes1)};
    BucketStats* d_stats = s.at<BucketStats>(o_stats);
    Split* d_splits = s.at<Split>(o_splits);
    uint32_t* d_node_prefix = s.at<uint32_t>(o_node_prefix);
    uint32_t* d_tile_hist = s.at<uint32_t>(o_tile_hist);
    BvhInterior* d_interiors = s.at<BvhInterior>(o_interiors);
    uint32_t* d_flags = s.at<uint32_t>(o_flags);  // [0] nodes of the next level, [1] error
    uint32_t h_flags[2] = {0, 0};
    lap("allocations");
    BV(cudaMemcpyAsync(d_boxes, boxes6, (size_t)n * sizeof(Box6), cudaMemcpyHostToDevice, stream));
    LevelNode root;
    root.start = 0; root.count = n; root.rank_base = 0; root.interior = 0;
    for (int k = 0; k < 3; ++k) {
        root.hull.lo[k] = root_hull12[k]; root.hull.hi[k] = root_hull12[3 + k];
        root.hull.c_lo[k] = root_hull12[6 + k]; root.hull.c_hi[k] = root_hull12[9 + k];
    }
    BV(cudaMemcpyAsync(d_nodes[0], &root, sizeof(root), cudaMemcpyHostToDevice, stream));
    BV(cudaMemsetAsync(d_flags, 0, 4 * sizeof(uint32_t), stream));
    k_bvh_init<<<(n + 255) / 256, 256, 0, stream>>>(d_ids[0], d_node[0], n);

    if (timing) { BV(cudaStreamSynchronize(stream)); lap("upload of the boxes"); }
    uint32_t n_active = 1;
    int level = 0, cur = 0;
    while (n_active) {
        if (level >= 4096) throw ir::BuildError("the BVH does not finish (more than 4096 levels)");
        BV(cudaMemsetAsync(d_stats, 0, (size_t)n_active * sizeof(BucketStats), stream));
        BV(cudaMemsetAsync(d_flags, 0, sizeof(uint32_t), stream));
        k_bvh_assign<<<tiles, TILE, 0, stream>>>(d_boxes, d_ids[cur], d_node[cur], d_nodes[cur], n, d_bucket, d_stats, d_tile_hist);
        k_bvh_split<<<(n_active + 127) / 128, 128, 0, stream>>>(d_nodes[cur], n_active, d_stats, d_splits, d_nodes[cur ^ 1], d_flags, d_flags + 1, d_interiors, d_rank_at);
        k_bvh_tile_scan<<<1, 1024, 0, stream>>>(d_tile_hist, tiles);
        k_bvh_rank<<<tiles, TILE, 0, stream>>>(d_bucket, d_node[cur], d_nodes[cur], n, d_tile_hist, d_pre, d_node_prefix);
        k_bvh_scatter<<<tiles, TILE, 0, stream>>>(d_bucket, d_ids[cur], d_node[cur], d_nodes[cur], d_splits, d_pre, d_node_prefix, n, d_ids[cur ^ 1], d_node[cur ^ 1]);
        BV(cudaMemcpyAsync(h_flags, d_flags, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));  // (8 bytes: the level count decides whether to go on)
        BV(cudaStreamSynchronize(stream));
        BV(cudaGetLastError());
        if (h_flags[1]) throw ir::BuildError("BVH split produced an empty side");
        n_active = h_flags[0];
        if (n_active > node_cap) throw std::runtime_error("GPU BVH build: node count out of range");
        cur ^= 1;
        ++level;
    }
    lap("levels");
    k_bvh_order<<<(n + 255) / 256, 256, 0, stream>>>(d_ids[cur], d_rank_at, n, d_order);
    out.order.resize(n);
    out.allocate_interiors((size_t)n - 1);
    BV(cudaMemcpyAsync(out.order.data(), d_order, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    BV(cudaMemcpyAsync(out.interiors.get(), d_interiors, ((size_t)n - 1) * sizeof(BvhInterior), cudaMemcpyDeviceToHost, stream));
    BV(cudaStreamSynchronize(stream));
    BV(cudaGetLastError());
    lap("download");
    out.root = 0;
    out.max_depth = level;
}

}  // namespace pyr
