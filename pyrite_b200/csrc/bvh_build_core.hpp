// Bvh::new (pyrite/src/spatial/bvh.rs:13-155) as a LEVEL-SYNCHRONOUS build: every level of the tree is split at once, one
// thread per item and one per node, instead of the reference's depth-first loop over a work stack.  The tree comes out
// identical - same item groups in the same order, same boxes, same leaf pre-order - because nothing in the reference's
// split depends on the order it works in:
//   * an item's bucket is a function of its own box and its node's hull (bvh.rs:93-103);
//   * bucket hulls and counts are unions / sums: minima, maxima and integer adds are exact in any order;
//   * the split choice reads the six bucket records only (bvh.rs:112-127) and is evaluated here with the reference's
//     operations in the reference's order, one thread per node;
//   * the two item groups keep the bucket-major, otherwise original order of `merge_buckets` (bvh.rs:129-131): a STABLE
//     partition by bucket index, done with prefix counts;
//   * the flattened pre-order (the `first` child is the subtree of the SECOND group, because the Join pops it first,
//     bvh.rs:39-50) only needs the sizes of the groups: a node with `rank_base` leaves before it and interior index `i` has
//     the second group's subtree at ranks [rank_base, rank_base + nB) / interior i + 1 and the first group's at
//     [rank_base + nB, ...) / interior i + nB.
// This header is the per-item and per-node logic, __host__ __device__: bvh_build.cu runs it in kernels on the GPU, and
// tests/host_emu.cpp runs the same functions level by level on the CPU against the depth-first host builder.
#pragma once
#include <stdint.h>

#include "device_types.h"

namespace pyr {
namespace bvhb {

constexpr int BUCKETS = 6;                     // bvh.rs:91
constexpr uint32_t NO_NODE = 0xFFFFFFFFu;      // a position whose item already is a leaf
constexpr uint32_t BUCKET_NONE = 0xFFu;

struct Hull {  // spatial/bvh.rs:318-370: the boxes' union and the bounds of their centres
    float lo[3], hi[3], c_lo[3], c_hi[3];
};

// An interior node of the current level: items [start, start + count) of the position array, count >= 2.
struct LevelNode {
    uint32_t start, count;
    uint32_t rank_base;   // leaves before this subtree in the flattened pre-order
    uint32_t interior;    // index of this node among the interior nodes (pre-order, like the depth-first builder numbers them)
    Hull hull;
};

// Order-preserving map float -> unsigned (so that integer atomics take minima and maxima of floats).  All slots are
// combined with MAX: the lower bounds are stored complemented.  0 = nothing recorded yet.
PYR_HD uint32_t order_key(float f) {
    uint32_t b;
#if defined(__CUDA_ARCH__)
    b = __float_as_uint(f);
#else
    __builtin_memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
PYR_HD float order_value(uint32_t k) {
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    __builtin_memcpy(&f, &b, 4);
    return f;
#endif
}

// What one level gathers per node and bucket: item count and the hull as 12 max-combined keys
// (0..2 ~lo, 3..5 hi, 6..8 ~c_lo, 9..11 c_hi).
struct BucketStats {
    uint32_t count[BUCKETS];
    uint32_t key[BUCKETS][12];
    uint32_t pad[2];
};
static_assert(sizeof(BucketStats) == 320, "BucketStats is 80 words");

// The outcome of a node's split, as the scatter pass needs it.
struct Split {
    uint32_t offset[BUCKETS];  // first position (relative to the node's start) of every bucket's items after the partition
    uint32_t cut;              // buckets [0, cut) form the first group (A), the rest the second (B)
    uint32_t child[2];         // next-level node index of group A / B, NO_NODE when the group is a single item (a leaf)
    uint32_t bad;              // 1: a side came out empty (the reference panics there, bvh.rs:139-146)
    uint32_t pad[2];
};

PYR_HD float box_middle(float lo, float hi) { return lo + (hi - lo) / 2.0f; }  // Aabb::center as host_math.hpp's Box::middle

// Hull::largest_axis (bvh.rs:344-358) on the bounds of the centres
PYR_HD void largest_axis(const Hull& h, float& width, int& axis) {
    const float dx = h.c_hi[0] - h.c_lo[0], dy = h.c_hi[1] - h.c_lo[1], dz = h.c_hi[2] - h.c_lo[2];
    if (dy > dx) { width = dy; axis = 1; } else { width = dx; axis = 0; }
    if (dz > width) { width = dz; axis = 2; }
}

// Rust's saturating `as usize`, then `.min(BUCKETS - 1)`
PYR_HD uint32_t bucket_index(float fi) {
    if (!(fi > 0.0f)) return 0u;
    if (fi >= (float)BUCKETS) return (uint32_t)(BUCKETS - 1);
    return (uint32_t)fi;
}

// The bucket of the item at position `p` (box lo/hi = box6[0..2] / [3..5]) in its node (bvh.rs:66-68, 93-103).  Nodes whose
// centres (nearly) coincide are halved by position: "bucket" 0 = first half, 1 = second half.
PYR_HD uint32_t item_bucket(const LevelNode& nd, uint32_t p, const float* box6) {
    float width; int axis;
    largest_axis(nd.hull, width, axis);
    if (width < DIST_EPSILON) return (p - nd.start) < nd.count / 2u ? 0u : 1u;
    const float where = box_middle(box6[axis], box6[3 + axis]);
    const float fi = (float)BUCKETS * (where - nd.hull.c_lo[axis]) / width;
    return bucket_index(fi);
}

// The twelve keys an item adds to its bucket's record
PYR_HD void item_keys(const float* box6, uint32_t* key12) {
    for (int a = 0; a < 3; ++a) {
        const float c = box_middle(box6[a], box6[3 + a]);
        key12[a] = ~order_key(box6[a]);
        key12[3 + a] = order_key(box6[3 + a]);
        key12[6 + a] = ~order_key(c);
        key12[9 + a] = order_key(c);
    }
}

PYR_HD Hull hull_of_keys(const uint32_t* key12) {
    Hull h;
    for (int a = 0; a < 3; ++a) {
        h.lo[a] = order_value(~key12[a]);
        h.hi[a] = order_value(key12[3 + a]);
        h.c_lo[a] = order_value(~key12[6 + a]);
        h.c_hi[a] = order_value(key12[9 + a]);
    }
    return h;
}
PYR_HD Hull hull_join(const Hull& a, const Hull& b) {
    Hull h;
    for (int k = 0; k < 3; ++k) {
        h.lo[k] = a.lo[k] < b.lo[k] ? a.lo[k] : b.lo[k];
        h.hi[k] = a.hi[k] > b.hi[k] ? a.hi[k] : b.hi[k];
        h.c_lo[k] = a.c_lo[k] < b.c_lo[k] ? a.c_lo[k] : b.c_lo[k];
        h.c_hi[k] = a.c_hi[k] > b.c_hi[k] ? a.c_hi[k] : b.c_hi[k];
    }
    return h;
}
PYR_HD float hull_area(const Hull& h) {  // Aabb3::surface_area
    const float dx = h.hi[0] - h.lo[0], dy = h.hi[1] - h.lo[1], dz = h.hi[2] - h.lo[2];
    return 2.0f * ((dx * dy) + (dx * dz) + (dy * dz));
}

// get_bucket_stats (bvh.rs:277-296) over buckets [from, to)
PYR_HD void tally(const BucketStats& st, int from, int to, uint32_t& count, float& area) {
    count = 0;
    bool any = false;
    Hull acc;
    for (int s = from; s < to; ++s) {
        if (!st.count[s]) continue;
        const Hull h = hull_of_keys(st.key[s]);
        acc = any ? hull_join(acc, h) : h;
        any = true;
        count += st.count[s];
    }
    area = any ? hull_area(acc) : 0.0f;
}

// The split of one node from its bucket records (bvh.rs:66-88 for coinciding centres, :105-147 otherwise): where the cut goes,
// where every bucket's items land, the two groups' sizes and hulls.
struct SplitChoice { uint32_t cut, n_a, n_b; Hull hull_a, hull_b; uint32_t offset[BUCKETS]; };
PYR_HD SplitChoice choose_split(const LevelNode& nd, const BucketStats& st) {
    SplitChoice c;
    float width; int axis;
    largest_axis(nd.hull, width, axis);
    if (width < DIST_EPSILON) {
        c.cut = 1;  // the two halves were recorded as buckets 0 and 1
    } else {
        float best = order_value(0xFF800000u);  // +inf (f32::INFINITY, bvh.rs:110)
        c.cut = 0;
        const float whole = hull_area(nd.hull);
        for (int s = 1; s < BUCKETS; ++s) {
            uint32_t n1, n2; float a1, a2;
            tally(st, 0, s, n1, a1);
            tally(st, s, BUCKETS, n2, a2);
            const float cost = (a1 * (float)n1 + a2 * (float)n2) / whole;
            if (cost < best) { best = cost; c.cut = (uint32_t)s; }
        }
    }
    uint32_t at = 0;
    c.n_a = 0; c.n_b = 0;
    bool any_a = false, any_b = false;
    for (int s = 0; s < BUCKETS; ++s) {
        c.offset[s] = at;
        at += st.count[s];
        if (!st.count[s]) continue;
        const Hull h = hull_of_keys(st.key[s]);
        if ((uint32_t)s < c.cut) { c.hull_a = any_a ? hull_join(h, c.hull_a) : h; any_a = true; c.n_a += st.count[s]; }
        else { c.hull_b = any_b ? hull_join(h, c.hull_b) : h; any_b = true; c.n_b += st.count[s]; }
    }
    return c;
}

}  // namespace bvhb
}  // namespace pyr
